"""Stand-in for torchrl (generator-only): only the ReplayBuffer type name is needed."""
