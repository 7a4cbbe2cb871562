class ReplayBuffer:  # agents/agent.py:15 uses it as a type annotation only
    pass
