"""Hyper-parameter container with the key set of the reference's flat YAML configs
(tasks/defaults/sac.yml, tasks/defaults/td3.yml). ``Agent`` accepts this class, an
``omegaconf.DictConfig`` (when installed) or any object/mapping with the same keys."""
from __future__ import annotations

from typing import Any, Mapping

_SHARED = {
    # resources
    "cuda": True, "compile": False, "cudagraphs": True,
    # env
    "sync_vec_env": True, "num_envs": 4, "action_repeat": 1, "capture_video": False,
    "normalize_observations": False,
    # logging
    "wandb_project": "calico", "measure_burnin": 3,
    # training / evaluation mode
    "num_timesteps": 10_000_000, "eval_steps": 10, "eval_every": 10_000,
    "gather_trajectories": False, "pixels_too": False,
    # model / optimisation
    "layer_norm": True, "actor_lr": 3e-4, "clip_norm": 0.0,
    # algorithm
    "segment_len": 1, "batch_size": 256, "gamma": 0.99, "rb_capacity": 1_000_000, "polyak": 0.005,
    "actor_update_delay": 2,
}
SAC_DEFAULTS = dict(_SHARED, learning_starts=5000, num_episodes=16, qnets_lr=1e-3, prefer_td3_over_sac=False,
                    bcq_style_targ_mix=False, crit_targ_update_freq=1, alpha_init=0.2, autotune=True,
                    log_alpha_lr=1e-3)
TD3_DEFAULTS = dict(_SHARED, learning_starts=25000, num_episodes=10, qnets_lr=3e-4, prefer_td3_over_sac=True,
                    bcq_style_targ_mix=True, actor_noise_std=0.1, targ_actor_smoothing=True, td3_std=0.2,
                    td3_c=0.5)


class Hps(dict):
    """dict with attribute access; read-only once frozen (the reference freezes its cfg, main.py:107)."""

    _frozen = False

    def __getattr__(self, k: str) -> Any:
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(f"config has no key {k!r}") from e

    def __setattr__(self, k, v):
        if k == "_frozen":
            object.__setattr__(self, k, v)
        elif self._frozen:
            raise TypeError("config is read-only")
        else:
            self[k] = v

    def freeze(self) -> "Hps":
        self._frozen = True
        return self


def sac_hps(**over) -> Hps:
    return Hps(dict(SAC_DEFAULTS, **over))


def td3_hps(**over) -> Hps:
    return Hps(dict(TD3_DEFAULTS, **over))


def load_hps(path, **over) -> Hps:
    """Read one of the reference's YAML files (or any file with the same keys)."""
    import yaml
    with open(path) as f:
        return Hps(dict(yaml.safe_load(f), **over))


def hp_get(hps, key: str, default=None):
    """Read a key from a namespace-like or mapping-like config without touching absent keys."""
    if isinstance(hps, Mapping):
        return hps.get(key, default)
    return getattr(hps, key, default)
