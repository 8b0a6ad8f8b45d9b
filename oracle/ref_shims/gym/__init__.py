"""Minimal stand-in for gym 0.21 — TEST INFRASTRUCTURE ONLY.

The reference (`/root/reference/my_environment/envs/rocket_env.py:9`) imports
`gym.Env` and `gym.spaces`; gym is not installed in this image.  This shim gives the
reference just enough surface to run UNMODIFIED inside `oracle/make_golden.py`.
It follows gym 0.21's `Box` semantics (float32 default dtype, float64 uniform draw then
cast, `contains` = castable & shape & low<=x<=high).  The RNG stream is NOT gym's
(gym hashes the seed); parity runs inject the sampled initial conditions, so stream
equality is not needed.
"""
from . import spaces  # noqa: F401


class Env:
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    def seed(self, seed=None):
        return [seed]

    def close(self):
        return None

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kw):
        return self.env.reset(**kw)

    @property
    def unwrapped(self):
        return self.env.unwrapped


class ObservationWrapper(Wrapper):
    def reset(self, **kw):
        return self.observation(self.env.reset(**kw))

    def step(self, action):
        o, r, d, i = self.env.step(action)
        return self.observation(o), r, d, i


class RewardWrapper(Wrapper):
    def step(self, action):
        o, r, d, i = self.env.step(action)
        return o, self.reward(r), d, i


def make(id, **kwargs):
    """gym.make for ids registered through gym.envs.registration.register (entry_point 'module:Class')."""
    import importlib
    from .envs.registration import REGISTRY
    entry = REGISTRY[id]
    if isinstance(entry, str):
        mod, _, name = entry.partition(":")
        entry = getattr(importlib.import_module(mod), name)
    return entry(**kwargs)


from . import wrappers  # noqa: E402,F401
