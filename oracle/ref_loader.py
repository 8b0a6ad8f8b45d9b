"""Load the UNMODIFIED reference from /root/reference — TEST INFRASTRUCTURE ONLY.

Only `oracle/make_golden.py` and the optional `-m "not gpu"` cross-checks use this, and only in
the build container: /root/reference does not exist on the GPU box, so nothing on the product
path, in the `-m gpu` tests, in `smoke()` or in `bench.py` may import this module.
"""
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("R6_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "my_environment", "envs", "rocket_env.py"))


def load_reference():
    """Returns (Rocket6DOF, Simulator6DOF, env_config, sb3_config) from the reference tree."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import yaml
    from my_environment.envs.rocket_env import Rocket6DOF  # noqa: E402
    from my_environment.utils.simulator import Simulator6DOF  # noqa: E402

    with open(os.path.join(REFERENCE_ROOT, "config.yaml")) as f:
        cfg = yaml.safe_load(f)
    return Rocket6DOF, Simulator6DOF, cfg["env_config"], cfg["sb3_config"]
