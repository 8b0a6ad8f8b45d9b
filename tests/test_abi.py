"""The C-ABI library loads without a GPU, exports every entry point include/r6dof.h declares, and its
struct layouts match the ctypes mirrors.  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "r6dof.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(r6_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rl_rocket_6dof_b200 import _lib, build
    path = build.build()
    L = C.CDLL(path)
    names = declared_functions()
    assert "r6_step" in names and "r6_reset" in names and "r6_rollout" in names
    for n in names:
        assert hasattr(L, n), f"{n} declared in r6dof.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "ctypes binding and header disagree on the entry points"


def test_abi_version_and_struct_layouts():
    from rl_rocket_6dof_b200 import _lib
    from rl_rocket_6dof_b200.params import R6Params
    L = _lib.load()
    hdr = open(HEADER).read()
    assert int(re.search(r"#define R6_ABI_VERSION (\d+)", hdr).group(1)) == L.r6_abi_version() == _lib.ABI_VERSION
    assert L.r6_params_size() == C.sizeof(R6Params)
    assert L.r6_buffers_size() == C.sizeof(_lib.R6Buffers)
    assert L.r6_last_error() is not None


def test_bad_arguments_are_rejected_without_a_gpu():
    """Argument validation happens before any CUDA call: error code + message, never a crash."""
    from rl_rocket_6dof_b200 import _lib
    L = _lib.load()
    assert L.r6_step(None, None, 4, 0, None, 0, None) == -1
    assert b"null" in L.r6_last_error()
    assert L.r6_tgo(None, None, None, 1.0, 4, None, None, None) == -1
    assert L.r6_stats_reset(None, None) == -1


def test_missing_library_fails_loudly(tmp_path):
    from rl_rocket_6dof_b200 import _lib
    try:
        _lib.load(str(tmp_path / "nope.so"))
    except _lib.R6Error as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("a missing libr6dof.so must raise")


def test_params_match_reference_constants():
    """derive_params reproduces the constants dumped from Rocket6DOF.__init__ (tests/golden/constants.npz)."""
    from rl_rocket_6dof_b200.params import derive_params, load_config
    sb3, cfg = load_config()
    ep = derive_params(cfg, sb3)
    g = np.load(os.path.join(ROOT, "tests", "golden", "constants.npz"), allow_pickle=True)
    assert np.array_equal(ep.state_normalizer, g["state_normalizer"])
    for key, mine in (("ic_low", ep.ic_low), ("ic_high", ep.ic_high)):
        if key in g.files:
            assert np.array_equal(mine, g[key])


def test_build_is_keyed_on_a_source_hash(tmp_path, monkeypatch):
    """build() decides staleness by a hash of csrc/*.cu*, include/*.h and the flags (not by mtimes), and its
    dependency list covers every header the library includes."""
    from rl_rocket_6dof_b200 import build as b
    names = {os.path.basename(d) for d in b.deps()}
    assert {"r6_kernels.cu", "r6_core.cuh", "r6_mlp_tc.cuh", "r6_mlp_tcgen05.cuh", "r6dof.h"} <= names
    b.build()
    assert not b.needs_build() and b.built_hash() == b.source_hash()
    h0 = b.source_hash()
    monkeypatch.setattr(b, "NVCC_FLAGS", b.NVCC_FLAGS + ["-DSOMETHING"])
    assert b.source_hash() != h0 and b.needs_build()


def test_integration_stub_matches_the_library():
    """The ctypes stub printed in INTEGRATION.md §4: its ABI number and the r6_step / r6_step_range signatures are the
    library's (the snippet had drifted to an old ABI number)."""
    from rl_rocket_6dof_b200 import _lib
    L = _lib.load()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"assert lib\.r6_abi_version\(\) == (\d+)", text)
    assert m and int(m.group(1)) == L.r6_abi_version() == _lib.ABI_VERSION
    assert L.r6_params_size() == C.sizeof(_lib.R6Params) and L.r6_buffers_size() == C.sizeof(_lib.R6Buffers)
    hdr = open(os.path.join(ROOT, "include", "r6dof.h")).read()
    sig = re.search(r"int r6_step_range\((.*?)\);", hdr, re.S).group(1)
    assert len(sig.split(",")) == 11            # p, b, n, first, count, lane, env_offset, actions, seed, step_index, stream
    assert "r6_step_range(p, b, n, first, count, lane, env_offset, actions, seed, step_index, stream)" in text
