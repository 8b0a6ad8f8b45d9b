"""Independent NumPy statement of the device RNG contract (include/r6dof.h, r6_reset / r6_rollout):
Philox4x32-10, key = seed, counter = (global env id lo, hi, a, b).  Used to check the CUDA reset
sampler and the synthetic action stream bit-for-bit."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
STREAM_RESET = 0x52534554
STREAM_ACTION = 0x41435431


def philox4x32_10(ctr, key):
    """ctr: uint32 [..., 4], key: (k0, k1) -> uint32 [..., 4]"""
    c = np.array(ctr, dtype=np.uint32, copy=True)
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[..., 0].astype(np.uint64)
            p1 = M1 * c[..., 2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            n = np.empty_like(c)
            n[..., 0] = hi1 ^ c[..., 1] ^ k0
            n[..., 1] = lo1
            n[..., 2] = hi0 ^ c[..., 3] ^ k1
            n[..., 3] = lo0
            c = n
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c


def u53(a, b):
    return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0


def sample_ic(ic_low, ic_high, seed, genv, episode):
    """Box.sample for global env ids `genv` (int64 array) and episode counters -> float32 [n,14]
    BEFORE the quaternion normalisation."""
    genv = np.asarray(genv, np.uint64)
    episode = np.broadcast_to(np.asarray(episode, np.uint32), genv.shape)
    lo, hi = np.asarray(ic_low, np.float32).astype(np.float64), np.asarray(ic_high, np.float32).astype(np.float64)
    out = np.zeros(genv.shape + (14,), np.float32)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    for b in range(7):
        ctr = np.stack([(genv & np.uint64(0xFFFFFFFF)).astype(np.uint32), (genv >> np.uint64(32)).astype(np.uint32),
                        episode, np.full(genv.shape, STREAM_RESET + b, np.uint32)], -1)
        r = philox4x32_10(ctr, key)
        ua, ub = u53(r[..., 0], r[..., 1]), u53(r[..., 2], r[..., 3])
        i = 2 * b
        out[..., i] = (lo[i] + (hi[i] - lo[i]) * ua).astype(np.float32)
        out[..., i + 1] = (lo[i + 1] + (hi[i + 1] - lo[i + 1]) * ub).astype(np.float32)
    return out


def normalize_ic_quaternion(ic):
    """rocket_env.py:190 in float32 (sdot rule: f32 products, f64 accumulation, f32 rounding)."""
    ic = np.array(ic, np.float32, copy=True)
    q = ic[..., 6:10]
    acc = (q * q).astype(np.float64).sum(-1)
    n = np.sqrt(acc.astype(np.float32))
    ic[..., 6:10] = q / n[..., None]
    return ic


def actions(seed, genv, step):
    """uniform(-1,1) float32 actions of r6_rollout's R6_ACT_PHILOX for (global env, global step)."""
    genv = np.asarray(genv, np.uint64)
    step = np.broadcast_to(np.asarray(step, np.uint64), genv.shape)
    ctr = np.stack([(genv & np.uint64(0xFFFFFFFF)).astype(np.uint32), (genv >> np.uint64(32)).astype(np.uint32),
                    (step & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                    np.uint32(STREAM_ACTION) ^ (step >> np.uint64(32)).astype(np.uint32)], -1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return (-1.0 + 2.0 * ((r[..., :3].astype(np.float64) + 0.5) / 4294967296.0)).astype(np.float32)
