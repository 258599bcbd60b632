"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) in numpy + the device-RNG draw layout.

TEST INFRASTRUCTURE (see oracle/hp1_oracle.py header).  Restates the spec the CUDA kernel
implements in aerial_gym_simulator_b200/csrc/agx_math.cuh (philox4x32_10) and hp1.cu
(device_rng_reset).  Known-answer pinned against the Random123 test vectors in
tests/test_philox.py.  This RNG has NO counterpart in the reference (which draws with
torch.rand_like on reset); it exists so the fused single-launch step can reset envs without a
host round trip.  The reference-order torch draws remain available (agx_hp1_reset with draws).

Layout: counter = (env_gid, episode, block, 0), key = (seed & 0xffffffff, seed >> 32),
u = (x >> 8) * 2^-24.
  block 0..2 -> state[0..11]; block 3 -> state[12], bounds_lo[0..2]; block 4 -> bounds_hi[0..2], -
  block 5..8 -> K_pos, K_vel, K_rot, K_angvel (xyz, -); block 9+i -> motor i: tau_inc, tau_dec, thrust, k
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: uint32 [...,4]; key: (k0, k1) uint32 scalars.  Returns uint32 [...,4]."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0]
            p1 = M1 * c[2]
            hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
            n0 = hi1 ^ c[1] ^ np.uint64(k0)
            n2 = hi0 ^ c[3] ^ np.uint64(k1)
            c = [n0, lo1, n2, lo0]
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def u01(x):
    return ((x >> np.uint32(8)).astype(np.float32)) * np.float32(1.0 / 16777216.0)


def reset_uniforms(seed: int, env_gid, episode, num_motors: int):
    """numpy dict of uniform draws for envs with global ids ``env_gid`` at episode counters
    ``episode`` (both int arrays [n]).  Keys match AgxHp1ResetDraws / oracle ResetDraws."""
    env_gid = np.asarray(env_gid, dtype=np.uint32)
    episode = np.asarray(episode, dtype=np.uint32)
    n = env_gid.shape[0]
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)

    def block(b):
        ctr = np.stack([env_gid, episode, np.full(n, b, np.uint32), np.zeros(n, np.uint32)], axis=-1)
        return u01(philox4x32_10(ctr, key))

    b0, b1, b2, b3, b4 = block(0), block(1), block(2), block(3), block(4)
    state = np.concatenate([b0, b1, b2, b3[:, :1]], axis=1)
    out = {
        "state": state,
        "bounds_lo": b3[:, 1:4].copy(),
        "bounds_hi": b4[:, 0:3].copy(),
        "K_pos": block(5)[:, :3].copy(),
        "K_vel": block(6)[:, :3].copy(),
        "K_rot": block(7)[:, :3].copy(),
        "K_angvel": block(8)[:, :3].copy(),
    }
    mot = np.stack([block(9 + i) for i in range(num_motors)], axis=1)  # [n, M, 4]
    out["tau_inc"] = mot[:, :, 0].copy()
    out["tau_dec"] = mot[:, :, 1].copy()
    out["thrust"] = mot[:, :, 2].copy()
    out["k_thrust"] = mot[:, :, 3].copy()
    return out
