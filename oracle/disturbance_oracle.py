"""CPU restatement (numpy) of the device-RNG disturbance draw (csrc/disturbance_core.cuh, agx_disturbance_draw).

TEST INFRASTRUCTURE ONLY.  The DISTRIBUTION follows BaseMultirotor.apply_disturbance (robots/base_multirotor.py:213-234: a Bernoulli(p)
gate times U(-max, max) force and torque); the STREAM has no counterpart in the reference (torch.bernoulli + rand_like there; that
path is EnvManager._draw_disturbance with reset_rng='torch').  Philox4x32-10 (oracle/philox.py), counter = (global env id, draw counter,
block 0|1, 'DIST'), key = seed; gate = word 0 of block 0, the six uniforms = words 1..3 of block 0 and words 0..2 of block 1."""
import numpy as np

from . import philox

TAG = 0x44495354


def draw(num_envs, env_id_offset, prob, max6, seed, counter):
    gid = (np.arange(num_envs, dtype=np.uint64) + np.uint64(env_id_offset)).astype(np.uint32)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)

    def block(b):
        ctr = np.stack([gid, np.full(num_envs, counter, np.uint32), np.full(num_envs, b, np.uint32), np.full(num_envs, TAG, np.uint32)], axis=-1)
        return philox.u01(philox.philox4x32_10(ctr, key))
    a, b = block(0), block(1)
    gate = (a[:, 0] < np.float32(prob)).astype(np.float32)
    u = np.concatenate([a[:, 1:4], b[:, 0:3]], axis=1)
    mx = np.asarray(max6, np.float32)
    return (((mx - (-mx)) * u + (-mx)) * gate[:, None]).astype(np.float32)
