"""Device-RNG disturbance draw on CPU: the device code of agx_disturbance_draw compiled for the host against the numpy oracle, its
distribution against the reference's (Bernoulli gate x uniform wrench), and EnvManager's use of it through the CPU twin of the stack."""
import ctypes as C

import numpy as np
import torch

from oracle import disturbance_oracle as D

from . import _shadow
from ._cpu_stack import cpu_stack


def _draw(n, off, prob, mx, seed, counter):
    out = np.zeros((n, 6), np.float32)
    m = (C.c_float * 6)(*mx)
    assert _shadow.load().shadow_disturbance_draw(n, off, prob, C.cast(m, C.c_void_p), seed, counter, out.ctypes.data_as(C.c_void_p)) == 0
    return out


def test_shadow_matches_oracle_and_is_sharding_invariant():
    mx = [4.75, 4.75, 4.75, 0.03, 0.03, 0.03]
    seed = 0xFEED_0000_1234
    a = _draw(1000, 64, 0.05, mx, seed, 7)
    assert np.array_equal(a, D.draw(1000, 64, 0.05, mx, seed, 7))
    assert np.array_equal(a[500:], _draw(500, 564, 0.05, mx, seed, 7))         # keyed by the GLOBAL env id
    assert not np.array_equal(a, _draw(1000, 64, 0.05, mx, seed, 8)) and not np.array_equal(a, _draw(1000, 64, 0.05, mx, seed + 1, 7))


def test_distribution():
    mx = [1.5, 1.5, 1.5, 0.25, 0.25, 0.25]
    a = _draw(400_000, 0, 0.05, mx, 3, 0)
    hit = (a != 0).any(axis=1)
    assert abs(hit.mean() - 0.05) < 0.002 and (a[~hit] == 0).all()
    h = a[hit]
    assert (np.abs(h) <= np.asarray(mx, np.float32)).all()
    assert np.abs(h.mean(axis=0) / np.asarray(mx)).max() < 0.02                 # centred
    assert np.abs(h.std(axis=0) / (np.asarray(mx) / np.sqrt(3.0)) - 1).max() < 0.02  # uniform on [-max, max]
    assert np.abs(np.corrcoef(h.T) - np.eye(6)).max() < 0.03                   # independent components
    assert (_draw(1000, 0, 0.0, mx, 3, 0) == 0).all() and (_draw(1000, 0, 1.0, mx, 3, 0) != 0).any(axis=1).all()


def test_env_manager_uses_the_device_draw():
    """lmf2 (disturbance on, p = 0.05) in the CPU twin: one agx_disturbance_draw per physics step, no torch RNG consumed by the step"""
    from aerial_gym_simulator_b200.sim import SimBuilder
    import aerial_gym_simulator_b200.task  # noqa: F401

    with cpu_stack() as proxy:
        env = SimBuilder().build_env("base_sim", "empty_env", "lmf2", "lmf2_velocity_control", "cpu", args={"seed": 11}, num_envs=200,
                                     use_warp=False, headless=True)
        env.reset()
        s0 = torch.get_rng_state().clone()
        for _ in range(3):
            env.step(actions=torch.zeros(200, 4))
        assert proxy.calls["agx_disturbance_draw"] == 3 and env._dist_counter == 3
        assert torch.equal(torch.get_rng_state(), s0)
        want = D.draw(200, 0, 0.05, env.spec.max_disturbance, env._dist_seed, 2)
        assert np.array_equal(env._dist_buf.numpy(), want) and 0 < (want != 0).any(axis=1).sum() < 40
        env2 = SimBuilder().build_env("base_sim", "empty_env", "lmf2", "lmf2_velocity_control", "cpu", args={"seed": 11, "reset_rng": "torch"},
                                      num_envs=200, use_warp=False, headless=True)
        env2.reset()
        n0 = proxy.calls["agx_disturbance_draw"]
        s1 = torch.get_rng_state().clone()
        env2.step(actions=torch.zeros(200, 4))
        assert proxy.calls["agx_disturbance_draw"] == n0 and not torch.equal(torch.get_rng_state(), s1)  # torch mode: the reference's draws


def test_in_kernel_draw_equals_explicit_disturbance_tensor():
    """CPU twin of test_graph_step_gpu.py::test_in_kernel_disturbance_equals_separate_draw: the physics sub-step text of
    hp1_step_kernel with AgxHp1Buffers.dist_counter set (draw inside the step: counter word + offset + sub-step) against the same
    text fed the ORACLE's draw as an explicit [N,6] tensor -- single steps and three fused sub-steps in one call."""
    from aerial_gym_simulator_b200.hp1 import build_config
    from tests import _hp1_common as H
    from tests._shadow_hp1 import ShadowHp1Engine

    spec = H.spec_for("octa_velocity")
    spec.enable_disturbance, spec.prob_apply_disturbance, spec.max_disturbance = True, 0.1, [4.75, 4.75, 4.75, 0.03, 0.03, 0.03]
    n, off = 513, 1000
    root, actions, params = H.random_inputs(spec, n, seed=21)
    e1, e2 = (ShadowHp1Engine(spec, n, seed=9, env_id_offset=off) for _ in range(2))
    for e in (e1, e2):
        H.load_engine_state(e, root, params)
    cfg = e1.cfg
    assert cfg.dist_prob == np.float32(0.1) and cfg.dist_seed == build_config(spec, n, seed=9).dist_seed
    ctr = torch.tensor([0], dtype=torch.int32)
    hits = 0
    for step in range(6):
        d = D.draw(n, off, float(cfg.dist_prob), list(cfg.dist_max), int(cfg.dist_seed), step)
        hits += int((d != 0).any(axis=1).sum())
        e1.physics_step(actions, disturbance=torch.from_numpy(d), physics_steps=1)
        e2.physics_step(actions, physics_steps=1, dist_counter=ctr, dist_offset=step)
        assert torch.equal(e1.root_state, e2.root_state), f"step {step}"
    assert hits > 100
    for step in (6, 7, 8):
        e1.physics_step(actions, disturbance=torch.from_numpy(D.draw(n, off, float(cfg.dist_prob), list(cfg.dist_max), int(cfg.dist_seed), step)),
                        physics_steps=1)
    ctr.fill_(6)
    e2.physics_step(actions, physics_steps=3, dist_counter=ctr, dist_offset=0)
    assert torch.equal(e1.root_state, e2.root_state)
    # no counter, no tensor: no disturbance
    e3 = ShadowHp1Engine(spec, n, seed=9, env_id_offset=off)
    H.load_engine_state(e3, root, params)
    e4 = ShadowHp1Engine(spec, n, seed=9, env_id_offset=off)
    H.load_engine_state(e4, root, params)
    e3.physics_step(actions)
    e4.physics_step(actions, disturbance=torch.zeros(n, 6))
    assert torch.equal(e3.root_state, e4.root_state)
