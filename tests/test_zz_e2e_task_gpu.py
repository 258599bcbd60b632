"""Motor-command position tasks (sim2real_end_to_end on tinyprop, sim2real_px4 on x500) on the GPU, through the C ABI: agx_e2e_reward /
agx_e2e_obs against the fixtures produced by the reference's own functions, and the two tasks end to end against the oracle.

(Named test_zz_*: written after the round's GPU budget was spent; the device code is verified on CPU through the host shadow build,
tests/test_e2e_task_cpu.py, and the host logic through tests/test_host_stack_cpu.py.)"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from aerial_gym_simulator_b200 import _lib
from oracle import e2e_task_oracle as E

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden", "e2e_task_epilogue.npz")
DEV = "cuda:0"


def _d(x, dtype=torch.float32):
    return torch.tensor(np.asarray(x), dtype=dtype, device=DEV).contiguous()


def _p(t):
    return C.c_void_p(t.data_ptr())


def _params(tag, crash_dist):
    p = _lib.AgxE2ERewardParams()
    for k, v in E.E2E_PARAMS[tag].items():
        setattr(p, k, float(v))
    p.crash_dist = float(crash_dist)
    return p


@pytest.mark.parametrize("tag", ["end_to_end", "px4"])
def test_kernels_match_reference_fixture(tag):
    d = np.load(G)
    lib, n = _lib.load(), d["pos"].shape[0]
    st = torch.zeros(n, 13, device=DEV)
    st[:, 0:3], st[:, 3:7], st[:, 7:10] = _d(d["pos"]), _d(d["quat"]), _d(d["linvel"])
    t = [_d(d["body_angvel"]), _d(d[f"{tag}_actions"]), _d(d[f"{tag}_prev_actions"]), _d(d["prev_pos_error"])]
    params = _params(tag, d[f"{tag}_crash_dist"])
    cr, rew = _d(d["crashes_in"], torch.uint8), torch.zeros(n, device=DEV)
    _lib.check(lib.agx_e2e_reward(n, _p(st), 13, _p(t[0]), None, _p(t[1]), _p(t[2]), _p(t[3]), C.byref(params), _p(cr), _p(rew), None),
               "agx_e2e_reward")
    torch.cuda.synchronize()
    assert torch.equal(cr.cpu().bool(), torch.tensor(d[f"{tag}_crashes_out"]))
    ref = torch.tensor(d[f"{tag}_reward"])
    # (towards_goal multiplies a difference of two nearly equal fp32 norms by up to 100: see tests/test_e2e_task_cpu.py)
    assert torch.allclose(rew.cpu(), ref, rtol=1e-5, atol=3e-4 if tag == "px4" else 5e-5), (rew.cpu() - ref).abs().max()
    for stride in (15, 24):
        nz = _d(d[f"{tag}_noise"])
        obs = torch.full((n, stride), 7.0, device=DEV)
        _lib.check(lib.agx_e2e_obs(n, _p(st), 13, _p(t[0]), None, _p(nz), _p(obs), stride, None), "agx_e2e_obs")
        torch.cuda.synchronize()
        assert torch.allclose(obs.cpu()[:, :15], torch.tensor(d[f"{tag}_obs"]), rtol=1e-5, atol=1e-5)
        assert (obs.cpu()[:, 15:] == 7.0).all()
    assert lib.agx_e2e_obs(n, _p(st), 13, _p(t[0]), None, None, _p(obs), 24, None) == -3
    assert lib.agx_e2e_obs(n, _p(st), 13, _p(t[0]), None, _p(nz), _p(obs), 14, None) == -1


def test_kernels_match_oracle_on_random_inputs():
    g = torch.Generator().manual_seed(4)
    n = 30001
    r = lambda *s: torch.randn(*s, generator=g)
    q = r(n, 4)
    q = q / q.norm(dim=1, keepdim=True)
    pos, vel, w, target = r(n, 3), r(n, 3), r(n, 3), r(n, 3) * 0.2
    act, prev_act = torch.rand(n, 4, generator=g) * 1.0 + 0.2, torch.rand(n, 4, generator=g) * 1.0 + 0.2
    prev_err = (target - pos) + 0.05 * r(n, 3)
    crashes = torch.rand(n, generator=g) < 0.03
    p = dict(E.E2E_PARAMS["end_to_end"])
    want, cr = E.compute_reward(target - pos, q, vel, w, crashes, act, prev_act, prev_err, 1.5, p)
    st = torch.zeros(n, 13)
    st[:, 0:3], st[:, 3:7], st[:, 7:10] = pos, q, vel
    std = st.to(DEV)
    t = [x.to(DEV).contiguous() for x in (w, target, act, prev_act, prev_err)]
    crd, rew = crashes.to(DEV).to(torch.uint8), torch.zeros(n, device=DEV)
    params = _params("end_to_end", 1.5)
    _lib.check(_lib.load().agx_e2e_reward(n, _p(std), 13, _p(t[0]), _p(t[1]), _p(t[2]), _p(t[3]), _p(t[4]), C.byref(params), _p(crd), _p(rew), None),
               "agx_e2e_reward")
    near = ((target - pos).norm(dim=1) - 1.5).abs() < 1e-5  # a position error within 10 um of the crash radius may fall either way
    assert torch.equal(crd.cpu().bool()[~near], cr[~near])
    assert torch.allclose(rew.cpu(), want, rtol=1e-5, atol=5e-5), (rew.cpu() - want).abs().max()
    noise = r(n, 12) * 0.01
    obs = torch.zeros(n, 15, device=DEV)
    nzd = noise.to(DEV)
    _lib.check(_lib.load().agx_e2e_obs(n, _p(std), 13, _p(t[0]), _p(t[1]), _p(nzd), _p(obs), 15, None), "agx_e2e_obs")
    want_obs = E.process_obs(pos, q, vel, w, target, noise)
    ok = (2 * (q[:, 3] * q[:, 1] - q[:, 2] * q[:, 0])).abs() < 0.9999  # asin slope near gimbal lock (as for the HP1 Euler angles)
    assert torch.allclose(obs.cpu()[ok], want_obs[ok], rtol=1e-5, atol=2e-5), (obs.cpu()[ok] - want_obs[ok]).abs().max()
    assert torch.isfinite(obs).all()


@pytest.mark.parametrize("name,tag", [("position_setpoint_task_sim2real_end_to_end", "end_to_end"), ("position_setpoint_task_sim2real_px4", "px4")])
def test_motor_command_tasks_end_to_end(name, tag):
    import aerial_gym_simulator_b200.task  # noqa: F401
    from aerial_gym_simulator_b200.registry._core import task_registry

    N = 300
    task = task_registry.make_task(name, seed=3, num_envs=N, headless=True)
    obs, rew, term, trunc, info = task.reset()
    assert obs["observations"].shape == (N, 15)
    od = task.obs_dict
    g = torch.Generator(device=DEV).manual_seed(1)
    for step in range(5):
        a = torch.rand(N, 4, device=DEV, generator=g) * 2.4 - 1.2
        if step == 3:
            task.sim_env.sim_steps[5] = task.task_config.episode_len_steps
        ep = int(task.sim_env.engine.episode_count[5])
        out = task.step(a)
        assert out[0] is obs and out[1] is rew
        if step == 3:
            assert bool(trunc[5]) and int(task.sim_env.engine.episode_count[5]) == ep + 2 and int(task.sim_env.sim_steps[5]) == 0
    torch.cuda.synchronize()
    assert torch.isfinite(rew).all() and torch.isfinite(obs["observations"]).all()
    assert torch.allclose(task.prev_pos_error, task.target_position - od["robot_position"]) and torch.equal(task.prev_actions, task.actions)
    p = dict(E.E2E_PARAMS[tag])
    crashes_in = od["crashes"].cpu().clone()
    task.compute_rewards_and_crashes(od)
    torch.cuda.synchronize()
    c = lambda t: t.cpu()
    want, cr = E.compute_reward(c(task.target_position - od["robot_position"]), c(od["robot_orientation"]), c(od["robot_linvel"]),
                                c(od["robot_body_angvel"]), crashes_in, c(task.actions), c(task.prev_actions), c(task.prev_pos_error),
                                task.task_config.crash_dist, p)
    assert torch.allclose(c(task.rewards), want, rtol=1e-5, atol=3e-4) and torch.equal(c(od["crashes"]), cr)
    noise = torch.randn(N, 12, device=DEV) * 0.01
    task.process_obs_for_task(noise)
    torch.cuda.synchronize()
    want_obs = E.process_obs(c(od["robot_position"]), c(od["robot_orientation"]), c(od["robot_linvel"]), c(od["robot_body_angvel"]),
                             c(task.target_position), c(noise))
    assert torch.allclose(c(obs["observations"]), want_obs, rtol=1e-5, atol=2e-5)
    task.close()
