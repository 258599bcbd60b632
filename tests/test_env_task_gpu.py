"""GPU tests of the reference-facing surface: task_registry.make_task -> task.step / reset,
SimBuilder().build_env -> EnvManager.step / reset / render, against the oracle."""
import numpy as np
import pytest
import torch

from oracle import hp1_oracle as O
from oracle import hp2_oracle as RO
import aerial_gym_simulator_b200.task  # noqa: F401
from aerial_gym_simulator_b200.registry._core import task_registry
from aerial_gym_simulator_b200.sim import SimBuilder
from tests import _hp1_common as H
from tests import _hp2_common as H2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_task_surface_and_in_place_semantics():
    task = task_registry.make_task("position_setpoint_task", seed=3, num_envs=300, headless=True)
    out0 = task.reset()
    obs, rew, term, trunc, info = out0
    assert obs["observations"].shape == (300, 13) and rew.shape == (300,)
    assert term.dtype == torch.bool and trunc.dtype == torch.bool and isinstance(info, dict)
    assert task.num_envs == 300 and task.action_space.shape == (4,) and task.task_config.observation_space_dim == 13
    gtd = task.sim_env.get_obs()
    # position / orientation / velocities are views of ONE [N,13] row
    base = gtd["robot_state_tensor"]
    for k in ("robot_position", "robot_orientation", "robot_linvel", "robot_angvel"):
        assert gtd[k].data_ptr() >= base.data_ptr() and gtd[k]._base is base
    pos0 = gtd["robot_position"].clone()
    assert (pos0[:, 0] >= -0.8 - 1e-6).all() and (pos0[:, 0] <= -0.6 + 1e-6).all()  # ratio 0.1..0.2 of [-1,1]
    g = torch.Generator(device=DEV).manual_seed(0)
    for i in range(505):
        a = (torch.rand(300, 4, device=DEV, generator=g) * 2 - 1) * 0.05  # near hover: nobody leaves the 8 m ball
        out = task.step(a)
        assert out[0] is obs and out[1] is rew and out[2] is term and out[3] is trunc  # same objects, mutated in place
        if i == 499:
            assert not trunc.any() and not term.any() and (task.sim_env.sim_steps == 500).all()
    torch.cuda.synchronize()
    assert torch.isfinite(obs["observations"]).all() and torch.isfinite(rew).all()
    assert (task.sim_env.sim_steps <= 501).all()
    assert int(task.sim_env.engine.episode_count.min()) == 2  # initial reset + truncation at step 501
    assert (task.sim_env.sim_steps == 4).all()
    assert torch.equal(obs["observations"][:, 3:7], gtd["robot_orientation"])


def test_task_torch_rng_mode_follows_reference_call_order():
    """reset_rng='torch': uniforms come from torch.rand in the reference's order (bounds lo, bounds
    hi, state, motor tau_inc, tau_dec, thrust, k), consumed only when an env resets."""
    cfg = task_registry.get_task_config("position_setpoint_task")
    old = cfg.args
    cfg.args = {"reset_rng": "torch"}
    try:
        task = task_registry.make_task("position_setpoint_task", seed=5, num_envs=64, headless=True)
    finally:
        cfg.args = old
    N = 64
    # record every torch.rand the reset makes: shapes + order must be the reference's call order
    # (IGE bounds x2 [N,3] -> robot state [N,13] -> motor tau_inc, tau_dec, thrust, k [N,M])
    calls, orig = [], torch.rand

    def spy(*a, **k):
        out = orig(*a, **k)
        calls.append(out.clone())
        return out

    torch.rand = spy
    try:
        torch.manual_seed(1234)
        task.reset()
    finally:
        torch.rand = orig
    assert [tuple(c.shape) for c in calls] == [(N, 3), (N, 3), (N, 13), (N, 4), (N, 4), (N, 4), (N, 4)]
    eng = task.sim_env.engine
    model = H.oracle_model_from_spec(task.sim_env.spec)
    c = [x.cpu() for x in calls]
    draws = O.ResetDraws(c[0], c[1], c[2], None, None, None, None, c[3], c[4], c[5], c[6])
    st = O.make_state(model, N)
    O.reset_envs(model, st, torch.ones(N, dtype=torch.bool), draws)
    H.assert_close(eng.root_state, st.root, "reset state (torch rng order)", scale=1.0)
    H.assert_close(eng.motor_thrust, st.thrust, "reset thrust")
    H.assert_close(eng.k_thrust, st.k_thrust, "reset k", scale=1e-5)
    # same seed -> same episode start (determinism of the torch-RNG mode)
    first = eng.root_state.clone()
    torch.manual_seed(1234)
    task.reset()
    assert torch.equal(eng.root_state, first)
    # a step without resets must not consume the generator
    s0 = torch.cuda.get_rng_state(0).clone()
    task.step(torch.zeros(N, 4, device=DEV))
    assert torch.equal(torch.cuda.get_rng_state(0), s0)
    # force one truncation -> exactly one reset_idx worth of draws is consumed and that env restarts
    eng.sim_steps[7] = 500
    task.step(torch.zeros(N, 4, device=DEV))
    assert not torch.equal(torch.cuda.get_rng_state(0), s0)
    assert int(eng.sim_steps[7]) == 0 and bool(task.truncations[7]) and int(task.truncations.sum()) == 1


def test_env_manager_generic_step_and_depth_camera():
    """EnvManager surface on env_with_obstacles with a camera robot: 44 URDF boxes per env, 10
    fused physics sub-steps per env step, depth + segmentation render checked bit-exactly against
    the brute-force oracle fed with the env's own asset states."""
    env = SimBuilder().build_env("base_sim", "env_with_obstacles", "base_quadrotor_with_camera", "lee_velocity_control",
                                 DEV, args={"seed": 1}, num_envs=24, use_warp=True, headless=True)
    gtd = env.get_obs()
    assert gtd["num_obstacles_in_env"] == 44 and env.keep_in_env == 9  # 3 panels + 6 walls kept
    assert gtd["depth_range_pixels"].shape == (24, 1, 135, 240) and gtd["segmentation_pixels"].dtype == torch.int32
    env.reset()
    ast = gtd["env_asset_state_tensor"]
    assert ast.shape == (24, 44, 13)
    # walls sit on the bounds, the un-kept tail may be parked at -1000 only through the curriculum
    bmin, bmax = gtd["env_bounds_min"], gtd["env_bounds_max"]
    inside = (ast[..., 0:3] >= bmin.unsqueeze(1) - 1e-4) & (ast[..., 0:3] <= bmax.unsqueeze(1) + 1e-4)
    parked = ast[..., 0] < -900
    assert (inside.all(-1) | parked).all()
    a = torch.zeros(24, 4, device=DEV)
    a[:, 0] = 0.5
    p0 = gtd["robot_position"].clone()
    for _ in range(5):
        env.step(actions=a)
        env.post_reward_calculation_step()
    torch.cuda.synchronize()
    assert int(env.sim_steps[0]) == 5
    assert (gtd["robot_position"] - p0).norm(dim=1).mean() > 0.02  # 50 physics steps of forward velocity command
    # oracle render from the env's tensors
    sc = env.scene
    tris, segs, cnt = RO.build_world_tris(env._obj_pose.cpu().numpy(), sc.obj_template.cpu().numpy(), sc.obj_seg_counter.cpu().numpy(),
                                          sc.tmpl_tri_offset.cpu().numpy(), sc.tmpl_tris.cpu().numpy(), sc.tmpl_seg_base.cpu().numpy(),
                                          sc.tmpl_seg_mask.cpu().numpy(), sc.K * sc.L)
    so, _ = H2.oracle_sensor(env.sensor_cfg)
    ref_pix, ref_seg = RO.cast(so, gtd["robot_state_tensor"][:, :7].cpu().numpy(), env.sensor_mount.cpu().numpy(), None, tris, segs, cnt)
    assert np.array_equal(gtd["depth_range_pixels"].cpu().numpy(), ref_pix)
    assert np.array_equal(gtd["segmentation_pixels"].cpu().numpy(), ref_seg)
    seg = gtd["segmentation_pixels"]
    assert ((seg >= 9) & (seg <= 14)).any() and (seg >= 100).any()  # walls keep fixed ids, panels/objects get counters
    # masked reset re-randomises only those envs' obstacles and camera mounts
    before = ast.clone()
    env.reset_idx(torch.tensor([2, 5], device=DEV))
    changed = (ast != before).flatten(1).any(1)
    assert changed[2] and changed[5] and int(changed.sum()) == 2
