"""Reference-facing object graph on the GPU (added after the round's GPU budget was spent; the CPU twins of these tests are in
tests/test_host_stack_cpu.py): env.robot_manager / env.IGE_env views, controller gains read / set between steps."""
import pytest
import torch

import aerial_gym_simulator_b200.task  # noqa: F401
from aerial_gym_simulator_b200.sim import SimBuilder

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("per_env", [False, True])
def test_controller_gains_read_and_set(per_env):
    args = {"seed": 1, "per_env_params": "all"} if per_env else {"seed": 1}
    env = SimBuilder().build_env("base_sim", "empty_env", "base_quadrotor", "lee_position_control", DEV, args=args, num_envs=64, headless=True)
    ctrl = env.robot_manager.robot.controller
    mid = torch.tensor([2.5, 2.5, 1.5], device=DEV)
    assert torch.allclose(ctrl.K_pos_tensor_current, mid.expand(64, 3)) and ctrl.K_pos_tensor_current.device == env.device
    act = torch.zeros(64, 4, device=DEV)
    act[:, 0] = 1.0

    def run(scale):
        env.reset()
        env.engine.root_state[:] = 0.0  # same start for both runs: origin, level, at rest
        env.engine.root_state[:, 6] = 1.0
        env.engine.motor_thrust[:] = 0.25 * 9.81 / 4.0
        env.engine.refresh()
        ctrl.set_controller_gains(scale * mid, ctrl.K_linvel_tensor_current.clone(), ctrl.K_rot_tensor_current.clone(),
                                  ctrl.K_angvel_tensor_current.clone())
        for _ in range(20):
            env.step(act)
        torch.cuda.synchronize()
        return env.global_tensor_dict["robot_position"][:, 0].clone()
    slow, fast = run(0.5), run(2.0)
    assert (fast > slow + 1e-4).all() and torch.allclose(ctrl.K_pos_tensor_current, 2.0 * mid.expand(64, 3))
    if not per_env:
        with pytest.raises(RuntimeError, match="per_env_params"):
            ctrl.set_controller_gains(torch.rand(64, 3) + 1.0, 2.5, 1.0, 0.15)
    env.delete_env()


def test_object_graph_views():
    env = SimBuilder().build_env("base_sim", "env_with_obstacles", "base_quadrotor_with_camera", "lee_velocity_control", DEV, args={"seed": 1},
                                 num_envs=4, headless=True, use_warp=True)
    rm = env.robot_manager
    assert rm.robot.cfg.sensor_config.enable_camera and rm.robot.controller_config.num_actions == 4
    assert rm.robot_masses.shape == (4,) and rm.robot_inertias.shape == (4, 3, 3) and rm.warp_sensor is env.sensor
    assert env.IGE_env.num_assets_per_env == env.num_obs_in_env + 1 and env.sim_config.sim.dt == 0.01
    env.reset()
    env.step(torch.zeros(4, 4, device=DEV))
    env.render()
    torch.cuda.synchronize()
    assert torch.isfinite(env.global_tensor_dict["depth_range_pixels"]).all()
    env.delete_env()


def test_position_task_hooks_overridden():
    """a subclass overriding compute_rewards_and_crashes runs the un-fused sequence (physics launch -> hook -> truncation -> reset ->
    observation hook); against the fused base class on the same seed: same flags and resets, rewards differ by the override's bonus"""
    from aerial_gym_simulator_b200.config.task_config import position_setpoint_task_config as C
    from aerial_gym_simulator_b200.task.position_setpoint_task import PositionSetpointTask

    class RewardToo(PositionSetpointTask):
        def compute_rewards_and_crashes(self, obs_dict):
            rew, crashes = super().compute_rewards_and_crashes(obs_dict)
            return rew + 1.25, crashes

    cfg = lambda: type("cfg", (C,), dict(device=DEV, num_envs=256, episode_len_steps=6, reward_parameters=dict(C.reward_parameters)))
    base, hooked = PositionSetpointTask(cfg(), seed=5, headless=True), RewardToo(cfg(), seed=5, headless=True)
    base.reset()
    hooked.reset()
    g = torch.Generator(device=DEV).manual_seed(0)
    for step in range(16):
        a = torch.rand(256, 4, device=DEV, generator=g) * 2 - 1
        if step == 3:
            base.sim_env.engine.root_state[5:8, 0] = 9.0
            hooked.sim_env.engine.root_state[5:8, 0] = 9.0
        o0, r0, te0, tr0, _ = base.step(a.clone())
        o1, r1, te1, tr1, _ = hooked.step(a.clone())
        torch.cuda.synchronize()
        assert torch.equal(te0, te1) and torch.equal(tr0, tr1), step
        assert torch.allclose(r1, r0 + 1.25, rtol=1e-5, atol=1e-5) and torch.allclose(o1["observations"], o0["observations"], rtol=1e-5, atol=1e-5), step
    assert torch.equal(base.sim_env.sim_steps, hooked.sim_env.sim_steps) and int(base.sim_env.engine.episode_count.min()) >= 2
    base.close()
    hooked.close()
