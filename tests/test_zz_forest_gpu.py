"""forest_env on the GPU: a tree of 13 tessellated cylinders (143 scene parts) + 35 boxes + the floor = 179 objects per env -- beyond the
128-leaf tile path, i.e. the per-ray BVH traversal -- rendered bit-identically to the brute-force oracle, per-link segmentation ids.

(Named test_zz_*: written after the round's GPU budget was spent; host logic verified on CPU, tests/test_host_stack_cpu.py.)"""
import numpy as np
import pytest
import torch

from oracle import hp2_oracle as RO
from tests import _hp2_common as H2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_forest_env_render_matches_oracle():
    import aerial_gym_simulator_b200.task  # noqa: F401
    from aerial_gym_simulator_b200.sim import SimBuilder

    N = 6
    env = SimBuilder().build_env("base_sim", "forest_env", "base_quadrotor_with_camera", "lee_velocity_control", DEV, args={"seed": 1}, num_envs=N,
                                 use_warp=True, headless=True)
    gtd, sc = env.get_obs(), env.scene
    assert sc.K == 13 * 11 + 35 + 1 and sc.L == 12 and gtd["num_obstacles_in_env"] == 37
    env.reset()
    a = torch.zeros(N, 4, device=DEV)
    for _ in range(2):
        env.step(actions=a)
        env.post_reward_calculation_step()
    torch.cuda.synchronize()
    tris, segs, cnt = RO.build_world_tris(env._obj_pose.cpu().numpy(), sc.obj_template.cpu().numpy(), sc.obj_seg_counter.cpu().numpy(),
                                          sc.tmpl_tri_offset.cpu().numpy(), sc.tmpl_tris.cpu().numpy(), sc.tmpl_seg_base.cpu().numpy(),
                                          sc.tmpl_seg_mask.cpu().numpy(), sc.K * sc.L)
    so, _ = H2.oracle_sensor(env.sensor_cfg)
    ref_pix, ref_seg = RO.cast(so, gtd["robot_state_tensor"][:, :7].cpu().numpy(), env.sensor_mount.cpu().numpy(), None, tris, segs, cnt)
    assert np.array_equal(gtd["depth_range_pixels"].cpu().numpy(), ref_pix)
    assert np.array_equal(gtd["segmentation_pixels"].cpu().numpy(), ref_seg)
    assert (ref_seg >= 100).any()
    # collision flags against the cylinders too
    flags, _ = RO.collide(gtd["robot_state_tensor"][:, :7].cpu().numpy(), env.robot.collision_radius, tris, cnt)
    crashes = torch.zeros(N, dtype=torch.bool, device=DEV)
    sc.collide(env.engine.root_state, env.robot.collision_radius, crashes)
    assert np.array_equal(crashes.cpu().numpy(), flags)
