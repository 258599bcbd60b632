"""Dynamic obstacles on the GPU: agx_obstacle_step against the oracle's spec (parity unpinned: PhysX moves the obstacles in
the reference) and the "dynamic_env" environment end to end -- env.step(actions, env_actions=twist) as in the reference's
examples/dynamic_env_example.py, with the ray-cast scene re-posed every env step.

(Named test_zz_*: written after the round's GPU budget was spent; the device code is verified on CPU through the host
shadow build, tests/test_obstacles_cpu.py.)"""
import ctypes as C

import numpy as np
import pytest
import torch

from aerial_gym_simulator_b200 import _lib
from oracle import hp2_oracle as RO
from oracle import obstacle_oracle as OB
from tests import _hp2_common as H2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _random_state(n, a, g, stride=13):
    s = torch.zeros(n, a, stride)
    s[..., 0:3] = torch.randn(n, a, 3, generator=g) * 4
    q = torch.randn(n, a, 4, generator=g)
    s[..., 3:7] = q / q.norm(dim=-1, keepdim=True)
    s[..., 7:13] = torch.randn(n, a, 6, generator=g)
    return s


@pytest.mark.parametrize("with_twist,substeps,stride", [(True, 1, 13), (True, 10, 13), (False, 3, 13), (True, 4, 16), (True, 0, 13)])
def test_obstacle_step_kernel_matches_oracle(with_twist, substeps, stride):
    g = torch.Generator().manual_seed(3)
    n, a = 301, 35
    s = _random_state(n, a, g, stride)
    s[..., 13:] = 42.0
    s[0, 0, 10:13] = 0.0
    tw = torch.randn(n, a, 6, generator=g) * 1.5 if with_twist else None
    if with_twist:
        tw[0, 0, 3:6] = 0.0
    sd = s.clone().to(DEV).contiguous()
    twd = tw.to(DEV).contiguous() if with_twist else None
    _lib.check(_lib.load().agx_obstacle_step(n, a, C.c_void_p(sd.data_ptr()), stride, C.c_void_p(twd.data_ptr()) if with_twist else None,
                                             0.01, substeps, 0.1, 0.1, None), "agx_obstacle_step")
    torch.cuda.synchronize()
    got = sd.cpu()
    want = OB.obstacle_step(s[..., :13], tw, 0.01, substeps)
    assert torch.allclose(got[..., :13], want, rtol=1e-5, atol=1e-6), (got[..., :13] - want).abs().max()
    assert (got[..., 13:] == 42.0).all()
    if substeps == 0:
        assert torch.equal(got, s)


def test_obstacle_step_argument_validation():
    lib = _lib.load()
    x = torch.zeros(64, device=DEV)
    assert lib.agx_obstacle_step(1, 1, C.c_void_p(x.data_ptr()), 12, None, 0.01, 1, 0.0, 0.0, None) == -1  # stride < 13
    assert lib.agx_obstacle_step(1, 1, C.c_void_p(x.data_ptr()), 13, None, 0.0, 1, 0.0, 0.0, None) == -1   # dt <= 0
    assert lib.agx_obstacle_step(1, 1, None, 13, None, 0.01, 1, 0.0, 0.0, None) == -3
    assert lib.agx_obstacle_step(0, 5, None, 13, None, 0.01, 1, 0.0, 0.0, None) == 0


def test_dynamic_env_end_to_end():
    """dynamic_env + lmf2 (depth camera): obstacles follow env_actions by the spec, and the next render sees them where they
    are (bit-identical to the brute-force oracle fed with the advanced poses)."""
    import aerial_gym_simulator_b200.task  # noqa: F401
    from aerial_gym_simulator_b200.sim import SimBuilder

    N = 6
    env = SimBuilder().build_env("base_sim", "dynamic_env", "lmf2", "lmf2_position_control", DEV, args={"seed": 2}, num_envs=N,
                                 use_warp=True, headless=True)
    gtd = env.get_obs()
    A = gtd["num_obstacles_in_env"]
    assert A == 40 and gtd["num_env_actions"] == 6
    env.reset()
    ast = gtd["env_asset_state_tensor"]
    twist = torch.zeros(N, A, 6, device=DEV)
    twist[..., 0], twist[..., 1], twist[..., 5] = -1.0, 0.5, 0.8
    actions = torch.zeros(N, 4, device=DEV)
    before = ast.cpu().clone()
    env.step(actions=actions, env_actions=twist)
    torch.cuda.synchronize()
    want = OB.obstacle_step(before, twist.cpu(), gtd["dt"], 10, 0.1, 0.1)  # 10 physics steps per env step (dynamic_environment.py:17)
    assert torch.allclose(ast.cpu(), want, rtol=1e-5, atol=1e-5), (ast.cpu() - want).abs().max()
    assert torch.allclose(ast[..., 0].cpu(), before[..., 0] - 0.1 * 0.999, atol=1e-4)
    assert gtd["env_actions"] is twist or torch.equal(gtd["env_actions"], twist)
    # a wrong shape is refused loudly
    with pytest.raises(ValueError, match="env_actions"):
        env.step(actions=actions, env_actions=torch.zeros(N, A + 1, 6, device=DEV))
    # render after the move: the scene was re-posed
    env.render()
    torch.cuda.synchronize()
    sc = env.scene
    assert torch.allclose(env._obj_pose.cpu(), torch.gather(ast[..., 0:7], 1, env._obj_asset.unsqueeze(-1).expand(-1, -1, 7)).cpu())
    tris, segs, cnt = RO.build_world_tris(env._obj_pose.cpu().numpy(), sc.obj_template.cpu().numpy(), sc.obj_seg_counter.cpu().numpy(),
                                          sc.tmpl_tri_offset.cpu().numpy(), sc.tmpl_tris.cpu().numpy(), sc.tmpl_seg_base.cpu().numpy(),
                                          sc.tmpl_seg_mask.cpu().numpy(), sc.K * sc.L)
    so, _ = H2.oracle_sensor(env.sensor_cfg)
    ref_pix, ref_seg = RO.cast(so, gtd["robot_state_tensor"][:, :7].cpu().numpy(), env.sensor_mount.cpu().numpy(), None, tris, segs, cnt)
    assert np.array_equal(gtd["depth_range_pixels"].cpu().numpy(), ref_pix)
    assert np.array_equal(gtd["segmentation_pixels"].cpu().numpy(), ref_seg)
    # refit switched off: the reference's behaviour (stale meshes until the next reset)
    env2 = SimBuilder().build_env("base_sim", "dynamic_env", "lmf2", "lmf2_position_control", DEV,
                                  args={"seed": 2, "refit_dynamic_obstacles": False}, num_envs=N, use_warp=True, headless=True)
    env2.reset()
    pose0, ast2 = env2._obj_pose.clone(), env2.get_obs()["env_asset_state_tensor"]
    before2 = ast2.clone()
    env2.step(actions=actions, env_actions=twist)
    torch.cuda.synchronize()
    assert torch.equal(env2._obj_pose, pose0)                                            # the ray-cast scene stays where it was ...
    assert torch.allclose(ast2[..., 0], before2[..., 0] - 0.1 * 0.999, atol=1e-4)        # ... while the obstacle states moved


def test_disturbance_draw_kernel_matches_oracle():
    """agx_disturbance_draw (device-RNG form of BaseMultirotor.apply_disturbance): integer Philox + two multiplies, so bit-exact"""
    from oracle import disturbance_oracle as D

    lib, n = _lib.load(), 70001
    mx = [4.75, 4.75, 4.75, 0.03, 0.03, 0.03]
    m = (C.c_float * 6)(*mx)
    out = torch.zeros(n, 6, device=DEV)
    for off, counter in ((0, 0), (123456, 9)):
        _lib.check(lib.agx_disturbance_draw(n, off, 0.05, m, 0xFEED_0000_1234, counter, C.c_void_p(out.data_ptr()), None), "agx_disturbance_draw")
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), D.draw(n, off, 0.05, mx, 0xFEED_0000_1234, counter))
    assert lib.agx_disturbance_draw(n, 0, 1.5, m, 1, 0, C.c_void_p(out.data_ptr()), None) == -1
    assert lib.agx_disturbance_draw(n, 0, 0.5, None, 1, 0, C.c_void_p(out.data_ptr()), None) == -3
