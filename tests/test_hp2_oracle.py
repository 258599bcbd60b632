"""CPU checks of the HP2 oracle (the brute-force ray caster) -- known answers + an independent
float64 numpy restatement.  The closest-hit query is PARITY UNPINNED w.r.t. Warp (not installable);
what is pinned here is that the oracle is a correct exhaustive closest hit."""
import numpy as np
import torch

from oracle import hp2_oracle as RO
from tests import _hp2_common as H


def test_known_answer_single_box():
    """Unit box centred 3 m ahead of a camera looking along +x: centre pixel depth 2.5 m ->
    0.25 after normalisation; misses -> far value / max_range = 1.0 (SURVEY 8a constants)."""
    templates = [RO.box_template((1.0, 1.0, 1.0))]
    pose = np.zeros((1, 1, 7), np.float32)
    pose[0, 0] = [3, 0, 0, 0, 0, 0, 1]
    tris, segs, cnt = RO.build_world_tris(pose, np.zeros((1, 1), np.int32), np.full((1, 1), 7, np.int32),
                                          np.array([0, 12], np.int32), templates[0], np.full(12, 100, np.int32), np.ones(12, np.int32), 12)
    assert cnt[0] == 12
    s, _ = H.oracle_sensor(H.CamCfg)
    robot = np.array([[0, 0, 0, 0, 0, 0, 1]], np.float32)
    mount = np.array([[[0, 0, 0, 0, 0, 0, 1]]], np.float32)
    pix, seg = RO.cast(s, robot, mount, None, tris, segs, cnt)
    cy, cx = H.CamCfg.height // 2, H.CamCfg.width // 2
    assert abs(pix[0, 0, cy, cx] - 0.25) < 1e-6
    assert seg[0, 0, cy, cx] == 107  # base 100 + counter 7 * mask 1
    assert pix[0, 0, 0, 0] == 1.0 and seg[0, 0, 0, 0] == -2
    # depth image: every hit pixel on the front face has the same depth (range would vary)
    front = seg[0, 0] == 107
    assert front.sum() > 20 and np.allclose(pix[0, 0][front], 0.25, atol=2e-6)
    s2, _ = H.oracle_sensor(H.cfg_variant(H.CamCfg, calculate_depth=False))
    pix2, _ = RO.cast(s2, robot, mount, None, tris, segs, cnt)
    assert pix2[0, 0][front].max() > 0.2501 and abs(pix2[0, 0, cy, cx] - 0.25) < 1e-5


def _numpy_closest_hit(tris, o, d, max_t):
    v0, e1, e2 = tris[:, 0:3].astype(np.float64), tris[:, 3:6].astype(np.float64), tris[:, 6:9].astype(np.float64)
    pv = np.cross(d, e2)
    det = (e1 * pv).sum(1)
    ok = np.abs(det) > 1e-20
    inv = np.where(ok, 1.0 / np.where(ok, det, 1), 0)
    tv = o - v0
    u = (tv * pv).sum(1) * inv
    qv = np.cross(tv, e1)
    v = (qv * d).sum(1) * inv
    t = (e2 * qv).sum(1) * inv
    ok &= (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t >= 0) & (t < max_t)
    if not ok.any():
        return -1, None
    t = np.where(ok, t, np.inf)
    i = int(np.argmin(t))
    return i, t[i]


def test_against_independent_numpy_restatement():
    sc = H.make_scene(2, 9, seed=4)
    tris, segs, cnt = H.oracle_tris(sc)
    cfg = H.cfg_variant(H.CamCfg, height=12, width=16, calculate_depth=False)
    s, _ = H.oracle_sensor(cfg, fuse=False)
    robot = H.robot_poses(2, 5)[:, :7].numpy()
    mount = np.zeros((2, 1, 7), np.float32)
    mount[..., 6] = 1
    pix, seg = RO.cast(s, robot, mount, None, tris, segs, cnt)
    # float64 restatement of pose compose + ray generation
    from oracle import hp1_oracle as O
    fq = torch.tensor(RO.quat_from_euler_deg(cfg.euler_frame_rot_deg)).double()[None]
    kinv, _, _ = RO.camera_kinv(cfg.width, cfg.height, cfg.horizontal_fov_deg)
    mism = 0
    for e in range(2):
        rq = torch.tensor(robot[e:e + 1, 3:7]).double()
        sq = O.quat_mul(rq, O.quat_mul(torch.tensor([[0, 0, 0, 1.0]]).double(), fq))
        for y in range(cfg.height):
            for x in range(cfg.width):
                uv = torch.tensor(kinv.astype(np.float64) @ np.array([x, y, 1.0]))[None]
                rd = O.quat_rotate(sq, uv)[0].numpy()
                rd = rd / np.linalg.norm(rd)
                i, t = _numpy_closest_hit(tris[e, :cnt[e]], robot[e, :3].astype(np.float64), rd, cfg.max_range)
                if i < 0:
                    mism += int(seg[e, 0, y, x] != -2)
                else:
                    if seg[e, 0, y, x] != segs[e, i] or abs(pix[e, 0, y, x] - t) > 1e-4 * max(1, t):
                        mism += 1
    assert mism <= 2, mism  # silhouette pixels may legitimately flip between fp32 and fp64


def test_lidar_table_and_intrinsics_match_product_host_code():
    from aerial_gym_simulator_b200 import hp2
    a = RO.lidar_ray_table(8, 32, -180, 180, 0, 90)
    b = hp2.lidar_ray_table(8, 32, -180, 180, 0, 90)
    assert np.abs(a - b).max() <= 1.2e-7
    ka, cxa, cya = RO.camera_kinv(64, 48, 87.0)
    kb, cxb, cyb = hp2.camera_intrinsics(64, 48, 87.0)
    assert np.array_equal(ka, kb) and (cxa, cya) == (cxb, cyb) == (32, 24)
    assert np.array_equal(RO.box_template((1, 2, 3)), hp2.box_triangles((1, 2, 3)))


def test_collision_oracle_known_answers():
    """a14 oracle: sphere vs unit box at the origin."""
    tris = RO.box_template((1.0, 1.0, 1.0))[None].copy()  # world == object frame; stored as v0,v1,v2
    T = tris.reshape(1, 12, 3, 3)
    slabs = np.concatenate([T[:, :, 0], T[:, :, 1] - T[:, :, 0], T[:, :, 2] - T[:, :, 0]], axis=-1).astype(np.float32)
    cnt = np.array([12], np.int32)
    for pos, want_d in [((2.0, 0, 0), 1.5), ((0.5 + 0.1, 0.2, -0.3), 0.1), ((1.0, 1.0, 1.0), np.sqrt(0.75)), ((0, 0, 0), 0.5)]:
        pose = np.array([[*pos, 0, 0, 0, 1]], np.float32)
        hit, d2 = RO.collide(pose, 0.18, slabs, cnt)
        assert abs(np.sqrt(d2[0]) - want_d) < 1e-6, (pos, np.sqrt(d2[0]), want_d)
        assert bool(hit[0]) == (want_d <= 0.18)
