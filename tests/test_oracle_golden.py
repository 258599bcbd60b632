"""Pin the HP1 oracle against fixtures produced by the reference's own code
(tests/golden/make_golden.py) and the reference's only in-repo known-answer file."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import hp1_oracle as O
from tests._models import oracle_model

GOLD = os.path.join(os.path.dirname(__file__), "golden")
STEP_FILES = sorted(glob.glob(os.path.join(GOLD, "hp1_step_*.npz")))


def _load(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def _state_from_fixture(model, z, meta, s, dtype=torch.float32):
    N = meta["N"]
    st = O.make_state(model, N, dtype)
    st.root = torch.tensor(z[f"s{s}_root"], dtype=dtype)
    st.thrust = torch.tensor(z[f"s{s}_thrust_in"], dtype=dtype)
    st.tau_inc = torch.tensor(z["tau_inc"], dtype=dtype)
    st.tau_dec = torch.tensor(z["tau_dec"], dtype=dtype)
    if "k_thrust" in z:
        st.k_thrust = torch.tensor(z["k_thrust"], dtype=dtype)
    if "K_pos" in z:
        st.K_pos = torch.tensor(z["K_pos"], dtype=dtype)
        st.K_vel = torch.tensor(z["K_vel"], dtype=dtype)
        st.K_rot = torch.tensor(z["K_rot"], dtype=dtype)
        st.K_angvel = torch.tensor(z["K_angvel"], dtype=dtype)
    return st


def _close(a, b, rtol=2e-6, scale=1.0, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    tol = rtol * np.maximum(np.abs(b), scale)
    bad = err > tol
    assert not bad.any(), f"{what}: max err {err.max():.3e} (tol {tol[bad].min():.3e}) at {np.argwhere(bad)[:3]}"


@pytest.mark.parametrize("path", STEP_FILES, ids=[os.path.basename(p)[9:-4] for p in STEP_FILES])
def test_robot_step_matches_reference(path):
    """a1-a12: derived states, controller wrench, allocation, motor model, per-link wrenches."""
    z, meta = _load(path)
    _check_oracle_against_step_fixture(oracle_model(meta["robot"], meta["controller"], meta["mass"], meta["inertia"]), z, meta)


REG_FILES = sorted(glob.glob(os.path.join(GOLD, "hp1_regstep_*.npz")))


@pytest.mark.parametrize("path", REG_FILES, ids=[os.path.basename(p)[12:-4] for p in REG_FILES])
def test_robot_step_matches_reference_registry_models(path):
    """the same for 13 more robot x controller pairs, the oracle model built from the product's registry spec
    (tests/golden/make_golden_registry.py)"""
    from tests import _hp1_common as H
    z, meta = _load(path)
    _check_oracle_against_step_fixture(H.oracle_model_from_spec(H.spec_from_registry(meta["robot"], meta["controller"])), z, meta)


def _check_oracle_against_step_fixture(model, z, meta):
    mask = meta["application_mask"]
    for s in range(meta["steps"]):
        st = _state_from_fixture(model, z, meta, s)
        actions = torch.tensor(z[f"s{s}_actions"])
        d = O.update_states(st.root)
        ref_euler = z[f"s{s}_euler"]
        if meta["robot"] == "base_rov":  # base_rov.py:245 keeps [0, 2 pi); the oracle / product wrap (documented difference)
            ref_euler = np.where(ref_euler > np.pi, ref_euler - 2.0 * np.pi, ref_euler)
        _close(d["euler"], ref_euler, what="euler", scale=1.0)
        _close(d["vehicle_orientation"], z[f"s{s}_vehicle_orientation"], what="veh q")
        _close(d["vehicle_linvel"], z[f"s{s}_vehicle_linvel"], what="veh v")
        _close(d["body_linvel"], z[f"s{s}_body_linvel"], what="body v")
        _close(d["body_angvel"], z[f"s{s}_body_angvel"], what="body w")
        a = torch.clamp(actions, -10.0, 10.0)
        cmd = O.controller_wrench(model, st, d, a)
        if f"s{s}_wrench_cmd" in z:
            ref_cmd = z[f"s{s}_wrench_cmd"]
            # torque magnitudes reach O(10) while individual terms cancel; scale by the row norm
            _close(cmd, ref_cmd, rtol=5e-6, scale=max(1.0, float(np.abs(ref_cmd).max())), what="wrench cmd")
        f_new, forces, torques = O.allocate(model, st, cmd)
        fscale = max(1.0, float(np.abs(z[f"s{s}_thrust_out"]).max()))
        _close(f_new, z[f"s{s}_thrust_out"], rtol=5e-6, scale=fscale, what="thrust")
        df, dtq = O.drag_wrench(model, d)
        ref_F = z[f"s{s}_force"]
        ref_T = z[f"s{s}_torque"]
        got_F = np.zeros_like(ref_F)
        got_T = np.zeros_like(ref_T)
        got_F[:, mask, :] = forces.numpy()
        got_T[:, mask, :] = torques.numpy()
        got_F[:, 0, :] += df.numpy()
        got_T[:, 0, :] += dtq.numpy()
        model.enable_disturbance = meta["enable_disturbance"]
        model.prob_apply_disturbance = meta["prob_apply_disturbance"]
        model.max_disturbance = tuple(meta["max_disturbance"])
        torch.manual_seed(int(z[f"s{s}_seed"]))
        dist = O.draw_disturbance(model, meta["N"])
        if dist is not None:
            got_F[:, 0, :] += dist[:, 0:3].numpy()
            got_T[:, 0, :] += dist[:, 3:6].numpy()
        _close(got_F, ref_F, rtol=5e-6, scale=fscale, what="link forces")
        _close(got_T, ref_T, rtol=5e-6, scale=fscale, what="link torques")


@pytest.mark.parametrize("tag", ["quad_attitude", "octa_velocity"])
def test_reset_matches_reference(tag):
    """a15: reset sampling in the reference's RNG call order (full-N draws, then gather)."""
    z, meta = _load(os.path.join(GOLD, f"hp1_reset_{tag}.npz"))
    step_z, step_meta = _load(os.path.join(GOLD, f"hp1_step_{tag}.npz"))
    model = oracle_model(meta["robot"], meta["controller"], step_meta["mass"], step_meta["inertia"])
    N = meta["N"]
    st = O.make_state(model, N)
    st.root = torch.tensor(z["root_before"])
    st.thrust = torch.tensor(z["thrust_before"])
    st.tau_inc = torch.tensor(z["before_tau_inc"])
    st.tau_dec = torch.tensor(z["before_tau_dec"])
    if "before_k_thrust" in z:
        st.k_thrust = torch.tensor(z["before_k_thrust"])
    for k in ("K_pos", "K_vel", "K_rot", "K_angvel"):
        setattr(st, k, torch.tensor(z["before_" + k]))
    mask = torch.zeros(N, dtype=torch.bool)
    mask[torch.tensor(z["env_ids"])] = True
    torch.manual_seed(int(z["seed"]))
    # robot-only part of the call order (the IGE bounds draws are not part of robot.reset_idx)
    r = lambda *s: torch.rand(*s)
    M = model.num_motors
    state = r(N, 13)
    if model.randomize_params:
        # the reference draws rand_like(upper[env_ids]) -> [k,3] per gain, not full-N
        k = int(mask.sum())
        gains = [r(k, 3) for _ in range(4)]
        full = []
        for g in gains:
            t = torch.zeros(N, 3)
            t[mask] = g
            full.append(t)
        kp, kv, kr, kw = full
    else:
        kp = kv = kr = kw = None
    ti, td, th = r(N, M), r(N, M), r(N, M)
    kt = r(N, M) if model.use_rps else None
    draws = O.ResetDraws(torch.zeros(N, 3), torch.zeros(N, 3), state, kp, kv, kr, kw, ti, td, th, kt)
    O.reset_envs(model, st, mask, draws)
    _close(st.root, z["root_after"], what="root")
    _close(st.thrust, z["thrust_after"], what="thrust", scale=1.0)
    _close(st.tau_inc, z["after_tau_inc"], what="tau_inc", scale=0.01)
    _close(st.tau_dec, z["after_tau_dec"], what="tau_dec", scale=0.01)
    if model.use_rps:
        _close(st.k_thrust, z["after_k_thrust"], what="k", scale=1e-5)
    for k in ("K_pos", "K_vel", "K_rot", "K_angvel"):
        _close(getattr(st, k), z["after_" + k], what=k)
    # derived states of ALL envs refreshed (base_multirotor.py:204-205)
    _close(st.derived["euler"], z["after_euler_angles"], what="euler")
    _close(st.derived["body_linvel"], z["after_body_linvel"], what="body_linvel")
    _close(st.derived["body_angvel"], z["after_body_angvel"], what="body_angvel")
    _close(st.derived["vehicle_orientation"], z["after_vehicle_orientation"], what="veh q")
    # bit-exact reset mask bookkeeping: untouched rows are bit-identical
    untouched = ~mask.numpy()
    assert np.array_equal(st.root.numpy()[untouched], z["root_before"][untouched])


def test_position_reward_matches_reference():
    """a16: reward + crash flags from the reference's compute_reward."""
    z = np.load(os.path.join(GOLD, "hp1_position_reward.npz"))
    n = z["pos"].shape[0]
    model = O.Hp1Model()
    st = O.make_state(model, n)
    st.root[:, 0:3] = torch.tensor(z["pos"])
    st.root[:, 3:7] = torch.tensor(z["quat"])
    st.derived = {
        "vehicle_orientation": torch.tensor(z["vehicle_orientation"]),
        "body_angvel": torch.tensor(z["body_angvel"]),
    }
    rew, cr = O.position_task_reward(st, torch.zeros(n, 3), torch.tensor(z["crashes_in"]))
    assert np.array_equal(cr.numpy(), z["crashes_out"])  # bit-exact flags
    _close(rew, z["reward"], rtol=2e-6, scale=1.0, what="reward")
    assert cr.numpy()[:6].any() and (rew.numpy()[cr.numpy()] == -20).all()


def test_motor_model_known_answer_csv():
    """a9 (Euler, RPS space, discrete mixing) vs sim2real/motorid_utilities/sample_sim_euler_integration.csv."""
    g = json.load(open(os.path.join(GOLD, "motor_euler_csv.json")))
    rows = np.array(g["rows_t_rps"])
    k, tau, dt = g["k"], g["tau"], g["dt"]
    model = O.Hp1Model(num_motors=1, integration_scheme="euler", use_rps=True, max_thrust=1e9, dt=dt,
                       allocation_matrix=np.ones((6, 1)), motor_directions=[1], link_r=np.zeros((1, 3)))
    st = O.make_state(model, 1, torch.float64)
    st.k_thrust[:] = k
    st.tau_inc[:] = tau
    st.tau_dec[:] = tau
    st.thrust[:] = k * rows[0, 1] ** 2
    ref = torch.full((1, 1), k * g["rps_ref"] ** 2, dtype=torch.float64)
    for i in range(1, len(rows)):
        st.thrust = O.motor_update(model, st, ref)
        rps = float(torch.sqrt(st.thrust / k))
        assert abs(rps - rows[i, 1]) < 2e-4 * max(1.0, rows[i, 1]), (i, rps, rows[i, 1])


def test_integrator_spec_invariants():
    """a13 is OUR spec (parity unpinned): check the physics it promises."""
    model = O.Hp1Model(linear_damping=0.0, angular_damping=0.0)
    N = 32
    g = torch.Generator().manual_seed(0)
    root = torch.zeros(N, 13, dtype=torch.float64)
    q = torch.randn(N, 4, generator=g, dtype=torch.float64)
    root[:, 3:7] = q / q.norm(dim=1, keepdim=True)
    root[:, 10:13] = torch.randn(N, 3, generator=g, dtype=torch.float64)
    # free fall, torque free: v_z decreases by g*dt, |q| = 1, rotational kinetic energy ~conserved
    J = torch.tensor(model.inertia)
    W0 = O.quat_rotate_inverse(root[:, 3:7], root[:, 10:13])
    E0 = 0.5 * ((W0 @ J) * W0).sum(1)
    r = root
    for _ in range(100):
        r = O.rigid_body_integrate(model, r, torch.zeros(N, 3, dtype=torch.float64), torch.zeros(N, 3, dtype=torch.float64))
    assert torch.allclose(r[:, 9], torch.full((N,), -9.81 * 1.0, dtype=torch.float64), atol=1e-9)
    assert torch.allclose(r[:, 3:7].norm(dim=1), torch.ones(N, dtype=torch.float64), atol=1e-12)
    W1 = O.quat_rotate_inverse(r[:, 3:7], r[:, 10:13])
    E1 = 0.5 * ((W1 @ J) * W1).sum(1)
    assert ((E1 - E0).abs() / E0 < 0.05).all()
    # hover thrust balances gravity exactly
    F = torch.zeros(N, 3, dtype=torch.float64)
    F[:, 2] = model.mass * 9.81
    r0 = torch.zeros(N, 13, dtype=torch.float64)
    r0[:, 6] = 1.0
    r1 = O.rigid_body_integrate(model, r0, F, torch.zeros(N, 3, dtype=torch.float64))
    assert r1[:, 7:10].abs().max() < 1e-12
    # velocity cap
    r0[:, 7] = 1e4
    r1 = O.rigid_body_integrate(model, r0, F, torch.zeros(N, 3, dtype=torch.float64))
    assert torch.allclose(r1[:, 7:10].norm(dim=1), torch.full((N,), 100.0, dtype=torch.float64))
