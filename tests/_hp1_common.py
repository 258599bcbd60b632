"""Shared helpers for the HP1 parity tests (CUDA path vs oracle on identical inputs)."""
import dataclasses

import numpy as np
import torch

from oracle import hp1_oracle as O
from aerial_gym_simulator_b200 import _lib
from aerial_gym_simulator_b200.hp1 import Hp1Engine, MultirotorSpec

# fp32 parity bar (BASELINE.json north_star): 1e-5 relative.  "Relative" is taken against
# max(|reference value|, characteristic scale of the quantity in the batch): entries that are
# near zero by cancellation carry the absolute rounding noise of the O(scale) terms they are
# made of, in the reference as much as here.
RTOL = 1e-5


def oracle_model_from_spec(spec: MultirotorSpec) -> O.Hp1Model:
    return O.Hp1Model(**{f.name: getattr(spec, f.name) for f in dataclasses.fields(spec)})


def assert_close(got, ref, what, rtol=RTOL, scale=None):
    got = np.asarray(got.detach().cpu() if torch.is_tensor(got) else got, dtype=np.float64)
    ref = np.asarray(ref.detach().cpu() if torch.is_tensor(ref) else ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if scale is None:
        scale = max(float(np.abs(ref).max()) if ref.size else 1.0, 1e-30)
    err = np.abs(got - ref)
    tol = rtol * np.maximum(np.abs(ref), scale)
    bad = ~(err <= tol)
    assert not bad.any(), (
        f"{what}: {bad.sum()}/{bad.size} outside tol; max err {np.nanmax(err):.3e} at {np.argwhere(bad)[:3].tolist()} "
        f"(scale {scale:.3e})"
    )


def random_inputs(spec: MultirotorSpec, N: int, seed: int, randomize_params=True):
    """Seeded root state / actions / motor state / per-env params (CPU fp32 tensors)."""
    g = torch.Generator().manual_seed(seed)
    M = spec.num_motors
    root = torch.zeros(N, 13)
    root[:, 0:3] = torch.randn(N, 3, generator=g) * 1.5
    q = torch.randn(N, 4, generator=g)
    root[:, 3:7] = q / q.norm(dim=1, keepdim=True)
    h = N // 2  # half the envs near hover
    rp = torch.randn(h, 2, generator=g) * 0.2
    yaw = torch.rand(h, generator=g) * 6.28 - 3.14
    root[:h, 3:7] = O.quat_from_euler_xyz(rp[:, 0], rp[:, 1], yaw)
    root[:, 7:10] = torch.randn(N, 3, generator=g)
    root[:, 10:13] = torch.randn(N, 3, generator=g) * 2.0
    A = spec.num_actions
    if spec.controller == _lib.CTRL_NONE:
        actions = torch.rand(N, A, generator=g) * (spec.max_thrust - spec.min_thrust) * 1.2 + spec.min_thrust - 0.1
    else:
        actions = torch.rand(N, A, generator=g) * 2.4 - 1.2
        if spec.controller == _lib.CTRL_FULLY_ACTUATED:
            actions[:, 3:7] = torch.randn(N, 4, generator=g)
    actions[0] = 25.0  # exercises clip_actions
    u = lambda *s: torch.rand(*s, generator=g)
    lerp = lambda r, uu: (r[1] - r[0]) * uu + r[0]
    params = {
        "thrust": lerp((spec.min_thrust, spec.max_thrust), u(N, M)),
        "tau_inc": lerp(spec.tau_inc_range, u(N, M)) if randomize_params else torch.full((N, M), spec.tau_inc_range[0]),
        "tau_dec": lerp(spec.tau_dec_range, u(N, M)) if randomize_params else torch.full((N, M), spec.tau_dec_range[0]),
        "k_thrust": lerp(spec.k_thrust_range, u(N, M)),
    }
    for k, r in (("K_pos", spec.K_pos_range), ("K_vel", spec.K_vel_range), ("K_rot", spec.K_rot_range),
                 ("K_angvel", spec.K_angvel_range)):
        lo, hi = torch.tensor(r[0]), torch.tensor(r[1])
        params[k] = (hi - lo) * u(N, 3) + lo
    return root, actions, params


def load_oracle_state(model, root, params, N):
    st = O.make_state(model, N)
    st.root = root.clone()
    st.thrust = params["thrust"].clone()
    st.tau_inc, st.tau_dec, st.k_thrust = params["tau_inc"].clone(), params["tau_dec"].clone(), params["k_thrust"].clone()
    st.K_pos, st.K_vel, st.K_rot, st.K_angvel = (params[k].clone() for k in ("K_pos", "K_vel", "K_rot", "K_angvel"))
    return st


def load_engine_state(eng: Hp1Engine, root, params):
    eng.root_state.copy_(root)
    eng.motor_thrust.copy_(params["thrust"])
    for k in ("tau_inc", "tau_dec", "k_thrust", "K_pos", "K_vel", "K_rot", "K_angvel"):
        t = getattr(eng, k)
        if t is not None:
            t.copy_(params[k])


def sync_engine_from_oracle(eng: Hp1Engine, st: O.Hp1State):
    """Teacher forcing: make the engine's state bit-identical to the oracle's."""
    eng.root_state.copy_(st.root)
    eng.motor_thrust.copy_(st.thrust)
    eng.sim_steps.copy_(st.sim_steps)
    for k in ("tau_inc", "tau_dec", "k_thrust", "K_pos", "K_vel", "K_rot", "K_angvel", "bounds_min", "bounds_max"):
        t = getattr(eng, k)
        if t is not None:
            t.copy_(getattr(st, k))


# robot x controller cases shared by several tests --------------------------------------------
OCTA_ALLOC = [
    [-0.78867513, 0.21132487, -0.21132487, 0.78867513, 0.78867513, -0.21132487, 0.21132487, -0.78867513],
    [0.21132487, 0.78867513, -0.78867513, -0.21132487, -0.21132487, -0.78867513, 0.78867513, 0.21132487],
    [0.57735027, -0.57735027, -0.57735027, 0.57735027, 0.57735027, -0.57735027, -0.57735027, 0.57735027],
    [0.14226497, -0.21547005, 0.25773503, 0.01547005, -0.01547005, -0.25773503, 0.21547005, -0.14226497],
    [-0.25773503, 0.01547005, 0.14226497, 0.21547005, -0.21547005, -0.14226497, -0.01547005, 0.25773503],
    [0.11547005, -0.23094011, -0.11547005, 0.23094011, -0.23094011, 0.11547005, 0.23094011, -0.11547005],
]


def _octa_links():
    """Tilted-rotor geometry consistent with OCTA_ALLOC: thrust axis = column[0:3], position from
    the torque rows (r x ez - cq*dir*ez); used only to exercise non-identity link rotations."""
    A = np.array(OCTA_ALLOC)
    dirs = np.array([1, -1, 1, -1, 1, -1, 1, -1], dtype=float)
    r = np.zeros((8, 3))
    R = np.zeros((8, 3, 3))
    for i in range(8):
        ez = A[0:3, i] / np.linalg.norm(A[0:3, i])
        t = A[3:6, i] + 0.01 * dirs[i] * ez
        r[i] = np.cross(ez, t)  # minimum-norm solution of r x ez = t
        x = np.cross([0.0, 0.0, 1.0], ez)
        x = x / np.linalg.norm(x)
        R[i] = np.stack([x, np.cross(ez, x), ez], axis=1)
    return r, R


def spec_for(case: str) -> MultirotorSpec:
    quad = dict(num_motors=4)
    if case.startswith("quad_"):
        ctrl = {
            "quad_attitude": _lib.CTRL_ATTITUDE, "quad_position": _lib.CTRL_POSITION,
            "quad_velocity": _lib.CTRL_VELOCITY, "quad_acceleration": _lib.CTRL_ACCELERATION,
            "quad_rates": _lib.CTRL_RATES, "quad_none": _lib.CTRL_NONE,
            "quad_velocity_steering": _lib.CTRL_VELOCITY_STEERING,
        }[case]
        return MultirotorSpec(controller=ctrl, **quad)
    if case == "quadroot_attitude_euler":
        return MultirotorSpec(controller=_lib.CTRL_ATTITUDE, force_application_level="root_link", max_thrust=10.0,
                              tau_inc_range=(0.01, 0.03), tau_dec_range=(0.005, 0.005), integration_scheme="euler",
                              drag_lin1=(0.1, 0.1, 0.2), drag_lin2=(0.05, 0.05, 0.1), drag_ang1=(0.001, 0.001, 0.002),
                              drag_ang2=(0.0005, 0.0005, 0.001), **quad)
    if case in ("octa_velocity", "octa_fully_actuated", "octa_position_continuous"):
        r, R = _octa_links()
        kw = dict(
            num_motors=8, allocation_matrix=OCTA_ALLOC, motor_directions=[1, -1, 1, -1, 1, -1, 1, -1],
            link_r=r, link_R=R, use_rps=False, min_thrust=-6.25, max_thrust=6.25, mass=1.5,
            inertia=np.array([[0.021, 0.001, -0.0005], [0.001, 0.019, 0.0008], [-0.0005, 0.0008, 0.033]]),
            com=np.array([0.004, -0.003, 0.006]),
            tau_inc_range=(0.01, 0.03), tau_dec_range=(0.005, 0.005),
            K_pos_range=((2.0, 2.0, 1.0), (3.0, 3.0, 2.0)), K_vel_range=((2.0, 2.0, 2.0), (3.0, 3.0, 3.0)),
            K_rot_range=((10.8, 10.8, 5.4), (10.2, 10.2, 5.6)), K_angvel_range=((2.1, 2.1, 2.1), (2.2, 2.2, 2.2)),
            randomize_params=True,
            min_init_state=(0, 0, 0, 0, 0, -np.pi, 1.0, -0.2, -0.2, -0.2, -0.2, -0.2, -0.2),
            max_init_state=(1.0, 1.0, 1.0, 0, 0, np.pi, 1.0, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2),
        )
        if case == "octa_velocity":
            return MultirotorSpec(controller=_lib.CTRL_VELOCITY, **kw)
        if case == "octa_position_continuous":
            return MultirotorSpec(controller=_lib.CTRL_POSITION, use_discrete_approximation=False, **kw)
        return MultirotorSpec(controller=_lib.CTRL_FULLY_ACTUATED, **kw)
    raise KeyError(case)


ALL_CASES = [
    "quad_attitude", "quad_position", "quad_velocity", "quad_acceleration", "quad_rates", "quad_none",
    "quad_velocity_steering", "quadroot_attitude_euler", "octa_velocity", "octa_fully_actuated",
    "octa_position_continuous",
]


def well_conditioned(model, st, actions):
    """bool [N]: envs whose step is well conditioned in fp32.  Two places in the REFERENCE's own
    arithmetic have unbounded slope, so two correct fp32 implementations may differ by much more than
    1e-5 there (and the reference would differ from itself under a 1-ulp input change):
      * get_euler_xyz: asin(sin_pitch) near |sin_pitch| = 1 (gimbal lock), utils/math.py:135;
      * MotorModel (use_rps): sqrt(ref_thrust / k) as ref_thrust -> 0+, motor_model.py:213-242.
    Such envs (~1e-4 of uniformly random inputs) are checked for finiteness only."""
    q = st.root[:, 3:7]
    sinp = 2.0 * (q[:, 3] * q[:, 1] - q[:, 2] * q[:, 0])
    ok = sinp.abs() < 0.99995
    if model.use_rps:
        d = O.update_states(st.root)
        cmd = O.controller_wrench(model, st, d, torch.clamp(actions, -10.0, 10.0))
        if model.controller == O.CTRL_NONE:
            ref = cmd
        else:
            ref = (model.pinv_allocation(cmd.dtype) @ cmd.T).T
        thr = 5e-3 * model.max_thrust
        ok &= ~((ref > 0) & (ref < thr)).any(dim=1)
    return ok


def spec_from_registry(robot_name, controller_name):
    """the product's own config -> spec path: registries + config mirror + URDF pipeline (robots/__init__.py: make_spec)"""
    import aerial_gym_simulator_b200.robots  # noqa: F401
    from aerial_gym_simulator_b200.config.env_config import EmptyEnvCfg
    from aerial_gym_simulator_b200.config.sim_config import BaseSimConfig
    from aerial_gym_simulator_b200.registry._core import robot_registry
    robot, _ = robot_registry.make_robot(robot_name, controller_name, EmptyEnvCfg, "cpu")
    return robot.make_spec(BaseSimConfig, EmptyEnvCfg)


def check_engine_against_step_fixture(eng, spec, om, z, meta, sync=lambda: None):
    """One engine (CUDA or host shadow, built with debug_wrench=True) against a fixture recorded from the REFERENCE'S OWN
    BaseMultirotor.step (tests/golden/make_golden*.py): derived states and motor thrusts directly, link forces / torques through the
    W f reduction.  om: the oracle model (only used to replay the reference's disturbance draws)."""
    N = meta["N"]
    dev = eng.root_state.device
    T = lambda a: torch.tensor(a, device=dev)
    for s in range(meta["steps"]):
        eng.root_state.copy_(T(z[f"s{s}_root"]))
        eng.motor_thrust.copy_(T(z[f"s{s}_thrust_in"]))
        # per-env parameters: engines built with per_env_params="auto" keep a non-randomised parameter as a constant in
        # AgxHp1Config instead of an [N, .] array -- then the reference must have used that same constant
        for k in ("tau_inc", "tau_dec", "k_thrust", "K_pos", "K_vel", "K_rot", "K_angvel"):
            if k not in z:
                continue
            if getattr(eng, k, None) is not None:
                getattr(eng, k).copy_(T(z[k]))
            else:
                lo, hi = getattr(spec, k + "_range")
                assert np.allclose(lo, hi) and np.allclose(z[k], np.broadcast_to(np.asarray(lo, dtype=np.float64), z[k].shape), rtol=1e-6), \
                    f"{k}: randomised in the fixture but the engine holds a constant -- build it with per_env_params='all'"
        dist = None
        if meta["enable_disturbance"]:
            om.enable_disturbance, om.prob_apply_disturbance = True, meta["prob_apply_disturbance"]
            om.max_disturbance = tuple(meta["max_disturbance"])
            torch.manual_seed(int(z[f"s{s}_seed"]))
            dist = O.draw_disturbance(om, N).contiguous().to(dev)
        eng.physics_step(T(z[f"s{s}_actions"]).contiguous(), disturbance=dist)
        sync()
        fs = max(1.0, float(np.abs(z[f"s{s}_thrust_out"]).max()))
        assert_close(eng.motor_thrust, z[f"s{s}_thrust_out"], "thrust vs reference", scale=fs)
        ref_euler = z[f"s{s}_euler"]
        if meta["robot"] == "base_rov":  # BaseROV.update_states leaves the angles in [0, 2 pi) (base_rov.py:245, no ssa); the fused
            ref_euler = np.where(ref_euler > np.pi, ref_euler - 2.0 * np.pi, ref_euler)  # update_states wraps them (robots/__init__.py)
        assert_close(eng.euler, ref_euler, "euler vs reference", scale=np.pi)
        assert_close(eng.vehicle_orientation, z[f"s{s}_vehicle_orientation"], "veh q vs reference", scale=1.0)
        assert_close(eng.body_linvel, z[f"s{s}_body_linvel"], "body v vs reference")
        assert_close(eng.body_angvel, z[f"s{s}_body_angvel"], "body w vs reference")
        assert_close(eng.vehicle_linvel, z[f"s{s}_vehicle_linvel"], "veh v vs reference")
        Fl, Tl = z[f"s{s}_force"].astype(np.float64), z[f"s{s}_torque"].astype(np.float64)
        mask = meta["application_mask"]
        F, Tq = Fl.sum(1), Tl.sum(1)
        if spec.force_application_level == "motor_link":
            # link-frame forces: rotate into the base frame and take their moment about the COM (identity rotations for every
            # shipped robot except the tilted octarotor)
            R = np.asarray(spec.link_R, dtype=np.float64)
            Fb = np.einsum("mij,nmj->nmi", R, Fl[:, mask, :])
            Tb = np.einsum("mij,nmj->nmi", R, Tl[:, mask, :])
            rest = [b for b in range(Fl.shape[1]) if b not in mask]
            F = Fb.sum(1) + Fl[:, rest, :].sum(1)
            com = np.asarray(spec.com, dtype=np.float64)
            Tq = Tb.sum(1) + Tl[:, rest, :].sum(1) + np.cross((np.asarray(spec.link_r) - com)[None], Fb).sum(1)
            # drag / disturbance sit on body 0: the integrator spec applies them at the base-link origin
            Tq = Tq + np.cross(-com[None], Fl[:, rest, :].sum(1))
        assert_close(eng.body_wrench[:, 0:3], F, "F_body vs reference", scale=fs)
        assert_close(eng.body_wrench[:, 3:6], Tq, "T_body vs reference", scale=max(fs * 0.13, float(np.abs(Tq).max())))
