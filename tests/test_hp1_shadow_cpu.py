"""HP1 on CPU: the device code of the step / reset kernels (csrc/hp1_core.cuh + the two inline blocks of hp1_step_kernel),
compiled for the host (tests/csrc/host_shadow_hp1.inc), against
  (a) the fixtures produced by the reference's own code (no oracle in between),
  (b) the oracle on the same seeded inputs: every controller / allocation / motor-model branch, fused sub-steps, the fused
      position-task step with in-kernel Philox resets (scalar and warp-cooperative form) and the stale-observation quirk,
      resets with reference-order draws.
These are the CPU twins of tests/test_hp1_gpu.py (same helpers, same bars).  What they cannot cover: the kernels' thread
mapping, tile staging, synchronisation protocol and the device's 2-ulp division / square root (-prec-div=false) -- the -m gpu
tests do."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from aerial_gym_simulator_b200 import _lib
from aerial_gym_simulator_b200.hp1 import MultirotorSpec
from oracle import hp1_oracle as O
from oracle import philox
from tests import _hp1_common as H
from tests._models import oracle_model
from tests._shadow_hp1 import ShadowHp1Engine

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(params=["dynamic", "specialised"])
def spec_text(request):
    """which text of hp1_core.cuh's templates runs: HpSpec<-1> (every switch read from the launch constants) or the compile-time
    specialised instantiation hp1.cu's pick_kernel would launch for the configuration (same table, tests/_shadow.py lifts it)"""
    from tests import _shadow
    lib = _shadow.load()
    lib.shadow_hp1_set_specialised(int(request.param == "specialised"))
    yield request.param
    lib.shadow_hp1_set_specialised(0)


def _derived_close(eng, d, tag):
    H.assert_close(eng.euler, d["euler"], f"{tag} euler", scale=np.pi)
    H.assert_close(eng.vehicle_orientation, d["vehicle_orientation"], f"{tag} vehicle q", scale=1.0)
    H.assert_close(eng.vehicle_linvel, d["vehicle_linvel"], f"{tag} vehicle v")
    H.assert_close(eng.body_linvel, d["body_linvel"], f"{tag} body v")
    H.assert_close(eng.body_angvel, d["body_angvel"], f"{tag} body w")


@pytest.mark.parametrize("case", H.ALL_CASES)
def test_physics_step_matches_oracle(case):
    spec = H.spec_for(case)
    model = H.oracle_model_from_spec(spec)
    N = 1000
    root, actions, params = H.random_inputs(spec, N, seed=sum(map(ord, case)) % 1000)
    st = H.load_oracle_state(model, root, params, N)
    info = O.physics_step(model, st, actions)
    eng = ShadowHp1Engine(spec, N, debug_wrench=True)
    H.load_engine_state(eng, root, params)
    eng.physics_step(actions)
    _derived_close(eng, st.derived, case)
    H.assert_close(eng.motor_thrust, st.thrust, f"{case} thrust")
    H.assert_close(eng.body_wrench[:, 0:3], info["F_body"], f"{case} F_body")
    H.assert_close(eng.body_wrench[:, 3:6], info["T_body"], f"{case} T_body")
    H.assert_close(eng.root_state[:, 0:3], st.root[:, 0:3], f"{case} pos")
    H.assert_close(eng.root_state[:, 3:7], st.root[:, 3:7], f"{case} quat", scale=1.0)
    H.assert_close(eng.root_state[:, 7:10], st.root[:, 7:10], f"{case} linvel")
    H.assert_close(eng.root_state[:, 10:13], st.root[:, 10:13], f"{case} angvel")


def test_multi_substep_matches_oracle_and_single_steps():
    spec = H.spec_for("quad_velocity")
    model = H.oracle_model_from_spec(spec)
    N = 256
    root, actions, params = H.random_inputs(spec, N, seed=5)
    st = H.load_oracle_state(model, root, params, N)
    for _ in range(10):
        O.physics_step(model, st, actions)
    eng = ShadowHp1Engine(spec, N, physics_steps=10)
    H.load_engine_state(eng, root, params)
    eng.physics_step(actions)
    for nm, sl in (("pos", slice(0, 3)), ("quat", slice(3, 7)), ("linvel", slice(7, 10)), ("angvel", slice(10, 13))):
        H.assert_close(eng.root_state[:, sl], st.root[:, sl], f"10 substeps root {nm}", rtol=2e-4)
    H.assert_close(eng.body_angvel, st.derived["body_angvel"], "10 substeps stale body angvel", rtol=2e-4)
    eng1 = ShadowHp1Engine(spec, N, physics_steps=1)
    H.load_engine_state(eng1, root, params)
    for _ in range(10):
        eng1.physics_step(actions)
    assert torch.equal(eng1.root_state, eng.root_state)


@pytest.mark.parametrize("case", H.ALL_CASES)
def test_specialised_and_dynamic_text_agree_bit_for_bit(case):
    """HpSpec only removes branches: the specialised instantiation of a configuration and the generic text give the same bits
    (10 fused sub-steps, per-env parameters); and the shipped robots' configurations do have an instantiation."""
    import ctypes as C
    spec = H.spec_for(case)
    N = 512
    root, actions, params = H.random_inputs(spec, N, seed=11 + sum(map(ord, case)) % 997)
    out = []
    for on in (0, 1):
        eng = ShadowHp1Engine(spec, N, physics_steps=10, debug_wrench=True)
        eng.lib.shadow_hp1_set_specialised(on)
        try:
            H.load_engine_state(eng, root, params)
            sid = eng.lib.shadow_hp1_spec_id(C.byref(eng.cfg))
            eng.physics_step(actions)
        finally:
            eng.lib.shadow_hp1_set_specialised(0)
        out.append((sid, eng.root_state.clone(), eng.motor_thrust.clone(), eng.body_wrench.clone(), eng.body_angvel.clone()))
    assert out[0][0] == -1
    if case in ("quad_attitude", "quad_velocity", "quad_position", "quad_acceleration"):
        assert out[1][0] >= 0, "the base_quadrotor controllers have specialised kernels"
    for a, b in zip(out[0][1:], out[1][1:]):
        assert torch.equal(a, b)


STEP_FILES = sorted(glob.glob(os.path.join(GOLD, "hp1_step_*.npz")))


@pytest.mark.parametrize("path", STEP_FILES, ids=[os.path.basename(p)[9:-4] for p in STEP_FILES])
def test_device_code_matches_reference_golden(path, spec_text):
    """the kernels' arithmetic against outputs of the REFERENCE'S OWN code: derived states and motor thrusts directly, link
    forces / torques through the W f reduction (Appendix B)"""
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    om = oracle_model(meta["robot"], meta["controller"], meta["mass"], meta["inertia"])
    spec = MultirotorSpec(**{f: getattr(om, f) for f in MultirotorSpec.__dataclass_fields__})
    H.check_engine_against_step_fixture(ShadowHp1Engine(spec, meta["N"], debug_wrench=True), spec, om, z, meta)


REG_FILES = sorted(glob.glob(os.path.join(GOLD, "hp1_regstep_*.npz")))


@pytest.mark.parametrize("path", REG_FILES, ids=[os.path.basename(p)[12:-4] for p in REG_FILES])
def test_registry_built_spec_matches_reference_golden(path, spec_text):
    """as above, for 13 more robot x controller pairs (magpie, x500, lmf1, lmf2, tinyprop, base_random, morphy_stiff, octarotor;
    steering-angle controller), with the spec built by the PRODUCT's registries / config mirror / URDF pipeline"""
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    spec = H.spec_from_registry(meta["robot"], meta["controller"])
    assert abs(spec.mass - meta["mass"]) < 1e-9 and np.allclose(spec.inertia, meta["inertia"], rtol=1e-12)
    om = H.oracle_model_from_spec(spec)
    H.check_engine_against_step_fixture(ShadowHp1Engine(spec, meta["N"], debug_wrench=True), spec, om, z, meta)


def test_position_reward_block_matches_reference():
    """the reward block of hp1_step_kernel (lifted from hp1.cu) against the reference's compute_reward fixture, with the
    fixture's own stale vehicle orientation / body rates"""
    import ctypes as C

    from aerial_gym_simulator_b200.hp1 import build_config
    from tests import _shadow

    z = np.load(os.path.join(GOLD, "hp1_position_reward.npz"))
    n = z["pos"].shape[0]
    root = np.zeros((n, 13), np.float32)
    root[:, 0:3], root[:, 3:7] = z["pos"], z["quat"]
    cfg = build_config(MultirotorSpec(), n)
    qv, wb = np.ascontiguousarray(z["vehicle_orientation"], np.float32), np.ascontiguousarray(z["body_angvel"], np.float32)
    rew, cr, tr = np.zeros(n, np.float32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    steps = np.full(n, 501, np.int32)
    steps[::2] = 500
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert _shadow.load().shadow_hp1_position_reward(C.byref(cfg), n, p(root), p(qv), p(wb), None, p(steps), p(rew), p(cr), p(tr)) == 0
    ok = ~z["crashes_in"]  # the fixture also pre-sets crashes (collisions); the fused step has no such input
    assert np.array_equal(cr.astype(bool)[ok], z["crashes_out"][ok])  # bit-exact flags
    H.assert_close(rew[ok], z["reward"][ok], "reward vs reference", rtol=1e-5, scale=1.0)
    assert cr.any() and (rew[cr.astype(bool)] == -20).all()
    assert np.array_equal(tr.astype(bool), steps > 500)  # strict: sim_steps > episode_len_steps (position_setpoint_task.py:172-174)


def _philox_draws(seed, gids, episodes, M):
    d = philox.reset_uniforms(seed, gids, episodes, M)
    t = {k: torch.tensor(v) for k, v in d.items()}
    return O.ResetDraws(t["bounds_lo"], t["bounds_hi"], t["state"], t["K_pos"], t["K_vel"], t["K_rot"], t["K_angvel"],
                        t["tau_inc"], t["tau_dec"], t["thrust"], t["k_thrust"])


@pytest.mark.parametrize("case", ["quad_attitude", "octa_velocity"])
@pytest.mark.parametrize("strict,coop", [(True, False), (True, True), (False, True)], ids=["strict-scalar_rng", "strict-coop_rng", "fresh-coop_rng"])
def test_fused_position_task_step(case, strict, coop, spec_text):
    spec = H.spec_for(case)
    model = H.oracle_model_from_spec(spec)
    N, M, seed, off = 333, spec.num_motors, 99, 1000
    root, actions, params = H.random_inputs(spec, N, seed=3)
    st = H.load_oracle_state(model, root, params, N)
    g = torch.Generator().manual_seed(17)
    st.sim_steps = torch.randint(480, 501, (N,), generator=g, dtype=torch.int32)
    st.root[:5, 0:3] = 7.9
    eng = ShadowHp1Engine(spec, N, seed=seed, env_id_offset=off, device_rng_reset=True, strict_stale_obs=strict, coop_reset=coop)
    episodes = np.zeros(N, dtype=np.int64)
    target = torch.zeros(N, 3)
    n_resets = n_ill = 0
    for step in range(30):
        actions = torch.rand(N, spec.num_actions, generator=g) * 2 - 1
        H.sync_engine_from_oracle(eng, st)
        ok = H.well_conditioned(model, st, actions).numpy()
        eng.position_task_step(actions)
        draws = _philox_draws(seed, off + np.arange(N), episodes, M)
        obs, rew, term, trunc, rmask = O.position_task_step(model, st, actions, target, 500, 1, draws=draws)
        if not strict:
            st.derived = O.update_states(st.root)
            obs = O.position_task_obs(st, target)
        episodes += rmask.numpy().astype(np.int64)
        n_resets += int(rmask.sum())
        assert torch.equal(eng.terminations, term) and torch.equal(eng.truncations, trunc) and torch.equal(eng.reset_mask, rmask), step
        assert torch.equal(eng.sim_steps, st.sim_steps) and np.array_equal(eng.episode_count.numpy(), episodes), step
        H.assert_close(eng.reward, rew, f"step {step} reward", scale=1.0)
        n_ill += int((~ok).sum())
        for nm, sl in (("pos", slice(0, 3)), ("quat", slice(3, 7)), ("linvel", slice(7, 10)), ("angvel", slice(10, 13))):
            H.assert_close(eng.root_state[:, sl][ok], st.root[:, sl][ok], f"step {step} root {nm}")
            H.assert_close(eng.obs[:, sl][ok], obs[:, sl][ok], f"step {step} obs {nm}")
        H.assert_close(eng.motor_thrust[ok], st.thrust[ok], f"step {step} thrust")
        if spec.use_rps:
            H.assert_close(eng.k_thrust, st.k_thrust, f"step {step} k", scale=1e-5)
        H.assert_close(eng.tau_inc, st.tau_inc, f"step {step} tau_inc", scale=0.01)
        if spec.randomize_params:
            H.assert_close(eng.K_rot, st.K_rot, f"step {step} K_rot")
        H.assert_close(eng.body_linvel[ok], st.derived["body_linvel"][ok], f"step {step} derived body_linvel")
    assert n_resets >= N and n_ill <= 0.01 * 30 * N


def test_scalar_and_cooperative_philox_resets_are_identical():
    """device_rng_reset (reset kernel) and coop_rng_draw + apply_reset_from_tile (fused step): same blocks, same uniforms"""
    for case in ("quad_attitude", "octa_velocity"):
        spec = H.spec_for(case)
        root, actions, params = H.random_inputs(spec, 64, seed=2)
        out = []
        for coop in (False, True):
            eng = ShadowHp1Engine(spec, 64, seed=0xABCDEF0123, env_id_offset=7, episode_len_steps=0, coop_reset=coop)
            H.load_engine_state(eng, root, params)
            eng.episode_count[:] = torch.arange(64, dtype=torch.int32)
            eng.position_task_step(actions)  # episode length 0: everybody truncates and resets
            assert eng.reset_mask.all() and (eng.sim_steps == 0).all()
            out.append((eng.root_state.clone(), eng.motor_thrust.clone(), eng.tau_inc.clone(), eng.K_rot.clone(), eng.bounds_min.clone()))
        for a, b in zip(*out):
            assert torch.equal(a, b)


@pytest.mark.parametrize("case", ["quad_attitude", "octa_velocity"])
def test_reset_with_reference_order_draws(case):
    spec = H.spec_for(case)
    model = H.oracle_model_from_spec(spec)
    N = 500
    root, actions, params = H.random_inputs(spec, N, seed=21)
    st = H.load_oracle_state(model, root, params, N)
    eng = ShadowHp1Engine(spec, N, device_rng_reset=False)
    H.load_engine_state(eng, root, params)
    g = torch.Generator().manual_seed(5)
    mask = torch.rand(N, generator=g) < 0.2
    draws = O.draw_reset_uniforms(model, N, generator=g)
    before = eng.root_state.clone()
    dd = {k: (getattr(draws, k).contiguous() if getattr(draws, k) is not None else None) for k in _lib._HP1_DRAW_FIELDS}
    eng.sim_steps.fill_(7)
    eng.reset(mask, dd)
    eng.refresh()
    st.sim_steps[:] = 7
    O.reset_envs(model, st, mask, draws)
    H.assert_close(eng.root_state, st.root, "reset root", scale=1.0)
    H.assert_close(eng.motor_thrust, st.thrust, "reset thrust")
    H.assert_close(eng.tau_inc, st.tau_inc, "reset tau_inc", scale=0.01)
    H.assert_close(eng.K_angvel, st.K_angvel, "reset K_angvel")
    assert torch.equal(eng.sim_steps, st.sim_steps)
    assert torch.equal(eng.root_state[~mask], before[~mask])
    H.assert_close(eng.body_linvel, st.derived["body_linvel"], "refreshed body linvel")
    H.assert_close(eng.euler, st.derived["euler"], "refreshed euler", scale=np.pi)
