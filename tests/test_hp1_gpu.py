"""HP1 parity tests proper: the CUDA path, called through the C ABI, against the oracle on the
same seeded inputs, against the reference-generated golden fixtures, and size-independent
properties at BASELINE.json's full env count."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import hp1_oracle as O
from oracle import philox
from aerial_gym_simulator_b200 import _lib
from aerial_gym_simulator_b200.hp1 import Hp1Engine, MultirotorSpec
from tests import _hp1_common as H
from tests._models import oracle_model

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def _derived_close(eng, d, tag):
    H.assert_close(eng.euler, d["euler"], f"{tag} euler", scale=np.pi)
    H.assert_close(eng.vehicle_orientation, d["vehicle_orientation"], f"{tag} vehicle q", scale=1.0)
    H.assert_close(eng.vehicle_linvel, d["vehicle_linvel"], f"{tag} vehicle v")
    H.assert_close(eng.body_linvel, d["body_linvel"], f"{tag} body v")
    H.assert_close(eng.body_angvel, d["body_angvel"], f"{tag} body w")


@pytest.mark.parametrize("case", H.ALL_CASES)
def test_physics_step_matches_oracle(case):
    """a1-a13: one physics step from identical inputs, every controller / allocation mode /
    motor-model branch.  N = 1000 also covers the ragged last tile (1000 = 31*32 + 8)."""
    spec = H.spec_for(case)
    model = H.oracle_model_from_spec(spec)
    N = 1000
    root, actions, params = H.random_inputs(spec, N, seed=sum(map(ord, case)) % 1000)
    st = H.load_oracle_state(model, root, params, N)
    info = O.physics_step(model, st, actions)
    eng = Hp1Engine(spec, N, DEV, per_env_params="all", debug_wrench=True)
    H.load_engine_state(eng, root, params)
    eng.physics_step(actions.to(DEV))
    torch.cuda.synchronize()
    _derived_close(eng, st.derived, case)
    H.assert_close(eng.motor_thrust, st.thrust, f"{case} thrust")
    H.assert_close(eng.body_wrench[:, 0:3], info["F_body"], f"{case} F_body")
    H.assert_close(eng.body_wrench[:, 3:6], info["T_body"], f"{case} T_body")
    H.assert_close(eng.root_state[:, 0:3], st.root[:, 0:3], f"{case} pos")
    H.assert_close(eng.root_state[:, 3:7], st.root[:, 3:7], f"{case} quat", scale=1.0)
    H.assert_close(eng.root_state[:, 7:10], st.root[:, 7:10], f"{case} linvel")
    H.assert_close(eng.root_state[:, 10:13], st.root[:, 10:13], f"{case} angvel")


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 64, 4097])
def test_ragged_sizes(n):
    spec = H.spec_for("quad_attitude")
    model = H.oracle_model_from_spec(spec)
    eng = Hp1Engine(spec, n, DEV, per_env_params="all")
    if n == 0:
        eng.physics_step(torch.zeros(0, 4, device=DEV))
        return
    root, actions, params = H.random_inputs(spec, n, seed=n)
    st = H.load_oracle_state(model, root, params, n)
    O.physics_step(model, st, actions)
    H.load_engine_state(eng, root, params)
    eng.physics_step(actions.to(DEV))
    H.assert_close(eng.root_state, st.root, f"n={n} root", scale=1.0)
    H.assert_close(eng.motor_thrust, st.thrust, f"n={n} thrust")


def test_multi_substep_matches_oracle():
    """env_with_obstacles runs 10 physics steps per env step (env_with_obstacles.py:29):
    fused sub-stepping must equal 10 single steps."""
    spec = H.spec_for("quad_velocity")
    model = H.oracle_model_from_spec(spec)
    N = 512
    root, actions, params = H.random_inputs(spec, N, seed=5)
    st = H.load_oracle_state(model, root, params, N)
    for _ in range(10):
        O.physics_step(model, st, actions)
    eng = Hp1Engine(spec, N, DEV, per_env_params="all", physics_steps=10)
    H.load_engine_state(eng, root, params)
    eng.physics_step(actions.to(DEV))
    # ten free-running steps: per-step differences (<= 1e-5 of each quantity's scale) compound through
    # the attitude loop, so the bar here is 2e-4 of scale; single-step parity is tested above
    for nm, sl in (("pos", slice(0, 3)), ("quat", slice(3, 7)), ("linvel", slice(7, 10)), ("angvel", slice(10, 13))):
        H.assert_close(eng.root_state[:, sl], st.root[:, sl], f"10 substeps root {nm}", rtol=2e-4)
    H.assert_close(eng.motor_thrust, st.thrust, "10 substeps thrust", rtol=2e-4)
    H.assert_close(eng.body_angvel, st.derived["body_angvel"], "10 substeps stale body angvel", rtol=2e-4)
    H.assert_close(eng.euler, st.derived["euler"], "10 substeps stale euler", rtol=2e-4, scale=np.pi)
    # and equals ten 1-step launches bit for bit (same kernel arithmetic)
    eng1 = Hp1Engine(spec, N, DEV, per_env_params="all", physics_steps=1)
    H.load_engine_state(eng1, root, params)
    a = actions.to(DEV)
    for _ in range(10):
        eng1.physics_step(a)
    assert torch.equal(eng1.root_state, eng.root_state)


STEP_FILES = sorted(glob.glob(os.path.join(GOLD, "hp1_step_*.npz")))


@pytest.mark.parametrize("path", STEP_FILES, ids=[os.path.basename(p)[9:-4] for p in STEP_FILES])
def test_kernel_matches_reference_golden(path):
    """The CUDA path against outputs of the REFERENCE'S OWN code (no oracle in between):
    derived states and motor thrusts directly; link forces/torques through the W f reduction."""
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    om = oracle_model(meta["robot"], meta["controller"], meta["mass"], meta["inertia"])
    spec = MultirotorSpec(**{f: getattr(om, f) for f in MultirotorSpec.__dataclass_fields__})
    N = meta["N"]
    eng = Hp1Engine(spec, N, DEV, per_env_params="all", debug_wrench=True)
    T = lambda a: torch.tensor(a, device=DEV)
    for s in range(meta["steps"]):
        eng.root_state.copy_(T(z[f"s{s}_root"]))
        eng.motor_thrust.copy_(T(z[f"s{s}_thrust_in"]))
        eng.tau_inc.copy_(T(z["tau_inc"]))
        eng.tau_dec.copy_(T(z["tau_dec"]))
        if "k_thrust" in z and eng.k_thrust is not None:
            eng.k_thrust.copy_(T(z["k_thrust"]))
        if "K_pos" in z:
            eng.K_pos.copy_(T(z["K_pos"])); eng.K_vel.copy_(T(z["K_vel"]))
            eng.K_rot.copy_(T(z["K_rot"])); eng.K_angvel.copy_(T(z["K_angvel"]))
        dist = None
        if meta["enable_disturbance"]:
            om.enable_disturbance, om.prob_apply_disturbance = True, meta["prob_apply_disturbance"]
            om.max_disturbance = tuple(meta["max_disturbance"])
            torch.manual_seed(int(z[f"s{s}_seed"]))
            dist = O.draw_disturbance(om, N).to(DEV).contiguous()  # torch draws, reference order
        eng.physics_step(T(z[f"s{s}_actions"]).contiguous(), disturbance=dist)
        fs = max(1.0, float(np.abs(z[f"s{s}_thrust_out"]).max()))
        H.assert_close(eng.motor_thrust, z[f"s{s}_thrust_out"], "thrust vs reference", scale=fs)
        H.assert_close(eng.euler, z[f"s{s}_euler"], "euler vs reference", scale=np.pi)
        H.assert_close(eng.vehicle_orientation, z[f"s{s}_vehicle_orientation"], "veh q vs reference", scale=1.0)
        H.assert_close(eng.body_linvel, z[f"s{s}_body_linvel"], "body v vs reference")
        H.assert_close(eng.body_angvel, z[f"s{s}_body_angvel"], "body w vs reference")
        H.assert_close(eng.vehicle_linvel, z[f"s{s}_vehicle_linvel"], "veh v vs reference")
        # reduce the reference's per-link tensors to the base-frame wrench (Appendix B, R_i = I here)
        Fl, Tl = z[f"s{s}_force"].astype(np.float64), z[f"s{s}_torque"].astype(np.float64)
        mask = meta["application_mask"]
        F = Fl.sum(1)
        Tq = Tl.sum(1)
        if spec.force_application_level == "motor_link":
            r = spec.link_r - spec.com
            Tq = Tq + np.cross(r[None], Fl[:, mask, :]).sum(1)
        H.assert_close(eng.body_wrench[:, 0:3], F, "F_body vs reference", scale=fs)
        H.assert_close(eng.body_wrench[:, 3:6], Tq, "T_body vs reference", scale=max(fs * 0.13, float(np.abs(Tq).max())))


def _philox_draws(seed, gids, episodes, M):
    d = philox.reset_uniforms(seed, gids, episodes, M)
    t = {k: torch.tensor(v) for k, v in d.items()}
    return O.ResetDraws(t["bounds_lo"], t["bounds_hi"], t["state"], t["K_pos"], t["K_vel"], t["K_rot"], t["K_angvel"],
                        t["tau_inc"], t["tau_dec"], t["thrust"], t["k_thrust"])


@pytest.mark.parametrize("case", ["quad_attitude", "octa_velocity"])
@pytest.mark.parametrize("strict,materialize,two_launch",
                         [(True, True, True), (True, False, True), (True, True, False), (True, False, False), (False, True, False)],
                         ids=["strict-refresh_pass", "strict-obs_patch_pass", "strict-cooperative-derived", "strict-cooperative", "fresh"])
def test_fused_position_task_step(case, strict, materialize, two_launch):
    """Whole PositionSetpointTask.step in one C-ABI call, device-RNG resets, teacher-forced for
    30 steps: reward/obs to 1e-5, termination / truncation / reset masks and sim_steps bit-exact,
    the stale-derived-state quirk reproduced (strict: in the single cooperative launch, or -- the
    path grids larger than one resident wave take, forced here by asking for the mid-step event --
    by the refresh / obs-patch pass) or disabled."""
    spec = H.spec_for(case)
    model = H.oracle_model_from_spec(spec)
    N, M, seed, off = 777, spec.num_motors, 99, 1000
    root, actions, params = H.random_inputs(spec, N, seed=3)
    st = H.load_oracle_state(model, root, params, N)
    g = torch.Generator().manual_seed(17)
    st.sim_steps = torch.randint(480, 501, (N,), generator=g, dtype=torch.int32)  # truncations within 30 steps
    st.root[:5, 0:3] = 7.9  # |x| > 8 soon -> crashes
    eng = Hp1Engine(spec, N, DEV, per_env_params="all", seed=seed, env_id_offset=off, device_rng_reset=True,
                    strict_stale_obs=strict, materialize_derived=materialize)
    episodes = np.zeros(N, dtype=np.int64)
    target = torch.zeros(N, 3)
    n_resets = n_ill = 0
    mid = None
    if two_launch:
        mid = torch.cuda.Event(enable_timing=True)
        mid.record()
    for step in range(30):
        actions = torch.rand(N, spec.num_actions, generator=g) * 2 - 1
        H.sync_engine_from_oracle(eng, st)
        # conditioning: asin() in get_euler_xyz has slope 1/sqrt(1-x^2); within ~0.6 deg of gimbal lock a
        # 1-ulp difference in sin(pitch) moves the Euler angles (and the torque built on them) by far
        # more than 1e-5 in ANY two fp32 implementations.  Such envs (expected ~1 per 10^4 uniformly
        # random attitudes) are checked for boundedness only.
        ok = H.well_conditioned(model, st, actions).numpy()
        eng.position_task_step(actions.to(DEV), mid_event=mid)
        draws = _philox_draws(seed, off + np.arange(N), episodes, M)
        st_ref = st
        obs, rew, term, trunc, rmask = O.position_task_step(model, st_ref, actions, target, 500, 1, draws=draws)
        if not strict:  # non-strict mode: derived states always refreshed before the observation
            st.derived = O.update_states(st.root)
            obs = O.position_task_obs(st, target)
        episodes += rmask.numpy().astype(np.int64)
        n_resets += int(rmask.sum())
        torch.cuda.synchronize()
        assert torch.equal(eng.terminations.cpu(), term), f"step {step} terminations"
        assert torch.equal(eng.truncations.cpu(), trunc), f"step {step} truncations"
        assert torch.equal(eng.reset_mask.cpu(), rmask), f"step {step} reset mask"
        assert torch.equal(eng.sim_steps.cpu(), st.sim_steps), f"step {step} sim_steps"
        assert np.array_equal(eng.episode_count.cpu().numpy(), episodes), f"step {step} episode counters"
        H.assert_close(eng.reward, rew, f"step {step} reward", scale=1.0)
        assert torch.isfinite(eng.root_state).all() and torch.isfinite(eng.motor_thrust).all()
        n_ill += int((~ok).sum())
        for nm, sl in (("pos", slice(0, 3)), ("quat", slice(3, 7)), ("linvel", slice(7, 10)), ("angvel", slice(10, 13))):
            H.assert_close(eng.root_state[:, sl].cpu()[ok], st.root[:, sl][ok], f"step {step} root {nm}")
            H.assert_close(eng.obs[:, sl].cpu()[ok], obs[:, sl][ok], f"step {step} obs {nm}")
        H.assert_close(eng.motor_thrust.cpu()[ok], st.thrust[ok], f"step {step} thrust")
        if eng.k_thrust is not None and spec.use_rps:
            H.assert_close(eng.k_thrust, st.k_thrust, f"step {step} k", scale=1e-5)
        H.assert_close(eng.tau_inc, st.tau_inc, f"step {step} tau_inc", scale=0.01)
        if spec.randomize_params:
            H.assert_close(eng.K_rot, st.K_rot, f"step {step} K_rot")
        if materialize:  # derived arrays follow the same stale / refreshed rule as the observation
            H.assert_close(eng.body_linvel.cpu()[ok], st.derived["body_linvel"][ok], f"step {step} derived body_linvel")
            H.assert_close(eng.body_angvel.cpu()[ok], st.derived["body_angvel"][ok], f"step {step} derived body_angvel")
    assert n_resets >= N  # every env truncated at least once in the window
    assert n_ill <= 0.01 * 30 * N  # ill-conditioned samples are rare


def test_stale_observation_quirk():
    """velocity slots of the observation are one physics step
    stale unless ANY env reset that step, in which case ALL envs are refreshed."""
    spec = H.spec_for("quad_attitude")
    N = 256
    root, actions, params = H.random_inputs(spec, N, seed=8)
    eng = Hp1Engine(spec, N, DEV, per_env_params="all")
    H.load_engine_state(eng, root, params)
    pre = O.update_states(root)
    eng.position_task_step(actions.to(DEV))  # sim_steps = 1: no reset anywhere (positions < 8 m)
    assert not eng.reset_mask.any()
    H.assert_close(eng.obs[:, 7:10], pre["body_linvel"], "stale body linvel")
    fresh = O.update_states(eng.root_state.cpu())
    assert (eng.obs[:, 7:10].cpu() - fresh["body_linvel"]).abs().max() > 1e-3
    # now force ONE env to truncate: everybody's observation becomes fresh
    eng.sim_steps[17] = 500
    eng.position_task_step(actions.to(DEV))
    assert eng.reset_mask.sum().item() == 1 and bool(eng.reset_mask[17])
    fresh = O.update_states(eng.root_state.cpu())
    H.assert_close(eng.obs[:, 7:10], fresh["body_linvel"], "fresh body linvel after a reset elsewhere")
    H.assert_close(eng.obs[:, 10:13], fresh["body_angvel"], "fresh body angvel after a reset elsewhere")
    assert int(eng.any_reset[0]) == 0 and int(eng.any_reset[1]) == 0  # flag consumed
    H.assert_close(eng.body_linvel, fresh["body_linvel"], "derived array refreshed as well")
    ar = eng.any_reset.cpu().tolist()
    # single-launch path: 8 warps counted in per step (64-bit counters of steps 0 and 1 at [8], [10]); step 0's flag
    # never rose, step 1's holds its tag (T + 1 = 2); every tile claimed and published two steps
    assert ar[8] == N // 32 and ar[10] == N // 32 and ar[4] == 0 and ar[5] == 2
    ts = eng.tile_sync.cpu().tolist()
    assert ts == [2] * (2 * (N // 32))


def test_host_io_step_is_identical():
    """Host I/O mode: the kernel reads actions from and writes obs / reward / flags to pinned,
    device-mapped host memory.  Same bits as the device-buffer engine, step after step."""
    spec = H.spec_for("quad_attitude")
    N = 4099
    root, actions, params = H.random_inputs(spec, N, seed=12)
    engs = [Hp1Engine(spec, N, DEV, seed=5, materialize_derived=False, host_io=h) for h in (False, True)]
    g = torch.Generator().manual_seed(2)
    steps0 = torch.randint(490, 501, (N,), generator=g, dtype=torch.int32)
    for e in engs:
        H.load_engine_state(e, root, params)
        e.sim_steps.copy_(steps0.to(DEV))
    host = engs[1]
    assert host.obs.device.type == "cpu" and host.obs.is_pinned()
    pinned = torch.empty(N, spec.num_actions).pin_memory()
    for step in range(25):
        a = torch.rand(N, spec.num_actions, generator=g) * 2 - 1
        engs[0].position_task_step(a.to(DEV))
        if step % 2:  # the engine's own mapped buffer ...
            host.host_actions.copy_(a)
            host.position_task_step(host.host_actions)
        else:         # ... or any pinned tensor
            pinned.copy_(a)
            host.position_task_step(pinned)
        torch.cuda.synchronize()
        assert torch.equal(engs[0].obs.cpu(), host.obs), f"step {step} obs"
        assert torch.equal(engs[0].reward.cpu(), host.reward), f"step {step} reward"
        assert torch.equal(engs[0].terminations.cpu(), host.terminations)
        assert torch.equal(engs[0].truncations.cpu(), host.truncations)
        assert torch.equal(engs[0].root_state, host.root_state)
    assert host.truncations.any() or engs[0].episode_count.sum() > 0
    with pytest.raises(ValueError):
        host.position_task_step(torch.zeros(N, spec.num_actions))  # pageable host memory is refused
    host.close()


@pytest.mark.parametrize("case", ["quad_attitude", "octa_velocity"])
def test_reset_with_reference_order_draws(case):
    """a15: agx_hp1_reset with torch uniforms drawn in the reference's call order, followed by
    the all-env refresh; untouched rows stay bit-identical."""
    spec = H.spec_for(case)
    model = H.oracle_model_from_spec(spec)
    N = 500
    root, actions, params = H.random_inputs(spec, N, seed=21)
    st = H.load_oracle_state(model, root, params, N)
    eng = Hp1Engine(spec, N, DEV, per_env_params="all", device_rng_reset=False)
    H.load_engine_state(eng, root, params)
    g = torch.Generator().manual_seed(5)
    mask = torch.rand(N, generator=g) < 0.2
    draws = O.draw_reset_uniforms(model, N, generator=g)
    before = eng.root_state.clone()
    dd = {k: (getattr(draws, k).to(DEV).contiguous() if getattr(draws, k) is not None else None)
          for k in _lib._HP1_DRAW_FIELDS}
    eng.sim_steps.fill_(7)
    eng.reset(mask.to(DEV), dd)
    eng.refresh()
    st.sim_steps[:] = 7
    O.reset_envs(model, st, mask, draws)
    H.assert_close(eng.root_state, st.root, "reset root", scale=1.0)
    H.assert_close(eng.motor_thrust, st.thrust, "reset thrust")
    H.assert_close(eng.tau_inc, st.tau_inc, "reset tau_inc", scale=0.01)
    H.assert_close(eng.K_angvel, st.K_angvel, "reset K_angvel")
    assert torch.equal(eng.sim_steps.cpu(), st.sim_steps)
    keep = ~mask
    assert torch.equal(eng.root_state.cpu()[keep], before.cpu()[keep])
    H.assert_close(eng.body_linvel, st.derived["body_linvel"], "refreshed body linvel")
    H.assert_close(eng.euler, st.derived["euler"], "refreshed euler", scale=np.pi)


def test_full_size_properties():
    """BASELINE config #2 size (65,536 envs): size-independent properties over 200 fused steps --
    unit quaternions, finite state, run-to-run bit determinism, sharding invariance (two half-size
    engines with env_id_offset reproduce the single-engine result bit for bit)."""
    spec = H.spec_for("quad_attitude")
    N = 65536
    g = torch.Generator().manual_seed(0)
    acts = [(torch.rand(N, 4, generator=g) * 2 - 1).to(DEV) for _ in range(8)]

    def run(n, off, sl):
        eng = Hp1Engine(spec, n, DEV, seed=42, env_id_offset=off, materialize_derived=False)
        eng.reset(torch.ones(n, dtype=torch.bool, device=DEV))
        eng.refresh()
        eng.sim_steps.copy_((torch.arange(off, off + n, device=DEV) % 500).int())  # staggered truncations
        tot_reset = 0
        for i in range(200):
            eng.position_task_step(acts[i % 8][sl].contiguous())
            tot_reset += int(eng.reset_mask.sum())
        torch.cuda.synchronize()
        return eng, tot_reset

    e1, r1 = run(N, 0, slice(0, N))
    assert torch.isfinite(e1.root_state).all() and torch.isfinite(e1.obs).all() and torch.isfinite(e1.reward).all()
    qn = e1.root_state[:, 3:7].norm(dim=1)
    assert (qn - 1).abs().max() < 1e-5
    assert r1 > N * 0.3  # truncations happened
    assert (e1.sim_steps >= 0).all() and (e1.sim_steps <= 501).all()
    e2, r2 = run(N, 0, slice(0, N))
    assert r1 == r2 and torch.equal(e1.root_state, e2.root_state) and torch.equal(e1.obs, e2.obs)
    ea, ra = run(N // 2, 0, slice(0, N // 2))
    eb, rb = run(N // 2, N // 2, slice(N // 2, N))
    # the obs-refresh quirk couples envs globally, so compare the physics state + bookkeeping
    assert torch.equal(torch.cat([ea.root_state, eb.root_state]), e1.root_state)
    assert torch.equal(torch.cat([ea.sim_steps, eb.sim_steps]), e1.sim_steps)
    assert ra + rb == r1


def test_error_paths():
    spec = H.spec_for("quad_attitude")
    eng = Hp1Engine(spec, 64, DEV)
    with pytest.raises(ValueError):
        eng.physics_step(torch.zeros(63, 4, device=DEV))
    with pytest.raises(ValueError):
        eng.physics_step(torch.zeros(64, 4, device=DEV, dtype=torch.float64))
    eng.cfg.num_motors = 5
    with pytest.raises(_lib.AgxError, match="num_motors"):
        eng.physics_step(torch.zeros(64, 4, device=DEV))
    with pytest.raises(_lib.AgxError):
        Hp1Engine(spec, 4, "cpu")


@pytest.mark.parametrize("n,two_launch", [(4096, False), (776, False), (4096, True), (65536, False)])
def test_pipelined_obs_gather_loopback(n, two_launch):
    """The observation all-gather beside the chained steps, on ONE GPU with an emulated world of 3 (every "peer" buffer is a local
    tensor): the step writes its rows into the ring slot, the push kernel (side stream, waiting on the step's completion counter in
    device memory -- or in stream order on the two-launch path) copies them into every peer's buffer and publishes the epoch;
    a step waits (stream event) only for the push that last read its ring slot.  Results equal those of an engine without a gather, bit for bit."""
    from aerial_gym_simulator_b200.distributed import PipelinedObsGather

    spec = H.spec_for("quad_attitude")
    root, actions, params = H.random_inputs(spec, n, seed=4)
    ref = Hp1Engine(spec, n, DEV, seed=9, materialize_derived=False)
    eng = Hp1Engine(spec, n, DEV, seed=9, materialize_derived=False)
    for e in (ref, eng):
        H.load_engine_state(e, root, params)
        e.sim_steps.copy_((torch.arange(n, device=DEV) % 500 + 480).int() % 501)
    gth = PipelinedObsGather(n, 13, DEV, num_buffers=4, loopback_world=3)
    eng.attach_obs_gather(gth)
    mid = None
    if two_launch:
        mid = torch.cuda.Event(enable_timing=True)
        mid.record()
    act = actions.to(DEV)
    # (1) synchronous use: wait for the epoch after every step
    for step in range(6):
        ref.position_task_step(act, mid_event=mid)
        eng.position_task_step(act, mid_event=mid)
        gth.loopback_complete(gth.epoch)  # the emulated peers "arrive"
        got = gth.wait()
        torch.cuda.synchronize()
        assert got is eng.gathered_obs and got is gth.outs[(step + 1) % 4]
        assert torch.equal(eng.obs, ref.obs), f"step {step}"
        assert torch.equal(got[:n], ref.obs)
        for peer in (1, 2):  # rank 0's rows landed in slot 0 of every peer's buffer
            assert torch.equal(gth.peer_outs[(step + 1) % 4][peer][:n], ref.obs), f"step {step} peer {peer}"
        fw = ((step + 1) % 4) * 16  # flag words of this epoch's ring slot
        assert gth.flags[fw].item() == step + 1 and gth.peer_flags[1][fw].item() == step + 1 and gth.peer_flags[2][fw].item() == step + 1
        assert gth.scratch.sum().item() == 0
    # (2) free running: 40 steps enqueued back to back, pushes overlap the following steps, the ring throttles
    for step in range(40):
        ref.position_task_step(act, mid_event=mid)
        eng.position_task_step(act, mid_event=mid)
    gth.loopback_complete(gth.epoch)
    got = gth.wait()
    gth.check()
    eng.check()
    assert torch.equal(got[:n], ref.obs) and torch.equal(eng.root_state, ref.root_state)
    assert torch.equal(gth.peer_outs[gth.epoch % 4][2][:n], ref.obs)
    assert gth.flags[(46 % 4) * 16].item() == 46 and gth.flags[(45 % 4) * 16].item() == 45
    # (3) detach: the engine owns its observation buffer again
    eng.attach_obs_gather(None)
    ref.position_task_step(act)
    eng.position_task_step(act)
    torch.cuda.synchronize()
    assert eng.obs is eng._own_obs and torch.equal(eng.obs, ref.obs)


def test_chain_counters_are_rebased_before_they_wrap():
    """the 32-bit step index / flags / per-tile counters of the chained step are zeroed at a quiet point every CHAIN_REBASE_STEPS
    launches (they would wrap after 2^32 steps); a rebase in the middle of a rollout changes nothing"""
    spec = H.spec_for("quad_attitude")
    n = 4096
    root, actions, params = H.random_inputs(spec, n, seed=6)
    a, b = Hp1Engine(spec, n, DEV, seed=2, materialize_derived=False), Hp1Engine(spec, n, DEV, seed=2, materialize_derived=False)
    for e in (a, b):
        H.load_engine_state(e, root, params)
        e.sim_steps.copy_((torch.arange(n, device=DEV) % 500).int())
    b.CHAIN_REBASE_STEPS = 7
    act = actions.to(DEV)
    for step in range(30):
        a.position_task_step(act)
        b.position_task_step(act)
    torch.cuda.synchronize()
    assert b._chain_T == 30 % 7 and a._chain_T == 30
    assert torch.equal(a.root_state, b.root_state) and torch.equal(a.obs, b.obs) and torch.equal(a.sim_steps, b.sim_steps)
    assert b.tile_sync.max().item() == b._chain_T
    a.check(); b.check()


def test_chained_step_wait_times_out_instead_of_trapping():
    """A wait that can never be satisfied (a tile whose done-counter is ahead of its claim counter) expires by wall clock, the step
    goes on, agx_hp1_check reports AGX_E_TIMEOUT and the CUDA context survives (no __trap on a late producer)."""
    from aerial_gym_simulator_b200 import _lib

    lib = _lib.load()
    spec = H.spec_for("quad_attitude")
    n = 2048
    eng = Hp1Engine(spec, n, DEV, seed=1, materialize_derived=False)
    eng.reset(torch.ones(n, dtype=torch.bool, device=DEV))
    _lib.check(lib.agx_set_spin_timeout_ms(50), "agx_set_spin_timeout_ms")
    try:
        eng.tile_sync[n // 32 + 5] = 7  # tile 5 "has published step 7": its step-0 warp waits for a done-counter of 0 that never comes
        eng.position_task_step(torch.zeros(n, 4, device=DEV))
        with pytest.raises(_lib.AgxError, match="timed out"):
            eng.check()
        assert int(eng.any_reset[2]) & 15 in (1, 3)  # code 1: a tile's previous step never published (3: the reset decision of a later step never closed)
    finally:
        _lib.check(lib.agx_set_spin_timeout_ms(20000), "agx_set_spin_timeout_ms")
    # the context is alive and a fresh engine steps normally
    eng2 = Hp1Engine(spec, n, DEV, seed=1, materialize_derived=False)
    eng2.reset(torch.ones(n, dtype=torch.bool, device=DEV))
    eng2.position_task_step(torch.zeros(n, 4, device=DEV))
    eng2.check()
