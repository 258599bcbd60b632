import sys, torch
sys.path.insert(0, '/root/repo')
from aerial_gym_simulator_b200.distributed import PipelinedObsGather
from aerial_gym_simulator_b200.hp1 import Hp1Engine, MultirotorSpec
from aerial_gym_simulator_b200 import _lib
from tests import _hp1_common as H
DEV='cuda:0'; n=65536
_lib.check(_lib.load().agx_set_spin_timeout_ms(2000), "t")
def run(with_ref, with_gather=True, lb=3, steps=46, ctas=24, nb=4, prio=-1):
    spec = H.spec_for("quad_attitude")
    root, actions, params = H.random_inputs(spec, n, seed=4)
    eng = Hp1Engine(spec, n, DEV, seed=9, materialize_derived=False)
    ref = Hp1Engine(spec, n, DEV, seed=9, materialize_derived=False)
    for e in (eng, ref):
        H.load_engine_state(e, root, params)
        e.sim_steps.copy_((torch.arange(n, device=DEV) % 500 + 480).int() % 501)
    gth = PipelinedObsGather(n, 13, DEV, num_buffers=nb, loopback_world=lb, max_ctas=ctas)
    if prio != -1:
        gth.streams = [torch.cuda.Stream(device=DEV) for _ in range(nb)]; gth._raw = [st.cuda_stream for st in gth.streams]
    if with_gather: eng.attach_obs_gather(gth)
    act = actions.to(DEV)
    for step in range(steps):
        if with_ref: ref.position_task_step(act)
        eng.position_task_step(act)
    if with_gather:
        gth.loopback_complete(gth.epoch); gth.wait()
    torch.cuda.synchronize()
    ar = eng.any_reset.cpu().tolist()
    ts = eng.tile_sync.cpu()
    nt = n // 32
    print("ctas", ctas, "nb", nb, "prio", prio, "with_ref", with_ref, "gather", with_gather, "err", hex(gth.error_word.item()), "throttled", gth.throttled, "T", eng._chain_T,
          "flags", gth.flags.view(4,16)[:, 0].tolist(), "arrivals", [ar[8+2*k] + (ar[9+2*k] << 32) for k in range(4)], "stepflags", ar[4:8], "engerr", hex(ar[2]),
          "claim min/max", ts[:nt].min().item(), ts[:nt].max().item(), "done min/max", ts[nt:].min().item(), ts[nt:].max().item(), flush=True)
run(False); run(True); run(False, ctas=37); run(False, ctas=16)
