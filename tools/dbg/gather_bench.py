"""torchrun micro-benchmark of bench.py's timed loop (16 rotating replicas of the 65,536-env step, pipelined observation gather) for a
list of push-kernel sizes: python -m torch.distributed.run --nproc-per-node N tools/dbg/gather_bench.py 16 24 48"""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from aerial_gym_simulator_b200.distributed import PipelinedObsGather
from aerial_gym_simulator_b200.hp1 import Hp1Engine, MultirotorSpec
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N, R, K = 65536, 16, 200
engines = []
for rep in range(R):
    e = Hp1Engine(MultirotorSpec(), N, dev, seed=1 + rep, env_id_offset=rank * N, materialize_derived=False)
    e.reset(torch.ones(N, dtype=torch.bool, device=dev)); e.refresh()
    e.sim_steps.copy_((torch.arange(N, device=dev) % 500).int())
    engines.append(e)
g = torch.Generator(device=dev).manual_seed(1234 + rank)
acts = [(torch.rand(N, 4, generator=g, device=dev) * 2 - 1).contiguous() for _ in range(8)]
def run(gather, label):
    for e in engines: e.attach_obs_gather(gather)
    def loop(k):
        for i in range(k): engines[i % R].position_task_step(acts[i % 8])
        if gather is not None: gather.fence(); gather.wait()
    loop(32); dist.barrier(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); e0.record(); loop(K); e1.record(); th = time.perf_counter() - t0
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / K], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t))
    errs = []
    for e in engines:
        w = int(e.any_reset[2].item())
        if w: errs.append(hex(w))
    gerr = hex(int(gather.error_word.item())) if gather is not None else None
    if errs or (gerr not in (None, "0x0")):
        print(f"[rank {rank}] {label}: engine error words {errs[:4]} gather error word {gerr}", flush=True)
        for e in engines: e.any_reset[2] = 0
        if gather is not None: gather.error_word.zero_()
    if rank == 0:
        print(f"GATHER_BENCH world={world} {label}: {best:.2f} us/step  ({world * N / best * 1e6:.3e} env-steps/s, host {th * 1e6 / K:.1f} us/step, nvlink floor {(world - 1) * N * 52 / 900e9 * 1e6:.1f} us)", flush=True)
    for e in engines: e.attach_obs_gather(None)
FRESH = os.environ.get("FRESH", "0") == "1"  # attach the gather to engines that never stepped (their first steps are "no reset yet" steps)
if not FRESH:
    run(None, "no gather")
for spec in sys.argv[1:]:
    ctas, nb = (int(x) for x in (spec.split("x") + ["4"])[:2])
    for mc in ((True, False) if os.environ.get("MC_AB", "0") == "1" else (None,)):
        gth = PipelinedObsGather(N, 13, dev, num_buffers=nb, max_ctas=ctas, multicast=mc)
        run(gth, f"push ctas={ctas} ring={nb} multicast={gth.multicast}")
dist.destroy_process_group()
