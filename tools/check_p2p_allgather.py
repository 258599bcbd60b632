"""torchrun self-check of the NVLink observation all-gathers against NCCL (needs >= 2 GPUs):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_p2p_allgather.py

1. agx_p2p_allgather (one synchronous kernel) == all_gather_into_tensor, bit for bit;
2. the pipelined gather beside the chained HP1 steps (Hp1Engine.attach_obs_gather + PipelinedObsGather):
   a. synchronous use (wait after every step): the gathered buffer == NCCL's gather of the same step's observation;
   b. free running (K steps enqueued back to back, pushes overlapping the next steps): the last gathered buffer == NCCL's,
      no wait timed out; and the step period of both modes, next to step-then-NCCL and step alone."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aerial_gym_simulator_b200.distributed import ObsAllGather, P2PObsAllGather, PipelinedObsGather  # noqa: E402
from aerial_gym_simulator_b200.hp1 import Hp1Engine, MultirotorSpec  # noqa: E402

rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N = int(os.environ.get("N_ENVS", "65536"))
ITERS = int(os.environ.get("ITERS", "200"))
p2p = P2PObsAllGather(N, 13, dev)
nccl = ObsAllGather(N, 13, world * N, dev)
g = torch.Generator(device=dev).manual_seed(100 + rank)
ok = True
for it in range(20):
    x = torch.rand(N, 13, device=dev, generator=g)
    a = p2p(x).clone()
    b = nccl(x).clone()
    ok &= bool(torch.equal(a, b))
torch.cuda.synchronize()
x = torch.rand(N, 13, device=dev, generator=g)


def timeit(fn, iters=ITERS, tail=None):
    for _ in range(10):
        fn(x)
    if tail:
        tail()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn(x)
    if tail:
        tail()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


tp, tn = timeit(p2p), timeit(nccl)

spec = MultirotorSpec()


def make_engine(seed):
    e = Hp1Engine(spec, N, dev, seed=seed, env_id_offset=rank * N, materialize_derived=False)
    e.reset(torch.ones(N, dtype=torch.bool, device=dev))
    e.refresh()
    e.sim_steps.copy_((torch.arange(N, device=dev) % 500).int())
    return e


eng = make_engine(3)
pg = PipelinedObsGather(N, 13, dev, num_buffers=4)
eng.attach_obs_gather(pg)
sync_ok = True
for it in range(20):
    act = torch.rand(N, 4, device=dev, generator=g) * 2 - 1
    eng.position_task_step(act)
    got = pg.wait().clone()
    want = nccl(eng.obs).clone()
    sync_ok &= bool(torch.equal(got, want))
act = torch.rand(N, 4, device=dev, generator=g) * 2 - 1
# free running: the pushes overlap the following steps; only the end is awaited
for it in range(50):
    eng.position_task_step(act)
got = pg.wait().clone()
want = nccl(eng.obs).clone()
free_ok = bool(torch.equal(got, want))
try:
    pg.check()
    eng.check()
except Exception as exc:  # noqa: BLE001
    free_ok = False
    sys.stderr.write(f"[rank {rank}] {exc}\n")


def step_sync(_):
    eng.position_task_step(act)
    pg.wait()


def step_free(_):
    eng.position_task_step(act)


t_sync = timeit(step_sync)
t_free = timeit(step_free, tail=lambda: (pg.fence(), pg.wait()))
eng.attach_obs_gather(None)
torch.cuda.synchronize()
dist.barrier()


def step_then_nccl(_):
    eng.position_task_step(act)
    nccl(eng.obs)


t_alone = timeit(step_free)
t_nccl = timeit(step_then_nccl)
ok &= sync_ok and free_ok
flag = torch.tensor([int(ok), int(sync_ok), int(free_ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    wire_us = (world - 1) * N * 52 / 900e9 * 1e6
    print(f"P2P_ALLGATHER world={world} N={N} equal_to_nccl={bool(flag[0].item())} p2p_us={tp:.1f} nccl_us={tn:.1f} "
          f"bytes_in_per_rank={(world - 1) * N * 52} nvlink_floor_us={wire_us:.1f} pipelined_sync_equal={bool(flag[1].item())} "
          f"pipelined_free_equal={bool(flag[2].item())} step_alone_us={t_alone:.1f} step_gather_sync_us={t_sync:.1f} "
          f"step_gather_pipelined_us={t_free:.1f} step_then_nccl_us={t_nccl:.1f}")
dist.destroy_process_group()
