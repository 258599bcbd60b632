"""torchrun self-check of the NVLink P2P observation all-gather against NCCL (needs >= 2 GPUs):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_p2p_allgather.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aerial_gym_simulator_b200.distributed import ObsAllGather, P2PObsAllGather  # noqa: E402

rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N = int(os.environ.get("N_ENVS", "65536"))
p2p = P2PObsAllGather(N, 13, dev)
p2p4 = P2PObsAllGather(N, 13, dev, num_buffers=4)
nccl = ObsAllGather(N, 13, world * N, dev)
g = torch.Generator(device=dev).manual_seed(100 + rank)
ok = True
for it in range(40):
    x = torch.rand(N, 13, device=dev, generator=g)
    a = p2p(x).clone()
    b = nccl(x).clone()
    ok &= bool(torch.equal(a, b))
torch.cuda.synchronize()
x = torch.rand(N, 13, device=dev, generator=g)


def timeit(fn, iters=200):
    for _ in range(10):
        fn(x)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn(x)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


tp, tn = timeit(p2p), timeit(nccl)

# ---- the all-gather fused into the HP1 step kernel (Hp1Engine.attach_obs_gather) -------------------
from aerial_gym_simulator_b200.hp1 import Hp1Engine, MultirotorSpec  # noqa: E402

spec = MultirotorSpec()
eng = Hp1Engine(spec, N, dev, seed=3, env_id_offset=rank * N, materialize_derived=False)
eng.reset(torch.ones(N, dtype=torch.bool, device=dev))
eng.refresh()
eng.sim_steps.copy_((torch.arange(N, device=dev) % 500).int())
eng.attach_obs_gather(p2p)
fused_ok = True
for it in range(40):
    act = torch.rand(N, 4, device=dev, generator=g) * 2 - 1
    eng.position_task_step(act)
    got = eng.gathered_obs.clone()
    want = nccl(eng.obs).clone()
    fused_ok &= bool(torch.equal(got, want))
act = torch.rand(N, 4, device=dev, generator=g) * 2 - 1


def step_fused(_):
    eng.position_task_step(act)


tf = timeit(step_fused)

# lag = 1: the kernel of step t retires when step t-1's rows are complete; gathered_obs is one step old
eng.attach_obs_gather(p2p4, lag=1)
prev_want, lag_ok = None, True
for it in range(40):
    act = torch.rand(N, 4, device=dev, generator=g) * 2 - 1
    eng.position_task_step(act)
    if prev_want is not None:
        lag_ok &= bool(torch.equal(eng.gathered_obs.clone(), prev_want))
    prev_want = nccl(eng.obs).clone()
tl = timeit(step_fused)
fused_ok &= lag_ok
eng.attach_obs_gather(None)


def step_then_p2p(_):
    eng.position_task_step(act)
    p2p(eng.obs)


ts = timeit(step_then_p2p)
ok &= fused_ok
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"P2P_ALLGATHER world={world} N={N} equal_to_nccl={bool(flag.item())} p2p_us={tp:.1f} nccl_us={tn:.1f} "
          f"bytes_in_per_rank={(world - 1) * N * 52} p2p_GBps_in={(world - 1) * N * 52 / tp / 1e3:.1f} "
          f"fused_equal={fused_ok} step_fused_us={tf:.1f} step_fused_lag1_us={tl:.1f} step_then_p2p_us={ts:.1f}")
dist.destroy_process_group()
