"""torchrun self-check of the env-sharded task API (needs >= 2 GPUs):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_task.py
Every rank builds position_setpoint_task with args={"shard": "torchrun"}, steps it with its own actions and checks that
task_obs["observations"] equals NCCL's all_gather of the ranks' local observations, step after step."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aerial_gym_simulator_b200.task  # noqa: E402,F401
from aerial_gym_simulator_b200.registry.task_registry import task_registry  # noqa: E402

rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N = int(os.environ.get("N_ENVS", "8192"))
cfg = task_registry.get_task_config("position_setpoint_task")
cfg.device = str(dev)
task = task_registry.make_task("position_setpoint_task", seed=11, num_envs=N, headless=True, args={"shard": "torchrun"})
obs = task.reset()[0]
g = torch.Generator(device=dev).manual_seed(50 + rank)
ok, inplace = True, True
want = torch.empty(world * N, 13, device=dev)
for step in range(20):
    a = torch.rand(N, 4, generator=g, device=dev) * 2 - 1
    obs = task.step(a)[0]
    dist.all_gather_into_tensor(want, obs["observations_local"].contiguous())
    ok &= bool(torch.equal(obs["observations"], want))
    inplace &= bool(torch.equal(obs["observations"][rank * N:(rank + 1) * N], obs["observations_local"]))
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 100
for step in range(K):
    task.step(a)
torch.cuda.synchronize()
us = (time.perf_counter() - t0) / K * 1e6
task.obs_gather.check()
flag = torch.tensor([int(ok), int(inplace)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
t = torch.tensor([us], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"SHARDED_TASK world={world} N={N} global_obs_equals_nccl={bool(flag[0].item())} local_rows_in_place={bool(flag[1].item())} "
          f"task_step_with_gather_us={float(t):.1f} env_steps_per_s={world * N / float(t) * 1e6:.3e}")
task.close()
dist.destroy_process_group()
