"""2-GPU micro-benchmark of the gather's push / wait kernels alone (no step kernels): latency of one push + wait in stream order,
and throughput of pushes rotating over the ring's side streams.  torchrun --nproc-per-node 2 tools/dbg/push_sweep.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from aerial_gym_simulator_b200.distributed import PipelinedObsGather
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N = 65536
for ctas in (8, 24, 64, 148, 296):
    g = PipelinedObsGather(N, 13, dev, num_buffers=4, max_ctas=ctas)
    cur = torch.cuda.current_stream(dev)
    def sync_loop(iters):
        for _ in range(iters):
            e, s = g.next_epoch()
            g.push(g.own_slot_ptr[s], e, s, stream=cur.cuda_stream)
            g.wait()
    def pipe_loop(iters):
        for _ in range(iters):
            e, s = g.next_epoch()
            g.push(g.own_slot_ptr[s], e, s)
        g.fence(); g.wait()
    res = {}
    for name, fn in (("sync", sync_loop), ("pipe", pipe_loop)):
        fn(10); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(200); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = round(float(t), 2)
    g.check()
    if rank == 0:
        print(f"PUSH_SWEEP world={world} ctas={ctas} us_per_push sync={res['sync']} pipelined={res['pipe']} GBps_pipe={(world-1)*N*52/res['pipe']/1e3:.0f}", flush=True)
    del g
dist.destroy_process_group()
