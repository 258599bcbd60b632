"""Where does the host-I/O step spend its time?"""
import sys, time
import torch
sys.path.insert(0, ".")
from aerial_gym_simulator_b200.hp1 import Hp1Engine, MultirotorSpec

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
spec = MultirotorSpec()
K = 200


def mk(host_io):
    e = Hp1Engine(spec, N, dev, seed=1, materialize_derived=False, host_io=host_io)
    e.reset(torch.ones(N, dtype=torch.bool, device=dev))
    e.refresh()
    e.sim_steps.copy_((torch.arange(N, device=dev) % 500).int())
    return e


def timeit(fn):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e6


acts_d = [torch.rand(N, 4, device=dev) * 2 - 1 for _ in range(8)]
acts_h = [a.cpu().pin_memory() for a in acts_d]
st = torch.cuda.current_stream()
eh, ed = mk(True), mk(False)
d_act = torch.empty(N, 4, device=dev)
h_out = torch.empty(N * 13 + N + N // 2).pin_memory()


def a(i):
    eh.position_task_step(acts_d[i % 8]); st.synchronize()
def b(i):
    eh.position_task_step(acts_h[i % 8]); st.synchronize()
def c(i):
    d_act.copy_(acts_h[i % 8], non_blocking=True); eh.position_task_step(d_act); st.synchronize()
def d(i):
    ed.position_task_step(acts_d[i % 8]); st.synchronize()
def e(i):
    ed.position_task_step(acts_h[i % 8]) if False else None
def f(i):  # pure copies: 1 MB H2D + 3.8 MB D2H in one piece
    d_act.copy_(acts_h[i % 8], non_blocking=True); h_out.copy_(big, non_blocking=True); st.synchronize()
big = torch.empty(N * 13 + N + N // 2, device=dev)
def g(i):
    h_out.copy_(big, non_blocking=True); st.synchronize()
def h(i):
    d_act.copy_(acts_h[i % 8], non_blocking=True); st.synchronize()

for name, fn in (("dev actions -> host outputs", a), ("host actions -> host outputs", b), ("memcpy actions -> host outputs", c),
                 ("all device + sync", d), ("copies only (H2D 1MB + D2H 3.8MB)", f), ("D2H 3.8 MB only", g), ("H2D 1 MB only", h)):
    print(f"{name:40s} {timeit(fn):8.1f} us/step")
