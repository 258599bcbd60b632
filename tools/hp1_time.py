"""Step period of the fused position-task step for the library named by AGX_LIB_PATH (A/B runs of build variants on one box):

    AGX_LIB_PATH=tools/dbg/libagx_<name>.so python tools/hp1_time.py [--envs 65536] [--steps 200]

cold16 = the timed loop of bench.py (16 rotating replicas, state cold in L2, launches chained per tile); hot = one replica back to back;
lean = materialize_derived=False (bench.py's `value`), api = materialize_derived=True (what EnvManager / the task API runs)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aerial_gym_simulator_b200.hp1 import Hp1Engine, MultirotorSpec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    N, K = a.envs, a.steps
    out = {"lib": os.path.basename(os.environ.get("AGX_LIB_PATH", "default")), "envs": N}
    g = torch.Generator(device=dev).manual_seed(1)
    acts = [(torch.rand(N, 4, generator=g, device=dev) * 2 - 1).contiguous() for _ in range(8)]
    stream = torch.cuda.current_stream(dev)
    for mode, derived in (("lean", False), ("api", True)):
        engines = []
        for rep in range(16):
            e = Hp1Engine(MultirotorSpec(), N, dev, seed=1 + rep, materialize_derived=derived)
            e.reset(torch.ones(N, dtype=torch.bool, device=dev))
            e.refresh()
            e.sim_steps.copy_((torch.arange(N, device=dev) % 500).int())
            engines.append(e)
        out["chained"] = bool(engines[0].lib.agx_hp1_task_step_is_chained(engines[0]._cfg_ref, engines[0]._buf_ref))
        for name, rot in (("cold16", True), ("hot", False)):
            best = 1e9
            for _ in range(a.reps):
                for i in range(32):
                    engines[i % 16 if rot else 0].position_task_step(acts[i % 8])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for i in range(K):
                    engines[i % 16 if rot else 0].position_task_step(acts[i % 8])
                e1.record(stream)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e3 / K)
            out[f"{name}_{mode}_us"] = round(best, 3)
        engines[0].check()
        # a checksum of the state after the same number of steps: variants must agree bit for bit
        out[f"sum_{mode}"] = float(engines[0].root_state.double().sum().item())
        del engines
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
