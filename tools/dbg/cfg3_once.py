"""A few frames of BASELINE configs[2] (270x480 depth+seg camera, 1024 boxes per env) at a reduced env count -- the target of the ncu capture
of the records-only tile path (tools/profile_gpu_r2.sh)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
dev = torch.device("cuda:0")
cam3 = type("Cam270x480", (bench._CamCfg,), {"height": 270, "width": 480})
r = bench.run_hp2_config(dev, 1, 0, label="cfg3 (ncu)", metric="rays/s", E=int(os.environ.get("E", "256")), K=1024, cfg=cam3, frames=2, extent=8.0, seed=11)
print({k: r[k] for k in ("value", "ms_per_frame", "traversal")})
