"""Keeps the TEXT of the GPU tests that have CPU twins runnable: a subprocess `pytest -m gpu --gpu-dryrun` (tests/conftest.py) over the
test_zz_* files and the env- / task-level GPU tests.  A change of the host code that breaks a GPU test shows up here, not on the one
GPU run at the end of a round.  (Tests of GPU-only mechanics -- tile counters, host-mapped I/O, peer gather, the ray-caster's own
kernels -- have no twin and are not part of this.)"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_test_text_runs_on_the_cpu_twins():
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "test_zz_*_gpu.py"))) + [os.path.join(ROOT, "tests", f) for f in
                                                                                 ("test_aux_gpu.py", "test_env_task_gpu.py")]
    no_twin = ["--deselect", "tests/test_env_task_gpu.py::test_task_torch_rng_mode_follows_reference_call_order"]  # reads the CUDA RNG state
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "--gpu-dryrun", "-p", "no:cacheprovider", *no_twin, *files],
                       cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    assert r.returncode == 0, tail
    assert " passed" in tail and " failed" not in tail, tail
