"""Build + load the CPU shadow of the per-env device code (tests/csrc/host_shadow.cu).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "host_shadow.cu")
LIB = os.path.join(HERE, "_build", "libagx_host_shadow.so")
_CSRC = os.path.join(HERE, "..", "aerial_gym_simulator_b200", "csrc")
_DEPS = [SRC, os.path.join(_CSRC, "lidar_nav_core.cuh"), os.path.join(_CSRC, "obstacle_core.cuh"), os.path.join(_CSRC, "noise_core.cuh"), os.path.join(_CSRC, "agx_math.cuh"),
         os.path.join(HERE, "..", "include", "aerial_gym_b200.h")]
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in _DEPS):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # host code only is used; nvcc is the compiler because the headers are CUDA headers (float4, __host__ __device__)
    cmd = [nvcc, "-x", "cu", "-DAGX_HOST_SHADOW", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
           "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", LIB, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("host shadow build failed:\n" + r.stdout + r.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.shadow_lidar_nav_pool.restype = C.c_int
        lib.shadow_lidar_nav_pool.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_void_p, C.c_int]
        lib.shadow_lidar_nav_reward.restype = None
        lib.shadow_lidar_nav_reward.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 10 + [C.c_float, C.c_void_p] + [C.c_void_p] * 3
        lib.shadow_lidar_nav_obs.restype = None
        lib.shadow_lidar_nav_obs.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 10 + [C.c_int, C.c_void_p, C.c_int]
        lib.shadow_obstacle_step.restype = None
        lib.shadow_obstacle_step.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_float, C.c_float]
        lib.shadow_noise_limits.restype = None
        lib.shadow_noise_limits.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32]
        lib.shadow_philox4x32_10.restype = None
        lib.shadow_philox4x32_10.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib = lib
    return _lib
