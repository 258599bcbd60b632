"""Build + load the CPU shadow of the per-env device code (tests/csrc/host_shadow.cu).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "host_shadow.cu")
LIB = os.path.join(HERE, "_build", "libagx_host_shadow.so")
_CSRC = os.path.join(HERE, "..", "aerial_gym_simulator_b200", "csrc")
_HP1_CU = os.path.join(_CSRC, "hp1.cu")
_DEPS = [SRC, os.path.join(HERE, "csrc", "host_shadow_hp1.inc"), _HP1_CU, os.path.join(_CSRC, "hp1_core.cuh"), os.path.join(_CSRC, "aux_core.cuh"), os.path.join(_CSRC, "e2e_task_core.cuh"), os.path.join(_CSRC, "sim2real_core.cuh"), os.path.join(_CSRC, "disturbance_core.cuh"), os.path.join(_CSRC, "lidar_nav_core.cuh"), os.path.join(_CSRC, "obstacle_core.cuh"), os.path.join(_CSRC, "noise_core.cuh"), os.path.join(_CSRC, "agx_math.cuh"),
         os.path.join(HERE, "..", "include", "aerial_gym_b200.h")]
_lib = None


def _extract_kernel_blocks():
    """The two inline blocks of hp1_step_kernel (one physics sub-step; the position-task reward) are not functions -- making them
    functions changes the kernel's SASS -- so their TEXT is lifted out of hp1.cu, between `// AGX_SHADOW_BEGIN(name)` and
    `// AGX_SHADOW_END(name)`, into tests/_build/hp1_<name>.inc, which the shadow includes in a scope with the same variable names."""
    import re
    txt = open(_HP1_CU).read()
    for name in ("physics_substep", "position_task_reward"):
        m = re.search(r"// AGX_SHADOW_BEGIN\(%s\)[^\n]*\n(.*?)\n[ \t]*// AGX_SHADOW_END\(%s\)" % (name, name), txt, re.S)
        if not m:
            raise RuntimeError(f"marker AGX_SHADOW_BEGIN({name}) not found in hp1.cu")
        with open(os.path.join(os.path.dirname(LIB), f"hp1_{name}.inc"), "w") as f:
            f.write(m.group(1) + "\n")
    # the instantiation table (which controller x motor-model combinations have a compile-time specialised kernel): the shadow
    # dispatches over the same list, so the CPU suite can run the SPECIALISED text of hp1_core.cuh as well as the dynamic one
    m = re.search(r"(#define AGX_SPEC_FLAGS_ALL[^\n]*\n#ifndef AGX_HP1_NO_SPEC\n.*?\n#endif)", txt, re.S)
    if not m:
        raise RuntimeError("instantiation table (AGX_HP1_SPEC_LIST) not found in hp1.cu")
    with open(os.path.join(os.path.dirname(LIB), "hp1_spec_list.inc"), "w") as f:
        f.write(m.group(1) + "\n")


def build(force=False):
    # AGX_SHADOW_FLAGS: extra -D flags, to run the CPU suite on a build VARIANT of the device code (tools/build_variant.py)
    extra = os.environ.get("AGX_SHADOW_FLAGS", "").split()
    stamp = LIB + ".flags"
    same_flags = os.path.exists(stamp) and open(stamp).read().split() == extra
    if not force and same_flags and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in _DEPS):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    _extract_kernel_blocks()
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # host code only is used; nvcc is the compiler because the headers are CUDA headers (float4, __host__ __device__)
    cmd = [nvcc, "-x", "cu", "-DAGX_HOST_SHADOW", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
           "-Xcompiler", "-fPIC,-ffp-contract=off", "-I", os.path.dirname(LIB), "-shared", "-o", LIB, SRC] + extra
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("host shadow build failed:\n" + r.stdout + r.stderr)
    open(stamp, "w").write(" ".join(extra))
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
        from aerial_gym_simulator_b200 import _lib as A
        for name, extra in (("shadow_hp1_physics_step", []), ("shadow_hp1_position_task_step", [C.c_int]),
                            ("shadow_hp1_reset", [C.c_void_p, C.POINTER(A.AgxHp1ResetDraws)]), ("shadow_hp1_refresh", [])):
            fn = getattr(lib, name)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(A.AgxHp1Config), C.POINTER(A.AgxHp1Buffers)] + extra
        lib.shadow_nav_reward.restype = lib.shadow_nav_obs.restype = lib.shadow_imu_update.restype = None
        lib.shadow_nav_reward.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_float] + [C.c_void_p] * 4
        lib.shadow_nav_obs.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 9 + [C.c_int]
        lib.shadow_imu_update.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 6
        lib.shadow_e2e_reward.restype = lib.shadow_e2e_obs.restype = None
        lib.shadow_e2e_reward.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 8
        lib.shadow_e2e_obs.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int]
        lib.shadow_s2r_reward.restype = lib.shadow_s2r_obs.restype = None
        lib.shadow_s2r_reward.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 9
        lib.shadow_s2r_obs.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int]
        lib.shadow_disturbance_draw.restype = C.c_int
        lib.shadow_disturbance_draw.argtypes = [C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        lib.shadow_hp1_set_specialised.restype = None
        lib.shadow_hp1_set_specialised.argtypes = [C.c_int]
        lib.shadow_hp1_spec_id.restype = C.c_int
        lib.shadow_hp1_spec_id.argtypes = [C.POINTER(A.AgxHp1Config)]
        lib.shadow_hp1_position_reward.restype = C.c_int
        lib.shadow_hp1_position_reward.argtypes = [C.POINTER(A.AgxHp1Config), C.c_int] + [C.c_void_p] * 8
        _lib = lib
    return _lib
