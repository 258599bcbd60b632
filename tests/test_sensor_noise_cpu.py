"""Ray-cast sensor noise / range limits / normalisation on CPU:
  (a) the host's torch path (sensors/noise.py: apply_noise_and_limits_torch) against fixtures recorded from the reference's OWN
      apply_noise / apply_range_limits / normalize_observation under the same torch seed -- bit for bit;
  (b) the device code of agx_hp2_noise_limits compiled for the host (tests/csrc/host_shadow.cu): its Philox against the Random123
      known answers, its output against the numpy oracle (oracle/sensor_noise_oracle.py), and its distribution against (a)."""
import ctypes as C
import os
import types

import numpy as np
import pytest
import torch

from aerial_gym_simulator_b200 import _lib
from aerial_gym_simulator_b200.sensors.noise import apply_noise_and_limits_torch, noise_struct
from oracle import sensor_noise_oracle as SN

from . import _shadow

G = os.path.join(os.path.dirname(__file__), "golden", "sensor_noise.npz")
CASES = ["lidar_range", "lidar_pc_sensor", "lidar_pc_world", "camera_depth_nonoise"]


def _cfg(spec):
    st, pc, world, norm, noise, drop = spec[0], spec[1] == "1", spec[2] == "1", spec[3] == "1", spec[4] == "1", float(spec[5])
    nz = types.SimpleNamespace(enable_sensor_noise=noise, std_a=0.00038089, std_b=-0.00343351, std_c=0.01553284, mean_offset=-0.025,
                               pixel_dropout_prob=drop)
    return types.SimpleNamespace(sensor_type=st, return_pointcloud=pc, pointcloud_in_world_frame=world, normalize_range=norm, max_range=10.0,
                                 min_range=0.2, far_out_of_range_value=10.0, near_out_of_range_value=-10.0, sensor_noise=nz)


@pytest.mark.parametrize("name", CASES)
def test_torch_path_reproduces_the_reference(name):
    d = np.load(G)
    px = torch.tensor(d[f"{name}_in"])
    torch.manual_seed(int(d["seed"]))
    out = apply_noise_and_limits_torch(px, _cfg(d[f"{name}_spec"]))
    assert torch.equal(out, torch.tensor(d[f"{name}_out"]))


def test_device_philox_known_answers():
    """the Philox4x32-10 text the kernels compile (agx_math.cuh), run on the host: Random123 kat_vectors"""
    lib = _shadow.load()
    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, want in kat:
        c, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 4)()
        lib.shadow_philox4x32_10(C.cast(c, C.c_void_p), key[0], key[1], C.cast(o, C.c_void_p))
        assert tuple(o) == want


def _shadow_noise(px, n, seed, frame, first_pixel=0):
    lib = _shadow.load()
    a = np.ascontiguousarray(px, np.float32).copy()
    lib.shadow_noise_limits(a.ctypes.data_as(C.c_void_p), a.size // n.components, first_pixel, C.cast(C.byref(n), C.c_void_p), seed, frame)
    return a


def _oracle_noise(px, n, seed, frame, first_pixel=0):
    return SN.noise_limits(px, n.components, n.enable_noise, n.apply_limits, n.normalize, n.std_a, n.std_b, n.std_c, n.mean_offset,
                           n.pixel_dropout_prob, n.max_range, n.min_range, n.far_out_of_range_value, n.near_out_of_range_value, seed, frame,
                           first_pixel)


@pytest.mark.parametrize("name", CASES)
def test_shadow_matches_oracle(name):
    d = np.load(G)
    n = noise_struct(_cfg(d[f"{name}_spec"]))
    assert (n.components, n.apply_limits, n.normalize) == {"lidar_range": (1, 1, 1), "lidar_pc_sensor": (3, 1, 1), "lidar_pc_world": (3, 0, 0),
                                                           "camera_depth_nonoise": (1, 1, 1)}[name]
    px = d[f"{name}_in"]
    seed, frame = 0x1234_5678_9ABC_DEF0, 7
    got, want = _shadow_noise(px, n, seed, frame), _oracle_noise(px, n, seed, frame)
    # same uniforms on both sides; libm vs numpy log/cos differ by ulps, and a 1000 m "no hit" pixel has std = 377 m
    assert np.allclose(got, want, rtol=1e-4, atol=2e-4 * max(1.0, float(np.abs(want).max()))), np.abs(got - want).max()
    if not n.enable_noise:
        assert np.array_equal(got, d[f"{name}_out"])  # no randomness: identical to the reference's limits + normalisation
    # determinism, and a different stream for a different frame / seed
    assert np.array_equal(got, _shadow_noise(px, n, seed, frame))
    if n.enable_noise:
        assert not np.array_equal(got, _shadow_noise(px, n, seed, frame + 1)) and not np.array_equal(got, _shadow_noise(px, n, seed + 1, frame))
    # sharding invariance: the second half of the image processed on its own, with its global pixel offset (beyond 2^32 too)
    for base in (0, (1 << 33) + 5):
        whole = _shadow_noise(px, n, seed, frame, base)
        half = px.reshape(-1, n.components).shape[0] // 2
        tail = _shadow_noise(px.reshape(-1, n.components)[half:], n, seed, frame, base + half)
        assert np.array_equal(whole.reshape(-1, n.components)[half:], tail)
        assert np.allclose(whole, _oracle_noise(px, n, seed, frame, base), rtol=1e-4, atol=2e-4 * max(1.0, float(np.abs(want).max())))


def test_device_rng_noise_has_the_reference_distribution():
    """a constant 4 m range image: mean = p - mean_offset, std = a p^2 + b p + c, dropout fraction = p_drop, against both the
    closed form and the torch (reference-order) path on the same input"""
    cfg = _cfg(np.array(["lidar", "0", "0", "0", "1", "0.05"]))
    cfg.near_out_of_range_value = -1.0
    n = noise_struct(cfg)
    P = 400_000
    px = np.full((P,), 4.0, np.float32)
    got = _shadow_noise(px, n, 99, 0)
    dropped = got == -1.0
    assert abs(dropped.mean() - 0.05) < 0.002
    std = 0.00038089 * 16 + -0.00343351 * 4 + 0.01553284
    kept = got[~dropped]
    assert abs(kept.mean() - 4.025) < 4 * std / np.sqrt(kept.size) and abs(kept.std() - std) < 0.01 * std
    torch.manual_seed(0)
    ref = apply_noise_and_limits_torch(torch.full((1, 1, 1, P), 4.0), cfg).numpy().ravel()
    rk = ref[ref != -1.0]
    assert abs(rk.mean() - kept.mean()) < 6 * std / np.sqrt(kept.size) and abs(rk.std() - kept.std()) < 0.01 * std
    # tails: Box-Muller with a 24-bit u1 reaches |z| ~ 5.7
    z = (kept - 4.025) / std
    assert 4.0 < np.abs(z).max() < 6.0 and abs((np.abs(z) > 2).mean() - 0.0455) < 0.003


def test_struct_matches_library():
    assert _lib.load().agx_sizeof(8) == C.sizeof(_lib.AgxHp2Noise)
