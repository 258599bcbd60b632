"""Dynamic obstacles on CPU: the device code of agx_obstacle_step, compiled for the host (tests/csrc/host_shadow.cu),
against the oracle's spec (parity unpinned: PhysX moves the obstacles in the reference) and its analytic properties."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import obstacle_oracle as OB

from . import _shadow


def _run(state, twist, dt, substeps, ld=0.1, ad=0.1):
    lib = _shadow.load()
    s = np.ascontiguousarray(state.numpy().astype(np.float32)).copy()
    tw = None if twist is None else np.ascontiguousarray(twist.numpy().astype(np.float32))
    n, a = s.shape[:2]
    lib.shadow_obstacle_step(n, a, s.ctypes.data_as(C.c_void_p), s.shape[2], tw.ctypes.data_as(C.c_void_p) if tw is not None else None,
                             dt, substeps, ld, ad)
    return torch.from_numpy(s)


def _random_state(n, a, g, stride=13):
    s = torch.zeros(n, a, stride)
    s[..., 0:3] = torch.randn(n, a, 3, generator=g) * 4
    q = torch.randn(n, a, 4, generator=g)
    s[..., 3:7] = q / q.norm(dim=-1, keepdim=True)
    s[..., 7:13] = torch.randn(n, a, 6, generator=g)
    return s


@pytest.mark.parametrize("with_twist,substeps", [(True, 1), (True, 10), (False, 3), (True, 0)])
def test_shadow_matches_oracle(with_twist, substeps):
    g = torch.Generator().manual_seed(3)
    s = _random_state(7, 35, g)
    s[0, 0, 10:13] = 0.0  # a non-rotating obstacle (the wn = 0 branch)
    tw = torch.randn(7, 35, 6, generator=g) * 1.5 if with_twist else None
    if with_twist:
        tw[0, 0, 3:6] = 0.0
    got = _run(s, tw, 0.01, substeps)
    want = OB.obstacle_step(s, tw, 0.01, substeps)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), (got - want).abs().max()
    assert torch.allclose(got[..., 3:7].norm(dim=-1), torch.ones(7, 35), atol=1e-6)
    if substeps == 0:
        assert torch.equal(got, s)


def test_padded_rows_and_analytic_motion():
    """stride > 13 (extra columns untouched); constant twist: straight line at the damped speed, rotation about a fixed axis."""
    g = torch.Generator().manual_seed(4)
    s = _random_state(2, 3, g, stride=16)
    s[..., 13:] = 42.0
    s[..., 3:7] = torch.tensor([0.0, 0.0, 0.0, 1.0])
    tw = torch.zeros(2, 3, 6)
    tw[..., 0], tw[..., 5] = 1.0, 0.5  # 1 m/s along x, 0.5 rad/s about z
    dt, n = 0.01, 100
    got = _run(s, tw, dt, n, ld=0.1, ad=0.1)
    k = 1.0 - dt * 0.1
    assert torch.allclose(got[..., 0], s[..., 0] + n * dt * k, atol=1e-4) and torch.equal(got[..., 1:3], s[..., 1:3])
    yaw = n * dt * 0.5 * k
    assert torch.allclose(got[..., 5], torch.full((2, 3), math.sin(yaw / 2)), atol=1e-5)
    assert torch.allclose(got[..., 6], torch.full((2, 3), math.cos(yaw / 2)), atol=1e-5)
    assert (got[..., 13:] == 42.0).all()
    assert torch.allclose(got[..., 7], torch.full((2, 3), k)) and torch.allclose(got[..., 12], torch.full((2, 3), 0.5 * k))
