"""Known-answer tests for the oracle's Philox4x32-10 (Random123 kat_vectors)."""
import numpy as np

from oracle import philox


def test_philox_known_answers():
    # Random123 examples/kat_vectors: philox4x32 10 rounds
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(x) for x in got) == want


def test_uniform_range_and_layout():
    d = philox.reset_uniforms(1234, np.arange(1000), np.zeros(1000, int), 4)
    for k, v in d.items():
        assert v.dtype == np.float32 and (v >= 0).all() and (v < 1).all(), k
    assert d["state"].shape == (1000, 13) and d["tau_inc"].shape == (1000, 4)
    assert abs(float(d["state"].mean()) - 0.5) < 0.02
    # sharding invariance: draws depend on the GLOBAL env id only
    a = philox.reset_uniforms(7, np.arange(64, 128), np.full(64, 3), 4)
    b = philox.reset_uniforms(7, np.arange(0, 128), np.full(128, 3), 4)
    for k in a:
        assert np.array_equal(a[k], b[k][64:])
