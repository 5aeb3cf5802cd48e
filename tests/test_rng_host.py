"""CPU: the numpy-compatible random streams of csrc/apg_rng.h (host instantiation of the same
__host__ __device__ code the LIDAR reset path runs on the GPU) against numpy's own Generator."""
import ctypes as C

import numpy as np
import pytest


def _stream(seed, ops):
    from ap_gym_b200 import _capi

    lib = _capi.lib()
    n = len(ops)
    op = np.array([o for o, _ in ops], dtype=np.int32)
    arg = np.array([a for _, a in ops], dtype=np.uint64)
    out = np.zeros(n, dtype=np.uint64)
    _capi.check(lib.apg_rng_host_stream(C.c_uint64(seed), n, op.ctypes.data_as(C.c_void_p), arg.ctypes.data_as(C.c_void_p),
                                        out.ctypes.data_as(C.c_void_p)))
    return out


@pytest.mark.parametrize("seed", [0, 1, 7, 12345, 2**32 - 1, 2**32, 2**40 + 17, 2**63 + 5])
def test_seedsequence_pcg64_raw_stream(seed):
    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
    want = g.bit_generator.random_raw(64)
    got = _stream(seed, [(0, 0)] * 64)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("seed", [0, 3, 99, 2**31 + 11])
def test_lidar_reset_chain_matches_numpy(seed):
    """lidar_localization2d.py:547-557 + :303-311 + dataset_iterator.py:30 over 200 episodes."""
    rng = np.random.default_rng(seed + 1000)
    n_free = rng.integers(1, 900, 200)
    n_free[::17] = 1  # integers(0, 1) consumes nothing
    M = int(rng.integers(1, 5000))
    g0 = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
    s2 = int(g0.integers(0, 2**32, endpoint=True))
    g2 = np.random.default_rng(s2)
    want_map = [int(g2.integers(0, M)) for _ in n_free]
    want_cell = [int(g0.integers(0, int(f))) for f in n_free]
    got0 = _stream(seed, [(3, 0)] + [(2, int(f)) for f in n_free])
    assert int(got0[0]) == s2
    assert got0[1:].tolist() == want_cell
    got2 = _stream(s2, [(2, M)] * len(n_free))
    assert got2.tolist() == want_map


def test_bounded_draws_special_ranges():
    seed = 5
    g = np.random.default_rng(seed)
    ops, want = [], []
    for hi in [2**32, 2**32 - 1, 2, 3, 1, 2**31 + 1, 10, 2**32, 60000]:
        ops.append((2, hi))
        want.append(int(g.integers(0, hi)))
    assert _stream(seed, ops).tolist() == want
    # rejection path: a range just above 2^31 rejects ~half of the raw words
    g = np.random.default_rng(11)
    hi = 2**31 + 12345
    want = g.integers(0, hi, 500).tolist()
    assert _stream(11, [(2, hi)] * 500).tolist() == want


def test_uniform_and_advance_match_numpy():
    """image_perception_module.py:159-161 / image_localization.py:153-156: uniform(-1, 1) doubles, interleaved with
    buffered 32-bit draws, and the LCG jump used to skip other shards' draws."""
    seed = 21
    g = np.random.default_rng(seed)
    ops, want = [], []
    for rep in range(3):
        for _ in range(5):
            ops.append((2, 2))
            want.append(int(g.integers(0, 2)))
        for _ in range(7):
            ops.append((5, 0))
            want.append(np.float64(g.uniform(-1, 1)).view(np.uint64).item())
    got = _stream(seed, ops)
    assert got.tolist() == want
    # advance(k) == k discarded raw draws
    for k in (1, 2, 17, 1000, 2**20 + 3):
        g = np.random.default_rng(seed)
        g.bit_generator.advance(k)
        want = g.uniform(-1, 1, 3).view(np.uint64).tolist()
        assert _stream(seed, [(6, k), (5, 0), (5, 0), (5, 0)])[1:].tolist() == want
