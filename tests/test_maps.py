"""Procedural maps (mazes, rooms): the iterative CPU restatements (oracle/floor_map_oracle.py) against the
reference-generated golden banks (CPU), the product's host/device rooms generator compiled as host code against the same
bank (CPU), and the device generators (csrc/apg_maps.cu, through the C-ABI) against both (-m gpu)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN


def _golden(name, count):
    z = np.load(os.path.join(GOLDEN, f"maps_{name}.npz"))
    return z["bits"][:count], int(z["width"]), int(z["height"])


def _pack(maps):
    from ap_gym_b200.floor_maps import pack_maps

    return pack_maps(np.stack(maps))


@pytest.mark.parametrize("name,count", [("maze21", 200), ("maze63", 12)])
def test_iterative_maze_oracle_matches_reference_banks(name, count):
    """The goldens are FloorMapDatasetMaze(...)[0..M-1] of the unmodified reference (tests/golden/make_golden_lidar.py)."""
    from oracle.floor_map_oracle import maze

    bits, W, H = _golden(name, count)
    assert np.array_equal(_pack([maze(i, W, H) for i in range(count)]), bits)


@pytest.mark.gpu
@pytest.mark.parametrize("name,count", [("maze21", 1024), ("maze63", 256)])
def test_device_maze_generator_matches_reference_banks(name, count):
    import ap_gym_b200 as apg

    bits, W, H = _golden(name, count)
    bank = apg.FloorMapBank.generate_maze(np.arange(count), W, H)
    assert bank.bits.shape == bits.shape and np.array_equal(bank.bits, bits)


@pytest.mark.gpu
def test_device_maze_generator_matches_oracle_other_parameters():
    """Large indices (two-word seeds), odd non-square sizes and branching_prob < 1 (exercises rng.random())."""
    import ap_gym_b200 as apg
    from oracle.floor_map_oracle import maze

    for W, H, p, idx in [(21, 21, 0.5, [0, 7, 2**31 + 5, 2**32 - 1]), (33, 15, 0.8, [3, 123456789]), (85, 85, 1.0, [1, 2]),
                         (127, 127, 1.0, [0])]:
        bank = apg.FloorMapBank.generate_maze(idx, W, H, branching_prob=p)
        want = _pack([maze(i, W, H, p) for i in idx])
        assert np.array_equal(bank.bits, want), (W, H, p)


# ------------------------------------------------------------------------------------------------ rooms
def test_iterative_rooms_oracle_matches_reference_bank():
    """The golden is FloorMapDatasetRooms()[0..4095] of the unmodified reference (tests/golden/make_golden_lidar.py)."""
    from oracle.floor_map_oracle import rooms

    bits, W, H = _golden("rooms32", 300)
    assert np.array_equal(_pack([rooms(i, W, H) for i in range(300)]), bits)


@pytest.mark.reference
def test_iterative_rooms_oracle_matches_live_reference_other_parameters():
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "shim"))
    sys.path.insert(0, "/root/reference")
    from ap_gym.envs.floor_map.floor_map_dataset_rooms import FloorMapDatasetRooms

    from oracle.floor_map_oracle import rooms

    for W, mr, dw, idx in [(48, 6, 2, [0, 1, 2, 3, 99]), (64, 10, 3, [5, 2**31 + 7]), (24, 4, 4, [11, 12]), (96, 14, 2, [3])]:
        ds = FloorMapDatasetRooms(W, W, mr, dw)
        for i in idx:
            assert np.array_equal(rooms(i, W, W, mr, dw), ds.get_data_point(i)), (W, mr, dw, i)


_ROOMS_LIB = {}


def _rooms_lib(tmp_path):
    """csrc/apg_rooms.cuh (the code the kernel runs per thread) compiled with g++."""
    import ctypes as C
    import subprocess

    if "lib" not in _ROOMS_LIB:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        lib_path = str(tmp_path / "librooms_harness.so")
        r = subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I",
                            os.path.join(root, "active-perception-gym_b200", "csrc"), "-o", lib_path,
                            os.path.join(root, "tests", "native", "rooms_harness.cpp")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        _ROOMS_LIB["lib"] = C.CDLL(lib_path)
    return _ROOMS_LIB["lib"]


def _rooms_host(idx, W, mr, dw, tmp_path):
    import ctypes as C

    lib = _rooms_lib(tmp_path)
    idx = np.ascontiguousarray(idx, dtype=np.uint64)
    out = np.zeros((len(idx), W, (W + 31) // 32), dtype=np.uint32)
    bad = lib.apg_rooms_host(idx.ctypes.data_as(C.c_void_p), C.c_longlong(len(idx)), W, W, mr, dw,
                             out.ctypes.data_as(C.c_void_p))
    assert bad == 0
    return out


def test_host_device_rooms_generator_matches_reference_bank(tmp_path):
    """All 4096 maps of the golden bank, and other parameters against the oracle."""
    from oracle.floor_map_oracle import rooms

    bits, W, H = _golden("rooms32", 4096)
    assert np.array_equal(_rooms_host(np.arange(4096), W, 10, 3, tmp_path), bits)
    for W, mr, dw, idx in [(48, 6, 2, [0, 1, 2, 3, 99]), (64, 10, 3, [5, 2**31 + 7, 2**32 - 1]), (24, 4, 4, [11, 12]),
                           (96, 14, 2, [3])]:
        assert np.array_equal(_rooms_host(idx, W, mr, dw, tmp_path), _pack([rooms(i, W, W, mr, dw) for i in idx])), (W, mr, dw)


@pytest.mark.gpu
def test_device_rooms_generator_matches_reference_bank():
    import ap_gym_b200 as apg

    bits, W, H = _golden("rooms32", 4096)
    bank = apg.FloorMapBank.generate_rooms(np.arange(4096), W, H)
    assert bank.bits.shape == bits.shape and np.array_equal(bank.bits, bits)


@pytest.mark.gpu
def test_device_rooms_generator_matches_oracle_other_parameters():
    import ap_gym_b200 as apg
    from oracle.floor_map_oracle import rooms

    for W, mr, dw, idx in [(48, 6, 2, [0, 1, 2, 3, 99]), (64, 10, 3, [5, 2**31 + 7, 2**32 - 1]), (24, 4, 4, [11, 12]),
                           (96, 14, 2, [3])]:
        bank = apg.FloorMapBank.generate_rooms(idx, W, W, max_rooms=mr, door_width=dw)
        assert np.array_equal(bank.bits, _pack([rooms(i, W, W, mr, dw) for i in idx])), (W, mr, dw)


def test_host_device_rooms_generator_property(tmp_path):
    """hypothesis: random sizes / max_rooms / door widths / 64-bit indices -- the host/device generator equals the oracle,
    and reports a failure exactly where the reference's arithmetic raises (numpy: low >= high for the door draw)."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from oracle.floor_map_oracle import rooms

    @settings(max_examples=60, deadline=None)
    @given(W=st.integers(12, 72), mr=st.integers(1, 14), dw=st.integers(1, 5), idx=st.integers(0, 2**64 - 1))
    def check(W, mr, dw, idx):
        try:
            want = rooms(idx, W, W, mr, dw)
        except (ValueError, IndexError):
            want = None
        import ctypes as C

        lib = _rooms_lib(tmp_path)
        i = np.array([idx], dtype=np.uint64)
        out = np.zeros((1, W, (W + 31) // 32), dtype=np.uint32)
        bad = lib.apg_rooms_host(i.ctypes.data_as(C.c_void_p), C.c_longlong(1), W, W, mr, dw, out.ctypes.data_as(C.c_void_p))
        if want is None:
            assert bad == 1, (W, mr, dw, idx)
        else:
            assert bad == 0 and np.array_equal(out, _pack([want])), (W, mr, dw, idx)

    check()
