"""Procedural maze maps: the iterative CPU restatement (oracle/floor_map_oracle.py) against the reference-generated
golden banks (CPU), and the device generator (csrc/apg_maps.cu, through the C-ABI) against both (-m gpu)."""
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
