"""CPU ORACLE (test infrastructure, NOT product code): iterative restatement of the reference's maze generator
`FloorMapDatasetMaze.get_data_point` (/root/reference/ap_gym/envs/floor_map/floor_map_dataset_maze.py:24-55).

The reference carves recursively (`carve` :31); here the recursion is an explicit stack of (cell, direction order,
next direction, first flag) frames -- the same random draws in the same order:
  rng = default_rng(idx); per visited cell ONE `rng.permutation(directions)` (numpy: shuffle of arange(4) with
  `random_interval`, i.e. masked 32-bit draws with rejection, i = 3, 2, 1), and for every eligible neighbour after the
  first carved one ONE `rng.random()` compared with branching_prob.
Pinned: bit-identical to the reference-generated banks tests/golden/maps_maze{21,63}.npz (tests/test_maps.py)."""
from __future__ import annotations

import numpy as np

DIRECTIONS = np.array([[2, 0], [-2, 0], [0, 2], [0, -2]])


def maze(idx: int, width: int = 21, height: int = 21, branching_prob: float = 1.0) -> np.ndarray:
    rng = np.random.default_rng(int(idx))
    m = np.ones((height, width), dtype=np.bool_)
    m[1, 1] = False
    # frame: [x, y, order (4 direction indices), k (next position in order), first]
    stack = [[1, 1, rng.permutation(4), 0, True]]
    while stack:
        fr = stack[-1]
        if fr[3] == 4:
            stack.pop()
            continue
        d = DIRECTIONS[fr[2][fr[3]]]
        fr[3] += 1
        nx, ny = fr[0] + d[0], fr[1] + d[1]
        if 0 < nx < width - 1 and 0 < ny < height - 1 and m[ny, nx]:
            if fr[4] or rng.random() < branching_prob:
                m[fr[1] + d[1] // 2, fr[0] + d[0] // 2] = False
                m[ny, nx] = False
                fr[4] = False  # (the reference clears `first` after the recursive call returns: same effect)
                stack.append([nx, ny, rng.permutation(4), 0, True])
    return m
