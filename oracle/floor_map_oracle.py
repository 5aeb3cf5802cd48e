"""CPU ORACLE (test infrastructure, NOT product code): iterative restatements of the reference's map generators --
`rooms` further down follows FloorMapDatasetRooms.get_data_point (floor_map_dataset_rooms.py:25-89); the maze generator
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


def _distribute_integers(n: int, k: int, rng: np.random.Generator) -> np.ndarray:
    """floor_map_dataset_rooms.py:40-44"""
    r = np.arange(1, n)
    r = np.concatenate([np.zeros(max(0, k - n), dtype=np.int_), r])
    cuts = np.sort(rng.choice(r, k - 1, replace=False))
    return np.diff(np.concatenate(([0], cuts, [n])))


def rooms(idx: int, width: int = 32, height: int = 32, max_rooms: int = 10, door_width: int = 3) -> np.ndarray:
    """Iterative restatement of `FloorMapDatasetRooms.get_data_point` (floor_map_dataset_rooms.py:25-89): the recursive
    `split_room` (:46) is an explicit stack of (origin, shape, transposed, capacity) room views, popped depth-first in the
    reference's order, with the same Generator calls in the same order (binomial :54, choice :43 twice, integers :70, and
    the final transpose draw :86).  Cells are marked 1 (wall) / -1 (door, never overwritten) like the reference's int8 map.
    Pinned: bit-identical to the reference-generated bank tests/golden/maps_rooms32.npz (tests/test_maps.py)."""
    rng = np.random.default_rng(int(idx))
    mrs = door_width + 2
    m = np.zeros((height, width), dtype=np.int8)
    m[0, :] = 1
    m[-1, :] = 1
    m[:, 0] = 1
    m[:, -1] = 1

    def cell(view, a, b):  # view = (y0, x0, A, B, transposed)
        y0, x0, _, _, tr = view
        return (y0 + b, x0 + a) if tr else (y0 + a, x0 + b)

    stack = [((1, 1, height - 2, width - 2, False), max_rooms)]
    while stack:
        view, cap = stack.pop()
        y0, x0, A, B, tr = view
        mrl = min(cap, (A - mrs) // (mrs + 1) + 1)
        if mrl <= 1:
            continue
        sub = int(rng.binomial(mrl - 2, 0.3)) + 2
        caps = _distribute_integers(mrl, sub, rng)
        sizes = _distribute_integers(A - sub * (1 + mrs) + 1, sub, rng) + mrs
        ends = np.cumsum(sizes + 1) - 1
        starts = np.concatenate(([0], ends[:-1] + 2))
        walls = starts[1:] - 1
        doors = rng.integers(0, B - door_width, size=sub - 1)
        for w in walls:
            for b in range(B):
                y, x = cell(view, int(w), b)
                if m[y, x] != -1:
                    m[y, x] = 1
        for w, d in zip(walls, doors):
            for da in range(-(door_width - 1), door_width):
                for db in range(door_width):
                    y, x = cell(view, (int(w) + da) % A, int(d) + db)
                    m[y, x] = -1
        children = []
        for s, e, c in zip(starts, ends, caps):
            e = min(int(e), A - 1)  # numpy clips the slice of the last room (the sizes add up to A + 1 rows)
            s = int(s)
            child = (y0, x0 + s, B, e - s + 1, False) if tr else (y0 + s, x0, B, e - s + 1, True)
            children.append((child, int(c)))
        stack.extend(reversed(children))
    m[m == -1] = 0
    if rng.integers(0, 2) == 0:
        m = m.T
    return m.astype(np.bool_)
