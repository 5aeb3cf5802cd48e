"""Floor-map banks: the device-resident replacement of the reference's per-episode map fetch
(ap_gym/envs/floor_map/*.py, ap_gym/envs/lidar_localization2d.py:279-298).

A bank is a FINITE list of occupancy maps stored bit-packed, (M, H, ceil(W/32)) uint32 with bit x&31 of word
x>>5 set for a wall.  Any object with the reference's FloorMapDataset surface (`map_width`, `map_height`,
`__len__`, `__getitem__`) can be turned into a bank.

The reference's PROCEDURAL datasets (`FloorMapDatasetRooms`, `FloorMapDatasetMaze`, len == 2**32) are mirrored here by
classes of the same names and constructor signatures; handed to `LIDARLocalization2DVectorEnv` -- these or the
reference's own objects -- they switch the engine to procedural-dataset mode: every (auto)reset draws
`integers(0, 2**32)` from the environment's DatasetIterator stream and generates that map on the device, exactly what the
reference's `next(data_loader)` yields (csrc/apg_rooms.cuh, csrc/apg_maze.cuh, apg_lidar_set_map_generator).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np

MAX_IMPLICIT_BANK = 1 << 20


def pack_maps(maps: np.ndarray) -> np.ndarray:
    maps = np.asarray(maps, dtype=bool)
    if maps.ndim == 2:
        maps = maps[None]
    M, H, W = maps.shape
    out = np.zeros((M, H, (W + 31) // 32), dtype=np.uint32)
    for x in range(W):
        out[:, :, x >> 5] |= maps[:, :, x].astype(np.uint32) << np.uint32(x & 31)
    return out


def unpack_maps(bits: np.ndarray, width: int) -> np.ndarray:
    x = np.arange(width)
    return ((bits[..., x >> 5] >> (x & 31).astype(np.uint32)) & 1).astype(bool)


class _ProceduralDataset:
    """Common surface of the procedural datasets (ap_gym/envs/floor_map/floor_map_dataset.py, envs/dataset/dataset.py)."""

    def __len__(self):
        return 2 ** 32

    def load(self):
        pass

    def get_data_point(self, idx) -> np.ndarray:
        return self.get_data_point_batch([idx])[0]

    def __getitem__(self, item):
        if isinstance(item, (list, tuple, np.ndarray)):
            return self.get_data_point_batch(item)
        return self.get_data_point(item)


class FloorMapDatasetRooms(_ProceduralDataset):
    """Mirror of the reference's `FloorMapDatasetRooms(width, height, max_rooms, door_width)`
    (floor_map_dataset_rooms.py:9-23); maps are generated on the GPU, bit-identical to the reference's."""

    def __init__(self, width: int = 32, height: int = 32, max_rooms: int = 10, door_width: int = 3, device: int = 0):
        self.map_width, self.map_height = int(width), int(height)
        self.max_rooms, self.door_width, self.device = int(max_rooms), int(door_width), int(device)

    def get_data_point_batch(self, idx) -> np.ndarray:
        bank = FloorMapBank.generate_rooms(idx, self.map_width, self.map_height, self.max_rooms, self.door_width, self.device)
        return unpack_maps(bank.bits, self.map_width)


class FloorMapDatasetMaze(_ProceduralDataset):
    """Mirror of the reference's `FloorMapDatasetMaze(width, height, branching_prob)` (floor_map_dataset_maze.py:10-22)."""

    def __init__(self, width: int = 21, height: int = 21, branching_prob: float = 1.0, device: int = 0):
        if width % 2 == 0 or height % 2 == 0:
            raise ValueError("Width and height must be odd.")
        self.map_width, self.map_height = int(width), int(height)
        self.branching_prob, self.device = float(branching_prob), int(device)

    def get_data_point_batch(self, idx) -> np.ndarray:
        bank = FloorMapBank.generate_maze(idx, self.map_width, self.map_height, self.branching_prob, self.device)
        return unpack_maps(bank.bits, self.map_width)


def procedural_spec(dataset):
    """("rooms" | "maze", parameters) when `dataset` is a procedural map dataset -- one of the mirrors above or the
    reference's own object (recognised by class name, parameters read from its name-mangled attributes) -- else None."""
    name = type(dataset).__name__
    if isinstance(dataset, FloorMapDatasetRooms):
        return "rooms", dict(width=dataset.map_width, height=dataset.map_height, max_rooms=dataset.max_rooms,
                             door_width=dataset.door_width)
    if isinstance(dataset, FloorMapDatasetMaze):
        return "maze", dict(width=dataset.map_width, height=dataset.map_height, branching_prob=dataset.branching_prob)
    try:
        if name == "FloorMapDatasetRooms":
            g = lambda a: getattr(dataset, "_FloorMapDatasetRooms__" + a)  # noqa: E731
            return "rooms", dict(width=int(g("width")), height=int(g("height")), max_rooms=int(g("max_rooms")),
                                 door_width=int(g("door_width")))
        if name == "FloorMapDatasetMaze":
            g = lambda a: getattr(dataset, "_FloorMapDatasetMaze__" + a)  # noqa: E731
            return "maze", dict(width=int(g("width")), height=int(g("height")), branching_prob=float(g("branching_prob")))
    except AttributeError:
        return None
    return None


@dataclass
class FloorMapBank:
    bits: np.ndarray  # (M, H, words) uint32
    map_width: int
    map_height: int

    def __post_init__(self):
        self.bits = np.ascontiguousarray(self.bits, dtype=np.uint32)
        if self.bits.ndim != 3 or self.bits.shape[1] != self.map_height or \
                self.bits.shape[2] != (self.map_width + 31) // 32:
            raise ValueError("bits must have shape (M, map_height, ceil(map_width/32))")

    def __len__(self):
        return len(self.bits)

    def load(self):
        pass

    def get_data_point(self, idx) -> np.ndarray:
        return unpack_maps(self.bits[int(idx)], self.map_width)

    def __getitem__(self, item):
        if isinstance(item, (list, tuple, np.ndarray)):
            return unpack_maps(self.bits[np.asarray(item)], self.map_width)
        return self.get_data_point(item)

    @classmethod
    def generate_maze(cls, indices, width: int = 21, height: int = 21, branching_prob: float = 1.0,
                      device: int = 0) -> "FloorMapBank":
        """`FloorMapDatasetMaze(width, height, branching_prob)[indices]` generated on the GPU (csrc/apg_maps.cu):
        bit-identical to the reference's generator, without its 6-300 ms of host time per map."""
        import ctypes as C

        from . import _capi

        if width % 2 == 0 or height % 2 == 0:
            raise ValueError("Width and height must be odd.")
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64).ravel())
        bits = np.empty((len(idx), height, (width + 31) // 32), dtype=np.uint32)
        _capi.check(_capi.lib().apg_maps_generate_maze(int(device), idx.ctypes.data_as(C.c_void_p), len(idx), int(width),
                                                       int(height), float(branching_prob), bits.ctypes.data_as(C.c_void_p)))
        return cls(bits, int(width), int(height))

    @classmethod
    def generate_rooms(cls, indices, width: int = 32, height: int = 32, max_rooms: int = 10, door_width: int = 3,
                       device: int = 0) -> "FloorMapBank":
        """`FloorMapDatasetRooms(width, height, max_rooms, door_width)[indices]` generated on the GPU (csrc/apg_rooms.cuh,
        csrc/apg_maps.cu): bit-identical to the reference's generator (its numpy draws restated on the PCG64 stream)."""
        import ctypes as C

        from . import _capi

        if width != height:
            raise ValueError("rooms maps must be square (the reference transposes half of them)")
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64).ravel())
        bits = np.empty((len(idx), height, (width + 31) // 32), dtype=np.uint32)
        _capi.check(_capi.lib().apg_maps_generate_rooms(int(device), idx.ctypes.data_as(C.c_void_p), len(idx), int(width),
                                                        int(height), int(max_rooms), int(door_width),
                                                        bits.ctypes.data_as(C.c_void_p)))
        return cls(bits, int(width), int(height))

    @classmethod
    def from_maps(cls, maps: np.ndarray) -> "FloorMapBank":
        maps = np.asarray(maps, dtype=bool)
        return cls(pack_maps(maps), maps.shape[-1], maps.shape[-2])

    @classmethod
    def from_dataset(cls, dataset, indices: Sequence[int] | None = None) -> "FloorMapBank":
        if isinstance(dataset, FloorMapBank) and indices is None:
            return dataset
        if hasattr(dataset, "load"):
            dataset.load()
        if indices is None:
            if len(dataset) > MAX_IMPLICIT_BANK:
                raise ValueError(
                    f"dataset has {len(dataset)} maps; pass a finite bank (FloorMapBank.from_dataset(ds, indices))")
            indices = range(len(dataset))
        maps = np.stack([np.asarray(dataset[int(i)], dtype=bool) for i in indices])
        return cls(pack_maps(maps), int(dataset.map_width), int(dataset.map_height))

    @classmethod
    def load_npz(cls, path: str, limit: int | None = None) -> "FloorMapBank":
        z = np.load(path)
        bits = z["bits"] if limit is None else z["bits"][:limit]
        return cls(bits, int(z["width"]), int(z["height"]))
