"""ctypes binding of libapg_b200.so (include/apg_b200.h).  No CPU fallback: importing works anywhere, but
`lib()` raises when the shared library is missing and every engine call fails loudly without a GPU."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# APG_B200_LIB: developer override (A/B-testing a differently compiled build); the default is the in-tree library
LIB_PATH = os.environ.get("APG_B200_LIB") or os.path.join(_PKG, "_lib", "libapg_b200.so")

APG_OK = 0
ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NEED_PAYLOAD, ERR_NAN_ACTION, ERR_NAN_PRED, ERR_OUT_OF_GRID = range(-1, -8, -1)
FLAG_NAN_ACTION, FLAG_NAN_PRED, FLAG_OUT_OF_GRID, FLAG_MAPGEN = 1, 2, 4, 8
TASK_CLASSIFICATION, TASK_LOCALIZATION = 0, 1
BANK_U8, BANK_F32 = 0, 1
PAYLOAD_RING_DEFAULT, PAYLOAD_RING_MAX = 4, 64
NUM_STATS = 8


class ApgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libapg_b200 error {code}: {msg}")
        self.code = code


class ImageCfg(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("num_envs", C.c_int32), ("task", C.c_int32),
        ("height", C.c_int32), ("width", C.c_int32), ("channels", C.c_int32),
        ("sensor_h", C.c_int32), ("sensor_w", C.c_int32),
        ("sensor_scale", C.c_double), ("max_step_length", C.c_double * 2),
        ("step_limit", C.c_int32), ("num_classes", C.c_int32), ("bank_dtype", C.c_int32),
        ("randomly_invert_labels", C.c_int32), ("track_visited", C.c_int32), ("payload_ring", C.c_int32), ("blend_f32", C.c_int32), ("sparse_reward", C.c_int32),
    ]


class ImageBuffers(C.Structure):
    _fields_ = [
        ("d_glimpse", C.c_void_p), ("d_target_glimpse", C.c_void_p), ("d_glimpse_pos", C.c_void_p),
        ("d_time_step", C.c_void_p), ("d_inverted_label", C.c_void_p), ("d_reward", C.c_void_p),
        ("d_base_reward", C.c_void_p), ("d_loss", C.c_void_p), ("d_target_label", C.c_void_p),
        ("d_target_xy", C.c_void_p), ("d_terminated", C.c_void_p), ("d_truncated", C.c_void_p),
        ("d_index", C.c_void_p), ("d_pos", C.c_void_p), ("d_cur_target_xy", C.c_void_p),
        ("d_visited", C.c_void_p), ("d_stamp_xy", C.c_void_p), ("d_stamp_q", C.c_void_p),
        ("d_ep_avg1", C.c_void_p), ("d_ep_final1", C.c_void_p), ("d_ep_avg2", C.c_void_p), ("d_ep_final2", C.c_void_p),
        ("d_ep_return", C.c_void_p), ("d_first_correct", C.c_void_p), ("d_last_incorrect", C.c_void_p),
        ("stamp_count", C.c_int32), ("elapsed", C.c_int32), ("needs_reset", C.c_int32),
    ]


class Pcg64State(C.Structure):
    _fields_ = [("state_hi", C.c_uint64), ("state_lo", C.c_uint64), ("inc_hi", C.c_uint64), ("inc_lo", C.c_uint64),
                ("has_uint32", C.c_uint32), ("uinteger", C.c_uint32)]

    @classmethod
    def from_numpy(cls, gen):
        st = gen.bit_generator.state
        s, inc = int(st["state"]["state"]), int(st["state"]["inc"])
        m = (1 << 64) - 1
        return cls(s >> 64, s & m, inc >> 64, inc & m, int(st["has_uint32"]), int(st["uinteger"]))

    def to_numpy(self, gen):
        st = gen.bit_generator.state
        st["state"]["state"] = (int(self.state_hi) << 64) | int(self.state_lo)
        st["state"]["inc"] = (int(self.inc_hi) << 64) | int(self.inc_lo)
        st["has_uint32"], st["uinteger"] = int(self.has_uint32), int(self.uinteger)
        gen.bit_generator.state = st


class LidarCfg(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("num_envs", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("beam_count", C.c_int32), ("lidar_range", C.c_float), ("step_limit", C.c_int32),
        ("static_map", C.c_int32), ("static_map_index", C.c_int32), ("emit_cells", C.c_int32),
        ("env_offset", C.c_int64), ("sparse_reward", C.c_int32), ("reserved", C.c_int32),
    ]


class LidarBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "d_lidar", "d_odometry", "d_time_step", "d_map", "d_reward", "d_base_reward", "d_loss", "d_target",
        "d_terminated", "d_truncated", "d_stepped", "d_map_idx", "d_pos", "d_initial_pos",
        "d_ep_avg1", "d_ep_final1", "d_ep_avg2", "d_ep_final2", "d_ep_return", "d_cells", "d_dists")]


# name -> (restype, argtypes); every symbol include/apg_b200.h declares must be listed here
_P = C.c_void_p
SYMBOLS = {
    "apg_version": (C.c_int, []),
    "apg_last_error": (C.c_char_p, []),
    "apg_launch_count": (C.c_int64, []),
    "apg_image_create": (C.c_int, [C.POINTER(ImageCfg), C.POINTER(_P)]),
    "apg_image_destroy": (None, [_P]),
    "apg_image_set_stream": (C.c_int, [_P, _P]),
    "apg_image_set_bank": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int]),
    "apg_image_reset": (C.c_int, [_P, _P, _P, _P, _P]),
    "apg_image_queue_reset_payload": (C.c_int, [_P, _P, _P, _P, _P]),
    "apg_image_pending_payloads": (C.c_int, [_P]),
    "apg_image_set_streams": (C.c_int, [_P, C.POINTER(Pcg64State), C.POINTER(Pcg64State), C.POINTER(Pcg64State),
                                        C.c_int64, C.c_int64]),
    "apg_image_get_streams": (C.c_int, [_P, C.POINTER(Pcg64State), C.POINTER(Pcg64State), C.POINTER(Pcg64State)]),
    "apg_image_prefetch_payloads": (C.c_int, [_P, C.c_int32]),
    "apg_image_step": (C.c_int, [_P, _P, _P]),
    "apg_image_rollout": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P]),
    "apg_image_step_host": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "apg_image_unique_topk": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int32, C.c_int32, _P]),
    "apg_image_get_buffers": (C.c_int, [_P, C.POINTER(ImageBuffers)]),
    "apg_image_get_stats": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int]),
    "apg_image_poll_error": (C.c_int, [_P, C.POINTER(C.c_uint32), C.c_int]),
    "apg_image_sync": (C.c_int, [_P]),
    "apg_image_variant": (C.c_int, [_P]),
    "apg_lidar_create": (C.c_int, [C.POINTER(LidarCfg), C.POINTER(_P)]),
    "apg_lidar_destroy": (None, [_P]),
    "apg_lidar_set_stream": (C.c_int, [_P, _P]),
    "apg_lidar_set_map_generator": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_double]),
    "apg_lidar_set_map_bank": (C.c_int, [_P, _P, C.c_int64]),
    "apg_lidar_set_beams": (C.c_int, [_P, _P]),
    "apg_lidar_reset": (C.c_int, [_P, C.c_uint64]),
    "apg_lidar_reset_all": (C.c_int, [_P]),
    "apg_lidar_step": (C.c_int, [_P, _P, _P]),
    "apg_lidar_rollout": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P]),
    "apg_lidar_step_host": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "apg_lidar_get_buffers": (C.c_int, [_P, C.POINTER(LidarBuffers)]),
    "apg_lidar_get_stats": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int]),
    "apg_lidar_poll_error": (C.c_int, [_P, C.POINTER(C.c_uint32), C.c_int]),
    "apg_lidar_sync": (C.c_int, [_P]),
    "apg_struct_sizes": (None, [C.POINTER(C.c_int32)]),
    "apg_rng_host_stream": (C.c_int, [C.c_uint64, C.c_int32, _P, _P, _P]),
    "apg_rng_device_payload": (C.c_int, [C.c_int32, _P, _P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_int32,
                                         C.c_int32, _P, _P, _P, _P]),
    "apg_maps_generate_maze": (C.c_int, [C.c_int32, _P, C.c_int64, C.c_int32, C.c_int32, C.c_double, _P]),
    "apg_maps_generate_rooms": (C.c_int, [C.c_int32, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "apg_dlpack_wrap": (_P, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads the library once.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m ap_gym_b200._build` (needs nvcc). "
                "ap_gym_b200 has no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int):
    if rc != APG_OK:
        raise ApgError(rc, lib().apg_last_error().decode("utf-8", "replace"))


# ---- DLPack: engine-owned device memory -> torch tensor (zero copy) ------------------------------------
_DL_CODES = {"int": 0, "uint": 1, "float": 2, "bool": 6}
C.pythonapi.PyCapsule_New.restype = C.py_object
C.pythonapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]


def dlpack_tensor(ptr: int, device: int, kind: str, bits: int, shape):
    """A torch tensor viewing `ptr` (no copy, no ownership transfer)."""
    import torch.utils.dlpack

    shape_arr = (C.c_int64 * len(shape))(*shape)
    managed = lib().apg_dlpack_wrap(ptr, device, _DL_CODES[kind], bits, len(shape), shape_arr)
    if not managed:
        raise ApgError(ERR_INVALID, "apg_dlpack_wrap failed")
    capsule = C.pythonapi.PyCapsule_New(managed, b"dltensor", None)
    return torch.utils.dlpack.from_dlpack(capsule)
