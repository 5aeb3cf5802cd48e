"""ap_gym_b200 -- B200-native batched stepping engine for Active Perception Gym (ap_gym 0.5.0).

Drop-in vector environments for the reference's image and LIDAR hot paths on hand-written sm_100a CUDA
kernels behind a C-ABI (include/apg_b200.h).  Importing this package never touches the GPU; constructing an
environment requires the built library and a CUDA device -- there is no CPU fallback.
"""
from .loss_fn import (CrossEntropyLossFn, LambdaLossFn, LossFn, LossFnAffineTransformation, MSELossFn,
                      WeightedLossFn, ZeroLossFn)
from .spaces import ActivePerceptionActionSpace, ImageSpace, LogitSpace
from .datasets import ArrayImageDataset, ImageBank, SyntheticImageDataset, bank_from_dataset
from .floor_maps import FloorMapBank, FloorMapDatasetMaze, FloorMapDatasetRooms, pack_maps, unpack_maps
from .lidar_env import LIDARLocalization2DVectorEnv, lidar_directions
from .image_env import ImageClassificationVectorEnv, ImageLocalizationVectorEnv, ImagePerceptionConfig
from .reference_compat import enable_reference_compat, register_envs
from .hide_and_seek import CircleSquareHideAndSeekVectorWrapper

__version__ = "0.1.0"
