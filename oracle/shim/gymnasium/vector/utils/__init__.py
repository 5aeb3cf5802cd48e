"""gymnasium.vector.utils subset (TEST-ONLY shim): batch_space, iterate, concatenate, create_empty_array."""
from __future__ import annotations

from . import space_utils
from .space_utils import batch_space, concatenate, create_empty_array, iterate
