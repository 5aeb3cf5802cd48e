"""TEST-ONLY minimal stand-in for the `gymnasium` package (absent from this image, no network).

It exists so that the UNMODIFIED reference (`/root/reference/ap_gym`) can be imported and
executed to pin the oracle and to generate the golden vectors under `tests/golden/`.
It restates only the documented gymnasium >= 1.1 semantics that the reference touches
(SURVEY.md section 4.3).  It is NOT product code: nothing under `active-perception-gym_b200/`
imports it.
"""
from . import error, spaces, utils, vector, envs
from .core import Env, Wrapper
from .spaces import Space
from .envs.registration import register, registry, make, make_vec, VectorizeMode, spec

__version__ = "1.1.0+shim"
