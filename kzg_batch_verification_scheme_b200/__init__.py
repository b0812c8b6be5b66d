"""kzgb200 -- B200-native KZG batch verification (BLS12-381).  See DESIGN.md.

Public host API: `load()` -> KzgLib (ctypes over include/kzgb200.h); `KzgLib.context()` -> Context with
`verify_kzg_proof` / `verify_kzg_proof_batch` (BASELINE.json:5).
"""
from .api import (CHUNK, KZGB_BADARGS, KZGB_ERROR, KZGB_OK, PARTIAL_BYTES, Context, KzgError, KzgLib, load)  # noqa: F401
