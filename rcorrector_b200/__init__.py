"""rcorrector_b200 — B200-native (sm_100a) implementation of Rcorrector's stage-3
per-read error-correction hot path.  The product is `librcx.so` (C ABI,
include/rcx.h) and the drop-in `bin/rcorrector` host binary; this package only
holds their sources (csrc/) and a ctypes binding (api.py)."""
from .api import INTERLEAVED, SINGLE, Corrector, RcxError, load_library  # noqa: F401
