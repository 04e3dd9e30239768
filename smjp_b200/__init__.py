"""smjp_b200: B200-native engine for the gauenk/smjp hot path (Rao-Teh FFBS sweep + particle filter).

Host code is Python and mirrors the reference's `pkg` call signatures; the arithmetic runs in
hand-written sm_100a CUDA kernels behind a C ABI (include/smjp_b200.h, smjp_b200/csrc)."""
from ._lib import EngineError, build, load  # noqa: F401

__version__ = "0.1.0"
