"""phnrec_b200 — B200-native PhnRec recognition hot path.

The product is `lib/libphnrec_b200.so` (hand-written sm_100a CUDA kernels behind the C ABI of
include/phnrec_b200.h) and `bin/phnrec` (the drop-in CLI, C++).  This package is the thin
Python binding used by tests and bench.py; there is no CPU implementation here.
"""
from .api import (LABEL_DTYPE, MLP_EXACT_FP32, MLP_TC_F16, WAVE_ALAW, WAVE_LIN16, PhnRecError, Recognizer, build,
                  format_mlf_entry, format_rec, lib_path, read_htk, write_htk)

__all__ = ["Recognizer", "PhnRecError", "LABEL_DTYPE", "MLP_EXACT_FP32", "MLP_TC_F16", "WAVE_LIN16", "WAVE_ALAW",
           "build", "lib_path", "format_rec", "format_mlf_entry", "read_htk", "write_htk"]
