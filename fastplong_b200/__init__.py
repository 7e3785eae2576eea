"""fastplong_b200 — B200-native per-read hot loop of fastplong behind a C ABI (include/fplgpu.h).

The package holds only what the path needs: csrc/ (sm_100a CUDA kernels + the C-ABI library),
host/ (the C++ SingleEndProcessor drop-in that links against the unmodified reference sources),
and a thin ctypes mirror of the ABI used by tests, bench.py and the multi-GPU driver.
"""
from .abi import (FplOptions, FplAdapters, FplBatch, RESULT_DTYPE, stats_words, counter_words,  # noqa: F401
                  MAX_ADAPTER_LEN)
from .options import Options  # noqa: F401
from .pack import PackedBatch, pack_reads  # noqa: F401
