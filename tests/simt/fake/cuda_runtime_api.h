// TEST INFRASTRUCTURE: <cuda_runtime_api.h> for the emulated build of the drop-in binary (tests/simt_emu.py:build_binary): the
// few runtime calls host/seprocessor_gpu.cpp makes itself (pinned buffers, device count, driver warm-up), on host memory.
// Deliberately free of the emulator's device-side names (min / max templates, threadIdx ...): this header is included next to
// the reference's own headers.
#pragma once
#include <stdlib.h>
#include <string.h>
enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated runtime error"; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) {      // EMU_MALLOC_FILL: pinned memory is not zeroed either
    static const char* f = getenv("EMU_MALLOC_FILL");
    *p = (T*)malloc(n ? n : 1);
    if (*p) memset(*p, f ? (int)strtol(f, NULL, 0) : 0, n ? n : 1);
    return *p ? cudaSuccess : cudaErrorInvalidValue;
}
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
