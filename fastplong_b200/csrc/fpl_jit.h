#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
struct DevBatch;
struct ReadState;
struct FplJitKernel { void* fn = nullptr; };
int fpl_jit_build_scan(int device, const char* a0, const char* a1, bool doAdapters, bool doCounts, bool doCplx, int qq,
                       FplJitKernel* out, char* err, size_t errlen);
int fpl_jit_launch_scan(const FplJitKernel* k, const DevBatch& b, ReadState* st, cudaStream_t stream);
