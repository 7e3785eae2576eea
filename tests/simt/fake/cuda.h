// TEST INFRASTRUCTURE: the handful of driver-API names fpl_jit.cu uses (resolved with dlsym; tests/simt/emu_jit.cpp defines them)
#pragma once
#include "cuda_runtime.h"
typedef void* CUmodule;
typedef void* CUfunction;
typedef void* CUstream;
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
