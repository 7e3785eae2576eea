// TEST INFRASTRUCTURE: NVRTC for the emulated build — "compiling" a kernel means building it for the host with g++ behind
// the SIMT emulator (tests/simt/emu_jit.cpp); the "cubin" is the path of the shared object.
#pragma once
#include <stddef.h>
typedef int nvrtcResult;
enum { NVRTC_SUCCESS = 0 };
typedef struct EmuNvrtcProgram* nvrtcProgram;
extern "C" {
nvrtcResult nvrtcCreateProgram(nvrtcProgram* prog, const char* src, const char* name, int nh, const char* const* headers, const char* const* names);
nvrtcResult nvrtcDestroyProgram(nvrtcProgram* prog);
nvrtcResult nvrtcCompileProgram(nvrtcProgram prog, int nopt, const char* const* opts);
nvrtcResult nvrtcGetProgramLogSize(nvrtcProgram prog, size_t* n);
nvrtcResult nvrtcGetProgramLog(nvrtcProgram prog, char* log);
nvrtcResult nvrtcGetCUBINSize(nvrtcProgram prog, size_t* n);
nvrtcResult nvrtcGetCUBIN(nvrtcProgram prog, char* cubin);
const char* nvrtcGetErrorString(nvrtcResult r);
}
