// Run-time specialisation of the adapter+quality scan (k_scan_jit, source in fpl_scan_jit_src.h): NVRTC compiles it
// for sm_100a with the two adapter strings and the option flags as constants, the driver API loads and launches it.
// Modules are cached per (device, specialisation) for the life of the process.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "fpl_device.cuh"
#include "fpl_jit.h"
#include "fpl_scan_jit_src.h"

#include <dlfcn.h>

namespace {
struct Cached { CUmodule mod; CUfunction fn; };
std::mutex g_mu;
std::map<std::string, Cached> g_cache;

// The driver library is resolved lazily: libfplgpu.so must stay loadable on a machine without a GPU driver
// (build check, symbol test), where only fpl_create() fails.
struct Driver {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                             void**, void**) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    bool ok = false;
    Driver() {
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        ModuleLoadData = (decltype(ModuleLoadData))dlsym(h, "cuModuleLoadData");
        ModuleGetFunction = (decltype(ModuleGetFunction))dlsym(h, "cuModuleGetFunction");
        LaunchKernel = (decltype(LaunchKernel))dlsym(h, "cuLaunchKernel");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "cuGetErrorString");
        ok = ModuleLoadData && ModuleGetFunction && LaunchKernel && GetErrorString;
    }
};
Driver& driver() { static Driver d; return d; }
}  // namespace

int fpl_jit_build_scan(int device, const char* a0, const char* a1, bool doAdapters, bool doCounts, bool doCplx, int qq,
                       FplJitKernel* out, char* err, size_t errlen) {
    out->fn = nullptr;
    std::string defs;
    defs += std::string("#define FPL_A0 \"") + (doAdapters ? a0 : "") + "\"\n";
    defs += std::string("#define FPL_A1 \"") + (doAdapters ? a1 : "") + "\"\n";
    defs += std::string("#define FPL_DO_ADAPTERS ") + (doAdapters ? "true" : "false") + "\n";
    defs += std::string("#define FPL_DO_COUNTS ") + (doCounts ? "true" : "false") + "\n";
    defs += std::string("#define FPL_DO_CPLX ") + (doCplx ? "true" : "false") + "\n";
    defs += "#define FPL_QQ " + std::to_string(qq & 0x7f) + "\n";
    const char* mb = getenv("FPL_JIT_MINBLOCKS");   // occupancy knob (registers per thread); 8 (64 registers) measured best on B200: 4/5/6/8 blocks -> 17.8/15.7/14.5/13.7 ms
    defs += std::string("#define FPL_MINBLOCKS ") + (mb && atoi(mb) > 0 ? std::to_string(atoi(mb)) : std::string("8")) + "\n";
    const std::string key = std::to_string(device) + "|" + defs;
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) { out->fn = (void*)it->second.fn; return 0; }

    const std::string src = defs + kScanJitSource;
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, src.c_str(), "fpl_scan_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
        snprintf(err, errlen, "nvrtcCreateProgram failed");
        return -1;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
    nvrtcResult rc = nvrtcCompileProgram(prog, 3, opts);
    if (rc != NVRTC_SUCCESS) {
        size_t n = 0;
        nvrtcGetProgramLogSize(prog, &n);
        std::vector<char> log(n + 1, 0);
        nvrtcGetProgramLog(prog, log.data());
        snprintf(err, errlen, "NVRTC: %s: %.300s", nvrtcGetErrorString(rc), log.data());
        nvrtcDestroyProgram(&prog);
        return -1;
    }
    size_t n = 0;
    nvrtcGetCUBINSize(prog, &n);
    std::vector<char> cubin(n);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    // the runtime API has already made the device's primary context current (cudaSetDevice + allocations in fpl_create)
    Cached c;
    if (!driver().ok) { snprintf(err, errlen, "libcuda.so.1 not loadable"); return -1; }
    CUresult cr = driver().ModuleLoadData(&c.mod, cubin.data());
    if (cr == CUDA_SUCCESS) cr = driver().ModuleGetFunction(&c.fn, c.mod, "k_scan_jit");
    if (cr != CUDA_SUCCESS) {
        const char* s = nullptr;
        driver().GetErrorString(cr, &s);
        snprintf(err, errlen, "loading the JIT module failed: %s", s ? s : "?");
        return -1;
    }
    g_cache[key] = c;
    out->fn = (void*)c.fn;
    return 0;
}

int fpl_jit_launch_scan(const FplJitKernel* k, const DevBatch& b, ReadState* st, cudaStream_t stream) {
    if (b.n_reads == 0) return 0;
    const uint8_t* seq = b.seq;
    const uint8_t* qual = b.qual;
    const int64_t* offsets = b.offsets;
    int64_t n = b.n_reads;
    void* args[] = {(void*)&seq, (void*)&qual, (void*)&offsets, (void*)&st, (void*)&n};
    // one CTA per read.  (Measured: persistent CTAs walking the reads round-robin are slower — 14.7 / 15.7 / 17.1 ms with
    // 32 / 16 / 8 CTAs per SM against 13.8 ms — the hardware scheduler balances the gamma-distributed read lengths better.)
    CUresult cr = driver().LaunchKernel((CUfunction)k->fn, (unsigned)b.n_reads, 1, 1, 128, 1, 1, 0, (CUstream)stream, args, nullptr);
    return cr == CUDA_SUCCESS ? 0 : -1;
}
