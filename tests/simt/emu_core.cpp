// TEST INFRASTRUCTURE: the one translation unit of the emulated libfplgpu that owns the emulator's state — the fiber scheduler
// (emu_cuda_impl.h), the cp.async / mbarrier emulation, and NVRTC + the three driver calls fpl_jit.cu makes, where "compiling"
// a kernel means building its source for the host with g++ behind the emulator and "loading the cubin" means dlopen.
#include "emu_cuda_impl.h"
#include "fake/cuda.h"
#include "fake/nvrtc.h"
#include <dlfcn.h>
#include <mutex>
#include <string>

// ---------------- cp.async: completion at the latest legal moment (default) or at issue ----------------
namespace emu {
namespace {
int cp_lazy = 1;
struct Copy { uint8_t* dst; const uint8_t* src; int n, size; };
struct Queue { unsigned long long serial = 0; std::vector<Copy> open; std::vector<std::vector<Copy>> groups; };
Queue queues[1024];
Queue& q() { Queue& x = queues[tid_in_block()]; if (x.serial != block_serial) { x.serial = block_serial; x.open.clear(); x.groups.clear(); } return x; }
void now(const Copy& c) { if (c.n > 0) memcpy(c.dst, c.src, c.n); memset(c.dst + c.n, 0, c.size - c.n); }
struct Mbar { uint32_t phase; int32_t pending; };       // pending: bytes still to arrive, +2^30 while the arrival is outstanding
void settle(Mbar* m) { if (m->pending == 0) { m->phase++; m->pending = 1 << 30; note_progress(); } }
}  // namespace
void cp_async(uint32_t dst, const void* src, int n, int size) {
    Copy c = {(uint8_t*)from_shared(dst), (const uint8_t*)src, n, size};
    if (cp_lazy) q().open.push_back(c); else now(c);
}
void cp_commit() { if (cp_lazy) { Queue& x = q(); x.groups.push_back(x.open); x.open.clear(); } }
void cp_wait(int n_pending) {
    if (!cp_lazy) return;
    Queue& x = q();
    while ((int)x.groups.size() > n_pending) { for (const Copy& c : x.groups.front()) now(c); x.groups.erase(x.groups.begin()); }
}
void mbar_init(uint32_t bar) { Mbar* m = (Mbar*)from_shared(bar); m->phase = 0; m->pending = 1 << 30; }
void mbar_expect(uint32_t bar, uint32_t bytes) { Mbar* m = (Mbar*)from_shared(bar); m->pending += (int32_t)bytes - (1 << 30); settle(m); }
void mbar_wait(uint32_t bar, uint32_t parity) { Mbar* m = (Mbar*)from_shared(bar); while ((m->phase & 1u) == parity) yield(); }
void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    memcpy(from_shared(dst), src, bytes);
    Mbar* m = (Mbar*)from_shared(bar); m->pending -= (int32_t)bytes; settle(m);
}
}  // namespace emu
extern "C" void emu_set_cp_async_lazy(int on) { emu::cp_lazy = on; }
extern "C" long long emu_collectives() { return emu::collectives; }
extern "C" long long emu_blocks() { return (long long)emu::block_serial; }

// ---------------- NVRTC + driver: kernels "compiled" for the host ----------------
struct EmuNvrtcProgram { std::string src, log, so; };
static std::mutex jit_mu;
extern "C" {
nvrtcResult nvrtcCreateProgram(nvrtcProgram* prog, const char* src, const char*, int, const char* const*, const char* const*) {
    *prog = new EmuNvrtcProgram();
    (*prog)->src = src;
    return NVRTC_SUCCESS;
}
nvrtcResult nvrtcDestroyProgram(nvrtcProgram* prog) { delete *prog; *prog = nullptr; return NVRTC_SUCCESS; }
nvrtcResult nvrtcCompileProgram(nvrtcProgram p, int, const char* const*) {
    std::lock_guard<std::mutex> lock(jit_mu);
    unsigned long long h = 1469598103934665603ull;
    for (unsigned char ch : p->src) h = (h ^ ch) * 1099511628211ull;
    char base[256];
    snprintf(base, sizeof(base), "/tmp/fpl_emu_jit_%016llx_%s", h, EMU_BUILD_TAG);
    p->so = std::string(base) + ".so";
    if (FILE* f = fopen(p->so.c_str(), "rb")) { fclose(f); return NVRTC_SUCCESS; }
    const std::string cpp = std::string(base) + ".cpp";
    FILE* f = fopen(cpp.c_str(), "w");
    if (!f) { p->log = "cannot write " + cpp; return 1; }
    fputs("#include \"emu_cuda.h\"\nnamespace jit {\n", f);
    fputs(p->src.c_str(), f);
    fputs("\n}  // namespace jit\n"
          "extern \"C\" void jit_scan(unsigned grid, unsigned block, unsigned smem, const uint8_t* seq, const uint8_t* qual, const int64_t* offsets,\n"
          "                         void* st, int64_t n, unsigned one) {\n"
          "    EMU_LAUNCH((grid), (block), (smem), jit::k_scan_jit(seq, qual, (const jit::int64_t*)offsets, (jit::ReadState*)st, (jit::int64_t)n, one));\n"
          "}\n", f);
    fclose(f);
    const std::string tmp = p->so + ".tmp";
    const std::string cmd = std::string("g++ -std=c++17 -O1 -fPIC -shared -w -I ") + EMU_SIMT_DIR + " -o " + tmp + " " + cpp + " " + EMU_LIB_PATH +
                            " -Wl,-rpath," + EMU_LIB_DIR + " > " + base + ".log 2>&1 && mv " + tmp + " " + p->so;
    if (system(cmd.c_str()) != 0) { p->log = "g++ failed: see " + std::string(base) + ".log"; return 1; }
    return NVRTC_SUCCESS;
}
nvrtcResult nvrtcGetProgramLogSize(nvrtcProgram p, size_t* n) { *n = p->log.size() + 1; return NVRTC_SUCCESS; }
nvrtcResult nvrtcGetProgramLog(nvrtcProgram p, char* log) { memcpy(log, p->log.c_str(), p->log.size() + 1); return NVRTC_SUCCESS; }
nvrtcResult nvrtcGetCUBINSize(nvrtcProgram p, size_t* n) { *n = p->so.size() + 1; return NVRTC_SUCCESS; }
nvrtcResult nvrtcGetCUBIN(nvrtcProgram p, char* cubin) { memcpy(cubin, p->so.c_str(), p->so.size() + 1); return NVRTC_SUCCESS; }
const char* nvrtcGetErrorString(nvrtcResult r) { return r == NVRTC_SUCCESS ? "NVRTC_SUCCESS" : "emulated NVRTC: host compile failed"; }

CUresult cuModuleLoadData(CUmodule* mod, const void* image) { *mod = dlopen((const char*)image, RTLD_NOW | RTLD_LOCAL); return *mod ? CUDA_SUCCESS : 200; }
CUresult cuModuleGetFunction(CUfunction* fn, CUmodule mod, const char*) { *fn = dlsym(mod, "jit_scan"); return *fn ? CUDA_SUCCESS : 500; }
CUresult cuLaunchKernel(CUfunction f, unsigned gx, unsigned, unsigned, unsigned bx, unsigned, unsigned, unsigned smem, CUstream, void** args, void**) {
    typedef void (*fn_t)(unsigned, unsigned, unsigned, const uint8_t*, const uint8_t*, const int64_t*, void*, int64_t, unsigned);
    ((fn_t)f)(gx, bx, smem, *(const uint8_t**)args[0], *(const uint8_t**)args[1], *(const int64_t**)args[2], *(void**)args[3], *(int64_t*)args[4],
              *(unsigned*)args[5]);
    return CUDA_SUCCESS;
}
CUresult cuGetErrorString(CUresult, const char** s) { *s = "emulated driver error"; return CUDA_SUCCESS; }
}
