// TEST INFRASTRUCTURE: the scheduler of tests/simt/emu_cuda.h (include once per harness translation unit).
#pragma once
#include "emu_cuda.h"
#include <mutex>

uint3 emu_threadIdx, emu_blockIdx;
dim3 emu_blockDim, emu_gridDim;

namespace emu {
long long collectives = 0;
unsigned long long block_serial = 0;
uint8_t* dynamic_smem = nullptr;
namespace {
std::vector<char*> windows;          // base of each 16 MB shared window
std::vector<uint8_t> dyn;
constexpr size_t STACK = 512 * 1024;
struct Fiber { ucontext_t ctx; bool done; };
ucontext_t sched_ctx;
std::vector<Fiber> fibers;
std::vector<char*> stacks;          // pooled, never shrunk
std::vector<Group> warps;
Group blk;
const std::function<void()>* body_ = nullptr;
int cur = -1;
bool progress = false;

void complete(Group& g, int n) {
    memcpy(g.res, g.slot, sizeof(uint64_t) * (size_t)n);
    memset(g.slot, 0, sizeof(uint64_t) * (size_t)n);
    g.res_mask = (unsigned)g.in_lo;
    g.in_lo = 0;
    g.arrived = 0;
    g.gen++;
    progress = true;
}
void fiber_main() {
    (*body_)();
    Fiber& f = fibers[cur];
    f.done = true;
    progress = true;
    Group& w = warps[cur >> 5];
    w.alive--; blk.alive--;
    if (w.arrived > 0 && w.arrived == w.alive) complete(w, 32);
    if (blk.arrived > 0 && blk.arrived == blk.alive) complete(blk, (int)emu_blockDim.x);
    // returning switches to sched_ctx (uc_link)
}
}  // namespace

uint32_t to_shared(const void* p) {
    const char* c = (const char*)p;
    for (size_t i = 0; i < windows.size(); i++)
        if (c >= windows[i] && c < windows[i] + (1 << 24)) return (uint32_t)(((i + 1) << 24) | (size_t)(c - windows[i]));
    windows.push_back((char*)((uintptr_t)c & ~(uintptr_t)0xFFFF));          // windows start on 64 KB boundaries
    if (windows.size() > 200) { fprintf(stderr, "emu: too many shared windows\n"); abort(); }
    return to_shared(p);
}
void* from_shared(uint32_t a) {
    const size_t i = (a >> 24);
    if (i == 0 || i > windows.size()) { fprintf(stderr, "emu: bad shared address %08x\n", a); abort(); }
    return windows[i - 1] + (a & 0xFFFFFFu);
}
void yield() { swapcontext(&fibers[cur].ctx, &sched_ctx); }
void note_progress() { progress = true; }
Group& warp() { return warps[cur >> 5]; }
Group& block() { return blk; }
int lane() { return cur & 31; }
int tid_in_block() { return cur; }

Group& collect(Group& g, int index, uint64_t v) {
    collectives++;
    g.slot[index] = v;
    if (index < 64) g.in_lo |= 1ull << index;
    g.arrived++;
    const unsigned my = g.gen;
    if (g.arrived == g.alive) complete(g, &g == &blk ? (int)emu_blockDim.x : 32);
    else while (g.gen == my) yield();
    return g;
}

static int sched_mode = [] {
    const char* m = getenv("EMU_SCHED");
    return !m ? 0 : !strncmp(m, "reverse", 7) ? 1 : !strncmp(m, "random", 6) ? 2 : 0;
}();
static unsigned long long sched_rng = [] { const char* m = getenv("EMU_SCHED"); const char* c = m ? strchr(m, ':') : nullptr; return c ? strtoull(c + 1, nullptr, 0) * 2654435761ull + 1 : 88172645463325252ull; }();
static std::vector<int> perm;
static void shuffle(int n) {
    perm.resize((size_t)n);
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int i = n - 1; i > 0; i--) {
        sched_rng ^= sched_rng << 13; sched_rng ^= sched_rng >> 7; sched_rng ^= sched_rng << 17;
        const int j = (int)(sched_rng % (unsigned long long)(i + 1));
        const int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
}
static std::mutex launch_mu;      // one kernel at a time: callers on several host threads (one context each) take turns
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    std::lock_guard<std::mutex> lock(launch_mu);
    if (block.y != 1 || block.z != 1 || grid.z != 1 || block.x > 1024) { fprintf(stderr, "emu: 1-D blocks, 2-D grids only\n"); abort(); }
    if (dyn.size() < smem_bytes + 64) dyn.resize(smem_bytes + 64);
    dynamic_smem = (uint8_t*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    const int n = (int)block.x;
    emu_blockDim = block; emu_gridDim = grid;
    body_ = &body;
    while ((int)stacks.size() < n) stacks.push_back((char*)malloc(STACK));
    for (unsigned long long bb = 0; bb < (unsigned long long)grid.x * grid.y; bb++) {
        const unsigned b = (unsigned)(bb % grid.x);
        emu_blockIdx = {b, (unsigned)(bb / grid.x), 0};
        block_serial++;
        { static const char* f = getenv("EMU_SMEM_FILL"); memset(dynamic_smem, f ? (int)strtol(f, nullptr, 0) : 0xCD, smem_bytes); }   // shared memory is not zero at block start
        fibers.assign((size_t)n, Fiber());
        warps.assign((size_t)((n + 31) / 32), Group());
        for (int w = 0; w < (int)warps.size(); w++) warps[w].alive = (w * 32 + 32 <= n) ? 32 : n - w * 32;
        blk = Group();
        blk.alive = n;
        for (int i = 0; i < n; i++) {
            fibers[i].done = false;
            getcontext(&fibers[i].ctx);
            fibers[i].ctx.uc_stack.ss_sp = stacks[i];
            fibers[i].ctx.uc_stack.ss_size = STACK;
            fibers[i].ctx.uc_link = &sched_ctx;
            makecontext(&fibers[i].ctx, (void (*)())fiber_main, 0);
        }
        int left = n;
        bool first = true;
        while (left > 0) {
            progress = first;
            first = false;
            left = 0;
            for (int k = 0; k < n; k++) {
                // which live thread runs next is the scheduler's business: EMU_SCHED=reverse / random[:seed] change the order in which
                // the threads of a block get their turns (results must not depend on it: an ordering assumption without a barrier shows)
                int i = k;
                if (sched_mode == 1) i = n - 1 - k;
                else if (sched_mode == 2) { if (k == 0) shuffle(n); i = perm[k]; }
                if (fibers[i].done) continue;
                cur = i;
                emu_threadIdx = {(unsigned)i, 0, 0};
                swapcontext(&sched_ctx, &fibers[i].ctx);
                if (!fibers[i].done) left++;
            }
            if (left > 0 && !progress) {
                fprintf(stderr, "emu: deadlock in block %u: %d threads wait at a collective the others never reach "
                                "(divergent __syncthreads / *_sync)\n", b, left);
                abort();
            }
        }
    }
    cur = -1;
}
}  // namespace emu
