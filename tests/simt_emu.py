"""TEST INFRASTRUCTURE: the default path of libfplgpu executed on the CPU.

The source text of fpl_device.cuh, fpl_trim.cu (k_trim, k_trim_fasta), fpl_scan.cu (the generic k_scan, k_final, k_count) and
fpl_stats.cu (k_make_preseg, k_cs_keys / k_cs_gather, k_cycle_stats, k_kmer_fix, k_read_qual) is preprocessed — CUDA includes
dropped, `mad.lo.u32` inline PTX rewritten as C, the small PTX wrappers (cp.async, ld.shared, red.shared, prmt, mbarrier / bulk
copy) replaced by host forms, `extern __shared__` bound to the emulator's buffer, `kernel<<<grid, block, smem, stream>>>(args)`
turned into a call of the SIMT emulator (tests/simt/emu_cuda.h: one fiber per CUDA thread, warp / block collectives with their
real semantics) — and compiled with g++ together with the launch functions of those files and the table builder cut out of
fpl_create (fpl_api.cu), so that adapters, thresholds, match masks, grids and kernel variants are chosen by the product's own
code.  `EmuEngine.process()` runs the kernels in run_batch's order on host memory and returns records, both Stats blocks and
the counter vector, to be compared with the oracle like a GPU result.

Also under the emulator: k_scan_jit v2 as fpl_jit.cu generates it for the options' adapters (jit_scan), k_scan_fast, the
--mask/--break kernels (fpl_ext.cu) and the FASTQ text path (fpl_ingest.cu, fpl_emit.cu).  Not covered: the NCCL merge, the
upload overlap of fpl_process_host, k_eval_kmers, cub's own kernels (host stand-ins); and what no functional emulation shows:
timing, bank conflicts, memory ordering between warps.
Nothing here is shipped or reachable from the product path.
"""
import ctypes as C
import hashlib
import os
import re
import subprocess

import numpy as np

from device_helpers import CSRC, ROOT, _asm_to_c, _match_brace
from fastplong_b200 import abi
from fastplong_b200.abi import FplBatch, RESULT_DTYPE

SIMT = os.path.join(ROOT, "tests", "simt")


def _drop_function(text, name):
    m = re.search(r"^(?:template\s*<[^>\n]*>\s*\n)?[A-Za-z_][^\n;{}()]*\b" + name + r"\s*\([^;{}]*\)\s*\{", text, re.M)
    if not m:
        raise KeyError(name)
    end = _match_brace(text, m.end() - 1)
    return text[:m.start()] + text[end + 1:]


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def _rewrite_launches(text):
    out, i = "", 0
    while True:
        j = text.find("<<<", i)
        if j < 0:
            return out + text[i:]
        k = text.index(">>>", j)
        # kernel name (with template arguments) in front of <<<
        m = re.search(r"([A-Za-z_]\w*(?:<[^<>;]*>)?)\s*$", text[:j])
        cfg = _split_top(text[j + 3:k])
        p = text.index("(", k)
        q = p
        depth = 0
        while True:
            depth += text[q] == "("
            depth -= text[q] == ")"
            if depth == 0:
                break
            q += 1
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += text[i:m.start()] + f"EMU_LAUNCH(({cfg[0]}), ({cfg[1]}), ({smem}), {m.group(1)}{text[p:q + 1]})"
        i = q + 1


def device_text(fn, drop=()):
    text = open(os.path.join(CSRC, fn)).read()
    text = re.sub(r'^\s*#include\s+[<"](cuda_runtime\.h|fpl_device\.cuh|cuda\.h|cub/[\w/.]+|fpl_ext\.h|fpl_ingest\.h|fpl_emit\.h)[>"].*$', "", text, flags=re.M)
    text = text.replace("#pragma once", "")
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];", r"\1* \2 = (\1*)emu::dynamic_smem;", text)
    text = re.sub(r'asm\s+volatile\s*\(\s*"fence[^"]*"[^;]*;', ";", text)
    for name in drop:
        text = _drop_function(text, name)
    return f"// ======== {fn} (preprocessed) ========\n" + _rewrite_launches(_asm_to_c(text))


def table_builder():
    """fpl_create's own table building: from `// host tables` to `stamp("tables");`"""
    text = open(os.path.join(CSRC, "fpl_api.cu")).read()
    a = text.index("    // host tables")
    b = text.index('    stamp("tables");', a)
    return text[a:b]


HARNESS = r"""
#include <math.h>
#include <stdarg.h>
#include <algorithm>
#include <string>
#include "emu_cuda_impl.h"
#include "fplgpu.h"
#include "fpl_scanplan.h"

// ---- a host-memory stand-in for the runtime calls inside the table builder and the launch functions ----
enum cudaError_t { cudaSuccess = 0 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaStreamNonBlocking = 1,
       cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
template <class T> static cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)calloc(n ? n : 1, 1); return cudaSuccess; }
static cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memcpy(d, s, n); return cudaSuccess; }
static cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return cudaSuccess; }
static cudaError_t cudaGetLastError() { return cudaSuccess; }
static cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class F> static cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
namespace cub {
struct DeviceRadixSort {      // a stable LSD radix sort on key bits [begin_bit, end_bit), like the library's
    template <class K, class V>
    static cudaError_t SortPairsDescending(void* tmp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, int n, int begin_bit,
                                           int end_bit, cudaStream_t) {
        if (!tmp) { bytes = 64; return cudaSuccess; }
        std::vector<int> idx((size_t)n);
        for (int i = 0; i < n; i++) idx[i] = i;
        const K mask = end_bit - begin_bit >= (int)(8 * sizeof(K)) ? ~(K)0 : (K)((((K)1 << (end_bit - begin_bit)) - 1) << begin_bit);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return (kin[a] & mask) > (kin[b] & mask); });
        for (int i = 0; i < n; i++) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
        return cudaSuccess;
    }
};
template <class T> struct CountingInputIterator {
    T base;
    explicit CountingInputIterator(T b) : base(b) {}
    T operator[](int64_t i) const { return base + (T)i; }
};
struct DeviceSelect {
    template <class In, class Out, class Num, class Pred>
    static cudaError_t If(void* tmp, size_t& bytes, In in, Out out, Num num_out, int64_t n, Pred pred, cudaStream_t = nullptr) {
        if (!tmp) { bytes = 64; return cudaSuccess; }
        int64_t k = 0;
        for (int64_t i = 0; i < n; i++) { const auto v = in[i]; if (pred(v)) out[k++] = v; }
        *num_out = k;
        return cudaSuccess;
    }
};
struct DeviceScan {
    template <class I, class O>
    static cudaError_t ExclusiveSum(void* tmp, size_t& bytes, I in, O out, int64_t n, cudaStream_t = nullptr) {
        if (!tmp) { bytes = 64; return cudaSuccess; }
        typename std::remove_reference<decltype(out[0])>::type acc = 0;
        for (int64_t i = 0; i < n; i++) { auto v = in[i]; out[i] = acc; acc += v; }       // in may alias out
        return cudaSuccess;
    }
};
}  // namespace cub

// ---- host forms of the PTX wrappers of fpl_device.cuh / fpl_stats.cu (their definitions are cut out of the text) ----
static inline uint32_t shared_addr(const void* p) { return emu::to_shared(p); }
static inline void red_shared_add(uint32_t a, uint32_t v) { *(uint32_t*)emu::from_shared(a) += v; }
template <int IMM> static inline void red_shared_add_imm(uint32_t a, uint32_t v) { *(uint32_t*)emu::from_shared(a + (uint32_t)IMM) += v; }
static inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {       // prmt.b32, default mode
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t d = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t c = (sel >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(src >> (8 * (c & 7))) & 0xFF;
        if (c & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        d |= byte << (8 * i);
    }
    return d;
}
// cp.async completes at an unspecified time between its issue and the wait that covers its group.  Both extremes are
// emulated: eager (at issue) and lazy (at that wait, the default) — under the lazy order a thread that reads bytes another
// thread copied, without a barrier behind that thread's wait, sees the stale slot (the hazard racecheck reports on the GPU).
static int g_cp_lazy = 1;
extern "C" void emu_set_cp_async_lazy(int on) { g_cp_lazy = on; }
struct EmuCopy { uint8_t* dst; const uint8_t* src; int n, size; };
struct EmuCopyQueue { unsigned long long serial = 0; std::vector<EmuCopy> open; std::vector<std::vector<EmuCopy>> groups; };
static EmuCopyQueue g_cpq[1024];
static inline EmuCopyQueue& cpq() { EmuCopyQueue& q = g_cpq[emu::tid_in_block()]; if (q.serial != emu::block_serial) { q.serial = emu::block_serial; q.open.clear(); q.groups.clear(); } return q; }
static inline void emu_copy_now(const EmuCopy& c) { if (c.n > 0) memcpy(c.dst, c.src, c.n); memset(c.dst + c.n, 0, c.size - c.n); }
static inline void cp_async_any(uint32_t dst, const void* src, int n, int size) {
    EmuCopy c = {(uint8_t*)emu::from_shared(dst), (const uint8_t*)src, n, size};
    if (g_cp_lazy) cpq().open.push_back(c); else emu_copy_now(c);
}
static inline void cp_async16(uint32_t dst, const void* src, int n) { cp_async_any(dst, src, n, 16); }
static inline void cp_async4(uint32_t dst, const void* src, int n) { cp_async_any(dst, src, n, 4); }
static inline void cp_async_commit() { if (g_cp_lazy) { EmuCopyQueue& q = cpq(); q.groups.push_back(q.open); q.open.clear(); } }
template <int N> static inline void cp_async_wait() {      // all but the N most recent groups of this thread are complete
    if (!g_cp_lazy) return;
    EmuCopyQueue& q = cpq();
    while ((int)q.groups.size() > N) { for (const EmuCopy& c : q.groups.front()) emu_copy_now(c); q.groups.erase(q.groups.begin()); }
}
static inline uint4 lds128(uint32_t a) { uint4 v; memcpy(&v, emu::from_shared(a), 16); return v; }
static inline uint32_t lds32(uint32_t a) { uint32_t v; memcpy(&v, emu::from_shared(a), 4); return v; }
// mbarrier with one expected arrival + a transaction count, in its 8 bytes of shared memory: {completed phases, pending}.
// A bulk copy completes at issue (one legal order); the phase completes when the arrival has happened and no bytes are pending.
struct EmuMbar { uint32_t phase; int32_t pending; };       // pending: bytes still to arrive, +2^30 while the arrival is outstanding
static inline void mbar_init(uint32_t bar, uint32_t) { EmuMbar* m = (EmuMbar*)emu::from_shared(bar); m->phase = 0; m->pending = 1 << 30; }
static inline void mbar_settle(EmuMbar* m) { if (m->pending == 0) { m->phase++; m->pending = 1 << 30; emu::note_progress(); } }
static inline void mbar_expect_tx(uint32_t bar, uint32_t bytes) { EmuMbar* m = (EmuMbar*)emu::from_shared(bar); m->pending += (int32_t)bytes - (1 << 30); mbar_settle(m); }
static inline void mbar_wait(uint32_t bar, uint32_t parity) { EmuMbar* m = (EmuMbar*)emu::from_shared(bar); while ((m->phase & 1u) == parity) emu::yield(); }
static inline void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    memcpy(emu::from_shared(dst), src, bytes);
    EmuMbar* m = (EmuMbar*)emu::from_shared(bar); m->pending -= (int32_t)bytes; mbar_settle(m);
}

@@DEVICE@@

struct EmuCtx {
    DevParams P; ScanPlan plan; int n_adapters = 0; cudaStream_t stream = nullptr;
    uint8_t* d_adapters = nullptr; int* d_alen = nullptr; uint4* d_peq = nullptr; uint32_t* d_peq16 = nullptr; uint32_t* d_acode = nullptr;
    unsigned long long* d_peq_long = nullptr; int* d_pf_order = nullptr; unsigned long long* d_counters = nullptr; int64_t counter_words = 0;
};
static char g_err[512];
static int fail(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); return -1; }
static void fpl_destroy(EmuCtx* c) {
    free(c->d_adapters); free(c->d_alen); free(c->d_peq); free(c->d_peq16); free(c->d_acode); free(c->d_peq_long); free(c->d_pf_order);
    free(c->d_counters); delete c;
}

extern "C" const char* emu_last_error() { return g_err; }
extern "C" long long emu_collectives() { return emu::collectives; }

typedef void (*scan_fn)(const uint8_t*, const uint8_t*, const int64_t*, void*, int64_t);

static EmuCtx* emu_create(const fpl_options* opt, const fpl_adapters* ad) {
    g_err[0] = 0;
    const int n = 2 + (ad->n_fasta > 0 ? ad->n_fasta : 0);
    EmuCtx* c = new EmuCtx();
    c->n_adapters = n;
    auto stamp = [](const char*) {};
    auto body = [&]() -> int {
@@BUILDER@@
        return 0;
    };
    if (body()) return nullptr;         // (the builder's error paths have destroyed c)
    return c;
}

// run_batch's kernel order (fpl_api.cu) over a batch in host memory.  stats0 / stats1: FPL_STATS_WORDS(C) words each, zeroed by the
// caller; null: records and counters only (the Stats kernels are skipped, the two median fields stay 0).  keep: where the --mask/--break
// lists stay alive for the caller (freed by the caller).
static int emu_run(EmuCtx* c, const DevBatch& b, int64_t n_bytes, fpl_read_result* results, unsigned long long* stats0,
                   unsigned long long* stats1, int64_t C, scan_fn scan_override, extns::FplExt* keep) {
    const int64_t nr = b.n_reads;
    std::vector<ReadState> st((size_t)nr + 1);
    std::vector<StatSeg> pre((size_t)nr + 1), post(2 * (size_t)nr + 2);
    int64_t tmax = 0;
    bool slots16 = true;
    for (int64_t i = 0; i < nr; i++) { if (b.lens[i] > tmax) tmax = b.lens[i]; if (b.offsets[i] & 15) slots16 = false; }
    if (stats0 && tmax > C) return fail("C too small");
    memset(results, 0xAB, sizeof(fpl_read_result) * (size_t)nr);          // k_trim must write every record
    CycleWs ws;
    cudaStream_t s = nullptr;
    const bool ext = c->P.opt.mask_enabled || c->P.opt.break_enabled;   // variable number of output reads: run_batch's other branch
    if (stats0) launch_make_preseg(b, pre.data(), s);
    launch_trim(c->P, b, st.data(), results, c->d_counters, s);
    if (stats0 && launch_cycle_stats(&ws, b.seq, b.qual, pre.data(), nr, tmax, stats0, C, true,
                                     ext ? nullptr : stats1 + 16 * C + FPL_STATS_KMER, slots16, s))
        return fail("launch_cycle_stats(pre) failed");
    if (scan_override == (scan_fn)1) {          // the precompiled bit-sliced kernel (FPL_NO_JIT), where fpl_create would use it
        if (c->plan.fast) scanfast::launch_scan_fast(c->P, c->plan, b, st.data(), s); else launch_scan(c->P, b, st.data(), s);
    } else if (scan_override) { if (nr) scan_override(b.seq, b.qual, b.offsets, st.data(), nr); }     // k_scan_jit (jit_scan below)
    else launch_scan(c->P, b, st.data(), s);
    launch_final(c->P, b, st.data(), results, post.data(), s);
    if (ext) {          // --mask / --break (needs the Stats blocks)
        if (!stats0) return fail("--mask/--break needs the Stats blocks");
        extns::FplExt local;
        extns::FplExt& x = keep ? *keep : local;
        char xerr[256] = "";
        const uint8_t* fseq = b.seq;
        launch_read_qual(b, stats0, stats1, C, results, true, s);
        launch_count(results, nr, c->d_counters, false, s);
        if (extns::fpl_ext_run(&x, c->P, b, n_bytes, results, c->d_counters, stats1, C, &fseq, s, xerr, sizeof(xerr))) return fail("ext: %s", xerr);
        if (launch_cycle_stats(&ws, fseq, b.qual, x.d_stat, x.n_segs, tmax, stats1, C, true, nullptr, false, s)) return fail("launch_cycle_stats(ext) failed");
        if (!keep) extns::fpl_ext_free(&x);
        fpl_cycle_ws_free(&ws);
        return 0;
    }
    launch_count(results, nr, c->d_counters, true, s);
    if (stats0) {
        if (launch_cycle_stats(&ws, b.seq, b.qual, post.data(), 2 * nr, tmax, stats1, C, false, nullptr, false, s))
            return fail("launch_cycle_stats(post) failed");
        launch_kmer_fix(b, results, stats1 + 16 * C + FPL_STATS_KMER, s);
        launch_read_qual(b, stats0, stats1, C, results, false, s);
        fpl_cycle_ws_free(&ws);
    }
    return 0;
}

// fpl_process_host's work (minus the copies): a packed batch -> records, counters, Stats, -N/-b lists
extern "C" int emu_process(const fpl_options* opt, const fpl_adapters* ad, const fpl_batch* hb, fpl_read_result* results,
                           unsigned long long* counters, int64_t n_counter_words, unsigned long long* stats0,
                           unsigned long long* stats1, int64_t C, scan_fn scan_override, int* plan_fast, fpl_segment* segs_out,
                           int64_t segs_cap, int64_t* n_segs, fpl_region* regs_out, int64_t regs_cap, int64_t* n_regs) {
    EmuCtx* c = emu_create(opt, ad);
    if (!c) return -1;
    if (plan_fast) *plan_fast = c->plan.fast;
    if (!hb) { fpl_destroy(c); return 0; }                              // only the tables were wanted
    if (n_counter_words != c->counter_words) { fpl_destroy(c); return fail("counter words %lld != %lld", (long long)n_counter_words, (long long)c->counter_words); }
    DevBatch b = {hb->seq, hb->qual, hb->offsets, hb->lens, hb->n_reads};
    extns::FplExt x;
    int rc = emu_run(c, b, hb->n_bytes, results, stats0, stats1, C, scan_override, &x);
    if (!rc) {
        if (n_segs) *n_segs = x.n_segs;
        if (n_regs) *n_regs = x.n_regs;
        if (x.n_segs > segs_cap || x.n_regs > regs_cap) rc = fail("segment / region capacity");
        else {
            if (x.n_segs) memcpy(segs_out, x.d_segs, sizeof(fpl_segment) * (size_t)x.n_segs);
            if (x.n_regs) memcpy(regs_out, x.d_regs, sizeof(fpl_region) * (size_t)x.n_regs);
        }
        memcpy(counters, c->d_counters, sizeof(unsigned long long) * (size_t)c->counter_words);
    }
    extns::fpl_ext_free(&x);
    fpl_destroy(c);
    return rc;
}

// fpl_process_fastq_host + fpl_emit_fastq_host (fpl_api.cu) on host memory: a chunk of FASTQ text -> record table, records,
// counters, Stats and the --out / --failed_out text.  Returns 1 where the device parser refuses the layout.
extern "C" int emu_process_fastq(const fpl_options* opt, const fpl_adapters* ad, const uint8_t* text, int64_t n_bytes, int is_last,
                                 fpl_fastq_record* records, fpl_read_result* results, int64_t max_records, int64_t* n_records,
                                 int64_t* consumed, unsigned long long* counters, unsigned long long* stats0, unsigned long long* stats1,
                                 int64_t C, scan_fn scan_override, int want_failed, uint8_t* out, int64_t out_cap, int64_t* out_bytes,
                                 uint8_t* failed, int64_t failed_cap, int64_t* failed_bytes) {
    EmuCtx* c = emu_create(opt, ad);
    if (!c) return -1;
    ingestns::FplIngest g;
    char err[256] = "";
    int64_t nrec = 0;
    *out_bytes = *failed_bytes = 0;
    int rc = ingestns::fpl_ingest_index(&g, text, n_bytes, is_last, nullptr, &nrec, consumed, err, sizeof(err));
    if (rc < 0) { fpl_destroy(c); return fail("ingest: %s", err); }
    if (rc > 0 || nrec > max_records) { ingestns::fpl_ingest_free(&g); fpl_destroy(c); return 1; }
    *n_records = nrec;
    std::vector<uint8_t> d_seq((size_t)g.packed_bytes + 64, 0), d_qual((size_t)g.packed_bytes + 64, 0);
    extns::FplExt x;
    emitns::FplEmit e;
    if (nrec) {
        if (ingestns::fpl_ingest_pack(&g, nrec, d_seq.data(), d_qual.data(), nullptr, err, sizeof(err))) rc = fail("pack: %s", err);
        memcpy(records, g.d_rec, sizeof(fpl_fastq_record) * (size_t)nrec);
        DevBatch b = {d_seq.data(), d_qual.data(), g.d_offsets, g.d_lens, nrec};
        if (!rc) rc = emu_run(c, b, g.packed_bytes, results, stats0, stats1, C, scan_override, &x);
        if (!rc) {
            const bool ext = c->P.opt.mask_enabled || c->P.opt.break_enabled;
            emitns::EmitSource src;
            src.text = g.d_text; src.rec = g.d_rec; src.res = results; src.n_reads = nrec;
            src.segs = ext ? x.d_segs : nullptr;
            src.seg_off = ext ? x.d_off : nullptr;
            src.mseq = (c->P.opt.mask_enabled && x.n_segs > 0) ? x.d_mseq : nullptr;
            src.offsets = g.d_offsets;
            if (emitns::fpl_emit_build(&e, src, want_failed != 0, nullptr, err, sizeof(err))) rc = fail("emit: %s", err);
            else if (e.out_bytes > out_cap || e.failed_bytes > failed_cap) rc = fail("text capacity");
            else {
                *out_bytes = e.out_bytes; *failed_bytes = e.failed_bytes;
                if (e.out_bytes) memcpy(out, e.d_out, (size_t)e.out_bytes);
                if (e.failed_bytes) memcpy(failed, e.d_failed, (size_t)e.failed_bytes);
            }
        }
    }
    memcpy(counters, c->d_counters, sizeof(unsigned long long) * (size_t)c->counter_words);
    emitns::fpl_emit_free(&e);
    extns::fpl_ext_free(&x);
    ingestns::fpl_ingest_free(&g);
    fpl_destroy(c);
    return rc;
}
"""

PTX_WRAPPERS = ("prmt", "cp_async16", "cp_async4", "cp_async_commit", "cp_async_wait", "lds128", "lds32", "mbar_init", "mbar_expect_tx",
                "mbar_wait", "bulk_g2s", "red_shared_add_imm")


def source():
    dev = "\n".join([device_text("fpl_device.cuh", drop=("red_shared_add", "shared_addr")), device_text("fpl_trim.cu"),
                     device_text("fpl_scan.cu"), "namespace scanfast {", device_text("fpl_scan_fast.cu"), "}  // namespace scanfast",
                     '#include "fpl_stats.h"', device_text("fpl_stats.cu", drop=PTX_WRAPPERS), "namespace extns {",
                     device_text("fpl_ext.h"), device_text("fpl_ext.cu"), "}  // namespace extns", "namespace ingestns {",
                     device_text("fpl_ingest.h"), device_text("fpl_ingest.cu"), "}  // namespace ingestns", "namespace emitns {",
                     device_text("fpl_emit.h"), device_text("fpl_emit.cu"), "}  // namespace emitns", device_text("fpl_eval.cu")])
    return HARNESS.replace("@@DEVICE@@", dev).replace("@@BUILDER@@", table_builder())


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = source()
    for h in ("emu_cuda.h", "emu_cuda_impl.h"):
        src_h = open(os.path.join(SIMT, h)).read()
        src += "\n// " + hashlib.md5(src_h.encode()).hexdigest()
    so = f"/tmp/fpl_simt_{hashlib.md5(src.encode()).hexdigest()[:12]}.so"
    if not os.path.exists(so):
        cpp = so[:-3] + ".cpp"
        open(cpp, "w").write(src)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-w", "-I", SIMT, "-I", os.path.join(ROOT, "include"),
                               "-I", CSRC, "-o", so, cpp])
    lib = C.CDLL(so)
    lib.emu_process_fastq.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.fpl_eval_adapter_kmers.argtypes = [C.c_int, C.POINTER(FplBatch), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    lib.emu_last_error.restype = C.c_char_p
    lib.emu_set_cp_async_lazy.argtypes = [C.c_int]
    lib.emu_collectives.restype = C.c_longlong
    lib.emu_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                C.c_int64, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p, C.c_int64,
                                C.POINTER(C.c_int64)]
    _lib = lib
    return lib


JIT_HARNESS = r"""
#include "emu_cuda_impl.h"
namespace jit {
@@DEFS@@
@@SOURCE@@
}  // namespace jit
extern "C" void jit_scan(const uint8_t* seq, const uint8_t* qual, const int64_t* offsets, void* st, int64_t n) {
    EMU_LAUNCH(((unsigned)n), (128), (0), jit::k_scan_jit(seq, qual, (const jit::int64_t*)offsets, (jit::ReadState*)st, (jit::int64_t)n, 1u));
}
"""

_jit_cache = {}


def jit_scan(options):
    """k_scan_jit v2 for these options, as fpl_create would have NVRTC build it (fpl_jit.cu: fpl_jit_build_scan's #defines +
    the generated source of fpl_jit_debug_source), compiled for the host under the emulator.  Returns the ctypes function to hand
    to emu_process, or None where the library would not specialise (an adapter that is empty, not ACGT-only or > 128 bp)."""
    from fastplong_b200.binding import load_library
    o, ad, keep = options.to_abi()
    a0, a1 = (ad.start or b"").decode(), (ad.end or b"").decode()
    do_adapters = bool(o.adapter_enabled)
    if do_adapters and not all(1 <= len(a) <= 128 and set(a) <= set("ACGT") for a in (a0, a1)):
        return None
    if not do_adapters:
        a0 = a1 = ""
    do_counts = bool(o.qual_filter_enabled or o.length_filter_enabled)
    key = (a0, a1, do_adapters, do_counts, bool(o.complexity_enabled), o.qualified_qual & 0x7F)
    if key in _jit_cache:
        return _jit_cache[key]
    lib = load_library()
    lib.fpl_jit_debug_source.restype = C.c_char_p
    lib.fpl_jit_debug_source.argtypes = [C.c_char_p, C.c_char_p]
    body = lib.fpl_jit_debug_source(a0.encode(), a1.encode()).decode()
    t = lambda x: "true" if x else "false"
    defs = (f'#define FPL_A0 "{a0}"\n#define FPL_A1 "{a1}"\n#define FPL_ALEN0 {len(a0)}\n#define FPL_ALEN1 {len(a1)}\n#define FPL_JIT_VERSION 2\n'
            f'#define FPL_DO_ADAPTERS {t(do_adapters)}\n#define FPL_DO_COUNTS {t(do_counts)}\n#define FPL_DO_CPLX {t(o.complexity_enabled)}\n'
            f'#define FPL_QQ {o.qualified_qual & 0x7F}\n#define FPL_MINBLOCKS 8\n')
    src = JIT_HARNESS.replace("@@DEFS@@", defs).replace("@@SOURCE@@", _asm_to_c(body))
    for h in ("emu_cuda.h", "emu_cuda_impl.h"):
        src += "\n// " + hashlib.md5(open(os.path.join(SIMT, h)).read().encode()).hexdigest()
    so = f"/tmp/fpl_simt_jit_{hashlib.md5(src.encode()).hexdigest()[:12]}.so"
    if not os.path.exists(so):
        cpp = so[:-3] + ".cpp"
        open(cpp, "w").write(src)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-w", "-I", SIMT, "-o", so, cpp])
    fn = C.CDLL(so).jit_scan
    _jit_cache[key] = fn
    return fn


class EmuEngine:
    """binding.Engine's process() / stats() / counters() for the emulated kernels.  with_stats=False: records and counters only
    (the Stats kernels are skipped; the two median fields of the records stay 0)."""

    def __init__(self, options, with_stats=True, scan="generic"):
        """scan = "generic": launch_scan (k_scan); "fast": launch_scan_fast (the precompiled k_scan_fast that FPL_NO_JIT selects) where
        ScanPlan.fast allows it; "jit": k_scan_jit v2 specialised on the options where the library would
        specialise (self.jit tells), the generic kernel elsewhere."""
        self.lib = load()
        self.scan_fn = jit_scan(options) if scan == "jit" else C.cast(1, C.c_void_p) if scan == "fast" else None
        self.jit = scan == "jit" and self.scan_fn is not None
        self.options = options
        self._abi = options.to_abi()
        self.n_adapters = 2 + len(options.adapter_fasta)
        self._counters = np.zeros(abi.counter_words(self.n_adapters), dtype=np.int64)
        self.with_stats = with_stats
        self._blocks = None        # [pre, post] at self._C cycles

    def process(self, batch):
        o, ad, keep = self._abi
        res = np.zeros(batch.n_reads, dtype=RESULT_DTYPE)
        cnt = np.zeros_like(self._counters)
        b = batch.to_abi()
        s0 = s1 = None
        cyc = 0
        if self.with_stats:
            need = max(64, int(batch.lens.max()) if batch.n_reads else 1)
            cyc = 1 << int(np.ceil(np.log2(need)))
            s0 = np.zeros(abi.stats_words(cyc), dtype=np.int64)
            s1 = np.zeros(abi.stats_words(cyc), dtype=np.int64)
        segs = np.zeros(64 * batch.n_reads + 4096 if (self.options.mask or self.options.break_reads) else 1, dtype=abi.SEGMENT_DTYPE)
        regs = np.zeros(len(segs), dtype=abi.REGION_DTYPE)
        n_segs, n_regs = C.c_int64(0), C.c_int64(0)
        rc = self.lib.emu_process(C.addressof(o), C.addressof(ad), C.addressof(b), res.ctypes.data, cnt.ctypes.data, cnt.shape[0],
                                  s0.ctypes.data if s0 is not None else None, s1.ctypes.data if s1 is not None else None, cyc,
                                  C.cast(self.scan_fn, C.c_void_p) if self.scan_fn is not None else None, None,
                                  segs.ctypes.data, segs.shape[0], C.byref(n_segs), regs.ctypes.data, regs.shape[0], C.byref(n_regs))
        self._segs, self._regs = segs[:n_segs.value].copy(), regs[:n_regs.value].copy()
        if rc != 0:
            raise RuntimeError(self.lib.emu_last_error().decode())
        self._counters += cnt
        if self.with_stats:
            from fastplong_b200.binding import relayout_stats
            if self._blocks is None:
                self._blocks, self._C = [s0, s1], cyc
            else:
                c2 = max(cyc, self._C)
                self._blocks = [relayout_stats(x, self._C, c2) + relayout_stats(y, cyc, c2) for x, y in zip(self._blocks, (s0, s1))]
                self._C = c2
        return res

    def process_fastq(self, text, is_last=True, want_failed=True):
        """binding.Engine.process_fastq + emit_fastq in one call: a chunk of plain FASTQ text through k_count_lines / the newline
        select / k_fastq_records / k_fastq_pack, the kernels of process(), k_emit_sizes / k_emit_copy.  Returns None where the
        device parser refuses the layout, else (record table, records, bytes consumed, --out text, --failed_out text)."""
        from fastplong_b200.abi import FASTQ_RECORD_DTYPE
        from fastplong_b200.binding import relayout_stats
        o, ad, keep = self._abi
        buf = np.frombuffer(bytes(text) + b"\0" * 64, dtype=np.uint8)[:len(text)]
        cap = max(16, len(text) // 8)
        recs = np.zeros(cap, dtype=FASTQ_RECORD_DTYPE)
        res = np.zeros(cap, dtype=RESULT_DTYPE)
        cnt = np.zeros_like(self._counters)
        longest = max((len(ln) for ln in bytes(text).split(b"\n")), default=1)
        cyc = 1 << int(np.ceil(np.log2(max(64, longest))))
        s0, s1 = np.zeros(abi.stats_words(cyc), dtype=np.int64), np.zeros(abi.stats_words(cyc), dtype=np.int64)
        out, failed = np.zeros(2 * len(text) + 4096, dtype=np.uint8), np.zeros(2 * len(text) + 4096, dtype=np.uint8)
        n, used, nout, nfail = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        rc = self.lib.emu_process_fastq(C.addressof(o), C.addressof(ad), buf.ctypes.data if len(text) else None, len(text), int(is_last),
                                        recs.ctypes.data, res.ctypes.data, cap, C.byref(n), C.byref(used), cnt.ctypes.data, s0.ctypes.data,
                                        s1.ctypes.data, cyc, C.cast(self.scan_fn, C.c_void_p) if self.scan_fn is not None else None,
                                        int(want_failed), out.ctypes.data, out.shape[0], C.byref(nout), failed.ctypes.data, failed.shape[0],
                                        C.byref(nfail))
        if rc == 1:
            return None
        if rc != 0:
            raise RuntimeError(self.lib.emu_last_error().decode())
        self._counters += cnt
        self._blocks, self._C = [s0, s1], cyc
        return recs[:n.value], res[:n.value], used.value, out[:nout.value].tobytes(), failed[:nfail.value].tobytes()

    def segments(self):
        """--mask/--break: every output read of the last process() call (fpl_last_segments)"""
        return self._segs

    def mask_regions(self):
        return self._regs

    def plan_fast(self):
        """fpl_create's own verdict (ScanPlan.fast): may the bit-sliced scan kernels be used for these adapters"""
        o, ad, keep = self._abi
        v = C.c_int(-1)
        rc = self.lib.emu_process(C.addressof(o), C.addressof(ad), None, None, None, 0, None, None, 0, None, C.byref(v), None, 0, None,
                                  None, 0, None)
        assert rc == 0
        return bool(v.value)

    def stats(self, which, cycles):
        from fastplong_b200.binding import relayout_stats
        if self._blocks is None:
            return np.zeros(abi.stats_words(cycles), dtype=np.int64)
        return relayout_stats(self._blocks[which], self._C, int(cycles))

    def counters(self):
        return self._counters.copy()


def eval_adapter_kmers(batch, side, shift_tail=1):
    """binding.eval_adapter_kmers through the emulated k_eval_kmers (fpl_eval.cu with its own host function)"""
    lib = load()
    counts = np.zeros(1 << 20, dtype=np.uint32)
    acc = np.zeros(1 << 20, dtype=np.uint64)
    total = C.c_int64()
    b = batch.to_abi()
    rc = lib.fpl_eval_adapter_kmers(0, C.byref(b), int(shift_tail), int(side), counts.ctypes.data, acc.ctypes.data, C.byref(total))
    if rc != 0:
        raise RuntimeError(f"fpl_eval_adapter_kmers failed ({rc})")
    return counts, acc, int(total.value)


# =====================================================================================================================
# The whole library on the CPU: every .cu of fastplong_b200/csrc — fpl_api.cu's C ABI and host logic included — built for
# the host against tests/simt/fake/ (a host-memory <cuda_runtime.h>, cub stand-ins, NVRTC that compiles with g++) and the
# SIMT emulator.  The result exports include/fplgpu.h like libfplgpu.so does, so binding.Engine, the C tests and the
# drop-in binary run on it unchanged.
# =====================================================================================================================
LIB_SOURCES = ("fpl_api.cu", "fpl_trim.cu", "fpl_scan.cu", "fpl_scan_fast.cu", "fpl_stats.cu", "fpl_jit.cu", "fpl_ingest.cu", "fpl_ext.cu",
               "fpl_eval.cu", "fpl_emit.cu")


def _lib_text(fn):
    text = open(os.path.join(CSRC, fn)).read()
    drops = {"fpl_device.cuh": ("red_shared_add", "shared_addr"), "fpl_stats.cu": PTX_WRAPPERS}.get(fn, ())
    for name in drops:
        text = _drop_function(text, name)
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];", r"\1* \2 = (\1*)emu::dynamic_smem;", text)
    text = re.sub(r'asm\s+volatile\s*\(\s*"fence[^"]*"[^;]*;', ";", text)
    text = text.replace('dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL)', "dlopen(EMU_LIB_PATH, RTLD_NOW)")
    return _rewrite_launches(_asm_to_c(text))


_emu_lib_path = None


def build_library():
    """-> path of libfplgpu_emu.so (built once per source state into /tmp)"""
    global _emu_lib_path
    if _emu_lib_path:
        return _emu_lib_path
    files = {}
    for fn in sorted(os.listdir(CSRC)):
        if fn.endswith((".cu", ".cuh", ".h")):
            files[fn] = _lib_text(fn)
    extra = ""
    for root, _, names in os.walk(SIMT):
        for nm in sorted(names):
            extra += open(os.path.join(root, nm)).read()
    san = os.environ.get("FPL_EMU_SANITIZE", "")          # e.g. "undefined": the kernels and fpl_api.cu's host code under UBSan
    tag = hashlib.md5(("".join(files[k] for k in sorted(files)) + extra + san).encode()).hexdigest()[:12]
    out = f"/tmp/fpl_emu_lib_{tag}"
    so = os.path.join(out, "libfplgpu_emu.so")
    if not os.path.exists(so):
        os.makedirs(out, exist_ok=True)
        for fn, text in files.items():
            open(os.path.join(out, fn[:-3] + ".cpp" if fn.endswith(".cu") else fn), "w").write(text)
        defs = [f'-DEMU_LIB_PATH="{so}"', f'-DEMU_LIB_DIR="{out}"', f'-DEMU_SIMT_DIR="{SIMT}"', f'-DEMU_BUILD_TAG="{tag}"']
        inc = ["-I", out, "-I", os.path.join(SIMT, "fake"), "-I", SIMT, "-I", os.path.join(ROOT, "include")]
        objs = []
        jobs = []
        for src in [os.path.join(out, fn[:-3] + ".cpp") for fn in LIB_SOURCES] + [os.path.join(SIMT, "emu_core.cpp")]:
            obj = os.path.join(out, os.path.basename(src)[:-4] + ".o")
            objs.append(obj)
            sflags = [f"-fsanitize={san}", "-fno-omit-frame-pointer"] if san else []        # reports go to stderr, the run goes on
            jobs.append(subprocess.Popen(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-w", "-c", *sflags, *defs, *inc, "-o", obj, src],
                                         stderr=subprocess.PIPE, text=True))
        for j in jobs:
            err = j.communicate()[1]
            if j.returncode != 0:
                raise RuntimeError("emulated build failed:\n" + err[:6000])
        subprocess.check_call(["g++", "-shared", *([f"-fsanitize={san}"] if san else []), "-o", so + ".tmp", *objs, "-ldl", "-lpthread"])
        os.replace(so + ".tmp", so)
    _emu_lib_path = so
    return so


_emu_bin_path = None


def build_binary():
    """-> path of fastplong_gpu_emu: host/seprocessor_gpu.cpp (unchanged) + the reference's unmodified objects (oracle/_ref/obj, as
    fastplong_b200/host/Makefile links them) + the emulated library.  The drop-in CLI without a GPU."""
    global _emu_bin_path
    if _emu_bin_path:
        return _emu_bin_path
    lib = build_library()
    out = os.path.dirname(lib)
    exe = os.path.join(out, "fastplong_gpu_emu")
    host = os.path.join(ROOT, "fastplong_b200", "host", "seprocessor_gpu.cpp")
    objdir = os.path.join(ROOT, "oracle", "_ref", "obj")
    stamp = os.path.join(out, "host.md5")
    tag = hashlib.md5(open(host, "rb").read()).hexdigest()
    if not os.path.exists(exe) or not os.path.exists(stamp) or open(stamp).read() != tag:
        if not os.path.exists(os.path.join(objdir, "main.o")):
            raise RuntimeError("oracle/_ref/obj is not built (make -C oracle)")
        obj = os.path.join(out, "seprocessor_gpu.o")
        subprocess.check_call(["g++", "-std=c++14", "-pthread", "-O2", "-w", "-I", os.path.join(SIMT, "fake"), "-I", os.path.join(ROOT, "oracle", "shim"),
                               "-I", "/root/reference/src", "-I", os.path.join(ROOT, "include"), "-c", host, "-o", obj])
        refobj = sorted(os.path.join(objdir, f) for f in os.listdir(objdir) if f.endswith(".o") and f != "seprocessor.o")
        subprocess.check_call(["g++", "-pthread", "-o", exe, obj, *refobj, lib, "-lz", "-ldl", f"-Wl,-rpath,{out}"])
        open(stamp, "w").write(tag)
    _emu_bin_path = exe
    return exe
