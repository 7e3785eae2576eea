"""TEST INFRASTRUCTURE: the record path of libfplgpu executed on the CPU.

The source text of fpl_device.cuh, fpl_trim.cu (k_trim, k_trim_fasta) and fpl_scan.cu (the generic k_scan, k_final, k_count)
is preprocessed — CUDA includes dropped, `mad.lo.u32` inline PTX rewritten as C, `kernel<<<grid, block, smem, stream>>>(args)`
turned into a call of the SIMT emulator (tests/simt/emu_cuda.h: one fiber per CUDA thread, warp / block collectives with their
real semantics) — and compiled with g++ together with the table builder cut out of fpl_create (fpl_api.cu), so that adapters,
thresholds and match masks are prepared by the product's own code.  `process(options, batch)` then runs
launch_trim -> launch_scan -> launch_final -> launch_count in run_batch's order and returns the per-read records and the
counter vector, to be compared with the oracle like a GPU result.

What this covers: everything the records and counters depend on except the NVRTC-specialised scan (the generic k_scan
computes the same ReadState by contract; the GPU tests hold the variants to each other).  Not covered: the Stats kernels
(cp.async rings, shared-memory reductions in PTX, cub sorts), FASTQ ingest / emit, --mask/--break, timing, races.
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
    m = re.search(r"^[A-Za-z_][^\n;{}()]*\b" + name + r"\s*\([^;{}]*\)\s*\{", text, re.M)
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
        out += text[i:m.start()] + f"EMU_LAUNCH(({cfg[0]}), ({cfg[1]}), {m.group(1)}{text[p:q + 1]})"
        i = q + 1


def device_text(fn, drop=()):
    text = open(os.path.join(CSRC, fn)).read()
    text = re.sub(r'^\s*#include\s+[<"](cuda_runtime\.h|fpl_device\.cuh|cuda\.h)[>"].*$', "", text, flags=re.M)
    text = text.replace("#pragma once", "")
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
#include <string>
#include "emu_cuda_impl.h"
#include "fplgpu.h"
#include "fpl_scanplan.h"
@@DEVICE@@

// a host-memory stand-in for the few runtime calls inside the table builder
enum cudaError_t { cudaSuccess = 0 };
enum { cudaMemcpyHostToDevice = 1, cudaStreamNonBlocking = 1 };
static const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
template <class T> static cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)calloc(n ? n : 1, 1); return cudaSuccess; }
static cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memcpy(d, s, n); return cudaSuccess; }
static cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return cudaSuccess; }

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

// processSingleEnd's record path in run_batch's order (fpl_api.cu), on host memory
extern "C" int emu_process(const fpl_options* opt, const fpl_adapters* ad, const fpl_batch* hb, fpl_read_result* results,
                           unsigned long long* counters, int64_t n_counter_words) {
    g_err[0] = 0;
    const int n = 2 + (ad->n_fasta > 0 ? ad->n_fasta : 0);
    EmuCtx* c = new EmuCtx();
    c->n_adapters = n;
    auto stamp = [](const char*) {};
@@BUILDER@@
    if (n_counter_words != c->counter_words) { fpl_destroy(c); return fail("counter words %lld != %lld", (long long)n_counter_words, (long long)c->counter_words); }
    const int64_t nr = hb->n_reads;
    DevBatch b = {hb->seq, hb->qual, hb->offsets, hb->lens, nr};
    std::vector<ReadState> st((size_t)nr + 1);
    std::vector<StatSeg> post(2 * (size_t)nr + 2);
    memset(results, 0xAB, sizeof(fpl_read_result) * (size_t)nr);          // k_trim must write every record
    launch_trim(c->P, b, st.data(), results, c->d_counters, nullptr);
    launch_scan(c->P, b, st.data(), nullptr);
    launch_final(c->P, b, st.data(), results, post.data(), nullptr);
    launch_count(results, nr, c->d_counters, true, nullptr);
    memcpy(counters, c->d_counters, sizeof(unsigned long long) * (size_t)c->counter_words);
    fpl_destroy(c);
    return 0;
}
"""


def source():
    dev = "\n".join([device_text("fpl_device.cuh", drop=("red_shared_add", "shared_addr")), device_text("fpl_trim.cu"),
                     device_text("fpl_scan.cu")])
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
    lib.emu_last_error.restype = C.c_char_p
    lib.emu_collectives.restype = C.c_longlong
    lib.emu_process.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(FplBatch), C.c_void_p, C.c_void_p, C.c_int64]
    _lib = lib
    return lib


class EmuEngine:
    """binding.Engine's process() / counters() for the emulated record path."""

    def __init__(self, options):
        self.lib = load()
        self.options = options
        self._abi = options.to_abi()
        self.n_adapters = 2 + len(options.adapter_fasta)
        self._counters = np.zeros(abi.counter_words(self.n_adapters), dtype=np.int64)

    def process(self, batch):
        o, ad, keep = self._abi
        res = np.zeros(batch.n_reads, dtype=RESULT_DTYPE)
        cnt = np.zeros_like(self._counters)
        b = batch.to_abi()
        rc = self.lib.emu_process(C.byref(o), C.byref(ad), C.byref(b), res.ctypes.data, cnt.ctypes.data, cnt.shape[0])
        if rc != 0:
            raise RuntimeError(self.lib.emu_last_error().decode())
        self._counters += cnt
        return res

    def counters(self):
        return self._counters.copy()
