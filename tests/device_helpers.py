"""TEST INFRASTRUCTURE: the kernels' own arithmetic helpers, compiled for the HOST from the source text of the .cu files.

The bit-parallel edit distances (Myers32 / myers16 / myers64 / myers128 / myers_long), the search-mode lower-bound filter
(SearchMyers), the exact byte classifiers (is_acgt, encode4's ACGTU test, kmer_code) and passFilter's integer forms
(pass_filter) are plain integer C++ inside `__device__` functions.  This module cuts their definitions out of
fastplong_b200/csrc/*.cu / *.cuh by name, rewrites the two inline-PTX forms they use (`mad.lo.u32`) as C, and builds them
with g++ behind small shims (`__device__` -> nothing, `__ldg`, `__dp4a`, `uint4`) into /tmp — so that the CPU tests run the
very text the GPU runs, over input spaces the GPU tests can only sample (every byte value, ties of every threshold,
the one-sidedness of the filters).  Nothing here is shipped or used by the product path.
"""
import ctypes as C
import hashlib
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "fastplong_b200", "csrc")

SHIMS = r"""
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include "fplgpu.h"
using std::min; using std::max;
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
struct uint4 { uint32_t x, y, z, w; };
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c) {
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xFFu) * ((b >> (8 * i)) & 0xFFu);
    return c;
}
static inline int __dp4a(int a, int b, int c) {
    for (int i = 0; i < 4; i++) c += (int)(int8_t)((uint32_t)a >> (8 * i)) * (int)(int8_t)((uint32_t)b >> (8 * i));
    return c;
}
"""

HARNESS = r"""
// peq tables as fpl_create builds them (fpl_api.cu: bit j of word j >> 5 for the adapter's byte at j; 16-mer tables in
// one word; 64-bit words for the long form)
static void build_peq(const uint8_t* pat, int m, uint4* peq) {
    memset(peq, 0, sizeof(uint4) * 256);
    for (int j = 0; j < m && j < 128; j++) reinterpret_cast<uint32_t*>(&peq[pat[j]])[j >> 5] |= 1u << (j & 31);
}
extern "C" {
int h_myers32(const uint8_t* text, int n, const uint8_t* pat, int m, int shift, int sub) {
    uint4 peq[256]; build_peq(pat, m, peq);
    return myers32(text, n, peq, shift, sub);
}
int h_myers64(const uint8_t* text, int n, const uint8_t* pat, int m, int shift, int sub) {
    uint4 peq[256]; build_peq(pat, m, peq);
    return myers64(text, n, peq, shift, sub);
}
int h_myers128(const uint8_t* text, int n, const uint8_t* pat, int m, int shift, int sub) {
    uint4 peq[256]; build_peq(pat, m, peq);
    return myers128(text, n, peq, shift, sub);
}
int h_myers16(const uint8_t* text, int n, const uint8_t* pat, int m) {
    uint32_t t[256]; memset(t, 0, sizeof(t));
    for (int j = 0; j < m; j++) t[pat[j]] |= 1u << j;
    return myers16(text, n, t, m);
}
int h_myers_long(const uint8_t* text, int n, const uint8_t* pat, int m, int shift, int sub) {
    const int words = (m + 63) / 64;
    unsigned long long* peq = new unsigned long long[256 * (size_t)words]();
    for (int j = 0; j < m; j++) peq[(size_t)pat[j] * words + (j >> 6)] |= 1ull << (j & 63);
    const int d = myers_long(text, n, peq, words, shift, sub);
    delete[] peq;
    return d;
}
// search-mode pass over the text: sg[e] = SearchMyers' score after column e (pattern m <= 32 bits)
void h_search_scores(const uint8_t* text, int n, const uint8_t* pat, int m, int* sg) {
    uint32_t t[256]; memset(t, 0, sizeof(t));
    for (int j = 0; j < m; j++) t[pat[j]] |= 1u << j;
    SearchMyers<uint32_t> Q; Q.init(m);
    for (int e = 0; e < n; e++) { Q.column(t[text[e]]); sg[e] = Q.score; }
}
void h_search_scores64(const uint8_t* text, int n, const uint8_t* pat, int m, int* sg) {
    unsigned long long t[256]; memset(t, 0, sizeof(t));
    for (int j = 0; j < m; j++) t[pat[j]] |= 1ull << j;
    SearchMyers<unsigned long long> Q; Q.init(m);
    for (int e = 0; e < n; e++) { Q.column(t[text[e]]); sg[e] = Q.score; }
}
int h_pass_filter(const fpl_options* o, int rlen, int lowq, int nn, int totalq, int diff) {
    Counts c = {lowq, nn, totalq, diff};
    return pass_filter(*o, rlen, c);
}
int h_is_acgt(uint32_t b) { return is_acgt(b) ? 1 : 0; }
uint32_t h_kmer_code(uint32_t b) { return kmer_code(b); }
uint32_t h_zero_bytes80(uint32_t d) { return zero_bytes80(d); }
void h_encode4(uint32_t w, uint32_t* nz, uint32_t* pc) { encode4(w, 1u, *nz, *pc); }
}
"""

# (file, name, kind)
WANTED = [("fpl_device.cuh", "Myers32", "struct"), ("fpl_device.cuh", "myers_eq_top", "func"), ("fpl_device.cuh", "myers32", "func"),
          ("fpl_device.cuh", "myers_long", "func"), ("fpl_trim.cu", "peq_sub", "func"), ("fpl_trim.cu", "myers128", "func"),
          ("fpl_trim.cu", "myers64", "func"), ("fpl_trim.cu", "myers16", "func"), ("fpl_trim.cu", "is_acgt", "func"),
          ("fpl_trim.cu", "SearchMyers", "struct"), ("fpl_scan.cu", "Counts", "struct"), ("fpl_scan.cu", "pass_filter", "func"),
          ("fpl_scan.cu", "zero_bytes80", "func"), ("fpl_stats.cu", "kmer_code", "func"), ("fpl_stats.cu", "mad_u32", "func"),
          ("fpl_stats.cu", "encode4", "func")]


def _match_brace(text, i):
    depth = 0
    while True:
        ch = text[i]
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                return i
        i += 1


def extract(path, name, kind):
    """The definition of function / struct `name` in `path`, template header included."""
    text = open(path).read()
    if kind == "struct":
        m = re.search(r"^(template\s*<[^>\n]*>\s*\n)?struct\s+(?:__align__\(\d+\)\s+)?" + name + r"\s*\{", text, re.M)
    else:
        m = re.search(r"^(template\s*<[^>\n]*>\s*\n)?[A-Za-z_][^\n;{}()]*\b" + name + r"\s*\([^;{}]*\)\s*(?:const\s*)?\{", text, re.M)
    if not m:
        raise KeyError(f"{name} not found in {path}")
    end = _match_brace(text, m.end() - 1)
    body = text[m.start():end + 1]
    return body + (";" if kind == "struct" else "") + "\n"


def _asm_to_c(src):
    """mad.lo.u32 d, a, b, c  (operands %k or literals)  ->  d = a * b + c"""
    def repl(m):
        tmpl, outs, ins = m.group(1), m.group(2), m.group(3)
        ops = re.findall(r'"[^"]*"\s*\(([^()]*(?:\([^()]*\))?[^()]*)\)', outs) + re.findall(r'"r"\s*\(((?:[^()]|\([^()]*\))*)\)', ins)
        t = re.match(r"mad\.lo\.u32\s+(%\d+|\w+),\s*(%\d+|\w+),\s*(%\d+|\w+),\s*(%\d+|\w+);", tmpl)
        if not t:
            raise ValueError("inline PTX this harness does not know: " + tmpl)
        val = [("(uint32_t)(" + ops[int(x[1:])] + ")") if x.startswith("%") else x + "u" for x in t.groups()]
        return f"{ops[0]} = {val[1]} * {val[2]} + {val[3]};"
    return re.sub(r'asm\s*\(\s*"([^"]*)"\s*:\s*([^:;]*?)\s*:\s*(.*?)\);', repl, src)


def source():
    parts = [SHIMS]
    for fn, name, kind in WANTED:
        parts.append(f"// ---- {name} ({fn}) ----\n" + _asm_to_c(extract(os.path.join(CSRC, fn), name, kind)))
    parts.append(HARNESS)
    return "\n".join(parts)


# ---- the whole-read scan (k_scan_jit v2): its helpers live in a raw string that NVRTC compiles at fpl_create ----
JIT_WANTED = [("nz7", "func"), ("plane", "func"), ("nibble_to_bytes", "func"), ("range_mask", "func"), ("ge_mask", "func")]

JIT_HARNESS = r"""
extern "C" {
void j_planes(const uint32_t* w8, uint32_t* out5) {
    uint32_t w[8]; for (int i = 0; i < 8; i++) w[i] = w8[i];
    out5[0] = plane<0>(w); out5[1] = plane<1>(w); out5[2] = plane<2>(w); out5[3] = plane<3>(w); out5[4] = plane<4>(w);
}
// the byte classification of the plane path, the kernel's own lines: masks of A / C / G / T / N among the lane's 32 bytes and
// the test that sends a lane down that path at all
void j_classify(const uint32_t* w8, uint32_t* out5, int* plane_path) {
    uint32_t w[8]; for (int i = 0; i < 8; i++) w[i] = w8[i];
    uint32_t okacc = 0xFFFFFFFFu;
    for (int k = 0; k < 8; k++) okacc &= w[k] ^ 0xA0A0A0A0u;
    *plane_path = @@PLANE_TEST@@ ? 1 : 0;
    uint32_t MA, MC, MG, MT, NM;
    @@CLASSIFY@@
    out5[0] = MA; out5[1] = MC; out5[2] = MG; out5[3] = MT; out5[4] = NM;
}
uint32_t j_nibble_to_bytes(uint32_t n) { return nibble_to_bytes(n); }
uint32_t j_range_mask(int p_first, int n) { return range_mask(p_first, n); }
uint32_t j_nz7(uint32_t d) { return nz7(d); }
uint32_t j_ge_mask(const uint32_t* c, const uint32_t* row) {
    uint32_t cc[NPL]; for (int i = 0; i < NPL; i++) cc[i] = c[i];
    alignas(16) uint32_t r[8]; for (int i = 0; i < 8; i++) r[i] = row[i];
    return ge_mask(cc, r);
}
int j_npl() { return NPL; }
}
"""


def jit_source(amax):
    """Host translation unit with the helpers of k_scan_jit v2 for adapters of up to `amax` letters (NPL count planes)."""
    raw = open(os.path.join(CSRC, "fpl_scan_jit2_src.h")).read()
    text = raw[raw.index('R"JITSRC(') + len('R"JITSRC('):raw.rindex(')JITSRC"')]
    tmp = "/tmp/fpl_jit_text.cu"
    open(tmp, "w").write(text)
    parts = [SHIMS, _asm_to_c(extract(os.path.join(CSRC, "fpl_stats.cu"), "mad_u32", "func")),
             f"constexpr int NPL = {5 if amax <= 31 else 6 if amax <= 63 else 7 if amax <= 127 else 8};\n"]
    for name, kind in JIT_WANTED:
        parts.append(_asm_to_c(extract(tmp, name, kind)))
    m = re.search(r"if \(\((okacc \| 0x1F1F1F1Fu\) == 0xFFFFFFFFu)\) \{\s*\n(\s*const uint32_t B0 = plane<0>\(w\).*?NM = [^;]*;)", text, re.S)
    if not m:
        raise KeyError("the plane-path classification block of k_scan_jit was not found")
    parts.append(JIT_HARNESS.replace("@@PLANE_TEST@@", "((" + m.group(1) + ")").replace("@@CLASSIFY@@", m.group(2)))
    return "\n".join(parts)


_jit_libs = {}


def load_jit(amax=30):
    if amax in _jit_libs:
        return _jit_libs[amax]
    src = jit_source(amax)
    so = f"/tmp/fpl_jit_helpers_{hashlib.md5(src.encode()).hexdigest()[:12]}.so"
    if not os.path.exists(so):
        cpp = so[:-3] + ".cpp"
        open(cpp, "w").write(src)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-w", "-I", os.path.join(ROOT, "include"), "-o", so, cpp])
    lib = C.CDLL(so)
    lib.j_planes.argtypes = [C.c_void_p, C.c_void_p]
    lib.j_classify.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.j_nibble_to_bytes.argtypes = lib.j_nz7.argtypes = [C.c_uint32]
    lib.j_nibble_to_bytes.restype = lib.j_nz7.restype = lib.j_range_mask.restype = lib.j_ge_mask.restype = C.c_uint32
    lib.j_range_mask.argtypes = [C.c_int, C.c_int]
    lib.j_ge_mask.argtypes = [C.c_void_p, C.c_void_p]
    _jit_libs[amax] = lib
    return lib


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = source()
    tag = hashlib.md5(src.encode()).hexdigest()[:12]
    so = f"/tmp/fpl_device_helpers_{tag}.so"
    if not os.path.exists(so):
        cpp = so[:-3] + ".cpp"
        open(cpp, "w").write(src)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-w", "-I", os.path.join(ROOT, "include"), "-o", so, cpp])
    lib = C.CDLL(so)
    for f in ("h_myers32", "h_myers64", "h_myers128", "h_myers_long"):
        getattr(lib, f).argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.h_myers16.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    lib.h_search_scores.argtypes = lib.h_search_scores64.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p]
    lib.h_pass_filter.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.h_is_acgt.argtypes = lib.h_kmer_code.argtypes = lib.h_zero_bytes80.argtypes = [C.c_uint32]
    lib.h_kmer_code.restype = lib.h_zero_bytes80.restype = C.c_uint32
    lib.h_encode4.argtypes = [C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    _lib = lib
    return lib
