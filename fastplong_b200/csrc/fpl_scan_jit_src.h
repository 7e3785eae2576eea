// Source of k_scan_jit, compiled at fpl_create() time with NVRTC for sm_100a and specialised on the two adapter
// strings (FPL_A0 / FPL_A1), their lengths and the option flags.  Same contract and same arithmetic as k_scan_fast
// (fpl_scan_fast.cu) — the specialisation only removes work that does not depend on the data:
//   * the letter of adapter position i is a compile-time constant, so "the match vector of letter a_i shifted by i"
//     is ONE funnel shift of two registers (the letter's mask of this lane and of its neighbour); k_scan_fast needs
//     three shifts and two LOP3 for the same bit vector because it must select the letter at run time;
//   * the carry-save adder tree is generated for the exact adapter length (no padding, no run-time block guards);
//   * option flags (filters on/off, qualified quality) are constants.
// The text is a raw string so that it stays readable here; fpl_jit.cpp prepends the #defines.
#pragma once

static const char* const kScanJitSource = R"JITSRC(
typedef unsigned char uint8_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;

struct ReadState {          // must match fpl_device.cuh
    int32_t lo, len;
    uint32_t alive, pad;
    unsigned long long best[2];
    int32_t lowq, nn, totalq, diff;
    int32_t reserved[4];
};

#define SF_WARPS 4
#define SF_THREADS (SF_WARPS * 32)

constexpr char A0[] = FPL_A0;
constexpr char A1[] = FPL_A1;
constexpr int ALEN0 = sizeof(A0) - 1;
constexpr int ALEN1 = sizeof(A1) - 1;
constexpr int AMAX = ALEN0 > ALEN1 ? ALEN0 : ALEN1;
constexpr int HL = ((AMAX - 1) >> 5) + 1;                         // halo words
constexpr int NPL = AMAX <= 31 ? 5 : AMAX <= 63 ? 6 : AMAX <= 127 ? 7 : 8;

#define CSA(h, l, a, b, c)                            \
    do {                                              \
        const uint32_t a_ = (a), b_ = (b), c_ = (c);  \
        l = a_ ^ b_ ^ c_;                             \
        h = (a_ & b_) | (c_ & (a_ | b_));             \
    } while (0)

__device__ __forceinline__ uint32_t plane_nibble(uint32_t w, uint32_t mask, uint32_t mul) { return (w & mask) * mul; }
__device__ __forceinline__ uint32_t nz7(uint32_t d) { return (((d & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d) & 0x80808080u; }

struct Masks { uint32_t A[HL + 1], C[HL + 1], G[HL + 1], T[HL + 1]; };

template <char L>
__device__ __forceinline__ uint32_t pick(const Masks& m, int w) {
    if constexpr (L == 'A') return m.A[w];
    else if constexpr (L == 'C') return m.C[w];
    else if constexpr (L == 'G') return m.G[w];
    else return m.T[w];
}

// match vector of adapter K's letter I for the lane's 32 positions (0 beyond the adapter's end)
template <int K, int I>
__device__ __forceinline__ uint32_t letter(const Masks& m) {
    constexpr int ALEN = K ? ALEN1 : ALEN0;
    if constexpr (I >= ALEN) return 0u;
    else {
        constexpr char L = K ? A1[I] : A0[I];
        constexpr int w = I >> 5, sh = I & 31;
        if constexpr (sh == 0) return pick<L>(m, w);
        else return __funnelshift_r(pick<L>(m, w), pick<L>(m, w + 1), sh);
    }
}

struct Counter {
    uint32_t p[NPL];   // bit-sliced match count
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int b = 0; b < NPL; b++) p[b] = 0;
    }
    __device__ __forceinline__ void ripple(uint32_t carry, int from) {
#pragma unroll
        for (int b = 0; b < NPL; b++)
            if (b >= from) { const uint32_t t = p[b] & carry; p[b] ^= carry; carry = t; }
    }
    __device__ __forceinline__ void add8(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, uint32_t x5,
                                         uint32_t x6, uint32_t x7) {
        uint32_t tA, tB, fA, fB, e;
        CSA(tA, p[0], p[0], x0, x1);
        CSA(tB, p[0], p[0], x2, x3);
        CSA(fA, p[1], p[1], tA, tB);
        CSA(tA, p[0], p[0], x4, x5);
        CSA(tB, p[0], p[0], x6, x7);
        CSA(fB, p[1], p[1], tA, tB);
        CSA(e, p[2], p[2], fA, fB);
        ripple(e, 3);
    }
};

// all blocks of 8 letters of adapter K, generated for its exact length (trailing zero inputs fold away)
template <int K, int I0>
__device__ __forceinline__ void add_blocks(Counter& c, const Masks& m) {
    constexpr int ALEN = K ? ALEN1 : ALEN0;
    if constexpr (I0 < ALEN) {
        c.add8(letter<K, I0>(m), letter<K, I0 + 1>(m), letter<K, I0 + 2>(m), letter<K, I0 + 3>(m),
               letter<K, I0 + 4>(m), letter<K, I0 + 5>(m), letter<K, I0 + 6>(m), letter<K, I0 + 7>(m));
        add_blocks<K, I0 + 8>(c, m);
    }
}

// gbest = best match count any lane of the CTA has seen for this adapter (shared memory): a position can only be the
// read's first arg-max if its count is >= gbest, which the top planes rule out for almost every lane.
__device__ __forceinline__ void argmax(const Counter& c, uint32_t valid, int64_t pos0, int& bestM, int64_t& bestPos,
                                       volatile int* gbest) {
    if (!valid) return;
    const int g = *gbest;
    if constexpr (NPL >= 2) {
        const uint32_t top = c.p[NPL - 1], nxt = c.p[NPL - 2];
        constexpr int T1 = 1 << (NPL - 1), T0 = 1 << (NPL - 2);
        const uint32_t f = g >= T1 + T0 ? (top & nxt) : g >= T1 ? top : g >= T0 ? (top | nxt) : 0xFFFFFFFFu;
        if (!(valid & f)) return;
    }
    uint32_t cand = valid;
    int val = 0;
#pragma unroll
    for (int b = NPL - 1; b >= 0; b--) {
        const uint32_t t = cand & c.p[b];
        if (t) { cand = t; val |= 1 << b; }
    }
    if (val > bestM) {
        bestM = val; bestPos = pos0 + (__ffs(cand) - 1);
        if (val > g) atomicMax((int*)gbest, val);
    }
}

extern "C" __global__ void __launch_bounds__(SF_THREADS, FPL_MINBLOCKS)
k_scan_jit(const uint8_t* __restrict__ seqbuf, const uint8_t* __restrict__ qualbuf, const int64_t* __restrict__ offsets,
           ReadState* __restrict__ st, int64_t n_reads) {
    __shared__ unsigned long long sh64[2][SF_WARPS];
    __shared__ int sh32[4][SF_WARPS];
    __shared__ int gbest[2];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = blockIdx.x;
    if (r >= n_reads) return;
    const ReadState s = st[r];
    if (!s.alive) return;
    if (threadIdx.x < 2) gbest[threadIdx.x] = -1;
    __syncthreads();
    const int len = s.len;
    const int64_t start = offsets[r] + s.lo;
    const int pre = (int)(start & 15);
    const uint8_t* sbase = seqbuf + (start - pre);
    const uint8_t* qbase = qualbuf + (start - pre);
    constexpr bool doAdapters = FPL_DO_ADAPTERS;
    constexpr bool doCounts = FPL_DO_COUNTS;
    constexpr bool doCplx = FPL_DO_CPLX;
    constexpr uint32_t qq4 = (uint32_t)(FPL_QQ & 0x7f) * 0x01010101u;
    const int np0 = (doAdapters && ALEN0 <= len) ? len - ALEN0 : 0;
    const int np1 = (doAdapters && ALEN1 <= len) ? len - ALEN1 : 0;
    constexpr int step = (32 - HL) * 32;
    const int total = pre + len;
    const int np_min = min(np0, np1);                          // < len; 0 when the scans are off

    int bestM0 = -1, bestM1 = -1;
    int64_t bestP0 = 0, bestP1 = 0;
    int nn = 0, totalq = 0, diff = 0, nbytes = 0;
    unsigned ge128 = 0;   // 128 * #(q >= qualified)

    // software pipeline: the sequence vectors of the warp's next tile are requested before this tile is processed
    uint4 nx = make_uint4(0, 0, 0, 0), ny = nx;
    {
        const int64_t a0 = (int64_t)wid * step + 32 * lane;
        if (a0 < total && a0 + 32 > pre) {
            const uint4* v = reinterpret_cast<const uint4*>(sbase + a0);
            nx = __ldg(v); ny = __ldg(v + 1);
        }
    }
    for (int64_t t0 = (int64_t)wid * step; t0 < total; t0 += (int64_t)SF_WARPS * step) {
        const int64_t a0 = t0 + 32 * lane;
        const bool inrange = a0 < total && a0 + 32 > pre;
        const uint32_t w[8] = {nx.x, nx.y, nx.z, nx.w, ny.x, ny.y, ny.z, ny.w};
        {
            const int64_t a1 = a0 + (int64_t)SF_WARPS * step;
            nx = make_uint4(0, 0, 0, 0); ny = nx;
            if (a1 < total && a1 + 32 > pre) {
                const uint4* v = reinterpret_cast<const uint4*>(sbase + a1);
                nx = __ldg(v); ny = __ldg(v + 1);
            }
        }
        // every byte in 0x40..0x5F  <=>  bits 7..5 of (byte ^ 0xA0) are all ones: one LOP3 per word
        uint32_t okacc = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < 8; k++) okacc &= w[k] ^ 0xA0A0A0A0u;
        uint32_t MA, MC, MG, MT, NM;
        if ((okacc | 0x1F1F1F1Fu) == 0xFFFFFFFFu) {
            uint32_t B0 = 0, B1 = 0, B2 = 0, B3 = 0, B4 = 0;
#pragma unroll
            for (int k = 7; k >= 0; k--) {
                B0 = __funnelshift_l(plane_nibble(w[k], 0x01010101u, 0x10204080u), B0, 4);
                B1 = __funnelshift_l(plane_nibble(w[k], 0x02020202u, 0x08102040u), B1, 4);
                B2 = __funnelshift_l(plane_nibble(w[k], 0x04040404u, 0x04081020u), B2, 4);
                B3 = __funnelshift_l(plane_nibble(w[k], 0x08080808u, 0x02040810u), B3, 4);
                B4 = __funnelshift_l(plane_nibble(w[k], 0x10101010u, 0x01020408u), B4, 4);
            }
            // bytes are 010 b4 b3 b2 b1 b0:  A 00001  C 00011  G 00111  T 10100  N 01110
            const uint32_t acg = ~B3 & ~B4 & B0;
            MA = acg & ~B1 & ~B2;
            MC = acg & B1 & ~B2;
            MG = acg & B1 & B2;
            MT = ~B3 & B4 & ~B0 & ~B1 & B2;
            NM = B3 & B2 & B1 & ~B0 & ~B4;
        } else {
            MA = MC = MG = MT = NM = 0;
            for (int j = 0; j < 32; j++) {
                const uint32_t ch = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                const uint32_t bit = 1u << j;
                MA |= ch == 'A' ? bit : 0u; MC |= ch == 'C' ? bit : 0u;
                MG |= ch == 'G' ? bit : 0u; MT |= ch == 'T' ? bit : 0u;
                NM |= ch == 'N' ? bit : 0u;
            }
        }
        const int64_t p_first = a0 - pre;
        const bool mine = lane < 32 - HL;
        uint32_t inwin = 0, v0 = 0, v1 = 0;
        if (mine && inrange) {
            if (p_first >= 0 && p_first + 32 <= np_min) {       // interior lane: everything is in range
                inwin = v0 = v1 = 0xFFFFFFFFu;
            } else {
                const uint32_t from0 = p_first >= 0 ? 0xFFFFFFFFu : (0xFFFFFFFFu << (int)(-p_first));
                auto range_mask = [&](int n) -> uint32_t {
                    const int64_t hi = (int64_t)n - p_first;
                    const uint32_t upto = hi >= 32 ? 0xFFFFFFFFu : hi <= 0 ? 0u : ((1u << (int)hi) - 1u);
                    return upto & from0;
                };
                inwin = range_mask(len);
                v0 = range_mask(np0);
                v1 = range_mask(np1);
            }
        }
        if constexpr (doCounts) {
            if (inwin) {
                const uint4* v = reinterpret_cast<const uint4*>(qbase + a0);
                const uint4 x = __ldg(v), y = __ldg(v + 1);
                const uint32_t q[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
                if (inwin == 0xFFFFFFFFu) {
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        totalq = (int)__dp4a(q[k], 0x01010101u, (unsigned)totalq);
                        // bytes are 0x80 where q >= qualified: their dp4a sum is 128 * count (divided out at the end)
                        ge128 = __dp4a(((q[k] | 0x80808080u) - qq4) & 0x80808080u, 0x01010101u, ge128);
                    }
                    nbytes += 32;
                } else {
                    for (int j = 0; j < 32; j++)
                        if (inwin >> j & 1u) {
                            const int qv = (int)((q[j >> 2] >> (8 * (j & 3))) & 0xFFu);
                            totalq += qv;
                            ge128 += ((qv & 0x7f) >= (int)(qq4 & 0x7f)) ? 128u : 0u;
                            nbytes++;
                        }
                }
                nn += __popc(NM & inwin);
            }
        }
        if constexpr (doCplx) {
            const uint32_t nxt = __shfl_down_sync(0xffffffffu, w[0], 1);
            if (inwin) {
                uint32_t pm = inwin;
                const int64_t last = (int64_t)len - 1 - p_first;
                if (last >= 0 && last < 32) pm &= ~(1u << last);
                if (pm == 0xFFFFFFFFu) {
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const uint32_t w2 = k < 7 ? w[k + 1] : nxt;
                        diff += __popc(nz7(w[k] ^ __funnelshift_r(w[k], w2, 8)));
                    }
                } else {
                    for (int j = 0; j < 32; j++)
                        if (pm >> j & 1u) {
                            const uint32_t c0 = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                            const uint32_t w2 = (j + 1) < 32 ? w[(j + 1) >> 2] : nxt;
                            const uint32_t c1 = (w2 >> (8 * ((j + 1) & 3))) & 0xFFu;
                            diff += c0 != c1;
                        }
                }
            }
        }
        if constexpr (doAdapters) {
            Masks m;
            m.A[0] = MA; m.C[0] = MC; m.G[0] = MG; m.T[0] = MT;
#pragma unroll
            for (int k = 1; k <= HL; k++) {
                m.A[k] = __shfl_down_sync(0xffffffffu, MA, k); m.C[k] = __shfl_down_sync(0xffffffffu, MC, k);
                m.G[k] = __shfl_down_sync(0xffffffffu, MG, k); m.T[k] = __shfl_down_sync(0xffffffffu, MT, k);
            }
            Counter c0, c1;
            c0.clear();
            add_blocks<0, 0>(c0, m);
            argmax(c0, v0, p_first, bestM0, bestP0, &gbest[0]);
            c1.clear();
            add_blocks<1, 0>(c1, m);
            argmax(c1, v1, p_first, bestM1, bestP1, &gbest[1]);
        }
    }
    unsigned long long k0 = bestM0 >= 0 ? (((unsigned long long)(unsigned)(ALEN0 - bestM0) << 32) | (unsigned)bestP0) : ~0ull;
    unsigned long long k1 = bestM1 >= 0 ? (((unsigned long long)(unsigned)(ALEN1 - bestM1) << 32) | (unsigned)bestP1) : ~0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, k0, d), o1 = __shfl_xor_sync(0xffffffffu, k1, d);
        k0 = o0 < k0 ? o0 : k0; k1 = o1 < k1 ? o1 : k1;
    }
    int lowq_ge = (int)(ge128 >> 7);
    lowq_ge = __reduce_add_sync(0xffffffffu, lowq_ge); nn = __reduce_add_sync(0xffffffffu, nn);
    totalq = __reduce_add_sync(0xffffffffu, totalq); diff = __reduce_add_sync(0xffffffffu, diff);
    nbytes = __reduce_add_sync(0xffffffffu, nbytes);
    if (lane == 0) {
        sh64[0][wid] = k0; sh64[1][wid] = k1;
        sh32[0][wid] = nbytes - lowq_ge;
        sh32[1][wid] = nn;
        sh32[2][wid] = totalq - 33 * nbytes;
        sh32[3][wid] = diff;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m0 = sh64[0][0], m1 = sh64[1][0];
        int c0 = sh32[0][0], c1 = sh32[1][0], c2 = sh32[2][0], c3 = sh32[3][0];
        for (int i = 1; i < SF_WARPS; i++) {
            m0 = sh64[0][i] < m0 ? sh64[0][i] : m0; m1 = sh64[1][i] < m1 ? sh64[1][i] : m1;
            c0 += sh32[0][i]; c1 += sh32[1][i]; c2 += sh32[2][i]; c3 += sh32[3][i];
        }
        ReadState* o = &st[r];
        o->best[0] = m0; o->best[1] = m1;
        o->lowq = c0; o->nn = c1; o->totalq = c2; o->diff = c3;
    }
}
)JITSRC";
