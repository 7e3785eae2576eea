// Stats::statRead (src/stats.cpp:265-375) as two kernels over a list of segments
// (pre-filter Stats: every input read; post-filter Stats: every passing segment):
//
//  k_cycle_stats  column-tiled: a CTA owns a tile of 512 cycles x a group of up to 4000 segments; its warps take the
//                 segments that reach the tile round-robin, one 16-byte vector of sequence and of quality per
//                 lane.  The per-(base-bin, cycle) counters live in shared memory (column j*32+lane: conflict-free)
//                 and are updated with shared-memory atomics; one 32-bit word packs (count << 20 | sum of raw
//                 quality chars).  The tile is flushed once to the global int64 arrays mCycleBaseContents /
//                 mCycleBaseQual.  Also counts the 5-mers (mKmer) with a SWAR fast path for all-ACGTU vectors.
//  k_read_qual    row-shaped: a warp owns a read, builds its quality histogram with shared-memory atomics
//                 (16-byte vector loads), adds it to mBaseQualHistogram and derives the per-read median quality
//                 (mMedianReadQualHistogram / mMedianReadQualBases / mReads / mLengthSum) for BOTH Stats objects:
//                 a passing segment's histogram is the read's minus the removed ends.
//  k_kmer_fix     post-filter 5-mer table = pre-filter table - the 5-mers outside the passing segments.
#include "fpl_device.cuh"

#define CS_THREADS 256
#define CS_WARPS (CS_THREADS / 32)
#define CS_TILE 512                              // cycles per CTA: one 16-byte vector per lane
#define CS_GROUP 4000                            // segments per CTA; packed counter: count <= 4095, sum of q < 2^20
#define CS_STAGE 1000                            // descriptors staged in shared memory at a time

namespace {

// byte -> (valid << 2 | 2-bit code) for the 5-mer table: A=0, T/U=1, C=2, G=3 (Stats::base2val, src/stats.cpp:411-425)
__device__ __forceinline__ uint32_t kmer_code(uint32_t b) {
    uint32_t v = 8;  // invalid
    v = b == 'A' ? 0u : v;
    v = (b == 'T' || b == 'U') ? 1u : v;
    v = b == 'C' ? 2u : v;
    v = b == 'G' ? 3u : v;
    return v;
}

// 16 bytes at an arbitrary address: two aligned 16-byte loads (load16_raw, issued early) + a byte shift
// (assemble16, at the point of use so that the loads stay in flight; sh = address & 15 is warp-uniform)
struct Raw16 { uint4 x, y; };
__device__ __forceinline__ Raw16 load16_raw(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint4* v = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    Raw16 r;
    r.x = __ldg(v);
    r.y = (a & 15) ? __ldg(v + 1) : make_uint4(0, 0, 0, 0);
    return r;
}
__device__ __forceinline__ void assemble16(const Raw16& r, unsigned sh, uint32_t (&o)[4]) {
    const uint4 x = r.x, y = r.y;
    if (sh == 0) { o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; return; }
    const unsigned bs = (sh & 3) * 8;
    switch (sh >> 2) {
        case 0:
            o[0] = __funnelshift_r(x.x, x.y, bs); o[1] = __funnelshift_r(x.y, x.z, bs);
            o[2] = __funnelshift_r(x.z, x.w, bs); o[3] = __funnelshift_r(x.w, y.x, bs); break;
        case 1:
            o[0] = __funnelshift_r(x.y, x.z, bs); o[1] = __funnelshift_r(x.z, x.w, bs);
            o[2] = __funnelshift_r(x.w, y.x, bs); o[3] = __funnelshift_r(y.x, y.y, bs); break;
        case 2:
            o[0] = __funnelshift_r(x.z, x.w, bs); o[1] = __funnelshift_r(x.w, y.x, bs);
            o[2] = __funnelshift_r(y.x, y.y, bs); o[3] = __funnelshift_r(y.y, y.z, bs); break;
        default:
            o[0] = __funnelshift_r(x.w, y.x, bs); o[1] = __funnelshift_r(y.x, y.y, bs);
            o[2] = __funnelshift_r(y.y, y.z, bs); o[3] = __funnelshift_r(y.z, y.w, bs); break;
    }
}

__device__ __forceinline__ uint32_t load4_before(const uint8_t* p) {   // the 4 bytes ending just before p
    const uintptr_t a = reinterpret_cast<uintptr_t>(p) - 4;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3) * 8;
    const uint32_t lo = __ldg(w);
    if (sh == 0) return lo;
    return __funnelshift_r(lo, __ldg(w + 1), sh);
}

// 0x01 in every byte of w that is one of A, C, G, T, U (exact): b & 0xE8 == 0x40 and, on bits (b4,b2,b1,b0),
// b4 ? (b2 & ~b1) : (b0 & (b1 | ~b2))
__device__ __forceinline__ uint32_t valid_acgtu(uint32_t w) {
    const uint32_t x1 = w >> 1, x2 = w >> 2, x4 = w >> 4;
    const uint32_t g1 = w & (x1 | ~x2);
    const uint32_t g2 = x2 & ~x1;
    const uint32_t v = (x4 & g2) | (~x4 & g1);
    const uint32_t t = (w & 0xE8E8E8E8u) ^ 0x40404040u;                         // zero byte <=> 010x0xxx
    const uint32_t nz = (((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) >> 7;            // bit0 of each byte: t != 0
    return v & ~nz & 0x01010101u;
}

// four 2-bit 5-mer codes of a word, oldest base in the top bits: code = bit1 << 1 | bit2 (A=0, T/U=1, C=2, G=3)
__device__ __forceinline__ uint32_t pack_codes(uint32_t w) {
    const uint32_t c = (w & 0x02020202u) | ((w >> 2) & 0x01010101u);
    return (c * 0x40100401u) >> 24;
}

}  // namespace

// DO_KMER: also count the 5-mers; their table is flushed to `stats` and, if given, to `kmer_also` (the post-filter
// block: post 5-mers = pre 5-mers - the ones k_kmer_fix finds outside the passing segments).
template <bool DO_KMER>
__global__ void __launch_bounds__(CS_THREADS, 4)
k_cycle_stats(const uint8_t* __restrict__ seqbuf, const uint8_t* __restrict__ qualbuf, const StatSeg* __restrict__ segs,
              int64_t nseg, unsigned long long* __restrict__ stats, int64_t C, unsigned long long* __restrict__ kmer_also) {
    // packed[bin][j*32 + lane]: cycle c0 + 16*lane + j; one word = count << 20 | sum of quality chars
    __shared__ uint32_t packed[8][CS_TILE];
    __shared__ uint32_t kmer[1024 + 32];   // [1024..1055]: sink for the lanes whose 5-mer is not valid
    __shared__ int64_t d_off[CS_STAGE];
    __shared__ int32_t d_len[CS_STAGE];
    __shared__ int d_n;
    const int wid = threadIdx.x >> 5, lane = lane_id();
    const int64_t c0 = (int64_t)blockIdx.x * CS_TILE;      // first cycle of this CTA
    const int64_t g0 = (int64_t)blockIdx.y * CS_GROUP;
    const int64_t g1 = min(nseg, g0 + CS_GROUP);
    for (int i = threadIdx.x; i < 1024; i += CS_THREADS) kmer[i] = 0;
    for (int i = threadIdx.x; i < 8 * CS_TILE; i += CS_THREADS) (&packed[0][0])[i] = 0;
    const int64_t cl = c0 + 16 * lane;                     // this lane's first cycle
    const uint32_t pk_lane = shared_addr(&packed[0][0]) + (uint32_t)lane * 4u;
    const uint32_t km_base = shared_addr(&kmer[0]);
    const uint32_t km_sink = km_base + (1024u + (uint32_t)lane) * 4u;
    bool any = false;
    for (int64_t s0 = g0; s0 < g1; s0 += CS_STAGE) {
        __syncthreads();
        if (threadIdx.x == 0) d_n = 0;
        __syncthreads();
        // stage the descriptors of the segments that reach this tile
        for (int64_t s = s0 + threadIdx.x; s < min(g1, s0 + CS_STAGE); s += CS_THREADS) {
            const StatSeg sg = segs[s];
            if ((int64_t)sg.len > c0) {
                const int k = atomicAdd(&d_n, 1);
                d_off[k] = sg.off; d_len[k] = sg.len;
            }
        }
        __syncthreads();
        const int n = d_n;
        if (n) any = true;
        // software pipeline: the vectors of segment k+CS_WARPS are in flight while segment k is processed
        Raw16 nrs, nrq;
        nrs.x = nrs.y = nrq.x = nrq.y = make_uint4(0, 0, 0, 0);
        uint32_t nprev0 = 0;
        unsigned nsh = 0;
        int nlen = 0;
        auto fetch = [&](int k) {
            const int64_t off = d_off[k];
            nlen = d_len[k];
            const uint8_t* sp = seqbuf + off + c0;
            nsh = (unsigned)(reinterpret_cast<uintptr_t>(sp) & 15);   // same for the quality buffer (both 16-byte aligned bases)
            if (cl < nlen) {
                nrs = load16_raw(sp + 16 * lane);
                nrq = load16_raw(qualbuf + off + c0 + 16 * lane);
            }
            if (DO_KMER && lane == 0) nprev0 = c0 >= 4 ? load4_before(sp) : 0u;
        };
        if (wid < n) fetch(wid);
        for (int k = wid; k < n; k += CS_WARPS) {
            const Raw16 rs = nrs, rq = nrq;
            const unsigned sh = nsh;
            const int len = nlen;
            const uint32_t prev0 = nprev0;
            if (k + CS_WARPS < n) fetch(k + CS_WARPS);
            const bool active = cl < len;
            uint32_t sw[4] = {0, 0, 0, 0}, qw[4] = {0, 0, 0, 0};
            if (active) { assemble16(rs, sh, sw); assemble16(rq, sh, qw); }
            // per word: validity nibble (A,C,G,T,U) and four 2-bit codes; the previous lane's last word supplies the
            // four bases in front of this lane's vector (lane 0: the word loaded in front of the tile)
            uint32_t vn[4] = {0, 0, 0, 0}, pc[4] = {0, 0, 0, 0}, pvn = 0, ppc = 0;
            if (DO_KMER) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    vn[i] = active ? (valid_acgtu(sw[i]) * 0x10204080u) >> 28 : 0u;
                    pc[i] = pack_codes(sw[i]);
                }
                pvn = __shfl_up_sync(0xffffffffu, vn[3], 1); ppc = __shfl_up_sync(0xffffffffu, pc[3], 1);
                if (lane == 0) { pvn = (valid_acgtu(prev0) * 0x10204080u) >> 28; ppc = pack_codes(prev0); }
            }
            if (!active) continue;
            const int nvalid = (int)min((int64_t)16, (int64_t)len - cl);
            // ---- per-(bin, cycle) counters: word = count << 20 | sum of quality chars ----
            // shared byte address of packed[base & 7][t*32 + lane] = pk_lane + ((base & 7) << 11) + (t << 7)
            auto count = [&](int t) {
                const uint32_t w = sw[t >> 2];
                const int j = t & 3;
                const uint32_t binoff = j == 0 ? (w << 11) : j == 1 ? (w << 3) : j == 2 ? (w >> 5) : (w >> 13);
                const uint32_t val = __byte_perm(qw[t >> 2], 0x00100000u, 0x7650 + j);   // q | 1 << 20
                red_shared_add(pk_lane + (binoff & 0x3800u) + ((uint32_t)t << 7), val);
            };
            if (nvalid == 16) {
#pragma unroll
                for (int t = 0; t < 16; t++) count(t);
            } else {
#pragma unroll
                for (int t = 0; t < 16; t++)
                    if (t < nvalid) count(t);
            }
            if (!DO_KMER) continue;
            // ---- 5-mers ending in this lane's 16 cycles: all five bases in ACGTU ----
            const uint32_t V20 = pvn | (vn[0] << 4) | (vn[1] << 8) | (vn[2] << 12) | (vn[3] << 16);
            uint32_t ok = V20 & (V20 >> 1) & (V20 >> 2) & (V20 >> 3) & (V20 >> 4);   // bit t: bytes t-4..t all valid
            if (nvalid < 16) ok &= (1u << nvalid) - 1u;
            if (ok) {
                // 20 codes, oldest first, 2 bits each: Phi = codes of bytes -4..11 (32 bits), Plo = bytes 4..15 low part
                const uint32_t Phi = (ppc << 24) | (pc[0] << 16) | (pc[1] << 8) | pc[2];   // bytes -4..11
                const uint32_t Plo = (pc[1] << 24) | (pc[2] << 16) | (pc[3] << 8);         // bytes 4..15, then 8 zero bits
#pragma unroll
                for (int t = 0; t < 16; t++) {
                    // 5-mer ending at byte t = codes of bytes t-4..t
                    const uint32_t idx = t <= 11 ? (Phi >> (2 * (11 - t))) & 0x3FFu : (Plo >> (2 * (15 - t) + 8)) & 0x3FFu;
                    // invalid 5-mers go to a per-lane sink slot instead of branching around the reduction
                    red_shared_add((ok >> t & 1u) ? km_base + idx * 4u : km_sink, 1u);
                }
            }
        }
    }
    __syncthreads();
    if (!any) return;   // block-uniform: d_n was read after a barrier by every thread
    // flush: content[b][c] += count ; qual[b][c] += sumq - 33*count
    unsigned long long* content = stats;
    unsigned long long* qualsum = stats + 8 * C;
    for (int i = threadIdx.x; i < 8 * CS_TILE; i += CS_THREADS) {
        const int bin = i / CS_TILE, col = i % CS_TILE;
        const uint32_t v = packed[bin][col];
        if (v) {
            const int j = col >> 5, ln = col & 31;
            const int64_t c = c0 + 16 * ln + j;
            const long long cnt = v >> 20, sq = v & 0xFFFFFu;
            atomicAdd(&content[(int64_t)bin * C + c], (unsigned long long)cnt);
            atomicAdd(&qualsum[(int64_t)bin * C + c], (unsigned long long)(sq - 33 * cnt));
        }
    }
    if (DO_KMER) {
        unsigned long long* tail = stats + 16 * C;
        for (int i = threadIdx.x; i < 1024; i += CS_THREADS)
            if (kmer[i]) {
                atomicAdd(&tail[FPL_STATS_KMER + i], (unsigned long long)kmer[i]);
                if (kmer_also) atomicAdd(&kmer_also[i], (unsigned long long)kmer[i]);
            }
    }
}

// kmer_to: where the 5-mer counts of this launch go besides `stats` (nullptr = nowhere else); do_kmer = false skips them
void launch_cycle_stats(const uint8_t* seq, const uint8_t* qual, const StatSeg* segs, int64_t nseg, int64_t max_len,
                        unsigned long long* stats, int64_t C, bool do_kmer, unsigned long long* kmer_also,
                        cudaStream_t stream) {
    if (nseg == 0 || max_len <= 0) return;
    dim3 grid((unsigned)((max_len + CS_TILE - 1) / CS_TILE), (unsigned)((nseg + CS_GROUP - 1) / CS_GROUP));
    if (do_kmer) k_cycle_stats<true><<<grid, CS_THREADS, 0, stream>>>(seq, qual, segs, nseg, stats, C, kmer_also);
    else k_cycle_stats<false><<<grid, CS_THREADS, 0, stream>>>(seq, qual, segs, nseg, stats, C, nullptr);
}

// ------------------------------------------------------------------------------------------------------------------
// k_kmer_fix: the post-filter 5-mer table is the pre-filter one minus the 5-mers that do not lie inside a passing
// segment (trimmed ends, split gaps, failed or dropped reads).  A 5-mer ending at read position e (e >= 4) survives
// iff a passing segment [a, b) has a + 4 <= e < b (Stats::statRead skips the first four cycles of every read it is
// given, src/stats.cpp:309-311).  The removed ranges are short for almost every read.
// ------------------------------------------------------------------------------------------------------------------
#define KF_WARPS 8
__global__ void __launch_bounds__(KF_WARPS * 32)
k_kmer_fix(DevBatch b, const fpl_read_result* __restrict__ res, unsigned long long* __restrict__ post_kmer) {
    __shared__ uint32_t rem[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) rem[i] = 0;
    __syncthreads();
    const int wid = threadIdx.x >> 5, lane = lane_id();
    const uint32_t rem_base = shared_addr(&rem[0]);
    const int64_t nwarps = (int64_t)gridDim.x * KF_WARPS;
    for (int64_t r = (int64_t)blockIdx.x * KF_WARPS + wid; r < b.n_reads; r += nwarps) {
        const int L = b.lens[r];
        if (L < 5) continue;
        const uint8_t* seq = b.seq + b.offsets[r];
        const fpl_read_result* o = &res[r];
        // kept intervals of ending positions, in read order
        int ks[2], ke[2], nk = 0;
        const int nseg = o->n_segments;
        for (int k = 0; k < nseg; k++)
            if (o->seg_result[k] == FPL_PASS_FILTER) { ks[nk] = o->seg_lo[k] + 4; ke[nk] = o->seg_lo[k] + o->seg_len[k]; nk++; }
        // removed ranges: [4, ks0) [ke0, ks1) [ke1, L)
        int from = 4;
        for (int k = 0; k <= nk; k++) {
            const int to = k < nk ? min(ks[k], L) : L;
            for (int e0 = from; e0 < to; e0 += 32) {
                const int e = e0 + lane;
                if (e < to) {
                    const uint32_t c4 = kmer_code(seq[e - 4]), c3 = kmer_code(seq[e - 3]), c2 = kmer_code(seq[e - 2]),
                                   c1 = kmer_code(seq[e - 1]), c0 = kmer_code(seq[e]);
                    if (((c4 | c3 | c2 | c1 | c0) & 8u) == 0u)
                        red_shared_add(rem_base + (((c4 << 8) | (c3 << 6) | (c2 << 4) | (c1 << 2) | c0) << 2), 1u);
                }
            }
            if (k < nk) from = max(from, ke[k]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x)
        if (rem[i]) atomicAdd(&post_kmer[i], 0ull - (unsigned long long)rem[i]);
}

void launch_kmer_fix(const DevBatch& b, const fpl_read_result* res, unsigned long long* post_kmer, cudaStream_t stream) {
    if (b.n_reads == 0) return;
    const int64_t want = (b.n_reads + KF_WARPS - 1) / KF_WARPS;
    const unsigned grid = (unsigned)(want < 148 * 16 ? want : 148 * 16);
    k_kmer_fix<<<grid, KF_WARPS * 32, 0, stream>>>(b, res, post_kmer);
}

// ------------------------------------------------------------------------------------------------------------------
#define RQ_WARPS 8

namespace {
// median: smallest char m with sum_{c<=m} hist[c] > len>>1 (src/stats.cpp:351-361); lane l owns bins 4l..4l+3
__device__ __forceinline__ uint8_t hist_median(const uint32_t* h, int len, int lane, uint32_t (&own)[4]) {
    own[0] = h[4 * lane]; own[1] = h[4 * lane + 1]; own[2] = h[4 * lane + 2]; own[3] = h[4 * lane + 3];
    if (len <= 0) return 0;
    const int half = len >> 1;
    const int tot = (int)(own[0] + own[1] + own[2] + own[3]);
    int run = warp_incl_scan(tot) - tot;
    int m = 1 << 30;
    run += own[0]; if (run > half) m = min(m, 4 * lane);
    run += own[1]; if (run > half) m = min(m, 4 * lane + 1);
    run += own[2]; if (run > half) m = min(m, 4 * lane + 2);
    run += own[3]; if (run > half) m = min(m, 4 * lane + 3);
    return (uint8_t)__reduce_min_sync(0xffffffffu, m);
}

// h[q] += delta for the bytes qp[0..n): 16-byte vector body, byte head/tail (delta = 1 or 0xFFFFFFFF)
__device__ __forceinline__ void hist_bytes(uint32_t hbase, const uint8_t* qp, int n, uint32_t delta, int lane) {
    const int head = min(n, (int)((16 - (reinterpret_cast<uintptr_t>(qp) & 15)) & 15));
    if (lane < head) red_shared_add(hbase + ((uint32_t)(qp[lane] & 127) << 2), delta);
    const int nvec = (n - head) >> 4;
    const uint4* vp = reinterpret_cast<const uint4*>(qp + head);
    for (int i = lane; i < nvec; i += 32) {
        const uint4 v = __ldg(vp + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int j = 0; j < 4; j++) red_shared_add(hbase + (((w[k] >> (8 * j)) & 127u) << 2), delta);
    }
    const int done = head + (nvec << 4);
    if (done + lane < n) red_shared_add(hbase + ((uint32_t)(qp[done + lane] & 127) << 2), delta);   // < 16 tail bytes
}
}  // namespace

// Shared-memory atomics are the fast way to histogram on this part (tools/ubench_hist.cu: a 128-bin table updated
// with atomicAdd by 8 warps streams quality bytes at HBM speed, 4x faster than lane-private read-modify-write).
//
// One warp per input read: the histogram of the whole read gives the pre-filter median and mBaseQualHistogram; the
// histogram of each passing segment is derived from it by subtracting the (short) removed ends — or counted directly
// when the segment is the smaller part — and gives the post-filter median and histogram.  One pass over the
// quality bytes serves both Stats objects.
__global__ void __launch_bounds__(RQ_WARPS * 32)
k_read_qual(DevBatch b, unsigned long long* __restrict__ stats_pre, unsigned long long* __restrict__ stats_post, int64_t C,
            fpl_read_result* __restrict__ res, bool pre_only) {
    __shared__ uint32_t hist[RQ_WARPS][2][128];        // per warp: [0] whole read, [1] current segment
    __shared__ uint32_t block_hist[2][128];            // this block's reads -> one flush per Stats block
    __shared__ unsigned long long block_misc[2][2];    // reads, length sum
    const int wid = threadIdx.x >> 5, lane = lane_id();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) (&block_hist[0][0])[i] = 0;
    if (threadIdx.x < 4) (&block_misc[0][0])[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long* tail[2] = {stats_pre + 16 * C, stats_post + 16 * C};
    uint32_t* hfull = hist[wid][0];
    uint32_t* hseg = hist[wid][1];
    const uint32_t hfull_s = shared_addr(hfull), hseg_s = shared_addr(hseg);
    const int64_t nwarps = (int64_t)gridDim.x * RQ_WARPS;
    for (int64_t r = (int64_t)blockIdx.x * RQ_WARPS + wid; r < b.n_reads; r += nwarps) {
        const uint8_t* qp = b.qual + b.offsets[r];
        const int L = b.lens[r];
        fpl_read_result* o = &res[r];
        for (int i = lane; i < 128; i += 32) hfull[i] = 0;
        __syncwarp();
        hist_bytes(hfull_s, qp, L, 1u, lane);
        __syncwarp();
        uint32_t own[4];
        const uint8_t med = hist_median(hfull, L, lane, own);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (own[k]) atomicAdd(&block_hist[0][4 * lane + k], own[k]);
        if (lane == 0) {
            atomicAdd(&block_misc[0][0], 1ull);
            atomicAdd(&block_misc[0][1], (unsigned long long)L);
            if (L > 0) {
                atomicAdd(&tail[0][FPL_STATS_MEDHIST + med], 1ull);
                atomicAdd(&tail[0][FPL_STATS_MEDBASES + med], (unsigned long long)L);
            }
            o->pre_median_qual = med;
        }
        const int nseg = pre_only ? 0 : o->n_segments;   // --mask/--break: the post-filter part is k_ext_seg_qual's
        for (int k = 0; k < nseg; k++) {
            if (o->seg_result[k] != FPL_PASS_FILTER) continue;      // warp-uniform
            const int a = o->seg_lo[k], n = o->seg_len[k];
            __syncwarp();
            if (L - n <= n) {            // copy the read's histogram and take the removed ends out
                for (int i = lane; i < 128; i += 32) hseg[i] = hfull[i];
                __syncwarp();
                hist_bytes(hseg_s, qp, a, 0xFFFFFFFFu, lane);
                hist_bytes(hseg_s, qp + a + n, L - a - n, 0xFFFFFFFFu, lane);
            } else {
                for (int i = lane; i < 128; i += 32) hseg[i] = 0;
                __syncwarp();
                hist_bytes(hseg_s, qp + a, n, 1u, lane);
            }
            __syncwarp();
            const uint8_t smed = hist_median(hseg, n, lane, own);
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (own[j]) atomicAdd(&block_hist[1][4 * lane + j], own[j]);
            if (lane == 0) {
                atomicAdd(&block_misc[1][0], 1ull);
                atomicAdd(&block_misc[1][1], (unsigned long long)n);
                if (n > 0) {
                    atomicAdd(&tail[1][FPL_STATS_MEDHIST + smed], 1ull);
                    atomicAdd(&tail[1][FPL_STATS_MEDBASES + smed], (unsigned long long)n);
                }
                o->seg_median_qual[k] = smed;
            }
        }
        __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int which = i >> 7, bin = i & 127;
        if (block_hist[which][bin]) atomicAdd(&tail[which][FPL_STATS_QUALHIST + bin], (unsigned long long)block_hist[which][bin]);
    }
    if (threadIdx.x < 2 && block_misc[threadIdx.x][0]) {
        atomicAdd(&tail[threadIdx.x][FPL_STATS_READS], block_misc[threadIdx.x][0]);
        atomicAdd(&tail[threadIdx.x][FPL_STATS_LENSUM], block_misc[threadIdx.x][1]);
    }
}

void launch_read_qual(const DevBatch& b, unsigned long long* stats_pre, unsigned long long* stats_post, int64_t C,
                      fpl_read_result* res, bool pre_only, cudaStream_t stream) {
    if (b.n_reads == 0) return;
    // persistent-ish grid: enough blocks to fill the GPU several times over, each warp strides over the reads
    const int64_t want = (b.n_reads + RQ_WARPS - 1) / RQ_WARPS;
    const unsigned grid = (unsigned)(want < 148 * 64 ? want : 148 * 64);
    k_read_qual<<<grid, RQ_WARPS * 32, 0, stream>>>(b, stats_pre, stats_post, C, res, pre_only);
}

// pre-stats segment list: every input read, full length
__global__ void k_make_preseg(DevBatch b, StatSeg* __restrict__ segs) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= b.n_reads) return;
    StatSeg s;
    s.off = b.offsets[r]; s.len = b.lens[r]; s.read = (int)r; s.slot = 2; s.pad = 0;
    segs[r] = s;
}

void launch_make_preseg(const DevBatch& b, StatSeg* segs, cudaStream_t stream) {
    if (b.n_reads == 0) return;
    k_make_preseg<<<(unsigned)((b.n_reads + 255) / 256), 256, 0, stream>>>(b, segs);
}
