// Stats::statRead (src/stats.cpp:265-375) as two kernels over a list of segments
// (pre-filter Stats: every input read; post-filter Stats: every passing segment):
//
//  k_cycle_stats  column-tiled: a CTA owns a tile of 1024 cycles x a group of segments.  Every lane owns its
//                 cycle columns exclusively, so the per-(base-bin, cycle) counters live in shared memory and are
//                 updated with plain read-modify-write (no atomics, no bank conflicts: column j*32+lane).
//                 One 32-bit word packs (count << 20 | sum of raw quality chars); the tile is flushed once to
//                 the global int64 arrays mCycleBaseContents / mCycleBaseQual.  Also counts the 5-mers (mKmer).
//  k_read_qual    row-shaped: a warp owns a segment, builds its quality histogram in lane-private shared memory
//                 columns, adds it to mBaseQualHistogram and derives the per-read median quality
//                 (mMedianReadQualHistogram / mMedianReadQualBases / mReads / mLengthSum).
#include "fpl_device.cuh"

#define CS_THREADS 256
#define CS_WARPS (CS_THREADS / 32)
#define CS_WARP_CYCLES 128                       // 4 bytes per lane
#define CS_TILE (CS_WARPS * CS_WARP_CYCLES)      // 1024 cycles per CTA
#define CS_GROUP 1024                            // segments per CTA; packed counter: count < 4096, sum < 2^20

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

// unaligned 4-byte fetch through two aligned 32-bit loads (segment starts are arbitrary byte offsets)
__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3) * 8;
    uint32_t lo = __ldg(w);
    if (sh == 0) return lo;
    uint32_t hi = __ldg(w + 1);
    return __funnelshift_r(lo, hi, sh);
}

}  // namespace

__global__ void __launch_bounds__(CS_THREADS)
k_cycle_stats(const uint8_t* __restrict__ seqbuf, const uint8_t* __restrict__ qualbuf, const StatSeg* __restrict__ segs,
              int64_t nseg, unsigned long long* __restrict__ stats, int64_t C) {
    // packed[bin][j*32 + lane] for the warp's 128 cycles: cycle c0w + 4*lane + j
    __shared__ uint32_t packed[CS_WARPS][8][CS_WARP_CYCLES];
    __shared__ uint32_t kmer[1024];
    const int wid = threadIdx.x >> 5, lane = lane_id();
    const int64_t c0w = (int64_t)blockIdx.x * CS_TILE + (int64_t)wid * CS_WARP_CYCLES;  // first cycle of this warp
    const int64_t g0 = (int64_t)blockIdx.y * CS_GROUP;
    const int64_t g1 = min(nseg, g0 + CS_GROUP);
    for (int i = threadIdx.x; i < 1024; i += CS_THREADS) kmer[i] = 0;
    uint32_t (*mine)[CS_WARP_CYCLES] = packed[wid];
    for (int i = lane; i < 8 * CS_WARP_CYCLES; i += 32) (&mine[0][0])[i] = 0;
    __syncthreads();
    const int64_t cl = c0w + 4 * lane;  // this lane's first cycle
    for (int64_t s = g0; s < g1; s++) {
        const StatSeg sg = segs[s];
        if ((int64_t)sg.len <= c0w) continue;   // warp-uniform
        const uint8_t* sp = seqbuf + sg.off;
        const uint8_t* qp = qualbuf + sg.off;
        uint32_t sw = 0, qw = 0, prev = 0;
        const bool active = cl < sg.len;
        if (active) {
            sw = load_u32_unaligned(sp + cl);
            qw = load_u32_unaligned(qp + cl);
        }
        // the 4 bases before this lane's word, for the 5-mers: previous lane's word, or a load for lane 0
        prev = __shfl_up_sync(0xffffffffu, sw, 1);
        if (lane == 0) prev = c0w >= 4 ? load_u32_unaligned(sp + c0w - 4) : 0u;
        if (!active) continue;
        const int nvalid = (int)min((int64_t)4, (int64_t)sg.len - cl);
        // codes of the 8 bases prev[0..3], sw[0..3]
        uint32_t code[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            code[j] = kmer_code((prev >> (8 * j)) & 0xFFu);
            code[4 + j] = kmer_code((sw >> (8 * j)) & 0xFFu);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (j < nvalid) {
                const uint32_t base = (sw >> (8 * j)) & 0xFFu;
                const uint32_t q = (qw >> (8 * j)) & 0xFFu;
                mine[base & 7u][j * 32 + lane] += (1u << 20) + q;
                // 5-mer ending at cycle cl+j: needs cl+j >= 4 and five valid bases
                if (cl + j >= 4) {
                    const uint32_t c4 = code[j] , c3 = code[j + 1], c2 = code[j + 2], c1 = code[j + 3], c0 = code[j + 4];
                    if (((c4 | c3 | c2 | c1 | c0) & 8u) == 0u)
                        atomicAdd(&kmer[(c4 << 8) | (c3 << 6) | (c2 << 4) | (c1 << 2) | c0], 1u);
                }
            }
        }
    }
    __syncwarp();
    // flush: content[b][c] += count ; qual[b][c] += sumq - 33*count
    unsigned long long* content = stats;
    unsigned long long* qualsum = stats + 8 * C;
    for (int i = lane; i < 8 * CS_WARP_CYCLES; i += 32) {
        const int bin = i / CS_WARP_CYCLES, col = i % CS_WARP_CYCLES;
        const uint32_t v = mine[bin][col];
        if (v) {
            const int j = col >> 5, ln = col & 31;
            const int64_t c = c0w + 4 * ln + j;
            const long long cnt = v >> 20, sq = v & 0xFFFFFu;
            atomicAdd(&content[(int64_t)bin * C + c], (unsigned long long)cnt);
            atomicAdd(&qualsum[(int64_t)bin * C + c], (unsigned long long)(sq - 33 * cnt));
        }
    }
    __syncthreads();
    unsigned long long* tail = stats + 16 * C;
    for (int i = threadIdx.x; i < 1024; i += CS_THREADS)
        if (kmer[i]) atomicAdd(&tail[FPL_STATS_KMER + i], (unsigned long long)kmer[i]);
}

void launch_cycle_stats(const uint8_t* seq, const uint8_t* qual, const StatSeg* segs, int64_t nseg, int64_t max_len,
                        unsigned long long* stats, int64_t C, cudaStream_t stream) {
    if (nseg == 0 || max_len <= 0) return;
    dim3 grid((unsigned)((max_len + CS_TILE - 1) / CS_TILE), (unsigned)((nseg + CS_GROUP - 1) / CS_GROUP));
    k_cycle_stats<<<grid, CS_THREADS, 0, stream>>>(seq, qual, segs, nseg, stats, C);
}

// ------------------------------------------------------------------------------------------------------------------
#define RQ_WARPS 4
#define RQ_CHUNK (32 * 60000)   // bases per flush of the 16-bit lane-private counters (< 65536 per lane)

__global__ void __launch_bounds__(RQ_WARPS * 32)
k_read_qual(const uint8_t* __restrict__ qualbuf, const StatSeg* __restrict__ segs, int64_t nseg,
            unsigned long long* __restrict__ stats, int64_t C, fpl_read_result* __restrict__ res) {
    __shared__ uint16_t hist[RQ_WARPS][128][32];       // lane-private columns: no atomics
    __shared__ uint32_t total[RQ_WARPS][128];          // per-segment histogram
    __shared__ uint32_t block_hist[128];               // all segments of this block -> one flush to mBaseQualHistogram
    __shared__ unsigned long long block_misc[2];       // reads, length sum
    const int wid = threadIdx.x >> 5, lane = lane_id();
    for (int i = threadIdx.x; i < 128; i += blockDim.x) block_hist[i] = 0;
    if (threadIdx.x < 2) block_misc[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long* tail = stats + 16 * C;
    const int64_t s = (int64_t)blockIdx.x * RQ_WARPS + wid;
    if (s < nseg) {
        const StatSeg sg = segs[s];
        if (sg.read >= 0) {   // a real segment (pre: every read, even empty; post: passing segments only)
            const uint8_t* qp = qualbuf + sg.off;
            const int len = sg.len;
            for (int b = lane; b < 128; b += 32) total[wid][b] = 0;
            for (int base0 = 0; base0 < len; base0 += RQ_CHUNK) {
                const int n = min(RQ_CHUNK, len - base0);
                uint32_t* z = reinterpret_cast<uint32_t*>(&hist[wid][0][0]);
                for (int i = lane; i < 128 * 32 / 2; i += 32) z[i] = 0;
                __syncwarp();
                for (int i = lane; i < n; i += 32) {
                    const uint8_t q = qp[base0 + i];
                    hist[wid][q & 127][lane]++;
                }
                __syncwarp();
                // lane b sums bins b, b+32, b+64, b+96 over the 32 columns (rotated start: conflict-free)
                for (int b = lane; b < 128; b += 32) {
                    uint32_t acc = 0;
                    for (int k = 0; k < 32; k++) acc += hist[wid][b][(k + lane) & 31];
                    total[wid][b] += acc;
                }
                __syncwarp();
            }
            // median: smallest char m with sum_{c<=m} hist[c] > len>>1 (src/stats.cpp:351-361)
            uint8_t median = 0;
            if (len > 0) {
                const int half = len >> 1;
                // 4 bins per lane, in order: lane l owns bins 4l..4l+3
                uint32_t h0 = total[wid][4 * lane], h1 = total[wid][4 * lane + 1], h2 = total[wid][4 * lane + 2],
                         h3 = total[wid][4 * lane + 3];
                int incl = warp_incl_scan((int)(h0 + h1 + h2 + h3));
                int excl = incl - (int)(h0 + h1 + h2 + h3);
                int m = 1 << 30;
                int run = excl;
                run += h0; if (run > half) m = min(m, 4 * lane);
                run += h1; if (run > half) m = min(m, 4 * lane + 1);
                run += h2; if (run > half) m = min(m, 4 * lane + 2);
                run += h3; if (run > half) m = min(m, 4 * lane + 3);
                m = __reduce_min_sync(0xffffffffu, m);
                median = (uint8_t)m;
            }
            for (int b = lane; b < 128; b += 32)
                if (total[wid][b]) atomicAdd(&block_hist[b], total[wid][b]);
            if (lane == 0) {
                atomicAdd(&block_misc[0], 1ull);
                atomicAdd(&block_misc[1], (unsigned long long)len);
                if (len > 0) {
                    atomicAdd(&tail[FPL_STATS_MEDHIST + median], 1ull);
                    atomicAdd(&tail[FPL_STATS_MEDBASES + median], (unsigned long long)len);
                }
                if (sg.slot == 2) res[sg.read].pre_median_qual = median;
                else res[sg.read].seg_median_qual[sg.slot] = median;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 128; i += blockDim.x)
        if (block_hist[i]) atomicAdd(&tail[FPL_STATS_QUALHIST + i], (unsigned long long)block_hist[i]);
    if (threadIdx.x == 0 && block_misc[0]) {
        atomicAdd(&tail[FPL_STATS_READS], block_misc[0]);
        atomicAdd(&tail[FPL_STATS_LENSUM], block_misc[1]);
    }
}

void launch_read_qual(const uint8_t* qual, const StatSeg* segs, int64_t nseg, unsigned long long* stats, int64_t C,
                      fpl_read_result* res, cudaStream_t stream) {
    if (nseg == 0) return;
    k_read_qual<<<(unsigned)((nseg + RQ_WARPS - 1) / RQ_WARPS), RQ_WARPS * 32, 0, stream>>>(qual, segs, nseg, stats, C, res);
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
