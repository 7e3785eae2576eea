// Stats::statRead (src/stats.cpp:265-375) as kernels over a list of segments
// (pre-filter Stats: every input read; post-filter Stats: every passing segment):
//
//  k_cycle_stats  column-tiled over the LENGTH-SORTED segment list: a CTA owns a 512-byte tile of the segments'
//                 16-byte-aligned byte ranges x a group of up to 4000 segments; its warps take the group's segments
//                 round-robin, one aligned 16-byte vector of sequence and of quality per lane (cp.async ring, several
//                 segments in flight).  The per-(base-bin, cycle) counters live in shared memory, one 32-bit word =
//                 (count << 20 | sum of raw quality chars), laid out so that the 32 lanes of a reduction hit 32 banks
//                 whatever the segment's misalignment; the tile is flushed once to the global int64 arrays
//                 mCycleBaseContents / mCycleBaseQual.  The 5-mer variant also fills lane-private 1024-bin tables
//                 (mKmer) from 2-bit codes packed by multiplication, with an exact A/C/G/T/U test.
//  k_read_qual    row-shaped: a warp owns a read, builds its quality histogram with shared-memory atomics
//                 (16-byte vector loads, four per lane in flight), adds it to mBaseQualHistogram and derives the
//                 per-read median quality (mMedianReadQualHistogram / mMedianReadQualBases / mReads / mLengthSum)
//                 for BOTH Stats objects: a passing segment's histogram is the read's minus the removed ends.
//  k_kmer_fix     post-filter 5-mer table = pre-filter table - the 5-mers outside the passing segments.
#include "fpl_device.cuh"
#include "fpl_stats.h"

#include <atomic>
#include <stdlib.h>
#include <cub/cub.cuh>

#define CS_NT_KMER 1024                          // threads per CTA with the 5-mer tables (one CTA per SM: 145 KB of shared memory)
#define CS_NT_PLAIN 512
#define CS_TILE 512                              // bytes of every segment a CTA looks at: one 16-byte vector per lane
#define CS_ROWW 33                               // words per counter row: 32 lanes + the carry column
#define CS_BINW (16 * CS_ROWW + 16)              // words per base bin: a multiple of 32, so the bank is the lane's whatever the bin
#define CS_GROUP 4000                            // segments per CTA; packed counter: count <= 4095, sum of q < 2^20
#define CS_STAGE 1000                            // descriptors staged in shared memory at a time
#define CS_SLOT 1040                             // ring slot: 16 B in front of the tile (its last 4 are used), 512 B sequence, 512 B quality
#define CS_DEPTH_KMER 1                          // segments in flight per warp (the 5-mer variant is bound by the logic pipe)
#define CS_DEPTH_PLAIN 3
#define CS_SMEM_BASE (8 * CS_BINW * 4 + CS_STAGE * 16)            // counters + staged descriptors
#define CS_MBAR_BYTES(NT, DEPTH) (((NT) / 32) * ((DEPTH) + 1) * 8)          // one mbarrier per ring slot (TMA staging)
#define CS_SMEM_PLAIN (CS_SMEM_BASE + (CS_NT_PLAIN / 32) * (CS_DEPTH_PLAIN + 1) * CS_SLOT + CS_MBAR_BYTES(CS_NT_PLAIN, CS_DEPTH_PLAIN))
#define CS_SMEM_KMER (CS_SMEM_BASE + (CS_NT_KMER / 32) * (CS_DEPTH_KMER + 1) * CS_SLOT + 1024 * 32 * 4 + CS_MBAR_BYTES(CS_NT_KMER, CS_DEPTH_KMER))   // + lane-private 5-mer tables

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

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {   // PTX semantics: selector bit 3 = sign fill
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {  // keeps the work on the IMAD pipe
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// ---- 1-D TMA (cp.async.bulk) staging: one lane copies a whole 512-byte tile row, completion on an mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int IMM>
__device__ __forceinline__ void red_shared_add_imm(uint32_t saddr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0 + %2], %1;" ::"r"(saddr), "r"(v), "n"(IMM) : "memory");
}

// Counter word of (bin, column): the CTA's columns are the cycles 512*tile - 15 + col, col in [0, 527).  Column col lives
// in row col & 15 at position col >> 4 (33 positions per row), so that the 32 lanes of one instruction — which hold the
// columns 16*lane + m for one m — hit 32 different banks (the bin stride is a multiple of 32 words).
// A lane's vector starts S columns into its 16-column stripe (S = 15 - (segment address & 15), warp-uniform): byte J
// is column 16*lane + J + S = row (J+S)&15, position lane + ((J+S)>>4).
//   wm: base & 7 in every byte; qm: quality chars; cw: 0x10 in every byte that counts (0 elsewhere, with qm = 0 there)
// Per base: one PRMT for the bin, one for the value (quality | 1 << 20), one IMAD for the address, one reduction.
template <int S, int J>
__device__ __forceinline__ void count1(const uint32_t (&wm)[4], const uint32_t (&qm)[4], const uint32_t (&cw)[4], uint32_t pk_lane) {
    constexpr int jj = J & 3, m = J + S, z = 8 | jj;                      // selector z: sign fill of a 7-bit byte = 0
    const uint32_t bin = prmt(wm[J >> 2], 0u, jj | z << 4 | z << 8 | z << 12);
    const uint32_t val = prmt(qm[J >> 2], cw[J >> 2], jj | z << 4 | (4 + jj) << 8 | z << 12);
    red_shared_add_imm<4 * ((m & 15) * CS_ROWW + (m >> 4))>(mad_u32(bin, CS_BINW * 4u, pk_lane), val);
}
template <int S>
__device__ __forceinline__ void count16(const uint32_t (&wm)[4], const uint32_t (&qm)[4], const uint32_t (&cw)[4], uint32_t pk) {
    count1<S, 0>(wm, qm, cw, pk); count1<S, 1>(wm, qm, cw, pk); count1<S, 2>(wm, qm, cw, pk); count1<S, 3>(wm, qm, cw, pk);
    count1<S, 4>(wm, qm, cw, pk); count1<S, 5>(wm, qm, cw, pk); count1<S, 6>(wm, qm, cw, pk); count1<S, 7>(wm, qm, cw, pk);
    count1<S, 8>(wm, qm, cw, pk); count1<S, 9>(wm, qm, cw, pk); count1<S, 10>(wm, qm, cw, pk); count1<S, 11>(wm, qm, cw, pk);
    count1<S, 12>(wm, qm, cw, pk); count1<S, 13>(wm, qm, cw, pk); count1<S, 14>(wm, qm, cw, pk); count1<S, 15>(wm, qm, cw, pk);
}

// The same for a vector whose 16 bytes all count (warp-uniform fast path): both operands of the reduction come from
// the FMA pipe, which issues beside the logic pipe (tools/ubench_pipes.cu: LOP3 + IDP.4A together run at 0.94
// warp-instructions per clock and sub-partition, LOP3 alone at 0.48).  wb: (base & 7) << 4 in every byte, so that
// one dp4a with the weight 136 on byte J is bin * CS_BINW * 4 (+ the lane's base address); qm: quality chars.
// wa[j] = 136 << 8j and wv[j] = 1 << 8j arrive as run-time values (uniform registers for the whole loop): as literals
// the compiler re-materialises the eight of them with UMOVs in front of every vector.
template <int S, int J>
__device__ __forceinline__ void count1f(const uint32_t (&wb)[4], const uint32_t (&qm)[4], uint32_t pk_lane,
                                        const uint32_t (&wa)[4], const uint32_t (&wv)[4]) {
    constexpr int jj = J & 3, m = J + S;
    static_assert(CS_BINW * 4 == 136 * 16, "dp4a weight of the bin stride");
    const uint32_t addr = __dp4a(wb[J >> 2], wa[jj], pk_lane);
    const uint32_t val = __dp4a(qm[J >> 2], wv[jj], 1u << 20);
    red_shared_add_imm<4 * ((m & 15) * CS_ROWW + (m >> 4))>(addr, val);
}
template <int S>
__device__ __forceinline__ void count16f(const uint32_t (&wb)[4], const uint32_t (&qm)[4], uint32_t pk, const uint32_t (&wa)[4],
                                         const uint32_t (&wv)[4]) {
    count1f<S, 0>(wb, qm, pk, wa, wv); count1f<S, 1>(wb, qm, pk, wa, wv); count1f<S, 2>(wb, qm, pk, wa, wv); count1f<S, 3>(wb, qm, pk, wa, wv);
    count1f<S, 4>(wb, qm, pk, wa, wv); count1f<S, 5>(wb, qm, pk, wa, wv); count1f<S, 6>(wb, qm, pk, wa, wv); count1f<S, 7>(wb, qm, pk, wa, wv);
    count1f<S, 8>(wb, qm, pk, wa, wv); count1f<S, 9>(wb, qm, pk, wa, wv); count1f<S, 10>(wb, qm, pk, wa, wv); count1f<S, 11>(wb, qm, pk, wa, wv);
    count1f<S, 12>(wb, qm, pk, wa, wv); count1f<S, 13>(wb, qm, pk, wa, wv); count1f<S, 14>(wb, qm, pk, wa, wv); count1f<S, 15>(wb, qm, pk, wa, wv);
}

// 0xFF in byte k of the result iff bit k of the nibble n
__device__ __forceinline__ uint32_t nibble_to_bytes(uint32_t n) { return ((n * 0x00204081u) & 0x01010101u) * 0xFFu; }

// 5-mer ending at byte T of the lane's vector.  P: 2-bit codes, oldest first, the 5-mer's in bits [22-SHL .. 32-SHL); the
// table word is [code][lane] (no bank conflicts); an invalid 5-mer adds 0 (no branch around the reduction — a reduction
// predicated on the validity bit is compiled into BSSY / BRA / BSYNC around it: four instructions instead of two).
// Left shifts as multiplies: they run on the IMAD pipe.  (Right shifts as mul.hi were measured too: IMAD.HI is slow on
// this part — 13.7 ms against 11.4 ms for the kernel; mask + shift-and-add compiles to the same four instructions.)
template <int SHL, int T>
__device__ __forceinline__ void kmer1(uint32_t P, uint32_t ok, uint32_t km_lane) {
    const uint32_t idx = (SHL ? mad_u32(P, 1u << SHL, 0u) : P) >> 22;
    red_shared_add(mad_u32(idx, 128u, km_lane), (ok >> T) & 1u);
}

// the same when the 5-mer is known to count: the value is the immediate 1
template <int SHL>
__device__ __forceinline__ void kmer1f(uint32_t P, uint32_t km_lane) {
    const uint32_t idx = (SHL ? mad_u32(P, 1u << SHL, 0u) : P) >> 22;
    red_shared_add(mad_u32(idx, 128u, km_lane), 1u);
}

// Per 4 bytes of sequence: in7 = (nibble of "byte is not one of A, C, G, T, U") << 7, exact for any byte value, and
// pc = four 2-bit codes c' = (b >> 1) & 3, oldest first.  g = b2 & ~b1 (set for T/U); with bit 7 and bits 2..1 masked
// off, a valid byte is 0x41 (A, C, G) or, with g, 0x51 (T, U): d == 0.  d < 0x80, so d + 0x7f sets bit 7 iff d != 0 and
// cannot carry into the next byte; the byte's own bit 7 is or-ed in by the same LOP3.  The two gathers are dp4a.
__device__ __forceinline__ void encode4(uint32_t w, uint32_t one, uint32_t& nz, uint32_t& pc) {
    const uint32_t x1 = w >> 1;
    const uint32_t g = (x1 >> 1) & ~x1 & 0x01010101u;
    const uint32_t d = ((w & 0x79797979u) | g) ^ 0x41414141u ^ mad_u32(g, 16u, 0u);
    nz = (mad_u32(d, one, 0x7f7f7f7fu) | w) & 0x80808080u;
    pc = __dp4a(x1 & 0x03030303u, 0x01041040u, 0u);
}

}  // namespace

// Sorted segment descriptor (by length, longest first): 16 bytes.
struct SegD {
    int64_t off;
    int32_t len;
    int32_t pad;
};
// The same as a CTA stages it for its tile: a = 16-byte aligned byte offset of the tile's first vector, lim = how many
// of the tile's bytes lie in front of the segment's end (> 0), sh = segment address & 15.
struct __align__(16) TileSeg {
    int64_t a;
    int32_t lim;
    int32_t sh;
};

// DO_KMER: also count the 5-mers; their table is flushed to `stats` and, if given, to `kmer_also` (the post-filter
// block: post 5-mers = pre 5-mers - the ones k_kmer_fix finds outside the passing segments).
//
// Grid: x = 512-byte tile of the segments' 16-byte-aligned byte ranges, y = group of CS_GROUP segments of the list
// sorted by length (so a group's segments end in the same few tiles and the CTAs beyond the longest one leave at once).
// Every load is an aligned 16-byte vector: the misalignment of a segment (its address & 15, different for every
// post-filter segment) moves the COLUMNS its bytes count into instead of the bytes (count16<S>).
// TMA = true: the tile rows are staged by 1-D bulk copies (cp.async.bulk + mbarrier) issued by one lane per warp instead
// of 32 lanes' cp.async — the variant FPL_CS_TMA=1 selects; both are measured in profiles/README.md.
template <bool DO_KMER, int NT, int DEPTH, bool TMA>
__global__ void __launch_bounds__(NT, DO_KMER ? 1 : 2048 / NT / 2)
k_cycle_stats(const uint8_t* __restrict__ seqbuf, const uint8_t* __restrict__ qualbuf, const SegD* __restrict__ segs,
              int64_t nseg, unsigned long long* __restrict__ stats, int64_t C, unsigned long long* __restrict__ kmer_also,
              uint32_t one, int aligned) {
    extern __shared__ __align__(16) uint8_t cs_smem[];
    uint32_t* packed = reinterpret_cast<uint32_t*>(cs_smem);                       // [8][16][33]: count << 20 | sum of q
    TileSeg* stage = reinterpret_cast<TileSeg*>(cs_smem + 8 * CS_BINW * 4);
    uint8_t* ring = reinterpret_cast<uint8_t*>(stage + CS_STAGE);                  // [warps][DEPTH+1] slots of CS_SLOT bytes
    uint32_t* kmer = reinterpret_cast<uint32_t*>(ring + (NT / 32) * (DEPTH + 1) * CS_SLOT);   // [1024][32]: one column per lane
    uint8_t* mbars = reinterpret_cast<uint8_t*>(kmer) + (DO_KMER ? 1024 * 32 * 4 : 0);       // [warps][DEPTH+1] mbarriers (TMA)
    const int wid = threadIdx.x >> 5, lane = lane_id();
    const int64_t t0 = (int64_t)blockIdx.x * CS_TILE;      // first byte of this tile, relative to the aligned segment start
    const int64_t g0 = (int64_t)blockIdx.y * CS_GROUP;
    const int64_t g1 = min(nseg, g0 + CS_GROUP);
    {
        const int maxlen = segs[g0].len;                   // sorted: the group's longest segment
        if (maxlen <= 0 || t0 >= (int64_t)maxlen + 15) return;
    }
    for (int i = threadIdx.x; i < 8 * CS_BINW; i += NT) packed[i] = 0;
    if (DO_KMER) {
        uint4* kz = reinterpret_cast<uint4*>(kmer);
        for (int i = threadIdx.x; i < 1024 * 32 / 4; i += NT) kz[i] = make_uint4(0, 0, 0, 0);
    }
    const uint32_t pk_lane = shared_addr(packed) + (uint32_t)lane * 4u;
    const uint32_t km_lane = shared_addr(kmer) + (uint32_t)lane * 4u;
    const uint8_t* seq_lane = seqbuf + 16 * lane;
    const uint8_t* qual_lane = qualbuf + 16 * lane;
    // a slot: [16 bytes in front of the tile][512 bytes of sequence][512 bytes of quality]; the word in front of a lane's
    // vector — the previous lane's last word, for lane 0 the front bytes — is at the same offset for every lane
    const uint32_t ring_warp = shared_addr(ring) + (uint32_t)wid * ((DEPTH + 1) * CS_SLOT);
    const uint32_t ring_lane = ring_warp + 16u + (uint32_t)lane * 16u;
    const uint32_t bar_warp = shared_addr(mbars) + (uint32_t)wid * ((DEPTH + 1) * 8);
    const uint32_t wa[4] = {136u * one, (136u << 8) * one, (136u << 16) * one, (136u << 24) * one};   // one == 1
    const uint32_t wv[4] = {one, one << 8, one << 16, one << 24};
    uint32_t phase = 0;                                   // TMA: the parity each slot's mbarrier completes next (bit per slot)
    if (TMA) {
        if (lane == 0)
            for (int d = 0; d <= DEPTH; d++) mbar_init(bar_warp + 8u * d, 1u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int64_t s0 = g0; s0 < g1; s0 += CS_STAGE) {
        __syncthreads();
        // stage the descriptors; the ones that can reach this tile are a prefix (sorted by length)
        int n = 0;
        for (int i0 = 0; i0 < CS_STAGE; i0 += NT) {
            const int i = i0 + threadIdx.x;
            bool reach = false;
            if (i < CS_STAGE && s0 + i < g1) {
                const SegD sg = segs[s0 + i];
                TileSeg ts;
                ts.sh = (int32_t)(sg.off & 15);
                ts.a = sg.off - ts.sh + t0;
                const int64_t lim = (int64_t)ts.sh + sg.len - t0;
                ts.lim = lim > 0 ? (int32_t)lim : 0;
                stage[i] = ts;
                reach = sg.len > 0 && (int64_t)sg.len + 15 > t0;
            }
            n += __syncthreads_count(reach);
        }
        if (n == 0) break;
        // software pipeline: the vectors of the next DEPTH segments of this warp are in flight (cp.async into the warp's
        // ring of DEPTH+1 slots: no registers held, every lane copies and later reads its own 16 bytes) while one is
        // processed; a lane beyond the segment's end copies 0 bytes, which zero-fills its slot bytes
        auto issue = [&](int k, int slot) {
            const TileSeg ts = stage[k];
            if (TMA) {
                // one lane: up to 512 bytes of sequence, the same of quality and (5-mers) the 16 bytes in front of the
                // tile; a tile the segment does not fill is copied up to its last vector, the rest of the slot keeps
                // stale bytes that the edge path masks
                __syncwarp();                              // every lane is done reading the slot that is refilled
                if (lane == 0) {
                    const uint32_t dst = ring_warp + (uint32_t)slot * CS_SLOT, bar = bar_warp + 8u * slot;
                    const uint32_t bytes = ts.lim > 0 ? (uint32_t)min(CS_TILE, (ts.lim + 15) & ~15) : 0u;
                    const uint32_t pb = (DO_KMER && t0 != 0 && bytes) ? 16u : 0u;
                    mbar_expect_tx(bar, 2u * bytes + pb);
                    if (bytes) { bulk_g2s(dst + 16u, seqbuf + ts.a, bytes, bar); bulk_g2s(dst + 528u, qualbuf + ts.a, bytes, bar); }
                    if (pb) bulk_g2s(dst, seqbuf + ts.a - 16, 16u, bar);
                }
                return;
            }
            const bool act = 16 * lane < ts.lim;
            const uint32_t dst = ring_lane + (uint32_t)slot * CS_SLOT;
            const int64_t la = ts.a + (act ? 16 * lane : 0);     // an inactive lane copies 0 bytes (zero fill); the address stays valid
            cp_async16(dst, seqbuf + la, act ? 16 : 0);
            cp_async16(dst + 512, qualbuf + la, act ? 16 : 0);
            // the four bases in front of the tile (lane 0's 5-mers): not needed in the first tile, where no 5-mer ends in
            // front of cycle 4; behind it a >= 512 - 15
            if (DO_KMER && lane == 0 && t0 != 0) cp_async4(dst - 4, seqbuf + ts.a - 4, 4);
        };
        int kk = wid;
#pragma unroll
        for (int d = 0; d < DEPTH; d++) {
            if (kk < n) issue(kk, d);
            if (!TMA) cp_async_commit();
            kk += NT / 32;
        }
        int slot = 0, fill = DEPTH;
        for (int k = wid; k < n; k += (NT / 32)) {
            if (TMA) { mbar_wait(bar_warp + 8u * slot, (phase >> slot) & 1u); phase ^= 1u << slot; }
            else {
                cp_async_wait<DEPTH - 1>();
                // the word in front of a lane's vector was copied by the lane below: a lane's wait covers its own copies
                // only, and the slot refilled further down was read by the neighbours in the previous round
                if (DO_KMER) __syncwarp();
            }
            const uint32_t src = ring_lane + (uint32_t)slot * CS_SLOT;
            const uint4 ns = lds128(src), nq = lds128(src + 512);
            const uint32_t pw = DO_KMER ? lds32(src - 4) : 0u;     // the four bases in front of this lane's vector
            const uint32_t sw[4] = {ns.x, ns.y, ns.z, ns.w};
            uint32_t qm[4] = {nq.x, nq.y, nq.z, nq.w};
            const int sh = stage[k].sh, lim = stage[k].lim;
            if (kk < n) issue(kk, fill);
            if (!TMA) cp_async_commit();
            kk += NT / 32;
            static_assert(((DEPTH + 1) & DEPTH) == 0, "ring slots: a power of two");
            slot = (slot + 1) & DEPTH;
            fill = (fill + 1) & DEPTH;
            if (lim <= 0) continue;                        // warp-uniform: the segment ends in front of this tile
            const bool full = lim >= CS_TILE && t0 != 0;   // warp-uniform: every byte of every lane is a cycle >= 4
            uint32_t kmask = 0xFFFFu;
            if (full) {
                // ---- per-(bin, cycle) counters, all 16 bytes of every lane: address and value by dp4a ----
                const uint32_t wb[4] = {mad_u32(sw[0], 16u, 0u) & 0x70707070u, mad_u32(sw[1], 16u, 0u) & 0x70707070u,
                                        mad_u32(sw[2], 16u, 0u) & 0x70707070u, mad_u32(sw[3], 16u, 0u) & 0x70707070u};
                // aligned: every segment starts on a 16-byte boundary (the pre-filter pass: whole reads in their slots)
                if (aligned) count16f<15>(wb, qm, pk_lane, wa, wv);
                else switch (sh) {
                    case 0: count16f<15>(wb, qm, pk_lane, wa, wv); break;
                    case 1: count16f<14>(wb, qm, pk_lane, wa, wv); break;
                    case 2: count16f<13>(wb, qm, pk_lane, wa, wv); break;
                    case 3: count16f<12>(wb, qm, pk_lane, wa, wv); break;
                    case 4: count16f<11>(wb, qm, pk_lane, wa, wv); break;
                    case 5: count16f<10>(wb, qm, pk_lane, wa, wv); break;
                    case 6: count16f<9>(wb, qm, pk_lane, wa, wv); break;
                    case 7: count16f<8>(wb, qm, pk_lane, wa, wv); break;
                    case 8: count16f<7>(wb, qm, pk_lane, wa, wv); break;
                    case 9: count16f<6>(wb, qm, pk_lane, wa, wv); break;
                    case 10: count16f<5>(wb, qm, pk_lane, wa, wv); break;
                    case 11: count16f<4>(wb, qm, pk_lane, wa, wv); break;
                    case 12: count16f<3>(wb, qm, pk_lane, wa, wv); break;
                    case 13: count16f<2>(wb, qm, pk_lane, wa, wv); break;
                    case 14: count16f<1>(wb, qm, pk_lane, wa, wv); break;
                    default: count16f<0>(wb, qm, pk_lane, wa, wv); break;
                }
            } else {
                // the segment starts or ends in this tile: which of the lane's 16 bytes are cycles of the segment (vmask)
                // and can end a 5-mer (kmask: cycle >= 4)
                uint32_t cw[4] = {0x10101010u, 0x10101010u, 0x10101010u, 0x10101010u};
                const int cyc0 = (int)t0 + 16 * lane - sh;                          // cycle of byte 0, negative in front
                const int lo = max(0, -cyc0), hi = min(16, max(0, lim - 16 * lane));
                const uint32_t vmask = hi > lo ? ((0xFFFFu >> (16 - hi)) & (0xFFFFu << lo)) : 0u;
                const int klo = min(16, max(0, 4 - cyc0));
                kmask = vmask & (0xFFFFu << klo);
                if (vmask != 0xFFFFu) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t bm = nibble_to_bytes((vmask >> (4 * i)) & 15u);
                        qm[i] &= bm; cw[i] &= bm;
                    }
                }
                const uint32_t wm[4] = {sw[0] & 0x07070707u, sw[1] & 0x07070707u, sw[2] & 0x07070707u, sw[3] & 0x07070707u};
                switch (sh) {
                    case 0: count16<15>(wm, qm, cw, pk_lane); break;
                    case 1: count16<14>(wm, qm, cw, pk_lane); break;
                    case 2: count16<13>(wm, qm, cw, pk_lane); break;
                    case 3: count16<12>(wm, qm, cw, pk_lane); break;
                    case 4: count16<11>(wm, qm, cw, pk_lane); break;
                    case 5: count16<10>(wm, qm, cw, pk_lane); break;
                    case 6: count16<9>(wm, qm, cw, pk_lane); break;
                    case 7: count16<8>(wm, qm, cw, pk_lane); break;
                    case 8: count16<7>(wm, qm, cw, pk_lane); break;
                    case 9: count16<6>(wm, qm, cw, pk_lane); break;
                    case 10: count16<5>(wm, qm, cw, pk_lane); break;
                    case 11: count16<4>(wm, qm, cw, pk_lane); break;
                    case 12: count16<3>(wm, qm, cw, pk_lane); break;
                    case 13: count16<2>(wm, qm, cw, pk_lane); break;
                    case 14: count16<1>(wm, qm, cw, pk_lane); break;
                    default: count16<0>(wm, qm, cw, pk_lane); break;
                }
            }
            if (!DO_KMER) continue;
            // ---- 5-mers ending in this lane's 16 bytes: all five bases in ACGTU, end cycle in [4, len) ----
            // per word: the "not ACGTU" flags and four 2-bit codes c' = (b >> 1) & 3 (A=0, C=1, T/U=2, G=3; the flush maps
            // them to base2val's); the word in front of the lane's vector (pw) supplies the four bases before it — every
            // lane encodes its own copy: cheaper than two shuffles plus a lane-0 branch around a fifth encode
            uint32_t nz[4], pc[4], pnz, ppc;
#pragma unroll
            for (int i = 0; i < 4; i++) encode4(sw[i], one, nz[i], pc[i]);
            encode4(pw, one, pnz, ppc);
            // flags << 7: bit 7+t = byte t-4 of the 20 bytes (previous four, then the lane's sixteen)
            const uint32_t pin7 = __dp4a(pnz, 0x08040201u, 0u);
            const uint32_t in01 = __dp4a(nz[0], 0x08040201u, __dp4a(nz[1], 0x80402010u, 0u));
            const uint32_t in23 = __dp4a(nz[2], 0x08040201u, __dp4a(nz[3], 0x80402010u, 0u));
            const uint32_t I27 = mad_u32(in23, 4096u, mad_u32(in01, 16u, pin7));     // 20 flags at bits 7..26
            // 20 codes, oldest first, 2 bits each: Phi = codes of bytes -4..11 (32 bits), Plo = bytes 4..15, then 8 zero bits
            const uint32_t Phi = mad_u32(ppc, 1u << 24, mad_u32(pc[0], 1u << 16, mad_u32(pc[1], 256u, pc[2])));
            const uint32_t Plo = mad_u32(pc[1], 1u << 24, mad_u32(pc[2], 1u << 16, pc[3] * 256u));
            if (full && !__any_sync(0xffffffffu, (I27 & 0x07FFFF80u) != 0u)) {
                // plain bases everywhere in the warp's 512 bytes: every 5-mer is counted with the immediate 1
                kmer1f<0>(Phi, km_lane); kmer1f<2>(Phi, km_lane); kmer1f<4>(Phi, km_lane); kmer1f<6>(Phi, km_lane);
                kmer1f<8>(Phi, km_lane); kmer1f<10>(Phi, km_lane); kmer1f<12>(Phi, km_lane); kmer1f<14>(Phi, km_lane);
                kmer1f<16>(Phi, km_lane); kmer1f<18>(Phi, km_lane); kmer1f<20>(Phi, km_lane); kmer1f<22>(Phi, km_lane);
                kmer1f<8>(Plo, km_lane); kmer1f<10>(Plo, km_lane); kmer1f<12>(Plo, km_lane); kmer1f<14>(Plo, km_lane);
            } else if (full) {
                // some lane holds an N (or another byte outside ACGTU): the reductions add the validity bit (a predicated
                // reduction is compiled into a branch around it: four instructions per 5-mer instead of two)
                const uint32_t I20 = I27 >> 7;
                const uint32_t ok = ~(I20 | (I20 >> 1) | (I20 >> 2) | (I20 >> 3) | (I20 >> 4));   // bit t: bytes t-4..t all valid
                kmer1<0, 0>(Phi, ok, km_lane); kmer1<2, 1>(Phi, ok, km_lane); kmer1<4, 2>(Phi, ok, km_lane); kmer1<6, 3>(Phi, ok, km_lane);
                kmer1<8, 4>(Phi, ok, km_lane); kmer1<10, 5>(Phi, ok, km_lane); kmer1<12, 6>(Phi, ok, km_lane); kmer1<14, 7>(Phi, ok, km_lane);
                kmer1<16, 8>(Phi, ok, km_lane); kmer1<18, 9>(Phi, ok, km_lane); kmer1<20, 10>(Phi, ok, km_lane); kmer1<22, 11>(Phi, ok, km_lane);
                kmer1<8, 12>(Plo, ok, km_lane); kmer1<10, 13>(Plo, ok, km_lane); kmer1<12, 14>(Plo, ok, km_lane); kmer1<14, 15>(Plo, ok, km_lane);
            } else {
                const uint32_t I20 = I27 >> 7;
                const uint32_t bad = I20 | (I20 >> 1) | (I20 >> 2) | (I20 >> 3) | (I20 >> 4);   // bit t: a byte of t-4..t is invalid
                const uint32_t ok = ~bad & kmask;
                kmer1<0, 0>(Phi, ok, km_lane); kmer1<2, 1>(Phi, ok, km_lane); kmer1<4, 2>(Phi, ok, km_lane); kmer1<6, 3>(Phi, ok, km_lane);
                kmer1<8, 4>(Phi, ok, km_lane); kmer1<10, 5>(Phi, ok, km_lane); kmer1<12, 6>(Phi, ok, km_lane); kmer1<14, 7>(Phi, ok, km_lane);
                kmer1<16, 8>(Phi, ok, km_lane); kmer1<18, 9>(Phi, ok, km_lane); kmer1<20, 10>(Phi, ok, km_lane); kmer1<22, 11>(Phi, ok, km_lane);
                kmer1<8, 12>(Plo, ok, km_lane); kmer1<10, 13>(Plo, ok, km_lane); kmer1<12, 14>(Plo, ok, km_lane); kmer1<14, 15>(Plo, ok, km_lane);
            }
        }
        if (n < CS_STAGE) break;     // the rest of the group is shorter still
    }
    __syncthreads();
    // flush: content[b][c] += count ; qual[b][c] += sumq - 33*count
    unsigned long long* content = stats;
    unsigned long long* qualsum = stats + 8 * C;
    for (int i = threadIdx.x; i < 8 * CS_BINW; i += NT) {
        const uint32_t v = packed[i];
        if (v) {
            const int bin = i / CS_BINW, p = i % CS_BINW;
            const int64_t c = t0 - 15 + 16 * (p % CS_ROWW) + p / CS_ROWW;
            const long long cnt = v >> 20, sq = v & 0xFFFFFu;
            atomicAdd(&content[(int64_t)bin * C + c], (unsigned long long)cnt);
            atomicAdd(&qualsum[(int64_t)bin * C + c], (unsigned long long)(sq - 33 * cnt));
        }
    }
    if (DO_KMER) {
        unsigned long long* tail = stats + 16 * C;
        for (int i = threadIdx.x; i < 1024; i += NT) {
            uint32_t sum = 0;
#pragma unroll 8
            for (int l = 0; l < 32; l++) sum += kmer[i * 32 + ((l + threadIdx.x) & 31)];
            if (sum) {
                // table index: five c' pairs (b2, b1); base2val's code is (b1, b2): swap the bits of every pair
                const int code = ((i & 0x155) << 1) | ((i >> 1) & 0x155);
                atomicAdd(&tail[FPL_STATS_KMER + code], (unsigned long long)sum);
                if (kmer_also) atomicAdd(&kmer_also[code], (unsigned long long)sum);
            }
        }
    }
}

namespace {
__global__ void k_cs_keys(const StatSeg* __restrict__ segs, int64_t n, uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = (uint32_t)max(segs[i].len, 0);
    vals[i] = (int32_t)i;
}
__global__ void k_cs_gather(const StatSeg* __restrict__ segs, const uint32_t* __restrict__ keys, const int32_t* __restrict__ vals,
                            int64_t n, SegD* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SegD d;
    d.off = segs[vals[i]].off; d.len = (int32_t)keys[i]; d.pad = 0;
    out[i] = d;
}
}  // namespace

void fpl_cycle_ws_free(CycleWs* ws) {
    cudaFree(ws->tmp); cudaFree(ws->k_in); cudaFree(ws->k_out); cudaFree(ws->v_in); cudaFree(ws->v_out); cudaFree(ws->sorted);
    *ws = CycleWs();
}

// kmer_also: where the 5-mer counts of this launch go besides `stats` (nullptr = nowhere else); do_kmer = false skips
// them.  Returns 0, or -1 when the workspace cannot be allocated.
int launch_cycle_stats(CycleWs* ws, const uint8_t* seq, const uint8_t* qual, const StatSeg* segs, int64_t nseg, int64_t max_len,
                       unsigned long long* stats, int64_t C, bool do_kmer, unsigned long long* kmer_also, bool aligned,
                       cudaStream_t stream) {
    if (nseg == 0 || max_len <= 0) return 0;
    // the opt-in to more than 48 KB of dynamic shared memory is a per-device function attribute
    static std::atomic<unsigned long long> attr_set{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_set.load() & bit)) {
        if (cudaFuncSetAttribute(k_cycle_stats<true, CS_NT_KMER, CS_DEPTH_KMER, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_KMER) != cudaSuccess ||
            cudaFuncSetAttribute(k_cycle_stats<false, CS_NT_PLAIN, CS_DEPTH_PLAIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_PLAIN) != cudaSuccess ||
            cudaFuncSetAttribute(k_cycle_stats<true, CS_NT_KMER, CS_DEPTH_KMER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_KMER) != cudaSuccess ||
            cudaFuncSetAttribute(k_cycle_stats<false, CS_NT_PLAIN, CS_DEPTH_PLAIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_PLAIN) != cudaSuccess)
            return -1;
        attr_set.fetch_or(bit);
    }
    int end_bit = 1;
    while (end_bit < 31 && (max_len >> end_bit)) end_bit++;
    size_t need = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, need, ws->k_in, ws->k_out, ws->v_in, ws->v_out, (int)nseg, 0, end_bit, stream);
    if (nseg > ws->cap || need > ws->tmp_bytes) {
        cudaStreamSynchronize(stream);
        const int64_t cap = nseg > ws->cap ? nseg : ws->cap;
        void* tmp = ws->tmp; size_t tmp_bytes = ws->tmp_bytes;
        if (nseg > ws->cap) {
            cudaFree(ws->k_in); cudaFree(ws->k_out); cudaFree(ws->v_in); cudaFree(ws->v_out); cudaFree(ws->sorted);
            ws->k_in = ws->k_out = nullptr; ws->v_in = ws->v_out = nullptr; ws->sorted = nullptr; ws->cap = 0;
            if (cudaMalloc(&ws->k_in, 4 * cap) != cudaSuccess || cudaMalloc(&ws->k_out, 4 * cap) != cudaSuccess ||
                cudaMalloc(&ws->v_in, 4 * cap) != cudaSuccess || cudaMalloc(&ws->v_out, 4 * cap) != cudaSuccess ||
                cudaMalloc(&ws->sorted, sizeof(SegD) * cap) != cudaSuccess)
                return -1;
            ws->cap = cap;
            // the temporary storage grows with the item count
            cub::DeviceRadixSort::SortPairsDescending(nullptr, need, ws->k_in, ws->k_out, ws->v_in, ws->v_out, (int)cap, 0, 31, stream);
        }
        if (need > tmp_bytes) {
            cudaFree(tmp); ws->tmp = nullptr; ws->tmp_bytes = 0;
            if (cudaMalloc(&ws->tmp, need) != cudaSuccess) return -1;
            ws->tmp_bytes = need;
        }
    }
    const unsigned blocks = (unsigned)((nseg + 255) / 256);
    k_cs_keys<<<blocks, 256, 0, stream>>>(segs, nseg, ws->k_in, ws->v_in);
    size_t bytes = ws->tmp_bytes;
    cub::DeviceRadixSort::SortPairsDescending(ws->tmp, bytes, ws->k_in, ws->k_out, ws->v_in, ws->v_out, (int)nseg, 0, end_bit, stream);
    k_cs_gather<<<blocks, 256, 0, stream>>>(segs, ws->k_out, ws->v_out, nseg, static_cast<SegD*>(ws->sorted));
    dim3 grid((unsigned)((max_len + 15 + CS_TILE - 1) / CS_TILE), (unsigned)((nseg + CS_GROUP - 1) / CS_GROUP));
    const SegD* sorted = static_cast<const SegD*>(ws->sorted);
    static const bool tma = getenv("FPL_CS_TMA") != nullptr && atoi(getenv("FPL_CS_TMA")) != 0;
    if (tma) {
        if (do_kmer) k_cycle_stats<true, CS_NT_KMER, CS_DEPTH_KMER, true><<<grid, CS_NT_KMER, CS_SMEM_KMER, stream>>>(seq, qual, sorted, nseg, stats, C, kmer_also, 1u, aligned ? 1 : 0);
        else k_cycle_stats<false, CS_NT_PLAIN, CS_DEPTH_PLAIN, true><<<grid, CS_NT_PLAIN, CS_SMEM_PLAIN, stream>>>(seq, qual, sorted, nseg, stats, C, nullptr, 1u, aligned ? 1 : 0);
    } else {
        if (do_kmer) k_cycle_stats<true, CS_NT_KMER, CS_DEPTH_KMER, false><<<grid, CS_NT_KMER, CS_SMEM_KMER, stream>>>(seq, qual, sorted, nseg, stats, C, kmer_also, 1u, aligned ? 1 : 0);
        else k_cycle_stats<false, CS_NT_PLAIN, CS_DEPTH_PLAIN, false><<<grid, CS_NT_PLAIN, CS_SMEM_PLAIN, stream>>>(seq, qual, sorted, nseg, stats, C, nullptr, 1u, aligned ? 1 : 0);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// k_kmer_fix: the post-filter 5-mer table is the pre-filter one minus the 5-mers that do not lie inside a passing
// segment (trimmed ends, split gaps, failed or dropped reads).  A 5-mer ending at read position e (e >= 4) survives
// iff a passing segment [a, b) has a + 4 <= e < b (Stats::statRead skips the first four cycles of every read it is
// given, src/stats.cpp:309-311).  The removed ranges are short for almost every read.
// ------------------------------------------------------------------------------------------------------------------
#define KF_WARPS 8
#define KF_LONG 2048          // a removed range longer than this is shared by the block's warps
#define KF_LIST 16

namespace {
// the 5-mers ending at e in [from, to) of one read, strided over `nlanes` lanes (lane id `me`): rem[code]++
__device__ __forceinline__ void kmer_remove_one(const uint8_t* seq, int e, uint32_t rem_base) {
    const uint32_t c4 = kmer_code(seq[e - 4]), c3 = kmer_code(seq[e - 3]), c2 = kmer_code(seq[e - 2]),
                   c1 = kmer_code(seq[e - 1]), c0 = kmer_code(seq[e]);
    if (((c4 | c3 | c2 | c1 | c0) & 8u) == 0u)
        red_shared_add(rem_base + (((c4 << 8) | (c3 << 6) | (c2 << 4) | (c1 << 2) | c0) << 2), 1u);
}
__device__ __forceinline__ void kmer_remove_range(const uint8_t* seq, int from, int to, int me, int nlanes, uint32_t rem_base) {
    if (to - from <= 4 * nlanes) {            // the usual case: a trimmed end or a short gap
        for (int e = from + me; e < to; e += nlanes) kmer_remove_one(seq, e, rem_base);
        return;
    }
    // a long range: four positions per lane in flight (the loads of one round are independent; a lone warp walking a
    // long range is bound by their latency otherwise)
    int e0 = from + me;
    for (; e0 + 3 * nlanes < to; e0 += 4 * nlanes) {
        uint32_t c[4][5];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 5; i++) c[u][i] = (uint32_t)seq[e0 + u * nlanes - 4 + i];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t c4 = kmer_code(c[u][0]), c3 = kmer_code(c[u][1]), c2 = kmer_code(c[u][2]), c1 = kmer_code(c[u][3]),
                           c0 = kmer_code(c[u][4]);
            if (((c4 | c3 | c2 | c1 | c0) & 8u) == 0u)
                red_shared_add(rem_base + (((c4 << 8) | (c3 << 6) | (c2 << 4) | (c1 << 2) | c0) << 2), 1u);
        }
    }
    for (; e0 < to; e0 += nlanes) kmer_remove_one(seq, e0, rem_base);
}
}  // namespace

__global__ void __launch_bounds__(KF_WARPS * 32)
k_kmer_fix(DevBatch b, const fpl_read_result* __restrict__ res, unsigned long long* __restrict__ post_kmer) {
    __shared__ uint32_t rem[1024];
    __shared__ int n_long;
    __shared__ long long long_off[KF_LIST];
    __shared__ int long_from[KF_LIST], long_to[KF_LIST];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) rem[i] = 0;
    if (threadIdx.x == 0) n_long = 0;
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
            if (to - from > KF_LONG) {
                // a failed long read, or the gap between two far-apart adapter hits of an ultra-long read: the block's
                // warps share it after the loop (a lone warp would walk it at the latency of its loads)
                int slot = -1;
                if (lane == 0) slot = atomicAdd(&n_long, 1);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (slot < KF_LIST) {
                    if (lane == 0) { long_off[slot] = b.offsets[r]; long_from[slot] = from; long_to[slot] = to; }
                } else {
                    kmer_remove_range(seq, from, to, lane, 32, rem_base);
                }
            } else {
                kmer_remove_range(seq, from, to, lane, 32, rem_base);
            }
            // the next removed range starts behind this one and behind the segment; a passing segment shorter than
            // four bases (only possible with --length_required < 5) ends in front of `to` and must not pull it back
            if (k < nk) from = max(from, max(to, ke[k]));
        }
    }
    __syncthreads();
    const int nl = min(n_long, KF_LIST);
    for (int i = 0; i < nl; i++)
        kmer_remove_range(b.seq + long_off[i], long_from[i], long_to[i], threadIdx.x, KF_WARPS * 32, rem_base);
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
__device__ __forceinline__ void hist_vec(uint32_t hbase, const uint4& v, uint32_t delta) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        // bin address = hbase + 4 * (byte & 127): one dp4a per byte (weight 4 on that byte) on the FMA pipe, after one
        // LOP3 per word that keeps the bytes inside the 128-bin table
        const uint32_t m = w[k] & 0x7f7f7f7fu;
#pragma unroll
        for (int j = 0; j < 4; j++) red_shared_add(__dp4a(m, 4u << (8 * j), hbase), delta);
    }
}
// h[q] += delta for the bytes qp[0..n) (delta = 1 or 0xFFFFFFFF): 16-byte vector body, byte head/tail; the work is
// strided over `nth` threads, this one being `me` (a warp: lane / 32; a whole block: threadIdx.x / blockDim.x)
__device__ __forceinline__ void hist_bytes(uint32_t hbase, const uint8_t* qp, int n, uint32_t delta, int me, int nth) {
    const int head = min(n, (int)((16 - (reinterpret_cast<uintptr_t>(qp) & 15)) & 15));
    if (me < head) red_shared_add(hbase + ((uint32_t)(qp[me] & 127) << 2), delta);
    const int nvec = (n - head) >> 4;
    const uint4* vp = reinterpret_cast<const uint4*>(qp + head);
    // four vectors per thread in flight: a warp alone keeps 2 KB of loads outstanding (the kernel is latency-bound otherwise)
    int i = me;
    for (; i + 3 * nth < nvec; i += 4 * nth) {
        const uint4 v0 = __ldg(vp + i), v1 = __ldg(vp + i + nth), v2 = __ldg(vp + i + 2 * nth), v3 = __ldg(vp + i + 3 * nth);
        hist_vec(hbase, v0, delta); hist_vec(hbase, v1, delta); hist_vec(hbase, v2, delta); hist_vec(hbase, v3, delta);
    }
    for (; i < nvec; i += nth) hist_vec(hbase, __ldg(vp + i), delta);
    const int done = head + (nvec << 4);
    if (me < 16 && done + me < n) red_shared_add(hbase + ((uint32_t)(qp[done + me] & 127) << 2), delta);   // < 16 tail bytes
}

// One read: the histogram of the whole read gives the pre-filter median and mBaseQualHistogram; the histogram of each
// passing segment is derived from it by subtracting the (short) removed ends — or counted directly when the segment is
// the smaller part — and gives the post-filter median and histogram.  One pass over the quality bytes serves both Stats
// objects.  COOP = false: one warp owns the read.  COOP = true: the whole block does (an ultra-long read streamed by
// one warp is bound by the latency of that warp's loads); every thread of the block calls it for the same read, the
// histograms are warp 0's, warp 0 does the medians.
template <bool COOP>
__device__ __forceinline__ void read_qual_one(const DevBatch& b, int64_t r, uint32_t* hfull, uint32_t* hseg, uint32_t (*block_hist)[128],
                                              unsigned long long (*block_misc)[2], unsigned long long* const* tail,
                                              fpl_read_result* __restrict__ res, bool pre_only) {
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    const int me = COOP ? (int)threadIdx.x : lane, nth = COOP ? (int)blockDim.x : 32;
    auto sync = [] { if (COOP) __syncthreads(); else __syncwarp(); };
    const bool lead = !COOP || wid == 0;                 // the warp that owns the medians and the flushes
    const uint32_t hfull_s = shared_addr(hfull), hseg_s = shared_addr(hseg);
    const uint8_t* qp = b.qual + b.offsets[r];
    const int L = b.lens[r];
    fpl_read_result* o = &res[r];
    sync();
    for (int i = me; i < 128; i += nth) hfull[i] = 0;
    sync();
    hist_bytes(hfull_s, qp, L, 1u, me, nth);
    sync();
    uint32_t own[4];
    if (lead) {
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
    }
    const int nseg = pre_only ? 0 : o->n_segments;   // --mask/--break: the post-filter part is k_ext_seg_qual's
    for (int k = 0; k < nseg; k++) {
        if (o->seg_result[k] != FPL_PASS_FILTER) continue;      // uniform
        const int a = o->seg_lo[k], n = o->seg_len[k];
        sync();
        if (L - n <= n) {            // copy the read's histogram and take the removed ends out
            for (int i = me; i < 128; i += nth) hseg[i] = hfull[i];
            sync();
            hist_bytes(hseg_s, qp, a, 0xFFFFFFFFu, me, nth);
            hist_bytes(hseg_s, qp + a + n, L - a - n, 0xFFFFFFFFu, me, nth);
        } else {
            for (int i = me; i < 128; i += nth) hseg[i] = 0;
            sync();
            hist_bytes(hseg_s, qp + a, n, 1u, me, nth);
        }
        sync();
        if (lead) {
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
    }
    sync();
}
}  // namespace

#define RQ_LONG (96 * 1024)      // reads longer than this are streamed by the whole block
#define RQ_LIST 24

// Shared-memory atomics are the fast way to histogram on this part (tools/ubench_hist.cu: a 128-bin table updated
// with atomicAdd by 8 warps streams quality bytes at HBM speed, 4x faster than lane-private read-modify-write).
__global__ void __launch_bounds__(RQ_WARPS * 32)
k_read_qual(DevBatch b, unsigned long long* __restrict__ stats_pre, unsigned long long* __restrict__ stats_post, int64_t C,
            fpl_read_result* __restrict__ res, bool pre_only) {
    __shared__ uint32_t hist[RQ_WARPS][2][128];        // per warp: [0] whole read, [1] current segment
    __shared__ uint32_t block_hist[2][128];            // this block's reads -> one flush per Stats block
    __shared__ unsigned long long block_misc[2][2];    // reads, length sum
    __shared__ int n_long;
    __shared__ long long long_read[RQ_LIST];
    const int wid = threadIdx.x >> 5, lane = lane_id();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) (&block_hist[0][0])[i] = 0;
    if (threadIdx.x < 4) (&block_misc[0][0])[threadIdx.x] = 0;
    if (threadIdx.x == 0) n_long = 0;
    __syncthreads();
    unsigned long long* tail[2] = {stats_pre + 16 * C, stats_post + 16 * C};
    const int64_t nwarps = (int64_t)gridDim.x * RQ_WARPS;
    for (int64_t r = (int64_t)blockIdx.x * RQ_WARPS + wid; r < b.n_reads; r += nwarps) {
        if (b.lens[r] > RQ_LONG) {
            int slot = -1;
            if (lane == 0) slot = atomicAdd(&n_long, 1);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot < RQ_LIST) { if (lane == 0) long_read[slot] = r; continue; }
        }
        read_qual_one<false>(b, r, hist[wid][0], hist[wid][1], block_hist, block_misc, tail, res, pre_only);
    }
    __syncthreads();
    const int nl = min(n_long, RQ_LIST);
    for (int i = 0; i < nl; i++)
        read_qual_one<true>(b, long_read[i], hist[0][0], hist[0][1], block_hist, block_misc, tail, res, pre_only);
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
