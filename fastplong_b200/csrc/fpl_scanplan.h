// Host-built description of the two middle-adapter scans for k_scan_fast (passed by value as a kernel parameter,
// i.e. it lives in the constant bank: statically indexed entries become direct c[0][imm] instruction operands).
#pragma once
#include <stdint.h>

struct ScanPlan {
    int fast;           // 1: both scans can use the bit-sliced kernel (adapters ACGT-only, 1..128 bp) or no scan at all
    int npl;            // bit planes of the match counter: 5 (alen <= 31) .. 8 (alen == 128)
    int halo_words;     // 32-position words each lane needs beyond its own: ((max alen - 1) >> 5) + 1
    int pad;
    // 2-bit code of adapter k's letter i, expanded to all-zeros / all-ones words: hm = code bit "HI" (byte bit 1),
    // lm = code bit "LO" (byte bit 2).  A=(0,0) C=(1,0) G=(1,1) T=(0,1).
    uint32_t hm[2][128];
    uint32_t lm[2][128];
    uint32_t vm[2][128];   // all-ones for i < alen, zero beyond (only read in the last, partial block of 8)
};
