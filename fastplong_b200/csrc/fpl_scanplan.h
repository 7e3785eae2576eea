// Host-built description of the two middle-adapter scans for k_scan_fast (passed by value as a kernel parameter).
#pragma once
#include <stdint.h>

struct ScanPlan {
    int fast;           // 1: both scans can use the bit-sliced kernel (adapters ACGT-only, 1..128 bp) or no scan at all
    int npl;            // bit planes of the match counter: 5 (alen <= 31) .. 8 (alen == 128)
    int halo_words;     // 32-position words each lane needs beyond its own: ((max alen - 1) >> 5) + 1
    // For adapter k, letter l (A,C,G,T), word w: the adapter positions i with a_i == l and (i >> 5) == w, as shift
    // amounts (i & 31); cnt = how many.  The kernel sums the letter-l mask shifted by every listed amount.
    uint8_t cnt[2][4][4];
    uint8_t shift[2][4][4][32];
};
