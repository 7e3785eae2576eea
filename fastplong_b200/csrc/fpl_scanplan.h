// Host-built description of the two middle-adapter scans for k_scan_fast (passed by value as a kernel parameter).
#pragma once
#include <stdint.h>

#define SCANPLAN_MASK_WORDS 40   // must equal SF_MASK_WORDS in fpl_scan_fast.cu

struct ScanPlan {
    int fast;           // 1: both scans can use the bit-sliced kernel (adapters ACGT-only, 1..128 bp) or no scan at all
    int npl;            // bit planes of the match counter: 5 (alen <= 31) .. 8 (alen == 128)
    int halo_words;     // 32-position words each lane needs beyond its own: ((max alen - 1) >> 5) + 1
    int n_in[2];        // CSA inputs per adapter, padded to a multiple of 8
    // per input: 0x8000 | (shared-memory word offset = letter * SCANPLAN_MASK_WORDS + (i >> 5)) << 5 | (i & 31); 0 = padding
    uint16_t in[2][128];
};
