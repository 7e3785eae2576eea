#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fplgpu.h"

// Device half of SURVEY §8f row 2 (output assembly): state of fpl_emit_fastq_host.
struct FplEmit {
    int64_t* d_size = nullptr; int64_t cap_size = 0;   // [0, n]: --out bytes of read r; [n+1, 2n+1]: --failed_out bytes
    int64_t* d_off = nullptr; int64_t cap_off = 0;     // exclusive sums of the two halves
    uint8_t* d_out = nullptr; int64_t cap_out = 0;
    uint8_t* d_failed = nullptr; int64_t cap_failed = 0;
    void* d_tmp = nullptr; int64_t cap_tmp = 0;
    int64_t out_bytes = 0, failed_bytes = 0;
    bool built = false, with_failed = false;
};

// What the text is cut from: the chunk, its record table, the per-read results, and (--mask / --break) the list of
// output reads with its per-read ranges and the masked copy of the packed sequence buffer.
struct EmitSource {
    const uint8_t* text;
    const fpl_fastq_record* rec;
    const fpl_read_result* res;
    int64_t n_reads;
    const fpl_segment* segs;        // null in the plain mode (the records' inline segments are the output reads)
    const int32_t* seg_off;         // segs of read r: [seg_off[2r], seg_off[2r+2])
    const uint8_t* mseq;            // masked bases in the packed layout (null: nothing was masked)
    const int64_t* offsets;         // packed-layout slot of read r
};

int fpl_emit_build(FplEmit* e, const EmitSource& src, bool want_failed, cudaStream_t s, char* err, size_t errlen);
void fpl_emit_free(FplEmit* e);
