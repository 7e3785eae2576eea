// Test infrastructure only (oracle build): the slice of the isa-l inflate API that
// src/fastqreader.{h,cpp} compiles against, backed by the system zlib so .gz inputs still work.
// Off the hot path: only the decompressed byte stream matters.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISAL_DECOMP_OK 0
#define ISAL_BLOCK_FINISH 1
#define ISAL_BLOCK_BUSY 0
#define ISAL_GZIP_NO_HDR_VER 5

struct inflate_state {
    uint8_t* next_out;
    uint32_t avail_out;
    uint8_t* next_in;
    uint32_t avail_in;
    int block_state;
    int bfinal;
    int crc_flag;
    void* zs;  // z_stream owned by the shim
};

struct isal_gzip_header { int unused; };

void isal_gzip_header_init(struct isal_gzip_header* h);
void isal_inflate_init(struct inflate_state* s);
void isal_inflate_reset(struct inflate_state* s);
int isal_read_gzip_header(struct inflate_state* s, struct isal_gzip_header* h);
int isal_inflate(struct inflate_state* s);

#ifdef __cplusplus
}
#endif
