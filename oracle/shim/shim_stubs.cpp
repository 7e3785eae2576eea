// Test infrastructure only (oracle build): link-time stand-ins for isa-l (zlib-backed inflate)
// and libdeflate (zlib-backed gzip compress) so that the unmodified reference sources link.
// API shapes follow src/fastqreader.cpp:79-175 (call sites) and the vendored src/libdeflate.h.
#include <zlib.h>
#include <stdlib.h>
#include <string.h>
#include "isa-l/igzip_lib.h"
#include "libdeflate.h"

extern "C" {

void isal_gzip_header_init(struct isal_gzip_header* h) { h->unused = 0; }

static void start_stream(struct inflate_state* s) {
    z_stream* z = (z_stream*)calloc(1, sizeof(z_stream));
    inflateInit2(z, 15 + 16);  // gzip wrapper; header parsed lazily by inflate()
    s->zs = z;
}
void isal_inflate_init(struct inflate_state* s) {
    memset(s, 0, sizeof(*s));
    start_stream(s);
}
void isal_inflate_reset(struct inflate_state* s) {
    z_stream* z = (z_stream*)s->zs;
    if (z) inflateReset2(z, 15 + 16);
    s->block_state = ISAL_BLOCK_BUSY;
    s->bfinal = 0;
    s->avail_in = 0;
    s->next_in = NULL;
}
int isal_read_gzip_header(struct inflate_state* s, struct isal_gzip_header*) {
    if (s->avail_in >= 2 && !(s->next_in[0] == 0x1f && s->next_in[1] == 0x8b)) return -1;
    return ISAL_DECOMP_OK;
}
int isal_inflate(struct inflate_state* s) {
    z_stream* z = (z_stream*)s->zs;
    z->next_in = s->next_in; z->avail_in = s->avail_in;
    z->next_out = s->next_out; z->avail_out = s->avail_out;
    int ret = inflate(z, Z_NO_FLUSH);
    s->next_in = z->next_in; s->avail_in = z->avail_in;
    s->next_out = z->next_out; s->avail_out = z->avail_out;
    if (ret == Z_STREAM_END) { s->block_state = ISAL_BLOCK_FINISH; s->bfinal = 1; return ISAL_DECOMP_OK; }
    if (ret == Z_OK || ret == Z_BUF_ERROR) return ISAL_DECOMP_OK;
    return -1;
}

struct libdeflate_compressor { int level; };

struct libdeflate_compressor* libdeflate_alloc_compressor(int level) {
    libdeflate_compressor* c = (libdeflate_compressor*)malloc(sizeof(*c));
    c->level = level > 9 ? 9 : level;
    return c;
}
size_t libdeflate_gzip_compress_bound(struct libdeflate_compressor*, size_t n) {
    return compressBound(n) + 32;
}
size_t libdeflate_gzip_compress(struct libdeflate_compressor* c, const void* in, size_t n, void* out, size_t cap) {
    z_stream z; memset(&z, 0, sizeof(z));
    if (deflateInit2(&z, c->level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return 0;
    z.next_in = (Bytef*)in; z.avail_in = n; z.next_out = (Bytef*)out; z.avail_out = cap;
    int ret = deflate(&z, Z_FINISH);
    size_t produced = cap - z.avail_out;
    deflateEnd(&z);
    return ret == Z_STREAM_END ? produced : 0;
}
void libdeflate_free_compressor(struct libdeflate_compressor* c) { free(c); }

}  // extern "C"
