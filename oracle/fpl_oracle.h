/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of the reference algorithm of fastplong's per-read
 * hot loop.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product (libfplgpu.so) never does.
 *
 * Parity is PINNED: this restatement is checked (tests/test_oracle_vs_reference.py) against
 *   - every known-answer vector of the reference's own unit tests (test/adaptertrimmer_test.cpp:4-57,
 *     test/filter_test.cpp:4-22, test/polyx_test.cpp:4-17, src/editdistance.cpp:141-172), and
 *   - the unmodified reference itself, compiled into oracle/_ref/libfplref.so, field by field on
 *     adversarial seeded batches, and whole-binary fastplong_ref runs (golden fixtures in tests/golden/).
 */
#ifndef FPL_ORACLE_H
#define FPL_ORACLE_H
#include "fplgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

orc_ctx* orc_create(const fpl_options* opt, const fpl_adapters* adapters);
void orc_destroy(orc_ctx* c);
/* processSingleEnd's per-read body over a packed HOST batch; accumulates pre/post stats + counters. */
int orc_process(orc_ctx* c, const fpl_batch* batch, fpl_read_result* results);
int64_t orc_stats_cycles(orc_ctx* c);
int orc_stats_download(orc_ctx* c, int which, int64_t* out, int64_t C);
int orc_counters_download(orc_ctx* c, int64_t* out, int64_t n_words);
/* --mask/--break: output reads / masked regions of the last orc_process call (returns -1 if cap is too small) */
int orc_last_segments(orc_ctx* c, fpl_segment* out, int64_t cap, int64_t* n);
int orc_last_mask_regions(orc_ctx* c, fpl_region* out, int64_t cap, int64_t* n);

/* single operators, for known-answer tests */
int orc_edit_distance(const char* a, int alen, const char* b, int blen);
int orc_search_adapter(const orc_ctx* c, const char* read, int rlen, const char* adapter, int alen,
                       int search_start, int search_len, int left, int right);
/* returns 0 and the kept window, or 1 if the read is dropped */
int orc_trim_and_cut(const orc_ctx* c, const char* seq, const char* qual, int len, int* lo, int* rlen);
/* returns new length; base and plen describe the event (base = -1 if none) */
int orc_trim_polyx(const char* seq, int len, int min_len, int* base, int* plen);

#ifdef __cplusplus
}
#endif
#endif
