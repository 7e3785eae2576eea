// Test infrastructure only — never linked into the product.
//
// Function-level harness around the UNMODIFIED reference translation units (compiled where they lie under
// /root/reference/src by oracle/Makefile into oracle/_ref/libfplref.so).  It feeds packed batches through the
// reference's own operators in the order of SingleEndProcessor::processSingleEnd
// (src/seprocessor.cpp:180-329) and reports what they decided in the C-ABI's record types
// (include/fplgpu.h), so tests can compare reference vs oracle restatement vs CUDA field by field.
// The order of operations below is the only logic of ours; every decision is made by reference code
// (Stats::statRead, Filter::trimAndCut, PolyX::trimPolyX, AdapterTrimmer::*, Read::breakByGap,
// Filter::passFilter).  Whole-binary runs of oracle/_ref/fastplong_ref cross-check this driver.
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>
// Stats and FilterResult keep their accumulators private (src/stats.h:56-113, src/filterresult.h:58-64);
// the harness only READS them.
#define private public
#include "stats.h"
#include "filterresult.h"
#undef private
#include "filter.h"
#include "adaptertrimmer.h"
#include "polyx.h"
#include "editdistance.h"
#include "options.h"
#include "read.h"
#include "util.h"
#include "fplgpu.h"

std::string command;
std::mutex logmtx;

namespace {

struct Harness {
    Options opt;
    Filter* filter;
    Stats* pre;
    Stats* post;
    FilterResult* fr;
    std::vector<std::string> adapters;  // fpl_adapters order
    int64_t dropped = 0, split = 0;
    std::map<uint32_t, int64_t> events;  // packed event -> count
    std::vector<fpl_segment> segs;       // --mask/--break: outReads of the last ref_process call
    std::vector<fpl_region> regs;        // --mask: regions handed to Read::maskRegionWithN (clamped like it does)
};

int medianOfLast(Stats* s, const long* before) {
    for (int c = 0; c < 128; c++)
        if (s->mMedianReadQualHistogram[c] != before[c]) return c;
    return 0;
}

// One adapter trim through the reference; records the event it counted (if any).
int trimOne(Harness* h, Read* r, int idx, int side, std::vector<uint32_t>& ev) {
    FilterResult probe(&h->opt);
    std::string& a = h->adapters[idx];
    int n = side == 0
        ? AdapterTrimmer::trimBySequenceStart(r, &probe, a, h->opt.adapter.edMax, h->opt.adapter.trimmingExtension)
        : AdapterTrimmer::trimBySequenceEnd(r, &probe, a, h->opt.adapter.edMax, h->opt.adapter.trimmingExtension);
    for (auto& kv : probe.mAdapter) {  // at most one entry: the exact (sub)string the reference counted
        const std::string& s = kv.first;
        int c = (int)s.length();
        std::string expect = side == 0 ? a.substr(a.length() - c, c) : a.substr(0, c);
        if (s != expect) { std::cerr << "ref_harness: unexpected adapter event string" << std::endl; abort(); }
        uint32_t e = FPL_EVENT(idx, side, c);
        ev.push_back(e);
        h->events[e] += kv.second;
        h->fr->addAdapterTrimmed(s);
    }
    return n;
}

}  // namespace

extern "C" {

void* ref_create(const fpl_options* o, const fpl_adapters* ad) {
    Harness* h = new Harness();
    Options& opt = h->opt;
    opt.seqLen = 0;
    opt.trim.front = o->trim_front;
    opt.trim.tail = o->trim_tail;
    opt.qualityCut.enabledFront = o->cut_front_enabled;
    opt.qualityCut.windowSizeFront = o->cut_front_window;
    opt.qualityCut.qualityFront = o->cut_front_quality;
    opt.qualityCut.enabledTail = o->cut_tail_enabled;
    opt.qualityCut.windowSizeTail = o->cut_tail_window;
    opt.qualityCut.qualityTail = o->cut_tail_quality;
    opt.polyXTrim.enabled = o->polyx_enabled;
    opt.polyXTrim.minLen = o->polyx_min_len;
    opt.adapter.enabled = o->adapter_enabled;
    opt.adapter.trimmingExtension = o->trimming_extension;
    opt.adapter.edMax = o->ed_max;
    opt.adapter.sequenceStart = ad->start ? ad->start : "";
    opt.adapter.sequenceEnd = ad->end ? ad->end : "";
    opt.adapter.hasFasta = ad->n_fasta > 0;
    for (int i = 0; i < ad->n_fasta; i++) opt.adapter.seqsInFasta.push_back(ad->fasta[i]);
    opt.qualfilter.enabled = o->qual_filter_enabled;
    opt.qualfilter.qualifiedQual = (char)o->qualified_qual;
    opt.qualfilter.unqualifiedPercentLimit = o->unqualified_percent_limit;
    opt.qualfilter.avgQualReq = o->avg_qual_req;
    opt.qualfilter.nBasePercentLimit = o->n_base_percent_limit;
    opt.qualfilter.nBaseLimit = o->n_base_limit;
    opt.lengthFilter.enabled = o->length_filter_enabled;
    opt.lengthFilter.requiredLength = o->length_required;
    opt.lengthFilter.maxLength = o->length_max;
    opt.complexityFilter.enabled = o->complexity_enabled;
    opt.complexityFilter.threshold = o->complexity_threshold_pct / 100.0;  // src/main.cpp:205
    opt.mask.enabled = o->mask_enabled;
    opt.mask.windowSize = o->mask_window;
    opt.mask.quality = o->mask_quality;
    opt.breakOpt.enabled = o->break_enabled;
    opt.breakOpt.windowSize = o->break_window;
    opt.breakOpt.quality = o->break_quality;
    h->adapters.push_back(opt.adapter.sequenceStart);
    h->adapters.push_back(opt.adapter.sequenceEnd);
    for (auto& s : opt.adapter.seqsInFasta) h->adapters.push_back(s);
    h->filter = new Filter(&h->opt);
    h->pre = new Stats(&h->opt);
    h->post = new Stats(&h->opt);
    h->fr = new FilterResult(&h->opt);
    return h;
}

void ref_destroy(void* hh) {
    Harness* h = (Harness*)hh;
    delete h->filter; delete h->pre; delete h->post; delete h->fr;
    delete h;
}

// processSingleEnd's per-read body (src/seprocessor.cpp:186-295) for n reads of a packed host batch.
int ref_process(void* hh, const fpl_batch* b, fpl_read_result* results) {
    Harness* h = (Harness*)hh;
    Options* opt = &h->opt;
    long before[128];
    const bool ext = opt->mask.enabled || opt->breakOpt.enabled;
    h->segs.clear(); h->regs.clear();
    for (int64_t i = 0; i < b->n_reads; i++) {
        fpl_read_result* out = &results[i];
        memset(out, 0, sizeof(*out));
        const int L = b->lens[i];
        std::string seq((const char*)b->seq + b->offsets[i], L);
        std::string qual((const char*)b->qual + b->offsets[i], L);
        Read* or1 = new Read("@r", seq.c_str(), "+", qual.c_str());
        // the Read(const char*...) constructor stops at NUL; rebuild exactly for arbitrary bytes
        *or1->mSeq = seq; *or1->mQuality = qual;

        memcpy(before, h->pre->mMedianReadQualHistogram, sizeof(before));
        h->pre->statRead(or1);                                             // :192
        out->pre_median_qual = (uint8_t)medianOfLast(h->pre, before);

        int frontTrimmed = 0;
        Read* r1 = h->filter->trimAndCut(or1, opt->trim.front, opt->trim.tail, frontTrimmed);  // :196
        // window bookkeeping: r1 is or1 mutated in place; lo = bases erased at the front so far
        int lo = frontTrimmed;
        if (r1 == NULL) { out->flags |= FPL_FLAG_DROPPED_BY_CUT; h->dropped++; }

        if (r1 != NULL && opt->polyXTrim.enabled) {                        // :198-201
            long rb[4], bb[4];
            memcpy(rb, h->fr->mTrimmedPolyXReads, sizeof(rb));
            memcpy(bb, h->fr->mTrimmedPolyXBases, sizeof(bb));
            PolyX::trimPolyX(r1, h->fr, opt->polyXTrim.minLen);
            for (int k = 0; k < 4; k++)
                if (h->fr->mTrimmedPolyXReads[k] != rb[k]) {
                    out->flags |= FPL_FLAG_POLYX;
                    out->polyx_base = (uint8_t)k;
                    out->polyx_len = (int32_t)(h->fr->mTrimmedPolyXBases[k] - bb[k]);
                }
        }

        std::vector<Read*> outReads;
        std::vector<uint32_t> ev;
        int segLo[2] = {0, 0};
        if (r1 != NULL && opt->adapter.enabled) {                          // :205
            int trimmed = 0;
            if (!opt->adapter.sequenceStart.empty()) {
                int before_len = r1->length();
                trimmed += trimOne(h, r1, 0, 0, ev);                       // :207-208
                lo += before_len - r1->length();
            }
            if (!opt->adapter.sequenceEnd.empty())
                trimmed += trimOne(h, r1, 1, 1, ev);                       // :209-210
            if (opt->adapter.hasFasta) {                                   // :211-213, adaptertrimmer.cpp:42-57
                for (size_t k = 0; k < opt->adapter.seqsInFasta.size(); k++) {
                    int before_len = r1->length();
                    trimmed += trimOne(h, r1, 2 + (int)k, 0, ev);
                    lo += before_len - r1->length();
                    trimmed += trimOne(h, r1, 2 + (int)k, 1, ev);
                }
            }
            if (trimmed > 0) h->fr->addReadTrimmed(trimmed);               // :214-216
            out->adapter_trimmed_bases = trimmed;

            int start = -1, len = 0;
            bool found = AdapterTrimmer::findMiddleAdapters(r1, opt->adapter.sequenceStart, opt->adapter.sequenceEnd,
                                                            start, len, opt->adapter.edMax, opt->adapter.trimmingExtension);  // :219-221
            if (found) {
                out->flags |= FPL_FLAG_MIDDLE_ADAPTER;
                h->split++;
                outReads = r1->breakByGap(start, len);                     // :224
                // recover the segment windows from the reference's own output reads
                int l1 = start, l2 = r1->length() - start - len;
                size_t k = 0;
                if (l1 > 0) { segLo[k++] = lo; }
                if (l2 > 0) { segLo[k++] = lo + start + len; }
                if (k != outReads.size()) { std::cerr << "ref_harness: breakByGap bookkeeping" << std::endl; abort(); }
                if (outReads.size() == 1 && l1 <= 0) out->flags |= FPL_FLAG_SEG0_IS_RIGHT;
            } else {
                outReads.push_back(r1);
                segLo[0] = lo;
            }
        } else if (r1 != NULL) {
            outReads.push_back(r1);
            segLo[0] = lo;
        }
        if (r1 != NULL) { out->trim_lo = lo; out->trim_len = r1->length(); }
        out->n_events = (uint16_t)ev.size();
        for (size_t k = 0; k < ev.size() && k < FPL_INLINE_EVENTS; k++) out->events[k] = ev[k];

        if (ext) {
            // ---- src/seprocessor.cpp:235-288 with the reference's own detectLowQualityRegions / breakByRegions /
            //      maskRegionWithN; our bookkeeping only tracks where each output read sits in the original bytes ----
            struct Piece { Read* r; int lo; int side; int is_r1; int bidx; };
            std::vector<Piece> reads;
            const bool wasSplit = (out->flags & FPL_FLAG_MIDDLE_ADAPTER) != 0;
            for (size_t k = 0; k < outReads.size(); k++) {
                int side = wasSplit ? ((k == 1 || (out->flags & FPL_FLAG_SEG0_IS_RIGHT)) ? 2 : 1) : 0;
                reads.push_back({outReads[k], segLo[k], side, outReads[k] == r1 ? 1 : 0, 0});
            }
            if (opt->breakOpt.enabled && !reads.empty()) {
                std::vector<Piece> tmp;
                for (auto& p : reads) {
                    vector<pair<int, int>> regions = h->filter->detectLowQualityRegions(p.r, opt->breakOpt.windowSize, opt->breakOpt.quality);
                    if (!regions.empty()) {
                        vector<Read*> brs = p.r->breakByRegions(regions);
                        // offsets of the pieces: the arithmetic of Read::breakByRegions (src/read.cpp:227-262), checked
                        // against the bytes of the pieces the reference actually made
                        int lastEnd = -1, length = p.r->length();
                        size_t bi = 0;
                        for (size_t ri = 0; ri <= regions.size(); ri++) {
                            int start, end;
                            if (ri < regions.size()) {
                                start = regions[ri].first < 0 ? 0 : regions[ri].first;
                                end = regions[ri].second >= length ? length - 1 : regions[ri].second;
                                if (start > end || start >= length) continue;
                                if (start > lastEnd + 1) {
                                    if (bi >= brs.size()) { std::cerr << "ref_harness: breakByRegions bookkeeping" << std::endl; abort(); }
                                    tmp.push_back({brs[bi++], p.lo + lastEnd + 1, p.side, 0, (int)ri + 1});
                                }
                                lastEnd = end;
                            } else if (lastEnd < length - 1) {
                                if (bi >= brs.size()) { std::cerr << "ref_harness: breakByRegions bookkeeping" << std::endl; abort(); }
                                tmp.push_back({brs[bi++], p.lo + lastEnd + 1, p.side, 0, (int)regions.size() + 1});
                            }
                        }
                        if (bi != brs.size()) { std::cerr << "ref_harness: breakByRegions piece count" << std::endl; abort(); }
                        if (p.r != or1 && p.r != r1) delete p.r;
                    } else {
                        tmp.push_back(p);
                    }
                }
                reads = tmp;
            }
            if (opt->mask.enabled && !reads.empty()) {
                for (auto& p : reads) {
                    vector<pair<int, int>> regions = h->filter->detectLowQualityRegions(p.r, opt->mask.windowSize, opt->mask.quality);
                    for (auto& rg : regions) {
                        int start = rg.first, len = rg.second - rg.first + 1, length = p.r->length();
                        p.r->maskRegionWithN(start, len);
                        if (start < 0 || len <= 0 || start >= length) continue;      // what maskRegionWithN itself skips
                        if (start + len > length) len = length - start;
                        h->regs.push_back({(int32_t)i, p.lo + start, len});
                    }
                }
            }
            out->n_segments = (int32_t)reads.size();
            for (size_t k = 0; k < reads.size(); k++) {
                Piece& p = reads[k];
                int result = h->filter->passFilter(p.r);
                h->fr->addFilterResult(result, 1);
                fpl_segment sg;
                memset(&sg, 0, sizeof(sg));
                sg.read = (int32_t)i; sg.lo = p.lo; sg.len = p.r->length(); sg.result = (uint8_t)result;
                sg.split_side = (uint8_t)p.side; sg.is_r1 = (uint8_t)p.is_r1; sg.break_index = p.bidx;
                if (result == PASS_FILTER) {
                    memcpy(before, h->post->mMedianReadQualHistogram, sizeof(before));
                    h->post->statRead(p.r);
                    sg.median_qual = (uint8_t)medianOfLast(h->post, before);
                }
                // the piece's QUALITY bytes are never modified: they must equal the original window
                if (sg.len > 0 && memcmp(p.r->mQuality->data(), b->qual + b->offsets[i] + p.lo, sg.len) != 0) {
                    std::cerr << "ref_harness: piece window mismatch on read " << i << std::endl; abort();
                }
                h->segs.push_back(sg);
                if (k < 2) {
                    out->seg_lo[k] = sg.lo; out->seg_len[k] = sg.len; out->seg_result[k] = sg.result;
                    out->seg_median_qual[k] = sg.median_qual;
                }
                if (p.r != or1 && p.r != r1) delete p.r;
            }
            delete or1;
            continue;
        }

        out->n_segments = (int32_t)outReads.size();
        for (size_t k = 0; k < outReads.size(); k++) {                     // :264-288
            Read* outr = outReads[k];
            int result = h->filter->passFilter(outr);
            h->fr->addFilterResult(result, 1);
            out->seg_lo[k] = segLo[k];
            out->seg_len[k] = outr->length();
            out->seg_result[k] = (uint8_t)result;
            if (result == PASS_FILTER) {
                memcpy(before, h->post->mMedianReadQualHistogram, sizeof(before));
                h->post->statRead(outr);
                out->seg_median_qual[k] = (uint8_t)medianOfLast(h->post, before);
            }
            // consistency: the segment bytes must be the window of the original bytes we report
            if (outr->length() > 0 &&
                memcmp(outr->mSeq->data(), b->seq + b->offsets[i] + segLo[k], outr->length()) != 0) {
                std::cerr << "ref_harness: window bookkeeping mismatch on read " << i << std::endl; abort();
            }
            if (outr != or1 && outr != r1) delete outr;
        }
        delete or1;
    }
    return 0;
}

int ref_last_segments(void* hh, fpl_segment* out, int64_t cap, int64_t* n) {
    Harness* h = (Harness*)hh;
    *n = (int64_t)h->segs.size();
    if (*n > cap) return -1;
    if (*n) memcpy(out, h->segs.data(), sizeof(fpl_segment) * h->segs.size());
    return 0;
}
int ref_last_mask_regions(void* hh, fpl_region* out, int64_t cap, int64_t* n) {
    Harness* h = (Harness*)hh;
    *n = (int64_t)h->regs.size();
    if (*n > cap) return -1;
    if (*n) memcpy(out, h->regs.data(), sizeof(fpl_region) * h->regs.size());
    return 0;
}

int64_t ref_stats_cycles(void* hh, int which) {
    Harness* h = (Harness*)hh;
    return (which == FPL_STATS_PRE ? h->pre : h->post)->mBufLen;
}

// Fill the FPL_STATS_WORDS(C) layout from the reference Stats object (C must be >= every cycle touched).
int ref_stats_download(void* hh, int which, int64_t* out, int64_t C) {
    Harness* h = (Harness*)hh;
    Stats* s = which == FPL_STATS_PRE ? h->pre : h->post;
    memset(out, 0, sizeof(int64_t) * FPL_STATS_WORDS(C));
    int n = s->mBufLen < C ? s->mBufLen : (int)C;
    for (int b = 0; b < 8; b++)
        for (int c = 0; c < n; c++) {
            out[(int64_t)b * C + c] = s->mCycleBaseContents[b][c];
            out[8 * C + (int64_t)b * C + c] = s->mCycleBaseQual[b][c];
        }
    for (int c = n; c < s->mBufLen; c++)
        for (int b = 0; b < 8; b++)
            if (s->mCycleBaseContents[b][c] != 0) return -1;  // C too small
    int64_t* t = out + 16 * C;
    for (int k = 0; k < 1024; k++) t[FPL_STATS_KMER + k] = s->mKmer[k];
    for (int k = 0; k < 128; k++) {
        t[FPL_STATS_QUALHIST + k] = s->mBaseQualHistogram[k];
        t[FPL_STATS_MEDHIST + k] = s->mMedianReadQualHistogram[k];
        t[FPL_STATS_MEDBASES + k] = s->mMedianReadQualBases[k];
    }
    t[FPL_STATS_READS] = s->mReads;
    t[FPL_STATS_LENSUM] = s->mLengthSum;
    // the two row-sum arrays must equal the sums over b — checked here against the reference
    for (int c = 0; c < n; c++) {
        long tb = 0, tq = 0;
        for (int b = 0; b < 8; b++) { tb += s->mCycleBaseContents[b][c]; tq += s->mCycleBaseQual[b][c]; }
        if (tb != s->mCycleTotalBase[c] || tq != s->mCycleTotalQual[c]) return -2;
    }
    return 0;
}

int ref_counters_download(void* hh, int64_t* out, int64_t n_words) {
    Harness* h = (Harness*)hh;
    int64_t need = FPL_COUNTER_WORDS((int64_t)h->adapters.size());
    if (n_words < need) return -1;
    memset(out, 0, sizeof(int64_t) * n_words);
    for (int k = 0; k < 32; k++) out[FPL_CNT_FILTER + k] = h->fr->mFilterReadStats[k];
    out[FPL_CNT_ADAPTER_READS] = h->fr->mTrimmedAdapterRead;
    out[FPL_CNT_ADAPTER_BASES] = h->fr->mTrimmedAdapterBases;
    for (int k = 0; k < 4; k++) {
        out[FPL_CNT_POLYX_READS + k] = h->fr->mTrimmedPolyXReads[k];
        out[FPL_CNT_POLYX_BASES + k] = h->fr->mTrimmedPolyXBases[k];
    }
    out[FPL_CNT_DROPPED] = h->dropped;
    out[FPL_CNT_SPLIT] = h->split;
    for (auto& kv : h->events) {
        uint32_t e = kv.first;
        int c = FPL_EVENT_CMPLEN(e);
        if (c > FPL_MAX_ADAPTER_LEN) return -2;
        out[FPL_CNT_FIXED + ((int64_t)FPL_EVENT_ADAPTER(e) * 2 + FPL_EVENT_SIDE(e)) * (FPL_MAX_ADAPTER_LEN + 1) + c] = kv.second;
    }
    return 0;
}

// Direct access to single reference operators for known-answer tests.
int ref_edit_distance(const char* a, int alen, const char* b, int blen) {
    return (int)edit_distance(a, alen, b, blen);
}
int ref_search_adapter(const char* read, int rlen, const char* adapter, double edMax, int start, int len, int left, int right) {
    std::string r(read, rlen), a(adapter);
    return AdapterTrimmer::searchAdapter(&r, a, edMax, start, len, left != 0, right != 0);
}

}  // extern "C"
