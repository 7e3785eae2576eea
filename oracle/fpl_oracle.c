/*
 * TEST INFRASTRUCTURE ONLY — see fpl_oracle.h.  Plain scalar C restatement of the reference algorithm;
 * every function cites the reference lines it follows (paths relative to /root/reference).
 * Sequence/quality bytes are treated as the reference treats them: `char` (signed on x86-64); bytes >= 0x80
 * index arrays out of bounds in the reference (src/stats.cpp:293) and are outside the domain.
 */
#include "fpl_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t C;        /* capacity in cycles */
    int64_t* content; /* [8][C] */
    int64_t* qual;    /* [8][C] sum of qual-33 */
    int64_t tail[FPL_STATS_TAIL];
} orc_stats;

struct orc_ctx {
    fpl_segment* segs; int64_t n_segs, cap_segs;       /* --mask/--break: output reads of the last orc_process call */
    fpl_region* regs; int64_t n_regs, cap_regs;        /* --mask: masked regions of the last call */
    fpl_options opt;
    int n_adapters;
    char** adapter;
    int* alen;
    orc_stats st[2];
    int64_t* counters;
    int64_t n_counter_words;
};

/* ---- thr(n) = (int)round(edMax * n): src/adaptertrimmer.cpp:73,204,222,275,291 ---- */
static int thr(const orc_ctx* c, int n) { return (int)round(c->opt.ed_max * n); }

/* ---- edit_distance: src/editdistance.cpp:100-126 computes the exact Levenshtein distance
 *      (bit-parallel for <=640 chars, full DP beyond); restated as the textbook two-row DP. ---- */
int orc_edit_distance(const char* a, int alen, const char* b, int blen) {
    if (alen == 0) return blen;
    if (blen == 0) return alen;
    int* prev = (int*)malloc(sizeof(int) * (size_t)(blen + 1));
    int* cur = (int*)malloc(sizeof(int) * (size_t)(blen + 1));
    for (int j = 0; j <= blen; j++) prev[j] = j;
    for (int i = 1; i <= alen; i++) {
        cur[0] = i;
        for (int j = 1; j <= blen; j++) {
            int v = prev[j - 1] + (a[i - 1] == b[j - 1] ? 0 : 1);
            if (prev[j] + 1 < v) v = prev[j] + 1;
            if (cur[j - 1] + 1 < v) v = cur[j - 1] + 1;
            cur[j] = v;
        }
        int* t = prev; prev = cur; cur = t;
    }
    int d = prev[blen];
    free(prev); free(cur);
    return d;
}

static int hamming(const char* r, const char* a, int alen) {
    int m = 0;
    for (int i = 0; i < alen; i++) m += r[i] != a[i];  /* hn::LoadN/!=/CountTrue, src/adaptertrimmer.cpp:93-96 */
    return m;
}

/* ---- AdapterTrimmer::searchAdapter: src/adaptertrimmer.cpp:59-166 ---- */
static int search_adapter(const orc_ctx* c, const char* rdata, int rlen, const char* adata, int alen,
                          int searchStart, int searchLen, int left, int right) {
    int minMismatch = 99999, pos = -1;
    int threshold = thr(c, alen);
    int searchEnd = rlen;
    if (searchLen > 0) searchEnd = rlen < searchLen + searchStart ? rlen : searchLen + searchStart;  /* :76-79 */
    if (searchStart + alen > rlen) return -1;                                                          /* :81-82 */
    if (left) {                                                                                         /* :84-109 */
        for (int p = searchStart; p < searchEnd - alen; p++) {
            int mm = hamming(rdata + p, adata, alen);
            if (mm <= threshold) return p;
            if (mm <= minMismatch) { minMismatch = mm; pos = p; }
        }
    } else if (right && searchEnd > alen) {                                                             /* :110-134 */
        for (int p = searchEnd - alen; p >= searchStart; p--) {
            int mm = hamming(rdata + p, adata, alen);
            if (mm <= threshold) return p;
            if (mm <= minMismatch) { minMismatch = mm; pos = p; }
        }
    } else {                                                                                            /* :135-154 */
        for (int p = searchStart; p < searchEnd - alen; p++) {
            int mm = hamming(rdata + p, adata, alen);
            if (mm < minMismatch) { minMismatch = mm; pos = p; }
        }
    }
    if (pos >= 0) {                                                                                     /* :156-165 */
        int ed = orc_edit_distance(rdata + pos, alen, adata, alen);
        return ed <= threshold ? pos : -1;
    }
    return -1;
}

int orc_search_adapter(const orc_ctx* c, const char* read, int rlen, const char* adapter, int alen,
                       int search_start, int search_len, int left, int right) {
    return search_adapter(c, read, rlen, adapter, alen, search_start, search_len, left, right);
}

/* a read is a window [lo, lo+len) on the original bytes */
typedef struct { int lo, len; } window;

/* Read::trimFront: src/read.cpp:69-73 (negative len wraps to npos in string::erase => everything goes) */
static void trim_front(window* w, int n) {
    if (w->len - 1 < n) n = w->len - 1;
    if (n < 0) { w->lo += w->len; w->len = 0; return; }
    w->lo += n; w->len -= n;
}
/* Read::resize: src/read.cpp:62-67 */
static void resize_(window* w, int n) {
    if (n > w->len || n < 0) return;
    w->len = n;
}

static void add_event(orc_ctx* c, fpl_read_result* out, int idx, int side, int cmplen) {
    uint32_t e = FPL_EVENT(idx, side, cmplen);
    if (out->n_events < FPL_INLINE_EVENTS) out->events[out->n_events] = e;
    out->n_events++;
    c->counters[FPL_CNT_FIXED + ((int64_t)idx * 2 + side) * (FPL_MAX_ADAPTER_LEN + 1) + cmplen]++;
}

/* ---- AdapterTrimmer::trimBySequenceStart: src/adaptertrimmer.cpp:168-236 ---- */
static int trim_start(orc_ctx* c, const char* seq, window* w, int idx, fpl_read_result* out) {
    const int WINDOW = 200, PATTERN_LEN = 16;
    const char* adata = c->adapter[idx];
    int alen = c->alen[idx], ext = c->opt.trimming_extension;
    const char* rdata = seq + w->lo;
    int rlen = w->len;
    if (rlen < PATTERN_LEN) return 0;
    int plen = PATTERN_LEN < alen ? PATTERN_LEN : alen;
    int mpos = search_adapter(c, rdata, rlen, adata, alen, 0, WINDOW, 0, 1);          /* :183 */
    if (mpos >= 0) {
        mpos = mpos + ext < rlen - alen ? mpos + ext : rlen - alen;                    /* :186 */
        add_event(c, out, idx, 0, alen);
        trim_front(w, mpos + alen);
        return mpos + alen;
    }
    int mined = -1, pos = -1;
    for (int p = 0; p < rlen - plen && p < WINDOW - plen; p++) {                       /* :202-216 */
        int ed = orc_edit_distance(rdata + p, plen, adata + alen - plen, plen);
        if (ed <= thr(c, plen)) {
            if (pos < 0) { pos = p; mined = ed; }
            else if (ed >= mined) { }
            else { pos = p; mined = ed; }
        }
    }
    if (pos >= 0) {                                                                    /* :218-233 */
        int cmplen = pos + plen < alen ? pos + plen : alen;
        int ed = orc_edit_distance(rdata + pos + plen - cmplen, cmplen, adata + alen - cmplen, cmplen);
        if (ed <= thr(c, cmplen)) {
            pos = pos + ext < rlen - alen ? pos + ext : rlen - alen;
            add_event(c, out, idx, 0, cmplen);
            trim_front(w, pos + plen);
            return pos + plen;
        }
    }
    return 0;
}

/* ---- AdapterTrimmer::trimBySequenceEnd: src/adaptertrimmer.cpp:238-302 ---- */
static int trim_end(orc_ctx* c, const char* seq, window* w, int idx, fpl_read_result* out) {
    const int WINDOW = 200, PATTERN_LEN = 16;
    const char* adata = c->adapter[idx];
    int alen = c->alen[idx], ext = c->opt.trimming_extension;
    const char* rdata = seq + w->lo;
    int rlen = w->len;
    if (rlen < PATTERN_LEN) return 0;
    int plen = PATTERN_LEN < alen ? PATTERN_LEN : alen;
    int searchStart = rlen - WINDOW > 0 ? rlen - WINDOW : 0;
    int mpos = search_adapter(c, rdata, rlen, adata, alen, searchStart, WINDOW, 1, 0); /* :254 */
    if (mpos >= 0) {
        mpos = mpos - ext > 0 ? mpos - ext : 0;                                        /* :257 */
        add_event(c, out, idx, 1, alen);
        resize_(w, mpos);
        return rlen - mpos;
    }
    int mined = -1, pos = -1;
    for (int p = 0; p < rlen - plen && p < WINDOW - plen; p++) {                       /* :273-286 */
        int ed = orc_edit_distance(rdata + rlen - plen - p, plen, adata, plen);
        if (ed <= thr(c, plen)) {
            if (pos < 0) { pos = p; mined = ed; }
            else if (ed > mined) break;
            else { pos = p; mined = ed; }
        }
    }
    if (pos > 0) {                                                                     /* :288 (> 0, not >= 0) */
        int cmplen = pos + plen < alen ? pos + plen : alen;
        if (orc_edit_distance(rdata + rlen - plen - pos, cmplen, adata, cmplen) <= thr(c, cmplen)) {
            pos = pos + ext < rlen - plen ? pos + ext : rlen - plen;
            add_event(c, out, idx, 1, cmplen);
            resize_(w, rlen - plen - pos);
            return pos + plen;
        }
    }
    return 0;
}

/* ---- AdapterTrimmer::findMiddleAdapters: src/adaptertrimmer.cpp:13-40 ---- */
static int find_middle(const orc_ctx* c, const char* seq, const window* w, int* start, int* len) {
    const char* r = seq + w->lo;
    int L = w->len, ext = c->opt.trimming_extension;
    int slen = c->alen[0], elen = c->alen[1];
    int sp = search_adapter(c, r, L, c->adapter[0], slen, 0, -1, 0, 0);
    int ep = search_adapter(c, r, L, c->adapter[1], elen, 0, -1, 0, 0);
    *len = -1;
    if (sp >= 0 && ep >= 0) {
        int st = sp < ep ? sp : ep;
        int en = sp + slen > ep + elen ? sp + slen : ep + elen;
        st = st - ext > 0 ? st - ext : 0;
        en = en + ext < L ? en + ext : L;
        *start = st; *len = en - st;
        return 1;
    }
    if (sp >= 0) {
        int en = sp + slen + ext < L ? sp + slen + ext : L;
        *start = sp - ext > 0 ? sp - ext : 0;
        *len = en - *start;
        return 1;
    }
    if (ep >= 0) {
        int en = ep + elen + ext < L ? ep + elen + ext : L;
        *start = ep - ext > 0 ? ep - ext : 0;
        *len = en - *start;
        return 1;
    }
    return 0;
}

/* ---- Filter::trimAndCut: src/filter.cpp:130-232.  Returns 1 if the read is dropped (NULL). ---- */
static int trim_and_cut(const orc_ctx* c, const char* seq, const char* qualstr, int l, window* w) {
    const fpl_options* o = &c->opt;
    int front = o->trim_front, tail = o->trim_tail;
    w->lo = 0; w->len = l;
    if (front == 0 && tail == 0 && !o->cut_front_enabled && !o->cut_tail_enabled) return 0;  /* :133-134 */
    int rlen = l - front - tail;
    if (rlen < 0) return 1;                                                                  /* :138-139 */
    if (front == 0 && !o->cut_front_enabled && !o->cut_tail_enabled) { resize_(w, rlen); return 0; }
    if (!o->cut_front_enabled && !o->cut_tail_enabled) { w->lo = front; w->len = rlen; return 0; }
    if (o->cut_front_enabled) {                                                              /* :159-189 */
        int wsz = o->cut_front_window;
        int s = front;
        if (l - front - tail - wsz <= 0) return 1;
        int totalQual = 0;
        for (int i = 0; i < wsz - 1; i++) totalQual += qualstr[s + i];
        for (s = front; s + wsz < l - tail; s++) {
            totalQual += qualstr[s + wsz - 1];
            if (s > front) totalQual -= qualstr[s - 1];
            if ((double)totalQual / (double)wsz >= 33 + o->cut_front_quality) break;
        }
        if (s > 0) s = s + wsz - 1;
        while (s < l && seq[s] == 'N') s++;
        front = s;
        rlen = l - front - tail;
    }
    if (o->cut_tail_enabled) {                                                               /* :191-219 */
        int wsz = o->cut_tail_window;
        if (l - front - tail - wsz <= 0) return 1;
        int totalQual = 0;
        int t = l - tail - 1;
        for (int i = 0; i < wsz - 1; i++) totalQual += qualstr[t - i];
        for (t = l - tail - 1; t - wsz >= front; t--) {
            totalQual += qualstr[t - wsz + 1];
            if (t < l - tail - 1) totalQual -= qualstr[t + 1];
            if ((double)totalQual / (double)wsz >= 33 + o->cut_tail_quality) break;
        }
        if (t < l - 1) t = t - wsz + 1;
        while (t >= 0 && seq[t] == 'N') t--;
        rlen = t - front + 1;
    }
    if (rlen <= 0 || front >= l - 1) return 1;                                               /* :221-222 */
    w->lo = front; w->len = rlen;
    return 0;
}

int orc_trim_and_cut(const orc_ctx* c, const char* seq, const char* qual, int len, int* lo, int* rlen) {
    window w;
    int dropped = trim_and_cut(c, seq, qual, len, &w);
    *lo = w.lo; *rlen = w.len;
    return dropped;
}

/* ---- PolyX::trimPolyX: src/polyx.cpp:11-78 ---- */
int orc_trim_polyx(const char* data, int rlen, int compareReq, int* base, int* plen) {
    const int allowOneMismatchForEach = 8, maxMismatch = 5;
    static const char ATCG[4] = {'A', 'T', 'C', 'G'};  /* src/common.h:29 */
    int n[4] = {0, 0, 0, 0};
    int pos;
    *base = -1; *plen = 0;
    for (pos = 0; pos < rlen; pos++) {
        switch (data[rlen - pos - 1]) {
            case 'A': n[0]++; break;
            case 'T': n[1]++; break;
            case 'C': n[2]++; break;
            case 'G': n[3]++; break;
            case 'N': n[0]++; n[1]++; n[2]++; n[3]++; break;
            default: break;
        }
        int cmp = pos + 1;
        int allowed = cmp / allowOneMismatchForEach < maxMismatch ? cmp / allowOneMismatchForEach : maxMismatch;
        int needToBreak = 1;
        for (int b = 0; b < 4; b++) if (cmp - n[b] <= allowed) needToBreak = 0;
        if (needToBreak && (pos >= allowOneMismatchForEach || pos + 1 >= compareReq - 1)) break;
    }
    if (pos + 1 >= compareReq) {                                                             /* :57 */
        int poly = 0, maxCount = -1;
        for (int b = 0; b < 4; b++) if (n[b] > maxCount) { maxCount = n[b]; poly = b; }
        char polyBase = ATCG[poly];
        /* :68: data[rlen-pos-1] with pos == rlen reads data[-1] (UB); in practice that byte is 0 => unequal.
         * with pos == -1 it reads the NUL terminator => unequal, then `pos>=0` ends the loop. */
        while (pos >= 0) {
            int idx = rlen - pos - 1;
            char ch = (idx < 0 || idx >= rlen) ? 0 : data[idx];
            if (ch == polyBase) break;
            pos--;
        }
        *base = poly; *plen = pos + 1;
        int nl = rlen - pos - 1;
        if (nl > rlen || nl < 0) return rlen;  /* Read::resize no-op, src/read.cpp:62-64 */
        return nl;
    }
    return rlen;
}

/* ---- Stats::statRead: src/stats.cpp:265-375 (live state only) ---- */
static void stats_reserve(orc_stats* s, int64_t need) {
    if (need <= s->C) return;
    int64_t nc = s->C ? s->C : 1024;
    while (nc < need) nc *= 2;
    int64_t* c2 = (int64_t*)calloc((size_t)(8 * nc), sizeof(int64_t));
    int64_t* q2 = (int64_t*)calloc((size_t)(8 * nc), sizeof(int64_t));
    for (int b = 0; b < 8; b++) {
        if (s->C) {
            memcpy(c2 + b * nc, s->content + b * s->C, sizeof(int64_t) * (size_t)s->C);
            memcpy(q2 + b * nc, s->qual + b * s->C, sizeof(int64_t) * (size_t)s->C);
        }
    }
    free(s->content); free(s->qual);
    s->content = c2; s->qual = q2; s->C = nc;
}

static int base2val(char base) {  /* src/stats.cpp:411-425 */
    switch (base) {
        case 'A': return 0;
        case 'T': case 'U': return 1;
        case 'C': return 2;
        case 'G': return 3;
        default: return -1;
    }
}

static int stat_read(orc_stats* s, const char* seqstr, const char* qualstr, int len) {
    stats_reserve(s, len);
    s->tail[FPL_STATS_LENSUM] += len;
    int qualHist[128];
    memset(qualHist, 0, sizeof(qualHist));
    int kmer = 0, needFullCompute = 1;
    for (int i = 0; i < len; i++) {
        char base = seqstr[i], qual = qualstr[i];
        int b = base & 0x07;
        s->tail[FPL_STATS_QUALHIST + qual]++;
        qualHist[(int)qual]++;
        s->content[b * s->C + i]++;
        s->qual[b * s->C + i] += qual - 33;
        if (base == 'N') { needFullCompute = 1; continue; }
        if (i < 4) continue;
        if (!needFullCompute) {
            int val = base2val(base);
            if (val < 0) { needFullCompute = 1; continue; }
            kmer = ((kmer << 2) & 0x3FC) | val;
            s->tail[FPL_STATS_KMER + kmer]++;
        } else {
            int valid = 1;
            kmer = 0;
            for (int k = 0; k < 5; k++) {
                int val = base2val(seqstr[i - 4 + k]);
                if (val < 0) { valid = 0; break; }
                kmer = ((kmer << 2) & 0x3FC) | val;
            }
            if (!valid) { needFullCompute = 1; continue; }
            s->tail[FPL_STATS_KMER + kmer]++;
            needFullCompute = 0;
        }
    }
    int median = 0;
    if (len > 0) {                                                                           /* :351-361 */
        int total = 0, half = len >> 1;
        while (1) {
            total += qualHist[median];
            if (total > half) break;
            median++;
        }
        s->tail[FPL_STATS_MEDHIST + median]++;
        s->tail[FPL_STATS_MEDBASES + median] += len;
    }
    s->tail[FPL_STATS_READS]++;
    return median;
}

/* ---- Filter::passFilter + passLowComplexityFilter: src/filter.cpp:12-81 ---- */
static int pass_filter(const orc_ctx* c, const char* seqstr, const char* qualstr, int rlen) {
    const fpl_options* o = &c->opt;
    if (rlen == 0) return FPL_FAIL_LENGTH;
    int lowQualNum = 0, nBaseNum = 0, totalQual = 0;
    if (o->qual_filter_enabled || o->length_filter_enabled) {
        for (int i = 0; i < rlen; i++) {
            char base = seqstr[i], qual = qualstr[i];
            totalQual += qual - 33;
            if (qual < (char)o->qualified_qual) lowQualNum++;
            if (base == 'N') nBaseNum++;
        }
    }
    if (o->qual_filter_enabled) {
        if (lowQualNum > (o->unqualified_percent_limit * rlen / 100.0)) return FPL_FAIL_QUALITY;
        else if (o->avg_qual_req > 0 && (totalQual / rlen) < o->avg_qual_req) return FPL_FAIL_QUALITY;
        else if (nBaseNum * 100 > rlen * o->n_base_percent_limit) return FPL_FAIL_N_BASE;
        else if (o->n_base_limit != 1000000 && nBaseNum > o->n_base_limit) return FPL_FAIL_N_BASE;
    }
    if (o->length_filter_enabled) {
        if (rlen < o->length_required) return FPL_FAIL_LENGTH;
        if (o->length_max > 0 && rlen > o->length_max) return FPL_FAIL_TOO_LONG;
    }
    if (o->complexity_enabled) {
        int diff = 0;
        if (rlen <= 1) return FPL_FAIL_COMPLEXITY;
        for (int i = 0; i < rlen - 1; i++) if (seqstr[i] != seqstr[i + 1]) diff++;
        double threshold = o->complexity_threshold_pct / 100.0;  /* src/main.cpp:205 */
        if (!((double)diff / (double)(rlen - 1) >= threshold)) return FPL_FAIL_COMPLEXITY;
    }
    return FPL_PASS_FILTER;
}

/* ---- Filter::detectLowQualityRegions: src/filter.cpp:83-128, transcribed literally (the pre-sum loop
 *      bound `i < windowSize-1` is absolute, so after the first region the running sum restarts from 0).
 *      Calls emit(first, last) (0-based, inclusive) per region; returns the number of regions. ---- */
typedef void (*region_cb)(void* u, int first, int last);
static int detect_low_quality_regions(const char* qualstr, int l, int windowSize, int quality, region_cb emit, void* u) {
    int n = 0;
    if (l == 0 || windowSize <= 0) return 0;
    int start = 0;
    while (start + windowSize <= l) {
        int totalQual = 0;
        for (int i = start; i < windowSize - 1 && i < l; i++) totalQual += qualstr[i];
        int windowStart = -1;
        for (int s = start; s + windowSize < l; s++) {
            if (totalQual < (33 + quality) * windowSize) { windowStart = s; break; }
            totalQual += qualstr[s + windowSize];
            totalQual -= qualstr[s];
        }
        if (windowStart == -1) break;
        int e;
        for (e = windowStart; e + windowSize < l; e++) {
            totalQual += qualstr[e + windowSize];
            totalQual -= qualstr[e];
            if (totalQual >= (33 + quality) * windowSize) break;
        }
        if (emit) emit(u, windowStart, e + windowSize - 1);
        n++;
        start = e + windowSize;
    }
    return n;
}

typedef struct { int lo, len, split_side, is_r1, break_index; } piece;
typedef struct { piece* v; int n, cap; } piece_list;
static void pl_push(piece_list* l, piece p) {
    if (l->n == l->cap) { l->cap = l->cap ? 2 * l->cap : 8; l->v = (piece*)realloc(l->v, sizeof(piece) * (size_t)l->cap); }
    l->v[l->n++] = p;
}
typedef struct { int (*v)[2]; int n, cap; } region_list;
static void rl_emit(void* u, int first, int last) {
    region_list* l = (region_list*)u;
    if (l->n == l->cap) { l->cap = l->cap ? 2 * l->cap : 8; l->v = (int(*)[2])realloc(l->v, sizeof(int[2]) * (size_t)l->cap); }
    l->v[l->n][0] = first; l->v[l->n][1] = last; l->n++;
}

/* Read::breakByRegions: src/read.cpp:227-262 */
static void break_by_regions(const piece* rr, const region_list* regions, piece_list* out) {
    int lastEnd = -1, length = rr->len;
    for (int i = 0; i < regions->n; i++) {
        int start = regions->v[i][0], end = regions->v[i][1];
        if (start < 0) start = 0;
        if (end >= length) end = length - 1;
        if (start > end || start >= length) continue;
        if (start > lastEnd + 1) {
            piece p = {rr->lo + lastEnd + 1, start - lastEnd - 1, rr->split_side, 0, i + 1};
            pl_push(out, p);
        }
        lastEnd = end;
    }
    if (lastEnd < length - 1) {
        piece p = {rr->lo + lastEnd + 1, length - lastEnd - 1, rr->split_side, 0, regions->n + 1};
        pl_push(out, p);
    }
}

static void push_segment(orc_ctx* c, fpl_segment sg) {
    if (c->n_segs == c->cap_segs) { c->cap_segs = c->cap_segs ? 2 * c->cap_segs : 1024; c->segs = (fpl_segment*)realloc(c->segs, sizeof(fpl_segment) * (size_t)c->cap_segs); }
    c->segs[c->n_segs++] = sg;
}
static void push_region(orc_ctx* c, fpl_region r) {
    if (c->n_regs == c->cap_regs) { c->cap_regs = c->cap_regs ? 2 * c->cap_regs : 1024; c->regs = (fpl_region*)realloc(c->regs, sizeof(fpl_region) * (size_t)c->cap_regs); }
    c->regs[c->n_regs++] = r;
}

/* --mask / --break variant of the tail of processSingleEnd (src/seprocessor.cpp:235-288) for one read whose outReads
 * after the adapter stage are seg[0..nseg).  seq is the read's ORIGINAL bases; masking works on a private copy. */
static void finish_read_ext(orc_ctx* c, int64_t ri, const char* seq, const char* qual, int L, const window* seg, int nseg,
                            int split, int seg0right, fpl_read_result* out) {
    const fpl_options* o = &c->opt;
    piece_list reads = {0, 0, 0};
    for (int k = 0; k < nseg; k++) {
        piece p = {seg[k].lo, seg[k].len, split ? ((k == 1 || seg0right) ? 2 : 1) : 0, split ? 0 : 1, 0};
        pl_push(&reads, p);
    }
    if (o->break_enabled && reads.n > 0) {                                                   /* :235-252 */
        piece_list tmp = {0, 0, 0};
        for (int i = 0; i < reads.n; i++) {
            region_list regions = {0, 0, 0};
            detect_low_quality_regions(qual + reads.v[i].lo, reads.v[i].len, o->break_window, o->break_quality, rl_emit, &regions);
            if (regions.n > 0) break_by_regions(&reads.v[i], &regions, &tmp);
            else pl_push(&tmp, reads.v[i]);
            free(regions.v);
        }
        free(reads.v);
        reads = tmp;
    }
    char* mseq = NULL;
    if (o->mask_enabled && reads.n > 0) {                                                    /* :254-262 */
        mseq = (char*)malloc((size_t)L + 1);
        memcpy(mseq, seq, (size_t)L);
        for (int i = 0; i < reads.n; i++) {
            region_list regions = {0, 0, 0};
            detect_low_quality_regions(qual + reads.v[i].lo, reads.v[i].len, o->mask_window, o->mask_quality, rl_emit, &regions);
            for (int j = 0; j < regions.n; j++) {                                            /* Read::maskRegionWithN, src/read.cpp:217-225 */
                int start = regions.v[j][0], len = regions.v[j][1] - regions.v[j][0] + 1, length = reads.v[i].len;
                if (start < 0 || len <= 0 || start >= length) continue;
                if (start + len > length) len = length - start;
                memset(mseq + reads.v[i].lo + start, 'N', (size_t)len);
                fpl_region r = {(int32_t)ri, reads.v[i].lo + start, len};
                push_region(c, r);
            }
            free(regions.v);
        }
    }
    const char* fseq = mseq ? mseq : seq;
    out->n_segments = reads.n;
    for (int k = 0; k < reads.n; k++) {                                                      /* :264-288 */
        const piece* p = &reads.v[k];
        int result = pass_filter(c, fseq + p->lo, qual + p->lo, p->len);
        c->counters[FPL_CNT_FILTER + result]++;
        fpl_segment sg;
        memset(&sg, 0, sizeof(sg));
        sg.read = (int32_t)ri; sg.lo = p->lo; sg.len = p->len; sg.result = (uint8_t)result;
        sg.split_side = (uint8_t)p->split_side; sg.is_r1 = (uint8_t)p->is_r1; sg.break_index = p->break_index;
        if (result == FPL_PASS_FILTER) sg.median_qual = (uint8_t)stat_read(&c->st[1], fseq + p->lo, qual + p->lo, p->len);
        push_segment(c, sg);
        if (k < 2) {
            out->seg_lo[k] = p->lo; out->seg_len[k] = p->len; out->seg_result[k] = (uint8_t)result;
            out->seg_median_qual[k] = sg.median_qual;
        }
    }
    free(reads.v);
    free(mseq);
}

/* ---- SingleEndProcessor::processSingleEnd per-read body: src/seprocessor.cpp:186-295 ---- */
int orc_process(orc_ctx* c, const fpl_batch* b, fpl_read_result* results) {
    const fpl_options* o = &c->opt;
    const int ext = o->mask_enabled || o->break_enabled;
    c->n_segs = 0; c->n_regs = 0;
    for (int64_t i = 0; i < b->n_reads; i++) {
        fpl_read_result* out = &results[i];
        memset(out, 0, sizeof(*out));
        const char* seq = (const char*)b->seq + b->offsets[i];
        const char* qual = (const char*)b->qual + b->offsets[i];
        int L = b->lens[i];
        out->pre_median_qual = (uint8_t)stat_read(&c->st[0], seq, qual, L);                  /* :192 */
        window w;
        int alive = !trim_and_cut(c, seq, qual, L, &w);                                     /* :196 */
        if (!alive) { out->flags |= FPL_FLAG_DROPPED_BY_CUT; c->counters[FPL_CNT_DROPPED]++; }
        if (alive && o->polyx_enabled) {                                                     /* :198-201 */
            int base, plen;
            int nl = orc_trim_polyx(seq + w.lo, w.len, o->polyx_min_len, &base, &plen);
            if (base >= 0) {
                w.len = nl;
                out->flags |= FPL_FLAG_POLYX;
                out->polyx_base = (uint8_t)base;
                out->polyx_len = plen;
                c->counters[FPL_CNT_POLYX_READS + base]++;
                c->counters[FPL_CNT_POLYX_BASES + base] += plen;
            }
        }
        window seg[2];
        int nseg = 0, split = 0, seg0right = 0;
        if (alive && o->adapter_enabled) {                                                   /* :205-229 */
            int trimmed = 0;
            if (c->alen[0] > 0) trimmed += trim_start(c, seq, &w, 0, out);
            if (c->alen[1] > 0) trimmed += trim_end(c, seq, &w, 1, out);
            for (int k = 2; k < c->n_adapters; k++) {                                        /* adaptertrimmer.cpp:51-54 */
                trimmed += trim_start(c, seq, &w, k, out);
                trimmed += trim_end(c, seq, &w, k, out);
            }
            if (trimmed > 0) {
                c->counters[FPL_CNT_ADAPTER_READS]++;
                c->counters[FPL_CNT_ADAPTER_BASES] += trimmed;
            }
            out->adapter_trimmed_bases = trimmed;
            int start, len;
            if (find_middle(c, seq, &w, &start, &len)) {
                out->flags |= FPL_FLAG_MIDDLE_ADAPTER;
                c->counters[FPL_CNT_SPLIT]++;
                int len1 = start, len2 = w.len - start - len;                                /* Read::breakByGap, src/read.cpp:192-215 */
                if (len1 > 0) { seg[nseg].lo = w.lo; seg[nseg].len = len1; nseg++; }
                if (len2 > 0) { seg[nseg].lo = w.lo + start + len; seg[nseg].len = len2; nseg++; }
                if (nseg == 1 && len1 <= 0) out->flags |= FPL_FLAG_SEG0_IS_RIGHT;
                split = 1; seg0right = (nseg == 1 && len1 <= 0);
            } else {
                seg[nseg++] = w;
            }
        } else if (alive) {
            seg[nseg++] = w;
        }
        if (alive) { out->trim_lo = w.lo; out->trim_len = w.len; }
        if (ext) { finish_read_ext(c, i, seq, qual, L, seg, nseg, split, seg0right, out); continue; }
        out->n_segments = nseg;
        for (int k = 0; k < nseg; k++) {                                                     /* :264-288 */
            int result = pass_filter(c, seq + seg[k].lo, qual + seg[k].lo, seg[k].len);
            c->counters[FPL_CNT_FILTER + result]++;
            out->seg_lo[k] = seg[k].lo;
            out->seg_len[k] = seg[k].len;
            out->seg_result[k] = (uint8_t)result;
            if (result == FPL_PASS_FILTER)
                out->seg_median_qual[k] = (uint8_t)stat_read(&c->st[1], seq + seg[k].lo, qual + seg[k].lo, seg[k].len);
        }
    }
    return 0;
}

orc_ctx* orc_create(const fpl_options* opt, const fpl_adapters* ad) {
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
    c->opt = *opt;
    c->n_adapters = 2 + ad->n_fasta;
    c->adapter = (char**)calloc((size_t)c->n_adapters, sizeof(char*));
    c->alen = (int*)calloc((size_t)c->n_adapters, sizeof(int));
    for (int k = 0; k < c->n_adapters; k++) {
        const char* s = k == 0 ? ad->start : k == 1 ? ad->end : ad->fasta[k - 2];
        if (!s) s = "";
        c->adapter[k] = strdup(s);
        c->alen[k] = (int)strlen(s);
    }
    c->n_counter_words = FPL_COUNTER_WORDS(c->n_adapters);
    c->counters = (int64_t*)calloc((size_t)c->n_counter_words, sizeof(int64_t));
    return c;
}

void orc_destroy(orc_ctx* c) {
    if (!c) return;
    for (int k = 0; k < c->n_adapters; k++) free(c->adapter[k]);
    free(c->adapter); free(c->alen); free(c->counters); free(c->segs); free(c->regs);
    for (int k = 0; k < 2; k++) { free(c->st[k].content); free(c->st[k].qual); }
    free(c);
}

int orc_last_segments(orc_ctx* c, fpl_segment* out, int64_t cap, int64_t* n) {
    *n = c->n_segs;
    if (c->n_segs > cap) return -1;
    if (c->n_segs) memcpy(out, c->segs, sizeof(fpl_segment) * (size_t)c->n_segs);
    return 0;
}
int orc_last_mask_regions(orc_ctx* c, fpl_region* out, int64_t cap, int64_t* n) {
    *n = c->n_regs;
    if (c->n_regs > cap) return -1;
    if (c->n_regs) memcpy(out, c->regs, sizeof(fpl_region) * (size_t)c->n_regs);
    return 0;
}

int64_t orc_stats_cycles(orc_ctx* c) { return c->st[0].C > c->st[1].C ? c->st[0].C : c->st[1].C; }

int orc_stats_download(orc_ctx* c, int which, int64_t* out, int64_t C) {
    orc_stats* s = &c->st[which];
    memset(out, 0, sizeof(int64_t) * (size_t)FPL_STATS_WORDS(C));
    int64_t n = s->C < C ? s->C : C;
    for (int b = 0; b < 8; b++)
        for (int64_t k = 0; k < s->C; k++) {
            if (k < n) {
                out[b * C + k] = s->content[b * s->C + k];
                out[8 * C + b * C + k] = s->qual[b * s->C + k];
            } else if (s->content[b * s->C + k] != 0) return -1;
        }
    memcpy(out + 16 * C, s->tail, sizeof(s->tail));
    return 0;
}

int orc_counters_download(orc_ctx* c, int64_t* out, int64_t n_words) {
    if (n_words < c->n_counter_words) return -1;
    memset(out, 0, sizeof(int64_t) * (size_t)n_words);
    memcpy(out, c->counters, sizeof(int64_t) * (size_t)c->n_counter_words);
    return 0;
}
