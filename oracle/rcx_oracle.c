/* rcx_oracle.c — CPU restatement (plain C) of Rcorrector's stage-3 hot path.
 *
 * TEST INFRASTRUCTURE ONLY — see rcx_oracle.h.  Every function cites the
 * reference file:line whose behaviour it restates; citations are into
 * /root/reference (mourisl/Rcorrector @ 46bc390).  The restatement is written
 * from the functional spec in SURVEY.md Appendix A, with its own data
 * structures (prefix-sum poly-A tests, an open-addressed table, a search
 * context struct) and is validated byte-for-byte against the reference binary
 * built by oracle/Makefile (tests/test_oracle_vs_ref.py).
 */
#include "rcx_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_LEN 1024   /* utils.h:7 */
#define ORC_MAX_TRIAL 1025 /* ErrorCorrection.cpp:7 */
#define ORC_INF 1000000000 /* utils.h:10 */

/* ------------------------------------------------------------------ k-mer codec
 * KmerCode.h:7-89, KmerCode.cpp:7-52: 2-bit code, newest base in
 * the low bits, ONE tracked invalid position `inv` (-1 = none). */
typedef struct {
    uint64_t code;
    int inv;
} kmer;

static inline int nuc(char c) {
    /* nucToNum, main.cpp:17-20: A0 C1 G2 T3, everything else -1 */
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return -1;
    }
}

static inline uint64_t kmask(int k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1); }

static inline void k_append(kmer* x, char c, int k, uint64_t mask) { /* KmerCode.cpp:7-21 */
    int v = nuc(c);
    if (x->inv != -1) ++x->inv;
    x->code = ((x->code << 2) & mask) | (uint64_t)(v & 3);
    if (v < 0) x->inv = 0;
    if (x->inv >= k) x->inv = -1;
}

static inline void k_shift_right(kmer* x, int n, uint64_t mask) { /* KmerCode.cpp:33-42 */
    if (x->inv != -1) x->inv -= n;
    x->code = (x->code >> (2 * n)) & (mask >> (2 * n));
    if (x->inv < 0) x->inv = -1;
}

static inline void k_prepend(kmer* x, char c, int k, uint64_t mask) { /* KmerCode.cpp:23-31 */
    int v = nuc(c);
    k_shift_right(x, 1, mask);
    if (v < 0) x->inv = k - 1;
    x->code = (x->code | ((uint64_t)(v & 3) << (2 * (k - 1)))) & mask;
}

static inline uint64_t canonical(uint64_t code, int k) { /* KmerCode.h:58-71 */
    uint64_t rc = 0, x = code;
    for (int i = 0; i < k; ++i) {
        rc = (rc << 2) | (3 - (x & 3));
        x >>= 2;
    }
    return rc < code ? rc : code;
}

/* ------------------------------------------------------------------ table
 * Store.h:17-88 semantics on an open-addressed array (keys are canonical
 * codes; ~0 is never canonical for k <= 32 because its reverse complement is 0). */
struct orc_table {
    int k;
    uint64_t cap, used;
    uint64_t* keys;
    int32_t* vals;
};

static inline uint64_t hash64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

static void table_alloc(orc_table* t, uint64_t cap) {
    t->cap = cap;
    t->keys = (uint64_t*)malloc(cap * sizeof(uint64_t));
    t->vals = (int32_t*)malloc(cap * sizeof(int32_t));
    memset(t->keys, 0xff, cap * sizeof(uint64_t));
}

orc_table* orc_table_new(int k) {
    orc_table* t = (orc_table*)calloc(1, sizeof(orc_table));
    t->k = k;
    table_alloc(t, 1u << 16);
    return t;
}

void orc_table_free(orc_table* t) {
    if (!t) return;
    free(t->keys);
    free(t->vals);
    free(t);
}

uint64_t orc_table_size(const orc_table* t) { return t->used; }

static void table_set(orc_table* t, uint64_t key, int32_t v) {
    uint64_t m = t->cap - 1, i = hash64(key) & m;
    while (t->keys[i] != ~0ull && t->keys[i] != key) i = (i + 1) & m;
    if (t->keys[i] == ~0ull) {
        t->keys[i] = key;
        ++t->used;
    }
    t->vals[i] = v; /* last writer wins, Store.h:55 */
}

static void table_grow(orc_table* t, uint64_t want) {
    if (want * 2 <= t->cap) return;
    uint64_t ncap = t->cap;
    while (want * 2 > ncap) ncap <<= 1;
    orc_table old = *t;
    table_alloc(t, ncap);
    t->used = 0;
    for (uint64_t i = 0; i < old.cap; ++i)
        if (old.keys[i] != ~0ull) table_set(t, old.keys[i], old.vals[i]);
    free(old.keys);
    free(old.vals);
}

static inline int32_t table_get(const orc_table* t, uint64_t key) {
    uint64_t m = t->cap - 1, i = hash64(key) & m;
    while (t->keys[i] != ~0ull) {
        if (t->keys[i] == key) return t->vals[i];
        i = (i + 1) & m;
    }
    return 0;
}

void orc_table_put(orc_table* t, const uint64_t* codes, const int32_t* counts, uint64_t n) {
    table_grow(t, t->used + n);
    for (uint64_t i = 0; i < n; ++i) {
        if (counts[i] <= 1) continue; /* main.cpp:299 */
        table_set(t, canonical(codes[i], t->k), counts[i]);
    }
}

int32_t orc_table_count(const orc_table* t, uint64_t fwd_code) { return table_get(t, canonical(fwd_code, t->k)); }

static inline int32_t k_count(const orc_table* t, const kmer* x) { /* Store::GetCount, Store.h:59-66 */
    if (x->inv != -1) return 0;
    return table_get(t, canonical(x->code, t->k));
}

int64_t orc_table_load_dump(orc_table* t, const char* path) { /* main.cpp:294-308 */
    FILE* fp = fopen(path, "r");
    if (!fp) return -1;
    char a[128], b[128];
    int64_t stored = 0;
    const uint64_t mask = kmask(t->k);
    while (fscanf(fp, "%99s", a) == 1) {
        int count = atoi(a + 1);
        if (fscanf(fp, "%99s", b) != 1) break;
        if (count <= 1) continue;
        kmer x = {0, -1};
        for (int i = 0; b[i]; ++i) k_append(&x, b[i], t->k, mask);
        if (x.inv == -1) { /* Store::Put ignores invalid k-mers, Store.h:53-54 */
            table_grow(t, t->used + 1);
            table_set(t, canonical(x.code, t->k), count);
        }
        ++stored; /* kmerCount counts the record either way, main.cpp:306 */
    }
    fclose(fp);
    return stored;
}

static int cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return x > y ? 1 : x < y ? -1 : 0;
}

double orc_estimate_error_rate(const orc_table* t, const char* dump_path, double wk) { /* main.cpp:310-358, SURVEY A.10 */
    FILE* fp = fopen(dump_path, "r");
    if (!fp) return -1.0;
    const int cap = 100000;
    const int k = t->k;
    const uint64_t mask = kmask(k);
    double* rates = (double*)malloc(sizeof(double) * (cap + 1));
    rates[0] = 0;
    int n = 0;
    char a[128], b[128];
    /* the reference tests IsValid() on the state left by the PREVIOUS record (main.cpp:323);
       before the first record that state is whatever the load loop left behind: the last
       record of the dump.  Reproduce by carrying `x` across records and seeding it the same way. */
    kmer x = {0, -1};
    {
        FILE* f2 = fopen(dump_path, "r");
        char la[128], lb[128];
        while (f2 && fscanf(f2, "%99s", la) == 1 && fscanf(f2, "%99s", lb) == 1) {
            if (atoi(la + 1) <= 1) continue;
            x.code = 0; x.inv = -1;
            for (int i = 0; lb[i]; ++i) k_append(&x, lb[i], k, mask);
        }
        if (f2) fclose(f2);
    }
    while (n < cap && fscanf(fp, "%99s", a) == 1) {
        if (fscanf(fp, "%99s", b) != 1) break;
        if (x.inv != -1) continue;
        x.code = 0; x.inv = -1;
        for (int i = 0; b[i]; ++i) k_append(&x, b[i], k, mask);
        int best = 0, second = 0;
        for (int c = 0; c < 4; ++c) {
            k_shift_right(&x, 1, mask);
            k_append(&x, "ACGT"[c], k, mask);
            int cnt = k_count(t, &x);
            if (cnt > best) { second = best; best = cnt; }
            else if (cnt > second) second = cnt;
        }
        if (best < 1000) continue;
        rates[n++] = (double)second / (double)best;
    }
    fclose(fp);
    double er;
    if (n > 0) {
        qsort(rates, n, sizeof(double), cmp_double);
        rates[n] = rates[n - 1];
        er = rates[(int)(n * wk)];
    } else {
        er = 0; /* reference reads rates[-1] here (UB); falls back below since n < 100 */
    }
    if (er == 0 || n < 100) er = 0.01;
    free(rates);
    return er;
}

int32_t orc_bad_quality(const char* qual, const uint32_t* off, int64_t n) { /* main.cpp:88-128 */
    int first[300], last[300];
    memset(first, 0, sizeof first);
    memset(last, 0, sizeof last);
    int64_t tot = n < 1000000 ? n : 1000000;
    for (int64_t r = 0; r < tot; ++r) {
        ++last[(int)qual[off[r + 1] - 1]];
        ++first[(int)qual[off[r]]];
    }
    int i, cnt = 0, t1, t2;
    for (i = 0; i < 300; ++i) { cnt += first[i]; if (cnt > tot * 0.05) break; }
    t1 = i - 1;
    cnt = 0;
    for (i = 0; i < 300; ++i) { cnt += last[i]; if (cnt > tot * 0.05) break; }
    t2 = i;
    return (int32_t)(char)(t2 < t1 ? t2 : t1);
}

/* ------------------------------------------------------------------ helpers */
static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; } /* ErrorCorrection.cpp:37-40 */

static inline double bound_d(int c, double er) { /* GetBound, ErrorCorrection.cpp:139-142 */
    return c * er + 6.0 * sqrt(c * er) + 1;
}
static inline int bound_i(int c, double er) {
    /* (int)GetBound(c): for c < 0 the value is NaN and x86 cvttsd2si yields INT_MIN */
    if (c < 0) return INT_MIN;
    return (int)bound_d(c, er);
}

typedef struct {
    int len, k, kcnt;
    int pa[ORC_MAX_LEN + 1], pt[ORC_MAX_LEN + 1]; /* prefix counts of 'A' and 'T' */
} polyctx;

static void poly_init(polyctx* pc, const char* seq, int len, int k) {
    pc->len = len; pc->k = k; pc->kcnt = len - k + 1;
    pc->pa[0] = pc->pt[0] = 0;
    for (int i = 0; i < len; ++i) {
        pc->pa[i + 1] = pc->pa[i] + (seq[i] == 'A');
        pc->pt[i + 1] = pc->pt[i] + (seq[i] == 'T');
    }
}
static inline int is_polya(const polyctx* pc, int j, int thr) { /* IsPolyA, ErrorCorrection.cpp:53-71 */
    int k = pc->k;
    return pc->pa[j + k] - pc->pa[j] >= k - thr || pc->pt[j + k] - pc->pt[j] >= k - thr;
}

int32_t orc_scan_counts(const orc_table* t, const char* seq, int32_t len, int32_t* counts) { /* ErrorCorrection.cpp:716-723 */
    const int k = t->k;
    const uint64_t mask = kmask(k);
    if (len < k) return 0;
    kmer x = {0, -1};
    int i, kcnt = 0;
    for (i = 0; i < k - 1; ++i) k_append(&x, seq[i], k, mask);
    for (; i < len; ++i, ++kcnt) {
        k_append(&x, seq[i], k, mask);
        counts[kcnt] = k_count(t, &x);
    }
    return kcnt;
}

/* common head of GetStrongTrustedThreshold (:1491-1527) and ErrorCorrection (:713-755): 1 = rejected */
static int rejected(const char* seq, int len, int k) {
    int nN = 0, nA = 0, nT = 0;
    for (int i = 0; i < len; ++i) {
        nN += seq[i] == 'N';
        nA += seq[i] == 'A';
        nT += seq[i] == 'T';
    }
    return nN > 5 || nA > len - k || nT > len - k;
}

/* sorted = ascending (loose-polyA k-mers as -1, others their count); ErrorCorrection.cpp:774-791 */
static int masked_sorted(const polyctx* pc, const int* counts, int* sorted) {
    int loose = pc->k / 2 > 7 ? pc->k / 2 : 7;
    for (int j = 0; j < pc->kcnt; ++j) sorted[j] = is_polya(pc, j, loose) ? -1 : counts[j];
    qsort(sorted, pc->kcnt, sizeof(int), cmp_int);
    int i;
    for (i = pc->kcnt - 1; i >= 1; --i)
        if (sorted[i] > 2 * sorted[i - 1] && sorted[i] > 10) break;
    return i; /* >= 1: index of the drop; <= 0: none */
}
static int median_positive(const int* sorted, int kcnt) { /* ErrorCorrection.cpp:811-814, 1559-1562 */
    int i;
    for (i = 0; i < kcnt; ++i)
        if (sorted[i] > 0) break;
    return sorted[(i + kcnt - 1) / 2];
}

int32_t orc_strong_threshold(const orc_table* t, const char* seq, int32_t len) { /* ErrorCorrection.cpp:1482-1565 */
    const int k = t->k;
    int counts[ORC_MAX_LEN], sorted[ORC_MAX_LEN];
    if (len < k) return -1;
    int kcnt = orc_scan_counts(t, seq, len, counts);
    if (rejected(seq, len, k)) return -1;
    polyctx pc;
    poly_init(&pc, seq, len, k);
    int i = masked_sorted(&pc, counts, sorted);
    if (i >= 1) return sorted[i];
    return median_positive(sorted, kcnt);
}

void orc_kmer_info(const orc_table* t, const char* seq, int32_t len, int32_t* l, int32_t* m, int32_t* h) { /* ErrorCorrection.cpp:1567-1602 */
    const int k = t->k;
    const uint64_t mask = kmask(k);
    int v[ORC_MAX_LEN], n = 0, i;
    kmer x = {0, -1};
    *l = *m = *h = 0;
    for (i = 0; i < len && i < k - 1; ++i) k_append(&x, seq[i], k, mask);
    for (; i < len; ++i) {
        k_append(&x, seq[i], k, mask);
        if (x.inv == -1) {
            int c = k_count(t, &x);
            v[n++] = c == 0 ? 1 : c;
        }
    }
    if (n == 0) return;
    qsort(v, n, sizeof(int), cmp_int);
    *l = v[0];
    *m = v[n / 2];
    *h = v[n - 1];
}

/* ------------------------------------------------------------------ path search
 * SearchPaths_Right / SearchPaths_Left, ErrorCorrection.cpp:201-442 / 444-678
 * (SURVEY A.5 / A.6).  State shared by the whole recursion lives in `srch`. */
typedef struct {
    const orc_table* tab;
    int k, len;
    uint64_t mask;
    double er;
    const char* seq;
    int* path;         /* the `fix` argument (iBuffer in the caller) */
    int* best;         /* the `bestFix` argument (fix[] in the caller) */
    int* top2;         /* segment's top2FixBottleNeck */
    const unsigned char* strong; /* isStrongTrusted */
    const unsigned char* polya;  /* isPolyAKmer (strict) */
    int max_fix, best_cnt, best_bn, trial;
} srch;

static inline char seq_at(const srch* s, int i) { return (i >= 0 && i < s->len) ? s->seq[i] : 0; }

static int infer_threshold(const srch* s, const kmer* kc, int dir, int upper) { /* InferPosThreshold, :144-173 */
    int mx = 0;
    for (int b = 0; b < 4; ++b) {
        kmer x = *kc;
        if (dir < 0) k_prepend(&x, "ACGT"[b], s->k, s->mask);
        else k_append(&x, "ACGT"[b], s->k, s->mask);
        int c = k_count(s->tab, &x);
        if (c > mx) mx = c;
    }
    int r = bound_i(mx, s->er);
    if (r < 1) r = 1;
    return (upper > r || upper <= 0) ? r : upper;
}

/* step 0 of both searches (:211-224 / :453-467): 1 = give up this node */
static int search_gate(srch* s, int fix_cnt) {
    if (s->trial > ORC_MAX_TRIAL) {
        if (s->max_fix > 2) {
            --s->max_fix;
            s->best_cnt = 0;
            s->best_bn = -1;
            s->trial = 0;
        } else {
            return 1;
        }
    }
    return fix_cnt > s->max_fix;
}

/* terminal bookkeeping (:243-284 / :483-523); lo..hi is the range of path[] to keep */
static void search_terminal(srch* s, int t, int fix_cnt, int bn, int lo, int hi) {
    if (bn < t) ++fix_cnt;
    if (fix_cnt < s->max_fix) {
        s->top2[0] = bn;
        s->top2[1] = -1;
    } else if (fix_cnt == s->max_fix) {
        if (bn > s->top2[0]) {
            s->top2[1] = s->top2[0];
            s->top2[0] = bn;
        } else if (bn > s->top2[1]) {
            s->top2[1] = bn;
        }
    }
    if (fix_cnt < s->max_fix || (fix_cnt == s->max_fix && bn > s->best_bn)) {
        if (fix_cnt < s->max_fix) s->trial = -(s->max_fix - fix_cnt + 1) * ORC_MAX_TRIAL;
        for (int i = lo; i <= hi; ++i) s->best[i] = s->path[i];
        s->max_fix = fix_cnt;
        s->best_bn = bn;
        s->best_cnt = 1;
    } else if (fix_cnt == s->max_fix && bn == s->best_bn) {
        s->best_cnt += 1;
    }
}

static void search_right(srch* s, int start, int to, int pos, int t, int fix_cnt, int bn, kmer kc) {
    const int k = s->k;
    if (search_gate(s, fix_cnt)) return;
    if (pos >= to) {
        search_terminal(s, t, fix_cnt, bn, start, pos - 1);
        return;
    }
    int thr = infer_threshold(s, &kc, +1, t);
    int ext = 0;
    const char here = seq_at(s, pos);
    if (nuc(here) != -1) { /* :294-339 */
        kmer x = kc;
        k_append(&x, here, k, s->mask);
        int c = k_count(s->tab, &x);
        if (c >= thr) {
            s->path[pos] = -1;
            ++ext;
            search_right(s, start, to, pos + 1, t, fix_cnt, c < bn ? c : bn, x);
        } else if (thr == 1 && t <= 2) { /* "accidentally below", :313-338 */
            int i = pos, steps = 0;
            for (; c < thr && steps < k; ++steps) {
                ++i;
                if (i >= to) break;
                k_append(&x, seq_at(s, i), k, s->mask);
                c = k_count(s->tab, &x);
            }
            if (steps < k && i < to) {
                for (int j = pos; j <= i; ++j) s->path[j] = -1;
                ++ext;
                search_right(s, start, to, i + 1, t, fix_cnt + 1, bn, x);
            }
        }
    }
    if (!s->strong[pos] && !s->polya[pos - k + 1]) { /* :343-372 */
        for (int b = 0; b < 4; ++b) {
            if ("ACGT"[b] == here) continue;
            kmer x = kc;
            k_append(&x, "ACGT"[b], k, s->mask);
            int c = k_count(s->tab, &x);
            if (c >= thr) {
                s->path[pos] = b;
                ++s->trial;
                ++ext;
                search_right(s, start, to, pos + 1, t, fix_cnt + (nuc(here) != -1), c < bn ? c : bn, x);
            }
        }
    }
    if (ext == 0) { /* jump through an unfixable region, :393-441 */
        kmer x = kc;
        int i;
        for (i = pos; i < to; ++i) {
            int th = infer_threshold(s, &x, +1, t);
            k_append(&x, seq_at(s, i), k, s->mask);
            int c = k_count(s->tab, &x);
            s->path[i] = -1;
            if (c >= th) break;
        }
        int kk = seq_at(s, i) ? i - pos - k + 1 : (i - pos) / 2;
        if (kk <= 0) kk = 1;
        if (i >= to) i -= 1;
        search_right(s, start, to, i + 1, t, fix_cnt + kk, bn, x);
    }
}

static void search_left(srch* s, int start, int to, int pos, int t, int fix_cnt, int bn, kmer kc) {
    const int k = s->k;
    if (search_gate(s, fix_cnt)) return;
    if (pos < to) {
        search_terminal(s, t, fix_cnt, bn, pos + 1, start);
        return;
    }
    int thr = infer_threshold(s, &kc, -1, t);
    int ext = 0;
    const char here = seq_at(s, pos);
    if (nuc(here) != -1) { /* :532-574 */
        kmer x = kc;
        k_prepend(&x, here, k, s->mask);
        int c = k_count(s->tab, &x);
        if (c >= thr) {
            s->path[pos] = -1;
            ++ext;
            search_left(s, start, to, pos - 1, thr, fix_cnt, c < bn ? c : bn, x);
        } else if (thr == 1 && t <= 2) { /* :550-573 */
            int i = pos, steps = 0;
            for (; c < thr && steps < k; ++steps) {
                --i;
                if (i < to) break;
                k_prepend(&x, seq_at(s, i), k, s->mask);
                c = k_count(s->tab, &x);
            }
            if (steps < k && i >= to) {
                for (int j = i; j <= pos; ++j) s->path[j] = -1;
                ++ext;
                search_left(s, start, to, i - 1, thr, fix_cnt + 1, bn, x);
            }
        }
    }
    if (!s->strong[pos] && !s->polya[pos]) { /* :577-608 */
        for (int b = 0; b < 4; ++b) {
            if ("ACGT"[b] == here) continue;
            kmer x = kc;
            k_prepend(&x, "ACGT"[b], k, s->mask);
            int c = k_count(s->tab, &x);
            if (c >= thr) {
                s->path[pos] = b;
                ++s->trial;
                ++ext;
                search_left(s, start, to, pos - 1, thr, fix_cnt + (nuc(here) != -1), c < bn ? c : bn, x);
            }
        }
    }
    if (ext == 0) { /* :629-677 */
        kmer x = kc;
        int i;
        for (i = pos; i >= to; --i) {
            thr = infer_threshold(s, &x, -1, t);
            k_prepend(&x, seq_at(s, i), k, s->mask);
            int c = k_count(s->tab, &x);
            s->path[i] = -1;
            if (c >= thr) break;
        }
        int kk = i >= 0 ? pos - i - k + 1 : pos - 1;
        if (kk <= 0) kk = 1;
        if (i <= to) ++i;
        search_left(s, start, to, i - 1, thr, fix_cnt + kk, bn, x);
    }
}

/* ------------------------------------------------------------------ per-read correction */
typedef struct {
    int from, to;
} span;
typedef struct {
    int from, to, lanchor, ranchor;
    int top2[2];
} segment;

int32_t orc_correct_read(const orc_table* tab, const orc_params* p, char* seq, const char* qual, int32_t len,
                         int32_t pair_threshold) { /* ErrorCorrection, ErrorCorrection.cpp:682-1480 */
    const int k = p->k, K = p->max_fix_per_k;
    const double er = p->error_rate;
    const int Q = p->bad_qual;
    const uint64_t mask = kmask(k);
    static __thread int counts[ORC_MAX_LEN], ibuf[ORC_MAX_LEN + 1], fix[ORC_MAX_LEN + 1];
    static __thread unsigned char strong_pos[ORC_MAX_LEN + 1], polya_strict[ORC_MAX_LEN + 1];
    static __thread span isl[ORC_MAX_LEN];
    static __thread segment seg[ORC_MAX_LEN];
    static __thread polyctx pc;

    if (len < k) return -1; /* :713 */
    const int kcnt = orc_scan_counts(tab, seq, len, counts);
    if (rejected(seq, len, k)) return -1; /* :735-755 */
    poly_init(&pc, seq, len, k);

    /* A.3 initial thresholds, :772-842 */
    int trust, strong;
    int flag = 0;
    int di = masked_sorted(&pc, counts, ibuf);
    if (di >= 1) {
        trust = bound_i(ibuf[di], er);
        strong = ibuf[di];
        if (strong >= 20 && ibuf[di - 1] == 2 && trust < 3) {
            flag = 1;
            trust = 3;
        }
    } else {
        strong = median_positive(ibuf, kcnt);
        trust = bound_i(strong, er);
    }
    if (pair_threshold >= 1 && strong > pair_threshold) {
        if (!flag || pair_threshold < 20) trust = bound_i(pair_threshold, er);
        strong = pair_threshold;
    }
    for (int j = 0; j < kcnt; ++j) polya_strict[j] = (unsigned char)is_polya(&pc, j, 2);
    if (trust < 2) trust = 2;

    int iter = 0, total_fix = 0, bad_segments = 0;
    for (;;) { /* A.4, :853-1291 */
        const int allowed = len;
        int unfixable = 0, force_next = 0;
        int nisl = 0, longest = -1, tstart = 0, tend = 0, run = 0, j;
        total_fix = 0;
        memset(strong_pos, 0, (size_t)len + 1);
        for (j = 0; j < len; ++j) ibuf[j] = -1;
        /* trusted k-mer runs, :870-931 */
        for (j = 0; j <= kcnt; ++j) {
            if (j < kcnt && counts[j] >= strong && !polya_strict[j]) {
                ++run;
                continue;
            }
            if (run > longest) {
                longest = run;
                tstart = j - run;
                tend = j - 1;
            }
            if (run >= 2) {
                isl[nisl].from = j - run;
                isl[nisl].to = j - 1;
                ++nisl;
            }
            run = 0;
        }
        /* island boundary adjustment, :934-965 */
        for (int i = 1; i < nisl; ++i) {
            if (isl[i].from > isl[i - 1].to + k) continue;
            int len1 = isl[i - 1].to - isl[i - 1].from, len2 = isl[i].to - isl[i].from;
            int overlap = isl[i - 1].to + k - isl[i].from;
            for (j = isl[i - 1].to + 1; j < isl[i].from; ++j)
                if (counts[j] <= 2 && counts[j] < trust) break;
            if (j >= isl[i].from) continue;
            if (overlap > 3) continue;
            if (len1 < len2) isl[i - 1].to -= overlap + 1;
            else isl[i].from += overlap + 1;
        }
        /* trusted positions and position islands, :968-1007 */
        for (int i = 0; i < nisl; ++i) {
            if (isl[i].from > isl[i].to) continue;
            for (j = isl[i].from; j <= isl[i].to + k - 1; ++j) strong_pos[j] = 1;
        }
        nisl = 0;
        for (j = 0; j < len;) {
            if (!strong_pos[j]) { ++j; continue; }
            int b = j;
            while (j < len && strong_pos[j]) ++j;
            isl[nisl].from = b;
            isl[nisl].to = j - 1;
            ++nisl;
        }
        if (nisl == 0) {
            isl[0].from = tstart;
            isl[0].to = tend + k - 1;
            nisl = 1;
        }
        /* segments, :1009-1046 */
        int nseg = 0;
        if (isl[0].from > 0) {
            segment* s = &seg[nseg++];
            s->from = 0; s->to = isl[0].from - 1;
            s->lanchor = 0; s->ranchor = isl[0].to - isl[0].from + 1;
        }
        for (int i = 0; i + 1 < nisl; ++i) {
            segment* s = &seg[nseg++];
            s->from = isl[i].to + 1; s->to = isl[i + 1].from - 1;
            s->lanchor = isl[i].to - isl[i].from + 1;
            s->ranchor = isl[i + 1].to - isl[i + 1].from + 1;
        }
        if (isl[nisl - 1].to < len - 1) {
            segment* s = &seg[nseg++];
            s->from = isl[nisl - 1].to + 1; s->to = len; /* [from, to) */
            s->lanchor = isl[nisl - 1].to - isl[nisl - 1].from + 1; s->ranchor = 0;
        }
        for (int i = 0; i < nseg; ++i) seg[i].top2[0] = seg[i].top2[1] = -1;

        if (longest == -1) return -1; /* :1107 (unreachable) */
        if (longest == kcnt) return 0; /* :1110 */
        for (j = 0; j < len; ++j) fix[j] = -1;
        bad_segments = 0;

        if (nseg > 0) { /* :1118-1230 */
            int best_fix_cnt = -1;
            int best_bn = ORC_INF;
            srch s;
            s.tab = tab; s.k = k; s.len = len; s.mask = mask; s.er = er; s.seq = seq;
            s.path = ibuf; s.best = fix; s.strong = strong_pos; s.polya = polya_strict;
            s.best_cnt = best_fix_cnt;
            for (int g = 0; g < nseg; ++g) {
                s.trial = 0;
                s.max_fix = (seg[g].to - seg[g].from + 1) * K / k * 2 + 1;
                if (s.max_fix < K) s.max_fix = K;
                s.best_bn = -1;
                s.top2 = seg[g].top2;
                kmer kc = {0, -1};
                if (seg[g].lanchor >= seg[g].ranchor) {
                    int extend = seg[g].to == len ? 0 : k - 1;
                    int a = seg[g].from - k;
                    for (int i = a; i < a + k; ++i) k_append(&kc, i < 0 ? 'A' : seq[i], k, mask);
                    search_right(&s, a + k, seg[g].to + extend, a + k, trust, 0, ORC_INF, kc);
                } else {
                    int extend = seg[g].from == 0 ? 0 : k - 1;
                    int a = seg[g].to + 1;
                    for (int i = a; i < a + k; ++i) k_append(&kc, seq_at(&s, i), k, mask);
                    search_left(&s, a - 1, seg[g].from - extend, a - 1, trust, 0, ORC_INF, kc);
                }
                if (s.best_bn == -1) {
                    ++bad_segments;
                    continue;
                }
                if (s.best_bn < best_bn) best_bn = s.best_bn;
                if (best_bn == -1) break;
                if (s.trial > ORC_MAX_TRIAL) return -1;
                total_fix += s.max_fix;
            }
            best_fix_cnt = s.best_cnt;
            if (best_bn != -1) {
                best_fix_cnt = 1;
                for (int g = 0; g < nseg; ++g)
                    if (seg[g].top2[1] >= best_bn) best_fix_cnt *= 2;
            }
            if (best_bn != -1 && iter == 0 && best_bn < bound_d(strong, er)) force_next = 1; /* :1195, FP64 compare */
            if (best_fix_cnt >= 2) return -1;
            else if (best_fix_cnt <= 0) unfixable = 1;
        }
        if (total_fix == 0 && force_next) return 0;
        if (total_fix > allowed) unfixable = 1;
        if (!unfixable && !force_next) break;
        if (trust < 10 && !force_next) return -1;

        /* re-threshold, :1247-1289 */
        {
            int loose = k / 2 > 7 ? k / 2 : 7, i, has_drop = 0;
            for (j = 0; j < kcnt; ++j) ibuf[j] = is_polya(&pc, j, loose) ? -1 : counts[j];
            qsort(ibuf, kcnt, sizeof(int), cmp_int);
            for (i = kcnt - 1; i >= 1; --i) {
                if (ibuf[i] > strong) continue;
                if ((ibuf[i] > 2 * ibuf[i - 1] && ibuf[i] > 10) || (ibuf[i - 1] == 0 && ibuf[i] >= 5)) {
                    has_drop = 1;
                    if (ibuf[i] < strong) break;
                }
            }
            if (!has_drop) break;
            ++iter;
            trust = bound_i(ibuf[i], er);
            strong = ibuf[i];
        }
    }

    /* A.7 post-filters, :1296-1466 */
    int* P = ibuf;
    int np = 0;
    for (int i = 0; i < len; ++i)
        if (seq[i] != 'N' && fix[i] != -1) P[np++] = i;
    for (int i = 1; i < np; ++i) { /* pair filter, :1314-1398 */
        if (qual[0] != '\0' && qual[P[i]] <= Q && qual[P[i - 1]] <= Q) continue;
        if (P[i] - P[i - 1] + 1 > k) continue;
        int j = P[i - 1] - k + 1;
        int min_single = ORC_INF, min_double = ORC_INF, ta = -1, tb = -1;
        if (j < 0) j = 0;
        for (; j < kcnt; ++j) {
            if (i >= 2 && j <= P[i - 2]) continue;
            if (i < np - 1 && j + k - 1 >= P[i + 1]) break;
            if (j + k - 1 >= P[i]) break;
            if (counts[j] < min_single) { min_single = counts[j]; ta = j; }
        }
        for (; j < kcnt; ++j) {
            if (i < np - 1 && j + k - 1 >= P[i + 1]) break;
            if (j > P[i - 1]) break;
            if (counts[j] < min_double) { min_double = counts[j]; tb = j; }
        }
        for (; j < kcnt; ++j) {
            if (i < np - 1 && j + k - 1 >= P[i + 1]) break;
            if (j > P[i]) break;
            if (counts[j] < min_single) { min_single = counts[j]; ta = j; }
        }
        if (min_single != ORC_INF && min_double != ORC_INF && counts[ta] > 1 && counts[tb] > 1 &&
            counts[ta] > counts[tb] / 2 && counts[ta] < 2 * counts[tb]) {
            fix[P[i]] = -1;
            fix[P[i - 1]] = -1;
            j = i - 2;
            while (j >= 0 && P[j + 1] - P[j] + 1 <= k) {
                fix[P[j]] = -1;
                --j;
            }
            while (i + 1 < np && P[i + 1] - P[i] + 1 <= k) {
                fix[P[i + 1]] = -1;
                ++i;
            }
        }
    }
    if (total_fix > 3 && len > 10) { /* edge filter, :1407-1430 */
        int n = 0;
        for (int i = 0; i < 10; ++i) n += fix[i] != -1 && seq[i] != 'N' && qual[i] > Q;
        if (n >= 2)
            for (int i = 0; i < 10; ++i)
                if (seq[i] != 'N') fix[i] = -1;
        n = 0;
        for (int i = len - 10; i < len; ++i) n += fix[i] != -1 && seq[i] != 'N' && qual[i] > Q;
        if (n >= 3)
            for (int i = len - 10; i < len; ++i)
                if (seq[i] != 'N') fix[i] = -1;
    }
    if (total_fix >= K) { /* density filter, :1432-1455; weights are halves, so 2x integers are exact */
        static __thread int w2[ORC_MAX_LEN + 1];
        w2[0] = 0;
        for (int i = 0; i < len; ++i)
            w2[i + 1] = w2[i] + ((seq[i] != 'N' && fix[i] != -1) ? (qual[i] > Q ? 2 : 1) : 0);
        for (int i = 0; i < kcnt; ++i)
            if (w2[i + k] - w2[i] > 2 * K) return -1;
    }
    int ret = 0; /* apply, :1468-1479 */
    for (int i = 0; i < len; ++i)
        if (fix[i] != -1) {
            seq[i] = "ACGT"[fix[i]];
            ++ret;
        }
    if (ret == 0 && bad_segments > 0) return -1;
    return ret;
}

/* ------------------------------------------------------------------ batch worker */
typedef struct {
    const orc_table* t;
    const orc_params* p;
    int mode;
    int64_t n;
    const char* seq_in;
    const char* qual;
    const uint32_t* off;
    char* seq_out;
    int32_t *corr, *l, *m, *h;
    int64_t next;
    pthread_mutex_t lock;
} batch_job;

static void do_record(batch_job* b, int64_t r, int pair_t) {
    int len = (int)(b->off[r + 1] - b->off[r]);
    char* s = b->seq_out + b->off[r];
    char q0[2] = {0, 0};
    const char* q = len > 0 ? b->qual + b->off[r] : q0;
    b->corr[r] = orc_correct_read(b->t, b->p, s, q, len, pair_t);
    orc_kmer_info(b->t, s, len, &b->l[r], &b->m[r], &b->h[r]);
}

static void* batch_thread(void* arg) { /* ErrorCorrection_Thread, ErrorCorrection.cpp:73-136 */
    batch_job* b = (batch_job*)arg;
    const int inc = b->mode == ORC_INTERLEAVED ? 2 : 1;
    for (;;) {
        pthread_mutex_lock(&b->lock);
        int64_t r = b->next;
        b->next += 64 * inc;
        pthread_mutex_unlock(&b->lock);
        if (r >= b->n) break;
        int64_t hi = r + 64 * inc < b->n ? r + 64 * inc : b->n;
        for (; r < hi; r += inc) {
            int t = -1;
            if (b->mode == ORC_INTERLEAVED) {
                int t1 = orc_strong_threshold(b->t, b->seq_out + b->off[r], (int)(b->off[r + 1] - b->off[r]));
                int t2 = orc_strong_threshold(b->t, b->seq_out + b->off[r + 1], (int)(b->off[r + 2] - b->off[r + 1]));
                t = t1 < t2 ? t1 : t2;
            }
            do_record(b, r, t);
            if (b->mode == ORC_INTERLEAVED) do_record(b, r + 1, t);
        }
    }
    return NULL;
}

void orc_correct_batch(const orc_table* t, const orc_params* p, int mode, int64_t n, const char* seq_in,
                       const char* qual, const uint32_t* off, char* seq_out, int32_t* correction, int32_t* l,
                       int32_t* m, int32_t* h, int threads) {
    batch_job b;
    b.t = t; b.p = p; b.mode = mode; b.n = n; b.seq_in = seq_in; b.qual = qual; b.off = off;
    b.seq_out = seq_out; b.corr = correction; b.l = l; b.m = m; b.h = h; b.next = 0;
    if (seq_out != seq_in) memcpy(seq_out, seq_in, off[n]);
    pthread_mutex_init(&b.lock, NULL);
    if (threads <= 1) {
        batch_thread(&b);
    } else {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
        for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, batch_thread, &b);
        for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&b.lock);
}
