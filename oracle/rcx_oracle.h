/* rcx_oracle.h — CPU restatement of Rcorrector's stage-3 hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 *
 * Parity status: PINNED.  The reference ships no golden vectors,
 * so the restatement is pinned against outputs of the reference itself,
 * compiled untouched from /root/reference into oracle/_ref/rcorrector by
 * oracle/Makefile (tests/test_oracle_vs_ref.py, tests/golden/).
 */
#ifndef RCX_ORACLE_H
#define RCX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_table orc_table;

/* The four live scalars of the hot path (main.cpp:17-30). */
typedef struct {
    int32_t k;              /* kmerLength                         */
    int32_t max_fix_per_k;  /* MAX_FIX_PER_K  (-maxcorK, default 4) */
    double error_rate;      /* ERROR_RATE                          */
    int32_t bad_qual;       /* badQualityThreshold as a char value */
} orc_params;

enum { ORC_SINGLE = 0, ORC_INTERLEAVED = 1 };

orc_table* orc_table_new(int k);
void orc_table_free(orc_table* t);
uint64_t orc_table_size(const orc_table* t);
/* Store::Put over (forward code, count) records, in order: canonicalise,
 * last writer wins (Store.h:51-57).  Records with count <= 1 are skipped
 * like main.cpp:299. */
void orc_table_put(orc_table* t, const uint64_t* codes, const int32_t* counts, uint64_t n);
/* main.cpp:294-308: parse a Jellyfish text dump; returns the "Stored %d kmers" number, -1 on I/O error. */
int64_t orc_table_load_dump(orc_table* t, const char* path);
/* Store::GetCount on a valid forward code. */
int32_t orc_table_count(const orc_table* t, uint64_t fwd_code);

/* main.cpp:310-358 */
double orc_estimate_error_rate(const orc_table* t, const char* dump_path, double wk);
/* main.cpp:88-128 over (qual[0], qual[len-1]) of the first min(n,1e6) records */
int32_t orc_bad_quality(const char* qual, const uint32_t* off, int64_t n);

/* ErrorCorrection.cpp:716-723 — counts[] of one read; returns kcnt (0 if shorter than k). */
int32_t orc_scan_counts(const orc_table* t, const char* seq, int32_t len, int32_t* counts);
/* ErrorCorrection.cpp:1482-1565 */
int32_t orc_strong_threshold(const orc_table* t, const char* seq, int32_t len);
/* ErrorCorrection.cpp:682-1480; seq is corrected in place; returns -1 / 0 / n. */
int32_t orc_correct_read(const orc_table* t, const orc_params* p, char* seq, const char* qual, int32_t len,
                         int32_t pair_threshold);
/* ErrorCorrection.cpp:1567-1602 */
void orc_kmer_info(const orc_table* t, const char* seq, int32_t len, int32_t* l, int32_t* m, int32_t* h);

/* ErrorCorrection.cpp:73-136 run to completion over one batch.  Reads are
 * concatenated without separators; read r occupies [off[r], off[r+1]).  In
 * ORC_INTERLEAVED mode mates are records 2i and 2i+1 (n must be even). */
void orc_correct_batch(const orc_table* t, const orc_params* p, int mode, int64_t n, const char* seq_in,
                       const char* qual, const uint32_t* off, char* seq_out, int32_t* correction, int32_t* l,
                       int32_t* m, int32_t* h, int threads);

#ifdef __cplusplus
}
#endif
#endif
