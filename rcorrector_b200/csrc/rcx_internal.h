// rcx_internal.h — host-side declarations shared by the kernel translation units and the C-ABI layer.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "rcx_device.cuh"

// per-read flags written by the scan kernel
#define RCX_F_DROP 1      // a "drop" was found in the sorted counts (ErrorCorrection.cpp:787-794)
#define RCX_F_PREV2 2     // the entry below the drop equals 2 (ErrorCorrection.cpp:800)
#define RCX_F_REJECT 4    // too short / too many N / A / T (ErrorCorrection.cpp:713,735-755)

// device counters (unsigned long long each)
enum { RCX_C_SCAN_PROBES = 0, RCX_C_SEARCH_PROBES, RCX_C_RESCAN_PROBES, RCX_C_SEARCH_NODES, RCX_C_WORK,
       RCX_C_NWORK, RCX_C_WORK2,               // work-list length, finalize kernel's fetch counter
       RCX_C_NWORK2, RCX_C_WORKX0, RCX_C_WORKX1, RCX_C_WORKX2, RCX_C_WORKX3, RCX_C_UNSUPPORTED,  // second-pass list, per-kernel fetch counters
       RCX_C_NVH,                                // work items with nearly every k-window untrusted (counted always; listed in vhlist[] in vh mode)
       RCX_C_WORKV,                              // search kernel, vh mode: next batch of 32 listed items
       RCX_C_WORKB,                              // search kernel: next batch of 32 light items
       RCX_C_NWORKH,                             // HEAVY work items (long untrusted spans): slots n-1, n-2, ... of the work list, searched first
       RCX_C_RESCAN_WINDOWS,                     // k-windows the finalize kernels looked up again (two per request on the thread path)
       RCX_C_NEDITS,                             // edits written by the gather kernel of the compact host interface
       RCX_C_DIAG0, RCX_C_DIAG_END = RCX_C_DIAG0 + 13,  // diagnostic builds (-DRCX_DIAG): search-walk length against untrusted-window count
       RCX_C_SINK,                               // scratch word of the random-read micro-benchmark
       RCX_C_COUNT = RCX_C_SINK + 4 };

struct RcxScanArgs {
    RcxTable t;
    const char* seq;
    const uint32_t* off;
    long long n;
    int max_len;
    int32_t* counts;   // [off[n]] counts[off[r] + j]
    int32_t* thr;      // [n] GetStrongTrustedThreshold
    uint8_t* flags;    // [n]
    unsigned long long* counters;
    // fused triage (triage != 0): clean / rejected reads get correction + l/m/h here, the others go to the work list
    int triage;
    int mode;          // rcx_mode: in interleaved mode a warp owns both mates (pair threshold)
    double error_rate;
    int32_t* correction;
    int32_t* l;
    int32_t* m;
    int32_t* h;
    uint32_t* worklist;
    // fused first pass of the A.4 preparation (thread-per-read kernel only)
    unsigned long long total_bases;
    int max_fix_per_k;
    int32_t* state;
    unsigned char* posbuf;
    signed char* fixbuf;
    int compact_fix;   // edit-list hand-over to the thread-per-item finalize: only the count word of a read's list is cleared
    unsigned short* fixlist;  // [off[n]] 16-bit view of the fix buffer in that mode: fixlist[off[r]] = number of edits of read r, then pos | base << 8
    int heavy_windows; // work items with at least this many untrusted k-windows go to the heavy end of the work list (0: none)
    int vheavy_windows;// ... and with at least this many are counted as VERY heavy (RCX_C_NVH; 0: none)
    int vh_mode;       // vh mode: the very heavy items are also listed in vhlist[] (status RCX_ST_SEARCH_VH) for the search kernel's batches
    uint32_t* vhlist;  // [n] work-list slots of the very heavy items
};
#define RCX_TS_MAX_KCNT 232  // reads with more k-windows use the warp-per-read kernels

struct RcxCorrectArgs {
    RcxTable t;
    char* seq;             // corrected in place
    const char* qual;
    const uint32_t* off;
    long long n;
    int max_len;
    int mode;              // rcx_mode
    int max_fix_per_k;
    double error_rate;
    int bad_qual;
    const int32_t* counts;
    const int32_t* thr;
    const uint8_t* flags;
    int32_t* correction;
    int32_t* l;
    int32_t* m;
    int32_t* h;
    unsigned long long* counters;
    // hand-over between the triage / search / finalize kernels
    uint32_t* worklist;    // [n] indices of reads that need the search
    uint32_t* worklist2;   // [n] work items that asked for a second pass of A.4
    uint32_t* vhlist;      // [n] vh mode: slots of the very heavy work items (status RCX_ST_SEARCH_VH until searched), taken first
    int vh_mode;           // the search kernel's first pass starts with warp batches over vhlist[]
    int32_t* state;        // [n * rcx_state_words()] per work item: header, masks, segment list
    int pass;              // 0 / 1 (set by the launcher)
    signed char* fixbuf;   // [off[n]] fix[] (bestFix of the searches), -1 = no change
    unsigned char* posbuf; // [off[n]] per base: code | not-ACGT << 2 | substitution allowed (Right search) << 3 | (Left search) << 4
    unsigned int* dfs_stack; // search scratch: explicit recursion stacks, interleaved by lane
    signed char* dfs_path;   // search scratch: path[] per thread, interleaved by lane
    int dfs_depth, dfs_lpad;
    int dfs_batch;           // search kernel: light items of the first pass are taken 32 consecutive slots per warp
    int compact_fix;         // the fix buffer holds the search kernel's edit list per read (thread-per-item finalize) instead of a dense fix[]
    unsigned short* fixlist; // [off[n]] that list: fixlist[off[r]] = count, then pos | base << 8 (at most len - 1 edits: one position is always trusted)
    // compact host interface (rcx_correct_batch_compact): quality as one bit per base (bit (i & 7) of byte i >> 3 for the
    // batch's base i: qual[i] > badQualityThreshold) instead of `qual`; has_qual = "qual[0] != 0" of every read (0 for FASTA)
    const unsigned char* qbits;
    int has_qual;
    int emit_edits;          // the finalize kernels leave the applied edits in fixbuf for the gather kernel
};
// qual[i] > badQualityThreshold for base i of the read at offset o (ErrorCorrection.cpp:1314-1316,1407-1455 read quality only so)
__device__ __forceinline__ bool rcx_qual_high(const RcxCorrectArgs& a, uint32_t o, int i) {
    if (a.qbits) { const unsigned long long x = (unsigned long long)o + (unsigned int)i; return (a.qbits[x >> 3] >> (x & 7)) & 1; }
    if (!a.qual) return false;  // compact interface without quality (FASTA): every qual[i] is 0 and so is the threshold
    return (int)a.qual[o + i] > a.bad_qual;
}
__device__ __forceinline__ bool rcx_has_qual(const RcxCorrectArgs& a, uint32_t o) {
    return a.qbits ? a.has_qual != 0 : (a.qual ? a.qual[o] != 0 : false);
}

// per-work-item state in HBM (int32 words): 8-word header, then ST_SEGW words per segment
#define ST_STATUS 0
#define ST_TOTALFIX 1  // in (first pass): rcx_force_bound(strong) for the search kernel; out: totalFix
#define ST_BADSEG 2
#define ST_STRONG 3
#define ST_TRUST 4
#define ST_NSEG 5
#define ST_HDR 8
#define ST_SEGW 6   // start | to << 16, right | maxFixCnt << 1, seed code lo, hi, seed invalidPos, top2FixBottleNeck[1]
#define RCX_ST_SEARCH (-4)       // waiting for the search kernel
#define RCX_ST_SEARCH_VH (-5)    // the same, listed in vhlist[] (vh mode: the sweep over the heavy slots passes it by)
#define RCX_ST_RETHRESHOLD (-3)  // first pass asked for a second one (forceNextRound)
#define RCX_ST_POSTFILTER (-2)   // searches done: post-filters decide
#define ST_OFF 6    // off[r] and length of the read: the search kernel gets them with the header instead of chasing worklist -> off[]
#define ST_LEN 7
// Work-list slots.  The thread-per-read scan kernel files the items whose search is likely to be long (many untrusted
// k-windows) from the END of the work list (slots n-1, n-2, ...; RCX_C_NWORKH of them) and the others from the front
// (RCX_C_NWORK); the search kernel takes the heavy ones first, so that the long walks do not start when the launch is
// nearly over.  Item number w of the RCX_C_NWORK + RCX_C_NWORKH items -> slot (= index into worklist[] and state[]).
__host__ __device__ inline long long rcx_work_slot(long long w, long long nheavy, long long n) { return w < nheavy ? n - 1 - w : w - nheavy; }
// multiple of 8 words: every item header is 32-byte aligned
__host__ __device__ inline int rcx_state_stride(int max_len, int k) { return (ST_HDR + ST_SEGW * (max_len / k + 4) + 7) & ~7; }

// table (rcx_table.cu)
cudaError_t rcx_launch_table_init(RcxEntry* e, unsigned long long nlines, cudaStream_t s);
cudaError_t rcx_launch_table_claim(RcxEntry* e, unsigned long long nlines, int k, const unsigned long long* codes, const int32_t* counts,
                                   unsigned long long n, int home_only, unsigned long long* stat, unsigned long long fresh_limit,
                                   cudaStream_t s);
cudaError_t rcx_launch_table_fill(RcxEntry* e, unsigned long long nlines, int k, const unsigned long long* codes, const int32_t* counts,
                                  unsigned long long n, unsigned long long first_index, unsigned int* win, unsigned long long part_lo,
                                  unsigned long long part_hi, int write, unsigned long long* stat, unsigned long long* big_keys,
                                  unsigned int* big_counts, unsigned long long big_cap, cudaStream_t s);
cudaError_t rcx_launch_table_lookup(RcxTable t, const unsigned long long* codes, int32_t* out, unsigned long long n,
                                    cudaStream_t s);
cudaError_t rcx_launch_table_lookup4(RcxTable t, const unsigned long long* codes, int32_t* out4, unsigned long long n, cudaStream_t s);
cudaError_t rcx_launch_random_read(RcxTable t, int bytes, int grid, unsigned long long iters, unsigned long long* sink, cudaStream_t s);
// scan (rcx_scan.cu)
size_t rcx_scan_smem_per_warp(int max_len, int k);
cudaError_t rcx_launch_scan(const RcxScanArgs& a, int sm_count, cudaStream_t s);
// thread-per-read scan + thresholds + triage + first-pass preparation (rcx_scan_thread.cu)
bool rcx_scan_thread_supported(int max_len, int k, const void* seq);
cudaError_t rcx_launch_scan_thread(const RcxScanArgs& a, int sm_count, cudaStream_t s, bool row24 = false);  // row24: experimental, rcx_scan_thread24.cu
cudaError_t rcx_launch_scan_thread24(const RcxScanArgs& a, int sm_count, cudaStream_t s);  // rcx_scan_thread24.cu
// correction (rcx_correct.cu)
size_t rcx_correct_smem_per_warp(int max_len, int k);
size_t rcx_state_words(int max_len, int k);
// kev: 7 events recorded before triage, prep0, dfs0, prep1, dfs1, finalize and after finalize (may be null)
cudaError_t rcx_launch_correct(const RcxCorrectArgs& a, int sm_count, int dfs_grid, cudaStream_t s, cudaEvent_t* kev, bool prep0_done);
// thread-per-item post-filters + apply + l/m/h (rcx_finalize_thread.cu)
bool rcx_finalize_thread_supported(int max_len, int k);
cudaError_t rcx_launch_finalize_thread(const RcxCorrectArgs& a, int sm_count, cudaStream_t s);
// compact host interface: applied edits of the changed reads gathered in record order (rcx_edits.cu)
size_t rcx_edits_temp_bytes(long long n);
cudaError_t rcx_launch_edits_gather(const int32_t* corr, const uint32_t* off, long long n, const signed char* fixbuf, const unsigned short* fixlist,
                                    uint32_t* scan, void* temp, size_t temp_bytes, uint16_t* edits, unsigned long long* counters, cudaStream_t s);
// k-mer counting, SURVEY 8f N3 (rcx_count.cu)
cudaError_t rcx_launch_count_init(unsigned long long* keys, unsigned int* counts, unsigned long long cap, cudaStream_t s);
cudaError_t rcx_launch_count_reads(const char* seq, const uint32_t* off, long long n, int k, unsigned long long* keys, unsigned int* counts,
                                   unsigned long long cap, unsigned long long* stat, int sm_count, cudaStream_t s);
cudaError_t rcx_launch_count_rehash(const unsigned long long* okeys, const unsigned int* ocounts, unsigned long long ocap, unsigned long long* keys,
                                    unsigned int* counts, unsigned long long cap, cudaStream_t s);
size_t rcx_count_select_temp_bytes(unsigned long long n);
cudaError_t rcx_launch_count_select(const unsigned long long* keys, const unsigned int* counts, unsigned long long lo, unsigned long long n,
                                    unsigned int min_count, unsigned int* flags, unsigned int* pos, void* temp, size_t temp_bytes,
                                    unsigned long long base, unsigned long long* out_codes, int32_t* out_counts, cudaStream_t s);
// search (rcx_dfs.cu)
void rcx_dfs_geometry(int sm_count, int max_len, int k, int max_fix_per_k, int* grid, int* depth, int* lpad);
size_t rcx_dfs_stack_words(int grid, int depth);
size_t rcx_dfs_path_bytes(int grid, int lpad);
cudaError_t rcx_launch_dfs(const RcxCorrectArgs& a, int grid, cudaStream_t s);
