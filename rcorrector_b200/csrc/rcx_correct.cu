// rcx_correct.cu — K2: per-read correction, one warp owns one read.
//
// Restates ErrorCorrection (ErrorCorrection.cpp:682-1480), SearchPaths_Right /
// SearchPaths_Left (:201-442 / :444-678), InferPosThreshold (:144-173) and
// GetKmerInformation (:1567-1602) as a warp-cooperative state machine:
//
//  * every lane keeps the (warp-uniform) scalar state of the algorithm in
//    registers, so control flow never diverges; lanes split the data-parallel
//    parts: bit-plane construction, trusted-k-mer ballots, range fills, the
//    four one-base-extension probes of a search node (lanes 0..3, one 32-byte
//    sector each), best-path copies, filters and the final re-scan;
//  * the recursion of the reference becomes an explicit stack of 48-byte
//    frames in shared memory (depth <= kcnt + 2); children are visited in the
//    reference's order (keep-base / rescue, A, C, G, T, jump) and all state
//    the reference shares by reference across the recursion (maxFixCnt,
//    bestFixCnt, bestFixBottleNeck, top2, trialCnt, the never-rolled-back
//    path[] / fix[] arrays) is shared here in the same way;
//  * the four extension k-mers of a node serve InferPosThreshold, the
//    keep-base test and the three alternatives at once (the CPU probes the
//    same k-mers 8 times).
//
// counts[] and the per-read GetStrongTrustedThreshold come from the scan
// kernel (rcx_scan.cu); the pair threshold is min over the two mates
// (ErrorCorrection.cpp:96-106).
#include "rcx_internal.h"

#define COR_WARPS 4
#define FRAME_WORDS 12

enum { ST_ENTER = 0, ST_KEEP = 1, ST_ALT0 = 2, ST_JUMP = 6, ST_DONE = 7 };

struct CorLayout {
    int pkw, mw, lpad, n2, nseg, depth;
    int words;  // total 32-bit words per warp
};

__host__ __device__ inline CorLayout cor_layout(int max_len, int k) {
    CorLayout L;
    int kc = max_len - k + 1;
    if (kc < 1) kc = 1;
    L.pkw = rcx_pk_words(max_len);
    L.mw = rcx_mask_words(max_len);
    L.lpad = ((max_len + 31) / 32) * 32 + 32;
    L.n2 = rcx_pow2_at_least(kc);
    L.nseg = max_len / k + 4;
    L.depth = kc + 3;
    int un = L.depth * FRAME_WORDS;
    if (un < L.n2) un = L.n2;
    L.words = L.pkw + 10 * L.mw + L.lpad /*counts*/ + 3 * L.lpad / 4 /*path, fix, posinfo as int8*/ + L.nseg + un;
    return L;
}

size_t rcx_correct_smem_per_warp(int max_len, int k) { return sizeof(unsigned int) * (size_t)cor_layout(max_len, k).words; }

struct WarpMem {
    RcxPlanes P;
    unsigned int* strong;   // isStrongTrusted, bit per base
    unsigned int* island;   // islands used to derive segments (== strong unless the fallback island is used)
    unsigned int* pstrict;  // isPolyAKmer (threshold 2), bit per k-mer
    unsigned int* trusted;  // trusted k-mers of this pass
    unsigned int* fixm;     // scratch masks for the post-filters
    int* counts;
    signed char* path;
    signed char* fix;
    unsigned char* posinfo;  // per base: code (bits 0-1) | not-ACGT (bit 2) | substitution allowed in a Right search (bit 3) / Left search (bit 4)
    int* segtop;            // top2FixBottleNeck[1] per segment
    int* sbuf;              // sort buffer (aliases the search stack)
    unsigned int* stack;
};

// state shared by the whole search of one segment (the by-reference parameters of SearchPaths_*)
struct SearchState {
    int max_fix, best_cnt, best_bn, trial, top2a, top2b;
};

__device__ __noinline__ void warp_sort_ool(int* s, int n2, int lane) { rcx_warp_sort(s, n2, lane); }
__device__ __noinline__ void build_planes_ool(const RcxPlanes& P, const char* __restrict__ seq, int len, int max_len, int lane, int& nN,
                                              int& nA, int& nT) {
    rcx_build_planes(P, seq, len, max_len, lane, nN, nA, nT);
}

__device__ __forceinline__ int getbit(const unsigned int* m, int i) { return (m[i >> 5] >> (i & 31)) & 1; }

// first index in [from, limit) whose bit equals `want`, or `limit`; warp-cooperative, all lanes must call
__device__ __noinline__ int next_bit(const unsigned int* m, int from, int limit, bool want, int lane) {
    int best = limit;
    if (from < limit) {
        const int w0 = from >> 5;
        #pragma unroll 1
        for (int w = w0 + lane; (w << 5) < limit; w += 32) {
            unsigned int x = want ? m[w] : ~m[w];
            if (w == w0) x &= 0xffffffffu << (from & 31);
            if (x) {
                const int idx = (w << 5) + __ffs(x) - 1;
                if (idx < best) best = idx;
                break;
            }
        }
    }
    best = __reduce_min_sync(RCX_FULL, best);
    return best < limit ? best : limit;
}

// set bits [a, b] (inclusive) of a mask plane; warp-cooperative
__device__ __noinline__ void set_range(unsigned int* m, int a, int b, int mw, int lane) {
    if (a > b) return;
    #pragma unroll 1
    for (int w = (a >> 5) + lane; w <= (b >> 5) && w < mw; w += 32) {
        const int lo = a > (w << 5) ? a - (w << 5) : 0;
        const int hi = b < (w << 5) + 31 ? b - (w << 5) : 31;
        const unsigned int bits = (hi - lo == 31) ? 0xffffffffu : (((1u << (hi - lo + 1)) - 1) << lo);
        m[w] |= bits;
    }
}

// window code / state for a seed k-mer built by Restart + k Appends of seq[j .. j+k) (j may be -1)
__device__ __forceinline__ RcxKmer seed_kmer(const RcxPlanes& P, int j, int k, unsigned int kmask32) {
    RcxKmer x;
    if (j >= 0) {
        x.code = rcx_window(P.pk, j, k);
        const unsigned int m = rcx_bits(P.inv, j) & kmask32;
        x.inv = m ? (k - 1) - (31 - __clz(m)) : -1;
    } else {
        // seq[-1] is read by the reference (ErrorCorrection.cpp:1139-1142 with tend = -1); whatever it holds is shifted
        // out / reset by the first Append of the search, so any value gives the same result.
        x.code = rcx_window(P.pk, 0, k) >> 2;
        const unsigned int m = rcx_bits(P.inv, 0) & (kmask32 >> 1);
        x.inv = m ? (k - 2) - (31 - __clz(m)) : -1;
    }
    return x;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
    #pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RCX_FULL, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------ kernels
// The correction is split into three kernels so that each one's hot code stays inside the SM's instruction cache
// (a single fused kernel was 130 KB of SASS and spent 2/3 of its issue slots in no_instruction stalls, profiles/):
//   rcx_triage_kernel   every read : thresholds, "all k-mers trusted?" -> result 0 / -1 and l/m/h at once,
//                                    otherwise the read index goes to a work list
//   rcx_search_kernel   work list  : A.4 iteration loop (islands, segments, ordered DFS, re-threshold)
//   rcx_finalize_kernel work list  : A.7 post-filters, apply, l/m/h re-scan
// State between search and finalize travels through HBM: status + totalFix + badSegmentCnt (12 B / read) and the
// fix[] array (1 B / base, only for reads that reach the post-filters).

__device__ __forceinline__ void setup_warpmem(WarpMem& M, unsigned int* smem, int warp, const CorLayout& L) {
    const int mw = L.mw;
    unsigned int* p = smem + (size_t)warp * L.words;
    M.P.pk = p; p += L.pkw;
    M.P.inv = p; p += mw;
    M.P.mA = p; p += mw;
    M.P.mT = p; p += mw;
    M.P.mN = p; p += mw;
    M.strong = p; p += mw;
    M.island = p; p += mw;
    M.pstrict = p; p += mw;
    M.trusted = p; p += mw;
    M.fixm = p; p += 2 * mw;
    M.counts = reinterpret_cast<int*>(p); p += L.lpad;
    M.path = reinterpret_cast<signed char*>(p);
    M.fix = M.path + L.lpad;
    M.posinfo = reinterpret_cast<unsigned char*>(M.fix + L.lpad);
    p += 3 * L.lpad / 4;
    M.segtop = reinterpret_cast<int*>(p); p += L.nseg;
    M.sbuf = reinterpret_cast<int*>(p);
    M.stack = p;
}

// k-th smallest (0-based) of n ints in shared memory, all >= 1: bisection on the value, no data movement
__device__ __noinline__ int warp_select(const int* v, int n, int kth, int lo, int hi, int lane) {
    #pragma unroll 1
    while (lo < hi) {
        const int mid = lo + (int)(((unsigned int)hi - (unsigned int)lo) >> 1);
        int c = 0;
        #pragma unroll 1
        for (int j = lane; j < n; j += 32) c += v[j] <= mid;
        c = __reduce_add_sync(RCX_FULL, c);
        if (c > kth) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// GetKmerInformation (:1567-1602) from counts[] and the validity plane: l = min, m = element [n/2], h = max of the
// counts of the VALID k-mers with 0 -> 1; all 0 when there is none.
__device__ __noinline__ void kmer_info(const WarpMem& M, int kcnt, unsigned int kmask32, int lane, int& l, int& m, int& h) {
    int nvalid = 0, mn = 0x7fffffff, mx = 0;
    #pragma unroll 1
    for (int j = lane; j < kcnt; j += 32) {
        int v = 0x7fffffff;
        if ((rcx_bits(M.P.inv, j) & kmask32) == 0) {
            v = M.counts[j];
            if (v == 0) v = 1;
            ++nvalid;
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        M.sbuf[j] = v;
    }
    nvalid = __reduce_add_sync(RCX_FULL, nvalid);
    mn = __reduce_min_sync(RCX_FULL, mn);
    mx = __reduce_max_sync(RCX_FULL, mx);
    __syncwarp();
    if (nvalid == 0) { l = m = h = 0; return; }
    l = mn; h = mx;
    m = warp_select(M.sbuf, kcnt, nvalid / 2, mn, mx, lane);  // invalid entries are INT_MAX: never counted below hi
    __syncwarp();
}

// thresholds of A.3 (:772-842) from the scan kernel's per-read summary
__device__ __forceinline__ void initial_thresholds(const RcxCorrectArgs& a, long long r, uint8_t fl, double er, int& strong, int& trust) {
    int pair_t = -1;
    if (a.mode == 1) {
        const int t1 = a.thr[r], t2 = a.thr[r ^ 1];
        pair_t = t1 < t2 ? t1 : t2;
    }
    strong = a.thr[r];
    bool flag = false;
    trust = rcx_bound_i_ool(strong, er);
    if ((fl & RCX_F_DROP) && strong >= 20 && (fl & RCX_F_PREV2) && trust < 3) { flag = true; trust = 3; }
    if (pair_t >= 1 && strong > pair_t) {
        if (!flag || pair_t < 20) trust = rcx_bound_i_ool(pair_t, er);
        strong = pair_t;
    }
    if (trust < 2) trust = 2;
}

// work-item fetch shared by the persistent warp-per-read kernels: pass 0 walks the triage work list, pass 1 the
// (short) list of items whose first pass asked for a second one (forceNextRound, ErrorCorrection.cpp:1195-1211; at
// most two passes can happen because forceNextRound needs iter == 0 and `unfixable` is never set)
__device__ __forceinline__ bool fetch_item(const RcxCorrectArgs& a, int counter, int lane, long long& item) {
    const long long n = (long long)a.counters[a.pass == 0 ? RCX_C_NWORK : RCX_C_NWORK2];
    long long w = 0;
    if (lane == 0) w = (long long)atomicAdd(&a.counters[counter], 1ull);
    w = __shfl_sync(RCX_FULL, w, 0);
    if (w >= n) return false;
    item = a.pass == 0 ? w : (long long)a.worklist2[w];
    return true;
}

// A.4 steps 1-4 (:863-1046): trusted k-mers -> islands -> trusted positions -> segments (+ their seed k-mers and
// maxFixCnt), written to the item state for the search kernel; on the second pass the re-threshold of :1247-1289 first.
__global__ void __launch_bounds__(COR_WARPS * 32) rcx_prep_kernel(RcxCorrectArgs a) {
    extern __shared__ unsigned int smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.t.k, K = a.max_fix_per_k;
    const CorLayout L = cor_layout(a.max_len, k);
    const int mw = L.mw;
    WarpMem M;
    setup_warpmem(M, smem, warp, L);
    const unsigned int kmask32 = k >= 32 ? 0xffffffffu : ((1u << k) - 1);
    const int stride = rcx_state_stride(a.max_len, k);
    for (;;) {
        long long item;
        if (!fetch_item(a, a.pass == 0 ? RCX_C_WORKX0 : RCX_C_WORKX2, lane, item)) break;
        const long long r = a.worklist[item];
        int* st = a.state + item * stride;
        const uint32_t o = a.off[r];
        const int len = (int)(a.off[r + 1] - o);
        const int kcnt = len - k + 1;
        int nN, nA, nT;
        build_planes_ool(M.P, a.seq + o, len, a.max_len, lane, nN, nA, nT);
        #pragma unroll 1
        for (int j = lane; j < kcnt; j += 32) M.counts[j] = a.counts[o + j] & 0x7fffffff;  // bit 31: validity flag of the thread-per-read scan
        __syncwarp();
        int strong, trust;
        if (a.pass == 0) {
            initial_thresholds(a, r, a.flags[r], a.error_rate, strong, trust);
        } else {
            // ---------------- re-threshold (:1247-1289)
            strong = st[ST_STRONG]; trust = st[ST_TRUST];
            const int loose = (k / 2 > 7) ? k / 2 : 7;
            const int n2 = rcx_pow2_at_least(kcnt);
            #pragma unroll 1
            for (int j = lane; j < n2; j += 32)
                M.sbuf[j] = j < kcnt ? (rcx_polya(M.P, j, k, kmask32, loose) ? -1 : M.counts[j]) : 0x7fffffff;
            __syncwarp();
            warp_sort_ool(M.sbuf, n2, lane);
            int ibreak = 0;
            bool any_drop = false;
            #pragma unroll 1
            for (int i = 1 + lane; i < kcnt; i += 32) {
                const int v = M.sbuf[i], pv = M.sbuf[i - 1];
                if (v <= strong && ((v > 2 * pv && v > 10) || (pv == 0 && v >= 5)) && v < strong) ibreak = i;
            }
            ibreak = __reduce_max_sync(RCX_FULL, ibreak);
            #pragma unroll 1
            for (int i = 1 + lane; i < kcnt; i += 32) {
                const int v = M.sbuf[i], pv = M.sbuf[i - 1];
                if (i >= ibreak && v <= strong && ((v > 2 * pv && v > 10) || (pv == 0 && v >= 5))) any_drop = true;
            }
            any_drop = __any_sync(RCX_FULL, any_drop);
            if (!any_drop) {  // no lower plateau: keep the first pass' result and go to the post-filters (:1288-1289)
                if (lane == 0) st[ST_STATUS] = RCX_ST_POSTFILTER;
                __syncwarp();
                continue;
            }
            const int nv = M.sbuf[ibreak];
            trust = rcx_bound_i_ool(nv, a.error_rate);  // no clamp to 2 here (:1282)
            strong = nv;
            __syncwarp();
        }
        #pragma unroll 1
        for (int c = lane; c < mw; c += 32) { M.pstrict[c] = 0; M.strong[c] = 0; M.trusted[c] = 0; }
        __syncwarp();
        #pragma unroll 1
        for (int c = 0; (c << 5) < kcnt; ++c) {
            const int j = (c << 5) + lane;
            const bool pa = j < kcnt && rcx_polya(M.P, j, k, kmask32, 2);
            const unsigned int bp = __ballot_sync(RCX_FULL, pa);
            const unsigned int bt = __ballot_sync(RCX_FULL, j < kcnt && M.counts[j] >= strong && !pa);  // :875
            if (lane == 0) { M.pstrict[c] = bp; M.trusted[c] = bt; }
        }
        __syncwarp();
        int status = RCX_ST_SEARCH, nseg = 0;
        if (next_bit(M.trusted, 0, kcnt, false, lane) >= kcnt) {
            status = 0;  // longestTrustedLength == kcnt (:1110)
        } else {
            // runs of trusted k-mers -> longest run, islands (runs >= 2), boundary adjustment, trusted positions
            int longest = 0, tstart = 0, tend = -1;
            {
                int pf = 0, pt = -1;
                bool have = false;
                int p = 0;
                #pragma unroll 1
                while (p < kcnt) {
                    const int s = next_bit(M.trusted, p, kcnt, true, lane);
                    if (s >= kcnt) break;
                    const int e = next_bit(M.trusted, s, kcnt, false, lane);
                    const int rl = e - s;
                    if (rl > longest) { longest = rl; tstart = s; tend = e - 1; }
                    if (rl >= 2) {
                        int cf = s, ct = e - 1;
                        if (have) {
                            if (cf <= pt + k) {  // :936-964
                                const int len1 = pt - pf, len2 = ct - cf, overlap = pt + k - cf;
                                const int j = pt + 1 + lane;
                                const bool low = j < cf && M.counts[j] <= 2 && M.counts[j] < trust;
                                if (__any_sync(RCX_FULL, low) && overlap <= 3) {
                                    if (len1 < len2) pt -= overlap + 1; else cf += overlap + 1;
                                }
                            }
                            if (pf <= pt) set_range(M.strong, pf, pt + k - 1, mw, lane);
                            __syncwarp();
                        }
                        pf = cf; pt = ct; have = true;
                    }
                    p = e;
                }
                if (have && pf <= pt) set_range(M.strong, pf, pt + k - 1, mw, lane);
                __syncwarp();
            }
            // islands that define the segments (:977-1007)
            #pragma unroll 1
            for (int c = lane; c < mw; c += 32) M.island[c] = M.strong[c];
            __syncwarp();
            if (next_bit(M.strong, 0, len, true, lane) >= len) {
                set_range(M.island, tstart, tend + k - 1, mw, lane);
                __syncwarp();
            }
            // segments (:1009-1046, :1125-1160): prefix, gaps between islands, suffix
            int ifrom = next_bit(M.island, 0, len, true, lane);
            int ito = ifrom < len ? next_bit(M.island, ifrom, len, false, lane) - 1 : len - 1;
            int prev_to = -1, prev_len = 0;
            bool first = true;
            for (;;) {
                int sf, sto, la, ra;
                bool is_seg;
                if (ifrom < len) {
                    sf = prev_to + 1; sto = ifrom - 1;
                    la = first ? 0 : prev_len; ra = ito - ifrom + 1;
                    is_seg = first ? ifrom > 0 : true;
                } else {
                    sf = prev_to + 1; sto = len; la = prev_len; ra = 0;
                    is_seg = prev_to < len - 1;
                }
                if (is_seg) {
                    if (nseg < L.nseg) {
                        const bool right = la >= ra;  // :1136
                        const int extend = right ? (sto == len ? 0 : k - 1) : (sf == 0 ? 0 : k - 1);
                        const RcxKmer kc = seed_kmer(M.P, right ? sf - k : sto + 1, k, kmask32);
                        int max_fix = (sto - sf + 1) * K / k * 2 + 1;  // :1128-1131
                        if (max_fix < K) max_fix = K;
                        if (lane == 0) {
                            int* sg = st + ST_HDR + ST_SEGW * nseg;
                            sg[0] = (right ? sf : sto) | ((right ? sto + extend : sf - extend) << 16);
                            sg[1] = (right ? 1 : 0) | (max_fix << 1);
                            sg[2] = (int)(unsigned int)kc.code;
                            sg[3] = (int)(unsigned int)(kc.code >> 32);
                            sg[4] = kc.inv;
                            sg[5] = -1;  // top2FixBottleNeck[1], filled by the search
                        }
                    }
                    ++nseg;
                }
                if (ifrom >= len) break;
                prev_to = ito; prev_len = ito - ifrom + 1; first = false;
                ifrom = next_bit(M.island, ito + 1, len, true, lane);
                ito = ifrom < len ? next_bit(M.island, ifrom, len, false, lane) - 1 : len - 1;
            }
            // per-base info for the search and a clean fix[] (:1113-1114)
            #pragma unroll 1
            for (int i = lane; i < len; i += 32) {
                const bool free_pos = !getbit(M.strong, i);
                const bool pr = i - k + 1 >= 0 && getbit(M.pstrict, i - k + 1), pl = getbit(M.pstrict, i);  // :343 / :577
                a.posbuf[o + i] = (unsigned char)(rcx_base(M.P.pk, i) | (getbit(M.P.inv, i) << 2) | ((free_pos && !pr) ? 8 : 0) |
                                                  ((free_pos && !pl) ? 16 : 0));
                if (!a.compact_fix) a.fixbuf[o + i] = -1;
            }
            if (a.compact_fix && lane == 0) a.fixlist[o] = 0;  // edit-list hand-over: an empty list
        }
        if (lane == 0) {
            st[ST_STATUS] = status; st[ST_TOTALFIX] = rcx_force_bound(strong, a.error_rate); st[ST_BADSEG] = 0;
            st[ST_STRONG] = strong; st[ST_TRUST] = trust; st[ST_NSEG] = nseg < L.nseg ? nseg : L.nseg;
            st[ST_OFF] = (int)o; st[ST_LEN] = len;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(COR_WARPS * 32) rcx_finalize_kernel(RcxCorrectArgs a) {
    extern __shared__ unsigned int smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.t.k, K = a.max_fix_per_k;
    const CorLayout L = cor_layout(a.max_len, k);
    const int mw = L.mw;
    WarpMem M;
    setup_warpmem(M, smem, warp, L);
    const unsigned int kmask32 = k >= 32 ? 0xffffffffu : ((1u << k) - 1);
    unsigned long long n_rprobes = 0;
    const long long nwork = (long long)a.counters[RCX_C_NWORK];
    for (;;) {
        long long w = 0;
        if (lane == 0) w = (long long)atomicAdd(&a.counters[RCX_C_WORK2], 1ull);
        w = __shfl_sync(RCX_FULL, w, 0);
        if (w >= nwork) break;
        const long long r = a.worklist[w];
        const uint32_t o = a.off[r];
        const int len = (int)(a.off[r + 1] - o);
        const int kcnt = len - k + 1;
        const int* st = a.state + w * rcx_state_stride(a.max_len, k);
        int result = st[ST_STATUS];
        const int total_fix = st[ST_TOTALFIX], bad_segments = st[ST_BADSEG];
        bool changed = false;
        int nN = 0, nA = 0, nT = 0;
        build_planes_ool(M.P, a.seq + o, len, a.max_len, lane, nN, nA, nT);
        #pragma unroll 1
        for (int j = lane; j < kcnt; j += 32) M.counts[j] = a.counts[o + j] & 0x7fffffff;  // bit 31: validity flag of the thread-per-read scan
        if (result == -2) {
            #pragma unroll 1
            for (int i = lane; i < len; i += 32) M.fix[i] = a.fixbuf[o + i];
        }
        __syncwarp();
        {
            if (result == -2) {
                // ---------------- A.7 post-filters (:1296-1466)
                const bool q0 = rcx_has_qual(a, o);
                auto qhi = [&](int i) { return rcx_qual_high(a, o, i); };  // qual[i] > badQualityThreshold
                int* P = M.sbuf;  // ascending positions with a fix on a non-N base
                int np = 0;
                #pragma unroll 1
                for (int c = 0; (c << 5) < len; ++c) {
                    const int i = (c << 5) + lane;
                    const bool f = i < len && M.fix[i] != -1 && !getbit(M.P.mN, i);
                    const unsigned int b = __ballot_sync(RCX_FULL, f);
                    if (f) P[np + __popc(b & ((1u << lane) - 1))] = i;
                    np += __popc(b);
                }
                __syncwarp();
                #pragma unroll 1
                for (int i = 1; i < np; ++i) {  // pair filter (:1314-1398)
                    const int pa = P[i - 1], pb = P[i];
                    if (q0 && !qhi(pb) && !qhi(pa)) continue;
                    if (pb - pa + 1 > k) continue;
                    const int lim = (i < np - 1) ? P[i + 1] - k + 1 : 0x7fffffff;  // j + k - 1 >= P[i+1]  <=>  j >= lim
                    int j0 = pa - k + 1;
                    if (j0 < 0) j0 = 0;
                    // loop 1: [j0, e1), skipping j <= P[i-2]
                    const int skip = (i >= 2) ? P[i - 2] : -1;  // j <= P[i-2] is skipped BEFORE the break tests
                    int e1 = pb - k + 1;
                    if (e1 > lim) e1 = lim;
                    if (e1 < skip + 1) e1 = skip + 1;
                    if (e1 < j0) e1 = j0;
                    if (e1 > kcnt) e1 = kcnt;
                    int e2 = pa + 1;
                    if (e2 > lim) e2 = lim;
                    if (e2 > kcnt) e2 = kcnt;
                    if (e2 < e1) e2 = e1;
                    int e3 = pb + 1;
                    if (e3 > lim) e3 = lim;
                    if (e3 > kcnt) e3 = kcnt;
                    if (e3 < e2) e3 = e2;
                    int ms = RCX_INF, md = RCX_INF;
                    #pragma unroll 1
                    for (int j = j0 + lane; j < e3; j += 32) {
                        const int c = M.counts[j];
                        if (j < e1) { if (j > skip && c < ms) ms = c; }
                        else if (j < e2) { if (c < md) md = c; }
                        else { if (c < ms) ms = c; }
                    }
                    ms = __reduce_min_sync(RCX_FULL, ms);
                    md = __reduce_min_sync(RCX_FULL, md);
                    if (ms != RCX_INF && md != RCX_INF && ms > 1 && md > 1 && ms > md / 2 && ms < 2 * md) {
                        if (lane == 0) { M.fix[pb] = -1; M.fix[pa] = -1; }
                        int j = i - 2;
                        #pragma unroll 1
                        while (j >= 0 && P[j + 1] - P[j] + 1 <= k) { if (lane == 0) M.fix[P[j]] = -1; --j; }
                        #pragma unroll 1
                        while (i + 1 < np && P[i + 1] - P[i] + 1 <= k) { if (lane == 0) M.fix[P[i + 1]] = -1; ++i; }
                    }
                }
                __syncwarp();
                if (total_fix > 3 && len > 10) {  // edge filter (:1407-1430)
                    {
                        const int i = lane;
                        const bool nn = i < 10 && !getbit(M.P.mN, i);
                        const bool f = nn && M.fix[i] != -1 && qhi(i);
                        if (__popc(__ballot_sync(RCX_FULL, f)) >= 2 && nn) M.fix[i] = -1;
                    }
                    __syncwarp();
                    {
                        const int i = len - 10 + lane;
                        const bool nn = lane < 10 && !getbit(M.P.mN, i);
                        const bool f = nn && M.fix[i] != -1 && qhi(i);
                        if (__popc(__ballot_sync(RCX_FULL, f)) >= 3 && nn) M.fix[i] = -1;
                    }
                    __syncwarp();
                }
                if (total_fix >= K) {  // density filter (:1432-1455), weights doubled to stay in integers
                    #pragma unroll 1
                    for (int c = lane; c < 2 * mw; c += 32) M.fixm[c] = 0;
                    __syncwarp();
                    #pragma unroll 1
                    for (int c = 0; (c << 5) < len; ++c) {
                        const int i = (c << 5) + lane;
                        const bool f = i < len && M.fix[i] != -1 && !getbit(M.P.mN, i);
                        const bool good = f && qhi(i);
                        const unsigned int bg = __ballot_sync(RCX_FULL, good), bb = __ballot_sync(RCX_FULL, f && !good);
                        if (lane == 0) { M.fixm[c] = bg; M.fixm[mw + c] = bb; }
                    }
                    __syncwarp();
                    bool dense = false;
                    #pragma unroll 1
                    for (int j = lane; j < kcnt; j += 32) {
                        const int w2 = 2 * __popc(rcx_bits(M.fixm, j) & kmask32) + __popc(rcx_bits(M.fixm + mw, j) & kmask32);
                        if (w2 > 2 * K) dense = true;
                    }
                    if (__any_sync(RCX_FULL, dense)) result = -1;
                }
                if (result == -2) {  // apply (:1468-1479)
                    int ret = 0;
                    #pragma unroll 1
                    for (int c = 0; (c << 5) < len; ++c) {
                        const int i = (c << 5) + lane;
                        const bool f = i < len && M.fix[i] != -1;
                        if (f) a.seq[o + i] = "ACGT"[M.fix[i]];
                        if (a.emit_edits && i < len) a.fixbuf[o + i] = M.fix[i];  // what survived the filters, for the gather kernel
                        ret += __popc(__ballot_sync(RCX_FULL, f));
                    }
                    if (ret == 0 && bad_segments > 0) result = -1;
                    else { result = ret; changed = ret > 0; }
                }
            }
        }
        // ---------------- GetKmerInformation on the final sequence (:1567-1602)
        __syncwarp();
        if (changed) {
            // fold the fixes into the bit planes, then re-probe only the windows that cover a fix
            #pragma unroll 1
            for (int c = lane; c < mw; c += 32) M.fixm[c] = 0;
            __syncwarp();
            #pragma unroll 1
            for (int c = 0; (c << 5) < len; ++c) {
                const int i = (c << 5) + lane;
                const bool f = i < len && M.fix[i] != -1;
                const unsigned int bf = __ballot_sync(RCX_FULL, f);
                const unsigned int sh = (unsigned int)(f ? M.fix[i] : 0) << (30 - 2 * (lane & 15));
                const unsigned int keep = f ? ~(3u << (30 - 2 * (lane & 15))) : 0xffffffffu;
                const unsigned int hi = __reduce_or_sync(RCX_FULL, lane < 16 ? sh : 0u), lo = __reduce_or_sync(RCX_FULL, lane >= 16 ? sh : 0u);
                const unsigned int khi = __reduce_and_sync(RCX_FULL, lane < 16 ? keep : 0xffffffffu);
                const unsigned int klo = __reduce_and_sync(RCX_FULL, lane >= 16 ? keep : 0xffffffffu);
                if (lane == 0) {
                    M.fixm[c] = bf;
                    M.P.inv[c] &= ~bf;
                    M.P.pk[2 * c] = (M.P.pk[2 * c] & khi) | hi;
                    M.P.pk[2 * c + 1] = (M.P.pk[2 * c + 1] & klo) | lo;
                }
            }
            __syncwarp();
            #pragma unroll 1
            for (int j = lane; j < kcnt; j += 32) {
                if ((rcx_bits(M.fixm, j) & kmask32) == 0) continue;
                int c = 0;
                if ((rcx_bits(M.P.inv, j) & kmask32) == 0) {
                    c = rcx_probe_kmer_ool(a.t, rcx_window(M.P.pk, j, k));
                    ++n_rprobes;
                }
                M.counts[j] = c;
            }
            __syncwarp();
        }
        {
            int l, m, h;
            kmer_info(M, kcnt, kmask32, lane, l, m, h);
            if (lane == 0) { a.correction[r] = result; a.l[r] = l; a.m[r] = m; a.h[r] = h; }
            __syncwarp();
        }
    }
    n_rprobes = warp_sum_u64(n_rprobes);
    if (lane == 0 && n_rprobes) atomicAdd(&a.counters[RCX_C_RESCAN_PROBES], n_rprobes);
    if (lane == 0 && n_rprobes) atomicAdd(&a.counters[RCX_C_RESCAN_WINDOWS], n_rprobes);  // (one window per request here)
}

static cudaError_t launch_one(void (*kern)(RcxCorrectArgs), const RcxCorrectArgs& a, int sm_count, long long items, size_t smem,
                              cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, COR_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long want = (items + COR_WARPS - 1) / COR_WARPS;
    long long grid = (long long)sm_count * per_sm;  // persistent grid: a multiple of the SM count
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    kern<<<(unsigned int)grid, COR_WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
}

size_t rcx_state_words(int max_len, int k) { return (size_t)rcx_state_stride(max_len, k); }

cudaError_t rcx_launch_correct(const RcxCorrectArgs& a0, int sm_count, int dfs_grid, cudaStream_t s, cudaEvent_t* kev, bool prep0_done) {
    if (a0.n <= 0) return cudaSuccess;
    RcxCorrectArgs a = a0;
    const size_t smem = rcx_correct_smem_per_warp(a.max_len, a.t.k) * COR_WARPS;
    int ke = 0;
#define MARK() do { if (kev) { cudaError_t _e = cudaEventRecord(kev[ke++], s); if (_e != cudaSuccess) return _e; } } while (0)
    a.pass = 0;
    MARK();  // (the triage is fused into the scan kernel: kev[0..1] brackets nothing)
    cudaError_t e = cudaSuccess;
    // the work-list lengths live on the device: the persistent kernels read them there (no host round trip)
    for (int pass = 0; pass < 2; ++pass) {
        a.pass = pass;
        MARK();
        if (!(pass == 0 && prep0_done)) {  // the thread-per-read scan kernel already prepared the first pass
            e = launch_one(rcx_prep_kernel, a, sm_count, pass == 0 ? a.n : a.n / 8 + 1, smem, s);
            if (e != cudaSuccess) return e;
        }
        MARK();
        e = rcx_launch_dfs(a, dfs_grid, s);
        if (e != cudaSuccess) return e;
    }
    MARK();
    if (prep0_done && rcx_finalize_thread_supported(a.max_len, a.t.k)) e = rcx_launch_finalize_thread(a, sm_count, s);
    else e = launch_one(rcx_finalize_kernel, a, sm_count, a.n, smem, s);
    if (e != cudaSuccess) return e;
    MARK();
#undef MARK
    return cudaSuccess;
}
