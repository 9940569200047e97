// rcx_scan.cu — K1, the lookup kernel: one warp owns one read (or the two mates of a pair).
//
// Restates, fused in one pass over the read while it sits in shared memory:
//   * the scan shared by ErrorCorrection (ErrorCorrection.cpp:716-723) and GetStrongTrustedThreshold (:1494-1501):
//     counts[j] = GetCount of the k-window starting at base j (0 for a window holding a non-ACGT base).  The reference
//     runs it twice per paired read; here it runs once and later kernels consume counts[] from HBM;
//   * GetStrongTrustedThreshold (:1482-1565): reject tests, loose-polyA mask, "drop" search in the sorted counts;
//   * the triage of ErrorCorrection: initial thresholds (:772-842) and the "every k-mer trusted" exit (:1110); reads
//     that leave here (clean or rejected) also get their l/m/h tags (GetKmerInformation, :1567-1602); the others go
//     to the work list of the correction kernels.
//
// Data flow per warp: coalesced byte loads of the read -> bit planes in shared memory (rcx_build_planes) -> every lane
// funnel-shifts its (k-1)-mer contexts out of the packed plane and issues SCAN_ILP independent one-sector lookups, each
// answering TWO windows -> counts to shared + HBM -> thresholds.
//
// No sort: the reference sorts the masked counts only to find the largest i with s[i] > 2*s[i-1] && s[i] > 10 and a
// median.  Two sorted neighbours with a ratio > 2 cannot share a power-of-two octave, so a drop can only sit between the
// smallest value of an occupied octave and the largest value of the next occupied octave below it: the warp walks the
// occupied octaves from the top (a 32-bit occupancy mask, two warp reductions per octave) and stops at the first
// drop; the median (only needed when there is no drop) is a bisection on the value.  Same results as the sort for
// counts < 2^30 (beyond that the reference's own 2*x overflows).
#include "rcx_internal.h"

#define SCAN_WARPS 8
#define SCAN_ILP 2

struct ScanSlot {
    RcxPlanes P;
    int* cnt;
};

static __host__ __device__ inline int scan_slot_words(int max_len, int k) {
    int kc = max_len - k + 1;
    if (kc < 1) kc = 1;
    return rcx_pk_words(max_len) + 4 * rcx_mask_words(max_len) + ((kc + 3) & ~3);
}
size_t rcx_scan_smem_per_warp(int max_len, int k) { return sizeof(unsigned int) * 2 * (size_t)scan_slot_words(max_len, k); }

// octave id of a masked count: 0 for <= 0, 1 + floor(log2 v) otherwise
__device__ __forceinline__ int scan_octave(int v) { return v <= 0 ? 0 : 32 - __clz(v); }

// The values of one read: 32 lanes x R registers (element j = r * 32 + lane) when the read has <= 32 R windows, in
// shared memory when R == 0.  Padding entries are INT_MAX: never <= a real value, and never marked as an occupied octave.
template <int R>
struct ScanVals {
    int v[R > 0 ? R : 1];
    const int* s;
    int n, lane;
    template <class F>
    __device__ __forceinline__ void each(F f) const {
        if (R > 0) {
#pragma unroll
            for (int r = 0; r < (R > 0 ? R : 1); ++r) f(v[r]);
        } else {
            for (int j = lane; j < n; j += 32) f(s[j]);
        }
    }
};

// smallest x with #{v <= x} > rank  (= element `rank` of the ascending order), lo/hi = bounds of the values
template <int R>
__device__ __forceinline__ int scan_select(const ScanVals<R>& V, int rank, int lo, int hi) {
    while (lo < hi) {
        const int mid = lo + (int)(((unsigned int)hi - (unsigned int)lo) >> 1);
        int c = 0;
        V.each([&](int x) { c += x <= mid; });
        c = __reduce_add_sync(RCX_FULL, c);
        if (c > rank) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// GetStrongTrustedThreshold's sorted-array logic (:1540-1563) on the masked counts: found = a drop exists, thr = s[i] at
// the largest drop or the median of the positive entries, below = s[i-1]
template <int R>
__device__ __forceinline__ void scan_threshold(const ScanVals<R>& V, int kcnt, bool& found, int& thr, int& below) {
    unsigned int present = 0;
    int nonpos = 0, gmin = 0x7fffffff, gmax = -1;
    V.each([&](int x) {
        if (x != 0x7fffffff) {
            present |= 1u << scan_octave(x);
            nonpos += x <= 0;
            gmin = x < gmin ? x : gmin;
            gmax = x > gmax ? x : gmax;
        }
    });
    present = __reduce_or_sync(RCX_FULL, present);
    found = false;
    thr = 0;
    below = 0;
    unsigned int left = present;
    while (left) {
        const int b = 31 - __clz(left);
        const unsigned int lower = left & ~(1u << b);
        if (b < 4 || lower == 0) break;  // values <= 10 cannot be a drop; the smallest element has no predecessor
        int mn = 0x7fffffff;
        V.each([&](int x) { if (x != 0x7fffffff && scan_octave(x) == b && x < mn) mn = x; });
        mn = __reduce_min_sync(RCX_FULL, mn);
        if (mn <= 10) break;  // everything further down is smaller still
        const int b2 = 31 - __clz(lower);
        int mx = (int)0x80000000;
        V.each([&](int x) { if (scan_octave(x) == b2 && x > mx) mx = x; });
        mx = __reduce_max_sync(RCX_FULL, mx);
        if (mn > 2 * mx) { found = true; thr = mn; below = mx; break; }
        left = lower;
    }
    if (!found) {
        nonpos = __reduce_add_sync(RCX_FULL, nonpos);
        gmin = __reduce_min_sync(RCX_FULL, gmin);
        gmax = __reduce_max_sync(RCX_FULL, gmax);
        thr = scan_select<R>(V, (nonpos + kcnt - 1) / 2, gmin, gmax);  // first index with a positive value == number of entries <= 0
    }
}

// GetKmerInformation (:1567-1602) from counts[] and the validity plane: l = min, m = element [n/2], h = max of the
// counts of the VALID k-mers with 0 -> 1; all 0 when there is none.  (Overwrites S.cnt when R == 0.)
template <int R>
__device__ __forceinline__ void scan_kmer_info(const ScanSlot& S, int kcnt, unsigned int kmask32, int lane, int& l, int& m, int& h) {
    ScanVals<R> V;
    V.s = S.cnt; V.n = kcnt; V.lane = lane;
    int nvalid = 0, mn = 0x7fffffff, mx = 0;
    auto one = [&](int j) {
        int x = 0x7fffffff;
        if (j < kcnt && (rcx_bits(S.P.inv, j) & kmask32) == 0) {
            x = S.cnt[j];
            if (x == 0) x = 1;
            ++nvalid;
            mn = x < mn ? x : mn;
            mx = x > mx ? x : mx;
        }
        return x;
    };
    if (R > 0) {
#pragma unroll
        for (int r = 0; r < (R > 0 ? R : 1); ++r) V.v[r] = one((r << 5) | lane);
    } else {
        for (int j = lane; j < kcnt; j += 32) { const int x = one(j); S.cnt[j] = x; }
        __syncwarp();
    }
    nvalid = __reduce_add_sync(RCX_FULL, nvalid);
    mn = __reduce_min_sync(RCX_FULL, mn);
    mx = __reduce_max_sync(RCX_FULL, mx);
    if (nvalid == 0) { l = m = h = 0; return; }
    l = mn; h = mx;
    m = scan_select<R>(V, nvalid / 2, mn, mx);
}

// thresholds of A.3 (:772-842) from GetStrongTrustedThreshold of the read (own) and of the pair (pair_t, -1 = unpaired)
__device__ __forceinline__ void scan_initial_thresholds(int own, bool drop, bool prev2, int pair_t, double er, int& strong, int& trust) {
    strong = own;
    bool flag = false;
    trust = rcx_bound_i(strong, er);
    if (drop && strong >= 20 && prev2 && trust < 3) { flag = true; trust = 3; }
    if (pair_t >= 1 && strong > pair_t) {
        if (!flag || pair_t < 20) trust = rcx_bound_i(pair_t, er);
        strong = pair_t;
    }
    if (trust < 2) trust = 2;
}

// phase A for one read: planes, lookups, counts, GetStrongTrustedThreshold -> thr (-1 = rejected) and flags
template <int R>
__device__ __forceinline__ void scan_read(const RcxScanArgs& a, const ScanSlot& S, uint32_t o, int len, int lane, int k,
                                          unsigned int kmask32, int loose, unsigned int& probes, int& thr, uint8_t& flags) {
    const int kcnt = len - k + 1;
    thr = -1;
    flags = RCX_F_REJECT;
    if (kcnt <= 0) return;  // strlen(seq) < kmerLength, :1491
    int nN, nA, nT;
    rcx_build_planes(S.P, a.seq + o, len, a.max_len, lane, nN, nA, nT);
    // ---- lookups: windows c-1 and c share the context seq[c .. c+k-1) (c odd), so one sector answers both: the left
    // extension by seq[c-1] and the right extension by seq[c+k-1].  SCAN_ILP contexts per lane in flight.
    const int nctx = (kcnt + 1) >> 1;
    for (int i0 = 0; i0 < nctx; i0 += 32 * SCAN_ILP) {
        RcxLookup L[SCAN_ILP];
        unsigned long long ctx[SCAN_ILP];
        bool live[SCAN_ILP];
#pragma unroll
        for (int u = 0; u < SCAN_ILP; ++u) {
            const int i = i0 + u * 32 + lane;
            const int c = 2 * i + 1;
            live[u] = i < nctx && (rcx_bits(S.P.inv, i < nctx ? c : 0) & (kmask32 >> 1)) == 0;  // no non-ACGT base in the context
            if (live[u]) {
                ctx[u] = rcx_window(S.P.pk, c, k - 1);
                rcx_lookup_begin(a.t, ctx[u], L[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < SCAN_ILP; ++u) {
            const int i = i0 + u * 32 + lane;
            const int c = 2 * i + 1;
            int cl = 0, cr = 0;
            if (live[u]) {
                ++probes;
                const bool found = rcx_lookup_end(a.t, L[u]);
                // the context is clean, so a window is valid iff its one extra base is
                if (!((S.P.inv[(c - 1) >> 5] >> ((c - 1) & 31)) & 1u)) cl = rcx_count1(a.t, ctx[u], true, rcx_base(S.P.pk, c - 1), L[u], found);  // window c-1
                if (c < kcnt && !((S.P.inv[(c + k - 1) >> 5] >> ((c + k - 1) & 31)) & 1u))
                    cr = rcx_count1(a.t, ctx[u], false, rcx_base(S.P.pk, c + k - 1), L[u], found);  // window c
            }
            if (i < nctx) {
                a.counts[o + c - 1] = cl;
                S.cnt[c - 1] = cl;
                if (c < kcnt) { a.counts[o + c] = cr; S.cnt[c] = cr; }
            }
        }
    }
    __syncwarp();
    // ---- reject tests (:1507-1527)
    if (nN > 5 || nA > len - k || nT > len - k) return;
    // ---- loose-polyA mask (:1530-1539) + drop search / median (:1540-1563)
    ScanVals<R> V;
    V.s = S.cnt; V.n = kcnt; V.lane = lane;
    bool found;
    int below;
    if (R > 0) {
#pragma unroll
        for (int q = 0; q < (R > 0 ? R : 1); ++q) {
            const int j = (q << 5) | lane;
            V.v[q] = j < kcnt ? (rcx_polya(S.P, j, k, kmask32, loose) ? -1 : S.cnt[j]) : 0x7fffffff;
        }
        scan_threshold<R>(V, kcnt, found, thr, below);
    } else {
        // long reads: mask in place, then restore the raw counts from HBM (written above by this warp)
        for (int j = lane; j < kcnt; j += 32)
            if (rcx_polya(S.P, j, k, kmask32, loose)) S.cnt[j] = -1;
        __syncwarp();
        scan_threshold<R>(V, kcnt, found, thr, below);
        __syncwarp();
        for (int j = lane; j < kcnt; j += 32) S.cnt[j] = a.counts[o + j];
        __syncwarp();
    }
    flags = found ? (uint8_t)(RCX_F_DROP | (below == 2 ? RCX_F_PREV2 : 0)) : (uint8_t)0;
}

// phase B for one read: triage.  Clean / rejected reads get their result and tags here, the others go to the work list.
template <int R>
__device__ __forceinline__ void scan_triage(const RcxScanArgs& a, const ScanSlot& S, long long r, int len, int lane, int k, unsigned int kmask32,
                                            int thr, uint8_t flags, int pair_t) {
    const int kcnt = len - k + 1;
    if (kcnt <= 0) {  // :713 -> -1 ; GetKmerInformation gives 0,0,0
        if (lane == 0) { a.correction[r] = -1; a.l[r] = 0; a.m[r] = 0; a.h[r] = 0; }
        return;
    }
    int result = -2;
    if (flags & RCX_F_REJECT) {
        result = -1;
    } else {
        int strong, trust;
        scan_initial_thresholds(thr, (flags & RCX_F_DROP) != 0, (flags & RCX_F_PREV2) != 0, pair_t, a.error_rate, strong, trust);
        bool bad = false;  // any untrusted k-mer (:875)?
        for (int j = lane; j < kcnt; j += 32) bad |= !(S.cnt[j] >= strong && !rcx_polya(S.P, j, k, kmask32, 2));
        if (!__any_sync(RCX_FULL, bad)) result = 0;  // longestTrustedLength == kcnt (:1110)
    }
    if (result == -2) {
        if (lane == 0) a.worklist[atomicAdd(&a.counters[RCX_C_NWORK], 1ull)] = (uint32_t)r;
    } else {
        int l, m, h;
        scan_kmer_info<R>(S, kcnt, kmask32, lane, l, m, h);
        if (lane == 0) { a.correction[r] = result; a.l[r] = l; a.m[r] = m; a.h[r] = h; }
    }
}

template <int R>
__global__ void __launch_bounds__(SCAN_WARPS * 32, 4) rcx_scan_kernel(RcxScanArgs a) {
    extern __shared__ unsigned int smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.t.k;
    const int pkw = rcx_pk_words(a.max_len), mw = rcx_mask_words(a.max_len);
    const int slot_words = scan_slot_words(a.max_len, k);
    ScanSlot S[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        unsigned int* base = smem + ((size_t)warp * 2 + q) * slot_words;
        S[q].P.pk = base;
        S[q].P.inv = base + pkw;
        S[q].P.mA = S[q].P.inv + mw;
        S[q].P.mT = S[q].P.mA + mw;
        S[q].P.mN = S[q].P.mT + mw;
        S[q].cnt = reinterpret_cast<int*>(S[q].P.mN + mw);
    }
    const unsigned int kmask32 = k >= 32 ? 0xffffffffu : ((1u << k) - 1);
    const int loose = (k / 2 > 7) ? k / 2 : 7;
    unsigned int probes = 0;
    const int per = a.mode == 1 ? 2 : 1;  // reads per unit: the two mates of a pair share a warp (pair threshold, :96-106)
    const long long units = (a.n + per - 1) / per;
    for (long long u = (long long)blockIdx.x * SCAN_WARPS + warp; u < units; u += (long long)gridDim.x * SCAN_WARPS) {
        int thr[2] = {-1, -1}, len[2] = {0, 0};
        uint8_t fl[2] = {0, 0};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const long long r = u * per + q;
            if (q < per && r < a.n) {
                const uint32_t o = a.off[r];
                len[q] = (int)(a.off[r + 1] - o);
                scan_read<R>(a, S[q], o, len[q], lane, k, kmask32, loose, probes, thr[q], fl[q]);
                if (lane == 0) { a.thr[r] = thr[q]; a.flags[r] = fl[q]; }
            }
        }
        if (a.triage) {
            const int pair_t = per == 2 ? (thr[0] < thr[1] ? thr[0] : thr[1]) : -1;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const long long r = u * per + q;
                if (q < per && r < a.n) scan_triage<R>(a, S[q], r, len[q], lane, k, kmask32, thr[q], fl[q], pair_t);
            }
        }
        __syncwarp();
    }
    unsigned long long total = __reduce_add_sync(RCX_FULL, probes);
    if (lane == 0 && total) atomicAdd(&a.counters[RCX_C_SCAN_PROBES], total);
}

template <int R>
static cudaError_t launch_scan_r(const RcxScanArgs& a, int sm_count, cudaStream_t s) {
    const size_t smem = rcx_scan_smem_per_warp(a.max_len, a.t.k) * SCAN_WARPS;
    cudaError_t e = cudaFuncSetAttribute(rcx_scan_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rcx_scan_kernel<R>, SCAN_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int per = a.mode == 1 ? 2 : 1;
    long long want = ((a.n + per - 1) / per + SCAN_WARPS - 1) / SCAN_WARPS;
    long long grid = (long long)sm_count * per_sm;  // persistent grid: a multiple of the SM count
    if (grid > want) grid = want;
    rcx_scan_kernel<R><<<(unsigned int)grid, SCAN_WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t rcx_launch_scan(const RcxScanArgs& a, int sm_count, cudaStream_t s) {
    if (a.n <= 0) return cudaSuccess;
    const int kcnt = a.max_len - a.t.k + 1;
    if (kcnt <= 96) return launch_scan_r<3>(a, sm_count, s);   // e.g. 100 bp at k = 23: 78 windows
    if (kcnt <= 128) return launch_scan_r<4>(a, sm_count, s);  // 150 bp at k = 23 / 31
    return launch_scan_r<0>(a, sm_count, s);                   // longer reads: the values stay in shared memory
}
