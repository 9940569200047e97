// rcx_scan.cu — K1, the lookup kernel: one warp owns one read.
//
// Restates the scan shared by ErrorCorrection (ErrorCorrection.cpp:716-723)
// and GetStrongTrustedThreshold (:1494-1501): counts[j] = GetCount of the
// k-window starting at base j (0 for a window holding a non-ACGT base), plus
// the whole of GetStrongTrustedThreshold (:1482-1565): reject tests, the
// loose-polyA-masked ascending sort and the "drop" search.  The reference
// runs this scan twice per paired read (pair threshold, then correction); on
// the GPU it runs once and the correction kernel consumes counts[] from HBM.
//
// Data flow per warp: coalesced byte loads of the read -> bit planes in shared
// memory (rcx_build_planes) -> every lane funnel-shifts its own k-windows out
// of the packed plane, canonicalises and issues SCAN_ILP independent 64-byte
// bucket loads before consuming any -> counts to shared + HBM -> warp bitonic sort.
#include "rcx_internal.h"

#define SCAN_WARPS 8
#define SCAN_ILP 2

size_t rcx_scan_smem_per_warp(int max_len, int k) {
    int kcnt = max_len - k + 1;
    if (kcnt < 1) kcnt = 1;
    return sizeof(unsigned int) * (size_t)(rcx_pk_words(max_len) + 4 * rcx_mask_words(max_len) + rcx_pow2_at_least(kcnt));
}

// sort the kcnt masked counts of one read in registers and pull out what GetStrongTrustedThreshold needs:
// di = largest index i >= 1 with s[i] > 2*s[i-1] && s[i] > 10 (0 if none), thr = s[di] or the median of the positive
// entries, below = s[di-1]
template <int R>
__device__ __forceinline__ void scan_sorted_summary(const int* sbuf, int kcnt, int lane, int& di, int& nonpos, int& thr, int& below) {
    int v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = (r << 5) | lane;
        v[r] = j < kcnt ? sbuf[j] : 0x7fffffff;
    }
    rcx_regsort<R>(v, lane);
    di = 0;
    nonpos = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = (r << 5) | lane;
        int pv = __shfl_up_sync(RCX_FULL, v[r], 1);
        if (r > 0) {
            const int pl = __shfl_sync(RCX_FULL, v[r - 1], 31);
            if (lane == 0) pv = pl;
        }
        if (i >= 1 && i < kcnt && v[r] > 2 * pv && v[r] > 10) di = i;
        nonpos += (i < kcnt && v[r] <= 0) ? 1 : 0;
    }
    di = __reduce_max_sync(RCX_FULL, di);
    nonpos = __reduce_add_sync(RCX_FULL, nonpos);
    below = 0;
    if (di >= 1) {
        thr = rcx_regsort_get<R>(v, di);
        below = rcx_regsort_get<R>(v, di - 1);
    } else {
        thr = rcx_regsort_get<R>(v, (nonpos + kcnt - 1) / 2);
    }
}

__global__ void __launch_bounds__(SCAN_WARPS * 32, 4) rcx_scan_kernel(RcxScanArgs a) {
    extern __shared__ unsigned int smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.t.k;
    const int pkw = rcx_pk_words(a.max_len), mw = rcx_mask_words(a.max_len);
    int maxk = a.max_len - k + 1;
    if (maxk < 1) maxk = 1;
    const int per_warp = pkw + 4 * mw + rcx_pow2_at_least(maxk);
    unsigned int* base = smem + (size_t)warp * per_warp;
    RcxPlanes P;
    P.pk = base;
    P.inv = base + pkw;
    P.mA = P.inv + mw;
    P.mT = P.mA + mw;
    P.mN = P.mT + mw;
    int* sbuf = reinterpret_cast<int*>(P.mN + mw);
    const unsigned int kmask32 = k >= 32 ? 0xffffffffu : ((1u << k) - 1);
    const int loose = (k / 2 > 7) ? k / 2 : 7;
    unsigned long long probes = 0;

    for (long long r = (long long)blockIdx.x * SCAN_WARPS + warp; r < a.n; r += (long long)gridDim.x * SCAN_WARPS) {
        const uint32_t o = a.off[r];
        const int len = (int)(a.off[r + 1] - o);
        const int kcnt = len - k + 1;
        if (kcnt <= 0) {  // strlen(seq) < kmerLength, :1491
            if (lane == 0) { a.thr[r] = -1; a.flags[r] = RCX_F_REJECT; }
            continue;
        }
        int nN, nA, nT;
        rcx_build_planes(P, a.seq + o, len, a.max_len, lane, nN, nA, nT);

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
                live[u] = i < nctx && (rcx_bits(P.inv, i < nctx ? c : 0) & (kmask32 >> 1)) == 0;  // no non-ACGT base in the context
                if (live[u]) {
                    ctx[u] = rcx_window(P.pk, c, k - 1);
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
                    if (!((P.inv[(c - 1) >> 5] >> ((c - 1) & 31)) & 1u)) cl = rcx_count1(a.t, ctx[u], true, rcx_base(P.pk, c - 1), L[u], found);  // window c-1
                    if (c < kcnt && !((P.inv[(c + k - 1) >> 5] >> ((c + k - 1) & 31)) & 1u))
                        cr = rcx_count1(a.t, ctx[u], false, rcx_base(P.pk, c + k - 1), L[u], found);  // window c
                }
                if (i < nctx) {
                    a.counts[o + c - 1] = cl;
                    sbuf[c - 1] = rcx_polya(P, c - 1, k, kmask32, loose) ? -1 : cl;  // :1530-1539
                    if (c < kcnt) {
                        a.counts[o + c] = cr;
                        sbuf[c] = rcx_polya(P, c, k, kmask32, loose) ? -1 : cr;
                    }
                }
            }
        }
        // ---- reject tests (:1507-1527)
        const bool rej = nN > 5 || nA > len - k || nT > len - k;
        if (rej) {
            if (lane == 0) { a.thr[r] = -1; a.flags[r] = RCX_F_REJECT; }
            __syncwarp();
            continue;
        }
        // ---- masked sort + drop search (:1540-1563)
        int thr, di = 0, nonpos = 0, below = 0;
        if (kcnt <= 128) {
            scan_sorted_summary<4>(sbuf, kcnt, lane, di, nonpos, thr, below);
        } else {  // longer reads: shared-memory network (keeps the common path at 4 sort registers per lane)
            const int n2 = rcx_pow2_at_least(kcnt);
            for (int j = kcnt + lane; j < n2; j += 32) sbuf[j] = 0x7fffffff;
            __syncwarp();
            rcx_warp_sort(sbuf, n2, lane);
            for (int i = lane; i < kcnt; i += 32) {
                const int v = sbuf[i];
                if (i >= 1) {
                    const int pv = sbuf[i - 1];
                    if (v > 2 * pv && v > 10) di = i;  // indices ascend per lane, so the last hit is the largest
                }
                nonpos += v <= 0;
            }
            di = __reduce_max_sync(RCX_FULL, di);
            nonpos = __reduce_add_sync(RCX_FULL, nonpos);
            if (di >= 1) { thr = sbuf[di]; below = sbuf[di - 1]; }
            else thr = sbuf[(nonpos + kcnt - 1) / 2];  // first index with a positive value == number of entries <= 0
        }
        if (lane == 0) {
            a.thr[r] = thr;
            a.flags[r] = di >= 1 ? (uint8_t)(RCX_F_DROP | (below == 2 ? RCX_F_PREV2 : 0)) : (uint8_t)0;
        }
        __syncwarp();
    }
    probes = __reduce_add_sync(RCX_FULL, (unsigned int)probes);
    if (lane == 0 && probes) atomicAdd(&a.counters[RCX_C_SCAN_PROBES], probes);
}

cudaError_t rcx_launch_scan(const RcxScanArgs& a, int sm_count, cudaStream_t s) {
    if (a.n <= 0) return cudaSuccess;
    const size_t smem = rcx_scan_smem_per_warp(a.max_len, a.t.k) * SCAN_WARPS;
    cudaError_t e = cudaFuncSetAttribute(rcx_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rcx_scan_kernel, SCAN_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long want = (a.n + SCAN_WARPS - 1) / SCAN_WARPS;
    long long grid = (long long)sm_count * per_sm;  // persistent grid: a multiple of the SM count
    if (grid > want) grid = want;
    rcx_scan_kernel<<<(unsigned int)grid, SCAN_WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
}
