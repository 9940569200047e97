// rcx_finalize_thread.cu — the end of ErrorCorrection for reads of ordinary length, ONE THREAD per work item:
// post-filters on fix[] (ErrorCorrection.cpp:1296-1466), apply (:1468-1479) and GetKmerInformation on the final
// sequence (:1567-1602).  Same organisation as rcx_scan_thread.cu: a block stages a tile of work items through
// shared memory (counts[], bases and fix[] of each item copied by a warp with coalesced loads), every thread then
// runs the scalar filter logic on its own conflict-free rows, re-probes only the k-windows that cover a changed base
// (software-pipelined one-sector lookups) and the few medians are served warp-cooperatively.
//
// The warp-per-item version of this kernel (rcx_correct.cu, still used for long reads) spent ~2 000 warp
// instructions per item on bit-plane rebuilds, ballots and reductions (profiles/r01_g).
#include "rcx_internal.h"

#define TF_THREADS 128
#ifndef TF_DEPTH
#define TF_DEPTH 4
#endif
#define TF_INVALID 0x80000000u
#define TF_CMASK 0x7fffffffu

struct TfLayout {
    int cw, bw;  // words per count row (odd), bytes per base / fix row
    size_t smem;
};
__host__ __device__ inline TfLayout tf_layout(int max_len, int k, int nt) {
    TfLayout L;
    int kc = max_len - k + 1;
    if (kc < 1) kc = 1;
    L.cw = kc | 1;
    L.bw = ((max_len + 3 + 3) & ~3) | 4;  // room for a word-aligned copy; multiple of 4 with an odd number of words (conflict-free across lanes)
    L.smem = (size_t)nt * (sizeof(int) * L.cw + (size_t)L.bw + 16) + 64;
    return L;
}

__device__ __forceinline__ bool tf_find(const RcxTable& t, unsigned long long key, RcxEntryRegs& d) {
    if ((d.q0 & RCX_KEY_MASK) == key) return true;
    if (d.q0 == ~0ull || !(d.q0 & RCX_OVERFLOW_BIT)) return false;
    const unsigned long long home = __umul64hi(rcx_hash(key), t.nlines << 2);
    for (unsigned int i = 1;; ++i) {
        rcx_issue(t, rcx_probe_sector(t, home, i), d);
        if (d.q0 == ~0ull) return false;
        if ((d.q0 & RCX_KEY_MASK) == key) return true;
        if (i >= 4 * t.nlines) return false;
    }
}

struct TfMeta {
    uint32_t r, o;
    int len, status;
};
// 4-byte global -> shared copy without a register round trip (LDGSTS)
__device__ __forceinline__ void tf_cp4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned int)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

struct TfPend {
    unsigned long long key;
    RcxEntryRegs d;
    int info;  // -1: empty; else window j | base << 10 | flip << 12
};

template <int NW, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) rcx_finalize_thread_kernel(RcxCorrectArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = a.t.k, K = a.max_fix_per_k;
    const TfLayout LY = tf_layout(a.max_len, k, NT);
    unsigned int* cnt_all = reinterpret_cast<unsigned int*>(smem_raw);
    unsigned char* seq_all = reinterpret_cast<unsigned char*>(cnt_all + (size_t)NT * LY.cw);
    TfMeta* meta = reinterpret_cast<TfMeta*>(seq_all + (size_t)NT * LY.bw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    const long long nheavy = (long long)a.counters[RCX_C_NWORKH];
    const long long nwork = (long long)a.counters[RCX_C_NWORK] + nheavy;
    const long long ntiles = (nwork + NT - 1) / NT;
    const int stride = rcx_state_stride(a.max_len, k);
    unsigned int n_rprobes = 0, n_rwindows = 0;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long w0 = tile * NT;
        const int nw = (int)((nwork - w0) < NT ? (nwork - w0) : NT);
        __syncthreads();
        // ---- stage: every thread fetches the metadata of its own item, then a warp copies one item at a time with
        // cp.async (global -> shared without a register round trip, so the copies of all items overlap)
        if (tid < nw) {
            const long long slot = rcx_work_slot(w0 + tid, nheavy, a.n);
            const long long rr = a.worklist[slot];
            const uint32_t oo = a.off[rr];
            meta[tid].r = (uint32_t)rr;
            meta[tid].o = oo;
            meta[tid].len = (int)(a.off[rr + 1] - oo);
            meta[tid].status = a.state[slot * stride + ST_STATUS];
        }
        __syncthreads();
        for (int t = warp; t < nw; t += NT / 32) {
            const uint32_t o = meta[t].o;
            const int len = meta[t].len, kc = len - k + 1;
            unsigned int* row = cnt_all + (size_t)t * LY.cw;
            unsigned char* Sw = seq_all + (size_t)t * LY.bw;   // word-aligned copies: the row starts at base (o & ~3)
            for (int j = lane; j < kc; j += 32) tf_cp4(row + j, a.counts + o + j);
            const uint32_t ob = o & ~3u;
            const int nwords = (int)(((o + len + 3) & ~3u) - ob) >> 2;
            for (int i = lane; i < nwords; i += 32) {
                tf_cp4(Sw + 4 * i, a.seq + ob + 4 * i);
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();

        const bool have = tid < nw;
        long long r = 0;
        uint32_t o = 0;
        int len = 0, kcnt = 0, result = 0;
        unsigned int* row = cnt_all + (size_t)tid * LY.cw;
        unsigned char* S = seq_all + (size_t)tid * LY.bw;
        int nvalid = 0, vmin = 0x7fffffff, vmax = 0, med = 0;
        if (have) {
            r = meta[tid].r;
            o = meta[tid].o;
            len = meta[tid].len;
            kcnt = len - k + 1;
            S += o & 3;
            const int* st = a.state + rcx_work_slot(w0 + tid, nheavy, a.n) * stride;
            result = meta[tid].status;
            const int total_fix = st[ST_TOTALFIX], bad_segments = st[ST_BADSEG];
            bool changed = false;
            RcxMask<NW> need;  // k-windows whose count must be looked up again
            need.clear();
            if (result == RCX_ST_POSTFILTER) {
                // ---------------- A.7 post-filters (:1296-1466).  fix[] arrives as the search kernel's edit list
                // (count, then pos / base pairs, in the read's fixbuf bytes): every filter works on the few edited positions.
                const bool q0 = rcx_has_qual(a, o);
                auto qhi = [&](int i) { return rcx_qual_high(a, o, i); };  // qual[i] > badQualityThreshold
                unsigned char P[256], V[256];    // ascending positions with a fix on a non-N base, their bases (0xff = fix removed)
                unsigned char NP[8], NV[8];      // fixes on literal-N positions (no filter touches them; a read with more than 5 N is rejected)
                int np = 0, nn = 0;
                {
                    const unsigned short* fl = a.fixlist + o;
                    const int cnt = fl[0];  // at most len - 1 <= 254
                    for (int e = 0; e < cnt; ++e) {  // insertion by position (segments come in order, a Left search lists its edits downwards)
                        const unsigned int xe = fl[1 + e];
                        const unsigned char pe = (unsigned char)(xe & 255u), be = (unsigned char)(xe >> 8);
                        if (S[pe] == 'N') {
                            int j = nn++;
                            while (j > 0 && NP[j - 1] > pe) { NP[j] = NP[j - 1]; NV[j] = NV[j - 1]; --j; }
                            NP[j] = pe; NV[j] = be;
                        } else {
                            int j = np++;
                            while (j > 0 && P[j - 1] > pe) { P[j] = P[j - 1]; V[j] = V[j - 1]; --j; }
                            P[j] = pe; V[j] = be;
                        }
                    }
                }
                for (int i = 1; i < np; ++i) {  // pair filter (:1314-1398)
                    const int pa = P[i - 1], pb = P[i];
                    if (q0 && !qhi(pb) && !qhi(pa)) continue;
                    if (pb - pa + 1 > k) continue;
                    const int lim = (i < np - 1) ? P[i + 1] - k + 1 : 0x7fffffff;  // j + k - 1 >= P[i+1]  <=>  j >= lim
                    int j0 = pa - k + 1;
                    if (j0 < 0) j0 = 0;
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
                    for (int j = j0; j < e3; ++j) {
                        const int c = (int)(row[j] & TF_CMASK);  // (bit 31 = validity flag; an invalid window counts 0)
                        if (j < e1) { if (j > skip && c < ms) ms = c; }
                        else if (j < e2) { if (c < md) md = c; }
                        else { if (c < ms) ms = c; }
                    }
                    if (ms != RCX_INF && md != RCX_INF && ms > 1 && md > 1 && ms > md / 2 && ms < 2 * md) {
                        V[i] = 0xff; V[i - 1] = 0xff;
                        int j = i - 2;
                        while (j >= 0 && P[j + 1] - P[j] + 1 <= k) { V[j] = 0xff; --j; }
                        while (i + 1 < np && P[i + 1] - P[i] + 1 <= k) { V[i + 1] = 0xff; ++i; }
                    }
                }
                if (total_fix > 3 && len > 10) {  // edge filter (:1407-1430)
                    int c = 0;
                    for (int i = 0; i < np; ++i) c += P[i] < 10 && V[i] != 0xff && qhi(P[i]);
                    if (c >= 2)
                        for (int i = 0; i < np; ++i) if (P[i] < 10) V[i] = 0xff;
                    c = 0;
                    for (int i = 0; i < np; ++i) c += P[i] >= len - 10 && V[i] != 0xff && qhi(P[i]);
                    if (c >= 3)
                        for (int i = 0; i < np; ++i) if (P[i] >= len - 10) V[i] = 0xff;
                }
                if (total_fix >= K) {  // density filter (:1432-1455): weight of the fixes inside any k consecutive positions
                    int sum = 0, lo = 0;  // (doubled to stay in integers; a window is bounded by its first fix)
                    bool dense = false;
                    for (int i = 0; i < np; ++i) {
                        if (V[i] == 0xff) continue;
                        sum += qhi(P[i]) ? 2 : 1;
                        while (P[i] - P[lo] >= k || V[lo] == 0xff) { if (V[lo] != 0xff) sum -= qhi(P[lo]) ? 2 : 1; ++lo; }
                        if (sum > 2 * K) dense = true;
                    }
                    if (dense) result = -1;
                }
                if (result == RCX_ST_POSTFILTER) {  // apply (:1468-1479): the global read and the staged copy
                    int ret = 0;
                    auto put = [&](int i, int base) {
                        const char ch = "ACGT"[base];
                        a.seq[o + i] = ch;
                        S[i] = (unsigned char)ch;
                        ++ret;
                        need.set(i - k + 1 > 0 ? i - k + 1 : 0, i < kcnt - 1 ? i : kcnt - 1);  // windows that cover the changed base
                    };
                    for (int i = 0; i < np; ++i) if (V[i] != 0xff) put(P[i], V[i]);
                    for (int i = 0; i < nn; ++i) put(NP[i], NV[i]);
                    if (ret == 0 && bad_segments > 0) result = -1;
                    else { result = ret; changed = ret > 0; }
                    if (a.emit_edits && changed) {  // the applied edits (a subset of the list that came in) for the gather kernel
                        unsigned short* fl = a.fixlist + o;
                        int e = 0;
                        for (int i = 0; i < np; ++i) if (V[i] != 0xff) fl[1 + e++] = (unsigned short)(P[i] | ((unsigned int)V[i] << 8));
                        for (int i = 0; i < nn; ++i) fl[1 + e++] = (unsigned short)(NP[i] | ((unsigned int)NV[i] << 8));
                        fl[0] = (unsigned short)e;
                    }
                }
            }
            // ---------------- GetKmerInformation on the final sequence (:1567-1602).  Unchanged windows keep the count and the
            // validity flag of the scan; the runs of windows that cover a changed base are rolled and looked up again
            // (TF_DEPTH one-sector lookups in flight).
            if (changed) {
                int jpos = 0;
                for (;;) {
                    const int s0 = need.next(jpos, kcnt, true);
                    if (s0 >= kcnt) break;
                    const int e0 = need.next(s0, kcnt, false);
                    jpos = e0;
                    n_rwindows += (unsigned int)(e0 - s0);
                    unsigned long long F = 0;
                    int run = 0, run_prev = 0, left = 0, i = s0;
                    auto append = [&]() {
                        const unsigned char ch = S[i];  // (the staged copy already holds the corrected bases)
                        const unsigned int idx = (unsigned int)ch - 'A';
                        const bool ok = idx < 26u && ((0x00080045u >> idx) & 1u);
                        const int code = ok ? (((ch >> 1) ^ (ch >> 2)) & 3) : 3;
                        left = (int)((F >> (2 * (k - 1))) & 3);
                        F = ((F << 2) | (unsigned long long)code) & kmask;
                        run_prev = run;
                        run = ok ? run + 1 : 0;
                        ++i;
                    };
                    for (int q = 0; q < k - 1; ++q) append();
                    int j = s0;
                    while (j < e0) {
                        TfPend p[TF_DEPTH];
#pragma unroll
                        for (int u = 0; u < TF_DEPTH; ++u) {
                            p[u].info = -1;
                            if (j < e0) {
                                // like the scan: the context seq[j+1 .. j+k) answers window j (its left extension by seq[j]) and
                                // window j+1 (its right extension by seq[j+k]) -- one sector for two windows of the run
                                append();  // completes window j
                                const bool pair = j + 1 < e0;
                                bool wantL, wantR;
                                int bl, br;
                                if (pair) {
                                    append();  // completes window j+1
                                    wantL = run_prev >= k; wantR = run >= k;
                                    bl = left; br = (int)(F & 3);
                                    if (!wantL) row[j] = TF_INVALID;
                                    if (!wantR) row[j + 1] = TF_INVALID;
                                } else {  // last window of an odd run: right extension of its prefix context
                                    wantL = false; wantR = run >= k;
                                    bl = 0; br = (int)(F & 3);
                                    if (!wantR) row[j] = TF_INVALID;
                                }
                                if (wantL || wantR) {
                                    const unsigned long long ctx = F >> 2;
                                    const unsigned long long rc = rcx_revcomp(ctx, k - 1);
                                    const bool flip = rc < ctx;
                                    p[u].key = flip ? rc : ctx;
                                    p[u].info = (pair ? j + 1 : j) | (bl << 10) | (br << 12) | (wantL ? 1 << 14 : 0) | (wantR ? 1 << 15 : 0) | (flip ? 1 << 16 : 0);
                                    rcx_issue(a.t, __umul64hi(rcx_hash(p[u].key), a.t.nlines << 2), p[u].d);
                                    ++n_rprobes;
                                }
                                j += pair ? 2 : 1;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < TF_DEPTH; ++u) {
                            const int info = p[u].info;
                            if (info >= 0) {
                                const bool found = tf_find(a.t, p[u].key, p[u].d);
                                const bool flip = (info >> 16) & 1;
                                const int jr = info & 1023;  // window answered by the right extension
                                auto field = [&](bool left_ext, int b) -> int {
                                    if (!found) return 0;
                                    unsigned int f0, f1, f2, f3;
                                    rcx_group(p[u].d, left_ext != flip, f0, f1, f2, f3);
                                    const int bb = flip ? 3 - b : b;
                                    const unsigned int f = bb == 0 ? f0 : bb == 1 ? f1 : bb == 2 ? f2 : f3;
                                    if (f == RCX_BIG) {
                                        const unsigned long long ctx = flip ? rcx_revcomp(p[u].key, k - 1) : p[u].key;
                                        return rcx_big_count(a.t, rcx_canonical(rcx_ext_kmer(ctx, left_ext, b, k), k));
                                    }
                                    return (int)f;
                                };
                                if (info & (1 << 14)) row[jr - 1] = (unsigned int)field(true, (info >> 10) & 3);
                                if (info & (1 << 15)) row[jr] = (unsigned int)field(false, (info >> 12) & 3);
                            }
                        }
                    }
                }
            }
            // GetKmerInformation takes min / median / max of the valid windows' counts with 0 -> 1: one pass normalises the
            // row in place (an invalid window becomes INT_MAX, beyond every pivot), so the selection passes below are a load,
            // a compare and three selects per window
            for (int j = 0; j < kcnt; ++j) {
                const unsigned int x = row[j];
                int v = (int)x;
                v = v == 0 ? 1 : v;
                const bool ok = !(x & TF_INVALID);
                nvalid += ok;
                vmin = (ok && v < vmin) ? v : vmin;
                vmax = (ok && v > vmax) ? v : vmax;
                row[j] = ok ? (unsigned int)v : 0x7fffffffu;
            }
            // element [nvalid / 2] of the ascending valid counts: bisection on the value, the bounds snap to values that
            // occur (every lane of the warp needs one, so each walks its own row)
            if (nvalid > 0) {
                const int rank = nvalid / 2;
                int lo = vmin, hi = vmax;
                while (lo < hi) {
                    const int mid = lo + (int)(((unsigned int)hi - (unsigned int)lo) >> 1);
                    int c = 0, maxle = lo, mingt = hi;
#pragma unroll 4
                    for (int j = 0; j < kcnt; ++j) {
                        const int v = (int)row[j];
                        const bool le = v <= mid;
                        c += le;
                        maxle = le ? max(maxle, v) : maxle;
                        mingt = le ? mingt : min(mingt, v);
                    }
                    if (c > rank) hi = maxle; else lo = mingt;
                }
                med = lo;
            }
        }
        if (have) {
            int l = 0, m = 0, h = 0;
            if (nvalid > 0) { l = vmin; h = vmax; m = med; }
            a.correction[r] = result; a.l[r] = l; a.m[r] = m; a.h[r] = h;
        }
    }
    n_rprobes = __reduce_add_sync(RCX_FULL, n_rprobes);
    n_rwindows = __reduce_add_sync(RCX_FULL, n_rwindows);
    if (lane == 0 && n_rprobes) atomicAdd(&a.counters[RCX_C_RESCAN_PROBES], (unsigned long long)n_rprobes);
    if (lane == 0 && n_rwindows) atomicAdd(&a.counters[RCX_C_RESCAN_WINDOWS], (unsigned long long)n_rwindows);
}

bool rcx_finalize_thread_supported(int max_len, int k) {
    const int kc = max_len - k + 1;
    return kc >= 1 && kc <= RCX_TS_MAX_KCNT && max_len <= 255 && tf_layout(max_len, k, max_len <= 128 ? TF_THREADS : 64).smem <= 200 * 1024;
}

template <int NW, int NT>
static cudaError_t launch_finalize_thread_nw(const RcxCorrectArgs& a, int sm_count, cudaStream_t s) {
    const TfLayout L = tf_layout(a.max_len, a.t.k, NT);
    cudaError_t e = cudaFuncSetAttribute(rcx_finalize_thread_kernel<NW, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rcx_finalize_thread_kernel<NW, NT>, NT, L.smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long want = (a.n + NT - 1) / NT;
    long long grid = (long long)sm_count * per_sm;  // persistent grid (the work-list length lives on the device)
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    rcx_finalize_thread_kernel<NW, NT><<<(unsigned int)grid, NT, L.smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t rcx_launch_finalize_thread(const RcxCorrectArgs& a, int sm_count, cudaStream_t s) {
    if (a.max_len <= 128) return launch_finalize_thread_nw<2, TF_THREADS>(a, sm_count, s);
    return launch_finalize_thread_nw<4, 64>(a, sm_count, s);  // 64-item tiles: more blocks fit the shared memory
}
