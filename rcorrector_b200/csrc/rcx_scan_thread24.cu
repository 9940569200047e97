// rcx_scan_thread24.cu — EXPERIMENTAL variant of the thread-per-read scan (rcx_scan_thread.cu, whose kernel body this file
// repeats) for the 64-read tiles of reads longer than 128 bases, selected by RCX_ROW24=1 only.  Not the default: it was
// measured at the very end of the round on one workload (profiles/r02_u_row24_experiment.md) and the GPU suite could not be
// rerun on it.
//
// What differs: the per-thread count rows in shared memory take 24 bits per k-window (two planes: low 16 bits, high 8 bits;
// 21-bit count + the three flags) instead of 32 -- 392 instead of 516 bytes per 150-bp read, which lets a 6th block (12 instead
// of 10 warps) onto an SM, where shared memory is the limit (profiles/r02_r_cfg3_full.md).  A count of 2^21 - 1 or more does not
// fit: the row keeps the all-ones ESCAPE value, the exact count goes straight to counts[] in global memory (where the write-out
// at the end of the tile would put it anyway), and the thread remembers that its row holds an escape: only such rows (a handful
// of poly-A / repeat k-mers of a deep data set) run the loops over the row in their checking version.
#include "rcx_internal.h"

namespace rcx24 {

#define TS_THREADS 128  // reads per tile for reads up to 128 bases; longer reads use 64-read tiles (more blocks fit the shared memory)
// TS_DEPTH (template parameter DEPTH): sector requests in flight per thread.  Measured for 100-bp reads (128-read tiles,
// 512 threads per SM) 2 / 4 / 8: 26.0 / 27.4 / 40.4 ms (registers and instructions against latency hiding) -> 2; for 150-bp
// reads (64-read tiles, 320 threads per SM) 2 / 4: 28.2 / 29.0 ms per 6 M reads -> 2 (profiles/r01_u_e2e_chunks.md).
#ifndef TS_ROTATE
// 1: rotating request pipeline (a sector is consumed one round after it was issued).  Measured equal or slightly slower
// (10 M x 100 bp: 22.98 -> 23.14 ms; 48 M x 150 bp: 224.6 -> 225.8 ms; 3 / 4 requests in flight: 24.4 / 26.8 ms,
// gpurun_out/r02_n2b): the scan is not waiting for its sectors, it is short of issue slots.
#define TS_ROTATE 0
#endif
#ifndef TS_DEPTH_DEFAULT
#define TS_DEPTH_DEFAULT 2
#endif
#define TS_LOOSE 0x80000000u   // flag bits stored with a count in shared memory
#define TS_STRICT 0x40000000u
#define TS_INVALID 0x20000000u
#define TS_CMASK 0x1fffffffu
// per-base byte in shared memory: code (bits 0-1) | not-ACGT (bit 2) | allow-Right (3) | allow-Left (4) | literal N (5) |
// strongly trusted position (6)
#define TB_BAD 4
#define TB_N 32
#define TB_STRONG 64

#define TS_BINS 256  // per-tile counting sort of the work items by their number of untrusted k-windows (<= RCX_TS_MAX_KCNT)
struct TsLayout {
    int nt;         // rows (threads of the block)
    int cw;         // words per 32-bit count row (odd)
    int lo_bytes, hi_bytes;  // 24-bit rows: bytes per row of the low-16 plane and of the high-8 plane (a multiple of 4, odd number of words)
    int bytes_cap;  // staged bases
    size_t smem;
};
// Count rows, one per thread: 32 bits per k-window (29-bit count + three flags; kept for reference, not instantiated here) or,
// W24, 24 bits in two planes (see the head of the file).
__host__ __device__ inline TsLayout ts_layout(int max_len, int k, int nt, bool w24) {
    TsLayout L;
    int kc = max_len - k + 1;
    if (kc < 1) kc = 1;
    L.nt = nt;
    L.cw = kc | 1;
    L.lo_bytes = 4 * (((2 * kc + 3) / 4) | 1);
    L.hi_bytes = 4 * (((kc + 3) / 4) | 1);
    L.bytes_cap = ((nt * max_len + 15) & ~15) + 32;
    const size_t row = w24 ? (size_t)(L.lo_bytes + L.hi_bytes) : sizeof(int) * (size_t)L.cw;
    L.smem = sizeof(uint32_t) * (nt + 4) + sizeof(int) * (16 + TS_BINS) + (size_t)nt * row + (size_t)L.bytes_cap;
    return L;
}
// flag bits and count mask of a row entry as the kernel sees it: the 32-bit rows' own layout, or the 24 bits of the two planes
template <bool W24>
struct TsBits {
    static constexpr unsigned int LOOSE = TS_LOOSE, STRICT = TS_STRICT, INVALID = TS_INVALID, CMASK = TS_CMASK;
    static constexpr int SIGN_SHIFT = 0;  // left shift that brings LOOSE to the sign bit
};
template <>
struct TsBits<true> {
    static constexpr unsigned int LOOSE = 0x00800000u, STRICT = 0x00400000u, INVALID = 0x00200000u, CMASK = 0x001fffffu;  // count CMASK = escape
    static constexpr int SIGN_SHIFT = 8;
};
template <bool W24>
struct TsRow;
template <>
struct TsRow<false> {
    unsigned int* p;
    __device__ __forceinline__ TsRow(unsigned char* rows, const TsLayout& L, int t) : p(reinterpret_cast<unsigned int*>(rows) + (size_t)t * L.cw) {}
    __device__ __forceinline__ unsigned int get(int j) const { return p[j]; }
    __device__ __forceinline__ void set_flags(int j, unsigned int f) const { p[j] = f; }         // count bits 0
    __device__ __forceinline__ void add_count(int j, unsigned int c) const { p[j] |= c; }         // once per entry, after set_flags
};
template <>
struct TsRow<true> {
    unsigned short* lo;
    unsigned char* hi;
    __device__ __forceinline__ TsRow(unsigned char* rows, const TsLayout& L, int t)
        : lo(reinterpret_cast<unsigned short*>(rows + (size_t)t * L.lo_bytes)), hi(rows + (size_t)L.nt * L.lo_bytes + (size_t)t * L.hi_bytes) {}
    __device__ __forceinline__ unsigned int get(int j) const { return (unsigned int)lo[j] | ((unsigned int)hi[j] << 16); }
    __device__ __forceinline__ void set_flags(int j, unsigned int f) const {
        lo[j] = 0;
        hi[j] = (unsigned char)(f >> 16);
    }
    __device__ __forceinline__ void add_count(int j, unsigned int c) const {
        lo[j] = (unsigned short)c;
        if (c >> 16) hi[j] = (unsigned char)(hi[j] | (c >> 16));
    }
};

struct TsPend {
    unsigned long long key;
    RcxEntryRegs d;
    int info;  // -1: empty;  else  j (right window index, 10 bits) | left base << 10 | right base << 12 | wantL << 14 | wantR << 15 | flip << 16
};

__device__ __forceinline__ bool ts_find_slow(const RcxTable& t, unsigned long long key, RcxEntryRegs& d) {
    const unsigned long long home = __umul64hi(rcx_hash(key), t.nlines << 2);
    for (unsigned int i = 1;; ++i) {
        rcx_issue(t, rcx_probe_sector(t, home, i), d);
        if (d.q0 == ~0ull) return false;
        if ((d.q0 & RCX_KEY_MASK) == key) return true;
        if (i >= 4 * t.nlines) return false;
    }
}
__device__ __forceinline__ bool ts_find(const RcxTable& t, unsigned long long key, RcxEntryRegs& d) {
    if ((d.q0 & RCX_KEY_MASK) == key) return true;
    if (d.q0 == ~0ull || !(d.q0 & RCX_OVERFLOW_BIT)) return false;
    return ts_find_slow(t, key, d);
}
// count of the extension by base b (caller's orientation) from a found entry
__device__ __forceinline__ int ts_field(const RcxTable& t, const RcxEntryRegs& d, unsigned long long key, bool flip, bool left, int b) {
    unsigned int f0, f1, f2, f3;
    rcx_group(d, left != flip, f0, f1, f2, f3);
    const int bb = flip ? 3 - b : b;
    const unsigned int f = bb == 0 ? f0 : bb == 1 ? f1 : bb == 2 ? f2 : f3;
    if (f == RCX_BIG) {
        const unsigned long long ctx = flip ? rcx_revcomp(key, t.k - 1) : key;
        return rcx_big_count(t, rcx_canonical(rcx_ext_kmer(ctx, left, b, t.k), t.k));
    }
    return (int)f;
}

__device__ __forceinline__ int ts_octave(int v) { return v <= 0 ? 0 : 32 - __clz(v); }
// count of row entry x = row.get(j).  ESC: the row holds escaped entries (24-bit rows only), gc = counts[] of the row's read
template <bool W24, bool ESC>
__device__ __forceinline__ int ts_count(unsigned int x, const int32_t* gc, int j) {
    int c = (int)(x & TsBits<W24>::CMASK);
    if (W24 && ESC) {
        if (c == (int)TsBits<W24>::CMASK) c = __ldcg(gc + j);  // (L2: the select below reads other lanes' rows)
    }
    return c;
}
struct TsPlain { static constexpr bool value = false; };
struct TsEscaped { static constexpr bool value = true; };

// Selection: smallest x with #{value <= x} > rank over a row; kind 0: loose-masked counts, kind 1: valid windows with 0 -> 1.
// Only a FEW lanes of a warp need one (reads without a drop, reads that leave at the triage):
// instead of every lane walking its own row while the others idle, the warp serves the requests one after the other,
// all 32 lanes counting a slice of the requester's row (conflict-free: consecutive words).  Must be called convergently.
template <int KIND, bool W24>
__device__ __forceinline__ int ts_select_coop(unsigned char* rows, const TsLayout& L, int warp, int lane, bool need, int kcnt, int rank, int lo, int hi,
                                              const int32_t* counts, uint32_t o, bool esc) {
    using TB = TsBits<W24>;
    int result = lo;
    __syncwarp();  // the rows were written by their owner lanes
    unsigned int pending = __ballot_sync(RCX_FULL, need);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const TsRow<W24> row(rows, L, warp * 32 + src);
        const int n = __shfl_sync(RCX_FULL, kcnt, src), rk = __shfl_sync(RCX_FULL, rank, src);
        const int32_t* gc = nullptr;
        bool src_esc = false;
        if (W24) {
            gc = counts + __shfl_sync(RCX_FULL, o, src);
            src_esc = __shfl_sync(RCX_FULL, (int)esc, src) != 0;  // (uniform over the warp)
        }
        int l = __shfl_sync(RCX_FULL, lo, src), h = __shfl_sync(RCX_FULL, hi, src);
        while (l < h) {
            const int mid = l + (int)(((unsigned int)h - (unsigned int)l) >> 1);
            int c = 0;
            if (W24 && src_esc) {
                for (int j = lane; j < n; j += 32) {
                    const unsigned int x = row.get(j);
                    const int v = ts_count<W24, true>(x, gc, j);
                    if (KIND == 0) c += ((x & TB::LOOSE) ? -1 : v) <= mid;
                    else if (!(x & TB::INVALID)) c += (v == 0 ? 1 : v) <= mid;
                }
            } else {
                for (int j = lane; j < n; j += 32) {
                    const unsigned int x = row.get(j);
                    // a loose entry (its flag shifted to the sign bit) reads as negative: like its -1, never above a pivot >= -1
                    if (KIND == 0) c += (int)((x & ~(TB::STRICT | TB::INVALID)) << TB::SIGN_SHIFT) <= mid * (1 << TB::SIGN_SHIFT);
                    else if (!(x & TB::INVALID)) { int v = (int)(x & TB::CMASK); v = v == 0 ? 1 : v; c += v <= mid; }
                }
            }
            c = __reduce_add_sync(RCX_FULL, c);
            if (c > rk) h = mid; else l = mid + 1;
        }
        if (lane == src) result = l;
    }
    return result;
}

template <int NW, int NT, int TS_DEPTH, bool W24>
__global__ void __launch_bounds__(NT, 512 / NT) rcx_scan_thread_kernel(RcxScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = a.t.k;
    const TsLayout LY = ts_layout(a.max_len, k, NT, W24);
    uint32_t* offs = reinterpret_cast<uint32_t*>(smem_raw);                       // [T + 1] (+3 pad)
    int* blk = reinterpret_cast<int*>(offs + NT + 4);                     // [16] block scratch
    int* bins = blk + 16;                                                 // [TS_BINS] work items of the tile per untrusted-window count
    unsigned char* cnt_all = reinterpret_cast<unsigned char*>(bins + TS_BINS);    // [T] count rows (TsRow)
    unsigned char* bytes = cnt_all + (size_t)NT * (W24 ? (size_t)(LY.lo_bytes + LY.hi_bytes) : sizeof(int) * (size_t)LY.cw);  // 16-byte aligned (all sizes above are multiples of 16)
    using TB = TsBits<W24>;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    const int loose_thr = (k / 2 > 7) ? k / 2 : 7;
    const double er = a.error_rate;
    const int K = a.max_fix_per_k;
    const int stride = rcx_state_stride(a.max_len, k);
    const int nseg_cap = a.max_len / k + 4;
    unsigned int probes = 0, toobig = 0;
    const long long ntiles = (a.n + NT - 1) / NT;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(blk + 12);  // transaction barrier of the tile copy (8-byte aligned)
    unsigned int tile_phase = 0;
    if (tid == 0) rcx_mbar_init(mbar, 1);
    __syncthreads();

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long r0 = tile * NT;
        const int nr = (int)((a.n - r0) < NT ? (a.n - r0) : NT);
        __syncthreads();  // previous tile fully written out
        for (int t = tid; t <= nr; t += NT) offs[t] = a.off[r0 + t];
        for (int t = tid; t < TS_BINS; t += NT) bins[t] = 0;
        __syncthreads();
        const unsigned long long g0 = offs[0], g1 = offs[nr];
        const unsigned long long a0 = g0 & ~15ull;
        {   // stage the tile's bases: ONE bulk copy (TMA, 1-D) of the 16-byte aligned body handed to the copy engine by thread 0
            // -- no registers, no staging loop -- and a byte tail (never reads past the batch); the block waits on the
            // transaction barrier
            unsigned long long e16 = (g1 + 15) & ~15ull;
            if (e16 > (a.total_bases & ~15ull)) e16 = a.total_bases & ~15ull;
            if (e16 < a0) e16 = a0;
            if (tid == 0) {
                const unsigned int nbytes = (unsigned int)(e16 - a0);
                rcx_fence_proxy_async();  // the previous tile's bytes were rewritten in place by the threads
                if (nbytes) rcx_bulk_load(bytes, a.seq + a0, nbytes, mbar);
                else rcx_mbar_arrive(mbar);
            }
            for (unsigned long long p = e16 + tid; p < g1; p += NT) bytes[p - a0] = (unsigned char)a.seq[p];
            rcx_mbar_wait(mbar, tile_phase);
            tile_phase ^= 1u;
        }
        __syncthreads();

        // ------------------------------------------------------------------ per-thread scan
        const long long r = r0 + tid;
        const bool have = tid < nr;
        const uint32_t o = have ? offs[tid] : 0;
        const int len = have ? (int)(offs[tid + 1] - o) : 0;
        const int kcnt = len - k + 1;
        unsigned char* B = bytes + (o - a0);
        const TsRow<W24> row(cnt_all, LY, tid);
        int32_t* gc = a.counts + o;  // (24-bit rows: exact value of an escaped count)
        bool esc = false;            // this row holds an escaped entry: its loops run in the checking version
        // runs body(TsEscaped) for a row with escapes, body(TsPlain) otherwise (the only version of the 32-bit rows)
        auto on_row = [&](auto body) {
            if (W24 && esc) body(TsEscaped());
            else body(TsPlain());
        };
        int thr = -1;
        uint8_t flags = RCX_F_REJECT;
        int nN = 0, nA = 0, nT = 0;
        if (have && kcnt > 0) {
            unsigned long long F = 0;
            int run = 0, run_prev = 0, wA = 0, wT = 0, left = 0;
            int i = 0;
            auto append = [&]() {
                const unsigned char ch = B[i];
                const unsigned int idx = (unsigned int)ch - 'A';
                const bool ok = idx < 26u && ((0x00080045u >> idx) & 1u);  // A, C, G, T
                const int code = ok ? (((ch >> 1) ^ (ch >> 2)) & 3) : 3;   // A0 C1 G2 T3, anything else coded as 3 (KmerCode.cpp:12-13)
                nN += ch == 'N';
                nA += ch == 'A';
                nT += ch == 'T';
                B[i] = (unsigned char)(code | (ok ? 0 : TB_BAD) | (ch == 'N' ? TB_N : 0));
                wA += ch == 'A';
                wT += ch == 'T';
                if (i >= k) {  // base leaving the window (already converted)
                    const unsigned char q = B[i - k];
                    wA -= (q & 7) == 0;
                    wT -= (q & 7) == 3;
                }
                left = (int)((F >> (2 * (k - 1))) & 3);
                F = ((F << 2) | (unsigned long long)code) & kmask;
                run_prev = run;
                run = ok ? run + 1 : 0;
                ++i;
            };
            while (i < k - 1) append();
            // advance to the next request (one per two windows) and issue it; p.info = -1 when the read ends first
            auto issue = [&](TsPend& p) {
                p.info = -1;
                bool trig = false;
                int j = 0;
                while (i < len && !trig) {
                    append();
                    j = i - k;  // window that ends at the base just appended
                    const bool lo_ = wA >= k - loose_thr || wT >= k - loose_thr;  // IsPolyA(seq + j, k, max(7, k/2))
                    const bool st_ = wA >= k - 2 || wT >= k - 2;                  // IsPolyA(seq + j, k, 2)
                    row.set_flags(j, (lo_ ? TB::LOOSE : 0u) | (st_ ? TB::STRICT : 0u) | (run >= k ? 0u : TB::INVALID));
                    trig = (j & 1) || i == len;
                }
                if (trig) {
                    // context = the k-1 bases before the newest one: window j-1 is its left extension (by the base that
                    // just left F), window j its right extension (by the newest base)
                    const bool wantL = (j & 1) && run_prev >= k;
                    const bool wantR = run >= k;
                    if (wantL || wantR) {
                        const unsigned long long ctx = F >> 2;
                        const unsigned long long rc = rcx_revcomp(ctx, k - 1);
                        const bool flip = rc < ctx;
                        p.key = flip ? rc : ctx;
                        p.info = j | (left << 10) | ((int)(F & 3) << 12) | (wantL ? 1 << 14 : 0) | (wantR ? 1 << 15 : 0) | (flip ? 1 << 16 : 0);
                        rcx_issue(a.t, __umul64hi(rcx_hash(p.key), a.t.nlines << 2), p.d);
                        ++probes;
                    }
                }
            };
            auto consume = [&](TsPend& p) {
                const int info = p.info;
                p.info = -1;
                if (info >= 0 && ts_find(a.t, p.key, p.d)) {
                    const int j = info & 1023;
                    const bool flip = (info >> 16) & 1;
                    if (info & (1 << 14)) {
                        const int c = ts_field(a.t, p.d, p.key, flip, true, (info >> 10) & 3);
                        toobig |= (unsigned int)c > TS_CMASK;
                        if (W24 && (unsigned int)c >= TB::CMASK) { gc[j - 1] = c; row.add_count(j - 1, TB::CMASK); esc = true; }
                        else row.add_count(j - 1, (unsigned int)c & TB::CMASK);
                    }
                    if (info & (1 << 15)) {
                        const int c = ts_field(a.t, p.d, p.key, flip, false, (info >> 12) & 3);
                        toobig |= (unsigned int)c > TS_CMASK;
                        if (W24 && (unsigned int)c >= TB::CMASK) { gc[j] = c; row.add_count(j, TB::CMASK); esc = true; }
                        else row.add_count(j, (unsigned int)c & TB::CMASK);
                    }
                }
            };
            TsPend p[TS_DEPTH];
#pragma unroll
            for (int u = 0; u < TS_DEPTH; ++u) p[u].info = -1;
#if TS_ROTATE
            // rotating pipeline: a request is consumed one round AFTER it was issued -- every sector has the time of
            // TS_DEPTH - 1 further requests (and of the other warps) to arrive, instead of being waited for right after the issue
            while (i < len) {
#pragma unroll
                for (int u = 0; u < TS_DEPTH; ++u) {
                    consume(p[u]);
                    issue(p[u]);
                }
            }
#else
            while (i < len) {
#pragma unroll
                for (int u = 0; u < TS_DEPTH; ++u) issue(p[u]);
#pragma unroll
                for (int u = 0; u < TS_DEPTH; ++u) consume(p[u]);
            }
#endif
#pragma unroll
            for (int u = 0; u < TS_DEPTH; ++u) consume(p[u]);
        } else if (have) {
            for (int i = 0; i < len; ++i) {  // shorter than k: only the per-base conversion (the bytes go out as per-base info)
                const unsigned char ch = B[i];
                const unsigned int idx = (unsigned int)ch - 'A';
                const bool ok = idx < 26u && ((0x00080045u >> idx) & 1u);
                B[i] = (unsigned char)((ok ? (((ch >> 1) ^ (ch >> 2)) & 3) : 3) | (ok ? 0 : TB_BAD) | (ch == 'N' ? TB_N : 0));
            }
        }

        // ------------------------------------------------------------------ GetStrongTrustedThreshold (:1507-1563)
        int minc = 0x7fffffff, nvalid = 0, vmin = 0x7fffffff, vmax = 0;
        bool any_strict = false, need_sel = false;
        int sel_rank = 0, sel_lo = 0, sel_hi = 0;
        if (have && kcnt > 0 && !(nN > 5 || nA > len - k || nT > len - k)) {
            unsigned int present = 0;
            int nonpos = 0, gmin = 0x7fffffff, gmax = -1;
            on_row([&](auto E) {
                for (int j = 0; j < kcnt; ++j) {
                    const unsigned int x = row.get(j);
                    const int c = ts_count<W24, decltype(E)::value>(x, gc, j);
                    const int m = (x & TB::LOOSE) ? -1 : c;
                    present |= 1u << ts_octave(m);
                    nonpos += m <= 0;
                    gmin = m < gmin ? m : gmin;
                    gmax = m > gmax ? m : gmax;
                    minc = c < minc ? c : minc;
                    any_strict |= (x & TB::STRICT) != 0;
                }
            });
            bool found = false;
            int below = 0;
            unsigned int leftm = present;
            while (leftm) {
                const int b = 31 - __clz(leftm);
                const unsigned int lower = leftm & ~(1u << b);
                if (b < 4 || lower == 0) break;  // values <= 10 cannot be a drop; the smallest element has no predecessor
                const int lo_b = 1 << (b - 1);   // octave b = [2^(b-1), 2^b)
                int mn = 0x7fffffff, mx = (int)0x80000000;
                on_row([&](auto E) {
                    for (int j = 0; j < kcnt; ++j) {
                        const unsigned int x = row.get(j);
                        const int m = (x & TB::LOOSE) ? -1 : ts_count<W24, decltype(E)::value>(x, gc, j);
                        if ((unsigned int)(m - lo_b) < (unsigned int)lo_b) mn = m < mn ? m : mn;  // smallest value of this octave
                        else if (m < lo_b) mx = m > mx ? m : mx;                                  // largest value below it
                    }
                });
                if (mn <= 10) break;  // everything further down is smaller still
                if (mn > 2 * mx) { found = true; thr = mn; below = mx; break; }
                leftm = lower;
            }
            if (!found) { need_sel = true; sel_rank = (nonpos + kcnt - 1) / 2; sel_lo = gmin; sel_hi = gmax; }  // median of the positive entries
            flags = found ? (uint8_t)(RCX_F_DROP | (below == 2 ? RCX_F_PREV2 : 0)) : (uint8_t)0;
        }
        {
            const int med = ts_select_coop<0, W24>(cnt_all, LY, warp, lane, need_sel, kcnt, sel_rank, sel_lo, sel_hi, a.counts, o, esc);
            if (need_sel) thr = med;
        }
        if (have) { a.thr[r] = thr; a.flags[r] = flags; }

        // ------------------------------------------------------------------ triage
        int result = -3;  // -3: no read / scan only; -2: work item; else final
        bool need_info = false;
        int strong = 0, trust = 0;
        const int mate_thr = __shfl_xor_sync(RCX_FULL, thr, 1);  // mates are records 2i, 2i+1 = neighbouring lanes
        if (have && a.triage) {
            if (kcnt <= 0) {  // :713 -> -1 ; GetKmerInformation gives 0,0,0
                a.correction[r] = -1; a.l[r] = 0; a.m[r] = 0; a.h[r] = 0;
                result = -1;
            } else {
                if (flags & RCX_F_REJECT) result = -1;
                else {
                    const int pair_t = a.mode == 1 ? (thr < mate_thr ? thr : mate_thr) : -1;
                    // A.3 (:772-842)
                    strong = thr;
                    bool flag = false;
                    trust = rcx_bound_i(strong, er);
                    if ((flags & RCX_F_DROP) && strong >= 20 && (flags & RCX_F_PREV2) && trust < 3) { flag = true; trust = 3; }
                    if (pair_t >= 1 && strong > pair_t) {
                        if (!flag || pair_t < 20) trust = rcx_bound_i(pair_t, er);
                        strong = pair_t;
                    }
                    if (trust < 2) trust = 2;
                    result = (!any_strict && minc >= strong) ? 0 : -2;  // every k-mer trusted (:875, :1110)
                }
                if (result != -2) {  // GetKmerInformation (:1567-1602) on the unchanged read
                    on_row([&](auto E) {
                        for (int j = 0; j < kcnt; ++j) {
                            const unsigned int x = row.get(j);
                            if (!(x & TB::INVALID)) {
                                int v = ts_count<W24, decltype(E)::value>(x, gc, j);
                                v = v == 0 ? 1 : v;
                                ++nvalid;
                                vmin = v < vmin ? v : vmin;
                                vmax = v > vmax ? v : vmax;
                            }
                        }
                    });
                    need_info = true;
                }
            }
        }
        {   // the median of the tags, served warp-cooperatively for the lanes that leave here
            const int med = ts_select_coop<1, W24>(cnt_all, LY, warp, lane, need_info && nvalid > 0, kcnt, nvalid / 2, vmin, vmax, a.counts, o, esc);
            if (need_info) {
                int l = 0, m = 0, h = 0;
                if (nvalid > 0) { l = vmin; h = vmax; m = med; }
                a.correction[r] = result; a.l[r] = l; a.m[r] = m; a.h[r] = h;
            }
        }
        // work-list slots: one atomic per block and class.  Items with many untrusted k-windows (more than one isolated
        // error explains) are the candidates for long searches: they are filed from the far end of the list and searched first.
        // The others are filed sorted by that number within the tile (counting sort in shared memory): the search kernel takes
        // them 32 consecutive slots per warp, and items with the same number of untrusted windows take nearly the same walk.
        const bool is_work = result == -2;
        bool heavy = false, vheavy = false;
        int nun = 0, pos_in_bin = 0;
        if (is_work) {
            on_row([&](auto E) {
                for (int j = 0; j < kcnt; ++j) {
                    const unsigned int x = row.get(j);
                    nun += ts_count<W24, decltype(E)::value>(x, gc, j) < strong || (x & TB::STRICT) != 0;
                }
            });
            heavy = a.heavy_windows > 0 && nun >= a.heavy_windows;
            vheavy = heavy && a.vheavy_windows > 0 && nun >= a.vheavy_windows;
            if (nun >= TS_BINS) nun = TS_BINS - 1;
            if (!heavy) pos_in_bin = atomicAdd(&bins[nun], 1);
        }
        const unsigned int wbh = __ballot_sync(RCX_FULL, is_work && heavy), wbv = __ballot_sync(RCX_FULL, vheavy);
        if (lane == 0) { blk[4 + warp] = __popc(wbh); blk[warp] = __popc(wbv); }  // (blk[12..13] is the tile copy's barrier)
        __syncthreads();
        if (warp == 0) {  // exclusive prefix over the bins (TS_BINS / 32 per lane), total = light items of the tile
            int v[TS_BINS / 32], sum = 0;
#pragma unroll
            for (int q = 0; q < TS_BINS / 32; ++q) { v[q] = bins[lane * (TS_BINS / 32) + q]; sum += v[q]; }
            int incl = sum;
            for (int o2 = 1; o2 < 32; o2 <<= 1) { const int y = __shfl_up_sync(RCX_FULL, incl, o2); if (lane >= o2) incl += y; }
            int run = incl - sum;
#pragma unroll
            for (int q = 0; q < TS_BINS / 32; ++q) { bins[lane * (TS_BINS / 32) + q] = run; run += v[q]; }
            const int tot = __shfl_sync(RCX_FULL, incl, 31);
            if (lane == 0) {
                int toth = 0, totv = 0;
                for (int w = 0; w < NT / 32; ++w) {
                    const int ch = blk[4 + w]; blk[4 + w] = toth; toth += ch;
                    const int cv = blk[w]; blk[w] = totv; totv += cv;
                }
                blk[8] = tot ? (int)atomicAdd(&a.counters[RCX_C_NWORK], (unsigned long long)tot) : 0;
                blk[9] = toth ? (int)atomicAdd(&a.counters[RCX_C_NWORKH], (unsigned long long)toth) : 0;
                blk[10] = totv ? (int)atomicAdd(&a.counters[RCX_C_NVH], (unsigned long long)totv) : 0;  // (counted always: the host picks the mode of the NEXT batch from it)
            }
        }
        __syncthreads();

        // ------------------------------------------------------------------ A.4 steps 1-4 for the work items (:863-1046)
        // Bit masks in registers (NW 64-bit words cover the read) keep the lanes of a warp in step: one uniform pass over
        // the k-mers builds them, runs and islands are then found with ffs/clz instead of per-element loops.
        if (is_work) {
            const long long item = heavy ? a.n - 1 - ((long long)(unsigned int)blk[9] + blk[4 + warp] + __popc(wbh & ((1u << lane) - 1)))
                                         : (long long)(unsigned int)blk[8] + bins[nun] + pos_in_bin;
            a.worklist[item] = (uint32_t)r;
            if (vheavy && a.vh_mode) a.vhlist[(unsigned int)blk[10] + blk[warp] + __popc(wbv & ((1u << lane) - 1))] = (uint32_t)item;
            int* st = a.state + item * stride;
            RcxMask<NW> T, S, LOWM, M;  // trusted k-mers (:875), strict poly-A k-mers, counts <= 2 && < trust, trusted positions
            T.clear(); S.clear(); LOWM.clear(); M.clear();
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                unsigned long long tw = 0, sw = 0, lw = 0;
                const int jn = kcnt - 64 * w < 64 ? kcnt - 64 * w : 64;
                on_row([&](auto E) {
                    for (int b = 0; b < jn; ++b) {
                        const unsigned int x = row.get(64 * w + b);
                        const int c = ts_count<W24, decltype(E)::value>(x, gc, 64 * w + b);
                        const bool strict = (x & TB::STRICT) != 0;
                        tw |= (unsigned long long)(c >= strong && !strict) << b;
                        sw |= (unsigned long long)strict << b;
                        lw |= (unsigned long long)(c <= 2 && c < trust) << b;
                    }
                });
                T.m[w] = tw; S.m[w] = sw; LOWM.m[w] = lw;
            }
            int longest = 0, tstart = 0, tend = -1;
            {
                int pf = 0, pt = -1;
                bool havei = false;
                int p = 0;
                while (p < kcnt) {
                    const int s = T.next(p, kcnt, true);
                    if (s >= kcnt) break;
                    const int e = T.next(s, kcnt, false);
                    const int rl = e - s;
                    if (rl > longest) { longest = rl; tstart = s; tend = e - 1; }
                    if (rl >= 2) {
                        int cf = s, ct = e - 1;
                        if (havei) {
                            if (cf <= pt + k) {  // :936-964
                                const int len1 = pt - pf, len2 = ct - cf, overlap = pt + k - cf;
                                if (overlap <= 3 && LOWM.any(pt + 1, cf - 1)) {
                                    if (len1 < len2) pt -= overlap + 1; else cf += overlap + 1;
                                }
                            }
                            if (pf <= pt) M.set(pf, pt + k - 1);
                        }
                        pf = cf; pt = ct; havei = true;
                    }
                    p = e;
                }
                if (havei && pf <= pt) M.set(pf, pt + k - 1);
            }
            // islands of trusted positions (:977-1007); none: the longest trusted run stands in (without becoming "strong")
            RcxMask<NW> I = M;
            if (!M.nonzero()) I.set(tstart, tend + k - 1);
            // seed k-mer of a search = Restart + k Appends of seq[j0 .. j0+k) (j0 may be -1, see :1139-1142)
            auto seed = [&](int j0, unsigned long long& code, int& inv) {
                code = 0; inv = -1;
                for (int q = 0; q < k; ++q) {
                    const int p = j0 + q;
                    const unsigned char bq = (p >= 0 && p < len) ? B[p] : 0;
                    if (inv != -1) ++inv;
                    code = (code << 2) | (unsigned long long)(bq & 3);
                    if (p >= 0 && (bq & TB_BAD)) inv = 0;
                }
                code &= kmask;
            };
            int nseg = 0;
            {
                int ifrom = I.next(0, len, true);
                int ito = ifrom < len ? I.next(ifrom, len, false) - 1 : len - 1;
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
                        if (nseg < nseg_cap) {
                            const bool right = la >= ra;  // :1136
                            const int extend = right ? (sto == len ? 0 : k - 1) : (sf == 0 ? 0 : k - 1);
                            unsigned long long code;
                            int inv;
                            seed(right ? sf - k : sto + 1, code, inv);
                            int max_fix = (sto - sf + 1) * K / k * 2 + 1;  // :1128-1131
                            if (max_fix < K) max_fix = K;
                            int* sg = st + ST_HDR + ST_SEGW * nseg;
                            sg[0] = (right ? sf : sto) | ((right ? sto + extend : sf - extend) << 16);
                            sg[1] = (right ? 1 : 0) | (max_fix << 1);
                            sg[2] = (int)(unsigned int)code;
                            sg[3] = (int)(unsigned int)(code >> 32);
                            sg[4] = inv;
                            sg[5] = -1;  // top2FixBottleNeck[1], filled by the search
                        }
                        ++nseg;
                    }
                    if (ifrom >= len) break;
                    prev_to = ito; prev_len = ito - ifrom + 1; first = false;
                    ifrom = I.next(ito + 1, len, true);
                    ito = ifrom < len ? I.next(ifrom, len, false) - 1 : len - 1;
                }
            }
            // per-base info for the search (:343 / :577): substitutions are allowed on untrusted positions outside strict
            // poly-A k-mers: Right tests k-mer i-k+1, Left k-mer i
            RcxMask<NW> SR = S;
            SR.shl(k - 1);
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const unsigned long long ar = ~M.m[w] & ~SR.m[w], al = ~M.m[w] & ~S.m[w];
                const int in = len - 64 * w < 64 ? len - 64 * w : 64;
                for (int b = 0; b < in; ++b) {
                    const int i = 64 * w + b;
                    B[i] = (unsigned char)((B[i] & 7) | (((ar >> b) & 1) << 3) | (((al >> b) & 1) << 4));
                }
            }
            st[ST_STATUS] = (vheavy && a.vh_mode) ? RCX_ST_SEARCH_VH : RCX_ST_SEARCH; st[ST_TOTALFIX] = rcx_force_bound(strong, er); st[ST_BADSEG] = 0;
#ifdef RCX_DIAG
            st[ST_BADSEG] = nun;  // (in: the search kernel reads it at the fetch and overwrites it at the item's end)
#endif
            st[ST_STRONG] = strong; st[ST_TRUST] = trust; st[ST_NSEG] = nseg < nseg_cap ? nseg : nseg_cap;
            st[ST_OFF] = (int)o; st[ST_LEN] = len;
        }
        __syncthreads();

        // ------------------------------------------------------------------ coalesced write-out: counts[], per-base info, cleared fix[]
        for (int t = warp; t < nr; t += NT / 32) {
            const uint32_t ot = offs[t];
            const int kc = (int)(offs[t + 1] - ot) - k + 1;
            const TsRow<W24> rw(cnt_all, LY, t);
            // (with the correction kernels downstream, bit 31 tells them that the window holds a non-ACGT base)
            for (int j = lane; j < kc; j += 32) {
                const unsigned int x = rw.get(j);
                if (W24 && (x & TB::CMASK) == TB::CMASK) continue;  // escaped: the exact count is already there
                a.counts[ot + j] = (int)((x & TB::CMASK) | ((a.triage && (x & TB::INVALID)) ? RCX_COUNT_INVALID : 0u));
            }
        }
        if (a.triage) {
            for (unsigned long long p = g0 + tid; p < g1; p += NT) a.posbuf[p] = bytes[p - a0] & 31;
            if (a.compact_fix) {
                if (tid < nr) a.fixlist[offs[tid]] = 0;  // no edits yet: the search kernel leaves its edit list here (first word = count)
            } else {
                for (unsigned long long p = g0 + tid; p < g1; p += NT) a.fixbuf[p] = -1;
            }
        }
    }
    probes = __reduce_add_sync(RCX_FULL, probes);
    toobig = __reduce_or_sync(RCX_FULL, toobig);
    if (lane == 0 && probes) atomicAdd(&a.counters[RCX_C_SCAN_PROBES], (unsigned long long)probes);
    if (lane == 0 && toobig) atomicAdd(&a.counters[RCX_C_UNSUPPORTED], 1ull);  // a count >= 2^29 (the host routes such tables to the warp-per-read kernels): flagged, not hidden
}

template <int NW, int NT, int DEPTH, bool W24>
static cudaError_t launch_scan_thread_nw(const RcxScanArgs& a, int sm_count, cudaStream_t s) {
    const TsLayout L = ts_layout(a.max_len, a.t.k, NT, W24);
    cudaError_t e = cudaFuncSetAttribute(rcx_scan_thread_kernel<NW, NT, DEPTH, W24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rcx_scan_thread_kernel<NW, NT, DEPTH, W24>, NT, L.smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long want = (a.n + NT - 1) / NT;
    long long grid = (long long)sm_count * per_sm;  // persistent grid: a multiple of the SM count
    if (grid > want) grid = want;
    rcx_scan_thread_kernel<NW, NT, DEPTH, W24><<<(unsigned int)grid, NT, L.smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace rcx24

cudaError_t rcx_launch_scan_thread24(const RcxScanArgs& a, int sm_count, cudaStream_t s) {
    if (a.n <= 0) return cudaSuccess;
    return rcx24::launch_scan_thread_nw<4, 64, TS_DEPTH_DEFAULT, true>(a, sm_count, s);  // 128 < max_len <= 255
}
