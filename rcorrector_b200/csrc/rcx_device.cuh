// rcx_device.cuh — device-side building blocks shared by the sm_100a kernels:
// the k-mer count table probe, the KmerCode state machine, the per-read
// bit-plane representation a warp builds in shared memory, FP64 GetBound and a
// warp bitonic sort.
//
// Reference behaviour restated here (no reference code is used):
//   KmerCode.h:7-89 / KmerCode.cpp:7-52  (2-bit code + single invalidPos)
//   Store.h:59-66                        (GetCount: 0 if invalid or absent)
//   ErrorCorrection.cpp:139-142          (GetBound, FP64 without FMA)
//   ErrorCorrection.cpp:53-71            (IsPolyA)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define RCX_FULL 0xffffffffu
#define RCX_EMPTY_KEY 0xffffffffffffffffull  // never canonical for k <= 32 (revcomp of T..T is 0)
#define RCX_INF 1000000000                   // utils.h:10
#define RCX_MAX_TRIAL 1025                   // ErrorCorrection.cpp:7
#define RCX_COUNT_INVALID 0x80000000u        // counts[] between the thread-per-read kernels: the window holds a non-ACGT base

// ----------------------------------------------------------------------------- table
// The table is keyed by (k-1)-mer CONTEXTS, not by k-mers.  One entry = 32 bytes = ONE L2 sector:
//     u64 key | u32 cw[6]
// key = canonical (k-1)-mer code (<= 62 bits) | overflow << 63; empty = ~0.
// cw[0..2] = RIGHT group: counts of the four k-mers  key+b  (b = A,C,G,T appended),
// cw[3..5] = LEFT  group: counts of the four k-mers  b+key  (b prepended);
// a group packs four 24-bit counts: word i = count_i | (byte i of count_3) << 24.  A count >= 0xFFFFFF is stored as
// 0xFFFFFF and resolved exactly through a (tiny, sorted) side list of canonical k-mers.
// Every stored k-mer appears twice: as a right extension of its prefix context and as a left extension of its suffix
// context (both normalised to the canonical strand of the context).
//
// Why this shape (measured on the B200, profiles/r01_f_membench*): (1) an L2 miss always fetches the whole 128-byte
// line from HBM, whatever the request size or cudaLimitMaxL2FetchGranularity; (2) independent random requests over a
// footprint > 1 GiB top out at ~45 G requests/s (TLB-miss / DRAM-line rate), whether a request is 16, 32 or 2 x 32
// bytes.  So the currency is REQUESTS, and a request should answer as many of the path's questions as possible:
//   * the scan needs count(window j) for consecutive j: windows j-1 and j share the context seq[j .. j+k-1), so ONE
//     sector answers two windows (left extension by seq[j-1], right extension by seq[j+k-1]);
//   * a search node needs the counts of the four one-base extensions of its k-mer (InferPosThreshold, the keep-base
//     test and the three alternatives): they share one context, so ONE sector answers the node (the k-mer-keyed
//     table of the first round needed four).
// Probe sequence of a key with home sector h: the four sectors of h's 128-byte line first (h, h^1, h^2, h^3: a second
// probe is an L2 hit, the line is already there), then the following lines.  Keys are placed home-first (every home
// sector that has a candidate gets one), the overflow bit of a home sector says that a key homed there lives further
// on, so a lookup -- hit or miss -- ends at its home sector unless that bit is set.
struct __align__(32) RcxEntry {
    unsigned long long key;
    unsigned int cw[6];
};
#define RCX_KEY_MASK 0x7fffffffffffffffull
#define RCX_OVERFLOW_BIT 0x8000000000000000ull
#define RCX_BIG 0xffffffu
struct RcxTable {
    const RcxEntry* e;
    unsigned long long nlines;   // 128-byte lines; sectors = 4 * nlines
    int k;
    const unsigned long long* big_keys;  // sorted canonical k-mers whose count >= RCX_BIG
    const unsigned int* big_counts;
    int nbig;
};
// an entry held in registers (one 256-bit load)
struct RcxEntryRegs {
    unsigned long long q0, q1, q2, q3;  // key, cw0|cw1<<32, cw2|cw3<<32, cw4|cw5<<32
};

__device__ __forceinline__ unsigned long long rcx_hash(unsigned long long x) {
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}

// reverse complement of an n-base code (A0 C1 G2 T3 => complement is 3-x = ~x on 2 bits), 1 <= n <= 32
__device__ __forceinline__ unsigned long long rcx_revcomp(unsigned long long code, int n) {
    unsigned long long r = __brevll(code);                                                  // reverses bits, also inside pairs
    r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);            // restore pair order
    return (~r) >> (64 - 2 * n);
}
__device__ __forceinline__ unsigned long long rcx_canonical(unsigned long long code, int k) {
    unsigned long long rc = rcx_revcomp(code, k);
    return rc < code ? rc : code;
}

// sector index of probe number i of a key whose home sector is h
__device__ __forceinline__ unsigned long long rcx_probe_sector(const RcxTable& t, unsigned long long h, unsigned int i) {
    unsigned long long line = (h >> 2) + (i >> 2);
    if (line >= t.nlines) line -= t.nlines;
    return (line << 2) | ((h ^ i) & 3);
}
// issue the 256-bit load of one entry (LDG.E.256 on sm_100a; read-only path, no L1 allocation)
__device__ __forceinline__ void rcx_issue(const RcxTable& t, unsigned long long sector, RcxEntryRegs& d) {
    const RcxEntry* p = t.e + sector;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(d.q0), "=l"(d.q1), "=l"(d.q2), "=l"(d.q3)
                 : "l"(p));
}
// one context lookup in flight
struct RcxLookup {
    unsigned long long key, home;
    bool flip;  // the caller's orientation is the reverse complement of the stored key
    RcxEntryRegs d;
};
// ctx = code of the (k-1)-mer in the caller's orientation
__device__ __forceinline__ void rcx_lookup_begin(const RcxTable& t, unsigned long long ctx, RcxLookup& L) {
    const unsigned long long rc = rcx_revcomp(ctx, t.k - 1);
    L.flip = rc < ctx;
    L.key = L.flip ? rc : ctx;
    L.home = __umul64hi(rcx_hash(L.key), t.nlines << 2);  // uniform over the sectors without a power-of-two table
    rcx_issue(t, L.home, L.d);
}
// finish: true if the context is stored (L.d then holds its entry)
__device__ __forceinline__ bool rcx_lookup_end(const RcxTable& t, RcxLookup& L) {
    if ((L.d.q0 & RCX_KEY_MASK) == L.key) return true;
    if (L.d.q0 == ~0ull || !(L.d.q0 & RCX_OVERFLOW_BIT)) return false;
    for (unsigned int i = 1;; ++i) {  // rare: a key homed here lives further along the probe sequence
        rcx_issue(t, rcx_probe_sector(t, L.home, i), L.d);
        if (L.d.q0 == ~0ull) return false;
        if ((L.d.q0 & RCX_KEY_MASK) == L.key) return true;
        if (i >= 4 * t.nlines) return false;
    }
}
// finish a lookup whose home sector was loaded into d: true if `key` is stored (d then holds its entry)
static __device__ __noinline__ bool rcx_find_slow(const RcxTable& t, unsigned long long key, RcxEntryRegs& d) {
    const unsigned long long home = __umul64hi(rcx_hash(key), t.nlines << 2);
    for (unsigned int i = 1;; ++i) {  // rare: a key homed there lives further along the probe sequence
        rcx_issue(t, rcx_probe_sector(t, home, i), d);
        if (d.q0 == ~0ull) return false;
        if ((d.q0 & RCX_KEY_MASK) == key) return true;
        if (i >= 4 * t.nlines) return false;
    }
}
__device__ __forceinline__ bool rcx_find(const RcxTable& t, unsigned long long key, RcxEntryRegs& d) {
    if ((d.q0 & RCX_KEY_MASK) == key) return true;
    if (d.q0 == ~0ull || !(d.q0 & RCX_OVERFLOW_BIT)) return false;
    return rcx_find_slow(t, key, d);
}
__device__ __forceinline__ unsigned long long rcx_home(const RcxTable& t, unsigned long long key) { return __umul64hi(rcx_hash(key), t.nlines << 2); }
// the four 24-bit fields of the stored RIGHT (left = false) or LEFT group
__device__ __forceinline__ void rcx_group(const RcxEntryRegs& d, bool left, unsigned int& f0, unsigned int& f1, unsigned int& f2, unsigned int& f3) {
    const unsigned int x0 = left ? (unsigned int)(d.q2 >> 32) : (unsigned int)d.q1;
    const unsigned int x1 = left ? (unsigned int)d.q3 : (unsigned int)(d.q1 >> 32);
    const unsigned int x2 = left ? (unsigned int)(d.q3 >> 32) : (unsigned int)d.q2;
    f0 = x0 & RCX_BIG; f1 = x1 & RCX_BIG; f2 = x2 & RCX_BIG;
    f3 = (x0 >> 24) | ((x1 >> 24) << 8) | ((x2 >> 24) << 16);
}
// exact count of a canonical k-mer whose 24-bit field saturated (binary search in the side list)
static __device__ __noinline__ int rcx_big_count(const RcxTable& t, unsigned long long canon) {
    int lo = 0, hi = t.nbig - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const unsigned long long v = t.big_keys[mid];
        if (v == canon) return (int)t.big_counts[mid];
        if (v < canon) lo = mid + 1; else hi = mid - 1;
    }
    return (int)RCX_BIG;
}
// k-mer formed by extending ctx (caller's orientation) with base b on the right / left
__device__ __forceinline__ unsigned long long rcx_ext_kmer(unsigned long long ctx, bool left, int b, int k) {
    return left ? (ctx | ((unsigned long long)b << (2 * (k - 1)))) : ((ctx << 2) | (unsigned long long)b);
}
// counts of the four extensions ctx+b (left = false) or b+ctx (left = true), b in the caller's orientation
__device__ __forceinline__ void rcx_counts4(const RcxTable& t, unsigned long long ctx, bool left, const RcxLookup& L, bool found,
                                            int& c0, int& c1, int& c2, int& c3) {
    c0 = c1 = c2 = c3 = 0;
    if (!found) return;
    unsigned int f0, f1, f2, f3;
    rcx_group(L.d, left != L.flip, f0, f1, f2, f3);
    c0 = (int)(L.flip ? f3 : f0); c1 = (int)(L.flip ? f2 : f1); c2 = (int)(L.flip ? f1 : f2); c3 = (int)(L.flip ? f0 : f3);
    if (((f0 + 1) | (f1 + 1) | (f2 + 1) | (f3 + 1)) & 0x1000000u) {  // some field == RCX_BIG
        if (c0 == (int)RCX_BIG) c0 = rcx_big_count(t, rcx_canonical(rcx_ext_kmer(ctx, left, 0, t.k), t.k));
        if (c1 == (int)RCX_BIG) c1 = rcx_big_count(t, rcx_canonical(rcx_ext_kmer(ctx, left, 1, t.k), t.k));
        if (c2 == (int)RCX_BIG) c2 = rcx_big_count(t, rcx_canonical(rcx_ext_kmer(ctx, left, 2, t.k), t.k));
        if (c3 == (int)RCX_BIG) c3 = rcx_big_count(t, rcx_canonical(rcx_ext_kmer(ctx, left, 3, t.k), t.k));
    }
}
// count of the single extension by base b
__device__ __forceinline__ int rcx_count1(const RcxTable& t, unsigned long long ctx, bool left, int b, const RcxLookup& L, bool found) {
    if (!found) return 0;
    unsigned int f0, f1, f2, f3;
    rcx_group(L.d, left != L.flip, f0, f1, f2, f3);
    const int bb = L.flip ? 3 - b : b;
    const unsigned int f = bb == 0 ? f0 : bb == 1 ? f1 : bb == 2 ? f2 : f3;
    if (f == RCX_BIG) return rcx_big_count(t, rcx_canonical(rcx_ext_kmer(ctx, left, b, t.k), t.k));
    return (int)f;
}
// Store::GetCount of a valid k-mer given by its forward code: right extension of its prefix context
__device__ __forceinline__ int rcx_probe_kmer(const RcxTable& t, unsigned long long code) {
    RcxLookup L;
    rcx_lookup_begin(t, code >> 2, L);
    const bool found = rcx_lookup_end(t, L);
    return rcx_count1(t, code >> 2, false, (int)(code & 3), L, found);
}

// ----------------------------------------------------------------------------- bulk tile copy (TMA, 1-D) + mbarrier
// One thread hands a contiguous, 16-byte aligned range of global memory to the copy engine; the bytes land in shared
// memory without passing through registers and complete a transaction barrier the whole block waits on.
__device__ __forceinline__ unsigned int rcx_smem_addr(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rcx_mbar_init(void* bar, unsigned int arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rcx_smem_addr(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// (shared memory the block wrote with ordinary stores is about to be written by the async proxy)
__device__ __forceinline__ void rcx_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void rcx_bulk_load(void* smem_dst, const void* gsrc, unsigned int bytes, void* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rcx_smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rcx_smem_addr(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(rcx_smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void rcx_mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rcx_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void rcx_mbar_wait(void* bar, unsigned int phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RCX_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RCX_MBAR_DONE;\n"
        "bra RCX_MBAR_WAIT;\n"
        "RCX_MBAR_DONE:\n"
        "}\n" ::"r"(rcx_smem_addr(bar)),
        "r"(phase)
        : "memory");
}

// ----------------------------------------------------------------------------- KmerCode state machine
struct RcxKmer {
    unsigned long long code;
    int inv;  // invalidPos, -1 = valid
};

// v = 0..3 for ACGT, bad = true for anything else (coded as 3 like nucToNum[..] & 3 of -1)
__device__ __forceinline__ void rcx_k_append(RcxKmer& x, int v, bool bad, int k, unsigned long long mask) {
    if (x.inv != -1) ++x.inv;
    x.code = ((x.code << 2) & mask) | (unsigned long long)(bad ? 3 : v);
    if (bad) x.inv = 0;
    if (x.inv >= k) x.inv = -1;
}
__device__ __forceinline__ void rcx_k_prepend(RcxKmer& x, int v, bool bad, int k, unsigned long long mask) {
    if (x.inv != -1) x.inv -= 1;
    x.code = (x.code >> 2) & (mask >> 2);
    if (x.inv < 0) x.inv = -1;
    if (bad) x.inv = k - 1;
    x.code = (x.code | ((unsigned long long)(bad ? 3 : v) << (2 * (k - 1)))) & mask;
}
__device__ __forceinline__ int rcx_k_count(const RcxTable& t, const RcxKmer& x) {
    if (x.inv != -1) return 0;
    return rcx_probe_kmer(t, x.code);
}

// ----------------------------------------------------------------------------- GetBound
// c*ER + 6*sqrt(c*ER) + 1 with every operation rounded separately (the x86-64
// reference build has no FMA contraction).  (int)NaN on x86 is INT_MIN.
__device__ __forceinline__ double rcx_bound_d(int c, double er) {
    double ce = __dmul_rn((double)c, er);
    double s = __dmul_rn(6.0, __dsqrt_rn(ce));
    return __dadd_rn(__dadd_rn(ce, s), 1.0);
}
__device__ __forceinline__ int rcx_bound_i(int c, double er) {
    if (c < 0) return (int)0x80000000;
    return __double2int_rz(rcx_bound_d(c, er));
}

// The search kernel's "force the next round" test (ErrorCorrection.cpp:1195) compares an int with GetBound(strong) in FP64:
// (double)x < B  <=>  x < ceil(B) for an integer x, so the preparation computes ceil(B) once per read (INT_MIN when B is
// NaN, i.e. never) and the search compares integers.
__device__ __forceinline__ int rcx_force_bound(int strong, double er) {
    const double B = rcx_bound_d(strong, er);
    if (!(B == B)) return (int)0x80000000;
    if (B >= 2147483647.0) return 0x7fffffff;
    int b = __double2int_rz(B);
    if ((double)b < B) ++b;
    return b;
}

// out-of-line copies for the (large) correction kernel: its hot loops must stay inside the instruction cache
static __device__ __noinline__ int rcx_bound_i_ool(int c, double er) { return rcx_bound_i(c, er); }
static __device__ __noinline__ int rcx_probe_kmer_ool(const RcxTable& t, unsigned long long code) { return rcx_probe_kmer(t, code); }

// ----------------------------------------------------------------------------- per-read bit planes
// A warp turns one read into bit planes in shared memory:
//   pk[w]   : 2-bit codes, 16 bases per 32-bit word, base 16w in the top bits
//   inv[w]  : bit i of word w = base 32w+i is not A/C/G/T
//   mA/mT/mN: same layout, base == 'A' / 'T' / 'N'
// All planes are padded with two zero words so any k-window can be read with
// funnel shifts without bounds checks.
struct RcxPlanes {
    unsigned int* pk;
    unsigned int* inv;
    unsigned int* mA;
    unsigned int* mT;
    unsigned int* mN;
};
__host__ __device__ __forceinline__ int rcx_pk_words(int max_len) { return (max_len + 15) / 16 + 3; }
__host__ __device__ __forceinline__ int rcx_mask_words(int max_len) { return (max_len + 31) / 32 + 2; }

// returns via out-params the number of N / A / T in the read
__device__ __forceinline__ void rcx_build_planes(const RcxPlanes& P, const char* __restrict__ seq, int len, int max_len,
                                                 int lane, int& nN, int& nA, int& nT) {
    nN = nA = nT = 0;
    const int chunks = (len + 31) >> 5;
    const int mw = rcx_mask_words(max_len), pw = rcx_pk_words(max_len);
    for (int c = 0; c < chunks; ++c) {
        const int p = (c << 5) + lane;
        const char ch = p < len ? seq[p] : 'A';
        int v;
        bool bad = false;
        switch (ch) {
            case 'A': v = 0; break;
            case 'C': v = 1; break;
            case 'G': v = 2; break;
            case 'T': v = 3; break;
            default: v = 3; bad = true; break;
        }
        const bool in = p < len;
        const unsigned int bI = __ballot_sync(RCX_FULL, in && bad);
        const unsigned int bA = __ballot_sync(RCX_FULL, in && ch == 'A');
        const unsigned int bT = __ballot_sync(RCX_FULL, in && ch == 'T');
        const unsigned int bN = __ballot_sync(RCX_FULL, in && ch == 'N');
        const unsigned int sh = (unsigned int)(in ? v : 0) << (30 - 2 * (lane & 15));
        const unsigned int hi = __reduce_or_sync(RCX_FULL, lane < 16 ? sh : 0u);
        const unsigned int lo = __reduce_or_sync(RCX_FULL, lane >= 16 ? sh : 0u);
        if (lane == 0) {
            P.inv[c] = bI; P.mA[c] = bA; P.mT[c] = bT; P.mN[c] = bN;
            P.pk[2 * c] = hi; P.pk[2 * c + 1] = lo;
        }
        nN += __popc(bN); nA += __popc(bA); nT += __popc(bT);
    }
    // zero padding
    for (int w = chunks + lane; w < mw; w += 32) { P.inv[w] = 0; P.mA[w] = 0; P.mT[w] = 0; P.mN[w] = 0; }
    for (int w = 2 * chunks + lane; w < pw; w += 32) P.pk[w] = 0;
    __syncwarp();
}

// bits [j, j+32) of a mask plane
__device__ __forceinline__ unsigned int rcx_bits(const unsigned int* m, int j) {
    return __funnelshift_r(m[j >> 5], m[(j >> 5) + 1], j & 31);
}
// 2-bit code of base p
__device__ __forceinline__ int rcx_base(const unsigned int* pk, int p) { return (pk[p >> 4] >> (30 - 2 * (p & 15))) & 3; }
// code of the k-window starting at base j >= 0 (first base in the highest bits)
__device__ __forceinline__ unsigned long long rcx_window(const unsigned int* pk, int j, int k) {
    const int w = j >> 4, s = (j & 15) * 2;
    const unsigned int x0 = pk[w], x1 = pk[w + 1], x2 = pk[w + 2];
    const unsigned int hi = __funnelshift_l(x1, x0, s), lo = __funnelshift_l(x2, x1, s);
    return (((unsigned long long)hi << 32) | lo) >> (64 - 2 * k);
}
// IsPolyA(seq + j, k, thr): #A >= k - thr or #T >= k - thr
__device__ __forceinline__ bool rcx_polya(const RcxPlanes& P, int j, int k, unsigned int kmask32, int thr) {
    return __popc(rcx_bits(P.mA, j) & kmask32) >= k - thr || __popc(rcx_bits(P.mT, j) & kmask32) >= k - thr;
}

// ----------------------------------------------------------------------------- warp bitonic sort (ascending, shared memory)
// n2 is a power of two >= 64; entries beyond the payload must be padded with INT_MAX by the caller.
__device__ __forceinline__ void rcx_warp_sort(int* s, int n2, int lane) {
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = lane; i < (n2 >> 1); i += 32) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const int a = s[lo], b = s[hi];
                if ((a > b) == up) { s[lo] = b; s[hi] = a; }
            }
            __syncwarp();
        }
    }
}
// a bit mask over the positions / k-mers of one read in NW 64-bit registers (all loops unrolled: no dynamic indexing)
template <int NW>
struct RcxMask {
    unsigned long long m[NW];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int w = 0; w < NW; ++w) m[w] = 0;
    }
    __device__ __forceinline__ bool nonzero() const {
        unsigned long long o = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) o |= m[w];
        return o != 0;
    }
    // bits [lo, hi] (inclusive) of word w as a mask, 0 if the range misses the word
    static __device__ __forceinline__ unsigned long long range_word(int w, int lo, int hi) {
        int l = lo - 64 * w, h = hi - 64 * w;
        if (l < 0) l = 0;
        if (h > 63) h = 63;
        if (l > h) return 0;
        return ((~0ull) >> (63 - (h - l))) << l;
    }
    __device__ __forceinline__ void set(int lo, int hi) {
#pragma unroll
        for (int w = 0; w < NW; ++w) m[w] |= range_word(w, lo, hi);
    }
    __device__ __forceinline__ bool any(int lo, int hi) const {
        unsigned long long o = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) o |= m[w] & range_word(w, lo, hi);
        return o != 0;
    }
    // first index in [from, limit) whose bit equals `want`, or `limit`
    __device__ __forceinline__ int next(int from, int limit, bool want) const {
        int best = limit;
        if (from < limit) {
#pragma unroll
            for (int w = NW - 1; w >= 0; --w) {
                unsigned long long x = (want ? m[w] : ~m[w]) & range_word(w, from, limit - 1);
                if (x) best = 64 * w + __ffsll((long long)x) - 1;
            }
        }
        return best;
    }
    // shift towards higher indices by s (0 <= s < 64)
    __device__ __forceinline__ void shl(int s) {
        if (s == 0) return;
#pragma unroll
        for (int w = NW - 1; w >= 0; --w) m[w] = (m[w] << s) | (w > 0 ? m[w - 1] >> (64 - s) : 0ull);
    }
};

__host__ __device__ __forceinline__ int rcx_pow2_at_least(int n) {
    int p = 64;
    while (p < n) p <<= 1;
    return p;
}
