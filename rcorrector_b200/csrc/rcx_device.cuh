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

// ----------------------------------------------------------------------------- table
// Open-addressed, bucketed.  One bucket = 32 bytes = ONE DRAM sector:
//     u64 key[2] | u32 count[2] | u64 meta            (16 + 8 + 8 bytes)
// key = canonical code, empty = ~0 (never canonical for k <= 32: revcomp of T..T is 0); slots fill in order.
// meta bit 0 = "overflow": some key whose HOME bucket this is lives in a later bucket.  A probe therefore ends at
// its home bucket unless that bit is set (0.45 keys per bucket on average: Poisson(0.45) >= 3 keys, 1.1 % of
// buckets) -- hits AND misses cost exactly one dependent memory round trip of one sector.
// Why this shape (round-1 measurements, profiles/): random reads on this part are limited to ~40 G sectors/s
// whether a probe takes 32 or 64 bytes (rcx_measure_random_read: 1.28 TB/s in both cases), so sectors per probe is
// what counts.  Tried and measured: (a) 16-byte slots, 2 per sector, linear probing at load 0.37: 30 % of miss
// probes needed a second DEPENDENT load; (b) one 64-byte bucket of 5 slots + overflow bit: no dependency, but 2
// sectors per probe; (c) three 10-byte slots with 16-bit saturated counts + a side table for big counts: RNA-seq
// reads concentrate on highly expressed k-mers, so the side lookup was hot.  (d) = this layout: exact 32-bit
// counts inline, one sector, 71 B of HBM per k-mer (1 G k-mers = 71 GB, fits one B200).
struct __align__(32) RcxBucket {
    unsigned long long key[2];
    unsigned int count[2];
    unsigned long long meta;
};
#define RCX_SLOTS 2
struct RcxTable {
    const RcxBucket* buckets;
    unsigned long long nbuckets;
    int k;
};
// a bucket held in registers (one 256-bit load)
struct RcxBucketRegs {
    unsigned long long a0, a1;  // key[0..1]
    unsigned long long w;       // count[0] | count[1] << 32
    unsigned long long meta;
};

__device__ __forceinline__ unsigned long long rcx_hash(unsigned long long x) {
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}
__device__ __forceinline__ unsigned long long rcx_home(const RcxTable& t, unsigned long long canon) {
    return __umul64hi(rcx_hash(canon), t.nbuckets);  // uniform over [0, nbuckets) without a power-of-two table
}

// reverse complement of a k-mer code (A0 C1 G2 T3 => complement is 3-x = ~x on 2 bits)
__device__ __forceinline__ unsigned long long rcx_revcomp(unsigned long long code, int k) {
    unsigned long long r = __brevll(code);                                                  // reverses bits, also inside pairs
    r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);            // restore pair order
    return (~r) >> (64 - 2 * k);
}
__device__ __forceinline__ unsigned long long rcx_canonical(unsigned long long code, int k) {
    unsigned long long rc = rcx_revcomp(code, k);
    return rc < code ? rc : code;
}

// issue the 256-bit load of one bucket (LDG.E.256 on sm_100a; read-only path, no L1 allocation)
__device__ __forceinline__ void rcx_issue(const RcxTable& t, unsigned long long b, RcxBucketRegs& d) {
    const RcxBucket* p = t.buckets + b;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(d.a0), "=l"(d.a1), "=l"(d.w), "=l"(d.meta)
                 : "l"(p));
}
// consume a loaded HOME bucket: Store::GetCount for a canonical key (follows the overflow chain if needed)
__device__ __forceinline__ int rcx_resolve(const RcxTable& t, unsigned long long canon, unsigned long long home, RcxBucketRegs& d) {
    unsigned long long b = home;
    bool at_home = true;
    for (;;) {
        if (d.a0 == canon) return (int)(unsigned int)d.w;
        if (d.a1 == canon) return (int)(unsigned int)(d.w >> 32);
        const bool more = at_home ? (d.meta & 1ull) != 0 : d.a1 != RCX_EMPTY_KEY;
        if (!more) return 0;
        at_home = false;
        b = b + 1 == t.nbuckets ? 0 : b + 1;
        rcx_issue(t, b, d);
    }
}
// Store::GetCount for a canonical key.
__device__ __forceinline__ int rcx_probe(const RcxTable& t, unsigned long long canon) {
    const unsigned long long home = rcx_home(t, canon);
    RcxBucketRegs d;
    rcx_issue(t, home, d);
    return rcx_resolve(t, canon, home, d);
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
    return rcx_probe(t, rcx_canonical(x.code, t.k));
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

// out-of-line copies for the (large) correction kernel: its hot loops must stay inside the instruction cache
static __device__ __noinline__ int rcx_bound_i_ool(int c, double er) { return rcx_bound_i(c, er); }
static __device__ __noinline__ double rcx_bound_d_ool(int c, double er) { return rcx_bound_d(c, er); }
static __device__ __noinline__ int rcx_probe_ool(const RcxTable& t, unsigned long long canon) { return rcx_probe(t, canon); }

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
// Warp bitonic sort in REGISTERS (ascending): element e = r * 32 + lane lives in v[r] of `lane`.  Strides below 32
// are one __shfl_xor per register, strides >= 32 are compare-exchanges between a lane's own registers; no shared
// memory, no barriers (25 shuffle stages + 3 register stages for 128 keys, ~3x fewer instructions than the
// shared-memory network above).
template <int R>
__device__ __forceinline__ void rcx_regsort(int (&v)[R], int lane) {
#pragma unroll
    for (int size = 2; size <= 32 * R; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int rs = stride >> 5;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & rs) == 0) {
                        const bool up = (((r << 5) & size) == 0);
                        const int a = v[r], b = v[r | rs];
                        const int lo = a < b ? a : b, hi = a < b ? b : a;
                        v[r] = up ? lo : hi;
                        v[r | rs] = up ? hi : lo;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int e = (r << 5) | lane;
                    const int p = __shfl_xor_sync(RCX_FULL, v[r], stride);
                    const bool keep_min = ((e & size) == 0) == ((e & stride) == 0);
                    v[r] = keep_min ? (v[r] < p ? v[r] : p) : (v[r] > p ? v[r] : p);
                }
            }
        }
    }
}
// element `idx` of a register-sorted array (warp-uniform idx)
template <int R>
__device__ __forceinline__ int rcx_regsort_get(const int (&v)[R], int idx) {
    int x = v[0];
#pragma unroll
    for (int r = 1; r < R; ++r)
        if ((idx >> 5) == r) x = v[r];
    return __shfl_sync(RCX_FULL, x, idx & 31);
}

__host__ __device__ __forceinline__ int rcx_pow2_at_least(int n) {
    int p = 64;
    while (p < n) p <<= 1;
    return p;
}
