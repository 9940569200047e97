// rcx_table.cu — build and bulk-lookup kernels of the context-keyed k-mer count table (layout: rcx_device.cuh).
//
// Replaces Store::Put / Store::GetCount (Store.h:51-66) and the load loop of main.cpp:295-307: skip count <= 1,
// canonical strand, last record of a duplicate canonical key wins.  Every record is placed twice: as a right
// extension of its prefix (k-1)-mer and as a left extension of its suffix (k-1)-mer, each normalised to the
// canonical strand of that context.  Passes over the records (all schedule-independent):
//   claim (home only)  every context key tries its home sector;
//   claim (sequence)   keys that lost their home walk the probe sequence, the home sector gets its overflow bit;
//   winner             32-bit atomicMax of (record index + 1) per count slot in a build-time side array
//                      ("last writer wins", Store.h:55);
//   write              the winning record of each slot ORs its 24-bit count into the entry.
#include "rcx_internal.h"

__global__ void __launch_bounds__(256) rcx_table_init_kernel(uint4* p, unsigned long long nsectors) {
    // entry = 2 x 16 bytes: key all ones + first two count words zero, then four count words zero
    const unsigned long long n = nsectors * 2;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        p[i] = (i & 1) ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
}

struct RcxPlacement {
    unsigned long long key;
    int slot0, slot1;  // slot = side * 4 + base (side 0 = RIGHT group, 1 = LEFT group); slot1 >= 0 only for a palindromic context
};
// the two places where the k-mer `code` (either strand) is stored
__device__ __forceinline__ void rcx_placements(unsigned long long code, int k, RcxPlacement (&p)[2]) {
    const int n = k - 1;
    const unsigned long long m1 = (1ull << (2 * n)) - 1;  // n <= 31
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        const unsigned long long ctx = which == 0 ? (code >> 2) : (code & m1);
        int side = which;                                                      // prefix context: right extension
        int base = which == 0 ? (int)(code & 3) : (int)((code >> (2 * n)) & 3);
        const unsigned long long rc = rcx_revcomp(ctx, n);
        p[which].slot1 = -1;
        if (rc < ctx) { p[which].key = rc; side ^= 1; base = 3 - base; }
        else {
            p[which].key = ctx;
            if (rc == ctx) p[which].slot1 = (side ^ 1) * 4 + (3 - base);  // ctx+b and comp(b)+ctx are the same k-mer
        }
        p[which].slot0 = side * 4 + base;
    }
}

__device__ __forceinline__ unsigned long long rcx_home_sector(unsigned long long key, unsigned long long nlines) {
    return __umul64hi(rcx_hash(key), nlines << 2);
}
__device__ __forceinline__ unsigned long long rcx_seq_sector(unsigned long long h, unsigned long long i, unsigned long long nlines) {
    unsigned long long line = ((h >> 2) + (i >> 2)) % nlines;
    return (line << 2) | ((h ^ i) & 3);
}

// pass 1: claim sectors for the context keys.  home_only: only the home sector is tried.
__global__ void __launch_bounds__(256)
rcx_table_claim_kernel(RcxEntry* e, unsigned long long nlines, int k, const unsigned long long* __restrict__ codes,
                       const int32_t* __restrict__ counts, unsigned long long n, int home_only, unsigned long long* stat,
                       unsigned long long fresh_limit) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    unsigned long long fresh = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (counts[i] <= 1) continue;  // main.cpp:299
        RcxPlacement p[2];
        rcx_placements(codes[i] & kmask, k, p);
        for (int w = 0; w < 2; ++w) {
            const unsigned long long key = p[w].key, h = rcx_home_sector(key, nlines);
            for (unsigned long long j = 0;; ++j) {
                if (*reinterpret_cast<volatile unsigned long long*>(&stat[1])) return;  // table too small: host retries
                const unsigned long long s = rcx_seq_sector(h, j, nlines);
                unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&e[s].key);
                if (cur == ~0ull) {
                    cur = atomicCAS(&e[s].key, ~0ull, key);
                    if (cur == ~0ull) { ++fresh; cur = key; }
                }
                if ((cur & RCX_KEY_MASK) == key) {
                    if (j) atomicOr(&e[h].key, RCX_OVERFLOW_BIT);
                    break;
                }
                if (home_only) break;
                if (j >= 4 * nlines) {  // every sector visited and none free: the table is full, whatever `fresh` says so far
                    atomicExch(&stat[1], 1ull);
                    return;
                }
            }
        }
    }
    if (fresh) {
        const unsigned long long tot = atomicAdd(&stat[0], fresh) + fresh;
        if (tot > fresh_limit) atomicExch(&stat[1], 1ull);
    }
}

__device__ __forceinline__ unsigned long long rcx_find_sector(const RcxEntry* e, unsigned long long nlines, unsigned long long key) {
    const unsigned long long h = rcx_home_sector(key, nlines);
    for (unsigned long long j = 0;; ++j) {
        const unsigned long long s = rcx_seq_sector(h, j, nlines);
        if ((e[s].key & RCX_KEY_MASK) == key) return s;
    }
}

// pass 2 (write = 0): atomicMax of record index + 1 per count slot;  pass 3 (write = 1): the winner writes its count.
// Only sectors in [part_lo, part_hi) are handled (the side array covers one partition of the table at a time).
__global__ void __launch_bounds__(256)
rcx_table_fill_kernel(RcxEntry* e, unsigned long long nlines, int k, const unsigned long long* __restrict__ codes,
                      const int32_t* __restrict__ counts, unsigned long long n, unsigned long long first_index, unsigned int* win,
                      unsigned long long part_lo, unsigned long long part_hi, int write, unsigned long long* stat,
                      unsigned long long* big_keys, unsigned int* big_counts, unsigned long long big_cap) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    unsigned long long winners = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int c = counts[i];
        if (c <= 1) continue;
        const unsigned long long code = codes[i] & kmask;
        RcxPlacement p[2];
        rcx_placements(code, k, p);
        const unsigned int tag = (unsigned int)(first_index + i + 1);
        for (int w = 0; w < 2; ++w) {
            const unsigned long long s = rcx_find_sector(e, nlines, p[w].key);
            if (s < part_lo || s >= part_hi) continue;
            for (int q = 0; q < 2; ++q) {
                const int slot = q == 0 ? p[w].slot0 : p[w].slot1;
                if (slot < 0) continue;
                unsigned int* wp = win + (s - part_lo) * 8 + slot;
                if (!write) { atomicMax(wp, tag); continue; }
                if (*wp != tag) continue;
                const unsigned int v = (unsigned int)c >= RCX_BIG ? RCX_BIG : (unsigned int)c;
                unsigned int* cw = e[s].cw + (slot >> 2) * 3;
                const int b = slot & 3;
                if (b < 3) atomicOr(&cw[b], v);
                else {
                    atomicOr(&cw[0], (v & 0xffu) << 24);
                    atomicOr(&cw[1], ((v >> 8) & 0xffu) << 24);
                    atomicOr(&cw[2], ((v >> 16) & 0xffu) << 24);
                }
                if (w == 0 && q == 0) {  // one count per distinct canonical k-mer
                    ++winners;
                    if (v == RCX_BIG) {
                        const unsigned long long at = atomicAdd(&stat[3], 1ull);
                        if (at < big_cap) { big_keys[at] = rcx_canonical(code, k); big_counts[at] = (unsigned int)c; }
                    }
                }
            }
        }
    }
    if (winners) atomicAdd(&stat[2], winners);
}

__global__ void __launch_bounds__(256)
rcx_table_lookup_kernel(RcxTable t, const unsigned long long* __restrict__ codes, int32_t* __restrict__ out, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long kmask = t.k >= 32 ? ~0ull : ((1ull << (2 * t.k)) - 1);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = rcx_probe_kmer(t, codes[i] & kmask);
}

// counts of the four last-base variants of every record (ERROR_RATE pass, main.cpp:324-340): ONE context lookup each
__global__ void __launch_bounds__(256)
rcx_table_lookup4_kernel(RcxTable t, const unsigned long long* __restrict__ codes, int4* __restrict__ out, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long kmask = t.k >= 32 ? ~0ull : ((1ull << (2 * t.k)) - 1);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long ctx = (codes[i] & kmask) >> 2;
        RcxLookup L;
        rcx_lookup_begin(t, ctx, L);
        const bool found = rcx_lookup_end(t, L);
        int4 r;
        rcx_counts4(t, ctx, false, L, found, r.x, r.y, r.z, r.w);
        out[i] = r;
    }
}

// Roofline micro-benchmark: independent, uniformly random 32- or 64-byte reads over the table (8 in flight per thread).
// This is the ceiling the lookup kernel is compared with ("the random-access ceiling must be measured on the box").
template <int BYTES>
__global__ void __launch_bounds__(256) rcx_random_read_kernel(RcxTable t, unsigned long long iters, unsigned long long seed,
                                                              unsigned long long* sink) {
    unsigned long long x = seed + ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
    unsigned long long acc = 0;
    const unsigned long long nsec = t.nlines << 2;
    for (unsigned long long it = 0; it < iters; ++it) {
        unsigned long long b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x = rcx_hash(x + u + 1);
            b[u] = __umul64hi(x, nsec - 1) & ~1ull;  // even sector: a 64-byte read stays inside the table and 64-byte aligned
        }
        unsigned long long v[8][BYTES / 8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const RcxEntry* p = t.e + b[u];
            asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(v[u][0]), "=l"(v[u][1]), "=l"(v[u][2]), "=l"(v[u][3]) : "l"(p));
            if (BYTES == 64)
                asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4+32];"
                             : "=l"(v[u][BYTES / 8 - 4]), "=l"(v[u][BYTES / 8 - 3]), "=l"(v[u][BYTES / 8 - 2]), "=l"(v[u][BYTES / 8 - 1]) : "l"(p));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int w = 0; w < BYTES / 8; ++w) acc ^= v[u][w];
    }
    if (acc == 0x1234567887654321ull) *sink = acc;  // keep the loads alive
}

cudaError_t rcx_launch_random_read(RcxTable t, int bytes, int grid, unsigned long long iters, unsigned long long* sink, cudaStream_t s) {
    if (bytes == 64) rcx_random_read_kernel<64><<<grid, 256, 0, s>>>(t, iters, 12345, sink);
    else rcx_random_read_kernel<32><<<grid, 256, 0, s>>>(t, iters, 12345, sink);
    return cudaGetLastError();
}

static inline int grid_for(unsigned long long n, int block, int cap) {
    unsigned long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > (unsigned long long)cap) g = cap;
    return (int)g;
}

cudaError_t rcx_launch_table_init(RcxEntry* e, unsigned long long nlines, cudaStream_t s) {
    rcx_table_init_kernel<<<grid_for(nlines * 8, 256, 148 * 16), 256, 0, s>>>(reinterpret_cast<uint4*>(e), nlines * 4);
    return cudaGetLastError();
}

cudaError_t rcx_launch_table_claim(RcxEntry* e, unsigned long long nlines, int k, const unsigned long long* codes, const int32_t* counts,
                                   unsigned long long n, int home_only, unsigned long long* stat, unsigned long long fresh_limit,
                                   cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    rcx_table_claim_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, s>>>(e, nlines, k, codes, counts, n, home_only, stat, fresh_limit);
    return cudaGetLastError();
}

cudaError_t rcx_launch_table_fill(RcxEntry* e, unsigned long long nlines, int k, const unsigned long long* codes, const int32_t* counts,
                                  unsigned long long n, unsigned long long first_index, unsigned int* win, unsigned long long part_lo,
                                  unsigned long long part_hi, int write, unsigned long long* stat, unsigned long long* big_keys,
                                  unsigned int* big_counts, unsigned long long big_cap, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    rcx_table_fill_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, s>>>(e, nlines, k, codes, counts, n, first_index, win, part_lo, part_hi,
                                                                    write, stat, big_keys, big_counts, big_cap);
    return cudaGetLastError();
}

cudaError_t rcx_launch_table_lookup(RcxTable t, const unsigned long long* codes, int32_t* out, unsigned long long n,
                                    cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    rcx_table_lookup_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, s>>>(t, codes, out, n);
    return cudaGetLastError();
}

cudaError_t rcx_launch_table_lookup4(RcxTable t, const unsigned long long* codes, int32_t* out4, unsigned long long n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    rcx_table_lookup4_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, s>>>(t, codes, reinterpret_cast<int4*>(out4), n);
    return cudaGetLastError();
}
