// rcx_count.cu — SURVEY §8f N3: stages 0-2 of run_rcorrector.pl (`jellyfish bc`, `jellyfish count -C --bc`, `jellyfish dump
// -L 2`, run_rcorrector.pl:262-281) on the GPU: exact counts of the canonical k-mers of the reads, the k-mers seen at
// least twice, handed to the context table without the 27-byte-per-k-mer text round trip.
//
// Jellyfish is a third-party dependency that the reference fetches at build time (Makefile:10-21, v2.3.0) and that is
// absent here; what the reference's call sites ask of it is restated: `-m k -C` = canonical k-mers (the smaller of a
// k-mer and its reverse complement), windows that hold a non-ACGT letter are skipped, the bloom counter only serves to
// keep singletons out of the hash (`--bc`), and `dump -L 2` lists the k-mers with a count of 2 or more.  The result is
// "every canonical k-mer that occurs at least twice, with its exact count" -- what tools/rcx_kmercount computes on the CPU
// (the checker of tests/test_gpu_count.py).
//
// Counting table: open addressing with linear probing over two arrays (u64 canonical code, u32 count), empty = ~0 (never
// canonical for k <= 32).  One thread walks one read with rolling forward / reverse-complement codes and claims or finds
// the slot of every valid window with a 64-bit atomicCAS, then adds 1.  The host keeps the table at most half full
// (it grows, by re-insertion, before a batch that could fill it further), so probe sequences stay short and an insertion
// cannot fail.
#include <cub/device/device_scan.cuh>

#include "rcx_internal.h"

#define CNT_EMPTY 0xffffffffffffffffull

__device__ __forceinline__ unsigned long long cnt_slot(unsigned long long key, unsigned long long cap) {
    return __umul64hi(rcx_hash(key), cap);
}
// claim or find the slot of `key`; returns the slot, sets fresh when this call created it
__device__ __forceinline__ unsigned long long cnt_claim(unsigned long long* keys, unsigned long long cap, unsigned long long key, bool& fresh) {
    unsigned long long s = cnt_slot(key, cap);
    fresh = false;
    for (;;) {
        unsigned long long cur = keys[s];
        if (cur == key) return s;
        if (cur == CNT_EMPTY) {
            cur = atomicCAS(&keys[s], CNT_EMPTY, key);
            if (cur == CNT_EMPTY) { fresh = true; return s; }
            if (cur == key) return s;
        }
        if (++s >= cap) s = 0;
    }
}

__global__ void __launch_bounds__(256) rcx_count_init_kernel(unsigned long long* keys, unsigned int* counts, unsigned long long cap) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (unsigned long long)gridDim.x * blockDim.x) {
        keys[i] = CNT_EMPTY;
        counts[i] = 0;
    }
}

__global__ void __launch_bounds__(128) rcx_count_reads_kernel(const char* __restrict__ seq, const uint32_t* __restrict__ off, long long n, int k,
                                                              unsigned long long* keys, unsigned int* counts, unsigned long long cap,
                                                              unsigned long long* stat) {
    const unsigned long long mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    unsigned int fresh_total = 0;
    unsigned long long windows = 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const uint32_t o = off[r];
        const int len = (int)(off[r + 1] - o);
        unsigned long long fw = 0, rv = 0;
        int run = 0;
        for (int i = 0; i < len; ++i) {
            const unsigned char ch = (unsigned char)seq[o + i];
            const unsigned int idx = (unsigned int)ch - 'A';
            const bool ok = idx < 26u && ((0x00080045u >> idx) & 1u);  // A, C, G, T
            const unsigned long long c = ((ch >> 1) ^ (ch >> 2)) & 3;  // A0 C1 G2 T3
            fw = ((fw << 2) | c) & mask;
            rv = (rv >> 2) | ((3 - c) << (2 * (k - 1)));
            run = ok ? run + 1 : 0;
            if (run >= k) {
                bool fresh;
                const unsigned long long s = cnt_claim(keys, cap, fw < rv ? fw : rv, fresh);
                const unsigned int before = atomicAdd(&counts[s], 1u);
                if (before == 0xffffffffu) atomicExch(&counts[s], 0xffffffffu);  // saturate instead of wrapping
                fresh_total += fresh;
                ++windows;
            }
        }
    }
    fresh_total = __reduce_add_sync(RCX_FULL, fresh_total);
    for (int o2 = 16; o2 > 0; o2 >>= 1) windows += __shfl_xor_sync(RCX_FULL, windows, o2);
    if ((threadIdx.x & 31) == 0) {
        if (fresh_total) atomicAdd(&stat[0], (unsigned long long)fresh_total);  // occupied slots
        if (windows) atomicAdd(&stat[1], windows);                              // k-mer instances counted
    }
}

// re-insertion into a larger table
__global__ void __launch_bounds__(256) rcx_count_rehash_kernel(const unsigned long long* __restrict__ okeys, const unsigned int* __restrict__ ocounts,
                                                               unsigned long long ocap, unsigned long long* keys, unsigned int* counts,
                                                               unsigned long long cap) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < ocap; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long key = okeys[i];
        if (key == CNT_EMPTY) continue;
        bool fresh;
        const unsigned long long s = cnt_claim(keys, cap, key, fresh);
        counts[s] = ocounts[i];  // every key is stored once in the old table: no other thread writes this slot
    }
}

__global__ void __launch_bounds__(256) rcx_count_flag_kernel(const unsigned int* __restrict__ counts, unsigned long long lo, unsigned long long n,
                                                             unsigned int min_count, unsigned int* flags) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = counts[lo + i] >= min_count ? 1u : 0u;
}
__global__ void __launch_bounds__(256) rcx_count_scatter_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ counts,
                                                                unsigned long long lo, unsigned long long n, unsigned int min_count,
                                                                const unsigned int* __restrict__ pos, unsigned long long base,
                                                                unsigned long long* out_codes, int32_t* out_counts) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int c = counts[lo + i];
    if (c >= min_count) {
        out_codes[base + pos[i]] = keys[lo + i];
        out_counts[base + pos[i]] = c > 0x7fffffffu ? 0x7fffffff : (int32_t)c;  // Store keeps an int (Store.h:17)
    }
}

cudaError_t rcx_launch_count_init(unsigned long long* keys, unsigned int* counts, unsigned long long cap, cudaStream_t s) {
    rcx_count_init_kernel<<<148 * 8, 256, 0, s>>>(keys, counts, cap);
    return cudaGetLastError();
}
cudaError_t rcx_launch_count_reads(const char* seq, const uint32_t* off, long long n, int k, unsigned long long* keys, unsigned int* counts,
                                   unsigned long long cap, unsigned long long* stat, int sm_count, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    long long grid = (n + 127) / 128;
    if (grid > (long long)sm_count * 16) grid = (long long)sm_count * 16;
    rcx_count_reads_kernel<<<(unsigned int)grid, 128, 0, s>>>(seq, off, n, k, keys, counts, cap, stat);
    return cudaGetLastError();
}
cudaError_t rcx_launch_count_rehash(const unsigned long long* okeys, const unsigned int* ocounts, unsigned long long ocap, unsigned long long* keys,
                                    unsigned int* counts, unsigned long long cap, cudaStream_t s) {
    rcx_count_rehash_kernel<<<148 * 8, 256, 0, s>>>(okeys, ocounts, ocap, keys, counts, cap);
    return cudaGetLastError();
}
size_t rcx_count_select_temp_bytes(unsigned long long n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const unsigned int*)nullptr, (unsigned int*)nullptr, (int)n);
    return bytes + 256;
}
// slots [lo, lo + n) with count >= min_count -> (out_codes, out_counts)[base ...] in slot order; *kept (device) = how many
cudaError_t rcx_launch_count_select(const unsigned long long* keys, const unsigned int* counts, unsigned long long lo, unsigned long long n,
                                    unsigned int min_count, unsigned int* flags, unsigned int* pos, void* temp, size_t temp_bytes,
                                    unsigned long long base, unsigned long long* out_codes, int32_t* out_counts, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const unsigned int grid = (unsigned int)((n + 255) / 256);
    rcx_count_flag_kernel<<<grid, 256, 0, s>>>(counts, lo, n, min_count, flags);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, flags, pos, (int)n, s);
    if (e != cudaSuccess) return e;
    if (out_codes) rcx_count_scatter_kernel<<<grid, 256, 0, s>>>(keys, counts, lo, n, min_count, pos, base, out_codes, out_counts);
    return cudaGetLastError();
}
