// rcx_table.cu — build and bulk-lookup kernels of the k-mer count table.
//
// Replaces Store::Put / Store::GetCount (Store.h:51-66) and the load loop of
// main.cpp:295-307: canonicalise, skip count <= 1, last record of a duplicate
// canonical key wins (resolved with a 64-bit atomicMax on {tag:count}, tag =
// record index + 1, so the result is independent of thread scheduling).
#include "rcx_internal.h"

__global__ void __launch_bounds__(256) rcx_table_init_kernel(uint4* p, unsigned long long nslots) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += stride)
        p[i] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);  // key = EMPTY, count = 0, tag = 0
}

__global__ void __launch_bounds__(256)
rcx_table_insert_kernel(RcxBucket* buckets, unsigned long long bucket_mask, int k, const unsigned long long* __restrict__ codes,
                        const int32_t* __restrict__ counts, unsigned long long n, unsigned long long first_index,
                        unsigned long long* new_keys) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    unsigned long long fresh = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int c = counts[i];
        if (c <= 1) continue;  // main.cpp:299
        const unsigned long long key = rcx_canonical(codes[i] & kmask, k);
        unsigned long long b = rcx_hash(key) & bucket_mask;
        RcxSlot* slot = nullptr;
        while (!slot) {
            for (int s = 0; s < 2 && !slot; ++s) {
                RcxSlot* cand = &buckets[b].s[s];
                unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&cand->key);
                if (cur == RCX_EMPTY_KEY) {
                    cur = atomicCAS(&cand->key, RCX_EMPTY_KEY, key);
                    if (cur == RCX_EMPTY_KEY) { ++fresh; cur = key; }
                }
                if (cur == key) slot = cand;
            }
            b = (b + 1) & bucket_mask;
        }
        const unsigned long long packed = ((first_index + i + 1) << 32) | (unsigned int)c;
        atomicMax(reinterpret_cast<unsigned long long*>(&slot->count), packed);
    }
    if (fresh) atomicAdd(new_keys, fresh);
}

__global__ void __launch_bounds__(256)
rcx_table_lookup_kernel(RcxTable t, const unsigned long long* __restrict__ codes, int32_t* __restrict__ out, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long kmask = t.k >= 32 ? ~0ull : ((1ull << (2 * t.k)) - 1);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = rcx_probe(t, rcx_canonical(codes[i] & kmask, t.k));
}

static inline int grid_for(unsigned long long n, int block, int cap) {
    unsigned long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > (unsigned long long)cap) g = cap;
    return (int)g;
}

cudaError_t rcx_launch_table_init(RcxBucket* buckets, unsigned long long nbuckets, cudaStream_t s) {
    const unsigned long long nslots = nbuckets * 2;
    rcx_table_init_kernel<<<grid_for(nslots, 256, 148 * 16), 256, 0, s>>>(reinterpret_cast<uint4*>(buckets), nslots);
    return cudaGetLastError();
}

cudaError_t rcx_launch_table_insert(RcxBucket* buckets, unsigned long long bucket_mask, int k, const unsigned long long* codes,
                                    const int32_t* counts, unsigned long long n, unsigned long long first_index,
                                    unsigned long long* new_keys, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    rcx_table_insert_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, s>>>(buckets, bucket_mask, k, codes, counts, n, first_index,
                                                                      new_keys);
    return cudaGetLastError();
}

cudaError_t rcx_launch_table_lookup(RcxTable t, const unsigned long long* codes, int32_t* out, unsigned long long n,
                                    cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    rcx_table_lookup_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, s>>>(t, codes, out, n);
    return cudaGetLastError();
}
