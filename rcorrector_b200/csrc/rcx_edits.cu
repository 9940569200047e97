// rcx_edits.cu — compact host interface (rcx_correct_batch_compact): instead of copying every corrected read back to the
// host, the applied substitutions of the changed reads are gathered in record order into one stream of 16-bit entries
// (position | base << 10).  Record r owns correction[r] entries (when positive), after those of the records before it:
// an exclusive prefix sum over max(correction, 0) gives every read its place, one thread per read copies its edits there
// from what the finalize kernels left in fixbuf (the search kernel's edit list filtered by the post-filters, or the
// dense fix[] of the warp-per-read path).
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include "rcx_internal.h"

struct RcxPositive {
    __host__ __device__ __forceinline__ uint32_t operator()(const int32_t& c) const { return c > 0 ? (uint32_t)c : 0u; }
};
typedef thrust::transform_iterator<RcxPositive, const int32_t*, uint32_t> RcxPositiveIt;

size_t rcx_edits_temp_bytes(long long n) {
    size_t bytes = 0;
    RcxPositiveIt it((const int32_t*)nullptr, RcxPositive());
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, (uint32_t*)nullptr, (int)n);
    return bytes + 256;
}

__global__ void __launch_bounds__(256) rcx_edits_gather_kernel(const int32_t* __restrict__ corr, const uint32_t* __restrict__ off, long long n,
                                                               const uint32_t* __restrict__ scan, const signed char* __restrict__ fixbuf,
                                                               const unsigned short* __restrict__ fixlist, uint16_t* __restrict__ edits,
                                                               unsigned long long* counters) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int c = corr[r];
    if (r == n - 1) counters[RCX_C_NEDITS] = (unsigned long long)scan[r] + (unsigned long long)(c > 0 ? c : 0);
    if (c <= 0) return;
    uint16_t* out = edits + scan[r];
    const uint32_t o = off[r];
    if (fixlist) {  // count, then pos | base << 8
        const unsigned short* fl = fixlist + o;
        for (int e = 0; e < c; ++e) { const unsigned int x = fl[1 + e]; out[e] = (uint16_t)((x & 255u) | ((x >> 8) << 10)); }
    } else {        // dense fix[]: -1 = unchanged
        const signed char* fx = fixbuf + o;
        const int len = (int)(off[r + 1] - o);
        int e = 0;
        for (int i = 0; i < len && e < c; ++i)
            if (fx[i] != -1) out[e++] = (uint16_t)(i | ((unsigned int)fx[i] << 10));
    }
}

cudaError_t rcx_launch_edits_gather(const int32_t* corr, const uint32_t* off, long long n, const signed char* fixbuf, const unsigned short* fixlist,
                                    uint32_t* scan, void* temp, size_t temp_bytes, uint16_t* edits, unsigned long long* counters, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    RcxPositiveIt it(corr, RcxPositive());
    cudaError_t e = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, it, scan, (int)n, s);
    if (e != cudaSuccess) return e;
    rcx_edits_gather_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, s>>>(corr, off, n, scan, fixbuf, fixlist, edits, counters);
    return cudaGetLastError();
}
