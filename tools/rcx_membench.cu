// rcx_membench.cu — random-access ceiling of one B200 ("the random-access ceiling must be measured on the
// box").  Independent uniformly random reads of 16 / 32 / 64 bytes over a buffer, swept over footprint, load flavour,
// loads in flight per thread, cudaLimitMaxL2FetchGranularity and a TLB-reach pattern.  Test/bench infrastructure, not
// part of librcx.so.  Prints one JSON object per configuration.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/rcx_membench rcx_membench.cu
// ncu  : ncu --metrics dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum
//            bin/rcx_membench quick      -> DRAM sectors fetched per requested sector
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long mix(unsigned long long x) {
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
    return x;
}

enum { V_NC_NOALLOC = 0, V_PLAIN = 1, V_CG = 2, V_NC_EVICT_FIRST = 3, V_LU = 4 };

template <int V>
__device__ __forceinline__ void ld32(const void* p, unsigned long long& a, unsigned long long& b, unsigned long long& c, unsigned long long& d) {
    if (V == V_NC_NOALLOC)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    else if (V == V_PLAIN)
        asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    else if (V == V_CG) {  // two 128-bit .cg loads (cache at L2 only)
        asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
        asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2+16];" : "=l"(c), "=l"(d) : "l"(p));
    } else if (V == V_NC_EVICT_FIRST)
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    else {
        asm volatile("ld.global.lu.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
        asm volatile("ld.global.lu.v2.u64 {%0,%1}, [%2+16];" : "=l"(c), "=l"(d) : "l"(p));
    }
}

// BYTES: 16 (one 128-bit load), 32, 64 (two adjacent 32-byte loads).  MLP loads in flight per thread.
// pattern 0: uniform over [0, nunits) units of 64 bytes.
// pattern 1: TLB probe - only `pages` 2 MiB pages are touched (page list = every `page_stride`-th page), `span` bytes in each.
template <int V, int BYTES, int MLP>
__global__ void __launch_bounds__(256) rr_kernel(const char* base, unsigned long long nunits, int pattern, unsigned long long pages,
                                                 unsigned long long page_stride, unsigned long long span_units, unsigned long long iters,
                                                 unsigned long long seed, unsigned long long* sink) {
    unsigned long long x = seed + ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
    unsigned long long acc = 0;
    for (unsigned long long it = 0; it < iters; ++it) {
        const char* p[MLP];
#pragma unroll
        for (int u = 0; u < MLP; ++u) {
            x = mix(x + u + 1);
            unsigned long long unit;
            if (pattern == 0) unit = __umul64hi(x, nunits);
            else {
                const unsigned long long pg = __umul64hi(x, pages) * page_stride;
                const unsigned long long in = __umul64hi(x * 0x9E3779B97F4A7C15ull, span_units);
                unit = pg * (2097152ull / 64) + in;
            }
            p[u] = base + unit * 64;
        }
        unsigned long long v[MLP][BYTES / 8];
#pragma unroll
        for (int u = 0; u < MLP; ++u) {
            if (BYTES == 16) asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(v[u][0]), "=l"(v[u][1]) : "l"(p[u]));
            else {
                ld32<V>(p[u], v[u][0], v[u][1], v[u][2], v[u][3]);
                if (BYTES == 64) ld32<V>(p[u] + 32, v[u][BYTES / 8 - 4], v[u][BYTES / 8 - 3], v[u][BYTES / 8 - 2], v[u][BYTES / 8 - 1]);
            }
        }
#pragma unroll
        for (int u = 0; u < MLP; ++u)
#pragma unroll
            for (int w = 0; w < BYTES / 8; ++w) acc ^= v[u][w];
    }
    if (acc == 0x1234567887654321ull) *sink = acc;
}

struct Cfg { int v, bytes, mlp, pattern; unsigned long long pages, page_stride, span; const char* name; };

template <int V, int BYTES, int MLP>
static float run(const char* buf, unsigned long long nunits, const Cfg& c, int grid, unsigned long long iters, unsigned long long* sink) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        rr_kernel<V, BYTES, MLP><<<grid, 256>>>(buf, nunits, c.pattern, c.pages, c.page_stride, c.span / 64, iters, 1234 + rep, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

static float dispatch(const char* buf, unsigned long long nunits, const Cfg& c, int grid, unsigned long long iters, unsigned long long* sink) {
#define D(V, B, M) if (c.v == V && c.bytes == B && c.mlp == M) return run<V, B, M>(buf, nunits, c, grid, iters, sink);
    D(0, 32, 8) D(0, 64, 8) D(0, 16, 8) D(0, 32, 4) D(0, 32, 16) D(0, 32, 2)
    D(1, 32, 8) D(2, 32, 8) D(3, 32, 8) D(4, 32, 8) D(1, 64, 8)
#undef D
    fprintf(stderr, "no instantiation for v=%d bytes=%d mlp=%d\n", c.v, c.bytes, c.mlp);
    exit(1);
}

int main(int argc, char** argv) {
    const bool quick = argc > 1 && !strcmp(argv[1], "quick");
    size_t gran_req = 0;
    for (int i = 1; i < argc; ++i) if (!strcmp(argv[i], "-gran") && i + 1 < argc) gran_req = (size_t)atoi(argv[i + 1]);
    CK(cudaSetDevice(0));
    if (gran_req) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran_req); if (e != cudaSuccess) { fprintf(stderr, "set gran: %s\n", cudaGetErrorString(e)); cudaGetLastError(); } }
    size_t gran = 0;
    CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int grid = prop.multiProcessorCount * 8;
    unsigned long long* sink; CK(cudaMalloc(&sink, 8));
    const size_t maxbytes = quick ? (4ull << 30) : (16ull << 30);
    char* buf; CK(cudaMalloc(&buf, maxbytes));
    CK(cudaMemset(buf, 1, maxbytes));
    const char* vn[] = {"nc.L1::no_allocate", "plain", "cg", "nc.L2::evict_first", "lu"};
    // footprint sweep with the product's load flavour
    const size_t sizes_full[] = {64ull << 20, 256ull << 20, 1ull << 30, 4ull << 30, 16ull << 30};
    const size_t sizes_quick[] = {4ull << 30};
    const size_t* sizes = quick ? sizes_quick : sizes_full;
    const int nsizes = quick ? 1 : 5;
    const unsigned long long iters = 128;
    for (int s = 0; s < nsizes; ++s) {
        const unsigned long long nunits = sizes[s] / 64;
        Cfg cfgs[] = {
            {0, 32, 8, 0, 0, 0, 0, "uniform"}, {0, 64, 8, 0, 0, 0, 0, "uniform"}, {0, 16, 8, 0, 0, 0, 0, "uniform"},
            {0, 32, 2, 0, 0, 0, 0, "uniform"}, {0, 32, 4, 0, 0, 0, 0, "uniform"}, {0, 32, 16, 0, 0, 0, 0, "uniform"},
            {1, 32, 8, 0, 0, 0, 0, "uniform"}, {1, 64, 8, 0, 0, 0, 0, "uniform"}, {2, 32, 8, 0, 0, 0, 0, "uniform"},
            {3, 32, 8, 0, 0, 0, 0, "uniform"}, {4, 32, 8, 0, 0, 0, 0, "uniform"},
        };
        for (const Cfg& c : cfgs) {
            if (!quick && sizes[s] != (4ull << 30) && !(c.v == 0 && c.mlp == 8)) continue;  // flavours only at 4 GiB
            const float ms = dispatch(buf, nunits, c, grid, iters, sink);
            const double req = (double)grid * 256 * iters * c.mlp;
            printf("{\"pattern\": \"%s\", \"footprint_mib\": %zu, \"load\": \"%s\", \"bytes\": %d, \"mlp\": %d, \"l2_fetch_gran\": %zu, \"ms\": %.3f, "
                   "\"greq_per_s\": %.2f, \"gbs\": %.1f}\n", c.name, sizes[s] >> 20, vn[c.v], c.bytes, c.mlp, gran, ms, req / ms / 1e6,
                   req * c.bytes / ms / 1e6);
            fflush(stdout);
        }
    }
    // TLB probe: the same 128 MiB of data (L2-resident either way) in 64 whole pages vs 64 KiB slices of 2048 pages,
    // and 1 GiB of data (DRAM-resident) in 512 whole pages vs 128 KiB slices of 8192 pages (needs the 16 GiB buffer)
    {
        Cfg t[] = {
            {0, 32, 8, 1, 64, 1, 2097152, "tlb:128MiB in 64 pages"},
            {0, 32, 8, 1, 2048, 1, 65536, "tlb:128MiB in 2048 pages"},
            {0, 32, 8, 1, 512, 1, 2097152, "tlb:1GiB in 512 pages"},
            {0, 32, 8, 1, 2040, 1, 524288, "tlb:1GiB in 2040 pages"},
        };
        for (const Cfg& c : t) {
            if ((c.pages * c.page_stride) * 2097152ull > maxbytes) continue;
            const float ms = dispatch(buf, maxbytes / 64, c, grid, iters, sink);
            const double req = (double)grid * 256 * iters * c.mlp;
            printf("{\"pattern\": \"%s\", \"load\": \"%s\", \"bytes\": %d, \"mlp\": %d, \"l2_fetch_gran\": %zu, \"ms\": %.3f, \"greq_per_s\": %.2f, \"gbs\": %.1f}\n",
                   c.name, vn[c.v], c.bytes, c.mlp, gran, ms, req / ms / 1e6, req * c.bytes / ms / 1e6);
            fflush(stdout);
        }
    }
    return 0;
}
