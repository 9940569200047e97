// rcx_dfs.cu — the ordered correction search: SearchPaths_Right / SearchPaths_Left
// (ErrorCorrection.cpp:201-442 / 444-678), InferPosThreshold (:144-173) and the per-read
// bookkeeping around them (:1118-1242), one THREAD per read.
//
// Why a thread, not a warp: the search is an order-dependent depth-first walk whose state
// (maxFixCnt, trialCnt, best path, tie counts) is shared by the whole recursion, so the only
// parallelism inside one read is the 4 one-base extensions of a node.  A warp-per-read version of
// this kernel (round 1, profiles/) executed ~4 000 warp-instructions per read with 31 of 32 lanes
// repeating the same scalar bookkeeping, ran at 16 warps/SM because of its register count, and
// spent 40 % of its issue slots waiting for instructions (63 KB of straight-line SASS executed once
// per read).  Here every lane walks its own read: the bookkeeping is SIMT-parallel across 32
// reads, the 4 extension probes of a node are 4 independent 64-byte loads per lane (128 per warp in
// flight), and lanes pull new work items on their own so that short searches do not wait for long
// ones.  Reads are independent, so results do not depend on the mapping.
//
// The recursion becomes an explicit stack of 11-word frames per thread, interleaved by lane in a
// global scratch buffer (same-depth accesses of a warp coalesce; depth <= kcnt + 2 because the jump
// child is a tail call).  Children are visited in the reference's order: keep-base (or the
// "accidentally below threshold" rescue), then A, C, G, T, then the jump through an unfixable
// region.  The four extension k-mers of a node serve InferPosThreshold, the keep-base test and the
// three alternatives at once (the CPU probes the same k-mers 8 times).
#include "rcx_internal.h"

#define DFS_THREADS 128
#define DFS_FRAME 11
#define DFS_LUT 1024

enum { T_FETCH = 0, T_SEG = 1, T_ENTER = 2, T_ALT0 = 3 /* ..T_ALT0+3 */, T_JUMP = 7, T_POP = 8, T_IDLE = 9 };

__device__ __forceinline__ int dfs_sel4(int b, int c0, int c1, int c2, int c3) { return b == 0 ? c0 : b == 1 ? c1 : b == 2 ? c2 : c3; }

// counts of the 4 one-base extensions of kc: they share one (k-1)-mer context, so ONE sector answers the node
// (InferPosThreshold, the keep-base test and the three alternatives)
__device__ __noinline__ void dfs_probe4(const RcxTable& t, const RcxKmer& kc, bool right, unsigned long long mask, int& c0, int& c1,
                                        int& c2, int& c3) {
    // validity of the children after Append / Prepend of a valid base (KmerCode.cpp:7-42)
    const bool valid = kc.inv == -1 || (right ? kc.inv + 1 >= t.k : kc.inv <= 0);
    c0 = c1 = c2 = c3 = 0;
    if (!valid) return;
    const unsigned long long ctx = right ? (kc.code & (mask >> 2)) : (kc.code >> 2);
    RcxLookup L;
    rcx_lookup_begin(t, ctx, L);
    const bool found = rcx_lookup_end(t, L);
    rcx_counts4(t, ctx, !right, L, found, c0, c1, c2, c3);
}

// the same from an entry already in registers: counts of the four extensions ctx+b (left = false) or b+ctx (left = true)
__device__ __forceinline__ void dfs_counts4(const RcxTable& t, unsigned long long ctx, bool left, const RcxEntryRegs& d, bool flip, bool found,
                                            int& c0, int& c1, int& c2, int& c3) {
    c0 = c1 = c2 = c3 = 0;
    if (!found) return;
    unsigned int f0, f1, f2, f3;
    rcx_group(d, left != flip, f0, f1, f2, f3);
    c0 = (int)(flip ? f3 : f0); c1 = (int)(flip ? f2 : f1); c2 = (int)(flip ? f1 : f2); c3 = (int)(flip ? f0 : f3);
    if (((f0 + 1) | (f1 + 1) | (f2 + 1) | (f3 + 1)) & 0x1000000u) {  // some field == RCX_BIG
        if (c0 == (int)RCX_BIG) c0 = rcx_big_count(t, rcx_canonical(rcx_ext_kmer(ctx, left, 0, t.k), t.k));
        if (c1 == (int)RCX_BIG) c1 = rcx_big_count(t, rcx_canonical(rcx_ext_kmer(ctx, left, 1, t.k), t.k));
        if (c2 == (int)RCX_BIG) c2 = rcx_big_count(t, rcx_canonical(rcx_ext_kmer(ctx, left, 2, t.k), t.k));
        if (c3 == (int)RCX_BIG) c3 = rcx_big_count(t, rcx_canonical(rcx_ext_kmer(ctx, left, 3, t.k), t.k));
    }
}

// InferPosThreshold (:144-173) given the four extension counts; lut[c] = max(1, (int)GetBound(c)) for small c
__device__ __forceinline__ int dfs_infer(int c0, int c1, int c2, int c3, int upper, double er, const int* lut) {
    int mx = c0 > c1 ? c0 : c1;
    const int my = c2 > c3 ? c2 : c3;
    mx = mx > my ? mx : my;
    if (mx < 0) mx = 0;
    int r;
    if (mx < DFS_LUT) r = lut[mx];
    else {
        // GetBound is non-decreasing in c (every rounded operation is monotone), so for a count beyond the table the
        // result is already known when the table's last entry reaches the cap: no FP64 sqrt for highly expressed k-mers
        if (upper > 0 && lut[DFS_LUT - 1] >= upper) return upper;
        r = rcx_bound_i(mx, er);
        if (r < 1) r = 1;
    }
    return (upper > r || upper <= 0) ? r : upper;
}

#ifndef DFS_MINLANES
#define DFS_MINLANES 20  // measured 0 / 6 / 12 / 20 / 24 / 28 / 32: 35.1 / 29.9 / 26.8 / 26.0 / 26.7 / 28.6 / 29.6 ms
#endif
#ifndef DFS_MINBLOCKS
#define DFS_MINBLOCKS 6
#endif
#ifndef DFS_AHEAD
// 1: keep-base chains request the child's sector before the node's own counts are consumed.  Measured slower (10 M x 100 bp:
// search 21.6 -> 25.5 ms at 6 blocks / SM, 24.8 at 5; 150-bp pairs 121 -> 143 ms; k = 31 / -maxcorK 8 56 -> 66 ms,
// profiles/r02_c): 11 more live registers and 5 % more requests cost more than the latency they hide -- six warps per
// scheduler already cover it; the kernel is bound by instruction issue under divergence, not by memory latency.
#define DFS_AHEAD 0
#endif
#ifndef DFS_TAIL_PREFETCH
#define DFS_TAIL_PREFETCH 0  // lanes: at or below this many active lanes in the node loop, the keep-base child's sector is prefetched into L2.  Measured: no gain (2 GPUs, rank 1 search 54.96 ms off, 56.12 at 4 lanes, 56.38 at 12, gpurun_out/r02_n2c) -> off
#endif
#ifndef DFS_BATCH
// 1: the LIGHT work items (the scan kernel files them sorted by their number of untrusted k-windows within every tile) are
// taken by a whole warp at a time, 32 consecutive slots, and a lane that finishes waits for the others: items of one
// batch take nearly the same walk, so the lanes go through fetch / segment / alternative / pop together instead of one or
// two at a time.  The heavy items (first) and the second pass keep the per-lane dynamic fetch.  Measured (profiles/r02_d):
// with the scan's 128-read tiles (reads up to 128 bases) search 18.7 -> 16.8 ms per 10 M reads; with its 64-read tiles the
// sorted runs are too short for homogeneous batches and batches lose (150-bp pairs 104 -> 115 ms, k = 31 / -maxcorK 8
// 51 -> 56 ms): the launcher asks for batches only in the first case (RcxCorrectArgs::dfs_batch).
#define DFS_BATCH 1
#endif
// SMEM (reads up to 255 bases): path[] / fix[] are kept as EDIT LISTS in shared memory.  Every accepted path of a segment
// copies the same range path[start .. to) into fix[], and everything in that range was assigned by the nodes of the
// current root-to-leaf path, so "path[] restricted to the current path" is just the list of substitutions made on it
// (all other entries are -1), and fix[] is, per segment, the edit list of the LAST accepted path (a segment that ends
// "bad" after a trial-overflow reset keeps it, like the reference's never-rolled-back fix[]).  This removes the dense
// per-position writes, the O(segment) copy at every acceptance and the O(read) write-back.  One entry = pos | base << 8.
// Row capacities (entries) for a launch: edits on one path / accepted edits of a whole read.
__host__ __device__ inline int dfs_cap_path(int max_len, int k, int K) {
    int c = (max_len * K / k) * 2 + 1;
    if (c < K) c = K;
    c += 8;  // the child that exceeds maxFixCnt is entered before it returns; up to 5 N positions; slack
    return c < max_len ? c : max_len;
}
// Shared memory holds only the edits of the CURRENT path (touched at every substitution and every pop); the accepted edits
// of the read (written at the few terminals that improve on the best path, read once at the end) live in the
// lane-interleaved global scratch, so the rows no longer limit the blocks per SM for long reads or a large -maxcorK
// (150 bp: 4 -> 6 blocks; k = 31 with -maxcorK 8: 3 -> 6).  Neither list can overflow: the row holds what maxFixCnt allows
// on one path, and a path with more substitutions than that -- they are free on non-ACGT letters other than N, which the
// read filter does not count -- continues in a spill area of the same global scratch; both global lists have max_len
// entries (a position is edited at most once: the segments are disjoint and substitutions happen on untrusted positions).
__host__ __device__ inline int dfs_row_entries(int max_len, int k, int K) { return dfs_cap_path(max_len, k, K) | 1; }  // odd: rows spread over the banks
// VH ("vh mode", first pass only): the launch starts with warp batches over the listed very heavy items (at least three quarters
// of their k-windows untrusted) -- 32 listed slots per warp, a lane that finishes waits for the others -- before the items taken
// one by one.  Where such items are the bulk of the work (k = 31 / -maxcorK 8 / 5 % errors) they are alike and long, and batches
// of them are a fifth faster than the per-lane fetch (search 52.5 -> 41.7 ms per 8 M reads); where they are few (0.7 % of the
// items of 150-bp pairs at 1 % errors) the batches' idle lanes cost more than they give, and the mere presence of the phase
// logic slows the kernel by 8-20 % (profiles/r02_m_search_order_experiments.md) -- hence a template instantiation of its own,
// which the host picks for a batch when the PREVIOUS batch counted at least a quarter of its work items as very heavy.
template <bool SMEM, bool BATCH, bool VH>
__global__ void __launch_bounds__(DFS_THREADS, DFS_MINBLOCKS) rcx_dfs_kernel(RcxCorrectArgs a) {
    __shared__ int lut[DFS_LUT];
    extern __shared__ __align__(16) unsigned short dfs_rows[];
    const int lane = threadIdx.x & 31;
    const int k = a.t.k;
    const double er = a.error_rate;
    const unsigned long long mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    for (int c = threadIdx.x; c < DFS_LUT; c += DFS_THREADS) {
        const int r = rcx_bound_i(c, er);
        lut[c] = r < 1 ? 1 : r;
    }
    __syncthreads();
    const int stride = rcx_state_stride(a.max_len, k);
    // items taken one by one by the lanes (heavy ones of the first pass, everything of the second) and items taken by
    // whole warps (light ones of the first pass)
    const long long nheavy = a.pass == 0 ? (long long)a.counters[RCX_C_NWORKH] : 0;
    const long long nlight = a.pass == 0 ? (long long)a.counters[RCX_C_NWORK] : 0;
    const bool batch = BATCH && a.pass == 0;
    const long long ndyn = a.pass == 0 ? (batch ? nheavy : nheavy + nlight) : (long long)a.counters[RCX_C_NWORK2];
    const long long nbatch = batch ? nlight : 0;
    const long long nvh = (VH && a.pass == 0) ? (long long)a.counters[RCX_C_NVH] : 0;
    const int fetch_counter = a.pass == 0 ? RCX_C_WORKX1 : RCX_C_WORKX3;
    // per-thread scratch, interleaved by lane
    const long long gwarp = ((long long)blockIdx.x * DFS_THREADS + threadIdx.x) >> 5;
    unsigned int* stack = a.dfs_stack + gwarp * (long long)a.dfs_depth * DFS_FRAME * 32 + lane;
    signed char* path = a.dfs_path + gwarp * (long long)a.dfs_lpad * 32 + lane;  // dense path[] (long reads only)
    const int cap_path = dfs_cap_path(a.max_len, k, a.max_fix_per_k);  // row entries; edits beyond them go to the spill area
    unsigned short* E = dfs_rows + (size_t)threadIdx.x * dfs_row_entries(a.max_len, k, a.max_fix_per_k);  // edits on the current path
    // accepted edits of the read (SMEM mode): address rebuilt at the few uses instead of a pointer kept live across the walk
#define A(e) (reinterpret_cast<unsigned short*>(a.dfs_path + ((((long long)blockIdx.x * DFS_THREADS + threadIdx.x) >> 5) * (long long)a.dfs_lpad << 5)) + (threadIdx.x & 31))[(e) * 32]
#define ESPILL(e) A(a.max_len + (e))  // entries cap_path, cap_path + 1, ... of the current path's list
#define EGET(e) ((e) < cap_path ? (unsigned int)E[e] : (unsigned int)ESPILL((e) - cap_path))
    int ne = 0, na = 0, na_base = 0;
    bool edit_overflow = false;
#define STK(f, w) stack[((f) * DFS_FRAME + (w)) * 32]
#define PATH(p) path[(p) * 32]

    unsigned long long n_nodes = 0, n_probes = 0;
#ifdef RCX_DIAG
    unsigned long long diag_n0 = 0;
    int diag_nun = 0;
#endif
    // ---- per-item state
    int* st = nullptr;
    const unsigned char* pinfo = nullptr;
    signed char* fix = nullptr;
    int len = 0, nseg = 0, g = 0, trust = 0, fbound = 0, max_top2 = -1, total_fix = 0, bad_segments = 0, best_bn_all = RCX_INF;
    bool abort_read = false;
    long long item = 0;
    // ---- per-segment state (the by-reference parameters of SearchPaths_*)
    bool right = true;
    int start = 0, to = 0;
    int max_fix = 0, best_cnt = -1, best_bn = -1, trial = 0, top2a = -1, top2b = -1;
    // ---- current node
    int pos = 0, t = 0, fix_cnt = 0, bn = 0, thr = 0, ext = 0, sp = 0;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    RcxKmer kc;
    kc.code = 0; kc.inv = -1;
    // the lookup issued one node ahead (DFS_AHEAD): canonical context and its home sector, possibly still in flight
    unsigned long long pf_key = 0;
    RcxEntryRegs pf_d;
    pf_d.q0 = pf_d.q1 = pf_d.q2 = pf_d.q3 = 0;
    bool pf_on = false;
    bool vh_phase = VH && nvh > 0;  // (warp-uniform) still in the batches over the listed very heavy items
    int stage = vh_phase ? T_IDLE : T_FETCH;
    int4 seg0a = make_int4(0, 0, 0, 0), seg0b = make_int4(0, 0, 0, 0);
    long long nx_w = vh_phase ? -1 : (long long)atomicAdd(&a.counters[fetch_counter], 1ull);

    for (;;) {
        bool have_item = false;
        int expect = RCX_ST_SEARCH;
        if (VH && vh_phase) {
            if (stage == T_FETCH) stage = T_IDLE;  // wait for the batch-mates
        } else if (stage == T_FETCH) {
            if (VH && nx_w < 0) nx_w = (long long)atomicAdd(&a.counters[fetch_counter], 1ull);
            if (nx_w < ndyn) {  // reserved while the previous item was searched
                const long long w = nx_w;
                nx_w = (long long)atomicAdd(&a.counters[fetch_counter], 1ull);  // the next one: its latency hides behind this search
                item = a.pass == 0 ? rcx_work_slot(w, nheavy, a.n) : (long long)a.worklist2[w];  // heavy items first
                have_item = true;
            } else {
                stage = T_IDLE;
            }
        }
        {   // every lane comes by here once per trip: when the whole warp is out of work it takes the next 32 light items
            const unsigned int idle = __ballot_sync(RCX_FULL, stage == T_IDLE);
            if (idle == RCX_FULL) {
                long long base = 0;
                if (VH && vh_phase) {
                    if (lane == 0) base = (long long)atomicAdd(&a.counters[RCX_C_WORKV], 32ull);
                    base = __shfl_sync(RCX_FULL, base, 0);
                    if (base < nvh) {
                        if (base + lane < nvh) { item = (long long)a.vhlist[base + lane]; expect = RCX_ST_SEARCH_VH; have_item = true; }
                    } else {  // (uniform) the list is exhausted: on to the items taken one by one (the sweep passes the listed ones by)
                        vh_phase = false;
                        stage = T_FETCH;
                        continue;
                    }
                } else {
                    if (lane == 0 && nbatch > 0) base = (long long)atomicAdd(&a.counters[RCX_C_WORKB], 32ull);
                    base = __shfl_sync(RCX_FULL, base, 0);
                    if (nbatch <= 0 || base >= nbatch) break;  // (uniform) nothing left for this warp
                    if (base + lane < nbatch) { item = rcx_work_slot(nheavy + base + lane, nheavy, a.n); have_item = true; }
                }
            }
        }
        if (have_item) {
            st = a.state + item * stride;
            // header (8 words) + first segment in one round trip: four independent 16-byte loads
            const int4 h0 = *reinterpret_cast<const int4*>(st), h1 = *reinterpret_cast<const int4*>(st + 4);
            seg0a = *reinterpret_cast<const int4*>(st + 8);
            seg0b = *reinterpret_cast<const int4*>(st + 12);
            if (h0.x != (VH ? expect : RCX_ST_SEARCH)) { stage = T_FETCH; continue; }  // decided by the prep kernel (second pass), or (vh mode) a listed item met by the sweep: next one
            const uint32_t o = (uint32_t)h1.z;    // ST_OFF / ST_LEN: written with the header, no worklist -> off[] chase
            len = h1.w;
            pinfo = a.posbuf + o;
            fix = a.fixbuf + o;
            nseg = h1.y; trust = h1.x; fbound = h0.y;  // (ST_TOTALFIX carries rcx_force_bound(strong) in)
#ifdef RCX_DIAG
            diag_nun = h0.z; diag_n0 = n_nodes;
#endif
            max_top2 = -1;
            na_base = 0;  // fix[] = -1 (:1113-1114): no accepted edits yet
            g = 0; total_fix = 0; bad_segments = 0; best_bn_all = RCX_INF; abort_read = false;
            stage = T_SEG;
        }
        if (stage == T_SEG) {
            if (g >= nseg || abort_read) {
                // ---------------- decisions after the searches (:1176-1242)
                int result = RCX_ST_POSTFILTER;
                bool rethreshold = false;
                if (abort_read) {
                    result = -1;  // :1172
                } else {
                    bool force_next = false;
                    if (nseg > 0) {
                        // :1186-1199: bestFixCnt doubles for every segment whose top2FixBottleNeck[1] reaches the read's best
                        // bottleneck (any one of them rejects the read); the FP64 compare with GetBound(strong) as integers
                        if (a.pass == 0 && best_bn_all < fbound) force_next = true;
                        if (max_top2 >= best_bn_all) result = -1;
                    }
                    if (result == RCX_ST_POSTFILTER) {
                        const bool unfixable = total_fix > len;  // :1234 (never true in practice)
                        if (total_fix == 0 && force_next) result = 0;  // :1231
                        else if (!unfixable && !force_next) { /* leave the loop: post-filters */ }
                        else if (trust < 10 && !force_next) result = -1;
                        else rethreshold = true;
                    }
                }
                st[ST_TOTALFIX] = total_fix;
                st[ST_BADSEG] = bad_segments;
#ifdef RCX_DIAG
                {   // walk length of the item against its number of untrusted k-windows (buckets of 32)
                    const unsigned long long dn = n_nodes - diag_n0;
                    const int bkt = diag_nun >= 128 ? 3 : (diag_nun >> 5);
                    atomicMax(&a.counters[RCX_C_DIAG0], dn);
                    atomicAdd(&a.counters[RCX_C_DIAG0 + 1 + bkt], 1ull);                 // items per bucket
                    if (dn > 2000) { atomicAdd(&a.counters[RCX_C_DIAG0 + 5 + bkt], 1ull); atomicAdd(&a.counters[RCX_C_DIAG0 + 9 + bkt], dn); }
                }
#endif
                if (rethreshold) {
                    if (a.pass == 0) {
                        st[ST_STATUS] = RCX_ST_RETHRESHOLD;  // the pass-1 prep kernel re-thresholds (:1247-1289)
                        a.worklist2[atomicAdd(&a.counters[RCX_C_NWORK2], 1ull)] = (uint32_t)item;
                    } else {
                        // a third pass needs `unfixable`, i.e. totalFix > read length: cannot happen; flagged, not hidden
                        atomicAdd(&a.counters[RCX_C_UNSUPPORTED], 1ull);
                        st[ST_STATUS] = -1;
                    }
                } else {
                    st[ST_STATUS] = result;
                }
                // the post-filters read fix[] from HBM; a read that asked for a second pass may still keep this pass' edits
                // (no lower plateau found, :1288-1289).  fixbuf is -1 everywhere, so only the edits are written.
                if (SMEM && (rethreshold || result == RCX_ST_POSTFILTER)) {
                    if (a.compact_fix) {  // thread-per-item finalize: the edit list itself (count, then pos | base << 8) in the read's words
                        unsigned short* fl = a.fixlist + (fix - a.fixbuf);
                        if (na_base < len) {
                            fl[0] = (unsigned short)na_base;
                            for (int e = 0; e < na_base; ++e) fl[1 + e] = A(e);
                        } else {
                            edit_overflow = true;  // cannot happen: at least one position of a work item is trusted
                        }
                    } else {
                        for (int e = 0; e < na_base; ++e) { const unsigned int x = A(e); fix[x & 255u] = (signed char)(x >> 8); }
                    }
                }
                stage = T_FETCH;
                continue;
            }
            int w0, w1, w2, w3, w4;
            if (g == 0) { w0 = seg0a.x; w1 = seg0a.y; w2 = seg0a.z; w3 = seg0a.w; w4 = seg0b.x; }  // came with the header
            else { const int* sg = st + ST_HDR + ST_SEGW * g; w0 = sg[0]; w1 = sg[1]; w2 = sg[2]; w3 = sg[3]; w4 = sg[4]; }
            start = w0 & 0xffff; to = (w0 >> 16) & 0xffff;
            right = (w1 & 1) != 0;
            max_fix = w1 >> 1;
            kc.code = (unsigned long long)(unsigned int)w2 | ((unsigned long long)(unsigned int)w3 << 32);
            kc.inv = w4;
            best_cnt = -1; best_bn = -1; trial = 0; top2a = -1; top2b = -1;
            pos = start; t = trust; fix_cnt = 0; bn = RCX_INF; sp = 0;
            ne = 0; na = 0;
            stage = T_ENTER;
        }
        // Node entries run as an explicit inner loop: lanes whose node has a single child keep walking while the lanes that
        // need a transition (alternative, pop, next segment, next item) wait after the loop.  When fewer than DFS_MINLANES
        // lanes are still walking, they pause so that everybody re-enters together after one pass over the transitions.
        while (stage == T_ENTER) {
            ++n_nodes;
            bool leave = false;
            // step 0 (:211-224 / :453-467)
            if (trial > RCX_MAX_TRIAL) {
                if (max_fix > 2) { --max_fix; best_cnt = 0; best_bn = -1; trial = 0; }
                else leave = true;
            }
            if (!leave && fix_cnt > max_fix) leave = true;
            if (!leave && (right ? pos >= to : pos < to)) {
                // terminal (:228-285 / :468-524)
                int fc = fix_cnt;
                if (bn < t) ++fc;
                if (fc < max_fix) { top2a = bn; top2b = -1; }
                else if (fc == max_fix) {
                    if (bn > top2a) { top2b = top2a; top2a = bn; }
                    else if (bn > top2b) top2b = bn;
                }
                if (fc < max_fix || (fc == max_fix && bn > best_bn)) {
                    if (fc < max_fix) trial = -(max_fix - fc + 1) * RCX_MAX_TRIAL;
                    const int lo = right ? start : pos + 1, hi = right ? pos - 1 : start;
                    if (SMEM) {  // the edits of this path replace the segment's previously accepted ones
                        if (na_base + ne <= a.max_len) { for (int e = 0; e < ne; ++e) A(na_base + e) = (unsigned short)EGET(e); na = ne; }
                        else edit_overflow = true;  // cannot happen: every position is edited at most once
                    } else {
                        for (int i = lo; i <= hi; ++i) fix[i] = PATH(i);
                    }
                    max_fix = fc; best_bn = bn; best_cnt = 1;
                } else if (fc == max_fix && bn == best_bn) {
                    best_cnt += 1;
                }
                leave = true;
            }
            if (leave) {
                stage = T_POP;
            } else {
                // ---- the node's sector (the four one-base extensions of kc): taken from the lookup issued ahead by the parent
                // when it is the one (same canonical context), else issued now
                const bool cvalid = kc.inv == -1 || (right ? kc.inv + 1 >= k : kc.inv <= 0);
                const unsigned long long ctx = right ? (kc.code & (mask >> 2)) : (kc.code >> 2);
                unsigned long long key = 0;
                bool flip = false;
                RcxEntryRegs d;
                if (cvalid) {
                    const unsigned long long rc = rcx_revcomp(ctx, k - 1);
                    flip = rc < ctx;
                    key = flip ? rc : ctx;
                    if (DFS_AHEAD && pf_on && pf_key == key) d = pf_d;
                    else { rcx_issue(a.t, rcx_home(a.t, key), d); n_probes += 1; }
                }
                const int pi = pinfo[pos];
                // ---- one node ahead: if the read's base is kept here, the child's context is known before this node's counts
                // arrive -- its sector travels while this node is decided (a substitution or a pop just drops it)
                RcxKmer x = kc;
                const int hb = pi & 3;
                if (!(pi & 4)) { if (right) rcx_k_append(x, hb, false, k, mask); else rcx_k_prepend(x, hb, false, k, mask); }
                if (DFS_TAIL_PREFETCH > 0 && !(pi & 4) && __popc(__activemask()) <= DFS_TAIL_PREFETCH) {
                    // Few lanes of this warp are still walking (the end of a launch, or one very long walk): the warp has
                    // issue slots to spare and every node costs it a full DRAM round trip.  The keep-base child's sector is
                    // asked into L2 now (no register holds it), so that the child finds it there.
                    const int np1 = right ? pos + 1 : pos - 1;
                    if ((right ? np1 < to : np1 >= to) && (x.inv == -1 || (right ? x.inv + 1 >= k : x.inv <= 0))) {
                        const unsigned long long nctx = right ? (x.code & (mask >> 2)) : (x.code >> 2);
                        const unsigned long long nrc = rcx_revcomp(nctx, k - 1);
                        const RcxEntry* pe = a.t.e + rcx_home(a.t, nrc < nctx ? nrc : nctx);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pe));
                    }
                }
                if (DFS_AHEAD) {
                    pf_on = false;
                    const int np1 = right ? pos + 1 : pos - 1;
                    if (!(pi & 4) && (right ? np1 < to : np1 >= to) && (x.inv == -1 || (right ? x.inv + 1 >= k : x.inv <= 0))) {
                        const unsigned long long nctx = right ? (x.code & (mask >> 2)) : (x.code >> 2);
                        const unsigned long long nrc = rcx_revcomp(nctx, k - 1);
                        pf_key = nrc < nctx ? nrc : nctx;
                        rcx_issue(a.t, rcx_home(a.t, pf_key), pf_d);
                        n_probes += 1;  // (counted when issued: a substitution or a pop may leave it unused)
                        pf_on = true;
                    }
                }
                dfs_counts4(a.t, ctx, !right, d, flip, cvalid && rcx_find(a.t, key, d), c0, c1, c2, c3);
                thr = dfs_infer(c0, c1, c2, c3, t, er, lut);
                ext = 0;
                stage = T_ALT0;
                // ---- keep the read's base (:294-339 / :532-574)
                if (!(pi & 4)) {
                    int c = dfs_sel4(hb, c0, c1, c2, c3);
                    bool go = false;
                    int npos = 0, nfix = fix_cnt, nbn = bn;
                    if (c >= thr) {
                        if (!SMEM) PATH(pos) = -1;
                        go = true; npos = right ? pos + 1 : pos - 1; nbn = c < bn ? c : bn;
                    } else if (thr == 1 && t <= 2) {
                        // "accidentally below the threshold" rescue (:313-338 / :550-573)
                        int i = pos, steps = 0;
                        for (; c < thr && steps < k; ++steps) {
                            if (right) { ++i; if (i >= to) break; } else { --i; if (i < to) break; }
                            const int qi = pinfo[i];
                            if (right) rcx_k_append(x, qi & 3, (qi & 4) != 0, k, mask); else rcx_k_prepend(x, qi & 3, (qi & 4) != 0, k, mask);
                            c = rcx_k_count(a.t, x);
                            ++n_probes;
                        }
                        if (steps < k && (right ? i < to : i >= to)) {
                            const int lo = right ? pos : i, hi = right ? i : pos;
                            if (!SMEM) { for (int j = lo; j <= hi; ++j) PATH(j) = -1; }
                            go = true; npos = right ? i + 1 : i - 1; nfix = fix_cnt + 1;
                        }
                    }
                    if (go) {
                        ext = 1;
                        // when no alternative base can pass, the parent has nothing left to do after this child
                        // (ext != 0 rules out the jump): tail call, no frame
                        const bool allowed = (pi & (right ? 8 : 16)) != 0;
                        const bool alt_any = allowed && ((hb != 0 && c0 >= thr) || (hb != 1 && c1 >= thr) || (hb != 2 && c2 >= thr) ||
                                                         (hb != 3 && c3 >= thr));
                        if (alt_any) {
                            STK(sp, 0) = (unsigned int)kc.code; STK(sp, 1) = (unsigned int)(kc.code >> 32);
                            STK(sp, 2) = (unsigned int)(pos + 1) | ((unsigned int)(kc.inv + 1) << 12) | ((unsigned int)T_ALT0 << 20) | (1u << 24);
                            STK(sp, 3) = (unsigned int)t; STK(sp, 4) = (unsigned int)fix_cnt | ((unsigned int)ne << 16); STK(sp, 5) = (unsigned int)bn;
                            STK(sp, 6) = (unsigned int)thr;
                            STK(sp, 7) = (unsigned int)c0; STK(sp, 8) = (unsigned int)c1; STK(sp, 9) = (unsigned int)c2; STK(sp, 10) = (unsigned int)c3;
                            ++sp;
                        }
                        if (!right) t = thr;  // Left hands its threshold down (:546,569)
                        pos = npos; fix_cnt = nfix; bn = nbn; kc = x;
                        stage = T_ENTER;
                        if (DFS_MINLANES > 0 && __popc(__activemask()) < DFS_MINLANES) break;  // few lanes still chaining: let the others catch up
                        continue;
                    }
                }
            }
        }
        if (stage >= T_ALT0 && stage < T_JUMP) {  // alternatives in the order A, C, G, T (:343-372 / :577-608)
            const int pi = pinfo[pos];
            const bool here_ok = !(pi & 4);
            const int hb = pi & 3;
            bool pushed = false;
            if (pi & (right ? 8 : 16)) {
                for (int b = stage - T_ALT0; b < 4; ++b) {
                    if (here_ok && b == hb) continue;
                    const int c = dfs_sel4(b, c0, c1, c2, c3);
                    if (c >= thr) {
                        if (SMEM) {
                            if (ne < cap_path) E[ne] = (unsigned short)(pos | (b << 8));
                            else if (ne - cap_path < a.max_len) ESPILL(ne - cap_path) = (unsigned short)(pos | (b << 8));
                            else edit_overflow = true;  // cannot happen: one substitution per position on a path
                        } else PATH(pos) = (signed char)b;
                        ++trial;
                        ext = 1;
                        // a frame is only needed if another alternative can still pass after this one: otherwise the parent
                        // has nothing left to do when this child returns (ext != 0 rules out the jump) -- tail call
                        bool more = false;
                        for (int b2 = b + 1; b2 < 4; ++b2) more |= !(here_ok && b2 == hb) && dfs_sel4(b2, c0, c1, c2, c3) >= thr;
                        if (more) {
                            STK(sp, 0) = (unsigned int)kc.code; STK(sp, 1) = (unsigned int)(kc.code >> 32);
                            STK(sp, 2) = (unsigned int)(pos + 1) | ((unsigned int)(kc.inv + 1) << 12) | ((unsigned int)(T_ALT0 + b + 1) << 20) | (1u << 24);
                            STK(sp, 3) = (unsigned int)t; STK(sp, 4) = (unsigned int)fix_cnt | ((unsigned int)ne << 16); STK(sp, 5) = (unsigned int)bn;
                            STK(sp, 6) = (unsigned int)thr;
                            STK(sp, 7) = (unsigned int)c0; STK(sp, 8) = (unsigned int)c1; STK(sp, 9) = (unsigned int)c2; STK(sp, 10) = (unsigned int)c3;
                            ++sp;
                        }
                        if (SMEM) ++ne;  // the substitution joins the current path (the frame keeps the count before it)
                        if (right) rcx_k_append(kc, b, false, k, mask); else rcx_k_prepend(kc, b, false, k, mask);
                        if (!right) t = thr;
                        pos = right ? pos + 1 : pos - 1;
                        if (here_ok) ++fix_cnt;
                        if (c < bn) bn = c;
                        stage = T_ENTER;
                        pushed = true;
                        break;
                    }
                }
            }
            if (pushed) continue;
            stage = T_JUMP;
        }
        if (stage == T_JUMP) {
            stage = T_POP;
            if (ext == 0) {  // jump through an unfixable region (:393-441 / :629-677); last child => tail call
                RcxKmer x = kc;
                int i, th = thr;
                if (right) {
                    for (i = pos; i < to; ++i) {
                        int d0, d1, d2, d3;
                        dfs_probe4(a.t, x, true, mask, d0, d1, d2, d3);
                        n_probes += 1;
                        th = dfs_infer(d0, d1, d2, d3, t, er, lut);
                        const int qi = pinfo[i];
                        const bool bad = (qi & 4) != 0;
                        rcx_k_append(x, qi & 3, bad, k, mask);
                        const int c = bad ? 0 : dfs_sel4(qi & 3, d0, d1, d2, d3);
                        if (!SMEM) PATH(i) = -1;
                        if (c >= th) break;
                    }
                    int kk = (i < len) ? i - pos - k + 1 : (i - pos) / 2;
                    if (kk <= 0) kk = 1;
                    if (i >= to) i -= 1;
                    pos = i + 1; fix_cnt += kk; kc = x;
                } else {
                    for (i = pos; i >= to; --i) {
                        int d0, d1, d2, d3;
                        dfs_probe4(a.t, x, false, mask, d0, d1, d2, d3);
                        n_probes += 1;
                        th = dfs_infer(d0, d1, d2, d3, t, er, lut);
                        const int qi = pinfo[i];
                        const bool bad = (qi & 4) != 0;
                        rcx_k_prepend(x, qi & 3, bad, k, mask);
                        const int c = bad ? 0 : dfs_sel4(qi & 3, d0, d1, d2, d3);
                        if (!SMEM) PATH(i) = -1;
                        if (c >= th) break;
                    }
                    int kk = (i >= 0) ? pos - i - k + 1 : pos - 1;
                    if (kk <= 0) kk = 1;
                    if (i <= to) ++i;
                    pos = i - 1; fix_cnt += kk; kc = x; t = th;
                }
                stage = T_ENTER;
                continue;
            }
        }
        if (stage == T_POP) {
            if (sp == 0) {
                // ---------------- segment finished (:1162-1174)
                if (top2b > max_top2) max_top2 = top2b;
                na_base += na;  // fix[] keeps the segment's last accepted edits whether or not the segment ends "bad"
                if (best_bn == -1) {
                    ++bad_segments;
                } else {
                    if (best_bn < best_bn_all) best_bn_all = best_bn;
                    if (trial > RCX_MAX_TRIAL) abort_read = true;
                    else total_fix += max_fix;
                }
                ++g;
                stage = T_SEG;
                continue;
            }
            --sp;
            kc.code = (unsigned long long)STK(sp, 0) | ((unsigned long long)STK(sp, 1) << 32);
            const unsigned int w = STK(sp, 2);
            pos = (int)(w & 0xfff) - 1;
            kc.inv = (int)((w >> 12) & 0xff) - 1;
            stage = (int)((w >> 20) & 0xf);
            ext = (int)((w >> 24) & 1);
            t = (int)STK(sp, 3); bn = (int)STK(sp, 5); thr = (int)STK(sp, 6);
            { const unsigned int w4 = STK(sp, 4); fix_cnt = (int)(w4 & 0xffffu); ne = (int)(w4 >> 16); }
            c0 = (int)STK(sp, 7); c1 = (int)STK(sp, 8); c2 = (int)STK(sp, 9); c3 = (int)STK(sp, 10);
        }
    }
#undef STK
#undef PATH
#undef EGET
#undef ESPILL
#undef A
    // counters: one atomic per warp
    for (int o = 16; o > 0; o >>= 1) {
        n_nodes += __shfl_xor_sync(RCX_FULL, n_nodes, o);
        n_probes += __shfl_xor_sync(RCX_FULL, n_probes, o);
    }
    if (edit_overflow) atomicAdd(&a.counters[RCX_C_UNSUPPORTED], 1ull);  // more substitutions on one path than the row holds: flagged, not hidden
    if (lane == 0) {
        if (n_nodes) atomicAdd(&a.counters[RCX_C_SEARCH_NODES], n_nodes);
        if (n_probes) atomicAdd(&a.counters[RCX_C_SEARCH_PROBES], n_probes);
    }
}

// scratch requirements of one launch
static bool dfs_smem_rows(int max_len) { return max_len <= 255; }
static size_t dfs_smem_bytes(int max_len, int k, int K) {
    return dfs_smem_rows(max_len) ? sizeof(unsigned short) * (size_t)DFS_THREADS * dfs_row_entries(max_len, k, K) : 0;
}
void rcx_dfs_geometry(int sm_count, int max_len, int k, int K, int* grid, int* depth, int* lpad) {
    int per_sm = 0;
    const size_t smem = dfs_smem_bytes(max_len, k, K);
    cudaError_t e;
    if (dfs_smem_rows(max_len)) {
        cudaFuncSetAttribute(rcx_dfs_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(rcx_dfs_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(rcx_dfs_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(rcx_dfs_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rcx_dfs_kernel<true, false, false>, DFS_THREADS, smem);
        int other = 0;  // one grid (and scratch) size for every instantiation
        if (e == cudaSuccess && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&other, rcx_dfs_kernel<true, true, false>, DFS_THREADS, smem) == cudaSuccess && other < per_sm) per_sm = other;
        if (e == cudaSuccess && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&other, rcx_dfs_kernel<true, false, true>, DFS_THREADS, smem) == cudaSuccess && other < per_sm) per_sm = other;
        if (e == cudaSuccess && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&other, rcx_dfs_kernel<true, true, true>, DFS_THREADS, smem) == cudaSuccess && other < per_sm) per_sm = other;
    } else {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rcx_dfs_kernel<false, false, false>, DFS_THREADS, 0);
    }
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;
    *grid = sm_count * per_sm;
    int kc = max_len - k + 1;
    if (kc < 1) kc = 1;
    *depth = kc + 3;
    // bytes per thread of the lane-interleaved scratch: the dense path[] of long reads, or the accepted edit list + the spill area of
    // the current path's list (max_len 16-bit entries each)
    *lpad = dfs_smem_rows(max_len) ? 4 * max_len + 8 : max_len + 2;
}
size_t rcx_dfs_stack_words(int grid, int depth) { return (size_t)grid * (DFS_THREADS / 32) * depth * DFS_FRAME * 32; }
size_t rcx_dfs_path_bytes(int grid, int lpad) { return (size_t)grid * (DFS_THREADS / 32) * lpad * 32; }

cudaError_t rcx_launch_dfs(const RcxCorrectArgs& a, int grid, cudaStream_t s) {
    const size_t smem = dfs_smem_bytes(a.max_len, a.t.k, a.max_fix_per_k);
    if (dfs_smem_rows(a.max_len)) {
        const bool batch = DFS_BATCH && a.dfs_batch && a.pass == 0, vh = a.vh_mode && a.pass == 0;
        if (batch && vh) rcx_dfs_kernel<true, true, true><<<grid, DFS_THREADS, smem, s>>>(a);
        else if (batch) rcx_dfs_kernel<true, true, false><<<grid, DFS_THREADS, smem, s>>>(a);
        else if (vh) rcx_dfs_kernel<true, false, true><<<grid, DFS_THREADS, smem, s>>>(a);
        else rcx_dfs_kernel<true, false, false><<<grid, DFS_THREADS, smem, s>>>(a);
    } else {
        rcx_dfs_kernel<false, false, false><<<grid, DFS_THREADS, 0, s>>>(a);
    }
    return cudaGetLastError();
}
