// rcx_correct.cu — K2: per-read correction, one warp owns one read.
//
// Restates ErrorCorrection (ErrorCorrection.cpp:682-1480), SearchPaths_Right /
// SearchPaths_Left (:201-442 / :444-678), InferPosThreshold (:144-173) and
// GetKmerInformation (:1567-1602) as a warp-cooperative state machine:
//
//  * every lane keeps the (warp-uniform) scalar state of the algorithm in
//    registers, so control flow never diverges; lanes split the data-parallel
//    parts: bit-plane construction, trusted-k-mer ballots, range fills, the
//    four one-base-extension probes of a search node (lanes 0..3, one 32-byte
//    sector each), best-path copies, filters and the final re-scan;
//  * the recursion of the reference becomes an explicit stack of 48-byte
//    frames in shared memory (depth <= kcnt + 2); children are visited in the
//    reference's order (keep-base / rescue, A, C, G, T, jump) and all state
//    the reference shares by reference across the recursion (maxFixCnt,
//    bestFixCnt, bestFixBottleNeck, top2, trialCnt, the never-rolled-back
//    path[] / fix[] arrays) is shared here in the same way;
//  * the four extension k-mers of a node serve InferPosThreshold, the
//    keep-base test and the three alternatives at once (the CPU probes the
//    same k-mers 8 times).
//
// counts[] and the per-read GetStrongTrustedThreshold come from the scan
// kernel (rcx_scan.cu); the pair threshold is min over the two mates
// (ErrorCorrection.cpp:96-106).
#include "rcx_internal.h"

#define COR_WARPS 4
#ifdef RCX_PHASE_TIMERS
#define PH(i) do { const long long _now = clock64(); ph[ph_cur] += (unsigned long long)(_now - ph_last); ph_last = _now; ph_cur = (i); } while (0)
#define DBG_INC(x) (++(x))
#else
#define PH(i) do { } while (0)
#define DBG_INC(x) do { } while (0)
#endif
#define FRAME_WORDS 12

enum { ST_ENTER = 0, ST_KEEP = 1, ST_ALT0 = 2, ST_JUMP = 6, ST_DONE = 7 };

struct CorLayout {
    int pkw, mw, lpad, n2, nseg, depth;
    int words;  // total 32-bit words per warp
};

__host__ __device__ inline CorLayout cor_layout(int max_len, int k) {
    CorLayout L;
    int kc = max_len - k + 1;
    if (kc < 1) kc = 1;
    L.pkw = rcx_pk_words(max_len);
    L.mw = rcx_mask_words(max_len);
    L.lpad = ((max_len + 31) / 32) * 32 + 32;
    L.n2 = rcx_pow2_at_least(kc);
    L.nseg = max_len / k + 4;
    L.depth = kc + 3;
    int un = L.depth * FRAME_WORDS;
    if (un < L.n2) un = L.n2;
    L.words = L.pkw + 10 * L.mw + L.lpad /*counts*/ + 3 * L.lpad / 4 /*path, fix, posinfo as int8*/ + L.nseg + un;
    return L;
}

size_t rcx_correct_smem_per_warp(int max_len, int k) { return sizeof(unsigned int) * (size_t)cor_layout(max_len, k).words; }

struct WarpMem {
    RcxPlanes P;
    unsigned int* strong;   // isStrongTrusted, bit per base
    unsigned int* island;   // islands used to derive segments (== strong unless the fallback island is used)
    unsigned int* pstrict;  // isPolyAKmer (threshold 2), bit per k-mer
    unsigned int* trusted;  // trusted k-mers of this pass
    unsigned int* fixm;     // scratch masks for the post-filters
    int* counts;
    signed char* path;
    signed char* fix;
    unsigned char* posinfo;  // per base: code (bits 0-1) | not-ACGT (bit 2) | substitution allowed in a Right search (bit 3) / Left search (bit 4)
    int* segtop;            // top2FixBottleNeck[1] per segment
    int* sbuf;              // sort buffer (aliases the search stack)
    unsigned int* stack;
};

// state shared by the whole search of one segment (the by-reference parameters of SearchPaths_*)
struct SearchState {
    int max_fix, best_cnt, best_bn, trial, top2a, top2b;
};

__device__ __noinline__ void warp_sort_ool(int* s, int n2, int lane) { rcx_warp_sort(s, n2, lane); }
__device__ __noinline__ void build_planes_ool(const RcxPlanes& P, const char* __restrict__ seq, int len, int max_len, int lane, int& nN,
                                              int& nA, int& nT) {
    rcx_build_planes(P, seq, len, max_len, lane, nN, nA, nT);
}

__device__ __forceinline__ int getbit(const unsigned int* m, int i) { return (m[i >> 5] >> (i & 31)) & 1; }

// first index in [from, limit) whose bit equals `want`, or `limit`; warp-cooperative, all lanes must call
__device__ __noinline__ int next_bit(const unsigned int* m, int from, int limit, bool want, int lane) {
    int best = limit;
    if (from < limit) {
        const int w0 = from >> 5;
        #pragma unroll 1
        for (int w = w0 + lane; (w << 5) < limit; w += 32) {
            unsigned int x = want ? m[w] : ~m[w];
            if (w == w0) x &= 0xffffffffu << (from & 31);
            if (x) {
                const int idx = (w << 5) + __ffs(x) - 1;
                if (idx < best) best = idx;
                break;
            }
        }
    }
    best = __reduce_min_sync(RCX_FULL, best);
    return best < limit ? best : limit;
}

// set bits [a, b] (inclusive) of a mask plane; warp-cooperative
__device__ __noinline__ void set_range(unsigned int* m, int a, int b, int mw, int lane) {
    if (a > b) return;
    #pragma unroll 1
    for (int w = (a >> 5) + lane; w <= (b >> 5) && w < mw; w += 32) {
        const int lo = a > (w << 5) ? a - (w << 5) : 0;
        const int hi = b < (w << 5) + 31 ? b - (w << 5) : 31;
        const unsigned int bits = (hi - lo == 31) ? 0xffffffffu : (((1u << (hi - lo + 1)) - 1) << lo);
        m[w] |= bits;
    }
}

__device__ __forceinline__ int sel4(int b, int c0, int c1, int c2, int c3) { return b == 0 ? c0 : b == 1 ? c1 : b == 2 ? c2 : c3; }

// counts of the four one-base extensions of kc (lanes 0..3 probe, everyone receives)
__device__ __noinline__ void probe4(const RcxTable& t, const RcxKmer& kc, bool right, unsigned long long mask, int lane, int& c0,
                                       int& c1, int& c2, int& c3) {
    int mine = 0;
    if (lane < 4) {
        RcxKmer x = kc;
        if (right) rcx_k_append(x, lane, false, t.k, mask);
        else rcx_k_prepend(x, lane, false, t.k, mask);
        mine = x.inv != -1 ? 0 : rcx_probe_ool(t, rcx_canonical(x.code, t.k));
    }
    __syncwarp();
    c0 = __shfl_sync(RCX_FULL, mine, 0);
    c1 = __shfl_sync(RCX_FULL, mine, 1);
    c2 = __shfl_sync(RCX_FULL, mine, 2);
    c3 = __shfl_sync(RCX_FULL, mine, 3);
}

// InferPosThreshold (:144-173) given the four extension counts
__device__ __forceinline__ int infer_thr(int c0, int c1, int c2, int c3, int upper, double er) {
    int mx = c0 > c1 ? c0 : c1;
    int my = c2 > c3 ? c2 : c3;
    mx = mx > my ? mx : my;
    if (mx < 0) mx = 0;
    int r = rcx_bound_i_ool(mx, er);
    if (r < 1) r = 1;
    return (upper > r || upper <= 0) ? r : upper;
}

// ---------------------------------------------------------------------------------------------- speculative frontier
// A search node needs the counts of its 4 one-base extensions, and the next node's k-mer depends on which child is
// taken, so a literal translation pays one dependent DRAM round trip per node.  Almost every node, however, is a
// "keep the read's base" step (after a substitution the search walks k-1 more positions to re-validate the
// following k-mers).  So one ROUND probes, in a single batch of independent loads, nodes along the read's own bases:
//   * FULL round: the 4 extensions of the current node and of the next 8*SPEC_U-1 nodes (lane 4*g+b probes base b of
//     level 8*u+g); results are transposed to one level per lane (4 counts + the InferPosThreshold bound, so
//     GetBound's FP64 sqrt runs once per round, not once per node);
//   * LITE round (node on a position where substitutions are not allowed: strong-trusted base or poly-A k-mer):
//     only the keep-base k-mer of up to 32 levels, one per lane.  A level passes without the other three counts when
//     keep >= t and (int)GetBound(keep) >= t: GetBound is monotone, so InferPosThreshold returns exactly t.
// Runs of nodes whose only viable child is the keep-base child are then consumed in one warp-parallel step
// ("fast-forward").  Results, order of side effects and tie-breaking are those of the node-by-node recursion.
#define SPEC_U 2
#define SPEC_LEVELS (8 * SPEC_U)

struct SpecRegs {
    int c0, c1, c2, c3;  // full round: counts of the 4 extensions of level `lane`; lite round: c0 = keep-base count
    int r;               // full: max(1, (int)GetBound(max of the 4)); lite: (int)GetBound(keep)
};

__device__ __noinline__ RcxKmer advance_kmer(const RcxPlanes& P, const RcxKmer& kc, bool right, int pos, int n, int k,
                                                unsigned long long mask) {
    // kc after n keep-base steps over VALID read bases starting at pos (n < k)
    if (n == 0) return kc;
    RcxKmer x;
    if (right) {
        const unsigned long long W = rcx_window(P.pk, pos, n);
        x.code = ((kc.code << (2 * n)) | W) & mask;
        x.inv = kc.inv == -1 ? -1 : (kc.inv + n >= k ? -1 : kc.inv + n);
    } else {
        const unsigned long long W = rcx_window(P.pk, pos - n + 1, n);
        x.code = (kc.code >> (2 * n)) | (W << (2 * (k - n)));
        x.inv = kc.inv == -1 ? -1 : (kc.inv - n < 0 ? -1 : kc.inv - n);
    }
    return x;
}

__device__ __noinline__ void spec_round_full(const RcxTable& tab, const RcxPlanes& P, const RcxKmer& kc, bool right, int pos, int D,
                                                unsigned long long mask, double er, int lane, SpecRegs& R,
                                                unsigned long long& n_probes) {
    const int k = tab.k;
    const int g = lane >> 2, b = lane & 3;
    unsigned long long canon[SPEC_U], home[SPEC_U];
    RcxBucketRegs d[SPEC_U];
    bool live[SPEC_U];
    int cnt[SPEC_U];
#pragma unroll
    for (int u = 0; u < SPEC_U; ++u) {
        const int lvl = u * 8 + g;
        live[u] = false;
        cnt[u] = 0;
        if (u * 8 < D) {  // warp-uniform: skip whole groups of levels that are not wanted
            if (lvl < D) {
                RcxKmer x = advance_kmer(P, kc, right, pos, lvl, k, mask);
                if (right) rcx_k_append(x, b, false, k, mask);
                else rcx_k_prepend(x, b, false, k, mask);
                if (x.inv == -1) {
                    live[u] = true;
                    canon[u] = rcx_canonical(x.code, k);
                    home[u] = rcx_home(tab, canon[u]);
                    rcx_issue(tab, home[u], d[u]);
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < SPEC_U; ++u) {
        if (u * 8 < D) {
            if (live[u]) cnt[u] = rcx_resolve(tab, canon[u], home[u], d[u]);
            n_probes += __popc(__ballot_sync(RCX_FULL, live[u]));
        }
    }
    __syncwarp();
    // transpose: lane i gathers the 4 counts of level i from lanes 4*(i&7)+b of register u = i>>3
    const int src = 4 * (lane & 7), myu = lane >> 3;
    int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int u = 0; u < SPEC_U; ++u) {
        if (u * 8 < D) {
            const int x0 = __shfl_sync(RCX_FULL, cnt[u], src), x1 = __shfl_sync(RCX_FULL, cnt[u], src + 1);
            const int x2 = __shfl_sync(RCX_FULL, cnt[u], src + 2), x3 = __shfl_sync(RCX_FULL, cnt[u], src + 3);
            if (myu == u) { a0 = x0; a1 = x1; a2 = x2; a3 = x3; }
        }
    }
    R.c0 = a0; R.c1 = a1; R.c2 = a2; R.c3 = a3;
    int mx = a0 > a1 ? a0 : a1;
    const int my = a2 > a3 ? a2 : a3;
    mx = mx > my ? mx : my;
    int r = rcx_bound_i_ool(mx < 0 ? 0 : mx, er);
    R.r = r < 1 ? 1 : r;
}

__device__ __noinline__ void spec_round_lite(const RcxTable& tab, const RcxPlanes& P, const RcxKmer& kc, bool right, int pos, int D,
                                                unsigned long long mask, double er, int lane, SpecRegs& R,
                                                unsigned long long& n_probes) {
    const int k = tab.k;
    int c = 0;
    bool live = false;
    if (lane < D) {  // level `lane`: the k-mer after keeping the read's bases pos .. pos+lane (all valid)
        const RcxKmer x = advance_kmer(P, kc, right, pos, lane + 1, k, mask);
        if (x.inv == -1) {
            live = true;
            c = rcx_probe_ool(tab, rcx_canonical(x.code, k));
        }
    }
    n_probes += __popc(__ballot_sync(RCX_FULL, live));
    R.c0 = c; R.c1 = 0; R.c2 = 0; R.c3 = 0;
    R.r = rcx_bound_i_ool(c < 0 ? 0 : c, er);
}

// One whole SearchPaths_Right (right = true) or SearchPaths_Left call tree.
__device__ __forceinline__ void search_segment(const RcxTable& tab, const WarpMem& M, SearchState& S, const bool right, const int start,
                               const int to, const int t0, RcxKmer kc, const int len, const double er, const int lane,
                               unsigned long long& n_nodes, unsigned long long& n_probes, unsigned long long* dbg) {
    const int k = tab.k;
    const unsigned long long mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    (void)dbg;
    int pos = start, t = t0, fix_cnt = 0, bn = RCX_INF, thr = 0, ext = 0, stage = ST_ENTER;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    int sp = 0;
    // speculation state (warp-uniform) + per-lane results
    SpecRegs R;
    R.c0 = R.c1 = R.c2 = R.c3 = 0; R.r = 1;
    int spec_lvl = -1, spec_n = 0, spec_pos0 = 0;
    bool first = true, spec_lite = false, force_full = false;

#define PUSH_CHILD(npos, nt, nfix, nbn, nkc, nlvl)                                                    \
    do {                                                                                              \
        if (lane == 0) {                                                                              \
            unsigned int* f = M.stack + sp * FRAME_WORDS;                                             \
            f[0] = (unsigned int)kc.code;                                                             \
            f[1] = (unsigned int)(kc.code >> 32);                                                     \
            f[2] = (unsigned int)(pos + 1) | ((unsigned int)(kc.inv + 1) << 12) | ((unsigned int)stage << 20) | \
                   ((unsigned int)(ext != 0) << 24);                                                  \
            f[3] = (unsigned int)t; f[4] = (unsigned int)fix_cnt; f[5] = (unsigned int)bn; f[6] = (unsigned int)thr; \
            f[7] = (unsigned int)c0; f[8] = (unsigned int)c1; f[9] = (unsigned int)c2; f[10] = (unsigned int)c3; \
        }                                                                                             \
        ++sp;                                                                                         \
        { const int _p = (npos), _t = (nt), _f = (nfix), _b = (nbn), _l = (nlvl); const RcxKmer _k = (nkc); \
          pos = _p; t = _t; fix_cnt = _f; bn = _b; kc = _k; spec_lvl = _l; }                          \
        stage = ST_ENTER;                                                                             \
        __syncwarp();                                                                                 \
    } while (0)
    // the jump child is always the last child of its node, so the parent frame is not needed again: no push.
    // (This also bounds the stack: a Left jump can re-enter the same position, ErrorCorrection.cpp:672-676.)
#define TAIL_CHILD(npos, nt, nfix, nkc)                                                               \
    do {                                                                                              \
        { const int _p = (npos), _t = (nt), _f = (nfix); const RcxKmer _k = (nkc);                    \
          pos = _p; t = _t; fix_cnt = _f; kc = _k; spec_lvl = -1; }                                   \
        stage = ST_ENTER;                                                                             \
        __syncwarp();                                                                                 \
    } while (0)

    for (;;) {
        if (stage == ST_ENTER) {
            ++n_nodes;
            bool leave = false;
            // step 0 (:211-224 / :453-467)
            if (S.trial > RCX_MAX_TRIAL) {
                if (S.max_fix > 2) { --S.max_fix; S.best_cnt = 0; S.best_bn = -1; S.trial = 0; }
                else leave = true;
            }
            if (!leave && fix_cnt > S.max_fix) leave = true;
            if (!leave && (right ? pos >= to : pos < to)) {
                // terminal (:228-285 / :468-524)
                int fc = fix_cnt;
                if (bn < t) ++fc;
                if (fc < S.max_fix) { S.top2a = bn; S.top2b = -1; }
                else if (fc == S.max_fix) {
                    if (bn > S.top2a) { S.top2b = S.top2a; S.top2a = bn; }
                    else if (bn > S.top2b) S.top2b = bn;
                }
                if (fc < S.max_fix || (fc == S.max_fix && bn > S.best_bn)) {
                    if (fc < S.max_fix) S.trial = -(S.max_fix - fc + 1) * RCX_MAX_TRIAL;
                    const int lo = right ? start : pos + 1, hi = right ? pos - 1 : start;
                    __syncwarp();
                    #pragma unroll 1
                    for (int i = lo + lane; i <= hi; i += 32) M.fix[i] = M.path[i];
                    __syncwarp();
                    S.max_fix = fc; S.best_bn = bn; S.best_cnt = 1;
                } else if (fc == S.max_fix && bn == S.best_bn) {
                    S.best_cnt += 1;
                }
                leave = true;
            }
            if (leave) {
                stage = ST_DONE;
            } else {
                if (!(spec_lvl >= 0 && spec_lvl < spec_n)) {
                    // new round from this node
                    const int remaining = right ? to - pos : pos - to + 1;
                    int run;  // consecutive valid read bases from pos in the search direction
                    if (right) {
                        const unsigned int m = rcx_bits(M.P.inv, pos);
                        run = m ? __ffs(m) - 1 : 32;
                    } else {
                        const unsigned int m = pos >= 31 ? rcx_bits(M.P.inv, pos - 31) : (M.P.inv[0] << (31 - pos));
                        run = m ? __clz(m) : 32;
                    }
                    const bool allowed0 = (M.posinfo[pos] & (right ? 8 : 16)) != 0;
                    spec_lite = !allowed0 && !force_full && t > 0 && run >= 1;
                    int D;
                    if (spec_lite) {
                        D = remaining < 32 ? remaining : 32;
                        if (D > k - 1) D = k - 1;
                        if (D > run) D = run;
                    } else {
                        // 1 level at the root of the search (its keep-base child usually fails: the root sits on the first
                        // untrusted position), full depth afterwards
                        D = first ? 1 : SPEC_LEVELS;
                        if (D > remaining) D = remaining;
                        if (D > k) D = k;
                        if (D > run + 1) D = run + 1;
                    }
#ifdef RCX_PHASE_TIMERS
                    const long long rt0 = clock64();
#endif
                    if (spec_lite) spec_round_lite(tab, M.P, kc, right, pos, D, mask, er, lane, R, n_probes);
                    else spec_round_full(tab, M.P, kc, right, pos, D, mask, er, lane, R, n_probes);
#ifdef RCX_PHASE_TIMERS
                    dbg[6] += (unsigned long long)(clock64() - rt0);
                    dbg[7] += D;
#endif
                    DBG_INC(dbg[0]);
                    spec_lvl = 0; spec_n = D; spec_pos0 = pos;
                    first = false;
                    force_full = false;
                }
                // ---- fast-forward over nodes whose only viable child is the keep-base child (lane i = level i)
                {
                    const int i = lane;
                    const bool inr = i >= spec_lvl && i < spec_n;
                    const int p = right ? spec_pos0 + i : spec_pos0 - i;
                    const int pi = inr ? M.posinfo[p] : 4;
                    const bool hok = !(pi & 4), allowed_i = (pi & (right ? 8 : 16)) != 0;
                    const int hb_i = pi & 3;
                    const int t_eff = t > 0 ? t : 0x7fffffff;
                    int thr_i = t_eff, keepc;
                    bool ok;
                    if (spec_lite) {
                        keepc = R.c0;
                        ok = inr && hok && !allowed_i && keepc >= t_eff && R.r >= t_eff;
                    } else {
                        if (right) {
                            thr_i = t_eff < R.r ? t_eff : R.r;
                        } else {  // Left hands its threshold down: prefix minimum from the current level
                            int v = inr ? R.r : 0x7fffffff;
#pragma unroll
                            for (int o = 1; o < SPEC_LEVELS; o <<= 1) {
                                const int u = __shfl_up_sync(RCX_FULL, v, o);
                                if (lane >= o && u < v) v = u;
                            }
                            thr_i = t_eff < v ? t_eff : v;
                        }
                        keepc = sel4(hb_i, R.c0, R.c1, R.c2, R.c3);
                        const bool alt_any = (R.c0 >= thr_i && !(hok && hb_i == 0)) || (R.c1 >= thr_i && !(hok && hb_i == 1)) ||
                                             (R.c2 >= thr_i && !(hok && hb_i == 2)) || (R.c3 >= thr_i && !(hok && hb_i == 3));
                        ok = inr && hok && keepc >= thr_i && !(allowed_i && alt_any);
                    }
                    const unsigned int okmask = __ballot_sync(RCX_FULL, ok);
                    const int n_ff = __ffs(~(okmask >> spec_lvl)) - 1;
                    if (n_ff > 0) {
                        const bool mine = i >= spec_lvl && i < spec_lvl + n_ff;
                        const int mn = __reduce_min_sync(RCX_FULL, mine ? keepc : 0x7fffffff);
                        if (mn < bn) bn = mn;
                        if (mine) M.path[p] = -1;
                        if (!right) t = __shfl_sync(RCX_FULL, thr_i, spec_lvl + n_ff - 1);
                        kc = advance_kmer(M.P, kc, right, pos, n_ff, k, mask);
                        pos += right ? n_ff : -n_ff;
                        spec_lvl += n_ff;
                        n_nodes += n_ff - 1;
                        DBG_INC(dbg[1]);
                    }
                    if (spec_lite && spec_lvl < spec_n) {  // a level failed the keep-only test: this node needs a full round
                        spec_lvl = -1;
                        force_full = true;
                    }
                    if (n_ff > 0 || spec_lite) {
                        if (n_ff == 0) --n_nodes;  // the same node is entered again
                        __syncwarp();
                        continue;  // enter the node the run stopped at (gate and terminal test are re-evaluated)
                    }
                }
                c0 = __shfl_sync(RCX_FULL, R.c0, spec_lvl);
                c1 = __shfl_sync(RCX_FULL, R.c1, spec_lvl);
                c2 = __shfl_sync(RCX_FULL, R.c2, spec_lvl);
                c3 = __shfl_sync(RCX_FULL, R.c3, spec_lvl);
                {
                    const int r = __shfl_sync(RCX_FULL, R.r, spec_lvl);
                    thr = (t > r || t <= 0) ? r : t;  // InferPosThreshold (:169-172)
                }
                ext = 0;
                stage = ST_KEEP;
                DBG_INC(dbg[2]);
            }
        }
        const int pinfo = stage != ST_DONE ? M.posinfo[pos] : 4;
        const bool here_ok = !(pinfo & 4);
        const int hb = pinfo & 3;
        if (stage == ST_KEEP) {
            stage = ST_ALT0;
            if (here_ok) {  // :294-339 / :532-574
                int c = sel4(hb, c0, c1, c2, c3);
                RcxKmer x = kc;
                if (right) rcx_k_append(x, hb, false, k, mask); else rcx_k_prepend(x, hb, false, k, mask);
                if (c >= thr) {
                    if (lane == 0) M.path[pos] = -1;
                    ext = 1;
                    PUSH_CHILD(right ? pos + 1 : pos - 1, right ? t : thr, fix_cnt, c < bn ? c : bn, x, spec_lvl + 1);
                    continue;
                } else if (thr == 1 && t <= 2) {
                    // "accidentally below the threshold" rescue (:313-338 / :550-573)
                    int i = pos, steps = 0;
                    #pragma unroll 1
                    for (; c < thr && steps < k; ++steps) {
                        if (right) { ++i; if (i >= to) break; } else { --i; if (i < to) break; }
                        const int qi = M.posinfo[i];
                        const bool bad = (qi & 4) != 0;
                        const int b = qi & 3;
                        if (right) rcx_k_append(x, b, bad, k, mask); else rcx_k_prepend(x, b, bad, k, mask);
                        c = x.inv != -1 ? 0 : rcx_probe_ool(tab, rcx_canonical(x.code, k));
                        n_probes += 1;
                        DBG_INC(dbg[5]);
                    }
                    if (steps < k && (right ? i < to : i >= to)) {
                        const int lo = right ? pos : i, hi = right ? i : pos;
                        #pragma unroll 1
                        for (int j = lo + lane; j <= hi; j += 32) M.path[j] = -1;
                        ext = 1;
                        PUSH_CHILD(right ? i + 1 : i - 1, right ? t : thr, fix_cnt + 1, bn, x, -1);
                        continue;
                    }
                }
            }
        }
        if (stage >= ST_ALT0 && stage < ST_JUMP) {  // :343-372 / :577-608
            const bool allowed = (pinfo & (right ? 8 : 16)) != 0;
            bool pushed = false;
            if (allowed) {
                #pragma unroll 1
                for (int b = stage - ST_ALT0; b < 4; ++b) {
                    if (here_ok && b == hb) continue;
                    const int c = sel4(b, c0, c1, c2, c3);
                    if (c >= thr) {
                        if (lane == 0) M.path[pos] = (signed char)b;
                        ++S.trial;
                        ext = 1;
                        RcxKmer x = kc;
                        if (right) rcx_k_append(x, b, false, k, mask); else rcx_k_prepend(x, b, false, k, mask);
                        stage = ST_ALT0 + b + 1;
                        PUSH_CHILD(right ? pos + 1 : pos - 1, right ? t : thr, fix_cnt + (here_ok ? 1 : 0), c < bn ? c : bn, x, -1);
                        pushed = true;
                        break;
                    }
                }
            }
            if (pushed) continue;
            stage = ST_JUMP;
        }
        if (stage == ST_JUMP) {
            stage = ST_DONE;
            if (ext == 0) {  // jump through an unfixable region (:393-441 / :629-677)
                RcxKmer x = kc;
                int i, th = thr;
#ifdef RCX_PHASE_TIMERS
                const long long jt0 = clock64();
                dbg[3] += 1;
#endif
                if (right) {
                    #pragma unroll 1
                    for (i = pos; i < to; ++i) {
                        int d0, d1, d2, d3;
                        probe4(tab, x, true, mask, lane, d0, d1, d2, d3);
                        n_probes += 4;
                        th = infer_thr(d0, d1, d2, d3, t, er);
                        const int qi = M.posinfo[i];
                        const bool bad = (qi & 4) != 0;
                        const int b = qi & 3;
                        rcx_k_append(x, b, bad, k, mask);
                        const int c = bad ? 0 : sel4(b, d0, d1, d2, d3);
                        if (lane == 0) M.path[i] = -1;
                        if (c >= th) break;
                    }
                    int kk = (i < len) ? i - pos - k + 1 : (i - pos) / 2;
                    if (kk <= 0) kk = 1;
                    if (i >= to) i -= 1;
                    TAIL_CHILD(i + 1, t, fix_cnt + kk, x);
                } else {
                    #pragma unroll 1
                    for (i = pos; i >= to; --i) {
                        int d0, d1, d2, d3;
                        probe4(tab, x, false, mask, lane, d0, d1, d2, d3);
                        n_probes += 4;
                        th = infer_thr(d0, d1, d2, d3, t, er);
                        const int qi = M.posinfo[i];
                        const bool bad = (qi & 4) != 0;
                        const int b = qi & 3;
                        rcx_k_prepend(x, b, bad, k, mask);
                        const int c = bad ? 0 : sel4(b, d0, d1, d2, d3);
                        if (lane == 0) M.path[i] = -1;
                        if (c >= th) break;
                    }
                    int kk = (i >= 0) ? pos - i - k + 1 : pos - 1;
                    if (kk <= 0) kk = 1;
                    if (i <= to) ++i;
                    TAIL_CHILD(i - 1, th, fix_cnt + kk, x);
                }
#ifdef RCX_PHASE_TIMERS
                dbg[4] += (unsigned long long)(clock64() - jt0);
#endif
                continue;
            }
        }
        // ST_DONE: return to the parent frame
        if (sp == 0) break;
        --sp;
        spec_lvl = -1;
        {
            const unsigned int* f = M.stack + sp * FRAME_WORDS;
            kc.code = (unsigned long long)f[0] | ((unsigned long long)f[1] << 32);
            const unsigned int w = f[2];
            pos = (int)(w & 0xfff) - 1;
            kc.inv = (int)((w >> 12) & 0xff) - 1;
            stage = (int)((w >> 20) & 0xf);
            ext = (int)((w >> 24) & 1);
            t = (int)f[3]; fix_cnt = (int)f[4]; bn = (int)f[5]; thr = (int)f[6];
            c0 = (int)f[7]; c1 = (int)f[8]; c2 = (int)f[9]; c3 = (int)f[10];
        }
    }
#undef PUSH_CHILD
#undef TAIL_CHILD
}

// window code / state for a seed k-mer built by Restart + k Appends of seq[j .. j+k) (j may be -1)
__device__ __forceinline__ RcxKmer seed_kmer(const RcxPlanes& P, int j, int k, unsigned int kmask32) {
    RcxKmer x;
    if (j >= 0) {
        x.code = rcx_window(P.pk, j, k);
        const unsigned int m = rcx_bits(P.inv, j) & kmask32;
        x.inv = m ? (k - 1) - (31 - __clz(m)) : -1;
    } else {
        // seq[-1] is read by the reference (ErrorCorrection.cpp:1139-1142 with tend = -1); whatever it holds is shifted
        // out / reset by the first Append of the search, so any value gives the same result.
        x.code = rcx_window(P.pk, 0, k) >> 2;
        const unsigned int m = rcx_bits(P.inv, 0) & (kmask32 >> 1);
        x.inv = m ? (k - 2) - (31 - __clz(m)) : -1;
    }
    return x;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
    #pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RCX_FULL, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------ kernels
// The correction is split into three kernels so that each one's hot code stays inside the SM's instruction cache
// (a single fused kernel was 130 KB of SASS and spent 2/3 of its issue slots in no_instruction stalls, profiles/):
//   rcx_triage_kernel   every read : thresholds, "all k-mers trusted?" -> result 0 / -1 and l/m/h at once,
//                                    otherwise the read index goes to a work list
//   rcx_search_kernel   work list  : A.4 iteration loop (islands, segments, ordered DFS, re-threshold)
//   rcx_finalize_kernel work list  : A.7 post-filters, apply, l/m/h re-scan
// State between search and finalize travels through HBM: status + totalFix + badSegmentCnt (12 B / read) and the
// fix[] array (1 B / base, only for reads that reach the post-filters).

__device__ __forceinline__ void setup_warpmem(WarpMem& M, unsigned int* smem, int warp, const CorLayout& L) {
    const int mw = L.mw;
    unsigned int* p = smem + (size_t)warp * L.words;
    M.P.pk = p; p += L.pkw;
    M.P.inv = p; p += mw;
    M.P.mA = p; p += mw;
    M.P.mT = p; p += mw;
    M.P.mN = p; p += mw;
    M.strong = p; p += mw;
    M.island = p; p += mw;
    M.pstrict = p; p += mw;
    M.trusted = p; p += mw;
    M.fixm = p; p += 2 * mw;
    M.counts = reinterpret_cast<int*>(p); p += L.lpad;
    M.path = reinterpret_cast<signed char*>(p);
    M.fix = M.path + L.lpad;
    M.posinfo = reinterpret_cast<unsigned char*>(M.fix + L.lpad);
    p += 3 * L.lpad / 4;
    M.segtop = reinterpret_cast<int*>(p); p += L.nseg;
    M.sbuf = reinterpret_cast<int*>(p);
    M.stack = p;
}

// k-th smallest (0-based) of n ints in shared memory, all >= 1: bisection on the value, no data movement
__device__ __noinline__ int warp_select(const int* v, int n, int kth, int lo, int hi, int lane) {
    #pragma unroll 1
    while (lo < hi) {
        const int mid = lo + (int)(((unsigned int)hi - (unsigned int)lo) >> 1);
        int c = 0;
        #pragma unroll 1
        for (int j = lane; j < n; j += 32) c += v[j] <= mid;
        c = __reduce_add_sync(RCX_FULL, c);
        if (c > kth) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// GetKmerInformation (:1567-1602) from counts[] and the validity plane: l = min, m = element [n/2], h = max of the
// counts of the VALID k-mers with 0 -> 1; all 0 when there is none.
__device__ __noinline__ void kmer_info(const WarpMem& M, int kcnt, unsigned int kmask32, int lane, int& l, int& m, int& h) {
    int nvalid = 0, mn = 0x7fffffff, mx = 0;
    #pragma unroll 1
    for (int j = lane; j < kcnt; j += 32) {
        int v = 0x7fffffff;
        if ((rcx_bits(M.P.inv, j) & kmask32) == 0) {
            v = M.counts[j];
            if (v == 0) v = 1;
            ++nvalid;
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        M.sbuf[j] = v;
    }
    nvalid = __reduce_add_sync(RCX_FULL, nvalid);
    mn = __reduce_min_sync(RCX_FULL, mn);
    mx = __reduce_max_sync(RCX_FULL, mx);
    __syncwarp();
    if (nvalid == 0) { l = m = h = 0; return; }
    l = mn; h = mx;
    m = warp_select(M.sbuf, kcnt, nvalid / 2, mn, mx, lane);  // invalid entries are INT_MAX: never counted below hi
    __syncwarp();
}

// thresholds of A.3 (:772-842) from the scan kernel's per-read summary
__device__ __forceinline__ void initial_thresholds(const RcxCorrectArgs& a, long long r, uint8_t fl, double er, int& strong, int& trust) {
    int pair_t = -1;
    if (a.mode == 1) {
        const int t1 = a.thr[r], t2 = a.thr[r ^ 1];
        pair_t = t1 < t2 ? t1 : t2;
    }
    strong = a.thr[r];
    bool flag = false;
    trust = rcx_bound_i_ool(strong, er);
    if ((fl & RCX_F_DROP) && strong >= 20 && (fl & RCX_F_PREV2) && trust < 3) { flag = true; trust = 3; }
    if (pair_t >= 1 && strong > pair_t) {
        if (!flag || pair_t < 20) trust = rcx_bound_i_ool(pair_t, er);
        strong = pair_t;
    }
    if (trust < 2) trust = 2;
}

__global__ void __launch_bounds__(COR_WARPS * 32) rcx_triage_kernel(RcxCorrectArgs a) {
    extern __shared__ unsigned int smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.t.k;
    const CorLayout L = cor_layout(a.max_len, k);
    WarpMem M;
    setup_warpmem(M, smem, warp, L);
    const unsigned int kmask32 = k >= 32 ? 0xffffffffu : ((1u << k) - 1);
    #pragma unroll 1
    for (long long r = (long long)blockIdx.x * COR_WARPS + warp; r < a.n; r += (long long)gridDim.x * COR_WARPS) {
        const uint32_t o = a.off[r];
        const int len = (int)(a.off[r + 1] - o);
        const int kcnt = len - k + 1;
        if (kcnt <= 0) {  // :713 -> -1 ; GetKmerInformation gives 0,0,0
            if (lane == 0) { a.correction[r] = -1; a.l[r] = 0; a.m[r] = 0; a.h[r] = 0; }
            continue;
        }
        int nN, nA, nT;
        build_planes_ool(M.P, a.seq + o, len, a.max_len, lane, nN, nA, nT);
        #pragma unroll 1
        for (int j = lane; j < kcnt; j += 32) M.counts[j] = a.counts[o + j];
        const uint8_t fl = a.flags[r];
        __syncwarp();
        int result = -2;
        if (fl & RCX_F_REJECT) {
            result = -1;
        } else {
            int strong, trust;
            initial_thresholds(a, r, fl, a.error_rate, strong, trust);
            bool bad = false;  // any untrusted k-mer (:875)?
            #pragma unroll 1
            for (int j = lane; j < kcnt; j += 32) bad |= !(M.counts[j] >= strong && !rcx_polya(M.P, j, k, kmask32, 2));
            if (!__any_sync(RCX_FULL, bad)) result = 0;  // longestTrustedLength == kcnt (:1110)
        }
        if (result == -2) {
            if (lane == 0) a.worklist[atomicAdd(&a.counters[RCX_C_NWORK], 1ull)] = (uint32_t)r;
        } else {
            int l, m, h;
            kmer_info(M, kcnt, kmask32, lane, l, m, h);
            if (lane == 0) { a.correction[r] = result; a.l[r] = l; a.m[r] = m; a.h[r] = h; }
        }
        __syncwarp();
    }
}

// per-work-item state in HBM (int32 words): header, isStrongTrusted plane, strict poly-A plane, segment list
#define ST_STATUS 0      // -2 = go to the post-filters, -1 / 0 = final result, -3 = second pass of A.4 requested
#define ST_TOTALFIX 1
#define ST_BADSEG 2
#define ST_STRONG 3
#define ST_TRUST 4
#define ST_NSEG 5
#define ST_HDR 8
__host__ __device__ inline int state_stride(int max_len, int k) { return ST_HDR + 2 * rcx_mask_words(max_len) + 2 * (max_len / k + 4); }

// work-item fetch shared by the persistent kernels: pass 0 walks the triage work list, pass 1 the (short) list of
// items whose first pass asked for a second one (forceNextRound, ErrorCorrection.cpp:1195-1211; at most two passes
// can happen because forceNextRound needs iter == 0 and `unfixable` is never set)
__device__ __forceinline__ bool fetch_item(const RcxCorrectArgs& a, int counter, int lane, long long& item) {
    const long long n = (long long)a.counters[a.pass == 0 ? RCX_C_NWORK : RCX_C_NWORK2];
    long long w = 0;
    if (lane == 0) w = (long long)atomicAdd(&a.counters[counter], 1ull);
    w = __shfl_sync(RCX_FULL, w, 0);
    if (w >= n) return false;
    item = a.pass == 0 ? w : (long long)a.worklist2[w];
    return true;
}

// A.4 steps 1-4 (:863-1046): trusted k-mers -> islands -> trusted positions -> segments, written to the item state.
__global__ void __launch_bounds__(COR_WARPS * 32) rcx_prep_kernel(RcxCorrectArgs a) {
    extern __shared__ unsigned int smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.t.k;
    const CorLayout L = cor_layout(a.max_len, k);
    const int mw = L.mw;
    WarpMem M;
    setup_warpmem(M, smem, warp, L);
    const unsigned int kmask32 = k >= 32 ? 0xffffffffu : ((1u << k) - 1);
    const int stride = state_stride(a.max_len, k);
    for (;;) {
        long long item;
        if (!fetch_item(a, a.pass == 0 ? RCX_C_WORKX0 : RCX_C_WORKX2, lane, item)) break;
        const long long r = a.worklist[item];
        int* st = a.state + item * stride;
        const uint32_t o = a.off[r];
        const int len = (int)(a.off[r + 1] - o);
        const int kcnt = len - k + 1;
        int nN, nA, nT;
        build_planes_ool(M.P, a.seq + o, len, a.max_len, lane, nN, nA, nT);
        #pragma unroll 1
        for (int j = lane; j < kcnt; j += 32) M.counts[j] = a.counts[o + j];
        int strong, trust;
        if (a.pass == 0) initial_thresholds(a, r, a.flags[r], a.error_rate, strong, trust);
        else { strong = st[ST_STRONG]; trust = st[ST_TRUST]; }
        #pragma unroll 1
        for (int c = lane; c < mw; c += 32) { M.pstrict[c] = 0; M.strong[c] = 0; M.trusted[c] = 0; }
        __syncwarp();
        #pragma unroll 1
        for (int c = 0; (c << 5) < kcnt; ++c) {
            const int j = (c << 5) + lane;
            const bool pa = j < kcnt && rcx_polya(M.P, j, k, kmask32, 2);
            const unsigned int bp = __ballot_sync(RCX_FULL, pa);
            const unsigned int bt = __ballot_sync(RCX_FULL, j < kcnt && M.counts[j] >= strong && !pa);  // :875
            if (lane == 0) { M.pstrict[c] = bp; M.trusted[c] = bt; }
        }
        __syncwarp();
        int status = -2, nseg = 0;
        if (next_bit(M.trusted, 0, kcnt, false, lane) >= kcnt) {
            status = 0;  // longestTrustedLength == kcnt (:1110)
        } else {
            // runs of trusted k-mers -> longest run, islands (runs >= 2), boundary adjustment, trusted positions
            int longest = 0, tstart = 0, tend = -1;
            {
                int pf = 0, pt = -1;
                bool have = false;
                int p = 0;
                #pragma unroll 1
                while (p < kcnt) {
                    const int s = next_bit(M.trusted, p, kcnt, true, lane);
                    if (s >= kcnt) break;
                    const int e = next_bit(M.trusted, s, kcnt, false, lane);
                    const int rl = e - s;
                    if (rl > longest) { longest = rl; tstart = s; tend = e - 1; }
                    if (rl >= 2) {
                        int cf = s, ct = e - 1;
                        if (have) {
                            if (cf <= pt + k) {  // :936-964
                                const int len1 = pt - pf, len2 = ct - cf, overlap = pt + k - cf;
                                const int j = pt + 1 + lane;
                                const bool low = j < cf && M.counts[j] <= 2 && M.counts[j] < trust;
                                if (__any_sync(RCX_FULL, low) && overlap <= 3) {
                                    if (len1 < len2) pt -= overlap + 1; else cf += overlap + 1;
                                }
                            }
                            if (pf <= pt) set_range(M.strong, pf, pt + k - 1, mw, lane);
                            __syncwarp();
                        }
                        pf = cf; pt = ct; have = true;
                    }
                    p = e;
                }
                if (have && pf <= pt) set_range(M.strong, pf, pt + k - 1, mw, lane);
                __syncwarp();
            }
            // islands that define the segments (:977-1007)
            #pragma unroll 1
            for (int c = lane; c < mw; c += 32) M.island[c] = M.strong[c];
            __syncwarp();
            if (next_bit(M.strong, 0, len, true, lane) >= len) {
                set_range(M.island, tstart, tend + k - 1, mw, lane);
                __syncwarp();
            }
            // segments (:1009-1046): prefix, gaps between islands, suffix; packed (from | to << 16), (lanchor | ranchor << 16)
            int* segs = st + ST_HDR + 2 * mw;
            int ifrom = next_bit(M.island, 0, len, true, lane);
            int ito = ifrom < len ? next_bit(M.island, ifrom, len, false, lane) - 1 : len - 1;
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
                    if (lane == 0 && nseg < L.nseg) { segs[2 * nseg] = sf | (sto << 16); segs[2 * nseg + 1] = la | (ra << 16); }
                    ++nseg;
                }
                if (ifrom >= len) break;
                prev_to = ito; prev_len = ito - ifrom + 1; first = false;
                ifrom = next_bit(M.island, ito + 1, len, true, lane);
                ito = ifrom < len ? next_bit(M.island, ifrom, len, false, lane) - 1 : len - 1;
            }
        }
        #pragma unroll 1
        for (int c = lane; c < mw; c += 32) { st[ST_HDR + c] = (int)M.strong[c]; st[ST_HDR + mw + c] = (int)M.pstrict[c]; }
        if (lane == 0) {
            st[ST_STATUS] = status; st[ST_TOTALFIX] = 0; st[ST_BADSEG] = 0;
            st[ST_STRONG] = strong; st[ST_TRUST] = trust; st[ST_NSEG] = nseg < L.nseg ? nseg : L.nseg;
        }
        __syncwarp();
    }
}

// A.4 steps 6-9 (:1113-1289): the ordered searches over the segments of one read, then the ambiguity / forceNextRound
// decisions and, when a second pass is due, the re-threshold.
__global__ void __launch_bounds__(COR_WARPS * 32, 4) rcx_search_kernel(RcxCorrectArgs a) {
    extern __shared__ unsigned int smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.t.k, K = a.max_fix_per_k;
    const double er = a.error_rate;
    const CorLayout L = cor_layout(a.max_len, k);
    const int mw = L.mw;
    WarpMem M;
    setup_warpmem(M, smem, warp, L);
    const unsigned int kmask32 = k >= 32 ? 0xffffffffu : ((1u << k) - 1);
    const int stride = state_stride(a.max_len, k);
    unsigned long long n_nodes = 0, n_sprobes = 0;
    unsigned long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#ifdef RCX_PHASE_TIMERS
    unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_last = clock64();
    int ph_cur = 0;
#endif
    for (;;) {
        PH(0);  // work fetch
        long long item;
        if (!fetch_item(a, a.pass == 0 ? RCX_C_WORKX1 : RCX_C_WORKX3, lane, item)) break;
        int* st = a.state + item * stride;
        if (st[ST_STATUS] != -2) continue;  // decided by the prep kernel (all k-mers trusted on the second pass)
        const long long r = a.worklist[item];
        const uint32_t o = a.off[r];
        const int len = (int)(a.off[r + 1] - o);
        const int kcnt = len - k + 1;
        PH(1);  // planes + state loads
        int nN, nA, nT;
        build_planes_ool(M.P, a.seq + o, len, a.max_len, lane, nN, nA, nT);
        #pragma unroll 1
        for (int c = lane; c < mw; c += 32) { M.strong[c] = (unsigned int)st[ST_HDR + c]; M.pstrict[c] = (unsigned int)st[ST_HDR + mw + c]; }
        __syncwarp();
        #pragma unroll 1
        for (int i = lane; i < len; i += 32) {
            M.fix[i] = -1; M.path[i] = -1;  // :868-869, :1113-1114
            const bool free_pos = !getbit(M.strong, i);
            const bool pr = i - k + 1 >= 0 && getbit(M.pstrict, i - k + 1), pl = getbit(M.pstrict, i);  // :343 / :577
            M.posinfo[i] = (unsigned char)(rcx_base(M.P.pk, i) | (getbit(M.P.inv, i) << 2) | ((free_pos && !pr) ? 8 : 0) |
                                           ((free_pos && !pl) ? 16 : 0));
        }
        const int strong = st[ST_STRONG], trust = st[ST_TRUST], nseg = st[ST_NSEG];
        const int* segs = st + ST_HDR + 2 * mw;
        __syncwarp();
        PH(3);  // searches
        int result = -2, total_fix = 0, bad_segments = 0, best_bn = RCX_INF;
        bool abort_read = false;
        #pragma unroll 1
        for (int g = 0; g < nseg && !abort_read; ++g) {  // :1125-1175
            const int w0 = segs[2 * g], w1 = segs[2 * g + 1];
            const int sf = w0 & 0xffff, sto = w0 >> 16, la = w1 & 0xffff, ra = w1 >> 16;
            SearchState S;
            S.trial = 0;
            S.max_fix = (sto - sf + 1) * K / k * 2 + 1;
            if (S.max_fix < K) S.max_fix = K;
            S.best_bn = -1; S.best_cnt = -1; S.top2a = -1; S.top2b = -1;
            {
                const bool right = la >= ra;  // :1136
                const int extend = right ? (sto == len ? 0 : k - 1) : (sf == 0 ? 0 : k - 1);
                const RcxKmer kc = seed_kmer(M.P, right ? sf - k : sto + 1, k, kmask32);
                search_segment(a.t, M, S, right, right ? sf : sto, right ? sto + extend : sf - extend, trust, kc, len, er, lane,
                               n_nodes, n_sprobes, dbg);
            }
            if (lane == 0) M.segtop[g] = S.top2b;
            if (S.best_bn == -1) {
                ++bad_segments;
            } else {
                if (S.best_bn < best_bn) best_bn = S.best_bn;
                if (S.trial > RCX_MAX_TRIAL) abort_read = true;  // :1172
                else total_fix += S.max_fix;
            }
        }
        __syncwarp();
        PH(4);  // decisions, re-threshold
        bool second_pass = false;
        int new_strong = strong, new_trust = trust;
        if (abort_read) {
            result = -1;
        } else {
            bool force_next = false;
            if (nseg > 0) {  // :1176-1230
                int best_fix_cnt = 1;
                #pragma unroll 1
                for (int g = 0; g < nseg; ++g)
                    if (M.segtop[g] >= best_bn) best_fix_cnt *= 2;
                if (a.pass == 0 && (double)best_bn < rcx_bound_d_ool(strong, er)) force_next = true;  // :1195 FP64 compare
                if (best_fix_cnt >= 2) result = -1;
            }
            if (result == -2) {
                const bool unfixable = total_fix > len;  // :1234 (never true in practice)
                if (total_fix == 0 && force_next) result = 0;  // :1231
                else if (!unfixable && !force_next) { /* leave the loop: post-filters */ }
                else if (trust < 10 && !force_next) result = -1;
                else {
                    // ---------------- re-threshold (:1247-1289)
                    const int loose = (k / 2 > 7) ? k / 2 : 7;
                    const int n2 = rcx_pow2_at_least(kcnt);
                    #pragma unroll 1
                    for (int j = lane; j < n2; j += 32)
                        M.sbuf[j] = j < kcnt ? (rcx_polya(M.P, j, k, kmask32, loose) ? -1 : a.counts[o + j]) : 0x7fffffff;
                    __syncwarp();
                    warp_sort_ool(M.sbuf, n2, lane);
                    int ibreak = 0;
                    bool any_drop = false;
                    #pragma unroll 1
                    for (int i = 1 + lane; i < kcnt; i += 32) {
                        const int v = M.sbuf[i], pv = M.sbuf[i - 1];
                        if (v <= strong && ((v > 2 * pv && v > 10) || (pv == 0 && v >= 5)) && v < strong) ibreak = i;
                    }
                    ibreak = __reduce_max_sync(RCX_FULL, ibreak);
                    #pragma unroll 1
                    for (int i = 1 + lane; i < kcnt; i += 32) {
                        const int v = M.sbuf[i], pv = M.sbuf[i - 1];
                        if (i >= ibreak && v <= strong && ((v > 2 * pv && v > 10) || (pv == 0 && v >= 5))) any_drop = true;
                    }
                    any_drop = __any_sync(RCX_FULL, any_drop);
                    if (any_drop) {
                        const int nv = M.sbuf[ibreak];
                        new_trust = rcx_bound_i_ool(nv, er);
                        new_strong = nv;
                        second_pass = a.pass == 0;  // a second pass never asks for a third (forceNextRound needs iter == 0)
                    }
                    __syncwarp();
                }
            }
        }
        PH(5);  // hand-over
        if (second_pass) {
            if (lane == 0) {
                st[ST_STATUS] = -3; st[ST_STRONG] = new_strong; st[ST_TRUST] = new_trust;
                a.worklist2[atomicAdd(&a.counters[RCX_C_NWORK2], 1ull)] = (uint32_t)item;
            }
        } else {
            if (result == -2) {
                #pragma unroll 1
                for (int i = lane; i < len; i += 32) a.fixbuf[o + i] = M.fix[i];
            }
            if (lane == 0) { st[ST_STATUS] = result; st[ST_TOTALFIX] = total_fix; st[ST_BADSEG] = bad_segments; }
        }
        __syncwarp();
    }
#ifdef RCX_PHASE_TIMERS
    PH(7);
    if (lane == 0) {
        for (int i = 0; i < 8; ++i) atomicAdd(&a.counters[RCX_C_PHASE0 + i], ph[i]);
        atomicAdd(&a.counters[RCX_C_ROUNDS], dbg[0]);
        atomicAdd(&a.counters[RCX_C_FF], dbg[1]);
        atomicAdd(&a.counters[RCX_C_REGULAR], dbg[2]);
        for (int i = 3; i < 8; ++i) atomicAdd(&a.counters[RCX_C_PHASE0 + 8 + i - 3], dbg[i]);
    }
#endif
    n_nodes = warp_sum_u64(lane == 0 ? n_nodes : 0);
    n_sprobes = warp_sum_u64(lane == 0 ? n_sprobes : 0);
    if (lane == 0) {
        if (n_nodes) atomicAdd(&a.counters[RCX_C_SEARCH_NODES], n_nodes);
        if (n_sprobes) atomicAdd(&a.counters[RCX_C_SEARCH_PROBES], n_sprobes);
    }
}

__global__ void __launch_bounds__(COR_WARPS * 32) rcx_finalize_kernel(RcxCorrectArgs a) {
    extern __shared__ unsigned int smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.t.k, K = a.max_fix_per_k, Q = a.bad_qual;
    const CorLayout L = cor_layout(a.max_len, k);
    const int mw = L.mw;
    WarpMem M;
    setup_warpmem(M, smem, warp, L);
    const unsigned int kmask32 = k >= 32 ? 0xffffffffu : ((1u << k) - 1);
    unsigned long long n_rprobes = 0;
    const long long nwork = (long long)a.counters[RCX_C_NWORK];
    for (;;) {
        long long w = 0;
        if (lane == 0) w = (long long)atomicAdd(&a.counters[RCX_C_WORK2], 1ull);
        w = __shfl_sync(RCX_FULL, w, 0);
        if (w >= nwork) break;
        const long long r = a.worklist[w];
        const uint32_t o = a.off[r];
        const int len = (int)(a.off[r + 1] - o);
        const int kcnt = len - k + 1;
        const int* st = a.state + w * state_stride(a.max_len, k);
        int result = st[ST_STATUS];
        const int total_fix = st[ST_TOTALFIX], bad_segments = st[ST_BADSEG];
        bool changed = false;
        int nN = 0, nA = 0, nT = 0;
        build_planes_ool(M.P, a.seq + o, len, a.max_len, lane, nN, nA, nT);
        #pragma unroll 1
        for (int j = lane; j < kcnt; j += 32) M.counts[j] = a.counts[o + j];
        if (result == -2) {
            #pragma unroll 1
            for (int i = lane; i < len; i += 32) M.fix[i] = a.fixbuf[o + i];
        }
        __syncwarp();
        {
            if (result == -2) {
                // ---------------- A.7 post-filters (:1296-1466)
                const char* q = a.qual + o;
                const int q0 = q[0];
                int* P = M.sbuf;  // ascending positions with a fix on a non-N base
                int np = 0;
                #pragma unroll 1
                for (int c = 0; (c << 5) < len; ++c) {
                    const int i = (c << 5) + lane;
                    const bool f = i < len && M.fix[i] != -1 && !getbit(M.P.mN, i);
                    const unsigned int b = __ballot_sync(RCX_FULL, f);
                    if (f) P[np + __popc(b & ((1u << lane) - 1))] = i;
                    np += __popc(b);
                }
                __syncwarp();
                #pragma unroll 1
                for (int i = 1; i < np; ++i) {  // pair filter (:1314-1398)
                    const int pa = P[i - 1], pb = P[i];
                    if (q0 != 0 && q[pb] <= Q && q[pa] <= Q) continue;
                    if (pb - pa + 1 > k) continue;
                    const int lim = (i < np - 1) ? P[i + 1] - k + 1 : 0x7fffffff;  // j + k - 1 >= P[i+1]  <=>  j >= lim
                    int j0 = pa - k + 1;
                    if (j0 < 0) j0 = 0;
                    // loop 1: [j0, e1), skipping j <= P[i-2]
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
                    #pragma unroll 1
                    for (int j = j0 + lane; j < e3; j += 32) {
                        const int c = M.counts[j];
                        if (j < e1) { if (j > skip && c < ms) ms = c; }
                        else if (j < e2) { if (c < md) md = c; }
                        else { if (c < ms) ms = c; }
                    }
                    ms = __reduce_min_sync(RCX_FULL, ms);
                    md = __reduce_min_sync(RCX_FULL, md);
                    if (ms != RCX_INF && md != RCX_INF && ms > 1 && md > 1 && ms > md / 2 && ms < 2 * md) {
                        if (lane == 0) { M.fix[pb] = -1; M.fix[pa] = -1; }
                        int j = i - 2;
                        #pragma unroll 1
                        while (j >= 0 && P[j + 1] - P[j] + 1 <= k) { if (lane == 0) M.fix[P[j]] = -1; --j; }
                        #pragma unroll 1
                        while (i + 1 < np && P[i + 1] - P[i] + 1 <= k) { if (lane == 0) M.fix[P[i + 1]] = -1; ++i; }
                    }
                }
                __syncwarp();
                if (total_fix > 3 && len > 10) {  // edge filter (:1407-1430)
                    {
                        const int i = lane;
                        const bool nn = i < 10 && !getbit(M.P.mN, i);
                        const bool f = nn && M.fix[i] != -1 && q[i] > Q;
                        if (__popc(__ballot_sync(RCX_FULL, f)) >= 2 && nn) M.fix[i] = -1;
                    }
                    __syncwarp();
                    {
                        const int i = len - 10 + lane;
                        const bool nn = lane < 10 && !getbit(M.P.mN, i);
                        const bool f = nn && M.fix[i] != -1 && q[i] > Q;
                        if (__popc(__ballot_sync(RCX_FULL, f)) >= 3 && nn) M.fix[i] = -1;
                    }
                    __syncwarp();
                }
                if (total_fix >= K) {  // density filter (:1432-1455), weights doubled to stay in integers
                    #pragma unroll 1
                    for (int c = lane; c < 2 * mw; c += 32) M.fixm[c] = 0;
                    __syncwarp();
                    #pragma unroll 1
                    for (int c = 0; (c << 5) < len; ++c) {
                        const int i = (c << 5) + lane;
                        const bool f = i < len && M.fix[i] != -1 && !getbit(M.P.mN, i);
                        const bool good = f && q[i] > Q;
                        const unsigned int bg = __ballot_sync(RCX_FULL, good), bb = __ballot_sync(RCX_FULL, f && !good);
                        if (lane == 0) { M.fixm[c] = bg; M.fixm[mw + c] = bb; }
                    }
                    __syncwarp();
                    bool dense = false;
                    #pragma unroll 1
                    for (int j = lane; j < kcnt; j += 32) {
                        const int w2 = 2 * __popc(rcx_bits(M.fixm, j) & kmask32) + __popc(rcx_bits(M.fixm + mw, j) & kmask32);
                        if (w2 > 2 * K) dense = true;
                    }
                    if (__any_sync(RCX_FULL, dense)) result = -1;
                }
                if (result == -2) {  // apply (:1468-1479)
                    int ret = 0;
                    #pragma unroll 1
                    for (int c = 0; (c << 5) < len; ++c) {
                        const int i = (c << 5) + lane;
                        const bool f = i < len && M.fix[i] != -1;
                        if (f) a.seq[o + i] = "ACGT"[M.fix[i]];
                        ret += __popc(__ballot_sync(RCX_FULL, f));
                    }
                    if (ret == 0 && bad_segments > 0) result = -1;
                    else { result = ret; changed = ret > 0; }
                }
            }
        }
        // ---------------- GetKmerInformation on the final sequence (:1567-1602)
        __syncwarp();
        if (changed) {
            // fold the fixes into the bit planes, then re-probe only the windows that cover a fix
            #pragma unroll 1
            for (int c = lane; c < mw; c += 32) M.fixm[c] = 0;
            __syncwarp();
            #pragma unroll 1
            for (int c = 0; (c << 5) < len; ++c) {
                const int i = (c << 5) + lane;
                const bool f = i < len && M.fix[i] != -1;
                const unsigned int bf = __ballot_sync(RCX_FULL, f);
                const unsigned int sh = (unsigned int)(f ? M.fix[i] : 0) << (30 - 2 * (lane & 15));
                const unsigned int keep = f ? ~(3u << (30 - 2 * (lane & 15))) : 0xffffffffu;
                const unsigned int hi = __reduce_or_sync(RCX_FULL, lane < 16 ? sh : 0u), lo = __reduce_or_sync(RCX_FULL, lane >= 16 ? sh : 0u);
                const unsigned int khi = __reduce_and_sync(RCX_FULL, lane < 16 ? keep : 0xffffffffu);
                const unsigned int klo = __reduce_and_sync(RCX_FULL, lane >= 16 ? keep : 0xffffffffu);
                if (lane == 0) {
                    M.fixm[c] = bf;
                    M.P.inv[c] &= ~bf;
                    M.P.pk[2 * c] = (M.P.pk[2 * c] & khi) | hi;
                    M.P.pk[2 * c + 1] = (M.P.pk[2 * c + 1] & klo) | lo;
                }
            }
            __syncwarp();
            #pragma unroll 1
            for (int j = lane; j < kcnt; j += 32) {
                if ((rcx_bits(M.fixm, j) & kmask32) == 0) continue;
                int c = 0;
                if ((rcx_bits(M.P.inv, j) & kmask32) == 0) {
                    c = rcx_probe_ool(a.t, rcx_canonical(rcx_window(M.P.pk, j, k), k));
                    ++n_rprobes;
                }
                M.counts[j] = c;
            }
            __syncwarp();
        }
        {
            int l, m, h;
            kmer_info(M, kcnt, kmask32, lane, l, m, h);
            if (lane == 0) { a.correction[r] = result; a.l[r] = l; a.m[r] = m; a.h[r] = h; }
            __syncwarp();
        }
    }
    n_rprobes = warp_sum_u64(n_rprobes);
    if (lane == 0 && n_rprobes) atomicAdd(&a.counters[RCX_C_RESCAN_PROBES], n_rprobes);
}

static cudaError_t launch_one(void (*kern)(RcxCorrectArgs), const RcxCorrectArgs& a, int sm_count, long long items, size_t smem,
                              cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, COR_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long want = (items + COR_WARPS - 1) / COR_WARPS;
    long long grid = (long long)sm_count * per_sm;  // persistent grid: a multiple of the SM count
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    kern<<<(unsigned int)grid, COR_WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
}

size_t rcx_state_words(int max_len, int k) { return (size_t)state_stride(max_len, k); }

cudaError_t rcx_launch_correct(const RcxCorrectArgs& a0, int sm_count, cudaStream_t s) {
    if (a0.n <= 0) return cudaSuccess;
    RcxCorrectArgs a = a0;
    const size_t smem = rcx_correct_smem_per_warp(a.max_len, a.t.k) * COR_WARPS;
    a.pass = 0;
    cudaError_t e = launch_one(rcx_triage_kernel, a, sm_count, a.n, smem, s);
    if (e != cudaSuccess) return e;
    // the work-list lengths live on the device: the persistent kernels read them there (no host round trip)
    for (int pass = 0; pass < 2; ++pass) {
        a.pass = pass;
        e = launch_one(rcx_prep_kernel, a, sm_count, pass == 0 ? a.n : a.n / 8 + 1, smem, s);
        if (e != cudaSuccess) return e;
        e = launch_one(rcx_search_kernel, a, sm_count, pass == 0 ? a.n : a.n / 8 + 1, smem, s);
        if (e != cudaSuccess) return e;
    }
    return launch_one(rcx_finalize_kernel, a, sm_count, a.n, smem, s);
}
