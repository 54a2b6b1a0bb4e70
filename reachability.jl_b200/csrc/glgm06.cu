// GLGM06 zonotope post-operator on the device.
//
// Replaces `reach_homog!` / `reach_inhomog!` (src/ReachSets/ContinuousPost/GLGM06/reach.jl:5-62) and the LazySets
// calls they make (linear_map, minkowski_sum, reduce_order -- SURVEY.md a8-a11, Appendix A.4/B.1-B.3):
//     R_k = reduce_order( (P_{k-1} c0 + cW, [P_{k-1} G0 | GW]), r )                                   (:48-49)
//     W   = reduce_order( (cW + P_{k-1} cV, [GW | P_{k-1} GV]), r )                                   (:54-55)
//     P_k = P_{k-1} Phi                                                                               (:57-58)
//
// B200 design
//  * The mapped generators do not depend on W, so they are produced for a whole BLOCK of s steps by one large DMMA
//    GEMM: with X = [G0 | GV | c0 | cV] (n x q) and Y_j = Phi^j X, the stack [Y_{bs+1} .. Y_{bs+s}] is Phi^s times the
//    previous stack (left multiplication; gemm_left keeps everything column-major, one generator = one contiguous
//    column).  Their Girard scores h = |g|_1 - |g|_inf, the within-segment stable order and the zero-generator flags
//    are batched per block as well (score_stack / sort_segments): off the sequential critical path.
//  * Per step only the data-dependent part remains.  W is kept as dense columns plus per-column descriptors
//    (score h, sorted rank, and an axis tag for box generators so that diag(d) is never stored or read densely).
//    Because every segment carries its stable sorted order, the sortperm of a concatenation is a MERGE:
//    rank = own rank + binary search in the other segment (rank_lists, O(p log p), ties resolved by list index as
//    Julia's stable sortperm does).  apply_step then moves each kept column once (warp per column, 16-byte
//    vectorised, coalesced) and accumulates |g| of the discarded ones in a fixed order (deterministic, no float
//    atomics); the last box CTA finalises descriptors, centres and counts for the next step.
//  * All counts live on the device (zero-generator removal makes them data dependent): the host enqueues the whole
//    flowpipe without ever synchronising.
#include "common.cuh"
#include "../../include/reach.h"
#include <algorithm>
#include <cmath>

namespace reach {

namespace {

constexpr int MODE_UNCHANGED = 0;   // r*n >= p: reduce_order returns Z unchanged
constexpr int MODE_REDUCE = 1;
constexpr int NBOX_R = 24;          // CTAs accumulating the discarded generators of the R / W reduction
constexpr int NBOX_W = 12;
constexpr int NBOX = NBOX_R + NBOX_W;
constexpr int APPLY_THREADS = 256;
constexpr int MAX_IT = 16;          // double2 per lane: rows <= 32*2*16 = 1024

struct WState {      // ping-pong by step parity
    int pW;          // generators of W
    int cur;         // which GW buffer holds them
    int pad0, pad1;
};

struct StepPlan {    // written by rank_lists, read by apply_step; kept for every step (selection records)
    int modeR, pinR, mR, pA;
    int modeW, pinW, mW, pC;
};

struct GP {
    int n, ld, p0, pV, q;
    int strip;
    int64_t nkeep, rn;      // n*(r-1), r*n
    int capW, capR, capListR, capListW;
    // mapped stack of the current block and its batched metadata
    const double* Y;
    double* hS; int* rkS; int* cposS; double* skS; int* cntS;
    // W: dense columns + descriptors
    double* GW[2];
    double* hW[2]; int* tagW[2]; double* valW[2]; int* rkW[2]; double* skW[2];
    WState* st;
    double* cW;
    int* orderR; int* orderW; int64_t order_stride;   // per-step stride (0 when selections are not recorded)
    double* recH; int64_t rec_stride;                  // per-step scores of both lists (NULL unless recorded)
    StepPlan* plans;
    // flowpipe
    double* Rg; int64_t slot;   // doubles per step slot (ld * capR)
    double* Rc; double* Rd; int* Rrow; int2* meta;
    double* partial; unsigned* ticket;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Girard score of one column: lanes stride the rows (double2), fixed order => a pure function of the column's bits.
__device__ __forceinline__ void column_score(const double* __restrict__ col, int n2, int lane, double& asum, double& amax) {
    const double2* c2 = reinterpret_cast<const double2*>(col);
    double s = 0.0, m = 0.0;
    for (int i = lane; i < n2; i += 32) {
        const double2 v = c2[i];
        const double ax = fabs(v.x), ay = fabs(v.y);
        s += ax;
        s += ay;
        m = fmax(m, fmax(ax, ay));
    }
    asum = warp_sum(s);
    amax = warp_max(m);
}

// h for every mapped generator of a block (columns c < p0 + pV of each of the cnt steps); dropped (all-zero, strip
// mode) columns get +inf.  One warp per column.
__global__ void __launch_bounds__(256) score_stack_kernel(GP g, int cnt) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int per = g.p0 + g.pV;
    if (warp >= cnt * per) return;
    const int t = warp / per, c = warp % per;
    const double* col = g.Y + (int64_t)(t * g.q + c) * g.ld;
    double asum, amax;
    column_score(col, g.ld / 2, lane, asum, amax);
    if (lane == 0) g.hS[t * g.q + c] = (g.strip && asum == 0.0) ? INFINITY : asum - amax;
}

// Stable order of every mapped segment (step t, seg 0 = P G0, seg 1 = P GV): rank by counting, O(len^2) but batched
// over the block and spread over (segment, 256-element chunk) CTAs.
__global__ void __launch_bounds__(256) sort_segments_kernel(GP g) {
    __shared__ double keys[1024];
    const int t = blockIdx.y >> 1, seg = blockIdx.y & 1;
    const int len = seg ? g.pV : g.p0;
    if (blockIdx.x * 256 >= len && !(blockIdx.x == 0)) return;
    const int base = t * g.q + (seg ? g.p0 : 0);
    const int e = blockIdx.x * 256 + threadIdx.x;
    const double he = e < len ? g.hS[base + e] : INFINITY;
    int less = 0, cpos = 0, nvalid = 0;
    for (int j0 = 0; j0 < len; j0 += 1024) {
        const int tile = min(1024, len - j0);
        __syncthreads();
        for (int j = threadIdx.x; j < tile; j += 256) keys[j] = g.hS[base + j0 + j];
        __syncthreads();
        for (int j = 0; j < tile; ++j) {
            const double hj = keys[j];
            const int gj = j0 + j;
            const bool valid = hj < INFINITY;
            less += (hj < he) || (hj == he && valid && gj < e);
            cpos += valid && gj < e;
            nvalid += valid;
        }
    }
    if (e < len) {
        const bool valid = he < INFINITY;
        g.rkS[base + e] = valid ? less : -1;
        g.cposS[base + e] = cpos;
        if (valid) g.skS[base + less] = he;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) g.cntS[t * 2 + seg] = nvalid;
}

__device__ __forceinline__ int lower_bound(const double* a, int len, double key) {   // # entries < key
    int lo = 0, hi = len;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int upper_bound(const double* a, int len, double key) {   // # entries <= key
    int lo = 0, hi = len;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// sortperm of the two concatenations of step k by merging sorted segments.
//   R list = [A = P G0 (mapped, step t of the block) | B = W]        (GLGM06/reach.jl:48)
//   W list = [B = W | C = P GV]                                       (GLGM06/reach.jl:54)
// order?[sigma] = source code: c < p0+pV -> mapped column c of the step; p0+pV+j -> column j of W.
__global__ void __launch_bounds__(256) rank_lists_kernel(GP g, int t, int k, int use_smem) {
    extern __shared__ double sk_s[];
    const WState st = g.st[k & 1];
    const int par = k & 1;
    const int pW = st.pW, pA = g.cntS[t * 2 + 0], pC = g.cntS[t * 2 + 1];
    const int baseA = t * g.q, baseC = t * g.q + g.p0;
    const double* skA = g.skS + baseA;
    const double* skB = g.skW[par];
    const double* skC = g.skS + baseC;
    if (use_smem) {
        double* sA = sk_s;
        double* sB = sA + pA;
        double* sC = sB + pW;
        for (int i = threadIdx.x; i < pA; i += blockDim.x) sA[i] = skA[i];
        for (int i = threadIdx.x; i < pW; i += blockDim.x) sB[i] = skB[i];
        for (int i = threadIdx.x; i < pC; i += blockDim.x) sC[i] = skC[i];
        __syncthreads();
        skA = sA; skB = sB; skC = sC;
    }
    const int pinR = pA + pW, pinW = pW + pC;
    const int modeR = (g.rn >= (int64_t)pinR) ? MODE_UNCHANGED : MODE_REDUCE;
    const int modeW = (g.rn >= (int64_t)pinW) ? MODE_UNCHANGED : MODE_REDUCE;
    int* orderR = g.orderR + (int64_t)k * g.order_stride;
    int* orderW = g.orderW + (int64_t)k * g.order_stride;
    double* recH = g.recH ? g.recH + (int64_t)k * g.rec_stride : nullptr;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nmap = g.p0 + g.pV;
    if (gid == 0) {
        StepPlan pl;
        pl.modeR = modeR; pl.pinR = pinR; pl.mR = modeR == MODE_REDUCE ? (int)(pinR - g.nkeep) : 0; pl.pA = pA;
        pl.modeW = modeW; pl.pinW = pinW; pl.mW = modeW == MODE_REDUCE ? (int)(pinW - g.nkeep) : 0; pl.pC = pC;
        g.plans[k] = pl;
    }
    int e = gid;
    if (e < g.p0) {                                   // R list, segment A
        const double h = g.hS[baseA + e];
        if (recH) recH[e] = h;
        if (h < INFINITY) orderR[g.rkS[baseA + e] + lower_bound(skB, pW, h)] = e;
        return;
    }
    e -= g.p0;
    if (e < g.capW) {                                 // R list, segment B
        if (e < pW) {
            const double h = g.hW[par][e];
            if (recH) recH[g.p0 + e] = h;
            orderR[g.rkW[par][e] + upper_bound(skA, pA, h)] = nmap + e;
        }
        return;
    }
    e -= g.capW;
    if (e < g.capW) {                                 // W list, segment B
        if (e < pW) {
            const double h = g.hW[par][e];
            if (recH) recH[g.capListR + e] = h;
            orderW[g.rkW[par][e] + lower_bound(skC, pC, h)] = nmap + e;
        }
        return;
    }
    e -= g.capW;
    if (e < g.pV) {                                   // W list, segment C
        const double h = g.hS[baseC + e];
        if (recH) recH[g.capListR + pW + e] = h;
        if (h < INFINITY) orderW[g.rkS[baseC + e] + upper_bound(skB, pW, h)] = g.p0 + e;
    }
}

struct Src {
    const double* col;   // dense column (nullptr when tagged)
    int tag;             // axis row of a box generator, -1 dense
    double val;
    double h;
};

__device__ __forceinline__ Src resolve(const GP& g, int code, int t, int par, int cur) {
    Src s;
    const int nmap = g.p0 + g.pV;
    if (code < nmap) {
        s.col = g.Y + (int64_t)(t * g.q + code) * g.ld;
        s.tag = -1; s.val = 0.0;
        s.h = g.hS[t * g.q + code];
    } else {
        const int j = code - nmap;
        s.tag = g.tagW[par][j];
        s.val = g.valW[par][j];
        s.h = g.hW[par][j];
        s.col = s.tag >= 0 ? nullptr : g.GW[cur] + (int64_t)j * g.ld;
    }
    return s;
}

__device__ __forceinline__ void copy_column(const GP& g, const Src& s, double* __restrict__ dst, int lane) {
    double2* d2 = reinterpret_cast<double2*>(dst);
    const int n2 = g.ld / 2;
    if (s.tag < 0) {
        const double2* c2 = reinterpret_cast<const double2*>(s.col);
#pragma unroll 4
        for (int i = lane; i < n2; i += 32) d2[i] = c2[i];
    } else {
        for (int i = lane; i < n2; i += 32) {
            double2 v = make_double2(0.0, 0.0);
            if (2 * i == s.tag) v.x = s.val;
            if (2 * i + 1 == s.tag) v.y = s.val;
            d2[i] = v;
        }
    }
}

__device__ void finalize_step(const GP& g, int t, int k);

// One launch per step: box CTAs (blockIdx.x < NBOX) accumulate the discarded generators, the others copy the kept
// ones, warp per column.
__global__ void __launch_bounds__(APPLY_THREADS) apply_step_kernel(GP g, int t, int k) {
    extern __shared__ double red_s[];       // [APPLY_THREADS/32][ld] cross-warp reduction of the box CTAs
    __shared__ unsigned last_flag;
    const StepPlan pl = g.plans[k];
    const int par = k & 1;
    const WState st = g.st[par];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n2 = g.ld / 2;
    const int* orderR = g.orderR + (int64_t)k * g.order_stride;
    const int* orderW = g.orderW + (int64_t)k * g.order_stride;
    if (blockIdx.x < NBOX) {
        const bool isR = blockIdx.x < NBOX_R;
        const int b = isR ? blockIdx.x : blockIdx.x - NBOX_R;
        const int nb = isR ? NBOX_R : NBOX_W;
        const int mode = isR ? pl.modeR : pl.modeW;
        const int m = isR ? pl.mR : pl.mW;
        const int* order = isR ? orderR : orderW;
        if (mode == MODE_REDUCE) {
            double2 acc[MAX_IT];
#pragma unroll
            for (int i = 0; i < MAX_IT; ++i) acc[i] = make_double2(0.0, 0.0);
            for (int rho = b * (APPLY_THREADS / 32) + warp; rho < m; rho += nb * (APPLY_THREADS / 32)) {
                const Src s = resolve(g, order[rho], t, par, st.cur);
                if (s.tag < 0) {
                    const double2* c2 = reinterpret_cast<const double2*>(s.col);
#pragma unroll
                    for (int i = 0; i < MAX_IT; ++i) {
                        const int idx = lane + 32 * i;
                        if (idx < n2) {
                            const double2 v = c2[idx];
                            acc[i].x += fabs(v.x);
                            acc[i].y += fabs(v.y);
                        }
                    }
                } else {
                    const int idx = s.tag >> 1;
#pragma unroll
                    for (int i = 0; i < MAX_IT; ++i)
                        if (lane + 32 * i == idx) {
                            if (s.tag & 1) acc[i].y += fabs(s.val); else acc[i].x += fabs(s.val);
                        }
                }
            }
#pragma unroll
            for (int i = 0; i < MAX_IT; ++i) {
                const int idx = lane + 32 * i;
                if (idx < n2) {
                    red_s[warp * g.ld + 2 * idx] = acc[i].x;
                    red_s[warp * g.ld + 2 * idx + 1] = acc[i].y;
                }
            }
            __syncthreads();
            double* part = g.partial + (int64_t)blockIdx.x * g.ld;
            for (int i = threadIdx.x; i < g.ld; i += APPLY_THREADS) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < APPLY_THREADS / 32; ++w) s += red_s[w * g.ld + i];
                part[i] = s;
            }
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last_flag = (atomicAdd(g.ticket, 1u) == (unsigned)(NBOX - 1));
        __syncthreads();
        if (last_flag) {
            __threadfence();
            finalize_step(g, t, k);
            if (threadIdx.x == 0) *g.ticket = 0u;
        }
        return;
    }
    // ---- copy jobs ----
    const int nwarps = (gridDim.x - NBOX) * (APPLY_THREADS / 32);
    const int w0 = (blockIdx.x - NBOX) * (APPLY_THREADS / 32) + warp;
    const int nmap = g.p0 + g.pV;
    double* slot = g.Rg + (int64_t)(k - 1) * g.slot;
    const int jobsR = g.capListR, jobs = g.capListR + g.capListW;
    for (int job = w0; job < jobs; job += nwarps) {
        if (job < jobsR) {
            if (pl.modeR == MODE_REDUCE) {
                if (job >= pl.mR && job < pl.pinR) {
                    const Src s = resolve(g, orderR[job], t, par, st.cur);
                    copy_column(g, s, slot + (int64_t)(job - pl.mR) * g.ld, lane);
                }
            } else {   // unchanged: [A valid | B] in list order
                if (job < g.p0) {
                    if (g.hS[t * g.q + job] < INFINITY) {
                        const Src s = resolve(g, job, t, par, st.cur);
                        copy_column(g, s, slot + (int64_t)g.cposS[t * g.q + job] * g.ld, lane);
                    }
                } else if (job - g.p0 < st.pW) {
                    const Src s = resolve(g, nmap + job - g.p0, t, par, st.cur);
                    copy_column(g, s, slot + (int64_t)(pl.pA + job - g.p0) * g.ld, lane);
                }
            }
        } else {
            const int j = job - jobsR;
            if (pl.modeW == MODE_REDUCE) {
                if (j >= pl.mW && j < pl.pinW) {
                    const Src s = resolve(g, orderW[j], t, par, st.cur);
                    if (s.tag < 0) copy_column(g, s, g.GW[st.cur ^ 1] + (int64_t)(j - pl.mW) * g.ld, lane);
                }
            } else if (j >= g.capW && j - g.capW < g.pV) {   // append the valid mapped input generators in place
                const int c = j - g.capW;
                if (g.hS[t * g.q + g.p0 + c] < INFINITY) {
                    const Src s = resolve(g, g.p0 + c, t, par, st.cur);
                    copy_column(g, s, g.GW[st.cur] + (int64_t)(st.pW + g.cposS[t * g.q + g.p0 + c]) * g.ld, lane);
                }
            }
        }
    }
}

// Runs in the last box CTA of the step: box vectors, descriptors of the next W, centres, counts.
__device__ void finalize_step(const GP& g, int t, int k) {
    __shared__ int scan[APPLY_THREADS];
    __shared__ int s_z;
    const StepPlan pl = g.plans[k];
    const int par = k & 1, nxt = par ^ 1;
    const WState st = g.st[par];
    const int tid = threadIdx.x;
    const int nmap = g.p0 + g.pV;
    const int* orderW = g.orderW + (int64_t)k * g.order_stride;
    const int per = (g.n + APPLY_THREADS - 1) / APPLY_THREADS;     // rows per thread (contiguous chunk)
    // ---------------- R: box vector of the discarded generators ----------------
    int ndR = 0;
    if (pl.modeR == MODE_REDUCE) {
        int cnt = 0;
        for (int i = tid * per; i < min(g.n, (tid + 1) * per); ++i) {
            double d = 0.0;
            for (int b = 0; b < NBOX_R; ++b) d += g.partial[(int64_t)b * g.ld + i];
            g.Rd[(int64_t)(k - 1) * g.ld + i] = d;
            cnt += (!g.strip || d != 0.0);
        }
        scan[tid] = cnt;
        __syncthreads();
        for (int o = 1; o < APPLY_THREADS; o <<= 1) {
            const int v = tid >= o ? scan[tid - o] : 0;
            __syncthreads();
            scan[tid] += v;
            __syncthreads();
        }
        ndR = scan[APPLY_THREADS - 1];
        int pos = scan[tid] - cnt;
        for (int i = tid * per; i < min(g.n, (tid + 1) * per); ++i) {
            const double d = g.Rd[(int64_t)(k - 1) * g.ld + i];
            if (!g.strip || d != 0.0) g.Rrow[(int64_t)(k - 1) * g.n + pos++] = i;
        }
        __syncthreads();
    }
    if (tid == 0) g.meta[k - 1] = make_int2(pl.modeR == MODE_REDUCE ? pl.pinR - pl.mR : pl.pinR, ndR);
    // ---------------- W ----------------
    if (pl.modeW == MODE_REDUCE) {
        const int nk = pl.pinW - pl.mW;
        // z = kept generators with score exactly 0 (they precede the new box generators in the stable order)
        if (tid == 0) s_z = 0;
        __syncthreads();
        int zl = 0;
        for (int pos = tid; pos < nk; pos += APPLY_THREADS) {
            const int code = orderW[pl.mW + pos];
            const double h = code < nmap ? g.hS[t * g.q + code] : g.hW[par][code - nmap];
            zl += (h == 0.0);
        }
        if (zl) atomicAdd(&s_z, zl);
        __syncthreads();
        const int z = s_z;
        // box vector, compacted over the non-zero rows in strip mode
        int cnt = 0;
        double dloc[8];
        for (int i = tid * per, u = 0; i < min(g.n, (tid + 1) * per); ++i, ++u) {
            double d = 0.0;
            for (int b = 0; b < NBOX_W; ++b) d += g.partial[(int64_t)(NBOX_R + b) * g.ld + i];
            if (u < 8) dloc[u] = d;
            cnt += (!g.strip || d != 0.0);
        }
        scan[tid] = cnt;
        __syncthreads();
        for (int o = 1; o < APPLY_THREADS; o <<= 1) {
            const int v = tid >= o ? scan[tid - o] : 0;
            __syncthreads();
            scan[tid] += v;
            __syncthreads();
        }
        const int nd = scan[APPLY_THREADS - 1];
        int ip = scan[tid] - cnt;
        for (int i = tid * per, u = 0; i < min(g.n, (tid + 1) * per); ++i, ++u) {
            const double d = dloc[u];
            if (!g.strip || d != 0.0) {
                const int pos = nk + ip;
                g.hW[nxt][pos] = 0.0;
                g.tagW[nxt][pos] = i;
                g.valW[nxt][pos] = d;
                g.rkW[nxt][pos] = z + ip;
                g.skW[nxt][z + ip] = 0.0;
                ++ip;
            }
        }
        for (int pos = tid; pos < nk; pos += APPLY_THREADS) {
            const int code = orderW[pl.mW + pos];
            double h, val = 0.0;
            int tag = -1;
            if (code < nmap) {
                h = g.hS[t * g.q + code];
            } else {
                const int j = code - nmap;
                h = g.hW[par][j]; tag = g.tagW[par][j]; val = g.valW[par][j];
            }
            const int rk = pos < z ? pos : pos + nd;
            g.hW[nxt][pos] = h;
            g.tagW[nxt][pos] = tag;
            g.valW[nxt][pos] = val;
            g.rkW[nxt][pos] = rk;
            g.skW[nxt][rk] = h;
        }
        if (tid == 0) {
            WState ns; ns.pW = nk + nd; ns.cur = st.cur ^ 1; ns.pad0 = ns.pad1 = 0;
            g.st[nxt] = ns;
        }
    } else {
        for (int sigma = tid; sigma < pl.pinW; sigma += APPLY_THREADS) {
            const int code = orderW[sigma];
            double h, val = 0.0;
            int tag = -1, pos;
            if (code < nmap) {
                h = g.hS[t * g.q + code];
                pos = st.pW + g.cposS[t * g.q + code];
            } else {
                const int j = code - nmap;
                h = g.hW[par][j]; tag = g.tagW[par][j]; val = g.valW[par][j];
                pos = j;
            }
            g.hW[nxt][pos] = h;
            g.tagW[nxt][pos] = tag;
            g.valW[nxt][pos] = val;
            g.rkW[nxt][pos] = sigma;
            g.skW[nxt][sigma] = h;
        }
        if (tid == 0) {
            WState ns; ns.pW = pl.pinW; ns.cur = st.cur; ns.pad0 = ns.pad1 = 0;
            g.st[nxt] = ns;
        }
    }
    // ---------------- centres: c_k = P c0 + cW ; cW += P cV (GLGM06/reach.jl:48,54) ----------------
    const double* yc0 = g.Y + (int64_t)(t * g.q + nmap) * g.ld;
    const double* ycV = yc0 + g.ld;
    for (int i = tid; i < g.n; i += APPLY_THREADS) {
        const double cw = g.cW[i];
        g.Rc[(int64_t)(k - 1) * g.ld + i] = yc0[i] + cw;
        g.cW[i] = cw + ycV[i];
    }
}

// Dense box generators of the reach sets of steps [k0, k1): column nk + c of slot k is d[row_c] e_{row_c}.
__global__ void __launch_bounds__(256) expand_diag_kernel(GP g, int k0, int k1) {
    const int k = k0 + blockIdx.y;
    if (k >= k1) return;
    const int2 mt = g.meta[k - 1];
    const int nk = mt.x, nd = mt.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n2 = g.ld / 2;
    double* slot = g.Rg + (int64_t)(k - 1) * g.slot;
    for (int c = blockIdx.x * 8 + warp; c < nd; c += gridDim.x * 8) {
        const int row = g.Rrow[(int64_t)(k - 1) * g.n + c];
        const double d = g.Rd[(int64_t)(k - 1) * g.ld + row];
        double2* d2 = reinterpret_cast<double2*>(slot + (int64_t)(nk + c) * g.ld);
        for (int i = lane; i < n2; i += 32) {
            double2 v = make_double2(0.0, 0.0);
            if (2 * i == row) v.x = d;
            if (2 * i + 1 == row) v.y = d;
            d2[i] = v;
        }
    }
}

// W_1 = V (GLGM06/reach.jl:41-42): the valid columns of the GV segment of X become the first W.
__global__ void __launch_bounds__(256) init_w_kernel(GP g, const double* X) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        WState s; s.pW = g.cntS[1]; s.cur = 0; s.pad0 = s.pad1 = 0;
        g.st[0] = s;
    }
    if (warp >= g.pV) return;
    const double h = g.hS[g.p0 + warp];
    if (!(h < INFINITY)) return;
    const int pos = g.cposS[g.p0 + warp];
    const double2* c2 = reinterpret_cast<const double2*>(X + (int64_t)(g.p0 + warp) * g.ld);
    double2* d2 = reinterpret_cast<double2*>(g.GW[0] + (int64_t)pos * g.ld);
    for (int i = lane; i < g.ld / 2; i += 32) d2[i] = c2[i];
    if (lane == 0) {
        const int rk = g.rkS[g.p0 + warp];
        g.hW[0][pos] = h;
        g.tagW[0][pos] = -1;
        g.valW[0][pos] = 0.0;
        g.rkW[0][pos] = rk;
        g.skW[0][rk] = h;
    }
}

// homogeneous flowpipes: number of non-zero columns per step is not tracked (linear_map of a non-singular Phi);
// box hull of zonotopes: lo/hi[i] = c[i] -/+ sum_j |G[i,j]|
__global__ void __launch_bounds__(256) zono_box_kernel(int n, int ld, int p, int64_t member_stride, int batch,
                                                       const double* __restrict__ G, const double* __restrict__ c,
                                                       int64_t c_stride, double* __restrict__ lo, double* __restrict__ hi) {
    const int b = blockIdx.y;
    if (b >= batch) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* Gb = G + (int64_t)b * member_stride;
    double s = 0.0;
    for (int j = 0; j < p; ++j) s += fabs(Gb[(int64_t)j * ld + i]);
    const double ci = c[(int64_t)b * c_stride + i];
    lo[(int64_t)b * n + i] = ci - s;
    hi[(int64_t)b * n + i] = ci + s;
}

// Dense copy of the current W (parity par): tagged box generators are synthesised.  out: n x pW, leading dim ldo.
__global__ void __launch_bounds__(256) materialize_w_kernel(GP g, int par, double* __restrict__ out, int ldo) {
    const WState st = g.st[par];
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= st.pW) return;
    const int tag = g.tagW[par][warp];
    const double val = g.valW[par][warp];
    const double* src = g.GW[st.cur] + (int64_t)warp * g.ld;
    double* dst = out + (int64_t)warp * ldo;
    for (int i = lane; i < g.n; i += 32) dst[i] = tag < 0 ? src[i] : (i == tag ? val : 0.0);
}

template <typename T>
T* dalloc(Ctx* ctx, size_t count, bool zero = true) {
    T* p = nullptr;
    if (count == 0) count = 1;
    REACH_CUDA(cudaMalloc(&p, sizeof(T) * count));
    if (zero) REACH_CUDA(cudaMemsetAsync(p, 0, sizeof(T) * count, ctx->stream));
    return p;
}

}  // namespace

// =========================================================================================================
struct Flowpipe {
    Ctx* ctx = nullptr;
    int64_t n = 0, ld = 0, p0 = 0, pV = 0, q = 0, N = 0, r = 0, s = 1;
    int flags = 0;
    bool homog = false, record = false;
    int64_t nc = 1;                 // homogeneous: number of centre columns (ensemble: one per member)
    GP g{};
    DMat Phi, PowA, PowB;
    double* X = nullptr;            // n x q  [G0 | GV | c0 | cV]
    double* stack[2] = {nullptr, nullptr};
    double* Hflow = nullptr;        // homogeneous: the flowpipe itself, N slots of q columns
    std::vector<void*> owned;
    std::vector<int2> meta_host;
    bool ran = false;
    double run_ms = 0, gemm_ms = 0, gemm_flops = 0, reduce_ms = 0, reduce_bytes = 0;
    int64_t gemm_launches = 0;
    std::vector<cudaEvent_t> evs;
    double* stage = nullptr;        // staging for strided downloads
};

void flowpipe_free(Flowpipe* F) {
    if (!F) return;
    dmat_free(F->Phi); dmat_free(F->PowA); dmat_free(F->PowB);
    for (void* p : F->owned) cudaFree(p);
    for (cudaEvent_t e : F->evs) cudaEventDestroy(e);
    delete F;
}

static int64_t pick_block(int64_t n, int64_t q, int64_t N) {
    // enough stacked columns for ~4 waves of 128-row tiles on 148 SMs
    const int64_t tiles_n = ceil_div(n, 96);
    int64_t s = 1;
    while (s < 64 && s < N - 1 && ceil_div(s * q, 128) * tiles_n < 148 * 4) s *= 2;
    return s;
}

Flowpipe* flowpipe_create(Ctx* ctx, int64_t n, const double* Phi, int64_t ldphi, int64_t p0, const double* c0,
                          const double* G0, int64_t ldg0, int64_t pV, const double* cV, const double* GV, int64_t ldgv,
                          int64_t N, int64_t max_order, int32_t flags, int64_t nc = 1) {
    REACH_REQUIRE(n > 0 && Phi && ldphi >= n, "glgm06: bad Phi");
    REACH_REQUIRE(p0 >= 0 && c0 && (p0 == 0 || (G0 && ldg0 >= n)), "glgm06: bad Omega0");
    REACH_REQUIRE(N >= 1, "glgm06: N must be >= 1");
    const bool homog = (cV == nullptr);
    if (homog) pV = 0;
    REACH_REQUIRE(pV >= 0 && (pV == 0 || (GV && ldgv >= n)), "glgm06: bad input zonotope");
    if (!homog && max_order < 1)
        throw Error(REACH_EUNSUPPORTED, "glgm06: max_order < 1 (reduce_order never reduces; unbounded growth) is not supported");
    if (!homog && n > 64 * MAX_IT)
        throw Error(REACH_EUNSUPPORTED, "glgm06: inhomogeneous path supports n <= 1024 in this build");
    Flowpipe* F = new Flowpipe();
    try {
        F->ctx = ctx; F->n = n; F->p0 = p0; F->pV = pV; F->N = N; F->r = max_order; F->flags = flags;
        F->homog = homog;
        F->record = (flags & REACH_GLGM06_RECORD_SELECTIONS) != 0;
        F->ld = round_up(n, 4);   // 32-byte aligned columns; the GEMM's K-tail over-read hits finite data times zero pad
        const int64_t ld = F->ld;
        F->nc = homog ? nc : 1;
        F->q = homog ? p0 + F->nc : p0 + pV + 2;
        const int64_t q = F->q;
        F->s = N > 2 ? pick_block(n, q, N) : 1;
        const int64_t s = F->s;
        cudaStream_t st = ctx->stream;
        F->Phi = dmat_alloc(ctx, n, n);
        dmat_upload(ctx, F->Phi, Phi, ldphi);
        if (s > 1) { F->PowA = dmat_alloc(ctx, n, n); F->PowB = dmat_alloc(ctx, n, n); }
        auto own = [&](void* p) { F->owned.push_back(p); return p; };
        auto up2d = [&](double* dst, const double* src, int64_t lds, int64_t cols) {
            if (cols > 0)
                REACH_CUDA(cudaMemcpy2DAsync(dst, ld * 8, src, lds * 8, n * 8, cols, cudaMemcpyHostToDevice, st));
        };
        GP& g = F->g;
        g.n = (int)n; g.ld = (int)ld; g.p0 = (int)p0; g.pV = (int)pV; g.q = (int)q;
        g.strip = (flags & REACH_GLGM06_KEEP_ZERO_GENERATORS) ? 0 : 1;
        if (homog) {
            // the flowpipe storage is the stack: slot k-1 = Phi^{k-1} [G0 | c0]  (ensemble: [G_0 .. G_{b-1} | c_0 .. c_{b-1}])
            const size_t cols = (size_t)N * q + 256;
            size_t free_b = 0, total_b = 0;
            REACH_CUDA(cudaMemGetInfo(&free_b, &total_b));
            if ((double)cols * ld * 8.0 > 0.92 * (double)free_b) {
                char buf[256];
                snprintf(buf, sizeof(buf), "glgm06: homogeneous flowpipe needs %.1f GB of HBM (N=%lld sets of %lld columns), "
                         "%.1f GB free: shard the ensemble across GPUs or shorten the horizon",
                         (double)cols * ld * 8.0 / 1e9, (long long)N, (long long)q, (double)free_b / 1e9);
                throw Error(REACH_EUNSUPPORTED, buf);
            }
            F->Hflow = (double*)own(dalloc<double>(ctx, cols * ld));
            up2d(F->Hflow, G0, ldg0, p0);
            up2d(F->Hflow + p0 * ld, c0, n, nc);
        } else {
            const int64_t rn = max_order * n;
            g.rn = rn; g.nkeep = n * (max_order - 1);
            g.capW = (int)std::max<int64_t>(rn, pV) + 8;
            g.capR = (int)std::max<int64_t>(rn, p0) + 8;
            g.capListR = (int)p0 + g.capW;
            g.capListW = g.capW + (int)pV;
            F->X = (double*)own(dalloc<double>(ctx, (size_t)(q + 128) * ld));
            up2d(F->X, G0, ldg0, p0);
            up2d(F->X + p0 * ld, GV, ldgv, pV);
            up2d(F->X + (p0 + pV) * ld, c0, n, 1);
            up2d(F->X + (p0 + pV + 1) * ld, cV, n, 1);
            for (int i = 0; i < 2; ++i) {
                F->stack[i] = (double*)own(dalloc<double>(ctx, (size_t)(s * q + 128) * ld));
                g.GW[i] = (double*)own(dalloc<double>(ctx, (size_t)(g.capW + (int)pV + 8) * ld));
                g.hW[i] = (double*)own(dalloc<double>(ctx, g.capListW + 8));
                g.valW[i] = (double*)own(dalloc<double>(ctx, g.capListW + 8));
                g.skW[i] = (double*)own(dalloc<double>(ctx, g.capListW + 8));
                g.tagW[i] = (int*)own(dalloc<int>(ctx, g.capListW + 8));
                g.rkW[i] = (int*)own(dalloc<int>(ctx, g.capListW + 8));
            }
            g.hS = (double*)own(dalloc<double>(ctx, s * q + 8));
            g.skS = (double*)own(dalloc<double>(ctx, s * q + 8));
            g.rkS = (int*)own(dalloc<int>(ctx, s * q + 8));
            g.cposS = (int*)own(dalloc<int>(ctx, s * q + 8));
            g.cntS = (int*)own(dalloc<int>(ctx, 2 * s + 8));
            g.st = (WState*)own(dalloc<WState>(ctx, 2));
            g.cW = (double*)own(dalloc<double>(ctx, ld));
            g.order_stride = F->record ? (int64_t)(g.capListR + g.capListW) : 0;
            const size_t order_len = (size_t)(g.capListR + g.capListW) * (F->record ? (size_t)(N + 1) : 1);
            g.orderR = (int*)own(dalloc<int>(ctx, order_len));
            g.orderW = g.orderR + g.capListR;
            g.rec_stride = g.order_stride;
            g.recH = F->record ? (double*)own(dalloc<double>(ctx, order_len)) : nullptr;
            g.plans = (StepPlan*)own(dalloc<StepPlan>(ctx, N + 2));
            g.slot = ld * (int64_t)g.capR;
            size_t free_b = 0, total_b = 0;
            REACH_CUDA(cudaMemGetInfo(&free_b, &total_b));
            const double need = (double)N * (double)g.slot * 8.0;
            if (need > 0.92 * (double)free_b) {
                char buf[256];
                snprintf(buf, sizeof(buf), "glgm06: flowpipe needs %.1f GB of HBM (N=%lld sets of %d x %d), %.1f GB free",
                         need / 1e9, (long long)N, (int)n, g.capR, (double)free_b / 1e9);
                throw Error(REACH_EUNSUPPORTED, buf);
            }
            g.Rg = (double*)own(dalloc<double>(ctx, (size_t)N * g.slot, false));
            g.Rc = (double*)own(dalloc<double>(ctx, (size_t)N * ld));
            g.Rd = (double*)own(dalloc<double>(ctx, (size_t)N * ld));
            g.Rrow = (int*)own(dalloc<int>(ctx, (size_t)N * n));
            g.meta = (int2*)own(dalloc<int2>(ctx, N + 1));
            g.partial = (double*)own(dalloc<double>(ctx, (size_t)NBOX * ld));
            g.ticket = (unsigned*)own(dalloc<unsigned>(ctx, 1));
            // R_1 = Omega0 (GLGM06/reach.jl:39-40)
            if (p0 > 0)
                REACH_CUDA(cudaMemcpy2DAsync(g.Rg, ld * 8, G0, ldg0 * 8, n * 8, p0, cudaMemcpyHostToDevice, st));
            REACH_CUDA(cudaMemcpyAsync(g.Rc, c0, n * 8, cudaMemcpyHostToDevice, st));
            if (ld > n) {   // pad rows of slot 0 are read by nobody, but keep them finite
                REACH_CUDA(cudaMemset2DAsync(g.Rg + n, ld * 8, 0, (ld - n) * 8, p0 > 0 ? p0 : 1, st));
            }
            const int2 m1 = make_int2((int)p0, 0);
            REACH_CUDA(cudaMemcpyAsync(g.meta, &m1, sizeof(int2), cudaMemcpyHostToDevice, st));
        }
        REACH_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        flowpipe_free(F);
        throw;
    }
    return F;
}

static cudaEvent_t next_event(Flowpipe* F, size_t& used) {
    if (used == F->evs.size()) {
        cudaEvent_t e;
        REACH_CUDA(cudaEventCreate(&e));
        F->evs.push_back(e);
    }
    return F->evs[used++];
}

void flowpipe_run(Flowpipe* F) {
    Ctx* ctx = F->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t n = F->n, ld = F->ld, q = F->q, N = F->N, s = F->s;
    GP g = F->g;
    size_t ev_used = 0;
    F->gemm_launches = 0; F->gemm_flops = 0;
    REACH_CUDA(cudaEventRecord(ctx->ev0, st));
    std::vector<std::pair<size_t, size_t>> gemm_ev, red_ev;
    auto timed_gemm = [&](int64_t Q, const DMat& P, const double* Xp, double* Yp) {
        cudaEvent_t e0 = next_event(F, ev_used), e1 = next_event(F, ev_used);
        gemm_ev.push_back({ev_used - 2, ev_used - 1});
        REACH_CUDA(cudaEventRecord(e0, st));
        gemm_left(ctx, n, Q, n, P.p, P.ld, Xp, ld, Yp, ld);
        REACH_CUDA(cudaEventRecord(e1, st));
        F->gemm_launches++;
        F->gemm_flops += 2.0 * (double)n * (double)n * (double)Q;
    };
    auto square = [&](const DMat& P, DMat& out) {
        cudaEvent_t e0 = next_event(F, ev_used), e1 = next_event(F, ev_used);
        gemm_ev.push_back({ev_used - 2, ev_used - 1});
        REACH_CUDA(cudaEventRecord(e0, st));
        gemm_left(ctx, n, n, n, P.p, P.ld, P.p, P.ld, out.p, out.ld);
        REACH_CUDA(cudaEventRecord(e1, st));
        F->gemm_launches++;
        F->gemm_flops += 2.0 * (double)n * n * n;
    };
    const int64_t npow = N - 1;          // Y_1 .. Y_{N-1}
    if (F->homog) {
        // slot j = Phi^j [G0 | c0] (R_{j+1}, GLGM06/reach.jl:15-19).  With the power Phi^w in hand, slots
        // [have, have + w) = Phi^w * slots [have - w, have): doubling until w = s, then s slots per GEMM.
        const DMat* pw = &F->Phi;
        DMat* spare[2] = {&F->PowA, &F->PowB};
        int sp = 0;
        int64_t have = 1, w = 1;
        while (have < N) {
            const int64_t cnt = std::min(w, N - have);
            timed_gemm(cnt * q, *pw, F->Hflow + (have - w) * q * ld, F->Hflow + have * q * ld);
            have += cnt;
            if (have < N && w < s) {
                square(*pw, *spare[sp]);
                pw = spare[sp];
                sp ^= 1;
                w *= 2;
            }
        }
        REACH_CUDA(cudaEventRecord(ctx->ev1, st));
        REACH_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        REACH_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        F->run_ms = ms;
        F->gemm_ms = 0; F->reduce_ms = 0;
        for (auto& pr : gemm_ev) {
            float t = 0;
            REACH_CUDA(cudaEventElapsedTime(&t, F->evs[pr.first], F->evs[pr.second]));
            F->gemm_ms += t;
        }
        F->meta_host.assign(N, make_int2((int)F->p0, 0));
        F->ran = true;
        return;
    }
    // ---------------- inhomogeneous ----------------
    const int per = (int)(g.p0 + g.pV);
    auto score_and_sort = [&](const double* Y, int cnt) {
        g.Y = Y;
        if (per > 0) {
            const int warps = cnt * per;
            score_stack_kernel<<<(unsigned)ceil_div(warps, 8), 256, 0, st>>>(g, cnt);
            REACH_CUDA(cudaGetLastError());
            const int maxlen = (int)std::max(g.p0, g.pV);
            dim3 grid((unsigned)std::max<int64_t>(1, ceil_div(maxlen, 256)), (unsigned)(cnt * 2));
            sort_segments_kernel<<<grid, 256, 0, st>>>(g);
            REACH_CUDA(cudaGetLastError());
            ctx->launches += 2;
        } else {
            REACH_CUDA(cudaMemsetAsync(g.cntS, 0, sizeof(int) * 2 * cnt, st));
        }
    };
    // W_1 = V (GLGM06/reach.jl:41-42): cW = cV, scores / order of X's own input segment
    REACH_CUDA(cudaMemcpyAsync(g.cW, F->X + (g.p0 + g.pV + 1) * ld, n * 8, cudaMemcpyDeviceToDevice, st));
    score_and_sort(F->X, 1);
    init_w_kernel<<<(unsigned)std::max<int64_t>(1, ceil_div(g.pV, 8)), 256, 0, st>>>(g, F->X);
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
    if (npow > 0) {
        // first block by doubling: stack0[0] = Phi X, stack0[have..2have) = Phi^have stack0[0..have)
        timed_gemm(q, F->Phi, F->X, F->stack[0]);
        const DMat* pw = &F->Phi;
        DMat* spare[2] = {&F->PowA, &F->PowB};
        int sp = 0;
        for (int64_t have = 1; have < s; have *= 2) {
            timed_gemm(have * q, *pw, F->stack[0], F->stack[0] + have * q * ld);
            square(*pw, *spare[sp]);
            pw = spare[sp];
            sp ^= 1;
        }
        // pw = Phi^s
        const size_t smem_keys = (size_t)(g.p0 + g.capW + g.pV) * sizeof(double);
        const int use_smem = smem_keys <= 200 * 1024 ? 1 : 0;
        static bool attr_set = false;
        if (use_smem && !attr_set) {
            REACH_CUDA(cudaFuncSetAttribute(rank_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set = true;
        }
        const size_t apply_smem = (size_t)(APPLY_THREADS / 32) * ld * sizeof(double);
        static bool apply_attr = false;
        if (!apply_attr) {
            REACH_CUDA(cudaFuncSetAttribute(apply_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
            apply_attr = true;
        }
        const int rank_threads = g.p0 + 2 * g.capW + g.pV;
        const unsigned rank_grid = (unsigned)ceil_div(rank_threads, 256);
        const unsigned apply_grid = NBOX + 2 * (unsigned)ctx->sm_count;
        const int64_t nblocks = ceil_div(npow, s);
        int cur = 0;
        for (int64_t b = 0; b < nblocks; ++b) {
            const int cnt = (int)std::min<int64_t>(s, npow - b * s);
            if (b > 0) {
                timed_gemm((int64_t)cnt * q, *pw, F->stack[cur ^ 1], F->stack[cur]);
            }
            cudaEvent_t e0 = next_event(F, ev_used), e1 = next_event(F, ev_used);
            red_ev.push_back({ev_used - 2, ev_used - 1});
            REACH_CUDA(cudaEventRecord(e0, st));
            score_and_sort(F->stack[cur], cnt);
            for (int t = 0; t < cnt; ++t) {
                const int k = (int)(b * s) + t + 2;
                rank_lists_kernel<<<rank_grid, 256, use_smem ? smem_keys : 0, st>>>(g, t, k, use_smem);
                apply_step_kernel<<<apply_grid, APPLY_THREADS, apply_smem, st>>>(g, t, k);
            }
            REACH_CUDA(cudaGetLastError());
            const int k0 = (int)(b * s) + 2;
            dim3 eg(8, (unsigned)cnt);
            expand_diag_kernel<<<eg, 256, 0, st>>>(g, k0, k0 + cnt);
            REACH_CUDA(cudaGetLastError());
            REACH_CUDA(cudaEventRecord(e1, st));
            ctx->launches += 2 * (uint64_t)cnt + 1;
            cur ^= 1;
        }
    }
    REACH_CUDA(cudaEventRecord(ctx->ev1, st));
    F->meta_host.resize(N);
    REACH_CUDA(cudaMemcpyAsync(F->meta_host.data(), g.meta, sizeof(int2) * N, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    REACH_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    F->run_ms = ms;
    F->gemm_ms = 0; F->reduce_ms = 0;
    for (auto& pr : gemm_ev) {
        float t = 0;
        REACH_CUDA(cudaEventElapsedTime(&t, F->evs[pr.first], F->evs[pr.second]));
        F->gemm_ms += t;
    }
    for (auto& pr : red_ev) {
        float t = 0;
        REACH_CUDA(cudaEventElapsedTime(&t, F->evs[pr.first], F->evs[pr.second]));
        F->reduce_ms += t;
    }
    F->ran = true;
}

// reduce_order(Z, r) of LazySets (Girard 2005; call site GLGM06/post.jl:20) as ONE step of the same
// machinery: W list = [ (empty W) | the columns of G ], so scores, stable order, boxing and zero-generator removal
// are exactly the kernels the time loop uses.
void reduce_order_once(Ctx* ctx, int64_t n, int64_t p, const double* c, const double* G, int64_t ldg, int64_t r,
                       int32_t flags, double* Gout, int64_t ldgo, int64_t* p_out, int32_t* ind) {
    REACH_REQUIRE(n > 0 && p >= 0 && c && (p == 0 || (G && ldg >= n)) && Gout && ldgo >= n && p_out,
                  "reach_reduce_order: bad arguments");
    cudaStream_t st = ctx->stream;
    if (r < 1 || p == 0) {      // reduce_order returns Z itself
        for (int64_t j = 0; j < p; ++j) std::memcpy(Gout + j * ldgo, G + j * ldg, sizeof(double) * n);
        *p_out = p;
        if (ind) for (int64_t j = 0; j < p; ++j) ind[j] = (int32_t)j;
        return;
    }
    std::vector<double> zeros((size_t)n * n, 0.0);
    Flowpipe* F = flowpipe_create(ctx, n, zeros.data(), n, 0, c, nullptr, n, p, c, G, ldg, 2, r, flags & REACH_GLGM06_KEEP_ZERO_GENERATORS);
    double* dout = nullptr;
    try {
        GP g = F->g;
        g.Y = F->X;
        const int per = (int)p;
        score_stack_kernel<<<(unsigned)ceil_div(per, 8), 256, 0, st>>>(g, 1);
        dim3 grid((unsigned)std::max<int64_t>(1, ceil_div(per, 256)), 2u);
        sort_segments_kernel<<<grid, 256, 0, st>>>(g);
        REACH_CUDA(cudaGetLastError());
        REACH_CUDA(cudaMemsetAsync(g.st, 0, 2 * sizeof(WState), st));       // W empty, buffer 0
        const int rank_threads = g.p0 + 2 * g.capW + g.pV;
        rank_lists_kernel<<<(unsigned)ceil_div(rank_threads, 256), 256, 0, st>>>(g, 0, 2, 0);
        const size_t apply_smem = (size_t)(APPLY_THREADS / 32) * F->ld * sizeof(double);
        REACH_CUDA(cudaFuncSetAttribute(apply_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        apply_step_kernel<<<NBOX + 2 * (unsigned)ctx->sm_count, APPLY_THREADS, apply_smem, st>>>(g, 0, 2);
        REACH_CUDA(cudaGetLastError());
        const int64_t cap = g.capW;
        REACH_CUDA(cudaMalloc(&dout, sizeof(double) * (size_t)cap * n));
        materialize_w_kernel<<<(unsigned)ceil_div(cap, 8), 256, 0, st>>>(g, 1, dout, (int)n);
        REACH_CUDA(cudaGetLastError());
        ctx->launches += 5;
        WState ns;
        StepPlan pl;
        REACH_CUDA(cudaMemcpyAsync(&ns, g.st + 1, sizeof(ns), cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaMemcpyAsync(&pl, g.plans + 2, sizeof(pl), cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaStreamSynchronize(st));
        *p_out = ns.pW;
        if (ns.pW > 0)
            REACH_CUDA(cudaMemcpy2DAsync(Gout, ldgo * 8, dout, n * 8, n * 8, ns.pW, cudaMemcpyDeviceToHost, st));
        if (ind) {
            std::vector<int> order(std::max(pl.pinW, 1));
            std::vector<int> cpos(p);
            REACH_CUDA(cudaMemcpyAsync(order.data(), g.orderW + 2 * g.order_stride, sizeof(int) * pl.pinW, cudaMemcpyDeviceToHost, st));
            REACH_CUDA(cudaMemcpyAsync(cpos.data(), g.cposS, sizeof(int) * p, cudaMemcpyDeviceToHost, st));
            REACH_CUDA(cudaStreamSynchronize(st));
            for (int sgm = 0; sgm < pl.pinW; ++sgm) ind[sgm] = cpos[order[sgm]];     // code == column (p0 = 0)
            for (int64_t j = pl.pinW; j < p; ++j) ind[j] = -1;
        }
        REACH_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        if (dout) cudaFree(dout);
        flowpipe_free(F);
        throw;
    }
    cudaFree(dout);
    flowpipe_free(F);
}

}  // namespace reach

// =========================================================================================================
// C ABI (include/reach.h)
// =========================================================================================================
namespace reach { Ctx* ctx_of(reach_ctx* h); }
using namespace reach;

struct reach_flowpipe {
    Flowpipe* f;
    Ctx* ctx;
};

#define FP_BEGIN(ctx)                                     \
    if (!(ctx)) return REACH_EINVAL;                      \
    try {                                                 \
        REACH_CUDA(cudaSetDevice((ctx)->device));
#define FP_END(ctx)                                                                               \
    }                                                                                             \
    catch (const reach::Error& e) { (ctx)->last_error = e.what(); return e.code; }                \
    catch (const std::exception& e) { (ctx)->last_error = e.what(); return REACH_ECUDA; }

namespace {

void require_ran(Flowpipe* F) { REACH_REQUIRE(F->ran, "flowpipe: call reach_glgm06_run first"); }

// device pointer of reach set k (1-based): generators, centre, counts
void locate(Flowpipe* F, int64_t k, const double** G, const double** c, int* p) {
    REACH_REQUIRE(k >= 1 && k <= F->N, "flowpipe: step index out of range");
    if (F->homog) {
        const double* slot = F->Hflow + (k - 1) * F->q * F->ld;
        *G = slot; *c = slot + F->p0 * F->ld; *p = (int)F->p0;
    } else {
        *G = F->g.Rg + (k - 1) * F->g.slot;
        *c = F->g.Rc + (k - 1) * F->ld;
        const int2 m = F->meta_host[k - 1];
        *p = m.x + m.y;
    }
}

}  // namespace

extern "C" {

int32_t reach_glgm06_create(reach_ctx* h, int64_t n, const double* Phi, int64_t ldphi, int64_t p0, const double* c0,
                            const double* G0, int64_t ldg0, int64_t pV, const double* cV, const double* GV, int64_t ldgv,
                            int64_t N, int64_t max_order, int32_t flags, reach_flowpipe** out) {
    if (!h || !out) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    *out = nullptr;
    FP_BEGIN(ctx)
    Flowpipe* f = flowpipe_create(ctx, n, Phi, ldphi, p0, c0, G0, ldg0, pV, cV, GV, ldgv, N, max_order, flags);
    *out = new reach_flowpipe{f, ctx};
    return REACH_OK;
    FP_END(ctx)
}

int32_t reach_glgm06_run(reach_flowpipe* fp) {
    if (!fp) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    flowpipe_run(fp->f);
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_flowpipe_ngens(reach_flowpipe* fp, int32_t* p) {
    if (!fp || !p) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    require_ran(fp->f);
    for (int64_t k = 0; k < fp->f->N; ++k) p[k] = fp->f->meta_host[k].x + fp->f->meta_host[k].y;
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_flowpipe_step(reach_flowpipe* fp, int64_t k, double* c, double* G, int64_t ldg) {
    if (!fp) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    Flowpipe* F = fp->f;
    require_ran(F);
    const double *dG, *dc;
    int p;
    locate(F, k, &dG, &dc, &p);
    cudaStream_t st = F->ctx->stream;
    if (c) REACH_CUDA(cudaMemcpyAsync(c, dc, F->n * 8, cudaMemcpyDeviceToHost, st));
    if (G && p > 0) {
        REACH_REQUIRE(ldg >= F->n, "reach_flowpipe_step: ldg < n");
        REACH_CUDA(cudaMemcpy2DAsync(G, ldg * 8, dG, F->ld * 8, F->n * 8, p, cudaMemcpyDeviceToHost, st));
    }
    REACH_CUDA(cudaStreamSynchronize(st));
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_flowpipe_fetch(reach_flowpipe* fp, double* centers, double* G, int64_t pmax) {
    if (!fp) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    Flowpipe* F = fp->f;
    require_ran(F);
    cudaStream_t st = F->ctx->stream;
    for (int64_t k = 1; k <= F->N; ++k) {
        const double *dG, *dc;
        int p;
        locate(F, k, &dG, &dc, &p);
        if (centers) REACH_CUDA(cudaMemcpyAsync(centers + (k - 1) * F->n, dc, F->n * 8, cudaMemcpyDeviceToHost, st));
        if (G && p > 0) {
            REACH_REQUIRE(p <= pmax, "reach_flowpipe_fetch: pmax smaller than a reach set's generator count");
            REACH_CUDA(cudaMemcpy2DAsync(G + (k - 1) * F->n * pmax, F->n * 8, dG, F->ld * 8, F->n * 8, p,
                                         cudaMemcpyDeviceToHost, st));
        }
    }
    REACH_CUDA(cudaStreamSynchronize(st));
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_flowpipe_box(reach_flowpipe* fp, int64_t k, double* lo, double* hi) {
    if (!fp || !lo || !hi) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    Flowpipe* F = fp->f;
    require_ran(F);
    const double *dG, *dc;
    int p;
    locate(F, k, &dG, &dc, &p);
    cudaStream_t st = F->ctx->stream;
    if (!F->stage) {
        REACH_CUDA(cudaMalloc(&F->stage, 2 * F->n * 8));
        F->owned.push_back(F->stage);
    }
    zono_box_kernel<<<dim3((unsigned)ceil_div(F->n, 256), 1), 256, 0, st>>>((int)F->n, (int)F->ld, p, 0, 1, dG, dc, 0,
                                                                             F->stage, F->stage + F->n);
    REACH_CUDA(cudaGetLastError());
    F->ctx->launches++;
    REACH_CUDA(cudaMemcpyAsync(lo, F->stage, F->n * 8, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaMemcpyAsync(hi, F->stage + F->n, F->n * 8, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaStreamSynchronize(st));
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_flowpipe_selection(reach_flowpipe* fp, int64_t k, int32_t which, int32_t* p_in, int32_t* m, int32_t* ind,
                                 double* hout) {
    if (!fp) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    Flowpipe* F = fp->f;
    require_ran(F);
    REACH_REQUIRE(!F->homog, "reach_flowpipe_selection: the homogeneous recurrence never reduces");
    REACH_REQUIRE(F->record, "reach_flowpipe_selection: create the flowpipe with REACH_GLGM06_RECORD_SELECTIONS");
    REACH_REQUIRE(k >= 2 && k <= F->N && (which == 0 || which == 1), "reach_flowpipe_selection: bad k / which");
    const GP& g = F->g;
    cudaStream_t st = F->ctx->stream;
    StepPlan pl;
    REACH_CUDA(cudaMemcpyAsync(&pl, g.plans + k, sizeof(pl), cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaStreamSynchronize(st));
    const int mode = which ? pl.modeW : pl.modeR;
    const int pin = which ? pl.pinW : pl.pinR;
    const int mm = which ? pl.mW : pl.mR;
    if (mode == MODE_UNCHANGED) {
        if (p_in) *p_in = 0;
        if (m) *m = 0;
        return REACH_OK;
    }
    if (p_in) *p_in = pin;
    if (m) *m = mm;
    if (!ind && !hout) return REACH_OK;
    const int cap = which ? g.capListW : g.capListR;
    std::vector<int> order(pin);
    std::vector<double> hh(cap);
    const int* dord = (which ? g.orderW : g.orderR) + k * g.order_stride;
    const double* dh = g.recH + k * g.rec_stride + (which ? g.capListR : 0);
    REACH_CUDA(cudaMemcpyAsync(order.data(), dord, sizeof(int) * pin, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaMemcpyAsync(hh.data(), dh, sizeof(double) * cap, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaStreamSynchronize(st));
    // list element index of a source code, then compaction over the dropped (all-zero) mapped generators
    const int nmap = g.p0 + g.pV;
    const int pW = which ? pin - pl.pC : pin - pl.pA;     // generators of W that went in
    const int seg_len = which ? g.pV : g.p0;              // mapped segment (uncompacted)
    const int seg_off = which ? pW : 0;                   // its offset in the list
    std::vector<int> cpos(seg_len);
    {
        int c = 0;
        for (int e = 0; e < seg_len; ++e) {
            cpos[e] = c;
            c += (hh[seg_off + e] < INFINITY);
        }
    }
    for (int sgm = 0; sgm < pin; ++sgm) {
        const int code = order[sgm];
        int idx;      // compacted list index
        double hv;
        if (code < nmap) {
            const int e = which ? code - g.p0 : code;
            idx = (which ? pW : 0) + cpos[e];
            hv = hh[seg_off + e];
        } else {
            const int j = code - nmap;
            idx = which ? j : pl.pA + j;
            hv = hh[which ? j : g.p0 + j];
        }
        if (ind) ind[sgm] = idx;
        if (hout) hout[idx] = hv;
    }
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_flowpipe_timing(reach_flowpipe* fp, double* run_ms, double* gemm_ms, int64_t* gemm_launches,
                              double* gemm_flops, double* reduce_ms, double* reduce_bytes) {
    if (!fp) return REACH_EINVAL;
    Flowpipe* F = fp->f;
    if (run_ms) *run_ms = F->run_ms;
    if (gemm_ms) *gemm_ms = F->gemm_ms;
    if (gemm_launches) *gemm_launches = F->gemm_launches;
    if (gemm_flops) *gemm_flops = F->gemm_flops;
    if (reduce_ms) *reduce_ms = F->reduce_ms;
    if (reduce_bytes) *reduce_bytes = F->reduce_bytes;
    return REACH_OK;
}

int32_t reach_reduce_order(reach_ctx* h, int64_t n, int64_t p, const double* c, const double* G, int64_t ldg, int64_t r,
                           int32_t flags, double* Gout, int64_t ldgo, int64_t* p_out, int32_t* ind) {
    if (!h) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    FP_BEGIN(ctx)
    reduce_order_once(ctx, n, p, c, G, ldg, r, flags, Gout, ldgo, p_out, ind);
    return REACH_OK;
    FP_END(ctx)
}

// ---- batched homogeneous ensemble (config C5): the same homogeneous recurrence over batch * (p + 1) columns ----
struct reach_ensemble {
    Flowpipe* f;
    Ctx* ctx;
    int64_t batch, p;
    double* stage;
};

int32_t reach_ensemble_create(reach_ctx* h, int64_t n, const double* Phi, int64_t ldphi, int64_t batch, int64_t p,
                              const double* c0s, const double* G0s, int64_t N, reach_ensemble** out) {
    if (!h || !out) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    *out = nullptr;
    FP_BEGIN(ctx)
    REACH_REQUIRE(batch >= 1 && p >= 0 && c0s && (p == 0 || G0s), "reach_ensemble_create: bad arguments");
    Flowpipe* f = flowpipe_create(ctx, n, Phi, ldphi, batch * p, c0s, G0s, n, 0, nullptr, nullptr, n, N, 1, 0, batch);
    *out = new reach_ensemble{f, ctx, batch, p, nullptr};
    return REACH_OK;
    FP_END(ctx)
}

int32_t reach_ensemble_run(reach_ensemble* en) {
    if (!en) return REACH_EINVAL;
    FP_BEGIN(en->ctx)
    flowpipe_run(en->f);
    return REACH_OK;
    FP_END(en->ctx)
}

int32_t reach_ensemble_step(reach_ensemble* en, int64_t b, int64_t k, double* c, double* G) {
    if (!en) return REACH_EINVAL;
    FP_BEGIN(en->ctx)
    Flowpipe* F = en->f;
    require_ran(F);
    REACH_REQUIRE(b >= 0 && b < en->batch && k >= 1 && k <= F->N, "reach_ensemble_step: member / step out of range");
    const double* slot = F->Hflow + (k - 1) * F->q * F->ld;
    cudaStream_t st = F->ctx->stream;
    if (c) REACH_CUDA(cudaMemcpyAsync(c, slot + (en->batch * en->p + b) * F->ld, F->n * 8, cudaMemcpyDeviceToHost, st));
    if (G && en->p > 0)
        REACH_CUDA(cudaMemcpy2DAsync(G, F->n * 8, slot + b * en->p * F->ld, F->ld * 8, F->n * 8, en->p, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaStreamSynchronize(st));
    return REACH_OK;
    FP_END(en->ctx)
}

int32_t reach_ensemble_boxes(reach_ensemble* en, int64_t k, double* lo, double* hi) {
    if (!en || !lo || !hi) return REACH_EINVAL;
    FP_BEGIN(en->ctx)
    Flowpipe* F = en->f;
    require_ran(F);
    REACH_REQUIRE(k >= 1 && k <= F->N, "reach_ensemble_boxes: step out of range");
    cudaStream_t st = F->ctx->stream;
    const size_t cnt = (size_t)F->n * en->batch;
    if (!en->stage) {
        REACH_CUDA(cudaMalloc(&en->stage, 2 * cnt * 8));
        F->owned.push_back(en->stage);
    }
    const double* slot = F->Hflow + (k - 1) * F->q * F->ld;
    zono_box_kernel<<<dim3((unsigned)ceil_div(F->n, 64), (unsigned)en->batch), 64, 0, st>>>(
        (int)F->n, (int)F->ld, (int)en->p, en->p * F->ld, (int)en->batch, slot, slot + en->batch * en->p * F->ld, F->ld, en->stage,
        en->stage + cnt);
    REACH_CUDA(cudaGetLastError());
    F->ctx->launches++;
    REACH_CUDA(cudaMemcpyAsync(lo, en->stage, cnt * 8, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaMemcpyAsync(hi, en->stage + cnt, cnt * 8, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaStreamSynchronize(st));
    return REACH_OK;
    FP_END(en->ctx)
}

int32_t reach_ensemble_timing(reach_ensemble* en, double* run_ms, double* bytes_per_step) {
    if (!en) return REACH_EINVAL;
    if (run_ms) *run_ms = en->f->run_ms;
    if (bytes_per_step) *bytes_per_step = 16.0 * (double)en->f->n * (double)(en->p + 1) * (double)en->batch;
    return REACH_OK;
}

void reach_ensemble_free(reach_ensemble* en) {
    if (!en) return;
    cudaSetDevice(en->ctx->device);
    cudaStreamSynchronize(en->ctx->stream);
    flowpipe_free(en->f);
    delete en;
}

void reach_flowpipe_free(reach_flowpipe* fp) {
    if (!fp) return;
    cudaSetDevice(fp->ctx->device);
    cudaStreamSynchronize(fp->ctx->stream);
    flowpipe_free(fp->f);
    delete fp;
}

}  // extern "C"
