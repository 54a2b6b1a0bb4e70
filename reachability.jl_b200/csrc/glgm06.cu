// GLGM06 zonotope post-operator on the device.
//
// Replaces `reach_homog!` / `reach_inhomog!` (src/ReachSets/ContinuousPost/GLGM06/reach.jl:5-62) and the LazySets
// calls they make (linear_map, minkowski_sum, reduce_order -- SURVEY.md a8-a11, Appendix A.4/B.1-B.3):
//     R_k = reduce_order( (P_{k-1} c0 + cW, [P_{k-1} G0 | GW]), r )                                   (:48-49)
//     W   = reduce_order( (cW + P_{k-1} cV, [GW | P_{k-1} GV]), r )                                   (:54-55)
//     P_k = P_{k-1} Phi                                                                               (:57-58)
//
// B200 design ("archive + chain + batched numerics")
//  * ARCHIVE.  With X = [G0 | GV | c0 | cV] (n x q) every mapped generator the recurrence ever touches is a column of
//    Y_j = Phi^j X.  All Y_j live in one archive (slot j), produced s slots at a time by ONE large DMMA GEMM
//    slots[have .. have+s) = Phi^s * slots[have-s .. have)   (left multiplication, gemm_left, columns stay contiguous).
//    A generator of W is therefore never copied: it is an immutable REFERENCE, either an archive column or a box
//    generator (step, row) whose value sits in the archive of W box vectors.
//  * CHAIN.  The only sequential, data-dependent part is the Girard SELECTION of W (which generators are boxed).  It is
//    pure key logic: every segment carries its stable sorted order, so the sortperm of [W | Phi^k GV] is a merge
//    (merge path: a binary search per thread along its diagonal, then a short sequential merge; ties by list index like
//    Julia's stable sortperm).  One persistent
//    CTA keeps the sorted (score, reference) list of W in shared memory and advances a whole block of steps without any
//    grid-wide synchronisation; zero-generator removal needs only row-support bit masks, never the values.
//  * BATCHED NUMERICS.  Everything numerical is then embarrassingly parallel over the steps of the block: box vectors of
//    the discarded generators (fixed summation order, no float atomics), the R-side selection [Phi^k G0 | W_{k-1}]
//    against the per-step snapshot of W, the gather of the kept generators into the flowpipe slots (columns staged through
//    shared memory with bulk copies, TMA engine), centres, and the dense box generators.
//  * PIPELINE.  Six streams: map of batch b+1 (main) | scores + segment order | selection chain of batch b | W box sums + W
//    recurrence | R sortperm + R box sums of batch b-1; the persistent GEMM leaves K SMs to the chain and the HBM-bound numerics.
//  * The reach sets themselves are stored as REFERENCES (kept generators) + box vector; dense matrices are materialised on
//    demand (fetch / step / box / projection kernels).  The steady-state selection runs as a PARALLEL chain (one warp per
//    generator, chain_fast_kernel), the sequential merge-path chain is the fallback.
//  * SPARSE PHI.  When Phi (and the powers the run uses) has at most 8 non-zeros per row, the map is an ELL sparse product
//    fused with the Girard scores instead of the GEMM (decoupled modes: the ISS-shaped C2).
//  * All counts live on the device; the host enqueues the whole flowpipe without synchronising.
#include "common.cuh"
#include "../../include/reach.h"
#include <algorithm>
#include <cmath>
#include <climits>
#include <mutex>
#include "tma.cuh"

namespace reach {

namespace {

constexpr int MODE_UNCHANGED = 0;   // r*n >= p: reduce_order returns Z unchanged
constexpr int MODE_REDUCE = 1;
constexpr int NBW = 4;              // box-sum chunks per step for the W reduction (wseq adds them up per row: keep it small)
constexpr int NBR = 16;
constexpr int COPY_CTAS = 8;        // CTAs per step that record the references of the kept generators of R_k
constexpr int COPY_JB = 8;          // columns a warp resolves (one per lane) and then moves back to back
constexpr int NUM_WARPS = 4;        // warps per CTA of the TMA-staged numerics kernels (wbox_partial, apply_r)
constexpr int STAGE_BYTES = 16384;  // staging per warp: JB = min(8, STAGE_BYTES / column bytes) columns in flight
constexpr int MAX_IT = 16;          // double2 per lane of a row WINDOW: the numerics kernels sweep the rows 1024 at a time
constexpr int WIN_ROWS = 64 * MAX_IT;   // rows per window
constexpr int CHAIN_MAX_IT = 64;    // the sequential selection chain keeps per-row state in shared memory: n <= 64 * 64 = 4096
#ifndef REACH_CHAIN_THREADS
#define REACH_CHAIN_THREADS 512
#endif
// threads of the single selection-chain CTA (it owns an SM in the pipelined run): C2 chain time 10.8 ms with 256 threads,
// 9.7 ms with 512, 11.2 ms with 1024 (-DREACH_CHAIN_THREADS=... to rebuild)
constexpr int CHAIN_THREADS = REACH_CHAIN_THREADS;
constexpr int CHAIN_ROWS = 64 * CHAIN_MAX_IT / CHAIN_THREADS;   // rows per thread when a loop runs over the n rows

struct PlanW { int mode, pin, m, pW, pC, nk, nd, z; };     // W reduction of step k  (GLGM06/reach.jl:54-55)
struct PlanR { int mode, pin, m, pA, pW, pad0, pad1, pad2; };   // R reduction of step k  (:48-49)

struct __align__(16) WEnt {     // one generator of W in the sorted list: score, reference, logical position
    double key;
    int ref;
    int pos;
};

struct GP {
    int n, ld, p0, pV, q, strip, nwords, nwords4, record;
    int64_t nkeep, rn;          // n*(r-1), r*n
    int capW, capR, capListR, capListW, capDisc, smax;
    // column sharding (multi-GPU): this rank owns the generators [a0, a1) of G0 and [c0r, c1r) of GV;
    // its archive stores only those (la + lc columns) plus the two centre columns, qloc per slot.  Single GPU: all.
    int a0, a1, c0r, c1r, la, lc, qloc, rank, nranks;
    // ---- archive
    const double* Yall;         // slot j = Phi^j X, q columns each
    double* Wd;                 // [N+2][ld] box vectors of the W reductions (values of the box-generator references)
    uint32_t* maskAll;          // [N+1][pV][nwords] row-support masks of the mapped input generators (strip mode)
    // ---- current batch of steps [k0, k0 + cnt): step k uses slot j = k - 1, block-local index t = k - k0
    const double* Y;            // = Yall + (k0 - 1) q ld
    int64_t col0;               // archive column of Y's first column
    double* hS; int* rkS; int* cposS; double* skS; int* cntS;
    int* ordS; int* cpsS; int* cntZ;   // by sorted rank: element index, its compact position; # zero scores per segment
    // ---- W chain state (sorted list) + per-step snapshots of the batch
    WEnt* wEnt; int* wCount;
    WEnt* wEntAlt;              // the next W list while the parallel chain still reads the current one
    int* fastFlag;              // batch-local: 1 = this batch took the parallel selection chain (chain_fast_kernel)
    int* fastCount;             // [2]: batches that took the parallel chain / all batches (diagnostics)
    WEnt* snEnt; int* snCount;
    PlanW* planW; PlanR* planR;                 // [N+2]
    int* dlistW;                // [smax][capDisc] references of the generators boxed by the W reduction, sigma order
    int* tcnt; int* tsrc;       // [smax][n], [smax][n][4]: creation steps of the box generators boxed again, per row
    int* orderR;                // [smax][capListR] sortperm codes of the R list
    double* partW; double* partR;               // [smax][NB*][ld]
    double* cW;
    // ---- flowpipe
    // R_k = [ kept generators, as REFERENCES in sortperm order (Rref) | box generators d[row] e_row (Rd, Rrow) ]: the reach sets
    // are never copied into dense matrices by the time loop -- every generator already sits in the archive -- the dense
    // n x p matrix of a reach set is materialised by the fetch / projection kernels (materialize_kernel) on demand
    int* Rref; double* Rc; double* Rd; int* Rrow; int2* meta;
    // ---- selection records
    int* recOrdW; double* recHW; int* recOrdR; double* recHR;
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

// Box-generator reference of (step k, row i) and its decoding.
__device__ __forceinline__ int tag_ref(int k, int i, int n) { return -1 - (k * n + i); }
__device__ __forceinline__ void tag_decode(int ref, int n, int& k, int& i) {
    const int v = -1 - ref;
    k = v / n;
    i = v - k * n;
}
// Row i <-> (word, bit) of a row-support mask: the score kernel produces the bits with warp ballots over double2 lanes.
__device__ __forceinline__ void mask_pos(int i, int& word, int& bit) {
    word = 2 * (i >> 6) + (i & 1);
    bit = (i & 63) >> 1;
}

// Archive column of a (global) reference, or nullptr when another rank owns it.
__device__ __forceinline__ const double* col_ptr(const GP& g, int ref) {
    const int j = ref / g.q, c = ref - j * g.q;
    int l;
    if (c < g.p0) {
        if (c < g.a0 || c >= g.a1) return nullptr;
        l = c - g.a0;
    } else if (c < g.p0 + g.pV) {
        const int cc = c - g.p0;
        if (cc < g.c0r || cc >= g.c1r) return nullptr;
        l = g.la + cc - g.c0r;
    } else {
        l = g.la + g.lc + (c - g.p0 - g.pV);      // centres are replicated
    }
    return g.Yall + ((int64_t)j * g.qloc + l) * g.ld;
}

// h = |g|_1 - |g|_inf of every mapped generator of the batch (+inf = all-zero column dropped in strip mode) and, for the
// input segment, the row-support mask.  One warp per column; a pure function of the column's bits.
__global__ void __launch_bounds__(256) score_stack_kernel(GP g, int cnt) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int per = g.la + g.lc;                 // the generators this rank owns
    if (warp >= cnt * per) return;
    const int t = warp / per, lcol = warp % per;
    const int c = lcol < g.la ? g.a0 + lcol : g.p0 + g.c0r + (lcol - g.la);     // global column
    const double2* c2 = reinterpret_cast<const double2*>(g.Y + (int64_t)(t * g.qloc + lcol) * g.ld);
    const int n2 = g.ld / 2;
    double s = 0.0, m = 0.0;
    uint32_t* mask = (g.strip && c >= g.p0) ? g.maskAll + ((g.col0 / g.q + t) * (int64_t)g.pV + (c - g.p0)) * g.nwords4 : nullptr;
    for (int it = 0; it * 32 < n2; ++it) {
        const int i = lane + 32 * it;
        double2 v = make_double2(0.0, 0.0);
        if (i < n2) v = c2[i];
        const double ax = fabs(v.x), ay = fabs(v.y);
        s += ax;
        s += ay;
        m = fmax(m, fmax(ax, ay));
        if (mask) {
            const unsigned bx = __ballot_sync(0xffffffffu, ax != 0.0), by = __ballot_sync(0xffffffffu, ay != 0.0);
            if (lane == 0) { mask[2 * it] = bx; mask[2 * it + 1] = by; }
        }
    }
    const double asum = warp_sum(s), amax = warp_max(m);
    if (lane == 0) g.hS[t * g.q + c] = (g.strip && asum == 0.0) ? INFINITY : asum - amax;
}

// Stable order of every mapped segment (step t, seg 0 = P G0, seg 1 = P GV): rank by counting, O(len^2) but batched
// over the batch and spread over (segment, 256-element chunk) CTAs.
__global__ void __launch_bounds__(256) sort_segments_kernel(GP g) {
    __shared__ double keys[1024];
    const int t = blockIdx.y >> 1, seg = blockIdx.y & 1;
    const int len = seg ? g.pV : g.p0;
    if (blockIdx.x * 256 >= len && blockIdx.x != 0) return;
    const int base = t * g.q + (seg ? g.p0 : 0);
    const int e = blockIdx.x * 256 + threadIdx.x;
    const double he = e < len ? g.hS[base + e] : INFINITY;
    int less = 0, cpos = 0, nvalid = 0, nzero = 0;
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
            nzero += (hj == 0.0);
        }
    }
    if (e < len) {
        const bool valid = he < INFINITY;
        g.rkS[base + e] = valid ? less : -1;
        g.cposS[base + e] = cpos;
        if (valid) { g.skS[base + less] = he; g.ordS[base + less] = e; g.cpsS[base + less] = cpos; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { g.cntS[t * 2 + seg] = nvalid; g.cntZ[t * 2 + seg] = nzero; }
}

// Same result as sort_segments_kernel for segments of up to BITONIC_MAX generators: bitonic sort of (score, index) pairs in
// shared memory, one CTA per (step, segment).  O(len log^2 len) instead of O(len^2).
constexpr int BITONIC_MAX = 16384;      // 12 B per element of dynamic shared memory: 192 KB (n = 1024 has p0 = 4n + m + 1 = 4100 generators)
__global__ void __launch_bounds__(512) sort_segments_bitonic_kernel(GP g, int P) {
    extern __shared__ __align__(16) unsigned char bsm[];
    double* kk = reinterpret_cast<double*>(bsm);            // [P] scores
    int* ix = reinterpret_cast<int*>(kk + P);               // [P] indices
    __shared__ int wsum[16];
    const int t = blockIdx.x >> 1, seg = blockIdx.x & 1, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int len = seg ? g.pV : g.p0;
    const int base = t * g.q + (seg ? g.p0 : 0);
    for (int i = tid; i < P; i += 512) {
        kk[i] = i < len ? g.hS[base + i] : INFINITY;
        ix[i] = i;
    }
    __syncthreads();
    // compact position of every valid generator (exclusive scan of the valid flags in index order)
    {
        const int per = (len + 511) / 512;
        const int lo = tid * per, hi = min(len, lo + per);
        int c = 0;
        for (int i = lo; i < hi; ++i) c += kk[i] < INFINITY;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int v = lane < 16 ? wsum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += u;
            }
            if (lane < 16) wsum[lane] = v;
        }
        __syncthreads();
        int run = inc - c + (wid ? wsum[wid - 1] : 0);
        for (int i = lo; i < hi; ++i) {
            g.cposS[base + i] = run;
            run += kk[i] < INFINITY;
        }
        if (tid == 0) g.cntS[t * 2 + seg] = wsum[15];
    }
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int x = tid; x < P / 2; x += 512) {
                const int i = ((x & ~(j - 1)) << 1) | (x & (j - 1));      // index with bit j clear
                const int l = i | j;
                const double a = kk[i], b = kk[l];
                const int ia = ix[i], ib = ix[l];
                const bool up = (i & k) == 0;
                const bool gt = a > b || (a == b && ia > ib);
                if (gt == up) { kk[i] = b; kk[l] = a; ix[i] = ib; ix[l] = ia; }
            }
        }
    }
    __syncthreads();
    if (tid == 0) wsum[0] = 0;
    __syncthreads();
    int nz = 0;
    for (int pos = tid; pos < P; pos += 512) {
        const int e = ix[pos];
        if (e < len) {
            const double h = kk[pos];
            const bool valid = h < INFINITY;
            g.rkS[base + e] = valid ? pos : -1;
            if (valid) {
                g.skS[base + pos] = h;
                g.ordS[base + pos] = e;
                g.cpsS[base + pos] = g.cposS[base + e];      // written by this CTA before the sort
                nz += (h == 0.0);
            }
        }
    }
    if (nz) atomicAdd(&wsum[0], nz);
    __syncthreads();
    if (tid == 0) g.cntZ[t * 2 + seg] = wsum[0];
}

__device__ __forceinline__ int lower_bound_ent(const WEnt* a, int len, double key) {   // # entries with key < key
    int lo = 0, hi = len;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid].key < key) lo = mid + 1; else hi = mid;
    }
    return lo;
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

// W_1 = V (GLGM06/reach.jl:41-42): the valid columns of the input segment of slot 0, in their stable sorted order.
__global__ void __launch_bounds__(256) init_w_kernel(GP g) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) *g.wCount = g.cntS[1];
    if (e >= g.pV) return;
    const double h = g.hS[g.p0 + e];
    if (!(h < INFINITY)) return;
    WEnt w;
    w.key = h;
    w.ref = g.p0 + e;                       // archive column of slot 0
    w.pos = g.cposS[g.p0 + e];
    g.wEnt[g.rkS[g.p0 + e]] = w;
}

// -----------------------------------------------------------------------------------------------------------------
// CHAIN: the W selections of steps [k0, k0 + cnt), one persistent CTA, state in shared memory (or global scratch when
// it does not fit).  W list of step k = [W_{k-1} (logical order) | valid columns of P_{k-1} GV]; sigma = stable rank.
// -----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int upper_bound_ent(const WEnt* a, int len, double key) {   // # entries with key <= key
    int lo = 0, hi = len;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid].key <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

#ifdef REACH_CHAIN_PROFILE
__device__ unsigned long long g_chain_prof[8];
#define CHAIN_TICK(i) do { __syncthreads(); if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_chain_prof[i], (unsigned long long)(now_ - tick_)); tick_ = now_; } } while (0)
#else
#define CHAIN_TICK(i) do { } while (0)
#endif

template <bool SMEM, bool RECORD>
__global__ void __launch_bounds__(CHAIN_THREADS, CHAIN_THREADS <= 256 ? 3 : 1) chain_kernel(GP g, int k0, int cnt, double* scr) {
    extern __shared__ __align__(16) unsigned char chain_sm[];
    __shared__ unsigned orm[2 * CHAIN_MAX_IT];
    __shared__ int tc_s[64 * CHAIN_MAX_IT];    // per row: box generators boxed again in this step
    __shared__ int rows_s[CHAIN_ROWS * (CHAIN_THREADS / 32)];
    __shared__ int s_pC[128], s_zC[128];
    __shared__ int s_nd;
    const int tid = threadIdx.x, cap = g.capW, pV = g.pV, lane = tid & 31, wid = tid >> 5;
    if (g.fastFlag != nullptr && *g.fastFlag != 0) return;      // the parallel chain handled this batch (uniform over the CTA)
    // working set (shared memory, or global scratch when the W list is too long): the sorted list of W, double-buffered;
    // the input segment C_t of two consecutive steps in sorted order (score, element, compact position).
    WEnt* ebuf0 = SMEM ? reinterpret_cast<WEnt*>(chain_sm) : reinterpret_cast<WEnt*>(scr);
    WEnt* ebuf1 = ebuf0 + cap;
    const int pV1 = pV + 2;                                    // room for the +inf sentinel (kept even)
    double* cSk0 = reinterpret_cast<double*>(ebuf1 + cap);     // [2][pV1]
    int* cOrd0 = reinterpret_cast<int*>(cSk0 + 2 * pV1);       // [2][pV1]
    int* cCp0 = cOrd0 + 2 * pV1;                               // [2][pV1]
    int* dls = cCp0 + 2 * pV1;                                  // [capDisc] boxed list of the step, staged for a coalesced store
    if (tid < cnt && tid < 128) { s_pC[tid] = g.cntS[tid * 2 + 1]; s_zC[tid] = g.cntZ[tid * 2 + 1]; }
    int pW = *g.wCount;
    for (int r = tid; r < pW; r += CHAIN_THREADS) ebuf0[r] = g.wEnt[r];
    for (int e = tid; e < pV; e += CHAIN_THREADS) {
        const int b0 = g.p0 + e;
        cSk0[e] = g.skS[b0]; cOrd0[e] = g.ordS[b0]; cCp0[e] = g.cpsS[b0];
    }
    __syncthreads();
    // number of zero-score generators at the head of the sorted W (carried from step to step)
    int czw = 0;
    for (int lo = 0, hi = pW; ; ) {            // every thread runs the same search once per launch
        if (lo >= hi) { czw = lo; break; }
        const int mid = (lo + hi) >> 1;
        if (ebuf0[mid].key <= 0.0) lo = mid + 1; else hi = mid;
    }
    int cur = 0;
#ifdef REACH_CHAIN_PROFILE
    long long tick_ = clock64();
#endif
    for (int t = 0; t < cnt; ++t) {
        const int k = k0 + t;
        const int cb = t & 1;
        const WEnt* ec = cur ? ebuf1 : ebuf0;
        WEnt* en = cur ? ebuf0 : ebuf1;
        const double* skC = cSk0 + cb * pV1;
        const int* cOrd = cOrd0 + cb * pV1;
        const int* cCp = cCp0 + cb * pV1;
        const int baseC = t * g.q + g.p0;
        const int refC0 = (int)(g.col0 + baseC);
        const int pC = s_pC[t];
        // prefetch the next step's segment into registers (stored to the other buffer at the end of this step)
        double pfS[2] = {0.0, 0.0};
        int pfO[2] = {0, 0}, pfC[2] = {0, 0};
        if (t + 1 < cnt) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = tid + u * CHAIN_THREADS;
                if (e < pV) {
                    const int b1 = (t + 1) * g.q + g.p0 + e;
                    pfS[u] = g.skS[b1]; pfO[u] = g.ordS[b1]; pfC[u] = g.cpsS[b1];
                }
            }
        }
        const int pin = pW + pC;
        const int mode = (g.rn >= (int64_t)pin) ? MODE_UNCHANGED : MODE_REDUCE;
        const int m = mode == MODE_REDUCE ? (int)(pin - g.nkeep) : 0;
        const int nk = pin - m;
        const int z = max(0, czw + s_zC[t] - m);       // zero-score generators that survive the cut: they stay in front
        int* dl = g.dlistW + (int64_t)t * g.capDisc;
        CHAIN_TICK(0);
        // ---- A: the box generators of W (score 0, hence sigma = rank) that are boxed again ----
        if (tid < 2 * CHAIN_MAX_IT) orm[tid] = 0u;
        for (int i = tid; i < g.n; i += CHAIN_THREADS) tc_s[i] = 0;
        if (tid == 0) {             // sentinels for the merge of this step
            const_cast<WEnt*>(ec)[pW] = WEnt{INFINITY, 0, 0};
            const_cast<double*>(skC)[pC] = INFINITY;
            if (SMEM) bulk_wait_read0();    // the previous step's snapshot has left the buffer this step's merge overwrites
        }
        // every thread orders its generic-proxy writes of the list (previous step's merge) before the async-proxy read below
        if (SMEM) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (SMEM && tid == 0 && pW > 0) {
            // snapshot of W for the batched numerics: ONE bulk copy shared -> global (TMA engine) instead of 43 KB of store
            // instructions on the critical path; the list is read-only for the rest of the step
            bulk_s2g(g.snEnt + (int64_t)t * cap, ec, (uint32_t)pW * (uint32_t)sizeof(WEnt));
            bulk_commit();
        }
        for (int r = tid; r < min(czw, m); r += CHAIN_THREADS) {
            const int ref = ec[r].ref;
            if (ref < 0) {          // its value is an earlier W box vector entry: remember the creation step (wseq_kernel)
                int kt, it, w, b;
                tag_decode(ref, g.n, kt, it);
                const int sl = atomicAdd(&tc_s[it], 1);
                if (sl < 4) g.tsrc[((int64_t)t * g.n + it) * 4 + sl] = kt;
                if (g.strip) {
                    mask_pos(it, w, b);
                    atomicOr(&orm[w], 1u << b);
                }
            }
        }
        __syncthreads();
        CHAIN_TICK(1);
        if (g.strip && mode == MODE_REDUCE) {
            // Row-support union of the boxed generators.  The boxed box generators alone usually cover every row
            // (steady state); only otherwise are the dense ones located (binary search) and their masks fetched.
            if (wid == 0) {
                int c = 0;
                for (int w = lane; w < g.nwords; w += 32) c += __popc(orm[w]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                if (lane == 0) s_nd = c;
            }
            __syncthreads();
            const int covered = s_nd;
            __syncthreads();
            if (covered < g.n) {
                const int nw4 = g.nwords4 / 4;
                for (int x = tid; x < pW + pC; x += CHAIN_THREADS) {
                    const uint32_t* mk = nullptr;
                    if (x < pW) {
                        const WEnt w = ec[x];
                        if (w.ref >= 0 && x + lower_bound(skC, pC, w.key) < m) {
                            const int j = w.ref / g.q, c = w.ref - j * g.q - g.p0;
                            mk = g.maskAll + ((int64_t)j * pV + c) * g.nwords4;
                        }
                    } else {
                        const int jc = x - pW;
                        if (jc + upper_bound_ent(ec, pW, skC[jc]) < m)
                            mk = g.maskAll + ((g.col0 / g.q + t) * (int64_t)pV + cOrd[jc]) * g.nwords4;
                    }
                    if (mk) {
#pragma unroll 4
                        for (int u = 0; u < nw4; ++u) {
                            const uint4 v = reinterpret_cast<const uint4*>(mk)[u];
                            if (v.x) atomicOr(&orm[4 * u + 0], v.x);
                            if (v.y) atomicOr(&orm[4 * u + 1], v.y);
                            if (v.z) atomicOr(&orm[4 * u + 2], v.z);
                            if (v.w) atomicOr(&orm[4 * u + 3], v.w);
                        }
                    }
                }
                __syncthreads();
            }
        }
        for (int i = tid; i < g.n; i += CHAIN_THREADS) g.tcnt[(int64_t)t * g.n + i] = tc_s[i];
        CHAIN_TICK(2);
        // ---- B: rows of the new box generators (all rows, or the non-zero rows of the box vector in strip mode):
        //      thread tid owns rows tid, tid + T, ...; iprime = number of flagged rows below, nd = their total ----
        int flag[CHAIN_ROWS], iprime[CHAIN_ROWS], nd = 0;
#pragma unroll
        for (int j = 0; j < CHAIN_ROWS; ++j) { flag[j] = 0; iprime[j] = 0; }
        if (mode == MODE_REDUCE) {
            unsigned bal[CHAIN_ROWS];
#pragma unroll
            for (int j = 0; j < CHAIN_ROWS; ++j) {
                const int i = tid + j * CHAIN_THREADS;
                if (i < g.n) {
                    if (g.strip) {
                        int w, b;
                        mask_pos(i, w, b);
                        flag[j] = (orm[w] >> b) & 1;
                    } else {
                        flag[j] = 1;
                    }
                }
                bal[j] = __ballot_sync(0xffffffffu, flag[j]);
                if (lane == 0) rows_s[j * (CHAIN_THREADS / 32) + wid] = __popc(bal[j]);
            }
            __syncthreads();
            if (wid == 0) {         // exclusive scan over the CHAIN_ROWS * warps counts, row-major = ascending rows
                int run = 0;
                for (int x0 = 0; x0 < CHAIN_ROWS * (CHAIN_THREADS / 32); x0 += 32) {
                    const int x = x0 + lane;
                    const int c = x < CHAIN_ROWS * (CHAIN_THREADS / 32) ? rows_s[x] : 0;
                    int v = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int u = __shfl_up_sync(0xffffffffu, v, o);
                        if (lane >= o) v += u;
                    }
                    if (x < CHAIN_ROWS * (CHAIN_THREADS / 32)) rows_s[x] = run + v - c;
                    run += __shfl_sync(0xffffffffu, v, 31);
                }
                if (lane == 0) s_nd = run;
            }
            __syncthreads();
            nd = s_nd;
#pragma unroll
            for (int j = 0; j < CHAIN_ROWS; ++j)
                iprime[j] = rows_s[j * (CHAIN_THREADS / 32) + wid] + __popc(bal[j] & ((1u << lane) - 1u));
        }
        CHAIN_TICK(3);
        // ---- C: MERGE PATH.  The stable sortperm of [W | C_t] is the merge of two sorted lists (ties: W first, it
        //      precedes C in the list).  Thread tid produces the sigma range [tid L, (tid+1) L): a binary search along its
        //      diagonal finds how many entries of W / C precede it, then L sequential merge steps emit, per entry: the
        //      snapshot of W_{k-1} (for the R reduction), the boxed list, and the next W. ----
        WEnt* sn = g.snEnt + (int64_t)t * cap;
        int* recO = RECORD ? g.recOrdW + (int64_t)k * g.capListW : nullptr;
        double* recH = RECORD ? g.recHW + (int64_t)k * g.capListW : nullptr;
        if (!SMEM) {
#pragma unroll 4
            for (int r = tid; r < pW; r += CHAIN_THREADS) sn[r] = ec[r];   // coalesced; the merge below touches only the list
        }
        {
            const int L = ((pin + CHAIN_THREADS - 1) / CHAIN_THREADS) | 1;     // odd: the 16-byte stores of a quarter-warp hit distinct banks
            const int d0 = min(pin, tid * L), d1 = min(pin, d0 + L);
            // i = # entries of W among the first d0 merged: largest i with (i == 0 || W[i-1] <= C[d0-i])  (W wins ties)
            int lo = max(0, d0 - pC), hi = min(d0, pW);
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;               // candidate i = mid: needs W[mid-1] <= C[d0-mid]
                const int jc = d0 - mid;
                const bool ok = jc >= pC || ec[mid - 1].key <= skC[jc];
                if (ok) lo = mid; else hi = mid - 1;
            }
            int i = lo, j = d0 - lo;
            CHAIN_TICK(6);
            // sentinels (+inf) close both lists, so the loop body needs no bounds checks
            const WEnt* wp = ec + i;
            WEnt wi = *wp;
            double cj = skC[j];
            for (int sigma = d0; sigma < d1; ++sigma) {
                WEnt w;
                if (wi.key <= cj) {          // W precedes C in the list: ties go to W
                    w = wi;
                    wi = *++wp;
                } else {
                    w.key = cj; w.ref = refC0 + cOrd[j]; w.pos = pW + cCp[j];
                    cj = skC[++j];
                }
                if (RECORD) { recO[sigma] = w.pos; recH[w.pos] = w.key; }
                if (mode == MODE_REDUCE) {
                    if (sigma < m) {
                        dls[sigma] = w.ref;
                    } else {
                        const int ps = sigma - m;
                        w.pos = ps;
                        en[ps < z ? ps : ps + nd] = w;
                    }
                } else {
                    en[sigma] = w;
                }
            }
            CHAIN_TICK(7);
            __syncthreads();
            for (int x = tid; x < m; x += CHAIN_THREADS) dl[x] = dls[x];
        }
#pragma unroll
        for (int j = 0; j < CHAIN_ROWS; ++j)
            if (flag[j]) {
                WEnt w;
                w.key = 0.0; w.ref = tag_ref(k, tid + j * CHAIN_THREADS, g.n); w.pos = nk + iprime[j];
                en[z + iprime[j]] = w;
            }
        if (tid == 0) {
            g.snCount[t] = pW;
            PlanW pl;
            pl.mode = mode; pl.pin = pin; pl.m = m; pl.pW = pW; pl.pC = pC; pl.nk = nk; pl.nd = nd; pl.z = z;
            g.planW[k] = pl;
        }
        CHAIN_TICK(4);
        czw = mode == MODE_REDUCE ? z + nd : czw + s_zC[t];
        pW = mode == MODE_REDUCE ? nk + nd : pin;
        cur ^= 1;
        if (t + 1 < cnt) {
            double* nS = cSk0 + (cb ^ 1) * pV1;
            int* nO = cOrd0 + (cb ^ 1) * pV1;
            int* nC = cCp0 + (cb ^ 1) * pV1;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = tid + u * CHAIN_THREADS;
                if (e < pV) { nS[e] = pfS[u]; nO[e] = pfO[u]; nC[e] = pfC[u]; }
            }
            for (int e = tid + 2 * CHAIN_THREADS; e < pV; e += CHAIN_THREADS) {
                const int b1 = (t + 1) * g.q + g.p0 + e;
                nS[e] = g.skS[b1]; nO[e] = g.ordS[b1]; nC[e] = g.cpsS[b1];
            }
        }
        __syncthreads();
        CHAIN_TICK(5);
    }
    const WEnt* ef = cur ? ebuf1 : ebuf0;
    for (int r = tid; r < pW; r += CHAIN_THREADS) g.wEnt[r] = ef[r];
    if (SMEM && tid == 0) bulk_wait0();       // the snapshots are in global memory before the kernel ends
    if (tid == 0) *g.wCount = pW;
}

// -----------------------------------------------------------------------------------------------------------------
// PARALLEL CHAIN (steady state).  Once W is full -- n(r-1) dense generators plus the n box generators of the previous step --
// and no mapped generator has score exactly 0, every step boxes the n box generators (score 0, always first) and the pC_t
// lowest dense candidates of D_t ∪ C_t (D_t: the dense generators of W_{k-1}, C_t: the valid columns of P_{k-1} GV).  Under
// Julia's stable sortperm the candidates are totally ordered by (score, age, column), ages never change, so whether and
// when a generator is boxed is a pure function of counts:
//     L_t(e)   = #candidates of step t that sort before e  = pos_t(e) + #{x in C_t : score_x < score_e}      (e in D_t)
//     boxed at t  <=>  L_t(e) < pC_t ;  otherwise  pos_{t+1}(e) = L_t(e) - pC_t.
// One warp per generator (the dense part of W at the start of the batch and every mapped column of the batch) evaluates its
// whole life over the batch -- a binary search per step, a warp prefix sum -- and writes what the batched numerics read: its
// entry in every per-step snapshot of W, its slot in the boxed list of the step that boxes it, and its place in the next W.
// No step waits for another: 64 dependent merges on one CTA (3.9 us each at C2) become two launches.
// chain_decide_kernel checks the preconditions on the device (the host never synchronises); when they do not hold (growth
// phase, zero scores, zero-generator removal still shrinking the box) the sequential chain_kernel below runs instead.
// -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) chain_decide_kernel(GP g, int k0, int cnt, int force_sequential) {
    const int lane = threadIdx.x, n = g.n, nkeep = (int)g.nkeep;
    int ok = !force_sequential && nkeep >= 1 && cnt <= 128 && *g.wCount == (int)g.rn;
    if (ok) {
        for (int i = lane; i < n; i += 32) {
            const WEnt w = g.wEnt[i];
            ok &= (w.ref == tag_ref(k0 - 1, i, n)) && (w.key == 0.0) && (w.pos == nkeep + i);
        }
        if (lane == 0) {
            const WEnt w = g.wEnt[n];                   // the lowest dense generator: a genuine generator with a positive score
            ok &= (w.ref >= 0) && (w.key > 0.0);
        }
        for (int t = lane; t < cnt; t += 32) ok &= (g.cntS[2 * t + 1] >= 1) && (g.cntZ[2 * t + 1] == 0);
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        *g.fastFlag = ok;
        g.fastCount[0] += ok;
        g.fastCount[1] += 1;
    }
}

template <bool RECORD>
__global__ void __launch_bounds__(256) chain_fast_kernel(GP g, int k0, int cnt) {
    if (*g.fastFlag == 0) return;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n = g.n, nkeep = (int)g.nkeep, pV = g.pV, rn = (int)g.rn;
    if (warp >= nkeep + cnt * pV) return;
    int b = -1, j = 0, cp = 0, ref;
    double key;
    if (warp < nkeep) {
        const WEnt w = g.wEnt[n + warp];
        key = w.key; ref = w.ref;
    } else {
        const int x = warp - nkeep;
        b = x / pV; j = x - b * pV;
        if (j >= g.cntS[2 * b + 1]) return;
        const int base = b * g.q + g.p0;
        key = g.skS[base + j]; ref = (int)(g.col0 + base + g.ordS[base + j]); cp = g.cpsS[base + j];
    }
    constexpr int U = 4;                                 // steps per lane: cnt <= 128
    int delta[U], pc[U], incl[U];
    int older_greater = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int t = lane + 32 * u;
        delta[u] = 0; pc[u] = 0;
        if (t < cnt) {
            pc[u] = g.cntS[2 * t + 1];
            const double* sk = g.skS + t * g.q + g.p0;
            if (t > b) delta[u] = lower_bound(sk, pc[u], key) - pc[u];         // younger columns sort after e on a tie
            else if (t < b) older_greater += pc[u] - upper_bound(sk, pc[u], key);   // older columns sort before e on a tie
        }
    }
    // position from which the recurrence starts: P0 = pos at step t0
    int t0 = 0, P0 = warp, entry_L = 0;
    if (b >= 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) older_greater += __shfl_xor_sync(0xffffffffu, older_greater, o);
        const int G = older_greater + (nkeep - upper_bound_ent(g.wEnt + n, nkeep, key));   // older generators that sort after e
        entry_L = (nkeep - min(nkeep, G)) + j;           // candidates of step b that sort before e
        t0 = b + 1;
        int pcb = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int v = __shfl_sync(0xffffffffu, pc[u], b & 31);
            if (u == (b >> 5)) pcb = v;
        }
        P0 = entry_L - pcb;                              // pos_{b+1}(e) if it survives its first step
    }
    // inclusive prefix sums of delta over t = 0 .. cnt-1 (t = lane + 32 u: lanes first, then u)
    int carry = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        int v = delta[u];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += x;
        }
        incl[u] = carry + v;
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
    // first step that boxes e
    int death = INT_MAX;
    if (b >= 0 && P0 < 0) death = b;                     // boxed by the very reduction it enters
    else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = lane + 32 * u;
            if (t < cnt && t >= t0 && P0 + incl[u] < 0) death = min(death, t);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) death = min(death, __shfl_xor_sync(0xffffffffu, death, o));
    }
    if (b >= 0 && lane == 0) {                           // the step the column enters: its place in the merged order
        if (death == b) g.dlistW[(int64_t)b * g.capDisc + n + entry_L] = ref;
        if (RECORD) {
            g.recOrdW[(int64_t)(k0 + b) * g.capListW + n + entry_L] = rn + cp;
            g.recHW[(int64_t)(k0 + b) * g.capListW + rn + cp] = key;
        }
    }
    if (death != b || b < 0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = lane + 32 * u;
            if (t < cnt && t >= t0 && t <= death) {
                const int pos = P0 + incl[u] - delta[u];     // position among the dense generators of W_{k-1}, ascending
                WEnt w;
                w.key = key; w.ref = ref; w.pos = pos;
                g.snEnt[(int64_t)t * g.capW + n + pos] = w;
                const int L = pos + delta[u] + pc[u];
                if (t == death) g.dlistW[(int64_t)t * g.capDisc + n + L] = ref;
                if (RECORD) {
                    g.recOrdW[(int64_t)(k0 + t) * g.capListW + n + L] = pos;
                    g.recHW[(int64_t)(k0 + t) * g.capListW + pos] = key;
                }
            }
        }
        if (death == INT_MAX && lane == 0) {             // survives the batch: its place in the next W
            WEnt w;
            w.key = key; w.ref = ref; w.pos = P0 + carry;
            g.wEntAlt[n + w.pos] = w;
        }
    }
}

// everything of the parallel chain that does not depend on scores: the box generators (references to the previous step's box
// vector), the plans, and the hand-over of the next W
template <bool RECORD>
__global__ void __launch_bounds__(256) chain_fast_fill_kernel(GP g, int k0, int cnt) {
    if (*g.fastFlag == 0) return;
    const int n = g.n, nkeep = (int)g.nkeep, rn = (int)g.rn;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.y < (unsigned)cnt) {
        const int t = blockIdx.y, k = k0 + t;
        if (gid < n) {
            const int i = gid;
            WEnt w;
            w.key = 0.0; w.ref = tag_ref(k - 1, i, n); w.pos = nkeep + i;
            g.snEnt[(int64_t)t * g.capW + i] = w;
            g.dlistW[(int64_t)t * g.capDisc + i] = w.ref;
            g.tcnt[(int64_t)t * n + i] = 1;
            g.tsrc[((int64_t)t * n + i) * 4] = k - 1;
            if (RECORD) {
                g.recOrdW[(int64_t)k * g.capListW + i] = nkeep + i;
                g.recHW[(int64_t)k * g.capListW + nkeep + i] = 0.0;
            }
        }
        if (gid == 0) {
            const int pC = g.cntS[2 * t + 1];
            g.snCount[t] = rn;
            PlanW pl;
            pl.mode = MODE_REDUCE; pl.pin = rn + pC; pl.m = n + pC; pl.pW = rn; pl.pC = pC; pl.nk = nkeep; pl.nd = n; pl.z = 0;
            g.planW[k] = pl;
        }
    } else {                                             // last row of the grid: the next W (box generators of the last step first)
        const int klast = k0 + cnt - 1;
        if (gid < n) {
            WEnt w;
            w.key = 0.0; w.ref = tag_ref(klast, gid, n); w.pos = nkeep + gid;
            g.wEnt[gid] = w;
        } else if (gid < rn) {
            g.wEnt[gid] = g.wEntAlt[gid];
        }
    }
}

static void launch_chain(GP& g, int k0, int cnt, bool use_smem, size_t smem, double* scr, cudaStream_t st,
                         bool exclusive = false) {
    // exclusive: ask for so much shared memory that neither a GEMM CTA nor a numerics CTA fits next to the chain -- it
    // then owns the SM the GEMM leaves free (Ctx::gemm_reserve_sms) and runs at its stand-alone speed
    if (use_smem && exclusive) smem = std::max<size_t>(smem, 200 * 1024);
    if (use_smem) {
        if (g.record) chain_kernel<true, true><<<1, CHAIN_THREADS, smem, st>>>(g, k0, cnt, scr);
        else chain_kernel<true, false><<<1, CHAIN_THREADS, smem, st>>>(g, k0, cnt, scr);
    } else {
        if (g.record) chain_kernel<false, true><<<1, CHAIN_THREADS, 0, st>>>(g, k0, cnt, scr);
        else chain_kernel<false, false><<<1, CHAIN_THREADS, 0, st>>>(g, k0, cnt, scr);
    }
}

// Selection chain of steps [k0, k0 + cnt): the parallel chain when its preconditions hold (decided on the device), else the
// sequential one -- both are enqueued, exactly one of them works.
static int launch_chain_auto(GP& g, int k0, int cnt, bool use_smem, size_t smem, double* scr, cudaStream_t st, bool exclusive) {
    const bool force_seq = getenv("REACH_CHAIN_SEQUENTIAL") != nullptr;      // diagnostics / A-B tests
    chain_decide_kernel<<<1, 32, 0, st>>>(g, k0, cnt, force_seq ? 1 : 0);
    int launches = 1;
    if (!force_seq && cnt <= 128 && g.nkeep >= 1) {
        const int64_t warps = g.nkeep + (int64_t)cnt * g.pV;
        const dim3 fill_grid((unsigned)ceil_div(g.rn, 256), (unsigned)(cnt + 1));
        if (g.record) {
            chain_fast_kernel<true><<<(unsigned)ceil_div(warps, 8), 256, 0, st>>>(g, k0, cnt);
            chain_fast_fill_kernel<true><<<fill_grid, 256, 0, st>>>(g, k0, cnt);
        } else {
            chain_fast_kernel<false><<<(unsigned)ceil_div(warps, 8), 256, 0, st>>>(g, k0, cnt);
            chain_fast_fill_kernel<false><<<fill_grid, 256, 0, st>>>(g, k0, cnt);
        }
        launches += 2;
    }
    launch_chain(g, k0, cnt, use_smem, smem, scr, st, exclusive);
    return launches + 1;
}

// -----------------------------------------------------------------------------------------------------------------
// batched numerics
// -----------------------------------------------------------------------------------------------------------------
template <int IT>
struct Acc {        // IT double2 per lane: rows <= 64 * IT.  Small IT keeps the numerics kernels co-resident with the GEMM.
    double2 v[IT];
};

template <int IT>
__device__ __forceinline__ void acc_zero(Acc<IT>& a) {
#pragma unroll
    for (int i = 0; i < IT; ++i) a.v[i] = make_double2(0.0, 0.0);
}
// The rows are swept in WINDOWS of WIN_ROWS (one window for n <= 1024): an accumulator covers rows [rbase, rbase + wl).
struct Win {
    int rbase, wl;
};
__device__ __forceinline__ int win_pitch(const GP& g) { return min(g.ld, WIN_ROWS); }     // column pitch of the shared-memory stages
__device__ __forceinline__ Win win_at(const GP& g, int rbase) { return Win{rbase, min(WIN_ROWS, g.ld - rbase)}; }
// CTA-wide sum of the warps' accumulators in warp order -> out[0 .. ld)
template <int IT>
__device__ __forceinline__ void acc_reduce_store(const GP& g, const Acc<IT>& a, double* red_s, double* __restrict__ out, const Win& win) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n2 = win.wl / 2, nw = blockDim.x >> 5, pitch = win_pitch(g);
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        const int idx = lane + 32 * i;
        if (idx < n2) {
            red_s[warp * pitch + 2 * idx] = a.v[i].x;
            red_s[warp * pitch + 2 * idx + 1] = a.v[i].y;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < win.wl; i += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += red_s[w * pitch + i];
        out[win.rbase + i] = s;
    }
    __syncthreads();        // the scratch is free for the next window
}

// ---- TMA-staged column access ------------------------------------------------------------------------------------
// The numerics kernels run next to the persistent DMMA GEMM of a later batch, which leaves room for ONE small CTA per
// SM.  With register-staged loads (one 2 KB column per warp in flight) that CTA could not cover the HBM latency
// (measured: a 64-step batch took 1.8 ms in situ against 0.24 ms alone).  Here every warp owns a shared-memory stage of
// JB columns; one lane per column issues a bulk copy (TMA engine, SASS UBLKCP) and the warp waits on an mbarrier, so a
// 4-warp CTA keeps ~70 KB in flight with no registers held.
struct WarpStage {
    double* buf;        // JB columns of ld doubles
    uint64_t* bar;
    uint32_t phase;
};
__device__ __forceinline__ WarpStage warp_stage_init(const GP& g, unsigned char* smem, int jb) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpStage ws;
    ws.buf = reinterpret_cast<double*>(smem) + (size_t)warp * jb * win_pitch(g);
    ws.bar = reinterpret_cast<uint64_t*>(reinterpret_cast<double*>(smem) + (size_t)NUM_WARPS * jb * win_pitch(g)) + warp;
    ws.phase = 0;
    if (lane == 0) mbar_init(ws.bar, 1);
    mbar_fence_init();
    __syncthreads();
    return ws;
}
// scratch of acc_reduce_store behind the stages and barriers
__device__ __forceinline__ double* stage_red(const GP& g, unsigned char* smem, int jb) {
    return reinterpret_cast<double*>(smem) + (size_t)NUM_WARPS * jb * win_pitch(g) + NUM_WARPS;
}
// Lane j < jb holds column pointer `src` (or nullptr): start the bulk loads of the warp's group.  Returns the ballot.
__device__ __forceinline__ unsigned stage_load(const GP& g, WarpStage& ws, const double* src, int lane, const Win& win) {
    const unsigned dense = __ballot_sync(0xffffffffu, src != nullptr);
    if (dense) {
        if (lane == 0) mbar_expect_tx(ws.bar, (uint32_t)(__popc(dense) * win.wl * 8));
        __syncwarp();
        if (src) bulk_g2s(ws.buf + (size_t)lane * win_pitch(g), src + win.rbase, (uint32_t)(win.wl * 8), ws.bar);
    }
    return dense;
}
__device__ __forceinline__ void stage_wait(WarpStage& ws) {
    mbar_wait(ws.bar, ws.phase);
    ws.phase ^= 1;
}
// a += |staged column j| for every j of `dense`, in lane order
template <int IT>
__device__ __forceinline__ void acc_add_staged(const GP& g, Acc<IT>& a, const WarpStage& ws, unsigned dense, int lane, const Win& win) {
    const int n2 = win.wl / 2;
    for (unsigned rest = dense; rest;) {
        const int j = __ffs(rest) - 1;
        rest &= rest - 1;
        const double2* c2 = reinterpret_cast<const double2*>(ws.buf + (size_t)j * win_pitch(g));
#pragma unroll
        for (int i = 0; i < IT; ++i) {
            const int idx = lane + 32 * i;
            if (idx < n2) {
                const double2 v = c2[idx];
                a.v[i].x += fabs(v.x);
                a.v[i].y += fabs(v.y);
            }
        }
    }
    __syncwarp();           // every lane is done reading before the next bulk load lands in the stage
}
// a += |box generator (tval at row trow)| of the lanes in `tags`, in lane order
template <int IT>
__device__ __forceinline__ void acc_add_tags(Acc<IT>& a, unsigned tags, double tval, int trow, int lane, const Win& win) {
    for (unsigned rest = tags; rest;) {
        const int j = __ffs(rest) - 1;
        rest &= rest - 1;
        const double val = fabs(__shfl_sync(0xffffffffu, tval, j));
        const int it = __shfl_sync(0xffffffffu, trow, j) - win.rbase;
        if (it < 0 || it >= win.wl) continue;           // the box generator's row lies in another window
        const int idx = it >> 1;
#pragma unroll
        for (int i = 0; i < IT; ++i)
            if (lane + 32 * i == idx) {
                if (it & 1) a.v[i].y += val; else a.v[i].x += val;
            }
    }
}

// Partial box vectors of the W reductions: chunk b of step t sums |dense discarded generators| in sigma order.
template <int IT>
__global__ void __launch_bounds__(NUM_WARPS * 32) wbox_partial_kernel(GP g, int k0, int jb) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int t = blockIdx.y, b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PlanW pl = g.planW[k0 + t];
    double* out = g.partW + ((int64_t)t * NBW + b) * g.ld;
    WarpStage ws = warp_stage_init(g, smem_raw, jb);
    for (int rbase = 0; rbase < g.ld; rbase += WIN_ROWS) {
        const Win win = win_at(g, rbase);
        Acc<IT> a;
        acc_zero(a);
        if (pl.mode == MODE_REDUCE) {
            const int chunk = (pl.m + NBW - 1) / NBW;
            const int lo = b * chunk, hi = min(pl.m, lo + chunk);
            const int* dl = g.dlistW + (int64_t)t * g.capDisc;
            const int wchunk = (chunk + NUM_WARPS - 1) / NUM_WARPS;
            const int wlo = lo + warp * wchunk, whi = min(hi, wlo + wchunk);
            for (int s0 = wlo; s0 < whi; s0 += jb) {
                const int ref = (lane < jb && s0 + lane < whi) ? dl[s0 + lane] : -1;     // box generators are wseq's business
                const double* src = ref >= 0 ? col_ptr(g, ref) : nullptr;   // nullptr: another rank sums it (partials are all-reduced)
                const unsigned dense = stage_load(g, ws, src, lane, win);
                if (dense) {
                    stage_wait(ws);
                    acc_add_staged(g, a, ws, dense, lane, win);
                }
            }
        }
        __syncthreads();
        acc_reduce_store(g, a, stage_red(g, smem_raw, jb), out, win);
    }
}

// Sequential over the steps of the batch, parallel over rows: W box vectors (dense partials + discarded box
// generators, whose values are earlier W box vectors) and the centres  c_k = P c0 + cW ; cW += P cV  (:48,:54).
// Software-pipelined: everything of step t+1 that does not depend on the recurrence is loaded during step t.
struct WseqIn {
    int mode, tc, m;
    int ks[4];
    double base, yc0, ycV;
};
__device__ __forceinline__ WseqIn wseq_load(const GP& g, int k0, int t, int i) {
    WseqIn w;
    const PlanW pl = g.planW[k0 + t];
    w.mode = pl.mode; w.m = pl.m;
    const double* yc0 = g.Y + (int64_t)(t * g.qloc + g.la + g.lc) * g.ld;
    w.yc0 = yc0[i];
    w.ycV = yc0[g.ld + i];
    w.base = 0.0; w.tc = 0;
    w.ks[0] = w.ks[1] = w.ks[2] = w.ks[3] = 0;
    if (pl.mode == MODE_REDUCE) {
        for (int b = 0; b < NBW; ++b) w.base += g.partW[((int64_t)t * NBW + b) * g.ld + i];
        w.tc = g.tcnt[(int64_t)t * g.n + i];
        const int4 kk = *reinterpret_cast<const int4*>(g.tsrc + ((int64_t)t * g.n + i) * 4);
        w.ks[0] = kk.x; w.ks[1] = kk.y; w.ks[2] = kk.z; w.ks[3] = kk.w;
    }
    return w;
}
// Steady state (the parallel chain handled the batch): every step boxes exactly the n box generators of the previous step, so
// d_k = base_k + |d_{k-1}| and the recurrences need no indirection.  One WARP per row, lanes = steps: the loads of 32 steps are
// issued at once (one round trip to memory instead of one per step), the sums then run lane by lane in the sequential
// kernel's order (identical bits).
__global__ void __launch_bounds__(256) wseq_fast_kernel(GP g, int k0, int cnt) {
    if (*g.fastFlag == 0) return;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= g.n) return;
    double cw = g.cW[i];
    double d = g.Wd[(int64_t)(k0 - 1) * g.ld + i];
    for (int t0 = 0; t0 < cnt; t0 += 32) {
        const int t = t0 + lane;
        double base = 0.0, yc0 = 0.0, ycV = 0.0;
        if (t < cnt) {
            double pb[NBW];
#pragma unroll
            for (int b = 0; b < NBW; ++b) pb[b] = g.partW[((int64_t)t * NBW + b) * g.ld + i];
            const double* yc = g.Y + (int64_t)(t * g.qloc + g.la + g.lc) * g.ld;
            yc0 = yc[i];
            ycV = yc[g.ld + i];
#pragma unroll
            for (int b = 0; b < NBW; ++b) base += pb[b];
        }
        double my_c = 0.0, my_d = 0.0;
        const int m = min(32, cnt - t0);
        for (int u = 0; u < m; ++u) {                       // the recurrences, in step order
            const double bu = __shfl_sync(0xffffffffu, base, u), y0 = __shfl_sync(0xffffffffu, yc0, u),
                         yV = __shfl_sync(0xffffffffu, ycV, u);
            const double cu = y0 + cw;
            cw += yV;
            d = bu + fabs(d);
            if (lane == u) { my_c = cu; my_d = d; }
        }
        if (t < cnt) {
            const int k = k0 + t;
            g.Rc[(int64_t)(k - 1) * g.ld + i] = my_c;
            g.Wd[(int64_t)k * g.ld + i] = my_d;
        }
    }
    if (lane == 0) g.cW[i] = cw;
}

__global__ void __launch_bounds__(128) wseq_kernel(GP g, int k0, int cnt) {
    if (g.fastFlag != nullptr && *g.fastFlag != 0) return;     // wseq_fast_kernel handled this batch
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    double cw = g.cW[i];
    double dprev = 0.0;
    int kprev = -1;
    WseqIn w = wseq_load(g, k0, 0, i);
    for (int t = 0; t < cnt; ++t) {
        const int k = k0 + t;
        WseqIn wn = w;
        if (t + 1 < cnt) wn = wseq_load(g, k0, t + 1, i);
        g.Rc[(int64_t)(k - 1) * g.ld + i] = w.yc0 + cw;
        cw += w.ycV;
        if (w.mode == MODE_REDUCE) {
            double d = w.base;
            if (w.tc <= 4) {          // creation steps ascending => fixed summation order
                int ks[4] = {w.ks[0], w.ks[1], w.ks[2], w.ks[3]};
                for (int u = 1; u < w.tc; ++u)
                    for (int v = u; v > 0 && ks[v - 1] > ks[v]; --v) { const int x = ks[v]; ks[v] = ks[v - 1]; ks[v - 1] = x; }
                for (int u = 0; u < w.tc; ++u)
                    d += fabs(ks[u] == kprev ? dprev : g.Wd[(int64_t)ks[u] * g.ld + i]);
            } else {                // rare: more than four box generators of one row boxed in the same step
                const int* dl = g.dlistW + (int64_t)t * g.capDisc;
                for (int s = 0; s < w.m; ++s) {
                    const int ref = dl[s];
                    if (ref < 0) {
                        int kt, it;
                        tag_decode(ref, g.n, kt, it);
                        if (it == i) d += fabs(g.Wd[(int64_t)kt * g.ld + i]);
                    }
                }
            }
            g.Wd[(int64_t)k * g.ld + i] = d;
            dprev = d;
            kprev = k;
        }
        w = wn;
    }
    g.cW[i] = cw;
}

// sortperm of the R lists [A = P_{k-1} G0 | W_{k-1}] of the batch by merging sorted segments (GLGM06/reach.jl:48-49).
// orderR[t][sigma] = code: e < p0 -> mapped column e of the step; p0 + rho -> entry of sorted rank rho of the snapshot.
__global__ void __launch_bounds__(256) rank_r_kernel(GP g, int k0) {
    const int t = blockIdx.y, k = k0 + t;
    const int pW = g.snCount[t], pA = g.cntS[t * 2 + 0];
    const int baseA = t * g.q;
    const double* skA = g.skS + baseA;
    const WEnt* sn = g.snEnt + (int64_t)t * g.capW;
    const int pin = pA + pW;
    const int mode = (g.rn >= (int64_t)pin) ? MODE_UNCHANGED : MODE_REDUCE;
    int* order = g.orderR + (int64_t)t * g.capListR;
    int* recO = g.record ? g.recOrdR + (int64_t)k * g.capListR : nullptr;
    double* recH = g.record ? g.recHR + (int64_t)k * g.capListR : nullptr;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid == 0) {
        PlanR pl;
        pl.mode = mode; pl.pin = pin; pl.m = mode == MODE_REDUCE ? (int)(pin - g.nkeep) : 0; pl.pA = pA; pl.pW = pW;
        pl.pad0 = pl.pad1 = pl.pad2 = 0;
        g.planR[k] = pl;
    }
    if (gid < g.p0) {
        const double h = g.hS[baseA + gid];
        if (h < INFINITY) {
            const int sigma = g.rkS[baseA + gid] + lower_bound_ent(sn, pW, h);
            order[sigma] = gid;
            if (recO) { const int li = g.cposS[baseA + gid]; recO[sigma] = li; recH[li] = h; }
        }
    } else if (gid - g.p0 < pW) {
        const int rho = gid - g.p0;
        const double h = sn[rho].key;
        const int sigma = rho + upper_bound(skA, pA, h);
        order[sigma] = g.p0 + rho;
        if (recO) { const int li = pA + sn[rho].pos; recO[sigma] = li; recH[li] = h; }
    }
}

// R gather of the batch: blockIdx.x < NBR accumulate the boxed generators (chunks of sigma), the other CTAs move the
// kept ones into the flowpipe slot, warp per column.
constexpr int NO_JOB = INT_MIN;

// (reference, destination column) of copy job `job` of step t, or NO_JOB
__device__ __forceinline__ void resolve_copy_job(const GP& g, const PlanR& pl, const int* order, const WEnt* sn, int t, int job,
                                                 int jobs, int& ref, int& dest) {
    ref = NO_JOB; dest = 0;
    if (job >= jobs) return;
    if (pl.mode == MODE_REDUCE) {
        if (job < pl.m) return;
        const int code = order[job];
        ref = code < g.p0 ? (int)(g.col0 + t * g.q + code) : sn[code - g.p0].ref;
        dest = job - pl.m;
    } else if (job < g.p0) {          // unchanged: [valid columns of A in order | W in logical order]
        if (!(g.hS[t * g.q + job] < INFINITY)) return;
        ref = (int)(g.col0 + t * g.q + job);
        dest = g.cposS[t * g.q + job];
    } else {
        const WEnt w = sn[job - g.p0];
        ref = w.ref;
        dest = pl.pA + w.pos;
    }
}

template <int IT>
__global__ void __launch_bounds__(NUM_WARPS * 32) apply_r_kernel(GP g, int k0, int jb) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int t = blockIdx.y, k = k0 + t, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PlanR pl = g.planR[k];
    const int* order = g.orderR + (int64_t)t * g.capListR;
    const WEnt* sn = g.snEnt + (int64_t)t * g.capW;
    WarpStage ws = warp_stage_init(g, smem_raw, jb);
    if (blockIdx.x < NBR) {
      for (int rbase = 0; rbase < g.ld; rbase += WIN_ROWS) {
        const Win win = win_at(g, rbase);
        Acc<IT> a;
        acc_zero(a);
        if (pl.mode == MODE_REDUCE) {
            // chunk of sigma per CTA, a contiguous sub-chunk per warp; jb lanes resolve jb references at once (order ->
            // snapshot -> column pointer) and the columns arrive together through the TMA engine
            const int chunk = (pl.m + NBR - 1) / NBR;
            const int lo = blockIdx.x * chunk, hi = min(pl.m, lo + chunk);
            const int wchunk = (chunk + NUM_WARPS - 1) / NUM_WARPS;
            const int wlo = lo + warp * wchunk, whi = min(hi, wlo + wchunk);
            for (int s0 = wlo; s0 < whi; s0 += jb) {
                int ref = NO_JOB;
                if (lane < jb && s0 + lane < whi) {
                    const int code = order[s0 + lane];
                    ref = code < g.p0 ? (int)(g.col0 + t * g.q + code) : sn[code - g.p0].ref;
                }
                const double* src = (ref != NO_JOB && ref >= 0) ? col_ptr(g, ref) : nullptr;
                const unsigned dense = stage_load(g, ws, src, lane, win);
                double tval = 0.0;
                int trow = -1;
                const bool is_tag = ref != NO_JOB && ref < 0 && g.rank == 0;   // box generators are replicated: rank 0 accounts for them
                if (is_tag) {
                    int kt;
                    tag_decode(ref, g.n, kt, trow);
                    tval = g.Wd[(int64_t)kt * g.ld + trow];
                }
                const unsigned tags = __ballot_sync(0xffffffffu, is_tag);
                if (dense) {
                    stage_wait(ws);
                    acc_add_staged(g, a, ws, dense, lane, win);
                }
                acc_add_tags(a, tags, tval, trow, lane, win);
            }
        }
        __syncthreads();
        acc_reduce_store(g, a, stage_red(g, smem_raw, jb), g.partR + ((int64_t)t * NBR + blockIdx.x) * g.ld, win);
      }
      return;
    }
    // the kept generators of R_k: their references, in the order of the reduced zonotope (sortperm order)
    int* rr = g.Rref + (int64_t)(k - 1) * g.capR;
    const int jobs = pl.mode == MODE_REDUCE ? pl.pin : g.p0 + pl.pW;
    const int nthreads = (gridDim.x - NBR) * blockDim.x;
    for (int job = (blockIdx.x - NBR) * blockDim.x + threadIdx.x; job < jobs; job += nthreads) {
        int ref, dest;
        resolve_copy_job(g, pl, order, sn, t, job, jobs, ref, dest);
        if (ref != NO_JOB) rr[dest] = ref;
    }
}

// Box vector, generator counts and (strip mode) the rows of the dense box generators of every R_k of the batch.
__global__ void __launch_bounds__(256) finalize_r_kernel(GP g, int k0) {
    __shared__ int scan[256];
    const int t = blockIdx.x, k = k0 + t, tid = threadIdx.x;
    const PlanR pl = g.planR[k];
    if (pl.mode != MODE_REDUCE) {
        if (tid == 0) g.meta[k - 1] = make_int2(pl.pin, 0);
        return;
    }
    const int per = (g.n + 255) / 256;
    int cnt = 0;
    for (int i = tid * per; i < min(g.n, (tid + 1) * per); ++i) {
        double d = 0.0;
        for (int b = 0; b < NBR; ++b) d += g.partR[((int64_t)t * NBR + b) * g.ld + i];
        g.Rd[(int64_t)(k - 1) * g.ld + i] = d;
        cnt += (!g.strip || d != 0.0);
    }
    scan[tid] = cnt;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const int v = tid >= o ? scan[tid - o] : 0;
        __syncthreads();
        scan[tid] += v;
        __syncthreads();
    }
    int p = scan[tid] - cnt;
    for (int i = tid * per; i < min(g.n, (tid + 1) * per); ++i) {
        const double d = g.Rd[(int64_t)(k - 1) * g.ld + i];
        if (!g.strip || d != 0.0) g.Rrow[(int64_t)(k - 1) * g.n + p++] = i;
    }
    if (tid == 0) g.meta[k - 1] = make_int2(pl.pin - pl.m, scan[255]);
}

// Dense generator matrices of the reach sets of steps [k0, k0 + cnt): slab t = pmax columns of leading dimension ldo (slab
// stride `stride` doubles); column j < nk is the generator Rref[j] refers to (an archive column, or a box generator of W
// whose value sits in the W box archive), the next nd columns are the box generators d[row] e_row of the reduction, the rest
// is zero.  One warp per column.  A column-sharded rank materialises only what it owns (the others' columns are zero; the
// replicated box generators belong to rank 0), so the slabs of the ranks sum to the reach set.
__global__ void __launch_bounds__(256) materialize_kernel(GP g, int k0, int cnt, int pmax, double* __restrict__ out, int ldo,
                                                          int64_t stride) {
    const int t = blockIdx.y, k = k0 + t, col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= cnt || col >= pmax) return;
    const int2 mt = g.meta[k - 1];
    const int nk = mt.x, nd = mt.y;
    double* dst = out + (int64_t)t * stride + (int64_t)col * ldo;
    const double* src = nullptr;
    int row = -1;
    double val = 0.0;
    if (col < nk) {
        const int ref = g.Rref[(int64_t)(k - 1) * g.capR + col];
        if (ref >= 0) {
            src = col_ptr(g, ref);
        } else {
            int kt;
            tag_decode(ref, g.n, kt, row);
            val = g.rank == 0 ? g.Wd[(int64_t)kt * g.ld + row] : 0.0;
        }
    } else if (col < nk + nd) {
        row = g.Rrow[(int64_t)(k - 1) * g.n + (col - nk)];
        val = g.rank == 0 ? g.Rd[(int64_t)(k - 1) * g.ld + row] : 0.0;
    }
    if (src) {
        for (int i = lane; i < g.n; i += 32) dst[i] = src[i];
    } else {
        for (int i = lane; i < g.n; i += 32) dst[i] = i == row ? val : 0.0;
    }
}

// Dense copy of the current W in logical order (box generators synthesised).  out: n x pW, leading dimension ldo.
__global__ void __launch_bounds__(256) materialize_w_kernel(GP g, double* __restrict__ out, int ldo) {
    const int pW = *g.wCount;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= pW) return;
    const int ref = g.wEnt[warp].ref;
    double* dst = out + (int64_t)g.wEnt[warp].pos * ldo;
    if (ref >= 0) {
        const double* src = col_ptr(g, ref);
        for (int i = lane; i < g.n; i += 32) dst[i] = src ? src[i] : 0.0;
    } else {
        int kt, it;
        tag_decode(ref, g.n, kt, it);
        const double val = g.Wd[(int64_t)kt * g.ld + it];
        for (int i = lane; i < g.n; i += 32) dst[i] = i == it ? val : 0.0;
    }
}

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

// Box hulls of EVERY member at EVERY step of an ensemble flowpipe in one pass (one warp per (member, step); lanes = rows, the
// member's p generator columns are contiguous): lo / hi [i + n (b + batch (k - 1))].  Reads the flowpipe once.
__global__ void __launch_bounds__(128) ensemble_boxes_all_kernel(int n, int ld, int p, int batch, int64_t N, int64_t q,
                                                                 const double* __restrict__ Y, double* __restrict__ lo,
                                                                 double* __restrict__ hi) {
    const int64_t job = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (job >= (int64_t)batch * N) return;
    const int64_t k = job / batch;
    const int b = (int)(job - k * batch);
    const double* slot = Y + k * q * ld;
    const double* G = slot + (int64_t)b * p * ld;
    const double* c = slot + ((int64_t)batch * p + b) * ld;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        if (i < n) {
            double s = 0.0;
#pragma unroll 8
            for (int j = 0; j < p; ++j) s += fabs(G[(int64_t)j * ld + i]);
            const double ci = c[i];
            lo[(k * batch + b) * n + i] = ci - s;
            hi[(k * batch + b) * n + i] = ci + s;
        }
    }
}

// ---- structurally sparse Phi: the generator map as a sparse product -------------------------------------------------------
// Decoupled / block-diagonal systems (the ISS-shaped C2: 135 independent 2 x 2 modes) have a Phi = e^{A delta} with a handful of
// non-zeros per row, and so have all its powers.  linear_map(Phi^w, Z) (GLGM06/reach.jl:16,48,54) is then an ELL sparse x
// generator-block product: HBM-bound (read X once, write Y once) instead of 2 n^2 flop per generator.  The products that
// are skipped are exact zeros, and the non-zero terms of a row are added in ascending k like the DMMA k-loop does, so the
// result equals the dense GEMM's.  Detection is automatic at flowpipe creation (REACH_GLGM06_DENSE=1 forces the GEMM).
constexpr int ELL_MAX = 8;

// One warp per row i of P (n x n, column-major, ld): non-zeros compacted in ascending k into idx/val[i*ELL_MAX + .]; count[i].
__global__ void __launch_bounds__(256) ell_build_kernel(int n, const double* __restrict__ P, int64_t ld, int* __restrict__ idx,
                                                        double* __restrict__ val, int* __restrict__ count) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    int base = 0;
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + lane;
        const double v = k < n ? P[(int64_t)k * ld + i] : 0.0;
        const unsigned mask = __ballot_sync(0xffffffffu, v != 0.0);
        const int pos = base + __popc(mask & ((1u << lane) - 1u));
        if (v != 0.0 && pos < ELL_MAX) {
            idx[i * ELL_MAX + pos] = k;
            val[i * ELL_MAX + pos] = v;
        }
        base += __popc(mask);
    }
    for (int w = base + lane; w < ELL_MAX; w += 32) {      // pad: multiplies x[0] by zero
        idx[i * ELL_MAX + w] = 0;
        val[i * ELL_MAX + w] = 0.0;
    }
    if (lane == 0) count[i] = base;
}

// Y[:, g] = P X[:, g] for Q generator columns (ld doubles each).  Thread = row i (its ELL row lives in registers), CTA =
// ELL_GPB generators: the gathered entries of a 2 KB column come from L1, the stores are coalesced.
constexpr int ELL_GPB = 16;
template <int W>
__global__ void __launch_bounds__(1024) ell_left_kernel(int n, int ld, int64_t Q, const int* __restrict__ idx,
                                                        const double* __restrict__ val, const double* __restrict__ X,
                                                        double* __restrict__ Y) {
    const int i = threadIdx.x;
    int ji[W];
    double vi[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        ji[w] = i < n ? idx[i * ELL_MAX + w] : 0;
        vi[w] = i < n ? val[i * ELL_MAX + w] : 0.0;
    }
    const int64_t g0 = (int64_t)blockIdx.x * ELL_GPB;
    constexpr int U = W <= 2 ? 8 : (W <= 4 ? 4 : 2);       // generators in flight per thread: U * W independent loads
#pragma unroll 1
    for (int u0 = 0; u0 < ELL_GPB; u0 += U) {
        double xv[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {                      // unconditional loads (clamped column) so that they all issue at once
            const int64_t g = min(g0 + u0 + u, Q - 1);
            const double* x = X + g * ld;
#pragma unroll
            for (int w = 0; w < W; ++w) xv[u][w] = x[ji[w]];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < W; ++w) acc = fma(vi[w], xv[u][w], acc);
            const int64_t g = g0 + u0 + u;
            if (g < Q && i < n) Y[g * ld + i] = acc;
        }
    }
}

// The same product FUSED with the Girard scores of the mapped generators (score_stack_kernel): the CTA keeps its 16 mapped
// columns in shared memory and its warps then run score_stack_kernel's column loop on them -- same lane striding, same
// warp tree, hence the same bits -- so the scores (and, in strip mode, the row-support masks of the input segment) never
// re-read the slots from HBM.  Single-rank flowpipes only (all columns of a slot are local: [G0 | GV | c0 | cV]).
template <int W>
__global__ void __launch_bounds__(1024) ell_left_score_kernel(GP g, int cnt, const int* __restrict__ idx,
                                                              const double* __restrict__ val, const double* __restrict__ X,
                                                              double* __restrict__ Yout) {
    // TMA-staged: the CTA's ELL_GPB input columns are contiguous in the archive, so ONE bulk copy (cp.async.bulk, SASS UBLKCP)
    // brings them into shared memory, the threads gather from there (thread = state row, its ELL row in registers), and ONE
    // bulk copy shared -> global writes the mapped columns while the warps compute their scores from the same tile.  No
    // per-thread global load or store is left on the path; three CTAs per SM keep ~200 KB in flight.
    extern __shared__ __align__(128) double esm[];          // [ELL_GPB * ld] input columns | [ELL_GPB * ld] mapped columns | mbarrier
    const int n = g.n, ld = g.ld;
    double* xin = esm;
    double* ycol = esm + ELL_GPB * ld;
    uint64_t* bar = reinterpret_cast<uint64_t*>(ycol + ELL_GPB * ld);
    const int i = threadIdx.x;
    const int64_t Q = (int64_t)cnt * g.q;
    const int64_t g0 = (int64_t)blockIdx.x * ELL_GPB;
    const int ncols = (int)min((int64_t)ELL_GPB, Q - g0);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, (uint32_t)(ncols * ld * 8));
        bulk_g2s(xin, X + g0 * ld, (uint32_t)(ncols * ld * 8), bar);
    }
    int ji[W];
    double vi[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        ji[w] = i < n ? idx[i * ELL_MAX + w] : 0;
        vi[w] = i < n ? val[i * ELL_MAX + w] : 0.0;
    }
    __syncthreads();                                         // the barrier is initialised before anyone polls it
    mbar_wait(bar, 0);
    if (i < ld) {
#pragma unroll 4
        for (int u = 0; u < ncols; ++u) {
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < W; ++w) acc = fma(vi[w], xin[u * ld + ji[w]], acc);
            ycol[u * ld + i] = i < n ? acc : 0.0;            // pad rows are zero, as in the archive
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        bulk_s2g(Yout + g0 * ld, ycol, (uint32_t)(ncols * ld * 8));
        bulk_commit();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int n2 = ld / 2;
    for (int u = warp; u < ncols; u += nwarps) {
        const int64_t gq = g0 + u;
        const int t = (int)(gq / g.q), c = (int)(gq - (int64_t)t * g.q);
        if (c >= g.p0 + g.pV) continue;                     // the two centre columns of a slot carry no score
        const double2* c2 = reinterpret_cast<const double2*>(ycol + u * ld);
        double s = 0.0, m = 0.0;
        uint32_t* mask = (g.strip && c >= g.p0) ? g.maskAll + ((g.col0 / g.q + t) * (int64_t)g.pV + (c - g.p0)) * g.nwords4 : nullptr;
        for (int it = 0; it * 32 < n2; ++it) {
            const int k = lane + 32 * it;
            double2 v = make_double2(0.0, 0.0);
            if (k < n2) v = c2[k];
            const double ax = fabs(v.x), ay = fabs(v.y);
            s += ax;
            s += ay;
            m = fmax(m, fmax(ax, ay));
            if (mask) {
                const unsigned bx = __ballot_sync(0xffffffffu, ax != 0.0), by = __ballot_sync(0xffffffffu, ay != 0.0);
                if (lane == 0) { mask[2 * it] = bx; mask[2 * it + 1] = by; }
            }
        }
        const double asum = warp_sum(s), amax = warp_max(m);
        if (lane == 0) g.hS[t * g.q + c] = (g.strip && asum == 0.0) ? INFINITY : asum - amax;
    }
    if (threadIdx.x == 0) bulk_wait0();                      // the tile has left shared memory before the CTA retires
}

// ---- projection of a zonotope flowpipe: M * R_k (project_reach.jl:18-99, linear map of every reach set) -----------
// One warp per generator column: lanes stride the column with 16-byte loads and carry the q <= 4 dot products with the
// rows of M (Mt: q rows of n doubles); out[((k-1)*pmax + col)*q + row].  Columns beyond the set's count are zeroed.
struct ProjIn {
    const double* G;        // generators of step 1
    int64_t slot;           // doubles between consecutive steps
    const double* c;        // centre of step 1
    int64_t cstride;
    const int2* meta;       // per-step generator counts (x + y), or nullptr: p_const
    int p_const;
    int n, ld, q, pmax, N;
};

__device__ __forceinline__ void proj_column(const double* __restrict__ col, const double* __restrict__ Mt, int n, int q,
                                            int lane, double acc[4]) {
    acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
    const int n2 = n >> 1;
#pragma unroll 4
    for (int i = lane; i < n2; i += 32) {
        const double2 v = reinterpret_cast<const double2*>(col)[i];
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r < q) {
                const double2 m = reinterpret_cast<const double2*>(Mt + (int64_t)r * (n + (n & 1)))[i];
                acc[r] = fma(m.y, v.y, fma(m.x, v.x, acc[r]));
            }
    }
    if ((n & 1) && lane == 0) {
        const double v = col[n - 1];
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r < q) acc[r] = fma(Mt[(int64_t)r * (n + 1) + n - 1], v, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (r < q)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
}

// `g` is only used when in.meta != nullptr (a reduced flowpipe: generators are references, see materialize_kernel)
__global__ void __launch_bounds__(256) project_gens_kernel(ProjIn in, GP g, const double* __restrict__ Mt, double* __restrict__ out) {
    const int k = blockIdx.y, col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= in.N || col >= in.pmax) return;
    const int p = in.meta ? in.meta[k].x + in.meta[k].y : in.p_const;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (col < p && in.meta == nullptr) {
        proj_column(in.G + (int64_t)k * in.slot + (int64_t)col * in.ld, Mt, in.n, in.q, lane, acc);
    } else if (col < p) {
        const int nk = in.meta[k].x;
        const double* src = nullptr;
        int row = -1;
        double val = 0.0;
        if (col < nk) {
            const int ref = g.Rref[(int64_t)k * g.capR + col];
            if (ref >= 0) {
                src = col_ptr(g, ref);
            } else {
                int kt;
                tag_decode(ref, g.n, kt, row);
                val = g.rank == 0 ? g.Wd[(int64_t)kt * g.ld + row] : 0.0;
            }
        } else {
            row = g.Rrow[(int64_t)k * g.n + (col - nk)];
            val = g.rank == 0 ? g.Rd[(int64_t)k * g.ld + row] : 0.0;
        }
        if (src) {
            proj_column(src, Mt, in.n, in.q, lane, acc);
        } else if (row >= 0) {                      // M (val e_row) = val M[:, row]
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (r < in.q) acc[r] = Mt[(int64_t)r * (in.n + (in.n & 1)) + row] * val;
        }
    }
    if (lane < in.q) {
        const double v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
        out[((int64_t)k * in.pmax + col) * in.q + lane] = v;
    }
}

// Centre M c_k and the box radius sum_j |M g_j| of every projected reach set (set_type_proj = Hyperrectangle,
// default_options.jl:69), fixed summation order: thread t adds columns t, t+256, ... and a shared-memory tree follows.
__global__ void __launch_bounds__(256) project_box_kernel(ProjIn in, const double* __restrict__ Mt, const double* __restrict__ Gp,
                                                          double* __restrict__ centers, double* __restrict__ radii) {
    __shared__ double red[256];
    const int k = blockIdx.x;
    const int p = in.meta ? in.meta[k].x + in.meta[k].y : in.p_const;
    for (int r = 0; r < in.q; ++r) {
        double s = 0.0;
        for (int j = threadIdx.x; j < p; j += 256) s += fabs(Gp[((int64_t)k * in.pmax + j) * in.q + r]);
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) radii[(int64_t)k * in.q + r] = red[0];
        __syncthreads();
    }
    if (threadIdx.x < 32) {
        double acc[4];
        proj_column(in.c + (int64_t)k * in.cstride, Mt, in.n, in.q, threadIdx.x, acc);
        if ((int)threadIdx.x < in.q) {
            const int l = threadIdx.x;
            centers[(int64_t)k * in.q + l] = l == 0 ? acc[0] : l == 1 ? acc[1] : l == 2 ? acc[2] : acc[3];
        }
    }
}

// Homogeneous recurrence for small state dimension (n <= 32; config C5: n = 28): Y_k = Phi Y_{k-1}, exactly the reference's
// R_k = linear_map(Phi, R_{k-1}) (GLGM06/reach.jl:15-19).  The whole flowpipe is ONE launch: every slot is written once
// and read never (HBM traffic = the flowpipe itself).  A thread-per-column FMA version reached 150 k steps/s on 512
// members of C5 (the FP64 vector pipe tops out near 11 TFLOP/s on B200); this tensor-core version reaches 270 k:
// a warp owns 32 consecutive columns.  Phi sits in registers as m8n8k4 A fragments for the whole flowpipe; every
// step reads the B fragments of the current columns from a padded shared-memory tile (pitch 36: conflict-free), runs
// MT x 4 x KT DMMAs, writes the accumulators back into the tile and streams the tile out with 256-byte stores.
__device__ __forceinline__ void dmma884_acc(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// pitch (doubles) of a column in the shared-memory tiles: the smallest P >= ld with P mod 16 in {4, 12}, which makes both the
// B-fragment loads (slot 12 g + t) and the accumulator stores (slot 8 t + g) conflict-free; P == ld (C5: 28) lets the
// whole tile leave with ONE bulk copy
static int homog_pitch(int ld, int kt) {
    int P = std::max(ld, 4 * kt);       // a B fragment reads k = 0 .. 4 KT - 1 of its column: they must stay inside the column
    while ((P & 15) != 4 && (P & 15) != 12) P += 2;
    return P;
}

// MT x KT: m8 / k4 tiles covering the n x n matrix; NT: n8 column tiles per warp (a warp owns 8 NT consecutive columns).
// Per step and warp: KT x NT fragment loads from the current tile, MT x NT x KT DMMAs, the accumulators go to the OTHER tile
// of the warp's pair, and that tile leaves as a bulk copy shared -> global (TMA engine, SASS UBLKCP.G.S) issued by one lane --
// no per-thread store instructions, no index arithmetic, and the next step's DMMAs start while the copy drains.
__device__ __forceinline__ void dmma884_first(double& c0, double& c1, double a, double b) {       // D = A B (C = 0)
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(c0), "=d"(c1) : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}

// The kernel is bound by the DMMA pipe of each SM, not by HBM (AI = 7 flop/B against a machine balance of 5.7), so the SMs
// must carry EQUAL work and every scheduler needs several warps to keep the pipe fed across the per-step epilogue: one CTA per
// SM with W warps of 8 NT columns each, (W, NT) chosen on the host so that one wave of CTAs covers all columns with the least
// slack (C5 on one GPU: 145 CTAs x 19 warps x 16 columns).  Phi lives in shared memory as m8n8k4 A fragments (pitch 36:
// conflict-free), which keeps the kernel under 107 registers -- 19 warps per SM.  The warps never synchronise after the prologue.
constexpr int HOMOG_PA = 36;
template <int MT, int KT, int NT>
__global__ void __launch_bounds__(640, 1) homog_small_dmma_kernel(const double* __restrict__ Phi, int ldphi, int n, int ld, int64_t q,
                                                                  int64_t nsteps, double* __restrict__ Y, int P) {
    extern __shared__ __align__(128) double hsm[];
    constexpr int NC = 8 * NT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    double* const As = hsm;                                         // As[k * HOMOG_PA + row], 4 KT x 8 MT, zero-padded
    for (int idx = threadIdx.x; idx < 4 * KT * HOMOG_PA; idx += blockDim.x) {
        const int k = idx / HOMOG_PA, row = idx - k * HOMOG_PA;
        As[idx] = (row < n && k < n) ? Phi[row + (int64_t)k * ldphi] : 0.0;
    }
    __syncthreads();
    const int64_t col0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * NC;
    if (col0 >= q) return;                                    // whole warps leave; nothing below synchronises the CTA
    const int ncol = (int)min((int64_t)NC, q - col0);
    double* const tiles = hsm + 4 * KT * HOMOG_PA + (size_t)(warp * 2) * NC * P;    // the warp's pair of tiles, NC * P doubles each
    const int tile_doubles = NC * P;
    const bool compact = (P == ld);
    for (int idx = lane; idx < 2 * tile_doubles; idx += 32) tiles[idx] = 0.0;
    __syncwarp();
    {
        const double* src = Y + col0 * ld;
        for (int c = 0; c < ncol; ++c)
            for (int i = lane; i < n; i += 32) tiles[c * P + i] = src[(int64_t)c * ld + i];
    }
    __syncwarp();
    const double* const afrag = As + t * HOMOG_PA + g;
    int cur = 0;
    for (int64_t step = 1; step <= nsteps; ++step) {
        const double* tin = tiles + cur * tile_doubles + g * P + t;
        double* tout = tiles + (cur ^ 1) * tile_doubles;
        double d[MT][NT][2];
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            double av[MT], bv[NT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) av[mt] = afrag[4 * kt * HOMOG_PA + 8 * mt];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) bv[nt] = tin[8 * nt * P + 4 * kt];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    if (kt == 0) dmma884_first(d[mt][nt][0], d[mt][nt][1], av[mt], bv[nt]);
                    else dmma884_acc(d[mt][nt][0], d[mt][nt][1], av[mt], bv[nt]);
                }
        }
        // the bulk copy that left `tout` two steps ago has finished reading it (at most the previous step's copy is pending)
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int row = 8 * mt + g;
            if (row < n) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    tout[(8 * nt + 2 * t) * P + row] = d[mt][nt][0];
                    tout[(8 * nt + 2 * t + 1) * P + row] = d[mt][nt][1];
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes before the async-proxy read
        __syncwarp();
        double* dst = Y + (step * q + col0) * ld;
        if (compact) {
            if (lane == 0) {
                bulk_s2g(dst, tout, (uint32_t)(ncol * ld * 8));
                bulk_commit();
            }
        } else if (lane < ncol) {
            bulk_s2g(dst + (int64_t)lane * ld, tout + lane * P, (uint32_t)(ld * 8));
            bulk_commit();
        }
        cur ^= 1;
    }
    bulk_wait0();
}

template <typename T>
T* dalloc(Ctx* ctx, size_t count, bool zero = true) {
    if (count == 0) count = 1;
    T* p = (T*)ctx_malloc(ctx, sizeof(T) * count);
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
    int64_t qloc = 0;               // archive columns per slot on this rank (= q unless column-sharded)
    int rank = 0, nranks = 1;
    // column-sharded phase machine (reach_glgm06_shard_*): batches are processed in order, collectives in between
    struct Batch { int64_t have, cnt, w; bool square_after; };
    std::vector<Batch> sched;
    int64_t next_batch = 0;         // next batch whose phases 1..3 are due
    int64_t next_p0 = 0;            // next batch whose phase 0 (map + scores) is due: may run ahead of next_batch on another stream
    std::vector<cudaEvent_t> shard_done;    // phase 3 of batch b issued (orders the reuse of the batch-local array sets)
    const DMat* pw = nullptr;
    int sp = 0, shard_lvl = 0;
    GP g{};
    DMat Phi, PowA, PowB;
    // ELL form of Phi^(2^l) for the first `ell_levels` doubling levels (structurally sparse Phi); width = non-zeros per row
    int ell_levels = 0, ell_width = 0;
    int* ell_idx = nullptr; double* ell_val = nullptr;       // [level][n][ELL_MAX]
    bool force_dense = false;
    double* Hflow = nullptr;        // archive of Phi^j X: the flowpipe itself in the homogeneous case
    double* scr = nullptr;          // chain scratch when the W list does not fit in shared memory
    size_t chain_smem = 0;
    int chain_use_smem = 1;
    bool chain_exclusive = false;   // the chain owns one SM (the GEMM grid leaves it free)
    cudaStream_t side = nullptr;    // selection chain of batch b (overlaps the GEMM of batch b+1) ...
    cudaStream_t side2 = nullptr;   // ... and the numerics of batch b (overlap the chain of batch b+1)
    cudaStream_t side3 = nullptr;   // scores + segment order of batch b (only need the GEMM of batch b)
    cudaStream_t side4 = nullptr;   // R sortperm of batch b (only needs the chain of batch b)
    // batch-local arrays, NSET-fold buffered: with two sets the loop numerics(b-1) -> map(b+1) -> sort -> chain(b+1) ->
    // numerics(b+1) spans two batches and left a gap between consecutive chains (0.31 ms per batch at C2); with three sets it
    // spans three and the chain (0.25 ms per batch) runs back to back
    static constexpr int NSET_MAX = 8;
    int nset = 3;                   // sets in use (REACH_GLGM06_SETS, diagnostics)
    GP gset[NSET_MAX];
    std::vector<void*> owned;
    std::vector<int2> meta_host;
    bool ran = false;
    double run_ms = 0, gemm_ms = 0, gemm_flops = 0, reduce_ms = 0, reduce_bytes = 0, chain_ms = 0;
    int64_t gemm_launches = 0;
    std::vector<cudaEvent_t> evs;
    double* stage = nullptr;        // staging for box queries
    // streaming fetch (reach_glgm06_run_fetch): host destination, copy stream, device staging of one batch
    double* sink_c = nullptr; double* sink_G = nullptr; int64_t sink_pmax = 0;
    cudaStream_t cpy = nullptr;
    double* pack = nullptr;
    std::vector<double> phi_host;   // n <= 32: Phi as a kernel parameter of homog_small_kernel
    bool serial = false;            // diagnostics: no stream overlap, every kernel timed alone (reach_flowpipe_set_serial)
    bool detail = false;            // REACH_TIMING_DETAIL
    std::vector<cudaEvent_t> detail_ev;
};

void flowpipe_free(Flowpipe* F) {
    if (!F) return;
    // the blocks go back to the context's pool at once: no stream of this flowpipe may still be using them (a run that
    // threw half-way leaves its side streams busy)
    for (cudaStream_t s : {F->side, F->side2, F->side3, F->side4, F->cpy})
        if (s) cudaStreamSynchronize(s);
    if (F->ctx) cudaStreamSynchronize(F->ctx->stream);
    dmat_free(F->Phi); dmat_free(F->PowA); dmat_free(F->PowB);
    for (void* p : F->owned) ctx_free(F->ctx, p);
    for (cudaEvent_t e : F->evs) cudaEventDestroy(e);
    for (cudaEvent_t e : F->shard_done) cudaEventDestroy(e);
    if (F->side) cudaStreamDestroy(F->side);
    if (F->side2) cudaStreamDestroy(F->side2);
    if (F->side3) cudaStreamDestroy(F->side3);
    if (F->side4) cudaStreamDestroy(F->side4);
    if (F->cpy) cudaStreamDestroy(F->cpy);
    delete F;
}

static int64_t pick_block(int64_t n, int64_t q, int64_t N) {
    // enough stacked columns for ~8 waves of 128-row tiles on 148 SMs (power of two: the doubling schedule needs it);
    // measured at C2: s = 16 / 32 / 64 / 128 -> 27.7 / 22.8 / 21.5 / 21.3 ms per 2000-step flowpipe
    const int64_t tiles_n = ceil_div(n, 96);
    int64_t s = 1;
    while (s < 64 && s < N - 1 && ceil_div(s * q, 128) * tiles_n < 148 * 8) s *= 2;
    return s;
}

Flowpipe* flowpipe_create(Ctx* ctx, int64_t n, const double* Phi, int64_t ldphi, int64_t p0, const double* c0,
                          const double* G0, int64_t ldg0, int64_t pV, const double* cV, const double* GV, int64_t ldgv,
                          int64_t N, int64_t max_order, int32_t flags, int64_t nc = 1, int rank = 0, int nranks = 1) {
    REACH_REQUIRE(n > 0 && Phi && ldphi >= n, "glgm06: bad Phi");
    REACH_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "glgm06: bad rank / nranks");
    REACH_REQUIRE(p0 >= 0 && c0 && (p0 == 0 || (G0 && ldg0 >= n)), "glgm06: bad Omega0");
    REACH_REQUIRE(N >= 1, "glgm06: N must be >= 1");
    const bool homog = (cV == nullptr);
    if (homog) pV = 0;
    REACH_REQUIRE(pV >= 0 && (pV == 0 || (GV && ldgv >= n)), "glgm06: bad input zonotope");
    if (!homog && max_order < 1)
        throw Error(REACH_EUNSUPPORTED, "glgm06: max_order < 1 (reduce_order never reduces; unbounded growth) is not supported");
    if (!homog && n > 64 * CHAIN_MAX_IT)
        throw Error(REACH_EUNSUPPORTED, "glgm06: inhomogeneous path supports n <= 4096 in this build");
    Flowpipe* F = new Flowpipe();
    try {
        F->ctx = ctx; F->n = n; F->p0 = p0; F->pV = pV; F->N = N; F->r = max_order; F->flags = flags;
        F->homog = homog;
        F->record = (flags & REACH_GLGM06_RECORD_SELECTIONS) != 0;
        F->ld = round_up(n, 4);   // 32-byte aligned columns
        const int64_t ld = F->ld;
        F->nc = homog ? nc : 1;
        F->q = homog ? p0 + F->nc : p0 + pV + 2;
        const int64_t q = F->q;
        F->s = N > 2 ? pick_block(n, q, N) : 1;
        if (!homog && N > 2) {
            // inhomogeneous path: every batch costs ~12 launches (and, column-sharded, 4 all-reduces), so long batches pay even
            // when the GEMM grid is full much earlier (n = 1024, N = 200: s = 4 / 16 / 64 -> 71.2 / 66.3 / 66.0 ms on one GPU,
            // 52.8 / 40.4 / 37.6 ms column-sharded over two); bounded by ~1 GB of batch-local arrays (3 sets)
            const int64_t cap = std::max<int64_t>(1, (int64_t)1e9 / (192 * std::max<int64_t>(1, max_order * n)));
            while (F->s < 64 && F->s < N - 1 && F->s * 2 <= cap) F->s *= 2;
        }
        if (getenv("REACH_GLGM06_BLOCK") && N > 2) {      // tuning knob, rounded down to a power of two
            int64_t want = std::max(1, std::min(128, atoi(getenv("REACH_GLGM06_BLOCK")))), pw2 = 1;
            while (pw2 * 2 <= want) pw2 *= 2;
            F->s = pw2;
        }
        const int64_t s = F->s;
        cudaStream_t st = ctx->stream;
        F->Phi = dmat_alloc(ctx, n, n);
        dmat_upload(ctx, F->Phi, Phi, ldphi);
        if (homog && n <= 32) {
            F->phi_host.resize((size_t)n * n);
            for (int64_t k = 0; k < n; ++k) std::memcpy(&F->phi_host[k * n], Phi + k * ldphi, sizeof(double) * n);
        }
        if (s > 1) { F->PowA = dmat_alloc(ctx, n, n); F->PowB = dmat_alloc(ctx, n, n); }
        auto own = [&](void* p) { F->owned.push_back(p); return p; };
        auto up2d = [&](double* dst, const double* src, int64_t lds, int64_t cols) {
            if (cols > 0)
                REACH_CUDA(cudaMemcpy2DAsync(dst, ld * 8, src, lds * 8, n * 8, cols, cudaMemcpyHostToDevice, st));
        };
        GP& g = F->g;
        g.n = (int)n; g.ld = (int)ld; g.p0 = (int)p0; g.pV = (int)pV; g.q = (int)q;
        F->rank = rank; F->nranks = nranks;
        g.rank = rank; g.nranks = nranks;
        if (homog) {
            REACH_REQUIRE(nranks == 1, "glgm06: homogeneous flowpipes are sharded by members, not by columns");
            g.a0 = 0; g.a1 = (int)p0; g.c0r = 0; g.c1r = 0; g.la = (int)p0; g.lc = 0; g.qloc = (int)q;
        } else {
            g.a0 = (int)(p0 * rank / nranks); g.a1 = (int)(p0 * (rank + 1) / nranks);
            g.c0r = (int)(pV * rank / nranks); g.c1r = (int)(pV * (rank + 1) / nranks);
            g.la = g.a1 - g.a0; g.lc = g.c1r - g.c0r; g.qloc = g.la + g.lc + 2;
        }
        F->qloc = g.qloc;
        const int64_t qloc = F->qloc;
        g.strip = (flags & REACH_GLGM06_KEEP_ZERO_GENERATORS) ? 0 : 1;
        g.record = F->record ? 1 : 0;
        g.nwords = 2 * (int)ceil_div(ld / 2, 32);
        g.nwords4 = (int)round_up(g.nwords, 4);
        // ---- archive: slot j = Phi^j X  (homogeneous: X = [G0 | c0], the flowpipe itself; ensemble: [G_0 .. | c_0 ..])
        const size_t cols = (size_t)N * qloc + 256;
        size_t free_b = 0, total_b = 0;
        REACH_CUDA(cudaMemGetInfo(&free_b, &total_b));
        free_b += ctx->pooled_bytes;       // cached blocks are released on demand (ctx_malloc)
        double need = (double)cols * ld * 8.0;
        if (!homog) {
            const int64_t rn = max_order * n;
            g.rn = rn; g.nkeep = n * (max_order - 1);
            g.capW = (int)std::max<int64_t>(rn, pV) + 8;
            g.capR = (int)std::max<int64_t>(rn, p0) + 8;
            g.capListR = (int)p0 + g.capW;
            g.capListW = g.capW + (int)pV;
            g.capDisc = (int)std::max<int64_t>(g.capW + pV - g.nkeep, 0) + 8;
            g.smax = (int)s;
            need += (double)N * (double)g.capR * 4.0;          // the reach sets are generator references, not copies
            REACH_REQUIRE((double)N * (double)std::max<int64_t>(q, n) < 2.0e9, "glgm06: N * columns exceeds the 32-bit reference range");
        }
        if (need > 0.92 * (double)free_b) {
            char buf[320];
            snprintf(buf, sizeof(buf), "glgm06: the flowpipe needs %.1f GB of HBM (N=%lld sets, %lld mapped columns per step%s), "
                     "%.1f GB free: shard the ensemble across GPUs or shorten the horizon",
                     need / 1e9, (long long)N, (long long)q, homog ? "" : " + the reduced zonotopes", (double)free_b / 1e9);
            throw Error(REACH_EUNSUPPORTED, buf);
        }
        F->Hflow = (double*)own(dalloc<double>(ctx, cols * ld));
        g.Yall = F->Hflow;
        if (g.la > 0) up2d(F->Hflow, G0 + (int64_t)g.a0 * ldg0, ldg0, g.la);
        if (homog) {
            up2d(F->Hflow + p0 * ld, c0, n, F->nc);
        } else {
            if (g.lc > 0) up2d(F->Hflow + (int64_t)g.la * ld, GV + (int64_t)g.c0r * ldgv, ldgv, g.lc);
            up2d(F->Hflow + (int64_t)(g.la + g.lc) * ld, c0, n, 1);
            up2d(F->Hflow + (int64_t)(g.la + g.lc + 1) * ld, cV, n, 1);
            const size_t sq = (size_t)s * q + 8;
            g.maskAll = (uint32_t*)own(dalloc<uint32_t>(ctx, g.strip ? (size_t)(N + 1) * std::max<int64_t>(pV, 1) * g.nwords4 + 4 : 4));
            g.Wd = (double*)own(dalloc<double>(ctx, (size_t)(N + 2) * ld));
            g.wEnt = (WEnt*)own(dalloc<WEnt>(ctx, g.capW));
            g.wCount = (int*)own(dalloc<int>(ctx, 1));
            g.wEntAlt = (WEnt*)own(dalloc<WEnt>(ctx, g.capW));
            g.fastCount = (int*)own(dalloc<int>(ctx, 2));
            g.planW = (PlanW*)own(dalloc<PlanW>(ctx, N + 2));
            g.planR = (PlanR*)own(dalloc<PlanR>(ctx, N + 2));
            g.cW = (double*)own(dalloc<double>(ctx, ld));
            if (F->record) {
                g.recOrdW = (int*)own(dalloc<int>(ctx, (size_t)(N + 2) * g.capListW));
                g.recHW = (double*)own(dalloc<double>(ctx, (size_t)(N + 2) * g.capListW));
                g.recOrdR = (int*)own(dalloc<int>(ctx, (size_t)(N + 2) * g.capListR));
                g.recHR = (double*)own(dalloc<double>(ctx, (size_t)(N + 2) * g.capListR));
            }
            g.Rref = (int*)own(dalloc<int>(ctx, (size_t)N * g.capR, false));
            g.Rc = (double*)own(dalloc<double>(ctx, (size_t)N * ld));
            g.Rd = (double*)own(dalloc<double>(ctx, (size_t)N * ld));
            g.Rrow = (int*)own(dalloc<int>(ctx, (size_t)N * n));
            g.meta = (int2*)own(dalloc<int2>(ctx, N + 1));
            // chain working set: keys 2 cap doubles, refs/pos 4 cap ints, sigma (cap + pV) ints
            // chain working set: two copies of the sorted list (16 B per generator), histogram cap + 2 ints, 52 bytes per input generator
            F->chain_smem = (size_t)g.capW * 32 + (size_t)(pV + 2) * 32 + (size_t)g.capDisc * 4 + 64;
            F->chain_use_smem = F->chain_smem <= 200 * 1024 ? 1 : 0;
            if (!F->chain_use_smem) F->scr = (double*)own(dalloc<double>(ctx, F->chain_smem / 8 + 8));
            if (const char* e = getenv("REACH_GLGM06_SETS")) F->nset = std::max(2, std::min(Flowpipe::NSET_MAX, atoi(e)));
            for (int b = 0; b < F->nset; ++b) {      // batch-local arrays, one set per batch in flight
                GP& h = F->gset[b];
                h.hS = (double*)own(dalloc<double>(ctx, sq));
                h.skS = (double*)own(dalloc<double>(ctx, sq));
                h.rkS = (int*)own(dalloc<int>(ctx, sq));
                h.cposS = (int*)own(dalloc<int>(ctx, sq));
                h.cntS = (int*)own(dalloc<int>(ctx, 2 * s + 8));
                h.ordS = (int*)own(dalloc<int>(ctx, sq));
                h.cpsS = (int*)own(dalloc<int>(ctx, sq));
                h.cntZ = (int*)own(dalloc<int>(ctx, 2 * s + 8));
                h.snEnt = (WEnt*)own(dalloc<WEnt>(ctx, (size_t)s * g.capW));
                h.snCount = (int*)own(dalloc<int>(ctx, s));
                h.fastFlag = (int*)own(dalloc<int>(ctx, 1));
                h.dlistW = (int*)own(dalloc<int>(ctx, (size_t)s * g.capDisc));
                h.tcnt = (int*)own(dalloc<int>(ctx, (size_t)s * n));
                h.tsrc = (int*)own(dalloc<int>(ctx, (size_t)s * n * 4));
                h.orderR = (int*)own(dalloc<int>(ctx, (size_t)s * g.capListR));
                h.partW = (double*)own(dalloc<double>(ctx, (size_t)s * NBW * ld));
                h.partR = (double*)own(dalloc<double>(ctx, (size_t)s * NBR * ld));
            }
            int prio_lo = 0, prio_hi = 0;
            REACH_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            REACH_CUDA(cudaStreamCreateWithPriority(&F->side, cudaStreamNonBlocking, prio_hi));      // the chain is the critical path
            // the map / GEMM of later batches (main stream, lowest priority) is far ahead of the pipeline; everything downstream
            // of it gets its CTAs scheduled first as SM slots free up
            REACH_CUDA(cudaStreamCreateWithPriority(&F->side2, cudaStreamNonBlocking, prio_hi));
            REACH_CUDA(cudaStreamCreateWithPriority(&F->side3, cudaStreamNonBlocking, prio_hi));
            REACH_CUDA(cudaStreamCreateWithPriority(&F->side4, cudaStreamNonBlocking, prio_hi));
            // R_1 = Omega0 (GLGM06/reach.jl:39-40)
            {       // R_1 = Omega0: the columns 0 .. p0-1 of archive slot 0
                std::vector<int> iota((size_t)std::max<int64_t>(p0, 1));
                for (int64_t j = 0; j < p0; ++j) iota[(size_t)j] = (int)j;
                REACH_CUDA(cudaMemcpyAsync(g.Rref, iota.data(), sizeof(int) * (size_t)p0, cudaMemcpyHostToDevice, st));
                REACH_CUDA(cudaStreamSynchronize(st));
            }
            REACH_CUDA(cudaMemcpyAsync(g.Rc, c0, n * 8, cudaMemcpyHostToDevice, st));
            const int2 m1 = make_int2((int)p0, 0);
            REACH_CUDA(cudaMemcpyAsync(g.meta, &m1, sizeof(int2), cudaMemcpyHostToDevice, st));
        }
        // structurally sparse Phi?  ELL forms of Phi^(2^l) for the doubling levels the run uses (see ell_left_kernel)
        F->force_dense = getenv("REACH_GLGM06_DENSE") != nullptr;
        if (F->phi_host.empty() && n >= 16 && n <= 1024 && !F->force_dense) {      // one thread per state row in the ELL kernels
            int nlev = 1;
            for (int64_t w = 1; w < s; w *= 2) ++nlev;
            int* idx = (int*)own(dalloc<int>(ctx, (size_t)nlev * n * ELL_MAX));
            double* val = (double*)own(dalloc<double>(ctx, (size_t)nlev * n * ELL_MAX));
            int* cnt = (int*)own(dalloc<int>(ctx, n));
            std::vector<int> hc(n);
            const DMat* P = &F->Phi;
            DMat* spare[2] = {&F->PowA, &F->PowB};
            int sp = 0, width = 0;
            for (int lvl = 0; lvl < nlev; ++lvl) {
                ell_build_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>((int)n, P->p, P->ld, idx + (size_t)lvl * n * ELL_MAX,
                                                                                 val + (size_t)lvl * n * ELL_MAX, cnt);
                REACH_CUDA(cudaGetLastError());
                ctx->launches++;
                REACH_CUDA(cudaMemcpyAsync(hc.data(), cnt, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
                REACH_CUDA(cudaStreamSynchronize(st));
                const int mx = *std::max_element(hc.begin(), hc.end());
                if (mx > ELL_MAX || (int64_t)mx * 8 > n) break;
                width = std::max(width, mx);
                F->ell_levels = lvl + 1;
                if (lvl + 1 < nlev) {
                    gemm_left(ctx, n, n, n, P->p, P->ld, P->p, P->ld, spare[sp]->p, spare[sp]->ld);
                    P = spare[sp];
                    sp ^= 1;
                }
            }
            if (F->ell_levels < nlev) F->ell_levels = 0;      // mixed sparse / dense levels: not worth the bookkeeping
            F->ell_width = width;
            F->ell_idx = idx;
            F->ell_val = val;
        }
        REACH_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        flowpipe_free(F);
        throw;
    }
    return F;
}

// ELL product of a whole batch fused with its scores (g bound to the batch's array set, g.Y / g.col0 set): true when launched
static bool map_slots_scored(Flowpipe* F, int lvl, GP& g, int64_t slot0, int cnt, const double* Xp, double* Yp) {
    Ctx* ctx = F->ctx;
    const int64_t n = F->n, ld = F->ld;
    const size_t smem = (size_t)2 * ELL_GPB * ld * sizeof(double) + 16;
    if (lvl >= F->ell_levels || F->nranks != 1 || smem > 200 * 1024) return false;
    g.Y = F->Hflow + slot0 * F->qloc * ld;
    g.col0 = slot0 * F->q;
    static bool attr[64] = {false};
    static std::mutex attr_mutex;
    std::lock_guard<std::mutex> attr_lock(attr_mutex);
    if (!attr[ctx->device & 63]) {
        REACH_CUDA(cudaFuncSetAttribute(ell_left_score_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        REACH_CUDA(cudaFuncSetAttribute(ell_left_score_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        REACH_CUDA(cudaFuncSetAttribute(ell_left_score_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr[ctx->device & 63] = true;
    }
    const int* idx = F->ell_idx + (size_t)lvl * n * ELL_MAX;
    const double* val = F->ell_val + (size_t)lvl * n * ELL_MAX;
    const unsigned grid = (unsigned)ceil_div((int64_t)cnt * F->q, ELL_GPB), block = (unsigned)round_up(ld, 32);
    if (F->ell_width <= 2) ell_left_score_kernel<2><<<grid, block, smem, ctx->stream>>>(g, cnt, idx, val, Xp, Yp);
    else if (F->ell_width <= 4) ell_left_score_kernel<4><<<grid, block, smem, ctx->stream>>>(g, cnt, idx, val, Xp, Yp);
    else ell_left_score_kernel<8><<<grid, block, smem, ctx->stream>>>(g, cnt, idx, val, Xp, Yp);
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
    return true;
}

// slots Y = Phi^(2^lvl) X for Q generator columns: the ELL product when Phi is structurally sparse, else the DMMA GEMM
static bool map_slots(Flowpipe* F, int lvl, const DMat& P, int64_t Q, const double* Xp, double* Yp) {
    Ctx* ctx = F->ctx;
    const int64_t n = F->n, ld = F->ld;
    if (lvl < F->ell_levels) {
        const int* idx = F->ell_idx + (size_t)lvl * n * ELL_MAX;
        const double* val = F->ell_val + (size_t)lvl * n * ELL_MAX;
        const unsigned grid = (unsigned)ceil_div(Q, ELL_GPB), block = (unsigned)round_up(n, 32);
        if (F->ell_width <= 2) ell_left_kernel<2><<<grid, block, 0, ctx->stream>>>((int)n, (int)ld, Q, idx, val, Xp, Yp);
        else if (F->ell_width <= 4) ell_left_kernel<4><<<grid, block, 0, ctx->stream>>>((int)n, (int)ld, Q, idx, val, Xp, Yp);
        else ell_left_kernel<8><<<grid, block, 0, ctx->stream>>>((int)n, (int)ld, Q, idx, val, Xp, Yp);
        REACH_CUDA(cudaGetLastError());
        ctx->launches++;
        return true;
    }
    gemm_left(ctx, n, Q, n, P.p, P.ld, Xp, ld, Yp, ld);
    return false;
}

static cudaEvent_t next_event(Flowpipe* F, size_t& used) {
    if (used == F->evs.size()) {
        cudaEvent_t e;
        REACH_CUDA(cudaEventCreate(&e));
        F->evs.push_back(e);
    }
    return F->evs[used++];
}

// scores (+ row-support masks) of the mapped generators this rank owns, for `cnt` steps whose first slot is `slot0`
static void score_only(Flowpipe* F, GP& g, int64_t slot0, int cnt, cudaStream_t st) {
    Ctx* ctx = F->ctx;
    g.Y = F->Hflow + slot0 * F->qloc * F->ld;
    g.col0 = slot0 * F->q;
    if (F->nranks > 1) {        // entries of the other ranks must be 0 for the sum all-reduce
        REACH_CUDA(cudaMemsetAsync(g.hS, 0, sizeof(double) * (size_t)cnt * F->q, st));
        if (g.strip && F->pV > 0)
            REACH_CUDA(cudaMemsetAsync(g.maskAll + (size_t)slot0 * F->pV * g.nwords4, 0, sizeof(uint32_t) * (size_t)cnt * F->pV * g.nwords4, st));
    }
    const int per = g.la + g.lc;
    if (per > 0) {
        score_stack_kernel<<<(unsigned)ceil_div((int64_t)cnt * per, 8), 256, 0, st>>>(g, cnt);
        REACH_CUDA(cudaGetLastError());
        ctx->launches++;
    }
}

// stable order of the (global) segments from the (all-reduced) scores
static void sort_only(Flowpipe* F, GP& g, int cnt, cudaStream_t st) {
    Ctx* ctx = F->ctx;
    const int maxlen = std::max(g.p0, g.pV);
    if (maxlen > 0) {
        if (maxlen <= BITONIC_MAX) {
            int P = 2;
            while (P < maxlen) P <<= 1;
            sort_segments_bitonic_kernel<<<(unsigned)(cnt * 2), 512, (size_t)P * 12, st>>>(g, P);
        } else {
            dim3 grid((unsigned)std::max<int64_t>(1, ceil_div(maxlen, 256)), (unsigned)(cnt * 2));
            sort_segments_kernel<<<grid, 256, 0, st>>>(g);
        }
        REACH_CUDA(cudaGetLastError());
        ctx->launches++;
    } else {
        REACH_CUDA(cudaMemsetAsync(g.cntS, 0, sizeof(int) * 2 * cnt, st));
    }
}

static void score_and_sort(Flowpipe* F, GP& g, int64_t slot0, int cnt, cudaStream_t st) {
    score_only(F, g, slot0, cnt, st);
    sort_only(F, g, cnt, st);
}

static void bind_set(GP& g, const GP& h) {
    g.hS = h.hS; g.skS = h.skS; g.rkS = h.rkS; g.cposS = h.cposS; g.cntS = h.cntS;
    g.ordS = h.ordS; g.cpsS = h.cpsS; g.cntZ = h.cntZ;
    g.snEnt = h.snEnt; g.snCount = h.snCount; g.fastFlag = h.fastFlag; g.dlistW = h.dlistW; g.tcnt = h.tcnt; g.tsrc = h.tsrc;
    g.orderR = h.orderR; g.partW = h.partW; g.partR = h.partR;
}

// columns per warp stage and dynamic shared memory of the TMA-staged numerics kernels: stages | mbarriers | reduction scratch
static int numerics_jb(int64_t ld) {
    const int64_t pitch = std::min<int64_t>(ld, WIN_ROWS);
    return (int)std::max<int64_t>(1, std::min<int64_t>(COPY_JB, STAGE_BYTES / (pitch * 8)));
}
static size_t numerics_smem(int64_t ld, int jb) {
    const int64_t pitch = std::min<int64_t>(ld, WIN_ROWS);
    return (size_t)NUM_WARPS * jb * pitch * 8 + NUM_WARPS * sizeof(uint64_t) + (size_t)NUM_WARPS * pitch * 8;
}

// selection chain of steps [k0, k0 + cnt) (their slots k-1 are in the archive already) ...
static void process_chain(Flowpipe* F, GP& g, int k0, int cnt, cudaStream_t st) {
    F->ctx->launches += launch_chain_auto(g, k0, cnt, F->chain_use_smem != 0, F->chain_smem, F->scr, st, F->chain_exclusive);
    REACH_CUDA(cudaGetLastError());
}

// ... and their batched numerics
// `tail_st` (optional): a second stream for the R reduction of the batch.  Only the W box recurrence (wseq) is sequential from
// batch to batch; the R sortperm needs nothing but the chain's snapshot, and the R gather (apply_r, finalize_r) needs the W
// box vectors of its OWN batch (`w_done`, recorded on `st` after wseq).  With the two halves on two streams the R reduction of
// batch b overlaps the W recurrence of batch b + 1 instead of queueing behind it (dense C2: the numerics stream was busy 15.4
// of 17.3 ms, its latency-bound kernels slowed 4x by the GEMM next door).  Returns the stream on which the batch completes.
static cudaStream_t process_numerics(Flowpipe* F, const GP& g, int k0, int cnt, cudaStream_t st, cudaStream_t tail_st = nullptr,
                                     cudaEvent_t w_done = nullptr) {
    const int jb = numerics_jb(F->ld);
    const size_t red_smem = numerics_smem(F->ld, jb);
    const int it = F->ld <= 256 ? 4 : (F->ld <= 512 ? 8 : 16);
    auto mark = [&]() {          // REACH_TIMING_DETAIL=1: CUDA events between the kernels (diagnostics only)
        if (!F->detail) return;
        cudaEvent_t e;
        REACH_CUDA(cudaEventCreate(&e));
        REACH_CUDA(cudaEventRecord(e, st));
        F->detail_ev.push_back(e);
    };
    mark();
    if (it == 4) wbox_partial_kernel<4><<<dim3(NBW, (unsigned)cnt), NUM_WARPS * 32, red_smem, st>>>(g, k0, jb);
    else if (it == 8) wbox_partial_kernel<8><<<dim3(NBW, (unsigned)cnt), NUM_WARPS * 32, red_smem, st>>>(g, k0, jb);
    else wbox_partial_kernel<16><<<dim3(NBW, (unsigned)cnt), NUM_WARPS * 32, red_smem, st>>>(g, k0, jb);
    mark();
    wseq_fast_kernel<<<(unsigned)ceil_div(F->n * 32, 256), 256, 0, st>>>(g, k0, cnt);
    wseq_kernel<<<(unsigned)ceil_div(F->n, 128), 128, 0, st>>>(g, k0, cnt);
    mark();
    cudaStream_t rs = st;
    if (tail_st && tail_st != st) {
        REACH_CUDA(cudaEventRecord(w_done, st));
        rs = tail_st;
        rank_r_kernel<<<dim3((unsigned)ceil_div(g.p0 + g.capW, 256), (unsigned)cnt), 256, 0, rs>>>(g, k0);
        REACH_CUDA(cudaStreamWaitEvent(rs, w_done, 0));
    } else {
        rank_r_kernel<<<dim3((unsigned)ceil_div(g.p0 + g.capW, 256), (unsigned)cnt), 256, 0, st>>>(g, k0);
    }
    mark();
    const dim3 ag(NBR + COPY_CTAS, (unsigned)cnt);
    if (it == 4) apply_r_kernel<4><<<ag, NUM_WARPS * 32, red_smem, rs>>>(g, k0, jb);
    else if (it == 8) apply_r_kernel<8><<<ag, NUM_WARPS * 32, red_smem, rs>>>(g, k0, jb);
    else apply_r_kernel<16><<<ag, NUM_WARPS * 32, red_smem, rs>>>(g, k0, jb);
    mark();
    finalize_r_kernel<<<(unsigned)cnt, 256, 0, rs>>>(g, k0);
    mark();
    REACH_CUDA(cudaGetLastError());
    F->ctx->launches += 6;
    return rs;
}

static void set_kernel_attributes(Flowpipe* F) {
    static bool done_on[64] = {false};         // function attributes are per device
    static std::mutex done_mutex;
    std::lock_guard<std::mutex> done_lock(done_mutex);
    bool& done = done_on[F->ctx->device & 63];
    if (done) return;
    REACH_CUDA(cudaFuncSetAttribute(chain_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    REACH_CUDA(cudaFuncSetAttribute(chain_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    REACH_CUDA(cudaFuncSetAttribute(sort_segments_bitonic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BITONIC_MAX * 12));
    const int num_smem = 112 * 1024;        // 4 warps x 16 KB stages + scratch for n <= 1024
    REACH_CUDA(cudaFuncSetAttribute(wbox_partial_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, num_smem));
    REACH_CUDA(cudaFuncSetAttribute(wbox_partial_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, num_smem));
    REACH_CUDA(cudaFuncSetAttribute(wbox_partial_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, num_smem));
    REACH_CUDA(cudaFuncSetAttribute(apply_r_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, num_smem));
    REACH_CUDA(cudaFuncSetAttribute(apply_r_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, num_smem));
    REACH_CUDA(cudaFuncSetAttribute(apply_r_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, num_smem));
    done = true;
}

// SMs the persistent GEMM leaves to the selection chain and the HBM-bound numerics while they overlap it: balances t_gemm at
// the DMMA peak against t_numerics at ~4 TB/s, scaled by 0.5 because the numerics CTAs that land next to GEMM CTAs still
// contribute (C2: K = 33 gave 106 k steps/s against 89 k with K = 0 in round 1; the dynamic tile scheduler of the GEMM makes
// any grid size free of wave quantisation).
static int partition_reserve(const Flowpipe* F, int64_t q_gemm, int64_t gens_numerics, bool chain_exclusive) {
    const double t_gemm = 2.0 * (double)F->n * F->n * q_gemm / 37.0e12;
    // HBM-side work per step since the reach sets are generator references: the score pass over the mapped generators and
    // the box sums over the ~(p0 + pV + n) boxed generators of the two reductions
    const double t_num = 8.0 * (double)F->ld * (2.0 * gens_numerics + F->n) / 4.0e12;
    int K = (int)(0.5 * F->ctx->sm_count * t_num / (t_gemm + t_num) + 0.5) + (chain_exclusive ? 1 : 0);
    return std::max(chain_exclusive ? 2 : 0, std::min(K, F->ctx->sm_count / 2));
}

void flowpipe_run(Flowpipe* F) {
    Ctx* ctx = F->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t n = F->n, ld = F->ld, q = F->qloc, N = F->N, s = F->s;
    REACH_REQUIRE(F->nranks == 1, "glgm06: a column-sharded flowpipe is driven through reach_glgm06_shard_phase");
    GP g = F->g;
    size_t ev_used = 0;
    F->gemm_launches = 0; F->gemm_flops = 0;
    REACH_CUDA(cudaEventRecord(ctx->ev0, st));
    std::vector<std::pair<size_t, size_t>> gemm_ev, red_ev;
    int lvl = 0;        // doubling level of the current power Phi^(2^lvl)
    auto timed_gemm = [&](int64_t Q, const DMat& P, const double* Xp, double* Yp) {
        cudaEvent_t e0 = next_event(F, ev_used), e1 = next_event(F, ev_used);
        gemm_ev.push_back({ev_used - 2, ev_used - 1});
        REACH_CUDA(cudaEventRecord(e0, st));
        const bool sparse = map_slots(F, lvl, P, Q, Xp, Yp);
        REACH_CUDA(cudaEventRecord(e1, st));
        F->gemm_launches++;
        if (!sparse) F->gemm_flops += 2.0 * (double)n * (double)n * (double)Q;
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
    struct ReserveGuard {       // one SM stays free for the selection chain while the GEMMs of later batches run
        Ctx* c; int old;
        ReserveGuard(Ctx* c_, int v) : c(c_), old(c_->gemm_reserve_sms) { c->gemm_reserve_sms = v; }
        ~ReserveGuard() { c->gemm_reserve_sms = old; c->gemm_exclusive = false; }
    } reserve(ctx, 0);
    if (!F->homog && !F->serial) {
        // SM partition of the pipelined run.  The persistent GEMM is issue-bound, so kernels that share its SMs crawl (a
        // 64-step batch of numerics took 1.8 ms next to it against 0.24 ms alone) and the chain lost half its speed.  The
        // GEMM therefore leaves K SMs free: one is owned by the selection chain (which asks for so much shared memory that
        // nothing fits beside it), the others run the HBM-bound numerics at full occupancy.  K balances the two: t_gemm at
        // the DMMA peak against t_numerics at ~4 TB/s, scaled by 0.5 because the numerics CTAs that land next to GEMM CTAs
        // still contribute.  (C2: K = 33 gave 106 k steps/s against 89 k with K = 0; the dynamic tile scheduler of the GEMM
        // makes any grid size free of wave quantisation.)  REACH_GEMM_RESERVE / REACH_CHAIN_EXCLUSIVE override (diagnostics).
        const char* e = getenv("REACH_CHAIN_EXCLUSIVE");
        F->chain_exclusive = F->chain_use_smem && (!e || atoi(e) != 0);
        const int K = partition_reserve(F, F->q, F->p0 + F->pV, F->chain_exclusive);
        const char* r = getenv("REACH_GEMM_RESERVE");
        ctx->gemm_reserve_sms = r ? atoi(r) : K;
        const char* x = getenv("REACH_GEMM_EXCLUSIVE");
        ctx->gemm_exclusive = x && atoi(x) != 0;
    }
    std::vector<std::pair<size_t, size_t>> chain_ev;
    F->detail = getenv("REACH_TIMING_DETAIL") != nullptr;
    if (F->homog && !F->phi_host.empty() && N > 1) {
        const int nn = (int)n;
        const int P = homog_pitch((int)ld, nn <= 8 ? 2 : nn <= 16 ? 4 : nn <= 24 ? 6 : nn <= 28 ? 7 : 8);
        // (W warps per CTA, NT column tiles per warp): the fewest waves of one-CTA-per-SM, then the least slack in the last wave,
        // then the most warps (latency hiding).  Budget: 65536 registers per SM at <= 107 per thread -> up to 19 warps.
        const int kt_tiles = nn <= 8 ? 2 : nn <= 16 ? 4 : nn <= 24 ? 6 : nn <= 28 ? 7 : 8;
        const size_t a_bytes = (size_t)4 * kt_tiles * HOMOG_PA * sizeof(double);
        int bestW = 4, nt = 1;
        double best_cost = 1e300;
        for (int cand_nt = 1; cand_nt <= 4; ++cand_nt)
            for (int W = 4; W <= 20; ++W) {          // __launch_bounds__(640): at most 20 warps, <= 102 registers
                if (a_bytes + (size_t)W * 2 * 8 * cand_nt * P * sizeof(double) > 200 * 1024) continue;
                const int64_t ctas = ceil_div(q, (int64_t)W * 8 * cand_nt);
                const int64_t waves = ceil_div(ctas, ctx->sm_count);
                const double cost = (double)waves * W * cand_nt + 1e-2 * waves - 1e-3 * W - 1e-4 * cand_nt;
                if (cost < best_cost) { best_cost = cost; bestW = W; nt = cand_nt; }
            }
        const unsigned grid = (unsigned)ceil_div(q, (int64_t)bestW * 8 * nt), block = (unsigned)bestW * 32;
        const size_t smem = a_bytes + (size_t)bestW * 2 * 8 * nt * P * sizeof(double);
#define REACH_HOMOG_LAUNCH(MT, KT, NT)                                                                                          \
    do {                                                                                                                       \
        REACH_CUDA(cudaFuncSetAttribute(homog_small_dmma_kernel<MT, KT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
        homog_small_dmma_kernel<MT, KT, NT><<<grid, block, smem, st>>>(F->Phi.p, (int)F->Phi.ld, nn, (int)ld, q, N - 1, F->Hflow, P); \
    } while (0)
#define REACH_HOMOG_DMMA(MT, KT)                                                                                               \
    do {                                                                                                                       \
        if (nt == 4) REACH_HOMOG_LAUNCH(MT, KT, 4);                                                                            \
        else if (nt == 3) REACH_HOMOG_LAUNCH(MT, KT, 3);                                                                       \
        else if (nt == 2) REACH_HOMOG_LAUNCH(MT, KT, 2);                                                                       \
        else REACH_HOMOG_LAUNCH(MT, KT, 1);                                                                                    \
    } while (0)
        if (nn <= 8) REACH_HOMOG_DMMA(1, 2);
        else if (nn <= 16) REACH_HOMOG_DMMA(2, 4);
        else if (nn <= 24) REACH_HOMOG_DMMA(3, 6);
        else if (nn <= 28) REACH_HOMOG_DMMA(4, 7);
        else REACH_HOMOG_DMMA(4, 8);
#undef REACH_HOMOG_DMMA
#undef REACH_HOMOG_LAUNCH
        REACH_CUDA(cudaGetLastError());
        ctx->launches++;
        REACH_CUDA(cudaEventRecord(ctx->ev1, st));
        REACH_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms0 = 0;
        REACH_CUDA(cudaEventElapsedTime(&ms0, ctx->ev0, ctx->ev1));
        F->run_ms = ms0; F->gemm_ms = 0; F->reduce_ms = 0; F->chain_ms = 0;
        F->meta_host.assign(N, make_int2((int)F->p0, 0));
        F->ran = true;
        return;
    }
    if (!F->homog) {
        set_kernel_attributes(F);
        bind_set(g, F->gset[0]);
        // W_1 = V, cW_1 = cV (GLGM06/reach.jl:41-42)
        REACH_CUDA(cudaMemcpyAsync(g.cW, F->Hflow + (int64_t)(g.la + g.lc + 1) * ld, n * 8, cudaMemcpyDeviceToDevice, st));
        REACH_CUDA(cudaMemsetAsync(g.fastCount, 0, 2 * sizeof(int), st));
        score_and_sort(F, g, 0, 1, st);
        init_w_kernel<<<(unsigned)std::max<int64_t>(1, ceil_div(g.pV, 256)), 256, 0, st>>>(g);
        REACH_CUDA(cudaGetLastError());
        ctx->launches++;
        if (F->sink_G) {            // R_1 = Omega0
            cudaEvent_t ev_init = next_event(F, ev_used);
            REACH_CUDA(cudaEventRecord(ev_init, st));
            REACH_CUDA(cudaStreamWaitEvent(F->cpy, ev_init, 0));
            materialize_kernel<<<dim3((unsigned)ceil_div(F->sink_pmax, 8), 1u), 256, 0, F->cpy>>>(g, 1, 1, (int)F->sink_pmax, F->pack, (int)n,
                                                                                                   n * F->sink_pmax);
            REACH_CUDA(cudaGetLastError());
            REACH_CUDA(cudaMemcpyAsync(F->sink_G, F->pack, (size_t)n * F->sink_pmax * 8, cudaMemcpyDeviceToHost, F->cpy));
        }
    }
    // slot j = Phi^j X.  With Phi^w in hand, slots [have, have + w) = Phi^w * slots [have - w, have): doubling until
    // w = s, then s slots per GEMM.  The steps k = have+1 .. have+cnt (slot k-1) form one batch.  The map of batch b+1
    // (main stream), the scores / segment order of batch b+1 (side3), the selection chain of batch b (side), its R sortperm
    // (side4) and the numerics of batch b-1 (side2) overlap; the batch-local arrays are double-buffered, so the chain of
    // batch b waits for the numerics of batch b-2.
    const DMat* pw = &F->Phi;
    DMat* spare[2] = {&F->PowA, &F->PowB};
    int sp = 0;
    int64_t have = 1, w = 1;
    std::vector<cudaEvent_t> ev_num;     // numerics of batch b done
    int bi = 0;
    cudaStream_t sChain = F->serial ? st : F->side, sNum = F->serial ? st : F->side2, sSort = F->serial ? st : F->side3;
    while (have < N) {
        const int64_t cnt = std::min(w, N - have);
        bool scored = false;
        if (!F->homog && lvl < F->ell_levels && !F->serial) {
            // sparse Phi: the map also produces the scores, so it writes into this batch's array set and has to wait for the
            // numerics that last used it (two batches back)
            bind_set(g, F->gset[bi % F->nset]);
            if (bi >= F->nset) REACH_CUDA(cudaStreamWaitEvent(st, ev_num[bi - F->nset], 0));
            cudaEvent_t e0 = next_event(F, ev_used), e1 = next_event(F, ev_used);
            gemm_ev.push_back({ev_used - 2, ev_used - 1});
            REACH_CUDA(cudaEventRecord(e0, st));
            scored = map_slots_scored(F, lvl, g, have, (int)cnt, F->Hflow + (have - w) * q * ld, F->Hflow + have * q * ld);
            REACH_CUDA(cudaEventRecord(e1, st));
            if (scored) F->gemm_launches++;
            else { gemm_ev.pop_back(); }
        }
        if (!scored) timed_gemm(cnt * q, *pw, F->Hflow + (have - w) * q * ld, F->Hflow + have * q * ld);
        if (!F->homog) {
            cudaEvent_t ev_gemm = next_event(F, ev_used);
            REACH_CUDA(cudaEventRecord(ev_gemm, st));
            // scores / segment order: after this batch's GEMM and after the numerics that last used this array set
            REACH_CUDA(cudaStreamWaitEvent(sSort, ev_gemm, 0));
            if (bi >= F->nset) REACH_CUDA(cudaStreamWaitEvent(sSort, ev_num[bi - F->nset], 0));
            bind_set(g, F->gset[bi % F->nset]);
            if (scored) {
                g.Y = F->Hflow + have * F->qloc * ld;
                g.col0 = have * F->q;
                sort_only(F, g, (int)cnt, sSort);
            } else {
                score_and_sort(F, g, have, (int)cnt, sSort);
            }
            cudaEvent_t ev_sort = next_event(F, ev_used);
            REACH_CUDA(cudaEventRecord(ev_sort, sSort));
            REACH_CUDA(cudaStreamWaitEvent(sChain, ev_sort, 0));
            cudaEvent_t c0 = next_event(F, ev_used), c1 = next_event(F, ev_used);
            chain_ev.push_back({ev_used - 2, ev_used - 1});
            REACH_CUDA(cudaEventRecord(c0, sChain));
            process_chain(F, g, (int)have + 1, (int)cnt, sChain);
            REACH_CUDA(cudaEventRecord(c1, sChain));
            REACH_CUDA(cudaStreamWaitEvent(sNum, c1, 0));
            cudaEvent_t e0 = next_event(F, ev_used), e1 = next_event(F, ev_used);
            red_ev.push_back({ev_used - 2, ev_used - 1});
            REACH_CUDA(cudaEventRecord(e0, sNum));
            cudaStream_t sDone = sNum;
            if (F->serial || F->detail || getenv("REACH_NUMERICS_ONE_STREAM")) {
                process_numerics(F, g, (int)have + 1, (int)cnt, sNum);
            } else {
                REACH_CUDA(cudaStreamWaitEvent(F->side4, c1, 0));
                sDone = process_numerics(F, g, (int)have + 1, (int)cnt, sNum, F->side4, next_event(F, ev_used));
            }
            REACH_CUDA(cudaEventRecord(e1, sDone));
            ev_num.push_back(e1);
            if (F->sink_G) {        // stream this batch's reach sets to the host while the later batches are computed
                REACH_CUDA(cudaStreamWaitEvent(F->cpy, e1, 0));
                dim3 pg((unsigned)ceil_div(F->sink_pmax, 8), (unsigned)cnt);
                materialize_kernel<<<pg, 256, 0, F->cpy>>>(g, (int)have + 1, (int)cnt, (int)F->sink_pmax, F->pack, (int)n, n * F->sink_pmax);
                REACH_CUDA(cudaGetLastError());
                ctx->launches++;
                REACH_CUDA(cudaMemcpyAsync(F->sink_G + have * n * F->sink_pmax, F->pack, (size_t)cnt * n * F->sink_pmax * 8,
                                           cudaMemcpyDeviceToHost, F->cpy));
            }
            ++bi;
        }
        have += cnt;
        if (have < N && w < s) {
            if (F->ell_levels == 0) square(*pw, *spare[sp]);      // the ELL forms of all levels were built at creation
            pw = spare[sp];
            sp ^= 1;
            w *= 2;
            ++lvl;
        }
    }
    if (!F->homog && !ev_num.empty()) REACH_CUDA(cudaStreamWaitEvent(st, ev_num.back(), 0));
    if (!F->homog && (F->sink_G || F->sink_c)) {
        cudaEvent_t ev_cpy = next_event(F, ev_used);
        if (F->sink_c)
            REACH_CUDA(cudaMemcpy2DAsync(F->sink_c, n * 8, g.Rc, ld * 8, n * 8, N, cudaMemcpyDeviceToHost, F->cpy));
        REACH_CUDA(cudaEventRecord(ev_cpy, F->cpy));
        REACH_CUDA(cudaStreamWaitEvent(st, ev_cpy, 0));
    }
    REACH_CUDA(cudaEventRecord(ctx->ev1, st));
    if (F->homog) {
        F->meta_host.assign(N, make_int2((int)F->p0, 0));
    } else {
        F->meta_host.resize(N);
        REACH_CUDA(cudaMemcpyAsync(F->meta_host.data(), g.meta, sizeof(int2) * N, cudaMemcpyDeviceToHost, st));
    }
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
    if (F->detail && !F->detail_ev.empty()) {
        const char* names[5] = {"wbox_partial", "wseq", "rank_r", "apply_r", "finalize_r"};
        double acc[5] = {0, 0, 0, 0, 0};
        for (size_t i = 0; i + 5 < F->detail_ev.size(); i += 6)
            for (int j = 0; j < 5; ++j) {
                float t = 0;
                cudaEventElapsedTime(&t, F->detail_ev[i + j], F->detail_ev[i + j + 1]);
                acc[j] += t;
            }
        for (int j = 0; j < 5; ++j) fprintf(stderr, "[reach detail] %-13s %8.3f ms in situ\n", names[j], acc[j]);
        for (cudaEvent_t e : F->detail_ev) cudaEventDestroy(e);
        F->detail_ev.clear();
    }
#ifdef REACH_CHAIN_PROFILE
    {
        unsigned long long h[8];
        cudaMemcpyFromSymbol(h, g_chain_prof, sizeof(h));
        const char* nm[8] = {"prefetch+setup", "A tags", "strip union", "B rows scan", "C diag+plan", "stage next", "C search", "C merge loop"};
        for (int i = 0; i < 8; ++i) fprintf(stderr, "[chain prof] %-16s %10.1f cycles/step\n", nm[i], (double)h[i] / (double)(F->N - 1));
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_chain_prof, z, sizeof(z));
    }
#endif
    if (getenv("REACH_TIMELINE")) {     // diagnostics: start/end of every batch's GEMM, chain and numerics relative to the run start
        auto rel = [&](size_t e) { float t = 0; cudaEventElapsedTime(&t, ctx->ev0, F->evs[e]); return t; };
        for (size_t b = 0; b < chain_ev.size(); ++b)
            fprintf(stderr, "[reach timeline] batch %3zu chain %8.3f - %8.3f   numerics %8.3f - %8.3f\n", b,
                    rel(chain_ev[b].first), rel(chain_ev[b].second), rel(red_ev[b].first), rel(red_ev[b].second));
        for (size_t b = 0; b < gemm_ev.size(); ++b)
            fprintf(stderr, "[reach timeline] gemm %3zu %8.3f - %8.3f\n", b, rel(gemm_ev[b].first), rel(gemm_ev[b].second));
    }
    F->chain_ms = 0;
    for (auto& pr : chain_ev) {
        float t = 0;
        REACH_CUDA(cudaEventElapsedTime(&t, F->evs[pr.first], F->evs[pr.second]));
        F->chain_ms += t;
    }
    F->ran = true;
}

// ---------------------------------------------------------------------------------------------------------------
// Column-sharded flowpipe (multi-GPU): every rank maps, scores and gathers only the generators it
// owns; the selection chain is key logic and runs replicated.  One batch = four phases with a sum all-reduce of a small
// buffer between them (issued by the caller with torch.distributed / NCCL over NVLink):
//   phase 0  GEMM of the batch on the own columns + their scores          -> all-reduce scores (and row-support masks)
//   phase 1  segment order, selection chain, own part of the W box sums   -> all-reduce W box partials
//   phase 2  W box recurrence + centres, R ranks, own part of the R gather -> all-reduce R box partials
//   phase 3  R box vectors, counts, dense box generators (rank 0)
// Batch 0 is the initialisation (W_1 = V); batches must be processed in order.
// ---------------------------------------------------------------------------------------------------------------
static void shard_build_schedule(Flowpipe* F) {
    F->sched.clear();
    int64_t have = 1, w = 1;
    while (have < F->N) {
        const int64_t cnt = std::min(w, F->N - have);
        Flowpipe::Batch b{have, cnt, w, false};
        have += cnt;
        if (have < F->N && w < F->s) { b.square_after = true; w *= 2; }
        F->sched.push_back(b);
    }
}

int64_t flowpipe_shard_nbatches(Flowpipe* F) {
    if (F->sched.empty() && F->N > 1) shard_build_schedule(F);
    return 1 + (int64_t)F->sched.size();
}

// `on` (optional): the CUDA stream the phase is issued on instead of the context's.  Phase 0 of batch b + 1 (the map of the
// own columns and their scores -- the DMMA GEMM for a dense Phi) needs nothing of batch b but its slots, so a caller may issue
// it on a second stream BEFORE phases 1..3 of batch b and the two overlap; the batch-local arrays exist NSET times and the
// library orders their reuse with events (phase 0 of batch b waits for phase 3 of batch b - NSET, which must have been issued).
void flowpipe_shard_phase(Flowpipe* F, int64_t b, int phase, cudaStream_t on) {
    REACH_REQUIRE(!F->homog, "glgm06: homogeneous flowpipes are not column-sharded");
    const int64_t nb = flowpipe_shard_nbatches(F);
    REACH_REQUIRE(b >= 0 && b < nb && phase >= 0 && phase < 4, "glgm06: bad batch / phase");
    Ctx* ctx = F->ctx;
    struct StreamGuard {            // the GEMM / map launchers issue on the context's stream
        Ctx* c; cudaStream_t old; int old_reserve;
        StreamGuard(Ctx* c_, cudaStream_t s) : c(c_), old(c_->stream), old_reserve(c_->gemm_reserve_sms) { if (s) c->stream = s; }
        ~StreamGuard() { c->stream = old; c->gemm_reserve_sms = old_reserve; }
    } guard(ctx, on);
    cudaStream_t st = ctx->stream;
    const int64_t n = F->n, ld = F->ld, N = F->N;
    set_kernel_attributes(F);
    if ((int64_t)F->shard_done.size() < nb) {
        const size_t have_ev = F->shard_done.size();
        F->shard_done.resize(nb);
        for (size_t i = have_ev; i < (size_t)nb; ++i) REACH_CUDA(cudaEventCreateWithFlags(&F->shard_done[i], cudaEventDisableTiming));
    }
    GP g = F->g;
    const int si = (int)(b % F->nset);
    bind_set(g, F->gset[si]);
    if (phase == 0) {
        REACH_REQUIRE(b == 0 || b == F->next_p0, "glgm06: phase 0 of the sharded batches must be issued in order");
        if (b >= F->nset) {
            REACH_REQUIRE(F->next_batch > b - F->nset,
                          "glgm06: phase 0 of a batch may run at most NSET - 1 batches ahead of phases 1..3");
            REACH_CUDA(cudaStreamWaitEvent(st, F->shard_done[b - F->nset], 0));
        }
    } else {
        REACH_REQUIRE(b == F->next_batch && b < F->next_p0, "glgm06: sharded batches must be processed in order (phase 0 first)");
    }
    if (b == 0) {
        g.Y = F->Hflow;
        g.col0 = 0;
        if (phase == 0) {
            F->next_batch = 0; F->next_p0 = 1; F->pw = &F->Phi; F->sp = 0; F->shard_lvl = 0; F->ran = false;
            REACH_CUDA(cudaMemsetAsync(g.fastCount, 0, 2 * sizeof(int), st));
            REACH_CUDA(cudaMemcpyAsync(g.cW, F->Hflow + (int64_t)(g.la + g.lc + 1) * ld, n * 8, cudaMemcpyDeviceToDevice, st));
            score_only(F, g, 0, 1, st);
        } else if (phase == 1) {
            sort_only(F, g, 1, st);
            init_w_kernel<<<(unsigned)std::max<int64_t>(1, ceil_div(g.pV, 256)), 256, 0, st>>>(g);
            REACH_CUDA(cudaGetLastError());
            ctx->launches++;
        } else if (phase == 3) {
            F->next_batch = 1;
            REACH_CUDA(cudaEventRecord(F->shard_done[0], st));
            if (N == 1) { F->meta_host.assign(1, make_int2((int)F->p0, 0)); F->ran = true; }
        }
        return;
    }
    const Flowpipe::Batch& B = F->sched[b - 1];
    const int k0 = (int)B.have + 1, cnt = (int)B.cnt;
    g.Y = F->Hflow + B.have * F->qloc * ld;
    g.col0 = B.have * F->q;
    const int jb = numerics_jb(ld);
    const size_t red_smem = numerics_smem(ld, jb);
    const int it = ld <= 256 ? 4 : (ld <= 512 ? 8 : 16);
    if (phase == 0) {
        if (on) {       // overlapped with phases 1..3 of the previous batch: leave them SMs (the chain runs replicated on every rank)
            const char* r = getenv("REACH_GEMM_RESERVE");
            // the selection, the segment sort and the R ranks run REPLICATED on every rank while the GEMM shrinks with the number
            // of ranks, so their share of the SMs has to grow with it (n = 1024, 8 GPUs, N = 600: K = 8 / 16 / 32 -> 37.9 / 33.9 /
            // 33.0 ms per flowpipe)
            const int k_rep = std::min(48, 8 + 3 * F->nranks);
            ctx->gemm_reserve_sms = r ? atoi(r) : std::max(k_rep, partition_reserve(F, F->qloc, (F->p0 + F->pV) / F->nranks + 1, false));
        }
        if (b >= 2 && F->sched[b - 2].square_after) {      // Phi^(2w) for this batch (doubling phase of the schedule)
            DMat* spare[2] = {&F->PowA, &F->PowB};
            if (F->ell_levels == 0)
                gemm_left(ctx, n, n, n, F->pw->p, F->pw->ld, F->pw->p, F->pw->ld, spare[F->sp]->p, spare[F->sp]->ld);
            F->pw = spare[F->sp];
            F->sp ^= 1;
            F->shard_lvl++;
        }
        map_slots(F, F->shard_lvl, *F->pw, B.cnt * F->qloc, F->Hflow + (B.have - B.w) * F->qloc * ld,
                  F->Hflow + B.have * F->qloc * ld);
        score_only(F, g, B.have, cnt, st);
        F->next_p0 = b + 1;
    } else if (phase == 1) {
        sort_only(F, g, cnt, st);
        ctx->launches += launch_chain_auto(g, k0, cnt, F->chain_use_smem != 0, F->chain_smem, F->scr, st, false) - 1;
        if (it == 4) wbox_partial_kernel<4><<<dim3(NBW, (unsigned)cnt), NUM_WARPS * 32, red_smem, st>>>(g, k0, jb);
        else if (it == 8) wbox_partial_kernel<8><<<dim3(NBW, (unsigned)cnt), NUM_WARPS * 32, red_smem, st>>>(g, k0, jb);
        else wbox_partial_kernel<16><<<dim3(NBW, (unsigned)cnt), NUM_WARPS * 32, red_smem, st>>>(g, k0, jb);
        REACH_CUDA(cudaGetLastError());
        ctx->launches += 2;
    } else if (phase == 2) {
        wseq_fast_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(g, k0, cnt);
        wseq_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, st>>>(g, k0, cnt);
        rank_r_kernel<<<dim3((unsigned)ceil_div(g.p0 + g.capW, 256), (unsigned)cnt), 256, 0, st>>>(g, k0);
        const dim3 ag(NBR + COPY_CTAS, (unsigned)cnt);
        if (it == 4) apply_r_kernel<4><<<ag, NUM_WARPS * 32, red_smem, st>>>(g, k0, jb);
        else if (it == 8) apply_r_kernel<8><<<ag, NUM_WARPS * 32, red_smem, st>>>(g, k0, jb);
        else apply_r_kernel<16><<<ag, NUM_WARPS * 32, red_smem, st>>>(g, k0, jb);
        REACH_CUDA(cudaGetLastError());
        ctx->launches += 3;
    } else {
        finalize_r_kernel<<<(unsigned)cnt, 256, 0, st>>>(g, k0);
        REACH_CUDA(cudaGetLastError());
        ctx->launches += 1;
        REACH_CUDA(cudaEventRecord(F->shard_done[b], st));
        F->next_batch = b + 1;
        if (b == nb - 1) {
            F->meta_host.resize(N);
            REACH_CUDA(cudaMemcpyAsync(F->meta_host.data(), g.meta, sizeof(int2) * N, cudaMemcpyDeviceToHost, st));
            REACH_CUDA(cudaStreamSynchronize(st));
            F->ran = true;
        }
    }
    // No host synchronisation: the phases and the caller's all-reduces are ordered on the stream(s) they are issued on (the
    // caller issues the collectives on that stream, or synchronises it itself -- sharding.run_sharded does either).
}

// Device buffer the caller all-reduces (sum) after phase `which`-dependent: 0 scores, 1 row-support masks, 2 W box
// partials, 3 R box partials.  dtype 0 = float64, 1 = int32.
void flowpipe_shard_buffer(Flowpipe* F, int64_t b, int which, void** ptr, int64_t* count, int* dtype) {
    const int64_t nb = flowpipe_shard_nbatches(F);
    REACH_REQUIRE(b >= 0 && b < nb && which >= 0 && which < 4, "glgm06: bad batch / buffer");
    const GP& h = F->gset[b % F->nset];
    const int64_t cnt = b == 0 ? 1 : F->sched[b - 1].cnt, slot0 = b == 0 ? 0 : F->sched[b - 1].have;
    *dtype = 0;
    if (which == 0) { *ptr = h.hS; *count = cnt * F->q; }
    else if (which == 1) {
        *dtype = 1;
        *ptr = F->g.maskAll + (size_t)slot0 * std::max<int64_t>(F->pV, 1) * F->g.nwords4;
        *count = F->g.strip ? cnt * F->pV * F->g.nwords4 : 0;
    }
    else if (which == 2) { *ptr = h.partW; *count = b == 0 ? 0 : cnt * NBW * F->ld; }
    else { *ptr = h.partR; *count = b == 0 ? 0 : cnt * NBR * F->ld; }
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
    Flowpipe* F = flowpipe_create(ctx, n, zeros.data(), n, 0, c, nullptr, n, p, c, G, ldg, 1, r,
                                  (flags & REACH_GLGM06_KEEP_ZERO_GENERATORS) | REACH_GLGM06_RECORD_SELECTIONS);
    double* dout = nullptr;
    try {
        set_kernel_attributes(F);
        GP g = F->g;
        bind_set(g, F->gset[0]);
        // step "k = 1" on slot 0 with an empty W
        REACH_CUDA(cudaMemsetAsync(g.wCount, 0, sizeof(int), st));
        score_and_sort(F, g, 0, 1, st);
        launch_chain(g, 1, 1, F->chain_use_smem != 0, F->chain_smem, F->scr, st);
        const int jb = numerics_jb(F->ld);
        const size_t red_smem = numerics_smem(F->ld, jb);
        wbox_partial_kernel<16><<<dim3(NBW, 1), NUM_WARPS * 32, red_smem, st>>>(g, 1, jb);
        wseq_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, st>>>(g, 1, 1);
        REACH_CUDA(cudaGetLastError());
        const int64_t cap = g.capW;
        REACH_CUDA(cudaMalloc(&dout, sizeof(double) * (size_t)cap * n));
        materialize_w_kernel<<<(unsigned)ceil_div(cap, 8), 256, 0, st>>>(g, dout, (int)n);
        REACH_CUDA(cudaGetLastError());
        ctx->launches += 6;
        int pW = 0;
        PlanW pl;
        REACH_CUDA(cudaMemcpyAsync(&pW, g.wCount, sizeof(int), cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaMemcpyAsync(&pl, g.planW + 1, sizeof(pl), cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaStreamSynchronize(st));
        *p_out = pW;
        if (pW > 0) REACH_CUDA(cudaMemcpy2DAsync(Gout, ldgo * 8, dout, n * 8, n * 8, pW, cudaMemcpyDeviceToHost, st));
        if (ind) {
            std::vector<int> order(std::max(pl.pin, 1));
            REACH_CUDA(cudaMemcpyAsync(order.data(), g.recOrdW + (int64_t)1 * g.capListW, sizeof(int) * pl.pin,
                                       cudaMemcpyDeviceToHost, st));
            REACH_CUDA(cudaStreamSynchronize(st));
            for (int sgm = 0; sgm < pl.pin; ++sgm) ind[sgm] = order[sgm];
            for (int64_t j = pl.pin; j < p; ++j) ind[j] = -1;
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
    catch (const std::exception& e) { (ctx)->last_error = e.what(); return REACH_ECUDA; } \
    catch (...) { (ctx)->last_error = "unknown exception"; return REACH_ECUDA; }

namespace {

void require_ran(Flowpipe* F) { REACH_REQUIRE(F->ran, "flowpipe: call reach_glgm06_run first"); }

// Reach sets [k0, k0 + cnt) (1-based) as dense generator matrices on the device: slab t = n x pmax, leading dimension n.
// Homogeneous flowpipes are stored densely already (the archive IS the flowpipe); reduced ones are materialised from their
// generator references.
void materialize_steps(Flowpipe* F, int64_t k0, int64_t cnt, int64_t pmax, double* out, cudaStream_t st) {
    const int64_t n = F->n;
    if (F->homog) {
        for (int64_t t = 0; t < cnt; ++t)
            REACH_CUDA(cudaMemcpy2DAsync(out + t * n * pmax, n * 8, F->Hflow + (k0 + t - 1) * F->q * F->ld, F->ld * 8, n * 8,
                                         std::min<int64_t>(pmax, F->p0), cudaMemcpyDeviceToDevice, st));
        return;
    }
    GP g = F->g;
    materialize_kernel<<<dim3((unsigned)ceil_div(pmax, 8), (unsigned)cnt), 256, 0, st>>>(g, (int)k0, (int)cnt, (int)pmax, out, (int)n,
                                                                                          n * pmax);
    REACH_CUDA(cudaGetLastError());
    F->ctx->launches++;
}

int ngens_of(Flowpipe* F, int64_t k) {
    REACH_REQUIRE(k >= 1 && k <= F->N, "flowpipe: step index out of range");
    if (F->homog) return (int)F->p0;
    const int2 m = F->meta_host[k - 1];
    return m.x + m.y;
}

const double* center_of(Flowpipe* F, int64_t k) {
    return F->homog ? F->Hflow + ((k - 1) * F->q + F->p0) * F->ld : F->g.Rc + (k - 1) * F->ld;
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

int32_t reach_glgm06_create_sharded(reach_ctx* h, int64_t n, const double* Phi, int64_t ldphi, int64_t p0, const double* c0,
                                    const double* G0, int64_t ldg0, int64_t pV, const double* cV, const double* GV,
                                    int64_t ldgv, int64_t N, int64_t max_order, int32_t flags, int32_t rank, int32_t nranks,
                                    reach_flowpipe** out) {
    if (!h || !out) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    *out = nullptr;
    FP_BEGIN(ctx)
    REACH_REQUIRE(cV != nullptr, "reach_glgm06_create_sharded: the homogeneous recurrence is sharded by members (reach_ensemble_*)");
    Flowpipe* f = flowpipe_create(ctx, n, Phi, ldphi, p0, c0, G0, ldg0, pV, cV, GV, ldgv, N, max_order, flags, 1, rank, nranks);
    *out = new reach_flowpipe{f, ctx};
    return REACH_OK;
    FP_END(ctx)
}

int32_t reach_glgm06_shard_nbatches(reach_flowpipe* fp, int64_t* nbatches) {
    if (!fp || !nbatches) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    *nbatches = flowpipe_shard_nbatches(fp->f);
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_glgm06_shard_phase(reach_flowpipe* fp, int64_t batch, int32_t phase) {
    if (!fp) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    flowpipe_shard_phase(fp->f, batch, phase, nullptr);
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_glgm06_shard_phase_on(reach_flowpipe* fp, int64_t batch, int32_t phase, void* cuda_stream) {
    if (!fp) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    flowpipe_shard_phase(fp->f, batch, phase, reinterpret_cast<cudaStream_t>(cuda_stream));
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_glgm06_shard_buffer(reach_flowpipe* fp, int64_t batch, int32_t which, void** dev_ptr, int64_t* count,
                                  int32_t* dtype) {
    if (!fp || !dev_ptr || !count || !dtype) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    int dt = 0;
    flowpipe_shard_buffer(fp->f, batch, which, dev_ptr, count, &dt);
    *dtype = dt;
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_glgm06_run(reach_flowpipe* fp) {
    if (!fp) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    flowpipe_run(fp->f);
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_glgm06_run_fetch(reach_flowpipe* fp, double* centers, double* G, int64_t pmax) {
    if (!fp) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    Flowpipe* F = fp->f;
    if (F->homog || !G) {           // nothing to overlap: run, then fetch
        flowpipe_run(F);
        return reach_flowpipe_fetch(fp, centers, G, pmax);
    }
    REACH_REQUIRE(pmax >= 1, "reach_glgm06_run_fetch: pmax must be positive");
    if (!F->cpy) REACH_CUDA(cudaStreamCreateWithFlags(&F->cpy, cudaStreamNonBlocking));
    F->pack = (double*)ctx_malloc(F->ctx, (size_t)F->s * F->n * pmax * 8);
    F->sink_c = centers; F->sink_G = G; F->sink_pmax = pmax;
    try {
        flowpipe_run(F);
    } catch (...) {
        F->sink_c = F->sink_G = nullptr;
        ctx_free(F->ctx, F->pack); F->pack = nullptr;
        throw;
    }
    F->sink_c = F->sink_G = nullptr;
    ctx_free(F->ctx, F->pack); F->pack = nullptr;
    for (int64_t k = 0; k < F->N; ++k)
        REACH_REQUIRE(F->meta_host[k].x + F->meta_host[k].y <= pmax,
                      "reach_glgm06_run_fetch: pmax smaller than a reach set's generator count (that set is truncated)");
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
    const int p = ngens_of(F, k);
    cudaStream_t st = F->ctx->stream;
    if (c) REACH_CUDA(cudaMemcpyAsync(c, center_of(F, k), F->n * 8, cudaMemcpyDeviceToHost, st));
    if (G && p > 0) {
        REACH_REQUIRE(ldg >= F->n, "reach_flowpipe_step: ldg < n");
        DevScratch scr(F->ctx);
        double* d = scr.get<double>((size_t)F->n * p);
        materialize_steps(F, k, 1, p, d, st);
        REACH_CUDA(cudaMemcpy2DAsync(G, ldg * 8, d, F->n * 8, F->n * 8, p, cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaStreamSynchronize(st));
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
    const int64_t n = F->n, N = F->N;
    if (centers) {
        if (F->homog) {
            for (int64_t k = 1; k <= N; ++k)
                REACH_CUDA(cudaMemcpyAsync(centers + (k - 1) * n, center_of(F, k), n * 8, cudaMemcpyDeviceToHost, st));
        } else {
            REACH_CUDA(cudaMemcpy2DAsync(centers, n * 8, F->g.Rc, F->ld * 8, n * 8, N, cudaMemcpyDeviceToHost, st));
        }
    }
    if (G) {
        int64_t need = 0;
        for (int64_t k = 1; k <= N; ++k) need = std::max<int64_t>(need, ngens_of(F, k));
        REACH_REQUIRE(need <= pmax, "reach_flowpipe_fetch: pmax smaller than a reach set's generator count");
        // whole slabs (columns beyond a set's count are zero), a chunk of steps at a time through a device staging buffer
        const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(N, (int64_t)(256 << 20) / std::max<int64_t>(1, n * pmax * 8)));
        DevScratch scr(F->ctx);
        double* d = scr.get<double>((size_t)chunk * n * pmax);
        for (int64_t k0 = 1; k0 <= N; k0 += chunk) {
            const int64_t cnt = std::min(chunk, N - k0 + 1);
            if (F->homog) REACH_CUDA(cudaMemsetAsync(d, 0, (size_t)cnt * n * pmax * 8, st));
            materialize_steps(F, k0, cnt, pmax, d, st);
            REACH_CUDA(cudaMemcpyAsync(G + (k0 - 1) * n * pmax, d, (size_t)cnt * n * pmax * 8, cudaMemcpyDeviceToHost, st));
            REACH_CUDA(cudaStreamSynchronize(st));       // the staging buffer is reused by the next chunk
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
    const int p = ngens_of(F, k);
    cudaStream_t st = F->ctx->stream;
    DevScratch scr(F->ctx);
    double* d = scr.get<double>((size_t)F->n * std::max(p, 1));
    double* stage = scr.get<double>((size_t)2 * F->n);
    materialize_steps(F, k, 1, p, d, st);
    zono_box_kernel<<<dim3((unsigned)ceil_div(F->n, 256), 1), 256, 0, st>>>((int)F->n, (int)F->n, p, 0, 1, d, center_of(F, k), 0, stage,
                                                                             stage + F->n);
    REACH_CUDA(cudaGetLastError());
    F->ctx->launches++;
    REACH_CUDA(cudaMemcpyAsync(lo, stage, F->n * 8, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaMemcpyAsync(hi, stage + F->n, F->n * 8, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaStreamSynchronize(st));
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_flowpipe_project(reach_flowpipe* fp, int64_t q, const double* M, int64_t ldm, double* centers, double* radii,
                               double* Gp, int64_t pmax) {
    if (!fp || !M) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    Flowpipe* F = fp->f;
    require_ran(F);
    REACH_REQUIRE(q >= 1 && q <= 4 && ldm >= q, "reach_flowpipe_project: 1 <= q <= 4 rows and ldm >= q are required");
    int64_t need = 0;
    for (const int2& m : F->meta_host) need = std::max<int64_t>(need, m.x + m.y);
    if (Gp) REACH_REQUIRE(pmax >= need, "reach_flowpipe_project: pmax smaller than a reach set's generator count");
    else pmax = std::max<int64_t>(need, 1);
    cudaStream_t st = F->ctx->stream;
    const int64_t n = F->n, N = F->N, npad = n + (n & 1);
    // M (q x n, column-major, ld ldm) -> q contiguous rows of npad doubles
    std::vector<double> mt((size_t)q * npad, 0.0);
    for (int64_t r = 0; r < q; ++r)
        for (int64_t i = 0; i < n; ++i) mt[(size_t)r * npad + i] = M[i * ldm + r];
    double* dM = (double*)ctx_malloc(F->ctx, mt.size() * 8);
    double* dG = (double*)ctx_malloc(F->ctx, (size_t)q * pmax * N * 8);
    double* dcr = (double*)ctx_malloc(F->ctx, (size_t)2 * q * N * 8);
    try {
        REACH_CUDA(cudaMemcpyAsync(dM, mt.data(), mt.size() * 8, cudaMemcpyHostToDevice, st));
        ProjIn in;
        if (F->homog) {
            in.G = F->Hflow; in.slot = F->q * F->ld; in.c = F->Hflow + F->p0 * F->ld; in.cstride = F->q * F->ld;
            in.meta = nullptr; in.p_const = (int)F->p0;
        } else {
            in.G = nullptr; in.slot = 0; in.c = F->g.Rc; in.cstride = F->ld; in.meta = F->g.meta; in.p_const = 0;
        }
        in.n = (int)n; in.ld = (int)F->ld; in.q = (int)q; in.pmax = (int)pmax; in.N = (int)N;
        project_gens_kernel<<<dim3((unsigned)ceil_div(pmax, 8), (unsigned)N), 256, 0, st>>>(in, F->g, dM, dG);
        REACH_CUDA(cudaGetLastError());
        project_box_kernel<<<(unsigned)N, 256, 0, st>>>(in, dM, dG, dcr, dcr + q * N);
        REACH_CUDA(cudaGetLastError());
        F->ctx->launches += 2;
        if (centers) REACH_CUDA(cudaMemcpyAsync(centers, dcr, (size_t)q * N * 8, cudaMemcpyDeviceToHost, st));
        if (radii) REACH_CUDA(cudaMemcpyAsync(radii, dcr + q * N, (size_t)q * N * 8, cudaMemcpyDeviceToHost, st));
        if (Gp) REACH_CUDA(cudaMemcpyAsync(Gp, dG, (size_t)q * pmax * N * 8, cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        cudaStreamSynchronize(st);
        ctx_free(F->ctx, dM); ctx_free(F->ctx, dG); ctx_free(F->ctx, dcr);
        throw;
    }
    ctx_free(F->ctx, dM); ctx_free(F->ctx, dG); ctx_free(F->ctx, dcr);
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
    int mode, pin, mm;
    if (which) {
        PlanW pl;
        REACH_CUDA(cudaMemcpyAsync(&pl, g.planW + k, sizeof(pl), cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaStreamSynchronize(st));
        mode = pl.mode; pin = pl.pin; mm = pl.m;
    } else {
        PlanR pl;
        REACH_CUDA(cudaMemcpyAsync(&pl, g.planR + k, sizeof(pl), cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaStreamSynchronize(st));
        mode = pl.mode; pin = pl.pin; mm = pl.m;
    }
    if (mode == MODE_UNCHANGED) {
        if (p_in) *p_in = 0;
        if (m) *m = 0;
        return REACH_OK;
    }
    if (p_in) *p_in = pin;
    if (m) *m = mm;
    // recorded per list element AFTER zero-generator removal, exactly LazySets' index space
    const int64_t cap = which ? g.capListW : g.capListR;
    const int* dord = (which ? g.recOrdW : g.recOrdR) + k * cap;
    const double* dh = (which ? g.recHW : g.recHR) + k * cap;
    if (ind) REACH_CUDA(cudaMemcpyAsync(ind, dord, sizeof(int) * pin, cudaMemcpyDeviceToHost, st));
    if (hout) REACH_CUDA(cudaMemcpyAsync(hout, dh, sizeof(double) * pin, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaStreamSynchronize(st));
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

int32_t reach_ensemble_boxes_all(reach_ensemble* en, double* lo, double* hi) {
    if (!en || !lo || !hi) return REACH_EINVAL;
    FP_BEGIN(en->ctx)
    Flowpipe* F = en->f;
    require_ran(F);
    cudaStream_t st = F->ctx->stream;
    const size_t cnt = (size_t)F->n * en->batch * F->N;
    DevScratch scr(F->ctx);
    double* d = scr.get<double>(2 * cnt);
    const int64_t jobs = en->batch * F->N;
    ensemble_boxes_all_kernel<<<(unsigned)ceil_div(jobs * 32, 128), 128, 0, st>>>((int)F->n, (int)F->ld, (int)en->p, (int)en->batch,
                                                                                 F->N, F->q, F->Hflow, d, d + cnt);
    REACH_CUDA(cudaGetLastError());
    F->ctx->launches++;
    REACH_CUDA(cudaMemcpyAsync(lo, d, cnt * 8, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaMemcpyAsync(hi, d + cnt, cnt * 8, cudaMemcpyDeviceToHost, st));
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

int32_t reach_flowpipe_map_info(reach_flowpipe* fp, int32_t* sparse_levels, int32_t* nnz_per_row) {
    if (!fp) return REACH_EINVAL;
    if (sparse_levels) *sparse_levels = fp->f->ell_levels;
    if (nnz_per_row) *nnz_per_row = fp->f->ell_levels ? fp->f->ell_width : 0;
    return REACH_OK;
}

int32_t reach_flowpipe_chain_info(reach_flowpipe* fp, int64_t* parallel_batches, int64_t* batches) {
    if (!fp) return REACH_EINVAL;
    FP_BEGIN(fp->ctx)
    int h[2] = {0, 0};
    if (!fp->f->homog && fp->f->g.fastCount) {
        REACH_CUDA(cudaMemcpyAsync(h, fp->f->g.fastCount, sizeof(h), cudaMemcpyDeviceToHost, fp->ctx->stream));
        REACH_CUDA(cudaStreamSynchronize(fp->ctx->stream));
    }
    if (parallel_batches) *parallel_batches = h[0];
    if (batches) *batches = h[1];
    return REACH_OK;
    FP_END(fp->ctx)
}

int32_t reach_flowpipe_set_serial(reach_flowpipe* fp, int32_t serial) {
    if (!fp) return REACH_EINVAL;
    fp->f->serial = serial != 0;
    return REACH_OK;
}

int32_t reach_flowpipe_phase_ms(reach_flowpipe* fp, double* chain_ms, double* numerics_ms) {
    if (!fp) return REACH_EINVAL;
    if (chain_ms) *chain_ms = fp->f->chain_ms;
    if (numerics_ms) *numerics_ms = fp->f->reduce_ms;
    return REACH_OK;
}

void reach_flowpipe_free(reach_flowpipe* fp) {
    if (!fp) return;
    cudaSetDevice(fp->ctx->device);
    cudaStreamSynchronize(fp->ctx->stream);
    flowpipe_free(fp->f);
    delete fp;
}

}  // extern "C"
