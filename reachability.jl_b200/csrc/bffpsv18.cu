// BFFPSV18 decomposed box post-operator on the device.
//
// Replaces the dense `reach_blocks!` loop (src/ReachSets/ContinuousPost/BFFPSV18/reach_blocks.jl:135-224)
// for Interval/Hyperrectangle block options.  Array form, rows R of the blocks of interest:
//     c_{k+1}[R] = P_k[R,:] c0 + cW_k          r_{k+1}[R] = |P_k[R,:]| r0 + rW_k              (:185-202)
//     cW_{k+1}   = cW_k + (P_k[R,:] δB) cU     rW_{k+1}   = rW_k + |P_k[R,:] δB| rU + |P_k[R,:]| rψ   (:209-215)
//     P_{k+1}    = P_k Φ                                                                      (:217-218)
//
// B200 design ("block stepping"): the recurrence is right-multiplication, so a STACK of s consecutive powers
// S = [P_{j+1}[R,:]; ...; P_{j+s}[R,:]] advances s steps with ONE large DMMA GEMM  S' = S · Φ^s  instead of
// s small dependent ones (n=1006 alone is 64 tiles on 148 SMs; the stack fills the machine).  The same GEMM
// carries the signed side products as extra columns of its B operand, B_aug = [Φ^s | δB | c0], so
// P_k δB and P_k c0 come out of the tensor pipe too.  The abs-sums |P_k| r0, |P_k| rψ are one streaming pass
// over the stack, and the (cW, rW) recurrences are a sequential scan per row (same summation order as the
// reference).  Φ^s and the first s powers are produced by repeated doubling with the same GEMM.
#include "common.cuh"
#include "../../include/reach.h"
#include <algorithm>
#include <climits>
#include <cstring>
#include <vector>

namespace reach {

struct BoxPlan {
    Ctx* ctx = nullptr;
    int64_t n = 0, m = 0, nrows = 0, N = 0, Rp = 0, s = 1;
    DMat Phi, PowA, PowB, PowT, AccA, AccB, BtAug, BtSmall, S[2], E, MU;
    double *c0 = nullptr, *r0 = nullptr, *rpsi = nullptr, *cU = nullptr, *rU = nullptr;
    double *bsum = nullptr, *fsum = nullptr, *cW = nullptr, *rW = nullptr;
    int32_t* rows = nullptr;
    double *centers = nullptr, *radii = nullptr;
    // timing of the last run
    double run_ms = 0, gemm_ms = 0, gemm_flops = 0;
    int64_t gemm_launches = 0;
    std::vector<cudaEvent_t> evs;
    // the streaming passes of block b (abs-row-sums, scan) overlap the GEMM of block b+1 on a side stream
    cudaStream_t side = nullptr;
    DMat E1;
    double *bsum1 = nullptr, *fsum1 = nullptr;
    std::vector<cudaEvent_t> sync_evs;
    // fused abs-sum epilogue of the stacked GEMM: partial |P_k| r0, |P_k| rψ per column group, double-buffered by block parity
    double* part[2] = {nullptr, nullptr};
    int64_t ngroups = 0, ldpart = 0;
    // ---- lazy-state variants: output_function (reach_blocks.jl:152-155,191-199) and check mode
    //      (check_blocks.jl:105-183).  The states stay lazy per block, so the radius is
    //      sum_b |M[:, b] P[b, :]| r0 + |M| rW   (support function of a CartesianProductArray is separable per block).
    int64_t qmap = 0, nchunks = 0;
    double* Mmap = nullptr;            // qmap x nrows, column-major
    int32_t *blk_len = nullptr;        // nrows: length of the block starting at row l, 0 if l is not a block start
    double *zpart = nullptr;           // [s][nchunks][qmap] partial sums of the current stack block
    double *out_c = nullptr, *out_r = nullptr;   // qmap x N
    bool check = false, eager = true;
    double bound = 0.0;
    long long* viol = nullptr;         // device: first violating step (1-based) or LLONG_MAX
    long long violation = 0;           // host copy after run
};

namespace {

__global__ void gather_rows_kernel(int nrows, int n, const int32_t* __restrict__ rows, const double* __restrict__ Phi,
                                   int64_t ldphi, double* __restrict__ S, int64_t lds) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (l < nrows && j < n) S[(int64_t)j * lds + l] = Phi[(int64_t)j * ldphi + rows[l]];
}

// step 1 (reach_blocks.jl:152-158) and the first input box (reach_blocks.jl:169-176)
__global__ void init_boxes_kernel(int nrows, int m, const int32_t* __restrict__ rows, const double* __restrict__ c0,
                                  const double* __restrict__ r0, const double* __restrict__ MU, int64_t ldmu,
                                  const double* __restrict__ cU, const double* __restrict__ rU,
                                  const double* __restrict__ rpsi, double* __restrict__ centers,
                                  double* __restrict__ radii, double* __restrict__ cW, double* __restrict__ rW) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nrows) return;
    const int row = rows[l];
    centers[l] = c0[row];
    radii[l] = r0[row];
    double cw = 0.0, rw = 0.0;
    for (int q = 0; q < m; ++q) {
        const double mu = MU[(int64_t)q * ldmu + row];
        cw += mu * cU[q];
        rw += fabs(mu) * rU[q];
    }
    if (m > 0) rw += rpsi[row];
    cW[l] = cw;
    rW[l] = rw;
}

// b[i] = sum_j |X[i,j]| r0[j],  f[i] = sum_j |X[i,j]| rpsi[j]   (streaming, HBM/L2-bound)
// CTA = 32 row-pairs (64 rows) x 8 column groups; columns are summed in ascending order inside a group and the
// 8 group partials in group order => deterministic.
template <bool WITH_F>
__global__ void __launch_bounds__(256) abs_rowsums_kernel(int64_t mrows, int n, const double* __restrict__ X, int64_t ldx,
                                                          const double* __restrict__ r0,
                                                          const double* __restrict__ rpsi, double* __restrict__ bsum,
                                                          double* __restrict__ fsum) {
    __shared__ double sb[8][64], sf[8][64];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * 64 + 2 * lane;   // ldx even, rows padded => 16-byte aligned double2
    double b0 = 0, b1 = 0, f0 = 0, f1 = 0;
    const double2* xp = reinterpret_cast<const double2*>(X + row);
    const int64_t ld2 = ldx / 2;
#pragma unroll 4
    for (int j = grp; j < n; j += 8) {
        const double2 v = xp[(int64_t)j * ld2];
        const double rj = r0[j];
        b0 = fma(fabs(v.x), rj, b0);
        b1 = fma(fabs(v.y), rj, b1);
        if (WITH_F) {
            const double pj = rpsi[j];
            f0 = fma(fabs(v.x), pj, f0);
            f1 = fma(fabs(v.y), pj, f1);
        }
    }
    sb[grp][2 * lane] = b0;
    sb[grp][2 * lane + 1] = b1;
    if (WITH_F) {
        sf[grp][2 * lane] = f0;
        sf[grp][2 * lane + 1] = f1;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        const int64_t r = (int64_t)blockIdx.x * 64 + threadIdx.x;
        if (r < mrows) {
            double b = 0, f = 0;
#pragma unroll
            for (int gq = 0; gq < 8; ++gq) {
                b += sb[gq][threadIdx.x];
                if (WITH_F) f += sf[gq][threadIdx.x];
            }
            bsum[r] = b;
            if (WITH_F) fsum[r] = f;
        }
    }
}

// bsum[i] = sum_q part[(2 q) ldpart + i], fsum[i] = sum_q part[(2 q + 1) ldpart + i]: the column-group partials that the stacked
// GEMM's epilogue produced (gemm_dmma.cu), added in ascending column order => deterministic.
__global__ void __launch_bounds__(256) combine_partials_kernel(int64_t mrows, int ngroups, const double* __restrict__ part,
                                                               int64_t ldpart, int with_f, double* __restrict__ bsum,
                                                               double* __restrict__ fsum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mrows) return;
    double b = 0.0, f = 0.0;
    for (int q = 0; q < ngroups; ++q) {
        b += part[(int64_t)(2 * q) * ldpart + i];
        if (with_f) f += part[(int64_t)(2 * q + 1) * ldpart + i];
    }
    bsum[i] = b;
    if (with_f) fsum[i] = f;
}

// Sequential scan over the cnt powers of one stack block, one thread per row of interest.
// E holds [P δB | P c0] for the block: E[(j*Rp + l) + q*ldE], q < m: (P δB)[:,q], q == m: P c0.
__global__ void scan_block_kernel(int nrows, int m, int64_t Rp, int cnt, int64_t first_power, int64_t N,
                                  const double* __restrict__ E, int64_t ldE, const double* __restrict__ bsum,
                                  const double* __restrict__ fsum, const double* __restrict__ cU,
                                  const double* __restrict__ rU, double* __restrict__ cW, double* __restrict__ rW,
                                  double* __restrict__ centers, double* __restrict__ radii, int add_b) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nrows) return;
    double cw = cW[l], rw = rW[l];
    for (int j = 0; j < cnt; ++j) {
        const int64_t q = first_power + j;          // this is P_q, it yields reach set q+1 (1-based), column q
        const int64_t idx = (int64_t)j * Rp + l;
        const double a = E[idx + (int64_t)m * ldE];
        const double b = bsum[idx];
        centers[q * nrows + l] = a + cw;
        radii[q * nrows + l] = add_b ? b + rw : rw;      // lazy-state variants keep rW_k alone (|M| rW term)
        if (m > 0 && q + 1 < N) {
            double dc = 0.0, dr = 0.0;
            for (int u = 0; u < m; ++u) {
                const double e = E[idx + (int64_t)u * ldE];
                dc += e * cU[u];
                dr += fabs(e) * rU[u];
            }
            cw += dc;
            rw += dr + fsum[idx];
        }
    }
    cW[l] = cw;
    rW[l] = rw;
}

// zpart[(j * nchunks + chunk) * qmap + i] = sum_{c in chunk} r0[c] * sum_b | sum_{l in b} M[i, l] S[j Rp + l, c] |
// CTA = one power j x one chunk of MAP_CHUNK columns; threads stride the rows (coalesced), products go through shared
// memory and the first row of every block sums its block (deterministic, no atomics).
constexpr int MAP_CHUNK = 32;
constexpr int MAP_Q = 8;
__global__ void __launch_bounds__(256) block_abs_map_kernel(int nrows, int n, int64_t Rp, int q, int nchunks,
                                                            const double* __restrict__ S, int64_t lds,
                                                            const double* __restrict__ M, const int32_t* __restrict__ blk_len,
                                                            const double* __restrict__ r0, double* __restrict__ zpart) {
    extern __shared__ double sm[];              // [q][nrows] products, then [256][MAP_Q] reduction
    const int chunk = blockIdx.x, j = blockIdx.y, tid = threadIdx.x;
    double acc[MAP_Q];
#pragma unroll
    for (int i = 0; i < MAP_Q; ++i) acc[i] = 0.0;
    const int c1 = min(n, (chunk + 1) * MAP_CHUNK);
    for (int c = chunk * MAP_CHUNK; c < c1; ++c) {
        const double* col = S + (int64_t)c * lds + (int64_t)j * Rp;
        const double rc = r0[c];
        for (int l = tid; l < nrows; l += 256) {
            const double v = col[l];
#pragma unroll
            for (int i = 0; i < MAP_Q; ++i)
                if (i < q) sm[i * nrows + l] = M[i + (int64_t)l * q] * v;
        }
        __syncthreads();
        for (int l = tid; l < nrows; l += 256) {
            const int len = blk_len[l];
            if (len > 0) {
#pragma unroll
                for (int i = 0; i < MAP_Q; ++i)
                    if (i < q) {
                        double t = 0.0;
                        for (int u = 0; u < len; ++u) t += sm[i * nrows + l + u];
                        acc[i] = fma(fabs(t), rc, acc[i]);
                    }
            }
        }
        __syncthreads();
    }
    double* red = sm;
#pragma unroll
    for (int i = 0; i < MAP_Q; ++i) red[tid * MAP_Q + i] = acc[i];
    __syncthreads();
    if (tid < q) {
        double t = 0.0;
        for (int u = 0; u < 256; ++u) t += red[u * MAP_Q + tid];
        zpart[((int64_t)j * nchunks + chunk) * q + tid] = t;
    }
}

// Reach sets of the powers [first_col, first_col + cnt): centre = M c_k, radius = Z + |M| rW_k   (reach_blocks.jl:197-199);
// check mode: rho(l, X_k) = centre + radius > bound  =>  violation (check_property.jl, check_blocks.jl:158-164).
__global__ void __launch_bounds__(128) map_finish_kernel(int nrows, int q, int nchunks, int cnt, int64_t first_col,
                                                         const double* __restrict__ M, const double* __restrict__ centers,
                                                         const double* __restrict__ rwk, const double* __restrict__ zpart,
                                                         double* __restrict__ out_c, double* __restrict__ out_r, int check,
                                                         double bound, long long* viol) {
    __shared__ double sc[128], sr[128];
    const int j = blockIdx.x, i = blockIdx.y, tid = threadIdx.x;
    if (j >= cnt) return;
    const int64_t col = first_col + j;
    double a = 0.0, b = 0.0;
    for (int l = tid; l < nrows; l += 128) {
        const double m = M[i + (int64_t)l * q];
        a = fma(m, centers[col * nrows + l], a);
        b = fma(fabs(m), rwk[col * nrows + l], b);
    }
    sc[tid] = a; sr[tid] = b;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (tid < o) { sc[tid] += sc[tid + o]; sr[tid] += sr[tid + o]; }
        __syncthreads();
    }
    if (tid == 0) {
        double z = 0.0;
        if (zpart) for (int ch = 0; ch < nchunks; ++ch) z += zpart[((int64_t)j * nchunks + ch) * q + i];
        const double c = sc[0], r = z + sr[0];
        out_c[col * q + i] = c;
        out_r[col * q + i] = r;
        if (check && c + r > bound) atomicMin(viol, (long long)(col + 1));
    }
}

// Same recurrences as scan_block_kernel for the latency-bound regime (few rows, long blocks: config C1): one WARP per
// row, every lane owns a contiguous chunk of the block's powers; chunk sums, a warp exclusive scan (shuffles) and a
// second pass that emits the boxes.  Summation order: chunk-local then across chunks (differs from the sequential
// order by rounding only; deterministic).
__global__ void __launch_bounds__(128) scan_block_warp_kernel(int nrows, int m, int64_t Rp, int cnt, int64_t first_power, int64_t N,
                                                              const double* __restrict__ E, int64_t ldE,
                                                              const double* __restrict__ bsum, const double* __restrict__ fsum,
                                                              const double* __restrict__ cU, const double* __restrict__ rU,
                                                              double* __restrict__ cW, double* __restrict__ rW,
                                                              double* __restrict__ centers, double* __restrict__ radii, int add_b) {
    const int l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (l >= nrows) return;
    const int chunk = (cnt + 31) / 32;
    const int j0 = lane * chunk, j1 = min(cnt, j0 + chunk);
    double sc = 0.0, sr = 0.0;
    if (m > 0)
        for (int j = j0; j < j1; ++j) {
            const int64_t q = first_power + j;
            if (q + 1 < N) {
                const int64_t idx = (int64_t)j * Rp + l;
                double dc = 0.0, dr = 0.0;
                for (int u = 0; u < m; ++u) {
                    const double e = E[idx + (int64_t)u * ldE];
                    dc += e * cU[u];
                    dr += fabs(e) * rU[u];
                }
                sc += dc;
                sr += dr + fsum[idx];
            }
        }
    double pc = sc, pr = sr;       // inclusive scan over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double vc = __shfl_up_sync(0xffffffffu, pc, o), vr = __shfl_up_sync(0xffffffffu, pr, o);
        if (lane >= o) { pc += vc; pr += vr; }
    }
    const double totc = __shfl_sync(0xffffffffu, pc, 31), totr = __shfl_sync(0xffffffffu, pr, 31);
    double cw = cW[l] + (pc - sc), rw = rW[l] + (pr - sr);
    for (int j = j0; j < j1; ++j) {
        const int64_t q = first_power + j;
        const int64_t idx = (int64_t)j * Rp + l;
        centers[q * nrows + l] = E[idx + (int64_t)m * ldE] + cw;
        radii[q * nrows + l] = add_b ? bsum[idx] + rw : rw;
        if (m > 0 && q + 1 < N) {
            double dc = 0.0, dr = 0.0;
            for (int u = 0; u < m; ++u) {
                const double e = E[idx + (int64_t)u * ldE];
                dc += e * cU[u];
                dr += fabs(e) * rU[u];
            }
            cw += dc;
            rw += dr + fsum[idx];
        }
    }
    __syncwarp();
    if (lane == 0) {
        cW[l] += totc;
        rW[l] += totr;
    }
}

double* dvec_alloc(Ctx* ctx, int64_t len, const double* host) {
    double* p = (double*)ctx_malloc(ctx, sizeof(double) * (size_t)(len > 0 ? len : 1));
    if (host && len > 0)
        REACH_CUDA(cudaMemcpyAsync(p, host, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    else
        REACH_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * (size_t)(len > 0 ? len : 1), ctx->stream));
    return p;
}

int64_t choose_block(int64_t n, int64_t Rp, int64_t N, int64_t m, int sms) {
    // enough stacked rows for ~6 waves of 128-row tiles on the SMs, capped by N-1, 1024 and ~24 GB of stacks ...
    const int64_t want_tiles = (int64_t)sms * 6;
    const int64_t tiles_n = gemm_nt_col_groups(n + m + 1) / 4;
    int64_t s0 = 1;
    while (s0 < 1024 && s0 < N - 1 && ceil_div(s0 * Rp, 128) * tiles_n < want_tiles) s0 *= 2;
    const double bytes_per_block = 2.0 * (double)Rp * (double)round_up(n, 128) * 8.0;
    while (s0 > 1 && bytes_per_block * (double)s0 > 24e9) s0 /= 2;
    if (s0 <= 2) return s0;
    // ... then ANY depth in [s0, 2 s0] (not only powers of two): the persistent GEMM hands out equal tiles, so the launch
    // takes ceil(tiles / SMs) waves -- pick the depth whose last wave is fullest and whose row padding is smallest
    // (C3: s = 16 -> 1008 tiles = 6.8 waves, 97.3 % of the tile slots used; s = 49 -> 3088 tiles = 20.9 waves, 99.3 %)
    // A depth that is not a power of two costs extra n^3 products for Phi^s and extra underfilled launches while the stack is
    // built, which matters when few rows are left per GPU (C3 on 8 GPUs: 126 rows, s0 = 128 already fills 98.8 % of its tile
    // slots): leave s0 only for a gain of more than 2 %.
    int64_t best = s0;
    double best_eff = 0.0;
    auto eff_of = [&](int64_t s) {
        const int64_t tiles = ceil_div(s * Rp, 128) * tiles_n;
        return (double)(s * Rp) / 128.0 * (double)tiles_n / (double)(ceil_div(tiles, sms) * sms);
    };
    const double eff0 = eff_of(s0);
    for (int64_t s = s0; s <= 2 * s0 && s <= N - 1 && s <= 1024 && bytes_per_block * (double)s <= 24e9; ++s) {
        const double eff = eff_of(s);
        if (eff > best_eff + 1e-3) { best_eff = eff; best = s; }
    }
    return best_eff > eff0 + 0.02 ? best : s0;
}

}  // namespace

void boxplan_destroy(BoxPlan* P);

// ---------------------------------------------------------------------------------------------------------------
// Structurally sparse Phi (sparse `reach_blocks!`, reach_blocks.jl:47-132, option :assume_sparse).  The sparse method of
// the reference skips the blocks of Phi^k that are zero.  When the sparsity graph of Phi falls into small connected
// components (decoupled subsystems: block-diagonal up to a permutation), EVERY power of Phi has the same pattern, so row l of
// P_k lives on the w <= 16 columns of its component and the whole recurrence of that row is
//     p <- p * Phi[comp, comp]                       (w^2 fma per step instead of 2 n^2)
// independent of every other row.  One thread per row of interest runs all N steps with p in registers; the sums run over
// ascending column index like the dense GEMM's k loop (the skipped terms are exact zeros).  Anything else (a component larger
// than BD_MAX) keeps the dense DMMA recurrence: Phi^k fills in and the zero blocks disappear after a few steps.
// ---------------------------------------------------------------------------------------------------------------
constexpr int BD_MAX = 16;

namespace {

struct BdRow { int w, jr, mem_off; int64_t blk_off; };      // component width, position of the row in it, offsets of its members / block

template <int W>
__global__ void __launch_bounds__(128) blockdiag_boxes_kernel(int nrows, int m, int64_t N, int wlo, const BdRow* __restrict__ info,
                                                              const int* __restrict__ members, const double* __restrict__ blocks,
                                                              const int32_t* __restrict__ rows, const double* __restrict__ c0,
                                                              const double* __restrict__ r0, const double* __restrict__ rpsi,
                                                              const double* __restrict__ MU, int64_t ldmu,
                                                              const double* __restrict__ cU, const double* __restrict__ rU,
                                                              double* __restrict__ centers, double* __restrict__ radii) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nrows) return;
    const BdRow ri = info[l];
    if (ri.w <= wlo || ri.w > W) return;              // rows are served by the narrowest instantiation that holds their component
    const int w = ri.w;
    const int* mem = members + ri.mem_off;
    const double* B = blocks + ri.blk_off;           // B[j * w + c] = Phi[mem[j], mem[c]]
    double p[W], c0w[W], r0w[W], psw[W];
    int col[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const bool on = j < w;
        col[j] = on ? mem[j] : 0;
        p[j] = on ? B[ri.jr * w + j] : 0.0;          // P_1[l, :] = Phi[row, :]
        c0w[j] = on ? c0[col[j]] : 0.0;
        r0w[j] = on ? r0[col[j]] : 0.0;
        psw[j] = (on && m > 0) ? rpsi[col[j]] : 0.0;
    }
    const int row = rows[l];
    // step 1: Xhat0[row]; What_1 = box(V)[row]  (reach_blocks.jl:152-158, 172-175)
    centers[l] = c0[row];
    radii[l] = r0[row];
    double cw = 0.0, rw = 0.0;
    for (int q = 0; q < m; ++q) {
        const double mu = MU[(int64_t)q * ldmu + row];
        cw += mu * cU[q];
        rw += fabs(mu) * rU[q];
    }
    if (m > 0) rw += rpsi[row];
    for (int64_t k = 1; k < N; ++k) {                // p = P_k[l, comp] yields reach set k + 1 (column k)
        double a = 0.0, b = 0.0, f = 0.0;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            a = fma(p[j], c0w[j], a);
            b = fma(fabs(p[j]), r0w[j], b);
            f = fma(fabs(p[j]), psw[j], f);
        }
        centers[k * nrows + l] = a + cw;
        radii[k * nrows + l] = b + rw;
        if (k + 1 < N) {
            if (m > 0) {
                double dc = 0.0, dr = 0.0;
                for (int u = 0; u < m; ++u) {
                    double e = 0.0;
#pragma unroll
                    for (int j = 0; j < W; ++j)
                        if (j < w) e = fma(p[j], MU[(int64_t)u * ldmu + col[j]], e);
                    dc += e * cU[u];
                    dr += fabs(e) * rU[u];
                }
                cw += dc;
                rw += dr + f;
            }
            double pn[W];
#pragma unroll
            for (int c = 0; c < W; ++c) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < W; ++j)
                    if (j < w && c < w) acc = fma(p[j], B[j * w + c], acc);
                pn[c] = acc;
            }
#pragma unroll
            for (int c = 0; c < W; ++c) p[c] = pn[c];
        }
    }
}

}  // namespace

// Connected components of the sparsity graph of Phi (CSC, 1-based): returns false as soon as one exceeds BD_MAX.
bool blockdiag_analyse(int64_t n, const PhiCsc& csc, std::vector<int>& comp_of, std::vector<std::vector<int>>& comps) {
    std::vector<int> parent(n), size(n, 1);
    for (int64_t i = 0; i < n; ++i) parent[i] = (int)i;
    auto find = [&](int x) {
        while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
        return x;
    };
    for (int64_t j = 0; j < n; ++j)
        for (int64_t q = csc.colptr[j] - 1; q < csc.colptr[j + 1] - 1; ++q) {
            if (csc.nzval[q] == 0.0) continue;
            const int64_t i = csc.rowval[q] - 1;
            REACH_REQUIRE(i >= 0 && i < n, "bffpsv18 (csc): row index out of range");
            int a = find((int)i), b = find((int)j);
            if (a == b) continue;
            if (size[a] < size[b]) std::swap(a, b);
            parent[b] = a;
            size[a] += size[b];
            if (size[a] > BD_MAX) return false;
        }
    comp_of.assign(n, -1);
    comps.clear();
    std::vector<int> id(n, -1);
    for (int64_t i = 0; i < n; ++i) {                // ascending i => members of every component come out sorted
        const int r = find((int)i);
        if (id[r] < 0) { id[r] = (int)comps.size(); comps.emplace_back(); }
        comp_of[i] = id[r];
        comps[id[r]].push_back((int)i);
    }
    return true;
}

// The whole flowpipe of the rows of interest for a block-diagonal Phi; results straight into the caller's host arrays.
// Returns false (nothing done) when Phi is not block-diagonal with components of at most BD_MAX states.
bool boxes_blockdiag(Ctx* ctx, int64_t n, const PhiCsc& csc, const double* c0, const double* r0, int64_t m, const double* MU,
                     int64_t ldmu, const double* cU, const double* rU, const double* rpsi, int64_t nrows, const int32_t* rows,
                     int64_t N, double* centers, double* radii, int64_t* max_component) {
    REACH_REQUIRE(n > 0 && c0 && r0 && csc.colptr && (csc.nnz == 0 || (csc.rowval && csc.nzval)), "bffpsv18 (csc): bad Phi/c0/r0");
    REACH_REQUIRE(N >= 1 && nrows > 0 && nrows <= n && rows, "bffpsv18 (csc): bad N / rows");
    if (!MU) m = 0;
    if (m > 0) REACH_REQUIRE(cU && rU && rpsi && ldmu >= n, "bffpsv18: inputs need MU, cU, rU, rpsi");
    for (int64_t i = 0; i < nrows; ++i) REACH_REQUIRE(rows[i] >= 0 && rows[i] < n, "bffpsv18: row index out of range");
    std::vector<int> comp_of;
    std::vector<std::vector<int>> comps;
    if (!blockdiag_analyse(n, csc, comp_of, comps)) return false;
    // members / dense blocks of the components that hold a row of interest
    std::vector<int> mem_off(comps.size(), -1), members;
    std::vector<int64_t> blk_off(comps.size(), -1);
    std::vector<double> blocks;
    std::vector<BdRow> info(nrows);
    std::vector<int> pos(n, -1);
    int maxw = 0;
    for (int64_t l = 0; l < nrows; ++l) {
        const int c = comp_of[rows[l]];
        const std::vector<int>& mb = comps[c];
        const int w = (int)mb.size();
        maxw = std::max(maxw, w);
        if (mem_off[c] < 0) {
            mem_off[c] = (int)members.size();
            blk_off[c] = (int64_t)blocks.size();
            members.insert(members.end(), mb.begin(), mb.end());
            blocks.resize(blocks.size() + (size_t)w * w, 0.0);
            for (int j = 0; j < w; ++j) pos[mb[j]] = j;
            double* Bk = blocks.data() + blk_off[c];
            for (int cc = 0; cc < w; ++cc) {          // column mb[cc] of Phi
                const int64_t col = mb[cc];
                for (int64_t q = csc.colptr[col] - 1; q < csc.colptr[col + 1] - 1; ++q) {
                    const int64_t i = csc.rowval[q] - 1;
                    if (csc.nzval[q] != 0.0) Bk[(size_t)pos[i] * w + cc] = csc.nzval[q];
                }
            }
        }
        info[l] = BdRow{w, pos[rows[l]], mem_off[c], blk_off[c]};
    }
    if (max_component) *max_component = maxw;
    DevScratch scr(ctx);
    cudaStream_t st = ctx->stream;
    auto up = [&](const void* src, size_t bytes) {
        void* d = scr.get<unsigned char>(bytes ? bytes : 8);
        if (bytes) REACH_CUDA(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st));
        return d;
    };
    const BdRow* dinfo = (const BdRow*)up(info.data(), info.size() * sizeof(BdRow));
    const int* dmem = (const int*)up(members.data(), members.size() * sizeof(int));
    const double* dblk = (const double*)up(blocks.data(), blocks.size() * sizeof(double));
    const int32_t* drows = (const int32_t*)up(rows, (size_t)nrows * sizeof(int32_t));
    const double* dc0 = (const double*)up(c0, (size_t)n * 8);
    const double* dr0 = (const double*)up(r0, (size_t)n * 8);
    const double *dpsi = nullptr, *dMU = nullptr, *dcU = nullptr, *drU = nullptr;
    std::vector<double> mu_packed;
    if (m > 0) {
        dpsi = (const double*)up(rpsi, (size_t)n * 8);
        mu_packed.resize((size_t)n * m);
        for (int64_t u = 0; u < m; ++u) std::memcpy(&mu_packed[(size_t)u * n], MU + u * ldmu, (size_t)n * 8);
        dMU = (const double*)up(mu_packed.data(), mu_packed.size() * 8);
        dcU = (const double*)up(cU, (size_t)m * 8);
        drU = (const double*)up(rU, (size_t)m * 8);
    }
    double* dC = scr.get<double>((size_t)nrows * N);
    double* dR = scr.get<double>((size_t)nrows * N);
    const unsigned grid = (unsigned)ceil_div(nrows, 128);
    REACH_CUDA(cudaEventRecord(ctx->ev0, st));
#define REACH_BD_LAUNCH(W, WLO)                                                                                                     \
    if (maxw > WLO) {                                                                                                               \
        blockdiag_boxes_kernel<W><<<grid, 128, 0, st>>>((int)nrows, (int)m, N, WLO, dinfo, dmem, dblk, drows, dc0, dr0, dpsi, dMU, n, \
                                                        dcU, drU, dC, dR);                                                          \
        ctx->launches++;                                                                                                            \
    }
    REACH_BD_LAUNCH(2, 0)
    REACH_BD_LAUNCH(4, 2)
    REACH_BD_LAUNCH(8, 4)
    REACH_BD_LAUNCH(16, 8)
#undef REACH_BD_LAUNCH
    REACH_CUDA(cudaGetLastError());
    REACH_CUDA(cudaEventRecord(ctx->ev1, st));
    REACH_CUDA(cudaMemcpyAsync(centers, dC, (size_t)nrows * N * 8, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaMemcpyAsync(radii, dR, (size_t)nrows * N * 8, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    REACH_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->t_loop_ms = ms;
    return true;
}

// Φ given as Julia's SparseMatrixCSC fields (1-based Int64): scattered into the zero-filled dense device matrix.  The sparse
// reach_blocks! (reach_blocks.jl:47-132) computes the same sets as the dense one -- it only skips blocks that are zero --
// and Φ^k fills in quickly, so the device keeps the dense DMMA recurrence and only the upload is sparse.
__global__ void csc_scatter_kernel(int n, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                                   const double* __restrict__ nzval, double* __restrict__ D, int64_t ld) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) D[(int64_t)j * ld + (rowval[p] - 1)] = nzval[p];
}

void upload_phi_csc(Ctx* ctx, DMat& Phi, int64_t n, const PhiCsc& csc) {
    REACH_REQUIRE(csc.nnz >= 0 && csc.colptr && (csc.nnz == 0 || (csc.rowval && csc.nzval)), "bffpsv18: bad CSC Phi");
    REACH_REQUIRE(csc.colptr[0] == 1 && csc.colptr[n] == csc.nnz + 1, "bffpsv18: CSC colptr must be 1-based with colptr[n+1] = nnz + 1");
    for (int64_t j = 0; j < n; ++j) REACH_REQUIRE(csc.colptr[j + 1] >= csc.colptr[j], "bffpsv18: CSC colptr must be non-decreasing");
    for (int64_t p = 0; p < csc.nnz; ++p)
        REACH_REQUIRE(csc.rowval[p] >= 1 && csc.rowval[p] <= n, "bffpsv18: CSC row index out of range");
    cudaStream_t st = ctx->stream;
    int64_t* dcp = (int64_t*)ctx_malloc(ctx, sizeof(int64_t) * (n + 1));
    int64_t* drv = (int64_t*)ctx_malloc(ctx, sizeof(int64_t) * std::max<int64_t>(csc.nnz, 1));
    double* dnz = (double*)ctx_malloc(ctx, sizeof(double) * std::max<int64_t>(csc.nnz, 1));
    try {
        REACH_CUDA(cudaMemcpyAsync(dcp, csc.colptr, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
        if (csc.nnz) {
            REACH_CUDA(cudaMemcpyAsync(drv, csc.rowval, sizeof(int64_t) * csc.nnz, cudaMemcpyHostToDevice, st));
            REACH_CUDA(cudaMemcpyAsync(dnz, csc.nzval, sizeof(double) * csc.nnz, cudaMemcpyHostToDevice, st));
        }
        csc_scatter_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, st>>>((int)n, dcp, drv, dnz, Phi.p, Phi.ld);
        REACH_CUDA(cudaGetLastError());
        ctx->launches++;
        REACH_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        cudaStreamSynchronize(st);
        ctx_free(ctx, dcp); ctx_free(ctx, drv); ctx_free(ctx, dnz);
        throw;
    }
    ctx_free(ctx, dcp); ctx_free(ctx, drv); ctx_free(ctx, dnz);
}

BoxPlan* boxplan_create(Ctx* ctx, int64_t n, const double* Phi, int64_t ldphi, const double* c0, const double* r0,
                        int64_t m, const double* MU, int64_t ldmu, const double* cU, const double* rU,
                        const double* rpsi, int64_t nrows, const int32_t* rows, int64_t N, const PhiCsc* csc) {
    REACH_REQUIRE(n > 0 && c0 && r0 && ((Phi && ldphi >= n) || csc), "bffpsv18: bad Phi/c0/r0");
    REACH_REQUIRE(N >= 1, "bffpsv18: N must be >= 1");
    REACH_REQUIRE(nrows > 0 && nrows <= n && rows, "bffpsv18: bad rows");
    if (!MU) m = 0;
    REACH_REQUIRE(m >= 0, "bffpsv18: bad m");
    if (m > 0) REACH_REQUIRE(cU && rU && rpsi && ldmu >= n, "bffpsv18: inputs need MU, cU, rU, rpsi");
    for (int64_t i = 0; i < nrows; ++i) REACH_REQUIRE(rows[i] >= 0 && rows[i] < n, "bffpsv18: row index out of range");
    BoxPlan* P = new BoxPlan();
    try {
        P->ctx = ctx; P->n = n; P->m = m; P->nrows = nrows; P->N = N;
        P->Rp = round_up(nrows, 8);
        P->s = N > 2 ? choose_block(n, P->Rp, N, m, ctx->sm_count) : 1;
        if (const char* e = getenv("REACH_BFFPSV18_BLOCK"))         // diagnostics: powers per stacked GEMM (any depth >= 1)
            if (N > 2) P->s = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(1024, N - 1), atoi(e)));
        const int64_t s = P->s;
        P->Phi = dmat_alloc(ctx, n, n);
        if (Phi) dmat_upload(ctx, P->Phi, Phi, ldphi);
        else upload_phi_csc(ctx, P->Phi, n, *csc);
        P->c0 = dvec_alloc(ctx, n, c0);
        P->r0 = dvec_alloc(ctx, n, r0);
        if (m > 0) {
            P->MU = dmat_alloc(ctx, n, m);
            dmat_upload(ctx, P->MU, MU, ldmu);
            P->cU = dvec_alloc(ctx, m, cU);
            P->rU = dvec_alloc(ctx, m, rU);
            P->rpsi = dvec_alloc(ctx, n, rpsi);
        }
        P->rows = (int32_t*)ctx_malloc(ctx, sizeof(int32_t) * nrows);
        REACH_CUDA(cudaMemcpyAsync(P->rows, rows, sizeof(int32_t) * nrows, cudaMemcpyHostToDevice, ctx->stream));
        P->centers = dvec_alloc(ctx, nrows * N, nullptr);
        P->radii = dvec_alloc(ctx, nrows * N, nullptr);
        P->cW = dvec_alloc(ctx, nrows, nullptr);
        P->rW = dvec_alloc(ctx, nrows, nullptr);
        if (N >= 2) {
            if (s > 1) {
                P->PowA = dmat_alloc(ctx, n, n);
                P->PowB = dmat_alloc(ctx, n, n);
                if (s & (s - 1)) {      // Phi^s of a depth that is not a power of two: product of the squares at its set bits
                    P->AccA = dmat_alloc(ctx, n, n);
                    P->AccB = dmat_alloc(ctx, n, n);
                }
            }
            P->PowT = dmat_alloc(ctx, n, n);
            P->BtAug = dmat_alloc(ctx, n + m + 1, n);
            P->BtSmall = dmat_alloc(ctx, m + 1, n);
            P->S[0] = dmat_alloc(ctx, s * P->Rp, n);
            if ((N - 1) > s) P->S[1] = dmat_alloc(ctx, s * P->Rp, n);
            P->E = dmat_alloc(ctx, s * P->Rp, m + 1);
            P->bsum = dvec_alloc(ctx, s * P->Rp + 128, nullptr);
            P->fsum = dvec_alloc(ctx, s * P->Rp + 128, nullptr);
            if ((N - 1) > s) {      // more than one block: double-buffer what the side stream consumes
                P->E1 = dmat_alloc(ctx, s * P->Rp, m + 1);
                P->bsum1 = dvec_alloc(ctx, s * P->Rp + 128, nullptr);
                P->fsum1 = dvec_alloc(ctx, s * P->Rp + 128, nullptr);
                REACH_CUDA(cudaStreamCreateWithFlags(&P->side, cudaStreamNonBlocking));
                P->ngroups = gemm_nt_col_groups(n + m + 1);
                P->ldpart = round_up(s * P->Rp, 16);
                for (int b = 0; b < 2; ++b) P->part[b] = dvec_alloc(ctx, 2 * P->ngroups * P->ldpart, nullptr);
            }
        }
        REACH_CUDA(cudaStreamSynchronize(ctx->stream));   // host pointers are not retained past this call
    } catch (...) {
        boxplan_destroy(P);
        throw;
    }
    return P;
}

void boxplan_destroy(BoxPlan* P) {
    if (!P) return;
    // the blocks go back to the context's pool at once: neither stream of this plan may still be using them
    if (P->side) cudaStreamSynchronize(P->side);
    if (P->ctx) cudaStreamSynchronize(P->ctx->stream);
    for (DMat* d : {&P->Phi, &P->PowA, &P->PowB, &P->PowT, &P->AccA, &P->AccB, &P->BtAug, &P->BtSmall, &P->S[0], &P->S[1], &P->E, &P->E1, &P->MU})
        dmat_free(*d);
    for (double* v : {P->bsum1, P->fsum1, P->part[0], P->part[1]})
        if (v) ctx_free(P->ctx, v);
    if (P->side) cudaStreamDestroy(P->side);
    for (cudaEvent_t e : P->sync_evs) cudaEventDestroy(e);
    for (double* v : {P->c0, P->r0, P->rpsi, P->cU, P->rU, P->bsum, P->fsum, P->cW, P->rW, P->centers, P->radii})
        if (v) ctx_free(P->ctx, v);
    if (P->rows) ctx_free(P->ctx, P->rows);
    for (void* v : {(void*)P->Mmap, (void*)P->blk_len, (void*)P->zpart, (void*)P->out_c, (void*)P->out_r, (void*)P->viol})
        if (v) ctx_free(P->ctx, v);
    for (cudaEvent_t e : P->evs) cudaEventDestroy(e);
    delete P;
}

void boxplan_run(BoxPlan* P) {
    Ctx* ctx = P->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t n = P->n, m = P->m, nrows = P->nrows, N = P->N, Rp = P->Rp, s = P->s;
    REACH_CUDA(cudaEventRecord(ctx->ev0, st));
    init_boxes_kernel<<<(unsigned)ceil_div(nrows, 128), 128, 0, st>>>(
        (int)nrows, (int)m, P->rows, P->c0, P->r0, P->MU.p, P->MU.ld, P->cU, P->rU, P->rpsi, P->centers, P->radii, P->cW,
        P->rW);
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
    if (P->qmap) {
        // k = 1 uses the boxed Xhat0 (reach_blocks.jl:152-155, check_blocks.jl:118-124): centre M c0[R], radius |M| r0[R]
        const long long none = LLONG_MAX;
        REACH_CUDA(cudaMemcpyAsync(P->viol, &none, sizeof(none), cudaMemcpyHostToDevice, st));
        dim3 gf(1u, (unsigned)P->qmap);
        map_finish_kernel<<<gf, 128, 0, st>>>((int)nrows, (int)P->qmap, 0, 1, 0, P->Mmap, P->centers, P->radii, nullptr,
                                              P->out_c, P->out_r, P->check ? 1 : 0, P->bound, P->viol);
        REACH_CUDA(cudaGetLastError());
        ctx->launches++;
        P->violation = 0;
        if (P->check && P->eager) {
            long long v;
            REACH_CUDA(cudaMemcpyAsync(&v, P->viol, sizeof(v), cudaMemcpyDeviceToHost, st));
            REACH_CUDA(cudaStreamSynchronize(st));
            if (v != LLONG_MAX) {
                P->violation = v;
                P->run_ms = 0;
                return;
            }
        }
    }
    P->gemm_launches = 0;
    P->gemm_flops = 0;
    size_t ev_used = 0;
    auto next_event = [&]() {
        if (ev_used == P->evs.size()) {
            cudaEvent_t e;
            REACH_CUDA(cudaEventCreate(&e));
            P->evs.push_back(e);
        }
        return P->evs[ev_used++];
    };
    if (N >= 2) {
        // P_1[R,:] = Φ[R,:]
        dim3 g((unsigned)ceil_div(nrows, 128), (unsigned)n);
        gather_rows_kernel<<<g, 128, 0, st>>>((int)nrows, (int)n, P->rows, P->Phi.p, P->Phi.ld, P->S[0].p, P->S[0].ld);
        REACH_CUDA(cudaGetLastError());
        ctx->launches++;
        // side operand [δB | c0], transposed
        if (m > 0) transpose(ctx, n, m, P->MU.p, P->MU.ld, P->BtSmall.p, P->BtSmall.ld);
        transpose(ctx, n, 1, P->c0, n, P->BtSmall.p + m, P->BtSmall.ld);
        // doubling: with the stack blocks [0, have) and Φ^have in hand, blocks [have, have + cnt) = blocks [0, cnt) · Φ^have with
        // cnt = min(have, s - have); Φ^{2 have} = Φ^have · Φ^have.  Φ^s is the product of the squares at the set bits of s
        // (for a power of two: the last square itself).
        const DMat* pow = &P->Phi;
        DMat* spare[2] = {&P->PowA, &P->PowB};
        DMat* accb[2] = {&P->AccA, &P->AccB};
        const DMat* acc = nullptr;
        int sp = 0, ap = 0;
        int64_t have = 1;
        for (int level = 0; s > 1; ++level) {
            const bool bit = ((s >> level) & 1) != 0, more = (s >> (level + 1)) != 0, grow = have < s;
            if (grow || more || (bit && acc)) transpose(ctx, n, n, pow->p, pow->ld, P->PowT.p, P->PowT.ld);
            cudaEvent_t e0 = next_event(), e1 = next_event();
            REACH_CUDA(cudaEventRecord(e0, st));
            if (grow) {
                const int64_t cnt = std::min(have, s - have);
                gemm_nt(ctx, cnt * Rp, n, n, P->S[0].p, P->S[0].ld, P->PowT.p, P->PowT.ld, P->S[0].p + have * Rp, P->S[0].ld);
                P->gemm_launches++;
                P->gemm_flops += 2.0 * (double)(cnt * nrows) * n * n;
                have += cnt;
            }
            if (bit) {
                if (!acc) {                 // lowest set bit: Φ^s starts as this square
                    if (more) {
                        REACH_CUDA(cudaMemcpyAsync(accb[ap]->p, pow->p, pow->bytes(), cudaMemcpyDeviceToDevice, st));
                        acc = accb[ap];
                        ap ^= 1;
                    } else {
                        acc = pow;          // s is a power of two: the last square is Φ^s
                    }
                } else {
                    gemm_nt(ctx, n, n, n, acc->p, acc->ld, P->PowT.p, P->PowT.ld, accb[ap]->p, accb[ap]->ld);
                    P->gemm_launches++;
                    P->gemm_flops += 2.0 * (double)n * n * n;
                    acc = accb[ap];
                    ap ^= 1;
                }
            }
            if (more) {
                gemm_nt(ctx, n, n, n, pow->p, pow->ld, P->PowT.p, P->PowT.ld, spare[sp]->p, spare[sp]->ld);
                P->gemm_launches++;
                P->gemm_flops += 2.0 * (double)n * n * n;
            }
            REACH_CUDA(cudaEventRecord(e1, st));
            if (!more) break;
            pow = spare[sp];
            sp ^= 1;
        }
        if (s > 1) pow = acc;
        // B_aug^T = [Φ^s | δB | c0]^T
        transpose(ctx, n, n, pow->p, pow->ld, P->BtAug.p, P->BtAug.ld);
        if (m > 0) transpose(ctx, n, m, P->MU.p, P->MU.ld, P->BtAug.p + n, P->BtAug.ld);
        transpose(ctx, n, 1, P->c0, n, P->BtAug.p + n + m, P->BtAug.ld);

        const int64_t npow = N - 1;
        const int64_t nblocks = ceil_div(npow, s);
        int cur = 0;
        // Overlap (plain box mode, more than one block).  The weighted abs-row-sums |P_k| r0, |P_k| rψ of block b + 1 come out
        // of the EPILOGUE of GEMM b (the accumulators of S' = S Phi^s are the entries of the next block's powers), as
        // per-column-group partials that a tiny kernel adds up; only block 0 -- produced by the doubling prologue -- takes
        // the streaming pass over its stack.  The scan of block b runs beside GEMM b + 1 on the side stream.  E, the
        // partials and bsum / fsum are double-buffered by block parity.
        const bool overlap = !P->qmap && P->side != nullptr && nblocks > 1;
        const bool fused = overlap && getenv("REACH_BFFPSV18_NO_FUSED_ABS") == nullptr;
        size_t sync_used = 0;
        auto sync_event = [&]() {
            if (sync_used == P->sync_evs.size()) {
                cudaEvent_t e;
                REACH_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                P->sync_evs.push_back(e);
            }
            return P->sync_evs[sync_used++];
        };
        std::vector<cudaEvent_t> evG, evAbs, evScan;
        if (overlap) {
            cudaEvent_t ready = sync_event();
            REACH_CUDA(cudaEventRecord(ready, st));
            REACH_CUDA(cudaStreamWaitEvent(P->side, ready, 0));
        }
        for (int64_t b = 0; b < nblocks; ++b) {
            const int64_t cnt = std::min<int64_t>(s, npow - b * s);
            const bool last = (b == nblocks - 1);
            const int par = overlap ? (int)(b & 1) : 0;
            DMat& Eb = par ? P->E1 : P->E;
            double* bsum = par ? P->bsum1 : P->bsum;
            double* fsum = par ? P->fsum1 : P->fsum;
            cudaStream_t ss = overlap ? P->side : st;
            if (overlap) {
                // GEMM b overwrites the stack abs(b-1) read (unfused / block 0) and the partials combine(b-1) read (same parity as b+1) ...
                if (b > 0) REACH_CUDA(cudaStreamWaitEvent(st, evAbs[b - 1], 0));
                if (b > 1) REACH_CUDA(cudaStreamWaitEvent(st, evScan[b - 2], 0));     // ... and the E buffer scan(b-2) read
            }
            cudaEvent_t e0 = next_event(), e1 = next_event();
            REACH_CUDA(cudaEventRecord(e0, st));
            if (!last) {
                GemmEpilogue ep;
                ep.C2 = Eb.p; ep.ldc2 = Eb.ld; ep.nsplit = n;
                if (fused) {            // partials of block b + 1
                    ep.w0 = P->r0; ep.w1 = m > 0 ? P->rpsi : nullptr;
                    ep.part = P->part[(b + 1) & 1]; ep.ldpart = P->ldpart;
                }
                gemm_nt(ctx, s * Rp, n + m + 1, n, P->S[cur].p, P->S[cur].ld, P->BtAug.p, P->BtAug.ld, P->S[cur ^ 1].p,
                        P->S[cur ^ 1].ld, ep);
                P->gemm_flops += 2.0 * (double)(s * nrows) * n * (double)(n + m + 1);
            } else {
                gemm_nt(ctx, cnt * Rp, m + 1, n, P->S[cur].p, P->S[cur].ld, P->BtSmall.p, P->BtSmall.ld, Eb.p, Eb.ld);
                P->gemm_flops += 2.0 * (double)(cnt * nrows) * n * (double)(m + 1);
            }
            REACH_CUDA(cudaEventRecord(e1, st));
            P->gemm_launches++;
            if (overlap) {
                evG.push_back(sync_event());
                REACH_CUDA(cudaEventRecord(evG[b], st));
                if (b > 0) REACH_CUDA(cudaStreamWaitEvent(ss, evG[b - 1], 0));        // S[cur] / the partials of this block are complete
            }
            if (fused && b > 0) {
                combine_partials_kernel<<<(unsigned)ceil_div(cnt * Rp, 256), 256, 0, ss>>>(cnt * Rp, (int)P->ngroups, P->part[b & 1],
                                                                                          P->ldpart, m > 0 ? 1 : 0, bsum, fsum);
            } else {
                const unsigned grid = (unsigned)ceil_div(cnt * Rp, 64);
                if (m > 0)
                    abs_rowsums_kernel<true><<<grid, 256, 0, ss>>>(cnt * Rp, (int)n, P->S[cur].p, P->S[cur].ld, P->r0, P->rpsi,
                                                                   bsum, fsum);
                else
                    abs_rowsums_kernel<false><<<grid, 256, 0, ss>>>(cnt * Rp, (int)n, P->S[cur].p, P->S[cur].ld, P->r0, nullptr,
                                                                    bsum, nullptr);
            }
            REACH_CUDA(cudaGetLastError());
            if (overlap) {
                evAbs.push_back(sync_event());
                REACH_CUDA(cudaEventRecord(evAbs[b], ss));
                REACH_CUDA(cudaStreamWaitEvent(ss, evG[b], 0));                       // E of this block is complete
            }
            if (cnt >= 64 && nrows <= 512)
                scan_block_warp_kernel<<<(unsigned)ceil_div(nrows * 32, 128), 128, 0, ss>>>(
                    (int)nrows, (int)m, Rp, (int)cnt, b * s + 1, N, Eb.p, Eb.ld, bsum, fsum, P->cU, P->rU, P->cW,
                    P->rW, P->centers, P->radii, P->qmap ? 0 : 1);
            else
                scan_block_kernel<<<(unsigned)ceil_div(nrows, 64), 64, 0, ss>>>(
                    (int)nrows, (int)m, Rp, (int)cnt, b * s + 1, N, Eb.p, Eb.ld, bsum, fsum, P->cU, P->rU, P->cW,
                    P->rW, P->centers, P->radii, P->qmap ? 0 : 1);
            REACH_CUDA(cudaGetLastError());
            ctx->launches += 2;
            if (overlap) {
                evScan.push_back(sync_event());
                REACH_CUDA(cudaEventRecord(evScan[b], ss));
            }
            if (P->qmap) {
                const size_t smem = sizeof(double) * std::max<size_t>((size_t)P->qmap * nrows, 256 * MAP_Q);
                dim3 gm((unsigned)P->nchunks, (unsigned)cnt);
                block_abs_map_kernel<<<gm, 256, smem, st>>>((int)nrows, (int)n, Rp, (int)P->qmap, (int)P->nchunks, P->S[cur].p,
                                                           P->S[cur].ld, P->Mmap, P->blk_len, P->r0, P->zpart);
                dim3 gf((unsigned)cnt, (unsigned)P->qmap);
                map_finish_kernel<<<gf, 128, 0, st>>>((int)nrows, (int)P->qmap, (int)P->nchunks, (int)cnt, b * s + 1, P->Mmap,
                                                      P->centers, P->radii, P->zpart, P->out_c, P->out_r, P->check ? 1 : 0,
                                                      P->bound, P->viol);
                REACH_CUDA(cudaGetLastError());
                ctx->launches += 2;
                if (P->check && P->eager) {        // eager_checking: stop at the first stack block with a violation
                    long long v;
                    REACH_CUDA(cudaMemcpyAsync(&v, P->viol, sizeof(v), cudaMemcpyDeviceToHost, st));
                    REACH_CUDA(cudaStreamSynchronize(st));
                    if (v != LLONG_MAX) break;
                }
            }
            cur ^= 1;
        }
        if (overlap && !evScan.empty()) REACH_CUDA(cudaStreamWaitEvent(st, evScan.back(), 0));
    }
    if (P->qmap) {
        if (P->check) {
            long long v;
            REACH_CUDA(cudaMemcpyAsync(&v, P->viol, sizeof(v), cudaMemcpyDeviceToHost, st));
            REACH_CUDA(cudaStreamSynchronize(st));
            P->violation = (v == LLONG_MAX) ? 0 : v;
        }
    }
    REACH_CUDA(cudaEventRecord(ctx->ev1, st));
    REACH_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    REACH_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    P->run_ms = ms;
    P->gemm_ms = 0;
    for (size_t i = 0; i + 1 < ev_used; i += 2) {
        float t = 0;
        REACH_CUDA(cudaEventElapsedTime(&t, P->evs[i], P->evs[i + 1]));
        P->gemm_ms += t;
    }
}

// Attach an output map (q x nrows, column-major, ldm) with the block structure of `rows` (block_sizes == NULL: 1-D
// blocks); check != 0 turns it into the half-space check  ell . x <= bound  (q must be 1).
void boxplan_attach_map(BoxPlan* P, int64_t q, const double* M, int64_t ldm, int64_t nblocks, const int32_t* block_sizes,
                        bool check, double bound, bool eager) {
    Ctx* ctx = P->ctx;
    REACH_REQUIRE(q >= 1 && M && ldm >= q, "bffpsv18: bad output map");
    if (q > MAP_Q) throw Error(REACH_EUNSUPPORTED, "bffpsv18: output maps with more than 8 rows are not supported in this build");
    const int64_t nrows = P->nrows;
    std::vector<int32_t> blen(nrows, 0);
    if (block_sizes) {
        int64_t o = 0;
        for (int64_t b = 0; b < nblocks; ++b) {
            REACH_REQUIRE(block_sizes[b] >= 1 && o + block_sizes[b] <= nrows, "bffpsv18: block sizes do not match the rows");
            blen[o] = block_sizes[b];
            o += block_sizes[b];
        }
        REACH_REQUIRE(o == nrows, "bffpsv18: block sizes do not cover the rows");
    } else {
        for (int64_t l = 0; l < nrows; ++l) blen[l] = 1;
    }
    std::vector<double> Mh((size_t)q * nrows);
    for (int64_t l = 0; l < nrows; ++l)
        for (int64_t i = 0; i < q; ++i) Mh[i + l * q] = M[i + l * ldm];
    P->qmap = q;
    P->nchunks = ceil_div(P->n, MAP_CHUNK);
    P->check = check; P->bound = bound; P->eager = eager;
    cudaStream_t st = ctx->stream;
    P->Mmap = (decltype(P->Mmap))ctx_malloc(ctx, sizeof(double) * q * nrows);
    P->blk_len = (decltype(P->blk_len))ctx_malloc(ctx, sizeof(int32_t) * nrows);
    P->zpart = (decltype(P->zpart))ctx_malloc(ctx, sizeof(double) * (size_t)P->s * P->nchunks * q);
    P->out_c = (decltype(P->out_c))ctx_malloc(ctx, sizeof(double) * q * P->N);
    P->out_r = (decltype(P->out_r))ctx_malloc(ctx, sizeof(double) * q * P->N);
    P->viol = (decltype(P->viol))ctx_malloc(ctx, sizeof(long long));
    REACH_CUDA(cudaMemsetAsync(P->out_c, 0, sizeof(double) * q * P->N, st));
    REACH_CUDA(cudaMemsetAsync(P->out_r, 0, sizeof(double) * q * P->N, st));
    REACH_CUDA(cudaMemcpyAsync(P->Mmap, Mh.data(), sizeof(double) * q * nrows, cudaMemcpyHostToDevice, st));
    REACH_CUDA(cudaMemcpyAsync(P->blk_len, blen.data(), sizeof(int32_t) * nrows, cudaMemcpyHostToDevice, st));
    const size_t smem = sizeof(double) * std::max<size_t>((size_t)q * nrows, 256 * MAP_Q);
    if (smem > 200 * 1024) throw Error(REACH_EUNSUPPORTED, "bffpsv18: output map too large for shared memory (q * nrows > 25600)");
    REACH_CUDA(cudaFuncSetAttribute(block_abs_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    REACH_CUDA(cudaStreamSynchronize(st));
}

void boxplan_fetch_map(BoxPlan* P, double* centers, double* radii, int64_t* violation) {
    Ctx* ctx = P->ctx;
    const size_t bytes = sizeof(double) * (size_t)P->qmap * (size_t)P->N;
    if (centers) REACH_CUDA(cudaMemcpyAsync(centers, P->out_c, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (radii) REACH_CUDA(cudaMemcpyAsync(radii, P->out_r, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    REACH_CUDA(cudaStreamSynchronize(ctx->stream));
    if (violation) *violation = P->violation;
}

// Rows of this plan into rows [0, nrows) of an ld_out x N column-major host array (ld_out >= nrows): the shards of a
// multi-device call land side by side in the caller's result arrays.
void boxplan_fetch_strided(BoxPlan* P, double* centers, double* radii, int64_t ld_out) {
    Ctx* ctx = P->ctx;
    REACH_REQUIRE(ld_out >= P->nrows, "bffpsv18: output leading dimension smaller than the shard");
    const size_t w = sizeof(double) * (size_t)P->nrows;
    if (centers)
        REACH_CUDA(cudaMemcpy2DAsync(centers, ld_out * sizeof(double), P->centers, w, w, P->N, cudaMemcpyDeviceToHost, ctx->stream));
    if (radii)
        REACH_CUDA(cudaMemcpy2DAsync(radii, ld_out * sizeof(double), P->radii, w, w, P->N, cudaMemcpyDeviceToHost, ctx->stream));
    REACH_CUDA(cudaStreamSynchronize(ctx->stream));
}

void boxplan_fetch(BoxPlan* P, double* centers, double* radii) {
    Ctx* ctx = P->ctx;
    const size_t bytes = sizeof(double) * (size_t)P->nrows * (size_t)P->N;
    if (centers) REACH_CUDA(cudaMemcpyAsync(centers, P->centers, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (radii) REACH_CUDA(cudaMemcpyAsync(radii, P->radii, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    REACH_CUDA(cudaStreamSynchronize(ctx->stream));
}

}  // namespace reach

namespace reach {
void boxplan_results(BoxPlan* P, double** c, double** r) {
    if (c) *c = P->centers;
    if (r) *r = P->radii;
}
void boxplan_timing(BoxPlan* P, double* run_ms, double* gemm_ms, int64_t* launches, double* flops) {
    if (run_ms) *run_ms = P->run_ms;
    if (gemm_ms) *gemm_ms = P->gemm_ms;
    if (launches) *launches = P->gemm_launches;
    if (flops) *flops = P->gemm_flops;
}
}  // namespace reach
