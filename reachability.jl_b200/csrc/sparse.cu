// Sparse-A ("lazy matrix exponential") variants of the path: Φ = e^{Aδ} is never formed.
//
//   BFFPSV18 with `:lazy_expm`  -- reach_blocks!(ϕ::SparseMatrixExp, ...)       BFFPSV18/reach_blocks.jl:227-397
//       rows of e^{kδA} come from `get_rows(ϕpowerk, bi)` (:37-38), ϕpowerk.M .= ϕpowerk.M + ϕ.M (:306)
//   discretize with `exp_method = "lazy"` -- ϕ = SparseMatrixExp(A*δ), Φ₂ through get_columns of the 3n x 3n
//       sparse exponential                                                      discretize.jl:153-154, 302-306
//
// The reference evaluates every row with a Krylov `expmv` (Expokit, tolerance 1e-7) from scratch at every step.  B200
// design: the block of rows of interest W_k = e^{kδA}[rows, :] is kept TRANSPOSED on the device (n x R, the R values of
// one state contiguous) and advanced by  W_kᵀ = e^{δAᵀ} W_{k-1}ᵀ  -- the same recurrence as the dense path
// (reach_blocks.jl:217-218) -- with a scaled truncated Taylor series (Horner form) of sparse-matrix x dense-block products
// carried to full double precision (‖δA‖/s <= 2, m terms until the remainder is below 2^-64).  Julia's SparseMatrixCSC of A *is* the
// CSR of Aᵀ, so the arrays cross the ABI untouched.  Every product is one HBM/L2-bound kernel: nnz·12 B of matrix +
// 24·n·R B of block traffic; the per-step box epilogue is a column reduction of the same block.
#include "common.cuh"
#include "../../include/reach.h"
#include <cmath>
#include <algorithm>
namespace reach {

void omega0_box(Ctx* ctx, int64_t n, const double* c, const double* r, const double* pc, const double* pr, const double* rE,
                const double* uc, const double* ur, const double* rpsi, double* c0, double* r0);   // discretize.cu

namespace {

struct Csr {
    int n = 0;
    int64_t nnz = 0;
    int* ptr = nullptr;
    int* idx = nullptr;
    double* val = nullptr;
};

// Tn = alpha * op(S) * T ;  Yout = Yin + Tn          op = |.| entrywise when ABS
// Blocks are n x ld2 double2 ("row-block" layout: the R values of state i are contiguous).  Thread = (row i, double2 l2);
// the nonzeros of a row are read by all its threads (broadcast), the gathered block rows are coalesced.
template <bool ABS>
__device__ __forceinline__ void spmm_item(const Csr& S, int ld2, const double2* T, double2* Tn, const double2* Yin, double2* Yout,
                                          double alpha, int64_t t) {
    const int i = (int)(t / ld2), l2 = (int)(t - (int64_t)i * ld2);
    double2 acc = make_double2(0.0, 0.0);
    const int p1 = __ldg(S.ptr + i + 1);
    // four nonzeros at a time: their (index, value) loads and then their four gathers are independent, so a row costs
    // three dependent memory round trips instead of two per nonzero; the sums keep CSR order
    for (int p = __ldg(S.ptr + i); p < p1; p += 4) {
        int jj[4];
        double vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool ok = p + u < p1;
            jj[u] = ok ? __ldg(S.idx + p + u) : i;
            vv[u] = ok ? __ldg(S.val + p + u) : 0.0;
        }
        double2 xx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xx[u] = T[(int64_t)jj[u] * ld2 + l2];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (p + u < p1) {
                const double v = ABS ? fabs(vv[u]) : vv[u];
                acc.x = fma(v, xx[u].x, acc.x);
                acc.y = fma(v, xx[u].y, acc.y);
            }
    }
    acc.x *= alpha;
    acc.y *= alpha;
    const int64_t o = (int64_t)i * ld2 + l2;
    if (Tn) Tn[o] = acc;
    if (Yout) {
        double2 y = Yin ? Yin[o] : make_double2(0.0, 0.0);
        y.x += acc.x;
        y.y += acc.y;
        Yout[o] = y;
    }
}

template <bool ABS>
__global__ void __launch_bounds__(256) spmm_taylor_kernel(Csr S, int ld2, const double2* T, double2* Tn, const double2* Yin,
                                                          double2* Yout, double alpha) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (int64_t)S.n * ld2) spmm_item<ABS>(S, ld2, T, Tn, Yin, Yout, alpha, t);
}

// X[i*ld + l] = (i == sel[l]) for l < cnt   (identity columns / unit rows of interest)
__global__ void unit_block_kernel(int n, int ld, int cnt, const int* __restrict__ sel, double* __restrict__ X) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * ld) return;
    const int i = (int)(t / ld), l = (int)(t - (int64_t)i * ld);
    X[t] = (l < cnt && sel[l] == i) ? 1.0 : 0.0;
}

// Row-wise box products of a row-block Z (n x cnt, ld):  sgn[i] (+)= sum_l Z[i,l] c[l],  mag[i] (+)= sum_l |Z[i,l]| w[l].
// One warp per row, lanes stride l, xor-shuffle tree (fixed order).
__global__ void __launch_bounds__(256) row_box_kernel(int n, int ld, int cnt, const double* __restrict__ Z,
                                                      const double* __restrict__ c, const double* __restrict__ w,
                                                      double* __restrict__ sgn, double* __restrict__ mag, int accumulate) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    double a = 0.0, b = 0.0;
    for (int l = lane; l < cnt; l += 32) {
        const double z = Z[(int64_t)i * ld + l];
        if (c) a = fma(z, c[l], a);
        if (w) b = fma(fabs(z), w[l], b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
        if (sgn) sgn[i] = (accumulate ? sgn[i] : 0.0) + a;
        if (mag) mag[i] = (accumulate ? mag[i] : 0.0) + b;
    }
}

__global__ void bump_kernel(int* ctr, int by) { *ctr += by; }
__global__ void gather_kernel(int cnt, const int* __restrict__ sel, const double* __restrict__ v, double* __restrict__ out) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < cnt) out[l] = v[sel[l]];
}
__global__ void sih_vec_kernel(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) s[i] = fabs(a[i]) + b[i];
}
__global__ void scale_vec_kernel(int n, double alpha, const double* __restrict__ x, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = alpha * x[i];
}

// One step of the lazy reach_blocks! recurrence for the R rows of interest (block Wt = P_{k-1}[rows,:]ᵀ, n x ld):
//   c_k = P c0 + cW,  r_k = |P| r0 + rW                              reach_blocks.jl:266-283 (box of the lazy sum)
//   cW += (P MU) cU,  rW += |P MU| rU + |P| rψ                       :284-287
// Two kernels.  `lazy_partial_kernel`: CTA (bx, by, bz) = 32 rows of interest x state slice by x quantity group bz
// (bz = 0: P c0, |P| r0, |P| rψ; bz >= 1: the products with 8 columns of MU), 8 sub-slices per CTA, ascending j inside a
// sub-slice; partial sums go to part[quantity][slice][l].  `lazy_finish_kernel` adds the slices in order and applies the
// recurrences.  Every sum has a fixed order.
constexpr int STEP_SUB = 8;
constexpr int STEP_MCHUNK = 8;
constexpr int FIN_Q = 64;

struct StepGeom {       // epilogue geometry and operands (device pointers)
    int n, ld, R, nsl, m, ldp, gx, gz;
    const double* c0; const double* r0; const double* rpsi;      // rpsi may be nullptr (homogeneous)
    const double* MU; const double* cU; const double* rU;        // MU: n x m, ld n
    double* cW; double* rW; double* part;
    const double* cbase;            // start of the centres array (column offsets of cWs / rWs follow the output column)
    double* cWs; double* rWs;       // optional (R x N): the input boxes Ŵ_k every step USES (check mode / output maps need them)
};

// virtual CTA (bx, sl, grp): 32 rows of interest x state slice sl x quantity group grp.  `red`: STEP_SUB x 32 doubles.
__device__ __forceinline__ void lazy_partial_block(const StepGeom& g, const double* Wt, int bx, int sl, int grp, double (*red)[32]) {
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int l = bx * 32 + lx;
    const bool live = l < g.R;
    const int chunk = (g.n + g.nsl - 1) / g.nsl;
    const int j0 = sl * chunk, j1 = min(g.n, j0 + chunk);
    auto combine_store = [&](double v, int quantity) {       // sub-slices in order -> part[quantity][sl][l]
        red[ly][lx] = v;
        __syncthreads();
        if (ly == 0 && live) {
            double s = 0.0;
            for (int q = 0; q < STEP_SUB; ++q) s += red[q][lx];
            g.part[((int64_t)quantity * g.nsl + sl) * g.ldp + l] = s;
        }
        __syncthreads();
    };
    if (grp == 0) {
        double sc = 0.0, sr = 0.0, sp = 0.0;
        if (live)
            for (int j = j0 + ly; j < j1; j += STEP_SUB) {
                const double w = Wt[(int64_t)j * g.ld + l];
                sc = fma(w, g.c0[j], sc);
                sr = fma(fabs(w), g.r0[j], sr);
                if (g.rpsi) sp = fma(fabs(w), g.rpsi[j], sp);
            }
        combine_store(sc, 0);
        combine_store(sr, 1);
        combine_store(sp, 2);
    } else {
        const int i0 = (grp - 1) * STEP_MCHUNK;
        double pm[STEP_MCHUNK];
#pragma unroll
        for (int u = 0; u < STEP_MCHUNK; ++u) pm[u] = 0.0;
        if (live)
            for (int j = j0 + ly; j < j1; j += STEP_SUB) {
                const double w = Wt[(int64_t)j * g.ld + l];
#pragma unroll
                for (int u = 0; u < STEP_MCHUNK; ++u)
                    if (i0 + u < g.m) pm[u] = fma(w, g.MU[(int64_t)(i0 + u) * g.n + j], pm[u]);
            }
#pragma unroll
        for (int u = 0; u < STEP_MCHUNK; ++u)
            if (i0 + u < g.m) combine_store(pm[u], 3 + i0 + u);
    }
}

// virtual CTA bx of the finish: sub-warp ly owns the quantities q = ly, ly + 8, ... of a pass (64 quantities per pass) and
// adds their state slices in slice order; sub-warp 0 then applies the input recurrences in input order.  `qs`: FIN_Q x 32.
__device__ __forceinline__ void lazy_finish_block(const StepGeom& g, int bx, double* centers, double* radii, double (*qs)[32]) {
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int l = bx * 32 + lx;
    const bool live = l < g.R;
    const int nq = 3 + g.m;
    double sc = 0.0, sr = 0.0, sp = 0.0, dc = 0.0, dr = 0.0;
    for (int q0 = 0; q0 < nq; q0 += FIN_Q) {
        for (int q = q0 + ly; q < min(nq, q0 + FIN_Q); q += STEP_SUB) {
            double s = 0.0;
            if (live) {
                const double* pq = g.part + (int64_t)q * g.nsl * g.ldp + l;
#pragma unroll 8
                for (int sl = 0; sl < g.nsl; ++sl) s += pq[(int64_t)sl * g.ldp];
            }
            qs[q - q0][lx] = s;
        }
        __syncthreads();
        if (ly == 0 && live)
            for (int q = q0; q < min(nq, q0 + FIN_Q); ++q) {
                const double v = qs[q - q0][lx];
                if (q == 0) sc = v;
                else if (q == 1) sr = v;
                else if (q == 2) sp = v;
                else {
                    dc = fma(v, g.cU[q - 3], dc);
                    dr = fma(fabs(v), g.rU[q - 3], dr);
                }
            }
        __syncthreads();
    }
    if (ly == 0 && live) {
        const double cw = g.m > 0 ? g.cW[l] : 0.0, rw = g.m > 0 ? g.rW[l] : 0.0;
        centers[l] = sc + cw;
        radii[l] = sr + rw;
        if (g.cWs) {
            g.cWs[(centers - g.cbase) + l] = cw;
            g.rWs[(centers - g.cbase) + l] = rw;
        }
        if (g.m > 0) {
            g.cW[l] = cw + dc;
            g.rW[l] = rw + dr + sp;
        }
    }
}

__global__ void __launch_bounds__(32 * STEP_SUB) lazy_partial_kernel(StepGeom g, const double* Wt) {
    __shared__ double red[STEP_SUB][32];
    lazy_partial_block(g, Wt, blockIdx.x, blockIdx.y, blockIdx.z, red);
}

// output column = device-side step counter + offset: the same captured graph serves every pair of steps
__global__ void __launch_bounds__(32 * STEP_SUB) lazy_finish_kernel(StepGeom g, double* centers, double* radii, const int* kctr,
                                                                    int koff) {
    __shared__ double qs[FIN_Q][32];
    const int64_t col = (int64_t)(*kctr + koff) * g.R;
    lazy_finish_block(g, blockIdx.x, centers + col, radii + col, qs);
}

// step 1 and the initial input boxes: centers/radii column 0 = X̂0[rows]; Ŵ_1 = box(MU U ⊕ Eψ)[rows]  (:241-262)
__global__ void lazy_init_kernel(int n, int R, const int* __restrict__ rows, const double* __restrict__ c0,
                                 const double* __restrict__ r0, const double* __restrict__ rpsi, int m,
                                 const double* __restrict__ MU, const double* __restrict__ cU, const double* __restrict__ rU,
                                 double* __restrict__ cW, double* __restrict__ rW, double* __restrict__ centers,
                                 double* __restrict__ radii) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= R) return;
    const int row = rows[l];
    centers[l] = c0[row];
    radii[l] = r0[row];
    if (m > 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < m; ++i) {
            const double v = MU[(int64_t)i * n + row];
            a = fma(v, cU[i], a);
            b = fma(fabs(v), rU[i], b);
        }
        cW[l] = a;
        rW[l] = b + rpsi[row];
    }
}

template <typename T>
T* dnew(Ctx* ctx, size_t count) { return (T*)ctx_malloc(ctx, std::max<size_t>(count, 1) * sizeof(T)); }

}  // namespace

struct SparseSys {
    Ctx* ctx = nullptr;
    int64_t n = 0, nnz = 0;
    Csr At, Ar;                 // CSR of Aᵀ (= the caller's CSC of A) and CSR of A
    double norm_row = 0, norm_col = 0;      // max abs row sum / column sum of A
    std::vector<void*> owned;
    double last_ms = 0;
    int64_t last_products = 0;
    double last_bytes = 0;
};

void sparse_free(SparseSys* S) {
    if (!S) return;
    for (void* p : S->owned) ctx_free(S->ctx, p);
    delete S;
}

SparseSys* sparse_create(Ctx* ctx, int64_t n, int64_t nnz, const int64_t* colptr, const int64_t* rowval, const double* nzval) {
    REACH_REQUIRE(n > 0 && n < (int64_t)1 << 30 && nnz >= 0 && nnz < (int64_t)1 << 31 && colptr && (nnz == 0 || (rowval && nzval)),
                  "reach_sparse_create: bad arguments");
    REACH_REQUIRE(colptr[0] == 1 && colptr[n] == nnz + 1, "reach_sparse_create: colptr must be 1-based with colptr[n+1] = nnz + 1");
    // host: 0-based int32 copies of the CSC (= CSR of Aᵀ) and the transposed structure (CSR of A), plus the norms
    std::vector<int> tp(n + 1), ti(nnz), rp(n + 1, 0), ri(nnz);
    std::vector<double> rv(nnz), rows_abs(n, 0.0);
    double norm_col = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        REACH_REQUIRE(colptr[j + 1] >= colptr[j], "reach_sparse_create: colptr must be non-decreasing");
        tp[j] = (int)(colptr[j] - 1);
        double s = 0.0;
        for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) {
            const int64_t i = rowval[p] - 1;
            REACH_REQUIRE(i >= 0 && i < n, "reach_sparse_create: row index out of range");
            REACH_REQUIRE(std::isfinite(nzval[p]), "reach_sparse_create: non-finite entry");
            ti[p] = (int)i;
            rp[i + 1]++;
            s += std::fabs(nzval[p]);
            rows_abs[i] += std::fabs(nzval[p]);
        }
        norm_col = std::max(norm_col, s);
    }
    tp[n] = (int)nnz;
    for (int64_t i = 0; i < n; ++i) rp[i + 1] += rp[i];
    std::vector<int> fill(rp.begin(), rp.end() - 1);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t p = tp[j]; p < tp[j + 1]; ++p) {
            const int q = fill[ti[p]]++;
            ri[q] = (int)j;
            rv[q] = nzval[p];
        }
    SparseSys* S = new SparseSys();
    S->ctx = ctx; S->n = n; S->nnz = nnz;
    S->norm_col = norm_col;
    S->norm_row = n ? *std::max_element(rows_abs.begin(), rows_abs.end()) : 0.0;
    try {
        cudaStream_t st = ctx->stream;
        auto up = [&](Csr& C, const std::vector<int>& p, const std::vector<int>& ix, const double* v) {
            C.n = (int)n; C.nnz = nnz;
            C.ptr = dnew<int>(ctx, n + 1); S->owned.push_back(C.ptr);
            C.idx = dnew<int>(ctx, nnz); S->owned.push_back(C.idx);
            C.val = dnew<double>(ctx, nnz); S->owned.push_back(C.val);
            REACH_CUDA(cudaMemcpyAsync(C.ptr, p.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, st));
            if (nnz) {
                REACH_CUDA(cudaMemcpyAsync(C.idx, ix.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice, st));
                REACH_CUDA(cudaMemcpyAsync(C.val, v, sizeof(double) * nnz, cudaMemcpyHostToDevice, st));
            }
        };
        up(S->At, tp, ti, nzval);
        up(S->Ar, rp, ri, rv.data());
        REACH_CUDA(cudaStreamSynchronize(st));      // the host vectors die here
    } catch (...) {
        sparse_free(S);
        throw;
    }
    return S;
}

namespace {

int64_t block_ld(int64_t R) { return round_up(std::max<int64_t>(R, 1), 4); }

template <bool ABS>
void spmm(SparseSys* S, const Csr& M, int64_t ld, const double* T, double* Tn, const double* Yin, double* Yout, double alpha) {
    const int ld2 = (int)(ld / 2);
    const int64_t threads = S->n * ld2;
    spmm_taylor_kernel<ABS><<<(unsigned)ceil_div(threads, 256), 256, 0, S->ctx->stream>>>(
        M, ld2, (const double2*)T, (double2*)Tn, (const double2*)Yin, (double2*)Yout, alpha);
    REACH_CUDA(cudaGetLastError());
    S->ctx->launches++;
    S->last_products++;
    S->last_bytes += 12.0 * (double)S->nnz + 8.0 * (double)S->n * ld * (1 + (Tn ? 1 : 0) + (Yin ? 1 : 0) + (Yout ? 1 : 0));
}

// scaling s and number of Taylor terms m for theta_total = ‖δ M‖: the pair with the fewest products s*m among
// theta = theta_total / s <= 2 (larger theta lets the alternating series cancel digits), remainder < 2^-64
int taylor_terms(double theta) {
    double term = 1.0;
    int m = 0;
    while (m < 80) {
        ++m;
        term *= theta / m;
        if (theta < m + 1 && term / (1.0 - theta / (m + 1)) < 5.4e-20) break;
    }
    return m;
}
void taylor_plan(double theta_total, int& s, int& m) {
    const int s0 = (int)std::max(1.0, std::ceil(theta_total / 2.0));
    s = s0;
    m = taylor_terms(theta_total / s0);
    for (int c = s0 + 1; c <= 4 * s0 + 4; ++c) {
        const int mc = taylor_terms(theta_total / c);
        if ((int64_t)c * mc < (int64_t)s * m) { s = c; m = mc; }
    }
}

// W <- e^{δ op(M)} W for a row-block W (n x ld); W2, T0, T1: workspaces of the same size.  Returns the buffer that holds
// the result (W or W2).  The truncated series is evaluated in HORNER form,
//     t_m = v,   t_{j-1} = v + (δ / (s j)) M t_j   (j = m .. 1),   e^{δM/s} v ≈ t_0,
// one fused product per term that gathers t_j, reads v and writes t_{j-1}: three block passes per term instead of the four
// of a running sum (term in, term out, accumulator in and out), which is what counts once the block no longer fits in L2.
double* expm_action(SparseSys* S, const Csr& M, double delta, int64_t ld, double* W, double* W2, double* T0, double* T1) {
    int s, m;
    taylor_plan(std::fabs(delta) * std::max(S->norm_row, S->norm_col), s, m);
    for (int rep = 0; rep < s; ++rep) {
        const double* t = W;                                    // t_m = v
        for (int j = m; j >= 1; --j) {
            double* out = j == 1 ? W2 : (t == T0 ? T1 : T0);    // the last term lands in the other W buffer
            spmm<false>(S, M, ld, t, nullptr, W, out, delta / (s * (double)j));
            t = out;
        }
        std::swap(W, W2);
    }
    return W;
}

// y = Φ₂(|A|, δ) v = Σ_{i>=0} δ^{i+2}/(i+2)! |A|^i v   (discretize.jl:296-301 for the nonnegative matrix |A|; all terms are
// nonnegative when v is, so the plain series is stable).  v, y, t0, t1: n x 4 row-blocks (ld 4, column 0 used).
void phi2abs_action(SparseSys* S, double delta, const double* v, double* y, double* t0, double* t1) {
    const int64_t n = S->n;
    cudaStream_t st = S->ctx->stream;
    const unsigned g = (unsigned)ceil_div(4 * n, 256);
    scale_vec_kernel<<<g, 256, 0, st>>>((int)(4 * n), delta * delta / 2.0, v, t0);        // term 0
    REACH_CUDA(cudaMemcpyAsync(y, t0, sizeof(double) * 4 * n, cudaMemcpyDeviceToDevice, st));
    const double theta = std::fabs(delta) * std::max(S->norm_row, S->norm_col);
    double bound = 1.0;     // (i+2)!-normalised size of term i relative to term 0
    double* t = t0;
    double* tn = t1;
    for (int i = 1; i < 400; ++i) {
        spmm<true>(S, S->Ar, 4, t, tn, y, y, delta / (i + 2.0));
        std::swap(t, tn);
        bound *= theta / (i + 2.0);
        if (bound < 5.4e-20 && theta / (i + 3.0) < 0.5) break;
    }
    S->ctx->launches++;
}

}  // namespace

// Y = e^{δA} X (transpose = 0) or e^{δAᵀ} X (1) for host column-major blocks X, Y (n x R).
void sparse_expmv(SparseSys* S, double delta, int transpose_op, int64_t R, const double* X, int64_t ldx, double* Y, int64_t ldy) {
    Ctx* ctx = S->ctx;
    const int64_t n = S->n, ld = block_ld(R);
    REACH_REQUIRE(R >= 1 && X && Y && ldx >= n && ldy >= n, "reach_expmv: bad arguments");
    cudaStream_t st = ctx->stream;
    std::vector<double> hb((size_t)n * ld, 0.0);
    for (int64_t l = 0; l < R; ++l)
        for (int64_t i = 0; i < n; ++i) hb[(size_t)i * ld + l] = X[l * ldx + i];
    double* buf[4];
    for (auto& b : buf) b = dnew<double>(ctx, (size_t)n * ld);
    try {
        REACH_CUDA(cudaMemcpyAsync(buf[0], hb.data(), sizeof(double) * n * ld, cudaMemcpyHostToDevice, st));
        S->last_products = 0; S->last_bytes = 0;
        REACH_CUDA(cudaEventRecord(ctx->ev0, st));
        double* res = expm_action(S, transpose_op ? S->At : S->Ar, delta, ld, buf[0], buf[1], buf[2], buf[3]);
        REACH_CUDA(cudaEventRecord(ctx->ev1, st));
        REACH_CUDA(cudaMemcpyAsync(hb.data(), res, sizeof(double) * n * ld, cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        REACH_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        S->last_ms = ms;
    } catch (...) {
        cudaStreamSynchronize(st);
        for (auto& b : buf) ctx_free(ctx, b);
        throw;
    }
    for (auto& b : buf) ctx_free(ctx, b);
    for (int64_t l = 0; l < R; ++l)
        for (int64_t i = 0; i < n; ++i) Y[l * ldy + i] = hb[(size_t)i * ld + l];
}

// discretize_interpolation (forward / backward) + box decomposition of Ω0 with A sparse and Φ never formed.
void sparse_discretize_box(SparseSys* S, double delta, int32_t algorithm, const double* c, const double* r, int64_t m,
                           const double* B, int64_t ldb, const double* cU, const double* rU, double* c0, double* r0,
                           double* rE, double* rpsi, double* MU) {
    Ctx* ctx = S->ctx;
    const int64_t n = S->n;
    if (algorithm != REACH_FORWARD && algorithm != REACH_BACKWARD) {
        if (algorithm == REACH_FIRSTORDER || algorithm == REACH_NOBLOATING)
            throw Error(REACH_EUNSUPPORTED, "reach_discretize_box_sparse: only the forward / backward interpolation models run on the sparse path");
        throw Error(REACH_EINVAL, "the algorithm is unknown");
    }
    REACH_REQUIRE(c && r && m >= 0 && (m == 0 || (cU && rU)), "reach_discretize_box_sparse: bad X0 / U");
    REACH_REQUIRE(!B || ldb >= n, "reach_discretize_box_sparse: bad ldb");
    REACH_REQUIRE(B || m == 0 || m == n, "reach_discretize_box_sparse: B == NULL means identity and needs m == n");
    cudaStream_t st = ctx->stream;
    const bool backward = algorithm == REACH_BACKWARD;
    // columns of Φ (and A²) that matter for |Φ| r, |A²| r: those with r_j != 0 (the others contribute exact zeros)
    std::vector<int> J;
    for (int64_t j = 0; j < n; ++j) {
        REACH_REQUIRE(r[j] >= 0.0 && std::isfinite(r[j]) && std::isfinite(c[j]), "reach_discretize_box_sparse: bad X0");
        if (r[j] != 0.0) J.push_back((int)j);
    }
    const int64_t CH = std::min<int64_t>(256, std::max<int64_t>(4, (int64_t)J.size()));
    const int64_t ldJ = block_ld(CH), ldm = block_ld(std::max<int64_t>(m, 1));
    std::vector<void*> tmp;
    auto alloc = [&](size_t count) { double* p = dnew<double>(ctx, count); tmp.push_back(p); return p; };
    auto free_all = [&]() { for (void* p : tmp) ctx_free(ctx, p); tmp.clear(); };
    try {
        // n-vectors live in n x 4 row-blocks (column 0 used): the same SpMM kernel serves as SpMV
        const size_t nv = (size_t)n * 4;
        double* vc = alloc(nv); double* va = alloc(nv); double* vb = alloc(nv); double* vt0 = alloc(nv); double* vt1 = alloc(nv);
        double* vw = alloc(nv); double* vw2 = alloc(nv);
        double* pc = alloc(n); double* pr = alloc(n); double* a1 = alloc(n); double* b1 = alloc(n); double* s1 = alloc(nv);
        double* dE = alloc(nv); double* dr = alloc(n); double* dc = alloc(n);
        double* uc = alloc(n); double* ur = alloc(n); double* a2 = alloc(n); double* b2 = alloc(n); double* s2 = alloc(nv);
        double* dpsi = alloc(nv); double* dc0 = alloc(n); double* dr0 = alloc(n);
        double* ZJ[4];
        for (auto& z : ZJ) z = alloc((size_t)n * ldJ);
        int* dsel = (int*)alloc((size_t)(CH + 1) / 2 + 1);
        double* dwt = alloc(CH);
        std::vector<double> hv(nv, 0.0);
        for (int64_t i = 0; i < n; ++i) hv[(size_t)i * 4] = c[i];
        REACH_CUDA(cudaMemcpyAsync(vc, hv.data(), sizeof(double) * nv, cudaMemcpyHostToDevice, st));
        REACH_CUDA(cudaMemcpyAsync(dc, c, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        REACH_CUDA(cudaMemcpyAsync(dr, r, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        REACH_CUDA(cudaStreamSynchronize(st));
        const unsigned gw = (unsigned)ceil_div(n * 32, 256), gv = (unsigned)ceil_div(n, 256);
        auto col0 = [&](const double* blk, double* out) {       // column 0 of an n x 4 block -> dense vector
            REACH_CUDA(cudaMemcpy2DAsync(out, 8, blk, 32, 8, n, cudaMemcpyDeviceToDevice, st));
        };
        // Φ c   (centre of ΦX0, discretize.jl:679,685)
        REACH_CUDA(cudaMemcpyAsync(vw, vc, sizeof(double) * nv, cudaMemcpyDeviceToDevice, st));
        double* phic = expm_action(S, S->Ar, delta, 4, vw, vw2, vt0, vt1);
        col0(phic, pc);
        // centre part of (A²)X0 (forward, :655) or (A²Φ)X0 (backward, :657)
        spmm<false>(S, S->Ar, 4, backward ? phic : vc, va, nullptr, nullptr, 1.0);
        spmm<false>(S, S->Ar, 4, va, vb, nullptr, nullptr, 1.0);
        col0(vb, a1);
        // radius parts: |Φ| r and |A²| r (|A²Φ| r), column chunk by column chunk of the identity
        REACH_CUDA(cudaMemsetAsync(pr, 0, sizeof(double) * n, st));
        REACH_CUDA(cudaMemsetAsync(b1, 0, sizeof(double) * n, st));
        for (size_t j0 = 0; j0 < J.size(); j0 += (size_t)CH) {
            const int cnt = (int)std::min<size_t>((size_t)CH, J.size() - j0);
            REACH_CUDA(cudaMemcpyAsync(dsel, J.data() + j0, sizeof(int) * cnt, cudaMemcpyHostToDevice, st));
            gather_kernel<<<(unsigned)ceil_div(cnt, 256), 256, 0, st>>>(cnt, dsel, dr, dwt);
            unit_block_kernel<<<(unsigned)ceil_div(n * ldJ, 256), 256, 0, st>>>((int)n, (int)ldJ, cnt, dsel, ZJ[0]);
            double* E = ZJ[0];
            double* PhiJ = nullptr;
            if (backward) {         // Φ E_J first, then A², sharing the block with |Φ| r
                PhiJ = expm_action(S, S->Ar, delta, ldJ, ZJ[0], ZJ[1], ZJ[2], ZJ[3]);
                row_box_kernel<<<gw, 256, 0, st>>>((int)n, (int)ldJ, cnt, PhiJ, nullptr, dwt, nullptr, pr, 1);
                double* other = PhiJ == ZJ[0] ? ZJ[1] : ZJ[0];
                spmm<false>(S, S->Ar, ldJ, PhiJ, ZJ[2], nullptr, nullptr, 1.0);
                spmm<false>(S, S->Ar, ldJ, ZJ[2], other, nullptr, nullptr, 1.0);
                row_box_kernel<<<gw, 256, 0, st>>>((int)n, (int)ldJ, cnt, other, nullptr, dwt, nullptr, b1, 1);
            } else {
                spmm<false>(S, S->Ar, ldJ, E, ZJ[2], nullptr, nullptr, 1.0);
                spmm<false>(S, S->Ar, ldJ, ZJ[2], ZJ[3], nullptr, nullptr, 1.0);
                row_box_kernel<<<gw, 256, 0, st>>>((int)n, (int)ldJ, cnt, ZJ[3], nullptr, dwt, nullptr, b1, 1);
                PhiJ = expm_action(S, S->Ar, delta, ldJ, ZJ[0], ZJ[1], ZJ[2], ZJ[3]);
                row_box_kernel<<<gw, 256, 0, st>>>((int)n, (int)ldJ, cnt, PhiJ, nullptr, dwt, nullptr, pr, 1);
            }
            REACH_CUDA(cudaGetLastError());
            ctx->launches += 4;
            REACH_CUDA(cudaStreamSynchronize(st));       // J chunk upload buffer is reused
        }
        // Einit radius = Φ₂(|A|, δ) sih((A²)X0)
        REACH_CUDA(cudaMemsetAsync(s1, 0, sizeof(double) * nv, st));
        sih_vec_kernel<<<gv, 256, 0, st>>>((int)n, a1, b1, a2);                   // a2 used as scratch
        REACH_CUDA(cudaMemcpy2DAsync(s1, 32, a2, 8, 8, n, cudaMemcpyDeviceToDevice, st));
        phi2abs_action(S, delta, s1, dE, vt0, vt1);
        double* rEv = alloc(n);
        col0(dE, rEv);
        double* rpsiv = alloc(n);
        if (m > 0) {
            // B in row-block layout (identity when B == NULL), A B, box of (A B) U and of δ B U
            std::vector<double> hB((size_t)n * ldm, 0.0);
            for (int64_t i = 0; i < n; ++i)
                for (int64_t l = 0; l < m; ++l) hB[(size_t)i * ldm + l] = B ? B[l * ldb + i] : (i == l ? 1.0 : 0.0);
            double* dB = alloc((size_t)n * ldm);
            double* dAB = alloc((size_t)n * ldm);
            double* dcu = alloc(m); double* dru = alloc(m);
            REACH_CUDA(cudaMemcpyAsync(dB, hB.data(), sizeof(double) * n * ldm, cudaMemcpyHostToDevice, st));
            REACH_CUDA(cudaMemcpyAsync(dcu, cU, sizeof(double) * m, cudaMemcpyHostToDevice, st));
            REACH_CUDA(cudaMemcpyAsync(dru, rU, sizeof(double) * m, cudaMemcpyHostToDevice, st));
            spmm<false>(S, S->Ar, ldm, dB, dAB, nullptr, nullptr, 1.0);            // A * U0 = (A B) U     :671
            row_box_kernel<<<gw, 256, 0, st>>>((int)n, (int)ldm, (int)m, dAB, dcu, dru, a2, b2, 0);
            REACH_CUDA(cudaMemsetAsync(s2, 0, sizeof(double) * nv, st));
            sih_vec_kernel<<<gv, 256, 0, st>>>((int)n, a2, b2, uc);               // uc as scratch
            REACH_CUDA(cudaMemcpy2DAsync(s2, 32, uc, 8, 8, n, cudaMemcpyDeviceToDevice, st));
            phi2abs_action(S, delta, s2, dpsi, vt0, vt1);                          // Eψ0 radius
            col0(dpsi, rpsiv);
            row_box_kernel<<<gw, 256, 0, st>>>((int)n, (int)ldm, (int)m, dB, dcu, dru, a2, b2, 0);   // box of B U
            scale_vec_kernel<<<gv, 256, 0, st>>>((int)n, delta, a2, uc);          // δ U0
            scale_vec_kernel<<<gv, 256, 0, st>>>((int)n, std::fabs(delta), b2, ur);
            REACH_CUDA(cudaGetLastError());
            ctx->launches += 5;
            REACH_CUDA(cudaStreamSynchronize(st));       // hB dies here
        }
        omega0_box(ctx, n, dc, dr, pc, pr, rEv, m > 0 ? uc : nullptr, ur, rpsiv, dc0, dr0);
        auto down = [&](double* dst, const double* src) {
            if (dst) REACH_CUDA(cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyDeviceToHost, st));
        };
        down(c0, dc0); down(r0, dr0); down(rE, rEv);
        if (m > 0) down(rpsi, rpsiv);
        REACH_CUDA(cudaStreamSynchronize(st));
        if (m > 0 && MU)
            for (int64_t l = 0; l < m; ++l)
                for (int64_t i = 0; i < n; ++i) MU[l * n + i] = delta * (B ? B[l * ldb + i] : (i == l ? 1.0 : 0.0));
    } catch (...) {
        cudaStreamSynchronize(st);
        free_all();
        throw;
    }
    free_all();
}

// reach_blocks! for a SparseMatrixExp (lazy rows), box blocks: see the header of this file.
// One panel of rows of interest (hrows[0 .. R)): the whole time loop for these rows; results go to the host arrays
// centers / radii (leading dimension ldo = all rows of the call) at row offset `row0`.
// V0 (optional, host n x R column-major): functional mode -- the block starts from the columns of V0 instead of unit vectors
// (homogeneous only; column 0 of the outputs is V0ᵀc0, |V0|ᵀr0).  cWs_out / rWs_out (optional): the input boxes per step.
static void lazy_boxes_panel(SparseSys* S, double delta, const double* c0, const double* r0, int64_t m, const double* MU,
                             int64_t ldmu, const double* cU, const double* rU, const double* rpsi, const int* hrows_ptr,
                             int64_t R, int64_t N, double* centers, double* radii, int64_t ldo, int64_t row0,
                             const double* V0 = nullptr, double* cWs_out = nullptr, double* rWs_out = nullptr) {
    REACH_REQUIRE(!V0 || m == 0, "lazy functionals are homogeneous");
    Ctx* ctx = S->ctx;
    const int64_t n = S->n;
    const int64_t ld = block_ld(R);
    cudaStream_t st = ctx->stream;
    struct { const int* p; const int* data() const { return p; } } hrows{hrows_ptr};
    std::vector<void*> tmp;
    auto alloc = [&](size_t count) { double* p = dnew<double>(ctx, count); tmp.push_back(p); return p; };
    auto free_all = [&]() { for (void* p : tmp) ctx_free(ctx, p); tmp.clear(); };
    try {
        double* W[4];
        for (auto& w : W) w = alloc((size_t)n * ld);
        double* dC = alloc((size_t)R * N); double* dR = alloc((size_t)R * N);
        double* dc0 = alloc(n); double* dr0 = alloc(n); double* dpsi = alloc(n);
        double* dMU = alloc((size_t)n * std::max<int64_t>(m, 1)); double* dcu = alloc(std::max<int64_t>(m, 1));
        double* dru = alloc(std::max<int64_t>(m, 1));
        double* cW = alloc(R); double* rW = alloc(R);
        double* dcWs = cWs_out ? alloc((size_t)R * N) : nullptr;
        double* drWs = cWs_out ? alloc((size_t)R * N) : nullptr;
        // epilogue geometry: ~2000 CTAs over (row chunks, state slices, quantity groups)
        const unsigned gx = (unsigned)ceil_div(R, 32);
        const int gz = 1 + (int)ceil_div(m, STEP_MCHUNK);
        const int nsl = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(64, ceil_div(n, 64)),
                                                                    std::max<int64_t>(8, 2048 / ((int64_t)gx * gz))));
        const int64_t ldp = round_up(R, 32);
        double* part = alloc((size_t)(3 + m) * nsl * ldp);
        int* drows = (int*)alloc((size_t)(R + 1) / 2 + 1);
        REACH_CUDA(cudaMemcpyAsync(dc0, c0, n * 8, cudaMemcpyHostToDevice, st));
        REACH_CUDA(cudaMemcpyAsync(dr0, r0, n * 8, cudaMemcpyHostToDevice, st));
        if (!V0) REACH_CUDA(cudaMemcpyAsync(drows, hrows.data(), sizeof(int) * R, cudaMemcpyHostToDevice, st));
        if (m > 0) {
            REACH_CUDA(cudaMemcpyAsync(dpsi, rpsi, n * 8, cudaMemcpyHostToDevice, st));
            REACH_CUDA(cudaMemcpy2DAsync(dMU, n * 8, MU, ldmu * 8, n * 8, m, cudaMemcpyHostToDevice, st));
            REACH_CUDA(cudaMemcpyAsync(dcu, cU, m * 8, cudaMemcpyHostToDevice, st));
            REACH_CUDA(cudaMemcpyAsync(dru, rU, m * 8, cudaMemcpyHostToDevice, st));
        }
        StepGeom geom;
        geom.n = (int)n; geom.ld = (int)ld; geom.R = (int)R; geom.nsl = nsl; geom.m = (int)m; geom.ldp = (int)ldp;
        geom.gx = (int)gx; geom.gz = gz;
        geom.c0 = dc0; geom.r0 = dr0; geom.rpsi = m > 0 ? dpsi : nullptr; geom.MU = dMU; geom.cU = dcu; geom.rU = dru;
        geom.cW = cW; geom.rW = rW; geom.part = part;
        geom.cbase = dC; geom.cWs = dcWs; geom.rWs = drWs;
        if (dcWs) {
            REACH_CUDA(cudaMemsetAsync(dcWs, 0, sizeof(double) * R * N, st));
            REACH_CUDA(cudaMemsetAsync(drWs, 0, sizeof(double) * R * N, st));
        }
        REACH_CUDA(cudaEventRecord(ctx->ev0, st));
        int* kzero = (int*)alloc(1);
        REACH_CUDA(cudaMemsetAsync(kzero, 0, sizeof(int), st));
        if (V0) {
            std::vector<double> hb((size_t)n * ld, 0.0);
            for (int64_t l = 0; l < R; ++l)
                for (int64_t i = 0; i < n; ++i) hb[(size_t)i * ld + l] = V0[l * n + i];
            REACH_CUDA(cudaMemcpyAsync(W[0], hb.data(), sizeof(double) * n * ld, cudaMemcpyHostToDevice, st));
            REACH_CUDA(cudaStreamSynchronize(st));
            // column 0: the functionals of the initial block itself
            lazy_partial_kernel<<<dim3(gx, (unsigned)nsl, (unsigned)gz), 32 * STEP_SUB, 0, st>>>(geom, W[0]);
            lazy_finish_kernel<<<gx, 32 * STEP_SUB, 0, st>>>(geom, dC, dR, kzero, 0);
        } else {
            lazy_init_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, st>>>((int)n, (int)R, drows, dc0, dr0, dpsi, (int)m, dMU, dcu,
                                                                         dru, cW, rW, dC, dR);
            unit_block_kernel<<<(unsigned)ceil_div(n * ld, 256), 256, 0, st>>>((int)n, (int)ld, (int)R, drows, W[0]);
        }
        REACH_CUDA(cudaGetLastError());
        ctx->launches += 2;
        int64_t steps = N - 1;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t gexec = nullptr;
        double* cur = W[0];
        int* kctr = (int*)alloc(1);
        const int one = 1;
        REACH_CUDA(cudaMemcpyAsync(kctr, &one, sizeof(int), cudaMemcpyHostToDevice, st));     // next output column (0-based)
        auto one_step = [&](int koff) {
            // rows of e^{(k-1)δA}: one more factor e^{δAᵀ} on the transposed block     (:306, :37-38)
            double* a = cur;
            double* b = nullptr;
            double* t0 = nullptr;
            double* t1 = nullptr;
            for (auto& w : W) {
                if (w == a) continue;
                if (!b) b = w; else if (!t0) t0 = w; else t1 = w;
            }
            cur = expm_action(S, S->At, delta, ld, a, b, t0, t1);
            lazy_partial_kernel<<<dim3(gx, (unsigned)nsl, (unsigned)gz), 32 * STEP_SUB, 0, st>>>(geom, cur);
            lazy_finish_kernel<<<gx, 32 * STEP_SUB, 0, st>>>(geom, dC, dR, kctr, koff);
            REACH_CUDA(cudaGetLastError());
            ctx->launches += 2;
        };
        // The step is a chain of two dozen short kernels: two consecutive steps -- after which the buffer rotation repeats
        // -- are captured once into a CUDA graph and replayed; the output column comes from a device counter.  (A single
        // persistent cooperative kernel with grid barriers between the Taylor terms was measured SLOWER, 5.8 k against
        // 7.0 k steps/s on C4 with 31 rows: a barrier over ~600 CTAs costs more than a kernel boundary inside a graph.)
        if (steps >= 6 && !getenv("REACH_LAZY_NO_GRAPH")) {
            double* cur0 = cur;
            const uint64_t l0 = ctx->launches;
            const int64_t p0 = S->last_products;
            const double b0 = S->last_bytes;
            REACH_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
            cudaError_t cap = cudaSuccess;
            try {
                one_step(0);
                one_step(1);
                bump_kernel<<<1, 1, 0, st>>>(kctr, 2);
            } catch (...) {
                cap = cudaErrorUnknown;
            }
            const cudaError_t endc = cudaStreamEndCapture(st, &graph);
            if (cap != cudaSuccess || endc != cudaSuccess || cur != cur0) {
                if (graph) cudaGraphDestroy(graph);
                throw Error(2, "reach_bffpsv18_boxes_lazy: CUDA graph capture of the step failed");
            }
            const cudaError_t inst = cudaGraphInstantiate(&gexec, graph, nullptr, nullptr, 0);
            if (inst != cudaSuccess) {
                cudaGraphDestroy(graph);
                throw Error(2, std::string("reach_bffpsv18_boxes_lazy: cudaGraphInstantiate: ") + cudaGetErrorString(inst));
            }
            const int64_t pairs = steps / 2;
            const uint64_t per_l = ctx->launches - l0 + 1;
            const int64_t per_p = S->last_products - p0;
            const double per_b = S->last_bytes - b0;
            ctx->launches = l0; S->last_products = p0; S->last_bytes = b0;
            for (int64_t q = 0; q < pairs; ++q) {
                const cudaError_t le = cudaGraphLaunch(gexec, st);
                if (le != cudaSuccess) {
                    cudaGraphExecDestroy(gexec); cudaGraphDestroy(graph);
                    throw Error(2, std::string("reach_bffpsv18_boxes_lazy: cudaGraphLaunch: ") + cudaGetErrorString(le));
                }
            }
            ctx->launches += per_l * pairs; S->last_products += per_p * pairs; S->last_bytes += per_b * pairs;
            steps -= 2 * pairs;
        }
        for (int64_t q = 0; q < steps; ++q) {
            one_step(0);
            bump_kernel<<<1, 1, 0, st>>>(kctr, 1);
            ctx->launches++;
        }
        REACH_CUDA(cudaGetLastError());
        REACH_CUDA(cudaEventRecord(ctx->ev1, st));
        REACH_CUDA(cudaMemcpy2DAsync(centers + row0, ldo * 8, dC, R * 8, R * 8, N, cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaMemcpy2DAsync(radii + row0, ldo * 8, dR, R * 8, R * 8, N, cudaMemcpyDeviceToHost, st));
        if (cWs_out) {
            REACH_CUDA(cudaMemcpy2DAsync(cWs_out + row0, ldo * 8, dcWs, R * 8, R * 8, N, cudaMemcpyDeviceToHost, st));
            REACH_CUDA(cudaMemcpy2DAsync(rWs_out + row0, ldo * 8, drWs, R * 8, R * 8, N, cudaMemcpyDeviceToHost, st));
        }
        REACH_CUDA(cudaStreamSynchronize(st));
        if (gexec) cudaGraphExecDestroy(gexec);
        if (graph) cudaGraphDestroy(graph);
        float ms = 0;
        REACH_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        S->last_ms += ms;
    } catch (...) {
        cudaStreamSynchronize(st);
        free_all();
        throw;
    }
    free_all();
}

// reach_blocks! for a SparseMatrixExp (lazy rows), box blocks: see the header of this file.  The rows of interest are
// independent of each other, so they are processed in PANELS whose working set (three n x panel blocks of the Horner
// recurrence) stays in the 126 MB L2: every gathered block row is then re-read from L2 instead of HBM by the ~nnz/n
// matrix rows that reference it (all 10913 rows of C4 at once: 39 steps/s, HBM-bound; in panels: see profiles/).
void sparse_bffpsv18_boxes(SparseSys* S, double delta, const double* c0, const double* r0, int64_t m, const double* MU,
                           int64_t ldmu, const double* cU, const double* rU, const double* rpsi, int64_t nrows,
                           const int32_t* rows, int64_t N, double* centers, double* radii, double* cWs = nullptr,
                           double* rWs = nullptr) {
    const int64_t n = S->n;
    REACH_REQUIRE(c0 && r0 && N >= 1 && centers && radii, "reach_bffpsv18_boxes_lazy: bad arguments");
    if (m == 0 || !MU) { m = 0; MU = nullptr; }
    REACH_REQUIRE(m == 0 || (cU && rU && rpsi && ldmu >= n), "reach_bffpsv18_boxes_lazy: inputs need MU, cU, rU, rpsi");
    std::vector<int> hrows;
    if (rows) {
        REACH_REQUIRE(nrows >= 1, "reach_bffpsv18_boxes_lazy: nrows must be positive");
        for (int64_t l = 0; l < nrows; ++l) {
            REACH_REQUIRE(rows[l] >= 0 && rows[l] < n, "reach_bffpsv18_boxes_lazy: row index out of range");
            hrows.push_back(rows[l]);
        }
    } else {
        nrows = n;
        for (int64_t l = 0; l < n; ++l) hrows.push_back((int)l);
    }
    // panel width: three blocks of n x panel doubles within ~64 MB (half of L2: measured best on C4), a multiple of 32
    int64_t panel = std::max<int64_t>(32, (((int64_t)64 << 20) / (3 * 8 * n)) / 32 * 32);
    if (const char* e = getenv("REACH_LAZY_PANEL")) panel = std::max<int64_t>(4, atoll(e));
    const int64_t npan = ceil_div(nrows, panel);
    const int64_t width = round_up(ceil_div(nrows, npan), 4);        // equal panels
    S->last_ms = 0; S->last_products = 0; S->last_bytes = 0;
    for (int64_t r0p = 0; r0p < nrows; r0p += width)
        lazy_boxes_panel(S, delta, c0, r0, m, MU, ldmu, cU, rU, rpsi, hrows.data() + r0p, std::min(width, nrows - r0p), N, centers,
                         radii, nrows, r0p, nullptr, cWs, rWs);
}

namespace {
// out[a, k] for k >= 1 (0-based column): centre = Σ_i sc[(a,i), k] + Σ_l M[a,l] cW[l,k], radius = Σ_i sr[(a,i), k] + Σ_l |M[a,l]| rW[l,k];
// column 0: M c0[rows], |M| r0[rows] (the boxed X̂0, reach_blocks.jl:152-155 / check_blocks.jl:122-129).  Thread per (a, k).
__global__ void lazy_map_combine_kernel(int q, int nb, int R, int N, const double* __restrict__ sc, const double* __restrict__ sr,
                                        const double* __restrict__ cWs, const double* __restrict__ rWs, const double* __restrict__ M,
                                        const double* __restrict__ c0r, const double* __restrict__ r0r, double* __restrict__ outc,
                                        double* __restrict__ outr) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= q * N) return;
    const int a = t % q, k = t / q;
    double c = 0.0, r = 0.0;
    if (k == 0) {
        for (int l = 0; l < R; ++l) {
            c = fma(M[(int64_t)l * q + a], c0r[l], c);
            r = fma(fabs(M[(int64_t)l * q + a]), r0r[l], r);
        }
    } else {
        const int F = q * nb;
        for (int i = 0; i < nb; ++i) {
            c += sc[(int64_t)k * F + a * nb + i];
            r += sr[(int64_t)k * F + a * nb + i];
        }
        if (cWs)
            for (int l = 0; l < R; ++l) {
                c = fma(M[(int64_t)l * q + a], cWs[(int64_t)k * R + l], c);
                r = fma(fabs(M[(int64_t)l * q + a]), rWs[(int64_t)k * R + l], r);
            }
    }
    outc[(int64_t)k * q + a] = c;
    outr[(int64_t)k * q + a] = r;
}
}  // namespace

// output_function / check-mode states with a lazy exponential: the blocks stay lazy (reach_blocks.jl:266-283 with
// output_function, check_blocks.jl), so per block i and output row a the functional v = Σ_{l in b_i} M[a,l] e^{kδA}[rows[l], :]
// enters as v c0 + |v| r0.  The functionals are just more columns for the same linear recurrence (block of q x nb
// combinations of unit vectors instead of unit vectors); the input boxes Ŵ_k come out of the ordinary row run.
void sparse_output_map(SparseSys* S, double delta, const double* c0, const double* r0, int64_t m, const double* MU, int64_t ldmu,
                       const double* cU, const double* rU, const double* rpsi, int64_t nrows, const int32_t* rows, int64_t N,
                       int64_t nblocks, const int32_t* block_sizes, int64_t q, const double* Mout, int64_t ldm, double* centers,
                       double* radii) {
    Ctx* ctx = S->ctx;
    const int64_t n = S->n, R = nrows;
    REACH_REQUIRE(rows && R >= 1 && q >= 1 && Mout && ldm >= q && centers && radii && N >= 1, "lazy output map: bad arguments");
    if (m == 0 || !MU) { m = 0; MU = nullptr; }
    std::vector<int64_t> boff(1, 0);
    if (block_sizes) {
        for (int64_t b = 0; b < nblocks; ++b) { REACH_REQUIRE(block_sizes[b] >= 1, "lazy output map: empty block"); boff.push_back(boff.back() + block_sizes[b]); }
        REACH_REQUIRE(boff.back() == R, "lazy output map: block sizes must add up to nrows");
    } else {
        for (int64_t l = 1; l <= R; ++l) boff.push_back(l);      // every row is its own block
    }
    const int64_t nb = (int64_t)boff.size() - 1, F = q * nb;
    for (int64_t l = 0; l < R; ++l) REACH_REQUIRE(rows[l] >= 0 && rows[l] < n, "lazy output map: row index out of range");
    double total_ms = 0;
    int64_t total_products = 0;
    double total_bytes = 0;
    // (1) input boxes Ŵ_k per row (only with inputs)
    std::vector<double> tc, tr, cWs, rWs;
    if (m > 0) {
        tc.resize((size_t)R * N); tr.resize((size_t)R * N); cWs.resize((size_t)R * N); rWs.resize((size_t)R * N);
        sparse_bffpsv18_boxes(S, delta, c0, r0, m, MU, ldmu, cU, rU, rpsi, R, rows, N, tc.data(), tr.data(), cWs.data(), rWs.data());
        total_ms += S->last_ms; total_products += S->last_products; total_bytes += S->last_bytes;
    }
    // (2) the functionals: column (a, i) starts as Σ_{l in b_i} M[a,l] e_{rows[l]}
    std::vector<double> sc((size_t)F * N), sr((size_t)F * N);
    {
        int64_t panel = std::max<int64_t>(32, (((int64_t)64 << 20) / (3 * 8 * n)) / 32 * 32);
        const int64_t npan = ceil_div(F, panel), width = round_up(ceil_div(F, npan), 4);
        for (int64_t f0 = 0; f0 < F; f0 += width) {
            const int64_t fc = std::min(width, F - f0);
            std::vector<double> V0((size_t)n * fc, 0.0);
            for (int64_t f = 0; f < fc; ++f) {
                const int64_t a = (f0 + f) / nb, i = (f0 + f) % nb;
                for (int64_t l = boff[i]; l < boff[i + 1]; ++l) V0[(size_t)f * n + rows[l]] += Mout[l * ldm + a];
            }
            S->last_ms = 0; S->last_products = 0; S->last_bytes = 0;
            lazy_boxes_panel(S, delta, c0, r0, 0, nullptr, 0, nullptr, nullptr, nullptr, nullptr, fc, N, sc.data(), sr.data(), F, f0,
                             V0.data());
            total_ms += S->last_ms; total_products += S->last_products; total_bytes += S->last_bytes;
        }
    }
    // (3) combine on the device
    cudaStream_t st = ctx->stream;
    std::vector<void*> tmp;
    auto up = [&](const double* h, size_t count) {
        double* d = dnew<double>(ctx, count);
        tmp.push_back(d);
        REACH_CUDA(cudaMemcpyAsync(d, h, sizeof(double) * count, cudaMemcpyHostToDevice, st));
        return d;
    };
    try {
        std::vector<double> hM((size_t)q * R), c0r(R), r0r(R);
        for (int64_t l = 0; l < R; ++l) {
            c0r[l] = c0[rows[l]]; r0r[l] = r0[rows[l]];
            for (int64_t a = 0; a < q; ++a) hM[(size_t)l * q + a] = Mout[l * ldm + a];
        }
        double* dsc = up(sc.data(), sc.size()); double* dsr = up(sr.data(), sr.size());
        double* dcw = m > 0 ? up(cWs.data(), cWs.size()) : nullptr; double* drw = m > 0 ? up(rWs.data(), rWs.size()) : nullptr;
        double* dM = up(hM.data(), hM.size()); double* dc0r = up(c0r.data(), R); double* dr0r = up(r0r.data(), R);
        double* dout = dnew<double>(ctx, (size_t)2 * q * N);
        tmp.push_back(dout);
        lazy_map_combine_kernel<<<(unsigned)ceil_div(q * N, 128), 128, 0, st>>>((int)q, (int)nb, (int)R, (int)N, dsc, dsr, dcw, drw, dM,
                                                                               dc0r, dr0r, dout, dout + q * N);
        REACH_CUDA(cudaGetLastError());
        ctx->launches++;
        REACH_CUDA(cudaMemcpyAsync(centers, dout, sizeof(double) * q * N, cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaMemcpyAsync(radii, dout + q * N, sizeof(double) * q * N, cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        cudaStreamSynchronize(st);
        for (void* p : tmp) ctx_free(ctx, p);
        throw;
    }
    for (void* p : tmp) ctx_free(ctx, p);
    S->last_ms = total_ms; S->last_products = total_products; S->last_bytes = total_bytes;
}

}  // namespace reach

// =========================================================================================================
// C ABI (include/reach.h)
// =========================================================================================================
namespace reach { Ctx* ctx_of(reach_ctx* h); }
using namespace reach;

struct reach_sparse {
    SparseSys* s;
    Ctx* ctx;
};

#define SP_BEGIN(ctx)                                      \
    if (!(ctx)) return REACH_EINVAL;                       \
    try {                                                  \
        REACH_CUDA(cudaSetDevice((ctx)->device));
#define SP_END(ctx)                                                                               \
    }                                                                                             \
    catch (const reach::Error& e) { (ctx)->last_error = e.what(); return e.code; }                \
    catch (const std::exception& e) { (ctx)->last_error = e.what(); return REACH_ECUDA; } \
    catch (...) { (ctx)->last_error = "unknown exception"; return REACH_ECUDA; }

extern "C" {

int32_t reach_sparse_create(reach_ctx* h, int64_t n, int64_t nnz, const int64_t* colptr, const int64_t* rowval,
                            const double* nzval, reach_sparse** out) {
    if (!h || !out) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    *out = nullptr;
    SP_BEGIN(ctx)
    SparseSys* s = sparse_create(ctx, n, nnz, colptr, rowval, nzval);
    *out = new reach_sparse{s, ctx};
    return REACH_OK;
    SP_END(ctx)
}

void reach_sparse_free(reach_sparse* sp) {
    if (!sp) return;
    cudaSetDevice(sp->ctx->device);
    cudaStreamSynchronize(sp->ctx->stream);
    sparse_free(sp->s);
    delete sp;
}

int32_t reach_expmv(reach_sparse* sp, double delta, int32_t transpose_op, int64_t R, const double* X, int64_t ldx, double* Y,
                    int64_t ldy) {
    if (!sp) return REACH_EINVAL;
    SP_BEGIN(sp->ctx)
    sparse_expmv(sp->s, delta, transpose_op, R, X, ldx, Y, ldy);
    return REACH_OK;
    SP_END(sp->ctx)
}

int32_t reach_discretize_box_sparse(reach_sparse* sp, double delta, int32_t algorithm, const double* c, const double* r,
                                    int64_t m, const double* B, int64_t ldb, const double* cU, const double* rU, double* c0,
                                    double* r0, double* rE, double* rpsi, double* MU) {
    if (!sp) return REACH_EINVAL;
    SP_BEGIN(sp->ctx)
    sparse_discretize_box(sp->s, delta, algorithm, c, r, m, B, ldb, cU, rU, c0, r0, rE, rpsi, MU);
    return REACH_OK;
    SP_END(sp->ctx)
}

int32_t reach_bffpsv18_boxes_lazy(reach_sparse* sp, double delta, const double* c0, const double* r0, int64_t m,
                                  const double* MU, int64_t ldmu, const double* cU, const double* rU, const double* rpsi,
                                  int64_t nrows, const int32_t* rows, int64_t N, double* centers, double* radii,
                                  int64_t* steps_done) {
    if (!sp) return REACH_EINVAL;
    SP_BEGIN(sp->ctx)
    sparse_bffpsv18_boxes(sp->s, delta, c0, r0, m, MU, ldmu, cU, rU, rpsi, nrows, rows, N, centers, radii);
    if (steps_done) *steps_done = N;
    return REACH_OK;
    SP_END(sp->ctx)
}

int32_t reach_bffpsv18_output_map_lazy(reach_sparse* sp, double delta, const double* c0, const double* r0, int64_t m,
                                       const double* MU, int64_t ldmu, const double* cU, const double* rU, const double* rpsi,
                                       int64_t nrows, const int32_t* rows, int64_t N, int64_t nblocks, const int32_t* block_sizes,
                                       int64_t q, const double* Mout, int64_t ldm, double* centers, double* radii) {
    if (!sp) return REACH_EINVAL;
    SP_BEGIN(sp->ctx)
    sparse_output_map(sp->s, delta, c0, r0, m, MU, ldmu, cU, rU, rpsi, nrows, rows, N, nblocks, block_sizes, q, Mout, ldm, centers,
                      radii);
    return REACH_OK;
    SP_END(sp->ctx)
}

int32_t reach_bffpsv18_check_lazy(reach_sparse* sp, double delta, const double* c0, const double* r0, int64_t m, const double* MU,
                                  int64_t ldmu, const double* cU, const double* rU, const double* rpsi, int64_t nrows,
                                  const int32_t* rows, int64_t N, int64_t nblocks, const int32_t* block_sizes, const double* ell,
                                  double bound, int32_t eager, int64_t* violation) {
    if (!sp) return REACH_EINVAL;
    SP_BEGIN(sp->ctx)
    REACH_REQUIRE(ell && violation && N >= 1, "reach_bffpsv18_check_lazy: ell and violation are required");
    std::vector<double> c(N), r(N);
    sparse_output_map(sp->s, delta, c0, r0, m, MU, ldmu, cU, rU, rpsi, nrows, rows, N, nblocks, block_sizes, 1, ell, 1, c.data(),
                      r.data());
    *violation = 0;
    for (int64_t k = 0; k < N; ++k)
        if (c[k] + r[k] > bound) { *violation = k + 1; break; }     // first violating step (check_blocks.jl:116); eager or not, the same index
    (void)eager;
    return REACH_OK;
    SP_END(sp->ctx)
}

int32_t reach_sparse_timing(reach_sparse* sp, double* run_ms, int64_t* products, double* bytes) {
    if (!sp) return REACH_EINVAL;
    if (run_ms) *run_ms = sp->s->last_ms;
    if (products) *products = sp->s->last_products;
    if (bytes) *bytes = sp->s->last_bytes;
    return REACH_OK;
}

}  // extern "C"
