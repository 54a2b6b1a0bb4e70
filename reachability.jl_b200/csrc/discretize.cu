// Discretisation on the device: src/ReachSets/discretize.jl of the reference.
//
//   exp_Aδ(A, δ)          = exp(Matrix(A*δ))                         discretize.jl:150-152
//   Φ₁(A, δ), Φ₂(A, δ)    = blocks of exp(Pmatrix(A, δ))              discretize.jl:162-166, 221-226, 296-301
//   discretize_interpolation (forward / backward) for box sets        discretize.jl:627-675
//   box decomposition of the lazy Ω0 (BFFPSV18/reach.jl:64-65)
//
// B200 design: everything is GEMM.  The reference forms the dense 3n x 3n exponential to read off Φ₁/Φ₂ (27x the
// flops of e^{Aδ}); here the three blocks come out of ONE Horner chain on the n x n matrix X = A δ / 2^s,
//     R_{m+1} = I,  R_j = I + (X / j) R_{j+1}      =>   e^X = R_1,  Σ X^i/(i+1)! = R_2,  Σ X^i/(i+2)! = R_3 / 2,
// one DMMA GEMM per term with the scale 1/j and the "+ I" fused into the epilogue, followed by s doubling steps
//     Φ₀(2t) = Φ₀(t)²,   Φ₁(2t) = (Φ₀(t) + I) Φ₁(t),   Φ₂(2t) = (Φ₀(t) + I) Φ₂(t) + t Φ₁(t)
// (the block rows of exp(Pmatrix)², read off the block-triangular structure).  ‖X‖₁ <= 1/2 and m = 18 terms put the
// truncation error below 1e-20, so the result agrees with Julia's Padé-13 `exp` to rounding (tests pin both against
// a 50-digit mpmath reference).
#include "common.cuh"
#include "../../include/reach.h"
#include <cmath>
#include <algorithm>

namespace reach {

namespace {

constexpr int TAYLOR_TERMS = 18;

__global__ void scale_copy_kernel(int n, const double* __restrict__ A, int64_t lda, double alpha, int take_abs,
                                  double* __restrict__ X, int64_t ldx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i < n) {
        double v = A[(int64_t)j * lda + i];
        if (take_abs) v = fabs(v);
        X[(int64_t)j * ldx + i] = alpha * v;
    }
}

// out[j] = sum_i |A[i, j]|  (column sums for the 1-norm); one warp per column
__global__ void col_abs_sums_kernel(int n, const double* __restrict__ A, int64_t lda, double* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += fabs(A[(int64_t)warp * lda + i]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[warp] = s;
}

// Y = a * X + b * Z + d * I   (n x n, elementwise; Z may be NULL)
__global__ void axpby_diag_kernel(int n, double a, const double* __restrict__ X, int64_t ldx, double b,
                                  const double* __restrict__ Z, int64_t ldz, double d, double* __restrict__ Y, int64_t ldy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i < n) {
        double v = a * X[(int64_t)j * ldx + i];
        if (Z) v += b * Z[(int64_t)j * ldz + i];
        if (i == j) v += d;
        Y[(int64_t)j * ldy + i] = v;
    }
}

// Row-wise products with a box (c, r):  sgn[i] = sum_j M[i,j] c[j],  mag[i] = sum_j |M[i,j]| r[j].
// CTA = 64 rows x 4 column groups; columns in ascending order inside a group, groups combined in order.
__global__ void __launch_bounds__(256) box_matvec_kernel(int rows, int cols, const double* __restrict__ M, int64_t ldm,
                                                         const double* __restrict__ c, const double* __restrict__ r,
                                                         double* __restrict__ sgn, double* __restrict__ mag) {
    __shared__ double ss[4][64], sm[4][64];
    const int li = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int i = blockIdx.x * 64 + li;
    double a = 0.0, b = 0.0;
    if (i < rows) {
        for (int j = grp; j < cols; j += 4) {
            const double v = M[(int64_t)j * ldm + i];
            if (c) a = fma(v, c[j], a);
            if (r) b = fma(fabs(v), r[j], b);
        }
    }
    ss[grp][li] = a;
    sm[grp][li] = b;
    __syncthreads();
    if (grp == 0 && i < rows) {
        if (sgn) sgn[i] = (ss[0][li] + ss[1][li]) + (ss[2][li] + ss[3][li]);
        if (mag) mag[i] = (sm[0][li] + sm[1][li]) + (sm[2][li] + sm[3][li]);
    }
}

// s[i] = |a[i]| + b[i]      (symmetric_interval_hull radius of a box with centre a, radius b)
__global__ void sih_kernel(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) s[i] = fabs(a[i]) + b[i];
}

// box decomposition of Ω0 = CH(X0, Y), Y = ΦX0 ⊕ δU0 ⊕ Eψ0 ⊕ Einit   (discretize.jl:679,685)
__global__ void omega0_box_kernel(int n, const double* __restrict__ c, const double* __restrict__ r,
                                  const double* __restrict__ pc, const double* __restrict__ pr,
                                  const double* __restrict__ rE, const double* __restrict__ uc,
                                  const double* __restrict__ ur, const double* __restrict__ rpsi,
                                  double* __restrict__ c0, double* __restrict__ r0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double cY = pc[i];
    double rY = pr[i] + rE[i];
    if (uc) {
        cY = cY + uc[i];
        rY = rY + ur[i] + rpsi[i];
    }
    const double hi = fmax(c[i] + r[i], cY + rY);
    const double lo = fmin(c[i] - r[i], cY - rY);
    c0[i] = (hi + lo) / 2;
    r0[i] = (hi - lo) / 2;
}


// ---- set_operations = "zonotope" (discretize.jl:705-743): concrete zonotope arithmetic on the device ------------------------
// Candidate generators of Z2, one warp per column, with a "has a non-zero entry" flag per column (the `Zonotope`
// constructor's zero-generator removal):
//   inhomogeneous (:718-722)  [ Φ diag(r) | δ (B diag(rU)) | diag(rψ) | diag(rE) ]      3n + m columns
//   homogeneous   (:708)      [ Φ diag(r) | Φ diag(rE) ]                                2n columns   (Φ also maps Einit)
__global__ void __launch_bounds__(256) zono_cand_kernel(int n, int m, int homog, const double* __restrict__ Phi, int64_t ldphi,
                                                        const double* __restrict__ r, const double* __restrict__ rE,
                                                        const double* __restrict__ B, int64_t ldb, const double* __restrict__ rU,
                                                        double delta, const double* __restrict__ rpsi, double* __restrict__ C,
                                                        int64_t ldc, int* __restrict__ flags) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int ncand = homog ? 2 * n : 3 * n + m;
    if (j >= ncand) return;
    int any = 0;
    for (int i = lane; i < n; i += 32) {
        double v;
        if (j < n) v = Phi[(int64_t)j * ldphi + i] * r[j];                               // linear_map(ϕ, Z1)
        else if (homog) v = Phi[(int64_t)(j - n) * ldphi + i] * rE[j - n];               // ϕ * Einit
        else if (j < n + m) v = delta * (B[(int64_t)(j - n) * ldb + i] * rU[j - n]);     // linear_map(δI, convert(Zonotope, B*U))
        else if (j < 2 * n + m) v = (i == j - n - m) ? rpsi[i] : 0.0;                    // Eψ0
        else v = (i == j - 2 * n - m) ? rE[i] : 0.0;                                     // Einit
        C[(int64_t)j * ldc + i] = v;
        any |= (v != 0.0);
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) flags[j] = any;
}

// overapproximate(ConvexHull(ZA, ZB), Zonotope) with ngens(ZA) = qA >= qB = ngens(ZB), one warp per column:
//   [ (GA[:, 1:qB] + GB)/2 | (cA - cB)/2 | (GA[:, 1:qB] - GB)/2 | GA[:, qB+1:end] ],   centre (cA + cB)/2.
// An operand is either Z1 = convert(Zonotope, X0) (type 0: column j = r[j] e_j) or Z2 (type 1: candidate column j of C).
struct ZonoOperand {
    int type, q;
    const int* idx;         // the generators that exist (zero-generator removal already applied)
    const double* c;        // centre
};
__device__ __forceinline__ double zono_entry(const ZonoOperand& z, int col, int i, const double* __restrict__ r,
                                             const double* __restrict__ C, int64_t ldc) {
    const int j = z.idx[col];
    return z.type == 0 ? (i == j ? r[j] : 0.0) : C[(int64_t)j * ldc + i];
}
__global__ void __launch_bounds__(256) zono_ch_kernel(int n, ZonoOperand za, ZonoOperand zb, const double* __restrict__ r,
                                                      const double* __restrict__ C, int64_t ldc, double* __restrict__ out,
                                                      int64_t ldo, int* __restrict__ flags, double* __restrict__ c0) {
    const int oc = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int qB = zb.q, total = za.q + zb.q + 1;
    if (oc >= total) return;
    int any = 0;
    for (int i = lane; i < n; i += 32) {
        double v;
        if (oc < qB) v = (zono_entry(za, oc, i, r, C, ldc) + zono_entry(zb, oc, i, r, C, ldc)) / 2;
        else if (oc == qB) {
            v = (za.c[i] - zb.c[i]) / 2;
            c0[i] = (za.c[i] + zb.c[i]) / 2;
        } else if (oc <= 2 * qB) v = (zono_entry(za, oc - qB - 1, i, r, C, ldc) - zono_entry(zb, oc - qB - 1, i, r, C, ldc)) / 2;
        else v = zono_entry(za, oc - qB - 1, i, r, C, ldc);
        out[(int64_t)oc * ldo + i] = v;
        any |= (v != 0.0);
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) flags[oc] = any;
}

// dst[:, k] = src[:, idx[k]]
__global__ void __launch_bounds__(256) gather_cols_kernel(int n, int count, const int* __restrict__ idx, const double* __restrict__ src,
                                                          int64_t lds, double* __restrict__ dst, int64_t ldd) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= count) return;
    const int j = idx[k];
    for (int i = lane; i < n; i += 32) dst[(int64_t)k * ldd + i] = src[(int64_t)j * lds + i];
}

// y = a + alpha * b
__global__ void axpy_vec_kernel(int n, const double* __restrict__ a, double alpha, const double* __restrict__ b, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = (a ? a[i] : 0.0) + alpha * b[i];
}


// Y = |X1| + X2 elementwise (n x K, leading dimension ld): the inner symmetric interval hull of K boxes at once
__global__ void abs_add_kernel(int n, const double* __restrict__ X1, const double* __restrict__ X2, int64_t ld, double* __restrict__ Y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (i < n) Y[(int64_t)k * ld + i] = fabs(X1[(int64_t)k * ld + i]) + X2[(int64_t)k * ld + i];
}

void launch2d(Ctx* ctx, int n, dim3& grid) { grid = dim3((unsigned)ceil_div(n, 256), (unsigned)n); (void)ctx; }

}  // namespace

// box decomposition of Ω0 on device vectors (shared with the sparse path, sparse.cu)
void omega0_box(Ctx* ctx, int64_t n, const double* c, const double* r, const double* pc, const double* pr, const double* rE,
                const double* uc, const double* ur, const double* rpsi, double* c0, double* r0) {
    omega0_box_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>((int)n, c, r, pc, pr, rE, uc, ur, rpsi, c0, r0);
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
}

// Horner chain + doubling.  dA: n x n on the device (ld lda).  Outputs are DMats allocated by the caller (any may
// be NULL): P0 = e^{Aδ}, P1 = Φ₁(A, δ), P2 = Φ₂(A, δ).  take_abs: use |A| entrywise (Φ₂(abs.(A), δ), discretize.jl:652).
void expm_family(Ctx* ctx, int64_t n, const double* dA, int64_t lda, double delta, bool take_abs, DMat* P0, DMat* P1,
                 DMat* P2) {
    cudaStream_t st = ctx->stream;
    const bool need12 = (P1 != nullptr) || (P2 != nullptr);
    // 1-norm of A δ
    double* dsum = nullptr;
    REACH_CUDA(cudaMalloc(&dsum, sizeof(double) * n));
    col_abs_sums_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>((int)n, dA, lda, dsum);
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
    std::vector<double> hs(n);
    REACH_CUDA(cudaMemcpyAsync(hs.data(), dsum, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaStreamSynchronize(st));
    cudaFree(dsum);
    double nrm = 0.0;
    for (double v : hs) {
        if (!std::isfinite(v)) throw Error(REACH_EINVAL, "expm: matrix has non-finite entries");
        nrm = std::max(nrm, v);
    }
    nrm *= std::fabs(delta);
    int s = 0;
    if (nrm > 0.5) s = (int)std::ceil(std::log2(nrm / 0.5));
    if (s > 60) throw Error(REACH_EINVAL, "expm: ||A delta|| too large");
    const double h = delta / std::ldexp(1.0, s);

    DMat X = dmat_alloc(ctx, n, n), Ra = dmat_alloc(ctx, n, n), Rb = dmat_alloc(ctx, n, n);
    DMat R2 = need12 ? dmat_alloc(ctx, n, n) : DMat(), R3 = need12 ? dmat_alloc(ctx, n, n) : DMat();
    dim3 g2;
    launch2d(ctx, (int)n, g2);
    scale_copy_kernel<<<g2, 256, 0, st>>>((int)n, dA, lda, h, take_abs ? 1 : 0, X.p, X.ld);
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
    // R_{m+1} = I
    DMat* cur = &Ra;
    DMat* nxt = &Rb;
    axpby_diag_kernel<<<g2, 256, 0, st>>>((int)n, 0.0, X.p, X.ld, 0.0, nullptr, 0, 1.0, cur->p, cur->ld);
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
    for (int j = TAYLOR_TERMS; j >= 1; --j) {
        GemmEpilogue ep;
        ep.alpha = 1.0 / j;
        ep.diag = 1.0;
        DMat* dst = nxt;
        if (need12 && j == 3) dst = &R3;
        if (need12 && j == 2) dst = &R2;
        gemm_left(ctx, n, n, n, X.p, X.ld, cur->p, cur->ld, dst->p, dst->ld, ep);      // R_j = I + (X/j) R_{j+1}
        if (dst == nxt) std::swap(cur, nxt); else cur = dst;
    }
    // cur = R_1 = e^X ; R2 = Σ X^i/(i+1)! ; R3/2 = Σ X^i/(i+2)!
    DMat W = need12 ? dmat_alloc(ctx, n, n) : DMat();
    DMat Q1, Q2, T1, T2;
    if (need12) {
        Q1 = dmat_alloc(ctx, n, n); Q2 = dmat_alloc(ctx, n, n); T1 = dmat_alloc(ctx, n, n); T2 = dmat_alloc(ctx, n, n);
        axpby_diag_kernel<<<g2, 256, 0, st>>>((int)n, h, R2.p, R2.ld, 0.0, nullptr, 0, 0.0, Q1.p, Q1.ld);
        axpby_diag_kernel<<<g2, 256, 0, st>>>((int)n, 0.5 * h * h, R3.p, R3.ld, 0.0, nullptr, 0, 0.0, Q2.p, Q2.ld);
        REACH_CUDA(cudaGetLastError());
        ctx->launches += 2;
    }
    // the three blocks live in: e = *cur (other of Ra/Rb is scratch), Q1, Q2
    DMat* e = cur;
    DMat* escr = (cur == &Ra) ? &Rb : &Ra;
    if (cur == &R2 || cur == &R3) throw Error(REACH_ECUDA, "expm: internal buffer rotation error");
    DMat *q1 = &Q1, *q2 = &Q2, *t1 = &T1, *t2 = &T2;
    double t = h;
    for (int i = 0; i < s; ++i) {
        if (need12) {
            axpby_diag_kernel<<<g2, 256, 0, st>>>((int)n, 1.0, e->p, e->ld, 0.0, nullptr, 0, 1.0, W.p, W.ld);   // W = Φ₀ + I
            REACH_CUDA(cudaGetLastError());
            ctx->launches++;
            gemm_left(ctx, n, n, n, W.p, W.ld, q1->p, q1->ld, t1->p, t1->ld);             // Φ₁(2t) = W Φ₁
            gemm_left(ctx, n, n, n, W.p, W.ld, q2->p, q2->ld, t2->p, t2->ld);             // W Φ₂
            axpby_diag_kernel<<<g2, 256, 0, st>>>((int)n, 1.0, t2->p, t2->ld, t, q1->p, q1->ld, 0.0, t2->p, t2->ld);  // + t Φ₁(t)
            REACH_CUDA(cudaGetLastError());
            ctx->launches++;
            std::swap(q1, t1);
            std::swap(q2, t2);
        }
        gemm_left(ctx, n, n, n, e->p, e->ld, e->p, e->ld, escr->p, escr->ld);             // Φ₀(2t) = Φ₀²
        std::swap(e, escr);
        t *= 2;
    }
    auto give = [&](DMat* out, DMat* src) {
        if (!out) return;
        REACH_CUDA(cudaMemcpy2DAsync(out->p, out->ld * 8, src->p, src->ld * 8, n * 8, n, cudaMemcpyDeviceToDevice, st));
    };
    give(P0, e);
    if (need12) { give(P1, q1); give(P2, q2); }
    REACH_CUDA(cudaStreamSynchronize(st));
    dmat_free(X); dmat_free(Ra); dmat_free(Rb);
    if (need12) { dmat_free(R2); dmat_free(R3); dmat_free(W); dmat_free(Q1); dmat_free(Q2); dmat_free(T1); dmat_free(T2); }
}

static void box_matvec(Ctx* ctx, int64_t rows, int64_t cols, const double* M, int64_t ldm, const double* c, const double* r,
                       double* sgn, double* mag) {
    if (rows <= 0) return;
    box_matvec_kernel<<<(unsigned)ceil_div(rows, 64), 256, 0, ctx->stream>>>((int)rows, (int)cols, M, ldm, c, r, sgn, mag);
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
}

}  // namespace reach

// =========================================================================================================
namespace reach { Ctx* ctx_of(reach_ctx* h); }
using namespace reach;

#define D_BEGIN(ctx)                                      \
    if (!(ctx)) return REACH_EINVAL;                      \
    try {                                                 \
        REACH_CUDA(cudaSetDevice((ctx)->device));
#define D_END(ctx)                                                                                \
    }                                                                                             \
    catch (const reach::Error& e) { (ctx)->last_error = e.what(); return e.code; }                \
    catch (const std::exception& e) { (ctx)->last_error = e.what(); return REACH_ECUDA; } \
    catch (...) { (ctx)->last_error = "unknown exception"; return REACH_ECUDA; }

// discretize_firstorder (p = Inf, discretize.jl:403-473) and discretize_nobloating (:527-552) for box sets, in the same
// array form as the interpolation models.
static int32_t discretize_box_other(Ctx* ctx, int64_t n, const double* A, int64_t lda, double delta, int32_t algorithm,
                                    const double* c, const double* r, int64_t m, const double* B, int64_t ldb, const double* cU,
                                    const double* rU, double* Phi, int64_t ldphi, double* c0, double* r0, double* rE,
                                    double* rpsi, double* MU) {
    cudaStream_t st = ctx->stream;
    DMat dA = dmat_alloc(ctx, n, n), dPhi = dmat_alloc(ctx, n, n), dB, dMU, dP1;
    const int64_t NV = 13;      // 0 c, 1 r, 2 Φc, 3 |Φ|r, 4 ones -> α, 5 |A|1 -> β, 6 |c|+r, 7 |Bc_U|+|B|r_U, 8 δBc_U, 9 |δB|r_U, 10 c0, 11 r0, 12 zeros
    double* vec = (double*)ctx_malloc(ctx, sizeof(double) * (NV * n + 2 * std::max<int64_t>(m, 1)));
    REACH_CUDA(cudaMemsetAsync(vec, 0, sizeof(double) * (NV * n + 2 * std::max<int64_t>(m, 1)), st));
    auto V = [&](int i) { return vec + (int64_t)i * n; };
    double* dcU = vec + NV * n;
    double* drU = dcU + std::max<int64_t>(m, 1);
    dmat_upload(ctx, dA, A, lda);
    REACH_CUDA(cudaMemcpyAsync(V(0), c, n * 8, cudaMemcpyHostToDevice, st));
    REACH_CUDA(cudaMemcpyAsync(V(1), r, n * 8, cudaMemcpyHostToDevice, st));
    expm_family(ctx, n, dA.p, dA.ld, delta, false, &dPhi, nullptr, nullptr);
    const unsigned gv = (unsigned)ceil_div(n, 256);
    std::vector<double> hn(n), hx(n), hu(n);
    if (m > 0) {
        REACH_CUDA(cudaMemcpyAsync(dcU, cU, m * 8, cudaMemcpyHostToDevice, st));
        REACH_CUDA(cudaMemcpyAsync(drU, rU, m * 8, cudaMemcpyHostToDevice, st));
        dB = dmat_alloc(ctx, n, m);
        if (B) {
            dmat_upload(ctx, dB, B, ldb);
        } else {
            dim3 g2((unsigned)ceil_div(n, 256), (unsigned)n);
            axpby_diag_kernel<<<g2, 256, 0, st>>>((int)n, 0.0, dA.p, dA.ld, 0.0, nullptr, 0, 1.0, dB.p, dB.ld);
        }
        dMU = dmat_alloc(ctx, n, m);
    }
    if (algorithm == REACH_NOBLOATING) {
        REACH_CUDA(cudaMemcpyAsync(V(10), c, n * 8, cudaMemcpyHostToDevice, st));       // Ω0 = X0
        REACH_CUDA(cudaMemcpyAsync(V(11), r, n * 8, cudaMemcpyHostToDevice, st));
        if (m > 0) {
            dP1 = dmat_alloc(ctx, n, n);
            expm_family(ctx, n, dA.p, dA.ld, delta, false, nullptr, &dP1, nullptr);
            gemm_left(ctx, n, m, n, dP1.p, dP1.ld, dB.p, dB.ld, dMU.p, dMU.ld);           // Φ₁(A, δ) * B
        }
        // rE = rpsi = 0 (vec is zero-filled): V(4) doubles as the zero vector
    } else {
        // ‖A‖_inf, ‖X0‖_inf, ‖B U‖_inf: row sums on the device, the three maxima on the host (scalars of the model)
        REACH_CUDA(cudaMemsetAsync(V(4), 0, n * 8, st));
        {
            std::vector<double> ones(n, 1.0);
            REACH_CUDA(cudaMemcpyAsync(V(4), ones.data(), n * 8, cudaMemcpyHostToDevice, st));
            REACH_CUDA(cudaStreamSynchronize(st));
        }
        box_matvec(ctx, n, n, dA.p, dA.ld, nullptr, V(4), nullptr, V(5));                 // |A| 1
        sih_kernel<<<gv, 256, 0, st>>>((int)n, V(0), V(1), V(6));                         // |c| + r
        if (m > 0) {
            box_matvec(ctx, n, m, dB.p, dB.ld, dcU, drU, V(8), V(9));                     // box of B U
            sih_kernel<<<gv, 256, 0, st>>>((int)n, V(8), V(9), V(7));
        }
        REACH_CUDA(cudaMemcpyAsync(hn.data(), V(5), n * 8, cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaMemcpyAsync(hx.data(), V(6), n * 8, cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaMemcpyAsync(hu.data(), V(7), n * 8, cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaStreamSynchronize(st));
        const double Anorm = *std::max_element(hn.begin(), hn.end());
        const double RX0 = *std::max_element(hx.begin(), hx.end());
        const double RU = m > 0 ? *std::max_element(hu.begin(), hu.end()) : 0.0;
        const double kappa = std::exp(delta * Anorm) - 1.0 - delta * Anorm;
        const double alpha = m > 0 ? kappa * (RX0 + RU / Anorm) : kappa * RX0;
        const double beta = m > 0 ? kappa * RU / Anorm : 0.0;
        std::vector<double> av(n, alpha), bv(n, beta);
        REACH_CUDA(cudaMemcpyAsync(V(4), av.data(), n * 8, cudaMemcpyHostToDevice, st));   // rE = α
        REACH_CUDA(cudaMemcpyAsync(V(5), bv.data(), n * 8, cudaMemcpyHostToDevice, st));   // rpsi = β
        REACH_CUDA(cudaStreamSynchronize(st));
        box_matvec(ctx, n, n, dPhi.p, dPhi.ld, V(0), V(1), V(2), V(3));                   // box of ΦX0
        if (m > 0) {
            dim3 gm((unsigned)ceil_div(n, 256), (unsigned)m);
            scale_copy_kernel<<<gm, 256, 0, st>>>((int)n, dB.p, dB.ld, delta, 0, dMU.p, dMU.ld);   // δ B
            box_matvec(ctx, n, m, dMU.p, dMU.ld, dcU, drU, V(8), V(9));
        }
        // Ω0 = CH(X0, ΦX0 ⊕ δU ⊕ □α): the kernel adds rE, and for inputs ur + rpsi; β must not enter Ω0 => pass zeros
        omega0_box_kernel<<<gv, 256, 0, st>>>((int)n, V(0), V(1), V(2), V(3), V(4), m > 0 ? V(8) : nullptr, V(9), V(12), V(10), V(11));
    }
    REACH_CUDA(cudaGetLastError());
    if (Phi) { REACH_REQUIRE(ldphi >= n, "reach_discretize_box: ldphi"); dmat_download(ctx, dPhi, Phi, ldphi); }
    auto down = [&](double* dst, const double* src) {
        if (dst) REACH_CUDA(cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyDeviceToHost, st));
    };
    down(c0, V(10)); down(r0, V(11));
    if (algorithm == REACH_NOBLOATING) {
        std::vector<double> z(n, 0.0);
        if (rE) std::memcpy(rE, z.data(), n * 8);
        if (rpsi && m > 0) std::memcpy(rpsi, z.data(), n * 8);
    } else {
        down(rE, V(4));
        if (m > 0) down(rpsi, V(5));
    }
    if (m > 0 && MU) dmat_download(ctx, dMU, MU, n);
    REACH_CUDA(cudaStreamSynchronize(st));
    dmat_free(dA); dmat_free(dPhi); dmat_free(dB); dmat_free(dMU); dmat_free(dP1);
    ctx_free(ctx, vec);
    return REACH_OK;
}


// Device results of discretize_interpolation (forward / backward) for a box X0 and a constant box input through B, kept on
// the device: the box entry point downloads them, the zonotope entry point keeps building on them.
struct BoxDev {
    Ctx* ctx = nullptr;
    int64_t n = 0, m = 0;
    DMat dA, dPhi, dP2, dM, dM2, dB, dMU;
    double* vec = nullptr;
    // n-vectors: 0 c, 1 r, 2 Φc, 3 |Φ|r, 4-6 inner sih of (A²)X0, 7 rE, 8 (δB)cU, 9 |δB|rU, 10-12 inner sih of (AB)U, 13 rψ,
    //            14 c0, 15 r0, 16 B cU, 17 Φc + δ(B cU) ; then the m-vectors cU, rU
    static constexpr int NV = 18;
    double* V(int i) const { return vec + (int64_t)i * n; }
    double* dcU() const { return vec + (int64_t)NV * n; }
    double* drU() const { return dcU() + std::max<int64_t>(m, 1); }
    ~BoxDev() {
        if (ctx) cudaStreamSynchronize(ctx->stream);      // nothing may still be using the blocks that go back to the pool
        dmat_free(dA); dmat_free(dPhi); dmat_free(dP2); dmat_free(dM); dmat_free(dM2); dmat_free(dB); dmat_free(dMU);
        if (vec) ctx_free(ctx, vec);
    }
};

static void interp_box_device(Ctx* ctx, int64_t n, const double* A, int64_t lda, double delta, int32_t algorithm, const double* c,
                              const double* r, int64_t m, const double* B, int64_t ldb, const double* cU, const double* rU,
                              BoxDev& d) {
    REACH_REQUIRE(m >= 0 && (m == 0 || (cU && rU)), "reach_discretize: inputs need cU, rU");
    REACH_REQUIRE(!B || ldb >= n, "reach_discretize: bad ldb");
    REACH_REQUIRE(B || m == 0 || m == n, "reach_discretize: B == NULL means identity and needs m == n");
    cudaStream_t st = ctx->stream;
    d.ctx = ctx; d.n = n; d.m = m;
    d.dA = dmat_alloc(ctx, n, n); d.dPhi = dmat_alloc(ctx, n, n); d.dP2 = dmat_alloc(ctx, n, n); d.dM = dmat_alloc(ctx, n, n);
    const size_t vbytes = sizeof(double) * (size_t)(BoxDev::NV * n + 2 * std::max<int64_t>(m, 1));
    d.vec = (double*)ctx_malloc(ctx, vbytes);
    REACH_CUDA(cudaMemsetAsync(d.vec, 0, vbytes, st));
    dmat_upload(ctx, d.dA, A, lda);
    REACH_CUDA(cudaMemcpyAsync(d.V(0), c, n * 8, cudaMemcpyHostToDevice, st));
    REACH_CUDA(cudaMemcpyAsync(d.V(1), r, n * 8, cudaMemcpyHostToDevice, st));
    expm_family(ctx, n, d.dA.p, d.dA.ld, delta, false, &d.dPhi, nullptr, nullptr);          // ϕ = exp_Aδ(A, δ)        :649
    expm_family(ctx, n, d.dA.p, d.dA.ld, delta, true, nullptr, nullptr, &d.dP2);            // Phi2Aabs = Φ₂(|A|, δ)   :652
    gemm_left(ctx, n, n, n, d.dA.p, d.dA.ld, d.dA.p, d.dA.ld, d.dM.p, d.dM.ld);              // A * A                   :655
    const DMat* Msel = &d.dM;
    if (algorithm == REACH_BACKWARD) {                                                       // A * A * ϕ               :657
        d.dM2 = dmat_alloc(ctx, n, n);
        gemm_left(ctx, n, n, n, d.dM.p, d.dM.ld, d.dPhi.p, d.dPhi.ld, d.dM2.p, d.dM2.ld);
        Msel = &d.dM2;
    }
    const unsigned gv = (unsigned)ceil_div(n, 256);
    box_matvec(ctx, n, n, Msel->p, Msel->ld, d.V(0), d.V(1), d.V(4), d.V(5));                // box of (A²)X0
    sih_kernel<<<gv, 256, 0, st>>>((int)n, d.V(4), d.V(5), d.V(6));                          // inner sih
    box_matvec(ctx, n, n, d.dP2.p, d.dP2.ld, nullptr, d.V(6), nullptr, d.V(7));              // Einit radius = Φ₂|A| s1
    box_matvec(ctx, n, n, d.dPhi.p, d.dPhi.ld, d.V(0), d.V(1), d.V(2), d.V(3));              // box of ΦX0
    ctx->launches += 1;
    if (m > 0) {
        REACH_CUDA(cudaMemcpyAsync(d.dcU(), cU, m * 8, cudaMemcpyHostToDevice, st));
        REACH_CUDA(cudaMemcpyAsync(d.drU(), rU, m * 8, cudaMemcpyHostToDevice, st));
        d.dB = dmat_alloc(ctx, n, m);
        if (B) {
            dmat_upload(ctx, d.dB, B, ldb);
        } else {
            dim3 g2((unsigned)ceil_div(n, 256), (unsigned)n);
            axpby_diag_kernel<<<g2, 256, 0, st>>>((int)n, 0.0, d.dA.p, d.dA.ld, 0.0, nullptr, 0, 1.0, d.dB.p, d.dB.ld);
            ctx->launches++;
        }
        d.dMU = dmat_alloc(ctx, n, m);
        gemm_left(ctx, n, m, n, d.dA.p, d.dA.ld, d.dB.p, d.dB.ld, d.dMU.p, d.dMU.ld);        // A * U0 = (A B) U        :671
        box_matvec(ctx, n, m, d.dMU.p, d.dMU.ld, d.dcU(), d.drU(), d.V(10), d.V(11));
        sih_kernel<<<gv, 256, 0, st>>>((int)n, d.V(10), d.V(11), d.V(12));
        box_matvec(ctx, n, n, d.dP2.p, d.dP2.ld, nullptr, d.V(12), nullptr, d.V(13));        // Eψ0 radius
        // δ U0: MU = δ B, box of MU * U
        dim3 gm((unsigned)ceil_div(n, 256), (unsigned)m);
        scale_copy_kernel<<<gm, 256, 0, st>>>((int)n, d.dB.p, d.dB.ld, delta, 0, d.dMU.p, d.dMU.ld);
        box_matvec(ctx, n, m, d.dMU.p, d.dMU.ld, d.dcU(), d.drU(), d.V(8), d.V(9));
        box_matvec(ctx, n, m, d.dB.p, d.dB.ld, d.dcU(), nullptr, d.V(16), nullptr);          // B cU (centre of convert(Zonotope, B*U))
        ctx->launches += 2;
    }
    omega0_box_kernel<<<gv, 256, 0, st>>>((int)n, d.V(0), d.V(1), d.V(2), d.V(3), d.V(7), m > 0 ? d.V(8) : nullptr, d.V(9), d.V(13),
                                          d.V(14), d.V(15));
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
}

extern "C" {

int32_t reach_expm(reach_ctx* h, int64_t n, const double* A, int64_t lda, double delta, double* Phi, int64_t ldphi) {
    if (!h) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    D_BEGIN(ctx)
    REACH_REQUIRE(n > 0 && A && Phi && lda >= n && ldphi >= n, "reach_expm: bad arguments");
    DMat dA = dmat_alloc(ctx, n, n), dP = dmat_alloc(ctx, n, n);
    dmat_upload(ctx, dA, A, lda);
    expm_family(ctx, n, dA.p, dA.ld, delta, false, &dP, nullptr, nullptr);
    dmat_download(ctx, dP, Phi, ldphi);
    REACH_CUDA(cudaStreamSynchronize(ctx->stream));
    dmat_free(dA); dmat_free(dP);
    return REACH_OK;
    D_END(ctx)
}

int32_t reach_phi12(reach_ctx* h, int64_t n, const double* A, int64_t lda, double delta, double* Phi1, int64_t ld1,
                    double* Phi2, int64_t ld2) {
    if (!h) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    D_BEGIN(ctx)
    REACH_REQUIRE(n > 0 && A && lda >= n && (!Phi1 || ld1 >= n) && (!Phi2 || ld2 >= n), "reach_phi12: bad arguments");
    DMat dA = dmat_alloc(ctx, n, n), d1 = dmat_alloc(ctx, n, n), d2 = dmat_alloc(ctx, n, n);
    dmat_upload(ctx, dA, A, lda);
    expm_family(ctx, n, dA.p, dA.ld, delta, false, nullptr, &d1, &d2);
    if (Phi1) dmat_download(ctx, d1, Phi1, ld1);
    if (Phi2) dmat_download(ctx, d2, Phi2, ld2);
    REACH_CUDA(cudaStreamSynchronize(ctx->stream));
    dmat_free(dA); dmat_free(d1); dmat_free(d2);
    return REACH_OK;
    D_END(ctx)
}

int32_t reach_discretize_box(reach_ctx* h, int64_t n, const double* A, int64_t lda, double delta, int32_t algorithm,
                             const double* c, const double* r, int64_t m, const double* B, int64_t ldb,
                             const double* cU, const double* rU, double* Phi, int64_t ldphi, double* Phi2abs,
                             int64_t ldphi2, double* c0, double* r0, double* rE, double* rpsi, double* MU) {
    if (!h) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    D_BEGIN(ctx)
    REACH_REQUIRE(n > 0 && A && lda >= n && c && r, "reach_discretize_box: bad A / X0");
    if (algorithm < REACH_FORWARD || algorithm > REACH_NOBLOATING)
        throw Error(REACH_EINVAL, "the algorithm is unknown");      // ArgumentError at discretize.jl:121,659
    if (algorithm == REACH_FIRSTORDER || algorithm == REACH_NOBLOATING)
        return discretize_box_other(ctx, n, A, lda, delta, algorithm, c, r, m, B, ldb, cU, rU, Phi, ldphi, c0, r0, rE, rpsi, MU);
    cudaStream_t st = ctx->stream;
    BoxDev d;
    interp_box_device(ctx, n, A, lda, delta, algorithm, c, r, m, B, ldb, cU, rU, d);
    if (Phi) { REACH_REQUIRE(ldphi >= n, "reach_discretize_box: ldphi"); dmat_download(ctx, d.dPhi, Phi, ldphi); }
    if (Phi2abs) { REACH_REQUIRE(ldphi2 >= n, "reach_discretize_box: ldphi2"); dmat_download(ctx, d.dP2, Phi2abs, ldphi2); }
    auto down = [&](double* dst, int i) {
        if (dst) REACH_CUDA(cudaMemcpyAsync(dst, d.V(i), n * 8, cudaMemcpyDeviceToHost, st));
    };
    down(c0, 14); down(r0, 15); down(rE, 7);
    if (m > 0) {
        down(rpsi, 13);
        if (MU) dmat_download(ctx, d.dMU, MU, n);
    }
    REACH_CUDA(cudaStreamSynchronize(st));
    return REACH_OK;
    D_END(ctx)
}


// discretize(...; set_operations = "zonotope") for a box X0 and a constant box input through B (discretize.jl:705-743,
// 882-890): Φ, Φ₂(|A|), Einit, Eψ0 as in reach_discretize_box, then -- on the device -- the concrete zonotope arithmetic
// (convert, linear_map, minkowski_sum, overapproximate(ConvexHull(Z1, Z2), Zonotope)) with the `Zonotope` constructor's
// zero-generator removal.  Only which generators EXIST is decided on the host, from the per-column flags the kernels emit.
int32_t reach_discretize_zonotope(reach_ctx* h, int64_t n, const double* A, int64_t lda, double delta, int32_t algorithm,
                                  const double* c, const double* r, int64_t m, const double* B, int64_t ldb, const double* cU,
                                  const double* rU, int32_t flags, double* Phi, int64_t ldphi, double* c0, double* G0, int64_t ldg0,
                                  int64_t* p0, double* cV, double* GV, int64_t ldgv, int64_t* pV) {
    if (!h) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    D_BEGIN(ctx)
    REACH_REQUIRE(n > 0 && A && lda >= n && c && r, "reach_discretize_zonotope: bad A / X0");
    REACH_REQUIRE(c0 && G0 && ldg0 >= n && p0, "reach_discretize_zonotope: c0, G0 (n x (4n + m + 1)) and p0 are required");
    if (algorithm == REACH_FIRSTORDER || algorithm == REACH_NOBLOATING)      // ArgumentError, discretize.jl:105-115
        throw Error(REACH_EINVAL, std::string("the algorithm ") + (algorithm == REACH_FIRSTORDER ? "firstorder" : "nobloating") +
                                      " with set operations=zonotope is not implemented");
    if (algorithm != REACH_FORWARD && algorithm != REACH_BACKWARD)
        throw Error(REACH_EINVAL, "the algorithm is unknown");               // discretize.jl:121
    if (m > 0) REACH_REQUIRE(cV && GV && ldgv >= n && pV, "reach_discretize_zonotope: inputs need cV, GV (n x (n + m)) and pV");
    const bool rzg = (flags & REACH_GLGM06_KEEP_ZERO_GENERATORS) == 0;
    cudaStream_t st = ctx->stream;
    BoxDev d;
    interp_box_device(ctx, n, A, lda, delta, algorithm, c, r, m, B, ldb, cU, rU, d);
    const int homog = m == 0 ? 1 : 0;
    const int64_t ncand = homog ? 2 * n : 3 * n + m, ld = round_up(n, 2), nout_max = n + ncand + 1;
    DevScratch scr(ctx);        // freed after the stream has drained (also on the error path)
    double* C = scr.get<double>((size_t)ncand * ld);
    double* Out = scr.get<double>((size_t)nout_max * ld);
    double* Fin = scr.get<double>((size_t)nout_max * ld);
    int* dflags = scr.get<int>(nout_max);
    int* didx = scr.get<int>(3 * nout_max);
    zono_cand_kernel<<<(unsigned)ceil_div(ncand * 32, 256), 256, 0, st>>>((int)n, (int)m, homog, d.dPhi.p, d.dPhi.ld, d.V(1), d.V(7),
                                                                         d.dB.p, d.dB.ld, d.drU(), delta, d.V(13), C, ld, dflags);
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
    std::vector<int> f2(ncand);
    REACH_CUDA(cudaMemcpyAsync(f2.data(), dflags, sizeof(int) * ncand, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaStreamSynchronize(st));
    std::vector<int> idx1, idx2, idxV;
    for (int64_t j = 0; j < n; ++j) if (!rzg || r[j] != 0.0) idx1.push_back((int)j);            // convert(Zonotope, X0)   :707 / :717
    for (int64_t j = 0; j < ncand; ++j) if (!rzg || f2[j]) idx2.push_back((int)j);
    if (!homog) for (int64_t j = n; j < 2 * n + m; ++j) if (!rzg || f2[j]) idxV.push_back((int)j);   // minkowski_sum(δU0, Eψ0)  :726
    // centre of Z2: Φ c (+ δ (B cU))
    const double* c2 = d.V(2);
    if (!homog) {
        axpy_vec_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>((int)n, d.V(2), delta, d.V(16), d.V(17));
        ctx->launches++;
        c2 = d.V(17);
    }
    const bool swap = idx2.size() > idx1.size();        // the operand with more generators goes first
    int* dA_idx = didx;
    int* dB_idx = didx + nout_max;
    const std::vector<int>& ia = swap ? idx2 : idx1;
    const std::vector<int>& ib = swap ? idx1 : idx2;
    if (!ia.empty()) REACH_CUDA(cudaMemcpyAsync(dA_idx, ia.data(), sizeof(int) * ia.size(), cudaMemcpyHostToDevice, st));
    if (!ib.empty()) REACH_CUDA(cudaMemcpyAsync(dB_idx, ib.data(), sizeof(int) * ib.size(), cudaMemcpyHostToDevice, st));
    ZonoOperand za{swap ? 1 : 0, (int)ia.size(), dA_idx, swap ? c2 : d.V(0)};
    ZonoOperand zb{swap ? 0 : 1, (int)ib.size(), dB_idx, swap ? d.V(0) : c2};
    const int64_t nout = (int64_t)ia.size() + (int64_t)ib.size() + 1;
    zono_ch_kernel<<<(unsigned)ceil_div(nout * 32, 256), 256, 0, st>>>((int)n, za, zb, d.V(1), C, ld, Out, ld, dflags, d.V(14));
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
    std::vector<int> fo(nout), idxO;
    REACH_CUDA(cudaMemcpyAsync(fo.data(), dflags, sizeof(int) * nout, cudaMemcpyDeviceToHost, st));
    REACH_CUDA(cudaStreamSynchronize(st));
    for (int64_t j = 0; j < nout; ++j) if (!rzg || fo[j]) idxO.push_back((int)j);
    int* dO_idx = didx + 2 * nout_max;
    auto gather_down = [&](const std::vector<int>& idx, const double* src, double* host, int64_t ldh) {
        if (idx.empty()) return;
        REACH_CUDA(cudaMemcpyAsync(dO_idx, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice, st));
        gather_cols_kernel<<<(unsigned)ceil_div((int64_t)idx.size() * 32, 256), 256, 0, st>>>((int)n, (int)idx.size(), dO_idx, src, ld, Fin, ld);
        REACH_CUDA(cudaGetLastError());
        ctx->launches++;
        REACH_CUDA(cudaMemcpy2DAsync(host, ldh * 8, Fin, ld * 8, n * 8, idx.size(), cudaMemcpyDeviceToHost, st));
        REACH_CUDA(cudaStreamSynchronize(st));        // dO_idx / Fin are reused by the next gather
    };
    gather_down(idxO, Out, G0, ldg0);
    *p0 = (int64_t)idxO.size();
    REACH_CUDA(cudaMemcpyAsync(c0, d.V(14), n * 8, cudaMemcpyDeviceToHost, st));
    if (!homog) {
        gather_down(idxV, C, GV, ldgv);
        *pV = (int64_t)idxV.size();
        axpy_vec_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>((int)n, nullptr, delta, d.V(16), d.V(17));     // δ (B cU)
        ctx->launches++;
        REACH_CUDA(cudaMemcpyAsync(cV, d.V(17), n * 8, cudaMemcpyDeviceToHost, st));
    } else if (pV) {
        *pV = 0;
    }
    if (Phi) { REACH_REQUIRE(ldphi >= n, "reach_discretize_zonotope: ldphi"); dmat_download(ctx, d.dPhi, Phi, ldphi); }
    REACH_CUDA(cudaStreamSynchronize(st));
    return REACH_OK;
    D_END(ctx)
}


// Time-varying inputs (VaryingInput branches of discretize.jl:690-696 / 729-737): for K input boxes U_k = B * Box(cU_k, rU_k)
//     Eψ_k = sih(Φ₂(|A|, δ) * sih(A * U_k))      radius  rψ_k = Φ₂(|A|, δ) (|A B cU_k| + |A B| rU_k)
// for all k at once: three thin GEMMs (A B CU, |A B| RU, Φ₂ S) on the DMMA kernel instead of K mat-vec sweeps.
int32_t reach_discretize_inputs_box(reach_ctx* h, int64_t n, const double* A, int64_t lda, double delta, int64_t m, const double* B,
                                    int64_t ldb, int64_t K, const double* cUs, int64_t ldcu, const double* rUs, int64_t ldru,
                                    double* rpsi, int64_t ldrpsi, double* MU) {
    if (!h) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    D_BEGIN(ctx)
    REACH_REQUIRE(n > 0 && A && lda >= n && m > 0 && K > 0 && cUs && rUs && ldcu >= m && ldru >= m && rpsi && ldrpsi >= n,
                  "reach_discretize_inputs_box: bad arguments");
    REACH_REQUIRE(!B || ldb >= n, "reach_discretize_inputs_box: bad ldb");
    REACH_REQUIRE(B || m == n, "reach_discretize_inputs_box: B == NULL means identity and needs m == n");
    cudaStream_t st = ctx->stream;
    DMatGuard gA(dmat_alloc(ctx, n, n)), gP2(dmat_alloc(ctx, n, n)), gB(dmat_alloc(ctx, n, m)), gAB(dmat_alloc(ctx, n, m)),
        gABabs(dmat_alloc(ctx, n, m));
    DMat &dA = gA.m, &dP2 = gP2.m, &dB = gB.m, &dAB = gAB.m, &dABabs = gABabs.m;
    // generator layout for the thin operands: columns contiguous, ld a multiple of 4, >= 128 zeroed columns of slack
    const int64_t lm = round_up(m, 4), ln = round_up(n, 4);
    DevScratch scr(ctx);
    double* dCU = scr.get<double>((size_t)(K + 128) * lm);
    double* dRU = scr.get<double>((size_t)(K + 128) * lm);
    double* dX1 = scr.get<double>((size_t)(K + 128) * ln);
    double* dX2 = scr.get<double>((size_t)(K + 128) * ln);
    double* dS = scr.get<double>((size_t)(K + 128) * ln);
    double* dR = scr.get<double>((size_t)(K + 128) * ln);
    for (double* p : {dCU, dRU}) REACH_CUDA(cudaMemsetAsync(p, 0, (size_t)(K + 128) * lm * 8, st));
    for (double* p : {dX1, dX2, dS, dR}) REACH_CUDA(cudaMemsetAsync(p, 0, (size_t)(K + 128) * ln * 8, st));
    dmat_upload(ctx, dA, A, lda);
    if (B) {
        dmat_upload(ctx, dB, B, ldb);
    } else {
        dim3 g2((unsigned)ceil_div(n, 256), (unsigned)n);
        axpby_diag_kernel<<<g2, 256, 0, st>>>((int)n, 0.0, dA.p, dA.ld, 0.0, nullptr, 0, 1.0, dB.p, dB.ld);
        ctx->launches++;
    }
    REACH_CUDA(cudaMemcpy2DAsync(dCU, lm * 8, cUs, ldcu * 8, m * 8, K, cudaMemcpyHostToDevice, st));
    REACH_CUDA(cudaMemcpy2DAsync(dRU, lm * 8, rUs, ldru * 8, m * 8, K, cudaMemcpyHostToDevice, st));
    expm_family(ctx, n, dA.p, dA.ld, delta, true, nullptr, nullptr, &dP2);             // Φ₂(|A|, δ)                 :652
    gemm_left(ctx, n, m, n, dA.p, dA.ld, dB.p, dB.ld, dAB.p, dAB.ld);                   // A * B
    dim3 gm((unsigned)ceil_div(n, 256), (unsigned)m);
    scale_copy_kernel<<<gm, 256, 0, st>>>((int)n, dAB.p, dAB.ld, 1.0, 1, dABabs.p, dABabs.ld);      // |A B|
    gemm_left(ctx, n, K, m, dAB.p, dAB.ld, dCU, lm, dX1, ln);                           // (A B) cU_k for every k
    gemm_left(ctx, n, K, m, dABabs.p, dABabs.ld, dRU, lm, dX2, ln);                     // |A B| rU_k
    abs_add_kernel<<<dim3((unsigned)ceil_div(n, 256), (unsigned)K), 256, 0, st>>>((int)n, dX1, dX2, ln, dS);   // sih(A * U_k)  :693,:731
    gemm_left(ctx, n, K, n, dP2.p, dP2.ld, dS, ln, dR, ln);                             // outer sih: Φ₂(|A|) s_k >= 0
    REACH_CUDA(cudaGetLastError());
    ctx->launches += 2;
    REACH_CUDA(cudaMemcpy2DAsync(rpsi, ldrpsi * 8, dR, ln * 8, n * 8, K, cudaMemcpyDeviceToHost, st));
    if (MU) {
        scale_copy_kernel<<<gm, 256, 0, st>>>((int)n, dB.p, dB.ld, delta, 0, dAB.p, dAB.ld);       // δ B
        ctx->launches++;
        dmat_download(ctx, dAB, MU, n);
    }
    REACH_CUDA(cudaStreamSynchronize(st));
    return REACH_OK;
    D_END(ctx)
}

}  // extern "C"
