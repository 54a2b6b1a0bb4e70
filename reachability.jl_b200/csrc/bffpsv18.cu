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

namespace reach {

struct BoxPlan {
    Ctx* ctx = nullptr;
    int64_t n = 0, m = 0, nrows = 0, N = 0, Rp = 0, s = 1;
    DMat Phi, PowA, PowB, PowT, BtAug, BtSmall, S[2], E, MU;
    double *c0 = nullptr, *r0 = nullptr, *rpsi = nullptr, *cU = nullptr, *rU = nullptr;
    double *bsum = nullptr, *fsum = nullptr, *cW = nullptr, *rW = nullptr;
    int32_t* rows = nullptr;
    double *centers = nullptr, *radii = nullptr;
    // timing of the last run
    double run_ms = 0, gemm_ms = 0, gemm_flops = 0;
    int64_t gemm_launches = 0;
    std::vector<cudaEvent_t> evs;
    // optional check-mode / output-map extras are handled by separate entry points
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

// Sequential scan over the cnt powers of one stack block, one thread per row of interest.
// E holds [P δB | P c0] for the block: E[(j*Rp + l) + q*ldE], q < m: (P δB)[:,q], q == m: P c0.
__global__ void scan_block_kernel(int nrows, int m, int64_t Rp, int cnt, int64_t first_power, int64_t N,
                                  const double* __restrict__ E, int64_t ldE, const double* __restrict__ bsum,
                                  const double* __restrict__ fsum, const double* __restrict__ cU,
                                  const double* __restrict__ rU, double* __restrict__ cW, double* __restrict__ rW,
                                  double* __restrict__ centers, double* __restrict__ radii) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nrows) return;
    double cw = cW[l], rw = rW[l];
    for (int j = 0; j < cnt; ++j) {
        const int64_t q = first_power + j;          // this is P_q, it yields reach set q+1 (1-based), column q
        const int64_t idx = (int64_t)j * Rp + l;
        const double a = E[idx + (int64_t)m * ldE];
        const double b = bsum[idx];
        centers[q * nrows + l] = a + cw;
        radii[q * nrows + l] = b + rw;
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

double* dvec_alloc(Ctx* ctx, int64_t len, const double* host) {
    double* p = nullptr;
    REACH_CUDA(cudaMalloc(&p, sizeof(double) * (size_t)(len > 0 ? len : 1)));
    if (host && len > 0)
        REACH_CUDA(cudaMemcpyAsync(p, host, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    else
        REACH_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * (size_t)(len > 0 ? len : 1), ctx->stream));
    return p;
}

int64_t choose_block(int64_t n, int64_t Rp, int64_t N, int64_t m) {
    // enough stacked rows for ~6 waves of 128x128 tiles on 148 SMs, capped by N-1, 1024 and ~24 GB of stacks
    const int64_t want_tiles = 148 * 6;
    const int64_t tiles_n = ceil_div(n + m + 1, 128);
    int64_t s = 1;
    while (s < 1024 && s < N - 1 && ceil_div(s * Rp, 128) * tiles_n < want_tiles) s *= 2;
    const double bytes_per_block = 2.0 * (double)Rp * (double)round_up(n, 128) * 8.0;
    while (s > 1 && bytes_per_block * (double)s > 24e9) s /= 2;
    return s;
}

}  // namespace

void boxplan_destroy(BoxPlan* P);

BoxPlan* boxplan_create(Ctx* ctx, int64_t n, const double* Phi, int64_t ldphi, const double* c0, const double* r0,
                        int64_t m, const double* MU, int64_t ldmu, const double* cU, const double* rU,
                        const double* rpsi, int64_t nrows, const int32_t* rows, int64_t N) {
    REACH_REQUIRE(n > 0 && Phi && c0 && r0 && ldphi >= n, "bffpsv18: bad Phi/c0/r0");
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
        P->s = N > 2 ? choose_block(n, P->Rp, N, m) : 1;
        const int64_t s = P->s;
        P->Phi = dmat_alloc(ctx, n, n);
        dmat_upload(ctx, P->Phi, Phi, ldphi);
        P->c0 = dvec_alloc(ctx, n, c0);
        P->r0 = dvec_alloc(ctx, n, r0);
        if (m > 0) {
            P->MU = dmat_alloc(ctx, n, m);
            dmat_upload(ctx, P->MU, MU, ldmu);
            P->cU = dvec_alloc(ctx, m, cU);
            P->rU = dvec_alloc(ctx, m, rU);
            P->rpsi = dvec_alloc(ctx, n, rpsi);
        }
        REACH_CUDA(cudaMalloc(&P->rows, sizeof(int32_t) * nrows));
        REACH_CUDA(cudaMemcpyAsync(P->rows, rows, sizeof(int32_t) * nrows, cudaMemcpyHostToDevice, ctx->stream));
        P->centers = dvec_alloc(ctx, nrows * N, nullptr);
        P->radii = dvec_alloc(ctx, nrows * N, nullptr);
        P->cW = dvec_alloc(ctx, nrows, nullptr);
        P->rW = dvec_alloc(ctx, nrows, nullptr);
        if (N >= 2) {
            if (s > 1) {
                P->PowA = dmat_alloc(ctx, n, n);
                P->PowB = dmat_alloc(ctx, n, n);
            }
            P->PowT = dmat_alloc(ctx, n, n);
            P->BtAug = dmat_alloc(ctx, n + m + 1, n);
            P->BtSmall = dmat_alloc(ctx, m + 1, n);
            P->S[0] = dmat_alloc(ctx, s * P->Rp, n);
            if ((N - 1) > s) P->S[1] = dmat_alloc(ctx, s * P->Rp, n);
            P->E = dmat_alloc(ctx, s * P->Rp, m + 1);
            P->bsum = dvec_alloc(ctx, s * P->Rp + 128, nullptr);
            P->fsum = dvec_alloc(ctx, s * P->Rp + 128, nullptr);
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
    for (DMat* d : {&P->Phi, &P->PowA, &P->PowB, &P->PowT, &P->BtAug, &P->BtSmall, &P->S[0], &P->S[1], &P->E, &P->MU})
        dmat_free(*d);
    for (double* v : {P->c0, P->r0, P->rpsi, P->cU, P->rU, P->bsum, P->fsum, P->cW, P->rW, P->centers, P->radii})
        if (v) cudaFree(v);
    if (P->rows) cudaFree(P->rows);
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
        // doubling: stack blocks [have, 2*have) = stack blocks [0, have) · Φ^have ;  Φ^{2 have} = Φ^have · Φ^have
        const DMat* pow = &P->Phi;
        DMat* spare[2] = {&P->PowA, &P->PowB};
        int sp = 0;
        for (int64_t have = 1; have < s; have *= 2) {
            transpose(ctx, n, n, pow->p, pow->ld, P->PowT.p, P->PowT.ld);
            cudaEvent_t e0 = next_event(), e1 = next_event();
            REACH_CUDA(cudaEventRecord(e0, st));
            gemm_nt(ctx, have * Rp, n, n, P->S[0].p, P->S[0].ld, P->PowT.p, P->PowT.ld, P->S[0].p + have * Rp, P->S[0].ld);
            gemm_nt(ctx, n, n, n, pow->p, pow->ld, P->PowT.p, P->PowT.ld, spare[sp]->p, spare[sp]->ld);
            REACH_CUDA(cudaEventRecord(e1, st));
            P->gemm_launches += 2;
            P->gemm_flops += 2.0 * (double)(have * nrows) * n * n + 2.0 * (double)n * n * n;
            pow = spare[sp];
            sp ^= 1;
        }
        // B_aug^T = [Φ^s | δB | c0]^T
        transpose(ctx, n, n, pow->p, pow->ld, P->BtAug.p, P->BtAug.ld);
        if (m > 0) transpose(ctx, n, m, P->MU.p, P->MU.ld, P->BtAug.p + n, P->BtAug.ld);
        transpose(ctx, n, 1, P->c0, n, P->BtAug.p + n + m, P->BtAug.ld);

        const int64_t npow = N - 1;
        const int64_t nblocks = ceil_div(npow, s);
        int cur = 0;
        for (int64_t b = 0; b < nblocks; ++b) {
            const int64_t cnt = std::min<int64_t>(s, npow - b * s);
            const bool last = (b == nblocks - 1);
            cudaEvent_t e0 = next_event(), e1 = next_event();
            REACH_CUDA(cudaEventRecord(e0, st));
            if (!last) {
                GemmEpilogue ep;
                ep.C2 = P->E.p; ep.ldc2 = P->E.ld; ep.nsplit = n;
                gemm_nt(ctx, s * Rp, n + m + 1, n, P->S[cur].p, P->S[cur].ld, P->BtAug.p, P->BtAug.ld, P->S[cur ^ 1].p,
                        P->S[cur ^ 1].ld, ep);
                P->gemm_flops += 2.0 * (double)(s * nrows) * n * (double)(n + m + 1);
            } else {
                gemm_nt(ctx, cnt * Rp, m + 1, n, P->S[cur].p, P->S[cur].ld, P->BtSmall.p, P->BtSmall.ld, P->E.p, P->E.ld);
                P->gemm_flops += 2.0 * (double)(cnt * nrows) * n * (double)(m + 1);
            }
            REACH_CUDA(cudaEventRecord(e1, st));
            P->gemm_launches++;
            const unsigned grid = (unsigned)ceil_div(cnt * Rp, 64);
            if (m > 0)
                abs_rowsums_kernel<true><<<grid, 256, 0, st>>>(cnt * Rp, (int)n, P->S[cur].p, P->S[cur].ld, P->r0, P->rpsi,
                                                               P->bsum, P->fsum);
            else
                abs_rowsums_kernel<false><<<grid, 256, 0, st>>>(cnt * Rp, (int)n, P->S[cur].p, P->S[cur].ld, P->r0, nullptr,
                                                                P->bsum, nullptr);
            REACH_CUDA(cudaGetLastError());
            scan_block_kernel<<<(unsigned)ceil_div(nrows, 64), 64, 0, st>>>(
                (int)nrows, (int)m, Rp, (int)cnt, b * s + 1, N, P->E.p, P->E.ld, P->bsum, P->fsum, P->cU, P->rU, P->cW,
                P->rW, P->centers, P->radii);
            REACH_CUDA(cudaGetLastError());
            ctx->launches += 2;
            cur ^= 1;
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
