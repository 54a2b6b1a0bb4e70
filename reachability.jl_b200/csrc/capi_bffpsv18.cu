// C ABI: BFFPSV18 entry points (include/reach.h).
#include "common.cuh"
#include <algorithm>
#include <thread>
#include "../../include/reach.h"

namespace reach {
struct BoxPlan;
BoxPlan* boxplan_create(Ctx* ctx, int64_t n, const double* Phi, int64_t ldphi, const double* c0, const double* r0,
                        int64_t m, const double* MU, int64_t ldmu, const double* cU, const double* rU,
                        const double* rpsi, int64_t nrows, const int32_t* rows, int64_t N, const PhiCsc* csc = nullptr);
void boxplan_destroy(BoxPlan* P);
void boxplan_run(BoxPlan* P);
void boxplan_fetch(BoxPlan* P, double* centers, double* radii);
void boxplan_fetch_strided(BoxPlan* P, double* centers, double* radii, int64_t ld_out);
std::vector<Ctx*> ctxs_of(reach_ctx* h);
void boxplan_results(BoxPlan* P, double** c, double** r);
void boxplan_timing(BoxPlan* P, double* run_ms, double* gemm_ms, int64_t* launches, double* flops);
void boxplan_attach_map(BoxPlan* P, int64_t q, const double* M, int64_t ldm, int64_t nblocks, const int32_t* block_sizes,
                        bool check, double bound, bool eager);
void boxplan_fetch_map(BoxPlan* P, double* centers, double* radii, int64_t* violation);
Ctx* ctx_of(reach_ctx* h);
bool boxes_blockdiag(Ctx* ctx, int64_t n, const PhiCsc& csc, const double* c0, const double* r0, int64_t m, const double* MU,
                     int64_t ldmu, const double* cU, const double* rU, const double* rpsi, int64_t nrows, const int32_t* rows,
                     int64_t N, double* centers, double* radii, int64_t* max_component);
}  // namespace reach

using namespace reach;

struct reach_boxplan {
    BoxPlan* p;
    Ctx* ctx;
};

#define GUARD_BEGIN(ctx) \
    if (!(ctx)) return REACH_EINVAL; \
    try { \
        REACH_CUDA(cudaSetDevice((ctx)->device));
#define GUARD_END(ctx) \
    } catch (const reach::Error& e) { (ctx)->last_error = e.what(); return e.code; } \
    catch (const std::exception& e) { (ctx)->last_error = e.what(); return REACH_ECUDA; } \
    catch (...) { (ctx)->last_error = "unknown exception"; return REACH_ECUDA; }

namespace {

// BFFPSV18 on a multi-device context (reach_ctx_create_multi): the rows of interest are sharded in
// contiguous ranges over the devices, Phi and the sets are replicated, and every device runs the complete recurrence on
// its rows -- P_k[R_g, :] = P_{k-1}[R_g, :] Phi needs nothing from the other shards, so there is NO collective: one host
// thread per device drives upload, time loop and the strided download into the caller's arrays ("gather only at the end").
// For the box set types every coordinate is independent, so a shard boundary may fall inside a partition block.
void boxes_multi(const std::vector<Ctx*>& cs, int64_t n, const double* Phi, int64_t ldphi, const PhiCsc* csc, const double* c0,
                 const double* r0, int64_t m, const double* MU, int64_t ldmu, const double* cU, const double* rU, const double* rpsi,
                 int64_t nrows, const int32_t* rows, int64_t N, double* centers, double* radii) {
    const int64_t G = std::min<int64_t>((int64_t)cs.size(), nrows);
    struct Part { int code = 0; std::string msg; };
    std::vector<Part> parts((size_t)G);
    std::vector<std::thread> th;
    for (int64_t g = 0; g < G; ++g) {
        th.emplace_back([&, g]() {
            Ctx* ctx = cs[(size_t)g];
            const int64_t lo = nrows * g / G, hi = nrows * (g + 1) / G;
            BoxPlan* p = nullptr;
            try {
                REACH_CUDA(cudaSetDevice(ctx->device));
                p = boxplan_create(ctx, n, Phi, ldphi, c0, r0, m, MU, ldmu, cU, rU, rpsi, hi - lo, rows + lo, N, csc);
                boxplan_run(p);
                boxplan_fetch_strided(p, centers + lo, radii + lo, nrows);
            } catch (const reach::Error& e) {
                parts[(size_t)g].code = e.code; parts[(size_t)g].msg = e.what();
            } catch (const std::exception& e) {
                parts[(size_t)g].code = REACH_ECUDA; parts[(size_t)g].msg = e.what();
            } catch (...) {
                parts[(size_t)g].code = REACH_ECUDA; parts[(size_t)g].msg = "unknown exception";
            }
            if (p) boxplan_destroy(p);
        });
    }
    for (std::thread& t : th) t.join();
    for (int64_t g = 0; g < G; ++g)
        if (parts[(size_t)g].code)
            throw reach::Error(parts[(size_t)g].code, "device " + std::to_string(cs[(size_t)g]->device) + ": " + parts[(size_t)g].msg);
}

}  // namespace

extern "C" {

int32_t reach_boxplan_create(reach_ctx* h, int64_t n, const double* Phi, int64_t ldphi, const double* c0,
                             const double* r0, int64_t m, const double* MU, int64_t ldmu, const double* cU,
                             const double* rU, const double* rpsi, int64_t nrows, const int32_t* rows, int64_t N,
                             reach_boxplan** out) {
    if (!h || !out) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    *out = nullptr;
    GUARD_BEGIN(ctx)
    BoxPlan* p = boxplan_create(ctx, n, Phi, ldphi, c0, r0, m, MU, ldmu, cU, rU, rpsi, nrows, rows, N);
    *out = new reach_boxplan{p, ctx};
    return REACH_OK;
    GUARD_END(ctx)
}

int32_t reach_boxplan_run(reach_boxplan* plan) {
    if (!plan) return REACH_EINVAL;
    GUARD_BEGIN(plan->ctx)
    boxplan_run(plan->p);
    return REACH_OK;
    GUARD_END(plan->ctx)
}

int32_t reach_boxplan_fetch(reach_boxplan* plan, double* centers, double* radii) {
    if (!plan) return REACH_EINVAL;
    GUARD_BEGIN(plan->ctx)
    boxplan_fetch(plan->p, centers, radii);
    return REACH_OK;
    GUARD_END(plan->ctx)
}

int32_t reach_boxplan_device_results(reach_boxplan* plan, double** centers_dev, double** radii_dev) {
    if (!plan) return REACH_EINVAL;
    boxplan_results(plan->p, centers_dev, radii_dev);
    return REACH_OK;
}

int32_t reach_boxplan_timing(reach_boxplan* plan, double* run_ms, double* gemm_ms, int64_t* gemm_launches,
                             double* gemm_flops) {
    if (!plan) return REACH_EINVAL;
    boxplan_timing(plan->p, run_ms, gemm_ms, gemm_launches, gemm_flops);
    return REACH_OK;
}

void reach_boxplan_destroy(reach_boxplan* plan) {
    if (!plan) return;
    cudaSetDevice(plan->ctx->device);
    cudaStreamSynchronize(plan->ctx->stream);
    boxplan_destroy(plan->p);
    delete plan;
}

int32_t reach_bffpsv18_boxes(reach_ctx* h, int64_t n, const double* Phi, int64_t ldphi, const double* c0,
                             const double* r0, int64_t m, const double* MU, int64_t ldmu, const double* cU,
                             const double* rU, const double* rpsi, int64_t nrows, const int32_t* rows, int64_t N,
                             double* centers, double* radii, int64_t* steps_done) {
    if (!h) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    GUARD_BEGIN(ctx)
    REACH_REQUIRE(centers && radii, "reach_bffpsv18_boxes: output buffers required");
    const std::vector<Ctx*> cs = ctxs_of(h);
    if (cs.size() > 1 && nrows > 1 && rows) {
        boxes_multi(cs, n, Phi, ldphi, nullptr, c0, r0, m, MU, ldmu, cU, rU, rpsi, nrows, rows, N, centers, radii);
        if (steps_done) *steps_done = N;
        return REACH_OK;
    }
    BoxPlan* p = boxplan_create(ctx, n, Phi, ldphi, c0, r0, m, MU, ldmu, cU, rU, rpsi, nrows, rows, N);
    try {
        boxplan_run(p);
        boxplan_fetch(p, centers, radii);
    } catch (...) {
        boxplan_destroy(p);
        throw;
    }
    boxplan_destroy(p);
    if (steps_done) *steps_done = N;
    return REACH_OK;
    GUARD_END(ctx)
}

int32_t reach_bffpsv18_csc_info(reach_ctx* h, int64_t* max_component, int32_t* sparse_path) {
    if (!h || !max_component || !sparse_path) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    *max_component = ctx->csc_max_component;
    *sparse_path = ctx->csc_sparse_path;
    return REACH_OK;
}

int32_t reach_bffpsv18_boxes_csc(reach_ctx* h, int64_t n, int64_t nnz, const int64_t* colptr, const int64_t* rowval,
                                 const double* nzval, const double* c0, const double* r0, int64_t m, const double* MU,
                                 int64_t ldmu, const double* cU, const double* rU, const double* rpsi, int64_t nrows,
                                 const int32_t* rows, int64_t N, double* centers, double* radii, int64_t* steps_done) {
    if (!h) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    GUARD_BEGIN(ctx)
    REACH_REQUIRE(centers && radii, "reach_bffpsv18_boxes_csc: output buffers required");
    const PhiCsc csc{nnz, colptr, rowval, nzval};
    // decoupled subsystems: Phi (and every power of it) is block-diagonal up to a permutation -> the zero blocks are skipped
    // for good (reach_blocks.jl:47-132); REACH_BFFPSV18_CSC_DENSE=1 keeps the dense DMMA recurrence (A/B tests)
    ctx->csc_max_component = 0;
    ctx->csc_sparse_path = 0;
    if (!getenv("REACH_BFFPSV18_CSC_DENSE") && colptr && rows &&
        boxes_blockdiag(ctx, n, csc, c0, r0, m, MU, ldmu, cU, rU, rpsi, nrows, rows, N, centers, radii, &ctx->csc_max_component)) {
        ctx->csc_sparse_path = 1;
        if (steps_done) *steps_done = N;
        return REACH_OK;
    }
    const std::vector<Ctx*> cs = ctxs_of(h);
    if (cs.size() > 1 && nrows > 1 && rows) {
        boxes_multi(cs, n, nullptr, 0, &csc, c0, r0, m, MU, ldmu, cU, rU, rpsi, nrows, rows, N, centers, radii);
        if (steps_done) *steps_done = N;
        return REACH_OK;
    }
    BoxPlan* p = boxplan_create(ctx, n, nullptr, 0, c0, r0, m, MU, ldmu, cU, rU, rpsi, nrows, rows, N, &csc);
    try {
        boxplan_run(p);
        boxplan_fetch(p, centers, radii);
    } catch (...) {
        boxplan_destroy(p);
        throw;
    }
    boxplan_destroy(p);
    if (steps_done) *steps_done = N;
    return REACH_OK;
    GUARD_END(ctx)
}

int32_t reach_bffpsv18_check(reach_ctx* h, int64_t n, const double* Phi, int64_t ldphi, const double* c0,
                             const double* r0, int64_t m, const double* MU, int64_t ldmu, const double* cU,
                             const double* rU, const double* rpsi, int64_t nrows, const int32_t* rows, int64_t N,
                             int64_t nblocks, const int32_t* block_sizes, const double* ell, double bound, int32_t eager,
                             int64_t* violation) {
    if (!h) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    GUARD_BEGIN(ctx)
    REACH_REQUIRE(ell && violation, "reach_bffpsv18_check: ell and violation are required");
    BoxPlan* p = boxplan_create(ctx, n, Phi, ldphi, c0, r0, m, MU, ldmu, cU, rU, rpsi, nrows, rows, N);
    try {
        boxplan_attach_map(p, 1, ell, 1, nblocks, block_sizes, true, bound, eager != 0);
        boxplan_run(p);
        boxplan_fetch_map(p, nullptr, nullptr, violation);
    } catch (...) {
        boxplan_destroy(p);
        throw;
    }
    boxplan_destroy(p);
    return REACH_OK;
    GUARD_END(ctx)
}

int32_t reach_bffpsv18_output_map(reach_ctx* h, int64_t n, const double* Phi, int64_t ldphi, const double* c0,
                                  const double* r0, int64_t m, const double* MU, int64_t ldmu, const double* cU,
                                  const double* rU, const double* rpsi, int64_t nrows, const int32_t* rows, int64_t N,
                                  int64_t nblocks, const int32_t* block_sizes, int64_t q, const double* Mout, int64_t ldm,
                                  double* centers, double* radii) {
    if (!h) return REACH_EINVAL;
    Ctx* ctx = ctx_of(h);
    GUARD_BEGIN(ctx)
    REACH_REQUIRE(centers && radii && q >= 1 && Mout && ldm >= q, "reach_bffpsv18_output_map: bad output map / buffers");
    // the map kernels carry up to 8 output rows; wider maps run in groups of 8 rows (the recurrence is repeated per group)
    constexpr int64_t QMAX = 8;
    std::vector<double> Mg, cg, rg;
    for (int64_t q0 = 0; q0 < q; q0 += QMAX) {
        const int64_t qc = std::min(QMAX, q - q0);
        const bool whole = (q <= QMAX);
        const double* Mp = Mout;
        int64_t ldp = ldm;
        if (!whole) {
            Mg.assign((size_t)qc * nrows, 0.0);
            for (int64_t l = 0; l < nrows; ++l)
                for (int64_t a = 0; a < qc; ++a) Mg[(size_t)l * qc + a] = Mout[l * ldm + q0 + a];
            cg.assign((size_t)qc * N, 0.0);
            rg.assign((size_t)qc * N, 0.0);
            Mp = Mg.data();
            ldp = qc;
        }
        BoxPlan* p = boxplan_create(ctx, n, Phi, ldphi, c0, r0, m, MU, ldmu, cU, rU, rpsi, nrows, rows, N);
        try {
            boxplan_attach_map(p, qc, Mp, ldp, nblocks, block_sizes, false, 0.0, false);
            boxplan_run(p);
            boxplan_fetch_map(p, whole ? centers : cg.data(), whole ? radii : rg.data(), nullptr);
        } catch (...) {
            boxplan_destroy(p);
            throw;
        }
        boxplan_destroy(p);
        if (!whole)
            for (int64_t k = 0; k < N; ++k)
                for (int64_t a = 0; a < qc; ++a) {
                    centers[k * q + q0 + a] = cg[(size_t)k * qc + a];
                    radii[k * q + q0 + a] = rg[(size_t)k * qc + a];
                }
    }
    return REACH_OK;
    GUARD_END(ctx)
}

}  // extern "C"
