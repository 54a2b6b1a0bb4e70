// C ABI: context management and the GEMM test hook (include/reach.h).
#include "common.cuh"
#include "../../include/reach.h"
#include <memory>

using namespace reach;

struct reach_ctx {
    Ctx c;
    std::vector<reach_ctx*> peers;      // reach_ctx_create_multi: the contexts of the other devices (owned)
};

namespace reach {
thread_local std::string g_create_error;
Ctx* ctx_of(reach_ctx* h) { return &h->c; }
// every device context of a handle, the primary one first
std::vector<Ctx*> ctxs_of(reach_ctx* h) {
    std::vector<Ctx*> v{&h->c};
    for (reach_ctx* p : h->peers) v.push_back(&p->c);
    return v;
}
}  // namespace reach

#define REACH_TRY(ctxptr) try {
#define REACH_CATCH(ctxptr)                                              \
    }                                                                    \
    catch (const reach::Error& e) {                                      \
        if (ctxptr) (ctxptr)->c.last_error = e.what();                   \
        return e.code;                                                   \
    }                                                                    \
    catch (const std::exception& e) {                                    \
        if (ctxptr) (ctxptr)->c.last_error = e.what();                   \
        return REACH_ECUDA;                                              \
    }                                                                    \
    catch (...) {                                                        \
        if (ctxptr) (ctxptr)->c.last_error = "unknown exception";        \
        return REACH_ECUDA;                                              \
    }

extern "C" {

int32_t reach_version(void) { return 100; }

int32_t reach_ctx_create(int32_t device, reach_ctx** out) { return reach_ctx_create_on_stream(device, nullptr, out); }

int32_t reach_ctx_create_on_stream(int32_t device, void* stream, reach_ctx** out) {
    if (!out) return REACH_EINVAL;
    *out = nullptr;
    try {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw Error(REACH_ECUDA, std::string("no CUDA device visible (libreach_sm100a has no CPU path): ") +
                                         cudaGetErrorString(e));
        REACH_REQUIRE(device >= 0 && device < count, "reach_ctx_create: device index out of range");
        cudaDeviceProp prop;
        REACH_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            throw Error(REACH_EUNSUPPORTED, std::string("device '") + prop.name + "' is not sm_100 (B200)");
        REACH_CUDA(cudaSetDevice(device));
        std::unique_ptr<reach_ctx> h(new reach_ctx());
        h->c.device = device;
        h->c.sm_count = prop.multiProcessorCount;
        if (stream) {
            h->c.stream = (cudaStream_t)stream;
            h->c.owns_stream = false;
        } else {
            REACH_CUDA(cudaStreamCreateWithFlags(&h->c.stream, cudaStreamNonBlocking));
        }
        REACH_CUDA(cudaEventCreate(&h->c.ev0));
        REACH_CUDA(cudaEventCreate(&h->c.ev1));
        *out = h.release();
        return REACH_OK;
    } catch (const Error& e) {
        g_create_error = e.what();
        return e.code;
    } catch (const std::exception& e) {       // nothing may unwind through the C ABI (a Julia ccall cannot catch it)
        g_create_error = e.what();
        return REACH_ECUDA;
    } catch (...) {
        g_create_error = "unknown exception in reach_ctx_create";
        return REACH_ECUDA;
    }
}

int32_t reach_ctx_create_multi(const int32_t* devices, int32_t ndev, reach_ctx** out) {
    if (!out) return REACH_EINVAL;
    *out = nullptr;
    if (!devices || ndev < 1) {
        g_create_error = "reach_ctx_create_multi: devices / ndev";
        return REACH_EINVAL;
    }
    reach_ctx* first = nullptr;
    int32_t rc = reach_ctx_create(devices[0], &first);
    if (rc != REACH_OK) return rc;
    for (int32_t i = 1; i < ndev; ++i) {
        reach_ctx* peer = nullptr;
        rc = reach_ctx_create(devices[i], &peer);
        if (rc != REACH_OK) {
            reach_ctx_destroy(first);
            return rc;
        }
        try {
            first->peers.push_back(peer);
        } catch (...) {
            reach_ctx_destroy(peer);
            reach_ctx_destroy(first);
            g_create_error = "out of memory";
            return REACH_ECUDA;
        }
    }
    cudaSetDevice(devices[0]);
    *out = first;
    return REACH_OK;
}

int32_t reach_ctx_device_count(reach_ctx* ctx, int32_t* ndev) {
    if (!ctx || !ndev) return REACH_EINVAL;
    *ndev = 1 + (int32_t)ctx->peers.size();
    return REACH_OK;
}

void reach_ctx_destroy(reach_ctx* ctx) {
    if (!ctx) return;
    for (reach_ctx* p : ctx->peers) reach_ctx_destroy(p);
    ctx->peers.clear();
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    for (void* p : ctx->c.scratch) cudaFree(p);
    ctx_release_pool(&ctx->c);
    if (ctx->c.ev0) cudaEventDestroy(ctx->c.ev0);
    if (ctx->c.ev1) cudaEventDestroy(ctx->c.ev1);
    if (ctx->c.stream && ctx->c.owns_stream) cudaStreamDestroy(ctx->c.stream);
    delete ctx;
}

const char* reach_last_error(const reach_ctx* ctx) {
    return ctx ? ctx->c.last_error.c_str() : g_create_error.c_str();
}

int32_t reach_sync(reach_ctx* ctx) {
    if (!ctx) return REACH_EINVAL;
    REACH_TRY(ctx)
    REACH_CUDA(cudaSetDevice(ctx->c.device));
    REACH_CUDA(cudaStreamSynchronize(ctx->c.stream));
    return REACH_OK;
    REACH_CATCH(ctx)
}

int32_t reach_launch_count(reach_ctx* ctx, uint64_t* launches) {
    if (!ctx || !launches) return REACH_EINVAL;
    *launches = ctx->c.launches;
    for (reach_ctx* p : ctx->peers) *launches += p->c.launches;      // multi-device context: all devices
    return REACH_OK;
}

int32_t reach_dgemm(reach_ctx* ctx, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* B,
                    int64_t ldb, double* C, int64_t ldc, int32_t reps, double* kernel_ms) {
    if (!ctx) return REACH_EINVAL;
    REACH_TRY(ctx)
    Ctx* c = &ctx->c;
    REACH_CUDA(cudaSetDevice(c->device));
    REACH_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && C, "reach_dgemm: bad arguments");
    DMatGuard gA(dmat_alloc(c, M, K)), gB(dmat_alloc(c, K, N)), gBt(dmat_alloc(c, N, K)), gC(dmat_alloc(c, M, N));
    DMat &dA = gA.m, &dB = gB.m, &dBt = gBt.m, &dC = gC.m;
    dmat_upload(c, dA, A, lda);
    dmat_upload(c, dB, B, ldb);
    transpose(c, K, N, dB.p, dB.ld, dBt.p, dBt.ld);
    gemm_nt(c, M, N, K, dA.p, dA.ld, dBt.p, dBt.ld, dC.p, dC.ld);
    if (reps > 1) {
        REACH_CUDA(cudaEventRecord(c->ev0, c->stream));
        for (int i = 0; i < reps; ++i) gemm_nt(c, M, N, K, dA.p, dA.ld, dBt.p, dBt.ld, dC.p, dC.ld);
        REACH_CUDA(cudaEventRecord(c->ev1, c->stream));
        REACH_CUDA(cudaEventSynchronize(c->ev1));
        float ms = 0;
        REACH_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (kernel_ms) *kernel_ms = ms / reps;
    } else if (kernel_ms) {
        *kernel_ms = 0;
    }
    dmat_download(c, dC, C, ldc);
    REACH_CUDA(cudaStreamSynchronize(c->stream));
    return REACH_OK;
    REACH_CATCH(ctx)
}

int32_t reach_dgemm_left(reach_ctx* ctx, int64_t n, int64_t Q, int64_t K, const double* P, int64_t ldp, const double* X,
                         int64_t ldx, double* Y, int64_t ldy, int32_t reps, double* kernel_ms) {
    if (!ctx) return REACH_EINVAL;
    REACH_TRY(ctx)
    Ctx* c = &ctx->c;
    REACH_CUDA(cudaSetDevice(c->device));
    REACH_REQUIRE(n > 0 && Q > 0 && K > 0 && P && X && Y, "reach_dgemm_left: bad arguments");
    // generator layout: ld = round_up(rows, 4), 128 zeroed columns of slack
    const int64_t lx = round_up(K, 4), ly = round_up(n, 4);
    DMatGuard gP(dmat_alloc(c, n, K));
    DMat& dP = gP.m;
    DevScratch scr(c);
    double* dX = scr.get<double>((size_t)(Q + 128) * lx);
    double* dY = scr.get<double>((size_t)(Q + 128) * ly);
    REACH_CUDA(cudaMemsetAsync(dX, 0, (size_t)(Q + 128) * lx * 8, c->stream));
    REACH_CUDA(cudaMemsetAsync(dY, 0, (size_t)(Q + 128) * ly * 8, c->stream));
    dmat_upload(c, dP, P, ldp);
    REACH_CUDA(cudaMemcpy2DAsync(dX, lx * 8, X, ldx * 8, K * 8, Q, cudaMemcpyHostToDevice, c->stream));
    gemm_left(c, n, Q, K, dP.p, dP.ld, dX, lx, dY, ly);
    if (reps > 1) {
        REACH_CUDA(cudaEventRecord(c->ev0, c->stream));
        for (int i = 0; i < reps; ++i) gemm_left(c, n, Q, K, dP.p, dP.ld, dX, lx, dY, ly);
        REACH_CUDA(cudaEventRecord(c->ev1, c->stream));
        REACH_CUDA(cudaEventSynchronize(c->ev1));
        float ms = 0;
        REACH_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (kernel_ms) *kernel_ms = ms / reps;
    } else if (kernel_ms) {
        *kernel_ms = 0;
    }
    REACH_CUDA(cudaMemcpy2DAsync(Y, ldy * 8, dY, ly * 8, n * 8, Q, cudaMemcpyDeviceToHost, c->stream));
    REACH_CUDA(cudaStreamSynchronize(c->stream));
    return REACH_OK;
    REACH_CATCH(ctx)
}

}  // extern "C"
