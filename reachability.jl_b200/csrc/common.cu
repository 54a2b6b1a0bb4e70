// Device-matrix helpers and small elementwise kernels shared by the path.
#include "common.cuh"

namespace reach {

void* ctx_malloc(Ctx* ctx, size_t bytes) {
    if (bytes == 0) bytes = 8;
    auto it = ctx->pool.find(bytes);
    if (it != ctx->pool.end()) {
        void* p = it->second;
        ctx->pool.erase(it);
        ctx->pooled_bytes -= bytes;
        return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess && !ctx->pool.empty()) {      // out of memory with blocks parked in the cache: release and retry
        cudaGetLastError();
        ctx_release_pool(ctx);
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        char buf[160];
        snprintf(buf, sizeof(buf), "cudaMalloc of %.3f GB failed: %s", (double)bytes / 1e9, cudaGetErrorString(e));
        throw Error(2, buf);
    }
    ctx->block_size[p] = bytes;
    return p;
}

void ctx_free(Ctx* ctx, void* p) {
    if (!p) return;
    auto it = ctx->block_size.find(p);
    if (it == ctx->block_size.end()) { cudaFree(p); return; }
    const size_t bytes = it->second;
    if (ctx->pooled_bytes + bytes <= ctx->pool_limit) {
        ctx->pool.emplace(bytes, p);
        ctx->pooled_bytes += bytes;
    } else {
        ctx->block_size.erase(it);
        cudaFree(p);
    }
}

void ctx_release_pool(Ctx* ctx) {
    for (auto& kv : ctx->pool) {
        ctx->block_size.erase(kv.second);
        cudaFree(kv.second);
    }
    ctx->pool.clear();
    ctx->pooled_bytes = 0;
}

DMat dmat_alloc(Ctx* ctx, int64_t rows, int64_t cols, int64_t extra_rows) {
    DMat m;
    m.ctx = ctx;
    m.rows = rows;
    m.cols = cols;
    m.ld = round_up(round_up(rows, 8) + 128 + extra_rows, 16);
    m.cols_alloc = round_up(cols > 0 ? cols : 1, 128);
    m.owned = true;
    m.p = (double*)ctx_malloc(ctx, m.bytes());
    REACH_CUDA(cudaMemsetAsync(m.p, 0, m.bytes(), ctx->stream));
    return m;
}

void dmat_free(DMat& m) {
    if (m.owned && m.p) ctx_free(m.ctx, m.p);
    m.p = nullptr;
    m.owned = false;
}

void dmat_zero(Ctx* ctx, DMat& m) { REACH_CUDA(cudaMemsetAsync(m.p, 0, m.bytes(), ctx->stream)); }

void dmat_upload(Ctx* ctx, DMat& dst, const double* host, int64_t ld_host) {
    if (dst.rows == 0 || dst.cols == 0) return;
    REACH_CUDA(cudaMemcpy2DAsync(dst.p, dst.ld * sizeof(double), host, ld_host * sizeof(double),
                                 dst.rows * sizeof(double), dst.cols, cudaMemcpyHostToDevice, ctx->stream));
}

void dmat_download(Ctx* ctx, const DMat& src, double* host, int64_t ld_host) {
    if (src.rows == 0 || src.cols == 0) return;
    REACH_CUDA(cudaMemcpy2DAsync(host, ld_host * sizeof(double), src.p, src.ld * sizeof(double),
                                 src.rows * sizeof(double), src.cols, cudaMemcpyDeviceToHost, ctx->stream));
}

}  // namespace reach
