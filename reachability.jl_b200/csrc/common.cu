// Device-matrix helpers and small elementwise kernels shared by the path.
#include "common.cuh"

namespace reach {

DMat dmat_alloc(Ctx* ctx, int64_t rows, int64_t cols, int64_t extra_rows) {
    DMat m;
    m.rows = rows;
    m.cols = cols;
    m.ld = round_up(round_up(rows, 8) + 128 + extra_rows, 16);
    m.cols_alloc = round_up(cols > 0 ? cols : 1, 128);
    m.owned = true;
    REACH_CUDA(cudaMalloc(&m.p, m.bytes()));
    REACH_CUDA(cudaMemsetAsync(m.p, 0, m.bytes(), ctx->stream));
    return m;
}

void dmat_free(DMat& m) {
    if (m.owned && m.p) cudaFree(m.p);
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
