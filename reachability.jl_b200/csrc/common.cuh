// libreach_sm100a -- shared declarations for the CUDA translation units.
// B200 (sm_100a) only; no CPU fallback exists anywhere in this library.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include <map>
#include <unordered_map>
#include <cuda_runtime.h>

namespace reach {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define REACH_CUDA(expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            char _buf[512];                                                                      \
            snprintf(_buf, sizeof(_buf), "CUDA error '%s' at %s:%d (%s)", cudaGetErrorString(_e), \
                     __FILE__, __LINE__, #expr);                                                 \
            throw ::reach::Error(2, _buf);                                                       \
        }                                                                                        \
    } while (0)

#define REACH_REQUIRE(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) throw ::reach::Error(1, std::string(msg));    \
    } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// Device matrix, column-major (Julia layout), zero-padded so that every tile the GEMM touches exists:
//   rows_alloc = round_up(rows, 8) + 128 slack, cols_alloc = round_up(cols, 128)  (see DESIGN.md "HBM layout").
struct Ctx;
struct DMat {
    Ctx* ctx = nullptr;
    double* p = nullptr;
    int64_t rows = 0, cols = 0;     // logical
    int64_t ld = 0;                 // leading dimension (doubles), multiple of 16
    int64_t cols_alloc = 0;
    bool owned = false;
    size_t bytes() const { return (size_t)ld * (size_t)cols_alloc * sizeof(double); }
};

struct Ctx {
    int device = 0;
    int sm_count = 148;
    bool gemm_exclusive = false;    // the persistent GEMM asks for all of an SM's shared memory while gemm_reserve_sms > 0: a real SM partition
    int gemm_reserve_sms = 0;       // SMs the persistent GEMM leaves free (GLGM06 keeps one for its selection chain)
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    int* gemm_sched = nullptr;      // dynamic tile-scheduler words of the persistent GEMM (gemm_dmma.cu)
    uint32_t gemm_sched_next = 0;
    std::string last_error;
    // launch accounting (bench.py "gpu_launches") and per-phase timings
    int64_t csc_max_component = 0;  // last reach_bffpsv18_boxes_csc call: largest connected component of Phi's sparsity graph (0: not block-diagonal)
    int csc_sparse_path = 0;        // ... and whether the block-diagonal recurrence served it
    uint64_t launches = 0;
    double t_discretize_ms = 0, t_loop_ms = 0, t_h2d_ms = 0, t_d2h_ms = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<void*> scratch;     // freed at destroy
    // size-keyed cache of device blocks: repeated calls with the same shapes (one-shot C-ABI entry points) skip
    // cudaMalloc / cudaFree.  ctx_free parks a block at once and ctx_malloc hands it out without synchronising, so the OWNER
    // must have drained every stream that touched the block before it calls ctx_free: boxplan_destroy / flowpipe_free
    // synchronise their side streams first, DevScratch / DMatGuard synchronise the context's stream in their destructors.
    std::multimap<size_t, void*> pool;
    std::unordered_map<void*, size_t> block_size;
    size_t pooled_bytes = 0;
    size_t pool_limit = (size_t)48 << 30;
};

void* ctx_malloc(Ctx* ctx, size_t bytes);       // never returns nullptr (throws)
void ctx_free(Ctx* ctx, void* p);               // back to the cache (or cudaFree beyond the limit)
void ctx_release_pool(Ctx* ctx);

// RAII scratch of pooled device blocks for one C-ABI call: the destructor drains the context's stream first, so a block never
// goes back to the pool while a kernel or copy of this call may still touch it (also when an exception unwinds the call).
struct DevScratch {
    Ctx* ctx;
    std::vector<void*> blocks;
    explicit DevScratch(Ctx* c) : ctx(c) {}
    DevScratch(const DevScratch&) = delete;
    DevScratch& operator=(const DevScratch&) = delete;
    ~DevScratch() {
        cudaStreamSynchronize(ctx->stream);
        for (void* p : blocks) ctx_free(ctx, p);
    }
    template <typename T>
    T* get(size_t count) {
        T* p = (T*)ctx_malloc(ctx, sizeof(T) * (count ? count : 1));
        blocks.push_back(p);
        return p;
    }
};

DMat dmat_alloc(Ctx* ctx, int64_t rows, int64_t cols, int64_t extra_rows = 0);
void dmat_free(DMat& m);
void dmat_upload(Ctx* ctx, DMat& dst, const double* host, int64_t ld_host);      // host col-major -> device
void dmat_download(Ctx* ctx, const DMat& src, double* host, int64_t ld_host);
void dmat_zero(Ctx* ctx, DMat& m);
// owns a DMat for the duration of a C-ABI call; like DevScratch it drains the stream before the block is recycled
struct DMatGuard {
    DMat m;
    explicit DMatGuard(DMat d) : m(d) {}
    DMatGuard(const DMatGuard&) = delete;
    DMatGuard& operator=(const DMatGuard&) = delete;
    ~DMatGuard() {
        if (m.ctx && m.p) cudaStreamSynchronize(m.ctx->stream);
        dmat_free(m);
    }
};

// A host matrix given as Julia's SparseMatrixCSC fields (1-based Int64 indices)
struct PhiCsc {
    int64_t nnz;
    const int64_t* colptr;
    const int64_t* rowval;
    const double* nzval;
};

// ---- primitives (each is one kernel launch on ctx->stream) -------------------------------------------
// C[M x N] = A[M x K] * B[K x N], A and C column-major, B given TRANSPOSED (Bt[k*ldbt + j] = B[k, j]).
// All operands zero-padded as DMat guarantees.  Fused options: see gemm_dmma.cu.
struct GemmEpilogue {
    // C = alpha * A*B + beta_I * I (diagonal add on global row==col), used by the Taylor/Horner expm.
    double alpha = 1.0;
    double diag = 0.0;
    // columns >= nsplit of the product go to C2 (ldc2) instead of C: side products such as P·[δB | c0]
    double* C2 = nullptr;
    int64_t ldc2 = 0;
    int64_t nsplit = -1;
    // gemm_nt only -- fused weighted abs-row-sums of the product's columns [0, nsplit) (the box radii |P_k| r0 and |P_k| rψ of
    // BFFPSV18, reach_blocks.jl:191-193, 209-215, straight from the accumulators):
    //   part[(2 q + w) * ldpart + row] = sum over the 8*WNT columns of column group q of |C[row, col]| * (w ? w1 : w0)[col],
    // q = 0 .. gemm_nt_col_groups(N) - 1 in ascending column order; the caller adds the groups up in that order.
    const double* w0 = nullptr;
    const double* w1 = nullptr;     // may stay nullptr (only w0 sums are produced)
    double* part = nullptr;
    int64_t ldpart = 0;
};
// number of column groups gemm_nt produces partial abs-row-sums for (4 consumer warps per column tile)
int64_t gemm_nt_col_groups(int64_t N);
void gemm_nt(Ctx* ctx, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* Bt,
             int64_t ldbt, double* C, int64_t ldc, const GemmEpilogue& ep = GemmEpilogue());
// Y[n x Q] = P[n x K] * X[K x Q], all column-major; X and Y are generator matrices (columns contiguous, ld a
// multiple of 16 and zero-padded).  P must be zero-padded like a DMat.  ep.alpha / ep.diag apply (diag on Y[i,i]).
void gemm_left(Ctx* ctx, int64_t n, int64_t Q, int64_t K, const double* P, int64_t ldp, const double* X, int64_t ldx,
               double* Y, int64_t ldy, const GemmEpilogue& ep = GemmEpilogue());
// Bt[j + i*ldt] = A[i + j*lda]   (rows x cols -> cols x rows), out-of-place
void transpose(Ctx* ctx, int64_t rows, int64_t cols, const double* A, int64_t lda, double* At, int64_t ldt);

}  // namespace reach
