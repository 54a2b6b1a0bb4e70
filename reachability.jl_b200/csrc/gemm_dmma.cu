// FP64 GEMM for sm_100a: persistent, warp-specialised, TMA-bulk-copy (UBLKCP) -> mbarrier ring -> DMMA.
//
// tcgen05 has no f64 kind, so the FP64 tensor path on B200 is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4),
// measured at 37.06 TFLOP/s register-resident on this pool's B200 (profiles/r01/fp64_peak_b200.jsonl).
// This kernel is the engine of every dense contraction on the path:
//   * BFFPSV18  P_{k+s}[R,:] = P_k[R,:] · Φ^s  on a stack of s row-blocks (reach_blocks.jl:217-218)
//   * GLGM06    Φ^k·[G0 | c0], Φ^k·[GV | cV], Φ^k·Φ (GLGM06/reach.jl:48,54,57)
//   * discretize  Taylor/Horner products of e^{Aδ}, Φ₁, Φ₂ (discretize.jl:150-166, 221-315), A·A (:655)
//
// Operands: A [M x K] column-major (M contiguous), B passed TRANSPOSED as Bt [K x N] with N contiguous,
// C [M x N] column-major.  Both operand tiles are therefore K "rows" of 1 KB contiguous memory, each moved
// by one cp.async.bulk (TMA engine, no tensor map needed) into a padded smem row (pitch 132 doubles) so that
// the m8n8k4 fragment loads (lane -> (k = lane%4, row = lane/4)) hit 16 distinct 8-byte bank slots per
// half-warp: slot = 4*(lane%4) + lane/4 (mod 16).
#include "common.cuh"
#include "tma.cuh"
#include <mutex>
#include <cuda.h>      // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)

namespace reach {

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int STAGES = 4;
constexpr int PITCH_A = BM + 4;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;

// 2-D tiled TMA load through a tensor map (coordinates: c0 = innermost = k, c1 = row)
__device__ __forceinline__ void tma_2d_g2s(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

struct GemmParams {
    const double* A;
    const double* Bt;
    double* C;
    int64_t lda, ldbt, ldc;
    int M, N, ksteps;
    int tiles_m, tiles_n;
    double alpha, diag;
    double* C2;      // columns >= nsplit are stored to C2[(col - nsplit) * ldc2 + row] (fused side products)
    int64_t ldc2;
    int nsplit;
    const double* w0;   // fused weighted abs-row-sums (NT variant): see GemmEpilogue
    const double* w1;
    double* part;
    int64_t ldpart;
    int* sched;      // sched[0] = dynamic tile counter, sched[1] = CTAs that finished fetching (both zero between launches)
};

// WNT = n8 tiles per warp along N; CTA tile is 128 x (32*WNT).
// AKM = false ("NT"): A column-major (M contiguous), C column-major           -> C[row + col*ldc]
// AKM = true  ("TN"): A given with K contiguous (row m of A = lda-strided run of K doubles), C stored with N
//                     contiguous -> C[row*ldc + col].  With A = X^T (X n x Q column-major generators) and
//                     Bt = Phi^s column-major this is  Y = Phi^s * X  with Y column-major again: the left
//                     multiplication of GLGM06 (linear_map, GLGM06/reach.jl:16,48,54) with no transposes.
//                     The 128 x 16 A tile is ONE 2-D tensor-map TMA (cp.async.bulk.tensor.2d, SASS UTMALDG) per stage
//                     with the 128-byte swizzle; out-of-range rows / k are zero-filled by the TMA unit.  (128 separate
//                     128-byte bulk copies per stage starved the DMMA pipe: 12 TFLOP/s, ncu: long-scoreboard on the
//                     full barrier.)  Bank conflicts are avoided by letting lane t = 2*th + b of a quad feed
//                     k = 2*(kk + 4*th) + b to the kk-th DMMA of a stage (any 4 distinct k per DMMA are legal as long
//                     as A and B agree): swizzled A chunk (kk + 4 th) ^ g is distinct over the half-warp, and the B rows
//                     k >= 8 are shifted by 8 doubles so that slot 8kk + 4b + 8th + g (mod 16) is distinct too.
template <int WNT, bool AKM>
__global__ void __launch_bounds__(THREADS, 1) gemm_dmma_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmapA) {
    constexpr int BN = 32 * WNT;
    constexpr int PITCH_B = BN + 4;
    constexpr int A_STAGE = AKM ? BM * BK : BK * PITCH_A;     // doubles (TN: 16 KB, 1024-byte aligned swizzle atoms)
    constexpr int B_STAGE = BK * PITCH_B + (AKM ? 8 : 0);
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sB = sA + STAGES * A_STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);
    uint64_t* empty = full + STAGES;
    int* tile_ring = reinterpret_cast<int*>(empty + STAGES);   // tile id travelling with the first k-step of every tile
    // fused abs-sum epilogue (NT variant): every consumer warp keeps the weights of ITS 8*WNT columns, [warp][w][column], fetched
    // with cp.async at the start of the tile so that the epilogue finds them in shared memory without a CTA-wide barrier
    double* wsm = reinterpret_cast<double*>(tile_ring + STAGES + (STAGES & 1));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int num_tiles = p.tiles_m * p.tiles_n;
    if (warp == CONSUMER_WARPS) {
        // ===== producer warp: lanes 0..15 move A rows, lanes 16..31 move Bt rows of each k-step =====
        // Tiles are handed out DYNAMICALLY (first one = blockIdx.x, the rest from an atomic counter): the other kernels of
        // the path share SMs with this one (GLGM06's selection chain owns most of one SM), and with a static round-robin
        // the slowest SM would set the duration of the whole launch.
        uint32_t it = 0;
        int tile = blockIdx.x;
        while (true) {
            if (tile >= num_tiles) {                       // sentinel: tells the consumers to leave
                const int s = it % STAGES;
                mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                if (lane == 0) {
                    tile_ring[s] = -1;
                    mbar_arrive(&full[s]);
                    // the last CTA to stop fetching re-arms the scheduler words for the next launch
                    __threadfence();
                    if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
                        p.sched[0] = 0;
                        __threadfence();
                        p.sched[1] = 0;
                    }
                }
                break;
            }
            const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
            const double* gA = AKM ? p.A + (int64_t)tm * BM * p.lda : p.A + (int64_t)tm * BM;
            const double* gB = p.Bt + (int64_t)tn * BN;
            for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                if (lane == 0) {
                    if (ks == 0) tile_ring[s] = tile;
                    mbar_expect_tx(&full[s], (uint32_t)(BK * (BM + BN) * sizeof(double)));
                }
                __syncwarp();
                const int k = ks * BK + (lane & 15);
                if (AKM) {
                    if (lane == 0) tma_2d_g2s(sA + s * A_STAGE, &tmapA, ks * BK, tm * BM, &full[s]);
                    if (lane < 16)
                        bulk_g2s(sB + s * B_STAGE + lane * PITCH_B + (lane >= 8 ? 8 : 0), gB + (int64_t)k * p.ldbt, BN * 8,
                                 &full[s]);
                } else if (lane < 16)
                    bulk_g2s(sA + s * A_STAGE + lane * PITCH_A, gA + (int64_t)k * p.lda, BM * 8, &full[s]);
                else
                    bulk_g2s(sB + s * B_STAGE + (lane - 16) * PITCH_B, gB + (int64_t)k * p.ldbt, BN * 8, &full[s]);
            }
            if (lane == 0) tile = (int)gridDim.x + atomicAdd(p.sched, 1);
            tile = __shfl_sync(0xffffffffu, tile, 0);
        }
    } else {
        // ===== consumer warps: 2 (M) x 4 (N); warp tile 64 x 8*WNT =====
        const int wm = warp >> 2, wn = warp & 3;
        const int g = lane >> 2, t = lane & 3;
        uint32_t it = 0;
        while (true) {
            mbar_wait(&full[it % STAGES], (it / STAGES) & 1);
            const int tile = tile_ring[it % STAGES];
            if (tile < 0) break;
            const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
            if (!AKM && p.part) {
                __syncwarp();                     // the previous tile's epilogue has read the warp's weights
                if (lane < 8 * WNT) {
                    const int col = tn * BN + wn * (8 * WNT) + lane;
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        const double* wv = w ? p.w1 : p.w0;
                        double* dst = wsm + (warp * 2 + w) * 32 + lane;
                        if (wv != nullptr && col < p.nsplit)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(wv + col) : "memory");
                        else
                            *dst = 0.0;
                    }
                }
            }
            double acc[8][WNT][2];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < WNT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

            for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full[s], ph);
                const int th = t >> 1, tb = t & 1;
                const double* a_s = AKM ? sA + s * A_STAGE + (wm * 64 + g) * BK + tb
                                        : sA + s * A_STAGE + t * PITCH_A + wm * 64 + g;
                const double* b_s = AKM ? sB + s * B_STAGE + (8 * th + tb) * PITCH_B + 8 * th + wn * (8 * WNT) + g
                                        : sB + s * B_STAGE + t * PITCH_B + wn * (8 * WNT) + g;
#pragma unroll
                for (int kk = 0; kk < BK / 4; ++kk) {
                    double a[8], b[WNT];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        a[i] = AKM ? a_s[i * 8 * BK + ((((kk + 4 * th) ^ g)) << 1)] : a_s[kk * 4 * PITCH_A + i * 8];
#pragma unroll
                    for (int j = 0; j < WNT; ++j) b[j] = AKM ? b_s[2 * kk * PITCH_B + j * 8] : b_s[kk * 4 * PITCH_B + j * 8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < WNT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
            // ---- epilogue: masked direct stores (8 rows x 8 B = two full 32 B sectors per column) ----
            const int row0 = tm * BM + wm * 64 + g;
            const int col0 = tn * BN + wn * (8 * WNT) + 2 * t;
            if constexpr (AKM) {
                // C[row*ldc + col]: each thread owns two adjacent columns -> one 16-byte store; 4 lanes cover 64 B.
                // (col, col+1) is stored whenever col < N: for odd N the pad entry receives the exact zero of the
                // zero-padded operand (ldc >= round_up(N, 2) by the DMat/generator layout).
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = row0 + i * 8;
                    if (row < p.M) {
                        double* crow = p.C + (int64_t)row * p.ldc;
#pragma unroll
                        for (int j = 0; j < WNT; ++j) {
                            const int col = col0 + j * 8;
                            if (col < p.N) {
                                double2 v = make_double2(p.alpha * acc[i][j][0], p.alpha * acc[i][j][1]);
                                if (row == col) v.x += p.diag;
                                if (row == col + 1) v.y += p.diag;
                                *reinterpret_cast<double2*>(crow + col) = v;
                            }
                        }
                    }
                }
            } else {
            if (p.part) {
                // Fused abs-sum epilogue: the warp holds a 64 x 8*WNT piece of P_k; every row's weighted abs-sum over these
                // columns is a per-thread fma chain (ascending columns) followed by a butterfly over the quad.  One partial
                // per (column group, row) leaves the SM instead of the 130 MB stack being re-read by a second kernel.
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();                     // this tile's weights (fetched by lanes 0 .. 8 WNT - 1) are visible to the warp
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    if ((w ? p.w1 : p.w0) == nullptr) continue;
                    const double* wsrc = wsm + (warp * 2 + w) * 32 + 2 * t;
                    double wc[WNT][2];
#pragma unroll
                    for (int j = 0; j < WNT; ++j) {
                        const double2 v = *reinterpret_cast<const double2*>(wsrc + j * 8);
                        wc[j][0] = v.x; wc[j][1] = v.y;
                    }
                    double* pout = p.part + (int64_t)((tn * 4 + wn) * 2 + w) * p.ldpart;
                    double sv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        double sacc = 0.0;
#pragma unroll
                        for (int j = 0; j < WNT; ++j)
#pragma unroll
                            for (int e = 0; e < 2; ++e) sacc = fma(fabs(acc[i][j][e]), wc[j][e], sacc);
                        sv[i] = sacc;
                    }
                    // reduce-scatter over the quad (6 shuffles instead of 16): lane t ends up with the sums of rows 2t, 2t+1
                    double k4[4], k2[2];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double send = (t & 2) ? sv[k] : sv[4 + k];
                        const double recv = __shfl_xor_sync(0xffffffffu, send, 2);
                        k4[k] = ((t & 2) ? sv[4 + k] : sv[k]) + recv;
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const double send = (t & 1) ? k4[k] : k4[2 + k];
                        const double recv = __shfl_xor_sync(0xffffffffu, send, 1);
                        k2[k] = ((t & 1) ? k4[2 + k] : k4[k]) + recv;
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int row = row0 + (2 * t + k) * 8;
                        if (row < p.M) pout[row] = k2[k];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < WNT; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = col0 + j * 8 + e;
                    if (col < p.N) {
                        double* cptr = col < p.nsplit ? p.C + (int64_t)col * p.ldc
                                                      : p.C2 + (int64_t)(col - p.nsplit) * p.ldc2;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = row0 + i * 8;
                            if (row < p.M) {
                                double v = p.alpha * acc[i][j][e];
                                if (row == col) v += p.diag;
                                cptr[row] = v;
                            }
                        }
                    }
                }
            }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

std::mutex g_config_mutex;      // guards the per-device "function attributes set" flags and the driver entry point below

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    std::lock_guard<std::mutex> lock(g_config_mutex);
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        REACH_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
        if (!sym || qres != cudaDriverEntryPointSuccess) throw Error(2, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// Tensor map over the K-contiguous operand: rows x K doubles, row pitch ld; box = 128 rows x 16 k, 128-byte swizzle,
// out-of-range elements read as zero.
CUtensorMap make_tmap_rows(const double* base, int64_t rows, int64_t K, int64_t ld) {
    CUtensorMap m;
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box,
                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d) for rows=%lld K=%lld ld=%lld", (int)r, (long long)rows,
                 (long long)K, (long long)ld);
        throw Error(2, buf);
    }
    return m;
}

template <int WNT, bool AKM>
void launch(Ctx* ctx, const GemmParams& p, const CUtensorMap& tmap) {
    constexpr int BN = 32 * WNT;
    constexpr int A_STAGE = AKM ? BM * BK : BK * PITCH_A;
    constexpr int B_STAGE = BK * (BN + 4) + (AKM ? 8 : 0);
    const size_t smem = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double) + 2 * STAGES * sizeof(uint64_t) +
                        (STAGES + (STAGES & 1)) * sizeof(int) + CONSUMER_WARPS * 2 * 32 * sizeof(double) + 1024;
    static bool configured[64] = {false};      // function attributes are per device
    const int dev = ctx->device & 63;
    constexpr size_t SMEM_ALL = 227 * 1024;    // with the 1 KB the system reserves per CTA: the whole SM, nothing fits beside it
    {
        std::lock_guard<std::mutex> lock(g_config_mutex);     // contexts of several devices may run on their own host threads
        if (!configured[dev]) {
            REACH_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel<WNT, AKM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALL));
            configured[dev] = true;
        }
    }
    const size_t smem_launch = (ctx->gemm_exclusive && ctx->gemm_reserve_sms > 0) ? SMEM_ALL : smem;
    const int tiles = p.tiles_m * p.tiles_n;
    const int sms = ctx->sm_count - ctx->gemm_reserve_sms > 0 ? ctx->sm_count - ctx->gemm_reserve_sms : 1;
    const int grid = tiles < sms ? tiles : sms;
    // scheduler words: a small ring so that two launches in flight never share a pair (each launch leaves its pair zeroed)
    constexpr int SCHED_RING = 64;
    if (!ctx->gemm_sched) {
        REACH_CUDA(cudaMalloc(&ctx->gemm_sched, 2 * SCHED_RING * sizeof(int)));
        ctx->scratch.push_back(ctx->gemm_sched);
        REACH_CUDA(cudaMemset(ctx->gemm_sched, 0, 2 * SCHED_RING * sizeof(int)));
    }
    GemmParams q = p;
    q.sched = ctx->gemm_sched + 2 * (ctx->gemm_sched_next++ % SCHED_RING);
    gemm_dmma_kernel<WNT, AKM><<<grid, THREADS, smem_launch, ctx->stream>>>(q, tmap);
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
}

__global__ void transpose_kernel(int rows, int cols, const double* __restrict__ A, int64_t lda, double* __restrict__ At,
                                 int64_t ldt) {
    __shared__ double tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + threadIdx.x, c = c0 + j;
        tile[j][threadIdx.x] = (r < rows && c < cols) ? A[(int64_t)c * lda + r] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + threadIdx.x, r = r0 + j;     // At[c, r] = A[r, c]
        if (r < rows && c < cols) At[(int64_t)r * ldt + c] = tile[threadIdx.x][j];
    }
}

}  // namespace

// widest column tile (32 * WNT columns) that wastes the least padding
static int nt_col_warp_tiles(int64_t N) {
    int best = 4;
    int64_t best_cols = round_up(N, 128);
    for (int wnt = 3; wnt >= 2; --wnt) {
        int64_t c = round_up(N, 32 * wnt);
        if (c < best_cols) { best_cols = c; best = wnt; }
    }
    return best;
}

int64_t gemm_nt_col_groups(int64_t N) { return 4 * ceil_div(N, 32 * nt_col_warp_tiles(N)); }

void gemm_nt(Ctx* ctx, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* Bt, int64_t ldbt,
             double* C, int64_t ldc, const GemmEpilogue& ep) {
    if (M <= 0 || N <= 0) return;
    REACH_REQUIRE(K > 0, "gemm_nt: K must be positive");
    REACH_REQUIRE((lda % 2) == 0 && (ldbt % 2) == 0, "gemm_nt: leading dimensions must be even (16-byte TMA rows)");
    REACH_REQUIRE((((uintptr_t)A) & 15) == 0 && (((uintptr_t)Bt) & 15) == 0, "gemm_nt: operands must be 16-byte aligned");
    GemmParams p;
    p.A = A; p.Bt = Bt; p.C = C;
    p.lda = lda; p.ldbt = ldbt; p.ldc = ldc;
    p.M = (int)M; p.N = (int)N; p.ksteps = (int)ceil_div(K, BK);
    p.alpha = ep.alpha; p.diag = ep.diag;
    p.C2 = ep.C2; p.ldc2 = ep.ldc2;
    p.nsplit = (ep.C2 != nullptr && ep.nsplit >= 0) ? (int)ep.nsplit : (int)N;
    p.w0 = ep.part ? ep.w0 : nullptr; p.w1 = ep.part ? ep.w1 : nullptr; p.part = ep.w0 ? ep.part : nullptr; p.ldpart = ep.ldpart;
    REACH_REQUIRE(!p.part || (p.ldpart >= M && ep.alpha == 1.0), "gemm_nt: fused abs-row-sums need ldpart >= M and alpha == 1");
    p.tiles_m = (int)ceil_div(M, BM);
    const int best = nt_col_warp_tiles(N);
    p.tiles_n = (int)ceil_div(N, 32 * best);
    CUtensorMap dummy;
    memset(&dummy, 0, sizeof(dummy));
    switch (best) {
        case 4: launch<4, false>(ctx, p, dummy); break;
        case 3: launch<3, false>(ctx, p, dummy); break;
        default: launch<2, false>(ctx, p, dummy); break;
    }
}

// Y[n x Q] = P[n x K] * X[K x Q], everything column-major (see the AKM = true variant above).
void gemm_left(Ctx* ctx, int64_t n, int64_t Q, int64_t K, const double* P, int64_t ldp, const double* X, int64_t ldx,
               double* Y, int64_t ldy, const GemmEpilogue& ep) {
    if (n <= 0 || Q <= 0) return;
    REACH_REQUIRE(K > 0, "gemm_left: K must be positive");
    // The K-tail of a generator row (k in [K, round_up(K,16))) may run into the next column: it multiplies the zero
    // pad rows of P, so it only has to be finite; the caller keeps >= 128 zeroed columns of slack behind X.
    REACH_REQUIRE((ldp % 2) == 0 && (ldx % 2) == 0 && (ldy % 2) == 0 && ldx >= K && ldy >= round_up(n, 2),
                  "gemm_left: leading dimensions must be even, ldx >= K, ldy >= round_up(n,2)");
    REACH_REQUIRE((((uintptr_t)P) & 15) == 0 && (((uintptr_t)X) & 15) == 0 && (((uintptr_t)Y) & 15) == 0,
                  "gemm_left: operands must be 16-byte aligned");
    GemmParams p;
    p.A = X; p.Bt = P; p.C = Y;
    p.lda = ldx; p.ldbt = ldp; p.ldc = ldy;
    p.M = (int)Q; p.N = (int)n; p.ksteps = (int)ceil_div(K, BK);
    p.alpha = ep.alpha; p.diag = ep.diag;
    p.C2 = nullptr; p.ldc2 = 0; p.nsplit = (int)n;
    p.w0 = p.w1 = nullptr; p.part = nullptr; p.ldpart = 0;
    p.tiles_m = (int)ceil_div(Q, BM);
    int best = 4;
    int64_t best_cols = round_up(n, 128);
    for (int wnt = 3; wnt >= 2; --wnt) {
        int64_t c = round_up(n, 32 * wnt);
        if (c < best_cols) { best_cols = c; best = wnt; }
    }
    p.tiles_n = (int)ceil_div(n, 32 * best);
    const CUtensorMap tmap = make_tmap_rows(X, Q, K, ldx);
    switch (best) {
        case 4: launch<4, true>(ctx, p, tmap); break;
        case 3: launch<3, true>(ctx, p, tmap); break;
        default: launch<2, true>(ctx, p, tmap); break;
    }
}

void transpose(Ctx* ctx, int64_t rows, int64_t cols, const double* A, int64_t lda, double* At, int64_t ldt) {
    if (rows <= 0 || cols <= 0) return;
    dim3 grid((unsigned)ceil_div(rows, 32), (unsigned)ceil_div(cols, 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>((int)rows, (int)cols, A, lda, At, ldt);
    REACH_CUDA(cudaGetLastError());
    ctx->launches++;
}

}  // namespace reach
