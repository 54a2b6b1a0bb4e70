/* libreach_sm100a -- C ABI of the B200-native linear-flowpipe hot path of JuliaReach/Reachability.jl v0.7.0.
 *
 * The reference has no FFI of its own (it is pure Julia); each entry point below replaces one Julia-level
 * seam on the hot path and is what the `ccall` shim (julia/ReachSM100a.jl) binds.
 * Citations are file:line under /root/reference/src.
 *
 * Conventions
 *   - every function returns int32 status: 0 = OK, 1 = invalid argument, 2 = CUDA/runtime error,
 *     3 = unsupported on this path (the caller must raise; there is NO CPU fallback in this library);
 *     the message is available through reach_last_error().
 *   - all host pointers are caller-owned, column-major Float64 (Julia layout) with an explicit leading
 *     dimension, valid only for the duration of the call; the library copies what it needs to the device
 *     before returning and never retains host pointers.
 *   - indices crossing the ABI are 0-based int32; sizes are int64 (Julia Int).
 *   - reach_ctx_create binds a context to one CUDA device and one stream; multi-GPU runs either use one process (rank)
 *     and one context per GPU and shard rows / steps / problems across ranks, or ONE process with a multi-device
 *     context (reach_ctx_create_multi) whose BFFPSV18 calls shard the rows inside the library (DESIGN.md "Multi-GPU").
 */
#ifndef REACH_SM100A_H
#define REACH_SM100A_H

#include <stdint.h>

#if defined(__GNUC__)
#define REACH_API __attribute__((visibility("default")))
#else
#define REACH_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define REACH_OK 0
#define REACH_EINVAL 1
#define REACH_ECUDA 2
#define REACH_EUNSUPPORTED 3

#define REACH_FORWARD 0  /* discretize(...; algorithm="forward")  discretize.jl:654-655 */
#define REACH_BACKWARD 1 /* discretize(...; algorithm="backward") discretize.jl:656-657 */
#define REACH_FIRSTORDER 2 /* discretize_firstorder, p = Inf      discretize.jl:403-473 */
#define REACH_NOBLOATING 3 /* discretize_nobloating               discretize.jl:527-552 */

/* zonotope flags (discretisation and GLGM06) */
#define REACH_GLGM06_KEEP_ZERO_GENERATORS 1 /* Zonotope(c, G; remove_zero_generators=false) semantics */
#define REACH_GLGM06_RECORD_SELECTIONS 2    /* keep sortperm + scores of every reduce_order call (reach_flowpipe_selection) */

typedef struct reach_ctx reach_ctx;
typedef struct reach_boxplan reach_boxplan;   /* BFFPSV18 problem resident on the device */
typedef struct reach_flowpipe reach_flowpipe; /* GLGM06 problem + zonotope flowpipe resident on the device */
typedef struct reach_ensemble reach_ensemble; /* batched homogeneous GLGM06 flowpipes on the device */

/* ---- context ----------------------------------------------------------------------------------------- */
REACH_API int32_t reach_version(void);
/* Module `__init__` of the Julia shim. Fails (REACH_ECUDA) when no sm_100 device is visible. */
REACH_API int32_t reach_ctx_create(int32_t device, reach_ctx** out);
/* Same, but all work is issued on the caller's CUDA stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream
 * or CUDA.jl's task stream); the stream is borrowed, not owned.  stream == NULL behaves like reach_ctx_create. */
REACH_API int32_t reach_ctx_create_on_stream(int32_t device, void* stream, reach_ctx** out);
/* One handle over SEVERAL devices of one process: devices[0] is the primary device -- every entry point
 * works on it exactly as with reach_ctx_create -- and reach_bffpsv18_boxes / reach_bffpsv18_boxes_csc additionally shard
 * their rows of interest in contiguous ranges over all ndev devices (Phi and the sets replicated, one host thread per
 * device, no collective: P_k[R_g, :] = P_{k-1}[R_g, :] Phi needs only the shard's own rows; the results are gathered by the
 * strided download into the caller's arrays).  This is how a single Julia task (`solve`, src/solve.jl:59-81) gets all GPUs of
 * a box through one `ccall`.  A device may be listed more than once (the shards then share it). */
REACH_API int32_t reach_ctx_create_multi(const int32_t* devices, int32_t ndev, reach_ctx** out);
REACH_API int32_t reach_ctx_device_count(reach_ctx* ctx, int32_t* ndev);
REACH_API void reach_ctx_destroy(reach_ctx* ctx);
/* Message of the last failure on this context (ctx == NULL: last failure of reach_ctx_create). */
REACH_API const char* reach_last_error(const reach_ctx* ctx);
REACH_API int32_t reach_sync(reach_ctx* ctx);
/* Counters for `@timing` (logging.jl:80-88) and bench.py: kernels launched so far on this context. */
REACH_API int32_t reach_launch_count(reach_ctx* ctx, uint64_t* launches);

/* ---- discretisation: src/ReachSets/discretize.jl ----------------------------------------------------- */
/* exp_Aδ(A, δ; exp_method="base") = exp(Matrix(A*δ))            discretize.jl:150-152 */
REACH_API int32_t reach_expm(reach_ctx* ctx, int64_t n, const double* A, int64_t lda, double delta, double* Phi, int64_t ldphi);
/* Φ₁(A, δ) = exp(Pmatrix)[1:n, n+1:2n]   discretize.jl:221-226;  Φ₂(A, δ) = exp(Pmatrix)[1:n, 2n+1:3n]  :296-301.
 * Computed from the block-triangular structure without forming the 3n x 3n exponential. Either output may be NULL. */
REACH_API int32_t reach_phi12(reach_ctx* ctx, int64_t n, const double* A, int64_t lda, double delta, double* Phi1, int64_t ld1,
                    double* Phi2, int64_t ld2);
/* discretize_interpolation (discretize.jl:627-675) for a box X0 = (c, r) and a constant box input U = (cU, rU)
 * entering through B (n x m; m = 0: homogeneous; B == NULL with m == n: identity), followed by the box
 * decomposition of the lazy Ω0 that BFFPSV18 performs (BFFPSV18/reach.jl:64-65).
 * Outputs (each may be NULL): Phi n x n; Phi2abs = Φ₂(|A|, δ) n x n; c0, r0 (n) box of Ω0; rE (n) radius of Einit;
 * rpsi (n) radius of Eψ0; MU = δ·B (n x m, ld n).
 * algorithm REACH_FIRSTORDER: the first-order model with infinity-norm balls (rE = α, rpsi = β uniform; Phi2abs is not
 * computed); REACH_NOBLOATING: Ω0 = X0, MU = Φ₁(A, δ)·B, rpsi = 0. */
REACH_API int32_t reach_discretize_box(reach_ctx* ctx, int64_t n, const double* A, int64_t lda, double delta, int32_t algorithm,
                             const double* c, const double* r, int64_t m, const double* B, int64_t ldb,
                             const double* cU, const double* rU, double* Phi, int64_t ldphi, double* Phi2abs,
                             int64_t ldphi2, double* c0, double* r0, double* rE, double* rpsi, double* MU);

/* discretize(...; set_operations = "zonotope") for a box X0 = (c, r) and a constant box input U = (cU, rU) through B
 * (discretize.jl:705-743 with `_convert_or_overapproximate`, :882-890) -- what GLGM06 calls (GLGM06/post.jl:15-18):
 *   homogeneous (m == 0):  Z2 = linear_map(Φ, Z1 ⊕ Einit),  Ω0 = overapproximate(ConvexHull(Z1, Z2), Zonotope)      :706-709
 *   inhomogeneous:         Z2 = ΦZ1 ⊕ δU0 ⊕ Eψ0 ⊕ Einit,  Ω0 as above,  V = δU0 ⊕ Eψ0                               :715-727
 * computed on the device, zero-generator removal of the `Zonotope` constructor included (flags:
 * REACH_GLGM06_KEEP_ZERO_GENERATORS keeps them).  algorithm: REACH_FORWARD / REACH_BACKWARD; REACH_FIRSTORDER and
 * REACH_NOBLOATING return REACH_EINVAL like the reference's ArgumentError (discretize.jl:105-115).
 * Outputs: Phi n x n (may be NULL); Ω0 = (c0, G0) with G0 n x p0, the caller provides room for 4n + m + 1 columns;
 * V = (cV, GV) with GV n x pV, room for n + m columns (ignored when m == 0; *pV = 0). */
REACH_API int32_t reach_discretize_zonotope(reach_ctx* ctx, int64_t n, const double* A, int64_t lda, double delta,
                                    int32_t algorithm, const double* c, const double* r, int64_t m, const double* B,
                                    int64_t ldb, const double* cU, const double* rU, int32_t flags, double* Phi, int64_t ldphi,
                                    double* c0, double* G0, int64_t ldg0, int64_t* p0, double* cV, double* GV, int64_t ldgv,
                                    int64_t* pV);

/* Time-varying inputs (`VaryingInput`; discretize.jl:690-696 with lazy sets, :729-737 with zonotopes): for the K input boxes
 * U_k = B * Box(cUs[:, k], rUs[:, k]) (B n x m, NULL with m == n: identity) the bloating terms
 *   Eψ_k = sih(Φ₂(|A|, δ) * sih(A * U_k)),   radius rpsi[:, k] = Φ₂(|A|, δ) (|A B cU_k| + |A B| rU_k)      (n x K, ld ldrpsi),
 * for all k at once (three thin DMMA GEMMs).  MU (may be NULL) receives δ B; the discretised input of step k is
 * Ud[k] = MU * Box(cU_k, rU_k) ⊕ Box(0, rpsi[:, k]).  (The reference's reach kernels themselves take a ConstantInput only,
 * reach_blocks.jl:137, GLGM06/reach.jl:32.) */
REACH_API int32_t reach_discretize_inputs_box(reach_ctx* ctx, int64_t n, const double* A, int64_t lda, double delta, int64_t m,
                                      const double* B, int64_t ldb, int64_t K, const double* cUs, int64_t ldcu,
                                      const double* rUs, int64_t ldru, double* rpsi, int64_t ldrpsi, double* MU);

/* ---- BFFPSV18: dense reach_blocks! (BFFPSV18/reach_blocks.jl:135-224) ---------------------------------- */
/* One-shot host-buffer call = the method the shim adds for
 *   reach_blocks!(ϕ::Matrix{Float64}, Xhat0::Vector{<:Union{Interval,Hyperrectangle}}, U, ..., res)
 * dispatched from the registry at BFFPSV18/reach.jl:191-193.
 *   Phi n x n; (c0, r0) decomposed Ω0 (all n coordinates, reach_blocks.jl:185-187);
 *   inputs: MU = δB (n x m), U = box(cU, rU), rpsi (n) -- m == 0 (or MU == NULL): homogeneous (U == nothing);
 *   rows: the nrows 0-based coordinates of the blocks in `:blocks` (compute_dimensions.jl:1-15);
 *   N = number of reach sets (reach.jl:54).
 * Outputs centers/radii: nrows x N column-major (column k-1 = step k; reach_blocks.jl:152-158,195-202).
 * steps_done receives N (the `index` of the (index, skip) tuple, reach_blocks.jl:62; skip is always false for the
 * Universe invariant, reach.jl:225-227). */
REACH_API int32_t reach_bffpsv18_boxes(reach_ctx* ctx, int64_t n, const double* Phi, int64_t ldphi, const double* c0,
                             const double* r0, int64_t m, const double* MU, int64_t ldmu, const double* cU,
                             const double* rU, const double* rpsi, int64_t nrows, const int32_t* rows, int64_t N,
                             double* centers, double* radii, int64_t* steps_done);
/* The sparse twin, reach_blocks!(ϕ::SparseMatrixCSC{Float64,Int}, ...) (reach_blocks.jl:47-132; `:assume_sparse`): Φ crosses the
 * ABI as its CSC fields (colptr n+1, rowval nnz, nzval nnz; 1-based Int64).  The sparse method computes the same sets as the
 * dense one -- it only skips the blocks of Φ^k that are zero.  Two device paths:
 *  (a) the sparsity graph of Φ falls into connected components of at most 16 states (decoupled subsystems, block-diagonal up
 *      to a permutation): every power of Φ keeps that pattern, so row l of Φ^k lives on the columns of its component and the
 *      zero blocks are skipped for good -- one thread per row of interest runs the whole recurrence, w^2 fma per step instead
 *      of 2 n^2 (same sets as the dense path to rounding: the sums run over the same non-zero terms in ascending order);
 *  (b) anything else: Φ^k fills in, the device scatters Φ into its dense layout and runs the DMMA recurrence; results are
 *      bit-identical to reach_bffpsv18_boxes on Matrix(Φ).
 * reach_bffpsv18_csc_info reports which path served the last call on this context (largest component, 0 = path (b)). */
REACH_API int32_t reach_bffpsv18_boxes_csc(reach_ctx* ctx, int64_t n, int64_t nnz, const int64_t* colptr, const int64_t* rowval,
                                   const double* nzval, const double* c0, const double* r0, int64_t m, const double* MU,
                                   int64_t ldmu, const double* cU, const double* rU, const double* rpsi, int64_t nrows,
                                   const int32_t* rows, int64_t N, double* centers, double* radii, int64_t* steps_done);
REACH_API int32_t reach_bffpsv18_csc_info(reach_ctx* ctx, int64_t* max_component, int32_t* sparse_path);
/* Same computation split into upload / device-resident run / download (what bench.py times as `value`). */
REACH_API int32_t reach_boxplan_create(reach_ctx* ctx, int64_t n, const double* Phi, int64_t ldphi, const double* c0,
                             const double* r0, int64_t m, const double* MU, int64_t ldmu, const double* cU,
                             const double* rU, const double* rpsi, int64_t nrows, const int32_t* rows, int64_t N,
                             reach_boxplan** out);
REACH_API int32_t reach_boxplan_run(reach_boxplan* plan);                      /* whole time loop on the device */
REACH_API int32_t reach_boxplan_fetch(reach_boxplan* plan, double* centers, double* radii); /* nrows x N each */
/* Device pointers of the result (nrows x N column-major, ld = nrows), for zero-copy consumers. */
REACH_API int32_t reach_boxplan_device_results(reach_boxplan* plan, double** centers_dev, double** radii_dev);
/* CUDA-event time of the last run and of its dominant kernel (the stacked DMMA GEMM), launches of each. */
REACH_API int32_t reach_boxplan_timing(reach_boxplan* plan, double* run_ms, double* gemm_ms, int64_t* gemm_launches,
                             double* gemm_flops);
REACH_API void reach_boxplan_destroy(reach_boxplan* plan);
/* check mode, dense check_blocks (BFFPSV18/check_blocks.jl:105-183) for a half-space property
 * is_contained_in(LinearConstraint(ell, bound)) already mapped to `rows` (partitions.jl:33-46).
 * The states stay lazy per block (check_blocks.jl:146-156), so the support function is separable over the blocks of
 * `:blocks`: block_sizes[0..nblocks) groups `rows` into those blocks, in order (NULL: every row is its own block).
 * eager != 0: stop at the first violation (`:eager_checking`).
 * violation: first violating step (1-based) or 0 (check_blocks.jl:116). */
REACH_API int32_t reach_bffpsv18_check(reach_ctx* ctx, int64_t n, const double* Phi, int64_t ldphi, const double* c0,
                             const double* r0, int64_t m, const double* MU, int64_t ldmu, const double* cU,
                             const double* rU, const double* rpsi, int64_t nrows, const int32_t* rows, int64_t N,
                             int64_t nblocks, const int32_t* block_sizes, const double* ell, double bound, int32_t eager,
                             int64_t* violation);
/* output_function branch (reach_blocks.jl:152-155,197-199; reach.jl:73-82): stored set is
 * box_approximation(Mout * CPA(Xhat_k)); Mout is q x nrows (any q; the kernels carry 8 output rows at a time), block_sizes as
 * above. centers/radii: q x N. */
REACH_API int32_t reach_bffpsv18_output_map(reach_ctx* ctx, int64_t n, const double* Phi, int64_t ldphi, const double* c0,
                                  const double* r0, int64_t m, const double* MU, int64_t ldmu, const double* cU,
                                  const double* rU, const double* rpsi, int64_t nrows, const int32_t* rows, int64_t N,
                                  int64_t nblocks, const int32_t* block_sizes, int64_t q, const double* Mout, int64_t ldm,
                                  double* centers, double* radii);

/* ---- sparse A, lazy matrix exponential: `:lazy_expm` / `exp_method = "lazy"` --------------------------------- */
/* The state matrix as Julia's SparseMatrixCSC fields (colptr n+1, rowval nnz, nzval nnz; 1-based Int64), kept on the device
 * as CSR of Aᵀ (the same arrays) and CSR of A.  Φ = e^{Aδ} is never formed on this path: every use of Φ is the ACTION
 * of e^{δA} / e^{δAᵀ} on a block of vectors, a scaled Taylor series of sparse x dense-block products carried to full
 * double precision -- the device counterpart of `SparseMatrixExp` + `get_rows` / `get_columns` / `expmv`
 * (discretize.jl:153-154, 302-306; BFFPSV18/reach_blocks.jl:37-38). */
typedef struct reach_sparse reach_sparse;
REACH_API int32_t reach_sparse_create(reach_ctx* ctx, int64_t n, int64_t nnz, const int64_t* colptr, const int64_t* rowval,
                              const double* nzval, reach_sparse** out);
REACH_API void reach_sparse_free(reach_sparse* sp);
/* Y = e^{δA} X (transpose_op = 0: what get_columns needs) or e^{δAᵀ} X (1: get_rows); X, Y n x R column-major. */
REACH_API int32_t reach_expmv(reach_sparse* sp, double delta, int32_t transpose_op, int64_t R, const double* X, int64_t ldx,
                      double* Y, int64_t ldy);
/* reach_discretize_box for a sparse A with REACH_FORWARD / REACH_BACKWARD (others: REACH_EUNSUPPORTED): same outputs except
 * the matrices -- c0, r0 (box of Ω0), rE, rpsi (n), MU = δB (n x m, ld n); each may be NULL. */
REACH_API int32_t reach_discretize_box_sparse(reach_sparse* sp, double delta, int32_t algorithm, const double* c, const double* r,
                                      int64_t m, const double* B, int64_t ldb, const double* cU, const double* rU,
                                      double* c0, double* r0, double* rE, double* rpsi, double* MU);
/* reach_blocks!(ϕ::SparseMatrixExp, ...) for Interval / Hyperrectangle blocks (BFFPSV18/reach_blocks.jl:227-397): the rows
 * of interest of e^{kδA} are advanced on the device, W_kᵀ = e^{δAᵀ} W_{k-1}ᵀ (the reference recomputes them with a Krylov
 * expmv from ϕpowerk.M = kδA at every step, :306).  Arguments and outputs as reach_bffpsv18_boxes, without Phi. */
REACH_API int32_t reach_bffpsv18_boxes_lazy(reach_sparse* sp, double delta, const double* c0, const double* r0, int64_t m,
                                    const double* MU, int64_t ldmu, const double* cU, const double* rU, const double* rpsi,
                                    int64_t nrows, const int32_t* rows, int64_t N, double* centers, double* radii,
                                    int64_t* steps_done);
/* The output_function branch and check mode with a lazy exponential (reach_blocks.jl:227-397 with output_function;
 * check_blocks.jl for a SparseMatrixExp): arguments as reach_bffpsv18_output_map / reach_bffpsv18_check.  The per-block
 * functionals Σ_{l in b_i} M[a,l] e^{kδA}[rows[l], :] are advanced as extra columns of the same recurrence. */
REACH_API int32_t reach_bffpsv18_output_map_lazy(reach_sparse* sp, double delta, const double* c0, const double* r0, int64_t m,
                                         const double* MU, int64_t ldmu, const double* cU, const double* rU, const double* rpsi,
                                         int64_t nrows, const int32_t* rows, int64_t N, int64_t nblocks,
                                         const int32_t* block_sizes, int64_t q, const double* Mout, int64_t ldm, double* centers,
                                         double* radii);
REACH_API int32_t reach_bffpsv18_check_lazy(reach_sparse* sp, double delta, const double* c0, const double* r0, int64_t m,
                                    const double* MU, int64_t ldmu, const double* cU, const double* rU, const double* rpsi,
                                    int64_t nrows, const int32_t* rows, int64_t N, int64_t nblocks, const int32_t* block_sizes,
                                    const double* ell, double bound, int32_t eager, int64_t* violation);
/* CUDA-event time of the last call on `sp`, its sparse x block products and their algorithmic HBM/L2 bytes. */
REACH_API int32_t reach_sparse_timing(reach_sparse* sp, double* run_ms, int64_t* products, double* bytes);

/* ---- GLGM06: reach_homog! / reach_inhomog! (GLGM06/reach.jl:5-62) -------------------------------------- */
/* flags: REACH_GLGM06_KEEP_ZERO_GENERATORS, REACH_GLGM06_RECORD_SELECTIONS (defined above) */
/* Upload Ω0red = (c0, G0 n x p0) (already reduce_order'ed, GLGM06/post.jl:20 -- or call reach_reduce_order) and
 * the discretised input V = (cV, GV n x pV); pV == 0 and cV == NULL selects reach_homog!. */
REACH_API int32_t reach_glgm06_create(reach_ctx* ctx, int64_t n, const double* Phi, int64_t ldphi, int64_t p0, const double* c0,
                            const double* G0, int64_t ldg0, int64_t pV, const double* cV, const double* GV,
                            int64_t ldgv, int64_t N, int64_t max_order, int32_t flags, reach_flowpipe** out);
REACH_API int32_t reach_glgm06_run(reach_flowpipe* fp); /* whole recurrence on the device; results stay there */
/* Column-sharded single system over `nranks` GPUs (one process and context per GPU): rank g keeps and maps only the
 * generators [p0 g/nranks, p0 (g+1)/nranks) of Omega0 and the same slice of V; the order-reduction SELECTION runs replicated
 * on the all-reduced scores, so every rank takes identical decisions (reduce_order, GLGM06/reach.jl:49,55).  The caller
 * drives the batches in order: for every batch b in [0, nbatches) and phase in 0..3 call reach_glgm06_shard_phase, then
 * sum-all-reduce (NCCL over NVLink) the buffers reach_glgm06_shard_buffer reports: after phase 0 `which` = 0 (scores,
 * float64) and 1 (row-support masks, int32); after phase 1 `which` = 2 (W box partials); after phase 2 `which` = 3 (R box
 * partials).  Afterwards reach_flowpipe_step returns this rank's columns of R_k (the others are zero): the slabs of the
 * ranks have disjoint support and their sum is the reach set; centres are replicated. */
REACH_API int32_t reach_glgm06_create_sharded(reach_ctx* ctx, int64_t n, const double* Phi, int64_t ldphi, int64_t p0, const double* c0,
                                    const double* G0, int64_t ldg0, int64_t pV, const double* cV, const double* GV,
                                    int64_t ldgv, int64_t N, int64_t max_order, int32_t flags, int32_t rank, int32_t nranks,
                                    reach_flowpipe** out);
REACH_API int32_t reach_glgm06_shard_nbatches(reach_flowpipe* fp, int64_t* nbatches);
REACH_API int32_t reach_glgm06_shard_phase(reach_flowpipe* fp, int64_t batch, int32_t phase);
REACH_API int32_t reach_glgm06_shard_buffer(reach_flowpipe* fp, int64_t batch, int32_t which, void** dev_ptr, int64_t* count,
                                  int32_t* dtype /* 0 = float64, 1 = int32 */);
/* reach_glgm06_shard_phase with the phase's kernels issued on `cuda_stream` (a cudaStream_t of the caller) instead of the
 * context's stream.  Phase 0 of batch b + 1 -- the map of the rank's own generator columns (the DMMA GEMM when Phi is dense)
 * and their scores -- needs nothing of batch b but its archive slots, so the caller may issue it on a second stream BEFORE
 * phases 1..3 of batch b and the two overlap on the device (sharding.run_sharded(..., gemm_stream=...)): the batch-local
 * arrays exist three times, the library orders their reuse with events, and the buffers reach_glgm06_shard_buffer reports
 * are those of the batch's own set.  Phase 0 may run at most two batches ahead of phases 1..3 (status 1 otherwise); the
 * all-reduces of a phase must be ordered after it on the stream it was issued on. */
REACH_API int32_t reach_glgm06_shard_phase_on(reach_flowpipe* fp, int64_t batch, int32_t phase, void* cuda_stream);
/* run + fetch in one call, streamed: every batch of reach sets is copied to the host (layout of reach_flowpipe_fetch;
 * pinned buffers overlap best) while the later batches are still being computed.  This is the shape of
 * `reach_inhomog!(R, ...)` filling its caller-allocated vector (GLGM06/reach.jl:30-36). */
REACH_API int32_t reach_glgm06_run_fetch(reach_flowpipe* fp, double* centers, double* G, int64_t pmax);
/* Number of generators of every reach set R_1..R_N. */
REACH_API int32_t reach_flowpipe_ngens(reach_flowpipe* fp, int32_t* p /* N */);
/* Reach set k (1-based): centre (n) and generators (n x p_k, leading dimension ldg >= n). */
REACH_API int32_t reach_flowpipe_step(reach_flowpipe* fp, int64_t k, double* c, double* G, int64_t ldg);
/* All reach sets at once (the `R` vector reach_inhomog! fills, GLGM06/reach.jl:30-36): centers n x N (ld n);
 * G holds N slabs of n x pmax (ld n, slab stride n*pmax doubles), slab k-1 = generators of R_k in its first p_k columns.
 * Either output may be NULL. */
REACH_API int32_t reach_flowpipe_fetch(reach_flowpipe* fp, double* centers, double* G, int64_t pmax);
/* Box hull of reach set k: lo/hi (n). */
REACH_API int32_t reach_flowpipe_box(reach_flowpipe* fp, int64_t k, double* lo, double* hi);
/* project(Rsets, options) for a zonotope flowpipe (project_reach.jl:18-99; `project(rs, M)` of ReachSet.jl): every reach
 * set is mapped through the q x n projection matrix M (1 <= q <= 4, column-major, ld ldm) ON THE DEVICE, so that only the
 * projected sets cross PCIe (C2: 86 MB instead of 11.7 GB).  centers, radii (q x N): the Hyperrectangle overapproximation
 * of M * R_k, `set_type_proj = Hyperrectangle` (default_options.jl:69): centre M c_k, radius sum_j |M g_j|.
 * Gp (may be NULL): the projected zonotopes themselves, N slabs of q x pmax (ld q), slab k-1 = M * G_k in its first p_k
 * columns (what `ε_proj < Inf` / HPolygon consumers need).  For a column-sharded flowpipe the radii (and Gp columns) are this
 * rank's partial results. */
REACH_API int32_t reach_flowpipe_project(reach_flowpipe* fp, int64_t q, const double* M, int64_t ldm, double* centers,
                                 double* radii, double* Gp, int64_t pmax);
/* Order-reduction record of step k >= 2 (which = 0: reduce_order of R_k, GLGM06/reach.jl:49; which = 1: of W, :55):
 * p_in columns went in, m were boxed; ind[0..p_in) = sortperm(h) 0-based; h[0..p_in) the scores. Outputs may be NULL.
 * m == 0 and p_in == 0 mean "no reduction happened" (r*n >= p).  Needs REACH_GLGM06_RECORD_SELECTIONS.
 * Indices refer to the concatenation after zero-generator removal, exactly like LazySets' `ind`. */
REACH_API int32_t reach_flowpipe_selection(reach_flowpipe* fp, int64_t k, int32_t which, int32_t* p_in, int32_t* m, int32_t* ind,
                                 double* h);
REACH_API int32_t reach_flowpipe_timing(reach_flowpipe* fp, double* run_ms, double* gemm_ms, int64_t* gemm_launches,
                              double* gemm_flops, double* reduce_ms, double* reduce_bytes);
/* CUDA-event time of the last run spent in the selection chain (side stream) and in the batched numerics (second side
 * stream); the two overlap each other and the GEMMs. */
REACH_API int32_t reach_flowpipe_phase_ms(reach_flowpipe* fp, double* chain_ms, double* numerics_ms);
/* How linear_map(Phi^w, .) runs for this flowpipe: sparse_levels > 0 means Phi and the powers the run uses are structurally
 * sparse (at most nnz_per_row <= 8 non-zeros per row, e.g. decoupled modes) and the generator map is an ELL sparse product
 * (HBM-bound) instead of the dense DMMA GEMM; the results are those of the dense product (the skipped terms are exact
 * zeros).  Detected at creation; the environment variable REACH_GLGM06_DENSE=1 forces the GEMM. */
REACH_API int32_t reach_flowpipe_map_info(reach_flowpipe* fp, int32_t* sparse_levels, int32_t* nnz_per_row);
/* How the Girard SELECTION of W ran in the last run: `batches` batches of steps, `parallel_batches` of them through the parallel
 * selection chain (steady state: W full, no zero scores; every generator's fate over the batch is evaluated independently),
 * the others through the sequential merge chain (growth phase, zero-generator removal still changing the box).  Same
 * selections either way; REACH_CHAIN_SEQUENTIAL=1 forces the sequential chain. */
REACH_API int32_t reach_flowpipe_chain_info(reach_flowpipe* fp, int64_t* parallel_batches, int64_t* batches);
/* Diagnostics: serial != 0 makes the next runs issue everything on one stream (no overlap), so that the per-kernel
 * CUDA-event times of reach_flowpipe_timing / _phase_ms are those of each kernel running alone. */
REACH_API int32_t reach_flowpipe_set_serial(reach_flowpipe* fp, int32_t serial);
REACH_API void reach_flowpipe_free(reach_flowpipe* fp);
/* reduce_order(Z, r) of LazySets (Girard), call site GLGM06/post.jl:20.  Gout must hold n x max(p, r*n) doubles;
 * p_out receives the new generator count; ind (p, may be NULL) the sortperm. */
REACH_API int32_t reach_reduce_order(reach_ctx* ctx, int64_t n, int64_t p, const double* c, const double* G, int64_t ldg,
                           int64_t r, int32_t flags, double* Gout, int64_t ldgo, int64_t* p_out, int32_t* ind);

/* Batched homogeneous ensemble (config C5): `batch` independent initial zonotopes (c0s n x batch, G0s n x p x batch)
 * under the same Φ; R_k = (Φ c_{k-1}, Φ G_{k-1}) for every member (GLGM06/reach.jl:15-19). */
REACH_API int32_t reach_ensemble_create(reach_ctx* ctx, int64_t n, const double* Phi, int64_t ldphi, int64_t batch, int64_t p,
                              const double* c0s, const double* G0s, int64_t N, reach_ensemble** out);
REACH_API int32_t reach_ensemble_run(reach_ensemble* e);
/* Member b (0-based), step k (1-based): centre n, generators n x p (ld n). */
REACH_API int32_t reach_ensemble_step(reach_ensemble* e, int64_t b, int64_t k, double* c, double* G);
/* Box hull of every member at step k: lo/hi n x batch. */
REACH_API int32_t reach_ensemble_boxes(reach_ensemble* e, int64_t k, double* lo, double* hi);
/* Box hull of every member at EVERY step in one pass over the flowpipe: lo/hi n x batch x N (index i + n (b + batch (k-1))). */
REACH_API int32_t reach_ensemble_boxes_all(reach_ensemble* e, double* lo, double* hi);
REACH_API int32_t reach_ensemble_timing(reach_ensemble* e, double* run_ms, double* bytes_per_step);
REACH_API void reach_ensemble_free(reach_ensemble* e);

/* ---- test/bench hooks ---------------------------------------------------------------------------------- */
/* C = A * B with the library's DMMA GEMM (all host column-major); reps > 1 re-runs the device kernel and
 * kernel_ms receives the mean CUDA-event time of one launch. */
REACH_API int32_t reach_dgemm(reach_ctx* ctx, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* B,
                    int64_t ldb, double* C, int64_t ldc, int32_t reps, double* kernel_ms);

/* Y (n x Q) = P (n x K) * X (K x Q) with the left-multiplication variant of the same kernel (generator layout,
 * everything column-major; the GEMM behind linear_map, GLGM06/reach.jl:16,48,54). */
REACH_API int32_t reach_dgemm_left(reach_ctx* ctx, int64_t n, int64_t Q, int64_t K, const double* P, int64_t ldp, const double* X,
                         int64_t ldx, double* Y, int64_t ldy, int32_t reps, double* kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* REACH_SM100A_H */
