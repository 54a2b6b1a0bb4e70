/* A plain C99 client of include/reach.h: what a non-Python, non-Julia caller links against.
 *
 *   gcc -std=c99 -Wall -Wextra -Werror -pedantic -I include tests/c_abi/flowpipe_smoke.c \
 *       -L reachability.jl_b200 -lreach_sm100a -lm -o flowpipe_smoke
 *
 * Runs one BFFPSV18 box flowpipe and one GLGM06 zonotope flowpipe of a diagonal system x' = diag(a) x + u, whose reach sets
 * have closed forms (Phi = diag(e^{a delta}); reach_blocks.jl:135-224, GLGM06/reach.jl:30-62), and checks every number to
 * 1e-10 relative + 1e-12 absolute.  Exit codes: 0 = all checks passed, 2 = a check failed, 3 = no B200 visible (the library
 * has no CPU path and says so). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "reach.h"

#define N_DIM 6
#define N_IN 2
#define N_STEPS 40

static int close_enough(double got, double want) { return fabs(got - want) <= 1e-10 * fabs(want) + 1e-12; }

#define CHECK(call)                                                                                  \
    do {                                                                                             \
        int32_t rc_ = (call);                                                                        \
        if (rc_ != REACH_OK) {                                                                       \
            fprintf(stderr, "%s failed with status %d: %s\n", #call, (int)rc_, reach_last_error(ctx)); \
            return 2;                                                                                \
        }                                                                                            \
    } while (0)

int main(void) {
    reach_ctx* ctx = NULL;
    int32_t rc = reach_ctx_create(0, &ctx);
    if (rc != REACH_OK) {
        fprintf(stderr, "reach_ctx_create: status %d: %s\n", (int)rc, reach_last_error(NULL));
        return 3;
    }
    const double delta = 0.05;
    double a[N_DIM], A[N_DIM * N_DIM], Phi[N_DIM * N_DIM];
    memset(A, 0, sizeof(A));
    for (int i = 0; i < N_DIM; ++i) {
        a[i] = -0.3 * (i + 1);
        A[i + i * N_DIM] = a[i];
    }
    /* exp_Adelta on the device (discretize.jl:150-152) */
    CHECK(reach_expm(ctx, N_DIM, A, N_DIM, delta, Phi, N_DIM));
    for (int i = 0; i < N_DIM; ++i)
        for (int j = 0; j < N_DIM; ++j)
            if (!close_enough(Phi[i + j * N_DIM], i == j ? exp(a[i] * delta) : 0.0)) {
                fprintf(stderr, "Phi[%d,%d] = %.17g\n", i, j, Phi[i + j * N_DIM]);
                return 2;
            }
    /* BFFPSV18: discretised problem given directly (Omega0 box, discretised input V = MU * Box(cU, rU) + Box(0, rpsi)) */
    double c0[N_DIM], r0[N_DIM], rpsi[N_DIM], MU[N_DIM * N_IN], cU[N_IN] = {0.5, -1.0}, rU[N_IN] = {0.25, 0.5};
    memset(MU, 0, sizeof(MU));
    for (int i = 0; i < N_DIM; ++i) {
        c0[i] = 1.0 + 0.1 * i;
        r0[i] = 0.01 * (i + 1);
        rpsi[i] = 1e-3 * (i + 1);
        MU[i + (i % N_IN) * N_DIM] = delta * (i % 2 ? -1.0 : 2.0);
    }
    int32_t rows[4] = {0, 2, 3, 5};
    double centers[4 * N_STEPS], radii[4 * N_STEPS];
    int64_t done = 0;
    CHECK(reach_bffpsv18_boxes(ctx, N_DIM, Phi, N_DIM, c0, r0, N_IN, MU, N_DIM, cU, rU, rpsi, 4, rows, N_STEPS, centers, radii,
                               &done));
    if (done != N_STEPS) return 2;
    for (int l = 0; l < 4; ++l) {
        const int i = rows[l];
        const double p = exp(a[i] * delta), mu = MU[i + (i % N_IN) * N_DIM], u_c = cU[i % N_IN], u_r = rU[i % N_IN];
        double pk = p, cw = mu * u_c, rw = fabs(mu) * u_r + rpsi[i];      /* reach_blocks.jl:172-175 */
        if (!close_enough(centers[l], c0[i]) || !close_enough(radii[l], r0[i])) return 2;
        for (int k = 2; k <= N_STEPS; ++k) {                             /* :185-218 */
            const double ck = pk * c0[i] + cw, rk = fabs(pk) * r0[i] + rw;
            if (!close_enough(centers[l + 4 * (k - 1)], ck) || !close_enough(radii[l + 4 * (k - 1)], rk)) {
                fprintf(stderr, "BFFPSV18 row %d step %d: (%.17g, %.17g) vs (%.17g, %.17g)\n", i, k, centers[l + 4 * (k - 1)],
                        radii[l + 4 * (k - 1)], ck, rk);
                return 2;
            }
            cw += pk * mu * u_c;
            rw += fabs(pk * mu) * u_r + fabs(pk) * rpsi[i];
            pk *= p;
        }
    }
    /* GLGM06, homogeneous: R_k = Phi^(k-1) R_1 (GLGM06/reach.jl:15-19) */
    double G0[N_DIM * 3], cz[N_DIM], Gk[N_DIM * 3];
    for (int i = 0; i < N_DIM; ++i)
        for (int j = 0; j < 3; ++j) G0[i + j * N_DIM] = 0.1 * (i + 1) - 0.07 * j * (i % 3);
    reach_flowpipe* fp = NULL;
    CHECK(reach_glgm06_create(ctx, N_DIM, Phi, N_DIM, 3, c0, G0, N_DIM, 0, NULL, NULL, N_DIM, N_STEPS, 10, 0, &fp));
    CHECK(reach_glgm06_run(fp));
    CHECK(reach_flowpipe_step(fp, N_STEPS, cz, Gk, N_DIM));
    for (int i = 0; i < N_DIM; ++i) {
        double pw = 1.0;
        const double p = exp(a[i] * delta);
        for (int k = 1; k < N_STEPS; ++k) pw *= p;
        if (fabs(cz[i] - pw * c0[i]) > 1e-9 * fabs(pw * c0[i]) + 1e-12) return 2;        /* product order differs: 1e-9 */
        for (int j = 0; j < 3; ++j)
            if (fabs(Gk[i + j * N_DIM] - pw * G0[i + j * N_DIM]) > 1e-9 * fabs(pw * G0[i + j * N_DIM]) + 1e-12) return 2;
    }
    reach_flowpipe_free(fp);
    uint64_t launches = 0;
    CHECK(reach_launch_count(ctx, &launches));
    reach_ctx_destroy(ctx);
    printf("c_abi flowpipe_smoke: OK (reach_version %d, %llu kernel launches)\n", (int)reach_version(), (unsigned long long)launches);
    return launches > 0 ? 0 : 2;
}
