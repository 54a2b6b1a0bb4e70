// Entry points declared in include/reach.h whose kernels are not written yet return REACH_EUNSUPPORTED
// (the library has no CPU fallback).  Each stub disappears when its implementation lands.
#include "common.cuh"
#include "../../include/reach.h"
namespace reach { Ctx* ctx_of(reach_ctx* h); }
#define STUB(ctx, name) do { if (ctx) reach::ctx_of(ctx)->last_error = name ": not implemented in this build"; return REACH_EUNSUPPORTED; } while (0)
extern "C" {
int32_t reach_bffpsv18_check(reach_ctx* ctx, int64_t, const double*, int64_t, const double*, const double*, int64_t, const double*, int64_t, const double*, const double*, const double*, int64_t, const int32_t*, int64_t, const double*, double, int32_t, int64_t*) { STUB(ctx, "reach_bffpsv18_check"); }
int32_t reach_bffpsv18_output_map(reach_ctx* ctx, int64_t, const double*, int64_t, const double*, const double*, int64_t, const double*, int64_t, const double*, const double*, const double*, int64_t, const int32_t*, int64_t, int64_t, const double*, int64_t, double*, double*) { STUB(ctx, "reach_bffpsv18_output_map"); }
}
