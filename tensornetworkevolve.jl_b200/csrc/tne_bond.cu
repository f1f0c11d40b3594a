// Bond update (K5): placeholder until the Jacobi SVD kernel lands.
#include "tne_internal.h"

extern "C" int tne_svd_trunc(tne_ctx* ctx, tne_buf, int, double, tne_buf*, tne_buf*, tne_buf*, int32_t*, double*) {
    return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "tne_svd_trunc: not built yet");
}
extern "C" int tne_apply_onbond_batched(tne_ctx* ctx, int, tne_bond_job*) {
    return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "tne_apply_onbond_batched: not built yet");
}
