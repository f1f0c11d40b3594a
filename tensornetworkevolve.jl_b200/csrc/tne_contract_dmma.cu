// FP64 DMMA fast paths of the contraction (see tne_contract.cu).  Filled in below.
#include "tne_gett.cuh"

int gett_fast_plan(tne_ctx*, ContractPlan*) { return TNE_OK; }
int gett_fast_run(tne_ctx* ctx, const ContractPlan&, const cplx*, const cplx*, cplx*) {
    return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "no fast path built");
}
