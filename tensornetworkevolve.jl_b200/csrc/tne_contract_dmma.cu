// FP64 DMMA fast paths of the contraction (see tne_contract.cu).  Filled in below.
#include "tne_gett.cuh"

int gett_try_fast(tne_ctx*, const GettDesc&, const cplx*, const cplx*, cplx*, bool* handled) {
    *handled = false;
    return TNE_OK;
}
