// Contraction descriptor shared by the GETT kernels (see tne_contract.cu for the scheme).
#pragma once
#include "tne_internal.h"

__device__ __forceinline__ void decomp2(const LegSet& L, long long idx, long long& o0, long long& o1) {
    o0 = 0; o1 = 0;
#pragma unroll 1
    for (int i = 0; i < L.n; ++i) {
        long long q, r;
        const int sh = L.shift[i];
        if (sh >= 0) { r = idx & (long long)(L.size[i] - 1); q = idx >> sh; }
        else if (idx < 0x7fffffffLL) { unsigned u = (unsigned)idx, s = (unsigned)L.size[i]; unsigned uq = u / s; q = uq; r = u - uq * s; }
        else { q = idx / L.size[i]; r = idx - q * L.size[i]; }
        o0 += r * L.s0[i];
        o1 += r * L.s1[i];
        idx = q;
    }
}
__device__ __forceinline__ void decomp3(const LegSet& L, long long idx, long long& o0, long long& o1, long long& o2) {
    o0 = 0; o1 = 0; o2 = 0;
#pragma unroll 1
    for (int i = 0; i < L.n; ++i) {
        long long q, r;
        const int sh = L.shift[i];
        if (sh >= 0) { r = idx & (long long)(L.size[i] - 1); q = idx >> sh; }
        else { q = idx / L.size[i]; r = idx - q * L.size[i]; }
        o0 += r * L.s0[i]; o1 += r * L.s1[i]; o2 += r * L.s2[i];
        idx = q;
    }
}

__device__ __forceinline__ void atomic_add_c(cplx* p, double re, double im) {
    atomicAdd(&p->x, re);
    atomicAdd(&p->y, im);
}


// DMMA fast paths: pick a kernel for the classified contraction (plan->kernel stays 0 if none applies)
int gett_fast_plan(tne_ctx* ctx, ContractPlan* plan);
int gett_fast_run(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C);
bool gett_multi_capable(const ContractPlan& plan);
int gett_fast_run_multi(tne_ctx* ctx, const ContractPlan& plan, const AbsorbMulti& ops, cplx* C);
