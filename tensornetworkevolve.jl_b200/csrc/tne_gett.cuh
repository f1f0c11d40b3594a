// Contraction descriptor shared by the GETT kernels (see tne_contract.cu for the scheme).
#pragma once
#include "tne_internal.h"

#define MAXLEG 16

struct LegSet {
    int n;
    int size[MAXLEG];
    long long s0[MAXLEG];
    long long s1[MAXLEG];
    long long s2[MAXLEG];
};

struct GettDesc {
    LegSet M;   // s0: A stride, s1: C stride
    LegSet N;   // s0: B stride, s1: C stride
    LegSet K;   // s0: A stride, s1: B stride
    LegSet Bt;  // s0: A, s1: B, s2: C
    long long Msz, Nsz, Ksz, Bsz;
    long long kPerSplit;
    int nsplit;
    int conjA, conjB, accumulate;
    double alphaRe, alphaIm;  // C (+)= alpha * A * B
    int aKfast;  // A's unit-stride direction is a K leg (else an M leg)
    int bKfast;
};

__device__ __forceinline__ void decomp2(const LegSet& L, long long idx, long long& o0, long long& o1) {
    o0 = 0; o1 = 0;
#pragma unroll 1
    for (int i = 0; i < L.n; ++i) {
        long long q, r;
        if (idx < 0x7fffffffLL) { unsigned u = (unsigned)idx, s = (unsigned)L.size[i]; unsigned uq = u / s; q = uq; r = u - uq * s; }
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
        long long q = idx / L.size[i], r = idx - q * L.size[i];
        o0 += r * L.s0[i]; o1 += r * L.s1[i]; o2 += r * L.s2[i];
        idx = q;
    }
}

__device__ __forceinline__ void atomic_add_c(cplx* p, double re, double im) {
    atomicAdd(&p->x, re);
    atomicAdd(&p->y, im);
}


int gett_try_fast(tne_ctx* ctx, const GettDesc& d, const cplx* A, const cplx* B, cplx* C, bool* handled);
