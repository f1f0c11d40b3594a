// FP64 DMMA fast paths of the contraction (K2/K3 of SURVEY.md 2c; see tne_contract.cu for the GETT scheme).
//
// Both kernels run the complex128 product as a REAL m8n8k4 DMMA GEMM on the interleaved (re, im) storage
// without ever de-interleaving the big operand in memory:
//
//   absorb :  C[m, n] (+)= alpha * sum_k op(A[m, k]) * op(B[k, n])      M huge, K and N <= 32 / 64
//             A row m is read as 2K reals (re0, im0, re1, im1, ...) straight from HBM into DMMA A
//             fragments (registers; every A element is used by exactly one warp, so there is no
//             shared-memory staging).  The small operand is expanded once per CTA into the real
//             2K x 2N matrix  B'[2k+p, 2n+t]  (p, t in {re, im}; conjugations and alpha folded into
//             its signs) and lives in shared memory.  The accumulator fragment of a lane is then exactly
//             one complex output element (re, im) and goes back to HBM with one 128-bit store.
//             Executed flops = 2 * M * 2K * 2N = 8 MNK: the complex count, nothing wasted.
//             This is the boundary-absorption step of the PEPS sweep (forward) and dA = dC * B^H (reverse).
//
//   reduce :  C[m, n] += alpha * sum_k op(A[k, m]) * op(B[k, n])        M x N small (a leaf), K huge
//             the leaf gradients dB = A^H * dC.  The real reduction index is (k, part); the B fragment
//             is (+-) one part of y = B[k, n] chosen per lane, so again no expanded copy exists.
//             Split-K over warps, reduced through shared memory, one atomicAdd per CTA and element.
#include <algorithm>

#include "tne_gett.cuh"

namespace {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

enum { KERNEL_ABSORB = 1, KERNEL_REDUCE = 2 };

// ---------------------------------------------------------------------------------------
// absorb
// ---------------------------------------------------------------------------------------
// KS: DMMA k-steps (4 reals = 2 complex k each); padded K = 2 * KS complex.
template <int KS>
__global__ void __launch_bounds__(256, 2) gett_absorb(const __grid_constant__ GettDesc d, const cplx* __restrict__ A,
                                                   const cplx* __restrict__ B, cplx* __restrict__ C, int Np) {
    constexpr int KP = 2 * KS;            // padded complex K
    constexpr int STRIDE = 4 * KS + 4;    // doubles per B' row (conflict-free fragment reads)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Bs = reinterpret_cast<double*>(smem_raw);                       // [2 Np][STRIDE]
    long long* offAk = reinterpret_cast<long long*>(Bs + 2 * Np * STRIDE);   // [KP]
    long long* offCn = offAk + KP;                                          // [Np]
    const int tid = threadIdx.x;
    for (int k = tid; k < KP; k += blockDim.x) {
        long long a = -1, b = 0;
        if (k < d.Ksz) decomp2(d.K, k, a, b);
        offAk[k] = a;
    }
    for (int n = tid; n < Np; n += blockDim.x) {
        long long b = 0, c = -1;
        if (n < d.Nsz) decomp2(d.N, n, b, c);
        offCn[n] = c;
    }
    const double sA = d.conjA ? -1.0 : 1.0;
    for (int e = tid; e < Np * KP; e += blockDim.x) {
        const int k = e % KP, n = e / KP;
        double br = 0.0, bi = 0.0;
        if (k < d.Ksz && n < d.Nsz) {
            long long ak, bk, bn, cn;
            decomp2(d.K, k, ak, bk);
            decomp2(d.N, n, bn, cn);
            cplx v = B[bk + bn];
            if (d.conjB) v.y = -v.y;
            br = d.alphaRe * v.x - d.alphaIm * v.y;
            bi = d.alphaRe * v.y + d.alphaIm * v.x;
        }
        // (ar + i sA ai)(br + i bi): Re = ar br - sA ai bi ; Im = ar bi + sA ai br
        double* r0 = Bs + (2 * n) * STRIDE + 2 * k;      // output column (n, re)
        double* r1 = Bs + (2 * n + 1) * STRIDE + 2 * k;  // output column (n, im)
        r0[0] = br; r0[1] = -sA * bi;
        r1[0] = bi; r1[1] = sA * br;
    }
    __syncthreads();

    const int lane = tid & 31;
    const int lrow = lane >> 2, lcol = lane & 3;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long ntiles = (d.Msz + 15) >> 4;
    const int NT = Np >> 2;
    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (tid >> 5); w < ntiles; w += nwarps) {
        long long oA[2], oC[2];
        bool ok[2];
        double a[2][KS];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const long long row = (w << 4) + (mt << 3) + lrow;
            ok[mt] = row < d.Msz;
            oA[mt] = 0; oC[mt] = 0;
            if (ok[mt]) decomp2(d.M, row, oA[mt], oC[mt]);
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const long long ofk = offAk[2 * ks + (lcol >> 1)];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                double v = 0.0;
                if (ok[mt] && ofk >= 0) v = reinterpret_cast<const double*>(A + oA[mt] + ofk)[lcol & 1];
                a[mt][ks] = v;
            }
        }
        for (int nt = 0; nt < NT; ++nt) {
            double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            const double* brow = Bs + (8 * nt + lrow) * STRIDE + lcol;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const double b = brow[4 * ks];
                dmma884(c[0][0], c[0][1], a[0][ks], b);
                dmma884(c[1][0], c[1][1], a[1][ks], b);
            }
            const long long ofn = offCn[4 * nt + lcol];
            if (ofn >= 0) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    if (!ok[mt]) continue;
                    cplx* p = C + oC[mt] + ofn;
                    cplx v = make_double2(c[mt][0], c[mt][1]);
                    if (d.accumulate) { cplx o = *p; v.x += o.x; v.y += o.y; }
                    *p = v;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// reduce
// ---------------------------------------------------------------------------------------
// MT x NT output tiles of 8 x 4 complex each; a warp owns a chunk of the reduction range.
template <int MT, int NT>
__global__ void __launch_bounds__(256, 2) gett_reduce(const __grid_constant__ GettDesc d, const cplx* __restrict__ A,
                                                   const cplx* __restrict__ B, cplx* __restrict__ C, int chunk,
                                                   long long nunits) {
    __shared__ double red[MT * NT * 2 * 32];
    for (int e = threadIdx.x; e < MT * NT * 2 * 32; e += blockDim.x) red[e] = 0.0;
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lrow = lane >> 2, lcol = lane & 3;
    const int part = lcol & 1, kin = lcol >> 1;
    // A fragment: x[k][m = 8 mt + lrow], part `part`;  B fragment column (n = 4 nt + (lrow >> 1), t = lrow & 1)
    long long oAm[MT];
    bool okm[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int m = 8 * mt + lrow;
        long long c_;
        okm[mt] = m < d.Msz;
        oAm[mt] = 0;
        if (okm[mt]) decomp2(d.M, m, oAm[mt], c_);
    }
    long long oBn[NT];
    bool okn[NT];
    const int t = lrow & 1;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int n = 4 * nt + (lrow >> 1);
        long long c_;
        okn[nt] = n < d.Nsz;
        oBn[nt] = 0;
        if (okn[nt]) decomp2(d.N, n, oBn[nt], c_);
    }
    // which part of y feeds this lane's B fragment, and with which sign
    const double sx = d.conjA ? -1.0 : 1.0, sy = d.conjB ? -1.0 : 1.0;
    // t = 0 (Re): p = 0 -> +yr ; p = 1 -> -sx sy yi.   t = 1 (Im): p = 0 -> sy yi ; p = 1 -> sx yr
    const int ysel = (t == 0) ? part : 1 - part;
    const double ysgn = (t == 0) ? (part == 0 ? 1.0 : -sx * sy) : (part == 0 ? sy : sx);

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    const int size0 = d.K.n > 0 ? d.K.size[0] : 1;
    const long long sa0 = d.K.n > 0 ? d.K.s0[0] : 0, sb0 = d.K.n > 0 ? d.K.s1[0] : 0;
    const int chunks_per_row = (size0 + chunk - 1) / chunk;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long u = (long long)blockIdx.x * (blockDim.x >> 5) + warp; u < nunits; u += nwarps) {
        const long long outer = u / chunks_per_row;
        const int cidx = (int)(u - outer * chunks_per_row);
        // offsets of the outer K legs (all but the fastest one)
        long long baseA = 0, baseB = 0;
        {
            long long idx = outer;
#pragma unroll 1
            for (int i = 1; i < d.K.n; ++i) {
                long long q = idx / d.K.size[i], r = idx - q * d.K.size[i];
                baseA += r * d.K.s0[i]; baseB += r * d.K.s1[i];
                idx = q;
            }
        }
        const int kbeg = cidx * chunk, kend = min(size0, kbeg + chunk);
#pragma unroll 2
        for (int k0 = kbeg; k0 < kend; k0 += 2) {
            const int k = k0 + kin;
            const bool okk = k < kend;
            const long long offa = baseA + (long long)k * sa0, offb = baseB + (long long)k * sb0;
            double af[MT], bf[NT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                af[mt] = (okk && okm[mt]) ? reinterpret_cast<const double*>(A + oAm[mt] + offa)[part] : 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                bf[nt] = (okk && okn[nt]) ? ysgn * reinterpret_cast<const double*>(B + oBn[nt] + offb)[ysel] : 0.0;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
    }
    // CTA reduction, then one atomic per element: lane holds complex (m = 8 mt + lrow, n = 4 nt + lcol)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            atomicAdd(&red[((mt * NT + nt) * 2 + 0) * 32 + lane], acc[mt][nt][0]);
            atomicAdd(&red[((mt * NT + nt) * 2 + 1) * 32 + lane], acc[mt][nt][1]);
        }
    __syncthreads();
    for (int e = tid; e < MT * NT * 32; e += blockDim.x) {
        const int tile = e >> 5, l = e & 31;
        const double re = red[(tile * 2 + 0) * 32 + l], im = red[(tile * 2 + 1) * 32 + l];
        const int mt = tile / NT, nt = tile - mt * NT;
        const int m = 8 * mt + (l >> 2), n = 4 * nt + (l & 3);
        if (m < d.Msz && n < d.Nsz) {
            long long am, cm, bn, cn;
            decomp2(d.M, m, am, cm);
            decomp2(d.N, n, bn, cn);
            atomic_add_c(C + cm + cn, d.alphaRe * re - d.alphaIm * im, d.alphaRe * im + d.alphaIm * re);
        }
    }
}

void sort_legs_by(LegSet& L, int which) {  // insertion sort of the legs by stride s0 / s1
    for (int i = 1; i < L.n; ++i)
        for (int j = i; j > 0; --j) {
            long long a = which == 0 ? L.s0[j - 1] : L.s1[j - 1], b = which == 0 ? L.s0[j] : L.s1[j];
            if (a <= b) break;
            std::swap(L.size[j - 1], L.size[j]);
            std::swap(L.s0[j - 1], L.s0[j]);
            std::swap(L.s1[j - 1], L.s1[j]);
            std::swap(L.s2[j - 1], L.s2[j]);
        }
}

template <int KS>
int launch_absorb(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gett_absorb<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr_set = true;
    }
    gett_absorb<KS><<<plan.gx, plan.threads, plan.smem, ctx->stream>>>(plan.d, A, B, C, plan.aux[1]);
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}

template <int MT, int NT>
int launch_reduce(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C) {
    long long nunits = ((long long)plan.aux[3] << 31) | (long long)plan.aux[2];
    gett_reduce<MT, NT><<<plan.gx, plan.threads, 0, ctx->stream>>>(plan.d, A, B, C, plan.aux[1], nunits);
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}

}  // namespace

int gett_fast_plan(tne_ctx* ctx, ContractPlan* plan) {
    GettDesc& d = plan->d;
    if (d.Bsz != 1) return TNE_OK;
    // ---- absorb: big M, small K and N ----
    if (d.Msz >= 2048 && d.Ksz <= 32 && d.Nsz <= 64 && d.Ksz * d.Nsz >= 4) {
        int ks = 1;
        while (2 * ks < d.Ksz) ks <<= 1;
        const int Np = (int)((d.Nsz + 3) / 4 * 4);
        sort_legs_by(d.N, 1);  // B is staged in shared memory: order the N legs by the OUTPUT stride
        plan->kernel = KERNEL_ABSORB;
        plan->aux[0] = ks;
        plan->aux[1] = Np;
        plan->threads = 256;
        plan->smem = (size_t)(2 * Np) * (4 * ks + 4) * sizeof(double) + (size_t)(2 * ks + Np) * sizeof(long long);
        const long long ntiles = (d.Msz + 15) / 16;
        const long long ctas = (ntiles + 7) / 8;
        plan->gx = (unsigned)std::min<long long>(ctas, (long long)ctx->sm_count * 2);
        return TNE_OK;
    }
    // ---- reduce: small M x N, long K ----
    if (d.Ksz >= 8192 && d.Msz <= 32 && d.Nsz <= 16 && d.K.n >= 1) {
        int MT = d.Msz <= 8 ? 1 : (d.Msz <= 16 ? 2 : 4);
        int NT = d.Nsz <= 4 ? 1 : (d.Nsz <= 8 ? 2 : 4);
        sort_legs_by(d.K, 0);  // iterate along A's fastest reduction leg
        const int size0 = d.K.size[0];
        const long long outer = d.Ksz / size0;
        const long long warps = (long long)ctx->sm_count * 2 * 8;
        // chunk of the fastest leg per work unit: ~3 units per warp, at least 64 k's, even
        long long chunk = (d.Ksz + 3 * warps - 1) / (3 * warps);
        chunk = std::max<long long>(64, std::min<long long>(chunk, size0));
        chunk = (chunk + 1) / 2 * 2;
        const long long cpr = (size0 + chunk - 1) / chunk;
        const long long nunits = outer * cpr;
        plan->kernel = KERNEL_REDUCE;
        plan->aux[0] = MT * 16 + NT;
        plan->aux[1] = (int)chunk;
        plan->aux[2] = (int)(nunits & 0x7fffffffLL);
        plan->aux[3] = (int)(nunits >> 31);
        plan->threads = 256;
        plan->gx = (unsigned)std::max<long long>(1, std::min<long long>((nunits + 7) / 8, (long long)ctx->sm_count * 2));
        plan->zero_first = !d.accumulate;
        return TNE_OK;
    }
    return TNE_OK;
}

int gett_fast_run(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C) {
    if (plan.kernel == KERNEL_ABSORB) {
        switch (plan.aux[0]) {
            case 1: return launch_absorb<1>(ctx, plan, A, B, C);
            case 2: return launch_absorb<2>(ctx, plan, A, B, C);
            case 4: return launch_absorb<4>(ctx, plan, A, B, C);
            case 8: return launch_absorb<8>(ctx, plan, A, B, C);
            case 16: return launch_absorb<16>(ctx, plan, A, B, C);
        }
        return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "absorb kernel: bad k-step count %d", plan.aux[0]);
    }
    if (plan.kernel == KERNEL_REDUCE) {
        if (plan.zero_first) TNE_TRY(launch_fill_zero(ctx, C, plan.c_nelem));
        switch (plan.aux[0]) {
            case 1 * 16 + 1: return launch_reduce<1, 1>(ctx, plan, A, B, C);
            case 1 * 16 + 2: return launch_reduce<1, 2>(ctx, plan, A, B, C);
            case 1 * 16 + 4: return launch_reduce<1, 4>(ctx, plan, A, B, C);
            case 2 * 16 + 1: return launch_reduce<2, 1>(ctx, plan, A, B, C);
            case 2 * 16 + 2: return launch_reduce<2, 2>(ctx, plan, A, B, C);
            case 2 * 16 + 4: return launch_reduce<2, 4>(ctx, plan, A, B, C);
            case 4 * 16 + 1: return launch_reduce<4, 1>(ctx, plan, A, B, C);
            case 4 * 16 + 2: return launch_reduce<4, 2>(ctx, plan, A, B, C);
            case 4 * 16 + 4: return launch_reduce<4, 4>(ctx, plan, A, B, C);
        }
        return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "reduce kernel: bad tile shape %d", plan.aux[0]);
    }
    return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "unknown fast kernel %d", plan.kernel);
}
