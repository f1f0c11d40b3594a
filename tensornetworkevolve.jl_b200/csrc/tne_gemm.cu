// General dense node of a contraction tree: tiled complex128 GEMM on FP64 DMMA for contractions whose M, N AND K are all
// large (and / or carry batch legs) -- the shapes a greedy / TreeSA order (src/peps.jl:84-93,244-245; test/peps.jl:9) or
// the paired boundary tree (N = K = 256) produces, which the skinny absorb / reduce kernels of tne_contract_dmma.cu do
// not cover.  Replaces OMEinsum's permutedims -> reshape -> batched zgemm -> permutedims for those nodes: operands are
// gathered straight from their own strided layouts (GETT, see tne_contract.cu), no permuted copies exist.
//
//   C[m, n, b] (+)= alpha * sum_k op(A[m, k, b]) * op(B[k, n, b])
//
// CTA tile TM x TN complex, K chunks of TK = 16.  The A tile travels global -> shared with 16-byte cp.async gathers (one
// complex element each, zero-filled outside the tensor) into a two-stage ring and is read as DMMA A fragments exactly
// like the absorb kernels do (row = 2 TK interleaved reals, quad-ordered real k index).  The B tile is fetched one chunk
// ahead into registers and expanded into the real 2 TK x 2 TN matrix B'[(k, part), (n, t)] (signs of conj(A), conj(B)
// and alpha folded in) when it is stored.  One block-wide barrier per chunk.  8 MNK real flops: the complex count (4M).
#include <algorithm>

#include "tne_gett.cuh"

namespace {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;     // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int GM_TK = 16;
constexpr int GM_ASTR = GM_TK + 4;            // complex per A row in shared memory (320 B: conflict-free 128-bit fragment reads)
constexpr int GM_BSTR = 2 * GM_TK + 4;        // doubles per B' row (conflict-free 64-bit fragment reads)

// WM x WN warps; a warp owns (TM / WM) rows = MT m-tiles and (TN / WN) complex columns = NTW n-tiles of 4
template <int TM, int TN, int WM, int WN, int MINB = 1>
__global__ void __launch_bounds__(32 * WM * WN, MINB) gett_gemm(const __grid_constant__ GettDesc d, const cplx* __restrict__ A,
                                                         const cplx* __restrict__ B, cplx* __restrict__ C) {
    constexpr int NTHR = 32 * WM * WN;
    constexpr int MT = TM / WM / 8, NTW = TN / WN / 4;
    constexpr int A_PER = TM * GM_TK / NTHR, B_PER = TN * GM_TK / NTHR;
    static_assert(TM % (8 * WM) == 0 && TN % (4 * WN) == 0 && (TM * GM_TK) % NTHR == 0 && (TN * GM_TK) % NTHR == 0, "tile shape");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* As = reinterpret_cast<cplx*>(smem_raw);                               // [2][TM][GM_ASTR]
    double* Bs = reinterpret_cast<double*>(As + 2 * TM * GM_ASTR);               // [2][2 TN][GM_BSTR]
    long long* offAm = reinterpret_cast<long long*>(Bs + 2 * 2 * TN * GM_BSTR);  // [TM]
    long long* offCm = offAm + TM;                                               // [TM]
    long long* offBn = offCm + TM;                                               // [TN]
    long long* offCn = offBn + TN;                                               // [TN]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lrow = lane >> 2, lcol = lane & 3;
    const int wm = warp % WM, wn = warp / WM;
    const long long m0 = (long long)blockIdx.x * TM, n0 = (long long)blockIdx.y * TN;
    long long oAb, oBb, oCb;
    decomp3(d.Bt, blockIdx.z, oAb, oBb, oCb);
    for (int i = tid; i < TM; i += NTHR) {
        long long a = 0, c = 0;
        if (m0 + i < d.Msz) decomp2(d.M, m0 + i, a, c);
        offAm[i] = a; offCm[i] = c;
    }
    for (int i = tid; i < TN; i += NTHR) {
        long long b = 0, c = 0;
        if (n0 + i < d.Nsz) decomp2(d.N, n0 + i, b, c);
        offBn[i] = b; offCn[i] = c;
    }
    __syncthreads();

    // element -> thread maps: consecutive threads walk the operand's unit-stride direction
    auto a_elem = [&](int r, int& i, int& kk) { const int e = tid + r * NTHR; if (d.aKfast) { kk = e % GM_TK; i = e / GM_TK; } else { i = e % TM; kk = e / TM; } };
    auto b_elem = [&](int r, int& j, int& kk) { const int e = tid + r * NTHR; if (d.bKfast) { kk = e % GM_TK; j = e / GM_TK; } else { j = e % TN; kk = e / TN; } };
    const double sA = d.conjA ? -1.0 : 1.0, sB = d.conjB ? -1.0 : 1.0;

    auto issue_a = [&](int stage, long long k0) {
        cplx* dst = As + stage * TM * GM_ASTR;
#pragma unroll
        for (int r = 0; r < A_PER; ++r) {
            int i, kk;
            a_elem(r, i, kk);
            long long ak = 0, bk = 0;
            const bool ok = (k0 + kk < d.Ksz) && (m0 + i < d.Msz);
            if (ok) decomp2(d.K, k0 + kk, ak, bk);
            cp_async16(smem_u32(dst + i * GM_ASTR + kk), A + oAb + offAm[i] + ak, ok);
        }
        cp_async_commit();
    };
    cplx breg[B_PER];
    auto load_b = [&](long long k0) {
#pragma unroll
        for (int r = 0; r < B_PER; ++r) {
            int j, kk;
            b_elem(r, j, kk);
            cplx v = make_double2(0.0, 0.0);
            if (k0 + kk < d.Ksz && n0 + j < d.Nsz) {
                long long ak, bk;
                decomp2(d.K, k0 + kk, ak, bk);
                v = __ldg(B + oBb + offBn[j] + bk);
            }
            breg[r] = v;
        }
    };
    auto store_b = [&](int stage) {
        double* dst = Bs + stage * 2 * TN * GM_BSTR;
#pragma unroll
        for (int r = 0; r < B_PER; ++r) {
            int j, kk;
            b_elem(r, j, kk);
            const double yr = breg[r].x, yi = sB * breg[r].y;
            const double br = d.alphaRe * yr - d.alphaIm * yi, bi = d.alphaRe * yi + d.alphaIm * yr;
            // (ar + i sA ai)(br + i bi): Re = ar br - sA ai bi ; Im = ar bi + sA ai br
            const int jre = 8 * (kk >> 2) + (kk & 3), jim = jre + 4;     // real k index of (kk, re) / (kk, im): quad-ordered
            double* r0 = dst + (2 * j) * GM_BSTR;
            double* r1 = r0 + GM_BSTR;
            r0[jre] = br; r0[jim] = -sA * bi;
            r1[jre] = bi; r1[jim] = sA * br;
        }
    };

    double acc[MT][NTW][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    const long long nchunks = (d.Ksz + GM_TK - 1) / GM_TK;
    issue_a(0, 0);
    load_b(0);
    store_b(0);
    for (long long c = 0; c < nchunks; ++c) {
        const int st = (int)(c & 1);
        cp_async_wait_all();
        __syncthreads();          // chunk c is complete in As[st] / Bs[st]; everybody has left chunk c - 1 (the other stage)
        if (c + 1 < nchunks) {
            issue_a(st ^ 1, (c + 1) * GM_TK);
            load_b((c + 1) * GM_TK);
        }
        const cplx* as = As + st * TM * GM_ASTR + (wm * MT * 8 + lrow) * GM_ASTR + lcol;
        const double* bs = Bs + st * 2 * TN * GM_BSTR + (wn * NTW * 8 + lrow) * GM_BSTR + lcol;
#pragma unroll
        for (int q = 0; q < GM_TK / 4; ++q) {
            cplx av[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) av[mt] = as[mt * 8 * GM_ASTR + 4 * q];
#pragma unroll
            for (int prt = 0; prt < 2; ++prt) {
                const int ks = 2 * q + prt;
                double bv[NTW];
#pragma unroll
                for (int nt = 0; nt < NTW; ++nt) bv[nt] = bs[nt * 8 * GM_BSTR + 4 * ks];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NTW; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], prt ? av[mt].y : av[mt].x, bv[nt]);
            }
        }
        if (c + 1 < nchunks) store_b(st ^ 1);     // readers of Bs[st ^ 1] (chunk c - 1) finished before this chunk's barrier
    }
    // epilogue: lane holds complex (row 8 mt + lrow, column 4 nt + lcol) of the warp tile
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int i = wm * MT * 8 + mt * 8 + lrow;
        if (m0 + i >= d.Msz) continue;
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt) {
            const int j = wn * NTW * 4 + nt * 4 + lcol;
            if (n0 + j >= d.Nsz) continue;
            cplx* p = C + oCb + offCm[i] + offCn[j];
            cplx v = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
            if (d.accumulate) { const cplx o = *p; v.x += o.x; v.y += o.y; }
            *p = v;
        }
    }
}

template <int TM, int TN>
size_t gemm_smem() {
    return (size_t)2 * TM * GM_ASTR * sizeof(cplx) + (size_t)2 * 2 * TN * GM_BSTR * sizeof(double) + (size_t)(2 * TM + 2 * TN) * sizeof(long long) + 128;
}

template <int TM, int TN, int WM, int WN, int MINB = 1>
int launch_gemm(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C) {
    ensure_smem_attr(ctx, gett_gemm<TM, TN, WM, WN, MINB>, (int)gemm_smem<TM, TN>());
    dim3 grid((unsigned)((plan.d.Msz + TM - 1) / TM), (unsigned)((plan.d.Nsz + TN - 1) / TN), plan.gz);
    gett_gemm<TM, TN, WM, WN, MINB><<<grid, 32 * WM * WN, gemm_smem<TM, TN>(), ctx->stream>>>(plan.d, A, B, C);
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}

}  // namespace

// plan->aux[0]: tile variant (0: 64 x 32, 1: 64 x 16, 2: 128 x 64 with 16 warps, one CTA per SM: for contractions large
// enough to fill the chip with such tiles)
bool gemm_plan(tne_ctx* ctx, ContractPlan* plan) {
    const GettDesc& d = plan->d;
    if (ctx->opt_no_gemm == 1) return false;
    if (d.Msz < 32 || d.Nsz < 8 || d.Ksz < 8) return false;
    if (8.0 * (double)d.Msz * (double)d.Nsz * (double)d.Ksz * (double)d.Bsz < 4.0e6) return false;   // tiny: the generic kernel's latency wins
    int variant = d.Nsz > 20 ? 0 : 1;
    if (d.Msz >= 256 && d.Nsz >= 48 && ((d.Msz + 127) / 128) * ((d.Nsz + 63) / 64) * d.Bsz >= ctx->sm_count) variant = 2;
    if (ctx->opt_no_gemm == 2 && variant == 2) variant = 0;     // tuning knob: small tiles only
    const int TM = variant == 2 ? 128 : 64, TN = variant == 2 ? 64 : (variant == 0 ? 32 : 16);
    const long long gm = (d.Msz + TM - 1) / TM, gn = (d.Nsz + TN - 1) / TN;
    if (gn > 65535 || d.Bsz > 65535 || gm > 0x7fffffffLL) return false;
    // a long reduction with too few output tiles to fill the chip stays with the split-K generic kernel
    if (gm * gn * d.Bsz < ctx->sm_count / 2 && d.Ksz >= 4096) return false;
    plan->kernel = 7;
    plan->aux[0] = variant;
    plan->gx = (unsigned)gm; plan->gy = (unsigned)gn; plan->gz = (unsigned)d.Bsz;
    plan->threads = 256;
    plan->zero_first = false;
    return true;
}

int gemm_run(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C) {
    // 128 x 64 tile: 16 warps (4 x 4, 8 accumulator tiles each, <= 128 registers) hide the per-chunk barrier better than 8 warps
    // with 32 tiles each: 24.7-25.6 vs 21.7-22.5 TFLOP/s on N = K = 256 / 4096^3 (tools/gemm_time.py; option no_gemm = 3: 8 warps)
    if (plan.aux[0] == 2 && ctx->opt_no_gemm == 3) return launch_gemm<128, 64, 4, 2>(ctx, plan, A, B, C);
    if (plan.aux[0] == 2) return launch_gemm<128, 64, 4, 4>(ctx, plan, A, B, C);
    if (plan.aux[0] == 0) return launch_gemm<64, 32, 4, 2>(ctx, plan, A, B, C);
    return launch_gemm<64, 16, 8, 1>(ctx, plan, A, B, C);
}
