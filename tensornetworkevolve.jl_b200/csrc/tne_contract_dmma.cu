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

// workspace of the CTA partials of the reduce kernels (stream-ordered: one launch uses it at a time)
int reduce_part_reserve(tne_ctx* ctx, int64_t elems) {
    if (ctx->plan_only || elems <= ctx->red_part_elems) return TNE_OK;
    if (ctx->red_part) { TNE_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->red_part); ctx->red_part = nullptr; ctx->red_part_elems = 0; }
    const int64_t want = std::max<int64_t>(elems, (int64_t)ctx->sm_count * 2 * 512);
    TNE_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->red_part), (size_t)want * sizeof(cplx)));
    ctx->red_part_elems = want;
    return TNE_OK;
}


namespace {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Programmatic dependent launch: the next absorb kernel may become resident while this one drains; it builds its
// shared-memory tables and B' (which never depend on the previous absorb kernel: small operands are leaves or are
// produced by non-triggering kernels) and then waits for full completion of its predecessors before touching A or C.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

enum { KERNEL_ABSORB = 1, KERNEL_REDUCE = 2, KERNEL_ABSORB_TMA = 3, KERNEL_REDUCE_TMA = 4 };

// ---------------------------------------------------------------------------------------
// absorb
// ---------------------------------------------------------------------------------------
// KS: DMMA k-steps (4 reals = 2 complex k each); padded K = 2 * KS complex.
// MTW: 8-row m-tiles per warp iteration.  The A fragments of the NEXT tile are fetched into a second
// register set before the current tile's DMMAs are issued, so every warp always has HBM loads in flight
// under its math and its stores (software pipeline; there is no shared-memory stage for A because each
// A element is consumed by exactly one warp).
template <int KS, int MTW>
struct AbsorbTile {
    long long oC[MTW];
    bool ok[MTW];
    double a[MTW][KS];
};

// Row -> (A offset, C offset).  A warp walks consecutive tiles, so the decomposition of everything above
// the fastest M leg is cached per lane and only redone when that outer index changes (a few hundred
// cycles of dependent integer latency per tile otherwise -- more than the tile's DMMA issue time).
struct RowCache {
    long long outer = -1, baseA = 0, baseC = 0;
};
__device__ __forceinline__ void row_offsets(const GettDesc& d, RowCache& rc, long long row, long long& oA, long long& oC) {
    const int sh = d.M.shift[0];
    long long i0, outer;
    if (d.M.n == 0) { oA = 0; oC = 0; return; }
    if (sh >= 0) { i0 = row & (long long)(d.M.size[0] - 1); outer = row >> sh; }
    else { outer = row / d.M.size[0]; i0 = row - outer * d.M.size[0]; }
    if (outer != rc.outer) {
        rc.outer = outer;
        long long a = 0, c = 0, idx = outer;
#pragma unroll 1
        for (int i = 1; i < d.M.n; ++i) {
            long long q, r;
            const int s2 = d.M.shift[i];
            if (s2 >= 0) { r = idx & (long long)(d.M.size[i] - 1); q = idx >> s2; }
            else { q = idx / d.M.size[i]; r = idx - q * d.M.size[i]; }
            a += r * d.M.s0[i]; c += r * d.M.s1[i];
            idx = q;
        }
        rc.baseA = a; rc.baseC = c;
    }
    oA = rc.baseA + i0 * d.M.s0[0];
    oC = rc.baseC + i0 * d.M.s1[0];
}

template <int KS, int MTW>
__device__ __forceinline__ void absorb_load(AbsorbTile<KS, MTW>& t, const GettDesc& d, const cplx* __restrict__ A,
                                            const long long* offAk, long long w, int lrow, int lcol, RowCache* rc) {
    long long oA[MTW];
#pragma unroll
    for (int mt = 0; mt < MTW; ++mt) {
        const long long row = w * (8 * MTW) + (mt << 3) + lrow;
        t.ok[mt] = row < d.Msz;
        oA[mt] = 0; t.oC[mt] = 0;
        if (t.ok[mt]) row_offsets(d, rc[mt], row, oA[mt], t.oC[mt]);
    }
    // one 128-bit load per lane and k-quad: lane (lrow, lcol) fetches the complex element (row, k = 4 q + lcol).
    // The real reduction index is ORDERED so that this is already the DMMA A fragment: k-step 2q takes the
    // four real parts of the quad, k-step 2q+1 the four imaginary parts (B' rows are stored in the same order).
#pragma unroll
    for (int q = 0; q < KS / 2; ++q) {
        const long long ofk = offAk[4 * q + lcol];
#pragma unroll
        for (int mt = 0; mt < MTW; ++mt) {
            cplx v = make_double2(0.0, 0.0);
            if (t.ok[mt] && ofk >= 0) v = __ldcs(A + oA[mt] + ofk);
            t.a[mt][2 * q] = v.x;
            t.a[mt][2 * q + 1] = v.y;
        }
    }
}

template <int KS, int MTW>
__device__ __forceinline__ void absorb_compute(const AbsorbTile<KS, MTW>& t, const GettDesc& d, cplx* __restrict__ C,
                                               const double* Bs, const long long* offCn, int NT, int lrow, int lcol) {
    constexpr int STRIDE = 4 * KS + 4;
    constexpr int NTU = 4 / MTW < 1 ? 1 : 4 / MTW;   // n-tiles in flight: >= 4 independent DMMA accumulator chains per warp
    for (int nt0 = 0; nt0 < NT; nt0 += NTU) {   // NT is a multiple of NTU (Np is padded by the host)
        double c[NTU][MTW][2];
#pragma unroll
        for (int u = 0; u < NTU; ++u)
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) { c[u][mt][0] = 0.0; c[u][mt][1] = 0.0; }
        const double* brow = Bs + (8 * nt0 + lrow) * STRIDE + lcol;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            double b[NTU];
#pragma unroll
            for (int u = 0; u < NTU; ++u) b[u] = brow[u * 8 * STRIDE + 4 * ks];
#pragma unroll
            for (int u = 0; u < NTU; ++u)
#pragma unroll
                for (int mt = 0; mt < MTW; ++mt) dmma884(c[u][mt][0], c[u][mt][1], t.a[mt][ks], b[u]);
        }
#pragma unroll
        for (int u = 0; u < NTU; ++u) {
            const long long ofn = offCn[4 * (nt0 + u) + lcol];
            if (ofn < 0) continue;
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) {
                if (!t.ok[mt]) continue;
                cplx* p = C + t.oC[mt] + ofn;
                cplx v = make_double2(c[u][mt][0], c[u][mt][1]);
                if (d.accumulate) { cplx o = *p; v.x += o.x; v.y += o.y; }
                __stcs(p, v);
            }
        }
    }
}

// ROWS: A's rows are contiguous (its fastest leg is the fastest M leg) and K is small.  A lane-per-fragment load
// would touch four far-apart 128-byte lines per instruction (measured 3.4 TB/s for 8 read planes, tools/mem_pattern2.cu);
// instead every instruction reads 32 consecutive rows of ONE k-plane (512 contiguous bytes, 5.3 TB/s) into a
// warp-private shared-memory slab from which the DMMA fragments of the four 8-row tiles are gathered.
template <int KS, int MTW, int MINB, bool ROWS>
__global__ void __launch_bounds__(256, MINB) gett_absorb(const __grid_constant__ GettDesc d, const cplx* __restrict__ A,
                                                         const cplx* __restrict__ B, cplx* __restrict__ C, int Np) {
    constexpr int KP = 2 * KS;            // padded complex K
    constexpr int STRIDE = 4 * KS + 4;    // doubles per B' row (conflict-free fragment reads)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Bs = reinterpret_cast<double*>(smem_raw);                       // [2 Np][STRIDE]
    long long* offAk = reinterpret_cast<long long*>(Bs + 2 * Np * STRIDE);   // [KP]
    long long* offCn = offAk + KP;                                          // [Np]
    cplx* slab = reinterpret_cast<cplx*>(offCn + Np + (Np & 1)) + (threadIdx.x >> 5) * (KP * 34);   // ROWS: [KP][32 + 2] per warp
    const int tid = threadIdx.x;
    pdl_trigger();
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
        const int jre = 8 * (k >> 2) + (k & 3), jim = jre + 4;   // real reduction index of (k, re) and (k, im)
        double* r0 = Bs + (2 * n) * STRIDE;      // output column (n, re)
        double* r1 = Bs + (2 * n + 1) * STRIDE;  // output column (n, im)
        r0[jre] = br; r0[jim] = -sA * bi;
        r1[jre] = bi; r1[jim] = sA * br;
    }
    __syncthreads();
    pdl_wait();

    const int lane = tid & 31;
    const int lrow = lane >> 2, lcol = lane & 3;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long ntiles = (d.Msz + 8 * MTW - 1) / (8 * MTW);
    const int NT = Np >> 2;
    const long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (tid >> 5);
    RowCache rc[MTW];
    AbsorbTile<KS, MTW> t0;
    {
        // chunk-interleaved distribution: warp g takes the chunks g, g + nwarps, ... of CH consecutive tiles, so the
        // whole chip sweeps one narrow window of rows (HBM locality, tools/mem_pattern.cu) while the cached row
        // decomposition still holds inside a chunk
        constexpr int CH = 4;
        const long long nchunks = (ntiles + CH - 1) / CH;
        for (long long c = g; c < nchunks; c += nwarps) {
            const long long wlast = min(ntiles, (c + 1) * CH);
            for (long long w = c * CH; w < wlast; ++w) {
                if (ROWS) {   // MTW == 4: the tile is 32 consecutive rows inside M leg 0 (the host guarantees size0 % 32 == 0)
                    long long oA0, oC0;
                    RowCache r0 = rc[0];
                    row_offsets(d, r0, w * 32, oA0, oC0);
                    const bool okrow = w * 32 + lane < d.Msz;
#pragma unroll
                    for (int k = 0; k < KP; ++k) {
                        const long long ofk = offAk[k];
                        cplx v = make_double2(0.0, 0.0);
                        if (okrow && ofk >= 0) v = __ldcs(A + oA0 + lane + ofk);
                        slab[k * 34 + lane] = v;
                    }
                    __syncwarp();
#pragma unroll
                    for (int mt = 0; mt < MTW; ++mt) {
                        const long long row = w * 32 + (mt << 3) + lrow;
                        long long oA;
                        t0.ok[mt] = row < d.Msz;
                        t0.oC[mt] = 0;
                        if (t0.ok[mt]) row_offsets(d, rc[mt], row, oA, t0.oC[mt]);
#pragma unroll
                        for (int q = 0; q < KS / 2; ++q) {
                            const cplx v = slab[(4 * q + lcol) * 34 + (mt << 3) + lrow];
                            t0.a[mt][2 * q] = v.x;
                            t0.a[mt][2 * q + 1] = v.y;
                        }
                    }
                    __syncwarp();
                } else {
                    absorb_load(t0, d, A, offAk, w, lrow, lcol, rc);
                }
                absorb_compute(t0, d, C, Bs, offCn, NT, lrow, lcol);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// absorb with several contributions into one output:  C (+)= sum_t alpha_t op(A_t) op(B_t)  where all A_t have
// the same layout (the variants of the term dynamic programme at one tree node, or the cotangents of all
// consumers of one value).  The accumulators stay in registers across the contributions, so C is written once
// instead of being read and re-written by every further contribution.
// ---------------------------------------------------------------------------------------
template <int KS, int NTMAX>
__global__ void __launch_bounds__(256, 3) gett_absorb_multi(const __grid_constant__ GettDesc d, const __grid_constant__ AbsorbMulti ops,
                                                            cplx* __restrict__ C, int Np) {
    constexpr int KP = 2 * KS;
    constexpr int STRIDE = 4 * KS + 4;    // NTMAX = Np / 4 n-tiles, all accumulated at once
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Bs = reinterpret_cast<double*>(smem_raw);                                   // [T][2 Np][STRIDE]
    long long* offAk = reinterpret_cast<long long*>(Bs + ops.T * 2 * Np * STRIDE);       // [KP]
    long long* offCn = offAk + KP;                                                      // [Np]
    const int tid = threadIdx.x;
    pdl_trigger();
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
    for (int e = tid; e < ops.T * Np * KP; e += blockDim.x) {
        const int t = e / (Np * KP), r = e - t * (Np * KP);
        const int k = r % KP, n = r / KP;
        const double sA = ops.conjA[t] ? -1.0 : 1.0;
        double br = 0.0, bi = 0.0;
        if (k < d.Ksz && n < d.Nsz) {
            long long ak, bk, bn, cn;
            decomp2(d.K, k, ak, bk);
            decomp2(d.N, n, bn, cn);
            cplx v = ops.B[t][bk + bn];
            if (ops.conjB[t]) v.y = -v.y;
            br = ops.are[t] * v.x - ops.aim[t] * v.y;
            bi = ops.are[t] * v.y + ops.aim[t] * v.x;
        }
        const int jre = 8 * (k >> 2) + (k & 3), jim = jre + 4;
        double* r0 = Bs + (t * 2 * Np + 2 * n) * STRIDE;
        double* r1 = r0 + STRIDE;
        r0[jre] = br; r0[jim] = -sA * bi;
        r1[jre] = bi; r1[jim] = sA * br;
    }
    __syncthreads();
    pdl_wait();

    const int lane = tid & 31;
    const int lrow = lane >> 2, lcol = lane & 3;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long ntiles = (d.Msz + 7) >> 3;
    const long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (tid >> 5);
    RowCache rc;
    constexpr int CH = 4;
    const long long nchunks = (ntiles + CH - 1) / CH;
    for (long long ch = g; ch < nchunks; ch += nwarps) {
        const long long wlast = min(ntiles, (ch + 1) * CH);
        for (long long w = ch * CH; w < wlast; ++w) {
            const long long row = (w << 3) + lrow;
            const bool ok = row < d.Msz;
            long long oA = 0, oC = 0;
            if (ok) row_offsets(d, rc, row, oA, oC);
            double c[NTMAX][2];
#pragma unroll
            for (int nt = 0; nt < NTMAX; ++nt) { c[nt][0] = 0.0; c[nt][1] = 0.0; }
            for (int t = 0; t < ops.T; ++t) {
                double a[KS];
                const cplx* At = ops.A[t];
#pragma unroll
                for (int q = 0; q < KS / 2; ++q) {
                    const long long ofk = offAk[4 * q + lcol];
                    cplx v = make_double2(0.0, 0.0);
                    if (ok && ofk >= 0) v = __ldcs(At + oA + ofk);
                    a[2 * q] = v.x; a[2 * q + 1] = v.y;
                }
                const double* brow = Bs + (t * 2 * Np + lrow) * STRIDE + lcol;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                    for (int nt = 0; nt < NTMAX; ++nt) dmma884(c[nt][0], c[nt][1], a[ks], brow[nt * 8 * STRIDE + 4 * ks]);
                }
            }
            if (ok) {
#pragma unroll
                for (int nt = 0; nt < NTMAX; ++nt) {
                    const long long ofn = offCn[4 * nt + lcol];
                    if (ofn < 0) continue;
                    cplx* p = C + oC + ofn;
                    cplx v = make_double2(c[nt][0], c[nt][1]);
                    if (d.accumulate) { cplx o = *p; v.x += o.x; v.y += o.y; }
                    __stcs(p, v);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// absorb, TMA-staged: the same math, but the A tiles travel HBM -> shared memory as bulk async copies
// (cp.async.bulk, completion on an mbarrier) issued by a producer warp into a STAGES-deep ring, so the
// HBM latency is covered by the ring instead of by registers of DMMA-heavy warps.  Applies when a tile of
// R consecutive rows is made of long contiguous runs in A:
//   kf > 1 : A's fastest leg is a K leg of size kf and the fastest M leg follows it (forward sweep step):
//            one run of R * kf elements per remaining K index ("plane");
//   kf = 1 : A's fastest leg is the fastest M leg (reverse step dA = dC * B^H): one run of R elements per k.
// The shared-memory image keeps the runs as they are in memory; a lane's 128-bit fragment read
// (row, k = 4 q + lcol) is conflict-free in both cases (32-byte pad per plane when kf = 1).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct TmaGeom {
    int tpc;            // tiles (stages) per CTA: a CTA walks one contiguous chunk of rows
    int kf;             // elements of the contiguous K run inside a row (1: rows themselves are contiguous)
    int nplanes;        // bulk copies per stage
    int plane_stride;   // bytes between planes in shared memory
    int Np;
};

constexpr int TMA_STAGES = 4;
constexpr int TMA_CONSUMERS = 4;   // consumer warps; warp TMA_CONSUMERS is the producer

template <int KS, int MTW>
__global__ void __launch_bounds__(32 * (TMA_CONSUMERS + 1), 2)
gett_absorb_tma(const __grid_constant__ GettDesc d, const cplx* __restrict__ A, const cplx* __restrict__ B,
                cplx* __restrict__ C, const TmaGeom g) {
    constexpr int KP = 2 * KS;
    constexpr int STRIDE = 4 * KS + 4;
    constexpr int R = 8 * MTW * TMA_CONSUMERS;   // rows per stage
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Np = g.Np;
    const int stage_bytes = g.nplanes * g.plane_stride;
    unsigned char* stages = smem_raw;
    double* Bs = reinterpret_cast<double*>(smem_raw + TMA_STAGES * stage_bytes);          // [2 Np][STRIDE]
    long long* planeoff = reinterpret_cast<long long*>(Bs + 2 * Np * STRIDE);              // [nplanes] A offset of each plane
    long long* offCn = planeoff + 32;                                                      // [Np]
    int* koff = reinterpret_cast<int*>(offCn + Np);                                        // [KP] smem byte offset of (row 0, k); -1: padding
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(koff + KP + (KP & 1)); // full[STAGES], empty[STAGES]
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    for (int pl = tid; pl < g.nplanes; pl += nthr) {
        long long a, b;
        decomp2(d.K, (long long)pl * g.kf, a, b);
        planeoff[pl] = a;
    }
    for (int k = tid; k < KP; k += nthr) koff[k] = k < d.Ksz ? (k / g.kf) * g.plane_stride + (k % g.kf) * 16 : -1;
    for (int n = tid; n < Np; n += nthr) {
        long long b = 0, c = -1;
        if (n < d.Nsz) decomp2(d.N, n, b, c);
        offCn[n] = c;
    }
    const double sA = d.conjA ? -1.0 : 1.0;
    for (int e = tid; e < Np * KP; e += nthr) {
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
        const int jre = 8 * (k >> 2) + (k & 3), jim = jre + 4;
        double* r0 = Bs + (2 * n) * STRIDE;
        double* r1 = Bs + (2 * n + 1) * STRIDE;
        r0[jre] = br; r0[jim] = -sA * bi;
        r1[jre] = bi; r1[jim] = sA * br;
    }
    if (tid == 0) {
        for (int s = 0; s < TMA_STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);                          // full: the producer's expect_tx arrive
            mbar_init(smem_u32(&bars[TMA_STAGES + s]), TMA_CONSUMERS);  // empty: one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    // Persistent grid; CTA b walks the tiles b, b + G, b + 2G, ... so that at any time the whole chip works on
    // one narrow window of rows per plane (tools/mem_pattern.cu: 5.8 TB/s for this plane pattern, against
    // 5.0 TB/s when every CTA walks its own far-away contiguous range).
    const long long ntiles = d.Msz / R;   // the host guarantees Msz % R == 0 (M leg 0 is a multiple of R)
    const long long tbeg = blockIdx.x, tend = ntiles, tstep = gridDim.x;
    if (warp == TMA_CONSUMERS) {
        // ---------------- producer warp: lane p copies plane p of every tile ----------------
        RowCache rc;
        int stage = 0;
        unsigned phase = 0;
        const unsigned chunk = (unsigned)(R * g.kf * 16);
        for (long long t = tbeg; t < tend; t += tstep) {
            mbar_wait(smem_u32(&bars[TMA_STAGES + stage]), phase ^ 1);
            long long oA, oC;
            row_offsets(d, rc, t * R, oA, oC);
            if (lane == 0) mbar_expect_tx(smem_u32(&bars[stage]), chunk * (unsigned)g.nplanes);
            __syncwarp();
            for (int pl = lane; pl < g.nplanes; pl += 32)
                bulk_g2s(smem_u32(stages + stage * stage_bytes + pl * g.plane_stride), A + oA + planeoff[pl], chunk,
                         smem_u32(&bars[stage]));
            if (++stage == TMA_STAGES) { stage = 0; phase ^= 1; }
        }
        return;
    }
    // ---------------- consumer warps ----------------
    const int lrow = lane >> 2, lcol = lane & 3;
    const int NT = Np >> 2;
    RowCache rc[MTW];
    int stage = 0;
    unsigned phase = 0;
    for (long long t = tbeg; t < tend; t += tstep) {
        AbsorbTile<KS, MTW> tl;
        mbar_wait(smem_u32(&bars[stage]), phase);
        const unsigned char* sbase = stages + stage * stage_bytes;
#pragma unroll
        for (int mt = 0; mt < MTW; ++mt) {
            const int rl = (warp * MTW + mt) * 8 + lrow;   // row inside the stage
            long long oA;
            tl.ok[mt] = true;
            row_offsets(d, rc[mt], t * R + rl, oA, tl.oC[mt]);
#pragma unroll
            for (int q = 0; q < KS / 2; ++q) {
                const int ko = koff[4 * q + lcol];
                cplx v = make_double2(0.0, 0.0);
                if (ko >= 0) v = *reinterpret_cast<const cplx*>(sbase + ko + rl * g.kf * 16);
                tl.a[mt][2 * q] = v.x;
                tl.a[mt][2 * q + 1] = v.y;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[TMA_STAGES + stage]));   // fragments are in registers: refill the slot
        if (++stage == TMA_STAGES) { stage = 0; phase ^= 1; }
        absorb_compute(tl, d, C, Bs, offCn, NT, lrow, lcol);
    }
}

// ---------------------------------------------------------------------------------------
// reduce
// ---------------------------------------------------------------------------------------
// MT x NT output tiles of 8 x 4 complex each; a warp owns a chunk of the reduction range.
template <int MT, int NT>
__global__ void __launch_bounds__(256, 2) gett_reduce(const __grid_constant__ GettDesc d, const cplx* __restrict__ A,
                                                   const cplx* __restrict__ B, cplx* __restrict__ cta_part, int chunk,
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
                long long q, r;
                if (d.K.shift[i] >= 0) { r = idx & (long long)(d.K.size[i] - 1); q = idx >> d.K.shift[i]; }
                else { q = idx / d.K.size[i]; r = idx - q * d.K.size[i]; }
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
    // CTA reduction in a FIXED warp order (no floating-point atomics: bit-reproducible gradients), then one partial per CTA;
    // lane holds complex (m = 8 mt + lrow, n = 4 nt + lcol)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        if (warp == w) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    red[((mt * NT + nt) * 2 + 0) * 32 + lane] += acc[mt][nt][0];
                    red[((mt * NT + nt) * 2 + 1) * 32 + lane] += acc[mt][nt][1];
                }
        }
        __syncthreads();
    }
    cplx* mine = cta_part + (long long)blockIdx.x * (MT * NT * 32);
    for (int e = tid; e < MT * NT * 32; e += blockDim.x) {
        const int tile = e >> 5, l = e & 31;
        mine[e] = make_double2(red[(tile * 2 + 0) * 32 + l], red[(tile * 2 + 1) * 32 + l]);
    }
}

// ---------------------------------------------------------------------------------------
// reduce, TMA-staged: chunks of KC reduction indices of BOTH big operands are bulk-copied into a shared-memory
// ring by a producer warp; four consumer warps split every chunk's k-steps and keep the whole M x N
// accumulator tile in registers.  For each operand a chunk is a handful of long contiguous runs
// ("planes"): either the reduction leg is the operand's fastest leg (one run of KC elements per free
// index), or a free leg of size f is fastest and the reduction leg follows it (one run of KC * f elements
// per remaining free index).
// ---------------------------------------------------------------------------------------
struct RedGeom {
    int fA, fB;               // contiguous free-leg run (1: the reduction leg itself is fastest)
    int npA, npB;             // planes per operand
    int strideA, strideB;     // bytes between planes in shared memory
    int offB;                 // byte offset of B's planes inside a stage
    int stage_bytes;
    long long nchunks;        // total chunks = (Ksz / size0) * (size0 / KC)
    int chunks_per_row;       // size0 / KC
};
constexpr int RED_KC = 128;      // reduction indices per stage (bulk copies of 1-4 KiB)
constexpr int RED_STAGES = 2;
constexpr int RED_CONSUMERS = 4;

template <int MT, int NT>
__global__ void __launch_bounds__(32 * (RED_CONSUMERS + 1), 1)
gett_reduce_tma(const __grid_constant__ GettDesc d, const cplx* __restrict__ A, const cplx* __restrict__ B,
                cplx* __restrict__ cta_part, const RedGeom g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* stages = smem_raw;
    long long* poffA = reinterpret_cast<long long*>(smem_raw + RED_STAGES * g.stage_bytes);   // [npA] global offset of plane
    long long* poffB = poffA + 32;                                                            // [npB]
    double* red = reinterpret_cast<double*>(poffB + 32);                                      // [MT*NT*2*32]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(red + MT * NT * 2 * 32);
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int p = tid; p < g.npA; p += nthr) { long long a, c; decomp2(d.M, (long long)p * g.fA, a, c); poffA[p] = a; }
    for (int p = tid; p < g.npB; p += nthr) { long long b, c; decomp2(d.N, (long long)p * g.fB, b, c); poffB[p] = b; }
    for (int e = tid; e < MT * NT * 2 * 32; e += nthr) red[e] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < RED_STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[RED_STAGES + s]), RED_CONSUMERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const long long sa0 = d.K.s0[0], sb0 = d.K.s1[0];
    if (warp == RED_CONSUMERS) {
        int stage = 0;
        unsigned phase = 0;
        const unsigned bytesA = (unsigned)(RED_KC * g.fA * 16), bytesB = (unsigned)(RED_KC * g.fB * 16);
        for (long long u = blockIdx.x; u < g.nchunks; u += gridDim.x) {
            const long long outer = u / g.chunks_per_row;
            const long long k0 = (u - outer * g.chunks_per_row) * RED_KC;
            long long baseA = k0 * sa0, baseB = k0 * sb0, idx = outer;
#pragma unroll 1
            for (int i = 1; i < d.K.n; ++i) {
                long long q, r;
                if (d.K.shift[i] >= 0) { r = idx & (long long)(d.K.size[i] - 1); q = idx >> d.K.shift[i]; }
                else { q = idx / d.K.size[i]; r = idx - q * d.K.size[i]; }
                baseA += r * d.K.s0[i]; baseB += r * d.K.s1[i];
                idx = q;
            }
            mbar_wait(smem_u32(&bars[RED_STAGES + stage]), phase ^ 1);
            if (lane == 0) mbar_expect_tx(smem_u32(&bars[stage]), bytesA * (unsigned)g.npA + bytesB * (unsigned)g.npB);
            __syncwarp();
            unsigned char* sb = stages + stage * g.stage_bytes;
            for (int p = lane; p < g.npA; p += 32)
                bulk_g2s(smem_u32(sb + p * g.strideA), A + baseA + poffA[p], bytesA, smem_u32(&bars[stage]));
            for (int p = lane; p < g.npB; p += 32)
                bulk_g2s(smem_u32(sb + g.offB + p * g.strideB), B + baseB + poffB[p], bytesB, smem_u32(&bars[stage]));
            if (++stage == RED_STAGES) { stage = 0; phase ^= 1; }
        }
        return;
    }
    const int lrow = lane >> 2, lcol = lane & 3;
    const int part = lcol & 1, kin = lcol >> 1;
    const int t = lrow & 1;
    const double sx = d.conjA ? -1.0 : 1.0, sy = d.conjB ? -1.0 : 1.0;
    const int ysel = (t == 0) ? part : 1 - part;
    const double ysgn = (t == 0) ? (part == 0 ? 1.0 : -sx * sy) : (part == 0 ? sy : sx);
    // byte offsets of this lane's fragment elements inside a stage, for k_local = 0
    int aoff[MT], boff[NT];
    bool okm[MT], okn[NT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int m = 8 * mt + lrow;
        okm[mt] = m < d.Msz;
        aoff[mt] = (m / g.fA) * g.strideA + (m % g.fA) * 16 + part * 8;
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int n = 4 * nt + (lrow >> 1);
        okn[nt] = n < d.Nsz;
        boff[nt] = g.offB + (n / g.fB) * g.strideB + (n % g.fB) * 16 + ysel * 8;
    }
    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    int stage = 0;
    unsigned phase = 0;
    for (long long u = blockIdx.x; u < g.nchunks; u += gridDim.x) {
        mbar_wait(smem_u32(&bars[stage]), phase);
        const unsigned char* sb = stages + stage * g.stage_bytes;
        constexpr int KSTEPS = RED_KC / 2 / RED_CONSUMERS;   // k-steps (2 complex k) per warp and chunk
        constexpr int KB = 4;                                // ... fetched and multiplied KB at a time
#pragma unroll
        for (int j0 = 0; j0 < KSTEPS; j0 += KB) {
            double af[KB][MT], bf[KB][NT];
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                const int kl = 2 * (warp + RED_CONSUMERS * (j0 + j)) + kin;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) af[j][mt] = okm[mt] ? *reinterpret_cast<const double*>(sb + aoff[mt] + kl * g.fA * 16) : 0.0;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) bf[j][nt] = okn[nt] ? ysgn * *reinterpret_cast<const double*>(sb + boff[nt] + kl * g.fB * 16) : 0.0;
            }
            if (j0 + KB >= KSTEPS) {   // last batch: the fragments are in registers, the slot can be refilled
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars[RED_STAGES + stage]));
            }
#pragma unroll
            for (int j = 0; j < KB; ++j)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[j][mt], bf[j][nt]);
        }
        if (++stage == RED_STAGES) { stage = 0; phase ^= 1; }
    }
    // consumer warps add their accumulators in a FIXED order (no floating-point atomics), one partial per CTA
    for (int w = 0; w < RED_CONSUMERS; ++w) {
        if (warp == w) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    red[((mt * NT + nt) * 2 + 0) * 32 + lane] += acc[mt][nt][0];
                    red[((mt * NT + nt) * 2 + 1) * 32 + lane] += acc[mt][nt][1];
                }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * RED_CONSUMERS));   // consumers only (the producer warp has left)
    }
    cplx* mine = cta_part + (long long)blockIdx.x * (MT * NT * 32);
    for (int e = tid; e < MT * NT * 32; e += 32 * RED_CONSUMERS) {
        const int tile = e >> 5, l = e & 31;
        mine[e] = make_double2(red[(tile * 2 + 0) * 32 + l], red[(tile * 2 + 1) * 32 + l]);
    }
}

// second pass of both reduce kernels: C[m, n] += alpha * sum of the CTA partials.  Eight lanes share one element (lane s adds
// the partials g = s, s + 8, ... in order; fixed shuffle tree): the same association order in every run (deterministic).
__global__ void __launch_bounds__(256) gett_reduce_final(const __grid_constant__ GettDesc d, const cplx* __restrict__ part, int nparts, int NT, int ntile_elems,
                                                        cplx* __restrict__ C) {
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7;
    double re = 0.0, im = 0.0;
    if (e < ntile_elems)
        for (int g = sub; g < nparts; g += 8) { const cplx v = part[(long long)g * ntile_elems + e]; re += v.x; im += v.y; }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        re += __shfl_down_sync(0xffffffffu, re, o, 8);
        im += __shfl_down_sync(0xffffffffu, im, o, 8);
    }
    if (e >= ntile_elems || sub != 0) return;
    const int tile = e >> 5, l = e & 31;
    const int mt = tile / NT, nt = tile - mt * NT;
    const int m = 8 * mt + (l >> 2), n = 4 * nt + (l & 3);
    if (m < d.Msz && n < d.Nsz) {
        long long am, cm, bn, cn;
        decomp2(d.M, m, am, cm);
        decomp2(d.N, n, bn, cn);
        cplx* p = C + cm + cn;
        const cplx o = *p;
        *p = make_double2(o.x + d.alphaRe * re - d.alphaIm * im, o.y + d.alphaRe * im + d.alphaIm * re);
    }
}

void sort_legs_by(LegSet& L, int which) {  // insertion sort of the legs by stride s0 / s1
    for (int i = 1; i < L.n; ++i)
        for (int j = i; j > 0; --j) {
            long long a = which == 0 ? L.s0[j - 1] : L.s1[j - 1], b = which == 0 ? L.s0[j] : L.s1[j];
            if (a <= b) break;
            std::swap(L.size[j - 1], L.size[j]);
            std::swap(L.shift[j - 1], L.shift[j]);
            std::swap(L.s0[j - 1], L.s0[j]);
            std::swap(L.s1[j - 1], L.s1[j]);
            std::swap(L.s2[j - 1], L.s2[j]);
        }
}

template <int KS, int MTW, int MINB, bool ROWS>
int launch_absorb(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C) {
    ensure_smem_attr(ctx, gett_absorb<KS, MTW, MINB, ROWS>, 100 * 1024);
    unsigned gx = (unsigned)std::min<long long>((long long)plan.gx / 2 * MINB, ((plan.d.Msz + 8 * MTW - 1) / (8 * MTW) + 7) / 8);
    TNE_CUDA(ctx, launch_pdl(gett_absorb<KS, MTW, MINB, ROWS>, gx, (unsigned)plan.threads, plan.smem + (ROWS ? (size_t)8 * (2 * KS) * 34 * sizeof(cplx) + 16 : 0), ctx->stream, !ctx->opt_no_pdl, plan.d, A, B, C, plan.aux[1]));
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}

template <int KS, int NT>
int launch_absorb_multi(tne_ctx* ctx, const ContractPlan& plan, const AbsorbMulti& ops, cplx* C) {
    ensure_smem_attr(ctx, gett_absorb_multi<KS, NT>, 75 * 1024);
    const int Np = plan.aux[1];
    const size_t smem = (size_t)ops.T * (2 * Np) * (4 * KS + 4) * sizeof(double) + (size_t)(2 * KS + Np) * sizeof(long long);
    unsigned gx = (unsigned)std::min<long long>((long long)plan.gx / 2 * 3, ((plan.d.Msz + 7) / 8 + 7) / 8);
    TNE_CUDA(ctx, launch_pdl(gett_absorb_multi<KS, NT>, gx, 256u, smem, ctx->stream, !ctx->opt_no_pdl, plan.d, ops, C, Np));
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}

template <int KS, int MTW>
int launch_absorb_tma(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C) {
    ensure_smem_attr(ctx, gett_absorb_tma<KS, MTW>, 110 * 1024);
    TmaGeom g;
    g.kf = plan.aux[4]; g.nplanes = plan.aux[5]; g.plane_stride = plan.aux[6]; g.Np = plan.aux[1]; g.tpc = plan.aux[2];
    gett_absorb_tma<KS, MTW><<<plan.gx, 32 * (TMA_CONSUMERS + 1), plan.smem, ctx->stream>>>(plan.d, A, B, C, g);
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}

template <int MT, int NT>
int launch_reduce_tma(tne_ctx* ctx, const ContractPlan& plan, const RedGeom& g, const cplx* A, const cplx* B, cplx* C) {
    ensure_smem_attr(ctx, gett_reduce_tma<MT, NT>, 220 * 1024);
    TNE_TRY(reduce_part_reserve(ctx, (int64_t)plan.gx * MT * NT * 32));
    gett_reduce_tma<MT, NT><<<plan.gx, 32 * (RED_CONSUMERS + 1), plan.smem, ctx->stream>>>(plan.d, A, B, ctx->red_part, g);
    TNE_LAUNCH_CHECK(ctx);
    gett_reduce_final<<<(MT * NT * 32 * 8 + 255) / 256, 256, 0, ctx->stream>>>(plan.d, ctx->red_part, (int)plan.gx, NT, MT * NT * 32, C);
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}

template <int MT, int NT>
int launch_reduce(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C) {
    long long nunits = ((long long)plan.aux[3] << 31) | (long long)plan.aux[2];
    TNE_TRY(reduce_part_reserve(ctx, (int64_t)plan.gx * MT * NT * 32));
    gett_reduce<MT, NT><<<plan.gx, plan.threads, 0, ctx->stream>>>(plan.d, A, B, ctx->red_part, plan.aux[1], nunits);
    TNE_LAUNCH_CHECK(ctx);
    gett_reduce_final<<<(MT * NT * 32 * 8 + 255) / 256, 256, 0, ctx->stream>>>(plan.d, ctx->red_part, (int)plan.gx, NT, MT * NT * 32, C);
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}

}  // namespace

int gett_fast_plan(tne_ctx* ctx, ContractPlan* plan) {
    GettDesc& d = plan->d;
    if (d.Bsz != 1) return TNE_OK;
    // ---- absorb: big M, small K and N ----
    if (d.Msz >= 2048 && d.Ksz <= 32 && d.Nsz <= 64 && d.Ksz * d.Nsz >= 4) {
        int ks = 2;   // k-steps come in pairs (a quad of complex k per 128-bit load)
        while (2 * ks < d.Ksz) ks <<= 1;
        // n-tiles are processed NTU = 4 / MTW at a time: pad N so that the tile count is a multiple of NTU
        // (zero columns of B', never stored) unless that would more than double the math
        const int ntu = ks >= 8 ? 4 : 1;   // n-tiles per pass = 4 / MTW (MTW = 1 for the dense steps, 4 for K <= 8)
        int Np = (int)((d.Nsz + 3) / 4 * 4);
        Np = (Np + 4 * ntu - 1) / (4 * ntu) * (4 * ntu);
        sort_legs_by(d.N, 1);  // B is staged in shared memory: order the N legs by the OUTPUT stride
        plan->kernel = KERNEL_ABSORB;
        plan->aux[0] = ks;
        plan->aux[1] = Np;
        plan->aux[3] = (ks <= 4 && d.M.n >= 1 && d.M.s0[0] == 1 && d.M.size[0] % 32 == 0 && !ctx->opt_no_rows) ? 1 : 0;   // ROWS load path
        plan->threads = 256;
        // TMA-staged variant when a tile of R rows is made of long contiguous runs of A (see gett_absorb_tma)
        if ((ctx->opt_absorb_tma == 1 || (ctx->opt_absorb_tma == 2 && ks <= 4)) && d.M.n >= 1) {
            const int mtw = ks >= 16 ? 1 : (ks >= 8 ? 2 : 4);
            const int R = 8 * mtw * TMA_CONSUMERS;
            LegSet Ks = d.K;
            sort_legs_by(Ks, 0);
            int kf = 0;
            if (d.M.s0[0] == 1) kf = 1;                                                  // rows contiguous
            else if (Ks.n >= 1 && Ks.s0[0] == 1 && d.M.s0[0] == Ks.size[0]) kf = Ks.size[0];  // K run, then rows
            if (kf > 0 && d.M.size[0] % R == 0 && d.Msz >= 4LL * R * ctx->sm_count) {
                d.K = Ks;  // plane p <-> K index p * kf needs the contiguous leg first
                const int nplanes = (int)(d.Ksz / kf);
                const int plane_stride = R * kf * 16 + (kf == 1 ? 32 : 0);
                const int Np2 = Np;   // already a multiple of 16 columns (4 n-tiles)
                size_t smem = (size_t)TMA_STAGES * nplanes * plane_stride + (size_t)(2 * Np2) * (4 * ks + 4) * sizeof(double) +
                              (32 + Np2) * sizeof(long long) + (2 * ks + 2) * sizeof(int) + 2 * TMA_STAGES * sizeof(long long) + 128;
                if (nplanes <= 32 && smem <= 100 * 1024) {
                    plan->kernel = KERNEL_ABSORB_TMA;
                    plan->aux[1] = Np2;
                    plan->aux[4] = kf; plan->aux[5] = nplanes; plan->aux[6] = plane_stride; plan->aux[7] = mtw;
                    plan->smem = smem;
                    plan->aux[2] = 1;
                    plan->gx = (unsigned)std::min<long long>((long long)ctx->sm_count * 2, d.Msz / R);
                    return TNE_OK;
                }
            }
        }
        plan->smem = (size_t)(2 * Np) * (4 * ks + 4) * sizeof(double) + (size_t)(2 * ks + Np) * sizeof(long long);
        plan->gx = (unsigned)ctx->sm_count * 2;   // resident CTAs at 2 per SM; the launcher scales with the variant's occupancy
        return TNE_OK;
    }
    // ---- reduce: small M x N, long K ----
    if (d.Ksz >= 8192 && d.Msz <= 32 && d.Nsz <= 16 && d.K.n >= 1) {
        int MT = d.Msz <= 8 ? 1 : (d.Msz <= 16 ? 2 : 4);
        int NT = d.Nsz <= 4 ? 1 : (d.Nsz <= 8 ? 2 : 4);
        sort_legs_by(d.K, 0);  // iterate along A's fastest reduction leg
        const int size0 = d.K.size[0];
        if (!ctx->opt_no_reduce_tma && d.M.n >= 1 && d.N.n >= 1) {
            // TMA-staged variant: both operands must be made of long contiguous runs along K leg 0
            LegSet Ms = d.M, Ns = d.N;
            sort_legs_by(Ms, 0);
            sort_legs_by(Ns, 0);
            int fA = 0, fB = 0;
            if (d.K.s0[0] == 1) fA = 1; else if (Ms.s0[0] == 1 && d.K.s0[0] == Ms.size[0]) fA = Ms.size[0];
            if (d.K.s1[0] == 1) fB = 1; else if (Ns.s0[0] == 1 && d.K.s1[0] == Ns.size[0]) fB = Ns.size[0];
            if (fA > 0 && fB > 0 && size0 % RED_KC == 0 && d.Msz % fA == 0 && d.Nsz % fB == 0 && d.Msz / fA <= 32 && d.Nsz / fB <= 32) {
                d.M = Ms; d.N = Ns;   // free index m = m0 + fA * plane needs the contiguous leg first
                RedGeom g;
                g.fA = fA; g.fB = fB; g.npA = (int)(d.Msz / fA); g.npB = (int)(d.Nsz / fB);
                g.strideA = RED_KC * fA * 16 + 32; g.strideB = RED_KC * fB * 16 + 32;
                g.offB = g.npA * g.strideA;
                g.stage_bytes = (g.offB + g.npB * g.strideB + 127) / 128 * 128;
                g.chunks_per_row = size0 / RED_KC;
                g.nchunks = (d.Ksz / size0) * g.chunks_per_row;
                size_t smem = (size_t)RED_STAGES * g.stage_bytes + 64 * sizeof(long long) + (size_t)MT * NT * 64 * sizeof(double) + 2 * RED_STAGES * sizeof(long long) + 128;
                if (smem <= 220 * 1024) {
                    plan->kernel = KERNEL_REDUCE_TMA;
                    plan->aux[0] = MT * 16 + NT;
                    static_assert(sizeof(RedGeom) <= sizeof(plan->blob), "RedGeom must fit the plan blob");
                    memcpy(plan->blob, &g, sizeof(g));
                    plan->smem = smem;
                    plan->threads = 32 * (RED_CONSUMERS + 1);
                    plan->gx = (unsigned)std::max<long long>(1, std::min<long long>(g.nchunks, (long long)ctx->sm_count * (smem > 110 * 1024 ? 1 : 2)));
                    plan->zero_first = !d.accumulate;
                    return TNE_OK;
                }
            }
        }
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

bool gett_multi_capable(const ContractPlan& plan) {
    // instantiated shapes: (K <= 32, N <= 16) and (K <= 16, N <= 32) -- the two dense steps of the sweep
    return plan.kernel == KERNEL_ABSORB && ((plan.aux[0] == 16 && plan.aux[1] == 16) || (plan.aux[0] == 8 && plan.aux[1] == 32));
}

int gett_fast_run_multi(tne_ctx* ctx, const ContractPlan& plan, const AbsorbMulti& ops, cplx* C) {
    if (plan.aux[0] == 16 && plan.aux[1] == 16) return launch_absorb_multi<16, 4>(ctx, plan, ops, C);
    if (plan.aux[0] == 8 && plan.aux[1] == 32) return launch_absorb_multi<8, 8>(ctx, plan, ops, C);
    return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "absorb (multi) kernel: shape (%d k-steps, %d columns) not instantiated", plan.aux[0], plan.aux[1]);
}

int gett_fast_run(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C) {
    if (plan.kernel == KERNEL_ABSORB_TMA) {
        switch (plan.aux[0]) {
            case 2: return launch_absorb_tma<2, 4>(ctx, plan, A, B, C);
            case 4: return launch_absorb_tma<4, 4>(ctx, plan, A, B, C);
            case 8: return launch_absorb_tma<8, 2>(ctx, plan, A, B, C);
            case 16: return launch_absorb_tma<16, 1>(ctx, plan, A, B, C);
        }
        return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "absorb (TMA) kernel: bad k-step count %d", plan.aux[0]);
    }
    if (plan.kernel == KERNEL_ABSORB) {
        switch (plan.aux[0]) {   // 3 CTAs of 8 warps per SM, one 8-row m-tile per warp iteration
            case 2: return plan.aux[3] ? launch_absorb<2, 4, 3, true>(ctx, plan, A, B, C) : launch_absorb<2, 4, 3, false>(ctx, plan, A, B, C);
            case 4: return plan.aux[3] ? launch_absorb<4, 4, 3, true>(ctx, plan, A, B, C) : launch_absorb<4, 4, 3, false>(ctx, plan, A, B, C);
            case 8: return launch_absorb<8, 1, 3, false>(ctx, plan, A, B, C);
            case 16: return launch_absorb<16, 1, 3, false>(ctx, plan, A, B, C);
        }
        return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "absorb kernel: bad k-step count %d", plan.aux[0]);
    }
    if (plan.kernel == KERNEL_REDUCE_TMA) {
        if (plan.zero_first) TNE_TRY(launch_fill_zero(ctx, C, plan.c_nelem));
        RedGeom g;
        memcpy(&g, plan.blob, sizeof(g));
        switch (plan.aux[0]) {
            case 1 * 16 + 1: return launch_reduce_tma<1, 1>(ctx, plan, g, A, B, C);
            case 1 * 16 + 2: return launch_reduce_tma<1, 2>(ctx, plan, g, A, B, C);
            case 1 * 16 + 4: return launch_reduce_tma<1, 4>(ctx, plan, g, A, B, C);
            case 2 * 16 + 1: return launch_reduce_tma<2, 1>(ctx, plan, g, A, B, C);
            case 2 * 16 + 2: return launch_reduce_tma<2, 2>(ctx, plan, g, A, B, C);
            case 2 * 16 + 4: return launch_reduce_tma<2, 4>(ctx, plan, g, A, B, C);
            case 4 * 16 + 1: return launch_reduce_tma<4, 1>(ctx, plan, g, A, B, C);
            case 4 * 16 + 2: return launch_reduce_tma<4, 2>(ctx, plan, g, A, B, C);
            case 4 * 16 + 4: return launch_reduce_tma<4, 4>(ctx, plan, g, A, B, C);
        }
        return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "reduce (TMA) kernel: bad tile shape %d", plan.aux[0]);
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
