// Fused REVERSE pass of one PEPS site of the boundary sweep (FP64 DMMA), the twin of tne_site2.cu:
//
//   forward (two steps of the tree):   T = A . B1   (bra tensor),      C = T . B2   (ket tensor)
//   reverse (three ops of the source-transformed reverse mode, src/TensorAD/rules.jl:1-8 / OMEinsum.einsum_grad):
//        dT  = a1 f(dC) f(B2)           M x (x, p)   <-  M x n2          (the ket tensor carries no gradient)
//        dA  = a2 f(dT) f(B1)           M x k1       <-  M x (p, n1)
//        dB1 += a3 sum_M f(A) f(dT)     k1 x (p, n1)                    (leaf gradient of the bra tensor)
//
// Run separately these move 9 |A| bytes (dT is twice as large as A and is written once and read twice); here dT never
// leaves the SM: dC and A are read once, dA is written once (3 |A| bytes, 16 flop/B), the same 3 x 8 |M| 512 flops.
//
// Index classes as in tne_site2.cu (labels of A/dA, B1/dB1, B2, C/dC):
//     K1 = A, B1     X = A, B2     P = B1, B2     N1 = B1, C     N2 = B2, C     M = A, C (boundary rows m'')
// instantiated for the bulk site of a D = 4 PEPS (|K1| = |X| = |N1| = |N2| = 16, |P| = 2).
//
// Work split: a CTA (4 warps) takes 8 consecutive boundary rows; warp w owns the four x values 4w .. 4w+3 in ALL three
// stages, so no stage depends on another warp's results and there is no block-wide barrier in the pipeline:
//   stage 1  rows (n1, m''), K = n2, N = the warp's 8 columns (xi, p): B2' fragments live in registers for the whole
//            kernel (16 doubles), dC fragments come straight from HBM (every warp reads the tile; L1 serves the repeats);
//   stage 2  rows (x, m''), K = (p, n1), N = k1: A fragments come from the warp's private 16 KiB scratch tile (XOR-swizzled:
//            conflict-free for the 128-bit stores of stage 1, the 128-bit loads here and the 64-bit loads of stage 3),
//            B1' from shared memory, dA fragments go straight to HBM;
//   stage 3  out (k1, (p, n1)), reduction over (x, m''): A fragments straight from HBM, dT fragments from the scratch tile;
//            the 16 x 32 accumulator tile stays in registers over all tiles of the warp.
// The per-warp partial leaf gradients are summed in a FIXED order (warps through shared memory, then CTAs by a second tiny
// kernel): no floating-point atomics, bit-reproducible gradients.
#include <algorithm>

#include "tne_gett.cuh"

namespace {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int SB_K1 = 16, SB_X = 16, SB_P = 2, SB_N1 = 16, SB_N2 = 16, SB_TB = 8;
constexpr int SB_WARPS = 4, SB_XW = SB_X / SB_WARPS;        // 4 x values per warp
constexpr int SB_STRB1 = 2 * SB_P * SB_N1 + 4;              // doubles per B1' row (64 reals of (p, n1) + pad)
constexpr int SB_NOUT = SB_K1 * SB_P * SB_N1;               // 512 elements of dB1

__device__ __noinline__ long long legs4_off(const Legs4& L, int idx, int which) {
    long long o = 0;
    for (int i = 0; i < L.n; ++i) {
        int r = idx % L.size[i];
        idx /= L.size[i];
        o += r * (which == 0 ? L.s0[i] : L.s1[i]);
    }
    return o;
}

// scratch tile of one warp: row (n1, m'') of 8 chunks of 16 B (the warp's columns c = 2 xi + p), chunk XOR-swizzled
__device__ __forceinline__ int scr_idx(int n1, int mrow, int c) {
    return ((n1 * SB_TB + mrow) << 3) + (c ^ ((mrow & 1) << 2) ^ ((n1 & 1) << 1));
}

__device__ __forceinline__ double flip_sign(double x, int mask) {   // x * (+-1) on the integer pipe (the FP64 pipe belongs to DMMA)
    return __hiloint2double(__double2hiint(x) ^ mask, __double2loint(x));
}

__global__ void __launch_bounds__(32 * SB_WARPS, 2)
gett_site2_bwd(const __grid_constant__ Site2BwdDesc d, const cplx* __restrict__ dC, const cplx* __restrict__ B2,
               const cplx* __restrict__ B1, const cplx* __restrict__ A, cplx* __restrict__ dA, cplx* __restrict__ part) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* B1s = reinterpret_cast<double*>(smem_raw);                                   // [2 K1 = 32][SB_STRB1]
    double* B2f = B1s + 2 * SB_K1 * SB_STRB1;                                              // [SB_WARPS][2 j][8 ks][32 lanes]: B2' fragments
    cplx* scratch = reinterpret_cast<cplx*>(B2f + SB_WARPS * 2 * 8 * 32);                  // [SB_WARPS][128 * 8]
    long long* offC_n1 = reinterpret_cast<long long*>(scratch + SB_WARPS * SB_N1 * SB_TB * 8);   // [16] each
    long long* offC_n2 = offC_n1 + 16;
    long long* offA_k1 = offC_n2 + 16;
    long long* offA_x = offA_k1 + 16;
    long long* offD_k1 = offA_x + 16;     // dA
    long long* offD_x = offD_k1 + 16;
    long long* offB1_k1 = offD_x + 16;    // small operands
    long long* offB1_n1 = offB1_k1 + 16;
    long long* offB2_x = offB1_n1 + 16;
    long long* offB2_n2 = offB2_x + 16;
    long long* offP = offB2_n2 + 16;      // [0..1]: B1 stride of p, [2..3]: B2 stride of p
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lrow = lane >> 2, lcol = lane & 3;
    if (tid < 16) {
        offC_n1[tid] = legs4_off(d.N1, tid, 1);
        offC_n2[tid] = legs4_off(d.N2, tid, 1);
        offA_k1[tid] = legs4_off(d.K1, tid, 0);
        offA_x[tid] = legs4_off(d.X, tid, 0);
        offD_k1[tid] = legs4_off(d.K1d, tid, 0);
        offD_x[tid] = legs4_off(d.Xd, tid, 0);
        offB1_k1[tid] = legs4_off(d.K1, tid, 1);
        offB1_n1[tid] = legs4_off(d.N1, tid, 0);
        offB2_x[tid] = legs4_off(d.X, tid, 1);
        offB2_n2[tid] = legs4_off(d.N2, tid, 0);
        if (tid < 2) { offP[tid] = legs4_off(d.P, tid, 0); offP[2 + tid] = legs4_off(d.P, tid, 1); }
    }
    __syncthreads();
    // B1'[(k1, t)][real index of k = p + 2 n1]:  dA = a2 * f(dT) * f(B1), contraction over (p, n1)
    {
        const double sT = d.conjT2 ? -1.0 : 1.0;
        for (int e = tid; e < SB_K1 * SB_P * SB_N1; e += blockDim.x) {
            const int k = e % (SB_P * SB_N1), k1 = e / (SB_P * SB_N1);
            const int p = k & 1, n1 = k >> 1;
            cplx v = B1[offB1_k1[k1] + offP[p] + offB1_n1[n1]];
            if (d.conjB1) v.y = -v.y;
            const double br = d.a2re * v.x - d.a2im * v.y, bi = d.a2re * v.y + d.a2im * v.x;
            const int jre = 8 * (k >> 2) + (k & 3), jim = jre + 4;
            double* r0 = B1s + (2 * k1) * SB_STRB1;
            double* r1 = r0 + SB_STRB1;
            r0[jre] = br; r0[jim] = -sT * bi;
            r1[jre] = bi; r1[jim] = sT * br;
        }
    }
    // B2' fragments of this warp's 8 columns c = 2 xi + p (x = 4 warp + xi), one double per (j, ks, lane):
    // B2'[real column 8 j + lrow][real k = 4 ks + lcol], dT = a1 * f(dC) * f(B2), contraction over n2
    double* b2w = B2f + warp * (2 * 8 * 32) + lane;
    {
        const double sC = d.conjC ? -1.0 : 1.0;
#pragma unroll 1
        for (int e = 0; e < 16; ++e) {
            const int j = e >> 3, ks = e & 7;
            const int c = 4 * j + (lrow >> 1), t = lrow & 1;
            const int x = SB_XW * warp + (c >> 1), p = c & 1;
            const int n2 = 4 * (ks >> 1) + lcol, prt = ks & 1;
            cplx v = B2[offB2_x[x] + offP[2 + p] + offB2_n2[n2]];
            if (d.conjB2) v.y = -v.y;
            const double br = d.a1re * v.x - d.a1im * v.y, bi = d.a1re * v.y + d.a1im * v.x;
            b2w[e * 32] = t == 0 ? (prt == 0 ? br : -sC * bi) : (prt == 0 ? bi : sC * br);
        }
    }
    __syncthreads();

    cplx* scr = scratch + warp * (SB_N1 * SB_TB * 8);
    // stage 3 operand selection (see gett_reduce): out[k1, n'] += sum_k fA(A[k, k1]) fT(dT[k, n'])
    const int prt3 = lcol & 1, kin = lcol >> 1, t3 = lrow & 1;
    const double sx = d.conjA ? -1.0 : 1.0, sy = d.conjT3 ? -1.0 : 1.0;
    const int ysel = (t3 == 0) ? prt3 : 1 - prt3;
    const double ysgn = (t3 == 0) ? (prt3 == 0 ? 1.0 : -sx * sy) : (prt3 == 0 ? sy : sx);
    const int ymask = ysgn < 0.0 ? (int)0x80000000 : 0;
    double acc3[2][8][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc3[i][j][0] = 0.0; acc3[i][j][1] = 0.0; }

    const long long ntiles = d.Msz / SB_TB;
    const long long sA0 = d.M.s0[0], sC0 = d.M.s1[0], sD0 = d.M.s2[0];

    // dC fragments of two consecutive n1 (rows m'' = lrow, k = n2): 8 x 128-bit loads per lane
    auto load_pair = [&](double (&a)[2][8], const cplx* row, int n1) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const cplx v = __ldg(row + offC_n1[n1 + u] + offC_n2[4 * q + lcol]);
                a[u][2 * q] = v.x; a[u][2 * q + 1] = v.y;
            }
    };
    // stage 1 for two n1: 4 accumulator chains, results to the scratch tile
    auto mma_pair = [&](const double (&a)[2][8], int n1) {
        double c[2][2][2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int j = 0; j < 2; ++j) { c[u][j][0] = 0.0; c[u][j][1] = 0.0; }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const double b0 = b2w[ks * 32], b1 = b2w[(8 + ks) * 32];
            dmma884(c[0][0][0], c[0][0][1], a[0][ks], b0);
            dmma884(c[1][0][0], c[1][0][1], a[1][ks], b0);
            dmma884(c[0][1][0], c[0][1][1], a[0][ks], b1);
            dmma884(c[1][1][0], c[1][1][1], a[1][ks], b1);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int j = 0; j < 2; ++j) scr[scr_idx(n1 + u, lrow, 4 * j + lcol)] = make_double2(c[u][j][0], c[u][j][1]);
    };

    double aA[2][8], aB[2][8];
    long long bA = 0, bC = 0, bD = 0;
    if ((long long)blockIdx.x < ntiles) {
        decomp3(d.M, (long long)blockIdx.x * SB_TB, bA, bC, bD);
        load_pair(aA, dC + bC + lrow * sC0, 0);
    }
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();   // keeps the four warps on the same tile: they all read the same dC rows (L1 reuse)
        const cplx* dCrow = dC + bC + lrow * sC0;
        const long long dArow = bD + lrow * sD0;
        const double* Ad = reinterpret_cast<const double*>(A + bA) + prt3;
        const bool more = tile + gridDim.x < ntiles;
        long long nA = 0, nC = 0, nD = 0;
        if (more) {
            // pull the NEXT tile into L2 while this one computes: its dC rows (256 lines of 128 B, 64 per warp) and this
            // warp's share of its A rows (4 x values x 16 k1 x 8 rows)
            decomp3(d.M, (tile + gridDim.x) * SB_TB, nA, nC, nD);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int line = 64 * warp + 32 * r + lane;        // (n1, n2) plane of the dC tile
                asm volatile("prefetch.global.L2 [%0];" ::"l"(dC + nC + offC_n1[line >> 4] + offC_n2[line & 15]));
                const int ak = 32 * r + lane;                        // (xi, k1 quad, row pair) of this warp's A rows
                asm volatile("prefetch.global.L2 [%0];" ::"l"(A + nA + offA_x[SB_XW * warp + (ak >> 4)] + offA_k1[4 * ((ak >> 2) & 3)] + 2 * (ak & 3) * sA0));
            }
        }
        // ---- stage 1: dT[m'', n1, c] for all n1; the fragments of the next pair of n1 are in flight under the DMMAs ----
#pragma unroll
        for (int pr = 0; pr < 8; pr += 2) {
            load_pair(aB, dCrow, 2 * (pr + 1));
            mma_pair(aA, 2 * pr);
            if (pr + 2 < 8) load_pair(aA, dCrow, 2 * (pr + 2));
            mma_pair(aB, 2 * (pr + 1));
        }
        __syncwarp();
        // ---- stages 2 and 3, two x values at a time.  The A fragments of stage 3 (the only consumer of A: every load is an
        // HBM / L2 access) are requested BEFORE the 128 DMMAs of stage 2, which hide their latency ----
#pragma unroll 1
        for (int xi0 = 0; xi0 < SB_XW; xi0 += 2) {
            double af[2][4][2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long ox = offA_x[SB_XW * warp + xi0 + u];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
                        af[u][jj][mt] = __ldg(Ad + 2 * ((2 * jj + kin) * sA0 + ox + offA_k1[8 * mt + lrow]));
            }
            // the first dC fragments of the next tile travel under the last 256 DMMAs of this one
            if (xi0 + 2 >= SB_XW && more) load_pair(aA, dC + nC + lrow * sC0, 0);
            // stage 2: dA[m'', k1, x] = sum over (p, n1): 8 accumulator chains, one B1' fragment feeds both x values
            double c2[2][4][2];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) { c2[u][nt][0] = 0.0; c2[u][nt][1] = 0.0; }
            const double* brow = B1s + lrow * SB_STRB1 + lcol;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int n1 = 2 * q + (lcol >> 1), p = lcol & 1;
                const cplx v0 = scr[scr_idx(n1, lrow, 2 * xi0 + p)];
                const cplx v1 = scr[scr_idx(n1, lrow, 2 * (xi0 + 1) + p)];
#pragma unroll
                for (int prt = 0; prt < 2; ++prt) {
                    const int ks = 2 * q + prt;
                    const double a0 = prt ? v0.y : v0.x, a1 = prt ? v1.y : v1.x;
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const double b = brow[nt * 8 * SB_STRB1 + 4 * ks];
                        dmma884(c2[0][nt][0], c2[0][nt][1], a0, b);
                        dmma884(c2[1][nt][0], c2[1][nt][1], a1, b);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long ox = offD_x[SB_XW * warp + xi0 + u];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    cplx* pd = dA + dArow + ox + offD_k1[4 * nt + lcol];
                    if (d.accumulate) {
                        // every element of dA belongs to exactly one thread of one launch and launches are ordered on the
                        // stream: these fire-and-forget reductions are deterministic and do not stall on the old value
                        atomicAdd(&pd->x, c2[u][nt][0]);
                        atomicAdd(&pd->y, c2[u][nt][1]);
                    } else {
                        __stcs(pd, make_double2(c2[u][nt][0], c2[u][nt][1]));
                    }
                }
            }
            // stage 3: dB1[k1, (p, n1)] += sum over (xi, m'') of fA(A) fT(dT); one k-step = rows m'' = 2 jj + kin
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int mrow = 2 * jj + kin;
                    double bf[8];
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        const int np = 4 * nt + (lrow >> 1), p = np & 1, n1 = np >> 1;
                        bf[nt] = flip_sign(reinterpret_cast<const double*>(scr + scr_idx(n1, mrow, 2 * (xi0 + u) + p))[ysel], ymask);
                    }
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        dmma884(acc3[0][nt][0], acc3[0][nt][1], af[u][jj][0], bf[nt]);
                        dmma884(acc3[1][nt][0], acc3[1][nt][1], af[u][jj][1], bf[nt]);
                    }
                }
        }
        __syncwarp();   // the scratch tile is rewritten by stage 1 of the next tile
        bA = nA; bC = nC; bD = nD;
    }

    // ---- leaf gradient: warps summed in a fixed order through shared memory, one partial per CTA ----
    __syncthreads();
    double* red = reinterpret_cast<double*>(scratch);     // [SB_WARPS][SB_NOUT][2]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int e = (8 * mt + lrow) * (SB_P * SB_N1) + 4 * nt + lcol;     // (k1, n' = p + 2 n1)
            red[(warp * SB_NOUT + e) * 2] = acc3[mt][nt][0];
            red[(warp * SB_NOUT + e) * 2 + 1] = acc3[mt][nt][1];
        }
    __syncthreads();
    cplx* mine = part + (long long)blockIdx.x * SB_NOUT;
    for (int e = tid; e < SB_NOUT; e += blockDim.x) {
        double re = 0.0, im = 0.0;
#pragma unroll
        for (int w = 0; w < SB_WARPS; ++w) { re += red[(w * SB_NOUT + e) * 2]; im += red[(w * SB_NOUT + e) * 2 + 1]; }
        if (d.part_add) { const cplx o = mine[e]; re += o.x; im += o.y; }
        mine[e] = make_double2(re, im);
    }
}

// dB1[k1, p, n1] += a3 * sum over the CTA partials.  Eight lanes share one element: lane s adds the partials g = s, s + 8, ...
// in order, the eight sub-sums are combined by a fixed shuffle tree -- the same association order in every run (deterministic).
__global__ void __launch_bounds__(256) site2_bwd_reduce(const __grid_constant__ Site2BwdDesc d, const cplx* __restrict__ part, int nparts,
                                                       cplx* __restrict__ dB1) {
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7;
    double re = 0.0, im = 0.0;
    if (e < SB_NOUT)
        for (int g = sub; g < nparts; g += 8) { const cplx v = part[(long long)g * SB_NOUT + e]; re += v.x; im += v.y; }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        re += __shfl_down_sync(0xffffffffu, re, o, 8);
        im += __shfl_down_sync(0xffffffffu, im, o, 8);
    }
    if (e >= SB_NOUT || sub != 0) return;
    const int k1 = e / (SB_P * SB_N1), np = e % (SB_P * SB_N1), p = np & 1, n1 = np >> 1;
    cplx* out = dB1 + legs4_off(d.K1g, k1, 0) + legs4_off(d.Pg, p, 0) + legs4_off(d.N1g, n1, 0);
    const cplx o = *out;
    *out = make_double2(o.x + d.a3re * re - d.a3im * im, o.y + d.a3re * im + d.a3im * re);
}

size_t site2b_smem() {
    return (size_t)(2 * SB_K1 * SB_STRB1 + SB_WARPS * 2 * 8 * 32) * sizeof(double) + (size_t)SB_WARPS * SB_N1 * SB_TB * 8 * sizeof(cplx) +
           (10 * 16 + 4) * sizeof(long long) + 128;
}

struct HL { int32_t label; int64_t size, s[6]; };   // strides in dC, B2, B1, A, dA, dB1 (0: absent)

int find_l(const std::vector<int32_t>& v, int32_t l) {
    for (size_t i = 0; i < v.size(); ++i) if (v[i] == l) return (int)i;
    return -1;
}

void strides_of(const TensorRef& t, std::vector<int64_t>& s) {
    if (!t.strides.empty()) { s = t.strides; return; }
    s.resize(t.dims.size());
    int64_t acc = 1;
    for (size_t i = 0; i < t.dims.size(); ++i) { s[i] = acc; acc *= t.dims[i]; }
}

bool fill4(Legs4& L, std::vector<HL> legs, int w0, int w1, int sort_by, int64_t want) {
    std::stable_sort(legs.begin(), legs.end(), [&](const HL& a, const HL& b) { return a.s[sort_by] < b.s[sort_by]; });
    if (legs.size() > 4) return false;
    int64_t tot = 1;
    L.n = (int)legs.size();
    for (int i = 0; i < L.n; ++i) { L.size[i] = (int)legs[i].size; L.s0[i] = legs[i].s[w0]; L.s1[i] = legs[i].s[w1]; tot *= legs[i].size; }
    return tot == want;
}

}  // namespace

// Classifies the reverse triple of a site; false if it does not have the instantiated bulk-site shape.
// Tensor order in HL.s: 0 dC, 1 B2, 2 B1, 3 A, 4 dA, 5 dB1.
bool site2b_plan(tne_ctx* ctx, const TensorRef& dC, const TensorRef& B2, const TensorRef& B1, const TensorRef& A, const TensorRef& dA,
                 const TensorRef& dB1, Site2BwdDesc* out) {
    (void)ctx;
    const TensorRef* T[6] = {&dC, &B2, &B1, &A, &dA, &dB1};
    std::vector<int64_t> st[6];
    for (int i = 0; i < 6; ++i) strides_of(*T[i], st[i]);
    if (A.labels.size() != dA.labels.size() || B1.labels.size() != dB1.labels.size()) return false;
    std::vector<int32_t> all = A.labels;
    for (auto l : B1.labels) if (find_l(all, l) < 0) all.push_back(l);
    for (auto l : B2.labels) if (find_l(all, l) < 0) all.push_back(l);
    for (auto l : dC.labels) if (find_l(all, l) < 0) return false;
    std::vector<HL> K1, X, P, N1, N2, M;
    for (auto l : all) {
        int idx[6];
        for (int i = 0; i < 6; ++i) idx[i] = find_l(T[i]->labels, l);
        if ((idx[3] >= 0) != (idx[4] >= 0) || (idx[2] >= 0) != (idx[5] >= 0)) return false;   // dA ~ A, dB1 ~ B1
        HL h;
        h.label = l;
        h.size = -1;
        for (int i = 0; i < 6; ++i) {
            h.s[i] = idx[i] >= 0 ? st[i][idx[i]] : 0;
            if (idx[i] >= 0) {
                if (h.size >= 0 && T[i]->dims[idx[i]] != h.size) return false;
                h.size = T[i]->dims[idx[i]];
            }
        }
        const int code = (idx[3] >= 0) | ((idx[2] >= 0) << 1) | ((idx[1] >= 0) << 2) | ((idx[0] >= 0) << 3);   // A, B1, B2, C
        switch (code) {
            case 0b0011: K1.push_back(h); break;
            case 0b0101: X.push_back(h); break;
            case 0b0110: P.push_back(h); break;
            case 0b1010: N1.push_back(h); break;
            case 0b1100: N2.push_back(h); break;
            case 0b1001: M.push_back(h); break;
            default: return false;
        }
    }
    Site2BwdDesc d;
    memset(&d, 0, sizeof(d));
    // s0 / s1 of each class: K1 (A, B1), X (A, B2), P (B1, B2), N1 (B1, dC), N2 (B2, dC); the *d copies carry dA strides,
    // the *g copies the strides of the leaf gradient dB1.  All copies of a class are sorted by the SAME key so that an
    // index means the same leg combination in every table.
    if (!fill4(d.K1, K1, 3, 2, 3, SB_K1) || !fill4(d.X, X, 3, 1, 3, SB_X) || !fill4(d.P, P, 2, 1, 2, SB_P) ||
        !fill4(d.N1, N1, 2, 0, 0, SB_N1) || !fill4(d.N2, N2, 1, 0, 0, SB_N2))
        return false;
    if (!fill4(d.K1d, K1, 4, 2, 3, SB_K1) || !fill4(d.Xd, X, 4, 1, 3, SB_X)) return false;
    if (!fill4(d.K1g, K1, 5, 2, 3, SB_K1) || !fill4(d.Pg, P, 5, 1, 2, SB_P) || !fill4(d.N1g, N1, 5, 0, 0, SB_N1)) return false;
    // boundary rows: sort by the A stride, merge legs adjacent in A, dA and dC
    std::stable_sort(M.begin(), M.end(), [](const HL& a, const HL& b) { return a.s[3] < b.s[3]; });
    std::vector<HL> Mm;
    for (auto& l : M) {
        if (l.size == 1) continue;
        if (!Mm.empty() && l.s[3] == Mm.back().s[3] * Mm.back().size && l.s[4] == Mm.back().s[4] * Mm.back().size &&
            l.s[0] == Mm.back().s[0] * Mm.back().size && Mm.back().size * l.size < (1LL << 30))
            Mm.back().size *= l.size;
        else
            Mm.push_back(l);
    }
    if (Mm.empty() || (int)Mm.size() > MAXLEG) return false;
    d.M.n = (int)Mm.size();
    d.Msz = 1;
    for (int i = 0; i < d.M.n; ++i) {
        d.M.size[i] = (int)Mm[i].size; d.M.s0[i] = Mm[i].s[3]; d.M.s1[i] = Mm[i].s[0]; d.M.s2[i] = Mm[i].s[4];
        d.M.shift[i] = -1;
        if ((Mm[i].size & (Mm[i].size - 1)) == 0) { int sh = 0; while ((1LL << sh) < Mm[i].size) ++sh; d.M.shift[i] = sh; }
        d.Msz *= Mm[i].size;
    }
    if (d.M.size[0] % SB_TB != 0 || d.Msz < 64) return false;
    *out = d;
    return true;
}

int64_t site2b_part_elems(tne_ctx* ctx) { return (int64_t)ctx->sm_count * 2 * SB_NOUT; }

// `part`: workspace of site2b_part_elems() complex numbers.  part_add: add this launch's CTA partials to what `part`
// holds (the variants of one site share a leaf gradient); finalize: reduce the partials into dB1 after the launch.
int site2b_run(tne_ctx* ctx, const Site2BwdDesc& d0, const cplx* dC, const cplx* B2, const cplx* B1, const cplx* A, cplx* dA, cplx* dB1,
               cplx* part, bool part_add, bool finalize) {
    ensure_smem_attr(ctx, gett_site2_bwd, (int)site2b_smem());
    Site2BwdDesc d = d0;
    d.part_add = part_add ? 1 : 0;
    const unsigned gx = (unsigned)ctx->sm_count * 2;    // FIXED grid: a partial slot belongs to one CTA index in every launch of a group
    gett_site2_bwd<<<gx, 32 * SB_WARPS, site2b_smem(), ctx->stream>>>(d, dC, B2, B1, A, dA, part);
    TNE_LAUNCH_CHECK(ctx);
    if (finalize) {
        site2_bwd_reduce<<<SB_NOUT * 8 / 256, 256, 0, ctx->stream>>>(d, part, (int)gx, dB1);
        TNE_LAUNCH_CHECK(ctx);
    }
    return TNE_OK;
}
