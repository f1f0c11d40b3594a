// Fused absorption of one PEPS site into the boundary:  C = (A . B1) . B2  in ONE kernel (FP64 DMMA).
//
// In the boundary sweep every site is two contraction steps (the bra tensor, then the ket tensor) whose intermediate
// T = A . B1 is twice as large as A and C and has exactly one consumer.  Run separately the two steps move
// |A| + 2 |T| + |C| = 6 |A| bytes; here a CTA takes 8 consecutive boundary rows m'', contracts them with B1
// (stage 1), keeps T in shared memory, contracts it with B2 (stage 2) and writes C: 2 |A| bytes, the same
// 8 * |K1| * |X| * |P| * |N1| (1 + |N2| / |K1|) flops.  Index classes (labels of A, B1, B2, C):
//     K1 = A, B1 (summed in stage 1)      X  = A, B2 (summed in stage 2)       P  = B1, B2 (physical; summed in stage 2)
//     N1 = B1, C                          N2 = B2, C                           M  = A, C (the boundary rows m'')
// Instantiated for the bulk site of a D = 4 PEPS: |K1| = |X| = |N1| = |N2| = 16, |P| = 2.  Both stages use the real
// embedding of tne_contract_dmma.cu (A rows read as interleaved reals, B' expanded in shared memory with the signs
// of the conjugations and alpha folded in); rows are ordered (x, m'') / (n1, m'') with m'' fastest, so that loads
// and stores of a warp instruction cover 8 consecutive boundary rows.
#include <algorithm>

#include "tne_gett.cuh"

namespace {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int S2_K1 = 16, S2_X = 16, S2_P = 2, S2_N1 = 16, S2_N2 = 16, S2_TB = 8;
constexpr int S2_STR1 = 2 * S2_K1 + 4;            // doubles per B1' row (2 K1 reals)
constexpr int S2_STR2 = 2 * S2_X * S2_P + 4;      // doubles per B2' row (2 X P reals)
constexpr int S2_TSTR = S2_X * S2_P + 4;          // complex per T row (576 B: conflict-free 128-bit fragment reads)

__device__ __forceinline__ long long legs4_off(const Legs4& L, int idx, int which) {
    long long o = 0;
    for (int i = 0; i < L.n; ++i) {
        int r = idx % L.size[i];
        idx /= L.size[i];
        o += r * (which == 0 ? L.s0[i] : L.s1[i]);
    }
    return o;
}

__global__ void __launch_bounds__(256, 2) gett_site2(const __grid_constant__ Site2Desc d, const cplx* __restrict__ A,
                                                     const cplx* __restrict__ B1, const cplx* __restrict__ B2, cplx* __restrict__ C) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* B1s = reinterpret_cast<double*>(smem_raw);                       // [2 P N1 = 64][S2_STR1]
    double* B2s = B1s + 2 * S2_P * S2_N1 * S2_STR1;                           // [2 N2 = 32][S2_STR2]
    cplx* Ts = reinterpret_cast<cplx*>(B2s + 2 * S2_N2 * S2_STR2);            // [N1 * TB = 128][S2_TSTR]
    long long* offA_k1 = reinterpret_cast<long long*>(Ts + S2_N1 * S2_TB * S2_TSTR);   // [16]
    long long* offA_x = offA_k1 + S2_K1;                                      // [16]
    long long* offC_n1 = offA_x + S2_X;                                       // [16]
    long long* offC_n2 = offC_n1 + S2_N1;                                     // [16]
    const int tid = threadIdx.x;
    if (tid < 16) {
        offA_k1[tid] = legs4_off(d.K1, tid, 0);
        offA_x[tid] = legs4_off(d.X, tid, 0);
        offC_n1[tid] = legs4_off(d.N1, tid, 1);
        offC_n2[tid] = legs4_off(d.N2, tid, 1);
    }
    // B1'[(j1 = p + 2 n1, t)][real index of k1]:  T = alpha1 * op(A) * op(B1)
    {
        const double sA = d.conjA ? -1.0 : 1.0;
        for (int e = tid; e < S2_P * S2_N1 * S2_K1; e += blockDim.x) {
            const int k = e % S2_K1, j1 = e / S2_K1;
            const int p = j1 & 1, n1 = j1 >> 1;
            cplx v = B1[legs4_off(d.K1, k, 1) + legs4_off(d.P, p, 0) + legs4_off(d.N1, n1, 0)];
            if (d.conjB1) v.y = -v.y;
            const double br = d.a1re * v.x - d.a1im * v.y, bi = d.a1re * v.y + d.a1im * v.x;
            const int jre = 8 * (k >> 2) + (k & 3), jim = jre + 4;
            double* r0 = B1s + (2 * j1) * S2_STR1;
            double* r1 = r0 + S2_STR1;
            r0[jre] = br; r0[jim] = -sA * bi;
            r1[jre] = bi; r1[jim] = sA * br;
        }
        // B2'[(n2, t)][real index of k2 = 2 x + p]:  C = alpha2 * T * op(B2)
        for (int e = tid; e < S2_N2 * S2_X * S2_P; e += blockDim.x) {
            const int k2 = e % (S2_X * S2_P), n2 = e / (S2_X * S2_P);
            const int p = k2 & 1, x = k2 >> 1;
            cplx v = B2[legs4_off(d.X, x, 1) + legs4_off(d.P, p, 1) + legs4_off(d.N2, n2, 0)];
            if (d.conjB2) v.y = -v.y;
            const double br = d.a2re * v.x - d.a2im * v.y, bi = d.a2re * v.y + d.a2im * v.x;
            const int jre = 8 * (k2 >> 2) + (k2 & 3), jim = jre + 4;
            double* r0 = B2s + (2 * n2) * S2_STR2;
            double* r1 = r0 + S2_STR2;
            r0[jre] = br; r0[jim] = -bi;
            r1[jre] = bi; r1[jim] = br;
        }
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int lrow = lane >> 2, lcol = lane & 3;
    const long long ntiles = d.g.Msz / S2_TB;
    long long outer_cached = -1, baseA = 0, baseC = 0;
    auto row_off = [&](long long tile, long long& oA, long long& oC) {   // offsets of this lane's boundary row m'' = 8 tile + lrow
        const long long row = tile * S2_TB + lrow;
        const int sh = d.g.M.shift[0];
        long long i0, outer;
        if (sh >= 0) { i0 = row & (long long)(d.g.M.size[0] - 1); outer = row >> sh; }
        else { outer = row / d.g.M.size[0]; i0 = row - outer * d.g.M.size[0]; }
        if (outer != outer_cached) {
            outer_cached = outer;
            long long a = 0, c = 0, idx = outer;
            for (int i = 1; i < d.g.M.n; ++i) {
                long long q, r;
                const int s2 = d.g.M.shift[i];
                if (s2 >= 0) { r = idx & (long long)(d.g.M.size[i] - 1); q = idx >> s2; }
                else { q = idx / d.g.M.size[i]; r = idx - q * d.g.M.size[i]; }
                a += r * d.g.M.s0[i]; c += r * d.g.M.s1[i];
                idx = q;
            }
            baseA = a; baseC = c;
        }
        oA = baseA + i0 * d.g.M.s0[0];
        oC = baseC + i0 * d.g.M.s1[0];
    };
    long long oA, oC;
    if ((long long)blockIdx.x < ntiles) row_off(blockIdx.x, oA, oC);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- stage 1: this warp takes x = 2 warp, 2 warp + 1; rows (x, m''), K1 = 16, columns (p, n1) = 32 ----
#pragma unroll
        for (int xi = 0; xi < 2; ++xi) {
            const int x = 2 * warp + xi;
            double a[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const cplx v = __ldcs(A + oA + offA_x[x] + offA_k1[4 * q + lcol]);
                a[2 * q] = v.x; a[2 * q + 1] = v.y;
            }
            double c[8][2];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { c[nt][0] = 0.0; c[nt][1] = 0.0; }
            const double* brow = B1s + lrow * S2_STR1 + lcol;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) dmma884(c[nt][0], c[nt][1], a[ks], brow[nt * 8 * S2_STR1 + 4 * ks]);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int j1 = 4 * nt + lcol, p = j1 & 1, n1 = j1 >> 1;
                Ts[(n1 * S2_TB + lrow) * S2_TSTR + 2 * x + p] = make_double2(c[nt][0], c[nt][1]);
            }
        }
        __syncthreads();
        const long long oCcur = oC;
        if (tile + gridDim.x < ntiles) {   // pull the next tile's A rows into L2 while stage 2 runs
            row_off(tile + gridDim.x, oA, oC);
#pragma unroll
            for (int xi = 0; xi < 2; ++xi)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(A + oA + offA_x[2 * warp + xi] + offA_k1[4 * q + lcol]));
        }
        // ---- stage 2: this warp takes n1 = 2 warp, 2 warp + 1; rows (n1, m''), K2 = (x, p) = 32, columns n2 = 16 ----
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
            const int n1 = 2 * warp + ni;
            double a[16];
            const cplx* trow = Ts + (n1 * S2_TB + lrow) * S2_TSTR + lcol;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const cplx v = trow[4 * q];
                a[2 * q] = v.x; a[2 * q + 1] = v.y;
            }
            double c[4][2];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) { c[nt][0] = 0.0; c[nt][1] = 0.0; }
            const double* brow = B2s + lrow * S2_STR2 + lcol;
#pragma unroll
            for (int ks = 0; ks < 16; ++ks)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(c[nt][0], c[nt][1], a[ks], brow[nt * 8 * S2_STR2 + 4 * ks]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                cplx* p = C + oCcur + offC_n1[n1] + offC_n2[4 * nt + lcol];
                cplx v = make_double2(c[nt][0], c[nt][1]);
                if (d.accumulate) { cplx o = *p; v.x += o.x; v.y += o.y; }
                __stcs(p, v);
            }
        }
        __syncthreads();
    }
}

size_t site2_smem() {
    return (size_t)(2 * S2_P * S2_N1 * S2_STR1 + 2 * S2_N2 * S2_STR2) * sizeof(double) + (size_t)S2_N1 * S2_TB * S2_TSTR * sizeof(cplx) +
           64 * sizeof(long long) + 128;
}

struct HL { int32_t label; int64_t size, s[4]; };   // strides in A, B1, B2, C (0: absent)

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

bool fill4(Legs4& L, std::vector<HL>& legs, int w0, int w1, int sort_by, int64_t want) {
    std::stable_sort(legs.begin(), legs.end(), [&](const HL& a, const HL& b) { return a.s[sort_by] < b.s[sort_by]; });
    if (legs.size() > 4) return false;
    int64_t tot = 1;
    L.n = (int)legs.size();
    for (int i = 0; i < L.n; ++i) { L.size[i] = (int)legs[i].size; L.s0[i] = legs[i].s[w0]; L.s1[i] = legs[i].s[w1]; tot *= legs[i].size; }
    return tot == want;
}

}  // namespace

// Classifies C = (A . B1) . B2; returns false if the pair does not have the instantiated bulk-site shape.
bool site2_plan(tne_ctx* ctx, const TensorRef& A, const TensorRef& B1, const TensorRef& B2, const TensorRef& C, bool accumulate,
                double a1re, double a1im, double a2re, double a2im, Site2Desc* out) {
    std::vector<int64_t> sa, s1, s2, sc;
    strides_of(A, sa); strides_of(B1, s1); strides_of(B2, s2); strides_of(C, sc);
    std::vector<HL> K1, X, P, N1, N2, M;
    std::vector<int32_t> all = A.labels;
    for (auto l : B1.labels) if (find_l(all, l) < 0) all.push_back(l);
    for (auto l : B2.labels) if (find_l(all, l) < 0) all.push_back(l);
    for (auto l : C.labels) if (find_l(all, l) < 0) return false;
    for (auto l : all) {
        const int ia = find_l(A.labels, l), i1 = find_l(B1.labels, l), i2 = find_l(B2.labels, l), ic = find_l(C.labels, l);
        HL h;
        h.label = l;
        h.size = ia >= 0 ? A.dims[ia] : (i1 >= 0 ? B1.dims[i1] : B2.dims[i2]);
        h.s[0] = ia >= 0 ? sa[ia] : 0; h.s[1] = i1 >= 0 ? s1[i1] : 0; h.s[2] = i2 >= 0 ? s2[i2] : 0; h.s[3] = ic >= 0 ? sc[ic] : 0;
        if ((i1 >= 0 && B1.dims[i1] != h.size) || (i2 >= 0 && B2.dims[i2] != h.size) || (ic >= 0 && C.dims[ic] != h.size)) return false;
        const int code = (ia >= 0) | ((i1 >= 0) << 1) | ((i2 >= 0) << 2) | ((ic >= 0) << 3);
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
    Site2Desc d;
    memset(&d, 0, sizeof(d));
    if (!fill4(d.K1, K1, 0, 1, 0, S2_K1) || !fill4(d.X, X, 0, 2, 0, S2_X) || !fill4(d.P, P, 1, 2, 1, S2_P) ||
        !fill4(d.N1, N1, 1, 3, 3, S2_N1) || !fill4(d.N2, N2, 2, 3, 3, S2_N2))
        return false;
    // boundary rows: sort by A stride, merge legs adjacent in A and in C
    std::stable_sort(M.begin(), M.end(), [](const HL& a, const HL& b) { return a.s[0] < b.s[0]; });
    std::vector<HL> Mm;
    for (auto& l : M) {
        if (l.size == 1) continue;
        if (!Mm.empty() && l.s[0] == Mm.back().s[0] * Mm.back().size && l.s[3] == Mm.back().s[3] * Mm.back().size && Mm.back().size * l.size < (1LL << 30))
            Mm.back().size *= l.size;
        else
            Mm.push_back(l);
    }
    if (Mm.empty() || (int)Mm.size() > MAXLEG) return false;
    d.g.M.n = (int)Mm.size();
    d.g.Msz = 1;
    for (int i = 0; i < d.g.M.n; ++i) {
        d.g.M.size[i] = (int)Mm[i].size; d.g.M.s0[i] = Mm[i].s[0]; d.g.M.s1[i] = Mm[i].s[3];
        d.g.M.shift[i] = -1;
        if ((Mm[i].size & (Mm[i].size - 1)) == 0) { int sh = 0; while ((1LL << sh) < Mm[i].size) ++sh; d.g.M.shift[i] = sh; }
        d.g.Msz *= Mm[i].size;
    }
    if (d.g.M.size[0] % S2_TB != 0 || d.g.Msz < 64) return false;
    d.conjA = A.conj; d.conjB1 = B1.conj; d.conjB2 = B2.conj; d.accumulate = accumulate;
    d.a1re = a1re; d.a1im = a1im; d.a2re = a2re; d.a2im = a2im;
    *out = d;
    return true;
}

int site2_run(tne_ctx* ctx, const Site2Desc& d, const cplx* A, const cplx* B1, const cplx* B2, cplx* C) {
    ensure_smem_attr(ctx, gett_site2, (int)site2_smem());
    const unsigned gx = (unsigned)std::min<long long>(d.g.Msz / S2_TB, (long long)ctx->sm_count * 2);
    gett_site2<<<gx, 256, site2_smem(), ctx->stream>>>(d, A, B1, B2, C);
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}
