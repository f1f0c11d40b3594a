// Binary tensor contraction C (+)= A * B without TTGT permutation passes (K2/K3 of SURVEY.md 2c).
//
// Replaces OMEinsum's binary einsum (permutedims -> reshape -> batched zgemm -> reshape ->
// permutedims; call sites src/peps.jl:109,321,334,344,373,376,429 and einsum_grad from
// src/TensorAD/rules.jl:4-5).  Instead of materialising permuted copies, the labels are split
// into four classes -- M (A and C), N (B and C), K (A and B, summed) and batch (all three) --
// and an element's address is the sum of per-class offsets, each a mixed-radix dot product of
// the class index with that tensor's strides.  Kernels gather operand tiles straight from the
// tensors' own layouts ("GETT"); conjugation is fused into the operand load.
//
// Kernels
//   gett_generic : any shape; shared-memory tiles, 4x2 complex register blocking, FP64 FMA.
//   gett_absorb  : big A (M huge) x small B (K*N fits shared memory) -> big C, FP64 DMMA
//                  (mma.sync m8n8k4): A fragments stream from HBM through registers, B lives in
//                  shared memory, C fragments go straight back to HBM.  Forward boundary
//                  absorption and the dA = dC * B^H backward step have this shape.
//   gett_reduce  : big A x big B -> small C (M*N small, K huge), FP64 DMMA with split-K:
//                  the dB = A^H * dC leaf gradients and the closing dot product.
#include <algorithm>
#include <cmath>

#include "tne_internal.h"

#include "tne_gett.cuh"

// ---------------------------------------------------------------------------------------
// generic kernel: CTA tile TM x TN, K chunks of TK, 256 threads, 4 x 2 complex per thread
// ---------------------------------------------------------------------------------------
template <int TM, int TN, int TK>
__global__ void __launch_bounds__(256) gett_generic(const __grid_constant__ GettDesc d, const cplx* __restrict__ A,
                                                    const cplx* __restrict__ B, cplx* __restrict__ C, cplx* __restrict__ splitws) {
    constexpr int RM = TM / 16, RN = TN / 16;
    __shared__ long long offAm[TM], offCm[TM], offBn[TN], offCn[TN], offAk[TK], offBk[TK];
    __shared__ cplx As[TK][TM + 1];
    __shared__ cplx Bs[TK][TN + 1];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * TM;
    const long long n0 = (long long)blockIdx.y * TN;
    const long long z = blockIdx.z;
    const long long bt = z / d.nsplit;
    const int split = (int)(z - bt * d.nsplit);
    long long oAb, oBb, oCb;
    decomp3(d.Bt, bt, oAb, oBb, oCb);
    for (int i = tid; i < TM; i += 256) {
        long long a = 0, c = 0;
        if (m0 + i < d.Msz) decomp2(d.M, m0 + i, a, c);
        offAm[i] = a; offCm[i] = c;
    }
    for (int i = tid; i < TN; i += 256) {
        long long b = 0, c = 0;
        if (n0 + i < d.Nsz) decomp2(d.N, n0 + i, b, c);
        offBn[i] = b; offCn[i] = c;
    }
    double accr[RM][RN], acci[RM][RN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) { accr[i][j] = 0; acci[i][j] = 0; }
    const long long kbeg = (long long)split * d.kPerSplit;
    const long long kend = min(d.Ksz, kbeg + d.kPerSplit);
    const double sA = d.conjA ? -1.0 : 1.0, sB = d.conjB ? -1.0 : 1.0;
    for (long long k0 = kbeg; k0 < kend; k0 += TK) {
        __syncthreads();
        if (tid < TK) {
            long long a = 0, b = 0;
            if (k0 + tid < kend) decomp2(d.K, k0 + tid, a, b);
            offAk[tid] = a; offBk[tid] = b;
        }
        __syncthreads();
        for (int e = tid; e < TM * TK; e += 256) {
            int i, kk;
            if (d.aKfast) { kk = e % TK; i = e / TK; } else { i = e % TM; kk = e / TM; }
            cplx v = make_double2(0.0, 0.0);
            if (m0 + i < d.Msz && k0 + kk < kend) { v = A[oAb + offAm[i] + offAk[kk]]; v.y *= sA; }
            As[kk][i] = v;
        }
        for (int e = tid; e < TN * TK; e += 256) {
            int j, kk;
            if (d.bKfast) { kk = e % TK; j = e / TK; } else { j = e % TN; kk = e / TN; }
            cplx v = make_double2(0.0, 0.0);
            if (n0 + j < d.Nsz && k0 + kk < kend) { v = B[oBb + offBn[j] + offBk[kk]]; v.y *= sB; }
            Bs[kk][j] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < TK; ++kk) {
            cplx a[RM], b[RN];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < RN; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < RN; ++j) {
                    accr[i][j] = fma(a[i].x, b[j].x, accr[i][j]);
                    accr[i][j] = fma(-a[i].y, b[j].y, accr[i][j]);
                    acci[i][j] = fma(a[i].x, b[j].y, acci[i][j]);
                    acci[i][j] = fma(a[i].y, b[j].x, acci[i][j]);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        int mi = ty + 16 * i;
        if (m0 + mi >= d.Msz) continue;
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            int nj = tx + 16 * j;
            if (n0 + nj >= d.Nsz) continue;
            cplx* p = C + oCb + offCm[mi] + offCn[nj];
            const double vr = d.alphaRe * accr[i][j] - d.alphaIm * acci[i][j];
            const double vi = d.alphaRe * acci[i][j] + d.alphaIm * accr[i][j];
            if (d.nsplit > 1) {
                // split K: every split writes its tile to the workspace; gett_generic_final adds the splits in a fixed order
                // (no floating-point atomics: bit-reproducible results)
                const long long tile = ((long long)z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;   // z = bt * nsplit + split
                splitws[tile * (TM * TN) + mi * TN + nj] = make_double2(vr, vi);
            }
            else if (d.accumulate) { cplx o = *p; *p = make_double2(o.x + vr, o.y + vi); }
            else *p = make_double2(vr, vi);
        }
    }
}

// second pass of a split-K launch: C (+)= sum over the splits, in split order
template <int TM, int TN>
__global__ void __launch_bounds__(256) gett_generic_final(const __grid_constant__ GettDesc d, const cplx* __restrict__ splitws, cplx* __restrict__ C) {
    const long long m0 = (long long)blockIdx.x * TM, n0 = (long long)blockIdx.y * TN, bt = blockIdx.z;
    long long oAb, oBb, oCb;
    decomp3(d.Bt, bt, oAb, oBb, oCb);
    for (int e = threadIdx.x; e < TM * TN; e += blockDim.x) {
        const int mi = e / TN, nj = e - mi * TN;
        if (m0 + mi >= d.Msz || n0 + nj >= d.Nsz) continue;
        double re = 0.0, im = 0.0;
        for (int sp = 0; sp < d.nsplit; ++sp) {
            const long long tile = ((bt * d.nsplit + sp) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            const cplx v = splitws[tile * (TM * TN) + e];
            re += v.x; im += v.y;
        }
        long long a_, cm, b_, cn;
        decomp2(d.M, m0 + mi, a_, cm);
        decomp2(d.N, n0 + nj, b_, cn);
        cplx* p = C + oCb + cm + cn;
        if (d.accumulate) { const cplx o = *p; re += o.x; im += o.y; }
        *p = make_double2(re, im);
    }
}

// ---------------------------------------------------------------------------------------
// host side: classify labels, order and merge legs, pick a kernel
// ---------------------------------------------------------------------------------------
struct HLeg { int32_t label; int64_t size, sa, sb, sc; };

static void default_strides(const TensorRef& t, std::vector<int64_t>& s) {
    if (!t.strides.empty()) { s = t.strides; return; }
    s.resize(t.dims.size());
    int64_t acc = 1;
    for (size_t i = 0; i < t.dims.size(); ++i) { s[i] = acc; acc *= t.dims[i]; }
}

static int find_label(const std::vector<int32_t>& v, int32_t l) {
    for (size_t i = 0; i < v.size(); ++i) if (v[i] == l) return (int)i;
    return -1;
}

// merge legs adjacent in every tensor that sees them (stride_next == stride * size)
static void merge_legs(std::vector<HLeg>& legs) {
    std::vector<HLeg> out;
    for (auto& l : legs) {
        if (l.size == 1) continue;
        if (!out.empty()) {
            HLeg& p = out.back();
            bool ok = (l.sa == p.sa * p.size) && (l.sb == p.sb * p.size) && (l.sc == p.sc * p.size);
            if (ok && p.size * l.size < (1LL << 30)) { p.size *= l.size; continue; }
        }
        out.push_back(l);
    }
    legs.swap(out);
}

static int fill_legset(tne_ctx* ctx, LegSet& L, const std::vector<HLeg>& legs, int which0, int which1, int which2) {
    if ((int)legs.size() > MAXLEG) return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "contraction has more than %d legs in one class", MAXLEG);
    L.n = (int)legs.size();
    auto pick = [](const HLeg& l, int w) -> long long { return w == 0 ? l.sa : (w == 1 ? l.sb : (w == 2 ? l.sc : 0)); };
    for (int i = 0; i < L.n; ++i) {
        L.size[i] = (int)legs[i].size;
        L.shift[i] = -1;
        if ((legs[i].size & (legs[i].size - 1)) == 0) { int sh = 0; while ((1LL << sh) < legs[i].size) ++sh; L.shift[i] = sh; }
        L.s0[i] = pick(legs[i], which0);
        L.s1[i] = pick(legs[i], which1);
        L.s2[i] = pick(legs[i], which2);
    }
    return TNE_OK;
}

struct Classified {
    std::vector<HLeg> M, N, K, Bt;
    int64_t Msz = 1, Nsz = 1, Ksz = 1, Bsz = 1;
};

static int classify(tne_ctx* ctx, const TensorRef& A, const TensorRef& B, const TensorRef& C, Classified& cl) {
    std::vector<int64_t> sa, sb, sc;
    default_strides(A, sa); default_strides(B, sb); default_strides(C, sc);
    auto dup = [](const std::vector<int32_t>& v) {
        for (size_t i = 0; i < v.size(); ++i) for (size_t j = i + 1; j < v.size(); ++j) if (v[i] == v[j]) return true;
        return false;
    };
    if (dup(A.labels) || dup(B.labels) || dup(C.labels))
        return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "repeated labels inside one tensor (trace/diagonal) are not supported");
    std::vector<int32_t> all = A.labels;
    for (auto l : B.labels) if (find_label(all, l) < 0) all.push_back(l);
    for (auto l : C.labels) if (find_label(all, l) < 0)
        return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "output label %d does not appear in any input", (int)l);
    for (auto l : all) {
        int ia = find_label(A.labels, l), ib = find_label(B.labels, l), ic = find_label(C.labels, l);
        HLeg h;
        h.label = l;
        h.size = ia >= 0 ? A.dims[ia] : B.dims[ib];
        if (ia >= 0 && ib >= 0 && A.dims[ia] != B.dims[ib])
            return tne_fail(ctx, TNE_ERR_INVALID, "label %d has sizes %lld and %lld", (int)l, (long long)A.dims[ia], (long long)B.dims[ib]);
        if (ic >= 0 && C.dims[ic] != h.size)
            return tne_fail(ctx, TNE_ERR_INVALID, "label %d: output size %lld != input size %lld", (int)l, (long long)C.dims[ic], (long long)h.size);
        h.sa = ia >= 0 ? sa[ia] : 0;
        h.sb = ib >= 0 ? sb[ib] : 0;
        h.sc = ic >= 0 ? sc[ic] : 0;
        if (ia >= 0 && ib >= 0 && ic >= 0) cl.Bt.push_back(h);
        else if (ia >= 0 && ic >= 0) cl.M.push_back(h);
        else if (ib >= 0 && ic >= 0) cl.N.push_back(h);
        else cl.K.push_back(h);  // contracted, or summed out of one operand (other stride 0)
    }
    std::stable_sort(cl.M.begin(), cl.M.end(), [](const HLeg& x, const HLeg& y) { return x.sa < y.sa; });
    std::stable_sort(cl.N.begin(), cl.N.end(), [](const HLeg& x, const HLeg& y) { return x.sb < y.sb; });
    std::stable_sort(cl.K.begin(), cl.K.end(), [](const HLeg& x, const HLeg& y) {
        long long kx = x.sa ? x.sa : x.sb, ky = y.sa ? y.sa : y.sb;
        return kx < ky;
    });
    std::stable_sort(cl.Bt.begin(), cl.Bt.end(), [](const HLeg& x, const HLeg& y) { return x.sc < y.sc; });
    merge_legs(cl.M); merge_legs(cl.N); merge_legs(cl.K); merge_legs(cl.Bt);
    for (auto& l : cl.M) cl.Msz *= l.size;
    for (auto& l : cl.N) cl.Nsz *= l.size;
    for (auto& l : cl.K) cl.Ksz *= l.size;
    for (auto& l : cl.Bt) cl.Bsz *= l.size;
    return TNE_OK;
}

void contract_cost(const TensorRef& A, const TensorRef& B, const TensorRef& C, double* flops, double* bytes) {
    double vol = 1;
    std::vector<int32_t> seen;
    for (size_t i = 0; i < A.labels.size(); ++i) { vol *= (double)A.dims[i]; seen.push_back(A.labels[i]); }
    for (size_t i = 0; i < B.labels.size(); ++i) if (find_label(seen, B.labels[i]) < 0) vol *= (double)B.dims[i];
    *flops = 8.0 * vol;
    *bytes = 16.0 * ((double)A.nelem() + (double)B.nelem() + (double)C.nelem());
}


int contract_plan(tne_ctx* ctx, const TensorRef& A0, const TensorRef& B0, const TensorRef& C, bool accumulate,
                  double alpha_re, double alpha_im, ContractPlan* plan) {
    const TensorRef* A = &A0;
    const TensorRef* B = &B0;
    Classified cl;
    TNE_TRY(classify(ctx, *A, *B, C, cl));
    plan->swapped = false;
    if (cl.Nsz > cl.Msz) {  // keep the larger free side on M (grid.x)
        std::swap(A, B);
        plan->swapped = true;
        cl = Classified();
        TNE_TRY(classify(ctx, *A, *B, C, cl));
    }
    contract_cost(A0, B0, C, &plan->flops, &plan->bytes);
    plan->c_nelem = C.nelem();
    if (!C.strides.empty()) {  // memory extent of a strided / padded output (cleared before split-K atomics)
        int64_t ext = 1;
        for (size_t i = 0; i < C.dims.size(); ++i) ext += (C.dims[i] - 1) * C.strides[i];
        plan->c_nelem = C.nelem() == 0 ? 0 : ext;
    }
    plan->empty = (cl.Msz * cl.Nsz * cl.Bsz == 0);
    if (plan->empty) return TNE_OK;
    GettDesc& d = plan->d;
    memset(&d, 0, sizeof(d));
    TNE_TRY(fill_legset(ctx, d.M, cl.M, 0, 2, 3));
    TNE_TRY(fill_legset(ctx, d.N, cl.N, 1, 2, 3));
    TNE_TRY(fill_legset(ctx, d.K, cl.K, 0, 1, 3));
    TNE_TRY(fill_legset(ctx, d.Bt, cl.Bt, 0, 1, 2));
    d.Msz = cl.Msz; d.Nsz = cl.Nsz; d.Ksz = cl.Ksz; d.Bsz = cl.Bsz;
    d.conjA = A->conj; d.conjB = B->conj; d.accumulate = accumulate;
    d.alphaRe = alpha_re; d.alphaIm = alpha_im;
    auto unit_is_k = [](const std::vector<HLeg>& K, const std::vector<HLeg>& O, bool isA) {
        long long bestK = -1, bestO = -1;
        for (auto& l : K) { long long s = isA ? l.sa : l.sb; if (s > 0 && (bestK < 0 || s < bestK)) bestK = s; }
        for (auto& l : O) { long long s = isA ? l.sa : l.sb; if (s > 0 && (bestO < 0 || s < bestO)) bestO = s; }
        if (bestK < 0) return 0;
        if (bestO < 0) return 1;
        return bestK < bestO ? 1 : 0;
    };
    d.aKfast = unit_is_k(cl.K, cl.M, true);
    d.bKfast = unit_is_k(cl.K, cl.N, false);
    d.nsplit = 1;
    d.kPerSplit = d.Ksz;
    plan->kernel = 0;
    if (!ctx->opt_force_generic) {
        TNE_TRY(gett_fast_plan(ctx, plan));
        if (plan->kernel != 0) return TNE_OK;
        if (gemm_plan(ctx, plan)) return TNE_OK;     // general dense node: tiled DMMA GEMM (tne_gemm.cu)
    }
    constexpr int TM = 64, TN = 32, TK = 16;
    long long gm = (d.Msz + TM - 1) / TM, gn = (d.Nsz + TN - 1) / TN;
    long long tiles = gm * gn * d.Bsz;
    // split K when the output is too small to fill the chip and K is long
    if (tiles < 2LL * ctx->sm_count && d.Ksz >= 4096) {
        long long want = (4LL * ctx->sm_count + tiles - 1) / tiles;
        long long maxsplit = d.Ksz / 1024;
        long long ns = std::max<long long>(1, std::min(want, maxsplit));
        long long per = (d.Ksz + ns - 1) / ns;
        per = (per + TK - 1) / TK * TK;
        ns = (d.Ksz + per - 1) / per;
        // the splits go through a workspace of nsplit x (tiles x TM x TN) elements: keep it below 64 MiB
        while (ns > 1 && ns * tiles * TM * TN * (long long)sizeof(cplx) > (64LL << 20)) {
            ns = (ns + 1) / 2;
            per = (d.Ksz + ns - 1) / ns;
            per = (per + TK - 1) / TK * TK;
            ns = (d.Ksz + per - 1) / per;
        }
        d.nsplit = (int)ns;
        d.kPerSplit = per;
    }
    if (gn > 65535 || d.Bsz * d.nsplit > 65535)
        return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "contraction grid too large (N tiles %lld, batch*split %lld)", gn, d.Bsz * d.nsplit);
    plan->zero_first = false;   // split K is reduced by a second pass (gett_generic_final), which also honours `accumulate`
    plan->gx = (unsigned)gm; plan->gy = (unsigned)gn; plan->gz = (unsigned)(d.Bsz * d.nsplit);
    return TNE_OK;
}

static int contract_run_inner(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C);

int contract_run(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C) {
    if (plan.empty) return TNE_OK;
    if (!ctx->opt_profile) return contract_run_inner(ctx, plan, A, B, C);
    tne_ctx::ProfRec r;
    cudaEvent_t* ev[2] = {&r.e0, &r.e1};
    for (auto e : ev) {
        if (!ctx->event_pool.empty()) { *e = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
        else TNE_CUDA(ctx, cudaEventCreate(e));
    }
    r.cls = 8 * plan.kernel + ((plan.kernel == 1 || plan.kernel == 3) ? (plan.aux[0] >= 16 ? 4 : plan.aux[0] >= 8 ? 3 : plan.aux[0] >= 4 ? 2 : 1) : 0);
    r.flops = plan.flops; r.bytes = plan.bytes;
    TNE_CUDA(ctx, cudaEventRecord(r.e0, ctx->stream));
    int rc = contract_run_inner(ctx, plan, A, B, C);
    TNE_CUDA(ctx, cudaEventRecord(r.e1, ctx->stream));
    ctx->prof.push_back(r);
    return rc;
}

static int contract_run_inner(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C) {
    if (plan.swapped) std::swap(A, B);
    if (plan.kernel == 7) return gemm_run(ctx, plan, A, B, C);
    if (plan.kernel != 0) return gett_fast_run(ctx, plan, A, B, C);
    if (plan.zero_first) TNE_TRY(launch_fill_zero(ctx, C, plan.c_nelem));
    dim3 grid(plan.gx, plan.gy, plan.gz);
    if (plan.d.nsplit > 1) TNE_TRY(reduce_part_reserve(ctx, (int64_t)plan.gx * plan.gy * plan.gz * 64 * 32));
    gett_generic<64, 32, 16><<<grid, 256, 0, ctx->stream>>>(plan.d, A, B, C, ctx->red_part);
    TNE_LAUNCH_CHECK(ctx);
    if (plan.d.nsplit > 1) {
        dim3 g2(plan.gx, plan.gy, (unsigned)plan.d.Bsz);
        gett_generic_final<64, 32><<<g2, 256, 0, ctx->stream>>>(plan.d, ctx->red_part, C);
        TNE_LAUNCH_CHECK(ctx);
    }
    return TNE_OK;
}

static bool legs_equal(const LegSet& x, const LegSet& y) {
    if (x.n != y.n) return false;
    for (int i = 0; i < x.n; ++i)
        if (x.size[i] != y.size[i] || x.s0[i] != y.s0[i] || x.s1[i] != y.s1[i] || x.s2[i] != y.s2[i]) return false;
    return true;
}

bool contract_multi_compatible(const ContractPlan& x, const ContractPlan& y) {
    if (x.empty || y.empty || !gett_multi_capable(x) || !gett_multi_capable(y)) return false;
    if (x.swapped != y.swapped || x.aux[0] != y.aux[0] || x.aux[1] != y.aux[1]) return false;
    const GettDesc &a = x.d, &b = y.d;
    return a.Msz == b.Msz && a.Nsz == b.Nsz && a.Ksz == b.Ksz && a.Bsz == b.Bsz && legs_equal(a.M, b.M) && legs_equal(a.N, b.N) &&
           legs_equal(a.K, b.K) && legs_equal(a.Bt, b.Bt);
}

int contract_run_multi(tne_ctx* ctx, int T, const ContractPlan* const* plans, const cplx* const* As, const cplx* const* Bs, cplx* C) {
    if (T == 1) return contract_run(ctx, *plans[0], As[0], Bs[0], C);
    AbsorbMulti ops;
    ops.T = T;
    for (int t = 0; t < T; ++t) {
        const ContractPlan& p = *plans[t];
        ops.A[t] = p.swapped ? Bs[t] : As[t];
        ops.B[t] = p.swapped ? As[t] : Bs[t];
        ops.conjA[t] = p.d.conjA; ops.conjB[t] = p.d.conjB;
        ops.are[t] = p.d.alphaRe; ops.aim[t] = p.d.alphaIm;
    }
    if (!ctx->opt_profile) return gett_fast_run_multi(ctx, *plans[0], ops, C);
    tne_ctx::ProfRec r;
    cudaEvent_t* ev[2] = {&r.e0, &r.e1};
    for (auto e : ev) {
        if (!ctx->event_pool.empty()) { *e = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
        else TNE_CUDA(ctx, cudaEventCreate(e));
    }
    r.cls = 8 * 1 + 5 + (plans[0]->aux[0] >= 16 ? 1 : 0); r.flops = 0; r.bytes = 0;   // buckets 5, 6: fused multi-contribution launches
    for (int t = 0; t < T; ++t) { r.flops += plans[t]->flops; r.bytes += plans[t]->bytes - (t ? 16.0 * (double)plans[t]->c_nelem : 0.0); }
    TNE_CUDA(ctx, cudaEventRecord(r.e0, ctx->stream));
    int rc = gett_fast_run_multi(ctx, *plans[0], ops, C);
    TNE_CUDA(ctx, cudaEventRecord(r.e1, ctx->stream));
    ctx->prof.push_back(r);
    return rc;
}

int contract_launch(tne_ctx* ctx, const TensorRef& A, const TensorRef& B, const TensorRef& C, bool accumulate,
                    double alpha_re, double alpha_im) {
    ContractPlan plan;
    TNE_TRY(contract_plan(ctx, A, B, C, accumulate, alpha_re, alpha_im, &plan));
    return contract_run(ctx, plan, A.ptr, B.ptr, C.ptr);
}

#ifndef TNE_HAVE_DMMA
bool gett_multi_capable(const ContractPlan&) { return false; }
int gett_fast_run_multi(tne_ctx* ctx, const ContractPlan&, const AbsorbMulti&, cplx*) { return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "no fast path built"); }
int gett_fast_plan(tne_ctx*, ContractPlan*) { return TNE_OK; }
int gett_fast_run(tne_ctx* ctx, const ContractPlan&, const cplx*, const cplx*, cplx*) { return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "no fast path built"); }
#endif

// ---------------------------------------------------------------------------------------
// C ABI: einsum
// ---------------------------------------------------------------------------------------
// Repeated labels inside one operand (OMEinsum's trace / diagonal unary rules, e.g. ein"ii->"): the operand is read
// through its diagonal view -- one leg per distinct label whose stride is the sum of the strides of its occurrences.
static int diagonal_view(tne_ctx* ctx, TensorRef& t, const char* what) {
    std::vector<int64_t> st;
    default_strides(t, st);
    TensorRef v;
    v.ptr = t.ptr; v.conj = t.conj;
    for (size_t q = 0; q < t.labels.size(); ++q) {
        int at = find_label(v.labels, t.labels[q]);
        if (at < 0) { v.labels.push_back(t.labels[q]); v.dims.push_back(t.dims[q]); v.strides.push_back(st[q]); continue; }
        if (v.dims[at] != t.dims[q])
            return tne_fail(ctx, TNE_ERR_INVALID, "einsum: label %d repeats inside %s with sizes %lld and %lld", (int)t.labels[q], what,
                            (long long)v.dims[at], (long long)t.dims[q]);
        v.strides[at] += st[q];
    }
    if (v.labels.size() != t.labels.size()) t = v;
    return TNE_OK;
}

extern "C" int tne_einsum_sized(tne_ctx* ctx, int n_in, const int32_t* labels_flat, const int32_t* ranks, const tne_buf* in,
                                const int32_t* conj_flags, const int32_t* out_labels, int out_rank, int n_sizes,
                                const int32_t* size_labels, const int64_t* size_values, tne_buf* out) {
    TNE_ENTER(ctx);
    if (n_in != 1 && n_in != 2) return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "einsum takes 1 or 2 tensors (got %d); use tne_contract_tree", n_in);
    TensorRef T[2];
    int pos = 0;
    for (int i = 0; i < n_in; ++i) {
        Buf* b = buf_get(ctx, in[i]);
        if (!b) return TNE_ERR_INVALID;
        if ((int)b->dims.size() != ranks[i])
            return tne_fail(ctx, TNE_ERR_INVALID, "einsum: operand %d has rank %d but %d labels", i, (int)b->dims.size(), ranks[i]);
        T[i].ptr = b->ptr();
        T[i].dims = b->dims;
        T[i].labels.assign(labels_flat + pos, labels_flat + pos + ranks[i]);
        T[i].conj = conj_flags ? conj_flags[i] != 0 : false;
        pos += ranks[i];
        TNE_TRY(diagonal_view(ctx, T[i], "an operand"));
    }
    if (n_in == 1) {  // unary: permutation / summation == contraction with the constant 1
        T[1].ptr = ctx->one_dev;
    }
    TensorRef C;   // the output as allocated: one leg per entry of out_labels, repeats included
    C.labels.assign(out_labels, out_labels + out_rank);
    for (int i = 0; i < out_rank; ++i) {
        int32_t l = out_labels[i];
        int ia = find_label(T[0].labels, l), ib = find_label(T[1].labels, l);
        if (ia < 0 && ib < 0) {
            // a label that only the output carries (OMEinsum's repeat rule; the pullback of a sum / trace): its size comes
            // from the size dictionary and the first operand is broadcast along it (a leg of stride 0)
            int64_t sz = -1;
            for (int q = 0; q < n_sizes; ++q) if (size_labels[q] == l) sz = size_values[q];
            if (sz < 0) return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "einsum: output label %d is not in the inputs and has no size", (int)l);
            std::vector<int64_t> st;
            default_strides(T[0], st);
            T[0].strides = st;
            T[0].labels.push_back(l); T[0].dims.push_back(sz); T[0].strides.push_back(0);
            ia = (int)T[0].labels.size() - 1;
        }
        C.dims.push_back(ia >= 0 ? T[0].dims[ia] : T[1].dims[ib]);
    }
    tne_buf h;
    TNE_TRY(buf_new(ctx, C.dims, &h));
    C.ptr = buf_get(ctx, h)->ptr();
    TensorRef Cv = C;
    int rc = diagonal_view(ctx, Cv, "the output");
    if (rc == TNE_OK && Cv.labels.size() != C.labels.size())   // ein"i->ii": only the diagonal is written
        rc = launch_fill_zero(ctx, C.ptr, buf_get(ctx, h)->nelem);
    if (rc == TNE_OK) {
        // a pure permutation goes to the shared-memory tiled transpose kernel (tne_permute.cu), anything else is a contraction
        if (n_in == 1 && !ctx->opt_force_generic && !ctx->opt_no_permute && permute_applicable(T[0], Cv)) rc = permute_launch(ctx, T[0], Cv);
        else rc = contract_launch(ctx, T[0], T[1], Cv, false, 1.0, 0.0);
    }
    if (rc != TNE_OK) { tne_free(ctx, h); return rc; }
    *out = h;
    return TNE_OK;
}

extern "C" int tne_einsum(tne_ctx* ctx, int n_in, const int32_t* labels_flat, const int32_t* ranks, const tne_buf* in,
                          const int32_t* conj_flags, const int32_t* out_labels, int out_rank, tne_buf* out) {
    return tne_einsum_sized(ctx, n_in, labels_flat, ranks, in, conj_flags, out_labels, out_rank, 0, nullptr, nullptr, out);
}
