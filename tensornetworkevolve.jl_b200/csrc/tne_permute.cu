// Tensor permutation (K1 of SURVEY.md 2c): C[perm(labels)] = op(A[labels]) for complex128 tensors of any rank.
// Replaces OMEinsum's `permutedims` (the unary einsum rule; call sites src/peps.jl:392-393 and the output permutation of a
// NestedEinsum root) where a permutation cannot be folded into the operand addressing of a contraction.
//
// HBM-bound, 16 algorithmic bytes read + 16 written per element.  One element is one 128-bit access.  The legs are split into
//   GA : the leading legs of A (unit stride first) up to 32 elements -- a warp reads 512 contiguous bytes,
//   GC : the leading legs of C not already in GA, up to 32 elements -- a warp writes 512 contiguous bytes,
//   outer legs: everything else, walked by a persistent grid.
// A CTA moves one (GA x GC) tile at a time through shared memory: loads with the GA index fastest (coalesced in A), stores with
// the GC index fastest (coalesced in C); rows are padded by one element so that both phases are bank-conflict free.  When A and C
// share their unit-stride leg no transposition is needed and the tile is copied with the GA index fastest on both sides.
// A leg that is too long for its group is split at a divisor (16384 = 32 x 512, 81 = 27 x 3).
#include <algorithm>

#include "tne_gett.cuh"

namespace {

constexpr int PM_MAXG = 32;                 // elements per group
constexpr int PM_THREADS = 256;

struct PermDesc {
    LegSet outer;                           // s0: A stride, s1: C stride
    long long nouter;
    int TA, TC;
    int nga, ngc;
    int gaSize[8], gcSize[8];
    long long gaA[8], gaC[8], gcA[8], gcC[8];
    int conj;
    int aligned;                            // A and C share the unit-stride leg: no transposition
};

__global__ void __launch_bounds__(PM_THREADS) permute_kernel(const __grid_constant__ PermDesc d, const cplx* __restrict__ A, cplx* __restrict__ C) {
    __shared__ cplx tile[PM_MAXG][PM_MAXG + 1];
    __shared__ long long offAa[PM_MAXG], offCa[PM_MAXG], offAc[PM_MAXG], offCc[PM_MAXG];
    const int tid = threadIdx.x;
    if (tid < d.TA) {
        long long a = 0, c = 0;
        int idx = tid;
        for (int i = 0; i < d.nga; ++i) { const int r = idx % d.gaSize[i]; idx /= d.gaSize[i]; a += r * d.gaA[i]; c += r * d.gaC[i]; }
        offAa[tid] = a; offCa[tid] = c;
    }
    if (tid >= 32 && tid < 32 + d.TC) {
        long long a = 0, c = 0;
        int idx = tid - 32;
        for (int i = 0; i < d.ngc; ++i) { const int r = idx % d.gcSize[i]; idx /= d.gcSize[i]; a += r * d.gcA[i]; c += r * d.gcC[i]; }
        offAc[tid - 32] = a; offCc[tid - 32] = c;
    }
    __syncthreads();
    const int n = d.TA * d.TC;
    const double sgn = d.conj ? -1.0 : 1.0;
    for (long long o = blockIdx.x; o < d.nouter; o += gridDim.x) {
        long long bA, bC;
        decomp2(d.outer, o, bA, bC);
        if (d.aligned) {
            for (int e = tid; e < n; e += PM_THREADS) {
                const int ia = e % d.TA, ic = e / d.TA;
                cplx v = __ldcs(A + bA + offAa[ia] + offAc[ic]);
                v.y *= sgn;
                __stcs(C + bC + offCa[ia] + offCc[ic], v);
            }
            continue;
        }
        for (int e = tid; e < n; e += PM_THREADS) {
            const int ia = e % d.TA, ic = e / d.TA;
            cplx v = __ldcs(A + bA + offAa[ia] + offAc[ic]);
            v.y *= sgn;
            tile[ic][ia] = v;
        }
        __syncthreads();
        for (int e = tid; e < n; e += PM_THREADS) {
            const int ic = e % d.TC, ia = e / d.TC;
            __stcs(C + bC + offCa[ia] + offCc[ic], tile[ic][ia]);
        }
        __syncthreads();
    }
}

struct PLeg { int64_t size, sa, sc; };

// takes legs (sorted by the group's stride) into a group of at most PM_MAXG elements; a leg that does not fit is split at
// its largest divisor that does, the remainder goes back to `rest`
void take_group(std::vector<PLeg>& sorted, std::vector<PLeg>& group, std::vector<PLeg>& rest) {
    int64_t prod = 1;
    size_t i = 0;
    for (; i < sorted.size(); ++i) {
        PLeg l = sorted[i];
        const int64_t cap = PM_MAXG / prod;
        if (cap <= 1) break;
        if (l.size <= cap) { group.push_back(l); prod *= l.size; continue; }
        int64_t div = 1;
        for (int64_t q = cap; q >= 2; --q) if (l.size % q == 0) { div = q; break; }
        if (div > 1) {
            group.push_back({div, l.sa, l.sc});
            rest.push_back({l.size / div, l.sa * div, l.sc * div});
            prod *= div;
            ++i;
        }
        break;
    }
    for (; i < sorted.size(); ++i) rest.push_back(sorted[i]);
}

}  // namespace

// true if (A, C) is a pure permutation (same labels, all distinct) -- then permute_launch applies
bool permute_applicable(const TensorRef& A, const TensorRef& C) {
    if (A.labels.size() != C.labels.size()) return false;
    for (size_t i = 0; i < A.labels.size(); ++i) {
        bool found = false;
        for (size_t j = 0; j < C.labels.size(); ++j)
            if (C.labels[j] == A.labels[i]) { found = C.dims[j] == A.dims[i]; break; }
        if (!found) return false;
        for (size_t j = i + 1; j < A.labels.size(); ++j) if (A.labels[j] == A.labels[i]) return false;
    }
    return true;
}

int permute_launch(tne_ctx* ctx, const TensorRef& A, const TensorRef& C) {
    std::vector<int64_t> sa = A.strides, sc = C.strides;
    if (sa.empty()) { sa.resize(A.dims.size()); int64_t acc = 1; for (size_t i = 0; i < A.dims.size(); ++i) { sa[i] = acc; acc *= A.dims[i]; } }
    if (sc.empty()) { sc.resize(C.dims.size()); int64_t acc = 1; for (size_t i = 0; i < C.dims.size(); ++i) { sc[i] = acc; acc *= C.dims[i]; } }
    std::vector<PLeg> legs;
    int64_t total = 1;
    for (size_t i = 0; i < A.labels.size(); ++i) {
        size_t j = 0;
        while (C.labels[j] != A.labels[i]) ++j;
        total *= A.dims[i];
        if (A.dims[i] > 1) legs.push_back({A.dims[i], sa[i], sc[j]});
    }
    if (total == 0) return TNE_OK;
    // merge legs adjacent in both tensors
    std::sort(legs.begin(), legs.end(), [](const PLeg& x, const PLeg& y) { return x.sa < y.sa; });
    std::vector<PLeg> m;
    for (auto& l : legs) {
        if (!m.empty() && l.sa == m.back().sa * m.back().size && l.sc == m.back().sc * m.back().size) m.back().size *= l.size;
        else m.push_back(l);
    }
    PermDesc d;
    memset(&d, 0, sizeof(d));
    d.conj = A.conj ? 1 : 0;
    std::vector<PLeg> ga, rest, gc, outer;
    take_group(m, ga, rest);                                             // m is sorted by the A stride
    std::sort(rest.begin(), rest.end(), [](const PLeg& x, const PLeg& y) { return x.sc < y.sc; });
    // aligned: the unit-stride leg of C is already in GA (it is A's unit-stride leg as well)
    int64_t min_sc_all = INT64_MAX, min_sc_ga = INT64_MAX;
    for (auto& l : ga) min_sc_ga = std::min(min_sc_ga, l.sc);
    for (auto& l : rest) min_sc_all = std::min(min_sc_all, l.sc);
    d.aligned = ga.empty() || rest.empty() || min_sc_ga < min_sc_all;
    take_group(rest, gc, outer);
    if (ga.size() > 8 || gc.size() > 8 || outer.size() > MAXLEG) return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "permutation has too many legs");
    d.TA = 1; d.TC = 1;
    d.nga = (int)ga.size(); d.ngc = (int)gc.size();
    for (int i = 0; i < d.nga; ++i) { d.gaSize[i] = (int)ga[i].size; d.gaA[i] = ga[i].sa; d.gaC[i] = ga[i].sc; d.TA *= (int)ga[i].size; }
    for (int i = 0; i < d.ngc; ++i) { d.gcSize[i] = (int)gc[i].size; d.gcA[i] = gc[i].sa; d.gcC[i] = gc[i].sc; d.TC *= (int)gc[i].size; }
    std::sort(outer.begin(), outer.end(), [](const PLeg& x, const PLeg& y) { return x.sa < y.sa; });
    d.outer.n = (int)outer.size();
    d.nouter = 1;
    for (int i = 0; i < d.outer.n; ++i) {
        d.outer.size[i] = (int)outer[i].size; d.outer.s0[i] = outer[i].sa; d.outer.s1[i] = outer[i].sc; d.outer.s2[i] = 0;
        d.outer.shift[i] = -1;
        if ((outer[i].size & (outer[i].size - 1)) == 0) { int sh = 0; while ((1LL << sh) < outer[i].size) ++sh; d.outer.shift[i] = sh; }
        if (outer[i].size >= (1LL << 31)) return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "permutation leg too long");
        d.nouter *= outer[i].size;
    }
    const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>(d.nouter, (long long)ctx->sm_count * 8));
    permute_kernel<<<gx, PM_THREADS, 0, ctx->stream>>>(d, A.ptr, C.ptr);
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}
