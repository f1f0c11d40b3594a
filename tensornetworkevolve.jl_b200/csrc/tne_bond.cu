// Bond update (K5 of SURVEY.md 2c): the device side of apply_onbond! (src/peps.jl:353-417) and of
// LinearAlgebra.svd + truncation (src/peps.jl:381-390).
//
// One CTA per bond job, many independent jobs per launch (a TEBD wavefront).  The reference forms the
// (up to 128 x 128) two-site matrix and calls LAPACK zgesdd; here the two site tensors are first
// orthogonalised against their outer legs (CGS2 QR of the O x (2 Ds) matricisation), so that the SVD runs on
// the small core  theta[(a,p'),(b,q')] = sum_{p,q,s} G[p'q',pq] R_i[a,(p,s)] R_j[b,(q,s)]  (at most 4Ds x 4Ds,
// 16 x 16 for D = 4) with a one-sided (Hestenes) Jacobi iteration in shared memory -- same singular values,
// U = Q_i u, conj(V) = Q_j conj(v).  Everything is complex128 FMA arithmetic; the matrices are far too small
// for the tensor pipe to matter and the step is latency-bound.
#include <algorithm>
#include <cmath>

#include <unordered_set>

#include "tne_internal.h"

namespace {

constexpr int BOND_THREADS = 256;
constexpr int MAXI = 32;          // 2 * Ds: columns of a site matricisation (shared bond Ds <= 16)
constexpr int MAXN = 2 * MAXI;    // rows / columns of the core theta
constexpr int QS_MAX = 1024;      // elements of one site matricisation kept in shared memory (16 KiB per side)

struct BondDev {
    const cplx* t[2];             // site tensors i, j
    int rank[2];
    int dims[2][TNE_MAX_RANK];
    int axis[2];                  // position of the shared bond
    const cplx* lambda[2][TNE_MAX_RANK];   // Vidal bond vectors per axis (null: none)
    double gate[32];              // 4 x 4 complex, column-major, M[o_i + 2 o_j, i_i + 2 i_j]
    int Dmax;
    double eps;
    int vidal;
    // scratch / outputs
    cplx* Q[2];                   // O x I (column-major): the matricisation, orthonormalised in place
    cplx* UV[2];                  // (2 O) x Dcap column-major: reshape(U, (sl..., D)) and reshape(conj(V), (sr..., D))
    cplx* S;                      // Dcap singular values (imaginary part 0)
    cplx* place[2];               // optional: the new site tensors themselves (bond leg = Dcap entries, site leg order);
                                  // used by the speculative sweep, which assumes kept == Dcap (no placement launches)
    int Dcap;
    int* kept;
    double* trunc_err;
};

__device__ __forceinline__ cplx cmulc(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cmulconj(cplx a, cplx b) {  // conj(a) * b
    return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ cplx csqrt_pos(cplx z) {  // principal square root
    if (z.y == 0.0) return z.x >= 0 ? make_double2(sqrt(z.x), 0.0) : make_double2(0.0, sqrt(-z.x));
    double r = hypot(z.x, z.y);
    double a = sqrt(0.5 * (r + fabs(z.x)));
    double b = 0.5 * z.y / a;
    return z.x >= 0 ? make_double2(a, b) : make_double2(fabs(b), copysign(a, z.y));
}
__device__ __forceinline__ cplx cinv(cplx z) {
    double dd = z.x * z.x + z.y * z.y;
    return make_double2(z.x / dd, -z.y / dd);
}
__device__ __forceinline__ cplx safe_inv_dev(cplx y) {  // src/peps.jl:419-426
    const double eps = 1e-10;
    if (hypot(y.x, y.y) < eps) y = make_double2(y.x >= 0 ? eps : -eps, 0.0);
    return cinv(y);
}

// block-wide sum of a complex value (all threads get the result)
__device__ cplx block_sum(cplx v, double* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    __syncthreads();
    if (lane == 0) { red[2 * warp] = v.x; red[2 * warp + 1] = v.y; }
    __syncthreads();
    cplx r = make_double2(0.0, 0.0);
    for (int w = 0; w < BOND_THREADS / 32; ++w) { r.x += red[2 * w]; r.y += red[2 * w + 1]; }
    return r;
}

// outer index o (the virtual legs other than the shared one, in tensor order) and (p, s) -> linear tensor index
struct SiteMap {
    int rank, ax, Ds, O;
    int dim[TNE_MAX_RANK];
    long long stride[TNE_MAX_RANK];
    __device__ void init(const int* dims, int rank_, int ax_) {
        rank = rank_; ax = ax_;
        long long s = 1;
        O = 1;
        for (int q = 0; q < rank; ++q) { dim[q] = dims[q]; stride[q] = s; s *= dims[q]; if (q != 0 && q != ax) O *= dims[q]; }
        Ds = dims[ax];
    }
    __device__ long long outer_offset(int o, int* leg_index /* may be null */) const {
        long long off = 0;
        for (int q = 1; q < rank; ++q) {
            if (q == ax) continue;
            int r = o % dim[q];
            o /= dim[q];
            off += r * stride[q];
            if (leg_index) leg_index[q] = r;
        }
        return off;
    }
};

// In-place QR of the O x I column-major matrix Q (shared or global memory) by classical Gram-Schmidt with one
// re-orthogonalisation pass ("twice is enough"); a column that is numerically in the span of the previous ones becomes
// zero with a zero pivot.  The triangular factor goes to Rm (leading dimension ld): Rm[k + c ld] = r_kc, or, with herm,
// its conjugate transpose Rm[c + k ld] = conj(r_kc).  Rm must be zeroed by the caller.  Block-wide; ends synchronised.
__device__ void cgs2(cplx* Q, int O, int I, cplx* Rm, int ld, bool herm, cplx* coef, double* red) {
    const int tid = threadIdx.x;
    for (int c = 0; c < I; ++c) {
        cplx* vc = Q + (long long)c * O;
        double nrm0 = 0.0;
        {
            cplx a = make_double2(0.0, 0.0);
            for (int o = tid; o < O; o += BOND_THREADS) { cplx x = vc[o]; a.x += x.x * x.x + x.y * x.y; }
            nrm0 = sqrt(block_sum(a, red).x);
        }
        for (int pass = 0; pass < 2; ++pass) {
            // coef[k] = <q_k, v> for k < c: warp w takes k = w, w + 8, ...
            const int lane = tid & 31, warp = tid >> 5;
            for (int k = warp; k < c; k += BOND_THREADS / 32) {
                const cplx* qk = Q + (long long)k * O;
                cplx a = make_double2(0.0, 0.0);
                for (int o = lane; o < O; o += 32) { cplx t = cmulconj(qk[o], vc[o]); a.x += t.x; a.y += t.y; }
                for (int sh = 16; sh > 0; sh >>= 1) { a.x += __shfl_xor_sync(0xffffffffu, a.x, sh); a.y += __shfl_xor_sync(0xffffffffu, a.y, sh); }
                if (lane == 0) coef[k] = a;
            }
            __syncthreads();
            for (int o = tid; o < O; o += BOND_THREADS) {
                cplx x = vc[o];
                for (int k = 0; k < c; ++k) { cplx t = cmulc(coef[k], Q[(long long)k * O + o]); x.x -= t.x; x.y -= t.y; }
                vc[o] = x;
            }
            if (tid < c) {
                cplx* dst = herm ? Rm + c + tid * ld : Rm + c * ld + tid;
                cplx r = *dst;
                r.x += coef[tid].x; r.y += herm ? -coef[tid].y : coef[tid].y;
                *dst = r;
            }
            __syncthreads();
        }
        cplx a = make_double2(0.0, 0.0);
        for (int o = tid; o < O; o += BOND_THREADS) { cplx x = vc[o]; a.x += x.x * x.x + x.y * x.y; }
        const double nrm = sqrt(block_sum(a, red).x);
        // a column that is (numerically) in the span of the previous ones contributes nothing
        const bool dead = !(nrm > 1e-13 * nrm0) || nrm0 == 0.0;
        const double inv = dead ? 0.0 : 1.0 / nrm;
        for (int o = tid; o < O; o += BOND_THREADS) { cplx x = vc[o]; vc[o] = make_double2(x.x * inv, x.y * inv); }
        if (tid == 0) Rm[c * ld + c] = make_double2(dead ? 0.0 : nrm, 0.0);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BOND_THREADS) bond_kernel(const BondDev* __restrict__ jobs) {
    extern __shared__ __align__(16) unsigned char smem[];
    const BondDev& J = jobs[blockIdx.x];
    const int tid = threadIdx.x;
    cplx* R[2];
    R[0] = reinterpret_cast<cplx*>(smem);                 // [MAXI][MAXI] column-major (I x I)
    R[1] = R[0] + MAXI * MAXI;
    cplx* W = R[1] + MAXI * MAXI;                         // theta, then W = theta * V : [ni][nj] column-major
    cplx* V = W + MAXN * MAXN;                            // [nj][nj]
    double* red = reinterpret_cast<double*>(V + MAXN * MAXN);   // 64 doubles
    double* sval = red + 64;                              // [MAXN]
    int* perm = reinterpret_cast<int*>(sval + MAXN);      // [MAXN]
    cplx* coef = reinterpret_cast<cplx*>(perm + MAXN);    // [MAXI]
    // the two matricisations live in shared memory when they fit (O * 2 Ds <= QS_MAX per side: every D <= 4 site of a
    // square lattice), else in the global scratch J.Q: the Gram-Schmidt loop is a chain of short dependent steps and
    // runs at memory latency
    cplx* Qs = coef + MAXI;                               // [2][QS_MAX]
    __shared__ int pidx[MAXN];                            // column order of the preconditioning QR
    __shared__ SiteMap map[2];
    __shared__ int s_flag, s_bad;
    if (tid == 0) s_bad = 0;
    if (tid < 2) map[tid].init(J.dims[tid], J.rank[tid], J.axis[tid]);
    __syncthreads();
    const long long t_start = clock64();

    // ---- 1. matricise (Vidal: absorb sqrt(lambda) on every virtual leg) and orthogonalise: T = Q R ----
    for (int side = 0; side < 2; ++side) {
        const SiteMap& M = map[side];
        const int O = M.O, Ds = M.Ds, I = 2 * Ds;
        cplx* Q = O * I <= QS_MAX ? Qs + side * QS_MAX : J.Q[side];
        for (int e = tid; e < O * I; e += BOND_THREADS) {
            const int o = e % O, c = e / O;       // column c = p + 2 s
            const int p = c & 1, s = c >> 1;
            int li[TNE_MAX_RANK];
            long long off = M.outer_offset(o, li) + p * M.stride[0] + (long long)s * M.stride[M.ax];
            cplx v = J.t[side][off];
            if (J.vidal) {
                li[M.ax] = s;
                for (int q = 1; q < M.rank; ++q)
                    if (J.lambda[side][q]) v = cmulc(v, csqrt_pos(J.lambda[side][q][li[q]]));
            }
            Q[(long long)c * O + o] = v;
        }
        for (int e = tid; e < MAXI * MAXI; e += BOND_THREADS) R[side][e] = make_double2(0.0, 0.0);
        __syncthreads();
        cgs2(Q, O, I, R[side], MAXI, false, coef, red);
    }

    const long long t_qr = clock64();
    // ---- 2. the core: theta[(a,p'),(b,q')] = sum_{p,q,s} G[p'+2q', p+2q] R_i[a,(p,s)] R_j[b,(q,s)] ----
    const int Ds = map[0].Ds;
    const int Ii = 2 * Ds, Ij = 2 * Ds;
    const int ni = 2 * Ii, nj = 2 * Ij;      // row index a + Ii * p' ... stored as (p' + 2 a) to follow (phys, outer) order
    for (int e = tid; e < ni * nj; e += BOND_THREADS) {
        const int row = e % ni, col = e / ni;
        const int pp = row & 1, a = row >> 1, qq = col & 1, b = col >> 1;
        cplx acc = make_double2(0.0, 0.0);
        for (int s = 0; s < Ds; ++s)
            for (int p = 0; p < 2; ++p) {
                const cplx ri = R[0][(p + 2 * s) * MAXI + a];
                if (ri.x == 0.0 && ri.y == 0.0) continue;
                for (int q = 0; q < 2; ++q) {
                    const int gi = (pp + 2 * qq) + 4 * (p + 2 * q);
                    const cplx g = make_double2(J.gate[2 * gi], J.gate[2 * gi + 1]);
                    cplx t = cmulc(cmulc(g, ri), R[1][(q + 2 * s) * MAXI + b]);
                    acc.x += t.x; acc.y += t.y;
                }
            }
        W[col * MAXN + row] = acc;
    }
    __syncthreads();
    // ---- 2b. preconditioning (Drmac-Veselic): theta P = Q1 Rt with the columns sorted by decreasing norm, then Jacobi on
    // X = Rt^H instead of theta.  The Rydberg gate at small dt makes theta strongly graded (sigma_16 / sigma_1 ~ 1e-9) and the
    // edge bonds make it rank deficient: plain one-sided Jacobi needs 12-13 sweeps (up to the cap of 40 for the
    // rank-deficient ones), on X it needs 5-7.  theta = (Q1 Vj) Sigma (P X' / Sigma)^H with X' = X Vj the rotated columns.
    // Needs ni, nj <= MAXI (Ds <= 8): Q1 lives in V's buffer, X and Vj in the two halves of W's buffer.
    const bool precond = ni <= MAXI && nj <= MAXI;
    cplx* jw = W;           // the matrix whose columns the Jacobi rotations orthogonalise (jr rows)
    cplx* jv = V;           // the accumulated rotations (nj rows)
    int jr = ni;
    if (precond) {
        for (int c = tid; c < nj; c += BOND_THREADS) {
            double a = 0.0;
            for (int r = 0; r < ni; ++r) { cplx u = W[c * MAXN + r]; a += u.x * u.x + u.y * u.y; }
            sval[c] = isfinite(a) ? a : 0.0;   // keeps the rank sort below a total order (pidx is always a permutation)
        }
        __syncthreads();
        for (int c = tid; c < nj; c += BOND_THREADS) {
            int rank = 0;
            for (int k = 0; k < nj; ++k) if (sval[k] > sval[c] || (sval[k] == sval[c] && k < c)) ++rank;
            pidx[rank] = c;
        }
        __syncthreads();
        cplx* Q1 = V;                                   // ni x nj, contiguous columns of ni elements
        for (int e = tid; e < ni * nj; e += BOND_THREADS) { const int r = e % ni, c = e / ni; Q1[c * ni + r] = W[pidx[c] * MAXN + r]; }
        __syncthreads();
        cplx* X = W;                                    // nj x nj, leading dimension MAXN
        cplx* Vj = W + MAXI * MAXN;
        for (int e = tid; e < nj * nj; e += BOND_THREADS) {
            const int r = e % nj, c = e / nj;
            X[c * MAXN + r] = make_double2(0.0, 0.0);
            Vj[c * MAXN + r] = make_double2(r == c ? 1.0 : 0.0, 0.0);
        }
        __syncthreads();
        cgs2(Q1, ni, nj, X, MAXN, true, coef, red);
        jw = X; jv = Vj; jr = nj;
    } else {
        for (int e = tid; e < nj * nj; e += BOND_THREADS) {
            const int r = e % nj, c = e / nj;
            V[c * MAXN + r] = make_double2(r == c ? 1.0 : 0.0, 0.0);
        }
        __syncthreads();
    }

    int n_sweeps = 0;
    const long long t_core = clock64();
    // ---- 3. one-sided Jacobi: rotate column pairs of W (and V) until all pairs are orthogonal ----
    {
        const int lane = tid & 31, warp = tid >> 5, nwarp = BOND_THREADS / 32;
        const int npad = nj + (nj & 1);
        for (int sweep = 0; sweep < 40; ++sweep) {
            n_sweeps = sweep + 1;
            int again = 0;      // block-uniform: some pair of this sweep was still off by more than the sweep tolerance
            for (int step = 0; step < npad - 1; ++step) {
                int need = 0;
                // round-robin tournament: npad / 2 disjoint pairs per step
                for (int pr = warp; pr < npad / 2; pr += nwarp) {
                    int x = pr == 0 ? npad - 1 : (step + pr) % (npad - 1);
                    int y = (step + npad - 1 - pr) % (npad - 1);
                    if (pr == 0) y = step % (npad - 1);
                    int p = min(x, y), q = max(x, y);
                    if (q >= nj || p == q) continue;
                    cplx* wp = jw + p * MAXN;
                    cplx* wq = jw + q * MAXN;
                    double al = 0.0, be = 0.0;
                    cplx ga = make_double2(0.0, 0.0);
                    for (int r = lane; r < jr; r += 32) {
                        cplx u = wp[r], v = wq[r];
                        al += u.x * u.x + u.y * u.y;
                        be += v.x * v.x + v.y * v.y;
                        cplx t = cmulconj(u, v);
                        ga.x += t.x; ga.y += t.y;
                    }
                    for (int sh = 16; sh > 0; sh >>= 1) {
                        al += __shfl_xor_sync(0xffffffffu, al, sh);
                        be += __shfl_xor_sync(0xffffffffu, be, sh);
                        ga.x += __shfl_xor_sync(0xffffffffu, ga.x, sh);
                        ga.y += __shfl_xor_sync(0xffffffffu, ga.y, sh);
                    }
                    // the scalar part of a rotation is a chain of dependent FP64 divisions and square roots: it sets the
                    // latency of a Jacobi step, so it is written with reciprocal square roots (2 rsqrt, 1 sqrt, 1 division)
                    const double g2 = ga.x * ga.x + ga.y * ga.y, ab = al * be;
                    if (!(g2 > 1e-30 * ab) || ab == 0.0) continue;   // also skips a pair with an exactly vanishing (or underflowing) column
                    // every pair above rounding level is still rotated, but only a pair that is off by more than 1e-12
                    // asks for another sweep: Jacobi converges quadratically, so the rotations of this sweep already
                    // bring such pairs to ~1e-24
                    if (g2 > 1e-24 * ab) need = 1;
                    const double ginv = rsqrt(g2);
                    // unitary rotation [c, s e^{i phi}; -s e^{-i phi}, c] that zeroes <w_p', w_q'>
                    const double zeta = 0.5 * (be - al) * ginv;
                    const double w = 1.0 + zeta * zeta;
                    const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(w));   // sqrt, not w * rsqrt(w): w may overflow to inf (t -> 0)
                    const double c = rsqrt(1.0 + t * t), s = c * t;
                    const cplx ph = make_double2(ga.x * ginv, ga.y * ginv);   // e^{i phi}, gamma = <w_p, w_q>
                    for (int r = lane; r < jr; r += 32) {
                        cplx u = wp[r], v = wq[r];
                        cplx vph = cmulconj(ph, v);            // e^{-i phi} v
                        cplx uph = cmulc(ph, u);               // e^{ i phi} u
                        wp[r] = make_double2(c * u.x - s * vph.x, c * u.y - s * vph.y);
                        wq[r] = make_double2(s * uph.x + c * v.x, s * uph.y + c * v.y);
                    }
                    cplx* vp = jv + p * MAXN;
                    cplx* vq = jv + q * MAXN;
                    for (int r = lane; r < nj; r += 32) {
                        cplx u = vp[r], v = vq[r];
                        cplx vph = cmulconj(ph, v);
                        cplx uph = cmulc(ph, u);
                        vp[r] = make_double2(c * u.x - s * vph.x, c * u.y - s * vph.y);
                        vq[r] = make_double2(s * uph.x + c * v.x, s * uph.y + c * v.y);
                    }
                }
                again |= __syncthreads_or(need);    // the barrier between two tournament steps doubles as the vote
            }
            if (!again) break;
        }
    }

    const long long t_jac = clock64();
    // ---- 4. singular values, descending order, truncation (src/peps.jl:383-390) ----
    for (int c = tid; c < nj; c += BOND_THREADS) {
        double a = 0.0;
        for (int r = 0; r < jr; ++r) { cplx u = jw[c * MAXN + r]; a += u.x * u.x + u.y * u.y; }
        // a non-finite input (diverged state, or garbage handed to a speculative wavefront) must not break the index
        // bookkeeping: it counts as a zero singular value here and is reported through trunc_err = NaN below
        if (!isfinite(a)) { a = 0.0; s_bad = 1; }
        sval[c] = sqrt(a);
    }
    __syncthreads();
    // Wf[:, c] = sigma_c u_c (ni rows), Vf[:, c] = v_c (nj rows)
    const cplx* Wf = W;
    const cplx* Vf = V;
    int ldw = MAXN, ldv = MAXN;
    if (precond) {
        const cplx* Q1 = V;
        cplx* wf = R[0];                                // the R factors of the sites are no longer needed
        cplx* vf = R[1];
        for (int e = tid; e < ni * nj; e += BOND_THREADS) {          // sigma u = sigma Q1 Vj[:, c]
            const int r = e % ni, c = e / ni;
            cplx acc = make_double2(0.0, 0.0);
            for (int k = 0; k < nj; ++k) { cplx t = cmulc(Q1[k * ni + r], jv[c * MAXN + k]); acc.x += t.x; acc.y += t.y; }
            wf[c * MAXI + r] = make_double2(acc.x * sval[c], acc.y * sval[c]);
        }
        for (int e = tid; e < nj * nj; e += BOND_THREADS) {          // v = P X'[:, c] / sigma_c
            const int r = e % nj, c = e / nj;
            const double inv = sval[c] > 0.0 ? 1.0 / sval[c] : 0.0;
            const cplx x = jw[c * MAXN + r];
            vf[c * MAXI + pidx[r]] = make_double2(x.x * inv, x.y * inv);
        }
        Wf = wf; Vf = vf; ldw = MAXI; ldv = MAXI;
        __syncthreads();
    }
    for (int c = tid; c < nj; c += BOND_THREADS) {
        int rank = 0;
        for (int k = 0; k < nj; ++k) if (sval[k] > sval[c] || (sval[k] == sval[c] && k < c)) ++rank;
        perm[rank] = c;
    }
    __syncthreads();
    const int mi = 2 * map[0].O, mj = 2 * map[1].O;
    const int full = min(mi, mj);             // size(U, 2) of the reference's thin SVD
    const int ncomp = min(nj, full);          // singular values beyond this are exactly zero
    if (tid == 0) {
        int D0 = -1;
        for (int k = 0; k < ncomp; ++k) if (sval[perm[k]] < J.eps) { D0 = k; break; }
        if (D0 < 0 && full > ncomp && 0.0 < J.eps) D0 = ncomp;
        int D = D0 < 0 ? min(J.Dmax, full) : min(D0, J.Dmax);
        D = min(D, ncomp);
        double err = 0.0;
        for (int k = D; k < ncomp; ++k) err += sval[perm[k]];
        *J.kept = D;
        // diagnostics in the spare words of the job's meta record (read back with option debug >= 2)
        J.kept[1] = n_sweeps; J.kept[2] = (int)((t_qr - t_start) >> 10); J.kept[3] = (int)((t_jac - t_core) >> 10);
        *J.trunc_err = D < full ? err : -1.0;   // -1: nothing was cut (the reference logs only when D < size(U,2))
        if (s_bad || !isfinite(err)) *J.trunc_err = nan("");   // non-finite input: the host turns this into an error status
        s_flag = D;
    }
    __syncthreads();
    const int D = s_flag;

    // ---- 5. U = Q_i u, conj(V) = Q_j conj(v); Simple: sqrt(S) into both; Vidal: divide the outer legs ----
    for (int side = 0; side < 2; ++side) {
        const SiteMap& M = map[side];
        const int O = M.O, I = 2 * M.Ds;
        const cplx* Q = O * I <= QS_MAX ? Qs + side * QS_MAX : J.Q[side];
        cplx* out = J.UV[side];
        const int rows = 2 * O;
        // columns D .. Dcap-1 (a bond that kept fewer singular values than its cap) are written as zeros: later wavefronts
        // of a speculative sweep read these buffers before the exact-rank redo replaces them
        for (int e = tid + rows * D; e < rows * J.Dcap; e += BOND_THREADS) {
            const int row = e % rows, dd = e / rows;
            out[(long long)dd * rows + row] = make_double2(0.0, 0.0);
            if (J.place[side]) {
                long long off = row & 1, st = M.dim[0];
                int oo = row >> 1;
                for (int q = 1; q < M.rank; ++q) {
                    if (q == M.ax) { off += (long long)dd * st; st *= J.Dcap; }
                    else { off += (long long)(oo % M.dim[q]) * st; oo /= M.dim[q]; st *= M.dim[q]; }
                }
                J.place[side][off] = make_double2(0.0, 0.0);
            }
        }
        for (int e = tid; e < rows * D; e += BOND_THREADS) {
            const int row = e % rows, dd = e / rows;
            const int pp = row & 1, o = row >> 1;
            const int col = perm[dd];
            const double s = sval[col];
            cplx acc = make_double2(0.0, 0.0);
            if (side == 0) {
                for (int a = 0; a < I; ++a) { cplx t = cmulc(Q[(long long)a * O + o], Wf[col * ldw + pp + 2 * a]); acc.x += t.x; acc.y += t.y; }
                const double inv = s > 0.0 ? 1.0 / s : 0.0;   // u = W / s
                acc.x *= inv; acc.y *= inv;
            } else {
                for (int b = 0; b < I; ++b) {
                    cplx v = Vf[col * ldv + pp + 2 * b];
                    v.y = -v.y;
                    cplx t = cmulc(Q[(long long)b * O + o], v);
                    acc.x += t.x; acc.y += t.y;
                }
            }
            if (J.vidal) {
                int li[TNE_MAX_RANK];
                M.outer_offset(o, li);
                for (int q = 1; q < M.rank; ++q)
                    if (q != M.ax && J.lambda[side][q]) acc = cmulc(acc, safe_inv_dev(csqrt_pos(J.lambda[side][q][li[q]])));
            } else {
                const double sq = sqrt(s);
                acc.x *= sq; acc.y *= sq;
            }
            out[(long long)dd * rows + row] = acc;
            if (J.place[side]) {
                // the same element in the site tensor's own layout (column-major over dims with dims[ax] = Dcap)
                long long off = pp, st = M.dim[0];
                int oo = o;
                for (int q = 1; q < M.rank; ++q) {
                    if (q == M.ax) { off += (long long)dd * st; st *= J.Dcap; }
                    else { off += (long long)(oo % M.dim[q]) * st; oo /= M.dim[q]; st *= M.dim[q]; }
                }
                J.place[side][off] = acc;
            }
        }
    }
    for (int dd = tid; dd < J.Dcap; dd += BOND_THREADS) J.S[dd] = make_double2(dd < D ? sval[perm[dd]] : 0.0, 0.0);
}

size_t bond_smem_bytes() {
    return (size_t)(2 * MAXI * MAXI + 2 * MAXN * MAXN + MAXI + 2 * QS_MAX) * sizeof(cplx) + (64 + MAXN) * sizeof(double) + MAXN * sizeof(int) + 64;
}

struct HostJob {
    tne_buf q[2] = {0, 0}, uv[2] = {0, 0}, s = 0;
    int O[2] = {1, 1};
    std::vector<int64_t> dims[2];
    int axis[2] = {0, 0};
};

}  // namespace

// A launched batch of bond jobs whose results (kept rank, truncation error) have not been read back yet.
struct BondBatch {
    std::vector<HostJob> hj;
    std::vector<BondDev> dev;      // stays alive until the batch has been fetched (source of an async H2D copy)
    tne_buf hmeta = 0, hdev = 0;
    int n = 0;
    std::vector<tne_buf> slabs;    // speculative sweeps: the scratch of all jobs is carved out of a few big blocks
};

// builds the device descriptors and launches the bond kernel for `n_jobs` site-disjoint jobs; does not synchronise
// defer: only build the descriptors (bb.dev); the caller uploads and launches them later (tne_tebd_sweep: one upload per sweep)
static int bond_enqueue(tne_ctx* ctx, int n_jobs, const tne_bond_job* in, BondBatch& bb, std::vector<tne_buf>* direct = nullptr, bool defer = false) {
    ensure_smem_attr(ctx, bond_kernel, (int)bond_smem_bytes());
    std::vector<BondDev>& dev = bb.dev;
    std::vector<HostJob>& hj = bb.hj;
    dev.assign(n_jobs, BondDev());
    hj.assign(n_jobs, HostJob());
    bb.n = n_jobs;
    tne_buf& hmeta = bb.hmeta;
    TNE_TRY(buf_new(ctx, {(int64_t)(2 * n_jobs + 1)}, &hmeta));   // kept (as int in .x slot) and error per job
    char* meta = reinterpret_cast<char*>(buf_get(ctx, hmeta)->ptr());
    cplx* slab_ptr = nullptr;
    int64_t slab_left = 0;
    bb.slabs.clear();
    for (int k = 0; k < n_jobs; ++k) {
        const tne_bond_job& jb = in[k];
        BondDev& d = dev[k];
        memset(&d, 0, sizeof(d));
        const tne_buf tb[2] = {jb.ti, jb.tj};
        const int ax[2] = {jb.axis_i, jb.axis_j};
        for (int side = 0; side < 2; ++side) {
            Buf* b = buf_get(ctx, tb[side]);
            if (!b) return TNE_ERR_INVALID;
            const int rk = (int)b->dims.size();
            if (rk < 2 || ax[side] < 1 || ax[side] >= rk) return tne_fail(ctx, TNE_ERR_INVALID, "bond job %d: bad shared axis", k);
            if (b->dims[0] != 2) return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "bond job %d: the physical leg must come first and have dimension 2", k);
            d.t[side] = b->ptr();
            d.rank[side] = rk;
            d.axis[side] = ax[side];
            int O = 1;
            for (int q = 0; q < rk; ++q) { d.dims[side][q] = (int)b->dims[q]; if (q != 0 && q != ax[side]) O *= (int)b->dims[q]; }
            hj[k].O[side] = O;
            hj[k].dims[side] = b->dims;
            hj[k].axis[side] = ax[side];
            if (jb.vidal) {
                const tne_buf* lam = side == 0 ? jb.lambda_i : jb.lambda_j;
                for (int q = 1; q < rk; ++q) {
                    if (!lam[q]) return tne_fail(ctx, TNE_ERR_INVALID, "bond job %d: missing bond vector for axis %d", k, q);
                    Buf* lb = buf_get(ctx, lam[q]);
                    if (!lb || lb->nelem != b->dims[q]) return tne_fail(ctx, TNE_ERR_INVALID, "bond job %d: bond vector of axis %d has the wrong length", k, q);
                    d.lambda[side][q] = lb->ptr();
                }
            }
        }
        const int Ds = d.dims[0][ax[0]];
        if (Ds != d.dims[1][ax[1]]) return tne_fail(ctx, TNE_ERR_INVALID, "bond job %d: shared bond sizes differ", k);
        if (2 * Ds > MAXI) return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "bond job %d: bond dimension %d > %d", k, Ds, MAXI / 2);
        memcpy(d.gate, jb.gate, sizeof(d.gate));
        d.Dmax = jb.Dmax; d.eps = jb.eps; d.vidal = jb.vidal;
        d.Dcap = std::min(4 * Ds, std::min(2 * hj[k].O[0], 2 * hj[k].O[1]));
        d.Dcap = std::max(1, std::min(d.Dcap, std::max(1, (int)jb.Dmax)));
        if (direct) {
            // scratch from a bump allocator over big blocks: one pool operation per ~4 MiB instead of five per gate
            auto carve = [&](int64_t elems, cplx** out) -> int {
                elems = (elems + 15) / 16 * 16;                       // 256-byte granules
                if (slab_left < elems) {
                    tne_buf hs;
                    const int64_t chunk = std::max<int64_t>(elems, 1 << 18);
                    TNE_TRY(buf_new(ctx, {chunk}, &hs));
                    bb.slabs.push_back(hs);
                    slab_ptr = buf_get(ctx, hs)->ptr();
                    slab_left = chunk;
                }
                *out = slab_ptr;
                slab_ptr += elems; slab_left -= elems;
                return TNE_OK;
            };
            for (int side = 0; side < 2; ++side) {
                TNE_TRY(carve((int64_t)hj[k].O[side] * 2 * Ds, &d.Q[side]));
                TNE_TRY(carve((int64_t)2 * hj[k].O[side] * d.Dcap, &d.UV[side]));
            }
        } else {
            for (int side = 0; side < 2; ++side) {
                TNE_TRY(buf_new(ctx, {(int64_t)hj[k].O[side] * 2 * Ds}, &hj[k].q[side]));
                TNE_TRY(buf_new(ctx, {(int64_t)2 * hj[k].O[side] * d.Dcap}, &hj[k].uv[side]));
                d.Q[side] = buf_get(ctx, hj[k].q[side])->ptr();
                d.UV[side] = buf_get(ctx, hj[k].uv[side])->ptr();
            }
            TNE_TRY(buf_new(ctx, {(int64_t)d.Dcap}, &hj[k].s));
            d.S = buf_get(ctx, hj[k].s)->ptr();
        }
        if (direct) {   // speculative sweep: allocate the final tensors now (bond leg = Dcap) and let the kernel fill them
            for (int side = 0; side < 2; ++side) {
                std::vector<int64_t> od = hj[k].dims[side];
                od[hj[k].axis[side]] = d.Dcap;
                tne_buf ho;
                TNE_TRY(buf_new(ctx, od, &ho));
                direct->push_back(ho);
                d.place[side] = buf_get(ctx, ho)->ptr();
            }
            tne_buf hs;
            TNE_TRY(buf_new(ctx, {(int64_t)d.Dcap}, &hs));
            direct->push_back(hs);
            d.S = buf_get(ctx, hs)->ptr();
        }
        d.kept = reinterpret_cast<int*>(meta + 16 * (2 * k));
        d.trunc_err = reinterpret_cast<double*>(meta + 16 * (2 * k + 1));
    }
    if (defer) return TNE_OK;
    tne_buf& hdev = bb.hdev;
    const int64_t dev_elems = ((int64_t)n_jobs * (int64_t)sizeof(BondDev) + 15) / 16;
    TNE_TRY(buf_new(ctx, {dev_elems}, &hdev));
    TNE_CUDA(ctx, cudaMemcpyAsync(buf_get(ctx, hdev)->ptr(), dev.data(), (size_t)n_jobs * sizeof(BondDev), cudaMemcpyHostToDevice, ctx->stream));
    bond_kernel<<<n_jobs, BOND_THREADS, bond_smem_bytes(), ctx->stream>>>(reinterpret_cast<const BondDev*>(buf_get(ctx, hdev)->ptr()));
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}

// reads back (kept, truncation error) of a launched batch: the only host synchronisation of a bond update
static int bond_fetch(tne_ctx* ctx, BondBatch& bb, std::vector<int>& kept, std::vector<double>& err) {
    const int n_jobs = bb.n;
    tne_buf& hmeta = bb.hmeta;
    tne_buf& hdev = bb.hdev;
    const char* meta = reinterpret_cast<const char*>(buf_get(ctx, hmeta)->ptr());
    std::vector<char> hmeta_host((size_t)32 * n_jobs);
    TNE_CUDA(ctx, cudaMemcpyAsync(hmeta_host.data(), meta, hmeta_host.size(), cudaMemcpyDeviceToHost, ctx->stream));
    TNE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // also guarantees `dev` was consumed
    kept.resize(n_jobs); err.resize(n_jobs);
    for (int k = 0; k < n_jobs; ++k) {
        kept[k] = *reinterpret_cast<int*>(hmeta_host.data() + 32 * k);
        if (ctx->opt_debug > 1) {
            const int* w = reinterpret_cast<const int*>(hmeta_host.data() + 32 * k);
            fprintf(stderr, "[tne-bond] job %d: kept %d, %d Jacobi sweeps, QR %d kclk, Jacobi %d kclk\n", k, w[0], w[1], w[2], w[3]);
        }
        err[k] = *reinterpret_cast<double*>(hmeta_host.data() + 32 * k + 16);
    }
    if (hdev) tne_free(ctx, hdev);
    hdev = 0;
    tne_free(ctx, hmeta);
    hmeta = 0;
    return TNE_OK;
}

// reshape(U, (sl..., D)) -> the site tensor in its own leg order (src/peps.jl:392-393)
static int bond_place(tne_ctx* ctx, tne_buf uv, const std::vector<int64_t>& dims, int axis, int D, tne_buf* out) {
    const int rk = (int)dims.size();
    TensorRef A, one, C;
    A.ptr = buf_get(ctx, uv)->ptr();
    std::vector<int64_t> odims = dims;
    odims[axis] = D;
    for (int q = 0; q < rk; ++q) if (q != axis) { A.labels.push_back(q); A.dims.push_back(dims[q]); }
    A.labels.push_back(axis); A.dims.push_back(D);
    for (int q = 0; q < rk; ++q) { C.labels.push_back(q); C.dims.push_back(odims[q]); }
    one.ptr = ctx->one_dev;
    TNE_TRY(buf_new(ctx, odims, out));
    C.ptr = buf_get(ctx, *out)->ptr();
    return contract_launch(ctx, A, one, C, false);
}

extern "C" int tne_apply_onbond_batched(tne_ctx* ctx, int n_jobs, tne_bond_job* jobs) {
    TNE_ENTER(ctx);
    if (n_jobs <= 0) return TNE_OK;
    if (!jobs) return tne_fail(ctx, TNE_ERR_INVALID, "null jobs");
    BondBatch bb;
    std::vector<int> kept;
    std::vector<double> err;
    TNE_TRY(bond_enqueue(ctx, n_jobs, jobs, bb));
    TNE_TRY(bond_fetch(ctx, bb, kept, err));
    std::vector<HostJob>& hj = bb.hj;
    for (int k = 0; k < n_jobs; ++k) {
        const int D = kept[k];
        TNE_TRY(bond_place(ctx, hj[k].uv[0], hj[k].dims[0], hj[k].axis[0], D, &jobs[k].ti_out));
        TNE_TRY(bond_place(ctx, hj[k].uv[1], hj[k].dims[1], hj[k].axis[1], D, &jobs[k].tj_out));
        TNE_TRY(buf_new(ctx, {(int64_t)D}, &jobs[k].s_out));
        if (D > 0)
            TNE_CUDA(ctx, cudaMemcpyAsync(buf_get(ctx, jobs[k].s_out)->ptr(), buf_get(ctx, hj[k].s)->ptr(), (size_t)D * sizeof(cplx),
                                          cudaMemcpyDeviceToDevice, ctx->stream));
        jobs[k].kept = D;
        jobs[k].trunc_err = err[k];
        for (int side = 0; side < 2; ++side) { tne_free(ctx, hj[k].q[side]); tne_free(ctx, hj[k].uv[side]); }
        tne_free(ctx, hj[k].s);
    }
    for (int k = 0; k < n_jobs; ++k)
        if (std::isnan(err[k])) return tne_fail(ctx, TNE_ERR_INVALID, "bond job %d: non-finite values in the site tensors", k);
    return TNE_OK;
}

// Whole TEBD sweeps (src/tebd.jl:15-23, and the nstep - 1 of them of src/tebd.jl:25-38) in one call.  The reference applies the
// gates strictly one after the other; gates on disjoint sites commute exactly (truncation included), so the sequence is levelled
// into wavefronts of site-disjoint gates (a gate's level = 1 + the highest level of an earlier gate on one of its sites) and every
// wavefront is one launch.  The rank a bond keeps is data dependent (src/peps.jl:383-390), but a generic gate always truncates to
// the cap min(Dmax, 4 Ds, 2 O_i, 2 O_j): ALL sweeps are therefore enqueued SPECULATIVELY with kept = cap, with no host
// synchronisation between wavefronts or between sweeps, and the kept ranks are checked once at the end.  If a bond of sweep k
// kept fewer (singular values below eps), the speculative results from sweep k on are dropped, sweep k is redone wavefront by
// wavefront with the exact ranks (the inputs are immutable buffers, so nothing has to be restored) and the rest is enqueued again.
namespace {

struct SweepCtx {
    tne_ctx* ctx;
    int n_sites, n_bonds, n_gates, Dmax, vidal, n_levels;
    double eps;
    const tne_sweep_gate* gates;
    std::vector<std::vector<int>> by_level;
};

struct SweepRun {                        // one sweep whose kernels are enqueued and whose ranks are not verified yet
    std::vector<BondBatch> batches;      // per level: descriptors + meta buffer
    std::vector<std::vector<int>> cap;
    std::vector<tne_buf> created;        // every tensor the sweep allocated (final ones and intermediate generations)
    std::vector<tne_buf> in_s, in_b, out_s, out_b;
    std::vector<BondDev> all_dev;
    tne_buf hall = 0;
};

void fill_job(const SweepCtx& S, const tne_sweep_gate& g, const std::vector<tne_buf>& cur_s, const std::vector<tne_buf>& cur_b, tne_bond_job& jb) {
    memset(&jb, 0, sizeof(jb));
    jb.ti = cur_s[g.site_i]; jb.tj = cur_s[g.site_j];
    jb.axis_i = g.axis_i; jb.axis_j = g.axis_j;
    memcpy(jb.gate, g.gate, sizeof(jb.gate));
    jb.Dmax = S.Dmax; jb.eps = S.eps; jb.vidal = S.vidal;
    if (S.vidal)
        for (int a = 1; a < TNE_MAX_RANK; ++a) {
            if (g.lambda_i[a] >= 0 && g.lambda_i[a] < S.n_bonds) jb.lambda_i[a] = cur_b[g.lambda_i[a]];
            if (g.lambda_j[a] >= 0 && g.lambda_j[a] < S.n_bonds) jb.lambda_j[a] = cur_b[g.lambda_j[a]];
        }
}

void free_batches(tne_ctx* ctx, SweepRun& R) {
    for (auto& bb : R.batches) { if (bb.hmeta) tne_free(ctx, bb.hmeta); if (bb.hdev) tne_free(ctx, bb.hdev); bb.hmeta = bb.hdev = 0; }
    if (R.hall) { tne_free(ctx, R.hall); R.hall = 0; }
}

// enqueue one sweep speculatively (kept = cap): descriptors of all wavefronts in ONE upload, then the launches back to back
int sweep_enqueue(const SweepCtx& S, const std::vector<tne_buf>& in_s, const std::vector<tne_buf>& in_b, SweepRun& R) {
    tne_ctx* ctx = S.ctx;
    R.in_s = in_s; R.in_b = in_b;
    std::vector<tne_buf> cur_s = in_s, cur_b = in_b, scratch;
    R.batches.assign(S.n_levels, BondBatch());
    R.cap.assign(S.n_levels, std::vector<int>());
    int rc = TNE_OK;
    for (int lv = 0; lv < S.n_levels && rc == TNE_OK; ++lv) {
        const std::vector<int>& ids = S.by_level[lv];
        std::vector<tne_bond_job> jobs(ids.size());
        for (size_t q = 0; q < ids.size(); ++q) fill_job(S, S.gates[ids[q]], cur_s, cur_b, jobs[q]);
        BondBatch& bb = R.batches[lv];
        std::vector<tne_buf> direct;
        rc = bond_enqueue(ctx, (int)jobs.size(), jobs.data(), bb, &direct, true);
        for (auto h : direct) R.created.push_back(h);
        // the launches are deferred: the scratch of this level must not go back to the pool before its kernel is enqueued
        // (a later buf_new -- the descriptor table itself -- could be handed the same block)
        for (auto hs : bb.slabs) scratch.push_back(hs);
        bb.slabs.clear();
        if (rc) break;
        R.cap[lv].resize(ids.size());
        for (size_t q = 0; q < ids.size(); ++q) {
            const tne_sweep_gate& g = S.gates[ids[q]];
            R.cap[lv][q] = bb.dev[q].Dcap;
            cur_s[g.site_i] = direct[3 * q]; cur_s[g.site_j] = direct[3 * q + 1];
            if (S.vidal) cur_b[g.bond] = direct[3 * q + 2];
        }
    }
    if (rc == TNE_OK) {
        std::vector<size_t> off(S.n_levels + 1, 0);
        for (int lv = 0; lv < S.n_levels; ++lv) off[lv + 1] = off[lv] + R.batches[lv].dev.size();
        R.all_dev.reserve(off[S.n_levels]);
        for (int lv = 0; lv < S.n_levels; ++lv) R.all_dev.insert(R.all_dev.end(), R.batches[lv].dev.begin(), R.batches[lv].dev.end());
        rc = buf_new(ctx, {(int64_t)((R.all_dev.size() * sizeof(BondDev) + 15) / 16)}, &R.hall);
        if (rc == TNE_OK) {
            BondDev* base = reinterpret_cast<BondDev*>(buf_get(ctx, R.hall)->ptr());
            cudaError_t e = cudaMemcpyAsync(base, R.all_dev.data(), R.all_dev.size() * sizeof(BondDev), cudaMemcpyHostToDevice, ctx->stream);
            if (e != cudaSuccess) rc = tne_fail(ctx, TNE_ERR_CUDA, "descriptor upload failed: %s", cudaGetErrorString(e));
            for (int lv = 0; lv < S.n_levels && rc == TNE_OK; ++lv) {
                bond_kernel<<<(unsigned)R.batches[lv].dev.size(), BOND_THREADS, bond_smem_bytes(), ctx->stream>>>(base + off[lv]);
                ctx->launches++;
                e = cudaGetLastError();
                if (e != cudaSuccess) rc = tne_fail(ctx, TNE_ERR_CUDA, "bond kernel launch failed: %s", cudaGetErrorString(e));
            }
        }
    }
    for (auto h : scratch) tne_free(ctx, h);   // stream-ordered pool: safe once the kernels that use them are enqueued
    R.out_s = cur_s; R.out_b = cur_b;
    return rc;
}

// read the ranks of an enqueued sweep back (synchronises); ok = every bond reached its cap and nothing was non-finite
int sweep_verify(const SweepCtx& S, SweepRun& R, int32_t* kept_out, double* err_out, bool* ok) {
    *ok = true;
    for (int lv = 0; lv < S.n_levels; ++lv) {
        std::vector<int> kept;
        std::vector<double> err;
        TNE_TRY(bond_fetch(S.ctx, R.batches[lv], kept, err));
        for (size_t q = 0; q < S.by_level[lv].size(); ++q) {
            const int gi = S.by_level[lv][q];
            kept_out[gi] = kept[q]; err_out[gi] = err[q];
            if (kept[q] != R.cap[lv][q] || std::isnan(err[q])) *ok = false;   // NaN: decided by the exact-rank pass
        }
    }
    if (R.hall) { tne_free(S.ctx, R.hall); R.hall = 0; }
    return TNE_OK;
}

// one sweep wavefront by wavefront with the exact ranks (one synchronisation per wavefront)
int sweep_exact(const SweepCtx& S, const std::vector<tne_buf>& in_s, const std::vector<tne_buf>& in_b, std::vector<tne_buf>& out_s,
                std::vector<tne_buf>& out_b, std::vector<tne_buf>& created, int32_t* kept_out, double* err_out) {
    tne_ctx* ctx = S.ctx;
    std::vector<tne_buf> cur_s = in_s, cur_b = in_b;
    int rc = TNE_OK;
    for (int lv = 0; lv < S.n_levels && rc == TNE_OK; ++lv) {
        const std::vector<int>& ids = S.by_level[lv];
        std::vector<tne_bond_job> jobs(ids.size());
        for (size_t q = 0; q < ids.size(); ++q) fill_job(S, S.gates[ids[q]], cur_s, cur_b, jobs[q]);
        BondBatch bb;
        rc = bond_enqueue(ctx, (int)jobs.size(), jobs.data(), bb);
        if (rc) break;
        std::vector<int> kept;
        std::vector<double> err;
        rc = bond_fetch(ctx, bb, kept, err);
        if (rc) break;
        for (size_t q = 0; q < ids.size() && rc == TNE_OK; ++q) {
            const tne_sweep_gate& g = S.gates[ids[q]];
            HostJob& h = bb.hj[q];
            const int D = kept[q];
            kept_out[ids[q]] = kept[q]; err_out[ids[q]] = err[q];
            tne_buf ti = 0, tj = 0, sb = 0;
            rc = bond_place(ctx, h.uv[0], h.dims[0], h.axis[0], D, &ti);
            if (!rc) { created.push_back(ti); rc = bond_place(ctx, h.uv[1], h.dims[1], h.axis[1], D, &tj); }
            if (!rc) { created.push_back(tj); rc = buf_new(ctx, {(int64_t)D}, &sb); }
            if (!rc) {
                created.push_back(sb);
                if (D > 0) {
                    cudaError_t e = cudaMemcpyAsync(buf_get(ctx, sb)->ptr(), buf_get(ctx, h.s)->ptr(), (size_t)D * sizeof(cplx), cudaMemcpyDeviceToDevice, ctx->stream);
                    if (e != cudaSuccess) rc = tne_fail(ctx, TNE_ERR_CUDA, "copy of the singular values failed: %s", cudaGetErrorString(e));
                }
                cur_s[g.site_i] = ti; cur_s[g.site_j] = tj;
                if (S.vidal) cur_b[g.bond] = sb;
            }
            for (int side = 0; side < 2; ++side) { tne_free(ctx, h.q[side]); tne_free(ctx, h.uv[side]); }
            tne_free(ctx, h.s);
        }
        for (size_t q = 0; q < err.size() && rc == TNE_OK; ++q)
            if (std::isnan(err[q])) rc = tne_fail(ctx, TNE_ERR_INVALID, "tebd sweep gate %d: non-finite values in the site tensors", ids[q]);
    }
    out_s = cur_s; out_b = cur_b;
    return rc;
}

}  // namespace

extern "C" int tne_tebd_sweeps(tne_ctx* ctx, int n_sites, tne_buf* sites, int n_bonds, tne_buf* bonds, int n_gates,
                               tne_sweep_gate* gates, int Dmax, double eps, int vidal, int n_sweeps, int32_t* kept_all, double* err_all) {
    TNE_ENTER(ctx);
    if (n_gates <= 0 || n_sweeps <= 0) return TNE_OK;
    if (!sites || !gates || (vidal && !bonds)) return tne_fail(ctx, TNE_ERR_INVALID, "null argument");
    SweepCtx S;
    S.ctx = ctx; S.n_sites = n_sites; S.n_bonds = n_bonds; S.n_gates = n_gates; S.Dmax = Dmax; S.vidal = vidal; S.eps = eps; S.gates = gates;
    // wavefront levels, order-preserving
    std::vector<int> site_level(n_sites, -1), level(n_gates, 0);
    S.n_levels = 0;
    for (int gi = 0; gi < n_gates; ++gi) {
        const tne_sweep_gate& g = gates[gi];
        if (g.site_i < 0 || g.site_i >= n_sites || g.site_j < 0 || g.site_j >= n_sites || g.site_i == g.site_j)
            return tne_fail(ctx, TNE_ERR_INVALID, "sweep gate %d: bad site index", gi);
        if (vidal && (g.bond < 0 || g.bond >= n_bonds)) return tne_fail(ctx, TNE_ERR_INVALID, "sweep gate %d: bad bond index", gi);
        level[gi] = 1 + std::max(site_level[g.site_i], site_level[g.site_j]);
        site_level[g.site_i] = site_level[g.site_j] = level[gi];
        S.n_levels = std::max(S.n_levels, level[gi] + 1);
    }
    S.by_level.assign(S.n_levels, std::vector<int>());
    for (int gi = 0; gi < n_gates; ++gi) S.by_level[level[gi]].push_back(gi);
    ensure_smem_attr(ctx, bond_kernel, (int)bond_smem_bytes());

    std::vector<int32_t> kept_loc((size_t)n_sweeps * n_gates, 0);
    std::vector<double> err_loc((size_t)n_sweeps * n_gates, 0.0);
    std::vector<tne_buf> cur_s(sites, sites + n_sites), cur_b(bonds, bonds + (vidal ? n_bonds : 0));
    std::unordered_set<tne_buf> mine;          // every live tensor this call allocated; whatever is not handed out is freed at the end
    auto drop = [&](const std::vector<tne_buf>& hs) { for (auto h : hs) if (mine.erase(h)) tne_free(ctx, h); };
    int rc = TNE_OK;
    int sw = 0;
    while (sw < n_sweeps && rc == TNE_OK) {
        // enqueue all remaining sweeps speculatively, back to back
        std::vector<SweepRun> runs((size_t)(n_sweeps - sw));
        std::vector<tne_buf> st_s = cur_s, st_b = cur_b;
        size_t enq = 0;
        for (; enq < runs.size() && rc == TNE_OK; ++enq) {
            rc = sweep_enqueue(S, st_s, st_b, runs[enq]);
            for (auto h : runs[enq].created) mine.insert(h);
            st_s = runs[enq].out_s; st_b = runs[enq].out_b;
        }
        if (rc != TNE_OK) { for (auto& R : runs) free_batches(ctx, R); break; }
        // verify in order (the first fetch synchronises the stream)
        size_t good = 0;
        for (; good < runs.size(); ++good) {
            bool ok = false;
            rc = sweep_verify(S, runs[good], kept_loc.data() + (size_t)(sw + good) * n_gates, err_loc.data() + (size_t)(sw + good) * n_gates, &ok);
            if (rc != TNE_OK || !ok) break;
        }
        if (rc != TNE_OK) { for (auto& R : runs) free_batches(ctx, R); break; }
        if (good > 0) { cur_s = runs[good - 1].out_s; cur_b = runs[good - 1].out_b; }
        if (good == runs.size()) { sw = n_sweeps; break; }
        // sweep sw + good kept fewer singular values than its cap somewhere: its results and everything after it are void
        for (size_t k = good; k < runs.size(); ++k) { free_batches(ctx, runs[k]); drop(runs[k].created); }
        std::vector<tne_buf> out_s, out_b, created;
        rc = sweep_exact(S, cur_s, cur_b, out_s, out_b, created, kept_loc.data() + (size_t)(sw + good) * n_gates,
                         err_loc.data() + (size_t)(sw + good) * n_gates);
        for (auto h : created) mine.insert(h);
        if (rc != TNE_OK) break;
        cur_s = out_s; cur_b = out_b;
        sw += (int)good + 1;
    }
    if (rc != TNE_OK) {
        cudaStreamSynchronize(ctx->stream);
        for (auto h : std::vector<tne_buf>(mine.begin(), mine.end())) tne_free(ctx, h);
        return rc;
    }
    // hand the final tensors out; every intermediate generation is released
    std::unordered_set<tne_buf> live(cur_s.begin(), cur_s.end());
    live.insert(cur_b.begin(), cur_b.end());
    for (auto h : mine)
        if (!live.count(h)) tne_free(ctx, h);
    for (int i = 0; i < n_sites; ++i) sites[i] = cur_s[i];
    for (size_t i = 0; i < cur_b.size(); ++i) bonds[i] = cur_b[i];
    for (int gi = 0; gi < n_gates; ++gi) {
        gates[gi].kept = kept_loc[(size_t)(n_sweeps - 1) * n_gates + gi];
        gates[gi].trunc_err = err_loc[(size_t)(n_sweeps - 1) * n_gates + gi];
    }
    if (kept_all) memcpy(kept_all, kept_loc.data(), kept_loc.size() * sizeof(int32_t));
    if (err_all) memcpy(err_all, err_loc.data(), err_loc.size() * sizeof(double));
    return TNE_OK;
}

extern "C" int tne_tebd_sweep(tne_ctx* ctx, int n_sites, tne_buf* sites, int n_bonds, tne_buf* bonds, int n_gates,
                              tne_sweep_gate* gates, int Dmax, double eps, int vidal) {
    return tne_tebd_sweeps(ctx, n_sites, sites, n_bonds, bonds, n_gates, gates, Dmax, eps, vidal, 1, nullptr, nullptr);
}

// LinearAlgebra.svd(A) + truncation for a matrix: the same kernel on the trivial "bond" whose sites are
// t_i[p, r, s] = A[(p, r), s] restricted ... (general matrices go through the two-site form with an identity gate)
extern "C" int tne_svd_trunc(tne_ctx* ctx, tne_buf hA, int Dmax, double eps, tne_buf* U, tne_buf* S, tne_buf* V, int32_t* kept,
                             double* trunc_err) {
    TNE_ENTER(ctx);
    Buf* a = buf_get(ctx, hA);
    if (!a) return TNE_ERR_INVALID;
    if (a->dims.size() != 2) return tne_fail(ctx, TNE_ERR_INVALID, "svd: a matrix is required");
    const int64_t m = a->dims[0], n = a->dims[1];
    if (m % 2 || n % 2 || n > MAXI / 2)
        return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "svd: supported for even m and even n <= %d (bond-update sizes)", MAXI / 2);
    // A = T_i * T_j^T with T_i = A (rows (p, o), shared s = column) and T_j = identity on (q, s):
    // site i: dims (2, m/2, n) with shared axis 2;  site j: dims (2, n/2, n) with T_j[(q, o'), s] = delta
    std::vector<cplx> eye((size_t)n * n, make_double2(0.0, 0.0));
    for (int64_t i = 0; i < n; ++i) eye[(size_t)(i * n + i)] = make_double2(1.0, 0.0);
    tne_buf hI, hAi;
    TNE_TRY(buf_new(ctx, {2, n / 2, n}, &hI));
    TNE_CUDA(ctx, cudaMemcpyAsync(buf_get(ctx, hI)->ptr(), eye.data(), eye.size() * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
    TNE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    Buf src = *a;
    TNE_TRY(buf_view(ctx, src, 0, {2, m / 2, n}, &hAi));
    tne_bond_job jb;
    memset(&jb, 0, sizeof(jb));
    jb.ti = hAi; jb.tj = hI; jb.axis_i = 2; jb.axis_j = 2;
    for (int g = 0; g < 4; ++g) jb.gate[2 * (g + 4 * g)] = 1.0;   // identity gate
    jb.Dmax = Dmax; jb.eps = eps; jb.vidal = 1;                    // "vidal" form returns bare U and conj(V)
    // unit bond vectors so that the Vidal scalings are no-ops
    std::vector<cplx> ones((size_t)std::max(m, n), make_double2(1.0, 0.0));
    tne_buf h1;
    TNE_TRY(buf_new(ctx, {(int64_t)ones.size()}, &h1));
    TNE_CUDA(ctx, cudaMemcpyAsync(buf_get(ctx, h1)->ptr(), ones.data(), ones.size() * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
    TNE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    tne_buf l_mi, l_ni, l_n;
    Buf one_src = *buf_get(ctx, h1);
    TNE_TRY(buf_view(ctx, one_src, 0, {m / 2}, &l_mi));
    TNE_TRY(buf_view(ctx, one_src, 0, {n / 2}, &l_ni));
    TNE_TRY(buf_view(ctx, one_src, 0, {n}, &l_n));
    jb.lambda_i[1] = l_mi; jb.lambda_i[2] = l_n;
    jb.lambda_j[1] = l_ni; jb.lambda_j[2] = l_n;
    int rc = tne_apply_onbond_batched(ctx, 1, &jb);
    tne_free(ctx, hI); tne_free(ctx, hAi); tne_free(ctx, h1); tne_free(ctx, l_mi); tne_free(ctx, l_ni); tne_free(ctx, l_n);
    if (rc) return rc;
    // ti_out: (2, m/2, D) = U as (m x D); tj_out: (2, n/2, D) = conj(V); return V itself
    const int D = jb.kept;
    Buf ub = *buf_get(ctx, jb.ti_out);
    TNE_TRY(buf_view(ctx, ub, 0, {m, (int64_t)D}, U));
    tne_free(ctx, jb.ti_out);
    tne_buf vc;
    Buf vb = *buf_get(ctx, jb.tj_out);
    TNE_TRY(buf_view(ctx, vb, 0, {n, (int64_t)D}, &vc));
    tne_free(ctx, jb.tj_out);
    rc = tne_map(ctx, TNE_CONJ, vc, 0.0, 0.0, V);
    tne_free(ctx, vc);
    if (rc) return rc;
    *S = jb.s_out;
    if (kept) *kept = D;
    if (trunc_err) *trunc_err = jb.trunc_err;
    return TNE_OK;
}
