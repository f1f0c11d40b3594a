// Internal declarations shared by the translation units of libtne_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <string>
#include <set>
#include <unordered_map>
#include <vector>

#include "tne.h"

typedef double2 cplx;  // interleaved (re, im), identical to Julia's ComplexF64

struct Block {          // one device allocation; views (reshape / slice) share it
    void* ptr = nullptr;
    int64_t bytes = 0;
    struct tne_ctx* ctx = nullptr;
    ~Block();
};

struct Buf {
    std::shared_ptr<Block> block;
    int64_t offset = 0;                 // in elements
    std::vector<int64_t> dims;
    int64_t nelem = 1;
    cplx* ptr() const { return reinterpret_cast<cplx*>(block->ptr) + offset; }
};

struct tne_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    std::unordered_map<int64_t, Buf> bufs;
    int64_t next_handle = 1;
    // caching allocator for user tensors
    std::multimap<int64_t, void*> free_blocks;
    int64_t in_use = 0, peak = 0, reserved = 0;
    // workspace arena for tree programs
    void* ws = nullptr;
    int64_t ws_bytes = 0;
    int64_t mem_limit = 0;
    int64_t total_mem = 0;
    int sm_count = 148;
    int64_t launches = 0;
    double stats[6] = {0, 0, 0, 0, 0, 0};
    // options
    int64_t opt_force_generic = 0;
    int64_t opt_debug = 0;
    int64_t opt_profile = 0;
    int64_t opt_pad = 0;                   // stride padding (elements) of workspace tensors, see tne_tree.cu
    int64_t opt_no_rows = 1;               // 0: row-contiguous staged loads in the small-K absorb kernel (measured slower: off)
    int64_t opt_no_pdl = 1;                // 1: no programmatic dependent launch of the absorb kernels
    int64_t opt_no_site2 = 0;              // 1: never fuse the bra and ket step of a site into one kernel
    int64_t opt_no_fuse = 0;               // 1: one launch per contribution (no gett_absorb_multi)
    int64_t opt_no_permute = 0;            // 1: unary permutations run as contractions with the constant 1 (gett_generic)
    int64_t opt_no_gemm = 0;               // 1: never use the tiled DMMA GEMM for general dense nodes (tne_gemm.cu)
    int64_t opt_no_site2b = 0;             // 1: never fuse the three reverse ops of a site into one kernel (tne_site2b.cu)
    int64_t opt_no_plan_cache = 0;         // 1: tne_expect_terms re-plans every call (default: compiled programs are cached, tne_tree.cu)
    std::list<std::pair<std::string, std::shared_ptr<void>>> plan_cache;   // key bytes -> compiled program, most recently used first
    int64_t plan_cache_hits = 0;
    cplx* s2b_part = nullptr;              // CTA partials of the fused reverse kernel's leaf gradient
    cplx* red_part = nullptr;              // CTA partials of the reduce kernels (deterministic two-pass reduction)
    int64_t red_part_elems = 0;
    int64_t opt_absorb_tma = 0;            // 1: TMA-staged absorb kernel where applicable (default: register-fed)
    int64_t opt_no_reduce_tma = 0;         // 1: register-fed reduce kernel instead of the TMA-staged one
    int64_t opt_rs_comm = 0;               // 1: reduce-scatter / all-gather of column checkpoints instead of all-reduce (tne_tree.cu)
    int64_t opt_tape_gib = -1;             // forward tapes kept for the reverse pass: -1 whatever memory is left, 0 none, n > 0: n GiB
    int64_t opt_comm_overlap = 1;          // 1: the all-gathers of the reverse pass run on a second stream behind the next column's recompute
    cudaStream_t comm_stream = nullptr;    // created on first use
    cudaStream_t comm_active = nullptr;    // stream the collectives of tne_comm.cu are issued on right now (nullptr: the compute stream)
    cudaEvent_t ev_comm_ready = nullptr, ev_comm_done = nullptr;
    struct ProfRec { cudaEvent_t e0, e1; int cls; double flops, bytes; };
    std::vector<ProfRec> prof;              // pending event pairs
    std::vector<cudaEvent_t> event_pool;
    cplx* one_dev = nullptr;               // device constant 1+0i
    void* nccl = nullptr;                  // ncclComm_t
    void* nccl_lib = nullptr;
    int world = 1, rank = 0;
    // planner context (TNE_PLAN_ONLY=1 and no device): programs are built, costed and dumped, nothing is ever
    // computed -- tne_expect_terms / tne_contract_tree return TNE_ERR_UNSUPPORTED after planning (tools/plan_dump.py)
    bool plan_only = false;
    // kernels whose dynamic shared-memory opt-in has been set on THIS context's device (the attribute is per device)
    std::set<const void*> smem_attr_done;
};

// every entry point binds the calling thread to the context's device first (a host program may switch devices or call
// from another thread between two calls)
#define TNE_ENTER(ctx)                                                      \
    do {                                                                    \
        if ((ctx) && !(ctx)->plan_only) cudaSetDevice((ctx)->device);       \
    } while (0)

template <typename K>
inline void ensure_smem_attr(tne_ctx* ctx, K kernel, int bytes) {
    const void* key = reinterpret_cast<const void*>(kernel);
    if (ctx->smem_attr_done.count(key)) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    ctx->smem_attr_done.insert(key);
}

// ---- error helpers ---------------------------------------------------------------------
int tne_fail(tne_ctx* ctx, int code, const char* fmt, ...);
#define TNE_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess && (ctx) && (ctx)->plan_only) cudaGetLastError();              \
        else if (e__ != cudaSuccess)                                                          \
            return tne_fail(ctx, e__ == cudaErrorMemoryAllocation ? TNE_ERR_OOM : TNE_ERR_CUDA, \
                            "CUDA error '%s' at %s:%d", cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define TNE_TRY(call)             \
    do {                          \
        int rc__ = (call);        \
        if (rc__ != TNE_OK) return rc__; \
    } while (0)
#define TNE_LAUNCH_CHECK(ctx)                 \
    do {                                      \
        (ctx)->launches++;                    \
        TNE_CUDA(ctx, cudaGetLastError());    \
    } while (0)

// ---- buffers ---------------------------------------------------------------------------
int buf_new(tne_ctx* ctx, const std::vector<int64_t>& dims, tne_buf* out);
Buf* buf_get(tne_ctx* ctx, tne_buf h);
int buf_view(tne_ctx* ctx, const Buf& src, int64_t offset, const std::vector<int64_t>& dims, tne_buf* out);
int ws_reserve(tne_ctx* ctx, int64_t bytes);

// ---- contraction (tne_contract.cu) ------------------------------------------------------
struct TensorRef {                  // a tensor operand: labels in memory order (fastest first)
    cplx* ptr = nullptr;
    std::vector<int32_t> labels;
    std::vector<int64_t> dims;
    std::vector<int64_t> strides;   // in elements; empty = column-major contiguous
    bool conj = false;
    int64_t nelem() const { int64_t n = 1; for (auto d : dims) n *= d; return n; }
};
#define MAXLEG 16
struct LegSet {
    int n;
    int size[MAXLEG];
    int shift[MAXLEG];   // log2(size) for power-of-two legs (index arithmetic by shift/mask), else -1
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
// A classified contraction, ready to launch any number of times with different base pointers
// (the slice loops of tne_tree.cu re-launch the same plan with shifted operands).
struct ContractPlan {
    GettDesc d;
    int kernel = 0;            // 0: generic FMA; >0: DMMA fast paths (tne_contract_dmma.cu)
    unsigned gx = 1, gy = 1, gz = 1;
    int threads = 256;
    size_t smem = 0;
    bool swapped = false;      // operands A and B exchanged (the larger free side sits on M)
    bool zero_first = false;   // split-K with atomics into a non-accumulating output
    int64_t c_nelem = 0;
    bool empty = false;
    int aux[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned char blob[64] = {0};   // kernel-specific geometry (tne_contract_dmma.cu)
    double flops = 0, bytes = 0;
};
// fused absorption of one PEPS site, C = (A . B1) . B2 in one kernel (tne_site2.cu)
struct Legs4 { int n; int size[4]; long long s0[4], s1[4]; };
struct Site2Desc {
    GettDesc g;                    // g.M (s0: A stride, s1: C stride), g.Msz: the boundary rows
    Legs4 K1, X, P, N1, N2;        // (A, B1), (A, B2), (B1, B2), (B1, C), (B2, C) strides
    int conjA, conjB1, conjB2, accumulate;
    double a1re, a1im, a2re, a2im;
};
bool site2_plan(tne_ctx* ctx, const TensorRef& A, const TensorRef& B1, const TensorRef& B2, const TensorRef& C, bool accumulate,
                double a1re, double a1im, double a2re, double a2im, Site2Desc* out);
int site2_run(tne_ctx* ctx, const Site2Desc& d, const cplx* A, const cplx* B1, const cplx* B2, cplx* C);

// fused reverse pass of one PEPS site: dT = dC . B2^H (kept on the SM), dA = dT . B1^H, dB1 += A^H . dT (tne_site2b.cu)
struct Site2BwdDesc {
    LegSet M;                      // boundary rows: s0 A stride, s1 dC stride, s2 dA stride
    long long Msz;
    Legs4 K1, X, P, N1, N2;        // (A, B1), (A, B2), (B1, B2), (B1, dC), (B2, dC) strides
    Legs4 K1d, Xd;                 // s0: dA strides
    Legs4 K1g, Pg, N1g;            // s0: strides of the leaf gradient dB1
    int conjC, conjB2;             // dT  = a1 f(dC) f(B2)
    int conjT2, conjB1;            // dA  = a2 f(dT) f(B1)
    int conjA, conjT3;             // dB1 += a3 f(A) f(dT)
    int accumulate;                // dA += ...
    int part_add;                  // add the CTA partials of the leaf gradient to the workspace instead of overwriting it
    double a1re, a1im, a2re, a2im, a3re, a3im;
};
bool site2b_plan(tne_ctx* ctx, const TensorRef& dC, const TensorRef& B2, const TensorRef& B1, const TensorRef& A, const TensorRef& dA,
                 const TensorRef& dB1, Site2BwdDesc* out);
int64_t site2b_part_elems(tne_ctx* ctx);
int site2b_run(tne_ctx* ctx, const Site2BwdDesc& d, const cplx* dC, const cplx* B2, const cplx* B1, const cplx* A, cplx* dA, cplx* dB1,
               cplx* part, bool part_add, bool finalize);

// several contributions into one output through the DMMA absorb kernel (see gett_absorb_multi)
#define TNE_MULTI_MAX 3
struct AbsorbMulti {
    int T;
    const cplx* A[TNE_MULTI_MAX];
    const cplx* B[TNE_MULTI_MAX];
    int conjA[TNE_MULTI_MAX], conjB[TNE_MULTI_MAX];
    double are[TNE_MULTI_MAX], aim[TNE_MULTI_MAX];
};
bool contract_multi_compatible(const ContractPlan& x, const ContractPlan& y);
// C (+)= sum_t: plans[t] are compatible plans; As/Bs are the operands in the ORIGINAL (a, b) order of each op
int contract_run_multi(tne_ctx* ctx, int T, const ContractPlan* const* plans, const cplx* const* As, const cplx* const* Bs, cplx* C);
int contract_plan(tne_ctx* ctx, const TensorRef& A, const TensorRef& B, const TensorRef& C, bool accumulate,
                  double alpha_re, double alpha_im, ContractPlan* plan);
int contract_run(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C);
// C (+)= A * B over the labels; C's layout is given by its labels/strides.
int contract_launch(tne_ctx* ctx, const TensorRef& A, const TensorRef& B, const TensorRef& C, bool accumulate,
                    double alpha_re = 1.0, double alpha_im = 0.0);
// tiled complex128 DMMA GEMM for general dense nodes (tne_gemm.cu); plan->kernel = 7
bool gemm_plan(tne_ctx* ctx, ContractPlan* plan);
int gemm_run(tne_ctx* ctx, const ContractPlan& plan, const cplx* A, const cplx* B, cplx* C);
// shared-memory tiled tensor permutation (tne_permute.cu)
bool permute_applicable(const TensorRef& A, const TensorRef& C);
int permute_launch(tne_ctx* ctx, const TensorRef& A, const TensorRef& C);
// workspace of CTA / split partials of the two-pass (deterministic) reductions; grows on demand (tne_contract_dmma.cu)
int reduce_part_reserve(tne_ctx* ctx, int64_t elems);
// algorithmic cost of one contraction (SURVEY.md 8(d) accounting)
void contract_cost(const TensorRef& A, const TensorRef& B, const TensorRef& C, double* flops, double* bytes);

// ---- multi-GPU (tne_comm.cu): in-place sum over the ranks of `n` complex elements at a device pointer
int comm_allreduce_ptr(tne_ctx* ctx, cplx* p, int64_t n);
int comm_allreduce_many(tne_ctx* ctx, int n, cplx* const* ptrs, const int64_t* counts);   // one grouped NCCL call
int comm_group_begin(tne_ctx* ctx);   // collectives until comm_group_end form one NCCL group
int comm_group_end(tne_ctx* ctx);
// `nblocks` contiguous blocks of `block` complex elements at p, block b owned by rank b % world:
// reduce-scatter: on return the blocks a rank owns hold the sum over ranks (the others are unspecified);
// all-gather: every rank contributes the blocks it owns, on return all blocks are complete everywhere.
int comm_reduce_scatter_blocks(tne_ctx* ctx, cplx* p, int64_t block, int64_t nblocks);
int comm_all_gather_blocks(tne_ctx* ctx, cplx* p, int64_t block, int64_t nblocks);
void comm_destroy(tne_ctx* ctx);   // ncclCommDestroy of the context's communicator, if any

// ---- elementwise helpers used across units ---------------------------------------------
int launch_fill_zero(tne_ctx* ctx, cplx* p, int64_t n);
int launch_fill_value(tne_ctx* ctx, cplx* p, int64_t n, double re, double im);
int launch_scale_inplace(tne_ctx* ctx, cplx* p, int64_t n, double re, double im);
int launch_axpy_dev(tne_ctx* ctx, cplx* y, const cplx* x, int64_t n, const cplx* alpha_dev);
