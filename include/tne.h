/*
 * tne.h -- C ABI of the B200-native engine for TensorNetworkEvolve.jl's hot path.
 *
 * The reference has no FFI; its plug-in point is Julia multiple dispatch on the tensor
 * array type (both in-tree examples hook `OMEinsum.einsum`: src/TensorAD/rules.jl:1-8,
 * src/cracker.jl:4-8).  Each entry point below names the reference call it stands behind;
 * the Julia-side `ccall` stubs are in INTEGRATION.md / julia/TensorNetworkEvolveB200.jl.
 *
 * Conventions
 *  - every tensor is ComplexF64 (interleaved re,im; 16 bytes), column-major, owned by the
 *    library and named by an opaque 64-bit handle (`tne_buf`, 0 = null).  Real-valued
 *    Julia arrays (norms, singular values) are carried with zero imaginary part.
 *  - host pointers are borrowed for the duration of the call only.
 *  - every function returns 0 on success and a non-zero `tne_status` otherwise; it never
 *    throws or aborts.  `tne_last_error` gives the message (the Julia shim raises
 *    ErrorException with it, mirroring `error(...)`/`@assert` at src/peps.jl:74,259,330,356).
 *  - one context per GPU / per process, driven from one host thread; all work is enqueued
 *    on the context's stream and is asynchronous until tne_download / tne_sync.
 *  - there is no CPU fallback: without a CUDA device tne_ctx_create fails.
 */
#ifndef TNE_H
#define TNE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct tne_ctx tne_ctx;
typedef int64_t tne_buf;

enum tne_status {
    TNE_OK = 0,
    TNE_ERR_INVALID = 1,
    TNE_ERR_CUDA = 2,
    TNE_ERR_OOM = 3,
    TNE_ERR_UNSUPPORTED = 4,
    TNE_ERR_NCCL = 5
};

#define TNE_MAX_RANK 32

/* ---- context (new; the reference is single-process CPU code) ------------------------- */
int tne_version(void);
int tne_ctx_create(int device, tne_ctx** out);
int tne_ctx_destroy(tne_ctx* ctx);
const char* tne_last_error(tne_ctx* ctx);
int tne_sync(tne_ctx* ctx);
/* workspace budget for tree programs (bytes); default = 85% of the device memory */
int tne_set_mem_limit(tne_ctx* ctx, int64_t bytes);
int tne_mem_info(tne_ctx* ctx, int64_t* in_use, int64_t* peak, int64_t* workspace);
/* cudaStream_t of the context (for CUDA-event timing by the caller) */
void* tne_stream(tne_ctx* ctx);
/* number of kernels this library has launched on the context so far */
int64_t tne_launch_count(tne_ctx* ctx);
/* statistics of the most recent tree program: [0] algorithmic real flops, [1] algorithmic bytes,
 * [2] contraction kernels launched, [3] workspace bytes, [4] flops of recomputed forward values (executed on
 * top of [0]), [5] number of values */
int tne_last_program_stats(tne_ctx* ctx, double* stats6);
/* tne_expect_terms / tne_contract_tree keep the programs they compile (op lists, classified launch plans, memory plan), keyed
 * by (tree, leaf shapes, term pattern and coefficients, gradient request, options): a TDVP driver that evaluates the same
 * <H> + gradient every step plans once.  entries: programs held now; hits: calls served from the cache so far.
 * Option "no_plan_cache" = 1 disables and clears it. */
int tne_plan_cache_info(tne_ctx* ctx, int64_t* entries, int64_t* hits);
/* option knobs: "force_generic_kernel", "debug", "profile" (1: bracket every contraction launch with a
 * CUDA-event pair on the context's stream; read the totals with tne_profile_collect), "absorb_tma" (1: use the
 * TMA-staged absorb kernel where it applies), "no_reduce_tma" (1: register-fed reduce kernel), "pad",
 * "no_permute" (1: unary permutations run on the FMA contraction kernel instead of the tiled transpose), "no_gemm" (1: general dense nodes run on the FMA kernel instead of the tiled DMMA GEMM), "no_plan_cache" (1: re-plan every tree call), "comm_overlap" (0: the all-gathers of the multi-GPU reverse pass stay on the compute stream), "no_site2" / "no_site2b" / "no_fuse" (1: never fuse the two steps of a site / the contributions into one output: the step-by-step
 * launch plan), "rs_comm" (1: multi-GPU checkpoints are combined by reduce-scatter / all-gather instead of all-reduce).
 * A context created with TNE_PLAN_ONLY=1 in the environment on a machine WITHOUT a device is a planner: tree programs are
 * built, costed and dumped to stderr, no kernel ever runs, tne_download and the tree entry points fail. */
int tne_set_option(tne_ctx* ctx, const char* key, int64_t value);
/* Per kernel class c = 8 * kernel + bucket (kernel: 0 generic FMA, 1 DMMA absorb, 2 DMMA reduce, 3 DMMA absorb with
 * TMA-staged operand, 4 DMMA reduce with TMA-staged operands, 5 fused two-stage site absorption, 6 fused reverse pass of a site, 7 tiled DMMA GEMM of a general dense node; bucket: log2 of the padded K / 2 of absorb kernels): out[4c .. 4c+3] = launches, summed
 * device milliseconds, algorithmic real flops, algorithmic bytes since the last collect.  Synchronises. */
#define TNE_PROFILE_CLASSES 64   /* class = 8 * kernel + log2(k-steps) bucket (0 for kernels without one) */
int tne_profile_collect(tne_ctx* ctx, double* out);

/* roofline denominators of the context's device, measured now (a few ms): issue peaks of FP64 DMMA (mma.sync.m8n8k4, the
 * tensor-pipe primitive of every dense contraction step here) and of DFMA, in TFLOP/s */
int tne_measure_fp64_peak(tne_ctx* ctx, double* dmma_tflops, double* dfma_tflops);

/* ---- device tensors: Array{ComplexF64,N} as used at src/peps.jl:41,132 ---------------- */
int tne_alloc(tne_ctx* ctx, int rank, const int64_t* dims, tne_buf* out);
int tne_free(tne_ctx* ctx, tne_buf buf);
int tne_shape(tne_ctx* ctx, tne_buf buf, int* rank, int64_t* dims);
int tne_upload(tne_ctx* ctx, tne_buf buf, const void* host_c128);      /* Array -> device */
int tne_download(tne_ctx* ctx, tne_buf buf, void* host_c128);          /* Array(x); synchronises */
void* tne_device_ptr(tne_ctx* ctx, tne_buf buf);

/* ---- OMEinsum.einsum(code, xs, size_dict), unary and binary ----------------------------
 * Call sites: src/peps.jl:334,344,373,376,392-393,429; src/TensorAD/rules.jl:2-5 (forward and
 * einsum_grad); src/timeevolve.jl:38.  `labels_flat` concatenates the inputs' labels,
 * `conj_flags[i] != 0` conjugates input i while it is loaded (fuses the conj copies of
 * src/peps.jl:107 and of einsum_grad). */
int tne_einsum(tne_ctx* ctx, int n_in, const int32_t* labels_flat, const int32_t* ranks,
               const tne_buf* in, const int32_t* conj_flags,
               const int32_t* out_labels, int out_rank, tne_buf* out);
/* The same with OMEinsum's `size_dict` argument: sizes of labels that only the output carries (the "repeat" rule, e.g.
 * the pullback of ein"ii->" is ein"->ii").  A label repeated inside one operand reads its diagonal (trace / diag
 * rules); repeated inside the output, only the diagonal is written and the rest is zero. */
int tne_einsum_sized(tne_ctx* ctx, int n_in, const int32_t* labels_flat, const int32_t* ranks,
                     const tne_buf* in, const int32_t* conj_flags,
                     const int32_t* out_labels, int out_rank,
                     int n_sizes, const int32_t* size_labels, const int64_t* size_values, tne_buf* out);

/* ---- Base / broadcast methods used by src/TensorAD/rules.jl and src/peps.jl:283-295 ----- */
enum tne_map_op {
    TNE_COPY = 0, TNE_CONJ = 1, TNE_REAL = 2, TNE_IMAG = 3, TNE_ABS = 4, TNE_SQRT = 5, TNE_INV = 6,
    TNE_NEG = 7, TNE_SCALE = 8,   /* x * s */
    TNE_POW = 9,                  /* x .^ s (s real), principal branch */
    TNE_SIN = 10, TNE_COS = 11, TNE_SAFE_INV = 12 /* src/peps.jl:419-426 */
};
enum tne_map2_op { TNE_ADD = 0, TNE_SUB = 1, TNE_MUL = 2, TNE_DIV = 3 };
int tne_map(tne_ctx* ctx, int op, tne_buf in, double s_re, double s_im, tne_buf* out);
/* elementwise binary; an operand with exactly one element broadcasts (0-dim * tensor, rules.jl:79-88) */
int tne_map2(tne_ctx* ctx, int op, tne_buf a, tne_buf b, tne_buf* out);
int tne_reduce_sum(tne_ctx* ctx, tne_buf in, tne_buf* out);                     /* sum(x) -> 0-dim */
int tne_fill(tne_ctx* ctx, tne_buf scalar0, int rank, const int64_t* dims, tne_buf* out); /* fill(x[], dims) */
/* x[lo+1 : lo+count] of the flattened tensor as a new rank-1 tensor (variables[a:b], src/peps.jl:217) */
int tne_slice(tne_ctx* ctx, tne_buf in, int64_t lo, int64_t count, tne_buf* out);
/* accum(zero(x), lo+1:lo+count, y): a zero vector of `total` with y added at the range (rules.jl:46-55) */
int tne_embed(tne_ctx* ctx, tne_buf y, int64_t lo, int64_t total, tne_buf* out);
int tne_reshape(tne_ctx* ctx, tne_buf in, int rank, const int64_t* dims, tne_buf* out);    /* shares storage */
int tne_concat(tne_ctx* ctx, int n, const tne_buf* in, tne_buf* out);            /* vcat(vec.(xs)...) (src/peps.jl:202) */

/* ---- (::NestedEinsum)(xs...) / (::SlicedEinsum)(xs...) fast path: src/peps.jl:109,321 ---- */
typedef struct {
    int32_t n_leaves;
    int32_t n_nodes;                 /* binary nodes in post-order (children before parents) */
    const int32_t* leaf_rank;        /* [n_leaves] */
    const int32_t* leaf_labels;      /* concatenated leaf labels */
    const int32_t* node_children;    /* [2*n_nodes]; id < n_leaves is a leaf, otherwise n_leaves + node index */
    const int32_t* node_out_rank;    /* [n_nodes] */
    const int32_t* node_out_labels;  /* concatenated; the order is honoured for the root only */
    int32_t n_sliced;                /* SlicedEinsum: labels looped over and summed */
    const int32_t* sliced_labels;
    int32_t shard_rank, shard_world; /* this process takes work items with index % world == rank */
    /* Optional checkpointed segments (n_segments <= 1: the whole tree is one segment).  Segment s covers
     * the post-order nodes [seg_first_node[s], seg_first_node[s+1]); seg_first_node[0] must be 0.  Values
     * crossing a segment boundary are kept as checkpoints; the reverse pass walks the segments backwards
     * and recomputes each segment's forward values from its checkpoints, so only one segment's tape is
     * alive at a time.  Inside segment s the labels seg_sliced_labels[seg_sliced_off[s] .. seg_sliced_off[s+1])
     * are looped over: every value of the segment that carries such a label is handled one index at a time
     * (1/size of the memory); checkpoints keep their full shape and are read / written through strided
     * views or accumulated.  A label that is open for the whole segment costs no extra flops this way. */
    /* seg_shard != 0: the segment-local slices are dealt round-robin to the ranks of the context's communicator
     * (tne_comm_init); after every sliced segment the checkpoints it produced are summed over the ranks with an
     * NCCL all-reduce on the context's stream, and so are the outputs, which are complete on every rank on
     * return.  Unsliced segments are replicated. */
    int32_t seg_shard;
    int32_t n_segments;
    const int32_t* seg_first_node;   /* [n_segments] */
    const int32_t* seg_sliced_off;   /* [n_segments + 1] */
    const int32_t* seg_sliced_labels;
} tne_tree;

int tne_contract_tree(tne_ctx* ctx, const tne_tree* tree, const tne_buf* leaves,
                      const int32_t* leaf_conj, tne_buf* out);

/* ---- Yao.expect(op::Add, pa, pb) + its TensorAD pullback: src/peps.jl:469-477,
 * src/timeevolve.jl:55-64.  A term replaces up to 4 leaves of the network (h_k applied to the ket
 * tensors) and carries a scalar coefficient; value = sum_k coef_k * tree(leaves with term k's
 * replacements).  Terms share every sub-contraction that does not contain a replaced leaf, and
 * completed terms are summed into one running partial result.  With n_terms == 0 the plain tree is
 * evaluated.  If n_grad > 0 the pullback for cotangent `seed` (convention g = dL/dre + i dL/dim)
 * w.r.t. the listed leaves is returned in new buffers shaped like those leaves.
 * Work is sharded over (sliced-label assignments) with index % shard_world == shard_rank; the
 * caller all-reduces `value` and the gradients (tne_allreduce_sum or torch.distributed). */
typedef struct {
    int32_t n_repl;
    int32_t leaf[4];
    tne_buf repl[4];
    double coef_re, coef_im;
} tne_term;

int tne_expect_terms(tne_ctx* ctx, const tne_tree* tree, const tne_buf* leaves, const int32_t* leaf_conj,
                     int n_terms, const tne_term* terms,
                     int n_grad, const int32_t* grad_leaves, double seed_re, double seed_im,
                     tne_buf* value_out, tne_buf* grads_out);

/* ---- LinearAlgebra.svd + truncation and the fused apply_onbond!: src/peps.jl:353-417 ------
 * A is m x n column-major.  Keeps D = min(Dmax, first index with S < eps minus one) (src/peps.jl:383-390);
 * U is m x D, S has D entries (real), V is n x D with A ~= U diag(S) V^H.  `kept` and `trunc_err`
 * (sum of the discarded singular values, the @info of src/peps.jl:386) are host outputs. */
int tne_svd_trunc(tne_ctx* ctx, tne_buf A, int Dmax, double eps,
                  tne_buf* U, tne_buf* S, tne_buf* V, int32_t* kept, double* trunc_err);

typedef struct {
    tne_buf ti, tj;            /* site tensors, physical leg first (src/peps.jl:253-254) */
    int32_t axis_i, axis_j;    /* 0-based position of the shared bond in ti / tj */
    double gate[32];           /* 4x4 complex, column-major, M[o_i + 2 o_j, i_i + 2 i_j] */
    int32_t Dmax;
    double eps;
    int32_t vidal;             /* 0: SimplePEPS (sqrt(S) into both); 1: VidalPEPS */
    tne_buf lambda_i[TNE_MAX_RANK]; /* Vidal: bond vector per axis of ti (entry for axis 0 ignored) */
    tne_buf lambda_j[TNE_MAX_RANK];
    /* outputs */
    tne_buf ti_out, tj_out, s_out;
    int32_t kept;
    double trunc_err;
} tne_bond_job;

int tne_apply_onbond_batched(tne_ctx* ctx, int n_jobs, tne_bond_job* jobs);

/* ---- one whole TEBD sweep: rydberg_tebd_sweep! (src/tebd.jl:15-23), i.e. apply_onbond! over edges(g) in order ------
 * `sites` / `bonds` are tables of handles (vertex tensors in vertex order; Vidal bond vectors in virtual_labels order);
 * gates refer to them by 0-based index and are given in the reference's sequential order.  The library levels them
 * into wavefronts of site-disjoint gates (exactly commuting), runs one launch per wavefront with NO host
 * synchronisation in between (the kept rank is assumed to reach its cap and verified once at the end; otherwise the
 * sweep is redone with exact ranks) and overwrites the tables with the handles of the new tensors.  The old handles
 * stay valid and owned by the caller.  `kept` and `trunc_err` are filled per gate (the @info of src/peps.jl:386). */
typedef struct {
    int32_t site_i, site_j;          /* 0-based indices into `sites` */
    int32_t axis_i, axis_j;          /* 0-based position of the shared bond in the two site tensors */
    int32_t bond;                    /* Vidal: index into `bonds` of the vector this gate rewrites; else ignored */
    double gate[32];                 /* 4x4 complex, column-major, M[o_i + 2 o_j, i_i + 2 i_j] */
    int32_t lambda_i[TNE_MAX_RANK];  /* Vidal: index into `bonds` per axis of site_i (-1: none; axis 0 ignored) */
    int32_t lambda_j[TNE_MAX_RANK];
    /* outputs */
    int32_t kept;
    double trunc_err;
} tne_sweep_gate;

int tne_tebd_sweep(tne_ctx* ctx, int n_sites, tne_buf* sites, int n_bonds, tne_buf* bonds, int n_gates,
                   tne_sweep_gate* gates, int Dmax, double eps, int vidal);
/* rydberg_tebd! (src/tebd.jl:25-38): `n_sweeps` identical sweeps back to back in one call.  All of them are enqueued
 * speculatively (no host synchronisation between wavefronts OR between sweeps) and verified once; a sweep whose ranks did not
 * reach their caps is redone exactly and the rest is enqueued again.  `kept_all` / `err_all` ([n_sweeps * n_gates], sweep-major;
 * may be NULL) receive the kept rank and the truncation error of every gate of every sweep; the gates' own output fields hold
 * those of the last sweep. */
int tne_tebd_sweeps(tne_ctx* ctx, int n_sites, tne_buf* sites, int n_bonds, tne_buf* bonds, int n_gates,
                    tne_sweep_gate* gates, int Dmax, double eps, int vidal, int n_sweeps, int32_t* kept_all, double* err_all);

/* ---- multi-GPU: sum of partial energies / gradients (serial `sum` at src/peps.jl:470) ------
 * comm is created from an NCCL unique id broadcast by the host program. */
int tne_nccl_unique_id(void* id128);
int tne_comm_init(tne_ctx* ctx, const void* id128, int rank, int world);
int tne_allreduce_sum(tne_ctx* ctx, int n, const tne_buf* bufs);

#ifdef __cplusplus
}
#endif
#endif /* TNE_H */
