// Multi-GPU sum of partial energies / gradients over NCCL (K6).  libnccl is loaded lazily with
// dlopen so that single-GPU use has no NCCL dependency.
#include <dlfcn.h>

#include "tne_internal.h"

namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_comm_init)(void**, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_reduce_scatter)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_group)(void);
typedef int (*fn_comm_destroy)(void*);
typedef const char* (*fn_errstr)(int);
struct NcclApi {
    void* lib = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_comm_init comm_init = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_reduce_scatter reduce_scatter = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
    fn_errstr errstr = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
};
NcclApi g_nccl;

// collectives go to the compute stream unless the tree executor has routed them to its communication stream
inline cudaStream_t comm_stream_of(tne_ctx* ctx) { return ctx->comm_active ? ctx->comm_active : ctx->stream; }

int load_nccl(tne_ctx* ctx) {
    if (g_nccl.lib) return TNE_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return tne_fail(ctx, TNE_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.comm_init = (fn_comm_init)dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.reduce_scatter = (fn_reduce_scatter)dlsym(g_nccl.lib, "ncclReduceScatter");
    g_nccl.all_gather = (fn_all_gather)dlsym(g_nccl.lib, "ncclAllGather");
    g_nccl.group_start = (fn_group)dlsym(g_nccl.lib, "ncclGroupStart");
    g_nccl.group_end = (fn_group)dlsym(g_nccl.lib, "ncclGroupEnd");
    g_nccl.errstr = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
    g_nccl.comm_destroy = (fn_comm_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
    if (!g_nccl.get_uid || !g_nccl.comm_init || !g_nccl.allreduce || !g_nccl.group_start || !g_nccl.group_end)
        return tne_fail(ctx, TNE_ERR_NCCL, "libnccl is missing required symbols");
    return TNE_OK;
}
}  // namespace

extern "C" int tne_nccl_unique_id(void* id128) {
    TNE_TRY(load_nccl(nullptr));
    int rc = g_nccl.get_uid((nccl_uid*)id128);
    return rc == 0 ? TNE_OK : tne_fail(nullptr, TNE_ERR_NCCL, "ncclGetUniqueId failed (%d)", rc);
}

extern "C" int tne_comm_init(tne_ctx* ctx, const void* id128, int rank, int world) {
    TNE_ENTER(ctx);
    TNE_TRY(load_nccl(ctx));
    TNE_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_uid uid;
    memcpy(&uid, id128, sizeof(uid));
    if (ctx->nccl) comm_destroy(ctx);   // re-initialisation replaces the communicator
    int rc = g_nccl.comm_init(&ctx->nccl, world, uid, rank);
    if (rc != 0) return tne_fail(ctx, TNE_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
    ctx->rank = rank;
    ctx->world = world;
    return TNE_OK;
}

void comm_destroy(tne_ctx* ctx) {
    if (ctx->nccl && g_nccl.comm_destroy) g_nccl.comm_destroy(ctx->nccl);
    ctx->nccl = nullptr;
    ctx->world = 1; ctx->rank = 0;
}

int comm_allreduce_ptr(tne_ctx* ctx, cplx* p, int64_t n) {
    if (ctx->world <= 1 || n <= 0) return TNE_OK;
    if (!ctx->nccl) return tne_fail(ctx, TNE_ERR_NCCL, "tne_comm_init has not been called");
    const int NCCL_DOUBLE = 8, NCCL_SUM = 0;
    int rc = g_nccl.allreduce(p, p, (size_t)(2 * n), NCCL_DOUBLE, NCCL_SUM, ctx->nccl, comm_stream_of(ctx));
    if (rc != 0) return tne_fail(ctx, TNE_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
    return TNE_OK;
}

// All collectives issued between the two calls form ONE NCCL group (nested groups are legal: the outermost end launches):
// the checkpoints of a column go out as one fused operation instead of one launch per tensor.
int comm_group_begin(tne_ctx* ctx) {
    if (ctx->world <= 1 || !ctx->nccl) return TNE_OK;
    int rc = g_nccl.group_start();
    return rc == 0 ? TNE_OK : tne_fail(ctx, TNE_ERR_NCCL, "ncclGroupStart failed: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
}
int comm_group_end(tne_ctx* ctx) {
    if (ctx->world <= 1 || !ctx->nccl) return TNE_OK;
    int rc = g_nccl.group_end();
    return rc == 0 ? TNE_OK : tne_fail(ctx, TNE_ERR_NCCL, "ncclGroupEnd failed: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
}

int comm_allreduce_many(tne_ctx* ctx, int n, cplx* const* ptrs, const int64_t* counts) {
    if (ctx->world <= 1 || n <= 0) return TNE_OK;
    if (!ctx->nccl) return tne_fail(ctx, TNE_ERR_NCCL, "tne_comm_init has not been called");
    const int NCCL_DOUBLE = 8, NCCL_SUM = 0;
    int rc = g_nccl.group_start();
    for (int i = 0; i < n && rc == 0; ++i)
        if (counts[i] > 0) rc = g_nccl.allreduce(ptrs[i], ptrs[i], (size_t)(2 * counts[i]), NCCL_DOUBLE, NCCL_SUM, ctx->nccl, comm_stream_of(ctx));
    int rc2 = g_nccl.group_end();
    if (rc != 0 || rc2 != 0) return tne_fail(ctx, TNE_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.errstr ? g_nccl.errstr(rc ? rc : rc2) : "?");
    return TNE_OK;
}

// Blocks dealt round-robin to the ranks (block b belongs to rank b % world), `world` consecutive blocks per NCCL call,
// in place (recvbuff = sendbuff + rank * count for the reduce-scatter, sendbuff = recvbuff + rank * count for the
// all-gather: the in-place forms NCCL documents).
static int comm_blocks(tne_ctx* ctx, cplx* p, int64_t block, int64_t nblocks, bool gather) {
    if (ctx->world <= 1 || block <= 0 || nblocks <= 0) return TNE_OK;
    if (!ctx->nccl) return tne_fail(ctx, TNE_ERR_NCCL, "tne_comm_init has not been called");
    if (!g_nccl.reduce_scatter || !g_nccl.all_gather) return tne_fail(ctx, TNE_ERR_NCCL, "libnccl lacks ncclReduceScatter / ncclAllGather");
    if (nblocks % ctx->world != 0) return tne_fail(ctx, TNE_ERR_INVALID, "%lld blocks cannot be dealt to %d ranks", (long long)nblocks, ctx->world);
    const int NCCL_DOUBLE = 8, NCCL_SUM = 0;
    const size_t count = (size_t)(2 * block);
    int rc = g_nccl.group_start();
    for (int64_t j = 0; j < nblocks / ctx->world && rc == 0; ++j) {
        cplx* base = p + j * ctx->world * block;
        cplx* mine = base + (int64_t)ctx->rank * block;
        rc = gather ? g_nccl.all_gather(mine, base, count, NCCL_DOUBLE, ctx->nccl, comm_stream_of(ctx))
                    : g_nccl.reduce_scatter(base, mine, count, NCCL_DOUBLE, NCCL_SUM, ctx->nccl, comm_stream_of(ctx));
    }
    int rc2 = g_nccl.group_end();
    if (rc != 0 || rc2 != 0)
        return tne_fail(ctx, TNE_ERR_NCCL, "%s failed: %s", gather ? "ncclAllGather" : "ncclReduceScatter", g_nccl.errstr ? g_nccl.errstr(rc ? rc : rc2) : "?");
    return TNE_OK;
}
int comm_reduce_scatter_blocks(tne_ctx* ctx, cplx* p, int64_t block, int64_t nblocks) { return comm_blocks(ctx, p, block, nblocks, false); }
int comm_all_gather_blocks(tne_ctx* ctx, cplx* p, int64_t block, int64_t nblocks) { return comm_blocks(ctx, p, block, nblocks, true); }

extern "C" int tne_allreduce_sum(tne_ctx* ctx, int n, const tne_buf* bufs) {
    TNE_ENTER(ctx);
    if (ctx->world <= 1) return TNE_OK;
    if (!ctx->nccl) return tne_fail(ctx, TNE_ERR_NCCL, "tne_comm_init has not been called");
    const int NCCL_DOUBLE = 8, NCCL_SUM = 0;
    int rc = g_nccl.group_start();
    for (int i = 0; i < n && rc == 0; ++i) {
        Buf* b = buf_get(ctx, bufs[i]);
        if (!b) { g_nccl.group_end(); return TNE_ERR_INVALID; }
        rc = g_nccl.allreduce(b->ptr(), b->ptr(), (size_t)(2 * b->nelem), NCCL_DOUBLE, NCCL_SUM, ctx->nccl, comm_stream_of(ctx));
    }
    int rc2 = g_nccl.group_end();
    if (rc != 0 || rc2 != 0) return tne_fail(ctx, TNE_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.errstr ? g_nccl.errstr(rc ? rc : rc2) : "?");
    return TNE_OK;
}
