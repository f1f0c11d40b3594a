// Context, device tensor pool and the elementwise kernels (K4 of SURVEY.md 2c):
// the Base/broadcast methods src/TensorAD/rules.jl and src/peps.jl:283-295 call on tensors.
#include <cstdarg>
#include <cstdlib>
#include <algorithm>
#include <cmath>

#include "tne_internal.h"

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local std::string g_noctx_err;

int tne_fail(tne_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_noctx_err = buf;
    return code;
}

// ---------------------------------------------------------------------------------------
// pool: size-bucketed caching allocator (stream-ordered: one stream per context, so a block
// freed by the host can be handed out again immediately -- later kernels run after earlier ones)
// ---------------------------------------------------------------------------------------
static int64_t round_size(int64_t bytes) {
    if (bytes < 512) return 512;
    if (bytes <= (1 << 20)) { int64_t r = 512; while (r < bytes) r <<= 1; return r; }
    const int64_t g = 2 << 20;
    return (bytes + g - 1) / g * g;
}

static int pool_alloc(tne_ctx* ctx, int64_t bytes, void** out, int64_t* rounded) {
    int64_t r = round_size(bytes);
    auto it = ctx->free_blocks.find(r);
    if (it != ctx->free_blocks.end()) {
        *out = it->second;
        ctx->free_blocks.erase(it);
    } else if (ctx->plan_only) {
        *out = calloc(1, (size_t)r);   // planner context: host placeholder, never dereferenced by a kernel
        if (!*out) return tne_fail(ctx, TNE_ERR_OOM, "planner: host allocation failed");
        ctx->reserved += r;
    } else {
        cudaError_t e = cudaMalloc(out, r);
        if (e != cudaSuccess) {
            cudaGetLastError();
            // release cached blocks and the workspace, then retry once
            for (auto& kv : ctx->free_blocks) { cudaFree(kv.second); ctx->reserved -= kv.first; }
            ctx->free_blocks.clear();
            e = cudaMalloc(out, r);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return tne_fail(ctx, TNE_ERR_OOM, "device pool: cannot allocate %lld bytes (in use %lld)",
                                (long long)r, (long long)ctx->in_use);
            }
        }
        ctx->reserved += r;
    }
    ctx->in_use += r;
    ctx->peak = std::max(ctx->peak, ctx->in_use);
    *rounded = r;
    return TNE_OK;
}

Block::~Block() {
    if (ptr && ctx) {
        ctx->free_blocks.emplace(bytes, ptr);
        ctx->in_use -= bytes;
    }
}

int buf_new(tne_ctx* ctx, const std::vector<int64_t>& dims, tne_buf* out) {
    int64_t n = 1;
    for (auto d : dims) {
        if (d < 0) return tne_fail(ctx, TNE_ERR_INVALID, "negative dimension");
        n *= d;
    }
    if (dims.size() > TNE_MAX_RANK) return tne_fail(ctx, TNE_ERR_INVALID, "rank > %d", TNE_MAX_RANK);
    auto blk = std::make_shared<Block>();
    void* p = nullptr;
    int64_t r = 0;
    TNE_TRY(pool_alloc(ctx, std::max<int64_t>(n, 1) * (int64_t)sizeof(cplx), &p, &r));
    blk->ptr = p;
    blk->bytes = r;
    blk->ctx = ctx;
    Buf b;
    b.block = blk;
    b.dims = dims;
    b.nelem = n;
    int64_t h = ctx->next_handle++;
    ctx->bufs.emplace(h, std::move(b));
    *out = h;
    return TNE_OK;
}

Buf* buf_get(tne_ctx* ctx, tne_buf h) {
    auto it = ctx->bufs.find(h);
    if (it == ctx->bufs.end()) {
        tne_fail(ctx, TNE_ERR_INVALID, "unknown tensor handle %lld", (long long)h);
        return nullptr;
    }
    return &it->second;
}

int buf_view(tne_ctx* ctx, const Buf& src, int64_t offset, const std::vector<int64_t>& dims, tne_buf* out) {
    Buf b;
    b.block = src.block;
    b.offset = src.offset + offset;
    b.dims = dims;
    b.nelem = 1;
    for (auto d : dims) b.nelem *= d;
    int64_t h = ctx->next_handle++;
    ctx->bufs.emplace(h, std::move(b));
    *out = h;
    return TNE_OK;
}

int ws_reserve(tne_ctx* ctx, int64_t bytes) {
    if (bytes <= ctx->ws_bytes) return TNE_OK;
    if (ctx->plan_only) { ctx->ws_bytes = bytes; return TNE_OK; }
    if (ctx->ws) { cudaFree(ctx->ws); ctx->ws = nullptr; ctx->ws_bytes = 0; }
    const int64_t g = 64 << 20;
    int64_t r = (bytes + g - 1) / g * g;
    cudaError_t e = cudaMalloc(&ctx->ws, r);
    if (e != cudaSuccess) {
        cudaGetLastError();
        for (auto& kv : ctx->free_blocks) { cudaFree(kv.second); ctx->reserved -= kv.first; }
        ctx->free_blocks.clear();
        e = cudaMalloc(&ctx->ws, r);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ctx->ws = nullptr;
            return tne_fail(ctx, TNE_ERR_OOM, "cannot reserve a %lld-byte workspace", (long long)r);
        }
    }
    ctx->ws_bytes = r;
    return TNE_OK;
}

// ---------------------------------------------------------------------------------------
// C ABI: context
// ---------------------------------------------------------------------------------------
extern "C" int tne_version(void) { return 100; }

extern "C" int tne_ctx_create(int device, tne_ctx** out) {
    if (!out) return TNE_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        const char* po = getenv("TNE_PLAN_ONLY");
        if (po && po[0] == '1') {
            // planner context for tools/plan_dump.py: builds and costs tree programs on a machine without a GPU.
            // It computes NOTHING: every kernel launch is a no-op, tne_download and the tree entry points fail.
            tne_ctx* ctx = new tne_ctx();
            ctx->plan_only = true;
            ctx->total_mem = 180LL << 30;
            ctx->mem_limit = (int64_t)(0.85 * (double)ctx->total_mem);
            *out = ctx;
            return TNE_OK;
        }
        return tne_fail(nullptr, TNE_ERR_CUDA, "no CUDA device: this engine has no CPU path");
    }
    if (device < 0 || device >= n) return tne_fail(nullptr, TNE_ERR_INVALID, "device %d out of range", device);
    tne_ctx* ctx = new tne_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return tne_fail(nullptr, TNE_ERR_CUDA, "cudaSetDevice failed"); }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10) {
        delete ctx;
        return tne_fail(nullptr, TNE_ERR_UNSUPPORTED, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->total_mem = (int64_t)prop.totalGlobalMem;
    ctx->mem_limit = (int64_t)(0.85 * (double)prop.totalGlobalMem);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return tne_fail(nullptr, TNE_ERR_CUDA, "cudaStreamCreate failed");
    }
    cplx one = make_double2(1.0, 0.0);
    if (cudaMalloc(&ctx->one_dev, sizeof(cplx)) != cudaSuccess ||
        cudaMemcpy(ctx->one_dev, &one, sizeof(cplx), cudaMemcpyHostToDevice) != cudaSuccess) {
        delete ctx;
        return tne_fail(nullptr, TNE_ERR_CUDA, "device constant setup failed");
    }
    *out = ctx;
    return TNE_OK;
}

extern "C" int tne_ctx_destroy(tne_ctx* ctx) {
    if (!ctx) return TNE_OK;
    if (ctx->plan_only) {   // host placeholders; leaked blocks of a debugging tool are not worth tracking
        for (auto& kv : ctx->bufs) if (kv.second.block) kv.second.block->ctx = nullptr;
        ctx->bufs.clear();
        delete ctx;
        return TNE_OK;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm_stream) { cudaStreamSynchronize(ctx->comm_stream); cudaStreamDestroy(ctx->comm_stream); }
    if (ctx->ev_comm_ready) cudaEventDestroy(ctx->ev_comm_ready);
    if (ctx->ev_comm_done) cudaEventDestroy(ctx->ev_comm_done);
    ctx->plan_cache.clear();
    comm_destroy(ctx);
    for (auto& kv : ctx->bufs) if (kv.second.block) kv.second.block->ctx = nullptr;  // freed below
    std::vector<void*> ptrs;
    for (auto& kv : ctx->bufs) if (kv.second.block && kv.second.block->ptr) ptrs.push_back(kv.second.block->ptr);
    std::sort(ptrs.begin(), ptrs.end());
    ptrs.erase(std::unique(ptrs.begin(), ptrs.end()), ptrs.end());
    ctx->bufs.clear();
    for (void* p : ptrs) cudaFree(p);
    for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->one_dev) cudaFree(ctx->one_dev);
    if (ctx->s2b_part) cudaFree(ctx->s2b_part);
    if (ctx->red_part) cudaFree(ctx->red_part);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TNE_OK;
}

extern "C" const char* tne_last_error(tne_ctx* ctx) { return ctx ? ctx->err.c_str() : g_noctx_err.c_str(); }

extern "C" int tne_sync(tne_ctx* ctx) {
    TNE_ENTER(ctx);
    TNE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TNE_OK;
}

extern "C" int tne_set_mem_limit(tne_ctx* ctx, int64_t bytes) {
    TNE_ENTER(ctx);
    if (bytes <= 0) return tne_fail(ctx, TNE_ERR_INVALID, "memory limit must be positive");
    ctx->mem_limit = bytes;
    return TNE_OK;
}

extern "C" int tne_mem_info(tne_ctx* ctx, int64_t* in_use, int64_t* peak, int64_t* workspace) {
    TNE_ENTER(ctx);
    if (in_use) *in_use = ctx->in_use;
    if (peak) *peak = ctx->peak;
    if (workspace) *workspace = ctx->ws_bytes;
    return TNE_OK;
}

extern "C" void* tne_stream(tne_ctx* ctx) { return (void*)ctx->stream; }
extern "C" int64_t tne_launch_count(tne_ctx* ctx) { return ctx->launches; }

extern "C" int tne_plan_cache_info(tne_ctx* ctx, int64_t* entries, int64_t* hits) {
    if (entries) *entries = (int64_t)ctx->plan_cache.size();
    if (hits) *hits = ctx->plan_cache_hits;
    return TNE_OK;
}

extern "C" int tne_last_program_stats(tne_ctx* ctx, double* s) {
    TNE_ENTER(ctx);
    for (int i = 0; i < 6; ++i) s[i] = ctx->stats[i];
    return TNE_OK;
}

extern "C" int tne_profile_collect(tne_ctx* ctx, double* out) {
    TNE_ENTER(ctx);
    for (int i = 0; i < 4 * TNE_PROFILE_CLASSES; ++i) out[i] = 0.0;
    TNE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& r : ctx->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        int c = std::min(std::max(r.cls, 0), TNE_PROFILE_CLASSES - 1);
        out[4 * c] += 1.0; out[4 * c + 1] += ms; out[4 * c + 2] += r.flops; out[4 * c + 3] += r.bytes;
        ctx->event_pool.push_back(r.e0);
        ctx->event_pool.push_back(r.e1);
    }
    ctx->prof.clear();
    return TNE_OK;
}

extern "C" int tne_set_option(tne_ctx* ctx, const char* key, int64_t value) {
    TNE_ENTER(ctx);
    std::string k(key ? key : "");
    if (k == "force_generic_kernel") ctx->opt_force_generic = value;
    else if (k == "debug") ctx->opt_debug = value;
    else if (k == "profile") ctx->opt_profile = value;
    else if (k == "absorb_tma") ctx->opt_absorb_tma = value;
    else if (k == "no_fuse") ctx->opt_no_fuse = value;
    else if (k == "no_site2") ctx->opt_no_site2 = value;
    else if (k == "no_site2b") ctx->opt_no_site2b = value;
    else if (k == "no_gemm") ctx->opt_no_gemm = value;
    else if (k == "comm_overlap") ctx->opt_comm_overlap = value;
    else if (k == "tape_gib") ctx->opt_tape_gib = value;
    else if (k == "no_plan_cache") { ctx->opt_no_plan_cache = value; if (value) ctx->plan_cache.clear(); }
    else if (k == "no_permute") ctx->opt_no_permute = value;
    else if (k == "no_rows") ctx->opt_no_rows = value;
    else if (k == "no_pdl") ctx->opt_no_pdl = value;   // default 1: measured no gain from programmatic dependent launch
    else if (k == "no_reduce_tma") ctx->opt_no_reduce_tma = value;
    else if (k == "pad") ctx->opt_pad = value;
    else if (k == "rs_comm") ctx->opt_rs_comm = value;
    else return tne_fail(ctx, TNE_ERR_INVALID, "unknown option '%s'", k.c_str());
    return TNE_OK;
}

// ---------------------------------------------------------------------------------------
// C ABI: tensors
// ---------------------------------------------------------------------------------------
extern "C" int tne_alloc(tne_ctx* ctx, int rank, const int64_t* dims, tne_buf* out) {
    TNE_ENTER(ctx);
    if (rank < 0 || rank > TNE_MAX_RANK) return tne_fail(ctx, TNE_ERR_INVALID, "bad rank %d", rank);
    std::vector<int64_t> d(dims, dims + rank);
    return buf_new(ctx, d, out);
}

extern "C" int tne_free(tne_ctx* ctx, tne_buf h) {
    TNE_ENTER(ctx);
    if (h == 0) return TNE_OK;
    auto it = ctx->bufs.find(h);
    if (it == ctx->bufs.end()) return tne_fail(ctx, TNE_ERR_INVALID, "free of unknown handle %lld", (long long)h);
    ctx->bufs.erase(it);
    return TNE_OK;
}

extern "C" int tne_shape(tne_ctx* ctx, tne_buf h, int* rank, int64_t* dims) {
    TNE_ENTER(ctx);
    Buf* b = buf_get(ctx, h);
    if (!b) return TNE_ERR_INVALID;
    *rank = (int)b->dims.size();
    for (size_t i = 0; i < b->dims.size(); ++i) dims[i] = b->dims[i];
    return TNE_OK;
}

extern "C" int tne_upload(tne_ctx* ctx, tne_buf h, const void* host) {
    TNE_ENTER(ctx);
    Buf* b = buf_get(ctx, h);
    if (!b) return TNE_ERR_INVALID;
    TNE_CUDA(ctx, cudaMemcpyAsync(b->ptr(), host, b->nelem * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
    // the host buffer is only borrowed for the duration of the call
    TNE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TNE_OK;
}

extern "C" int tne_download(tne_ctx* ctx, tne_buf h, void* host) {
    TNE_ENTER(ctx);
    Buf* b = buf_get(ctx, h);
    if (!b) return TNE_ERR_INVALID;
    if (ctx->plan_only) return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "planner context (TNE_PLAN_ONLY): nothing was computed, nothing to download");
    TNE_CUDA(ctx, cudaMemcpyAsync(host, b->ptr(), b->nelem * sizeof(cplx), cudaMemcpyDeviceToHost, ctx->stream));
    TNE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TNE_OK;
}

extern "C" void* tne_device_ptr(tne_ctx* ctx, tne_buf h) {
    TNE_ENTER(ctx);
    Buf* b = buf_get(ctx, h);
    return b ? (void*)b->ptr() : nullptr;
}

// ---------------------------------------------------------------------------------------
// elementwise kernels (HBM-bound; one complex128 = one 128-bit access per thread and step)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    double d = b.x * b.x + b.y * b.y;
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
__device__ __forceinline__ cplx csqrt_(cplx z) {
    if (z.y == 0.0) return z.x >= 0 ? make_double2(sqrt(z.x), 0.0) : make_double2(0.0, sqrt(-z.x));
    double r = hypot(z.x, z.y);
    double a = sqrt(0.5 * (r + fabs(z.x)));
    double b = 0.5 * z.y / a;
    return z.x >= 0 ? make_double2(a, b) : make_double2(fabs(b), copysign(a, z.y));
}
__device__ __forceinline__ cplx cpow_(cplx z, double p) {
    if (z.y == 0.0 && z.x >= 0.0) return make_double2(pow(z.x, p), 0.0);
    double r = hypot(z.x, z.y), th = atan2(z.y, z.x);
    double rp = pow(r, p);
    double s, c;
    sincos(p * th, &s, &c);
    return make_double2(rp * c, rp * s);
}

__global__ void map_kernel(int op, const cplx* __restrict__ in, cplx* __restrict__ out, int64_t n, double sre, double sim) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        cplx z = in[i], r;
        switch (op) {
            case TNE_COPY: r = z; break;
            case TNE_CONJ: r = make_double2(z.x, -z.y); break;
            case TNE_REAL: r = make_double2(z.x, 0.0); break;
            case TNE_IMAG: r = make_double2(z.y, 0.0); break;
            case TNE_ABS: r = make_double2(hypot(z.x, z.y), 0.0); break;
            case TNE_SQRT: r = csqrt_(z); break;
            case TNE_INV: r = cdiv(make_double2(1.0, 0.0), z); break;
            case TNE_NEG: r = make_double2(-z.x, -z.y); break;
            case TNE_SCALE: r = cmul(z, make_double2(sre, sim)); break;
            case TNE_POW: r = cpow_(z, sre); break;
            case TNE_SIN: r = make_double2(sin(z.x) * cosh(z.y), cos(z.x) * sinh(z.y)); break;
            case TNE_COS: r = make_double2(cos(z.x) * cosh(z.y), -sin(z.x) * sinh(z.y)); break;
            case TNE_SAFE_INV: {
                const double eps = 1e-10;
                if (hypot(z.x, z.y) < eps) z = make_double2(z.x >= 0 ? eps : -eps, 0.0);
                r = cdiv(make_double2(1.0, 0.0), z);
                break;
            }
            default: r = z;
        }
        out[i] = r;
    }
}

__global__ void map2_kernel(int op, const cplx* __restrict__ a, int64_t na, const cplx* __restrict__ b, int64_t nb,
                            cplx* __restrict__ out, int64_t n) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        cplx x = a[na == 1 ? 0 : i], y = b[nb == 1 ? 0 : i], r;
        switch (op) {
            case TNE_ADD: r = make_double2(x.x + y.x, x.y + y.y); break;
            case TNE_SUB: r = make_double2(x.x - y.x, x.y - y.y); break;
            case TNE_MUL: r = cmul(x, y); break;
            default: r = cdiv(x, y);
        }
        out[i] = r;
    }
}

__global__ void fill_zero_kernel(cplx* p, int64_t n) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = make_double2(0.0, 0.0);
}

__global__ void scale_inplace_kernel(cplx* p, int64_t n, double re, double im) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = cmul(p[i], make_double2(re, im));
}

__global__ void axpy_dev_kernel(cplx* y, const cplx* x, int64_t n, const cplx* alpha) {
    cplx a = *alpha;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        cplx v = cmul(a, x[i]);
        y[i] = make_double2(y[i].x + v.x, y[i].y + v.y);
    }
}

__global__ void fill_from_kernel(cplx* p, int64_t n, const cplx* s) {
    cplx v = *s;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

__global__ void fill_value_kernel(cplx* p, int64_t n, double re, double im) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = make_double2(re, im);
}

// block-wise tree reduction; second pass over <= 1024 partials
__global__ void reduce_sum_kernel(const cplx* __restrict__ in, int64_t n, cplx* __restrict__ out) {
    __shared__ double sre[256], sim[256];
    double re = 0, im = 0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { re += in[i].x; im += in[i].y; }
    sre[threadIdx.x] = re; sim[threadIdx.x] = im;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sre[threadIdx.x] += sre[threadIdx.x + s]; sim[threadIdx.x] += sim[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = make_double2(sre[0], sim[0]);
}

static int grid_for(tne_ctx* ctx, int64_t n) {
    int64_t b = (n + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    return (int)std::max<int64_t>(1, std::min(b, cap));
}

int launch_fill_zero(tne_ctx* ctx, cplx* p, int64_t n) {
    if (n <= 0) return TNE_OK;
    TNE_CUDA(ctx, cudaMemsetAsync(p, 0, n * sizeof(cplx), ctx->stream));
    return TNE_OK;
}

int launch_fill_value(tne_ctx* ctx, cplx* p, int64_t n, double re, double im) {
    if (n <= 0) return TNE_OK;
    fill_value_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(p, n, re, im);
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}

int launch_scale_inplace(tne_ctx* ctx, cplx* p, int64_t n, double re, double im) {
    if (n <= 0) return TNE_OK;
    scale_inplace_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(p, n, re, im);
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}

int launch_axpy_dev(tne_ctx* ctx, cplx* y, const cplx* x, int64_t n, const cplx* alpha_dev) {
    if (n <= 0) return TNE_OK;
    axpy_dev_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(y, x, n, alpha_dev);
    TNE_LAUNCH_CHECK(ctx);
    return TNE_OK;
}

extern "C" int tne_map(tne_ctx* ctx, int op, tne_buf in, double sre, double sim, tne_buf* out) {
    TNE_ENTER(ctx);
    Buf* a = buf_get(ctx, in);
    if (!a) return TNE_ERR_INVALID;
    std::vector<int64_t> dims = a->dims;
    int64_t n = a->nelem;
    const cplx* src = a->ptr();
    tne_buf h;
    TNE_TRY(buf_new(ctx, dims, &h));
    if (n > 0) {
        map_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(op, src, buf_get(ctx, h)->ptr(), n, sre, sim);
        TNE_LAUNCH_CHECK(ctx);
    }
    *out = h;
    return TNE_OK;
}

extern "C" int tne_map2(tne_ctx* ctx, int op, tne_buf ha, tne_buf hb, tne_buf* out) {
    TNE_ENTER(ctx);
    Buf* a = buf_get(ctx, ha);
    Buf* b = buf_get(ctx, hb);
    if (!a || !b) return TNE_ERR_INVALID;
    if (a->nelem != b->nelem && a->nelem != 1 && b->nelem != 1)
        return tne_fail(ctx, TNE_ERR_INVALID, "map2: sizes %lld and %lld do not broadcast", (long long)a->nelem, (long long)b->nelem);
    const Buf* big = (a->nelem >= b->nelem && !(a->nelem == 1 && b->dims.size() > a->dims.size())) ? a : b;
    std::vector<int64_t> dims = big->dims;
    int64_t n = std::max(a->nelem, b->nelem);
    const cplx *pa = a->ptr(), *pb = b->ptr();
    int64_t na = a->nelem, nb = b->nelem;
    tne_buf h;
    TNE_TRY(buf_new(ctx, dims, &h));
    if (n > 0) {
        map2_kernel<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(op, pa, na, pb, nb, buf_get(ctx, h)->ptr(), n);
        TNE_LAUNCH_CHECK(ctx);
    }
    *out = h;
    return TNE_OK;
}

extern "C" int tne_reduce_sum(tne_ctx* ctx, tne_buf in, tne_buf* out) {
    TNE_ENTER(ctx);
    Buf* a = buf_get(ctx, in);
    if (!a) return TNE_ERR_INVALID;
    int64_t n = a->nelem;
    const cplx* src = a->ptr();
    int g = (int)std::min<int64_t>(1024, std::max<int64_t>(1, (n + 255) / 256));
    tne_buf tmp, res;
    TNE_TRY(buf_new(ctx, {(int64_t)g}, &tmp));
    TNE_TRY(buf_new(ctx, {}, &res));
    reduce_sum_kernel<<<g, 256, 0, ctx->stream>>>(src, n, buf_get(ctx, tmp)->ptr());
    TNE_LAUNCH_CHECK(ctx);
    reduce_sum_kernel<<<1, 256, 0, ctx->stream>>>(buf_get(ctx, tmp)->ptr(), g, buf_get(ctx, res)->ptr());
    TNE_LAUNCH_CHECK(ctx);
    tne_free(ctx, tmp);
    *out = res;
    return TNE_OK;
}

extern "C" int tne_fill(tne_ctx* ctx, tne_buf scalar0, int rank, const int64_t* dims, tne_buf* out) {
    TNE_ENTER(ctx);
    Buf* s = buf_get(ctx, scalar0);
    if (!s) return TNE_ERR_INVALID;
    if (s->nelem != 1) return tne_fail(ctx, TNE_ERR_INVALID, "fill: source must have one element");
    const cplx* sp = s->ptr();
    tne_buf h;
    TNE_TRY(buf_new(ctx, std::vector<int64_t>(dims, dims + rank), &h));
    Buf* o = buf_get(ctx, h);
    if (o->nelem > 0) {
        fill_from_kernel<<<grid_for(ctx, o->nelem), 256, 0, ctx->stream>>>(o->ptr(), o->nelem, sp);
        TNE_LAUNCH_CHECK(ctx);
    }
    *out = h;
    return TNE_OK;
}

extern "C" int tne_slice(tne_ctx* ctx, tne_buf in, int64_t lo, int64_t count, tne_buf* out) {
    TNE_ENTER(ctx);
    Buf* a = buf_get(ctx, in);
    if (!a) return TNE_ERR_INVALID;
    if (lo < 0 || count < 0 || lo + count > a->nelem) return tne_fail(ctx, TNE_ERR_INVALID, "slice [%lld, %lld) out of range (%lld)", (long long)lo, (long long)(lo + count), (long long)a->nelem);
    Buf src = *a;
    return buf_view(ctx, src, lo, {count}, out);
}

extern "C" int tne_embed(tne_ctx* ctx, tne_buf y, int64_t lo, int64_t total, tne_buf* out) {
    TNE_ENTER(ctx);
    Buf* a = buf_get(ctx, y);
    if (!a) return TNE_ERR_INVALID;
    if (lo < 0 || lo + a->nelem > total) return tne_fail(ctx, TNE_ERR_INVALID, "embed range out of bounds");
    const cplx* src = a->ptr();
    int64_t n = a->nelem;
    tne_buf h;
    TNE_TRY(buf_new(ctx, {total}, &h));
    Buf* o = buf_get(ctx, h);
    TNE_TRY(launch_fill_zero(ctx, o->ptr(), total));
    if (n > 0) TNE_CUDA(ctx, cudaMemcpyAsync(o->ptr() + lo, src, n * sizeof(cplx), cudaMemcpyDeviceToDevice, ctx->stream));
    *out = h;
    return TNE_OK;
}

extern "C" int tne_reshape(tne_ctx* ctx, tne_buf in, int rank, const int64_t* dims, tne_buf* out) {
    TNE_ENTER(ctx);
    Buf* a = buf_get(ctx, in);
    if (!a) return TNE_ERR_INVALID;
    std::vector<int64_t> d(dims, dims + rank);
    int64_t n = 1;
    for (auto x : d) n *= x;
    if (n != a->nelem) return tne_fail(ctx, TNE_ERR_INVALID, "reshape: %lld elements into %lld", (long long)a->nelem, (long long)n);
    Buf src = *a;
    return buf_view(ctx, src, 0, d, out);
}

extern "C" int tne_concat(tne_ctx* ctx, int n, const tne_buf* in, tne_buf* out) {
    TNE_ENTER(ctx);
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
        Buf* b = buf_get(ctx, in[i]);
        if (!b) return TNE_ERR_INVALID;
        total += b->nelem;
    }
    tne_buf h;
    TNE_TRY(buf_new(ctx, {total}, &h));
    cplx* dst = buf_get(ctx, h)->ptr();
    int64_t off = 0;
    for (int i = 0; i < n; ++i) {
        Buf* b = buf_get(ctx, in[i]);
        if (b->nelem > 0)
            TNE_CUDA(ctx, cudaMemcpyAsync(dst + off, b->ptr(), b->nelem * sizeof(cplx), cudaMemcpyDeviceToDevice, ctx->stream));
        off += b->nelem;
    }
    *out = h;
    return TNE_OK;
}
