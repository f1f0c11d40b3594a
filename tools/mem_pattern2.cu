// HBM throughput vs the GRANULARITY of a warp's accesses in the sweep's plane pattern: P read planes, Q write planes.
//   gran = 32: one instruction = 32 consecutive rows (512 B) of ONE plane  (tools/mem_pattern.cu)
//   gran = 8 : one instruction = 8 consecutive rows (128 B) of FOUR planes (the DMMA fragment mapping: lane = (row, col))
// rows are dealt to warps in chunks of 4 tiles, interleaved over the grid (like gett_absorb).
#include <cuda_runtime.h>
#include <cstdio>
template <int GRAN_R, int GRAN_W>
__global__ void __launch_bounds__(256) pat(const double2* __restrict__ in, double2* __restrict__ out, long long rows, int P, int Q,
                                           long long rstride, long long wstride) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long ntiles = rows / 32;      // a warp always owns 32 rows per iteration (4 DMMA m-tiles)
    for (long long t = g; t < ntiles; t += nw) {
        double2 acc = make_double2(0, 0);
        if (GRAN_R == 32) {
            for (int p = 0; p < P; ++p) { double2 v = __ldcs(in + p * rstride + t * 32 + lane); acc.x += v.x; acc.y += v.y; }
        } else {
            for (int mt = 0; mt < 4; ++mt)
                for (int p = 0; p < P; p += 4) { double2 v = __ldcs(in + (p + (lane & 3)) * rstride + t * 32 + mt * 8 + (lane >> 2)); acc.x += v.x; acc.y += v.y; }
        }
        if (GRAN_W == 32) {
            for (int q = 0; q < Q; ++q) { acc.x += 1.0; __stcs(out + q * wstride + t * 32 + lane, acc); }
        } else {
            for (int mt = 0; mt < 4; ++mt)
                for (int q = 0; q < Q; q += 4) { acc.x += 1.0; __stcs(out + (q + (lane & 3)) * wstride + t * 32 + mt * 8 + (lane >> 2), acc); }
        }
    }
}
template <int GR, int GW> void run(const double2* in, double2* out, int sms, int P, int Q, long long rows, const char* name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 3;
    pat<GR, GW><<<blocks, 256>>>(in, out, rows, P, Q, rows, rows); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) pat<GR, GW><<<blocks, 256>>>(in, out, rows, P, Q, rows, rows);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double bytes = 16.0 * rows * (P + Q);
    printf("%-28s read gran %2d write gran %2d : %.2f TB/s (%.1f us)\n", name, GR, GW, bytes / (ms * 1e-3) / 1e12, ms * 1e3);
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0); int sms = prop.multiProcessorCount;
    size_t total = (size_t)1 << 27;
    double2 *in, *out; cudaMalloc(&in, total * 16); cudaMalloc(&out, total * 16);
    cudaMemset(in, 0, total * 16); cudaMemset(out, 0, total * 16);
    struct { int P, Q; long long rows; const char* name; } cfgs[] = {{32, 16, 1LL << 20, "32 rd, 16 wr (K=32,N=16)"}, {16, 32, 1LL << 20, "16 rd, 32 wr (K=16,N=32)"}, {4, 8, 1LL << 22, "4 rd, 8 wr (K=4,N=8)"}, {8, 4, 1LL << 22, "8 rd, 4 wr (K=8,N=4)"}};
    for (auto& c : cfgs) {
        run<32, 32>(in, out, sms, c.P, c.Q, c.rows, c.name);
        run<8, 32>(in, out, sms, c.P, c.Q, c.rows, c.name);
        run<32, 8>(in, out, sms, c.P, c.Q, c.rows, c.name);
        run<8, 8>(in, out, sms, c.P, c.Q, c.rows, c.name);
    }
    return 0;
}
