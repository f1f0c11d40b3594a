// HBM throughput of the sweep's access pattern: every warp reads `chunk` bytes from each of P read planes and
// writes `chunk` bytes to each of Q write planes (planes = index values of slow legs, `stride` bytes apart);
// consecutive warps take consecutive chunks.  P = Q = 1 is a plain copy.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__global__ void __launch_bounds__(256) pat(const double2* __restrict__ in, double2* __restrict__ out, long long rows, int P, int Q,
                                           long long rstride, long long wstride, int persistent) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const long long ntiles = rows / 32;   // 32 elements (512 B) per warp per plane
    long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    long long per = (ntiles + nw - 1) / nw;
    long long beg = persistent == 2 ? w * per : w, end = persistent == 2 ? min(ntiles, beg + per) : ntiles, step = persistent == 2 ? 1 : nw;
    for (long long t = beg; t < end; t += step) {
        double2 acc = make_double2(0, 0);
        for (int p = 0; p < P; ++p) { double2 v = __ldcs(in + p * rstride + t * 32 + lane); acc.x += v.x; acc.y += v.y; }
        for (int q = 0; q < Q; ++q) { acc.x += 1.0; __stcs(out + q * wstride + t * 32 + lane, acc); }
    }
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0); int sms = prop.multiProcessorCount;
    size_t total = (size_t)3 << 30;  // elements of 16 B: 48 GiB? no: keep 1.5 Gi elements = 24 GiB each
    total = (size_t)1 << 29;         // 512 Mi elements = 8 GiB per buffer
    double2 *in, *out; cudaMalloc(&in, total * 16); cudaMalloc(&out, total * 16);
    cudaMemset(in, 0, total * 16); cudaMemset(out, 0, total * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int P, Q; long long rows; const char* name; } cfgs[] = {
        {1, 1, 1LL << 26, "copy 1:1"}, {2, 1, 1LL << 25, "2:1"}, {1, 2, 1LL << 25, "1:2"},
        {8, 4, 1LL << 23, "8 rd planes, 4 wr planes"}, {32, 16, 1LL << 20, "32 rd, 16 wr (K=32,N=16, M=1M)"},
        {16, 32, 1LL << 20, "16 rd, 32 wr (K=16,N=32)"}, {8, 16, 1LL << 22, "8 rd, 16 wr"}, {4, 8, 1LL << 22, "4 rd, 8 wr (K=4,N=8, M=4M)"},
    };
    for (auto& c : cfgs) {
        for (int pad = 0; pad < 2; ++pad) for (int pers = 0; pers < 3; ++pers) {
            long long rs = c.rows + (pad ? 24 : 0), ws = c.rows + (pad ? 24 : 0);
            if ((size_t)c.P * rs > total || (size_t)c.Q * ws > total) continue;
            long long ntiles = c.rows / 32;
            int blocks = pers ? sms * 8 : (int)((ntiles + 7) / 8);
            pat<<<blocks, 256>>>(in, out, c.rows, c.P, c.Q, rs, ws, pers); cudaDeviceSynchronize();
            cudaEventRecord(e0);
            for (int r = 0; r < 5; ++r) pat<<<blocks, 256>>>(in, out, c.rows, c.P, c.Q, rs, ws, pers);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
            double bytes = 16.0 * c.rows * (c.P + c.Q);
            printf("%-36s pad %d %-10s : %.2f TB/s (%.0f MB, %.1f us)\n", c.name, pad, pers == 0 ? "one-shot" : (pers == 1 ? "persist-il" : "persist-blk"), bytes / (ms * 1e-3) / 1e12, bytes / 1e6, ms * 1e3);
        }
    }
    return 0;
}
