// DMMA issue-rate study: TFLOP/s as a function of warps per SM and independent accumulator chains per warp,
// with register operands and with the B fragment re-read from shared memory before every DMMA.
#include <cuda_runtime.h>
#include <cstdio>
template<int ILP, bool SMEM>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
    __shared__ double sb[64 * 36];
    for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) sb[i] = b + i * 1e-9;
    __syncthreads();
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
    double av = a + threadIdx.x * 1e-6, bv = b;
    const int lane = threadIdx.x & 31;
    const double* p = sb + (lane >> 2) * 36 + (lane & 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (SMEM) bv = p[(i & 3) * 8 * 36 + 4 * (it & 7)];
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(av), "d"(bv));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}
template<int ILP, bool SMEM> void run(double* d, int sms, int bps, int threads) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 1 << 13; int blocks = sms * bps;
    k<ILP, SMEM><<<blocks, threads>>>(d, 100, 1.0000001, 1e-9); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<ILP, SMEM><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 8 * 8 * 4 * (double)ILP * iters * (threads / 32.0) * blocks;
    printf("warps/SM %2d ILP %d smemB %d : %.2f TFLOP/s\n", bps * threads / 32, ILP, (int)SMEM, fl / (ms * 1e-3) / 1e12);
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0); int sms = prop.multiProcessorCount;
    double* d; cudaMalloc(&d, 1024);
    int cfg[][2] = {{1, 128}, {1, 256}, {2, 256}, {3, 256}, {4, 256}, {8, 256}};
    for (auto& c : cfg) {
        run<1, false>(d, sms, c[0], c[1]); run<2, false>(d, sms, c[0], c[1]); run<4, false>(d, sms, c[0], c[1]); run<8, false>(d, sms, c[0], c[1]);
        run<4, true>(d, sms, c[0], c[1]); run<8, true>(d, sms, c[0], c[1]);
    }
    return 0;
}
