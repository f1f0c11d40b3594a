// DMMA issue-rate study 2: distinct A registers per k-step (like the absorb kernel), B from shared memory,
// NCH independent accumulator chains.  Is the 37 TF register-operand peak reachable with changing operands?
#include <cuda_runtime.h>
#include <cstdio>
template<int KS, int NCH, bool SMEM, bool VARA>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a0, double b0) {
    __shared__ double sb[64 * 68];
    for (int i = threadIdx.x; i < 64 * 68; i += blockDim.x) sb[i] = b0 + i * 1e-9;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const double* p = sb + (lane >> 2) * 68 + (lane & 3);
    double a[KS];
#pragma unroll
    for (int i = 0; i < KS; ++i) a[i] = a0 + (threadIdx.x + i) * 1e-6;
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
        double c[NCH][2];
#pragma unroll
        for (int u = 0; u < NCH; ++u) { c[u][0] = 0; c[u][1] = 0; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            double b[NCH];
#pragma unroll
            for (int u = 0; u < NCH; ++u) b[u] = SMEM ? p[u * 8 * 68 + 4 * ks] : b0;
#pragma unroll
            for (int u = 0; u < NCH; ++u)
                asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                    : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(VARA ? a[ks] : a[0]), "d"(b[u]));
        }
#pragma unroll
        for (int u = 0; u < NCH; ++u) acc += c[u][0] + c[u][1];
#pragma unroll
        for (int i = 0; i < KS; ++i) a[i] += 1e-9;   // "new tile"
    }
    if (acc == 12345.678) out[0] = acc;
}
template<int KS, int NCH, bool SMEM, bool VARA> void run(double* d, int sms, int bps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 1 << 11; int blocks = sms * bps;
    k<KS, NCH, SMEM, VARA><<<blocks, 256>>>(d, 10, 1.0000001, 1e-9); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<KS, NCH, SMEM, VARA><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 512.0 * KS * NCH * iters * 8.0 * blocks;
    printf("warps/SM %2d KS %2d chains %d smemB %d varA %d : %.2f TFLOP/s\n", bps * 8, KS, NCH, (int)SMEM, (int)VARA, fl / (ms * 1e-3) / 1e12);
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0); int sms = prop.multiProcessorCount;
    double* d; cudaMalloc(&d, 1024);
    for (int bps = 1; bps <= 4; ++bps) {
        run<16, 4, false, false>(d, sms, bps); run<16, 4, false, true>(d, sms, bps); run<16, 4, true, false>(d, sms, bps); run<16, 4, true, true>(d, sms, bps);
        run<16, 2, true, true>(d, sms, bps); run<16, 1, true, true>(d, sms, bps); run<8, 4, true, true>(d, sms, bps); run<16, 8, true, true>(d, sms, bps);
    }
    return 0;
}
