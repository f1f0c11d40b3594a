// FP64 roofline denominators for this pool's B200: DFMA and DMMA issue peaks plus
// cuBLAS DGEMM/ZGEMM (burst = best of 10, sustained = back-to-back for ~4 s).
// Writes one JSON object to stdout. Build: see tools/Makefile (nvcc, sm_100a, -lcublas).
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cuComplex.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

template<int ILP>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
    double av = a + threadIdx.x * 1e-6, bv = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(av), "d"(bv));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1) { float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms; }

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* d; CK(cudaMalloc(&d, 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    // DFMA
    {
        const int ILP = 8, iters = 1 << 16; int blocks = sms * 8;
        dfma_kernel<ILP><<<blocks, 256>>>(d, 1000, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
        double best = 0;
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0)); dfma_kernel<ILP><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            double fl = 2.0 * ILP * iters * 256.0 * blocks; best = std::max(best, fl / (time_ms(e0, e1) * 1e-3) / 1e12);
        }
        printf(", \"dfma_tflops\": %.3f", best);
    }
    // DMMA
    {
        const int ILP = 8, iters = 1 << 14; int blocks = sms * 8;
        dmma_kernel<ILP><<<blocks, 256>>>(d, 1000, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
        double best = 0;
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0)); dmma_kernel<ILP><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            double fl = 2.0 * 8 * 8 * 4 * (double)ILP * iters * (256.0 / 32) * blocks; best = std::max(best, fl / (time_ms(e0, e1) * 1e-3) / 1e12);
        }
        printf(", \"dmma_tflops\": %.3f", best);
    }
    // DMMA + DFMA concurrently issued is not measured; cuBLAS below.
    cublasHandle_t h; cublasCreate(&h);
    {
        int n = 8192; size_t bytes = (size_t)n * n * sizeof(double);
        double *A, *B, *C; CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
        CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes));
        double one = 1, zero = 0;
        for (int w = 0; w < 3; ++w) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
        CK(cudaDeviceSynchronize());
        double best = 0; double fl = 2.0 * n * (double)n * n;
        for (int r = 0; r < 10; ++r) {
            CK(cudaEventRecord(e0)); cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            best = std::max(best, fl / (time_ms(e0, e1) * 1e-3) / 1e12);
        }
        printf(", \"dgemm8192_burst_tflops\": %.3f", best);
        int reps = (int)(4.0 / (fl / (best * 1e12))) + 1;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        printf(", \"dgemm8192_sustained_tflops\": %.3f", reps * fl / (time_ms(e0, e1) * 1e-3) / 1e12);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    {
        int n = 4096; size_t bytes = (size_t)n * n * sizeof(cuDoubleComplex);
        cuDoubleComplex *A, *B, *C; CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
        CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes));
        cuDoubleComplex one = make_cuDoubleComplex(1, 0), zero = make_cuDoubleComplex(0, 0);
        for (int w = 0; w < 3; ++w) cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
        CK(cudaDeviceSynchronize());
        double best = 0; double fl = 8.0 * n * (double)n * n;
        for (int r = 0; r < 10; ++r) {
            CK(cudaEventRecord(e0)); cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            best = std::max(best, fl / (time_ms(e0, e1) * 1e-3) / 1e12);
        }
        printf(", \"zgemm4096_burst_tflops\": %.3f", best);
        int reps = (int)(4.0 / (fl / (best * 1e12))) + 1;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        printf(", \"zgemm4096_sustained_tflops\": %.3f", reps * fl / (time_ms(e0, e1) * 1e-3) / 1e12);
        // skinny ZGEMM of the hot-path shape: M=2^22, N=16, K=32
        cudaFree(A); cudaFree(B); cudaFree(C);
        size_t M = 1 << 22; int N = 16, K = 32;
        CK(cudaMalloc(&A, M * K * 16)); CK(cudaMalloc(&B, (size_t)K * N * 16)); CK(cudaMalloc(&C, M * N * 16));
        CK(cudaMemset(A, 0, M * K * 16)); CK(cudaMemset(B, 0, (size_t)K * N * 16));
        for (int w = 0; w < 3; ++w) cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, (int)M, N, K, &one, A, (int)M, B, K, &zero, C, (int)M);
        CK(cudaDeviceSynchronize());
        best = 0; fl = 8.0 * M * N * K;
        float bestms = 1e9;
        for (int r = 0; r < 10; ++r) {
            CK(cudaEventRecord(e0)); cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, (int)M, N, K, &one, A, (int)M, B, K, &zero, C, (int)M); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            bestms = std::min(bestms, time_ms(e0, e1));
        }
        printf(", \"zgemm_skinny_M4M_N16_K32_tflops\": %.3f, \"zgemm_skinny_gbs\": %.1f", fl / (bestms * 1e-3) / 1e12, (M * (double)(K + N) * 16) / (bestms * 1e-3) / 1e9);
    }
    printf("}\n");
    return 0;
}
