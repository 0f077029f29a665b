// fp64_rates.cu -- B200 microbenchmarks that size the LDL^T kernels: DFMA vs DMMA (mma.sync m8n8k4 f64) throughput,
// dependent-chain latencies (DFMA, 1/x, shared-memory load, __syncthreads, shfl).  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_dmma(double *out, int iters, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lat(double *out, long long *cyc, int iters, double a, double b)
{
    __shared__ double sm[256];
    sm[threadIdx.x] = threadIdx.x == 0 ? 0.0 : 1.0;   // sm[0] = 0 -> pointer chase stays at 0
    __syncthreads();
    double x = 1.0 + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) x = fma(x, a, b);
    long long t1 = clock64();
    double y = x;
    for (int i = 0; i < iters; i++) y = 1.0 / (y + a);
    long long t2 = clock64();
    int idx = (int)sm[threadIdx.x & 0];
    for (int i = 0; i < iters; i++) idx = (int)sm[idx];
    long long t3 = clock64();
    for (int i = 0; i < iters; i++) __syncthreads();
    long long t4 = clock64();
    double z = y;
    for (int i = 0; i < iters; i++) z = __shfl_sync(0xffffffffu, z, (i + 1) & 31) + a;
    long long t5 = clock64();
    double c0 = z, c1 = y;
    for (int i = 0; i < iters; i++) dmma(c0, c1, a, b);
    long long t6 = clock64();
    double q = c0;
    for (int i = 0; i < iters; i++) q = rsqrt(q + 2.0);
    long long t7 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6;
    }
    out[threadIdx.x] = x + y + idx + z + c0 + c1 + q;
}

int main()
{
    double *out;
    long long *cyc, h[8];
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
    cudaMalloc(&cyc, 8 * sizeof(long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int threads : {128, 256, 512, 1024}) {
        const int blocks = 148 * (1024 / threads), iters = 20000;
        k_dfma<<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA  %4d thr/blk x %d blk: %.2f TFLOP/s\n", threads, blocks, 2.0 * blocks * threads * 8.0 * iters / (ms * 1e-3) / 1e12);
        k_dmma<<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        k_dmma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA  %4d thr/blk x %d blk: %.2f TFLOP/s\n", threads, blocks, 2.0 * blocks * (threads / 32) * 8.0 * 256.0 * iters / (ms * 1e-3) / 1e12);
    }
    for (int threads : {32, 256}) {
        const int iters = 2000;
        k_lat<<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("latency (cycles, %d threads): dfma %.1f  div %.1f  lds-chase %.1f  syncthreads %.1f  shfl+add %.1f  dmma %.1f  rsqrt %.1f\n", threads,
               h[0] / (double)iters, h[1] / (double)iters, h[2] / (double)iters, h[3] / (double)iters, h[4] / (double)iters, h[5] / (double)iters, h[6] / (double)iters);
    }
    // one CTA per SM with few warps: per-SM DFMA rate as a function of warps (latency hiding)
    for (int threads : {32, 64, 128, 256}) {
        const int iters = 20000;
        cudaEventRecord(e0);
        k_dfma<<<148, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA 1 CTA/SM %4d thr (ILP 8): %.2f FMA/clk/SM at 1.965 GHz\n", threads, 148.0 * threads * 8.0 * iters / (ms * 1e-3) / 148 / 1.965e9);
        cudaEventRecord(e0);
        k_dmma<<<148, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA 1 CTA/SM %4d thr (8 indep): %.2f FMA/clk/SM\n", threads, 148.0 * (threads / 32) * 8.0 * 256 * iters / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
