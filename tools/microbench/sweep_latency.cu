// sweep_latency.cu -- cycles per pivot of the register-resident unit-lower-triangular sweep used by the solves
// (one warp, panel in shared memory), and of variants.  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_sweep(double *out, long long *cyc, int reps)
{
    extern __shared__ double L[];          // nrow x w column-major
    const int nrow = 84, w = 48, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < nrow * w; i += blockDim.x) L[i] = ((i % nrow) > (i / nrow)) ? 1e-3 * ((i * 7) % 13) : 0.0;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    double y0 = 1.0 + lane, y1 = 2.0 + lane;
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        const double *Lk0 = L + lane, *Lk1 = L + lane + 32;
        int k = 0;
#pragma unroll 4
        for (; k < 32; k++) {
            const double l0 = *Lk0, l1 = *Lk1;
            const double yk = __shfl_sync(0xffffffffu, y0, k);
            y0 -= l0 * yk;
            y1 -= l1 * yk;
            Lk0 += nrow; Lk1 += nrow;
        }
#pragma unroll 4
        for (; k < w; k++) {
            const double l1 = *Lk1;
            const double yk = __shfl_sync(0xffffffffu, y1, k - 32);
            y1 -= l1 * yk;
            Lk1 += nrow;
        }
    }
    long long t1 = clock64();
    // variant: broadcast through shared memory instead of shuffles
    __shared__ double ybuf[64];
    for (int r = 0; r < reps; r++) {
        const double *Lk0 = L + lane, *Lk1 = L + lane + 32;
        for (int k = 0; k < w; k++) {
            if (lane == (k & 31)) ybuf[k] = k < 32 ? y0 : y1;
            __syncwarp();
            const double yk = ybuf[k];
            y0 -= Lk0[0] * yk;
            y1 -= Lk1[0] * yk;
            Lk0 += nrow; Lk1 += nrow;
        }
    }
    long long t2 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
    out[lane] = y0 + y1;
}

int main()
{
    double *out; long long *cyc, h[2];
    cudaMalloc(&out, 64 * sizeof(double)); cudaMalloc(&cyc, 2 * sizeof(long long));
    const int reps = 200;
    k_sweep<<<1, 256, 84 * 48 * 8>>>(out, cyc, reps);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("shuffle sweep: %.1f cycles per pivot; shared-memory broadcast sweep: %.1f cycles per pivot (48 pivots, %d reps)\n",
           h[0] / (48.0 * reps), h[1] / (48.0 * reps), reps);
    return 0;
}
