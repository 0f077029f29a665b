// sweep_blocked.cu -- cycles per pivot of the forward sweep variants (one warp, panel in shared memory), alone on an SM
// and with other warps busy.  nvcc -arch=sm_100a -O3 [-maxrregcount=80].
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void sweep4(const double *__restrict__ L, int nrow, int k0, int k1, int lane, double &y0, double &y1)
{
    const int r0 = min(lane, nrow - 1), r1 = min(lane + 32, nrow - 1);
    int kb = k0;
    for (; kb + 4 <= k1; kb += 4) {
        const double *c0 = L + kb * nrow, *c1 = c0 + nrow, *c2 = c1 + nrow, *c3 = c2 + nrow;
        const double l10 = c0[kb + 1], l20 = c0[kb + 2], l30 = c0[kb + 3], l21 = c1[kb + 2], l31 = c1[kb + 3], l32 = c2[kb + 3];
        const double a00 = c0[r0], a01 = c1[r0], a02 = c2[r0], a03 = c3[r0];
        const double a10 = c0[r1], a11 = c1[r1], a12 = c2[r1], a13 = c3[r1];
        const double v0 = __shfl_sync(0xffffffffu, kb < 32 ? y0 : y1, kb & 31);
        double v1 = __shfl_sync(0xffffffffu, kb + 1 < 32 ? y0 : y1, (kb + 1) & 31);
        double v2 = __shfl_sync(0xffffffffu, kb + 2 < 32 ? y0 : y1, (kb + 2) & 31);
        double v3 = __shfl_sync(0xffffffffu, kb + 3 < 32 ? y0 : y1, (kb + 3) & 31);
        v1 -= l10 * v0; v2 -= l20 * v0; v3 -= l30 * v0; y0 -= a00 * v0; y1 -= a10 * v0;
        v2 -= l21 * v1; v3 -= l31 * v1; y0 -= a01 * v1; y1 -= a11 * v1;
        v3 -= l32 * v2; y0 -= a02 * v2; y1 -= a12 * v2;
        y0 -= a03 * v3; y1 -= a13 * v3;
    }
}
__device__ __forceinline__ void sweep1(const double *__restrict__ L, int nrow, int k0, int k1, int lane, double &y0, double &y1)
{
    const double *Lk0 = L + k0 * nrow + min(lane, nrow - 1), *Lk1 = L + k0 * nrow + min(lane + 32, nrow - 1);
#pragma unroll 4
    for (int k = k0; k < k1; k++) {
        const double l0 = *Lk0, l1 = *Lk1;
        const double yk = __shfl_sync(0xffffffffu, k < 32 ? y0 : y1, k & 31);
        y0 -= l0 * yk; y1 -= l1 * yk;
        Lk0 += nrow; Lk1 += nrow;
    }
}

template <int VARIANT>
__global__ void k_sweep(double *out, long long *cyc, int reps, int busy)
{
    extern __shared__ double L[];          // nrow x w column-major
    const int nrow = 84, w = 48, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < nrow * w; i += blockDim.x) L[i] = ((i % nrow) > (i / nrow)) ? 1e-3 * ((i * 7) % 13) : 0.0;
    __syncthreads();
    if (threadIdx.x >= 32) {               // other warps: idle or busy with FMAs + loads
        if (!busy) return;
        double a = threadIdx.x, b = 1.0001;
        for (int r = 0; r < reps * 400; r++) { a = a * b + L[(threadIdx.x + r) % (nrow * w)]; }
        out[threadIdx.x] = a;
        return;
    }
    double y0 = 1.0 + lane, y1 = 2.0 + lane;
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        if (VARIANT == 4) { sweep4(L, nrow, 0, 24, lane, y0, y1); sweep4(L, nrow, 24, 48, lane, y0, y1); }
        else { sweep1(L, nrow, 0, 24, lane, y0, y1); sweep1(L, nrow, 24, 48, lane, y0, y1); }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    out[lane] = y0 + y1;
}

int main()
{
    double *out; long long *cyc, h[1];
    cudaMalloc(&out, 1024 * sizeof(double)); cudaMalloc(&cyc, sizeof(long long));
    const int reps = 200;
    for (int busy = 0; busy < 2; busy++) {
        k_sweep<1><<<1, 256, 84 * 48 * 8>>>(out, cyc, reps, busy);
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("busy=%d pivot-by-pivot sweep: %.1f cycles per pivot\n", busy, h[0] / (48.0 * reps));
        k_sweep<4><<<1, 256, 84 * 48 * 8>>>(out, cyc, reps, busy);
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("busy=%d four-pivot blocks:    %.1f cycles per pivot\n", busy, h[0] / (48.0 * reps));
    }
    return 0;
}
