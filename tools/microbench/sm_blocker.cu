// Occupies `n` SMs with one spinning CTA each (all of the SM's shared memory, so nothing else fits beside it) until released:
// lets a kernel be timed at full per-SM residency on a fraction of the chip (is a slowdown under load SM-local or chip-wide?).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --shared -Xcompiler -fPIC sm_blocker.cu -o sm_blocker.so
#include <cuda_runtime.h>
extern __shared__ char sm_block_smem[];
__global__ void k_sm_blocker(volatile int *flag, int *arrived)
{
    if (threadIdx.x == 0) {
        sm_block_smem[0] = 1;
        atomicAdd(arrived, 1);
        while (*flag == 0) __nanosleep(2000);
    }
}
static cudaStream_t g_stream = nullptr, g_copy = nullptr;
static int *g_flag = nullptr, *g_arrived = nullptr;
extern "C" int blocker_start(int n)
{
    const int smem = 200 * 1024;
    if (!g_stream) {
        if (cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking) != cudaSuccess) return -1;
        if (cudaStreamCreateWithFlags(&g_copy, cudaStreamNonBlocking) != cudaSuccess) return -1;
        if (cudaMalloc(&g_flag, sizeof(int)) != cudaSuccess || cudaMalloc(&g_arrived, sizeof(int)) != cudaSuccess) return -2;
        if (cudaFuncSetAttribute(k_sm_blocker, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -3;
    }
    if (cudaMemset(g_flag, 0, sizeof(int)) != cudaSuccess || cudaMemset(g_arrived, 0, sizeof(int)) != cudaSuccess) return -4;
    cudaDeviceSynchronize();
    k_sm_blocker<<<n, 32, smem, g_stream>>>(g_flag, g_arrived);
    if (cudaGetLastError() != cudaSuccess) return -5;
    int seen = 0;
    for (int spin = 0; spin < 200000 && seen < n; spin++) {
        cudaMemcpyAsync(&seen, g_arrived, sizeof(int), cudaMemcpyDeviceToHost, g_copy);
        cudaStreamSynchronize(g_copy);
    }
    return seen;
}
extern "C" int blocker_release()
{
    int one = 1;
    if (cudaMemcpyAsync(g_flag, &one, sizeof(int), cudaMemcpyHostToDevice, g_copy) != cudaSuccess || cudaStreamSynchronize(g_copy) != cudaSuccess) return -1;
    return cudaStreamSynchronize(g_stream) == cudaSuccess ? 0 : -2;
}
