// Which physical warp slot (%warpid; slot % 4 = SM sub-partition) do the warps of co-resident CTAs get?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o warpslots warpslots.cu && ./warpslots THREADS CTAS_PER_SM
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k(int *out, int spin)
{
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if ((threadIdx.x & 31) == 0) {
        int *o = out + 3 * (blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32);
        o[0] = smid; o[1] = warpid; o[2] = threadIdx.x / 32;
    }
}
int main(int argc, char **argv)
{
    int threads = atoi(argv[1]), per_sm = atoi(argv[2]);
    int blocks = 148 * per_sm, wpc = threads / 32;
    int *d, *h = (int *)malloc(blocks * wpc * 12);
    cudaMalloc(&d, blocks * wpc * 12);
    k<<<blocks, threads>>>(d, 2000000);
    cudaMemcpy(h, d, blocks * wpc * 12, cudaMemcpyDeviceToHost);
    for (int b = 0; b < blocks; ++b) {
        if (h[3 * b * wpc] != 0 && h[3 * b * wpc] != 77) continue;
        printf("sm %3d cta %4d: slots", h[3 * b * wpc], b);
        for (int w = 0; w < wpc; ++w) printf(" %2d", h[3 * (b * wpc + w) + 1]);
        printf("   (slot %% 4:");
        for (int w = 0; w < wpc; ++w) printf(" %d", h[3 * (b * wpc + w) + 1] % 4);
        printf(")\n");
    }
    return 0;
}
