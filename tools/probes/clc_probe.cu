// clc_probe.cu -- cluster launch control (clusterlaunchcontrol.try_cancel) as a dynamic tile scheduler for a persistent-style kernel:
// the grid has one CTA per tile; a running CTA, after its own tile, cancels a not-yet-launched CTA and does that CTA's tile instead.
// Checks: every tile done exactly once; how many CTAs actually launched; cycles from try_cancel to the response.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o clc_probe clc_probe.cu && ./clc_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) clc_kernel(int *done, int *launched, long long *lat, int spin) {
    __shared__ __align__(16) uint4 resp;
    __shared__ __align__(8) uint64_t bar;
    __shared__ int next;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        atomicAdd(launched, 1);
    }
    __syncthreads();
    int tile = blockIdx.x;
    uint32_t parity = 0;
    long long tsum = 0;
    int n = 0;
    while (tile >= 0) {
        if (threadIdx.x == 0) {
            long long t0 = clock64();
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(s32(&bar)) : "memory");
            asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(s32(&resp)),
                         "r"(s32(&bar))
                         : "memory");
            atomicAdd(&done[tile], 1);
            for (volatile int i = 0; i < spin; ++i) {}
            asm volatile(
                "{\n\t.reg .pred p;\n\tW_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(&bar)),
                "r"(parity)
                : "memory");
            tsum += clock64() - t0;
            ++n;
            uint32_t valid, x, y, z;
            asm volatile(
                "{\n\t.reg .pred p1;\n\t.reg .b128 r;\n\t"
                "ld.shared.b128 r, [%4];\n\t"
                "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\t"
                "selp.u32 %3, 1, 0, p1;\n\t"
                "mov.u32 %0, 0; mov.u32 %1, 0; mov.u32 %2, 0;\n\t"
                "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, r;\n\t}"
                : "=r"(x), "=r"(y), "=r"(z), "=r"(valid)
                : "r"(s32(&resp))
                : "memory");
            next = valid ? (int)x : -1;
        }
        parity ^= 1;
        __syncthreads();
        tile = next;
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) { lat[0] = tsum; lat[1] = n; }
}

int main() {
    const int ntiles = 4096;
    int *done, *launched;
    long long *lat;
    cudaMalloc(&done, ntiles * sizeof(int));
    cudaMalloc(&launched, sizeof(int));
    cudaMalloc(&lat, 16);
    for (int spin : {0, 2000}) {
        cudaMemset(done, 0, ntiles * sizeof(int));
        cudaMemset(launched, 0, sizeof(int));
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        clc_kernel<<<ntiles, 128>>>(done, launched, lat, spin);
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        static int h[ntiles]; int hl; long long hlat[2];
        cudaMemcpy(h, done, sizeof(h), cudaMemcpyDeviceToHost);
        cudaMemcpy(&hl, launched, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hlat, lat, 16, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < ntiles; ++i) bad += h[i] != 1;
        printf("spin %d: %s, %d tiles, %d not done exactly once, %d CTAs launched, CTA 0 did %lld tiles, %.0f cycles per try_cancel round trip (incl. spin), %.3f ms\n",
               spin, cudaGetErrorString(e), ntiles, bad, hl, hlat[1], hlat[1] ? (double)hlat[0] / hlat[1] : 0.0, ms);
    }
    return 0;
}
