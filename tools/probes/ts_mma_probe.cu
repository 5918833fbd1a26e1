// ts_mma_probe.cu -- stand-alone probe (not part of the library): does tcgen05.mma take its A operand from TMEM the way the
// chain kernel's "activations stay in tensor memory" design assumes?
//   layout under test: row m of the 128-row tile = TMEM lane m; elements k = 2j, 2j+1 of the row = low / high half of 32-bit column j
//   (written by tcgen05.st.32x32b, thread = row).  A K=16 slice is 8 columns.
// Also measures the round-trip latency of tcgen05.ld/st used as scratch memory (for the single-CTA FPS design).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ts_mma_probe ts_mma_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mb_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
        "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tc_ld1(uint32_t taddr, uint32_t &v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
}
__device__ __forceinline__ void tc_st1(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr uint32_t kDescHi = 0x40004040u;  // SBO = 1024 B, version 1, SWIZZLE_128B (as gspn_b200/csrc/mlp_tc.cu)
__device__ __forceinline__ uint64_t desc64(uint32_t saddr) { return ((uint64_t)kDescHi << 32) | (((saddr & 0x3FFFFu) >> 4) | (1u << 16)); }
__device__ __forceinline__ uint32_t instr_desc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ size_t sw_off(int r, int k) {  // K-major SWIZZLE_128B, 64 bf16 per row
    return (size_t)(r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7))) << 4) + (k & 7) * 2;
}

// mode 0: A from shared memory (sanity of the probe itself); mode 1: A from TMEM
__global__ void __launch_bounds__(128) probe_kernel(const float *A, const float *B, float *D, int N, int mode, long long *lat) {
    extern __shared__ unsigned char raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const uint32_t base = (s_u32(raw) + 1023u) & ~1023u;
    unsigned char *sm = raw + (base - s_u32(raw));
    unsigned char *sA = sm, *sB = sm + 16384;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mb_init(s_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = tid; e < 128 * 64; e += 128) {
        int r = e >> 6, k = e & 63;
        *reinterpret_cast<__nv_bfloat16 *>(sA + sw_off(r, k)) = __float2bfloat16_rn(A[e]);
    }
    for (int e = tid; e < N * 64; e += 128) {
        int r = e >> 6, k = e & 63;
        *reinterpret_cast<__nv_bfloat16 *>(sB + sw_off(r, k)) = __float2bfloat16_rn(B[e]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t colA = 128;  // A operand columns [128, 160): 64 bf16 per row
    if (mode == 1) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(A[tid * 64 + 2 * j], A[tid * 64 + 2 * j + 1]);  // .x (low half) = element 2j
            v[j] = *reinterpret_cast<uint32_t *>(&h);
        }
        tc_st32(lane_base + colA, v);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc = instr_desc(128, N);
        for (int k = 0; k < 4; ++k) {
            if (mode == 0) mma_ss(tmem, desc64(base) + 2 * k, desc64(base + 16384) + 2 * k, idesc, k != 0);
            else mma_ts(tmem, tmem + colA + 8 * k, desc64(base + 16384) + 2 * k, idesc, k != 0);
        }
        tc_commit(s_u32(&bar));
    }
    mb_wait(s_u32(&bar), 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tc_ld32(lane_base + c0, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    // ---- scratch latency: dependent chain of st -> wait -> ld -> wait on a dynamic column
    if (lat) {
        uint32_t x = tid, col = 200;
        long long t0 = clock64();
        for (int i = 0; i < 64; ++i) {
            tc_st1(lane_base + col, x);
            tc_wait_st();
            tc_ld1(lane_base + col, x);
            tc_wait_ld();
            col = 200 + (x & 7);
            x += 1;
        }
        long long t1 = clock64();
        uint32_t y = x;
        for (int i = 0; i < 64; ++i) {
            tc_ld1(lane_base + 200 + (y & 7), y);
            tc_wait_ld();
        }
        long long t2 = clock64();
        // shared-memory comparison: dependent ld.shared chain
        volatile uint32_t *sp = reinterpret_cast<volatile uint32_t *>(sm + 40000);
        sp[tid] = tid;
        __syncwarp();
        uint32_t z = tid;
        long long t3 = clock64();
        for (int i = 0; i < 64; ++i) z = sp[z & 127];
        long long t4 = clock64();
        if (tid == 0) { lat[0] = (t1 - t0) / 64; lat[1] = (t2 - t1) / 64; lat[2] = (t4 - t3) / 64; lat[3] = x + y + z; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main() {
    const int M = 128, K = 64;
    int fails = 0;
    for (int N : {64, 128}) {
        std::vector<float> A(M * K), B(N * K), ref(M * N), refswap(M * N);
        srand(3);
        for (auto &x : A) x = (rand() % 2001 - 1000) / 500.f;
        for (auto &x : B) x = (rand() % 2001 - 1000) / 500.f;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double s = 0, s2 = 0;
                for (int k = 0; k < K; ++k) { s += (double)bf(A[m * K + k]) * bf(B[n * K + k]); s2 += (double)bf(A[m * K + (k ^ 1)]) * bf(B[n * K + k]); }
                ref[m * N + n] = (float)s; refswap[m * N + n] = (float)s2;
            }
        float *dA, *dB, *dD; long long *dl;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dl, 64));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        for (int mode = 0; mode < 2; ++mode) {
            CK(cudaMemset(dD, 0, M * N * 4));
            probe_kernel<<<1, 128, 64 * 1024>>>(dA, dB, dD, N, mode, mode == 1 ? dl : nullptr);
            CK(cudaDeviceSynchronize());
            std::vector<float> D(M * N);
            CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
            double e = 0, es = 0;
            for (int i = 0; i < M * N; ++i) { e = fmax(e, fabs(D[i] - ref[i])); es = fmax(es, fabs(D[i] - refswap[i])); }
            printf("N=%d mode=%s max|err| vs ref %.3e, vs half-swapped ref %.3e  -> %s\n", N, mode ? "A in TMEM" : "A in smem", e, es,
                   e < 1e-3 ? "OK" : (es < 1e-3 ? "HALVES SWAPPED" : "MISMATCH"));
            if (e >= 1e-3) ++fails;
            if (mode == 1) {
                long long l[4];
                CK(cudaMemcpy(l, dl, 32, cudaMemcpyDeviceToHost));
                printf("  scratch latency (cycles): tcgen05 st+wait+ld+wait %lld, ld+wait %lld, ld.shared dependent %lld\n", l[0], l[1], l[2]);
            }
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dl);
    }
    printf("ts_mma_probe: %s\n", fails ? "FAIL" : "PASS");
    return fails != 0;
}
