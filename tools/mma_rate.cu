// Microbenchmark (development aid, not part of the library): cycles per tcgen05.mma for the operand layout the
// engine uses (canonical K-major, no swizzle; A [128 x 16] from shared memory ".ss" or tensor memory ".ts"),
// kind::f16 and kind::f8f6f4, N = 32..192.   Build + run:  tools/mma_rate.sh   (needs a B200)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../catfish_b200/csrc/tc_ptx.cuh"

using namespace cf::ptx;

__device__ __forceinline__ void umma_f16_ss_inline_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(int n, int mode, int kinds, int reps, int commit_each, int nacc, int variant, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* a_s = smem;                       // 8 chunks x {4 KB main, 4 KB second plane}
    uint8_t* w_s = smem + 65536;               // {main, second plane} x [16 kgroups][n][16 B]
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (65536 + 2 * 128 * n * 2) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t tmem = tmem_base;
    if (warp == 0) {
        const uint32_t elected = elect_one();
        const uint32_t a_u = smem_u32(a_s), w_u = smem_u32(w_s);
        const uint32_t id16 = make_idesc_f16(128, n), id8 = make_idesc_e5m2(128, n);
        const long long t0 = clock64();
        if (variant == 0) {
        for (int r = 0; r < reps; ++r) {
            for (int kk = 0; kk < 8; ++kk) {
                const uint32_t tmem = tmem_base + (uint32_t)((kk % nacc) * (n <= 64 ? 64 : n));     // round-robin over independent accumulators
                const uint32_t a0 = a_u + kk * 8192;
                const uint32_t wx = w_u + kk * 2 * (n * 16);
                if (mode == 0) {
                    if (kinds & 1) umma_bf16_pred(tmem, make_smem_desc(a0, 2048, 128), make_smem_desc(wx, n * 16, 128), id16, 1, elected);
                    if (kinds & 2) umma_f8_pred(tmem, make_smem_desc(a0 + 4096, 2048, 128), make_smem_desc(wx + 128 * n * 2, n * 16, 128), id8, 1, elected);
                } else {
                    if (kinds & 1) umma_bf16_ts_pred(tmem, tmem_base + 448 + (kk & 3) * 8, make_smem_desc(wx, n * 16, 128), id16, 1, elected);
                    if (kinds & 2) umma_f8_ts_pred(tmem, tmem_base + 448 + 32 + (kk & 3) * 8, make_smem_desc(wx + 128 * n * 2, n * 16, 128), id8, 1, elected);
                }
                if (commit_each) umma_commit_pred(&bar[1], elected);
            }
        }
        } else if (variant == 1) {
            if (elected) {
                for (int r = 0; r < reps; ++r)
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_bf16(tmem, make_smem_desc(a_u + kk * 8192, 2048, 128), make_smem_desc(w_u + kk * 2 * (n * 16), n * 16, 128), id16, 1);
            }
            __syncwarp();
        } else if (variant == 2) {
            if (threadIdx.x == 0) {
                for (int r = 0; r < reps; ++r)
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_bf16(tmem, make_smem_desc(a_u + kk * 8192, 2048, 128), make_smem_desc(w_u + kk * 2 * (n * 16), n * 16, 128), id16, 1);
            }
            __syncwarp();
        } else if (variant == 3) {
            for (int r = 0; r < reps; ++r)
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_f16_ss_inline_elect(tmem, make_smem_desc(a_u + kk * 8192, 2048, 128), make_smem_desc(w_u + kk * 2 * (n * 16), n * 16, 128), id16, 1);
        } else {
            // descriptors precomputed outside the loop, predicated issue
            uint64_t ad[8], bd[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) { ad[kk] = make_smem_desc(a_u + kk * 8192, 2048, 128); bd[kk] = make_smem_desc(w_u + kk * 2 * (n * 16), n * 16, 128); }
            for (int r = 0; r < reps; ++r)
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) umma_bf16_pred(tmem, ad[kk], bd[kk], id16, 1, elected);
        }
        const long long t1 = clock64();
        umma_commit_pred(&bar[0], elected);
        mbar_wait(&bar[0], 0);
        const long long t2 = clock64();
        if (threadIdx.x == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = t2 - t0; }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * 2 * sizeof(long long));
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 2 * 128 * 192 * 2);
    const int reps = 64;
    printf("%4s %8s | %10s %10s\n", "N", "variant", "issue/MMA", "total/MMA");
    const char* names[5] = {"pred-reg", "if-elect", "if-tid0", "asm-elect", "pre-desc"};
    for (int n : {192, 64, 32})
        for (int variant : {0, 1, 2, 3, 4}) {
            const int grid = 148;
            const size_t smem = 65536 + 2 * 128 * n * 2;
            mma_rate_kernel<<<grid, 128, smem>>>(n, 0, 1, reps, 0, 1, variant, out);
            mma_rate_kernel<<<grid, 128, smem>>>(n, 0, 1, reps, 0, 1, variant, out);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            long long h[2 * 148];
            cudaMemcpy(h, out, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
            double issue = 0, total = 0;
            for (int b = 0; b < grid; ++b) { issue += h[2 * b]; total += h[2 * b + 1]; }
            const double mmas = reps * 8.0;
            printf("%4d %8s | %10.1f %10.1f\n", n, names[variant], issue / grid / mmas, total / grid / mmas);
        }
    return 0;
}
