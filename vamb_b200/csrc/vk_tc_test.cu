// Stand-alone check of the tcgen05 3xTF32 tile GEMM (vk_tc.cuh): C[M, N] = A * B^T for every
// combination of operand storage orders.  Used by tests/test_tc_gpu.py before the core is trusted
// inside the VAE kernels.
#include "vk_tc.cuh"

namespace {

struct Ld4Plain {
    const float *p;
    int ld, rows, cols, aligned;
    __device__ __forceinline__ float4 ld4(int r, int c4) const {
        const int c = c4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r >= rows || c >= cols) return v;
        const float *q = p + (int64_t)r * ld + c;
        if (aligned && c + 3 < cols) return __ldg(reinterpret_cast<const float4 *>(q));
        v.x = __ldg(q);
        if (c + 1 < cols) v.y = __ldg(q + 1);
        if (c + 2 < cols) v.z = __ldg(q + 2);
        if (c + 3 < cols) v.w = __ldg(q + 3);
        return v;
    }
};

// a_mn: A stored [K][M] (else [M][K]); b_mn: B stored [K][N] (else [N][K])
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(tc::TC_THREADS, 1)
tc_gemm_test_kernel(const float *A, const float *B, float *C, int M, int N, int K, int lda, int ldb, int variant) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ tc::TcShared sh;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
    int bn = N - n0;
    bn = bn > 128 ? 128 : ((bn + 15) & ~15);
    const int al_a = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const int al_b = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    Ld4Plain la{A, lda, A_MN ? K : M, A_MN ? M : K, al_a};
    Ld4Plain lb{B, ldb, B_MN ? K : N, B_MN ? N : K, al_b};
    tc::tc_tile_mainloop<A_MN, B_MN>(K, m0, n0, bn, la, lb, smem, &sh);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = m0 + (warp & 3) * 32 + lane;
    for (int c = (warp >> 2) * 64; c < (warp >> 2) * 64 + 64 && c < bn; c += 32) {
        float v[32];
        tc::tc_read_acc(&sh, c, v);
        if (m < M) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + c + j < N) C[(int64_t)m * N + n0 + c + j] = v[j];
        }
    }
    tc::tc_tile_end(&sh);
}

}  // namespace

static int g_variant = -1;
extern "C" void vk_tc_set_variant(int v) { g_variant = v; }

extern "C" int vk_tc_gemm_test(const float *A, const float *B, float *C, int M, int N, int K, int a_mn, int b_mn,
                               void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int lda = a_mn ? M : K, ldb = b_mn ? N : K;
    dim3 grid((N + 127) / 128, (M + 127) / 128);
    const int smem = tc::tc_smem_bytes(128);
#define LAUNCH(AM, BM_)                                                                                     \
    do {                                                                                                    \
        VK_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel<AM, BM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        tc_gemm_test_kernel<AM, BM_><<<grid, tc::TC_THREADS, smem, s>>>(A, B, C, M, N, K, lda, ldb, g_variant);        \
    } while (0)
    if (!a_mn && !b_mn) LAUNCH(false, false);
    else if (!a_mn && b_mn) LAUNCH(false, true);
    else if (a_mn && !b_mn) LAUNCH(true, false);
    else LAUNCH(true, true);
#undef LAUNCH
    VK_LAUNCH_CHECK();
    return 0;
}

// ---- launch / prologue overhead probes (tools/tc_fixed_cost.py) ----
namespace {
// mode 0: nothing; 1: + barrier init + TMEM alloc/dealloc; 2: + one MMA + commit + wait; 3: + __threadfence + atomic
__global__ void __launch_bounds__(tc::TC_THREADS, 1) tc_overhead_kernel(int mode, int *counter) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ tc::TcShared sh;
    if (mode == 0) return;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        tc::mbar_init(&sh.bar_done, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(&sh.tmem_base, 128);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (mode >= 2) {
        for (int i = tid; i < 16384; i += tc::TC_THREADS) reinterpret_cast<float *>(smem)[i] = 1.0f;
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            const uint64_t d = tc::make_smem_desc(tc::smem_u32(smem), 128, 1024);
            tc::umma_tf32(sh.tmem_base, d, d + (32768 >> 4), tc::make_idesc_tf32(128, 32, 0, 0), 0u);
            tc::umma_commit(&sh.bar_done);
        }
        tc::mbar_wait(&sh.bar_done, 0);
        tc::tc_fence_after();
    }
    if (mode == 3) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(counter, 1);
        }
    }
    if (mode == 4) {  // + accumulator read-back and a global store
        float v[32];
        tc::tc_read_acc(&sh, 0, v);
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) t += v[j];
        if (t == 123.456f) counter[1] = 1;
    }
    if (mode == 5) {  // sixteen dependent MMA+commit+wait round trips instead of one
        for (int it = 1; it < 16; ++it) {
            if (tid == 0) {
                const uint64_t d = tc::make_smem_desc(tc::smem_u32(smem), 128, 1024);
                tc::umma_tf32(sh.tmem_base, d, d + (32768 >> 4), tc::make_idesc_tf32(128, 32, 0, 0), 1u);
                tc::umma_commit(&sh.bar_done);
            }
            tc::mbar_wait(&sh.bar_done, (uint32_t)(it & 1));
            tc::tc_fence_after();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(sh.tmem_base, 128);
}
}  // namespace

extern "C" int vk_tc_overhead_test(int mode, int grid, int smem_bytes, int *counter, void *stream) {
    static bool attr = false;
    if (!attr) {
        VK_CUDA(cudaFuncSetAttribute(tc_overhead_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    tc_overhead_kernel<<<grid, tc::TC_THREADS, smem_bytes, (cudaStream_t)stream>>>(mode, counter);
    VK_LAUNCH_CHECK();
    return 0;
}

// ---- issue / execution rate of back-to-back tcgen05.mma (tools/tc_fixed_cost.py) ----
namespace {
__global__ void __launch_bounds__(tc::TC_THREADS, 1) tc_mma_rate_kernel(int n_mma, int n, int swizzle, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ tc::TcShared sh;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        tc::mbar_init(&sh.bar_done, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(&sh.tmem_base, 256);
    for (int i = tid; i < 16384; i += tc::TC_THREADS) reinterpret_cast<float *>(smem)[i] = 1.0f;
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (tid == 0) {
        uint64_t d = tc::make_smem_desc(tc::smem_u32(smem), swizzle ? 16 : 128, 1024);
        if (swizzle) d |= (uint64_t)2 << 61;  // SWIZZLE_128B
        const uint64_t da = d, db = d + (32768 >> 4);
        const uint32_t idesc = tc::make_idesc_tf32(128, n, 0, 0);
        const long long t0 = clock64();
        if (swizzle == 2)  // A operand from tensor memory (columns 128..), B from shared memory
            for (int i = 0; i < n_mma; ++i) tc::umma_tf32_ts(sh.tmem_base, sh.tmem_base + 128u + 8u * (i & 3), db, idesc, 1u);
        else
            for (int i = 0; i < n_mma; ++i) tc::umma_tf32(sh.tmem_base, da + (uint64_t)((i & 3) * (swizzle ? 2 : 16)), db, idesc, 1u);
        tc::umma_commit(&sh.bar_done);
        const long long t1 = clock64();
        tc::mbar_wait(&sh.bar_done, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(sh.tmem_base, 256);
}
}  // namespace

extern "C" int vk_tc_mma_rate_test(int n_mma, int n, int swizzle, long long *out_dev, void *stream) {
    VK_CUDA(cudaFuncSetAttribute(tc_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    tc_mma_rate_kernel<<<1, tc::TC_THREADS, 100 * 1024, (cudaStream_t)stream>>>(n_mma, n, swizzle, out_dev);
    VK_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t vk_lane_major_index(int r, int k, int ld) { return (int64_t)tc::lane_major_index(r, k, ld); }
