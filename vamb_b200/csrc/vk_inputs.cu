// The step before the hot path: the TNF projection of vamb/parsecontigs.py:141-150
// (Composition._project): per contig, 256 four-mer counts -> frequencies (row / rowsum, zero rows stay zero) shifted by
// -1/256, times the 256 x 103 projection kernel.  One pass over the counts: HBM-bound on paper (1,024 B in + 412 B out
// per contig), here limited by fp32 FMA issue (26,368 FMA per contig) -- a one-off of ~1.5 ms per million contigs.
#include "vk_common.cuh"

namespace {
constexpr int TP_ROWS = 32, TP_K = 256, TP_THREADS = 256, TP_NOUT_MAX = 104;

// thread = (row r = tid / 8, output group jg = tid % 8): outputs j = jg + 8 i, i < 13; the row's 256 normalised
// frequencies sit in shared memory and are broadcast to the 8 threads of the row; the kernel matrix streams through L1.
__global__ void __launch_bounds__(TP_THREADS)
tnf_project_kernel(const float *__restrict__ counts, const float *__restrict__ kernel, float *__restrict__ out, int64_t n,
                   int n_out) {
    __shared__ float xs[TP_ROWS][TP_K + 1];
    __shared__ float inv_sum[TP_ROWS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * TP_ROWS;
    // load 32 rows (coalesced), row sums in fp32 pairwise-free order: lanes then butterfly (documented tolerance 2e-6)
    for (int r = warp; r < TP_ROWS; r += TP_THREADS / 32) {
        const int64_t row = row0 + r;
        float s = 0.0f;
        for (int k = lane; k < TP_K; k += 32) {
            const float v = row < n ? __ldg(counts + row * TP_K + k) : 0.0f;
            xs[r][k] = v;
            s += v;
        }
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) inv_sum[r] = 1.0f / (s == 0.0f ? 1.0f : s);  // parsecontigs.py:144-146
    }
    __syncthreads();
    for (int i = tid; i < TP_ROWS * TP_K; i += TP_THREADS) {
        const int r = i / TP_K, k = i - r * TP_K;
        xs[r][k] = __fmaf_rn(xs[r][k], inv_sum[r], -(1.0f / 256.0f));  // fourmers *= 1/s; fourmers += -(1/256)
    }
    __syncthreads();
    const int r = tid >> 3, jg = tid & 7;
    float acc[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) acc[i] = 0.0f;
    for (int k = 0; k < TP_K; ++k) {
        const float x = xs[r][k];
        const float *kr = kernel + (int64_t)k * n_out;
#pragma unroll
        for (int i = 0; i < 13; ++i) {
            const int j = jg + 8 * i;
            if (j < n_out) acc[i] = __fmaf_rn(x, __ldg(kr + j), acc[i]);
        }
    }
    const int64_t row = row0 + r;
    if (row < n) {
#pragma unroll
        for (int i = 0; i < 13; ++i) {
            const int j = jg + 8 * i;
            if (j < n_out) out[row * n_out + j] = acc[i];
        }
    }
}
}  // namespace

extern "C" int vk_tnf_project(const float *counts, const float *kernel, float *out, int64_t n, int n_out, void *stream) {
    if (n < 0 || n_out < 1 || n_out > TP_NOUT_MAX) {
        vk_set_error("vk_tnf_project: n_out=%d outside [1, %d]", n_out, TP_NOUT_MAX);
        return 1;
    }
    if (n == 0) return 0;
    const int64_t blocks = (n + TP_ROWS - 1) / TP_ROWS;
    tnf_project_kernel<<<(unsigned)blocks, TP_THREADS, 0, (cudaStream_t)stream>>>(counts, kernel, out, n, n_out);
    VK_LAUNCH_CHECK();
    return 0;
}
