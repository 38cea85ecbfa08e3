// tcgen05 (5th-gen tensor core) building blocks for sm_100a: raw PTX wrappers, UMMA shared-memory /
// instruction descriptors, and a 128 x BN x K tile GEMM in error-compensated 3xTF32.
//
// Operands are materialised ONCE per layer by the producing kernel (vk_vae.cu: fused staging / prep kernels) as
// plain zero-padded fp32 arrays; the warp-specialised main loop below (ws_mainloop) keeps the 128-row operand in
// tensor memory and streams the narrow operand through a shared-memory ring.
//
// 3xTF32: kind::tf32 reads the top 19 bits of each fp32 operand word (low 13 mantissa bits ignored),
// so hi = x as stored, lo = x - (x & 0xFFFFE000) (exact).  D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo leaves
// a relative error of ~2^-21 per product, i.e. fp32-grade results from the tensor pipe.
#pragma once
#include "vk_common.cuh"

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// The bare try_wait loop (the hardware suspends the thread inside try_wait for a bounded time), plus a poll counter: a
// wait that never completes (a lost arrival, a bulk copy that never lands) traps after 2^26 failed polls (seconds)
// instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .u32 n;\n\t"
        "mov.u32 n, 0;\n\t"
        "LAB_WAIT%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LAB_DONE%=;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.lt.u32 q, n, 0x4000000;\n\t"
        "@q bra LAB_WAIT%=;\n\t"
        "trap;\n\t"
        "LAB_DONE%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---- fences ----
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- tensor memory ----
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp as alloc
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive columns -> 32 registers per thread (thread = lane of its warp's quadrant)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors (cute/arch/mma_sm100_desc.hpp: SmemDescriptor / InstrDescriptor) ----
// shared-memory matrix descriptor, no swizzle: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout_type=0 [61,64).  SBO = byte stride between core matrices along M/N,
// LBO = byte stride between core matrices along K.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor for kind::tf32, fp32 accumulate: c_format=F32 [4,6), a/b_format=TF32(2) [7,10)/[10,13),
// a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= 2u << 7;
    d |= 2u << 10;
    d |= (uint32_t)(a_mn_major & 1) << 15;
    d |= (uint32_t)(b_mn_major & 1) << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---- operand tiles in shared memory (fp32 words, no swizzle) ----
// Tile = ROWS (M or N extent, multiple of 8) x KT (=32) elements.
//  K-major  : core matrix = 8 rows x 16 B; offset(r, k) = (r/8)*1024 + (k/4)*128 + (r%8)*16 + (k%4)*4
//             -> SBO = 1024 (next 8 rows), LBO = 128 (next 16-byte K chunk); k8-step j starts at +256*j.
//  MN-major : core matrix = 8 k x 16 B (4 consecutive m); offset(m, k) = (k/8)*(ROWS*32) + (m/4)*128 + (k%8)*16 + (m%4)*4
//             -> SBO = 128 (next 4 m), LBO = ROWS*32 (next 8 k, unused by a K=8 instruction); k8-step j starts at +ROWS*32*j.
constexpr int KT = 32;
__host__ __device__ constexpr int b_tile_bytes(int bn) { return bn * KT * 4; }

__device__ __forceinline__ uint32_t off_kmajor(int r, int k4) {  // k4 = k / 4 (one float4 per call)
    return (uint32_t)((r >> 3) * 1024 + k4 * 128 + (r & 7) * 16);
}
__device__ __forceinline__ uint32_t off_mnmajor(int m4, int k, int rows) {  // m4 = m / 4
    return (uint32_t)((k >> 3) * (rows * 32) + m4 * 128 + (k & 7) * 16);
}

__device__ __forceinline__ float4 tf32_lo(const float4 v) {
    float4 r;
    r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    return r;
}

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 256;

struct TcShared {  // micro-benchmark kernels of vk_tc_test.cu (tools/tc_fixed_cost.py)
    uint64_t bar_stage[2];
    uint64_t bar_done;
    uint32_t tmem_base;
};

// Read the accumulator rows of this thread's TMEM quadrant: warp w owns lanes 32*(w%4)..+31; v receives 32
// consecutive columns starting at `col`.
__device__ __forceinline__ void tc_read_acc(const TcShared *sh, int col, float (&v)[32]) {
    const int warp = threadIdx.x >> 5;
    const uint32_t taddr = sh->tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col;
    tmem_ld32(taddr, v);
}

}  // namespace tc

// =====================================================================================================
// Layer GEMM main loop: operands are PLAIN K-major fp32 arrays in global memory, zero padded to 128 rows /
// 32 columns with 16-byte aligned rows -- produced once per layer by the prep kernels (vk_vae.cu).  No
// transform and no bounds checks on the way in; the tf32 remainders are derived in shared memory.
namespace tc {

struct OpRef {
    const float *hi, *lo;
    int ld;  // floats per row (multiple of 4)
};

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- TMA: bulk tensor copies global -> shared, completion counted in bytes on an mbarrier ----
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// box of the 2-D tensor map `tmap` at (c0 = innermost element, c1 = row) -> shared memory at smem_dst
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void *tmap, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// shared-memory matrix descriptor of a K-major tile written by TMA with CU_TENSOR_MAP_SWIZZLE_128B: rows of 128 bytes
// (one 32-float k-tile), 8-row swizzle atoms of 1024 bytes (SBO), layout_type = SWIZZLE_128B (2); the tile base must be
// 1024-byte aligned; a K = 8 (32-byte) step advances the start address by 32 bytes inside the atom.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;                  // LBO: unused for swizzled K-major operands
    d |= (uint64_t)(1024u >> 4) << 32;       // SBO: next group of 8 rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 consecutive columns, registers -> tensor memory
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem, 128 lanes x 8 columns] * B[smem descriptor]^T
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// "Lane-major" storage of an A-role operand (the 128-row side of the GEMM): 128-row panels of an
// [rows][ld] array, each k-tile of a panel is one 16 KB block  [k / 4 (8)][row (128)][k % 4 (4)]  so that
// the 32 lanes of a warp read / write consecutive float4.  Same footprint as the row-major array.
__host__ __device__ __forceinline__ size_t lane_major_index(int r, int k, int ld) {
    return (size_t)(r >> 7) * 128 * ld + (size_t)(k >> 5) * 4096 + (size_t)((k & 31) >> 2) * 512 + (size_t)(r & 127) * 4 + (k & 3);
}

// Warp-specialised pipeline.  A CTA is 8 producer/epilogue warps + 1 MMA warp (WS_THREADS = 288).
// Shared memory bandwidth (128 B/clk) is what bounds a 3xTF32 tile loop that keeps both operands in shared
// memory (each k-tile: copy in, read + write for the split, and twelve MMA reads of the 128-row operand), so
// the 128-row operand A never touches shared memory:
//   producers : warp w owns TMEM lanes 32 (w % 4).. and the k-half (w / 4) of each k-tile: it loads its 16
//               values per row from the lane-major array (coalesced float4), splits them into the tf32 value
//               ("hi" = the fp32 word, the MMA ignores the low mantissa bits) and the remainder ("lo") in
//               registers and writes both to the A stage in TENSOR memory (tcgen05.st); the narrow operand
//               B goes global -> shared with cp.async, each thread deriving "lo" of the chunk it copied;
//               then proxy/tcgen05 fences and one arrival per warp on full[stage];
//   MMA warp  : one lane waits full[stage], issues the 12 tcgen05.mma (A from tensor memory, B from shared
//               memory; small terms first) of the k-tile and commits them to empty[stage].
// Tensor memory: columns [0, 128) accumulator, then S stages of 64 columns (A_hi 32 | A_lo 32).
// A stage is refilled once the MMAs that read it have completed; two tiles are in flight.
constexpr int WS_THREADS = 288;
constexpr int WS_EPI_THREADS = 256;
constexpr int WS_STAGES = 4;
constexpr uint32_t WS_TMEM_COLS = 512;  // 128 + 4 * 64 rounded up to a power of two
__host__ __device__ constexpr int ws_smem_bytes(int bn) { return WS_STAGES * 2 * b_tile_bytes(bn) + 1024; }  // bn = slot rows

struct WsShared {
    uint64_t full[WS_STAGES], empty[WS_STAGES], done;
    uint64_t acc_full[2], acc_empty[2];  // FLUSH: hand-over of the two alternating accumulators
    uint32_t tmem_base;
};

// FLUSH (wgrad: the reduction runs over the batch, hundreds of rows): the tensor core's fp32 accumulator TRUNCATES on
// every accumulation, so the error of a long chain grows linearly with its length and is biased.  With FLUSH the chain
// is cut every WS_FLUSH_KT k-tiles (128 batch rows): the MMA warp alternates between two accumulators in tensor memory
// (columns [0, 128) and [384, 512)), and the epilogue warps add each finished group into fp32 REGISTERS (round to
// nearest) while the next group accumulates -- `racc[i][j]` = column 32 i + 16 (warp / 4) + j of this thread's row.
constexpr int WS_FLUSH_KT = 4;
constexpr int WS_DRAIN_LAG = 2;  // a group is drained once the producers are this many k-tiles into the next one
constexpr uint32_t WS_ACC1_COL = 384u;

// Stacked B: the tf32 remainder tile of B lies directly behind its hi tile in every ring stage, with the same layout, so
// ONE MMA with N' = 2 bn rows of B computes A_hi x B_hi (accumulator columns [0, bn)) and A_hi x B_lo (columns
// [bn, 2 bn)); a second one adds A_lo x B_hi into [0, bn); the epilogue adds the two column ranges.  A k-step is then 2
// tcgen05.mma instead of 3 -- and an MMA with M = 128 costs the issuing thread ~47 cycles for any N <= 64 (64 cycles at
// N = 128; tools/tc_fixed_cost.py, also with alternating accumulators: an issue cost, not a dependency), which is what
// bounds the main loop of the narrow tiles small batches use.  Needs 2 bn <= 128 accumulator columns, a single
// accumulator (no FLUSH) and, with TMA boxes of slot_rows rows, a full tile (bn == slot_rows).
constexpr bool WS_STACK_B = true;
template <bool TMA_B, bool FLUSH>
__device__ __forceinline__ bool ws_stacked(int bn, int slot_rows) {
    return WS_STACK_B && !FLUSH && 2 * bn <= 128 && (!TMA_B || bn == slot_rows);
}

// D[128 x bn] = sum over k-tiles [kt0, kt0 + nk) of A[m0.., k] * B[n0.., k]^T.  A: lane-major (m0 a multiple
// of 128), B: plain K-major; both fp32, zero padded to whole tiles (ld = floats per row, multiple of 32).
// Returns true for the 256 epilogue threads (accumulator complete), false for the MMA warp (which is done).
//
// TMA_B: the B operand (weights) is fetched by the TMA engine instead: one elected thread arms full[stage] with the
// byte count and issues two bulk tensor copies (the fp32 tile and its pre-split tf32 remainder, written once per step by
// prep_weights) through 128B-swizzled tensor maps whose box is 32 floats x `slot_rows` rows; the producer warps then
// only feed A.  `slot_rows` (= the launch's tile width, >= bn) sizes the ring slots; tm_hi / tm_lo point to CUtensorMap
// objects in kernel-parameter space (__grid_constant__).
template <bool TMA_B = false, bool FLUSH = false>
__device__ __forceinline__ bool ws_mainloop(const float *A, int lda, int m0, const float *B, int ldb, int n0, int bn,
                                            int kt0, int nk, uint8_t *smem, WsShared *sh, int slot_rows = 0,
                                            const void *tm_hi = nullptr, const void *tm_lo = nullptr,
                                            float (*racc)[16] = nullptr) {
    // PA: A tiles in flight as global loads into registers; PB: B tiles issued ahead into the shared-memory ring.
    // A k-tile of A is 16 KB per CTA read from L2 with ~0.8 us of latency: two tiles in flight held the main loop to
    // 0.43-0.47 us per k-tile (latency / 2) against the 0.29 us the 12 MMAs need (tools/kernel_timeline.py).
    constexpr int S = WS_STAGES, PA = 4, PB = 3;
    static_assert(PB < S && PA <= S, "ring depth");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bbytes = b_tile_bytes(TMA_B ? slot_rows : bn), sbytes = 2 * bbytes;
    const uint32_t smem_base = smem_u32(smem);
    // B chunk q = tid + 256 i of a tile: row ((q >> 6) << 3) | (q & 7) = r0 + 32 i, 16-byte column k4, at byte q * 16
    const int r0 = ((tid >> 6) << 3) | (tid & 7), k4 = (tid >> 3) & 7;
    const float *pb = B + (size_t)(n0 + r0) * ldb + (size_t)kt0 * KT + k4 * 4;
    const size_t b_step = (size_t)32 * ldb;
    const int nbi = bn >> 5;                                // whole 32-row groups of the B tile
    const bool b_tail = tid < (bn & 31) * 8;                // + 16 rows when bn is an odd multiple of 16
    const uint32_t soff = (uint32_t)tid * 16u;
    // A: row 32 (warp % 4) + lane of the panel, float4 groups 4 (warp / 4) .. + 3 of each k-tile
    const float4 *pa = reinterpret_cast<const float4 *>(A + (size_t)(m0 >> 7) * 128 * lda + (size_t)kt0 * 4096) +
                       (warp >> 2) * 4 * 128 + (warp & 3) * 32 + lane;
    auto issue_b = [&](int kt, int slot) {
        if (TMA_B) {
            if (tid == 0) {
                mbar_expect_tx(&sh->full[slot], (uint32_t)sbytes);
                tma_load_2d(smem_base + slot * sbytes, tm_hi, &sh->full[slot], (kt0 + kt) * KT, n0);
                tma_load_2d(smem_base + slot * sbytes + bbytes, tm_lo, &sh->full[slot], (kt0 + kt) * KT, n0);
            }
            return;
        }
        const uint32_t st = smem_base + slot * sbytes + soff;
        const float *gb = pb + (size_t)kt * KT;
        for (int i = 0; i < nbi; ++i) cp_async16(st + i * 4096, gb + i * b_step);
        if (b_tail) cp_async16(st + nbi * 4096, gb + nbi * b_step);
    };
    auto split_b = [&](int slot) {
        uint8_t *st = smem + slot * sbytes + soff;
        for (int i = 0; i < nbi; ++i)
            *reinterpret_cast<float4 *>(st + bbytes + i * 4096) = tf32_lo(*reinterpret_cast<const float4 *>(st + i * 4096));
        if (b_tail)
            *reinterpret_cast<float4 *>(st + bbytes + nbi * 4096) = tf32_lo(*reinterpret_cast<const float4 *>(st + nbi * 4096));
    };
    auto load_a = [&](int kt, float4 (&v)[4]) {
        const float4 *g = pa + (size_t)kt * 1024;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __ldg(g + j * 128);
    };
    // the first tiles are in flight while barriers / tensor memory are set up
    float4 ra[PA][4];
    if (warp < 8) {
#pragma unroll
        for (int t = 0; t < PB; ++t) {
            if (t < nk && !TMA_B) issue_b(t, t);  // (the bulk copies need the barriers: issued after the set-up below)
            if (!TMA_B) cp_async_commit();
        }
#pragma unroll
        for (int t = 0; t < PA; ++t)
            if (t < nk) load_a(t, ra[t]);
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            mbar_init(&sh->full[s], 8);
            mbar_init(&sh->empty[s], 1);
        }
        mbar_init(&sh->done, 1);
        if (FLUSH) {
            mbar_init(&sh->acc_full[0], 1);
            mbar_init(&sh->acc_full[1], 1);
            mbar_init(&sh->acc_empty[0], 8);
            mbar_init(&sh->acc_empty[1], 8);
        }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&sh->tmem_base, WS_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tl_mark(8);
    if (TMA_B && warp < 8) {
#pragma unroll
        for (int t = 0; t < PB; ++t)
            if (t < nk) issue_b(t, t);
    }
    const uint32_t tmem_d = sh->tmem_base;
    if (warp == 8) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BM, bn, 0, 0);
            const bool stk = ws_stacked<TMA_B, FLUSH>(bn, slot_rows);
            const uint32_t idesc2 = make_idesc_tf32(TC_BM, 2 * bn, 0, 0);
            const uint64_t desc0 = TMA_B ? make_smem_desc_sw128(smem_base) : make_smem_desc(smem_base, 128, 1024);
            const uint32_t kstep = TMA_B ? 2u : 16u;  // start-address units (16 bytes) per K = 8 step
            int slot = 0;
            uint32_t phase = 0;
            for (int kt = 0; kt < nk; ++kt) {
                const int grp = FLUSH ? kt / WS_FLUSH_KT : 0;
                const bool grp_first = FLUSH ? (kt % WS_FLUSH_KT == 0) : (kt == 0);
                const bool grp_last = FLUSH ? ((kt + 1) % WS_FLUSH_KT == 0 || kt == nk - 1) : (kt == nk - 1);
                const uint32_t tmem_acc = tmem_d + ((FLUSH && (grp & 1)) ? WS_ACC1_COL : 0u);
                // the epilogue warps must have drained this accumulator's previous group (two groups ago)
                if (FLUSH && grp_first && grp >= 2) mbar_wait(&sh->acc_empty[grp & 1], (uint32_t)(((grp >> 1) - 1) & 1));
                mbar_wait(&sh->full[slot], phase);
                tc_fence_after();
                const uint32_t ah = tmem_d + 128u + 64u * slot, al = ah + 32u;
                // descriptors differ only in the 14-bit start-address field: add (byte offset >> 4)
                const uint64_t bh = desc0 + (uint64_t)((uint32_t)(slot * sbytes) >> 4);
                const uint64_t bl = bh + (uint64_t)((uint32_t)bbytes >> 4);
                if (stk) {
#pragma unroll
                    for (int j = 0; j < KT / 8; ++j) {
                        umma_tf32_ts(tmem_acc, ah + 8u * j, bh + kstep * j, idesc2, (grp_first && j == 0) ? 0u : 1u);  // hi x [hi; lo]
                        umma_tf32_ts(tmem_acc, al + 8u * j, bh + kstep * j, idesc, 1u);                                // lo x hi
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < KT / 8; ++j) {
                        umma_tf32_ts(tmem_acc, al + 8u * j, bh + kstep * j, idesc, (grp_first && j == 0) ? 0u : 1u);  // small terms first
                        umma_tf32_ts(tmem_acc, ah + 8u * j, bl + kstep * j, idesc, 1u);
                        umma_tf32_ts(tmem_acc, ah + 8u * j, bh + kstep * j, idesc, 1u);
                    }
                }
                umma_commit(&sh->empty[slot]);
                if (FLUSH && grp_last && kt != nk - 1) umma_commit(&sh->acc_full[grp & 1]);
                if (kt == nk - 1) umma_commit(&sh->done);
                if (++slot == S) {
                    slot = 0;
                    phase ^= 1u;
                }
            }
        }
        return false;
    }
    const uint32_t a_lane = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + 128u + (uint32_t)(warp >> 2) * 16u;
    auto produce = [&](int kt, float4(&v)[4]) {
        const int slot = kt % S;
        // A tile kt: registers -> tensor memory.  Its stage is free: every producer thread waited for the MMAs of tile
        // kt - S before B of tile kt was issued (PB tiles ago, below).
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 l4 = tf32_lo(v[j]);
            hi[4 * j] = v[j].x; hi[4 * j + 1] = v[j].y; hi[4 * j + 2] = v[j].z; hi[4 * j + 3] = v[j].w;
            lo[4 * j] = l4.x; lo[4 * j + 1] = l4.y; lo[4 * j + 2] = l4.z; lo[4 * j + 3] = l4.w;
        }
        tc_fence_after();
        tmem_st16(a_lane + 64u * slot, hi);
        tmem_st16(a_lane + 64u * slot + 32u, lo);
        if (kt + PA < nk) load_a(kt + PA, v);  // v's registers are free again
        if (!TMA_B) {
            cp_async_wait<PB - 1>();  // this thread's B copies of tile kt have landed (one group per produced tile)
            split_b(slot);
        }
        tmem_wait_st();
        if (!TMA_B) fence_async_smem();  // shared-memory writes visible to the tensor core (async proxy)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->full[slot]);
        // B of tile kt + PB, AFTER this tile has been handed over: its stage was read by the MMAs of tile
        // kt + PB - S (the previous tile for PB = S - 1), which run while the A part above is produced
        const int nt = kt + PB;
        if (nt < nk) {
            if (nt >= S) mbar_wait(&sh->empty[nt % S], (uint32_t)(((nt / S) - 1) & 1));
            issue_b(nt, nt % S);
        }
        if (!TMA_B) cp_async_commit();
        if (kt < 20) tl_mark(10 + kt);
    };
    // FLUSH: add a finished group (all but the last, which the epilogue reads) into the register accumulators
    int next_drain = 0;
    auto drain_ready = [&](int kt_done) {  // after producing tile kt_done
        if (!FLUSH) return;
        while ((next_drain + 1) * WS_FLUSH_KT + WS_DRAIN_LAG <= kt_done + 1 && (next_drain + 1) * WS_FLUSH_KT < nk) {
            const int a = next_drain & 1;
            mbar_wait(&sh->acc_full[a], (uint32_t)((next_drain >> 1) & 1));
            tc_fence_after();
            const uint32_t base = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (a ? WS_ACC1_COL : 0u);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = (warp >> 2) * 16 + 32 * i;
                if (c < bn) {
                    float v[16];
                    tmem_ld16(base + (uint32_t)c, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) racc[i][j] += v[j];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh->acc_empty[a]);
            ++next_drain;
        }
    };
    if (FLUSH) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) racc[i][j] = 0.0f;
    }
    for (int kt = 0; kt < nk; kt += PA) {
#pragma unroll
        for (int i = 0; i < PA; ++i)
            if (kt + i < nk) {
                produce(kt + i, ra[i]);
                drain_ready(kt + i);
            }
    }
    if (FLUSH) {  // groups that completed inside the last WS_DRAIN_LAG tiles
        while ((next_drain + 1) * WS_FLUSH_KT < nk) drain_ready(1 << 28);
    }
    tl_mark(9);
    if (nk > 0) mbar_wait(&sh->done, 0);
    tc_fence_after();
    return true;
}

// Accumulator -> shared tile [128][ts] (raw values): warp w owns TMEM lanes 32*(w%4)..+31, the two warp
// groups take alternate 16-column chunks.  nk == 0 (an empty split) yields zeros.
// racc != nullptr (FLUSH): the last group sits in accumulator ((nk - 1) / WS_FLUSH_KT) & 1 and the earlier groups in
// the register sums of ws_mainloop.
__device__ __forceinline__ void ws_acc_to_tile(const WsShared *sh, int bn, int nk, float *tile, int ts,
                                               const float (*racc)[16] = nullptr, bool stacked = false) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t acc_col = (racc && nk > 0 && (((nk - 1) / WS_FLUSH_KT) & 1)) ? WS_ACC1_COL : 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = (warp >> 2) * 16 + 32 * i;
        if (c >= bn) break;
        float v[16];
        if (nk > 0) tmem_ld16(sh->tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc_col + (uint32_t)c, v);
        else
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.0f;
        if (racc) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += racc[i][j];
        }
        if (stacked && nk > 0) {  // + A_hi x B_lo, accumulated in the columns behind the tile's own (ws_stacked)
            float w[16];
            tmem_ld16(sh->tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(bn + c), w);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += w[j];
        }
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
            *reinterpret_cast<float4 *>(tile + row * ts + c + 4 * j4) =
                make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
    }
}

// all epilogue threads: tile complete for everyone, tensor memory released
__device__ __forceinline__ void ws_tile_end(WsShared *sh) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc(sh->tmem_base, WS_TMEM_COLS);
}

}  // namespace tc
