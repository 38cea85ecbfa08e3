// Host side of the TMA-fed B operand: CUtensorMap objects for the K-major weight arrays (vk_tc.cuh: ws_mainloop<true>).
// The driver's encoder is fetched at run time (cudaGetDriverEntryPoint), so the library has no link-time dependency on
// libcuda and still loads on a machine without a driver (the CPU test-suite checks the export list there).
#include <cuda.h>

#include <map>
#include <tuple>

#include "vk_common.cuh"

typedef CUresult (*vk_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                       const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static vk_encode_tiled_fn encode_fn() {
    static vk_encode_tiled_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (vk_encode_tiled_fn)p;
    }();
    return fn;
}

// fp32 array [rows][ld] (ld floats per row, a multiple of 32; base 16-byte aligned); box = 32 floats (one k-tile,
// 128 bytes = the swizzle span) x box_rows rows; CU_TENSOR_MAP_SWIZZLE_128B; out-of-bounds elements read as zero.
// The maps are cached per (base, ld, rows, box_rows): a map only encodes geometry, so it stays valid for the lifetime
// of the allocation.  out128 receives a copy (the kernels take them in parameter space).
int vk_make_tmap_2d(void *out128, const float *base, int ld, int rows, int box_rows) {
    typedef std::tuple<const float *, int, int, int> Key;
    static std::map<Key, CUtensorMap> cache;
    const Key key(base, ld, rows, box_rows);
    auto it = cache.find(key);
    if (it == cache.end()) {
        vk_encode_tiled_fn fn = encode_fn();
        if (!fn) {
            vk_set_error("vk_make_tmap_2d: cuTensorMapEncodeTiled is not available from this driver");
            return 1;
        }
        if ((ld & 31) || box_rows < 8 || box_rows > 256 || (reinterpret_cast<uintptr_t>(base) & 15)) {
            vk_set_error("vk_make_tmap_2d: bad geometry (ld=%d, box_rows=%d)", ld, box_rows);
            return 1;
        }
        CUtensorMap m;
        const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
        const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
        const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
        const cuuint32_t estride[2] = {1u, 1u};
        const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            vk_set_error("cuTensorMapEncodeTiled failed (%d) for [%d x %d], box 32 x %d", (int)r, rows, ld, box_rows);
            return 1;
        }
        it = cache.emplace(key, m).first;
    }
    memcpy(out128, &it->second, sizeof(CUtensorMap));
    return 0;
}
