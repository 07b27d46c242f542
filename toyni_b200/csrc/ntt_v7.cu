// Host side of the TMA-staged two-pass transform (ntt_pass_v7.cuh): tensor maps, launch configuration.
#include "ntt_v7.cuh"

#include "lde_expand.cuh"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

namespace bb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        // resolved through the runtime, so the library keeps linking against libcudart only (build.rs links cudart)
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

static int g_l2promo = -1;

// 5-D view of `batch` row-major [4096][ncols] u32 matrices: {column, d0, d1, d2, batch} with row d = 256 d2 + 16 d1 + d0;
// one box = C columns x 16 d0 x one d1 x 16 d2 = a 256-row chunk in the order round 1 reads it
static int make_map(const uint32_t* in, size_t ncols, size_t batch, size_t batch_stride, int cols, int box_d2, CUtensorMap* out) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    if (g_l2promo < 0) {
        const char* e = getenv("TOYNI_V7_L2PROMO");
        g_l2promo = e ? atoi(e) : 2;
    }
    const cuuint64_t row = (cuuint64_t)ncols * 4u;
    cuuint64_t dims[5] = {(cuuint64_t)ncols, 16, 16, 16, (cuuint64_t)batch};
    cuuint64_t strides[4] = {row, 16 * row, 256 * row, batch > 1 ? (cuuint64_t)batch_stride * 4u : (cuuint64_t)V7_R * row};
    cuuint32_t box[5] = {(cuuint32_t)cols, 16, 1, (cuuint32_t)box_d2, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUtensorMapL2promotion promo = g_l2promo == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                   : g_l2promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                   : g_l2promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                    : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    CUresult rc = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, const_cast<uint32_t*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

struct MapKey {
    int dev, cols, box_d2;
    const void* in;
    size_t ncols, batch, stride;
    bool operator<(const MapKey& o) const {
        return std::tie(dev, cols, box_d2, in, ncols, batch, stride) < std::tie(o.dev, o.cols, o.box_d2, o.in, o.ncols, o.batch, o.stride);
    }
};

// Descriptors live in device memory, one immutable 128-byte entry per distinct (buffer, shape): the kernel gets a pointer.
// (A descriptor passed by value in the parameter space can share its address with the different descriptor of the
// kernel launched just before it; entries here are written once, before their first use, and never change.)
static int map_get(const uint32_t* in, size_t ncols, size_t batch, size_t stride, int cols, int box_d2, const CUtensorMap** out) {
    static std::mutex mu;
    static std::map<MapKey, CUtensorMap*> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    const MapKey key{dev, cols, box_d2, in, ncols, batch, stride};
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it == cache.end()) {
        if (cache.size() >= 4096) {  // a process that walks through thousands of distinct buffers: start over (entries are 128 B each)
            cudaDeviceSynchronize();
            for (auto& kv : cache) cudaFree(kv.second);
            cache.clear();
        }
        alignas(64) CUtensorMap m;
        int rc = make_map(in, ncols, batch, stride, cols, box_d2, &m);
        if (rc) return rc;
        CUtensorMap* d = nullptr;
        cudaError_t e = cudaMalloc(&d, sizeof(CUtensorMap));
        if (e == cudaSuccess) e = cudaMemcpy(d, &m, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return (int)e;
        it = cache.emplace(key, d).first;
    }
    *out = it->second;
    return 0;
}

extern int g_pdl;

template <bool PASS2, int C>
static int launch_one(const CUtensorMap* map, const CUtensorMap* omap, V7Params p, size_t ncols, size_t batch, bool pdl, cudaStream_t s) {
    using T = V7<C>;
    static int n_ctas[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (n_ctas[dev] == 0) {
        cudaError_t e = cudaFuncSetAttribute(ntt_pass_v7_kernel<PASS2, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM);
        if (e != cudaSuccess) return (int)e;
        int n_sm = 0, occ = 0;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ntt_pass_v7_kernel<PASS2, C>, T::NT, T::SMEM);
        if (occ < 1) return (int)cudaErrorLaunchOutOfResources;
        n_ctas[dev] = n_sm * occ;
    }
    p.tiles_x = (uint32_t)(ncols / C);
    p.total_tiles = (uint32_t)(p.tiles_x * batch);
    uint32_t ctas = (uint32_t)n_ctas[dev];
    if (ctas > p.total_tiles) ctas = p.total_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(T::NT);
    cfg.dynamicSmemBytes = T::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (g_pdl && pdl) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, ntt_pass_v7_kernel<PASS2, C>, map, omap, p);
}

int launch_pass_v7(bool pass2, const uint32_t* in, size_t ncols, size_t batch, size_t in_batch_stride, V7Params p, bool pdl, cudaStream_t s) {
    // 8 columns (32-byte rows, one CTA of 16 warps per SM).  The kernel template also builds with 4 columns (two CTAs of
    // 8 warps per SM); that shape measured 152 us against 116 us — rows below a 32-byte sector — and is not instantiated.
    constexpr int cols = 8;
    static int flags = -1;
    if (flags < 0) {
        const char* f = getenv("TOYNI_V7_FLAGS");
        flags = f ? atoi(f) : 0;
    }
    p.flags = (uint32_t)flags;
    const CUtensorMap* map = nullptr;
    int rc = map_get(in, ncols, batch, in_batch_stride, cols, 16, &map);
    if (rc) return rc;
    // last pass of the square plan (out[e * ncols + col]): the result rows leave through the TMA as well
    const CUtensorMap* omap = nullptr;
    static int tma_store = -1;
    if (tma_store < 0) {
        const char* e = getenv("TOYNI_V7_TMA_STORE");
        tma_store = e ? atoi(e) : 1;
    }
    if (pass2 && cols == 8 && tma_store && ((size_t)1 << p.log_pfull) == ncols && (batch == 1 || p.out_batch_stride % 4 == 0)) {
        rc = map_get(p.out, ncols, batch, (size_t)p.out_batch_stride, cols, 8, &omap);
        if (rc) return rc;
    }
    return pass2 ? launch_one<true, 8>(map, omap, p, ncols, batch, pdl, s) : launch_one<false, 8>(map, omap, p, ncols, batch, pdl, s);
}

int launch_lde_expand(LdeParams p, bool pdl, cudaStream_t s) {
    static int n_ctas[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (n_ctas[dev] == 0) {
        cudaError_t e = cudaFuncSetAttribute(lde_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LDE::SMEM);
        if (e != cudaSuccess) return (int)e;
        int n_sm = 0;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        n_ctas[dev] = n_sm;  // one CTA of 16 warps and 180 KB per SM
    }
    p.total_tiles = 1u << (LDE::LOG_COLS - 2);
    uint32_t ctas = (uint32_t)n_ctas[dev];
    if (ctas > p.total_tiles) ctas = p.total_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(LDE::NT);
    cfg.dynamicSmemBytes = LDE::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (g_pdl && pdl) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, lde_expand_kernel, p);
}

}  // namespace bb
