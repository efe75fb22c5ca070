// stands in for <cuda.h> (driver API types used to build TMA tensor maps) when tc_gemm.cu is compiled for the host
#pragma once
#include "../cuda_emu.h"

typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 = 9 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_128B = 3 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_256B = 3 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };

// what cuTensorMapEncodeTiled records for a 2-D tiled map (the real one is an opaque 128-byte blob)
struct alignas(64) CUtensorMap {
    char* base;
    uint64_t cols, rows, row_pitch;     // elements, elements, bytes
    uint32_t box_cols, box_rows, esize;
    int swizzle;
    char pad[128 - 8 - 24 - 12 - 4];
};
static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is a 128-byte object");

enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
enum { cudaEnableDefault = 0 };
CUresult svae_emu_encode_tiled(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline cudaError_t cudaGetDriverEntryPoint(const char*, void** fn, int, cudaDriverEntryPointQueryResult* q) {
    *fn = (void*)&svae_emu_encode_tiled;
    *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
}

// cudaLaunchKernelEx with a cluster-dimension attribute
enum { cudaLaunchAttributeClusterDimension = 4 };
struct cudaLaunchAttribute {
    int id;
    struct { struct { unsigned x, y, z; } clusterDim; } val;
};
struct cudaLaunchConfig_t {
    dim3 gridDim, blockDim;
    size_t dynamicSmemBytes;
    cudaStream_t stream;
    cudaLaunchAttribute* attrs;
    unsigned numAttrs;
};
template <typename... P, typename... A>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*k)(P...), A&&... a) {
    unsigned cluster = 1;
    for (unsigned i = 0; i < cfg->numAttrs; ++i)
        if (cfg->attrs[i].id == cudaLaunchAttributeClusterDimension) cluster = cfg->attrs[i].val.clusterDim.x;
    svae_emu::Launcher(cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, cfg->stream, cluster).run(k, a...);
    return cudaSuccess;
}
#define __grid_constant__
