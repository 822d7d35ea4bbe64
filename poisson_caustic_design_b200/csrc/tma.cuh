// TMA plumbing shared by the kernels that stage tiles with bulk tensor copies (sm_100a): mbarrier + cp.async.bulk.tensor.2d
// device helpers, and the host-side encoding of a 2-D fp64 tensor map through the driver entry point (no -lcuda).
#pragma once

#include <cuda.h>

#include "pcd_internal.h"

namespace pcd {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, int n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect(unsigned long long *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long *b, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a copy that never lands must not hang the device (the error word voids the launch, like a dead neighbour)
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity, int *err) {
    if (mbar_try(b, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(b, parity)) {
        if (clock64() - t0 > 4000000000ll) {
            if (err) atomicExch(err, 1);
            break;
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *b) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(b)), "r"(c0), "r"(c1)
                 : "memory");
}

// 2-D row-major fp64 tensor: dim0 (innermost, contiguous) x dim1, row pitch `stride1_bytes` (a multiple of 16), box
// box0 x box1 (each <= 256; box0 * 8 a multiple of 16); out-of-range parts of a box are zero-filled.  `map_out`: 128 bytes.
// Returns PCD_ERR_UNSUPPORTED when the driver entry point is missing or refuses the shape (callers keep a cp.async path).
inline int tma_encode_2d_f64(void *map_out, const void *base, unsigned long long dim0, unsigned long long dim1,
                             unsigned long long stride1_bytes, unsigned box0, unsigned box1) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const EncodeFn enc = []() -> EncodeFn {   // resolved once (thread-safe static initialisation)
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            return (EncodeFn)fn;
        cudaGetLastError();
        return nullptr;
    }();
    if (!enc || (stride1_bytes & 15ull) || ((unsigned long long)base & 15ull) || box0 > 256 || box1 > 256 || ((box0 * 8u) & 15u))
        return PCD_ERR_UNSUPPORTED;
    CUtensorMap m;
    const cuuint64_t dims[2] = {dim0, dim1}, strides[1] = {stride1_bytes};
    const cuuint32_t box[2] = {box0, box1}, estr[2] = {1, 1};
    const CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return PCD_ERR_UNSUPPORTED;
    memcpy(map_out, &m, sizeof(m));
    return PCD_OK;
}

}  // namespace pcd
