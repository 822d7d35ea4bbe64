// Device helpers shared by the stage kernels.
#pragma once

#include "pcd_internal.h"

namespace pcd {

constexpr int OWNER_NONE = 0x7fffffff;
constexpr int RED_BLOCKS = 296;  // 2 x 148 SMs; fixed so that reductions are reproducible run to run
constexpr int RED_THREADS = 256;

// Barycentric coordinates exactly as src/bvh.cpp:162-190 is called from :215
// (get_barycentric_coordinates(tri[2], tri[1], tri[0], p)): out[k] weights triangle vertex k.
__device__ __forceinline__ void barycentric(double t0x, double t0y, double t1x, double t1y, double t2x, double t2y,
                                            double px, double py, double &u, double &v, double &w) {
    // reference argument order: (t0,t1,t2) := (tri[2], tri[1], tri[0])
    const double ax = t2x, ay = t2y, bx = t1x, by = t1y, cx = t0x, cy = t0y;
    const double v0x = cx - ax, v0y = cy - ay, v1x = bx - ax, v1y = by - ay, v2x = px - ax, v2y = py - ay;
    const double dot00 = v0x * v0x + v0y * v0y, dot01 = v0x * v1x + v0y * v1y, dot02 = v0x * v2x + v0y * v2y;
    const double dot11 = v1x * v1x + v1y * v1y, dot12 = v1x * v2x + v1y * v2y;
    const double denom = dot00 * dot11 - dot01 * dot01;
    const double inv_denom = 1 / denom;
    u = (dot11 * dot02 - dot01 * dot12) * inv_denom;
    v = (dot00 * dot12 - dot01 * dot02) * inv_denom;
    w = 1.0 - u - v;
}

// inside test of src/bvh.cpp:208-217
__device__ __forceinline__ bool bary_inside(double u, double v) {
    const double eps = 1e-12;
    return (u >= -eps && v >= -eps) && ((u + v) <= 1.0 + eps);
}

// vertices of triangle t of the structured mesh (src/mesh.cpp:56-63)
__device__ __forceinline__ void tri_vertices(int t, int nx, int &a, int &b, int &c) {
    const int qd = t >> 1, qi = qd / (nx - 1), qj = qd - qi * (nx - 1);
    const int idx = qi * nx + qj;
    if (t & 1) { a = idx + nx; b = idx + 1; c = idx + nx + 1; }
    else       { a = idx;      b = idx + 1; c = idx + nx; }
}

// ---- reproducible block reductions ----------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double *scratch /* >= 32 doubles */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < (blockDim.x >> 5) ? scratch[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    return v;  // valid in thread 0
}

__device__ __forceinline__ double block_max(double v, double *scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double b = __shfl_down_sync(0xffffffffu, v, o);
        v = b > v ? b : v;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < (blockDim.x >> 5) ? scratch[lane] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double b = __shfl_down_sync(0xffffffffu, v, o);
            v = b > v ? b : v;
        }
    }
    return v;
}

// order-preserving map double -> uint64 (so that atomicMin/atomicMax give exact fp min/max)
__device__ __forceinline__ unsigned long long ordered_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double ordered_unkey(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double v;
#ifdef __CUDA_ARCH__
    v = __longlong_as_double((long long)b);
#else
    memcpy(&v, &b, sizeof(v));
#endif
    return v;
}

// two-stage reproducible sum of n doubles: partials[RED_BLOCKS] then out[0]
int reduce_sum(const double *in, long n, double *partials, double *out, cudaStream_t st);

}  // namespace pcd
