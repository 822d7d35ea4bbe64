// Microbenchmark (scratch tool): latency of CTA-to-CTA signalling through L2 on B200, the primitive under
// the resident SOR kernel's halo exchange.  nvcc -arch=sm_100a -O3 -o ll_latency tools/ll_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int MODE> __device__ __forceinline__ void st_flag(unsigned long long *p, unsigned long long v) {
    if (MODE == 0) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    else if (MODE == 1) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    else asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <int MODE> __device__ __forceinline__ unsigned long long ld_flag(const unsigned long long *p) {
    unsigned long long v;
    if (MODE == 0) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else if (MODE == 1) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ping-pong between CTA 0 and CTA `peer`, one thread each
template <int MODE>
__global__ void pingpong(unsigned long long *slots, int iters, int peer, long long *cycles) {
    if (threadIdx.x != 0) return;
    unsigned long long *mine = slots + 32 * blockIdx.x, *theirs;
    if (blockIdx.x == 0) {
        theirs = slots + 32 * peer;
        long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) {
            st_flag<MODE>(theirs, i);
            while (ld_flag<MODE>(mine) != (unsigned long long)i) { }
        }
        *cycles = clock64() - t0;
    } else if (blockIdx.x == peer) {
        theirs = slots;
        for (int i = 1; i <= iters; ++i) {
            while (ld_flag<MODE>(mine) != (unsigned long long)i) { }
            st_flag<MODE>(theirs, i);
        }
    }
}

// chain like the solver: every CTA, every phase: `nthreads` threads each store one 16 B LL message to both
// neighbours, then poll their own two slots; __syncthreads between phases.  No compute.
__device__ __forceinline__ void ll_store(uint4 *p, unsigned lo, unsigned hi, unsigned seq) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(seq), "r"(hi), "r"(seq) : "memory");
}
__device__ __forceinline__ uint4 ll_load(const uint4 *p) {
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__global__ void chain(uint4 *ll, int W, int phases, int active, int backoff, long long *cycles) {
    const int cta = blockIdx.x, P = gridDim.x, t = threadIdx.x;
    uint4 *up = cta > 0 ? ll + ((size_t)(cta - 1) * 2 + 1) * W : nullptr;
    uint4 *dn = cta + 1 < P ? ll + ((size_t)(cta + 1) * 2) * W : nullptr;
    const uint4 *it = ll + ((size_t)cta * 2) * W, *ib = ll + ((size_t)cta * 2 + 1) * W;
    long long t0 = clock64();
    for (int ph = 1; ph <= phases; ++ph) {
        if (t < active) {
            if (up) ll_store(up + t, ph, t, ph);
            if (dn) ll_store(dn + t, ph, t, ph);
            if (cta > 0) { uint4 r = ll_load(it + t); while (r.y < (unsigned)ph || r.w < (unsigned)ph) { if (backoff) __nanosleep(backoff); r = ll_load(it + t); } }
            if (cta + 1 < P) { uint4 r = ll_load(ib + t); while (r.y < (unsigned)ph || r.w < (unsigned)ph) { if (backoff) __nanosleep(backoff); r = ll_load(ib + t); } }
        }
        __syncthreads();
    }
    if (t == 0 && cta == P / 2) *cycles = clock64() - t0;
}


// chain2: the solver's phase structure with emulated compute.  Per phase: polls for the neighbours' previous
// messages are issued at the start and then every `gap` cycles (DEPTH polls in flight per slot), the thread
// "computes" for I cycles, consumes the messages (re-polling one at a time if none of the early polls saw them),
// "computes" B cycles, sends its own messages, __syncthreads.
__device__ __forceinline__ void spin_until(long long t) { while (clock64() < t) { } }
template <int DEPTH>
__global__ void chain2(uint4 *ll, int W, int phases, int I, int B, int gap, long long *cycles, unsigned *sink) {
    const int cta = blockIdx.x, P = gridDim.x, t = threadIdx.x;
    uint4 *up = cta > 0 ? ll + ((size_t)(cta - 1) * 2 + 1) * W : nullptr;
    uint4 *dn = cta + 1 < P ? ll + ((size_t)(cta + 1) * 2) * W : nullptr;
    const uint4 *it = ll + ((size_t)cta * 2) * W + t, *ib = ll + ((size_t)cta * 2 + 1) * W + t;
    const bool hu = cta > 0, hd = cta + 1 < P;
    unsigned acc = 0;
    long long t0 = clock64();
    for (int ph = 1; ph <= phases; ++ph) {
        const long long s0 = clock64();
        const unsigned want = (unsigned)(ph - 1);
        uint4 rt[DEPTH], rb[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            if (d) spin_until(s0 + (long long)d * gap);
            if (hu) rt[d] = ll_load(it);
            if (hd) rb[d] = ll_load(ib);
        }
        spin_until(s0 + I);
        if (hu) {
            bool got = false;
            uint4 r;
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) if (!got && rt[d].y >= want && rt[d].w >= want) { got = true; r = rt[d]; }
            if (!got) { r = ll_load(it); while (r.y < want || r.w < want) r = ll_load(it); }
            acc += r.x;
        }
        if (hd) {
            bool got = false;
            uint4 r;
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) if (!got && rb[d].y >= want && rb[d].w >= want) { got = true; r = rb[d]; }
            if (!got) { r = ll_load(ib); while (r.y < want || r.w < want) r = ll_load(ib); }
            acc += r.x;
        }
        spin_until(clock64() + B);
        if (up) ll_store(up + t, ph + acc * 0, t, ph);
        if (dn) ll_store(dn + t, ph, t, ph);
        __syncthreads();
    }
    if (t == 0 && cta == P / 2) *cycles = clock64() - t0;
    if (acc == 0xdeadbeef) *sink = acc;
}


// chain3: like chain2 but the halo is received by NPOLL dedicated warps that poll continuously (MSG loads in flight
// per lane) and hand the values to the compute warps through shared memory at the phase barrier.  Compute warps:
// "boundary" B cycles, send, "interior" I cycles.  blockDim = 512 + 32*NPOLL.
template <int NPOLL>
__global__ void __launch_bounds__(512 + 32 * NPOLL, 1) chain3(uint4 *ll, int W, int phases, int I, int B, long long *cycles, unsigned *sink) {
    __shared__ unsigned halo[2][1024];
    const int cta = blockIdx.x, P = gridDim.x, t = threadIdx.x;
    uint4 *up = cta > 0 ? ll + ((size_t)(cta - 1) * 2 + 1) * W : nullptr;
    uint4 *dn = cta + 1 < P ? ll + ((size_t)(cta + 1) * 2) * W : nullptr;
    const bool hu = cta > 0, hd = cta + 1 < P;
    unsigned acc = 0;
    long long t0 = clock64();
    if (t < 512) {
        for (int ph = 1; ph <= phases; ++ph) {
            if (ph > 1) acc += halo[0][t] + halo[1][t];     // halo of the previous phase, delivered by the poll warps
            spin_until(clock64() + B);
            if (up) ll_store(up + t, ph + acc * 0, t, ph);
            if (dn) ll_store(dn + t, ph, t, ph);
            spin_until(clock64() + I);
            __syncthreads();
        }
    } else {
        constexpr int MSG = 1024 / (32 * NPOLL);            // messages per lane (512 from above + 512 from below)
        const int lane = t - 512;
        for (int ph = 1; ph <= phases; ++ph) {
            // messages of THIS phase (the compute warps of the neighbours send them B cycles into the phase)
            const uint4 *base = ll + (size_t)cta * 2 * W;   // message id (0..1023) lives at base[(id >> 9) * W + (id & 511)]
            unsigned need = 0;
            uint4 r[MSG];
#pragma unroll
            for (int m = 0; m < MSG; ++m) {
                const int id = lane + m * 32 * NPOLL;
                if (id < 512 ? hu : hd) need |= 1u << m;
            }
            while (need) {
#pragma unroll
                for (int m = 0; m < MSG; ++m) {
                    const int id = lane + m * 32 * NPOLL;
                    if (need >> m & 1) r[m] = ll_load(base + (size_t)(id >> 9) * W + (id & 511));
                }
#pragma unroll
                for (int m = 0; m < MSG; ++m)
                    if ((need >> m & 1) && r[m].y >= (unsigned)ph && r[m].w >= (unsigned)ph) {
                        const int id = lane + m * 32 * NPOLL;
                        halo[id >> 9][id & 511] = r[m].x;
                        need &= ~(1u << m);
                    }
            }
            __syncthreads();
        }
    }
    if (t == 0 && cta == P / 2) *cycles = clock64() - t0;
    if (acc == 0xdeadbeef) *sink = acc;
}

// plain L2 load latency (pointer chase, one thread)
__global__ void chase(const unsigned *next, int iters, long long *cycles, unsigned *sink) {
    unsigned i = 0;
    long long t0 = clock64();
    for (int n = 0; n < iters; ++n) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(i) : "l"(next + i) : "memory");
    *cycles = clock64() - t0;
    *sink = i;
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    unsigned long long *slots; long long *cyc;
    CK(cudaMalloc(&slots, 32 * 8 * 256)); CK(cudaMallocManaged(&cyc, 8));
    const int iters = 20000;
    int peers[] = {1, 2, 37, 74, 147};
    for (int mode = 0; mode < 3; ++mode)
        for (int peer : peers) {
            CK(cudaMemset(slots, 0, 32 * 8 * 256));
            void *args[] = {&slots, (void *)&iters, &peer, &cyc};
            void *fn = mode == 0 ? (void *)pingpong<0> : mode == 1 ? (void *)pingpong<1> : (void *)pingpong<2>;
            CK(cudaLaunchCooperativeKernel(fn, dim3(148), dim3(32), args, 0, 0));
            CK(cudaDeviceSynchronize());
            printf("pingpong mode %d (0 volatile, 1 relaxed.gpu, 2 rel/acq.gpu) peer CTA %3d: %.0f cycles per one-way hop\n", mode, peer, (double)*cyc / iters / 2);
        }
    // pointer chase over 64 MB (L2 resident, > L1)
    {
        const int n = 1 << 24; unsigned *h = (unsigned *)malloc(n * 4), *d, *sink;
        for (int i = 0; i < n; ++i) h[i] = (unsigned)(((long long)i * 40503 + 12345) % n);  // scattered
        CK(cudaMalloc(&d, n * 4)); CK(cudaMalloc(&sink, 4)); CK(cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice));
        chase<<<1, 1>>>(d, 4000, cyc, sink); CK(cudaDeviceSynchronize());
        chase<<<1, 1>>>(d, 4000, cyc, sink); CK(cudaDeviceSynchronize());
        printf("dependent ld.volatile chase (64 MB footprint): %.0f cycles per load\n", (double)*cyc / 4000);
    }
    // solver-like chain
    uint4 *ll; const int W = 1024;
    CK(cudaMalloc(&ll, (size_t)148 * 2 * W * 16));
    int actives[] = {1, 32, 128, 512};
    int backs[] = {0, 100};
    for (int b : backs)
        for (int a : actives) {
            CK(cudaMemset(ll, 0, (size_t)148 * 2 * W * 16));
            int phases = 4000;
            void *args[] = {&ll, (void *)&W, &phases, &a, &b, &cyc};
            CK(cudaLaunchCooperativeKernel((void *)chain, dim3(148), dim3(512), args, 0, 0));
            CK(cudaDeviceSynchronize());
            printf("chain: 148 CTAs x %3d polling threads, backoff %3d ns: %.0f cycles per phase\n", a, b, (double)*cyc / phases);
        }
    {   // solver-shaped phases with emulated compute
        unsigned *sink; CK(cudaMalloc(&sink, 4));
        struct Cfg { int depth, I, B, gap; };
        Cfg cfgs[] = {{1, 0, 0, 0}, {1, 700, 300, 0}, {2, 700, 300, 350}, {3, 700, 300, 230}, {4, 700, 300, 175}, {1, 1000, 300, 0}, {3, 1000, 300, 330},
                      {4, 1000, 300, 250}, {1, 400, 200, 0}, {3, 400, 200, 130}, {4, 1400, 300, 350}};
        for (Cfg c : cfgs) {
            CK(cudaMemset(ll, 0, (size_t)148 * 2 * W * 16));
            int phases = 4000;
            void *args[] = {&ll, (void *)&W, &phases, &c.I, &c.B, &c.gap, &cyc, &sink};
            void *fn = c.depth == 1 ? (void *)chain2<1> : c.depth == 2 ? (void *)chain2<2> : c.depth == 3 ? (void *)chain2<3> : (void *)chain2<4>;
            CK(cudaLaunchCooperativeKernel(fn, dim3(148), dim3(512), args, 0, 0));
            CK(cudaDeviceSynchronize());
            printf("chain2: depth %d interior %4d boundary %3d gap %3d: %.0f cycles per phase (compute alone %d)\n", c.depth, c.I, c.B, c.gap,
                   (double)*cyc / phases, c.I + c.B);
        }
    }
    {   // dedicated polling warps
        unsigned *sink; CK(cudaMalloc(&sink, 4));
        struct Cfg { int npoll, I, B; };
        Cfg cfgs[] = {{2, 0, 0}, {2, 700, 300}, {2, 1070, 300}, {1, 1070, 300}, {4, 1070, 300}, {2, 1400, 300}, {2, 400, 200}};
        for (Cfg c : cfgs) {
            CK(cudaMemset(ll, 0, (size_t)148 * 2 * W * 16));
            int phases = 4000;
            void *args[] = {&ll, (void *)&W, &phases, &c.I, &c.B, &cyc, &sink};
            void *fn = c.npoll == 1 ? (void *)chain3<1> : c.npoll == 2 ? (void *)chain3<2> : (void *)chain3<4>;
            CK(cudaLaunchCooperativeKernel(fn, dim3(148), dim3(512 + 32 * c.npoll), args, 0, 0));
            CK(cudaDeviceSynchronize());
            printf("chain3: %d poll warps, boundary %3d interior %4d: %.0f cycles per phase (compute alone %d)\n", c.npoll, c.B, c.I,
                   (double)*cyc / phases, c.I + c.B);
        }
    }
    return 0;
}
