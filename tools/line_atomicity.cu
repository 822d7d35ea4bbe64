// Scratch tool for round 2 (candidate i): is a warp's store to one 128-byte line (8 lanes x 16 B, one
// instruction) observed WHOLE by another SM's warp load of that line (8 lanes x 16 B, one instruction)?  And a 32-byte
// sector (2 lanes x 16 B)?  Line-granular halo messages (payload + one flag per line / sector instead of a flag in every
// 8 bytes) are only admissible if no torn read ever shows up here.
//   nvcc -arch=sm_100a -O3 -o tools/line_atomicity tools/line_atomicity.cu && tools/line_atomicity [seconds]
// Every producer CTA hammers its own set of lines with ever increasing sequence numbers (all 16 u64 words of a line
// carry the same number); every consumer CTA reads the lines of "its" producer and counts lines whose words differ
// (a) within one 16 B lane access, (b) within a 32 B sector, (c) within the 128 B line.  (a) must be 0 by the PTX rules
// for vector accesses of naturally aligned 8-byte halves only if it is 0 here as well -- the LL protocol in the
// solver does not rely on it (each 8-byte half carries its own flag).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int LINES = 64;            // lines per producer
struct Counts { unsigned long long reads, torn16, torn32, torn128; };

__device__ __forceinline__ void st16(void *p, unsigned long long a, unsigned long long b) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)a), "r"((unsigned)(a >> 32)),
                 "r"((unsigned)b), "r"((unsigned)(b >> 32)) : "memory");
}
__device__ __forceinline__ void ld16(const void *p, unsigned long long &a, unsigned long long &b) {
    unsigned x, y, z, w;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "l"(p) : "memory");
    a = ((unsigned long long)y << 32) | x;
    b = ((unsigned long long)w << 32) | z;
}

// grid = 2 * pairs CTAs of 32 threads: even CTAs produce, odd CTAs consume the lines of CTA-1 (different SMs: 1 CTA/SM
// is not enforced, the scheduler spreads 32-thread CTAs over all SMs first)
__global__ void hammer(unsigned char *buf, volatile int *stop, Counts *out) {
    const int pair = blockIdx.x >> 1, lane = threadIdx.x;
    unsigned char *mine = buf + (size_t)pair * LINES * 128;
    const int sub = lane & 7, grp = lane >> 3;          // 4 groups of 8 lanes: each group handles one line per step
    if ((blockIdx.x & 1) == 0) {
        unsigned long long seq = 1;
        while (!*stop) {
            for (int l = grp; l < LINES; l += 4) st16(mine + l * 128 + sub * 16, seq, seq);
            ++seq;
        }
    } else {
        unsigned long long reads = 0, t16 = 0, t32 = 0, t128 = 0;
        while (!*stop) {
            for (int l = grp; l < LINES; l += 4) {
                unsigned long long a, b;
                ld16(mine + l * 128 + sub * 16, a, b);
                const unsigned m8 = 0xffu << (grp * 8);
                const bool bad16 = a != b;
                // sector = lanes (2s, 2s+1); line = the 8 lanes of the group
                const unsigned long long nb = __shfl_xor_sync(0xffffffffu, a, 1);
                const bool bad32 = a != nb;
                const unsigned long long first = __shfl_sync(0xffffffffu, a, grp * 8);
                const bool bad128 = a != first;
                const unsigned v16 = __ballot_sync(0xffffffffu, bad16) & m8, v32 = __ballot_sync(0xffffffffu, bad32) & m8,
                               v128 = __ballot_sync(0xffffffffu, bad128) & m8;
                if (sub == 0) { ++reads; t16 += v16 != 0; t32 += v32 != 0; t128 += v128 != 0; }
            }
        }
        if (sub == 0) {
            atomicAdd(&out->reads, reads); atomicAdd(&out->torn16, t16); atomicAdd(&out->torn32, t32); atomicAdd(&out->torn128, t128);
        }
    }
}

int main(int argc, char **argv) {
    const double seconds = argc > 1 ? atof(argv[1]) : 5.0;
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int pairs = sms / 2;
    unsigned char *buf; int *stop; Counts *out;
    CK(cudaMalloc(&buf, (size_t)pairs * LINES * 128)); CK(cudaMemset(buf, 0, (size_t)pairs * LINES * 128));
    CK(cudaMallocManaged(&out, sizeof(Counts))); *out = Counts{};
    CK(cudaHostAlloc(&stop, sizeof(int), cudaHostAllocMapped)); *stop = 0;
    int *dstop; CK(cudaHostGetDevicePointer(&dstop, stop, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    hammer<<<2 * pairs, 32>>>(buf, dstop, out);
    CK(cudaGetLastError());
    struct timespec ts = {(time_t)seconds, (long)((seconds - (long)seconds) * 1e9)};
    nanosleep(&ts, nullptr);
    *(volatile int *)stop = 1;
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%d producer/consumer pairs, %.1f s: %llu line reads, torn within 16 B: %llu, within a 32 B sector: %llu, within the 128 B line: %llu\n",
           pairs, ms * 1e-3, out->reads, out->torn16, out->torn32, out->torn128);
    return 0;
}
