// K-SOR: red-black SOR for the 5-point Poisson problem, fp64, sm_100a.
//
// Restates src/solver.cpp:12-61 (patial_relax) + :70-147 (poisson_solver) of the reference with a
// red-black ordering ((x+y) even first, then odd): same omega (:71), same per-cell update (:47),
// same neighbour rule (:29-44: a neighbour counts when it is inside the grid and its D is not NaN),
// same stopping rule (:50-53,:142: max|delta| of a full sweep < tol), phi warm-started in place.
// The expression order of the update is kept literally and the library is compiled with
// -fmad=false, so the field after n sweeps is bit-identical to the red-black restatement
// kept under oracle/ (test infrastructure; nothing here includes or links it).
//
// Two kernel families:
//   * streaming : one launch per colour, phi/D streamed through L2/HBM; any grid size.
//   * resident  : ONE persistent cooperative kernel per solve.  Each CTA (one per SM) owns a slab of
//                 rows; phi lives in shared memory (column-parity split, conflict-free), D / phi /
//                 neighbour masks of the thread's own cells live in registers; slab boundary rows are
//                 exchanged between neighbouring CTAs through L2 with flag-in-data ("LL") 16-byte
//                 messages, so there is no grid-wide barrier and no fence on the critical path; the
//                 convergence test is a per-sweep atomicMax + arrival counter evaluated with a fixed
//                 lag by a dedicated control warp.
#include "pcd_internal.h"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace pcd {

struct SorW {
    double w[5];  // omega / cnt, cnt = 0..4 (w[0] = +inf as in the reference's division)
};

static SorW make_w(int W) {
    SorW r;
    double omega = sor_omega(W);
    for (int c = 0; c < 5; ++c) r.w[c] = omega / (double)c;
    return r;
}

__device__ __forceinline__ double wsel(const SorW &w, int cnt) {
    return cnt == 4 ? w.w[4] : (cnt == 3 ? w.w[3] : (cnt == 2 ? w.w[2] : (cnt == 1 ? w.w[1] : w.w[0])));
}

__device__ __forceinline__ double warp_max(double a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double b = __shfl_xor_sync(0xffffffffu, a, o);
        a = b > a ? b : a;
    }
    return a;
}

// ------------------------------------------------------------------------------------------------
// neighbour masks (bit0 left, bit1 up, bit2 right, bit3 down), only needed when D has NaN holes
// ------------------------------------------------------------------------------------------------
__global__ void detect_nan_kernel(const double *__restrict__ D, long n, int *flag) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long stride = (long)gridDim.x * blockDim.x;
    int found = 0;
    for (; i < n; i += stride) found |= isnan(D[i]);
    if (__any_sync(0xffffffffu, found) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

__global__ void build_mask_kernel(const double *__restrict__ D, unsigned char *__restrict__ mask, int W, int H) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    size_t i = (size_t)y * W + x;
    unsigned m = 0;
    if (x != 0 && !isnan(D[i - 1])) m |= 1;
    if (y != 0 && !isnan(D[i - W])) m |= 2;
    if (x != W - 1 && !isnan(D[i + 1])) m |= 4;
    if (y != H - 1 && !isnan(D[i + W])) m |= 8;
    mask[i] = (unsigned char)m;
}

// ------------------------------------------------------------------------------------------------
// streaming path
// ------------------------------------------------------------------------------------------------
template <bool MASKED>
__global__ void __launch_bounds__(256)
sor_colour_kernel(double *__restrict__ phi, const double *__restrict__ D, const unsigned char *__restrict__ mask,
                  int W, int H, int colour, SorW w, unsigned long long *__restrict__ slot) {
    const int y = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = 2 * k + ((y + colour) & 1);
    double a = 0.0;
    if (x < W) {
        const size_t i = (size_t)y * W + x;
        unsigned m;
        if (MASKED) m = mask[i];
        else m = (x != 0 ? 1u : 0u) | (y != 0 ? 2u : 0u) | (x != W - 1 ? 4u : 0u) | (y != H - 1 ? 8u : 0u);
        const double val = phi[i];
        double s = 0.0;
        if (m & 1) s += phi[i - 1];
        if (m & 2) s += phi[i - W];
        if (m & 4) s += phi[i + 1];
        if (m & 8) s += phi[i + W];
        const int cnt = __popc(m);
        const double delta = wsel(w, cnt) * (s - (double)cnt * val - D[i]);
        const double ad = fabs(delta);
        a = ad > 0.0 ? ad : 0.0;  // NaN never enters the max (src/solver.cpp:50-53)
        phi[i] = val + delta;
    }
    a = warp_max(a);
    __shared__ double wm[8];
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        a = threadIdx.x < (blockDim.x >> 5) ? wm[threadIdx.x] : 0.0;
        a = warp_max(a);
        if (threadIdx.x == 0 && a > 0.0) atomicMax(slot, (unsigned long long)__double_as_longlong(a));
    }
}

static int run_streaming(pcd_solver *s, const double *D, double *phi, int max_it, double tol, bool masked,
                         pcd_solve_info *info) {
    const int W = s->W, H = s->H;
    const SorW w = make_w(W);
    const int K = (W + 1) / 2;
    dim3 block(256), grid((K + 255) / 256, H);
    int chunk = s->check_lag > 0 ? s->check_lag : 64;
    if (chunk > 4096) chunk = 4096;  // size of the pinned mirror
    int done = 0, conv = 0;
    double last = 0.0;
    while (done < max_it && !conv) {
        const int k = max_it - done < chunk ? max_it - done : chunk;
        PCD_CUDA(cudaMemsetAsync(s->sweep_max, 0, sizeof(unsigned long long) * k, s->stream));
        PCD_CUDA(cudaEventRecord(s->evk0, s->stream));
        for (int j = 0; j < k; ++j)
            for (int colour = 0; colour < 2; ++colour) {
                if (masked)
                    sor_colour_kernel<true><<<grid, block, 0, s->stream>>>(phi, D, s->mask, W, H, colour, w, s->sweep_max + j);
                else
                    sor_colour_kernel<false><<<grid, block, 0, s->stream>>>(phi, D, nullptr, W, H, colour, w, s->sweep_max + j);
                PCD_LAUNCHED();
                info->launches++;
            }
        PCD_CUDA(cudaEventRecord(s->evk1, s->stream));
        PCD_CUDA(cudaMemcpyAsync(s->h_sweep_max, s->sweep_max, sizeof(unsigned long long) * k, cudaMemcpyDeviceToHost, s->stream));
        PCD_CUDA(cudaStreamSynchronize(s->stream));
        {
            float kms = 0.f;
            PCD_CUDA(cudaEventElapsedTime(&kms, s->evk0, s->evk1));
            info->kernel_ms += kms;
        }
        for (int j = 0; j < k; ++j) {
            double m;
            memcpy(&m, &s->h_sweep_max[j], sizeof(double));
            if (!conv && m < tol) {
                conv = done + j + 1;
                last = m;
            }
            if (!conv && j == k - 1) last = m;
        }
        done += k;
    }
    info->sweeps = done;
    info->converged_at = conv;
    info->last_max_update = last;
    return PCD_OK;
}

// ------------------------------------------------------------------------------------------------
// resident path
// ------------------------------------------------------------------------------------------------
constexpr int RES_COMPUTE_THREADS = 480;  // 15 warps + 1 control warp = 16 warps = 4 per SM sub-partition (128 regs each)
constexpr int RES_THREADS = RES_COMPUTE_THREADS + 32;  // + control warp
constexpr int RES_MAX_SWEEPS_PER_LAUNCH = 1 << 17;

struct ResState {  // device control block (also mirrored in pinned host memory)
    int sweeps;
    int converged_at;
    int pad0, pad1;
};

struct ResParams {
    double *phi;              // global field, in/out
    const double *D;
    int W, H, K, Kp;          // K = ceil(W/2) column pairs per row, Kp = padded pitch (doubles) of one parity array
    int P;                    // CTAs
    int max_it;               // sweeps this launch may execute
    int lag;                  // convergence lag L
    double tol;
    SorW w;
    uint4 *ll;                // LL halo slots: [P][2][W] x 16 B
    unsigned long long *g_max;  // [max_it]
    unsigned int *g_cnt;        // [max_it]
    ResState *state;
};

__device__ __forceinline__ void ll_store(uint4 *p, double v, unsigned seq) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)b), "r"(seq),
                 "r"((unsigned)(b >> 32)), "r"(seq)
                 : "memory");
}

__device__ __forceinline__ double ll_wait(const uint4 *p, unsigned seq) {
    unsigned lo, f0, hi, f1;
    do {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(p) : "memory");
    } while (f0 != seq || f1 != seq);
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// One item = the column pair (2k, 2k+1) of one slab row; exactly one of its two cells is updated per
// colour phase.  Everything that is invariant across sweeps sits in registers.
template <int MAXI>
__global__ void __launch_bounds__(RES_THREADS, 1) sor_resident_kernel(ResParams p) {
    extern __shared__ double smem[];
    __shared__ unsigned long long blkmax[2];
    __shared__ int s_stop_after;  // number of sweeps after which every CTA leaves the loop (0 = keep going)

    const int tid = threadIdx.x;
    const int cta = blockIdx.x;
    const int W = p.W, H = p.H, K = p.K, Kp = p.Kp;
    const int r0 = (int)(((long long)cta * H) / p.P), r1 = (int)(((long long)(cta + 1) * H) / p.P);
    const int nr = r1 - r0;
    // smem: local row ly in [0, nr+2) (ly = 0 / nr+1 are the halo rows), parity q in {0,1}:
    //   sm[(ly*2 + q)*Kp + 1 + k]   (+1: one pad double in front so that k-1 = -1 is addressable)
    double *sm = smem;
    auto sidx = [Kp](int ly, int q, int k) { return (ly * 2 + q) * Kp + 1 + k; };

    if (tid == 0) {
        blkmax[0] = 0ull;
        blkmax[1] = 0ull;
        s_stop_after = 0;
    }

    const bool is_compute = tid < RES_COMPUTE_THREADS;
    // ---- per-item invariants ---------------------------------------------------------------
    // packed per item: bits 0-7 local row ly (1..nr, 0 = unused) | bits 8-19 flags m | bits 20-31 column pair k
    //   m: bits 0-3 neighbour mask of cell 0 (x = 2k), 4-7 of cell 1, 8/9 cell valid, 10 top row, 11 bottom row
    unsigned it_meta[MAXI];
    double it_D[MAXI][2], it_v[MAXI][2];
#pragma unroll
    for (int j = 0; j < MAXI; ++j) {
        it_meta[j] = 0;
        it_D[j][0] = it_D[j][1] = it_v[j][0] = it_v[j][1] = 0.0;
        const int id = tid + j * RES_COMPUTE_THREADS;
        if (is_compute && id < nr * K) {
            const int ro = id / K, k = id - ro * K;
            // row order: top boundary row, bottom boundary row, then the interior rows
            const int ly = ro == 0 ? 1 : (ro == 1 ? nr : ro);
            const int y = r0 + ly - 1;
            unsigned m = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int x = 2 * k + q;
                if (x < W) {
                    const size_t i = (size_t)y * W + x;
                    unsigned mm = 0;
                    if (x != 0 && !isnan(p.D[i - 1])) mm |= 1;
                    if (y != 0 && !isnan(p.D[i - W])) mm |= 2;
                    if (x != W - 1 && !isnan(p.D[i + 1])) mm |= 4;
                    if (y != H - 1 && !isnan(p.D[i + W])) mm |= 8;
                    m |= mm << (4 * q);
                    m |= 1u << (8 + q);
                    it_D[j][q] = p.D[i];
                    it_v[j][q] = p.phi[i];
                    sm[sidx(ly, q, k)] = it_v[j][q];
                }
            }
            if (ly == 1) m |= 1u << 10;
            if (ly == nr) m |= 1u << 11;
            it_meta[j] = (unsigned)ly | (m << 8) | ((unsigned)k << 20);
        }
    }
    // halo rows from the global field (state before the first phase)
    if (is_compute) {
        for (int k = tid; k < K; k += RES_COMPUTE_THREADS)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int x = 2 * k + q;
                if (x < W) {
                    if (r0 > 0) sm[sidx(0, q, k)] = p.phi[(size_t)(r0 - 1) * W + x];
                    if (r1 < H) sm[sidx(nr + 1, q, k)] = p.phi[(size_t)r1 * W + x];
                }
            }
    }
    __syncthreads();

    uint4 *ll_up = cta > 0 ? p.ll + ((size_t)(cta - 1) * 2 + 1) * W : nullptr;      // neighbour above: its "from below" slots
    uint4 *ll_dn = cta + 1 < p.P ? p.ll + ((size_t)(cta + 1) * 2 + 0) * W : nullptr;  // neighbour below: its "from above" slots
    const uint4 *ll_in_top = p.ll + ((size_t)cta * 2 + 0) * W;
    const uint4 *ll_in_bot = p.ll + ((size_t)cta * 2 + 1) * W;

    const int max_it = p.max_it;
    int s = 0;
    if (is_compute) {
        // =============================== compute warps ===========================================
        for (;; ++s) {
            double lmax = 0.0;
#pragma unroll
            for (int colour = 0; colour < 2; ++colour) {
                const unsigned seq = 2u * (unsigned)s + (unsigned)colour + 1u;
#pragma unroll
                for (int j = 0; j < MAXI; ++j) {
                    const int ly = (int)(it_meta[j] & 255u);
                    if (ly == 0) continue;
                    const int k = (int)(it_meta[j] >> 20);
                    const int y = r0 + ly - 1;
                    const int q = (y + colour) & 1;  // column parity of this row's active cell
                    const unsigned m = (it_meta[j] >> 8) & 0xfffu;
                    if (!((m >> (8 + q)) & 1u)) continue;
                    const unsigned mm = (m >> (4 * q)) & 15u;
                    // neighbours: left/right live in the other parity array of the same row
                    const int kl = q ? k : k - 1, kr = q ? k + 1 : k;
                    const double vl = (mm & 1) ? sm[sidx(ly, q ^ 1, kl)] : 0.0;
                    const double vu = (mm & 2) ? sm[sidx(ly - 1, q, k)] : 0.0;
                    const double vr = (mm & 4) ? sm[sidx(ly, q ^ 1, kr)] : 0.0;
                    const double vd = (mm & 8) ? sm[sidx(ly + 1, q, k)] : 0.0;
                    double sum = 0.0;
                    if (mm & 1) sum += vl;
                    if (mm & 2) sum += vu;
                    if (mm & 4) sum += vr;
                    if (mm & 8) sum += vd;
                    const int cnt = __popc(mm);
                    const double val = q ? it_v[j][1] : it_v[j][0];
                    const double Dv = q ? it_D[j][1] : it_D[j][0];
                    const double delta = wsel(p.w, cnt) * (sum - (double)cnt * val - Dv);
                    const double ad = fabs(delta);
                    if (ad > lmax) lmax = ad;
                    const double nv = val + delta;
                    if (q) it_v[j][1] = nv; else it_v[j][0] = nv;
                    sm[sidx(ly, q, k)] = nv;
                    const int x = 2 * k + q;
                    if ((m & (1u << 10)) && ll_up) ll_store(ll_up + x, nv, seq);
                    if ((m & (1u << 11)) && ll_dn) ll_store(ll_dn + x, nv, seq);
                }
                // import the neighbours' boundary cells of this colour (produced in their phase `seq`)
                for (int k = tid; k < K; k += RES_COMPUTE_THREADS) {
                    if (cta > 0) {
                        const int q = (r0 - 1 + colour) & 1, x = 2 * k + q;
                        if (x < W) sm[sidx(0, q, k)] = ll_wait(ll_in_top + x, seq);
                    }
                    if (cta + 1 < p.P) {
                        const int q = (r1 + colour) & 1, x = 2 * k + q;
                        if (x < W) sm[sidx(nr + 1, q, k)] = ll_wait(ll_in_bot + x, seq);
                    }
                }
                if (colour == 0) {
                    bar_sync(1, RES_COMPUTE_THREADS);  // compute warps only
                } else {
                    lmax = warp_max(lmax);
                    if ((tid & 31) == 0 && lmax > 0.0)
                        atomicMax(&blkmax[s & 1], (unsigned long long)__double_as_longlong(lmax));
                    bar_sync(0, RES_THREADS);          // end of sweep, with the control warp
                }
            }
            const int stop = s_stop_after;
            if (stop == s + 1 || s + 1 >= max_it) break;
        }
        // ---- write the slab back ----------------------------------------------------------------
#pragma unroll
        for (int j = 0; j < MAXI; ++j) {
            const int ly = (int)(it_meta[j] & 255u);
            if (ly == 0) continue;
            const int y = r0 + ly - 1, k = (int)(it_meta[j] >> 20);
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if ((it_meta[j] >> (16 + q)) & 1u) p.phi[(size_t)y * W + 2 * k + q] = it_v[j][q];
        }
    } else {
        // =============================== control warp =============================================
        const int lane = tid & 31;
        int conv_at = 0;
        for (;; ++s) {
            bar_sync(0, RES_THREADS);  // end of sweep s
            const int stop = s_stop_after;
            if (lane == 0) {
                const unsigned long long bm = blkmax[s & 1];
                blkmax[s & 1] = 0ull;
                if (bm) atomicMax(p.g_max + s, bm);
                __threadfence();
                atomicAdd(p.g_cnt + s, 1u);
            }
            if (stop == s + 1 || s + 1 >= max_it) break;
            // evaluate sweep e = s - lag: acted on by everybody after sweep s+1
            const int e = s - p.lag;
            if (lane == 0 && e >= 0 && conv_at == 0) {
                const volatile unsigned int *c = p.g_cnt + e;
                while (*c != (unsigned)p.P) { }
                __threadfence();
                const unsigned long long bits = *((const volatile unsigned long long *)(p.g_max + e));
                if (__longlong_as_double((long long)bits) < p.tol) {
                    conv_at = e + 1;
                    s_stop_after = s + 2;
                }
            }
        }
        // tail: sweeps whose convergence test was never acted on still have to be scanned by the host
        if (lane == 0 && cta == 0) {
            p.state->sweeps = s + 1;
            p.state->converged_at = conv_at;
        }
    }
}

static size_t resident_smem_bytes(int nr_max, int Kp) { return (size_t)(nr_max + 2) * 2 * Kp * sizeof(double); }

static int resident_plan(pcd_solver *s) {
    const int W = s->W, H = s->H;
    const int P = s->sm_count < H ? s->sm_count : H;
    const int nr_max = (H + P - 1) / P;
    const int K = (W + 1) / 2;
    const int Kp = ((K + 2 + 3) / 4) * 4;  // pad: 1 double in front, >=1 behind, pitch multiple of 32 B
    const long items = (long)nr_max * K;
    int maxi = 0;
    if (items <= 2L * RES_COMPUTE_THREADS) maxi = 2;
    else if (items <= 4L * RES_COMPUTE_THREADS) maxi = 4;
    else if (items <= 8L * RES_COMPUTE_THREADS) maxi = 8;
    const size_t smem = resident_smem_bytes(nr_max, Kp);
    if (!maxi || smem > 200 * 1024 || nr_max > 250 || K > 4095) return 0;
    s->res_ctas = P;
    s->res_rows_per_cta = nr_max;
    s->res_smem = smem;
    s->res_threads = maxi;  // reused: items per thread
    return 1;
}

template <int MAXI>
static int launch_resident(pcd_solver *s, ResParams &prm) {
    PCD_CUDA(cudaFuncSetAttribute(sor_resident_kernel<MAXI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->res_smem));
    void *args[] = {&prm};
    PCD_CUDA(cudaLaunchCooperativeKernel((void *)sor_resident_kernel<MAXI>, dim3(s->res_ctas), dim3(RES_THREADS), args,
                                         s->res_smem, s->stream));
    PCD_LAUNCHED();
    return PCD_OK;
}

static int run_resident(pcd_solver *s, const double *D, double *phi, int max_it, double tol, pcd_solve_info *info) {
    const int W = s->W, H = s->H;
    const int P = s->res_ctas;
    const int lag = s->check_lag > 0 ? s->check_lag : 4;
    int done = 0, conv = 0;
    double last = 0.0;
    unsigned long long *g_max = s->sweep_max;
    unsigned int *g_cnt = (unsigned int *)(s->sweep_max + s->ring);
    while (done < max_it && !conv) {
        const int k = max_it - done < RES_MAX_SWEEPS_PER_LAUNCH ? max_it - done : RES_MAX_SWEEPS_PER_LAUNCH;
        PCD_CUDA(cudaMemsetAsync(s->sweep_max, 0, (sizeof(unsigned long long) + sizeof(unsigned int)) * (size_t)s->ring, s->stream));
        PCD_CUDA(cudaMemsetAsync(s->halo, 0, (size_t)P * 2 * W * sizeof(uint4), s->stream));
        ResParams prm;
        prm.phi = phi; prm.D = D; prm.W = W; prm.H = H; prm.K = (W + 1) / 2;
        prm.Kp = ((prm.K + 2 + 3) / 4) * 4;
        prm.P = P; prm.max_it = k; prm.lag = lag; prm.tol = tol; prm.w = make_w(W);
        prm.ll = (uint4 *)s->halo; prm.g_max = g_max; prm.g_cnt = g_cnt; prm.state = (ResState *)s->res_state;
        int rc;
        PCD_CUDA(cudaEventRecord(s->evk0, s->stream));
        switch (s->res_threads) {
            case 2: rc = launch_resident<2>(s, prm); break;
            case 4: rc = launch_resident<4>(s, prm); break;
            default: rc = launch_resident<8>(s, prm); break;
        }
        PCD_TRY(rc);
        PCD_CUDA(cudaEventRecord(s->evk1, s->stream));
        info->launches++;
        PCD_CUDA(cudaMemcpyAsync(s->h_res_state, s->res_state, sizeof(ResState), cudaMemcpyDeviceToHost, s->stream));
        PCD_CUDA(cudaStreamSynchronize(s->stream));
        const ResState st = *(ResState *)s->h_res_state;
        {
            float kms = 0.f;
            PCD_CUDA(cudaEventElapsedTime(&kms, s->evk0, s->evk1));
            info->kernel_ms += kms;
        }
        // the device acts on the test with a lag: when the cap ends the launch, the last lag+1 sweeps
        // were executed but never tested -- scan them here so converged_at is exact
        int conv_local = st.converged_at;
        const int first = conv_local ? conv_local - 1 : (st.sweeps - (lag + 2) > 0 ? st.sweeps - (lag + 2) : 0);
        const int cnt = conv_local ? 1 : st.sweeps - first;
        PCD_CUDA(cudaMemcpyAsync(s->h_sweep_max, g_max + first, sizeof(unsigned long long) * cnt, cudaMemcpyDeviceToHost, s->stream));
        PCD_CUDA(cudaStreamSynchronize(s->stream));
        for (int j = 0; j < cnt; ++j) {
            memcpy(&last, &s->h_sweep_max[j], sizeof(double));
            if (!conv_local && last < tol) { conv_local = first + j + 1; break; }
        }
        if (conv_local) conv = done + conv_local;
        done += st.sweeps;
    }
    info->sweeps = done;
    info->converged_at = conv;
    info->last_max_update = last;
    return PCD_OK;
}

// ------------------------------------------------------------------------------------------------
// solver object
// ------------------------------------------------------------------------------------------------
int solver_init(pcd_solver *s, int W, int H, int device, int path, cudaStream_t stream) {
    if (W < 1 || H < 1 || H > 65535) {
        set_error("poisson solver: unsupported grid %dx%d", W, H);
        return PCD_ERR_INVALID;
    }
    s->W = W; s->H = H; s->device = device; s->path_req = path;
    PCD_TRY(select_device(device));
    cudaDeviceProp prop;
    PCD_CUDA(cudaGetDeviceProperties(&prop, device));
    s->sm_count = prop.multiProcessorCount;
    if (stream) {
        s->stream = stream;
        s->own_stream = false;
    } else {
        PCD_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        s->own_stream = true;
    }
    s->ring = RES_MAX_SWEEPS_PER_LAUNCH;
    PCD_CUDA(cudaMalloc(&s->sweep_max, (sizeof(unsigned long long) + sizeof(unsigned int)) * (size_t)s->ring));
    PCD_CUDA(cudaMallocHost(&s->h_sweep_max, sizeof(unsigned long long) * 4096));
    PCD_CUDA(cudaMalloc(&s->d_flags, sizeof(int) * 4));
    PCD_CUDA(cudaMallocHost(&s->h_flags, sizeof(int) * 4));
    PCD_CUDA(cudaMalloc(&s->res_state, sizeof(ResState)));
    PCD_CUDA(cudaMallocHost(&s->h_res_state, sizeof(ResState)));
    PCD_CUDA(cudaEventCreate(&s->ev0));
    PCD_CUDA(cudaEventCreate(&s->ev1));
    PCD_CUDA(cudaEventCreate(&s->evk0));
    PCD_CUDA(cudaEventCreate(&s->evk1));
    s->path_used = PCD_SOLVER_STREAMING;
    if (path != PCD_SOLVER_STREAMING) {
        int coop = 0;
        PCD_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        if (coop && resident_plan(s)) {
            s->path_used = PCD_SOLVER_RESIDENT;
            PCD_CUDA(cudaMalloc(&s->halo, (size_t)s->res_ctas * 2 * W * sizeof(uint4)));
        } else if (path == PCD_SOLVER_RESIDENT) {
            set_error("poisson solver: grid %dx%d does not fit the resident path on this device", W, H);
            return PCD_ERR_UNSUPPORTED;
        }
    }
    return PCD_OK;
}

void solver_free(pcd_solver *s) {
    if (!s) return;
    cudaFree(s->sweep_max); cudaFreeHost(s->h_sweep_max);
    cudaFree(s->mask); cudaFree(s->d_flags); cudaFreeHost(s->h_flags);
    cudaFree(s->res_state); cudaFreeHost(s->h_res_state); cudaFree(s->halo);
    if (s->own_fields) { cudaFree(s->D); cudaFree(s->phi); }
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->evk0) cudaEventDestroy(s->evk0);
    if (s->evk1) cudaEventDestroy(s->evk1);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    *s = pcd_solver();
}

int solver_run(pcd_solver *s, const double *D, double *phi, int max_iterations, double tol, pcd_solve_info *info) {
    pcd_solve_info local{};
    if (!info) info = &local;
    *info = pcd_solve_info{};
    info->path = s->path_used;
    PCD_TRY(select_device(s->device));
    if (max_iterations <= 0) return PCD_OK;  // the reference's loop body never runs (src/solver.cpp:92)
    PCD_CUDA(cudaEventRecord(s->ev0, s->stream));
    int rc;
    if (s->path_used == PCD_SOLVER_RESIDENT) {
        rc = run_resident(s, D, phi, max_iterations, tol, info);
    } else {
        // NaN holes change the neighbour rule (src/solver.cpp:29-44); only then is a mask array built
        const long n = (long)s->W * s->H;
        PCD_CUDA(cudaMemsetAsync(s->d_flags, 0, sizeof(int), s->stream));
        detect_nan_kernel<<<296, 256, 0, s->stream>>>(D, n, s->d_flags);
        PCD_LAUNCHED();
        info->launches++;
        PCD_CUDA(cudaMemcpyAsync(s->h_flags, s->d_flags, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        PCD_CUDA(cudaStreamSynchronize(s->stream));
        const bool masked = s->h_flags[0] != 0;
        if (masked) {
            if (!s->mask) PCD_CUDA(cudaMalloc(&s->mask, (size_t)n));
            build_mask_kernel<<<dim3((s->W + 255) / 256, s->H), 256, 0, s->stream>>>(D, s->mask, s->W, s->H);
            PCD_LAUNCHED();
            info->launches++;
        }
        rc = run_streaming(s, D, phi, max_iterations, tol, masked, info);
    }
    PCD_TRY(rc);
    PCD_CUDA(cudaEventRecord(s->ev1, s->stream));
    PCD_CUDA(cudaEventSynchronize(s->ev1));
    float ms = 0.f;
    PCD_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    info->device_ms = ms;
    return PCD_OK;
}

}  // namespace pcd
