// Shared between the streaming (sor_kernels.cu) and resident (sor_resident.cu) K-SOR kernels.
#pragma once

#include "pcd_internal.h"

namespace pcd {

struct SorW {
    double w[5];  // omega / cnt, cnt = 0..4 (w[0] = +inf as in the reference's division)
};

inline SorW make_w(int W) {
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


constexpr int RES_MAX_SWEEPS_PER_LAUNCH = 1 << 17;

struct ResState {  // device control block of the resident kernel (mirrored in pinned host memory)
    int sweeps;
    int converged_at;
    int error;   // a CTA gave up waiting for a neighbour (CTAs not co-resident): the launch is void, phi untouched
    int done;    // CTAs that finished their sweeps; phi is written back only after all of them have
};

// sor_resident.cu
int resident_plan(pcd_solver *s);   // 1 when the grid fits the on-chip path (fills s->res_*)
int run_resident(pcd_solver *s, const double *D, double *phi, int max_it, double tol, pcd_solve_info *info);


// sor_tiled.cu
int tiled_sweeps_per_pass();
// sm_count: SMs of the device the pass runs on; sm_reserve: SMs the pass leaves free (overlapped multi-GPU exchange)
int tiled_pass(const double *phi_in, double *phi_out, const double *D, int W, int H, int row_first, int rows, int grow0,
               int nsweeps, unsigned long long *slots, int sm_count, int sm_reserve, cudaStream_t stream);

// Ghost-row exchange fused into the wavefront pass (multi-GPU slabs): the CTAs that finish the `gh` owned rows next to
// a slab edge store them to the neighbour's output field as well (peer memory over NVLink) and the last of them
// raises the neighbour's flag to `seq`; the CTAs that read ghost rows first wait until their flag reached seq-1.
struct WavePeer {
    double *up_out = nullptr, *dn_out = nullptr;  // neighbours' phi_out arrays (local row 0 = global row *_grow0)
    int up_grow0 = 0, dn_grow0 = 0;
    int gh = 0;
    const unsigned *wait_up = nullptr, *wait_dn = nullptr;  // local flags, written by the neighbours
    unsigned *sig_up = nullptr, *sig_dn = nullptr;           // the neighbours' flags
    unsigned *cnt = nullptr;                                  // local arrival counters [2]
    int *err = nullptr;                                       // local: set when a wait ran into its time limit
    unsigned seq = 0;                                         // sequence number of this pass (first pass: 1)
    int tail_rows = 0;                                        // rows of the short last chunk (filled by the launcher)
};
int tiled_pass_peer(const double *phi_in, double *phi_out, const double *D, int W, int H, int row_first, int rows, int grow0,
                    int nsweeps, unsigned long long *slots, const WavePeer &peer, int sm_count, int sm_reserve,
                    cudaStream_t stream);

}  // namespace pcd
