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
int resident_slots();               // halo slots (16 B each) per CTA link and direction
int run_resident(pcd_solver *s, const double *D, double *phi, int max_it, double tol, pcd_solve_info *info);


// dct_solver.cu: opt-in direct backend (PCD_SOLVER_DCT); D must be free of NaN
int run_dct(pcd_solver *s, const double *D, double *phi, pcd_solve_info *info);
void dct_free(pcd_solver *s);

// sor_tiled.cu
int tiled_sweeps_per_pass();
// sm_count: SMs of the device the pass runs on; sm_reserve: SMs the pass leaves free (overlapped multi-GPU exchange)
int tiled_pass(const double *phi_in, double *phi_out, const double *D, int W, int H, int row_first, int rows, int grow0,
               int nsweeps, unsigned long long *slots, int sm_count, int sm_reserve, cudaStream_t stream);

// Persistent multi-pass launch of the wavefront kernel with the ghost-row exchange fused in (multi-GPU slabs): the CTAs
// that finish the `gh` owned rows next to a slab edge store them to the neighbour's output field as well (peer memory
// over NVLink) and raise THEIR STRIP's flag in the neighbour's memory to the pass's sequence number; the edge CTAs of
// the next pass wait until the flags of their own and the two adjacent strips have reached the previous sequence number.
constexpr int WAVE_MAX_CTAS = 1024;   // size of the per-CTA sequence array of a slab
constexpr int WAVE_MAX_STRIPS = 512;  // size of a per-strip flag array
struct WavePeer {
    double *buf[2] = {nullptr, nullptr};                 // this slab's two field buffers (local row 0 = global row grow0)
    double *up_buf[2] = {nullptr, nullptr};              // the neighbours' buffers, same numbering (nullptr: no neighbour)
    double *dn_buf[2] = {nullptr, nullptr};
    int up_grow0 = 0, dn_grow0 = 0;
    int gh = 0;
    int cur = 0;                                          // buffer that holds the field before the first pass
    int npass = 0;
    const unsigned *wait_up = nullptr, *wait_dn = nullptr;  // local per-strip flags, written by the neighbours
    unsigned *sig_up = nullptr, *sig_dn = nullptr;           // the neighbours' per-strip flags this slab raises
    unsigned *done = nullptr;                                 // local: last published pass per CTA [WAVE_MAX_CTAS]
    int *err = nullptr;                                       // local: set when a wait ran into its time limit
    unsigned seq0 = 0;                                        // sequence number of the last pass before this launch
    int tail_rows = 0;                                        // rows of the short last chunk (filled by the launcher)
    unsigned long long *trace = nullptr;                      // diagnostics (PCD_WAVE_TRACE): per CTA and pass 4 x globaltimer
                                                              // [wait begin, pass begin, pass end, published], else nullptr
};
// dmap: tensor map (128 bytes, tiled_dmap_encode) of the parity-split copy of D, or nullptr for the cp.async staging
int tiled_run_peer(const double *D, int W, int H, int row_first, int rows, int grow0, int sweeps_per_pass,
                   unsigned long long *slots, const WavePeer &peer, const void *dmap, int sm_count, int sm_reserve,
                   cudaStream_t stream);
int tiled_dsplit_pitch(int W);                                                       // column pairs per row of the split copy
int tiled_dsplit(const double *D, double *S, int W, int rows, cudaStream_t stream);  // S[rows][2][Kp] <- D[rows][W]
int tiled_dmap_encode(void *map_out, const double *S, int W, int rows);
int tiled_strips(int W);

}  // namespace pcd
