// K-SOR, wavefront streaming path for grids that do not fit on chip (C5: 8192 x 8192): temporal blocking by
// time-skewed row streaming.
//
// One pass advances the field by TS full red-black sweeps (NP = 2*TS colour phases) while phi and D cross HBM
// once in and phi once out (ping-pong fields).  A CTA owns a strip of 2*NT columns (minus a 2*TS-column skirt
// on each side that goes stale by one column per phase and is never written) and a chunk of rows, and walks
// DOWN the rows: at step f, phase p updates row f - 1 - 2p.  With a skew of two rows per phase every phase of a
// step only reads values produced in earlier steps (phase p at row r needs phase p-1 at rows r-1..r+1, done at
// step f-1 at the latest), so the whole step needs ONE barrier and updates are in place.  Thread k owns the
// column pair (2k, 2k+1): its rolling window of R rows lives in registers (ring indices AND the colour parity
// of every update are compile-time: the row loop is unrolled by the even period R and chunks start on a fixed
// parity), so up/down/own values are register reads; the single left/right neighbour owned by thread k-1/k+1
// goes through a shared-memory ring that also receives the rows streamed in by cp.async PF steps ahead (no
// staging registers, no extra barrier).  Blocks of R steps that lie entirely inside the chunk and away from the
// domain's top/bottom rows run a check-free variant.  The chunk starts NP rows early and ends NP rows late
// (staleness advances one row per phase, so owned rows stay exact); per-sweep maxima are taken over owned cells
// only.  Every owned value equals the global red-black iteration bit for bit.
//
// PEER variant (multi-GPU row slabs, and any run of several passes): ONE persistent cooperative launch executes a
// whole block of passes.  There is no grid-wide barrier between passes: a CTA starts pass q as soon as its (up to)
// eight neighbouring CTAs have published pass q-1 (a per-CTA sequence word, st.release / ld.acquire at gpu scope) --
// they are the only CTAs whose output it reads and whose input it overwrites (ping-pong fields) -- and, for the CTAs
// next to a slab edge, as soon as the neighbouring GPU's edge CTAs of the same and the adjacent strips have raised
// their per-strip flag for pass q-1 (st.release.sys into this GPU's memory over NVLink).  The edge CTAs store the GH
// rows next to the slab edge into the neighbour's field as well (plain peer stores) and raise their own strip's flag
// there as soon as those rows are complete, so the transfer overlaps the rest of the pass.
//
// Neighbour rule / update / max: src/solver.cpp:29-56 (no NaN holes on this path: the solver falls back to the
// masked colour kernels when D contains NaN).
#include <cuda_pipeline_primitives.h>

#include <cstdlib>
#include <cstring>

#include "sor_common.cuh"
#include "tma.cuh"

namespace pcd {

#ifndef PCD_WAVE_NT
#define PCD_WAVE_NT 256
#endif
constexpr int WAVE_NT = PCD_WAVE_NT;      // threads = column pairs per strip window
constexpr int WAVE_CTAS = 512 / WAVE_NT;   // CTAs per SM (128 registers per thread)
constexpr int WAVE_PF = 3;    // rows prefetched ahead by cp.async

template <int TS>
struct WaveCfg {
    static constexpr int NP = 2 * TS;                            // colour phases per pass
    static constexpr int R = ((2 * NP + WAVE_PF + 1) + 1) & ~1;  // ring period (rows), even
    static constexpr int HX = 4;                                 // stale skirt columns on each side of the window: 2*TS are needed;
                                                                 // fixed at the TS = 2 value so that every launch cuts the
                                                                 // grid into the SAME strips (the per-strip peer flags rely on it)
    static constexpr int LXW = 2 * WAVE_NT;                      // window columns
    static constexpr int CORE = LXW - 2 * HX;                    // columns written by the strip
    static constexpr int PITCH = WAVE_NT + 2;                    // smem pitch of one parity row (pad on each side)
};

struct WaveParams {
    const double *phi_in;   // single-pass launches; a PEER launch derives both per pass from peer.buf[]
    double *phi_out;
    const double *D;
    int W, H;            // global grid
    int row_first, rows; // owned global rows [row_first, row_first + rows)
    int grow0;           // global row of local row 0 of the arrays
    int nchunks;         // row chunks (not counting the short last chunk of a PEER pass)
    int top_credit;      // PEER: rows the first chunk gives to the others (it also feeds the upper neighbour's ghost rows)
    SorW w;
    unsigned long long *slots;  // per-sweep max: TS entries per pass
    WavePeer peer;              // persistent multi-pass launch + fused ghost-row exchange (PEER kernels only)
    alignas(64) CUtensorMap dmap;   // DTMA kernels: D in the parity-split layout [row][parity][Kp] as a 2-D tensor (Kp, 2*rows)
};

// what changes from pass to pass inside one launch
struct WaveDyn {
    const double *phi_in;
    double *phi_out;
    double *up_out, *dn_out;     // the neighbouring slabs' output fields (nullptr: no neighbour)
    unsigned long long *slots;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// ---- TMA staging of D (DTMA kernels) -------------------------------------------------------------------------------
// D is read-only and only ever read by the thread that owns the cell, so its layout is free: the solver keeps a copy in
// a column-parity-split layout D_split[row][parity][Kp] (Kp = ceil(W/2) rounded up to even) and one elected thread
// fetches BOTH parity planes of a row's 256 column pairs with ONE bulk tensor copy (cp.async.bulk.tensor.2d, box
// {256, 2}: 4 KB, zero-filled outside the array, completion on an mbarrier) instead of two 8-byte cp.async per thread.
// one thread: wait until *flag has reached `want` (sequence numbers only grow).  SYS: the flag is written by another
// GPU.  A writer that never shows up must not hang the device: after ~3 s the slab's error word is set instead.
template <bool SYS>
__device__ __forceinline__ void seq_wait(const unsigned *flag, unsigned want, int *err) {
    const long long t0 = clock64();
    while ((int)((SYS ? ld_acquire_sys(flag) : ld_acquire_gpu(flag)) - want) < 0) {
        if (clock64() - t0 > 6000000000ll) {
            atomicExch(err, 1);
            break;
        }
        __nanosleep(SYS ? 64 : 20);
    }
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// all threads of the CTA: the edge rows of this CTA's strip are stored in the neighbour's field; raise its flag there
__device__ __forceinline__ void strip_signal(unsigned *flag, unsigned seq) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(flag, seq);
    }
}

struct WaveThread {  // per-thread invariants
    int k, gx0;
    bool ex0, ex1, core;
    double cx[2], wx[2];
    int ys, ye, y0, y1;
    int k0, f_last;   // DTMA: first column pair of the window (tensor coordinate, may be negative); last step of the chunk
};

// One row step.  C = (f - ys) mod R is compile-time; chunks start so that the active cell of row r in phase ph
// has window parity (r - ys + 1 + ph) & 1, i.e. (C + ph) & 1 for the row f-1-2ph updated at offset C.
// MODE (per block of R steps, CTA-uniform):
//   WAVE_STEADY   every row any phase touches is owned by the chunk and no domain edge is near: no checks at all;
//   WAVE_CHECKED  a block at the top / bottom of a chunk.  Rows outside the dependency pyramid of the owned rows are
//                 simply updated as well -- their (stale) values are never read by a row inside the pyramid of a later
//                 phase, so they cannot reach an owned value -- and only what leaves the CTA is guarded: the per-sweep
//                 maximum and the store to the output field (owned rows only), the rows streamed in (they must exist).
//                 Rows outside the grid stay 0.0 (a missing neighbour reads 0.0) and the domain's first / last row use
//                 their own neighbour count.
constexpr int WAVE_STEADY = 0, WAVE_CHECKED = 1;

template <int TS, int C, int MODE, bool DTMA>
__device__ __forceinline__ void wave_step(const WaveParams &p, const WaveDyn &d, const WaveThread &t, const int f,
                                          double (&v)[WaveCfg<TS>::R][2], double *__restrict__ sphi,
                                          double *__restrict__ sD, double (&lmax)[TS], unsigned long long *dbar, unsigned &dpar) {
    using Cfg = WaveCfg<TS>;
    constexpr int R = Cfg::R, NP = Cfg::NP, PITCH = Cfg::PITCH;
    constexpr int DP = DTMA ? WAVE_NT : PITCH, DO = DTMA ? 0 : 1;   // pitch / first-column offset of a parity row of D
    const int W = p.W, H = p.H, k = t.k;
    // (a) stream row f+PF into ring slot (C+PF)%R, 8-byte cp.async per cell.  Columns outside the grid are never copied
    // and never updated in STEADY blocks: their slots keep the 0.0 the ring was initialised with; CHECKED blocks zero-fill
    // rows that lie outside the grid or behind the chunk's last needed row.
    {
        constexpr int SL = (C + WAVE_PF) % R;
        const int fr = f + WAVE_PF;
        if constexpr (MODE == WAVE_CHECKED) {
            const bool row_ok = fr <= t.ye && fr >= 0 && fr < H;
            const size_t base = row_ok ? (size_t)(fr - p.grow0) * W : 0;
            const bool a0 = row_ok && t.ex0, a1 = row_ok && t.ex1;
            const size_t o0 = a0 ? base + t.gx0 : 0, o1 = a1 ? base + t.gx0 + 1 : 0;
            __pipeline_memcpy_async(sphi + (SL * 2 + 0) * PITCH + 1 + k, d.phi_in + o0, 8, a0 ? 0 : 8);
            __pipeline_memcpy_async(sphi + (SL * 2 + 1) * PITCH + 1 + k, d.phi_in + o1, 8, a1 ? 0 : 8);
            if constexpr (!DTMA) {
                __pipeline_memcpy_async(sD + (SL * 2 + 0) * PITCH + 1 + k, p.D + o0, 8, a0 ? 0 : 8);
                __pipeline_memcpy_async(sD + (SL * 2 + 1) * PITCH + 1 + k, p.D + o1, 8, a1 ? 0 : 8);
            }
        } else {
            const size_t o = (size_t)(fr - p.grow0) * W + t.gx0;
            if (t.ex0) {
                __pipeline_memcpy_async(sphi + (SL * 2 + 0) * PITCH + 1 + k, d.phi_in + o, 8);
                if constexpr (!DTMA) __pipeline_memcpy_async(sD + (SL * 2 + 0) * PITCH + 1 + k, p.D + o, 8);
            }
            if (t.ex1) {
                __pipeline_memcpy_async(sphi + (SL * 2 + 1) * PITCH + 1 + k, d.phi_in + o + 1, 8);
                if constexpr (!DTMA) __pipeline_memcpy_async(sD + (SL * 2 + 1) * PITCH + 1 + k, p.D + o + 1, 8);
            }
        }
        __pipeline_commit();
        if constexpr (DTMA) {
            // both parity planes of D's row fr for the window's 256 column pairs: ONE bulk tensor copy by one thread,
            // zero-filled outside the array.  Only rows that a later step waits for are requested (<= f_last + 1), so
            // every mbarrier phase that completes is consumed exactly once.
            if (k == 0 && (MODE == WAVE_STEADY || fr <= t.f_last + 1)) {
                mbar_expect(dbar + SL, 2 * WAVE_NT * (unsigned)sizeof(double));
                tma_load_2d(sD + SL * 2 * WAVE_NT, &p.dmap, t.k0, 2 * (fr - p.grow0), dbar + SL);
            }
        }
    }
    // (b) row f becomes active: own pair from the smem ring into the register ring
    v[C][0] = sphi[(C * 2 + 0) * PITCH + 1 + k];
    v[C][1] = sphi[(C * 2 + 1) * PITCH + 1 + k];
    // (c) phases: phase ph updates row f - 1 - 2*ph (row f, just activated, is phase 0's lower neighbour).
    // The NP updates of a step are independent of each other: all shared-memory reads are issued first, then
    // the arithmetic, then all writes, so the NP dependency chains overlap instead of being serialised by the
    // (possible-alias) ordering of smem stores and loads.
    double nbv[NP], Dvv[NP], nvv[NP];
    bool act[NP];
#pragma unroll
    for (int ph = 0; ph < NP; ++ph) {
        const int SL = (C - 1 - 2 * ph + 2 * R) % R;
        const int q = (C + ph) & 1;  // compile-time after unrolling
        const int r = f - 1 - 2 * ph;
        bool a = q ? t.ex1 : t.ex0;                           // columns outside the grid stay 0.0
        if (MODE == WAVE_CHECKED) a = a && r >= 0 && r < H;   // ... and so do rows outside the grid
        act[ph] = a;
        nbv[ph] = q ? sphi[(SL * 2 + 0) * PITCH + 1 + k + 1] : sphi[(SL * 2 + 1) * PITCH + 1 + k - 1];
        Dvv[ph] = sD[(SL * 2 + q) * DP + DO + k];
    }
#pragma unroll
    for (int ph = 0; ph < NP; ++ph) {
        const int SL = (C - 1 - 2 * ph + 2 * R) % R, SU = (SL + R - 1) % R, SD = (SL + 1) % R;
        const int q = (C + ph) & 1;
        const int r = f - 1 - 2 * ph;
        const double own = v[SL][q ^ 1];
        const double l = q ? own : nbv[ph], rr = q ? nbv[ph] : own;
        const double u = v[SU][q], dn = v[SD][q];
        const double val = v[SL][q];
        const double sum = ((l + u) + rr) + dn;  // ghost rows / columns read 0.0: identical to skipping them
        double delta;
        if (MODE == WAVE_CHECKED && (r == 0 || r == H - 1)) {  // CTA-uniform: domain top / bottom row (two rows of the grid)
            const int cnt = (int)t.cx[q] - (r == 0 ? 1 : 0) - (r == H - 1 ? 1 : 0);
            delta = wsel(p.w, cnt) * ((sum - (double)cnt * val) - Dvv[ph]);
        } else {
            delta = t.wx[q] * ((sum - t.cx[q] * val) - Dvv[ph]);
        }
        const double nv = act[ph] ? val + delta : val;
        nvv[ph] = nv;
        v[SL][q] = nv;
        if (act[ph] && t.core && (MODE == WAVE_STEADY || (r >= t.y0 && r < t.y1))) {
            const double ad = fabs(delta);
            if (ad > lmax[ph / 2]) lmax[ph / 2] = ad;
        }
    }
#pragma unroll
    for (int ph = 0; ph < NP; ++ph) {
        const int SL = (C - 1 - 2 * ph + 2 * R) % R;
        const int q = (C + ph) & 1;
        if (act[ph]) sphi[(SL * 2 + q) * PITCH + 1 + k] = nvv[ph];
    }
    // (d) the row that just finished its last phase leaves through the output field
    {
        constexpr int SL = (C - 1 - 2 * (NP - 1) + 2 * R) % R;
        const int r = f - 1 - 2 * (NP - 1);
        if (t.core && (MODE == WAVE_STEADY || (r >= t.y0 && r < t.y1))) {
            const size_t o = (size_t)(r - p.grow0) * W + t.gx0;
            if (t.ex0 && t.ex1 && ((W & 1) == 0)) {
                *reinterpret_cast<double2 *>(d.phi_out + o) = make_double2(v[SL][0], v[SL][1]);
            } else {
                if (t.ex0) d.phi_out[o] = v[SL][0];
                if (t.ex1) d.phi_out[o + 1] = v[SL][1];
            }
        }
    }
    // (e) row f+1 must have landed before the next step reads it
    __pipeline_wait_prior(WAVE_PF - 1);
    if constexpr (DTMA) {
        constexpr int S1 = (C + 1) % R;
        mbar_wait(dbar + S1, (dpar >> S1) & 1u, p.peer.err);
        dpar ^= 1u << S1;
    }
    __syncthreads();
}

template <int TS, int C, int MODE, bool DTMA>
struct WaveUnroll {
    static __device__ __forceinline__ void run(const WaveParams &p, const WaveDyn &d, const WaveThread &t, const int f,
                                               const int f_last, double (&v)[WaveCfg<TS>::R][2], double *sphi, double *sD,
                                               double (&lmax)[TS], unsigned long long *dbar, unsigned &dpar) {
        if (MODE != WAVE_STEADY && f + C > f_last) return;  // CTA-uniform: the chunk's last block stops at its last step
        wave_step<TS, C, MODE, DTMA>(p, d, t, f + C, v, sphi, sD, lmax, dbar, dpar);
        if constexpr (C + 1 < WaveCfg<TS>::R) WaveUnroll<TS, C + 1, MODE, DTMA>::run(p, d, t, f, f_last, v, sphi, sD, lmax, dbar, dpar);
    }
};

// Ghost rows of a neighbouring slab: every thread copies ITS OWN two columns of `nrows` finished rows of this pass's
// output field (it stored them itself a few steps ago) into the neighbour's output field -- plain stores to peer memory
// over NVLink.  Kept out of the row step so that the hot loop carries neither the code nor the registers for it.
template <int NROWS>
__device__ __forceinline__ void push_rows(const WaveParams &p, const WaveDyn &d, const WaveThread &t, double *dst_base,
                                          const int dst_grow0, const int first_row) {
    if (!t.core || !(t.ex0 || t.ex1)) return;
    const double *src = d.phi_out + (size_t)(first_row - p.grow0) * p.W + t.gx0;
    double *dst = dst_base + (size_t)(first_row - dst_grow0) * p.W + t.gx0;
    if (t.ex0 && t.ex1 && ((p.W & 1) == 0)) {      // all loads first (L2 round trips overlap), then all peer stores
        double2 v[NROWS];
#pragma unroll
        for (int r = 0; r < NROWS; ++r) v[r] = __ldcg(reinterpret_cast<const double2 *>(src + (size_t)r * p.W));
#pragma unroll
        for (int r = 0; r < NROWS; ++r) *reinterpret_cast<double2 *>(dst + (size_t)r * p.W) = v[r];
    } else {
#pragma unroll
        for (int r = 0; r < NROWS; ++r) {
            if (t.ex0) dst[(size_t)r * p.W] = __ldcg(src + (size_t)r * p.W);
            if (t.ex1) dst[(size_t)r * p.W + 1] = __ldcg(src + (size_t)r * p.W + 1);
        }
    }
}

// DTMA: D of the first WAVE_PF rows of a chunk (ring slots 0..PF-1).  D never changes during a solve, so a persistent
// launch requests them BEFORE it waits for its neighbours: the copy's latency hides behind the pass boundary.
__device__ __forceinline__ void dtma_prologue(const WaveParams &p, const WaveThread &t, double *sD, unsigned long long *dbar) {
    if (t.k != 0) return;
#pragma unroll
    for (int j = 0; j < WAVE_PF; ++j) {
        const int fr = t.ys + j;
        if (fr <= t.f_last + 1) {
            mbar_expect(dbar + j, 2 * WAVE_NT * (unsigned)sizeof(double));
            tma_load_2d(sD + j * 2 * WAVE_NT, &p.dmap, t.k0, 2 * (fr - p.grow0), dbar + j);
        }
    }
}

// One pass of one CTA over its chunk.  top / bot: this chunk touches the slab's first / last owned rows AND a
// neighbouring slab is attached there (its ghost rows are fed from here, flag value `seq`).
template <int TS, bool PEER, bool DTMA>
__device__ __forceinline__ void wave_chunk(const WaveParams &p, const WaveDyn &d, const WaveThread &t, double *sphi, double *sD,
                                           double (*wred)[WAVE_NT / 32], const bool top, const bool bot, const unsigned seq,
                                           unsigned long long *dbar, unsigned &dpar) {
    using Cfg = WaveCfg<TS>;
    constexpr int R = Cfg::R, NP = Cfg::NP, PITCH = Cfg::PITCH;
    const int H = p.H;
    double v[R][2];
#pragma unroll
    for (int i = 0; i < R; ++i) { v[i][0] = 0.0; v[i][1] = 0.0; }
    double lmax[TS];
#pragma unroll
    for (int s = 0; s < TS; ++s) lmax[s] = 0.0;

    // prologue: rows ys .. ys+PF-1 (slots 0..PF-1)
#pragma unroll
    for (int j = 0; j < WAVE_PF; ++j) {
        const int fr = t.ys + j;
        const bool row_ok = fr <= t.ye && fr >= 0 && fr < H;
        const size_t base = row_ok ? (size_t)(fr - p.grow0) * p.W : 0;
        const bool a0 = row_ok && t.ex0, a1 = row_ok && t.ex1;
        const size_t o0 = a0 ? base + t.gx0 : 0, o1 = a1 ? base + t.gx0 + 1 : 0;
        __pipeline_memcpy_async(sphi + (j * 2 + 0) * PITCH + 1 + t.k, d.phi_in + o0, 8, a0 ? 0 : 8);
        __pipeline_memcpy_async(sphi + (j * 2 + 1) * PITCH + 1 + t.k, d.phi_in + o1, 8, a1 ? 0 : 8);
        if constexpr (!DTMA) {
            __pipeline_memcpy_async(sD + (j * 2 + 0) * PITCH + 1 + t.k, p.D + o0, 8, a0 ? 0 : 8);
            __pipeline_memcpy_async(sD + (j * 2 + 1) * PITCH + 1 + t.k, p.D + o1, 8, a1 ? 0 : 8);
        }
        __pipeline_commit();
    }
    if constexpr (DTMA && !PEER) dtma_prologue(p, t, sD, dbar);   // (a persistent launch requested these rows before its waits)
    __pipeline_wait_prior(WAVE_PF - 1);
    if constexpr (DTMA) {   // D of row ys (slot 0)
        mbar_wait(dbar + 0, dpar & 1u, p.peer.err);
        dpar ^= 1u;
    }
    __syncthreads();

    const int f_last = t.y1 + 2 * (NP - 1);
    bool up_sent = false, up_pushed = false;
    for (int fb = t.ys; fb <= f_last; fb += R) {
        // a block of R steps is steady when every row any phase touches is owned and not a domain edge row, and
        // every row streamed in exists
        const int r_min = fb - 1 - 2 * (NP - 1), r_max = fb + R - 2, f_max = fb + R - 1 + WAVE_PF;
        const bool steady = r_min >= max(t.y0, 1) && r_max <= min(t.y1 - 1, H - 2) && f_max <= min(t.ye, H - 1);
        if (steady) WaveUnroll<TS, 0, WAVE_STEADY, DTMA>::run(p, d, t, fb, f_last, v, sphi, sD, lmax, dbar, dpar);
        else WaveUnroll<TS, 0, WAVE_CHECKED, DTMA>::run(p, d, t, fb, f_last, v, sphi, sD, lmax, dbar, dpar);
        if constexpr (PEER) {
            // The top edge rows are complete long before the chunk is: copy them into the upper neighbour's ghost rows
            // now, and raise its flag ONE BLOCK LATER, when the peer stores have long been acknowledged -- the
            // system-scope fence in front of the flag then costs nothing (this CTA is the critical one of the pass).
            if (top && up_pushed && !up_sent) {
                strip_signal(p.peer.sig_up + blockIdx.x, seq);
                up_sent = true;
            }
            if (top && !up_pushed && r_min + R - 1 >= p.row_first + p.peer.gh - 1) {
                push_rows<2 * TS + 1>(p, d, t, d.up_out, p.peer.up_grow0, p.row_first);
                up_pushed = true;
            }
        }
    }
    if constexpr (PEER) {
        if (top && !up_pushed) push_rows<2 * TS + 1>(p, d, t, d.up_out, p.peer.up_grow0, p.row_first);
        if (top && !up_sent) strip_signal(p.peer.sig_up + blockIdx.x, seq);
        if (bot) {
            push_rows<2 * TS + 1>(p, d, t, d.dn_out, p.peer.dn_grow0, p.row_first + p.rows - p.peer.gh);
            strip_signal(p.peer.sig_dn + blockIdx.x, seq);
        }
    }

    // publish the per-sweep maxima
    __pipeline_wait_prior(0);
#pragma unroll
    for (int s = 0; s < TS; ++s) {
        const double m = warp_max(lmax[s]);
        if ((t.k & 31) == 0) wred[s][t.k >> 5] = m;
    }
    __syncthreads();
    if (t.k < 32) {
#pragma unroll
        for (int s = 0; s < TS; ++s) {
            double m = t.k < WAVE_NT / 32 ? wred[s][t.k] : 0.0;
            m = warp_max(m);
            if (t.k == 0 && m > 0.0) atomicMax(d.slots + s, (unsigned long long)__double_as_longlong(m));
        }
    }
}

template <int TS, bool PEER, bool DTMA>
__global__ void __launch_bounds__(WAVE_NT, WAVE_CTAS) sor_wave_kernel(const __grid_constant__ WaveParams p) {
    using Cfg = WaveCfg<TS>;
    constexpr int R = Cfg::R, NP = Cfg::NP, PITCH = Cfg::PITCH;
    constexpr int SD_ELEMS = DTMA ? R * 2 * WAVE_NT : R * 2 * PITCH;
    extern __shared__ __align__(128) double smem[];
    double *sphi = smem;                   // [R][2][PITCH]
    double *sD = smem + R * 2 * PITCH;     // [R][2][PITCH], DTMA: [R][2][WAVE_NT] dense (128-byte aligned: R*2*PITCH*8 is)
    __shared__ double wred[TS][WAVE_NT / 32];
    __shared__ unsigned long long dbar[R];   // DTMA: one mbarrier per ring slot of D
    static_assert(!DTMA || (R * 2 * PITCH * 8) % 128 == 0, "the D ring must start on a 128-byte boundary");
    unsigned dpar = 0u;                      // DTMA: phase parity of every slot's mbarrier (bit per slot)
    if constexpr (DTMA) {
        if (threadIdx.x == 0) {
            for (int i = 0; i < R; ++i) mbar_init(dbar + i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&p.dmap) : "memory");
        }
    }

    WaveThread t;
    t.k = threadIdx.x;
    const int W = p.W;
    const int xw0 = blockIdx.x * Cfg::CORE - Cfg::HX;              // window origin (global x, even)
    if (PEER && p.peer.tail_rows > 0 && blockIdx.y == gridDim.y - 1) {  // short last chunk: the bottom edge rows leave early
        t.y1 = p.row_first + p.rows;
        t.y0 = t.y1 - p.peer.tail_rows;
    } else {
        const int main_rows = p.rows - (PEER ? p.peer.tail_rows : 0);   // balanced split: chunk lengths differ by <= 1 row
        // ... except that the first chunk of a slab with an upper neighbour is `top_credit` rows shorter: it spends that time
        // copying its first rows into the neighbour and raising the flag, and a pass moves at the pace of its slowest CTA
        const int c = (int)blockIdx.y, n = p.nchunks, cr = (PEER && n > 1) ? p.top_credit : 0;
        auto cut = [&](int i) { return i == 0 ? 0 : (int)((long long)i * (main_rows + cr) / n) - cr; };
        t.y0 = p.row_first + cut(c);
        t.y1 = p.row_first + (c + 1 == n ? main_rows : cut(c + 1));
    }
    const bool empty = t.y0 >= t.y1;
    if (!PEER && empty) return;
    // first row streamed in: NP rows of warm-up, moved one row earlier when needed so that the colour parity of
    // every (step offset, phase) pair is a compile-time constant: active parity of row r in phase ph is
    // (xw0 + r + ph) & 1 = (r + ph) & 1, and r = ys + C - 1 - 2ph at offset C  =>  need ys odd
    t.ys = t.y0 - NP;
    if ((t.ys & 1) == 0) t.ys -= 1;
    t.ye = t.y1 - 1 + NP;                                           // last row streamed in
    t.gx0 = xw0 + 2 * t.k;
    t.ex0 = t.gx0 >= 0 && t.gx0 < W;
    t.ex1 = t.gx0 + 1 >= 0 && t.gx0 + 1 < W;
    t.core = (2 * t.k >= Cfg::HX) && (2 * t.k < Cfg::LXW - Cfg::HX);
    t.k0 = xw0 / 2;                                                 // xw0 is even (and may be -HX)
    t.f_last = t.y1 + 2 * (NP - 1);
    {
        const int c0 = 4 - (t.gx0 == 0 ? 1 : 0) - (t.gx0 == W - 1 ? 1 : 0), c1 = 4 - (t.gx0 + 1 == 0 ? 1 : 0) - (t.gx0 + 1 == W - 1 ? 1 : 0);
        t.cx[0] = (double)c0; t.cx[1] = (double)c1;
        t.wx[0] = wsel(p.w, c0); t.wx[1] = wsel(p.w, c1);
    }
    // the ring is zeroed once: a pass overwrites every slot it reads for a live result, and the two pad columns of a
    // parity row are never written
    for (int i = t.k; i < R * 2 * PITCH + SD_ELEMS; i += WAVE_NT) smem[i] = 0.0;
    if constexpr (DTMA) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores above, bulk copies into the same bytes below
    __syncthreads();

    if constexpr (!PEER) {
        WaveDyn d;
        d.phi_in = p.phi_in; d.phi_out = p.phi_out; d.up_out = nullptr; d.dn_out = nullptr; d.slots = p.slots;
        wave_chunk<TS, false, DTMA>(p, d, t, sphi, sD, wred, false, false, 0u, dbar, dpar);
    } else {
        // ---- persistent launch: p.peer.npass passes, CTAs synchronise with their neighbours only ----
        const int nbx = (int)gridDim.x, nby = (int)gridDim.y, bx = (int)blockIdx.x, by = (int)blockIdx.y;
        unsigned *const my_done = p.peer.done + by * nbx + bx;
        const bool has_up = p.peer.up_buf[0] != nullptr, has_dn = p.peer.dn_buf[0] != nullptr;
        const bool top = !empty && has_up && t.y0 == p.row_first;
        const bool bot = !empty && has_dn && t.y1 == p.row_first + p.rows;
        for (int i = 0; i < p.peer.npass; ++i) {
            const unsigned seq = p.peer.seq0 + (unsigned)i + 1u;
            const int oid = (p.peer.cur ^ (i & 1)) ^ 1;          // buffer this pass writes
            WaveDyn d;
            d.phi_in = p.peer.buf[oid ^ 1];
            d.phi_out = p.peer.buf[oid];
            d.up_out = has_up ? p.peer.up_buf[oid] : nullptr;
            d.dn_out = has_dn ? p.peer.dn_buf[oid] : nullptr;
            d.slots = p.slots + (size_t)i * TS;
            if constexpr (DTMA) {
                if (!empty) dtma_prologue(p, t, sD, dbar);   // every warp is past its last read of the D ring (barrier below / kernel start)
            }
            unsigned long long *tr = nullptr;
            if (p.peer.trace && threadIdx.x == 0) {
                tr = p.peer.trace + ((size_t)i * WAVE_MAX_CTAS + (by * nbx + bx)) * 4;
                tr[0] = global_ns();
            }
            // (1) dependencies of this pass.  Inside the GPU: the eight neighbouring CTAs have published pass seq-1 --
            // the only ones whose output this CTA reads and whose input it is about to overwrite (the launch boundary
            // covers the first pass).  Across GPUs: the neighbour's edge CTAs of strips bx-1..bx+1 have raised their
            // flags for pass seq-1: their rows are in this slab's ghost rows, and they no longer read the ghost rows
            // this pass overwrites over there.
            {
                const int tid = (int)threadIdx.x;
                if (i > 0 && tid < 9 && tid != 4) {
                    const int nx = bx + tid % 3 - 1, ny = by + tid / 3 - 1;
                    if (nx >= 0 && nx < nbx && ny >= 0 && ny < nby) seq_wait<false>(p.peer.done + ny * nbx + nx, seq - 1u, p.peer.err);
                } else if (tid >= 32 && tid < 35) {
                    const int nx = bx + tid - 33;
                    if (top && nx >= 0 && nx < nbx) seq_wait<true>(p.peer.wait_up + nx, seq - 1u, p.peer.err);
                } else if (tid >= 64 && tid < 67) {
                    const int nx = bx + tid - 65;
                    if (bot && nx >= 0 && nx < nbx) seq_wait<true>(p.peer.wait_dn + nx, seq - 1u, p.peer.err);
                }
                // a wait that ran into its limit (a neighbour died or never started) voids the whole launch: every CTA
                // of this GPU leaves at its next pass boundary instead of spinning ~3 s per pass (the error word is
                // per GPU; the neighbouring GPUs notice through their own waits)
                const int dead = __syncthreads_or(threadIdx.x == 0 ? *((volatile int *)p.peer.err) : 0);
                if (dead) {
                    if constexpr (DTMA) {   // do not leave with bulk copies in flight into this CTA's shared memory
                        if (!empty)
                            for (int j = 0; j < WAVE_PF; ++j)
                                if (t.ys + j <= t.f_last + 1) mbar_wait(dbar + j, (dpar >> j) & 1u, nullptr);
                    }
                    break;
                }
                __threadfence();   // every thread's loads of this pass are ordered behind the flags observed above
            }
            // (2) the pass
            if (tr) tr[1] = global_ns();
            if (!empty) wave_chunk<TS, true, DTMA>(p, d, t, sphi, sD, wred, top, bot, seq, dbar, dpar);
            // (3) publish: every store of this CTA's pass is visible before its sequence word moves
            __syncthreads();
            if (threadIdx.x == 0) {
                if (tr) tr[2] = global_ns();
                __threadfence();
                st_release_gpu(my_done, seq);
                if (tr) tr[3] = global_ns();
            }
        }
    }
}

constexpr int TILED_TS = 2;
int tiled_sweeps_per_pass() { return TILED_TS; }

template <int TS, bool PEER, bool DTMA>
static int wave_smem_optin(size_t smem) {
    // the opt-in is per device: remember which devices have it (contexts on several threads may race here: atomic)
    static std::atomic<unsigned long long> done_mask{0ull};
    int dev = 0;
    PCD_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !((done_mask.load(std::memory_order_acquire) >> dev) & 1ull)) {
        PCD_CUDA(cudaFuncSetAttribute(sor_wave_kernel<TS, PEER, DTMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev < 64) done_mask.fetch_or(1ull << dev, std::memory_order_release);
    }
    return PCD_OK;
}

template <int TS, bool PEER, bool DTMA = false>
static int launch_wave(const WaveParams &prm, int sm_count, int sm_reserve, cudaStream_t stream) {
    using Cfg = WaveCfg<TS>;
    const size_t smem = (size_t)Cfg::R * 2 * (Cfg::PITCH + (DTMA ? WAVE_NT : Cfg::PITCH)) * sizeof(double);
    PCD_TRY((wave_smem_optin<TS, PEER, DTMA>(smem)));
    WaveParams p = prm;
    const int strips = (p.W + Cfg::CORE - 1) / Cfg::CORE;
    // ONE wave of two CTAs per SM (never a second, nearly empty wave).  Measured on B200 (2048^2: 59 chunks of 35
    // rows 19.8 us/sweep, 32 chunks of 64 rows 28.6): the row step is latency-bound, so the shortest chunks that
    // still fit one wave win even though the 2*NP warm-up rows are then a larger share of the work
    // (multi-GPU runs keep a few SMs free so that the NCCL kernels of the overlapped ghost-row exchange can run)
    const int avail = (sm_count - sm_reserve >= 8) ? sm_count - sm_reserve : sm_count;
    int chunks = (WAVE_CTAS * avail) / strips;
    int min_rows = 4 * Cfg::NP;  // a step costs the same latency whatever the chunk length: fill the wave first
    static const int dbg_chunks = getenv("PCD_WAVE_CHUNKS") ? atoi(getenv("PCD_WAVE_CHUNKS")) : 0;  // tuning knob
    if (dbg_chunks > 0) { chunks = dbg_chunks; min_rows = 8; }
    int tail = 0;
    if constexpr (PEER) {
        // with a lower neighbour the bottom rows get a short chunk of their own (they are the LAST rows a chunk
        // walking down would finish), so that the neighbour has them long before the pass ends
        // (its CTAs would otherwise idle for most of the pass: 32 rows where the slab is thick enough, measured -1.3 %)
        static const int tail_env = getenv("PCD_WAVE_TAIL_ROWS") ? atoi(getenv("PCD_WAVE_TAIL_ROWS")) : 0;     // tuning knob
        if (p.peer.dn_buf[0] && p.rows >= 8 * min_rows && chunks >= 3) {
            tail = tail_env >= 2 * Cfg::NP + 1 ? tail_env : (p.rows >= 16 * 32 ? 32 : 2 * Cfg::NP + 4);
            chunks -= 1;
        }
        p.peer.tail_rows = tail;
    }
    // The first chunk of a slab with an upper neighbour also copies its first rows into the neighbour and raises the flag:
    // ~8 us per pass, and a pass moves at the pace of its slowest CTA (profiles/r02_wave_trace_*: 52 us against 42-46 for
    // the other chunks).  It gets 10 rows fewer, the others share them: 8192 x 1024 slabs on 4 GPUs 30.6 -> 28.8 us/sweep.
    static const int credit_env = getenv("PCD_WAVE_TOP_CREDIT") ? atoi(getenv("PCD_WAVE_TOP_CREDIT")) : 10;   // tuning knob
    p.top_credit = 0;
    if (PEER && p.peer.up_buf[0]) p.top_credit = credit_env;
    const int main_rows = p.rows - tail;
    if (chunks * min_rows > main_rows) chunks = main_rows / min_rows;
    if (chunks < 1) chunks = 1;
    p.nchunks = chunks;
    if (chunks < 2 || p.top_credit < 0 || (main_rows + p.top_credit) / chunks - p.top_credit < 2 * min_rows) p.top_credit = 0;
    const dim3 grid(strips, chunks + (tail ? 1 : 0));
    if constexpr (PEER) {
        // persistent: every CTA must be resident (they wait for each other)
        if ((int)(grid.x * grid.y) > WAVE_CTAS * sm_count || (int)(grid.x * grid.y) > WAVE_MAX_CTAS) {
            set_error("wavefront kernel: %u x %u CTAs cannot be co-resident on %d SMs", grid.x, grid.y, sm_count);
            return PCD_ERR_UNSUPPORTED;
        }
        void *args[] = {&p};
        PCD_CUDA(cudaLaunchCooperativeKernel((void *)sor_wave_kernel<TS, true, DTMA>, grid, dim3(WAVE_NT), args, smem, stream));
    } else {
        sor_wave_kernel<TS, false, false><<<grid, WAVE_NT, smem, stream>>>(p);
    }
    PCD_LAUNCHED();
    return PCD_OK;
}

// One pass over global rows [row_first, row_first+rows): phi_out <- nsweeps (<= TILED_TS) sweeps applied to
// phi_in.  Arrays have local row 0 = global row grow0 and must hold rows row_first-2*nsweeps-1 ..
// row_first+rows+2*nsweeps-1 (clipped to the grid) of the current field.
int tiled_pass(const double *phi_in, double *phi_out, const double *D, int W, int H, int row_first, int rows, int grow0,
               int nsweeps, unsigned long long *slots, int sm_count, int sm_reserve, cudaStream_t stream) {
    WaveParams prm;
    prm.phi_in = phi_in; prm.phi_out = phi_out; prm.D = D; prm.W = W; prm.H = H;
    prm.row_first = row_first; prm.rows = rows; prm.grow0 = grow0; prm.nchunks = 1; prm.top_credit = 0;
    prm.w = make_w(W); prm.slots = slots;
    if (nsweeps >= 2) return launch_wave<2, false>(prm, sm_count, sm_reserve, stream);
    return launch_wave<1, false>(prm, sm_count, sm_reserve, stream);
}

// ONE persistent launch: peer.npass passes of `sweeps_per_pass` (1 or TILED_TS) sweeps each over a whole slab, the field
// ping-ponging between peer.buf[0] and peer.buf[1] (peer.cur holds it first), maxima into slots[0 .. npass *
// sweeps_per_pass), ghost rows pushed into the attached neighbours (see WavePeer).
int tiled_run_peer(const double *D, int W, int H, int row_first, int rows, int grow0, int sweeps_per_pass,
                   unsigned long long *slots, const WavePeer &peer, const void *dmap, int sm_count, int sm_reserve,
                   cudaStream_t stream) {
    WaveParams prm;
    prm.phi_in = nullptr; prm.phi_out = nullptr; prm.D = D; prm.W = W; prm.H = H;
    prm.row_first = row_first; prm.rows = rows; prm.grow0 = grow0; prm.nchunks = 1; prm.top_credit = 0;
    prm.w = make_w(W); prm.slots = slots; prm.peer = peer;
    // D staged by TMA from its parity-split copy -- where it pays.  Measured on B200 (tools/wave_time.py, us per sweep,
    // TMA / cp.async): 8192^2 (482 rows per chunk) 153.0 / 156.1, 4096^2 (113) 47.3 / 45.4, 8192 x 1024 (60) 30.2 / 26.4,
    // 2048^2 (35) 21.9 / 18.0: the bulk copy saves ~50 ns per row step and costs ~9 us per pass (its waits are polled,
    // the cp.async ones are scoreboard waits), so it wins from roughly 190 rows per chunk on.
    bool use_tma = dmap != nullptr;
    if (use_tma) {
        using Cfg = WaveCfg<TILED_TS>;
        const int strips = (W + Cfg::CORE - 1) / Cfg::CORE;
        int chunks = (WAVE_CTAS * sm_count) / strips;
        if (chunks * 4 * Cfg::NP > rows) chunks = rows / (4 * Cfg::NP);
        if (chunks < 1) chunks = 1;
        static const int min_rows = getenv("PCD_WAVE_TMA_MIN_ROWS") ? atoi(getenv("PCD_WAVE_TMA_MIN_ROWS")) : 192;   // tuning knob
        use_tma = rows / chunks >= min_rows;
    }
    if (use_tma) {
        memcpy(&prm.dmap, dmap, sizeof(CUtensorMap));
        if (sweeps_per_pass >= 2) return launch_wave<2, true, true>(prm, sm_count, sm_reserve, stream);
        return launch_wave<1, true, true>(prm, sm_count, sm_reserve, stream);
    }
    memset(&prm.dmap, 0, sizeof(CUtensorMap));
    if (sweeps_per_pass >= 2) return launch_wave<2, true>(prm, sm_count, sm_reserve, stream);
    return launch_wave<1, true>(prm, sm_count, sm_reserve, stream);
}

// ---- parity-split copy of D for the TMA staging ------------------------------------------------------------------
int tiled_dsplit_pitch(int W) { return (((W + 1) / 2) + 1) & ~1; }   // Kp: column pairs per row, rounded up to even

__global__ void dsplit_kernel(const double *__restrict__ D, double *__restrict__ S, int W, int rows, int Kp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (k >= Kp) return;
    const size_t i = (size_t)r * W + 2 * k;
    S[((size_t)r * 2 + 0) * Kp + k] = 2 * k < W ? D[i] : 0.0;
    S[((size_t)r * 2 + 1) * Kp + k] = 2 * k + 1 < W ? D[i + 1] : 0.0;
}

// S[rows][2][Kp] <- D[rows][W] (on `stream`)
int tiled_dsplit(const double *D, double *S, int W, int rows, cudaStream_t stream) {
    const int Kp = tiled_dsplit_pitch(W);
    dsplit_kernel<<<dim3((Kp + 255) / 256, rows), 256, 0, stream>>>(D, S, W, rows, Kp);
    PCD_LAUNCHED();
    return PCD_OK;
}

// The tensor map of a parity-split D: 2-D (Kp, 2 * rows) fp64, box {WAVE_NT, 2}.  `map_out`: 128 bytes, 64-byte aligned
// not required here (it is copied into the kernel parameters).  Returns PCD_ERR_UNSUPPORTED when the driver entry
// point is missing or refuses the shape: the caller then runs the cp.async staging.
int tiled_dmap_encode(void *map_out, const double *S, int W, int rows) {
    static const bool off = getenv("PCD_WAVE_NO_TMA") != nullptr;   // diagnostics: force the cp.async staging
    if (off || WAVE_NT > 256) return PCD_ERR_UNSUPPORTED;
    const int Kp = tiled_dsplit_pitch(W);
    return tma_encode_2d_f64(map_out, S, (unsigned long long)Kp, (unsigned long long)2 * rows, (unsigned long long)Kp * sizeof(double), WAVE_NT, 2);
}

int tiled_strips(int W) { return (W + WaveCfg<TILED_TS>::CORE - 1) / WaveCfg<TILED_TS>::CORE; }

}  // namespace pcd
