// K-SOR, resident path: ONE persistent kernel per Poisson solve (sm_100a), one CTA per SM, all CTAs co-resident.
//
// Decomposition: the H rows are split as evenly as possible over P = ceil(H/NR) CTAs (slabs of NR and NR-1
// rows); thread k of a CTA's 512 threads owns the column pair (2k, 2k+1) of every row of the slab -- a vertical
// strip of 2*NR cells whose phi and D values stay in REGISTERS for the whole solve.
//   * up / down neighbours of a cell are the thread's own registers (compile-time indices);
//   * the left / right neighbour that belongs to thread k-1 / k+1 goes through shared memory
//     (one STS + one LDS per cell update, unit stride, conflict-free);
//   * the rows above / below the slab belong to the neighbouring CTAs: each boundary cell is pushed to
//     the neighbour as one 16-byte flag-in-data message {lo, seq, hi, seq} ("LL" protocol: each 8-byte
//     half is written atomically and carries the phase number, so the reader needs no fence and no
//     separate flag) and is polled by exactly the thread that consumes it -> registers.  The kernel runs as
//     clusters of two CTAs: the link inside a pair uses slots in the partner's shared memory (DSMEM), the
//     other link slots in global memory (L2).  There is no grid-wide barrier; CTAs only wait for their
//     two neighbours;
//   * inside a CTA there is no CTA-wide barrier in the sweep loop either: after a phase a warp syncs with
//     its two neighbouring warps only (pairwise named barriers);
//   * convergence: per sweep every warp adds its verdict to the sweep's arrival word in shared memory, the
//     last warp adds (1 | not_converged<<32) to that sweep's global slot with ONE atomic (arrival count and
//     verdict travel together, no fence); the slot of sweep s-lag is read at the end of sweep s by thread
//     0 of every CTA and the stop is announced RES_STOP_AHEAD sweeps ahead, so all warps of all CTAs leave
//     after the same sweep.
// Domain edges cost nothing in the hot loop: a missing neighbour reads a 0.0 ghost (x + 0.0 == x, so the
// reference's "skip the neighbour" sum is reproduced bit for bit) and the per-cell neighbour count /
// omega/cnt factors are per-thread registers selected at compile time by (row class, column parity).
// Threads whose strip touches NaN holes or phantom cells (odd W) run a per-cell masked path.
//
// Two kernels share this layout.  `sor_resident_kernel` (first half of the file) exchanges halos once per COLOUR PHASE and
// handles everything: odd widths, NaN holes, slabs of one or two rows.  `sor_resident_deep_kernel` (second half; the one
// that normally runs) exchanges once per SWEEP by updating the colour-0 cells of the rows just outside its slab
// redundantly, keeps D in shared memory (up to 9 rows per CTA), and also solves grids wider than 1024 columns on the
// transposed grid; it needs an even width and no NaN holes.  resident_plan / run_resident at the end choose between them.
//
// Update formula, ordering and stopping rule: src/solver.cpp:12-61,70-147 (see sor_kernels.cu header).
#include <cooperative_groups.h>

#include <cstdlib>
#include <type_traits>

#include "sor_common.cuh"

namespace pcd {

constexpr int RES_NT = 512;     // 16 warps = 4 per SM sub-partition -> 128 registers per thread
constexpr int RES_NR_MAX = 7;   // rows per slab (2*NR phi + 2*NR D doubles per thread)
// The deep-halo kernel keeps D in shared memory, so its strips can be taller: 9 rows = 36 phi registers per thread and
// 209 KB of shared memory per CTA (grids up to 1024 x 1332; the exchange-per-phase kernel stops at RES_NR_MAX).
constexpr int RES_NR_DEEP_MAX = 9;
constexpr int RES_KP = RES_NT + 4;  // smem pitch of one parity row: compile-time so every smem offset is an immediate
// Halo slots of one link and direction: [parity][RES_NT] x 16 B, the message of column x = 2k+q in slot q * RES_NT + k.
// Dense per parity: the 32 messages a warp sends in a phase fill 16 whole 32-byte sectors (with one slot per COLUMN, as
// in round 1, every message half-filled a sector of its own: twice the L2 sector traffic and partial-sector writes).
constexpr int RES_SLOTS = 2 * RES_NT;

struct ResParams {
    double *phi;              // global field, in/out
    const double *D;
    int W, H, K, Kp;          // K = ceil(W/2) column pairs per row, Kp = padded pitch of one parity row in smem
    int P;                    // CTAs
    int n_big;                // CTAs [0, n_big) own NR rows, the others NR-1 (rows split as evenly as possible)
    int nr_big;               // NR of the big slabs
    int max_it;               // sweeps this launch may execute
    int lag;                  // convergence lag
    double tol;
    SorW w;
    uint4 *ll;                // LL halo slots: [P][2][RES_SLOTS] x 16 B  ([.][0] = from the CTA above, [.][1] = from below)
    unsigned long long *g_max;   // [max_it] per-sweep max|delta| bit patterns
    unsigned long long *g_slot;  // [max_it] low 32: CTAs arrived, high 32: CTAs whose max >= tol
    ResState *state;
    int pair;                 // cluster size of the launch (0: none): links inside a cluster go through DSMEM
};

// (generic addressing: a slot is either in global memory or -- inside a CTA pair -- in the partner's shared memory)
__device__ __forceinline__ void ll_store(uint4 *p, double v, unsigned seq) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)b), "r"(seq),
                 "r"((unsigned)(b >> 32)), "r"(seq)
                 : "memory");
}

__device__ __forceinline__ uint4 ll_issue(const uint4 *p) {
    uint4 r;
    asm volatile("ld.volatile.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

// A neighbour that never shows up (the pair launch is not cooperative: CTAs might not all be resident) must not hang
// the GPU: after ~1 s of re-polling the launch is declared void (ResState::error) and every CTA bails out.
constexpr int RES_STOP_AHEAD = 10;   // sweeps between the stop decision and the stop (> the warps' maximum drift of 7.5 sweeps)
constexpr unsigned RES_SPIN_CHECK = 1023u, RES_SPIN_LIMIT = 1u << 21;
__device__ __forceinline__ bool res_give_up(unsigned &spins, int *err) {
    if ((++spins & RES_SPIN_CHECK) != 0u) return false;
    if (spins >= RES_SPIN_LIMIT) atomicExch(err, 1);
    return *((volatile int *)err) != 0;
}

__device__ __forceinline__ double ll_consume(uint4 r, const uint4 *p, unsigned seq, int *err) {
    unsigned spins = 0u;
    while (r.y != seq || r.w != seq) {
        r = ll_issue(p);
        if (res_give_up(spins, err)) break;
    }
    return __longlong_as_double((long long)(((unsigned long long)r.z << 32) | r.x));
}

// the same for a thread that keeps going after a voided launch: once it has given up it does not wait again
__device__ __forceinline__ double ll_consume(uint4 r, const uint4 *p, unsigned seq, int *err, bool &dead) {
    unsigned spins = 0u;
    while ((r.y != seq || r.w != seq) && !dead) {
        r = ll_issue(p);
        if (res_give_up(spins, err)) dead = true;
    }
    return __longlong_as_double((long long)(((unsigned long long)r.z << 32) | r.x));
}

// Per-thread strip state.
template <int NR>
struct Strip {
    double v[NR][2];   // phi of (row j, column 2k+q)
    double D[NR][2];
    // neighbour count and omega/cnt by row class (T = row 0, M = rows 1..NR-2, B = row NR-1) and parity
    double cT[2], cM[2], cB[2], wT[2], wM[2], wB[2];
    unsigned long long masks;  // 4-bit neighbour mask per cell at bit 8j+4q (bit0 left, 1 up, 2 right, 3 down)
    unsigned valid;            // bit 2j+q
};

// Arithmetic of the update of the active cell of row J in a colour phase (no shared-memory traffic: `nb` was read
// before, the new value is stored afterwards, so the NR independent dependency chains of a phase overlap).
// P0 = column parity of the active cell of row 0; row j's active cell has parity (P0+j)&1.
template <int NR, int P0, bool FAST, bool EDGE, int J>
__device__ __forceinline__ bool res_cell(Strip<NR> &s, const double nb, const double hu, const double hd,
                                         const ResParams &p, double &lmax) {
    constexpr int j = J;
    constexpr int q = (P0 + j) & 1;
    const double own = s.v[j][q ^ 1];
    const double l = (q == 0) ? nb : own, r = (q == 0) ? own : nb;
    const double u = (j == 0) ? hu : s.v[(j == 0) ? 0 : j - 1][q];
    const double d = (j == NR - 1) ? hd : s.v[(j == NR - 1) ? j : j + 1][q];
    const double val = s.v[j][q];
    if (FAST) {
        // only the first / last slab (EDGE) has rows whose neighbour count differs from the interior rows'
        const double cnt = (EDGE && j == 0) ? s.cT[q] : ((EDGE && j == NR - 1) ? s.cB[q] : s.cM[q]);
        const double wv = (EDGE && j == 0) ? s.wT[q] : ((EDGE && j == NR - 1) ? s.wB[q] : s.wM[q]);
        const double sum = ((l + u) + r) + d;  // ghosts are 0.0: identical to skipping them
        const double delta = wv * ((sum - cnt * val) - s.D[j][q]);
        const double ad = fabs(delta);
        if (ad > lmax) lmax = ad;
        s.v[j][q] = val + delta;
        return true;
    } else {
        if (!((s.valid >> (2 * j + q)) & 1u)) return false;
        const unsigned mm = (unsigned)(s.masks >> (8 * j + 4 * q)) & 15u;
        double sum = 0.0;
        if (mm & 1) sum += l;
        if (mm & 2) sum += u;
        if (mm & 4) sum += r;
        if (mm & 8) sum += d;
        const int cnt = __popc(mm);
        const double delta = wsel(p.w, cnt) * (sum - (double)cnt * val - s.D[j][q]);
        const double ad = fabs(delta);
        if (ad > lmax) lmax = ad;
        s.v[j][q] = val + delta;
        return true;
    }
}

template <int NR, int P0, int J>
__device__ __forceinline__ double res_nb(const double *__restrict__ smk) {
    constexpr int q = (P0 + J) & 1, Kp = RES_KP;
    // the one neighbour owned by another thread: left (q = 0) or right (q = 1), other parity row in smem
    return (q == 0) ? smk[(J * 2 + 1) * Kp - 1] : smk[(J * 2 + 0) * Kp + 1];
}

template <int NR, int P0, bool FAST, bool EDGE, int J>
struct InteriorRows {  // rows J .. NR-2 (compile-time recursion keeps every register index static)
    static __device__ __forceinline__ void load(const double *__restrict__ smk, double (&nb)[NR]) {
        if constexpr (J <= NR - 2) {
            nb[J] = res_nb<NR, P0, J>(smk);
            InteriorRows<NR, P0, FAST, EDGE, J + 1>::load(smk, nb);
        }
    }
    static __device__ __forceinline__ void compute(Strip<NR> &s, const double (&nb)[NR], bool (&ok)[NR], const ResParams &p,
                                                   double &lmax) {
        if constexpr (J <= NR - 2) {
            ok[J] = res_cell<NR, P0, FAST, EDGE, J>(s, nb[J], 0.0, 0.0, p, lmax);
            InteriorRows<NR, P0, FAST, EDGE, J + 1>::compute(s, nb, ok, p, lmax);
        }
    }
    static __device__ __forceinline__ void store(const Strip<NR> &s, double *__restrict__ smk, const bool (&ok)[NR]) {
        if constexpr (J <= NR - 2) {
            constexpr int q = (P0 + J) & 1;
            if (ok[J]) smk[(J * 2 + q) * RES_KP] = s.v[J][q];
            InteriorRows<NR, P0, FAST, EDGE, J + 1>::store(s, smk, ok);
        }
    }
};

// One colour phase of one slab: interior rows first (they need nothing from other CTAs), then the halo messages of
// the neighbours' previous phase are consumed and the two boundary rows are updated and exported -- so a message is
// in flight while both CTAs work on their interiors.  Within each group all shared-memory reads come first, then the
// arithmetic, then the writes.  (Measured alternatives, all bit-exact and none faster -- tools/experiments/README.md:
// boundary rows first with the polls for the next phase issued mid-interior; polls delayed by 1..5 interior rows;
// re-polling both messages with both loads in flight.)
template <int NR, int P0, bool FAST, bool EDGE>
__device__ __forceinline__ void res_phase(Strip<NR> &s, double *__restrict__ smk, double &hu, double &hd, const ResParams &p,
                                          double &lmax, const int k, uint4 *ll_up, uint4 *ll_dn, const unsigned seq,
                                          const uint4 *in_t, const uint4 *in_b, const bool first) {
    uint4 rt = make_uint4(0, 0, 0, 0), rb = make_uint4(0, 0, 0, 0);
    const uint4 *pt = first ? nullptr : in_t, *pb = first ? nullptr : in_b;
    if (pt) rt = ll_issue(pt);  // polls for the cells the neighbours produced in their previous phase (seq-1)
    if (pb) rb = ll_issue(pb);
    double nb[NR];
    bool ok[NR];
    nb[0] = res_nb<NR, P0, 0>(smk);
    if constexpr (NR >= 2) nb[NR - 1] = res_nb<NR, P0, NR - 1>(smk);
    InteriorRows<NR, P0, FAST, EDGE, 1>::load(smk, nb);
    InteriorRows<NR, P0, FAST, EDGE, 1>::compute(s, nb, ok, p, lmax);
    InteriorRows<NR, P0, FAST, EDGE, 1>::store(s, smk, ok);
    if (pt) hu = ll_consume(rt, pt, seq - 1u, &p.state->error);
    if (pb) hd = ll_consume(rb, pb, seq - 1u, &p.state->error);
    constexpr int q0 = P0 & 1, qb = (P0 + NR - 1) & 1;
    const bool ok0 = res_cell<NR, P0, FAST, EDGE, 0>(s, nb[0], hu, hd, p, lmax);
    bool okb = false;
    if constexpr (NR >= 2) okb = res_cell<NR, P0, FAST, EDGE, NR - 1>(s, nb[NR - 1], hu, hd, p, lmax);
    if (ok0) {
        smk[(0 * 2 + q0) * RES_KP] = s.v[0][q0];
        if (ll_up) ll_store(ll_up + q0 * RES_NT + k, s.v[0][q0], seq);
        if (NR == 1 && ll_dn) ll_store(ll_dn + q0 * RES_NT + k, s.v[0][q0], seq);
    }
    if constexpr (NR >= 2) {
        if (okb) {
            smk[((NR - 1) * 2 + qb) * RES_KP] = s.v[NR - 1][qb];
            if (ll_dn) ll_store(ll_dn + qb * RES_NT + k, s.v[NR - 1][qb], seq);
        }
    }
}

template <int NR, bool EDGE>
__device__ __forceinline__ void res_body(const ResParams &p, const int r0) {
    extern __shared__ double smem[];  // [NR][2][Kp], element (j, q, k) at (j*2+q)*Kp + 1 + k; pads stay 0.0
    __shared__ unsigned arrive[16];   // per sweep (ring): low half = warps done, high half = warps with a |delta| >= tol
    __shared__ int s_stop;

    const int tid = threadIdx.x, cta = blockIdx.x, k = tid;
    const int W = p.W, H = p.H, K = p.K;
    constexpr int Kp = RES_KP;
    const bool has_up = cta > 0, has_dn = cta + 1 < p.P;

    for (int i = tid; i < NR * 2 * Kp; i += RES_NT) smem[i] = 0.0;
    // CTA pairs (clusters of two): the slots of the link INSIDE the pair live in shared memory behind the strip arrays
    // (same offset in both CTAs); the partner writes them through DSMEM and this CTA polls its own shared memory
    uint4 *pslot = reinterpret_cast<uint4 *>(smem + p.nr_big * 2 * Kp);   // [2][RES_SLOTS]: from the CTA above / from below
    if (p.pair)
        for (int i = tid; i < 2 * RES_SLOTS; i += RES_NT) pslot[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 16) arrive[tid] = 0u;
    if (tid == 0) s_stop = 0;
    __syncthreads();

    // ---- load the strip -----------------------------------------------------------------------
    Strip<NR> s;
    s.masks = 0ull;
    s.valid = 0u;
    bool fast = NR >= 2;
#pragma unroll
    for (int j = 0; j < NR; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int x = 2 * k + q, y = r0 + j;
            s.v[j][q] = 0.0;
            s.D[j][q] = 0.0;
            if (k < K && x < W && y < H) {
                const size_t i = (size_t)y * W + x;
                unsigned geo = (x != 0 ? 1u : 0u) | (y != 0 ? 2u : 0u) | (x != W - 1 ? 4u : 0u) | (y != H - 1 ? 8u : 0u);
                unsigned mm = 0;
                if ((geo & 1) && !isnan(p.D[i - 1])) mm |= 1;
                if ((geo & 2) && !isnan(p.D[i - W])) mm |= 2;
                if ((geo & 4) && !isnan(p.D[i + 1])) mm |= 4;
                if ((geo & 8) && !isnan(p.D[i + W])) mm |= 8;
                if (mm != geo) fast = false;
                s.masks |= (unsigned long long)mm << (8 * j + 4 * q);
                s.valid |= 1u << (2 * j + q);
                s.D[j][q] = p.D[i];
                s.v[j][q] = p.phi[i];
                smem[(j * 2 + q) * Kp + 1 + k] = s.v[j][q];
            } else {
                fast = false;
            }
        }
    const bool idle = s.valid == 0u;  // k >= K, or a slab without rows for this thread
    // neighbour counts / weights by row class and parity (fast path only; geometric masks)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int cT = __popc((unsigned)(s.masks >> (4 * q)) & 15u);
        const int cM = __popc((unsigned)(s.masks >> (8 * (NR >= 3 ? 1 : 0) + 4 * q)) & 15u);
        const int cB = __popc((unsigned)(s.masks >> (8 * (NR - 1) + 4 * q)) & 15u);
        s.cT[q] = (double)cT; s.cM[q] = (double)cM; s.cB[q] = (double)cB;
        s.wT[q] = wsel(p.w, cT); s.wM[q] = wsel(p.w, cM); s.wB[q] = wsel(p.w, cB);
    }

    uint4 *ll_up = has_up ? p.ll + ((size_t)(cta - 1) * 2 + 1) * RES_SLOTS : nullptr;   // neighbour above: its "from below" slots
    uint4 *ll_dn = has_dn ? p.ll + ((size_t)(cta + 1) * 2 + 0) * RES_SLOTS : nullptr;   // neighbour below: its "from above" slots
    const uint4 *in_top = p.ll + ((size_t)cta * 2 + 0) * RES_SLOTS;
    const uint4 *in_bot = p.ll + ((size_t)cta * 2 + 1) * RES_SLOTS;
    if (p.pair) {
        cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
        const int rank = cta % p.pair;
        if (rank != 0) {                          // the CTA above is in this cluster: its "from below" slots, my "from above"
            ll_up = cl.map_shared_rank(pslot + RES_SLOTS, rank - 1); in_top = pslot;
        }
        if (rank != p.pair - 1 && has_dn) {       // the CTA below is in this cluster
            ll_dn = cl.map_shared_rank(pslot, rank + 1); in_bot = pslot + RES_SLOTS;
        }
    }
    if (idle) { ll_up = nullptr; ll_dn = nullptr; }
    double *smk = smem + 1 + k;

    // halo values for the first phase come from the global field; later ones arrive as LL messages
    const int p0_first = r0 & 1;  // colour 0: active parity of row y is y&1
    double hu = 0.0, hd = 0.0;
    {
        const int xt = 2 * k + p0_first, xb = 2 * k + ((p0_first + NR - 1) & 1);
        if (has_up && !idle && xt < W) hu = p.phi[(size_t)(r0 - 1) * W + xt];
        if (has_dn && !idle && xb < W) hd = p.phi[(size_t)(r0 + NR) * W + xb];
    }
    __syncthreads();
    if (p.pair) cooperative_groups::this_cluster().sync();   // the partner's slots are zeroed before anything is sent

    const int max_it = p.max_it;
    int sweep = 0, conv_at = 0;
    for (;; ++sweep) {
        double lmax = 0.0;
        unsigned long long pend = 0ull;
        const int e = sweep - p.lag;
        int errv = 0;
        if (tid == 0) {  // both consumed at the end of the sweep
            if (e >= 0) pend = *((const volatile unsigned long long *)(p.g_slot + e));
            errv = *((volatile int *)&p.state->error);
        }
#pragma unroll
        for (int colour = 0; colour < 2; ++colour) {
            const unsigned seq = 2u * (unsigned)sweep + (unsigned)colour + 1u;
            const int p0 = (r0 + colour) & 1;
            if (!idle) {
                // slots of this phase's halo cells: column of the active cell of row 0 / row NR-1
                const int qb = (p0 + NR - 1) & 1, xt = 2 * k + p0, xb = 2 * k + qb;
                const uint4 *it = (has_up && xt < W) ? in_top + p0 * RES_NT + k : nullptr;
                const uint4 *ib = (has_dn && xb < W) ? in_bot + qb * RES_NT + k : nullptr;
                const bool first = seq == 1u;
                if (p0 == 0) {
                    if (fast) res_phase<NR, 0, true, EDGE>(s, smk, hu, hd, p, lmax, k, ll_up, ll_dn, seq, it, ib, first);
                    else res_phase<NR, 0, false, EDGE>(s, smk, hu, hd, p, lmax, k, ll_up, ll_dn, seq, it, ib, first);
                } else {
                    if (fast) res_phase<NR, 1, true, EDGE>(s, smk, hu, hd, p, lmax, k, ll_up, ll_dn, seq, it, ib, first);
                    else res_phase<NR, 1, false, EDGE>(s, smk, hu, hd, p, lmax, k, ll_up, ll_dn, seq, it, ib, first);
                }
            }
            if (colour == 1 && tid == 0) {   // convergence duty: the verdict on sweep e = sweep - lag
                if (e >= 0 && conv_at == 0 && !errv) {
                    unsigned spins = 0u;
                    while ((unsigned)pend != (unsigned)p.P) {
                        pend = *((const volatile unsigned long long *)(p.g_slot + e));
                        if (res_give_up(spins, &p.state->error)) { errv = 1; break; }
                    }
                    if (!errv && (pend >> 32) == 0ull) conv_at = e + 1;
                }
                // the warps of a CTA drift apart by up to 15 phases (below): the stop is announced RES_STOP_AHEAD sweeps
                // ahead so that every warp of every CTA leaves after the same sweep
                // (announced once: a void launch keeps errv set, and re-announcing every sweep would never stop)
                if ((conv_at == e + 1 && e >= 0) || (errv && *((volatile int *)&s_stop) == 0))
                    *((volatile int *)&s_stop) = sweep + 1 + RES_STOP_AHEAD;
            }
            // After a phase a warp only depends on its two neighbouring warps (the left / right cells of its first and
            // last lane): two rounds of pairwise named barriers (pairs (0,1),(2,3).. then (1,2),(3,4)..) instead of a
            // CTA-wide barrier, so a warp whose halo message is late holds up two warps, not sixteen.
            {
                const int wp = tid >> 5;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + (wp & ~1)) : "memory");
                if (wp >= 1 && wp <= RES_NT / 32 - 2) asm volatile("bar.sync %0, 64;" ::"r"((wp & 1) ? 1 + wp : wp) : "memory");
            }
        }
        // End of the sweep, per WARP (no CTA-wide barrier): the warp's verdict goes into this sweep's arrival word; the
        // last warp to arrive publishes the CTA's arrival + verdict with one atomic.  The maximum itself is only ever
        // read for sweeps at which every warp is below tol (the sweep the solve stops at) and for the last sweeps of a
        // launch that hits its cap (the host scans those): reduce it only then.
        {
            const bool wany = __any_sync(0xffffffffu, lmax >= p.tol);
            if (!wany || sweep + p.lag + 3 >= max_it) {
                const double wm = warp_max(lmax);
                if ((tid & 31) == 0 && wm > 0.0) atomicMax(p.g_max + sweep, (unsigned long long)__double_as_longlong(wm));
            }
            if ((tid & 31) == 0) {
                const unsigned add = 1u | (wany ? 0x10000u : 0u);
                const unsigned old = atomicAdd(&arrive[sweep & 15], add);
                if ((old & 0xffffu) == RES_NT / 32 - 1) {
                    arrive[sweep & 15] = 0u;   // next used 16 sweeps from now
                    atomicAdd(p.g_slot + sweep, 1ull | (((old + add) >> 16) ? (1ull << 32) : 0ull));
                }
            }
        }
        if (*((volatile int *)&s_stop) == sweep + 1 || sweep + 1 >= max_it) break;
    }

    if (p.pair) cooperative_groups::this_cluster().sync();   // no CTA leaves while its partner may still store into it
    // ---- write the strip back: only after EVERY CTA has finished its sweeps, and not when the launch was declared
    // void -- so the input field is intact whenever the host has to repeat the launch --------------------------------
    if (tid == 0) {
        int ok = *((volatile int *)&p.state->error) == 0;
        if (ok) {
            atomicAdd(&p.state->done, 1);
            unsigned spins = 0u;
            while (*((volatile int *)&p.state->done) != p.P)
                if (res_give_up(spins, &p.state->error)) { ok = 0; break; }
        }
        s_stop = ok ? -1 : -2;
    }
    __syncthreads();
    if (s_stop != -1) return;
#pragma unroll
    for (int j = 0; j < NR; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if ((s.valid >> (2 * j + q)) & 1u) p.phi[(size_t)(r0 + j) * W + 2 * k + q] = s.v[j][q];
    if (tid == 0 && cta == 0) {
        p.state->sweeps = sweep + 1;
        p.state->converged_at = conv_at;
    }
}

template <int NR>
__global__ void __launch_bounds__(RES_NT, 1) sor_resident_kernel(const __grid_constant__ ResParams p) {
    // the first and the last slab carry the domain's top / bottom row (different neighbour counts)
    const int c = (int)blockIdx.x;
    const bool edge = c == 0 || c + 1 == p.P;
    if (c >= p.P) {  // padding CTA of the last pair: only keeps the cluster barriers balanced
        cooperative_groups::this_cluster().sync();
        cooperative_groups::this_cluster().sync();
    } else if (c < p.n_big) {
        if (edge) res_body<NR, true>(p, c * NR);
        else res_body<NR, false>(p, c * NR);
    } else if constexpr (NR >= 2) {  // slabs one row shorter: every slab is full, every CTA stays on the fast path
        const int r0 = p.n_big * NR + (c - p.n_big) * (NR - 1);
        if (edge) res_body<NR - 1, true>(p, r0);
        else res_body<NR - 1, false>(p, r0);
    }
}

// ------------------------------------------------------------------------------------------------
// Deep-halo variant (round 2): ONE neighbour exchange per sweep instead of one per colour phase.
//
// The chain that paces the kernel above is message latency: a boundary cell of phase p needs the neighbour's boundary
// cell of phase p-1, i.e. two L2 round trips per sweep lie on the critical path (~1500 cycles each, tools/ll_latency.cu)
// against ~600 cycles of arithmetic.  Here every CTA also keeps the colour-0 cells of the row just outside its slab
// on either side (rows r0-1 and r0+NR) and updates them REDUNDANTLY -- same inputs, same expression, same bits as the
// owner's update -- so the colour-1 phase needs nothing from outside, and a sweep exchanges once: after its colour-1
// phase a CTA sends the colour-1 cells of its two outermost rows on either side (2 x K messages per link and sweep,
// the same volume as before).  The schedule of a sweep:
//   colour 0: rows 1..NR-2 (own data only) | consume the messages | halo row above, rows 0 and NR-1, halo row below
//   colour 1: rows 0, 1, NR-2, NR-1 -> send | rows 2..NR-3
// so a message is in flight during (NR-4) + (NR-2) row updates, and one latency per sweep is left on the chain.
// The colour-1 cell of the halo row that belongs to thread k-1 / k+1 comes from the neighbouring lane by shuffle; the
// first / last lane of a warp polls that slot itself.  Slots are double-buffered by sweep parity: the sender of
// message s+2 has consumed the receiver's message s+1, which was sent after message s had been consumed.
// Limits: even W, no NaN holes, every slab >= 2 rows and the first >= 3 (H > 2 * CTAs): anything else runs the kernel
// above (a NaN found while loading voids the launch with ResState::error = 2 and the host repeats it there).
// ------------------------------------------------------------------------------------------------
constexpr int RES2_SLOTS = 4 * RES_NT;   // [sweep parity][0: row next to the receiver, 1: the row behind it][RES_NT]
constexpr int RES_ERR_UNSUPPORTED = 2;
// Build switches of the deep-halo kernel (all bit-identical; the defaults are the measured winners, tools/experiments):
//   PCD_DEEP_DSMEM    D lives in shared memory instead of registers (28 registers less at 7 rows: no spills, no
//                     re-materialised neighbour counts, room to overlap the cells' dependency chains)
//   PCD_DEEP_FASTPOLL the six messages of a sweep are tested with ONE branch; the per-message poll loop is the slow path
//   PCD_DEEP_PRESEND  the halo rows of the first sweep travel as messages too (sent before the loop from the loaded
//                     strip), so the loop has no first-sweep case
#ifndef PCD_DEEP_DSMEM
#define PCD_DEEP_DSMEM 1
#endif
#ifndef PCD_DEEP_FASTPOLL
#define PCD_DEEP_FASTPOLL 1
#endif
#ifndef PCD_DEEP_PRESEND
#define PCD_DEEP_PRESEND 1
#endif
#ifndef PCD_DEEP_SPLITBAR
#define PCD_DEEP_SPLITBAR 0     // sweep-end warp-pair barrier as mbarrier arrive (after the last stores) / wait (before the next reads)
#endif
#ifndef PCD_DEEP_POLL_AFTER
#define PCD_DEEP_POLL_AFTER 0   // interior rows of colour 0 updated before the first poll of a sweep is issued
#endif
constexpr bool DEEP_DSMEM = PCD_DEEP_DSMEM != 0, DEEP_FASTPOLL = PCD_DEEP_FASTPOLL != 0, DEEP_PRESEND = PCD_DEEP_PRESEND != 0;

__device__ __forceinline__ double ll_value(const uint4 r) {
    return __longlong_as_double((long long)(((unsigned long long)r.z << 32) | r.x));
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra MBAR_WAIT_%=;\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

template <int A, int B, class F>
__device__ __forceinline__ void static_rows(F &&f) {
    if constexpr (A < B) {
        f(std::integral_constant<int, A>{});
        static_rows<A + 1, B>(f);
    }
}

__device__ __forceinline__ double sor_cell(const double val, const double l, const double u, const double r, const double d,
                                           const double cnt, const double wv, const double Dv, double &lmax) {
    const double sum = ((l + u) + r) + d;  // ghosts are 0.0: identical to skipping them
    const double delta = wv * ((sum - cnt * val) - Dv);
    // (a SIGNED running maximum -- |delta| > |lmax| ? delta : lmax, 3 instructions instead of 4 -- was measured: 7 % slower
    // at 1024^2 once a ptxas miscompile of its first two comparisons was worked around; tools/experiments/README.md)
    const double ad = fabs(delta);
    if (ad > lmax) lmax = ad;
    return val + delta;
}

// Transposed solves (TR): the kernel's left/up/right/down are the grid's up/left/down/right, so the reference's summation
// order ((l + u) + r) + d reads ((l' + u') + d') + r' here (l + u commutes bit for bit; the last two additions swap)
template <bool TR>
__device__ __forceinline__ double sor_cell_o(const double val, const double l, const double u, const double r, const double d,
                                             const double cnt, const double wv, const double Dv, double &lmax) {
    return TR ? sor_cell(val, l, u, d, r, cnt, wv, Dv, lmax) : sor_cell(val, l, u, r, d, cnt, wv, Dv, lmax);
}

template <int NR>
struct DeepStrip {
    double v[NR][2], D[DEEP_DSMEM ? 1 : NR][2];   // D: registers, or shared memory (DEEP_DSMEM)
    double cM[2], wM[2], cT[2], wT[2], cB[2], wB[2];   // T / B only live in the first / last slab
    double hu, hd, Du, Dd;   // colour-0 cell of the halo rows r0-1 / r0+NR in this thread's column pair, and its D
};

// update of the colour-C cell of slab row J (P0 = column parity of the colour-0 cell of row 0)
template <int NR, int P0, bool EDGE, bool TR, int C, int J>
__device__ __forceinline__ void deep_cell(DeepStrip<NR> &s, const double nb, const double Dv, const double upv, const double dnv, double &lmax) {
    constexpr int q = (P0 + C + J) & 1;
    const double own = s.v[J][q ^ 1];
    const double l = (q == 0) ? nb : own, r = (q == 0) ? own : nb;
    const double u = (J == 0) ? upv : s.v[(J == 0) ? 0 : J - 1][q];
    const double d = (J == NR - 1) ? dnv : s.v[(J == NR - 1) ? J : J + 1][q];
    const double cnt = (EDGE && J == 0) ? s.cT[q] : ((EDGE && J == NR - 1) ? s.cB[q] : s.cM[q]);
    const double wv = (EDGE && J == 0) ? s.wT[q] : ((EDGE && J == NR - 1) ? s.wB[q] : s.wM[q]);
    s.v[J][q] = sor_cell_o<TR>(s.v[J][q], l, u, r, d, cnt, wv, Dv, lmax);
}

// D of the colour-C cell of row J: a register, or this thread's own shared-memory word (unit stride over the warp)
template <int NR, int P0, int C, int J>
__device__ __forceinline__ double deep_D(const DeepStrip<NR> &s, const double *__restrict__ smD) {
    constexpr int q = (P0 + C + J) & 1;
    if constexpr (DEEP_DSMEM) return smD[(J * 2 + q) * RES_KP];
    else return s.D[J][q];
}

template <int P0, int C, int J>
__device__ __forceinline__ double deep_nb(const double *__restrict__ smk) {
    constexpr int q = (P0 + C + J) & 1;
    return (q == 0) ? smk[(J * 2 + 1) * RES_KP - 1] : smk[(J * 2 + 0) * RES_KP + 1];
}

// rows [A, B) of colour C: all shared-memory reads, then the arithmetic, then the writes
template <int NR, int P0, bool EDGE, bool TR, int C, int A, int B>
__device__ __forceinline__ void deep_rows(DeepStrip<NR> &s, double *__restrict__ smk, const double *__restrict__ smD,
                                          const double upv, const double dnv, double &lmax) {
    if constexpr (A < B) {
        double nb[B - A], dv[B - A];
        static_rows<A, B>([&](auto Jc) {
            constexpr int J = decltype(Jc)::value;
            nb[J - A] = deep_nb<P0, C, J>(smk);
            dv[J - A] = deep_D<NR, P0, C, J>(s, smD);
        });
        static_rows<A, B>([&](auto Jc) {
            constexpr int J = decltype(Jc)::value;
            deep_cell<NR, P0, EDGE, TR, C, J>(s, nb[J - A], dv[J - A], upv, dnv, lmax);
        });
        static_rows<A, B>([&](auto Jc) {
            constexpr int J = decltype(Jc)::value, q = (P0 + C + J) & 1;
            smk[(J * 2 + q) * RES_KP] = s.v[J][q];
        });
    }
}

template <int NR, int P0, bool EDGE, bool TR>
__device__ __forceinline__ void deep_body(const ResParams &p, const int r0) {
    static_assert(NR >= 2, "deep halos need two rows per slab");
    extern __shared__ double smem[];  // phi [nr_big][2][Kp] as in res_body, D likewise (DEEP_DSMEM), then the slots of the links inside a cluster
    __shared__ unsigned arrive[16];
    __shared__ unsigned long long pairbar[16];   // DEEP_SPLITBAR: one mbarrier per pair of neighbouring warps (w, w+1), two arrivals each
    __shared__ int s_stop;

    const int tid = threadIdx.x, cta = blockIdx.x, k = tid, lane = tid & 31;
    const int W = p.W, H = p.H, K = p.K;
    constexpr int Kp = RES_KP;
    const bool has_up = !EDGE || cta > 0, has_dn = !EDGE || cta + 1 < p.P;
    int *const err = &p.state->error;
    constexpr int PLANES = DEEP_DSMEM ? 2 : 1;

    for (int i = tid; i < NR * 2 * Kp; i += RES_NT) smem[i] = 0.0;
    uint4 *pslot = reinterpret_cast<uint4 *>(smem + p.nr_big * 2 * Kp * PLANES);   // [2][RES2_SLOTS]: from the CTA above / from below
    double *const smD = smem + p.nr_big * 2 * Kp + 1 + k;   // D of (row j, parity q) of this thread at smD[(j*2+q)*Kp]
    if (p.pair)
        for (int i = tid; i < 2 * RES2_SLOTS; i += RES_NT) pslot[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 16) arrive[tid] = 0u;
    if (PCD_DEEP_SPLITBAR && tid < 16) mbar_init(&pairbar[tid], 2u);
    if (tid == 0) s_stop = 0;
    __syncthreads();

    // ---- load the strip -----------------------------------------------------------------------
    DeepStrip<NR> s;
    const bool idle = k >= K;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < NR; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int x = 2 * k + q, y = r0 + j;
            s.v[j][q] = 0.0;
            if constexpr (!DEEP_DSMEM) s.D[j][q] = 0.0;
            if (!idle) {
                const size_t i = (size_t)y * W + x;
                const double Dv = p.D[i];
                if constexpr (DEEP_DSMEM) smD[(j * 2 + q) * Kp] = Dv;   // only ever read by this thread
                else s.D[j][q] = Dv;
                s.v[j][q] = p.phi[i];
                bad = bad || isnan(Dv);
                smem[(j * 2 + q) * Kp + 1 + k] = s.v[j][q];
            }
        }
    constexpr int qa = (P0 + 1) & 1, qd = (P0 + NR) & 1;   // column parity of the colour-0 cell of the halo row above / below
    s.hu = 0.0; s.hd = 0.0; s.Du = 0.0; s.Dd = 0.0;
    if (!idle && has_up) {   // NaN anywhere in D is seen by its owner above; the halo cells only need their own D
        const size_t i = (size_t)(r0 - 1) * W + 2 * k + qa;
        s.hu = p.phi[i]; s.Du = p.D[i];
    }
    if (!idle && has_dn) {
        const size_t i = (size_t)(r0 + NR) * W + 2 * k + qd;
        s.hd = p.phi[i]; s.Dd = p.D[i];
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int x = 2 * k + q;
        const int cx = (x != 0 ? 1 : 0) + (x != W - 1 ? 1 : 0);
        const int cT = cx + (r0 != 0 ? 1 : 0) + 1, cB = cx + 1 + (r0 + NR != H ? 1 : 0), cM = cx + 2;
        s.cT[q] = (double)cT; s.cM[q] = (double)cM; s.cB[q] = (double)cB;
        s.wT[q] = wsel(p.w, cT); s.wM[q] = wsel(p.w, cM); s.wB[q] = wsel(p.w, cB);
    }
    const unsigned amask = __ballot_sync(0xffffffffu, !idle);
    if (__syncthreads_or(bad ? 1 : 0)) {   // the masked neighbour rule is the other kernel's business: void the launch
        if (tid == 0) atomicExch(err, RES_ERR_UNSUPPORTED);
        if (p.pair) {
            cooperative_groups::this_cluster().sync();
            cooperative_groups::this_cluster().sync();
        }
        return;
    }

    uint4 *ll_up = has_up ? p.ll + ((size_t)(cta - 1) * 2 + 1) * RES2_SLOTS : nullptr;
    uint4 *ll_dn = has_dn ? p.ll + ((size_t)(cta + 1) * 2 + 0) * RES2_SLOTS : nullptr;
    const uint4 *in_top = p.ll + ((size_t)cta * 2 + 0) * RES2_SLOTS;
    const uint4 *in_bot = p.ll + ((size_t)cta * 2 + 1) * RES2_SLOTS;
    if (p.pair) {
        cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
        const int rank = cta % p.pair;
        if (rank != 0) { ll_up = cl.map_shared_rank(pslot + RES2_SLOTS, rank - 1); in_top = pslot; }
        if (rank != p.pair - 1 && has_dn) { ll_dn = cl.map_shared_rank(pslot, rank + 1); in_bot = pslot + RES2_SLOTS; }
    }
    double *smk = smem + 1 + k;
    // the colour-1 cell left / right of the halo cell that belongs to another thread: lane +-1, or -- for the first /
    // last lane of a warp -- that thread's slot polled directly
    const int keu = (qa == 0) ? k - 1 : k + 1, ked = (qd == 0) ? k - 1 : k + 1;
    const bool nbu_ok = keu >= 0 && keu < K, nbd_ok = ked >= 0 && ked < K;
    const bool edge_u = has_up && nbu_ok && lane == ((qa == 0) ? 0 : 31);
    const bool edge_d = has_dn && nbd_ok && lane == ((qd == 0) ? 0 : 31);
    __syncthreads();
    if (p.pair) cooperative_groups::this_cluster().sync();

    // sequence number of the message consumed by sweep s: s (the halo rows of sweep 0 are read from the input field), or
    // s + 1 when the first sweep's halo rows travel as messages as well (slots are zeroed by the host: 0 is never valid)
    constexpr unsigned SEQ0 = DEEP_PRESEND ? 1u : 0u;
    if constexpr (DEEP_PRESEND) {
        if (!idle) {
            const int b1 = 2 * RES_NT + k;   // the buffer sweep 0 reads: ((0 + 1) & 1)
            if (ll_up) {
                ll_store(ll_up + b1, s.v[0][(P0 + 1) & 1], SEQ0);
                ll_store(ll_up + b1 + RES_NT, s.v[1][P0 & 1], SEQ0);
            }
            if (ll_dn) {
                ll_store(ll_dn + b1, s.v[NR - 1][(P0 + NR) & 1], SEQ0);
                ll_store(ll_dn + b1 + RES_NT, s.v[NR - 2][(P0 + NR - 1) & 1], SEQ0);
            }
        }
    }

    const int max_it = p.max_it;
    int sweep = 0, conv_at = 0;
    bool dead = false;   // this thread has seen the launch declared void
    for (;; ++sweep) {
        double lmax = 0.0;
        unsigned long long pend = 0ull;
        const int e = sweep - p.lag;
        int errv = 0;
        if (tid == 0) {
            if (e >= 0) pend = *((const volatile unsigned long long *)(p.g_slot + e));
            errv = *((volatile int *)err);
        }
        if (!idle) {
            const bool first = !DEEP_PRESEND && sweep == 0;
            const unsigned seq = (unsigned)sweep + SEQ0;      // the message consumed now was sent after sweep-1
            const int bin = ((sweep + 1) & 1) * 2 * RES_NT + k;
            // ---- colour 0 --------------------------------------------------------------------------
            double m1u = 0.0, m2u = 0.0, nbu = 0.0, m1d = 0.0, m2d = 0.0, nbd = 0.0;
            if constexpr (DEEP_FASTPOLL) {
                uint4 ru1 = make_uint4(0u, 0u, 0u, 0u), ru2 = ru1, rd1 = ru1, rd2 = ru1;
                uint4 rue, rde;   // only the first / last lane of a warp has these (read under edge_u / edge_d alone)
                // the first poll is issued after SPLIT - 1 of the interior rows: late enough to find the message, early
                // enough for the rest of the interior to cover its latency
                constexpr int SPLIT = (1 + PCD_DEEP_POLL_AFTER < NR - 1) ? 1 + PCD_DEEP_POLL_AFTER : (NR - 1 > 1 ? NR - 1 : 1);
                deep_rows<NR, P0, EDGE, TR, 0, 1, SPLIT>(s, smk, smD, 0.0, 0.0, lmax);
                if (!first) {
                    if (has_up) { ru1 = ll_issue(in_top + bin); ru2 = ll_issue(in_top + bin + RES_NT); if (edge_u) rue = ll_issue(in_top + bin + (keu - k)); }
                    if (has_dn) { rd1 = ll_issue(in_bot + bin); rd2 = ll_issue(in_bot + bin + RES_NT); if (edge_d) rde = ll_issue(in_bot + bin + (ked - k)); }
                }
                if (PCD_DEEP_SPLITBAR && sweep > 0) {   // the neighbouring warps' colour-1 values of the previous sweep
                    const int wp = tid >> 5;
                    const unsigned par = (unsigned)(sweep - 1) & 1u;
                    if (wp > 0) mbar_wait(&pairbar[wp - 1], par);
                    if (wp < RES_NT / 32 - 1) mbar_wait(&pairbar[wp], par);
                }
                deep_rows<NR, P0, EDGE, TR, 0, SPLIT, NR - 1>(s, smk, smD, 0.0, 0.0, lmax);
                if (first) {   // (!DEEP_PRESEND) the halo rows of the first sweep come from the input field
                    if (has_up) {
                        const double *row = p.phi + (size_t)(r0 - 1) * W;
                        m1u = row[2 * k + (qa ^ 1)]; m2u = row[2 * k + qa - W];
                        if (nbu_ok) nbu = row[2 * keu + (qa ^ 1)];
                    }
                    if (has_dn) {
                        const double *row = p.phi + (size_t)(r0 + NR) * W;
                        m1d = row[2 * k + (qd ^ 1)]; m2d = row[2 * k + qd + W];
                        if (nbd_ok) nbd = row[2 * ked + (qd ^ 1)];
                    }
                } else {
                    // ONE test for all messages of the sweep.  Whatever is still missing is then re-polled TOGETHER (all
                    // loads in flight at once: a round of re-polls costs one L2 round trip, not one per message)
                    unsigned spins = 0u;
                    for (;;) {
                        unsigned su1 = 0u, su2 = 0u, sue = 0u, sd1 = 0u, sd2 = 0u, sde = 0u;
                        if (has_up) { su1 = (ru1.y ^ seq) | (ru1.w ^ seq); su2 = (ru2.y ^ seq) | (ru2.w ^ seq); if (edge_u) sue = (rue.y ^ seq) | (rue.w ^ seq); }
                        if (has_dn) { sd1 = (rd1.y ^ seq) | (rd1.w ^ seq); sd2 = (rd2.y ^ seq) | (rd2.w ^ seq); if (edge_d) sde = (rde.y ^ seq) | (rde.w ^ seq); }
                        if ((((su1 | su2) | sue) | ((sd1 | sd2) | sde)) == 0u || dead) break;
                        if (su1) ru1 = ll_issue(in_top + bin);
                        if (su2) ru2 = ll_issue(in_top + bin + RES_NT);
                        if (sue) rue = ll_issue(in_top + bin + (keu - k));
                        if (sd1) rd1 = ll_issue(in_bot + bin);
                        if (sd2) rd2 = ll_issue(in_bot + bin + RES_NT);
                        if (sde) rde = ll_issue(in_bot + bin + (ked - k));
                        if (res_give_up(spins, err)) dead = true;
                    }
                    if (has_up) {
                        m1u = ll_value(ru1); m2u = ll_value(ru2);
                        nbu = (qa == 0) ? __shfl_up_sync(amask, m1u, 1) : __shfl_down_sync(amask, m1u, 1);
                        if (edge_u) nbu = ll_value(rue);
                        if (!nbu_ok) nbu = 0.0;
                    }
                    if (has_dn) {
                        m1d = ll_value(rd1); m2d = ll_value(rd2);
                        nbd = (qd == 0) ? __shfl_up_sync(amask, m1d, 1) : __shfl_down_sync(amask, m1d, 1);
                        if (edge_d) nbd = ll_value(rde);
                        if (!nbd_ok) nbd = 0.0;
                    }
                }
            } else {
                uint4 ru1 = make_uint4(0, 0, 0, 0), ru2 = ru1, rue = ru1, rd1 = ru1, rd2 = ru1, rde = ru1;
                if (!first) {
                    if (has_up) { ru1 = ll_issue(in_top + bin); ru2 = ll_issue(in_top + bin + RES_NT); if (edge_u) rue = ll_issue(in_top + bin + (keu - k)); }
                    if (has_dn) { rd1 = ll_issue(in_bot + bin); rd2 = ll_issue(in_bot + bin + RES_NT); if (edge_d) rde = ll_issue(in_bot + bin + (ked - k)); }
                }
                deep_rows<NR, P0, EDGE, TR, 0, 1, NR - 1>(s, smk, smD, 0.0, 0.0, lmax);
                if (first) {   // the halo rows of the first sweep come from the input field (untouched until every CTA is done)
                    if (has_up) {
                        const double *row = p.phi + (size_t)(r0 - 1) * W;
                        m1u = row[2 * k + (qa ^ 1)]; m2u = row[2 * k + qa - W];
                        if (nbu_ok) nbu = row[2 * keu + (qa ^ 1)];
                    }
                    if (has_dn) {
                        const double *row = p.phi + (size_t)(r0 + NR) * W;
                        m1d = row[2 * k + (qd ^ 1)]; m2d = row[2 * k + qd + W];
                        if (nbd_ok) nbd = row[2 * ked + (qd ^ 1)];
                    }
                } else {
                    if (has_up) {
                        m1u = ll_consume(ru1, in_top + bin, seq, err, dead);
                        m2u = ll_consume(ru2, in_top + bin + RES_NT, seq, err, dead);
                        nbu = (qa == 0) ? __shfl_up_sync(amask, m1u, 1) : __shfl_down_sync(amask, m1u, 1);
                        if (edge_u) nbu = ll_consume(rue, in_top + bin + (keu - k), seq, err, dead);
                        if (!nbu_ok) nbu = 0.0;
                    }
                    if (has_dn) {
                        m1d = ll_consume(rd1, in_bot + bin, seq, err, dead);
                        m2d = ll_consume(rd2, in_bot + bin + RES_NT, seq, err, dead);
                        nbd = (qd == 0) ? __shfl_up_sync(amask, m1d, 1) : __shfl_down_sync(amask, m1d, 1);
                        if (edge_d) nbd = ll_consume(rde, in_bot + bin + (ked - k), seq, err, dead);
                        if (!nbd_ok) nbd = 0.0;
                    }
                }
            }
            {
                double halo_max = 0.0;   // the owner of a halo cell counts its update
                const double nb0 = deep_nb<P0, 0, 0>(smk), nbl = deep_nb<P0, 0, NR - 1>(smk);
                const double D0 = deep_D<NR, P0, 0, 0>(s, smD), Dl = deep_D<NR, P0, 0, NR - 1>(s, smD);
                if (has_up)
                    s.hu = sor_cell_o<TR>(s.hu, (qa == 0) ? nbu : m1u, m2u, (qa == 0) ? m1u : nbu, s.v[0][qa], s.cM[qa], s.wM[qa], s.Du, halo_max);
                if (has_dn)
                    s.hd = sor_cell_o<TR>(s.hd, (qd == 0) ? nbd : m1d, s.v[NR - 1][qd], (qd == 0) ? m1d : nbd, m2d, s.cM[qd], s.wM[qd], s.Dd, halo_max);
                deep_cell<NR, P0, EDGE, TR, 0, 0>(s, nb0, D0, m1u, m1d, lmax);
                deep_cell<NR, P0, EDGE, TR, 0, NR - 1>(s, nbl, Dl, m1u, m1d, lmax);
                smk[(0 * 2 + (P0 & 1)) * Kp] = s.v[0][P0 & 1];
                smk[((NR - 1) * 2 + ((P0 + NR - 1) & 1)) * Kp] = s.v[NR - 1][(P0 + NR - 1) & 1];
            }
        }
        {
            const int wp = tid >> 5;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + (wp & ~1)) : "memory");
            if (wp >= 1 && wp <= RES_NT / 32 - 2) asm volatile("bar.sync %0, 64;" ::"r"((wp & 1) ? 1 + wp : wp) : "memory");
        }
        if (!idle) {
            // ---- colour 1: the rows the neighbours wait for first ------------------------------------
            const int bout = (sweep & 1) * 2 * RES_NT + k;
            const unsigned seq_out = (unsigned)sweep + 1u + SEQ0;
            constexpr int TOP = NR < 2 ? NR : 2;                    // rows [0, TOP) and [BOT, NR) are sent
            constexpr int BOT = (NR - 2 > TOP) ? NR - 2 : TOP;
            deep_rows<NR, P0, EDGE, TR, 1, 0, TOP>(s, smk, smD, s.hu, s.hd, lmax);
            deep_rows<NR, P0, EDGE, TR, 1, BOT, NR>(s, smk, smD, s.hu, s.hd, lmax);
            if (ll_up) {
                ll_store(ll_up + bout, s.v[0][(P0 + 1) & 1], seq_out);
                ll_store(ll_up + bout + RES_NT, s.v[1][P0 & 1], seq_out);
            }
            if (ll_dn) {
                ll_store(ll_dn + bout, s.v[NR - 1][(P0 + NR) & 1], seq_out);
                ll_store(ll_dn + bout + RES_NT, s.v[NR - 2][(P0 + NR - 1) & 1], seq_out);
            }
            deep_rows<NR, P0, EDGE, TR, 1, TOP, BOT>(s, smk, smD, s.hu, s.hd, lmax);
        }
        if (tid == 0) {   // convergence duty: the verdict on sweep e = sweep - lag (as in res_body)
            if (e >= 0 && conv_at == 0 && !errv) {
                unsigned spins = 0u;
                while ((unsigned)pend != (unsigned)p.P) {
                    pend = *((const volatile unsigned long long *)(p.g_slot + e));
                    if (res_give_up(spins, err)) { errv = 1; break; }
                }
                if (!errv && (pend >> 32) == 0ull) conv_at = e + 1;
            }
            if ((conv_at == e + 1 && e >= 0) || (errv && *((volatile int *)&s_stop) == 0))
                *((volatile int *)&s_stop) = sweep + 1 + RES_STOP_AHEAD;
        }
        if (PCD_DEEP_SPLITBAR) {   // arrive now (this warp's stores of the sweep are done), wait before the next sweep's first reads
            const int wp = tid >> 5;
            __syncwarp();
            if ((tid & 31) == 0) {
                if (wp > 0) mbar_arrive(&pairbar[wp - 1]);
                if (wp < RES_NT / 32 - 1) mbar_arrive(&pairbar[wp]);
            }
        } else {
            const int wp = tid >> 5;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + (wp & ~1)) : "memory");
            if (wp >= 1 && wp <= RES_NT / 32 - 2) asm volatile("bar.sync %0, 64;" ::"r"((wp & 1) ? 1 + wp : wp) : "memory");
        }
        {
            const bool wany = __any_sync(0xffffffffu, lmax >= p.tol);
            if (!wany || sweep + p.lag + 3 >= max_it) {
                const double wm = warp_max(lmax);
                if ((tid & 31) == 0 && wm > 0.0) atomicMax(p.g_max + sweep, (unsigned long long)__double_as_longlong(wm));
            }
            if ((tid & 31) == 0) {
                const unsigned add = 1u | (wany ? 0x10000u : 0u);
                const unsigned old = atomicAdd(&arrive[sweep & 15], add);
                if ((old & 0xffffu) == RES_NT / 32 - 1) {
                    arrive[sweep & 15] = 0u;
                    atomicAdd(p.g_slot + sweep, 1ull | (((old + add) >> 16) ? (1ull << 32) : 0ull));
                }
            }
        }
        if (*((volatile int *)&s_stop) == sweep + 1 || sweep + 1 >= max_it) break;
    }

    if (p.pair) cooperative_groups::this_cluster().sync();
    if (tid == 0) {
        int ok = *((volatile int *)err) == 0;
        if (ok) {
            atomicAdd(&p.state->done, 1);
            unsigned spins = 0u;
            while (*((volatile int *)&p.state->done) != p.P)
                if (res_give_up(spins, err)) { ok = 0; break; }
        }
        s_stop = ok ? -1 : -2;
    }
    __syncthreads();
    if (s_stop != -1) return;
    if (!idle) {
#pragma unroll
        for (int j = 0; j < NR; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) p.phi[(size_t)(r0 + j) * W + 2 * k + q] = s.v[j][q];
    }
    if (tid == 0 && cta == 0) {
        p.state->sweeps = sweep + 1;
        p.state->converged_at = conv_at;
    }
}

template <int NR, bool TR>
__global__ void __launch_bounds__(RES_NT, 1) sor_resident_deep_kernel(const __grid_constant__ ResParams p) {
    const int c = (int)blockIdx.x;
    const bool edge = c == 0 || c + 1 == p.P;
    if (c >= p.P) {
        cooperative_groups::this_cluster().sync();
        cooperative_groups::this_cluster().sync();
        return;
    }
    const bool big = c < p.n_big;
    const int r0 = big ? c * NR : p.n_big * NR + (c - p.n_big) * (NR - 1);
    if (big) {
        if (edge) { if (r0 & 1) deep_body<NR, 1, true, TR>(p, r0); else deep_body<NR, 0, true, TR>(p, r0); }
        else      { if (r0 & 1) deep_body<NR, 1, false, TR>(p, r0); else deep_body<NR, 0, false, TR>(p, r0); }
    } else {
        if (edge) { if (r0 & 1) deep_body<NR - 1, 1, true, TR>(p, r0); else deep_body<NR - 1, 0, true, TR>(p, r0); }
        else      { if (r0 & 1) deep_body<NR - 1, 1, false, TR>(p, r0); else deep_body<NR - 1, 0, false, TR>(p, r0); }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int resident_slots() { return RES2_SLOTS; }

// rows per CTA for a Wk x Hk kernel grid (0: does not fit on chip)
static int resident_rows(const pcd_solver *s, int Wk, int Hk, bool deep_only) {
    if ((Wk + 1) / 2 > RES_NT) return 0;
    int nr = (Hk + s->sm_count - 1) / s->sm_count;
    // 8 or 9 rows per CTA: only the deep-halo kernel (even W); NaN holes then go to the large-grid paths (run_resident)
    if (nr > RES_NR_DEEP_MAX || ((nr > RES_NR_MAX || deep_only) && Wk % 2 != 0)) return 0;
    // short grids: fewer CTAs with three rows each rather than one or two rows on every SM, so that the kernel with one
    // exchange per sweep applies (it needs even W and slabs of 3 / 2 rows; measured: three rows per CTA 1.18 us/sweep at
    // 400^2 against 1.49 for two rows per CTA with an exchange per colour phase at 300 x 157)
    if (nr < 3 && Wk % 2 == 0 && Hk >= 3) nr = 3;
    if (deep_only && nr < 3) return 0;
    return nr;
}

int resident_plan(pcd_solver *s) {
    // Wide grids (more than 1024 columns) whose HEIGHT fits the thread layout are solved transposed, by the deep-halo
    // kernel only (its summation order is a template switch, sor_cell_o): 1280 x 720 runs as 720 x 1280.
    s->res_tr = false;
    int nr = resident_rows(s, s->W, s->H, false);
    if (!nr) {
        nr = resident_rows(s, s->H, s->W, true);
        if (!nr) return 0;
        s->res_tr = true;
    }
    const int H = s->res_tr ? s->W : s->H;   // rows of the kernel's grid
    // rows are split as evenly as possible: P slabs, the first n_big of nr rows and the rest of nr-1, so no slab has
    // phantom rows (a partly filled slab would run the per-cell masked path and set the pace of the whole chain:
    // 1000 x 1000 took 4.98 us/sweep with one 6-of-7 slab against 3.0 us for 1024 x 1024)
    const int pick = nr;
    const int P = (H + pick - 1) / pick;
    s->res_n_big = P - (P * pick - H);
    const int Kp = RES_KP;
    s->res_ctas = P;
    s->res_rows_per_cta = pick;
    // phi (and, in the deep-halo kernel, D) of the slab + the slots of the links inside a cluster
    s->res_smem = (size_t)pick * 2 * Kp * sizeof(double) * (DEEP_DSMEM ? 2 : 1) + (size_t)2 * RES2_SLOTS * sizeof(uint4);
    s->res_threads = RES_NT;
    return 1;
}

// out[x][y] = in[y][x] for a W x H row-major field (32 x 32 tiles through shared memory, both sides coalesced)
__global__ void transpose_kernel(const double *__restrict__ in, double *__restrict__ out, int W, int H) {
    __shared__ double tile[32][33];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int x = x0 + threadIdx.x, y = y0 + j;
        if (x < W && y < H) tile[j][threadIdx.x] = in[(size_t)y * W + x];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int y = y0 + threadIdx.x, x = x0 + j;
        if (x < W && y < H) out[(size_t)x * H + y] = tile[threadIdx.x][j];
    }
}

static int transpose_field(const double *in, double *out, int W, int H, cudaStream_t stream) {
    transpose_kernel<<<dim3((W + 31) / 32, (H + 31) / 32), dim3(32, 8), 0, stream>>>(in, out, W, H);
    PCD_LAUNCHED();
    return PCD_OK;
}

template <int NR>
static int launch_resident(pcd_solver *s, ResParams &prm, bool deep) {
    void (*kernel)(const ResParams) = nullptr;
    if constexpr (NR <= RES_NR_MAX) kernel = sor_resident_kernel<NR>;
    if constexpr (NR >= 3) {
        if (deep) kernel = s->res_tr ? sor_resident_deep_kernel<NR, true> : sor_resident_deep_kernel<NR, false>;
    }
    if (!kernel) { set_error("resident K-SOR: no kernel for %d rows per CTA without deep halos", NR); return PCD_ERR_UNSUPPORTED; }
    PCD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->res_smem));
    // preferred: clusters of two CTAs (one TPC each); the link inside a pair then runs over DSMEM and only every other
    // link crosses L2.  This launch carries the cluster attribute only (cluster + cooperative is refused under the
    // profiler): with one CTA per SM and at most one CTA per SM in the grid all pairs are resident on an idle device,
    // which cudaOccupancyMaxActiveClusters confirms up front; should they ever not be, the kernel gives up after ~1 s
    // (ResState::error), leaves phi untouched, and run_resident repeats the launch the cooperative way.
    // (Measured at the end of round 2, deep-halo kernel, pairs / plain cooperative launch (PCD_RES_NO_PAIRS=1): 1024^2
    // 1.557 / 1.552, 1024 x 512 1.543 / 1.436, 400^2 1.184 / 1.169 us/sweep -- the pairs no longer pay now that the
    // exchange does not pace the kernel.  Making the cooperative launch the default was tried and taken back: the
    // exchange-per-phase kernel hung on an odd-width grid (1023 x 1024) in that mode, with no GPU time left to find out why;
    // the pair launch is the configuration every test and soak of both rounds ran.)
    static const bool no_pairs = getenv("PCD_RES_NO_PAIRS") != nullptr;
    static const int cs_env = getenv("PCD_RES_CLUSTER") ? atoi(getenv("PCD_RES_CLUSTER")) : 2;   // tuning knob
    const int cs = (cs_env == 4 || cs_env == 8) ? cs_env : 2;
    if (s->res_pairs >= 0 && !no_pairs && s->res_ctas >= 2) {
        const int grid = (s->res_ctas + cs - 1) / cs * cs;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(RES_NT); cfg.dynamicSmemBytes = s->res_smem; cfg.stream = s->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (s->res_pairs == 0) {   // decide once per solver
            int n = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
            s->res_pairs = (e == cudaSuccess && n * cs >= grid) ? 1 : -1;
            if (e != cudaSuccess) cudaGetLastError();
        }
        if (s->res_pairs == 1) {
            prm.pair = cs;
            cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, prm);
            if (e == cudaSuccess) { PCD_LAUNCHED(); return PCD_OK; }
            cudaGetLastError();
            s->res_pairs = -1;
        }
    }
    prm.pair = 0;
    void *args[] = {&prm};
    PCD_CUDA(cudaLaunchCooperativeKernel((void *)kernel, dim3(s->res_ctas), dim3(RES_NT), args,
                                         s->res_smem, s->stream));
    PCD_LAUNCHED();
    return PCD_OK;
}

int run_resident(pcd_solver *s, const double *D, double *phi, int max_it, double tol, pcd_solve_info *info) {
    const bool tr = s->res_tr;
    const int W = tr ? s->H : s->W, H = tr ? s->W : s->H;   // the kernel's grid (the transposed one for wide grids)
    const int P = s->res_ctas;
    const int lag = s->check_lag > 0 ? s->check_lag : 4;  // >= 1: a sweep's slot is complete only after every CTA left it
    int done = 0, conv = 0;
    double last = 0.0;
    unsigned long long *g_max = s->sweep_max;
    unsigned long long *g_slot = s->sweep_max + s->ring;
    // one exchange per sweep (deep halos) wherever its preconditions hold; a NaN in D is only found by the kernel itself
    // Measured (tools/res_time.py, us per sweep deep / per-phase): slabs of 7 rows 1.65 / 2.35, 6 rows 1.72 / 2.12, 5 rows
    // 1.64 / 1.97, 4 rows 1.86 / 1.71 (1024 x 512), 1.69 / 1.61 (512^2), 3 rows 1.64 / 1.58 (400^2): with fewer than five
    // rows there is too little interior work to cover the one long exchange, and two short ones win.
    const bool no_deep = getenv("PCD_RES_NO_DEEP") != nullptr;   // read per solve: the tests compare both kernels
    const char *min_rows_env = getenv("PCD_RES_DEEP_MIN_ROWS");  // tests: 3 = wherever the kernel is valid
    const int min_rows = min_rows_env ? (atoi(min_rows_env) < 3 ? 3 : atoi(min_rows_env)) : 3;
    bool deep = !no_deep && W % 2 == 0 && s->res_rows_per_cta >= min_rows;
    const bool deep_only = s->res_rows_per_cta > RES_NR_MAX || tr;   // strips too tall for the exchange-per-phase kernel, or transposed
    if (deep_only && !deep) return PCD_RES_FALLBACK;
    if (tr) {   // the kernel works on transposed copies; phi itself is only written after a completed solve
        const size_t bytes = sizeof(double) * (size_t)W * H;
        if (!s->tr_D) PCD_CUDA(cudaMalloc(&s->tr_D, bytes));
        if (!s->tr_phi) PCD_CUDA(cudaMalloc(&s->tr_phi, bytes));
        PCD_TRY(transpose_field(D, s->tr_D, s->W, s->H, s->stream));
        PCD_TRY(transpose_field(phi, s->tr_phi, s->W, s->H, s->stream));
        info->launches += 2;
    }
    const double *Dk = tr ? s->tr_D : D;
    double *phik = tr ? s->tr_phi : phi;
    while (done < max_it && !conv) {
        const int k = max_it - done < RES_MAX_SWEEPS_PER_LAUNCH ? max_it - done : RES_MAX_SWEEPS_PER_LAUNCH;
        PCD_CUDA(cudaMemsetAsync(g_max, 0, sizeof(unsigned long long) * (size_t)k, s->stream));
        PCD_CUDA(cudaMemsetAsync(g_slot, 0, sizeof(unsigned long long) * (size_t)k, s->stream));
        PCD_CUDA(cudaMemsetAsync(s->halo, 0, (size_t)P * 2 * RES2_SLOTS * sizeof(uint4), s->stream));
        PCD_CUDA(cudaMemsetAsync(s->res_state, 0, sizeof(ResState), s->stream));
        ResParams prm;
        prm.phi = phik; prm.D = Dk; prm.W = W; prm.H = H; prm.K = (W + 1) / 2;
        prm.Kp = RES_KP;
        prm.P = P; prm.n_big = s->res_n_big; prm.nr_big = s->res_rows_per_cta; prm.max_it = k; prm.lag = lag; prm.tol = tol; prm.w = make_w(s->W);   // omega follows the GRID's width (src/solver.cpp:71), transposed or not
        prm.ll = (uint4 *)s->halo; prm.g_max = g_max; prm.g_slot = g_slot; prm.state = (ResState *)s->res_state;
        int rc;
        PCD_CUDA(cudaEventRecord(s->evk0, s->stream));
        switch (s->res_rows_per_cta) {
            case 1: rc = launch_resident<1>(s, prm, deep); break;
            case 2: rc = launch_resident<2>(s, prm, deep); break;
            case 3: rc = launch_resident<3>(s, prm, deep); break;
            case 4: rc = launch_resident<4>(s, prm, deep); break;
            case 5: rc = launch_resident<5>(s, prm, deep); break;
            case 6: rc = launch_resident<6>(s, prm, deep); break;
            case 7: rc = launch_resident<7>(s, prm, deep); break;
            case 8: rc = launch_resident<8>(s, prm, deep); break;
            default: rc = launch_resident<9>(s, prm, deep); break;
        }
        PCD_TRY(rc);
        PCD_CUDA(cudaEventRecord(s->evk1, s->stream));
        info->launches++;
        PCD_CUDA(cudaMemcpyAsync(s->h_res_state, s->res_state, sizeof(ResState), cudaMemcpyDeviceToHost, s->stream));
        PCD_CUDA(cudaStreamSynchronize(s->stream));
        const ResState st = *(ResState *)s->h_res_state;
        if (st.error == RES_ERR_UNSUPPORTED && deep) {   // NaN holes: phi is untouched
            if (deep_only) return PCD_RES_FALLBACK;
            deep = false;
            continue;
        }
        if (st.error) {  // a CTA waited in vain: phi is untouched
            if (s->res_pairs == 1) { s->res_pairs = -1; continue; }   // once more, cooperatively
            set_error("resident K-SOR kernel: a CTA gave up waiting for its neighbour");
            return PCD_ERR_CUDA;
        }
        {
            float kms = 0.f;
            PCD_CUDA(cudaEventElapsedTime(&kms, s->evk0, s->evk1));
            info->kernel_ms += kms;
        }
        // the device acts on the test with a lag: when the cap ends the launch, the last `lag` sweeps
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
        s->res_exchange = deep ? 2 : 1;
    }
    if (tr) {
        PCD_TRY(transpose_field(s->tr_phi, phi, W, H, s->stream));
        info->launches++;
    }
    info->sweeps = done;
    info->converged_at = conv;
    info->last_max_update = last;
    return PCD_OK;
}

}  // namespace pcd
