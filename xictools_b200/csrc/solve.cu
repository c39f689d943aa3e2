// solve.cu -- the linear solve of one Newton tick for every trial that asked for one:
//
//   k_assemble  sCKT::load's accumulation + nodeset/ic pinning + loadGmin (ckt.cc:1177-1263,
//               spspice.cc:259-300).  contrib[NC][ldt] (trial fastest, written by the load
//               kernels in ASSEMBLY ORDER: the terms of one matrix entry / rhs row are
//               consecutive rows, in the reference's load order) -> image[trial][img], one
//               contiguous A | b | scratch block per trial.  lane = trial, so every read is a
//               256-B line; the transpose goes through shared memory so every write is too.
//               No atomics: a value is the left-to-right sum of its run.
//   k_lu        spFactor/klu_refactor + spSolve/klu_solve with the frozen pivot order
//               (kluif.cc:594-628).  ONE WARP PER TRIAL: the image is pulled into shared
//               memory by a bulk copy, the 32 lanes interpret the level-scheduled task program
//               (b200_internal.h) with __syncwarp between levels -- no CTA barrier anywhere.
//               Persistent CTAs; warps take trials from a work counter over the live list.
//   k_finish    unpermute (ysol[trial][k] -> X[node][trial], transposed through shared
//               memory) + the node part of NIconvTest (niconv.cc:58-91) + per-trial flags.
#include <algorithm>
#include <cuda_runtime.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "device_state.h"
#include "dev_load.cuh"
#include "kernels.h"

// ---- 1-D bulk copy (TMA engine) global -> shared, completion on an mbarrier ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    // dst, src 16-byte aligned, bytes a multiple of 16
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// ---------------------------------------------------------------------------------------
// k_assemble
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double asm_finalize(double v, int info, const Consts &K, const SolveView &Sv,
    bool ns_on, bool ic_on, double diagGmin)
{
    int node = info & 0x00ffffff;
    if (!node) return v;
    if (info & ASM_DIAG) {
        // ckt.cc:1177-1204 nodeset / ic pinning, then loadGmin (ckt.cc:1221-1263, spspice.cc:259-300)
        if (ns_on && Sv.ns_given[node]) v += 1.0;
        if (ic_on && Sv.ic_given[node]) v += 1.0;
        if (info & ASM_VNODE) {
            if (diagGmin != 0.0)
                v += diagGmin;
            else if (K.opt.forcegmin && (v >= 0.0 && v < K.opt.gmin))
                v = K.opt.gmin;
            else if (v > K.opt.gmax)
                v = K.opt.gmax;
        }
    }
    else {
        if (ns_on && Sv.ns_given[node]) v += Sv.ns_val[node];
        if (ic_on && Sv.ic_given[node]) v += Sv.ic_val[node];
    }
    return v;
}

// left-to-right sum of len rows (pitch ldt) with U loads in flight
template <int U>
__device__ __forceinline__ double run_sum(double v, const double * __restrict__ x, size_t ldt, int &k, int len)
{
    for (; k + U <= len; k += U) {
        double xx[U];
#pragma unroll
        for (int u = 0; u < U; u++) xx[u] = __ldcs(x + (size_t)(k + u)*ldt);
#pragma unroll
        for (int u = 0; u < U; u++) v += xx[u];
    }
    return v;
}

// One WARP = 32 adjacent trials (lane = trial) x one tile of ASM_TS consecutive image slots; the
// warps of a CTA share nothing (no CTA barrier: a warp that owns a long run -- the supply node's
// diagonal collects a term from every MOSFET -- must not hold seven others at a barrier).
// Tiles are visited in the order of the host-built list, heaviest first.
#define ASM_TS 16               // image slots per tile (one 128-B line of a trial's image row)
#define ASM_WARPS 8
__global__ void __launch_bounds__(ASM_WARPS*32)
k_assemble(Consts K, HotView hot, VecView V, SolveView Sv)
{
    __shared__ double tiles[ASM_WARPS][ASM_TS][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x*32 + lane;
    // Sv.cls: -1 every solving trial; 1 only trials whose load was in a DC mode, 0 only the
    // transient ones (two programs, b200_build_programs)
    const bool live = (t < K.ntrials) && hot.do_solve[t] &&
        (Sv.cls < 0 || ((hot.mode[t] & MODEDC) != 0) == (Sv.cls != 0));
    const unsigned mask = __ballot_sync(0xffffffffu, live);
    if (!mask) return;                                   // same 32 trials in every warp: uniform exit
    if (blockIdx.y == 0 && warp == 0) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&Sv.ctr[0], __popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (live) Sv.live[base + __popc(mask & ((1u << lane) - 1u))] = t;
    }
    const int ti = blockIdx.y*ASM_WARPS + warp;
    if (ti >= Sv.ntiles) return;
    const int s0 = Sv.tile_order[ti]*ASM_TS;
    const int cnt = min(ASM_TS, Sv.img - s0);
    const int mode = live ? hot.mode[t] : 0;
    const double diagGmin = live ? hot.diaggmin[t] : 0.0;
    const bool ns_on = (mode & MODEDC) && (mode & (MODEINITJCT | MODEINITFIX));
    const bool ic_on = (mode & MODEDC) && (mode & MODETRANOP) && !(mode & MODEUIC);
    const size_t ldt = (size_t)K.ldt;
    const double *cb = V.contrib + t;
    double (*tile)[33] = tiles[warp];
    for (int sl = 0; sl < cnt; sl++) {
        const int4 run = __ldg(reinterpret_cast<const int4*>(Sv.runs + s0 + sl));
        double v = 0.0;
        if (live) {
            const double *x = cb + (size_t)run.x*ldt;
            const int len = run.y;
            int k = 0;
            if (len >= 64) v = run_sum<32>(v, x, ldt, k, len);
            v = run_sum<8>(v, x, ldt, k, len);
            for (; k < len; k++) v += __ldcs(x + (size_t)k*ldt);
            v = asm_finalize(v, run.z, K, Sv, ns_on, ic_on, diagGmin);
        }
        tile[sl][lane] = v;
    }
    __syncwarp();
    // two trial rows per store instruction: lanes 0..15 one row, lanes 16..31 the next
    const int half = lane >> 4, c = lane & 15;
    for (int tr = half; tr < 32; tr += 2) {
        if (!((mask >> tr) & 1u) || c >= cnt) continue;
        Sv.image[(size_t)(blockIdx.x*32 + tr)*(size_t)Sv.img + s0 + c] = tile[c][tr];
    }
}

// long runs, first step: one warp = 32 adjacent trials x one chunk (b200_internal.h ASM_LONG)
__global__ void __launch_bounds__(ASM_WARPS*32)
k_assemble_long(Consts K, HotView hot, VecView V, SolveView Sv)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x*32 + lane;
    const int c = blockIdx.y*ASM_WARPS + warp;
    if (c >= Sv.nlong) return;
    const bool live = (t < K.ntrials) && hot.do_solve[t] &&
        (Sv.cls < 0 || ((hot.mode[t] & MODEDC) != 0) == (Sv.cls != 0));
    if (!live) return;
    const size_t ldt = (size_t)K.ldt;
    const AsmChunk ch = Sv.long_chunks[c];
    const double *x = V.contrib + t + (size_t)ch.pos*ldt;
    double v = 0.0;
    int k = 0;
    v = run_sum<32>(v, x, ldt, k, ch.len);
    for (; k < ch.len; k++) v += __ldcs(x + (size_t)k*ldt);
    V.contrib[((size_t)Sv.NC + c)*ldt + t] = v;
}

// ---------------------------------------------------------------------------------------
// k_lu
// ---------------------------------------------------------------------------------------
#define LU_MAXW 8
template <bool PSMEM>
__global__ void __launch_bounds__(LU_MAXW*32, 1)
k_lu(SolveView Sv, int n)
{
    extern __shared__ double smem[];
    __shared__ unsigned long long s_bar[LU_MAXW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = Sv.lu_warps;
    const int img = Sv.img;
    double *S = smem + (size_t)warp*img;
    unsigned short *s_round = reinterpret_cast<unsigned short*>(smem + (size_t)W*img);
    unsigned *s_rows = reinterpret_cast<unsigned*>(s_round + ((Sv.nrounds + 7) & ~7));
    const int nlive = Sv.ctr[0];
    if ((int)(blockIdx.x*W) >= nlive) return;           // the CTAs before this one have a warp per live trial
    for (int i = threadIdx.x; i < Sv.nrounds; i += blockDim.x) s_round[i] = Sv.rounds[i];
    if (PSMEM) {
        const uint4 *src = reinterpret_cast<const uint4*>(Sv.rows);
        uint4 *dst = reinterpret_cast<uint4*>(s_rows);
        for (int i = threadIdx.x; i < Sv.nrows*(LU_LANES/4); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x < W) mbar_init(&s_bar[threadIdx.x], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const unsigned *rows0 = (PSMEM ? s_rows : Sv.rows) + lane;
    const unsigned bytes = (unsigned)img*8u;
    unsigned phase = 0;
    const int nrounds = Sv.nrounds;
    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(&Sv.ctr[1], 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= nlive) break;
        const int t = Sv.live[i];
        if (lane == 0) {
            // the previous trial's image was read and written through the generic proxy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s_bar[warp], bytes);
            bulk_g2s(S, Sv.image + (size_t)t*img, bytes, &s_bar[warp]);
        }
        mbar_wait(&s_bar[warp], phase);
        phase ^= 1u;
        int bad = 0;
        const unsigned *rp = rows0;
        for (int r = 0; r < nrounds; r++) {
            const unsigned meta = s_round[r];
            const int nsteps = (int)(meta & 0x7fffu);
            if (meta & LU_RSYNC) __syncwarp();
            const unsigned h = PSMEM ? rp[0] : __ldg(rp);
            rp += LU_LANES;
            const unsigned dst = h & 0xffffu, div = h >> 16;
            double acc = S[dst];
            for (int s = 0; s < nsteps; s += 4) {
                unsigned w[4];
                double a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (s + u < nsteps) w[u] = PSMEM ? rp[(s + u)*LU_LANES] : __ldg(rp + (s + u)*LU_LANES);
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (s + u < nsteps) { a[u] = S[w[u] & 0xffffu]; b[u] = S[w[u] >> 16]; }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (s + u < nsteps) acc -= a[u]*b[u];
            }
            rp += nsteps*LU_LANES;
            if (div != LU_NONE) {
                const double piv = S[div];
                if (piv == 0.0 || !isfinite(piv)) bad = 1;
                acc = acc/piv;
            }
            S[dst] = acc;
        }
        __syncwarp();
        double *y = Sv.ysol + (size_t)t*Sv.npad;
        const double *Sy = S + Sv.nlu;
        for (int k = lane; k < n; k += 32) y[k] = Sy[k];
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) Sv.lu_bad[t] = bad;
        __syncwarp();                                    // every lane is done with S before the next copy lands
    }
}

// ---------------------------------------------------------------------------------------
// k_lu_team: k_lu with P warps per resident image.  One warp per image leaves the SM's issue
// slots idle (RO-501: 5 images fit, a row of the program is a chain of ~150 dependent cycles);
// here the P warps of a team take the rounds of a level in turn (round r of the level -> warp
// r mod P) and meet at a named barrier of their own between levels, so the other teams of the
// CTA never wait.  Image by one bulk copy as in k_lu; program rows addressed through the
// round -> row table.  Same tasks, same operand order: bit-identical results.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void team_sync(int team, int nthreads)
{
    asm volatile("bar.sync %0, %1;" :: "r"(1 + team), "r"(nthreads) : "memory");
}

template <bool PSMEM, int P>
__global__ void __launch_bounds__(LU_MAXW*P*32, 1)
k_lu_team(SolveView Sv, int n)
{
    extern __shared__ double smem[];
    __shared__ unsigned long long s_bar[LU_MAXW];
    __shared__ int s_next[LU_MAXW], s_badf[LU_MAXW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int team = warp/P, h = warp%P;
    const int W = Sv.lu_warps;                          // teams (= resident images) per CTA
    const int img = Sv.img;
    double *S = smem + (size_t)team*img;
    unsigned short *s_round = reinterpret_cast<unsigned short*>(smem + (size_t)W*img);
    unsigned *s_rows = reinterpret_cast<unsigned*>(s_round + ((Sv.nrounds + 7) & ~7));
    const int nlive = Sv.ctr[0];
    if ((int)(blockIdx.x*W) >= nlive) return;
    for (int i = threadIdx.x; i < Sv.nrounds; i += blockDim.x) s_round[i] = Sv.rounds[i];
    if (PSMEM) {
        const uint4 *src = reinterpret_cast<const uint4*>(Sv.rows);
        uint4 *dst = reinterpret_cast<uint4*>(s_rows);
        for (int i = threadIdx.x; i < Sv.nrows*(LU_LANES/4); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x < W) mbar_init(&s_bar[threadIdx.x], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const unsigned *rows0 = (PSMEM ? s_rows : Sv.rows) + lane;
    const unsigned bytes = (unsigned)img*8u;
    unsigned phase = 0;
    const int nlevels = Sv.nlevels;
    for (;;) {
        team_sync(team, P*32);                           // the team is done with the previous image
        if (h == 0 && lane == 0) {
            s_next[team] = atomicAdd(&Sv.ctr[1], 1);
            s_badf[team] = 0;
        }
        team_sync(team, P*32);
        const int i = s_next[team];
        if (i >= nlive) break;
        const int t = Sv.live[i];
        if (h == 0 && lane == 0) {
            // the previous trial's image was read and written through the generic proxy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s_bar[team], bytes);
            bulk_g2s(S, Sv.image + (size_t)t*img, bytes, &s_bar[team]);
        }
        mbar_wait(&s_bar[team], phase);
        phase ^= 1u;
        int bad = 0;
        for (int lv = 0; lv < nlevels; lv++) {
            const int r0 = __ldg(Sv.level_ptr + lv), r1 = __ldg(Sv.level_ptr + lv + 1);
            for (int r = r0 + h; r < r1; r += P) {
                const int nsteps = (int)(s_round[r] & 0x7fffu);
                const unsigned *rp = rows0 + (size_t)__ldg(Sv.round_row + r)*LU_LANES;
                const unsigned hd = PSMEM ? rp[0] : __ldg(rp);
                rp += LU_LANES;
                const unsigned dst = hd & 0xffffu, div = hd >> 16;
                double acc = S[dst];
                for (int s = 0; s < nsteps; s += 4) {
                    unsigned w[4];
                    double a[4], b[4];
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (s + u < nsteps) w[u] = PSMEM ? rp[(s + u)*LU_LANES] : __ldg(rp + (s + u)*LU_LANES);
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (s + u < nsteps) { a[u] = S[w[u] & 0xffffu]; b[u] = S[w[u] >> 16]; }
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (s + u < nsteps) acc -= a[u]*b[u];
                }
                if (div != LU_NONE) {
                    const double piv = S[div];
                    if (piv == 0.0 || !isfinite(piv)) bad = 1;
                    acc = acc/piv;
                }
                S[dst] = acc;
            }
            team_sync(team, P*32);                       // the level is complete
        }
        double *y = Sv.ysol + (size_t)t*Sv.npad;
        const double *Sy = S + Sv.nlu;
        for (int k = h*32 + lane; k < n; k += P*32) y[k] = Sy[k];
        if (bad) s_badf[team] = 1;
        team_sync(team, P*32);
        if (h == 0 && lane == 0) Sv.lu_bad[t] = s_badf[team];
    }
}

// ---------------------------------------------------------------------------------------
// k_lu_wide: the same task program for images beyond one warp's shared-memory window (a few
// thousand unknowns upward).  The image stays where k_assemble put it -- global memory, in
// practice L2 -- and the program rows carry 32-bit slot numbers.  A large system has wide
// levels (50 000 unknowns: ~26 000 rounds in ~50 levels), so here one CTA works on one trial:
// its warps share out the rounds of a level (round r of the level goes to warp r mod W) and
// meet at a CTA barrier between levels, which also orders the global-memory writes of one
// level before the reads of the next.  Same tasks, same operand order inside every task, so
// the arithmetic is bit-identical to k_lu's.  Persistent CTAs pull live trials from the work
// counter.
// ---------------------------------------------------------------------------------------
#define LUW_MAXW 32
__global__ void __launch_bounds__(LUW_MAXW*32)
k_lu_wide(SolveView Sv, int n)
{
    __shared__ int s_next;
    __shared__ int s_bad;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = blockDim.x >> 5;
    const int nlive = Sv.ctr[0];
    const int img = Sv.img;
    const int nlevels = Sv.nlevels;
    const uint2 *rows = reinterpret_cast<const uint2*>(Sv.rows);
    for (;;) {
        __syncthreads();                                 // everyone is done with the previous trial (and s_next)
        if (threadIdx.x == 0) { s_next = atomicAdd(&Sv.ctr[1], 1); s_bad = 0; }
        __syncthreads();
        const int i = s_next;
        if (i >= nlive) break;
        const int t = Sv.live[i];
        double *S = Sv.image + (size_t)t*img;
        int bad = 0;
        for (int lv = 0; lv < nlevels; lv++) {
            const int r0 = __ldg(Sv.level_ptr + lv), r1 = __ldg(Sv.level_ptr + lv + 1);
            for (int r = r0 + warp; r < r1; r += W) {
                const int nsteps = (int)(__ldg(Sv.rounds + r) & 0x7fffu);
                const uint2 *rp = rows + (size_t)__ldg(Sv.round_row + r)*LU_LANES + lane;
                const uint2 h = __ldg(rp);
                rp += LU_LANES;
                double acc = S[h.x];
                for (int s = 0; s < nsteps; s += 4) {
                    uint2 w[4];
                    double a[4], b[4];
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (s + u < nsteps) w[u] = __ldg(rp + (s + u)*LU_LANES);
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (s + u < nsteps) { a[u] = S[w[u].x]; b[u] = S[w[u].y]; }
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (s + u < nsteps) acc -= a[u]*b[u];
                }
                if (h.y != LU_NONE_W) {
                    const double piv = S[h.y];
                    if (piv == 0.0 || !isfinite(piv)) bad = 1;
                    acc = acc/piv;
                }
                S[h.x] = acc;
            }
            __syncthreads();
        }
        double *y = Sv.ysol + (size_t)t*Sv.npad;
        const double *Sy = S + Sv.nlu;
        for (int k = threadIdx.x; k < n; k += blockDim.x) y[k] = Sy[k];
        if (bad) s_bad = 1;
        __syncthreads();
        if (threadIdx.x == 0) Sv.lu_bad[t] = s_bad;
    }
}

// ---------------------------------------------------------------------------------------
// k_finish: CTA = one group of 32 adjacent trials, all n unknowns
// ---------------------------------------------------------------------------------------
#define FIN_WARPS 8
// what a trial's solve leaves for the control kernel
__device__ __forceinline__ void finish_trial(const HotView &hot, int t, int fl)
{
    const int err = (fl & 1) ? B200_E_SINGULAR : ((fl & 2) ? B200_E_MATHOVF : 0);
    const int fail = (fl & 4) ? 1 : 0;
    const int mode = hot.mode[t];
    hot.solve_err[t] = err;
    hot.nodeconv[t] = fail;
    hot.devconv[t] = 0;
    const int it1 = hot.iterno[t] + 1;
    hot.need_devconv[t] = (!err && !fail && (mode & (MODEINITFIX | MODEINITFLOAT)) &&
        hot.noncon[t] == 0 && it1 != 1) ? 1 : 0;
}

// CHUNKED = false: one CTA walks all n unknowns of its 32 trials and finishes them.
// CHUNKED = true (few trial groups, or many unknowns: the 32-trial groups alone would not fill
// the GPU): grid.y CTAs share the unknowns, kchunk each, and leave their flags in Sv.fin_flags;
// k_finish_tail closes the trials.
template <bool CHUNKED>
__global__ void __launch_bounds__(FIN_WARPS*32)
k_finish(Consts K, HotView hot, SolveView Sv, const double * __restrict__ Xold, double * __restrict__ Xnew, int kchunk)
{
    __shared__ double tile[2][32][33];
    __shared__ int s_flags[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t0 = blockIdx.x*32;
    const int t = t0 + lane;
    const bool live = (t < K.ntrials) && hot.do_solve[t] &&
        (Sv.cls < 0 || ((hot.mode[t] & MODEDC) != 0) == (Sv.cls != 0));
    const unsigned mask = __ballot_sync(0xffffffffu, live);
    if (!mask) return;
    if (threadIdx.x < 32) s_flags[threadIdx.x] = 0;
    const int n = K.n;
    const size_t ldt = (size_t)K.ldt;
    int fl = 0;
    int buf = 0;
    const int kbeg = CHUNKED ? blockIdx.y*kchunk : 0;
    const int kend = CHUNKED ? min(n, kbeg + kchunk) : n;
    for (int k0 = kbeg; k0 < kend; k0 += 32, buf ^= 1) {
        // rows of ysol: trial tr, unknowns k0 .. k0+31 (one 256-B line)
        for (int tr = warp; tr < 32; tr += FIN_WARPS) {
            double v = 0.0;
            if (((mask >> tr) & 1u) && k0 + lane < n) v = Sv.ysol[(size_t)(t0 + tr)*Sv.npad + k0 + lane];
            tile[buf][tr][lane] = v;
        }
        __syncthreads();
        for (int kk = warp; kk < 32; kk += FIN_WARPS) {
            const int k = k0 + kk;
            if (k >= n || !live) continue;
            const int node = Sv.yperm_out[k];
            const double cur = tile[buf][lane][kk];
            const double old = Xold[(size_t)node*ldt + t];
            Xnew[(size_t)node*ldt + t] = cur;
            if (!isfinite(cur)) fl |= 2;
            const double dd = fabs(old - cur);
            const double tol = K.opt.reltol*DMAX(fabs(old), fabs(cur)) +
                (Sv.ntype[node] == 0 ? K.opt.vntol : K.opt.abstol);
            if (dd > tol) fl |= 4;
        }
        // the other buffer is refilled next; it was last read two iterations ago
    }
    if (fl) atomicOr(&s_flags[lane], fl);
    __syncthreads();
    if (warp == 0 && live) {
        if (CHUNKED) {
            if (s_flags[lane]) atomicOr(&Sv.fin_flags[t], s_flags[lane]);
        }
        else {
            fl = s_flags[lane] | (Sv.lu_bad[t] ? 1 : 0);
            Xnew[t] = 0.0;                               // ground slot
            finish_trial(hot, t, fl);
        }
    }
}

__global__ void k_finish_tail(Consts K, HotView hot, SolveView Sv, double * __restrict__ Xnew)
{
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= K.ntrials || !hot.do_solve[t]) return;
    if (!(Sv.cls < 0 || ((hot.mode[t] & MODEDC) != 0) == (Sv.cls != 0))) return;
    const int fl = Sv.fin_flags[t] | (Sv.lu_bad[t] ? 1 : 0);
    Sv.fin_flags[t] = 0;
    Xnew[t] = 0.0;                                       // ground slot
    finish_trial(hot, t, fl);
}

// ---------------------------------------------------------------------------------------
// launch side
// ---------------------------------------------------------------------------------------
int solve_configure(SolveView &Sv)
{
    int dev = 0, nsm = 0, maxsm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return B200_E_NOGPU;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (Sv.ntiles != (Sv.img + ASM_TS - 1)/ASM_TS) return B200_E_PANIC;     // b200_internal.h ASM_TILE_SLOTS
    Sv.nsm = nsm;
    if (Sv.wide) {
        // image in global memory: no window to size; as many warps per SM as the registers allow
        // warps per trial: what an average level can keep busy
        int w = Sv.nlevels > 0 ? (Sv.nrounds + Sv.nlevels - 1)/Sv.nlevels : 1;
        if (const char *e = getenv("B200_LUW_WARPS")) w = atoi(e);      // measurement override
        w = w >= 24 ? 32 : w >= 12 ? 16 : w >= 6 ? 8 : 4;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lu_wide, w*32, 0) != cudaSuccess || per_sm < 1)
            return B200_E_NOGPU;
        Sv.prog_in_smem = 0;
        Sv.lu_warps = w;
        Sv.lu_smem = 0;
        Sv.lu_grid = nsm*per_sm;
        return B200_OK;
    }
    cudaFuncAttributes fa, ft;
    if (cudaFuncGetAttributes(&fa, k_lu<true>) != cudaSuccess ||
        cudaFuncGetAttributes(&ft, k_lu_team<true, 4>) != cudaSuccess) return B200_E_NOGPU;
    const size_t budget = (size_t)maxsm - std::max(fa.sharedSizeBytes, ft.sharedSizeBytes);
    const size_t imgb = (size_t)Sv.img*8;
    const size_t rnd = (size_t)((Sv.nrounds + 7) & ~7)*sizeof(unsigned short);
    const size_t prog = (size_t)Sv.nrows*LU_LANES*sizeof(unsigned);
    if (imgb + rnd > budget) return B200_E_UNSUPP;
    int w_glob = (int)((budget - rnd)/imgb);
    if (w_glob > LU_MAXW) w_glob = LU_MAXW;
    int w_smem = (budget > rnd + prog) ? (int)((budget - rnd - prog)/imgb) : 0;
    if (w_smem > LU_MAXW) w_smem = LU_MAXW;
    // the program in shared memory costs trials in flight; worth it while it costs at most a third
    const char *force = getenv("B200_LU_PROG");        // "smem" | "global": measurement override
    bool psmem = (w_smem >= 1) && (3*w_smem >= 2*w_glob);
    if (force && !strcmp(force, "global")) psmem = false;
    if (force && !strcmp(force, "smem") && w_smem >= 1) psmem = true;
    Sv.prog_in_smem = psmem ? 1 : 0;
    Sv.lu_warps = psmem ? w_smem : w_glob;
    Sv.lu_smem = (int)((size_t)Sv.lu_warps*imgb + rnd + (psmem ? prog : 0));
    Sv.lu_grid = nsm;
    // warps per resident image: 2 when the levels are wide enough to feed them
    Sv.lu_team = (Sv.nlevels > 0 && Sv.nrounds >= 2*Sv.nlevels) ? 2 : 1;
    if (const char *e = getenv("B200_LU_TEAM")) Sv.lu_team = atoi(e) == 2 ? 2 : (atoi(e) == 4 ? 4 : 1);   // measurement override
    // opt in to the whole window: a batch may run two programs of different size through the
    // same kernel, and the attribute is per function, not per launch
    cudaError_t e = psmem
        ? cudaFuncSetAttribute(k_lu<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget)
        : cudaFuncSetAttribute(k_lu<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    if (e == cudaSuccess && Sv.lu_team == 2)
        e = psmem ? cudaFuncSetAttribute(k_lu_team<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget)
                  : cudaFuncSetAttribute(k_lu_team<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    if (e == cudaSuccess && Sv.lu_team == 4)
        e = psmem ? cudaFuncSetAttribute(k_lu_team<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget)
                  : cudaFuncSetAttribute(k_lu_team<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    return e == cudaSuccess ? B200_OK : B200_E_NOMEM;
}

int launch_solve(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const SolveView &Sv, const double *Xold, double *Xnew)
{
    if (cudaMemsetAsync(Sv.ctr, 0, 2*sizeof(int), st) != cudaSuccess) return B200_E_NOGPU;
    const int ngroups = K.ldt/32;
    dim3 ga(ngroups, (Sv.ntiles + ASM_WARPS - 1)/ASM_WARPS);
    if (Sv.nlong)
        k_assemble_long<<<dim3(ngroups, (Sv.nlong + ASM_WARPS - 1)/ASM_WARPS), ASM_WARPS*32, 0, st>>>(K, hot, V, Sv);
    k_assemble<<<ga, ASM_WARPS*32, 0, st>>>(K, hot, V, Sv);
    if (Sv.wide) k_lu_wide<<<Sv.lu_grid, Sv.lu_warps*32, 0, st>>>(Sv, K.n);
    else if (Sv.lu_team == 2) {
        if (Sv.prog_in_smem) k_lu_team<true, 2><<<Sv.lu_grid, Sv.lu_warps*64, Sv.lu_smem, st>>>(Sv, K.n);
        else k_lu_team<false, 2><<<Sv.lu_grid, Sv.lu_warps*64, Sv.lu_smem, st>>>(Sv, K.n);
    }
    else if (Sv.lu_team == 4) {
        if (Sv.prog_in_smem) k_lu_team<true, 4><<<Sv.lu_grid, Sv.lu_warps*128, Sv.lu_smem, st>>>(Sv, K.n);
        else k_lu_team<false, 4><<<Sv.lu_grid, Sv.lu_warps*128, Sv.lu_smem, st>>>(Sv, K.n);
    }
    else if (Sv.prog_in_smem) k_lu<true><<<Sv.lu_grid, Sv.lu_warps*32, Sv.lu_smem, st>>>(Sv, K.n);
    else k_lu<false><<<Sv.lu_grid, Sv.lu_warps*32, Sv.lu_smem, st>>>(Sv, K.n);
    // share the unknowns among CTAs when the trial groups alone leave SMs idle: aim at ~4 CTAs per SM
    int nchunk = (4*Sv.nsm + ngroups - 1)/ngroups;
    nchunk = std::min(nchunk, (K.n + 63)/64);
    if (nchunk > 2) {                                    // (2 would trade a half-empty GPU for an extra launch)
        const int kchunk = (((K.n + nchunk - 1)/nchunk) + 31) & ~31;
        k_finish<true><<<dim3(ngroups, (K.n + kchunk - 1)/kchunk), FIN_WARPS*32, 0, st>>>(K, hot, Sv, Xold, Xnew, kchunk);
        k_finish_tail<<<(K.ntrials + 127)/128, 128, 0, st>>>(K, hot, Sv, Xnew);
    }
    else
        k_finish<false><<<ngroups, FIN_WARPS*32, 0, st>>>(K, hot, Sv, Xold, Xnew, 0);
    return 0;
}
