// kernels.cu -- CUDA kernels of the batched Newton engine (sm_100a, FP64, -fmad=false).
//
// Thread mappings:
//   k_load_*    one thread per (instance, trial), trial fastest: every global access of a warp
//               is 32 consecutive doubles of one [slot][trial] row -> one 256-B line.  Stamps go
//               to contrib[NC][ldt] in assembly order (no atomics).
//   k_ctl       one thread per trial: the flattened NIiter / op / tran-step state machine.
//   k_vec, k_statevec, k_devconv: one thread per (element, trial), gated by per-trial
//               request flags.
//   The solve of a tick (k_assemble -> k_lu -> k_finish) lives in solve.cu.
#include <algorithm>
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include "device_state.h"
#include "dev_load.cuh"
#include "dev_ctl.cuh"
#include "kernels.h"

#define PR(ref, i, t) ((ref).base[(size_t)(i)*(ref).ld + (size_t)(t)*(ref).ts])


// ---------------------------------------------------------------------------------------
__global__ void k_mos_temp(int n, int ntrials, int ldt, const MosCard *cards,
    RawParams raw, MosParamsOut out)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;
    if (t >= ntrials || i >= n) return;
    double r[MR_NRAW], o[MP_NPARAMS];
    for (int k = 0; k < MR_NRAW; k++) r[k] = PR(raw.p[k], i, t);
    dev_mos_temp(cards[i], r, o);
    for (int k = 0; k < MP_NPARAMS; k++) {
        if (out.ts[k])
            out.base[k][(size_t)i*ldt + t] = o[k];
        else if (t == 0)
            out.base[k][i] = o[k];
    }
}

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void make_ctx(TrialCtx &c, const Consts &K, const HotView &hot,
    const VecView &V, int t, int mode)
{
    c.mode = mode;
    c.order = hot.order[t];
    c.method = K.method;
    c.delta = hot.delta[t];
    c.ldt = K.ldt;
    c.trapratio = K.opt.trapratio;
    c.trapbad = &hot.trapbad[t];
    c.trapCheck = hot.trapcheck[t] && (mode & MODEINITFLOAT) && (mode & MODETRAN);
    c.ag[0] = hot.ag[t]; c.ag[1] = hot.ag[(size_t)K.ldt + t];
    c.agp = hot.ag + t;
    c.pred[0] = hot.pred[t]; c.pred[1] = hot.pred[K.ldt + t];
    c.base = hot.st_base[t]; c.NS = K.NS; c.numStates = K.numStates;
    c.states = V.states + t;
    c.dOldp = hot.deltaOld + t;
    c.rcp = hot.lte_rcp + t;
    c.lte = hot.lte + t;
}

// MOSdev::load for every (instance, trial): CTA = one instance x MOS_TPB adjacent trials.
// Stage 1: every per-trial input of the thread (4 node voltages, old state-0 values, the state-1
// block, the cached operating point, per-trial parameters) is copied to the thread's column of
// the shared-memory staging area with cp.async -- ~50 independent 8-byte copies in flight per
// thread, nothing parked in registers.  Stage 2: the model is evaluated out of shared memory;
// state-0 results and stamps go straight to HBM, the operating point is updated in place in
// the staging area and written back at the end.
template <int MINB, bool L6>
__global__ void __launch_bounds__(MOS_TPB, MINB)
k_load_mos(Consts K, HotView hot, VecView V, const __grid_constant__ MosView M, const double *rhsOld)
{
    const MosStageMap &map = M.smap;
    extern __shared__ double stage[];
    const int tid = threadIdx.x;
    int t = blockIdx.x*MOS_TPB + tid;
    int i = GRID_ROW;
    if (i >= M.n) return;
    // shared parameter block (values common to all trials)
    constexpr int NSHARED = L6 ? MP_NPARAMS : MP_NBASE;      // a level-1 deck stages the base set only
    for (int k = tid; k < NSHARED; k += MOS_TPB)
        if (!map.ts[k]) stage[map.off[k]] = M.prm[k].base[i];
    // this device's contribution slots (one copy per CTA: every thread stamps the same device)
    int *sslot = (int*)(stage + M.stage_doubles) - MOS_NSLOT;
    for (int k = tid; k < MOS_NSLOT; k += MOS_TPB)
        sslot[k] = (k < MT_NSTAMP) ? M.cslot[(size_t)i*MT_NSTAMP + k] : M.rslot[(size_t)i*MRH_NRHS + (k - MT_NSTAMP)];
    int mode = (t < K.ntrials) ? hot.mode[t] : 0;
    // trials that are finished (or past the end) stage nothing but still meet the barrier
    TrialCtx c;
    MosInst m;
    m.type = M.type[i]; m.off = M.off[i]; m.st = M.state[i];
    m.level = L6 ? M.level[i] : 1;               // decks without level-6 devices get the lean kernel
    m.g = M.g[i]; m.b = M.b[i]; m.dp = M.dp[i]; m.sp = M.sp[i];
    m.col = stage + tid; m.shp = stage; m.map = &map; m.tid8 = 8u*tid;
    const int st = m.st;
    const size_t ldt = K.ldt;
    const size_t cstride = (size_t)M.n*ldt;
    double *cp = M.cache + (size_t)i*ldt + t;
    // this device's rows of state planes 0..3 (row k of a plane at +k*ldt; 17 rows fit int offsets)
    double *sg0 = nullptr, *sg1 = nullptr, *sg2 = nullptr, *sg3 = nullptr;
    const int ldti = K.ldt;
#define SG(p, k) ((p) + (k)*ldti)
    // Stage 1a: what the bypass test needs -- node voltages, old state-0 voltages, operating point.
    if (mode) {
        make_ctx(c, K, hot, V, t, mode);
        sg0 = &ST(c, 0, st); sg1 = &ST(c, 1, st); sg2 = &ST(c, 2, st); sg3 = &ST(c, 3, st);
        m.sg2 = sg2;
        if (!(mode & MODEINITPRED)) {
            const double *x = rhsOld + t;
            __pipeline_memcpy_async(&m.R(SR_V_B), x + (size_t)m.b*ldt, 8);
            __pipeline_memcpy_async(&m.R(SR_V_SP), x + (size_t)m.sp*ldt, 8);
            __pipeline_memcpy_async(&m.R(SR_V_G), x + (size_t)m.g*ldt, 8);
            __pipeline_memcpy_async(&m.R(SR_V_DP), x + (size_t)m.dp*ldt, 8);
        }
        __pipeline_memcpy_async(&m.s0(MS_vbs), SG(sg0, MS_vbs), 8);
        __pipeline_memcpy_async(&m.s0(MS_vgs), SG(sg0, MS_vgs), 8);
        __pipeline_memcpy_async(&m.s0(MS_vds), SG(sg0, MS_vds), 8);
        __pipeline_memcpy_async(&m.s0(MS_vbd), SG(sg0, MS_vbd), 8);
#pragma unroll 1
        for (int k = 0; k < MC_NCACHE; k++)
            __pipeline_memcpy_async(&m.R(SR_CACHE + k), cp + k*cstride, 8);
        __pipeline_commit();
    }
    // MC_mode carries the channel direction in its sign and, as magnitude 2, the note "the last
    // load of this device was a bypass": its stamps, Meyer currents and LTE bounds are then
    // already in HBM, bit for bit what a second bypass load would recompute from the same
    // cached operating point, state and integration coefficients (any non-bypass load, and the
    // predictor load that opens every time step, rewrites the operating point with magnitude 1).
    // Only FLOAT-mode loads can repeat a bypass; the others go straight to stage 1b.
    bool prev_bypass = false, skip = false;
    double lte0 = 0, lte1 = 0;
    const bool maybe = mode && (mode & MODEINITFLOAT) && K.opt.bypass;
    if (__syncthreads_or(maybe)) {
        __pipeline_wait_prior(0);
        if (maybe) {
            double raw = m.R(SR_CACHE + MC_mode);
            prev_bypass = (fabs(raw) == 2.0);
            if (prev_bypass && !(hot.trapcheck[t] && (mode & MODETRAN))) {
                m.mode = (raw > 0) ? 1 : -1;
                if (mode & MODETRAN) {                  // cached bounds, in flight during the test
                    lte0 = M.lte[(size_t)i*ldt + t];
                    lte1 = M.lte[(size_t)i*ldt + t + cstride];
                }
                MosStuff ms;
                dev_mos_get_v(m, ms);
                ms.vbd = ms.vbs - ms.vds;
                ms.vgd = ms.vgs - ms.vds;
                skip = dev_mos_bypass_ok(m, ms, K.opt.reltol, K.opt.abstol, K.opt.vntol);
            }
        }
    }
    // Stage 1b: everything else the evaluation needs, unless the device repeats its bypass.
    if (mode && !skip) {
        if (mode & (MODETRAN | MODEINITPRED)) {
            // old Meyer capacitances and gate charges (bypass re-integrates them)
            __pipeline_memcpy_async(&m.s0(MS_capgs), SG(sg0, MS_capgs), 8);
            __pipeline_memcpy_async(&m.s0(MS_capgd), SG(sg0, MS_capgd), 8);
            __pipeline_memcpy_async(&m.s0(MS_capgb), SG(sg0, MS_capgb), 8);
            __pipeline_memcpy_async(&m.s0(MS_qgs), SG(sg0, MS_qgs), 8);
            __pipeline_memcpy_async(&m.s0(MS_qgd), SG(sg0, MS_qgd), 8);
            __pipeline_memcpy_async(&m.s0(MS_qgb), SG(sg0, MS_qgb), 8);
#pragma unroll 1
            for (int k = 0; k < MS_NSTATES; k++)
                __pipeline_memcpy_async(&m.s1(k), SG(sg1, k), 8);
        }
        if ((mode & MODETRAN) && (mode & MODEINITFLOAT)) {
            const int qs[3] = { MS_qgs, MS_qgd, MS_qgb };
#pragma unroll
            for (int k = 0; k < 3; k++) {
                __pipeline_memcpy_async(&m.R(SR_Q2 + k), SG(sg2, qs[k]), 8);
                __pipeline_memcpy_async(&m.R(SR_Q3 + k), SG(sg3, qs[k]), 8);
            }
        }
        if (mode & MODEINITPRED) {
            __pipeline_memcpy_async(&m.R(SR_S2_vbs), SG(sg2, MS_vbs), 8);
            __pipeline_memcpy_async(&m.R(SR_S2_vgs), SG(sg2, MS_vgs), 8);
            __pipeline_memcpy_async(&m.R(SR_S2_vds), SG(sg2, MS_vds), 8);
        }
#pragma unroll 1
        for (int q = 0; q < map.nvar; q++) {
            int k = map.var[q];
            __pipeline_memcpy_async(m.col + map.off[k], M.prm[k].base + (size_t)i*ldt + t, 8);
        }
        __pipeline_commit();
    }
    MosOut o;
    o.cbase = V.contrib + t;
    o.cs = sslot;
    o.rs = sslot + MT_NSTAMP;
    o.ldt = (unsigned)K.ldt;
    double *plte = M.lte + (size_t)i*ldt + t;              // this device's LTE bounds: [2][nmos][ldt]
    __pipeline_wait_prior(0);
    __syncthreads();                                     // shared parameter block visible
    if (!mode) return;
    if (skip) {
        if (mode & MODETRAN) {
            atomic_min_pos(c.lte, lte0);
            if (c.order + 1 <= K.opt.maxord) atomic_min_pos(c.lte + K.ldt, lte1);
        }
        return;
    }
    {
        double raw = m.R(SR_CACHE + MC_mode);
        prev_bypass = (fabs(raw) == 2.0);
        m.mode = (raw > 0) ? 1 : -1;
    }

    int r = dev_mos_load(c, m, o, K.opt);
    // ---- results: state 0, operating point, flags ----
    const bool charges = (mode & MODEINITTRAN) || ((mode & (MODEINITFLOAT | MODEINITPRED)) && (mode & MODETRAN));
    if (r & 8) {
        // bypassed: voltages and charges are unchanged, only the Meyer currents were re-integrated
        if (charges) {
            *SG(sg0, MS_cqgs) = m.s0(MS_cqgs);
            *SG(sg0, MS_cqgd) = m.s0(MS_cqgd);
            *SG(sg0, MS_cqgb) = m.s0(MS_cqgb);
        }
    }
    else if (r & 1) {
        if (charges) {
#pragma unroll 1
            for (int k = 0; k < MS_NSTATES; k++) *SG(sg0, k) = m.s0(k);
        }
        else {
            *SG(sg0, MS_vbs) = m.s0(MS_vbs); *SG(sg0, MS_vbd) = m.s0(MS_vbd);
            *SG(sg0, MS_vgs) = m.s0(MS_vgs); *SG(sg0, MS_vds) = m.s0(MS_vds);
        }
    }
    else if (r & 16) {
        *SG(sg0, MS_vds) = m.s0(MS_vds);
        *SG(sg0, MS_vgs) = m.s0(MS_vgs);
        *SG(sg0, MS_vbs) = m.s0(MS_vbs);
    }
    if ((mode & MODETRAN) && (mode & MODEINITFLOAT) && (r & 1)) {
        // MOSdev::trunc (mostrun.cc:52-65): the three gate charges
        // qgs, qgd, qgb are MS_qgs + 3k (state layout: cap, q, cq per gate charge)
        double m0 = DBL_MAX, m1 = DBL_MAX;
        double lr[LR_NROWS];                             // the trial's divisor rows, read once
#pragma unroll
        for (int k = 0; k < LR_NROWS; k++) lr[k] = c.rcp[(size_t)k*ldt];
#pragma unroll 1
        for (int k = 0; k < 3; k++) {
            const int qk = MS_qgs + 3*k;
            double q[4] = { m.s0(qk), m.s1(qk), m.R(SR_Q2 + k), m.R(SR_Q3 + k) };
            dev_lte_charge<false>(c, K.opt, q, m.s0(qk + 1), m.s1(qk + 1), st + qk, m0, m1, lr, 1);
        }
        atomic_min_pos(c.lte, m0);
        if (c.order + 1 <= K.opt.maxord) atomic_min_pos(c.lte + K.ldt, m1);
        plte[0] = m0; plte[cstride] = m1;
    }
    if ((r & 8) && !prev_bypass)
        cp[MC_mode*cstride] = 2.0*(double)m.mode;        // bypassed: see the note above
    else if (!(r & 2) && prev_bypass && !(r & 8))
        cp[MC_mode*cstride] = (double)m.mode;
    if (r & 2) {
        m.R(SR_CACHE + MC_mode) = (double)m.mode;
#pragma unroll 1
        for (int k = 0; k < MC_NCACHE; k++)
            cp[k*cstride] = m.R(SR_CACHE + k);
    }
    if (r & 4)
        hot.noncon[t] = 1;
    if (!(r & 1)) {
#pragma unroll 1
        for (int k = 0; k < MT_NSTAMP; k++) o.M(k, 0.0);
#pragma unroll 1
        for (int k = 0; k < MRH_NRHS; k++) o.Rw(k, 0.0);
    }
#undef SG
}

__global__ void __launch_bounds__(128, 6)
k_load_cap(Consts K, HotView hot, VecView V, CapView Cp, const double *rhsOld)
{
    // capload.cc:75-205 (linear capacitor)
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;
    if (i >= Cp.n) return;
    if (t >= K.ntrials) return;
    int mode = hot.mode[t];
    if (!mode) return;
    const int *cs = Cp.cslot + (size_t)i*4;
    const int *rs = Cp.rslot + (size_t)i*2;
    double geq = 0.0, ceq = 0.0;
    bool stamp = false;
    if (!(mode & MODEDC)) {
        TrialCtx c;
        make_ctx(c, K, hot, V, t, mode);
        int st = Cp.state[i];
        int pos = Cp.pos[i], neg = Cp.neg[i];
        double capac = PR(Cp.capac, i, t);
        double *pg = Cp.geq + (size_t)i*K.ldt + t, *pc = Cp.ceq + (size_t)i*K.ldt + t;
        if (mode & MODEINITFLOAT) {
            // every load of the iteration is issued before the first store, so they overlap
            // (the kernel is latency bound: a dozen values per thread, all in different rows)
            const double vp = rhsOld[(size_t)pos*K.ldt + t], vn = rhsOld[(size_t)neg*K.ldt + t];
            const double q1 = ST(c, 1, st), c1 = ST(c, 1, st + 1);
            const double cq_old = *pc;
            geq = *pg;
            double q2 = 0, q3 = 0;
            if (mode & MODETRAN) { q2 = ST(c, 2, st); q3 = ST(c, 3, st); }
            double qcap = capac*(vp - vn);
            ST(c, 0, st) = qcap;
            // sCKT::integrate's current update + trapcheck (dev_integrate2) on the values at hand
            const double c0 = c.ag[0]*qcap + cq_old;
            ST(c, 0, st + 1) = c0;
            if (c.trapCheck && !*c.trapbad && !(fabs(c0) < 1e-6) && !(fabs(qcap) < 1e-16)) {
                double c2 = ST(c, 2, st + 1);
                double dqdt = fabs(qcap - q1)/c.delta;
                if (fabs(c1) > c.trapratio*dqdt && fabs(c0 - c1) > c.trapratio*fabs(c0 - c2))
                    *c.trapbad = 1;
            }
            if (c.method == 0)
                ceq = (c.order <= 1 ? q1*c.ag[1] : - c1*c.ag[1] - q1*c.ag[0]);
            else
                ceq = dev_NIceq(c, st);
            *pc = ceq;
            stamp = true;
            if (mode & MODETRAN) {
                // CAPdev::trunc (captrun.cc:58)
                double q[4] = { qcap, q1, q2, q3 };
                double m0 = DBL_MAX, m1 = DBL_MAX;
                dev_lte_charge<true>(c, K.opt, q, c0, c1, st, m0, m1, c.rcp, c.ldt);
                atomic_min_pos(c.lte, m0);
                if (c.order + 1 <= K.opt.maxord) atomic_min_pos(c.lte + K.ldt, m1);
            }
        }
        else if (mode & MODEINITPRED) {
            geq = c.ag[0]*capac;
            ceq = dev_find_ceq(c, st);
            *pg = geq; *pc = ceq;
            stamp = true;
        }
        else if (mode & MODEINITTRAN) {
            double vcap;
            if (mode & MODEUIC)
                vcap = PR(Cp.ic, i, t);
            else
                vcap = rhsOld[(size_t)pos*K.ldt + t] - rhsOld[(size_t)neg*K.ldt + t];
            ST(c, 1, st) = capac*vcap;
            geq = c.ag[0]*capac;
            ceq = dev_find_ceq(c, st);
            *pg = geq; *pc = ceq;
            stamp = true;
        }
    }
    if (!stamp) { geq = 0.0; ceq = 0.0; }
    // load_cap(): pos-only / neg-only / both (slots exist only for the live stamps)
    if (cs[0] >= 0) V.contrib[CADDR(K, cs[0], t)] = geq;
    if (cs[1] >= 0) V.contrib[CADDR(K, cs[1], t)] = geq;
    if (cs[2] >= 0) V.contrib[CADDR(K, cs[2], t)] = -geq;
    if (cs[3] >= 0) V.contrib[CADDR(K, cs[3], t)] = -geq;
    if (rs[0] >= 0) V.contrib[CADDR(K, rs[0], t)] = stamp ? -ceq : 0.0;
    if (rs[1] >= 0) V.contrib[CADDR(K, rs[1], t)] = stamp ? ceq : 0.0;
}

__global__ void __launch_bounds__(128)
k_load_ind(Consts K, HotView hot, VecView V, IndView L, const double *rhsOld)
{
    // indload.cc:84-263 (linear inductor; the +-1 couplings are preloaded constants)
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;
    if (i >= L.n) return;
    if (t >= K.ntrials) return;
    int mode = hot.mode[t];
    if (!mode) return;
    int cs = L.cslot[i], rs = L.rslot[i];
    double req = 0.0, veq = 0.0;
    if (!(mode & MODEDC)) {
        TrialCtx c;
        make_ctx(c, K, hot, V, t, mode);
        int st = L.state[i], br = L.branch[i];
        double induct = PR(L.induct, i, t);
        size_t at = (size_t)i*K.ldt + t;
        if (mode & MODEINITFLOAT) {
            double prev = L.prevFlux[at];
            double f0 = ST(c, 0, st);
            // MUTdev::load runs first (mutload.cc:96-104): cross fluxes, in MUT load order
            for (int q = L.mptr[i]; q < L.mptr[i+1]; q++) {
                int o = L.mother[q];
                double factor = L.mk[q]*sqrt(induct*PR(L.induct, o, t));
                f0 += factor*rhsOld[(size_t)L.branch[o]*K.ldt + t];
            }
            double flux = induct*rhsOld[(size_t)br*K.ldt + t] - prev;
            f0 = f0 + flux;
            ST(c, 0, st) = f0;
            L.prevFlux[at] = f0;
            veq = L.veq[at];
            dev_integrate2(c, st, veq);
            req = L.req[at];
            if (mode & MODETRAN) {
                // INDdev::trunc (indtrun.cc:58)
                double q[4] = { f0, ST(c, 1, st), ST(c, 2, st), ST(c, 3, st) };
                double m0 = DBL_MAX, m1 = DBL_MAX;
                dev_lte_charge<true>(c, K.opt, q, ST(c, 0, st + 1), ST(c, 1, st + 1), st, m0, m1, c.rcp, c.ldt);
                atomic_min_pos(c.lte, m0);
                if (c.order + 1 <= K.opt.maxord) atomic_min_pos(c.lte + K.ldt, m1);
            }
        }
        else if (mode & MODEINITPRED) {
            req = c.ag[0]*induct;
            veq = dev_find_ceq(c, st);
            L.req[at] = req; L.veq[at] = veq;
            L.prevFlux[at] = 0.0;
            ST(c, 0, st) = 0.0;
        }
        else if (mode & MODEINITTRAN) {
            double ival;
            if (mode & MODEUIC)
                ival = PR(L.ic, i, t);
            else
                ival = rhsOld[(size_t)br*K.ldt + t];
            double f1 = ST(c, 1, st);
            for (int q = L.mptr[i]; q < L.mptr[i+1]; q++) {
                // mutload.cc:118-133
                int o = L.mother[q];
                double factor = L.mk[q]*sqrt(induct*PR(L.induct, o, t));
                double io = (mode & MODEUIC) ? PR(L.ic, o, t) : rhsOld[(size_t)L.branch[o]*K.ldt + t];
                f1 += factor*io;
            }
            ST(c, 1, st) = f1 + induct*ival;
            L.prevFlux[at] = 0.0;
            ST(c, 0, st) = 0.0;
            req = c.ag[0]*induct;
            veq = dev_find_ceq(c, st);
            L.req[at] = req; L.veq[at] = veq;
        }
    }
    else if (K.jjDCphase) {
        // phase-mode DC (indload.cc:88-96): the inductor's "voltage" is its phase
        req = 2*M_PI*PR(L.induct, i, t)/JJ_PHI0;
    }
    // other DC modes: nothing but the preloaded +-1 (the branch equation shorts the inductor)
    if (rs >= 0) V.contrib[CADDR(K, rs, t)] = veq;
    if (cs >= 0) V.contrib[CADDR(K, cs, t)] = -req;
}

__global__ void __launch_bounds__(128)
k_load_mut(Consts K, HotView hot, VecView V, IndView L)
{
    // mutload.cc:55-140, matrix part (the flux part runs inside k_load_ind, in load order)
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;
    if (i >= L.nmut) return;
    if (t >= K.ntrials) return;
    int mode = hot.mode[t];
    if (!mode) return;
    double v = 0.0;
    if (!(mode & MODEDC) && (mode & (MODEINITFLOAT | MODEINITPRED | MODEINITTRAN))) {
        double factor = L.mutk[i]*sqrt(PR(L.induct, L.mut1[i], t)*PR(L.induct, L.mut2[i], t));
        factor *= hot.ag[t];
        v = -factor;
    }
    else if ((mode & MODEDC) && K.jjDCphase) {
        // mutload.cc:75-87
        double factor = L.mutk[i]*sqrt(PR(L.induct, L.mut1[i], t)*PR(L.induct, L.mut2[i], t));
        v = -(2*M_PI*factor/JJ_PHI0);
    }
    int c0 = L.mut_cslot[2*i], c1 = L.mut_cslot[2*i + 1];
    if (c0 >= 0) V.contrib[CADDR(K, c0, t)] = v;
    if (c1 >= 0) V.contrib[CADDR(K, c1, t)] = v;
}

__global__ void __launch_bounds__(128)
k_load_src(Consts K, HotView hot, VecView V, SrcView S)
{
    // srcload.cc:67-105, 215-262
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;
    if (i >= S.n) return;
    if (t >= K.ntrials) return;
    int mode = hot.mode[t];
    if (!mode) return;
    bool is_dc = mode & (MODEDCOP | MODEDCTRANCURVE);
    int fn = S.fn[i];
    int ft = (fn == B200_FN_NONE) ? 1 : (is_dc ? (S.dcgiven[i] ? 1 : 2) : 2);
    double srcFact = hot.srcfact[t];
    double rhs;
    if (ft == 1)
        rhs = srcFact*PR(S.dc, i, t);
    else {
        rhs = dev_src_eval(fn, S.pool + S.coef_off[i], S.ncoef[i], hot.time[t]);
        if (K.analysis == 1 && fn == B200_FN_GPULSE)
            rhs = S.pool[S.coef_off[i]];           // `.op`: td_enable_tran is false, V1 (inptran.cc:1736-1737)
        if (mode & (MODEDCOP | MODETRANOP))
            rhs *= srcFact;
    }
    const int *rs = S.rslot + (size_t)i*2;
    if (S.kind[i] == B200_SRC_V) {
        if (rs[0] >= 0) V.contrib[CADDR(K, rs[0], t)] = rhs;
    }
    else {
        if (rs[0] >= 0) V.contrib[CADDR(K, rs[0], t)] = -rhs;
        if (rs[1] >= 0) V.contrib[CADDR(K, rs[1], t)] = rhs;
    }
}

// ---------------------------------------------------------------------------------------
// MOSdev::convTest, gated per trial; each thread walks a chunk of instances and stops at the
// first failure ("no reason to continue", mosconv.cc:112)
#define DEVCONV_CHUNK 8
__global__ void __launch_bounds__(128)
k_devconv(Consts K, HotView hot, VecView V, MosView M, const double *rhs)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i0 = GRID_ROW*DEVCONV_CHUNK;
    if (t >= K.ntrials) return;
    if (!hot.need_devconv[t]) return;
    TrialCtx c;
    make_ctx(c, K, hot, V, t, hot.mode[t]);
    size_t cstride = (size_t)M.n*K.ldt;
    int i1 = min(M.n, i0 + DEVCONV_CHUNK);
    for (int i = i0; i < i1; i++) {
        MosConvIn m;
        m.type = M.type[i]; m.off = M.off[i]; m.st = M.state[i];
        m.g = M.g[i]; m.b = M.b[i]; m.dp = M.dp[i]; m.sp = M.sp[i];
        const double *cp = M.cache + (size_t)i*K.ldt + t;
        m.mode = (int)cp[MC_mode*cstride];
        m.cd = cp[MC_cd*cstride]; m.cbs = cp[MC_cbs*cstride]; m.cbd = cp[MC_cbd*cstride];
        m.gm = cp[MC_gm*cstride]; m.gds = cp[MC_gds*cstride]; m.gmbs = cp[MC_gmbs*cstride];
        m.gbd = cp[MC_gbd*cstride]; m.gbs = cp[MC_gbs*cstride];
        if (dev_mos_conv(c, m, rhs + t, K.opt.reltol, K.opt.abstol)) {
            hot.devconv[t] = 1;
            return;
        }
    }
}

// ---------------------------------------------------------------------------------------
// control
__global__ void k_ctl(Consts K, HotView hot, TrialCtl *ctl, CtlShared sh, CtlAux aux, int phase)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= K.ntrials) return;
    sh.K = &K;
    TrialCtl s = ctl[t];
    if (phase == 0) {
        // begin
        aux.vo_stbase[t] = s.st_base;
        s.vecops = 0;
        s.want_load = 0;
        ctl_begin(s, K, sh);
    }
    else if (phase == 1) {
        // st_base in force when state-vector ops are requested = value at the start of the
        // control pass (a rotation, if any, comes after the request inside the same tick)
        aux.vo_stbase[t] = s.st_base;
        if (s.status != ST_RUNNING) {
            // finished trials: no load, no solve, and no stale vector-op requests
            hot.mode[t] = 0; hot.do_solve[t] = 0; aux.vecops[t] = 0; hot.sync[t] = 1;
            return;
        }
        s.vecops = (s.pc == PC_SYNC) ? VO_KEEP_X : 0;      // parked since an earlier tick
        s.want_load = 0;
        CtlIn in;
        in.solve_err = hot.do_solve[t] ? hot.solve_err[t] : 0;
        in.accel_exit = 0;
        if (K.jjaccel) { in.accel_exit = hot.accel_exit[t]; hot.accel_exit[t] = 0; }
        in.nodeconv = hot.nodeconv[t];
        in.devconv = hot.devconv[t];
        in.noncon_load = hot.noncon[t];
        in.trapbad = hot.trapbad[t];
        in.lte0 = hot.lte[t];
        in.lte1 = hot.lte[K.ldt + t];
        ctl_after_solve(s, K, sh, in, hot, t);
        hot.sync[t] = (s.status != ST_RUNNING || s.pc == PC_SYNC) ? 1 : 0;
    }
    else {
        // phase 2: release the groups whose running trials have all reached the time-point
        // boundary (ctl_sync_point); everybody else was published by phase 1
        if (s.status != ST_RUNNING || s.pc != PC_SYNC) return;
        int g0 = t & ~31;
        for (int k = 0; k < 32; k++) {
            int u = g0 + k;
            if (u < K.ntrials && !hot.sync[u]) return;
        }
        // JJdev::accept's bound on the next step (jjacct.cc:52-141): per model, the largest |vj|
        // the trial's last load left in state 0, floored at vdpbak/tsaccl
        s.jj_dmax = 0.0;
        for (int m = 0; m < K.njjmod; m++) {
            if (!K.jj_limits[m]) continue;
            double vmax = hot.jj_vmax[(size_t)m*K.ldt + t];
            if (vmax < K.jj_vth[m]) vmax = K.jj_vth[m];
            if (vmax > 0.0) {
                double delmax = K.jj_tsfact[m]*JJ_PHI0/vmax;
                if (s.jj_dmax == 0.0 || delmax < s.jj_dmax) s.jj_dmax = delmax;
            }
        }
        ctl_tran_loop_top(s, K, sh);
    }
    // publish for the vector / load kernels of this tick
    {
        bool load = s.want_load && s.status == ST_RUNNING;
        hot.mode[t] = load ? s.mode : 0;
        hot.do_solve[t] = (load && s.pc == PC_NIITER) ? 1 : 0;
        hot.active[t] = (s.status == ST_RUNNING);
        if (load) {
            hot.order[t] = s.order;
            hot.st_base[t] = s.st_base;
            hot.noncon[t] = 0;
            for (int m = 0; m < K.njjmod; m++) hot.jj_vmax[(size_t)m*K.ldt + t] = 0.0;
            hot.accepted[t] = s.did_accept;
            s.did_accept = 0;
            hot.trapbad[t] = 0;
            hot.trapcheck[t] = s.trapCheck;
            hot.iterno[t] = s.iterno;
            hot.time[t] = s.time;
            hot.delta[t] = s.delta;
            hot.srcfact[t] = s.srcFact;
            hot.diaggmin[t] = s.diagGmin;
            int nag = (K.method == 0) ? 2 : 7;
            for (int k = 0; k < nag; k++) hot.ag[(size_t)k*K.ldt + t] = s.ag[k];
            hot.pred[t] = s.pred[0]; hot.pred[K.ldt + t] = s.pred[1];
            if (s.mode & MODETRAN) {
                // LTE bounds are accumulated by this tick's loads (min over all charges)
                hot.lte[t] = DBL_MAX;
                hot.lte[K.ldt + t] = DBL_MAX;
                for (int k = 0; k <= s.order + 1 && k < 7; k++)
                    hot.deltaOld[(size_t)k*K.ldt + t] = s.deltaOld[k];
                // divisors of terr's divided-difference table, in its own association order
                double dv[LR_NDIV];
                dv[LR_d0] = s.deltaOld[0]; dv[LR_d1] = s.deltaOld[1]; dv[LR_d2] = s.deltaOld[2];
                dv[LR_e0] = dv[LR_d1] + dv[LR_d0];
                dv[LR_e1] = dv[LR_d2] + dv[LR_d1];
                dv[LR_f0] = dv[LR_e1] + dv[LR_d0];
                for (int k = 0; k < LR_NDIV; k++) {
                    hot.lte_rcp[(size_t)k*K.ldt + t] = dv[k];
                    hot.lte_rcp[(size_t)(LR_NDIV + k)*K.ldt + t] = 1.0/dv[k];
                }
                hot.lte_rcp[(size_t)LR_CT*K.ldt + t] = K.opt.reltol/s.delta;
            }
        }
    }
    // vector-op requests
    {
        aux.vecops[t] = s.vecops;
        aux.sols_base[t] = s.sols_base;
        if (s.vecops & (VO_INTERP | VO_DUMP_DIRECT)) {
            aux.ip_n[t] = (s.vecops & VO_INTERP) ? s.ip_n : 0;
            aux.ip_out0[t] = s.ip_out0;
            aux.ip_degree[t] = s.ip_degree;
            aux.ip_d[t] = s.ip_time;
            aux.ip_d[K.ldt + t] = s.ip_c2;
            aux.ip_d[2*(size_t)K.ldt + t] = s.ip_c3;
            aux.ip_d[3*(size_t)K.ldt + t] = s.ip_c4;
            aux.ip_d[4*(size_t)K.ldt + t] = s.ip_tcheck0;
        }
    }
    ctl[t] = s;
}

__global__ void __launch_bounds__(128)
k_vec(Consts K, VecView V, CtlAux aux, double *rhsOld)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;                          // node 0..n
    if (i >= K.n + 1) return;
    if (t >= K.ntrials) return;
    int ops = aux.vecops[t];
    if (!ops) return;
    size_t ldt = K.ldt;
    size_t at = (size_t)i*ldt + t;
    size_t vplane = (size_t)(K.n + 1)*ldt;
    if (ops & VO_KEEP_X) {
        const double *other = (rhsOld == V.X[0]) ? V.X[1] : V.X[0];
        rhsOld[at] = other[at];
    }
    if ((ops & VO_DCMU) && i) {
        // option dcmu: mix in a little of the last solution (DC operating point damping)
        const double *other = (rhsOld == V.X[0]) ? V.X[1] : V.X[0];
        const double A = K.opt.dcmu + 0.5, B = 1.0 - A;
        rhsOld[at] = A*rhsOld[at] + B*other[at];
    }
    if ((ops & VO_ZERO_PHASE) && V.phase && V.phase[i]) rhsOld[at] = 0.0;
    if (ops & VO_BK_SAVE) V.oldSol[at] = rhsOld[at];
    if (ops & VO_BK_RESTORE) rhsOld[at] = V.oldSol[at];
    if (ops & VO_RESTORE_OLD) { if (i) rhsOld[at] = V.spare[at]; }
    if (ops & VO_SAVE_SPARE) { if (i) V.spare[at] = rhsOld[at]; }
    int sb = aux.sols_base[t];
    if (ops & VO_ACCEPT_SOLS) V.sols[(size_t)sb*vplane + at] = rhsOld[at];
    if (!i) return;
    const int os = V.outslot ? V.outslot[i] : i - 1;      // the `.save` list (b200_batch_set_saves)
    if (os < 0) return;
    size_t oplane = (size_t)K.nout*ldt;
    if (ops & VO_DUMP_DIRECT)
        V.out[(size_t)aux.ip_out0[t]*oplane + (size_t)os*ldt + t] = rhsOld[at];
    if (ops & VO_INTERP) {
        int cnt = aux.ip_n[t], out0 = aux.ip_out0[t], degree = aux.ip_degree[t];
        if (ops & VO_DUMP_DIRECT) out0 += 1;
        double time = aux.ip_d[t], c2 = aux.ip_d[ldt + t], c3 = aux.ip_d[2*ldt + t];
        double c4 = aux.ip_d[3*ldt + t], tc = aux.ip_d[4*ldt + t];
        double s0 = V.sols[(size_t)sb*vplane + at];
        double s1 = V.sols[(size_t)((sb + 1) & 7)*vplane + at];
        double s2 = V.sols[(size_t)((sb + 2) & 7)*vplane + at];
        double s3 = V.sols[(size_t)((sb + 3) & 7)*vplane + at];
        for (int p = 0; p < cnt; p++, tc += K.t_step) {
            // tranan.cc:999-1101
            double d1 = time - tc, v;
            if (degree <= 1) {
                double d2 = d1 - c2;
                double f0 = d2/(-c2), f1 = d1/(c2);
                v = f0*s0 + f1*s1;
            }
            else if (degree == 2) {
                double d2 = d1 - c2, d3 = d1 - c3;
                double f0 = d2*d3/((c2)*(c3));
                double f1 = d1*d3/((c2)*(c2-c3));
                double f2 = d1*d2/((c3)*(c3-c2));
                v = f0*s0 + f1*s1 + f2*s2;
            }
            else {
                double d2 = d1 - c2, d3 = d1 - c3, d4 = d1 - c4;
                double f0 = d2*d3*d4/((-c2)*(c3)*(c4));
                double f1 = d1*d3*d4/((c2)*(c2-c3)*(c2-c4));
                double f2 = d1*d2*d4/((c3)*(c3-c2)*(c3-c4));
                double f3 = d1*d2*d3/((c4)*(c4-c2)*(c4-c3));
                v = f0*s0 + f1*s1 + f2*s2 + f3*s3;
            }
            V.out[(size_t)(out0 + p)*oplane + (size_t)os*ldt + t] = v;
        }
    }
}

#define SV_CHUNKS 64
__global__ void __launch_bounds__(128)
k_statevec(Consts K, VecView V, CtlAux aux, MosView M)
{
    // grid = (trial blocks, chunks): a block with no request among its 128 trials exits at
    // once (the common case in the transient); otherwise it walks its share of the slots.
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    const int mask = VO_ZERO_STATE0 | VO_BK_SAVE | VO_BK_RESTORE | VO_HIST_INIT | VO_FIRSTTIME;
    int ops = (t < K.ntrials) ? (aux.vecops[t] & mask) : 0;
    if (!__syncthreads_or(ops)) return;
    if (!ops) return;
    size_t ldt = K.ldt;
    int total = K.numStates + MC_NCACHE*M.n;
    const int nchunks = gridDim.y*gridDim.z;
    int per = (total + nchunks - 1)/nchunks;
    int e0 = GRID_ROW*per, e1 = min(total, e0 + per);
    int base = aux.vo_stbase[t];
    size_t plane = (size_t)K.numStates*ldt;
    double *sp[8];
    for (int k = 0; k < K.NS; k++) {
        int phys = base + k; if (phys >= K.NS) phys -= K.NS;
        sp[k] = V.states + phys*plane;
    }
    for (int e = e0; e < e1; e++) {
        if (e >= K.numStates) {
            size_t at = (size_t)(e - K.numStates)*ldt + t;
            if (ops & VO_BK_SAVE) M.save[at] = M.cache[at];
            if (ops & VO_BK_RESTORE) {
                double v = M.save[at];
                // MC_mode rows: drop the "last load was a bypass" note (see k_load_mos) -- the
                // stamps in HBM belong to the abandoned attempt, not to the restored state
                if (e - K.numStates < M.n) v = (v > 0) ? 1.0 : -1.0;
                M.cache[at] = v;
            }
            continue;
        }
        size_t at = (size_t)e*ldt + t;
        if (ops & VO_ZERO_STATE0) sp[0][at] = 0.0;
        if (ops & VO_BK_SAVE) V.oldState0[at] = sp[0][at];
        if (ops & VO_BK_RESTORE) sp[0][at] = V.oldState0[at];
        if (ops & VO_HIST_INIT) {
            double v = sp[0][at];
            for (int k = 1; k <= K.opt.maxord + 1 && k < K.NS; k++) sp[k][at] = v;
        }
        if (ops & VO_FIRSTTIME) {
            double v = sp[1][at];
            sp[2][at] = v;
            if (K.opt.maxord > 1) sp[3][at] = v;
        }
    }
}

// jjaccel (niiter.cc:535-539): a transient FLOAT load in which no device flagged non-convergence
// ends NIiter there and then -- its matrix is never solved.  Runs between the load and the solve.
__global__ void k_jjaccel_gate(Consts K, HotView hot)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= K.ntrials || !hot.do_solve[t]) return;
    const int mode = hot.mode[t];
    if (hot.noncon[t] == 0 && (mode & MODEINITFLOAT) && !(mode & (MODEDCOP | MODETRANOP | MODEINITTRAN))) {
        hot.do_solve[t] = 0;
        hot.accel_exit[t] = 1;
    }
}

// Per-trial resistor values: the preloaded constant of every entry that holds resistor
// conductances, re-summed per trial in the order the nominal one was (restemp.cc:209-229)
__global__ void k_res_ainit(int ne, int ldt, const int *cslot, const double *base, const int *ptr,
    const int *res, const signed char *sign, ParamRef r, const double *m, double resmin, double *contrib)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int e = GRID_ROW;
    if (t >= ldt || e >= ne || cslot[e] < 0) return;
    double v = base[e];
    for (int k = ptr[e]; k < ptr[e+1]; k++) {
        double rv = PR(r, res[k], t);
        if (fabs(rv) < resmin) rv = (rv < 0.0) ? -resmin : resmin;
        const double g = m[res[k]]/(rv*1.0);
        v += (sign[k] > 0) ? g : -g;
    }
    contrib[(size_t)cslot[e]*ldt + t] = v;
}

void launch_res_ainit(cudaStream_t st, int ne, int ldt, const int *cslot, const double *base, const int *ptr,
    const int *res, const signed char *sign, ParamRef r, const double *m, double resmin, double *contrib)
{
    if (ne) k_res_ainit<<<grid_rows((ldt + 127)/128, ne), 128, 0, st>>>(ne, ldt, cslot, base, ptr, res, sign, r, m, resmin, contrib);
}

__global__ void k_count_active(int ntrials, const int *active, int *count)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int a = (t < ntrials) ? active[t] : 0;
    unsigned m = __ballot_sync(0xffffffffu, a);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}

__global__ void k_fill(double *p, size_t n, double v)
{
    size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// out[npts][n][ldt] -> waves[ntrials][npts][n] (host layout), tiled transpose over (unknown, trial)
__global__ void k_gather_waves(int n, int ldt, int ntrials, int npts_max, const double *out,
    const int *npts, int nsel, const int *sel, double *waves)
{
    __shared__ double tile[32][33];
    int p = blockIdx.z;
    int t0 = blockIdx.x*32, i0 = blockIdx.y*32;
    int nn = sel ? nsel : n;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int i = i0 + r, t = t0 + threadIdx.x;
        double v = 0.0;
        if (i < nn && t < ntrials) {
            int src = sel ? sel[i] : i;
            v = out[((size_t)p*n + src)*ldt + t];
        }
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int t = t0 + r, i = i0 + threadIdx.x;
        if (i < nn && t < ntrials)
            waves[((size_t)t*npts_max + p)*nn + i] = (p < npts[t]) ? tile[threadIdx.x][r] : 0.0;
    }
}

__global__ void k_export_counters(int ntrials, const TrialCtl *ctl, int *counters, int *npts)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= ntrials) return;
    const TrialCtl &s = ctl[t];
    int *c = counters + (size_t)t*B200_NCOUNTERS;
    c[B200_C_STATUS] = (s.status == ST_DONE) ? 0 : (s.status == ST_RUNNING ? -1 : s.status);
    c[B200_C_ACCEPTED] = s.accepted; c[B200_C_REJECTED] = s.rejected;
    c[B200_C_NUMITER] = s.numIter; c[B200_C_TIMEPTS] = s.timePts; c[B200_C_TRANITER] = s.tranIter;
    c[B200_C_ITERCUT] = s.iterCut; c[B200_C_TRAPCUT] = s.trapCut; c[B200_C_NPOINTS] = s.npts;
    npts[t] = s.npts;
}

// ---------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------
static inline dim3 grid2(int ntrials, int ny, int bs = 128)
{
    return grid_rows((ntrials + bs - 1)/bs, ny);
}


// ---------------------------------------------------------------------------------------
// RESdev::load for the resistors a JJ deck does not preload (a terminal on a phase node,
// restemp.cc:213-223): resload.cc:250-292.  In the phase-mode DC modes a resistor between two
// phase / ground nodes leaves only gmin, one between a voltage node and a phase node loads the
// voltage node's side (the phase node looks grounded from there).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_load_rph(Consts K, HotView hot, VecView V, ResPhView R)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;
    if (i >= R.n) return;
    if (t >= K.ntrials) return;
    int mode = hot.mode[t];
    if (!mode) return;
    const int *cs = R.cslot + (size_t)i*4;          // PP NN PN NP
    const int ri = R.res[i];
    double rv = PR(R.r, ri, t);
    if (fabs(rv) < R.resmin) rv = (rv < 0.0) ? -R.resmin : R.resmin;
    const double g = R.m[ri]/(rv*1.0);              // restemp.cc:200-209
    double v[4] = { g, g, -g, -g };
    if ((mode & MODEDC) && K.jjDCphase) {
        const int cls = R.cls[i];
        if (cls == 0) { v[0] = v[1] = v[2] = v[3] = K.opt.gmin; }
        else if (cls == 1) { v[0] = g; v[3] = -g; v[1] = 0.0; v[2] = 0.0; }       // volt - phase: PP, NP
        else if (cls == 2) { v[1] = g; v[2] = -g; v[0] = 0.0; v[3] = 0.0; }       // phase - volt: NN, PN
    }
    for (int k = 0; k < 4; k++) if (cs[k] >= 0) V.contrib[CADDR(K, cs[k], t)] = v[k];
}

// ---------------------------------------------------------------------------------------
// Josephson junction, RSJ level 1.  JJdev::load + jjstuff::{jj_limiting, jj_iv, jj_load}
// (devlib/jj/jjload.cc:77-871), the step-bound part of JJdev::accept (jjacct.cc:120-141) and
// JJdev::trunc (jjtrun.cc:52-64).  One thread per (junction, trial).
//
// JJdev::accept also folds the phase into (-2pi, 2pi] with an integer count of 4pi
// (jjacct.cc:63-80) before the state rotation.  Here the fold is applied where the folded
// values are first read -- the predictor / first load of the next step reads state 1 -- and
// written back to state 1; it is idempotent, so a rejected step that re-runs that load changes
// nothing.  The phase node's value phi + 4pi*n is the same number before and after the fold
// (up to one rounding of the sum).
// ---------------------------------------------------------------------------------------

struct JJOp { double vj, phi, gqt, crhs, crt, pfac; };

__device__ __forceinline__ void jj_iv(const JJView &J, int i, int m, int t, JJOp &js)
{
    // jjload.cc:620-739, rtype 1 / 2 / 3 (0: no quasiparticle branch)
    const double vj = js.vj;
    const double absvj = vj < 0 ? -vj : vj;
    const int rtype = J.rtype[m];
    const double g0 = PR(J.prm[JP_g0], i, t), gn = PR(J.prm[JP_gn], i, t);
    if (rtype == 1) {
        if (absvj < J.vless[i])
            js.gqt = g0;
        else if (absvj < J.vmore[i]) {
            js.gqt = PR(J.prm[JP_gs], i, t);
            const double cr1 = PR(J.prm[JP_cr1], i, t);
            js.crhs = (vj >= 0 ? cr1 : -cr1);
        }
        else {
            js.gqt = gn;
            const double cr2 = PR(J.prm[JP_cr2], i, t);
            js.crhs = (vj >= 0 ? cr2 : -cr2);
        }
    }
    else if (rtype == 2) {
        const double dv = 0.5*J.delv[m];
        const double avj = absvj/dv;
        double vgdv = J.vg[i]/dv;
        double gam = avj - vgdv;
        if (gam > 30.0) gam = 30.0;
        else if (gam < -30.0) gam = -30.0;
        const double xp = exp(gam);
        if (vgdv > 30.0) vgdv = 30.0;
        const double xp0 = exp(-vgdv);
        const double A = 0.5*(1.0 - J.icfactor[m])*J.vg[i];
        const double As = vj < 0.0 ? -A : A;
        const double xp1 = 1 + xp - xp0;
        const double I = ( (g0 + gn*xp0*A/dv)*vj + gn*(vj - As)*(xp - xp0) ) / xp1;
        const double dIdv = (g0 + gn*xp0*A/dv)*(xp1 - absvj*xp)/(xp1*xp1) +
            gn*(xp*(absvj - A)/(dv*xp1*xp1) + (xp - xp0)/xp1) ;
        js.crhs = I - dIdv*vj;
        js.gqt = dIdv;
    }
    else if (rtype == 3) {
        if (absvj < J.vmore[i]) {
            const double g1 = PR(J.prm[JP_g1], i, t), g2 = PR(J.prm[JP_g2], i, t);
            double temp1 = vj*vj;
            double temp2 = temp1*temp1;
            js.gqt = g0 + 3.0*g1*temp1 + 5.0*g2*temp2;
            temp1 *= vj;
            temp2 *= vj;
            js.crhs = -2.0*temp1*g1 - 4.0*temp2*g2;
        }
        else {
            js.gqt = gn;
            const double cr1 = PR(J.prm[JP_cr1], i, t);
            js.crhs = (vj >= 0 ? cr1 : -cr1);
        }
    }
    else
        js.gqt = 0;
}

__global__ void __launch_bounds__(128)
k_load_jj(Consts K, HotView hot, VecView V, JJView J, const double *rhsOld)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;
    if (i >= J.n) return;
    if (t >= K.ntrials) return;
    int mode = hot.mode[t];
    if (!mode) return;
    TrialCtx c;
    make_ctx(c, K, hot, V, t, mode);
    const int st = J.state[i], m = J.model[i];
    const int pos = J.pos[i], neg = J.neg[i];
    const size_t ldt = K.ldt;
    const double criti = PR(J.prm[JP_criti], i, t);
    const double cap = PR(J.prm[JP_cap], i, t);
    const double gshunt = PR(J.prm[JP_gshunt], i, t);
    double *pdv = J.dvdt + (size_t)i*ldt + t;
    JJOp js;
    js.gqt = 0; js.crhs = 0; js.crt = 0; js.pfac = 0; js.vj = 0; js.phi = 0;
    bool shunt = true;
    double pint = 0.0;                          // the 4pi count the phase node adds (state 1)
    if (mode & MODEDC) {
        // jjload.cc:154-250.  Only the single MODEUIC|MODETRANOP|MODEINITJCT load ahead of a `uic`
        // transient gets here (tranan.cc:342-350); no solve follows, the states are what counts.
        js.vj = rhsOld[(size_t)pos*ldt + t] - rhsOld[(size_t)neg*ldt + t];
        if (J.ictype[m] != 0) {
            js.phi = js.vj;
            js.pfac = 1.0;
            js.crt = criti;
            shunt = false;
        }
        else
            jj_iv(J, i, m, t, js);
    }
    else {
        js.pfac = c.delta/JJ_PHI0_2PI;
        if (c.order > 1) js.pfac *= .5;
        // state 1 as JJdev::accept left it: phase folded into (-2pi, 2pi], 4pi count adjusted
        double phi1 = ST(c, 1, st + JS_phase);
        pint = ST(c, 1, st + JS_phsInt);
        if (!(mode & MODEINITFLOAT)) {
            const double twopi = M_PI + M_PI, fourpi = twopi + twopi;
            bool fold = false;
            if (phi1 >= twopi) { phi1 -= fourpi; pint += 1.0; fold = true; }
            else if (phi1 < -twopi) { phi1 += fourpi; pint -= 1.0; fold = true; }
            if (fold) { ST(c, 1, st + JS_phase) = phi1; ST(c, 1, st + JS_phsInt) = pint; }
        }
        if (mode & MODEINITFLOAT) {
            js.vj = rhsOld[(size_t)pos*ldt + t] - rhsOld[(size_t)neg*ldt + t];
            const double vprev = ST(c, 0, st + JS_voltage);
            double absvj = fabs(js.vj);
            const double absold = fabs(vprev);
            double maxvj = DMAX(absvj, absold);
            double temp = vprev - js.vj;
            double absdvj = fabs(temp);
            const double vless = J.vless[i], vmore = J.vmore[i];
            if (maxvj >= vless) {
                temp = J.delv[m]*0.5;
                if (absdvj > temp) {
                    // jj_limiting, jjload.cc:593-617
                    double vtop = vprev + temp;
                    double vbot = vprev - temp;
                    if (absold < vless) {
                        vtop = DMAX(vtop, vless);
                        vbot = DMIN(vbot, -vless);
                    }
                    else if (vprev > vmore) {
                        vtop += 1.0;
                        vbot = DMIN(vbot, vmore);
                    }
                    else if (vprev < -vmore) {
                        vtop = DMAX(vtop, -vmore);
                        vbot -= 1.0;
                    }
                    if (js.vj > vtop) js.vj = vtop;
                    else if (js.vj < vbot) js.vj = vbot;
                    absvj = fabs(js.vj);
                    maxvj = DMAX(absvj, absold);
                    temp = js.vj - vprev;
                    absdvj = fabs(temp);
                }
            }
            // the convergence test inside the load (jjload.cc:335-348): noncon only ever matters
            // as zero / nonzero, so "if nobody has flagged yet" reduces to a flag
            if (absdvj > K.opt.reltol*maxvj + K.opt.abstol) hot.noncon[t] = 1;
            js.phi = phi1;
            temp = js.vj;
            if (c.order > 1) temp += ST(c, 1, st + JS_voltage);
            js.phi += js.pfac*temp;
            js.crt = criti;
        }
        else if (mode & MODEINITPRED) {
            double y0 = ST(c, 1, st + JS_dvdt)*c.pred[0] + ST(c, 2, st + JS_dvdt)*c.pred[1];
            if (c.order != 1) y0 += ST(c, 1, st + JS_dvdt);
            double temp = ST(c, 1, st + JS_voltage);
            const double rag0 = c.order == 1 ? c.delta : .5*c.delta;
            js.vj = temp + rag0*y0;
            js.phi = phi1;
            temp = js.vj;
            if (c.order > 1) temp += ST(c, 1, st + JS_voltage);
            js.phi += js.pfac*temp;
            *pdv = dev_find_ceq(c, st + JS_voltage);
            js.crt = criti;
        }
        else if (mode & MODEINITTRAN) {
            if (mode & MODEUIC) {
                js.vj = J.initV[i];
                js.phi = J.initPhi[i];
                ST(c, 1, st + JS_voltage) = js.vj;
                ST(c, 1, st + JS_phase) = js.phi;
                ST(c, 1, st + JS_conI) = 0.0;
            }
            else {
                js.vj = 0;
                ST(c, 1, st + JS_voltage) = 0.0;
                js.phi = phi1;
                ST(c, 1, st + JS_conI) = 0.0;
            }
            *pdv = dev_find_ceq(c, st + JS_voltage);
            js.crt = criti;
        }
        else
            return;
        jj_iv(J, i, m, t, js);
        if (J.ictype[m] != 1) js.crt = 0.0;        // jj_ic, jjload.cc:747-787 (ictype 0)
    }
    // ---- jjstuff::jj_load, jjload.cc:790-871 ----
    double gqt = js.gqt, crhs = js.crhs, crt = js.crt;
    ST(c, 0, st + JS_voltage) = js.vj;          // (the 0.0 of jjload.cc:163 is overwritten here, :800)
    ST(c, 0, st + JS_phase) = js.phi;
    ST(c, 0, st + JS_conI) = 0.0;
    ST(c, 0, st + JS_crti) = js.crt;
    ST(c, 0, st + JS_qpi) = (mode & MODEDC) ? 0.0 : crhs + gqt*js.vj;
    if (!(mode & MODEDC)) ST(c, 0, st + JS_phsInt) = pint;
    double si, co;
    sincos(js.phi, &si, &co);
    double gcs = js.pfac*crt*co;
    if (J.pi[m]) { si = -si; gcs = -gcs; }
    crt *= si;
    const double delVdelT = *pdv;
    if (mode & MODEDC) {
        crhs += crt - gcs*js.phi;
        gqt = gcs;
    }
    else {
        crhs += crt - gcs*js.vj;
        gqt += gcs + c.ag[0]*cap;
        crhs += delVdelT*cap;
    }
    const int *cs = J.cslot + (size_t)i*JT_NSTAMP;
    const int *rs = J.rslot + (size_t)i*JR_NRHS;
    if (rs[JR_phs] >= 0) V.contrib[CADDR(K, rs[JR_phs], t)] = js.phi + (4*M_PI)*pint;
    if (cs[JT_PP] >= 0) V.contrib[CADDR(K, cs[JT_PP], t)] = gqt;
    if (cs[JT_PN] >= 0) V.contrib[CADDR(K, cs[JT_PN], t)] = -gqt;
    if (cs[JT_NP] >= 0) V.contrib[CADDR(K, cs[JT_NP], t)] = -gqt;
    if (cs[JT_NN] >= 0) V.contrib[CADDR(K, cs[JT_NN], t)] = gqt;
    if (rs[JR_pos] >= 0) V.contrib[CADDR(K, rs[JR_pos], t)] = -crhs;
    if (rs[JR_neg] >= 0) V.contrib[CADDR(K, rs[JR_neg], t)] = crhs;
    // the shunt implied by vshunt (jjload.cc:351-356, 462-467, 537-542); not in phase-mode DC
    const double gsh = (shunt && gshunt > 1e-12) ? gshunt : 0.0;
    if (cs[JT_RPP] >= 0) V.contrib[CADDR(K, cs[JT_RPP], t)] = gsh;
    if (cs[JT_RPN] >= 0) V.contrib[CADDR(K, cs[JT_RPN], t)] = -gsh;
    if (cs[JT_RNP] >= 0) V.contrib[CADDR(K, cs[JT_RNP], t)] = -gsh;
    if (cs[JT_RNN] >= 0) V.contrib[CADDR(K, cs[JT_RNN], t)] = gsh;
    if ((mode & MODEINITFLOAT) && !(mode & MODEDC)) {
        // ckt->integrate(JJvoltage, JJdelVdelT), jjload.cc:391
        const double q0 = js.vj;
        const double c0 = c.ag[0]*q0 + delVdelT;
        ST(c, 0, st + JS_dvdt) = c0;
        dev_trapcheck(c, st + JS_voltage);
        if ((mode & MODETRAN) && J.ictype[m] == 0) {
            // JJdev::trunc: a junction without supercurrent bounds the step by LTE (jjtrun.cc:52-64)
            double q[4] = { q0, ST(c, 1, st + JS_voltage), ST(c, 2, st + JS_voltage), ST(c, 3, st + JS_voltage) };
            double m0 = DBL_MAX, m1 = DBL_MAX;
            dev_lte_charge<true>(c, K.opt, q, c0, ST(c, 1, st + JS_dvdt), st + JS_voltage, m0, m1, c.rcp, c.ldt);
            atomic_min_pos(c.lte, m0);
            if (c.order + 1 <= K.opt.maxord) atomic_min_pos(c.lte + K.ldt, m1);
        }
    }
    // what JJdev::accept will scan if this load turns out to be the last of an accepted step
    if (J.ictype[m] != 0 && criti > 0) {
        const double av = fabs(js.vj);
        atomicMax((unsigned long long*)(hot.jj_vmax + (size_t)m*ldt + t), (unsigned long long)__double_as_longlong(av));
    }
}

// The instance part of JJdev::setup (jjset.cc:613-790) per (junction, trial): ics / area ->
// criti, cap, g0, gn, gs, the piecewise-linear offsets, the rtype 3 polynomial and the shunt.
__global__ void k_jj_temp(JJTempIn in, int ntrials, int ldt)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;
    if (i >= in.n) return;
    if (t >= (in.ts ? ldt : 1)) return;
    const int m = in.model[i];
    const int tt = t < ntrials ? t : 0;         // padding trials repeat trial 0
    double area;
    const double ics = PR(in.ics, i, tt);
    if (!isnan(ics)) {
        double v = ics;
        if (v < in.criti[m]/50 || v > in.criti[m]*50) v = in.criti[m];
        area = v/in.criti[m];
    }
    else if (!isnan(in.area0[i])) area = in.area0[i];
    else area = 1;
    const double criti = in.criti[m]*in.tcf[i]*area;
    const double sqrta = sqrt(area);
    const double cap = in.cap[m]*(area*(1.0 - in.cmu[m]) + sqrta*in.cmu[m]);
    const double gfac = area*(1.0 - in.gmu[m]) + sqrta*in.gmu[m];
    const double g0 = gfac/in.r0[m];
    const double gn = gfac/in.rn[m];
    const double gs = criti/(in.icfactor[m]*in.delv[m]);
    const double vless = in.vless[i], vmore = in.vmore[i];
    double cr1 = (g0 - gs)*vless;
    const double cr2 = criti/in.icfactor[m] + vless*g0 - vmore*gn;
    double gshunt = 0.0;
    if (in.vshunt[i] > 1e-12) {
        double g = criti/in.vshunt[i] - g0;
        if (g > 0.0) gshunt = g;
    }
    double g1 = 0, g2 = 0;
    if (in.rtype[m] == 3) {
        const double Gn = (criti/in.icfactor[m] + vless*g0)/vmore;
        const double temp = vmore*vmore;
        g1 = 0.5*(5.0*Gn - gs - 4.0*g0);
        g2 = (Gn - g0 - g1)/(temp*temp);
        g1 /= temp;
        cr1 = (Gn - gn)*vmore;
    }
    const size_t at = in.ts ? (size_t)i*ldt + t : (size_t)i;
    in.out[JP_criti][at] = criti; in.out[JP_cap][at] = cap; in.out[JP_g0][at] = g0;
    in.out[JP_gn][at] = gn; in.out[JP_gs][at] = gs; in.out[JP_g1][at] = g1; in.out[JP_g2][at] = g2;
    in.out[JP_cr1][at] = cr1; in.out[JP_cr2][at] = cr2; in.out[JP_gshunt][at] = gshunt;
}

void launch_jj_temp(cudaStream_t st, const JJTempIn &in, int ntrials, int ldt)
{
    if (!in.n) return;
    int nt = in.ts ? ldt : 1;
    k_jj_temp<<<grid2(nt, in.n), 128, 0, st>>>(in, ntrials, ldt);
}


// ---------------------------------------------------------------------------------------
// Tunnel-junction model (JJ level 3, mmjco): TJMdev::load + sTJMinstance::{tjm_load, tjm_init,
// tjm_newstep, tjm_update} (devlib/tjm/tjmload.cc:65-553) and the device half of TJMdev::accept
// (tjmacct.cc:43-169: phase fold, F -> Fprev, sin / cos of the accepted phase) which runs inside
// the first load after an accepted point (hot.accepted).  Complex arithmetic as include/ifcx.h.
// One thread per (junction, trial); the 2*nterms complex recurrences live in work[..][n][ldt].
// ---------------------------------------------------------------------------------------
struct Cxd { double re, im; };
__device__ __forceinline__ Cxd cx_mul(Cxd a, Cxd b) { Cxd r; r.re = a.re*b.re - a.im*b.im; r.im = a.im*b.re + a.re*b.im; return r; }

__global__ void __launch_bounds__(128)
k_load_tjm(Consts K, HotView hot, VecView V, TJMView J, const double *rhsOld)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;
    if (i >= J.n) return;
    if (t >= K.ntrials) return;
    int mode = hot.mode[t];
    if (!mode) return;
    if ((mode & MODEDC) && (mode & MODEUIC)) return;     // tjmload.cc:75-76: nothing in the `uic` start load
    TrialCtx c;
    make_ctx(c, K, hot, V, t, mode);
    const int st = J.state[i], m = J.model[i];
    const int pos = J.pos[i], neg = J.neg[i];
    if (mode & MODEDC) {
        // phase-mode DC, tjmload.cc:78-88 + tjm_load :419-430: I = Ic sin(phi), no shunt
        const double crt0 = PR(J.prm[TP_criti], i, t);
        const double ph = rhsOld[(size_t)pos*K.ldt + t] - rhsOld[(size_t)neg*K.ldt + t];
        double si, co;
        sincos(ph, &si, &co);
        const double cr = crt0*si;
        const double gq = 1.0*crt0*co;
        const double crh = cr - gq*ph;
        ST(c, 0, st + TS_voltage) = ph;              // tjm_load stores ts_vj (= the phase here)
        ST(c, 0, st + TS_phase) = ph;
        ST(c, 0, st + TS_crti) = cr;
        ST(c, 0, st + TS_qpi) = 0.0;
        const int *cs0 = J.cslot + (size_t)i*JT_NSTAMP;
        const int *rs0 = J.rslot + (size_t)i*JR_NRHS;
        if (rs0[JR_phs] >= 0) V.contrib[CADDR(K, rs0[JR_phs], t)] = ph + (4*M_PI)*ST(c, 1, st + TS_phsInt);
        if (cs0[JT_PP] >= 0) V.contrib[CADDR(K, cs0[JT_PP], t)] = gq;
        if (cs0[JT_PN] >= 0) V.contrib[CADDR(K, cs0[JT_PN], t)] = -gq;
        if (cs0[JT_NP] >= 0) V.contrib[CADDR(K, cs0[JT_NP], t)] = -gq;
        if (cs0[JT_NN] >= 0) V.contrib[CADDR(K, cs0[JT_NN], t)] = gq;
        if (rs0[JR_pos] >= 0) V.contrib[CADDR(K, rs0[JR_pos], t)] = -crh;
        if (rs0[JR_neg] >= 0) V.contrib[CADDR(K, rs0[JR_neg], t)] = crh;
        for (int k = JT_RPP; k <= JT_RNN; k++) if (cs0[k] >= 0) V.contrib[CADDR(K, cs0[k], t)] = 0.0;
        return;
    }
    const int NT = J.nterms, nt = J.nt[m];
    const size_t ldt = K.ldt;
    const size_t wstride = (size_t)J.n*ldt;     // one double plane of work[]
    double *W = J.work + (size_t)i*ldt + t;
#define WK(a, k, cpt) W[(size_t)(((a)*NT + (k))*2 + (cpt))*wstride]
    const double criti = PR(J.prm[TP_criti], i, t);
    const double cap = PR(J.prm[TP_cap], i, t);
    const double g0 = PR(J.prm[TP_g0], i, t);
    const double gshunt = PR(J.prm[TP_gshunt], i, t);
    const double gcrit = PR(J.prm[TP_gcrit], i, t);
    double *pdv = J.dvdt + (size_t)i*ldt + t;
    double *psin = J.sc_old + (size_t)i*ldt + t, *pcos = psin + wstride;
    double pfac = c.delta/JJ_PHI0_2PI;
    if (c.order > 1) pfac *= .5;
    double phi1 = ST(c, 1, st + TS_phase);
    double pint = ST(c, 1, st + TS_phsInt);
    if (hot.accepted[t] && !(mode & MODEINITTRAN)) {
        // TJMdev::accept for the point just accepted (its values are in state 1 now)
        const double twopi = M_PI + M_PI, fourpi = twopi + twopi;
        if (phi1 >= twopi) { phi1 -= fourpi; pint += 1.0; ST(c, 1, st + TS_phase) = phi1; ST(c, 1, st + TS_phsInt) = pint; }
        else if (phi1 < -twopi) { phi1 += fourpi; pint -= 1.0; ST(c, 1, st + TS_phase) = phi1; ST(c, 1, st + TS_phsInt) = pint; }
        double so, co;
        sincos(0.5*phi1, &so, &co);
        *psin = so; *pcos = co;
        for (int k = 0; k < nt; k++) {
            WK(TF_Fcprev, k, 0) = WK(TF_Fc, k, 0); WK(TF_Fcprev, k, 1) = WK(TF_Fc, k, 1);
            WK(TF_Fsprev, k, 0) = WK(TF_Fs, k, 0); WK(TF_Fsprev, k, 1) = WK(TF_Fs, k, 1);
        }
    }
    double vj, phi;
    bool newstep = false;
    if (mode & MODEINITFLOAT) {
        vj = rhsOld[(size_t)pos*ldt + t] - rhsOld[(size_t)neg*ldt + t];
        const double vprev = ST(c, 0, st + TS_voltage);
        const double maxvj = DMAX(fabs(vj), fabs(vprev));
        if (fabs(vprev - vj) > K.opt.reltol*maxvj + K.opt.abstol) hot.noncon[t] = 1;
        phi = phi1;
        double temp = vj;
        if (c.order > 1) temp += ST(c, 1, st + TS_voltage);
        phi += pfac*temp;
        // ckt->integrate BEFORE tjm_load stores the new vj (tjmload.cc:209-210): state 0 still
        // holds the previous iteration's voltage
        ST(c, 0, st + TS_dvdt) = c.ag[0]*vprev + *pdv;
        dev_trapcheck(c, st + TS_voltage);
    }
    else if (mode & MODEINITPRED) {
        double y0 = ST(c, 1, st + TS_dvdt)*c.pred[0] + ST(c, 2, st + TS_dvdt)*c.pred[1];
        if (c.order != 1) y0 += ST(c, 1, st + TS_dvdt);
        double temp = ST(c, 1, st + TS_voltage);
        const double rag0 = c.order == 1 ? c.delta : .5*c.delta;
        vj = temp + rag0*y0;
        phi = phi1;
        temp = vj;
        if (c.order > 1) temp += ST(c, 1, st + TS_voltage);
        phi += pfac*temp;
        *pdv = dev_find_ceq(c, st + TS_voltage);
        newstep = true;
    }
    else if (mode & MODEINITTRAN) {
        if (mode & MODEUIC) {
            vj = J.initV[i]; phi = J.initPhi[i];
            ST(c, 1, st + TS_voltage) = vj;
            ST(c, 1, st + TS_phase) = phi;
        }
        else { vj = 0; ST(c, 1, st + TS_voltage) = 0.0; phi = phi1; }     // tjmload.cc:310-316
        *pdv = dev_find_ceq(c, st + TS_voltage);
        // tjm_init, tjmload.cc:469-487
        double so, co;
        sincos(0.5*phi, &so, &co);
        *psin = so; *pcos = co;
        for (int k = 0; k < nt; k++) {
            WK(TF_Fcprev, k, 0) = co; WK(TF_Fcprev, k, 1) = 0.0;
            WK(TF_Fsprev, k, 0) = so; WK(TF_Fsprev, k, 1) = 0.0;
        }
        newstep = true;
    }
    else
        return;
    const double *cf = J.coef + (size_t)i*6*NT;
    if (newstep) {
        // tjm_newstep, tjmload.cc:490-505
        const double kdt = J.kgap[i]*J.omegaJ[i]*c.delta;
        for (int k = 0; k < nt; k++) {
            Cxd z; z.re = cf[2*k]*kdt; z.im = cf[2*k + 1]*kdt;
            const double d = exp(z.re);
            double sz, cz;
            sincos(z.im, &sz, &cz);
            Cxd ez; ez.re = d*cz; ez.im = d*sz;
            const double dd = z.re*z.re + z.im*z.im;
            const double er = ez.re - 1.0;
            Cxd zt; zt.re = (er*z.re + ez.im*z.im)/dd; zt.im = (ez.im*z.re - er*z.im)/dd;
            WK(TF_expz, k, 0) = ez.re; WK(TF_expz, k, 1) = ez.im;
            WK(TF_alpha0, k, 0) = zt.re - ez.re; WK(TF_alpha0, k, 1) = zt.im - ez.im;
            WK(TF_alpha1, k, 0) = 1.0 - zt.re;   WK(TF_alpha1, k, 1) = -zt.im;
        }
    }
    // tjm_update, tjmload.cc:508-553
    double s2, c2;
    sincos(0.5*phi, &s2, &c2);
    const double so = *psin, co = *pcos;
    double FcSp = 0.0, FcSq = 0.0, FsSp = 0.0, FsSq = 0.0;
    for (int k = 0; k < nt; k++) {
        Cxd ez, a0, a1, fcp, fsp, A, B;
        ez.re = WK(TF_expz, k, 0); ez.im = WK(TF_expz, k, 1);
        a0.re = WK(TF_alpha0, k, 0); a0.im = WK(TF_alpha0, k, 1);
        a1.re = WK(TF_alpha1, k, 0); a1.im = WK(TF_alpha1, k, 1);
        fcp.re = WK(TF_Fcprev, k, 0); fcp.im = WK(TF_Fcprev, k, 1);
        fsp.re = WK(TF_Fsprev, k, 0); fsp.im = WK(TF_Fsprev, k, 1);
        A.re = cf[2*(NT + k)]; A.im = cf[2*(NT + k) + 1];
        B.re = cf[2*(2*NT + k)]; B.im = cf[2*(2*NT + k) + 1];
        Cxd fc = cx_mul(ez, fcp);
        fc.re = (fc.re + a0.re*co) + a1.re*c2; fc.im = (fc.im + a0.im*co) + a1.im*c2;
        Cxd fs = cx_mul(ez, fsp);
        fs.re = (fs.re + a0.re*so) + a1.re*s2; fs.im = (fs.im + a0.im*so) + a1.im*s2;
        WK(TF_Fc, k, 0) = fc.re; WK(TF_Fc, k, 1) = fc.im;
        WK(TF_Fs, k, 0) = fs.re; WK(TF_Fs, k, 1) = fs.im;
        FcSp += A.re*fc.re - A.im*fc.im;
        FcSq += B.re*fc.re - B.im*fc.im;
        FsSp += A.re*fs.re - A.im*fs.im;
        FsSq += B.re*fs.re - B.im*fs.im;
    }
    const double fct = criti*J.kgap_rejpt[i];
    const double cp = J.ictype[m] > 0 ? fct*(s2*FcSp + c2*FsSp) : 0.0;
    const double cqp = J.rtype[m] > 0 ? fct*(s2*FcSq - c2*FsSq) : 0.0;
    // tjm_load, tjmload.cc:412-456
    ST(c, 0, st + TS_voltage) = vj;
    ST(c, 0, st + TS_phase) = phi;
    ST(c, 0, st + TS_phsInt) = pint;
    const double crhs = cp + cqp + (*pdv)*cap;
    const double gqt = gcrit + g0 + c.ag[0]*cap;
    ST(c, 0, st + TS_crti) = cp;
    ST(c, 0, st + TS_qpi) = cqp + (gcrit + g0)*vj;
    const int *cs = J.cslot + (size_t)i*JT_NSTAMP;
    const int *rs = J.rslot + (size_t)i*JR_NRHS;
    if (rs[JR_phs] >= 0) V.contrib[CADDR(K, rs[JR_phs], t)] = phi + (4*M_PI)*pint;
    if (cs[JT_PP] >= 0) V.contrib[CADDR(K, cs[JT_PP], t)] = gqt;
    if (cs[JT_PN] >= 0) V.contrib[CADDR(K, cs[JT_PN], t)] = -gqt;
    if (cs[JT_NP] >= 0) V.contrib[CADDR(K, cs[JT_NP], t)] = -gqt;
    if (cs[JT_NN] >= 0) V.contrib[CADDR(K, cs[JT_NN], t)] = gqt;
    if (rs[JR_pos] >= 0) V.contrib[CADDR(K, rs[JR_pos], t)] = -crhs;
    if (rs[JR_neg] >= 0) V.contrib[CADDR(K, rs[JR_neg], t)] = crhs;
    const double gsh = gshunt > 1e-12 ? gshunt : 0.0;
    if (cs[JT_RPP] >= 0) V.contrib[CADDR(K, cs[JT_RPP], t)] = gsh;
    if (cs[JT_RPN] >= 0) V.contrib[CADDR(K, cs[JT_RPN], t)] = -gsh;
    if (cs[JT_RNP] >= 0) V.contrib[CADDR(K, cs[JT_RNP], t)] = -gsh;
    if (cs[JT_RNN] >= 0) V.contrib[CADDR(K, cs[JT_RNN], t)] = gsh;
    if ((mode & MODEINITFLOAT) && (mode & MODETRAN) && J.ictype[m] == 0) {
        // TJMdev::trunc (tjmtrun.cc:52-64): the LTE of the voltage state as this load leaves it
        double q[4] = { vj, ST(c, 1, st + TS_voltage), ST(c, 2, st + TS_voltage), ST(c, 3, st + TS_voltage) };
        double m0 = DBL_MAX, m1 = DBL_MAX;
        dev_lte_charge<true>(c, K.opt, q, ST(c, 0, st + TS_dvdt), ST(c, 1, st + TS_dvdt), st + TS_voltage, m0, m1, c.rcp, c.ldt);
        atomic_min_pos(c.lte, m0);
        if (c.order + 1 <= K.opt.maxord) atomic_min_pos(c.lte + K.ldt, m1);
    }
    // what TJMdev::accept scans (every junction counts, tjmacct.cc:127-135)
    atomicMax((unsigned long long*)(hot.jj_vmax + (size_t)(J.model0 + m)*ldt + t), (unsigned long long)__double_as_longlong(fabs(vj)));
#undef WK
}

// What of sTJMinstance depends on the junction's area (tjmset.cc:600-660, 714; tjmload.cc:476-479)
__global__ void k_tjm_temp(TJMTempIn in, int ntrials, int ldt)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;
    if (i >= in.n) return;
    if (t >= (in.ts ? ldt : 1)) return;
    const int m = in.model[i];
    const int tt = t < ntrials ? t : 0;
    double area;
    const double ics = PR(in.ics, i, tt);
    if (!isnan(ics)) {
        double v = ics;
        if (v < in.criti[m]/50 || v > in.criti[m]*50) v = in.criti[m];
        area = v/in.criti[m];
    }
    else if (!isnan(in.area0[i])) area = in.area0[i];
    else area = 1;
    const double criti = in.criti[m]*area*in.tcf[i];
    const double sqrta = sqrt(area);
    const double gfac = area*(1.0 - in.gmu[m]) + sqrta*in.gmu[m];
    const double gqp = gfac/in.rsint[m];
    double g0 = in.r0[m] > 0.0 ? gfac/in.r0[m] : 0.0;
    const double tg0 = g0;
    if (g0 >= gqp) g0 -= gqp;
    double gshunt = 0.0;
    if (in.vshunt[i] > 1e-12) {
        double g = criti/in.vshunt[i] - DMAX(tg0, gqp);
        if (g > 0.0) gshunt = g;
    }
    const double cap = in.cap[m]*(area*(1.0 - in.cmu[m]) + sqrta*in.cmu[m]);
    const double gcrit = in.rtype[m] > 0 ? in.alphaN[i]*sqrt(cap*criti/JJ_PHI0_2PI) : 0.0;
    const size_t at = in.ts ? (size_t)i*ldt + t : (size_t)i;
    in.out[TP_criti][at] = criti; in.out[TP_cap][at] = cap; in.out[TP_g0][at] = g0;
    in.out[TP_gshunt][at] = gshunt; in.out[TP_gcrit][at] = gcrit;
}

void launch_tjm_temp(cudaStream_t st, const TJMTempIn &in, int ntrials, int ldt)
{
    if (!in.n) return;
    int nt = in.ts ? ldt : 1;
    k_tjm_temp<<<grid2(nt, in.n), 128, 0, st>>>(in, ntrials, ldt);
}

void launch_mos_temp(cudaStream_t st, int n, int ntrials, int ldt, const MosCard *cards,
    const RawParams &raw, const MosParamsOut &out)
{
    if (n) k_mos_temp<<<grid2(ntrials, n), 128, 0, st>>>(n, ntrials, ldt, cards, raw, out);
}
void launch_jjaccel_gate(cudaStream_t st, const Consts &K, const HotView &hot)
{
    k_jjaccel_gate<<<(K.ntrials + 127)/128, 128, 0, st>>>(K, hot);
}

void launch_ctl(cudaStream_t st, const Consts &K, const HotView &hot, TrialCtl *ctl,
    const CtlShared &sh, const CtlAux &aux, int phase)
{
    k_ctl<<<(K.ntrials + 127)/128, 128, 0, st>>>(K, hot, ctl, sh, aux, phase);
}
int loads_configure(const MosView &M)
{
    // per device: called from prepare_params with the batch's device current
    if (!M.n) return 0;
    int smem = (int)((size_t)M.stage_doubles*sizeof(double));
    if (cudaFuncSetAttribute(k_load_mos<MOS_MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
        cudaFuncSetAttribute(k_load_mos<MOS_MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return B200_E_NOMEM;
    return 0;
}
void launch_loads(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const MosView &M, const CapView &Cp, const SrcView &S, const double *rhsOld)
{
    if (M.n) {
        size_t smem = (size_t)M.stage_doubles*sizeof(double);
        dim3 g = grid2(K.ntrials, M.n, MOS_TPB);
        if (M.has_level6) k_load_mos<MOS_MINB, true><<<g, MOS_TPB, smem, st>>>(K, hot, V, M, rhsOld);
        else k_load_mos<MOS_MINB, false><<<g, MOS_TPB, smem, st>>>(K, hot, V, M, rhsOld);
    }
    if (Cp.n) k_load_cap<<<grid2(K.ntrials, Cp.n), 128, 0, st>>>(K, hot, V, Cp, rhsOld);
    if (Cp.L.nmut) k_load_mut<<<grid2(K.ntrials, Cp.L.nmut), 128, 0, st>>>(K, hot, V, Cp.L);
    if (Cp.L.n) k_load_ind<<<grid2(K.ntrials, Cp.L.n), 128, 0, st>>>(K, hot, V, Cp.L, rhsOld);
    if (S.n) k_load_src<<<grid2(K.ntrials, S.n), 128, 0, st>>>(K, hot, V, S);
    if (Cp.J.n) k_load_jj<<<grid2(K.ntrials, Cp.J.n), 128, 0, st>>>(K, hot, V, Cp.J, rhsOld);
    if (Cp.T.n) k_load_tjm<<<grid2(K.ntrials, Cp.T.n), 128, 0, st>>>(K, hot, V, Cp.T, rhsOld);
    if (Cp.R.n) k_load_rph<<<grid2(K.ntrials, Cp.R.n), 128, 0, st>>>(K, hot, V, Cp.R);
}
void launch_devconv(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const MosView &M, const double *rhs)
{
    if (M.n) k_devconv<<<grid2(K.ntrials, (M.n + DEVCONV_CHUNK - 1)/DEVCONV_CHUNK), 128, 0, st>>>(K, hot, V, M, rhs);
}
void launch_vec(cudaStream_t st, const Consts &K, const VecView &V, const CtlAux &aux,
    const MosView &M, double *rhsOld)
{
    k_vec<<<grid2(K.ntrials, K.n + 1), 128, 0, st>>>(K, V, aux, rhsOld);
    {
        // at most ~512 slots per thread, at least SV_CHUNKS chunks
        const int total = K.numStates + MC_NCACHE*M.n;
        const int chunks = std::max(SV_CHUNKS, (total + 511)/512);
        k_statevec<<<grid2(K.ntrials, chunks), 128, 0, st>>>(K, V, aux, M);
    }
}
void launch_count_active(cudaStream_t st, int ntrials, const int *active, int *count)
{
    cudaMemsetAsync(count, 0, sizeof(int), st);
    k_count_active<<<(ntrials + 127)/128, 128, 0, st>>>(ntrials, active, count);
}
// preloaded matrix constants (resistor conductances, source +-1): pseudo-contributions written once
__global__ void k_fill_const(Consts K, double *contrib, int nconst, const int *idx, const double *val)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= K.ldt) return;
    for (int i = 0; i < nconst; i++) contrib[CADDR(K, idx[i], t)] = val[i];
}

void launch_fill_const(cudaStream_t st, const Consts &K, double *contrib, int nconst, const int *idx,
    const double *val)
{
    if (nconst) k_fill_const<<<(K.ldt + 127)/128, 128, 0, st>>>(K, contrib, nconst, idx, val);
}

void launch_fill(cudaStream_t st, double *p, size_t n, double v)
{
    if (n) k_fill<<<(unsigned)((n + 255)/256), 256, 0, st>>>(p, n, v);
}
void launch_gather_waves(cudaStream_t st, int n, int ldt, int ntrials, int npts_max,
    const double *out, const int *npts, int nsel, const int *sel, double *waves)
{
    int nn = sel ? nsel : n;
    dim3 g((ntrials + 31)/32, (nn + 31)/32, npts_max), b(32, 8);
    k_gather_waves<<<g, b, 0, st>>>(n, ldt, ntrials, npts_max, out, npts, nsel, sel, waves);
}
void launch_export_counters(cudaStream_t st, int ntrials, const TrialCtl *ctl, int *counters, int *npts)
{
    k_export_counters<<<(ntrials + 127)/128, 128, 0, st>>>(ntrials, ctl, counters, npts);
}
