// kernels.cu -- CUDA kernels of the batched Newton engine (sm_100a, FP64, -fmad=false).
//
// Thread mappings:
//   k_load_*    one thread per (instance, trial), trial fastest: every global access of a warp
//               is 32 consecutive doubles of one [slot][trial] row -> one 256-B line.
//   k_solve     one warp per trial: gather-assemble A and b into shared memory (ordered,
//               atomics-free), level-scheduled LU refactor with the frozen pivot order,
//               triangular solves, node convergence test.  Lanes = independent columns/rows
//               of one elimination level.
//   k_ctl       one thread per trial: the flattened NIiter / op / tran-step state machine.
//   k_vec, k_statevec, k_trunc, k_devconv: one thread per (element, trial), gated by
//               per-trial request flags.
#include <cuda_runtime.h>
#include <float.h>
#include <stdio.h>

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
    int i = blockIdx.y;
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
    int nag = (K.method == 0) ? 2 : 7;
    for (int k = 0; k < nag; k++) c.ag[k] = hot.ag[(size_t)k*K.ldt + t];
    c.pred[0] = hot.pred[t]; c.pred[1] = hot.pred[K.ldt + t];
    int base = hot.st_base[t];
    size_t plane = (size_t)K.numStates*K.ldt;
    for (int k = 0; k < K.NS; k++) {
        int phys = base + k; if (phys >= K.NS) phys -= K.NS;
        c.sp[k] = V.states + phys*plane + t;
    }
}

__global__ void __launch_bounds__(128)
k_load_mos(Consts K, HotView hot, VecView V, MosView M, const double *rhsOld)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (t >= K.ntrials) return;
    int mode = hot.mode[t];
    if (!mode) return;
    TrialCtx c;
    make_ctx(c, K, hot, V, t, mode);
    MosInst m;
    m.type = M.type[i]; m.off = M.off[i]; m.st = M.state[i];
    m.d = M.d[i]; m.g = M.g[i]; m.s = M.s[i]; m.b = M.b[i]; m.dp = M.dp[i]; m.sp = M.sp[i];
    m.tPhi = PR(M.prm[MP_tPhi], i, t); m.tVbi = PR(M.prm[MP_tVbi], i, t);
    m.tVto = PR(M.prm[MP_tVto], i, t); m.gamma = PR(M.prm[MP_gamma], i, t);
    m.lambda = PR(M.prm[MP_lambda], i, t); m.beta = PR(M.prm[MP_beta], i, t);
    m.oxideCap = PR(M.prm[MP_oxideCap], i, t);
    m.tSourceSatCur = PR(M.prm[MP_tSourceSatCur], i, t);
    m.tDrainSatCur = PR(M.prm[MP_tDrainSatCur], i, t);
    m.sourceVcrit = PR(M.prm[MP_sourceVcrit], i, t); m.drainVcrit = PR(M.prm[MP_drainVcrit], i, t);
    m.Cbd = PR(M.prm[MP_Cbd], i, t); m.Cbdsw = PR(M.prm[MP_Cbdsw], i, t);
    m.Cbs = PR(M.prm[MP_Cbs], i, t); m.Cbssw = PR(M.prm[MP_Cbssw], i, t);
    m.tBulkPot = PR(M.prm[MP_tBulkPot], i, t); m.tDepCap = PR(M.prm[MP_tDepCap], i, t);
    m.f2d = PR(M.prm[MP_f2d], i, t); m.f3d = PR(M.prm[MP_f3d], i, t); m.f4d = PR(M.prm[MP_f4d], i, t);
    m.f2s = PR(M.prm[MP_f2s], i, t); m.f3s = PR(M.prm[MP_f3s], i, t); m.f4s = PR(M.prm[MP_f4s], i, t);
    m.gsOvl = PR(M.prm[MP_gsOvl], i, t); m.gdOvl = PR(M.prm[MP_gdOvl], i, t);
    m.gbOvl = PR(M.prm[MP_gbOvl], i, t); m.dcon = PR(M.prm[MP_dcon], i, t);
    m.scon = PR(M.prm[MP_scon], i, t); m.mult = PR(M.prm[MP_mult], i, t);
    m.vt = PR(M.prm[MP_vt], i, t); m.mj = PR(M.prm[MP_mj], i, t); m.mjsw = PR(M.prm[MP_mjsw], i, t);
    m.icVDS = PR(M.prm[MP_icVDS], i, t); m.icVGS = PR(M.prm[MP_icVGS], i, t);
    m.icVBS = PR(M.prm[MP_icVBS], i, t);
    size_t cstride = (size_t)M.n*K.ldt;
    double *cp = M.cache + (size_t)i*K.ldt + t;
    m.mode = (int)cp[MC_mode*cstride];
    m.von = cp[MC_von*cstride]; m.vdsat = cp[MC_vdsat*cstride]; m.cd = cp[MC_cd*cstride];
    m.cbs = cp[MC_cbs*cstride]; m.cbd = cp[MC_cbd*cstride]; m.gm = cp[MC_gm*cstride];
    m.gds = cp[MC_gds*cstride]; m.gmbs = cp[MC_gmbs*cstride]; m.gbd = cp[MC_gbd*cstride];
    m.gbs = cp[MC_gbs*cstride]; m.capbd = cp[MC_capbd*cstride]; m.capbs = cp[MC_capbs*cstride];

    MosOut o;
    int r = dev_mos_load(c, m, o, rhsOld + t, K.opt);
    if (r & 2) {
        cp[MC_mode*cstride] = (double)m.mode;
        cp[MC_von*cstride] = m.von; cp[MC_vdsat*cstride] = m.vdsat; cp[MC_cd*cstride] = m.cd;
        cp[MC_cbs*cstride] = m.cbs; cp[MC_cbd*cstride] = m.cbd; cp[MC_gm*cstride] = m.gm;
        cp[MC_gds*cstride] = m.gds; cp[MC_gmbs*cstride] = m.gmbs; cp[MC_gbd*cstride] = m.gbd;
        cp[MC_gbs*cstride] = m.gbs; cp[MC_capbd*cstride] = m.capbd; cp[MC_capbs*cstride] = m.capbs;
    }
    if (r & 4)
        hot.noncon[t] = 1;
    const int *cs = M.cslot + (size_t)i*MT_NSTAMP;
    const int *rs = M.rslot + (size_t)i*MRH_NRHS;
    if (r & 1) {
#pragma unroll
        for (int k = 0; k < MT_NSTAMP; k++) {
            int s = cs[k];
            if (s >= 0) V.contrib[(size_t)s*K.ldt + t] = o.m[k];
        }
#pragma unroll
        for (int k = 0; k < MRH_NRHS; k++) {
            int s = rs[k];
            if (s >= 0) V.rcontrib[(size_t)s*K.ldt + t] = o.r[k];
        }
    }
    else {
        for (int k = 0; k < MT_NSTAMP; k++) {
            int s = cs[k];
            if (s >= 0) V.contrib[(size_t)s*K.ldt + t] = 0.0;
        }
        for (int k = 0; k < MRH_NRHS; k++) {
            int s = rs[k];
            if (s >= 0) V.rcontrib[(size_t)s*K.ldt + t] = 0.0;
        }
    }
}

__global__ void __launch_bounds__(128)
k_load_cap(Consts K, HotView hot, VecView V, CapView Cp, const double *rhsOld)
{
    // capload.cc:75-205 (linear capacitor)
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = blockIdx.y;
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
            double qcap = capac*(rhsOld[(size_t)pos*K.ldt + t] - rhsOld[(size_t)neg*K.ldt + t]);
            ST(c, 0, st) = qcap;
            dev_integrate2(c, st, *pc);
            ceq = dev_find_ceq(c, st);
            *pc = ceq;
            geq = *pg;
            stamp = true;
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
    if (cs[0] >= 0) V.contrib[(size_t)cs[0]*K.ldt + t] = geq;
    if (cs[1] >= 0) V.contrib[(size_t)cs[1]*K.ldt + t] = geq;
    if (cs[2] >= 0) V.contrib[(size_t)cs[2]*K.ldt + t] = -geq;
    if (cs[3] >= 0) V.contrib[(size_t)cs[3]*K.ldt + t] = -geq;
    if (rs[0] >= 0) V.rcontrib[(size_t)rs[0]*K.ldt + t] = stamp ? -ceq : 0.0;
    if (rs[1] >= 0) V.rcontrib[(size_t)rs[1]*K.ldt + t] = stamp ? ceq : 0.0;
}

__global__ void __launch_bounds__(128)
k_load_src(Consts K, HotView hot, VecView V, SrcView S)
{
    // srcload.cc:67-105, 215-262
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = blockIdx.y;
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
        if (mode & (MODEDCOP | MODETRANOP))
            rhs *= srcFact;
    }
    const int *rs = S.rslot + (size_t)i*2;
    if (S.kind[i] == B200_SRC_V) {
        if (rs[0] >= 0) V.rcontrib[(size_t)rs[0]*K.ldt + t] = rhs;
    }
    else {
        if (rs[0] >= 0) V.rcontrib[(size_t)rs[0]*K.ldt + t] = -rhs;
        if (rs[1] >= 0) V.rcontrib[(size_t)rs[1]*K.ldt + t] = rhs;
    }
}

// ---------------------------------------------------------------------------------------
// MOSdev::convTest, gated per trial
__global__ void __launch_bounds__(128)
k_devconv(Consts K, HotView hot, VecView V, MosView M, const double *rhs)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (t >= K.ntrials) return;
    if (!hot.need_devconv[t]) return;
    if (hot.devconv[t]) return;                  // "no reason to continue" (mosconv.cc:112)
    TrialCtx c;
    make_ctx(c, K, hot, V, t, hot.mode[t]);
    MosInst m;
    m.type = M.type[i]; m.off = M.off[i]; m.st = M.state[i];
    m.g = M.g[i]; m.b = M.b[i]; m.dp = M.dp[i]; m.sp = M.sp[i];
    size_t cstride = (size_t)M.n*K.ldt;
    const double *cp = M.cache + (size_t)i*K.ldt + t;
    m.mode = (int)cp[MC_mode*cstride];
    m.cd = cp[MC_cd*cstride]; m.cbs = cp[MC_cbs*cstride]; m.cbd = cp[MC_cbd*cstride];
    m.gm = cp[MC_gm*cstride]; m.gds = cp[MC_gds*cstride]; m.gmbs = cp[MC_gmbs*cstride];
    m.gbd = cp[MC_gbd*cstride]; m.gbs = cp[MC_gbs*cstride];
    if (dev_mos_conv(c, m, rhs + t, K.opt.reltol, K.opt.abstol))
        hot.devconv[t] = 1;
}

// ---------------------------------------------------------------------------------------
// LTE: sCKT::terr (ckt.cc:2326-2389) for every integrated charge, min-reduced per trial.
__device__ __forceinline__ void atomic_min_pos(double *addr, double v)
{
    // positive finite doubles order like their bit patterns
    atomicMin((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}

__device__ inline double dev_terr(const Consts &K, const double *const *sp, int ldt, int qcap,
    int order, const double *deltaOld, double delta)
{
    const double gearCoeff[] = { .5, .2222222222, .1363636364, .096, .07299270073, .05830903790 };
    const double trapCoeff[] = { .5, .08333333333 };
    int ccap = qcap + 1;
    double y0 = fabs(sp[0][(size_t)ccap*ldt]), y1 = fabs(sp[1][(size_t)ccap*ldt]);
    double volttol = K.opt.abstol + K.opt.reltol*DMAX(y0, y1);
    y0 = fabs(sp[0][(size_t)qcap*ldt]); y1 = fabs(sp[1][(size_t)qcap*ldt]);
    double chargetol = DMAX(y0, y1);
    if (chargetol < K.opt.chgtol) chargetol = K.opt.chgtol;
    chargetol *= K.opt.reltol/delta;
    double tol = DMAX(volttol, chargetol);
    double diff[8], deltmp[8];
    for (int i = order + 1; i >= 0; i--) diff[i] = sp[i][(size_t)qcap*ldt];
    for (int i = 0; i <= order; i++) deltmp[i] = deltaOld[i];
    int j = order;
    for (;;) {
        for (int i = 0; i <= j; i++) diff[i] = (diff[i] - diff[i+1])/deltmp[i];
        if (--j < 0) break;
        for (int i = 0; i <= j; i++) deltmp[i] = deltmp[i+1] + deltaOld[i];
    }
    double del;
    if (K.method == 0) del = fabs(diff[0]*trapCoeff[order-1]);
    else del = fabs(diff[0]*gearCoeff[order-1]);
    del = K.opt.trtol*tol/DMAX(K.opt.abstol, del);
    return del;
}

__global__ void __launch_bounds__(128)
k_trunc(Consts K, HotView hot, VecView V, int nq, const int *qslots)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int q = blockIdx.y;
    if (t >= K.ntrials || q >= nq) return;
    int want = hot.want_lte[t];
    if (!want) return;
    int base = hot.st_base[t];
    size_t plane = (size_t)K.numStates*K.ldt;
    const double *sp[8];
    for (int k = 0; k < K.NS; k++) {
        int phys = base + k; if (phys >= K.NS) phys -= K.NS;
        sp[k] = V.states + phys*plane + t;
    }
    double dOld[7];
    for (int k = 0; k < 7; k++) dOld[k] = hot.deltaOld[(size_t)k*K.ldt + t];
    int order = hot.order[t];
    double delta = hot.delta[t];
    double d0 = dev_terr(K, sp, K.ldt, qslots[q], order, dOld, delta);
    atomic_min_pos(&hot.lte[t], d0);
    if (want == 2 && order + 1 <= K.opt.maxord) {
        double d1 = dev_terr(K, sp, K.ldt, qslots[q], order + 1, dOld, delta);
        atomic_min_pos(&hot.lte[K.ldt + t], d1);
    }
}

// ---------------------------------------------------------------------------------------
// assemble + loadGmin + LU refactor + solve + NIconvTest node part.  One warp per trial.
__global__ void __launch_bounds__(SOLVE_WARPS*32)
k_solve(Consts K, HotView hot, VecView V, SolveView Sv, const double *Xold, double *Xnew)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x*SOLVE_WARPS + warp;
    if (t >= K.ntrials) return;
    if (!hot.do_solve[t]) return;
    const int ldt = K.ldt, n = K.n, nlu = K.nnzLU;
    double *LU = smem + (size_t)warp*(nlu + n);
    double *y = LU + nlu;
    const int mode = hot.mode[t];
    const double diagGmin = hot.diaggmin[t];
    const bool ns_on = (mode & MODEDC) && (mode & (MODEINITJCT | MODEINITFIX));
    const bool ic_on = (mode & MODEDC) && (mode & MODETRANOP) && !(mode & MODEUIC);

    // ---- matrix: A[k] = Ainit[k] + sum of contributions in reference load order ----
    for (int s = lane; s < nlu; s += 32) {
        int p = Sv.lu_src[s];
        double v = 0.0;
        if (p >= 0) {
            v = Sv.Ainit[p];
            int qe = Sv.cg_ptr[p+1];
            for (int q = Sv.cg_ptr[p]; q < qe; q++)
                v += V.contrib[(size_t)Sv.cg_idx[q]*ldt + t];
            int dn = Sv.diag_node[p];           // node number if p is a diagonal, else 0
            if (dn) {
                // ckt.cc:1177-1204 nodeset / ic pinning
                if (ns_on && Sv.ns_given[dn]) v += 1.0;
                if (ic_on && Sv.ic_given[dn]) v += 1.0;
                if (Sv.ntype[dn] == 0) {
                    // loadGmin, ckt.cc:1221-1263 / spspice.cc:259-300
                    if (diagGmin != 0.0)
                        v += diagGmin;
                    else if (K.opt.forcegmin && (v >= 0.0 && v < K.opt.gmin))
                        v = K.opt.gmin;
                    else if (v > K.opt.gmax)
                        v = K.opt.gmax;
                }
            }
        }
        LU[s] = v;
    }
    // ---- rhs (permuted) ----
    for (int k = lane; k < n; k += 32) {
        int row = Sv.yperm_in[k];
        double v = 0.0;
        int qe = Sv.rg_ptr[row+1];
        for (int q = Sv.rg_ptr[row]; q < qe; q++)
            v += V.rcontrib[(size_t)Sv.rg_idx[q]*ldt + t];
        if (ns_on && Sv.ns_given[row]) v += Sv.ns_val[row];
        if (ic_on && Sv.ic_given[row]) v += Sv.ic_val[row];
        y[k] = v;
    }
    __syncwarp();

    // ---- refactor: levels of independent columns, lane per column ----
    int bad = 0;
    for (int r = 0; r < Sv.fac_nrounds; r++) {
        if (Sv.fac_sync[r]) __syncwarp();
        const LuRound rd = Sv.fac_rounds[r];
        const LuOp *ops = Sv.fac_ops + rd.op_off + lane;
        for (int s = 0; s < rd.nsteps; s++) {
            LuOp op = ops[s*32];
            if (op.dst >= 0)
                LU[op.dst] -= LU[op.a]*LU[op.b];
        }
        int dg = Sv.fac_fin_diag[rd.fin_off + lane];
        if (dg >= 0) {
            double piv = LU[dg];
            if (piv == 0.0 || !isfinite(piv))
                bad = 1;
            int lb = Sv.fac_fin_lbeg[rd.fin_off + lane], le = Sv.fac_fin_lend[rd.fin_off + lane];
            for (int q = lb; q < le; q++)
                LU[q] = LU[q]/piv;
        }
    }
    __syncwarp();
    // ---- forward: lane per row ----
    for (int r = 0; r < Sv.fwd_nrounds; r++) {
        if (Sv.fwd_sync[r]) __syncwarp();
        const LuRound rd = Sv.fwd_rounds[r];
        int row = Sv.fwd_fin_row[rd.fin_off + lane];
        if (row >= 0) {
            const LuOp *ops = Sv.fwd_ops + rd.op_off + lane;
            double acc = y[row];
            for (int s = 0; s < rd.nsteps; s++) {
                LuOp op = ops[s*32];
                if (op.dst >= 0)
                    acc -= LU[op.a]*y[op.b];
            }
            y[row] = acc;
        }
    }
    __syncwarp();
    // ---- backward ----
    for (int r = 0; r < Sv.bwd_nrounds; r++) {
        if (Sv.bwd_sync[r]) __syncwarp();
        const LuRound rd = Sv.bwd_rounds[r];
        int row = Sv.bwd_fin_row[rd.fin_off + lane];
        if (row >= 0) {
            const LuOp *ops = Sv.bwd_ops + rd.op_off + lane;
            double acc = y[row];
            for (int s = 0; s < rd.nsteps; s++) {
                LuOp op = ops[s*32];
                if (op.dst >= 0)
                    acc -= LU[op.a]*y[op.b];
            }
            y[row] = acc/LU[Sv.bwd_fin_diag[rd.fin_off + lane]];
        }
    }
    __syncwarp();
    // ---- unpermute, store, NIconvTest node part (niconv.cc:58-91) ----
    int fail = 0, nonfin = 0;
    for (int k = lane; k < n; k += 32) {
        int node = Sv.yperm_out[k];
        double cur = y[k];
        double old = Xold[(size_t)node*ldt + t];
        Xnew[(size_t)node*ldt + t] = cur;
        if (!isfinite(cur)) nonfin = 1;
        double dd = fabs(old - cur);
        double tol = K.opt.reltol*DMAX(fabs(old), fabs(cur)) +
            (Sv.ntype[node] == 0 ? K.opt.vntol : K.opt.abstol);
        if (dd > tol) fail = 1;
    }
    fail = __any_sync(0xffffffffu, fail);
    nonfin = __any_sync(0xffffffffu, nonfin);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        Xnew[t] = 0.0;                           // ground slot
        int err = bad ? B200_E_SINGULAR : (nonfin ? B200_E_MATHOVF : 0);
        hot.solve_err[t] = err;
        hot.nodeconv[t] = fail;
        hot.devconv[t] = 0;
        int it1 = hot.iterno[t] + 1;
        hot.need_devconv[t] = (!err && !fail && (mode & (MODEINITFIX | MODEINITFLOAT)) &&
            hot.noncon[t] == 0 && it1 != 1) ? 1 : 0;
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
    // st_base in force when state-vector ops are requested = value at the start of the
    // control phase (a rotation, if any, comes after the request inside the same call)
    if (phase != 2) aux.vo_stbase[t] = s.st_base;
    if (phase == 0) {
        // begin
        s.vecops = 0;
        s.want_load = 0;
        ctl_begin(s, K, sh);
    }
    else if (phase == 1) {
        if (s.status != ST_RUNNING) {
            // finished trials: no load, no solve, and no stale vector-op requests
            hot.mode[t] = 0; hot.do_solve[t] = 0; hot.want_lte[t] = 0; aux.vecops[t] = 0;
            return;
        }
        s.vecops = 0;
        s.want_load = 0;
        CtlIn in;
        in.solve_err = hot.do_solve[t] ? hot.solve_err[t] : 0;
        in.nodeconv = hot.nodeconv[t];
        in.devconv = hot.devconv[t];
        in.noncon_load = hot.noncon[t];
        in.trapbad = hot.trapbad[t];
        in.lte0 = in.lte1 = 0;
        hot.want_lte[t] = 0;
        ctl_after_solve(s, K, sh, in, hot, t);
        if (s.pc == PC_WAIT_LTE) {
            hot.lte[t] = DBL_MAX;
            hot.lte[K.ldt + t] = DBL_MAX;
            // k_trunc needs order/delta/deltaOld/st_base of the step just solved
            hot.order[t] = s.order;
            hot.delta[t] = s.delta;
            hot.st_base[t] = s.st_base;
            for (int k = 0; k < 7; k++) hot.deltaOld[(size_t)k*K.ldt + t] = s.deltaOld[k];
        }
    }
    else {
        if (s.status != ST_RUNNING || s.pc != PC_WAIT_LTE) return;
        CtlIn in;
        in.solve_err = 0; in.nodeconv = 0; in.devconv = 0; in.noncon_load = 0;
        in.trapbad = hot.trapbad[t];
        in.lte0 = hot.lte[t];
        in.lte1 = hot.lte[K.ldt + t];
        hot.want_lte[t] = 0;
        int keep = s.vecops;
        ctl_lte_ret(s, K, sh, in);
        s.vecops |= keep;
    }
    // publish for the vector / load kernels of this tick
    if (!(phase == 1 && s.pc == PC_WAIT_LTE)) {
        bool load = s.want_load && s.status == ST_RUNNING;
        hot.mode[t] = load ? s.mode : 0;
        hot.do_solve[t] = (load && s.pc == PC_NIITER) ? 1 : 0;
        hot.active[t] = (s.status == ST_RUNNING);
        if (load) {
            hot.order[t] = s.order;
            hot.st_base[t] = s.st_base;
            hot.noncon[t] = 0;
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
        }
    }
    // vector-op requests
    if (phase != 2 || s.vecops) {
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
    int i = blockIdx.y;                          // node 0..n
    if (t >= K.ntrials) return;
    int ops = aux.vecops[t];
    if (!ops) return;
    size_t ldt = K.ldt;
    size_t at = (size_t)i*ldt + t;
    size_t vplane = (size_t)(K.n + 1)*ldt;
    if (ops & VO_BK_SAVE) V.oldSol[at] = rhsOld[at];
    if (ops & VO_BK_RESTORE) rhsOld[at] = V.oldSol[at];
    if (ops & VO_RESTORE_OLD) { if (i) rhsOld[at] = V.spare[at]; }
    if (ops & VO_SAVE_SPARE) { if (i) V.spare[at] = rhsOld[at]; }
    int sb = aux.sols_base[t];
    if (ops & VO_ACCEPT_SOLS) V.sols[(size_t)sb*vplane + at] = rhsOld[at];
    if (!i) return;
    size_t oplane = (size_t)K.n*ldt;
    if (ops & VO_DUMP_DIRECT)
        V.out[(size_t)aux.ip_out0[t]*oplane + (size_t)(i-1)*ldt + t] = rhsOld[at];
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
            V.out[(size_t)(out0 + p)*oplane + (size_t)(i-1)*ldt + t] = v;
        }
    }
}

__global__ void __launch_bounds__(128)
k_statevec(Consts K, VecView V, CtlAux aux, MosView M)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int e = blockIdx.y;                          // state slot, then MOS cache entries
    if (t >= K.ntrials) return;
    int ops = aux.vecops[t];
    if (!(ops & (VO_ZERO_STATE0 | VO_BK_SAVE | VO_BK_RESTORE | VO_HIST_INIT | VO_FIRSTTIME))) return;
    size_t ldt = K.ldt;
    if (e >= K.numStates) {
        size_t at = (size_t)(e - K.numStates)*ldt + t;
        if (ops & VO_BK_SAVE) M.save[at] = M.cache[at];
        if (ops & VO_BK_RESTORE) M.cache[at] = M.save[at];
        return;
    }
    int base = aux.vo_stbase[t];
    size_t plane = (size_t)K.numStates*ldt;
    size_t at = (size_t)e*ldt + t;
    double *sp[8];
    for (int k = 0; k < K.NS; k++) {
        int phys = base + k; if (phys >= K.NS) phys -= K.NS;
        sp[k] = V.states + phys*plane;
    }
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
static inline dim3 grid2(int ntrials, int ny, int bs = 128) { return dim3((ntrials + bs - 1)/bs, ny > 0 ? ny : 1); }

void launch_mos_temp(cudaStream_t st, int n, int ntrials, int ldt, const MosCard *cards,
    const RawParams &raw, const MosParamsOut &out)
{
    if (n) k_mos_temp<<<grid2(ntrials, n), 128, 0, st>>>(n, ntrials, ldt, cards, raw, out);
}
void launch_ctl(cudaStream_t st, const Consts &K, const HotView &hot, TrialCtl *ctl,
    const CtlShared &sh, const CtlAux &aux, int phase)
{
    k_ctl<<<(K.ntrials + 127)/128, 128, 0, st>>>(K, hot, ctl, sh, aux, phase);
}
void launch_loads(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const MosView &M, const CapView &Cp, const SrcView &S, const double *rhsOld)
{
    if (M.n) k_load_mos<<<grid2(K.ntrials, M.n), 128, 0, st>>>(K, hot, V, M, rhsOld);
    if (Cp.n) k_load_cap<<<grid2(K.ntrials, Cp.n), 128, 0, st>>>(K, hot, V, Cp, rhsOld);
    if (S.n) k_load_src<<<grid2(K.ntrials, S.n), 128, 0, st>>>(K, hot, V, S);
}
void launch_devconv(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const MosView &M, const double *rhs)
{
    if (M.n) k_devconv<<<grid2(K.ntrials, M.n), 128, 0, st>>>(K, hot, V, M, rhs);
}
void launch_trunc(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    int nq, const int *qslots)
{
    if (nq) k_trunc<<<grid2(K.ntrials, nq), 128, 0, st>>>(K, hot, V, nq, qslots);
}
int solve_smem_bytes(const Consts &K) { return (int)((size_t)SOLVE_WARPS*(K.nnzLU + K.n)*sizeof(double)); }
int launch_solve(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const SolveView &Sv, const double *Xold, double *Xnew)
{
    int smem = solve_smem_bytes(K);
    static int configured = -1;
    if (configured != smem) {
        if (cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return B200_E_NOMEM;
        configured = smem;
    }
    k_solve<<<(K.ntrials + SOLVE_WARPS - 1)/SOLVE_WARPS, SOLVE_WARPS*32, smem, st>>>(K, hot, V, Sv, Xold, Xnew);
    return 0;
}
void launch_vec(cudaStream_t st, const Consts &K, const VecView &V, const CtlAux &aux,
    const MosView &M, double *rhsOld)
{
    k_vec<<<grid2(K.ntrials, K.n + 1), 128, 0, st>>>(K, V, aux, rhsOld);
    k_statevec<<<grid2(K.ntrials, K.numStates + MC_NCACHE*M.n), 128, 0, st>>>(K, V, aux, M);
}
void launch_count_active(cudaStream_t st, int ntrials, const int *active, int *count)
{
    cudaMemsetAsync(count, 0, sizeof(int), st);
    k_count_active<<<(ntrials + 127)/128, 128, 0, st>>>(ntrials, active, count);
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
