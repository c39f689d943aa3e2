// dev_ctl.cuh -- per-trial control state machine (one thread per trial).
//
// The reference runs, per trial, the nested loops
//     TRANanalysis::tran_dcoperation  (analysis/tranan.cc:321-468)
//       sCKT::op -> dynamic_src / dynamic_gmin  (ckt/ckt.cc:1492-1776, 1859-2011)
//       sTRANint::accept / setbp / step / interpolate  (tranan.cc:483-1116)
//         sCKT::NIiter  (ckt/niiter.cc:160-555)
// with a load and a factor+solve in the innermost body.  Here every trial advances through
// the SAME loops independently: one batch "tick" = one solve (of the matrix the previous
// tick loaded) + everything the reference does between that solve and the next load.  The
// loops are flattened into an explicit state machine: `pc` says where the trial is waiting,
// `ret` is the continuation NIiter returns to.  No trial ever waits for another.
#pragma once

#include "device_state.h"
#include "dev_load.cuh"

struct CtlIn {              // what the kernels of this tick produced for this trial
    int solve_err, nodeconv, devconv, noncon_load, trapbad, accel_exit;
    double lte0, lte1;
};

struct CtlShared {          // shared tables
    const Consts *K;
    const double *brks, *lat_offs, *lat_per;
    // SIN sources bound the step (srctrun.cc:45-66, inptran.cc:2666-2680)
    int nsin; const double *sin_tdl, *sin_per;
    // EXP sources: IFexpData::time_limit, inptran.cc:3064-3082 (two edges per source)
    int nexp; const double *exp_td, *exp_tau;
};

#define ST_RUNNING 0
#define ST_DONE    1        // status >= 100: reference error code

#define PC_UIC_LOAD 10

#define NI_PENDING (-1000)

__device__ inline bool ctl_is_math_or_iterlim(int e)
{
    return e == B200_E_ITERLIM || (e >= 117 && e <= 120);
}

// ---- breakpoints: ckt/breakpt.cc:177-221 ----
__device__ inline double ctl_nextbreak(const CtlShared &sh, double t)
{
    const Consts &K = *sh.K;
    double minbrk = K.minBreak;
    double dt0 = 0.0;
    for (int i = 0; i < K.nlat; i++) {
        double dt1 = sh.lat_offs[i];
        if (t > sh.lat_offs[i])
            dt1 += sh.lat_per[i]*ceil((t - sh.lat_offs[i])/sh.lat_per[i]);
        if (fabs(dt1 - t) < minbrk)
            dt1 += sh.lat_per[i];
        if (dt0 == 0.0 || dt1 < dt0)
            dt0 = dt1;
    }
    int i;
    for (i = 0; i < K.nbrks; i++)
        if (sh.brks[i] > t)
            break;
    while (i < K.nbrks && fabs(t - sh.brks[i]) < minbrk)
        i++;
    if (i == K.nbrks)
        return dt0;
    if (dt0 > 0.0 && dt0 < sh.brks[i])
        return dt0;
    return sh.brks[i];
}

// ---- nicomcof.cc:58-147 ----
__device__ inline int ctl_NIcomCof(TrialCtl &s, const Consts &K)
{
    if (K.method == 0) {
        if (s.order == 2) {
            double t0 = 1.0/(1.0 - K.opt.xmu);
            s.ag[0] = t0/s.delta;
            s.ag[1] = t0*K.opt.xmu;
            return 0;
        }
        if (s.order == 1) {
            s.ag[0] = 1.0/s.delta;
            s.ag[1] = -s.ag[0];
            return 0;
        }
        return B200_E_ORDER;
    }
    if (s.order < 1 || s.order > 6) return B200_E_ORDER;
    double mat[8][8];
    for (int i = 0; i < 7; i++) s.ag[i] = 0;
    s.ag[1] = -1.0/s.delta;
    int i, ord = s.order;
    for (i = 0; i <= ord; i++) mat[0][i] = 1;
    for (i = 1; i <= ord; i++) mat[i][0] = 0;
    double arg = 0;
    for (i = 1; i <= ord; i++) {
        arg += s.deltaOld[i-1];
        double arg1 = 1;
        for (int j = 1; j <= ord; j++) { arg1 *= arg/s.delta; mat[j][i] = arg1; }
    }
    for (i = 1; i <= ord; i++)
        for (int j = i+1; j <= ord; j++) {
            mat[j][i] /= mat[i][i];
            for (int k = i+1; k <= ord; k++) mat[j][k] -= mat[j][i]*mat[i][k];
        }
    for (i = 1; i <= ord; i++)
        for (int j = i+1; j <= ord; j++) s.ag[j] = s.ag[j] - mat[j][i]*s.ag[i];
    s.ag[ord] /= mat[ord][ord];
    for (i = ord-1; i >= 0; i--) {
        for (int j = i+1; j <= ord; j++) s.ag[i] = s.ag[i] - mat[i][j]*s.ag[j];
        s.ag[i] /= mat[i][i];
    }
    return 0;
}

__device__ inline double ctl_true_time(double tmp, int order)
{
    // ckt.cc:2310-2322
    if (order == 2) return sqrt(tmp);
    if (order == 1) return tmp;
    if (order == 3) return cbrt(tmp);
    if (order == 4) return sqrt(sqrt(tmp));
    return exp(log(tmp)/order);
}

// ---- NIiter entry: niiter.cc:160-183 (the load it requests runs later in this tick) ----
__device__ inline void ctl_ni_start(TrialCtl &s, int maxIter, int ret)
{
    s.iterno = 0;
    s.ipass = 0;
    s.maxIter = maxIter;
    s.ret = ret;
    s.pc = PC_NIITER;
    s.want_load = 1;
    if (s.mode & (MODEINITJCT | MODEINITTRAN))
        s.ni_reorder = 1;               // niiter.cc:181
}

// ---- NIiter after the solve: niiter.cc:353-540.  Returns NI_PENDING or the NIiter result ----
__device__ inline int ctl_ni_post_solve(TrialCtl &s, const Consts &K, const CtlIn &in)
{
    if (in.accel_exit) {
        // niiter.cc:535-539: left the loop after the load; rhsOld is the last solution
        s.vecops |= VO_KEEP_X;
        s.numIter += s.iterno;
        if (!(s.mode & MODEDC) && (s.mode & MODETRAN))
            s.tranIter += s.iterno;
        return 0;
    }
    s.iterno++;
    int error = NI_PENDING;
    // niiter.cc:257-331.  The pivot sequence is frozen here, so "order and factor" and "factor"
    // are the same numeric pass; what remains of NISHOULDREORDER is its effect on the iteration
    // count: a singular matrix met by a plain factor on the first iteration is loaded and
    // factored once more (now "reordering") before NIiter gives up.
    const bool reordering = s.ni_reorder != 0;
    if (in.solve_err == B200_E_SINGULAR && !reordering) {
        s.ni_reorder = 1;
        if (s.iterno == 1) {
            s.want_load = 1;            // load() again from the same rhsOld, then `continue`
            s.vecops |= VO_KEEP_X;
            return NI_PENDING;
        }
    }
    if (!in.solve_err && reordering)
        s.ni_reorder = 0;
    if (in.solve_err)
        error = in.solve_err;
    else if (s.iterno > s.maxIter)
        error = B200_E_ITERLIM;
    else {
        int noncon = in.noncon_load;
        // with jjaccel the transient relies on the devices' own convergence tests (niiter.cc:456)
        if ((!K.jjaccel || (s.mode & MODEDC)) && (s.mode & (MODEINITFIX | MODEINITFLOAT))) {
            if (noncon == 0 && s.iterno != 1)
                noncon = (in.nodeconv || in.devconv) ? 1 : 0;
            else
                noncon = 1;
        }
        if (s.mode & MODEINITFLOAT) {
            if ((s.mode & MODEDC) && K.hadNodeset) {
                if (s.ipass) noncon = s.ipass;
                s.ipass = 0;
            }
            if (noncon == 0)
                error = 0;
            else if ((s.mode & MODEDC) && s.iterno > 2 && K.opt.dcmu != 0.5)
                s.vecops |= VO_DCMU;    // niiter.cc:480-497
        }
        else if (s.mode & MODEINITPRED)
            s.mode = (s.mode & (~INITF)) | MODEINITFLOAT;
        else if (s.mode & MODEINITTRAN) {
            if (s.iterno <= 1) s.ni_reorder = 1;
            s.mode = (s.mode & (~INITF)) | MODEINITFLOAT;
        }
        else if (s.mode & MODEINITSMSIG)
            s.mode = (s.mode & (~INITF)) | MODEINITFLOAT;
        else if (s.mode & MODEINITJCT) {
            s.mode = (s.mode & (~INITF)) | MODEINITFIX;
            s.ni_reorder = 1;
        }
        else if (s.mode & MODEINITFIX) {
            if (noncon == 0)
                s.mode = (s.mode & (~INITF)) | MODEINITFLOAT;
            s.ipass = 1;
        }
        else
            error = B200_E_PANIC;
        if (error == NI_PENDING)
            s.want_load = 1;            // CKTnoncon = 0; load()   (niiter.cc:522-523)
    }
    if (error != NI_PENDING) {
        s.numIter += s.iterno;
        if (!(s.mode & MODEDC) && (s.mode & MODETRAN))
            s.tranIter += s.iterno;
    }
    return error;
}

// forward declarations of the flattened loops
__device__ inline void ctl_fail(TrialCtl &s, int code) { s.status = code; s.pc = PC_DONE; s.want_load = 0; }
__device__ inline void ctl_gmin_iter(TrialCtl &s, const Consts &K);
__device__ inline void ctl_src_iter(TrialCtl &s, const Consts &K);
__device__ inline void ctl_op_method_done(TrialCtl &s, const Consts &K, const CtlShared &sh, int err);
__device__ inline void ctl_tran_init(TrialCtl &s, const Consts &K, const CtlShared &sh);
__device__ inline void ctl_tran_loop_top(TrialCtl &s, const Consts &K, const CtlShared &sh);
__device__ inline void ctl_step_begin(TrialCtl &s, const Consts &K);
// A time point has been accepted.  The trial parks here until the other running trials of its
// group of 32 adjacent trials (one warp of every per-trial kernel) have accepted theirs too,
// then all of them run accept + the predictor load of the next step in the same tick: the lanes
// of a warp stay in the same Newton phase (same load mode, same history ring position), which
// is what keeps the device-load warps convergent and their state accesses coalesced.  Pure
// scheduling: no trial's arithmetic depends on it.
__device__ inline void ctl_sync_point(TrialCtl &s) { s.pc = PC_SYNC; s.want_load = 0; }

// ---- backup(): ckt.cc:1323-1362 ----
__device__ inline void ctl_backup_save(TrialCtl &s) { s.vecops |= VO_BK_SAVE; s.have_save = 1; }
__device__ inline void ctl_backup_restore(TrialCtl &s) { if (s.have_save) s.vecops |= VO_BK_RESTORE; }
__device__ inline void ctl_backup_clear(TrialCtl &s) { s.have_save = 0; }

// ---- dynamic_gmin: ckt.cc:1492-1610 ----
__device__ inline void ctl_gmin_start(TrialCtl &s, const Consts &K, int srcstep)
{
    s.g_srcstep = srcstep;
    s.mode = s.firstmode;
    s.vecops |= VO_ZERO_STATE0;
    s.g_factor = 10.0;
    s.g_oldgmin = 1e-2;
    s.diagGmin = s.g_oldgmin/s.g_factor;
    double gtarget = DMAX(K.opt.gmin, 0.0);
    if (gtarget < 1e-15) gtarget = 1e-15;
    s.g_gtarget = gtarget;
    s.g_succeeded = 0;
    ctl_gmin_iter(s, K);
}

__device__ inline void ctl_gmin_iter(TrialCtl &s, const Consts &K)
{
    s.g_iters_mark = s.numIter;
    ctl_ni_start(s, K.opt.itl_gmin, RET_GMIN_LOOP);
}

__device__ inline void ctl_gmin_done(TrialCtl &s, const Consts &K, const CtlShared &sh, int nonconv);

__device__ inline void ctl_gmin_after_loop(TrialCtl &s, const Consts &K, const CtlShared &sh, int nonconv)
{
    s.diagGmin = 0.0;
    ctl_backup_clear(s);
    if (s.g_succeeded)
        ctl_ni_start(s, s.iterlim, RET_GMIN_FINAL);
    else
        ctl_gmin_done(s, K, sh, nonconv);
}

__device__ inline void ctl_gmin_loop_ret(TrialCtl &s, const Consts &K, const CtlShared &sh, int nonconv)
{
    const double Factor = 10.0;
    int iters = s.numIter - s.g_iters_mark;
    if (nonconv) {
        if (!ctl_is_math_or_iterlim(nonconv)) {
            s.diagGmin = 0.0;
            ctl_backup_clear(s);
            ctl_gmin_done(s, K, sh, nonconv);
            return;
        }
        if (s.g_factor < 1.00005) {
            ctl_gmin_after_loop(s, K, sh, nonconv);
            return;
        }
        s.g_factor = sqrt(sqrt(s.g_factor));
        s.diagGmin = s.g_oldgmin/s.g_factor;
        ctl_backup_restore(s);
    }
    else {
        s.mode = s.continuemode;
        if (s.diagGmin <= s.g_gtarget) {
            s.g_succeeded = 1;
            ctl_gmin_after_loop(s, K, sh, nonconv);
            return;
        }
        ctl_backup_save(s);
        if (iters <= K.opt.itl_gmin/4) {
            s.g_factor *= sqrt(s.g_factor);
            if (s.g_factor > Factor) s.g_factor = Factor;
        }
        else if (iters > 3*K.opt.itl_gmin/4)
            s.g_factor = sqrt(s.g_factor);
        s.g_oldgmin = s.diagGmin;
        if (s.diagGmin < s.g_factor*s.g_gtarget) {
            s.g_factor = s.diagGmin/s.g_gtarget;
            s.diagGmin = s.g_gtarget;
        }
        else
            s.diagGmin /= s.g_factor;
    }
    ctl_gmin_iter(s, K);
}

// ---- dynamic_src: ckt.cc:1680-1776 ----
__device__ inline void ctl_src_start(TrialCtl &s, const Consts &K)
{
    s.mode = s.firstmode;
    s.srcFact = 0.001;
    s.s_raise = 0.001;
    s.s_lastraise = s.s_raise;
    s.s_convfact = 0;
    ctl_gmin_start(s, K, 1);
}

__device__ inline void ctl_src_finish(TrialCtl &s, const Consts &K, const CtlShared &sh, int nonconv)
{
    ctl_backup_clear(s);
    s.srcFact = 1;
    ctl_op_method_done(s, K, sh, (s.s_convfact != 1) ? nonconv : 0);
}

__device__ inline void ctl_src_iter(TrialCtl &s, const Consts &K)
{
    s.s_iters_mark = s.numIter;
    ctl_ni_start(s, K.opt.itl_src, RET_SRC_LOOP);
}

__device__ inline void ctl_src_after_gmin(TrialCtl &s, const Consts &K, const CtlShared &sh, int nonconv)
{
    if (!nonconv) {
        ctl_backup_save(s);
        s.srcFact = s.s_convfact + s.s_raise;
        ctl_src_iter(s, K);
    }
    else
        ctl_src_finish(s, K, sh, nonconv);
}

__device__ inline void ctl_src_loop_ret(TrialCtl &s, const Consts &K, const CtlShared &sh, int nonconv)
{
    int iters = s.numIter - s.s_iters_mark;
    s.mode = s.continuemode;
    if (nonconv) {
        if (!ctl_is_math_or_iterlim(nonconv) && nonconv != B200_E_SINGULAR) {
            ctl_src_finish(s, K, sh, nonconv);
            return;
        }
        if (s.srcFact - s.s_convfact < 1e-8) {
            ctl_src_finish(s, K, sh, nonconv);
            return;
        }
        s.s_raise = 0.1*s.s_lastraise;
        if (s.s_raise > 0.01) s.s_raise = 0.01;
        s.srcFact = s.s_convfact + s.s_raise;
        s.s_lastraise = s.s_raise;
        ctl_backup_restore(s);
    }
    else {
        s.s_convfact = s.srcFact;
        ctl_backup_save(s);
        s.srcFact = s.s_convfact + s.s_raise;
        s.s_lastraise = s.s_raise;
        if (iters <= K.opt.itl_src/4) s.s_raise *= 1.5;
        else if (iters > 3*K.opt.itl_src/4) s.s_raise *= 0.5;
    }
    if (s.srcFact > 1) s.srcFact = 1;
    if (s.s_raise >= 1e-7 && s.s_convfact < 1)
        ctl_src_iter(s, K);
    else
        ctl_src_finish(s, K, sh, nonconv);
}

__device__ inline void ctl_gmin_done(TrialCtl &s, const Consts &K, const CtlShared &sh, int nonconv)
{
    if (s.g_srcstep)
        ctl_src_after_gmin(s, K, sh, nonconv);
    else
        ctl_op_method_done(s, K, sh, nonconv);
}

// ---- sCKT::op: ckt.cc:1859-2011 (default dynamic stepping paths) ----
// op_stage: 0 = first method running, 1 = second method running
__device__ inline int ctl_op_method(const Consts &K, int stage)
{
    // returns 1 = dynamic_src, 2 = dynamic_gmin, 0 = none
    int first = K.opt.gminfirst ? 2 : 1, second = K.opt.gminfirst ? 1 : 2;
    int m = stage == 0 ? first : second;
    if (m == 1 && K.opt.srcsteps < 0) return 0;
    if (m == 2 && K.opt.gminsteps < 0) return 0;
    return m;
}

__device__ inline void ctl_op_finish(TrialCtl &s, const Consts &K, const CtlShared &sh, int err)
{
    if (err) { ctl_fail(s, err); return; }
    if (K.analysis == 1) {
        // dcoan.cc:117-119: dump(0.0) = CKTrhsOld (the INITSMSIG load only feeds AC analysis)
        s.vecops |= VO_DUMP_DIRECT;
        s.ip_out0 = 0;
        s.npts = 1;
        s.status = ST_DONE;
        s.pc = PC_DONE;
        s.want_load = 0;
        return;
    }
    // tranan.cc:378-389: a phase-mode operating point leaves PHASE on the nodes of junctions and
    // inductors; their voltage, going into the transient, is zero
    if (K.jjDCphase) s.vecops |= VO_ZERO_PHASE;
    ctl_tran_init(s, K, sh);
}

__device__ inline void ctl_op_run_stage(TrialCtl &s, const Consts &K, const CtlShared &sh)
{
    while (s.op_stage < 2) {
        int m = ctl_op_method(K, s.op_stage);
        if (m == 1) { ctl_src_start(s, K); return; }
        if (m == 2) { ctl_gmin_start(s, K, 0); return; }
        s.op_stage++;
    }
    ctl_op_finish(s, K, sh, 0);
}

__device__ inline void ctl_op_method_done(TrialCtl &s, const Consts &K, const CtlShared &sh, int err)
{
    // err_return(), ckt.cc:1825-1852
    bool first = (s.op_stage == 0);
    bool stop = true;
    if (err == B200_E_ITERLIM && first) stop = false;
    if (stop) { ctl_op_finish(s, K, sh, err); return; }
    s.op_stage++;
    ctl_op_run_stage(s, K, sh);
}

__device__ inline void ctl_op_start(TrialCtl &s, const Consts &K, const CtlShared &sh,
    int firstmode, int continuemode, int iterlim)
{
    s.firstmode = firstmode; s.continuemode = continuemode; s.iterlim = iterlim;
    if (K.opt.gminsteps < 0 && K.opt.srcsteps < 0) {
        s.mode = firstmode;
        ctl_ni_start(s, iterlim, RET_OP_DIRECT);
        return;
    }
    s.op_stage = 0;
    ctl_op_run_stage(s, K, sh);
}

// ---- transient: tranan.cc ----
__device__ inline void ctl_tran_init(TrialCtl &s, const Consts &K, const CtlShared &sh)
{
    // tranan.cc:407-435
    s.vecops |= VO_HIST_INIT;
    s.timePts++;
    s.order = 1;
    if (K.initDelta > 0)                          // CKTinitDelta, tranan.cc:416-417
        s.delta = K.initDelta;
    else {
        s.delta = DMIN(s.t_delmax, K.t_step);
        if (!K.nocut)
            s.delta *= 0.1;
    }
    for (int i = 0; i < 7; i++) s.deltaOld[i] = s.delta;
    s.saveDelta = s.delta;
    s.mode = (s.mode & MODEUIC) | MODETRAN | MODEINITTRAN;
    s.ag[0] = s.ag[1] = 0;
    ctl_sync_point(s);
}

__device__ inline void ctl_check_delmin_and_step(TrialCtl &s, const Consts &K)
{
    // tranan.cc:940-949
    if (s.delta <= K.delmin) {
        if (s.olddelta > K.delmin)
            s.delta = K.delmin;
        else {
            ctl_fail(s, B200_E_TIMESTEP);
            return;
        }
    }
    ctl_step_begin(s, K);
}

__device__ inline void ctl_step_begin(TrialCtl &s, const Consts &K)
{
    // tranan.cc:686-712
    s.olddelta = s.delta;
    if (K.maxStep < s.t_delmax)
        s.t_delmax = K.maxStep;
    s.time += s.delta;
    s.deltaOld[0] = s.delta;
    int e = ctl_NIcomCof(s, K);
    if (e) { ctl_fail(s, e); return; }
    // predict(), ckt.cc:2020-2087
    s.pred[1] = -s.delta/s.deltaOld[1];
    s.pred[0] = 1 - s.pred[1];
    s.trapCheck = K.opt.trapcheck && K.method == 0 && s.order == 2;
    s.niter_mark = s.numIter;
    ctl_ni_start(s, K.opt.itl4, RET_TRAN_STEP);
}

__device__ inline void ctl_tran_loop_top(TrialCtl &s, const Consts &K, const CtlShared &sh)
{
    // ---- sTRANint::accept, tranan.cc:483-554 ----
    s.vecops |= VO_SAVE_SPARE;
    s.devMaxDelta = s.jj_dmax;                    // sCKT::accept, ckt.cc:363 + JJdev::accept
    s.did_accept = 1;
    s.sols_base = (s.sols_base + 7) & 7;          // rotate sols, ckt.cc:372-377
    s.vecops |= VO_ACCEPT_SOLS;
    if (s.time > s.breaks[0]) {
        // breakClr, breakpt.cc:57-72
        s.breaks[0] = ctl_nextbreak(sh, s.breaks[0]);
        s.breaks[1] = ctl_nextbreak(sh, s.breaks[0]);
    }
    s.accepted++;
    s.brk = 0;
    // ---- interpolate, tranan.cc:963-1116 ----
    bool endit = false;
    bool skip = false;
    if (s.mode & MODEUIC) {
        if (s.t_firsttime)
            skip = true;
        else if (s.npts == 0 && K.t_start == 0.0) {
            s.vecops |= VO_DUMP_DIRECT;
            s.ip_out0 = s.npts;
            s.npts++;
            s.t_check += K.t_step;
            skip = true;
        }
    }
    if (!skip) {
        double hd = .5*K.delmin;
        double finalmax = K.t_stop + hd, finalmin = K.t_stop - hd;
        double time = s.time + hd;
        int degree = K.opt.interplev;
        if (degree > K.opt.maxord) degree = K.opt.maxord;
        s.ip_degree = degree;
        s.ip_time = s.time;
        s.ip_c2 = s.deltaOld[0];
        s.ip_c3 = s.ip_c2 + s.deltaOld[1];
        s.ip_c4 = s.ip_c3 + s.deltaOld[2];
        s.ip_tcheck0 = s.t_check;
        s.ip_out0 = s.npts;
        int cnt = 0;
        bool nochange = false;
        for (; s.t_check <= time; s.t_check += K.t_step) {
            if (s.t_check < finalmax) {
                cnt++;
                if (finalmin < s.t_check) { nochange = true; break; }
            }
            else { nochange = true; break; }
        }
        if (!nochange && s.t_check >= finalmax)
            nochange = true;
        if (cnt > 0) {
            s.ip_n = cnt;
            s.npts += cnt;
            s.vecops |= VO_INTERP;
        }
        endit = nochange;
    }
    if (endit) {
        s.status = ST_DONE;
        s.pc = PC_DONE;
        s.want_load = 0;
        return;
    }
    // ---- setbp, tranan.cc:558-654 (STEP_NORMAL) ----
    s.brk = 0;
    if (s.time == 0.0 || s.breaks[0] > 0) {
        if ((s.time == s.breaks[0]) || (s.breaks[0] - s.time <= K.delmin)) {
            s.order = 1;
            s.t_ordcnt = K.opt.spice3 ? 0 : 1;
            double ttmp = s.breaks[1] - s.breaks[0];
            if (s.saveDelta < ttmp) ttmp = s.saveDelta;
            if (!K.nocut && !K.jjaccel) ttmp *= 0.1;
            double td1 = 0.1*s.saveDelta;
            if (ttmp < td1) ttmp = td1;
            if (ttmp < s.delta) s.delta = ttmp;
            ttmp = 2.0*K.delmin;
            if (s.delta < ttmp) s.delta = ttmp;
        }
        else if (s.time + 1.01*s.delta >= s.breaks[0]) {
            double del = s.breaks[0] - s.time;
            if (del > K.delmin) {
                s.saveDelta = s.delta;
                s.delta = del;
                s.brk = 1;
            }
        }
    }
    // rotate delta vector and states, tranan.cc:451-460
    for (int i = 5; i >= 0; i--) s.deltaOld[i+1] = s.deltaOld[i];
    s.deltaOld[0] = s.delta;
    s.st_base = (s.st_base + K.NS - 1) % K.NS;
    // ---- sTRANint::step entry, tranan.cc:683 ----
    s.intDevMaxDelta = s.devMaxDelta;
    ctl_step_begin(s, K);
}

__device__ inline void ctl_reject_common(TrialCtl &s)
{
    s.vecops |= VO_RESTORE_OLD;
}

__device__ inline void ctl_lte_ret(TrialCtl &s, const Consts &K, const CtlShared &sh, const CtlIn &in);

__device__ inline void ctl_tran_step_ret(TrialCtl &s, const Consts &K, const CtlShared &sh,
    int cverr, const CtlIn &in)
{
    // tranan.cc:714-742
    double bound_step = s.devMaxDelta;
    if (bound_step == 0.0 || (s.intDevMaxDelta > 0.0 && s.intDevMaxDelta < bound_step))
        s.devMaxDelta = s.intDevMaxDelta;
    s.timePts++;
    s.lastIters = s.numIter - s.niter_mark;
    s.mode = (s.mode & MODEUIC) | MODETRAN | MODEINITPRED;
    if (s.t_firsttime)
        s.vecops |= VO_FIRSTTIME;
    if (cverr) {
        if (cverr != B200_E_ITERLIM) { ctl_fail(s, cverr); return; }
        s.time -= s.delta;
        s.rejected++;
        s.iterCut++;
        s.delta *= 0.125;
        if (s.t_firsttime)
            s.mode = (s.mode & MODEUIC) | MODETRAN | MODEINITTRAN;
        s.order = 1;
        s.brk = 0;
        ctl_reject_common(s);
        ctl_check_delmin_and_step(s, K);
        return;
    }
    if (s.t_firsttime) {
        s.t_firsttime = 0;
        ctl_sync_point(s);
        return;
    }
    if (K.jjaccel) {
        // tranan.cc:787-792: no LTE; the junctions' phase-change bound is the step
        s.delta = s.devMaxDelta;
        if (s.order < K.opt.maxord) s.order++;
        ctl_sync_point(s);
        return;
    }
    // sCKT::trunc: the per-charge bounds were min-reduced by the load kernels of the tick whose
    // matrix has just converged (in.lte0: this order, in.lte1: order + 1)
    ctl_lte_ret(s, K, sh, in);
}

// device-independent part of sCKT::trunc for one order: SRC trunc (srctrun.cc:45-66) + root
__device__ inline double ctl_finish_trunc(const TrialCtl &s, const CtlShared &sh, double timetemp,
    int order)
{
    double time_step = HUGE_VAL;
    for (int i = 0; i < sh.nsin; i++) {
        if (s.time >= sh.sin_tdl[i]) {
            double per = sh.sin_per[i];              // dphimax over the angular frequency (host-evaluated)
            if (time_step > per) time_step = per;
        }
    }
    for (int i = 0; i < sh.nexp; i++) {
        const double td = sh.exp_td[i], tau = sh.exp_tau[i];
        if (s.time >= td && s.time < td + 5.0*tau) {
            double f = (.2*M_PI)/(2*M_PI);
            double per = tau*f*exp((s.time - td)/tau);
            if (time_step > per) time_step = per;
        }
    }
    if (time_step < HUGE_VAL) {
        double tmp = time_step;
        for (int i = 1; i < order; i++) time_step *= tmp;
    }
    if (timetemp > time_step) timetemp = time_step;
    return ctl_true_time(timetemp, order);
}

__device__ inline void ctl_lte_ret(TrialCtl &s, const Consts &K, const CtlShared &sh, const CtlIn &in)
{
    // tranan.cc:794-938
    double newd = ctl_finish_trunc(s, sh, in.lte0, s.order);
    if (in.trapbad) {
        s.time -= s.delta;
        s.rejected++;
        s.trapCut++;
        s.delta *= 0.125;
        s.order = 1;
        ctl_reject_common(s);
        s.t_ordcnt = K.opt.spice3 ? 0 : 1;
        // `continue`: skips the delmin check (tranan.cc:825)
        ctl_step_begin(s, K);
        return;
    }
    if (K.opt.spice3) {
        if (newd > 2.0*s.delta) newd = 2.0*s.delta;
        if (s.devMaxDelta > 0.0 && s.devMaxDelta < newd) newd = s.devMaxDelta;
        if (newd > .9*s.delta) {
            if (s.order == 1) {
                if (s.t_ordcnt) s.t_ordcnt--;
                else {
                    s.order++;
                    if (s.order > K.opt.maxord) s.order = K.opt.maxord;
                    newd = ctl_finish_trunc(s, sh, s.order == 1 ? in.lte0 : in.lte1, s.order);
                    if (newd > 2.0*s.delta) newd = 2.0*s.delta;
                    if (s.devMaxDelta > 0.0 && s.devMaxDelta < newd) newd = s.devMaxDelta;
                    if (newd <= 1.05*s.delta) s.order = 1;
                }
            }
            s.delta = newd;
            if (s.delta > s.t_delmax) s.delta = s.t_delmax;
            ctl_sync_point(s);
            return;
        }
        s.time -= s.delta;
        s.rejected++;
        s.delta = newd;
        ctl_reject_common(s);
        ctl_check_delmin_and_step(s, K);
        return;
    }
    if (newd > .5*s.delta) {
        double twodelta = 2.0*s.delta;
        if (s.order < K.opt.maxord) {
            if (s.t_ordcnt) s.t_ordcnt--;
            else {
                s.order++;
                double newdtmp = ctl_finish_trunc(s, sh, in.lte1, s.order);
                if (newdtmp <= 1.05*DMIN(newd, twodelta)) s.order--;
                else newd = newdtmp;
            }
        }
        if (newd > twodelta) newd = twodelta;
        if (s.devMaxDelta > 0.0 && s.devMaxDelta < newd) newd = s.devMaxDelta;
        if (newd > s.t_delmax) newd = s.t_delmax;
        s.delta = newd;
        ctl_sync_point(s);
        return;
    }
    s.time -= s.delta;
    s.rejected++;
    if (s.devMaxDelta > 0.0 && s.devMaxDelta < newd) newd = s.devMaxDelta;
    if (newd > s.t_delmax) newd = s.t_delmax;
    s.delta = newd;
    s.brk = 0;
    ctl_reject_common(s);
    ctl_check_delmin_and_step(s, K);
}

// ---- entry points called by the control kernels ----
__device__ inline void ctl_begin(TrialCtl &s, const Consts &K, const CtlShared &sh)
{
    // tran_dcoperation, tranan.cc:328-372
    s.time = 0; s.delta = 0; s.brk = 0;
    s.breaks[0] = K.break0; s.breaks[1] = K.break1;
    s.t_firsttime = 1;
    s.t_check = K.t_start;
    s.t_delmax = K.maxStep;
    s.srcFact = 1.0;
    s.status = ST_RUNNING;
    if (K.analysis == 1)
        ctl_op_start(s, K, sh, MODEDCOP | MODEINITJCT, MODEDCOP | MODEINITFLOAT, K.opt.itl1);
    else if (K.uic) {
        s.mode = MODEUIC | MODETRANOP | MODEINITJCT;
        s.want_load = 1;
        s.pc = PC_UIC_LOAD;
    }
    else
        ctl_op_start(s, K, sh, MODETRANOP | MODEINITJCT, MODETRANOP | MODEINITFLOAT, K.opt.itl1);
}

__device__ inline void ctl_after_solve(TrialCtl &s, const Consts &K, const CtlShared &sh,
    const CtlIn &in, HotView &hot, int t)
{
    if (s.pc == PC_UIC_LOAD) {
        ctl_tran_init(s, K, sh);
        return;
    }
    if (s.pc != PC_NIITER)
        return;
    int r = ctl_ni_post_solve(s, K, in);
    if (r == NI_PENDING)
        return;
    switch (s.ret) {
    case RET_OP_DIRECT:
        if (!r) s.mode = s.continuemode;
        ctl_op_finish(s, K, sh, r);
        break;
    case RET_GMIN_LOOP:  ctl_gmin_loop_ret(s, K, sh, r); break;
    case RET_GMIN_FINAL: ctl_gmin_done(s, K, sh, r); break;
    case RET_SRC_LOOP:   ctl_src_loop_ret(s, K, sh, r); break;
    case RET_TRAN_STEP:  ctl_tran_step_ret(s, K, sh, r, in); break;
    default: ctl_fail(s, B200_E_PANIC); break;
    }
}
