// device_state.h -- device-visible views of the batch (plain pointers; passed to kernels by
// value).  All per-trial arrays are [slot][ldt], trial fastest.
#pragma once

#include "b200_internal.h"

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#endif
#endif

// Row (instance / node / entry) index of a CTA: grids are (trial blocks, rows) and a deck may hold
// more rows than gridDim.y allows (65535), so the rows continue in z.
#define GRID_ROW ((int)(blockIdx.y + blockIdx.z*gridDim.y))
#define GRID_Y_MAX 65535
#ifdef __CUDACC__
static inline dim3 grid_rows(unsigned nx, int ny)
{
    if (ny < 1) ny = 1;
    return dim3(nx, ny < GRID_Y_MAX ? ny : GRID_Y_MAX, (ny + GRID_Y_MAX - 1)/GRID_Y_MAX);
}
#endif

// ---- per-trial control block (AoS; touched only by the control kernel, one thread/trial) ----
enum { PC_IDLE = 0, PC_NIITER, PC_DONE, PC_SYNC };
enum { RET_NONE = 0, RET_OP_DIRECT, RET_GMIN_LOOP, RET_GMIN_FINAL, RET_SRC_LOOP, RET_TRAN_STEP };

// vector-op request bits (consumed by k_vec / k_statevec right after the control kernel)
#define VO_SAVE_SPARE   0x001   // rhsSpare <- rhsOld                 (tranan.cc:501-502)
#define VO_ACCEPT_SOLS  0x002   // sols[0] <- rhsOld after rotation   (ckt.cc:372-381)
#define VO_RESTORE_OLD  0x004   // rhsOld <- rhsSpare                 (tranan.cc:775-776)
#define VO_BK_SAVE      0x008   // backup(DEV_SAVE)                   (ckt.cc:1329-1338)
#define VO_BK_RESTORE   0x010   // backup(DEV_RESTORE)                (ckt.cc:1339-1352)
#define VO_ZERO_STATE0  0x020   // memset(CKTstate0,0)                (ckt.cc:1499)
#define VO_HIST_INIT    0x040   // states[1..maxord+1] <- state0      (tranan.cc:408-411,431)
#define VO_FIRSTTIME    0x080   // state2 (,state3) <- state1         (tranan.cc:736-742)
#define VO_INTERP       0x100   // emit interpolated output points    (tranan.cc:963-1116)
#define VO_DUMP_DIRECT  0x200   // emit rhsOld as an output point     (tranan.cc:977-981)
#define VO_ZERO_PHASE   0x800   // zero the phase nodes of rhsOld after a phase-mode DCOP (tranan.cc:378-389)
#define VO_DCMU         0x1000  // rhsOld <- (dcmu+.5)*rhsOld + (.5-dcmu)*rhs   (niiter.cc:480-497)
#define VO_KEEP_X       0x400   // trial parked at a sync point: carry its rhsOld over the global
                                //  rhs/rhsOld buffer flip of this tick

struct TrialCtl {
    // sCKTPOD / sTRANint scalars (circuit.h:1562-1662, trandefs.h:93-151)
    double time, delta, deltaOld[7], saveDelta, ag[7], pred[2];
    double srcFact, diagGmin, devMaxDelta, breaks[2];
    double t_check, t_delmax, olddelta, intDevMaxDelta;
    // DCOP steppers (ckt.cc:1492-1776)
    double g_factor, g_oldgmin, g_gtarget, s_raise, s_lastraise, s_convfact;
    // interpolation request for k_vec
    double ip_time, ip_c2, ip_c3, ip_c4, ip_tcheck0;
    int ip_n, ip_out0, ip_degree;
    int mode, order, iterno, ipass, maxIter, pc, ret, status;
    int firstmode, continuemode, iterlim;       // op() arguments
    int op_stage;                               // which stepper of op() is running
    int g_srcstep, g_iters_mark, g_succeeded, s_iters_mark;
    int t_firsttime, t_ordcnt, brk, trapCheck;
    int sols_base, st_base;
    int accepted, rejected, numIter, timePts, tranIter, iterCut, trapCut, npts, lastIters;
    int niter_mark;
    double jj_dmax;                             // JJdev::accept's step bound for the accept being processed (0 = none)
    int did_accept;                             // ctl_tran_loop_top ran in this control pass
    int ni_reorder;                             // CKTniState & NISHOULDREORDER (niiter.cc:181, 303-331, 501-509)
    int vecops;
    int have_save;
    int want_load;                              // a load must follow in this tick
};

// ---- "hot" per-trial scalars read by the load / solve kernels (SoA, coalesced) ----
struct HotView {
    int    *mode;       // [ldt] CKTmode, or 0 if the trial takes no load this tick
    int    *order;
    int    *st_base;
    int    *noncon;     // set (|=1) by load kernels (incNoncon), cleared by control
    int    *trapbad;    // set by integrate() when trapcheck fires
    int    *trapcheck;  // CKTtrapCheck for this load
    int    *active;     // 1 while the trial is running
    int    *do_solve;   // 1: the load of this tick is a NIiter load, solve it next tick
    int    *iterno;     // value before the solve of this tick
    double *time, *delta, *srcfact, *diaggmin;
    double *ag;         // [7][ldt]
    double *pred;       // [2][ldt]
    double *deltaOld;   // [7][ldt]   (LTE kernel)
    // results of the solve kernel
    int    *nodeconv;   // 1 = node test failed
    int    *solve_err;  // 0 / B200_E_SINGULAR / B200_E_MATHINV
    int    *need_devconv;
    int    *devconv;    // 1 = device convTest failed
    // LTE
    double *lte_rcp;    // [LR_NROWS][ldt] divisors of the LTE divided differences + reciprocals
    int    *sync;       // 1: trial is parked at a time-point boundary (or finished)
    double *lte;        // [2][ldt] min over charges for order, order+1 (un-normalised)
    // Josephson junctions: max |vj| over the junctions of a model in the trial's latest load
    // (what JJdev::accept scans state 0 for, jjacct.cc:120-133); zeroed when a load is published
    double *jj_vmax;    // [nmodels][ldt]   (RSJ models first, then the TJM models)
    int    *accel_exit; // jjaccel: the load of the previous tick converged by the devices' own test, no solve (niiter.cc:535-539)
    int    *accepted;   // 1: the load published in this tick directly follows an accepted time point
                        //    (the device-side half of XXXdev::accept runs inside that load)
};

struct Consts {
    b200_options opt;
    int n, nnz, numStates, ldt, ntrials, NS /* state ring size = maxord+2 */;
    int nnzLU, ncslot, nrslot, NC /* contributions per trial, assembly order */;
    double t_stop, t_step, t_start, maxStep, delmin, minBreak, break0, break1;
    int nocut, uic, hadNodeset, npoints_max, method;
    int analysis;       // 0: transient (tranan.cc), 1: DC operating point only (dcoan.cc)
    int nbrks, nlat;
    // Josephson junctions: CKTinitDelta (jjic.cc:78-104) and the per-model accept constants
    double initDelta;
    int nout;           // unknowns kept per output point (n, or the `.save` list)
    int jjaccel;        // sCKT::jjaccel(): option jjaccel and a junction in the deck (circuit.h:2119)
    int jjDCphase;      // CKTjjDCphase: the DC modes compute phase on the nodes of junctions and inductors
    int njjmod; int jj_limits[JJ_MAXMODELS];    // model has a supercurrent (ictype != 0)
    double jj_vth[JJ_MAXMODELS], jj_tsfact[JJ_MAXMODELS];
};

struct MosView {
    int n, has_level6;
    const int *d, *g, *s, *b, *dp, *sp, *type, *off, *state, *level, *cslot, *rslot;
    ParamRef prm[MP_NPARAMS];
    double *cache;      // [MC_NCACHE][n][ldt]
    double *save;       // backup copy of cache
    double *lte;        // [2][n][ldt] LTE bounds of the device's last transient load (this order, order+1)
    MosStageMap smap;   // shared-memory staging layout of k_load_mos
    int stage_doubles;  // staging area size
};
// Linear inductors (indload.cc).  Carried inside CapView: the two linear reactive elements
// travel together through the launchers.
struct IndView {
    int n;
    const int *branch, *state, *cslot, *rslot;
    ParamRef induct, ic;
    double *req, *veq, *prevFlux;   // [n][ldt]  (INDreq / INDveq / INDprevFlux instance fields)
    // couplings onto each inductor in MUT load order (CSR), and the mutual stamps
    const int *mptr, *mother; const double *mk;
    int nmut; const int *mut1, *mut2, *mut_cslot; const double *mutk;
};
struct JJView {
    int n, nmodels;
    const int *pos, *neg, *phs, *model, *state, *cslot, *rslot;
    const double *vless, *vmore, *vg, *initV, *initPhi;     // [n] shared by all trials
    ParamRef prm[JP_NPARAMS];
    double *dvdt;       // [n][ldt] JJdelVdelT (instance field, jjload.cc:438)
    int rtype[JJ_MAXMODELS], ictype[JJ_MAXMODELS], pi[JJ_MAXMODELS];
    double delv[JJ_MAXMODELS], icfactor[JJ_MAXMODELS];
};
// resistors on phase nodes: loaded every iteration (resload.cc:250-292)
struct ResPhView {
    int n;
    const int *cslot;   // [n*4] PP NN PN NP
    const int *cls;     // 0: both terminals phase / ground, 1: volt - phase, 2: phase - volt
    const int *res;     // [n] index of the resistor
    ParamRef r;         // card values of ALL resistors (per trial where B200_P_RES_R was given)
    const double *m;    // multipliers
    double resmin;
};
// Tunnel-junction model (JJ level 3).  work[(a*nterms + k)*2 + c][n][ldt], a in TF_*, c = re / im
struct TJMView {
    int n, nmodels, nterms, model0;     // model0: index of its first model in the shared jj_vmax / Consts tables
    const int *pos, *neg, *phs, *model, *state, *cslot, *rslot;
    const double *initV, *initPhi, *omegaJ, *kgap, *kgap_rejpt;     // [n]
    const double *coef;                 // [n][3][nterms][2]: p, A, B (shared by all trials)
    ParamRef prm[TP_NPARAMS];
    double *dvdt, *sc_old;              // [n][ldt], [2][n][ldt] (sin, cos of phi/2 at the last accepted point)
    double *work;
    int rtype[JJ_MAXMODELS], ictype[JJ_MAXMODELS], nt[JJ_MAXMODELS];
};
struct CapView {
    int n;
    const int *pos, *neg, *state, *cslot, *rslot;
    ParamRef capac, ic;
    double *geq, *ceq;  // [n][ldt]  (CAPgeq / CAPceq instance fields, capload.cc:150-151)
    IndView L;
    JJView J;           // Josephson junctions travel with the other two-terminal reactive elements
    TJMView T;
    ResPhView R;        // the resistors a JJ deck loads in place
};
struct SrcView {
    int n;
    const int *kind, *fn, *dcgiven, *rslot, *coef_off, *ncoef;
    ParamRef dc;
    const double *pool;
    double *dcval;      // [n][ldt] run-time SRCdcValue (srcload.cc:236-237)
};

struct VecView {
    double *X[2];       // rhs / rhsOld ping-pong, [n+1][ldt]
    double *spare;      // [n+1][ldt]
    double *sols;       // [8][n+1][ldt]
    double *oldSol;     // [n+1][ldt]
    double *states;     // [NS][numStates][ldt]
    double *oldState0;  // [numStates][ldt]
    double *contrib;    // [NC][ldt]: matrix + rhs contributions in assembly order
    double *out;        // [npoints_max][nout][ldt] -> transposed on fetch
    const int *outslot; // [n+1] row of unknown i in `out`, -1 = not saved; NULL = all, row i-1 (b200_batch_set_saves)
    const char *phase;  // [n+1] node carries the phase flag, or NULL
};
