/*
 * b200spice.h -- C ABI of the B200-native batched Newton engine for WRspice's hot path
 * (per-iteration device load + fixed-pattern sparse LU/solve + the per-trial Newton /
 * time-step driver, batched over Monte-Carlo / sweep trials that share one sparsity pattern).
 *
 * Plain C: pointers and sizes only, no torch / C++ types.  Every entry point cites the
 * reference interface (wrcad/xictools, paths relative to the reference root) it stands in for.
 *
 * Three groups:
 *   1. deck description  (what the reference builds in sCKT::setup()/temp() from parsed cards:
 *      wrspice/src/ckt/ckt.cc:2110-2307, devlib/<dev>/<dev>set.cc, <dev>temp.cc)
 *   2. batched run       (replaces sCHECKprms::loop / sDCTprms::loop_mt over
 *      TRANanalysis::tran_dcoperation: fte/check.cc:1472-1733, analysis/dctprms.cc:841-1082,
 *      analysis/tranan.cc:321-468)
 *   3. KLU-hook seam     (the spMatlabMatrix / klu_* contract:
 *      xt_base/include/sparse/spmatrix.h:815-868, wrspice/src/sparse/kluif.cc:569-628)
 *
 * There is NO CPU fallback behind any compute entry point: without a CUDA device they
 * return B200_E_NOGPU.
 */
#ifndef B200SPICE_H
#define B200SPICE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes: numerically identical to wrspice/include/errors.h:52-100 ---- */
#define B200_OK            0
#define B200_E_NOMEM       1
#define B200_E_PANIC       2
#define B200_E_BADPARM    11
#define B200_E_UNSUPP     13
#define B200_E_NODECON    18
#define B200_E_SINGULAR  101
#define B200_E_ITERLIM   103
#define B200_E_ORDER     104
#define B200_E_METHOD    105
#define B200_E_TIMESTEP  106
#define B200_E_MATHDBZ   117   /* errors.h:117-120 E_MATH* (FPEcheck emulation) */
#define B200_E_MATHINV   118
#define B200_E_MATHOVF   119
#define B200_E_NOGPU    1001   /* ours: no CUDA device / extension unusable */

/* ---- device kinds ---- */
enum { B200_DEV_RES = 0, B200_DEV_CAP = 1, B200_DEV_IND = 2, B200_DEV_MUT = 3,
       B200_DEV_SRC = 4, B200_DEV_MOS = 5, B200_DEV_JJ = 6, B200_DEV_TJM = 7, B200_DEV_NKINDS = 8 };

/* NaN in a double field == "not given" (the reference's xxxGiven flags). */

/* .options that the hot path reads: wrspice/include/circuit.h:461-631 (defaults),
 * 999-1069 (TSK* macros). */
typedef struct b200_options {
    double abstol, reltol, vntol, chgtol, trtol, gmin, gmax, pivrel, pivtol;
    double temp, tnom;          /* Kelvin */
    double xmu, dcmu, trapratio, delmin, minbreak, resmin;
    double defl, defw, defad, defas;
    int    bypass, maxord, itl1, itl2, itl4, itl_gmin, itl_src;
    int    gminsteps, srcsteps, gminfirst;
    int    method;              /* 0 = TRAPEZOIDAL, 1 = GEAR */
    int    trapcheck, spice3, oldsteplim, oldlimit /* TSKfixLimit */, forcegmin;
    int    interplev, jjaccel, nojjtp, noopiter;
    int    steptype;            /* 0 STEP_NORMAL (interpolate to user points) */
} b200_options;

/* Resistor: devlib/res/restemp.cc:186-230 (conduct = m/(r*tcfactor), preloaded). */
typedef struct b200_res { int pos, neg; double r, m, tc1, tc2, temp; } b200_res;

/* Capacitor: devlib/cap/capset.cc, captemp.cc, capload.cc:104-205. */
typedef struct b200_cap { int pos, neg; double c, m, ic; } b200_cap;

/* Inductor / mutual: devlib/ind/indset.cc, indload.cc:84-150, mutload.cc:56-108. */
typedef struct b200_ind { int pos, neg; double l, m, ic; } b200_ind;
typedef struct b200_mut { int ind1, ind2; double k; } b200_mut;

/* Independent source: devlib/src/srcset.cc:64-370, srcload.cc:67-262,
 * src/inp/inptran.cc (PULSE :1099-1414, GPULSE :1567-1827, PWL :1980-2388, SIN :2496-2624,
 * SPULSE :2708-2870, EXP :2902-3082, SFFM :3113-3260, AM :3292-3442).  A PULSE / GPULSE bit pattern
 * (inptran.cc:285-520, `bprbs6 r=-1`, `b0110 rb=2 r=3` ...) follows the arguments in `coef` after a
 * NaN: length, repeat start, then the expanded bits as 0 / 1 (shim/harvest.cc expands it with the
 * reference's own pbitAry).  GAUSS (a random source), TABLE and INTERP are refused. */
enum { B200_SRC_V = 0, B200_SRC_I = 1 };
enum { B200_FN_NONE = 0, B200_FN_PULSE = 1, B200_FN_PWL = 2, B200_FN_SIN = 3,
       B200_FN_EXP = 4, B200_FN_SFFM = 5, B200_FN_AM = 6, B200_FN_GAUSS = 7,
       B200_FN_GPULSE = 8, B200_FN_SPULSE = 9 };
typedef struct b200_src {
    int kind;                   /* B200_SRC_V / B200_SRC_I */
    int pos, neg;
    double dc;                  /* NaN if not given */
    int fn;                     /* B200_FN_* */
    int ncoef;                  /* number of doubles in coef */
    const double *coef;         /* function arguments exactly as on the card */
} b200_src;

/* MOSFET level 1/2/3/6 card: devlib/mos/mosdefs.h:415-600 (model), :115-328 (instance),
 * defaults mosset.cc:83-390, preprocessing mostemp.cc:65-389. */
typedef struct b200_mos_model {
    int type;                   /* +1 NMOS, -1 PMOS */
    int level;                  /* 1,2,3,6 */
    double vto, kp, gamma, phi, lambda, rd, rs, cbd, cbs, is, pb, cgso, cgdo, cgbo;
    double rsh, cj, mj, cjsw, mjsw, js, tox, ld, u0, fc, nsub, tpg, nss, tnom;
    double nfs, delta, uexp, ucrit, vmax, xj, neff, eta, theta, kappa;   /* level 2/3 */
    double kv, nv, kc, nc, lamda0, lamda1, sigma, gamma1;                /* level 6 (moseq.cc:1091) */
} b200_mos_model;
typedef struct b200_mos {
    int d, g, s, b;             /* external node numbers (0 = ground) */
    int model;                  /* index into b200_deck.mos_models */
    double l, w, ad, as, pd, ps, nrd, nrs, m, temp;
    double icvds, icvgs, icvbs;
    int off;
} b200_mos;

/* Josephson junction, RSJ level 1: devlib/jj/jjdefs.h, jjload.cc, jjacct.cc.
 * The MODEL carries what sJJmodel holds after the model pass of JJdev::setup
 * (jjset.cc:196-545: defaults, range clamps, BCS gap when vg is not on the card), i.e. values a
 * host reads off the set-up sJJmodel (shim/harvest.cc does).  The per-INSTANCE derivation
 * (jjset.cc:609-715: ics -> area -> criti, cap, g0, gn, gs, cr1, cr2, gshunt, vless, vmore) runs
 * on the device, per trial where ics varies.
 * Built: rtype 0..3, ictype 0 / 1 (no control current), vshunt, the phase node; `.tran ... uic`
 * starts and the phase-mode DC operating point (jjload.cc:154-250 with the phase terms it puts
 * on inductors, mutuals and resistors: indload.cc:88-96, mutload.cc:75-87, resload.cc:261-283).
 * Refused (B200_E_UNSUPP): lser / lsh parasitic inductances, rtype 4, ictype 2..4, control. */
typedef struct b200_jj_model {
    int rtype, ictype, pi;          /* JJrtype, JJictype, JJpi */
    double criti;                   /* JJcriti: critical current of the reference junction */
    double vg;                      /* JJvgnom: gap voltage at tnom */
    double delv, r0, rn, cap, icfactor, cmu, gmu;
    double vshunt;                  /* JJvShunt: model default of the instance vshunt (0 = none) */
    double vdpbak, tsaccl, tsfact;  /* time-step control of JJdev::accept (jjacct.cc:57,138) */
} b200_jj_model;
typedef struct b200_jj {
    int pos, neg, phase_node;       /* phase_node 0 = absent */
    int model;
    double area, ics;               /* NaN = not given (ics wins, jjset.cc:613-631) */
    double vshunt;                  /* NaN = the model's */
    double vg, tcf;                 /* JJvg, JJtcf after setup (NaN = model vg, 1.0: temp_k == tnom) */
    double icv, icphi;              /* ic=vj,phi for the UIC start (NaN = 0) */
} b200_jj;

/* Josephson junction, microscopic tunnel-junction model (mmjco / TJM, JJ level 3):
 * devlib/tjm/tjmdefs.h, tjmload.cc:65-553, tjmacct.cc:43-169.  The current is the Werthamer
 * expression with the pair and quasiparticle kernels fitted by nterms complex exponentials
 * (coefficients p, A, B); every Newton iteration advances 2*nterms complex recurrences.
 * The coefficient set comes from the reference's set-up instance (TJMcoeffSet, tjmcoeffs.cc:
 * built-in tables `tjm1` / `tjm2`, or fitted by the external `mmjco` program), AFTER
 * sTJMinstance::tjm_init1 has renormalised it (tjmset.cc:869-918) -- none of it depends on the
 * junction's area, so it is shared by all trials.  Per trial (ics): criti, cap, gqp, g0,
 * gshunt, gcrit (tjmset.cc:600-660, tjmload.cc:467-470).
 * Built like the RSJ model (`uic` and phase-mode DC starts, tjmload.cc:78-120); no lser / lsh. */
typedef struct b200_tjm_model {
    int rtype, ictype, nterms;
    double criti, cap, cmu, gmu, r0, rsint, vshunt, tsaccl, tsfact, vgnom;
} b200_tjm_model;
typedef struct b200_tjm {
    int pos, neg, phase_node;       /* phase_node 0 = absent */
    int model;
    double area, ics, vshunt;       /* NaN = not given */
    double tcf;                     /* TJMicTempFactor */
    double vdpbak, omegaJ, kgap, kgap_rejpt, alphaN;    /* set-up instance values */
    const double *coef;             /* 6*nterms doubles: p[n], A[n], B[n] as (re, im) pairs */
    double icv, icphi;              /* ic=vj,phi of the UIC start (NaN = from the node ICs / 0) */
} b200_tjm;

/* One deck = one topology (shared by all trials).  Node numbers: 0 is ground, 1..nnodes-1 are
 * the external nodes in the order the reference's parser meets them (src/inp/inpdeck.cc pass2);
 * internal nodes and branch equations are appended by b200_batch_create in the reference's
 * setup order.  `order` lists every device in the reference's LOAD order
 * (ckt.cc:991-1005 mutuals first, then :1146-1161 model-type list -> models -> instances). */
typedef struct b200_deck {
    int nnodes;
    b200_options opt;
    /* .tran tstep tstop [tstart [tmax]] [uic]: analysis/tranan.cc:95-217 */
    double tstep, tstop, tstart, tmax;
    int uic;
    /* .ic / .nodeset: ckt.cc:908-925, 1177-1204 */
    int nic;      const int *ic_node;  const double *ic_val;
    int nnodeset; const int *ns_node;  const double *ns_val;
    /* device tables */
    int nres;  const b200_res *res;
    int ncap;  const b200_cap *cap;
    int nind;  const b200_ind *ind;
    int nmut;  const b200_mut *mut;
    int nsrc;  const b200_src *src;
    int nmosmod; const b200_mos_model *mos_models;
    int nmos;  const b200_mos *mos;
    int njjmod; const b200_jj_model *jj_models;
    int njj;   const b200_jj *jj;
    /* load order: order_kind[i] in B200_DEV_*, order_index[i] into that kind's table */
    int norder; const int *order_kind; const int *order_index;
    int ntjmmod; const b200_tjm_model *tjm_models;
    int ntjm;  const b200_tjm *tjm;
} b200_deck;

void b200_options_default(b200_options *o);        /* circuit.h:461-631 */

/* ---- per-trial parameter variation (what the reference's .exec block / .param text does
 * before each trial's re-parse: fte/check.cc:1649-1700, simulate.cc:828-837) ---- */
enum {
    B200_P_MOS_VTO = 1, B200_P_MOS_KP, B200_P_MOS_L, B200_P_MOS_W, B200_P_MOS_GAMMA,
    B200_P_MOS_LAMBDA, B200_P_MOS_PHI,
    B200_P_CAP_C = 32, B200_P_RES_R = 48, B200_P_IND_L = 56, B200_P_SRC_DC = 64,
    B200_P_JJ_ICS = 80,            /* instance ics: let @b1[ics] = ... of an .exec block */
    B200_P_TJM_ICS = 88
};

typedef struct b200_batch b200_batch;

/* Create a batch of `ntrials` trials of one deck on CUDA device `device`.
 * Does what sCKT::setup() + spSaveForInitialization() do once per trial in the reference
 * (ckt.cc:627-674, spspice.cc:79-100): allocates nodes/branches/states, builds the CSC
 * pattern from every device's TSTALLOC list, computes the ordering + symbolic LU.
 * Errors: B200_E_BADPARM (null deck, ntrials <= 0, opt.method not TRAPEZOIDAL 0 / GEAR 1),
 * B200_E_NOGPU (no CUDA device: there is no CPU fallback), B200_E_UNSUPP (a device model or
 * option that is not built; the caller keeps its CPU loop), B200_E_NOMEM. */
int  b200_batch_create(const b200_deck *deck, int ntrials, int device, b200_batch **out);
void b200_batch_destroy(b200_batch *b);

/* values[ntrials] (host) replaces instance `index`'s parameter `param` for every trial.
 * B200_P_CAP_C is the card value (the instance multiplier m still applies, captemp.cc:98);
 * B200_P_SRC_DC makes the source's DC value "given" (it then wins over the source function in the
 * DC modes, srcload.cc:215-237); B200_P_RES_R is the card value of a resistor (conduct = m/r with
 * the resmin clamp, restemp.cc:200-209): every trial re-sums its preloaded conductances on the
 * device in the reference's accumulation order, resistors on phase nodes stamp their own value.
 * For B200_P_MOS_VTO/KP/GAMMA/LAMBDA/PHI the value is the instance's private model card value
 * (the reference MC decks give each varied device its own .model). */
int  b200_batch_set_param(b200_batch *b, int param, int index, const double *values);

/* The same for many rows at once: values[nrows][ntrials] (host, row r = parameter param[r] of
 * instance index[r]) goes to the device in ONE transfer on the batch's stream -- a DMA when the
 * buffer is page-locked -- and a device kernel scatters the rows into the trial-minor tables.
 * The buffer must stay unchanged until the next b200_batch_run_* call returns.  Replaces the
 * rows of an earlier call; single rows (b200_batch_set_param) override it per instance. */
int  b200_batch_set_params(b200_batch *b, int nrows, const int *param, const int *index,
                           const double *values);

/* Runs DCOP (or UIC start) + transient for every trial to completion on the GPU:
 * TRANanalysis::tran_dcoperation (tranan.cc:321-468) with sCKT::op (ckt.cc:1859-2011),
 * NIiter (niiter.cc:160-555), sTRANint::step/accept/interpolate (tranan.cc:483-1116). */
int  b200_batch_run_tran(b200_batch *b);

/* DC operating point of every trial: DCOanalysis::anFunc (analysis/dcoan.cc:100-121) = ic() +
 * sCKT::op(MODEDCOP|MODEINITJCT, MODEDCOP|MODEINITFLOAT, itl1) with the gmin / source stepping
 * fall-backs (ckt.cc:1859-2011).  One output point per trial (time 0 = CKTrhsOld after op());
 * counters as for run_tran.  With per-trial parameter arrays (b200_batch_set_param, including
 * B200_P_SRC_DC) this is the batched form of a chained-DC / parameter sweep
 * (sDCTprms::loop_mt, analysis/dctprms.cc:924-1011): every sweep point is a trial. */
int  b200_batch_run_op(b200_batch *b);

/* Results (host buffers).  npoints = number of user time points (count of OP.appendData calls
 * per trial in the reference, tranan.cc:963-1116 -> ckt.cc:680-708). */
int  b200_batch_num_unknowns(const b200_batch *b);   /* matrix size n (nodes + branches) */
int  b200_batch_num_points(const b200_batch *b);
/* names: '\0'-separated, reference-style ("vdd#branch", "m1#internal#drain") */
int  b200_batch_unknown_names(const b200_batch *b, char *buf, int buflen);
/* counters[ntrials][B200_NCOUNTERS]: sSTATS fields, circuit.h:1093-1163 */
enum { B200_C_STATUS = 0, B200_C_ACCEPTED, B200_C_REJECTED, B200_C_NUMITER, B200_C_TIMEPTS,
       B200_C_TRANITER, B200_C_ITERCUT, B200_C_TRAPCUT, B200_C_NPOINTS, B200_NCOUNTERS };
int  b200_batch_get_counters(const b200_batch *b, int *counters);
/* waves[ntrials][npoints][n] : unknown vector at each user time point; times[npoints] */
int  b200_batch_get_waves(const b200_batch *b, double *times, double *waves);
/* Same, restricted to a subset of unknowns (what `.save`/`.print` would keep). */
int  b200_batch_get_waves_sel(const b200_batch *b, int nsel, const int *sel, double *waves);

/* The same into DEVICE buffers of the caller, dwaves[ntrials][npoints][nsel] and (optional)
 * dcounters[ntrials][B200_NCOUNTERS]: what the multi-GPU dispatcher hands to the final gather
 * (NCCL over NVLink) without a host round trip. */
int  b200_batch_get_results_dev(const b200_batch *b, int nsel, const int *sel, double *dwaves,
                                int *dcounters);

/* Kernel statistics of the last run (for bench.py): number of kernel launches, Newton ticks,
 * device-load instance evaluations, seconds spent per kernel class (CUDA events). */
typedef struct b200_run_stats {
    long long launches, ticks, load_instances;
    double ms_total, ms_load, ms_solve, ms_control;
    long long solve_launches, load_launches;
    double algbytes_load_per_launch, algbytes_solve_per_launch;
} b200_run_stats;
/* `.save`: keep only these unknowns (0-based indices as in b200_batch_unknown_names) at the output
 * points -- the output buffer is [points][saved][trials], so a long transient of many trials fits
 * (the reference keeps saved vectors only as well: the `.save` lines of a deck).  Before the first
 * run.  b200_batch_get_waves then returns the saved columns in list order; get_waves_sel /
 * get_results_dev still take unknown indices (B200_E_BADPARM for one that is not saved). */
int  b200_batch_set_saves(b200_batch *b, int nsave, const int *unknown);
int  b200_batch_num_saved(const b200_batch *b);
int  b200_batch_get_stats(const b200_batch *b, b200_run_stats *st);

/* ---- Monte-Carlo trial generation (host code of the dispatcher) ----
 * The deviates a WRspice Monte-Carlo run draws: sRnd::random / sRnd::gauss
 * (xt_base/include/miscutil/random.h:76-79, xt_base/miscutil/random.cc:51-72) over the BSD
 * random(3) TYPE_3 stream of xt_base/miscutil/randval.cc:196-258, seeded like sRnd::seed(seed).
 * out[n] receives the first n values of the stream.  Replaces the per-trial gauss() calls of a
 * deck's .exec block in the batched dispatcher (fte/check.cc:1649-1733). */
int  b200_mc_uniform(int seed, long long n, double *out);
int  b200_mc_gauss(int seed, long long n, double *out);

/* ---- debug / parity hooks used by tests (device state -> host) ---- */
/* Host-side analysis products: CSC pattern of A, and the LU ordering the batch uses. */
int  b200_batch_pattern(const b200_batch *b, int *n, int *nnz, const int **Ap, const int **Ai);
int  b200_batch_ordering(const b200_batch *b, const int **rowperm, const int **colperm);
/* A deck whose DC and transient matrices need different pivots (inductors: the branch diagonal
 * is 0 in DC and -L*ag0 in a transient) keeps two sequences, as the reference re-orders at the
 * first transient step (niiter.cc:501-509): this is the one of the transient modes (equal to
 * the one above for a batch with a single sequence). */
int  b200_batch_ordering_tran(const b200_batch *b, const int **rowperm, const int **colperm);
/* Single load+assemble at a forced state: the per-trial matrix values Ax[nnz][ntrials] and
 * RHS[n+1][ntrials] after sCKT::load + loadGmin, for stamp-for-stamp comparison. */
int  b200_batch_debug_load(b200_batch *b, int mode, double time, double delta, int order,
                           const double *rhsold /* [n+1][ntrials] or NULL */,
                           double *Ax, double *rhs);

/* Host-only (no CUDA call): sizes of the once-per-batch analysis products of a deck -- the
 * pattern (spbuild.cc:922-958), the LU under the static pivot order and the level-scheduled
 * task program of the refactor + forward + backward solve.  out[min(nout, B200_AS_NSTATS)]. */
enum { B200_AS_N = 0, B200_AS_NNZ, B200_AS_NNZLU, B200_AS_NC, B200_AS_NSTATES,
       B200_AS_IMG /* doubles per trial image */, B200_AS_SCRATCH /* partial-sum slots */,
       B200_AS_TASKS, B200_AS_OPS, B200_AS_LEVELS, B200_AS_ROUNDS, B200_AS_STEPS,
       B200_AS_PROG_BYTES, B200_AS_LONGEST_RUN,
       B200_AS_LU_SPLIT, B200_AS_LU_CHUNK /* the split rule of long sums (b200_internal.h) */,
       B200_AS_WIDE /* 1: image beyond the shared-memory window, wide program (k_lu_wide) */, B200_AS_NSTATS };
int  b200_deck_analysis_stats(const b200_deck *deck, long long *out, int nout);
/* Host-only: the CSC pattern and the static pivot order (position -> original row / column,
 * 0-based) into caller buffers Ap[n+1], Ai[nnz], rowperm[n], colperm[n] (any may be NULL). */
int  b200_deck_analysis_arrays(const b200_deck *deck, int *Ap, int *Ai, int *rowperm, int *colperm);
/* Host-only: the pivot order of the DC modes (which = 0) or the transient modes (which = 1);
 * *nprograms = 2 when the two differ (the batch then keeps two solve programs). */
int  b200_deck_analysis_order(const b200_deck *deck, int which, int *rowperm, int *colperm,
                              int *nprograms);

/* Test hook for the host logic (no CUDA call): orders a CSC matrix the way b200_klu_analyze
 * does, compiles the assembly runs and the task program, and executes them on ONE value set in
 * the order the kernels do, so the schedule can be checked bit for bit against the oracle's
 * static LU on a machine without a GPU.  x[n], rowperm[n], colperm[n] (may be NULL),
 * stats[8] = {nnzLU, image doubles, scratch slots, tasks, products, levels, rounds, steps}.
 * Not a solver and not on any product path. */
int  b200_debug_program_check(int n, const int *Ap, const int *Ai, const double *Ax, const double *b,
                              double *x, int *rowperm, int *colperm, long long *stats, int nstats);

/* Test hook: the LTE divided differences divide by per-trial constants through a correctly
 * rounded reciprocal + FMA refinement; this compares that sequence with IEEE division on n
 * pseudo-random operand pairs with binary exponents in [emin, emax]. */
int  b200_debug_divcheck(long long n, unsigned long long seed, int emin, int emax,
                         long long *mismatches);

/* ---- KLU-hook seam (batched): spMatlabMatrix::factor/refactor/solve semantics ----
 * xt_base/include/sparse/spmatrix.h:853-857; kluif.cc:569-628.  Pattern is CSC, 0-based,
 * row indices ascending per column (spbuild.cc:922-958). */
typedef struct b200_klu b200_klu;
/* klu_analyze + the pivot choice of the first klu_factor on trial 0's values Ax0 (host, nnz
 * doubles); the pivot sequence is then frozen, which is exactly klu_refactor's contract.
 * Order: the engine's static Markowitz order (constant +-1 entries, conductance diagonals) if
 * every pivot of it passes |pivot| >= pivabs and |pivot| >= pivrel * max|column| on Ax0
 * (spfactor.cc RelThreshold/AbsThreshold, KLU Common.tol), else threshold Markowitz on Ax0.
 * pivrel outside (0,1] -> 1e-3, pivabs < 0 -> 1e-13 (the task defaults).
 * Errors: B200_E_BADPARM, B200_E_NOGPU, B200_E_SINGULAR (no acceptable pivot sequence for
 * Ax0), B200_E_UNSUPP (LU image beyond the shared-memory program's reach), B200_E_NOMEM. */
int  b200_klu_analyze(int n, const int *Ap, const int *Ai, const double *Ax0,
                      double pivrel, double pivabs, int device, b200_klu **out);
void b200_klu_free(b200_klu *k);
/* Batched klu_refactor + klu_solve, HOST buffers: Ax[ntrials][nnz], B[ntrials][n] overwritten
 * with the solutions; status[ntrials] = B200_OK / B200_E_SINGULAR (a zero or non-finite pivot under
 * the frozen sequence: klu_refactor's test) / B200_E_MATHOVF (non-finite solution).  (e2e path) */
int  b200_klu_refactor_solve_host(b200_klu *k, int ntrials, const double *Ax, double *B,
                                  int *status);
/* Same with DEVICE buffers in the engine's trial-minor layout Ax[nnz][ldt], B[n][ldt]. */
int  b200_klu_refactor_solve_dev(b200_klu *k, int ntrials, int ldt, const double *dAx,
                                 double *dB, int *dstatus, void *cuda_stream);
int  b200_klu_lu_nnz(const b200_klu *k);
/* KLUmatrix::rowmap / colmap (kluif.cc:666-681): position -> original row / column, 0-based. */
int  b200_klu_get_ordering(const b200_klu *k, int *rowperm, int *colperm);

const char *b200_version(void);
int  b200_cuda_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
