/*
 * wrs_oracle.c -- TEST INFRASTRUCTURE ONLY (see wrs_oracle.h).
 *
 * Scalar CPU restatement of the WRspice Newton/transient hot path.  Written from the
 * reference's behaviour, function by function; every block cites the file:line it follows
 * (relative to /root/reference/wrspice).  Expression order is kept identical to the
 * reference wherever floating-point results depend on it (compile with -ffp-contract=off).
 */
#define _GNU_SOURCE
#include <fenv.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "wrs_oracle.h"

/* circuit.h:2158-2192 */
#define MODETRAN        0x1
#define MODEDC          0x70
#define MODEDCOP        0x10
#define MODETRANOP      0x20
#define MODEDCTRANCURVE 0x40
#define INITF           0x7f00
#define MODEINITFLOAT   0x100
#define MODEINITJCT     0x200
#define MODEINITFIX     0x400
#define MODEINITSMSIG   0x800
#define MODEINITTRAN    0x1000
#define MODEINITPRED    0x2000
#define MODEUIC         0x10000
#define NISHOULDREORDER 0x1

/* circuit.h:111-125 */
#define REFTEMP     298.15
#define CHARGE      1.602176634e-19
#define CONSTboltz  1.380649e-23
#define CONSTKoverQ (CONSTboltz/CHARGE)
#define CONSTroot2  M_SQRT2
#define EPSSIL      (11.7 * 8.854214871e-12)
#define CM3PM3 1e6
#define CM2PM2 1e4
#define M2PCM2 1e-4

#define OK 0
#define E_SINGULAR B200_E_SINGULAR
#define E_ITERLIM  B200_E_ITERLIM
#define E_TIMESTEP B200_E_TIMESTEP
#define E_MATHDBZ  117
#define E_MATHINV  118
#define E_MATHOVF  119
#define E_MATHUNF  120
#define E_NOCHANGE 16

#define SPMAX(a,b) ((a) > (b) ? (a) : (b))
#define SPMIN(a,b) ((a) < (b) ? (a) : (b))
#define GIVEN(x) (!isnan(x))

enum { SP_VOLTAGE = 0, SP_CURRENT = 1 };
#define PHI0      2.067833848e-15            /* wrsCONSTphi0, circuit.h:123 */
#define PHI0_2PI  (PHI0/(2*M_PI))

/* ------------------------------------------------------------------------- */
/* device records                                                             */
/* ------------------------------------------------------------------------- */

/* MOS states (mosdefs.h:330-375, minus the 18 write-only "ask" mirrors) */
enum { MS_vbd, MS_vbs, MS_vgs, MS_vds, MS_capgs, MS_qgs, MS_cqgs, MS_capgd, MS_qgd, MS_cqgd,
       MS_capgb, MS_qgb, MS_cqgb, MS_qbd, MS_cqbd, MS_qbs, MS_cqbs, MS_NSTATES };

/* matrix stamp slots in the order of mosset.cc:53-77 get_node_ptr */
enum { M_Dd, M_Gg, M_Ss, M_Bb, M_DPdp, M_SPsp, M_Ddp, M_Gb, M_Gdp, M_Gsp, M_Ssp, M_Bg, M_Bdp,
       M_Bsp, M_DPsp, M_DPd, M_Bd_unused, M_DPg, M_DPb, M_SPg, M_SPs, M_SPb, M_SPdp, M_NPTR };

typedef struct {
    /* card */
    b200_mos_model mod;      /* private copy so per-trial vto/kp overrides are local */
    b200_mos_model card;     /* the card as given (NaN = not given): the xxxGiven tests of mostemp */
    b200_mos in;
    /* model-derived (mostemp.cc:65-139) */
    double oxideCapFactor, surfaceMobility, kp_eff, gamma_eff, phi_eff, vt0_eff;
    double xd, alpha, narrowFactor, nsub;      /* level 2/3 (mostemp.cc:113-143) */
    /* nodes */
    int d, g, s, b, dp, sp;
    int ptr[M_NPTR];
    int st;
    /* instance-derived (mosset.cc:283-373, mostemp.cc:141-300) */
    double mult, temp, l, w, drainArea, sourceArea, drainPerimeter, sourcePerimeter;
    double drainSquares, sourceSquares, icVBS, icVDS, icVGS;
    double drainConductance, sourceConductance, effectiveLength;
    double tTransconductance, tSurfMob, tPhi, tVbi, tVto, tSatCur, tSatCurDens;
    double tCbd, tCbs, tCj, tCjsw, tBulkPot, tDepCap;
    double sourceVcrit, drainVcrit, tDrainSatCur, tSourceSatCur;
    double Cbd, Cbdsw, Cbs, Cbssw, f2d, f3d, f4d, f2s, f3s, f4s;
    double gateSourceOverlapCap, gateDrainOverlapCap, gateBulkOverlapCap, beta, oxideCap;
    /* cached op point (mosdefs.h:125,160-173) */
    int mode;
    double von, vdsat, cd, cbs, cbd, gm, gds, gmbs, gbd, gbs, capbd, capbs;
} Mos;

typedef struct { int pos, neg; int pPP, pNN, pPN, pNP; int st; int icGiven; double capac, ic, geq, ceq, m; } Cap;
typedef struct { int pos, neg; int pPP, pNN, pPN, pNP; double conduct, m; int usePreload; } Res;
/* sINDinstancePOD (inddefs.h): linear inductor, no series resistance / poly / expression */
typedef struct { int pos, neg, br; int pBB, pPB, pNB, pBN, pBP; int st; int icGiven;
    double induct, ic, req, veq, prevFlux; } Ind;
/* sMUTinstancePOD (inddefs.h) */
typedef struct { int ind1, ind2; int p12, p21; double coupling, factor; } Mut;
typedef struct {
    int kind, pos, neg, branch;
    int pPB, pNB, pBN, pBP;
    int dcGiven; double dc;
    int fn, ncoef; double *coef;
    /* run-time tran function parameters (inptran.cc) */
    double parms[8], cache[2];
    double *pwl_cache; int pwl_n, pwl_index;
    int enable_tran;
    int npat, prst, nargs, pat_on; const double *pat;   /* PULSE / GPULSE bit pattern */
} Src;

/* sJJinstancePOD + the model values it reads (jjdefs.h).  No lser / lsh / control. */
enum { JS_voltage, JS_dvdt, JS_phase, JS_conI, JS_phsInt, JS_crti, JS_qpi, JS_NSTATES = 11 };  /* jjdefs.h:237-250 */
typedef struct {
    b200_jj_model mod; b200_jj in;
    int pos, neg, phs;
    int pPP, pNN, pNP, pPN, pHH;             /* jjset.cc:136-139, :195 TSTALLOC order */
    int pRPP, pRPN, pRNP, pRNN;              /* shunt resistor stamps, jjset.cc:178-181 */
    int st;
    double area, criti, cap, g0, gn, gs, g1, g2, cr1, cr2, gshunt, vless, vmore, vg, tcf, vshunt;
    double initVoltage, initPhase, delVdelT;
    int phsN, phsF; double phsT;
} JJ;

/* sTJMinstancePOD (tjmdefs.h:100-245): JJ level 3, the mmjco tunnel-junction model */
enum { TS_voltage, TS_dvdt, TS_phase, TS_phsInt, TS_crti, TS_qpi, TS_NSTATES = 10 };    /* tjmdefs.h:245-257 */
#define TJM_MAXTERMS 20
typedef struct { double re, im; } Cx;
typedef struct {
    b200_tjm_model mod; b200_tjm in;
    int pos, neg, phs;
    int pPP, pNN, pNP, pPN, pHH, pRPP, pRPN, pRNP, pRNN;
    int st, n;
    double area, criti, cap, gqp, g0, gshunt, vshunt, gcrit;
    double initVoltage, initPhase, delVdelT;
    double sinphi_2_old, cosphi_2_old, cqp, cp;
    Cx p[TJM_MAXTERMS], A[TJM_MAXTERMS], B[TJM_MAXTERMS];
    Cx Fc[TJM_MAXTERMS], Fs[TJM_MAXTERMS], Fcprev[TJM_MAXTERMS], Fsprev[TJM_MAXTERMS];
    Cx alpha0[TJM_MAXTERMS], alpha1[TJM_MAXTERMS], exp_z[TJM_MAXTERMS];
} TJM;

typedef struct { int kind, index; } DevRef;

/* saved op-point for sCKT::backup (ckt.cc:1323-1362, mosdefs.h:386-402) */
typedef struct { int mode; double v[12]; } MosSave;

struct wo_ckt {
    b200_options opt;
    double tstep, tstop, tstart, tmax; int uic;
    int nnodes_ext;          /* external nodes incl. ground */
    int n;                   /* unknowns (matrix size); node numbers 1..n */
    int *ntype;              /* [n+1] */
    double *ic_val; char *ic_given; double *ns_val; char *ns_given; int hadNodeset;
    char **names;

    int nmos, ncap, nres, nsrc, nind, nmut, njj, ntjm;
    Mos *mos; Cap *cap; Res *res; Src *src; Ind *ind; Mut *mut; JJ *jj; TJM *tjm;
    double initDelta;        /* CKTinitDelta (circuit.h:1581), set by JJdev::getic */
    char *phase;             /* [n+1] node carries the phase flag (jjset.cc:656-681, indset.cc:104-120), or NULL */
    int nord; DevRef *ord;

    /* matrix: CSC, rows ascending per column (spbuild.cc:922-958) */
    int nnz; int *Ap, *Ai; double *Ax, *Ainit;
    int *diag;               /* [n+1] index of (i,i) in Ax or -1 */
    /* coordinate collection during setup */
    int ncoo, capcoo; int *coo_r, *coo_c;

    /* LU */
    int have_order; int *rowperm, *colperm;   /* position -> original (0-based) */
    int lu_ready;
    /* the engine keeps a second sequence for the transient modes when its DC one does not fit */
    int have_order2; int *rowperm2, *colperm2; int lu_ready2; void *slu2;
    int *Lp, *Li; double *Lx;   /* strictly lower, by column, in permuted positions */
    int *Up, *Ui; double *Ux;   /* strictly upper, by column */
    double *Udiag;
    int *cpat_p, *cpat_i;       /* per permuted column: positions < j that are nonzero in U(:,j) */
    int *posA;                  /* for own-pivot mode */
    double *work;
    double *dense;              /* own-pivot dense LU, n*n */
    int *dpiv;

    /* vectors (slot 0 = ground) */
    double *rhs, *rhsOld, *rhsSpare, *sols[8];
    int numStates; double *states[8];
    double *oldSol, *oldState0; MosSave *mosSave; int haveSave;

    /* sCKTPOD scalars (circuit.h:1562-1662) */
    double time, delta, deltaOld[7], saveDelta, ag[7], pred[4];
    double srcFact, diagGmin, devMaxDelta, vt;
    int order, mode, niState, noncon, brk, trapCheck, trapBad;
    double breaks[2];
    /* breakpoint table (breakpt.cc) */
    int numBreaks; double *brks; int numLat; double *lat_offs, *lat_per;
    /* sTRANint (trandefs.h:93-151) */
    double t_stop, t_step, t_start, t_check, t_delmax, t_delmin, t_fixed_step;
    int t_firsttime, t_dumpit, t_ordcnt, t_nocut, t_spice3, t_polydegree;
    double maxStep, minBreak;
    /* stats */
    int accepted, rejected, numIter, timePts, tranIter, iterCut, trapCut;
    /* outputs */
    int npts, cappts; double *pt_t, *pt_v;
    int nacc, capacc; double *acc_t, *acc_v; int *acc_it;
    int lastIters;
    int endit;
};

/* ------------------------------------------------------------------------- */
/* helpers                                                                    */
/* ------------------------------------------------------------------------- */
static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s); if (!p) abort(); return p; }

void b200_options_default_oracle(b200_options *o)
{
    /* circuit.h:461-631 */
    memset(o, 0, sizeof(*o));
    o->abstol = 1e-12; o->reltol = 1e-3; o->vntol = 1e-6; o->chgtol = 1e-14; o->trtol = 7.0;
    o->gmin = 1e-12; o->gmax = 1e6; o->pivrel = 1e-3; o->pivtol = 1e-13;
    o->temp = REFTEMP; o->tnom = REFTEMP;
    o->xmu = 0.5; o->dcmu = 0.5; o->trapratio = 10.0; o->delmin = 0.0; o->minbreak = 0.0;
    o->resmin = 1e-3;
    o->bypass = 1; o->maxord = 2; o->itl1 = 400; o->itl2 = 100; o->itl4 = 20;
    o->itl_gmin = 20; o->itl_src = 20; o->gminsteps = 0; o->srcsteps = 0; o->gminfirst = 0;
    o->method = 0; o->trapcheck = 0; o->spice3 = 0; o->oldsteplim = 0; o->oldlimit = 0;
    o->forcegmin = 0; o->interplev = 1; o->jjaccel = 0; o->nojjtp = 0; o->noopiter = 0;
    o->steptype = 0;
}

/* add a structural entry; returns coo index (resolved to Ax index after sorting) */
static int coo_add(wo_ckt *c, int r, int col)
{
    if (r == 0 || col == 0)
        return -1;
    for (int i = 0; i < c->ncoo; i++)
        if (c->coo_r[i] == r && c->coo_c[i] == col)
            return i;
    if (c->ncoo == c->capcoo) {
        c->capcoo = c->capcoo ? 2*c->capcoo : 256;
        c->coo_r = realloc(c->coo_r, c->capcoo*sizeof(int));
        c->coo_c = realloc(c->coo_c, c->capcoo*sizeof(int));
    }
    c->coo_r[c->ncoo] = r; c->coo_c[c->ncoo] = col;
    return c->ncoo++;
}

static int new_node(wo_ckt *c, int type, const char *name)
{
    c->n++;
    c->ntype = realloc(c->ntype, (c->n+1)*sizeof(int));
    c->names = realloc(c->names, (c->n+1)*sizeof(char*));
    c->ntype[c->n] = type;
    c->names[c->n] = strdup(name);
    return c->n;
}

static inline void ldadd(wo_ckt *c, int p, double v) { if (p >= 0) c->Ax[p] += v; }
static inline void rhsadd(wo_ckt *c, int node, double v) { c->rhs[node] += v; }

/* ------------------------------------------------------------------------- */
/* breakpoints: ckt/breakpt.cc:58-221                                         */
/* ------------------------------------------------------------------------- */
static double nextbreak(wo_ckt *c, double t, double minbrk)
{
    double dt0 = 0.0;
    int i;
    for (i = 0; i < c->numLat; i++) {
        double dt1 = c->lat_offs[i];
        if (t > c->lat_offs[i])
            dt1 += c->lat_per[i]*ceil((t - c->lat_offs[i])/c->lat_per[i]);
        if (fabs(dt1 - t) < minbrk)
            dt1 += c->lat_per[i];
        if (dt0 == 0.0 || dt1 < dt0)
            dt0 = dt1;
    }
    for (i = 0; i < c->numBreaks; i++)
        if (c->brks[i] > t)
            break;
    while (i < c->numBreaks && fabs(t - c->brks[i]) < minbrk)
        i++;
    if (i == c->numBreaks)
        return dt0;
    if (dt0 > 0.0 && dt0 < c->brks[i])
        return c->brks[i] < dt0 ? c->brks[i] : dt0;
    return c->brks[i];
}

static void break_clr(wo_ckt *c)
{
    /* breakpt.cc:57-72 clear_break (the table trim does not change nextbreak results) */
    c->breaks[0] = nextbreak(c, c->breaks[0], c->minBreak);
    c->breaks[1] = nextbreak(c, c->breaks[0], c->minBreak);
}

static void break_set(wo_ckt *c, double time)
{
    /* breakpt.cc:78-160 set_break */
    double minbrk = c->minBreak;
    for (int i = 0; i < c->numLat; i++) {
        double dt0, dt1;
        if (time < c->lat_offs[i]) {
            dt0 = c->lat_offs[i] + c->lat_per[i];
            dt1 = c->lat_offs[i];
        }
        else {
            dt0 = c->lat_per[i]*ceil((time - c->lat_offs[i])/c->lat_per[i]) + c->lat_offs[i];
            dt1 = c->lat_per[i]*floor((time - c->lat_offs[i])/c->lat_per[i]) + c->lat_offs[i];
        }
        if (fabs(time - dt0) < minbrk || fabs(time - dt1) < minbrk) {
            if (time < c->breaks[0])
                c->breaks[0] = nextbreak(c, c->time, minbrk);
            c->breaks[1] = nextbreak(c, c->breaks[0], minbrk);
            return;
        }
    }
    if (!c->brks) {
        c->brks = xcalloc(1, sizeof(double));
        c->brks[0] = time;
        c->numBreaks = 1;
        return;
    }
    int done = 0;
    for (int i = 0; i < c->numBreaks && !done; i++) {
        if (c->brks[i] > time) {
            if ((c->brks[i] - time) <= minbrk)
                c->brks[i] = time;
            else if (time - c->brks[i-1] <= minbrk)
                ;
            else {
                c->brks = realloc(c->brks, (c->numBreaks+1)*sizeof(double));
                memmove(c->brks + i + 1, c->brks + i, (c->numBreaks - i)*sizeof(double));
                c->brks[i] = time;
                c->numBreaks++;
            }
            done = 1;
        }
    }
    if (!done) {
        if (time - c->brks[c->numBreaks-1] <= minbrk)
            ;
        else {
            c->brks = realloc(c->brks, (c->numBreaks+1)*sizeof(double));
            c->brks[c->numBreaks++] = time;
        }
    }
    if (time < c->breaks[0])
        c->breaks[0] = nextbreak(c, c->time, minbrk);
    c->breaks[1] = nextbreak(c, c->breaks[0], minbrk);
}

static void break_set_lattice(wo_ckt *c, double offs, double per)
{
    for (int i = 0; i < c->numLat; i++)
        if (offs == c->lat_offs[i] && per == c->lat_per[i])
            return;
    c->lat_offs = realloc(c->lat_offs, (c->numLat+1)*sizeof(double));
    c->lat_per = realloc(c->lat_per, (c->numLat+1)*sizeof(double));
    c->lat_offs[c->numLat] = offs;
    c->lat_per[c->numLat] = per;
    c->numLat++;
}

/* ------------------------------------------------------------------------- */
/* source functions: src/inp/inptran.cc                                       */
/* ------------------------------------------------------------------------- */
static void src_setup_func(wo_ckt *c, Src *s, double step, double finaltime)
{
    s->enable_tran = (step > 0.0 && finaltime > 0.0);
    if (s->fn == B200_FN_PULSE) {
        /* inptran.cc:1203-1318 */
        double *p = s->parms; const double *co = s->coef; int nc = s->ncoef;
        /* a bit pattern follows a NaN: length, repeat start, bits (include/b200spice.h) */
        s->npat = 0; s->pat = 0;
        for (int k = 0; k < s->ncoef; k++)
            if (isnan(co[k])) { nc = k; s->npat = (int)co[k+1]; s->prst = (int)co[k+2]; s->pat = co + k + 3; break; }
        s->nargs = nc;
        memset(p, 0, 8*sizeof(double));
        p[0] = co[0];
        p[1] = nc >= 2 ? co[1] : co[0];
        p[2] = nc >= 3 ? co[2] : 0.0;
        p[3] = nc >= 4 ? co[3] : 0.0;
        p[4] = nc >= 5 ? co[4] : 0.0;
        p[5] = nc >= 6 ? co[5] : 0.0;
        p[6] = nc >= 7 ? co[6] : 0.0;
        s->cache[0] = p[3] > 0.0 ? (p[1] - p[0])/p[3] : 0.0;
        s->cache[1] = p[4] > 0.0 ? (p[1] - p[0])/p[4] : 0.0;
        if (s->enable_tran) {
            if (p[3] <= 0.0) p[3] = step;
            if (p[4] > 0.0 && p[5] <= 0.0)
                p[5] = 0.0;
            else {
                if (p[4] <= 0.0) p[4] = step;
                if (p[5] <= 0.0) p[5] = FLT_MAX;     /* MAXFLOAT */
            }
            if (p[6] <= 0.0) p[6] = FLT_MAX;
            else if (p[6] < p[3] + p[5] + p[4]) p[6] = p[3] + p[5] + p[4];
            s->cache[0] = (p[1] - p[0])/p[3];
            s->cache[1] = (p[0] - p[1])/p[4];
            int nrep = nc > 7 ? nc - 7 : 0;
            s->pat_on = 0;
            for (int r = -1; r < nrep; r++) {
                double time = (r < 0) ? p[2] : co[7 + r];
                if (p[6] < finaltime) {
                    break_set_lattice(c, time, p[6]);
                    time += p[3]; break_set_lattice(c, time, p[6]);
                    if (p[5] != 0) { time += p[5]; break_set_lattice(c, time, p[6]); }
                    if (p[4] != 0) { time += p[4]; break_set_lattice(c, time, p[6]); }
                }
                else {
                    break_set(c, time);
                    time += p[3]; break_set(c, time);
                    if (p[5] != 0) { time += p[5]; break_set(c, time); }
                    if (p[4] != 0) { time += p[4]; break_set(c, time); }
                }
            }
            s->pat_on = p[6] < finaltime;            /* :1305: the pattern array exists only then */
        }
    }
    else if (s->fn == B200_FN_PWL) {
        /* inptran.cc:2312-2345 (no R= repeat, no TD) */
        s->pwl_index = 0;
        if (s->enable_tran) {
            int n = s->pwl_n/2;
            for (int i = 0; i < n; i++) {
                double t = s->coef[2*i];
                if (t > finaltime)
                    return;
                break_set(c, t);
            }
        }
    }
    else if (s->fn == B200_FN_SIN) {
        /* inptran.cc:2579-2598 */
        double *p = s->parms; const double *co = s->coef; int nc = s->ncoef;
        memset(p, 0, 8*sizeof(double));
        p[0] = co[0];
        p[1] = nc >= 2 ? co[1] : 0.0;
        p[2] = nc >= 3 ? co[2] : 0.0;
        p[3] = nc >= 4 ? co[3] : 0.0;
        p[4] = nc >= 5 ? co[4] : 0.0;
        p[5] = nc >= 6 ? co[5] : 0.0;
        if (s->enable_tran) {
            if (!p[2]) p[2] = 1.0/finaltime;
            if (p[3] > 0.0) break_set(c, p[3]);
        }
    }
    else if (s->fn == B200_FN_EXP) {
        /* IFexpData::setup, inptran.cc:2985-3013 */
        double *p = s->parms; const double *co = s->coef; int nc = s->ncoef;
        memset(p, 0, 8*sizeof(double));
        p[0] = co[0];
        p[1] = nc >= 2 ? co[1] : co[0];
        p[2] = nc >= 3 ? co[2] : 0.0;
        p[3] = nc >= 4 ? co[3] : 0.0;
        p[4] = nc >= 5 ? co[4] : 0.0;
        p[5] = nc >= 6 ? co[5] : 0.0;
        if (s->enable_tran) {
            if (!p[5]) p[5] = step;
            if (!p[2]) p[2] = step;
            if (!p[3]) p[3] = step;
            if (!p[4]) p[4] = p[2] + step;
            if (p[2] > 0.0) break_set(c, p[2]);
            if (p[4] > p[2]) break_set(c, p[4]);
        }
    }
    else if (s->fn == B200_FN_GPULSE) {
        /* IFgpulseData::setup, inptran.cc:1672-1730 (bit patterns not restated) */
        double *p = s->parms; const double *co = s->coef; int nc = s->ncoef;
        const double TWOSQRTLN2 = 2.0*sqrt(M_LN2), PHI0_SQRTPI = PHI0/sqrt(M_PI);  /* :74-75 */
        /* a bit pattern follows a NaN: length, repeat start, bits (include/b200spice.h) */
        s->npat = 0; s->pat = 0;
        for (int k = 0; k < s->ncoef; k++)
            if (isnan(co[k])) { nc = k; s->npat = (int)co[k+1]; s->prst = (int)co[k+2]; s->pat = co + k + 3; break; }
        s->nargs = nc;
        memset(p, 0, 8*sizeof(double));
        p[0] = co[0];
        p[1] = nc >= 2 ? co[1] : co[0];
        p[2] = nc >= 3 ? co[2] : 0.0;
        p[3] = nc >= 4 ? (co[3] >= 0.0 ? co[3]/TWOSQRTLN2 : -co[3]) : 0.0;
        p[4] = nc >= 5 ? co[4] : 0.0;
        if (p[4] < 2*p[3]) p[4] = 0.0;
        if (s->enable_tran) {
            if (p[3] == 0.0) {
                if (p[1] != p[0]) p[3] = PHI0_SQRTPI/fabs(p[1] - p[0]);
                else p[3] = step/TWOSQRTLN2;
            }
            if (p[1] == p[0]) p[1] = p[0] + PHI0_SQRTPI/p[3];
        }
    }
    else if (s->fn == B200_FN_SPULSE) {
        /* IFspulseData::setup, inptran.cc:2790-2808 */
        double *p = s->parms; const double *co = s->coef; int nc = s->ncoef;
        memset(p, 0, 8*sizeof(double));
        p[0] = co[0];
        p[1] = nc >= 2 ? co[1] : co[0];
        p[2] = nc >= 3 ? co[2] : 0.0;
        p[3] = nc >= 4 ? co[3] : 0.0;
        p[4] = nc >= 5 ? co[4] : 0.0;
        if (s->enable_tran) {
            if (!p[2]) p[2] = finaltime;
            if (p[3] > 0.0) break_set(c, p[3]);
        }
    }
    else if (s->fn == B200_FN_SFFM || s->fn == B200_FN_AM) {
        /* IFsffmData::setup inptran.cc:3194-3211, IFamData::setup :3373-3388 */
        double *p = s->parms; const double *co = s->coef; int nc = s->ncoef;
        memset(p, 0, 8*sizeof(double));
        for (int k = 0; k < 5; k++) p[k] = nc > k ? co[k] : 0.0;
        if (s->enable_tran) {
            if (!p[2]) p[2] = 1.0/finaltime;
            if (s->fn == B200_FN_SFFM && !p[4]) p[4] = 1.0/finaltime;
        }
    }
}

static double src_eval(Src *s, double t)
{
    if (s->fn == B200_FN_PULSE) {
        /* inptran.cc:1322-1414 */
        const double *p = s->parms;
        if (!s->enable_tran)
            return p[0];
        double value = p[0];
        double pw = p[3] + p[5];
        for (int i = 6; i < s->nargs || i == 6; i++) {
            /* i = 6: the pulse itself; then the additional offsets */
            double time = t - (i == 6 ? p[2] : s->coef[i]);
            int skip = 0;
            if (p[6] > 0.0) {
                int pnum = 0;
                if (time > p[6]) {
                    pnum = (int)(time/p[6]);
                    time -= (p[6]*pnum);
                }
                if (s->npat && s->pat_on) {
                    int j = pnum;
                    if (pnum >= s->npat) j = s->prst ? s->prst + (pnum - s->prst)%(s->npat - s->prst) : -1;
                    skip = j < 0 || s->pat[j] == 0.0;
                }
            }
            if (skip) continue;
            if (time < p[3] || time >= pw + p[4]) {
                if (i == 6) value = p[0];
                if (time > 0 && time < p[3])
                    value += time*s->cache[0];
            }
            else {
                if (i == 6) value = p[1]; else value += (p[1] - p[0]);
                if (time > pw)
                    value += (time - pw)*s->cache[1];
            }
        }
        return value;
    }
    if (s->fn == B200_FN_PWL) {
        /* inptran.cc:2350-2388 (no repeat) */
        const double *co = s->coef;
        int nvals = s->pwl_n/2;
        if (nvals == 2 && co[3] == co[1])
            return co[1];
        if (t >= co[2*(nvals-1)])
            return co[2*(nvals-1)+1];
        int i = s->pwl_index;
        while (i + 1 < nvals) {
            if (t <= co[2*(i+1)])
                break;
            i++;
        }
        while (t < co[2*i] && i)
            i--;
        s->pwl_index = i;
        if (!i && t <= co[0])
            return co[1];
        return co[2*i+1] + s->pwl_cache[i]*(t - co[2*i]);
    }
    if (s->fn == B200_FN_SIN) {
        /* inptran.cc:2602-2624 */
        const double *p = s->parms;
        if (!s->enable_tran) {
            if (p[5] != 0.0)
                return p[0] + p[1]*sin(M_PI*(p[5]/180));
            return p[0];
        }
        double time = t - p[3];
        if (time <= 0) {
            if (p[5] != 0.0)
                return p[0] + p[1]*sin(M_PI*(p[5]/180));
            return p[0];
        }
        double a;
        if (p[5] != 0.0)
            a = p[1]*sin(2*M_PI*(p[2]*time + (p[5]/360)));
        else
            a = p[1]*sin(2*M_PI*p[2]*time);
        if (p[4] != 0.0)
            a *= exp(-time*p[4]);
        return p[0] + a;
    }
    if (s->fn == B200_FN_EXP) {
        /* inptran.cc:3016-3028 */
        const double *p = s->parms;
        if (!s->enable_tran) return p[0];
        double value = p[0];
        if (t <= p[2]) return value;
        value += (p[1]-p[0])*(1-exp(-(t-p[2])/p[3]));
        if (t <= p[4]) return value;
        value += (p[0]-p[1])*(1-exp(-(t-p[4])/p[5]));
        return value;
    }
    if (s->fn == B200_FN_SFFM) {
        /* inptran.cc:3215-3222 */
        const double *p = s->parms;
        if (!s->enable_tran) return p[0];
        double w1 = 2*M_PI*p[2];
        double w2 = 2*M_PI*p[4];
        return p[0] + p[1]*sin(w1*t + p[3]*sin(w2*t));
    }
    if (s->fn == B200_FN_GPULSE) {
        /* inptran.cc:1734-1827 */
        const double *p = s->parms;
        if (!s->enable_tran) return p[0];
        double V = 0;
        const double rchk = 5.0;
        if (p[4] > 0.0) {
            double per = p[4];
            for (int cidx = 4; cidx < s->nargs || cidx == 4; cidx++) {
                double td = cidx == 4 ? p[2] : s->coef[cidx];
                int ifirst, ilast;
                if (t < td) { ifirst = 0; ilast = 0; }
                else {
                    double del = rchk*p[3];
                    ifirst = (int)rint((t - td - del)/per);
                    ilast = (int)rint((t - td + del)/per);
                }
                for (int i = ifirst; i <= ilast; i++) {
                    if (s->npat) {                  /* inptran.cc:1755-1772 */
                        int j = i;
                        if (i >= s->npat) j = s->prst ? s->prst + (i - s->prst)%(s->npat - s->prst) : -1;
                        if (j < 0 || s->pat[j] == 0.0) continue;
                    }
                    double time = t - i*per;
                    double a = (time - td)/p[3];
                    V += exp(-a*a);
                }
            }
        }
        else {
            double a = (t - p[2])/p[3];
            if (fabs(a) < rchk) V += exp(-a*a);
            for (int i = 5; i < s->nargs; i++) {
                a = (t - s->coef[i])/p[3];
                if (fabs(a) < rchk) V += exp(-a*a);
            }
        }
        return p[0] + (p[1] - p[0])*V;
    }
    if (s->fn == B200_FN_SPULSE) {
        /* inptran.cc:2812-2824 */
        const double *p = s->parms;
        if (!s->enable_tran) return p[0];
        double time = t - p[3];
        if (time <= 0) return p[0];
        if (p[4] != 0.0)
            return p[0] + (p[1]-p[0])*( 1 - cos(2*M_PI*time/p[2])*exp(-time*p[4]) )/2;
        return p[0] + (p[1]-p[0])*( 1 - cos(2*M_PI*time/p[2]) )/2;
    }
    if (s->fn == B200_FN_AM) {
        /* inptran.cc:3391-3399 */
        const double *p = s->parms;
        if (!s->enable_tran) return 0.0;
        if (t <= p[4]) return 0.0;
        double w = 2*M_PI*(t - p[4]);
        return p[0]*(p[1] + sin(w*p[2]))*sin(w*p[3]);
    }
    return 0.0;
}

/* PWL construction: inptran.cc:1980-2074 (no TD, no R) */
static void src_build_pwl(Src *s)
{
    int num = s->ncoef;
    double *list = malloc((num + 3)*sizeof(double));
    memcpy(list, s->coef, num*sizeof(double));
    int add0 = (list[0] > 0.0);
    int add1 = (num & 1);
    if (add0 || add1) {
        int nnum = num + 2*add0 + add1;
        double *nl = malloc((nnum + 1)*sizeof(double)), *d = nl;
        if (add0) { *d++ = 0.0; *d++ = num > 1 ? list[1] : 0.0; }
        for (int i = 0; i < num; i++) *d++ = list[i];
        if (add1) { if (num > 1) *d = *(d-2); else *d = 0.0; }
        free(list); list = nl; num = nnum;
    }
    for (int i = 2; i < num; ) {
        if (list[i] <= list[i-2]) {
            for (int k = i; k+2 < num; k += 2) { list[k] = list[k+2]; list[k+1] = list[k+3]; }
            num -= 2;
            continue;
        }
        i += 2;
    }
    free(s->coef);
    s->coef = list;
    s->pwl_n = num;
    int i = num/2 - 1;
    if (i > 0) {
        s->pwl_cache = malloc(i*sizeof(double));
        for (int k = 0; k < i; k++)
            s->pwl_cache[k] = (list[2*k+3] - list[2*k+1])/(list[2*k+2] - list[2*k]);
    }
    else {
        s->pwl_cache = malloc(sizeof(double));
        s->pwl_cache[0] = 0.0;
    }
}

/* ------------------------------------------------------------------------- */
/* setup: ckt.cc:2110-2227 + per-device setup                                  */
/* ------------------------------------------------------------------------- */
static void mos_setup(wo_ckt *c, Mos *m, int idx)
{
    /* mosset.cc:83-390 */
    b200_mos_model *md = &m->mod;
    if (md->level != 1 && md->level != 2 && md->level != 3 && md->level != 6) md->level = 1;
    if (!GIVEN(md->vto)) md->vto = 0;
    /* kp default 2e-5 applies only if still not given after mostemp's u0/tox rule */
    if (!GIVEN(md->gamma)) md->gamma = 0;
    if (!GIVEN(md->phi)) md->phi = .6;
    if (!GIVEN(md->cbd)) md->cbd = 0;         /* "given" tested on the deck copy below */
    if (!GIVEN(md->cbs)) md->cbs = 0;
    if (!GIVEN(md->is)) md->is = 1e-14;
    if (!GIVEN(md->pb)) md->pb = .8;
    if (!GIVEN(md->cgso)) md->cgso = 0;
    if (!GIVEN(md->cgdo)) md->cgdo = 0;
    if (!GIVEN(md->cgbo)) md->cgbo = 0;
    if (!GIVEN(md->mj)) md->mj = .5;
    if (!GIVEN(md->mjsw)) md->mjsw = (md->level == 1 || md->level == 6) ? .5 : .33;
    if (!GIVEN(md->js)) md->js = 0;
    if (!GIVEN(md->ld)) md->ld = 0;
    if (!GIVEN(md->fc)) md->fc = .5;
    /* level 2/3, mosset.cc:208-253 */
    if (!GIVEN(md->uexp)) md->uexp = 0;
    if (!GIVEN(md->neff)) md->neff = 1;
    if (!GIVEN(md->ucrit)) md->ucrit = 1e4;
    if (!GIVEN(md->nfs)) md->nfs = 0;
    if (!GIVEN(md->delta)) md->delta = 0;
    if (!GIVEN(md->vmax)) md->vmax = 0;
    if (!GIVEN(md->xj)) md->xj = 0;
    if (!GIVEN(md->eta)) md->eta = 0;
    if (!GIVEN(md->theta)) md->theta = 0;
    if (!GIVEN(md->kappa)) md->kappa = .2;
    /* level 6, mosset.cc:255-272 (the "given" tests of lambda / kc are made on the deck card) */
    if (!GIVEN(md->kv)) md->kv = 2;
    if (!GIVEN(md->nv)) md->nv = 0.5;
    if (!GIVEN(md->nc)) md->nc = 1;
    if (!GIVEN(md->lamda0)) md->lamda0 = GIVEN(md->lambda) ? md->lambda : 0;
    if (!GIVEN(md->lamda1)) md->lamda1 = 0;
    if (!GIVEN(md->sigma)) md->sigma = 0;
    if (!GIVEN(md->gamma1)) md->gamma1 = 0;
    if (!GIVEN(md->lambda)) md->lambda = 0;

    m->mode = 1;
    m->von = 0;
    m->st = c->numStates;
    c->numStates += MS_NSTATES;

    const b200_mos *in = &m->in;
    m->mult = GIVEN(in->m) ? in->m : 1.0;
    m->temp = GIVEN(in->temp) ? in->temp : c->opt.temp;
    m->drainPerimeter = GIVEN(in->pd) ? in->pd : 0;
    m->drainSquares = (!GIVEN(in->nrd) || in->nrd == 0) ? 1 : in->nrd;
    m->sourcePerimeter = GIVEN(in->ps) ? in->ps : 0;
    m->sourceSquares = (!GIVEN(in->nrs) || in->nrs == 0) ? 1 : in->nrs;
    m->icVBS = GIVEN(in->icvbs) ? in->icvbs : 0;
    m->icVDS = GIVEN(in->icvds) ? in->icvds : 0;
    m->icVGS = GIVEN(in->icvgs) ? in->icvgs : 0;
    m->vdsat = 0;
    /* circuit.h:1957-1991 mos_default_* */
    m->l = GIVEN(in->l) ? in->l : (c->opt.defl > 0 ? c->opt.defl : 1e-6);
    m->w = GIVEN(in->w) ? in->w : (c->opt.defw > 0 ? c->opt.defw : 1e-6);
    m->drainArea = GIVEN(in->ad) ? in->ad : (c->opt.defad > 0 ? c->opt.defad : 0.0);
    m->sourceArea = GIVEN(in->as) ? in->as : (c->opt.defas > 0 ? c->opt.defas : 0.0);

    double rd = GIVEN(md->rd) ? md->rd : 0, rs = GIVEN(md->rs) ? md->rs : 0;
    double rsh = GIVEN(md->rsh) ? md->rsh : 0;
    if (GIVEN(md->rd))
        m->drainConductance = (md->rd != 0) ? 1/md->rd : 0;
    else if (GIVEN(md->rsh))
        m->drainConductance = (md->rsh != 0) ? 1/(md->rsh*m->drainSquares) : 0;
    else
        m->drainConductance = 0;
    if (GIVEN(md->rs))
        m->sourceConductance = (md->rs != 0) ? 1/md->rs : 0;
    else if (GIVEN(md->rsh))
        m->sourceConductance = (md->rsh != 0) ? 1/(md->rsh*m->sourceSquares) : 0;
    else
        m->sourceConductance = 0;

    m->effectiveLength = m->l - 2*md->ld;
    if (m->effectiveLength < 0)
        m->effectiveLength = 0;

    char nm[64];
    m->d = in->d; m->g = in->g; m->s = in->s; m->b = in->b;
    if (rd != 0 || rsh != 0) {
        snprintf(nm, sizeof(nm), "mos%d#internal#drain", idx);
        m->dp = new_node(c, SP_VOLTAGE, nm);
    }
    else
        m->dp = m->d;
    if (rs != 0 || (m->sourceSquares != 0 && rsh != 0)) {
        snprintf(nm, sizeof(nm), "mos%d#internal#source", idx);
        m->sp = new_node(c, SP_VOLTAGE, nm);
    }
    else
        m->sp = m->s;

    /* mosset.cc:53-77 get_node_ptr */
    int *p = m->ptr;
    for (int i = 0; i < M_NPTR; i++) p[i] = -1;
    p[M_Dd]   = coo_add(c, m->d,  m->d);
    p[M_Gg]   = coo_add(c, m->g,  m->g);
    p[M_Ss]   = coo_add(c, m->s,  m->s);
    p[M_Bb]   = coo_add(c, m->b,  m->b);
    p[M_DPdp] = coo_add(c, m->dp, m->dp);
    p[M_SPsp] = coo_add(c, m->sp, m->sp);
    p[M_Ddp]  = coo_add(c, m->d,  m->dp);
    p[M_Gb]   = coo_add(c, m->g,  m->b);
    p[M_Gdp]  = coo_add(c, m->g,  m->dp);
    p[M_Gsp]  = coo_add(c, m->g,  m->sp);
    p[M_Ssp]  = coo_add(c, m->s,  m->sp);
    p[M_Bg]   = coo_add(c, m->b,  m->g);
    p[M_Bdp]  = coo_add(c, m->b,  m->dp);
    p[M_Bsp]  = coo_add(c, m->b,  m->sp);
    p[M_DPsp] = coo_add(c, m->dp, m->sp);
    p[M_DPd]  = coo_add(c, m->dp, m->d);
    p[M_DPg]  = coo_add(c, m->dp, m->g);
    p[M_DPb]  = coo_add(c, m->dp, m->b);
    p[M_SPg]  = coo_add(c, m->sp, m->g);
    p[M_SPs]  = coo_add(c, m->sp, m->s);
    p[M_SPb]  = coo_add(c, m->sp, m->b);
    p[M_SPdp] = coo_add(c, m->sp, m->dp);
}

static void mos_temp(wo_ckt *c, Mos *m, const b200_mos_model *deckmod)
{
    /* mostemp.cc:65-300.  deckmod = the card as given (for xxxGiven tests). */
    b200_mos_model *md = &m->mod;
    double tnom = GIVEN(deckmod->tnom) ? deckmod->tnom : c->opt.tnom;
    double fact1  = tnom/REFTEMP;
    double vtnom  = tnom*CONSTKoverQ;
    double egfet1 = 1.16 - (7.02e-4*tnom*tnom)/(tnom + 1108);
    double tmp    = tnom*CONSTboltz;
    tmp = -egfet1/(tmp+tmp) + 1.1150877/(CONSTboltz*(REFTEMP+REFTEMP));
    double pbfact1 = -2*vtnom*(1.5*log(fact1) + CHARGE*tmp);

    double tox = GIVEN(deckmod->tox) ? deckmod->tox : ((md->level == 2 || md->level == 3) ? 1e-7 : 0);
    double kp = md->kp;      /* may be NaN = not given */
    double kc = md->kc;      /* level 6; default 5e-5 (mosset.cc:260-261) */
    int skip = 0;
    if (!GIVEN(deckmod->tox) || tox == 0) {
        if (md->level == 1 || md->level == 6) {
            m->oxideCapFactor = 0;
            skip = 1;
        }
        else
            tox = 1e-7;
    }
    if (!skip) {
        m->oxideCapFactor = 3.9*8.854214871e-12/tox;
        m->surfaceMobility = GIVEN(deckmod->u0) ? deckmod->u0 : 600;
        if (md->level == 6) {
            if (!GIVEN(kc))
                kc = .5*m->surfaceMobility*m->oxideCapFactor*M2PCM2;     /* mostemp.cc:100-105 */
        }
        else if (!GIVEN(kp))
            kp = m->surfaceMobility*m->oxideCapFactor*M2PCM2;
        if (GIVEN(deckmod->nsub)) {
            /* mos_dopsetup, mostemp.cc:303-358 */
            double doping = deckmod->nsub*CM3PM3;
            if (doping > 1.45e16) {
                if (!GIVEN(deckmod->phi)) {
                    md->phi = 2*vtnom*log(doping/1.45e16);
                    md->phi = SPMAX(.1, md->phi);
                }
                double gateType = GIVEN(deckmod->tpg) ? deckmod->tpg : 1;
                if (!GIVEN(deckmod->gamma))
                    md->gamma = sqrt(2*EPSSIL*CHARGE*doping)/m->oxideCapFactor;
                if (!GIVEN(deckmod->vto)) {
                    double wkfng = 3.2;
                    if (gateType != 0) {
                        if (md->type > 0)
                            wkfng = 3.25 + .5*egfet1*(1 - gateType);
                        else
                            wkfng = 3.25 + .5*egfet1*(1 + gateType);
                    }
                    double wkfngs;
                    if (md->type > 0)
                        wkfngs = wkfng - (3.25 + .5*(egfet1 + md->phi));
                    else
                        wkfngs = wkfng - (3.25 + .5*(egfet1 - md->phi));
                    double vfb;
                    if (!GIVEN(deckmod->nss))
                        vfb = wkfngs;
                    else
                        vfb = wkfngs - deckmod->nss*CM2PM2*CHARGE/m->oxideCapFactor;
                    if (md->type > 0)
                        md->vto = vfb + (md->gamma*sqrt(md->phi) + md->phi);
                    else
                        md->vto = vfb - (md->gamma*sqrt(md->phi) + md->phi);
                }
            }
        }
        if (GIVEN(deckmod->nsub)) {
            /* mostemp.cc:119-128 */
            if (md->level == 2)
                m->xd = sqrt((EPSSIL+EPSSIL)/(CHARGE*deckmod->nsub*CM3PM3));
            else if (md->level == 3) {
                m->alpha = (EPSSIL+EPSSIL)/(CHARGE*deckmod->nsub*CM3PM3);
                m->xd = sqrt(m->alpha);
            }
        }
    }
    else
        m->surfaceMobility = GIVEN(deckmod->u0) ? deckmod->u0 : 0;
    /* mostemp.cc:138-142 computes narrowFactor = MOSdelta*pi/2*EPSSIL/oxideCapFactor for level 3, but
     * the card's "delta" is stored in MOSnarrowFactor (mossetm.cc:216-219) and MOSdelta keeps its
     * default 0 (mosset.cc:230-231): the narrow-width term is always 0 in the reference. */
    m->narrowFactor = (md->level == 2 && GIVEN(deckmod->delta)) ? deckmod->delta : 0.0;   /* level 2: mossetm.cc:216-219 */
    if (!GIVEN(kp))
        kp = 2e-5;                       /* mosset.cc:101-103 */
    if (!GIVEN(kc))
        kc = 5e-5;
    if (md->level == 6)
        kp = kc;                         /* mostemp.cc:156-159: tTransconductance = kc*tmp */
    m->kp_eff = kp;
    m->nsub = GIVEN(deckmod->nsub) ? deckmod->nsub : 0;
    double cj = GIVEN(deckmod->cj) ? deckmod->cj : 0;
    double cjsw = GIVEN(deckmod->cjsw) ? deckmod->cjsw : 0;

    double fact2 = m->temp/REFTEMP;
    double vt    = m->temp*CONSTKoverQ;
    double ratio = m->temp/tnom;
    double egfet = 1.16 - (7.02e-4*m->temp*m->temp)/(m->temp + 1108);
    tmp = m->temp*CONSTboltz;
    tmp = -egfet/(tmp+tmp) + 1.1150877/(CONSTboltz*(REFTEMP+REFTEMP));
    double pbfact = -2*vt*(1.5*log(fact2) + CHARGE*tmp);

    tmp = 1/(ratio*sqrt(ratio));
    m->tTransconductance = kp*tmp;
    m->tSurfMob = m->surfaceMobility*tmp;

    tmp = (md->phi - pbfact1)/fact1;
    m->tPhi = fact2*tmp + pbfact;

    if (md->type > 0) {
        m->tVbi = md->vto - md->gamma*sqrt(md->phi) + .5*(egfet1 - egfet + m->tPhi - md->phi);
        m->tVto = m->tVbi + md->gamma*sqrt(m->tPhi);
    }
    else {
        m->tVbi = md->vto + md->gamma*sqrt(md->phi) + .5*(egfet1 - egfet - (m->tPhi - md->phi));
        m->tVto = m->tVbi - md->gamma*sqrt(m->tPhi);
    }

    tmp = exp(-egfet/vt + egfet1/vtnom);
    m->tSatCur = md->is*tmp;
    m->tSatCurDens = md->js*tmp;

    double pbo = (md->pb - pbfact1)/fact1;
    double gmaold = (md->pb - pbo)/pbo;
    double capfact = 1/(1 + md->mj*(4e-4*(tnom - REFTEMP) - gmaold));
    m->tCbd = md->cbd*capfact;
    m->tCbs = md->cbs*capfact;
    m->tCj  = cj*capfact;
    capfact = 1/(1 + md->mjsw*(4e-4*(tnom - REFTEMP) - gmaold));
    m->tCjsw = cjsw*capfact;
    m->tBulkPot = fact2*pbo + pbfact;
    double gmanew = (m->tBulkPot - pbo)/pbo;
    capfact = (1 + md->mj*(4e-4*(m->temp - REFTEMP) - gmanew));
    m->tCbd *= capfact;
    m->tCbs *= capfact;
    m->tCj  *= capfact;
    capfact = (1 + md->mjsw*(4e-4*(m->temp - REFTEMP) - gmanew));
    m->tCjsw *= capfact;
    m->tDepCap = md->fc*m->tBulkPot;

    if ((m->tSatCurDens == 0) || (m->drainArea == 0) || (m->sourceArea == 0)) {
        m->sourceVcrit = m->drainVcrit = vt*log(vt/(CONSTroot2*m->tSatCur));
        m->tDrainSatCur = m->tSatCur;
        m->tSourceSatCur = m->tSatCur;
    }
    else {
        m->tDrainSatCur = m->tSatCurDens*m->drainArea;
        m->tSourceSatCur = m->tSatCurDens*m->sourceArea;
        m->drainVcrit = vt*log(vt/(CONSTroot2*m->tDrainSatCur));
        m->sourceVcrit = vt*log(vt/(CONSTroot2*m->tSourceSatCur));
    }

    if (GIVEN(deckmod->cbd)) m->Cbd = m->tCbd;
    else m->Cbd = GIVEN(deckmod->cj) ? m->tCj*m->drainArea : 0;
    m->Cbdsw = GIVEN(deckmod->cjsw) ? m->tCjsw*m->drainPerimeter : 0;
    if (GIVEN(deckmod->cbs)) m->Cbs = m->tCbs;
    else m->Cbs = GIVEN(deckmod->cj) ? m->tCj*m->sourceArea : 0;
    m->Cbssw = GIVEN(deckmod->cjsw) ? m->tCjsw*m->sourcePerimeter : 0;

    /* MOSdev::fd / fs, mostemp.cc:335-389 */
    {
        double arg = 1 - md->fc;
        double sarg = exp((-md->mj)*log(arg));
        double sargsw = exp((-md->mjsw)*log(arg));
        double pot = m->tBulkPot;
        double tmp1 = (1 - md->fc*(1 + md->mj))/arg;
        double tmp2 = 1/(arg*pot);
        m->f2d = (m->Cbd*sarg + m->Cbdsw*sargsw)*tmp1;
        m->f3d = (m->Cbd*md->mj*sarg + m->Cbdsw*md->mjsw*sargsw)*tmp2;
        m->f4d = m->Cbd*pot*(1 - arg*sarg)/(1 - md->mj)
            + m->Cbdsw*pot*(1 - arg*sargsw)/(1 - md->mjsw)
            - m->tDepCap*(m->f2d + .5*m->f3d*m->tDepCap);
        m->f2s = (m->Cbs*sarg + m->Cbssw*sargsw)*tmp1;
        m->f3s = (m->Cbs*md->mj*sarg + m->Cbssw*md->mjsw*sargsw)*tmp2;
        m->f4s = m->Cbs*pot*(1 - arg*sarg)/(1 - md->mj)
            + m->Cbssw*pot*(1 - arg*sargsw)/(1 - md->mjsw)
            - m->tDepCap*(m->f2s + .5*m->f3s*m->tDepCap);
    }

    m->gateSourceOverlapCap = md->cgso*m->w;
    m->gateDrainOverlapCap = md->cgdo*m->w;
    m->gateBulkOverlapCap = md->cgbo*m->effectiveLength;
    m->beta = m->tTransconductance*m->w/m->effectiveLength;
    m->oxideCap = m->oxideCapFactor*m->effectiveLength*m->w;
}

static void jj_setup_instance(JJ *j);
static void tjm_setup_instance(TJM *j);
static void wo_sincos(double x, double *s, double *co) { *s = sin(x); *co = cos(x); }

static void build_matrix(wo_ckt *c)
{
    /* spbuild.cc:922-958: CSC, columns ascending, rows ascending within a column */
    int n = c->n;
    c->nnz = c->ncoo;
    c->Ap = xcalloc(n + 2, sizeof(int));
    c->Ai = xcalloc(c->nnz, sizeof(int));
    c->Ax = xcalloc(c->nnz, sizeof(double));
    c->Ainit = xcalloc(c->nnz, sizeof(double));
    int *cnt = xcalloc(n + 2, sizeof(int));
    for (int i = 0; i < c->ncoo; i++) cnt[c->coo_c[i]]++;     /* cols 1..n */
    c->Ap[0] = 0;
    for (int j = 1; j <= n; j++) c->Ap[j] = c->Ap[j-1] + cnt[j];
    int *map = xcalloc(c->ncoo, sizeof(int));
    int *fill = xcalloc(n + 2, sizeof(int));
    /* insertion sort rows within each column */
    for (int i = 0; i < c->ncoo; i++) {
        int col = c->coo_c[i] - 1, r = c->coo_r[i] - 1;
        int base = c->Ap[col], k = fill[col]++;
        int pos = base + k;
        while (pos > base && c->Ai[pos-1] > r) { c->Ai[pos] = c->Ai[pos-1]; pos--; }
        c->Ai[pos] = r;
    }
    for (int i = 0; i < c->ncoo; i++) {
        int col = c->coo_c[i] - 1, r = c->coo_r[i] - 1;
        for (int p = c->Ap[col]; p < c->Ap[col+1]; p++)
            if (c->Ai[p] == r) { map[i] = p; break; }
    }
#define REMAP(x) do { if ((x) >= 0) (x) = map[(x)]; } while (0)
    for (int i = 0; i < c->nmos; i++)
        for (int k = 0; k < M_NPTR; k++) REMAP(c->mos[i].ptr[k]);
    for (int i = 0; i < c->ncap; i++) {
        REMAP(c->cap[i].pPP); REMAP(c->cap[i].pNN); REMAP(c->cap[i].pPN); REMAP(c->cap[i].pNP);
    }
    for (int i = 0; i < c->nres; i++) {
        REMAP(c->res[i].pPP); REMAP(c->res[i].pNN); REMAP(c->res[i].pPN); REMAP(c->res[i].pNP);
    }
    for (int i = 0; i < c->nind; i++) {
        REMAP(c->ind[i].pBB); REMAP(c->ind[i].pPB); REMAP(c->ind[i].pNB); REMAP(c->ind[i].pBN); REMAP(c->ind[i].pBP);
    }
    for (int i = 0; i < c->nmut; i++) { REMAP(c->mut[i].p12); REMAP(c->mut[i].p21); }
    for (int i = 0; i < c->njj; i++) {
        JJ *j = &c->jj[i];
        REMAP(j->pPP); REMAP(j->pNN); REMAP(j->pNP); REMAP(j->pPN); REMAP(j->pHH);
        REMAP(j->pRPP); REMAP(j->pRPN); REMAP(j->pRNP); REMAP(j->pRNN);
    }
    for (int i = 0; i < c->ntjm; i++) {
        TJM *j = &c->tjm[i];
        REMAP(j->pPP); REMAP(j->pNN); REMAP(j->pNP); REMAP(j->pPN); REMAP(j->pHH);
        REMAP(j->pRPP); REMAP(j->pRPN); REMAP(j->pRNP); REMAP(j->pRNN);
    }
    for (int i = 0; i < c->nsrc; i++) {
        REMAP(c->src[i].pPB); REMAP(c->src[i].pNB); REMAP(c->src[i].pBN); REMAP(c->src[i].pBP);
    }
#undef REMAP
    c->diag = xcalloc(n + 1, sizeof(int));
    for (int j = 1; j <= n; j++) {
        c->diag[j] = -1;
        for (int p = c->Ap[j-1]; p < c->Ap[j]; p++)
            if (c->Ai[p] == j-1) c->diag[j] = p;
    }
    free(cnt); free(map); free(fill);
}

wo_ckt *wo_create(const b200_deck *dk)
{
    wo_ckt *c = xcalloc(1, sizeof(*c));
    c->opt = dk->opt;
    c->tstep = dk->tstep; c->tstop = dk->tstop; c->tstart = dk->tstart; c->tmax = dk->tmax;
    c->uic = dk->uic;
    c->nnodes_ext = dk->nnodes;
    c->n = 0;
    c->ntype = xcalloc(1, sizeof(int));
    c->names = xcalloc(1, sizeof(char*));
    c->names[0] = strdup("0");
    for (int i = 1; i < dk->nnodes; i++) {
        char nm[32];
        snprintf(nm, sizeof(nm), "n%d", i);
        new_node(c, SP_VOLTAGE, nm);
    }
    c->nmos = dk->nmos; c->ncap = dk->ncap; c->nres = dk->nres; c->nsrc = dk->nsrc;
    c->mos = xcalloc(c->nmos, sizeof(Mos));
    c->cap = xcalloc(c->ncap, sizeof(Cap));
    c->res = xcalloc(c->nres, sizeof(Res));
    c->src = xcalloc(c->nsrc, sizeof(Src));
    c->nind = dk->nind;
    c->ind = xcalloc(c->nind, sizeof(Ind));
    c->nmut = dk->nmut;
    c->mut = xcalloc(c->nmut, sizeof(Mut));
    c->njj = dk->njj;
    c->jj = xcalloc(c->njj, sizeof(JJ));
    c->ntjm = dk->ntjm;
    c->tjm = xcalloc(c->ntjm, sizeof(TJM));
    c->nord = dk->norder;
    c->ord = xcalloc(c->nord, sizeof(DevRef));
    for (int i = 0; i < c->nord; i++) {
        c->ord[i].kind = dk->order_kind[i];
        c->ord[i].index = dk->order_index[i];
    }
    for (int i = 0; i < c->nmos; i++) {
        c->mos[i].in = dk->mos[i];
        c->mos[i].mod = dk->mos_models[dk->mos[i].model];
        c->mos[i].card = c->mos[i].mod;
    }

    /* setup in load order (ckt.cc:2202-2210) */
    for (int o = 0; o < c->nord; o++) {
        int i = c->ord[o].index;
        char nm[64];
        switch (c->ord[o].kind) {
        case B200_DEV_MOS:
            mos_setup(c, &c->mos[i], i);
            break;
        case B200_DEV_CAP: {
            /* capset.cc:62-103 */
            Cap *p = &c->cap[i]; const b200_cap *in = &dk->cap[i];
            p->pos = in->pos; p->neg = in->neg;
            double mm = GIVEN(in->m) ? in->m : 1.0;
            p->capac = in->c*1.0*mm;          /* captemp.cc:98 nom*tcFactor*m */
            p->m = mm;
            p->icGiven = GIVEN(in->ic);
            p->ic = p->icGiven ? in->ic : 0.0;
            p->st = c->numStates; c->numStates += 2;
            p->pPP = coo_add(c, p->pos, p->pos);
            p->pNN = coo_add(c, p->neg, p->neg);
            p->pPN = coo_add(c, p->pos, p->neg);
            p->pNP = coo_add(c, p->neg, p->pos);
            break; }
        case B200_DEV_RES: {
            /* resset.cc:74-77, restemp.cc:171-209 */
            Res *p = &c->res[i]; const b200_res *in = &dk->res[i];
            p->pos = in->pos; p->neg = in->neg;
            p->pPP = coo_add(c, p->pos, p->pos);
            p->pNN = coo_add(c, p->neg, p->neg);
            p->pPN = coo_add(c, p->pos, p->neg);
            p->pNP = coo_add(c, p->neg, p->pos);
            double r = in->r, mm = GIVEN(in->m) ? in->m : 1.0;
            if (fabs(r) < c->opt.resmin)
                r = (r < 0.0) ? -c->opt.resmin : c->opt.resmin;
            p->conduct = mm/(r*1.0);
            p->m = mm;
            break; }
        case B200_DEV_SRC: {
            /* srcset.cc:64-273 */
            Src *p = &c->src[i]; const b200_src *in = &dk->src[i];
            p->kind = in->kind; p->pos = in->pos; p->neg = in->neg;
            p->dcGiven = GIVEN(in->dc); p->dc = p->dcGiven ? in->dc : 0.0;
            p->fn = in->fn; p->ncoef = in->ncoef;
            p->coef = malloc((in->ncoef + 1)*sizeof(double));
            memcpy(p->coef, in->coef, in->ncoef*sizeof(double));
            if (p->fn == B200_FN_PWL) src_build_pwl(p);
            p->pPB = p->pNB = p->pBN = p->pBP = -1;
            if (p->kind == B200_SRC_V) {
                snprintf(nm, sizeof(nm), "src%d#branch", i);
                p->branch = new_node(c, SP_CURRENT, nm);
                p->pPB = coo_add(c, p->pos, p->branch);
                p->pNB = coo_add(c, p->neg, p->branch);
                p->pBN = coo_add(c, p->branch, p->neg);
                p->pBP = coo_add(c, p->branch, p->pos);
            }
            if (p->kind == B200_SRC_I && p->fn != B200_FN_NONE)
                c->numStates += 1;
            break; }
        case B200_DEV_IND: {
            /* indset.cc:72-135: branch "name#branch", 2 states, TSTALLOC order :50-58 */
            Ind *p = &c->ind[i]; const b200_ind *in = &dk->ind[i];
            p->pos = in->pos; p->neg = in->neg;
            double mm = GIVEN(in->m) ? in->m : 1.0;
            p->induct = in->l/mm;
            p->icGiven = GIVEN(in->ic);
            p->ic = p->icGiven ? in->ic : 0.0;
            snprintf(nm, sizeof(nm), "ind%d#branch", i);
            p->br = new_node(c, SP_CURRENT, nm);
            p->st = c->numStates; c->numStates += 2;
            p->pBB = coo_add(c, p->br, p->br);
            p->pPB = coo_add(c, p->pos, p->br);
            p->pNB = coo_add(c, p->neg, p->br);
            p->pBN = coo_add(c, p->br, p->neg);
            p->pBP = coo_add(c, p->br, p->pos);
            break; }
        case B200_DEV_MUT:
            break;              /* set up after the inductors, below */
        case B200_DEV_JJ: {
            /* jjset.cc:130-200 (pointers), :609-715 + :794 (instance part of JJdev::setup).
             * The reference sets JJ models up before every other type (ckt.cc:2140-2157); with
             * no lser / lsh they create no nodes, so only the state offsets differ. */
            JJ *j = &c->jj[i];
            j->in = dk->jj[i]; j->mod = dk->jj_models[dk->jj[i].model];
            j->pos = j->in.pos; j->neg = j->in.neg; j->phs = j->in.phase_node;
            j->pPP = coo_add(c, j->pos, j->pos);
            j->pNN = coo_add(c, j->neg, j->neg);
            j->pNP = coo_add(c, j->neg, j->pos);
            j->pPN = coo_add(c, j->pos, j->neg);
            j->pRPP = coo_add(c, j->pos, j->pos);
            j->pRPN = coo_add(c, j->pos, j->neg);
            j->pRNP = coo_add(c, j->neg, j->pos);
            j->pRNN = coo_add(c, j->neg, j->neg);
            j->pHH = j->phs > 0 ? coo_add(c, j->phs, j->phs) : -1;
            j->st = c->numStates; c->numStates += JS_NSTATES;
            jj_setup_instance(j);
            break; }
        case B200_DEV_TJM: {
            /* tjmset.cc:130-200 pointers (same order as the RSJ model), :794 states */
            TJM *j = &c->tjm[i];
            j->in = dk->tjm[i]; j->mod = dk->tjm_models[dk->tjm[i].model];
            j->n = j->mod.nterms;
            if (j->n > TJM_MAXTERMS) { fprintf(stderr, "wrs_oracle: TJM nterms > %d\n", TJM_MAXTERMS); abort(); }
            for (int k = 0; k < j->n; k++) {
                j->p[k].re = j->in.coef[2*k];             j->p[k].im = j->in.coef[2*k+1];
                j->A[k].re = j->in.coef[2*(j->n + k)];    j->A[k].im = j->in.coef[2*(j->n + k)+1];
                j->B[k].re = j->in.coef[2*(2*j->n + k)];  j->B[k].im = j->in.coef[2*(2*j->n + k)+1];
            }
            j->pos = j->in.pos; j->neg = j->in.neg; j->phs = j->in.phase_node;
            j->pPP = coo_add(c, j->pos, j->pos);
            j->pNN = coo_add(c, j->neg, j->neg);
            j->pNP = coo_add(c, j->neg, j->pos);
            j->pPN = coo_add(c, j->pos, j->neg);
            j->pRPP = coo_add(c, j->pos, j->pos);
            j->pRPN = coo_add(c, j->pos, j->neg);
            j->pRNP = coo_add(c, j->neg, j->pos);
            j->pRNN = coo_add(c, j->neg, j->neg);
            j->pHH = j->phs > 0 ? coo_add(c, j->phs, j->phs) : -1;
            j->st = c->numStates; c->numStates += TS_NSTATES;
            tjm_setup_instance(j);
            break; }
        default:
            fprintf(stderr, "wrs_oracle: device kind %d not supported\n", c->ord[o].kind);
            abort();
        }
    }
    /* mutual inductors: INDdev::setup does them after all inductors (indset.cc:186-238) */
    for (int o = 0; o < c->nord; o++) {
        if (c->ord[o].kind != B200_DEV_MUT) continue;
        int i = c->ord[o].index;
        Mut *p = &c->mut[i]; const b200_mut *in = &dk->mut[i];
        p->ind1 = in->ind1; p->ind2 = in->ind2; p->coupling = in->k; p->factor = 0.0;
        p->p12 = coo_add(c, c->ind[p->ind1].br, c->ind[p->ind2].br);
        p->p21 = coo_add(c, c->ind[p->ind2].br, c->ind[p->ind1].br);
    }
    build_matrix(c);

    /* .ic / .nodeset */
    c->ic_val = xcalloc(c->n + 1, sizeof(double)); c->ic_given = xcalloc(c->n + 1, 1);
    c->ns_val = xcalloc(c->n + 1, sizeof(double)); c->ns_given = xcalloc(c->n + 1, 1);
    for (int i = 0; i < dk->nic; i++) { c->ic_val[dk->ic_node[i]] = dk->ic_val[i]; c->ic_given[dk->ic_node[i]] = 1; }
    for (int i = 0; i < dk->nnodeset; i++) { c->ns_val[dk->ns_node[i]] = dk->ns_val[i]; c->ns_given[dk->ns_node[i]] = 1; }

    int n = c->n;
    c->rhs = xcalloc(n + 1, sizeof(double));
    c->rhsOld = xcalloc(n + 1, sizeof(double));
    c->rhsSpare = xcalloc(n + 1, sizeof(double));
    for (int i = 0; i < 8; i++) c->sols[i] = xcalloc(n + 1, sizeof(double));
    for (int i = 0; i <= c->opt.maxord + 1 && i < 8; i++)
        c->states[i] = xcalloc(c->numStates, sizeof(double));
    c->work = xcalloc(n + 1, sizeof(double));
    c->srcFact = 1.0;
    c->phase = 0;
    for (int i = 0; i < c->njj; i++)
        if (c->jj[i].mod.ictype > 0) {
            if (!c->phase) c->phase = xcalloc(n + 1, 1);
            c->phase[c->jj[i].pos] = 1; c->phase[c->jj[i].neg] = 1;
        }
    for (int i = 0; i < c->ntjm; i++) {                          /* tjmset.cc:690-708 (any ictype) */
        if (!c->phase) c->phase = xcalloc(n + 1, 1);
        c->phase[c->tjm[i].pos] = 1; c->phase[c->tjm[i].neg] = 1;
    }
    if (c->phase) {
        for (int i = 0; i < c->nind; i++) { c->phase[c->ind[i].pos] = 1; c->phase[c->ind[i].neg] = 1; }
        c->phase[0] = 0;
    }
    return c;
}

int wo_set_param(wo_ckt *c, int param, int index, double value)
{
    switch (param) {
    case B200_P_MOS_VTO:    c->mos[index].mod.vto = c->mos[index].card.vto = value; break;
    case B200_P_MOS_KP:     /* level 6: the transconductance card is kc */
        if (c->mos[index].mod.level == 6) c->mos[index].mod.kc = c->mos[index].card.kc = value;
        else c->mos[index].mod.kp = c->mos[index].card.kp = value;
        break;
    case B200_P_MOS_GAMMA:  c->mos[index].mod.gamma = c->mos[index].card.gamma = value; break;
    case B200_P_MOS_LAMBDA: c->mos[index].mod.lambda = c->mos[index].card.lambda = value; break;
    case B200_P_MOS_PHI:    c->mos[index].mod.phi = c->mos[index].card.phi = value; break;
    case B200_P_MOS_L:      c->mos[index].in.l = value; c->mos[index].l = value;
        c->mos[index].effectiveLength = value - 2*c->mos[index].mod.ld;
        if (c->mos[index].effectiveLength < 0) c->mos[index].effectiveLength = 0;
        break;
    case B200_P_MOS_W:      c->mos[index].in.w = value; c->mos[index].w = value; break;
    case B200_P_CAP_C:      c->cap[index].capac = value*1.0*c->cap[index].m; break;
    case B200_P_RES_R: {
        double r = value;
        if (fabs(r) < c->opt.resmin) r = (r < 0.0) ? -c->opt.resmin : c->opt.resmin;
        c->res[index].conduct = c->res[index].m/(r*1.0);        /* restemp.cc:200-209 */
        break; }
    case B200_P_SRC_DC:     c->src[index].dc = value; c->src[index].dcGiven = 1; break;
    case B200_P_JJ_ICS:     c->jj[index].in.ics = value; break;
    case B200_P_TJM_ICS:    c->tjm[index].in.ics = value; break;
    case B200_P_IND_L:      c->ind[index].induct = value/1.0;
        for (int k = 0; k < c->nmut; k++) c->mut[k].factor = 0.0;   /* recomputed at the next load */
        break;
    default: return B200_E_BADPARM;
    }
    return OK;
}

void wo_set_ordering(wo_ckt *c, const int *rowperm, const int *colperm)
{
    free(c->rowperm); free(c->colperm);
    c->rowperm = malloc(c->n*sizeof(int));
    c->colperm = malloc(c->n*sizeof(int));
    memcpy(c->rowperm, rowperm, c->n*sizeof(int));
    memcpy(c->colperm, colperm, c->n*sizeof(int));
    c->have_order = 1;
    c->lu_ready = 0;
}

void wo_set_ordering_tran(wo_ckt *c, const int *rowperm, const int *colperm)
{
    free(c->rowperm2); free(c->colperm2);
    c->rowperm2 = malloc(c->n*sizeof(int));
    c->colperm2 = malloc(c->n*sizeof(int));
    memcpy(c->rowperm2, rowperm, c->n*sizeof(int));
    memcpy(c->colperm2, colperm, c->n*sizeof(int));
    c->have_order2 = 1;
    c->lu_ready2 = 0;
}

int wo_num_unknowns(const wo_ckt *c) { return c->n; }
int wo_nnz(const wo_ckt *c) { return c->nnz; }
void wo_pattern(const wo_ckt *c, const int **Ap, const int **Ai) { *Ap = c->Ap; *Ai = c->Ai; }
int wo_unknown_names(const wo_ckt *c, char *buf, int buflen)
{
    int pos = 0;
    for (int i = 1; i <= c->n; i++) {
        int l = (int)strlen(c->names[i]) + 1;
        if (buf && pos + l <= buflen) memcpy(buf + pos, c->names[i], l);
        pos += l;
    }
    return pos;
}

/* ------------------------------------------------------------------------- */
/* integration: circuit.h:1995-2029, niinteg.cc:58-152, nicomcof.cc:58-147      */
/* ------------------------------------------------------------------------- */
static double NIceq(wo_ckt *c, int q)
{
    double cq = 0;
    if (c->opt.method == 1) {
        switch (c->order) {
        case 6: cq += c->ag[6]*c->states[6][q]; /* fallthrough */
        case 5: cq += c->ag[5]*c->states[5][q]; /* fallthrough */
        case 4: cq += c->ag[4]*c->states[4][q]; /* fallthrough */
        case 3: cq += c->ag[3]*c->states[3][q]; /* fallthrough */
        default:
        case 2: cq += c->ag[2]*c->states[2][q]; /* fallthrough */
        case 1: cq += c->ag[1]*c->states[1][q]; break;
        }
    }
    else {
        if (c->order == 1)
            cq = c->ag[1]*c->states[1][q];
        else
            cq = - c->states[1][q+1]*c->ag[1] - c->ag[0]*c->states[1][q];
    }
    return cq;
}

static double find_ceq(wo_ckt *c, int q)
{
    if (c->opt.method == 0)
        return (c->order <= 1 ? c->states[1][q]*c->ag[1] :
            - c->states[1][q+1]*c->ag[1] - c->states[1][q]*c->ag[0]);
    return NIceq(c, q);
}

static int NItrapCheck(wo_ckt *c, int q)
{
    int cc = q + 1;
    double c0 = c->states[0][cc];
    if (fabs(c0) < 1e-6) return 0;
    double q0 = c->states[0][q];
    if (fabs(q0) < 1e-16) return 0;
    double c1 = c->states[1][cc], c2 = c->states[2][cc], q1 = c->states[1][q];
    double dqdt = fabs(q0 - q1)/c->delta;
    double tr = c->opt.trapratio;
    return (fabs(c1) > tr*dqdt && fabs(c0 - c1) > tr*fabs(c0 - c2));
}

static void integrate4(wo_ckt *c, double *geq, double *ceq, double cap, int q)
{
    double cq = find_ceq(c, q);
    c->states[0][q+1] = c->ag[0]*c->states[0][q] + cq;
    *ceq = cq;
    *geq = c->ag[0]*cap;
    if (c->trapCheck && !c->trapBad)
        c->trapBad = NItrapCheck(c, q);
}

static void integrate2(wo_ckt *c, int q, double cq)
{
    c->states[0][q+1] = c->ag[0]*c->states[0][q] + cq;
    if (c->trapCheck && !c->trapBad)
        c->trapBad = NItrapCheck(c, q);
}

static int NIcomCof(wo_ckt *c)
{
    if (c->opt.method == 0) {
        if (c->order == 2) {
            double t0 = 1.0/(1.0 - c->opt.xmu);
            c->ag[0] = t0/c->delta;
            c->ag[1] = t0*c->opt.xmu;
            return OK;
        }
        if (c->order == 1) {
            c->ag[0] = 1.0/c->delta;
            c->ag[1] = -c->ag[0];
            return OK;
        }
        return B200_E_ORDER;
    }
    /* GEAR: nicomcof.cc:79-147 */
    if (c->order < 1 || c->order > 6) return B200_E_ORDER;
    double mat[8][8];
    for (int i = 0; i < 7; i++) c->ag[i] = 0;
    c->ag[1] = -1.0/c->delta;
    int i, ord = c->order;
    for (i = 0; i <= ord; i++) mat[0][i] = 1;
    for (i = 1; i <= ord; i++) mat[i][0] = 0;
    double arg = 0;
    for (i = 1; i <= ord; i++) {
        arg += c->deltaOld[i-1];
        double arg1 = 1;
        for (int j = 1; j <= ord; j++) { arg1 *= arg/c->delta; mat[j][i] = arg1; }
    }
    for (i = 1; i <= ord; i++)
        for (int j = i+1; j <= ord; j++) {
            mat[j][i] /= mat[i][i];
            for (int k = i+1; k <= ord; k++) mat[j][k] -= mat[j][i]*mat[i][k];
        }
    for (i = 1; i <= ord; i++)
        for (int j = i+1; j <= ord; j++) c->ag[j] = c->ag[j] - mat[j][i]*c->ag[i];
    c->ag[ord] /= mat[ord][ord];
    for (i = ord-1; i >= 0; i--) {
        for (int j = i+1; j <= ord; j++) c->ag[i] = c->ag[i] - mat[i][j]*c->ag[j];
        c->ag[i] /= mat[i][i];
    }
    return OK;
}

static void predict(wo_ckt *c)
{
    /* ckt.cc:2020-2087 */
    if (c->mode & MODETRAN) {
        double d1 = c->delta;
        c->pred[1] = -d1/c->deltaOld[1];
        c->pred[0] = 1 - c->pred[1];
        c->pred[2] = c->pred[3] = 0;
    }
    else {
        c->pred[0] = 1.0; c->pred[1] = c->pred[2] = c->pred[3] = 0.0;
    }
}

/* ------------------------------------------------------------------------- */
/* limiters: devlib/devlib.cc:475-620                                          */
/* ------------------------------------------------------------------------- */
static double limvds(double vnew, double vold)
{
    if (vold >= 3.5) {
        if (vnew > vold)
            vnew = SPMIN(vnew, 3*vold + 2);
        else {
            if (vnew < 3.5)
                vnew = SPMAX(vnew, 2);
        }
    }
    else {
        if (vnew > vold)
            vnew = SPMIN(vnew, 4);
        else
            vnew = SPMAX(vnew, -.5);
    }
    return vnew;
}

static double pnjlim(double vnew, double vold, double vt, double vcrit, int *icheck)
{
    if (vnew > vcrit && fabs(vnew - vold) > (vt + vt)) {
        if (vold > 0) {
            double arg = (vnew - vold)/vt;
            if (arg > 0)
                vnew = vold + vt*(2.0 + log(arg - 1.8));
            else
                vnew = vold - vt*(2.0 + log(-arg + 2.0));
        }
        else
            vnew = vt*log(vnew/vt);
        *icheck = 1;
    }
    else {
        if (vnew < 0) {
            double arg;
            if (vold > 0)
                arg = -1*vold - 1;
            else
                arg = 2*vold - 1;
            if (vnew < arg) {
                vnew = arg;
                *icheck = 1;
            }
            else
                *icheck = 0;
        }
        else
            *icheck = 0;
    }
    return vnew;
}

static double fetlim(double vnew, double vold, double vto)
{
    double tmp = (vold - vto);
    if (tmp < 0.0)
        tmp = -tmp;
    double vtsthi = tmp + tmp + 2.0;
    double vtstlo = tmp + 1.0;
    double vtox = vto + 3.5;
    double delv = vnew - vold;
    if (vold >= vto) {
        if (vold >= vtox) {
            if (delv <= 0.0) {
                if (vnew >= vtox) {
                    if (-delv > vtstlo)
                        vnew = vold - vtstlo;
                }
                else
                    vnew = SPMAX(vnew, vto + 2);
            }
            else {
                if (delv >= vtsthi)
                    vnew = vold + vtsthi;
            }
        }
        else {
            if (delv <= 0)
                vnew = SPMAX(vnew, vto - 0.5);
            else
                vnew = SPMIN(vnew, vto + 4.0);
        }
    }
    else {
        if (delv <= 0) {
            if (-delv > vtsthi)
                vnew = vold - vtsthi;
        }
        else {
            double vtemp = vto + .5;
            if (vnew <= vtemp) {
                if (delv > vtstlo)
                    vnew = vold + vtstlo;
            }
            else
                vnew = vtemp;
        }
    }
    return vnew;
}

/* ------------------------------------------------------------------------- */
/* MOS load: devlib/mos/mosload.cc:107-1041, moseq.cc:76-141                    */
/* ------------------------------------------------------------------------- */
typedef struct {
    double vt, vbs, vgs, vds, vbd, vgd, vgb, von, vdsat, cdrain;
    double gcgs, ceqgs, gcgd, ceqgd, gcgb, ceqgb, capgs, capgd, capgb;
} MosStuff;

static void sargs(double arg, double bot, double side, double *sarg, double *sargsw)
{
    /* mosdefs.h:726-735 */
    if (bot == side) {
        if (bot == .5) *sarg = *sargsw = 1/sqrt(arg);
        else *sarg = *sargsw = exp(-bot*log(arg));
    }
    else {
        if (bot == .5) *sarg = 1/sqrt(arg);
        else *sarg = exp(-bot*log(arg));
        if (side == .5) *sargsw = 1/sqrt(arg);
        else *sargsw = exp(-side*log(arg));
    }
}

static void qmeyer(double vgs, double vgd, double von, double vdsat,
    double *capgs, double *capgd, double *capgb, double phi, double cox)
{
    /* mosload.cc:59-102 */
    double vgst = vgs - von;
    if (vgst <= -phi) {
        *capgb = cox/2; *capgs = 0; *capgd = 0;
    }
    else if (vgst <= -phi/2) {
        *capgb = -vgst*cox/(2*phi); *capgs = 0; *capgd = 0;
    }
    else if (vgst <= 0) {
        *capgb = -vgst*cox/(2*phi);
        *capgs = vgst*cox/(1.5*phi) + cox/3;
        *capgd = 0;
    }
    else {
        double vds = vgs - vgd;
        if (vdsat <= vds) {
            *capgs = cox/3; *capgd = 0; *capgb = 0;
        }
        else {
            double vddif = 2.0*vdsat - vds;
            double vddif1 = vdsat - vds;
            double vddif2 = vddif*vddif;
            *capgd = cox*(1.0 - vdsat*vdsat/vddif2)/3;
            *capgs = cox*(1.0 - vddif1*vddif1/vddif2)/3;
            *capgb = 0;
        }
    }
}

static double mos_eq1(Mos *m, MosStuff *ms)
{
    /* moseq.cc:76-141 */
    const b200_mos_model *md = &m->mod;
    double vbx = ((m->mode == 1) ? ms->vbs : ms->vbd);
    double vgx = ((m->mode == 1) ? ms->vgs : ms->vgd);
    double sarg;
    if (vbx <= 0)
        sarg = sqrt(m->tPhi - vbx);
    else {
        sarg = sqrt(m->tPhi);
        sarg = sarg - vbx/(sarg+sarg);
        sarg = SPMAX(0, sarg);
    }
    if (md->type > 0)
        ms->von = m->tVbi + md->gamma*sarg;
    else
        ms->von = -m->tVbi + md->gamma*sarg;
    double vgst = vgx - ms->von;
    ms->vdsat = SPMAX(vgst, 0);
    double cdrain;
    if (vgst <= 0) {
        cdrain = 0; m->gm = 0; m->gds = 0; m->gmbs = 0;
    }
    else {
        double vds = (m->mode > 0) ? ms->vds : -ms->vds;
        double arg;
        if (sarg <= 0)
            arg = 0;
        else
            arg = md->gamma/(sarg+sarg);
        double betap = m->beta*(1 + md->lambda*vds);
        if (vgst <= vds) {
            m->gm = betap*vgst;
            vgst *= vgst*.5;
            cdrain = betap*vgst;
            m->gds = md->lambda*m->beta*vgst;
            m->gmbs = m->gm*arg;
        }
        else {
            m->gm = betap*vds;
            cdrain = m->gm*(vgst - .5*vds);
            m->gds = betap*(vgst - vds) + md->lambda*m->beta*vds*(vgst - .5*vds);
            m->gmbs = m->gm*arg;
        }
    }
    return cdrain;
}

static double mos_eq6(Mos *m, MosStuff *ms)
{
    /* moseq.cc:1091-1172: n-th power law model */
    const b200_mos_model *md = &m->mod;
    double lvds, lvbs, lvgs;
    if (m->mode > 0) { lvbs = ms->vbs; lvds = ms->vds; lvgs = ms->vgs; }
    else { lvbs = ms->vbd; lvds = -ms->vds; lvgs = ms->vgd; }
    double sarg;
    if (lvbs <= 0)
        sarg = sqrt(m->tPhi - lvbs);
    else {
        sarg = sqrt(m->tPhi);
        sarg = sarg - lvbs/(sarg+sarg);
        sarg = SPMAX(0, sarg);
    }
    ms->von = (m->tVbi*md->type) + md->gamma*sarg - md->gamma1*lvbs - md->sigma*lvds;
    double vgon = lvgs - ms->von;
    double cdrain;
    if (vgon <= 0) {
        ms->vdsat = 0; m->gm = 0; m->gds = 0; m->gmbs = 0; cdrain = 0;
    }
    else {
        double vonbm;
        if (sarg <= 0)
            vonbm = 0;
        else if (lvbs <= 0)
            vonbm = md->gamma1 + md->gamma/(sarg + sarg);
        else
            vonbm = md->gamma1 + .5*md->gamma/sqrt(m->tPhi);
        sarg = log(vgon);
        ms->vdsat = md->kv*exp(sarg*md->nv);
        double idsat = m->beta*exp(sarg*md->nc);
        double lambda = md->lamda0 - md->lamda1*lvbs;
        cdrain = idsat*(1 + lambda*lvds);
        m->gm = cdrain*md->nc/vgon;
        m->gds = m->gm*md->sigma + idsat*lambda;
        m->gmbs = m->gm*vonbm - idsat*md->lamda1*lvds;
        if (ms->vdsat > lvds) {
            double vdst = lvds/ms->vdsat;
            double vdst2 = (2 - vdst)*vdst;
            double vdstg = -vdst*md->nv/vgon;
            double ivdst1 = cdrain*(2 - vdst - vdst);
            cdrain = cdrain*vdst2;
            m->gm = m->gm*vdst2 + ivdst1*vdstg;
            m->gds = m->gds*vdst2 + ivdst1*(1/ms->vdsat + vdstg*md->sigma);
            m->gmbs = m->gmbs*vdst2 + ivdst1*vdstg*vonbm;
        }
    }
    return cdrain;
}

#define P66 .66666666666667
#define P33 .33333333333333
#define P1X3 1.3333333333333

static void mos2_vsat(double *vsat, double *baum)
{
    /* moseq.cc:602-688: Baum's scattering-limited velocity saturation (quartic) */
    double a1 = baum[0], b1 = baum[1], c1 = baum[2], d1 = baum[3];
    double tmp = a1*c1 - 4.0*d1;
    double tmp1 = -P33*b1*b1 + tmp;
    double s = -2.0*b1*b1*b1/27.0 + P33*b1*tmp;
    tmp = -d1*(a1*a1 - 4.0*b1) - c1*c1;
    s += tmp;
    double s2 = s*s;
    tmp = P33*tmp1;
    tmp *= tmp*tmp;
    double p = .25*s2 + tmp;
    double p0 = fabs(p);
    double p2 = sqrt(p0);
    double y3;
    if (p < 0) {
        tmp = sqrt(.25*s2 + p0);
        tmp = exp(P33*log(tmp));
        y3 = 2.0*tmp*cos(P33*atan(-2.0*p2/s)) + P33*b1;
    }
    else {
        tmp = (-.5*s + p2);
        tmp = exp(P33*log(fabs(tmp)));
        tmp1 = (-.5*s - p2);
        tmp1 = exp(P33*log(fabs(tmp1)));
        y3 = tmp + tmp1 + P33*b1;
    }
    int jknt = 0;
    double a3 = sqrt(.25*a1*a1 - b1 + y3);
    double b3 = sqrt(.25*y3*y3 - d1);
    double b00 = .5*y3 + b3, b01 = .5*y3 - b3;
    double a00 = .5*a1 + a3, a01 = .5*a1 - a3;
    double a02 = .25*a00*a00, a03 = .25*a01*a01;
    double a04 = -.5*a00, a05 = -.5*a01;
    double az[4] = { a02 - b00, a03 - b00, a02 - b01, a03 - b01 };
    double bz[4] = { a04, a05, a04, a05 };
    for (int i = 0; i < 4; i++) {
        tmp = az[i];
        if (tmp >= 0) {
            tmp = sqrt(tmp);
            double ptmp = bz[i] + tmp;
            if (ptmp > 0) {
                double poly4 = d1 + ptmp*(c1 + ptmp*(b1 + ptmp*(a1 + ptmp)));
                if (fabs(poly4) <= 1.0e-6) {
                    if (!jknt || ptmp <= tmp1) tmp1 = ptmp;
                    jknt++;
                }
            }
            ptmp = bz[i] - tmp;
            if (ptmp > 0) {
                double poly4 = d1 + ptmp*(c1 + ptmp*(b1 + ptmp*(a1 + ptmp)));
                if (fabs(poly4) <= 1.0e-6) {
                    if (!jknt || ptmp <= tmp1) tmp1 = ptmp;
                    jknt++;
                }
            }
        }
    }
    if (jknt > 0)
        *vsat = tmp1*tmp1 - baum[4];
}

static double mos_eq2(Mos *m, MosStuff *ms)
{
    /* moseq.cc:147-597: level 2 (Grove-Frohman / Meyer analytical model) */
    const b200_mos_model *md = &m->mod;
    double lvbs, lvds, lvgs;
    if (m->mode > 0) { lvbs = ms->vbs; lvds = ms->vds; lvgs = ms->vgs; }
    else { lvbs = ms->vbd; lvds = -ms->vds; lvgs = ms->vgd; }
    const double xl = m->effectiveLength;
    const double nsub = m->nsub;                 /* MOSsubstrateDoping, 0 if not given */
    double xlamda = md->lambda;
    double phiMinVbs = m->tPhi - lvbs;
    double oneoverl = 1.0/xl;
    double sarg, dsrgdb, d2sdb2;
    double sphi = sqrt(m->tPhi);
    double sphi3 = m->tPhi*sphi;
    if (lvbs <= 0.0) {
        sarg = sqrt(phiMinVbs);
        dsrgdb = -0.5/sarg;
        d2sdb2 = 0.5*dsrgdb/phiMinVbs;
    }
    else {
        sarg = sphi/(1.0 + 0.5*lvbs/m->tPhi);
        double tmp = sarg/sphi3;
        dsrgdb = -0.5*sarg*tmp;
        d2sdb2 = -dsrgdb*tmp;
    }
    double barg, dbrgdb, d2bdb2;
    if ((lvds - lvbs) >= 0) {
        barg = sqrt(phiMinVbs + lvds);
        dbrgdb = -0.5/barg;
        d2bdb2 = 0.5*dbrgdb/(phiMinVbs + lvds);
    }
    else {
        barg = sphi/(1.0 + 0.5*(lvbs - lvds)/m->tPhi);
        double tmp = barg/sphi3;
        dbrgdb = -0.5*barg*tmp;
        d2bdb2 = -dbrgdb*tmp;
    }
    int fssandcap = (md->nfs != 0.0 && m->oxideCap != 0.0);
    double factor = 0.125*m->narrowFactor*2.0*M_PI*EPSSIL/m->oxideCap*xl;
    double eta = 1.0 + factor;
    double oneovere = 1.0/eta;
    double vbin = m->tVbi*md->type + factor*phiMinVbs;
    double gamasd = md->gamma, gammad = md->gamma;
    double dgddvb = 0.0, dgdvds = 0.0, dgddb2 = 0.0;
    if ((md->gamma > 0.0) || (nsub > 0.0)) {
        if (md->xj > 0) {
            double tbxwd = m->xd*dbrgdb;
            double tbxws = m->xd*dsrgdb;
            double tmp = 2.0/md->xj;
            double targxs = 1.0 + tmp*m->xd*sarg;
            double targxd = 1.0 + tmp*m->xd*barg;
            double targs = sqrt(targxs);
            double targd = sqrt(targxd);
            tmp = .5*md->xj*oneoverl;
            gamasd *= (1.0 - tmp*(targs + targd - 2.0));
            targs = oneoverl/targs;
            targd = oneoverl/targd;
            dgddvb = -.5*md->gamma*(tbxws*targs + tbxwd*targd);
            dgdvds = md->gamma*0.5*tbxwd*targd;
            tmp = -m->xd*(d2sdb2 + dsrgdb*dsrgdb*m->xd/(md->xj*targxs))*targs;
            double tmp1 = -m->xd*(d2bdb2 + dbrgdb*dbrgdb*m->xd/(md->xj*targxd))*targd;
            dgddb2 = -0.5*md->gamma*(tmp + tmp1);
        }
    }
    ms->von = vbin + gamasd*sarg;
    double vth = ms->von;
    ms->vdsat = 0.0;
    double argg;
    double xn = 1.0;
    if (fssandcap) {
        double tmp = CHARGE*md->nfs*CM2PM2;
        double tmp1 = -(gamasd*dsrgdb + dgddvb*sarg) + factor;
        xn += tmp/m->oxideCap*m->w*xl + tmp1;
        tmp = ms->vt*xn;
        ms->von += tmp;
        argg = 1.0/tmp;
    }
    else {
        argg = 0;
        if (lvgs <= vbin)
            m->gds = 0.0;
    }
    double cdrain = 0;
    if (fssandcap || (lvgs > vbin)) {
        double vgst = lvgs - ms->von;
        double sarg3 = sarg*sarg*sarg;
        double sbiarg = sqrt(m->tBulkPot);
        gammad = gamasd;
        double dgdvbs = dgddvb;
        double body = barg*barg*barg - sarg3;
        double gdbdv = 2.0*gammad*(barg*barg*dbrgdb - sarg*sarg*dsrgdb);
        double dodvbs = -factor + dgdvbs*sarg + gammad*dsrgdb;
        double dxndvb, dxndvd, dodvds;
        if (fssandcap) {
            dxndvb = 2.0*dgdvbs*dsrgdb + gammad*d2sdb2 + dgddb2*sarg;
            dodvbs = dodvbs + ms->vt*dxndvb;
            dxndvd = dgdvds*dsrgdb;
            dodvds = dgdvds*sarg + ms->vt*dxndvd;
        }
        else {
            dxndvb = 0; dxndvd = 0; dodvds = 0;
        }
        double ufact, ueff, dudvgs, dudvds, dudvbs;
        if (m->oxideCap <= 0.0) {
            ufact = 1.0;
            ueff = m->surfaceMobility*M2PCM2;
            dudvgs = 0.0; dudvds = 0.0; dudvbs = 0.0;
        }
        else {
            double tmp = md->ucrit*100*EPSSIL/m->oxideCapFactor;
            if (vgst <= tmp) {
                ufact = 1.0;
                ueff = m->surfaceMobility*M2PCM2;
                dudvgs = 0.0; dudvds = 0.0; dudvbs = 0.0;
            }
            else {
                ufact = exp(md->uexp*log(tmp/vgst));
                ueff = m->surfaceMobility*M2PCM2*ufact;
                dudvgs = -ufact*md->uexp/vgst;
                dudvds = 0.0;
                dudvbs = md->uexp*ufact*dodvbs/vgst;
            }
        }
        gammad = gamasd*oneovere;
        dgdvbs = dgddvb;
        double vgsx = lvgs;
        if (fssandcap)
            vgsx = SPMAX(lvgs, ms->von);
        double dsdvgs, dsdvbs;
        if (gammad > 0) {
            double tmp = gammad*gammad;
            double tmp1 = (vgsx - vbin)*oneovere + phiMinVbs;
            if (tmp1 <= 0.0) {
                ms->vdsat = 0.0;
                dsdvgs = 0.0;
                dsdvbs = 0.0;
            }
            else {
                double arg = sqrt(1.0 + 4.0*tmp1/tmp);
                ms->vdsat = (vgsx - vbin)*oneovere + .5*tmp*(1.0 - arg);
                ms->vdsat = SPMAX(ms->vdsat, 0.0);
                dsdvgs = (1.0 - 1.0/arg)*oneovere;
                dsdvbs = (gammad*(1.0 - arg) + 2.0*tmp1/(gammad*arg))*oneovere*dgdvbs +
                    1.0/arg + factor*dsdvgs;
            }
        }
        else {
            ms->vdsat = (vgsx - vbin)*oneovere;
            ms->vdsat = SPMAX(ms->vdsat, 0.0);
            dsdvgs = 1.0;
            dsdvbs = 0.0;
        }
        if (md->vmax > 0) {
            double tmp = md->vmax*xl/ueff;
            double tmp1 = (vgsx - vbin)*oneovere + phiMinVbs;
            double param[5];
            param[0] = P1X3*gammad;
            param[1] = -2.0*(tmp1 + tmp);
            param[2] = -2.0*gammad*tmp;
            param[3] = 2.0*tmp1*(phiMinVbs + tmp) - phiMinVbs*phiMinVbs - P1X3*gammad*sarg3;
            param[4] = phiMinVbs;
            mos2_vsat(&ms->vdsat, param);
        }
        double dldvgs = 0.0, dldvds = 0.0, dldvbs = 0.0;
        double bsarg, bodys, gdbdvs, dbsrdb = 0;
        if (lvds == 0.0) {
            bsarg = 0; bodys = 0; gdbdvs = 0;
        }
        else {
            gammad = gamasd;
            if ((lvbs - ms->vdsat) <= 0) {
                bsarg = sqrt(ms->vdsat + phiMinVbs);
                dbsrdb = -0.5/bsarg;
            }
            else {
                bsarg = sphi/(1.0 + 0.5*(lvbs - ms->vdsat)/m->tPhi);
                dbsrdb = -0.5*bsarg*bsarg/sphi3;
            }
            bodys = bsarg*bsarg*bsarg - sarg3;
            gdbdvs = 2.0*gammad*(bsarg*bsarg*dbsrdb - sarg*sarg*dsrgdb);
            if (md->vmax <= 0) {
                if (nsub != 0.0 && xlamda <= 0.0) {
                    double tmp = .25*(lvds - ms->vdsat);
                    double tmp1 = sqrt(1.0 + tmp*tmp);
                    tmp = sqrt(tmp + tmp1);
                    xlamda = m->xd*tmp*oneoverl/lvds;
                    tmp = lvds*xlamda/(8.0*tmp1);
                    dldvgs = tmp*dsdvgs;
                    dldvds = -xlamda + tmp;
                    dldvbs = tmp*dsdvbs;
                }
            }
            else {
                double txdv = m->xd/sqrt(md->neff);
                double txlv = md->vmax*txdv/(2.0*ueff);
                double tqdsat = -1.0 + gammad*dbsrdb;
                double tmp = md->vmax*xl;
                double tmp1 = (vgsx - vbin)*oneovere - ms->vdsat - gammad*bsarg;
                tmp1 = 1.0/(tmp*tqdsat - ueff*tmp1);
                dsdvgs = -(tmp - ueff*ms->vdsat)*oneovere*tmp1;
                dsdvbs = -(-tmp*(1.0 + tqdsat - factor*oneovere) +
                    ueff*(gdbdvs - P66*dgdvbs*bodys)*oneovere)*tmp1;
                if (nsub != 0.0 && xlamda <= 0.0) {
                    tmp = lvds - ms->vdsat;
                    tmp = SPMAX(tmp, 0.0);
                    tmp1 = sqrt(txlv*txlv + tmp);
                    tmp = txdv*oneoverl/(2.0*tmp1);
                    xlamda = txdv*(tmp1 - txlv)*oneoverl/lvds;
                    dldvgs = tmp*dsdvgs;
                    dldvds = -xlamda + tmp;
                    dldvbs = tmp*dsdvbs;
                }
            }
        }
        double xwb = m->xd*sbiarg;
        double clfact = 1.0 - xlamda*lvds;
        dldvds = -xlamda - dldvds;
        double tmp1 = xl*clfact;
        double tmp = xlamda*lvds*xl;
        if (nsub == 0.0) xwb = 0.25e-6;
        if (tmp1 < xwb) {
            tmp1 = xwb/(1.0 + (tmp - xl - xwb)/xwb);
            clfact = tmp1*oneoverl;
            double dfact = tmp1/xwb;
            dfact *= dfact;
            dldvgs *= dfact;
            dldvds *= dfact;
            dldvbs *= dfact;
        }
        clfact = 1.0/clfact;
        double beta1 = m->beta*ufact*clfact;
        ufact = 1.0/ufact;
        gammad = gamasd;
        dgdvbs = dgddvb;
        if (lvds > 1.0e-10) {
            if (lvgs > ms->von) {
                if (lvds <= ms->vdsat) {
                    cdrain = beta1*((lvgs - vbin - .5*eta*lvds)*lvds - P66*gammad*body);
                    tmp = cdrain*(dudvgs*ufact - dldvgs*clfact);
                    m->gm = tmp + beta1*lvds;
                    tmp = cdrain*(dudvds*ufact - dldvds*clfact);
                    m->gds = tmp + beta1*(lvgs - vbin - eta*lvds - gammad*barg - P66*dgdvds*body);
                    tmp = cdrain*(dudvbs*ufact - dldvbs*clfact);
                    m->gmbs = tmp - beta1*(gdbdv + P66*dgdvbs*body - factor*lvds);
                }
                else {
                    cdrain = beta1*((lvgs - vbin - eta*.5*ms->vdsat)*ms->vdsat - P66*gammad*bodys);
                    tmp = cdrain*(dudvgs*ufact - dldvgs*clfact);
                    m->gm = tmp + beta1*ms->vdsat +
                        beta1*(lvgs - vbin - eta*ms->vdsat - gammad*bsarg)*dsdvgs;
                    m->gds = -cdrain*dldvds*clfact - P66*beta1*dgdvds*bodys;
                    tmp = cdrain*(dudvbs*ufact - dldvbs*clfact);
                    m->gmbs = tmp - beta1*(gdbdvs + P66*dgdvbs*bodys - factor*ms->vdsat) +
                        beta1*(lvgs - vbin - eta*ms->vdsat - gammad*bsarg)*dsdvbs;
                }
            }
            else {
                if (ms->vdsat <= 0) {
                    m->gds = 0.0;
                    if (lvgs > vth)
                        return cdrain;
                    m->gm = 0.0;
                    m->gmbs = 0.0;
                    return 0.0;
                }
                double vdson = SPMIN(ms->vdsat, lvds);
                if (lvds > ms->vdsat) {
                    barg = bsarg;
                    dbrgdb = dbsrdb;
                    body = bodys;
                    gdbdv = gdbdvs;
                }
                double cdson = beta1*((ms->von - vbin - .5*eta*vdson)*vdson - P66*gammad*body);
                double didvds = beta1*(ms->von - vbin - eta*vdson - gammad*barg);
                double gdson = -cdson*dldvds*clfact - P66*beta1*dgdvds*body;
                if (lvds < ms->vdsat)
                    gdson += didvds;
                double gbson = -cdson*dldvbs*clfact +
                    beta1*(dodvbs*vdson + factor*vdson - P66*dgdvbs*body - gdbdv);
                if (lvds > ms->vdsat)
                    gbson += didvds*dsdvbs;
                tmp = exp(argg*(lvgs - ms->von));
                cdrain = cdson*tmp;
                double gmw = cdrain*argg;
                m->gm = gmw;
                if (lvds > ms->vdsat) m->gm = gmw + didvds*dsdvgs*tmp;
                tmp1 = gmw*(lvgs - ms->von)/xn;
                m->gds = gdson*tmp - m->gm*dodvds - tmp1*dxndvd;
                m->gmbs = gbson*tmp - m->gm*dodvbs - tmp1*dxndvb;
            }
            return cdrain;
        }
        else {
            if (lvgs <= ms->von) {
                if (!fssandcap)
                    m->gds = 0.0;
                else
                    m->gds = beta1*(ms->von - vbin - gammad*sarg)*exp(argg*(lvgs - ms->von));
            }
            else
                m->gds = beta1*(lvgs - vbin - gammad*sarg);
        }
    }
    m->gm = 0.0;
    m->gmbs = 0.0;
    return 0.0;
}

static double mos_eq3(Mos *m, MosStuff *ms)
{
    /* moseq.cc:693-1083: level 3 (semi-empirical short-channel model) */
    const b200_mos_model *md = &m->mod;
    double lvbs, lvds, lvgs;
    if (m->mode > 0) { lvbs = ms->vbs; lvds = ms->vds; lvgs = ms->vgs; }
    else { lvbs = ms->vbd; lvds = -ms->vds; lvgs = ms->vgd; }
    const double xl = m->effectiveLength;
    ms->vdsat = 0.0;
    double oneoverxl = 1.0/xl;
    double eta = md->eta*8.15e-22/(m->oxideCapFactor*xl*xl*xl);
    double phibs, sqphbs, dsqdvb;
    if (lvbs <= 0.0) {
        phibs = m->tPhi - lvbs;
        sqphbs = sqrt(phibs);
        dsqdvb = -0.5/sqphbs;
    }
    else {
        double sqphis = sqrt(m->tPhi);
        double sqphs3 = m->tPhi*sqphis;
        sqphbs = sqphis/(1.0 + lvbs/(m->tPhi + m->tPhi));
        phibs = sqphbs*sqphbs;
        dsqdvb = -phibs/(sqphs3+sqphs3);
    }
    double fshort, dfsdvb;
    if (md->xj != 0.0 && m->xd != 0.0) {
        double coeff0 = 0.0631353e0, coeff1 = 0.8013292e0, coeff2 = -0.01110777e0;
        double wps = m->xd*sqphbs;
        double oneoverxj = 1.0/md->xj;
        double xjonxl = md->xj*oneoverxl;
        double djonxj = md->ld*oneoverxj;
        double wponxj = wps*oneoverxj;
        double wconxj = coeff0 + coeff1*wponxj + coeff2*wponxj*wponxj;
        double arga = wconxj + djonxj;
        double argc = wponxj/(1.0 + wponxj);
        double argb = sqrt(1.0 - argc*argc);
        fshort = 1.0 - xjonxl*(arga*argb - djonxj);
        double dwpdvb = m->xd*dsqdvb;
        double dadvb = (coeff1 + coeff2*(wponxj+wponxj))*dwpdvb*oneoverxj;
        double dbdvb = -argc*argc*(1.0 - argc)*dwpdvb/(argb*wps);
        dfsdvb = -xjonxl*(dadvb*argb + arga*dbdvb);
    }
    else {
        fshort = 1.0;
        dfsdvb = 0.0;
    }
    double gammas = md->gamma*fshort;
    double fbodys = 0.5*gammas/(sqphbs+sqphbs);
    double fbody = fbodys + m->narrowFactor/m->w;
    double onfbdy = 1.0/(1.0 + fbody);
    double dfbdvb = -fbodys*dsqdvb/sqphbs + fbodys*dfsdvb/fshort;
    double qbonco = gammas*sqphbs + m->narrowFactor*phibs/m->w;
    double dqbdvb = gammas*dsqdvb + md->gamma*dfsdvb*sqphbs - m->narrowFactor/m->w;
    double vth = qbonco + m->tVbi*md->type - eta*lvds;
    double dvtdvd = -eta;
    double dvtdvb = dqbdvb;
    ms->von = vth;
    double cdrain;
    double xn = 1.0, dxndvb, dvodvd, dvodvb;
    if (md->nfs != 0.0) {
        double csonco = CHARGE*md->nfs*CM2PM2*xl*m->w/m->oxideCap;
        double cdonco = qbonco/(phibs + phibs);
        xn += csonco + cdonco;
        ms->von += ms->vt*xn;
        dxndvb = dqbdvb/(phibs + phibs) - qbonco*dsqdvb/(phibs*sqphbs);
        dvodvd = dvtdvd;
        dvodvb = dvtdvb + ms->vt*dxndvb;
    }
    else {
        dxndvb = 0; dvodvd = 0; dvodvb = 0;
        if (lvgs <= ms->von) {
            cdrain = 0.0; m->gm = 0.0; m->gds = 0.0; m->gmbs = 0.0;
            return cdrain;
        }
    }
    double vgsx = SPMAX(lvgs, ms->von);
    double onfg = 1.0 + md->theta*(vgsx - vth);
    double fgate = 1.0/onfg;
    double us = m->tSurfMob*M2PCM2*fgate;
    double dfgdvg = -md->theta*fgate*fgate;
    double dfgdvd = -dfgdvg*dvtdvd;
    double dfgdvb = -dfgdvg*dvtdvb;
    double dvsdvg, dvsdvd, dvsdvb, onvdsc;
    ms->vdsat = (vgsx - vth)*onfbdy;
    if (md->vmax <= 0.0) {
        dvsdvg = onfbdy;
        dvsdvd = -dvsdvg*dvtdvd;
        dvsdvb = -dvsdvg*dvtdvb - ms->vdsat*dfbdvb*onfbdy;
        onvdsc = 0;
    }
    else {
        double vdsc = xl*md->vmax/us;
        onvdsc = 1.0/vdsc;
        double arga = (vgsx - vth)*onfbdy;
        double argb = sqrt(arga*arga + vdsc*vdsc);
        ms->vdsat = arga + vdsc - argb;
        double dvsdga = (1.0 - arga/argb)*onfbdy;
        dvsdvg = dvsdga - (1.0 - vdsc/argb)*vdsc*dfgdvg*onfg;
        dvsdvd = -dvsdvg*dvtdvd;
        dvsdvb = -dvsdvg*dvtdvb - arga*dvsdga*dfbdvb;
    }
    double beta = m->beta;
    double vdsx = SPMIN(lvds, ms->vdsat);
    if (vdsx == 0.0) {
        beta *= fgate;
        cdrain = 0.0;
        m->gm = 0.0;
        m->gds = beta*(vgsx - vth);
        m->gmbs = 0.0;
        if (md->nfs != 0.0 && lvgs < ms->von)
            m->gds *= exp((lvgs - ms->von)/(ms->vt*xn));
        return cdrain;
    }
    double cdo = vgsx - vth - 0.5*(1.0 + fbody)*vdsx;
    double dcodvb = -dvtdvb - 0.5*dfbdvb*vdsx;
    double cdnorm = cdo*vdsx;
    m->gm = vdsx;
    if ((m->mode*ms->vds) > ms->vdsat)
        m->gds = -dvtdvd*vdsx;
    else
        m->gds = vgsx - vth - (1.0 + fbody + dvtdvd)*vdsx;
    m->gmbs = dcodvb*vdsx;
    double cd1 = beta*cdnorm;
    beta *= fgate;
    cdrain = beta*cdnorm;
    m->gm = beta*m->gm + dfgdvg*cd1;
    m->gds = beta*m->gds + dfgdvd*cd1;
    m->gmbs = beta*m->gmbs + dfgdvb*cd1;
    double fdrain, dfddvg, dfddvb, dfddvd;
    if (md->vmax > 0.0) {
        fdrain = 1.0/(1.0 + vdsx*onvdsc);
        double fd2 = fdrain*fdrain;
        double arga = fd2*vdsx*onvdsc*onfg;
        dfddvg = -dfgdvg*arga;
        if (m->mode*ms->vds > ms->vdsat)
            dfddvd = -dfgdvd*arga;
        else
            dfddvd = -dfgdvd*arga - fd2*onvdsc;
        dfddvb = -dfgdvb*arga;
        m->gm = fdrain*m->gm + dfddvg*cdrain;
        m->gds = fdrain*m->gds + dfddvd*cdrain;
        m->gmbs = fdrain*m->gmbs + dfddvb*cdrain;
        cdrain = fdrain*cdrain;
        beta *= fdrain;
    }
    else {
        fdrain = 0; dfddvg = 0; dfddvd = 0; dfddvb = 0;
    }
    double gds0 = 0;
    if (lvds > ms->vdsat && !(md->vmax != 0.0 && m->alpha == 0)) {
        double delxl, dldvd, ddldvg, ddldvd, ddldvb;
        if (md->vmax == 0.0) {
            delxl = sqrt(md->kappa*(lvds - ms->vdsat)*m->alpha);
            dldvd = 0.5*delxl/(lvds - ms->vdsat);
            ddldvg = 0.0;
            ddldvd = -dldvd;
            ddldvb = 0.0;
        }
        else {
            double cdsat = cdrain;
            double gdsat = cdsat*(1.0 - fdrain)*onvdsc;
            gdsat = SPMAX(1.0e-12, gdsat);
            double gdoncd = gdsat/cdsat;
            double gdonfd = gdsat/(1.0 - fdrain);
            double gdonfg = gdsat*onfg;
            double dgdvg = gdoncd*m->gm - gdonfd*dfddvg + gdonfg*dfgdvg;
            double dgdvd = gdoncd*m->gds - gdonfd*dfddvd + gdonfg*dfgdvd;
            double dgdvb = gdoncd*m->gmbs - gdonfd*dfddvb + gdonfg*dfgdvb;
            double emax = md->kappa*cdsat*oneoverxl/gdsat;
            double emoncd = emax/cdsat;
            double emongd = emax/gdsat;
            double demdvg = emoncd*m->gm - emongd*dgdvg;
            double demdvd = emoncd*m->gds - emongd*dgdvd;
            double demdvb = emoncd*m->gmbs - emongd*dgdvb;
            double arga = 0.5*emax*m->alpha;
            double argc = md->kappa*m->alpha;
            double argb = sqrt(arga*arga + argc*(lvds - ms->vdsat));
            delxl = argb - arga;
            dldvd = argc/(argb+argb);
            double dldem = 0.5*(arga/argb - 1.0)*m->alpha;
            ddldvg = dldem*demdvg;
            ddldvd = dldem*demdvd - dldvd;
            ddldvb = dldem*demdvb;
        }
        if (delxl > (0.5*xl)) {
            delxl = xl - (xl*xl/(4.0*delxl));
            double arga = 4.0*(xl - delxl)*(xl - delxl)/(xl*xl);
            ddldvg = ddldvg*arga;
            ddldvd = ddldvd*arga;
            ddldvb = ddldvb*arga;
            dldvd = dldvd*arga;
        }
        double dlonxl = delxl*oneoverxl;
        double xlfact = 1.0/(1.0 - dlonxl);
        cd1 = cdrain;
        cdrain = cdrain*xlfact;
        double diddl = cdrain/(xl - delxl);
        m->gm = m->gm*xlfact + diddl*ddldvg;
        m->gmbs = m->gmbs*xlfact + diddl*ddldvb;
        gds0 = diddl*ddldvd;
        m->gm = m->gm + gds0*dvsdvg;
        m->gmbs = m->gmbs + gds0*dvsdvb;
        m->gds = m->gds*xlfact + diddl*dldvd + gds0*dvsdvd;
    }
    if (lvgs < ms->von) {
        double onxn = 1.0/xn;
        double ondvt = onxn/ms->vt;
        double wfact = exp((lvgs - ms->von)*ondvt);
        cdrain = cdrain*wfact;
        double gms = m->gm*wfact;
        double gmw = cdrain*ondvt;
        m->gm = gmw;
        if (lvds > ms->vdsat)
            m->gm = m->gm + gds0*dvsdvg*wfact;
        m->gds = m->gds*wfact + (gms - gmw)*dvodvd;
        m->gmbs = m->gmbs*wfact + (gms - gmw)*dvodvb - gmw*(lvgs - ms->von)*onxn*dxndvb;
    }
    (void)cd1;
    return cdrain;
}

static int mos_limiting(wo_ckt *c, Mos *m, MosStuff *ms)
{
    /* mosload.cc:346-384 */
    double *s0 = c->states[0] + m->st;
    if (s0[MS_vds] >= 0) {
        ms->vgs = fetlim(ms->vgs, s0[MS_vgs], ms->von);
        ms->vds = ms->vgs - ms->vgd;
        ms->vds = limvds(ms->vds, s0[MS_vds]);
        ms->vgd = ms->vgs - ms->vds;
    }
    else {
        double vgdo = s0[MS_vgs] - s0[MS_vds];
        ms->vgd = fetlim(ms->vgd, vgdo, ms->von);
        ms->vds = ms->vgs - ms->vgd;
        if (!c->opt.oldlimit)
            ms->vds = -limvds(-ms->vds, -(s0[MS_vds]));
        ms->vgs = ms->vgd + ms->vds;
    }
    int check;
    if (ms->vds >= 0) {
        ms->vbs = pnjlim(ms->vbs, s0[MS_vbs], ms->vt, m->sourceVcrit, &check);
        ms->vbd = ms->vbs - ms->vds;
    }
    else {
        ms->vbd = pnjlim(ms->vbd, s0[MS_vbd], ms->vt, m->drainVcrit, &check);
        ms->vbs = ms->vbd + ms->vds;
    }
    ms->vgb = ms->vgs - ms->vbs;
    return check;
}

static void mos_iv(wo_ckt *c, Mos *m, MosStuff *ms)
{
    /* mosload.cc:498-593 */
    ms->gcgs = 0; ms->ceqgs = 0; ms->gcgd = 0; ms->ceqgd = 0; ms->gcgb = 0; ms->ceqgb = 0;
    double gmin = c->opt.gmin;
    if (ms->vbs <= -3*ms->vt) {
        m->gbs = gmin;
        m->cbs = m->gbs*ms->vbs - m->tSourceSatCur;
    }
    else {
        double evbs = exp(ms->vbs/ms->vt);
        m->gbs = m->tSourceSatCur*evbs/ms->vt + gmin;
        m->cbs = m->tSourceSatCur*(evbs-1) + gmin*ms->vbs;
    }
    if (ms->vbd <= -3*ms->vt) {
        m->gbd = gmin;
        m->cbd = m->gbd*ms->vbd - m->tDrainSatCur;
    }
    else {
        double evbd = exp(ms->vbd/ms->vt);
        m->gbd = m->tDrainSatCur*evbd/ms->vt + gmin;
        m->cbd = m->tDrainSatCur*(evbd-1) + gmin*ms->vbd;
    }
    m->mode = (ms->vds >= 0) ? 1 : -1;
    /* mosload.cc:570-577 */
    ms->cdrain = (m->mod.level == 6) ? mos_eq6(m, ms) : (m->mod.level == 3) ? mos_eq3(m, ms) :
        (m->mod.level == 2) ? mos_eq2(m, ms) : mos_eq1(m, ms);
    if (m->mod.type > 0) {
        m->von = ms->von; m->vdsat = ms->vdsat;
    }
    else {
        m->von = -ms->von; m->vdsat = -ms->vdsat;
    }
    if (m->mode > 0)
        m->cd = ms->cdrain - m->cbd;
    else
        m->cd = -ms->cdrain - m->cbd;
}

static void mos_cap(wo_ckt *c, Mos *m, MosStuff *ms)
{
    /* mosload.cc:597-750 */
    const b200_mos_model *md = &m->mod;
    double *s0 = c->states[0] + m->st, *s1 = c->states[1] + m->st;
    if (m->Cbs != 0 || m->Cbssw != 0) {
        if (ms->vbs < m->tDepCap) {
            double arg = 1 - ms->vbs/m->tBulkPot, sarg, sargsw;
            sargs(arg, md->mj, md->mjsw, &sarg, &sargsw);
            s0[MS_qbs] = m->tBulkPot*(m->Cbs*(1 - arg*sarg)/(1 - md->mj)
                + m->Cbssw*(1 - arg*sargsw)/(1 - md->mjsw));
            m->capbs = m->Cbs*sarg + m->Cbssw*sargsw;
        }
        else {
            s0[MS_qbs] = m->f4s + ms->vbs*(m->f2s + .5*ms->vbs*m->f3s);
            m->capbs = m->f2s + m->f3s*ms->vbs;
        }
    }
    else {
        s0[MS_qbs] = 0; m->capbs = 0;
    }
    if (m->Cbd != 0 || m->Cbdsw != 0) {
        if (ms->vbd < m->tDepCap) {
            double arg = 1 - ms->vbd/m->tBulkPot, sarg, sargsw;
            sargs(arg, md->mj, md->mjsw, &sarg, &sargsw);
            s0[MS_qbd] = m->tBulkPot*(m->Cbd*(1 - arg*sarg)/(1 - md->mj)
                + m->Cbdsw*(1 - arg*sargsw)/(1 - md->mjsw));
            m->capbd = m->Cbd*sarg + m->Cbdsw*sargsw;
        }
        else {
            s0[MS_qbd] = m->f4d + ms->vbd*(m->f2d + .5*ms->vbd*m->f3d);
            m->capbd = m->f2d + ms->vbd*m->f3d;
        }
    }
    else {
        s0[MS_qbd] = 0; m->capbd = 0;
    }

    double vgx, vgy, *px, *py;
    if (m->mode > 0) {
        vgx = ms->vgs; vgy = ms->vgd; px = &s0[MS_capgs]; py = &s0[MS_capgd];
    }
    else {
        vgx = ms->vgd; vgy = ms->vgs; px = &s0[MS_capgd]; py = &s0[MS_capgs];
    }
    qmeyer(vgx, vgy, ms->von, ms->vdsat, px, py, &s0[MS_capgb], m->tPhi, m->oxideCap);

    if (c->mode & (MODEINITTRAN|MODEINITSMSIG)) {
        s1[MS_capgs] = s0[MS_capgs];
        s1[MS_capgd] = s0[MS_capgd];
        s1[MS_capgb] = s0[MS_capgb];
        ms->capgs = s0[MS_capgs] + s0[MS_capgs] + m->gateSourceOverlapCap;
        ms->capgd = s0[MS_capgd] + s0[MS_capgd] + m->gateDrainOverlapCap;
        ms->capgb = s0[MS_capgb] + s0[MS_capgb] + m->gateBulkOverlapCap;
        s0[MS_qgs] = ms->capgs*ms->vgs;
        s0[MS_qgd] = ms->capgd*ms->vgd;
        s0[MS_qgb] = ms->capgb*ms->vgb;
        s1[MS_qbd] = s0[MS_qbd];
        s1[MS_qbs] = s0[MS_qbs];
        s1[MS_qgs] = s0[MS_qgs];
        s1[MS_qgd] = s0[MS_qgd];
        s1[MS_qgb] = s0[MS_qgb];
    }
    else {
        ms->capgs = s0[MS_capgs] + s1[MS_capgs] + m->gateSourceOverlapCap;
        ms->capgd = s0[MS_capgd] + s1[MS_capgd] + m->gateDrainOverlapCap;
        ms->capgb = s0[MS_capgb] + s1[MS_capgb] + m->gateBulkOverlapCap;
        vgx = s1[MS_vgs];
        vgy = vgx - s1[MS_vds];
        double vxy = vgx - s1[MS_vbs];
        s0[MS_qgs] = (ms->vgs - vgx)*ms->capgs + s1[MS_qgs];
        s0[MS_qgd] = (ms->vgd - vgy)*ms->capgd + s1[MS_qgd];
        s0[MS_qgb] = (ms->vgb - vxy)*ms->capgb + s1[MS_qgb];
    }
}

static void mos_meyer_integ(wo_ckt *c, Mos *m, MosStuff *ms)
{
    /* mosload.cc:477-486 and :784-793 */
    double *s0 = c->states[0] + m->st;
    integrate4(c, &ms->gcgs, &ms->ceqgs, ms->capgs, m->st + MS_qgs);
    integrate4(c, &ms->gcgd, &ms->ceqgd, ms->capgd, m->st + MS_qgd);
    integrate4(c, &ms->gcgb, &ms->ceqgb, ms->capgb, m->st + MS_qgb);
    ms->ceqgs += c->ag[0]*s0[MS_qgs] - ms->gcgs*ms->vgs;
    ms->ceqgd += c->ag[0]*s0[MS_qgd] - ms->gcgd*ms->vgd;
    ms->ceqgb += c->ag[0]*s0[MS_qgb] - ms->gcgb*ms->vgb;
}

static void mos_integ(wo_ckt *c, Mos *m, MosStuff *ms)
{
    /* mosload.cc:754-796 */
    double *s0 = c->states[0] + m->st;
    double geq, ceq;
    integrate4(c, &geq, &ceq, m->capbd, m->st + MS_qbd);
    m->gbd += geq;
    m->cbd += s0[MS_cqbd];
    m->cd  -= s0[MS_cqbd];
    integrate4(c, &geq, &ceq, m->capbs, m->st + MS_qbs);
    m->gbs += geq;
    m->cbs += s0[MS_cqbs];
    if (ms->capgs == 0) s0[MS_cqgs] = 0;
    if (ms->capgd == 0) s0[MS_cqgd] = 0;
    if (ms->capgb == 0) s0[MS_cqgb] = 0;
    mos_meyer_integ(c, m, ms);
}

static void mos_stamp(wo_ckt *c, Mos *m, MosStuff *ms, int tran)
{
    /* mosload.cc:801-924 (tran) and :928-1041 (dc) */
    double *s0 = c->states[0] + m->st;
    const int *p = m->ptr;
#define MSC(xx) (m->mult*(xx))
    s0[MS_vbs] = ms->vbs;
    s0[MS_vbd] = ms->vbd;
    s0[MS_vgs] = ms->vgs;
    s0[MS_vds] = ms->vds;
    double gbd = MSC(m->gbd), gbs = MSC(m->gbs), gds = MSC(m->gds), gm = MSC(m->gm);
    double gmbs = MSC(m->gmbs), cbd = MSC(m->cbd), cbs = MSC(m->cbs);
    double ceqbs = cbs - gbs*ms->vbs;
    double ceqbd = cbd - gbd*ms->vbd;
    double cdreq;
    if (m->mode > 0)
        cdreq = MSC(ms->cdrain) - gds*ms->vds - gm*ms->vgs - gmbs*ms->vbs;
    else
        cdreq = -(MSC(ms->cdrain) + gds*ms->vds - gm*ms->vgd - gmbs*ms->vbd);
    double dcon = MSC(m->drainConductance), scon = MSC(m->sourceConductance);
    if (tran) {
        if (m->mod.type > 0) {
            rhsadd(c, m->g, -MSC(ms->ceqgs + ms->ceqgb + ms->ceqgd));
            rhsadd(c, m->b, -(ceqbs + ceqbd - MSC(ms->ceqgb)));
            rhsadd(c, m->dp, ceqbd - cdreq + MSC(ms->ceqgd));
            rhsadd(c, m->sp, cdreq + ceqbs + MSC(ms->ceqgs));
        }
        else {
            rhsadd(c, m->g, MSC(ms->ceqgs + ms->ceqgb + ms->ceqgd));
            rhsadd(c, m->b, ceqbs + ceqbd - MSC(ms->ceqgb));
            rhsadd(c, m->dp, -(ceqbd - cdreq + MSC(ms->ceqgd)));
            rhsadd(c, m->sp, -(cdreq + ceqbs + MSC(ms->ceqgs)));
        }
        double gcgd = MSC(ms->gcgd), gcgs = MSC(ms->gcgs), gcgb = MSC(ms->gcgb);
        ldadd(c, p[M_Dd], dcon);
        ldadd(c, p[M_Gg], gcgd + gcgs + gcgb);
        ldadd(c, p[M_Ss], scon);
        ldadd(c, p[M_Bb], gbd + gbs + gcgb);
        ldadd(c, p[M_Ddp], -dcon);
        ldadd(c, p[M_Gb], -gcgb);
        ldadd(c, p[M_Gdp], -gcgd);
        ldadd(c, p[M_Gsp], -gcgs);
        ldadd(c, p[M_Ssp], -scon);
        ldadd(c, p[M_Bg], -gcgb);
        ldadd(c, p[M_Bdp], -gbd);
        ldadd(c, p[M_Bsp], -gbs);
        ldadd(c, p[M_DPd], -dcon);
        ldadd(c, p[M_SPs], -scon);
        if (m->mode > 0) {
            ldadd(c, p[M_DPdp], dcon + gds + gbd + gcgd);
            ldadd(c, p[M_SPsp], scon + gds + gbs + gm + gmbs + gcgs);
            ldadd(c, p[M_DPg], gm - gcgd);
            ldadd(c, p[M_DPb], -gbd + gmbs);
            ldadd(c, p[M_DPsp], -(gds + gm + gmbs));
            ldadd(c, p[M_SPg], -(gm + gcgs));
            ldadd(c, p[M_SPb], -(gbs + gmbs));
            ldadd(c, p[M_SPdp], -gds);
        }
        else {
            ldadd(c, p[M_DPdp], dcon + gds + gbd + gm + gmbs + gcgd);
            ldadd(c, p[M_SPsp], scon + gds + gbs + gcgs);
            ldadd(c, p[M_DPg], -(gm + gcgd));
            ldadd(c, p[M_DPb], -(gbd + gmbs));
            ldadd(c, p[M_DPsp], -gds);
            ldadd(c, p[M_SPg], -(-gm + gcgs));
            ldadd(c, p[M_SPb], -(gbs - gmbs));
            ldadd(c, p[M_SPdp], -(gds + gm + gmbs));
        }
    }
    else {
        if (m->mod.type > 0) {
            rhsadd(c, m->b, -(ceqbs + ceqbd));
            rhsadd(c, m->dp, ceqbd - cdreq);
            rhsadd(c, m->sp, cdreq + ceqbs);
        }
        else {
            rhsadd(c, m->b, ceqbs + ceqbd);
            rhsadd(c, m->dp, -(ceqbd - cdreq));
            rhsadd(c, m->sp, -(cdreq + ceqbs));
        }
        ldadd(c, p[M_Dd], dcon);
        ldadd(c, p[M_Ss], scon);
        ldadd(c, p[M_Bb], gbd + gbs);
        ldadd(c, p[M_Ddp], -dcon);
        ldadd(c, p[M_Ssp], -scon);
        ldadd(c, p[M_Bdp], -gbd);
        ldadd(c, p[M_Bsp], -gbs);
        ldadd(c, p[M_DPd], -dcon);
        ldadd(c, p[M_SPs], -scon);
        if (m->mode > 0) {
            ldadd(c, p[M_DPdp], dcon + gds + gbd);
            ldadd(c, p[M_SPsp], scon + gds + gbs + gm + gmbs);
            ldadd(c, p[M_DPg], gm);
            ldadd(c, p[M_DPb], -gbd + gmbs);
            ldadd(c, p[M_DPsp], -(gds + gm + gmbs));
            ldadd(c, p[M_SPg], -gm);
            ldadd(c, p[M_SPb], -(gbs + gmbs));
            ldadd(c, p[M_SPdp], -gds);
        }
        else {
            ldadd(c, p[M_DPdp], dcon + gds + gbd + gm + gmbs);
            ldadd(c, p[M_SPsp], scon + gds + gbs);
            ldadd(c, p[M_DPg], -gm);
            ldadd(c, p[M_DPb], -(gbd + gmbs));
            ldadd(c, p[M_DPsp], -gds);
            ldadd(c, p[M_SPg], gm);
            ldadd(c, p[M_SPb], -(gbs - gmbs));
            ldadd(c, p[M_SPdp], -(gds + gm + gmbs));
        }
    }
#undef MSC
}

static int mos_bypass(wo_ckt *c, Mos *m, MosStuff *ms)
{
    /* mosload.cc:388-494 */
    double *s0 = c->states[0] + m->st, *s1 = c->states[1] + m->st;
    const b200_options *o = &c->opt;
    double delvbs = ms->vbs - s0[MS_vbs];
    double delvgs = ms->vgs - s0[MS_vgs];
    double delvds = ms->vds - s0[MS_vds];
    double delvbd = ms->vbd - s0[MS_vbd];
    double delvgd = ms->vgd - (s0[MS_vgs] - s0[MS_vds]);
    double cbhat = m->cbs + m->cbd + m->gbd*delvbd + m->gbs*delvbs;
    double A1 = fabs(cbhat);
    double A2 = m->cbs + m->cbd;
    A2 = fabs(A2);
    double A3 = SPMAX(A1, A2) + o->abstol;
    A1 = cbhat - (m->cbs + m->cbd);
    if (fabs(A1) >= o->reltol*A3) return 0;
    A1 = fabs(ms->vbs); A2 = fabs(s0[MS_vbs]);
    if (fabs(delvbs) >= (o->reltol*SPMAX(A1, A2) + o->vntol)) return 0;
    A1 = fabs(ms->vbd); A2 = fabs(s0[MS_vbd]);
    if (fabs(delvbd) >= (o->reltol*SPMAX(A1, A2) + o->vntol)) return 0;
    A1 = fabs(ms->vgs); A2 = fabs(s0[MS_vgs]);
    if (fabs(delvgs) >= (o->reltol*SPMAX(A1, A2) + o->vntol)) return 0;
    A1 = fabs(ms->vds); A2 = fabs(s0[MS_vds]);
    if (fabs(delvds) >= (o->reltol*SPMAX(A1, A2) + o->vntol)) return 0;
    double cdhat;
    if (m->mode >= 0)
        cdhat = m->cd - m->gbd*delvbd + m->gmbs*delvbs + m->gm*delvgs + m->gds*delvds;
    else
        cdhat = m->cd + (m->gmbs - m->gbd)*delvbd - m->gm*delvgd + m->gds*delvds;
    A1 = cdhat - m->cd;
    A2 = fabs(cdhat);
    A3 = fabs(m->cd);
    if ((fabs(A1) >= o->reltol*SPMAX(A2, A3) + o->abstol)) return 0;

    ms->vbs = s0[MS_vbs];
    ms->vgs = s0[MS_vgs];
    ms->vds = s0[MS_vds];
    ms->vbd = s0[MS_vbd];
    ms->vgd = ms->vgs - ms->vds;
    ms->vgb = ms->vgs - ms->vbs;
    ms->cdrain = m->mode*(m->cd + m->cbd);
    if (c->mode & MODETRAN) {
        ms->capgs = s0[MS_capgs] + s1[MS_capgs] + m->gateSourceOverlapCap;
        ms->capgd = s0[MS_capgd] + s1[MS_capgd] + m->gateDrainOverlapCap;
        ms->capgb = s0[MS_capgb] + s1[MS_capgb] + m->gateBulkOverlapCap;
        mos_meyer_integ(c, m, ms);
        mos_stamp(c, m, ms, 1);
    }
    else
        mos_stamp(c, m, ms, 0);
    return 1;
}

static void mos_get_v(wo_ckt *c, Mos *m, MosStuff *ms, const double *v)
{
    if (m->mod.type > 0) {
        ms->vbs = v[m->b] - v[m->sp];
        ms->vgs = v[m->g] - v[m->sp];
        ms->vds = v[m->dp] - v[m->sp];
    }
    else {
        ms->vbs = v[m->sp] - v[m->b];
        ms->vgs = v[m->sp] - v[m->g];
        ms->vds = v[m->sp] - v[m->dp];
    }
}

static void mos_load(wo_ckt *c, Mos *m)
{
    /* mosload.cc:107-342 */
    MosStuff ms;
    memset(&ms, 0, sizeof(ms));
    double *s0 = c->states[0] + m->st, *s1 = c->states[1] + m->st, *s2 = c->states[2] + m->st;
    int type = m->mod.type;
    ms.vt = CONSTKoverQ*m->temp;
    if (c->mode & MODEINITFLOAT) {
        mos_get_v(c, m, &ms, c->rhsOld);
        ms.von = (type > 0) ? m->von : -m->von;
        ms.vbd = ms.vbs - ms.vds;
        ms.vgd = ms.vgs - ms.vds;
        ms.vgb = ms.vgs - ms.vbs;
        if (c->opt.bypass && mos_bypass(c, m, &ms))
            return;
        mos_limiting(c, m, &ms);
        mos_iv(c, m, &ms);
        if (c->mode & MODETRAN) {
            mos_cap(c, m, &ms);
            mos_integ(c, m, &ms);
            mos_stamp(c, m, &ms, 1);
        }
        else
            mos_stamp(c, m, &ms, 0);
        return;
    }
    if (c->mode & MODEINITPRED) {
        /* device.h:171-175 DEV.pred */
        ms.vbs = s1[MS_vbs]*c->pred[0] + s2[MS_vbs]*c->pred[1];
        ms.vgs = s1[MS_vgs]*c->pred[0] + s2[MS_vgs]*c->pred[1];
        ms.vds = s1[MS_vds]*c->pred[0] + s2[MS_vds]*c->pred[1];
        ms.vbd = ms.vbs - ms.vds;
        ms.vgd = ms.vgs - ms.vds;
        ms.vgb = ms.vgs - ms.vbs;
        ms.von = (type > 0) ? m->von : -m->von;
        s0[MS_vbs] = s1[MS_vbs];
        s0[MS_vgs] = s1[MS_vgs];
        s0[MS_vds] = s1[MS_vds];
        s0[MS_vbd] = s0[MS_vbs] - s0[MS_vds];
        mos_limiting(c, m, &ms);
        mos_iv(c, m, &ms);
        if (c->mode & MODETRAN) {
            mos_cap(c, m, &ms);
            mos_integ(c, m, &ms);
            mos_stamp(c, m, &ms, 1);
        }
        else
            mos_stamp(c, m, &ms, 0);
        return;
    }
    if (c->mode & MODEINITFIX) {
        ms.von = (type > 0) ? m->von : -m->von;
        if (m->in.off) {
            ms.vbs = ms.vgs = ms.vds = ms.vbd = ms.vgd = ms.vgb = 0;
        }
        else {
            mos_get_v(c, m, &ms, c->rhsOld);
            ms.vbd = ms.vbs - ms.vds;
            ms.vgd = ms.vgs - ms.vds;
            ms.vgb = ms.vgs - ms.vbs;
            if (mos_limiting(c, m, &ms))
                c->noncon++;
        }
        mos_iv(c, m, &ms);
        mos_stamp(c, m, &ms, 0);
        return;
    }
    if (c->mode & MODEINITTRAN) {
        ms.von = (type > 0) ? m->von : -m->von;
        ms.vbs = s1[MS_vbs];
        ms.vgs = s1[MS_vgs];
        ms.vds = s1[MS_vds];
        ms.vbd = ms.vbs - ms.vds;
        ms.vgd = ms.vgs - ms.vds;
        ms.vgb = ms.vgs - ms.vbs;
        mos_iv(c, m, &ms);
        mos_cap(c, m, &ms);
        mos_integ(c, m, &ms);
        mos_stamp(c, m, &ms, 1);
        return;
    }
    if ((c->mode & MODEINITJCT) && (c->mode & MODETRANOP) && (c->mode & MODEUIC)) {
        if (type > 0) {
            ms.vds = m->icVDS; ms.vgs = m->icVGS; ms.vbs = m->icVBS;
        }
        else {
            ms.vds = -m->icVDS; ms.vgs = -m->icVGS; ms.vbs = -m->icVBS;
        }
        s0[MS_vds] = ms.vds;
        s0[MS_vgs] = ms.vgs;
        s0[MS_vbs] = ms.vbs;
        return;
    }
    if (c->mode & MODEINITJCT) {
        ms.von = (type > 0) ? m->von : -m->von;
        if (m->in.off) {
            ms.vbs = ms.vgs = ms.vds = ms.vbd = ms.vgd = ms.vgb = 0;
        }
        else {
            if (m->icVDS == 0 && m->icVGS == 0 && m->icVBS == 0) {
                ms.vbs = -1;
                ms.vgs = (type > 0) ? m->tVto : -m->tVto;
                ms.vds = 0;
            }
            else {
                if (type > 0) {
                    ms.vds = m->icVDS; ms.vgs = m->icVGS; ms.vbs = m->icVBS;
                }
                else {
                    ms.vds = -m->icVDS; ms.vgs = -m->icVGS; ms.vbs = -m->icVBS;
                }
            }
            ms.vbd = ms.vbs - ms.vds;
            ms.vgd = ms.vgs - ms.vds;
            ms.vgb = ms.vgs - ms.vbs;
        }
        mos_iv(c, m, &ms);
        mos_stamp(c, m, &ms, 0);
    }
}

/* MOSdev::convTest per instance: mosconv.cc:52-141.  Returns 1 if not converged. */
static int mos_conv(wo_ckt *c, Mos *m)
{
    if (m->in.off && (c->mode & MODEINITFIX))
        return 0;
    MosStuff ms;
    mos_get_v(c, m, &ms, c->rhs);           /* NOTE: CKTrhs (previous iterate), as in the reference */
    double *s0 = c->states[0] + m->st;
    double vbd = ms.vbs - ms.vds, vgd = ms.vgs - ms.vds;
    double delvbs = ms.vbs - s0[MS_vbs];
    double delvbd = vbd - s0[MS_vbd];
    double delvgs = ms.vgs - s0[MS_vgs];
    double delvds = ms.vds - s0[MS_vds];
    double delvgd = vgd - (s0[MS_vgs] - s0[MS_vds]);
    double cdhat;
    if (m->mode >= 0)
        cdhat = m->cd - m->gbd*delvbd + m->gmbs*delvbs + m->gm*delvgs + m->gds*delvds;
    else
        cdhat = m->cd - (m->gbd - m->gmbs)*delvbd - m->gm*delvgd + m->gds*delvds;
    double A1 = fabs(cdhat), A2 = fabs(m->cd), A3 = cdhat - m->cd;
    A3 = fabs(A3);
    double tol = c->opt.reltol*SPMAX(A1, A2) + c->opt.abstol;
    if (A3 >= tol)
        return 1;
    double cbhat = m->cbs + m->cbd + m->gbd*delvbd + m->gbs*delvbs;
    A1 = fabs(cbhat);
    A2 = m->cbs + m->cbd;
    A2 = fabs(A2);
    A3 = cbhat - (m->cbs + m->cbd);
    A3 = fabs(A3);
    tol = c->opt.reltol*SPMAX(A1, A2) + c->opt.abstol;
    if (A3 > tol)
        return 1;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* CAP / SRC load                                                             */
/* ------------------------------------------------------------------------- */

/* ------------------------------------------------------------------------- */
/* Josephson junction, RSJ level 1: devlib/jj/{jjset,jjload,jjacct,jjic,jjtrun}.cc */
/* ------------------------------------------------------------------------- */

static void jj_setup_instance(JJ *j)
{
    /* jjset.cc:560-612: temperature factor (carried by the deck: depends on temp_k only),
     * vshunt default, vless / vmore */
    const b200_jj_model *m = &j->mod;
    j->vg = GIVEN(j->in.vg) ? j->in.vg : m->vg;
    j->tcf = GIVEN(j->in.tcf) ? j->in.tcf : 1.0;
    j->vshunt = GIVEN(j->in.vshunt) ? j->in.vshunt : m->vshunt;
    if (j->vshunt < 0.0 || j->vshunt > m->vg) j->vshunt = 0.0;
    double halfdv = 0.5*m->delv;
    j->vless = j->vg - halfdv;
    j->vmore = j->vg + halfdv;
    /* :613-635 ics / area */
    if (GIVEN(j->in.ics)) {
        double ics = j->in.ics;
        if (ics < m->criti/50 || ics > m->criti*50) ics = m->criti;
        j->area = ics/m->criti;
    }
    else if (GIVEN(j->in.area)) j->area = j->in.area;
    else j->area = 1;
    j->criti = m->criti*j->tcf*j->area;
    /* :683-705 */
    double sqrta = sqrt(j->area);
    j->cap = m->cap*(j->area*(1.0 - m->cmu) + sqrta*m->cmu);
    double gfac = j->area*(1.0 - m->gmu) + sqrta*m->gmu;
    j->g0 = gfac/m->r0;
    j->gn = gfac/m->rn;
    j->gs = j->criti/(m->icfactor*m->delv);
    j->cr1 = (j->g0 - j->gs)*j->vless;
    j->cr2 = j->criti/m->icfactor + j->vless*j->g0 - j->vmore*j->gn;
    j->gshunt = 0.0;
    if (j->vshunt > 1e-12) {
        double gshunt = j->criti/j->vshunt - j->g0;
        if (gshunt > 0.0) j->gshunt = gshunt;
    }
    /* :746-790 rtype 3 polynomial */
    if (m->rtype == 3) {
        double Gn = (j->criti/m->icfactor + j->vless*j->g0)/j->vmore;
        double temp = j->vmore*j->vmore;
        j->g1 = 0.5*(5.0*Gn - j->gs - 4.0*j->g0);
        j->g2 = (Gn - j->g0 - j->g1)/(temp*temp);
        j->g1 /= temp;
        j->cr1 = (Gn - j->gn)*j->vmore;
    }
    else { j->g1 = 0; j->g2 = 0; }
}

typedef struct { double vj, phi, ci, gqt, crhs, crt, dcrt, pfac, ddv; } JJStuff;

static void jj_limiting(wo_ckt *c, JJ *j, JJStuff *js)
{
    /* jjload.cc:593-617 */
    double vprev = c->states[0][j->st + JS_voltage];
    double absold = vprev < 0 ? -vprev : vprev;
    double vtop = vprev + js->ddv;
    double vbot = vprev - js->ddv;
    if (absold < j->vless) {
        vtop = SPMAX(vtop, j->vless);
        vbot = SPMIN(vbot, -j->vless);
    }
    else if (vprev > j->vmore) {
        vtop += 1.0;
        vbot = SPMIN(vbot, j->vmore);
    }
    else if (vprev < -j->vmore) {
        vtop = SPMAX(vtop, -j->vmore);
        vbot -= 1.0;
    }
    if (js->vj > vtop) js->vj = vtop;
    else if (js->vj < vbot) js->vj = vbot;
}

static void jj_iv(JJ *j, JJStuff *js)
{
    /* jjload.cc:620-739 (rtype 1, 2, 3; 0 and others: no quasiparticle branch) */
    const b200_jj_model *m = &j->mod;
    double vj = js->vj;
    double absvj = vj < 0 ? -vj : vj;
    if (m->rtype == 1) {
        if (absvj < j->vless)
            js->gqt = j->g0;
        else if (absvj < j->vmore) {
            js->gqt = j->gs;
            js->crhs = (vj >= 0 ? j->cr1 : -j->cr1);
        }
        else {
            js->gqt = j->gn;
            js->crhs = (vj >= 0 ? j->cr2 : -j->cr2);
        }
    }
    else if (m->rtype == 2) {
        double dv = 0.5*m->delv;
        double avj = absvj/dv;
        double vgdv = j->vg/dv;
        double gam = avj - vgdv;
        if (gam > 30.0) gam = 30.0;
        else if (gam < -30.0) gam = -30.0;
        double xp = exp(gam);
        if (vgdv > 30.0) vgdv = 30.0;
        double xp0 = exp(-vgdv);
        double A = 0.5*(1.0 - m->icfactor)*j->vg;
        double As = vj < 0.0 ? -A : A;
        double xp1 = 1 + xp - xp0;
        double I = ( (j->g0 + j->gn*xp0*A/dv)*vj + j->gn*(vj - As)*(xp - xp0) ) / xp1;
        double dIdv = (j->g0 + j->gn*xp0*A/dv)*(xp1 - absvj*xp)/(xp1*xp1) +
            j->gn*(xp*(absvj - A)/(dv*xp1*xp1) + (xp - xp0)/xp1) ;
        js->crhs = I - dIdv*vj;
        js->gqt = dIdv;
    }
    else if (m->rtype == 3) {
        if (absvj < j->vmore) {
            double temp1 = vj*vj;
            double temp2 = temp1*temp1;
            js->gqt = j->g0 + 3.0*j->g1*temp1 + 5.0*j->g2*temp2;
            temp1 *= vj;
            temp2 *= vj;
            js->crhs = -2.0*temp1*j->g1 - 4.0*temp2*j->g2;
        }
        else {
            js->gqt = j->gn;
            js->crhs = (vj >= 0 ? j->cr1 : -j->cr1);
        }
    }
    else
        js->gqt = 0;
}

static void jj_stamp(wo_ckt *c, JJ *j, JJStuff *js)
{
    /* jjstuff::jj_load, jjload.cc:790-871 (sin / cos through libm; the reference build uses
     * the x87 fsincos instruction, jjload.cc:55,808 -- same value to the last bit or two) */
    double gqt = js->gqt, crhs = js->crhs, crt = js->crt;
    double *s0 = c->states[0] + j->st;
    s0[JS_voltage] = js->vj;
    s0[JS_phase] = js->phi;
    s0[JS_conI] = js->ci;
    s0[JS_crti] = js->crt;
    if (c->mode & MODEDC) s0[JS_qpi] = 0.0;
    else s0[JS_qpi] = crhs + gqt*js->vj;
    double si, sctemp;
    wo_sincos(js->phi, &si, &sctemp);
    double gcs = js->pfac*crt*sctemp;
    if (j->mod.pi) { si = -si; gcs = -gcs; }
    crt *= si;
    if (c->mode & MODEDC) {
        crhs += crt - gcs*js->phi;
        gqt = gcs;
    }
    else {
        crhs += crt - gcs*js->vj;
        gqt += gcs + c->ag[0]*j->cap;
        crhs += j->delVdelT*j->cap;
    }
    if (j->phs > 0) {
        int pint;
        memcpy(&pint, &c->states[1][j->st + JS_phsInt], sizeof(int));
        c->rhs[j->phs] = js->phi + (4*M_PI)*pint;
    }
    if (!j->neg) {
        ldadd(c, j->pPP, gqt);
        rhsadd(c, j->pos, -crhs);
    }
    else if (!j->pos) {
        ldadd(c, j->pNN, gqt);
        rhsadd(c, j->neg, crhs);
    }
    else {
        ldadd(c, j->pPP, gqt);
        ldadd(c, j->pPN, -gqt);
        ldadd(c, j->pNP, -gqt);
        ldadd(c, j->pNN, gqt);
        rhsadd(c, j->pos, -crhs);
        rhsadd(c, j->neg, crhs);
    }
}

static void jj_shunt(wo_ckt *c, JJ *j)
{
    if (j->gshunt > 1e-12) {
        ldadd(c, j->pRPP, j->gshunt);
        ldadd(c, j->pRPN, -j->gshunt);
        ldadd(c, j->pRNP, -j->gshunt);
        ldadd(c, j->pRNN, j->gshunt);
    }
}

static void jj_load(wo_ckt *c, JJ *j)
{
    /* JJdev::load, jjload.cc:77-590 */
    JJStuff js;
    memset(&js, 0, sizeof(js));
    double *s0 = c->states[0] + j->st, *s1 = c->states[1] + j->st;
    if (c->mode & MODEDC) {
        /* :154-250 phase-mode DC.  Reached here only by the single MODEUIC|MODETRANOP|MODEINITJCT
         * load that precedes a `uic` transient (tranan.cc:342-350); no solve follows it. */
        js.ci = 0;
        js.vj = c->rhsOld[j->pos] - c->rhsOld[j->neg];
        if (j->mod.ictype != 0) {
            js.phi = js.vj;
            s0[JS_phase] = js.phi;
            s0[JS_voltage] = 0.0;
            js.pfac = 1.0;
            js.gqt = 0; js.crhs = 0; js.dcrt = 0;
            js.crt = j->criti;
            jj_stamp(c, j, &js);
        }
        else {
            js.phi = 0.0; js.crhs = 0.0; js.dcrt = 0.0; js.crt = 0.0;
            jj_iv(j, &js);
            jj_stamp(c, j, &js);
            jj_shunt(c, j);
        }
        return;
    }
    js.pfac = c->delta/PHI0_2PI;
    if (c->order > 1) js.pfac *= .5;

    if (c->mode & MODEINITFLOAT) {
        js.ci = 0;
        js.vj = c->rhsOld[j->pos] - c->rhsOld[j->neg];
        double absvj = fabs(js.vj);
        double temp = s0[JS_voltage];
        double absold = fabs(temp);
        double maxvj = SPMAX(absvj, absold);
        temp -= js.vj;
        double absdvj = fabs(temp);
        if (maxvj >= j->vless) {
            temp = j->mod.delv*0.5;
            if (absdvj > temp) {
                js.ddv = temp;
                jj_limiting(c, j, &js);
                absvj = fabs(js.vj);
                maxvj = SPMAX(absvj, absold);
                temp = js.vj - s0[JS_voltage];
                absdvj = fabs(temp);
            }
        }
        /* :335-348 convergence test inside the load */
        if (!c->noncon) {
            double tol = c->opt.reltol*maxvj + c->opt.abstol;
            if (absdvj > tol) c->noncon++;
        }
        js.phi = s1[JS_phase];
        temp = js.vj;
        if (c->order > 1) temp += s1[JS_voltage];
        js.phi += js.pfac*temp;
        js.crhs = 0; js.dcrt = 0;
        js.crt = j->criti;
        jj_iv(j, &js);
        if (j->mod.ictype != 1) js.crt = 0.0;            /* jj_ic, :747-787, ictype 0 */
        jj_stamp(c, j, &js);
        jj_shunt(c, j);
        integrate2(c, j->st + JS_voltage, j->delVdelT);
        return;
    }
    if (c->mode & MODEINITPRED) {
        /* :416-462 */
        double y0 = s1[JS_dvdt]*c->pred[0] + c->states[2][j->st + JS_dvdt]*c->pred[1];
        if (c->order != 1) y0 += s1[JS_dvdt];
        double temp = s1[JS_voltage];
        double rag0 = c->order == 1 ? c->delta : .5*c->delta;
        js.vj = temp + rag0*y0;
        js.phi = s1[JS_phase];
        temp = js.vj;
        if (c->order > 1) temp += s1[JS_voltage];
        js.phi += js.pfac*temp;
        js.ci = 0;
        j->delVdelT = find_ceq(c, j->st + JS_voltage);
        js.crhs = 0; js.dcrt = 0;
        js.crt = j->criti;
        jj_iv(j, &js);
        if (j->mod.ictype != 1) js.crt = 0.0;
        jj_stamp(c, j, &js);
        jj_shunt(c, j);
        return;
    }
    if (c->mode & MODEINITTRAN) {
        /* :495-587 */
        if (c->mode & MODEUIC) {
            js.vj = j->initVoltage;
            js.phi = j->initPhase;
            js.ci = 0;
            s1[JS_voltage] = js.vj;
            s1[JS_phase] = js.phi;
            s1[JS_conI] = js.ci;
        }
        else {
            js.vj = 0;
            s1[JS_voltage] = 0.0;
            js.phi = s1[JS_phase];
            js.ci = 0;
            s1[JS_conI] = js.ci;
            s0[JS_voltage] = 0.0;
        }
        j->delVdelT = find_ceq(c, j->st + JS_voltage);
        js.crhs = 0; js.dcrt = 0;
        js.crt = j->criti;
        jj_iv(j, &js);
        if (j->mod.ictype != 1) js.crt = 0.0;
        jj_stamp(c, j, &js);
        jj_shunt(c, j);
    }
}

static void jj_accept(wo_ckt *c)
{
    /* JJdev::accept, jjacct.cc:49-145.  The reference loops model by model; the time-step
     * bound only needs the largest |vj| per model (all instances of a deck's models here share
     * vdpbak / tsaccl / tsfact per model index). */
    for (int mi = 0; mi < c->njj; mi++) {
        /* models in instance order: process each distinct model once */
        int first = 1;
        for (int k = 0; k < mi; k++) if (c->jj[k].in.model == c->jj[mi].in.model) { first = 0; break; }
        if (!first) continue;
        const b200_jj_model *m = &c->jj[mi].mod;
        int didm = 0;
        double vmax = 0;
        double vth = m->vdpbak/m->tsaccl;
        for (int i = 0; i < c->njj; i++) {
            JJ *j = &c->jj[i];
            if (j->in.model != c->jj[mi].in.model) continue;
            double *s0 = c->states[0] + j->st;
            double phi = s0[JS_phase];
            int pint;
            memcpy(&pint, &c->states[1][j->st + JS_phsInt], sizeof(int));
            double twopi = M_PI + M_PI;
            double fourpi = twopi + twopi;
            if (phi >= twopi) { phi -= fourpi; pint++; }
            else if (phi < -twopi) { phi += fourpi; pint--; }
            s0[JS_phase] = phi;
            { double z = 0.0; memcpy(&z, &pint, sizeof(int)); memcpy(&s0[JS_phsInt], &pint, sizeof(int)); (void)z; }
            double phitot = phi + fourpi*pint;
            if (j->phs > 0) c->rhsOld[j->phs] = phitot;
            if (j->mod.ictype != 0 && j->criti > 0) {
                if (!didm) { didm = 1; if (vmax < vth) vmax = vth; }
                double vj = s0[JS_voltage];
                if (vj < 0) vj = -vj;
                if (vmax < vj) vmax = vj;
            }
        }
        if (vmax > 0.0) {
            double delmax = m->tsfact*PHI0/vmax;
            if (c->devMaxDelta == 0.0 || delmax < c->devMaxDelta) c->devMaxDelta = delmax;
        }
    }
}

static void jj_getic(wo_ckt *c)
{
    /* JJdev::getic, jjic.cc:49-107 (vmax is NOT reset between models, as there) */
    for (int i = 0; i < c->njj; i++) {
        JJ *j = &c->jj[i];
        j->initVoltage = GIVEN(j->in.icv) ? j->in.icv : c->rhs[j->pos] - c->rhs[j->neg];
        j->initPhase = GIVEN(j->in.icphi) ? j->in.icphi : (j->phs > 0 ? c->rhs[j->phs] : 0);
    }
    double vmax = 0;
    for (int mi = 0; mi < c->njj; mi++) {
        int first = 1;
        for (int k = 0; k < mi; k++) if (c->jj[k].in.model == c->jj[mi].in.model) { first = 0; break; }
        if (!first) continue;
        const b200_jj_model *m = &c->jj[mi].mod;
        double vth = m->vdpbak/m->tsaccl;
        if (m->ictype != 0) {
            if (vmax < vth) vmax = vth;
            for (int i = 0; i < c->njj; i++) {
                JJ *j = &c->jj[i];
                if (j->in.model != c->jj[mi].in.model) continue;
                if (j->criti > 0) {
                    double temp = j->initVoltage;
                    if (temp < 0) temp = -temp;
                    if (vmax < temp) vmax = temp;
                }
            }
        }
        if (vmax > 0) {
            double delmax = 0.1*m->tsfact*PHI0/vmax;
            if (c->initDelta == 0.0 || c->initDelta > delmax) c->initDelta = delmax;
        }
    }
}


/* ------------------------------------------------------------------------- */
/* Josephson junction, tunnel-junction model (JJ level 3): devlib/tjm/{tjmset,tjmload,tjmacct,tjmic,tjmtrun}.cc */
/* complex arithmetic as include/ifcx.h:100-150                               */
/* ------------------------------------------------------------------------- */
static inline Cx cx_mul(Cx a, Cx b) { Cx r = { a.re*b.re - a.im*b.im, a.im*b.re + a.re*b.im }; return r; }
static inline Cx cx_muld(Cx a, double d) { Cx r = { a.re*d, a.im*d }; return r; }
static inline Cx cx_add(Cx a, Cx b) { Cx r = { a.re + b.re, a.im + b.im }; return r; }
static inline Cx cx_div(Cx a, Cx b)
{
    double d = b.re*b.re + b.im*b.im;
    Cx r = { (a.re*b.re + a.im*b.im)/d, (a.im*b.re - a.re*b.im)/d };
    return r;
}

static void tjm_setup_instance(TJM *j)
{
    /* tjmset.cc:572-660, :714: what depends on the junction's area */
    const b200_tjm_model *m = &j->mod;
    j->vshunt = GIVEN(j->in.vshunt) ? j->in.vshunt : m->vshunt;
    if (GIVEN(j->in.vshunt) && (j->in.vshunt < 0.0 || j->in.vshunt > m->vgnom)) j->vshunt = 0.0;
    if (GIVEN(j->in.ics)) {
        double ics = j->in.ics;
        if (ics < m->criti/50 || ics > m->criti*50) ics = m->criti;
        j->area = ics/m->criti;
    }
    else if (GIVEN(j->in.area)) j->area = j->in.area;
    else j->area = 1;
    j->criti = m->criti*j->area*j->in.tcf;
    double sqrta = sqrt(j->area);
    double gfac = j->area*(1.0 - m->gmu) + sqrta*m->gmu;
    j->gqp = gfac/m->rsint;
    j->g0 = m->r0 > 0.0 ? gfac/m->r0 : 0.0;
    double tg0 = j->g0;
    if (j->g0 >= j->gqp) j->g0 -= j->gqp;
    j->gshunt = 0.0;
    if (j->vshunt > 1e-12) {
        double gshunt = j->criti/j->vshunt - SPMAX(tg0, j->gqp);
        if (gshunt > 0.0) j->gshunt = gshunt;
    }
    j->cap = m->cap*(j->area*(1.0 - m->cmu) + sqrta*m->cmu);
}

static void tjm_init(TJM *j, double phi)
{
    /* sTJMinstance::tjm_init, tjmload.cc:459-487 */
    double s2, c2;
    wo_sincos(0.5*phi, &s2, &c2);
    j->sinphi_2_old = s2; j->cosphi_2_old = c2;
    if (j->mod.rtype > 0) j->gcrit = j->in.alphaN*sqrt(j->cap*j->criti/PHI0_2PI);
    else j->gcrit = 0.0;
    for (int i = 0; i < j->n; i++) {
        j->Fcprev[i].re = c2; j->Fcprev[i].im = 0.0;
        j->Fsprev[i].re = s2; j->Fsprev[i].im = 0.0;
        j->Fc[i].re = j->Fc[i].im = j->Fs[i].re = j->Fs[i].im = 0.0;
    }
}

static void tjm_newstep(wo_ckt *c, TJM *j)
{
    /* tjmload.cc:490-505 */
    double kdt = j->in.kgap*j->in.omegaJ*c->delta;
    for (int i = 0; i < j->n; i++) {
        Cx z = cx_muld(j->p[i], kdt);
        double d = exp(z.re);
        Cx ez = { d*cos(z.im), d*sin(z.im) };
        j->exp_z[i] = ez;
        Cx em1 = { ez.re - 1.0, ez.im };
        Cx zt = cx_div(em1, z);
        j->alpha0[i].re = zt.re - ez.re; j->alpha0[i].im = zt.im - ez.im;
        j->alpha1[i].re = 1.0 - zt.re;   j->alpha1[i].im = -zt.im;
    }
}

static void tjm_update(TJM *j, double phi)
{
    /* tjmload.cc:508-553 */
    double s2, c2;
    wo_sincos(0.5*phi, &s2, &c2);
    double FcSp = 0.0, FcSq = 0.0, FsSp = 0.0, FsSq = 0.0;
    for (int i = 0; i < j->n; i++) {
        j->Fc[i] = cx_add(cx_add(cx_mul(j->exp_z[i], j->Fcprev[i]), cx_muld(j->alpha0[i], j->cosphi_2_old)),
            cx_muld(j->alpha1[i], c2));
        j->Fs[i] = cx_add(cx_add(cx_mul(j->exp_z[i], j->Fsprev[i]), cx_muld(j->alpha0[i], j->sinphi_2_old)),
            cx_muld(j->alpha1[i], s2));
        FcSp += cx_mul(j->A[i], j->Fc[i]).re;
        FcSq += cx_mul(j->B[i], j->Fc[i]).re;
        FsSp += cx_mul(j->A[i], j->Fs[i]).re;
        FsSq += cx_mul(j->B[i], j->Fs[i]).re;
    }
    double fct = j->criti*j->in.kgap_rejpt;
    j->cp = j->mod.ictype > 0 ? fct*(s2*FcSp + c2*FsSp) : 0.0;
    j->cqp = j->mod.rtype > 0 ? fct*(s2*FcSq - c2*FsSq) : 0.0;
}

static void tjm_stamp(wo_ckt *c, TJM *j, double vj, double phi, double pfac, double crt)
{
    /* sTJMinstance::tjm_load, tjmload.cc:412-456 */
    double *s0 = c->states[0] + j->st;
    s0[TS_voltage] = vj;
    s0[TS_phase] = phi;
    double crhs, gqt;
    if (c->mode & MODEDC) {
        double cr = crt*sin(phi);
        gqt = pfac*crt*cos(phi);
        crhs = cr - gqt*phi;
        j->cp = crt*sin(phi);
        s0[TS_crti] = j->cp;
        s0[TS_qpi] = 0.0;
    }
    else {
        tjm_update(j, phi);
        crhs = j->cp + j->cqp + j->delVdelT*j->cap;
        gqt = j->gcrit + j->g0 + c->ag[0]*j->cap;
        s0[TS_crti] = j->cp;
        s0[TS_qpi] = j->cqp + (j->gcrit + j->g0)*vj;
    }
    if (j->phs > 0) {
        int pint;
        memcpy(&pint, &c->states[1][j->st + TS_phsInt], sizeof(int));
        c->rhs[j->phs] = phi + (4*M_PI)*pint;
    }
    if (!j->neg) { ldadd(c, j->pPP, gqt); rhsadd(c, j->pos, -crhs); }
    else if (!j->pos) { ldadd(c, j->pNN, gqt); rhsadd(c, j->neg, crhs); }
    else {
        ldadd(c, j->pPP, gqt); ldadd(c, j->pPN, -gqt); ldadd(c, j->pNP, -gqt); ldadd(c, j->pNN, gqt);
        rhsadd(c, j->pos, -crhs); rhsadd(c, j->neg, crhs);
    }
}

static void tjm_shunt(wo_ckt *c, TJM *j)
{
    if (j->gshunt > 1e-12) {
        ldadd(c, j->pRPP, j->gshunt); ldadd(c, j->pRPN, -j->gshunt);
        ldadd(c, j->pRNP, -j->gshunt); ldadd(c, j->pRNN, j->gshunt);
    }
}

static void tjm_load(wo_ckt *c, TJM *j)
{
    /* TJMdev::load, tjmload.cc:65-392 */
    double *s0 = c->states[0] + j->st, *s1 = c->states[1] + j->st;
    if (c->mode & MODEDC) {
        if (c->mode & MODEUIC) return;                        /* :75-76 */
        double vj = c->rhsOld[j->pos] - c->rhsOld[j->neg];
        s0[TS_phase] = vj; s0[TS_voltage] = 0.0;
        tjm_stamp(c, j, vj, vj, 1.0, j->criti);
        return;
    }
    double pfac = c->delta/PHI0_2PI;
    if (c->order > 1) pfac *= .5;
    if (c->mode & MODEINITFLOAT) {
        double vj = c->rhsOld[j->pos] - c->rhsOld[j->neg];
        double absvj = fabs(vj);
        double temp = s0[TS_voltage];
        double absold = fabs(temp);
        double maxvj = SPMAX(absvj, absold);
        temp -= vj;
        double absdvj = fabs(temp);
        if (!c->noncon) {
            double tol = c->opt.reltol*maxvj + c->opt.abstol;
            if (absdvj > tol) c->noncon++;
        }
        double phi = s1[TS_phase];
        temp = vj;
        if (c->order > 1) temp += s1[TS_voltage];
        phi += pfac*temp;
        /* the reference integrates BEFORE tjm_load overwrites state 0's voltage with the new vj
         * (tjmload.cc:209-210): dvdt comes from the PREVIOUS iteration's voltage */
        integrate2(c, j->st + TS_voltage, j->delVdelT);
        tjm_stamp(c, j, vj, phi, pfac, j->criti);
        tjm_shunt(c, j);
        return;
    }
    if (c->mode & MODEINITPRED) {
        double y0 = s1[TS_dvdt]*c->pred[0] + c->states[2][j->st + TS_dvdt]*c->pred[1];
        if (c->order != 1) y0 += s1[TS_dvdt];
        double temp = s1[TS_voltage];
        double rag0 = c->order == 1 ? c->delta : .5*c->delta;
        double vj = temp + rag0*y0;
        double phi = s1[TS_phase];
        temp = vj;
        if (c->order > 1) temp += s1[TS_voltage];
        phi += pfac*temp;
        j->delVdelT = find_ceq(c, j->st + TS_voltage);
        tjm_newstep(c, j);
        tjm_stamp(c, j, vj, phi, pfac, j->criti);
        tjm_shunt(c, j);
        return;
    }
    if (c->mode & MODEINITTRAN) {
        double vj, phi;
        if (c->mode & MODEUIC) {
            vj = j->initVoltage; phi = j->initPhase;
            s1[TS_voltage] = vj; s1[TS_phase] = phi;
        }
        else {
            /* tjmload.cc:310-316 (NEWJJDC): the phase computed by the DC operating point carries over */
            vj = 0; s1[TS_voltage] = 0.0; phi = s1[TS_phase]; s0[TS_voltage] = 0.0;
        }
        j->delVdelT = find_ceq(c, j->st + TS_voltage);
        tjm_init(j, phi);
        tjm_newstep(c, j);
        tjm_stamp(c, j, vj, phi, pfac, j->criti);
        tjm_shunt(c, j);
    }
}

static void tjm_accept(wo_ckt *c)
{
    /* TJMdev::accept + sTJMinstance::tjm_accept, tjmacct.cc:43-169.  Instances are visited in
     * load order; the floor vth is the FIRST visited instance's (didm is set once per model). */
    for (int mi = 0; mi < c->ntjm; mi++) {
        int first = 1;
        for (int k = 0; k < mi; k++) if (c->tjm[k].in.model == c->tjm[mi].in.model) { first = 0; break; }
        if (!first) continue;
        const b200_tjm_model *m = &c->tjm[mi].mod;
        int didm = 0;
        double vmax = 0;
        for (int o = 0; o < c->nord; o++) {
            if (c->ord[o].kind != B200_DEV_TJM) continue;
            TJM *j = &c->tjm[c->ord[o].index];
            if (j->in.model != c->tjm[mi].in.model) continue;
            double vth = j->in.vdpbak/m->tsaccl;
            double *s0 = c->states[0] + j->st;
            double phi = s0[TS_phase];
            int pint;
            memcpy(&pint, &c->states[1][j->st + TS_phsInt], sizeof(int));
            double twopi = M_PI + M_PI, fourpi = twopi + twopi;
            if (phi >= twopi) { phi -= fourpi; pint++; }
            else if (phi < -twopi) { phi += fourpi; pint--; }
            s0[TS_phase] = phi;
            memcpy(&s0[TS_phsInt], &pint, sizeof(int));
            double phitot = phi + fourpi*pint;
            if (j->phs > 0) c->rhsOld[j->phs] = phitot;
            wo_sincos(0.5*phi, &j->sinphi_2_old, &j->cosphi_2_old);
            for (int i = 0; i < j->n; i++) { j->Fcprev[i] = j->Fc[i]; j->Fsprev[i] = j->Fs[i]; }
            if (!didm) { didm = 1; if (vmax < vth) vmax = vth; }
            double vj = s0[TS_voltage];
            if (vj < 0) vj = -vj;
            if (vmax < vj) vmax = vj;
        }
        if (vmax > 0.0) {
            double delmax = m->tsfact*PHI0/vmax;
            if (c->devMaxDelta == 0.0 || delmax < c->devMaxDelta) c->devMaxDelta = delmax;
        }
    }
}

static void tjm_getic(wo_ckt *c)
{
    /* TJMdev::getic, tjmic.cc:45-95 */
    for (int i = 0; i < c->ntjm; i++) {
        TJM *j = &c->tjm[i];
        j->initVoltage = GIVEN(j->in.icv) ? j->in.icv : c->rhs[j->pos] - c->rhs[j->neg];
        j->initPhase = GIVEN(j->in.icphi) ? j->in.icphi : (j->phs > 0 ? c->rhs[j->phs] : 0);
    }
    double vmax = 0;
    for (int mi = 0; mi < c->ntjm; mi++) {
        int first = 1;
        for (int k = 0; k < mi; k++) if (c->tjm[k].in.model == c->tjm[mi].in.model) { first = 0; break; }
        if (!first) continue;
        const b200_tjm_model *m = &c->tjm[mi].mod;
        if (m->ictype != 0) {
            for (int i = 0; i < c->ntjm; i++) {
                TJM *j = &c->tjm[i];
                if (j->in.model != c->tjm[mi].in.model) continue;
                double vth = j->in.vdpbak/m->tsaccl;
                if (vmax < vth) vmax = vth;
                if (j->criti > 0) {
                    double temp = fabs(j->initVoltage);
                    if (vmax < temp) vmax = temp;
                }
            }
        }
        if (vmax > 0) {
            double delmax = 0.1*m->tsfact*PHI0/vmax;
            if (c->initDelta == 0.0 || c->initDelta > delmax) c->initDelta = delmax;
        }
    }
}

static void res_load(wo_ckt *c, Res *r)
{
    /* RESdev::load, resload.cc:250-292 (linear resistor).  In the phase-mode DC load only the
     * single `uic` MODEINITJCT load gets here and no solve follows; its stamps are restated for
     * the phase / ground case that the built decks have (both terminals phase or ground). */
    if (r->usePreload) return;
    if ((c->mode & MODEDC) && c->phase) {
        int vp = r->pos && !c->phase[r->pos], vn = r->neg && !c->phase[r->neg];
        if (!vp && !vn) {
            ldadd(c, r->pPP, c->opt.gmin); ldadd(c, r->pNN, c->opt.gmin);
            ldadd(c, r->pPN, c->opt.gmin); ldadd(c, r->pNP, c->opt.gmin);
            return;
        }
        if (vp && r->neg && c->phase[r->neg]) { ldadd(c, r->pPP, r->conduct); ldadd(c, r->pNP, -r->conduct); return; }
        if (vn && r->pos && c->phase[r->pos]) { ldadd(c, r->pNN, r->conduct); ldadd(c, r->pPN, -r->conduct); return; }
    }
    ldadd(c, r->pPP, r->conduct);
    ldadd(c, r->pNN, r->conduct);
    ldadd(c, r->pPN, -r->conduct);
    ldadd(c, r->pNP, -r->conduct);
}

static void cap_stamp(wo_ckt *c, Cap *p)
{
    /* capload.cc:75-97 */
    double geq = p->geq, ceq = p->ceq;
    if (p->pos == 0) {
        ldadd(c, p->pNN, geq);
        rhsadd(c, p->neg, ceq);
        return;
    }
    if (p->neg == 0) {
        ldadd(c, p->pPP, geq);
        rhsadd(c, p->pos, -ceq);
        return;
    }
    ldadd(c, p->pPP, geq);
    ldadd(c, p->pNN, geq);
    ldadd(c, p->pPN, -geq);
    ldadd(c, p->pNP, -geq);
    rhsadd(c, p->pos, -ceq);
    rhsadd(c, p->neg, ceq);
}

static void cap_load(wo_ckt *c, Cap *p)
{
    /* capload.cc:104-205 (linear capacitor) */
    if (c->mode & MODEDC)
        return;
    if (c->mode & MODEINITFLOAT) {
        double qcap = p->capac*(c->rhsOld[p->pos] - c->rhsOld[p->neg]);
        c->states[0][p->st] = qcap;
        integrate2(c, p->st, p->ceq);
        p->ceq = find_ceq(c, p->st);
        cap_stamp(c, p);
        return;
    }
    if (c->mode & MODEINITPRED) {
        p->geq = c->ag[0]*p->capac;
        p->ceq = find_ceq(c, p->st);
        cap_stamp(c, p);
        return;
    }
    if (c->mode & MODEINITTRAN) {
        double vcap;
        if (c->mode & MODEUIC)
            vcap = p->ic;
        else
            vcap = c->rhsOld[p->pos] - c->rhsOld[p->neg];
        c->states[1][p->st] = p->capac*vcap;
        p->geq = c->ag[0]*p->capac;
        p->ceq = find_ceq(c, p->st);
        cap_stamp(c, p);
    }
}

static void mut_load(wo_ckt *c, Mut *p)
{
    /* mutload.cc:55-140: loaded BEFORE the inductors, adds the cross flux to their states */
    Ind *l1 = &c->ind[p->ind1], *l2 = &c->ind[p->ind2];
    if (p->factor == 0.0)
        p->factor = p->coupling*sqrt(l1->induct*l2->induct);
    if (c->mode & MODEDC) {
        if (c->phase) {                                   /* mutload.cc:75-87 */
            double res = 2*M_PI*p->factor/PHI0;
            ldadd(c, p->p12, -res);
            ldadd(c, p->p21, -res);
        }
        return;
    }
    if (c->mode & MODEINITFLOAT) {
        double factor = p->factor;
        c->states[0][l1->st] += factor*c->rhsOld[l2->br];
        c->states[0][l2->st] += factor*c->rhsOld[l1->br];
        factor *= c->ag[0];
        ldadd(c, p->p12, -factor);
        ldadd(c, p->p21, -factor);
    }
    else if (c->mode & MODEINITPRED) {
        double factor = p->factor*c->ag[0];
        ldadd(c, p->p12, -factor);
        ldadd(c, p->p21, -factor);
    }
    else if (c->mode & MODEINITTRAN) {
        double factor = p->factor;
        if (c->mode & MODEUIC) {
            c->states[1][l1->st] += factor*l2->ic;
            c->states[1][l2->st] += factor*l1->ic;
        }
        else {
            c->states[1][l1->st] += factor*c->rhsOld[l2->br];
            c->states[1][l2->st] += factor*c->rhsOld[l1->br];
        }
        factor *= c->ag[0];
        ldadd(c, p->p12, -factor);
        ldadd(c, p->p21, -factor);
    }
}

static void ind_load(wo_ckt *c, Ind *p)
{
    /* indload.cc:84-263 (linear inductor; the +-1 couplings are preloaded) */
    if (c->mode & MODEDC) {
        /* :88-96 phase-mode DC (CKTjjDCphase): the inductor's "voltage" is its phase */
        if (c->phase) {
            double res = 2*M_PI*p->induct/PHI0;
            ldadd(c, p->pBB, -res);
        }
        return;
    }
    if (c->mode & MODEINITFLOAT) {
        double flux = p->induct*c->rhsOld[p->br] - p->prevFlux;
        c->states[0][p->st] += flux;
        p->prevFlux = c->states[0][p->st];
        integrate2(c, p->st, p->veq);
        rhsadd(c, p->br, p->veq);
        ldadd(c, p->pBB, -p->req);
        return;
    }
    if (c->mode & MODEINITPRED) {
        p->req = c->ag[0]*p->induct;
        p->veq = find_ceq(c, p->st);
        rhsadd(c, p->br, p->veq);
        ldadd(c, p->pBB, -p->req);
        p->prevFlux = 0;
        c->states[0][p->st] = 0;
        return;
    }
    if (c->mode & MODEINITTRAN) {
        double ival;
        if (c->mode & MODEUIC)
            ival = p->ic;
        else
            ival = c->rhsOld[p->br];
        c->states[1][p->st] += p->induct*ival;
        p->prevFlux = 0;
        c->states[0][p->st] = 0;
        p->req = c->ag[0]*p->induct;
        p->veq = find_ceq(c, p->st);
        rhsadd(c, p->br, p->veq);
        ldadd(c, p->pBB, -p->req);
    }
}

static void src_load(wo_ckt *c, Src *s)
{
    /* srcload.cc:67-105, :215-262 (independent sources, no controlling variables) */
    int is_dc = c->mode & (MODEDCOP | MODEDCTRANCURVE);
    int ft;  /* 1 = SRC_dc, 2 = SRC_func */
    if (s->fn == B200_FN_NONE)
        ft = 1;
    else
        ft = is_dc ? (s->dcGiven ? 1 : 2) : 2;
    double rhs;
    if (ft == 1)
        rhs = c->srcFact*s->dc;
    else {
        rhs = src_eval(s, c->time);
        s->dc = rhs;                         /* srcload.cc:236-237 */
        if (c->mode & (MODEDCOP | MODETRANOP))
            rhs *= c->srcFact;
    }
    if (s->kind == B200_SRC_V)
        rhsadd(c, s->branch, rhs);
    else {
        rhsadd(c, s->pos, -rhs);
        rhsadd(c, s->neg, rhs);
    }
}

/* ------------------------------------------------------------------------- */
/* sCKT::load: ckt.cc:963-1207; loadGmin: ckt.cc:1221-1263, spspice.cc:259-368  */
/* ------------------------------------------------------------------------- */
static void preload(wo_ckt *c)
{
    /* spSaveForInitialization snapshot of the preloaded constants:
     * srcset.cc:158-170 (ldset +-1), restemp.cc:218-229 (ldadd +-g) */
    memset(c->Ainit, 0, c->nnz*sizeof(double));
    double *save = c->Ax;
    c->Ax = c->Ainit;
    for (int o = 0; o < c->nord; o++) {
        int i = c->ord[o].index;
        if (c->ord[o].kind == B200_DEV_SRC) {
            Src *s = &c->src[i];
            if (s->kind == B200_SRC_V) {
                if (s->pos) { if (s->pPB >= 0) c->Ax[s->pPB] = 1.0; if (s->pBP >= 0) c->Ax[s->pBP] = 1.0; }
                if (s->neg) { if (s->pNB >= 0) c->Ax[s->pNB] = -1.0; if (s->pBN >= 0) c->Ax[s->pBN] = -1.0; }
            }
        }
        else if (c->ord[o].kind == B200_DEV_IND) {
            /* indset.cc:117-131 */
            Ind *p = &c->ind[i];
            if (p->pos) { if (p->pPB >= 0) c->Ax[p->pPB] = 1.0; if (p->pBP >= 0) c->Ax[p->pBP] = 1.0; }
            if (p->neg) { if (p->pNB >= 0) c->Ax[p->pNB] = -1.0; if (p->pBN >= 0) c->Ax[p->pBN] = -1.0; }
        }
        else if (c->ord[o].kind == B200_DEV_JJ) {
            /* jjset.cc:815-816: the phase node's own equation, x[phs] = rhs[phs] */
            if (c->jj[i].phs > 0 && c->jj[i].pHH >= 0) c->Ax[c->jj[i].pHH] = 1.0;
        }
        else if (c->ord[o].kind == B200_DEV_TJM) {
            if (c->tjm[i].phs > 0 && c->tjm[i].pHH >= 0) c->Ax[c->tjm[i].pHH] = 1.0;
        }
    }
    for (int o = 0; o < c->nord; o++) {
        int i = c->ord[o].index;
        if (c->ord[o].kind == B200_DEV_RES) {
            Res *r = &c->res[i];
            /* restemp.cc:213-223: a resistor on a phase node is loaded, not preloaded */
            r->usePreload = !(c->phase && ((r->pos && c->phase[r->pos]) || (r->neg && c->phase[r->neg])));
            if (r->usePreload && r->conduct != 0.0) {
                ldadd(c, r->pPP, r->conduct);
                ldadd(c, r->pNN, r->conduct);
                ldadd(c, r->pPN, -r->conduct);
                ldadd(c, r->pNP, -r->conduct);
            }
        }
    }
    c->Ax = save;
}

static int ckt_load(wo_ckt *c)
{
    c->devMaxDelta = 0.0;
    memset(c->rhs, 0, (c->n + 1)*sizeof(double));
    memcpy(c->Ax, c->Ainit, c->nnz*sizeof(double));
    int tchk = c->trapCheck;
    c->trapCheck = tchk && (c->mode & MODEINITFLOAT) && (c->mode & MODETRAN);
    c->trapBad = 0;
    for (int o = 0; o < c->nord; o++) {
        int i = c->ord[o].index;
        switch (c->ord[o].kind) {
        case B200_DEV_MOS: mos_load(c, &c->mos[i]); break;
        case B200_DEV_CAP: cap_load(c, &c->cap[i]); break;
        case B200_DEV_IND: ind_load(c, &c->ind[i]); break;
        case B200_DEV_MUT: mut_load(c, &c->mut[i]); break;
        case B200_DEV_SRC: src_load(c, &c->src[i]); break;
        case B200_DEV_JJ: jj_load(c, &c->jj[i]); break;
        case B200_DEV_TJM: tjm_load(c, &c->tjm[i]); break;
        case B200_DEV_RES: res_load(c, &c->res[i]); break;   /* preloaded ones return at once (LOAD_SKIP_FLAG) */          /* preloaded, LOAD_SKIP_FLAG */
        }
    }
    c->trapCheck = tchk;
    if (c->mode & MODEDC) {
        /* ckt.cc:1177-1204 */
        if (c->mode & (MODEINITJCT | MODEINITFIX)) {
            for (int i = 1; i <= c->n; i++)
                if (c->ns_given[i]) {
                    c->rhs[i] += c->ns_val[i];
                    ldadd(c, c->diag[i], 1.0);
                }
        }
        if ((c->mode & MODETRANOP) && !(c->mode & MODEUIC)) {
            for (int i = 1; i <= c->n; i++)
                if (c->ic_given[i]) {
                    c->rhs[i] += c->ic_val[i];
                    ldadd(c, c->diag[i], 1.0);
                }
        }
    }
    return OK;
}

static int load_gmin(wo_ckt *c)
{
    /* ckt.cc:1221-1263 -> spspice.cc:259-300.  A voltage node with no diagonal is accepted
     * when its column holds a +-1 (junction of branch devices); we skip it. */
    if (c->diagGmin != 0.0) {
        for (int i = 1; i <= c->n; i++)
            if (c->ntype[i] == SP_VOLTAGE && c->diag[i] >= 0)
                c->Ax[c->diag[i]] += c->diagGmin;
        return OK;
    }
    int checkmin = c->opt.forcegmin;
    for (int i = 1; i <= c->n; i++)
        if (c->ntype[i] == SP_VOLTAGE && c->diag[i] >= 0) {
            double *r = &c->Ax[c->diag[i]];
            if (checkmin && (*r >= 0.0 && *r < c->opt.gmin))
                *r = c->opt.gmin;
            else if (*r > c->opt.gmax)
                *r = c->opt.gmax;
        }
    return OK;
}

/* ------------------------------------------------------------------------- */
/* Sparse LU.  Restated klu_refactor/klu_solve contract (kluif.cc:594-628):     */
/* left-looking column LU with a FIXED pivot sequence (static row/col order),  */
/* no scaling; L unit-diagonal stored as x[i]/pivot; updates applied in         */
/* ascending pivot order so every entry sees the same operation sequence as the */
/* engine's compiled program.  PARITY UNPINNED vs SuiteSparse (source absent).  */
/* ------------------------------------------------------------------------- */
typedef struct {
    int n;
    int *Lp, *Li, *Up, *Ui;     /* patterns in permuted positions, rows ascending */
    double *Lx, *Ux, *Ud;
    int *rinv;                  /* original row -> position */
} StaticLU;

static StaticLU *slu_symbolic(int n, const int *Ap, const int *Ai, const int *rowperm,
    const int *colperm)
{
    StaticLU *f = xcalloc(1, sizeof(*f));
    f->n = n;
    f->rinv = xcalloc(n, sizeof(int));
    for (int k = 0; k < n; k++) f->rinv[rowperm[k]] = k;
    char *mark = xcalloc(n, 1);
    int capL = 4*n + 16, capU = 4*n + 16;
    f->Lp = xcalloc(n + 1, sizeof(int)); f->Up = xcalloc(n + 1, sizeof(int));
    f->Li = xcalloc(capL, sizeof(int)); f->Ui = xcalloc(capU, sizeof(int));
    int nl = 0, nu = 0;
    for (int j = 0; j < n; j++) {
        int cj = colperm[j];
        memset(mark, 0, n);
        for (int p = Ap[cj]; p < Ap[cj+1]; p++) mark[f->rinv[Ai[p]]] = 1;
        for (int k = 0; k < j; k++) {
            if (!mark[k]) continue;
            for (int p = f->Lp[k]; p < f->Lp[k+1]; p++) mark[f->Li[p]] = 1;
        }
        for (int k = 0; k < j; k++)
            if (mark[k]) {
                if (nu == capU) { capU *= 2; f->Ui = realloc(f->Ui, capU*sizeof(int)); }
                f->Ui[nu++] = k;
            }
        f->Up[j+1] = nu;
        for (int k = j + 1; k < n; k++)
            if (mark[k]) {
                if (nl == capL) { capL *= 2; f->Li = realloc(f->Li, capL*sizeof(int)); }
                f->Li[nl++] = k;
            }
        f->Lp[j+1] = nl;
    }
    f->Lx = xcalloc(nl, sizeof(double));
    f->Ux = xcalloc(nu, sizeof(double));
    f->Ud = xcalloc(n, sizeof(double));
    free(mark);
    return f;
}

static void slu_free(StaticLU *f)
{
    if (!f) return;
    free(f->Lp); free(f->Li); free(f->Up); free(f->Ui); free(f->Lx); free(f->Ux); free(f->Ud);
    free(f->rinv); free(f);
}

/* The engine evaluates every LU entry / solve row as ONE left-to-right sum
 *      x0 - a[0]*b[0] - a[1]*b[1] - ...          (products in ascending pivot order)
 * and splits a sum of more than WO_LU_SPLIT products so that no single lane carries a long
 * dependent chain: chunks of WO_LU_CHUNK consecutive products are accumulated from 0.0, and the
 * partial sums p_c are then subtracted as products p_c*(-1.0) under the same rule (recursively).
 * Restated here so both sides round identically (xictools_b200/csrc/b200_internal.h LU_SPLIT,
 * LU_CHUNK; tests/test_abi.py checks the constants agree). */
#define WO_LU_SPLIT 32
#define WO_LU_CHUNK 16
static double split_eval(double x0, const double *a, const double *b, int m)
{
    if (m <= WO_LU_SPLIT) {
        double acc = x0;
        for (int q = 0; q < m; q++) acc -= a[q]*b[q];
        return acc;
    }
    int nch = (m + WO_LU_CHUNK - 1)/WO_LU_CHUNK;
    double *pa = xcalloc(nch, sizeof(double)), *pb = xcalloc(nch, sizeof(double));
    for (int c = 0; c < nch; c++) {
        int len = m - c*WO_LU_CHUNK; if (len > WO_LU_CHUNK) len = WO_LU_CHUNK;
        pa[c] = split_eval(0.0, a + c*WO_LU_CHUNK, b + c*WO_LU_CHUNK, len);
        pb[c] = -1.0;
    }
    double r = split_eval(x0, pa, pb, nch);
    free(pa); free(pb);
    return r;
}
void wo_lu_split_rule(int *split, int *chunk) { *split = WO_LU_SPLIT; *chunk = WO_LU_CHUNK; }

static int slu_refactor(StaticLU *f, const int *Ap, const int *Ai, const double *Ax,
    const int *colperm, double *x /* n work */)
{
    int n = f->n;
    /* per-row product lists of the current column: (L value index, U value index) */
    int cap = 64, *tr = xcalloc(3*(size_t)cap, sizeof(int));
    int *cnt = xcalloc(n + 1, sizeof(int)), *start = xcalloc(n + 1, sizeof(int));
    double *va = NULL, *vb = NULL; int vcap = 0;
    int err = OK;
    for (int j = 0; j < n && !err; j++) {
        int cj = colperm[j];
        /* scatter (pattern of column j of L+U is a superset) */
        for (int p = f->Up[j]; p < f->Up[j+1]; p++) x[f->Ui[p]] = 0.0;
        x[j] = 0.0;
        for (int p = f->Lp[j]; p < f->Lp[j+1]; p++) x[f->Li[p]] = 0.0;
        for (int p = Ap[cj]; p < Ap[cj+1]; p++) x[f->rinv[Ai[p]]] = Ax[p];
        /* products in ascending pivot order k, bucketed by target row (stable) */
        int nt = 0;
        for (int p = f->Up[j]; p < f->Up[j+1]; p++) {
            int k = f->Ui[p];
            for (int q = f->Lp[k]; q < f->Lp[k+1]; q++) {
                if (nt == cap) { cap *= 2; tr = realloc(tr, 3*(size_t)cap*sizeof(int)); }
                tr[3*nt] = f->Li[q]; tr[3*nt+1] = q; tr[3*nt+2] = p; nt++;
            }
        }
        if (nt > vcap) { vcap = 2*nt + 16; va = realloc(va, vcap*sizeof(double)); vb = realloc(vb, vcap*sizeof(double)); }
        for (int i = 0; i <= n; i++) cnt[i] = 0;
        for (int e = 0; e < nt; e++) cnt[tr[3*e]]++;
        start[0] = 0;
        for (int i = 0; i < n; i++) start[i+1] = start[i] + cnt[i];
        int *ord = xcalloc(nt + 1, sizeof(int)), *fill = xcalloc(n + 1, sizeof(int));
        for (int e = 0; e < nt; e++) { int i = tr[3*e]; ord[start[i] + fill[i]++] = e; }
        /* rows top to bottom: U(i,j) for i < j, the pivot, then L(i,j) = value / pivot */
#define ROW_VALUE(i, dst) do { \
            int m_ = cnt[i]; \
            for (int z = 0; z < m_; z++) { int e = ord[start[i] + z]; va[z] = f->Lx[tr[3*e+1]]; vb[z] = f->Ux[tr[3*e+2]]; } \
            (dst) = split_eval(x[i], va, vb, m_); } while (0)
        for (int p = f->Up[j]; p < f->Up[j+1]; p++) { int i = f->Ui[p]; ROW_VALUE(i, f->Ux[p]); }
        double piv;
        ROW_VALUE(j, piv);
        f->Ud[j] = piv;
        if (piv == 0.0 || !isfinite(piv)) err = E_SINGULAR;
        else
            for (int p = f->Lp[j]; p < f->Lp[j+1]; p++) { int i = f->Li[p]; double v; ROW_VALUE(i, v); f->Lx[p] = v/piv; }
#undef ROW_VALUE
        free(ord); free(fill);
    }
    free(tr); free(cnt); free(start); free(va); free(vb);
    return err;
}

static void slu_solve(StaticLU *f, const int *rowperm, const int *colperm, double *b /* n */,
    double *y /* n work */)
{
    int n = f->n;
    for (int k = 0; k < n; k++) y[k] = b[rowperm[k]];
    /* row-wise product lists from the column structure */
    int nl = f->Lp[n], nu = f->Up[n];
    int m = (nl > nu ? nl : nu) + 1;
    int *ptr = xcalloc(n + 2, sizeof(int)), *fill = xcalloc(n + 1, sizeof(int));
    int *ev = xcalloc(m, sizeof(int)), *ej = xcalloc(m, sizeof(int));
    double *va = xcalloc(m, sizeof(double)), *vb = xcalloc(m, sizeof(double));
    /* forward: y(i) = b(i) - sum_k L(i,k)*y(k), ascending k */
    for (int k = 0; k < n; k++) for (int q = f->Lp[k]; q < f->Lp[k+1]; q++) ptr[f->Li[q] + 1]++;
    for (int i = 0; i < n; i++) ptr[i+1] += ptr[i];
    for (int k = 0; k < n; k++)
        for (int q = f->Lp[k]; q < f->Lp[k+1]; q++) { int i = f->Li[q]; int at = ptr[i] + fill[i]++; ev[at] = q; ej[at] = k; }
    for (int i = 0; i < n; i++) {
        int c = ptr[i+1] - ptr[i];
        for (int z = 0; z < c; z++) { va[z] = f->Lx[ev[ptr[i] + z]]; vb[z] = y[ej[ptr[i] + z]]; }
        y[i] = split_eval(y[i], va, vb, c);
    }
    /* backward: x(i) = (y(i) - sum_j U(i,j)*x(j)) / U(i,i), descending j */
    for (int i = 0; i <= n + 1; i++) ptr[i] = 0;
    for (int i = 0; i <= n; i++) fill[i] = 0;
    for (int j = 0; j < n; j++) for (int p = f->Up[j]; p < f->Up[j+1]; p++) ptr[f->Ui[p] + 1]++;
    for (int i = 0; i < n; i++) ptr[i+1] += ptr[i];
    for (int j = n - 1; j >= 0; j--)
        for (int p = f->Up[j+1] - 1; p >= f->Up[j]; p--) { int i = f->Ui[p]; int at = ptr[i] + fill[i]++; ev[at] = p; ej[at] = j; }
    for (int i = n - 1; i >= 0; i--) {
        int c = ptr[i+1] - ptr[i];
        for (int z = 0; z < c; z++) { va[z] = f->Ux[ev[ptr[i] + z]]; vb[z] = y[ej[ptr[i] + z]]; }
        y[i] = split_eval(y[i], va, vb, c)/f->Ud[i];
    }
    for (int j = 0; j < n; j++) b[colperm[j]] = y[j];
    free(ptr); free(fill); free(ev); free(ej); free(va); free(vb);
}

int wo_lu_solve_static(int n, const int *Ap, const int *Ai, const double *Ax,
    const int *rowperm, const int *colperm, double *b)
{
    StaticLU *f = slu_symbolic(n, Ap, Ai, rowperm, colperm);
    double *x = xcalloc(n, sizeof(double)), *y = xcalloc(n, sizeof(double));
    int err = slu_refactor(f, Ap, Ai, Ax, colperm, x);
    if (!err) slu_solve(f, rowperm, colperm, b, y);
    free(x); free(y); slu_free(f);
    return err;
}

/* Own-order mode (no engine ordering supplied): dense LU with Markowitz-style complete
 * pivoting, the strategy of the solver the compiled reference actually runs (Sparse 1.3,
 * xt_base/sparse/spfactor.cc:286-423: among entries passing the relative threshold
 * |a| >= pivrel*max|column|, take the smallest (r-1)(c-1), singletons first).  This is a
 * robustness restatement, not an op-for-op one: results agree with oracle/_ref to roundoff. */
static int dense_factor(wo_ckt *c)
{
    int n = c->n;
    if (!c->dense) {
        c->dense = xcalloc((size_t)n*n, sizeof(double));
        c->dpiv = xcalloc(4*(size_t)n, sizeof(int));
    }
    double *M = c->dense;
    int *rp = c->dpiv, *cp = c->dpiv + n, *ract = c->dpiv + 2*n, *cact = c->dpiv + 3*n;
    memset(M, 0, (size_t)n*n*sizeof(double));
    for (int j = 0; j < n; j++)
        for (int p = c->Ap[j]; p < c->Ap[j+1]; p++)
            M[(size_t)c->Ai[p]*n + j] = c->Ax[p];
    for (int i = 0; i < n; i++) { ract[i] = 1; cact[i] = 1; }
    int *rc = xcalloc(n, sizeof(int)), *cc = xcalloc(n, sizeof(int));
    for (int k = 0; k < n; k++) {
        for (int i = 0; i < n; i++) { rc[i] = 0; cc[i] = 0; }
        for (int i = 0; i < n; i++) {
            if (!ract[i]) continue;
            for (int j = 0; j < n; j++)
                if (cact[j] && M[(size_t)i*n + j] != 0.0) { rc[i]++; cc[j]++; }
        }
        long best = -1; int bi = -1, bj = -1; double bmag = 0.0;
        for (int j = 0; j < n; j++) {
            if (!cact[j]) continue;
            double colmax = 0.0;
            for (int i = 0; i < n; i++)
                if (ract[i]) { double a = fabs(M[(size_t)i*n + j]); if (a > colmax) colmax = a; }
            if (colmax == 0.0 || !isfinite(colmax)) continue;
            for (int i = 0; i < n; i++) {
                if (!ract[i]) continue;
                double a = fabs(M[(size_t)i*n + j]);
                if (a == 0.0 || a < c->opt.pivrel*colmax || a <= c->opt.pivtol) continue;
                long mk = (long)(rc[i]-1)*(cc[j]-1);
                if (best < 0 || mk < best || (mk == best && a > bmag)) {
                    best = mk; bi = i; bj = j; bmag = a;
                }
            }
        }
        if (bi < 0) { free(rc); free(cc); return E_SINGULAR; }
        rp[k] = bi; cp[k] = bj;
        ract[bi] = 0; cact[bj] = 0;
        double d = M[(size_t)bi*n + bj];
        for (int i = 0; i < n; i++) {
            if (!ract[i]) continue;
            double l = M[(size_t)i*n + bj];
            if (l == 0.0) continue;
            l /= d;
            M[(size_t)i*n + bj] = l;
            for (int j = 0; j < n; j++)
                if (cact[j]) M[(size_t)i*n + j] -= l*M[(size_t)bi*n + j];
        }
    }
    free(rc); free(cc);
    return OK;
}

static void dense_solve(wo_ckt *c, double *b /* n */)
{
    int n = c->n; double *M = c->dense;
    int *rp = c->dpiv, *cp = c->dpiv + n;
    double *x = xcalloc(n, sizeof(double));
    char *done = xcalloc(n, 1);
    /* forward: rows eliminated after step k carry multiplier M[i][cp[k]] */
    for (int k = 0; k < n; k++) {
        int pr = rp[k], pc = cp[k];
        done[pr] = 1;
        double bk = b[pr];
        if (bk != 0.0)
            for (int i = 0; i < n; i++)
                if (!done[i]) b[i] -= M[(size_t)i*n + pc]*bk;
    }
    /* backward */
    for (int k = n - 1; k >= 0; k--) {
        int pr = rp[k], pc = cp[k];
        double s = b[pr];
        for (int q = k + 1; q < n; q++)
            s -= M[(size_t)pr*n + cp[q]]*x[cp[q]];
        x[pc] = s/M[(size_t)pr*n + pc];
    }
    memcpy(b, x, n*sizeof(double));
    free(x); free(done);
}

static StaticLU *g_dummy;   /* silence unused warnings on some compilers */

static int ckt_factor(wo_ckt *c)
{
    if (c->have_order && c->have_order2 && !(c->mode & MODEDC)) {
        if (!c->lu_ready2) {
            c->slu2 = slu_symbolic(c->n, c->Ap, c->Ai, c->rowperm2, c->colperm2);
            c->lu_ready2 = 1;
        }
        return slu_refactor((StaticLU*)c->slu2, c->Ap, c->Ai, c->Ax, c->colperm2, c->work);
    }
    if (c->have_order) {
        if (!c->lu_ready) {
            StaticLU *f = slu_symbolic(c->n, c->Ap, c->Ai, c->rowperm, c->colperm);
            c->Lp = (int*)f;       /* stash */
            c->lu_ready = 1;
        }
        return slu_refactor((StaticLU*)c->Lp, c->Ap, c->Ai, c->Ax, c->colperm, c->work);
    }
    (void)g_dummy;
    return dense_factor(c);
}

static void ckt_solve_inner(wo_ckt *c);
static void ckt_solve(wo_ckt *c)
{
    static int ncall;
    const char *dn = getenv("WO_DUMP_SOLVE");
    if (dn && ++ncall == atoi(dn)) {
        FILE *f = fopen("/tmp/wo_dump.txt", "w");
        fprintf(f, "%d %d\n", c->n, c->nnz);
        for (int j = 0; j <= c->n; j++) fprintf(f, "%d\n", c->Ap[j]);
        for (int p = 0; p < c->nnz; p++) fprintf(f, "%d %.17g\n", c->Ai[p], c->Ax[p]);
        for (int i = 1; i <= c->n; i++) fprintf(f, "%.17g\n", c->rhs[i]);
        ckt_solve_inner(c);
        for (int i = 1; i <= c->n; i++) fprintf(f, "%.17g\n", c->rhs[i]);
        fclose(f);
        return;
    }
    ckt_solve_inner(c);
}
static void ckt_solve_inner(wo_ckt *c)
{
    /* spSolve(CKTrhs, CKTrhs): slot 0 is ground (spsolve.cc:173-245) */
    if (c->have_order && c->have_order2 && !(c->mode & MODEDC)) {
        double *y = xcalloc(c->n, sizeof(double));
        slu_solve((StaticLU*)c->slu2, c->rowperm2, c->colperm2, c->rhs + 1, y);
        free(y);
    }
    else if (c->have_order) {
        double *y = xcalloc(c->n, sizeof(double));
        slu_solve((StaticLU*)c->Lp, c->rowperm, c->colperm, c->rhs + 1, y);
        free(y);
    }
    else
        dense_solve(c, c->rhs + 1);
}

/* ------------------------------------------------------------------------- */
/* FPE check: fpe_check.h:51-71 (FPEcheck mode)                                */
/* ------------------------------------------------------------------------- */
static int check_fpe(int noerrret)
{
    int err = OK;
    if (!noerrret) {
        if (fetestexcept(FE_DIVBYZERO)) err = E_MATHDBZ;
        else if (fetestexcept(FE_OVERFLOW)) err = E_MATHOVF;
        else if (fetestexcept(FE_INVALID)) err = E_MATHINV;
    }
    feclearexcept(FE_ALL_EXCEPT);
    return err;
}

/* ------------------------------------------------------------------------- */
/* NIconvTest: niconv.cc:58-100; sCKT::convTest: ckt.cc:466-478                 */
/* ------------------------------------------------------------------------- */
static int NIconvTest(wo_ckt *c)
{
    for (int i = 1; i <= c->n; i++) {
        double cur = c->rhs[i], old = c->rhsOld[i];
        double acur = fabs(cur), aold = fabs(old), dd = fabs(cur - old);
        double rel = c->opt.reltol*SPMAX(aold, acur);
        double tol = rel + (c->ntype[i] == SP_VOLTAGE ? c->opt.vntol : c->opt.abstol);
        if (dd > tol) {
            if (getenv("WO_TRACE2")) fprintf(stderr, "  nonconv node %d type %d old %.17g cur %.17g\n", i, c->ntype[i], old, cur);
            return 1;
        }
    }
    /* device convTest in model-type order; only MOS has one among the supported kinds
     * (cap/res/src inherit the no-op). mosconv.cc:55-57: skipped if noncon > 0. */
    if (c->noncon == 0) {
        for (int o = 0; o < c->nord; o++)
            if (c->ord[o].kind == B200_DEV_MOS)
                if (mos_conv(c, &c->mos[c->ord[o].index])) { if (getenv("WO_TRACE2")) fprintf(stderr, "  nonconv mos %d\n", c->ord[o].index); c->noncon++; break; }
    }
    return (c->noncon != 0);
}

/* ------------------------------------------------------------------------- */
/* NIiter: niiter.cc:160-555                                                  */
/* ------------------------------------------------------------------------- */
static int NIiter(wo_ckt *c, int maxIter)
{
    int iterno = 0, ipass = 0;
    c->noncon = 0;
    check_fpe(1);
    if ((c->mode & MODEINITJCT) || (c->mode & MODEINITTRAN))
        c->niState |= NISHOULDREORDER;
    int error = ckt_load(c);
    if (!error) error = check_fpe(0);
    if (error) return error;

    for (;;) {
        check_fpe(1);
        iterno++;
        error = load_gmin(c);
        if (error) break;
        if (c->niState & NISHOULDREORDER) {
            /* spOrderAndFactor: with a frozen pivot sequence this is a refactor */
            error = ckt_factor(c);
            if (error) break;
            error = check_fpe(0);
            if (error) break;
            c->niState &= ~NISHOULDREORDER;
        }
        else {
            error = ckt_factor(c);
            if (error) {
                if (error == E_SINGULAR) {
                    c->niState |= NISHOULDREORDER;
                    if (iterno == 1) {
                        error = ckt_load(c);
                        if (error) break;
                        continue;
                    }
                }
                break;
            }
            error = check_fpe(0);
            if (error) break;
        }
        ckt_solve(c);
        error = check_fpe(0);
        if (error) break;

        double *tmp = c->rhsOld; c->rhsOld = c->rhs; c->rhs = tmp;
        c->rhs[0] = 0; c->rhsSpare[0] = 0; c->rhsOld[0] = 0;

        if (iterno > maxIter) { error = E_ITERLIM; break; }

        if (!c->opt.jjaccel || (c->mode & MODEDC)) {
            if (c->mode & (MODEINITFIX | MODEINITFLOAT)) {
                if (c->noncon == 0 && iterno != 1)
                    c->noncon = NIconvTest(c);
                else
                    c->noncon = 1;
            }
        }
        if (c->mode & MODEINITFLOAT) {
            if ((c->mode & MODEDC) && c->hadNodeset) {
                if (ipass) c->noncon = ipass;
                ipass = 0;
            }
            if (c->noncon == 0) { error = OK; break; }
            if ((c->mode & MODEDC) && iterno > 2 && c->opt.dcmu != 0.5) {
                double A = c->opt.dcmu + 0.5, B = 1.0 - A;
                for (int i = 1; i <= c->n; i++)
                    c->rhsOld[i] = A*c->rhsOld[i] + B*c->rhs[i];
            }
        }
        else if (c->mode & MODEINITPRED)
            c->mode = (c->mode & (~INITF)) | MODEINITFLOAT;
        else if (c->mode & MODEINITTRAN) {
            if (iterno <= 1) c->niState |= NISHOULDREORDER;
            c->mode = (c->mode & (~INITF)) | MODEINITFLOAT;
        }
        else if (c->mode & MODEINITSMSIG)
            c->mode = (c->mode & (~INITF)) | MODEINITFLOAT;
        else if (c->mode & MODEINITJCT) {
            c->mode = (c->mode & (~INITF)) | MODEINITFIX;
            c->niState |= NISHOULDREORDER;
        }
        else if (c->mode & MODEINITFIX) {
            if (c->noncon == 0)
                c->mode = (c->mode & (~INITF)) | MODEINITFLOAT;
            ipass = 1;
        }
        else { error = B200_E_PANIC; break; }
        c->noncon = 0;
        error = ckt_load(c);
        if (error) break;
        error = check_fpe(0);
        if (error) break;
        if (c->opt.jjaccel && c->noncon == 0 && (c->mode & MODEINITFLOAT) &&
                !(c->mode & (MODEDCOP | MODETRANOP | MODEINITTRAN))) {
            error = OK; break;
        }
    }
    c->numIter += iterno;
    if (!(c->mode & MODEDC) && (c->mode & MODETRAN))
        c->tranIter += iterno;
    return error;
}

/* ------------------------------------------------------------------------- */
/* DCOP: ckt.cc:1323-1362 (backup), 1492-1610 (dynamic_gmin), 1680-1776         */
/* (dynamic_src), 1859-2011 (op)                                              */
/* ------------------------------------------------------------------------- */
enum { DEV_SAVE, DEV_RESTORE, DEV_CLEAR };
static void mos_pack(const Mos *m, MosSave *s)
{
    s->mode = m->mode;
    double v[12] = { m->von, m->vdsat, m->cd, m->cbs, m->cbd, m->gm, m->gds, m->gmbs, m->gbd,
        m->gbs, m->capbd, m->capbs };
    memcpy(s->v, v, sizeof(v));
}
static void mos_unpack(Mos *m, const MosSave *s)
{
    m->mode = s->mode;
    m->von = s->v[0]; m->vdsat = s->v[1]; m->cd = s->v[2]; m->cbs = s->v[3]; m->cbd = s->v[4];
    m->gm = s->v[5]; m->gds = s->v[6]; m->gmbs = s->v[7]; m->gbd = s->v[8]; m->gbs = s->v[9];
    m->capbd = s->v[10]; m->capbs = s->v[11];
}

static void ckt_backup(wo_ckt *c, int bmd)
{
    /* device part: MOSdev::backup copies the instance struct (mos.cc:395-410,
     * mosdefs.h:386-402): SAVE stores, RESTORE restores if a save exists, CLEAR drops it */
    if (bmd == DEV_SAVE) {
        if (!c->mosSave) c->mosSave = xcalloc(c->nmos, sizeof(MosSave));
        for (int i = 0; i < c->nmos; i++) mos_pack(&c->mos[i], &c->mosSave[i]);
        c->haveSave = 1;
        if (!c->oldSol) c->oldSol = xcalloc(c->n + 1, sizeof(double));
        if (!c->oldState0) c->oldState0 = xcalloc(c->numStates + 1, sizeof(double));
        memcpy(c->oldSol, c->rhsOld, (c->n + 1)*sizeof(double));
        memcpy(c->oldState0, c->states[0], c->numStates*sizeof(double));
    }
    else if (bmd == DEV_RESTORE) {
        if (c->haveSave)
            for (int i = 0; i < c->nmos; i++) mos_unpack(&c->mos[i], &c->mosSave[i]);
        if (c->oldSol) memcpy(c->rhsOld, c->oldSol, (c->n + 1)*sizeof(double));
        if (c->oldState0) memcpy(c->states[0], c->oldState0, c->numStates*sizeof(double));
    }
    else {
        c->haveSave = 0;
        free(c->oldSol); c->oldSol = 0;
        free(c->oldState0); c->oldState0 = 0;
    }
}

static int is_math_or_iterlim(int e)
{
    return e == E_ITERLIM || e == E_MATHDBZ || e == E_MATHOVF || e == E_MATHUNF || e == E_MATHINV;
}

static int dynamic_gmin(wo_ckt *c, int firstmode, int continuemode, int iterlim)
{
    c->mode = firstmode;
    memset(c->states[0], 0, c->numStates*sizeof(double));
    const double gshunt = 0.0, Factor = 10.0;
    double factor = Factor, OldGmin = 1e-2;
    c->diagGmin = OldGmin/factor;
    double gtarget = SPMAX(c->opt.gmin, gshunt);
    if (gtarget < 1e-15) gtarget = 1e-15;
    int succeeded = 0, nonconv = OK;
    for (;;) {
        int iters = c->numIter;
        c->noncon = 1;
        nonconv = NIiter(c, c->opt.itl_gmin);
        iters = c->numIter - iters;
        if (nonconv) {
            if (!is_math_or_iterlim(nonconv)) {
                c->diagGmin = gshunt;
                ckt_backup(c, DEV_CLEAR);
                return nonconv;
            }
            if (getenv("WO_TRACE")) fprintf(stderr, "Gmin step failed, Gshunt=%g iters=%d E%d\n", c->diagGmin, iters, nonconv);
            if (factor < 1.00005) break;
            factor = sqrt(sqrt(factor));
            c->diagGmin = OldGmin/factor;
            ckt_backup(c, DEV_RESTORE);
        }
        else {
            if (getenv("WO_TRACE")) fprintf(stderr, "Gmin step succeeded, Gshunt=%g iters=%d\n", c->diagGmin, iters);
            c->mode = continuemode;
            if (c->diagGmin <= gtarget) { succeeded = 1; break; }
            ckt_backup(c, DEV_SAVE);
            if (iters <= c->opt.itl_gmin/4) {
                factor *= sqrt(factor);
                if (factor > Factor) factor = Factor;
            }
            else if (iters > 3*c->opt.itl_gmin/4)
                factor = sqrt(factor);
            OldGmin = c->diagGmin;
            if (c->diagGmin < factor*gtarget) {
                factor = c->diagGmin/gtarget;
                c->diagGmin = gtarget;
            }
            else
                c->diagGmin /= factor;
        }
    }
    c->diagGmin = gshunt;
    ckt_backup(c, DEV_CLEAR);
    if (succeeded) {
        nonconv = NIiter(c, iterlim);
        if (!nonconv) return 0;
    }
    return nonconv;
}

static int dynamic_src(wo_ckt *c, int firstmode, int continuemode, int iterlim)
{
    c->mode = firstmode;
    c->srcFact = 0.001;
    double raise = 0.001, last_raise = raise, ConvFact = 0;
    int nonconv = dynamic_gmin(c, firstmode, continuemode, iterlim);
    if (!nonconv) {
        ckt_backup(c, DEV_SAVE);
        c->srcFact = ConvFact + raise;
        do {
            int iters = c->numIter;
            nonconv = NIiter(c, c->opt.itl_src);
            iters = c->numIter - iters;
            c->mode = continuemode;
            if (getenv("WO_TRACE")) fprintf(stderr, "Source step %s, factor=%g iters=%d\n", nonconv ? "failed" : "succeeded", c->srcFact, iters);
            if (nonconv) {
                if (!is_math_or_iterlim(nonconv) && nonconv != E_SINGULAR) break;
                if (c->srcFact - ConvFact < 1e-8) break;
                raise = 0.1*last_raise;
                if (raise > 0.01) raise = 0.01;
                c->srcFact = ConvFact + raise;
                last_raise = raise;
                ckt_backup(c, DEV_RESTORE);
            }
            else {
                ConvFact = c->srcFact;
                ckt_backup(c, DEV_SAVE);
                c->srcFact = ConvFact + raise;
                last_raise = raise;
                if (iters <= c->opt.itl_src/4) raise *= 1.5;
                else if (iters > 3*c->opt.itl_src/4) raise *= 0.5;
            }
            if (c->srcFact > 1) c->srcFact = 1;
        } while (raise >= 1e-7 && ConvFact < 1);
    }
    ckt_backup(c, DEV_CLEAR);
    c->srcFact = 1;
    if (ConvFact != 1) return nonconv;
    return 0;
}

static int err_return(int err, int first)
{
    /* ckt.cc:1825-1852: true => stop trying */
    if (err == OK) return 1;
    if (err == E_SINGULAR) return 1;
    if (err == E_ITERLIM) return first ? 0 : 1;
    return 1;
}

static int ckt_op(wo_ckt *c, int firstmode, int continuemode, int iterlim)
{
    const b200_options *o = &c->opt;
    if (o->gminsteps < 0 && o->srcsteps < 0) {
        c->mode = firstmode;
        int err = NIiter(c, iterlim);
        if (!err) c->mode = continuemode;
        return err;
    }
    if (o->gminsteps <= 0 && o->srcsteps <= 0) {
        if (o->gminfirst) {
            if (o->gminsteps >= 0) {
                int err = dynamic_gmin(c, firstmode, continuemode, iterlim);
                if (err_return(err, 1)) return err;
            }
            if (o->srcsteps >= 0) {
                int err = dynamic_src(c, firstmode, continuemode, iterlim);
                if (err_return(err, 0)) return err;
            }
        }
        else {
            if (o->srcsteps >= 0) {
                int err = dynamic_src(c, firstmode, continuemode, iterlim);
                if (err_return(err, 1)) return err;
            }
            if (o->gminsteps >= 0) {
                int err = dynamic_gmin(c, firstmode, continuemode, iterlim);
                if (err_return(err, 0)) return err;
            }
        }
        return OK;
    }
    /* spice3-style fixed stepping (gminsteps/srcsteps > 0) is not restated */
    return B200_E_UNSUPP;
}

/* ------------------------------------------------------------------------- */
/* LTE: ckt.cc:2310-2463, mostrun.cc:52-65, captrun.cc:52-63                    */
/* ------------------------------------------------------------------------- */
static double true_time(double tmp, int order)
{
    if (order == 2) return sqrt(tmp);
    if (order == 1) return tmp;
    if (order == 3) return cbrt(tmp);
    if (order == 4) return sqrt(sqrt(tmp));
    return exp(log(tmp)/order);
}

static void terr(wo_ckt *c, int qcap, double *timeStep)
{
    static const double gearCoeff[] = { .5, .2222222222, .1363636364, .096, .07299270073, .05830903790 };
    static const double trapCoeff[] = { .5, .08333333333 };
    int ccap = qcap + 1;
    double y0 = fabs(c->states[0][ccap]), y1 = fabs(c->states[1][ccap]);
    double volttol = c->opt.abstol + c->opt.reltol*SPMAX(y0, y1);
    y0 = fabs(c->states[0][qcap]); y1 = fabs(c->states[1][qcap]);
    double chargetol = SPMAX(y0, y1);
    if (chargetol < c->opt.chgtol) chargetol = c->opt.chgtol;
    chargetol *= c->opt.reltol/c->delta;
    double tol = SPMAX(volttol, chargetol);
    double diff[8], deltmp[8];
    for (int i = c->order + 1; i >= 0; i--) diff[i] = c->states[i][qcap];
    for (int i = 0; i <= c->order; i++) deltmp[i] = c->deltaOld[i];
    int j = c->order;
    for (;;) {
        for (int i = 0; i <= j; i++) diff[i] = (diff[i] - diff[i+1])/deltmp[i];
        if (--j < 0) break;
        for (int i = 0; i <= j; i++) deltmp[i] = deltmp[i+1] + c->deltaOld[i];
    }
    double del;
    if (c->opt.method == 0) del = fabs(diff[0]*trapCoeff[c->order-1]);
    else del = fabs(diff[0]*gearCoeff[c->order-1]);
    del = c->opt.trtol*tol/SPMAX(c->opt.abstol, del);
    if (del < *timeStep) *timeStep = del;
}

static int ckt_trunc(wo_ckt *c, double *timeStep)
{
    double timetemp = DBL_MAX;
    for (int o = 0; o < c->nord; o++) {
        int i = c->ord[o].index;
        if (c->ord[o].kind == B200_DEV_MOS) {
            terr(c, c->mos[i].st + MS_qgs, &timetemp);
            terr(c, c->mos[i].st + MS_qgd, &timetemp);
            terr(c, c->mos[i].st + MS_qgb, &timetemp);
        }
        else if (c->ord[o].kind == B200_DEV_CAP)
            terr(c, c->cap[i].st, &timetemp);
        else if (c->ord[o].kind == B200_DEV_IND)
            terr(c, c->ind[i].st, &timetemp);       /* indtrun.cc:58 */
        else if (c->ord[o].kind == B200_DEV_JJ) {
            /* jjtrun.cc:52-64: only junctions without a supercurrent bound the step by LTE */
            if (c->jj[i].mod.ictype == 0 || c->jj[i].criti == 0.0)
                terr(c, c->jj[i].st + JS_voltage, &timetemp);
        }
        else if (c->ord[o].kind == B200_DEV_TJM) {
            if (c->tjm[i].mod.ictype == 0 || c->tjm[i].criti == 0.0)       /* tjmtrun.cc:52-64 */
                terr(c, c->tjm[i].st + TS_voltage, &timetemp);
        }
    }
    /* srctrun.cc:45-66: SIN sources bound the step through IFsinData::time_limit
     * (inptran.cc:2666-2680, uses dphimax = .2*pi, circuit.h:495) */
    {
        double time_step = HUGE_VAL;
        for (int i = 0; i < c->nsrc; i++) {
            Src *s = &c->src[i];
            if (s->fn == B200_FN_SIN && c->time >= s->parms[3]) {
                double per = (.2*M_PI)/(2*M_PI*s->parms[2]);
                if (time_step > per) time_step = per;
            }
            else if (s->fn == B200_FN_SPULSE && c->time >= s->parms[3]) {     /* inptran.cc:2861-2870 */
                double per = (.2*M_PI)*s->parms[2]/(2*M_PI);
                if (time_step > per) time_step = per;
            }
            else if (s->fn == B200_FN_SFFM) {                       /* inptran.cc:3253-3260 */
                double per = (.2*M_PI)/(2*M_PI*(s->parms[2] + s->parms[4]));
                if (time_step > per) time_step = per;
            }
            else if (s->fn == B200_FN_AM) {                         /* inptran.cc:3435-3442 */
                double per = (.2*M_PI)/(2*M_PI*(s->parms[3] + s->parms[2]));
                if (time_step > per) time_step = per;
            }
            else if (s->fn == B200_FN_EXP) {                        /* inptran.cc:3064-3082 */
                double t = c->time, f = (.2*M_PI)/(2*M_PI);
                const double *p = s->parms;
                if (t >= p[2] && t < p[2] + 5.0*p[3]) {
                    double per = p[3]*f*exp((t - p[2])/p[3]);
                    if (time_step > per) time_step = per;
                }
                if (t >= p[4] && t < p[4] + 5.0*p[5]) {
                    double per = p[5]*f*exp((t - p[4])/p[5]);
                    if (time_step > per) time_step = per;
                }
            }
        }
        if (time_step < HUGE_VAL) {
            double tmp = time_step;
            for (int i = 1; i < c->order; i++) time_step *= tmp;
        }
        if (timetemp > time_step) timetemp = time_step;
    }
    *timeStep = true_time(timetemp, c->order);
    return OK;
}

/* ------------------------------------------------------------------------- */
/* transient driver: analysis/tranan.cc                                       */
/* ------------------------------------------------------------------------- */
static void record_point(wo_ckt *c, double t, const double *v /* [n+1] */)
{
    if (c->npts == c->cappts) {
        c->cappts = c->cappts ? 2*c->cappts : 256;
        c->pt_t = realloc(c->pt_t, c->cappts*sizeof(double));
        c->pt_v = realloc(c->pt_v, (size_t)c->cappts*c->n*sizeof(double));
    }
    c->pt_t[c->npts] = t;
    memcpy(c->pt_v + (size_t)c->npts*c->n, v + 1, c->n*sizeof(double));
    c->npts++;
}

static void record_accept(wo_ckt *c)
{
    if (c->nacc == c->capacc) {
        c->capacc = c->capacc ? 2*c->capacc : 256;
        c->acc_t = realloc(c->acc_t, c->capacc*sizeof(double));
        c->acc_v = realloc(c->acc_v, (size_t)c->capacc*c->n*sizeof(double));
        c->acc_it = realloc(c->acc_it, c->capacc*sizeof(int));
    }
    c->acc_t[c->nacc] = c->time;
    memcpy(c->acc_v + (size_t)c->nacc*c->n, c->rhsOld + 1, c->n*sizeof(double));
    c->acc_it[c->nacc] = c->lastIters;
    c->nacc++;
}

static void inc_check(wo_ckt *c)
{
    /* tranan.cc:472-479 (single .tran segment) */
    c->t_check += c->t_step;
}

static int tran_interpolate(wo_ckt *c)
{
    /* tranan.cc:963-1116 */
    if (c->mode & MODEUIC) {
        if (c->t_firsttime) return OK;
        if (c->npts == 0 && c->t_start == 0.0) {
            record_point(c, c->t_check, c->rhsOld);
            inc_check(c);
            return OK;
        }
    }
    double delta = .5*c->t_delmin;
    double finalmax = c->t_stop + delta, finalmin = c->t_stop - delta;
    double time = c->time + delta;
    int degree = c->t_polydegree;
    if (degree > c->opt.maxord) degree = c->opt.maxord;
    double diff[4];
    for (; c->t_check <= time; inc_check(c)) {
        double c2 = c->deltaOld[0];
        double d1 = c->time - c->t_check;
        if (degree <= 1) {
            double d2 = d1 - c2;
            diff[0] = d2/(-c2);
            diff[1] = d1/(c2);
            for (int i = 1; i <= c->n; i++)
                c->rhs[i] = diff[0]*c->sols[0][i] + diff[1]*c->sols[1][i];
        }
        else if (degree == 2) {
            double c3 = c2 + c->deltaOld[1];
            double d2 = d1 - c2, d3 = d1 - c3;
            diff[0] = d2*d3/((c2)*(c3));
            diff[1] = d1*d3/((c2)*(c2-c3));
            diff[2] = d1*d2/((c3)*(c3-c2));
            for (int i = 1; i <= c->n; i++)
                c->rhs[i] = diff[0]*c->sols[0][i] + diff[1]*c->sols[1][i] + diff[2]*c->sols[2][i];
        }
        else {
            double c3 = c2 + c->deltaOld[1], c4 = c3 + c->deltaOld[2];
            double d2 = d1 - c2, d3 = d1 - c3, d4 = d1 - c4;
            diff[0] = d2*d3*d4/((-c2)*(c3)*(c4));
            diff[1] = d1*d3*d4/((c2)*(c2-c3)*(c2-c4));
            diff[2] = d1*d2*d4/((c3)*(c3-c2)*(c3-c4));
            diff[3] = d1*d2*d3/((c4)*(c4-c2)*(c4-c3));
            for (int i = 1; i <= c->n; i++)
                c->rhs[i] = diff[0]*c->sols[0][i] + diff[1]*c->sols[1][i] +
                    diff[2]*c->sols[2][i] + diff[3]*c->sols[3][i];
        }
        if (c->t_check < finalmax) {
            record_point(c, c->t_check, c->rhs);
            if (finalmin < c->t_check) return E_NOCHANGE;
        }
        else
            return E_NOCHANGE;
    }
    if (c->t_check >= finalmax) return E_NOCHANGE;
    return OK;
}

static int ckt_accept(wo_ckt *c)
{
    /* ckt.cc:355-383 (device accept: JJ only among the built models) */
    c->devMaxDelta = 0.0;
    if (c->njj) jj_accept(c);
    if (c->ntjm) tjm_accept(c);
    double *tmp = c->sols[7];
    for (int i = 7; i > 0; i--) c->sols[i] = c->sols[i-1];
    c->sols[0] = tmp;
    for (int i = 0; i <= c->n; i++) c->sols[0][i] = c->rhsOld[i];
    return OK;
}

static int tran_accept(wo_ckt *c, int *done)
{
    /* tranan.cc:483-554 */
    memcpy(c->rhsSpare + 1, c->rhsOld + 1, c->n*sizeof(double));
    ckt_accept(c);
    if (c->time > c->breaks[0])
        break_clr(c);
    c->accepted++;
    c->brk = 0;
    record_accept(c);
    int error = tran_interpolate(c);
    if (error == E_NOCHANGE)
        c->endit = 1;
    if (c->endit) { c->endit = 0; *done = 1; }
    return OK;
}

static void tran_setbp(wo_ckt *c)
{
    /* tranan.cc:558-654 (STEP_NORMAL) */
    c->brk = 0;
    if (c->time == 0.0 || c->breaks[0] > 0) {
        if ((c->time == c->breaks[0]) || (c->breaks[0] - c->time <= c->t_delmin)) {
            c->order = 1;
            c->t_ordcnt = c->t_spice3 ? 0 : 1;
            double ttmp = c->breaks[1] - c->breaks[0];
            if (c->saveDelta < ttmp) ttmp = c->saveDelta;
            if (!c->t_nocut && !c->opt.jjaccel) ttmp *= 0.1;
            double td1 = 0.1*c->saveDelta;
            if (ttmp < td1) ttmp = td1;
            if (ttmp < c->delta) c->delta = ttmp;
            ttmp = 2.0*c->t_delmin;
            if (c->delta < ttmp) c->delta = ttmp;
        }
        else if (c->time + 1.01*c->delta >= c->breaks[0]) {
            double del = c->breaks[0] - c->time;
            if (del > c->t_delmin) {
                c->saveDelta = c->delta;
                c->delta = del;
                c->brk = 1;
            }
        }
    }
}

#define DELTA_FACTOR 0.125

static int tran_step(wo_ckt *c)
{
    /* tranan.cc:658-953 */
    double intDevMaxDelta = c->devMaxDelta;
    for (;;) {
        double olddelta = c->delta;
        if (c->maxStep < c->t_delmax) c->t_delmax = c->maxStep;
        c->time += c->delta;
        c->deltaOld[0] = c->delta;
        NIcomCof(c);
        predict(c);
        c->trapCheck = c->opt.trapcheck && c->opt.method == 0 && c->order == 2;
        int niter = c->numIter;
        int cverr = NIiter(c, c->opt.itl4);
        double bound_step = c->devMaxDelta;
        if (bound_step == 0.0 || (intDevMaxDelta > 0.0 && intDevMaxDelta < bound_step))
            c->devMaxDelta = intDevMaxDelta;
        c->timePts++;
        c->lastIters = c->numIter - niter;
        if (getenv("WO_TRACE"))
            fprintf(stderr, "time=%12g delta=%12g iters= %d order= %2d\n", c->time, c->delta, c->lastIters, c->order);
        c->mode = (c->mode & MODEUIC) | MODETRAN | MODEINITPRED;
        if (c->t_firsttime) {
            for (int i = 0; i < c->numStates; i++) {
                c->states[2][i] = c->states[1][i];
                if (c->opt.maxord > 1) c->states[3][i] = c->states[1][i];
            }
        }
        if (cverr) {
            if (cverr != E_ITERLIM) return cverr;
            c->time -= c->delta;
            c->rejected++;
            c->iterCut++;
            c->delta *= DELTA_FACTOR;
            if (c->t_firsttime)
                c->mode = (c->mode & MODEUIC) | MODETRAN | MODEINITTRAN;
            c->order = 1;
            c->brk = 0;
            memcpy(c->rhsOld + 1, c->rhsSpare + 1, c->n*sizeof(double));
        }
        else {
            if (c->t_firsttime) { c->t_firsttime = 0; break; }
            if (c->opt.jjaccel) {
                c->delta = c->devMaxDelta;
                if (c->order < c->opt.maxord) c->order++;
                break;
            }
            double newd;
            ckt_trunc(c, &newd);
            if (getenv("WO_TRACE")) fprintf(stderr, "LTE delta=%12g\n", newd);
            if (c->trapBad) {
                c->time -= c->delta;
                c->rejected++;
                c->trapCut++;
                c->delta *= DELTA_FACTOR;
                c->order = 1;
                memcpy(c->rhsOld + 1, c->rhsSpare + 1, c->n*sizeof(double));
                c->t_ordcnt = c->t_spice3 ? 0 : 1;
                continue;
            }
            if (c->t_spice3) {
                if (newd > 2.0*c->delta) newd = 2.0*c->delta;
                if (c->devMaxDelta > 0.0 && c->devMaxDelta < newd) newd = c->devMaxDelta;
                if (newd > .9*c->delta) {
                    if (c->order == 1) {
                        if (c->t_ordcnt) c->t_ordcnt--;
                        else {
                            c->order++;
                            if (c->order > c->opt.maxord) c->order = c->opt.maxord;
                            ckt_trunc(c, &newd);
                            if (newd > 2.0*c->delta) newd = 2.0*c->delta;
                            if (c->devMaxDelta > 0.0 && c->devMaxDelta < newd) newd = c->devMaxDelta;
                            if (newd <= 1.05*c->delta) c->order = 1;
                        }
                    }
                    c->delta = newd;
                    if (c->delta > c->t_delmax) c->delta = c->t_delmax;
                    break;
                }
                else {
                    c->time -= c->delta;
                    c->rejected++;
                    c->delta = newd;
                    memcpy(c->rhsOld + 1, c->rhsSpare + 1, c->n*sizeof(double));
                }
            }
            else {
                if (newd > .5*c->delta) {
                    double twodelta = 2.0*c->delta;
                    if (c->order < c->opt.maxord) {
                        if (c->t_ordcnt) c->t_ordcnt--;
                        else {
                            c->order++;
                            double newdtmp;
                            ckt_trunc(c, &newdtmp);
                            if (getenv("WO_TRACE")) fprintf(stderr, "LTE delta=%12g order= %d\n", newdtmp, c->order);
                            if (newdtmp <= 1.05*SPMIN(newd, twodelta)) c->order--;
                            else newd = newdtmp;
                        }
                    }
                    if (newd > twodelta) newd = twodelta;
                    if (c->devMaxDelta > 0.0 && c->devMaxDelta < newd) newd = c->devMaxDelta;
                    if (newd > c->t_delmax) newd = c->t_delmax;
                    c->delta = newd;
                    break;
                }
                else {
                    c->time -= c->delta;
                    c->rejected++;
                    if (c->devMaxDelta > 0.0 && c->devMaxDelta < newd) newd = c->devMaxDelta;
                    if (newd > c->t_delmax) newd = c->t_delmax;
                    c->delta = newd;
                    c->brk = 0;
                    memcpy(c->rhsOld + 1, c->rhsSpare + 1, c->n*sizeof(double));
                }
            }
        }
        if (c->delta <= c->t_delmin) {
            if (olddelta > c->t_delmin) c->delta = c->t_delmin;
            else return E_TIMESTEP;
        }
    }
    return OK;
}

static int has_trunc_dev(const wo_ckt *c)
{
    /* tranan.cc:68-77: any model type with DV_TRUNC (mos.cc:347, cap.cc:147) */
    return (c->nmos > 0 || c->ncap > 0 || c->nind > 0 || c->njj > 0 || c->ntjm > 0);      /* ind.cc:128, jj.cc:269, tjm.cc:280 */
}

int wo_run_tran(wo_ckt *c)
{
    const b200_options *o = &c->opt;
    feclearexcept(FE_ALL_EXCEPT);
    /* temperature pass (ckt.cc:2297-2307) -- after any wo_set_param */
    for (int i = 0; i < c->nmos; i++) {
        b200_mos_model deckmod = c->mos[i].card;
        mos_temp(c, &c->mos[i], &deckmod);
    }
    for (int i = 0; i < c->njj; i++) jj_setup_instance(&c->jj[i]);
    for (int i = 0; i < c->ntjm; i++) tjm_setup_instance(&c->tjm[i]);
    c->initDelta = 0.0;
    preload(c);

    /* sTRANAN::init, tranan.cc:94-217 */
    c->mode = (c->uic ? MODEUIC : 0) | MODETRAN;       /* TRANmode */
    c->t_stop = c->tstop; c->t_step = c->tstep; c->t_start = c->tstart;
    c->maxStep = c->tmax;
    c->t_spice3 = o->spice3;
    c->t_nocut = !has_trunc_dev(c);
    if (c->maxStep == 0) {
        c->maxStep = (c->t_stop - c->t_start)/50;
        if (o->oldsteplim) { if (c->t_step > c->maxStep) c->maxStep = c->t_step; }
        else { if (c->t_step < c->maxStep) c->maxStep = 1.001*c->t_step; }
    }
    c->t_delmax = c->maxStep;
    c->t_polydegree = o->interplev;
    double delmin = o->delmin;
    if (!(delmin > 0)) {
        if (o->oldsteplim) {
            delmin = c->maxStep*1e-9;
            double T1 = c->t_stop*1e-12;
            if (delmin < T1) delmin = T1;
        }
        else
            delmin = c->maxStep*1e-6;
    }
    c->t_delmin = delmin;
    c->minBreak = (o->minbreak > 0) ? o->minbreak : delmin*50.0;

    /* tran_dcoperation, tranan.cc:321-436 */
    c->time = 0; c->delta = 0; c->brk = 0;
    free(c->brks); c->brks = 0; c->numBreaks = 0; c->numLat = 0;
    c->breaks[0] = c->breaks[1] = 0.0;
    break_set(c, 0.0);
    break_set(c, c->t_start);
    break_set(c, c->t_stop);
    for (int oi = 0; oi < c->nord; oi++)
        if (c->ord[oi].kind == B200_DEV_SRC)
            src_setup_func(c, &c->src[c->ord[oi].index], c->t_step, c->t_stop);
    c->t_dumpit = 0;
    c->t_firsttime = 1;
    c->t_check = c->t_start;

    /* ckt.cc:908-925 ic() */
    for (int i = 0; i <= c->n; i++) c->rhs[i] = 0;
    c->hadNodeset = 0;
    for (int i = 1; i <= c->n; i++) {
        if (c->ns_given[i]) { c->hadNodeset = 1; c->rhs[i] = c->ns_val[i]; }
        if (c->ic_given[i]) c->rhs[i] = c->ic_val[i];
    }
    int error;
    if (c->mode & MODEUIC) {
        /* setic(): MOS getic (mosic.cc) fills missing device ICs from node ICs in CKTrhs */
        for (int i = 0; i < c->nmos; i++) {
            Mos *m = &c->mos[i];
            if (!GIVEN(m->in.icvbs)) m->icVBS = c->rhs[m->b] - c->rhs[m->s];
            if (!GIVEN(m->in.icvds)) m->icVDS = c->rhs[m->d] - c->rhs[m->s];
            if (!GIVEN(m->in.icvgs)) m->icVGS = c->rhs[m->g] - c->rhs[m->s];
        }
        for (int i = 0; i < c->ncap; i++) {
            /* capic.cc:52-71: if no ic given use node ICs */
            Cap *p = &c->cap[i];
            if (!p->icGiven) p->ic = c->rhs[p->pos] - c->rhs[p->neg];
        }
        for (int i = 0; i < c->nind; i++)
            if (c->ind[i].icGiven) c->rhs[c->ind[i].br] = c->ind[i].ic;      /* indic.cc:62-63 */
        if (c->njj) jj_getic(c);
        if (c->ntjm) tjm_getic(c);
        double *tmp = c->rhs; c->rhs = c->rhsOld; c->rhsOld = tmp;
        c->rhs[0] = 0.0; c->rhsOld[0] = 0.0;
        c->mode = MODEUIC | MODETRANOP | MODEINITJCT;
        error = ckt_load(c);
    }
    else
    {
        error = ckt_op(c, MODETRANOP | MODEINITJCT, MODETRANOP | MODEINITFLOAT, o->itl1);
        /* tranan.cc:378-389: what was computed on the phase nodes is phase; their voltage is 0 */
        if (!error && c->phase)
            for (int i = 1; i <= c->n; i++) if (c->phase[i]) c->rhsOld[i] = 0.0;
    }
    if (error) { return error; }

    for (int i = 1; i <= o->maxord + 1; i++)
        memcpy(c->states[i], c->states[0], c->numStates*sizeof(double));
    c->timePts++;
    c->order = 1;
    if (c->initDelta > 0)                              /* tranan.cc:416-422 */
        c->delta = c->initDelta;
    else {
        c->delta = SPMIN(c->t_delmax, c->t_step);
        if (!c->t_nocut) c->delta *= 0.1;
    }
    for (int i = 0; i < 7; i++) c->deltaOld[i] = c->delta;
    c->saveDelta = c->delta;
    c->mode = (c->mode & MODEUIC) | MODETRAN | MODEINITTRAN;
    c->ag[0] = c->ag[1] = 0;
    memcpy(c->states[1], c->states[0], c->numStates*sizeof(double));
    if (c->time >= c->t_start) c->t_dumpit = 1;

    int done = 0;
    for (;;) {
        error = tran_accept(c, &done);
        if (error || done) break;
        tran_setbp(c);
        for (int i = 5; i >= 0; i--) c->deltaOld[i+1] = c->deltaOld[i];
        c->deltaOld[0] = c->delta;
        double *temp = c->states[o->maxord + 1];
        for (int i = o->maxord; i >= 0; i--) c->states[i+1] = c->states[i];
        c->states[0] = temp;
        error = tran_step(c);
        if (error) break;
    }
    return error;
}

/* DCOanalysis::anFunc, analysis/dcoan.cc:100-121: ic(); op(MODEDCOP|INITJCT, MODEDCOP|INITFLOAT,
 * dcMaxIter); [INITSMSIG load: AC small-signal data only]; dump(0.0) = CKTrhsOld */
int wo_run_op(wo_ckt *c)
{
    const b200_options *o = &c->opt;
    feclearexcept(FE_ALL_EXCEPT);
    for (int i = 0; i < c->nmos; i++) {
        b200_mos_model deckmod = c->mos[i].card;
        mos_temp(c, &c->mos[i], &deckmod);
    }
    preload(c);
    c->time = 0; c->delta = 0;
    /* initTranFuncs(0.0, 0.0) */
    for (int oi = 0; oi < c->nord; oi++)
        if (c->ord[oi].kind == B200_DEV_SRC)
            src_setup_func(c, &c->src[c->ord[oi].index], 0.0, 0.0);
    for (int i = 0; i <= c->n; i++) c->rhs[i] = 0;
    c->hadNodeset = 0;
    for (int i = 1; i <= c->n; i++) {
        if (c->ns_given[i]) { c->hadNodeset = 1; c->rhs[i] = c->ns_val[i]; }
        if (c->ic_given[i]) c->rhs[i] = c->ic_val[i];
    }
    c->mode = 0;
    int error = ckt_op(c, MODEDCOP | MODEINITJCT, MODEDCOP | MODEINITFLOAT, o->itl1);
    if (error) return error;
    record_point(c, 0.0, c->rhsOld);
    return OK;
}

int wo_num_points(const wo_ckt *c) { return c->npts; }
void wo_get_points(const wo_ckt *c, double *times, double *vals)
{
    if (times) memcpy(times, c->pt_t, c->npts*sizeof(double));
    if (vals) memcpy(vals, c->pt_v, (size_t)c->npts*c->n*sizeof(double));
}
void wo_get_counters(const wo_ckt *c, int *k)
{
    k[B200_C_STATUS] = 0; k[B200_C_ACCEPTED] = c->accepted; k[B200_C_REJECTED] = c->rejected;
    k[B200_C_NUMITER] = c->numIter; k[B200_C_TIMEPTS] = c->timePts; k[B200_C_TRANITER] = c->tranIter;
    k[B200_C_ITERCUT] = c->iterCut; k[B200_C_TRAPCUT] = c->trapCut; k[B200_C_NPOINTS] = c->npts;
}
int wo_num_accepted(const wo_ckt *c) { return c->nacc; }
void wo_get_accepted(const wo_ckt *c, double *times, double *vals, int *iters)
{
    if (times) memcpy(times, c->acc_t, c->nacc*sizeof(double));
    if (vals) memcpy(vals, c->acc_v, (size_t)c->nacc*c->n*sizeof(double));
    if (iters) memcpy(iters, c->acc_it, c->nacc*sizeof(int));
}

int wo_debug_load(wo_ckt *c, int mode, double time, double delta, int order,
    const double *rhsold, double *Ax, double *rhs)
{
    for (int i = 0; i < c->nmos; i++) {
        b200_mos_model deckmod = c->mos[i].card;
        mos_temp(c, &c->mos[i], &deckmod);
    }
    preload(c);
    c->mode = mode; c->time = time; c->delta = delta; c->order = order;
    for (int i = 0; i < 7; i++) c->deltaOld[i] = delta;
    if (mode & MODETRAN) { NIcomCof(c); predict(c); }
    if (rhsold) memcpy(c->rhsOld, rhsold, (c->n + 1)*sizeof(double));
    int err = ckt_load(c);
    if (!err) err = load_gmin(c);
    if (Ax) memcpy(Ax, c->Ax, c->nnz*sizeof(double));
    if (rhs) memcpy(rhs, c->rhs, (c->n + 1)*sizeof(double));
    return err;
}

void wo_destroy(wo_ckt *c)
{
    if (!c) return;
    /* test infrastructure: process-lifetime leaks of the small tables are acceptable,
     * the large arrays are released */
    free(c->Ax); free(c->Ainit); free(c->Ap); free(c->Ai); free(c->dense); free(c->dpiv);
    free(c->rhs); free(c->rhsOld); free(c->rhsSpare);
    for (int i = 0; i < 8; i++) { free(c->sols[i]); free(c->states[i]); }
    free(c->pt_t); free(c->pt_v); free(c->acc_t); free(c->acc_v); free(c->acc_it);
    free(c->mos); free(c->cap); free(c->res); free(c->src); free(c->ord); free(c->ind); free(c->mut);
    free(c);
}
