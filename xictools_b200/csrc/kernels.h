// kernels.h -- kernel-side views and launch wrappers shared by kernels.cu and capi.cu
#pragma once
#include <cuda_runtime.h>
#include "device_state.h"

// contribution (s) of trial (t): [NC][ldt], trial fastest -- a warp of the load kernels writes one
// 256-B line per stamp, a warp of k_assemble reads one per term
#define CADDR(K, s, t) ((size_t)(s)*(size_t)(K).ldt + (size_t)(t))

struct MosCard;
struct CtlShared;

struct RawParams { ParamRef p[MR_NRAW]; };
struct MosParamsOut { double *base[MP_NPARAMS]; int ts[MP_NPARAMS]; };

// per-trial request arrays written by the control kernel for k_vec / k_statevec
struct CtlAux {
    int *vecops, *sols_base, *vo_stbase, *ip_n, *ip_out0, *ip_degree;
    double *ip_d;       // [5][ldt]: time, c2, c3, c4, tcheck0
};

// ---- the solve of one tick: k_assemble -> k_lu -> k_finish (solve.cu) ----
struct SolveView {
    const int *ntype;           // [n+1]
    const char *ic_given, *ns_given;
    const double *ic_val, *ns_val;
    const int *yperm_out;       // x[colperm[k]+1] = y[k]
    const AsmRun *runs;         // [img] assembly run of every image slot
    const int *tile_order;      // [ntiles] tiles of ASM_TILE_SLOTS slots, most terms first
    int ntiles;
    int img;                    // doubles per trial image (even: rows are 16-byte granular)
    int nlu;                    // LU values; the rhs / y / x part of the image starts here
    int npad;                   // row pitch of ysol (n rounded up to a multiple of 32)
    const unsigned *rows;       // task program, [nrows][LU_LANES] (b200_internal.h)
    const unsigned short *rounds;
    int nrows, nrounds;
    int cls;                    // which solving trials this program serves: -1 all, 1 DC modes, 0 transient
    double *image;              // [ldt][img] assembled A | b of every trial, trial-major
    double *ysol;               // [ldt][npad] solutions in permuted column order
    int *lu_bad;                // [ldt] 1: a pivot was zero or not finite
    int *fin_flags;             // [ldt] flags of the chunked k_finish (zero between solves)
    int *live;                  // [ldt] trials that take a solve this tick (built by k_assemble)
    int *ctr;                   // [0] number of live trials, [1] work counter of the k_lu warps
    int nsm;                    // SMs of the device (grid sizing)
    int lu_team;                // warps per resident image in k_lu_team (1: k_lu)
    int lu_warps, lu_grid;      // warps (= trials in flight) per CTA, CTAs of the persistent k_lu
    int lu_smem;                // dynamic shared memory of k_lu
    int prog_in_smem;           // 1: every CTA stages the program rows in shared memory
    int wide;                   // 1: wide program (b200_internal.h), image stays in global memory
    const AsmChunk *long_chunks; int nlong; int NC;     // long assembly runs (b200_internal.h), chunk c -> row NC + c
    const int *round_row;       // wide: first program row of every round
    const int *level_ptr;       // wide: [nlevels+1] first round of every level
    int nlevels;
};

void launch_res_ainit(cudaStream_t st, int ne, int ldt, const int *cslot, const double *base, const int *ptr,
    const int *res, const signed char *sign, ParamRef r, const double *m, double resmin, double *contrib);
void launch_jjaccel_gate(cudaStream_t st, const Consts &K, const HotView &hot);
void launch_mos_temp(cudaStream_t st, int n, int ntrials, int ldt, const MosCard *cards,
    const RawParams &raw, const MosParamsOut &out);
struct JJTempIn {
    int n, nmodels;
    const int *model;
    ParamRef ics;                       // NaN = not given
    const double *area0, *vshunt, *vless, *vmore, *tcf;
    double criti[JJ_MAXMODELS], cap[JJ_MAXMODELS], cmu[JJ_MAXMODELS], gmu[JJ_MAXMODELS], r0[JJ_MAXMODELS],
        rn[JJ_MAXMODELS], icfactor[JJ_MAXMODELS], delv[JJ_MAXMODELS];
    int rtype[JJ_MAXMODELS];
    double *out[JP_NPARAMS]; int ts;    // out[k][i*ld + t*ts]
};
void launch_jj_temp(cudaStream_t st, const JJTempIn &in, int ntrials, int ldt);
struct TJMTempIn {
    int n;
    const int *model;
    ParamRef ics;                       // NaN = not given
    const double *area0, *vshunt, *tcf, *alphaN;
    double criti[JJ_MAXMODELS], cap[JJ_MAXMODELS], cmu[JJ_MAXMODELS], gmu[JJ_MAXMODELS], r0[JJ_MAXMODELS],
        rsint[JJ_MAXMODELS];
    int rtype[JJ_MAXMODELS];
    double *out[TP_NPARAMS]; int ts;
};
void launch_tjm_temp(cudaStream_t st, const TJMTempIn &in, int ntrials, int ldt);
void launch_ctl(cudaStream_t st, const Consts &K, const HotView &hot, TrialCtl *ctl,
    const CtlShared &sh, const CtlAux &aux, int phase);
void launch_loads(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const MosView &M, const CapView &Cp, const SrcView &S, const double *rhsOld);
void launch_devconv(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const MosView &M, const double *rhs);
// per device, at batch creation / parameter upload: dynamic shared-memory opt-ins and launch shapes
int  loads_configure(const MosView &M);
// sizes the k_lu launch (warps per CTA, program placement) for the current device;
// B200_E_UNSUPP if one trial image does not fit shared memory
int  solve_configure(SolveView &Sv);
int  launch_solve(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const SolveView &Sv, const double *Xold, double *Xnew);
void launch_vec(cudaStream_t st, const Consts &K, const VecView &V, const CtlAux &aux,
    const MosView &M, double *rhsOld);
void launch_count_active(cudaStream_t st, int ntrials, const int *active, int *count);
void launch_fill(cudaStream_t st, double *p, size_t n, double v);
void launch_fill_const(cudaStream_t st, const Consts &K, double *contrib, int nconst, const int *idx,
    const double *val);
void launch_gather_waves(cudaStream_t st, int n, int ldt, int ntrials, int npts_max,
    const double *out, const int *npts, int nsel, const int *sel, double *waves);
void launch_export_counters(cudaStream_t st, int ntrials, const TrialCtl *ctl, int *counters,
    int *npts);
