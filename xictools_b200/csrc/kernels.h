// kernels.h -- kernel-side views and launch wrappers shared by kernels.cu and capi.cu
#pragma once
#include <cuda_runtime.h>
#include "device_state.h"

#define SOLVE_TB_MAX 4      // trials per CTA of k_solve: 4 (one 32-B sector per value), 2 or 1
                            // (Consts::tb), the largest whose LU image fits shared memory
#define SEG_NBUF 3           // streamed-assembly ring depth (chunks in flight)
// contribution (s) of trial (t): tile layout, TB adjacent trials of one slot = one 32-B sector
#define CADDR(K, s, t) ((((size_t)((t) >> (K).tbshift)*(size_t)(K).NC + (size_t)(s)) << (K).tbshift) + ((t) & ((K).tb - 1)))

struct MosCard;
struct CtlShared;

struct RawParams { ParamRef p[MR_NRAW]; };
struct MosParamsOut { double *base[MP_NPARAMS]; int ts[MP_NPARAMS]; };

// per-trial request arrays written by the control kernel for k_vec / k_statevec
struct CtlAux {
    int *vecops, *sols_base, *vo_stbase, *ip_n, *ip_out0, *ip_degree;
    double *ip_d;       // [5][ldt]: time, c2, c3, c4, tcheck0
};

#define RING_HROWS_PAD 28        // x3 >= (RING_NSLOT-1)*RING_HROWS of k_solve: the op-stream prefetch runs ahead
#define SOLVE_MAXCH 1024         // chunks of the streamed assembly (table cached in shared memory)
struct SolveView {
    const int *ntype;           // [n+1]
    const char *ic_given, *ns_given;
    const double *ic_val, *ns_val;
    const int *yperm_in, *yperm_out;
    const int *strm; const int *strm_round; int strm_nrounds;   // phase-2 op stream
    const SegOp *seg_ops; const int *seg_chunk; int nchunks, seg_maxsteps;   // streamed assembly
    long long *prof;            // optional per-warp phase cycle counters (B200_SOLVE_PROF=1)
};


void launch_mos_temp(cudaStream_t st, int n, int ntrials, int ldt, const MosCard *cards,
    const RawParams &raw, const MosParamsOut &out);
void launch_ctl(cudaStream_t st, const Consts &K, const HotView &hot, TrialCtl *ctl,
    const CtlShared &sh, const CtlAux &aux, int phase);
void launch_loads(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const MosView &M, const CapView &Cp, const SrcView &S, const double *rhsOld);
void launch_devconv(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const MosView &M, const double *rhs);
int  solve_smem_bytes(const Consts &K, int seg_maxsteps);   // for K.tb
int  solve_pick_tb(Consts &K, int seg_maxsteps);               // sets K.tb; 0 if nothing fits
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
