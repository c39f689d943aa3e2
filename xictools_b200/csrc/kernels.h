// kernels.h -- kernel-side views and launch wrappers shared by kernels.cu and capi.cu
#pragma once
#include <cuda_runtime.h>
#include "device_state.h"

#define SOLVE_WARPS 4

struct MosCard;
struct CtlShared;

struct RawParams { ParamRef p[MR_NRAW]; };
struct MosParamsOut { double *base[MP_NPARAMS]; int ts[MP_NPARAMS]; };

// per-trial request arrays written by the control kernel for k_vec / k_statevec
struct CtlAux {
    int *vecops, *sols_base, *vo_stbase, *ip_n, *ip_out0, *ip_degree;
    double *ip_d;       // [5][ldt]: time, c2, c3, c4, tcheck0
};

struct SolveView {
    const int *lu_src;          // [nnzLU] Ax index or -1
    const double *Ainit;        // [nnz]
    const int *cg_ptr, *cg_idx; // gather lists per Ax entry
    const int *rg_ptr, *rg_idx; // gather lists per rhs row (1-based rows)
    const int *diag_node;       // [nnz] node number if the entry is (node,node), else 0
    const int *ntype;           // [n+1]
    const char *ic_given, *ns_given;
    const double *ic_val, *ns_val;
    const int *yperm_in, *yperm_out;
    int fac_nrounds, fwd_nrounds, bwd_nrounds;
    const LuRound *fac_rounds, *fwd_rounds, *bwd_rounds;
    const LuOp *fac_ops, *fwd_ops, *bwd_ops;
    const char *fac_sync, *fwd_sync, *bwd_sync;     // 1: round starts a new level
    const int *fac_fin_diag, *fac_fin_lbeg, *fac_fin_lend;
    const int *fwd_fin_row, *bwd_fin_row, *bwd_fin_diag;
};

void launch_mos_temp(cudaStream_t st, int n, int ntrials, int ldt, const MosCard *cards,
    const RawParams &raw, const MosParamsOut &out);
void launch_ctl(cudaStream_t st, const Consts &K, const HotView &hot, TrialCtl *ctl,
    const CtlShared &sh, const CtlAux &aux, int phase);
void launch_loads(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const MosView &M, const CapView &Cp, const SrcView &S, const double *rhsOld);
void launch_devconv(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const MosView &M, const double *rhs);
void launch_trunc(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    int nq, const int *qslots);
int  solve_smem_bytes(const Consts &K);
int  launch_solve(cudaStream_t st, const Consts &K, const HotView &hot, const VecView &V,
    const SolveView &Sv, const double *Xold, double *Xnew);
void launch_vec(cudaStream_t st, const Consts &K, const VecView &V, const CtlAux &aux,
    const MosView &M, double *rhsOld);
void launch_count_active(cudaStream_t st, int ntrials, const int *active, int *count);
void launch_fill(cudaStream_t st, double *p, size_t n, double v);
void launch_gather_waves(cudaStream_t st, int n, int ldt, int ntrials, int npts_max,
    const double *out, const int *npts, int nsel, const int *sel, double *waves);
void launch_export_counters(cudaStream_t st, int ntrials, const TrialCtl *ctl, int *counters,
    int *npts);
