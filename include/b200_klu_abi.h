/*
 * b200_klu_abi.h -- the KLU plug-in C ABI as WRspice binds it, served by libb200spice.so.
 *
 * WRspice finds its fast sparse solver at run time: KLUif::find_klu
 * (wrspice/src/sparse/kluif.cc:117-266) dlopen()s the file named by $XT_KLU_PATH (or
 * <startup>/klu_wr.so), requires `SuiteSparse_version` to report the compile-time version
 * 4.4.6 (kluif.cc:150-166, KLU/Makefile.in:11), and resolves the 18 functions whose
 * prototypes are wrspice/include/kluif.h:196-226.  KLUmatrix (include/klumatrix.h:53-82,
 * kluif.cc:268-681), the spMatlabMatrix subclass that spMatrixFrame installs
 * (src/sparse/spspice.cc:83-97), then calls klu_defaults / klu_analyze / klu_factor /
 * klu_refactor / klu_solve / klu_free_* through those pointers and reads
 * Common.{ordering,status,singular_col} and Symbolic->{P,Q} directly
 * (kluif.cc:276,573,589,656-680), so the leading layout of those structs is part of the ABI.
 *
 * SuiteSparse itself is NOT in the reference tree (KLU/ holds only a Makefile that downloads
 * 4.4.6).  The declarations below restate the PUBLIC interface of KLU 1.3.x as shipped in
 * SuiteSparse 4.4.6 (klu.h: klu_common, klu_symbolic, status codes) -- field order and types of
 * the published header, nothing else.  klu_numeric is opaque to the caller.
 *
 * Pointing XT_KLU_PATH at libb200spice.so makes an unmodified WRspice factor and solve on the
 * B200 (one trial per call -- the batched entry points are b200_klu_* in b200spice.h; this
 * face exists so the existing hook loads the library without a source change).  No CPU
 * fallback: without a CUDA device klu_factor fails with KLU_INVALID and WRspice reports it.
 */
#ifndef B200_KLU_ABI_H
#define B200_KLU_ABI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SUITESPARSE_MAIN_VERSION   4
#define SUITESPARSE_SUB_VERSION    4
#define SUITESPARSE_SUBSUB_VERSION 6

#define KLU_OK             0
#define KLU_SINGULAR       1      /* status > 0 is a warning */
#define KLU_OUT_OF_MEMORY (-2)
#define KLU_INVALID       (-3)
#define KLU_TOO_LARGE     (-4)

typedef struct klu_common_struct {
    /* parameters */
    double tol;               /* pivot tolerance for diagonal preference (0.001) */
    double memgrow, initmem_amd, initmem, maxwork;
    int btf;                  /* use BTF pre-ordering */
    int ordering;             /* 0 AMD, 1 COLAMD, 2 user P and Q, 3 user function */
    int scale;                /* row scaling: -1 none (no error check), 0 none, 1 sum, 2 max */
    int (*user_order)(int, int *, int *, int *, struct klu_common_struct *);
    void *user_data;
    int halt_if_singular;
    /* statistics */
    int status;               /* KLU_OK if OK, < 0 if error */
    int nrealloc;
    int structural_rank;
    int numerical_rank;
    int singular_col;         /* n if the matrix is not singular */
    int noffdiag;
    double flops, rcond, condest, rgrowth, work;
    size_t memusage, mempeak;
} klu_common;

typedef struct {
    double symmetry, est_flops, lnz, unz;
    double *Lnz;
    int n, nz;
    int *P;                   /* size n: position -> original row */
    int *Q;                   /* size n: position -> original column */
    int *R;
    int nzoff, nblocks, maxblock, ordering, do_btf, structural_rank;
} klu_symbolic;

typedef struct klu_numeric_struct klu_numeric;      /* opaque here */

int SuiteSparse_version(int version[3]);

int klu_defaults(klu_common *Common);
klu_symbolic *klu_analyze(int n, int *Ap, int *Ai, klu_common *Common);
klu_numeric *klu_factor(int *Ap, int *Ai, double *Ax, klu_symbolic *Symbolic, klu_common *Common);
int klu_refactor(int *Ap, int *Ai, double *Ax, klu_symbolic *Symbolic, klu_numeric *Numeric,
                 klu_common *Common);
int klu_solve(klu_symbolic *Symbolic, klu_numeric *Numeric, int ldim, int nrhs, double *B,
              klu_common *Common);
int klu_tsolve(klu_symbolic *Symbolic, klu_numeric *Numeric, int ldim, int nrhs, double *B,
               klu_common *Common);
void klu_free_symbolic(klu_symbolic **Symbolic, klu_common *Common);
void klu_free_numeric(klu_numeric **Numeric, klu_common *Common);

/* Complex (AC analysis) and long-double (`extprec`) variants: resolvable, so find_klu accepts
 * the library, but they are outside the batched real-FP64 path and fail with KLU_INVALID
 * (WRspice then reports spPANIC; run such analyses with the `noklu` option). */
klu_numeric *klu_z_factor(int *, int *, double *, klu_symbolic *, klu_common *);
klu_numeric *klu_ld_factor(int *, int *, long double *, klu_symbolic *, klu_common *);
int klu_z_refactor(int *, int *, double *, klu_symbolic *, klu_numeric *, klu_common *);
int klu_ld_refactor(int *, int *, long double *, klu_symbolic *, klu_numeric *, klu_common *);
int klu_z_solve(klu_symbolic *, klu_numeric *, int, int, double *, klu_common *);
int klu_ld_solve(klu_symbolic *, klu_numeric *, int, int, long double *, klu_common *);
int klu_z_tsolve(klu_symbolic *, klu_numeric *, int, int, double *, int, klu_common *);
int klu_ld_tsolve(klu_symbolic *, klu_numeric *, int, int, long double *, klu_common *);
void klu_z_free_numeric(klu_numeric **, klu_common *);
void klu_ld_free_numeric(klu_numeric **, klu_common *);

#ifdef __cplusplus
}
#endif
#endif
