// klu_abi.cpp -- the 18 klu_* symbols + SuiteSparse_version that WRspice's plug-in loader
// resolves (wrspice/src/sparse/kluif.cc:117-266, prototypes include/kluif.h:196-226), served
// over the batched seam (b200_klu_* in klu_seam.cu).  XT_KLU_PATH=<this library> makes the
// reference's own KLUmatrix (kluif.cc:268-681) factor and solve on the B200, one trial per
// call.  See include/b200_klu_abi.h for what of the struct layouts is contractual.
//
// Call pattern of KLUmatrix (kluif.cc:569-628):
//   factor()   = klu_free_symbolic, klu_analyze(n, Ap, Ai), klu_free_numeric, klu_factor(Ap, Ai, Ax)
//   refactor() = klu_refactor(Ap, Ai, Ax, Symbolic, Numeric)       values changed, pattern and
//                                                                  pivot sequence kept
//   solve(rhs) = klu_solve(Symbolic, Numeric, n, 1, rhs)           rhs overwritten by x
// klu_analyze sees the pattern only; the batched seam orders on the first matrix's values, so
// the ordering is made in klu_factor and Symbolic->P/Q are filled there.  The device program
// is one fused refactor + forward + backward pass and the LU never travels back to the host:
// klu_refactor runs it on a zero right-hand side to report a singular pivot where the
// reference looks for it, klu_solve runs it again with the real one.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "b200spice.h"
#include "b200_klu_abi.h"

struct klu_numeric_struct {
    b200_klu *k = nullptr;
    int n = 0, nnz = 0;
    std::vector<double> Ax;          // the values of the last factor / refactor
};

namespace {

bool trace_on() { static int t = getenv("B200_KLU_TRACE") ? 1 : 0; return t != 0; }

int fail(klu_common *c, int status)
{
    if (trace_on()) fprintf(stderr, "b200 klu: status %d\n", status);
    if (c) c->status = status;
    return 0;
}

int device_from_env()
{
    const char *s = getenv("B200_KLU_DEVICE");
    return s ? atoi(s) : 0;
}

} // namespace

extern "C" {

int SuiteSparse_version(int version[3])
{
    if (version) {
        version[0] = SUITESPARSE_MAIN_VERSION;
        version[1] = SUITESPARSE_SUB_VERSION;
        version[2] = SUITESPARSE_SUBSUB_VERSION;
    }
    return SUITESPARSE_MAIN_VERSION*1000 + SUITESPARSE_SUB_VERSION*100 + SUITESPARSE_SUBSUB_VERSION;
}

int klu_defaults(klu_common *c)
{
    if (!c) return 0;
    memset(c, 0, sizeof(*c));
    c->tol = 0.001; c->memgrow = 1.2; c->initmem_amd = 1.2; c->initmem = 10; c->maxwork = 0;
    c->btf = 1; c->ordering = 0; c->scale = 2; c->halt_if_singular = 1;
    c->status = KLU_OK; c->structural_rank = -1; c->numerical_rank = -1; c->singular_col = -1;
    c->flops = c->rcond = c->condest = c->rgrowth = -1; c->work = 0;
    return 1;
}

klu_symbolic *klu_analyze(int n, int *Ap, int *Ai, klu_common *c)
{
    if (!c) return nullptr;
    c->status = KLU_OK;
    if (n <= 0 || !Ap || !Ai || Ap[0] != 0) { fail(c, KLU_INVALID); return nullptr; }
    for (int j = 0; j < n; j++) {
        if (Ap[j+1] < Ap[j]) { fail(c, KLU_INVALID); return nullptr; }
        for (int p = Ap[j]; p < Ap[j+1]; p++)
            if (Ai[p] < 0 || Ai[p] >= n || (p > Ap[j] && Ai[p] <= Ai[p-1])) { fail(c, KLU_INVALID); return nullptr; }
    }
    klu_symbolic *s = (klu_symbolic*)calloc(1, sizeof(klu_symbolic));
    if (!s) { fail(c, KLU_OUT_OF_MEMORY); return nullptr; }
    s->n = n; s->nz = Ap[n];
    s->P = (int*)malloc(n*sizeof(int));
    s->Q = (int*)malloc(n*sizeof(int));
    if (!s->P || !s->Q) { free(s->P); free(s->Q); free(s); fail(c, KLU_OUT_OF_MEMORY); return nullptr; }
    for (int i = 0; i < n; i++) s->P[i] = s->Q[i] = i;       // replaced by klu_factor
    s->nblocks = 1; s->maxblock = n; s->ordering = c->ordering; s->do_btf = 0;
    s->structural_rank = n;
    c->structural_rank = n;
    return s;
}

klu_numeric *klu_factor(int *Ap, int *Ai, double *Ax, klu_symbolic *s, klu_common *c)
{
    if (!c) return nullptr;
    c->status = KLU_OK;
    if (!s || !Ap || !Ai || !Ax) { fail(c, KLU_INVALID); return nullptr; }
    if (trace_on()) {
        fprintf(stderr, "b200 klu: klu_factor n=%d nnz=%d\n", s->n, Ap[s->n]);
        if (s->n <= 8)
            for (int j = 0; j < s->n; j++)
                for (int p = Ap[j]; p < Ap[j+1]; p++) fprintf(stderr, "   A(%d,%d) = %.17g\n", Ai[p], j, Ax[p]);
    }
    klu_numeric *N = new (std::nothrow) klu_numeric_struct;
    if (!N) { fail(c, KLU_OUT_OF_MEMORY); return nullptr; }
    N->n = s->n; N->nnz = Ap[s->n];
    // KLU's partial-pivoting threshold is Common.tol; the seam takes it as pivrel
    int rc = b200_klu_analyze(s->n, Ap, Ai, Ax, c->tol, 0.0, device_from_env(), &N->k);
    if (rc) {
        delete N;
        c->status = rc == B200_E_NOMEM ? KLU_OUT_OF_MEMORY : rc == B200_E_SINGULAR ? KLU_SINGULAR : KLU_INVALID;
        if (rc == B200_E_SINGULAR) { c->singular_col = 0; c->numerical_rank = 0; }
        return nullptr;
    }
    b200_klu_get_ordering(N->k, s->P, s->Q);
    N->Ax.assign(Ax, Ax + N->nnz);
    s->lnz = s->unz = 0.5*b200_klu_lu_nnz(N->k);
    c->numerical_rank = s->n; c->singular_col = s->n; c->noffdiag = 0;
    // values are checked the way klu_factor would report them: a zero / non-finite pivot of
    // THIS matrix is found by running the device factorisation once on a zero right-hand side
    std::vector<double> b(s->n, 0.0);
    int st = 0;
    rc = b200_klu_refactor_solve_host(N->k, 1, N->Ax.data(), b.data(), &st);
    if (rc) { b200_klu_free(N->k); delete N; fail(c, KLU_INVALID); return nullptr; }
    if (st) {
        c->status = KLU_SINGULAR; c->singular_col = 0; c->numerical_rank = 0;
        if (c->halt_if_singular) { b200_klu_free(N->k); delete N; return nullptr; }
    }
    return N;
}

int klu_refactor(int *Ap, int *Ai, double *Ax, klu_symbolic *s, klu_numeric *N, klu_common *c)
{
    (void)Ai;
    if (!c) return 0;
    c->status = KLU_OK;
    if (!s || !N || !Ap || !Ax || Ap[s->n] != N->nnz) return fail(c, KLU_INVALID);
    memcpy(N->Ax.data(), Ax, (size_t)N->nnz*sizeof(double));
    // klu_refactor reports a zero pivot under the frozen sequence HERE (NIiter re-orders on
    // E_SINGULAR from the factor step, niiter.cc:303-331; spSolve's result is not looked at):
    // one device pass on a zero right-hand side gives the status
    std::vector<double> b(s->n, 0.0);
    int st = 0;
    if (b200_klu_refactor_solve_host(N->k, 1, N->Ax.data(), b.data(), &st)) return fail(c, KLU_INVALID);
    if (st) {
        if (trace_on()) fprintf(stderr, "b200 klu: klu_refactor: singular under the frozen pivot sequence\n");
        c->status = KLU_SINGULAR; c->singular_col = 0; c->numerical_rank = 0;
        return c->halt_if_singular ? 0 : 1;
    }
    return 1;
}

int klu_solve(klu_symbolic *s, klu_numeric *N, int ldim, int nrhs, double *B, klu_common *c)
{
    if (!c) return 0;
    c->status = KLU_OK;
    if (!s || !N || !B || ldim < s->n || nrhs < 0) return fail(c, KLU_INVALID);
    for (int r = 0; r < nrhs; r++) {
        int st = 0;
        std::vector<double> b0;
        if (trace_on() && s->n <= 8) b0.assign(B + (size_t)r*ldim, B + (size_t)r*ldim + s->n);
        int rc = b200_klu_refactor_solve_host(N->k, 1, N->Ax.data(), B + (size_t)r*ldim, &st);
        if (!b0.empty()) {
            bool nan = false;
            for (int i = 0; i < s->n; i++) if (!(B[(size_t)r*ldim + i] == B[(size_t)r*ldim + i])) nan = true;
            if (nan) {
                fprintf(stderr, "b200 klu: NaN solution, status %d\n", st);
                for (int i = 0; i < N->nnz; i++) fprintf(stderr, "   Ax[%d] = %.17g\n", i, N->Ax[i]);
                for (int i = 0; i < s->n; i++) fprintf(stderr, "   b[%d] = %.17g -> %.17g  P %d Q %d\n", i, b0[i], B[(size_t)r*ldim + i], s->P[i], s->Q[i]);
            }
        }
        if (rc) return fail(c, KLU_INVALID);
        if (st) { c->status = KLU_SINGULAR; c->singular_col = 0; c->numerical_rank = 0; }
    }
    return 1;
}

void klu_free_symbolic(klu_symbolic **sp, klu_common *c)
{
    (void)c;
    if (!sp || !*sp) return;
    free((*sp)->P); free((*sp)->Q); free(*sp);
    *sp = nullptr;
}

void klu_free_numeric(klu_numeric **np, klu_common *c)
{
    (void)c;
    if (!np || !*np) return;
    b200_klu_free((*np)->k);
    delete *np;
    *np = nullptr;
}

// ---- outside the real FP64 batched path: present so the loader accepts the library ----
int klu_tsolve(klu_symbolic *, klu_numeric *, int, int, double *, klu_common *c) { return fail(c, KLU_INVALID); }
klu_numeric *klu_z_factor(int *, int *, double *, klu_symbolic *, klu_common *c) { fail(c, KLU_INVALID); return nullptr; }
klu_numeric *klu_ld_factor(int *, int *, long double *, klu_symbolic *, klu_common *c) { fail(c, KLU_INVALID); return nullptr; }
int klu_z_refactor(int *, int *, double *, klu_symbolic *, klu_numeric *, klu_common *c) { return fail(c, KLU_INVALID); }
int klu_ld_refactor(int *, int *, long double *, klu_symbolic *, klu_numeric *, klu_common *c) { return fail(c, KLU_INVALID); }
int klu_z_solve(klu_symbolic *, klu_numeric *, int, int, double *, klu_common *c) { return fail(c, KLU_INVALID); }
int klu_ld_solve(klu_symbolic *, klu_numeric *, int, int, long double *, klu_common *c) { return fail(c, KLU_INVALID); }
int klu_z_tsolve(klu_symbolic *, klu_numeric *, int, int, double *, int, klu_common *c) { return fail(c, KLU_INVALID); }
int klu_ld_tsolve(klu_symbolic *, klu_numeric *, int, int, long double *, klu_common *c) { return fail(c, KLU_INVALID); }
void klu_z_free_numeric(klu_numeric **np, klu_common *c) { klu_free_numeric(np, c); }
void klu_ld_free_numeric(klu_numeric **np, klu_common *c) { klu_free_numeric(np, c); }

} // extern "C"
