// klu_seam.cu -- (1) the KLU-hook seam as a batched C ABI, (2) the debug load hook.
//
// (1) spMatlabMatrix::factor/refactor/solve (xt_base/include/sparse/spmatrix.h:853-857) as
//     KLUmatrix implements them (wrspice/src/sparse/kluif.cc:569-628): klu_analyze +
//     klu_factor once, then klu_refactor + klu_solve per Newton iteration with the pivot
//     sequence frozen.  Batched: one CSC pattern, ntrials value arrays.  The same k_solve
//     kernel as the engine is used, with identity gather lists (A[p] = 0 + Ax[p]).
// (2) b200_batch_debug_load: one sCKT::load + loadGmin at a forced state, matrix/RHS
//     assembled to global memory for stamp-for-stamp comparison with the oracle.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "b200_internal.h"
#include "device_state.h"
#include "dev_load.cuh"
#include "dev_ctl.cuh"
#include "kernels.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "b200spice: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
    return B200_E_NOGPU; } } while (0)

struct b200_klu {
    Analysis an;
    Consts K;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;
    SolveView Sv;
    int cap_trials = 0;
    HotView hot; VecView V;
    double *dX0 = nullptr, *dX1 = nullptr, *dAxh = nullptr, *dBh = nullptr;
    int *d_packmap = nullptr; double *dpack = nullptr;
    std::vector<void*> trial_allocs;            // buffers sized by cap_trials (re-made when it grows)
    int d_const_n = 0; int *d_const_idx = nullptr; double *d_const_val = nullptr;

    template <class T> T *dalloc(size_t n) {
        void *p = nullptr;
        if (cudaMalloc(&p, (n ? n : 1)*sizeof(T)) != cudaSuccess) return nullptr;
        cudaMemset(p, 0, (n ? n : 1)*sizeof(T));
        allocs.push_back(p);
        return (T*)p;
    }
    template <class T> T *dupload(const std::vector<T> &v) {
        T *p = dalloc<T>(v.size());
        if (p && !v.empty()) cudaMemcpy(p, v.data(), v.size()*sizeof(T), cudaMemcpyHostToDevice);
        return p;
    }
    template <class T> T *talloc(size_t n) {        // uninitialised; klu_ensure clears on the solve's stream
        void *p = nullptr;
        if (cudaMalloc(&p, (n ? n : 1)*sizeof(T)) != cudaSuccess) return nullptr;
        trial_allocs.push_back(p);
        return (T*)p;
    }
};

// [rows][ntrials-major host] <-> [rows][ldt] device transposes
__global__ void k_transpose_in(int rows, int ntrials, int ldt, const double *src /* [nt][rows] */,
    double *dst /* [rows][ldt] */)
{
    __shared__ double tile[32][33];
    int r0 = blockIdx.y*32, t0 = blockIdx.x*32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int t = t0 + k, r = r0 + threadIdx.x;
        tile[k][threadIdx.x] = (t < ntrials && r < rows) ? src[(size_t)t*rows + r] : 0.0;
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int r = r0 + k, t = t0 + threadIdx.x;
        if (r < rows && t < ldt) dst[(size_t)r*ldt + t] = (t < ntrials) ? tile[threadIdx.x][k] : 0.0;
    }
}

__global__ void k_transpose_out(int rows, int ntrials, int ldt, const double *src /* [rows][ldt] */,
    double *dst /* [nt][rows] */)
{
    __shared__ double tile[32][33];
    int r0 = blockIdx.y*32, t0 = blockIdx.x*32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int r = r0 + k, t = t0 + threadIdx.x;
        tile[k][threadIdx.x] = (r < rows && t < ntrials) ? src[(size_t)r*ldt + t] : 0.0;
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int t = t0 + k, r = r0 + threadIdx.x;
        if (r < rows && t < ntrials) dst[(size_t)t*rows + r] = tile[threadIdx.x][k];
    }
}

__global__ void k_set_int(int *p, int n, int v)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// debug assemble: Ax[p][t], rhs[row][t] exactly as k_solve builds them (before the LU)
struct DebugTables {
    const double *Ainit; const int *cg_ptr, *cg_idx, *rg_ptr, *rg_idx, *diag_node;
};
__global__ void k_debug_assemble(Consts K, HotView hot, VecView V, SolveView Sv, DebugTables D, int nnz,
    double *Ax, double *rhs)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int e = GRID_ROW;
    if (t >= K.ntrials || e > nnz + K.n) return;
    int ldt = K.ldt;
    int mode = hot.mode[t];
    double diagGmin = hot.diaggmin[t];
    bool ns_on = (mode & MODEDC) && (mode & (MODEINITJCT | MODEINITFIX));
    bool ic_on = (mode & MODEDC) && (mode & MODETRANOP) && !(mode & MODEUIC);
    if (e < nnz) {
        int p = e;
        double v = 0.0;
        if (D.Ainit[p] != 0.0) v += D.Ainit[p];
        for (int q = D.cg_ptr[p]; q < D.cg_ptr[p+1]; q++)
            v += V.contrib[CADDR(K, D.cg_idx[q], t)];
        int dn = D.diag_node[p];
        if (dn) {
            if (ns_on && Sv.ns_given[dn]) v += 1.0;
            if (ic_on && Sv.ic_given[dn]) v += 1.0;
            if (Sv.ntype[dn] == 0) {
                if (diagGmin != 0.0) v += diagGmin;
                else if (K.opt.forcegmin && (v >= 0.0 && v < K.opt.gmin)) v = K.opt.gmin;
                else if (v > K.opt.gmax) v = K.opt.gmax;
            }
        }
        Ax[(size_t)p*ldt + t] = v;
    }
    else {
        int row = e - nnz;                   // 0..n
        double v = 0.0;
        if (row > 0) {
            for (int q = D.rg_ptr[row]; q < D.rg_ptr[row+1]; q++)
                v += V.contrib[CADDR(K, D.rg_idx[q], t)];
            if (ns_on && Sv.ns_given[row]) v += Sv.ns_val[row];
            if (ic_on && Sv.ic_given[row]) v += Sv.ic_val[row];
        }
        rhs[(size_t)row*ldt + t] = v;
    }
}

// KLU seam: [entry][ldt] value arrays -> rows of contrib[NC][ldt] in assembly order
__global__ void k_pack_stream(Consts K, int nent, const int *map, const double *src, double *dst)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int e = GRID_ROW;
    if (t >= K.ldt || e >= nent) return;
    dst[CADDR(K, map[e], t)] = src[(size_t)e*K.ldt + t];
}

__global__ void k_debug_set_hot(Consts K, HotView hot, int mode, double time, double delta,
    int order, double ag0, double ag1, double p0, double p1)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= K.ldt) return;
    hot.mode[t] = mode; hot.order[t] = order; hot.st_base[t] = 0; hot.noncon[t] = 0;
    hot.trapbad[t] = 0; hot.trapcheck[t] = 0; hot.time[t] = time; hot.delta[t] = delta;
    hot.srcfact[t] = 1.0; hot.diaggmin[t] = 0.0;
    hot.ag[t] = ag0; hot.ag[K.ldt + t] = ag1;
    hot.pred[t] = p0; hot.pred[K.ldt + t] = p1;
}

// accessors implemented in capi.cu
int b200_internal_prepare(b200_batch *b);
int b200_internal_views(b200_batch *b, Consts **K, HotView **hot, VecView **V, MosView **M,
    CapView **Cp, SrcView **S, SolveView **Sv, cudaStream_t *st, int **cur);
const Analysis *b200_internal_analysis(b200_batch *b);

// dev_div_rcp (dev_load.cuh) against IEEE division on pseudo-random operands
__global__ void k_divcheck(long long n, unsigned long long seed, int emin, int emax, unsigned long long *bad)
{
    long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    // splitmix64
    auto next = [](unsigned long long &x) {
        x += 0x9e3779b97f4a7c15ull;
        unsigned long long z = x;
        z = (z ^ (z >> 30))*0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27))*0x94d049bb133111ebull;
        return z ^ (z >> 31);
    };
    unsigned long long st = seed + 0x632be59bd9b4e019ull*(unsigned long long)i;
    auto rnd = [&](void) {
        unsigned long long m = next(st), e = next(st), sg = next(st);
        double f = 1.0 + (double)(m >> 12)*(1.0/4503599627370496.0);      // [1,2), all 52 bits
        int ex = emin + (int)(e % (unsigned long long)(emax - emin + 1));
        double v = ldexp(f, ex);
        return (sg & 1) ? -v : v;
    };
    double a = rnd(), b = fabs(rnd());
    double y = 1.0/b;
    double q = dev_div_rcp(a, b, y);
    if (__double_as_longlong(q) != __double_as_longlong(a/b)) atomicAdd(bad, 1ull);
}

extern "C" {

int b200_debug_divcheck(long long n, unsigned long long seed, int emin, int emax, long long *mismatches)
{
    unsigned long long *d = nullptr, h = 0;
    CK(cudaMalloc(&d, sizeof(h)));
    CK(cudaMemset(d, 0, sizeof(h)));
    k_divcheck<<<(unsigned)((n + 255)/256), 256>>>(n, seed, emin, emax, d);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    *mismatches = (long long)h;
    return B200_OK;
}

int b200_batch_debug_load(b200_batch *b, int mode, double time, double delta, int order,
    const double *rhsold, double *Ax, double *rhs)
{
    if (!b) return B200_E_BADPARM;
    int rc = b200_internal_prepare(b);
    if (rc) return rc;
    Consts *K; HotView *hot; VecView *V; MosView *M; CapView *Cp; SrcView *S; SolveView *Sv;
    cudaStream_t st; int *cur;
    b200_internal_views(b, &K, &hot, &V, &M, &Cp, &S, &Sv, &st, &cur);
    if (K->method != 0) return B200_E_UNSUPP;
    size_t ldt = K->ldt;
    int n = K->n, nnz = K->nnz, nt = K->ntrials;
    double ag0 = 0, ag1 = 0, p0 = 1.0, p1 = 0.0;
    if (mode & MODETRAN) {
        // NIcomCof + predict with deltaOld[*] = delta (oracle wo_debug_load does the same)
        if (order == 2) { double t0 = 1.0/(1.0 - K->opt.xmu); ag0 = t0/delta; ag1 = t0*K->opt.xmu; }
        else { ag0 = 1.0/delta; ag1 = -ag0; }
        p1 = -delta/delta; p0 = 1 - p1;
    }
    k_debug_set_hot<<<(unsigned)((ldt + 127)/128), 128, 0, st>>>(*K, *hot, mode, time, delta, order, ag0, ag1, p0, p1);
    double *X = V->X[0];
    if (rhsold) {
        std::vector<double> row(ldt);
        for (int i = 0; i <= n; i++) {
            std::fill(row.begin(), row.end(), 0.0);
            memcpy(row.data(), rhsold + (size_t)i*nt, nt*sizeof(double));
            CK(cudaMemcpyAsync(X + (size_t)i*ldt, row.data(), ldt*sizeof(double), cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
        }
    }
    else
        CK(cudaMemsetAsync(X, 0, (size_t)(n + 1)*ldt*sizeof(double), st));
    launch_loads(st, *K, *hot, *V, *M, *Cp, *S, X);
    double *dAx = nullptr, *drhs = nullptr;
    CK(cudaMalloc(&dAx, (size_t)nnz*ldt*sizeof(double)));
    CK(cudaMalloc(&drhs, (size_t)(n + 1)*ldt*sizeof(double)));
    dim3 g = grid_rows((nt + 127)/128, nnz + n + 1);
    const Analysis &an = *b200_internal_analysis(b);
    std::vector<int> dn(an.nnz, 0);
    for (int i = 1; i <= an.n; i++) if (an.diag[i] >= 0) dn[an.diag[i]] = i;
    std::vector<void*> tmp;
    auto up = [&](const void *h, size_t bytes) -> void* {
        void *p = nullptr;
        if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) return nullptr;
        if (bytes) cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice);
        tmp.push_back(p);
        return p;
    };
    DebugTables D;
    D.Ainit = (const double*)up(an.Ainit.data(), an.Ainit.size()*sizeof(double));
    D.cg_ptr = (const int*)up(an.cg_ptr.data(), an.cg_ptr.size()*sizeof(int));
    D.cg_idx = (const int*)up(an.cg_idx.data(), an.cg_idx.size()*sizeof(int));
    D.rg_ptr = (const int*)up(an.rg_ptr.data(), an.rg_ptr.size()*sizeof(int));
    D.rg_idx = (const int*)up(an.rg_idx.data(), an.rg_idx.size()*sizeof(int));
    D.diag_node = (const int*)up(dn.data(), dn.size()*sizeof(int));
    k_debug_assemble<<<g, 128, 0, st>>>(*K, *hot, *V, *Sv, D, nnz, dAx, drhs);
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (void *p : tmp) cudaFree(p);
    std::vector<double> hb((size_t)std::max(nnz, n + 1)*ldt);
    CK(cudaMemcpy(hb.data(), dAx, (size_t)nnz*ldt*sizeof(double), cudaMemcpyDeviceToHost));
    for (int p = 0; p < nnz; p++) memcpy(Ax + (size_t)p*nt, hb.data() + (size_t)p*ldt, nt*sizeof(double));
    CK(cudaMemcpy(hb.data(), drhs, (size_t)(n + 1)*ldt*sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i <= n; i++) memcpy(rhs + (size_t)i*nt, hb.data() + (size_t)i*ldt, nt*sizeof(double));
    cudaFree(dAx); cudaFree(drhs);
    return B200_OK;
}

// ---------------------------------------------------------------------------------------
int b200_klu_analyze(int n, const int *Ap, const int *Ai, const double *Ax0, double pivrel,
    double pivabs, int device, b200_klu **out)
{
    *out = nullptr;
    if (!(pivrel > 0.0) || pivrel > 1.0) pivrel = 1e-3;     // circuit.h:461-631 defaults, clamped like the task options
    if (!(pivabs >= 0.0)) pivabs = 1e-13;
    if (n <= 0 || !Ap || !Ai || !Ax0) return B200_E_BADPARM;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev) {
        fprintf(stderr, "b200spice: no CUDA device (no CPU fallback exists)\n");
        return B200_E_NOGPU;
    }
    CK(cudaSetDevice(device));
    b200_klu *k = new b200_klu;
    k->device = device;
    Analysis &an = k->an;
    an.n = n; an.nnz = Ap[n];
    an.Ap.assign(Ap, Ap + n + 1); an.Ai.assign(Ai, Ai + an.nnz);
    an.ntype.assign(n + 1, 0);
    an.diag.assign(n + 1, -1);
    for (int j = 0; j < n; j++)
        for (int p = Ap[j]; p < Ap[j+1]; p++) if (Ai[p] == j) an.diag[j+1] = p;
    // ordering sees the first matrix: +-1 entries count as constants, nonzero diagonals as safe
    an.Ainit.assign(Ax0, Ax0 + an.nnz);
    an.cg_ptr.assign(an.nnz + 1, 0);
    an.rg_ptr.assign(n + 2, 0);
    // entries that are exactly zero in this matrix are not live for the choice of pivots
    an.dc_active.assign(an.nnz, 0);
    for (int p = 0; p < an.nnz; p++) an.dc_active[p] = (Ax0[p] != 0.0);
    b200_order_markowitz(an);
    // klu_factor semantics on the matrix handed in: the static order is kept only if every one
    // of its pivots passes the pivrel / pivabs test on THESE values; otherwise threshold
    // Markowitz on the values decides (what spOrderAndFactor / klu_factor would do)
    if (b200_order_numeric(an, Ax0, pivrel, pivabs, 0) >= 0 &&
            b200_order_numeric(an, Ax0, pivrel, pivabs, 1) >= 0) {
        delete k;
        return B200_E_SINGULAR;
    }
    // run-time gather lists are the identity: A[p] = 0 + Ax[p], b[row] = 0 + B[row-1]
    an.Ainit.assign(an.nnz, 0.0);
    an.cg_idx.resize(an.nnz);
    for (int p = 0; p < an.nnz; p++) { an.cg_ptr[p] = p; an.cg_idx[p] = p; }
    an.cg_ptr[an.nnz] = an.nnz;
    an.rg_idx.resize(n);
    an.rg_ptr[0] = 0; an.rg_ptr[1] = 0;
    for (int r = 1; r <= n; r++) { an.rg_idx[r-1] = r - 1; an.rg_ptr[r+1] = r; }
    an.ic_given.assign(n + 1, 0); an.ns_given.assign(n + 1, 0);
    an.ic_val.assign(n + 1, 0.0); an.ns_val.assign(n + 1, 0.0);
    an.ncslot = an.nnz; an.nrslot = n;
    int rc = b200_build_lu(an);                // renumbers cg_idx / rg_idx into assembly order
    if (rc) { delete k; return rc; }

    Consts &K = k->K;
    memset(&K, 0, sizeof(K));
    b200_options_default(&K.opt);
    K.opt.gmax = HUGE_VAL;
    K.n = n; K.nnz = an.nnz; K.nnzLU = an.nnzLU; K.NS = 4; K.NC = an.NC;
    CK(cudaStreamCreate(&k->stream));
    SolveView &Sv = k->Sv;
    memset(&Sv, 0, sizeof(Sv));
    Sv.cls = -1;
    Sv.ntype = k->dupload(an.ntype);
    Sv.ic_given = k->dupload(an.ic_given); Sv.ns_given = k->dupload(an.ns_given);
    Sv.ic_val = k->dupload(an.ic_val); Sv.ns_val = k->dupload(an.ns_val);
    Sv.yperm_out = k->dupload(an.yperm_out);
    Sv.runs = k->dupload(an.runs);
    Sv.tile_order = k->dupload(an.tile_order); Sv.ntiles = (int)an.tile_order.size();
    Sv.img = an.img; Sv.nlu = an.nnzLU; Sv.npad = (n + 31) & ~31;
    Sv.rows = k->dupload(an.prog_rows); Sv.rounds = k->dupload(an.prog_round);
    Sv.wide = an.wide;
    Sv.long_chunks = k->dupload(an.long_chunks); Sv.nlong = (int)an.long_chunks.size(); Sv.NC = an.NC;
    Sv.round_row = k->dupload(an.round_row); Sv.level_ptr = k->dupload(an.level_ptr); Sv.nlevels = (int)an.level_ptr.size() - 1;
    Sv.nrows = (int)(an.prog_rows.size()/(LU_LANES*(an.wide ? 2 : 1))); Sv.nrounds = (int)an.prog_round.size();
    Sv.ctr = k->dalloc<int>(2);
    rc = solve_configure(Sv);
    if (rc) { b200_klu_free(k); return rc; }
    {
        std::vector<int> map(an.cg_idx);
        map.insert(map.end(), an.rg_idx.begin(), an.rg_idx.end());
        k->d_packmap = k->dupload(map);
    }
    k->d_const_n = (int)an.const_idx.size();
    k->d_const_idx = k->dupload(an.const_idx); k->d_const_val = k->dupload(an.const_val);
    CK(cudaDeviceSynchronize());
    *out = k;
    return B200_OK;
}

void b200_klu_free(b200_klu *k)
{
    if (!k) return;
    cudaSetDevice(k->device);
    for (void *p : k->allocs) cudaFree(p);
    for (void *p : k->trial_allocs) cudaFree(p);
    if (k->stream) cudaStreamDestroy(k->stream);
    delete k;
}

int b200_klu_lu_nnz(const b200_klu *k) { return k ? k->an.nnzLU : 0; }

int b200_klu_get_ordering(const b200_klu *k, int *rowperm, int *colperm)
{
    if (!k) return B200_E_BADPARM;
    if (rowperm) std::copy(k->an.rowperm.begin(), k->an.rowperm.end(), rowperm);
    if (colperm) std::copy(k->an.colperm.begin(), k->an.colperm.end(), colperm);
    return B200_OK;
}

// (re)allocates everything sized by the trial count and initialises it ON THE SOLVE'S STREAM
static int klu_ensure(b200_klu *k, int ldt, cudaStream_t st)
{
    if (k->cap_trials == ldt) return B200_OK;
    if (!k->trial_allocs.empty()) {
        CK(cudaDeviceSynchronize());            // nothing in flight may still use the old buffers
        for (void *p : k->trial_allocs) cudaFree(p);
        k->trial_allocs.clear();
    }
    k->cap_trials = 0;
    HotView &h = k->hot;
    memset(&h, 0, sizeof(h));
    h.mode = k->talloc<int>(ldt); h.do_solve = k->talloc<int>(ldt); h.iterno = k->talloc<int>(ldt);
    h.noncon = k->talloc<int>(ldt); h.diaggmin = k->talloc<double>(ldt);
    h.solve_err = k->talloc<int>(ldt); h.nodeconv = k->talloc<int>(ldt);
    h.devconv = k->talloc<int>(ldt); h.need_devconv = k->talloc<int>(ldt);
    k->dX0 = k->talloc<double>((size_t)(k->an.n + 1)*ldt);
    k->dX1 = k->talloc<double>((size_t)(k->an.n + 1)*ldt);
    k->dpack = k->talloc<double>(((size_t)k->an.NC + k->an.long_chunks.size())*ldt);
    SolveView &Sv = k->Sv;
    Sv.image = k->talloc<double>((size_t)ldt*Sv.img);
    Sv.ysol = k->talloc<double>((size_t)ldt*Sv.npad);
    Sv.lu_bad = k->talloc<int>(ldt); Sv.fin_flags = k->talloc<int>(ldt); Sv.live = k->talloc<int>(ldt);
    if (!k->dX1 || !k->dpack || !Sv.image || !Sv.ysol || !Sv.live) return B200_E_NOMEM;
    int *iz[] = { h.mode, h.iterno, h.noncon, h.solve_err, h.nodeconv, h.devconv, h.need_devconv, Sv.lu_bad, Sv.fin_flags };
    for (int *p : iz) CK(cudaMemsetAsync(p, 0, ldt*sizeof(int), st));
    CK(cudaMemsetAsync(h.diaggmin, 0, ldt*sizeof(double), st));
    CK(cudaMemsetAsync(k->dX0, 0, (size_t)(k->an.n + 1)*ldt*sizeof(double), st));
    CK(cudaMemsetAsync(k->dpack, 0, (size_t)k->an.NC*ldt*sizeof(double), st));
    k_set_int<<<(ldt + 127)/128, 128, 0, st>>>(h.do_solve, ldt, 1);
    {
        // the constant terms of the assembly (the -1.0 of the split sums)
        Consts K = k->K; K.ldt = ldt; K.ntrials = ldt;
        launch_fill_const(st, K, k->dpack, k->d_const_n, k->d_const_idx, k->d_const_val);
    }
    k->cap_trials = ldt;
    return B200_OK;
}

int b200_klu_refactor_solve_dev(b200_klu *k, int ntrials, int ldt, const double *dAx, double *dB,
    int *dstatus, void *cuda_stream)
{
    if (!k || ntrials <= 0 || ldt < ntrials || (ldt % 32)) return B200_E_BADPARM;
    CK(cudaSetDevice(k->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : k->stream;
    // the scratch buffers (and the constant rows written into them) are pitched for one ldt
    if (ldt != k->cap_trials) {
        int rc = klu_ensure(k, ldt, st);
        if (rc) return rc;
    }
    Consts K = k->K;
    K.ntrials = ntrials; K.ldt = ldt;
    VecView V;
    memset(&V, 0, sizeof(V));
    V.contrib = k->dpack;
    int nnz = k->an.nnz, n = k->an.n;
    k_pack_stream<<<grid_rows((ldt + 127)/128, nnz), 128, 0, st>>>(K, nnz, k->d_packmap, dAx, k->dpack);
    k_pack_stream<<<grid_rows((ldt + 127)/128, n), 128, 0, st>>>(K, n, k->d_packmap + nnz, dB, k->dpack);
    // solution goes to dX1 ([n+1][ldt], row 0 = ground), then back into dB
    int rc = launch_solve(st, K, k->hot, V, k->Sv, k->dX0, k->dX1);
    if (rc) return rc;
    CK(cudaMemcpyAsync(dB, k->dX1 + ldt, (size_t)k->an.n*ldt*sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (dstatus)
        CK(cudaMemcpyAsync(dstatus, k->hot.solve_err, ntrials*sizeof(int), cudaMemcpyDeviceToDevice, st));
    return B200_OK;
}

int b200_klu_refactor_solve_host(b200_klu *k, int ntrials, const double *Ax, double *B, int *status)
{
    if (!k || ntrials <= 0 || !Ax || !B) return B200_E_BADPARM;
    CK(cudaSetDevice(k->device));
    int ldt = (ntrials + 31) & ~31;
    int n = k->an.n, nnz = k->an.nnz;
    cudaStream_t st = k->stream;
    struct Tmp {                                   // freed on every exit path
        std::vector<void*> p;
        ~Tmp() { for (void *q : p) cudaFree(q); }
        void *get(size_t bytes) { void *q = nullptr; if (cudaMalloc(&q, bytes) != cudaSuccess) return nullptr; p.push_back(q); return q; }
    } tmp;
    double *hAx = (double*)tmp.get((size_t)ntrials*nnz*sizeof(double));
    double *hB = (double*)tmp.get((size_t)ntrials*n*sizeof(double));
    double *dAx = (double*)tmp.get((size_t)nnz*ldt*sizeof(double));
    double *dB = (double*)tmp.get((size_t)n*ldt*sizeof(double));
    int *dst = (int*)tmp.get(ldt*sizeof(int));
    if (!hAx || !hB || !dAx || !dB || !dst) return B200_E_NOMEM;
    CK(cudaMemcpyAsync(hAx, Ax, (size_t)ntrials*nnz*sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(hB, B, (size_t)ntrials*n*sizeof(double), cudaMemcpyHostToDevice, st));
    dim3 blk(32, 8);
    k_transpose_in<<<dim3(ldt/32, (nnz + 31)/32), blk, 0, st>>>(nnz, ntrials, ldt, hAx, dAx);
    k_transpose_in<<<dim3(ldt/32, (n + 31)/32), blk, 0, st>>>(n, ntrials, ldt, hB, dB);
    int rc = b200_klu_refactor_solve_dev(k, ntrials, ldt, dAx, dB, dst, st);
    if (rc) return rc;
    // dB now holds x as [n][ldt]
    k_transpose_out<<<dim3(ldt/32, (n + 31)/32), blk, 0, st>>>(n, ntrials, ldt, dB, hB);
    CK(cudaMemcpyAsync(B, hB, (size_t)ntrials*n*sizeof(double), cudaMemcpyDeviceToHost, st));
    if (status) CK(cudaMemcpyAsync(status, dst, ntrials*sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return B200_OK;
}

} // extern "C"
