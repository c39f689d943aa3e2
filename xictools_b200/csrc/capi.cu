// capi.cu -- C ABI (include/b200spice.h): batch construction, the tick loop, result export.
// No CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "b200_internal.h"
#include "device_state.h"
#include "dev_load.cuh"
#include "dev_ctl.cuh"
#include "kernels.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "b200spice: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
    return B200_E_NOGPU; } } while (0)

namespace {
inline bool given(double x) { return !std::isnan(x); }

// dependency of derived MOS parameters on the raw ones that may vary per trial
// (mostemp.cc:141-300): derived parameter k is per-trial iff it depends on a varied raw one.
const int kDep[MR_NRAW][8] = {
    /* vto    */ { MP_tVbi, MP_tVto, -1 },
    /* kp     */ { MP_beta, -1 },
    /* gamma  */ { MP_gamma, MP_tVbi, MP_tVto, -1 },
    /* phi    */ { MP_tPhi, MP_tVbi, MP_tVto, MP_r2phi, MP_r15phi, -1 },
    /* lambda */ { MP_lambda, -1 },
    /* l      */ { MP_beta, MP_oxideCap, MP_gbOvl, MP_xl, -1 },
    /* w      */ { MP_beta, MP_oxideCap, MP_gsOvl, MP_gdOvl, MP_w, -1 },
};
}

struct b200_batch {
    Analysis an;
    Consts K;
    int device = 0;
    int ntrials = 0, ldt = 0;
    bool uic = false;
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;

    // host-side parameter tables
    std::vector<MosCard> cards;
    std::vector<double> mos_raw[MR_NRAW];                   // [nmos] shared values
    std::map<int, std::vector<double>> mos_var[MR_NRAW];    // inst -> [ntrials]
    std::vector<double> cap_c, cap_ic, src_dc, ind_l, ind_ic;
    std::vector<double> cap_craw, cap_m;                    // card value and multiplier (per-trial rows carry the card value)
    bool cap_has_m = false;
    std::vector<int> src_dcgiven; int *d_src_dcgiven = nullptr;
    std::map<int, std::vector<double>> cap_var, src_var, ind_var, jj_var, res_var;
    std::vector<int> save_slot;                             // unknown -> row of the output buffer (empty: all saved)
    std::vector<double> jj_ics, tjm_ics;                    // [njj] / [ntjm] card ics (NaN = not given)
    std::map<int, std::vector<double>> tjm_var;
    double *d_tjmics = nullptr; int tjmics_ts = 0;
    double *d_tjmprm[TP_NPARAMS] = {0}; int tjmprm_ts = -1;
    const double *d_tjm_area0 = nullptr, *d_tjm_vshunt = nullptr, *d_tjm_tcf = nullptr, *d_tjm_alphaN = nullptr;
    double *d_jjics = nullptr; int jjics_ts = 0;
    double *d_jjprm[JP_NPARAMS] = {0}; int jjprm_ts = -1;
    const double *d_jj_area0 = nullptr, *d_jj_vshunt = nullptr, *d_jj_tcf = nullptr;
    bool params_dirty = true;
    // packed per-trial rows (b200_batch_set_params): table id (0..MR_NRAW-1 MOS raw, then
    // PT_CAP / PT_IND / PT_SRC), instance index, values on the device as [nrows][ntrials]
    std::vector<int> pk_tbl, pk_idx;
    double *d_pack = nullptr; size_t pack_cap = 0;
    int *d_pk_desc = nullptr; size_t pk_desc_cap = 0;
    // persistent output staging of get_waves
    mutable double *d_wout = nullptr; mutable size_t wout_cap = 0;
    mutable int *d_wsel = nullptr; mutable int wsel_cap = 0;

    // device views
    HotView hot; VecView V; MosView M; CapView Cp; SrcView S; SolveView Sv, Sv2; CtlAux aux;
    bool two_programs = false;
    CtlShared sh;
    TrialCtl *ctl = nullptr;
    MosCard *d_cards = nullptr;
    int nq = 0; int *d_qslots = nullptr;
    int nconst = 0; int *d_const_idx = nullptr; double *d_const_val = nullptr;   // preloaded matrix constants
    int *d_count = nullptr;
    int cur = 0;
    double *d_raw[MR_NRAW] = {0}; int raw_ts[MR_NRAW] = {0};
    double *d_prm[MP_NPARAMS] = {0}; int prm_ts[MP_NPARAMS] = {0};
    double *d_capc = nullptr, *d_srcdc = nullptr; int capc_ts = 0, srcdc_ts = 0;
    double *d_capic = nullptr;
    double *d_indl = nullptr, *d_indic = nullptr; int indl_ts = 0;
    double *d_resr = nullptr, *d_resm = nullptr, *d_rt_base = nullptr; int resr_ts = 0;
    int *d_rt_cslot = nullptr, *d_rt_ptr = nullptr, *d_rt_res = nullptr; signed char *d_rt_sign = nullptr;
    int *d_counters = nullptr, *d_npts = nullptr;

    bool ran = false;
    b200_run_stats stats;

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
};

enum { PT_CAP = MR_NRAW, PT_IND = MR_NRAW + 1, PT_SRC = MR_NRAW + 2, PT_JJ = MR_NRAW + 3,
       PT_TJM = MR_NRAW + 4, PT_RES = MR_NRAW + 5, PT_NTABLES = MR_NRAW + 6 };

// row i of a [n][ldt] per-trial table <- the shared value of instance i
__global__ void k_fill_rows(double *dst, const double *shared, int n, int ldt)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;
    if (t < ldt && i < n) dst[(size_t)i*ldt + t] = shared[i];
}

// packed rows [nrows][ntrials] -> row desc[2r+1] of the table whose base is bases[desc[2r]];
// the padding trials repeat trial 0 (they run the nominal arithmetic and are never read back)
struct TableBases { double *p[PT_NTABLES]; };
// per-trial capacitance rows hold the card value: CAPcapac = nomCapac*tcFactor*m (captemp.cc:98)
__global__ void k_scale_rows(double *dst, const double *scale, int n, int ldt)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int i = GRID_ROW;
    if (t < ldt && i < n) dst[(size_t)i*ldt + t] = dst[(size_t)i*ldt + t]*scale[i];
}

__global__ void k_scatter_rows(TableBases tb, const int *desc, const double *vals, int nrows, int ntrials, int ldt)
{
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int r = GRID_ROW;
    if (t >= ldt || r >= nrows) return;
    double *base = tb.p[desc[2*r]];
    if (!base || desc[2*r + 1] < 0) return;
    const double *src = vals + (size_t)r*ntrials;
    base[(size_t)desc[2*r + 1]*ldt + t] = src[t < ntrials ? t : 0];
}

static int param_table(const b200_batch *b, int param, int index)
{
    int tbl = -1, n = 0;
    switch (param) {
    case B200_P_MOS_VTO: tbl = MR_vto; break;
    case B200_P_MOS_KP: tbl = MR_kp; break;
    case B200_P_MOS_GAMMA: tbl = MR_gamma; break;
    case B200_P_MOS_PHI: tbl = MR_phi; break;
    case B200_P_MOS_LAMBDA: tbl = MR_lambda; break;
    case B200_P_MOS_L: tbl = MR_l; break;
    case B200_P_MOS_W: tbl = MR_w; break;
    case B200_P_CAP_C: tbl = PT_CAP; break;
    case B200_P_IND_L: tbl = PT_IND; break;
    case B200_P_SRC_DC: tbl = PT_SRC; break;
    case B200_P_JJ_ICS: tbl = PT_JJ; break;
    case B200_P_TJM_ICS: tbl = PT_TJM; break;
    case B200_P_RES_R: tbl = PT_RES; break;
    default: return -B200_E_UNSUPP;
    }
    n = tbl < MR_NRAW ? b->an.mos.n : tbl == PT_CAP ? b->an.cap.n : tbl == PT_IND ? b->an.ind.n :
        tbl == PT_JJ ? b->an.jj.n : tbl == PT_TJM ? b->an.tjm.n : tbl == PT_RES ? b->an.nres : b->an.src.n;
    if (index < 0 || index >= n) return -B200_E_BADPARM;
    return tbl;
}

extern "C" {

const char *b200_version(void) { return "b200spice 0.1 (sm_100a)"; }

int b200_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void b200_options_default(b200_options *o)
{
    // wrspice/include/circuit.h:461-631
    memset(o, 0, sizeof(*o));
    o->abstol = 1e-12; o->reltol = 1e-3; o->vntol = 1e-6; o->chgtol = 1e-14; o->trtol = 7.0;
    o->gmin = 1e-12; o->gmax = 1e6; o->pivrel = 1e-3; o->pivtol = 1e-13;
    o->temp = 298.15; o->tnom = 298.15;
    o->xmu = 0.5; o->dcmu = 0.5; o->trapratio = 10.0; o->resmin = 1e-3;
    o->bypass = 1; o->maxord = 2; o->itl1 = 400; o->itl2 = 100; o->itl4 = 20;
    o->itl_gmin = 20; o->itl_src = 20; o->interplev = 1;
}

static void build_cards(b200_batch *b, const b200_deck *dk)
{
    const Analysis &an = b->an;
    const b200_options &o = dk->opt;
    int nm = dk->nmos;
    b->cards.resize(nm);
    for (int k = 0; k < MR_NRAW; k++) b->mos_raw[k].assign(nm, 0.0);
    for (int i = 0; i < nm; i++) {
        const b200_mos &in = dk->mos[i];
        const b200_mos_model &md = dk->mos_models[in.model];
        MosCard &c = b->cards[i];
        memset(&c, 0, sizeof(c));
        c.type = md.type > 0 ? 1 : -1;
        c.level = (md.level == 2 || md.level == 3 || md.level == 6) ? md.level : 1;
        c.g_vto = given(md.vto); c.g_kp = given(c.level == 6 ? md.kc : md.kp); c.g_gamma = given(md.gamma);
        c.g_phi = given(md.phi); c.g_tox = given(md.tox); c.g_u0 = given(md.u0);
        c.g_nsub = given(md.nsub); c.g_cbd = given(md.cbd); c.g_cbs = given(md.cbs);
        c.g_cj = given(md.cj); c.g_cjsw = given(md.cjsw); c.g_nss = given(md.nss);
        c.g_tpg = given(md.tpg);
        // mosset.cc:83-200 defaults
        c.cbd = given(md.cbd) ? md.cbd : 0; c.cbs = given(md.cbs) ? md.cbs : 0;
        c.is = given(md.is) ? md.is : 1e-14; c.pb = given(md.pb) ? md.pb : .8;
        c.cgso = given(md.cgso) ? md.cgso : 0; c.cgdo = given(md.cgdo) ? md.cgdo : 0;
        c.cgbo = given(md.cgbo) ? md.cgbo : 0;
        c.cj = given(md.cj) ? md.cj : 0; c.mj = given(md.mj) ? md.mj : .5;
        c.cjsw = given(md.cjsw) ? md.cjsw : 0;
        c.mjsw = given(md.mjsw) ? md.mjsw : ((c.level == 1 || c.level == 6) ? .5 : .33);
        c.js = given(md.js) ? md.js : 0;
        c.tox = given(md.tox) ? md.tox : ((c.level == 2 || c.level == 3) ? 1e-7 : 0);
        c.ld = given(md.ld) ? md.ld : 0; c.u0 = given(md.u0) ? md.u0 : 0;
        c.fc = given(md.fc) ? md.fc : .5; c.nsub = given(md.nsub) ? md.nsub : 0;
        c.tpg = given(md.tpg) ? md.tpg : 1; c.nss = given(md.nss) ? md.nss : 0;
        c.tnom = given(md.tnom) ? md.tnom : o.tnom;
        // levels 2/3, mosset.cc:208-253
        c.uexp = given(md.uexp) ? md.uexp : 0; c.neff = given(md.neff) ? md.neff : 1;
        c.ucrit = given(md.ucrit) ? md.ucrit : 1e4;
        // the card's delta is the narrow-channel factor itself (mossetm.cc:216-219); level 3 then
        // overwrites it with 0 (mostemp.cc:138-142, MOSdelta is never set)
        c.narrow = (c.level == 2 && given(md.delta)) ? md.delta : 0;
        c.nfs = given(md.nfs) ? md.nfs : 0; c.vmax = given(md.vmax) ? md.vmax : 0;
        c.xj = given(md.xj) ? md.xj : 0; c.eta = given(md.eta) ? md.eta : 0;
        c.theta = given(md.theta) ? md.theta : 0; c.kappa = given(md.kappa) ? md.kappa : .2;
        // level 6, mosset.cc:255-272
        c.kv = given(md.kv) ? md.kv : 2; c.nv = given(md.nv) ? md.nv : 0.5;
        c.nc = given(md.nc) ? md.nc : 1;
        c.lamda0 = given(md.lamda0) ? md.lamda0 : (given(md.lambda) ? md.lambda : 0);
        c.lamda1 = given(md.lamda1) ? md.lamda1 : 0; c.sigma = given(md.sigma) ? md.sigma : 0;
        c.gamma1 = given(md.gamma1) ? md.gamma1 : 0;
        // instance, mosset.cc:283-373
        c.mult = given(in.m) ? in.m : 1.0;
        c.temp = given(in.temp) ? in.temp : o.temp;
        c.drainPerimeter = given(in.pd) ? in.pd : 0;
        c.sourcePerimeter = given(in.ps) ? in.ps : 0;
        double nrd = (!given(in.nrd) || in.nrd == 0) ? 1 : in.nrd;
        double nrs = (!given(in.nrs) || in.nrs == 0) ? 1 : in.nrs;
        c.drainArea = given(in.ad) ? in.ad : (o.defad > 0 ? o.defad : 0.0);
        c.sourceArea = given(in.as) ? in.as : (o.defas > 0 ? o.defas : 0.0);
        if (given(md.rd)) c.dcon = (md.rd != 0) ? 1/md.rd : 0;
        else if (given(md.rsh)) c.dcon = (md.rsh != 0) ? 1/(md.rsh*nrd) : 0;
        else c.dcon = 0;
        if (given(md.rs)) c.scon = (md.rs != 0) ? 1/md.rs : 0;
        else if (given(md.rsh)) c.scon = (md.rsh != 0) ? 1/(md.rsh*nrs) : 0;
        else c.scon = 0;
        c.icVDS = given(in.icvds) ? in.icvds : 0;
        c.icVGS = given(in.icvgs) ? in.icvgs : 0;
        c.icVBS = given(in.icvbs) ? in.icvbs : 0;
        if (dk->uic) {
            // setic() -> MOSdev::getic, mosic.cc:52-81 (node ICs are shared by all trials)
            auto icv = [&](int node) { return node ? (an.ns_given[node] && !an.ic_given[node] ? an.ns_val[node] : an.ic_val[node]) : 0.0; };
            if (!given(in.icvbs)) c.icVBS = icv(in.b) - icv(in.s);
            if (!given(in.icvds)) c.icVDS = icv(in.d) - icv(in.s);
            if (!given(in.icvgs)) c.icVGS = icv(in.g) - icv(in.s);
        }
        b->mos_raw[MR_vto][i] = md.vto; b->mos_raw[MR_kp][i] = (c.level == 6) ? md.kc : md.kp;
        b->mos_raw[MR_gamma][i] = md.gamma; b->mos_raw[MR_phi][i] = md.phi;
        b->mos_raw[MR_lambda][i] = given(md.lambda) ? md.lambda : 0;
        b->mos_raw[MR_l][i] = given(in.l) ? in.l : (o.defl > 0 ? o.defl : 1e-6);
        b->mos_raw[MR_w][i] = given(in.w) ? in.w : (o.defw > 0 ? o.defw : 1e-6);
    }
    b->cap_c.resize(dk->ncap); b->cap_ic.resize(dk->ncap);
    b->cap_craw.resize(dk->ncap); b->cap_m.resize(dk->ncap);
    for (int i = 0; i < dk->ncap; i++) {
        const b200_cap &in = dk->cap[i];
        double mm = given(in.m) ? in.m : 1.0;
        b->cap_c[i] = in.c*1.0*mm;                  // captemp.cc:98
        b->cap_craw[i] = in.c*1.0; b->cap_m[i] = mm;
        if (mm != 1.0) b->cap_has_m = true;
        double ic = given(in.ic) ? in.ic : 0.0;
        if (dk->uic && !given(in.ic)) {
            auto icv = [&](int node) { return node ? (an.ns_given[node] && !an.ic_given[node] ? an.ns_val[node] : an.ic_val[node]) : 0.0; };
            ic = icv(in.pos) - icv(in.neg);         // capic.cc:52-71
        }
        b->cap_ic[i] = ic;
    }
    b->ind_l.resize(dk->nind); b->ind_ic.resize(dk->nind);
    for (int i = 0; i < dk->nind; i++) {
        b->ind_l[i] = dk->ind[i].l;                 // indset.cc:145 INDnomInduct/INDm (m = 1)
        b->ind_ic[i] = given(dk->ind[i].ic) ? dk->ind[i].ic : 0.0;
    }
    b->src_dc = b->an.src.dc;
    b->jj_ics = b->an.jj.ics0;
    b->tjm_ics = b->an.tjm.ics0;
}

int b200_batch_create(const b200_deck *deck, int ntrials, int device, b200_batch **out)
{
    *out = nullptr;
    if (!deck || ntrials <= 0) return B200_E_BADPARM;
    if (deck->opt.method != 0 && deck->opt.method != 1) return B200_E_BADPARM;   // TRAPEZOIDAL | GEAR
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev) {
        fprintf(stderr, "b200spice: no CUDA device (no CPU fallback exists)\n");
        return B200_E_NOGPU;
    }
    if (deck->opt.steptype != 0 ||
            deck->opt.gminsteps > 0 || deck->opt.srcsteps > 0) {
        fprintf(stderr, "b200spice: option not built yet (steptype/spice3 stepping)\n");
        return B200_E_UNSUPP;
    }
    CK(cudaSetDevice(device));
    b200_batch *b = new b200_batch;
    b->device = device;
    b->ntrials = ntrials;
    b->ldt = (ntrials + 31) & ~31;
    b->uic = deck->uic != 0;
    std::string err;
    int rc = b200_analyze(deck, b->an, err);
    if (rc) { fprintf(stderr, "b200spice: %s\n", err.c_str()); delete b; return rc; }
    rc = b200_build_programs(b->an, deck->uic != 0);
    if (rc) {
        fprintf(stderr, "b200spice: LU (%d values) + rhs (%d) of one trial exceed the 16-bit slot "
            "indices of the solve program\n", b->an.nnzLU, b->an.n);
        delete b; return rc;
    }
    build_cards(b, deck);
    Analysis &an = b->an;
    CK(cudaStreamCreate(&b->stream));

    Consts &K = b->K;
    memset(&K, 0, sizeof(K));
    K.opt = deck->opt;
    K.n = an.n; K.nnz = an.nnz; K.numStates = an.numStates; K.ldt = b->ldt; K.ntrials = ntrials;
    K.NS = deck->opt.maxord + 2; if (K.NS > 8) K.NS = 8; if (K.NS < 4) K.NS = 4;
    K.nnzLU = an.nnzLU; K.ncslot = an.ncslot; K.nrslot = an.nrslot; K.NC = an.NC;
    K.t_stop = an.t_stop; K.t_step = an.t_step; K.t_start = an.t_start; K.maxStep = an.maxStep;
    K.delmin = an.delmin; K.minBreak = an.minBreak; K.break0 = an.break0; K.break1 = an.break1;
    K.nocut = an.nocut; K.uic = deck->uic; K.hadNodeset = an.hadNodeset;
    K.npoints_max = an.npoints_max; K.method = deck->opt.method;
    K.nbrks = (int)an.brks.size(); K.nlat = (int)an.lat_offs.size();

    size_t ldt = b->ldt;
    // ---- hot arrays ----
    HotView &h = b->hot;
    h.mode = b->dalloc<int>(ldt); h.order = b->dalloc<int>(ldt); h.st_base = b->dalloc<int>(ldt);
    h.noncon = b->dalloc<int>(ldt); h.trapbad = b->dalloc<int>(ldt); h.trapcheck = b->dalloc<int>(ldt);
    h.active = b->dalloc<int>(ldt); h.do_solve = b->dalloc<int>(ldt); h.iterno = b->dalloc<int>(ldt);
    h.time = b->dalloc<double>(ldt); h.delta = b->dalloc<double>(ldt);
    h.srcfact = b->dalloc<double>(ldt); h.diaggmin = b->dalloc<double>(ldt);
    h.ag = b->dalloc<double>(7*ldt); h.pred = b->dalloc<double>(2*ldt);
    h.deltaOld = b->dalloc<double>(7*ldt);
    h.nodeconv = b->dalloc<int>(ldt); h.solve_err = b->dalloc<int>(ldt);
    h.need_devconv = b->dalloc<int>(ldt); h.devconv = b->dalloc<int>(ldt);
    h.accel_exit = b->dalloc<int>(ldt);
    h.lte = b->dalloc<double>(2*ldt); h.sync = b->dalloc<int>(ldt);
    h.lte_rcp = b->dalloc<double>((size_t)LR_NROWS*ldt);
    // ---- vectors ----
    VecView &V = b->V;
    size_t vplane = (size_t)(an.n + 1)*ldt;
    V.X[0] = b->dalloc<double>(vplane); V.X[1] = b->dalloc<double>(vplane);
    V.spare = b->dalloc<double>(vplane); V.sols = b->dalloc<double>(8*vplane);
    V.oldSol = b->dalloc<double>(vplane);
    V.states = b->dalloc<double>((size_t)K.NS*an.numStates*ldt);
    V.oldState0 = b->dalloc<double>((size_t)an.numStates*ldt);
    V.contrib = b->dalloc<double>(((size_t)an.NC + an.long_chunks.size())*ldt);   // [NC (+ long-run chunk sums)][ldt]
    V.out = nullptr;                                     // sized at the first run: the save list may shrink it
    V.outslot = nullptr;
    K.nout = an.n;
    if (!V.states || !V.sols) { b200_batch_destroy(b); return B200_E_NOMEM; }
    // ---- topology ----
    MosView &M = b->M; MosTopo &mt = an.mos;
    M.n = mt.n;
    M.d = b->dupload(mt.d); M.g = b->dupload(mt.g); M.s = b->dupload(mt.s); M.b = b->dupload(mt.b);
    M.dp = b->dupload(mt.dp); M.sp = b->dupload(mt.sp); M.type = b->dupload(mt.type);
    M.off = b->dupload(mt.off); M.state = b->dupload(mt.state); M.level = b->dupload(mt.level);
    M.has_level6 = 0;              // any device that is not level 1
    for (int lv : mt.level) if (lv != 1) M.has_level6 = 1;
    M.cslot = b->dupload(mt.cslot); M.rslot = b->dupload(mt.rslot);
    M.cache = b->dalloc<double>((size_t)MC_NCACHE*mt.n*ldt);
    M.save = b->dalloc<double>((size_t)MC_NCACHE*mt.n*ldt);
    M.lte = b->dalloc<double>((size_t)2*mt.n*ldt);
    b->d_cards = b->dupload(b->cards);
    CapView &Cp = b->Cp; CapTopo &ct = an.cap;
    Cp.n = ct.n;
    Cp.pos = b->dupload(ct.pos); Cp.neg = b->dupload(ct.neg); Cp.state = b->dupload(ct.state);
    Cp.cslot = b->dupload(ct.cslot); Cp.rslot = b->dupload(ct.rslot);
    Cp.geq = b->dalloc<double>((size_t)ct.n*ldt); Cp.ceq = b->dalloc<double>((size_t)ct.n*ldt);
    IndView &Li = Cp.L; IndTopo &lt = an.ind;
    Li.n = lt.n;
    Li.branch = b->dupload(lt.branch); Li.state = b->dupload(lt.state);
    Li.cslot = b->dupload(lt.cslot); Li.rslot = b->dupload(lt.rslot);
    Li.req = b->dalloc<double>((size_t)lt.n*ldt); Li.veq = b->dalloc<double>((size_t)lt.n*ldt);
    Li.prevFlux = b->dalloc<double>((size_t)lt.n*ldt);
    Li.mptr = b->dupload(lt.mptr); Li.mother = b->dupload(lt.mother); Li.mk = b->dupload(lt.mk);
    Li.nmut = an.mut.n;
    Li.mut1 = b->dupload(an.mut.ind1); Li.mut2 = b->dupload(an.mut.ind2);
    Li.mut_cslot = b->dupload(an.mut.cslot); Li.mutk = b->dupload(an.mut.k);
    JJView &Jv = Cp.J; JJTopo &jt = an.jj;
    memset(&Jv, 0, sizeof(Jv));
    Jv.n = jt.n; Jv.nmodels = jt.nmodels;
    if (jt.n) {
        Jv.pos = b->dupload(jt.pos); Jv.neg = b->dupload(jt.neg); Jv.phs = b->dupload(jt.phs);
        Jv.model = b->dupload(jt.model); Jv.state = b->dupload(jt.state);
        Jv.cslot = b->dupload(jt.cslot); Jv.rslot = b->dupload(jt.rslot);
        Jv.vless = b->dupload(jt.vless); Jv.vmore = b->dupload(jt.vmore); Jv.vg = b->dupload(jt.vg);
        Jv.initV = b->dupload(jt.initV); Jv.initPhi = b->dupload(jt.initPhi);
        Jv.dvdt = b->dalloc<double>((size_t)jt.n*ldt);
        b->d_jj_area0 = b->dupload(jt.area0); b->d_jj_vshunt = b->dupload(jt.vshunt);
        b->d_jj_tcf = b->dupload(jt.tcf);
        K.njjmod = jt.nmodels;
        for (int m = 0; m < jt.nmodels; m++) {
            const b200_jj_model &md = jt.models[m];
            Jv.rtype[m] = md.rtype; Jv.ictype[m] = md.ictype; Jv.pi[m] = md.pi;
            Jv.delv[m] = md.delv; Jv.icfactor[m] = md.icfactor;
            K.jj_limits[m] = md.ictype != 0;
            K.jj_vth[m] = md.vdpbak/md.tsaccl;
            K.jj_tsfact[m] = md.tsfact;
        }
    }
    TJMView &Tv = Cp.T; TJMTopo &tt = an.tjm;
    memset(&Tv, 0, sizeof(Tv));
    Tv.n = tt.n; Tv.nmodels = tt.nmodels; Tv.nterms = tt.nterms; Tv.model0 = jt.nmodels;
    if (tt.n) {
        Tv.pos = b->dupload(tt.pos); Tv.neg = b->dupload(tt.neg); Tv.phs = b->dupload(tt.phs);
        Tv.model = b->dupload(tt.model); Tv.state = b->dupload(tt.state);
        Tv.cslot = b->dupload(tt.cslot); Tv.rslot = b->dupload(tt.rslot);
        Tv.initV = b->dupload(tt.initV); Tv.initPhi = b->dupload(tt.initPhi);
        Tv.omegaJ = b->dupload(tt.omegaJ); Tv.kgap = b->dupload(tt.kgap); Tv.kgap_rejpt = b->dupload(tt.kgap_rejpt);
        Tv.coef = b->dupload(tt.coef);
        Tv.dvdt = b->dalloc<double>((size_t)tt.n*ldt);
        Tv.sc_old = b->dalloc<double>((size_t)2*tt.n*ldt);
        Tv.work = b->dalloc<double>((size_t)TF_NARR*tt.nterms*2*tt.n*ldt);
        if (!Tv.work) { b200_batch_destroy(b); return B200_E_NOMEM; }
        b->d_tjm_area0 = b->dupload(tt.area0); b->d_tjm_vshunt = b->dupload(tt.vshunt);
        b->d_tjm_tcf = b->dupload(tt.tcf); b->d_tjm_alphaN = b->dupload(tt.alphaN);
        std::vector<char> seen(tt.nmodels, 0);
        for (int oi = 0; oi < deck->norder; oi++) {         // TJMdev::accept: the floor is the first
            if (deck->order_kind[oi] != B200_DEV_TJM) continue;   // visited instance's vdpbak/tsaccl
            int i = deck->order_index[oi], m = tt.model[i];
            if (seen[m]) continue;
            seen[m] = 1;
            const b200_tjm_model &md = tt.models[m];
            Tv.rtype[m] = md.rtype; Tv.ictype[m] = md.ictype; Tv.nt[m] = md.nterms;
            K.jj_limits[jt.nmodels + m] = 1;
            K.jj_vth[jt.nmodels + m] = tt.vdpbak[i]/md.tsaccl;
            K.jj_tsfact[jt.nmodels + m] = md.tsfact;
        }
        K.njjmod = jt.nmodels + tt.nmodels;
    }
    ResPhView &Rv = Cp.R;
    memset(&Rv, 0, sizeof(Rv));
    Rv.n = (int)an.rph_res.size();
    if (Rv.n) { Rv.cslot = b->dupload(an.rph_cslot); Rv.cls = b->dupload(an.rph_cls); Rv.res = b->dupload(an.rph_res); }
    b->d_resm = b->dupload(an.res_m);
    Rv.m = b->d_resm; Rv.resmin = an.resmin;
    {
        // per-trial resistor values (B200_P_RES_R): which preloaded constants to re-sum
        std::vector<int> cs(an.rt_entry.size());
        for (size_t e = 0; e < cs.size(); e++) cs[e] = an.ainit_cslot[an.rt_entry[e]];
        b->d_rt_cslot = b->dupload(cs); b->d_rt_base = b->dupload(an.rt_base);
        b->d_rt_ptr = b->dupload(an.rt_ptr); b->d_rt_res = b->dupload(an.rt_res); b->d_rt_sign = b->dupload(an.rt_sign);
    }
    K.jjDCphase = an.jjDCphase;
    K.jjaccel = (deck->opt.jjaccel && (an.jj.n + an.tjm.n) > 0) ? 1 : 0;
    V.phase = an.jjDCphase ? b->dupload(an.phase_mask) : nullptr;
    K.initDelta = an.initDelta;
    h.jj_vmax = b->dalloc<double>((size_t)std::max(jt.nmodels + tt.nmodels, 1)*ldt);
    h.accepted = b->dalloc<int>(ldt);
    SrcView &S = b->S; SrcTopo &st = an.src;
    S.n = st.n;
    S.kind = b->dupload(st.kind); S.fn = b->dupload(st.fn);
    b->src_dcgiven = st.dcgiven;
    b->d_src_dcgiven = b->dupload(st.dcgiven);
    S.dcgiven = b->d_src_dcgiven;
    S.rslot = b->dupload(st.rslot); S.coef_off = b->dupload(st.coef_off); S.ncoef = b->dupload(st.ncoef);
    S.pool = b->dupload(st.pool);
    S.dcval = nullptr;
    // ---- charges for the LTE kernel (mostrun.cc:59-61, captrun.cc:58) ----
    {
        std::vector<int> q;
        for (int oi = 0; oi < deck->norder; oi++) {
            int i = deck->order_index[oi];
            if (deck->order_kind[oi] == B200_DEV_MOS) {
                q.push_back(mt.state[i] + MS_qgs); q.push_back(mt.state[i] + MS_qgd);
                q.push_back(mt.state[i] + MS_qgb);
            }
            else if (deck->order_kind[oi] == B200_DEV_CAP)
                q.push_back(ct.state[i]);
        }
        b->nq = (int)q.size();
        b->d_qslots = b->dupload(q);
    }
    // ---- solve view ----
    SolveView &Sv = b->Sv;
    memset(&Sv, 0, sizeof(Sv));
    Sv.cls = an.primary_class;
    Sv.ntype = b->dupload(an.ntype);
    Sv.ic_given = b->dupload(an.ic_given); Sv.ns_given = b->dupload(an.ns_given);
    Sv.ic_val = b->dupload(an.ic_val); Sv.ns_val = b->dupload(an.ns_val);
    Sv.yperm_out = b->dupload(an.yperm_out);
    Sv.runs = b->dupload(an.runs);
    Sv.tile_order = b->dupload(an.tile_order); Sv.ntiles = (int)an.tile_order.size();
    Sv.img = an.img; Sv.nlu = an.nnzLU; Sv.npad = (an.n + 31) & ~31;
    Sv.rows = b->dupload(an.prog_rows); Sv.rounds = b->dupload(an.prog_round);
    Sv.wide = an.wide;
    Sv.long_chunks = b->dupload(an.long_chunks); Sv.nlong = (int)an.long_chunks.size(); Sv.NC = an.NC;
    Sv.round_row = b->dupload(an.round_row); Sv.level_ptr = b->dupload(an.level_ptr); Sv.nlevels = (int)an.level_ptr.size() - 1;
    Sv.nrows = (int)(an.prog_rows.size()/(LU_LANES*(an.wide ? 2 : 1))); Sv.nrounds = (int)an.prog_round.size();
    Sv.image = b->dalloc<double>((size_t)ldt*std::max(an.img, an.has_tranp ? an.tranp.img : 0));
    Sv.ysol = b->dalloc<double>((size_t)ldt*Sv.npad);
    Sv.lu_bad = b->dalloc<int>(ldt); Sv.fin_flags = b->dalloc<int>(ldt); Sv.live = b->dalloc<int>(ldt); Sv.ctr = b->dalloc<int>(2);
    if (!Sv.image || !Sv.ysol || !Sv.live) { b200_batch_destroy(b); return B200_E_NOMEM; }
    rc = solve_configure(Sv);
    if (rc) {
        if (rc == B200_E_UNSUPP)
            fprintf(stderr, "b200spice: LU (%d values) + rhs (%d) of one trial exceed shared memory; "
                "the supernodal large-matrix path is not built yet\n", an.nnzLU, an.n);
        b200_batch_destroy(b); return rc;
    }
    if (an.has_tranp) {
        // the transient program: same contribution stream, image / ysol buffers shared with the
        // DC program (a trial is in one class per tick), its own live list and counters
        const LuProgram &P = an.tranp;
        SolveView &S2 = b->Sv2;
        S2 = Sv;
        S2.cls = 0;
        S2.yperm_out = b->dupload(P.yperm_out);
        S2.runs = b->dupload(P.runs);
        S2.tile_order = b->dupload(P.tile_order); S2.ntiles = (int)P.tile_order.size();
        S2.img = P.img; S2.nlu = P.nnzLU;
        S2.rows = b->dupload(P.prog_rows); S2.rounds = b->dupload(P.prog_round);
        S2.wide = P.wide;
        S2.long_chunks = Sv.long_chunks; S2.nlong = Sv.nlong; S2.NC = Sv.NC;
        S2.round_row = b->dupload(P.round_row); S2.level_ptr = b->dupload(P.level_ptr); S2.nlevels = (int)P.level_ptr.size() - 1;
        S2.nrows = (int)(P.prog_rows.size()/(LU_LANES*(P.wide ? 2 : 1))); S2.nrounds = (int)P.prog_round.size();
        S2.live = b->dalloc<int>(ldt); S2.ctr = b->dalloc<int>(2);
        rc = solve_configure(S2);
        if (rc) { b200_batch_destroy(b); return rc; }
        b->two_programs = true;
    }
    b->nconst = (int)an.const_idx.size();
    b->d_const_idx = b->dupload(an.const_idx); b->d_const_val = b->dupload(an.const_val);
    // ---- control ----
    b->ctl = b->dalloc<TrialCtl>(ldt);
    CtlAux &ax = b->aux;
    ax.vecops = b->dalloc<int>(ldt); ax.sols_base = b->dalloc<int>(ldt);
    ax.vo_stbase = b->dalloc<int>(ldt); ax.ip_n = b->dalloc<int>(ldt);
    ax.ip_out0 = b->dalloc<int>(ldt); ax.ip_degree = b->dalloc<int>(ldt);
    ax.ip_d = b->dalloc<double>(5*ldt);
    CtlShared &sh = b->sh;
    sh.K = nullptr;
    sh.brks = b->dupload(an.brks); sh.lat_offs = b->dupload(an.lat_offs);
    sh.lat_per = b->dupload(an.lat_per);
    {
        // sources that bound the step through IFtranData::time_limit (srctrun.cc:45-66):
        // SIN after its delay, SFFM / AM always (period of the sum frequency), EXP inside its
        // two exponential edges, SPULSE after its delay.  fr: the bound itself, dphimax / omega
        std::vector<double> tdl, fr, etd, etau;
        for (int i = 0; i < st.n; i++) {
            const double *pp = st.pool.data() + st.coef_off[i];
            if (st.fn[i] == B200_FN_SIN) { tdl.push_back(pp[3]); fr.push_back((.2*M_PI)/(2*M_PI*pp[2])); }
            else if (st.fn[i] == B200_FN_SFFM) { tdl.push_back(-HUGE_VAL); fr.push_back((.2*M_PI)/(2*M_PI*(pp[2] + pp[4]))); }
            else if (st.fn[i] == B200_FN_AM) { tdl.push_back(-HUGE_VAL); fr.push_back((.2*M_PI)/(2*M_PI*(pp[3] + pp[2]))); }
            else if (st.fn[i] == B200_FN_SPULSE) { tdl.push_back(pp[3]); fr.push_back((.2*M_PI)*pp[2]/(2*M_PI)); }    // inptran.cc:2861-2870
            else if (st.fn[i] == B200_FN_EXP) {
                etd.push_back(pp[2]); etau.push_back(pp[3]);
                etd.push_back(pp[4]); etau.push_back(pp[5]);
            }
        }
        sh.nsin = (int)tdl.size();
        sh.sin_tdl = b->dupload(tdl); sh.sin_per = b->dupload(fr);
        sh.nexp = (int)etd.size();
        sh.exp_td = b->dupload(etd); sh.exp_tau = b->dupload(etau);
    }
    b->d_count = b->dalloc<int>(1);
    b->d_counters = b->dalloc<int>((size_t)ldt*B200_NCOUNTERS);
    b->d_npts = b->dalloc<int>(ldt);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (getenv("B200_VERBOSE")) {
        int maxrun = 0;
        for (auto &r : an.runs) maxrun = std::max(maxrun, r.len);
        fprintf(stderr, "b200spice: n=%d nnz=%d nnzLU=%d states=%d | contributions NC=%d (%d constants, longest run %d)\n"
            "  image %d doubles (%d scratch) | program: %d tasks, %lld products, %d levels, %d rounds, %lld steps, %d rows (%zu B)\n"
            "  k_lu: %d warps/CTA x %d CTAs, %d B shared, program in %s memory\n",
            an.n, an.nnz, an.nnzLU, an.numStates, an.NC, b->nconst, maxrun,
            an.img, an.nscratch, an.prog_tasks, an.prog_ops, an.prog_levels, Sv.nrounds, an.prog_steps,
            Sv.nrows, an.prog_rows.size()*sizeof(unsigned),
            Sv.lu_warps, Sv.lu_grid, Sv.lu_smem, Sv.prog_in_smem ? "shared" : "global");
    }
    *out = b;
    return B200_OK;
}

void b200_batch_destroy(b200_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    for (void *p : b->allocs) cudaFree(p);
    if (b->d_wout) cudaFree(b->d_wout);
    if (b->d_wsel) cudaFree(b->d_wsel);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

int b200_batch_set_param(b200_batch *b, int param, int index, const double *values)
{
    if (!b || !values) return B200_E_BADPARM;
    {
        // a row given singly replaces a packed row of the same (table, instance)
        int t = param_table(b, param, index);
        if (t >= 0)
            for (size_t r = 0; r < b->pk_tbl.size(); r++)
                if (b->pk_tbl[r] == t && b->pk_idx[r] == index) b->pk_idx[r] = -1;
    }
    // one copy, into the vector the batch already owns (repeated runs reuse its storage)
    auto put = [&](std::vector<double> &dst) { dst.assign(values, values + b->ntrials); };
    int raw = -1;
    switch (param) {
    case B200_P_MOS_VTO: raw = MR_vto; break;
    case B200_P_MOS_KP: raw = MR_kp; break;
    case B200_P_MOS_GAMMA: raw = MR_gamma; break;
    case B200_P_MOS_PHI: raw = MR_phi; break;
    case B200_P_MOS_LAMBDA: raw = MR_lambda; break;
    case B200_P_MOS_L: raw = MR_l; break;
    case B200_P_MOS_W: raw = MR_w; break;
    case B200_P_CAP_C:
        if (index < 0 || index >= b->an.cap.n) return B200_E_BADPARM;
        put(b->cap_var[index]); b->params_dirty = true; return B200_OK;
    case B200_P_IND_L:
        if (index < 0 || index >= b->an.ind.n) return B200_E_BADPARM;
        put(b->ind_var[index]); b->params_dirty = true; return B200_OK;
    case B200_P_SRC_DC:
        if (index < 0 || index >= b->an.src.n) return B200_E_BADPARM;
        b->src_dcgiven[index] = 1;                  // a DC value makes it given (SRCdcGiven)
        put(b->src_var[index]); b->params_dirty = true; return B200_OK;
    case B200_P_JJ_ICS:
        if (index < 0 || index >= b->an.jj.n) return B200_E_BADPARM;
        put(b->jj_var[index]); b->params_dirty = true; return B200_OK;
    case B200_P_TJM_ICS:
        if (index < 0 || index >= b->an.tjm.n) return B200_E_BADPARM;
        put(b->tjm_var[index]); b->params_dirty = true; return B200_OK;
    case B200_P_RES_R:
        if (index < 0 || index >= b->an.nres) return B200_E_BADPARM;
        put(b->res_var[index]); b->params_dirty = true; return B200_OK;
    default:
        return B200_E_UNSUPP;
    }
    if (index < 0 || index >= b->an.mos.n) return B200_E_BADPARM;
    // a per-trial model value makes the card "given" for that device
    MosCard &c = b->cards[index];
    if (raw == MR_vto) c.g_vto = 1;
    if (raw == MR_kp) c.g_kp = 1;
    if (raw == MR_gamma) c.g_gamma = 1;
    if (raw == MR_phi) c.g_phi = 1;
    put(b->mos_var[raw][index]);
    b->params_dirty = true;
    return B200_OK;
}

int b200_batch_set_params(b200_batch *b, int nrows, const int *param, const int *index,
    const double *values)
{
    if (!b || nrows < 0 || (nrows && (!param || !index || !values))) return B200_E_BADPARM;
    CK(cudaSetDevice(b->device));
    std::vector<int> tbl(nrows);
    for (int r = 0; r < nrows; r++) {
        int t = param_table(b, param[r], index[r]);
        if (t < 0) return -t;
        tbl[r] = t;
    }
    for (int r = 0; r < nrows; r++) {
        if (tbl[r] < MR_NRAW) {
            // a per-trial model value makes the card "given" for that device
            MosCard &c = b->cards[index[r]];
            if (tbl[r] == MR_vto) c.g_vto = 1;
            if (tbl[r] == MR_kp) c.g_kp = 1;
            if (tbl[r] == MR_gamma) c.g_gamma = 1;
            if (tbl[r] == MR_phi) c.g_phi = 1;
            b->mos_var[tbl[r]].erase(index[r]);
        }
        else if (tbl[r] == PT_CAP) b->cap_var.erase(index[r]);
        else if (tbl[r] == PT_IND) b->ind_var.erase(index[r]);
        else if (tbl[r] == PT_JJ) b->jj_var.erase(index[r]);
        else if (tbl[r] == PT_TJM) b->tjm_var.erase(index[r]);
        else if (tbl[r] == PT_RES) b->res_var.erase(index[r]);
        else { b->src_var.erase(index[r]); b->src_dcgiven[index[r]] = 1; }    // a DC value makes it given
    }
    size_t need = (size_t)nrows*b->ntrials;
    if (need > b->pack_cap) {
        b->d_pack = b->dalloc<double>(need);
        if (!b->d_pack) return B200_E_NOMEM;
        b->pack_cap = need;
    }
    // ONE transfer for all rows: a DMA when the caller's buffer is page-locked
    if (need) CK(cudaMemcpyAsync(b->d_pack, values, need*sizeof(double), cudaMemcpyHostToDevice, b->stream));
    b->pk_tbl = tbl;
    b->pk_idx.assign(index, index + nrows);
    b->params_dirty = true;
    return B200_OK;
}

// upload parameters and run the temperature kernel (sCKT::temp, ckt.cc:2297-2307)
static int prepare_params(b200_batch *b)
{
    if (!b->params_dirty) return B200_OK;
    size_t ldt = b->ldt;
    int nm = b->an.mos.n;
    cudaStream_t st = b->stream;
    std::vector<char> packed_in(PT_NTABLES, 0);
    for (int t : b->pk_tbl) packed_in[t] = 1;
    TableBases bases;
    for (int k = 0; k < PT_NTABLES; k++) bases.p[k] = nullptr;
    std::vector<void*> tmp_shared;
    auto upload_strided = [&](int tbl, const std::vector<double> &shared,
            const std::map<int, std::vector<double>> &var, double *&dptr, int &ts) -> int {
        int n = (int)shared.size();
        int want_ts = (var.empty() && !packed_in[tbl]) ? 0 : 1;
        if (!dptr || ts != want_ts) {
            dptr = b->dalloc<double>(want_ts ? (size_t)n*ldt : (size_t)n);
            if (!dptr) return B200_E_NOMEM;
            ts = want_ts;
        }
        if (!ts) {
            if (n) CK(cudaMemcpyAsync(dptr, shared.data(), n*sizeof(double), cudaMemcpyHostToDevice, st));
        }
        else {
            // every row starts as the instance's shared value (filled on the device), then the
            // rows given one by one (set_param), then the packed rows (set_params)
            double *dsh = nullptr;
            CK(cudaMalloc(&dsh, (n ? n : 1)*sizeof(double)));
            tmp_shared.push_back(dsh);
            CK(cudaMemcpyAsync(dsh, shared.data(), n*sizeof(double), cudaMemcpyHostToDevice, st));
            k_fill_rows<<<grid_rows((unsigned)((ldt + 255)/256), n), 256, 0, st>>>(dptr, dsh, n, (int)ldt);
            std::vector<double> row(ldt);
            for (auto &kv : var) {
                std::copy(kv.second.begin(), kv.second.end(), row.begin());
                std::fill(row.begin() + b->ntrials, row.end(), kv.second[0]);
                CK(cudaMemcpyAsync(dptr + (size_t)kv.first*ldt, row.data(), ldt*sizeof(double), cudaMemcpyHostToDevice, st));
                CK(cudaStreamSynchronize(st));          // `row` is reused
            }
            bases.p[tbl] = dptr;
        }
        return B200_OK;
    };
    RawParams raw;
    int tsd[MP_NPARAMS];
    for (int k = 0; k < MP_NPARAMS; k++) tsd[k] = 0;
    for (int k = 0; k < MR_NRAW; k++) {
        int rc = upload_strided(k, b->mos_raw[k], b->mos_var[k], b->d_raw[k], b->raw_ts[k]);
        if (rc) return rc;
        raw.p[k].base = b->d_raw[k];
        raw.p[k].ts = b->raw_ts[k];
        raw.p[k].ld = b->raw_ts[k] ? (int)ldt : 1;
        if (b->raw_ts[k])
            for (int q = 0; kDep[k][q] >= 0; q++) tsd[kDep[k][q]] = 1;
    }
    MosParamsOut po;
    for (int k = 0; k < MP_NPARAMS; k++) {
        if (!b->d_prm[k] || b->prm_ts[k] != tsd[k]) {
            b->d_prm[k] = b->dalloc<double>(tsd[k] ? (size_t)nm*ldt : (size_t)nm);
            if (!b->d_prm[k]) return B200_E_NOMEM;
            b->prm_ts[k] = tsd[k];
        }
        po.base[k] = b->d_prm[k]; po.ts[k] = tsd[k];
        b->M.prm[k].base = b->d_prm[k];
        b->M.prm[k].ts = tsd[k];
        b->M.prm[k].ld = tsd[k] ? (int)ldt : 1;
    }
    {
        // staging layout of k_load_mos: per-trial parameters get a row each, the others one
        // slot of the block shared by the CTA
        MosStageMap &sm = b->M.smap;
        sm.nvar = 0;
        // parameters beyond MP_NBASE are read by the level-3 / level-6 equations only: a deck of
        // level-1 devices does not stage their per-trial rows
        bool other = b->M.has_level6 != 0;
        for (int k = 0; k < MP_NPARAMS; k++)
            if (tsd[k] && (other || k < MP_NBASE)) sm.var[sm.nvar++] = k;
        int shared_base = (SR_PRM + sm.nvar)*MOS_TPB;
        int q = 0;
        for (int k = 0; k < MP_NPARAMS; k++) {
            bool row = tsd[k] && (other || k < MP_NBASE);
            sm.ts[k] = row;
            sm.off[k] = row ? (SR_PRM + q++)*MOS_TPB : shared_base + k;
            sm.aoff[k] = sm.off[k]*(int)sizeof(double);
            sm.amask[k] = row ? 0xffffffffu : 0u;
        }
        // + the device's contribution slots (MOS_NSLOT ints) at the end of the area
        b->M.stage_doubles = shared_base + (other ? MP_NPARAMS : MP_NBASE) + (MOS_NSLOT + 1)/2;
        if (loads_configure(b->M)) return B200_E_NOMEM;
    }
    // a per-trial capacitance table holds card values (the shared fill too) and is scaled by the
    // instance multipliers after the rows have landed
    bool cap_rows = !b->cap_var.empty() || packed_in[PT_CAP];
    int rc = upload_strided(PT_CAP, (cap_rows && b->cap_has_m) ? b->cap_craw : b->cap_c, b->cap_var, b->d_capc, b->capc_ts);
    if (rc) return rc;
    b->Cp.capac.base = b->d_capc; b->Cp.capac.ts = b->capc_ts; b->Cp.capac.ld = b->capc_ts ? (int)ldt : 1;
    if (!b->d_capic) b->d_capic = b->dalloc<double>(b->cap_ic.size());
    if (!b->cap_ic.empty())
        CK(cudaMemcpy(b->d_capic, b->cap_ic.data(), b->cap_ic.size()*sizeof(double), cudaMemcpyHostToDevice));
    b->Cp.ic.base = b->d_capic; b->Cp.ic.ts = 0; b->Cp.ic.ld = 1;
    rc = upload_strided(PT_IND, b->ind_l, b->ind_var, b->d_indl, b->indl_ts);
    if (rc) return rc;
    b->Cp.L.induct.base = b->d_indl; b->Cp.L.induct.ts = b->indl_ts; b->Cp.L.induct.ld = b->indl_ts ? (int)ldt : 1;
    if (!b->d_indic) b->d_indic = b->dalloc<double>(b->ind_ic.size());
    if (!b->ind_ic.empty())
        CK(cudaMemcpy(b->d_indic, b->ind_ic.data(), b->ind_ic.size()*sizeof(double), cudaMemcpyHostToDevice));
    b->Cp.L.ic.base = b->d_indic; b->Cp.L.ic.ts = 0; b->Cp.L.ic.ld = 1;
    rc = upload_strided(PT_RES, b->an.res_r, b->res_var, b->d_resr, b->resr_ts);
    if (rc) return rc;
    b->Cp.R.r.base = b->d_resr; b->Cp.R.r.ts = b->resr_ts; b->Cp.R.r.ld = b->resr_ts ? (int)ldt : 1;
    rc = upload_strided(PT_SRC, b->src_dc, b->src_var, b->d_srcdc, b->srcdc_ts);
    if (rc) return rc;
    b->S.dc.base = b->d_srcdc; b->S.dc.ts = b->srcdc_ts; b->S.dc.ld = b->srcdc_ts ? (int)ldt : 1;
    if (b->an.jj.n) {
        rc = upload_strided(PT_JJ, b->jj_ics, b->jj_var, b->d_jjics, b->jjics_ts);
        if (rc) return rc;
        if (b->jjprm_ts != b->jjics_ts) {
            for (int k = 0; k < JP_NPARAMS; k++) {
                b->d_jjprm[k] = b->dalloc<double>(b->jjics_ts ? (size_t)b->an.jj.n*ldt : (size_t)b->an.jj.n);
                if (!b->d_jjprm[k]) return B200_E_NOMEM;
                b->Cp.J.prm[k].base = b->d_jjprm[k];
                b->Cp.J.prm[k].ts = b->jjics_ts;
                b->Cp.J.prm[k].ld = b->jjics_ts ? (int)ldt : 1;
            }
            b->jjprm_ts = b->jjics_ts;
        }
    }
    if (b->an.tjm.n) {
        rc = upload_strided(PT_TJM, b->tjm_ics, b->tjm_var, b->d_tjmics, b->tjmics_ts);
        if (rc) return rc;
        if (b->tjmprm_ts != b->tjmics_ts) {
            for (int k = 0; k < TP_NPARAMS; k++) {
                b->d_tjmprm[k] = b->dalloc<double>(b->tjmics_ts ? (size_t)b->an.tjm.n*ldt : (size_t)b->an.tjm.n);
                if (!b->d_tjmprm[k]) return B200_E_NOMEM;
                b->Cp.T.prm[k].base = b->d_tjmprm[k];
                b->Cp.T.prm[k].ts = b->tjmics_ts;
                b->Cp.T.prm[k].ld = b->tjmics_ts ? (int)ldt : 1;
            }
            b->tjmprm_ts = b->tjmics_ts;
        }
    }
    // packed rows into their tables (after the default fill, before the temperature kernel)
    if (!b->pk_tbl.empty()) {
        int nrows = (int)b->pk_tbl.size();
        if ((size_t)nrows > b->pk_desc_cap) {
            b->d_pk_desc = b->dalloc<int>(2*(size_t)nrows);
            if (!b->d_pk_desc) return B200_E_NOMEM;
            b->pk_desc_cap = nrows;
        }
        std::vector<int> desc(2*(size_t)nrows);
        for (int r = 0; r < nrows; r++) { desc[2*r] = b->pk_tbl[r]; desc[2*r + 1] = b->pk_idx[r]; }
        CK(cudaMemcpyAsync(b->d_pk_desc, desc.data(), desc.size()*sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));              // `desc` is a local
        k_scatter_rows<<<grid_rows((unsigned)((ldt + 255)/256), nrows), 256, 0, st>>>(bases, b->d_pk_desc,
            b->d_pack, nrows, b->ntrials, (int)ldt);
    }
    if (b->capc_ts && b->cap_has_m) {
        double *dm = nullptr;
        CK(cudaMalloc(&dm, b->cap_m.size()*sizeof(double)));
        tmp_shared.push_back(dm);
        CK(cudaMemcpyAsync(dm, b->cap_m.data(), b->cap_m.size()*sizeof(double), cudaMemcpyHostToDevice, st));
        k_scale_rows<<<grid_rows((unsigned)((ldt + 255)/256), (int)b->cap_m.size()), 256, 0, st>>>(b->d_capc, dm,
            (int)b->cap_m.size(), (int)ldt);
    }
    if (!b->src_dcgiven.empty() && b->d_src_dcgiven)
        CK(cudaMemcpyAsync(b->d_src_dcgiven, b->src_dcgiven.data(), b->src_dcgiven.size()*sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b->d_cards, b->cards.data(), b->cards.size()*sizeof(MosCard), cudaMemcpyHostToDevice, st));
    launch_mos_temp(st, nm, b->ntrials, (int)ldt, b->d_cards, raw, po);
    if (b->an.jj.n) {
        const JJTopo &jt = b->an.jj;
        JJTempIn in;
        memset(&in, 0, sizeof(in));
        in.n = jt.n; in.nmodels = jt.nmodels; in.model = b->Cp.J.model;
        in.ics.base = b->d_jjics; in.ics.ts = b->jjics_ts; in.ics.ld = b->jjics_ts ? (int)ldt : 1;
        in.area0 = b->d_jj_area0; in.vshunt = b->d_jj_vshunt; in.tcf = b->d_jj_tcf;
        in.vless = b->Cp.J.vless; in.vmore = b->Cp.J.vmore;
        for (int m = 0; m < jt.nmodels; m++) {
            const b200_jj_model &md = jt.models[m];
            in.criti[m] = md.criti; in.cap[m] = md.cap; in.cmu[m] = md.cmu; in.gmu[m] = md.gmu;
            in.r0[m] = md.r0; in.rn[m] = md.rn; in.icfactor[m] = md.icfactor; in.delv[m] = md.delv;
            in.rtype[m] = md.rtype;
        }
        for (int k = 0; k < JP_NPARAMS; k++) in.out[k] = b->d_jjprm[k];
        in.ts = b->jjics_ts;
        launch_jj_temp(st, in, b->ntrials, (int)ldt);
    }
    if (b->an.tjm.n) {
        const TJMTopo &tt = b->an.tjm;
        TJMTempIn in;
        memset(&in, 0, sizeof(in));
        in.n = tt.n; in.model = b->Cp.T.model;
        in.ics.base = b->d_tjmics; in.ics.ts = b->tjmics_ts; in.ics.ld = b->tjmics_ts ? (int)ldt : 1;
        in.area0 = b->d_tjm_area0; in.vshunt = b->d_tjm_vshunt; in.tcf = b->d_tjm_tcf; in.alphaN = b->d_tjm_alphaN;
        for (int m = 0; m < tt.nmodels; m++) {
            const b200_tjm_model &md = tt.models[m];
            in.criti[m] = md.criti; in.cap[m] = md.cap; in.cmu[m] = md.cmu; in.gmu[m] = md.gmu;
            in.r0[m] = md.r0; in.rsint[m] = md.rsint; in.rtype[m] = md.rtype;
        }
        for (int k = 0; k < TP_NPARAMS; k++) in.out[k] = b->d_tjmprm[k];
        in.ts = b->tjmics_ts;
        launch_tjm_temp(st, in, b->ntrials, (int)ldt);
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (void *q : tmp_shared) cudaFree(q);
    b->params_dirty = false;
    return B200_OK;
}

static int reset_state(b200_batch *b)
{
    const Analysis &an = b->an;
    size_t ldt = b->ldt;
    size_t vplane = (size_t)(an.n + 1)*ldt;
    cudaStream_t st = b->stream;
    CK(cudaMemsetAsync(b->V.X[0], 0, vplane*sizeof(double), st));
    CK(cudaMemsetAsync(b->V.X[1], 0, vplane*sizeof(double), st));
    CK(cudaMemsetAsync(b->V.spare, 0, vplane*sizeof(double), st));
    CK(cudaMemsetAsync(b->V.sols, 0, 8*vplane*sizeof(double), st));
    CK(cudaMemsetAsync(b->V.states, 0, (size_t)b->K.NS*an.numStates*ldt*sizeof(double), st));
    CK(cudaMemsetAsync(b->V.contrib, 0, (size_t)an.NC*ldt*sizeof(double), st));
    launch_fill_const(st, b->K, b->V.contrib, b->nconst, b->d_const_idx, b->d_const_val);
    if (b->resr_ts)
        launch_res_ainit(st, (int)an.rt_entry.size(), (int)ldt, b->d_rt_cslot, b->d_rt_base, b->d_rt_ptr, b->d_rt_res,
            b->d_rt_sign, b->Cp.R.r, b->d_resm, an.resmin, b->V.contrib);
    CK(cudaMemsetAsync(b->M.cache, 0, (size_t)MC_NCACHE*an.mos.n*ldt*sizeof(double), st));
    // MOSmode = 1, MOSvon = 0 (mosset.cc:283-284)
    launch_fill(st, b->M.cache + (size_t)MC_mode*an.mos.n*ldt, (size_t)an.mos.n*ldt, 1.0);
    CK(cudaMemsetAsync(b->Cp.geq, 0, (size_t)an.cap.n*ldt*sizeof(double), st));
    CK(cudaMemsetAsync(b->Cp.ceq, 0, (size_t)an.cap.n*ldt*sizeof(double), st));
    CK(cudaMemsetAsync(b->Cp.L.req, 0, (size_t)an.ind.n*ldt*sizeof(double), st));
    CK(cudaMemsetAsync(b->Cp.L.veq, 0, (size_t)an.ind.n*ldt*sizeof(double), st));
    CK(cudaMemsetAsync(b->Cp.L.prevFlux, 0, (size_t)an.ind.n*ldt*sizeof(double), st));
    CK(cudaMemsetAsync(b->ctl, 0, ldt*sizeof(TrialCtl), st));
    int *hz[] = { b->hot.mode, b->hot.order, b->hot.st_base, b->hot.noncon, b->hot.trapbad,
        b->hot.trapcheck, b->hot.active, b->hot.do_solve, b->hot.iterno, b->hot.nodeconv,
        b->hot.solve_err, b->hot.need_devconv, b->hot.devconv, b->hot.sync, b->aux.vecops };
    for (int *p : hz) CK(cudaMemsetAsync(p, 0, ldt*sizeof(int), st));
    b->cur = 0;
    if (b->uic && b->K.analysis == 0) {
        // ic() + swaprhs: CKTrhsOld starts as the .ic/.nodeset vector (ckt.cc:908-925, tranan.cc:357-362)
        std::vector<double> row(ldt);
        for (int i = 1; i <= an.n; i++) {
            double v = 0.0;
            if (an.ns_given[i]) v = an.ns_val[i];
            if (an.ic_given[i]) v = an.ic_val[i];
            if (v != 0.0) {
                std::fill(row.begin(), row.end(), v);
                CK(cudaMemcpyAsync(b->V.X[0] + (size_t)i*ldt, row.data(), ldt*sizeof(double),
                    cudaMemcpyHostToDevice, st));
                CK(cudaStreamSynchronize(st));
            }
        }
        for (int k = 0; k < an.ind.n; k++) {
            // INDdev::getic (indic.cc:62-63): a given ic is the branch current
            if (an.ind.icgiven[k] && an.ind.ic[k] != 0.0) {
                std::fill(row.begin(), row.end(), an.ind.ic[k]);
                CK(cudaMemcpyAsync(b->V.X[0] + (size_t)an.ind.branch[k]*ldt, row.data(),
                    ldt*sizeof(double), cudaMemcpyHostToDevice, st));
                CK(cudaStreamSynchronize(st));
            }
        }
    }
    return B200_OK;
}

static int run_analysis(b200_batch *b, int analysis);
int b200_batch_run_tran(b200_batch *b) { return run_analysis(b, 0); }
int b200_batch_run_op(b200_batch *b) { return run_analysis(b, 1); }

namespace {
// the CUDA events of one run: destroyed on every exit path
struct RunEvents {
    static const int WIN = 32;
    cudaEvent_t ev[WIN][4], begin = nullptr, end = nullptr;
    bool ok = true;
    RunEvents() {
        for (int i = 0; i < WIN; i++) for (int k = 0; k < 4; k++) { ev[i][k] = nullptr; ok = ok && cudaEventCreate(&ev[i][k]) == cudaSuccess; }
        ok = ok && cudaEventCreate(&begin) == cudaSuccess && cudaEventCreate(&end) == cudaSuccess;
    }
    ~RunEvents() {
        for (int i = 0; i < WIN; i++) for (int k = 0; k < 4; k++) if (ev[i][k]) cudaEventDestroy(ev[i][k]);
        if (begin) cudaEventDestroy(begin);
        if (end) cudaEventDestroy(end);
    }
};
}

static int run_analysis(b200_batch *b, int analysis)
{
    if (!b) return B200_E_BADPARM;
    CK(cudaSetDevice(b->device));
    b->K.analysis = analysis;
    int rc = prepare_params(b);
    if (rc) return rc;
    if (!b->V.out) {
        b->V.out = b->dalloc<double>((size_t)b->an.npoints_max*b->K.nout*b->ldt);
        if (!b->V.out) return B200_E_NOMEM;
    }
    rc = reset_state(b);
    if (rc) return rc;
    cudaStream_t st = b->stream;
    const Consts &K = b->K;
    memset(&b->stats, 0, sizeof(b->stats));

    const int WIN = RunEvents::WIN;
    RunEvents E;
    if (!E.ok) return B200_E_NOGPU;
    long long launches = 0, ticks = 0, load_launches = 0, solve_launches = 0;
    double ms_load = 0, ms_solve = 0;
    int nload_kernels = (b->M.n ? 1 : 0) + (b->Cp.n ? 1 : 0) + (b->Cp.L.n ? 1 : 0) + (b->Cp.L.nmut ? 1 : 0) + (b->S.n ? 1 : 0) + (b->Cp.J.n ? 1 : 0) + (b->Cp.T.n ? 1 : 0) + (b->Cp.R.n ? 1 : 0);

    CK(cudaEventRecord(E.begin, st));
    // tick 0: begin + first load
    launch_ctl(st, K, b->hot, b->ctl, b->sh, b->aux, 0);
    launch_vec(st, K, b->V, b->aux, b->M, b->V.X[b->cur]);
    launch_loads(st, K, b->hot, b->V, b->M, b->Cp, b->S, b->V.X[b->cur]);
    launches += 3 + nload_kernels;
    int active = b->ntrials;
    const long long max_ticks = 20000000;
    // one tick; `timed`: bracket the solve and the load group with events (w = slot in the window)
    auto tick = [&](int w, bool timed) -> int {
        if (timed) cudaEventRecord(E.ev[w][0], st);
        if (K.jjaccel) launch_jjaccel_gate(st, K, b->hot);
        int r = launch_solve(st, K, b->hot, b->V, b->Sv, b->V.X[b->cur], b->V.X[b->cur ^ 1]);
        if (!r && b->two_programs)
            r = launch_solve(st, K, b->hot, b->V, b->Sv2, b->V.X[b->cur], b->V.X[b->cur ^ 1]);
        if (r) return r;
        if (timed) cudaEventRecord(E.ev[w][1], st);
        launch_devconv(st, K, b->hot, b->V, b->M, b->V.X[b->cur]);
        b->cur ^= 1;
        launch_ctl(st, K, b->hot, b->ctl, b->sh, b->aux, 1);
        launch_ctl(st, K, b->hot, b->ctl, b->sh, b->aux, 2);
        launch_vec(st, K, b->V, b->aux, b->M, b->V.X[b->cur]);
        if (timed) cudaEventRecord(E.ev[w][2], st);
        launch_loads(st, K, b->hot, b->V, b->M, b->Cp, b->S, b->V.X[b->cur]);
        if (timed) cudaEventRecord(E.ev[w][3], st);
        return 0;
    };
    const int per_tick = 8 + nload_kernels + (b->two_programs ? 3 : 0) + (K.jjaccel ? 1 : 0);   // 3 solve kernels (per program), devconv, 2 x ctl, vec, statevec
    // Small systems are launch-bound (a tick of the 177-unknown NOR gate x 2048 trials is 14
    // kernels of a few microseconds each): two ticks -- one per parity of the rhs / rhsOld flip --
    // are captured once into a CUDA graph and the window replays it.  The load / solve split is
    // not timed in this mode.
    bool use_graph = (double)K.ntrials*K.n < 2.0e6 && K.n < 20000;     // (a huge system with few trials has long ticks)
    if (const char *e = getenv("B200_GRAPH")) use_graph = atoi(e) != 0;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    if (use_graph) {
        const int cur0 = b->cur;
        bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        int r1 = 0;
        if (ok) { r1 = tick(0, false); if (!r1) r1 = tick(1, false); }
        if (ok && cudaStreamEndCapture(st, &graph) != cudaSuccess) ok = false;
        if (ok && !r1 && cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) ok = false;
        b->cur = cur0;
        if (!ok || r1 || !gexec) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            graph = nullptr; gexec = nullptr;
            use_graph = false;
            if (r1) return r1;
        }
    }
    while (active > 0 && ticks < max_ticks) {
        if (use_graph) {
            for (int w = 0; w < WIN; w += 2) CK(cudaGraphLaunch(gexec, st));   // two ticks each; cur is back where it was
        }
        else {
            for (int w = 0; w < WIN; w++) {
                rc = tick(w, true);
                if (rc) return rc;
            }
        }
        launches += (long long)per_tick*WIN;
        load_launches += WIN; solve_launches += WIN;
        ticks += WIN;
        launch_count_active(st, K.ntrials, b->hot.active, b->d_count);
        launches++;
        CK(cudaMemcpyAsync(&active, b->d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());                          // a failed launch must not spin the loop
        if (!use_graph)
            for (int w = 0; w < WIN; w++) {
                float a = 0, c = 0;
                cudaEventElapsedTime(&a, E.ev[w][0], E.ev[w][1]);
                cudaEventElapsedTime(&c, E.ev[w][2], E.ev[w][3]);
                ms_solve += a; ms_load += c;
            }
    }
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    launch_export_counters(st, K.ntrials, b->ctl, b->d_counters, b->d_npts);
    CK(cudaEventRecord(E.end, st));
    CK(cudaStreamSynchronize(st));
    float tot = 0;
    cudaEventElapsedTime(&tot, E.begin, E.end);
    CK(cudaGetLastError());
    b->stats.launches = launches; b->stats.ticks = ticks;
    b->stats.ms_total = tot; b->stats.ms_load = ms_load; b->stats.ms_solve = ms_solve;
    b->stats.ms_control = tot - ms_load - ms_solve;
    b->stats.load_launches = load_launches; b->stats.solve_launches = solve_launches;
    b->ran = true;
    return (ticks >= max_ticks) ? B200_E_PANIC : B200_OK;
}

int b200_batch_num_unknowns(const b200_batch *b) { return b ? b->an.n : 0; }
int b200_batch_num_points(const b200_batch *b) { return b ? b->an.npoints_max : 0; }

int b200_batch_unknown_names(const b200_batch *b, char *buf, int buflen)
{
    int pos = 0;
    for (int i = 1; i <= b->an.n; i++) {
        int l = (int)b->an.names[i].size() + 1;
        if (buf && pos + l <= buflen) memcpy(buf + pos, b->an.names[i].c_str(), l);
        pos += l;
    }
    return pos;
}

int b200_batch_get_counters(const b200_batch *b, int *counters)
{
    if (!b || !b->ran) return B200_E_BADPARM;
    CK(cudaSetDevice(b->device));
    CK(cudaMemcpy(counters, b->d_counters, (size_t)b->ntrials*B200_NCOUNTERS*sizeof(int),
        cudaMemcpyDeviceToHost));
    return B200_OK;
}

// unknown indices -> rows of the output buffer (identity without a save list)
static int out_rows(const b200_batch *b, int nsel, const int *sel, std::vector<int> &rows)
{
    rows.assign(sel, sel + nsel);
    for (int &r : rows) {
        if (r < 0 || r >= b->an.n) return B200_E_BADPARM;
        if (!b->save_slot.empty()) {
            r = b->save_slot[r];
            if (r < 0) return B200_E_BADPARM;            // not on the save list
        }
    }
    return B200_OK;
}

int b200_batch_set_saves(b200_batch *b, int nsave, const int *unknown)
{
    if (!b || nsave <= 0 || !unknown) return B200_E_BADPARM;
    if (b->V.out) return B200_E_BADPARM;                 // before the first run: the buffer is sized by it
    CK(cudaSetDevice(b->device));
    std::vector<int> slot(b->an.n, -1), dev(b->an.n + 1, -1);
    for (int k = 0; k < nsave; k++) {
        if (unknown[k] < 0 || unknown[k] >= b->an.n || slot[unknown[k]] >= 0) return B200_E_BADPARM;
        slot[unknown[k]] = k;
        dev[unknown[k] + 1] = k;
    }
    b->save_slot = slot;
    b->V.outslot = b->dupload(dev);
    b->K.nout = nsave;
    return B200_OK;
}

int b200_batch_num_saved(const b200_batch *b) { return b ? b->K.nout : 0; }

static int get_waves(const b200_batch *b, int nsel, const int *sel, double *times, double *waves)
{
    if (!b || !b->ran) return B200_E_BADPARM;
    CK(cudaSetDevice(b->device));
    const Analysis &an = b->an;
    int np = an.npoints_max;
    if (times) {
        // user time points: t_check sequence, tranan.cc:472-479
        double tc = an.t_start;
        for (int p = 0; p < np; p++, tc += an.t_step) times[p] = tc;
    }
    if (!waves) return B200_OK;
    int nn = sel ? nsel : b->K.nout;
    size_t total = (size_t)b->ntrials*np*nn;
    std::vector<int> rows;
    if (sel) { int rc = out_rows(b, nsel, sel, rows); if (rc) return rc; sel = rows.data(); }
    if (total > b->wout_cap) {
        if (b->d_wout) cudaFree(b->d_wout);
        b->d_wout = nullptr; b->wout_cap = 0;
        CK(cudaMalloc(&b->d_wout, total*sizeof(double)));
        b->wout_cap = total;
    }
    if (sel) {
        if (nsel > b->wsel_cap) {
            if (b->d_wsel) cudaFree(b->d_wsel);
            b->d_wsel = nullptr; b->wsel_cap = 0;
            CK(cudaMalloc(&b->d_wsel, nsel*sizeof(int)));
            b->wsel_cap = nsel;
        }
        CK(cudaMemcpyAsync(b->d_wsel, sel, nsel*sizeof(int), cudaMemcpyHostToDevice, b->stream));
    }
    launch_gather_waves(b->stream, b->K.nout, b->ldt, b->ntrials, np, b->V.out, b->d_npts, nsel,
        sel ? b->d_wsel : nullptr, b->d_wout);
    CK(cudaMemcpyAsync(waves, b->d_wout, total*sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return B200_OK;
}

// results straight into DEVICE buffers of the caller (the multi-GPU gather hands them to NCCL
// without a host round trip): dwaves[ntrials][npoints][nsel], dcounters[ntrials][B200_NCOUNTERS]
int b200_batch_get_results_dev(const b200_batch *b, int nsel, const int *sel, double *dwaves,
    int *dcounters)
{
    if (!b || !b->ran || !sel || nsel <= 0 || !dwaves) return B200_E_BADPARM;
    CK(cudaSetDevice(b->device));
    const Analysis &an = b->an;
    std::vector<int> rows;
    { int rc = out_rows(b, nsel, sel, rows); if (rc) return rc; sel = rows.data(); }
    if (nsel > b->wsel_cap) {
        if (b->d_wsel) cudaFree(b->d_wsel);
        b->d_wsel = nullptr; b->wsel_cap = 0;
        CK(cudaMalloc(&b->d_wsel, nsel*sizeof(int)));
        b->wsel_cap = nsel;
    }
    CK(cudaMemcpyAsync(b->d_wsel, sel, nsel*sizeof(int), cudaMemcpyHostToDevice, b->stream));
    CK(cudaStreamSynchronize(b->stream));               // `rows` is a local
    launch_gather_waves(b->stream, b->K.nout, b->ldt, b->ntrials, an.npoints_max, b->V.out, b->d_npts, nsel,
        b->d_wsel, dwaves);
    if (dcounters)
        CK(cudaMemcpyAsync(dcounters, b->d_counters, (size_t)b->ntrials*B200_NCOUNTERS*sizeof(int),
            cudaMemcpyDeviceToDevice, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return B200_OK;
}

int b200_batch_get_waves(const b200_batch *b, double *times, double *waves)
{
    return get_waves(b, 0, nullptr, times, waves);
}

int b200_batch_get_waves_sel(const b200_batch *b, int nsel, const int *sel, double *waves)
{
    if (!sel || nsel <= 0) return B200_E_BADPARM;
    return get_waves(b, nsel, sel, nullptr, waves);
}

int b200_batch_get_stats(const b200_batch *b, b200_run_stats *st)
{
    if (!b || !st) return B200_E_BADPARM;
    *st = b->stats;
    // algorithmic bytes per launch: trial-varying FP64 data that must move
    const Analysis &an = b->an;
    double per_mos = 8.0*(4 /*node V*/ + 17 /*state0 w*/ + 12 /*state1 r*/ + 2*MC_NCACHE);
    int live = 0;
    for (int s : an.mos.cslot) if (s >= 0) live++;
    for (int s : an.mos.rslot) if (s >= 0) live++;
    double mosbytes = per_mos*an.mos.n + 8.0*live;
    int nvar = 0;
    for (int k = 0; k < MP_NPARAMS; k++) nvar += b->prm_ts[k];
    mosbytes += 8.0*nvar*an.mos.n;
    double capbytes = 8.0*(2 + 2 + 1 + 2)*an.cap.n;
    for (int s : an.cap.cslot) if (s >= 0) capbytes += 8.0;
    for (int s : an.cap.rslot) if (s >= 0) capbytes += 8.0;
    capbytes += 8.0*(1 + 3 + 3 + 2 + 2)*an.ind.n;   // branch current, req/veq/prevFlux, 2 states r/w, 2 stamps
    double srcbytes = 0;
    for (int s : an.src.rslot) if (s >= 0) srcbytes += 8.0;
    // JJ: 2 node V, 5 state-1/2 reads, 7 state-0 writes, delVdelT r/w, live stamps,
    // per-trial derived parameters where ics varies
    double jjbytes = 8.0*(2 + 5 + 7 + 2)*an.jj.n + (b->jjics_ts ? 8.0*6*an.jj.n : 0.0);
    for (int sl : an.jj.cslot) if (sl >= 0) jjbytes += 8.0;
    for (int sl : an.jj.rslot) if (sl >= 0) jjbytes += 8.0;
    // TJM: the RSJ figure plus its recurrences: 4 x nterms complex read + 2 x nterms written per
    // iteration (the step coefficients are written once per step)
    jjbytes += 8.0*(2 + 4 + 6 + 2 + 2)*an.tjm.n + 8.0*(2*(4 + 3) + 2*2)*an.tjm.nterms*an.tjm.n;
    for (int sl : an.tjm.cslot) if (sl >= 0) jjbytes += 8.0;
    for (int sl : an.tjm.rslot) if (sl >= 0) jjbytes += 8.0;
    st->algbytes_load_per_launch = (mosbytes + capbytes + srcbytes + jjbytes)*b->ntrials;
    st->algbytes_solve_per_launch = 8.0*((double)an.NC + 2.0*an.n)*b->ntrials;
    return B200_OK;
}

int b200_batch_pattern(const b200_batch *b, int *n, int *nnz, const int **Ap, const int **Ai)
{
    if (!b) return B200_E_BADPARM;
    *n = b->an.n; *nnz = b->an.nnz; *Ap = b->an.Ap.data(); *Ai = b->an.Ai.data();
    return B200_OK;
}

// the pivot order of the transient modes when the batch keeps two (else the one order)
int b200_batch_ordering_tran(const b200_batch *b, const int **rowperm, const int **colperm)
{
    if (!b) return B200_E_BADPARM;
    const Analysis &an = b->an;
    *rowperm = an.has_tranp ? an.tranp.rowperm.data() : an.rowperm.data();
    *colperm = an.has_tranp ? an.tranp.colperm.data() : an.colperm.data();
    return B200_OK;
}

int b200_batch_ordering(const b200_batch *b, const int **rowperm, const int **colperm)
{
    if (!b) return B200_E_BADPARM;
    *rowperm = b->an.rowperm.data(); *colperm = b->an.colperm.data();
    return B200_OK;
}

} // extern "C"

// ---- accessors for klu_seam.cu (debug load hook) ----
int b200_internal_prepare(b200_batch *b)
{
    if (cudaSetDevice(b->device) != cudaSuccess) return B200_E_NOGPU;
    int rc = prepare_params(b);
    if (rc) return rc;
    return reset_state(b);
}

const Analysis *b200_internal_analysis(b200_batch *b) { return &b->an; }
int b200_internal_views(b200_batch *b, Consts **K, HotView **hot, VecView **V, MosView **M,
    CapView **Cp, SrcView **S, SolveView **Sv, cudaStream_t *st, int **cur)
{
    *K = &b->K; *hot = &b->hot; *V = &b->V; *M = &b->M; *Cp = &b->Cp; *S = &b->S; *Sv = &b->Sv;
    *st = b->stream; *cur = &b->cur;
    return 0;
}
