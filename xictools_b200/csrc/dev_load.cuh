// dev_load.cuh -- device-load arithmetic, one thread per (instance, trial).
//
// Restates, in FP64 with the reference's expression order (compile with -fmad=false):
//   MOSdev::load + mosstuff::{limiting,bypass,iv,cap,integ,load,load_dc}, eq1, qmeyer
//       wrspice/devlib/mos/mosload.cc:59-1041, moseq.cc:76-141
//   MOSdev::temperature                     devlib/mos/mostemp.cc:65-389
//   MOSdev::convTest                        devlib/mos/mosconv.cc:52-141
//   sDevLib::{pnjlim,fetlim,limvds}, pred   devlib/devlib.cc:475-620, include/device.h:171-175
//   sCKT::{integrate,find_ceq,NIceq,NItrapCheck}  include/circuit.h:1995-2029, ckt/niinteg.cc
//   CAPdev::load                            devlib/cap/capload.cc:75-205
//   SRCdev::load + IF{pulse,pwl,sin}Data::eval_func  devlib/src/srcload.cc:67-262,
//                                           src/inp/inptran.cc:1322-1414,2350-2388,2602-2624
// Stamps are not added into the matrix here: each one is written to its own contribution
// slot (contrib[slot][trial]); the solve kernel sums them per matrix entry in the
// reference's load order (atomics-free, deterministic).
#pragma once

#include <math.h>
#include "device_state.h"

#define REFTEMP     298.15
#define CHARGE      1.602176634e-19
#define CONSTboltz  1.380649e-23
#define CONSTKoverQ (CONSTboltz/CHARGE)
#define CONSTroot2  1.41421356237309504880
#define EPSSIL      (11.7 * 8.854214871e-12)
#define CM3PM3 1e6
#define CM2PM2 1e4
#define M2PCM2 1e-4
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define DMAX(a,b) ((a) > (b) ? (a) : (b))
#define DMIN(a,b) ((a) < (b) ? (a) : (b))

// Per-thread integration context.
struct TrialCtx {
    int mode, order, method, trapCheck;
    double delta, ag[2], pred[2];   // ag[0], ag[1]: all the trapezoidal rule needs
    const double *agp;      // &hot.ag[trial]: Gear coefficients k >= 2 at agp[k*ldt]
    double trapratio;
    double *states;         // &states[0][0][trial]; history plane k of this trial is ring slot
    int base, NS;           // (base + k) mod NS:  S(k,slot) = states[(phys*numStates + slot)*ldt]
    int numStates;
    int ldt;
    int *trapbad;           // per-trial flag
    const double *dOldp;    // &hot.deltaOld[trial]: deltaOld[k] at dOldp[k*ldt]
    const double *rcp;      // &hot.lte_rcp[trial]: LR_* rows (terr divisors, reciprocals, reltol/delta)
    double *lte;            // &hot.lte[trial]: LTE step bounds (this order, order + 1), min-reduced
};

__device__ __forceinline__ double &ST(const TrialCtx &c, int k, int slot)
{
    int phys = c.base + k;
    if (phys >= c.NS) phys -= c.NS;
    return c.states[((size_t)phys*c.numStates + slot)*c.ldt];
}
#define AGK(c, k) ((c).agp[(size_t)(k)*(c).ldt])

// Gear history sum (niinteg.cc:84-106), out of line with scalar arguments: the trapezoidal
// rule never runs it and the load kernels are instruction-fetch bound.
static __device__ __noinline__ double dev_NIceq_gear(const double *states, int base, int NS, int numStates,
    int ldt, const double *agp, double ag1, int order, int q)
{
    double cq = 0;
    const int kmax = (order == 1) ? 1 : (order >= 3 && order <= 6) ? order : 2;
    for (int k = kmax; k >= 2; k--) {
        int phys = base + k;
        if (phys >= NS) phys -= NS;
        cq += agp[(size_t)k*ldt]*states[((size_t)phys*numStates + q)*ldt];
    }
    int phys = base + 1;
    if (phys >= NS) phys -= NS;
    cq += ag1*states[((size_t)phys*numStates + q)*ldt];
    return cq;
}

__device__ __forceinline__ double dev_NIceq(const TrialCtx &c, int q)
{
    if (c.method == 1)
        return dev_NIceq_gear(c.states, c.base, c.NS, c.numStates, c.ldt, c.agp, c.ag[1], c.order, q);
    if (c.order == 1)
        return c.ag[1]*ST(c, 1, q);
    return - ST(c, 1, q+1)*c.ag[1] - c.ag[0]*ST(c, 1, q);
}

__device__ inline double dev_find_ceq(const TrialCtx &c, int q)
{
    if (c.method == 0)
        return (c.order <= 1 ? ST(c, 1, q)*c.ag[1] :
            - ST(c, 1, q+1)*c.ag[1] - ST(c, 1, q)*c.ag[0]);
    return dev_NIceq(c, q);
}

__device__ inline void dev_trapcheck(const TrialCtx &c, int q)
{
    // niinteg.cc:117-152
    if (!c.trapCheck || *c.trapbad) return;
    double c0 = ST(c, 0, q+1);
    if (fabs(c0) < 1e-6) return;
    double q0 = ST(c, 0, q);
    if (fabs(q0) < 1e-16) return;
    double c1 = ST(c, 1, q+1), c2 = ST(c, 2, q+1), q1 = ST(c, 1, q);
    double dqdt = fabs(q0 - q1)/c.delta;
    if (fabs(c1) > c.trapratio*dqdt && fabs(c0 - c1) > c.trapratio*fabs(c0 - c2))
        *c.trapbad = 1;
}

__device__ inline void dev_integrate2(const TrialCtx &c, int q, double cq)
{
    ST(c, 0, q+1) = c.ag[0]*ST(c, 0, q) + cq;
    dev_trapcheck(c, q);
}

// sCKT::terr (ckt.cc:2326-2389) for one integrated charge, evaluated inside the load of a
// transient Newton iteration: if this iteration turns out to be the converged one, the value is
// what sCKT::trunc would compute from the states this load leaves behind.  q[0..3]/c[0..1] are
// the charge / current history already at hand; deeper history (Gear > 2) is read from HBM.
__device__ __forceinline__ void atomic_min_pos(double *addr, double v)
{
    // positive finite doubles order like their bit patterns
    atomicMin((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}

// a/b from a correctly rounded reciprocal y = RN(1/b) (Markstein): two FMA refinement steps leave a
// faithful quotient, the third returns RN(a/b).  Bit-identical to the IEEE division the reference
// executes for the finite, normal operands of the divided-difference table (checked against
// true division on the device by tests/test_engine_gpu.py); 5 instructions instead of ~35.
__device__ __forceinline__ double dev_div_rcp(double a, double b, double y)
{
    double q = a*y;
    double r = __fma_rn(-b, q, a);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-b, q, a);
    return __fma_rn(r, y, q);
}

// The divisors of sCKT::terr's divided-difference table depend only on the step history of the
// trial: the control kernel publishes them with their reciprocals (LR_* rows of hot.lte_rcp).
enum { LR_d0, LR_d1, LR_d2, LR_e0, LR_e1, LR_f0, LR_NDIV,       // d_i = deltaOld[i], e0 = d1+d0, e1 = d2+d1, f0 = e1+d0
       LR_CT = 2*LR_NDIV, LR_NROWS };                            // reltol/delta

// Orders 1 and 2 (everything the trapezoidal rule uses).  rcp = &hot.lte_rcp[trial] (LR_* rows,
// stride ldt).  The load kernels are instruction-fetch bound: callers keep ONE inlined copy of this
// by looping (unroll 1) over charges and over the two orders instead of repeating the call.
__device__ __forceinline__ double dev_terr12(int order, double q0, double q1, double q2, double q3,
    double tol, const double *rcp, int ldt, int method, double trtol, double abstol)
{
    const size_t L = (size_t)ldt;
    double d0 = dev_div_rcp(q0 - q1, rcp[LR_d0*L], rcp[(LR_NDIV + LR_d0)*L]);
    double d1 = dev_div_rcp(q1 - q2, rcp[LR_d1*L], rcp[(LR_NDIV + LR_d1)*L]);
    d0 = dev_div_rcp(d0 - d1, rcp[LR_e0*L], rcp[(LR_NDIV + LR_e0)*L]);
    if (order != 1) {
        double d2 = dev_div_rcp(q2 - q3, rcp[LR_d2*L], rcp[(LR_NDIV + LR_d2)*L]);
        d1 = dev_div_rcp(d1 - d2, rcp[LR_e1*L], rcp[(LR_NDIV + LR_e1)*L]);
        d0 = dev_div_rcp(d0 - d1, rcp[LR_f0*L], rcp[(LR_NDIV + LR_f0)*L]);
    }
    const double coef = (method == 0) ? (order == 1 ? .5 : .08333333333) : (order == 1 ? .5 : .2222222222);
    double del = fabs(d0*coef);
    return trtol*tol/DMAX(abstol, del);
}

// Gear orders 3..6: one generic routine (local arrays; not on the trapezoidal hot path)
static __device__ __noinline__ double dev_terr_generic(const double *states, int base, int NS, int numStates,
    int ldt, const double *dOldp, int method, double trtol, double abstol, int order,
    double q0, double q1, double q2, double q3, double tol, int qslot)
{
    const double q[4] = { q0, q1, q2, q3 };
    const double gearCoeff[] = { .5, .2222222222, .1363636364, .096, .07299270073, .05830903790 };
    const double trapCoeff[] = { .5, .08333333333 };
    double diff[8], deltmp[8], dOld[8];
    for (int i = 0; i <= order + 1; i++) {
        if (i < 4) { diff[i] = q[i]; continue; }
        int phys = base + i;                            // ST(c, i, qslot)
        if (phys >= NS) phys -= NS;
        diff[i] = states[((size_t)phys*numStates + qslot)*ldt];
    }
    for (int i = 0; i <= order; i++) { dOld[i] = dOldp[(size_t)i*ldt]; deltmp[i] = dOld[i]; }
    for (int j = order; j >= 0; j--) {
        for (int i = 0; i <= j; i++) diff[i] = (diff[i] - diff[i+1])/deltmp[i];
        if (j > 0)
            for (int i = 0; i <= j - 1; i++) deltmp[i] = deltmp[i+1] + dOld[i];
    }
    double del;
    if (method == 0) del = fabs(diff[0]*trapCoeff[order <= 2 ? order - 1 : 1]);
    else del = fabs(diff[0]*gearCoeff[order - 1]);
    return trtol*tol/DMAX(abstol, del);
}

// out-of-line copy for the small kernels (capacitor, inductor), where a call is cheaper than the
// rolled loop the MOSFET kernel uses
static __device__ __noinline__ double dev_terr12_ool(int order, double q0, double q1, double q2, double q3,
    double tol, const double *rcp, int ldt, int method, double trtol, double abstol)
{
    return dev_terr12(order, q0, q1, q2, q3, tol, rcp, ldt, method, trtol, abstol);
}

template <bool OOL>
__device__ __forceinline__ double dev_terr_order(const TrialCtx &c, const b200_options &opt, int order,
    const double *q, double tol, int qslot, const double *rcp, int rld)
{
    if (order <= 2) {
        if (OOL) return dev_terr12_ool(order, q[0], q[1], q[2], q[3], tol, rcp, rld, c.method, opt.trtol, opt.abstol);
        return dev_terr12(order, q[0], q[1], q[2], q[3], tol, rcp, rld, c.method, opt.trtol, opt.abstol);
    }
    return dev_terr_generic(c.states, c.base, c.NS, c.numStates, c.ldt, c.dOldp, c.method, opt.trtol, opt.abstol,
        order, q[0], q[1], q[2], q[3], tol, qslot);
}

// OOL false: one inlined copy of the table code, looped over the two orders (MOSFET kernel, which
// is instruction-fetch bound); OOL true: two calls of the out-of-line copy.
template <bool OOL>
__device__ __forceinline__ void dev_lte_charge(const TrialCtx &c, const b200_options &opt, const double *q,
    double c0, double c1, int qslot, double &m0, double &m1, const double *rcp, int rld)
{
    // rcp/rld: the trial's LR_* rows and their stride (hot.lte_rcp: stride ldt; a register copy: 1)
    double y0 = fabs(c0), y1 = fabs(c1);
    double volttol = opt.abstol + opt.reltol*DMAX(y0, y1);
    y0 = fabs(q[0]); y1 = fabs(q[1]);
    double chargetol = DMAX(y0, y1);
    if (chargetol < opt.chgtol) chargetol = opt.chgtol;
    chargetol *= rcp[(size_t)LR_CT*rld];              // reltol/delta, published per trial
    double tol = DMAX(volttol, chargetol);
    if (OOL) {
        double d0 = dev_terr_order<true>(c, opt, c.order, q, tol, qslot, rcp, rld);
        if (d0 < m0) m0 = d0;
        if (c.order + 1 <= opt.maxord) {
            double d1 = dev_terr_order<true>(c, opt, c.order + 1, q, tol, qslot, rcp, rld);
            if (d1 < m1) m1 = d1;
        }
    }
    else {
#pragma unroll 1
        for (int o = 0; o < 2; o++) {
            const int ord = c.order + o;
            if (o && ord > opt.maxord) break;
            double d = dev_terr_order<false>(c, opt, ord, q, tol, qslot, rcp, rld);
            if (o == 0) { if (d < m0) m0 = d; }
            else { if (d < m1) m1 = d; }
        }
    }
}

// ---- limiters: devlib/devlib.cc:475-620 ----
__device__ __forceinline__ double dev_limvds(double vnew, double vold)
{
    if (vold >= 3.5) {
        if (vnew > vold)
            vnew = DMIN(vnew, 3*vold + 2);
        else {
            if (vnew < 3.5)
                vnew = DMAX(vnew, 2);
        }
    }
    else {
        if (vnew > vold)
            vnew = DMIN(vnew, 4);
        else
            vnew = DMAX(vnew, -.5);
    }
    return vnew;
}

__device__ __forceinline__ double dev_pnjlim(double vnew, double vold, double vt, double rvt, double vcrit, int *icheck)
{
    if (vnew > vcrit && fabs(vnew - vold) > (vt + vt)) {
        if (vold > 0) {
            double arg = dev_div_rcp(vnew - vold, vt, rvt);
            if (arg > 0)
                vnew = vold + vt*(2.0 + log(arg - 1.8));
            else
                vnew = vold - vt*(2.0 + log(-arg + 2.0));
        }
        else
            vnew = vt*log(dev_div_rcp(vnew, vt, rvt));
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

__device__ __forceinline__ double dev_fetlim(double vnew, double vold, double vto)
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
                    vnew = DMAX(vnew, vto + 2);
            }
            else {
                if (delv >= vtsthi)
                    vnew = vold + vtsthi;
            }
        }
        else {
            if (delv <= 0)
                vnew = DMAX(vnew, vto - 0.5);
            else
                vnew = DMIN(vnew, vto + 4.0);
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

// ---------------------------------------------------------------------------------------
// MOS
// ---------------------------------------------------------------------------------------
// One (instance, trial) of the MOS load.  Everything that is per trial is staged by the kernel
// into shared memory with one burst of asynchronous copies (row r of this thread at
// col[r*MOS_TPB]), so the loads are all in flight at once and nothing has to be parked in
// registers while the model is evaluated.  Derived parameters are either per-trial rows of the
// same staging area or values shared by all trials (MosStageMap tells which).
struct MosInst {
    int type, off, st, mode, level;
    int g, b, dp, sp;
    double *col;            // staging column of this thread
    const double *sg2;      // this device's rows of state plane 2 (row k at sg2[k*ldt])
    const double *shp;      // staging area base (shared parameter block addressed from here)
    const MosStageMap *map;
    unsigned tid8;          // 8*thread: byte offset of this thread's staging column
    __device__ __forceinline__ double P(int k) const
        { return *(const double*)((const char*)shp + (map->aoff[k] + (int)(tid8 & map->amask[k]))); }
    __device__ __forceinline__ double &R(int row) const { return col[row*MOS_TPB]; }
    __device__ __forceinline__ double eta() const { return P(MP_eta); }
    __device__ __forceinline__ double theta() const { return P(MP_theta); }
    __device__ __forceinline__ double kappa() const { return P(MP_kappa); }
    __device__ __forceinline__ double vmax() const { return P(MP_vmax); }
    __device__ __forceinline__ double xj() const { return P(MP_xj); }
    __device__ __forceinline__ double ld() const { return P(MP_ld); }
    __device__ __forceinline__ double nfs() const { return P(MP_nfs); }
    __device__ __forceinline__ double xd() const { return P(MP_xd); }
    __device__ __forceinline__ double alpha() const { return P(MP_alpha); }
    __device__ __forceinline__ double oxideCapFactor() const { return P(MP_oxideCapFactor); }
    __device__ __forceinline__ double tSurfMob() const { return P(MP_tSurfMob); }
    __device__ __forceinline__ double xl() const { return P(MP_xl); }
    __device__ __forceinline__ double w() const { return P(MP_w); }
    __device__ __forceinline__ double uexp() const { return P(MP_uexp); }
    __device__ __forceinline__ double ucrit() const { return P(MP_ucrit); }
    __device__ __forceinline__ double neff() const { return P(MP_neff); }
    __device__ __forceinline__ double narrow() const { return P(MP_narrow); }
    __device__ __forceinline__ double nsub() const { return P(MP_nsub); }
    __device__ __forceinline__ double surfMob() const { return P(MP_surfMob); }
    __device__ __forceinline__ double rvt() const { return P(MP_rvt); }
    __device__ __forceinline__ double rBulkPot() const { return P(MP_rBulkPot); }
    __device__ __forceinline__ double r1mj() const { return P(MP_r1mj); }
    __device__ __forceinline__ double r1mjsw() const { return P(MP_r1mjsw); }
    __device__ __forceinline__ double r2phi() const { return P(MP_r2phi); }
    __device__ __forceinline__ double r15phi() const { return P(MP_r15phi); }
    __device__ __forceinline__ double kv() const { return P(MP_kv); }
    __device__ __forceinline__ double nv() const { return P(MP_nv); }
    __device__ __forceinline__ double nc() const { return P(MP_nc); }
    __device__ __forceinline__ double lamda0() const { return P(MP_lamda0); }
    __device__ __forceinline__ double lamda1() const { return P(MP_lamda1); }
    __device__ __forceinline__ double sigma() const { return P(MP_sigma); }
    __device__ __forceinline__ double gamma1() const { return P(MP_gamma1); }
    __device__ __forceinline__ double &s1(int k) const { return col[(SR_S1 + k)*MOS_TPB]; }
    __device__ __forceinline__ double &s0(int k) const { return col[(SR_S0 + k)*MOS_TPB]; }
    __device__ __forceinline__ double tPhi() const { return P(MP_tPhi); }
    __device__ __forceinline__ double tVbi() const { return P(MP_tVbi); }
    __device__ __forceinline__ double tVto() const { return P(MP_tVto); }
    __device__ __forceinline__ double gamma() const { return P(MP_gamma); }
    __device__ __forceinline__ double lambda() const { return P(MP_lambda); }
    __device__ __forceinline__ double beta() const { return P(MP_beta); }
    __device__ __forceinline__ double oxideCap() const { return P(MP_oxideCap); }
    __device__ __forceinline__ double tSourceSatCur() const { return P(MP_tSourceSatCur); }
    __device__ __forceinline__ double tDrainSatCur() const { return P(MP_tDrainSatCur); }
    __device__ __forceinline__ double sourceVcrit() const { return P(MP_sourceVcrit); }
    __device__ __forceinline__ double drainVcrit() const { return P(MP_drainVcrit); }
    __device__ __forceinline__ double Cbd() const { return P(MP_Cbd); }
    __device__ __forceinline__ double Cbdsw() const { return P(MP_Cbdsw); }
    __device__ __forceinline__ double Cbs() const { return P(MP_Cbs); }
    __device__ __forceinline__ double Cbssw() const { return P(MP_Cbssw); }
    __device__ __forceinline__ double tBulkPot() const { return P(MP_tBulkPot); }
    __device__ __forceinline__ double tDepCap() const { return P(MP_tDepCap); }
    __device__ __forceinline__ double f2d() const { return P(MP_f2d); }
    __device__ __forceinline__ double f3d() const { return P(MP_f3d); }
    __device__ __forceinline__ double f4d() const { return P(MP_f4d); }
    __device__ __forceinline__ double f2s() const { return P(MP_f2s); }
    __device__ __forceinline__ double f3s() const { return P(MP_f3s); }
    __device__ __forceinline__ double f4s() const { return P(MP_f4s); }
    __device__ __forceinline__ double gsOvl() const { return P(MP_gsOvl); }
    __device__ __forceinline__ double gdOvl() const { return P(MP_gdOvl); }
    __device__ __forceinline__ double gbOvl() const { return P(MP_gbOvl); }
    __device__ __forceinline__ double dcon() const { return P(MP_dcon); }
    __device__ __forceinline__ double scon() const { return P(MP_scon); }
    __device__ __forceinline__ double mult() const { return P(MP_mult); }
    __device__ __forceinline__ double vt() const { return P(MP_vt); }
    __device__ __forceinline__ double mj() const { return P(MP_mj); }
    __device__ __forceinline__ double mjsw() const { return P(MP_mjsw); }
    __device__ __forceinline__ double icVDS() const { return P(MP_icVDS); }
    __device__ __forceinline__ double icVGS() const { return P(MP_icVGS); }
    __device__ __forceinline__ double icVBS() const { return P(MP_icVBS); }
    __device__ __forceinline__ double &von() const { return col[(SR_CACHE + MC_von)*MOS_TPB]; }
    __device__ __forceinline__ double &vdsat() const { return col[(SR_CACHE + MC_vdsat)*MOS_TPB]; }
    __device__ __forceinline__ double &cd() const { return col[(SR_CACHE + MC_cd)*MOS_TPB]; }
    __device__ __forceinline__ double &cbs() const { return col[(SR_CACHE + MC_cbs)*MOS_TPB]; }
    __device__ __forceinline__ double &cbd() const { return col[(SR_CACHE + MC_cbd)*MOS_TPB]; }
    __device__ __forceinline__ double &gm() const { return col[(SR_CACHE + MC_gm)*MOS_TPB]; }
    __device__ __forceinline__ double &gds() const { return col[(SR_CACHE + MC_gds)*MOS_TPB]; }
    __device__ __forceinline__ double &gmbs() const { return col[(SR_CACHE + MC_gmbs)*MOS_TPB]; }
    __device__ __forceinline__ double &gbd() const { return col[(SR_CACHE + MC_gbd)*MOS_TPB]; }
    __device__ __forceinline__ double &gbs() const { return col[(SR_CACHE + MC_gbs)*MOS_TPB]; }
    __device__ __forceinline__ double &capbd() const { return col[(SR_CACHE + MC_capbd)*MOS_TPB]; }
    __device__ __forceinline__ double &capbs() const { return col[(SR_CACHE + MC_capbs)*MOS_TPB]; }
    __device__ __forceinline__ double &o_vbs() const { return col[(SR_S0 + MS_vbs)*MOS_TPB]; }
    __device__ __forceinline__ double &o_vgs() const { return col[(SR_S0 + MS_vgs)*MOS_TPB]; }
    __device__ __forceinline__ double &o_vds() const { return col[(SR_S0 + MS_vds)*MOS_TPB]; }
    __device__ __forceinline__ double &o_vbd() const { return col[(SR_S0 + MS_vbd)*MOS_TPB]; }
    __device__ __forceinline__ double &o_capgs() const { return col[(SR_S0 + MS_capgs)*MOS_TPB]; }
    __device__ __forceinline__ double &o_capgd() const { return col[(SR_S0 + MS_capgd)*MOS_TPB]; }
    __device__ __forceinline__ double &o_capgb() const { return col[(SR_S0 + MS_capgb)*MOS_TPB]; }
};

struct MosStuff {
    double vt, vbs, vgs, vds, vbd, vgd, vgb, von, vdsat, cdrain;
    double gcgs, ceqgs, gcgd, ceqgd, gcgb, ceqgb, capgs, capgd, capgb;
};

struct MosOut {             // stamps go straight to this trial's column of contrib[NC][ldt]
    double *cbase;          // &contrib[0][trial]
    const int *cs, *rs;     // contribution slots of this instance (-1: ground row/column)
    unsigned ldt;
    __device__ __forceinline__ void M(int k, double v) const { int q = cs[k]; if (q >= 0) cbase[(size_t)(unsigned)q*ldt] = v; }
    __device__ __forceinline__ void Rw(int k, double v) const { int q = rs[k]; if (q >= 0) cbase[(size_t)(unsigned)q*ldt] = v; }
};

__device__ __forceinline__ void dev_sargs(double arg, double bot, double side, double *sarg, double *sargsw)
{
    // devlib.cc SARGS: the grading exponents .5 take the sqrt form; one copy of the code, the
    // sidewall trip skipped when both exponents agree
    double r = 0;
#pragma unroll 1
    for (int j = 0; j < 2; j++) {
        const double e = j ? side : bot;
        if (j && side == bot) break;
        if (e == .5) r = 1/sqrt(arg);
        else r = exp(-e*log(arg));
        if (j == 0) *sarg = r;
    }
    *sargsw = r;
}

__device__ inline void dev_qmeyer(double vgs, double vgd, double von, double vdsat,
    double *capgs, double *capgd, double *capgb, double phi, double cox, double r2phi, double r15phi)
{
    double vgst = vgs - von;
    if (vgst <= -phi) {
        *capgb = cox/2; *capgs = 0; *capgd = 0;
    }
    else if (vgst <= -phi/2) {
        *capgb = dev_div_rcp(-vgst*cox, 2*phi, r2phi); *capgs = 0; *capgd = 0;
    }
    else if (vgst <= 0) {
        *capgb = dev_div_rcp(-vgst*cox, 2*phi, r2phi);
        *capgs = dev_div_rcp(vgst*cox, 1.5*phi, r15phi) + dev_div_rcp(cox, 3.0, 1.0/3.0);
        *capgd = 0;
    }
    else {
        double vds = vgs - vgd;
        if (vdsat <= vds) {
            *capgs = dev_div_rcp(cox, 3.0, 1.0/3.0); *capgd = 0; *capgb = 0;
        }
        else {
            double vddif = 2.0*vdsat - vds;
            double vddif1 = vdsat - vds;
            double vddif2 = vddif*vddif;
            double rv = 1.0/vddif2;              // two divisions by vddif2 share its reciprocal
            *capgd = dev_div_rcp(cox*(1.0 - dev_div_rcp(vdsat*vdsat, vddif2, rv)), 3.0, 1.0/3.0);
            *capgs = dev_div_rcp(cox*(1.0 - dev_div_rcp(vddif1*vddif1, vddif2, rv)), 3.0, 1.0/3.0);
            *capgb = 0;
        }
    }
}

__device__ inline double dev_mos_eq1(MosInst &m, MosStuff &ms)
{
    double vbx = ((m.mode == 1) ? ms.vbs : ms.vbd);
    double vgx = ((m.mode == 1) ? ms.vgs : ms.vgd);
    double sarg;
    if (vbx <= 0)
        sarg = sqrt(m.tPhi() - vbx);
    else {
        sarg = sqrt(m.tPhi());
        sarg = sarg - vbx/(sarg+sarg);
        sarg = DMAX(0, sarg);
    }
    if (m.type > 0)
        ms.von = m.tVbi() + m.gamma()*sarg;
    else
        ms.von = -m.tVbi() + m.gamma()*sarg;
    double vgst = vgx - ms.von;
    ms.vdsat = DMAX(vgst, 0);
    double cdrain;
    if (vgst <= 0) {
        cdrain = 0; m.gm() = 0; m.gds() = 0; m.gmbs() = 0;
    }
    else {
        double vds = (m.mode > 0) ? ms.vds : -ms.vds;
        double arg;
        if (sarg <= 0)
            arg = 0;
        else
            arg = m.gamma()/(sarg+sarg);
        double betap = m.beta()*(1 + m.lambda()*vds);
        if (vgst <= vds) {
            m.gm() = betap*vgst;
            vgst *= vgst*.5;
            cdrain = betap*vgst;
            m.gds() = m.lambda()*m.beta()*vgst;
            m.gmbs() = m.gm()*arg;
        }
        else {
            m.gm() = betap*vds;
            cdrain = m.gm()*(vgst - .5*vds);
            m.gds() = betap*(vgst - vds) + m.lambda()*m.beta()*vds*(vgst - .5*vds);
            m.gmbs() = m.gm()*arg;
        }
    }
    return cdrain;
}

#define P66 .66666666666667
#define P33 .33333333333333
#define P1X3 1.3333333333333

// moseq.cc:602-688: Baum's scattering-limited velocity saturation (quartic)
__device__ inline void dev_mos2_vsat(double *vsat, const double *baum)
{
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
#pragma unroll
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

// moseq.cc:147-597: level 2 (Grove-Frohman / Meyer analytical model)
__device__ inline double dev_mos_eq2(MosInst &m, MosStuff &ms)
{
    double lvbs, lvds, lvgs;
    if (m.mode > 0) { lvbs = ms.vbs; lvds = ms.vds; lvgs = ms.vgs; }
    else { lvbs = ms.vbd; lvds = -ms.vds; lvgs = ms.vgd; }
    const double xl = m.xl(), nsub = m.nsub(), tPhi = m.tPhi(), xd = m.xd(), xj = m.xj(), gamma = m.gamma();
    const double oxideCap = m.oxideCap(), vmax = m.vmax(), nfs = m.nfs();
    double xlamda = m.lambda();
    double phiMinVbs = tPhi - lvbs;
    double oneoverl = 1.0/xl;
    double sarg, dsrgdb, d2sdb2;
    double sphi = sqrt(tPhi);
    double sphi3 = tPhi*sphi;
    if (lvbs <= 0.0) {
        sarg = sqrt(phiMinVbs);
        dsrgdb = -0.5/sarg;
        d2sdb2 = 0.5*dsrgdb/phiMinVbs;
    }
    else {
        sarg = sphi/(1.0 + 0.5*lvbs/tPhi);
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
        barg = sphi/(1.0 + 0.5*(lvbs - lvds)/tPhi);
        double tmp = barg/sphi3;
        dbrgdb = -0.5*barg*tmp;
        d2bdb2 = -dbrgdb*tmp;
    }
    int fssandcap = (nfs != 0.0 && oxideCap != 0.0);
    double factor = 0.125*m.narrow()*2.0*M_PI*EPSSIL/oxideCap*xl;
    double eta = 1.0 + factor;
    double oneovere = 1.0/eta;
    double vbin = m.tVbi()*m.type + factor*phiMinVbs;
    double gamasd = gamma, gammad = gamma;
    double dgddvb = 0.0, dgdvds = 0.0, dgddb2 = 0.0;
    if ((gamma > 0.0) || (nsub > 0.0)) {
        if (xj > 0) {
            double tbxwd = xd*dbrgdb;
            double tbxws = xd*dsrgdb;
            double tmp = 2.0/xj;
            double targxs = 1.0 + tmp*xd*sarg;
            double targxd = 1.0 + tmp*xd*barg;
            double targs = sqrt(targxs);
            double targd = sqrt(targxd);
            tmp = .5*xj*oneoverl;
            gamasd *= (1.0 - tmp*(targs + targd - 2.0));
            targs = oneoverl/targs;
            targd = oneoverl/targd;
            dgddvb = -.5*gamma*(tbxws*targs + tbxwd*targd);
            dgdvds = gamma*0.5*tbxwd*targd;
            tmp = -xd*(d2sdb2 + dsrgdb*dsrgdb*xd/(xj*targxs))*targs;
            double tmp1 = -xd*(d2bdb2 + dbrgdb*dbrgdb*xd/(xj*targxd))*targd;
            dgddb2 = -0.5*gamma*(tmp + tmp1);
        }
    }
    ms.von = vbin + gamasd*sarg;
    double vth = ms.von;
    ms.vdsat = 0.0;
    double argg;
    double xn = 1.0;
    if (fssandcap) {
        double tmp = CHARGE*nfs*CM2PM2;
        double tmp1 = -(gamasd*dsrgdb + dgddvb*sarg) + factor;
        xn += tmp/oxideCap*m.w()*xl + tmp1;
        tmp = ms.vt*xn;
        ms.von += tmp;
        argg = 1.0/tmp;
    }
    else {
        argg = 0;
        if (lvgs <= vbin)
            m.gds() = 0.0;
    }
    double cdrain = 0;
    if (fssandcap || (lvgs > vbin)) {
        double vgst = lvgs - ms.von;
        double sarg3 = sarg*sarg*sarg;
        double sbiarg = sqrt(m.tBulkPot());
        gammad = gamasd;
        double dgdvbs = dgddvb;
        double body = barg*barg*barg - sarg3;
        double gdbdv = 2.0*gammad*(barg*barg*dbrgdb - sarg*sarg*dsrgdb);
        double dodvbs = -factor + dgdvbs*sarg + gammad*dsrgdb;
        double dxndvb, dxndvd, dodvds;
        if (fssandcap) {
            dxndvb = 2.0*dgdvbs*dsrgdb + gammad*d2sdb2 + dgddb2*sarg;
            dodvbs = dodvbs + ms.vt*dxndvb;
            dxndvd = dgdvds*dsrgdb;
            dodvds = dgdvds*sarg + ms.vt*dxndvd;
        }
        else {
            dxndvb = 0; dxndvd = 0; dodvds = 0;
        }
        double ufact, ueff, dudvgs, dudvds, dudvbs;
        if (oxideCap <= 0.0) {
            ufact = 1.0;
            ueff = m.surfMob()*M2PCM2;
            dudvgs = 0.0; dudvds = 0.0; dudvbs = 0.0;
        }
        else {
            double tmp = m.ucrit()*100*EPSSIL/m.oxideCapFactor();
            if (vgst <= tmp) {
                ufact = 1.0;
                ueff = m.surfMob()*M2PCM2;
                dudvgs = 0.0; dudvds = 0.0; dudvbs = 0.0;
            }
            else {
                ufact = exp(m.uexp()*log(tmp/vgst));
                ueff = m.surfMob()*M2PCM2*ufact;
                dudvgs = -ufact*m.uexp()/vgst;
                dudvds = 0.0;
                dudvbs = m.uexp()*ufact*dodvbs/vgst;
            }
        }
        gammad = gamasd*oneovere;
        dgdvbs = dgddvb;
        double vgsx = lvgs;
        if (fssandcap)
            vgsx = DMAX(lvgs, ms.von);
        double dsdvgs, dsdvbs;
        if (gammad > 0) {
            double tmp = gammad*gammad;
            double tmp1 = (vgsx - vbin)*oneovere + phiMinVbs;
            if (tmp1 <= 0.0) {
                ms.vdsat = 0.0;
                dsdvgs = 0.0;
                dsdvbs = 0.0;
            }
            else {
                double arg = sqrt(1.0 + 4.0*tmp1/tmp);
                ms.vdsat = (vgsx - vbin)*oneovere + .5*tmp*(1.0 - arg);
                ms.vdsat = DMAX(ms.vdsat, 0.0);
                dsdvgs = (1.0 - 1.0/arg)*oneovere;
                dsdvbs = (gammad*(1.0 - arg) + 2.0*tmp1/(gammad*arg))*oneovere*dgdvbs +
                    1.0/arg + factor*dsdvgs;
            }
        }
        else {
            ms.vdsat = (vgsx - vbin)*oneovere;
            ms.vdsat = DMAX(ms.vdsat, 0.0);
            dsdvgs = 1.0;
            dsdvbs = 0.0;
        }
        if (vmax > 0) {
            double tmp = vmax*xl/ueff;
            double tmp1 = (vgsx - vbin)*oneovere + phiMinVbs;
            double param[5];
            param[0] = P1X3*gammad;
            param[1] = -2.0*(tmp1 + tmp);
            param[2] = -2.0*gammad*tmp;
            param[3] = 2.0*tmp1*(phiMinVbs + tmp) - phiMinVbs*phiMinVbs - P1X3*gammad*sarg3;
            param[4] = phiMinVbs;
            dev_mos2_vsat(&ms.vdsat, param);
        }
        double dldvgs = 0.0, dldvds = 0.0, dldvbs = 0.0;
        double bsarg, bodys, gdbdvs, dbsrdb = 0;
        if (lvds == 0.0) {
            bsarg = 0; bodys = 0; gdbdvs = 0;
        }
        else {
            gammad = gamasd;
            if ((lvbs - ms.vdsat) <= 0) {
                bsarg = sqrt(ms.vdsat + phiMinVbs);
                dbsrdb = -0.5/bsarg;
            }
            else {
                bsarg = sphi/(1.0 + 0.5*(lvbs - ms.vdsat)/tPhi);
                dbsrdb = -0.5*bsarg*bsarg/sphi3;
            }
            bodys = bsarg*bsarg*bsarg - sarg3;
            gdbdvs = 2.0*gammad*(bsarg*bsarg*dbsrdb - sarg*sarg*dsrgdb);
            if (vmax <= 0) {
                if (nsub != 0.0 && xlamda <= 0.0) {
                    double tmp = .25*(lvds - ms.vdsat);
                    double tmp1 = sqrt(1.0 + tmp*tmp);
                    tmp = sqrt(tmp + tmp1);
                    xlamda = xd*tmp*oneoverl/lvds;
                    tmp = lvds*xlamda/(8.0*tmp1);
                    dldvgs = tmp*dsdvgs;
                    dldvds = -xlamda + tmp;
                    dldvbs = tmp*dsdvbs;
                }
            }
            else {
                double txdv = xd/sqrt(m.neff());
                double txlv = vmax*txdv/(2.0*ueff);
                double tqdsat = -1.0 + gammad*dbsrdb;
                double tmp = vmax*xl;
                double tmp1 = (vgsx - vbin)*oneovere - ms.vdsat - gammad*bsarg;
                tmp1 = 1.0/(tmp*tqdsat - ueff*tmp1);
                dsdvgs = -(tmp - ueff*ms.vdsat)*oneovere*tmp1;
                dsdvbs = -(-tmp*(1.0 + tqdsat - factor*oneovere) +
                    ueff*(gdbdvs - P66*dgdvbs*bodys)*oneovere)*tmp1;
                if (nsub != 0.0 && xlamda <= 0.0) {
                    tmp = lvds - ms.vdsat;
                    tmp = DMAX(tmp, 0.0);
                    tmp1 = sqrt(txlv*txlv + tmp);
                    tmp = txdv*oneoverl/(2.0*tmp1);
                    xlamda = txdv*(tmp1 - txlv)*oneoverl/lvds;
                    dldvgs = tmp*dsdvgs;
                    dldvds = -xlamda + tmp;
                    dldvbs = tmp*dsdvbs;
                }
            }
        }
        double xwb = xd*sbiarg;
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
        double beta1 = m.beta()*ufact*clfact;
        ufact = 1.0/ufact;
        gammad = gamasd;
        dgdvbs = dgddvb;
        if (lvds > 1.0e-10) {
            if (lvgs > ms.von) {
                if (lvds <= ms.vdsat) {
                    cdrain = beta1*((lvgs - vbin - .5*eta*lvds)*lvds - P66*gammad*body);
                    tmp = cdrain*(dudvgs*ufact - dldvgs*clfact);
                    m.gm() = tmp + beta1*lvds;
                    tmp = cdrain*(dudvds*ufact - dldvds*clfact);
                    m.gds() = tmp + beta1*(lvgs - vbin - eta*lvds - gammad*barg - P66*dgdvds*body);
                    tmp = cdrain*(dudvbs*ufact - dldvbs*clfact);
                    m.gmbs() = tmp - beta1*(gdbdv + P66*dgdvbs*body - factor*lvds);
                }
                else {
                    cdrain = beta1*((lvgs - vbin - eta*.5*ms.vdsat)*ms.vdsat - P66*gammad*bodys);
                    tmp = cdrain*(dudvgs*ufact - dldvgs*clfact);
                    m.gm() = tmp + beta1*ms.vdsat +
                        beta1*(lvgs - vbin - eta*ms.vdsat - gammad*bsarg)*dsdvgs;
                    m.gds() = -cdrain*dldvds*clfact - P66*beta1*dgdvds*bodys;
                    tmp = cdrain*(dudvbs*ufact - dldvbs*clfact);
                    m.gmbs() = tmp - beta1*(gdbdvs + P66*dgdvbs*bodys - factor*ms.vdsat) +
                        beta1*(lvgs - vbin - eta*ms.vdsat - gammad*bsarg)*dsdvbs;
                }
            }
            else {
                if (ms.vdsat <= 0) {
                    m.gds() = 0.0;
                    if (lvgs > vth)
                        return cdrain;
                    m.gm() = 0.0;
                    m.gmbs() = 0.0;
                    return 0.0;
                }
                double vdson = DMIN(ms.vdsat, lvds);
                if (lvds > ms.vdsat) {
                    barg = bsarg;
                    dbrgdb = dbsrdb;
                    body = bodys;
                    gdbdv = gdbdvs;
                }
                double cdson = beta1*((ms.von - vbin - .5*eta*vdson)*vdson - P66*gammad*body);
                double didvds = beta1*(ms.von - vbin - eta*vdson - gammad*barg);
                double gdson = -cdson*dldvds*clfact - P66*beta1*dgdvds*body;
                if (lvds < ms.vdsat)
                    gdson += didvds;
                double gbson = -cdson*dldvbs*clfact +
                    beta1*(dodvbs*vdson + factor*vdson - P66*dgdvbs*body - gdbdv);
                if (lvds > ms.vdsat)
                    gbson += didvds*dsdvbs;
                tmp = exp(argg*(lvgs - ms.von));
                cdrain = cdson*tmp;
                double gmw = cdrain*argg;
                m.gm() = gmw;
                if (lvds > ms.vdsat) m.gm() = gmw + didvds*dsdvgs*tmp;
                tmp1 = gmw*(lvgs - ms.von)/xn;
                m.gds() = gdson*tmp - m.gm()*dodvds - tmp1*dxndvd;
                m.gmbs() = gbson*tmp - m.gm()*dodvbs - tmp1*dxndvb;
            }
            return cdrain;
        }
        else {
            if (lvgs <= ms.von) {
                if (!fssandcap)
                    m.gds() = 0.0;
                else
                    m.gds() = beta1*(ms.von - vbin - gammad*sarg)*exp(argg*(lvgs - ms.von));
            }
            else
                m.gds() = beta1*(lvgs - vbin - gammad*sarg);
        }
    }
    m.gm() = 0.0;
    m.gmbs() = 0.0;
    return 0.0;
}

// moseq.cc:693-1083: level 3 (semi-empirical short-channel model)
__device__ inline double dev_mos_eq3(MosInst &m, MosStuff &ms)
{
    double gm, gds, gmbs;       // the operating-point conductances, stored on every exit
    double lvbs, lvds, lvgs;
    if (m.mode > 0) { lvbs = ms.vbs; lvds = ms.vds; lvgs = ms.vgs; }
    else { lvbs = ms.vbd; lvds = -ms.vds; lvgs = ms.vgd; }
    const double xl = m.xl(), tPhi = m.tPhi(), xd = m.xd(), alpha = m.alpha(), xj = m.xj(), gamma = m.gamma();
    const double nfs = m.nfs(), theta = m.theta(), vmax = m.vmax(), kappa = m.kappa();
    ms.vdsat = 0.0;
    double oneoverxl = 1.0/xl;
    double eta = m.eta()*8.15e-22/(m.oxideCapFactor()*xl*xl*xl);
    double phibs, sqphbs, dsqdvb;
    if (lvbs <= 0.0) {
        phibs = tPhi - lvbs;
        sqphbs = sqrt(phibs);
        dsqdvb = -0.5/sqphbs;
    }
    else {
        double sqphis = sqrt(tPhi);
        double sqphs3 = tPhi*sqphis;
        sqphbs = sqphis/(1.0 + lvbs/(tPhi + tPhi));
        phibs = sqphbs*sqphbs;
        dsqdvb = -phibs/(sqphs3+sqphs3);
    }
    double fshort, dfsdvb;
    if (xj != 0.0 && xd != 0.0) {
        double coeff0 = 0.0631353e0, coeff1 = 0.8013292e0, coeff2 = -0.01110777e0;
        double wps = xd*sqphbs;
        double oneoverxj = 1.0/xj;
        double xjonxl = xj*oneoverxl;
        double djonxj = m.ld()*oneoverxj;
        double wponxj = wps*oneoverxj;
        double wconxj = coeff0 + coeff1*wponxj + coeff2*wponxj*wponxj;
        double arga = wconxj + djonxj;
        double argc = wponxj/(1.0 + wponxj);
        double argb = sqrt(1.0 - argc*argc);
        fshort = 1.0 - xjonxl*(arga*argb - djonxj);
        double dwpdvb = xd*dsqdvb;
        double dadvb = (coeff1 + coeff2*(wponxj+wponxj))*dwpdvb*oneoverxj;
        double dbdvb = -argc*argc*(1.0 - argc)*dwpdvb/(argb*wps);
        dfsdvb = -xjonxl*(dadvb*argb + arga*dbdvb);
    }
    else {
        fshort = 1.0;
        dfsdvb = 0.0;
    }
    double gammas = gamma*fshort;
    double fbodys = 0.5*gammas/(sqphbs+sqphbs);
    double fbody = fbodys + 0.0/m.w();
    double onfbdy = 1.0/(1.0 + fbody);
    double dfbdvb = -fbodys*dsqdvb/sqphbs + fbodys*dfsdvb/fshort;
    double qbonco = gammas*sqphbs + 0.0*phibs/m.w();
    double dqbdvb = gammas*dsqdvb + gamma*dfsdvb*sqphbs - 0.0/m.w();
    double vth = qbonco + m.tVbi()*m.type - eta*lvds;
    double dvtdvd = -eta;
    double dvtdvb = dqbdvb;
    ms.von = vth;
    double cdrain;
    double xn = 1.0, dxndvb, dvodvd, dvodvb;
    if (nfs != 0.0) {
        double csonco = CHARGE*nfs*CM2PM2*xl*m.w()/m.oxideCap();
        double cdonco = qbonco/(phibs + phibs);
        xn += csonco + cdonco;
        ms.von += ms.vt*xn;
        dxndvb = dqbdvb/(phibs + phibs) - qbonco*dsqdvb/(phibs*sqphbs);
        dvodvd = dvtdvd;
        dvodvb = dvtdvb + ms.vt*dxndvb;
    }
    else {
        dxndvb = 0; dvodvd = 0; dvodvb = 0;
        if (lvgs <= ms.von) {
            m.gm() = 0.0; m.gds() = 0.0; m.gmbs() = 0.0;
            return 0.0;
        }
    }
    double vgsx = DMAX(lvgs, ms.von);
    double onfg = 1.0 + theta*(vgsx - vth);
    double fgate = 1.0/onfg;
    double us = m.tSurfMob()*M2PCM2*fgate;
    double dfgdvg = -theta*fgate*fgate;
    double dfgdvd = -dfgdvg*dvtdvd;
    double dfgdvb = -dfgdvg*dvtdvb;
    double dvsdvg, dvsdvd, dvsdvb, onvdsc;
    ms.vdsat = (vgsx - vth)*onfbdy;
    if (vmax <= 0.0) {
        dvsdvg = onfbdy;
        dvsdvd = -dvsdvg*dvtdvd;
        dvsdvb = -dvsdvg*dvtdvb - ms.vdsat*dfbdvb*onfbdy;
        onvdsc = 0;
    }
    else {
        double vdsc = xl*vmax/us;
        onvdsc = 1.0/vdsc;
        double arga = (vgsx - vth)*onfbdy;
        double argb = sqrt(arga*arga + vdsc*vdsc);
        ms.vdsat = arga + vdsc - argb;
        double dvsdga = (1.0 - arga/argb)*onfbdy;
        dvsdvg = dvsdga - (1.0 - vdsc/argb)*vdsc*dfgdvg*onfg;
        dvsdvd = -dvsdvg*dvtdvd;
        dvsdvb = -dvsdvg*dvtdvb - arga*dvsdga*dfbdvb;
    }
    double beta = m.beta();
    double vdsx = DMIN(lvds, ms.vdsat);
    if (vdsx == 0.0) {
        beta *= fgate;
        cdrain = 0.0;
        gm = 0.0;
        gds = beta*(vgsx - vth);
        gmbs = 0.0;
        if (nfs != 0.0 && lvgs < ms.von)
            gds *= exp((lvgs - ms.von)/(ms.vt*xn));
        m.gm() = gm; m.gds() = gds; m.gmbs() = gmbs;
        return cdrain;
    }
    double cdo = vgsx - vth - 0.5*(1.0 + fbody)*vdsx;
    double dcodvb = -dvtdvb - 0.5*dfbdvb*vdsx;
    double cdnorm = cdo*vdsx;
    gm = vdsx;
    if ((m.mode*ms.vds) > ms.vdsat)
        gds = -dvtdvd*vdsx;
    else
        gds = vgsx - vth - (1.0 + fbody + dvtdvd)*vdsx;
    gmbs = dcodvb*vdsx;
    double cd1 = beta*cdnorm;
    beta *= fgate;
    cdrain = beta*cdnorm;
    gm = beta*gm + dfgdvg*cd1;
    gds = beta*gds + dfgdvd*cd1;
    gmbs = beta*gmbs + dfgdvb*cd1;
    double fdrain, dfddvg, dfddvb, dfddvd;
    if (vmax > 0.0) {
        fdrain = 1.0/(1.0 + vdsx*onvdsc);
        double fd2 = fdrain*fdrain;
        double arga = fd2*vdsx*onvdsc*onfg;
        dfddvg = -dfgdvg*arga;
        if (m.mode*ms.vds > ms.vdsat)
            dfddvd = -dfgdvd*arga;
        else
            dfddvd = -dfgdvd*arga - fd2*onvdsc;
        dfddvb = -dfgdvb*arga;
        gm = fdrain*gm + dfddvg*cdrain;
        gds = fdrain*gds + dfddvd*cdrain;
        gmbs = fdrain*gmbs + dfddvb*cdrain;
        cdrain = fdrain*cdrain;
        beta *= fdrain;
    }
    else {
        fdrain = 0; dfddvg = 0; dfddvd = 0; dfddvb = 0;
    }
    double gds0 = 0;
    if (lvds > ms.vdsat && !(vmax != 0.0 && alpha == 0)) {
        double delxl, dldvd, ddldvg, ddldvd, ddldvb;
        if (vmax == 0.0) {
            delxl = sqrt(kappa*(lvds - ms.vdsat)*alpha);
            dldvd = 0.5*delxl/(lvds - ms.vdsat);
            ddldvg = 0.0;
            ddldvd = -dldvd;
            ddldvb = 0.0;
        }
        else {
            double cdsat = cdrain;
            double gdsat = cdsat*(1.0 - fdrain)*onvdsc;
            gdsat = DMAX(1.0e-12, gdsat);
            double gdoncd = gdsat/cdsat;
            double gdonfd = gdsat/(1.0 - fdrain);
            double gdonfg = gdsat*onfg;
            double dgdvg = gdoncd*gm - gdonfd*dfddvg + gdonfg*dfgdvg;
            double dgdvd = gdoncd*gds - gdonfd*dfddvd + gdonfg*dfgdvd;
            double dgdvb = gdoncd*gmbs - gdonfd*dfddvb + gdonfg*dfgdvb;
            double emax = kappa*cdsat*oneoverxl/gdsat;
            double emoncd = emax/cdsat;
            double emongd = emax/gdsat;
            double demdvg = emoncd*gm - emongd*dgdvg;
            double demdvd = emoncd*gds - emongd*dgdvd;
            double demdvb = emoncd*gmbs - emongd*dgdvb;
            double arga = 0.5*emax*alpha;
            double argc = kappa*alpha;
            double argb = sqrt(arga*arga + argc*(lvds - ms.vdsat));
            delxl = argb - arga;
            dldvd = argc/(argb+argb);
            double dldem = 0.5*(arga/argb - 1.0)*alpha;
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
        gm = gm*xlfact + diddl*ddldvg;
        gmbs = gmbs*xlfact + diddl*ddldvb;
        gds0 = diddl*ddldvd;
        gm = gm + gds0*dvsdvg;
        gmbs = gmbs + gds0*dvsdvb;
        gds = gds*xlfact + diddl*dldvd + gds0*dvsdvd;
    }
    if (lvgs < ms.von) {
        double onxn = 1.0/xn;
        double ondvt = onxn/ms.vt;
        double wfact = exp((lvgs - ms.von)*ondvt);
        cdrain = cdrain*wfact;
        double gms = gm*wfact;
        double gmw = cdrain*ondvt;
        gm = gmw;
        if (lvds > ms.vdsat)
            gm = gm + gds0*dvsdvg*wfact;
        gds = gds*wfact + (gms - gmw)*dvodvd;
        gmbs = gmbs*wfact + (gms - gmw)*dvodvb - gmw*(lvgs - ms.von)*onxn*dxndvb;
    }
    (void)cd1;
    m.gm() = gm; m.gds() = gds; m.gmbs() = gmbs;
    return cdrain;
}

// moseq.cc:1091-1172: level 6, n-th power law
__device__ inline double dev_mos_eq6(MosInst &m, MosStuff &ms)
{
    double lvds, lvbs, lvgs;
    if (m.mode > 0) { lvbs = ms.vbs; lvds = ms.vds; lvgs = ms.vgs; }
    else { lvbs = ms.vbd; lvds = -ms.vds; lvgs = ms.vgd; }
    double tPhi = m.tPhi(), gamma = m.gamma(), gamma1 = m.gamma1(), sigma = m.sigma();
    double sarg;
    if (lvbs <= 0)
        sarg = sqrt(tPhi - lvbs);
    else {
        sarg = sqrt(tPhi);
        sarg = sarg - lvbs/(sarg+sarg);
        sarg = DMAX(0, sarg);
    }
    ms.von = (m.tVbi()*m.type) + gamma*sarg - gamma1*lvbs - sigma*lvds;
    double vgon = lvgs - ms.von;
    double cdrain;
    if (vgon <= 0) {
        ms.vdsat = 0; m.gm() = 0; m.gds() = 0; m.gmbs() = 0; cdrain = 0;
    }
    else {
        double vonbm;
        if (sarg <= 0)
            vonbm = 0;
        else if (lvbs <= 0)
            vonbm = gamma1 + gamma/(sarg + sarg);
        else
            vonbm = gamma1 + .5*gamma/sqrt(tPhi);
        double nv = m.nv(), nc = m.nc(), lamda1 = m.lamda1();
        sarg = log(vgon);
        ms.vdsat = m.kv()*exp(sarg*nv);
        double idsat = m.beta()*exp(sarg*nc);
        double lambda = m.lamda0() - lamda1*lvbs;
        cdrain = idsat*(1 + lambda*lvds);
        double gm = cdrain*nc/vgon;
        double gds = gm*sigma + idsat*lambda;
        double gmbs = gm*vonbm - idsat*lamda1*lvds;
        if (ms.vdsat > lvds) {
            double vdst = lvds/ms.vdsat;
            double vdst2 = (2 - vdst)*vdst;
            double vdstg = -vdst*nv/vgon;
            double ivdst1 = cdrain*(2 - vdst - vdst);
            cdrain = cdrain*vdst2;
            gm = gm*vdst2 + ivdst1*vdstg;
            gds = gds*vdst2 + ivdst1*(1/ms.vdsat + vdstg*sigma);
            gmbs = gmbs*vdst2 + ivdst1*vdstg*vonbm;
        }
        m.gm() = gm; m.gds() = gds; m.gmbs() = gmbs;
    }
    return cdrain;
}

// mosload.cc:440-486.  Each limiter has ONE call site (operands selected by the channel direction
// first; negation is exact) so that a single inlined copy exists.
__device__ inline int dev_mos_limiting(const TrialCtx &c, MosInst &m, MosStuff &ms, int oldlimit)
{
    const double s0vds = m.o_vds(), s0vgs = m.o_vgs();
    const bool fwd = (s0vds >= 0);
    {
        double vg = dev_fetlim(fwd ? ms.vgs : ms.vgd, fwd ? s0vgs : s0vgs - s0vds, ms.von);
        if (fwd) ms.vgs = vg; else ms.vgd = vg;
    }
    ms.vds = ms.vgs - ms.vgd;
    if (fwd || !oldlimit) {
        double r = dev_limvds(fwd ? ms.vds : -ms.vds, fwd ? s0vds : -s0vds);
        ms.vds = fwd ? r : -r;
    }
    if (fwd) ms.vgd = ms.vgs - ms.vds;
    else ms.vgs = ms.vgd + ms.vds;
    int check;
    const bool pos = (ms.vds >= 0);
    {
        double vb = dev_pnjlim(pos ? ms.vbs : ms.vbd, pos ? m.o_vbs() : m.o_vbd(), ms.vt, m.rvt(),
            pos ? m.sourceVcrit() : m.drainVcrit(), &check);
        if (pos) { ms.vbs = vb; ms.vbd = ms.vbs - ms.vds; }
        else { ms.vbd = vb; ms.vbs = ms.vbd + ms.vds; }
    }
    ms.vgb = ms.vgs - ms.vbs;
    return check;
}

__device__ inline void dev_mos_iv(MosInst &m, MosStuff &ms, double gmin)
{
    ms.gcgs = 0; ms.ceqgs = 0; ms.gcgd = 0; ms.ceqgd = 0; ms.gcgb = 0; ms.ceqgb = 0;
    // bulk-source, then bulk-drain diode (mosload.cc:520-547): one copy of the code, two trips
#pragma unroll 1
    for (int j = 0; j < 2; j++) {
        const double v = j ? ms.vbd : ms.vbs;
        const double isat = m.P(j ? MP_tDrainSatCur : MP_tSourceSatCur);
        double g, cur;
        if (v <= -3*ms.vt) {
            g = gmin;
            cur = g*v - isat;
        }
        else {
            double rvt = m.rvt();
            double ev = exp(dev_div_rcp(v, ms.vt, rvt));
            g = dev_div_rcp(isat*ev, ms.vt, rvt) + gmin;
            cur = isat*(ev-1) + gmin*v;
        }
        m.R(SR_CACHE + (j ? MC_gbd : MC_gbs)) = g;
        m.R(SR_CACHE + (j ? MC_cbd : MC_cbs)) = cur;
    }
    m.mode = (ms.vds >= 0) ? 1 : -1;
    // mosload.cc:570-577
    ms.cdrain = (m.level == 6) ? dev_mos_eq6(m, ms) : (m.level == 3) ? dev_mos_eq3(m, ms) :
        (m.level == 2) ? dev_mos_eq2(m, ms) : dev_mos_eq1(m, ms);
    if (m.type > 0) {
        m.von() = ms.von; m.vdsat() = ms.vdsat;
    }
    else {
        m.von() = -ms.von; m.vdsat() = -ms.vdsat;
    }
    if (m.mode > 0)
        m.cd() = ms.cdrain - m.cbd();
    else
        m.cd() = -ms.cdrain - m.cbd();
}

__device__ inline void dev_mos_cap(const TrialCtx &c, MosInst &m, MosStuff &ms)
{
    const int st = m.st;
    // bulk-source, then bulk-drain depletion charge (mosload.cc:610-690): one copy, two trips
#pragma unroll 1
    for (int j = 0; j < 2; j++) {
        const double v = j ? ms.vbd : ms.vbs;
        const double C = m.P(j ? MP_Cbd : MP_Cbs), Csw = m.P(j ? MP_Cbdsw : MP_Cbssw);
        double qv = 0, capv = 0;
        if (C != 0 || Csw != 0) {
            if (v < m.tDepCap()) {
                double arg = 1 - dev_div_rcp(v, m.tBulkPot(), m.rBulkPot()), sarg, sargsw;
                dev_sargs(arg, m.mj(), m.mjsw(), &sarg, &sargsw);
                qv = m.tBulkPot()*(dev_div_rcp(C*(1 - arg*sarg), 1 - m.mj(), m.r1mj())
                    + dev_div_rcp(Csw*(1 - arg*sargsw), 1 - m.mjsw(), m.r1mjsw()));
                capv = C*sarg + Csw*sargsw;
            }
            else {
                const double f2 = m.P(j ? MP_f2d : MP_f2s), f3 = m.P(j ? MP_f3d : MP_f3s), f4 = m.P(j ? MP_f4d : MP_f4s);
                qv = f4 + v*(f2 + .5*v*f3);
                capv = f2 + f3*v;
            }
        }
        m.s0(j ? MS_qbd : MS_qbs) = qv;
        m.R(SR_CACHE + (j ? MC_capbd : MC_capbs)) = capv;
    }

    double vgx, vgy, cgx, cgy, cgb;
    if (m.mode > 0) { vgx = ms.vgs; vgy = ms.vgd; }
    else { vgx = ms.vgd; vgy = ms.vgs; }
    dev_qmeyer(vgx, vgy, ms.von, ms.vdsat, &cgx, &cgy, &cgb, m.tPhi(), m.oxideCap(), m.r2phi(), m.r15phi());
    double s0capgs, s0capgd;
    if (m.mode > 0) { s0capgs = cgx; s0capgd = cgy; }
    else { s0capgd = cgx; s0capgs = cgy; }
    m.s0(MS_capgs) = s0capgs;
    m.s0(MS_capgd) = s0capgd;
    m.s0(MS_capgb) = cgb;

    if (c.mode & (MODEINITTRAN|MODEINITSMSIG)) {
        ST(c, 1, st + MS_capgs) = s0capgs;
        ST(c, 1, st + MS_capgd) = s0capgd;
        ST(c, 1, st + MS_capgb) = cgb;
        m.s1(MS_capgs) = s0capgs; m.s1(MS_capgd) = s0capgd; m.s1(MS_capgb) = cgb;
        ms.capgs = s0capgs + s0capgs + m.gsOvl();
        ms.capgd = s0capgd + s0capgd + m.gdOvl();
        ms.capgb = cgb + cgb + m.gbOvl();
        double qgs = ms.capgs*ms.vgs, qgd = ms.capgd*ms.vgd, qgb = ms.capgb*ms.vgb;
        m.s0(MS_qgs) = qgs;
        m.s0(MS_qgd) = qgd;
        m.s0(MS_qgb) = qgb;
        double qbd0 = m.s0(MS_qbd), qbs0 = m.s0(MS_qbs);
        ST(c, 1, st + MS_qbd) = qbd0;
        ST(c, 1, st + MS_qbs) = qbs0;
        ST(c, 1, st + MS_qgs) = qgs;
        ST(c, 1, st + MS_qgd) = qgd;
        ST(c, 1, st + MS_qgb) = qgb;
        m.s1(MS_qbd) = qbd0; m.s1(MS_qbs) = qbs0;
        m.s1(MS_qgs) = qgs; m.s1(MS_qgd) = qgd; m.s1(MS_qgb) = qgb;
    }
    else {
        ms.capgs = s0capgs + m.s1(MS_capgs) + m.gsOvl();
        ms.capgd = s0capgd + m.s1(MS_capgd) + m.gdOvl();
        ms.capgb = cgb + m.s1(MS_capgb) + m.gbOvl();
        vgx = m.s1(MS_vgs);
        vgy = vgx - m.s1(MS_vds);
        double vxy = vgx - m.s1(MS_vbs);
        m.s0(MS_qgs) = (ms.vgs - vgx)*ms.capgs + m.s1(MS_qgs);
        m.s0(MS_qgd) = (ms.vgd - vgy)*ms.capgd + m.s1(MS_qgd);
        m.s0(MS_qgb) = (ms.vgb - vxy)*ms.capgb + m.s1(MS_qgb);
    }
}

// sCKT::integrate (niinteg.cc:60-152) on the staged state rows of one charge: k = MS_q*,
// k+1 = its current.
__device__ inline void dev_mos_integrate(const TrialCtx &c, const MosInst &m, double *geq, double *ceq,
    double cap, int k)
{
    double s1q = m.s1(k), s1c = m.s1(k + 1);
    double cq;
    if (c.method == 0)
        cq = (c.order <= 1 ? s1q*c.ag[1] : - s1c*c.ag[1] - s1q*c.ag[0]);
    else        // Gear (the only other method)
        cq = dev_NIceq_gear(c.states, c.base, c.NS, c.numStates, c.ldt, c.agp, c.ag[1], c.order, m.st + k);
    double q0 = m.s0(k);
    double c0 = c.ag[0]*q0 + cq;
    m.s0(k + 1) = c0;
    *ceq = cq;
    *geq = c.ag[0]*cap;
    // trapezoid ringing check (niinteg.cc:117-152)
    if (c.trapCheck && !*c.trapbad && !(fabs(c0) < 1e-6) && !(fabs(q0) < 1e-16)) {
        double c2 = m.sg2[(k + 1)*c.ldt];
        double dqdt = fabs(q0 - s1q)/c.delta;
        if (fabs(s1c) > c.trapratio*dqdt && fabs(c0 - s1c) > c.trapratio*fabs(c0 - c2))
            *c.trapbad = 1;
    }
}

__device__ inline void dev_mos_meyer_integ(const TrialCtx &c, MosInst &m, MosStuff &ms)
{
    dev_mos_integrate(c, m, &ms.gcgs, &ms.ceqgs, ms.capgs, MS_qgs);
    dev_mos_integrate(c, m, &ms.gcgd, &ms.ceqgd, ms.capgd, MS_qgd);
    dev_mos_integrate(c, m, &ms.gcgb, &ms.ceqgb, ms.capgb, MS_qgb);
    ms.ceqgs += c.ag[0]*m.s0(MS_qgs) - ms.gcgs*ms.vgs;
    ms.ceqgd += c.ag[0]*m.s0(MS_qgd) - ms.gcgd*ms.vgd;
    ms.ceqgb += c.ag[0]*m.s0(MS_qgb) - ms.gcgb*ms.vgb;
}

__device__ inline void dev_mos_integ(const TrialCtx &c, MosInst &m, MosStuff &ms)
{
    double geq, ceq;
    dev_mos_integrate(c, m, &geq, &ceq, m.capbd(), MS_qbd);
    m.gbd() += geq;
    double cqbd = m.s0(MS_cqbd);
    m.cbd() += cqbd;
    m.cd()  -= cqbd;
    dev_mos_integrate(c, m, &geq, &ceq, m.capbs(), MS_qbs);
    m.gbs() += geq;
    m.cbs() += m.s0(MS_cqbs);
    if (ms.capgs == 0) m.s0(MS_cqgs) = 0;
    if (ms.capgd == 0) m.s0(MS_cqgd) = 0;
    if (ms.capgb == 0) m.s0(MS_cqgb) = 0;
}

// Unified stamp (tran form); with the Meyer terms zero it reproduces load_dc exactly
// (x+0 == x, -(x+0) == -x), see DESIGN.md section 4.
__device__ inline void dev_mos_stamp(const TrialCtx &c, MosInst &m, MosStuff &ms, const MosOut &o)
{
#define MSC(xx) (m.mult()*(xx))
    m.s0(MS_vbs) = ms.vbs;
    m.s0(MS_vbd) = ms.vbd;
    m.s0(MS_vgs) = ms.vgs;
    m.s0(MS_vds) = ms.vds;
    double gbd = MSC(m.gbd()), gbs = MSC(m.gbs()), gds = MSC(m.gds()), gm = MSC(m.gm());
    double gmbs = MSC(m.gmbs()), cbd = MSC(m.cbd()), cbs = MSC(m.cbs());
    double ceqbs = cbs - gbs*ms.vbs;
    double ceqbd = cbd - gbd*ms.vbd;
    double cdreq;
    if (m.mode > 0)
        cdreq = MSC(ms.cdrain) - gds*ms.vds - gm*ms.vgs - gmbs*ms.vbs;
    else
        cdreq = -(MSC(ms.cdrain) + gds*ms.vds - gm*ms.vgd - gmbs*ms.vbd);
    double dcon = MSC(m.dcon()), scon = MSC(m.scon());
    if (m.type > 0) {
        o.Rw(MRH_g, -MSC(ms.ceqgs + ms.ceqgb + ms.ceqgd));
        o.Rw(MRH_b, -(ceqbs + ceqbd - MSC(ms.ceqgb)));
        o.Rw(MRH_dp, ceqbd - cdreq + MSC(ms.ceqgd));
        o.Rw(MRH_sp, cdreq + ceqbs + MSC(ms.ceqgs));
    }
    else {
        o.Rw(MRH_g, MSC(ms.ceqgs + ms.ceqgb + ms.ceqgd));
        o.Rw(MRH_b, ceqbs + ceqbd - MSC(ms.ceqgb));
        o.Rw(MRH_dp, -(ceqbd - cdreq + MSC(ms.ceqgd)));
        o.Rw(MRH_sp, -(cdreq + ceqbs + MSC(ms.ceqgs)));
    }
    double gcgd = MSC(ms.gcgd), gcgs = MSC(ms.gcgs), gcgb = MSC(ms.gcgb);
    o.M(MT_Dd, dcon);
    o.M(MT_Gg, gcgd + gcgs + gcgb);
    o.M(MT_Ss, scon);
    o.M(MT_Bb, gbd + gbs + gcgb);
    o.M(MT_Ddp, -dcon);
    o.M(MT_Gb, -gcgb);
    o.M(MT_Gdp, -gcgd);
    o.M(MT_Gsp, -gcgs);
    o.M(MT_Ssp, -scon);
    o.M(MT_Bg, -gcgb);
    o.M(MT_Bdp, -gbd);
    o.M(MT_Bsp, -gbs);
    o.M(MT_DPd, -dcon);
    o.M(MT_SPs, -scon);
    if (m.mode > 0) {
        o.M(MT_DPdp, dcon + gds + gbd + gcgd);
        o.M(MT_SPsp, scon + gds + gbs + gm + gmbs + gcgs);
        o.M(MT_DPg, gm - gcgd);
        o.M(MT_DPb, -gbd + gmbs);
        o.M(MT_DPsp, -(gds + gm + gmbs));
        o.M(MT_SPg, -(gm + gcgs));
        o.M(MT_SPb, -(gbs + gmbs));
        o.M(MT_SPdp, -gds);
    }
    else {
        o.M(MT_DPdp, dcon + gds + gbd + gm + gmbs + gcgd);
        o.M(MT_SPsp, scon + gds + gbs + gcgs);
        o.M(MT_DPg, -(gm + gcgd));
        o.M(MT_DPb, -(gbd + gmbs));
        o.M(MT_DPsp, -gds);
        o.M(MT_SPg, -(-gm + gcgs));
        o.M(MT_SPb, -(gbs - gmbs));
        o.M(MT_SPdp, -(gds + gm + gmbs));
    }
#undef MSC
}

// MOS bypass test (mosload.cc:381-470): true if the cached operating point may be reused.
__device__ inline bool dev_mos_bypass_ok(const MosInst &m, const MosStuff &ms,
    double reltol, double abstol, double vntol)
{
    double s0vbs = m.o_vbs(), s0vgs = m.o_vgs();
    double s0vds = m.o_vds(), s0vbd = m.o_vbd();
    double delvbs = ms.vbs - s0vbs;
    double delvgs = ms.vgs - s0vgs;
    double delvds = ms.vds - s0vds;
    double delvbd = ms.vbd - s0vbd;
    double delvgd = ms.vgd - (s0vgs - s0vds);
    double cbs = m.cbs(), cbd = m.cbd(), gbd = m.gbd(), gbs = m.gbs();
    double cbhat = cbs + cbd + gbd*delvbd + gbs*delvbs;
    double A1 = fabs(cbhat);
    double A2 = cbs + cbd;
    A2 = fabs(A2);
    double A3 = DMAX(A1, A2) + abstol;
    A1 = cbhat - (cbs + cbd);
    if (fabs(A1) >= reltol*A3) return false;
    A1 = fabs(ms.vbs); A2 = fabs(s0vbs);
    if (fabs(delvbs) >= (reltol*DMAX(A1, A2) + vntol)) return false;
    A1 = fabs(ms.vbd); A2 = fabs(s0vbd);
    if (fabs(delvbd) >= (reltol*DMAX(A1, A2) + vntol)) return false;
    A1 = fabs(ms.vgs); A2 = fabs(s0vgs);
    if (fabs(delvgs) >= (reltol*DMAX(A1, A2) + vntol)) return false;
    A1 = fabs(ms.vds); A2 = fabs(s0vds);
    if (fabs(delvds) >= (reltol*DMAX(A1, A2) + vntol)) return false;
    double cd = m.cd(), gmbs = m.gmbs(), gm = m.gm(), gds = m.gds();
    double cdhat;
    if (m.mode >= 0)
        cdhat = cd - gbd*delvbd + gmbs*delvbs + gm*delvgs + gds*delvds;
    else
        cdhat = cd + (gmbs - gbd)*delvbd - gm*delvgd + gds*delvds;
    A1 = cdhat - cd;
    A2 = fabs(cdhat);
    A3 = fabs(cd);
    if ((fabs(A1) >= reltol*DMAX(A2, A3) + abstol)) return false;
    return true;
}

__device__ inline void dev_mos_get_v(const MosInst &m, MosStuff &ms)
{
    double vb = m.R(SR_V_B), vsp = m.R(SR_V_SP);
    double vg = m.R(SR_V_G), vdp = m.R(SR_V_DP);
    if (m.type > 0) {
        ms.vbs = vb - vsp;
        ms.vgs = vg - vsp;
        ms.vds = vdp - vsp;
    }
    else {
        ms.vbs = vsp - vb;
        ms.vgs = vsp - vg;
        ms.vds = vsp - vdp;
    }
}

// MOSdev::load for one (instance, trial) (mosload.cc:100-930).  The mode branches only choose
// the terminal voltages and which stages run; limiting -> I/V -> charges -> integration -> stamp
// exist once, in the order the reference executes them.
// Returns: bit0 = stamps valid (contributions written), bit1 = cache dirty, bit2 = incNoncon,
// bit3 = bypassed, bit4 = only the .ic voltages were stored (UIC)
__device__ inline int dev_mos_load(const TrialCtx &c, MosInst &m, const MosOut &o, const b200_options &opt)
{
    MosStuff ms;
    ms.vt = m.vt();
    ms.gcgs = ms.ceqgs = ms.gcgd = ms.ceqgd = ms.gcgb = ms.ceqgb = 0;
    ms.capgs = ms.capgd = ms.capgb = 0; ms.cdrain = 0; ms.vdsat = 0;
    bool limit = false, limit_noncon = false, charges = false;
    int ret = 3;
    ms.von = (m.type > 0) ? m.von() : -m.von();
    if (c.mode & MODEINITFLOAT) {
        dev_mos_get_v(m, ms);
        limit = true;
        charges = (c.mode & MODETRAN) != 0;
    }
    else if (c.mode & MODEINITPRED) {
        double s1vbs = m.s1(MS_vbs), s1vgs = m.s1(MS_vgs), s1vds = m.s1(MS_vds);
        ms.vbs = s1vbs*c.pred[0] + m.R(SR_S2_vbs)*c.pred[1];
        ms.vgs = s1vgs*c.pred[0] + m.R(SR_S2_vgs)*c.pred[1];
        ms.vds = s1vds*c.pred[0] + m.R(SR_S2_vds)*c.pred[1];
        // limiting reads the state-0 voltages the reference has just copied from state 1
        // (mosload.cc:172-182); the stamp stage then stores the limited values to state 0
        m.o_vbs() = s1vbs; m.o_vgs() = s1vgs; m.o_vds() = s1vds; m.o_vbd() = s1vbs - s1vds;
        limit = true;
        charges = (c.mode & MODETRAN) != 0;
    }
    else if (c.mode & MODEINITFIX) {
        if (m.off)
            ms.vbs = ms.vgs = ms.vds = 0;
        else {
            dev_mos_get_v(m, ms);
            limit = true; limit_noncon = true;
        }
    }
    else if (c.mode & MODEINITTRAN) {
        ms.vbs = m.s1(MS_vbs);
        ms.vgs = m.s1(MS_vgs);
        ms.vds = m.s1(MS_vds);
        charges = true;
    }
    else if (c.mode & MODEINITJCT) {
        double icVDS = m.icVDS(), icVGS = m.icVGS(), icVBS = m.icVBS();
        if ((c.mode & MODETRANOP) && (c.mode & MODEUIC)) {
            if (m.type > 0) {
                ms.vds = icVDS; ms.vgs = icVGS; ms.vbs = icVBS;
            }
            else {
                ms.vds = -icVDS; ms.vgs = -icVGS; ms.vbs = -icVBS;
            }
            m.s0(MS_vds) = ms.vds;
            m.s0(MS_vgs) = ms.vgs;
            m.s0(MS_vbs) = ms.vbs;
            return 16;
        }
        if (m.off)
            ms.vbs = ms.vgs = ms.vds = 0;
        else if (icVDS == 0 && icVGS == 0 && icVBS == 0) {
            ms.vbs = -1;
            ms.vgs = (m.type > 0) ? m.tVto() : -m.tVto();
            ms.vds = 0;
        }
        else if (m.type > 0) {
            ms.vds = icVDS; ms.vgs = icVGS; ms.vbs = icVBS;
        }
        else {
            ms.vds = -icVDS; ms.vgs = -icVGS; ms.vbs = -icVBS;
        }
    }
    else
        return 0;
    ms.vbd = ms.vbs - ms.vds;
    ms.vgd = ms.vgs - ms.vds;
    ms.vgb = ms.vgs - ms.vbs;

    if ((c.mode & MODEINITFLOAT) && opt.bypass &&
            dev_mos_bypass_ok(m, ms, opt.reltol, opt.abstol, opt.vntol)) {
        ret = 1 | 8;
        ms.vbs = m.o_vbs();
        ms.vgs = m.o_vgs();
        ms.vds = m.o_vds();
        ms.vbd = m.o_vbd();
        ms.vgd = ms.vgs - ms.vds;
        ms.vgb = ms.vgs - ms.vbs;
        ms.cdrain = m.mode*(m.cd() + m.cbd());
        if (charges) {
            ms.capgs = m.o_capgs() + m.s1(MS_capgs) + m.gsOvl();
            ms.capgd = m.o_capgd() + m.s1(MS_capgd) + m.gdOvl();
            ms.capgb = m.o_capgb() + m.s1(MS_capgb) + m.gbOvl();
        }
    }
    else {
        if (limit) {
            int check = dev_mos_limiting(c, m, ms, opt.oldlimit);
            if (limit_noncon && check) ret |= 4;
        }
        dev_mos_iv(m, ms, opt.gmin);
        if (charges) {
            dev_mos_cap(c, m, ms);
            dev_mos_integ(c, m, ms);
        }
    }
    if (charges)
        dev_mos_meyer_integ(c, m, ms);
    dev_mos_stamp(c, m, ms, o);
    return ret;
}

// MOSdev::convTest for one instance (mosconv.cc:52-141).  rhs = CKTrhs (PREVIOUS iterate).
struct MosConvIn {
    int type, off, st, mode;
    int g, b, dp, sp;
    double cd, cbs, cbd, gm, gds, gmbs, gbd, gbs;
};
__device__ inline int dev_mos_conv(const TrialCtx &c, const MosConvIn &m, const double *rhs,
    double reltol, double abstol)
{
    if (m.off && (c.mode & MODEINITFIX))
        return 0;
    double vb = rhs[(size_t)m.b*c.ldt], vsp = rhs[(size_t)m.sp*c.ldt];
    double vg = rhs[(size_t)m.g*c.ldt], vdp = rhs[(size_t)m.dp*c.ldt];
    double vbs, vgs, vds;
    if (m.type > 0) {
        vbs = vb - vsp; vgs = vg - vsp; vds = vdp - vsp;
    }
    else {
        vbs = vsp - vb; vgs = vsp - vg; vds = vsp - vdp;
    }
    const int st = m.st;
    double s0vbs = ST(c, 0, st + MS_vbs), s0vgs = ST(c, 0, st + MS_vgs);
    double s0vds = ST(c, 0, st + MS_vds), s0vbd = ST(c, 0, st + MS_vbd);
    double vbd = vbs - vds, vgd = vgs - vds;
    double delvbs = vbs - s0vbs;
    double delvbd = vbd - s0vbd;
    double delvgs = vgs - s0vgs;
    double delvds = vds - s0vds;
    double delvgd = vgd - (s0vgs - s0vds);
    double cdhat;
    if (m.mode >= 0)
        cdhat = m.cd - m.gbd*delvbd + m.gmbs*delvbs + m.gm*delvgs + m.gds*delvds;
    else
        cdhat = m.cd - (m.gbd - m.gmbs)*delvbd - m.gm*delvgd + m.gds*delvds;
    double A1 = fabs(cdhat), A2 = fabs(m.cd), A3 = cdhat - m.cd;
    A3 = fabs(A3);
    double tol = reltol*DMAX(A1, A2) + abstol;
    if (A3 >= tol)
        return 1;
    double cbhat = m.cbs + m.cbd + m.gbd*delvbd + m.gbs*delvbs;
    A1 = fabs(cbhat);
    A2 = m.cbs + m.cbd;
    A2 = fabs(A2);
    A3 = cbhat - (m.cbs + m.cbd);
    A3 = fabs(A3);
    tol = reltol*DMAX(A1, A2) + abstol;
    if (A3 > tol)
        return 1;
    return 0;
}

// ---- MOS temperature preprocessing (mosset.cc:283-373 + mostemp.cc:65-389) ----
struct MosCard {            // per instance, shared by all trials (device array)
    int type, level;
    int g_vto, g_kp, g_gamma, g_phi, g_tox, g_u0, g_nsub, g_cbd, g_cbs, g_cj, g_cjsw, g_nss, g_tpg;
    double cbd, cbs, is, pb, cgso, cgdo, cgbo, cj, mj, cjsw, mjsw, js, tox, ld, u0, fc, nsub,
        tpg, nss, tnom;
    double drainArea, sourceArea, drainPerimeter, sourcePerimeter, mult, temp;
    double dcon, scon, icVDS, icVGS, icVBS;
    double kv, nv, nc, lamda0, lamda1, sigma, gamma1;    // level 6
    double nfs, vmax, xj, eta, theta, kappa;             // level 3
    double uexp, ucrit, neff, narrow;                    // level 2
};

// raw[] = vto kp gamma phi lambda l w (NaN = not given for vto/kp/gamma/phi)
__device__ inline void dev_mos_temp(const MosCard &k, const double *raw, double *out)
{
    double vto = raw[MR_vto], kp = raw[MR_kp], gamma = raw[MR_gamma], phi = raw[MR_phi];
    double lambda = raw[MR_lambda], l = raw[MR_l], w = raw[MR_w];
    // mosset.cc defaults
    if (!k.g_vto) vto = 0;
    if (!k.g_gamma) gamma = 0;
    if (!k.g_phi) phi = .6;
    double effectiveLength = l - 2*k.ld;
    if (effectiveLength < 0) effectiveLength = 0;

    double tnom = k.tnom;
    double fact1  = tnom/REFTEMP;
    double vtnom  = tnom*CONSTKoverQ;
    double egfet1 = 1.16 - (7.02e-4*tnom*tnom)/(tnom + 1108);
    double tmp    = tnom*CONSTboltz;
    tmp = -egfet1/(tmp+tmp) + 1.1150877/(CONSTboltz*(REFTEMP+REFTEMP));
    double pbfact1 = -2*vtnom*(1.5*log(fact1) + CHARGE*tmp);

    double tox = k.tox, oxideCapFactor, surfaceMobility;
    int skip = 0;
    if (!k.g_tox || tox == 0) {
        if (k.level == 1 || k.level == 6) {
            oxideCapFactor = 0;
            skip = 1;
        }
        else
            tox = 1e-7;
    }
    int have_kp = k.g_kp;
    if (!skip) {
        oxideCapFactor = 3.9*8.854214871e-12/tox;
        surfaceMobility = k.g_u0 ? k.u0 : 600;
        if (!have_kp) {
            // level 6: the raw "kp" slot carries kc (mostemp.cc:100-111)
            if (k.level == 6) kp = .5*surfaceMobility*oxideCapFactor*M2PCM2;
            else kp = surfaceMobility*oxideCapFactor*M2PCM2;
            have_kp = 1;
        }
        if (k.g_nsub) {
            double doping = k.nsub*CM3PM3;
            if (doping > 1.45e16) {
                if (!k.g_phi) {
                    phi = 2*vtnom*log(doping/1.45e16);
                    phi = DMAX(.1, phi);
                }
                double gateType = k.g_tpg ? k.tpg : 1;
                if (!k.g_gamma)
                    gamma = sqrt(2*EPSSIL*CHARGE*doping)/oxideCapFactor;
                if (!k.g_vto) {
                    double wkfng = 3.2;
                    if (gateType != 0) {
                        if (k.type > 0)
                            wkfng = 3.25 + .5*egfet1*(1 - gateType);
                        else
                            wkfng = 3.25 + .5*egfet1*(1 + gateType);
                    }
                    double wkfngs;
                    if (k.type > 0)
                        wkfngs = wkfng - (3.25 + .5*(egfet1 + phi));
                    else
                        wkfngs = wkfng - (3.25 + .5*(egfet1 - phi));
                    double vfb;
                    if (!k.g_nss)
                        vfb = wkfngs;
                    else
                        vfb = wkfngs - k.nss*CM2PM2*CHARGE/oxideCapFactor;
                    if (k.type > 0)
                        vto = vfb + (gamma*sqrt(phi) + phi);
                    else
                        vto = vfb - (gamma*sqrt(phi) + phi);
                }
            }
        }
    }
    else
        surfaceMobility = k.g_u0 ? k.u0 : 0;
    if (!have_kp)
        kp = (k.level == 6) ? 5e-5 : 2e-5;       // mosset.cc:101-103, 260-261

    double temp = k.temp;
    double fact2 = temp/REFTEMP;
    double vt    = temp*CONSTKoverQ;
    double ratio = temp/tnom;
    double egfet = 1.16 - (7.02e-4*temp*temp)/(temp + 1108);
    tmp = temp*CONSTboltz;
    tmp = -egfet/(tmp+tmp) + 1.1150877/(CONSTboltz*(REFTEMP+REFTEMP));
    double pbfact = -2*vt*(1.5*log(fact2) + CHARGE*tmp);

    tmp = 1/(ratio*sqrt(ratio));
    double tTransconductance = kp*tmp;
    const double tmp_mob = tmp;          // 1/(ratio*sqrt(ratio)), mostemp.cc:155

    tmp = (phi - pbfact1)/fact1;
    double tPhi = fact2*tmp + pbfact;
    double tVbi, tVto;
    if (k.type > 0) {
        tVbi = vto - gamma*sqrt(phi) + .5*(egfet1 - egfet + tPhi - phi);
        tVto = tVbi + gamma*sqrt(tPhi);
    }
    else {
        tVbi = vto + gamma*sqrt(phi) + .5*(egfet1 - egfet - (tPhi - phi));
        tVto = tVbi - gamma*sqrt(tPhi);
    }
    tmp = exp(-egfet/vt + egfet1/vtnom);
    double tSatCur = k.is*tmp;
    double tSatCurDens = k.js*tmp;

    double pbo = (k.pb - pbfact1)/fact1;
    double gmaold = (k.pb - pbo)/pbo;
    double capfact = 1/(1 + k.mj*(4e-4*(tnom - REFTEMP) - gmaold));
    double tCbd = k.cbd*capfact;
    double tCbs = k.cbs*capfact;
    double tCj  = k.cj*capfact;
    capfact = 1/(1 + k.mjsw*(4e-4*(tnom - REFTEMP) - gmaold));
    double tCjsw = k.cjsw*capfact;
    double tBulkPot = fact2*pbo + pbfact;
    double gmanew = (tBulkPot - pbo)/pbo;
    capfact = (1 + k.mj*(4e-4*(temp - REFTEMP) - gmanew));
    tCbd *= capfact;
    tCbs *= capfact;
    tCj  *= capfact;
    capfact = (1 + k.mjsw*(4e-4*(temp - REFTEMP) - gmanew));
    tCjsw *= capfact;
    double tDepCap = k.fc*tBulkPot;

    double sourceVcrit, drainVcrit, tDrainSatCur, tSourceSatCur;
    if ((tSatCurDens == 0) || (k.drainArea == 0) || (k.sourceArea == 0)) {
        sourceVcrit = drainVcrit = vt*log(vt/(CONSTroot2*tSatCur));
        tDrainSatCur = tSatCur;
        tSourceSatCur = tSatCur;
    }
    else {
        tDrainSatCur = tSatCurDens*k.drainArea;
        tSourceSatCur = tSatCurDens*k.sourceArea;
        drainVcrit = vt*log(vt/(CONSTroot2*tDrainSatCur));
        sourceVcrit = vt*log(vt/(CONSTroot2*tSourceSatCur));
    }
    double Cbd, Cbdsw, Cbs, Cbssw;
    if (k.g_cbd) Cbd = tCbd;
    else Cbd = k.g_cj ? tCj*k.drainArea : 0;
    Cbdsw = k.g_cjsw ? tCjsw*k.drainPerimeter : 0;
    if (k.g_cbs) Cbs = tCbs;
    else Cbs = k.g_cj ? tCj*k.sourceArea : 0;
    Cbssw = k.g_cjsw ? tCjsw*k.sourcePerimeter : 0;

    double arg = 1 - k.fc;
    double sarg = exp((-k.mj)*log(arg));
    double sargsw = exp((-k.mjsw)*log(arg));
    double pot = tBulkPot;
    double tmp1 = (1 - k.fc*(1 + k.mj))/arg;
    double tmp2 = 1/(arg*pot);
    double f2d = (Cbd*sarg + Cbdsw*sargsw)*tmp1;
    double f3d = (Cbd*k.mj*sarg + Cbdsw*k.mjsw*sargsw)*tmp2;
    double f4d = Cbd*pot*(1 - arg*sarg)/(1 - k.mj)
        + Cbdsw*pot*(1 - arg*sargsw)/(1 - k.mjsw)
        - tDepCap*(f2d + .5*f3d*tDepCap);
    double f2s = (Cbs*sarg + Cbssw*sargsw)*tmp1;
    double f3s = (Cbs*k.mj*sarg + Cbssw*k.mjsw*sargsw)*tmp2;
    double f4s = Cbs*pot*(1 - arg*sarg)/(1 - k.mj)
        + Cbssw*pot*(1 - arg*sargsw)/(1 - k.mjsw)
        - tDepCap*(f2s + .5*f3s*tDepCap);

    out[MP_tPhi] = tPhi; out[MP_tVbi] = tVbi; out[MP_tVto] = tVto;
    out[MP_gamma] = gamma; out[MP_lambda] = lambda;
    out[MP_beta] = tTransconductance*w/effectiveLength;
    out[MP_oxideCap] = oxideCapFactor*effectiveLength*w;
    out[MP_tSourceSatCur] = tSourceSatCur; out[MP_tDrainSatCur] = tDrainSatCur;
    out[MP_sourceVcrit] = sourceVcrit; out[MP_drainVcrit] = drainVcrit;
    out[MP_Cbd] = Cbd; out[MP_Cbdsw] = Cbdsw; out[MP_Cbs] = Cbs; out[MP_Cbssw] = Cbssw;
    out[MP_tBulkPot] = tBulkPot; out[MP_tDepCap] = tDepCap;
    out[MP_f2d] = f2d; out[MP_f3d] = f3d; out[MP_f4d] = f4d;
    out[MP_f2s] = f2s; out[MP_f3s] = f3s; out[MP_f4s] = f4s;
    out[MP_gsOvl] = k.cgso*w; out[MP_gdOvl] = k.cgdo*w; out[MP_gbOvl] = k.cgbo*effectiveLength;
    out[MP_dcon] = k.dcon; out[MP_scon] = k.scon; out[MP_mult] = k.mult; out[MP_vt] = vt;
    out[MP_mj] = k.mj; out[MP_mjsw] = k.mjsw;
    out[MP_rvt] = 1.0/out[MP_vt]; out[MP_rBulkPot] = 1.0/out[MP_tBulkPot];
    out[MP_r1mj] = 1.0/(1 - k.mj); out[MP_r1mjsw] = 1.0/(1 - k.mjsw);
    out[MP_r2phi] = 1.0/(2*out[MP_tPhi]); out[MP_r15phi] = 1.0/(1.5*out[MP_tPhi]);
    // level 3: mostemp.cc:119-142.  The narrow-width factor is always 0 in the reference: the
    // card's "delta" lands in MOSnarrowFactor (mossetm.cc:216-219) and mostemp overwrites it
    // with MOSdelta (default 0) * pi/2 * EPSSIL / oxideCapFactor.
    {
        double alpha = 0, xd = 0;
        if (!skip && k.g_nsub) {
            if (k.level == 2)
                xd = sqrt((EPSSIL+EPSSIL)/(CHARGE*k.nsub*CM3PM3));
            else if (k.level == 3) {
                alpha = (EPSSIL+EPSSIL)/(CHARGE*k.nsub*CM3PM3);
                xd = sqrt(alpha);
            }
        }
        out[MP_xd] = xd; out[MP_alpha] = alpha;
    }
    out[MP_eta] = k.eta; out[MP_theta] = k.theta; out[MP_kappa] = k.kappa; out[MP_vmax] = k.vmax;
    out[MP_xj] = k.xj; out[MP_ld] = k.ld; out[MP_nfs] = k.nfs;
    out[MP_oxideCapFactor] = oxideCapFactor; out[MP_tSurfMob] = surfaceMobility*tmp_mob;
    out[MP_xl] = effectiveLength; out[MP_w] = w;
    out[MP_uexp] = k.uexp; out[MP_ucrit] = k.ucrit; out[MP_neff] = k.neff; out[MP_narrow] = k.narrow;
    out[MP_nsub] = k.g_nsub ? k.nsub : 0; out[MP_surfMob] = surfaceMobility;
    out[MP_kv] = k.kv; out[MP_nv] = k.nv; out[MP_nc] = k.nc; out[MP_lamda0] = k.lamda0;
    out[MP_lamda1] = k.lamda1; out[MP_sigma] = k.sigma; out[MP_gamma1] = k.gamma1;
    out[MP_icVDS] = k.icVDS; out[MP_icVGS] = k.icVGS; out[MP_icVBS] = k.icVBS;
}

// ---------------------------------------------------------------------------------------
// independent source value at time t (run-time parameter pool laid out by analysis.cpp)
// ---------------------------------------------------------------------------------------
__device__ inline double dev_src_eval(int fn, const double *p, int nextra, double t)
{
    if (fn == B200_FN_PULSE) {
        const double *cache = p + 8, *offs = p + 10;
        const int plen = (int)p[7];                       // bit pattern: which periods carry a pulse
        const int prst = plen ? (int)offs[nextra] : 0;
        const double *bits = offs + nextra + 1;
        double value = p[0];
        double pw = p[3] + p[5];
        for (int i = -1; i < nextra; i++) {
            // the pulse itself (i = -1), then the additional offsets: pulse(0 V2-V1 TDi ...) added
            double time = t - (i < 0 ? p[2] : offs[i]);
            bool skip = false;
            if (p[6] > 0.0) {
                int pnum = 0;
                if (time > p[6]) {
                    pnum = (int)(time/p[6]);
                    time -= (p[6]*pnum);
                }
                if (plen) {                               // inptran.cc:1338-1354
                    int j = pnum;
                    if (pnum >= plen) j = prst ? prst + (pnum - prst)%(plen - prst) : -1;
                    skip = j < 0 || bits[j] == 0.0;
                }
            }
            if (skip) continue;
            if (time < p[3] || time >= pw + p[4]) {
                if (i < 0) value = p[0];
                if (time > 0 && time < p[3])
                    value += time*cache[0];
            }
            else {
                if (i < 0) value = p[1]; else value += (p[1] - p[0]);
                if (time > pw)
                    value += (time - pw)*cache[1];
            }
        }
        return value;
    }
    if (fn == B200_FN_PWL) {
        // pool: nv (t,v) pairs then slopes; stateless search (td_pwlindex is only a cursor)
        int nvals = nextra;
        const double *co = p, *slope = p + 2*nvals;
        if (nvals == 2 && co[3] == co[1])
            return co[1];
        if (t >= co[2*(nvals-1)])
            return co[2*(nvals-1)+1];
        int i = 0;
        while (i + 1 < nvals) {
            if (t <= co[2*(i+1)])
                break;
            i++;
        }
        if (!i && t <= co[0])
            return co[1];
        return co[2*i+1] + slope[i]*(t - co[2*i]);
    }
    if (fn == B200_FN_SIN) {
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
    if (fn == B200_FN_EXP) {
        // IFexpData::eval_func, inptran.cc:3016-3028 (V1 V2 TD1 TAU1 TD2 TAU2)
        double value = p[0];
        if (t <= p[2])
            return value;
        value += (p[1]-p[0])*(1-exp(-(t-p[2])/p[3]));
        if (t <= p[4])
            return value;
        value += (p[0]-p[1])*(1-exp(-(t-p[4])/p[5]));
        return value;
    }
    if (fn == B200_FN_SFFM) {
        // IFsffmData::eval_func, inptran.cc:3215-3222 (VO VA FC MDI FS)
        double w1 = 2*M_PI*p[2];
        double w2 = 2*M_PI*p[4];
        return p[0] + p[1]*sin(w1*t + p[3]*sin(w2*t));
    }
    if (fn == B200_FN_GPULSE) {
        // IFgpulseData::eval_func, inptran.cc:1734-1827 (V1 V2 TD GPW RPT, extra delays; no bit pattern)
        double V = 0;
        const double rchk = 5.0;
        const double *offs = p + 8;
        if (p[4] > 0.0) {
            const double per = p[4];
            const int plen = (int)p[5], prst = (int)p[6];     // bit pattern: which periods carry a pulse
            const double *bits = offs + nextra;
            for (int c = -1; c < nextra; c++) {
                const double td = c < 0 ? p[2] : offs[c];
                int ifirst = 0, ilast = 0;
                if (!(t < td)) {
                    double del = rchk*p[3];
                    ifirst = (int)rint((t - td - del)/per);
                    ilast = (int)rint((t - td + del)/per);
                }
                for (int i = ifirst; i <= ilast; i++) {
                    if (plen) {
                        // inptran.cc:1755-1772 (a negative index reads no set bit)
                        int j = i;
                        if (i >= plen) j = prst ? prst + (i - prst)%(plen - prst) : -1;
                        if (j < 0 || bits[j] == 0.0) continue;
                    }
                    double time = t - i*per;
                    double a = (time - td)/p[3];
                    V += exp(-a*a);
                }
            }
        }
        else {
            for (int c = -1; c < nextra; c++) {
                double a = (t - (c < 0 ? p[2] : offs[c]))/p[3];
                if (fabs(a) < rchk)
                    V += exp(-a*a);
            }
        }
        return p[0] + (p[1] - p[0])*V;
    }
    if (fn == B200_FN_SPULSE) {
        // IFspulseData::eval_func, inptran.cc:2812-2824 (V1 V2 SPER SDEL THETA)
        double time = t - p[3];
        if (time <= 0)
            return p[0];
        if (p[4] != 0.0)
            return p[0] + (p[1]-p[0])*( 1 - cos(2*M_PI*time/p[2])*exp(-time*p[4]) )/2;
        return p[0] + (p[1]-p[0])*( 1 - cos(2*M_PI*time/p[2]) )/2;
    }
    if (fn == B200_FN_AM) {
        // IFamData::eval_func, inptran.cc:3391-3399 (SA OC MF CF DL)
        if (t <= p[4])
            return 0.0;
        double w = 2*M_PI*(t - p[4]);
        return p[0]*(p[1] + sin(w*p[2]))*sin(w*p[3]);
    }
    return 0.0;
}
