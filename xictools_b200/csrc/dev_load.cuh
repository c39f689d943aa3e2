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
    double delta, ag[7], pred[2];
    double trapratio;
    double *sp[8];          // state planes for this trial: S(k,slot) = sp[k][slot*ldt]
    int ldt;
    int *trapbad;           // per-trial flag
};

__device__ __forceinline__ double &ST(const TrialCtx &c, int k, int slot)
{
    return c.sp[k][(size_t)slot*c.ldt];
}

__device__ inline double dev_NIceq(const TrialCtx &c, int q)
{
    double cq = 0;
    if (c.method == 1) {
        switch (c.order) {
        case 6: cq += c.ag[6]*ST(c, 6, q);
        case 5: cq += c.ag[5]*ST(c, 5, q);
        case 4: cq += c.ag[4]*ST(c, 4, q);
        case 3: cq += c.ag[3]*ST(c, 3, q);
        default:
        case 2: cq += c.ag[2]*ST(c, 2, q);
        case 1: cq += c.ag[1]*ST(c, 1, q); break;
        }
    }
    else {
        if (c.order == 1)
            cq = c.ag[1]*ST(c, 1, q);
        else
            cq = - ST(c, 1, q+1)*c.ag[1] - c.ag[0]*ST(c, 1, q);
    }
    return cq;
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

__device__ inline void dev_integrate4(const TrialCtx &c, double *geq, double *ceq, double cap, int q)
{
    double cq = dev_find_ceq(c, q);
    ST(c, 0, q+1) = c.ag[0]*ST(c, 0, q) + cq;
    *ceq = cq;
    *geq = c.ag[0]*cap;
    dev_trapcheck(c, q);
}

// Same as dev_integrate4 with state1[q], state1[q+1] already in registers: the state-1 block
// of a device is fetched with one burst of independent loads at the top of the load function,
// so the five charge integrations do not serialise five DRAM round trips.
__device__ inline void dev_integrate4v(const TrialCtx &c, double *geq, double *ceq, double cap, int q,
    double s1q, double s1c)
{
    double cq;
    if (c.method == 0)
        cq = (c.order <= 1 ? s1q*c.ag[1] : - s1c*c.ag[1] - s1q*c.ag[0]);
    else
        cq = dev_NIceq(c, q);
    ST(c, 0, q+1) = c.ag[0]*ST(c, 0, q) + cq;
    *ceq = cq;
    *geq = c.ag[0]*cap;
    dev_trapcheck(c, q);
}

__device__ inline void dev_integrate2(const TrialCtx &c, int q, double cq)
{
    ST(c, 0, q+1) = c.ag[0]*ST(c, 0, q) + cq;
    dev_trapcheck(c, q);
}

// ---- limiters: devlib/devlib.cc:475-620 ----
__device__ inline double dev_limvds(double vnew, double vold)
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

__device__ inline double dev_pnjlim(double vnew, double vold, double vt, double vcrit, int *icheck)
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

__device__ inline double dev_fetlim(double vnew, double vold, double vto)
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
struct MosInst {            // registers of one (instance, trial)
    int type, off, st;
    int d, g, s, b, dp, sp;
    // derived parameters
    double tPhi, tVbi, tVto, gamma, lambda, beta, oxideCap;
    double tSourceSatCur, tDrainSatCur, sourceVcrit, drainVcrit;
    double Cbd, Cbdsw, Cbs, Cbssw, tBulkPot, tDepCap, f2d, f3d, f4d, f2s, f3s, f4s;
    double gsOvl, gdOvl, gbOvl, dcon, scon, mult, vt, mj, mjsw, icVDS, icVGS, icVBS;
    // cached op point
    int mode;
    double von, vdsat, cd, cbs, cbd, gm, gds, gmbs, gbd, gbs, capbd, capbs;
    // state-1 block and the old state-0 values, prefetched (see dev_mos_prefetch)
    double s1[MS_NSTATES];
    double o_vbs, o_vgs, o_vds, o_vbd, o_capgs, o_capgd, o_capgb;
};

struct MosStuff {
    double vt, vbs, vgs, vds, vbd, vgd, vgb, von, vdsat, cdrain;
    double gcgs, ceqgs, gcgd, ceqgd, gcgb, ceqgb, capgs, capgd, capgb;
};

struct MosOut {             // stamp values in load order
    double m[MT_NSTAMP];
    double r[MRH_NRHS];
};

__device__ inline void dev_sargs(double arg, double bot, double side, double *sarg, double *sargsw)
{
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

__device__ inline void dev_qmeyer(double vgs, double vgd, double von, double vdsat,
    double *capgs, double *capgd, double *capgb, double phi, double cox)
{
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

__device__ inline double dev_mos_eq1(MosInst &m, MosStuff &ms)
{
    double vbx = ((m.mode == 1) ? ms.vbs : ms.vbd);
    double vgx = ((m.mode == 1) ? ms.vgs : ms.vgd);
    double sarg;
    if (vbx <= 0)
        sarg = sqrt(m.tPhi - vbx);
    else {
        sarg = sqrt(m.tPhi);
        sarg = sarg - vbx/(sarg+sarg);
        sarg = DMAX(0, sarg);
    }
    if (m.type > 0)
        ms.von = m.tVbi + m.gamma*sarg;
    else
        ms.von = -m.tVbi + m.gamma*sarg;
    double vgst = vgx - ms.von;
    ms.vdsat = DMAX(vgst, 0);
    double cdrain;
    if (vgst <= 0) {
        cdrain = 0; m.gm = 0; m.gds = 0; m.gmbs = 0;
    }
    else {
        double vds = (m.mode > 0) ? ms.vds : -ms.vds;
        double arg;
        if (sarg <= 0)
            arg = 0;
        else
            arg = m.gamma/(sarg+sarg);
        double betap = m.beta*(1 + m.lambda*vds);
        if (vgst <= vds) {
            m.gm = betap*vgst;
            vgst *= vgst*.5;
            cdrain = betap*vgst;
            m.gds = m.lambda*m.beta*vgst;
            m.gmbs = m.gm*arg;
        }
        else {
            m.gm = betap*vds;
            cdrain = m.gm*(vgst - .5*vds);
            m.gds = betap*(vgst - vds) + m.lambda*m.beta*vds*(vgst - .5*vds);
            m.gmbs = m.gm*arg;
        }
    }
    return cdrain;
}

__device__ inline int dev_mos_limiting(const TrialCtx &c, MosInst &m, MosStuff &ms, int oldlimit)
{
    double s0vds = m.o_vds, s0vgs = m.o_vgs;
    if (s0vds >= 0) {
        ms.vgs = dev_fetlim(ms.vgs, s0vgs, ms.von);
        ms.vds = ms.vgs - ms.vgd;
        ms.vds = dev_limvds(ms.vds, s0vds);
        ms.vgd = ms.vgs - ms.vds;
    }
    else {
        double vgdo = s0vgs - s0vds;
        ms.vgd = dev_fetlim(ms.vgd, vgdo, ms.von);
        ms.vds = ms.vgs - ms.vgd;
        if (!oldlimit)
            ms.vds = -dev_limvds(-ms.vds, -(s0vds));
        ms.vgs = ms.vgd + ms.vds;
    }
    int check;
    if (ms.vds >= 0) {
        ms.vbs = dev_pnjlim(ms.vbs, m.o_vbs, ms.vt, m.sourceVcrit, &check);
        ms.vbd = ms.vbs - ms.vds;
    }
    else {
        ms.vbd = dev_pnjlim(ms.vbd, m.o_vbd, ms.vt, m.drainVcrit, &check);
        ms.vbs = ms.vbd + ms.vds;
    }
    ms.vgb = ms.vgs - ms.vbs;
    return check;
}

__device__ inline void dev_mos_iv(MosInst &m, MosStuff &ms, double gmin)
{
    ms.gcgs = 0; ms.ceqgs = 0; ms.gcgd = 0; ms.ceqgd = 0; ms.gcgb = 0; ms.ceqgb = 0;
    if (ms.vbs <= -3*ms.vt) {
        m.gbs = gmin;
        m.cbs = m.gbs*ms.vbs - m.tSourceSatCur;
    }
    else {
        double evbs = exp(ms.vbs/ms.vt);
        m.gbs = m.tSourceSatCur*evbs/ms.vt + gmin;
        m.cbs = m.tSourceSatCur*(evbs-1) + gmin*ms.vbs;
    }
    if (ms.vbd <= -3*ms.vt) {
        m.gbd = gmin;
        m.cbd = m.gbd*ms.vbd - m.tDrainSatCur;
    }
    else {
        double evbd = exp(ms.vbd/ms.vt);
        m.gbd = m.tDrainSatCur*evbd/ms.vt + gmin;
        m.cbd = m.tDrainSatCur*(evbd-1) + gmin*ms.vbd;
    }
    m.mode = (ms.vds >= 0) ? 1 : -1;
    ms.cdrain = dev_mos_eq1(m, ms);
    if (m.type > 0) {
        m.von = ms.von; m.vdsat = ms.vdsat;
    }
    else {
        m.von = -ms.von; m.vdsat = -ms.vdsat;
    }
    if (m.mode > 0)
        m.cd = ms.cdrain - m.cbd;
    else
        m.cd = -ms.cdrain - m.cbd;
}

__device__ inline void dev_mos_cap(const TrialCtx &c, MosInst &m, MosStuff &ms)
{
    const int st = m.st;
    if (m.Cbs != 0 || m.Cbssw != 0) {
        if (ms.vbs < m.tDepCap) {
            double arg = 1 - ms.vbs/m.tBulkPot, sarg, sargsw;
            dev_sargs(arg, m.mj, m.mjsw, &sarg, &sargsw);
            ST(c, 0, st + MS_qbs) = m.tBulkPot*(m.Cbs*(1 - arg*sarg)/(1 - m.mj)
                + m.Cbssw*(1 - arg*sargsw)/(1 - m.mjsw));
            m.capbs = m.Cbs*sarg + m.Cbssw*sargsw;
        }
        else {
            ST(c, 0, st + MS_qbs) = m.f4s + ms.vbs*(m.f2s + .5*ms.vbs*m.f3s);
            m.capbs = m.f2s + m.f3s*ms.vbs;
        }
    }
    else {
        ST(c, 0, st + MS_qbs) = 0; m.capbs = 0;
    }
    if (m.Cbd != 0 || m.Cbdsw != 0) {
        if (ms.vbd < m.tDepCap) {
            double arg = 1 - ms.vbd/m.tBulkPot, sarg, sargsw;
            dev_sargs(arg, m.mj, m.mjsw, &sarg, &sargsw);
            ST(c, 0, st + MS_qbd) = m.tBulkPot*(m.Cbd*(1 - arg*sarg)/(1 - m.mj)
                + m.Cbdsw*(1 - arg*sargsw)/(1 - m.mjsw));
            m.capbd = m.Cbd*sarg + m.Cbdsw*sargsw;
        }
        else {
            ST(c, 0, st + MS_qbd) = m.f4d + ms.vbd*(m.f2d + .5*ms.vbd*m.f3d);
            m.capbd = m.f2d + ms.vbd*m.f3d;
        }
    }
    else {
        ST(c, 0, st + MS_qbd) = 0; m.capbd = 0;
    }

    double vgx, vgy, cgx, cgy, cgb;
    if (m.mode > 0) { vgx = ms.vgs; vgy = ms.vgd; }
    else { vgx = ms.vgd; vgy = ms.vgs; }
    dev_qmeyer(vgx, vgy, ms.von, ms.vdsat, &cgx, &cgy, &cgb, m.tPhi, m.oxideCap);
    double s0capgs, s0capgd;
    if (m.mode > 0) { s0capgs = cgx; s0capgd = cgy; }
    else { s0capgd = cgx; s0capgs = cgy; }
    ST(c, 0, st + MS_capgs) = s0capgs;
    ST(c, 0, st + MS_capgd) = s0capgd;
    ST(c, 0, st + MS_capgb) = cgb;

    if (c.mode & (MODEINITTRAN|MODEINITSMSIG)) {
        ST(c, 1, st + MS_capgs) = s0capgs;
        ST(c, 1, st + MS_capgd) = s0capgd;
        ST(c, 1, st + MS_capgb) = cgb;
        m.s1[MS_capgs] = s0capgs; m.s1[MS_capgd] = s0capgd; m.s1[MS_capgb] = cgb;
        ms.capgs = s0capgs + s0capgs + m.gsOvl;
        ms.capgd = s0capgd + s0capgd + m.gdOvl;
        ms.capgb = cgb + cgb + m.gbOvl;
        double qgs = ms.capgs*ms.vgs, qgd = ms.capgd*ms.vgd, qgb = ms.capgb*ms.vgb;
        ST(c, 0, st + MS_qgs) = qgs;
        ST(c, 0, st + MS_qgd) = qgd;
        ST(c, 0, st + MS_qgb) = qgb;
        double qbd0 = ST(c, 0, st + MS_qbd), qbs0 = ST(c, 0, st + MS_qbs);
        ST(c, 1, st + MS_qbd) = qbd0;
        ST(c, 1, st + MS_qbs) = qbs0;
        ST(c, 1, st + MS_qgs) = qgs;
        ST(c, 1, st + MS_qgd) = qgd;
        ST(c, 1, st + MS_qgb) = qgb;
        m.s1[MS_qbd] = qbd0; m.s1[MS_qbs] = qbs0;
        m.s1[MS_qgs] = qgs; m.s1[MS_qgd] = qgd; m.s1[MS_qgb] = qgb;
    }
    else {
        ms.capgs = s0capgs + m.s1[MS_capgs] + m.gsOvl;
        ms.capgd = s0capgd + m.s1[MS_capgd] + m.gdOvl;
        ms.capgb = cgb + m.s1[MS_capgb] + m.gbOvl;
        vgx = m.s1[MS_vgs];
        vgy = vgx - m.s1[MS_vds];
        double vxy = vgx - m.s1[MS_vbs];
        ST(c, 0, st + MS_qgs) = (ms.vgs - vgx)*ms.capgs + m.s1[MS_qgs];
        ST(c, 0, st + MS_qgd) = (ms.vgd - vgy)*ms.capgd + m.s1[MS_qgd];
        ST(c, 0, st + MS_qgb) = (ms.vgb - vxy)*ms.capgb + m.s1[MS_qgb];
    }
}

__device__ inline void dev_mos_meyer_integ(const TrialCtx &c, MosInst &m, MosStuff &ms)
{
    const int st = m.st;
    dev_integrate4v(c, &ms.gcgs, &ms.ceqgs, ms.capgs, st + MS_qgs, m.s1[MS_qgs], m.s1[MS_cqgs]);
    dev_integrate4v(c, &ms.gcgd, &ms.ceqgd, ms.capgd, st + MS_qgd, m.s1[MS_qgd], m.s1[MS_cqgd]);
    dev_integrate4v(c, &ms.gcgb, &ms.ceqgb, ms.capgb, st + MS_qgb, m.s1[MS_qgb], m.s1[MS_cqgb]);
    ms.ceqgs += c.ag[0]*ST(c, 0, st + MS_qgs) - ms.gcgs*ms.vgs;
    ms.ceqgd += c.ag[0]*ST(c, 0, st + MS_qgd) - ms.gcgd*ms.vgd;
    ms.ceqgb += c.ag[0]*ST(c, 0, st + MS_qgb) - ms.gcgb*ms.vgb;
}

__device__ inline void dev_mos_integ(const TrialCtx &c, MosInst &m, MosStuff &ms)
{
    const int st = m.st;
    double geq, ceq;
    dev_integrate4v(c, &geq, &ceq, m.capbd, st + MS_qbd, m.s1[MS_qbd], m.s1[MS_cqbd]);
    m.gbd += geq;
    double cqbd = ST(c, 0, st + MS_cqbd);
    m.cbd += cqbd;
    m.cd  -= cqbd;
    dev_integrate4v(c, &geq, &ceq, m.capbs, st + MS_qbs, m.s1[MS_qbs], m.s1[MS_cqbs]);
    m.gbs += geq;
    m.cbs += ST(c, 0, st + MS_cqbs);
    if (ms.capgs == 0) ST(c, 0, st + MS_cqgs) = 0;
    if (ms.capgd == 0) ST(c, 0, st + MS_cqgd) = 0;
    if (ms.capgb == 0) ST(c, 0, st + MS_cqgb) = 0;
    dev_mos_meyer_integ(c, m, ms);
}

// Unified stamp (tran form); with the Meyer terms zero it reproduces load_dc exactly
// (x+0 == x, -(x+0) == -x), see DESIGN.md section 4.
__device__ inline void dev_mos_stamp(const TrialCtx &c, MosInst &m, MosStuff &ms, MosOut &o)
{
    const int st = m.st;
#define MSC(xx) (m.mult*(xx))
    ST(c, 0, st + MS_vbs) = ms.vbs;
    ST(c, 0, st + MS_vbd) = ms.vbd;
    ST(c, 0, st + MS_vgs) = ms.vgs;
    ST(c, 0, st + MS_vds) = ms.vds;
    double gbd = MSC(m.gbd), gbs = MSC(m.gbs), gds = MSC(m.gds), gm = MSC(m.gm);
    double gmbs = MSC(m.gmbs), cbd = MSC(m.cbd), cbs = MSC(m.cbs);
    double ceqbs = cbs - gbs*ms.vbs;
    double ceqbd = cbd - gbd*ms.vbd;
    double cdreq;
    if (m.mode > 0)
        cdreq = MSC(ms.cdrain) - gds*ms.vds - gm*ms.vgs - gmbs*ms.vbs;
    else
        cdreq = -(MSC(ms.cdrain) + gds*ms.vds - gm*ms.vgd - gmbs*ms.vbd);
    double dcon = MSC(m.dcon), scon = MSC(m.scon);
    if (m.type > 0) {
        o.r[MRH_g]  = -MSC(ms.ceqgs + ms.ceqgb + ms.ceqgd);
        o.r[MRH_b]  = -(ceqbs + ceqbd - MSC(ms.ceqgb));
        o.r[MRH_dp] = ceqbd - cdreq + MSC(ms.ceqgd);
        o.r[MRH_sp] = cdreq + ceqbs + MSC(ms.ceqgs);
    }
    else {
        o.r[MRH_g]  = MSC(ms.ceqgs + ms.ceqgb + ms.ceqgd);
        o.r[MRH_b]  = ceqbs + ceqbd - MSC(ms.ceqgb);
        o.r[MRH_dp] = -(ceqbd - cdreq + MSC(ms.ceqgd));
        o.r[MRH_sp] = -(cdreq + ceqbs + MSC(ms.ceqgs));
    }
    double gcgd = MSC(ms.gcgd), gcgs = MSC(ms.gcgs), gcgb = MSC(ms.gcgb);
    o.m[MT_Dd] = dcon;
    o.m[MT_Gg] = gcgd + gcgs + gcgb;
    o.m[MT_Ss] = scon;
    o.m[MT_Bb] = gbd + gbs + gcgb;
    o.m[MT_Ddp] = -dcon;
    o.m[MT_Gb] = -gcgb;
    o.m[MT_Gdp] = -gcgd;
    o.m[MT_Gsp] = -gcgs;
    o.m[MT_Ssp] = -scon;
    o.m[MT_Bg] = -gcgb;
    o.m[MT_Bdp] = -gbd;
    o.m[MT_Bsp] = -gbs;
    o.m[MT_DPd] = -dcon;
    o.m[MT_SPs] = -scon;
    if (m.mode > 0) {
        o.m[MT_DPdp] = dcon + gds + gbd + gcgd;
        o.m[MT_SPsp] = scon + gds + gbs + gm + gmbs + gcgs;
        o.m[MT_DPg] = gm - gcgd;
        o.m[MT_DPb] = -gbd + gmbs;
        o.m[MT_DPsp] = -(gds + gm + gmbs);
        o.m[MT_SPg] = -(gm + gcgs);
        o.m[MT_SPb] = -(gbs + gmbs);
        o.m[MT_SPdp] = -gds;
    }
    else {
        o.m[MT_DPdp] = dcon + gds + gbd + gm + gmbs + gcgd;
        o.m[MT_SPsp] = scon + gds + gbs + gcgs;
        o.m[MT_DPg] = -(gm + gcgd);
        o.m[MT_DPb] = -(gbd + gmbs);
        o.m[MT_DPsp] = -gds;
        o.m[MT_SPg] = -(-gm + gcgs);
        o.m[MT_SPb] = -(gbs - gmbs);
        o.m[MT_SPdp] = -(gds + gm + gmbs);
    }
#undef MSC
}

__device__ inline int dev_mos_bypass(const TrialCtx &c, MosInst &m, MosStuff &ms, MosOut &o,
    double reltol, double abstol, double vntol)
{
    const int st = m.st;
    double s0vbs = m.o_vbs, s0vgs = m.o_vgs;
    double s0vds = m.o_vds, s0vbd = m.o_vbd;
    double delvbs = ms.vbs - s0vbs;
    double delvgs = ms.vgs - s0vgs;
    double delvds = ms.vds - s0vds;
    double delvbd = ms.vbd - s0vbd;
    double delvgd = ms.vgd - (s0vgs - s0vds);
    double cbhat = m.cbs + m.cbd + m.gbd*delvbd + m.gbs*delvbs;
    double A1 = fabs(cbhat);
    double A2 = m.cbs + m.cbd;
    A2 = fabs(A2);
    double A3 = DMAX(A1, A2) + abstol;
    A1 = cbhat - (m.cbs + m.cbd);
    if (fabs(A1) >= reltol*A3) return 0;
    A1 = fabs(ms.vbs); A2 = fabs(s0vbs);
    if (fabs(delvbs) >= (reltol*DMAX(A1, A2) + vntol)) return 0;
    A1 = fabs(ms.vbd); A2 = fabs(s0vbd);
    if (fabs(delvbd) >= (reltol*DMAX(A1, A2) + vntol)) return 0;
    A1 = fabs(ms.vgs); A2 = fabs(s0vgs);
    if (fabs(delvgs) >= (reltol*DMAX(A1, A2) + vntol)) return 0;
    A1 = fabs(ms.vds); A2 = fabs(s0vds);
    if (fabs(delvds) >= (reltol*DMAX(A1, A2) + vntol)) return 0;
    double cdhat;
    if (m.mode >= 0)
        cdhat = m.cd - m.gbd*delvbd + m.gmbs*delvbs + m.gm*delvgs + m.gds*delvds;
    else
        cdhat = m.cd + (m.gmbs - m.gbd)*delvbd - m.gm*delvgd + m.gds*delvds;
    A1 = cdhat - m.cd;
    A2 = fabs(cdhat);
    A3 = fabs(m.cd);
    if ((fabs(A1) >= reltol*DMAX(A2, A3) + abstol)) return 0;

    ms.vbs = s0vbs;
    ms.vgs = s0vgs;
    ms.vds = s0vds;
    ms.vbd = s0vbd;
    ms.vgd = ms.vgs - ms.vds;
    ms.vgb = ms.vgs - ms.vbs;
    ms.cdrain = m.mode*(m.cd + m.cbd);
    if (c.mode & MODETRAN) {
        ms.capgs = m.o_capgs + m.s1[MS_capgs] + m.gsOvl;
        ms.capgd = m.o_capgd + m.s1[MS_capgd] + m.gdOvl;
        ms.capgb = m.o_capgb + m.s1[MS_capgb] + m.gbOvl;
        dev_mos_meyer_integ(c, m, ms);
    }
    dev_mos_stamp(c, m, ms, o);
    return 1;
}

__device__ inline void dev_mos_get_v(const MosInst &m, MosStuff &ms, const double *v, int ldt)
{
    double vb = v[(size_t)m.b*ldt], vsp = v[(size_t)m.sp*ldt];
    double vg = v[(size_t)m.g*ldt], vdp = v[(size_t)m.dp*ldt];
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

// Returns: bit0 = stamps valid (write contributions), bit1 = cache dirty, bit2 = incNoncon
__device__ inline int dev_mos_load(const TrialCtx &c, MosInst &m, MosOut &o,
    const double *rhsOld /* + trial */, const b200_options &opt)
{
    MosStuff ms;
    ms.vt = m.vt;
    ms.gcgs = ms.ceqgs = ms.gcgd = ms.ceqgd = ms.gcgb = ms.ceqgb = 0;
    ms.capgs = ms.capgd = ms.capgb = 0; ms.cdrain = 0; ms.vdsat = 0;
    const int st = m.st;
    int ret = 0;
    // one burst of independent loads: old state-0 values (limiting / bypass) and, for the
    // transient modes, the whole state-1 block of this device
    m.o_vbs = ST(c, 0, st + MS_vbs); m.o_vgs = ST(c, 0, st + MS_vgs);
    m.o_vds = ST(c, 0, st + MS_vds); m.o_vbd = ST(c, 0, st + MS_vbd);
    if (c.mode & (MODETRAN | MODEINITPRED)) {
        m.o_capgs = ST(c, 0, st + MS_capgs); m.o_capgd = ST(c, 0, st + MS_capgd);
        m.o_capgb = ST(c, 0, st + MS_capgb);
#pragma unroll
        for (int k = 0; k < MS_NSTATES; k++) m.s1[k] = ST(c, 1, st + k);
    }
    if (c.mode & MODEINITFLOAT) {
        dev_mos_get_v(m, ms, rhsOld, c.ldt);
        ms.von = (m.type > 0) ? m.von : -m.von;
        ms.vbd = ms.vbs - ms.vds;
        ms.vgd = ms.vgs - ms.vds;
        ms.vgb = ms.vgs - ms.vbs;
        if (opt.bypass && dev_mos_bypass(c, m, ms, o, opt.reltol, opt.abstol, opt.vntol))
            return 1;
        dev_mos_limiting(c, m, ms, opt.oldlimit);
        dev_mos_iv(m, ms, opt.gmin);
        if (c.mode & MODETRAN) {
            dev_mos_cap(c, m, ms);
            dev_mos_integ(c, m, ms);
        }
        dev_mos_stamp(c, m, ms, o);
        return 3;
    }
    if (c.mode & MODEINITPRED) {
        double s1vbs = m.s1[MS_vbs], s1vgs = m.s1[MS_vgs], s1vds = m.s1[MS_vds];
        ms.vbs = s1vbs*c.pred[0] + ST(c, 2, st + MS_vbs)*c.pred[1];
        ms.vgs = s1vgs*c.pred[0] + ST(c, 2, st + MS_vgs)*c.pred[1];
        ms.vds = s1vds*c.pred[0] + ST(c, 2, st + MS_vds)*c.pred[1];
        ms.vbd = ms.vbs - ms.vds;
        ms.vgd = ms.vgs - ms.vds;
        ms.vgb = ms.vgs - ms.vbs;
        ms.von = (m.type > 0) ? m.von : -m.von;
        ST(c, 0, st + MS_vbs) = s1vbs;
        ST(c, 0, st + MS_vgs) = s1vgs;
        ST(c, 0, st + MS_vds) = s1vds;
        ST(c, 0, st + MS_vbd) = s1vbs - s1vds;
        // limiting reads the state-0 voltages just written (mosload.cc:172-182)
        m.o_vbs = s1vbs; m.o_vgs = s1vgs; m.o_vds = s1vds; m.o_vbd = s1vbs - s1vds;
        dev_mos_limiting(c, m, ms, opt.oldlimit);
        dev_mos_iv(m, ms, opt.gmin);
        if (c.mode & MODETRAN) {
            dev_mos_cap(c, m, ms);
            dev_mos_integ(c, m, ms);
        }
        dev_mos_stamp(c, m, ms, o);
        return 3;
    }
    if (c.mode & MODEINITFIX) {
        ms.von = (m.type > 0) ? m.von : -m.von;
        if (m.off) {
            ms.vbs = ms.vgs = ms.vds = ms.vbd = ms.vgd = ms.vgb = 0;
        }
        else {
            dev_mos_get_v(m, ms, rhsOld, c.ldt);
            ms.vbd = ms.vbs - ms.vds;
            ms.vgd = ms.vgs - ms.vds;
            ms.vgb = ms.vgs - ms.vbs;
            if (dev_mos_limiting(c, m, ms, opt.oldlimit))
                ret |= 4;
        }
        dev_mos_iv(m, ms, opt.gmin);
        dev_mos_stamp(c, m, ms, o);
        return ret | 3;
    }
    if (c.mode & MODEINITTRAN) {
        ms.von = (m.type > 0) ? m.von : -m.von;
        ms.vbs = m.s1[MS_vbs];
        ms.vgs = m.s1[MS_vgs];
        ms.vds = m.s1[MS_vds];
        ms.vbd = ms.vbs - ms.vds;
        ms.vgd = ms.vgs - ms.vds;
        ms.vgb = ms.vgs - ms.vbs;
        dev_mos_iv(m, ms, opt.gmin);
        dev_mos_cap(c, m, ms);
        dev_mos_integ(c, m, ms);
        dev_mos_stamp(c, m, ms, o);
        return 3;
    }
    if ((c.mode & MODEINITJCT) && (c.mode & MODETRANOP) && (c.mode & MODEUIC)) {
        if (m.type > 0) {
            ms.vds = m.icVDS; ms.vgs = m.icVGS; ms.vbs = m.icVBS;
        }
        else {
            ms.vds = -m.icVDS; ms.vgs = -m.icVGS; ms.vbs = -m.icVBS;
        }
        ST(c, 0, st + MS_vds) = ms.vds;
        ST(c, 0, st + MS_vgs) = ms.vgs;
        ST(c, 0, st + MS_vbs) = ms.vbs;
        return 0;
    }
    if (c.mode & MODEINITJCT) {
        ms.von = (m.type > 0) ? m.von : -m.von;
        if (m.off) {
            ms.vbs = ms.vgs = ms.vds = ms.vbd = ms.vgd = ms.vgb = 0;
        }
        else {
            if (m.icVDS == 0 && m.icVGS == 0 && m.icVBS == 0) {
                ms.vbs = -1;
                ms.vgs = (m.type > 0) ? m.tVto : -m.tVto;
                ms.vds = 0;
            }
            else {
                if (m.type > 0) {
                    ms.vds = m.icVDS; ms.vgs = m.icVGS; ms.vbs = m.icVBS;
                }
                else {
                    ms.vds = -m.icVDS; ms.vgs = -m.icVGS; ms.vbs = -m.icVBS;
                }
            }
            ms.vbd = ms.vbs - ms.vds;
            ms.vgd = ms.vgs - ms.vds;
            ms.vgb = ms.vgs - ms.vbs;
        }
        dev_mos_iv(m, ms, opt.gmin);
        dev_mos_stamp(c, m, ms, o);
        return 3;
    }
    return 0;
}

// MOSdev::convTest for one instance (mosconv.cc:52-141).  rhs = CKTrhs (PREVIOUS iterate).
__device__ inline int dev_mos_conv(const TrialCtx &c, const MosInst &m, const double *rhs,
    double reltol, double abstol)
{
    if (m.off && (c.mode & MODEINITFIX))
        return 0;
    MosStuff ms;
    dev_mos_get_v(m, ms, rhs, c.ldt);
    const int st = m.st;
    double s0vbs = ST(c, 0, st + MS_vbs), s0vgs = ST(c, 0, st + MS_vgs);
    double s0vds = ST(c, 0, st + MS_vds), s0vbd = ST(c, 0, st + MS_vbd);
    double vbd = ms.vbs - ms.vds, vgd = ms.vgs - ms.vds;
    double delvbs = ms.vbs - s0vbs;
    double delvbd = vbd - s0vbd;
    double delvgs = ms.vgs - s0vgs;
    double delvds = ms.vds - s0vds;
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
            kp = surfaceMobility*oxideCapFactor*M2PCM2;
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
        kp = 2e-5;

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
    out[MP_icVDS] = k.icVDS; out[MP_icVGS] = k.icVGS; out[MP_icVBS] = k.icVBS;
}

// ---------------------------------------------------------------------------------------
// independent source value at time t (run-time parameter pool laid out by analysis.cpp)
// ---------------------------------------------------------------------------------------
__device__ inline double dev_src_eval(int fn, const double *p, int nextra, double t)
{
    if (fn == B200_FN_PULSE) {
        const double *cache = p + 8, *offs = p + 10;
        double value = p[0];
        double pw = p[3] + p[5];
        double time = t - p[2];
        if (p[6] > 0.0) {
            if (time > p[6]) {
                int pnum = (int)(time/p[6]);
                time -= (p[6]*pnum);
            }
        }
        if (time < p[3] || time >= pw + p[4]) {
            value = p[0];
            if (time > 0 && time < p[3])
                value += time*cache[0];
        }
        else {
            value = p[1];
            if (time > pw)
                value += (time - pw)*cache[1];
        }
        for (int i = 0; i < nextra; i++) {
            time = t - offs[i];
            if (p[6] > 0.0) {
                if (time > p[6]) {
                    int pnum = (int)(time/p[6]);
                    time -= (p[6]*pnum);
                }
            }
            if (time < p[3] || time >= pw + p[4]) {
                if (time > 0 && time < p[3])
                    value += time*cache[0];
            }
            else {
                value += (p[1] - p[0]);
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
    return 0.0;
}
