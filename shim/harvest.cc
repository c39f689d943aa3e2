// shim/harvest.cc -- the reference-side half of the drop-in boundary: walk a set-up sCKT and
// hand its topology, cards and options to the batched engine through the C ABI
// (include/b200spice.h).  This is the code a WRspice maintainer adds next to
// sCHECKprms::loop (wrspice/src/fte/check.cc:1472-1733) / sDCTprms::loop_mt
// (src/analysis/dctprms.cc:924-1011); it is compiled here AGAINST THE REFERENCE'S OWN HEADERS
// (nothing of the reference is copied) and linked into oracle/_ref/libwrsref*.so by
// oracle/refbuild/Makefile so the tests can drive it: parser -> sCKT -> harvest ->
// b200_batch_create/run -> counters and waveforms, compared with the same sCKT's doTask.
//
// What is read, and where the reference defines it:
//   models / instances    sCKTmodGen over ckt->CKTmodels (circuit.h:1560-2140), instance lists
//                         GENinstances / GENnextInstance (circuit.h:154-201); the walk order IS
//                         sCKT::load's order (ckt.cc:991-1005 mutuals first, :1146-1161 the rest)
//   R / C / L / K / V / I devlib/{res,cap,ind,src}/*defs.h instance PODs, "Given" bits -> NaN
//   MOS levels 1/2/3/6    devlib/mos/mosdefs.h:115-328 (instance), :415-600 (model)
//   JJ level 1            devlib/jj/jjdefs.h:80-235 (instance), :270-335 (model, post-setup values)
//   source functions      IFparseNode::tranData() (inpptree.h:262), IFtranData (inptran.h:62-95)
//   nodes, .ic, .nodeset  sCKTnodeTab (circuit.h:1386-1430)
//   options               sTASK TSK* (circuit.h:999-1069), .tran job sTRANAN (trandefs.h:148-190)
// The engine library is bound with dlopen (the KLU plug-in is bound the same way, kluif.cc:117).
#include "config.h"
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "circuit.h"
#include "device.h"
// the bit pattern of a PULSE / GPULSE source is a private member (inptran.h:236-276; inpptree.h
// pulls that header in)
#define private public
#include "input.h"
#include "inpptree.h"
#include "inptran.h"
#undef private
#include "trandefs.h"
#include "dcodefs.h"
#include "res/resdefs.h"
#include "cap/capdefs.h"
#include "ind/inddefs.h"
#include "src/srcdefs.h"
#include "mos/mosdefs.h"
#include "jj/jjdefs.h"
#include "tjm/tjmdefs.h"

#include "b200spice.h"
#include "harvest.h"

namespace {

const double NaN = NAN;
inline double opt(bool given, double v) { return given ? v : NaN; }

// IFtranData keeps its card arguments protected; a maintainer would add two accessors.  Here a
// derived view reads them.
struct TranPeek : public IFtranData {
    static const double *coeffs(IFtranData *t) { return static_cast<TranPeek*>(t)->td_coeffs; }
    static int ncoeffs(IFtranData *t) { return static_cast<TranPeek*>(t)->td_numcoeffs; }
    static int type(IFtranData *t) { return static_cast<TranPeek*>(t)->td_type; }
};

struct Api {
    void *h = nullptr;
    void (*options_default)(b200_options*) = nullptr;
    int (*create)(const b200_deck*, int, int, b200_batch**) = nullptr;
    void (*destroy)(b200_batch*) = nullptr;
    int (*set_param)(b200_batch*, int, int, const double*) = nullptr;
    int (*run_tran)(b200_batch*) = nullptr;
    int (*get_stats)(const b200_batch*, b200_run_stats*) = nullptr;
    int (*set_saves)(b200_batch*, int, const int*) = nullptr;
    int (*num_saved)(const b200_batch*) = nullptr;
    int (*run_op)(b200_batch*) = nullptr;
    int (*num_unknowns)(const b200_batch*) = nullptr;
    int (*num_points)(const b200_batch*) = nullptr;
    int (*get_counters)(const b200_batch*, int*) = nullptr;
    int (*get_waves)(const b200_batch*, double*, double*) = nullptr;
    bool load(const char *path) {
        h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (!h) { fprintf(stderr, "b200 shim: %s\n", dlerror()); return false; }
#define SYM(field, name) *(void**)&field = dlsym(h, name); if (!field) { fprintf(stderr, "b200 shim: %s missing\n", name); return false; }
        SYM(options_default, "b200_options_default") SYM(create, "b200_batch_create")
        SYM(destroy, "b200_batch_destroy") SYM(set_param, "b200_batch_set_param")
        SYM(run_tran, "b200_batch_run_tran") SYM(get_stats, "b200_batch_get_stats") SYM(set_saves, "b200_batch_set_saves") SYM(num_saved, "b200_batch_num_saved") SYM(run_op, "b200_batch_run_op")
        SYM(num_unknowns, "b200_batch_num_unknowns") SYM(num_points, "b200_batch_num_points")
        SYM(get_counters, "b200_batch_get_counters") SYM(get_waves, "b200_batch_get_waves")
#undef SYM
        return true;
    }
};

} // namespace

// Everything b200_deck points at lives here.
struct b200_harvest {
    b200_deck deck;
    std::vector<b200_res> res; std::vector<b200_cap> cap; std::vector<b200_ind> ind;
    std::vector<b200_mut> mut; std::vector<b200_src> src;
    std::vector<b200_mos_model> mosmod; std::vector<b200_mos> mos;
    std::vector<b200_jj_model> jjmod; std::vector<b200_jj> jj;
    std::vector<b200_tjm_model> tjmmod; std::vector<b200_tjm> tjm;
    std::vector<std::vector<double>> tjmcoef;
    std::vector<std::vector<double>> coefs;
    std::vector<int> ic_node, ns_node, okind, oindex;
    std::vector<double> ic_val, ns_val;
    std::map<const void*, int> ind_index, mosmod_index, jjmod_index, tjmmod_index;
    std::vector<std::string> names[B200_DEV_NKINDS];      // instance names per kind, table order
    std::string error;
    int has_tran = 0;
};

static void harvest_options(const sTASK *t, b200_options &o)
{
    o.abstol = t->TSKabstol; o.reltol = t->TSKreltol; o.vntol = t->TSKvoltTol; o.chgtol = t->TSKchgtol;
    o.trtol = t->TSKtrtol; o.gmin = t->TSKgmin; o.gmax = t->TSKgmax;
    o.pivrel = t->TSKpivotRelTol; o.pivtol = t->TSKpivotAbsTol;
    o.temp = t->TSKtemp; o.tnom = t->TSKnomTemp;
    o.xmu = t->TSKxmu; o.dcmu = t->TSKdcMu; o.trapratio = t->TSKtrapRatio;
    o.delmin = t->TSKdelMin; o.minbreak = t->TSKminBreak; o.resmin = t->TSKresmin;
    o.defl = t->TSKdefaultMosL; o.defw = t->TSKdefaultMosW;
    o.defad = t->TSKdefaultMosAD; o.defas = t->TSKdefaultMosAS;
    o.bypass = t->TSKbypass; o.maxord = t->TSKmaxOrder;
    o.itl1 = t->TSKdcMaxIter; o.itl2 = t->TSKdcTrcvMaxIter; o.itl4 = t->TSKtranMaxIter;
    o.itl_gmin = t->TSKdcOpGminMaxIter; o.itl_src = t->TSKdcOpSrcMaxIter;
    o.gminsteps = t->TSKnumGminSteps; o.srcsteps = t->TSKnumSrcSteps; o.gminfirst = t->TSKgminFirst;
    o.method = t->TSKintegrateMethod == GEAR ? 1 : 0;         // circuit.h:451-452 TRAPEZOIDAL 1, GEAR 2
    o.trapcheck = t->TSKtrapCheck; o.spice3 = t->TSKspice3; o.oldsteplim = t->TSKoldStepLim;
    o.oldlimit = t->TSKfixLimit; o.forcegmin = t->TSKforceGmin;
    o.interplev = t->TSKinterpLev; o.jjaccel = t->TSKjjaccel; o.nojjtp = t->TSKnojjtp;
    o.noopiter = t->TSKnoOpIter; o.steptype = t->TSKtranStepType;
    if (getenv("B200_SHIM_TRACE")) {
        const double *d = &o.abstol;
        fprintf(stderr, "b200 shim options:");
        for (int i = 0; i < 21; i++) fprintf(stderr, " %g", d[i]);
        const int *k = &o.bypass;
        for (int i = 0; i < 20; i++) fprintf(stderr, " %d", k[i]);
        fprintf(stderr, "\n");
    }
}

extern "C" void b200_harvest_free(b200_harvest *h) { delete h; }

extern "C" const char *b200_harvest_error(const b200_harvest *h) { return h ? h->error.c_str() : "null"; }

extern "C" const b200_deck *b200_harvest_deck(const b200_harvest *h) { return h ? &h->deck : nullptr; }

extern "C" int b200_harvest_find(const b200_harvest *h, int kind, const char *name)
{
    if (!h || kind < 0 || kind >= B200_DEV_NKINDS || !name) return -1;
    for (size_t i = 0; i < h->names[kind].size(); i++)
        if (!strcasecmp(h->names[kind][i].c_str(), name)) return (int)i;
    return -1;
}

// The walk.  `ckt` must have been set up (sCKT::setup via doTask/doTaskSetup): branch and
// internal nodes exist, MUTind1/2 are bound.  Returns 0 (and *out) or a B200_E_* code.
extern "C" int b200_harvest_ckt(sCKT *ckt, b200_harvest **out)
{
    *out = nullptr;
    if (!ckt || !ckt->CKTcurTask) return B200_E_BADPARM;
    b200_harvest *H = new b200_harvest;
    *out = H;
    b200_deck &D = H->deck;
    memset(&D, 0, sizeof(D));
    harvest_options(ckt->CKTcurTask, D.opt);

    // analysis: the first .tran of the task (tranan.cc:95-217)
    for (sJOB *j = ckt->CKTcurTask->TSKjobs; j; j = j->JOBnextJob) {
        sTRANAN *tj = dynamic_cast<sTRANAN*>(j);
        if (!tj || !tj->TRANspec) continue;
        if (tj->TRANspec->nparts != 1) { H->error = "multi-range .tran"; return B200_E_UNSUPP; }
        D.tstep = tj->TRANspec->step(0); D.tstop = tj->TRANspec->end(0);
        D.tstart = tj->TRANspec->tstart; D.tmax = tj->TRANmaxStep;
        D.uic = (tj->TRANmode & MODEUIC) ? 1 : 0;
        // an analysis that already ran has written its derived delmin / minbreak back into the
        // task (tranan.cc:183-212, 662-667); "not given" is what the engine must see
        if (!tj->TS.t_delmin_given) D.opt.delmin = 0.0;
        if (!tj->TS.t_minbreak_given) D.opt.minbreak = 0.0;
        H->has_tran = 1;
        break;
    }

    // external nodes keep the parser's numbers; .ic / .nodeset ride on the node table
    int maxnode = 0;
    std::vector<char> is_ext;
    for (const sCKTnode *n = ckt->CKTnodeTab.find(1); n; n = ckt->CKTnodeTab.nextNode(n)) {
        if (n->icGiven()) { H->ic_node.push_back(n->number()); H->ic_val.push_back(n->ic()); }
        if (n->nsGiven()) { H->ns_node.push_back(n->number()); H->ns_val.push_back(n->nodeset()); }
    }

    int muttype = -1;
    std::vector<sGENmodel*> heads;
    {
        sGENmodel *mutm = nullptr;
        ckt->typelook("mutual", &mutm);
        if (mutm) { muttype = mutm->GENmodType; }
        sCKTmodGen mgen(ckt->CKTmodels);
        // inductors must be numbered before the mutuals that name them
        for (sGENmodel *m = mgen.next(); m; m = mgen.next()) heads.push_back(m);
    }
    auto ext = [&](int node) { if (node > maxnode) maxnode = node; return node; };

    struct Ord { int kind, index; };
    std::vector<Ord> order_rest, order_mut;
    for (int pass = 0; pass < 2; pass++) {
        for (sGENmodel *head : heads) {
            const bool is_mut = (head->GENmodType == muttype);
            if ((pass == 0) == is_mut) continue;              // pass 0: everything but mutuals
            const char *dn = DEV.device(head->GENmodType)->name();
            for (sGENmodel *dm = head; dm; dm = dm->GENnextModel) {
                for (sGENinstance *gi = dm->GENinstances; gi; gi = gi->GENnextInstance) {
                    const char *iname = (const char*)gi->GENname;
                    if (!strcmp(dn, "Resistor")) {
                        sRESinstance *r = (sRESinstance*)gi;
                        if (r->REStree || r->RESpolyCoeffs) { H->error = "nonlinear resistor"; return B200_E_UNSUPP; }
                        b200_res x = { ext(r->RESposNode), ext(r->RESnegNode), r->RESresist,
                                       opt(r->RESmGiven, r->RESm), opt(r->REStc1Given, r->REStc1),
                                       opt(r->REStc2Given, r->REStc2), opt(r->REStempGiven, r->REStemp) };
                        if (!r->RESresGiven) { H->error = "resistor from geometry"; return B200_E_UNSUPP; }
                        order_rest.push_back({B200_DEV_RES, (int)H->res.size()});
                        H->res.push_back(x); H->names[B200_DEV_RES].push_back(iname);
                    }
                    else if (!strcmp(dn, "Capacitor")) {
                        sCAPinstance *c = (sCAPinstance*)gi;
                        if (c->CAPtree || c->CAPpolyCoeffs || !c->CAPcapGiven) { H->error = "nonlinear / geometric capacitor"; return B200_E_UNSUPP; }
                        b200_cap x = { ext(c->CAPposNode), ext(c->CAPnegNode), c->CAPnomCapac,
                                       opt(c->CAPmGiven, c->CAPm), opt(c->CAPicGiven, c->CAPinitCond) };
                        order_rest.push_back({B200_DEV_CAP, (int)H->cap.size()});
                        H->cap.push_back(x); H->names[B200_DEV_CAP].push_back(iname);
                    }
                    else if (!strcmp(dn, "Inductor")) {
                        sINDinstance *l = (sINDinstance*)gi;
                        if (l->INDtree || l->INDpolyCoeffs || l->INDresGiven) { H->error = "nonlinear / resistive inductor"; return B200_E_UNSUPP; }
                        b200_ind x = { ext(l->INDposNode), ext(l->INDnegNode), l->INDnomInduct,
                                       opt(l->INDmGiven, l->INDm), opt(l->INDicGiven, l->INDinitCond) };
                        H->ind_index[l] = (int)H->ind.size();
                        order_rest.push_back({B200_DEV_IND, (int)H->ind.size()});
                        H->ind.push_back(x); H->names[B200_DEV_IND].push_back(iname);
                    }
                    else if (!strcmp(dn, "Mutual")) {
                        sMUTinstance *k = (sMUTinstance*)gi;
                        auto a = H->ind_index.find(k->MUTind1), b = H->ind_index.find(k->MUTind2);
                        if (a == H->ind_index.end() || b == H->ind_index.end()) { H->error = "mutual: unbound inductor"; return B200_E_BADPARM; }
                        b200_mut x = { a->second, b->second, k->MUTcoupling };
                        order_mut.push_back({B200_DEV_MUT, (int)H->mut.size()});
                        H->mut.push_back(x); H->names[B200_DEV_MUT].push_back(iname);
                    }
                    else if (!strcmp(dn, "Source")) {
                        sSRCinstance *s = (sSRCinstance*)gi;
                        if (s->SRCdep) { H->error = "controlled source"; return B200_E_UNSUPP; }
                        b200_src x;
                        memset(&x, 0, sizeof(x));
                        x.kind = s->SRCtype == GENSRC_V ? B200_SRC_V : B200_SRC_I;
                        x.pos = ext(s->SRCposNode); x.neg = ext(s->SRCnegNode);
                        x.dc = opt(s->SRCdcGiven, s->SRCdcValue);
                        x.fn = B200_FN_NONE;
                        if (s->SRCtree) {
                            IFtranData *td = s->SRCtree->tree() ? s->SRCtree->tree()->tranData() : nullptr;
                            if (!td) { H->error = "source expression"; return B200_E_UNSUPP; }
                            switch (TranPeek::type(td)) {
                            case PTF_tPULSE: x.fn = B200_FN_PULSE; break;
                            case PTF_tPWL:   x.fn = B200_FN_PWL; break;
                            case PTF_tSIN:   x.fn = B200_FN_SIN; break;
                            case PTF_tEXP:   x.fn = B200_FN_EXP; break;
                            case PTF_tSFFM:  x.fn = B200_FN_SFFM; break;
                            case PTF_tAM:    x.fn = B200_FN_AM; break;
                            case PTF_tGPULSE: x.fn = B200_FN_GPULSE; break;     // (a bit pattern is private to the class: not seen)
                            case PTF_tSPULSE: x.fn = B200_FN_SPULSE; break;
                            case PTF_tGAUSS: x.fn = B200_FN_GAUSS; break;
                            default: H->error = "source function"; return B200_E_UNSUPP;
                            }
                            int nc = TranPeek::ncoeffs(td);
                            const double *cf = TranPeek::coeffs(td);
                            H->coefs.emplace_back(cf, cf + nc);
                            pbitList *pbl = x.fn == B200_FN_PULSE ? ((IFpulseData*)td)->td_pblist :
                                            x.fn == B200_FN_GPULSE ? ((IFgpulseData*)td)->td_pblist : nullptr;
                            if (pbl) {
                                // expanded as IF[g]pulseData::setup does (inptran.cc:1305-1316,
                                // 1714-1727): the reference's own pbitAry; handed on after a NaN
                                // separator as length, repeat start, bits
                                pbitAry ba;
                                ba.add(pbl);
                                if (ba.count()) {
                                    unsigned long *bits = ba.final();
                                    std::vector<double> &cv = H->coefs.back();
                                    cv.push_back(NAN);
                                    cv.push_back((double)ba.count());
                                    cv.push_back((double)ba.rep_start());
                                    for (int q = 0; q < ba.count(); q++)
                                        cv.push_back((bits[q/sizeof(unsigned long)] >> (q % sizeof(unsigned long))) & 1 ? 1.0 : 0.0);
                                    delete [] bits;
                                }
                            }
                            x.ncoef = (int)H->coefs.back().size();
                        }
                        else
                            H->coefs.emplace_back();
                        order_rest.push_back({B200_DEV_SRC, (int)H->src.size()});
                        H->src.push_back(x); H->names[B200_DEV_SRC].push_back(iname);
                    }
                    else if (!strcmp(dn, "MOS")) {
                        sMOSmodel *m = (sMOSmodel*)dm;
                        sMOSinstance *q = (sMOSinstance*)gi;
                        auto it = H->mosmod_index.find(m);
                        if (it == H->mosmod_index.end()) {
                            b200_mos_model y;
                            memset(&y, 0, sizeof(y));
                            y.type = m->MOStype; y.level = m->MOSlevel;
#define MP(field, flag, val) y.field = opt(m->flag, m->val)
                            MP(vto, MOSvt0Given, MOSvt0); MP(kp, MOStransconductanceGiven, MOStransconductance);
                            MP(gamma, MOSgammaGiven, MOSgamma); MP(phi, MOSphiGiven, MOSphi);
                            MP(lambda, MOSlambdaGiven, MOSlambda);
                            MP(rd, MOSdrainResistanceGiven, MOSdrainResistance);
                            MP(rs, MOSsourceResistanceGiven, MOSsourceResistance);
                            MP(cbd, MOScapBDGiven, MOScapBD); MP(cbs, MOScapBSGiven, MOScapBS);
                            MP(is, MOSjctSatCurGiven, MOSjctSatCur);
                            MP(pb, MOSbulkJctPotentialGiven, MOSbulkJctPotential);
                            MP(cgso, MOSgateSourceOverlapCapFactorGiven, MOSgateSourceOverlapCapFactor);
                            MP(cgdo, MOSgateDrainOverlapCapFactorGiven, MOSgateDrainOverlapCapFactor);
                            MP(cgbo, MOSgateBulkOverlapCapFactorGiven, MOSgateBulkOverlapCapFactor);
                            MP(rsh, MOSsheetResistanceGiven, MOSsheetResistance);
                            MP(cj, MOSbulkCapFactorGiven, MOSbulkCapFactor);
                            MP(mj, MOSbulkJctBotGradingCoeffGiven, MOSbulkJctBotGradingCoeff);
                            MP(cjsw, MOSsideWallCapFactorGiven, MOSsideWallCapFactor);
                            MP(mjsw, MOSbulkJctSideGradingCoeffGiven, MOSbulkJctSideGradingCoeff);
                            MP(js, MOSjctSatCurDensityGiven, MOSjctSatCurDensity);
                            MP(tox, MOSoxideThicknessGiven, MOSoxideThickness);
                            MP(ld, MOSlatDiffGiven, MOSlatDiff);
                            MP(u0, MOSsurfaceMobilityGiven, MOSsurfaceMobility);
                            MP(fc, MOSfwdCapDepCoeffGiven, MOSfwdCapDepCoeff);
                            MP(nsub, MOSsubstrateDopingGiven, MOSsubstrateDoping);
                            MP(tpg, MOSgateTypeGiven, MOSgateType);
                            MP(nss, MOSsurfaceStateDensityGiven, MOSsurfaceStateDensity);
                            MP(tnom, MOStnomGiven, MOStnom);
                            MP(nfs, MOSfastSurfaceStateDensityGiven, MOSfastSurfaceStateDensity);
                            MP(uexp, MOScritFieldExpGiven, MOScritFieldExp);
                            MP(ucrit, MOScritFieldGiven, MOScritField);
                            MP(vmax, MOSmaxDriftVelGiven, MOSmaxDriftVel);
                            MP(xj, MOSjunctionDepthGiven, MOSjunctionDepth);
                            MP(neff, MOSchannelChargeGiven, MOSchannelCharge);
                            MP(eta, MOSetaGiven, MOSeta); MP(theta, MOSthetaGiven, MOStheta);
                            MP(kappa, MOSkappaGiven, MOSkappa);
                            MP(kv, MOSkvGiven, MOSkv); MP(nv, MOSnvGiven, MOSnv);
                            MP(kc, MOSkcGiven, MOSkc); MP(nc, MOSncGiven, MOSnc);
                            MP(lamda0, MOSlamda0Given, MOSlamda0); MP(lamda1, MOSlamda1Given, MOSlamda1);
                            MP(sigma, MOSsigmaGiven, MOSsigma); MP(gamma1, MOSgamma1Given, MOSgamma1);
#undef MP
                            // "delta" on the card: levels 2 and 3 store it in different members
                            // (mossetm.cc:216-219, :292-296)
                            y.delta = m->MOSnarrowFactorGiven ? m->MOSnarrowFactor :
                                      (m->MOSdeltaGiven ? m->MOSdelta : NaN);
                            it = H->mosmod_index.emplace(m, (int)H->mosmod.size()).first;
                            H->mosmod.push_back(y);
                        }
                        b200_mos x;
                        memset(&x, 0, sizeof(x));
                        x.d = ext(q->MOSdNode); x.g = ext(q->MOSgNode); x.s = ext(q->MOSsNode); x.b = ext(q->MOSbNode);
                        x.model = it->second;
                        x.l = opt(q->MOSlGiven, q->MOSl); x.w = opt(q->MOSwGiven, q->MOSw);
                        x.ad = opt(q->MOSdrainAreaGiven, q->MOSdrainArea);
                        x.as = opt(q->MOSsourceAreaGiven, q->MOSsourceArea);
                        x.pd = opt(q->MOSdrainPerimeterGiven, q->MOSdrainPerimeter);
                        x.ps = opt(q->MOSsourcePerimeterGiven, q->MOSsourcePerimeter);
                        x.nrd = opt(q->MOSdrainSquaresGiven, q->MOSdrainSquares);
                        x.nrs = opt(q->MOSsourceSquaresGiven, q->MOSsourceSquares);
                        x.m = opt(q->MOSmGiven, q->MOSmult);
                        x.temp = opt(q->MOStempGiven, q->MOStemp);
                        x.icvds = opt(q->MOSicVDSGiven, q->MOSicVDS);
                        x.icvgs = opt(q->MOSicVGSGiven, q->MOSicVGS);
                        x.icvbs = opt(q->MOSicVBSGiven, q->MOSicVBS);
                        x.off = q->MOSoff ? 1 : 0;
                        order_rest.push_back({B200_DEV_MOS, (int)H->mos.size()});
                        H->mos.push_back(x); H->names[B200_DEV_MOS].push_back(iname);
                    }
                    else if (!strcmp(dn, "JJ")) {
                        // the model AFTER JJdev::setup's model pass (jjset.cc:196-545); the
                        // instance's temperature factor and gap (jjset.cc:560-596)
                        sJJmodel *m = (sJJmodel*)dm;
                        sJJinstance *q = (sJJinstance*)gi;
                        if (q->JJcontrol || q->JJlser > 0.0 || q->JJlsh > 0.0) { H->error = "JJ control / lser / lsh"; return B200_E_UNSUPP; }
                        auto it = H->jjmod_index.find(m);
                        if (it == H->jjmod_index.end()) {
                            b200_jj_model y;
                            memset(&y, 0, sizeof(y));
                            y.rtype = m->JJrtype; y.ictype = m->JJictype; y.pi = m->JJpi;
                            y.criti = m->JJcriti; y.vg = m->JJvgnom; y.delv = m->JJdelv;
                            y.r0 = m->JJr0; y.rn = m->JJrn; y.cap = m->JJcap;
                            y.icfactor = m->JJicFactor; y.cmu = m->JJcmu; y.gmu = m->JJgmu;
                            y.vshunt = m->JJvShuntGiven ? m->JJvShunt : 0.0;
                            y.vdpbak = m->JJvdpbak; y.tsaccl = m->JJtsaccl; y.tsfact = m->JJtsfact;
                            it = H->jjmod_index.emplace(m, (int)H->jjmod.size()).first;
                            H->jjmod.push_back(y);
                        }
                        b200_jj x;
                        memset(&x, 0, sizeof(x));
                        x.pos = ext(q->JJrealPosNode); x.neg = ext(q->JJnegNode);
                        x.phase_node = q->JJphsNode > 0 ? ext(q->JJphsNode) : 0;
                        x.model = it->second;
                        x.area = opt(q->JJareaGiven && !q->JJicsGiven, q->JJarea);
                        x.ics = opt(q->JJicsGiven, q->JJics);
                        x.vshunt = opt(q->JJvshuntGiven, q->JJvshunt);
                        x.vg = q->JJvg; x.tcf = q->JJtcf;
                        x.icv = opt(q->JJinitVoltGiven, q->JJinitVoltage);
                        x.icphi = opt(q->JJinitPhaseGiven, q->JJinitPhase);
                        order_rest.push_back({B200_DEV_JJ, (int)H->jj.size()});
                        H->jj.push_back(x); H->names[B200_DEV_JJ].push_back(iname);
                    }
                    else if (!strcmp(dn, "TJM")) {
                        // JJ level 3.  The set-up instance carries its coefficient set after
                        // sTJMinstance::tjm_init1 (tjmset.cc:869-918) and the quantities that do
                        // not depend on the junction's area.
                        sTJMmodel *m = (sTJMmodel*)dm;
                        sTJMinstance *q = (sTJMinstance*)gi;
                        if (q->TJMlser > 0.0 || q->TJMlsh > 0.0) { H->error = "TJM lser / lsh"; return B200_E_UNSUPP; }
                        if (!q->tjm_p) { H->error = "TJM instance without a coefficient set (setup failed)"; return B200_E_BADPARM; }
                        auto it = H->tjmmod_index.find(m);
                        if (it == H->tjmmod_index.end()) {
                            b200_tjm_model y;
                            memset(&y, 0, sizeof(y));
                            y.rtype = m->TJMrtype; y.ictype = m->TJMictype; y.nterms = m->TJMnterms;
                            y.criti = m->TJMcriti; y.cap = m->TJMcap; y.cmu = m->TJMcmu; y.gmu = m->TJMgmu;
                            y.r0 = m->TJMr0; y.rsint = m->TJMrsint; y.vshunt = m->TJMvShunt;
                            y.tsaccl = m->TJMtsaccl; y.tsfact = m->TJMtsfact; y.vgnom = m->TJMvgNom;
                            it = H->tjmmod_index.emplace(m, (int)H->tjmmod.size()).first;
                            H->tjmmod.push_back(y);
                        }
                        b200_tjm x;
                        memset(&x, 0, sizeof(x));
                        x.pos = ext(q->TJMrealPosNode); x.neg = ext(q->TJMnegNode);
                        x.phase_node = q->TJMphsNode > 0 ? ext(q->TJMphsNode) : 0;
                        x.model = it->second;
                        x.area = opt(q->TJMareaGiven && !q->TJMicsGiven, q->TJMarea);
                        x.ics = opt(q->TJMicsGiven, q->TJMics);
                        x.vshunt = opt(q->TJMvshuntGiven, q->TJMvshunt);
                        x.tcf = q->TJMicTempFactor;
                        x.vdpbak = q->TJMvdpbak; x.omegaJ = q->TJMomegaJ;
                        x.kgap = q->tjm_kgap; x.kgap_rejpt = q->tjm_kgap_rejpt; x.alphaN = q->tjm_alphaN;
                        std::vector<double> cf;
                        const IFcomplex *arr[3] = { q->tjm_p, q->tjm_A, q->tjm_B };
                        for (int a = 0; a < 3; a++)
                            for (int k = 0; k < m->TJMnterms; k++) { cf.push_back(arr[a][k].real); cf.push_back(arr[a][k].imag); }
                        H->tjmcoef.push_back(cf);
                        x.icv = opt(q->TJMinitVoltGiven, q->TJMinitVoltage);
                        x.icphi = opt(q->TJMinitPhaseGiven, q->TJMinitPhase);
                        order_rest.push_back({B200_DEV_TJM, (int)H->tjm.size()});
                        H->tjm.push_back(x); H->names[B200_DEV_TJM].push_back(iname);
                    }
                    else {
                        H->error = std::string("device not built: ") + dn;
                        return B200_E_UNSUPP;
                    }
                }
            }
        }
    }
    // the highest EXTERNAL node: internal and branch nodes were appended by setup after them,
    // so every terminal number collected above is external
    D.nnodes = maxnode + 1;
    for (size_t i = 0; i < H->src.size(); i++) H->src[i].coef = H->coefs[i].data();
    for (auto &o : order_mut) { H->okind.push_back(o.kind); H->oindex.push_back(o.index); }
    for (auto &o : order_rest) { H->okind.push_back(o.kind); H->oindex.push_back(o.index); }
    D.nic = (int)H->ic_node.size(); D.ic_node = H->ic_node.data(); D.ic_val = H->ic_val.data();
    D.nnodeset = (int)H->ns_node.size(); D.ns_node = H->ns_node.data(); D.ns_val = H->ns_val.data();
    D.nres = (int)H->res.size(); D.res = H->res.data();
    D.ncap = (int)H->cap.size(); D.cap = H->cap.data();
    D.nind = (int)H->ind.size(); D.ind = H->ind.data();
    D.nmut = (int)H->mut.size(); D.mut = H->mut.data();
    D.nsrc = (int)H->src.size(); D.src = H->src.data();
    D.nmosmod = (int)H->mosmod.size(); D.mos_models = H->mosmod.data();
    D.nmos = (int)H->mos.size(); D.mos = H->mos.data();
    D.njjmod = (int)H->jjmod.size(); D.jj_models = H->jjmod.data();
    D.njj = (int)H->jj.size(); D.jj = H->jj.data();
    for (size_t i = 0; i < H->tjm.size(); i++) H->tjm[i].coef = H->tjmcoef[i].data();
    D.ntjmmod = (int)H->tjmmod.size(); D.tjm_models = H->tjmmod.data();
    D.ntjm = (int)H->tjm.size(); D.tjm = H->tjm.data();
    D.norder = (int)H->okind.size(); D.order_kind = H->okind.data(); D.order_index = H->oindex.data();
    return B200_OK;
}

// device time (CUDA events, b200_run_stats.ms_total) and ticks of the latest b200_harvest_run
static b200_run_stats last_run_stats;
extern "C" void b200_harvest_last_stats(double *ms_total, long long *ticks)
{
    if (ms_total) *ms_total = last_run_stats.ms_total;
    if (ticks) *ticks = last_run_stats.ticks;
}

// the `.save` list of the next b200_harvest_run: unknown indices in the reference's output order
// (n = 0: everything).  The reference keeps its saves in the script front end (sFtCirc), which a
// host passes on here.
static std::vector<int> next_saves;
extern "C" void b200_harvest_set_saves(int n, const int *unknown)
{
    next_saves.assign(unknown, unknown + (n > 0 ? n : 0));
}

// harvest -> b200_batch_create -> set per-trial parameters -> run -> results, through the
// dlopen'ed engine.  nparams parameter rows: (param id, instance name, values[ntrials]).
// counters[ntrials][B200_NCOUNTERS]; *waves (malloc'ed, caller frees) [ntrials][npoints][n].
extern "C" int b200_harvest_run(sCKT *ckt, const char *libpath, int ntrials, int device, int nparams,
    const int *param_id, const char *const *inst_name, const double *const *values,
    int *counters, int *n_out, int *npoints_out, double **times_out, double **waves_out, int op_only)
{
    Api api;
    if (!api.load(libpath)) return B200_E_NOGPU;
    b200_harvest *H = nullptr;
    int rc = b200_harvest_ckt(ckt, &H);
    if (rc) { fprintf(stderr, "b200 shim: harvest: %s\n", b200_harvest_error(H)); b200_harvest_free(H); return rc; }
    if (!op_only && !H->has_tran) { b200_harvest_free(H); return B200_E_BADPARM; }
    b200_batch *b = nullptr;
    rc = api.create(&H->deck, ntrials, device, &b);
    if (rc) { b200_harvest_free(H); return rc; }
    for (int i = 0; i < nparams && !rc; i++) {
        int kind = param_id[i] >= B200_P_TJM_ICS ? B200_DEV_TJM : param_id[i] >= B200_P_JJ_ICS ? B200_DEV_JJ : param_id[i] >= B200_P_SRC_DC ? B200_DEV_SRC : param_id[i] >= B200_P_IND_L ? B200_DEV_IND :
                   param_id[i] >= B200_P_RES_R ? B200_DEV_RES : param_id[i] >= B200_P_CAP_C ? B200_DEV_CAP : B200_DEV_MOS;
        int idx = b200_harvest_find(H, kind, inst_name[i]);
        rc = idx < 0 ? B200_E_BADPARM : api.set_param(b, param_id[i], idx, values[i]);
    }
    if (!rc && !next_saves.empty()) rc = api.set_saves(b, (int)next_saves.size(), next_saves.data());
    if (!rc) rc = op_only ? api.run_op(b) : api.run_tran(b);
    memset(&last_run_stats, 0, sizeof(last_run_stats));
    if (!rc) api.get_stats(b, &last_run_stats);
    if (!rc) {
        int n = api.num_saved(b), np = api.num_points(b);     // columns returned: the saved unknowns
        *n_out = n; *npoints_out = np;
        api.get_counters(b, counters);
        *times_out = (double*)malloc(sizeof(double)*(np ? np : 1));
        *waves_out = (double*)malloc(sizeof(double)*(size_t)ntrials*(np ? np : 1)*n);
        rc = api.get_waves(b, *times_out, *waves_out);
    }
    api.destroy(b);
    b200_harvest_free(H);
    return rc;
}

// Host-only check of a harvest (no CUDA call): the engine's once-per-batch analysis of the
// harvested deck -- sizes of pattern, LU and task program (b200_deck_analysis_stats).
extern "C" int b200_harvest_stats(sCKT *ckt, const char *libpath, long long *out, int nout)
{
    void *h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
    if (!h) return B200_E_NOGPU;
    int (*stats)(const b200_deck*, long long*, int) =
        (int(*)(const b200_deck*, long long*, int))dlsym(h, "b200_deck_analysis_stats");
    if (!stats) return B200_E_NOGPU;
    b200_harvest *H = nullptr;
    int rc = b200_harvest_ckt(ckt, &H);
    if (rc) fprintf(stderr, "b200 shim: harvest: %s\n", b200_harvest_error(H));
    else rc = stats(&H->deck, out, nout);
    b200_harvest_free(H);
    return rc;
}

// Test plumbing: the same harvested deck through the plain-C oracle (oracle/libwrsoracle.so,
// dlopen'ed) -- decks whose device tables only the reference's setup can make (TJM coefficient
// sets) reach the oracle this way.  counters[9] as wo_get_counters; *vals malloc'ed [npts][n].
extern "C" int b200_harvest_oracle(sCKT *ckt, const char *oraclepath, int nparams, const int *param_id,
    const char *const *inst_name, const double *value, int *counters, int *n_out, int *npts_out,
    double **times_out, double **vals_out)
{
    void *h = dlopen(oraclepath, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "b200 shim: %s\n", dlerror()); return B200_E_PANIC; }
    typedef struct wo_ckt wo_ckt;
    wo_ckt *(*create)(const b200_deck*) = (wo_ckt*(*)(const b200_deck*))dlsym(h, "wo_create");
    int (*setp)(wo_ckt*, int, int, double) = (int(*)(wo_ckt*, int, int, double))dlsym(h, "wo_set_param");
    int (*run)(wo_ckt*) = (int(*)(wo_ckt*))dlsym(h, "wo_run_tran");
    int (*nunk)(const wo_ckt*) = (int(*)(const wo_ckt*))dlsym(h, "wo_num_unknowns");
    int (*npts)(const wo_ckt*) = (int(*)(const wo_ckt*))dlsym(h, "wo_num_points");
    void (*getp)(const wo_ckt*, double*, double*) = (void(*)(const wo_ckt*, double*, double*))dlsym(h, "wo_get_points");
    void (*getc)(const wo_ckt*, int*) = (void(*)(const wo_ckt*, int*))dlsym(h, "wo_get_counters");
    void (*destroy)(wo_ckt*) = (void(*)(wo_ckt*))dlsym(h, "wo_destroy");
    if (!create || !setp || !run || !nunk || !npts || !getp || !getc || !destroy) return B200_E_PANIC;
    b200_harvest *H = nullptr;
    int rc = b200_harvest_ckt(ckt, &H);
    if (rc) { fprintf(stderr, "b200 shim: harvest: %s\n", b200_harvest_error(H)); b200_harvest_free(H); return rc; }
    wo_ckt *c = create(&H->deck);
    for (int i = 0; i < nparams; i++) {
        int kind = param_id[i] >= B200_P_TJM_ICS ? B200_DEV_TJM : param_id[i] >= B200_P_JJ_ICS ? B200_DEV_JJ : B200_DEV_MOS;
        int idx = b200_harvest_find(H, kind, inst_name[i]);
        if (idx < 0 || setp(c, param_id[i], idx, value[i])) { destroy(c); b200_harvest_free(H); return B200_E_BADPARM; }
    }
    int err = run(c);
    getc(c, counters);
    counters[0] = err;
    *n_out = nunk(c); *npts_out = npts(c);
    *times_out = (double*)malloc(sizeof(double)*(*npts_out ? *npts_out : 1));
    *vals_out = (double*)malloc(sizeof(double)*(size_t)(*npts_out ? *npts_out : 1)*(*n_out));
    getp(c, *times_out, *vals_out);
    destroy(c);
    b200_harvest_free(H);
    return B200_OK;
}
