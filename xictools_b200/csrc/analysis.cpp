// analysis.cpp -- host-side, once-per-batch work: what sCKT::setup(), spSaveForInitialization()
// and klu_analyze do once per trial in the reference (ckt.cc:2110-2227, spspice.cc:79-100,
// kluif.cc:569-590), done ONCE for the whole batch because every trial shares the pattern.
//
//  * node / branch / state allocation in the reference's setup order
//  * every device's TSTALLOC list -> CSC pattern (spbuild.cc:922-958) and, instead of
//    pointers into Ax, ordered GATHER lists: for each matrix entry and RHS row the
//    contribution slots in the reference's load order (ckt.cc:1146-1161), so that assembly
//    is an atomics-free ordered sum that reproduces the reference's summation order
//  * preloaded constants (Ainit): srcset.cc:158-170, restemp.cc:218-229
//  * breakpoint table: breakpt.cc, inptran.cc setup functions
//  * static pivot order (Markowitz on the structure, restricted to value-independent safe
//    pivots) and the compiled, level-scheduled LU / triangular-solve programs.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <set>
#include <queue>
#include <unordered_map>
#include <unordered_set>

#include "b200_internal.h"

namespace {

inline bool given(double x) { return !std::isnan(x); }

struct Coo {
    std::map<std::pair<int,int>, int> idx;     // (row, col) 1-based -> coo id
    std::vector<int> r, c;
    int add(int row, int col) {
        if (row == 0 || col == 0) return -1;
        auto it = idx.find({row, col});
        if (it != idx.end()) return it->second;
        int id = (int)r.size();
        idx[{row, col}] = id;
        r.push_back(row); c.push_back(col);
        return id;
    }
};

// ---- breakpoint table construction: ckt/breakpt.cc:58-221 (host replica) ----
struct BreakTab {
    std::vector<double> brks, lat_offs, lat_per;
    double breaks[2] = {0.0, 0.0};
    double minbrk = 0.0;
    double nextbreak(double t) const {
        double dt0 = 0.0;
        for (size_t i = 0; i < lat_offs.size(); i++) {
            double dt1 = lat_offs[i];
            if (t > lat_offs[i])
                dt1 += lat_per[i]*ceil((t - lat_offs[i])/lat_per[i]);
            if (fabs(dt1 - t) < minbrk)
                dt1 += lat_per[i];
            if (dt0 == 0.0 || dt1 < dt0)
                dt0 = dt1;
        }
        size_t i;
        for (i = 0; i < brks.size(); i++)
            if (brks[i] > t)
                break;
        while (i < brks.size() && fabs(t - brks[i]) < minbrk)
            i++;
        if (i == brks.size())
            return dt0;
        if (dt0 > 0.0 && dt0 < brks[i])
            return dt0;
        return brks[i];
    }
    void set_break(double time, double ckt_time) {
        for (size_t i = 0; i < lat_offs.size(); i++) {
            double dt0, dt1;
            if (time < lat_offs[i]) {
                dt0 = lat_offs[i] + lat_per[i];
                dt1 = lat_offs[i];
            }
            else {
                dt0 = lat_per[i]*ceil((time - lat_offs[i])/lat_per[i]) + lat_offs[i];
                dt1 = lat_per[i]*floor((time - lat_offs[i])/lat_per[i]) + lat_offs[i];
            }
            if (fabs(time - dt0) < minbrk || fabs(time - dt1) < minbrk) {
                if (time < breaks[0]) breaks[0] = nextbreak(ckt_time);
                breaks[1] = nextbreak(breaks[0]);
                return;
            }
        }
        if (brks.empty()) { brks.push_back(time); return; }
        bool done = false;
        for (size_t i = 0; i < brks.size() && !done; i++) {
            if (brks[i] > time) {
                if ((brks[i] - time) <= minbrk)
                    brks[i] = time;
                else if (i > 0 && time - brks[i-1] <= minbrk)
                    ;
                else
                    brks.insert(brks.begin() + i, time);
                done = true;
            }
        }
        if (!done) {
            if (time - brks.back() <= minbrk)
                ;
            else
                brks.push_back(time);
        }
        if (time < breaks[0]) breaks[0] = nextbreak(ckt_time);
        breaks[1] = nextbreak(breaks[0]);
    }
    void set_lattice(double offs, double per) {
        for (size_t i = 0; i < lat_offs.size(); i++)
            if (offs == lat_offs[i] && per == lat_per[i]) return;
        lat_offs.push_back(offs); lat_per.push_back(per);
    }
};

// PWL list normalisation: inptran.cc:1980-2074 (no TD=, no R=)
void build_pwl(const double *coef, int num, std::vector<double> &pts, std::vector<double> &slopes)
{
    std::vector<double> list(coef, coef + num);
    bool add0 = (list[0] > 0.0), add1 = (num & 1);
    if (add0 || add1) {
        std::vector<double> nl;
        if (add0) { nl.push_back(0.0); nl.push_back(num > 1 ? list[1] : 0.0); }
        nl.insert(nl.end(), list.begin(), list.end());
        if (add1) nl.push_back(num > 1 ? nl[nl.size() - 2] : 0.0);
        list.swap(nl);
    }
    int n = (int)list.size();
    for (int i = 2; i < n; ) {
        if (list[i] <= list[i-2]) {
            for (int k = i; k + 2 < n; k += 2) { list[k] = list[k+2]; list[k+1] = list[k+3]; }
            n -= 2;
            continue;
        }
        i += 2;
    }
    list.resize(n);
    pts = list;
    int ns = n/2 - 1;
    slopes.clear();
    if (ns > 0)
        for (int k = 0; k < ns; k++)
            slopes.push_back((list[2*k+3] - list[2*k+1])/(list[2*k+2] - list[2*k]));
    else
        slopes.push_back(0.0);
}

} // namespace

int b200_analyze(const b200_deck *dk, Analysis &an, std::string &err)
{
    const b200_options &o = dk->opt;
    if (dk->njj) {
        // what of devlib/jj is built (include/b200spice.h, b200_jj_model)
        if (dk->njjmod > JJ_MAXMODELS) { err = "more JJ models than JJ_MAXMODELS"; return B200_E_UNSUPP; }
        for (int i = 0; i < dk->njjmod; i++) {
            const b200_jj_model &m = dk->jj_models[i];
            if (m.rtype < 0 || m.rtype > 3) { err = "JJ rtype 4 (control current) is not built"; return B200_E_UNSUPP; }
            if (m.ictype < 0 || m.ictype > 1) { err = "JJ ictype 2..4 (control current) is not built"; return B200_E_UNSUPP; }
        }
    }
    if (dk->ntjm) {
        if (dk->njjmod + dk->ntjmmod > JJ_MAXMODELS) { err = "more JJ + TJM models than JJ_MAXMODELS"; return B200_E_UNSUPP; }
        for (int i = 0; i < dk->ntjmmod; i++)
            if (dk->tjm_models[i].nterms < 1 || dk->tjm_models[i].nterms > TJM_MAXTERMS) { err = "TJM nterms out of range"; return B200_E_UNSUPP; }
    }
    an.n = 0;
    an.ntype.assign(1, 0);
    an.names.assign(1, "0");
    auto new_node = [&](int type, const std::string &nm) {
        an.n++;
        an.ntype.push_back(type);
        an.names.push_back(nm);
        return an.n;
    };
    for (int i = 1; i < dk->nnodes; i++)
        new_node(0, "n" + std::to_string(i));

    MosTopo &M = an.mos; CapTopo &Cp = an.cap; SrcTopo &S = an.src; IndTopo &L = an.ind;
    M.n = dk->nmos; Cp.n = dk->ncap; S.n = dk->nsrc; an.nres = dk->nres; L.n = dk->nind;
    L.pos.resize(L.n); L.neg.resize(L.n); L.branch.assign(L.n, 0); L.state.resize(L.n);
    L.cslot.assign(L.n, -1); L.rslot.assign(L.n, -1); L.icgiven.assign(L.n, 0); L.ic.assign(L.n, 0.0);
    M.d.resize(M.n); M.g.resize(M.n); M.s.resize(M.n); M.b.resize(M.n); M.dp.resize(M.n);
    M.sp.resize(M.n); M.type.resize(M.n); M.level.resize(M.n); M.off.resize(M.n);
    M.state.resize(M.n); M.cslot.assign((size_t)M.n*MT_NSTAMP, -1);
    M.rslot.assign((size_t)M.n*MRH_NRHS, -1);
    Cp.pos.resize(Cp.n); Cp.neg.resize(Cp.n); Cp.state.resize(Cp.n);
    Cp.cslot.assign((size_t)Cp.n*4, -1); Cp.rslot.assign((size_t)Cp.n*2, -1);
    S.kind.resize(S.n); S.pos.resize(S.n); S.neg.resize(S.n); S.branch.assign(S.n, 0);
    S.fn.resize(S.n); S.dcgiven.resize(S.n); S.dc.resize(S.n);
    S.rslot.assign((size_t)S.n*2, -1); S.coef_off.resize(S.n); S.ncoef.resize(S.n);

    Coo coo;
    struct Stamp { int coo; };
    // per device, the coo ids of its stamps (resolved to Ax indices after CSC build)
    std::vector<int> mos_coo((size_t)M.n*MT_NSTAMP, -1), cap_coo((size_t)Cp.n*4, -1);
    std::vector<int> res_coo((size_t)dk->nres*4, -1), src_coo((size_t)S.n*4, -1);
    std::vector<int> ind_coo((size_t)L.n*5, -1);
    MutTopo &MU = an.mut;
    MU.n = dk->nmut; MU.ind1.resize(MU.n); MU.ind2.resize(MU.n); MU.k.resize(MU.n);
    MU.cslot.assign((size_t)MU.n*2, -1);
    std::vector<int> mut_coo((size_t)MU.n*2, -1);
    JJTopo &J = an.jj;
    J.n = dk->njj; J.nmodels = dk->njjmod;
    J.models.assign(dk->jj_models, dk->jj_models + dk->njjmod);
    J.pos.resize(J.n); J.neg.resize(J.n); J.phs.resize(J.n); J.model.resize(J.n); J.state.resize(J.n);
    J.cslot.assign((size_t)J.n*JT_NSTAMP, -1); J.rslot.assign((size_t)J.n*JR_NRHS, -1);
    J.vless.resize(J.n); J.vmore.resize(J.n); J.vg.resize(J.n); J.initV.assign(J.n, 0.0); J.initPhi.assign(J.n, 0.0);
    J.area0.resize(J.n); J.ics0.resize(J.n); J.vshunt.resize(J.n); J.tcf.resize(J.n);
    std::vector<int> jj_coo((size_t)J.n*9, -1);
    TJMTopo &TJ = an.tjm;
    TJ.n = dk->ntjm; TJ.nmodels = dk->ntjmmod; TJ.nterms = 0;
    if (TJ.n) TJ.models.assign(dk->tjm_models, dk->tjm_models + dk->ntjmmod);
    for (auto &md : TJ.models) TJ.nterms = std::max(TJ.nterms, md.nterms);
    TJ.pos.resize(TJ.n); TJ.neg.resize(TJ.n); TJ.phs.resize(TJ.n); TJ.model.resize(TJ.n); TJ.state.resize(TJ.n);
    TJ.cslot.assign((size_t)TJ.n*JT_NSTAMP, -1); TJ.rslot.assign((size_t)TJ.n*JR_NRHS, -1);
    TJ.initV.assign(TJ.n, 0.0); TJ.initPhi.assign(TJ.n, 0.0);
    TJ.area0.resize(TJ.n); TJ.ics0.resize(TJ.n); TJ.vshunt.resize(TJ.n); TJ.tcf.resize(TJ.n);
    TJ.omegaJ.resize(TJ.n); TJ.kgap.resize(TJ.n); TJ.kgap_rejpt.resize(TJ.n); TJ.alphaN.resize(TJ.n); TJ.vdpbak.resize(TJ.n);
    TJ.coef.assign((size_t)TJ.n*6*std::max(TJ.nterms, 1), 0.0);
    std::vector<int> tjm_coo((size_t)TJ.n*9, -1);
    // phase flags (jjset.cc:656-681, indset.cc:104-120): nodes of junctions with a supercurrent,
    // and -- once any such junction exists -- nodes of inductors
    std::vector<char> phase_node(dk->nnodes + 1, 0);
    bool jjDCphase = false;
    for (int i = 0; i < dk->njj; i++) {
        const b200_jj &in = dk->jj[i];
        if (in.model < 0 || in.model >= dk->njjmod) { err = "bad JJ model index"; return B200_E_BADPARM; }
        if (dk->jj_models[in.model].ictype > 0) {
            if (in.pos > 0) phase_node[in.pos] = 1;
            if (in.neg > 0) phase_node[in.neg] = 1;
            jjDCphase = true;
        }
    }
    for (int i = 0; i < dk->ntjm; i++) {                // tjmset.cc:690-708: whatever the ictype
        const b200_tjm &in = dk->tjm[i];
        if (in.model < 0 || in.model >= dk->ntjmmod) { err = "bad TJM model index"; return B200_E_BADPARM; }
        if (in.pos > 0) phase_node[in.pos] = 1;
        if (in.neg > 0) phase_node[in.neg] = 1;
        jjDCphase = true;
    }
    if (jjDCphase)
        for (int i = 0; i < dk->nind; i++) {
            if (dk->ind[i].pos > 0) phase_node[dk->ind[i].pos] = 1;
            if (dk->ind[i].neg > 0) phase_node[dk->ind[i].neg] = 1;
        }
    std::vector<char> res_preload(dk->nres, 1);
    an.jjDCphase = jjDCphase ? 1 : 0;
    an.rph_res.clear(); an.rph_cls.clear(); an.rph_g.clear(); an.rph_cslot.clear();
    std::vector<double> res_g(dk->nres, 0.0);
    std::vector<char> mos_has_dcon(M.n, 0), mos_has_scon(M.n, 0);

    an.numStates = 0;
    // ---- setup pass in load order (ckt.cc:2202-2210) ----
    for (int oi = 0; oi < dk->norder; oi++) {
        int i = dk->order_index[oi];
        switch (dk->order_kind[oi]) {
        case B200_DEV_MOS: {
            const b200_mos &in = dk->mos[i];
            if (in.model < 0 || in.model >= dk->nmosmod) { err = "bad MOS model index"; return B200_E_BADPARM; }
            const b200_mos_model &md = dk->mos_models[in.model];
            int level = md.level;
            if (level != 1 && level != 2 && level != 3 && level != 6) level = 1;
            M.type[i] = md.type > 0 ? 1 : -1;
            M.level[i] = level;
            M.off[i] = in.off;
            M.state[i] = an.numStates;
            an.numStates += MS_NSTATES;
            double rd = given(md.rd) ? md.rd : 0, rs = given(md.rs) ? md.rs : 0;
            double rsh = given(md.rsh) ? md.rsh : 0;
            double nrs = (!given(in.nrs) || in.nrs == 0) ? 1 : in.nrs;
            M.d[i] = in.d; M.g[i] = in.g; M.s[i] = in.s; M.b[i] = in.b;
            // mosset.cc:375-402
            if (rd != 0 || rsh != 0)
                M.dp[i] = new_node(0, "mos" + std::to_string(i) + "#internal#drain");
            else
                M.dp[i] = in.d;
            if (rs != 0 || (nrs != 0 && rsh != 0))
                M.sp[i] = new_node(0, "mos" + std::to_string(i) + "#internal#source");
            else
                M.sp[i] = in.s;
            mos_has_dcon[i] = (given(md.rd) && md.rd != 0) || (!given(md.rd) && given(md.rsh) && md.rsh != 0);
            mos_has_scon[i] = (given(md.rs) && md.rs != 0) || (!given(md.rs) && given(md.rsh) && md.rsh != 0);
            int d = M.d[i], g = M.g[i], s = M.s[i], b = M.b[i], dp = M.dp[i], sp = M.sp[i];
            int *c = &mos_coo[(size_t)i*MT_NSTAMP];
            // mosset.cc:53-77 (allocation order), indexed by our stamp enum (load order)
            c[MT_Dd] = coo.add(d, d);     c[MT_Gg] = coo.add(g, g);     c[MT_Ss] = coo.add(s, s);
            c[MT_Bb] = coo.add(b, b);     c[MT_DPdp] = coo.add(dp, dp); c[MT_SPsp] = coo.add(sp, sp);
            c[MT_Ddp] = coo.add(d, dp);   c[MT_Gb] = coo.add(g, b);     c[MT_Gdp] = coo.add(g, dp);
            c[MT_Gsp] = coo.add(g, sp);   c[MT_Ssp] = coo.add(s, sp);   c[MT_Bg] = coo.add(b, g);
            c[MT_Bdp] = coo.add(b, dp);   c[MT_Bsp] = coo.add(b, sp);   c[MT_DPsp] = coo.add(dp, sp);
            c[MT_DPd] = coo.add(dp, d);   c[MT_DPg] = coo.add(dp, g);   c[MT_DPb] = coo.add(dp, b);
            c[MT_SPg] = coo.add(sp, g);   c[MT_SPs] = coo.add(sp, s);   c[MT_SPb] = coo.add(sp, b);
            c[MT_SPdp] = coo.add(sp, dp);
            break; }
        case B200_DEV_CAP: {
            const b200_cap &in = dk->cap[i];
            Cp.pos[i] = in.pos; Cp.neg[i] = in.neg;
            Cp.state[i] = an.numStates; an.numStates += 2;
            int *c = &cap_coo[(size_t)i*4];
            c[0] = coo.add(in.pos, in.pos); c[1] = coo.add(in.neg, in.neg);
            c[2] = coo.add(in.pos, in.neg); c[3] = coo.add(in.neg, in.pos);
            break; }
        case B200_DEV_RES: {
            const b200_res &in = dk->res[i];
            int *c = &res_coo[(size_t)i*4];
            c[0] = coo.add(in.pos, in.pos); c[1] = coo.add(in.neg, in.neg);
            c[2] = coo.add(in.pos, in.neg); c[3] = coo.add(in.neg, in.pos);
            double r = in.r, mm = given(in.m) ? in.m : 1.0;
            if (fabs(r) < o.resmin)
                r = (r < 0.0) ? -o.resmin : o.resmin;
            res_g[i] = mm/(r*1.0);                       // restemp.cc:209
            // restemp.cc:213-223: no preload when a terminal is a phase node
            if (jjDCphase && ((in.pos > 0 && phase_node[in.pos]) || (in.neg > 0 && phase_node[in.neg])))
                res_preload[i] = 0;
            break; }
        case B200_DEV_TJM: {
            // tjmset.cc: same pointer set as the RSJ model; what does not depend on the area comes
            // with the instance (b200_tjm), the rest is derived per trial by k_tjm_temp
            const b200_tjm &in = dk->tjm[i];
            const b200_tjm_model &md = dk->tjm_models[in.model];
            TJ.pos[i] = in.pos; TJ.neg[i] = in.neg; TJ.phs[i] = in.phase_node; TJ.model[i] = in.model;
            TJ.state[i] = an.numStates; an.numStates += TS_NSTATES;
            double vsh = given(in.vshunt) ? in.vshunt : md.vshunt;
            if (given(in.vshunt) && (in.vshunt < 0.0 || in.vshunt > md.vgnom)) vsh = 0.0;
            TJ.vshunt[i] = vsh;
            TJ.area0[i] = in.area; TJ.ics0[i] = in.ics; TJ.tcf[i] = in.tcf;
            TJ.omegaJ[i] = in.omegaJ; TJ.kgap[i] = in.kgap; TJ.kgap_rejpt[i] = in.kgap_rejpt;
            TJ.alphaN[i] = in.alphaN; TJ.vdpbak[i] = in.vdpbak;
            if (!in.coef) { err = "TJM instance without a coefficient set"; return B200_E_BADPARM; }
            for (int a = 0; a < 3; a++)
                for (int k = 0; k < md.nterms; k++) {
                    TJ.coef[(((size_t)i*3 + a)*TJ.nterms + k)*2] = in.coef[2*(a*md.nterms + k)];
                    TJ.coef[(((size_t)i*3 + a)*TJ.nterms + k)*2 + 1] = in.coef[2*(a*md.nterms + k) + 1];
                }
            int *c = &tjm_coo[(size_t)i*9];
            c[JT_PP] = coo.add(in.pos, in.pos); c[JT_NN] = coo.add(in.neg, in.neg);
            c[JT_NP] = coo.add(in.neg, in.pos); c[JT_PN] = coo.add(in.pos, in.neg);
            c[JT_RPP] = coo.add(in.pos, in.pos); c[JT_RPN] = coo.add(in.pos, in.neg);
            c[JT_RNP] = coo.add(in.neg, in.pos); c[JT_RNN] = coo.add(in.neg, in.neg);
            c[8] = in.phase_node > 0 ? coo.add(in.phase_node, in.phase_node) : -1;
            break; }
        case B200_DEV_JJ: {
            // jjset.cc:130-200 (TSTALLOC order), :560-612 (vshunt, vless / vmore), :794 (states)
            const b200_jj &in = dk->jj[i];
            const b200_jj_model &md = dk->jj_models[in.model];
            J.pos[i] = in.pos; J.neg[i] = in.neg; J.phs[i] = in.phase_node; J.model[i] = in.model;
            J.state[i] = an.numStates; an.numStates += JS_NSTATES;
            J.vg[i] = given(in.vg) ? in.vg : md.vg;
            J.tcf[i] = given(in.tcf) ? in.tcf : 1.0;
            double vsh = given(in.vshunt) ? in.vshunt : md.vshunt;
            if (vsh < 0.0 || vsh > md.vg) vsh = 0.0;
            J.vshunt[i] = vsh;
            J.vless[i] = J.vg[i] - 0.5*md.delv;
            J.vmore[i] = J.vg[i] + 0.5*md.delv;
            J.area0[i] = in.area; J.ics0[i] = in.ics;
            int *c = &jj_coo[(size_t)i*9];
            c[JT_PP] = coo.add(in.pos, in.pos); c[JT_NN] = coo.add(in.neg, in.neg);
            c[JT_NP] = coo.add(in.neg, in.pos); c[JT_PN] = coo.add(in.pos, in.neg);
            c[JT_RPP] = coo.add(in.pos, in.pos); c[JT_RPN] = coo.add(in.pos, in.neg);
            c[JT_RNP] = coo.add(in.neg, in.pos); c[JT_RNN] = coo.add(in.neg, in.neg);
            c[8] = in.phase_node > 0 ? coo.add(in.phase_node, in.phase_node) : -1;
            break; }
        case B200_DEV_SRC: {
            const b200_src &in = dk->src[i];
            S.kind[i] = in.kind; S.pos[i] = in.pos; S.neg[i] = in.neg;
            S.fn[i] = in.fn; S.dcgiven[i] = given(in.dc); S.dc[i] = given(in.dc) ? in.dc : 0.0;
            if (in.kind == B200_SRC_V) {
                S.branch[i] = new_node(1, "src" + std::to_string(i) + "#branch");
                int *c = &src_coo[(size_t)i*4];
                c[0] = coo.add(in.pos, S.branch[i]); c[1] = coo.add(in.neg, S.branch[i]);
                c[2] = coo.add(S.branch[i], in.neg); c[3] = coo.add(S.branch[i], in.pos);
            }
            if (in.kind == B200_SRC_I && in.fn != B200_FN_NONE)
                an.numStates += 1;
            break; }
        case B200_DEV_IND: {
            // indset.cc:72-135: branch unknown, 2 states (flux, volt), TSTALLOC order :50-58
            const b200_ind &in = dk->ind[i];
            if (given(in.m) && in.m != 1.0) { err = "inductor M != 1 is not built"; return B200_E_UNSUPP; }
            L.pos[i] = in.pos; L.neg[i] = in.neg;
            L.icgiven[i] = given(in.ic); L.ic[i] = given(in.ic) ? in.ic : 0.0;
            L.branch[i] = new_node(1, "ind" + std::to_string(i) + "#branch");
            L.state[i] = an.numStates; an.numStates += 2;
            int *c = &ind_coo[(size_t)i*5];
            int br = L.branch[i];
            c[0] = coo.add(br, br); c[1] = coo.add(in.pos, br); c[2] = coo.add(in.neg, br);
            c[3] = coo.add(br, in.neg); c[4] = coo.add(br, in.pos);
            break; }
        case B200_DEV_MUT:
            break;              // set up after the inductors, below
        default:
            err = "unsupported device kind in load order";
            return B200_E_UNSUPP;
        }
    }
    // mutual inductors: INDdev::setup does them after all inductors (indset.cc:186-238)
    L.mptr.assign(L.n + 1, 0);
    {
        std::vector<std::vector<std::pair<int,double>>> per(L.n);
        for (int oi = 0; oi < dk->norder; oi++) {
            if (dk->order_kind[oi] != B200_DEV_MUT) continue;
            int i = dk->order_index[oi];
            const b200_mut &in = dk->mut[i];
            if (in.ind1 < 0 || in.ind1 >= L.n || in.ind2 < 0 || in.ind2 >= L.n) {
                err = "mutual inductor refers to an unknown inductor"; return B200_E_BADPARM;
            }
            MU.ind1[i] = in.ind1; MU.ind2[i] = in.ind2; MU.k[i] = in.k;
            mut_coo[(size_t)i*2] = coo.add(L.branch[in.ind1], L.branch[in.ind2]);
            mut_coo[(size_t)i*2 + 1] = coo.add(L.branch[in.ind2], L.branch[in.ind1]);
            per[in.ind1].push_back({in.ind2, in.k});
            per[in.ind2].push_back({in.ind1, in.k});
        }
        for (int i = 0; i < L.n; i++) {
            L.mptr[i+1] = L.mptr[i] + (int)per[i].size();
            for (auto &pr : per[i]) { L.mother.push_back(pr.first); L.mk.push_back(pr.second); }
        }
    }

    // ---- CSC pattern: columns ascending, rows ascending (spbuild.cc:922-958) ----
    int n = an.n;
    an.nnz = (int)coo.r.size();
    an.Ap.assign(n + 1, 0);
    for (int e = 0; e < an.nnz; e++) an.Ap[coo.c[e]]++;          // count per col (1-based)
    for (int j = 1; j <= n; j++) an.Ap[j] += an.Ap[j-1];
    {
        std::vector<std::vector<int>> colrows(n);
        for (int e = 0; e < an.nnz; e++) colrows[coo.c[e] - 1].push_back(coo.r[e] - 1);
        an.Ai.clear();
        for (int j = 0; j < n; j++) {
            std::sort(colrows[j].begin(), colrows[j].end());
            an.Ai.insert(an.Ai.end(), colrows[j].begin(), colrows[j].end());
        }
    }
    std::vector<int> coo2ax(an.nnz);
    for (int e = 0; e < an.nnz; e++) {
        int j = coo.c[e] - 1, r = coo.r[e] - 1;
        auto b = an.Ai.begin() + an.Ap[j], en = an.Ai.begin() + an.Ap[j+1];
        coo2ax[e] = (int)(std::lower_bound(b, en, r) - an.Ai.begin());
    }
    an.diag.assign(n + 1, -1);
    for (int j = 0; j < n; j++)
        for (int p = an.Ap[j]; p < an.Ap[j+1]; p++)
            if (an.Ai[p] == j) an.diag[j+1] = p;

    // ---- contribution slots + ordered gather lists, preload ----
    std::vector<std::vector<int>> cg(an.nnz), rg(n + 1);
    an.ncslot = an.nrslot = 0;
    an.Ainit.assign(an.nnz, 0.0);
    an.dc_active.assign(an.nnz, 0);
    auto add_c = [&](int cooid) -> int {
        if (cooid < 0) return -1;
        int k = coo2ax[cooid];
        cg[k].push_back(an.ncslot);
        return an.ncslot++;
    };
    auto add_r = [&](int node) -> int {
        if (node == 0) return -1;
        rg[node].push_back(an.nrslot);
        return an.nrslot++;
    };
    // sources' +-1 are ldset in setup (srcset.cc:158-170) ...
    for (int oi = 0; oi < dk->norder; oi++) {
        int i = dk->order_index[oi];
        if (dk->order_kind[oi] == B200_DEV_SRC && S.kind[i] == B200_SRC_V) {
            const int *c = &src_coo[(size_t)i*4];
            if (S.pos[i]) { if (c[0] >= 0) an.Ainit[coo2ax[c[0]]] = 1.0; if (c[3] >= 0) an.Ainit[coo2ax[c[3]]] = 1.0; }
            if (S.neg[i]) { if (c[1] >= 0) an.Ainit[coo2ax[c[1]]] = -1.0; if (c[2] >= 0) an.Ainit[coo2ax[c[2]]] = -1.0; }
        }
    }
    // inductor couplings are preldset in setup (indset.cc:117-131)
    for (int i = 0; i < L.n; i++) {
        const int *c = &ind_coo[(size_t)i*5];
        if (L.pos[i]) { if (c[1] >= 0) an.Ainit[coo2ax[c[1]]] = 1.0; if (c[4] >= 0) an.Ainit[coo2ax[c[4]]] = 1.0; }
        if (L.neg[i]) { if (c[2] >= 0) an.Ainit[coo2ax[c[2]]] = -1.0; if (c[3] >= 0) an.Ainit[coo2ax[c[3]]] = -1.0; }
    }
    // ... resistor conductances are ldadd in temp() (restemp.cc:218-229), load order
    an.nres = dk->nres; an.resmin = o.resmin;
    an.res_r.resize(dk->nres); an.res_m.resize(dk->nres);
    for (int i = 0; i < dk->nres; i++) { an.res_r[i] = dk->res[i].r; an.res_m[i] = given(dk->res[i].m) ? dk->res[i].m : 1.0; }
    {
        std::vector<std::vector<std::pair<int, signed char>>> terms(an.nnz);
        std::vector<double> base(an.Ainit);
        for (int oi = 0; oi < dk->norder; oi++) {
            int i = dk->order_index[oi];
            if (dk->order_kind[oi] == B200_DEV_RES && res_g[i] != 0.0 && res_preload[i]) {
                const int *c = &res_coo[(size_t)i*4];
                if (c[0] >= 0) { an.Ainit[coo2ax[c[0]]] += res_g[i]; terms[coo2ax[c[0]]].push_back({i, 1}); }
                if (c[1] >= 0) { an.Ainit[coo2ax[c[1]]] += res_g[i]; terms[coo2ax[c[1]]].push_back({i, 1}); }
                if (c[2] >= 0) { an.Ainit[coo2ax[c[2]]] += -res_g[i]; terms[coo2ax[c[2]]].push_back({i, -1}); }
                if (c[3] >= 0) { an.Ainit[coo2ax[c[3]]] += -res_g[i]; terms[coo2ax[c[3]]].push_back({i, -1}); }
            }
        }
        an.rt_entry.clear(); an.rt_base.clear(); an.rt_ptr.assign(1, 0); an.rt_res.clear(); an.rt_sign.clear();
        for (int p = 0; p < an.nnz; p++) {
            if (terms[p].empty()) continue;
            an.rt_entry.push_back(p); an.rt_base.push_back(base[p]);
            for (auto &tm : terms[p]) { an.rt_res.push_back(tm.first); an.rt_sign.push_back(tm.second); }
            an.rt_ptr.push_back((int)an.rt_res.size());
        }
    }
    // the phase node's own equation x[phs] = rhs[phs] is preldset (jjset.cc:815-816)
    for (int i = 0; i < J.n; i++)
        if (jj_coo[(size_t)i*9 + 8] >= 0) an.Ainit[coo2ax[jj_coo[(size_t)i*9 + 8]]] = 1.0;
    for (int i = 0; i < TJ.n; i++)
        if (tjm_coo[(size_t)i*9 + 8] >= 0) an.Ainit[coo2ax[tjm_coo[(size_t)i*9 + 8]]] = 1.0;
    an.rload_slot.clear(); an.rload_val.clear();
    for (int oi = 0; oi < dk->norder; oi++) {
        int i = dk->order_index[oi];
        switch (dk->order_kind[oi]) {
        case B200_DEV_RES: {
            // RESdev::load for a resistor that was not preloaded (resload.cc:288-292): four
            // constant terms at this place of the load order
            if (res_preload[i]) break;
            const int *c = &res_coo[(size_t)i*4];
            const b200_res &in = dk->res[i];
            // nodetype(): ground, voltage or phase (resload.cc:53-64)
            bool vp = in.pos > 0 && !phase_node[in.pos], vn = in.neg > 0 && !phase_node[in.neg];
            bool pp = in.pos > 0 && phase_node[in.pos], pn = in.neg > 0 && phase_node[in.neg];
            int cls = (!vp && !vn) ? 0 : (vp && pn) ? 1 : (pp && vn) ? 2 : 3;
            an.rph_res.push_back(i); an.rph_cls.push_back(cls); an.rph_g.push_back(res_g[i]);
            for (int k = 0; k < 4; k++) {
                int sl = add_c(c[k]);
                an.rph_cslot.push_back(sl);
                // which of the four are live in the phase-mode DC modes
                bool dc = cls == 0 || cls == 3 || (cls == 1 && (k == 0 || k == 3)) || (cls == 2 && (k == 1 || k == 2));
                if (sl >= 0 && dc) an.dc_active[coo2ax[c[k]]] = 1;
            }
            break; }
        case B200_DEV_TJM: {
            // tjm_load (tjmload.cc:436-456) stamps like jj_load, then the shunt (tjmload.cc:213-218)
            const int *c = &tjm_coo[(size_t)i*9];
            int *cs = &TJ.cslot[(size_t)i*JT_NSTAMP];
            int *rs = &TJ.rslot[(size_t)i*JR_NRHS];
            if (TJ.phs[i] > 0) rs[JR_phs] = add_r(TJ.phs[i]);
            if (!TJ.neg[i]) { cs[JT_PP] = add_c(c[JT_PP]); rs[JR_pos] = add_r(TJ.pos[i]); }
            else if (!TJ.pos[i]) { cs[JT_NN] = add_c(c[JT_NN]); rs[JR_neg] = add_r(TJ.neg[i]); }
            else {
                cs[JT_PP] = add_c(c[JT_PP]); cs[JT_PN] = add_c(c[JT_PN]);
                cs[JT_NP] = add_c(c[JT_NP]); cs[JT_NN] = add_c(c[JT_NN]);
                rs[JR_pos] = add_r(TJ.pos[i]); rs[JR_neg] = add_r(TJ.neg[i]);
            }
            if (TJ.vshunt[i] > 1e-12) {
                cs[JT_RPP] = add_c(c[JT_RPP]); cs[JT_RPN] = add_c(c[JT_RPN]);
                cs[JT_RNP] = add_c(c[JT_RNP]); cs[JT_RNN] = add_c(c[JT_RNN]);
            }
            for (int k = 0; k < JT_NSTAMP; k++) if (cs[k] >= 0) an.dc_active[coo2ax[c[k]]] = 1;
            break; }
        case B200_DEV_JJ: {
            const int *c = &jj_coo[(size_t)i*9];
            int *cs = &J.cslot[(size_t)i*JT_NSTAMP];
            int *rs = &J.rslot[(size_t)i*JR_NRHS];
            // jj_load (jjload.cc:826-866): the phase node's rhs is stored first
            if (J.phs[i] > 0) rs[JR_phs] = add_r(J.phs[i]);
            if (!J.neg[i]) { cs[JT_PP] = add_c(c[JT_PP]); rs[JR_pos] = add_r(J.pos[i]); }
            else if (!J.pos[i]) { cs[JT_NN] = add_c(c[JT_NN]); rs[JR_neg] = add_r(J.neg[i]); }
            else {
                cs[JT_PP] = add_c(c[JT_PP]); cs[JT_PN] = add_c(c[JT_PN]);
                cs[JT_NP] = add_c(c[JT_NP]); cs[JT_NN] = add_c(c[JT_NN]);
                rs[JR_pos] = add_r(J.pos[i]); rs[JR_neg] = add_r(J.neg[i]);
            }
            // the shunt implied by vshunt (jjload.cc:351-356); a trial whose gshunt is not
            // positive adds an exact 0.0
            if (J.vshunt[i] > 1e-12) {
                cs[JT_RPP] = add_c(c[JT_RPP]); cs[JT_RPN] = add_c(c[JT_RPN]);
                cs[JT_RNP] = add_c(c[JT_RNP]); cs[JT_RNN] = add_c(c[JT_RNN]);
            }
            for (int k = 0; k < JT_NSTAMP; k++) if (cs[k] >= 0) an.dc_active[coo2ax[c[k]]] = 1;
            break; }
        case B200_DEV_MOS: {
            const int *c = &mos_coo[(size_t)i*MT_NSTAMP];
            int *cs = &M.cslot[(size_t)i*MT_NSTAMP];
            for (int k = 0; k < MT_NSTAMP; k++) {
                // stamps that only ever carry +-dcon / +-scon contribute exact zeros when the
                // device has no series resistance (x + 0 == x): drop them.
                bool dconly = (k == MT_Dd || k == MT_Ddp || k == MT_DPd);
                bool sconly = (k == MT_Ss || k == MT_Ssp || k == MT_SPs);
                if ((dconly && !mos_has_dcon[i]) || (sconly && !mos_has_scon[i])) continue;
                cs[k] = add_c(c[k]);
                if (c[k] >= 0) an.dc_active[coo2ax[c[k]]] = 1;
            }
            int *rs = &M.rslot[(size_t)i*MRH_NRHS];
            rs[MRH_g] = add_r(M.g[i]);   rs[MRH_b] = add_r(M.b[i]);
            rs[MRH_dp] = add_r(M.dp[i]); rs[MRH_sp] = add_r(M.sp[i]);
            break; }
        case B200_DEV_CAP: {
            const int *c = &cap_coo[(size_t)i*4];
            int *cs = &Cp.cslot[(size_t)i*4];
            int *rs = &Cp.rslot[(size_t)i*2];
            // capload.cc:75-97
            if (Cp.pos[i] == 0) { cs[1] = add_c(c[1]); rs[1] = add_r(Cp.neg[i]); }
            else if (Cp.neg[i] == 0) { cs[0] = add_c(c[0]); rs[0] = add_r(Cp.pos[i]); }
            else {
                cs[0] = add_c(c[0]); cs[1] = add_c(c[1]); cs[2] = add_c(c[2]); cs[3] = add_c(c[3]);
                rs[0] = add_r(Cp.pos[i]); rs[1] = add_r(Cp.neg[i]);
            }
            break; }
        case B200_DEV_MUT:
            // mutload.cc: ldadd(br1br2, -factor), ldadd(br2br1, -factor)
            MU.cslot[(size_t)i*2] = add_c(mut_coo[(size_t)i*2]);
            MU.cslot[(size_t)i*2 + 1] = add_c(mut_coo[(size_t)i*2 + 1]);
            if (jjDCphase)
                for (int k = 0; k < 2; k++) if (mut_coo[(size_t)i*2 + k] >= 0) an.dc_active[coo2ax[mut_coo[(size_t)i*2 + k]]] = 1;
            break;
        case B200_DEV_IND:
            // indload.cc: rhsadd(brEq, veq) then ldadd(ibrIbr, -req)
            L.rslot[i] = add_r(L.branch[i]);
            L.cslot[i] = add_c(ind_coo[(size_t)i*5]);
            if (jjDCphase && ind_coo[(size_t)i*5] >= 0) an.dc_active[coo2ax[ind_coo[(size_t)i*5]]] = 1;   // -2*pi*L/Phi0 in phase-mode DC
            break;
        case B200_DEV_SRC: {
            int *rs = &S.rslot[(size_t)i*2];
            if (S.kind[i] == B200_SRC_V) rs[0] = add_r(S.branch[i]);
            else { rs[0] = add_r(S.pos[i]); rs[1] = add_r(S.neg[i]); }
            break; }
        default: break;
        }
    }
    for (int k = 0; k < an.nnz; k++) if (an.Ainit[k] != 0.0) an.dc_active[k] = 1;
    an.cg_ptr.assign(an.nnz + 1, 0);
    for (int k = 0; k < an.nnz; k++) {
        an.cg_ptr[k+1] = an.cg_ptr[k] + (int)cg[k].size();
        an.cg_idx.insert(an.cg_idx.end(), cg[k].begin(), cg[k].end());
    }
    an.rg_ptr.assign(n + 2, 0);
    for (int r = 1; r <= n; r++) {
        an.rg_ptr[r+1] = an.rg_ptr[r] + (int)rg[r].size();
        an.rg_idx.insert(an.rg_idx.end(), rg[r].begin(), rg[r].end());
    }

    an.phase_mask.assign(n + 1, 0);
    if (jjDCphase) for (int i = 1; i < dk->nnodes && i <= n; i++) an.phase_mask[i] = phase_node[i];

    // ---- .ic / .nodeset ----
    an.ic_given.assign(n + 1, 0); an.ns_given.assign(n + 1, 0);
    an.ic_val.assign(n + 1, 0.0); an.ns_val.assign(n + 1, 0.0);
    for (int i = 0; i < dk->nic; i++) { an.ic_given[dk->ic_node[i]] = 1; an.ic_val[dk->ic_node[i]] = dk->ic_val[i]; }
    for (int i = 0; i < dk->nnodeset; i++) { an.ns_given[dk->ns_node[i]] = 1; an.ns_val[dk->ns_node[i]] = dk->ns_val[i]; an.hadNodeset = 1; }

    // ---- sTRANAN::init, tranan.cc:94-217 ----
    an.t_stop = dk->tstop; an.t_step = dk->tstep; an.t_start = dk->tstart;
    if (!(an.t_step > 0) || an.t_start < 0 || an.t_stop <= an.t_start) {
        err = "bad .tran parameters"; return B200_E_BADPARM;
    }
    an.maxStep = dk->tmax;
    an.nocut = !(M.n > 0 || Cp.n > 0 || L.n > 0 || J.n > 0 || TJ.n > 0);  // tranan.cc:68-77 DV_TRUNC devices (jj.cc:269)
    if (an.maxStep == 0) {
        an.maxStep = (an.t_stop - an.t_start)/50;
        if (o.oldsteplim) { if (an.t_step > an.maxStep) an.maxStep = an.t_step; }
        else { if (an.t_step < an.maxStep) an.maxStep = 1.001*an.t_step; }
    }
    double delmin = o.delmin;
    if (!(delmin > 0)) {
        if (o.oldsteplim) {
            delmin = an.maxStep*1e-9;
            double T1 = an.t_stop*1e-12;
            if (delmin < T1) delmin = T1;
        }
        else
            delmin = an.maxStep*1e-6;
    }
    an.delmin = delmin;
    an.minBreak = (o.minbreak > 0) ? o.minbreak : delmin*50.0;
    an.npoints_max = (int)floor((an.t_stop - an.t_start)/an.t_step + 0.5) + 3;

    // ---- JJdev::getic (jjic.cc:49-107): start values of a `uic` run and the first time step.
    // ic() has put the .ic / nodeset values into CKTrhs (ckt.cc:908-925); vmax is carried
    // across models as there.  criti > 0 for every junction with a supercurrent (area > 0).
    an.initDelta = 0.0;
    if (J.n && dk->uic) {
        auto rhs0 = [&](int node) { return node > 0 ? (an.ic_given[node] ? an.ic_val[node] : (an.ns_given[node] ? an.ns_val[node] : 0.0)) : 0.0; };
        for (int i = 0; i < J.n; i++) {
            const b200_jj &in = dk->jj[i];
            J.initV[i] = given(in.icv) ? in.icv : rhs0(J.pos[i]) - rhs0(J.neg[i]);
            J.initPhi[i] = given(in.icphi) ? in.icphi : (J.phs[i] > 0 ? rhs0(J.phs[i]) : 0.0);
        }
        const double PHI0 = 2.067833848e-15;                 // circuit.h:123
        double vmax = 0;
        std::vector<char> seen(J.nmodels, 0);
        for (int a = 0; a < J.n; a++) {                      // models in first-instance order
            int m = J.model[a];
            if (seen[m]) continue;
            seen[m] = 1;
            const b200_jj_model &md = J.models[m];
            double vth = md.vdpbak/md.tsaccl;
            if (md.ictype != 0) {
                if (vmax < vth) vmax = vth;
                for (int i = 0; i < J.n; i++)
                    if (J.model[i] == m) { double v = fabs(J.initV[i]); if (vmax < v) vmax = v; }
            }
            if (vmax > 0) {
                double delmax = 0.1*md.tsfact*PHI0/vmax;
                if (an.initDelta == 0.0 || an.initDelta > delmax) an.initDelta = delmax;
            }
        }
    }
    if (TJ.n && dk->uic) {
        // TJMdev::getic, tjmic.cc:45-95 (its own vmax; criti > 0 for every junction)
        auto rhs0 = [&](int node) { return node > 0 ? (an.ic_given[node] ? an.ic_val[node] : (an.ns_given[node] ? an.ns_val[node] : 0.0)) : 0.0; };
        for (int i = 0; i < TJ.n; i++) {
            const b200_tjm &in = dk->tjm[i];
            TJ.initV[i] = given(in.icv) ? in.icv : rhs0(TJ.pos[i]) - rhs0(TJ.neg[i]);
            TJ.initPhi[i] = given(in.icphi) ? in.icphi : (TJ.phs[i] > 0 ? rhs0(TJ.phs[i]) : 0.0);
        }
        const double PHI0 = 2.067833848e-15;
        double vmax = 0;
        std::vector<char> seen(TJ.nmodels, 0);
        for (int a = 0; a < TJ.n; a++) {
            int m = TJ.model[a];
            if (seen[m]) continue;
            seen[m] = 1;
            const b200_tjm_model &md = TJ.models[m];
            if (md.ictype != 0)
                for (int i = 0; i < TJ.n; i++)
                    if (TJ.model[i] == m) {
                        double vth = TJ.vdpbak[i]/md.tsaccl;
                        if (vmax < vth) vmax = vth;
                        double v = fabs(TJ.initV[i]);
                        if (vmax < v) vmax = v;
                    }
            if (vmax > 0) {
                double delmax = 0.1*md.tsfact*PHI0/vmax;
                if (an.initDelta == 0.0 || an.initDelta > delmax) an.initDelta = delmax;
            }
        }
    }

    // ---- breakpoints + source run-time parameters (tranan.cc:330-340, inptran.cc) ----
    BreakTab bt;
    bt.minbrk = an.minBreak;
    bt.set_break(0.0, 0.0);
    bt.set_break(an.t_start, 0.0);
    bt.set_break(an.t_stop, 0.0);
    S.pool.clear();
    double step = an.t_step, finaltime = an.t_stop;
    for (int oi = 0; oi < dk->norder; oi++) {
        if (dk->order_kind[oi] != B200_DEV_SRC) continue;
        int i = dk->order_index[oi];
        const b200_src &in = dk->src[i];
        S.coef_off[i] = (int)S.pool.size();
        S.ncoef[i] = in.ncoef;
        if (in.fn == B200_FN_PULSE) {
            // inptran.cc:1203-1318.  pool: parms[8], cache[2], then the extra offsets
            if (in.ncoef < 1) { err = "pulse needs arguments"; return B200_E_BADPARM; }
            double p[8] = {0,0,0,0,0,0,0,0}, cache[2];
            const double *co = in.coef; int nc = in.ncoef;
            // a bit pattern follows a NaN: length, repeat start, bits (include/b200spice.h)
            int npat = 0; const double *pat = nullptr;
            for (int k = 0; k < in.ncoef; k++)
                if (std::isnan(co[k])) {
                    nc = k;
                    if (k + 2 >= in.ncoef || (int)co[k+1] != in.ncoef - k - 3) { err = "pulse pattern encoding"; return B200_E_BADPARM; }
                    npat = (int)co[k+1]; pat = co + k + 2;
                    break;
                }
            p[0] = co[0];
            p[1] = nc >= 2 ? co[1] : co[0];
            p[2] = nc >= 3 ? co[2] : 0.0;
            p[3] = nc >= 4 ? co[3] : 0.0;
            p[4] = nc >= 5 ? co[4] : 0.0;
            p[5] = nc >= 6 ? co[5] : 0.0;
            p[6] = nc >= 7 ? co[6] : 0.0;
            if (p[3] <= 0.0) p[3] = step;
            if (p[4] > 0.0 && p[5] <= 0.0)
                p[5] = 0.0;
            else {
                if (p[4] <= 0.0) p[4] = step;
                if (p[5] <= 0.0) p[5] = FLT_MAX;
            }
            if (p[6] <= 0.0) p[6] = FLT_MAX;
            else if (p[6] < p[3] + p[5] + p[4]) p[6] = p[3] + p[5] + p[4];
            cache[0] = (p[1] - p[0])/p[3];
            cache[1] = (p[0] - p[1])/p[4];
            int nrep = nc > 7 ? nc - 7 : 0;
            for (int r = -1; r < nrep; r++) {
                double time = (r < 0) ? p[2] : co[7 + r];
                if (p[6] < finaltime) {
                    bt.set_lattice(time, p[6]);
                    time += p[3]; bt.set_lattice(time, p[6]);
                    if (p[5] != 0) { time += p[5]; bt.set_lattice(time, p[6]); }
                    if (p[4] != 0) { time += p[4]; bt.set_lattice(time, p[6]); }
                }
                else {
                    bt.set_break(time, 0.0);
                    time += p[3]; bt.set_break(time, 0.0);
                    if (p[5] != 0) { time += p[5]; bt.set_break(time, 0.0); }
                    if (p[4] != 0) { time += p[4]; bt.set_break(time, 0.0); }
                }
            }
            if (!(p[6] < finaltime)) npat = 0;       // the pattern exists only for a period inside the run (:1305)
            p[7] = npat;                             // pattern: [repeat start, bits...] after the offsets
            S.pool.insert(S.pool.end(), p, p + 8);
            S.pool.insert(S.pool.end(), cache, cache + 2);
            for (int r = 0; r < nrep; r++) S.pool.push_back(co[7 + r]);
            for (int k = 0; k <= npat && npat; k++) S.pool.push_back(pat[k]);
            S.ncoef[i] = nrep;                       // number of extra offsets
        }
        else if (in.fn == B200_FN_PWL) {
            std::vector<double> pts, slopes;
            if (in.ncoef < 1) { err = "pwl needs arguments"; return B200_E_BADPARM; }
            build_pwl(in.coef, in.ncoef, pts, slopes);
            int nv = (int)pts.size()/2;
            for (int k = 0; k < nv; k++) {
                double t = pts[2*k];
                if (t > finaltime) break;
                bt.set_break(t, 0.0);
            }
            S.pool.insert(S.pool.end(), pts.begin(), pts.end());
            S.pool.insert(S.pool.end(), slopes.begin(), slopes.end());
            S.ncoef[i] = nv;                         // number of (t,v) pairs
        }
        else if (in.fn == B200_FN_SIN) {
            double p[8] = {0,0,0,0,0,0,0,0};
            const double *co = in.coef; int nc = in.ncoef;
            if (nc < 1) { err = "sin needs arguments"; return B200_E_BADPARM; }
            p[0] = co[0];
            p[1] = nc >= 2 ? co[1] : 0.0;
            p[2] = nc >= 3 ? co[2] : 0.0;
            p[3] = nc >= 4 ? co[3] : 0.0;
            p[4] = nc >= 5 ? co[4] : 0.0;
            p[5] = nc >= 6 ? co[5] : 0.0;
            if (!p[2]) p[2] = 1.0/finaltime;
            if (p[3] > 0.0) bt.set_break(p[3], 0.0);
            S.pool.insert(S.pool.end(), p, p + 8);
        }
        else if (in.fn == B200_FN_EXP) {
            // IFexpData::setup, inptran.cc:2985-3013: V1 V2 TD1 TAU1 TD2 TAU2
            double p[8] = {0,0,0,0,0,0,0,0};
            const double *co = in.coef; int nc = in.ncoef;
            if (nc < 1) { err = "exp needs arguments"; return B200_E_BADPARM; }
            p[0] = co[0];
            p[1] = nc >= 2 ? co[1] : co[0];
            p[2] = nc >= 3 ? co[2] : 0.0;
            p[3] = nc >= 4 ? co[3] : 0.0;
            p[4] = nc >= 5 ? co[4] : 0.0;
            p[5] = nc >= 6 ? co[5] : 0.0;
            if (!p[5]) p[5] = step;
            if (!p[2]) p[2] = step;
            if (!p[3]) p[3] = step;
            if (!p[4]) p[4] = p[2] + step;
            if (p[2] > 0.0) bt.set_break(p[2], 0.0);
            if (p[4] > p[2]) bt.set_break(p[4], 0.0);
            S.pool.insert(S.pool.end(), p, p + 8);
        }
        else if (in.fn == B200_FN_SFFM || in.fn == B200_FN_AM) {
            // IFsffmData::setup (inptran.cc:3194-3211): VO VA FC MDI FS, FC / FS default 1/finaltime;
            // IFamData::setup (:3373-3388): SA OC MF CF DL, MF default 1/finaltime.  No breakpoints.
            double p[8] = {0,0,0,0,0,0,0,0};
            const double *co = in.coef; int nc = in.ncoef;
            if (nc < 1) { err = "sffm / am needs arguments"; return B200_E_BADPARM; }
            for (int k = 0; k < 5; k++) p[k] = nc > k ? co[k] : 0.0;
            if (in.fn == B200_FN_SFFM) { if (!p[2]) p[2] = 1.0/finaltime; if (!p[4]) p[4] = 1.0/finaltime; }
            else { if (!p[2]) p[2] = 1.0/finaltime; }
            S.pool.insert(S.pool.end(), p, p + 8);
        }
        else if (in.fn == B200_FN_GPULSE) {
            // IFgpulseData::setup, inptran.cc:1672-1730: V1 V2 TD GPW RPT [TD2 ...].  GPW >= 0 is
            // the FWHM, < 0 the old "variance"; no width -> a single-flux-quantum pulse of the
            // given amplitude (or, with no amplitude either, of FWHM = TSTEP).  pool: parms[8], extras
            double p[8] = {0,0,0,0,0,0,0,0};
            const double *co = in.coef; int nc = in.ncoef;
            // a bit pattern follows a NaN: length, repeat start, bits (include/b200spice.h)
            int npat = 0; const double *pat = nullptr;
            for (int k = 0; k < in.ncoef; k++)
                if (std::isnan(co[k])) {
                    nc = k;
                    if (k + 2 >= in.ncoef || (int)co[k+1] != in.ncoef - k - 3) { err = "gpulse pattern encoding"; return B200_E_BADPARM; }
                    npat = (int)co[k+1]; pat = co + k + 2;      // pat[0] = repeat start, then the bits
                    break;
                }
            if (nc < 1) { err = "gpulse needs arguments"; return B200_E_BADPARM; }
            const double twosqrtln2 = 2.0*sqrt(M_LN2), phi0_sqrtpi = JJ_PHI0/sqrt(M_PI);
            p[0] = co[0];
            p[1] = nc >= 2 ? co[1] : co[0];
            p[2] = nc >= 3 ? co[2] : 0.0;
            p[3] = nc >= 4 ? (co[3] >= 0.0 ? co[3]/twosqrtln2 : -co[3]) : 0.0;
            p[4] = nc >= 5 ? co[4] : 0.0;
            if (p[4] < 2*p[3]) p[4] = 0.0;
            if (p[3] == 0.0) {
                if (p[1] != p[0]) p[3] = phi0_sqrtpi/fabs(p[1] - p[0]);
                else p[3] = step/twosqrtln2;
            }
            if (p[1] == p[0]) p[1] = p[0] + phi0_sqrtpi/p[3];
            p[5] = npat; p[6] = npat ? pat[0] : 0.0;             // pattern length and repeat start
            S.pool.insert(S.pool.end(), p, p + 8);
            for (int k = 5; k < nc; k++) S.pool.push_back(co[k]);
            for (int k = 0; k < npat; k++) S.pool.push_back(pat[1 + k]);
            S.ncoef[i] = nc > 5 ? nc - 5 : 0;        // number of extra delays (the pattern bits follow them)
        }
        else if (in.fn == B200_FN_SPULSE) {
            // IFspulseData::setup, inptran.cc:2790-2808: V1 V2 SPER SDEL THETA
            double p[8] = {0,0,0,0,0,0,0,0};
            const double *co = in.coef; int nc = in.ncoef;
            if (nc < 1) { err = "spulse needs arguments"; return B200_E_BADPARM; }
            p[0] = co[0];
            p[1] = nc >= 2 ? co[1] : co[0];
            p[2] = nc >= 3 ? co[2] : 0.0;
            p[3] = nc >= 4 ? co[3] : 0.0;
            p[4] = nc >= 5 ? co[4] : 0.0;
            if (!p[2]) p[2] = finaltime;
            if (p[3] > 0.0) bt.set_break(p[3], 0.0);
            S.pool.insert(S.pool.end(), p, p + 8);
        }
        else if (in.fn != B200_FN_NONE && in.fn != B200_FN_PULSE && in.fn != B200_FN_PWL) {
            err = "source function not built (GAUSS / TABLE / INTERP)";
            return B200_E_UNSUPP;
        }
    }
    an.brks = bt.brks; an.lat_offs = bt.lat_offs; an.lat_per = bt.lat_per;
    an.break0 = bt.breaks[0]; an.break1 = bt.breaks[1];
    return B200_OK;
}

// ---------------------------------------------------------------------------------------
// Static pivot order.  The reference's KLU path fixes the pivot sequence at klu_factor and
// reuses it in every klu_refactor (kluif.cc:569-606); NIiter asks for a new one when the matrix
// changes character -- at the start of the DC operating point and again at the first transient
// step (NISHOULDREORDER: niiter.cc:181, 501-509).  The batch needs sequences that are valid for
// all trials, so it may only use pivots whose acceptability does not depend on the trial's
// values, and it keeps up to TWO of them: one for the DC modes (capacitor and inductor
// companions vanish there) and one for the transient modes (b200_build_programs).
//   (a) entries that are compile-time constants +-1 (source / inductor branch couplings,
//       srcset.cc:158-170, indset.cc:117-131) while no elimination has touched them, and only
//       where they cannot cause growth: every other live entry of the pivot's column is such a
//       constant too (all multipliers are +-1), or the pivot's row has no other live entry
//       (nothing propagates).  An inductor's coupling is such a pivot in DC (its (br,br) entry
//       is zero there) but not in a transient, where (br,br) = -L*ag0 would make multipliers of
//       that size: a chain of inductors multiplies them up (4e12 over 12 stages).
//   (b) diagonal entries that receive conductance stamps: voltage-node diagonals, and in the
//       transient order the branch diagonals of inductors (-req): the transient MNA matrix
//       [G+C*ag0, B; B^T, -L*ag0] is quasi-definite, diagonal pivots are stable in any order.
// Among those, Markowitz's (r-1)(c-1) rule (the criterion of the Sparse 1.3 solver the
// shipped reference runs, xt_base/sparse/spfactor.cc:1218-1560) picks the next pivot; ties go
// to the candidate touched least recently, which yields wide, shallow elimination levels
// (independent pivots first) for the level-scheduled GPU program.
// ---------------------------------------------------------------------------------------
void b200_order_markowitz(Analysis &an, bool tran_order)
{
    const int n = an.n;
    typedef long long i64;
    auto key = [n](int i, int j) { return (i64)i*(i64)(n + 1) + j; };
    std::vector<std::set<int>> rows(n), cols(n);
    // constant: 1 if the entry is a never-updated +-1 constant; 0 otherwise
    std::unordered_map<i64, char> constant;
    // dcnz: the entry is structurally nonzero in the DC modes as well (capacitor and inductor
    // companion conductances vanish there).  The pivot order is frozen for the whole run, so a
    // pivot must be structurally nonzero in DC (hence also in transient).
    std::unordered_map<i64, char> dcnz;
    std::unordered_set<i64> orig_const;          // the +-1 couplings as the deck gives them
    constant.reserve((size_t)an.nnz*4); dcnz.reserve((size_t)an.nnz*4);
    const bool have_dc = !tran_order && (an.dc_active.size() == (size_t)an.nnz);
    for (int j = 0; j < n; j++)
        for (int p = an.Ap[j]; p < an.Ap[j+1]; p++) {
            int i = an.Ai[p];
            rows[i].insert(j); cols[j].insert(i);
            bool is_const = (an.cg_ptr[p+1] == an.cg_ptr[p]) &&
                (an.Ainit[p] == 1.0 || an.Ainit[p] == -1.0);
            constant[key(i, j)] = is_const ? 1 : 0;
            if (is_const) orig_const.insert(key(i, j));
            dcnz[key(i, j)] = have_dc ? an.dc_active[p] : 1;
        }
    auto is_dcnz = [&](int i, int j) { auto it = dcnz.find(key(i, j)); return it != dcnz.end() && it->second; };
    auto is_constant = [&](int i, int j) { auto it = constant.find(key(i, j)); return it != constant.end() && it->second; };
    std::vector<char> ract(n, 1), cact(n, 1);
    std::vector<char> diag_ok(n, 0);
    for (int i = 0; i < n; i++) {
        int p = an.diag[i+1];
        if (p >= 0 && (an.ntype[i+1] == 0 || tran_order || an.jjDCphase) &&   // phase-mode DC: (br,br) = -2*pi*L/Phi0
                (an.cg_ptr[p+1] > an.cg_ptr[p] || an.Ainit[p] != 0.0) &&
                (!have_dc || an.dc_active[p]))
            diag_ok[i] = 1;
    }
    std::vector<long> rstamp(n, 0), cstamp(n, 0);
    an.rowperm.assign(n, -1); an.colperm.assign(n, -1);
    // Markowitz products up to mk_floor count as equal, so that the least-recently-touched rule
    // decides among them: on a chain (a Josephson transmission line, an LC ladder) the two end
    // nodes always have product 0, and following them strictly unzips the chain one pivot per
    // level -- 3000 levels for 1000 junctions, one busy lane of 32.  With the floor the interior
    // nodes (product 4 or less) compete on age: odd-even elimination, ~log n levels, for a fill
    // of about one entry per pivot.
    long mk_floor = 4;
    if (const char *e = getenv("B200_MK_FLOOR")) mk_floor = atol(e);

    // Per-column count of live non-constant entries and per-row count of live entries ("live":
    // structurally nonzero in DC when the DC order is built) -- what the growth-safety rule of a
    // +-1 constant needs; recomputed for the rows / columns a pivot step touches.
    std::vector<int> col_nonconst(n, 0), row_live(n, 0);
    auto recount_col = [&](int j) {
        int c = 0;
        for (int r : cols[j]) if ((!have_dc || is_dcnz(r, j)) && !is_constant(r, j)) c++;
        col_nonconst[j] = c;
    };
    auto recount_row = [&](int i) {
        int c = 0;
        for (int q : rows[i]) if (!have_dc || is_dcnz(i, q)) c++;
        row_live[i] = c;
    };
    for (int i = 0; i < n; i++) { recount_row(i); recount_col(i); }

    // The safe candidates (first choice) live in a heap ordered exactly as the exhaustive scan
    // would rank them: smaller Markowitz product, then older, then constants before diagonals,
    // then row-major position.  Entries are re-pushed whenever their row or column changes and
    // validated when popped (lazy deletion), so a pivot costs the size of what it touches, not
    // a pass over the whole matrix -- the analysis of a 50 000-unknown deck stays in seconds.
    struct Cand { long mk, st; int ncls, i, j; };
    struct Worse {
        bool operator()(const Cand &a, const Cand &b) const {
            if (a.mk != b.mk) return a.mk > b.mk;
            if (a.st != b.st) return a.st > b.st;
            if (a.ncls != b.ncls) return a.ncls > b.ncls;
            if (a.i != b.i) return a.i > b.i;
            return a.j > b.j;
        }
    };
    std::priority_queue<Cand, std::vector<Cand>, Worse> heap;
    // rank of entry (i,j) as a safe candidate; false if it is none
    auto safe_rank = [&](int i, int j, Cand &c) -> bool {
        if (!ract[i] || !cact[j] || !rows[i].count(j) || !is_dcnz(i, j)) return false;
        int cls;
        if (is_constant(i, j)) {
            // growth-safe only if the multipliers are +-1 or nothing propagates
            bool col_const = col_nonconst[j] == 0;
            bool row_single = row_live[i] - 1 == 0;
            if (!col_const && !row_single) return false;
            cls = 2;
        }
        else if (i == j && diag_ok[i]) cls = 1;
        else return false;
        long mk = (long)(rows[i].size() - 1)*(long)(cols[j].size() - 1);
        if (mk < mk_floor) mk = mk_floor;
        c.mk = mk; c.st = std::max(rstamp[i], cstamp[j]); c.ncls = -cls; c.i = i; c.j = j;
        return true;
    };
    auto push_row = [&](int i) { Cand c; for (int j : rows[i]) if (safe_rank(i, j, c)) heap.push(c); };
    auto push_col = [&](int j) { Cand c; for (int i : cols[j]) if (safe_rank(i, j, c)) heap.push(c); };
    for (int i = 0; i < n; i++) push_row(i);

    for (int k = 0; k < n; k++) {
        int bi = -1, bj = -1;
        while (!heap.empty()) {
            Cand c = heap.top(), now;
            heap.pop();
            if (!safe_rank(c.i, c.j, now)) continue;
            if (now.mk != c.mk || now.st != c.st || now.ncls != c.ncls) continue;    // stale: its current rank is in the heap
            bi = c.i; bj = c.j;
            break;
        }
        if (bi < 0) {
            // no safe candidate.  pass 0: a coupling that was +-1 in the deck (an elimination has
            // touched it since: the current-sum row of a node only inductors reach, whose node
            // voltage went with a branch row -- what is left of such an entry is 1 + a ratio of
            // inductances, never a difference); pass 1: any remaining diagonal; pass 2: any entry
            // -- all structurally nonzero in DC; pass 3: any entry at all (exhaustive scans: rare)
            long best = -1, bstamp = 0;
            for (int pass = 0; pass < 4 && bi < 0; pass++) {
                for (int i = 0; i < n; i++) {
                    if (!ract[i]) continue;
                    for (int j : rows[i]) {
                        if (!cact[j]) continue;
                        if (pass < 3 && !is_dcnz(i, j)) continue;
                        if (pass == 0 && !orig_const.count(key(i, j))) continue;
                        if (pass == 1 && i != j) continue;
                        long mk = (long)(rows[i].size() - 1)*(long)(cols[j].size() - 1);
                        if (mk < mk_floor) mk = mk_floor;
                        long st = std::max(rstamp[i], cstamp[j]);
                        if (best < 0 || mk < best || (mk == best && st < bstamp)) { best = mk; bi = i; bj = j; bstamp = st; }
                    }
                }
            }
        }
        if (bi < 0) {
            // structurally singular: finish with whatever is left (refactor will flag it)
            for (int i = 0; i < n && bi < 0; i++) if (ract[i]) bi = i;
            for (int j = 0; j < n && bj < 0; j++) if (cact[j]) bj = j;
        }
        an.rowperm[k] = bi; an.colperm[k] = bj;
        ract[bi] = 0; cact[bj] = 0;
        // eliminate: every active row r with an entry in column bj gets row bi's pattern
        std::vector<int> prow;
        for (int j : rows[bi]) if (cact[j]) prow.push_back(j);
        std::vector<int> pcol;
        for (int r : cols[bj]) if (ract[r]) pcol.push_back(r);
        for (int r : pcol) {
            for (int j : prow) {
                rows[r].insert(j); cols[j].insert(r);
                if (r != bi && j != bj && is_dcnz(r, bj) && is_dcnz(bi, j)) dcnz[key(r, j)] = 1;
                constant[key(r, j)] = 0;            // touched by an update: no longer constant
                cstamp[j] = k + 1;
            }
            rstamp[r] = k + 1;
        }
        // remove pivot row/col from active structure
        for (int j : rows[bi]) cols[j].erase(bi);
        for (int r : cols[bj]) rows[r].erase(bj);
        rows[bi].clear(); cols[bj].clear();
        // what changed: rows pcol, columns prow
        for (int r : pcol) recount_row(r);
        for (int j : prow) recount_col(j);
        for (int r : pcol) push_row(r);
        for (int j : prow) push_col(j);
    }
}

// ---------------------------------------------------------------------------------------
// Numerical pivot test on ONE value set -- what klu_factor / spOrderAndFactor do for the matrix
// they are handed (kluif.cc:569-590, Common.tol; spfactor.cc:257-423 with RelThreshold =
// pivrel, AbsThreshold = pivtol, circuit.h:461-631): a pivot must satisfy
//     |pivot| >= pivabs   and   |pivot| >= pivrel * max|entries left in its column|.
// mode 0: check the order already in an.rowperm/colperm on the values Ax (elimination carried
//         out numerically); returns the first position that fails, or -1 if all pass.
// mode 1: re-order: Markowitz restricted to entries that pass the test (threshold pivoting;
//         ties to the larger |pivot|/colmax), overwriting an.rowperm/colperm; returns -1, or
//         the position at which no acceptable pivot exists (numerically singular).
// ---------------------------------------------------------------------------------------
int b200_order_numeric(Analysis &an, const double *Ax, double pivrel, double pivabs, int mode)
{
    const int n = an.n;
    std::vector<std::map<int,double>> rows(n);          // row -> (col -> value), active part only
    std::vector<std::set<int>> cols(n);
    for (int j = 0; j < n; j++)
        for (int p = an.Ap[j]; p < an.Ap[j+1]; p++) { rows[an.Ai[p]][j] = Ax[p]; cols[j].insert(an.Ai[p]); }
    std::vector<char> ract(n, 1), cact(n, 1);
    if (mode == 1) { an.rowperm.assign(n, -1); an.colperm.assign(n, -1); }
    auto colmax = [&](int j) {
        double m = 0.0;
        for (int r : cols[j]) if (ract[r]) { double v = fabs(rows[r][j]); if (v > m) m = v; }
        return m;
    };
    for (int k = 0; k < n; k++) {
        int bi = -1, bj = -1;
        if (mode == 0) {
            bi = an.rowperm[k]; bj = an.colperm[k];
            auto it = rows[bi].find(bj);
            double pv = it == rows[bi].end() ? 0.0 : fabs(it->second);
            double cm = colmax(bj);
            if (!(pv >= pivabs) || !(pv >= pivrel*cm) || pv == 0.0 || !std::isfinite(pv)) return k;
        }
        else {
            long best = -1; double bratio = 0.0;
            std::vector<double> cm(n, -1.0);
            for (int i = 0; i < n; i++) {
                if (!ract[i]) continue;
                for (auto &e : rows[i]) {
                    int j = e.first;
                    if (!cact[j]) continue;
                    if (cm[j] < 0.0) cm[j] = colmax(j);
                    double pv = fabs(e.second);
                    if (pv == 0.0 || !std::isfinite(pv) || pv < pivabs || pv < pivrel*cm[j]) continue;
                    long rc = 0, cc = 0;
                    for (auto &q : rows[i]) if (cact[q.first]) rc++;
                    for (int r : cols[j]) if (ract[r]) cc++;
                    long mk = (rc - 1)*(cc - 1);
                    double ratio = pv/cm[j];
                    if (best < 0 || mk < best || (mk == best && ratio > bratio)) {
                        best = mk; bi = i; bj = j; bratio = ratio;
                    }
                }
            }
            if (bi < 0) return k;
            an.rowperm[k] = bi; an.colperm[k] = bj;
        }
        ract[bi] = 0; cact[bj] = 0;
        const double pv = rows[bi][bj];
        for (int r : cols[bj]) {
            if (!ract[r]) continue;
            const double l = rows[r][bj]/pv;
            if (l == 0.0) continue;
            for (auto &e : rows[bi]) {
                int j = e.first;
                if (!cact[j]) continue;
                rows[r][j] -= l*e.second;               // creates the fill entry if absent
                cols[j].insert(r);
            }
        }
    }
    return -1;
}

// ---------------------------------------------------------------------------------------
// Symbolic LU for the fixed order, the task program of k_lu and the assembly runs of
// k_assemble (see b200_internal.h).
// ---------------------------------------------------------------------------------------
namespace {

struct HTask {
    int dst = -1, div = -1, level = 0;
    std::vector<std::pair<int,int>> ops;        // (a slot, b slot) in subtraction order
};

} // namespace

int b200_build_lu(Analysis &an)
{
    int n = an.n;
    std::vector<int> rinv(n), cinv(n);
    for (int k = 0; k < n; k++) { rinv[an.rowperm[k]] = k; cinv[an.colperm[k]] = k; }
    // symbolic: column patterns in permuted positions
    an.Lp.assign(n + 1, 0); an.Up.assign(n + 1, 0); an.Li.clear(); an.Ui.clear();
    std::vector<char> mark(n);
    for (int j = 0; j < n; j++) {
        int cj = an.colperm[j];
        std::fill(mark.begin(), mark.end(), 0);
        for (int p = an.Ap[cj]; p < an.Ap[cj+1]; p++) mark[rinv[an.Ai[p]]] = 1;
        for (int k = 0; k < j; k++) {
            if (!mark[k]) continue;
            for (int p = an.Lp[k]; p < an.Lp[k+1]; p++) mark[an.Li[p]] = 1;
        }
        for (int k = 0; k < j; k++) if (mark[k]) an.Ui.push_back(k);
        an.Up[j+1] = (int)an.Ui.size();
        for (int k = j + 1; k < n; k++) if (mark[k]) an.Li.push_back(k);
        an.Lp[j+1] = (int)an.Li.size();
    }
    // value slots: per column [U entries..., diag, L entries...]
    an.slotU.resize(an.Ui.size()); an.slotL.resize(an.Li.size()); an.slotD.resize(n);
    int slot = 0;
    for (int j = 0; j < n; j++) {
        for (int p = an.Up[j]; p < an.Up[j+1]; p++) an.slotU[p] = slot++;
        an.slotD[j] = slot++;
        for (int p = an.Lp[j]; p < an.Lp[j+1]; p++) an.slotL[p] = slot++;
    }
    an.nnzLU = slot;
    an.lu_src.assign(slot, -1);
    an.a2lu.assign(an.nnz, -1);
    {
        std::vector<int> pos2slot(n, -1);
        for (int j = 0; j < n; j++) {
            for (int p = an.Up[j]; p < an.Up[j+1]; p++) pos2slot[an.Ui[p]] = an.slotU[p];
            pos2slot[j] = an.slotD[j];
            for (int p = an.Lp[j]; p < an.Lp[j+1]; p++) pos2slot[an.Li[p]] = an.slotL[p];
            int cj = an.colperm[j];
            for (int p = an.Ap[cj]; p < an.Ap[cj+1]; p++) {
                int s = pos2slot[rinv[an.Ai[p]]];
                an.a2lu[p] = s;
                an.lu_src[s] = p;
            }
        }
    }
    an.yperm_in.resize(n); an.yperm_out.resize(n);
    for (int k = 0; k < n; k++) { an.yperm_in[k] = an.rowperm[k] + 1; an.yperm_out[k] = an.colperm[k] + 1; }

    // ---- tasks, in an order in which every operand's producer precedes its consumers ----
    const int nlu = an.nnzLU;
    auto Y = [&](int k) { return nlu + k; };
    std::vector<HTask> tasks;
    int nscratch = 0;
    const int scratch0 = nlu + n;
    // emit(dst, div, ops): splits long sums (LU_SPLIT / LU_CHUNK rule), recursively
    std::function<void(int,int,std::vector<std::pair<int,int>>&)> emit =
        [&](int dst, int div, std::vector<std::pair<int,int>> &ops) {
        if ((int)ops.size() <= LU_SPLIT) {
            if (ops.empty() && div < 0) return;              // the assembled value stands
            HTask t; t.dst = dst; t.div = div; t.ops = ops;
            tasks.push_back(std::move(t));
            return;
        }
        std::vector<std::pair<int,int>> comb;
        for (size_t c = 0; c < ops.size(); c += LU_CHUNK) {
            size_t e = std::min(ops.size(), c + LU_CHUNK);
            std::vector<std::pair<int,int>> part(ops.begin() + c, ops.begin() + e);
            int ps = scratch0 + nscratch++;
            emit(ps, -1, part);
            comb.push_back({ps, -2});                        // -2: the -1.0 slot, patched below
        }
        emit(dst, div, comb);
    };
    {
        // factor: entry (i,j) = A(i,j) - sum over pivots k < min(i,j), ascending, of L(i,k)*U(k,j);
        // L entries are then divided by the pivot (division, as Sparse and KLU do)
        std::vector<std::vector<std::pair<int,int>>> rowops(n);
        std::vector<int> pos2slot(n, -1);
        for (int j = 0; j < n; j++) {
            for (int p = an.Up[j]; p < an.Up[j+1]; p++) pos2slot[an.Ui[p]] = an.slotU[p];
            pos2slot[j] = an.slotD[j];
            for (int p = an.Lp[j]; p < an.Lp[j+1]; p++) pos2slot[an.Li[p]] = an.slotL[p];
            for (int p = an.Up[j]; p < an.Up[j+1]; p++) {
                int k = an.Ui[p];
                for (int q = an.Lp[k]; q < an.Lp[k+1]; q++)
                    rowops[an.Li[q]].push_back({an.slotL[q], an.slotU[p]});
            }
            for (int p = an.Up[j]; p < an.Up[j+1]; p++) { int i = an.Ui[p]; emit(pos2slot[i], -1, rowops[i]); rowops[i].clear(); }
            emit(an.slotD[j], -1, rowops[j]); rowops[j].clear();
            for (int p = an.Lp[j]; p < an.Lp[j+1]; p++) { int i = an.Li[p]; emit(pos2slot[i], an.slotD[j], rowops[i]); rowops[i].clear(); }
        }
        // forward solve: y(i) = b(i) - sum_k L(i,k)*y(k), ascending k
        std::vector<std::vector<std::pair<int,int>>> rows(n);
        for (int k = 0; k < n; k++)
            for (int q = an.Lp[k]; q < an.Lp[k+1]; q++) rows[an.Li[q]].push_back({an.slotL[q], Y(k)});
        for (int i = 0; i < n; i++) emit(Y(i), -1, rows[i]);
        // backward solve: x(i) = (y(i) - sum_j U(i,j)*x(j)) / U(i,i), descending j
        for (auto &r : rows) r.clear();
        for (int j = n - 1; j >= 0; j--)
            for (int p = an.Up[j+1] - 1; p >= an.Up[j]; p--) rows[an.Ui[p]].push_back({an.slotU[p], Y(j)});
        for (int i = n - 1; i >= 0; i--) emit(Y(i), an.slotD[i], rows[i]);
    }
    an.nscratch = nscratch;
    an.slot_neg1 = scratch0 + nscratch;
    an.slot_zero = an.slot_neg1 + 1;
    an.img = (an.slot_zero + 1 + 1) & ~1;                   // 16-byte granular rows (bulk copy)
    for (auto &t : tasks) for (auto &op : t.ops) if (op.second == -2) op.second = an.slot_neg1;

    // ---- levels: RAW on every operand, and WAR on a y slot the backward solve overwrites ----
    {
        std::vector<int> wlevel(an.img, -1), rlevel(an.img, -1);
        for (auto &t : tasks) {
            int lv = std::max(wlevel[t.dst], rlevel[t.dst]);
            for (auto &op : t.ops) lv = std::max(lv, std::max(wlevel[op.first], wlevel[op.second]));
            if (t.div >= 0) lv = std::max(lv, wlevel[t.div]);
            t.level = lv + 1;
            for (auto &op : t.ops) {
                rlevel[op.first] = std::max(rlevel[op.first], t.level);
                rlevel[op.second] = std::max(rlevel[op.second], t.level);
            }
            if (t.div >= 0) rlevel[t.div] = std::max(rlevel[t.div], t.level);
            wlevel[t.dst] = t.level;
        }
    }
    int nlev = 0;
    for (auto &t : tasks) nlev = std::max(nlev, t.level + 1);
    an.prog_levels = nlev;

    // ---- rounds ----
    an.prog_rows.clear(); an.prog_round.clear();
    an.prog_steps = 0; an.prog_ops = 0; an.prog_tasks = (int)tasks.size();
    std::vector<std::vector<int>> by_level(nlev);
    for (int t = 0; t < (int)tasks.size(); t++) by_level[tasks[t].level].push_back(t);
    // rows as (first, second) pairs; packed below once the encoding is known.  second = -1 in a
    // header: no division
    std::vector<std::pair<int,int>> prows;
    for (int l = 0; l < nlev; l++) {
        auto &v = by_level[l];
        std::stable_sort(v.begin(), v.end(), [&](int a, int b) { return tasks[a].ops.size() > tasks[b].ops.size(); });
        for (size_t off = 0; off < v.size(); off += LU_LANES) {
            size_t cnt = std::min<size_t>(LU_LANES, v.size() - off);
            size_t nsteps = tasks[v[off]].ops.size();
            an.prog_round.push_back((unsigned short)(nsteps | ((off == 0 && l > 0) ? LU_RSYNC : 0)));
            size_t row0 = prows.size();
            // idle lanes run the exact no-op  S[zero] = S[zero] - S[zero]*S[zero]
            prows.resize(row0 + LU_LANES, std::make_pair(an.slot_zero, -1));
            prows.resize(row0 + (1 + nsteps)*LU_LANES, std::make_pair(an.slot_zero, an.slot_zero));
            for (size_t q = 0; q < cnt; q++) {
                const HTask &t = tasks[v[off + q]];
                prows[row0 + q] = std::make_pair(t.dst, t.div >= 0 ? t.div : -1);
                for (size_t s = 0; s < t.ops.size(); s++)
                    prows[row0 + (1 + s)*LU_LANES + q] = t.ops[s];
                an.prog_ops += (long long)t.ops.size();
            }
            an.prog_steps += (long long)nsteps;
        }
    }
    // ---- encoding: narrow (16-bit slots, image staged in one warp's shared-memory window) while
    // it fits, wide otherwise ----
    {
        const size_t rnd = ((an.prog_round.size() + 7) & ~(size_t)7)*sizeof(unsigned short);
        // Narrow while at least LU_MIN_RESIDENT images fit one SM's window: k_lu keeps one warp per
        // resident image, and below that it leaves the SM idle -- measured: the 1 000-junction
        // line (96 KB image, 2 resident) runs 1.7x faster through k_lu_wide, RO-501 (32 KB, 5-7
        // resident) 10 % slower.
        int min_resident = LU_MIN_RESIDENT;
        if (const char *e = getenv("B200_LU_MIN_RESIDENT")) min_resident = atoi(e);   // measurement override
        an.wide = (an.img > 0xfff0 || (size_t)an.img*8*min_resident + rnd > LU_SMEM_BUDGET) ? 1 : 0;
        if (const char *e = getenv("B200_LU_WIDE")) an.wide = (*e && *e != '0') ? 1 : (an.img > 0xfff0 || (size_t)an.img*8 + rnd > LU_SMEM_BUDGET);    // measurement / test override
        size_t row = 0;
        if (an.wide) an.prog_rows.resize(prows.size()*2);
        else an.prog_rows.resize(prows.size());
        for (size_t r = 0; r < an.prog_round.size(); r++) {
            const size_t nsteps = an.prog_round[r] & 0x7fff;
            for (size_t q = 0; q < (1 + nsteps)*LU_LANES; q++) {
                const std::pair<int,int> &w = prows[row*LU_LANES + q];
                const bool header = q < LU_LANES;
                if (an.wide) {
                    an.prog_rows[(row*LU_LANES + q)*2] = (unsigned)w.first;
                    an.prog_rows[(row*LU_LANES + q)*2 + 1] = (header && w.second < 0) ? LU_NONE_W : (unsigned)w.second;
                }
                else
                    an.prog_rows[row*LU_LANES + q] = (unsigned)w.first |
                        (((header && w.second < 0) ? LU_NONE : (unsigned)w.second) << 16);
            }
            row += 1 + nsteps;
        }
        // round -> first row, level -> first round (a CTA's warps share out the rounds of a level)
        an.round_row.clear(); an.level_ptr.clear();
        row = 0;
        for (size_t r = 0; r < an.prog_round.size(); r++) {
            if (r == 0 || (an.prog_round[r] & LU_RSYNC)) an.level_ptr.push_back((int)r);
            an.round_row.push_back((int)row);
            row += 1 + (an.prog_round[r] & 0x7fff);
        }
        an.level_ptr.push_back((int)an.prog_round.size());
    }

    // ---- assembly runs: contributions renumbered in ASSEMBLY ORDER (all terms of one image
    // slot consecutive, in the reference's load order; slots in image order) ----
    std::vector<int> newc(an.ncslot, -1), newr(an.nrslot, -1);
    an.const_idx.clear(); an.const_val.clear();
    an.ainit_cslot.assign(an.nnz, -1);
    an.runs.assign(an.img, AsmRun{0, 0, 0, 0});
    std::vector<int> p2node(an.nnz, 0);
    for (int i = 1; i <= n; i++) if (an.diag[i] >= 0) p2node[an.diag[i]] = i;
    int next = 0;
    for (int s = 0; s < nlu; s++) {
        int p = an.lu_src[s];
        AsmRun &r = an.runs[s];
        r.pos = next;
        if (p >= 0) {
            int dn = p2node[p];
            if (dn) r.info = dn | ASM_DIAG | (an.ntype[dn] == 0 ? ASM_VNODE : 0);
            // a nonzero preloaded constant (Ainit) is the first term of the run: it lives in a
            // contribution slot of its own that is written once at batch creation
            // (0.0 + Ainit == Ainit exactly, so the sum is the reference's Ainit + c1 + c2 ...)
            if (an.Ainit[p] != 0.0) { an.ainit_cslot[p] = next; an.const_idx.push_back(next++); an.const_val.push_back(an.Ainit[p]); r.len++; }
            for (int q = an.cg_ptr[p]; q < an.cg_ptr[p+1]; q++) { newc[an.cg_idx[q]] = next++; r.len++; }
        }
    }
    for (int k = 0; k < n; k++) {
        int row = an.yperm_in[k];
        AsmRun &r = an.runs[nlu + k];
        r.pos = next; r.info = row | ASM_ROW;
        for (int q = an.rg_ptr[row]; q < an.rg_ptr[row+1]; q++) { newr[an.rg_idx[q]] = next++; r.len++; }
    }
    for (int s = scratch0; s < an.img; s++) an.runs[s].pos = next;
    an.runs[an.slot_neg1].len = 1;
    an.const_idx.push_back(next++); an.const_val.push_back(-1.0);
    an.NC = next;
    an.long_chunks.clear();
    for (int s = 0; s < an.img; s++) {
        AsmRun &r = an.runs[s];
        if (r.len <= ASM_LONG) continue;
        const int first = (int)an.long_chunks.size();
        for (int c = 0; c < r.len; c += ASM_LCHUNK)
            an.long_chunks.push_back(AsmChunk{r.pos + c, std::min(ASM_LCHUNK, r.len - c)});
        r.pos = an.NC + first;
        r.len = (int)an.long_chunks.size() - first;
    }
    {
        int ntiles = (an.img + ASM_TILE_SLOTS - 1)/ASM_TILE_SLOTS;
        std::vector<int> terms(ntiles, 0);
        for (int s = 0; s < an.img; s++) terms[s/ASM_TILE_SLOTS] += an.runs[s].len;
        an.tile_order.resize(ntiles);
        for (int i = 0; i < ntiles; i++) an.tile_order[i] = i;
        std::stable_sort(an.tile_order.begin(), an.tile_order.end(), [&](int a, int b) { return terms[a] > terms[b]; });
    }
    for (auto &c : an.cg_idx) c = newc[c];
    for (auto &c : an.rg_idx) c = newr[c];
    auto remap = [](std::vector<int> &v, const std::vector<int> &m) { for (auto &x : v) if (x >= 0) x = m[x]; };
    remap(an.mos.cslot, newc); remap(an.mos.rslot, newr);
    remap(an.cap.cslot, newc); remap(an.cap.rslot, newr);
    remap(an.src.rslot, newr);
    remap(an.ind.cslot, newc); remap(an.ind.rslot, newr);
    remap(an.mut.cslot, newc);
    remap(an.jj.cslot, newc); remap(an.jj.rslot, newr);
    remap(an.tjm.cslot, newc); remap(an.tjm.rslot, newr);
    remap(an.rph_cslot, newc);
    return B200_OK;
}

// ---------------------------------------------------------------------------------------
// One or two solve programs (see the comment above b200_order_markowitz).
// ---------------------------------------------------------------------------------------
int b200_build_programs(Analysis &an, bool uic)
{
    an.has_tranp = false;
    an.primary_class = -1;
    if (uic) {
        // no solve ever happens in a DC mode (tranan.cc:342-350 loads once, without NIiter)
        b200_order_markowitz(an, true);
        return b200_build_lu(an);
    }
    Analysis tr = an;                            // before build_lu renumbers the gather lists
    b200_order_markowitz(an, false);
    b200_order_markowitz(tr, true);
    if (tr.rowperm == an.rowperm && tr.colperm == an.colperm)
        return b200_build_lu(an);
    int rc = b200_build_lu(an);
    if (rc) return rc;
    rc = b200_build_lu(tr);
    if (rc) return rc;
    // the transient program's runs in the primary's contribution numbering
    std::vector<int> krow(an.n + 1, -1);
    for (int k = 0; k < an.n; k++) krow[an.yperm_in[k]] = k;
    for (int s = 0; s < tr.nnzLU; s++) {
        int p = tr.lu_src[s];
        if (p < 0) continue;
        const AsmRun &r1 = an.runs[an.a2lu[p]];
        if (r1.len != tr.runs[s].len) return B200_E_PANIC;
        tr.runs[s].pos = r1.pos;
    }
    for (int k = 0; k < tr.n; k++) {
        const AsmRun &r1 = an.runs[an.nnzLU + krow[tr.yperm_in[k]]];
        if (r1.len != tr.runs[tr.nnzLU + k].len) return B200_E_PANIC;
        tr.runs[tr.nnzLU + k].pos = r1.pos;
    }
    tr.runs[tr.slot_neg1].pos = an.runs[an.slot_neg1].pos;
    LuProgram &P = an.tranp;
    P.rowperm = tr.rowperm; P.colperm = tr.colperm; P.yperm_out = tr.yperm_out;
    P.tile_order = tr.tile_order; P.runs = tr.runs; P.prog_rows = tr.prog_rows;
    P.prog_round = tr.prog_round; P.nnzLU = tr.nnzLU; P.img = tr.img; P.wide = tr.wide;
    P.round_row = tr.round_row; P.level_ptr = tr.level_ptr;
    P.prog_levels = tr.prog_levels; P.prog_tasks = tr.prog_tasks;
    P.prog_steps = tr.prog_steps; P.prog_ops = tr.prog_ops;
    an.has_tranp = true;
    an.primary_class = 1;
    return B200_OK;
}

// ---------------------------------------------------------------------------------------
// Host-only view of the once-per-batch analysis (no CUDA call): sizes of the pattern, the LU
// and the compiled programs.  Used by the CPU tests of the host logic and by profiles/.
// ---------------------------------------------------------------------------------------
extern "C" int b200_deck_analysis_stats(const b200_deck *deck, long long *out, int nout)
{
    if (!deck || !out) return B200_E_BADPARM;
    Analysis an;
    std::string err;
    int rc = b200_analyze(deck, an, err);
    if (rc) { fprintf(stderr, "b200spice: %s\n", err.c_str()); return rc; }
    rc = b200_build_programs(an, deck->uic != 0);
    if (rc) return rc;
    long long v[B200_AS_NSTATS] = {0};
    v[B200_AS_N] = an.n; v[B200_AS_NNZ] = an.nnz; v[B200_AS_NNZLU] = an.nnzLU; v[B200_AS_NC] = an.NC;
    v[B200_AS_NSTATES] = an.numStates;
    v[B200_AS_IMG] = an.img; v[B200_AS_SCRATCH] = an.nscratch;
    v[B200_AS_TASKS] = an.prog_tasks; v[B200_AS_OPS] = an.prog_ops;
    v[B200_AS_LEVELS] = an.prog_levels; v[B200_AS_ROUNDS] = (long long)an.prog_round.size();
    v[B200_AS_STEPS] = an.prog_steps;
    v[B200_AS_PROG_BYTES] = (long long)(an.prog_rows.size()*sizeof(unsigned) + an.prog_round.size()*sizeof(unsigned short));
    { int m = 0; for (auto &r : an.runs) m = std::max(m, r.len); v[B200_AS_LONGEST_RUN] = m; }
    v[B200_AS_LU_SPLIT] = LU_SPLIT; v[B200_AS_LU_CHUNK] = LU_CHUNK;
    v[B200_AS_WIDE] = an.wide;
    for (int i = 0; i < nout && i < B200_AS_NSTATS; i++) out[i] = v[i];
    return B200_OK;
}

// Host-only: the CSC pattern and the static pivot order of a deck (caller-sized buffers:
// Ap[n+1], Ai[nnz], rowperm[n], colperm[n]; sizes from b200_deck_analysis_stats).
// Host-only: the pivot order of the DC modes (which = 0) or of the transient modes (which = 1)
// as b200_build_programs chooses them; *nprograms = 1 or 2.
extern "C" int b200_deck_analysis_order(const b200_deck *deck, int which, int *rowperm, int *colperm,
    int *nprograms)
{
    if (!deck) return B200_E_BADPARM;
    Analysis an;
    std::string err;
    int rc = b200_analyze(deck, an, err);
    if (rc) { fprintf(stderr, "b200spice: %s\n", err.c_str()); return rc; }
    rc = b200_build_programs(an, deck->uic != 0);
    if (rc) return rc;
    const std::vector<int> &rp = (which && an.has_tranp) ? an.tranp.rowperm : an.rowperm;
    const std::vector<int> &cp = (which && an.has_tranp) ? an.tranp.colperm : an.colperm;
    if (rowperm) std::copy(rp.begin(), rp.end(), rowperm);
    if (colperm) std::copy(cp.begin(), cp.end(), colperm);
    if (nprograms) *nprograms = an.has_tranp ? 2 : 1;
    return B200_OK;
}

extern "C" int b200_deck_analysis_arrays(const b200_deck *deck, int *Ap, int *Ai, int *rowperm, int *colperm)
{
    if (!deck) return B200_E_BADPARM;
    Analysis an;
    std::string err;
    int rc = b200_analyze(deck, an, err);
    if (rc) { fprintf(stderr, "b200spice: %s\n", err.c_str()); return rc; }
    b200_order_markowitz(an, deck->uic != 0);
    if (Ap) std::copy(an.Ap.begin(), an.Ap.end(), Ap);
    if (Ai) std::copy(an.Ai.begin(), an.Ai.end(), Ai);
    if (rowperm) std::copy(an.rowperm.begin(), an.rowperm.end(), rowperm);
    if (colperm) std::copy(an.colperm.begin(), an.colperm.end(), colperm);
    return B200_OK;
}

// ---------------------------------------------------------------------------------------
// Test hook for the host logic (CPU tests, no CUDA call): executes the compiled assembly runs
// and task program of a CSC matrix the way k_assemble / k_lu do -- round by round, lanes in
// turn -- on ONE value set, so the schedule (levels, split sums, padding, slot numbering) can
// be checked against the oracle's static LU bit for bit without a GPU.  Not a solver: nothing
// in the product path calls it.
// ---------------------------------------------------------------------------------------
extern "C" int b200_debug_program_check(int n, const int *Ap, const int *Ai, const double *Ax,
    const double *b, double *x, int *rowperm, int *colperm, long long *stats, int nstats)
{
    if (n <= 0 || !Ap || !Ai || !Ax || !b || !x) return B200_E_BADPARM;
    Analysis an;
    an.n = n; an.nnz = Ap[n];
    an.Ap.assign(Ap, Ap + n + 1); an.Ai.assign(Ai, Ai + an.nnz);
    an.ntype.assign(n + 1, 0);
    an.diag.assign(n + 1, -1);
    for (int j = 0; j < n; j++)
        for (int p = Ap[j]; p < Ap[j+1]; p++) if (Ai[p] == j) an.diag[j+1] = p;
    an.Ainit.assign(Ax, Ax + an.nnz);
    an.cg_ptr.assign(an.nnz + 1, 0);
    an.rg_ptr.assign(n + 2, 0);
    // as b200_klu_analyze: entries that are exactly zero in this matrix are not live, and the
    // static order stands only if it passes the numerical pivot test on these values
    an.dc_active.assign(an.nnz, 0);
    for (int p = 0; p < an.nnz; p++) an.dc_active[p] = (Ax[p] != 0.0);
    b200_order_markowitz(an);
    if (b200_order_numeric(an, Ax, 1e-3, 1e-13, 0) >= 0 && b200_order_numeric(an, Ax, 1e-3, 1e-13, 1) >= 0)
        return B200_E_SINGULAR;
    an.Ainit.assign(an.nnz, 0.0);
    an.cg_idx.resize(an.nnz);
    for (int p = 0; p < an.nnz; p++) { an.cg_ptr[p] = p; an.cg_idx[p] = p; }
    an.cg_ptr[an.nnz] = an.nnz;
    an.rg_idx.resize(n);
    an.rg_ptr[0] = 0; an.rg_ptr[1] = 0;
    for (int r = 1; r <= n; r++) { an.rg_idx[r-1] = r - 1; an.rg_ptr[r+1] = r; }
    an.ncslot = an.nnz; an.nrslot = n;
    int rc = b200_build_lu(an);
    if (rc) return rc;
    // contributions in assembly order
    std::vector<double> contrib(an.NC + an.long_chunks.size(), 0.0);
    for (int p = 0; p < an.nnz; p++) contrib[an.cg_idx[p]] = Ax[p];
    for (int r = 0; r < n; r++) contrib[an.rg_idx[r]] = b[r];
    for (size_t i = 0; i < an.const_idx.size(); i++) contrib[an.const_idx[i]] = an.const_val[i];
    for (size_t c = 0; c < an.long_chunks.size(); c++) {
        double v = 0.0;
        for (int k = 0; k < an.long_chunks[c].len; k++) v += contrib[an.long_chunks[c].pos + k];
        contrib[an.NC + c] = v;
    }
    std::vector<double> S(an.img, 0.0);
    for (int s = 0; s < an.img; s++) {
        double v = 0.0;
        for (int k = 0; k < an.runs[s].len; k++) v += contrib[an.runs[s].pos + k];
        S[s] = v;
    }
    int bad = 0;
    size_t row = 0;
    for (size_t r = 0; r < an.prog_round.size(); r++) {
        int nsteps = an.prog_round[r] & 0x7fff;
        for (int e = 0; e < LU_LANES; e++) {
            unsigned dst, div;
            if (an.wide) { dst = an.prog_rows[(row*LU_LANES + e)*2]; div = an.prog_rows[(row*LU_LANES + e)*2 + 1]; }
            else { unsigned h = an.prog_rows[row*LU_LANES + e]; dst = h & 0xffffu; div = h >> 16; }
            double acc = S[dst];
            for (int st = 0; st < nsteps; st++) {
                if (an.wide) {
                    const unsigned *w = &an.prog_rows[((row + 1 + st)*LU_LANES + e)*2];
                    acc -= S[w[0]]*S[w[1]];
                }
                else {
                    unsigned w = an.prog_rows[(row + 1 + st)*LU_LANES + e];
                    acc -= S[w & 0xffffu]*S[w >> 16];
                }
            }
            if (div != (an.wide ? LU_NONE_W : LU_NONE)) {
                double piv = S[div];
                if (piv == 0.0 || !std::isfinite(piv)) bad = 1;
                acc = acc/piv;
            }
            S[dst] = acc;
        }
        row += 1 + nsteps;
    }
    for (int k = 0; k < n; k++) x[an.yperm_out[k] - 1] = S[an.nnzLU + k];
    if (rowperm) std::copy(an.rowperm.begin(), an.rowperm.end(), rowperm);
    if (colperm) std::copy(an.colperm.begin(), an.colperm.end(), colperm);
    if (stats) {
        long long v[8] = { an.nnzLU, an.img, an.nscratch, an.prog_tasks, an.prog_ops, an.prog_levels,
            (long long)an.prog_round.size(), an.prog_steps };
        for (int i = 0; i < nstats && i < 8; i++) stats[i] = v[i];
    }
    return bad ? B200_E_SINGULAR : B200_OK;
}
