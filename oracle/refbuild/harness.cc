// TEST INFRASTRUCTURE ONLY (oracle/_ref).  Drives the UNMODIFIED reference exactly the way
// its own front end does (fte/source.cc:2197-2262 sFtCirc::newCKT, fte/simulate.cc:551-620
// sFtCirc::run): deck lines -> IP.parseDeck -> sTASK::setOptions -> sCKT::doTask(true),
// with output points recorded from OP.appendData.  One call = one trial, as the
// reference's Monte-Carlo loop runs them (fte/check.cc:1649-1733, serial, deck re-parsed).

#include "config.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "simulator.h"
#include "input.h"
#include "inpline.h"
#include "circuit.h"
#include "optdefs.h"
#include "trandefs.h"
#include "output.h"
#include "sparse/spmatrix.h"

#include "miscutil/random.h"
#ifdef USE_KLU
#include "klumatrix.h"
#include "kluif.h"
#endif
#include "harness.h"
#include "harvest.h"

int wrsref_verbose = 0;

// set by wrsref_run_deck_harvest: called with the live, set-up circuit after doTask
static int (*after_task_hook)(sCKT*) = 0;

namespace {
    struct Recorder {
        int numnames = 0;
        std::vector<double> ref;
        std::vector<double> vals;   // [pt][numnames]
    };
    Recorder rec;

    struct RunResult {
        int error = 0;
        int size = 0;
        int accepted = 0, rejected = 0, numiter = 0, timepts = 0, traniter = 0,
            itercut = 0, trapcut = 0;
        double loadtime = 0, decomptime = 0, solvetime = 0, reordertime = 0, cvchktime = 0,
            trantime = 0, tottime = 0;
        std::vector<double> names_dummy;
    };
    RunResult last;
    int last_hook_rc = 0;
    std::vector<char> last_names;   // '\0'-separated node names in output column order
    std::vector<char> last_order;   // '\0'-separated instance names in sCKT::load order
}

extern "C" void wrsref_begin_plot(int numnames)
{
    rec.numnames = numnames;
    rec.ref.clear();
    rec.vals.clear();
}

extern "C" void wrsref_append(double ref, const double *vals, int n)
{
    rec.ref.push_back(ref);
    rec.vals.insert(rec.vals.end(), vals, vals + n);
}

extern "C" {

void wrsref_set_verbose(int v) { wrsref_verbose = v; }
// the reference's own per-NIiter trace on stderr (sCKT::CKTstepDebug, niiter.cc:541-546)
void wrsref_set_stepdebug(int v) { sCKT::CKTstepDebug = v; }

// the reference's own deviates (sRnd, xt_base/miscutil/random.cc): pins the dispatcher's
// restatement of the Monte-Carlo random stream
void wrsref_gauss(int seed, long n, double *out)
{
    sRnd r;
    r.seed(seed);
    for (long i = 0; i < n; i++) out[i] = r.gauss();
}
void wrsref_uniform(int seed, long n, double *out)
{
    sRnd r;
    r.seed(seed);
    for (long i = 0; i < n; i++) out[i] = r.random();
}
void wrsref_set_trantrace(int v) { Sp.SetTranTrace(v); }

// KLU plug-in: what the front end does once at start-up (KLUif::find_klu, kluif.cc:117-266,
// looks at $XT_KLU_PATH).  Returns 1 if a plug-in was accepted (libwrsref_klu.so only).
int wrsref_find_klu(void)
{
#ifdef USE_KLU
    if (!klu_if.is_ok())
        klu_if.find_klu();
    return (klu_if.is_ok() ? 1 : 0);
#else
    return (0);
#endif
}

// Run one deck.  deck_text: newline-separated SPICE lines, first line = title, must
// contain one .tran (or .op) line; ".options" lines are honoured the way
// fte/source.cc does (collected into an sOPTIONS handed to sTASK::setOptions).
// Returns the reference's error code from doTask (0 = OK).
int wrsref_run_deck(const char *deck_text)
{
    rec = Recorder();
    last = RunResult();
    last_names.clear();
    last_order.clear();

    // Split into sLine list.
    sLine *head = 0, *tail = 0;
    const char *s = deck_text;
    int lnum = 0;
    while (*s) {
        const char *e = strchr(s, '\n');
        size_t len = e ? (size_t)(e - s) : strlen(s);
        char *buf = new char[len + 1];
        memcpy(buf, s, len);
        buf[len] = 0;
        // strip trailing CR/space
        while (len && (buf[len-1] == '\r' || buf[len-1] == ' '))
            buf[--len] = 0;
        sLine *l = new sLine;
        l->set_line(buf);
        l->set_line_num(++lnum);
        delete [] buf;
        if (!head)
            head = tail = l;
        else {
            tail->set_next(l);
            tail = l;
        }
        if (!e)
            break;
        s = e + 1;
    }
    if (!head || !head->next()) {
        sLine::destroy(head);
        return (-1);
    }

    // .options lines -> sOPTIONS (what fte/source.cc builds as ci_defOpt).
    sOPTIONS *opts = new sOPTIONS;
    for (sLine *l = head->next(); l; l = l->next()) {
        const char *t = l->line();
        while (*t == ' ' || *t == '\t')
            t++;
        if (lstring::cimatch(".options", t) || lstring::cimatch(".option", t) ||
                lstring::cimatch(".opt", t)) {
            char *tok = IP.getTok(&t, true);
            delete [] tok;
            OPTinfo.parse(l, 0, 0, &t, (sTASK*)opts);
            l->comment_out();
        }
    }

    sFtCirc *circ = new sFtCirc;
    sCKT *ckt = new sCKT;
    ckt->CKTbackPtr = circ;
    IFuid taskUid;
    int ret = ckt->newUid(&taskUid, 0, "default", UID_TASK);
    if (ret) {
        delete ckt;
        delete circ;
        sLine::destroy(head);
        return (ret);
    }
    sTASK *task = new sTASK(taskUid);
    IP.parseDeck(ckt, head->next(), task, false);

    int nerr = 0;
    for (sLine *l = head; l; l = l->next()) {
        if (l->error()) {
            nerr++;
            if (wrsref_verbose)
                fprintf(stderr, "wrsref: line %d: %s\n  %s\n", l->line_num(), l->line(),
                    l->error());
        }
    }
    if (ckt->CKTnogo) {
        delete ckt;
        delete circ;
        sLine::destroy(head);
        return (-2);
    }

    task->setOptions(opts, false);
    ckt->setTask(task);

    int err = ckt->doTask(true);

    last.error = err;
    if (ckt->CKTmatrix)
        last.size = ckt->CKTmatrix->spGetSize(1);
    sSTATS *st = ckt->CKTstat;
    if (st) {
        last.accepted = st->STATaccepted;
        last.rejected = st->STATrejected;
        last.numiter = st->STATnumIter;
        last.timepts = st->STATtimePts;
        last.traniter = st->STATtranIter;
        last.itercut = st->STATtranIterCut;
        last.trapcut = st->STATtranTrapCut;
        last.loadtime = st->STATloadTime;
        last.decomptime = st->STATdecompTime;
        last.solvetime = st->STATsolveTime;
        last.reordertime = st->STATreorderTime;
        last.cvchktime = st->STATcvChkTime;
        last.trantime = st->STATtranTime;
        last.tottime = st->STATtotAnalTime;
    }
    // the hooks walk the set-up circuit: only after a task that got through its setup
    if (after_task_hook && (err == OK || ckt->CKTstat->STATnumIter > 0))
        last_hook_rc = (*after_task_hook)(ckt);

    // Output column names.
    {
        const sCKTnode *node = ckt->CKTnodeTab.find(1);
        for ( ; node; node = ckt->CKTnodeTab.nextNode(node)) {
            const char *nm = (const char*)node->name();
            last_names.insert(last_names.end(), nm, nm + strlen(nm) + 1);
        }
    }

    // Instance names in the order sCKT::load visits them (ckt.cc:996-1005, 1146-1161).
    {
        int muttype = -1;
        sGENmodel *mutm = 0;
        ckt->typelook("mutual", &mutm);
        if (mutm) {
            muttype = mutm->GENmodType;
            for (sGENmodel *dm = mutm; dm; dm = dm->GENnextModel)
                for (sGENinstance *d = dm->GENinstances; d; d = d->GENnextInstance) {
                    const char *nm = (const char*)d->GENname;
                    last_order.insert(last_order.end(), nm, nm + strlen(nm) + 1);
                }
        }
        sCKTmodGen mgen(ckt->CKTmodels);
        for (sGENmodel *m = mgen.next(); m; m = mgen.next()) {
            if (m->GENmodType == muttype)
                continue;
            for (sGENmodel *dm = m; dm; dm = dm->GENnextModel)
                for (sGENinstance *d = dm->GENinstances; d; d = d->GENnextInstance) {
                    const char *nm = (const char*)d->GENname;
                    last_order.insert(last_order.end(), nm, nm + strlen(nm) + 1);
                }
        }
    }

    delete ckt;
    delete circ;
    delete opts;
    sLine::destroy(head);
    return (err);
}

// ---- shim/harvest.cc driven from the reference's own parser and set-up circuit ----
// Runs the deck through doTask (the reference's own result stays readable through the
// wrsref_get_* calls), then hands the SAME sCKT to b200_harvest_run: harvest -> C ABI ->
// batched run on the GPU.  Results of the batch are kept until the next call.
namespace {
    struct HarvestArgs {
        const char *libpath; int ntrials, nparams, op_only;
        const int *param_id; const char *const *inst_name; const double *const *values;
    } hargs;
    struct HarvestOut {
        int rc = 0, n = 0, npoints = 0;
        std::vector<int> counters; std::vector<double> times, waves;
        long long stats[B200_AS_NSTATS];
    } hout;

    int harvest_hook(sCKT *ckt)
    {
        hout = HarvestOut();
        if (hargs.ntrials <= 0) {       // host-only: harvest + the engine's analysis sizes
            hout.rc = b200_harvest_stats(ckt, hargs.libpath, hout.stats, B200_AS_NSTATS);
            return (hout.rc);
        }
        hout.counters.assign((size_t)hargs.ntrials*B200_NCOUNTERS, 0);
        double *t = 0, *w = 0;
        hout.rc = b200_harvest_run(ckt, hargs.libpath, hargs.ntrials, 0, hargs.nparams, hargs.param_id,
            hargs.inst_name, hargs.values, hout.counters.data(), &hout.n, &hout.npoints, &t, &w,
            hargs.op_only);
        if (!hout.rc) {
            hout.times.assign(t, t + hout.npoints);
            hout.waves.assign(w, w + (size_t)hargs.ntrials*hout.npoints*hout.n);
        }
        free(t); free(w);
        return (hout.rc);
    }
}

int wrsref_run_deck_harvest(const char *deck_text, const char *libpath, int ntrials, int nparams,
    const int *param_id, const char *const *inst_name, const double *const *values, int op_only)
{
    hargs = { libpath, ntrials, nparams, op_only, param_id, inst_name, values };
    after_task_hook = harvest_hook;
    int err = wrsref_run_deck(deck_text);
    after_task_hook = 0;
    return (err);
}

void wrsref_harvest_stats(long long *out) { memcpy(out, hout.stats, sizeof(hout.stats)); }
void wrsref_harvest_last_run(double *ms_total, long long *ticks) { b200_harvest_last_stats(ms_total, ticks); }
void wrsref_harvest_set_saves(int n, const int *unknown) { b200_harvest_set_saves(n, unknown); }
// the harvested deck through the plain-C oracle (decks only the reference's setup can make)
namespace {
    struct OracleArgs { const char *path; int nparams; const int *param_id; const char *const *inst_name; const double *value; } oargs;
    struct OracleOut { int rc = 0, n = 0, npts = 0; int counters[B200_NCOUNTERS]; std::vector<double> times, vals; } oout;
    int oracle_hook(sCKT *ckt)
    {
        oout = OracleOut();
        double *t = 0, *v = 0;
        oout.rc = b200_harvest_oracle(ckt, oargs.path, oargs.nparams, oargs.param_id, oargs.inst_name,
            oargs.value, oout.counters, &oout.n, &oout.npts, &t, &v);
        if (!oout.rc) { oout.times.assign(t, t + oout.npts); oout.vals.assign(v, v + (size_t)oout.npts*oout.n); }
        free(t); free(v);
        return (oout.rc);
    }
}

int wrsref_run_deck_harvest_oracle(const char *deck_text, const char *oraclepath, int nparams,
    const int *param_id, const char *const *inst_name, const double *value)
{
    oargs = { oraclepath, nparams, param_id, inst_name, value };
    after_task_hook = oracle_hook;
    int err = wrsref_run_deck(deck_text);
    after_task_hook = 0;
    return (err);
}
int wrsref_oracle_rc(void) { return (oout.rc); }
int wrsref_oracle_n(void) { return (oout.n); }
int wrsref_oracle_npoints(void) { return (oout.npts); }
void wrsref_oracle_get(int *counters, double *times, double *vals)
{
    if (counters) memcpy(counters, oout.counters, sizeof(oout.counters));
    if (times) memcpy(times, oout.times.data(), oout.times.size()*sizeof(double));
    if (vals) memcpy(vals, oout.vals.data(), oout.vals.size()*sizeof(double));
}

// the set-up values of the first TJM instance of the last harvested circuit (fixture generation:
// tools/make_tjm_fixture.py).  out: 13 model doubles (rtype, ictype, nterms, criti, cap, cmu, gmu,
// r0, rsint, vshunt, tsaccl, tsfact, vgnom), 6 instance doubles (tcf, vdpbak, omegaJ, kgap,
// kgap_rejpt, alphaN), then 6*nterms coefficient doubles.  Returns the count or -1.
namespace {
    std::vector<double> tjm_dump;
    int tjm_dump_hook(sCKT *ckt)
    {
        tjm_dump.clear();
        b200_harvest *H = 0;
        int rc = b200_harvest_ckt(ckt, &H);
        if (!rc) {
            const b200_deck *d = b200_harvest_deck(H);
            if (d->ntjm > 0) {
                const b200_tjm &x = d->tjm[0];
                const b200_tjm_model &m = d->tjm_models[x.model];
                double mv[13] = { (double)m.rtype, (double)m.ictype, (double)m.nterms, m.criti, m.cap, m.cmu, m.gmu,
                    m.r0, m.rsint, m.vshunt, m.tsaccl, m.tsfact, m.vgnom };
                tjm_dump.assign(mv, mv + 13);
                double iv[6] = { x.tcf, x.vdpbak, x.omegaJ, x.kgap, x.kgap_rejpt, x.alphaN };
                tjm_dump.insert(tjm_dump.end(), iv, iv + 6);
                tjm_dump.insert(tjm_dump.end(), x.coef, x.coef + 6*m.nterms);
            }
        }
        b200_harvest_free(H);
        return (rc);
    }
}
int wrsref_tjm_setup_dump(const char *deck_text, double *out, int nout)
{
    after_task_hook = tjm_dump_hook;
    wrsref_run_deck(deck_text);
    after_task_hook = 0;
    if (tjm_dump.empty()) return (-1);
    for (int i = 0; i < nout && i < (int)tjm_dump.size(); i++) out[i] = tjm_dump[i];
    return ((int)tjm_dump.size());
}

int wrsref_harvest_rc(void) { return (hout.rc); }
int wrsref_harvest_n(void) { return (hout.n); }
int wrsref_harvest_npoints(void) { return (hout.npoints); }
void wrsref_harvest_get(int *counters, double *times, double *waves)
{
    if (counters) memcpy(counters, hout.counters.data(), hout.counters.size()*sizeof(int));
    if (times) memcpy(times, hout.times.data(), hout.times.size()*sizeof(double));
    if (waves) memcpy(waves, hout.waves.data(), hout.waves.size()*sizeof(double));
}

int wrsref_num_points(void) { return ((int)rec.ref.size()); }
int wrsref_num_names(void) { return (rec.numnames); }
const char *wrsref_names(int *nbytes) { *nbytes = (int)last_names.size(); return (last_names.data()); }

const char *wrsref_load_order(int *nbytes) { *nbytes = (int)last_order.size(); return (last_order.data()); }

// Copy recorded output: ref[npts], vals[npts*numnames] (row-major by point).
void wrsref_get_points(double *ref, double *vals)
{
    if (ref)
        memcpy(ref, rec.ref.data(), rec.ref.size()*sizeof(double));
    if (vals)
        memcpy(vals, rec.vals.data(), rec.vals.size()*sizeof(double));
}

// stats[0..7]: error, size, accepted, rejected, numIter, timePts, tranIter, iterCut, trapCut
void wrsref_get_counters(int *c)
{
    c[0] = last.error; c[1] = last.size; c[2] = last.accepted; c[3] = last.rejected;
    c[4] = last.numiter; c[5] = last.timepts; c[6] = last.traniter; c[7] = last.itercut;
    c[8] = last.trapcut;
}

// times[0..6]: load, decomp, solve, reorder, cvchk, tran, total   (sSTATS, circuit.h:1093-1163)
void wrsref_get_times(double *t)
{
    t[0] = last.loadtime; t[1] = last.decomptime; t[2] = last.solvetime;
    t[3] = last.reordertime; t[4] = last.cvchktime; t[5] = last.trantime; t[6] = last.tottime;
}

} // extern "C"
