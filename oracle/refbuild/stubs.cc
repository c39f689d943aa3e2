// TEST INFRASTRUCTURE ONLY (oracle/_ref).  Headless host layer for the UNMODIFIED reference
// numerical core (wrcad/xictools WRspice: src/ckt, src/analysis/tranan.cc, xt_base/sparse,
// devlib/*), which is compiled verbatim from /root/reference by oracle/refbuild/Makefile.
//
// The reference's front end (shell, plots, GUI: wrspice/src/{cp,fte,plot,...}, bin/wrspice.cc)
// is not built.  This file supplies the handful of host singletons the core links against
// (Sp, OP, TTY, GP, CP, DVO, Global, DEV table, analysis table) with minimal behaviour:
// no shell variables, no plots, output points recorded into a C array (harness.cc).
// Nothing here implements circuit arithmetic.

#include "config.h"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <fenv.h>

#include "simulator.h"
#include "graph.h"
#include "output.h"
#include "cshell.h"
#include "ttyio.h"
#include "device.h"
#include "input.h"
#include "circuit.h"
#include "trandefs.h"
#include "dcodefs.h"
#include "dctdefs.h"
#include "optdefs.h"
#include "subexpand.h"
#include "verilog.h"
#include "variable.h"
#include "noisdefs.h"
#include "spglobal.h"
#include "parser.h"
#include "commands.h"
#include "runop.h"
#include "datavec.h"
#include "ginterf/graphics.h"
#include "miscutil/pathlist.h"
#include "miscutil/random.h"
#include "spnumber/spnumber.h"

#include "harness.h"

static void die(const char *what)
{
    fprintf(stderr, "wrsref stub: unsupported host call: %s\n", what);
    abort();
}

// ---------------------------------------------------------------------------
// Device table (what devlib/Makefile.in:128-154 generates as toc.cc).
// ---------------------------------------------------------------------------
extern "C" {
    extern void mos_c(IFdevice**, int*);
    extern void cap_c(IFdevice**, int*);
    extern void res_c(IFdevice**, int*);
    extern void src_c(IFdevice**, int*);
    extern void ind_c(IFdevice**, int*);
    extern void jj_c(IFdevice**, int*);
    extern void tjm_c(IFdevice**, int*);
}

const char *WRS_ModuleName;
const char *WRS_ModuleVersion;
const char *sDevLib::dl_version = DEVLIB_VERSION_STR;
sDevLib DEV;

void
sDevLib::init()
{
    dl_devices = 0;
    dl_numdevs = 0;
    dl_size = 0;
    // Same relative order as devlib/Makefile.in DEVICES (res, cap, ind, ... src ..., mos, jj, tjm).
    loadDev(&res_c);
    loadDev(&cap_c);
    loadDev(&ind_c);
    loadDev(&src_c);
    loadDev(&mos_c);
    loadDev(&jj_c);
    loadDev(&tjm_c);
}

// ---------------------------------------------------------------------------
// Analysis table (bin/wrspice.cc:2208-2241).
// ---------------------------------------------------------------------------
IFanalysis *IFanalysis::analyses[] = { &DCOinfo, &TRANinfo, 0 };
int IFanalysis::num_analyses = 3;

IFparm sCKTnode::params[] = {
    IP("nodeset", PARM_NS, IF_REAL, "suggested initial voltage"),
    IP("ic",      PARM_IC, IF_REAL, "initial voltage"),
    IP("type",    PARM_NODETYPE, IF_INTEGER, "output type of equation")
};
int sCKTnode::num_params = 3;

const char *dckw_name1 = "name1", *dckw_start1 = "start1", *dckw_stop1 = "stop1",
    *dckw_step1 = "step1", *dckw_name2 = "name2", *dckw_start2 = "start2",
    *dckw_stop2 = "stop2", *dckw_step2 = "step2", *kw_dc = "dc", *kw_sweep = "sweep";

// ---------------------------------------------------------------------------
// Host singletons.
// ---------------------------------------------------------------------------
sGlobal Global;
sRnd Rnd;
sSPcx SPcx;
SPgraphics GP;
sDevOut DVO;
SPinput IP;

sConstant IFsimulator::ft_constants[] = { sConstant(0, 0.0, 0) };

IFsimulator::IFsimulator()
{
    memset((void*)this, 0, sizeof(*this));
    ft_fpe_mode = FPEdefault;       // simulator.h:588-595: default is FPEcheck
    DEV.init();
    ft_subc_catchar = DEF_SUBC_CATCHAR;
    ft_units_catchar = DEF_UNITS_CATCHAR;
    ft_units_sepchar = DEF_UNITS_SEPCHAR;
    ft_plot_catchar = DEF_PLOT_CATCHAR;
    ft_spec_catchar = DEF_SPEC_CATCHAR;
    ft_subc_catmode = SUBC_CATMODE_WR;
}
IFsimulator Sp;

// No shell variables are ever set in the headless oracle.
bool IFsimulator::GetVar(const char*, VTYPenum, VTvalue*, sFtCirc*) { return (false); }
bool IFsimulator::HspiceFriendly() { return (false); }
void IFsimulator::SetFPEmode(FPEmode m) { ft_fpe_mode = m; }
FPEmode IFsimulator::SetCircuitFPEmode() { return (ft_fpe_mode); }
const char *IFsimulator::ErrorShort(int code)
{
    static char buf[32];
    snprintf(buf, sizeof(buf), "E%d", code);
    return (buf);
}
void IFsimulator::Error(int code, const char *mesg, const char *badone)
{
    if (wrsref_verbose)
        fprintf(stderr, "wrsref: Error %d %s %s\n", code, mesg ? mesg : "", badone ? badone : "");
}
sDataVec *IFsimulator::Evaluate(pnode*, bool) { die("Sp.Evaluate"); return (0); }
void IFsimulator::ExecCmds(wordlist*) { die("Sp.ExecCmds"); }
pnode *IFsimulator::GetPnode(const char**, bool, bool, bool) { die("Sp.GetPnode"); return (0); }
bool IFsimulator::UserFunc(const char*, int, char**, const char**) { return (false); }
pnode::~pnode() { }

int SPgraphics::Checkup() { return (0); }

sTTYio TTY;
sTTYio::sTTYio() { memset((void*)this, 0, sizeof(*this)); }
void sTTYio::printf(const char *fmt, ...)
{
    if (!wrsref_verbose) return;
    va_list args; va_start(args, fmt); vfprintf(stdout, fmt, args); va_end(args);
}
void sTTYio::err_printf(const char *fmt, ...)
{
    if (!wrsref_verbose) return;
    va_list args; va_start(args, fmt); vfprintf(stderr, fmt, args); va_end(args);
}

CshPar CP;
CshPar::CshPar() { memset((void*)this, 0, sizeof(*this)); }
wordlist *CshPar::BracExpand(const char*) { die("CP.BracExpand"); return (0); }
bool CshPar::GlobMatch(const char *p, const char *s) { return (p && s && !strcmp(p, s)); }

void sDevOut::dumpStrobe() { }
void sDevOut::cleanup() { }
void sDevOut::textOut(OUTerrType, const char *fmt, ...)
{
    if (!wrsref_verbose) return;
    va_list args; va_start(args, fmt); vfprintf(stderr, fmt, args); va_end(args);
    fputc('\n', stderr);
}

namespace ginterf {
    GRpkg *GRpkg::instancePtr = 0;
    void GRpkg::on_null_ptr() { die("GRpkg::self"); }
    void GRpkg::ErrPrintf(Etype, const char *fmt, ...)
    {
        if (!wrsref_verbose) return;
        va_list args; va_start(args, fmt); vfprintf(stderr, fmt, args); va_end(args);
    }
}

bool sSPcx::kwMatchModel(const char *s) { return (lstring::cimatch(".model", s)); }
bool sSPcx::kwMatchSubstart(const char *s)
    { return (lstring::cimatch(".subckt", s) || lstring::cimatch(".macro", s)); }

VerilogBlock::~VerilogBlock() { }
void VerilogBlock::initialize(sOUTdata*) { die("VerilogBlock"); }
void VerilogBlock::finalize(bool) { die("VerilogBlock"); }
void VerilogBlock::run_step(sOUTdata*) { die("VerilogBlock"); }
bool VerilogBlock::query_var(const char*, const char*, double*) { return (false); }

int sFtCirc::newCKT(sCKT**, sTASK**) { die("sFtCirc::newCKT"); return (0); }
// Only the UID symbol table is used (ckt/symtab.cc:63-68 goes through CKTbackPtr->symtab()).
sFtCirc::sFtCirc()
{
    ci_next = 0; ci_name = 0; ci_descr = 0; ci_filename = 0;
    ci_deck = 0; ci_origdeck = 0; ci_verilog = 0; ci_options = 0; ci_params = 0;
    ci_vars = 0; ci_defines = 0; ci_commands = 0; ci_nodes = 0; ci_devices = 0;
    ci_models = 0; ci_defOpt = 0; ci_deferred = 0; ci_trial_deferred = 0;
    ci_symtab = new sSymTab(true);      // fte/circuit.cc:528: case-insensitive
    ci_runckt = 0; ci_sweep = 0; ci_check = 0; ci_runplot = 0; ci_runtype = 0;
    ci_inprogress = false; ci_runonce = false; ci_keep_deferred = false;
    ci_use_trial_deferred = false;
}
sFtCirc::~sFtCirc() { delete ci_symtab; }
void sRunopDb::clear() { }
void CommandTab::com_dumpnodes(wordlist*) { }

FILE *pathlist::open_path_file(const char*, const char*, const char*, char**, bool)
    { return (0); }

variable::variable(const char *n) { memset((void*)this, 0, sizeof(*this)); }
void variable::set_boolean(bool) { }

void NevalSrc(double*, double*, sCKT*, int, int, int, double, double) { die("noise"); }
double sNdata::integrate(double, double, double) { die("noise"); return (0); }
int IFanalysis::parseDC(sLine*, sCKT*, const char**, sJOB*, int) { die("parseDC"); return (0); }

bool sUnits::set(const char*) { return (true); }
char *sUnits::unitstr() const { return (lstring::copy("")); }

// ---------------------------------------------------------------------------
// fte/interface.cc look-ups needed by the deck parser (keyword -> IFparm).
// ---------------------------------------------------------------------------
IFparm *IFdevice::findInstanceParm(const char *name, int bits)
{
    for (int i = 0; i < dv_numInstanceParms; i++) {
        IFparm *p = dv_instanceParms + i;
        if ((p->dataType & bits) && lstring::cieq(name, p->keyword))
            return (p);
    }
    return (0);
}
IFparm *IFdevice::findModelParm(const char *name, int bits)
{
    for (int i = 0; i < dv_numModelParms; i++) {
        IFparm *p = dv_modelParms + i;
        if ((p->dataType & bits) && lstring::cieq(name, p->keyword))
            return (p);
    }
    return (0);
}
void IFdata::cleanup() { }
variable *IFparm::p_to_v(const IFvalue*, int) { die("p_to_v"); return (0); }
int IFspecial::evaluate(const char*, sCKT*, IFdata*, int) { die("IFspecial"); return (0); }
variable *IFanalysis::getOpt(const sCKT*, const sJOB*, const char*, IFspecial*) const
    { die("getOpt"); return (0); }
int IFanalysis::getOpt(const sCKT*, const sJOB*, const char*, IFdata*, IFspecial*) const
    { die("getOpt"); return (0); }

// ---------------------------------------------------------------------------
// Output: record every point the reference dumps (ckt.cc:680-708 -> OP.appendData).
// ---------------------------------------------------------------------------
IFoutput OP;
IFoutput::IFoutput() { memset((void*)this, 0, sizeof(*this)); }

double IFoutput::seconds()
{
    struct timeval tv;
    gettimeofday(&tv, 0);
    return (tv.tv_sec + 1e-6*tv.tv_usec);
}
sRunDesc *IFoutput::beginPlot(sOUTdata *outd, int, const char*, double)
{
    wrsref_begin_plot(outd->numNames);
    return ((sRunDesc*)1);
}
int IFoutput::appendData(sRunDesc*, IFvalue *ref, IFvalue *val)
{
    wrsref_append(ref->rValue, val->v.vec.rVec, val->v.numValue);
    return (OK);
}
int IFoutput::insertData(sCKT*, sRunDesc*, IFvalue*, IFvalue*, unsigned int)
    { die("insertData"); return (0); }
int IFoutput::setDims(sRunDesc*, int*, int, bool) { return (OK); }
int IFoutput::setDC(sRunDesc*, sDCTprms*) { return (OK); }
void IFoutput::unrollPlot(sRunDesc*) { }
void IFoutput::endPlot(sRunDesc*, bool) { }
void IFoutput::initRunops(sRunDesc*) { }
void IFoutput::checkRunops(sRunDesc*, double) { }
int IFoutput::pauseTest(sRunDesc*) { return (0); }
sDataVec *IFoutput::vecGet(const char*, const sCKT*, bool) { return (0); }
int IFoutput::error(ERRtype, const char *fmt, ...)
{
    if (wrsref_verbose) {
        va_list args; va_start(args, fmt); vfprintf(stderr, fmt, args); va_end(args);
        fputc('\n', stderr);
    }
    return (0);
}
