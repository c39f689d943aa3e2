// b200_internal.h -- shared declarations of the B200 batched Newton engine.
//
// Layout rule: everything that varies per trial lives in HBM as
// [slot][trial] with trial fastest (ldt = ntrials rounded up to 32), FP64.  Topology
// (node indices, gather lists, LU programs) is shared by all trials and read through L1/L2.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "b200spice.h"

// ---- mode bits: identical to wrspice/include/circuit.h:2158-2192 ----
#define MODETRAN        0x1
#define MODEDC          0x70
#define MODEDCOP        0x10
#define MODETRANOP      0x20
#define MODEDCTRANCURVE 0x40
#define INITF           0x7f00
#define MODEINITFLOAT   0x100
#define MODEINITJCT     0x200
#define MODEINITFIX     0x400
#define MODEINITSMSIG   0x800
#define MODEINITTRAN    0x1000
#define MODEINITPRED    0x2000
#define MODEUIC         0x10000

// ---- MOS per-instance state slots (mosdefs.h:330-375 minus the write-only "ask" mirrors) ----
enum { MS_vbd, MS_vbs, MS_vgs, MS_vds, MS_capgs, MS_qgs, MS_cqgs, MS_capgd, MS_qgd, MS_cqgd,
       MS_capgb, MS_qgb, MS_cqgb, MS_qbd, MS_cqbd, MS_qbs, MS_cqbs, MS_NSTATES };

// ---- MOS cached operating point (mosdefs.h:125,160-173) ----
enum { MC_mode, MC_von, MC_vdsat, MC_cd, MC_cbs, MC_cbd, MC_gm, MC_gds, MC_gmbs, MC_gbd, MC_gbs,
       MC_capbd, MC_capbs, MC_NCACHE };

// ---- MOS derived parameters produced by the temperature kernel (mostemp.cc:141-300) ----
enum { MP_tPhi, MP_tVbi, MP_tVto, MP_gamma, MP_lambda, MP_beta, MP_oxideCap,
       MP_tSourceSatCur, MP_tDrainSatCur, MP_sourceVcrit, MP_drainVcrit,
       MP_Cbd, MP_Cbdsw, MP_Cbs, MP_Cbssw, MP_tBulkPot, MP_tDepCap,
       MP_f2d, MP_f3d, MP_f4d, MP_f2s, MP_f3s, MP_f4s,
       MP_gsOvl, MP_gdOvl, MP_gbOvl, MP_dcon, MP_scon, MP_mult, MP_vt, MP_mj, MP_mjsw,
       MP_icVDS, MP_icVGS, MP_icVBS,
       // correctly rounded reciprocals of divisors that are constant for a device: a/b is then
       // formed by dev_div_rcp (bit-identical to the IEEE division, 5 instructions)
       MP_rvt, MP_rBulkPot, MP_r1mj, MP_r1mjsw, MP_r2phi, MP_r15phi,
       MP_NBASE,                 // everything above is all a level-1 deck needs
       // level 6 (moseq.cc:1091)
       MP_kv = MP_NBASE, MP_nv, MP_nc, MP_lamda0, MP_lamda1, MP_sigma, MP_gamma1,
       // level 3 (moseq.cc:693): model cards, doping-derived xd / alpha, and the geometry
       MP_eta, MP_theta, MP_kappa, MP_vmax, MP_xj, MP_ld, MP_nfs, MP_xd, MP_alpha,
       MP_oxideCapFactor, MP_tSurfMob, MP_xl, MP_w,
       // level 2 (moseq.cc:147)
       MP_uexp, MP_ucrit, MP_neff, MP_narrow, MP_nsub, MP_surfMob,
       MP_NPARAMS };

// ---- MOS raw (card) parameters that may vary per trial ----
enum { MR_vto, MR_kp, MR_gamma, MR_phi, MR_lambda, MR_l, MR_w, MR_NRAW };

// matrix stamps of one MOSFET in the reference's ldadd order (mosload.cc:887-923)
enum { MT_Dd, MT_Gg, MT_Ss, MT_Bb, MT_Ddp, MT_Gb, MT_Gdp, MT_Gsp, MT_Ssp, MT_Bg, MT_Bdp, MT_Bsp,
       MT_DPd, MT_SPs, MT_DPdp, MT_SPsp, MT_DPg, MT_DPb, MT_DPsp, MT_SPg, MT_SPb, MT_SPdp,
       MT_NSTAMP };
// rhs stamps (mosload.cc:862-875)
enum { MRH_g, MRH_b, MRH_dp, MRH_sp, MRH_NRHS };

// ---- shared-memory staging of the MOS load kernel (rows of MOS_TPB doubles) ----
#ifndef MOS_TPB
#define MOS_TPB 64                               // trials per CTA of k_load_mos
#endif
#ifndef MOS_MINB
#define MOS_MINB 8                               // __launch_bounds__ minimum: caps the kernel at 128 registers without
                                                 // spills: 7 CTAs per SM instead of 6, +1.6 % (shrinking the staging
                                                 // area by reading three rows late from HBM instead measured -1 %)
#endif
enum { SR_V_B, SR_V_SP, SR_V_G, SR_V_DP,                // node voltages (rhsOld); the predictor's
       SR_S2_vbs = SR_V_B, SR_S2_vgs = SR_V_SP, SR_S2_vds = SR_V_G,   // state-2 voltages share the rows
       SR_S0 = 4,                                       // state 0 block (old values in, new values out)
       SR_S1 = SR_S0 + MS_NSTATES,                      // state 1 block
       SR_CACHE = SR_S1 + MS_NSTATES,                   // cached operating point
       SR_Q2 = SR_CACHE + MC_NCACHE,                    // gate charges qgs,qgd,qgb two and three
       SR_Q3 = SR_Q2 + 3,                               //  steps back (LTE)
       SR_PRM = SR_Q3 + 3,                              // per-trial parameters
       SR_NFIXED = SR_PRM };
#define MOS_NSLOT (MT_NSTAMP + MRH_NRHS)    // contribution slots of one MOSFET (matrix + rhs)
struct MosStageMap {
    int off[MP_NPARAMS];    // per-trial: row*MOS_TPB; shared: offset of the value in the shared block
    int ts[MP_NPARAMS];     // 1 per-trial, 0 shared
    // accessor form: byte address of the value in the staging area = aoff + (8*thread & amask)
    int aoff[MP_NPARAMS];
    unsigned amask[MP_NPARAMS];
    int nvar;               // number of per-trial parameter rows
    int var[MP_NPARAMS];    // their parameter ids
};

// Strided parameter accessor: value(inst, trial) = base[inst*ld + trial*ts]
struct ParamRef {
    const double *base;
    int ld;     // 1 if shared across trials (then ts = 0), else ldt
    int ts;     // 0 or 1
};

// ---- per-device-kind topology tables (device memory mirrors are plain int arrays) ----
struct MosTopo {
    int n = 0;
    std::vector<int> d, g, s, b, dp, sp;        // node numbers (0 = ground)
    std::vector<int> type;                      // +1 / -1
    std::vector<int> level;
    std::vector<int> off;
    std::vector<int> state;                     // base state slot
    std::vector<int> cslot;                     // [n*MT_NSTAMP] contribution slot or -1
    std::vector<int> rslot;                     // [n*MRH_NRHS] rhs contribution slot or -1
};
struct CapTopo {
    int n = 0;
    std::vector<int> pos, neg, state;
    std::vector<int> cslot;                     // [n*4] PP NN PN NP
    std::vector<int> rslot;                     // [n*2] pos neg
};
struct IndTopo {
    int n = 0;
    std::vector<int> pos, neg, branch, state;
    std::vector<int> cslot;                     // [n] (br,br) contribution: -req
    std::vector<int> rslot;                     // [n] branch row: veq
    std::vector<int> icgiven; std::vector<double> ic;
    // couplings onto each inductor, in MUT load order (mutload.cc adds the cross flux to the
    // inductor's state before the inductor itself loads): CSR over inductors
    std::vector<int> mptr, mother;              // other inductor of the pair
    std::vector<double> mk;                     // coupling coefficient
};
struct MutTopo {
    int n = 0;
    std::vector<int> ind1, ind2;
    std::vector<double> k;
    std::vector<int> cslot;                     // [2n]: (br1,br2), (br2,br1)
};
// Josephson junctions, RSJ level 1 (jjdefs.h:237-250 states; no lser / lsh / control)
enum { JS_voltage, JS_dvdt, JS_phase, JS_conI, JS_phsInt, JS_crti, JS_qpi, JS_NSTATES };
// per-instance derived parameters (the instance part of JJdev::setup, jjset.cc:609-790)
enum { JP_criti, JP_cap, JP_g0, JP_gn, JP_gs, JP_g1, JP_g2, JP_cr1, JP_cr2, JP_gshunt, JP_NPARAMS };
// contribution slots of one junction in the reference's ldadd / rhsadd order (jjload.cc:826-871,
// then the shunt resistor :351-356)
enum { JT_PP, JT_PN, JT_NP, JT_NN, JT_RPP, JT_RPN, JT_RNP, JT_RNN, JT_NSTAMP };
enum { JR_phs, JR_pos, JR_neg, JR_NRHS };
#define JJ_MAXMODELS 8
#define JJ_PHI0      2.067833848e-15          /* wrsCONSTphi0, circuit.h:123 */
#define JJ_PHI0_2PI  (JJ_PHI0/(2*3.14159265358979323846))
struct JJTopo {
    int n = 0, nmodels = 0;
    std::vector<int> pos, neg, phs, model, state;
    std::vector<int> cslot;                     // [n*JT_NSTAMP]
    std::vector<int> rslot;                     // [n*JR_NRHS]
    std::vector<double> vless, vmore, vg, initV, initPhi;   // per instance, shared by all trials
    std::vector<double> area0, ics0, vshunt, tcf;           // card values (NaN = not given)
    std::vector<b200_jj_model> models;
};
// Tunnel-junction model (JJ level 3, devlib/tjm).  States tjmdefs.h:245-250.
enum { TS_voltage, TS_dvdt, TS_phase, TS_phsInt, TS_crti, TS_qpi, TS_NSTATES };
enum { TP_criti, TP_cap, TP_g0, TP_gshunt, TP_gcrit, TP_NPARAMS };
#define TJM_MAXTERMS 20
// per (junction, trial) work arrays, in units of nterms complex numbers: the recurrences of this
// iteration and of the last accepted point, and the step coefficients of tjm_newstep
enum { TF_Fc, TF_Fs, TF_Fcprev, TF_Fsprev, TF_alpha0, TF_alpha1, TF_expz, TF_NARR };
struct TJMTopo {
    int n = 0, nmodels = 0, nterms = 0;         // nterms: the largest of the models
    std::vector<int> pos, neg, phs, model, state;
    std::vector<int> cslot;                     // [n*JT_NSTAMP]
    std::vector<int> rslot;                     // [n*JR_NRHS]
    std::vector<double> initV, initPhi, area0, ics0, vshunt, tcf, omegaJ, kgap, kgap_rejpt, alphaN, vdpbak;
    std::vector<double> coef;                   // [n][3][nterms][2]: p, A, B
    std::vector<b200_tjm_model> models;
};
struct SrcTopo {
    int n = 0;
    std::vector<int> kind, pos, neg, branch, fn, dcgiven;
    std::vector<int> rslot;                     // [n*2]: V: branch,-1 ; I: pos,neg
    std::vector<int> coef_off, ncoef;           // into coef pool (run-time parms)
    std::vector<double> dc;
    std::vector<double> pool;                   // per source: parms[8] cache[2] then pwl pts/slopes
};

// ---- solve program (k_lu) ----
// One trial's LU refactor + forward + backward solve is ONE task graph over the trial's value
// image S[]: a task computes one value
//      S[dst] = ( S[dst] - sum_q S[a_q]*S[b_q] ) [ / S[div] ]
// with the products subtracted in list order.  Factor entries (dst = the LU slot of (i,j), ops
// over the pivots k < min(i,j) ascending, L entries divided by the pivot), rows of the forward
// solve (dst = y slot, ascending column) and rows of the backward solve (descending column,
// divided by the diagonal) are all tasks of this one form, level-scheduled together.
// A sum of more than LU_SPLIT products is split (the same rule is restated by the oracle's
// static LU so both sides round identically): chunks of LU_CHUNK consecutive products are
// accumulated from 0.0 into scratch slots, then added to the target in chunk order (as
// products with the constant -1.0 slot: x - p*(-1.0) == x + p exactly); applied recursively.
#define LU_SPLIT 32
#define LU_CHUNK 16
#define LU_LANES 32                              // task lanes per trial: one warp runs one trial
#define LU_NONE 0xffffu                          // "no divisor" in a header word
// The program is a stream of ROWS of LU_LANES 32-bit words, one word per lane, grouped in
// rounds.  A round = up to LU_LANES tasks of one level (longest first; lane e runs task e):
//   1 header row  : dst | div << 16   (LU_NONE: no division)
//   nsteps op rows: a | b << 16
// Lanes without a task, and lanes whose task has run out of products, execute the exact no-op
// S[zero] -= S[zero]*S[zero] on a slot that holds 0.0, so the interpreter has no predicates.
// prog_round[r] = nsteps | LU_RSYNC if the round opens a new level (its operands were written
// by earlier rounds: the warp synchronises before it reads them).
#define LU_RSYNC 0x8000u
// WIDE programs: an image that does not fit one warp's shared-memory window (or 16-bit slot
// numbers) stays in global memory / L2 and its program carries 32-bit slot numbers -- two
// words per lane and row: header {dst, div (LU_NONE_W: none)}, op {a, b}.  Same tasks, same
// order, same arithmetic; k_lu_wide (solve.cu) interprets it.
#define LU_NONE_W 0xffffffffu
#define LU_SMEM_BUDGET 229000                    // bytes of the 227 KB window left to image + round table
#define LU_MIN_RESIDENT 4                        // images that must fit the window for the narrow path
// Assembly run of one image slot: S[slot] = finalize( sum of contrib[pos .. pos+len) in order )
struct AsmRun { int pos, len, info, pad; };
#define ASM_TILE_SLOTS 16                        // slots per k_assemble warp tile
// A run of more than ASM_LONG terms (the supply node of a 100 000-transistor deck collects one
// term per transistor) is summed in two steps: chunks of ASM_LCHUNK consecutive terms are
// accumulated from 0.0 by k_assemble_long, one warp per chunk, into extra rows of the
// contribution array [NC, NC + nlong); the slot's run then names those rows, so k_assemble adds
// the chunk sums in order and finalizes the whole.  Decks without such runs never see it.
#define ASM_LONG 4096
#define ASM_LCHUNK 1024
struct AsmChunk { int pos, len; };
#define ASM_DIAG 0x40000000
#define ASM_ROW  0x20000000
#define ASM_VNODE 0x10000000  // diagonal of a voltage node (loadGmin applies)

// the per-order products of b200_build_lu that the solve kernels read
struct LuProgram {
    std::vector<int> rowperm, colperm, yperm_out, tile_order;
    std::vector<AsmRun> runs;
    std::vector<unsigned> prog_rows;
    std::vector<unsigned short> prog_round;
    int nnzLU = 0, img = 0, wide = 0;
    std::vector<int> round_row, level_ptr;
    int prog_levels = 0, prog_tasks = 0; long long prog_steps = 0, prog_ops = 0;
};

struct Analysis {
    // which trials the primary program serves: -1 all, 1 the DC modes only (then tranp serves
    // the transient modes)
    int primary_class = -1;
    bool has_tranp = false;
    LuProgram tranp;
    // unknowns
    int n = 0;                                  // matrix size; node numbers 1..n
    std::vector<int> ntype;                     // [n+1] 0 voltage, 1 current
    std::vector<std::string> names;             // [n+1]
    int numStates = 0;
    // CSC pattern of A (0-based, rows ascending per column) -- spbuild.cc:922-958
    int nnz = 0;
    std::vector<int> Ap, Ai;
    std::vector<int> diag;                      // [n+1] Ax index of (i,i) or -1
    std::vector<double> Ainit;                  // preloaded constants (shared by all trials)
    std::vector<char> dc_active;                // [nnz] entry is structurally nonzero in DC modes too
                                                //  (capacitor / inductor companions vanish in DC)
    // gather lists: contributions in reference load order
    int ncslot = 0, nrslot = 0;
    std::vector<int> cg_ptr, cg_idx;            // per Ax entry: cslots
    std::vector<int> rg_ptr, rg_idx;            // per rhs row (1..n): rslots
    // .ic / nodeset
    std::vector<char> ic_given, ns_given;
    std::vector<double> ic_val, ns_val;
    int hadNodeset = 0;
    // device tables
    MosTopo mos; CapTopo cap; SrcTopo src; IndTopo ind; MutTopo mut; JJTopo jj; TJMTopo tjm;
    int nres = 0;
    double initDelta = 0;                       // CKTinitDelta from JJdev::getic (jjic.cc:78-104), 0 = unset
    // resistors that touch a phase node are loaded every iteration instead of preloaded
    // (restemp.cc:213-248): their +-g are constant contributions at the resistor's place in the
    // load order (contribution slot, value)
    std::vector<int> rload_slot; std::vector<double> rload_val;
    std::vector<int> rph_res, rph_cls, rph_cslot; std::vector<double> rph_g;   // [n], [n], [4n], [n]
    int jjDCphase = 0;                          // CKTjjDCphase (ckt.cc:2137-2176)
    std::vector<char> phase_mask;               // [n+1] node carries the phase flag
    // ordering: position -> original (0-based)
    std::vector<int> rowperm, colperm;
    // LU structure in permuted positions
    int nnzLU = 0;
    std::vector<int> Lp, Li, Up, Ui;            // strictly lower / upper by column
    std::vector<int> slotL, slotU, slotD;       // LU value slot of each L entry / U entry / diag
    std::vector<int> a2lu;                      // Ax index -> LU slot
    std::vector<int> lu_src;                    // LU slot -> Ax index or -1 (fill)
    // image layout: [0,nnzLU) LU values, [nnzLU, nnzLU+n) rhs -> y -> x (permuted positions),
    // then nscratch partial sums, the constant -1.0 and a 0.0; img = padded total
    int nscratch = 0, slot_neg1 = 0, slot_zero = 0, img = 0;
    int NC = 0;                                 // total contributions (matrix + rhs + constants), assembly order
    std::vector<AsmRun> runs;                   // [img]
    std::vector<int> tile_order;                // tiles of ASM_TILE_SLOTS slots, most terms first
    std::vector<int> const_idx; std::vector<double> const_val;   // Ainit pseudo-terms, the -1.0
    std::vector<AsmChunk> long_chunks;          // chunk c -> contribution row NC + c
    // per-trial resistor values (B200_P_RES_R): the preloaded entries that hold resistor
    // conductances, with their terms in accumulation (= load) order, so the per-trial constant
    // can be re-summed exactly as the nominal one was (restemp.cc:209-229)
    double resmin = 0.0;
    std::vector<double> res_r, res_m;           // card value and multiplier of every resistor
    std::vector<int> rt_entry;                  // [ne] CSC entry p
    std::vector<double> rt_base;                // [ne] the entry's value before the resistor terms
    std::vector<int> rt_ptr, rt_res;            // CSR: terms of entry e
    std::vector<signed char> rt_sign;
    std::vector<int> ainit_cslot;               // [nnz] contribution row of the entry's preloaded constant, or -1
    // task program
    std::vector<unsigned> prog_rows;            // [nrows][LU_LANES] ([nrows][LU_LANES][2] if wide)
    int wide = 0;                               // 1: 32-bit slot numbers, image in global memory (k_lu_wide)
    std::vector<int> round_row;                 // wide: first program row of every round
    std::vector<int> level_ptr;                 // wide: [nlevels+1] first round of every level
    std::vector<unsigned short> prog_round;     // per round: nsteps | LU_RSYNC
    int prog_levels = 0, prog_tasks = 0; long long prog_steps = 0, prog_ops = 0;
    std::vector<int> yperm_in;                  // y[k] = rhs[rowperm[k]+1]
    std::vector<int> yperm_out;                 // x[colperm[k]+1] = y[k]
    // breakpoints (shared by all trials: sources are not varied per trial)
    std::vector<double> brks;
    std::vector<double> lat_offs, lat_per;
    // tran constants (sTRANAN::init, tranan.cc:94-217)
    double t_stop = 0, t_step = 0, t_start = 0, maxStep = 0, delmin = 0, minBreak = 0;
    int nocut = 0, npoints_max = 0;
    double break0 = 0, break1 = 0;
};

// Orders + builds the solve program(s) of a deck: one program when the run never solves in a DC
// mode (`uic`) or when the DC and the transient pivot orders coincide; otherwise the primary
// program (an.*) serves the DC modes and an.tranp the transient modes.  Both read the SAME
// contribution layout (the primary's assembly order).
int  b200_build_programs(Analysis &an, bool uic);
int  b200_analyze(const b200_deck *deck, Analysis &an, std::string &err);
void b200_order_markowitz(Analysis &an, bool tran_order = false);
// numerical pivot test / threshold re-ordering on one value set (analysis.cpp)
int  b200_order_numeric(Analysis &an, const double *Ax, double pivrel, double pivabs, int mode);
// symbolic LU for the fixed order, the task program and the assembly runs; returns 0 or
// B200_E_UNSUPP when the image does not fit the 16-bit slot indices of the program
int  b200_build_lu(Analysis &an);
