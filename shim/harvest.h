/* shim/harvest.h -- reference-side binding of the batched engine (see harvest.cc). */
#ifndef B200_HARVEST_H
#define B200_HARVEST_H

#include "b200spice.h"

struct sCKT;
typedef struct b200_harvest b200_harvest;

#ifdef __cplusplus
extern "C" {
#endif

/* Walk a set-up sCKT (after sCKT::setup / resetup, ckt.cc:627-674) into a b200_deck.  The deck
 * and everything it points at belong to *out until b200_harvest_free. */
int  b200_harvest_ckt(struct sCKT *ckt, b200_harvest **out);
void b200_harvest_free(b200_harvest *h);
const b200_deck *b200_harvest_deck(const b200_harvest *h);
const char *b200_harvest_error(const b200_harvest *h);
/* table index of instance `name` within its kind (what b200_batch_set_param takes), or -1 */
int  b200_harvest_find(const b200_harvest *h, int kind, const char *name);

/* harvest + create + set_param rows + run_tran (or run_op) + fetch, through dlopen(libpath). */
int  b200_harvest_run(struct sCKT *ckt, const char *libpath, int ntrials, int device, int nparams,
                      const int *param_id, const char *const *inst_name, const double *const *values,
                      int *counters, int *n_out, int *npoints_out, double **times_out,
                      double **waves_out, int op_only);

/* test plumbing: the harvested deck through the plain-C oracle (oracle/libwrsoracle.so) */
int  b200_harvest_oracle(struct sCKT *ckt, const char *oraclepath, int nparams, const int *param_id,
                         const char *const *inst_name, const double *value, int *counters, int *n_out,
                         int *npts_out, double **times_out, double **vals_out);

/* host-only: b200_deck_analysis_stats of the harvested deck (sizes of pattern / LU / program) */
int  b200_harvest_stats(struct sCKT *ckt, const char *libpath, long long *out, int nout);
/* `.save` list of the next b200_harvest_run: unknown indices in the reference's output order (0: all) */
void b200_harvest_set_saves(int n, const int *unknown);
/* device time (CUDA events) and ticks of the latest b200_harvest_run */
void b200_harvest_last_stats(double *ms_total, long long *ticks);

#ifdef __cplusplus
}
#endif
#endif
