// TEST INFRASTRUCTURE ONLY (oracle/_ref): C interface of the compiled-reference oracle.
#ifndef WRSREF_HARNESS_H
#define WRSREF_HARNESS_H
#ifdef __cplusplus
extern "C" {
#endif

extern int wrsref_verbose;
void wrsref_begin_plot(int numnames);
void wrsref_append(double ref, const double *vals, int n);

#ifdef __cplusplus
}
#endif
#endif
