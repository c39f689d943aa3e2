/* TEST INFRASTRUCTURE ONLY.  The reference's KLU adapter (wrspice/include/klumatrix.h:44-46,
 * src/sparse/kluif.cc) includes "klu.h" from SuiteSparse 4.4.6, which is not in the reference
 * tree (KLU/Makefile.in downloads it).  This stand-in forwards to the restated public
 * declarations in include/b200_klu_abi.h so the UNMODIFIED kluif.cc / spspice.cc compile with
 * -DUSE_KLU into oracle/_ref/libwrsref_klu.so. */
#ifndef KLU_H_STANDIN
#define KLU_H_STANDIN
#include "b200_klu_abi.h"
#endif
