/* Minimal stand-in for <lapacke.h> (not installed in this image).
 * The reference's hot path never calls LAPACK: LAPACK_zgetrf is referenced only by
 * determinant() (/root/reference/hmc.c:314-336), which nothing calls.  Test infrastructure only. */
#ifndef TB_ORACLE_LAPACKE_SHIM_H
#define TB_ORACLE_LAPACKE_SHIM_H
#include <complex.h>
void LAPACK_zgetrf(int *, int *, double _Complex *, int *, int *, int *);
void LAPACK_dgetrf(int *, int *, double *, int *, int *, int *);
void LAPACK_dgetri(int *, double *, int *, int *, double *, int *, int *);
#endif
