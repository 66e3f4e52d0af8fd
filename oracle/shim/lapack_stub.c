/* Aborting stubs: satisfy the link of the reference translation unit; never reached on the hot path. */
#include <stdio.h>
#include <stdlib.h>
#include <complex.h>
void LAPACK_zgetrf(int *a, int *b, double _Complex *c, int *d, int *e, int *f)
{ (void)a;(void)b;(void)c;(void)d;(void)e;(void)f; fprintf(stderr, "LAPACK_zgetrf stub called\n"); abort(); }
void LAPACK_dgetrf(int *a, int *b, double *c, int *d, int *e, int *f)
{ (void)a;(void)b;(void)c;(void)d;(void)e;(void)f; fprintf(stderr, "LAPACK_dgetrf stub called\n"); abort(); }
void LAPACK_dgetri(int *a, double *b, int *c, int *d, double *e, int *f, int *g)
{ (void)a;(void)b;(void)c;(void)d;(void)e;(void)f;(void)g; fprintf(stderr, "LAPACK_dgetri stub called\n"); abort(); }
