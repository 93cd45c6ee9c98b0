/* CPU restatement of FISHPACK90 HSTCRT for the oracle build (test infrastructure only).
 *
 * The reference calls the Fortran routine hstcrt_ (declared FiniteVolumeSolver.hpp:44, called
 * FiniteVolumeSolver.cpp:267) with MBDCND = NBDCND = 1 only.  There is no Fortran compiler in
 * the build image, so this file restates the routine in C with the same symbol and signature:
 *   - set-up of the linear system follows extern/fishpack90/src/hstcrt.f:386-446 line by line
 *     (DELXSQ = 2/dx^2, TWDYSQ = 2/dy^2, boundary rows, scaling of F by dy^2, PERTRB = 0,
 *     IERROR = 6 when ELMBDA > 0);
 *   - the solve replaces POISTGG (poistg.f:267-385, cyclic reduction) by another direct method
 *     for the *same* system  A(I)X(I-1,J)+B(I)X(I,J)+C(I)X(I+1,J)+X(I,J-1)-2X(I,J)+X(I,J+1)=Y(I,J)
 *     with X(I,0) = -X(I,1), X(I,N+1) = -X(I,N) (poistg.f:48-68, NPEROD = 1): diagonalise the
 *     J-operator (eigenvectors sin((j-1/2) k pi / N), eigenvalues 2cos(k pi/N)-2) and solve one
 *     tridiagonal system in I per mode.  Both are backward-stable direct solvers; results
 *     agree with a dense LU of the same system to ~1e-15 relative (tests/test_oracle.py).
 * PARITY NOTE: this is a restatement, not the Fortran; the FivePointStencil branch of the
 * reference (fully reference-sourced through the PETSc shim) solves the identical system for
 * alpha = beta = 1 and is used to cross-check it.
 */
#include <math.h>
#include <stdlib.h>
#include <stdio.h>

void hstcrt_(double* A, double* B, int* M_, int* MBDCND, double* BDA, double* BDB,
             double* C, double* D, int* N_, int* NBDCND, double* BDC, double* BDD,
             double* ELMBDA, double* F, int* IDIMF_, double* PERTRB, int* IERROR)
{
    const int M = *M_, N = *N_, IDIMF = *IDIMF_;
    *IERROR = 0;
    if (*A >= *B) *IERROR = 1;
    if (*MBDCND < 0 || *MBDCND > 4) *IERROR = 2;
    if (*C >= *D) *IERROR = 3;
    if (N <= 2) *IERROR = 4;
    if (*NBDCND < 0 || *NBDCND > 4) *IERROR = 5;
    if (IDIMF < M) *IERROR = 7;
    if (M <= 2) *IERROR = 8;
    if (*IERROR != 0) return;
    if (*MBDCND != 1 || *NBDCND != 1) {
        fprintf(stderr, "[hstcrt restatement] only MBDCND = NBDCND = 1 is restated\n");
        abort();
    }
#define Fij(i, j) F[(size_t)(j) * IDIMF + (i)]   /* column-major F(IDIMF,*), 0-based */
    const double deltax = (*B - *A) / (double)M;
    const double delxsq = 2.0 / (deltax * deltax);
    const double deltay = (*D - *C) / (double)N;
    const double delysq = deltay * deltay;
    const double twdysq = 2.0 / delysq;
    const double s = (deltay / deltax) * (deltay / deltax);
    const double st2 = 2.0 * s;
    double* wa = (double*)malloc(sizeof(double) * (size_t)M * 5);
    double* wb = wa + M; double* wc = wb + M; double* cp = wc + M; double* xh = cp + M;
    for (int i = 0; i < M; i++) { wa[i] = s; wb[i] = -st2 + (*ELMBDA) * delysq; wc[i] = s; }
    /* x-boundaries, MBDCND = 1 (hstcrt.f:411-424) */
    for (int j = 0; j < N; j++) Fij(0, j) -= BDA[j] * delxsq;
    wb[0] -= wa[0];
    for (int j = 0; j < N; j++) Fij(M - 1, j) -= BDB[j] * delxsq;
    wb[M - 1] -= wa[0];
    /* y-boundaries, NBDCND = 1 (hstcrt.f:430-439) */
    for (int i = 0; i < M; i++) Fij(i, 0) -= BDC[i] * twdysq;
    for (int i = 0; i < M; i++) Fij(i, N - 1) -= BDD[i] * twdysq;
    for (int j = 0; j < N; j++) for (int i = 0; i < M; i++) Fij(i, j) *= delysq;
    wa[0] = 0.0; wc[M - 1] = 0.0;                  /* MPEROD = 1 (hstcrt.f:445-448) */
    *PERTRB = 0.0;
    if (*ELMBDA > 0.0) *IERROR = 6;                /* hstcrt.f:450-452; still solves */

    /* direct solve: sine transform in J, Thomas in I */
    double* V = (double*)malloc(sizeof(double) * (size_t)N * N);   /* V[k*N + j] */
    double* Y = (double*)malloc(sizeof(double) * (size_t)M * N);   /* Y[k*M + i] */
    const double pi = 3.14159265358979323846264338327950288;
    for (int k = 1; k <= N; k++)
        for (int j = 0; j < N; j++) V[(size_t)(k - 1) * N + j] = sin((j + 0.5) * k * pi / N);
    for (int k = 0; k < N; k++)
        for (int i = 0; i < M; i++) {
            double acc = 0.0;
            for (int j = 0; j < N; j++) acc += Fij(i, j) * V[(size_t)k * N + j];
            Y[(size_t)k * M + i] = acc;
        }
    for (int k = 0; k < N; k++) {
        const double mu = 2.0 * cos((k + 1) * pi / N) - 2.0;
        double* y = Y + (size_t)k * M;
        double den = wb[0] + mu;
        cp[0] = wc[0] / den; xh[0] = y[0] / den;
        for (int i = 1; i < M; i++) {
            den = (wb[i] + mu) - wa[i] * cp[i - 1];
            cp[i] = wc[i] / den;
            xh[i] = (y[i] - wa[i] * xh[i - 1]) / den;
        }
        for (int i = M - 2; i >= 0; i--) xh[i] -= cp[i] * xh[i + 1];
        const double nrm = (k == N - 1) ? 1.0 / N : 2.0 / N;
        for (int i = 0; i < M; i++) y[i] = xh[i] * nrm;
    }
    for (int j = 0; j < N; j++)
        for (int i = 0; i < M; i++) {
            double acc = 0.0;
            for (int k = 0; k < N; k++) acc += Y[(size_t)k * M + i] * V[(size_t)k * N + j];
            Fij(i, j) = acc;
        }
#undef Fij
    free(V); free(Y); free(wa);
}
