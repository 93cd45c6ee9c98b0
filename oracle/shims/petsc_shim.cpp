// Implementation of the sequential PETSc shim declared in petsc.h (oracle build only).
// MatLUFactor / MatSolve forward to LAPACK dgetrf / dgetrs exactly as PETSc's MATSEQDENSE does.
#include <petsc.h>
#include <cstdio>
#include <cstdlib>
extern "C" {
void dgetrf_(int* m, int* n, double* a, int* lda, int* ipiv, int* info);
void dgetrs_(char* trans, int* n, int* nrhs, double* a, int* lda, int* ipiv, double* b, int* ldb, int* info);
}
static void unsupported(const char* what) { std::fprintf(stderr, "[petsc shim] %s is not supported in the oracle build\n", what); std::abort(); }

PetscErrorCode PetscInitialize(int*, char***, const char*, const char*) { return 0; }
PetscErrorCode PetscFinalize() { return 0; }
PetscErrorCode PetscGetArgs(int* argc, char*** argv) { static int c = 0; static char** v = nullptr; *argc = c; *argv = v; return 0; }

PetscErrorCode MatCreate(MPI_Comm, Mat* A) { *A = new _ef_shim_Mat; return 0; }
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N) { A->m = (M >= 0 ? M : m); A->n = (N >= 0 ? N : n); return 0; }
PetscErrorCode MatSetType(Mat, MatType) { return 0; }
PetscErrorCode MatSetFromOptions(Mat) { return 0; }
PetscErrorCode MatSetUp(Mat A) { A->a.assign((std::size_t)A->m * A->n, 0.0); A->factored = false; return 0; }
PetscErrorCode MatSetValue(Mat A, PetscInt i, PetscInt j, PetscScalar v, InsertMode mode) { return MatSetValues(A, 1, &i, 1, &j, &v, mode); }
PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt idxm[], PetscInt n, const PetscInt idxn[], const PetscScalar v[], InsertMode mode) {
    if (A->a.empty()) MatSetUp(A);
    for (int r = 0; r < m; r++) for (int c = 0; c < n; c++) {
        if (idxm[r] < 0 || idxn[c] < 0) continue; // PETSc ignores negative indices
        double& dst = A->a[(std::size_t)idxn[c] * A->m + idxm[r]];
        if (mode == ADD_VALUES) dst += v[r * n + c]; else dst = v[r * n + c];
    }
    return 0;
}
PetscErrorCode MatGetValue(Mat A, PetscInt i, PetscInt j, PetscScalar* v) { *v = A->a[(std::size_t)j * A->m + i]; return 0; }
PetscErrorCode MatGetValues(Mat A, PetscInt m, const PetscInt idxm[], PetscInt n, const PetscInt idxn[], PetscScalar v[]) {
    for (int r = 0; r < m; r++) for (int c = 0; c < n; c++) v[r * n + c] = A->a[(std::size_t)idxn[c] * A->m + idxm[r]];
    return 0;
}
PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType) { return 0; }
PetscErrorCode MatAssemblyEnd(Mat, MatAssemblyType) { return 0; }
PetscErrorCode MatGetSize(Mat A, PetscInt* M, PetscInt* N) { if (M) *M = A->m; if (N) *N = A->n; return 0; }
PetscErrorCode MatGetLocalSize(Mat A, PetscInt* m, PetscInt* n) { return MatGetSize(A, m, n); }
PetscErrorCode MatGetOwnershipRange(Mat A, PetscInt* first, PetscInt* last) { *first = 0; *last = A->m; return 0; }
PetscErrorCode MatFactorInfoInitialize(MatFactorInfo* info) { info->fill = 0; return 0; }
PetscErrorCode MatLUFactor(Mat A, IS, IS, const MatFactorInfo*) {
    int info = 0; A->ipiv.assign(A->m, 0);
    dgetrf_(&A->m, &A->n, A->a.data(), &A->m, A->ipiv.data(), &info);
    if (info != 0) std::fprintf(stderr, "[petsc shim] dgetrf info = %d\n", info);
    A->factored = true; return 0;
}
PetscErrorCode MatSolve(Mat A, Vec b, Vec x) {
    if (!A->factored) unsupported("MatSolve on an unfactored matrix");
    x->n = b->n; x->v = b->v; int one = 1, info = 0; char trans = 'N';
    dgetrs_(&trans, &A->m, &one, A->a.data(), &A->m, A->ipiv.data(), x->v.data(), &A->m, &info);
    return info;
}
PetscErrorCode MatDestroy(Mat* A) { delete *A; *A = nullptr; return 0; }
PetscErrorCode MatCreateConstantDiagonal(MPI_Comm, PetscInt, PetscInt, PetscInt, PetscInt, PetscScalar, Mat*) { unsupported("MatCreateConstantDiagonal"); return 1; }
PetscErrorCode MatCreateSubMatrix(Mat, IS, IS, MatReuse, Mat*) { unsupported("MatCreateSubMatrix"); return 1; }
PetscErrorCode MatCreateSubMatrices(Mat, PetscInt, const IS[], const IS[], MatReuse, Mat*[]) { unsupported("MatCreateSubMatrices"); return 1; }
PetscErrorCode MatCreateMPIMatConcatenateSeqMat(MPI_Comm, Mat, PetscInt, MatReuse, Mat*) { unsupported("MatCreateMPIMatConcatenateSeqMat"); return 1; }

PetscErrorCode VecCreate(MPI_Comm, Vec* v) { *v = new _ef_shim_Vec; return 0; }
PetscErrorCode VecSetSizes(Vec v, PetscInt n, PetscInt N) { v->n = (N >= 0 ? N : n); v->v.assign(v->n, 0.0); return 0; }
PetscErrorCode VecSetType(Vec, VecType) { return 0; }
PetscErrorCode VecSetFromOptions(Vec) { return 0; }
PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar y, InsertMode mode) { if (mode == ADD_VALUES) v->v[i] += y; else v->v[i] = y; return 0; }
PetscErrorCode VecSetValues(Vec v, PetscInt ni, const PetscInt ix[], const PetscScalar y[], InsertMode mode) { for (int k = 0; k < ni; k++) VecSetValue(v, ix[k], y[k], mode); return 0; }
PetscErrorCode VecAssemblyBegin(Vec) { return 0; }
PetscErrorCode VecAssemblyEnd(Vec) { return 0; }
PetscErrorCode VecDuplicate(Vec v, Vec* newv) { *newv = new _ef_shim_Vec; (*newv)->n = v->n; (*newv)->v.assign(v->n, 0.0); return 0; }
PetscErrorCode VecGetArray(Vec v, PetscScalar** a) { *a = v->v.data(); return 0; }
PetscErrorCode VecDestroy(Vec* v) { delete *v; *v = nullptr; return 0; }
PetscErrorCode KSPDestroy(KSP*) { return 0; }
PetscErrorCode ISCreateGeneral(MPI_Comm, PetscInt n, const PetscInt idx[], PetscCopyMode, IS* is) { *is = new _ef_shim_IS; (*is)->idx.assign(idx, idx + n); return 0; }
PetscErrorCode ISCreateStride(MPI_Comm, PetscInt n, PetscInt first, PetscInt step, IS* is) { *is = new _ef_shim_IS; for (int k = 0; k < n; k++) (*is)->idx.push_back(first + k * step); return 0; }
PetscErrorCode ISDestroy(IS* is) { delete *is; *is = nullptr; return 0; }
