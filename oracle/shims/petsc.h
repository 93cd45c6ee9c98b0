/* Sequential PETSc shim for the oracle build of the reference (test infrastructure only).
 * PETSc is an un-vendored ExternalProject of the reference (CMakeLists.txt:150-160, tag
 * "release", unpinned).  The hot path touches it at exactly one place: the FivePointStencil
 * leaf solve (src/Patches/FiniteVolume/FiniteVolumeSolver.cpp:44-57,185-219), which builds a
 * sequential MATDENSE, MatLUFactor + MatSolve (PETSc forwards those to LAPACK dgetrf/dgetrs).
 * This header provides that behaviour with the same names; everything else the reference
 * merely *declares* against PETSc (ParallelMatrix/ParallelVector, never used by HPSAlgorithm)
 * is given inert definitions so the headers compile unchanged. */
#ifndef EF_ORACLE_PETSC_SHIM_H
#define EF_ORACLE_PETSC_SHIM_H
#include <mpi.h>
#include <vector>
#include <cstddef>
#include <functional>
#include <cmath>
#include <cstring>
#include <string>
#include <map>

typedef int PetscInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef int PetscErrorCode;
typedef int PetscBool;
typedef const char* MatType;
typedef const char* VecType;
#define MATDENSE "dense"
#define MATAIJ "aij"
#define MATMPIAIJ "mpiaij"
#define MATMPIDENSE "mpidense"
#define MATSEQDENSE "seqdense"
#define VECSTANDARD "standard"
#define VECMPI "mpi"
#define VECSEQ "seq"
#define PETSC_DECIDE (-1)
#define PETSC_DETERMINE (-1)
#define PETSC_NULLPTR nullptr
#define PETSC_TRUE 1
#define PETSC_FALSE 0
enum InsertMode { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES };
enum MatAssemblyType { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 };
enum MatReuse { MAT_INITIAL_MATRIX, MAT_REUSE_MATRIX, MAT_IGNORE_MATRIX, MAT_INPLACE_MATRIX };
enum PetscCopyMode { PETSC_COPY_VALUES, PETSC_OWN_POINTER, PETSC_USE_POINTER };

struct _ef_shim_Mat { int m = 0, n = 0; std::vector<double> a; /* column-major */ std::vector<int> ipiv; bool factored = false; };
struct _ef_shim_Vec { int n = 0; std::vector<double> v; };
struct _ef_shim_IS { std::vector<int> idx; };
typedef _ef_shim_Mat* Mat;
typedef _ef_shim_Vec* Vec;
typedef _ef_shim_IS* IS;
typedef void* KSP;
typedef void* PC;
struct MatFactorInfo { double fill = 0; };

PetscErrorCode PetscInitialize(int* argc, char*** argv, const char* file, const char* help);
PetscErrorCode PetscFinalize();
PetscErrorCode PetscGetArgs(int* argc, char*** argv);

PetscErrorCode MatCreate(MPI_Comm comm, Mat* A);
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N);
PetscErrorCode MatSetType(Mat A, MatType type);
PetscErrorCode MatSetFromOptions(Mat A);
PetscErrorCode MatSetUp(Mat A);
PetscErrorCode MatSetValue(Mat A, PetscInt i, PetscInt j, PetscScalar v, InsertMode mode);
PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt idxm[], PetscInt n, const PetscInt idxn[], const PetscScalar v[], InsertMode mode);
PetscErrorCode MatGetValue(Mat A, PetscInt i, PetscInt j, PetscScalar* v);
PetscErrorCode MatGetValues(Mat A, PetscInt m, const PetscInt idxm[], PetscInt n, const PetscInt idxn[], PetscScalar v[]);
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t);
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t);
PetscErrorCode MatGetSize(Mat A, PetscInt* M, PetscInt* N);
PetscErrorCode MatGetLocalSize(Mat A, PetscInt* m, PetscInt* n);
PetscErrorCode MatGetOwnershipRange(Mat A, PetscInt* first, PetscInt* last);
PetscErrorCode MatFactorInfoInitialize(MatFactorInfo* info);
PetscErrorCode MatLUFactor(Mat A, IS row, IS col, const MatFactorInfo* info);
PetscErrorCode MatSolve(Mat A, Vec b, Vec x);
PetscErrorCode MatDestroy(Mat* A);
PetscErrorCode MatCreateConstantDiagonal(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt M, PetscInt N, PetscScalar diag, Mat* J);
PetscErrorCode MatCreateSubMatrix(Mat A, IS isrow, IS iscol, MatReuse cll, Mat* newmat);
PetscErrorCode MatCreateSubMatrices(Mat A, PetscInt n, const IS irow[], const IS icol[], MatReuse scall, Mat* submat[]);
PetscErrorCode MatCreateMPIMatConcatenateSeqMat(MPI_Comm comm, Mat seqmat, PetscInt n, MatReuse reuse, Mat* mpimat);

PetscErrorCode VecCreate(MPI_Comm comm, Vec* v);
PetscErrorCode VecSetSizes(Vec v, PetscInt n, PetscInt N);
PetscErrorCode VecSetType(Vec v, VecType t);
PetscErrorCode VecSetFromOptions(Vec v);
PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar y, InsertMode mode);
PetscErrorCode VecSetValues(Vec v, PetscInt ni, const PetscInt ix[], const PetscScalar y[], InsertMode mode);
PetscErrorCode VecAssemblyBegin(Vec v);
PetscErrorCode VecAssemblyEnd(Vec v);
PetscErrorCode VecDuplicate(Vec v, Vec* newv);
PetscErrorCode VecGetArray(Vec v, PetscScalar** a);
PetscErrorCode VecDestroy(Vec* v);
PetscErrorCode KSPDestroy(KSP* ksp);

PetscErrorCode ISCreateGeneral(MPI_Comm comm, PetscInt n, const PetscInt idx[], PetscCopyMode mode, IS* is);
PetscErrorCode ISCreateStride(MPI_Comm comm, PetscInt n, PetscInt first, PetscInt step, IS* is);
PetscErrorCode ISDestroy(IS* is);
#endif
