/* Single-rank MPI shim for the oracle build of the reference (test infrastructure only).
 * Maps the MPI names used by /root/reference/src (census: SURVEY.md 8(c)) onto libsc's
 * built-in serial emulation (extern/p4est/sc/src/sc_mpi.h) and stubs the few it lacks. */
#ifndef EF_ORACLE_MPI_SHIM_H
#define EF_ORACLE_MPI_SHIM_H
#include <sc.h>
#include <string.h>

typedef sc_MPI_Comm MPI_Comm;
typedef sc_MPI_Group MPI_Group;
typedef sc_MPI_Datatype MPI_Datatype;
typedef sc_MPI_Op MPI_Op;
typedef sc_MPI_Status MPI_Status;

#define MPI_COMM_WORLD sc_MPI_COMM_WORLD
#define MPI_COMM_SELF sc_MPI_COMM_SELF
#define MPI_COMM_NULL sc_MPI_COMM_NULL
#define MPI_UNDEFINED sc_MPI_UNDEFINED
#define MPI_STATUS_IGNORE sc_MPI_STATUS_IGNORE
#define MPI_CHAR sc_MPI_CHAR
#define MPI_SIGNED_CHAR sc_MPI_SIGNED_CHAR
#define MPI_UNSIGNED_CHAR sc_MPI_UNSIGNED_CHAR
#define MPI_WCHAR sc_MPI_INT
#define MPI_SHORT sc_MPI_SHORT
#define MPI_UNSIGNED_SHORT sc_MPI_UNSIGNED_SHORT
#define MPI_INT sc_MPI_INT
#define MPI_UNSIGNED sc_MPI_UNSIGNED
#define MPI_LONG sc_MPI_LONG
#define MPI_UNSIGNED_LONG sc_MPI_UNSIGNED_LONG
#define MPI_UNSIGNED_LONG_LONG sc_MPI_UNSIGNED_LONG_LONG
#define MPI_FLOAT sc_MPI_FLOAT
#define MPI_DOUBLE sc_MPI_DOUBLE
#define MPI_LONG_DOUBLE sc_MPI_LONG_DOUBLE
#define MPI_CXX_BOOL sc_MPI_UNSIGNED_CHAR
#define MPI_MAX sc_MPI_MAX
#define MPI_MIN sc_MPI_MIN
#define MPI_SUM sc_MPI_SUM

#define MPI_Init sc_MPI_Init
#define MPI_Finalize sc_MPI_Finalize
#define MPI_Comm_rank sc_MPI_Comm_rank
#define MPI_Comm_size sc_MPI_Comm_size
#define MPI_Barrier sc_MPI_Barrier
#define MPI_Bcast sc_MPI_Bcast
#define MPI_Allreduce sc_MPI_Allreduce
#define MPI_Allgather sc_MPI_Allgather
#define MPI_Allgatherv sc_MPI_Allgatherv
#define MPI_Comm_split sc_MPI_Comm_split
#define MPI_Comm_group sc_MPI_Comm_group
#define MPI_Group_free sc_MPI_Group_free
#define MPI_Comm_free sc_MPI_Comm_free
#define MPI_Wtime sc_MPI_Wtime

/* single rank: a point-to-point message can never be matched */
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) { (void)b;(void)n;(void)t;(void)dst;(void)tag;(void)c; return 0; }
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* s) { (void)b;(void)n;(void)t;(void)src;(void)tag;(void)c;(void)s; return 0; }
static inline int MPI_Group_range_incl(MPI_Group g, int n, int ranges[][3], MPI_Group* out) { (void)n;(void)ranges; *out = g; return 0; }
static inline int MPI_Comm_create_group(MPI_Comm c, MPI_Group g, int tag, MPI_Comm* out) { (void)g;(void)tag; *out = c; return 0; }
static inline int MPI_Comm_set_name(MPI_Comm c, const char* name) { (void)c;(void)name; return 0; }
static inline int MPI_Comm_get_name(MPI_Comm c, char* name, int* len) { (void)c; name[0] = 0; *len = 0; return 0; }
static inline int MPI_Initialized(int* flag) { *flag = 1; return 0; }
static inline int MPI_Finalized(int* flag) { *flag = 0; return 0; }
#ifndef MPI_MAX_OBJECT_NAME
#define MPI_MAX_OBJECT_NAME 128
#endif
#endif
