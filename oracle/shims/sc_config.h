/* Hand-written libsc configuration for the oracle build (test infrastructure only).
 * Replaces the header the reference's cmake would generate from
 * extern/p4est/sc/cmake/sc_config.h.in: serial (no MPI), no zlib, no threads. */
#ifndef _SRC_SC_CONFIG_H
#define _SRC_SC_CONFIG_H 1
#define SC_CC "gcc"
#define SC_CFLAGS "-O2"
#define SC_CPP "gcc -E"
#define SC_CPPFLAGS ""
#define SC_ENABLE_USE_COUNTERS 1
#define SC_ENABLE_USE_REALLOC 1
#define SC_HAVE_ALIGNED_ALLOC 1
#define SC_HAVE_FCNTL_H 1
#define SC_HAVE_FSYNC 1
#define SC_HAVE_INTTYPES_H 1
#define SC_HAVE_MEMORY_H 1
#define SC_HAVE_POSIX_MEMALIGN 1
#define SC_HAVE_FABS 1
#define SC_HAVE_QSORT_R 1
#define SC_HAVE_GNU_QSORT_R 1
#define SC_HAVE_SIGNAL_H 1
#define SC_HAVE_STDINT_H 1
#define SC_HAVE_STDLIB_H 1
#define SC_HAVE_STRING_H 1
#define SC_HAVE_LIBGEN_H 1
#define SC_HAVE_STRTOLL 1
#define SC_HAVE_SYS_IOCTL_H 1
#define SC_HAVE_SYS_SELECT_H 1
#define SC_HAVE_SYS_STAT_H 1
#define SC_HAVE_SYS_TIME_H 1
#define SC_HAVE_SYS_TYPES_H 1
#define SC_HAVE_TIME_H 1
#define SC_MEMALIGN_BYTES (SC_SIZEOF_VOID_P)
#define SC_HAVE_UNISTD_H 1
#define SC_LDFLAGS ""
#define SC_LIBS ""
#define SC_PACKAGE "libsc"
#define SC_PACKAGE_BUGREPORT "p4est@ins.uni-bonn.de"
#define SC_PACKAGE_NAME "libsc"
#define SC_PACKAGE_STRING "libsc 0.0.0"
#define SC_PACKAGE_TARNAME "libsc"
#define SC_PACKAGE_URL ""
#define SC_PACKAGE_VERSION "0.0.0"
#define SC_SIZEOF_INT 4
#define SC_SIZEOF_UNSIGNED_INT 4
#define SC_SIZEOF_LONG 8
#define SC_SIZEOF_LONG_LONG 8
#define SC_SIZEOF_UNSIGNED_LONG 8
#define SC_SIZEOF_UNSIGNED_LONG_LONG 8
#define SC_SIZEOF_VOID_P 8
#define SC_VERSION "0.0.0"
#define SC_VERSION_MAJOR 0
#define SC_VERSION_MINOR 0
#define SC_VERSION_POINT 0
#endif
