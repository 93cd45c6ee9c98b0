/* Hand-written p4est configuration for the oracle build (test infrastructure only).
 * Replaces the header generated from extern/p4est/cmake/p4est_config.h.in:
 * 2D only, serial, no zlib. */
#ifndef _SRC_P_EST_CONFIG_H
#define _SRC_P_EST_CONFIG_H 1
#define P4EST_CC "gcc"
#define P4EST_CFLAGS "-O2"
#define P4EST_CPP "gcc -E"
#define P4EST_CPPFLAGS ""
#define P4EST_ENABLE_BUILD_2D 1
#define P4EST_F77_FUNC(name,NAME) name ## _
#define P4EST_F77_FUNC_(name,NAME) name ## _
#define P4EST_FC_FUNC(name,NAME) name ## _
#define P4EST_FC_FUNC_(name,NAME) name ## _
#define P4EST_HAVE_ARPA_INET_H 1
#define P4EST_HAVE_DLFCN_H 1
#define P4EST_HAVE_FSYNC 1
#define P4EST_HAVE_INTTYPES_H 1
#define P4EST_HAVE_MEMORY_H 1
#define P4EST_HAVE_NETINET_IN_H 1
#define P4EST_HAVE_POSIX_MEMALIGN 1
#define P4EST_HAVE_STDINT_H 1
#define P4EST_HAVE_STDLIB_H 1
#define P4EST_HAVE_STRINGS_H 1
#define P4EST_HAVE_STRING_H 1
#define P4EST_HAVE_SYS_STAT_H 1
#define P4EST_HAVE_SYS_TYPES_H 1
#define P4EST_HAVE_UNISTD_H 1
#define P4EST_LDFLAGS ""
#define P4EST_LIBS ""
#define P4EST_PACKAGE "p4est"
#define P4EST_PACKAGE_BUGREPORT "p4est@ins.uni-bonn.de"
#define P4EST_PACKAGE_NAME "p4est"
#define P4EST_PACKAGE_STRING "p4est 0.0.0"
#define P4EST_PACKAGE_TARNAME "p4est"
#define P4EST_PACKAGE_URL ""
#define P4EST_PACKAGE_VERSION "0.0.0"
#define P4EST_VERSION "0.0.0"
#define P4EST_VERSION_MAJOR 0
#define P4EST_VERSION_MINOR 0
#define P4EST_VERSION_POINT 0
#endif
