/* Hand-written stand-in for the autoconf-generated lis_config.h of the reference
 * (include/lis_config.h.in).  Default double-precision, 32-bit LIS_INT, no MPI, no quad, no
 * Fortran, no SA-AMG -- i.e. what `./configure [--enable-omp]` produces on x86-64 Linux.
 * The C sources only test HAVE_MALLOC_H; the rest is listed for completeness. */
#ifndef LIS_REF_CONFIG_H
#define LIS_REF_CONFIG_H
#define HAVE_MALLOC_H 1
#define HAVE_STDIO_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_UNISTD_H 1
#define STDC_HEADERS 1
#define PACKAGE "lis"
#define VERSION "2.1.11"
#endif
