/* lis_config.h -- the reference drivers include this under -DHAVE_CONFIG_H
 * (e.g. test/test3.c:27-29).  lis_b200 has no configure step. */
#ifndef LIS_B200_CONFIG_H
#define LIS_B200_CONFIG_H
/* test/test7.c prints a complex literal through <complex.h> when the configure step found it */
#define HAVE_COMPLEX_H 1
#endif
