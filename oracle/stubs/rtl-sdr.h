/* oracle/stubs/rtl-sdr.h -- TEST INFRASTRUCTURE.  librtlsdr is not in this image; this
 * is the one declaration of its public header that the reference's driver shim
 * (src/sdr/rtlsdr_lib.h) and mock (test/rtlsdr_lib_mock.c) need in order to COMPILE.
 * No library code: the reference's own mock supplies every function. */
#ifndef XL_STUB_RTL_SDR_H
#define XL_STUB_RTL_SDR_H
#include <stdint.h>
typedef struct rtlsdr_dev rtlsdr_dev_t;
#endif
