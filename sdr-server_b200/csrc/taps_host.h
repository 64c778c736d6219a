/*
 * csrc/taps_host.h -- host-side preparation of a client's filter constants.
 * Stays on the CPU (same libm calls as the reference, so the constants are
 * bit-identical): /root/reference/src/xlating.c:519-549.
 */
#ifndef XLATING_B200_TAPS_HOST_H_
#define XLATING_B200_TAPS_HOST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  float *rev_cf32;   /* 2*taps_len floats: reversed band-pass taps, (re,im) interleaved */
  int16_t *rev_q15;  /* 2*taps_len int16: the same taps in Q15 */
  float incr_re, incr_im;       /* oscillator step exp(-j*w0*D) */
  int16_t qincr_re, qincr_im;   /* Q15 oscillator step */
} xl_client_consts;

/* Returns 0, or -ENOMEM.  Free with xl_client_consts_free. */
int xl_client_consts_build(const float *lpf_taps, size_t taps_len, uint32_t decimation,
                           int32_t center_freq, uint32_t sampling_freq, xl_client_consts *out);
void xl_client_consts_free(xl_client_consts *c);

#ifdef __cplusplus
}
#endif
#endif
