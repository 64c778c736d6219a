/*
 * oracle/lpf_oracle.c -- restatement of the reference low-pass tap designer.
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * Follows /root/reference/src/lpf.c:
 *   :12-29  argument checks
 *   :31-38  tap count  ntaps = (int)(53 * fs / (22.0f * tw)), forced odd
 *   :40-51  Hamming window, evaluated in double, stored as float
 *   :70-81  windowed sinc, evaluated in double, stored as float
 *   :85-98  DC-gain normalisation, evaluated in float
 * The precision of every intermediate (float vs double) is part of the
 * contract: the taps must come out bit-identical.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static int orc_lpf_ntaps(uint32_t fs, uint32_t tw) {
  /* lpf.c:33 -- the denominator is a FLOAT product, the quotient a double */
  float denom = 22.0f * (float)tw;
  double q = 53.0 * (double)fs / (double)denom;
  int n = (int)q;
  return (n % 2 == 0) ? n + 1 : n;
}

int orc_lpf_design(float gain, uint32_t fs, uint32_t cutoff, uint32_t tw,
                   float **taps_out, size_t *len_out) {
  if (fs == 0) return -1;                                   /* lpf.c:13 */
  if (cutoff == 0 || (float)cutoff > (float)fs / 2) return -1; /* lpf.c:18 */
  if (tw == 0) return -1;                                   /* lpf.c:23 */

  const int ntaps = orc_lpf_ntaps(fs, tw);
  const int half = (ntaps - 1) / 2;
  float *h = (float *)malloc(sizeof(float) * (size_t)ntaps);
  if (h == NULL) return -1;

  /* lpf.c:67  cutoff in rad/sample, double expression rounded to float */
  const float wc = (float)(2 * M_PI * (double)cutoff / (double)fs);

  for (int k = 0; k < ntaps; k++) {
    /* lpf.c:45-48 Hamming, double -> float */
    const float win = (float)(0.54 - 0.46 * cos((2 * M_PI * k) / (ntaps - 1)));
    const int n = k - half;
    if (n == 0) {
      h[k] = (float)((double)wc / M_PI * (double)win);      /* lpf.c:72 */
    } else {
      h[k] = (float)(sin((double)n * (double)wc) / (n * M_PI) * (double)win); /* lpf.c:75 */
    }
  }

  /* lpf.c:85-88: float accumulation of the (symmetric) DC gain */
  float dc = h[half];
  for (int n = 1; n <= half; n++) {
    dc += 2 * h[half + n];
  }
  gain /= dc;
  for (int k = 0; k < ntaps; k++) h[k] *= gain;            /* lpf.c:92-94 */

  *taps_out = h;
  *len_out = (size_t)ntaps;
  return 0;
}
