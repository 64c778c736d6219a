/*
 * csrc/lpf.c -- host-side low-pass tap designer of libxlating_b200 (PRODUCT code;
 * the tap designer stays on the CPU by design).
 *
 * Behavioural contract = /root/reference/src/lpf.c:
 *   argument checks and "<3>" messages          :12-29
 *   tap count (int)(53*fs/(22.0f*tw)), made odd :31-38
 *   Hamming window, double -> float              :40-51
 *   windowed sinc, double -> float               :70-81
 *   unity-DC-gain normalisation in float         :85-94
 * Bit-identical taps are required for parity of everything downstream, so every
 * intermediate keeps the reference's precision (float where it rounds to float,
 * double where it evaluates in double).  Compiled with -ffp-contract=off.
 */
#include "lpf.h"

#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

static const double XL_PI = 3.14159265358979323846;

static int lpf_args_ok(uint32_t fs, uint32_t cutoff, uint32_t tw) {
  if (fs == 0) {
    fprintf(stderr, "<3>sampling frequency should be positive\n");
    return 0;
  }
  if (cutoff == 0 || (float)cutoff > (float)fs / 2) {
    fprintf(stderr, "<3>cutoff frequency should be positive and less than sampling freq / 2. got: %u\n", cutoff);
    return 0;
  }
  if (tw == 0) {
    fprintf(stderr, "<3>transition width should be positive\n");
    return 0;
  }
  return 1;
}

/* number of taps for a Hamming design (53 dB), always odd */
static int lpf_tap_count(uint32_t fs, uint32_t tw) {
  const float width_term = 22.0F * (float)tw; /* float product, as lpf.c:33 */
  int n = (int)(53.0 * (double)fs / (double)width_term);
  return n | 1;
}

int create_low_pass_filter(float gain, uint32_t sampling_freq, uint32_t cutoff_freq,
                           uint32_t transition_width, float **taps, size_t *len) {
  if (!lpf_args_ok(sampling_freq, cutoff_freq, transition_width)) {
    return -1;
  }
  const int ntaps = lpf_tap_count(sampling_freq, transition_width);
  const int centre = (ntaps - 1) / 2;
  float *h = (float *)malloc(sizeof(float) * (size_t)ntaps);
  if (h == NULL) {
    return -ENOMEM;
  }
  const float omega_c = (float)(2 * XL_PI * (double)cutoff_freq / (double)sampling_freq);
  const double omega_c_d = (double)omega_c;

  /* centre tap, then the two symmetric halves are evaluated tap by tap exactly as
   * a straight -M..M walk would (each tap is an independent expression) */
  for (int k = 0; k < ntaps; k++) {
    const float hamming = (float)(0.54 - 0.46 * cos((2 * XL_PI * k) / (ntaps - 1)));
    const int n = k - centre;
    double ideal;
    if (n == 0) {
      ideal = omega_c_d / XL_PI;
    } else {
      ideal = sin((double)n * omega_c_d) / (n * XL_PI);
    }
    h[k] = (float)(ideal * (double)hamming);
  }

  float dc_gain = h[centre];
  for (int n = 1; n <= centre; n++) {
    dc_gain += 2 * h[centre + n];
  }
  const float scale = gain / dc_gain;
  for (int k = 0; k < ntaps; k++) {
    h[k] *= scale;
  }
  *taps = h;
  *len = (size_t)ntaps;
  return 0;
}
