/*
 * csrc/xlating_dropin.cpp -- the reference's per-filter C ABI (include/xlating.h)
 * on top of the batch engine: each filter is a private one-client group with its
 * own CUDA streams, so filters owned by different dsp threads run concurrently,
 * exactly like the reference's thread-per-client model (src/dsp_worker.c:41-88).
 *
 * Replaces, symbol for symbol: src/xlating.c:495-582 (create), :384-447 and
 * :352-382 (the twelve process_* entry points), :584-616 (destroy) and the
 * SIMD_STATUS string (:145-156, :268).
 *
 * A process_* call is synchronous, as in the reference: stage the block
 * (pinned copy + async H2D), run convert -> oscillator pre-pass -> FIR on the
 * device, copy the outputs back into filter-owned pinned memory and return a
 * pointer to it.  There is no CPU implementation behind these symbols.
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>

#include "xlating.h"
#include "xlating_group.h"

extern "C" {

const char *SIMD_STATUS = "CUDA sm_100a";

struct xlating_t {
  xlg_group *group;
  int client;
  float *adopted_taps;  // freed on destroy, like src/xlating.c:600-602
};

int create_frequency_xlating_filter(uint32_t decimation, float *taps, size_t taps_len, int32_t center_freq,
                                    uint32_t sampling_freq, uint32_t max_input_buffer_length, xlating **filter) {
  if (taps_len == 0) {
    return -1;  // src/xlating.c:496-498 (taps NOT adopted on this path)
  }
  struct xlating_t *f = (struct xlating_t *)calloc(1, sizeof(*f));
  if (f == NULL) {
    return -ENOMEM;
  }
  f->adopted_taps = taps;
  int device = 0;
  const char *env = getenv("XLATING_B200_DEVICE");
  if (env != NULL) {
    device = atoi(env);
  }
  const uint32_t max_in = max_input_buffer_length < 2 ? 2 : max_input_buffer_length;
  int rc = xlg_create(device, sampling_freq, max_in, 0, &f->group);
  if (rc != 0) {
    destroy_xlating(f);
    return rc;
  }
  rc = xlg_add_client(f->group, decimation, taps, taps_len, center_freq, &f->client);
  if (rc != 0) {
    destroy_xlating(f);
    return rc;
  }
  *filter = f;
  return 0;
}

void destroy_xlating(xlating *filter) {
  if (filter == NULL) {
    return;
  }
  if (filter->group != NULL) {
    xlg_destroy(filter->group);
  }
  if (filter->adopted_taps != NULL) {
    free(filter->adopted_taps);
  }
  free(filter);
}

static void run_block(xlating *f, int fmt, const void *input, size_t input_len, uint32_t path, void **output,
                      size_t *output_len) {
  *output_len = 0;
  const int64_t ticket = xlg_submit(f->group, fmt, input, input_len, path);
  if (ticket < 0) {
    fprintf(stderr, "<3>xlating_b200: block dropped (submit -> %lld)\n", (long long)ticket);
    return;
  }
  int rc = xlg_wait(f->group, ticket);
  if (rc != 0) {
    fprintf(stderr, "<3>xlating_b200: block dropped (wait -> %d)\n", rc);
    return;
  }
  const void *out = NULL;
  size_t n = 0;
  rc = xlg_output(f->group, ticket, f->client, &out, &n);
  if (rc != 0) {
    fprintf(stderr, "<3>xlating_b200: block dropped (output -> %d)\n", rc);
    return;
  }
  *output = (void *)out;
  *output_len = n;
}

#define XL_DEFINE_CF32(variant, name, ctype, fmt)                                                              \
  void process_##variant##_##name##_cf32(const ctype *input, size_t input_len, xlating_cf32 **output,           \
                                         size_t *output_len, xlating *filter) {                                 \
    run_block(filter, fmt, input, input_len, 0, (void **)output, output_len);                                   \
  }
#define XL_DEFINE_Q15(variant, name, ctype, fmt)                                                               \
  void process_##variant##_##name##_cs16(const ctype *input, size_t input_len, int16_t **output,                \
                                         size_t *output_len, xlating *filter) {                                 \
    run_block(filter, fmt, input, input_len, XLG_PATH_Q15, (void **)output, output_len);                        \
  }

XL_DEFINE_CF32(native, cu8, uint8_t, XLG_FMT_CU8)
XL_DEFINE_CF32(native, cs8, int8_t, XLG_FMT_CS8)
XL_DEFINE_CF32(native, cs16, int16_t, XLG_FMT_CS16)
XL_DEFINE_CF32(optimized, cu8, uint8_t, XLG_FMT_CU8)
XL_DEFINE_CF32(optimized, cs8, int8_t, XLG_FMT_CS8)
XL_DEFINE_CF32(optimized, cs16, int16_t, XLG_FMT_CS16)
XL_DEFINE_Q15(native, cu8, uint8_t, XLG_FMT_CU8)
XL_DEFINE_Q15(native, cs8, int8_t, XLG_FMT_CS8)
XL_DEFINE_Q15(native, cs16, int16_t, XLG_FMT_CS16)
XL_DEFINE_Q15(optimized, cu8, uint8_t, XLG_FMT_CU8)
XL_DEFINE_Q15(optimized, cs8, int8_t, XLG_FMT_CS8)
XL_DEFINE_Q15(optimized, cs16, int16_t, XLG_FMT_CS16)

}  // extern "C"
