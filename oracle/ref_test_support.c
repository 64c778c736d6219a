/*
 * oracle/ref_test_support.c -- TEST INFRASTRUCTURE.  The handful of helpers from the
 * reference's test/utils.h that its unit test test/test_xlating.c needs, so that the
 * reference's OWN test sources can be compiled unmodified (oracle/Makefile) and run
 * against libxlating_b200.so.  The reference's test/utils.c cannot be used: it pulls
 * in libpng, zlib and the server's file writers, none of which exist here.
 *
 * Behaviour restated from /root/reference/test/utils.c:137-196:
 *   inputs   a byte / int16 ramp starting at `input_offset` (int16: centred on len/2)
 *   asserts  element counts must match; floats are compared after truncating x*10000
 *            to int32, int16 outputs exactly
 */
#include <complex.h>
#include <stdint.h>
#include <stdlib.h>
#include <unity.h>

void setup_input_cu8(uint8_t **input, size_t input_offset, size_t len) {
  uint8_t *ramp = (uint8_t *)malloc(len ? len : 1);
  TEST_ASSERT_NOT_NULL(ramp);
  for (size_t k = 0; k < len; k++) ramp[k] = (uint8_t)(input_offset + k);
  *input = ramp;
}

void setup_input_cs8(int8_t **input, size_t input_offset, size_t len) {
  int8_t *ramp = (int8_t *)malloc(len ? len : 1);
  TEST_ASSERT_NOT_NULL(ramp);
  for (size_t k = 0; k < len; k++) ramp[k] = (int8_t)(input_offset + k);
  *input = ramp;
}

void setup_input_cs16(int16_t **input, size_t input_offset, size_t len) {
  int16_t *ramp = (int16_t *)malloc(sizeof(int16_t) * (len ? len : 1));
  TEST_ASSERT_NOT_NULL(ramp);
  const int16_t half = (int16_t)(len / 2);
  for (size_t k = 0; k < len; k++) ramp[k] = (int16_t)((int16_t)(input_offset + k) - half);
  *input = ramp;
}

static int32_t four_decimals(float x) { return (int32_t)(x * 10000); }

void assert_float_array(const float expected[], size_t expected_size, const float *actual, size_t actual_size) {
  TEST_ASSERT_EQUAL_INT(expected_size, actual_size);
  for (size_t k = 0; k < expected_size; k++) TEST_ASSERT_EQUAL_INT(four_decimals(expected[k]), four_decimals(actual[k]));
}

void assert_cf32(const float expected[], size_t expected_size, float complex *actual, size_t actual_size) {
  TEST_ASSERT_EQUAL_INT(expected_size, actual_size);
  for (size_t k = 0; k < expected_size; k++) {
    TEST_ASSERT_EQUAL_INT(four_decimals(expected[2 * k]), four_decimals(crealf(actual[k])));
    TEST_ASSERT_EQUAL_INT(four_decimals(expected[2 * k + 1]), four_decimals(cimagf(actual[k])));
  }
}

void assert_cs16(const int16_t expected[], size_t expected_size, int16_t *actual, size_t actual_size) {
  TEST_ASSERT_EQUAL_INT(expected_size, actual_size);
  for (size_t k = 0; k < 2 * expected_size; k++) TEST_ASSERT_EQUAL_INT(expected[k], actual[k]);
}

/* ---- helpers of the server test (test/utils.c:198-252): the file a client's dsp
 * thread wrote, plain or gzip'd, must hold the expected samples ---- */
#include <stdio.h>
#include <zlib.h>

#include "config.h"

void assert_file(struct server_config *config, int id, const float expected[], size_t expected_size) {
  char path[4096];
  snprintf(path, sizeof(path), "%s/%d.cf32", config->base_path, id);
  FILE *f = fopen(path, "rb");
  TEST_ASSERT_NOT_NULL(f);
  float complex *samples = (float complex *)malloc(sizeof(float complex) * (expected_size + 1));
  TEST_ASSERT_NOT_NULL(samples);
  const size_t got = fread(samples, sizeof(float complex), expected_size + 1, f);
  fclose(f);
  assert_cf32(expected, expected_size, samples, got);
  free(samples);
}

void assert_gzfile(struct server_config *config, int id, const float expected[], size_t expected_size) {
  char path[4096];
  snprintf(path, sizeof(path), "%s/%d.cf32.gz", config->base_path, id);
  gzFile f = gzopen(path, "rb");
  TEST_ASSERT_NOT_NULL(f);
  float complex *samples = (float complex *)malloc(sizeof(float complex) * expected_size);
  TEST_ASSERT_NOT_NULL(samples);
  size_t have = 0;
  const size_t want = sizeof(float complex) * expected_size;
  while (have < want) {
    const int n = gzread(f, (char *)samples + have, (unsigned)(want - have));
    if (n <= 0) break;
    have += (size_t)n;
  }
  gzclose(f);
  TEST_ASSERT_EQUAL_INT(want, have);
  assert_cf32(expected, expected_size, samples, expected_size);
  free(samples);
}
