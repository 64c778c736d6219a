/* oracle/stubs/airspy.h -- TEST INFRASTRUCTURE.  libairspy is not in this image; these
 * are the declarations of its public header (airspy.h, public API of libairspy 1.0.x)
 * that the reference's shim (src/sdr/airspy_lib.h), device wrapper
 * (src/sdr/airspy_device.c) and mock (test/airspy_lib_mock.c) need to COMPILE.  No
 * library code: the reference's own mock supplies every function. */
#ifndef XL_STUB_AIRSPY_H
#define XL_STUB_AIRSPY_H
#include <stdint.h>

enum airspy_error { AIRSPY_SUCCESS = 0, AIRSPY_TRUE = 1, AIRSPY_ERROR_INVALID_PARAM = -2, AIRSPY_ERROR_OTHER = -9999 };

enum airspy_sample_type {
  AIRSPY_SAMPLE_FLOAT32_IQ = 0,
  AIRSPY_SAMPLE_FLOAT32_REAL = 1,
  AIRSPY_SAMPLE_INT16_IQ = 2,
  AIRSPY_SAMPLE_INT16_REAL = 3,
  AIRSPY_SAMPLE_UINT16_REAL = 4,
  AIRSPY_SAMPLE_RAW = 5,
  AIRSPY_SAMPLE_END = 6
};

struct airspy_device;

typedef struct {
  struct airspy_device *device;
  void *ctx;
  void *samples;
  int sample_count;
  uint64_t dropped_samples;
  enum airspy_sample_type sample_type;
} airspy_transfer_t, airspy_transfer;

typedef struct {
  uint32_t major_version;
  uint32_t minor_version;
  uint32_t revision;
} airspy_lib_version_t;

typedef int (*airspy_sample_block_cb_fn)(airspy_transfer *transfer);
#endif
