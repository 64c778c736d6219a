/* oracle/stubs/hackrf.h -- TEST INFRASTRUCTURE.  libhackrf is not in this image; these
 * are the declarations of its public header that the reference's shim
 * (src/sdr/hackrf_lib.h), device wrapper (src/sdr/hackrf_device.c) and mock
 * (test/hackrf_lib_mock.c) need to COMPILE.  No library code: the reference's own mock
 * supplies every function. */
#ifndef XL_STUB_HACKRF_H
#define XL_STUB_HACKRF_H
#include <stdint.h>

enum hackrf_error {
  HACKRF_SUCCESS = 0,
  HACKRF_TRUE = 1,
  HACKRF_ERROR_INVALID_PARAM = -2,
  HACKRF_ERROR_NOT_FOUND = -5,
  HACKRF_ERROR_BUSY = -6,
  HACKRF_ERROR_NO_MEM = -11,
  HACKRF_ERROR_LIBUSB = -1000,
  HACKRF_ERROR_THREAD = -1001,
  HACKRF_ERROR_STREAMING_THREAD_ERR = -1002,
  HACKRF_ERROR_STREAMING_STOPPED = -1003,
  HACKRF_ERROR_STREAMING_EXIT_CALLED = -1004,
  HACKRF_ERROR_USB_API_VERSION = -1005,
  HACKRF_ERROR_NOT_LAST_DEVICE = -2000,
  HACKRF_ERROR_OTHER = -9999
};

typedef struct hackrf_device hackrf_device;

typedef struct {
  hackrf_device *device;
  uint8_t *buffer;
  int buffer_length;
  int valid_length;
  void *rx_ctx;
  void *tx_ctx;
} hackrf_transfer;

typedef int (*hackrf_sample_block_cb_fn)(hackrf_transfer *transfer);
#endif
