/*
 * csrc/xl_log.h -- error reporting of libxlating_b200: one "<3>..." line on stderr, the
 * reference's journald convention (src/dsp_worker.c:17).
 */
#ifndef XLATING_B200_LOG_H_
#define XLATING_B200_LOG_H_

#include <stdio.h>

#define XL_LOG(...)                                   \
  do {                                                \
    fprintf(stderr, "<3>xlating_b200: " __VA_ARGS__); \
    fprintf(stderr, "\n");                            \
  } while (0)

#endif
