/*
 * include/lpf.h -- host-side low-pass tap designer (stays on the CPU, as
 * BASELINE.json's north_star requires).  Same symbol and contract as the
 * reference's src/lpf.h:6 / src/lpf.c:53-99 so src/dsp_worker.c:98 links
 * unchanged: Hamming-windowed sinc, ntaps = (int)(53*fs/(22*tw)) forced odd,
 * unity DC gain times `gain`.  Returns 0 and a malloc()'d vector the caller (or
 * the filter that adopts it) frees; -1 with a "<3>" log on bad arguments;
 * -ENOMEM.  The taps are bit-identical to the reference's (same float/double
 * evaluation order), which tests/test_host_taps.py checks.
 */
#ifndef XLATING_B200_LPF_H_
#define XLATING_B200_LPF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int create_low_pass_filter(float gain, uint32_t sampling_freq, uint32_t cutoff_freq,
                           uint32_t transition_width, float **taps, size_t *len);

#ifdef __cplusplus
}
#endif
#endif /* XLATING_B200_LPF_H_ */
