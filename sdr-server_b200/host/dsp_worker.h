/*
 * host/dsp_worker.h -- the reference's thread-per-client model on top of the batch
 * GPU engine (include/xlating_group.h).
 *
 * Mirrors src/dsp_worker.h:41-45 (dsp_worker_start / dsp_worker_process /
 * dsp_worker_destroy) and the callback loop at src/dsp_worker.c:41-88, with one
 * change of substance: the per-client CPU filter call (:55-72) is gone.  The ingest
 * side submitted the SDR block ONCE for all clients (xl_stream_push, the counterpart
 * of sdr_callback, src/tcp_server.c:257-271); a client's dsp thread only takes a
 * ticket from its queue, waits for the GPU, and writes ITS output to the socket or
 * file exactly as the reference does (:10-39, :73-85).
 */
#ifndef XL_DSP_WORKER_H_
#define XL_DSP_WORKER_H_

#include <stdint.h>

#include "xlating_group.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { XL_DESTINATION_FILE = 0, XL_DESTINATION_SOCKET = 1 }; /* src/api.h REQUEST_DESTINATION_* */

/* the fields of the reference's client_config that the dsp path uses (src/dsp_worker.h:14-23) */
typedef struct {
  uint32_t center_freq;
  uint32_t sampling_rate;
  uint32_t band_freq;
  uint8_t destination;
  int client_socket; /* any writable fd */
  uint32_t id;
} xl_client_config;

typedef struct xl_dsp_worker xl_dsp_worker;

/* Designs the client's low-pass (create_low_pass_filter(1.0, band_rate, rate/2,
 * rate/lpf_cutoff_rate), src/dsp_worker.c:98), attaches it to the group, opens
 * <base_path>/<id>.cf32 -- or <id>.cf32.gz with use_gzip -- for file destinations
 * (:126-144) and starts the thread.  max_block_elements sizes the client's private
 * output buffer (largest block, in scalar elements). */
int xl_dsp_worker_start(const xl_client_config *config, xlg_group *group, uint32_t band_sampling_rate,
                        uint32_t max_block_elements, int lpf_cutoff_rate, int queue_size, const char *base_path,
                        int use_gzip, xl_dsp_worker **worker);

/* called by the ingest side for every submitted block (dsp_worker_process, :202-204) */
void xl_dsp_worker_post(xl_dsp_worker *worker, int64_t ticket);

/* poison pill, join, detach from the group, close the file (:172-200) */
void xl_dsp_worker_destroy(xl_dsp_worker *worker);

uint64_t xl_dsp_worker_blocks_written(xl_dsp_worker *worker);
uint64_t xl_dsp_worker_blocks_lost(xl_dsp_worker *worker);
uint64_t xl_dsp_worker_blocks_failed(xl_dsp_worker *worker);  /* write errors (socket closed, disk full) */
uint64_t xl_dsp_worker_queue_overruns(xl_dsp_worker *worker); /* tickets overwritten in a full queue */

#ifdef __cplusplus
}
#endif
#endif
