/*
 * host/stream.h -- ingest side of the server model: one wideband SDR stream, its
 * GPU group and the clients attached to it.  It is the part of
 * src/tcp_server.c that touches the data path -- sdr_callback (:257-271) and the
 * client add/remove bookkeeping around dsp_worker_start/destroy (:301-384,
 * :178-189) -- without the sockets, protocol and SDR drivers, which are out of
 * scope.
 */
#ifndef XL_STREAM_H_
#define XL_STREAM_H_

#include <stdint.h>

#include "dsp_worker.h"

#ifdef __cplusplus
extern "C" {
#endif

/* the subset of struct server_config the data path reads (src/config.h) */
typedef struct {
  int sdr_type;                /* XLG_FMT_CU8 (rtl-sdr), XLG_FMT_CS8 (hackrf), XLG_FMT_CS16 (airspy) */
  uint32_t band_sampling_rate;
  uint32_t buffer_size;        /* bytes per SDR block, default 262144 (src/config.c:208) */
  int queue_size;              /* default 64 (src/config.c:183) */
  int lpf_cutoff_rate;         /* default 5  (src/config.c:213) */
  const char *base_path;       /* file destinations */
  int device;                  /* CUDA device ordinal */
  int use_gzip;                /* file destinations are <id>.cf32.gz (src/config.c:250, default true there) */
} xl_stream_config;

typedef struct xl_stream xl_stream;

int xl_stream_create(const xl_stream_config *config, xl_stream **stream);
/* handle_new_client's data-path half: start a dsp worker for this client */
int xl_stream_add_client(xl_stream *stream, const xl_client_config *client);
int xl_stream_remove_client(xl_stream *stream, uint32_t client_id);
/* sdr_callback: one submit for all clients, then an 8-byte ticket per client */
int xl_stream_push(xl_stream *stream, const uint8_t *buf, uint32_t buf_len);
/* wait until every posted block has been accounted for by every client -- written, lost
 * (-ESTALE), failed to write, or overwritten in a full queue -- or 30 s have passed
 * (tests, shutdown).  Returns 0, or -ETIMEDOUT. */
int xl_stream_flush(xl_stream *stream);
void xl_stream_destroy(xl_stream *stream);
int xl_stream_client_count(xl_stream *stream);

#ifdef __cplusplus
}
#endif
#endif
