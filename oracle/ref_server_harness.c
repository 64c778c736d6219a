/*
 * oracle/ref_server_harness.c -- TEST INFRASTRUCTURE.  Drives the reference's REAL
 * per-client machinery -- src/dsp_worker.c and src/queue.c, compiled unmodified where
 * they lie -- the way its tcp_server does, without the TCP server and the SDR drivers
 * (which need libraries this image does not have):
 *
 *   dsp_worker_start()            once per client   (src/tcp_server.c:327)
 *   dsp_worker_process(block)     every block to every client, i.e. the fan-out of
 *                                 sdr_callback (src/tcp_server.c:257-271)
 *   dsp_worker_destroy()          drains the client's queue, joins its dsp thread
 *
 * oracle/build_ref.sh links it with the reference's src/xlating.c + src/lpf.c
 * (server_harness_ref), with libxlating_b200.so instead (server_harness_b200; _pinnedq =
 * with sdr-server_b200/host/queue_pinned.c in place of src/queue.c), and -- compiled with
 * -DHARNESS_CUDA_CF32 against the reference tree PATCHED with integration/cuda_cf32.patch --
 * in the patch's cpu_optimization = CUDA_CF32 mode (server_harness_patched_b200): one
 * xlg_submit per block, an 8-byte ticket per client, exactly what the patched
 * sdr_callback does (src/tcp_server.c:257-271 after the patch).
 * Both write <base_path>/<id>.cf32 exactly as the server does, so the two can be
 * compared file by file, and both print (last line of stdout) the wall-clock input rate
 * they sustained.
 *
 * usage: server_harness <clients> <blocks> <queue_size> <base_path> [rtl|airspy|hackrf] [dump]
 *   queue_size >= blocks keeps the reference queue from overwriting blocks
 *   (src/queue.c:90-94), which is what makes the outputs comparable.
 *   `dump` also writes the input blocks to <base_path>/input.raw.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "dsp_worker.h"
#include "api.h"

extern const char *SIMD_STATUS; /* src/xlating.c:145-156 */

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <clients> <blocks> <queue_size> <base_path> [rtl|airspy|hackrf] [dump]\n", argv[0]);
    return 2;
  }
  const int n_clients = atoi(argv[1]);
  const int n_blocks = atoi(argv[2]);
  struct server_config server;
  memset(&server, 0, sizeof(server));
#ifdef HARNESS_CUDA_CF32
  server.optimization = CUDA_CF32; /* integration/cuda_cf32.patch */
#else
  server.optimization = OPTIMIZED_CF32; /* the shipped default, src/config.c:252-264 */
#endif
  server.sdr_type = SDR_TYPE_RTL;
  server.band_sampling_rate = 2016000;
  if (argc > 5 && strcmp(argv[5], "airspy") == 0) {
    server.sdr_type = SDR_TYPE_AIRSPY;
    server.band_sampling_rate = 10000000;
  } else if (argc > 5 && strcmp(argv[5], "hackrf") == 0) {
    server.sdr_type = SDR_TYPE_HACKRF;
  }
  const int dump = argc > 6 && strcmp(argv[6], "dump") == 0;
  server.buffer_size = 262144; /* src/config.c:208 */
  server.queue_size = atoi(argv[3]);
  server.lpf_cutoff_rate = 5;
  server.base_path = argv[4];
  server.use_gzip = false;

  /* distinct pseudo-random blocks, the same in every build of this harness */
  uint8_t *blocks = (uint8_t *)malloc((size_t)n_blocks * server.buffer_size);
  if (blocks == NULL) return 1;
  uint64_t s = 0x9E3779B97F4A7C15ull;
  for (size_t i = 0; i < (size_t)n_blocks * server.buffer_size; i++) {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    blocks[i] = (uint8_t)s;
  }
  if (server.sdr_type == SDR_TYPE_AIRSPY) {
    /* 12-bit-ish int16 samples: keep 14 bits of every word, sign-extended */
    int16_t *w = (int16_t *)blocks;
    for (size_t i = 0; i < (size_t)n_blocks * server.buffer_size / 2; i++) w[i] = (int16_t)(w[i] >> 2);
  }
  if (dump) {
    char path[4096];
    snprintf(path, sizeof(path), "%s/input.raw", server.base_path);
    FILE *f = fopen(path, "wb");
    if (f == NULL || fwrite(blocks, server.buffer_size, (size_t)n_blocks, f) != (size_t)n_blocks) return 1;
    fclose(f);
  }

#ifdef HARNESS_CUDA_CF32
  /* what the patched handle_new_client does for the first client (src/tcp_server.c after the patch) */
  xlg_group *group = NULL;
  {
    const uint32_t max_elements = server.sdr_type == SDR_TYPE_AIRSPY ? server.buffer_size / 2 : server.buffer_size;
    const int code = xlg_create_ex(0, server.band_sampling_rate, max_elements, 0, (uint32_t)server.queue_size, &group);
    if (code != 0) {
      fprintf(stderr, "xlg_create_ex -> %d\n", code);
      return 1;
    }
  }
#endif
  const uint32_t band_freq = 100000000u;
  const uint32_t fs = server.band_sampling_rate;
  client_config *configs = (client_config *)calloc((size_t)n_clients, sizeof(client_config));
  dsp_worker **workers = (dsp_worker **)calloc((size_t)n_clients, sizeof(dsp_worker *));
  for (int c = 0; c < n_clients; c++) {
    const uint32_t rate = fs == 10000000 ? 250000u : ((c % 2 == 0) ? 48000u : 96000u);
    const int64_t offset = -(int64_t)fs / 2 + rate / 2 + (int64_t)c * (fs - rate) / (n_clients > 1 ? n_clients - 1 : 1);
    configs[c].center_freq = (uint32_t)((int64_t)band_freq + offset);
    configs[c].sampling_rate = rate;
    configs[c].band_freq = band_freq;
    configs[c].destination = REQUEST_DESTINATION_FILE;
    configs[c].client_socket = -1;
    configs[c].id = (uint32_t)c;
    configs[c].sdr_type = server.sdr_type;
    configs[c].is_running = true;
    const int code = dsp_worker_start(&configs[c], &server, &workers[c]);
    if (code != 0) {
      fprintf(stderr, "dsp_worker_start(client %d) -> %d\n", c, code);
      return 1;
    }
#ifdef HARNESS_CUDA_CF32
    if (dsp_worker_attach_group(workers[c], group) != 0) {
      fprintf(stderr, "dsp_worker_attach_group(client %d) failed\n", c);
      return 1;
    }
#endif
  }

  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
#ifdef HARNESS_CUDA_CF32
  const int fmt = server.sdr_type == SDR_TYPE_AIRSPY ? XLG_FMT_CS16 : (server.sdr_type == SDR_TYPE_HACKRF ? XLG_FMT_CS8 : XLG_FMT_CU8);
  const size_t elements = server.sdr_type == SDR_TYPE_AIRSPY ? server.buffer_size / 2 : server.buffer_size;
  for (int b = 0; b < n_blocks; b++) { /* the patched sdr_callback: ONE submit, a ticket per client */
    const int64_t ticket = xlg_submit(group, fmt, blocks + (size_t)b * server.buffer_size, elements, 0);
    if (ticket < 0) {
      fprintf(stderr, "xlg_submit -> %lld\n", (long long)ticket);
      return 1;
    }
    for (int c = 0; c < n_clients; c++) dsp_worker_process_ticket(ticket, workers[c]);
  }
  for (int c = 0; c < n_clients; c++) dsp_worker_destroy(workers[c]); /* drain + join */
  xlg_destroy(group);
#else
  for (int b = 0; b < n_blocks; b++) /* sdr_callback: the same block to every client */
    for (int c = 0; c < n_clients; c++)
      dsp_worker_process(blocks + (size_t)b * server.buffer_size, server.buffer_size, workers[c]);
  for (int c = 0; c < n_clients; c++) dsp_worker_destroy(workers[c]); /* drain + join */
#endif
  clock_gettime(CLOCK_MONOTONIC, &t1);
  const double dt = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  const double samples = (double)n_blocks * server.buffer_size / (server.sdr_type == SDR_TYPE_AIRSPY ? 4 : 2);
  fflush(stdout);
  fprintf(stdout, "{\"harness\": \"reference dsp_worker.c + queue.c\", \"simd_status\": \"%s\", \"clients\": %d, "
                  "\"blocks\": %d, \"queue_size\": %d, \"seconds\": %.4f, \"input_msps\": %.3f}\n",
          SIMD_STATUS, n_clients, n_blocks, server.queue_size, dt, samples / dt / 1e6);
  free(configs);
  free(workers);
  free(blocks);
  return 0;
}
