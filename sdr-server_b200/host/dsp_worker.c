/* host/dsp_worker.c -- see dsp_worker.h */
#include "dsp_worker.h"

#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

#include "lpf.h"
#include "ticket_queue.h"

struct xl_dsp_worker {
  xl_client_config config;
  xlg_group *group;
  int group_client;
  xl_ticket_queue *queue;
  pthread_t thread;
  int thread_started;
  FILE *file;
  volatile uint64_t written;
  volatile uint64_t lost;
};

/* src/dsp_worker.c:28-39: write everything or fail */
static int write_all(int fd, const void *data, size_t bytes) {
  const char *p = (const char *)data;
  while (bytes > 0) {
    ssize_t n = write(fd, p, bytes);
    if (n < 0) {
      if (errno == EINTR) {
        continue;
      }
      return -1;
    }
    p += n;
    bytes -= (size_t)n;
  }
  return 0;
}

static void *worker_main(void *arg) {
  xl_dsp_worker *w = (xl_dsp_worker *)arg;
  fprintf(stdout, "[%u] dsp_worker started\n", w->config.id);
  for (;;) {
    const int64_t ticket = xl_tq_take(w->queue);
    if (ticket == XL_TICKET_POISON) {
      break;
    }
    const void *out = NULL;
    size_t out_len = 0;
    int code = xlg_wait(w->group, ticket);
    if (code == 0) {
      code = xlg_output(w->group, ticket, w->group_client, &out, &out_len);
    }
    if (code != 0) {
      /* -ESTALE: this client fell more than host_ring blocks behind; the block is gone,
       * like a block overwritten in the reference's queue */
      fprintf(stderr, "<3>[%u] block %lld lost (%d)\n", w->config.id, (long long)ticket, code);
      w->lost++;
      xl_tq_complete(w->queue);
      continue;
    }
    const size_t bytes = out_len * 2 * sizeof(float); /* interleaved cf32, src/dsp_worker.c:13 */
    if (w->config.destination == XL_DESTINATION_FILE) {
      code = (w->file != NULL && fwrite(out, 1, bytes, w->file) == bytes) ? 0 : -1;
    } else if (w->config.destination == XL_DESTINATION_SOCKET) {
      code = write_all(w->config.client_socket, out, bytes);
    } else {
      fprintf(stderr, "<3>unknown destination: %d\n", w->config.destination);
      code = -1;
    }
    xl_tq_complete(w->queue);
    if (code != 0) {
      close(w->config.client_socket); /* src/dsp_worker.c:83-85: the tcp thread tears the client down */
    } else {
      w->written++;
    }
  }
  return NULL;
}

int xl_dsp_worker_start(const xl_client_config *config, xlg_group *group, uint32_t band_sampling_rate,
                        int lpf_cutoff_rate, int queue_size, const char *base_path, xl_dsp_worker **worker) {
  if (config == NULL || group == NULL || worker == NULL || config->sampling_rate == 0 || lpf_cutoff_rate <= 0) {
    return -EINVAL;
  }
  xl_dsp_worker *w = (xl_dsp_worker *)calloc(1, sizeof(*w));
  if (w == NULL) {
    return -ENOMEM;
  }
  w->config = *config;
  w->group = group;
  w->group_client = -1;

  float *taps = NULL;
  size_t len = 0;
  int code = create_low_pass_filter(1.0F, band_sampling_rate, config->sampling_rate / 2,
                                    config->sampling_rate / (uint32_t)lpf_cutoff_rate, &taps, &len);
  if (code != 0) {
    xl_dsp_worker_destroy(w);
    return code;
  }
  code = xlg_add_client(group, band_sampling_rate / config->sampling_rate, taps, len,
                        (int32_t)((int64_t)config->center_freq - (int64_t)config->band_freq), &w->group_client);
  free(taps); /* xlg_add_client copies (the per-filter create adopts, src/xlating.c:507-508) */
  if (code != 0) {
    xl_dsp_worker_destroy(w);
    return code;
  }
  if (config->destination == XL_DESTINATION_FILE) {
    char path[4096];
    snprintf(path, sizeof(path), "%s/%u.cf32", base_path != NULL ? base_path : ".", config->id);
    w->file = fopen(path, "wb");
    if (w->file == NULL) {
      fprintf(stderr, "<3>unable to open file for output: %s\n", path);
      xl_dsp_worker_destroy(w);
      return -1;
    }
  }
  code = xl_tq_create(queue_size, &w->queue);
  if (code != 0) {
    xl_dsp_worker_destroy(w);
    return code;
  }
  if (pthread_create(&w->thread, NULL, worker_main, w) != 0) {
    xl_dsp_worker_destroy(w);
    return -1;
  }
  w->thread_started = 1;
  *worker = w;
  return 0;
}

void xl_dsp_worker_post(xl_dsp_worker *w, int64_t ticket) { xl_tq_put(w->queue, ticket); }

void xl_dsp_worker_destroy(xl_dsp_worker *w) {
  if (w == NULL) {
    return;
  }
  fprintf(stdout, "[%u] dsp_worker is stopping\n", w->config.id);
  if (w->queue != NULL) {
    xl_tq_interrupt(w->queue);
  }
  if (w->thread_started) {
    pthread_join(w->thread, NULL);
  }
  if (w->queue != NULL) {
    xl_tq_destroy(w->queue);
  }
  if (w->file != NULL) {
    fclose(w->file);
  }
  if (w->group_client >= 0) {
    xlg_remove_client(w->group, w->group_client);
  }
  fprintf(stdout, "[%u] dsp_worker stopped\n", w->config.id);
  free(w);
}

uint64_t xl_dsp_worker_blocks_written(xl_dsp_worker *w) { return w->written; }
uint64_t xl_dsp_worker_blocks_lost(xl_dsp_worker *w) { return w->lost; }
