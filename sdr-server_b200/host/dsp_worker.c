/* host/dsp_worker.c -- see dsp_worker.h */
#include "dsp_worker.h"

#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

#include <zlib.h>

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
  gzFile gz;                /* use_gzip file destination (src/dsp_worker.c:126-133) */
  float *output;            /* this client's own copy of a block's result (see worker_main) */
  size_t output_cap;        /* complex samples */
  volatile uint64_t written;
  volatile uint64_t lost;
  volatile uint64_t failed; /* blocks whose write failed (full disk, closed socket) */
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
    /* Copy this client's result out of the group's result ring BEFORE the (possibly slow)
     * write: the ring entry is recycled host_ring submits later whatever this thread is
     * doing, and xlg_read_output re-validates the entry after the copy (-ESTALE instead of
     * a torn block).  The reference gets the same guarantee from its detached queue node
     * (src/queue.c:150-158). */
    const void *out = w->output;
    size_t out_len = 0;
    int code = xlg_read_output(w->group, ticket, w->group_client, w->output, w->output_cap, &out_len);
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
      /* src/dsp_worker.c:10-26: plain or gzip; a short write means the disk is full */
      if (w->file != NULL) {
        code = fwrite(out, 1, bytes, w->file) == bytes ? 0 : -1;
      } else if (w->gz != NULL) {
        code = (bytes == 0 || gzwrite(w->gz, out, (unsigned)bytes) == (int)bytes) ? 0 : -1;
      } else {
        fprintf(stderr, "<3>unknown file output\n");
        code = -1;
      }
    } else if (w->config.destination == XL_DESTINATION_SOCKET) {
      code = write_all(w->config.client_socket, out, bytes);
    } else {
      fprintf(stderr, "<3>unknown destination: %d\n", w->config.destination);
      code = -1;
    }
    xl_tq_complete(w->queue);
    if (code != 0) {
      close(w->config.client_socket); /* src/dsp_worker.c:83-85: the tcp thread tears the client down */
      w->failed++;
    } else {
      w->written++;
    }
  }
  return NULL;
}

int xl_dsp_worker_start(const xl_client_config *config, xlg_group *group, uint32_t band_sampling_rate,
                        uint32_t max_block_elements, int lpf_cutoff_rate, int queue_size, const char *base_path,
                        int use_gzip, xl_dsp_worker **worker) {
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
  w->output_cap = (size_t)max_block_elements / 2 / (band_sampling_rate / config->sampling_rate) + 2;
  w->output = (float *)malloc(w->output_cap * 2 * sizeof(float));
  if (w->output == NULL) {
    xl_dsp_worker_destroy(w);
    return -ENOMEM;
  }
  if (config->destination == XL_DESTINATION_FILE) {
    char path[4096];
    if (use_gzip) { /* src/dsp_worker.c:126-133 */
      snprintf(path, sizeof(path), "%s/%u.cf32.gz", base_path != NULL ? base_path : ".", config->id);
      w->gz = gzopen(path, "wb");
      if (w->gz == NULL) {
        fprintf(stderr, "<3>unable to open gz file for output: %s\n", path);
        xl_dsp_worker_destroy(w);
        return -1;
      }
    } else {
      snprintf(path, sizeof(path), "%s/%u.cf32", base_path != NULL ? base_path : ".", config->id);
      w->file = fopen(path, "wb");
      if (w->file == NULL) {
        fprintf(stderr, "<3>unable to open file for output: %s\n", path);
        xl_dsp_worker_destroy(w);
        return -1;
      }
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
  if (w->gz != NULL) {
    gzclose(w->gz);
  }
  free(w->output);
  if (w->group_client >= 0) {
    xlg_remove_client(w->group, w->group_client);
  }
  fprintf(stdout, "[%u] dsp_worker stopped\n", w->config.id);
  free(w);
}

uint64_t xl_dsp_worker_blocks_written(xl_dsp_worker *w) { return w->written; }
uint64_t xl_dsp_worker_blocks_lost(xl_dsp_worker *w) { return w->lost; }
uint64_t xl_dsp_worker_blocks_failed(xl_dsp_worker *w) { return w->failed; }
uint64_t xl_dsp_worker_queue_overruns(xl_dsp_worker *w) { return w->queue != NULL ? xl_tq_overruns(w->queue) : 0; }
