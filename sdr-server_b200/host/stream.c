/* host/stream.c -- see stream.h */
#include "stream.h"

#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

struct client_node {
  xl_dsp_worker *worker;
  uint32_t id;
  uint64_t posted;
  struct client_node *next;
};

struct xl_stream {
  xl_stream_config config;
  xlg_group *group;
  struct client_node *clients;
  pthread_mutex_t mutex; /* the reference holds server->mutex for the whole fan-out (:258-270) */
};

int xl_stream_create(const xl_stream_config *config, xl_stream **stream) {
  if (config == NULL || stream == NULL || config->band_sampling_rate == 0 || config->buffer_size < 4 ||
      config->queue_size <= 0) {
    return -EINVAL;
  }
  xl_stream *s = (xl_stream *)calloc(1, sizeof(*s));
  if (s == NULL) {
    return -ENOMEM;
  }
  s->config = *config;
  pthread_mutex_init(&s->mutex, NULL);
  /* max block in scalar elements: airspy delivers int16 (src/dsp_worker.c:65) */
  const uint32_t max_elems =
      config->sdr_type == XLG_FMT_CS16 ? config->buffer_size / (uint32_t)sizeof(int16_t) : config->buffer_size;
  /* results stay readable for queue_size blocks: the slack the reference's per-client
   * queues give a slow consumer */
  int code = xlg_create_ex(config->device, config->band_sampling_rate, max_elems, 0, (uint32_t)config->queue_size,
                           &s->group);
  if (code != 0) {
    pthread_mutex_destroy(&s->mutex);
    free(s);
    return code;
  }
  *stream = s;
  return 0;
}

int xl_stream_add_client(xl_stream *s, const xl_client_config *client) {
  if (s == NULL || client == NULL) {
    return -EINVAL;
  }
  /* src/tcp_server.c:101: the client rate must divide the band rate */
  if (client->sampling_rate == 0 || s->config.band_sampling_rate % client->sampling_rate != 0) {
    fprintf(stderr, "<3>[%u] sampling rate %u does not divide the band rate %u\n", client->id,
            client->sampling_rate, s->config.band_sampling_rate);
    return -EINVAL;
  }
  struct client_node *node = (struct client_node *)calloc(1, sizeof(*node));
  if (node == NULL) {
    return -ENOMEM;
  }
  pthread_mutex_lock(&s->mutex);
  const uint32_t max_elems = s->config.sdr_type == XLG_FMT_CS16 ? s->config.buffer_size / (uint32_t)sizeof(int16_t)
                                                                : s->config.buffer_size;
  int code = xl_dsp_worker_start(client, s->group, s->config.band_sampling_rate, max_elems, s->config.lpf_cutoff_rate,
                                 s->config.queue_size, s->config.base_path, s->config.use_gzip, &node->worker);
  if (code == 0) {
    node->id = client->id;
    node->next = s->clients;
    s->clients = node;
  }
  pthread_mutex_unlock(&s->mutex);
  if (code != 0) {
    free(node);
  }
  return code;
}

int xl_stream_remove_client(xl_stream *s, uint32_t client_id) {
  pthread_mutex_lock(&s->mutex);
  struct client_node **pp = &s->clients;
  while (*pp != NULL && (*pp)->id != client_id) {
    pp = &(*pp)->next;
  }
  struct client_node *node = *pp;
  if (node != NULL) {
    *pp = node->next;
  }
  if (node != NULL) {
    /* still under the mutex: xlg_remove_client must not race xlg_submit (one producer) */
    xl_dsp_worker_destroy(node->worker);
  }
  pthread_mutex_unlock(&s->mutex);
  if (node == NULL) {
    return -ENOENT;
  }
  free(node);
  return 0;
}

int xl_stream_push(xl_stream *s, const uint8_t *buf, uint32_t buf_len) {
  pthread_mutex_lock(&s->mutex);
  if (s->clients == NULL) {
    pthread_mutex_unlock(&s->mutex);
    return 0; /* nobody listens: the reference stops the SDR in that case (:245-250) */
  }
  const size_t elems = s->config.sdr_type == XLG_FMT_CS16 ? buf_len / sizeof(int16_t) : buf_len;
  const int64_t ticket = xlg_submit(s->group, s->config.sdr_type, buf, elems, 0);
  if (ticket < 0) {
    pthread_mutex_unlock(&s->mutex);
    fprintf(stderr, "<3>block dropped: submit failed (%lld)\n", (long long)ticket);
    return (int)ticket;
  }
  for (struct client_node *n = s->clients; n != NULL; n = n->next) {
    xl_dsp_worker_post(n->worker, ticket);
    n->posted++;
  }
  pthread_mutex_unlock(&s->mutex);
  return 0;
}

int xl_stream_flush(xl_stream *s) {
  for (int spins = 0; spins < 150000; spins++) { /* 30 s */
    int pending = 0;
    pthread_mutex_lock(&s->mutex);
    for (struct client_node *n = s->clients; n != NULL; n = n->next) {
      const uint64_t accounted = xl_dsp_worker_blocks_written(n->worker) + xl_dsp_worker_blocks_lost(n->worker) +
                                 xl_dsp_worker_blocks_failed(n->worker) + xl_dsp_worker_queue_overruns(n->worker);
      if (accounted < n->posted) {
        pending = 1;
      }
    }
    pthread_mutex_unlock(&s->mutex);
    if (!pending) {
      return 0;
    }
    usleep(200);
  }
  fprintf(stderr, "<3>xl_stream_flush: clients still busy after 30 s\n");
  return -ETIMEDOUT;
}

int xl_stream_client_count(xl_stream *s) {
  int n = 0;
  pthread_mutex_lock(&s->mutex);
  for (struct client_node *c = s->clients; c != NULL; c = c->next) {
    n++;
  }
  pthread_mutex_unlock(&s->mutex);
  return n;
}

void xl_stream_destroy(xl_stream *s) {
  if (s == NULL) {
    return;
  }
  pthread_mutex_lock(&s->mutex);
  struct client_node *n = s->clients;
  s->clients = NULL;
  while (n != NULL) {
    struct client_node *next = n->next;
    xl_dsp_worker_destroy(n->worker);
    free(n);
    n = next;
  }
  pthread_mutex_unlock(&s->mutex);
  xlg_destroy(s->group);
  pthread_mutex_destroy(&s->mutex);
  free(s);
}
