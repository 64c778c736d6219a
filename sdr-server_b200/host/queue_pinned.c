/*
 * host/queue_pinned.c -- the reference's block queue (src/queue.h:8-15) with its blocks
 * in PAGE-LOCKED host memory, so that the SDR block a dsp thread hands to process_* (or
 * an ingest thread hands to xlg_submit) can be DMA'd to the GPU without a staging copy
 * (SURVEY.md section 8f-2).
 *
 * Drop-in for src/queue.c: same six functions, same semantics --
 *   FIFO of `queue_size` pre-allocated blocks of `buffer_size` bytes (src/queue.c:42-85);
 *   queue_put copies the block in; when no block is free it overwrites the NEWEST queued
 *   block and logs "<3>queue is full" (:87-119);
 *   take_buffer_for_processing blocks until a block is queued, detaches it (the producer
 *   can neither read nor overwrite it while the consumer works on it, :150-158) and hands
 *   out its buffer; after interrupt_waiting_the_data the consumer still drains what is
 *   queued and only then receives buffer == NULL (:136-148, test/test_queue.c:23-59);
 *   complete_buffer_processing returns the detached block to the free pool (:167-177).
 * The reference's own test/test_queue.c is compiled unmodified against this file by
 * oracle/Makefile (_ref/test_queue_pinned / _ref/test_queue_pageable).
 *
 * Different inside: one slab of pinned memory and index rings instead of three linked
 * lists of malloc'd nodes.  -DXL_QUEUE_PAGEABLE builds the same logic on malloc (the CPU
 * check of the semantics); the default build fails loudly (create_queue -> -ENOMEM,
 * "<3>..." from xlg_alloc_pinned) when no CUDA device is usable.
 */
#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "xlating_group.h"

typedef struct queue_t queue;

struct queue_t {
  uint8_t *slab;        /* queue_size blocks of stride bytes, page-locked */
  size_t stride;
  size_t *lens;         /* bytes held by each block */
  int capacity;
  int *fifo;            /* ring of queued block indices, oldest at head */
  int head, count;
  int *free_stack;      /* indices of free blocks */
  int n_free;
  int detached;         /* block being processed, or -1 */
  int poison_pill;
  pthread_mutex_t mutex;
  pthread_cond_t condition;
};

static void *block_alloc(size_t bytes) {
#ifdef XL_QUEUE_PAGEABLE
  return malloc(bytes);
#else
  return xlg_alloc_pinned(bytes);
#endif
}

static void block_free(void *p) {
#ifdef XL_QUEUE_PAGEABLE
  free(p);
#else
  xlg_free_pinned(p);
#endif
}

int create_queue(uint32_t buffer_size, int queue_size, queue **result) {
  if (queue_size <= 0 || result == NULL) {
    return -EINVAL;
  }
  queue *q = (queue *)calloc(1, sizeof(*q));
  if (q == NULL) {
    return -ENOMEM;
  }
  q->stride = ((size_t)buffer_size + 255u) & ~(size_t)255u; /* blocks start on 256-byte boundaries */
  if (q->stride == 0) {
    q->stride = 256;
  }
  q->capacity = queue_size;
  q->slab = (uint8_t *)block_alloc(q->stride * (size_t)queue_size);
  q->lens = (size_t *)calloc((size_t)queue_size, sizeof(size_t));
  q->fifo = (int *)calloc((size_t)queue_size, sizeof(int));
  q->free_stack = (int *)calloc((size_t)queue_size, sizeof(int));
  if (q->slab == NULL || q->lens == NULL || q->fifo == NULL || q->free_stack == NULL) {
    if (q->slab != NULL) {
      block_free(q->slab);
    }
    free(q->lens);
    free(q->fifo);
    free(q->free_stack);
    free(q);
    return -ENOMEM;
  }
  for (int i = 0; i < queue_size; i++) {
    q->free_stack[i] = queue_size - 1 - i; /* block 0 is handed out first */
  }
  q->n_free = queue_size;
  q->detached = -1;
  pthread_mutex_init(&q->mutex, NULL);
  pthread_cond_init(&q->condition, NULL);
  *result = q;
  return 0;
}

void queue_put(const uint8_t *buffer, const size_t len, queue *q) {
  pthread_mutex_lock(&q->mutex);
  int slot;
  if (q->n_free == 0) {
    fprintf(stderr, "<3>queue is full\n");
    if (q->count == 0) {
      /* the only block is being processed: nowhere to put it (the reference would
       * dereference a NULL last_filled_node here, src/queue.c:93) */
      pthread_mutex_unlock(&q->mutex);
      return;
    }
    slot = q->fifo[(q->head + q->count - 1) % q->capacity]; /* overwrite the newest */
  } else {
    slot = q->free_stack[--q->n_free];
    q->fifo[(q->head + q->count) % q->capacity] = slot;
    q->count++;
  }
  const size_t n = len <= q->stride ? len : q->stride;
  memcpy(q->slab + (size_t)slot * q->stride, buffer, n);
  q->lens[slot] = n;
  pthread_cond_broadcast(&q->condition);
  pthread_mutex_unlock(&q->mutex);
}

void take_buffer_for_processing(uint8_t **buffer, size_t *len, queue *q) {
  pthread_mutex_lock(&q->mutex);
  while (q->count == 0 && !q->poison_pill) {
    pthread_cond_wait(&q->condition, &q->mutex);
  }
  if (q->count == 0) { /* interrupted and drained */
    pthread_mutex_unlock(&q->mutex);
    *buffer = NULL;
    return;
  }
  const int slot = q->fifo[q->head];
  q->head = (q->head + 1) % q->capacity;
  q->count--;
  q->detached = slot;
  *buffer = q->slab + (size_t)slot * q->stride;
  *len = q->lens[slot];
  pthread_mutex_unlock(&q->mutex);
}

void complete_buffer_processing(queue *q) {
  pthread_mutex_lock(&q->mutex);
  if (q->detached >= 0) {
    q->free_stack[q->n_free++] = q->detached;
    q->detached = -1;
  }
  pthread_mutex_unlock(&q->mutex);
}

void interrupt_waiting_the_data(queue *q) {
  if (q == NULL) {
    return;
  }
  pthread_mutex_lock(&q->mutex);
  q->poison_pill = 1;
  pthread_cond_broadcast(&q->condition);
  pthread_mutex_unlock(&q->mutex);
}

void destroy_queue(queue *q) {
  if (q == NULL) {
    return;
  }
  pthread_mutex_destroy(&q->mutex);
  pthread_cond_destroy(&q->condition);
  block_free(q->slab);
  free(q->lens);
  free(q->fifo);
  free(q->free_stack);
  free(q);
}
