/* host/ticket_queue.c -- see ticket_queue.h.  A ring of tickets guarded by one
 * mutex + condition variable; `detached` accounts for the entry that is being
 * processed so that capacity matches the reference's node pool (free + filled +
 * detached == queue_size, src/queue.c:15-28). */
#include "ticket_queue.h"

#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

struct xl_ticket_queue {
  int64_t *ring;
  int capacity;
  int head;      /* index of the oldest queued entry */
  int count;     /* queued entries */
  int detached;  /* 1 while the consumer holds an entry */
  int poisoned;
  uint64_t overruns;
  pthread_mutex_t mutex;
  pthread_cond_t nonempty;
};

int xl_tq_create(int queue_size, xl_ticket_queue **queue) {
  if (queue_size <= 0 || queue == NULL) {
    return -EINVAL;
  }
  xl_ticket_queue *q = (xl_ticket_queue *)calloc(1, sizeof(*q));
  if (q == NULL) {
    return -ENOMEM;
  }
  q->ring = (int64_t *)malloc(sizeof(int64_t) * (size_t)queue_size);
  if (q->ring == NULL) {
    free(q);
    return -ENOMEM;
  }
  q->capacity = queue_size;
  pthread_mutex_init(&q->mutex, NULL);
  pthread_cond_init(&q->nonempty, NULL);
  *queue = q;
  return 0;
}

void xl_tq_put(xl_ticket_queue *q, int64_t ticket) {
  pthread_mutex_lock(&q->mutex);
  if (q->count + q->detached >= q->capacity) {
    /* every slot is queued or being processed: overwrite the newest queued entry */
    fprintf(stderr, "<3>queue is full\n");
    q->overruns++;
    if (q->count > 0) {
      q->ring[(q->head + q->count - 1) % q->capacity] = ticket;
    }
    /* (a queue whose only slot is detached has nowhere to put it: the entry is dropped) */
  } else {
    q->ring[(q->head + q->count) % q->capacity] = ticket;
    q->count++;
  }
  pthread_cond_broadcast(&q->nonempty);
  pthread_mutex_unlock(&q->mutex);
}

int64_t xl_tq_take(xl_ticket_queue *q) {
  pthread_mutex_lock(&q->mutex);
  while (q->count == 0 && !q->poisoned) {
    pthread_cond_wait(&q->nonempty, &q->mutex);
  }
  if (q->count == 0) { /* poisoned and drained */
    pthread_mutex_unlock(&q->mutex);
    return XL_TICKET_POISON;
  }
  const int64_t ticket = q->ring[q->head];
  q->head = (q->head + 1) % q->capacity;
  q->count--;
  q->detached = 1;
  pthread_mutex_unlock(&q->mutex);
  return ticket;
}

void xl_tq_complete(xl_ticket_queue *q) {
  pthread_mutex_lock(&q->mutex);
  q->detached = 0;
  pthread_mutex_unlock(&q->mutex);
}

void xl_tq_interrupt(xl_ticket_queue *q) {
  if (q == NULL) {
    return;
  }
  pthread_mutex_lock(&q->mutex);
  q->poisoned = 1;
  pthread_cond_broadcast(&q->nonempty);
  pthread_mutex_unlock(&q->mutex);
}

uint64_t xl_tq_overruns(xl_ticket_queue *q) {
  pthread_mutex_lock(&q->mutex);
  const uint64_t n = q->overruns;
  pthread_mutex_unlock(&q->mutex);
  return n;
}

void xl_tq_destroy(xl_ticket_queue *q) {
  if (q == NULL) {
    return;
  }
  pthread_mutex_destroy(&q->mutex);
  pthread_cond_destroy(&q->nonempty);
  free(q->ring);
  free(q);
}
