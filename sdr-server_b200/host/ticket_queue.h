/*
 * host/ticket_queue.h -- per-client bounded queue between the ingest thread and a
 * client's dsp thread.
 *
 * Same contract as the reference's block queue (src/queue.h:8-15, src/queue.c):
 * FIFO; a fixed number of slots; when every slot is taken the NEWEST queued entry is
 * overwritten and "<3>queue is full" is logged (:90-94); the entry being processed
 * is detached and cannot be overwritten (:150-158); after interrupt the consumer
 * still drains what is queued and only then gets the poison pill (:136-148,
 * test/test_queue.c:42-59).  The difference is the payload: the reference memcpy's
 * the whole 256 KiB SDR block into every client's queue (:114); here the block was
 * submitted once to the GPU and the queue carries the 8-byte ticket.
 */
#ifndef XL_TICKET_QUEUE_H_
#define XL_TICKET_QUEUE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xl_ticket_queue xl_ticket_queue;

#define XL_TICKET_POISON (-1)

int xl_tq_create(int queue_size, xl_ticket_queue **queue);
void xl_tq_put(xl_ticket_queue *queue, int64_t ticket);
/* blocks until an entry is available; XL_TICKET_POISON after xl_tq_interrupt once drained */
int64_t xl_tq_take(xl_ticket_queue *queue);
/* the entry returned by the last xl_tq_take has been processed; its slot is free again */
void xl_tq_complete(xl_ticket_queue *queue);
void xl_tq_interrupt(xl_ticket_queue *queue);
void xl_tq_destroy(xl_ticket_queue *queue);
/* entries overwritten because the consumer was too slow (for tests / metrics) */
uint64_t xl_tq_overruns(xl_ticket_queue *queue);

#ifdef __cplusplus
}
#endif
#endif
