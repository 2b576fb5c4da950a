/* TEST INFRASTRUCTURE ONLY — oracle/_ref/libperseus_refqueue.so.
 *
 * The reference's own delivery path, UNMODIFIED (/root/reference/perseus-in.c and perseuserr.c are
 * compiled from where they lie, see oracle/Makefile), running over a synthetic USB device instead of
 * libusb + hardware.  This is SURVEY.md §8(f) row n1: it lets the tests hand the product's
 * perseus_gpu_input_callback to the reference's real queue code (in-order check perseus-in.c:204,
 * resubmit :263, cancel/complete handshake :120-158) and compare the product's virtual receiver
 * (perseus_vrx_*) against that code's behaviour, fault cases included.
 *
 * The synthetic device: every submitted bulk-IN transfer joins a FIFO; fakeusb_pump() completes them
 * in submission order, filling transfer number n of the stream with bytes [n*len, (n+1)*len) of the
 * synthetic recording (same definition as oracle/perseus_oracle.c), optionally short (drop_every) or
 * with a neighbouring pair swapped (swap_every), then runs the transfer's callback — which is the
 * reference's static input_queue_callback.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <libusb-1.0/libusb.h>

#include "perseus-sdr.h"
#include "perseus-in.h"

void perseus_oracle_synth_random(uint8_t *dst, size_t nbytes, uint64_t seed, uint64_t byte_offset);   /* perseus_oracle.c */

struct libusb_device_handle {
	struct libusb_transfer *head, *tail;   /* submitted, not yet completed */
	uint64_t seed;
	uint64_t completed;                    /* data-carrying completions so far == stream position in transfers */
	uint32_t drop_every, swap_every;
};

struct libusb_transfer *libusb_alloc_transfer(int iso_packets)
{
	(void)iso_packets;
	return (struct libusb_transfer *)calloc(1, sizeof(struct libusb_transfer));
}

void libusb_free_transfer(struct libusb_transfer *t) { free(t); }

int libusb_submit_transfer(struct libusb_transfer *t)
{
	libusb_device_handle *d = t->dev_handle;
	if (t->fake_pending) return -6;        /* LIBUSB_ERROR_BUSY */
	t->fake_pending = 1;
	t->fake_cancel = 0;
	t->fake_next = NULL;
	if (d->tail) d->tail->fake_next = t; else d->head = t;
	d->tail = t;
	return 0;
}

int libusb_cancel_transfer(struct libusb_transfer *t)
{
	if (!t->fake_pending) return -5;       /* LIBUSB_ERROR_NOT_FOUND */
	t->fake_cancel = 1;
	return 0;
}

static struct libusb_transfer *pop(libusb_device_handle *d)
{
	struct libusb_transfer *t = d->head;
	if (!t) return NULL;
	d->head = t->fake_next;
	if (!d->head) d->tail = NULL;
	t->fake_next = NULL;
	t->fake_pending = 0;
	return t;
}

static void complete(libusb_device_handle *d, struct libusb_transfer *t, uint64_t stream_index)
{
	if (t->fake_cancel) {
		t->status = LIBUSB_TRANSFER_CANCELLED;
		t->actual_length = 0;
	} else {
		perseus_oracle_synth_random(t->buffer, (size_t)t->length, d->seed, stream_index * (uint64_t)t->length);
		t->status = LIBUSB_TRANSFER_COMPLETED;
		t->actual_length = t->length;
		if (d->drop_every && (stream_index + 1) % d->drop_every == 0) t->actual_length = t->length - 6;
	}
	t->callback(t);
}

/* ---- exported to the tests --------------------------------------------------------------- */

libusb_device_handle *fakeusb_open(uint64_t seed, uint32_t drop_every, uint32_t swap_every)
{
	libusb_device_handle *d = (libusb_device_handle *)calloc(1, sizeof(*d));
	if (d) { d->seed = seed; d->drop_every = drop_every; d->swap_every = swap_every; }
	return d;
}

void fakeusb_close(libusb_device_handle *d) { free(d); }

/* Completes up to `n` pending transfers; returns how many completed. */
uint64_t fakeusb_pump(libusb_device_handle *d, uint64_t n)
{
	uint64_t done = 0;
	while (done < n && d->head) {
		struct libusb_transfer *a = pop(d);
		const int swap = d->swap_every && !a->fake_cancel && (d->completed + 1) % d->swap_every == 0 && d->head &&
		                 !d->head->fake_cancel && done + 1 < n;
		if (swap) {   /* the NEXT slot's transfer completes first; each keeps the data of its own stream position */
			struct libusb_transfer *b = pop(d);
			const uint64_t ia = d->completed, ib = d->completed + 1;
			d->completed += 2;
			complete(d, b, ib);
			complete(d, a, ia);
			done += 2;
			continue;
		}
		const uint64_t idx = d->completed;
		if (!a->fake_cancel) d->completed++;
		complete(d, a, idx);
		done++;
	}
	return done;
}

/* The reference queue, driven exactly as perseus_start_async_input / perseus_stop_async_input drive it
 * (perseus-sdr.c:683 and :708-726). */
perseus_input_queue *refq_start(libusb_device_handle *d, int transfer_buf_size, perseus_input_callback cb, void *extra)
{
	perseus_input_queue *q = (perseus_input_queue *)calloc(1, sizeof(*q));
	if (!q) return NULL;
	if (perseus_input_queue_create(q, 8, d, transfer_buf_size, cb, extra) < 0) { free(q); return NULL; }
	return q;
}

uint64_t refq_stop(libusb_device_handle *d, perseus_input_queue *q)
{
	perseus_input_queue_cancel(q);
	while (perseus_input_queue_completed(q) == FALSE)
		if (fakeusb_pump(d, 8) == 0) break;
	const uint64_t bytes = q->bytes_received;
	perseus_input_queue_free(q);
	free(q);
	return bytes;
}

uint64_t refq_bytes_received(const perseus_input_queue *q) { return q->bytes_received; }
const void *refq_ring(const perseus_input_queue *q) { return q->buf; }
int refq_idx_expected(const perseus_input_queue *q) { return q->idx_expected; }
