/* TEST INFRASTRUCTURE ONLY — part of oracle/_ref/libperseus_sdr_ref.so.
 *
 * A synthetic Perseus receiver behind the libusb-1.0 API, so that the reference's OWN library
 * (/root/reference/perseus-sdr.c, perseusfx2.c, perseus-in.c, perseuserr.c -- compiled UNMODIFIED from
 * where they lie, see oracle/Makefile) runs without hardware.  SURVEY.md §8(f) row n1 as written:
 *   enumeration     perseus_init's device scan                           perseus-sdr.c:126-158
 *   open            configuration / interface / clear_halt probes        perseus-sdr.c:254-290
 *   FX2 firmware    vendor request 0xA0 RAM writes + CPUCS reset         perseusfx2.c:81-123,164-202
 *   EP1 commands    FPGARESET / FPGACONFIG / FPGACHECK / FPGASIO / FX2PORTE / EEPROMREAD / SHUTDOWN
 *                   with their EP 0x81 status replies                    perseusfx2.c:125-359
 *   EP 0x82         bulk-IN transfers completed in submission order      perseus-in.c:83-96,187-264
 *   event loop      libusb_handle_events_timeout, called by the reference's SCHED_FIFO poll thread
 *                                                                         perseus-sdr.c:736-774
 * What it lets the tests do: hand the product's perseus_gpu_input_callback to the reference's real
 * perseus_start_async_input and have it called on the reference's own poll thread; compare the product's
 * virtual receiver (perseus_vrx_*) with the reference's queue code under every transfer status; run the
 * reference's static getFpgaFile through perseus_set_sampling_rate and see which bitstream arrives.
 *
 * The device: once the FPGA is configured (the bytes received through FPGACONFIG are recognised by length and
 * FNV-1a hash against the table oracle/gen_fpga_data.py writes next to fpgaImgTbl) and FIFOEN is set
 * (perseus-sdr.c:687-688), transfer number n of the stream is filled with bytes [n*len, (n+1)*len) of the synthetic
 * recording (same definition as oracle/perseus_oracle.c), optionally short / swapped / timed out / failed, optionally
 * paced at the bitstream's sample rate; after `limit` transfers the stream stalls and pending transfers time out
 * like real ones.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <libusb-1.0/libusb.h>

#include "perseus-sdr.h"
#include "perseus-in.h"
#include "fakeusb.h"

void perseus_oracle_synth_random(uint8_t *dst, size_t nbytes, uint64_t seed, uint64_t byte_offset);   /* perseus_oracle.c */

/* written by oracle/gen_fpga_data.py into oracle/_ref/fpga_data.c, next to the reference's fpgaImgTbl */
extern const fakeusb_bitstream_id fakeusb_bitstream_ids[];
extern const int fakeusb_n_bitstream_ids;

struct libusb_device_handle {
	pthread_mutex_t mu;
	struct libusb_transfer *head, *tail;   /* submitted bulk-IN transfers, oldest first */
	fakeusb_config cfg;
	int manual;                            /* made by fakeusb_open(): no FX2 in front, pumped by hand */
	uint64_t stream_t0_ns;
	unsigned char reply[64];               /* next EP 0x81 status packet */
	int reply_len;
	fakeusb_state st;
};

struct libusb_device {
	libusb_device_handle *rx;
	uint8_t bus, addr;
};

struct libusb_context { int unused; };

static struct libusb_context g_ctx;
static libusb_device_handle g_rx = { .mu = PTHREAD_MUTEX_INITIALIZER };
static libusb_device g_dev = { &g_rx, 3, 7 };
static int g_present;

static uint64_t now_ns(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

static void sleep_ns(uint64_t ns)
{
	struct timespec ts = { (time_t)(ns / 1000000000ull), (long)(ns % 1000000000ull) };
	nanosleep(&ts, NULL);
}

static uint64_t fnv1a(uint64_t h, const unsigned char *p, size_t n)
{
	for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001B3ull; }
	return h;
}
#define FNV_INIT 0xCBF29CE484222325ull

/* ================================================================= asynchronous bulk-IN (EP 0x82) */

struct libusb_transfer *libusb_alloc_transfer(int iso_packets)
{
	(void)iso_packets;
	return (struct libusb_transfer *)calloc(1, sizeof(struct libusb_transfer));
}

void libusb_free_transfer(struct libusb_transfer *t) { free(t); }

int libusb_submit_transfer(struct libusb_transfer *t)
{
	libusb_device_handle *d = t->dev_handle;
	pthread_mutex_lock(&d->mu);
	if (t->fake_pending) { pthread_mutex_unlock(&d->mu); return LIBUSB_ERROR_BUSY; }
	t->fake_pending = 1;
	t->fake_cancel = 0;
	t->fake_next = NULL;
	t->fake_submit_ns = now_ns();
	if (d->tail) d->tail->fake_next = t; else d->head = t;
	d->tail = t;
	d->st.submits++;
	pthread_mutex_unlock(&d->mu);
	return 0;
}

int libusb_cancel_transfer(struct libusb_transfer *t)
{
	libusb_device_handle *d = t->dev_handle;
	pthread_mutex_lock(&d->mu);
	const int pending = t->fake_pending;
	if (pending) t->fake_cancel = 1;       /* reported asynchronously, by the event loop, as in libusb */
	pthread_mutex_unlock(&d->mu);
	return pending ? 0 : LIBUSB_ERROR_NOT_FOUND;
}

/* caller holds d->mu */
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

/* Fills in what the device did with transfer number `stream_index` of the stream.  Caller holds d->mu. */
static void decide(libusb_device_handle *d, struct libusb_transfer *t, uint64_t stream_index)
{
	const uint64_t seq = stream_index + 1;
	const fakeusb_config *c = &d->cfg;
	t->actual_length = 0;
	if (t->fake_cancel) {
		t->status = LIBUSB_TRANSFER_CANCELLED;
		d->st.cancelled++;
	} else if (c->fail_at && seq == c->fail_at) {
		t->status = (enum libusb_transfer_status)c->fail_status;
		d->st.failed++;
	} else if (c->timeout_every && seq % c->timeout_every == 0) {
		t->status = LIBUSB_TRANSFER_TIMED_OUT;
		d->st.timed_out++;
	} else {
		perseus_oracle_synth_random(t->buffer, (size_t)t->length, c->seed, stream_index * (uint64_t)t->length);
		t->status = LIBUSB_TRANSFER_COMPLETED;
		t->actual_length = t->length;
		if (c->drop_every && seq % c->drop_every == 0) t->actual_length = t->length - 6;
		d->st.completed_ok++;
	}
}

/* Completes the oldest pending transfer -- or, on a swap position, the two oldest in the wrong order.  Called with
 * d->mu held; returns with it released (the transfer callbacks resubmit, which takes the lock).  Returns the number
 * of transfers completed. */
static int complete_next(libusb_device_handle *d, int allow_pair)
{
	struct libusb_transfer *a = pop(d), *b = NULL;
	if (!a) { pthread_mutex_unlock(&d->mu); return 0; }
	if (a->fake_cancel) {
		decide(d, a, 0);
		pthread_mutex_unlock(&d->mu);
		a->callback(a);
		return 1;
	}
	const fakeusb_config *c = &d->cfg;
	const int swap = allow_pair && c->swap_every && (d->st.stream_pos + 1) % c->swap_every == 0 && d->head && !d->head->fake_cancel;
	if (swap) {   /* the NEXT slot's transfer completes first; each keeps the data of its own stream position */
		b = pop(d);
		const uint64_t ia = d->st.stream_pos, ib = d->st.stream_pos + 1;
		d->st.stream_pos += 2;
		decide(d, b, ib);
		decide(d, a, ia);
		pthread_mutex_unlock(&d->mu);
		b->callback(b);
		a->callback(a);
		return 2;
	}
	decide(d, a, d->st.stream_pos++);
	pthread_mutex_unlock(&d->mu);
	a->callback(a);
	return 1;
}

static void note_events_thread(libusb_device_handle *d)
{
	if (d->st.events_calls++ == 0) {
		struct sched_param sp;
		int pol = -1;
		if (pthread_getschedparam(pthread_self(), &pol, &sp) == 0) {
			d->st.events_thread_policy = pol;
			d->st.events_thread_priority = sp.sched_priority;
		}
		d->st.events_thread_is_fifo = (pol == SCHED_FIFO);
	}
}

int libusb_handle_events_timeout(libusb_context *ctx, struct timeval *tv)
{
	(void)ctx;
	libusb_device_handle *d = &g_rx;
	const uint64_t deadline = now_ns() + (tv ? (uint64_t)tv->tv_sec * 1000000000ull + (uint64_t)tv->tv_usec * 1000ull : 0);
	for (;;) {
		uint64_t wait = 1000000ull;   /* idle poll: 1 ms */
		pthread_mutex_lock(&d->mu);
		note_events_thread(d);
		struct libusb_transfer *t = g_present ? d->head : NULL;
		const uint64_t now = now_ns();
		if (t) {
			const int streaming = d->st.fifo_enabled && d->st.fpga_rate && (!d->cfg.limit || d->st.stream_pos < d->cfg.limit);
			if (t->fake_cancel) return complete_next(d, 0), 0;
			if (streaming) {
				uint64_t due = 0;
				if (d->cfg.realtime)   /* the FPGA fills one transfer every length/6 samples at the bitstream's rate */
					due = d->stream_t0_ns + (uint64_t)((double)(d->st.stream_pos + 1) * (double)(t->length / 6) * 1e9 / (double)d->st.fpga_rate);
				if (now >= due) return complete_next(d, 1), 0;
				if (due - now < wait) wait = due - now;
			} else if (t->timeout && now >= t->fake_submit_ns + (uint64_t)t->timeout * 1000000ull) {
				/* nothing arrives (FIFO disabled, or the stream stalled after `limit`): the transfer times out like a real one */
				pop(d);
				t->status = LIBUSB_TRANSFER_TIMED_OUT;
				t->actual_length = 0;
				d->st.timed_out++;
				pthread_mutex_unlock(&d->mu);
				t->callback(t);
				return 0;
			}
		}
		pthread_mutex_unlock(&d->mu);
		if (now >= deadline) return 0;
		if (deadline - now < wait) wait = deadline - now;
		sleep_ns(wait);
	}
}

/* ================================================================= enumeration and handles */

int fakeusb_plug(const fakeusb_config *cfg);

/* The reference's own APPLICATION (examples/perseustest.c, built unmodified as oracle/_ref/perseustest_ref) has no place to
 * call fakeusb_plug(), so the receiver can also be plugged in from the environment: FAKEUSB_AUTOPLUG=1 and optionally
 * FAKEUSB_LIMIT, FAKEUSB_SEED, FAKEUSB_EP, FAKEUSB_REALTIME, FAKEUSB_DROP_EVERY, FAKEUSB_SWAP_EVERY, FAKEUSB_TIMEOUT_EVERY. */
static uint64_t env_u64(const char *name, uint64_t dflt)
{
	const char *v = getenv(name);
	return v && *v ? strtoull(v, NULL, 0) : dflt;
}

static void autoplug(void)
{
	if (g_present || !env_u64("FAKEUSB_AUTOPLUG", 0)) return;
	fakeusb_config c;
	memset(&c, 0, sizeof(c));
	c.struct_size = sizeof(c);
	c.seed = env_u64("FAKEUSB_SEED", 0x5045525345555300ull);
	c.limit = env_u64("FAKEUSB_LIMIT", 0);
	c.ep_max_packet = (uint32_t)env_u64("FAKEUSB_EP", 512);
	c.realtime = (uint32_t)env_u64("FAKEUSB_REALTIME", 0);
	c.drop_every = (uint32_t)env_u64("FAKEUSB_DROP_EVERY", 0);
	c.swap_every = (uint32_t)env_u64("FAKEUSB_SWAP_EVERY", 0);
	c.timeout_every = (uint32_t)env_u64("FAKEUSB_TIMEOUT_EVERY", 0);
	c.serial = (uint32_t)env_u64("FAKEUSB_SERIAL", 1234);
	fakeusb_plug(&c);
}

int libusb_init(libusb_context **ctx)
{
	if (ctx) *ctx = &g_ctx;
	autoplug();
	g_rx.st.inits++;
	return 0;
}

void libusb_exit(libusb_context *ctx) { (void)ctx; g_rx.st.exits++; }
void libusb_set_debug(libusb_context *ctx, int level) { (void)ctx; (void)level; }

ssize_t libusb_get_device_list(libusb_context *ctx, libusb_device ***list)
{
	(void)ctx;
	const int n = g_present ? 1 : 0;
	libusb_device **l = (libusb_device **)calloc((size_t)n + 1, sizeof(*l));
	if (!l) return LIBUSB_ERROR_NO_MEM;
	if (n) { l[0] = &g_dev; g_rx.st.device_refs++; }
	*list = l;
	return n;
}

void libusb_free_device_list(libusb_device **list, int unref_devices)
{
	if (!list) return;
	if (unref_devices)
		for (libusb_device **p = list; *p; ++p) (*p)->rx->st.device_refs--;
	free(list);
}

int libusb_get_device_descriptor(libusb_device *dev, struct libusb_device_descriptor *desc)
{
	memset(desc, 0, sizeof(*desc));
	desc->bLength = 18;
	desc->bDescriptorType = 1;
	desc->idVendor = 0x04B4;                                            /* PERSEUS_VID, perseusfx2.h:37 */
	desc->idProduct = dev->rx->st.firmware_loaded ? 0x325C : 0x8613;   /* PERSEUS_PID : PERSEUS_PID_BLANKEEPROM */
	return 0;
}

uint8_t libusb_get_bus_number(libusb_device *dev) { return dev->bus; }
uint8_t libusb_get_device_address(libusb_device *dev) { return dev->addr; }
libusb_device *libusb_ref_device(libusb_device *dev) { dev->rx->st.device_refs++; return dev; }
void libusb_unref_device(libusb_device *dev) { dev->rx->st.device_refs--; }

int libusb_get_max_packet_size(libusb_device *dev, unsigned char endpoint)
{
	if (endpoint != 0x82) return LIBUSB_ERROR_NOT_FOUND;
	return (int)dev->rx->cfg.ep_max_packet;
}

int libusb_open(libusb_device *dev, libusb_device_handle **handle)
{
	if (!g_present) return LIBUSB_ERROR_NO_DEVICE;
	dev->rx->st.opens++;
	*handle = dev->rx;
	return 0;
}

void libusb_close(libusb_device_handle *h) { h->st.closes++; }
int libusb_kernel_driver_active(libusb_device_handle *h, int i) { (void)h; (void)i; return 0; }
int libusb_detach_kernel_driver(libusb_device_handle *h, int i) { (void)h; (void)i; return 0; }
int libusb_set_configuration(libusb_device_handle *h, int c) { (void)h; return c == 1 ? 0 : LIBUSB_ERROR_INVALID_PARAM; }
int libusb_claim_interface(libusb_device_handle *h, int i) { h->st.claimed = 1; return i == 0 ? 0 : LIBUSB_ERROR_NOT_FOUND; }
int libusb_release_interface(libusb_device_handle *h, int i) { (void)i; h->st.claimed = 0; return 0; }
int libusb_set_interface_alt_setting(libusb_device_handle *h, int i, int a) { (void)h; (void)i; (void)a; return 0; }

/* An FX2 without firmware has only its default control endpoint: the three Perseus endpoints do not exist yet, which
 * is how perseus_open tells whether the firmware is there (perseus-sdr.c:276-285). */
int libusb_clear_halt(libusb_device_handle *h, unsigned char endpoint)
{
	(void)endpoint;
	return h->st.firmware_loaded ? 0 : LIBUSB_ERROR_NOT_FOUND;
}

const char *libusb_error_name(int e)
{
	switch (e) {
	case 0: return "LIBUSB_SUCCESS";
	case LIBUSB_ERROR_IO: return "LIBUSB_ERROR_IO";
	case LIBUSB_ERROR_NO_DEVICE: return "LIBUSB_ERROR_NO_DEVICE";
	case LIBUSB_ERROR_NOT_FOUND: return "LIBUSB_ERROR_NOT_FOUND";
	case LIBUSB_ERROR_TIMEOUT: return "LIBUSB_ERROR_TIMEOUT";
	case LIBUSB_ERROR_PIPE: return "LIBUSB_ERROR_PIPE";
	default: return "LIBUSB_ERROR_OTHER";
	}
}

/* ================================================================= FX2: firmware download and EP1 command protocol */

int libusb_control_transfer(libusb_device_handle *h, uint8_t request_type, uint8_t bRequest, uint16_t wValue, uint16_t wIndex,
                            unsigned char *data, uint16_t wLength, unsigned int timeout)
{
	(void)wIndex; (void)timeout;
	if (request_type != 0x40 || bRequest != 0xA0) return LIBUSB_ERROR_PIPE;   /* FX2_BM_VENDOR_REQUEST / FX2_REQUEST_FIRMWARE_LOAD */
	pthread_mutex_lock(&h->mu);
	if (wValue == 0xE600) {                                                   /* FX2_ADDR_CPUCS: 1 = hold the 8051 in reset, 0 = run */
		const int hold = wLength ? data[0] : 0;
		if (hold) {
			h->st.cpu_resets++;
			h->st.fw_records = h->st.fw_bytes = 0;
			h->st.fw_hash = FNV_INIT;
		} else if (h->st.fw_records) {
			h->st.firmware_loaded = 1;                                        /* the new firmware runs and re-enumerates as 04B4:325C */
		}
	} else {
		const unsigned char hdr[4] = { (unsigned char)wValue, (unsigned char)(wValue >> 8), (unsigned char)wLength, (unsigned char)(wLength >> 8) };
		h->st.fw_hash = fnv1a(fnv1a(h->st.fw_hash, hdr, 4), data, wLength);
		h->st.fw_records++;
		h->st.fw_bytes += wLength;
	}
	pthread_mutex_unlock(&h->mu);
	return wLength;
}

static void set_reply(libusb_device_handle *h, const void *p, int n)
{
	if (n > (int)sizeof(h->reply)) n = (int)sizeof(h->reply);
	memcpy(h->reply, p, (size_t)n);
	h->reply_len = n;
}

static void fpga_check(libusb_device_handle *h)
{
	h->st.fpga_rate = 0;
	for (int i = 0; i < fakeusb_n_bitstream_ids; i++)
		if (fakeusb_bitstream_ids[i].size == h->st.fpga_bytes && fakeusb_bitstream_ids[i].fnv1a64 == h->st.fpga_hash)
			h->st.fpga_rate = fakeusb_bitstream_ids[i].rate;
	const unsigned char r[2] = { 0x02, (unsigned char)(h->st.fpga_rate ? 1 : 0) };   /* DONE line, perseusfx2.c:347-352 */
	set_reply(h, r, 2);
}

static int ep1_command(libusb_device_handle *h, const unsigned char *p, int n)
{
	if (n < 1) return LIBUSB_ERROR_IO;
	h->st.commands++;
	switch (p[0]) {
	case 0x00:   /* PERSEUS_CMD_FPGACONFIG: up to 63 bitstream bytes, perseusfx2.c:318-332 */
		h->st.fpga_hash = fnv1a(h->st.fpga_hash, p + 1, (size_t)n - 1);
		h->st.fpga_bytes += (uint64_t)n - 1;
		break;
	case 0x01:   /* PERSEUS_CMD_FPGARESET */
		h->st.fpga_bytes = 0;
		h->st.fpga_hash = FNV_INIT;
		h->st.fpga_rate = 0;
		h->st.fpga_resets++;
		h->st.fifo_enabled = 0;
		break;
	case 0x02:   /* PERSEUS_CMD_FPGACHECK */
		fpga_check(h);
		break;
	case 0x03: { /* PERSEUS_CMD_FPGASIO: {ctl, freg} out, the same shape back on EP 0x81, perseusfx2.c:231-252 */
		if (n >= 6) {
			const int was = h->st.fifo_enabled;
			h->st.sio_ctl = p[1];
			memcpy(&h->st.sio_freg, p + 2, 4);
			h->st.sio_writes++;
			h->st.fifo_enabled = p[1] & 0x01;                                 /* PERSEUS_SIO_FIFOEN */
			if (!was && h->st.fifo_enabled) h->stream_t0_ns = now_ns();
		}
		set_reply(h, p, n);
		break;
	}
	case 0x04:   /* PERSEUS_CMD_FX2PORTE: attenuator + preselector bits */
		if (n >= 2) { h->st.porte = p[1]; h->st.porte_writes++; }
		break;
	case 0x06: { /* PERSEUS_CMD_EEPROMREAD {op, addr16, count}, reply {op, TRUE, data...}, perseusfx2.c:125-162 */
		unsigned char r[2 + 32] = { 0x06, 1 };
		const unsigned addr = n >= 3 ? (unsigned)(p[1] | (p[2] << 8)) : 0;
		unsigned count = n >= 4 ? p[3] : 0;
		if (count > 32) count = 32;
		unsigned char eeprom[64];
		memset(eeprom, 0xFF, sizeof(eeprom));
		eeprom_prodid id;
		memset(&id, 0, sizeof(id));
		id.sn = (uint16_t)h->cfg.serial;
		id.prodcode = h->cfg.preserie ? 0x0000 : 0x8014;                      /* PERSEUS_PRODCODE */
		id.hwrel = 2; id.hwver = 1;
		memcpy(id.signature, "PERSEU", 6);
		memcpy(eeprom + 8, &id, sizeof(id));                                  /* PERSEUS_EEPROMADR_PRODID */
		for (unsigned k = 0; k < count; k++) r[2 + k] = addr + k < sizeof(eeprom) ? eeprom[addr + k] : 0xFF;
		set_reply(h, r, 2 + (int)count);
		h->st.eeprom_reads++;
		break;
	}
	case 0x08:   /* PERSEUS_CMD_SHUTDOWN */
		h->st.shutdowns++;
		h->st.fifo_enabled = 0;
		break;
	default:
		return LIBUSB_ERROR_PIPE;
	}
	return 0;
}

int libusb_bulk_transfer(libusb_device_handle *h, unsigned char endpoint, unsigned char *data, int length, int *actual_length,
                         unsigned int timeout)
{
	(void)timeout;
	int rc = 0, done = 0;
	pthread_mutex_lock(&h->mu);
	if (!h->st.firmware_loaded) {
		rc = LIBUSB_ERROR_NOT_FOUND;
	} else if (endpoint == 0x01) {           /* PERSEUS_EP_CMD */
		rc = ep1_command(h, data, length);
		done = rc ? 0 : length;
	} else if (endpoint == 0x81) {           /* PERSEUS_EP_STATUS */
		if (!h->reply_len) rc = LIBUSB_ERROR_TIMEOUT;
		else {
			done = h->reply_len < length ? h->reply_len : length;
			memcpy(data, h->reply, (size_t)done);
			h->reply_len = 0;
		}
	} else {
		rc = LIBUSB_ERROR_NOT_FOUND;
	}
	pthread_mutex_unlock(&h->mu);
	if (actual_length) *actual_length = done;
	return rc;
}

/* ================================================================= exported to the tests */

int fakeusb_plug(const fakeusb_config *cfg)
{
	libusb_device_handle *d = &g_rx;
	pthread_mutex_lock(&d->mu);
	if (d->head) { pthread_mutex_unlock(&d->mu); return -1; }   /* transfers still pending: stop streaming first */
	memset(&d->cfg, 0, sizeof(d->cfg));
	if (cfg) memcpy(&d->cfg, cfg, cfg->struct_size < sizeof(d->cfg) ? cfg->struct_size : sizeof(d->cfg));
	if (!d->cfg.ep_max_packet) d->cfg.ep_max_packet = 512;
	memset(&d->st, 0, sizeof(d->st));
	d->st.firmware_loaded = d->cfg.blank_eeprom ? 0 : 1;
	d->st.fw_hash = d->st.fpga_hash = FNV_INIT;
	d->reply_len = 0;
	g_present = cfg ? 1 : 0;
	pthread_mutex_unlock(&d->mu);
	return 0;
}

void fakeusb_unplug(void) { g_present = 0; }

void fakeusb_get_state(fakeusb_state *out)
{
	libusb_device_handle *d = &g_rx;
	pthread_mutex_lock(&d->mu);
	*out = d->st;
	pthread_mutex_unlock(&d->mu);
}

/* ---- the hand-pumped variant (tests/test_refqueue_cpu.py): a bare EP 0x82 with no FX2 in front of it */

libusb_device_handle *fakeusb_open(uint64_t seed, uint32_t drop_every, uint32_t swap_every)
{
	libusb_device_handle *d = (libusb_device_handle *)calloc(1, sizeof(*d));
	if (!d) return NULL;
	pthread_mutex_init(&d->mu, NULL);
	d->manual = 1;
	d->cfg.seed = seed;
	d->cfg.drop_every = drop_every;
	d->cfg.swap_every = swap_every;
	return d;
}

void fakeusb_set_faults(libusb_device_handle *d, uint32_t timeout_every, uint32_t fail_at, uint32_t fail_status)
{
	d->cfg.timeout_every = timeout_every;
	d->cfg.fail_at = fail_at;
	d->cfg.fail_status = fail_status;
}

void fakeusb_close(libusb_device_handle *d)
{
	pthread_mutex_destroy(&d->mu);
	free(d);
}

/* Completes up to `n` pending transfers; returns how many completed. */
uint64_t fakeusb_pump(libusb_device_handle *d, uint64_t n)
{
	uint64_t done = 0;
	while (done < n) {
		pthread_mutex_lock(&d->mu);
		if (!d->head) { pthread_mutex_unlock(&d->mu); break; }
		done += (uint64_t)complete_next(d, done + 1 < n);
	}
	return done;
}

uint64_t fakeusb_pending(libusb_device_handle *d)
{
	uint64_t n = 0;
	pthread_mutex_lock(&d->mu);
	for (struct libusb_transfer *t = d->head; t; t = t->fake_next) n++;
	pthread_mutex_unlock(&d->mu);
	return n;
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
int refq_completed(const perseus_input_queue *q) { return q->completed; }

/* ---- accessors into the reference's descriptor, so the tests need not mirror its layout */
int  reflib_descr_firmware_downloaded(const perseus_descr *d) { return d->firmware_downloaded; }
int  reflib_descr_fpga_configured(const perseus_descr *d) { return d->fpga_configured; }
int  reflib_descr_is_preserie(const perseus_descr *d) { return d->is_preserie; }
const void *reflib_descr_ring(const perseus_descr *d) { return d->input_queue.buf; }
uint64_t reflib_descr_bytes_received(const perseus_descr *d) { return d->input_queue.bytes_received; }
int  reflib_descr_queue_active(const perseus_descr *d) { return d->input_queue.transfer_queue != NULL; }
