/* TEST INFRASTRUCTURE ONLY (oracle/): a fake libusb-1.0 asynchronous API, just large enough for
 * the reference's OWN transfer queue (/root/reference/perseus-in.c) to compile and run without
 * hardware.  The names and the public fields of struct libusb_transfer follow the libusb-1.0 API
 * (that is what perseus-in.c is written against); the implementation (oracle/fakeusb.c) is a
 * synthetic device that completes submitted bulk-IN transfers in submission order with generated
 * wire data.  Never included by the product. */
#ifndef PERSEUS_ORACLE_FAKE_LIBUSB_H
#define PERSEUS_ORACLE_FAKE_LIBUSB_H
#include <stdint.h>

#define LIBUSB_CALL

typedef struct libusb_context libusb_context;
typedef struct libusb_device libusb_device;
typedef struct libusb_device_handle libusb_device_handle;   /* here: the synthetic device */

enum libusb_transfer_status {
	LIBUSB_TRANSFER_COMPLETED,
	LIBUSB_TRANSFER_ERROR,
	LIBUSB_TRANSFER_TIMED_OUT,
	LIBUSB_TRANSFER_CANCELLED,
	LIBUSB_TRANSFER_STALL,
	LIBUSB_TRANSFER_NO_DEVICE,
	LIBUSB_TRANSFER_OVERFLOW
};

enum libusb_transfer_type { LIBUSB_TRANSFER_TYPE_CONTROL = 0, LIBUSB_TRANSFER_TYPE_ISOCHRONOUS = 1,
                            LIBUSB_TRANSFER_TYPE_BULK = 2, LIBUSB_TRANSFER_TYPE_INTERRUPT = 3 };

struct libusb_transfer;
typedef void (LIBUSB_CALL *libusb_transfer_cb_fn)(struct libusb_transfer *transfer);

struct libusb_transfer {
	libusb_device_handle *dev_handle;
	uint8_t flags;
	unsigned char endpoint;
	unsigned char type;
	unsigned int timeout;
	enum libusb_transfer_status status;
	int length;
	int actual_length;
	libusb_transfer_cb_fn callback;
	void *user_data;
	unsigned char *buffer;
	int num_iso_packets;
	/* fake-device bookkeeping */
	struct libusb_transfer *fake_next;
	int fake_pending, fake_cancel;
};

struct libusb_transfer *libusb_alloc_transfer(int iso_packets);
void libusb_free_transfer(struct libusb_transfer *transfer);
int libusb_submit_transfer(struct libusb_transfer *transfer);
int libusb_cancel_transfer(struct libusb_transfer *transfer);

static inline void libusb_fill_bulk_transfer(struct libusb_transfer *transfer, libusb_device_handle *dev_handle,
                                             unsigned char endpoint, unsigned char *buffer, int length,
                                             libusb_transfer_cb_fn callback, void *user_data, unsigned int timeout)
{
	transfer->dev_handle = dev_handle;
	transfer->endpoint = endpoint;
	transfer->type = LIBUSB_TRANSFER_TYPE_BULK;
	transfer->timeout = timeout;
	transfer->buffer = buffer;
	transfer->length = length;
	transfer->user_data = user_data;
	transfer->callback = callback;
}
#endif
