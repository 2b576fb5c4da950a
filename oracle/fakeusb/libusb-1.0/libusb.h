/* TEST INFRASTRUCTURE ONLY (oracle/): a fake libusb-1.0, just large enough for the reference's OWN
 * library (/root/reference/perseus-sdr.c, perseusfx2.c, perseus-in.c, perseuserr.c) to compile
 * and run UNMODIFIED without hardware.  The names and the public fields follow the libusb-1.0 API
 * (that is what the reference is written against); the implementation (oracle/fakeusb.c) is a
 * synthetic Perseus receiver: an FX2 that accepts the firmware download and the EP1 command
 * protocol of perseusfx2.c, and a bulk-IN endpoint 0x82 that completes submitted transfers in
 * submission order with generated wire data.  Never included by the product. */
#ifndef PERSEUS_ORACLE_FAKE_LIBUSB_H
#define PERSEUS_ORACLE_FAKE_LIBUSB_H
#include <stdint.h>
#include <sys/types.h>
#include <limits.h>   /* the real libusb.h includes it too; examples/perseustest.c relies on that for INT_MAX */
#include <sys/time.h>

#define LIBUSB_CALL

typedef struct libusb_context libusb_context;
typedef struct libusb_device libusb_device;
typedef struct libusb_device_handle libusb_device_handle;   /* here: the synthetic receiver */

enum libusb_error {
	LIBUSB_SUCCESS = 0, LIBUSB_ERROR_IO = -1, LIBUSB_ERROR_INVALID_PARAM = -2, LIBUSB_ERROR_ACCESS = -3,
	LIBUSB_ERROR_NO_DEVICE = -4, LIBUSB_ERROR_NOT_FOUND = -5, LIBUSB_ERROR_BUSY = -6, LIBUSB_ERROR_TIMEOUT = -7,
	LIBUSB_ERROR_OVERFLOW = -8, LIBUSB_ERROR_PIPE = -9, LIBUSB_ERROR_INTERRUPTED = -10, LIBUSB_ERROR_NO_MEM = -11,
	LIBUSB_ERROR_NOT_SUPPORTED = -12, LIBUSB_ERROR_OTHER = -99
};

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

struct libusb_device_descriptor {
	uint8_t  bLength, bDescriptorType;
	uint16_t bcdUSB;
	uint8_t  bDeviceClass, bDeviceSubClass, bDeviceProtocol, bMaxPacketSize0;
	uint16_t idVendor, idProduct, bcdDevice;
	uint8_t  iManufacturer, iProduct, iSerialNumber, bNumConfigurations;
};

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
	uint64_t fake_submit_ns;
};

/* library + enumeration */
int  libusb_init(libusb_context **ctx);
void libusb_exit(libusb_context *ctx);
void libusb_set_debug(libusb_context *ctx, int level);
ssize_t libusb_get_device_list(libusb_context *ctx, libusb_device ***list);
void libusb_free_device_list(libusb_device **list, int unref_devices);
int  libusb_get_device_descriptor(libusb_device *dev, struct libusb_device_descriptor *desc);
uint8_t libusb_get_bus_number(libusb_device *dev);
uint8_t libusb_get_device_address(libusb_device *dev);
libusb_device *libusb_ref_device(libusb_device *dev);
void libusb_unref_device(libusb_device *dev);
int  libusb_get_max_packet_size(libusb_device *dev, unsigned char endpoint);
/* device handles */
int  libusb_open(libusb_device *dev, libusb_device_handle **handle);
void libusb_close(libusb_device_handle *handle);
int  libusb_kernel_driver_active(libusb_device_handle *handle, int interface_number);
int  libusb_detach_kernel_driver(libusb_device_handle *handle, int interface_number);
int  libusb_set_configuration(libusb_device_handle *handle, int configuration);
int  libusb_claim_interface(libusb_device_handle *handle, int interface_number);
int  libusb_release_interface(libusb_device_handle *handle, int interface_number);
int  libusb_set_interface_alt_setting(libusb_device_handle *handle, int interface_number, int alternate_setting);
int  libusb_clear_halt(libusb_device_handle *handle, unsigned char endpoint);
/* synchronous I/O */
int  libusb_control_transfer(libusb_device_handle *handle, uint8_t request_type, uint8_t bRequest, uint16_t wValue, uint16_t wIndex,
                             unsigned char *data, uint16_t wLength, unsigned int timeout);
int  libusb_bulk_transfer(libusb_device_handle *handle, unsigned char endpoint, unsigned char *data, int length, int *actual_length,
                          unsigned int timeout);
const char *libusb_error_name(int errcode);
/* libusb_strerror is deliberately NOT declared: without config.h the reference supplies its own (perseusfx2.c:70-72). */
/* asynchronous I/O */
struct libusb_transfer *libusb_alloc_transfer(int iso_packets);
void libusb_free_transfer(struct libusb_transfer *transfer);
int  libusb_submit_transfer(struct libusb_transfer *transfer);
int  libusb_cancel_transfer(struct libusb_transfer *transfer);
int  libusb_handle_events_timeout(libusb_context *ctx, struct timeval *tv);

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
