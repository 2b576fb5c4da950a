/* TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <libusb-1.0/libusb.h>.
 *
 * libusb is not installed in the build image.  The reference's public header
 * (perseus-sdr.h:35) includes it only for a handful of opaque type names, so
 * this stub declares exactly those and nothing else.  It lets oracle/ref_harness.c
 * compile the reference's example callbacks where they lie under /root/reference.
 * Never included by the product (libperseus-sdr_b200/, include/). */
#ifndef PERSEUS_ORACLE_LIBUSB_STUB_H
#define PERSEUS_ORACLE_LIBUSB_STUB_H
#include <stdint.h>
#define LIBUSB_CALL
typedef struct libusb_device        libusb_device;
typedef struct libusb_device_handle libusb_device_handle;
typedef struct libusb_context       libusb_context;
struct libusb_transfer;
#endif
