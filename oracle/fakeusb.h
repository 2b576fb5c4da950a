/* TEST INFRASTRUCTURE ONLY — control surface of the synthetic receiver in oracle/fakeusb.c
 * (mirrored with ctypes by oracle/oracle.py). */
#ifndef PERSEUS_ORACLE_FAKEUSB_H
#define PERSEUS_ORACLE_FAKEUSB_H
#include <stdint.h>

typedef struct fakeusb_config {
	uint32_t struct_size;
	uint32_t blank_eeprom;    /* 1: enumerates as a bare Cypress FX2 (04B4:8613) that needs the firmware download */
	uint32_t ep_max_packet;   /* wMaxPacketSize of EP 0x82: 512 (fw 24v41) or 510 (legacy fw); 0 = 512 */
	uint32_t realtime;        /* pace the stream at the configured bitstream's sample rate */
	uint64_t seed;            /* of the synthetic stream */
	uint64_t limit;           /* the stream stalls after this many transfers (0 = endless) */
	uint32_t drop_every;      /* every Nth transfer completes 6 bytes short */
	uint32_t swap_every;      /* every Nth transfer completes after its successor */
	uint32_t timeout_every;   /* every Nth transfer completes LIBUSB_TRANSFER_TIMED_OUT */
	uint32_t fail_at;         /* the Nth transfer (1-based) completes with fail_status */
	uint32_t fail_status;     /* enum libusb_transfer_status: ERROR, STALL, NO_DEVICE or OVERFLOW */
	uint32_t preserie;        /* EEPROM product code != 0x8014 (perseus-sdr.c:375-378) */
	uint32_t serial;
	uint32_t reserved;
} fakeusb_config;

typedef struct fakeusb_state {
	/* enumeration / handles */
	uint64_t inits, exits, opens, closes;
	int64_t  device_refs;
	int32_t  claimed, firmware_loaded;
	/* firmware download (perseusfx2.c:164-202) */
	uint64_t cpu_resets, fw_records, fw_bytes, fw_hash;
	/* FPGA configuration (perseusfx2.c:291-359) */
	uint64_t fpga_resets, fpga_bytes, fpga_hash;
	int32_t  fpga_rate;       /* sample rate of the recognised bitstream, 0 = none / unknown */
	int32_t  fifo_enabled;    /* PERSEUS_SIO_FIFOEN as last written */
	/* control state as last written */
	uint32_t sio_freg;
	uint8_t  sio_ctl, porte, pad[2];
	uint64_t sio_writes, porte_writes, eeprom_reads, shutdowns, commands;
	/* EP 0x82 */
	uint64_t submits, stream_pos, completed_ok, cancelled, timed_out, failed;
	/* the thread that runs the event loop (the reference's poll thread, perseus-sdr.c:736-774) */
	uint64_t events_calls;
	int32_t  events_thread_policy, events_thread_priority, events_thread_is_fifo, pad2;
} fakeusb_state;

typedef struct fakeusb_bitstream_id { int32_t rate; uint32_t pad; uint64_t size, fnv1a64; } fakeusb_bitstream_id;

#endif
