/*
 * perseus-gpu.h — B200-native (sm_100a) I/Q unpack for the Perseus SDR sample stream.
 *
 * Sits alongside libperseus-sdr's perseus-sdr.h.  It replaces exactly one thing: the
 * per-transfer sample arithmetic that the reference leaves to the user callback
 * registered with perseus_start_async_input() (perseus-sdr.h:247-248), as written in
 *     examples/perseustest.c:432-460   user_data_callback_c_u   24-bit I/Q -> int32 (MSB aligned)
 *     examples/perseustest.c:466-502   user_data_callback_c_f   24-bit I/Q -> float, /(INT_MAX-256)
 *     examples/simple.c:33-61          user_data_callback_c_u   (duplicate)
 * plus the hand-off of completed transfers into it (perseus-in.c:187-264).
 * Device control (firmware, FPGA, attenuators, libusb) is untouched and stays in
 * libperseus-sdr.
 *
 * Plain C ABI: opaque handles, pointers and sizes only.  No CUDA or C++ types appear
 * in any signature, so the header can be bound from C, cgo, ctypes, JNI...
 * It deliberately does NOT include perseus-sdr.h (which drags in <libusb-1.0/libusb.h>,
 * perseus-sdr.h:35) and re-declares none of its names, so the two headers can be included
 * in either order; the one shared type, perseus_input_callback (perseus-sdr.h:81), appears
 * here under the library's own name perseus_gpu_input_fn (identical signature).
 *
 * Conventions mirror the reference (perseus-sdr.h:317-366): functions return 0 (or a
 * non-negative count) on success and a negative code on failure; a human-readable
 * message for the calling thread's last failure is returned by perseus_gpu_errorstr().
 * There is NO CPU fallback: every entry point that computes fails with
 * PERSEUS_GPU_NODEVICE / PERSEUS_GPU_BADARCH when no sm_100 device is usable.
 *
 * Threads: a handle is a monitor -- the callback thread (ONE per handle: the reference calls
 * back strictly serially, perseus-sdr.c:736-770), application threads and the library's own
 * latency watchdog may all touch the same handle; they serialise.  Ownership is recursive: a
 * sink may call the synchronous plumbing (sync, memcpy, get_stats) of its own handle, but not
 * flush/close/unpack.  perseus_gpu_input_callback itself takes no lock in the common case
 * (the hand-off uses sys_membarrier, csrc/handle.h), so it costs nothing next to its
 * copy; entry points called while a handle is streaming pay one membarrier (~us).  Every entry
 * point that takes a handle leaves the CALLING thread's current CUDA device set to the
 * handle's device (cudaSetDevice, not restored).
 *
 * Wire format (perseustest.c:434,450-455): 6 bytes per complex sample,
 *     I0 I1 I2 Q0 Q1 Q2      (24-bit little-endian two's complement I, then Q)
 * Outputs, 8 bytes per complex sample, interleaved I,Q:
 *     int32  : I0<<8 | I1<<16 | I2<<24          (value << 8; low byte always 0)
 *     float  : (float)int32 / (INT_MAX - 256)   (reference scale; range [-1.00000012, +1.0])
 *     float (POW2): (float)int32 * 2^-31        (NOT what the reference computes; opt-in)
 * A buffer of n bytes holds n/6 samples; n%6 trailing bytes are ignored (perseustest.c:443).
 */
#ifndef _perseus_gpu_h
#define _perseus_gpu_h

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PERSEUS_GPU_ABI_VERSION 3   /* 3: host sink, perseus_gpu_config.direct_bytes / copy_threads (were reserved, 0 = default) and
                                       eager_gap_us (appended: struct_size), perseus_gpu_stats.host_blocks (was reserved): callers
                                       built against 2 keep working */

/* Same signature as perseus_input_callback (perseus-sdr.h:81): a pointer of either type converts to the
 * other without a cast.  Declared under its own name so this header never collides with perseus-sdr.h. */
typedef int (*perseus_gpu_input_fn)(void *buf, int buf_size, void *extra);

typedef struct perseus_gpu perseus_gpu;               /* one per (device, receiver stream) */
typedef struct perseus_gpu_plan perseus_gpu_plan;     /* a reusable batched-launch layout  */
typedef struct perseus_vrx perseus_vrx;               /* synthetic receiver (stands in for the USB device) */

/* ---- error codes -------------------------------------------------------------------
 * Shared meanings reuse the reference's values (perseus-sdr.h:317-343); GPU-specific ones
 * continue below PERSEUS_SNNOTAVAILABLE (-26). */
#define PERSEUS_GPU_NOERROR        0
#define PERSEUS_GPU_NULLHANDLE    -2   /* = PERSEUS_NULLDESCR   */
#define PERSEUS_GPU_ASYNCSTARTED -19   /* = PERSEUS_ASYNCSTARTED */
#define PERSEUS_GPU_NOMEM        -20   /* = PERSEUS_NOMEM       */
#define PERSEUS_GPU_ERRPARAM     -22   /* = PERSEUS_ERRPARAM    */
#define PERSEUS_GPU_BUFFERSIZE   -24   /* = PERSEUS_BUFFERSIZE  */
#define PERSEUS_GPU_IOERROR      -13   /* = PERSEUS_IOERROR     */
#define PERSEUS_GPU_CUDAERR      -40   /* a CUDA runtime call or kernel failed (message has the CUDA text) */
#define PERSEUS_GPU_NODEVICE     -41   /* no CUDA device / driver */
#define PERSEUS_GPU_BADARCH      -42   /* device is not compute capability 10.x (the library only carries sm_100a code) */
#define PERSEUS_GPU_MISMATCH     -43   /* perseus_gpu_verify found differing samples */

/* ---- format / behaviour flags for the `flags` arguments -------------------------------- */
#define PERSEUS_GPU_OUT_INT32       0x0001u  /* write out_i32 (user_data_callback_c_u) */
#define PERSEUS_GPU_OUT_FLOAT       0x0002u  /* write out_f32, reference scale (user_data_callback_c_f) */
#define PERSEUS_GPU_OUT_FLOAT_POW2  0x0004u  /* write out_f32, 2^-31 scale (mutually exclusive with OUT_FLOAT) */
#define PERSEUS_GPU_ASYNC           0x0100u  /* enqueue only; complete with perseus_gpu_sync() */
#define PERSEUS_GPU_CHECKSUM        0x0200u  /* perseus_gpu_unpack: also accumulate the checksum (see perseus_gpu_checksum) of
                                                every output it produces, piece by piece on the streams that produce them, so
                                                it overlaps the copies; read the totals with perseus_gpu_get_checksums() */
/* flags == 0 means: produce whatever non-NULL output pointers were passed (float = reference scale). */

/* kernel selection (perseus_gpu_tuning.variant) */
#define PERSEUS_GPU_VARIANT_AUTO    0  /* = STREAM */
#define PERSEUS_GPU_VARIANT_STREAM  1  /* TMA bulk copy -> shared-memory ring -> coalesced stores; any pointer alignment */
#define PERSEUS_GPU_VARIANT_DIRECT  2  /* register-only kernel, any alignment: kept as the A/B comparison */

typedef struct perseus_gpu_tuning {
	int variant;       /* PERSEUS_GPU_VARIANT_*                                      (0 = auto)   */
	int tile_bytes;    /* input bytes per pipeline stage: 6144, 9216, 12288, 18432 or 24576         */
	int stages;        /* shared-memory ring depth, 2..8                                            */
	int ctas_per_sm;   /* persistent CTAs per SM, 1..8                                              */
	                   /* 0 in the three fields above = chosen per output format at launch time     */
	int store_mode;    /* 0 = default, 1 = st.global.cs (streaming), 2 = plain st.global           */
	int reserved[3];
} perseus_gpu_tuning;

typedef struct perseus_gpu_config {
	uint32_t struct_size;     /* = sizeof(perseus_gpu_config); lets the struct grow compatibly */
	int32_t  device;          /* CUDA ordinal                                                  */
	uint32_t stream_flags;    /* PERSEUS_GPU_OUT_* produced by the callback (streaming) path; 0 = INT32 */
	uint32_t nslabs;          /* pinned slabs in the hand-off ring, >= 2          (0 = 4)       */
	uint64_t slab_bytes;      /* bytes per slab, rounded down to a multiple of 48 (0 = 8 MiB)   */
	uint32_t nstreams;        /* CUDA streams the streaming path rotates its slabs over, 1..8 (0 = 2) */
	uint32_t max_latency_us;  /* streaming path: a partly filled slab is submitted once its oldest transfer has
	                             waited this long, checked at every callback and by the watchdog (0 = 50 000 us;
	                             0xFFFFFFFF = only when full).  See also eager_gap_us: at a real receiver's rates
	                             transfers do not wait at all.                                         */
	uint64_t chunk_bytes;     /* host<->device staging chunk for perseus_gpu_unpack with host pointers, rounded down to a
	                             multiple of 12288 (whole pages in and out; of 48 below that)  (0 = 32 MiB)  */
	perseus_gpu_tuning tuning;
	uint32_t options;         /* PERSEUS_GPU_OPT_*                                              */
	uint32_t stage_slots;     /* staging slots of the host-pointer pipeline, 2..8   (0 = 3): chunk c is copied in while
	                             chunk c-1 is unpacked and chunk c-2 is copied out, each on its own stream */
	uint32_t direct_bytes;    /* streaming path: a slab of at most this many bytes is unpacked by ONE kernel launch that
	                             reads the pinned slab over the link itself (no copy into HBM first) and, when only the host
	                             wants the samples (file / host sink, no device sink), stores them straight into pinned host
	                             memory: the short way for the small, latency-bounded slabs of a real receiver.  Larger
	                             slabs are copied by the copy engine.  (0 = 256 KiB; 0xFFFFFFFF = never)             */
	uint32_t copy_threads;    /* perseus_gpu_unpack with PAGEABLE host buffers (malloc, mmap): threads, the caller included, that move
	                             each chunk between the application's memory and pinned bounce buffers while the copy engines
	                             work on the neighbouring chunks  (0 = min(8, cores/2); 1 = the caller alone; 0xFFFFFFFF = hand
	                             pageable pointers to the CUDA runtime, which stages them on the calling thread)           */
	uint32_t eager_gap_us;    /* streaming path: a transfer that arrives more than this long after the previous callback returned
	                             is submitted at once instead of waiting for its slab to fill or to age: the stream
	                             is slower than the GPU path (any real receiver: a transfer every 0.5 ms at 2 MS/s, every
	                             10.8 ms at 95 kS/s), so every transfer is in device memory -- and at the host sink -- 15-30 us after its callback; transfers
	                             that arrive back to back (replayed recordings, bursts) still fill slabs.  Only with a latency
	                             bound (max_latency_us != 0xFFFFFFFF).  (0 = 100 us; 0xFFFFFFFF = never: slabs go out full or
	                             over age only)                                                                            */
	uint32_t reserved2[3];    /* the struct may grow: struct_size tells the library which fields the caller knows     */
} perseus_gpu_config;

/* perseus_gpu_config.options */
#define PERSEUS_GPU_OPT_NO_WATCHDOG 0x0001u  /* do not start the latency watchdog thread (see perseus_gpu_poll) */

/* ---- life cycle (names fixed by BASELINE.json north_star) ------------------------------ */

/* Creates a handle bound to cfg->device (cfg == NULL: device 0, all defaults).  Allocates the
 * CUDA streams now; staging memory is allocated on first use of the path that needs it. */
int perseus_gpu_open(perseus_gpu **h, const perseus_gpu_config *cfg);

/* Flushes the streaming path, waits for all queued work, frees everything.  Returns the
 * first latched asynchronous error, if any (the reference ignores callback return values,
 * perseus-in.c:207, so errors raised inside perseus_gpu_input_callback surface here or
 * at perseus_gpu_flush). */
int perseus_gpu_close(perseus_gpu *h);

/* Bulk unpack: the whole of user_data_callback_c_u / _c_f for one buffer of any size.
 *   buf      nbytes of wire data; device, pinned-host or pageable-host memory (detected).  Pageable memory is
 *            staged through pinned bounce buffers by cfg->copy_threads threads; a call with pageable OUTPUTS has
 *            completed them when it returns, PERSEUS_GPU_ASYNC or not.
 *   out_i32  2*(nbytes/6) int32, or NULL;  out_f32  2*(nbytes/6) float, or NULL;
 *            device or host memory (detected).  Host pointers are staged through the
 *            device in cfg->chunk_bytes pieces with copies and kernels overlapped.
 *   flags    PERSEUS_GPU_OUT_* | PERSEUS_GPU_ASYNC, or 0.
 * Returns the number of complex samples produced (>= 0) or a negative error.
 * The wire pointer may have ANY alignment and never matters for speed.  Output pointers must be 4-byte aligned
 * (PERSEUS_GPU_ERRPARAM otherwise; they hold int32 / float) and then run at full speed whatever their phase within 16
 * bytes -- except when out_i32 and out_f32 sit at DIFFERENT phases, which the same kernel serves with 32-bit stores. */
int64_t perseus_gpu_unpack(perseus_gpu *h, const void *buf, size_t nbytes,
                           void *out_i32, void *out_f32, unsigned flags);

/* Waits for everything queued on the handle; returns the first latched error. */
int perseus_gpu_sync(perseus_gpu *h);

/* Checksums of the int32 and float outputs of the most recent perseus_gpu_unpack(..., PERSEUS_GPU_CHECKSUM) call
 * (word index 0 = first output word of that call; a format that was not produced reports 0).  Waits for the call. */
int perseus_gpu_get_checksums(perseus_gpu *h, uint64_t *sum_i32, uint64_t *sum_f32);

/* ---- batched launch: many independent receivers in ONE kernel launch --------------------- */
typedef struct perseus_gpu_seg {
	const void *in;       /* device pointer, wire bytes of this receiver's window */
	size_t      nbytes;   /* nbytes/6 samples are produced                         */
	void       *out_i32;  /* device pointer or NULL                                */
	void       *out_f32;  /* device pointer or NULL                                */
} perseus_gpu_seg;

/* One launch over all segments (device pointers only).  Returns total samples or < 0. */
int64_t perseus_gpu_unpack_batch(perseus_gpu *h, const perseus_gpu_seg *segs, int nseg, unsigned flags);

/* Same, with the tile -> segment map built and uploaded once and reused across launches. */
int     perseus_gpu_plan_create(perseus_gpu *h, const perseus_gpu_seg *segs, int nseg, unsigned flags,
                                perseus_gpu_plan **plan);
int64_t perseus_gpu_plan_run(perseus_gpu *h, perseus_gpu_plan *plan, unsigned flags /* 0 or PERSEUS_GPU_ASYNC */);
int     perseus_gpu_plan_destroy(perseus_gpu *h, perseus_gpu_plan *plan);

/* ---- streaming hand-off: the drop-in for the reference's user callback -------------------
 *
 * perseus_gpu_input_callback has the signature of perseus_input_callback (perseus-sdr.h:81) with
 * extra = perseus_gpu*, so existing code does
 *     perseus_start_async_input(descr, 6144, perseus_gpu_input_callback, h);
 * It copies buf (valid only during the call: the transfer is resubmitted right after,
 * perseus-in.c:263) into the current pinned slab; a full slab is sent H2D and unpacked on
 * one of the handle's streams while the next slab fills.  Never blocks on the GPU except
 * for back-pressure when every slab is in flight (counted in perseus_gpu_stats.stalls; the
 * wait sleeps, it does not spin: the reference calls this on a SCHED_FIFO thread,
 * perseus-sdr.c:749-753).  Always returns 0, like the reference's callbacks
 * (perseustest.c:459,501); after an error has been latched, further transfers are counted
 * in perseus_gpu_stats.dropped_callbacks / dropped_bytes until flush/sync/close reports it. */
int perseus_gpu_input_callback(void *buf, int buf_size, void *extra);

/* Optional: does now what the streaming path would otherwise do inside its FIRST callback -- pinned slabs, device buffers,
 * events, the pinned output copies when a file / host sink is set, the first (code-loading) kernel launch, the watchdog: tens of
 * milliseconds -- so that the first transfers of a real receiver are not held up on libperseus-sdr's poll thread (its ring of 8
 * transfers covers 4 ms at 2 MS/s, perseus-sdr.c:683).  Call after perseus_gpu_set_host_sink / perseus_gpu_stream_to_file and
 * before perseus_start_async_input.  Harmless at any other time (in mid-stream it only allocates what is still missing). */
int perseus_gpu_prepare(perseus_gpu *h);

/* Latency bound without a following callback.  A partly filled slab is submitted once its oldest
 * transfer has waited cfg.max_latency_us.  That is checked at every callback and, because a stream
 * can stall (USB error, the last transfers before perseus_stop_async_input), also by a small
 * watchdog thread the handle starts with its first callback (it sleeps until the partial slab's deadline;
 * PERSEUS_GPU_OPT_NO_WATCHDOG disables it).  Applications that prefer to drive it themselves call
 * perseus_gpu_poll() from any thread: it submits the partial slab if it is over age and returns the
 * number of slabs it submitted (0 or 1), or a negative error. */
int perseus_gpu_poll(perseus_gpu *h);

/* A block of unpacked samples resident in device memory.  Passed to the sink on the thread
 * that submitted the slab (the callback thread, the caller of flush/poll, or the watchdog),
 * under the handle's lock, right after the unpack kernel was ENQUEUED on `stream`
 * (a cudaStream_t): work the sink enqueues on that stream runs after the unpack and before
 * the block's memory is reused.  (With a device sink set, slabs always land in device memory;
 * see perseus_gpu_config.direct_bytes.) */
typedef struct perseus_gpu_block {
	uint64_t first_sample;   /* index of the block's first complex sample since open */
	uint64_t nsamples;
	void    *dev_i32;        /* device pointers (NULL when the format is not produced) */
	void    *dev_f32;
	void    *stream;         /* cudaStream_t */
} perseus_gpu_block;
typedef void (*perseus_gpu_sink)(const perseus_gpu_block *blk, void *extra);
int perseus_gpu_set_sink(perseus_gpu *h, perseus_gpu_sink sink, void *extra);

/* A block of unpacked samples in (pinned) HOST memory: what the reference's callbacks have in hand when they fwrite
 * (perseustest.c:457,499), for applications whose consumer stays on the CPU.  The host sink is called once per slab, in
 * stream order, as soon as the slab's samples have arrived in host memory -- no flush needed -- on the handle's delivery
 * thread, NOT under the handle's lock: it must not call perseus_gpu_* on this handle (a flush waits for the sink), and the
 * pointers are valid only during the call.  Setting or clearing the sink flushes the stream first.  */
typedef struct perseus_gpu_host_block {
	uint64_t    first_sample;  /* index of the block's first complex sample since open */
	uint64_t    nsamples;
	const void *i32;           /* nsamples x {int32 I, int32 Q}, or NULL when the handle does not stream int32 */
	const void *f32;           /* nsamples x {float I, float Q}, or NULL when the handle does not stream floats */
} perseus_gpu_host_block;
typedef void (*perseus_gpu_host_sink)(const perseus_gpu_host_block *blk, void *extra);
int perseus_gpu_set_host_sink(perseus_gpu *h, perseus_gpu_host_sink sink, void *extra);

/* Also writes the stream to `path` exactly as `perseustest -o path [-p]` would: a raw,
 * headerless sequence of {int32 I,int32 Q} (or {float I,float Q} when the handle streams
 * floats), 8 bytes per sample (perseustest.c:337-343,457,499).  path "-" is standard output, as there (perseustest.c:98,337): a consumer on a
 * pipe receives the stream as it is unpacked.  path == NULL stops.  Blocks are written as they complete (same
 * delivery as the host sink); perseus_gpu_flush / close make the file complete. */
int perseus_gpu_stream_to_file(perseus_gpu *h, const char *path);

/* Submits the partly filled slab and waits until every submitted slab has been unpacked
 * (and written to the file sink).  Returns the first latched error. */
int perseus_gpu_flush(perseus_gpu *h);

/* ---- observability (mirrors the counters perseus_stop_async_input prints, perseus-sdr.c:719-722) */
typedef struct perseus_gpu_stats {
	uint64_t kernel_launches;   /* launches of this library's kernels                 */
	uint64_t samples;           /* complex samples unpacked                            */
	uint64_t bytes_in;          /* wire bytes consumed                                 */
	uint64_t h2d_bytes;         /* bytes copied host -> device by the library          */
	uint64_t d2h_bytes;         /* bytes copied device -> host by the library          */
	uint64_t callbacks;         /* perseus_gpu_input_callback invocations              */
	uint64_t slabs;             /* slabs submitted by the streaming path               */
	uint64_t stalls;            /* times the callback had to wait for a free slab      */
	uint64_t dropped_callbacks; /* callbacks ignored because an error was latched      */
	uint64_t dropped_bytes;     /* wire bytes of those callbacks                       */
	uint64_t watchdog_submits;  /* partial slabs submitted by the watchdog / poll      */
	uint64_t host_blocks;       /* blocks written to the stream file / handed to the host sink */
} perseus_gpu_stats;
int perseus_gpu_get_stats(perseus_gpu *h, perseus_gpu_stats *out);

/* Optional self-calibration.  The pipeline's best geometry is a narrow optimum of the wire bytes in flight per SM
 * (DESIGN.md §4); the built-in defaults were measured on B200.  This call measures a dozen candidate geometries on
 * scratch memory of THIS device (about 50 ms, ~2 GB of temporary device memory) for one-format and for fused
 * launches and keeps the winners for the handle; fields set explicitly with perseus_gpu_set_tuning still win.
 * gbs_single / gbs_fused (optional) receive the algorithmic GB/s of the winners. */
int perseus_gpu_autotune(perseus_gpu *h, double *gbs_single, double *gbs_fused);
int perseus_gpu_set_tuning(perseus_gpu *h, const perseus_gpu_tuning *t);   /* NULL = defaults */
int perseus_gpu_get_tuning(perseus_gpu *h, perseus_gpu_tuning *t);         /* what was set; 0 = automatic */
/* The geometry a launch producing `flags` (PERSEUS_GPU_OUT_*) would use now: explicit > autotuned > default. */
int perseus_gpu_get_geometry(perseus_gpu *h, unsigned flags, int *tile_bytes, int *stages, int *ctas_per_sm);

/* Last error message of the calling thread (cf. perseus_errorstr(), perseuserr.c:36-42). */
const char *perseus_gpu_errorstr(void);
/* "perseus-gpu <abi> sm_100a ..." */
const char *perseus_gpu_version(void);

/* ---- device plumbing, so a C caller needs no CUDA of its own ------------------------------ */
int   perseus_gpu_device_count(void);
int   perseus_gpu_device_info(int device, char *name, size_t name_len, int *sm_count,
                              int *cc_major, int *cc_minor, uint64_t *total_mem);
void *perseus_gpu_dev_alloc(perseus_gpu *h, size_t nbytes);          /* NULL on failure */
int   perseus_gpu_dev_free(perseus_gpu *h, void *p);
void *perseus_gpu_host_alloc(perseus_gpu *h, size_t nbytes);         /* pinned host memory */
int   perseus_gpu_host_free(perseus_gpu *h, void *p);
int   perseus_gpu_memcpy(perseus_gpu *h, void *dst, const void *src, size_t nbytes);  /* any direction, synchronous */
int   perseus_gpu_memset(perseus_gpu *h, void *dev, int byte, size_t nbytes);
void *perseus_gpu_get_stream(perseus_gpu *h, int idx);               /* cudaStream_t idx of the handle */
/* Timing on the stream the kernels are launched on: records event slot `slot` (0..31) on
 * stream 0 of the handle; elapsed time between two recorded slots in milliseconds.  All 32
 * slots belong to the caller (autotune and the probes time with private events). */
int   perseus_gpu_event_record(perseus_gpu *h, int slot);
int   perseus_gpu_event_elapsed_ms(perseus_gpu *h, int slot_start, int slot_stop, float *ms);

/* ---- synthetic wire data (replaces the USB device for benchmarks) ------------
 * RANDOM: the byte stream whose 64-bit little-endian word w is splitmix64(seed + w); every
 *         24-bit field is an independent uniform draw.  Random access via byte_offset.
 * RAMP:   sample v has I = v mod 2^24, Q = (uint32)(v*2654435761) >> 8; byte_offset must be a
 *         multiple of 6 (it selects the first sample). */
#define PERSEUS_SYNTH_RANDOM 0
#define PERSEUS_SYNTH_RAMP   1
#define PERSEUS_SYNTH_SEED   0x5045525345555300ull   /* "PERSEUS\0" */
int perseus_gpu_generate(perseus_gpu *h, void *dev_dst, size_t nbytes, int pattern, uint64_t seed, uint64_t byte_offset);
int perseus_synth_fill(void *host_dst, size_t nbytes, int pattern, uint64_t seed, uint64_t byte_offset);

/* ---- on-device verification for recordings too large to bring back ---------------------------
 * checksum: sum over 32-bit words w[i] of (splitmix64(first_index+i)|1) * (w[i]+1) mod 2^64;
 *           additive over shards, so per-GPU results add up to the whole recording's. */
int perseus_gpu_checksum(perseus_gpu *h, const void *dev_words, size_t nwords, uint64_t first_index, uint64_t *sum);
/* verify: recomputes every sample from the wire bytes with an independent one-thread-per-sample
 *         kernel (byte loads, IEEE division) and counts 32-bit words that differ from the given
 *         outputs.  Returns 0 if none, PERSEUS_GPU_MISMATCH otherwise; *nmismatch and
 *         *first_bad_word (index into the interleaved output) are filled when non-NULL. */
int perseus_gpu_verify(perseus_gpu *h, const void *dev_in, size_t nbytes, const void *dev_i32, const void *dev_f32,
                       unsigned flags, uint64_t *nmismatch, uint64_t *first_bad_word);

/* ---- in-run roofline context: what THIS device sustains on one-directional HBM streams -----------------
 * kind 0 = read nbytes, 1 = write nbytes, 2 = copy nbytes (traffic counted = 2*nbytes).  Runs `reps` launches of a plain
 * streaming kernel (best grid of a small built-in set) on scratch memory and returns the best GB/s in *gbs. */
#define PERSEUS_GPU_PROBE_READ  0
#define PERSEUS_GPU_PROBE_WRITE 1
#define PERSEUS_GPU_PROBE_COPY  2
int perseus_gpu_probe_hbm(perseus_gpu *h, int kind, size_t nbytes, int reps, double *gbs);

/* ---- in-run PCIe roofline: plain pinned-memory copies, best of `reps` ---------------------------------------
 * kind H2D: nbytes host->device alone; D2H: d2h_nbytes device->host alone; DUPLEX: both queued at once on two
 * streams -- each direction's rate over its own copy, so with d2h_nbytes = nbytes/6*16 (the fused round trip's
 * mix) max(nbytes/h2d, d2h_nbytes/d2h) is what plain copies need for that traffic.  d2h_nbytes 0 = nbytes.
 * *h2d_gbs / *d2h_gbs receive GB/s (0 for a direction not run). */
#define PERSEUS_GPU_PCIE_H2D    0
#define PERSEUS_GPU_PCIE_D2H    1
#define PERSEUS_GPU_PCIE_DUPLEX 2
int perseus_gpu_probe_pcie(perseus_gpu *h, int kind, size_t nbytes, size_t d2h_nbytes, int reps, double *h2d_gbs, double *d2h_gbs);

/* ---- multi-GPU sharding: contiguous ranges of whole transfers, no exchange --- */
/* Shard `shard` of `nshards` of a recording of `total_buffers` transfers gets
 * [*first, *first + *count) with first = floor(shard*total/nshards). */
int perseus_gpu_shard_range(uint64_t total_buffers, int nshards, int shard, uint64_t *first, uint64_t *count);
/* Same, with shards proportional to weights[0..nshards) (>= 0, not all zero) -- e.g. each GPU's measured host-link rate, for
 * host-fed recordings on a box whose GPUs do not all reach host memory equally fast: shard s gets
 * [floor(total * W(s) / W(nshards)), floor(total * W(s+1) / W(nshards))) with W(k) = weights[0] + ... + weights[k-1].
 * Every rank must pass the same weights.  Equal weights give exactly perseus_gpu_shard_range. */
int perseus_gpu_shard_range_weighted(uint64_t total_buffers, int nshards, const double *weights, int shard,
                                     uint64_t *first, uint64_t *count);

/* ---- virtual receiver: the reference's delivery semantics without the hardware ----------------
 * Reproduces what perseus_start_async_input() + perseus-in.c do with a real device: a ring of
 * 8 equal buffers in ONE contiguous pageable allocation (perseus-in.c:68, perseus-sdr.c:683),
 * filled with synthetic wire data and handed to the callback strictly in ring order, each
 * buffer valid only during its callback (it is refilled right after, perseus-in.c:263). */
#define PERSEUS_VRX_QUEUE_SIZE 8          /* perseus-sdr.c:683 */
#define PERSEUS_VRX_MAX_BUFFER 16320      /* perseus-sdr.c:662 */

typedef struct perseus_vrx_config {
	uint32_t struct_size;
	int32_t  sample_rate;      /* requested S/s; snapped like perseus_set_sampling_rate (perseus-sdr.c:776-811) */
	int32_t  ep_max_packet;    /* 512 (fw 24v41: sizes % 6144) or 510 (legacy fw: sizes % 510); 0 = 512 */
	int32_t  pattern;          /* PERSEUS_SYNTH_*                                               */
	uint64_t seed;             /* 0 = PERSEUS_SYNTH_SEED                                        */
	int32_t  realtime;         /* non-zero: pace deliveries at sample_rate; 0: as fast as possible */
	/* fault injection (perseus-in.c:204-216 log-and-drop cases); 0 = never */
	uint32_t drop_every;       /* every Nth transfer completes short -> not delivered           */
	uint32_t swap_every;       /* every Nth transfer completes out of sequence -> not delivered */
	uint32_t replay;           /* non-zero: generate only the first 8 transfers and re-deliver the ring's contents
	                              (stream repeats every 8 transfers): isolates the hand-off cost in benchmarks */
	/* transfer statuses other than COMPLETED (perseus-in.c:218-257) */
	uint32_t timeout_every;    /* every Nth transfer completes TIMED_OUT: nothing delivered or counted, slot re-armed */
	uint32_t fail_at;          /* the Nth transfer (1-based; 0 = never) completes with fail_status: its slot is retired */
	uint32_t fail_status;      /* PERSEUS_VRX_STATUS_*: ERROR, STALL, NO_DEVICE or OVERFLOW                             */
	uint32_t reserved[1];
} perseus_vrx_config;

/* libusb transfer statuses the reference's completion handler distinguishes (values = enum libusb_transfer_status) */
#define PERSEUS_VRX_STATUS_COMPLETED 0
#define PERSEUS_VRX_STATUS_ERROR     1
#define PERSEUS_VRX_STATUS_TIMED_OUT 2
#define PERSEUS_VRX_STATUS_CANCELLED 3
#define PERSEUS_VRX_STATUS_STALL     4
#define PERSEUS_VRX_STATUS_NO_DEVICE 5
#define PERSEUS_VRX_STATUS_OVERFLOW  6

typedef struct perseus_vrx_stats {        /* cf. perseus-sdr.c:719-722 */
	uint64_t bytes_received;   /* counts every completed transfer, delivered or not (perseus-in.c:202) */
	uint64_t delivered;        /* callbacks made                                   */
	uint64_t dropped_short;    /* perseus-in.c:209-212                             */
	uint64_t dropped_sequence; /* perseus-in.c:213-216                             */
	double   elapsed_s;
	double   ksamples_per_s;   /* bytes_received / elapsed / 6000, as the reference prints it */
	uint64_t timed_out;        /* perseus-in.c:218-221: logged, slot re-armed                 */
	uint64_t retired;          /* perseus-in.c:222-257: slots taken out of the ring for good  */
} perseus_vrx_stats;

int perseus_vrx_open(perseus_vrx **v, const perseus_vrx_config *cfg);
int perseus_vrx_close(perseus_vrx *v);
/* The ten FPGA bitstream rates (perseus-sdr.h:282-285; table order = ascending). */
int perseus_vrx_get_sampling_rates(int *buf, unsigned int size);
/* Nearest-rate selection with the reference's midpoint rule; returns the rate chosen. */
int perseus_vrx_nearest_rate(int requested);
int perseus_vrx_get_sampling_rate(perseus_vrx *v);
/* Name of the FPGA bitstream that produces `rate` (one of the ten table rates), e.g. 2000000 -> "perseus2m24v21";
 * the names are the reference's *.rbs files (generate_fpga_code.sh:71-97).  NULL for a rate not in the table.
 * All ten are 24-bit formats ("...24v..."): the wire layout is the same at every rate. */
const char *perseus_vrx_bitstream_name(int rate);
/* Same validation and error codes as perseus_start_async_input (perseus-sdr.c:638-692):
 * buffersize <= 16320 and a multiple of 6144 (EP 512) / 510 (EP 510).  Starts a delivery
 * thread.  stop cancels, joins, and fills the statistics. */
int perseus_vrx_start_async_input(perseus_vrx *v, uint32_t buffersize, perseus_gpu_input_fn callback, void *cb_extra);
int perseus_vrx_stop_async_input(perseus_vrx *v);
/* Synchronous variant for benchmarks and tests: delivers exactly `ntransfers` completed
 * transfers on the calling thread (dropped ones count), then returns. */
int perseus_vrx_run(perseus_vrx *v, uint32_t buffersize, perseus_gpu_input_fn callback, void *cb_extra,
                    uint64_t ntransfers);
int perseus_vrx_get_stats(perseus_vrx *v, perseus_vrx_stats *out);

#ifdef __cplusplus
}
#endif

#endif /* _perseus_gpu_h */
