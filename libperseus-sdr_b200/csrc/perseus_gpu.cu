// C-ABI layer of libperseus_gpu.so (include/perseus-gpu.h): handle, streams, staging of host
// buffers, the perseus_input_callback trampoline with its pinned slab ring, batched plans,
// and the small device utilities.  All sample arithmetic lives in unpack_kernels.cu.
//
// Reference anchors: the callback contract is perseus-sdr.h:81 / perseus-in.c:204-207,263
// (buffer valid only during the call, return value ignored, strictly serial, in ring order);
// the error convention mirrors perseus-sdr.h:317-366 + perseuserr.c:36-42.
#include "../../include/perseus-gpu.h"
#include "copy_pool.h"
#include "kernels.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include <pthread.h>
#include <sched.h>
#include <time.h>
#include <unistd.h>
#if defined(__linux__)
#include <linux/membarrier.h>
#include <sys/syscall.h>
#endif
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__x86_64__)
#include <cpuid.h>
#include <x86intrin.h>
#endif

namespace {

using pg::fail;
inline int ok(int v = 0) { return v; }

constexpr int kMaxStreams = 8;
constexpr int kMaxSlabs = 64;
constexpr int kEventSlots = 32;
constexpr int kMaxStageSlots = 8;
constexpr size_t kDefaultDirectBytes = 256u << 10;   // perseus_gpu_config.direct_bytes = 0

struct Slab {
	uint8_t *host = nullptr;       // pinned wire bytes being filled by the callback
	uint8_t *dev_in = nullptr;     // device copy
	uint8_t *dev_i32 = nullptr;    // device outputs of this slab
	uint8_t *dev_f32 = nullptr;
	uint8_t *host_out[2] = {nullptr, nullptr};   // pinned copies of the outputs (int32, float): only with a file / host sink
	cudaEvent_t unpacked = nullptr;   // recorded behind the slab's kernel (and whatever the device sink queued): host delivery waits for it
	cudaEvent_t done = nullptr;    // recorded after the slab's last device operation (sleeping waits: back-pressure)
	cudaEvent_t ready = nullptr;   // the same point, for the delivery thread (spinning wait: lowest latency)
	uint64_t dlv_seq = 0;          // host delivery: this slab's place in the delivery order
	bool to_deliver = false;       // host delivery was queued for this use of the slab
	uint64_t first_sample = 0;
	uint64_t nsamples = 0;
	size_t file_bytes = 0;         // bytes deliver_slab writes to the file sink
	int file_fmt = 0;              // which host_out[] the file holds (perseustest writes ONE format per run)
	bool busy = false;
};

}  // namespace

struct perseus_gpu {
	// Ownership of the handle's state.  Every C-ABI entry point takes `mu` (recursive: a sink runs under it and may call the
	// plumbing of its own handle) -- EXCEPT perseus_gpu_input_callback, whose fast path must not execute a single locked
	// instruction: the slab copy uses non-temporal stores, and a locked instruction (mutex, atomic read-modify-write, mfence)
	// stalls until the write-combining buffers have drained, ~0.25 us per 6144-byte transfer, a quarter of the path's
	// throughput.  The callback thread and everybody else therefore meet in an ASYMMETRIC Dekker handshake:
	//   callback:  cb_active = 1;  <compiler barrier>;  if (others_want) -> slow path (take mu like everybody else)
	//   others:    lock mu;  others_want++;  membarrier(PRIVATE_EXPEDITED);  wait until cb_active == 0
	// sys_membarrier makes every running thread of the process execute a full barrier at that instant, which supplies the
	// store-load ordering the callback side leaves out (and drains that core's write-combining buffers).  Where the system call
	// is not available both sides fall back to a real fence.
	std::recursive_mutex mu;
	int lock_depth = 0;                          // under mu: nested entries of the owning thread
	std::atomic<uint32_t> cb_active{0};          // 1 while the callback thread is inside its lock-free fast path
	std::atomic<uint32_t> others_want{0};        // threads that hold or wait for `mu`
	std::atomic<unsigned long> cb_thread{0};     // the thread that set cb_active (a sink called from it already owns the handle)
	std::atomic<uint64_t> partial_since_ns{0};   // published by the owner for the watchdog: age stamp of the partial slab, 0 = none
	bool asym = false;                           // membarrier available: the callback side needs no fence
	int device = 0;
	int sm_count = 0;
	perseus_gpu_config cfg{};
	pg::Tuning tune{};
	int nstreams = 0;
	cudaStream_t streams[kMaxStreams]{};        // [0] = the launch stream of everything device-resident; slabs rotate over all
	cudaStream_t s_in = nullptr, s_out = nullptr;   // copy-in / copy-out streams of the host-pointer pipeline
	cudaStream_t s_dlv = nullptr;               // host delivery of the streaming path: every slab's copy-out + deliver_slab, in stream order
	cudaEvent_t events[kEventSlots]{};          // the caller's timing slots (perseus_gpu_event_*)
	cudaEvent_t tev[4]{};                       // private timing events (autotune, probes)
	unsigned long long *d_scratch = nullptr;   // 2 x u64: checksum / verify results
	unsigned long long *h_scratch = nullptr;   // pinned mirror
	unsigned long long *d_sums = nullptr;      // 2 x u64: running checksums of PERSEUS_GPU_CHECKSUM calls (int32, float)
	// host-pointer pipeline of perseus_gpu_unpack: chunk c lives in slot c % nslots; per slot three events order
	// copy-in -> kernel -> copy-out and protect the slot's reuse
	size_t chunk_bytes = 0;
	int nslots = 0;
	uint64_t stage_seq = 0;                     // chunks staged so far (slot rotation continues across calls)
	uint8_t *stage_in[kMaxStageSlots]{};
	uint8_t *stage_out[kMaxStageSlots][2]{};
	cudaEvent_t ev_in[kMaxStageSlots]{}, ev_k[kMaxStageSlots]{}, ev_out[kMaxStageSlots]{};
	// PAGEABLE host buffers are staged through pinned bounce buffers by the caller and the pool's helper threads
	// (copy_threads participants; 0 = leave pageable memory to the CUDA runtime, which stages it on the calling thread alone)
	int copy_threads = 0;
	pg::CopyPool *pool = nullptr;
	uint8_t *bounce_in[kMaxStageSlots]{};
	uint8_t *bounce_out[kMaxStageSlots][2]{};
	// streaming (callback) path
	unsigned stream_fmt = 0;
	size_t slab_bytes = 0;
	int nslabs = 0;
	Slab slabs[kMaxSlabs];
	int cur = 0;               // slab being filled
	size_t fill = 0;           // bytes in it
	uint64_t fill_started_ns = 0;   // monotonic time the first transfer of the current slab arrived
	uint64_t max_latency_ns = 0;    // 0 = submit only full slabs
	uint64_t eager_gap_ns = 0;      // a transfer arriving this long after the previous one is submitted at once; 0 = never
	uint64_t last_push_ns = 0;      // monotonic time the previous callback began (0 = none yet)
	// the callback's clock: CLOCK_MONOTONIC carried forward by the time-stamp counter between anchors (see callback_now_ns)
	bool use_tsc = false;
	uint64_t tsc_anchor = 0, ns_anchor = 0;      // one reading of both clocks
	uint64_t tsc_reanchor = 0;                   // ticks after which the anchor is renewed (about a second)
	uint64_t ns_per_tick_q32 = 0;                // nanoseconds per tick, 32.32 fixed point
	int next_to_write = 0;     // oldest slab whose output has not reached the file sink
	bool streaming_ready = false;
	uint64_t samples_submitted = 0;
	perseus_gpu_sink sink = nullptr;
	void *sink_extra = nullptr;
	// host delivery (file sink, host sink): written by the owner only while nothing is in flight (after a flush), read by
	// deliver_slab on the handle's delivery thread
	perseus_gpu_host_sink host_sink = nullptr;
	void *host_sink_extra = nullptr;
	FILE *fout = nullptr;
	bool fout_is_stdout = false;                 // path "-" (perseustest.c:98,337): flushed, never closed
	std::atomic<int> io_error{0};                // deliver_slab could not write the file: surfaced at the next retire / flush
	std::atomic<uint64_t> host_blocks{0};        // blocks deliver_slab has handed over
	// The delivery thread: takes slabs in submission order, waits (spinning) for each one's outputs to have reached pinned host
	// memory, writes the file / calls the host sink.  A thread of the handle's own rather than cudaLaunchHostFunc because the
	// runtime dispatches host functions 0.15 ms late (measured, profiles/r2_latency_probe.jsonl); this one is usually there
	// before the data is.  dlv_mu guards the three counters below and pairs with the two condition variables.
	std::thread dlv_thread;
	std::mutex dlv_mu;
	std::condition_variable dlv_cv;              // delivery thread: something was submitted / stop
	std::condition_variable dlv_done_cv;         // owner: something was delivered
	uint64_t dlv_submitted = 0, dlv_delivered = 0;
	std::atomic<uint64_t> dlv_submitted_hint{0}; // = dlv_submitted, for the thread's short spin before it sleeps
	int dlv_ring[kMaxSlabs]{};                   // slab index of delivery number n at [n % kMaxSlabs]
	bool dlv_started = false, dlv_stop = false;
	size_t direct_bytes = 0;                     // slabs up to this size are unpacked straight from the pinned slab (no H2D copy)
	// latency watchdog (started with the first callback unless PERSEUS_GPU_OPT_NO_WATCHDOG)
	std::thread watchdog;
	std::mutex wd_mu;                            // only for wd_cv / wd_stop
	std::condition_variable wd_cv;
	bool wd_started = false, wd_stop = false;
	// bookkeeping
	perseus_gpu_stats stats{};
	int latched = 0;           // first asynchronous error (surfaced at flush/sync/close)
	char latched_msg[512] = "";
};

struct perseus_gpu_plan {
	pg::SegDesc *d_segs = nullptr;
	pg::TileRef *d_tiles = nullptr;            // tiles of segments with 16-byte aligned outputs, then the others
	uint64_t ntiles_stream = 0, ntiles_direct = 0, nsamples = 0, nbytes = 0;
	unsigned fmt = 0;
	int tile_bytes = 0;
};

namespace {

#define CU(h, call)                                                                                               \
	do {                                                                                                          \
		cudaError_t e__ = (call);                                                                                 \
		if (e__ != cudaSuccess) {                                                                                 \
			cudaGetLastError(); /* non-sticky errors must not show up at the next kernel launch check */          \
			return fail(PERSEUS_GPU_CUDAERR, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
		}                                                                                                         \
	} while (0)

void latch(perseus_gpu *h, int code)
{
	if (h && !h->latched) {
		h->latched = code;
		snprintf(h->latched_msg, sizeof(h->latched_msg), "%.*s", (int)sizeof(h->latched_msg) - 1, pg::last_error());
	}
}

int surface_latched(perseus_gpu *h)
{
	if (!h->latched) return 0;
	const int c = h->latched;
	fail(c, "%s", h->latched_msg);
	h->latched = 0;
	return c;
}

pg::Tuning resolve_tuning(const perseus_gpu_tuning *t)
{
	pg::Tuning r{};
	if (t) {
		r.variant = t->variant;
		r.tile_bytes = t->tile_bytes;      // 0 = chosen per output format at launch (kernels.h resolve_geometry)
		r.stages = t->stages;
		r.ctas_per_sm = t->ctas_per_sm;
		r.store_mode = t->store_mode;
	}
	if (r.store_mode == 0) r.store_mode = 1;
	return r;
}

int check_tuning(const pg::Tuning &t)
{
	if (t.variant < 0 || t.variant > 2) return fail(PERSEUS_GPU_ERRPARAM, "tuning.variant %d not in 0..2", t.variant);
	if (t.tile_bytes && !pg::valid_tile(t.tile_bytes))
		return fail(PERSEUS_GPU_ERRPARAM, "tuning.tile_bytes %d must be 6144, 9216, 12288, 18432 or 24576", t.tile_bytes);
	if (t.stages && (t.stages < 2 || t.stages > pg::kMaxStages)) return fail(PERSEUS_GPU_ERRPARAM, "tuning.stages %d not in 2..%d", t.stages, pg::kMaxStages);
	if (t.ctas_per_sm < 0 || t.ctas_per_sm > 8) return fail(PERSEUS_GPU_ERRPARAM, "tuning.ctas_per_sm %d not in 1..8", t.ctas_per_sm);
	if (t.store_mode < 1 || t.store_mode > 2) return fail(PERSEUS_GPU_ERRPARAM, "tuning.store_mode %d not in 0..2", t.store_mode);
	// what actually fits, for either default the unset fields could resolve to: 227 KiB of shared memory, 2048 threads per SM
	for (unsigned fmt : {1u, 3u}) {
		const pg::Geometry g = pg::resolve_geometry(t, fmt);
		if ((size_t)g.stages * g.tile_bytes > 200 * 1024)
			return fail(PERSEUS_GPU_ERRPARAM, "tuning: stages*tile_bytes = %zu exceeds 200 KiB of shared memory", (size_t)g.stages * g.tile_bytes);
		const int by_smem = (int)((227 * 1024) / ((size_t)g.stages * (g.tile_bytes + 16) + 1024));
		const int by_threads = 2048 / (pg::kConsumerThreads + pg::kProducerThreads);
		if (g.ctas_per_sm > by_smem || g.ctas_per_sm > by_threads)
			return fail(PERSEUS_GPU_ERRPARAM, "tuning: %d CTAs/SM do not fit (shared memory allows %d, threads allow %d)", g.ctas_per_sm, by_smem, by_threads);
	}
	return 0;
}

// flags -> format bits; 0 means "whatever output pointers are non-NULL"
int resolve_fmt(unsigned flags, const void *out_i32, const void *out_f32, unsigned *fmt)
{
	unsigned f = flags & (PERSEUS_GPU_OUT_INT32 | PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2);
	if (f == 0) f = (out_i32 ? PERSEUS_GPU_OUT_INT32 : 0u) | (out_f32 ? PERSEUS_GPU_OUT_FLOAT : 0u);
	if ((f & PERSEUS_GPU_OUT_FLOAT) && (f & PERSEUS_GPU_OUT_FLOAT_POW2))
		return fail(PERSEUS_GPU_ERRPARAM, "OUT_FLOAT and OUT_FLOAT_POW2 are mutually exclusive");
	if (f == 0) return fail(PERSEUS_GPU_ERRPARAM, "no output requested");
	if ((f & PERSEUS_GPU_OUT_INT32) && !out_i32) return fail(PERSEUS_GPU_ERRPARAM, "OUT_INT32 requested but out_i32 is NULL");
	if ((f & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2)) && !out_f32)
		return fail(PERSEUS_GPU_ERRPARAM, "float output requested but out_f32 is NULL");
	*fmt = f;
	return 0;
}

enum class Mem { Device, PinnedHost, PageableHost };

Mem classify(const void *p)
{
	cudaPointerAttributes a{};
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
		cudaGetLastError();
		return Mem::PageableHost;
	}
	switch (a.type) {
	case cudaMemoryTypeDevice: return Mem::Device;
	case cudaMemoryTypeManaged: return Mem::Device;
	case cudaMemoryTypeHost: return Mem::PinnedHost;
	default: return Mem::PageableHost;
	}
}

int bind(perseus_gpu *h)
{
	if (!h) return fail(PERSEUS_GPU_NULLHANDLE, "null handle");
	CU(h, cudaSetDevice(h->device));
	return 0;
}

int ensure_bounce(perseus_gpu *h, bool in, bool i32, bool f32)
{
	if (!(in || i32 || f32)) return 0;
	if (!h->pool) {
		h->pool = new (std::nothrow) pg::CopyPool(h->copy_threads - 1);
		if (!h->pool) return fail(PERSEUS_GPU_NOMEM, "out of memory");
	}
	for (int s = 0; s < h->nslots; ++s) {
		if (in && !h->bounce_in[s]) CU(h, cudaHostAlloc(&h->bounce_in[s], h->chunk_bytes, cudaHostAllocDefault));
		if (i32 && !h->bounce_out[s][0]) CU(h, cudaHostAlloc(&h->bounce_out[s][0], h->chunk_bytes / 6 * 8, cudaHostAllocDefault));
		if (f32 && !h->bounce_out[s][1]) CU(h, cudaHostAlloc(&h->bounce_out[s][1], h->chunk_bytes / 6 * 8, cudaHostAllocDefault));
	}
	return 0;
}

int ensure_staging(perseus_gpu *h, bool need_in, bool need_i32, bool need_f32)
{
	for (int s = 0; s < h->nslots; ++s) {
		if (need_in && !h->stage_in[s]) CU(h, cudaMalloc(&h->stage_in[s], h->chunk_bytes));
		if (need_i32 && !h->stage_out[s][0]) CU(h, cudaMalloc(&h->stage_out[s][0], h->chunk_bytes / 6 * 8));
		if (need_f32 && !h->stage_out[s][1]) CU(h, cudaMalloc(&h->stage_out[s][1], h->chunk_bytes / 6 * 8));
		if (!h->ev_in[s]) CU(h, cudaEventCreateWithFlags(&h->ev_in[s], cudaEventDisableTiming));
		if (!h->ev_k[s]) CU(h, cudaEventCreateWithFlags(&h->ev_k[s], cudaEventDisableTiming));
		if (!h->ev_out[s]) CU(h, cudaEventCreateWithFlags(&h->ev_out[s], cudaEventDisableTiming));
	}
	return 0;
}

int do_launch(perseus_gpu *h, const void *in, size_t nbytes, void *o_i32, void *o_f32, unsigned fmt, cudaStream_t st)
{
	int n = 0;
	cudaError_t e = pg::launch_unpack(in, nbytes, o_i32, o_f32, fmt, h->tune, h->sm_count, st, &n);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "unpack kernel launch failed: %s", cudaGetErrorString(e));
	h->stats.kernel_launches += (uint64_t)n;
	h->stats.samples += nbytes / 6;
	h->stats.bytes_in += nbytes / 6 * 6;
	return 0;
}

// Queues the checksum of the piece just unpacked behind its kernel, on the same stream (PERSEUS_GPU_CHECKSUM).
int queue_checksums(perseus_gpu *h, const void *o_i32, const void *o_f32, uint64_t nsamples, uint64_t first_sample, cudaStream_t st)
{
	const void *outs[2] = {o_i32, o_f32};
	for (int k = 0; k < 2; ++k) {
		if (!outs[k] || nsamples == 0) continue;
		cudaError_t e = pg::launch_checksum(outs[k], nsamples * 2, first_sample * 2, h->d_sums + k, h->sm_count, st, /*accumulate=*/true);
		if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "checksum launch failed: %s", cudaGetErrorString(e));
		h->stats.kernel_launches++;
	}
	return 0;
}

// ---- streaming path -------------------------------------------------------------------------
// All functions below this line that take a handle expect the caller to OWN the handle (struct perseus_gpu: the mutex, or the
// callback thread inside its fast path).

int ensure_streaming(perseus_gpu *h)
{
	if (h->streaming_ready) return 0;
	CU(h, cudaSetDevice(h->device));   // first callback on a foreign thread (the reference's libusb poll thread)
	const bool want_i32 = h->stream_fmt & PERSEUS_GPU_OUT_INT32;
	const bool want_f32 = h->stream_fmt & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2);
	for (int k = 0; k < h->nslabs; ++k) {
		Slab &s = h->slabs[k];
		// each guarded, so a call that failed half way (out of memory) can be retried without leaking
		if (!s.host) CU(h, cudaHostAlloc(&s.host, h->slab_bytes, cudaHostAllocDefault));
		if (!s.dev_in) CU(h, cudaMalloc(&s.dev_in, h->slab_bytes));
		if (want_i32 && !s.dev_i32) CU(h, cudaMalloc(&s.dev_i32, h->slab_bytes / 6 * 8));
		if (want_f32 && !s.dev_f32) CU(h, cudaMalloc(&s.dev_f32, h->slab_bytes / 6 * 8));
		// BlockingSync: back-pressure waits happen on the caller of the callback -- in the reference a SCHED_FIFO
		// thread (perseus-sdr.c:749-753) -- and must sleep, not spin
		if (!s.done) CU(h, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming | cudaEventBlockingSync));
		if (!s.unpacked) CU(h, cudaEventCreateWithFlags(&s.unpacked, cudaEventDisableTiming));
		if (!s.ready) CU(h, cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
	}
	h->streaming_ready = true;
	return 0;
}

// Host delivery of one slab, once its outputs have reached pinned host memory; slabs strictly in submission order.  Does what the
// reference's callbacks do with their samples -- fwrite them (perseustest.c:457,499) -- and/or hands them to the application's
// host sink.
void deliver_slab(perseus_gpu *h, Slab *s)
{
	if (h->fout && s->file_bytes && !h->io_error.load(std::memory_order_relaxed)) {
		if (fwrite(s->host_out[s->file_fmt], 1, s->file_bytes, h->fout) != s->file_bytes) h->io_error.store(1, std::memory_order_relaxed);
	}
	if (h->host_sink) {
		const perseus_gpu_host_block b{s->first_sample, s->nsamples, (h->stream_fmt & PERSEUS_GPU_OUT_INT32) ? s->host_out[0] : nullptr,
		                               (h->stream_fmt & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2)) ? s->host_out[1] : nullptr};
		h->host_sink(&b, h->host_sink_extra);
	}
	h->host_blocks.fetch_add(1, std::memory_order_relaxed);
}

void delivery_main(perseus_gpu *h)
{
	cudaSetDevice(h->device);
	std::unique_lock<std::mutex> lk(h->dlv_mu);
	for (;;) {
		if (h->dlv_delivered == h->dlv_submitted) {
			if (h->dlv_stop) return;
			// transfers of a stream usually follow each other closely: look for the next slab a little while before sleeping
			const uint64_t seen = h->dlv_submitted;
			lk.unlock();
			const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(100);
			while (h->dlv_submitted_hint.load(std::memory_order_acquire) == seen && std::chrono::steady_clock::now() < until) sched_yield();
			lk.lock();
			h->dlv_cv.wait(lk, [&] { return h->dlv_stop || h->dlv_delivered != h->dlv_submitted; });
			continue;
		}
		Slab *s = &h->slabs[h->dlv_ring[h->dlv_delivered % kMaxSlabs]];
		lk.unlock();
		// a spinning wait (the event is not BlockingSync): the thread is on the data as soon as the copy engine / the kernel is done
		if (cudaEventSynchronize(s->ready) != cudaSuccess) {
			cudaGetLastError();
			h->io_error.store(2, std::memory_order_relaxed);   // the owner meets the CUDA error itself at its next call; do not deliver garbage
		} else {
			deliver_slab(h, s);
		}
		lk.lock();
		h->dlv_delivered++;
		h->dlv_done_cv.notify_all();
	}
}

int start_delivery(perseus_gpu *h)   // owner only
{
	if (h->dlv_started) return 0;
	h->dlv_started = true;
	try {
		h->dlv_thread = std::thread(delivery_main, h);
	} catch (...) {
		h->dlv_started = false;
		return fail(PERSEUS_GPU_NOMEM, "cannot start the delivery thread");
	}
	return 0;
}

// Hands the slab just submitted to the delivery thread (started with the first one, or by perseus_gpu_prepare).  Owner only.
int queue_delivery(perseus_gpu *h, Slab &s, int index)
{
	int rc = start_delivery(h);
	if (rc) return rc;
	{
		std::lock_guard<std::mutex> lk(h->dlv_mu);
		s.dlv_seq = h->dlv_submitted;
		h->dlv_ring[h->dlv_submitted % kMaxSlabs] = index;
		h->dlv_submitted++;
		h->dlv_submitted_hint.store(h->dlv_submitted, std::memory_order_release);
	}
	h->dlv_cv.notify_one();
	s.to_deliver = true;
	return 0;
}

void stop_delivery(perseus_gpu *h)   // after a flush: nothing is queued
{
	{
		std::lock_guard<std::mutex> lk(h->dlv_mu);
		h->dlv_stop = true;
	}
	h->dlv_cv.notify_all();
	if (h->dlv_thread.joinable()) h->dlv_thread.join();
}

// Retires the `count` oldest slabs in ring order: waits until each one's device work is done and, with a file / host sink,
// until it has been delivered.  Both waits sleep (this may be the reference's SCHED_FIFO thread in back-pressure).
int retire_slabs(perseus_gpu *h, int count)
{
	for (int n = 0; n < count; ++n) {
		Slab &s = h->slabs[h->next_to_write];
		if (s.busy) {
			CU(h, cudaEventSynchronize(s.done));
			if (s.to_deliver) {
				std::unique_lock<std::mutex> lk(h->dlv_mu);
				h->dlv_done_cv.wait(lk, [&] { return h->dlv_delivered > s.dlv_seq; });
				s.to_deliver = false;
			}
			s.busy = false;
		}
		h->next_to_write = (h->next_to_write + 1) % h->nslabs;
	}
	if (h->io_error.exchange(0, std::memory_order_relaxed) == 1) return fail(PERSEUS_GPU_IOERROR, "short write to stream file");
	return 0;
}

int submit_slab(perseus_gpu *h)
{
	Slab &s = h->slabs[h->cur];
	const size_t nbytes = h->fill;
	if (nbytes == 0) return 0;
	CU(h, cudaSetDevice(h->device));   // the only place the callback path needs the device: once per slab, not per transfer
	cudaStream_t st = h->streams[h->cur % h->nstreams];
	const uint64_t ns = nbytes / 6;
	const bool produced[2] = {(h->stream_fmt & PERSEUS_GPU_OUT_INT32) != 0,
	                          (h->stream_fmt & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2)) != 0};
	const bool host_delivery = h->fout || h->host_sink;
	// A small slab (a real receiver delivers 0.6-12 MB/s: slabs are cut by the latency bound, not by their size) is unpacked
	// by ONE launch that reads the pinned slab over the link itself -- the producer's bulk copies take host addresses as they
	// take device ones -- instead of a copy and a launch that waits for it; and when only the host wants the samples the
	// same launch stores them straight into the pinned output.  Large slabs go through HBM: the copy engine moves them
	// without occupying the SMs.
	const bool direct = nbytes <= h->direct_bytes;
	const bool direct_out = direct && host_delivery && !h->sink;
	s.first_sample = h->samples_submitted;
	s.nsamples = ns;
	if (host_delivery)
		for (int k = 0; k < 2; ++k)
			if (produced[k] && !s.host_out[k]) CU(h, cudaHostAlloc(&s.host_out[k], h->slab_bytes / 6 * 8, cudaHostAllocDefault));
#if defined(__SSE2__)
	_mm_sfence();   // the slab was filled with non-temporal stores: make them visible before the device reads it
#endif
	const uint8_t *kin = s.host;
	if (!direct) {
		CU(h, cudaMemcpyAsync(s.dev_in, s.host, nbytes, cudaMemcpyHostToDevice, st));
		kin = s.dev_in;
	}
	h->stats.h2d_bytes += nbytes;
	uint8_t *kout[2] = {produced[0] ? (direct_out ? s.host_out[0] : s.dev_i32) : nullptr,
	                    produced[1] ? (direct_out ? s.host_out[1] : s.dev_f32) : nullptr};
	int rc = do_launch(h, kin, nbytes, kout[0], kout[1], h->stream_fmt, st);
	if (rc) return rc;
	if (h->sink) {
		perseus_gpu_block b{s.first_sample, ns, s.dev_i32, s.dev_f32, (void *)st};
		h->sink(&b, h->sink_extra);
	}
	if (host_delivery) {
		// copy-outs go through the ONE delivery stream (the copy engine takes them in order anyway); the delivery thread takes the
		// slabs in submission order, so blocks reach the host in stream order whichever of the handle's streams unpacked them
		cudaStream_t last = st;
		if (!direct_out) {
			CU(h, cudaEventRecord(s.unpacked, st));
			CU(h, cudaStreamWaitEvent(h->s_dlv, s.unpacked, 0));
			for (int k = 0; k < 2; ++k)
				if (produced[k]) CU(h, cudaMemcpyAsync(s.host_out[k], k ? s.dev_f32 : s.dev_i32, ns * 8, cudaMemcpyDeviceToHost, h->s_dlv));
			last = h->s_dlv;
		}
		h->stats.d2h_bytes += ns * 8 * ((produced[0] ? 1 : 0) + (produced[1] ? 1 : 0));
		s.file_fmt = produced[0] ? 0 : 1;   // a stream file holds one format (perseus_gpu_stream_to_file checks)
		s.file_bytes = h->fout ? ns * 8 : 0;
		CU(h, cudaEventRecord(s.ready, last));
		CU(h, cudaEventRecord(s.done, last));
		rc = queue_delivery(h, s, h->cur);
		if (rc) return rc;
	} else {
		CU(h, cudaEventRecord(s.done, st));
	}
	s.busy = true;
	h->samples_submitted += ns;
	h->stats.slabs++;
	h->fill = 0;
	h->partial_since_ns.store(0, std::memory_order_relaxed);
	h->cur = (h->cur + 1) % h->nslabs;
	// the slab we are about to fill must be free: back-pressure only when the ring is full
	Slab &next = h->slabs[h->cur];
	if (next.busy) {
		if (cudaEventQuery(next.done) == cudaErrorNotReady) h->stats.stalls++;
		cudaGetLastError();
		// everything older than `next` (inclusive) completes in order
		int count = (h->cur - h->next_to_write + h->nslabs) % h->nslabs + 1;
		rc = retire_slabs(h, count);
		if (rc) return rc;
	}
	return 0;
}

// Transfer -> pinned slab.  The slab is written once by this thread and then only read by the device, so the copy uses
// non-temporal stores (pg::copy_nontemporal): no read-for-ownership of the destination lines, about twice the bandwidth of
// memcpy for 6144-byte pieces on one core (the callback is single-threaded by contract, perseus-sdr.c:736-770).
inline void copy_to_slab(uint8_t *dst, const uint8_t *src, size_t n) { pg::copy_nontemporal(dst, src, n); }

inline uint64_t monotonic_ns()
{
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

// The callback reads the time once per transfer (age bound, eager submission).  clock_gettime costs 40-50 ns on these hosts -- a
// fifth of the whole 6144-byte hand-off -- so between anchors CLOCK_MONOTONIC is carried forward by the invariant time-stamp
// counter (RDTSC, ~8 ns).  The rate is measured against CLOCK_MONOTONIC itself (over perseus_gpu_open, then over every anchor
// interval) and the anchor is renewed about once a second, which bounds the disagreement with the watchdog's clock_gettime
// to microseconds.  Anything unexpected (no invariant TSC, a counter that stalls or jumps) falls back to clock_gettime.
#if defined(__x86_64__)
inline uint64_t read_tsc() { return __rdtsc(); }

bool tsc_usable()
{
	static const bool ok = [] {
		const char *off = getenv("PERSEUS_GPU_NO_TSC");
		if (off && *off && *off != '0') return false;
		unsigned a = 0, b = 0, c = 0, d = 0;
		if (!__get_cpuid(0x80000007u, &a, &b, &c, &d)) return false;
		return (d & (1u << 8)) != 0;   // invariant TSC
	}();
	return ok;
}
#else
inline uint64_t read_tsc() { return 0; }
bool tsc_usable() { return false; }
#endif

void tsc_anchor_now(perseus_gpu *h)
{
	const uint64_t ns = monotonic_ns(), c = read_tsc();
	if (h->tsc_anchor && c > h->tsc_anchor && ns > h->ns_anchor) {
		const uint64_t dt = ns - h->ns_anchor, dc = c - h->tsc_anchor;
		if (dt >= 50000 && dt < (1ull << 31)) {                           // 50 us .. 2 s: a rate worth having, no overflow below
			const uint64_t q = (dt << 32) / dc;
			if (q < (1ull << 20) || q > (1ull << 36)) h->use_tsc = false;   // outside 4 THz .. 60 MHz: not a time-stamp counter
			else {
				h->ns_per_tick_q32 = q;
				h->tsc_reanchor = ((1ull << 30) << 32) / q;               // ~1.07 s worth of ticks
			}
		}
	}
	h->ns_anchor = ns;
	h->tsc_anchor = c;
}

inline uint64_t callback_now_ns(perseus_gpu *h)
{
	if (!h->use_tsc || !h->ns_per_tick_q32) return monotonic_ns();
	const uint64_t c = read_tsc();
	const uint64_t dc = c - h->tsc_anchor;
	if (c < h->tsc_anchor || dc > h->tsc_reanchor) {
		tsc_anchor_now(h);
		return h->ns_anchor;
	}
	return h->ns_anchor + (uint64_t)(((unsigned __int128)dc * h->ns_per_tick_q32) >> 32);
}

// Submits the partial slab if its oldest transfer is over age.  Returns 1 if it did, 0 if not, < 0 on error.
int submit_if_over_age(perseus_gpu *h, uint64_t now)
{
	if (!h->max_latency_ns || !h->fill || h->latched) return 0;
	if (now - h->fill_started_ns < h->max_latency_ns) return 0;
	const int rc = submit_slab(h);
	return rc ? rc : 1;
}

// ---- ownership hand-off between the callback thread and everybody else (see struct perseus_gpu) ----------------------

bool membarrier_available()
{
#if defined(__linux__) && defined(__NR_membarrier)
	static const bool ok = [] {
		const char *off = getenv("PERSEUS_GPU_NO_MEMBARRIER");          // tests: force the fence fallback
		if (off && *off && *off != '0') return false;
		const long cmds = syscall(__NR_membarrier, MEMBARRIER_CMD_QUERY, 0, 0);
		if (cmds < 0 || !(cmds & MEMBARRIER_CMD_PRIVATE_EXPEDITED)) return false;
		return syscall(__NR_membarrier, MEMBARRIER_CMD_REGISTER_PRIVATE_EXPEDITED, 0, 0) == 0;
	}();
	return ok;
#else
	return false;
#endif
}

inline void light_barrier(const perseus_gpu *h)      // callback side
{
	if (h->asym) asm volatile("" ::: "memory");
	else std::atomic_thread_fence(std::memory_order_seq_cst);
}

void heavy_barrier(const perseus_gpu *h)             // everybody else
{
#if defined(__linux__) && defined(__NR_membarrier)
	if (h->asym) {
		if (syscall(__NR_membarrier, MEMBARRIER_CMD_PRIVATE_EXPEDITED, 0, 0) != 0) {
			// cannot happen after a successful registration; without the barrier the hand-off would be unsound
			fprintf(stderr, "perseus-gpu: membarrier(PRIVATE_EXPEDITED) failed\n");
			abort();
		}
		return;
	}
#endif
	std::atomic_thread_fence(std::memory_order_seq_cst);
}

// Exclusive ownership for any thread but the callback's fast path.  Also binds the device (bind = true).
struct Entry {
	perseus_gpu *h = nullptr;
	bool locked = false;
	int rc = 0;
	explicit Entry(perseus_gpu *hh, bool bind_device = true)
	{
		if (!hh) { rc = fail(PERSEUS_GPU_NULLHANDLE, "null handle"); return; }
		h = hh;
		const bool inside_fast_path = h->cb_active.load(std::memory_order_relaxed) &&
		                              h->cb_thread.load(std::memory_order_relaxed) == (unsigned long)pthread_self();
		if (!inside_fast_path) {                 // otherwise: a sink called from the callback -- this thread owns the handle already
			h->mu.lock();
			locked = true;
			if (h->lock_depth++ == 0) {
				h->others_want.fetch_add(1, std::memory_order_seq_cst);
				heavy_barrier(h);
				while (h->cb_active.load(std::memory_order_acquire)) sched_yield();   // a callback that was already running finishes first
			}
		}
		if (bind_device) rc = bind(h);
	}
	~Entry()
	{
		if (!locked) return;
		if (--h->lock_depth == 0) h->others_want.fetch_sub(1, std::memory_order_release);
		h->mu.unlock();
	}
	Entry(const Entry &) = delete;
	Entry &operator=(const Entry &) = delete;
};

// The latency bound must hold when no further callback comes (a stalled stream, the tail before
// perseus_stop_async_input): this thread sleeps until the partial slab's deadline and submits it.
void watchdog_main(perseus_gpu *h)
{
	std::unique_lock<std::mutex> lk(h->wd_mu);
	while (!h->wd_stop) {
		const uint64_t since = h->partial_since_ns.load(std::memory_order_relaxed), now = monotonic_ns();
		uint64_t wait_ns = h->max_latency_ns;
		if (since) wait_ns = since + h->max_latency_ns > now ? since + h->max_latency_ns - now : 0;
		if (wait_ns) {
			h->wd_cv.wait_for(lk, std::chrono::nanoseconds(wait_ns));
			continue;                                                   // look again: the slab may have gone out meanwhile
		}
		lk.unlock();
		int rc;
		{
			Entry en(h);                                                // takes the handle away from the callback thread
			rc = en.rc ? en.rc : submit_if_over_age(h, monotonic_ns());
			if (rc < 0) latch(h, rc);
			else if (rc > 0) h->stats.watchdog_submits++;
		}
		lk.lock();
		if (rc <= 0 && !h->wd_stop) h->wd_cv.wait_for(lk, std::chrono::milliseconds(1));   // latched error / raced with a callback: do not spin
	}
}

void start_watchdog(perseus_gpu *h)
{
	if (h->wd_started || !h->max_latency_ns || (h->cfg.options & PERSEUS_GPU_OPT_NO_WATCHDOG)) return;
	h->wd_started = true;   // one attempt only
	try {
		h->watchdog = std::thread(watchdog_main, h);
	} catch (...) {
		// no thread available: the bound is still checked at every callback and by perseus_gpu_poll
	}
}

void stop_watchdog(perseus_gpu *h)   // called without owning the handle
{
	{
		std::lock_guard<std::mutex> lk(h->wd_mu);
		h->wd_stop = true;
	}
	h->wd_cv.notify_all();
	if (h->watchdog.joinable()) h->watchdog.join();
}

int stream_push(perseus_gpu *h, const uint8_t *buf, size_t nbytes)
{
	int rc = ensure_streaming(h);
	if (rc) return rc;
	start_watchdog(h);
	const uint64_t now = h->max_latency_ns ? callback_now_ns(h) : 0;
	// This transfer comes a while after the previous one: the stream is slower than the GPU path -- any real receiver is (a
	// transfer every 0.5 ms at 2 MS/s, every 10.8 ms at 95 kS/s) -- and nothing is gained by letting the transfer wait for
	// company.  It goes out at once (a small slab: one launch, perseus_gpu_config.direct_bytes).  Transfers that arrive back
	// to back (a replayed recording, a burst) keep filling slabs; batching sets in by itself when the path is the bottleneck.
	// Measured start to start with the one clock reading the age bound needs anyway (a second reading per callback costs
	// the 6144-byte path a quarter of its rate); a previous callback that itself took long -- it allocated, or waited for a
	// free slab -- only makes one more small slab.
	const bool eager = h->eager_gap_ns && h->last_push_ns && now - h->last_push_ns > h->eager_gap_ns;
	h->last_push_ns = now;
	if (h->fill == 0) {
		h->fill_started_ns = now;
		h->partial_since_ns.store(now, std::memory_order_relaxed);
	}
	while (nbytes) {
		size_t room = h->slab_bytes - h->fill;
		size_t n = nbytes < room ? nbytes : room;
		copy_to_slab(h->slabs[h->cur].host + h->fill, buf, n);
		h->fill += n;
		buf += n;
		nbytes -= n;
		if (h->fill == h->slab_bytes) {
			rc = submit_slab(h);
			if (rc) return rc;
			h->fill_started_ns = now;
			if (nbytes) h->partial_since_ns.store(now, std::memory_order_relaxed);
		}
	}
	// latency bound: on a real receiver transfers trickle in (10.8 ms apart at 95 kS/s); do not sit on them
	rc = (eager && h->fill && !h->latched) ? submit_slab(h) : submit_if_over_age(h, now);
	return rc < 0 ? rc : 0;
}

// Waits for the streams of the streaming path (the slab streams and the delivery stream) and / or of the bulk host-pointer
// pipeline (streams[0] runs its kernels, s_in / s_out its copies); CUDA errors are latched.
void wait_streams(perseus_gpu *h, bool streaming, bool bulk, bool launch_stream_done = false)
{
	cudaStream_t all[kMaxStreams + 3];
	int n = 0;
	for (int s = 0; s < h->nstreams; ++s)
		if (streaming || (bulk && s == 0 && !launch_stream_done)) all[n++] = h->streams[s];
	if (bulk) {
		all[n++] = h->s_in;
		all[n++] = h->s_out;
	}
	if (streaming) all[n++] = h->s_dlv;
	for (int s = 0; s < n; ++s) {
		if (!all[s]) continue;
		cudaError_t e = cudaStreamSynchronize(all[s]);
		if (e != cudaSuccess) {
			fail(PERSEUS_GPU_CUDAERR, "stream synchronisation: %s", cudaGetErrorString(e));
			latch(h, PERSEUS_GPU_CUDAERR);
		}
	}
}

int sync_locked(perseus_gpu *h)
{
	wait_streams(h, true, true);
	return surface_latched(h);
}

int flush_locked(perseus_gpu *h)
{
	int rc;
	if (h->streaming_ready) {
		if (h->fill && !h->latched) {
			rc = submit_slab(h);
			if (rc) latch(h, rc);
		}
		// The slab events are BlockingSync (a back-pressure wait must sleep, see ensure_streaming), and a sleeping wait
		// costs ~0.2 ms of wake-up latency.  flush is called by the application and wants the result now: spin on the
		// streams first, after which every slab event is already complete and retiring never sleeps on the device.
		wait_streams(h, true, false);
		rc = retire_slabs(h, h->nslabs);   // ... and every block has been written / handed to the host sink
		if (rc) latch(h, rc);
		h->next_to_write = h->cur;       // nothing in flight: the next slab submitted is the oldest
		if (h->fout) fflush(h->fout);
		wait_streams(h, false, true, true);   // whatever the bulk pipeline still has queued (streams[0] was waited for above)
		return surface_latched(h);
	}
	return sync_locked(h);
}

void destroy_plan(perseus_gpu_plan *p)
{
	if (p->d_segs) cudaFree(p->d_segs);
	if (p->d_tiles) cudaFree(p->d_tiles);
	delete p;
}

int plan_create_locked(perseus_gpu *h, const perseus_gpu_seg *segs, int nseg, unsigned flags, perseus_gpu_plan **out)
{
	if (!out) return fail(PERSEUS_GPU_ERRPARAM, "null plan pointer");
	*out = nullptr;
	if (nseg < 0 || (nseg > 0 && !segs)) return fail(PERSEUS_GPU_ERRPARAM, "bad segment table");
	unsigned fmt = flags & (PERSEUS_GPU_OUT_INT32 | PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2);
	if (fmt == 0 && nseg > 0) fmt = (segs[0].out_i32 ? PERSEUS_GPU_OUT_INT32 : 0u) | (segs[0].out_f32 ? PERSEUS_GPU_OUT_FLOAT : 0u);
	if ((fmt & PERSEUS_GPU_OUT_FLOAT) && (fmt & PERSEUS_GPU_OUT_FLOAT_POW2))
		return fail(PERSEUS_GPU_ERRPARAM, "OUT_FLOAT and OUT_FLOAT_POW2 are mutually exclusive");
	if (fmt == 0 && nseg > 0) return fail(PERSEUS_GPU_ERRPARAM, "no output requested");

	const int tile = pg::resolve_geometry(h->tune, fmt).tile_bytes;
	std::vector<pg::SegDesc> hs((size_t)nseg);
	// Every segment takes the bulk-copy pipeline with 128-bit stores; outputs that are not 16-byte aligned (an {I,Q}
	// array at its natural 8-byte alignment, or any multiple of 4) get them after a pre-roll of 1..3 output words.  Per
	// segment, so one odd receiver does not slow the batch down.  Wire pointers may have any alignment.  (`slow`: the register-only
	// kernel, only when the handle is tuned to PERSEUS_GPU_VARIANT_DIRECT.)
	std::vector<pg::TileRef> fast, slow;
	uint64_t nsamples = 0, nbytes = 0;
	for (int i = 0; i < nseg; ++i) {
		const perseus_gpu_seg &s = segs[i];
		const uint64_t used = (uint64_t)s.nbytes / 6 * 6;
		if (used && !s.in) return fail(PERSEUS_GPU_ERRPARAM, "segment %d: in is NULL", i);
		if (used && (fmt & PERSEUS_GPU_OUT_INT32) && !s.out_i32) return fail(PERSEUS_GPU_ERRPARAM, "segment %d: out_i32 is NULL", i);
		if (used && (fmt & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2)) && !s.out_f32)
			return fail(PERSEUS_GPU_ERRPARAM, "segment %d: out_f32 is NULL", i);
		if (((uintptr_t)s.out_i32 & 3) || ((uintptr_t)s.out_f32 & 3)) return fail(PERSEUS_GPU_ERRPARAM, "segment %d: outputs must be 4-byte aligned", i);
		pg::SegDesc &d = hs[(size_t)i];
		d = pg::SegDesc{static_cast<const uint8_t *>(s.in), used, (fmt & PERSEUS_GPU_OUT_INT32) ? s.out_i32 : nullptr,
		                (fmt & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2)) ? s.out_f32 : nullptr, 0u, 0u};
		const bool direct = h->tune.variant == PERSEUS_GPU_VARIANT_DIRECT;
		const int pre = direct ? 0 : pg::stream_preroll(d.out_i32, d.out_f32);
		if (pre < 0) d.word_stores = 1u;       // the two outputs at different phases: same kernel, 32-bit stores for this segment's tiles
		uint64_t span = used;                  // wire bytes the segment's tiles cover
		if (pre > 0 && used) {                 // outputs 4m bytes past a 16-byte boundary: the segment starts m words early (kernels.h)
			d.in -= pre;
			d.nbytes = used + (uint64_t)pre;
			if (d.out_i32) d.out_i32 = static_cast<uint8_t *>(d.out_i32) - pre / 3 * 4;
			if (d.out_f32) d.out_f32 = static_cast<uint8_t *>(d.out_f32) - pre / 3 * 4;
			d.preroll = (uint32_t)pre;
			span = d.nbytes;
		}
		const uint64_t nt = (span + (uint64_t)tile - 1) / (uint64_t)tile;
		if (nt > 0xFFFFFFFFull) return fail(PERSEUS_GPU_BUFFERSIZE, "segment %d too large", i);
		std::vector<pg::TileRef> &dst = direct ? slow : fast;
		for (uint64_t t = 0; t < nt; ++t) dst.push_back(pg::TileRef{(uint32_t)i, (uint32_t)t});
		nsamples += used / 6;
		nbytes += used;
	}
	perseus_gpu_plan *p = new (std::nothrow) perseus_gpu_plan();
	if (!p) return fail(PERSEUS_GPU_NOMEM, "out of memory");
	p->ntiles_stream = fast.size();
	p->ntiles_direct = slow.size();
	p->nsamples = nsamples;
	p->nbytes = nbytes;
	p->fmt = fmt;
	p->tile_bytes = tile;
	fast.insert(fast.end(), slow.begin(), slow.end());   // one upload: [stream tiles | direct tiles]
	cudaError_t e = cudaSuccess;
	if (nseg) e = cudaMalloc(&p->d_segs, hs.size() * sizeof(pg::SegDesc));
	if (e == cudaSuccess && !fast.empty()) e = cudaMalloc(&p->d_tiles, fast.size() * sizeof(pg::TileRef));
	if (e == cudaSuccess && nseg) e = cudaMemcpyAsync(p->d_segs, hs.data(), hs.size() * sizeof(pg::SegDesc), cudaMemcpyHostToDevice, h->streams[0]);
	if (e == cudaSuccess && !fast.empty())
		e = cudaMemcpyAsync(p->d_tiles, fast.data(), fast.size() * sizeof(pg::TileRef), cudaMemcpyHostToDevice, h->streams[0]);
	if (e == cudaSuccess) e = cudaStreamSynchronize(h->streams[0]);   // hs/fast die with this frame
	if (e != cudaSuccess) {
		cudaGetLastError();
		destroy_plan(p);
		return fail(PERSEUS_GPU_CUDAERR, "plan upload failed: %s", cudaGetErrorString(e));
	}
	*out = p;
	return 0;
}

int64_t plan_run_locked(perseus_gpu *h, perseus_gpu_plan *p, unsigned flags)
{
	if (!p) return fail(PERSEUS_GPU_ERRPARAM, "null plan");
	int n = 0;   // the plan keeps the tile size and the split it was built with; stages / CTAs per SM follow the handle's current tuning
	cudaError_t e = pg::launch_unpack_batch(p->d_segs, p->d_tiles, p->ntiles_stream, p->d_tiles + p->ntiles_stream, p->ntiles_direct,
	                                        p->tile_bytes, p->fmt, h->tune, h->sm_count, h->streams[0], &n);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "batched unpack launch failed: %s", cudaGetErrorString(e));
	h->stats.kernel_launches += (uint64_t)n;
	h->stats.samples += p->nsamples;
	h->stats.bytes_in += p->nbytes;
	if (!(flags & PERSEUS_GPU_ASYNC)) {
		int rc = sync_locked(h);
		if (rc) return rc;
	}
	return (int64_t)p->nsamples;
}

}  // namespace

// =============================================================================== C ABI

extern "C" {

#define PG_STR2(x) #x
#define PG_STR(x) PG_STR2(x)
const char *perseus_gpu_version(void) { return "perseus-gpu abi " PG_STR(PERSEUS_GPU_ABI_VERSION) ", sm_100a, " __DATE__; }

int perseus_gpu_device_count(void)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(PERSEUS_GPU_NODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
	}
	return n;
}

int perseus_gpu_device_info(int device, char *name, size_t name_len, int *sm_count, int *cc_major, int *cc_minor, uint64_t *total_mem)
{
	cudaDeviceProp p{};
	cudaError_t e = cudaGetDeviceProperties(&p, device);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(PERSEUS_GPU_NODEVICE, "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
	}
	if (name && name_len) snprintf(name, name_len, "%s", p.name);
	if (sm_count) *sm_count = p.multiProcessorCount;
	if (cc_major) *cc_major = p.major;
	if (cc_minor) *cc_minor = p.minor;
	if (total_mem) *total_mem = p.totalGlobalMem;
	return 0;
}

int perseus_gpu_open(perseus_gpu **out, const perseus_gpu_config *ucfg)
{
	if (!out) return fail(PERSEUS_GPU_ERRPARAM, "null handle pointer");
	*out = nullptr;
	perseus_gpu_config cfg{};
	if (ucfg) {
		if (ucfg->struct_size < 8 || ucfg->struct_size > sizeof(cfg))
			return fail(PERSEUS_GPU_ERRPARAM, "perseus_gpu_config.struct_size %u not understood (this library: %zu)", ucfg->struct_size, sizeof(cfg));
		memcpy(&cfg, ucfg, ucfg->struct_size);
	}
	int ndev = perseus_gpu_device_count();
	if (ndev <= 0) return ndev < 0 ? ndev : fail(PERSEUS_GPU_NODEVICE, "no CUDA device");
	if (cfg.device < 0 || cfg.device >= ndev) return fail(PERSEUS_GPU_ERRPARAM, "device %d out of range (0..%d)", cfg.device, ndev - 1);
	cudaDeviceProp prop{};
	CU(nullptr, cudaGetDeviceProperties(&prop, cfg.device));
	if (prop.major != 10)
		return fail(PERSEUS_GPU_BADARCH, "device %d (%s) is sm_%d%d; this library carries sm_100a code only and has no fallback", cfg.device, prop.name,
		            prop.major, prop.minor);

	pg::Tuning tune = resolve_tuning(&cfg.tuning);
	int rc = check_tuning(tune);
	if (rc) return rc;
	unsigned sfmt = cfg.stream_flags ? cfg.stream_flags : PERSEUS_GPU_OUT_INT32;
	if (sfmt & ~(PERSEUS_GPU_OUT_INT32 | PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2))
		return fail(PERSEUS_GPU_ERRPARAM, "stream_flags 0x%x has unknown bits", sfmt);
	if ((sfmt & PERSEUS_GPU_OUT_FLOAT) && (sfmt & PERSEUS_GPU_OUT_FLOAT_POW2))
		return fail(PERSEUS_GPU_ERRPARAM, "stream_flags: OUT_FLOAT and OUT_FLOAT_POW2 are mutually exclusive");
	if (cfg.options & ~PERSEUS_GPU_OPT_NO_WATCHDOG) return fail(PERSEUS_GPU_ERRPARAM, "options 0x%x has unknown bits", cfg.options);
	const uint32_t nslabs = cfg.nslabs ? cfg.nslabs : 4;
	if (nslabs < 2 || nslabs > (uint32_t)kMaxSlabs) return fail(PERSEUS_GPU_ERRPARAM, "nslabs %u not in 2..%d", nslabs, kMaxSlabs);
	const uint32_t nstreams = cfg.nstreams ? cfg.nstreams : 2;
	if (nstreams < 1 || nstreams > (uint32_t)kMaxStreams) return fail(PERSEUS_GPU_ERRPARAM, "nstreams %u not in 1..%d", nstreams, kMaxStreams);
	const uint32_t nslots = cfg.stage_slots ? cfg.stage_slots : 3;
	if (nslots < 2 || nslots > (uint32_t)kMaxStageSlots) return fail(PERSEUS_GPU_ERRPARAM, "stage_slots %u not in 2..%d", nslots, kMaxStageSlots);
	uint64_t slab = cfg.slab_bytes ? cfg.slab_bytes : (8ull << 20);
	slab -= slab % 48;
	uint64_t chunk = cfg.chunk_bytes ? cfg.chunk_bytes : (32ull << 20);
	// Whole pages on both sides of the link: 12288 wire bytes (3 pages) become 16384 output bytes (4 pages), so every staged
	// copy starts and ends on a page boundary of the caller's pinned buffers.  The copy engines need that: chunks that are
	// only 128-byte aligned reach 52.5-53.5 GB/s device->host, page-aligned ones 56.1-56.6 (tools/d2h_probe.py).
	chunk -= chunk % (chunk >= 12288 ? 12288 : 48);
	if (slab < 48 || chunk < 48) return fail(PERSEUS_GPU_BUFFERSIZE, "slab_bytes/chunk_bytes must be at least 48");

	perseus_gpu *h = new (std::nothrow) perseus_gpu();
	if (!h) return fail(PERSEUS_GPU_NOMEM, "out of memory");
	h->device = cfg.device;
	h->sm_count = prop.multiProcessorCount;
	h->cfg = cfg;
	h->tune = tune;
	h->stream_fmt = sfmt;
	h->nslabs = (int)nslabs;
	h->nstreams = (int)nstreams;
	h->nslots = (int)nslots;
	h->slab_bytes = (size_t)slab;
	h->max_latency_ns = cfg.max_latency_us == 0xFFFFFFFFu ? 0 : (uint64_t)(cfg.max_latency_us ? cfg.max_latency_us : 50000u) * 1000ull;
	h->chunk_bytes = (size_t)chunk;
	// only with a latency bound: max_latency_us = 0xFFFFFFFF asks for full slabs and nothing else
	h->eager_gap_ns = (!h->max_latency_ns || cfg.eager_gap_us == 0xFFFFFFFFu) ? 0 : (uint64_t)(cfg.eager_gap_us ? cfg.eager_gap_us : 100u) * 1000ull;
	{
		const unsigned hw = std::thread::hardware_concurrency();
		const unsigned autot = hw >= 4 ? (hw / 2 > 8 ? 8 : hw / 2) : 1;
		h->copy_threads = cfg.copy_threads == 0xFFFFFFFFu ? 0 : (int)(cfg.copy_threads ? (cfg.copy_threads > 64 ? 64 : cfg.copy_threads) : autot);
	}
	h->direct_bytes = cfg.direct_bytes == 0xFFFFFFFFu ? 0 : cfg.direct_bytes ? (size_t)cfg.direct_bytes : kDefaultDirectBytes;
	h->asym = membarrier_available();
	h->use_tsc = tsc_usable();
	if (h->use_tsc) tsc_anchor_now(h);   // first reading; the rate comes from the second one, at the end of this function
	auto bail = [&](int code) {   // free what exists, keep the message of the original failure
		const std::string keep = pg::last_error();
		perseus_gpu_close(h);
		pg::set_last_error(keep.c_str());
		return code;
	};
	cudaError_t e = cudaSetDevice(h->device);
	if (e != cudaSuccess) return bail(fail(PERSEUS_GPU_CUDAERR, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e)));
	for (int s = 0; s < h->nstreams; ++s) {
		e = cudaStreamCreateWithFlags(&h->streams[s], cudaStreamNonBlocking);
		if (e != cudaSuccess) return bail(fail(PERSEUS_GPU_CUDAERR, "cudaStreamCreate: %s", cudaGetErrorString(e)));
	}
	e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_dlv, cudaStreamNonBlocking);
	if (e != cudaSuccess) return bail(fail(PERSEUS_GPU_CUDAERR, "cudaStreamCreate: %s", cudaGetErrorString(e)));
	for (int k = 0; k < kEventSlots && e == cudaSuccess; ++k) e = cudaEventCreate(&h->events[k]);
	for (int k = 0; k < 4 && e == cudaSuccess; ++k) e = cudaEventCreate(&h->tev[k]);
	if (e != cudaSuccess) return bail(fail(PERSEUS_GPU_CUDAERR, "cudaEventCreate: %s", cudaGetErrorString(e)));
	e = cudaMalloc(&h->d_scratch, 2 * sizeof(unsigned long long));
	if (e == cudaSuccess) e = cudaMalloc(&h->d_sums, 2 * sizeof(unsigned long long));
	if (e == cudaSuccess) e = cudaMemset(h->d_sums, 0, 2 * sizeof(unsigned long long));
	if (e == cudaSuccess) e = cudaHostAlloc(&h->h_scratch, 2 * sizeof(unsigned long long), cudaHostAllocDefault);
	if (e != cudaSuccess) return bail(fail(PERSEUS_GPU_CUDAERR, "scratch allocation: %s", cudaGetErrorString(e)));
	if (h->use_tsc) {
		while (monotonic_ns() - h->ns_anchor < 60000) { }   // streams and events took milliseconds; make sure of 60 us anyway
		tsc_anchor_now(h);
		if (!h->ns_per_tick_q32) h->use_tsc = false;
	}
	*out = h;
	return ok();
}

int perseus_gpu_close(perseus_gpu *h)
{
	if (!h) return fail(PERSEUS_GPU_NULLHANDLE, "null handle");
	stop_watchdog(h);
	int rc = 0;
	{
		Entry en(h, false);
		if (cudaSetDevice(h->device) == cudaSuccess) {
			rc = flush_locked(h);
			stop_delivery(h);
			for (int k = 0; k < kMaxSlabs; ++k) {
				Slab &s = h->slabs[k];
				if (s.host) cudaFreeHost(s.host);
				for (uint8_t *p : s.host_out)
					if (p) cudaFreeHost(p);
				if (s.dev_in) cudaFree(s.dev_in);
				if (s.dev_i32) cudaFree(s.dev_i32);
				if (s.dev_f32) cudaFree(s.dev_f32);
				if (s.done) cudaEventDestroy(s.done);
				if (s.unpacked) cudaEventDestroy(s.unpacked);
				if (s.ready) cudaEventDestroy(s.ready);
			}
			for (int s = 0; s < kMaxStageSlots; ++s) {
				if (h->stage_in[s]) cudaFree(h->stage_in[s]);
				if (h->stage_out[s][0]) cudaFree(h->stage_out[s][0]);
				if (h->stage_out[s][1]) cudaFree(h->stage_out[s][1]);
				if (h->ev_in[s]) cudaEventDestroy(h->ev_in[s]);
				if (h->ev_k[s]) cudaEventDestroy(h->ev_k[s]);
				if (h->ev_out[s]) cudaEventDestroy(h->ev_out[s]);
				if (h->bounce_in[s]) cudaFreeHost(h->bounce_in[s]);
				if (h->bounce_out[s][0]) cudaFreeHost(h->bounce_out[s][0]);
				if (h->bounce_out[s][1]) cudaFreeHost(h->bounce_out[s][1]);
			}
			if (h->d_scratch) cudaFree(h->d_scratch);
			if (h->d_sums) cudaFree(h->d_sums);
			if (h->h_scratch) cudaFreeHost(h->h_scratch);
			for (int k = 0; k < kEventSlots; ++k)
				if (h->events[k]) cudaEventDestroy(h->events[k]);
			for (int k = 0; k < 4; ++k)
				if (h->tev[k]) cudaEventDestroy(h->tev[k]);
			for (int s = 0; s < h->nstreams; ++s)
				if (h->streams[s]) cudaStreamDestroy(h->streams[s]);
			if (h->s_in) cudaStreamDestroy(h->s_in);
			if (h->s_out) cudaStreamDestroy(h->s_out);
			if (h->s_dlv) cudaStreamDestroy(h->s_dlv);
			cudaGetLastError();
		}
		stop_delivery(h);   // also when the device could not be bound above: the thread must be gone before the handle is
		if (h->fout) {
			if ((h->fout_is_stdout ? fflush(h->fout) : fclose(h->fout)) != 0 && !rc) rc = fail(PERSEUS_GPU_IOERROR, "closing stream file failed");
		}
	}
	delete h->pool;   // joins the helper threads
	delete h;
	return rc;
}

int perseus_gpu_sync(perseus_gpu *h)
{
	Entry en(h);
	if (en.rc) return en.rc;
	return sync_locked(h);
}

int64_t perseus_gpu_unpack(perseus_gpu *h, const void *buf, size_t nbytes, void *out_i32, void *out_f32, unsigned flags)
{
	Entry en(h);
	int rc = en.rc;
	if (rc) return rc;
	if (flags & ~(PERSEUS_GPU_OUT_INT32 | PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2 | PERSEUS_GPU_ASYNC | PERSEUS_GPU_CHECKSUM))
		return fail(PERSEUS_GPU_ERRPARAM, "unknown flag bits 0x%x", flags);
	const bool want_sums = flags & PERSEUS_GPU_CHECKSUM;
	unsigned fmt = 0;
	rc = resolve_fmt(flags, out_i32, out_f32, &fmt);
	if (rc) return rc;
	if (!(fmt & PERSEUS_GPU_OUT_INT32)) out_i32 = nullptr;
	if (!(fmt & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2))) out_f32 = nullptr;
	const uint64_t ns = nbytes / 6;
	if (ns && !buf) return fail(PERSEUS_GPU_ERRPARAM, "buf is NULL");
	if ((out_i32 && ((uintptr_t)out_i32 & 3)) || (out_f32 && ((uintptr_t)out_f32 & 3)))
		return fail(PERSEUS_GPU_ERRPARAM, "output pointers must be 4-byte aligned");

	if (want_sums) {   // totals restart with this call (even an empty one); everything queued earlier has to be done with them first
		rc = sync_locked(h);
		if (rc) return rc;
		CU(h, cudaMemsetAsync(h->d_sums, 0, 2 * sizeof(unsigned long long), h->streams[0]));
		CU(h, cudaStreamSynchronize(h->streams[0]));
	}
	if (ns == 0) return 0;
	const Mem min_ = classify(buf);
	const Mem mi = out_i32 ? classify(out_i32) : Mem::Device;
	const Mem mf = out_f32 ? classify(out_f32) : Mem::Device;
	const bool in_dev = min_ == Mem::Device, oi_dev = mi == Mem::Device, of_dev = mf == Mem::Device;
	cudaStream_t sk = h->streams[0];

	if (in_dev && oi_dev && of_dev) {
		rc = do_launch(h, buf, ns * 6, out_i32, out_f32, fmt, sk);
		if (rc) return rc;
		if (want_sums && (rc = queue_checksums(h, out_i32, out_f32, ns, 0, sk))) return rc;
	} else {
		// Three-stage pipeline over `nslots` staging slots: copy-in on s_in, kernel (+ checksums) on streams[0], copy-out
		// on s_out, ordered per slot by events.  The copy engines of both directions and the SMs each work on a
		// different chunk at the same time; neither copy stream ever waits behind a copy of the other direction.
		// PAGEABLE host buffers get two more stages at the ends: the caller and the handle's helper threads move each chunk
		// between the application's memory and a pinned bounce buffer of the slot while the engines work on the neighbours.
		const bool host_out = (out_i32 && !oi_dev) || (out_f32 && !of_dev);
		const bool in_page = min_ == Mem::PageableHost && h->copy_threads > 0;
		const bool oi_page = out_i32 && mi == Mem::PageableHost && h->copy_threads > 0;
		const bool of_page = out_f32 && mf == Mem::PageableHost && h->copy_threads > 0;
		rc = ensure_staging(h, !in_dev, out_i32 && !oi_dev, out_f32 && !of_dev);
		if (!rc) rc = ensure_bounce(h, in_page, oi_page, of_page);
		if (rc) return rc;
		const uint8_t *src = static_cast<const uint8_t *>(buf);
		const size_t total = ns * 6;
		struct { bool any; size_t o, on; } pend[kMaxStageSlots] = {};   // outputs waiting in a slot's bounce buffers
		// bounce_out[s] -> the application's memory, once the copy-out that fills it has finished
		auto drain = [&](int s) -> int {
			if (!pend[s].any) return 0;
			CU(h, cudaEventSynchronize(h->ev_out[s]));
			if (oi_page) h->pool->copy(static_cast<uint8_t *>(out_i32) + pend[s].o, h->bounce_out[s][0], pend[s].on, false);
			if (of_page) h->pool->copy(static_cast<uint8_t *>(out_f32) + pend[s].o, h->bounce_out[s][1], pend[s].on, false);
			pend[s].any = false;
			return 0;
		};
		for (size_t off = 0; off < total;) {
			const int s = (int)(h->stage_seq++ % (uint64_t)h->nslots);
			const size_t n = total - off < h->chunk_bytes ? total - off : h->chunk_bytes;
			const size_t o = off / 6 * 8, on = n / 6 * 8;
			if ((rc = drain(s))) return rc;                                   // the chunk that used this slot nslots chunks ago
			const uint8_t *kin = src + off;
			if (!in_dev) {
				const uint8_t *hsrc = src + off;
				if (in_page) {
					CU(h, cudaEventSynchronize(h->ev_in[s]));                 // the copy-in that last read this bounce buffer is done
					h->pool->copy(h->bounce_in[s], hsrc, n, true);
					hsrc = h->bounce_in[s];
				}
				CU(h, cudaStreamWaitEvent(h->s_in, h->ev_k[s], 0));       // the kernel that last read this slot's input is done
				CU(h, cudaMemcpyAsync(h->stage_in[s], hsrc, n, cudaMemcpyHostToDevice, h->s_in));
				CU(h, cudaEventRecord(h->ev_in[s], h->s_in));
				CU(h, cudaStreamWaitEvent(sk, h->ev_in[s], 0));
				h->stats.h2d_bytes += n;
				kin = h->stage_in[s];
			}
			if (host_out) CU(h, cudaStreamWaitEvent(sk, h->ev_out[s], 0));   // the copy-out that last read this slot's outputs is done
			uint8_t *ki = out_i32 ? (oi_dev ? static_cast<uint8_t *>(out_i32) + o : h->stage_out[s][0]) : nullptr;
			uint8_t *kf = out_f32 ? (of_dev ? static_cast<uint8_t *>(out_f32) + o : h->stage_out[s][1]) : nullptr;
			rc = do_launch(h, kin, n, ki, kf, fmt, sk);
			if (rc) return rc;
			if (want_sums && (rc = queue_checksums(h, ki, kf, n / 6, off / 6, sk))) return rc;
			CU(h, cudaEventRecord(h->ev_k[s], sk));
			if (host_out) {
				CU(h, cudaStreamWaitEvent(h->s_out, h->ev_k[s], 0));
				if (out_i32 && !oi_dev) {
					CU(h, cudaMemcpyAsync(oi_page ? h->bounce_out[s][0] : static_cast<uint8_t *>(out_i32) + o, ki, on, cudaMemcpyDeviceToHost, h->s_out));
					h->stats.d2h_bytes += on;
				}
				if (out_f32 && !of_dev) {
					CU(h, cudaMemcpyAsync(of_page ? h->bounce_out[s][1] : static_cast<uint8_t *>(out_f32) + o, kf, on, cudaMemcpyDeviceToHost, h->s_out));
					h->stats.d2h_bytes += on;
				}
				CU(h, cudaEventRecord(h->ev_out[s], h->s_out));
				if (oi_page || of_page) { pend[s].any = true; pend[s].o = o; pend[s].on = on; }
			}
			off += n;
		}
		// pageable outputs are complete when the call returns, PERSEUS_GPU_ASYNC or not (like the CUDA runtime's own pageable copies)
		for (int k = 0; k < h->nslots; ++k)
			if ((rc = drain((int)((h->stage_seq + (uint64_t)k) % (uint64_t)h->nslots)))) return rc;   // oldest chunk first
	}
	if (!(flags & PERSEUS_GPU_ASYNC)) {
		rc = sync_locked(h);
		if (rc) return rc;
	}
	return (int64_t)ns;
}

int perseus_gpu_get_checksums(perseus_gpu *h, uint64_t *sum_i32, uint64_t *sum_f32)
{
	Entry en(h);
	if (en.rc) return en.rc;
	int rc = sync_locked(h);
	if (rc) return rc;
	CU(h, cudaMemcpyAsync(h->h_scratch, h->d_sums, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->streams[0]));
	CU(h, cudaStreamSynchronize(h->streams[0]));
	h->stats.d2h_bytes += 2 * sizeof(unsigned long long);
	if (sum_i32) *sum_i32 = h->h_scratch[0];
	if (sum_f32) *sum_f32 = h->h_scratch[1];
	return 0;
}

// ---- batched ------------------------------------------------------------------------------------

int perseus_gpu_plan_create(perseus_gpu *h, const perseus_gpu_seg *segs, int nseg, unsigned flags, perseus_gpu_plan **out)
{
	Entry en(h);
	if (en.rc) return en.rc;
	return plan_create_locked(h, segs, nseg, flags, out);
}

int64_t perseus_gpu_plan_run(perseus_gpu *h, perseus_gpu_plan *p, unsigned flags)
{
	Entry en(h);
	if (en.rc) return en.rc;
	return plan_run_locked(h, p, flags);
}

int perseus_gpu_plan_destroy(perseus_gpu *h, perseus_gpu_plan *p)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (!p) return 0;
	cudaStreamSynchronize(h->streams[0]);
	destroy_plan(p);
	return 0;
}

int64_t perseus_gpu_unpack_batch(perseus_gpu *h, const perseus_gpu_seg *segs, int nseg, unsigned flags)
{
	Entry en(h);
	if (en.rc) return en.rc;
	perseus_gpu_plan *p = nullptr;
	int rc = plan_create_locked(h, segs, nseg, flags, &p);
	if (rc) return rc;
	int64_t n = plan_run_locked(h, p, 0);   // the plan is freed below, so always synchronous
	cudaStreamSynchronize(h->streams[0]);
	destroy_plan(p);
	return n;
}

// ---- streaming hand-off -----------------------------------------------------------------------------

namespace {
inline void callback_body(perseus_gpu *h, const uint8_t *buf, size_t n)
{
	if (h->latched) {                              // a previous failure is waiting to be reported: count what is lost
		h->stats.dropped_callbacks++;
		h->stats.dropped_bytes += n;
		return;
	}
	h->stats.callbacks++;
	int rc = stream_push(h, buf, n);
	if (rc) latch(h, rc);
}
}  // namespace

int perseus_gpu_input_callback(void *buf, int buf_size, void *extra)
{
	perseus_gpu *h = static_cast<perseus_gpu *>(extra);
	if (!h || !buf || buf_size < 6) return 0;
	// perseustest.c:443 — only whole samples of THIS transfer count
	const size_t n = (size_t)buf_size / 6 * 6;
	// fast path: announce, then look whether anybody else holds or wants the handle (no locked instruction, see struct perseus_gpu)
	h->cb_thread.store((unsigned long)pthread_self(), std::memory_order_relaxed);
	h->cb_active.store(1, std::memory_order_relaxed);
	light_barrier(h);
	if (h->others_want.load(std::memory_order_acquire) != 0) {
		h->cb_active.store(0, std::memory_order_release);
		Entry en(h, false);                         // slow path: queue up behind them like any other thread
		callback_body(h, static_cast<const uint8_t *>(buf), n);
		return 0;
	}
	callback_body(h, static_cast<const uint8_t *>(buf), n);
#if defined(__SSE2__)
	// somebody arrived meanwhile and will take over when cb_active drops: the slab bytes written with non-temporal stores must
	// be globally visible before that (a thread that arrives later than this check drains them with its membarrier)
	if (h->others_want.load(std::memory_order_relaxed) != 0) _mm_sfence();
#endif
	h->cb_active.store(0, std::memory_order_release);
	return 0;
}

int perseus_gpu_prepare(perseus_gpu *h)
{
	Entry en(h);
	if (en.rc) return en.rc;
	int rc = ensure_streaming(h);
	if (rc) return rc;
	if (h->fout || h->host_sink) {
		const bool produced[2] = {(h->stream_fmt & PERSEUS_GPU_OUT_INT32) != 0,
		                          (h->stream_fmt & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2)) != 0};
		for (int k = 0; k < h->nslabs; ++k)
			for (int f = 0; f < 2; ++f)
				if (produced[f] && !h->slabs[k].host_out[f]) CU(h, cudaHostAlloc(&h->slabs[k].host_out[f], h->slab_bytes / 6 * 8, cudaHostAllocDefault));
		if ((rc = start_delivery(h))) return rc;
	}
	// the first launch of a kernel loads its code: do that here, on two samples of slab 0 (nothing is delivered or counted)
	Slab &s = h->slabs[0];
	memset(s.host, 0, 12);
	int n = 0;
	cudaError_t e = pg::launch_unpack(s.host, 12, s.dev_i32, s.dev_f32, h->stream_fmt, h->tune, h->sm_count, h->streams[0], &n);
	if (e == cudaSuccess) e = cudaStreamSynchronize(h->streams[0]);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "warm-up launch failed: %s", cudaGetErrorString(e));
	h->stats.kernel_launches += (uint64_t)n;
	start_watchdog(h);
	return 0;
}

int perseus_gpu_poll(perseus_gpu *h)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (!h->streaming_ready) return 0;
	const int rc = submit_if_over_age(h, monotonic_ns());
	if (rc > 0) h->stats.watchdog_submits++;
	return rc;
}

int perseus_gpu_set_sink(perseus_gpu *h, perseus_gpu_sink sink, void *extra)
{
	Entry en(h, false);
	if (en.rc) return en.rc;
	h->sink = sink;
	h->sink_extra = extra;
	return 0;
}

int perseus_gpu_set_host_sink(perseus_gpu *h, perseus_gpu_host_sink sink, void *extra)
{
	Entry en(h);
	if (en.rc) return en.rc;
	int rc = flush_locked(h);   // blocks in flight still belong to the previous sink; deliver_slab reads these fields unlocked
	if (rc) return rc;
	h->host_sink = sink;
	h->host_sink_extra = extra;
	return 0;
}

int perseus_gpu_stream_to_file(perseus_gpu *h, const char *path)
{
	Entry en(h);
	if (en.rc) return en.rc;
	int rc = flush_locked(h);
	if (rc) return rc;
	if (h->fout) {
		FILE *f = h->fout;
		h->fout = nullptr;
		if ((h->fout_is_stdout ? fflush(f) : fclose(f)) != 0) return fail(PERSEUS_GPU_IOERROR, "closing stream file failed");
	}
	if (!path) return 0;
	const unsigned f = h->stream_fmt;
	if ((f & PERSEUS_GPU_OUT_INT32) && (f & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2)))
		return fail(PERSEUS_GPU_ERRPARAM, "a stream file holds one format (perseustest -p selects it); open the handle with a single stream format");
	h->fout_is_stdout = strcmp(path, "-") == 0;   // perseustest -o - : the stream goes to standard output, for a consumer on a pipe
	h->fout = h->fout_is_stdout ? stdout : fopen(path, "wb");
	if (!h->fout) return fail(PERSEUS_GPU_IOERROR, "cannot open %s for writing", path);
	return 0;
}

int perseus_gpu_flush(perseus_gpu *h)
{
	Entry en(h);
	if (en.rc) return en.rc;
	return flush_locked(h);
}

int perseus_gpu_get_stats(perseus_gpu *h, perseus_gpu_stats *out)
{
	Entry en(h, false);
	if (en.rc) return en.rc;
	if (!out) return fail(PERSEUS_GPU_ERRPARAM, "null stats pointer");
	*out = h->stats;
	out->host_blocks = h->host_blocks.load(std::memory_order_relaxed);
	return 0;
}

int perseus_gpu_set_tuning(perseus_gpu *h, const perseus_gpu_tuning *t)
{
	if (!h) return fail(PERSEUS_GPU_NULLHANDLE, "null handle");
	pg::Tuning r = resolve_tuning(t);
	int rc = check_tuning(r);
	if (rc) return rc;
	Entry en(h, false);
	memcpy(r.tuned, h->tune.tuned, sizeof(r.tuned));   // autotune results survive; explicit fields take precedence anyway
	h->tune = r;
	return 0;
}

int perseus_gpu_get_geometry(perseus_gpu *h, unsigned flags, int *tile_bytes, int *stages, int *ctas_per_sm)
{
	if (!h) return fail(PERSEUS_GPU_NULLHANDLE, "null handle");
	Entry en(h, false);
	const pg::Geometry g = pg::resolve_geometry(h->tune, flags & 7u);
	if (tile_bytes) *tile_bytes = g.tile_bytes;
	if (stages) *stages = g.stages;
	if (ctas_per_sm) *ctas_per_sm = g.ctas_per_sm;
	return 0;
}

int perseus_gpu_autotune(perseus_gpu *h, double *gbs_single, double *gbs_fused)
{
	Entry en(h);
	if (en.rc) return en.rc;
	int rc = sync_locked(h);
	if (rc) return rc;
	const size_t nbytes = (size_t)87381 * 6144;            // 512 MiB of wire: far beyond L2, 0.15-0.3 ms per launch
	const size_t ns = nbytes / 6;
	uint8_t *in = nullptr, *oi = nullptr, *of = nullptr;
	cudaError_t e = cudaMalloc(&in, nbytes);
	if (e == cudaSuccess) e = cudaMalloc(&oi, ns * 8);
	if (e == cudaSuccess) e = cudaMalloc(&of, ns * 8);
	cudaStream_t st = h->streams[0];
	if (e == cudaSuccess) e = cudaMemsetAsync(in, 0x5A, nbytes, st);
	static const pg::Geometry cand[] = {{12288, 2, 1}, {12288, 3, 1}, {12288, 4, 1}, {12288, 5, 1}, {12288, 6, 1}, {6144, 5, 1},
	                                    {6144, 6, 1},  {6144, 8, 1},  {18432, 2, 1}, {18432, 3, 1}, {24576, 2, 1}, {12288, 2, 2}};
	cudaEvent_t e0 = h->tev[0], e1 = h->tev[1];
	double best_gbs[2] = {0.0, 0.0};
	pg::Geometry best[2] = {};
	for (int cls = 0; cls < 2 && e == cudaSuccess; ++cls) {
		const unsigned fmt = cls ? (pg::FMT_I32 | pg::FMT_F32) : pg::FMT_F32;
		for (const pg::Geometry &g : cand) {
			pg::Tuning t{};
			t.store_mode = h->tune.store_mode;
			t.tile_bytes = g.tile_bytes; t.stages = g.stages; t.ctas_per_sm = g.ctas_per_sm;
			int n = 0;
			float ms = 0.f;
			e = pg::launch_unpack(in, nbytes, cls ? oi : nullptr, of, fmt, t, h->sm_count, st, &n);        // warm
			if (e == cudaSuccess) e = cudaEventRecord(e0, st);
			for (int r = 0; r < 3 && e == cudaSuccess; ++r) e = pg::launch_unpack(in, nbytes, cls ? oi : nullptr, of, fmt, t, h->sm_count, st, &n);
			if (e == cudaSuccess) e = cudaEventRecord(e1, st);
			if (e == cudaSuccess) e = cudaEventSynchronize(e1);
			if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
			if (e != cudaSuccess) break;
			h->stats.kernel_launches += 4;
			const double gbs = (cls ? 22.0 : 14.0) * (double)ns * 3.0 / (ms * 1e-3) / 1e9;
			if (gbs > best_gbs[cls]) { best_gbs[cls] = gbs; best[cls] = g; }
		}
	}
	if (in) cudaFree(in);
	if (oi) cudaFree(oi);
	if (of) cudaFree(of);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(PERSEUS_GPU_CUDAERR, "autotune failed: %s", cudaGetErrorString(e));
	}
	h->tune.tuned[0] = best[0];
	h->tune.tuned[1] = best[1];
	if (gbs_single) *gbs_single = best_gbs[0];
	if (gbs_fused) *gbs_fused = best_gbs[1];
	return 0;
}

int perseus_gpu_get_tuning(perseus_gpu *h, perseus_gpu_tuning *t)
{
	if (!h) return fail(PERSEUS_GPU_NULLHANDLE, "null handle");
	if (!t) return fail(PERSEUS_GPU_ERRPARAM, "null tuning pointer");
	Entry en(h, false);
	memset(t, 0, sizeof(*t));
	t->variant = h->tune.variant;
	t->tile_bytes = h->tune.tile_bytes;
	t->stages = h->tune.stages;
	t->ctas_per_sm = h->tune.ctas_per_sm;
	t->store_mode = h->tune.store_mode;
	return 0;
}

// ---- plumbing ---------------------------------------------------------------------------------------

void *perseus_gpu_dev_alloc(perseus_gpu *h, size_t nbytes)
{
	Entry en(h);
	if (en.rc) return nullptr;
	void *p = nullptr;
	cudaError_t e = cudaMalloc(&p, nbytes ? nbytes : 1);
	if (e != cudaSuccess) {
		cudaGetLastError();
		fail(PERSEUS_GPU_NOMEM, "cudaMalloc(%zu): %s", nbytes, cudaGetErrorString(e));
		return nullptr;
	}
	return p;
}

int perseus_gpu_dev_free(perseus_gpu *h, void *p)
{
	Entry en(h);
	if (en.rc) return en.rc;
	CU(h, cudaFree(p));
	return 0;
}

void *perseus_gpu_host_alloc(perseus_gpu *h, size_t nbytes)
{
	Entry en(h);
	if (en.rc) return nullptr;
	void *p = nullptr;
	cudaError_t e = cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocPortable);
	if (e != cudaSuccess) {
		cudaGetLastError();
		fail(PERSEUS_GPU_NOMEM, "cudaHostAlloc(%zu): %s", nbytes, cudaGetErrorString(e));
		return nullptr;
	}
	return p;
}

int perseus_gpu_host_free(perseus_gpu *h, void *p)
{
	Entry en(h);
	if (en.rc) return en.rc;
	CU(h, cudaFreeHost(p));
	return 0;
}

int perseus_gpu_memcpy(perseus_gpu *h, void *dst, const void *src, size_t nbytes)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (nbytes == 0) return 0;
	CU(h, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDefault, h->streams[0]));
	CU(h, cudaStreamSynchronize(h->streams[0]));
	return 0;
}

int perseus_gpu_memset(perseus_gpu *h, void *dev, int byte, size_t nbytes)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (nbytes == 0) return 0;
	CU(h, cudaMemsetAsync(dev, byte, nbytes, h->streams[0]));
	CU(h, cudaStreamSynchronize(h->streams[0]));
	return 0;
}

void *perseus_gpu_get_stream(perseus_gpu *h, int idx)
{
	if (!h || idx < 0 || idx >= h->nstreams) {
		fail(PERSEUS_GPU_ERRPARAM, "stream index out of range");
		return nullptr;
	}
	return (void *)h->streams[idx];
}

int perseus_gpu_event_record(perseus_gpu *h, int slot)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (slot < 0 || slot >= kEventSlots) return fail(PERSEUS_GPU_ERRPARAM, "event slot %d not in 0..%d", slot, kEventSlots - 1);
	CU(h, cudaEventRecord(h->events[slot], h->streams[0]));
	return 0;
}

int perseus_gpu_event_elapsed_ms(perseus_gpu *h, int a, int b, float *ms)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (a < 0 || a >= kEventSlots || b < 0 || b >= kEventSlots || !ms) return fail(PERSEUS_GPU_ERRPARAM, "bad event slots");
	CU(h, cudaEventSynchronize(h->events[b]));
	CU(h, cudaEventElapsedTime(ms, h->events[a], h->events[b]));
	return 0;
}

// ---- synthetic data / verification --------------------------------------------------------------------

int perseus_gpu_generate(perseus_gpu *h, void *dev_dst, size_t nbytes, int pattern, uint64_t seed, uint64_t byte_offset)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (pattern != PERSEUS_SYNTH_RANDOM && pattern != PERSEUS_SYNTH_RAMP) return fail(PERSEUS_GPU_ERRPARAM, "unknown pattern %d", pattern);
	if (pattern == PERSEUS_SYNTH_RAMP && byte_offset % 6) return fail(PERSEUS_GPU_ERRPARAM, "RAMP byte_offset must be a multiple of 6");
	if (nbytes && !dev_dst) return fail(PERSEUS_GPU_ERRPARAM, "null destination");
	cudaError_t e = pg::launch_generate(dev_dst, nbytes, pattern, seed, byte_offset, h->sm_count, h->streams[0]);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "generate launch failed: %s", cudaGetErrorString(e));
	h->stats.kernel_launches += nbytes ? 1 : 0;
	CU(h, cudaStreamSynchronize(h->streams[0]));
	return 0;
}

int perseus_gpu_checksum(perseus_gpu *h, const void *dev_words, size_t nwords, uint64_t first_index, uint64_t *sum)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (!sum || (nwords && !dev_words)) return fail(PERSEUS_GPU_ERRPARAM, "null argument");
	if ((uintptr_t)dev_words & 3) return fail(PERSEUS_GPU_ERRPARAM, "words must be 4-byte aligned");
	cudaError_t e = pg::launch_checksum(dev_words, nwords, first_index, h->d_scratch, h->sm_count, h->streams[0]);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "checksum launch failed: %s", cudaGetErrorString(e));
	h->stats.kernel_launches += nwords ? 1 : 0;
	CU(h, cudaMemcpyAsync(h->h_scratch, h->d_scratch, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->streams[0]));
	CU(h, cudaStreamSynchronize(h->streams[0]));
	*sum = h->h_scratch[0];
	return 0;
}

int perseus_gpu_verify(perseus_gpu *h, const void *dev_in, size_t nbytes, const void *dev_i32, const void *dev_f32, unsigned flags,
                       uint64_t *nmismatch, uint64_t *first_bad_word)
{
	Entry en(h);
	int rc = en.rc;
	if (rc) return rc;
	unsigned fmt = 0;
	rc = resolve_fmt(flags & ~PERSEUS_GPU_ASYNC, dev_i32, dev_f32, &fmt);
	if (rc) return rc;
	cudaError_t e = pg::launch_verify(dev_in, nbytes, dev_i32, dev_f32, fmt, h->d_scratch, h->sm_count, h->streams[0]);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "verify launch failed: %s", cudaGetErrorString(e));
	h->stats.kernel_launches += nbytes / 6 ? 1 : 0;
	CU(h, cudaMemcpyAsync(h->h_scratch, h->d_scratch, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->streams[0]));
	CU(h, cudaStreamSynchronize(h->streams[0]));
	if (nmismatch) *nmismatch = h->h_scratch[0];
	if (first_bad_word) *first_bad_word = h->h_scratch[1];
	if (h->h_scratch[0])
		return fail(PERSEUS_GPU_MISMATCH, "%llu output words differ from the per-sample recomputation (first at word %llu)", h->h_scratch[0], h->h_scratch[1]);
	return 0;
}

int perseus_gpu_probe_hbm(perseus_gpu *h, int kind, size_t nbytes, int reps, double *gbs)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (kind < 0 || kind > 2 || !gbs || reps < 1 || nbytes < (1u << 20)) return fail(PERSEUS_GPU_ERRPARAM, "bad probe arguments");
	nbytes -= nbytes % 16;
	void *a = nullptr, *b = nullptr;
	cudaError_t e = cudaMalloc(&a, nbytes);
	if (e == cudaSuccess && kind == 2) e = cudaMalloc(&b, nbytes);
	if (e != cudaSuccess) {
		if (a) cudaFree(a);
		cudaGetLastError();
		return fail(PERSEUS_GPU_NOMEM, "probe scratch: %s", cudaGetErrorString(e));
	}
	cudaStream_t st = h->streams[0];
	cudaMemsetAsync(a, 0x5A, nbytes, st);
	double best = 0.0;
	cudaEvent_t e0 = h->tev[0], e1 = h->tev[1];
	for (int ctas : {2, 4, 8, 16}) {
		float ms = 0.f;
		e = pg::launch_probe(kind, a, kind == 2 ? b : a, nbytes, h->sm_count, ctas, st);   // warm
		if (e == cudaSuccess) e = cudaEventRecord(e0, st);
		for (int r = 0; r < reps && e == cudaSuccess; ++r) e = pg::launch_probe(kind, a, kind == 2 ? b : a, nbytes, h->sm_count, ctas, st);
		if (e == cudaSuccess) e = cudaEventRecord(e1, st);
		if (e == cudaSuccess) e = cudaEventSynchronize(e1);
		if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
		if (e != cudaSuccess) break;
		h->stats.kernel_launches += (uint64_t)reps + 1;
		const double g = (kind == 2 ? 2.0 : 1.0) * (double)nbytes * reps / (ms * 1e-3) / 1e9;
		if (g > best) best = g;
	}
	cudaFree(a);
	if (b) cudaFree(b);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "probe failed: %s", cudaGetErrorString(e));
	*gbs = best;
	return 0;
}

int perseus_gpu_probe_pcie(perseus_gpu *h, int kind, size_t nbytes, size_t d2h_nbytes, int reps, double *h2d_gbs, double *d2h_gbs)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (d2h_nbytes == 0) d2h_nbytes = nbytes;
	if (kind < 0 || kind > 2 || reps < 1 || nbytes < (1u << 20) || d2h_nbytes < (1u << 20)) return fail(PERSEUS_GPU_ERRPARAM, "bad probe arguments");
	int rc = sync_locked(h);
	if (rc) return rc;
	const bool up = kind != PERSEUS_GPU_PCIE_D2H, down = kind != PERSEUS_GPU_PCIE_H2D;
	void *hu = nullptr, *du = nullptr, *hd = nullptr, *dd = nullptr;
	cudaError_t e = cudaSuccess;
	if (up) {
		e = cudaHostAlloc(&hu, nbytes, cudaHostAllocDefault);
		if (e == cudaSuccess) e = cudaMalloc(&du, nbytes);
		if (e == cudaSuccess) memset(hu, 0x5A, nbytes);
	}
	if (down && e == cudaSuccess) {
		e = cudaHostAlloc(&hd, d2h_nbytes, cudaHostAllocDefault);
		if (e == cudaSuccess) e = cudaMalloc(&dd, d2h_nbytes);
		if (e == cudaSuccess) memset(hd, 0, d2h_nbytes);   // touch the pages before timing
	}
	double best_up = 0.0, best_down = 0.0;
	for (int r = 0; r <= reps && e == cudaSuccess; ++r) {   // r == 0 warms up
		// both directions are queued before either is waited for, each on its own stream, so in DUPLEX mode they overlap
		if (up) {
			e = cudaEventRecord(h->tev[0], h->s_in);
			if (e == cudaSuccess) e = cudaMemcpyAsync(du, hu, nbytes, cudaMemcpyHostToDevice, h->s_in);
			if (e == cudaSuccess) e = cudaEventRecord(h->tev[1], h->s_in);
		}
		if (down && e == cudaSuccess) {
			e = cudaEventRecord(h->tev[2], h->s_out);
			if (e == cudaSuccess) e = cudaMemcpyAsync(hd, dd, d2h_nbytes, cudaMemcpyDeviceToHost, h->s_out);
			if (e == cudaSuccess) e = cudaEventRecord(h->tev[3], h->s_out);
		}
		if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_in);
		if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_out);
		float ms = 0.f;
		if (up && e == cudaSuccess) {
			e = cudaEventElapsedTime(&ms, h->tev[0], h->tev[1]);
			const double g = (double)nbytes / (ms * 1e-3) / 1e9;
			if (r && g > best_up) best_up = g;
		}
		if (down && e == cudaSuccess) {
			e = cudaEventElapsedTime(&ms, h->tev[2], h->tev[3]);
			const double g = (double)d2h_nbytes / (ms * 1e-3) / 1e9;
			if (r && g > best_down) best_down = g;
		}
		if (e == cudaSuccess) {
			if (up) h->stats.h2d_bytes += nbytes;
			if (down) h->stats.d2h_bytes += d2h_nbytes;
		}
	}
	if (hu) cudaFreeHost(hu);
	if (hd) cudaFreeHost(hd);
	if (du) cudaFree(du);
	if (dd) cudaFree(dd);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(PERSEUS_GPU_CUDAERR, "PCIe probe failed: %s", cudaGetErrorString(e));
	}
	if (h2d_gbs) *h2d_gbs = best_up;
	if (d2h_gbs) *d2h_gbs = best_down;
	return 0;
}

}  // extern "C"
