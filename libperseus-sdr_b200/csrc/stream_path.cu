// C-ABI layer of libperseus_gpu.so, part 2 of 3 (see handle.h): the streaming hand-off -- the drop-in for the reference's user callback.
#include "handle.h"

namespace pgh {

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

// Submits the partial slab if its oldest transfer is over age.  Returns 1 if it did, 0 if not, < 0 on error.
int submit_if_over_age(perseus_gpu *h, uint64_t now)
{
	if (!h->max_latency_ns || !h->fill || h->latched) return 0;
	if (now - h->fill_started_ns < h->max_latency_ns) return 0;
	const int rc = submit_slab(h);
	return rc ? rc : 1;
}

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
	// The gap runs from the END of the previous callback: time that callback spent submitting a slab, or waiting for a free one,
	// is the path's, not the stream's.  (Counted from its start, a path that has become the bottleneck -- a GPU busy with other
	// work, a profiler that makes every launch take milliseconds -- would see every transfer arrive "late", submit each one
	// alone and make itself slower still.)  A callback that only copied ended ~0.2 us after it began: its one clock reading
	// serves; only a callback that submitted something reads the clock again.
	const bool eager = h->eager_gap_ns && h->last_push_ns && now - h->last_push_ns > h->eager_gap_ns;
	const uint64_t slabs_before = h->stats.slabs;
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
	h->last_push_ns = (h->eager_gap_ns && h->stats.slabs != slabs_before) ? callback_now_ns(h) : now;
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

}  // namespace pgh

using namespace pgh;

// =============================================================================== C ABI

extern "C" {

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
	const uint64_t before = h->samples_submitted * 6 + h->fill;   // bytes the stream has taken in so far
	int rc = stream_push(h, buf, n);
	if (rc) {
		latch(h, rc);
		// the failing transfer itself: whatever part of it did not make it into a slab is lost like the ones that follow
		const uint64_t taken = h->samples_submitted * 6 + h->fill - before;
		if (taken < n) {
			h->stats.dropped_bytes += n - taken;
			if (taken == 0) {
				h->stats.callbacks--;
				h->stats.dropped_callbacks++;
			}
		}
	}
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
	// The first launch of a kernel loads its code: do that here, on two samples' worth of slab 0 (nothing is delivered or
	// counted).  Only while the stream has not begun -- afterwards slab 0 holds live data, and the code is loaded anyway.
	if (h->stats.slabs == 0 && h->fill == 0 && h->stats.callbacks == 0) {
		Slab &s = h->slabs[0];
		memset(s.host, 0, 12);
		int n = 0;
		cudaError_t e = pg::launch_unpack(s.host, 12, s.dev_i32, s.dev_f32, h->stream_fmt, h->tune, h->sm_count, h->streams[0], &n);
		if (e == cudaSuccess) e = cudaStreamSynchronize(h->streams[0]);
		if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "warm-up launch failed: %s", cudaGetErrorString(e));
		h->stats.kernel_launches += (uint64_t)n;
	}
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

}  // extern "C"
