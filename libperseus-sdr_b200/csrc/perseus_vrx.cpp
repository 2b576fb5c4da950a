// Virtual receiver: the reference's transfer-delivery semantics without the hardware.
//
// What is reproduced (and where it is in the reference):
//   * perseus_start_async_input's argument checks and error codes     perseus-sdr.c:638-680
//   * a queue of 8 transfers over ONE contiguous pageable ring          perseus-sdr.c:683, perseus-in.c:68,83-91
//   * per completed transfer: count bytes, deliver only if it is the expected slot AND full
//     length, otherwise log-and-drop; then expect (idx+1)%8 and re-arm  perseus-in.c:199-216,260-263
//   * stop: cancel, wait, report elapsed / kSamples / kS/s              perseus-sdr.c:694-734
//   * nearest-rate selection over the ten bitstream rates               perseus-sdr.c:776-811
// What replaces the USB device: the synthetic wire-data generator (kernels.h host_generate), so
// transfer number n of a stream carries bytes [n*size, (n+1)*size) of the synthetic recording.
#include "../../include/perseus-gpu.h"
#include "kernels.h"

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>

namespace {

int vfail(int code, const char *fmt, ...);

// perseus-sdr.h:282-285 / generate_fpga_code.sh:71-97 — rates encoded in the bitstream file names
const int kRates[] = {48000, 95000, 96000, 125000, 192000, 250000, 500000, 1000000, 1600000, 2000000};
constexpr int kNumRates = sizeof(kRates) / sizeof(kRates[0]);
// the bitstream files those rates come from (/root/reference/*.rbs), same order
const char *const kBitstreams[] = {"perseus48k24v31",  "perseus95k24v31",  "perseus96k24v31", "perseus125k24v21", "perseus192k24v31",
                                   "perseus250k24v21", "perseus500k24v21", "perseus1m24v21",  "perseus1d6m24v21", "perseus2m24v21"};

using Clock = std::chrono::steady_clock;

}  // namespace

struct perseus_vrx {
	perseus_vrx_config cfg{};
	int rate = 0;
	uint8_t *ring = nullptr;              // PERSEUS_VRX_QUEUE_SIZE * size bytes, pageable on purpose
	uint32_t size = 0;
	perseus_input_callback cb = nullptr;
	void *cb_extra = nullptr;
	int idx_expected = 0;
	int fifo[PERSEUS_VRX_QUEUE_SIZE] = {0, 1, 2, 3, 4, 5, 6, 7};   // slots in the order they were (re)submitted to the device
	int fifo_head = 0;
	uint64_t submitted = 0;               // transfers handed to the "device" so far == stream position
	std::atomic<bool> cancelling{false};
	bool started = false;
	std::thread worker;
	Clock::time_point t_start, t_stop;
	perseus_vrx_stats stats{};
};

extern "C" const char *perseus_gpu_errorstr(void);

namespace {

// the message lands in the same thread-local string perseus_gpu_errorstr() returns
int vfail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(const_cast<char *>(perseus_gpu_errorstr()), 512, fmt, ap);
	va_end(ap);
	return code;
}

int validate_size(const perseus_vrx *v, uint32_t buffersize)
{
	if (buffersize > PERSEUS_VRX_MAX_BUFFER) return vfail(PERSEUS_GPU_ERRPARAM, "max libusb bulk buffer size is 16320 bytes");
	const int maxps = v->cfg.ep_max_packet ? v->cfg.ep_max_packet : 512;
	if (maxps == 512) {
		if (buffersize % 6144) return vfail(PERSEUS_GPU_BUFFERSIZE, "buffer size should be an integer multiple of 6144 bytes (1024 I/Q samples)");
	} else if (maxps == 510) {
		if (buffersize % 510) return vfail(PERSEUS_GPU_BUFFERSIZE, "buffer size should be an integer multiple of 510 bytes (85 IQ samples)");
	} else {
		return vfail(PERSEUS_GPU_ERRPARAM, "Unexpected max packet size: %d", maxps);
	}
	if (buffersize == 0) return vfail(PERSEUS_GPU_BUFFERSIZE, "buffer size is zero");
	return 0;
}

// One completed transfer in slot `idx` with `actual` bytes: the body of the reference's
// completion handler for LIBUSB_TRANSFER_COMPLETED.
void complete_transfer(perseus_vrx *v, int idx, uint32_t actual)
{
	v->stats.bytes_received += actual;
	if (idx == v->idx_expected) {
		if (actual == v->size) {
			if (v->cb) v->cb(v->ring + (size_t)idx * v->size, (int)v->size, v->cb_extra);
			v->stats.delivered++;
		} else {
			v->stats.dropped_short++;
		}
	} else {
		v->stats.dropped_sequence++;
	}
	v->idx_expected = (idx + 1) % PERSEUS_VRX_QUEUE_SIZE;
}

// "Device side": fill slot idx with the next `size` bytes of the synthetic stream.
void arm_transfer(perseus_vrx *v, int idx)
{
	const uint64_t seed = v->cfg.seed ? v->cfg.seed : PERSEUS_SYNTH_SEED;
	if (v->cfg.replay && v->submitted >= PERSEUS_VRX_QUEUE_SIZE) {   // ring already holds transfers 0..7: replay them
		v->submitted++;
		return;
	}
	pg::host_generate(v->ring + (size_t)idx * v->size, v->size, v->cfg.pattern, seed, v->submitted * (uint64_t)v->size);
	v->submitted++;
}

void pace(perseus_vrx *v, uint64_t transfers_done)
{
	if (!v->cfg.realtime) return;
	const double due = (double)transfers_done * (double)(v->size / 6) / (double)v->rate;
	std::this_thread::sleep_until(v->t_start + std::chrono::duration_cast<Clock::duration>(std::chrono::duration<double>(due)));
}

// The device completes transfers in the order they were submitted; a completed transfer is resubmitted at once
// (perseus-in.c:263), so it re-enters the FIFO at the tail.  With no faults that is the cyclic order 0..7,0..;
// after an out-of-order completion the submission order stays permuted, exactly as it would with libusb.
int fifo_pop(perseus_vrx *v)
{
	const int idx = v->fifo[v->fifo_head];
	v->fifo_head = (v->fifo_head + 1) % PERSEUS_VRX_QUEUE_SIZE;
	return idx;
}
void fifo_push(perseus_vrx *v, int idx, int free_slots_before)
{
	// the FIFO always holds QUEUE_SIZE - free_slots_before entries starting at fifo_head
	v->fifo[(v->fifo_head + PERSEUS_VRX_QUEUE_SIZE - free_slots_before) % PERSEUS_VRX_QUEUE_SIZE] = idx;
}

// Delivers `limit` completions (UINT64_MAX: until cancelled).
void deliver(perseus_vrx *v, uint64_t limit)
{
	uint64_t n = 0;            // completions so far in this run
	while (n < limit && !v->cancelling.load(std::memory_order_acquire)) {
		const uint64_t seq = v->submitted + 1;   // 1-based number of the transfer about to complete
		const bool swap = v->cfg.swap_every && seq % v->cfg.swap_every == 0 && n + 1 < limit;
		const int a = fifo_pop(v);
		if (swap) {
			// the two oldest submissions complete in the wrong order; each carries the data of its own stream position
			const int b = fifo_pop(v);
			arm_transfer(v, a);
			arm_transfer(v, b);
			pace(v, n + 2);
			const bool a_short = v->cfg.drop_every && seq % v->cfg.drop_every == 0;
			const bool b_short = v->cfg.drop_every && (seq + 1) % v->cfg.drop_every == 0;
			complete_transfer(v, b, b_short ? v->size - 6 : v->size);
			fifo_push(v, b, 2);
			complete_transfer(v, a, a_short ? v->size - 6 : v->size);
			fifo_push(v, a, 1);
			n += 2;
			continue;
		}
		arm_transfer(v, a);
		pace(v, n + 1);
		const bool is_short = v->cfg.drop_every && seq % v->cfg.drop_every == 0;
		complete_transfer(v, a, is_short ? v->size - 6 : v->size);
		fifo_push(v, a, 1);
		++n;
	}
}

void finish_stats(perseus_vrx *v)
{
	v->stats.elapsed_s = std::chrono::duration<double>(v->t_stop - v->t_start).count();
	v->stats.ksamples_per_s = v->stats.elapsed_s > 0 ? 1.0 * (double)v->stats.bytes_received / v->stats.elapsed_s / 6000.0 : 0.0;
}

int setup(perseus_vrx *v, uint32_t buffersize, perseus_input_callback cb, void *extra)
{
	if (v->started) return vfail(PERSEUS_GPU_ASYNCSTARTED, "async input already started");
	int rc = validate_size(v, buffersize);
	if (rc) return rc;
	uint8_t *ring = static_cast<uint8_t *>(malloc((size_t)PERSEUS_VRX_QUEUE_SIZE * buffersize));
	if (!ring) return vfail(PERSEUS_GPU_NOMEM, "can't allocate datain buffer");
	free(v->ring);
	v->ring = ring;
	v->size = buffersize;
	v->cb = cb;
	v->cb_extra = extra;
	v->idx_expected = 0;
	v->submitted = 0;                                                  // every start is a new stream
	for (int k = 0; k < PERSEUS_VRX_QUEUE_SIZE; ++k) v->fifo[k] = k;   // perseus-in.c:95-96 submits slots 0..7 in order
	v->fifo_head = 0;
	v->cancelling.store(false);
	v->stats = perseus_vrx_stats{};
	v->t_start = Clock::now();
	return 0;
}

}  // namespace

extern "C" {

int perseus_vrx_get_sampling_rates(int *buf, unsigned int size)
{
	if (size == 0 || !buf) return vfail(PERSEUS_GPU_ERRPARAM, "Zero lenght buffer");
	for (unsigned i = 0; i < size; ++i) buf[i] = 0;
	if (size < (unsigned)kNumRates) {
		for (unsigned i = 0; i < size; ++i) buf[i] = kRates[i];
		return vfail(PERSEUS_GPU_BUFFERSIZE, "Insufficient buffer size");
	}
	for (int i = 0; i < kNumRates; ++i) buf[i] = kRates[i];
	return 0;
}

const char *perseus_vrx_bitstream_name(int rate)
{
	for (int i = 0; i < kNumRates; ++i)
		if (kRates[i] == rate) return kBitstreams[i];
	return nullptr;
}

int perseus_vrx_nearest_rate(int requested)
{
	// closest table entry; an exact midpoint goes to the LOWER rate; out-of-range requests clamp
	int best = kRates[0];
	long long best_d = llabs((long long)requested - best);
	for (int i = 1; i < kNumRates; ++i) {
		const long long d = llabs((long long)requested - kRates[i]);
		if (d < best_d) { best = kRates[i]; best_d = d; }
	}
	return best;
}

int perseus_vrx_open(perseus_vrx **out, const perseus_vrx_config *ucfg)
{
	if (!out) return vfail(PERSEUS_GPU_ERRPARAM, "null handle pointer");
	*out = nullptr;
	perseus_vrx_config cfg{};
	if (ucfg) {
		if (ucfg->struct_size < 8 || ucfg->struct_size > sizeof(cfg)) return vfail(PERSEUS_GPU_ERRPARAM, "perseus_vrx_config.struct_size %u not understood", ucfg->struct_size);
		memcpy(&cfg, ucfg, ucfg->struct_size);
	}
	if (cfg.pattern != PERSEUS_SYNTH_RANDOM && cfg.pattern != PERSEUS_SYNTH_RAMP) return vfail(PERSEUS_GPU_ERRPARAM, "unknown pattern %d", cfg.pattern);
	if (cfg.ep_max_packet != 0 && cfg.ep_max_packet != 512 && cfg.ep_max_packet != 510)
		return vfail(PERSEUS_GPU_ERRPARAM, "Unexpected max packet size: %d", cfg.ep_max_packet);
	perseus_vrx *v = new (std::nothrow) perseus_vrx();
	if (!v) return vfail(PERSEUS_GPU_NOMEM, "out of memory");
	v->cfg = cfg;
	v->rate = perseus_vrx_nearest_rate(cfg.sample_rate ? cfg.sample_rate : 95000);   // perseustest.c:100 default
	*out = v;
	return 0;
}

int perseus_vrx_close(perseus_vrx *v)
{
	if (!v) return vfail(PERSEUS_GPU_NULLHANDLE, "null descriptor");
	if (v->started) perseus_vrx_stop_async_input(v);
	free(v->ring);
	delete v;
	return 0;
}

int perseus_vrx_get_sampling_rate(perseus_vrx *v)
{
	if (!v) return vfail(PERSEUS_GPU_NULLHANDLE, "null descriptor");
	return v->rate;
}

int perseus_vrx_start_async_input(perseus_vrx *v, uint32_t buffersize, perseus_input_callback callback, void *cb_extra)
{
	if (!v) return vfail(PERSEUS_GPU_NULLHANDLE, "null descriptor");
	int rc = setup(v, buffersize, callback, cb_extra);
	if (rc) return rc;
	v->started = true;
	try {
		v->worker = std::thread([v] { deliver(v, UINT64_MAX); });
	} catch (...) {
		v->started = false;
		return vfail(PERSEUS_GPU_NOMEM, "can't create delivery thread");
	}
	return 0;
}

int perseus_vrx_stop_async_input(perseus_vrx *v)
{
	if (!v) return vfail(PERSEUS_GPU_NULLHANDLE, "null descriptor");
	if (!v->started) return vfail(PERSEUS_GPU_ASYNCSTARTED, "async input not started");
	v->cancelling.store(true, std::memory_order_release);
	if (v->worker.joinable()) v->worker.join();
	v->t_stop = Clock::now();
	v->cb = nullptr;
	v->started = false;
	finish_stats(v);
	return 0;
}

int perseus_vrx_run(perseus_vrx *v, uint32_t buffersize, perseus_input_callback callback, void *cb_extra, uint64_t ntransfers)
{
	if (!v) return vfail(PERSEUS_GPU_NULLHANDLE, "null descriptor");
	int rc = setup(v, buffersize, callback, cb_extra);
	if (rc) return rc;
	deliver(v, ntransfers);
	v->t_stop = Clock::now();
	v->cb = nullptr;
	finish_stats(v);
	return 0;
}

int perseus_vrx_get_stats(perseus_vrx *v, perseus_vrx_stats *out)
{
	if (!v) return vfail(PERSEUS_GPU_NULLHANDLE, "null descriptor");
	if (!out) return vfail(PERSEUS_GPU_ERRPARAM, "null stats pointer");
	if (v->started) {   // live view
		perseus_vrx_stats s = v->stats;
		s.elapsed_s = std::chrono::duration<double>(Clock::now() - v->t_start).count();
		s.ksamples_per_s = s.elapsed_s > 0 ? (double)s.bytes_received / s.elapsed_s / 6000.0 : 0.0;
		*out = s;
	} else {
		*out = v->stats;
	}
	return 0;
}

}  // extern "C"
