// Virtual receiver: the reference's transfer-delivery semantics without the hardware.
//
// What is reproduced (and where it is in the reference):
//   * perseus_start_async_input's argument checks and error codes     perseus-sdr.c:638-680
//   * a queue of 8 transfers over ONE contiguous pageable ring          perseus-sdr.c:683, perseus-in.c:68,83-91
//   * per completed transfer: count bytes, deliver only if it is the expected slot AND full
//     length, otherwise log-and-drop; then expect (idx+1)%8 and re-arm  perseus-in.c:199-216,260-263
//   * the other completion statuses: TIMED_OUT is logged and the slot re-armed; ERROR / STALL /
//     NO_DEVICE / OVERFLOW retire the slot for good WITHOUT advancing the expected index   perseus-in.c:218-257
//   * stop: cancel, wait, report elapsed / kSamples / kS/s              perseus-sdr.c:694-734
//   * nearest-rate selection over the ten bitstream rates               perseus-sdr.c:776-811
// What replaces the USB device: the synthetic wire-data generator (host_common.h host_generate), so
// transfer number n of a stream carries bytes [n*size, (n+1)*size) of the synthetic recording.
// Checked against the reference's own code (perseus-in.c and perseus-sdr.c compiled unmodified over a
// fake libusb, oracle/_ref) by tests/test_refqueue_cpu.py and tests/test_reflib_cpu.py.
//
// No CUDA in this file: it builds with a plain C++ compiler (tests/sanitize/sanitize.sh runs it under TSAN/ASAN).
#include "../../include/perseus-gpu.h"
#include "host_common.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <new>
#include <thread>

namespace {

using pg::fail;

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
	perseus_gpu_input_fn cb = nullptr;
	void *cb_extra = nullptr;
	int idx_expected = 0;
	// slots in the order they were (re)submitted to the device; a retired slot never comes back (perseus-in.c:222-257)
	int fifo[PERSEUS_VRX_QUEUE_SIZE] = {0, 1, 2, 3, 4, 5, 6, 7};
	int fifo_head = 0, fifo_count = PERSEUS_VRX_QUEUE_SIZE;
	uint64_t submitted = 0;               // transfers the "device" has filled so far == stream position
	std::atomic<bool> cancelling{false};
	bool started = false;
	std::thread worker;
	Clock::time_point t_start;
	// written by the delivery thread ONLY, read by perseus_vrx_get_stats on the application thread: atomic so the read is
	// defined, but bumped with a plain load + store (see bump()), never a locked read-modify-write
	std::atomic<uint64_t> bytes_received{0}, delivered{0}, dropped_short{0}, dropped_sequence{0}, timed_out{0}, retired{0};
	double elapsed_s = 0.0;               // frozen by stop / run
};

namespace {

int validate_size(const perseus_vrx *v, uint32_t buffersize)
{
	if (buffersize > PERSEUS_VRX_MAX_BUFFER) return fail(PERSEUS_GPU_ERRPARAM, "max libusb bulk buffer size is 16320 bytes");
	const int maxps = v->cfg.ep_max_packet ? v->cfg.ep_max_packet : 512;
	if (maxps == 512) {
		if (buffersize % 6144) return fail(PERSEUS_GPU_BUFFERSIZE, "buffer size should be an integer multiple of 6144 bytes (1024 I/Q samples)");
	} else if (maxps == 510) {
		if (buffersize % 510) return fail(PERSEUS_GPU_BUFFERSIZE, "buffer size should be an integer multiple of 510 bytes (85 IQ samples)");
	} else {
		return fail(PERSEUS_GPU_ERRPARAM, "Unexpected max packet size: %d", maxps);
	}
	// Deliberate difference: the reference lets 0 through (0 % 6144 == 0, perseus-sdr.c:671) and then spins on
	// empty transfers; a zero-length stream is refused here.
	if (buffersize == 0) return fail(PERSEUS_GPU_BUFFERSIZE, "buffer size is zero");
	return 0;
}

// Single-writer counter increment.  A `lock xadd` here would sit on the callback path of the GPU trampoline, whose slab copy
// uses non-temporal stores: every locked instruction waits for the write-combining buffers to drain (~0.25 us per transfer).
inline void bump(std::atomic<uint64_t> &c, uint64_t by = 1) { c.store(c.load(std::memory_order_relaxed) + by, std::memory_order_relaxed); }

// What the device does with transfer number `seq` (1-based) of the stream.
struct Outcome { int status; uint32_t actual; };

Outcome outcome_of(const perseus_vrx *v, uint64_t seq)
{
	if (v->cfg.fail_at && seq == v->cfg.fail_at) return {(int)v->cfg.fail_status, 0u};
	if (v->cfg.timeout_every && seq % v->cfg.timeout_every == 0) return {PERSEUS_VRX_STATUS_TIMED_OUT, 0u};
	if (v->cfg.drop_every && seq % v->cfg.drop_every == 0) return {PERSEUS_VRX_STATUS_COMPLETED, v->size - 6};
	return {PERSEUS_VRX_STATUS_COMPLETED, v->size};
}

// The reference's completion handler (perseus-in.c:187-264) for slot `idx`.  Returns true when the slot is
// re-armed (resubmitted), false when it is retired.
bool complete_transfer(perseus_vrx *v, int idx, Outcome o)
{
	switch (o.status) {
	case PERSEUS_VRX_STATUS_COMPLETED:
		bump(v->bytes_received, o.actual);
		if (idx == v->idx_expected) {
			if (o.actual == v->size) {
				if (v->cb) v->cb(v->ring + (size_t)idx * v->size, (int)v->size, v->cb_extra);
				bump(v->delivered);
			} else {
				bump(v->dropped_short);
			}
		} else {
			bump(v->dropped_sequence);
		}
		break;
	case PERSEUS_VRX_STATUS_TIMED_OUT:   // logged only; falls out of the switch to the index update and the resubmit
		bump(v->timed_out);
		break;
	default:                            // ERROR, STALL, NO_DEVICE, OVERFLOW: slot marked cancelled, early return
		bump(v->retired);
		return false;
	}
	v->idx_expected = (idx + 1) % PERSEUS_VRX_QUEUE_SIZE;
	return true;
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
	v->fifo_count--;
	return idx;
}
void fifo_push(perseus_vrx *v, int idx)
{
	v->fifo[(v->fifo_head + v->fifo_count) % PERSEUS_VRX_QUEUE_SIZE] = idx;
	v->fifo_count++;
}

// Delivers `limit` completions (UINT64_MAX: until cancelled, or until every slot has been retired).
void deliver(perseus_vrx *v, uint64_t limit)
{
	uint64_t n = 0;            // completions so far in this run
	while (n < limit && v->fifo_count > 0 && !v->cancelling.load(std::memory_order_acquire)) {
		const uint64_t seq = v->submitted + 1;   // 1-based number of the transfer about to complete
		const bool swap = v->cfg.swap_every && seq % v->cfg.swap_every == 0 && n + 1 < limit && v->fifo_count >= 2;
		const int a = fifo_pop(v);
		if (swap) {
			// the two oldest submissions complete in the wrong order; each carries the data of its own stream position
			const int b = fifo_pop(v);
			arm_transfer(v, a);
			arm_transfer(v, b);
			pace(v, n + 2);
			if (complete_transfer(v, b, outcome_of(v, seq + 1))) fifo_push(v, b);
			if (complete_transfer(v, a, outcome_of(v, seq))) fifo_push(v, a);
			n += 2;
			continue;
		}
		arm_transfer(v, a);
		pace(v, n + 1);
		if (complete_transfer(v, a, outcome_of(v, seq))) fifo_push(v, a);
		++n;
	}
}

int setup(perseus_vrx *v, uint32_t buffersize, perseus_gpu_input_fn cb, void *extra)
{
	if (v->started) return fail(PERSEUS_GPU_ASYNCSTARTED, "async input already started");
	int rc = validate_size(v, buffersize);
	if (rc) return rc;
	uint8_t *ring = static_cast<uint8_t *>(malloc((size_t)PERSEUS_VRX_QUEUE_SIZE * buffersize));
	if (!ring) return fail(PERSEUS_GPU_NOMEM, "can't allocate datain buffer");
	free(v->ring);
	v->ring = ring;
	v->size = buffersize;
	v->cb = cb;
	v->cb_extra = extra;
	v->idx_expected = 0;
	v->submitted = 0;                                                  // every start is a new stream
	for (int k = 0; k < PERSEUS_VRX_QUEUE_SIZE; ++k) v->fifo[k] = k;   // perseus-in.c:95-96 submits slots 0..7 in order
	v->fifo_head = 0;
	v->fifo_count = PERSEUS_VRX_QUEUE_SIZE;
	v->cancelling.store(false);
	for (std::atomic<uint64_t> *c : {&v->bytes_received, &v->delivered, &v->dropped_short, &v->dropped_sequence, &v->timed_out, &v->retired})
		c->store(0, std::memory_order_relaxed);
	v->elapsed_s = 0.0;
	v->t_start = Clock::now();
	return 0;
}

void fill_stats(const perseus_vrx *v, double elapsed_s, perseus_vrx_stats *out)
{
	out->bytes_received = v->bytes_received.load(std::memory_order_relaxed);
	out->delivered = v->delivered.load(std::memory_order_relaxed);
	out->dropped_short = v->dropped_short.load(std::memory_order_relaxed);
	out->dropped_sequence = v->dropped_sequence.load(std::memory_order_relaxed);
	out->timed_out = v->timed_out.load(std::memory_order_relaxed);
	out->retired = v->retired.load(std::memory_order_relaxed);
	out->elapsed_s = elapsed_s;
	out->ksamples_per_s = elapsed_s > 0 ? 1.0 * (double)out->bytes_received / elapsed_s / 6000.0 : 0.0;   // perseus-sdr.c:721-722
}

}  // namespace

extern "C" {

int perseus_vrx_get_sampling_rates(int *buf, unsigned int size)
{
	if (size == 0 || !buf) return fail(PERSEUS_GPU_ERRPARAM, "Zero lenght buffer");
	for (unsigned i = 0; i < size; ++i) buf[i] = 0;
	if (size < (unsigned)kNumRates) {
		for (unsigned i = 0; i < size; ++i) buf[i] = kRates[i];
		return fail(PERSEUS_GPU_BUFFERSIZE, "Insufficient buffer size");
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
	if (!out) return fail(PERSEUS_GPU_ERRPARAM, "null handle pointer");
	*out = nullptr;
	perseus_vrx_config cfg{};
	if (ucfg) {
		if (ucfg->struct_size < 8 || ucfg->struct_size > sizeof(cfg)) return fail(PERSEUS_GPU_ERRPARAM, "perseus_vrx_config.struct_size %u not understood", ucfg->struct_size);
		memcpy(&cfg, ucfg, ucfg->struct_size);
	}
	if (cfg.pattern != PERSEUS_SYNTH_RANDOM && cfg.pattern != PERSEUS_SYNTH_RAMP) return fail(PERSEUS_GPU_ERRPARAM, "unknown pattern %d", cfg.pattern);
	if (cfg.ep_max_packet != 0 && cfg.ep_max_packet != 512 && cfg.ep_max_packet != 510)
		return fail(PERSEUS_GPU_ERRPARAM, "Unexpected max packet size: %d", cfg.ep_max_packet);
	if (cfg.fail_at) {
		const uint32_t st = cfg.fail_status;
		if (st != PERSEUS_VRX_STATUS_ERROR && st != PERSEUS_VRX_STATUS_STALL && st != PERSEUS_VRX_STATUS_NO_DEVICE && st != PERSEUS_VRX_STATUS_OVERFLOW)
			return fail(PERSEUS_GPU_ERRPARAM, "fail_status %u is not ERROR, STALL, NO_DEVICE or OVERFLOW", st);
	}
	perseus_vrx *v = new (std::nothrow) perseus_vrx();
	if (!v) return fail(PERSEUS_GPU_NOMEM, "out of memory");
	v->cfg = cfg;
	v->rate = perseus_vrx_nearest_rate(cfg.sample_rate ? cfg.sample_rate : 95000);   // perseustest.c:100 default
	*out = v;
	return 0;
}

int perseus_vrx_close(perseus_vrx *v)
{
	if (!v) return fail(PERSEUS_GPU_NULLHANDLE, "null descriptor");
	if (v->started) perseus_vrx_stop_async_input(v);
	free(v->ring);
	delete v;
	return 0;
}

int perseus_vrx_get_sampling_rate(perseus_vrx *v)
{
	if (!v) return fail(PERSEUS_GPU_NULLHANDLE, "null descriptor");
	return v->rate;
}

int perseus_vrx_start_async_input(perseus_vrx *v, uint32_t buffersize, perseus_gpu_input_fn callback, void *cb_extra)
{
	if (!v) return fail(PERSEUS_GPU_NULLHANDLE, "null descriptor");
	int rc = setup(v, buffersize, callback, cb_extra);
	if (rc) return rc;
	v->started = true;
	try {
		v->worker = std::thread([v] { deliver(v, UINT64_MAX); });
	} catch (...) {
		v->started = false;
		return fail(PERSEUS_GPU_NOMEM, "can't create delivery thread");
	}
	return 0;
}

int perseus_vrx_stop_async_input(perseus_vrx *v)
{
	if (!v) return fail(PERSEUS_GPU_NULLHANDLE, "null descriptor");
	if (!v->started) return fail(PERSEUS_GPU_ASYNCSTARTED, "async input not started");
	v->cancelling.store(true, std::memory_order_release);
	if (v->worker.joinable()) v->worker.join();
	v->elapsed_s = std::chrono::duration<double>(Clock::now() - v->t_start).count();
	v->cb = nullptr;
	v->started = false;
	return 0;
}

int perseus_vrx_run(perseus_vrx *v, uint32_t buffersize, perseus_gpu_input_fn callback, void *cb_extra, uint64_t ntransfers)
{
	if (!v) return fail(PERSEUS_GPU_NULLHANDLE, "null descriptor");
	int rc = setup(v, buffersize, callback, cb_extra);
	if (rc) return rc;
	deliver(v, ntransfers);
	v->elapsed_s = std::chrono::duration<double>(Clock::now() - v->t_start).count();
	v->cb = nullptr;
	return 0;
}

int perseus_vrx_get_stats(perseus_vrx *v, perseus_vrx_stats *out)
{
	if (!v) return fail(PERSEUS_GPU_NULLHANDLE, "null descriptor");
	if (!out) return fail(PERSEUS_GPU_ERRPARAM, "null stats pointer");
	// live view while streaming, frozen elapsed time afterwards
	fill_stats(v, v->started ? std::chrono::duration<double>(Clock::now() - v->t_start).count() : v->elapsed_s, out);
	return 0;
}

}  // extern "C"
