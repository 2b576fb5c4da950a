// TEST INFRASTRUCTURE ONLY: drives the host layer of libperseus_gpu (handle.cu, stream_path.cu, bulk_path.cu, perseus_vrx.cpp,
// perseus_host.cpp) from several threads at once so ThreadSanitizer / AddressSanitizer can see it.  Built by
// tests/sanitize/sanitize.sh against tests/sanitize/fake_cuda (no GPU involved); results are also checked against the CPU oracle.
#include "../../include/perseus-gpu.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

extern "C" size_t perseus_oracle_unpack(int mode, const uint8_t *in, size_t nbytes, void *out);
namespace pg { void copy_nontemporal(uint8_t *dst, const uint8_t *src, size_t n); }   // csrc/copy_pool.h
extern "C" void fake_cuda_fail_alloc_in(long n);                                        // tests/sanitize/fake_cuda.cpp

#define CHECK(c)                                                                          \
	do {                                                                                  \
		if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s  [%s]\n", __FILE__, __LINE__, #c, perseus_gpu_errorstr()); exit(1); } \
	} while (0)

static void sleep_ms(int ms) { std::this_thread::sleep_for(std::chrono::milliseconds(ms)); }

struct Collected {
	std::mutex mu;
	std::vector<uint8_t> bytes;
	perseus_gpu *h = nullptr;
};

static void sink(const perseus_gpu_block *b, void *extra)
{
	Collected *c = static_cast<Collected *>(extra);
	std::lock_guard<std::mutex> lk(c->mu);
	const size_t n = b->nsamples * 8, at = c->bytes.size();
	CHECK(at == b->first_sample * 8);
	c->bytes.resize(at + n);
	CHECK(perseus_gpu_memcpy(c->h, c->bytes.data() + at, b->dev_i32, n) == 0);   // re-enters the handle from its own sink
}

// A: the virtual receiver's delivery thread against a stats reader
static void scenario_vrx_stats()
{
	perseus_vrx_config cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.struct_size = sizeof cfg;
	cfg.sample_rate = 2000000;
	cfg.drop_every = 7; cfg.swap_every = 11; cfg.timeout_every = 13;
	perseus_vrx *v = nullptr;
	CHECK(perseus_vrx_open(&v, &cfg) == 0);
	std::atomic<uint64_t> calls{0};
	auto cb = [](void *, int, void *extra) -> int { static_cast<std::atomic<uint64_t> *>(extra)->fetch_add(1); return 0; };
	CHECK(perseus_vrx_start_async_input(v, 6144, cb, &calls) == 0);
	CHECK(perseus_vrx_start_async_input(v, 6144, cb, &calls) == PERSEUS_GPU_ASYNCSTARTED);
	uint64_t last = 0;
	for (int k = 0; k < 2000 || last < 200; ++k) {
		perseus_vrx_stats s;
		CHECK(perseus_vrx_get_stats(v, &s) == 0);
		CHECK(s.delivered >= last);
		last = s.delivered;
	}
	CHECK(perseus_vrx_stop_async_input(v) == 0);
	perseus_vrx_stats s;
	CHECK(perseus_vrx_get_stats(v, &s) == 0 && s.delivered == calls.load() && s.timed_out > 0 && s.dropped_short > 0);
	CHECK(perseus_vrx_close(v) == 0);
}

// B: callbacks from the receiver's thread, watchdog inside the handle, and an application thread that polls, reads
//    statistics and syncs -- all on ONE handle
static void scenario_trampoline()
{
	perseus_gpu_config cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.struct_size = sizeof cfg;
	cfg.stream_flags = PERSEUS_GPU_OUT_INT32;
	cfg.slab_bytes = 6144 * 5;
	cfg.nslabs = 3;
	cfg.max_latency_us = 300;
	cfg.eager_gap_us = 0xFFFFFFFFu;    // slabs by size and age only: this scenario wants the watchdog busy (E runs the eager path)
	perseus_gpu *h = nullptr;
	CHECK(perseus_gpu_open(&h, &cfg) == 0);
	Collected col;
	col.h = h;
	CHECK(perseus_gpu_set_sink(h, sink, &col) == 0);

	perseus_vrx_config vc;
	memset(&vc, 0, sizeof vc);
	vc.struct_size = sizeof vc;
	vc.sample_rate = 2000000;
	vc.realtime = 1;
	vc.seed = 42;
	perseus_vrx *v = nullptr;
	CHECK(perseus_vrx_open(&v, &vc) == 0);
	CHECK(perseus_vrx_start_async_input(v, 6144, perseus_gpu_input_callback, h) == 0);
	std::atomic<bool> stop{false};
	std::thread app([&] {
		while (!stop.load()) {
			perseus_gpu_stats s;
			CHECK(perseus_gpu_get_stats(h, &s) == 0);
			CHECK(perseus_gpu_poll(h) >= 0);
			if (s.callbacks % 7 == 0) CHECK(perseus_gpu_sync(h) == 0);
			int t, st, c;
			CHECK(perseus_gpu_get_geometry(h, PERSEUS_GPU_OUT_INT32, &t, &st, &c) == 0);
		}
	});
	sleep_ms(300);
	CHECK(perseus_vrx_stop_async_input(v) == 0);
	sleep_ms(5);                       // the stream has stopped: the watchdog submits the tail on its own
	stop.store(true);
	app.join();
	perseus_vrx_stats vs;
	perseus_gpu_stats gs;
	CHECK(perseus_vrx_get_stats(v, &vs) == 0);
	CHECK(perseus_gpu_flush(h) == 0);
	CHECK(perseus_gpu_get_stats(h, &gs) == 0);
	CHECK(gs.callbacks == vs.delivered && gs.samples == vs.delivered * 1024 && gs.watchdog_submits > 0);
	// the stream the sink collected == the oracle's unpack of the synthetic stream
	std::vector<uint8_t> wire(vs.delivered * 6144), want(vs.delivered * 8192);
	CHECK(perseus_synth_fill(wire.data(), wire.size(), PERSEUS_SYNTH_RANDOM, 42, 0) == 0);
	perseus_oracle_unpack(0, wire.data(), wire.size(), want.data());
	CHECK(col.bytes.size() == want.size() && memcmp(col.bytes.data(), want.data(), want.size()) == 0);
	CHECK(perseus_vrx_close(v) == 0);
	CHECK(perseus_gpu_close(h) == 0);
}

// C: two handles on two threads through the host-pointer pipeline (slot rotation, events, statistics)
static void scenario_two_handles()
{
	auto work = [](int k) {
		perseus_gpu_config cfg;
		memset(&cfg, 0, sizeof cfg);
		cfg.struct_size = sizeof cfg;
		cfg.chunk_bytes = 48 * 100;
		cfg.stage_slots = 2 + k;
		perseus_gpu *h = nullptr;
		CHECK(perseus_gpu_open(&h, &cfg) == 0);
		std::vector<uint8_t> wire(6144 * 9 + 30), oi(wire.size() / 6 * 8), of(wire.size() / 6 * 8), want(oi.size());
		CHECK(perseus_synth_fill(wire.data(), wire.size(), PERSEUS_SYNTH_RANDOM, 7 + k, 0) == 0);
		for (int r = 0; r < 20; ++r) {
			CHECK(perseus_gpu_unpack(h, wire.data(), wire.size(), oi.data(), of.data(), PERSEUS_GPU_CHECKSUM) == (int64_t)(wire.size() / 6));
			uint64_t a, b;
			CHECK(perseus_gpu_get_checksums(h, &a, &b) == 0);
		}
		perseus_oracle_unpack(0, wire.data(), wire.size(), want.data());
		CHECK(memcmp(oi.data(), want.data(), want.size()) == 0);
		perseus_oracle_unpack(1, wire.data(), wire.size(), want.data());
		CHECK(memcmp(of.data(), want.data(), want.size()) == 0);
		perseus_gpu_seg seg = {wire.data(), wire.size(), oi.data(), nullptr};
		CHECK(perseus_gpu_unpack_batch(h, &seg, 1, PERSEUS_GPU_OUT_INT32) == (int64_t)(wire.size() / 6));
		CHECK(perseus_gpu_close(h) == 0);
	};
	std::thread a(work, 0), b(work, 1);
	a.join();
	b.join();
}

// D: close while the watchdog sleeps on a partial slab; error latching with its counters; the file sink
static void scenario_edges()
{
	perseus_gpu_config cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.struct_size = sizeof cfg;
	cfg.max_latency_us = 200000;
	perseus_gpu *h = nullptr;
	CHECK(perseus_gpu_open(&h, &cfg) == 0);
	std::vector<uint8_t> t(6144, 0x55);
	CHECK(perseus_gpu_stream_to_file(h, "/tmp/perseus_sanitize.bin") == 0);
	for (int k = 0; k < 3; ++k) perseus_gpu_input_callback(t.data(), 6144, h);
	CHECK(perseus_gpu_close(h) == 0);                    // flushes the partial slab, joins the sleeping watchdog
	FILE *f = fopen("/tmp/perseus_sanitize.bin", "rb");
	CHECK(f != nullptr);
	fseek(f, 0, SEEK_END);
	CHECK(ftell(f) == 3 * 8192);
	fclose(f);
	remove("/tmp/perseus_sanitize.bin");

	cfg.slab_bytes = 1ull << 44;                         // cannot be allocated: the callback latches the failure
	CHECK(perseus_gpu_open(&h, &cfg) == 0);
	for (int k = 0; k < 4; ++k) CHECK(perseus_gpu_input_callback(t.data(), 6144, h) == 0);
	perseus_gpu_stats s;
	CHECK(perseus_gpu_get_stats(h, &s) == 0 && s.callbacks == 0 && s.dropped_callbacks == 4 && s.dropped_bytes == 4 * 6144);   // the failing one too
	CHECK(perseus_gpu_flush(h) == PERSEUS_GPU_CUDAERR);
	CHECK(perseus_gpu_flush(h) == 0);
	CHECK(perseus_gpu_close(h) == 0);
	CHECK(perseus_gpu_close(nullptr) == PERSEUS_GPU_NULLHANDLE);
}

// E: the host sink and the file sink are served on the runtime's callback thread while the receiver's thread keeps submitting
//    slabs and an application thread reads statistics, polls and flushes
static void host_sink(const perseus_gpu_host_block *b, void *extra)
{
	Collected *c = static_cast<Collected *>(extra);
	std::lock_guard<std::mutex> lk(c->mu);
	CHECK(c->bytes.size() == b->first_sample * 8 && b->f32 && !b->i32);
	const uint8_t *p = static_cast<const uint8_t *>(b->f32);
	c->bytes.insert(c->bytes.end(), p, p + b->nsamples * 8);
}

static void scenario_host_delivery()
{
	for (uint32_t direct : {0u, 0xFFFFFFFFu}) {
		perseus_gpu_config cfg;
		memset(&cfg, 0, sizeof cfg);
		cfg.struct_size = sizeof cfg;
		cfg.stream_flags = PERSEUS_GPU_OUT_FLOAT;
		cfg.slab_bytes = 6144 * 3;
		cfg.nslabs = 2;
		cfg.max_latency_us = 500;
		cfg.direct_bytes = direct;
		perseus_gpu *h = nullptr;
		CHECK(perseus_gpu_open(&h, &cfg) == 0);
		Collected col;
		CHECK(perseus_gpu_set_host_sink(h, host_sink, &col) == 0);
		CHECK(perseus_gpu_stream_to_file(h, "/tmp/perseus_sanitize_e.bin") == 0);
		CHECK(perseus_gpu_prepare(h) == 0 && perseus_gpu_prepare(h) == 0);   // allocations, warm-up launch, helper threads: now, not in the first callback
		perseus_vrx_config vc;
		memset(&vc, 0, sizeof vc);
		vc.struct_size = sizeof vc;
		vc.sample_rate = 2000000;
		vc.realtime = 1;
		vc.seed = 9;
		perseus_vrx *v = nullptr;
		CHECK(perseus_vrx_open(&v, &vc) == 0);
		CHECK(perseus_vrx_start_async_input(v, 6144, perseus_gpu_input_callback, h) == 0);
		std::atomic<bool> stop{false};
		std::thread app([&] {
			while (!stop.load()) {
				perseus_gpu_stats s;
				CHECK(perseus_gpu_get_stats(h, &s) == 0 && s.host_blocks <= s.slabs);
				CHECK(perseus_gpu_poll(h) >= 0);
				if (s.slabs % 5 == 0) CHECK(perseus_gpu_flush(h) == 0);
			}
		});
		sleep_ms(200);
		CHECK(perseus_vrx_stop_async_input(v) == 0);
		stop.store(true);
		app.join();
		perseus_vrx_stats vs;
		perseus_gpu_stats gs;
		CHECK(perseus_vrx_get_stats(v, &vs) == 0);
		CHECK(perseus_gpu_set_host_sink(h, nullptr, nullptr) == 0);      // flushes: everything submitted has been delivered
		CHECK(perseus_gpu_get_stats(h, &gs) == 0 && gs.host_blocks == gs.slabs && gs.samples == vs.delivered * 1024);
		CHECK(perseus_gpu_close(h) == 0);
		std::vector<uint8_t> wire(vs.delivered * 6144), want(vs.delivered * 8192), file(want.size());
		CHECK(perseus_synth_fill(wire.data(), wire.size(), PERSEUS_SYNTH_RANDOM, 9, 0) == 0);
		perseus_oracle_unpack(1, wire.data(), wire.size(), want.data());
		CHECK(col.bytes.size() == want.size() && memcmp(col.bytes.data(), want.data(), want.size()) == 0);
		FILE *f = fopen("/tmp/perseus_sanitize_e.bin", "rb");
		CHECK(f != nullptr && fread(file.data(), 1, file.size() + 1, f) == want.size() && memcmp(file.data(), want.data(), want.size()) == 0);
		fclose(f);
		remove("/tmp/perseus_sanitize_e.bin");
		CHECK(perseus_vrx_close(v) == 0);
	}
}

// F: pageable buffers through the pinned bounce buffers: helper threads of two handles at once, chunks large enough to be split
static void scenario_pageable_bounce()
{
	auto work = [](int k) {
		perseus_gpu_config cfg;
		memset(&cfg, 0, sizeof cfg);
		cfg.struct_size = sizeof cfg;
		cfg.chunk_bytes = 12288 * 100;                   // 1.2 MB chunks: 4 slices of >= 256 KiB
		cfg.stage_slots = 2 + k;
		cfg.copy_threads = 3 + k;
		perseus_gpu *h = nullptr;
		CHECK(perseus_gpu_open(&h, &cfg) == 0);
		std::vector<uint8_t> wire(12288 * 530 + 30 + k), oi(wire.size() / 6 * 8), of(wire.size() / 6 * 8), want(oi.size());
		CHECK(perseus_synth_fill(wire.data(), wire.size(), PERSEUS_SYNTH_RANDOM, 70 + k, 0) == 0);
		for (int r = 0; r < 6; ++r) {
			memset(oi.data(), 0, oi.size());
			memset(of.data(), 0, of.size());
			const unsigned flags = (r & 1) ? PERSEUS_GPU_ASYNC : 0;      // pageable outputs are complete at return either way
			CHECK(perseus_gpu_unpack(h, wire.data(), wire.size(), oi.data(), of.data(), flags) == (int64_t)(wire.size() / 6));
			perseus_oracle_unpack(0, wire.data(), wire.size(), want.data());
			CHECK(memcmp(oi.data(), want.data(), want.size()) == 0);
			perseus_oracle_unpack(1, wire.data(), wire.size(), want.data());
			CHECK(memcmp(of.data(), want.data(), want.size()) == 0);
		}
		CHECK(perseus_gpu_close(h) == 0);
	};
	std::thread a(work, 0), b(work, 1);
	a.join();
	b.join();
}

// H: randomised: a receiver thread calling back as fast as it can while the application thread switches the host sink off and on,
//    polls, flushes, reads statistics, prepares -- three threads (with the delivery thread) on one handle, many short streams
struct Spans {
	std::mutex mu;
	std::vector<std::pair<uint64_t, uint64_t>> got;      // [first, end) of every block, merged when contiguous
	uint64_t words = 0, bad = 0;
	const uint32_t *want = nullptr;
};

static void span_sink(const perseus_gpu_host_block *b, void *extra)
{
	Spans *sp = static_cast<Spans *>(extra);
	std::lock_guard<std::mutex> lk(sp->mu);
	const uint32_t *w = static_cast<const uint32_t *>(b->i32);
	for (uint64_t k = 0; k < 2 * b->nsamples; ++k) sp->bad += w[k] != sp->want[2 * b->first_sample + k];
	sp->words += 2 * b->nsamples;
	if (!sp->got.empty() && sp->got.back().second == b->first_sample) sp->got.back().second += b->nsamples;
	else sp->got.emplace_back(b->first_sample, b->first_sample + b->nsamples);
}

static void scenario_random_toggling()
{
	unsigned rs = 12345;
	auto rnd = [&rs](unsigned n) { rs = rs * 1103515245u + 12345u; return (rs >> 16) % n; };
	for (int it = 0; it < 40; ++it) {
		const int ntransfers = 50 + (int)rnd(400);
		std::vector<uint8_t> wire((size_t)ntransfers * 6144), want((size_t)ntransfers * 8192);
		CHECK(perseus_synth_fill(wire.data(), wire.size(), PERSEUS_SYNTH_RANDOM, 1000 + it, 0) == 0);
		perseus_oracle_unpack(0, wire.data(), wire.size(), want.data());
		perseus_gpu_config cfg;
		memset(&cfg, 0, sizeof cfg);
		cfg.struct_size = sizeof cfg;
		cfg.stream_flags = PERSEUS_GPU_OUT_INT32;
		cfg.slab_bytes = 6144 * (1 + rnd(9)) + 48 * rnd(50);
		cfg.nslabs = 2 + rnd(4);
		cfg.nstreams = 1 + rnd(3);
		cfg.max_latency_us = rnd(2) ? 200 : 0;
		cfg.direct_bytes = rnd(2) ? 0 : 0xFFFFFFFFu;
		cfg.eager_gap_us = rnd(2) ? 20 : 0xFFFFFFFFu;
		perseus_gpu *h = nullptr;
		CHECK(perseus_gpu_open(&h, &cfg) == 0);
		Spans sp;
		sp.want = reinterpret_cast<const uint32_t *>(want.data());
		CHECK(perseus_gpu_set_host_sink(h, span_sink, &sp) == 0);
		if (rnd(2)) CHECK(perseus_gpu_prepare(h) == 0);
		std::atomic<bool> done{false};
		std::thread receiver([&] {
			for (int k = 0; k < ntransfers; ++k) {
				perseus_gpu_input_callback(wire.data() + (size_t)k * 6144, 6144, h);
				if (k % 37 == 0) std::this_thread::yield();
			}
			done.store(true);
		});
		bool on = true;
		unsigned toggles = 0;
		while (!done.load()) {
			switch (rnd(6)) {
			case 0: CHECK(perseus_gpu_set_host_sink(h, on ? nullptr : span_sink, &sp) == 0); on = !on; ++toggles; break;
			case 1: CHECK(perseus_gpu_poll(h) >= 0); break;
			case 2: CHECK(perseus_gpu_flush(h) == 0); break;
			case 3: { perseus_gpu_stats st; CHECK(perseus_gpu_get_stats(h, &st) == 0 && st.host_blocks <= st.slabs); break; }
			case 4: CHECK(perseus_gpu_prepare(h) == 0); break;
			default: std::this_thread::yield();
			}
		}
		receiver.join();
		CHECK(perseus_gpu_flush(h) == 0);
		perseus_gpu_stats st;
		CHECK(perseus_gpu_get_stats(h, &st) == 0 && st.callbacks == (uint64_t)ntransfers && st.samples == (uint64_t)ntransfers * 1024);
		CHECK(perseus_gpu_close(h) == 0);
		// whatever the interleaving: every block carried the samples its first_sample says, blocks came in stream order, and with
		// no toggling at all the sink saw the whole stream
		CHECK(sp.bad == 0);
		for (size_t k = 1; k < sp.got.size(); ++k) CHECK(sp.got[k].first > sp.got[k - 1].second);
		if (toggles == 0) CHECK(sp.got.size() == 1 && sp.got[0].first == 0 && sp.got[0].second == (uint64_t)ntransfers * 1024);
	}
}

// I: error paths under the sanitizers: the n-th device / pinned allocation fails, for every n that a short life of a handle reaches
//    (open, prepare, streaming with host sink + file, a bulk call on pageable memory, a batched plan, a probe).  Whatever fails must
//    fail cleanly: an error code, no leak (LeakSanitizer at exit), no use of what was not allocated, and a close that still works.
static void scenario_allocation_failures()
{
	std::vector<uint8_t> wire(6144 * 40 + 30), oi(wire.size() / 6 * 8), of(wire.size() / 6 * 8);
	CHECK(perseus_synth_fill(wire.data(), wire.size(), PERSEUS_SYNTH_RANDOM, 77, 0) == 0);
	int failed_calls = 0;
	for (long n = 1; n <= 70; ++n) {
		fake_cuda_fail_alloc_in(n);
		perseus_gpu_config cfg;
		memset(&cfg, 0, sizeof cfg);
		cfg.struct_size = sizeof cfg;
		cfg.stream_flags = PERSEUS_GPU_OUT_FLOAT;
		cfg.slab_bytes = 6144 * 2;
		cfg.nslabs = 2;
		cfg.chunk_bytes = 12288 * 4;
		cfg.copy_threads = 2;
		perseus_gpu *h = nullptr;
		if (perseus_gpu_open(&h, &cfg) != 0) { ++failed_calls; CHECK(h == nullptr); continue; }
		Collected col;
		failed_calls += perseus_gpu_set_host_sink(h, host_sink, &col) != 0;
		failed_calls += perseus_gpu_stream_to_file(h, "/tmp/perseus_sanitize_i.bin") != 0;
		failed_calls += perseus_gpu_prepare(h) != 0;
		for (int k = 0; k < 7; ++k) perseus_gpu_input_callback(wire.data() + k * 6144, 6144, h);
		failed_calls += perseus_gpu_flush(h) != 0;
		failed_calls += perseus_gpu_unpack(h, wire.data(), wire.size(), oi.data(), of.data(), PERSEUS_GPU_CHECKSUM) < 0;
		perseus_gpu_seg seg = {wire.data(), wire.size(), oi.data(), nullptr};
		failed_calls += perseus_gpu_unpack_batch(h, &seg, 1, PERSEUS_GPU_OUT_INT32) < 0;
		double gbs = 0;
		failed_calls += perseus_gpu_probe_hbm(h, PERSEUS_GPU_PROBE_COPY, 1 << 20, 1, &gbs) != 0;
		failed_calls += perseus_gpu_probe_pcie(h, PERSEUS_GPU_PCIE_DUPLEX, 1 << 20, 0, 1, &gbs, &gbs) != 0;
		void *p = perseus_gpu_dev_alloc(h, 4096);
		if (p) CHECK(perseus_gpu_dev_free(h, p) == 0);
		else ++failed_calls;
		perseus_gpu_close(h);                        // may report the latched failure; must free everything either way
	}
	fake_cuda_fail_alloc_in(0);
	remove("/tmp/perseus_sanitize_i.bin");
	CHECK(failed_calls >= 30);                       // the injected failures did land in API calls, not only in the void
}

// G: the non-temporal copy (16/32/64-byte stores picked at run time; PERSEUS_GPU_NT_COPY forces one) at every destination and
//    source phase and every length around its block sizes: exactly the bytes asked for, nothing outside
static void scenario_nontemporal_copy()
{
	std::vector<uint8_t> src(4096 + 128), dst(4096 + 256), want(dst.size());
	for (size_t i = 0; i < src.size(); ++i) src[i] = (uint8_t)(i * 131 + 7);
	for (size_t doff = 0; doff < 70; doff += 3)
		for (size_t soff = 0; soff < 5; ++soff)
			for (size_t n : {0u, 1u, 15u, 63u, 64u, 65u, 127u, 128u, 200u, 255u, 256u, 257u, 510u, 511u, 767u, 1000u, 4000u}) {
				memset(dst.data(), 0xEE, dst.size());
				memset(want.data(), 0xEE, want.size());
				memcpy(want.data() + 64 + doff, src.data() + soff, n);
				pg::copy_nontemporal(dst.data() + 64 + doff, src.data() + soff, n);
				CHECK(memcmp(dst.data(), want.data(), dst.size()) == 0);
			}
}

int main()
{
	scenario_vrx_stats();
	scenario_trampoline();
	scenario_two_handles();
	scenario_edges();
	scenario_host_delivery();
	scenario_pageable_bounce();
	scenario_nontemporal_copy();
	scenario_random_toggling();
	scenario_allocation_failures();
	printf("host_stress: all scenarios passed\n");
	return 0;
}
