/* Hand-off latency of the streaming path, measured from C so no interpreter is in the way.
 *
 * For slab sizes from one 6144-byte transfer to 8 MiB, on both slab routes (perseus_gpu_config.direct_bytes: the kernel reads
 * the pinned slab over the link itself / the copy engine moves it into HBM first):
 *   device : perseus_gpu_input_callback(last transfer of the slab) + perseus_gpu_flush      -> samples resident in HBM
 *   host   : perseus_gpu_input_callback(last transfer of the slab) -> the host sink is called -> samples in host memory
 * The slab's earlier transfers are pushed before the clock starts (their copy is the same on both routes).
 * One JSON line per (mode, route, slab size): median / p95 / min in microseconds over REPS slabs.
 *
 *   gcc -O2 -std=c11 tools/latency_probe.c -Iinclude -Llibperseus-sdr_b200/lib -lperseus_gpu -o tools/build/latency_probe
 *   LD_LIBRARY_PATH=libperseus-sdr_b200/lib tools/build/latency_probe > gpurun_out/latency.jsonl
 */
#define _POSIX_C_SOURCE 200809L
#include "perseus-gpu.h"

#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define XFER 6144
#define REPS 300

static double now_us(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

static _Atomic unsigned long long g_blocks;
static _Atomic unsigned g_word;

static void host_sink(const perseus_gpu_host_block *b, void *extra)
{
	(void)extra;
	/* touch the block like a consumer would: its last word has arrived */
	atomic_store(&g_word, ((const unsigned *)b->f32)[2 * b->nsamples - 1]);
	atomic_fetch_add(&g_blocks, 1);
}

static int cmp(const void *a, const void *b)
{
	const double x = *(const double *)a, y = *(const double *)b;
	return x < y ? -1 : x > y;
}

static int run(int host_mode, unsigned direct, size_t slab)
{
	perseus_gpu_config cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.struct_size = sizeof cfg;
	cfg.stream_flags = PERSEUS_GPU_OUT_FLOAT;
	cfg.slab_bytes = slab;
	cfg.nslabs = 4;
	cfg.nstreams = 2;
	cfg.max_latency_us = 0xFFFFFFFFu;
	cfg.direct_bytes = direct;
	perseus_gpu *h = NULL;
	if (perseus_gpu_open(&h, &cfg) < 0) { fprintf(stderr, "open: %s\n", perseus_gpu_errorstr()); return 1; }
	if (host_mode && perseus_gpu_set_host_sink(h, host_sink, NULL) < 0) { fprintf(stderr, "sink: %s\n", perseus_gpu_errorstr()); return 1; }
	static unsigned char xfer[XFER];
	perseus_synth_fill(xfer, XFER, PERSEUS_SYNTH_RANDOM, PERSEUS_SYNTH_SEED, 0);
	const size_t per_slab = slab / XFER;
	static double lat[REPS];
	for (int r = -20; r < REPS; ++r) {          /* negative r: warm-up (slab allocation, page touching) */
		const unsigned long long seen = atomic_load(&g_blocks);
		for (size_t k = 0; k + 1 < per_slab; ++k) perseus_gpu_input_callback(xfer, XFER, h);
		const double t0 = now_us();
		perseus_gpu_input_callback(xfer, XFER, h);   /* fills the slab: submitted inside this call */
		if (host_mode) {
			while (atomic_load(&g_blocks) == seen) { }
		} else if (perseus_gpu_flush(h) < 0) {
			fprintf(stderr, "flush: %s\n", perseus_gpu_errorstr());
			return 1;
		}
		const double t1 = now_us();
		if (r >= 0) lat[r] = t1 - t0;
		if (host_mode) perseus_gpu_flush(h);
	}
	qsort(lat, REPS, sizeof lat[0], cmp);
	printf("{\"mode\": \"%s\", \"route\": \"%s\", \"slab_bytes\": %zu, \"median_us\": %.1f, \"p95_us\": %.1f, \"min_us\": %.1f, \"reps\": %d}\n",
	       host_mode ? "host_sink" : "device_flush", direct == 0xFFFFFFFFu ? "staged" : "direct", slab, lat[REPS / 2], lat[REPS * 95 / 100], lat[0], REPS);
	fflush(stdout);
	return perseus_gpu_close(h) < 0;
}

/* The DEFAULT handle (8 MiB slabs, 50 ms bound, eager_gap_us = 100) fed like a real receiver: one 6144-byte transfer every
 * `period_us`.  Latency of each transfer = callback start -> host sink sees its block (nobody flushes or polls). */
static int run_paced(unsigned period_us)
{
	perseus_gpu_config cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.struct_size = sizeof cfg;
	cfg.stream_flags = PERSEUS_GPU_OUT_FLOAT;
	perseus_gpu *h = NULL;
	if (perseus_gpu_open(&h, &cfg) < 0 || perseus_gpu_set_host_sink(h, host_sink, NULL) < 0) { fprintf(stderr, "open: %s\n", perseus_gpu_errorstr()); return 1; }
	static unsigned char xfer[XFER];
	perseus_synth_fill(xfer, XFER, PERSEUS_SYNTH_RANDOM, PERSEUS_SYNTH_SEED, 0);
	perseus_gpu_input_callback(xfer, XFER, h);
	perseus_gpu_flush(h);
	static double lat[REPS];
	for (int r = -10; r < REPS; ++r) {
		const double until = now_us() + period_us;
		while (now_us() < until) { }
		const unsigned long long seen = atomic_load(&g_blocks);
		const double t0 = now_us();
		perseus_gpu_input_callback(xfer, XFER, h);
		const double t1 = now_us();
		while (atomic_load(&g_blocks) == seen && now_us() - t0 < 2e5) { }
		if (atomic_load(&g_blocks) == seen) { fprintf(stderr, "paced: a transfer was not delivered within 0.2 s\n"); return 1; }
		if (r >= 0) lat[r] = now_us() - t0;
		(void)t1;
	}
	qsort(lat, REPS, sizeof lat[0], cmp);
	perseus_gpu_stats st;
	perseus_gpu_get_stats(h, &st);
	printf("{\"mode\": \"paced_default_handle_host_sink\", \"period_us\": %u, \"median_us\": %.1f, \"p95_us\": %.1f, \"min_us\": %.1f, \"slabs\": %llu, "
	       "\"callbacks\": %llu, \"watchdog_submits\": %llu}\n", period_us, lat[REPS / 2], lat[REPS * 95 / 100], lat[0],
	       (unsigned long long)st.slabs, (unsigned long long)st.callbacks, (unsigned long long)st.watchdog_submits);
	fflush(stdout);
	return perseus_gpu_close(h) < 0;
}

/* How long the FIRST callback of a handle takes -- on a real receiver that is time spent on libperseus-sdr's poll thread while
 * the ring of 8 transfers fills up -- with and without perseus_gpu_prepare(); default handle, host sink set. */
static int run_first(int prepare)
{
	perseus_gpu_config cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.struct_size = sizeof cfg;
	cfg.stream_flags = PERSEUS_GPU_OUT_FLOAT;
	perseus_gpu *h = NULL;
	if (perseus_gpu_open(&h, &cfg) < 0 || perseus_gpu_set_host_sink(h, host_sink, NULL) < 0) { fprintf(stderr, "open: %s\n", perseus_gpu_errorstr()); return 1; }
	static unsigned char xfer[XFER];
	perseus_synth_fill(xfer, XFER, PERSEUS_SYNTH_RANDOM, PERSEUS_SYNTH_SEED, 0);
	double tp = 0.0;
	if (prepare) {
		const double t = now_us();
		if (perseus_gpu_prepare(h) < 0) { fprintf(stderr, "prepare: %s\n", perseus_gpu_errorstr()); return 1; }
		tp = now_us() - t;
	}
	double t[3];
	for (int k = 0; k < 3; ++k) {
		const double t0 = now_us();
		perseus_gpu_input_callback(xfer, XFER, h);
		t[k] = now_us() - t0;
		const double until = now_us() + 512;
		while (now_us() < until) { }
	}
	printf("{\"mode\": \"first_callbacks\", \"prepared\": %s, \"prepare_us\": %.0f, \"callback_us\": [%.1f, %.1f, %.1f]}\n", prepare ? "true" : "false", tp,
	       t[0], t[1], t[2]);
	return perseus_gpu_close(h) < 0;
}

int main(int argc, char **argv)
{
	if (argc > 1 && !strcmp(argv[1], "first")) return run_first(0) || run_first(1) || run_first(0) || run_first(1);
	if (argc > 1 && !strcmp(argv[1], "paced")) return run_paced(512) || run_paced(10779);   /* 2 MS/s and 95 kS/s transfer periods */
	static const size_t slabs[] = {XFER, 2 * XFER, 8 * XFER, 32 * XFER, 128 * XFER, 512 * XFER, 1365 * XFER};
	for (int host_mode = 0; host_mode < 2; ++host_mode)
		for (size_t i = 0; i < sizeof slabs / sizeof slabs[0]; ++i)
			for (int staged = 0; staged < 2; ++staged)
				if (run(host_mode, staged ? 0xFFFFFFFFu : 0x7FFFFFFFu, slabs[i])) return 1;
	return 0;
}
