/* perseus_gpu_shard — a recording sharded by contiguous transfer ranges over every B200 in the box, from ONE host
 * thread, in plain C against include/perseus-gpu.h (SURVEY.md §8e / BASELINE config 4).
 *
 * There is nothing to exchange between shards: each perseus_gpu handle owns one device, generates (or would receive)
 * only the bytes of its own transfer range, unpacks them with one asynchronous launch, and reports a checksum whose
 * per-shard values simply add up (mod 2^64) to the checksum of the whole recording.  No NCCL, no peer access.
 *
 *   gcc -std=c99 -I include examples/perseus_gpu_shard.c -L libperseus-sdr_b200/lib -lperseus_gpu -o perseus_gpu_shard
 *   ./perseus_gpu_shard [-n transfers] [-g gpus] [-r reps] [-p]        (-p: float output instead of int32)
 */
#define _POSIX_C_SOURCE 200809L
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "perseus-gpu.h"

#define MAXGPU 16
#define XFER 6144u

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define CHECK(call)                                                                  \
	do {                                                                             \
		if ((call) < 0) {                                                            \
			fprintf(stderr, "%s: %s\n", #call, perseus_gpu_errorstr());              \
			return 1;                                                                \
		}                                                                            \
	} while (0)

int main(int argc, char **argv)
{
	uint64_t total = 174762;          /* BASELINE config 2 size by default; config 4 is -n 11184810 */
	int ngpu = perseus_gpu_device_count(), reps = 20, use_float = 0, opt;
	while ((opt = getopt(argc, argv, "n:g:r:ph")) != -1) {
		switch (opt) {
		case 'n': total = strtoull(optarg, NULL, 10); break;
		case 'g': ngpu = atoi(optarg); break;
		case 'r': reps = atoi(optarg); break;
		case 'p': use_float = 1; break;
		default: fprintf(stderr, "usage: %s [-n transfers] [-g gpus] [-r reps] [-p]\n", argv[0]); return opt == 'h' ? 0 : 2;
		}
	}
	if (ngpu < 1) {
		fprintf(stderr, "no usable GPU: %s\n", perseus_gpu_errorstr());
		return 1;
	}
	if (ngpu > MAXGPU) ngpu = MAXGPU;

	perseus_gpu *h[MAXGPU];
	void *d_in[MAXGPU], *d_out[MAXGPU];
	uint64_t first[MAXGPU], count[MAXGPU];
	const unsigned flags = use_float ? PERSEUS_GPU_OUT_FLOAT : PERSEUS_GPU_OUT_INT32;

	for (int g = 0; g < ngpu; g++) {
		perseus_gpu_config cfg;
		memset(&cfg, 0, sizeof cfg);
		cfg.struct_size = sizeof cfg;
		cfg.device = g;
		CHECK(perseus_gpu_open(&h[g], &cfg));
		CHECK(perseus_gpu_shard_range(total, ngpu, g, &first[g], &count[g]));
		d_in[g] = perseus_gpu_dev_alloc(h[g], count[g] * XFER + 16);
		d_out[g] = perseus_gpu_dev_alloc(h[g], count[g] * 8192 + 16);
		if (!d_in[g] || !d_out[g]) {
			fprintf(stderr, "device %d: %s\n", g, perseus_gpu_errorstr());
			return 1;
		}
		/* the shard's own byte range of the recording, generated where it is needed */
		CHECK(perseus_gpu_generate(h[g], d_in[g], count[g] * XFER, PERSEUS_SYNTH_RANDOM, PERSEUS_SYNTH_SEED, first[g] * XFER));
	}

	/* warm-up + correctness: every shard against the independent per-sample kernel */
	for (int g = 0; g < ngpu; g++)
		CHECK(perseus_gpu_unpack(h[g], d_in[g], count[g] * XFER, use_float ? NULL : d_out[g], use_float ? d_out[g] : NULL, flags | PERSEUS_GPU_ASYNC));
	uint64_t sum = 0;
	for (int g = 0; g < ngpu; g++) {
		uint64_t bad = 0, where = 0, s = 0;
		CHECK(perseus_gpu_sync(h[g]));
		CHECK(perseus_gpu_verify(h[g], d_in[g], count[g] * XFER, use_float ? NULL : d_out[g], use_float ? d_out[g] : NULL, flags, &bad, &where));
		CHECK(perseus_gpu_checksum(h[g], d_out[g], count[g] * 2048, first[g] * 2048, &s));
		sum += s;                                          /* checksum of checksums */
	}

	/* timed: all devices launched back to back from this one thread, then joined */
	const double t0 = now_s();
	for (int r = 0; r < reps; r++)
		for (int g = 0; g < ngpu; g++)
			CHECK(perseus_gpu_unpack(h[g], d_in[g], count[g] * XFER, use_float ? NULL : d_out[g], use_float ? d_out[g] : NULL, flags | PERSEUS_GPU_ASYNC));
	for (int g = 0; g < ngpu; g++) CHECK(perseus_gpu_sync(h[g]));
	const double dt = (now_s() - t0) / reps;

	const double samples = (double)total * 1024.0;
	printf("{\"gpus\": %d, \"transfers\": %" PRIu64 ", \"format\": \"%s\", \"ms_per_pass\": %.4f, \"msamples_per_s\": %.1f, "
	       "\"hbm_gbs_per_gpu\": %.1f, \"recording_checksum\": \"%016" PRIx64 "\"}\n",
	       ngpu, total, use_float ? "float" : "int32", dt * 1e3, samples / dt / 1e6, 14.0 * samples / dt / 1e9 / ngpu, sum);

	for (int g = 0; g < ngpu; g++) {
		perseus_gpu_dev_free(h[g], d_in[g]);
		perseus_gpu_dev_free(h[g], d_out[g]);
		CHECK(perseus_gpu_close(h[g]));
	}
	return 0;
}
