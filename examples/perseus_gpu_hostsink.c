/* perseus_gpu_hostsink — a consumer that stays on the CPU.
 *
 * The reference's callbacks have the unpacked samples in hand where they fwrite them (examples/perseustest.c:457,499); an
 * application that does something else with them there -- a level meter, a demodulator, a network sender -- keeps that code and
 * moves it into a host sink: perseus_gpu_set_host_sink() hands it the same {int32 I, int32 Q} / {float I, float Q} samples in
 * blocks, in stream order, a few tens of microseconds after each transfer arrived, while the unpack itself runs on the B200.
 *
 * This example is a level meter: per block it accumulates the exact 64-bit sum of I and of Q and the peak magnitude of the int32
 * stream, and prints the totals (tests/test_examples.py recomputes them from the oracle's unpack of the same wire stream).
 * The receiver is the virtual one, paced in real time like a device (-t seconds) or replayed as fast as possible (-N transfers).
 *
 *   gcc -std=c99 -I include examples/perseus_gpu_hostsink.c -L libperseus-sdr_b200/lib -lperseus_gpu -o perseus_gpu_hostsink
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "perseus-gpu.h"

struct meter {
	long long sum_i, sum_q;
	unsigned long long nsamples, blocks, out_of_order;
	unsigned peak;                     /* largest |value| seen, as the unsigned magnitude of the int32 */
};

/* Runs on the handle's delivery thread; the block is valid only during the call. */
static void level_meter(const perseus_gpu_host_block *b, void *extra)
{
	struct meter *m = (struct meter *)extra;
	const int32_t *iq = (const int32_t *)b->i32;
	uint64_t k;
	if (b->first_sample != m->nsamples) m->out_of_order++;
	for (k = 0; k < b->nsamples; k++) {
		const int32_t i = iq[2 * k], q = iq[2 * k + 1];
		const unsigned ai = i < 0 ? 0u - (unsigned)i : (unsigned)i, aq = q < 0 ? 0u - (unsigned)q : (unsigned)q;
		m->sum_i += i;
		m->sum_q += q;
		if (ai > m->peak) m->peak = ai;
		if (aq > m->peak) m->peak = aq;
	}
	m->nsamples += b->nsamples;
	m->blocks++;
}

int main(int argc, char **argv)
{
	int sr = 95000, seconds = 0, device = 0, opt, rc;
	unsigned long long ntransfers = 1000;
	struct meter m;
	perseus_gpu *gpu = NULL;
	perseus_vrx *rx = NULL;
	perseus_gpu_config gcfg;
	perseus_vrx_config vcfg;
	perseus_gpu_stats gs;

	while ((opt = getopt(argc, argv, "s:t:N:g:h")) != -1) {
		switch (opt) {
		case 's': sr = atoi(optarg); break;
		case 't': seconds = atoi(optarg); break;
		case 'N': ntransfers = strtoull(optarg, NULL, 10); break;
		case 'g': device = atoi(optarg); break;
		default:
			fprintf(stderr, "Usage: perseus_gpu_hostsink [-s rate] [-t seconds | -N transfers] [-g device]\n");
			return opt == 'h' ? 0 : 2;
		}
	}
	memset(&m, 0, sizeof m);
	memset(&gcfg, 0, sizeof gcfg);
	gcfg.struct_size = sizeof gcfg;
	gcfg.device = device;
	gcfg.stream_flags = PERSEUS_GPU_OUT_INT32;
	gcfg.slab_bytes = 1u << 20;
	if (perseus_gpu_open(&gpu, &gcfg) < 0 || perseus_gpu_set_host_sink(gpu, level_meter, &m) < 0 || perseus_gpu_prepare(gpu) < 0) {
		fprintf(stderr, "perseus_gpu_open: %s\n", perseus_gpu_errorstr());
		return 1;
	}
	memset(&vcfg, 0, sizeof vcfg);
	vcfg.struct_size = sizeof vcfg;
	vcfg.sample_rate = sr;
	vcfg.realtime = seconds > 0;
	if (perseus_vrx_open(&rx, &vcfg) < 0) {
		fprintf(stderr, "perseus_vrx_open: %s\n", perseus_gpu_errorstr());
		perseus_gpu_close(gpu);
		return 1;
	}
	if (seconds > 0) {
		rc = perseus_vrx_start_async_input(rx, 6144, perseus_gpu_input_callback, gpu);
		if (rc == 0) {
			sleep((unsigned)seconds);
			rc = perseus_vrx_stop_async_input(rx);
		}
	} else {
		rc = perseus_vrx_run(rx, 6144, perseus_gpu_input_callback, gpu, ntransfers);
	}
	if (rc < 0 || perseus_gpu_flush(gpu) < 0) {          /* after the flush every block has been through the sink */
		fprintf(stderr, "streaming failed: %s\n", perseus_gpu_errorstr());
		return 1;
	}
	perseus_gpu_get_stats(gpu, &gs);
	printf("samples %llu blocks %llu out_of_order %llu sum_i %lld sum_q %lld peak %u\n", m.nsamples, m.blocks, m.out_of_order, m.sum_i, m.sum_q,
	       m.peak);
	fprintf(stderr, "GPU: %llu transfers in %llu slabs, %llu blocks delivered to the host sink\n", (unsigned long long)gs.callbacks,
	        (unsigned long long)gs.slabs, (unsigned long long)gs.host_blocks);
	perseus_vrx_close(rx);
	return perseus_gpu_close(gpu) < 0;
}
