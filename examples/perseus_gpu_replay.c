/* perseus_gpu_replay — the shape of the reference's examples/perseustest.c with the two things this repository
 * provides swapped in:
 *   - the USB receiver  -> the virtual receiver (perseus_vrx_*: same 8-slot ring, same buffersize checks,
 *                          same in-order delivery as perseus_start_async_input + perseus-in.c)
 *   - the CPU callbacks -> perseus_gpu_input_callback (user_data_callback_c_u / _c_f on a B200)
 * Everything else reads like perseustest.c:93-409: parse -s -n -b -t -o -p, start the async input with
 * nb*bs-byte buffers, let it run, stop, print the kS/s line perseus_stop_async_input prints
 * (perseus-sdr.c:719-722).  The output file is byte-identical to what `perseustest -o file [-p]` writes for
 * the same wire stream (tests/test_examples.py checks that against the reference's own callbacks).
 *
 * Plain C99; links only against libperseus_gpu.so:
 *   gcc -std=c99 -I include examples/perseus_gpu_replay.c -L libperseus-sdr_b200/lib -lperseus_gpu -o perseus_gpu_replay
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "perseus-gpu.h"

static void usage(void)
{
	int rates[16], i;
	fprintf(stderr, "Usage: perseus_gpu_replay [options]\n-s ............ sample rate (");
	if (perseus_vrx_get_sampling_rates(rates, 16) == 0)
		for (i = 0; rates[i]; i++) fprintf(stderr, " %d", rates[i]);
	fprintf(stderr, ")\n"
	        "-n ............ number of buffers (default 6)\n"
	        "-b ............ buffer size in bytes (default 1024)\n"
	        "-t ............ duration in seconds, paced at the sample rate (default: not paced, see -N)\n"
	        "-N ............ number of transfers to replay as fast as possible (default 1000)\n"
	        "-o ............ output file name (default perseusdata; - = standard output, as perseustest)\n"
	        "-p ............ I/Q samples emitted as floating point instead of 32 bit integers\n"
	        "-g ............ CUDA device ordinal (default 0)\n"
	        "-h ............ this help\n");
}

int main(int argc, char **argv)
{
	int sr = 95000, nb = 6, bs = 1024, seconds = 0, use_float = 0, device = 0, opt;
	unsigned long long ntransfers = 1000;
	const char *fname = "perseusdata";

	while ((opt = getopt(argc, argv, "s:n:b:t:N:o:pg:h")) != -1) {
		switch (opt) {
		case 's': sr = atoi(optarg); break;
		case 'n': nb = atoi(optarg); break;
		case 'b': bs = atoi(optarg); break;
		case 't': seconds = atoi(optarg); break;
		case 'N': ntransfers = strtoull(optarg, NULL, 10); break;
		case 'o': fname = optarg; break;
		case 'p': use_float = 1; break;
		case 'g': device = atoi(optarg); break;
		default: usage(); return opt == 'h' ? 0 : 2;
		}
	}

	perseus_gpu *gpu = NULL;
	perseus_gpu_config gcfg;
	memset(&gcfg, 0, sizeof gcfg);
	gcfg.struct_size = sizeof gcfg;
	gcfg.device = device;
	gcfg.stream_flags = use_float ? PERSEUS_GPU_OUT_FLOAT : PERSEUS_GPU_OUT_INT32;
	gcfg.slab_bytes = 1u << 20;
	if (perseus_gpu_open(&gpu, &gcfg) < 0) {
		fprintf(stderr, "perseus_gpu_open: %s\n", perseus_gpu_errorstr());
		return 1;
	}
	if (perseus_gpu_stream_to_file(gpu, fname) < 0) {
		fprintf(stderr, "cannot write %s: %s\n", fname, perseus_gpu_errorstr());
		perseus_gpu_close(gpu);
		return 1;
	}

	if (perseus_gpu_prepare(gpu) < 0) {   /* allocate and warm up now, not inside the first callback on the receiver's thread */
		fprintf(stderr, "perseus_gpu_prepare: %s\n", perseus_gpu_errorstr());
		perseus_gpu_close(gpu);
		return 1;
	}

	perseus_vrx *rx = NULL;
	perseus_vrx_config vcfg;
	memset(&vcfg, 0, sizeof vcfg);
	vcfg.struct_size = sizeof vcfg;
	vcfg.sample_rate = sr;
	vcfg.realtime = seconds > 0;
	if (perseus_vrx_open(&rx, &vcfg) < 0) {
		fprintf(stderr, "perseus_vrx_open: %s\n", perseus_gpu_errorstr());
		perseus_gpu_close(gpu);
		return 1;
	}
	fprintf(stderr, "Sample rate %d S/s, buffers of %d bytes, output %s (%s)\n", perseus_vrx_get_sampling_rate(rx), nb * bs, fname,
	        use_float ? "float" : "int32");

	int rc;
	if (seconds > 0) {
		/* as perseustest.c:349-376: start, sleep, stop */
		rc = perseus_vrx_start_async_input(rx, (uint32_t)(nb * bs), perseus_gpu_input_callback, gpu);
		if (rc == 0) {
			sleep((unsigned)seconds);
			rc = perseus_vrx_stop_async_input(rx);
		}
	} else {
		rc = perseus_vrx_run(rx, (uint32_t)(nb * bs), perseus_gpu_input_callback, gpu, ntransfers);
	}
	if (rc < 0) {
		fprintf(stderr, "start async input error: %s\n", perseus_gpu_errorstr());
		perseus_vrx_close(rx);
		perseus_gpu_close(gpu);
		return 1;
	}

	perseus_vrx_stats vs;
	perseus_gpu_stats gs;
	perseus_vrx_get_stats(rx, &vs);
	if (perseus_gpu_flush(gpu) < 0) fprintf(stderr, "perseus_gpu_flush: %s\n", perseus_gpu_errorstr());
	perseus_gpu_get_stats(gpu, &gs);
	/* perseus-sdr.c:719-722 */
	fprintf(stderr, "Elapsed time: %f s - kSamples read: %llu - Rate: %.1f kS/s\n", vs.elapsed_s,
	        (unsigned long long)(vs.bytes_received / 6000), vs.ksamples_per_s);
	fprintf(stderr, "GPU: %llu samples in %llu slabs, %llu kernel launches, %llu slab stalls\n", (unsigned long long)gs.samples,
	        (unsigned long long)gs.slabs, (unsigned long long)gs.kernel_launches, (unsigned long long)gs.stalls);

	perseus_vrx_close(rx);
	rc = perseus_gpu_close(gpu);
	if (rc < 0) {
		fprintf(stderr, "perseus_gpu_close: %s\n", perseus_gpu_errorstr());
		return 1;
	}
	fprintf(stderr, "Bye\n");
	return 0;
}
