/* perseus_gpu_libperseus — INTEGRATION.md §1 as a program: the reference's examples/perseustest.c flow, written against
 * the reference's OWN header and library, with exactly one thing changed: the callback handed to
 * perseus_start_async_input() is perseus_gpu_input_callback (extra = perseus_gpu*) instead of
 * user_data_callback_c_u / _c_f (perseustest.c:349,354), and the FILE* the examples open becomes
 * perseus_gpu_stream_to_file().
 *
 *     perseus_init -> perseus_open -> perseus_firmware_download -> perseus_set_sampling_rate      perseustest.c:188-266
 *     perseus_start_async_input(descr, nb*bs, perseus_gpu_input_callback, gpu)                    perseustest.c:349
 *     ... perseus_stop_async_input -> perseus_close -> perseus_exit                               perseustest.c:381-405
 *
 * It needs /root/reference/perseus-sdr.h to compile, so it is built by oracle/Makefile into oracle/_ref/ (next to the
 * reference library it links: libperseus_sdr_ref.so = the reference's sources, unmodified, over a fake libusb).  The only
 * lines a real deployment would not have are the two that plug in the synthetic receiver (fakeusb_plug): with real
 * hardware and the real libusb the rest is unchanged.
 *
 *   gcc -std=gnu99 -I include -I /root/reference -I oracle -I oracle/fakeusb tests/integration/perseus_gpu_libperseus.c \
 *       -L oracle/_ref -lperseus_sdr_ref -L libperseus-sdr_b200/lib -lperseus_gpu
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "perseus-gpu.h"          /* either order works: the header re-declares nothing of perseus-sdr.h */
#include "perseus-sdr.h"
#include "fakeusb.h"              /* test infrastructure: the synthetic receiver behind the fake libusb */

int fakeusb_plug(const fakeusb_config *cfg);
void fakeusb_get_state(fakeusb_state *out);

int main(int argc, char **argv)
{
	int sr = 95000, nb = 6, bs = 1024, use_float = 0, opt;
	unsigned long long ntransfers = 100;
	const char *fname = "perseusdata";
	while ((opt = getopt(argc, argv, "s:n:b:t:o:p")) != -1) {
		switch (opt) {
		case 's': sr = atoi(optarg); break;
		case 'n': nb = atoi(optarg); break;
		case 'b': bs = atoi(optarg); break;
		case 't': ntransfers = strtoull(optarg, NULL, 10); break;   /* here: transfers the synthetic receiver sends */
		case 'o': fname = optarg; break;
		case 'p': use_float = 1; break;
		default: return 2;
		}
	}

	fakeusb_config dev;
	memset(&dev, 0, sizeof dev);
	dev.struct_size = sizeof dev;
	dev.seed = PERSEUS_SYNTH_SEED;
	dev.limit = ntransfers;
	fakeusb_plug(&dev);

	/* ---- libperseus-sdr, as in perseustest.c ---- */
	if (perseus_init() < 1) { fprintf(stderr, "No Perseus receivers detected\n"); return 1; }
	perseus_descr *descr = perseus_open(0);
	if (!descr) { fprintf(stderr, "error: %s\n", perseus_errorstr()); return 1; }
	if (perseus_firmware_download(descr, NULL) < 0) { fprintf(stderr, "firmware download error: %s\n", perseus_errorstr()); return 1; }
	if (perseus_set_sampling_rate(descr, sr) < 0) { fprintf(stderr, "fpga configuration error: %s\n", perseus_errorstr()); return 1; }

	/* ---- the B200 path: what replaces `fout = fopen(...)` + the CPU callback ---- */
	perseus_gpu *gpu = NULL;
	perseus_gpu_config cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.struct_size = sizeof cfg;
	cfg.stream_flags = use_float ? PERSEUS_GPU_OUT_FLOAT : PERSEUS_GPU_OUT_INT32;
	if (perseus_gpu_open(&gpu, &cfg) < 0) { fprintf(stderr, "perseus-gpu: %s\n", perseus_gpu_errorstr()); return 1; }
	if (perseus_gpu_stream_to_file(gpu, fname) < 0) { fprintf(stderr, "perseus-gpu: %s\n", perseus_gpu_errorstr()); return 1; }
	if (perseus_gpu_prepare(gpu) < 0) { fprintf(stderr, "perseus-gpu: %s\n", perseus_gpu_errorstr()); return 1; }   /* not on the poll thread */

	/* was: perseus_start_async_input(descr, nb*bs, user_data_callback_c_u, fout) */
	if (perseus_start_async_input(descr, (uint32_t)(nb * bs), perseus_gpu_input_callback, gpu) < 0) {
		fprintf(stderr, "start async input error: %s\n", perseus_errorstr());
		return 1;
	}

	fakeusb_state st;   /* perseustest sleeps for its test time here; the synthetic receiver says when it has sent everything */
	do { usleep(1000); fakeusb_get_state(&st); } while (st.stream_pos < ntransfers);

	perseus_stop_async_input(descr);             /* blocks until the queue is cancelled, perseus-sdr.c:714-716 */
	perseus_gpu_stats gs;
	int rc = perseus_gpu_flush(gpu);             /* what fclose(fout) guarantees in perseustest.c:387-390 */
	perseus_gpu_get_stats(gpu, &gs);
	if (rc == 0) rc = perseus_gpu_close(gpu);
	if (rc < 0) { fprintf(stderr, "perseus-gpu: %s\n", perseus_gpu_errorstr()); return 1; }
	perseus_close(descr);
	perseus_exit();
	printf("%llu transfers, %llu samples unpacked on the GPU in %llu slabs (%llu submitted by the latency watchdog)\n",
	       (unsigned long long)gs.callbacks, (unsigned long long)gs.samples, (unsigned long long)gs.slabs,
	       (unsigned long long)gs.watchdog_submits);
	return 0;
}
