/* TEST INFRASTRUCTURE ONLY — never linked into, loaded by, or shipped with the product.
 *
 * oracle/_ref/libperseus_ref.so: the reference's OWN unpack callbacks, compiled
 * verbatim from /root/reference/examples/perseustest.c (nothing is copied into
 * this repository; the file is #included from where it lies).  The two callbacks
 *     user_data_callback_c_u   examples/perseustest.c:432-460  (24-bit -> MSB-aligned int32)
 *     user_data_callback_c_f   examples/perseustest.c:466-502  (24-bit -> float, /(INT_MAX-256))
 * are `static` there (forward declarations at :90-91), so the only way to reach
 * them unmodified is to include the translation unit.  `main` is renamed and the
 * library/fifo symbols it references get inert definitions below so the object
 * links with -z defs; none of them is ever called.
 *
 * The wrappers hand the callbacks a FILE* (their `extra` argument, perseustest.c:442)
 * backed by caller memory, so "what the reference would have written to its output
 * file" lands in a buffer the tests can compare byte for byte.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <pthread.h>

#define main perseustest_reference_main
#include "examples/perseustest.c"
#undef main

/* ---- inert definitions for the symbols perseustest.c's main() references ---- */
int  perseus_dbg_level = 0;
int  perseus_error = 0;
char perseus_error_str[1024];
char *perseus_errorstr(void) { return perseus_error_str; }
void perseus_set_debug(int level) { (void)level; }
int  perseus_init(void) { return 0; }
int  perseus_exit(void) { return 0; }
perseus_descr *perseus_open(int nDev) { (void)nDev; return NULL; }
int  perseus_close(perseus_descr *d) { (void)d; return 0; }
int  perseus_firmware_download(perseus_descr *d, char *f) { (void)d; (void)f; return 0; }
int  perseus_get_product_id(perseus_descr *d, eeprom_prodid *p) { (void)d; (void)p; return 0; }
int  perseus_set_attenuator_in_db(perseus_descr *d, int v) { (void)d; (void)v; return 0; }
int  perseus_get_attenuator_values(perseus_descr *d, int *b, unsigned int s) { (void)d; (void)b; (void)s; return 0; }
int  perseus_set_attenuator_n(perseus_descr *d, int v) { (void)d; (void)v; return 0; }
int  perseus_set_attenuator(perseus_descr *d, uint8_t v) { (void)d; (void)v; return 0; }
int  perseus_set_adc(perseus_descr *d, int a, int b) { (void)d; (void)a; (void)b; return 0; }
int  perseus_set_ddc_center_freq(perseus_descr *d, double f, int e) { (void)d; (void)f; (void)e; return 0; }
int  perseus_start_async_input(perseus_descr *d, uint32_t n, perseus_input_callback cb, void *x)
     { (void)d; (void)n; (void)cb; (void)x; return 0; }
int  perseus_stop_async_input(perseus_descr *d) { (void)d; return 0; }
int  perseus_set_sampling_rate(perseus_descr *d, int r) { (void)d; (void)r; return 0; }
int  perseus_set_sampling_rate_n(perseus_descr *d, unsigned int r) { (void)d; (void)r; return 0; }
int  perseus_get_sampling_rates(perseus_descr *d, int *b, unsigned int s) { (void)d; (void)b; (void)s; return 0; }
int  perseus_is_preserie(perseus_descr *d, int *f) { (void)d; (void)f; return 0; }
int  make_fifo(const char *n, perseus_descr *pd) { (void)n; (void)pd; return 0; }
int  run_fifo(void) { return 0; }
void stop_fifo(void) { }

/* ---- exported wrappers ---------------------------------------------------- */

/* Runs the reference callback once per `chunk` bytes of `in` (the way
 * input_queue_callback, perseus-in.c:206-207, hands it one transfer at a time) and
 * captures everything it fwrite()s into out[0..out_cap).  Returns bytes written,
 * or (size_t)-1 if the sink overflowed / could not be opened. */
typedef struct { unsigned char *dst; size_t cap, pos; int overflow; } mem_sink;

/* fopencookie sink: plain memcpy into caller memory (fmemopen is unusable here: glibc
 * reserves the last byte of the buffer for a terminating NUL). */
static ssize_t mem_sink_write(void *cookie, const char *data, size_t n)
{
	mem_sink *s = (mem_sink *)cookie;
	if (n > s->cap - s->pos) { s->overflow = 1; return 0; }
	memcpy(s->dst + s->pos, data, n);
	s->pos += n;
	return (ssize_t)n;
}

static size_t run_callback(perseus_input_callback cb, const void *in, size_t nbytes,
                           size_t chunk, void *out, size_t out_cap)
{
	if (chunk == 0 || chunk > (size_t)INT_MAX) return (size_t)-1;
	mem_sink sink = { (unsigned char *)out, out_cap, 0, 0 };
	cookie_io_functions_t io = { NULL, mem_sink_write, NULL, NULL };
	FILE *f = fopencookie(&sink, "w", io);
	if (!f) return (size_t)-1;
	setvbuf(f, NULL, _IOFBF, 1u << 20);
	const unsigned char *p = (const unsigned char *)in;
	size_t left = nbytes;
	while (left) {
		size_t n = left < chunk ? left : chunk;
		cb((void *)p, (int)n, f);
		p += n; left -= n;
	}
	int bad = fflush(f) != 0;
	fclose(f);
	return (bad || sink.overflow) ? (size_t)-1 : sink.pos;
}

size_t perseus_ref_unpack_i32(const void *in, size_t nbytes, size_t chunk, void *out, size_t out_cap)
{ return run_callback(user_data_callback_c_u, in, nbytes, chunk, out, out_cap); }

size_t perseus_ref_unpack_f32(const void *in, size_t nbytes, size_t chunk, void *out, size_t out_cap)
{ return run_callback(user_data_callback_c_f, in, nbytes, chunk, out, out_cap); }

/* Timing entry for bench.py --impl reference: `nthreads` independent streams, each
 * thread pushing its own slice of `in` through the reference callback in `chunk`-byte
 * transfers into its own slice of `out` (8 output bytes per 6 input bytes).
 * The callback is inherently serial per stream (one fwrite per sample), so threads
 * model independent receivers.  Returns total bytes written, (size_t)-1 on error. */
typedef struct { perseus_input_callback cb; const unsigned char *in; size_t nbytes, chunk;
                 unsigned char *out; size_t out_cap, written; } ref_job;

static void *ref_job_main(void *arg)
{
	ref_job *j = (ref_job *)arg;
	j->written = run_callback(j->cb, j->in, j->nbytes, j->chunk, j->out, j->out_cap);
	return NULL;
}

size_t perseus_ref_unpack_mt(int want_float, const void *in, size_t nbytes, size_t chunk,
                             void *out, size_t out_cap, int nthreads)
{
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 1024) nthreads = 1024;
	size_t nchunks = nbytes / chunk;            /* whole transfers only */
	ref_job *jobs = (ref_job *)calloc((size_t)nthreads, sizeof(ref_job));
	pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
	if (!jobs || !th) { free(jobs); free(th); return (size_t)-1; }
	size_t total = 0; int ok = 1;
	for (int t = 0; t < nthreads; t++) {
		size_t c0 = nchunks * (size_t)t / (size_t)nthreads, c1 = nchunks * (size_t)(t + 1) / (size_t)nthreads;
		jobs[t].cb = want_float ? user_data_callback_c_f : user_data_callback_c_u;
		jobs[t].in = (const unsigned char *)in + c0 * chunk;
		jobs[t].nbytes = (c1 - c0) * chunk;
		jobs[t].chunk = chunk;
		size_t o0 = c0 * (chunk / 6) * 8, o1 = c1 * (chunk / 6) * 8;
		if (o1 > out_cap) { ok = 0; o1 = o0; jobs[t].nbytes = 0; }
		jobs[t].out = (unsigned char *)out + o0;
		jobs[t].out_cap = o1 - o0;
	}
	for (int t = 0; t < nthreads; t++)
		if (jobs[t].nbytes == 0) { jobs[t].written = 0; th[t] = 0; }
		else if (pthread_create(&th[t], NULL, ref_job_main, &jobs[t]) != 0) { ok = 0; th[t] = 0; }
	for (int t = 0; t < nthreads; t++) {
		if (th[t]) pthread_join(th[t], NULL);
		if (jobs[t].written == (size_t)-1) ok = 0; else total += jobs[t].written;
	}
	free(jobs); free(th);
	return ok ? total : (size_t)-1;
}

const char *perseus_ref_build_info(void)
{
	return "verbatim examples/perseustest.c callbacks; " __VERSION__
#ifdef __FAST_MATH__
	       "; fast-math"
#endif
#ifdef __OPTIMIZE__
	       "; optimized"
#endif
	;
}
