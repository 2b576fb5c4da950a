/* TEST INFRASTRUCTURE ONLY — the CPU oracle.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this.  The product
 * (libperseus-sdr_b200/, include/) never links, loads or calls it.
 *
 * A memory->memory restatement, in our own words, of the one piece of sample
 * arithmetic in Microtelecom/libperseus-sdr: the two example callbacks
 *     user_data_callback_c_u   /root/reference/examples/perseustest.c:432-460
 *                              (duplicate: examples/simple.c:33-61)
 *     user_data_callback_c_f   /root/reference/examples/perseustest.c:466-502
 * which the library's delivery path (perseus-in.c:206-207) calls once per
 * completed USB transfer.
 *
 * Wire format (perseustest.c:434, :450-455): 6 bytes per complex sample,
 * I0 I1 I2 Q0 Q1 Q2, each field a little-endian 24-bit two's-complement value.
 *
 * PARITY PINNING: the reference ships no golden vectors or known-answer tests for
 * this path (its `make check` only builds examples/simple.c).  The restatement is
 * pinned instead against the reference's own callbacks compiled verbatim
 * (oracle/ref_harness.c -> oracle/_ref/libperseus_ref.so) over ALL 2^24 codes of a
 * 24-bit field, and against tests/golden/ vectors produced by that library
 * (tests/golden/make_golden.py).  See tests/test_oracle.py.
 */
#include <stdint.h>
#include <stddef.h>
#include <limits.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

/* perseustest.c:411-426 overlays the three wire bytes on bytes 1..3 of a
 * little-endian int32 whose byte 0 is cleared (:447 / :486), i.e. the 24-bit
 * value is MSB-aligned: the sign is the top bit of the third wire byte. */
static inline int32_t field24_msb_aligned(const uint8_t *p)
{
	uint32_t u = ((uint32_t)p[0] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 24);
	return (int32_t)u;
}

/* perseustest.c:443 — nSamples = buf_size/6, trailing buf_size%6 bytes are ignored. */
size_t perseus_oracle_nsamples(size_t nbytes) { return nbytes / 6; }

/* perseustest.c:449-457: int32 I then int32 Q per sample, 8 output bytes. */
size_t perseus_oracle_unpack_i32(const uint8_t *in, size_t nbytes, int32_t *out)
{
	size_t ns = nbytes / 6;
	for (size_t k = 0; k < ns; k++) {
		out[2 * k]     = field24_msb_aligned(in + 6 * k);
		out[2 * k + 1] = field24_msb_aligned(in + 6 * k + 3);
	}
	return ns;
}

/* perseustest.c:496-497: (float)int32 / (INT_MAX - 256).  The divisor is an int
 * expression that the compiler converts to float (2147483392.0f, exactly
 * representable); it is spelled the same way here so that conversion is the
 * compiler's, as in the reference.  NOT 2^-31. */
size_t perseus_oracle_unpack_f32(const uint8_t *in, size_t nbytes, float *out)
{
	size_t ns = nbytes / 6;
	for (size_t k = 0; k < ns; k++) {
		int32_t i = field24_msb_aligned(in + 6 * k);
		int32_t q = field24_msb_aligned(in + 6 * k + 3);
		out[2 * k]     = (float)i / (INT_MAX - 256);
		out[2 * k + 1] = (float)q / (INT_MAX - 256);
	}
	return ns;
}

/* The scale BASELINE.json's north_star names (x * 2^-31).  It is NOT what the
 * reference computes (differs for every non-zero input); provided so the
 * explicitly-named POW2 mode of the product has an oracle too. */
size_t perseus_oracle_unpack_f32_pow2(const uint8_t *in, size_t nbytes, float *out)
{
	size_t ns = nbytes / 6;
	for (size_t k = 0; k < ns; k++) {
		int32_t i = field24_msb_aligned(in + 6 * k);
		int32_t q = field24_msb_aligned(in + 6 * k + 3);
		out[2 * k]     = (float)i * 0x1p-31f;
		out[2 * k + 1] = (float)q * 0x1p-31f;
	}
	return ns;
}

/* mode: 0 = int32, 1 = float (reference scale), 2 = float (2^-31). */
size_t perseus_oracle_unpack(int mode, const uint8_t *in, size_t nbytes, void *out)
{
	switch (mode) {
	case 0: return perseus_oracle_unpack_i32(in, nbytes, (int32_t *)out);
	case 1: return perseus_oracle_unpack_f32(in, nbytes, (float *)out);
	case 2: return perseus_oracle_unpack_f32_pow2(in, nbytes, (float *)out);
	default: return (size_t)-1;
	}
}

/* All-cores CPU baseline: static split by whole samples; each thread runs the
 * scalar loop above on its range (samples are independent, perseustest.c:449-458). */
typedef struct { int mode; const uint8_t *in; size_t nbytes; uint8_t *out; } oracle_job;

static void *oracle_job_main(void *arg)
{
	oracle_job *j = (oracle_job *)arg;
	perseus_oracle_unpack(j->mode, j->in, j->nbytes, j->out);
	return NULL;
}

size_t perseus_oracle_unpack_mt(int mode, const uint8_t *in, size_t nbytes, void *out, int nthreads)
{
	size_t ns = nbytes / 6;
	if (mode < 0 || mode > 2) return (size_t)-1;
	if (nthreads <= 1 || ns < (size_t)nthreads * 1024) return perseus_oracle_unpack(mode, in, nbytes, out);
	if (nthreads > 1024) nthreads = 1024;
	oracle_job jobs[1024];
	pthread_t th[1024];
	int started[1024];
	for (int t = 0; t < nthreads; t++) {
		size_t s0 = ns * (size_t)t / (size_t)nthreads, s1 = ns * (size_t)(t + 1) / (size_t)nthreads;
		jobs[t].mode = mode;
		jobs[t].in = in + 6 * s0;
		jobs[t].nbytes = 6 * (s1 - s0);
		jobs[t].out = (uint8_t *)out + 8 * s0;
		started[t] = pthread_create(&th[t], NULL, oracle_job_main, &jobs[t]) == 0;
		if (!started[t]) oracle_job_main(&jobs[t]);
	}
	for (int t = 0; t < nthreads; t++) if (started[t]) pthread_join(th[t], NULL);
	return ns;
}

/* ---- hashes used by the golden fixtures and the full-size GPU parity tests ---- */

/* FNV-1a 64 over a byte stream (offset basis 0xcbf29ce484222325, prime 0x100000001b3);
 * pass the previous return value as `h` to continue a stream, 0 to start one. */
uint64_t perseus_oracle_fnv1a64(const void *data, size_t n, uint64_t h)
{
	const uint8_t *p = (const uint8_t *)data;
	if (h == 0) h = 0xcbf29ce484222325ull;
	for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
	return h;
}

/* Position-weighted 64-bit checksum of 32-bit output words: order-independent sum
 * (mod 2^64) of mix(index) * (word + 1), so shards/tiles can be summed in any order
 * and a checksum of a whole recording equals the sum of its shards' checksums.
 * The product's device-side checksum kernel (csrc/unpack_kernels.cu) follows the
 * same definition; this is the independent CPU statement of it. */
static inline uint64_t splitmix64(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}

uint64_t perseus_oracle_checksum32(const uint32_t *w, size_t nwords, uint64_t first_index)
{
	uint64_t acc = 0;
	for (size_t i = 0; i < nwords; i++)
		acc += (splitmix64(first_index + i) | 1ull) * ((uint64_t)w[i] + 1ull);
	return acc;
}

/* ---- synthetic wire data (definition shared with the product's generator) ------
 * byte stream whose 64-bit little-endian word number w is
 * splitmix64(seed + w); every 24-bit field is therefore an independent uniform draw.
 * Random access: any [byte_offset, byte_offset+nbytes) range can be regenerated. */
void perseus_oracle_synth_random(uint8_t *dst, size_t nbytes, uint64_t seed, uint64_t byte_offset)
{
	for (size_t i = 0; i < nbytes; i++) {
		uint64_t pos = byte_offset + i;
		uint64_t word = splitmix64(seed + (pos >> 3));
		dst[i] = (uint8_t)(word >> (8 * (pos & 7)));
	}
}

/* Exhaustive ramp: sample number v has I = v mod 2^24 and
 * Q = ((uint32)(v * 2654435761u)) >> 8, both as 3-byte little-endian fields. */
void perseus_oracle_synth_ramp(uint8_t *dst, size_t nsamples, uint64_t first_sample)
{
	for (size_t k = 0; k < nsamples; k++) {
		uint32_t v = (uint32_t)(first_sample + k);
		uint32_t i24 = v & 0xFFFFFFu;
		uint32_t q24 = ((uint32_t)(v * 2654435761u)) >> 8;
		uint8_t *p = dst + 6 * k;
		p[0] = (uint8_t)i24; p[1] = (uint8_t)(i24 >> 8); p[2] = (uint8_t)(i24 >> 16);
		p[3] = (uint8_t)q24; p[4] = (uint8_t)(q24 >> 8); p[5] = (uint8_t)(q24 >> 16);
	}
}
