// CUDA-free part of the C ABI (include/perseus-gpu.h): last-error string, synthetic wire data on
// the host, the shard planner.  Nothing here computes samples.
//
// Reference anchors: the error convention (negative code + message getter) mirrors
// /root/reference/perseus-sdr.h:317-366 and perseuserr.c:36-42, except that the message is
// per thread (the reference's globals are not thread-safe, perseus-sdr.h:362-364).
#include "../../include/perseus-gpu.h"
#include "host_common.h"

#include <cstdio>
#include <cstring>

namespace pg {

namespace {
thread_local char g_errstr[1024] = "";
}

int vfail(int code, const char *fmt, va_list ap)
{
	vsnprintf(g_errstr, sizeof(g_errstr), fmt, ap);
	return code;
}

int fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_errstr, sizeof(g_errstr), fmt, ap);
	va_end(ap);
	return code;
}

const char *last_error() { return g_errstr; }

void set_last_error(const char *msg) { snprintf(g_errstr, sizeof(g_errstr), "%s", msg ? msg : ""); }

void host_generate(uint8_t *dst, size_t nbytes, int pattern, uint64_t seed, uint64_t byte_offset)
{
	if (pattern == PERSEUS_SYNTH_RAMP) {
		const size_t ns = nbytes / 6;
		for (size_t k = 0; k < ns; ++k) ramp_sample(byte_offset / 6 + k, dst + 6 * k);
		memset(dst + ns * 6, 0, nbytes - ns * 6);
		return;
	}
	size_t i = 0;
	// head: up to the next 8-byte stream-word boundary
	while (i < nbytes && ((byte_offset + i) & 7)) {
		const uint64_t pos = byte_offset + i;
		dst[i++] = (uint8_t)(splitmix64(seed + (pos >> 3)) >> (8 * (pos & 7)));
	}
	for (; i + 8 <= nbytes; i += 8) {
		const uint64_t word = splitmix64(seed + ((byte_offset + i) >> 3));   // little-endian host
		memcpy(dst + i, &word, 8);
	}
	for (; i < nbytes; ++i) {
		const uint64_t pos = byte_offset + i;
		dst[i] = (uint8_t)(splitmix64(seed + (pos >> 3)) >> (8 * (pos & 7)));
	}
}

}  // namespace pg

extern "C" {

const char *perseus_gpu_errorstr(void) { return pg::last_error(); }

int perseus_synth_fill(void *host_dst, size_t nbytes, int pattern, uint64_t seed, uint64_t byte_offset)
{
	if (pattern != PERSEUS_SYNTH_RANDOM && pattern != PERSEUS_SYNTH_RAMP) return pg::fail(PERSEUS_GPU_ERRPARAM, "unknown pattern %d", pattern);
	if (pattern == PERSEUS_SYNTH_RAMP && byte_offset % 6) return pg::fail(PERSEUS_GPU_ERRPARAM, "RAMP byte_offset must be a multiple of 6");
	if (nbytes && !host_dst) return pg::fail(PERSEUS_GPU_ERRPARAM, "null destination");
	pg::host_generate(static_cast<uint8_t *>(host_dst), nbytes, pattern, seed, byte_offset);
	return 0;
}

int perseus_gpu_shard_range(uint64_t total, int nshards, int shard, uint64_t *first, uint64_t *count)
{
	if (nshards < 1 || shard < 0 || shard >= nshards || !first || !count) return pg::fail(PERSEUS_GPU_ERRPARAM, "bad shard arguments");
	// floor(shard*total/nshards) without overflowing 64 bits
	const unsigned __int128 t = total;
	const uint64_t a = (uint64_t)(t * (unsigned)shard / (unsigned)nshards);
	const uint64_t b = (uint64_t)(t * (unsigned)(shard + 1) / (unsigned)nshards);
	*first = a;
	*count = b - a;
	return 0;
}

int perseus_gpu_shard_range_weighted(uint64_t total, int nshards, const double *weights, int shard, uint64_t *first, uint64_t *count)
{
	if (nshards < 1 || shard < 0 || shard >= nshards || !weights || !first || !count) return pg::fail(PERSEUS_GPU_ERRPARAM, "bad shard arguments");
	long double sum = 0.0L, before = 0.0L;
	bool equal = true;
	for (int k = 0; k < nshards; ++k) {
		if (!(weights[k] >= 0.0) || weights[k] > 1e300) return pg::fail(PERSEUS_GPU_ERRPARAM, "weight %d is negative or not finite", k);
		if (weights[k] != weights[0]) equal = false;
		if (k < shard) before += (long double)weights[k];
		sum += (long double)weights[k];
	}
	if (!(sum > 0.0L)) return pg::fail(PERSEUS_GPU_ERRPARAM, "all weights are zero");
	if (equal) return perseus_gpu_shard_range(total, nshards, shard, first, count);   // exact integer arithmetic
	auto cut = [&](long double w) -> uint64_t {
		long double x = (long double)total * (w / sum);
		uint64_t b = x <= 0.0L ? 0 : (uint64_t)x;
		return b > total ? total : b;
	};
	const uint64_t a = shard == 0 ? 0 : cut(before);
	const uint64_t b = shard == nshards - 1 ? total : cut(before + (long double)weights[shard]);
	*first = a;
	*count = b > a ? b - a : 0;
	return 0;
}

}  // extern "C"
