// Pieces shared by every translation unit of libperseus_gpu.so that need no CUDA: the calling
// thread's last-error string (cf. perseus_errorstr(), /root/reference/perseuserr.c:36-42) and
// the definition of the synthetic wire stream.  perseus_vrx.cpp and perseus_host.cpp include
// only this header, so the host layer can be built (and run under TSAN/ASAN) without nvcc.
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define PG_HD __host__ __device__ __forceinline__
#else
#define PG_HD inline
#endif

namespace pg {

// Formats the calling thread's error message and returns `code` (always negative).
int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
int vfail(int code, const char *fmt, va_list ap);
// Read/replace the calling thread's message (used to carry a message across a clean-up that may fail itself).
const char *last_error();
void set_last_error(const char *msg);

// ---- synthetic stream --------------------------------------------------------------------
// RANDOM: 64-bit little-endian word w of the byte stream = splitmix64(seed + w).
PG_HD uint64_t splitmix64(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}

// RAMP: sample v has I = v mod 2^24, Q = (uint32)(v*2654435761) >> 8, each 3 bytes little endian.
PG_HD void ramp_sample(uint64_t v64, uint8_t *p)
{
	const uint32_t v = (uint32_t)v64;
	const uint32_t i24 = v & 0xFFFFFFu;
	const uint32_t q24 = (uint32_t)(v * 2654435761u) >> 8;
	p[0] = (uint8_t)i24; p[1] = (uint8_t)(i24 >> 8); p[2] = (uint8_t)(i24 >> 16);
	p[3] = (uint8_t)q24; p[4] = (uint8_t)(q24 >> 8); p[5] = (uint8_t)(q24 >> 16);
}

// Host mirror of the device generator (bit-identical), for perseus_synth_fill and the virtual receiver.
void host_generate(uint8_t *dst, size_t nbytes, int pattern, uint64_t seed, uint64_t byte_offset);

}  // namespace pg
