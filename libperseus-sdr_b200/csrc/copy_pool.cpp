// See copy_pool.h.
#include "copy_pool.h"

#include <chrono>
#include <cstring>
#include <sched.h>
#include <cstdlib>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace pg {

// Three bodies of the same loop -- 16-, 32- and 64-byte non-temporal stores -- picked once from what the CPU has: one 64-byte
// store fills a write-combining buffer that takes four 16-byte ones (about a tenth faster per core where it exists).
#if defined(__x86_64__)
namespace {

using Body = size_t (*)(uint8_t *dst, const uint8_t *src, size_t n);   // copies whole blocks, returns the bytes it took

size_t body_sse2(uint8_t *dst, const uint8_t *src, size_t n)
{
	const size_t n0 = n;
	for (; n >= 64; n -= 64, src += 64, dst += 64) {
		const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src));
		const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 16));
		const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 32));
		const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 48));
		_mm_stream_si128(reinterpret_cast<__m128i *>(dst), a);
		_mm_stream_si128(reinterpret_cast<__m128i *>(dst + 16), b);
		_mm_stream_si128(reinterpret_cast<__m128i *>(dst + 32), c);
		_mm_stream_si128(reinterpret_cast<__m128i *>(dst + 48), d);
	}
	return n0 - n;
}

__attribute__((target("avx2"))) size_t body_avx2(uint8_t *dst, const uint8_t *src, size_t n)
{
	const size_t n0 = n;
	for (; n >= 128; n -= 128, src += 128, dst += 128) {
		const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src));
		const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 32));
		const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 64));
		const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 96));
		_mm256_stream_si256(reinterpret_cast<__m256i *>(dst), a);
		_mm256_stream_si256(reinterpret_cast<__m256i *>(dst + 32), b);
		_mm256_stream_si256(reinterpret_cast<__m256i *>(dst + 64), c);
		_mm256_stream_si256(reinterpret_cast<__m256i *>(dst + 96), d);
	}
	return n0 - n;
}

__attribute__((target("avx512f"))) size_t body_avx512(uint8_t *dst, const uint8_t *src, size_t n)
{
	const size_t n0 = n;
	for (; n >= 256; n -= 256, src += 256, dst += 256) {
		const __m512i a = _mm512_loadu_si512(src);
		const __m512i b = _mm512_loadu_si512(src + 64);
		const __m512i c = _mm512_loadu_si512(src + 128);
		const __m512i d = _mm512_loadu_si512(src + 192);
		_mm512_stream_si512(reinterpret_cast<__m512i *>(dst), a);
		_mm512_stream_si512(reinterpret_cast<__m512i *>(dst + 64), b);
		_mm512_stream_si512(reinterpret_cast<__m512i *>(dst + 128), c);
		_mm512_stream_si512(reinterpret_cast<__m512i *>(dst + 192), d);
	}
	return n0 - n;
}

Body pick_body()
{
	const char *force = getenv("PERSEUS_GPU_NT_COPY");                 // tests / A-B: sse2, avx2 or avx512
	__builtin_cpu_init();
	const bool has512 = __builtin_cpu_supports("avx512f"), has2 = __builtin_cpu_supports("avx2");
	if (force && !strcmp(force, "sse2")) return body_sse2;
	if (force && !strcmp(force, "avx2") && has2) return body_avx2;
	if (has512) return body_avx512;
	if (has2) return body_avx2;
	return body_sse2;
}

}  // namespace
#endif

void copy_nontemporal(uint8_t *dst, const uint8_t *src, size_t n)
{
#if defined(__x86_64__)
	static const Body body = pick_body();
	size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;   // the widest store wants a 64-byte aligned destination
	if (head > n) head = n;
	memcpy(dst, src, head);
	dst += head; src += head; n -= head;
	size_t done = body(dst, src, n);
	dst += done; src += done; n -= done;
	done = body_sse2(dst, src, n);                                       // what is left of a wider block, in 64-byte steps
	dst += done; src += done; n -= done;
#endif
	memcpy(dst, src, n);
}

CopyPool::CopyPool(int helpers)
{
	for (int k = 0; k < helpers; ++k) {
		try {
			th_.emplace_back(&CopyPool::worker, this, k + 1);
		} catch (...) {
			break;   // fewer helpers than asked for: copy() splits over the ones that exist
		}
	}
}

CopyPool::~CopyPool()
{
	{
		std::lock_guard<std::mutex> lk(mu_);
		stop_ = true;
	}
	cv_.notify_all();
	for (std::thread &t : th_) t.join();
}

void CopyPool::run_slice(const Job &j, int part)
{
	// slices start on 64-byte boundaries of the destination offset, so the non-temporal loop of each one stays aligned
	const size_t per = ((j.n + (size_t)j.parts - 1) / (size_t)j.parts + 63) & ~(size_t)63;
	const size_t a = per * (size_t)part;
	if (a >= j.n) return;
	const size_t len = j.n - a < per ? j.n - a : per;
	if (j.nt) {
		copy_nontemporal(j.dst + a, j.src + a, len);
#if defined(__x86_64__)
		_mm_sfence();   // this thread's non-temporal stores are visible before it reports the slice done
#endif
	} else {
		memcpy(j.dst + a, j.src + a, len);
	}
}

void CopyPool::worker(int index)
{
	uint64_t seen = 0;
	for (;;) {
		// a job usually follows the previous one within microseconds (chunk after chunk of one perseus_gpu_unpack): look for
		// it for a little while before going to sleep
		const auto spin_until = std::chrono::steady_clock::now() + std::chrono::microseconds(200);
		while (gen_.load(std::memory_order_acquire) == seen && std::chrono::steady_clock::now() < spin_until) sched_yield();
		Job j;
		{
			std::unique_lock<std::mutex> lk(mu_);
			cv_.wait(lk, [&] { return stop_ || gen_.load(std::memory_order_relaxed) != seen; });
			if (stop_) return;
			seen = gen_.load(std::memory_order_relaxed);
			j = job_;
		}
		if (index < j.parts) run_slice(j, index);
		done_.fetch_add(1, std::memory_order_release);
	}
}

void CopyPool::copy(void *dst, const void *src, size_t n, bool nontemporal)
{
	if (n == 0) return;
	const int helpers = (int)th_.size();
	// below 256 KiB per participant the hand-off costs more than it saves
	int parts = (int)(n / (256u << 10));
	if (parts > helpers + 1) parts = helpers + 1;
	Job j{static_cast<uint8_t *>(dst), static_cast<const uint8_t *>(src), n, nontemporal, parts < 1 ? 1 : parts};
	if (j.parts == 1) {
		run_slice(j, 0);
		return;
	}
	{
		std::lock_guard<std::mutex> lk(mu_);
		job_ = j;
		done_.store(0, std::memory_order_relaxed);
		gen_.fetch_add(1, std::memory_order_release);
	}
	cv_.notify_all();
	run_slice(j, 0);
	// every helper reports, also those without a slice: nobody is still looking at this job when the next one is posted
	while (done_.load(std::memory_order_acquire) != helpers) sched_yield();
}

}  // namespace pg
