// See copy_pool.h.
#include "copy_pool.h"

#include <chrono>
#include <cstring>
#include <sched.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace pg {

void copy_nontemporal(uint8_t *dst, const uint8_t *src, size_t n)
{
#if defined(__SSE2__)
	size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
	if (head > n) head = n;
	memcpy(dst, src, head);
	dst += head; src += head; n -= head;
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
#if defined(__SSE2__)
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
