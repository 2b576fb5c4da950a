// Helper threads that move bytes between an application's PAGEABLE memory (malloc, mmap: what a libperseus-sdr user holds) and
// the handle's pinned bounce buffers, so that perseus_gpu_unpack can feed the copy engines from ordinary memory at several
// times the rate of one memcpy thread (the CUDA runtime stages pageable copies on the calling thread alone).
// No CUDA in here; part of the host layer that tests/sanitize/ builds with TSAN/ASAN.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

namespace pg {

// dst is written once and next read by a DMA engine: non-temporal stores (no read-for-ownership, no cache pollution).
void copy_nontemporal(uint8_t *dst, const uint8_t *src, size_t n);

class CopyPool {
public:
	// `helpers` threads besides the caller; they sleep between jobs (after a short spin, so back-to-back chunks of one call do
	// not pay a wake-up each).  helpers == 0: copy() is a plain copy on the calling thread.
	explicit CopyPool(int helpers);
	~CopyPool();
	CopyPool(const CopyPool &) = delete;
	CopyPool &operator=(const CopyPool &) = delete;
	// Copies n bytes with the caller and the helpers working on disjoint slices; returns when all of it is done.
	// ONE caller at a time (the handle's lock sees to that).
	void copy(void *dst, const void *src, size_t n, bool nontemporal);
	int helpers() const { return (int)th_.size(); }

private:
	struct Job { uint8_t *dst; const uint8_t *src; size_t n; bool nt; int parts; };
	static void run_slice(const Job &j, int part);
	void worker(int index);
	std::vector<std::thread> th_;
	std::mutex mu_;
	std::condition_variable cv_;
	Job job_{};                          // written under mu_ before gen_ moves on
	std::atomic<uint64_t> gen_{0};       // job number; helpers spin on it briefly, then sleep on cv_
	std::atomic<int> done_{0};           // helpers that have finished the current job
	bool stop_ = false;                  // under mu_
};

}  // namespace pg
