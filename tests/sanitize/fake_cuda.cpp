// TEST INFRASTRUCTURE ONLY: implementation of tests/sanitize/fake_cuda/cuda_runtime.h and of the kernel launchers of
// libperseus-sdr_b200/csrc/kernels.h, so the host layer runs under TSAN/ASAN without a GPU.  The "kernels" call the CPU
// oracle (oracle/perseus_oracle.c) -- tests may; the product never links this file.
#include "kernels.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <thread>

extern "C" size_t perseus_oracle_unpack(int mode, const uint8_t *in, size_t nbytes, void *out);
extern "C" uint64_t perseus_oracle_checksum32(const uint32_t *words, size_t nwords, uint64_t first_index);
extern "C" void perseus_oracle_synth_random(uint8_t *dst, size_t nbytes, uint64_t seed, uint64_t byte_offset);

namespace {
std::mutex g_mu;
std::map<const void *, std::pair<size_t, cudaMemoryType>> g_allocs;   // base -> (size, kind)
struct Ev { double t = 0; uint64_t ticket = 0; };

// cudaLaunchHostFunc: like the real runtime, host functions run on a thread of their own, in order, concurrently with the
// caller; everything else "on the device" completes at once.  Events remember how many host functions were queued before
// them, so waiting for an event / a stream waits for those functions, as it does on a real stream.
struct HostFuncs {
	std::mutex mu;
	std::condition_variable cv;
	std::deque<std::pair<cudaHostFn_t, void *>> q;
	uint64_t queued = 0, done = 0;
	bool stop = false;
	std::thread th;
	void run()
	{
		std::unique_lock<std::mutex> lk(mu);
		for (;;) {
			cv.wait(lk, [&] { return stop || !q.empty(); });
			if (q.empty()) return;
			auto f = q.front();
			q.pop_front();
			lk.unlock();
			f.first(f.second);
			lk.lock();
			++done;
			cv.notify_all();
		}
	}
	void push(cudaHostFn_t fn, void *p)
	{
		std::lock_guard<std::mutex> lk(mu);
		if (!th.joinable()) th = std::thread([this] { run(); });
		q.emplace_back(fn, p);
		++queued;
		cv.notify_all();
	}
	uint64_t ticket() { std::lock_guard<std::mutex> lk(mu); return queued; }
	bool reached(uint64_t t) { std::lock_guard<std::mutex> lk(mu); return done >= t; }
	void wait(uint64_t t) { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return done >= t; }); }
	~HostFuncs()
	{
		{ std::lock_guard<std::mutex> lk(mu); stop = true; }
		cv.notify_all();
		if (th.joinable()) th.join();
	}
} g_hostfuncs;
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
std::atomic<long> g_fail_alloc_in{0};   // fake_cuda_fail_alloc_in(n): the n-th allocation from now fails once (0 = none)
cudaError_t alloc(void **p, size_t n, cudaMemoryType kind)
{
	if (n > (size_t(1) << 40)) return cudaErrorMemoryAllocation;
	if (g_fail_alloc_in.load() > 0 && g_fail_alloc_in.fetch_sub(1) == 1) return cudaErrorMemoryAllocation;
	*p = malloc(n ? n : 1);
	if (!*p) return cudaErrorMemoryAllocation;
	std::lock_guard<std::mutex> lk(g_mu);
	g_allocs[*p] = {n, kind};
	return cudaSuccess;
}
cudaError_t release(void *p)
{
	if (!p) return cudaSuccess;
	{
		std::lock_guard<std::mutex> lk(g_mu);
		g_allocs.erase(p);
	}
	free(p);
	return cudaSuccess;
}
}  // namespace

// test hook (host-simulation build only): make the n-th device / pinned allocation from now fail, once
extern "C" void fake_cuda_fail_alloc_in(long n) { g_fail_alloc_in.store(n); }

cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "fake CUDA error"; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
	memset(p, 0, sizeof(*p));
	snprintf(p->name, sizeof(p->name), "fake sm_100 (sanitizer build)");
	p->multiProcessorCount = 148; p->major = 10; p->minor = 0; p->totalGlobalMem = size_t(180) << 30;
	return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = reinterpret_cast<cudaStream_t>(new int(0)); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { g_hostfuncs.wait(g_hostfuncs.ticket()); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete reinterpret_cast<int *>(s); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaLaunchHostFunc(cudaStream_t, cudaHostFn_t fn, void *p) { g_hostfuncs.push(fn, p); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = reinterpret_cast<cudaEvent_t>(new Ev()); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t)
{
	Ev *ev = reinterpret_cast<Ev *>(e);
	ev->t = now_ms();
	ev->ticket = g_hostfuncs.ticket();
	return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) { g_hostfuncs.wait(reinterpret_cast<Ev *>(e)->ticket); return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t e) { return g_hostfuncs.reached(reinterpret_cast<Ev *>(e)->ticket) ? cudaSuccess : cudaErrorNotReady; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = float(reinterpret_cast<Ev *>(b)->t - reinterpret_cast<Ev *>(a)->t) + 1e-3f; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete reinterpret_cast<Ev *>(e); return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n) { return alloc(p, n, cudaMemoryTypeDevice); }
cudaError_t cudaFree(void *p) { return release(p); }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return alloc(p, n, cudaMemoryTypeHost); }
cudaError_t cudaFreeHost(void *p) { return release(p); }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p)
{
	std::lock_guard<std::mutex> lk(g_mu);
	a->type = cudaMemoryTypeUnregistered;
	auto it = g_allocs.upper_bound(p);
	if (it != g_allocs.begin()) {
		--it;
		const char *base = static_cast<const char *>(it->first);
		if (static_cast<const char *>(p) < base + it->second.first) a->type = it->second.second;
	}
	return cudaSuccess;
}

namespace pg {

int stream_preroll(const void *, const void *) { return 0; }

cudaError_t launch_unpack(const void *in, size_t nbytes, void *out_i32, void *out_f32, unsigned fmt, const Tuning &, int, cudaStream_t, int *launches)
{
	*launches = 0;
	if (nbytes / 6 == 0) return cudaSuccess;
	// PERSEUS_FAKE_LAUNCH_US: every launch takes this long on the calling thread -- a submission path that has become the
	// bottleneck (a profiler serialising launches, a GPU busy elsewhere), for the tests of how the streaming path batches then
	static const long slow_us = [] { const char *e = getenv("PERSEUS_FAKE_LAUNCH_US"); return e ? atol(e) : 0L; }();
	if (slow_us > 0) {
		const double until = now_ms() + slow_us * 1e-3;
		while (now_ms() < until) { }
	}
	if (fmt & FMT_I32) perseus_oracle_unpack(0, static_cast<const uint8_t *>(in), nbytes, out_i32);
	if (fmt & FMT_F32) perseus_oracle_unpack(1, static_cast<const uint8_t *>(in), nbytes, out_f32);
	if (fmt & FMT_POW2) perseus_oracle_unpack(2, static_cast<const uint8_t *>(in), nbytes, out_f32);
	*launches = 1;
	return cudaSuccess;
}

cudaError_t launch_unpack_batch(const SegDesc *segs, const TileRef *ts, uint64_t ns, const TileRef *td, uint64_t nd, int tile, unsigned fmt,
                                const Tuning &t, int sm, cudaStream_t st, int *launches)
{
	*launches = 0;
	for (int pass = 0; pass < 2; ++pass) {
		const TileRef *tiles = pass ? td : ts;
		const uint64_t n = pass ? nd : ns;
		for (uint64_t k = 0; k < n; ++k) {
			const SegDesc &sd = segs[tiles[k].seg];
			const uint64_t off = (uint64_t)tiles[k].tile * tile, used = sd.nbytes / 6 * 6;
			const uint64_t len = used - off < (uint64_t)tile ? used - off : (uint64_t)tile;
			int one = 0;
			launch_unpack(sd.in + off, len, sd.out_i32 ? (char *)sd.out_i32 + off / 6 * 8 : nullptr, sd.out_f32 ? (char *)sd.out_f32 + off / 6 * 8 : nullptr,
			              fmt, t, sm, st, &one);
		}
		if (n) ++*launches;
	}
	return cudaSuccess;
}

cudaError_t launch_generate(void *dst, size_t nbytes, int pattern, uint64_t seed, uint64_t off, int, cudaStream_t)
{
	host_generate(static_cast<uint8_t *>(dst), nbytes, pattern, seed, off);
	return cudaSuccess;
}

cudaError_t launch_checksum(const void *w, size_t n, uint64_t first, unsigned long long *sum, int, cudaStream_t, bool accumulate)
{
	const uint64_t c = perseus_oracle_checksum32(static_cast<const uint32_t *>(w), n, first);
	*sum = (accumulate ? *sum : 0ull) + c;
	return cudaSuccess;
}

cudaError_t launch_verify(const void *, size_t, const void *, const void *, unsigned, unsigned long long *r, int, cudaStream_t)
{
	r[0] = 0; r[1] = ~0ull;
	return cudaSuccess;
}

cudaError_t launch_probe(int, const void *, void *, size_t, int, int, cudaStream_t) { return cudaSuccess; }

}  // namespace pg
