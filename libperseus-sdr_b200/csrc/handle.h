// Internals of the C-ABI layer of libperseus_gpu.so shared by its three translation units:
//   handle.cu       life cycle, ownership of a handle (lock + callback hand-off), the callback's clock, tuning, device plumbing
//   stream_path.cu  the perseus_input_callback trampoline: slab ring, both slab routes, eager submission, age bound + watchdog,
//                   delivery thread, device / host / file sinks, flush
//   bulk_path.cu    perseus_gpu_unpack with its staging pipeline and copy pool, checksums, batched plans, generator / verify,
//                   probes, autotune
// No kernels here or there: all sample arithmetic lives in unpack_kernels.cu, so tests/sanitize and tests/hostsim compile these
// files as plain C++ against a CUDA stand-in.  Not installed; the public surface is include/perseus-gpu.h.
//
// Reference anchors: the callback contract is perseus-sdr.h:81 / perseus-in.c:204-207,263 (buffer valid only during the call,
// return value ignored, strictly serial, in ring order); the error convention mirrors perseus-sdr.h:317-366 + perseuserr.c:36-42.
#pragma once
#include "../../include/perseus-gpu.h"
#include "copy_pool.h"
#include "kernels.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include <pthread.h>
#include <sched.h>
#include <time.h>
#include <unistd.h>
#if defined(__linux__)
#include <linux/membarrier.h>
#include <sys/syscall.h>
#endif
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__x86_64__)
#include <cpuid.h>
#include <x86intrin.h>
#endif

namespace pgh {   // perseus-gpu host layer

using pg::fail;
inline int ok(int v = 0) { return v; }

constexpr int kMaxStreams = 8;
constexpr int kMaxSlabs = 64;
constexpr int kEventSlots = 32;
constexpr int kMaxStageSlots = 8;
constexpr size_t kDefaultDirectBytes = 256u << 10;   // perseus_gpu_config.direct_bytes = 0

struct Slab {
	uint8_t *host = nullptr;       // pinned wire bytes being filled by the callback
	uint8_t *dev_in = nullptr;     // device copy
	uint8_t *dev_i32 = nullptr;    // device outputs of this slab
	uint8_t *dev_f32 = nullptr;
	uint8_t *host_out[2] = {nullptr, nullptr};   // pinned copies of the outputs (int32, float): only with a file / host sink
	cudaEvent_t unpacked = nullptr;   // recorded behind the slab's kernel (and whatever the device sink queued): host delivery waits for it
	cudaEvent_t done = nullptr;    // recorded after the slab's last device operation (sleeping waits: back-pressure)
	cudaEvent_t ready = nullptr;   // the same point, for the delivery thread (spinning wait: lowest latency)
	uint64_t dlv_seq = 0;          // host delivery: this slab's place in the delivery order
	bool to_deliver = false;       // host delivery was queued for this use of the slab
	uint64_t first_sample = 0;
	uint64_t nsamples = 0;
	size_t file_bytes = 0;         // bytes deliver_slab writes to the file sink
	int file_fmt = 0;              // which host_out[] the file holds (perseustest writes ONE format per run)
	bool busy = false;
};

}  // namespace pgh

struct perseus_gpu {
	// Ownership of the handle's state.  Every C-ABI entry point takes `mu` (recursive: a sink runs under it and may call the
	// plumbing of its own handle) -- EXCEPT perseus_gpu_input_callback, whose fast path must not execute a single locked
	// instruction: the slab copy uses non-temporal stores, and a locked instruction (mutex, atomic read-modify-write, mfence)
	// stalls until the write-combining buffers have drained, ~0.25 us per 6144-byte transfer, a quarter of the path's
	// throughput.  The callback thread and everybody else therefore meet in an ASYMMETRIC Dekker handshake:
	//   callback:  cb_active = 1;  <compiler barrier>;  if (others_want) -> slow path (take mu like everybody else)
	//   others:    lock mu;  others_want++;  membarrier(PRIVATE_EXPEDITED);  wait until cb_active == 0
	// sys_membarrier makes every running thread of the process execute a full barrier at that instant, which supplies the
	// store-load ordering the callback side leaves out (and drains that core's write-combining buffers).  Where the system call
	// is not available both sides fall back to a real fence.
	std::recursive_mutex mu;
	int lock_depth = 0;                          // under mu: nested entries of the owning thread
	std::atomic<uint32_t> cb_active{0};          // 1 while the callback thread is inside its lock-free fast path
	std::atomic<uint32_t> others_want{0};        // threads that hold or wait for `mu`
	std::atomic<unsigned long> cb_thread{0};     // the thread that set cb_active (a sink called from it already owns the handle)
	std::atomic<uint64_t> partial_since_ns{0};   // published by the owner for the watchdog: age stamp of the partial slab, 0 = none
	bool asym = false;                           // membarrier available: the callback side needs no fence
	int device = 0;
	int sm_count = 0;
	perseus_gpu_config cfg{};
	pg::Tuning tune{};
	int nstreams = 0;
	cudaStream_t streams[pgh::kMaxStreams]{};        // [0] = the launch stream of everything device-resident; slabs rotate over all
	cudaStream_t s_in = nullptr, s_out = nullptr;   // copy-in / copy-out streams of the host-pointer pipeline
	cudaStream_t s_dlv = nullptr;               // host delivery of the streaming path: every slab's copy-out + deliver_slab, in stream order
	cudaEvent_t events[pgh::kEventSlots]{};          // the caller's timing slots (perseus_gpu_event_*)
	cudaEvent_t tev[4]{};                       // private timing events (autotune, probes)
	unsigned long long *d_scratch = nullptr;   // 2 x u64: checksum / verify results
	unsigned long long *h_scratch = nullptr;   // pinned mirror
	unsigned long long *d_sums = nullptr;      // 2 x u64: running checksums of PERSEUS_GPU_CHECKSUM calls (int32, float)
	// host-pointer pipeline of perseus_gpu_unpack: chunk c lives in slot c % nslots; per slot three events order
	// copy-in -> kernel -> copy-out and protect the slot's reuse
	size_t chunk_bytes = 0;
	int nslots = 0;
	uint64_t stage_seq = 0;                     // chunks staged so far (slot rotation continues across calls)
	uint8_t *stage_in[pgh::kMaxStageSlots]{};
	uint8_t *stage_out[pgh::kMaxStageSlots][2]{};
	cudaEvent_t ev_in[pgh::kMaxStageSlots]{}, ev_k[pgh::kMaxStageSlots]{}, ev_out[pgh::kMaxStageSlots]{};
	// PAGEABLE host buffers are staged through pinned bounce buffers by the caller and the pool's helper threads
	// (copy_threads participants; 0 = leave pageable memory to the CUDA runtime, which stages it on the calling thread alone)
	int copy_threads = 0;
	pg::CopyPool *pool = nullptr;
	uint8_t *bounce_in[pgh::kMaxStageSlots]{};
	uint8_t *bounce_out[pgh::kMaxStageSlots][2]{};
	// streaming (callback) path
	unsigned stream_fmt = 0;
	size_t slab_bytes = 0;
	int nslabs = 0;
	pgh::Slab slabs[pgh::kMaxSlabs];
	int cur = 0;               // slab being filled
	size_t fill = 0;           // bytes in it
	uint64_t fill_started_ns = 0;   // monotonic time the first transfer of the current slab arrived
	uint64_t max_latency_ns = 0;    // 0 = submit only full slabs
	uint64_t eager_gap_ns = 0;      // a transfer arriving this long after the previous one is submitted at once; 0 = never
	uint64_t last_push_ns = 0;      // monotonic time the previous callback began (0 = none yet)
	// the callback's clock: CLOCK_MONOTONIC carried forward by the time-stamp counter between anchors (see callback_now_ns)
	bool use_tsc = false;
	uint64_t tsc_anchor = 0, ns_anchor = 0;      // one reading of both clocks
	uint64_t tsc_reanchor = 0;                   // ticks after which the anchor is renewed (about a second)
	uint64_t ns_per_tick_q32 = 0;                // nanoseconds per tick, 32.32 fixed point
	int next_to_write = 0;     // oldest slab whose output has not reached the file sink
	bool streaming_ready = false;
	uint64_t samples_submitted = 0;
	perseus_gpu_sink sink = nullptr;
	void *sink_extra = nullptr;
	// host delivery (file sink, host sink): written by the owner only while nothing is in flight (after a flush), read by
	// deliver_slab on the handle's delivery thread
	perseus_gpu_host_sink host_sink = nullptr;
	void *host_sink_extra = nullptr;
	FILE *fout = nullptr;
	bool fout_is_stdout = false;                 // path "-" (perseustest.c:98,337): flushed, never closed
	std::atomic<int> io_error{0};                // deliver_slab could not write the file: surfaced at the next retire / flush
	std::atomic<uint64_t> host_blocks{0};        // blocks deliver_slab has handed over
	// The delivery thread: takes slabs in submission order, waits (spinning) for each one's outputs to have reached pinned host
	// memory, writes the file / calls the host sink.  A thread of the handle's own rather than cudaLaunchHostFunc because the
	// runtime dispatches host functions 0.15 ms late (measured, profiles/r2_latency_probe.jsonl); this one is usually there
	// before the data is.  dlv_mu guards the three counters below and pairs with the two condition variables.
	std::thread dlv_thread;
	std::mutex dlv_mu;
	std::condition_variable dlv_cv;              // delivery thread: something was submitted / stop
	std::condition_variable dlv_done_cv;         // owner: something was delivered
	uint64_t dlv_submitted = 0, dlv_delivered = 0;
	std::atomic<uint64_t> dlv_submitted_hint{0}; // = dlv_submitted, for the thread's short spin before it sleeps
	int dlv_ring[pgh::kMaxSlabs]{};                   // slab index of delivery number n at [n % pgh::kMaxSlabs]
	bool dlv_started = false, dlv_stop = false;
	size_t direct_bytes = 0;                     // slabs up to this size are unpacked straight from the pinned slab (no H2D copy)
	// latency watchdog (started with the first callback unless PERSEUS_GPU_OPT_NO_WATCHDOG)
	std::thread watchdog;
	std::mutex wd_mu;                            // only for wd_cv / wd_stop
	std::condition_variable wd_cv;
	bool wd_started = false, wd_stop = false;
	// bookkeeping
	perseus_gpu_stats stats{};
	int latched = 0;           // first asynchronous error (surfaced at flush/sync/close)
	char latched_msg[512] = "";
};

struct perseus_gpu_plan {
	pg::SegDesc *d_segs = nullptr;
	pg::TileRef *d_tiles = nullptr;            // tiles of segments with 16-byte aligned outputs, then the others
	uint64_t ntiles_stream = 0, ntiles_direct = 0, nsamples = 0, nbytes = 0;
	unsigned fmt = 0;
	int tile_bytes = 0;
};

namespace pgh {

#define CU(h, call)                                                                                               \
	do {                                                                                                          \
		cudaError_t e__ = (call);                                                                                 \
		if (e__ != cudaSuccess) {                                                                                 \
			cudaGetLastError(); /* non-sticky errors must not show up at the next kernel launch check */          \
			return fail(PERSEUS_GPU_CUDAERR, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
		}                                                                                                         \
	} while (0)

// ---- handle.cu
void latch(perseus_gpu *h, int code);
int surface_latched(perseus_gpu *h);
pg::Tuning resolve_tuning(const perseus_gpu_tuning *t);
int check_tuning(const pg::Tuning &t);
int resolve_fmt(unsigned flags, const void *out_i32, const void *out_f32, unsigned *fmt);
enum class Mem { Device, PinnedHost, PageableHost };
Mem classify(const void *p);
int bind(perseus_gpu *h);
int do_launch(perseus_gpu *h, const void *in, size_t nbytes, void *o_i32, void *o_f32, unsigned fmt, cudaStream_t st);
bool tsc_usable();
void tsc_anchor_now(perseus_gpu *h);
bool membarrier_available();
void heavy_barrier(const perseus_gpu *h);
// ---- stream_path.cu
void stop_watchdog(perseus_gpu *h);
void stop_delivery(perseus_gpu *h);
int sync_locked(perseus_gpu *h);
int flush_locked(perseus_gpu *h);

inline uint64_t monotonic_ns()
{
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

#if defined(__x86_64__)
inline uint64_t read_tsc() { return __rdtsc(); }
#else
inline uint64_t read_tsc() { return 0; }
#endif

// CLOCK_MONOTONIC carried forward by the time-stamp counter since the last anchor (handle.cu: tsc_anchor_now)
inline uint64_t callback_now_ns(perseus_gpu *h)
{
	if (!h->use_tsc || !h->ns_per_tick_q32) return monotonic_ns();
	const uint64_t c = read_tsc();
	const uint64_t dc = c - h->tsc_anchor;
	if (c < h->tsc_anchor || dc > h->tsc_reanchor) {
		tsc_anchor_now(h);
		return h->ns_anchor;
	}
	return h->ns_anchor + (uint64_t)(((unsigned __int128)dc * h->ns_per_tick_q32) >> 32);
}

inline void light_barrier(const perseus_gpu *h)      // callback side
{
	if (h->asym) asm volatile("" ::: "memory");
	else std::atomic_thread_fence(std::memory_order_seq_cst);
}

// Exclusive ownership for any thread but the callback's fast path.  Also binds the device (bind = true).
struct Entry {
	perseus_gpu *h = nullptr;
	bool locked = false;
	int rc = 0;
	explicit Entry(perseus_gpu *hh, bool bind_device = true)
	{
		if (!hh) { rc = fail(PERSEUS_GPU_NULLHANDLE, "null handle"); return; }
		h = hh;
		const bool inside_fast_path = h->cb_active.load(std::memory_order_relaxed) &&
		                              h->cb_thread.load(std::memory_order_relaxed) == (unsigned long)pthread_self();
		if (!inside_fast_path) {                 // otherwise: a sink called from the callback -- this thread owns the handle already
			h->mu.lock();
			locked = true;
			if (h->lock_depth++ == 0) {
				h->others_want.fetch_add(1, std::memory_order_seq_cst);
				heavy_barrier(h);
				while (h->cb_active.load(std::memory_order_acquire)) sched_yield();   // a callback that was already running finishes first
			}
		}
		if (bind_device) rc = bind(h);
	}
	~Entry()
	{
		if (!locked) return;
		if (--h->lock_depth == 0) h->others_want.fetch_sub(1, std::memory_order_release);
		h->mu.unlock();
	}
	Entry(const Entry &) = delete;
	Entry &operator=(const Entry &) = delete;
};

}  // namespace pgh
