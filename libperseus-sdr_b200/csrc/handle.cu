// C-ABI layer of libperseus_gpu.so, part 1 of 3 (see handle.h): life cycle, ownership of a handle, the callback's clock, tuning,
// statistics, device plumbing.
#include "handle.h"

namespace pgh {

void latch(perseus_gpu *h, int code)
{
	if (h && !h->latched) {
		h->latched = code;
		snprintf(h->latched_msg, sizeof(h->latched_msg), "%.*s", (int)sizeof(h->latched_msg) - 1, pg::last_error());
	}
}

int surface_latched(perseus_gpu *h)
{
	if (!h->latched) return 0;
	const int c = h->latched;
	fail(c, "%s", h->latched_msg);
	h->latched = 0;
	return c;
}

pg::Tuning resolve_tuning(const perseus_gpu_tuning *t)
{
	pg::Tuning r{};
	if (t) {
		r.variant = t->variant;
		r.tile_bytes = t->tile_bytes;      // 0 = chosen per output format at launch (kernels.h resolve_geometry)
		r.stages = t->stages;
		r.ctas_per_sm = t->ctas_per_sm;
		r.store_mode = t->store_mode;
	}
	if (r.store_mode == 0) r.store_mode = 1;
	return r;
}

int check_tuning(const pg::Tuning &t)
{
	if (t.variant < 0 || t.variant > 2) return fail(PERSEUS_GPU_ERRPARAM, "tuning.variant %d not in 0..2", t.variant);
	if (t.tile_bytes && !pg::valid_tile(t.tile_bytes))
		return fail(PERSEUS_GPU_ERRPARAM, "tuning.tile_bytes %d must be 6144, 9216, 12288, 18432 or 24576", t.tile_bytes);
	if (t.stages && (t.stages < 2 || t.stages > pg::kMaxStages)) return fail(PERSEUS_GPU_ERRPARAM, "tuning.stages %d not in 2..%d", t.stages, pg::kMaxStages);
	if (t.ctas_per_sm < 0 || t.ctas_per_sm > 8) return fail(PERSEUS_GPU_ERRPARAM, "tuning.ctas_per_sm %d not in 1..8", t.ctas_per_sm);
	if (t.store_mode < 1 || t.store_mode > 2) return fail(PERSEUS_GPU_ERRPARAM, "tuning.store_mode %d not in 0..2", t.store_mode);
	// what actually fits, for either default the unset fields could resolve to: 227 KiB of shared memory, 2048 threads per SM
	for (unsigned fmt : {1u, 3u}) {
		const pg::Geometry g = pg::resolve_geometry(t, fmt);
		if ((size_t)g.stages * g.tile_bytes > 200 * 1024)
			return fail(PERSEUS_GPU_ERRPARAM, "tuning: stages*tile_bytes = %zu exceeds 200 KiB of shared memory", (size_t)g.stages * g.tile_bytes);
		const int by_smem = (int)((227 * 1024) / ((size_t)g.stages * (g.tile_bytes + 16) + 1024));
		const int by_threads = 2048 / (pg::kConsumerThreads + pg::kProducerThreads);
		if (g.ctas_per_sm > by_smem || g.ctas_per_sm > by_threads)
			return fail(PERSEUS_GPU_ERRPARAM, "tuning: %d CTAs/SM do not fit (shared memory allows %d, threads allow %d)", g.ctas_per_sm, by_smem, by_threads);
	}
	return 0;
}

// flags -> format bits; 0 means "whatever output pointers are non-NULL"
int resolve_fmt(unsigned flags, const void *out_i32, const void *out_f32, unsigned *fmt)
{
	unsigned f = flags & (PERSEUS_GPU_OUT_INT32 | PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2);
	if (f == 0) f = (out_i32 ? PERSEUS_GPU_OUT_INT32 : 0u) | (out_f32 ? PERSEUS_GPU_OUT_FLOAT : 0u);
	if ((f & PERSEUS_GPU_OUT_FLOAT) && (f & PERSEUS_GPU_OUT_FLOAT_POW2))
		return fail(PERSEUS_GPU_ERRPARAM, "OUT_FLOAT and OUT_FLOAT_POW2 are mutually exclusive");
	if (f == 0) return fail(PERSEUS_GPU_ERRPARAM, "no output requested");
	if ((f & PERSEUS_GPU_OUT_INT32) && !out_i32) return fail(PERSEUS_GPU_ERRPARAM, "OUT_INT32 requested but out_i32 is NULL");
	if ((f & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2)) && !out_f32)
		return fail(PERSEUS_GPU_ERRPARAM, "float output requested but out_f32 is NULL");
	*fmt = f;
	return 0;
}

Mem classify(const void *p)
{
	cudaPointerAttributes a{};
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
		cudaGetLastError();
		return Mem::PageableHost;
	}
	switch (a.type) {
	case cudaMemoryTypeDevice: return Mem::Device;
	case cudaMemoryTypeManaged: return Mem::Device;
	case cudaMemoryTypeHost: return Mem::PinnedHost;
	default: return Mem::PageableHost;
	}
}

int bind(perseus_gpu *h)
{
	if (!h) return fail(PERSEUS_GPU_NULLHANDLE, "null handle");
	CU(h, cudaSetDevice(h->device));
	return 0;
}

int do_launch(perseus_gpu *h, const void *in, size_t nbytes, void *o_i32, void *o_f32, unsigned fmt, cudaStream_t st)
{
	int n = 0;
	cudaError_t e = pg::launch_unpack(in, nbytes, o_i32, o_f32, fmt, h->tune, h->sm_count, st, &n);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "unpack kernel launch failed: %s", cudaGetErrorString(e));
	h->stats.kernel_launches += (uint64_t)n;
	h->stats.samples += nbytes / 6;
	h->stats.bytes_in += nbytes / 6 * 6;
	return 0;
}

// The callback reads the time once per transfer (age bound, eager submission).  clock_gettime costs 40-50 ns on these hosts -- a
// fifth of the whole 6144-byte hand-off -- so between anchors CLOCK_MONOTONIC is carried forward by the invariant time-stamp
// counter (RDTSC, ~8 ns).  The rate is measured against CLOCK_MONOTONIC itself (over perseus_gpu_open, then over every anchor
// interval) and the anchor is renewed about once a second, which bounds the disagreement with the watchdog's clock_gettime
// to microseconds.  Anything unexpected (no invariant TSC, a counter that stalls or jumps) falls back to clock_gettime.
#if defined(__x86_64__)
bool tsc_usable()
{
	static const bool ok = [] {
		const char *off = getenv("PERSEUS_GPU_NO_TSC");
		if (off && *off && *off != '0') return false;
		unsigned a = 0, b = 0, c = 0, d = 0;
		if (!__get_cpuid(0x80000007u, &a, &b, &c, &d)) return false;
		return (d & (1u << 8)) != 0;   // invariant TSC
	}();
	return ok;
}
#else
bool tsc_usable() { return false; }
#endif

void tsc_anchor_now(perseus_gpu *h)
{
	const uint64_t ns = monotonic_ns(), c = read_tsc();
	if (h->tsc_anchor && c > h->tsc_anchor && ns > h->ns_anchor) {
		const uint64_t dt = ns - h->ns_anchor, dc = c - h->tsc_anchor;
		if (dt >= 50000 && dt < (1ull << 31)) {                           // 50 us .. 2 s: a rate worth having, no overflow below
			const uint64_t q = (dt << 32) / dc;
			if (q < (1ull << 20) || q > (1ull << 36)) h->use_tsc = false;   // outside 4 THz .. 60 MHz: not a time-stamp counter
			else {
				h->ns_per_tick_q32 = q;
				h->tsc_reanchor = ((1ull << 30) << 32) / q;               // ~1.07 s worth of ticks
			}
		}
	}
	h->ns_anchor = ns;
	h->tsc_anchor = c;
}

// ---- ownership hand-off between the callback thread and everybody else (see struct perseus_gpu) ----------------------

bool membarrier_available()
{
#if defined(__linux__) && defined(__NR_membarrier)
	static const bool ok = [] {
		const char *off = getenv("PERSEUS_GPU_NO_MEMBARRIER");          // tests: force the fence fallback
		if (off && *off && *off != '0') return false;
		const long cmds = syscall(__NR_membarrier, MEMBARRIER_CMD_QUERY, 0, 0);
		if (cmds < 0 || !(cmds & MEMBARRIER_CMD_PRIVATE_EXPEDITED)) return false;
		return syscall(__NR_membarrier, MEMBARRIER_CMD_REGISTER_PRIVATE_EXPEDITED, 0, 0) == 0;
	}();
	return ok;
#else
	return false;
#endif
}

void heavy_barrier(const perseus_gpu *h)             // everybody else
{
#if defined(__linux__) && defined(__NR_membarrier)
	if (h->asym) {
		if (syscall(__NR_membarrier, MEMBARRIER_CMD_PRIVATE_EXPEDITED, 0, 0) != 0) {
			// cannot happen after a successful registration; without the barrier the hand-off would be unsound
			fprintf(stderr, "perseus-gpu: membarrier(PRIVATE_EXPEDITED) failed\n");
			abort();
		}
		return;
	}
#endif
	std::atomic_thread_fence(std::memory_order_seq_cst);
}

}  // namespace pgh

using namespace pgh;

// =============================================================================== C ABI

extern "C" {

#define PG_STR2(x) #x
#define PG_STR(x) PG_STR2(x)
const char *perseus_gpu_version(void) { return "perseus-gpu abi " PG_STR(PERSEUS_GPU_ABI_VERSION) ", sm_100a, " __DATE__; }

int perseus_gpu_device_count(void)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(PERSEUS_GPU_NODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
	}
	return n;
}

int perseus_gpu_device_info(int device, char *name, size_t name_len, int *sm_count, int *cc_major, int *cc_minor, uint64_t *total_mem)
{
	cudaDeviceProp p{};
	cudaError_t e = cudaGetDeviceProperties(&p, device);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(PERSEUS_GPU_NODEVICE, "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
	}
	if (name && name_len) snprintf(name, name_len, "%s", p.name);
	if (sm_count) *sm_count = p.multiProcessorCount;
	if (cc_major) *cc_major = p.major;
	if (cc_minor) *cc_minor = p.minor;
	if (total_mem) *total_mem = p.totalGlobalMem;
	return 0;
}

int perseus_gpu_open(perseus_gpu **out, const perseus_gpu_config *ucfg)
{
	if (!out) return fail(PERSEUS_GPU_ERRPARAM, "null handle pointer");
	*out = nullptr;
	perseus_gpu_config cfg{};
	if (ucfg) {
		if (ucfg->struct_size < 8 || ucfg->struct_size > sizeof(cfg))
			return fail(PERSEUS_GPU_ERRPARAM, "perseus_gpu_config.struct_size %u not understood (this library: %zu)", ucfg->struct_size, sizeof(cfg));
		memcpy(&cfg, ucfg, ucfg->struct_size);
	}
	int ndev = perseus_gpu_device_count();
	if (ndev <= 0) return ndev < 0 ? ndev : fail(PERSEUS_GPU_NODEVICE, "no CUDA device");
	if (cfg.device < 0 || cfg.device >= ndev) return fail(PERSEUS_GPU_ERRPARAM, "device %d out of range (0..%d)", cfg.device, ndev - 1);
	cudaDeviceProp prop{};
	CU(nullptr, cudaGetDeviceProperties(&prop, cfg.device));
	if (prop.major != 10)
		return fail(PERSEUS_GPU_BADARCH, "device %d (%s) is sm_%d%d; this library carries sm_100a code only and has no fallback", cfg.device, prop.name,
		            prop.major, prop.minor);

	pg::Tuning tune = resolve_tuning(&cfg.tuning);
	int rc = check_tuning(tune);
	if (rc) return rc;
	unsigned sfmt = cfg.stream_flags ? cfg.stream_flags : PERSEUS_GPU_OUT_INT32;
	if (sfmt & ~(PERSEUS_GPU_OUT_INT32 | PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2))
		return fail(PERSEUS_GPU_ERRPARAM, "stream_flags 0x%x has unknown bits", sfmt);
	if ((sfmt & PERSEUS_GPU_OUT_FLOAT) && (sfmt & PERSEUS_GPU_OUT_FLOAT_POW2))
		return fail(PERSEUS_GPU_ERRPARAM, "stream_flags: OUT_FLOAT and OUT_FLOAT_POW2 are mutually exclusive");
	if (cfg.options & ~PERSEUS_GPU_OPT_NO_WATCHDOG) return fail(PERSEUS_GPU_ERRPARAM, "options 0x%x has unknown bits", cfg.options);
	const uint32_t nslabs = cfg.nslabs ? cfg.nslabs : 4;
	if (nslabs < 2 || nslabs > (uint32_t)kMaxSlabs) return fail(PERSEUS_GPU_ERRPARAM, "nslabs %u not in 2..%d", nslabs, kMaxSlabs);
	const uint32_t nstreams = cfg.nstreams ? cfg.nstreams : 2;
	if (nstreams < 1 || nstreams > (uint32_t)kMaxStreams) return fail(PERSEUS_GPU_ERRPARAM, "nstreams %u not in 1..%d", nstreams, kMaxStreams);
	const uint32_t nslots = cfg.stage_slots ? cfg.stage_slots : 3;
	if (nslots < 2 || nslots > (uint32_t)kMaxStageSlots) return fail(PERSEUS_GPU_ERRPARAM, "stage_slots %u not in 2..%d", nslots, kMaxStageSlots);
	uint64_t slab = cfg.slab_bytes ? cfg.slab_bytes : (8ull << 20);
	slab -= slab % 48;
	uint64_t chunk = cfg.chunk_bytes ? cfg.chunk_bytes : (32ull << 20);
	// Whole pages on both sides of the link: 12288 wire bytes (3 pages) become 16384 output bytes (4 pages), so every staged
	// copy starts and ends on a page boundary of the caller's pinned buffers.  The copy engines need that: chunks that are
	// only 128-byte aligned reach 52.5-53.5 GB/s device->host, page-aligned ones 56.1-56.6 (tools/d2h_probe.py).
	chunk -= chunk % (chunk >= 12288 ? 12288 : 48);
	if (slab < 48 || chunk < 48) return fail(PERSEUS_GPU_BUFFERSIZE, "slab_bytes/chunk_bytes must be at least 48");

	perseus_gpu *h = new (std::nothrow) perseus_gpu();
	if (!h) return fail(PERSEUS_GPU_NOMEM, "out of memory");
	h->device = cfg.device;
	h->sm_count = prop.multiProcessorCount;
	h->cfg = cfg;
	h->tune = tune;
	h->stream_fmt = sfmt;
	h->nslabs = (int)nslabs;
	h->nstreams = (int)nstreams;
	h->nslots = (int)nslots;
	h->slab_bytes = (size_t)slab;
	h->max_latency_ns = cfg.max_latency_us == 0xFFFFFFFFu ? 0 : (uint64_t)(cfg.max_latency_us ? cfg.max_latency_us : 50000u) * 1000ull;
	h->chunk_bytes = (size_t)chunk;
	// only with a latency bound: max_latency_us = 0xFFFFFFFF asks for full slabs and nothing else
	h->eager_gap_ns = (!h->max_latency_ns || cfg.eager_gap_us == 0xFFFFFFFFu) ? 0 : (uint64_t)(cfg.eager_gap_us ? cfg.eager_gap_us : 100u) * 1000ull;
	{
		const unsigned hw = std::thread::hardware_concurrency();
		const unsigned autot = hw >= 4 ? (hw / 2 > 8 ? 8 : hw / 2) : 1;
		h->copy_threads = cfg.copy_threads == 0xFFFFFFFFu ? 0 : (int)(cfg.copy_threads ? (cfg.copy_threads > 64 ? 64 : cfg.copy_threads) : autot);
	}
	h->direct_bytes = cfg.direct_bytes == 0xFFFFFFFFu ? 0 : cfg.direct_bytes ? (size_t)cfg.direct_bytes : kDefaultDirectBytes;
	h->asym = membarrier_available();
	h->use_tsc = tsc_usable();
	if (h->use_tsc) tsc_anchor_now(h);   // first reading; the rate comes from the second one, at the end of this function
	auto bail = [&](int code) {   // free what exists, keep the message of the original failure
		const std::string keep = pg::last_error();
		perseus_gpu_close(h);
		pg::set_last_error(keep.c_str());
		return code;
	};
	cudaError_t e = cudaSetDevice(h->device);
	if (e != cudaSuccess) return bail(fail(PERSEUS_GPU_CUDAERR, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e)));
	for (int s = 0; s < h->nstreams; ++s) {
		e = cudaStreamCreateWithFlags(&h->streams[s], cudaStreamNonBlocking);
		if (e != cudaSuccess) return bail(fail(PERSEUS_GPU_CUDAERR, "cudaStreamCreate: %s", cudaGetErrorString(e)));
	}
	e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_dlv, cudaStreamNonBlocking);
	if (e != cudaSuccess) return bail(fail(PERSEUS_GPU_CUDAERR, "cudaStreamCreate: %s", cudaGetErrorString(e)));
	for (int k = 0; k < kEventSlots && e == cudaSuccess; ++k) e = cudaEventCreate(&h->events[k]);
	for (int k = 0; k < 4 && e == cudaSuccess; ++k) e = cudaEventCreate(&h->tev[k]);
	if (e != cudaSuccess) return bail(fail(PERSEUS_GPU_CUDAERR, "cudaEventCreate: %s", cudaGetErrorString(e)));
	e = cudaMalloc(&h->d_scratch, 2 * sizeof(unsigned long long));
	if (e == cudaSuccess) e = cudaMalloc(&h->d_sums, 2 * sizeof(unsigned long long));
	if (e == cudaSuccess) e = cudaMemset(h->d_sums, 0, 2 * sizeof(unsigned long long));
	if (e == cudaSuccess) e = cudaHostAlloc(&h->h_scratch, 2 * sizeof(unsigned long long), cudaHostAllocDefault);
	if (e != cudaSuccess) return bail(fail(PERSEUS_GPU_CUDAERR, "scratch allocation: %s", cudaGetErrorString(e)));
	if (h->use_tsc) {
		while (monotonic_ns() - h->ns_anchor < 60000) { }   // streams and events took milliseconds; make sure of 60 us anyway
		tsc_anchor_now(h);
		if (!h->ns_per_tick_q32) h->use_tsc = false;
	}
	*out = h;
	return ok();
}

int perseus_gpu_close(perseus_gpu *h)
{
	if (!h) return fail(PERSEUS_GPU_NULLHANDLE, "null handle");
	stop_watchdog(h);
	int rc = 0;
	{
		Entry en(h, false);
		if (cudaSetDevice(h->device) == cudaSuccess) {
			rc = flush_locked(h);
			stop_delivery(h);
			for (int k = 0; k < kMaxSlabs; ++k) {
				Slab &s = h->slabs[k];
				if (s.host) cudaFreeHost(s.host);
				for (uint8_t *p : s.host_out)
					if (p) cudaFreeHost(p);
				if (s.dev_in) cudaFree(s.dev_in);
				if (s.dev_i32) cudaFree(s.dev_i32);
				if (s.dev_f32) cudaFree(s.dev_f32);
				if (s.done) cudaEventDestroy(s.done);
				if (s.unpacked) cudaEventDestroy(s.unpacked);
				if (s.ready) cudaEventDestroy(s.ready);
			}
			for (int s = 0; s < kMaxStageSlots; ++s) {
				if (h->stage_in[s]) cudaFree(h->stage_in[s]);
				if (h->stage_out[s][0]) cudaFree(h->stage_out[s][0]);
				if (h->stage_out[s][1]) cudaFree(h->stage_out[s][1]);
				if (h->ev_in[s]) cudaEventDestroy(h->ev_in[s]);
				if (h->ev_k[s]) cudaEventDestroy(h->ev_k[s]);
				if (h->ev_out[s]) cudaEventDestroy(h->ev_out[s]);
				if (h->bounce_in[s]) cudaFreeHost(h->bounce_in[s]);
				if (h->bounce_out[s][0]) cudaFreeHost(h->bounce_out[s][0]);
				if (h->bounce_out[s][1]) cudaFreeHost(h->bounce_out[s][1]);
			}
			if (h->d_scratch) cudaFree(h->d_scratch);
			if (h->d_sums) cudaFree(h->d_sums);
			if (h->h_scratch) cudaFreeHost(h->h_scratch);
			for (int k = 0; k < kEventSlots; ++k)
				if (h->events[k]) cudaEventDestroy(h->events[k]);
			for (int k = 0; k < 4; ++k)
				if (h->tev[k]) cudaEventDestroy(h->tev[k]);
			for (int s = 0; s < h->nstreams; ++s)
				if (h->streams[s]) cudaStreamDestroy(h->streams[s]);
			if (h->s_in) cudaStreamDestroy(h->s_in);
			if (h->s_out) cudaStreamDestroy(h->s_out);
			if (h->s_dlv) cudaStreamDestroy(h->s_dlv);
			cudaGetLastError();
		}
		stop_delivery(h);   // also when the device could not be bound above: the thread must be gone before the handle is
		if (h->fout) {
			if ((h->fout_is_stdout ? fflush(h->fout) : fclose(h->fout)) != 0 && !rc) rc = fail(PERSEUS_GPU_IOERROR, "closing stream file failed");
		}
	}
	delete h->pool;   // joins the helper threads
	delete h;
	return rc;
}

int perseus_gpu_sync(perseus_gpu *h)
{
	Entry en(h);
	if (en.rc) return en.rc;
	return sync_locked(h);
}

int perseus_gpu_get_stats(perseus_gpu *h, perseus_gpu_stats *out)
{
	Entry en(h, false);
	if (en.rc) return en.rc;
	if (!out) return fail(PERSEUS_GPU_ERRPARAM, "null stats pointer");
	*out = h->stats;
	out->host_blocks = h->host_blocks.load(std::memory_order_relaxed);
	return 0;
}

int perseus_gpu_set_tuning(perseus_gpu *h, const perseus_gpu_tuning *t)
{
	if (!h) return fail(PERSEUS_GPU_NULLHANDLE, "null handle");
	pg::Tuning r = resolve_tuning(t);
	int rc = check_tuning(r);
	if (rc) return rc;
	Entry en(h, false);
	memcpy(r.tuned, h->tune.tuned, sizeof(r.tuned));   // autotune results survive; explicit fields take precedence anyway
	h->tune = r;
	return 0;
}

int perseus_gpu_get_geometry(perseus_gpu *h, unsigned flags, int *tile_bytes, int *stages, int *ctas_per_sm)
{
	if (!h) return fail(PERSEUS_GPU_NULLHANDLE, "null handle");
	Entry en(h, false);
	const pg::Geometry g = pg::resolve_geometry(h->tune, flags & 7u);
	if (tile_bytes) *tile_bytes = g.tile_bytes;
	if (stages) *stages = g.stages;
	if (ctas_per_sm) *ctas_per_sm = g.ctas_per_sm;
	return 0;
}

int perseus_gpu_get_tuning(perseus_gpu *h, perseus_gpu_tuning *t)
{
	if (!h) return fail(PERSEUS_GPU_NULLHANDLE, "null handle");
	if (!t) return fail(PERSEUS_GPU_ERRPARAM, "null tuning pointer");
	Entry en(h, false);
	memset(t, 0, sizeof(*t));
	t->variant = h->tune.variant;
	t->tile_bytes = h->tune.tile_bytes;
	t->stages = h->tune.stages;
	t->ctas_per_sm = h->tune.ctas_per_sm;
	t->store_mode = h->tune.store_mode;
	return 0;
}

// ---- plumbing ---------------------------------------------------------------------------------------

void *perseus_gpu_dev_alloc(perseus_gpu *h, size_t nbytes)
{
	Entry en(h);
	if (en.rc) return nullptr;
	void *p = nullptr;
	cudaError_t e = cudaMalloc(&p, nbytes ? nbytes : 1);
	if (e != cudaSuccess) {
		cudaGetLastError();
		fail(PERSEUS_GPU_NOMEM, "cudaMalloc(%zu): %s", nbytes, cudaGetErrorString(e));
		return nullptr;
	}
	return p;
}

int perseus_gpu_dev_free(perseus_gpu *h, void *p)
{
	Entry en(h);
	if (en.rc) return en.rc;
	CU(h, cudaFree(p));
	return 0;
}

void *perseus_gpu_host_alloc(perseus_gpu *h, size_t nbytes)
{
	Entry en(h);
	if (en.rc) return nullptr;
	void *p = nullptr;
	cudaError_t e = cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocPortable);
	if (e != cudaSuccess) {
		cudaGetLastError();
		fail(PERSEUS_GPU_NOMEM, "cudaHostAlloc(%zu): %s", nbytes, cudaGetErrorString(e));
		return nullptr;
	}
	return p;
}

int perseus_gpu_host_free(perseus_gpu *h, void *p)
{
	Entry en(h);
	if (en.rc) return en.rc;
	CU(h, cudaFreeHost(p));
	return 0;
}

int perseus_gpu_memcpy(perseus_gpu *h, void *dst, const void *src, size_t nbytes)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (nbytes == 0) return 0;
	CU(h, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDefault, h->streams[0]));
	CU(h, cudaStreamSynchronize(h->streams[0]));
	return 0;
}

int perseus_gpu_memset(perseus_gpu *h, void *dev, int byte, size_t nbytes)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (nbytes == 0) return 0;
	CU(h, cudaMemsetAsync(dev, byte, nbytes, h->streams[0]));
	CU(h, cudaStreamSynchronize(h->streams[0]));
	return 0;
}

void *perseus_gpu_get_stream(perseus_gpu *h, int idx)
{
	if (!h || idx < 0 || idx >= h->nstreams) {
		fail(PERSEUS_GPU_ERRPARAM, "stream index out of range");
		return nullptr;
	}
	return (void *)h->streams[idx];
}

int perseus_gpu_event_record(perseus_gpu *h, int slot)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (slot < 0 || slot >= kEventSlots) return fail(PERSEUS_GPU_ERRPARAM, "event slot %d not in 0..%d", slot, kEventSlots - 1);
	CU(h, cudaEventRecord(h->events[slot], h->streams[0]));
	return 0;
}

int perseus_gpu_event_elapsed_ms(perseus_gpu *h, int a, int b, float *ms)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (a < 0 || a >= kEventSlots || b < 0 || b >= kEventSlots || !ms) return fail(PERSEUS_GPU_ERRPARAM, "bad event slots");
	CU(h, cudaEventSynchronize(h->events[b]));
	CU(h, cudaEventElapsedTime(ms, h->events[a], h->events[b]));
	return 0;
}

}  // extern "C"
