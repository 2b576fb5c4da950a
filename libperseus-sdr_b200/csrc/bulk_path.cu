// C-ABI layer of libperseus_gpu.so, part 3 of 3 (see handle.h): bulk unpack of whole buffers (device, pinned or pageable memory), batched
// plans, synthetic data, on-device verification, in-run rooflines, autotune.
#include "handle.h"

namespace pgh {

int ensure_bounce(perseus_gpu *h, bool in, bool i32, bool f32)
{
	if (!(in || i32 || f32)) return 0;
	if (!h->pool) {
		h->pool = new (std::nothrow) pg::CopyPool(h->copy_threads - 1);
		if (!h->pool) return fail(PERSEUS_GPU_NOMEM, "out of memory");
	}
	for (int s = 0; s < h->nslots; ++s) {
		if (in && !h->bounce_in[s]) CU(h, cudaHostAlloc(&h->bounce_in[s], h->chunk_bytes, cudaHostAllocDefault));
		if (i32 && !h->bounce_out[s][0]) CU(h, cudaHostAlloc(&h->bounce_out[s][0], h->chunk_bytes / 6 * 8, cudaHostAllocDefault));
		if (f32 && !h->bounce_out[s][1]) CU(h, cudaHostAlloc(&h->bounce_out[s][1], h->chunk_bytes / 6 * 8, cudaHostAllocDefault));
	}
	return 0;
}

int ensure_staging(perseus_gpu *h, bool need_in, bool need_i32, bool need_f32)
{
	for (int s = 0; s < h->nslots; ++s) {
		if (need_in && !h->stage_in[s]) CU(h, cudaMalloc(&h->stage_in[s], h->chunk_bytes));
		if (need_i32 && !h->stage_out[s][0]) CU(h, cudaMalloc(&h->stage_out[s][0], h->chunk_bytes / 6 * 8));
		if (need_f32 && !h->stage_out[s][1]) CU(h, cudaMalloc(&h->stage_out[s][1], h->chunk_bytes / 6 * 8));
		if (!h->ev_in[s]) CU(h, cudaEventCreateWithFlags(&h->ev_in[s], cudaEventDisableTiming));
		if (!h->ev_k[s]) CU(h, cudaEventCreateWithFlags(&h->ev_k[s], cudaEventDisableTiming));
		if (!h->ev_out[s]) CU(h, cudaEventCreateWithFlags(&h->ev_out[s], cudaEventDisableTiming));
	}
	return 0;
}

// Queues the checksum of the piece just unpacked behind its kernel, on the same stream (PERSEUS_GPU_CHECKSUM).
int queue_checksums(perseus_gpu *h, const void *o_i32, const void *o_f32, uint64_t nsamples, uint64_t first_sample, cudaStream_t st)
{
	const void *outs[2] = {o_i32, o_f32};
	for (int k = 0; k < 2; ++k) {
		if (!outs[k] || nsamples == 0) continue;
		cudaError_t e = pg::launch_checksum(outs[k], nsamples * 2, first_sample * 2, h->d_sums + k, h->sm_count, st, /*accumulate=*/true);
		if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "checksum launch failed: %s", cudaGetErrorString(e));
		h->stats.kernel_launches++;
	}
	return 0;
}

void destroy_plan(perseus_gpu_plan *p)
{
	if (p->d_segs) cudaFree(p->d_segs);
	if (p->d_tiles) cudaFree(p->d_tiles);
	delete p;
}

int plan_create_locked(perseus_gpu *h, const perseus_gpu_seg *segs, int nseg, unsigned flags, perseus_gpu_plan **out)
{
	if (!out) return fail(PERSEUS_GPU_ERRPARAM, "null plan pointer");
	*out = nullptr;
	if (nseg < 0 || (nseg > 0 && !segs)) return fail(PERSEUS_GPU_ERRPARAM, "bad segment table");
	unsigned fmt = flags & (PERSEUS_GPU_OUT_INT32 | PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2);
	if (fmt == 0 && nseg > 0) fmt = (segs[0].out_i32 ? PERSEUS_GPU_OUT_INT32 : 0u) | (segs[0].out_f32 ? PERSEUS_GPU_OUT_FLOAT : 0u);
	if ((fmt & PERSEUS_GPU_OUT_FLOAT) && (fmt & PERSEUS_GPU_OUT_FLOAT_POW2))
		return fail(PERSEUS_GPU_ERRPARAM, "OUT_FLOAT and OUT_FLOAT_POW2 are mutually exclusive");
	if (fmt == 0 && nseg > 0) return fail(PERSEUS_GPU_ERRPARAM, "no output requested");

	const int tile = pg::resolve_geometry(h->tune, fmt).tile_bytes;
	std::vector<pg::SegDesc> hs((size_t)nseg);
	// Every segment takes the bulk-copy pipeline with 128-bit stores; outputs that are not 16-byte aligned (an {I,Q}
	// array at its natural 8-byte alignment, or any multiple of 4) get them after a pre-roll of 1..3 output words.  Per
	// segment, so one odd receiver does not slow the batch down.  Wire pointers may have any alignment.  (`slow`: the register-only
	// kernel, only when the handle is tuned to PERSEUS_GPU_VARIANT_DIRECT.)
	std::vector<pg::TileRef> fast, slow;
	uint64_t nsamples = 0, nbytes = 0;
	for (int i = 0; i < nseg; ++i) {
		const perseus_gpu_seg &s = segs[i];
		const uint64_t used = (uint64_t)s.nbytes / 6 * 6;
		if (used && !s.in) return fail(PERSEUS_GPU_ERRPARAM, "segment %d: in is NULL", i);
		if (used && (fmt & PERSEUS_GPU_OUT_INT32) && !s.out_i32) return fail(PERSEUS_GPU_ERRPARAM, "segment %d: out_i32 is NULL", i);
		if (used && (fmt & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2)) && !s.out_f32)
			return fail(PERSEUS_GPU_ERRPARAM, "segment %d: out_f32 is NULL", i);
		if (((uintptr_t)s.out_i32 & 3) || ((uintptr_t)s.out_f32 & 3)) return fail(PERSEUS_GPU_ERRPARAM, "segment %d: outputs must be 4-byte aligned", i);
		pg::SegDesc &d = hs[(size_t)i];
		d = pg::SegDesc{static_cast<const uint8_t *>(s.in), used, (fmt & PERSEUS_GPU_OUT_INT32) ? s.out_i32 : nullptr,
		                (fmt & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2)) ? s.out_f32 : nullptr, 0u, 0u};
		const bool direct = h->tune.variant == PERSEUS_GPU_VARIANT_DIRECT;
		const int pre = direct ? 0 : pg::stream_preroll(d.out_i32, d.out_f32);
		if (pre < 0) d.word_stores = 1u;       // the two outputs at different phases: same kernel, 32-bit stores for this segment's tiles
		uint64_t span = used;                  // wire bytes the segment's tiles cover
		if (pre > 0 && used) {                 // outputs 4m bytes past a 16-byte boundary: the segment starts m words early (kernels.h)
			d.in -= pre;
			d.nbytes = used + (uint64_t)pre;
			if (d.out_i32) d.out_i32 = static_cast<uint8_t *>(d.out_i32) - pre / 3 * 4;
			if (d.out_f32) d.out_f32 = static_cast<uint8_t *>(d.out_f32) - pre / 3 * 4;
			d.preroll = (uint32_t)pre;
			span = d.nbytes;
		}
		const uint64_t nt = (span + (uint64_t)tile - 1) / (uint64_t)tile;
		if (nt > 0xFFFFFFFFull) return fail(PERSEUS_GPU_BUFFERSIZE, "segment %d too large", i);
		std::vector<pg::TileRef> &dst = direct ? slow : fast;
		for (uint64_t t = 0; t < nt; ++t) dst.push_back(pg::TileRef{(uint32_t)i, (uint32_t)t});
		nsamples += used / 6;
		nbytes += used;
	}
	perseus_gpu_plan *p = new (std::nothrow) perseus_gpu_plan();
	if (!p) return fail(PERSEUS_GPU_NOMEM, "out of memory");
	p->ntiles_stream = fast.size();
	p->ntiles_direct = slow.size();
	p->nsamples = nsamples;
	p->nbytes = nbytes;
	p->fmt = fmt;
	p->tile_bytes = tile;
	fast.insert(fast.end(), slow.begin(), slow.end());   // one upload: [stream tiles | direct tiles]
	cudaError_t e = cudaSuccess;
	if (nseg) e = cudaMalloc(&p->d_segs, hs.size() * sizeof(pg::SegDesc));
	if (e == cudaSuccess && !fast.empty()) e = cudaMalloc(&p->d_tiles, fast.size() * sizeof(pg::TileRef));
	if (e == cudaSuccess && nseg) e = cudaMemcpyAsync(p->d_segs, hs.data(), hs.size() * sizeof(pg::SegDesc), cudaMemcpyHostToDevice, h->streams[0]);
	if (e == cudaSuccess && !fast.empty())
		e = cudaMemcpyAsync(p->d_tiles, fast.data(), fast.size() * sizeof(pg::TileRef), cudaMemcpyHostToDevice, h->streams[0]);
	if (e == cudaSuccess) e = cudaStreamSynchronize(h->streams[0]);   // hs/fast die with this frame
	if (e != cudaSuccess) {
		cudaGetLastError();
		destroy_plan(p);
		return fail(PERSEUS_GPU_CUDAERR, "plan upload failed: %s", cudaGetErrorString(e));
	}
	*out = p;
	return 0;
}

int64_t plan_run_locked(perseus_gpu *h, perseus_gpu_plan *p, unsigned flags)
{
	if (!p) return fail(PERSEUS_GPU_ERRPARAM, "null plan");
	int n = 0;   // the plan keeps the tile size and the split it was built with; stages / CTAs per SM follow the handle's current tuning
	cudaError_t e = pg::launch_unpack_batch(p->d_segs, p->d_tiles, p->ntiles_stream, p->d_tiles + p->ntiles_stream, p->ntiles_direct,
	                                        p->tile_bytes, p->fmt, h->tune, h->sm_count, h->streams[0], &n);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "batched unpack launch failed: %s", cudaGetErrorString(e));
	h->stats.kernel_launches += (uint64_t)n;
	h->stats.samples += p->nsamples;
	h->stats.bytes_in += p->nbytes;
	if (!(flags & PERSEUS_GPU_ASYNC)) {
		int rc = sync_locked(h);
		if (rc) return rc;
	}
	return (int64_t)p->nsamples;
}

}  // namespace pgh

using namespace pgh;

// =============================================================================== C ABI

extern "C" {

int64_t perseus_gpu_unpack(perseus_gpu *h, const void *buf, size_t nbytes, void *out_i32, void *out_f32, unsigned flags)
{
	Entry en(h);
	int rc = en.rc;
	if (rc) return rc;
	if (flags & ~(PERSEUS_GPU_OUT_INT32 | PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2 | PERSEUS_GPU_ASYNC | PERSEUS_GPU_CHECKSUM))
		return fail(PERSEUS_GPU_ERRPARAM, "unknown flag bits 0x%x", flags);
	const bool want_sums = flags & PERSEUS_GPU_CHECKSUM;
	unsigned fmt = 0;
	rc = resolve_fmt(flags, out_i32, out_f32, &fmt);
	if (rc) return rc;
	if (!(fmt & PERSEUS_GPU_OUT_INT32)) out_i32 = nullptr;
	if (!(fmt & (PERSEUS_GPU_OUT_FLOAT | PERSEUS_GPU_OUT_FLOAT_POW2))) out_f32 = nullptr;
	const uint64_t ns = nbytes / 6;
	if (ns && !buf) return fail(PERSEUS_GPU_ERRPARAM, "buf is NULL");
	if ((out_i32 && ((uintptr_t)out_i32 & 3)) || (out_f32 && ((uintptr_t)out_f32 & 3)))
		return fail(PERSEUS_GPU_ERRPARAM, "output pointers must be 4-byte aligned");

	if (want_sums) {   // totals restart with this call (even an empty one); everything queued earlier has to be done with them first
		rc = sync_locked(h);
		if (rc) return rc;
		CU(h, cudaMemsetAsync(h->d_sums, 0, 2 * sizeof(unsigned long long), h->streams[0]));
		CU(h, cudaStreamSynchronize(h->streams[0]));
	}
	if (ns == 0) return 0;
	const Mem min_ = classify(buf);
	const Mem mi = out_i32 ? classify(out_i32) : Mem::Device;
	const Mem mf = out_f32 ? classify(out_f32) : Mem::Device;
	const bool in_dev = min_ == Mem::Device, oi_dev = mi == Mem::Device, of_dev = mf == Mem::Device;
	cudaStream_t sk = h->streams[0];

	if (in_dev && oi_dev && of_dev) {
		rc = do_launch(h, buf, ns * 6, out_i32, out_f32, fmt, sk);
		if (rc) return rc;
		if (want_sums && (rc = queue_checksums(h, out_i32, out_f32, ns, 0, sk))) return rc;
	} else {
		// Three-stage pipeline over `nslots` staging slots: copy-in on s_in, kernel (+ checksums) on streams[0], copy-out
		// on s_out, ordered per slot by events.  The copy engines of both directions and the SMs each work on a
		// different chunk at the same time; neither copy stream ever waits behind a copy of the other direction.
		// PAGEABLE host buffers get two more stages at the ends: the caller and the handle's helper threads move each chunk
		// between the application's memory and a pinned bounce buffer of the slot while the engines work on the neighbours.
		const bool host_out = (out_i32 && !oi_dev) || (out_f32 && !of_dev);
		const bool in_page = min_ == Mem::PageableHost && h->copy_threads > 0;
		const bool oi_page = out_i32 && mi == Mem::PageableHost && h->copy_threads > 0;
		const bool of_page = out_f32 && mf == Mem::PageableHost && h->copy_threads > 0;
		rc = ensure_staging(h, !in_dev, out_i32 && !oi_dev, out_f32 && !of_dev);
		if (!rc) rc = ensure_bounce(h, in_page, oi_page, of_page);
		if (rc) return rc;
		const uint8_t *src = static_cast<const uint8_t *>(buf);
		const size_t total = ns * 6;
		struct { bool any; size_t o, on; } pend[kMaxStageSlots] = {};   // outputs waiting in a slot's bounce buffers
		// bounce_out[s] -> the application's memory, once the copy-out that fills it has finished
		auto drain = [&](int s) -> int {
			if (!pend[s].any) return 0;
			CU(h, cudaEventSynchronize(h->ev_out[s]));
			if (oi_page) h->pool->copy(static_cast<uint8_t *>(out_i32) + pend[s].o, h->bounce_out[s][0], pend[s].on, false);
			if (of_page) h->pool->copy(static_cast<uint8_t *>(out_f32) + pend[s].o, h->bounce_out[s][1], pend[s].on, false);
			pend[s].any = false;
			return 0;
		};
		for (size_t off = 0; off < total;) {
			const int s = (int)(h->stage_seq++ % (uint64_t)h->nslots);
			const size_t n = total - off < h->chunk_bytes ? total - off : h->chunk_bytes;
			const size_t o = off / 6 * 8, on = n / 6 * 8;
			if ((rc = drain(s))) return rc;                                   // the chunk that used this slot nslots chunks ago
			const uint8_t *kin = src + off;
			if (!in_dev) {
				const uint8_t *hsrc = src + off;
				if (in_page) {
					CU(h, cudaEventSynchronize(h->ev_in[s]));                 // the copy-in that last read this bounce buffer is done
					h->pool->copy(h->bounce_in[s], hsrc, n, true);
					hsrc = h->bounce_in[s];
				}
				CU(h, cudaStreamWaitEvent(h->s_in, h->ev_k[s], 0));       // the kernel that last read this slot's input is done
				CU(h, cudaMemcpyAsync(h->stage_in[s], hsrc, n, cudaMemcpyHostToDevice, h->s_in));
				CU(h, cudaEventRecord(h->ev_in[s], h->s_in));
				CU(h, cudaStreamWaitEvent(sk, h->ev_in[s], 0));
				h->stats.h2d_bytes += n;
				kin = h->stage_in[s];
			}
			if (host_out) CU(h, cudaStreamWaitEvent(sk, h->ev_out[s], 0));   // the copy-out that last read this slot's outputs is done
			uint8_t *ki = out_i32 ? (oi_dev ? static_cast<uint8_t *>(out_i32) + o : h->stage_out[s][0]) : nullptr;
			uint8_t *kf = out_f32 ? (of_dev ? static_cast<uint8_t *>(out_f32) + o : h->stage_out[s][1]) : nullptr;
			rc = do_launch(h, kin, n, ki, kf, fmt, sk);
			if (rc) return rc;
			if (want_sums && (rc = queue_checksums(h, ki, kf, n / 6, off / 6, sk))) return rc;
			CU(h, cudaEventRecord(h->ev_k[s], sk));
			if (host_out) {
				CU(h, cudaStreamWaitEvent(h->s_out, h->ev_k[s], 0));
				if (out_i32 && !oi_dev) {
					CU(h, cudaMemcpyAsync(oi_page ? h->bounce_out[s][0] : static_cast<uint8_t *>(out_i32) + o, ki, on, cudaMemcpyDeviceToHost, h->s_out));
					h->stats.d2h_bytes += on;
				}
				if (out_f32 && !of_dev) {
					CU(h, cudaMemcpyAsync(of_page ? h->bounce_out[s][1] : static_cast<uint8_t *>(out_f32) + o, kf, on, cudaMemcpyDeviceToHost, h->s_out));
					h->stats.d2h_bytes += on;
				}
				CU(h, cudaEventRecord(h->ev_out[s], h->s_out));
				if (oi_page || of_page) { pend[s].any = true; pend[s].o = o; pend[s].on = on; }
			}
			off += n;
		}
		// pageable outputs are complete when the call returns, PERSEUS_GPU_ASYNC or not (like the CUDA runtime's own pageable copies)
		for (int k = 0; k < h->nslots; ++k)
			if ((rc = drain((int)((h->stage_seq + (uint64_t)k) % (uint64_t)h->nslots)))) return rc;   // oldest chunk first
	}
	if (!(flags & PERSEUS_GPU_ASYNC)) {
		rc = sync_locked(h);
		if (rc) return rc;
	}
	return (int64_t)ns;
}

int perseus_gpu_get_checksums(perseus_gpu *h, uint64_t *sum_i32, uint64_t *sum_f32)
{
	Entry en(h);
	if (en.rc) return en.rc;
	int rc = sync_locked(h);
	if (rc) return rc;
	CU(h, cudaMemcpyAsync(h->h_scratch, h->d_sums, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->streams[0]));
	CU(h, cudaStreamSynchronize(h->streams[0]));
	h->stats.d2h_bytes += 2 * sizeof(unsigned long long);
	if (sum_i32) *sum_i32 = h->h_scratch[0];
	if (sum_f32) *sum_f32 = h->h_scratch[1];
	return 0;
}

// ---- batched ------------------------------------------------------------------------------------

int perseus_gpu_plan_create(perseus_gpu *h, const perseus_gpu_seg *segs, int nseg, unsigned flags, perseus_gpu_plan **out)
{
	Entry en(h);
	if (en.rc) return en.rc;
	return plan_create_locked(h, segs, nseg, flags, out);
}

int64_t perseus_gpu_plan_run(perseus_gpu *h, perseus_gpu_plan *p, unsigned flags)
{
	Entry en(h);
	if (en.rc) return en.rc;
	return plan_run_locked(h, p, flags);
}

int perseus_gpu_plan_destroy(perseus_gpu *h, perseus_gpu_plan *p)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (!p) return 0;
	cudaStreamSynchronize(h->streams[0]);
	destroy_plan(p);
	return 0;
}

int64_t perseus_gpu_unpack_batch(perseus_gpu *h, const perseus_gpu_seg *segs, int nseg, unsigned flags)
{
	Entry en(h);
	if (en.rc) return en.rc;
	perseus_gpu_plan *p = nullptr;
	int rc = plan_create_locked(h, segs, nseg, flags, &p);
	if (rc) return rc;
	int64_t n = plan_run_locked(h, p, 0);   // the plan is freed below, so always synchronous
	cudaStreamSynchronize(h->streams[0]);
	destroy_plan(p);
	return n;
}

int perseus_gpu_autotune(perseus_gpu *h, double *gbs_single, double *gbs_fused)
{
	Entry en(h);
	if (en.rc) return en.rc;
	int rc = sync_locked(h);
	if (rc) return rc;
	const size_t nbytes = (size_t)87381 * 6144;            // 512 MiB of wire: far beyond L2, 0.15-0.3 ms per launch
	const size_t ns = nbytes / 6;
	uint8_t *in = nullptr, *oi = nullptr, *of = nullptr;
	cudaError_t e = cudaMalloc(&in, nbytes);
	if (e == cudaSuccess) e = cudaMalloc(&oi, ns * 8);
	if (e == cudaSuccess) e = cudaMalloc(&of, ns * 8);
	cudaStream_t st = h->streams[0];
	if (e == cudaSuccess) e = cudaMemsetAsync(in, 0x5A, nbytes, st);
	static const pg::Geometry cand[] = {{12288, 2, 1}, {12288, 3, 1}, {12288, 4, 1}, {12288, 5, 1}, {12288, 6, 1}, {6144, 5, 1},
	                                    {6144, 6, 1},  {6144, 8, 1},  {18432, 2, 1}, {18432, 3, 1}, {24576, 2, 1}, {12288, 2, 2}};
	cudaEvent_t e0 = h->tev[0], e1 = h->tev[1];
	double best_gbs[2] = {0.0, 0.0};
	pg::Geometry best[2] = {};
	for (int cls = 0; cls < 2 && e == cudaSuccess; ++cls) {
		const unsigned fmt = cls ? (pg::FMT_I32 | pg::FMT_F32) : pg::FMT_F32;
		for (const pg::Geometry &g : cand) {
			pg::Tuning t{};
			t.store_mode = h->tune.store_mode;
			t.tile_bytes = g.tile_bytes; t.stages = g.stages; t.ctas_per_sm = g.ctas_per_sm;
			int n = 0;
			float ms = 0.f;
			e = pg::launch_unpack(in, nbytes, cls ? oi : nullptr, of, fmt, t, h->sm_count, st, &n);        // warm
			if (e == cudaSuccess) e = cudaEventRecord(e0, st);
			for (int r = 0; r < 3 && e == cudaSuccess; ++r) e = pg::launch_unpack(in, nbytes, cls ? oi : nullptr, of, fmt, t, h->sm_count, st, &n);
			if (e == cudaSuccess) e = cudaEventRecord(e1, st);
			if (e == cudaSuccess) e = cudaEventSynchronize(e1);
			if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
			if (e != cudaSuccess) break;
			h->stats.kernel_launches += 4;
			const double gbs = (cls ? 22.0 : 14.0) * (double)ns * 3.0 / (ms * 1e-3) / 1e9;
			if (gbs > best_gbs[cls]) { best_gbs[cls] = gbs; best[cls] = g; }
		}
	}
	if (in) cudaFree(in);
	if (oi) cudaFree(oi);
	if (of) cudaFree(of);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(PERSEUS_GPU_CUDAERR, "autotune failed: %s", cudaGetErrorString(e));
	}
	h->tune.tuned[0] = best[0];
	h->tune.tuned[1] = best[1];
	if (gbs_single) *gbs_single = best_gbs[0];
	if (gbs_fused) *gbs_fused = best_gbs[1];
	return 0;
}

// ---- synthetic data / verification --------------------------------------------------------------------

int perseus_gpu_generate(perseus_gpu *h, void *dev_dst, size_t nbytes, int pattern, uint64_t seed, uint64_t byte_offset)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (pattern != PERSEUS_SYNTH_RANDOM && pattern != PERSEUS_SYNTH_RAMP) return fail(PERSEUS_GPU_ERRPARAM, "unknown pattern %d", pattern);
	if (pattern == PERSEUS_SYNTH_RAMP && byte_offset % 6) return fail(PERSEUS_GPU_ERRPARAM, "RAMP byte_offset must be a multiple of 6");
	if (nbytes && !dev_dst) return fail(PERSEUS_GPU_ERRPARAM, "null destination");
	cudaError_t e = pg::launch_generate(dev_dst, nbytes, pattern, seed, byte_offset, h->sm_count, h->streams[0]);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "generate launch failed: %s", cudaGetErrorString(e));
	h->stats.kernel_launches += nbytes ? 1 : 0;
	CU(h, cudaStreamSynchronize(h->streams[0]));
	return 0;
}

int perseus_gpu_checksum(perseus_gpu *h, const void *dev_words, size_t nwords, uint64_t first_index, uint64_t *sum)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (!sum || (nwords && !dev_words)) return fail(PERSEUS_GPU_ERRPARAM, "null argument");
	if ((uintptr_t)dev_words & 3) return fail(PERSEUS_GPU_ERRPARAM, "words must be 4-byte aligned");
	cudaError_t e = pg::launch_checksum(dev_words, nwords, first_index, h->d_scratch, h->sm_count, h->streams[0]);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "checksum launch failed: %s", cudaGetErrorString(e));
	h->stats.kernel_launches += nwords ? 1 : 0;
	CU(h, cudaMemcpyAsync(h->h_scratch, h->d_scratch, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->streams[0]));
	CU(h, cudaStreamSynchronize(h->streams[0]));
	*sum = h->h_scratch[0];
	return 0;
}

int perseus_gpu_verify(perseus_gpu *h, const void *dev_in, size_t nbytes, const void *dev_i32, const void *dev_f32, unsigned flags,
                       uint64_t *nmismatch, uint64_t *first_bad_word)
{
	Entry en(h);
	int rc = en.rc;
	if (rc) return rc;
	unsigned fmt = 0;
	rc = resolve_fmt(flags & ~PERSEUS_GPU_ASYNC, dev_i32, dev_f32, &fmt);
	if (rc) return rc;
	cudaError_t e = pg::launch_verify(dev_in, nbytes, dev_i32, dev_f32, fmt, h->d_scratch, h->sm_count, h->streams[0]);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "verify launch failed: %s", cudaGetErrorString(e));
	h->stats.kernel_launches += nbytes / 6 ? 1 : 0;
	CU(h, cudaMemcpyAsync(h->h_scratch, h->d_scratch, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->streams[0]));
	CU(h, cudaStreamSynchronize(h->streams[0]));
	if (nmismatch) *nmismatch = h->h_scratch[0];
	if (first_bad_word) *first_bad_word = h->h_scratch[1];
	if (h->h_scratch[0])
		return fail(PERSEUS_GPU_MISMATCH, "%llu output words differ from the per-sample recomputation (first at word %llu)", h->h_scratch[0], h->h_scratch[1]);
	return 0;
}

int perseus_gpu_probe_hbm(perseus_gpu *h, int kind, size_t nbytes, int reps, double *gbs)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (kind < 0 || kind > 2 || !gbs || reps < 1 || nbytes < (1u << 20)) return fail(PERSEUS_GPU_ERRPARAM, "bad probe arguments");
	nbytes -= nbytes % 16;
	void *a = nullptr, *b = nullptr;
	cudaError_t e = cudaMalloc(&a, nbytes);
	if (e == cudaSuccess && kind == 2) e = cudaMalloc(&b, nbytes);
	if (e != cudaSuccess) {
		if (a) cudaFree(a);
		cudaGetLastError();
		return fail(PERSEUS_GPU_NOMEM, "probe scratch: %s", cudaGetErrorString(e));
	}
	cudaStream_t st = h->streams[0];
	cudaMemsetAsync(a, 0x5A, nbytes, st);
	double best = 0.0;
	cudaEvent_t e0 = h->tev[0], e1 = h->tev[1];
	for (int ctas : {2, 4, 8, 16}) {
		float ms = 0.f;
		e = pg::launch_probe(kind, a, kind == 2 ? b : a, nbytes, h->sm_count, ctas, st);   // warm
		if (e == cudaSuccess) e = cudaEventRecord(e0, st);
		for (int r = 0; r < reps && e == cudaSuccess; ++r) e = pg::launch_probe(kind, a, kind == 2 ? b : a, nbytes, h->sm_count, ctas, st);
		if (e == cudaSuccess) e = cudaEventRecord(e1, st);
		if (e == cudaSuccess) e = cudaEventSynchronize(e1);
		if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
		if (e != cudaSuccess) break;
		h->stats.kernel_launches += (uint64_t)reps + 1;
		const double g = (kind == 2 ? 2.0 : 1.0) * (double)nbytes * reps / (ms * 1e-3) / 1e9;
		if (g > best) best = g;
	}
	cudaFree(a);
	if (b) cudaFree(b);
	if (e != cudaSuccess) return fail(PERSEUS_GPU_CUDAERR, "probe failed: %s", cudaGetErrorString(e));
	*gbs = best;
	return 0;
}

int perseus_gpu_probe_pcie(perseus_gpu *h, int kind, size_t nbytes, size_t d2h_nbytes, int reps, double *h2d_gbs, double *d2h_gbs)
{
	Entry en(h);
	if (en.rc) return en.rc;
	if (d2h_nbytes == 0) d2h_nbytes = nbytes;
	if (kind < 0 || kind > 2 || reps < 1 || nbytes < (1u << 20) || d2h_nbytes < (1u << 20)) return fail(PERSEUS_GPU_ERRPARAM, "bad probe arguments");
	int rc = sync_locked(h);
	if (rc) return rc;
	const bool up = kind != PERSEUS_GPU_PCIE_D2H, down = kind != PERSEUS_GPU_PCIE_H2D;
	void *hu = nullptr, *du = nullptr, *hd = nullptr, *dd = nullptr;
	cudaError_t e = cudaSuccess;
	if (up) {
		e = cudaHostAlloc(&hu, nbytes, cudaHostAllocDefault);
		if (e == cudaSuccess) e = cudaMalloc(&du, nbytes);
		if (e == cudaSuccess) memset(hu, 0x5A, nbytes);
	}
	if (down && e == cudaSuccess) {
		e = cudaHostAlloc(&hd, d2h_nbytes, cudaHostAllocDefault);
		if (e == cudaSuccess) e = cudaMalloc(&dd, d2h_nbytes);
		if (e == cudaSuccess) memset(hd, 0, d2h_nbytes);   // touch the pages before timing
	}
	double best_up = 0.0, best_down = 0.0;
	for (int r = 0; r <= reps && e == cudaSuccess; ++r) {   // r == 0 warms up
		// both directions are queued before either is waited for, each on its own stream, so in DUPLEX mode they overlap
		if (up) {
			e = cudaEventRecord(h->tev[0], h->s_in);
			if (e == cudaSuccess) e = cudaMemcpyAsync(du, hu, nbytes, cudaMemcpyHostToDevice, h->s_in);
			if (e == cudaSuccess) e = cudaEventRecord(h->tev[1], h->s_in);
		}
		if (down && e == cudaSuccess) {
			e = cudaEventRecord(h->tev[2], h->s_out);
			if (e == cudaSuccess) e = cudaMemcpyAsync(hd, dd, d2h_nbytes, cudaMemcpyDeviceToHost, h->s_out);
			if (e == cudaSuccess) e = cudaEventRecord(h->tev[3], h->s_out);
		}
		if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_in);
		if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_out);
		float ms = 0.f;
		if (up && e == cudaSuccess) {
			e = cudaEventElapsedTime(&ms, h->tev[0], h->tev[1]);
			const double g = (double)nbytes / (ms * 1e-3) / 1e9;
			if (r && g > best_up) best_up = g;
		}
		if (down && e == cudaSuccess) {
			e = cudaEventElapsedTime(&ms, h->tev[2], h->tev[3]);
			const double g = (double)d2h_nbytes / (ms * 1e-3) / 1e9;
			if (r && g > best_down) best_down = g;
		}
		if (e == cudaSuccess) {
			if (up) h->stats.h2d_bytes += nbytes;
			if (down) h->stats.d2h_bytes += d2h_nbytes;
		}
	}
	if (hu) cudaFreeHost(hu);
	if (hd) cudaFreeHost(hd);
	if (du) cudaFree(du);
	if (dd) cudaFree(dd);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(PERSEUS_GPU_CUDAERR, "PCIe probe failed: %s", cudaGetErrorString(e));
	}
	if (h2d_gbs) *h2d_gbs = best_up;
	if (d2h_gbs) *d2h_gbs = best_down;
	return 0;
}

}  // extern "C"
