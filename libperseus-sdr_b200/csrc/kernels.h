// Internal interface between the C-ABI layer (handle.cu, stream_path.cu, bulk_path.cu) and the sm_100a kernels
// (unpack_kernels.cu).  Not installed; the public surface is include/perseus-gpu.h.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "host_common.h"

namespace pg {

// output-format bits, same values as PERSEUS_GPU_OUT_* in include/perseus-gpu.h
enum : unsigned { FMT_I32 = 1u, FMT_F32 = 2u, FMT_POW2 = 4u };

constexpr int kConsumerThreads = 256;                  // threads that convert + store, per CTA
constexpr int kProducerThreads = 32;                   // one warp; lane 0 issues the bulk copies
constexpr int kMaxStages       = 8;

struct Geometry { int tile_bytes, stages, ctas_per_sm; };

struct Tuning {            // copy of perseus_gpu_tuning; 0 in tile_bytes/stages/ctas_per_sm = pick by output format
	int variant, tile_bytes, stages, ctas_per_sm, store_mode;
	Geometry tuned[2];     // perseus_gpu_autotune results: [0] one output format, [1] int32+float fused; zeros = none
};

// Pipeline geometry actually used for a launch.  The defaults come from the sweep committed in
// profiles/ (round 1): what matters is the wire bytes in flight per SM -- about 48 KiB when one
// 8-byte output is written per sample, about 36 KiB when both are (fewer read bytes per byte of
// traffic); more read-ahead than that starves the write stream and costs 5-10 % of HBM bandwidth.
inline Geometry resolve_geometry(const Tuning &t, unsigned fmt)
{
	const bool fused = (fmt & FMT_I32) && (fmt & (FMT_F32 | FMT_POW2));
	const Geometry &m = t.tuned[fused ? 1 : 0];             // measured on this device, if the caller asked for it
	Geometry g;                                             // explicit settings win, then measured, then the defaults
	// fused: 3 stages of two 6144-byte transfers (36 KiB in flight); one format: 8 stages of one transfer (48 KiB) -- the same
	// bytes in flight as 4 x 12288 but 0.5 % faster in the round-2 sweep (profiles/r2_sweep_fine.jsonl)
	g.tile_bytes = t.tile_bytes ? t.tile_bytes : m.tile_bytes ? m.tile_bytes : (fused ? 12288 : 6144);
	g.stages = t.stages ? t.stages : m.stages ? m.stages : (fused ? 3 : (g.tile_bytes == 6144 ? 8 : 4));
	g.ctas_per_sm = t.ctas_per_sm ? t.ctas_per_sm : m.ctas_per_sm ? m.ctas_per_sm : 1;
	return g;
}
inline bool valid_tile(int tile) { return tile == 6144 || tile == 9216 || tile == 12288 || tile == 18432 || tile == 24576; }

// One tile of a batched launch: which segment, which tile of it.
struct TileRef { uint32_t seg, tile; };
// Device-side copy of a perseus_gpu_seg; nbytes is rounded down to whole samples by the host.  For a segment that takes
// the pipeline with a pre-roll (stream_preroll() = 3m > 0) the pointers and the size are the virtual ones: moved back by
// 3m wire bytes / 4m output bytes, 3m bytes longer.  word_stores: the two outputs sit at different phases, the pipeline
// writes this segment with 32-bit stores.
struct SegDesc { const uint8_t *in; uint64_t nbytes; void *out_i32; void *out_f32; uint32_t preroll, word_stores; };

// 0 / 3 / 6 / 9: 128-bit stores are legal for these output pointers after that pre-roll; -1: only 32-bit stores are.
int stream_preroll(const void *out_i32, const void *out_f32);

// Flat unpack of nbytes/6 samples.  Picks the kernel from `t.variant` and the pointers'
// alignment; returns the number of kernels it launched through *launches.
cudaError_t launch_unpack(const void *in, size_t nbytes, void *out_i32, void *out_f32, unsigned fmt,
                          const Tuning &t, int sm_count, cudaStream_t stream, int *launches);

// Batched unpack over segments described on the device; the caller built two tile maps with tile size
// `tile_bytes`: `d_tiles_stream` lists the tiles that go through the bulk-copy pipeline (every segment, whatever the
// alignment of its pointers: SegDesc.preroll / word_stores say how its stores are made, so one odd receiver costs only
// its own tiles) and `d_tiles_direct` the tiles for the register-only kernel (PERSEUS_GPU_VARIANT_DIRECT, the A/B
// variant).  The two launches, when both lists are non-empty, go back to back on `stream`.
cudaError_t launch_unpack_batch(const SegDesc *d_segs, const TileRef *d_tiles_stream, uint64_t ntiles_stream,
                                const TileRef *d_tiles_direct, uint64_t ntiles_direct, int tile_bytes, unsigned fmt,
                                const Tuning &t, int sm_count, cudaStream_t stream, int *launches);

cudaError_t launch_generate(void *dst, size_t nbytes, int pattern, uint64_t seed, uint64_t byte_offset, int sm_count,
                            cudaStream_t stream);
// *d_sum (device; zeroed first unless `accumulate`) += checksum of nwords 32-bit words
cudaError_t launch_checksum(const void *words, size_t nwords, uint64_t first_index, unsigned long long *d_sum, int sm_count,
                            cudaStream_t stream, bool accumulate = false);
// d_result[0] = mismatching words, d_result[1] = lowest mismatching word index (callee initialises)
cudaError_t launch_verify(const void *in, size_t nbytes, const void *out_i32, const void *out_f32, unsigned fmt,
                          unsigned long long *d_result, int sm_count, cudaStream_t stream);

// One-directional HBM streams for in-run roofline context: kind 0 = read src, 1 = write dst, 2 = copy src -> dst.
cudaError_t launch_probe(int kind, const void *src, void *dst, size_t nbytes, int sm_count, int ctas_per_sm, cudaStream_t stream);

}  // namespace pg
