// device_common.cuh — declarations shared by the CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "b2_recon.h"

namespace b2 {

void set_error(const char *fmt, ...);
#define B2_CUDA_OK(expr)                                                                    \
	do {                                                                                    \
		cudaError_t e__ = (expr);                                                           \
		if (e__ != cudaSuccess) {                                                           \
			b2::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
			return -1;                                                                      \
		}                                                                                   \
	} while (0)

// One MPEG-2 picture as the kernel sees it (all pointers are device pointers).
struct M2Job {
	uint8_t *dst_y, *dst_c;
	const uint8_t *fwd_y, *fwd_c, *bwd_y, *bwd_c;
	const b2_m2_mb *mbs;
	const int16_t *coefs;
	int32_t mb_w, mb_h;
};

// Launch reconstruction of n_jobs independent pictures of identical size in one grid.
int launch_mpeg2(cudaStream_t stream, const M2Job *d_jobs, int n_jobs, int mb_w, int mb_h);

// One H.264 picture as the kernels see it (device pointers).  Frames of the pool are addressed
// as frames + idx * frame_bytes (luma) and + W*H (NV12 chroma).
struct H4Job {
	uint8_t *frames;
	size_t frame_bytes;
	const b2_h264_mb *mbs;
	const b2_h264_dbk *dbk;
	const b2_h264_mc *mcs;
	const int16_t *coefs;
	int32_t dst, mb_w, mb_h, n_intra;
	const uint32_t *items;      // b2_h264_picture.items
	uint32_t n_items, n_first;  // b2_h264_picture.n_items / n_items_first
	uint32_t derive_bs;         // b2_h264_picture.bs_on_device
	// b2_h264_picture.mv_on_device: the partitions as parsed; mcs / items / n_items above are then device scratch that
	// h264_mv_kernel and h264_items_kernel fill before the inter kernel runs
	uint32_t derive_mv;
	const b2_h264_part *parts;
	uint32_t n_parts, pad_;
	int8_t *field_ref;          // [n_mbs][16] reference index of every 4x4 block (-1 intra), raster inside the macroblock
	uint32_t *field_mv;         // [n_mbs][16] motion vector (x | y << 16)
	uint32_t *owner;            // [n_parts] macroblock of a partition
};
// bytes of device scratch a picture with mv_on_device needs behind its records: motion records, work list (bound),
// owner table, motion field
inline size_t h264_mv_scratch_bytes(size_t n_mbs, size_t n_parts, size_t item_bound) {
	return n_parts * 32 + (item_bound + 64) * 4 + n_parts * 4 + 256 + n_mbs * 16 + 256 + n_mbs * 64 + 1024;
}
// A picture with few intra macroblocks (P / B pictures) goes through the sparse intra kernel, which
// orders macroblocks by per-macroblock flags; a mostly-intra picture through the row wavefront.
__host__ __device__ inline bool h264_intra_sparse(int n_intra, int n_mbs) { return n_intra * 3 < n_mbs; }
// The inter kernel fetches reference rows as whole 16-byte chunks: the chunk that holds the last bytes of a frame
// pool's last row may be read 16 bytes past it (never used), so pools are allocated with this many spare bytes.
constexpr size_t H4_POOL_TAIL = 256;
// phase: 0 inter (MC + residual), 1 intra wavefront (dense pictures), 2 deblock wavefront,
// 3 intra, sparse pictures, 4 deblocking strengths (the jobs with b2_h264_picture.bs_on_device), 5 motion vectors,
// motion records and work list (the jobs with mv_on_device; d_jobs is written: n_items)
size_t h264_sync_bytes(int n_jobs, int mb_w, int mb_h);
// max_items: largest n_items among the jobs (phase 0 only)
int launch_h264_phase(cudaStream_t stream, int phase, const H4Job *d_jobs, int n_jobs, int mb_w, int mb_h, int *sync_area, uint64_t *launches,
                      uint32_t max_items = 0, uint32_t max_first = 0);
int launch_h264(cudaStream_t stream, const H4Job *d_jobs, int n_jobs, int mb_w, int mb_h, uint32_t max_items, bool intra_dense,
                bool intra_sparse, bool deblock, int *sync_area, uint64_t *launches, uint32_t max_first = 0, bool derive_bs = false, bool derive_mv = false);

// md5.cu: MD5 of a cropped window of n frames (one thread per frame); d_slots / d_out are device pointers
int md5_pictures(cudaStream_t stream, const uint8_t *frames, size_t frame_bytes, const int *d_slots, int n, int stride, int rows, int left, int top, int w, int h,
                 uint8_t *d_out);
// the same for one frame whose planes the application addresses with row pitch `stride` (m2d_frame_t.width)
int md5_frame(cudaStream_t stream, const uint8_t *frames, size_t frame_bytes, int slot, int stride, size_t plane_bytes, int left, int top, int w, int h, uint8_t *d_out);

__device__ __forceinline__ int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

}  // namespace b2
