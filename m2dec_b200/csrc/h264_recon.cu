// h264_recon.cu — K1..K4: H.264 picture reconstruction on sm_100a.
//
//   h264_inter5_kernel       K2+K1  motion compensation (luma 6-tap quarter-pel, chroma eighth-pel bilinear, edge
//                                   replication by coordinate clamp, both lists, weighted prediction) fused with
//                                   dequantisation + 4x4 / 8x8 inverse transform + residual add/clip for every
//                                   inter macroblock; work item = one vertical strip of a motion record, from a
//                                   host-sorted list; each output byte is written once.
//   h264_wave_kernel<1>      K3+K1  intra 4x4 / 8x8 / 16x16 / chroma prediction + residual, and I_PCM, as a
//   h264_intra_sparse_kernel        persistent MB-row wavefront (I pictures) / per-macroblock flags (P, B pictures).
//   h264_deblock3_kernel     K4     in-loop filter from the per-MB strength records: two pictures (packed 16x2 lanes)
//   h264_wave_kernel<2>             and two macroblock rows per warp (batched launches) / one picture per launch
//                                   (plugin path).
//   h264_bs_kernel                  deblocking strengths from the motion records: pictures whose
//                                   host left them out (b2_h264_picture.bs_on_device)
//   h264_mv_kernel           \      motion vector prediction in a macroblock-row wavefront, motion records and the inter
//   h264_items_kernel        /      kernel's work list: pictures with mv_on_device (Baseline streams)
//
// Arithmetic restated from the reference (src/lib/h264.cpp): dequant tables 964-1054 and
// coeff_writeback 2005-2022; 4x4 transform 2145-2269; 8x8 transform 3942-4080; DC transforms
// 4309-4404; intra predictors 2510-2997, 3315-3929 (8x8), 4224-4304, 4561-4705; I_PCM 4736-4761; luma MC 5381-6406 (UMV
// 5932-6116); chroma MC 4859-5330; deblock 10253-10663 with tables h264vld.h:4536-4627.
// HBM-bound integer/byte work: no tensor cores.
#include <stdlib.h>
#include "device_common.cuh"

namespace b2 {
namespace {

__constant__ uint8_t c_blk_x[16] = {0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3};
__constant__ uint8_t c_blk_y[16] = {0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3};
// Intra4x4 blocks a warp's two halves predict in the same step (-1: the half idles)
__constant__ int8_t c_i4_pair[2][10] = {{0, 1, 2, 3, 6, 7, 10, 11, 14, 15}, {-1, -1, 4, 5, 8, 9, 12, 13, -1, -1}};
__constant__ uint8_t c_r2b[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};
__constant__ uint8_t c_norm_adjust[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};
__constant__ uint8_t c_norm_adjust8[6][6] = {{20, 18, 32, 19, 25, 24}, {22, 19, 35, 21, 28, 26}, {26, 23, 42, 24, 33, 31},
                                             {28, 25, 45, 26, 35, 33}, {32, 28, 51, 30, 40, 38}, {36, 32, 58, 34, 46, 43}};
// class of position (x & 3) + 4 * (y & 3) of an 8x8 block in the normAdjust8x8 pattern (h264.cpp:1036-1046)
__constant__ uint8_t c_cls8[16] = {0, 3, 4, 3, 3, 1, 5, 1, 4, 5, 2, 5, 3, 1, 5, 1};
__constant__ uint8_t c_alpha[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20,
	22, 25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255};
__constant__ uint8_t c_beta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 8, 8,
	9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18};
__constant__ uint8_t c_tc0[52][3] = {
	{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0},
	{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 1, 1},
	{0, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 2, 3}, {1, 2, 3},
	{2, 2, 3}, {2, 2, 4}, {2, 3, 4}, {2, 3, 4}, {3, 3, 5}, {3, 4, 6}, {3, 4, 6}, {4, 5, 7}, {4, 5, 8}, {4, 6, 9}, {5, 7, 10},
	{6, 8, 11}, {6, 8, 13}, {7, 10, 14}, {8, 11, 16}, {9, 12, 18}, {10, 13, 20}, {11, 15, 23}, {13, 17, 25}};

__device__ __forceinline__ int level_scale4(int qp, int pos) {
	int x = pos & 3, y = pos >> 2;
	int cls = (!(x & 1) && !(y & 1)) ? 0 : (((x & 1) && (y & 1)) ? 1 : 2);
	return (int)c_norm_adjust[qp % 6][cls] << (qp / 6);
}

__device__ __forceinline__ const int16_t *rec_of(const H4Job &job, const b2_h264_mb &m, int bit) {
	return job.coefs + ((size_t)m.coef_idx + __popc(m.nz & ((1u << bit) - 1))) * 16;
}

// 4x4 inverse transform of dequantised coefficients (DC already carries +32); res = value >> 6
__device__ __forceinline__ void idct4x4(int *c, int16_t *res) {
#pragma unroll
	for (int y = 0; y < 4; ++y) {
		int d0 = c[4 * y], d1 = c[4 * y + 1], d2 = c[4 * y + 2], d3 = c[4 * y + 3];
		int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
		c[4 * y] = e0 + e3; c[4 * y + 1] = e1 + e2; c[4 * y + 2] = e1 - e2; c[4 * y + 3] = e0 - e3;
	}
#pragma unroll
	for (int x = 0; x < 4; ++x) {
		int d0 = c[x], d1 = c[4 + x], d2 = c[8 + x], d3 = c[12 + x];
		int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
		res[x] = (int16_t)((e0 + e3) >> 6); res[4 + x] = (int16_t)((e1 + e2) >> 6);
		res[8 + x] = (int16_t)((e1 - e2) >> 6); res[12 + x] = (int16_t)((e0 - e3) >> 6);
	}
}

// Residual of block `b` of a macroblock (0..15 luma in decode order, 16..19 Cb, 20..23 Cr) into
// res[16] (raster 4x4).  Returns false (res untouched) when the block has no residual at all.
__device__ bool block_residual(const H4Job &job, const b2_h264_mb &m, int b, int16_t *res) {
	int c[16];
	bool has_ac;
	int dc = 0;
	bool has_dc = false;
	int qp;
	const int16_t *lv;
	if (b < 16) {
		qp = m.qp_y;
		has_ac = (m.nz >> (1 + b)) & 1;
		lv = rec_of(job, m, 1 + b);
		if (m.type == B2_H4_I16x16 && (m.nz & 1)) {
			// intra16x16_dc_transform: 4x4 Hadamard of the dequantised DC levels, (x + 2) >> 2
			const int16_t *dl = rec_of(job, m, 0);
			int f[16];
			int ls = level_scale4(qp, 0);
#pragma unroll
			for (int i = 0; i < 16; ++i) f[i] = dl[i] * ls;
#pragma unroll
			for (int y = 0; y < 4; ++y) {
				int a0 = f[4 * y], a1 = f[4 * y + 1], a2 = f[4 * y + 2], a3 = f[4 * y + 3];
				f[4 * y] = a0 + a1 + a2 + a3; f[4 * y + 1] = a0 + a1 - a2 - a3;
				f[4 * y + 2] = a0 - a1 - a2 + a3; f[4 * y + 3] = a0 - a1 + a2 - a3;
			}
			int bx = c_blk_x[b], by = c_blk_y[b];
			int a0 = f[bx], a1 = f[4 + bx], a2 = f[8 + bx], a3 = f[12 + bx];
			int v = by == 0 ? a0 + a1 + a2 + a3 : (by == 1 ? a0 + a1 - a2 - a3 : (by == 2 ? a0 - a1 - a2 + a3 : a0 - a1 + a2 - a3));
			dc = (v + 2) >> 2;
			has_dc = true;
		}
	} else {
		int comp = (b - 16) >> 2, k = (b - 16) & 3;
		qp = comp ? m.qp_cr : m.qp_cb;
		has_ac = (m.nz >> (2 + b)) & 1;
		lv = rec_of(job, m, 2 + b);
		if (m.nz & (1u << 17)) {
			// intra_chroma_dc_transform: 2x2 Hadamard of the dequantised DC levels, >> 1
			const int16_t *dl = rec_of(job, m, 17) + comp * 4;
			int ls = level_scale4(qp, 0);
			int f0 = dl[0] * ls, f1 = dl[1] * ls, f2 = dl[2] * ls, f3 = dl[3] * ls;
			int v = k == 0 ? f0 + f1 + f2 + f3 : (k == 1 ? f0 - f1 + f2 - f3 : (k == 2 ? f0 + f1 - f2 - f3 : f0 - f1 - f2 + f3));
			dc = v >> 1;
			has_dc = true;
		}
	}
	if (!has_ac && !(has_dc && dc != 0)) return false;
	if (has_ac) {
		int4 r0 = __ldg(reinterpret_cast<const int4 *>(lv)), r1 = __ldg(reinterpret_cast<const int4 *>(lv) + 1);
		int16_t l[16];
		*reinterpret_cast<int4 *>(l) = r0;
		*reinterpret_cast<int4 *>(l + 8) = r1;
		// the three scale factors of the position classes, looked up once per block
		const int sh = qp / 6, rem = qp - 6 * sh;
		const int s0 = (int)c_norm_adjust[rem][0] << sh, s1 = (int)c_norm_adjust[rem][1] << sh, s2 = (int)c_norm_adjust[rem][2] << sh;
#pragma unroll
		for (int i = 0; i < 16; ++i) {
			const int x = i & 3, y = i >> 2;
			c[i] = l[i] * ((!(x & 1) && !(y & 1)) ? s0 : (((x & 1) && (y & 1)) ? s1 : s2));
		}
	} else {
#pragma unroll
		for (int i = 0; i < 16; ++i) c[i] = 0;
	}
	if (has_dc) c[0] = dc;
	c[0] += 32;
	idct4x4(c, res);
	return true;
}

// one dimension of the 8x8 inverse transform (8.5.13; ac8x8transform_interim, h264.cpp:3942-3983)
__device__ __forceinline__ void idct8_1d(int x0, int x1, int x2, int x3, int x4, int x5, int x6, int x7, int *o) {
	int a0 = x0 + x4, a4 = x0 - x4, a2 = (x2 >> 1) - x6, a6 = x2 + (x6 >> 1);
	int b0 = a0 + a6, b2 = a4 + a2, b4 = a4 - a2, b6 = a0 - a6;
	int a1 = x5 - x3 - x7 - (x7 >> 1), a3 = x1 + x7 - x3 - (x3 >> 1), a5 = x5 + (x5 >> 1) + x7 - x1, a7 = x3 + x5 + x1 + (x1 >> 1);
	int b1 = a1 + (a7 >> 2), b7 = a7 - (a1 >> 2), b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5;
	o[0] = b0 + b7; o[1] = b2 + b5; o[2] = b4 + b3; o[3] = b6 + b1;
	o[4] = b6 - b1; o[5] = b4 - b3; o[6] = b2 - b5; o[7] = b0 - b7;
}

// Residual of the 8x8-transform luma block q (64 levels, raster, in four consecutive records)
// into the four 4x4 residual slots res[4q .. 4q+3] (sub-block i at (i & 1, i >> 1), raster 4x4).
// Dequantisation as qp_matrix8x8 (h264.cpp:997-1054): normAdjust8x8 << (qp/6 - 2), and the
// reference's truncating right shift of the table entry below qp 12.
__device__ void block_residual8x8(const H4Job &job, const b2_h264_mb &m, int q, int16_t (*res)[16]) {
	const int16_t *lv = rec_of(job, m, 1 + 4 * q);
	const int qp = m.qp_y, sh = qp / 6 - 2;
	int t[64];
#pragma unroll
	for (int y = 0; y < 8; ++y) {
		int4 raw = __ldg(reinterpret_cast<const int4 *>(lv) + y);
		int16_t l[8];
		*reinterpret_cast<int4 *>(l) = raw;
		int c[8];
#pragma unroll
		for (int x = 0; x < 8; ++x) {
			int na = c_norm_adjust8[qp % 6][c_cls8[(x & 3) + 4 * (y & 3)]];
			c[x] = l[x] * (sh >= 0 ? na << sh : na >> -sh);
		}
		if (y == 0) c[0] += 32;
		idct8_1d(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], &t[8 * y]);
	}
#pragma unroll
	for (int x = 0; x < 8; ++x) {
		int o[8];
		idct8_1d(t[x], t[8 + x], t[16 + x], t[24 + x], t[32 + x], t[40 + x], t[48 + x], t[56 + x], o);
#pragma unroll
		for (int y = 0; y < 8; ++y) res[4 * q + (y >> 2) * 2 + (x >> 2)][(y & 3) * 4 + (x & 3)] = (int16_t)(o[y] >> 6);
	}
}

// ------------------------------------------------------------------------------------------------
// inter prediction
// ------------------------------------------------------------------------------------------------

// ---- helpers of the inter kernel: packed reference rows, 6-tap filters --------------------------------

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
	int d;
	asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
#define PACK4(a, b, c, d) ((uint32_t)((a) & 255) | ((uint32_t)((b) & 255) << 8) | ((uint32_t)((c) & 255) << 16) | ((uint32_t)((d) & 255) << 24))

// (1,-5,20,20,-5,1) across window bytes c..c+5 for output columns c = 0..3 of a 12-byte window
template <int BIAS = 0>
__device__ __forceinline__ void hfilter4(const uint32_t pk[3], int out[4]) {
	out[0] = dp4a_us(pk[0], PACK4(1, -5, 20, 20), dp4a_us(pk[1], PACK4(-5, 1, 0, 0), BIAS));
	out[1] = dp4a_us(pk[0], PACK4(0, 1, -5, 20), dp4a_us(pk[1], PACK4(20, -5, 1, 0), BIAS));
	out[2] = dp4a_us(pk[0], PACK4(0, 0, 1, -5), dp4a_us(pk[1], PACK4(20, 20, -5, 1), BIAS));
	out[3] = dp4a_us(pk[0], PACK4(0, 0, 0, 1), dp4a_us(pk[1], PACK4(-5, 20, 20, -5), dp4a_us(pk[2], PACK4(1, 0, 0, 0), BIAS)));
}

// vertical 6-tap on two packed 16-bit lanes; result lanes carry a +2560 bias
__device__ __forceinline__ uint32_t vtap_packed(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f) {
	return (a + f) + 20u * (c + d) + 0x0A000A00u - 5u * (b + e);
}
// clip255((v + 16) >> 5) of both lanes at once (lanes carry v + 2560, and 2560 = 80 * 32): results in bytes 0 and 2
__device__ __forceinline__ uint32_t vfinish2(uint32_t t) {
	const uint32_t u = ((t + 0x00100010u) >> 5) & 0x07FF07FFu;
	return __vimin_s16x2_relu(__vsub2(u, 0x00500050u), 0x00FF00FFu);
}

// ------------------------------------------------------------------------------------------------
// inter kernel, strip form.  Work item = a vertical strip of one motion-compensation record: 8 (or 4)
// luma samples wide, the full height of the partition (4..16 rows), plus its chroma.  The strip is
// produced row by row, so reference rows, filter state and code are shared by all rows of the strip
// instead of being set up again for every 4x4 block; the residual of the blocks the strip crosses is
// transformed by the same thread into a private shared-memory scratch, so there is no separate
// residual phase and only the sort needs block-wide barriers.
// ------------------------------------------------------------------------------------------------
#ifndef B2_I5_THREADS
#define B2_I5_THREADS 128
#endif
#ifndef B2_I5_CTAS
#define B2_I5_CTAS 4
#endif
constexpr int I5_THREADS = B2_I5_THREADS;

__device__ __forceinline__ int frac_class(int fx, int fy) {
	if (!fx && !fy) return 0;
	if (!fy) return 1;
	if (!fx) return 2;
	if (fx == 2) return 4;          // b / s with j, or j alone
	if (fy == 2) return 5;          // h / m with j
	return 3;                       // b / s with h / m
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async16_ca(void *smem, const void *gmem) {      // through L1: neighbouring strips share sectors
	asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

// ---- reference rows as packed words -------------------------------------------------------------------
__device__ __noinline__ uint4 load_words_clamped(const uint8_t *__restrict__ row, int W, int x, int nwords) {
	uint32_t out[4] = {0, 0, 0, 0};
#pragma unroll
	for (int i = 0; i < 16; ++i)
		if (i < nwords * 4) out[i >> 2] |= (uint32_t)__ldg(row + min(max(x + i, 0), W - 1)) << ((i & 3) * 8);
	return make_uint4(out[0], out[1], out[2], out[3]);
}
// NV12 chroma row: clamping is per sample pair, bytes keep their component (x is even)
__device__ __noinline__ uint4 load_chroma_clamped(const uint8_t *__restrict__ row, int W, int x, int nwords) {
	uint32_t out[4] = {0, 0, 0, 0};
	const int cw = W >> 1;
#pragma unroll
	for (int i = 0; i < 16; ++i)
		if (i < nwords * 4) out[i >> 2] |= (uint32_t)__ldg(row + min(max((x >> 1) + (i >> 1), 0), cw - 1) * 2 + (i & 1)) << ((i & 3) * 8);
	return make_uint4(out[0], out[1], out[2], out[3]);
}
__device__ __forceinline__ bool words_fast(int x, int n, int W) { return x >= 0 && (x & ~3) + 4 * (n + 1) <= W; }

// Reference rows of a strip: N words (N <= 4) of bytes x .. x + 4N - 1 of rows y0, y0 + 1, ... (rows clamped into the
// plane).  The ncu source view of the register-prefetch form (next row's LDGs issued one row ahead) had 30 % of the
// kernel's stall samples on the first use of those loads: a row of work is shorter than a trip to L2 / HBM, and the
// registers for a deeper prefetch are not there (96 per thread at 20 warps per SM).  Here the rows travel through a
// per-lane ring in shared memory instead: the two 16-byte chunks that hold the window are copied asynchronously
// (cp.async, I5_DEPTH - 1 rows ahead, no registers in flight, 2 requests per row instead of N + 1), and read back as two
// 128-bit words from which the window is cut with selects and funnel shifts.  The chunk behind a row's last bytes
// may belong to the next row or to the pool's spare tail (H4_POOL_TAIL): it is fetched, never used.
// Strips whose window leaves the picture (fast == false; bucketed apart by the host) load clamped bytes as before.
constexpr int I5_DEPTH = 4;
template <int N, bool CHROMA>
struct RowFetch {
	const uint8_t *ref;
	int W, H, x;
	bool fast;
	unsigned o;
	uint4 *ring;            // this lane's entry of chunk 0 of slot 0; chunk c of slot s is 32 * (2 s + c) entries on
	__device__ __forceinline__ RowFetch(const uint8_t *ref_, int W_, int H_, int x_, uint4 *ring_)
	    : ref(ref_), W(W_), H(H_), x(x_), fast(words_fast(x_, N, W_)), o((unsigned)x_ & 15u), ring(ring_) {}
	__device__ __forceinline__ const uint8_t *rowp(int yy) const { return ref + (size_t)min(max(yy, 0), H - 1) * W; }
	// row yy is the k-th row of the strip
	__device__ __forceinline__ void issue(int k, int yy, bool on) {
		if (on && fast) {
			const uint8_t *a = rowp(yy) + (x & ~15);
			uint4 *d = ring + 64 * (k & (I5_DEPTH - 1));
			cp_async16_ca(d, a);
			cp_async16_ca(d + 32, a + 16);
		}
		cp_async_commit();
	}
	__device__ __forceinline__ void start(int y0, int total) {
#pragma unroll
		for (int i = 0; i < I5_DEPTH - 1; ++i) issue(i, y0 + i, i < total);
	}
	// asks for row k + I5_DEPTH - 1 (if the strip has one) and hands out row k
	__device__ __forceinline__ void next(int k, int y0, int total, uint32_t *out) {
		issue(k + I5_DEPTH - 1, y0 + k + I5_DEPTH - 1, k + I5_DEPTH - 1 < total);
		if (fast) {
			cp_async_wait<I5_DEPTH - 1>();
			const uint4 *d = ring + 64 * (k & (I5_DEPTH - 1));
			const uint4 c0 = d[0], c1 = d[32];
			uint32_t w[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
			if (o & 8) {
#pragma unroll
				for (int j = 0; j <= N + 1; ++j) w[j] = w[j + 2];
			}
			if (o & 4) {
#pragma unroll
				for (int j = 0; j <= N; ++j) w[j] = w[j + 1];
			}
			const unsigned sh = (o & 3) * 8;
#pragma unroll
			for (int j = 0; j < N; ++j) out[j] = __funnelshift_r(w[j], w[j + 1], sh);
		} else {
			const uint4 v = CHROMA ? load_chroma_clamped(rowp(y0 + k), W, x, N) : load_words_clamped(rowp(y0 + k), W, x, N);
			out[0] = v.x;
			if (N > 1) out[1] = v.y;
			if (N > 2) out[2] = v.z;
			if (N > 3) out[3] = v.w;
		}
	}
};
__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) { return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24); }
__device__ __forceinline__ int dp2a_lo_us(uint32_t a, uint32_t b, int c) {
	int d;
	asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
// four clipped horizontal half samples b = clip((b1 + 16) >> 5) of window words pk[0..2], packed
__device__ __forceinline__ uint32_t hhalf4(const uint32_t pk[3]) {
	int hb[4];
	hfilter4<16>(pk, hb);           // the rounding term rides in the accumulator
	return pack4(clip255(hb[0] >> 5), clip255(hb[1] >> 5), clip255(hb[2] >> 5), clip255(hb[3] >> 5));
}
// four clipped vertical half samples from six rows split into even / odd byte lanes, packed
// (the six rows are a ring: row i of the window is entry (J + 1 + i) % 6, J = the entry written last)
__device__ __forceinline__ uint32_t vhalf4(const uint32_t E[6], const uint32_t O[6], const int J) {
	const uint32_t te = vtap_packed(E[(J + 1) % 6], E[(J + 2) % 6], E[(J + 3) % 6], E[(J + 4) % 6], E[(J + 5) % 6], E[J]);
	const uint32_t to = vtap_packed(O[(J + 1) % 6], O[(J + 2) % 6], O[(J + 3) % 6], O[(J + 4) % 6], O[(J + 5) % 6], O[J]);
	return __byte_perm(vfinish2(te), vfinish2(to), 0x6240);
}

// Luma prediction of a strip NQ words wide and h rows high at integer position (x, y) + fraction (fx, fy),
// produced row by row: sink(r, words) is called for r = 0 .. h - 1 in order.  Classes 4 and 5 (the centre
// sample j) exist for NQ == 1 only: their filter state would not fit the registers at 8 samples.
template <int NQ, class Sink>
__device__ __forceinline__ void i5_luma(const uint8_t *__restrict__ ref, int W, int H, int x, int y, int fx, int fy, int h, uint4 *ring, Sink &&sink) {
	const int cls = frac_class(fx, fy);
	const uint32_t M = 0x00FF00FFu;
	if (cls == 0) {
		RowFetch<NQ, false> f(ref, W, H, x, ring);
		f.start(y, h);
		for (int r = 0; r < h; ++r) {
			uint32_t pw[NQ];
			f.next(r, y, h, pw);
			sink(r, pw);
		}
	} else if (cls == 1) {
		RowFetch<NQ + 2, false> f(ref, W, H, x - 2, ring);
		f.start(y, h);
		for (int r = 0; r < h; ++r) {
			uint32_t w[NQ + 2], pw[NQ];
			f.next(r, y, h, w);
#pragma unroll
			for (int q = 0; q < NQ; ++q) {
				uint32_t b = hhalf4(&w[q]);
				if (fx != 2) b = __vavgu4(__funnelshift_r(w[q], w[q + 1], fx == 1 ? 16 : 24), b);
				pw[q] = b;
			}
			sink(r, pw);
		}
	} else if (cls == 2) {
		RowFetch<NQ, false> f(ref, W, H, x, ring);
		uint32_t E[NQ][6], O[NQ][6];
#pragma unroll
		for (int q = 0; q < NQ; ++q)
#pragma unroll
			for (int j = 0; j < 6; ++j) E[q][j] = O[q][j] = 0;
		f.start(y - 2, h + 5);
		for (int k = 0; k < h + 5; ++k) {
			uint32_t w[NQ], pw[NQ];
			f.next(k, y - 2, h + 5, w);
#pragma unroll
			for (int q = 0; q < NQ; ++q) {
#pragma unroll
				for (int j = 0; j < 5; ++j) { E[q][j] = E[q][j + 1]; O[q][j] = O[q][j + 1]; }
				E[q][5] = w[q] & M;
				O[q][5] = (w[q] >> 8) & M;
			}
			if (k < 5) continue;
#pragma unroll
			for (int q = 0; q < NQ; ++q) {
				uint32_t hw = vhalf4(E[q], O[q], 5);
				if (fy != 2) hw = __vavgu4(fy == 1 ? (E[q][2] | (O[q][2] << 8)) : (E[q][3] | (O[q][3] << 8)), hw);
				pw[q] = hw;
			}
			sink(k - 5, pw);
		}
	} else if (cls == 3) {
		RowFetch<NQ + 2, false> f(ref, W, H, x - 2, ring);
		const bool xo = fx == 3, yo = fy == 3;
		uint32_t E[NQ][6], O[NQ][6], bq[NQ][4];
#pragma unroll
		for (int q = 0; q < NQ; ++q) {
#pragma unroll
			for (int j = 0; j < 6; ++j) E[q][j] = O[q][j] = 0;
#pragma unroll
			for (int j = 0; j < 4; ++j) bq[q][j] = 0;
		}
		f.start(y - 2, h + 5);
		for (int k = 0; k < h + 5; ++k) {
			uint32_t w[NQ + 2], pw[NQ];
			f.next(k, y - 2, h + 5, w);
#pragma unroll
			for (int q = 0; q < NQ; ++q) {
#pragma unroll
				for (int j = 0; j < 5; ++j) { E[q][j] = E[q][j + 1]; O[q][j] = O[q][j + 1]; }
#pragma unroll
				for (int j = 0; j < 3; ++j) bq[q][j] = bq[q][j + 1];
				const uint32_t vr = __funnelshift_r(w[q], w[q + 1], xo ? 24 : 16);      // samples at x + xo
				E[q][5] = vr & M;
				O[q][5] = (vr >> 8) & M;
				bq[q][3] = hhalf4(&w[q]);                                            // b of this row
			}
			if (k < 5) continue;
			// output row r = k - 5: vertical half sample over rows r .. r + 5, b of row r + 2 + yo = k - 3 + yo
#pragma unroll
			for (int q = 0; q < NQ; ++q) pw[q] = __vavgu4(yo ? bq[q][1] : bq[q][0], vhalf4(E[q], O[q], 5));
			sink(k - 5, pw);
		}
	} else if (NQ == 1) {
		// centre sample j from unrounded horizontal intermediates (biased to unsigned 16 bit, two vertically
		// adjacent rows per register) with dp2a down the column; cls 4: alone or averaged with b / s,
		// cls 5: averaged with the vertical half sample h / m
		RowFetch<3, false> f(ref, W, H, x - 2, ring);
		const bool xo = fx == 3, yo = fy == 3, with_b = cls == 4 && fy != 2, with_h = cls == 5;
		uint32_t PW[5][4], prevb[4] = {0, 0, 0, 0}, bq[4] = {0, 0, 0, 0}, E[6], O[6];
#pragma unroll
		for (int j = 0; j < 5; ++j)
#pragma unroll
			for (int c = 0; c < 4; ++c) PW[j][c] = 0;
#pragma unroll
		for (int j = 0; j < 6; ++j) E[j] = O[j] = 0;
		const uint32_t T0 = (1u & 255u) | ((uint32_t)(-5 & 255) << 8), T1 = 20u | (20u << 8), T2 = (uint32_t)(-5 & 255) | (1u << 8);
		f.start(y - 2, h + 5);
		for (int k = 0; k < h + 5; ++k) {
			uint32_t w[3];
			f.next(k, y - 2, h + 5, w);
			int hb[4];
			hfilter4(w, hb);
#pragma unroll
			for (int j = 0; j < 4; ++j)
#pragma unroll
				for (int c = 0; c < 4; ++c) PW[j][c] = PW[j + 1][c];
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				const uint32_t cur = (uint32_t)(hb[c] + 2560);
				PW[4][c] = prevb[c] | (cur << 16);
				prevb[c] = cur;
			}
			if (with_b) {
#pragma unroll
				for (int j = 0; j < 3; ++j) bq[j] = bq[j + 1];
				bq[3] = pack4(clip255((hb[0] + 16) >> 5), clip255((hb[1] + 16) >> 5), clip255((hb[2] + 16) >> 5), clip255((hb[3] + 16) >> 5));
			}
			if (with_h) {
#pragma unroll
				for (int j = 0; j < 5; ++j) { E[j] = E[j + 1]; O[j] = O[j + 1]; }
				const uint32_t vr = __funnelshift_r(w[0], w[1], xo ? 24 : 16);
				E[5] = vr & M;
				O[5] = (vr >> 8) & M;
			}
			if (k < 5) continue;
			int jv[4];
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				const int j1 = dp2a_lo_us(PW[0][c], T0, dp2a_lo_us(PW[2][c], T1, dp2a_lo_us(PW[4][c], T2, -32 * 2560)));
				jv[c] = clip255((j1 + 512) >> 10);
			}
			uint32_t pw = pack4(jv[0], jv[1], jv[2], jv[3]);
			if (with_b) pw = __vavgu4(yo ? bq[1] : bq[0], pw);
			if (with_h) pw = __vavgu4(vhalf4(E, O, 5), pw);
			sink(k - 5, &pw);
		}
	}
}

// Chroma prediction (eighth-sample bilinear, 16-bit lanes): NQ words (Cb Cr Cb Cr) per row, `rows` rows, at
// byte column xb (even) / chroma row cy of the picture
template <int NQ, class Sink>
__device__ __forceinline__ void i5_chroma(const uint8_t *__restrict__ refc, int W, int CH, int xb, int cy, int mvx, int mvy, int rows, uint4 *ring, Sink &&sink) {
	const uint32_t M = 0x00FF00FFu;
	const uint32_t cfx = mvx & 7, cfy = mvy & 7;
	const int x = xb + (mvx >> 3) * 2, y = cy + (mvy >> 3);
	RowFetch<NQ + 1, true> f(refc, W, CH, x, ring);
	uint32_t pcb[NQ], pcr[NQ];
#pragma unroll
	for (int q = 0; q < NQ; ++q) pcb[q] = pcr[q] = 0;
	f.start(y, rows + 1);
	for (int k = 0; k <= rows; ++k) {
		uint32_t w[NQ + 1], pw[NQ];
		f.next(k, y, rows + 1, w);
#pragma unroll
		for (int q = 0; q < NQ; ++q) {
			const uint32_t a = w[q], b = __funnelshift_r(w[q], w[q + 1], 16);
			const uint32_t hcb = (8u - cfx) * (a & M) + cfx * (b & M), hcr = (8u - cfx) * ((a >> 8) & M) + cfx * ((b >> 8) & M);
			const uint32_t vcb = ((8u - cfy) * pcb[q] + cfy * hcb + 0x00200020u) >> 6, vcr = ((8u - cfy) * pcr[q] + cfy * hcr + 0x00200020u) >> 6;
			pw[q] = (vcb & M) | ((vcr & M) << 8);
			pcb[q] = hcb;
			pcr[q] = hcr;
		}
		if (k) sink(k - 1, pw);
	}
}

// One strip, in phases.  The class-specific predictors only park their rows in shared memory (rows = this
// thread's column of the row buffer, word k at rows[k * I5_THREADS]); weights and the second list are
// folded in there; the residual is then added by the whole warp at once: every lane lists the coded
// blocks of its strip in a per-warp queue and the lanes take the queued blocks 32 at a time (dequantise,
// transform, add into the owner's parked rows), so that the transform runs with full warps although only
// a quarter of the blocks carries a residual.  Strips of 8x8-transform macroblocks (rare) keep their
// residual in the owning thread.
constexpr int I5_QUEUE = 256;         // queue entries per warp: 32 lanes x (up to 8 luma blocks | 2 chroma block rows)

template <int NQ>
__device__ __forceinline__ void i5_first_list(const uint8_t *__restrict__ ref, uint8_t *dst_y, uint8_t *dst_c, int W, int H, int px0, int py0,
                                              int cyp, int mvx, int mvy, int h, uint4 *ring) {
	// first list of a bi-predicted strip: straight to the destination, the second list is combined with it
	i5_luma<NQ>(ref, W, H, px0 + (mvx >> 2), py0 + (mvy >> 2), mvx & 3, mvy & 3, h, ring, [&](int r, const uint32_t *pw) {
#pragma unroll
		for (int q = 0; q < NQ; ++q) *reinterpret_cast<uint32_t *>(dst_y + (size_t)(py0 + r) * W + px0 + 4 * q) = pw[q];
	});
	i5_chroma<NQ>(ref + (size_t)W * H, W, H >> 1, px0, cyp, mvx, mvy, h / 2, ring, [&](int r, const uint32_t *pw) {
#pragma unroll
		for (int q = 0; q < NQ; ++q) *reinterpret_cast<uint32_t *>(dst_c + (size_t)(cyp + r) * W + px0 + 4 * q) = pw[q];
	});
}

template <int NQ>
__device__ __forceinline__ void i5_park_luma(const uint8_t *__restrict__ ref, int W, int H, int px0, int py0, int mvx, int mvy, int h, uint32_t *rows, uint4 *ring) {
	i5_luma<NQ>(ref, W, H, px0 + (mvx >> 2), py0 + (mvy >> 2), mvx & 3, mvy & 3, h, ring, [&](int r, const uint32_t *pw) {
#pragma unroll
		for (int q = 0; q < NQ; ++q) rows[(r * NQ + q) * I5_THREADS] = pw[q];
	});
}
template <int NQ>
__device__ __forceinline__ void i5_park_chroma(const uint8_t *__restrict__ refc, int W, int CH, int px0, int cyp, int mvx, int mvy, int rows_n, uint32_t *rows, uint4 *ring) {
	i5_chroma<NQ>(refc, W, CH, px0, cyp, mvx, mvy, rows_n, ring, [&](int r, const uint32_t *pw) {
#pragma unroll
		for (int q = 0; q < NQ; ++q) rows[(r * NQ + q) * I5_THREADS] = pw[q];
	});
}

// four samples + four residuals, clipped
__device__ __forceinline__ uint32_t add_res4(uint32_t w, int r0, int r1, int r2, int r3) {
	return pack4(clip255((int)(w & 255) + r0), clip255((int)((w >> 8) & 255) + r1), clip255((int)((w >> 16) & 255) + r2), clip255((int)(w >> 24) + r3));
}

// MODE 0: one launch does everything (the first list of a bi-predicted strip inside the thread) -- used when no
// picture of the launch has bi-predicted strips, where it is the same code as MODE 2 but allocates its registers
// better (895 vs 825 GB/s on P pictures); MODE 1: first-list passes; MODE 2: finishing passes.
template <int MODE>
__global__ void __launch_bounds__(I5_THREADS, B2_I5_CTAS)
h264_inter5_kernel(const H4Job *__restrict__ jobs) {
	constexpr bool first_pass = MODE == 1;
	// dynamic shared memory (more than the 48 KB a kernel may declare statically)
	extern __shared__ __align__(16) uint8_t s_dyn[];
	uint32_t (*s_rows)[I5_THREADS] = reinterpret_cast<uint32_t(*)[I5_THREADS]>(s_dyn);                       // per thread: up to 16 rows x 2 words of prediction
	int16_t (*s_rs)[66] = reinterpret_cast<int16_t(*)[66]>(s_dyn + sizeof(uint32_t) * 32 * I5_THREADS);     // per thread: residual scratch
	uint16_t (*s_queue)[I5_QUEUE] = reinterpret_cast<uint16_t(*)[I5_QUEUE]>(s_dyn + sizeof(uint32_t) * 32 * I5_THREADS + sizeof(int16_t) * 66 * I5_THREADS);
	const H4Job job = jobs[blockIdx.y];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t it = blockIdx.x * I5_THREADS + tid;
	// first_pass: the first-list passes of bi-predicted strips (items [0, n_first)); else one item per strip
	// for its last list (items [n_first, n_items)).  Each part is sorted by the class of ITS list.
	const uint32_t base = MODE == 2 ? job.n_first : 0u, count = MODE == 0 ? job.n_items : (MODE == 1 ? job.n_first : job.n_items - job.n_first);
	// whole warps beyond the list leave; in the list's last warp the lanes without an item stay (as strips of
	// height 0) for the warp-wide parts
	if ((it & ~31u) >= count) return;
	const bool valid = it < count;
	// the list comes sorted by (class, height) from the host: a warp runs one interpolation path
	const uint32_t item = __ldg(job.items + base + (valid ? it : count - 1));
	const int mb = item & 0xfffff, ri = (item >> 20) & 15, k = (item >> 24) & 3, wide = ((item >> 26) & 15) >= 6;
	const b2_h264_mb m = job.mbs[mb];
	const b2_h264_mc *mcp = job.mcs + m.mc_idx + ri;
	b2_h264_mc mc;
	reinterpret_cast<int4 *>(&mc)[0] = __ldg(reinterpret_cast<const int4 *>(mcp));
	reinterpret_cast<int4 *>(&mc)[1] = __ldg(reinterpret_cast<const int4 *>(mcp) + 1);
	const int nq = wide ? 2 : 1;
	const int x4 = (mc.pos & 3) + (wide ? 2 * k : k), y4 = mc.pos >> 2, h4 = valid ? mc.size >> 4 : 0, h = h4 * 4;
	const int mbx = mb % job.mb_w, mby = mb / job.mb_w;
	const int W = job.mb_w * 16, H = job.mb_h * 16;
	const size_t luma_size = (size_t)W * H;
	uint8_t *dst_y = job.frames + job.frame_bytes * job.dst;
	uint8_t *dst_c = dst_y + luma_size;
	const int px0 = mbx * 16 + x4 * 4, py0 = mby * 16 + y4 * 4;
	const int cy0 = y4 * 2, cx0 = x4 * 2, cyp = mby * 8 + cy0;         // chroma sample position inside the MB / chroma row in the picture
	const bool bi = mc.ref_frame[0] >= 0 && mc.ref_frame[1] >= 0;
	const int first = mc.ref_frame[0] >= 0 ? 0 : 1, lastl = bi ? 1 : first;
	const bool t8 = m.flags & B2_H4_T8x8;
	const bool plain_w = !(mc.log2wd_y | mc.log2wd_c | (mc.w[0][0] ^ 1) | (mc.w[0][1] ^ 1) | (mc.w[0][2] ^ 1) | (mc.w[1][0] ^ 1) | (mc.w[1][1] ^ 1) | (mc.w[1][2] ^ 1) |
	                       mc.o[0][0] | mc.o[0][1] | mc.o[0][2] | mc.o[1][0] | mc.o[1][1] | mc.o[1][2]);      // default prediction for whatever lists are used
	uint32_t *rows = &s_rows[0][tid];
	int16_t *rs = s_rs[tid];
	// the reference-row ring of the warp's lanes (RowFetch) lies over the warp's residual scratch: interpolation and the
	// residual phases that use the scratch are separated by __syncwarp()
	static_assert(sizeof(uint4) * 64 * I5_DEPTH <= sizeof(int16_t) * 66 * 32, "ring fits the warp's scratch");
	uint4 *ring = reinterpret_cast<uint4 *>(s_rs[tid & ~31]) + lane;
	uint16_t *queue = s_queue[warp];
	const unsigned lt_mask = (1u << lane) - 1;
	const uint8_t *refl = job.frames + job.frame_bytes * mc.ref_frame[lastl];
	const int mvx = mc.mv[lastl][0], mvy = mc.mv[lastl][1];

	// ---- first list of a bi-predicted strip: a launch of its own, before the one that finishes the strips ----
	if (MODE != 2) {
		if (bi && h) {
			const uint8_t *ref0 = job.frames + job.frame_bytes * mc.ref_frame[0];
			if (wide) i5_first_list<2>(ref0, dst_y, dst_c, W, H, px0, py0, cyp, mc.mv[0][0], mc.mv[0][1], h, ring);
			else i5_first_list<1>(ref0, dst_y, dst_c, W, H, px0, py0, cyp, mc.mv[0][0], mc.mv[0][1], h, ring);
		}
		if (first_pass) return;
	}
	// ---- luma: prediction of the (last) list, parked -------------------------------------------------------
	if (h) {
		if (wide) i5_park_luma<2>(refl, W, H, px0, py0, mvx, mvy, h, rows, ring);
		else i5_park_luma<1>(refl, W, H, px0, py0, mvx, mvy, h, rows, ring);
	}
	if (bi || !plain_w) {
		for (int r = 0; r < h; ++r)
			for (int q = 0; q < nq; ++q) {
				uint32_t w = rows[(r * nq + q) * I5_THREADS];
				const uint8_t *drow = dst_y + (size_t)(py0 + r) * W + px0 + 4 * q;
				if (plain_w) w = __vavgu4(*reinterpret_cast<const uint32_t *>(drow), w);
				else {
					int p[4];
#pragma unroll
					for (int i = 0; i < 4; ++i) p[i] = (w >> (8 * i)) & 255;
					if (bi) {
						const uint32_t w0 = *reinterpret_cast<const uint32_t *>(drow);
						const int lw = mc.log2wd_y, o = (mc.o[0][0] + mc.o[1][0] + 1) >> 1;
#pragma unroll
						for (int i = 0; i < 4; ++i) p[i] = clip255((((int)((w0 >> (8 * i)) & 255) * mc.w[0][0] + p[i] * mc.w[1][0] + (1 << lw)) >> (lw + 1)) + o);
					} else {
						const int lw = mc.log2wd_y, rnd = lw ? 1 << (lw - 1) : 0;
#pragma unroll
						for (int i = 0; i < 4; ++i) p[i] = clip255(((p[i] * mc.w[first][0] + rnd) >> lw) + mc.o[first][0]);
					}
					w = pack4(p[0], p[1], p[2], p[3]);
				}
				rows[(r * nq + q) * I5_THREADS] = w;
			}
	}
	// ---- luma residual, warp-wide ---------------------------------------------------------------------------
	{
		int n_jobs = 0;
#pragma unroll
		for (int s = 0; s < 8; ++s) {
			const int g = s >> 1, q = s & 1;
			bool coded = false;
			int b = 0;
			if (g < h4 && q < nq && !t8) {
				b = c_r2b[(y4 + g) * 4 + x4 + q];
				coded = (m.nz >> (1 + b)) & 1;
			}
			const unsigned bal = __ballot_sync(0xffffffffu, coded);
			if (coded) queue[n_jobs + __popc(bal & lt_mask)] = (uint16_t)(lane | (g << 5) | (q << 7) | (b << 8));
			n_jobs += __popc(bal);
		}
		__syncwarp();
		for (int j0 = 0; j0 < n_jobs; j0 += 32) {
			const bool has = j0 + lane < n_jobs;
			const int e = has ? queue[j0 + lane] : 0, o = e & 31;
			b2_h264_mb mo;           // what the transform needs to know of the owner's macroblock
			mo.type = B2_H4_INTER;
			mo.nz = __shfl_sync(0xffffffffu, m.nz, o);
			mo.coef_idx = __shfl_sync(0xffffffffu, m.coef_idx, o);
			mo.qp_y = (uint8_t)__shfl_sync(0xffffffffu, (int)m.qp_y, o);
			const int nqo = __shfl_sync(0xffffffffu, nq, o);
			if (has) {
				const int g = (e >> 5) & 3, q = (e >> 7) & 1;
				int16_t res[16];
				block_residual(job, mo, (e >> 8) & 15, res);
				uint32_t *orow = &s_rows[0][(tid & ~31) + o] + ((4 * g) * nqo + q) * I5_THREADS;
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					uint32_t *pw = orow + j * nqo * I5_THREADS;
					*pw = add_res4(*pw, res[4 * j], res[4 * j + 1], res[4 * j + 2], res[4 * j + 3]);
				}
			}
		}
		__syncwarp();
	}
	if (t8) {
		// 8x8 transform: partitions are at least 8x8 here, the strip crosses whole 8-row groups
		for (int r = 0; r < h; r += 8) {
			const int q8 = ((y4 + (r >> 2)) >> 1) * 2 + (x4 >> 1);
			if (!((m.nz >> (1 + 4 * q8)) & 1)) continue;
			block_residual8x8(job, m, q8, reinterpret_cast<int16_t(*)[16]>(rs) - 4 * q8);
			for (int rr = r; rr < r + 8 && rr < h; ++rr)
				for (int q = 0; q < nq; ++q) {
					const int16_t *res = rs + (((rr >> 2) & 1) * 2 + (x4 & 1) + q) * 16 + (rr & 3) * 4;      // narrow strips may sit in the right half of an 8x8 block
					uint32_t *pw = rows + (rr * nq + q) * I5_THREADS;
					*pw = add_res4(*pw, res[0], res[1], res[2], res[3]);
				}
		}
	}
	for (int r = 0; r < h; ++r)
		for (int q = 0; q < nq; ++q) *reinterpret_cast<uint32_t *>(dst_y + (size_t)(py0 + r) * W + px0 + 4 * q) = rows[(r * nq + q) * I5_THREADS];
	__syncwarp();       // the parked rows are reused for chroma: nobody may still be adding luma residuals

	// ---- chroma: nq words (Cb Cr Cb Cr) per row, h / 2 rows ------------------------------------------------------
	const int hc = h / 2;
	if (h) {
		if (wide) i5_park_chroma<2>(refl + luma_size, W, H >> 1, px0, cyp, mvx, mvy, hc, rows, ring);
		else i5_park_chroma<1>(refl + luma_size, W, H >> 1, px0, cyp, mvx, mvy, hc, rows, ring);
	}
	if (bi || !plain_w) {
		for (int r = 0; r < hc; ++r)
			for (int q = 0; q < nq; ++q) {
				uint32_t w = rows[(r * nq + q) * I5_THREADS];
				const uint8_t *drow = dst_c + (size_t)(cyp + r) * W + px0 + 4 * q;
				if (plain_w) w = __vavgu4(*reinterpret_cast<const uint32_t *>(drow), w);
				else {
					int p[4];
#pragma unroll
					for (int i = 0; i < 4; ++i) p[i] = (w >> (8 * i)) & 255;
					if (bi) {
						const uint32_t w0 = *reinterpret_cast<const uint32_t *>(drow);
						const int lc = mc.log2wd_c;
#pragma unroll
						for (int i = 0; i < 4; ++i) {
							const int c = 1 + (i & 1);
							p[i] = clip255((((int)((w0 >> (8 * i)) & 255) * mc.w[0][c] + p[i] * mc.w[1][c] + (1 << lc)) >> (lc + 1)) + ((mc.o[0][c] + mc.o[1][c] + 1) >> 1));
						}
					} else {
						const int lc = mc.log2wd_c, rc = lc ? 1 << (lc - 1) : 0;
#pragma unroll
						for (int i = 0; i < 4; ++i) {
							const int c = 1 + (i & 1);
							p[i] = clip255(((p[i] * mc.w[first][c] + rc) >> lc) + mc.o[first][c]);
						}
					}
					w = pack4(p[0], p[1], p[2], p[3]);
				}
				rows[(r * nq + q) * I5_THREADS] = w;
			}
	}
	// ---- chroma residual, warp-wide: one queue entry per row of chroma blocks the strip touches (both components) ----
	{
		int n_jobs = 0;
#pragma unroll
		for (int s = 0; s < 2; ++s) {
			const bool coded = hc > 0 && ((m.nz >> 17) & 0x1ff) && s * 4 < cy0 + hc && s * 4 + 4 > cy0;
			const unsigned bal = __ballot_sync(0xffffffffu, coded);
			if (coded) queue[n_jobs + __popc(bal & lt_mask)] = (uint16_t)(lane | (s << 5));
			n_jobs += __popc(bal);
		}
		__syncwarp();
		for (int j0 = 0; j0 < n_jobs; j0 += 32) {
			const bool has = j0 + lane < n_jobs;
			const int e = has ? queue[j0 + lane] : 0, o = e & 31;
			b2_h264_mb mo;
			mo.type = B2_H4_INTER;
			mo.nz = __shfl_sync(0xffffffffu, m.nz, o);
			mo.coef_idx = __shfl_sync(0xffffffffu, m.coef_idx, o);
			mo.qp_cb = (uint8_t)__shfl_sync(0xffffffffu, (int)m.qp_cb, o);
			mo.qp_cr = (uint8_t)__shfl_sync(0xffffffffu, (int)m.qp_cr, o);
			const int geo = __shfl_sync(0xffffffffu, nq | (cx0 << 2) | (cy0 << 6) | (hc << 10), o);
			if (has) {
				const int nqo = geo & 3, cxo = (geo >> 2) & 15, cyo = (geo >> 6) & 15, hco = geo >> 10, br = (e >> 5) & 1;
				const int cblk = br * 2 + (cxo >> 2);
				unsigned ok = 0;
				for (int comp = 0; comp < 2; ++comp)
					if (block_residual(job, mo, 16 + comp * 4 + cblk, rs + comp * 16)) ok |= 1u << comp;
				if (ok) {
					const int ca = max(cyo, br * 4), cb = min(cyo + hco, br * 4 + 4);
					for (int cy = ca; cy < cb; ++cy)
						for (int q = 0; q < nqo; ++q) {
							uint32_t *pw = &s_rows[0][(tid & ~31) + o] + ((cy - cyo) * nqo + q) * I5_THREADS;
							const uint32_t w = *pw;
							int p[4];
#pragma unroll
							for (int i = 0; i < 4; ++i) {
								const int comp = i & 1, cx = cxo + 2 * q + (i >> 1);
								p[i] = (w >> (8 * i)) & 255;
								if ((ok >> comp) & 1) p[i] = clip255(p[i] + rs[comp * 16 + (cy & 3) * 4 + (cx & 3)]);
							}
							*pw = pack4(p[0], p[1], p[2], p[3]);
						}
				}
			}
		}
		__syncwarp();
	}
	for (int r = 0; r < hc; ++r)
		for (int q = 0; q < nq; ++q) *reinterpret_cast<uint32_t *>(dst_c + (size_t)(cyp + r) * W + px0 + 4 * q) = rows[(r * nq + q) * I5_THREADS];
}

// ------------------------------------------------------------------------------------------------
// intra wavefront: one warp per macroblock of the diagonal x + 2y = d
// ------------------------------------------------------------------------------------------------
struct IntraTile {
	uint8_t y[17][28];      // row 0 = samples above (col 0 = top-left, cols 1..16 above, 17..24 above-right); col 0 = left
	uint8_t f8t[20], f8l[12];   // Intra8x8: filtered reference samples p'[-1..15, -1] and p'[-1, -1..7]
	uint8_t c[2][9][12];    // per component: row 0 above (col 0 top-left), col 0 left
	int16_t res[24][16];
	uint32_t has;
};

__device__ __forceinline__ int blk_avail(int k, int a) {
	int x = c_blk_x[k], y = c_blk_y[k];
	int left = x > 0 || (a & 1), top = y > 0 || (a & 2), tr;
	if (y == 0) tr = x < 3 ? (a & 2) != 0 : (a & 4) != 0;
	else if (x == 3) tr = 0;
	else tr = c_r2b[(y - 1) * 4 + x + 1] < k;
	return (left ? 1 : 0) | (top ? 2 : 0) | (tr ? 4 : 0);
}

// one sample of Intra4x4 prediction; T(i) = p[i,-1] (i = -1..7), L(i) = p[-1,i] (i = -1..3)
__device__ int pred4x4_sample(const IntraTile &tl, int bx, int by, int mode, int avail, int x, int y) {
	const uint8_t *top = &tl.y[by * 4][bx * 4 + 1];      // top[i] = p[i,-1], top[-1] = p[-1,-1]
#define T(i) ((avail & 4) || (i) < 4 ? (int)top[(i)] : (int)top[3])
#define L(i) ((int)tl.y[by * 4 + 1 + (i)][bx * 4])
	switch (mode) {
	case 0: return T(x);
	case 1: return L(y);
	case 2:
		if ((avail & 3) == 3) return (T(0) + T(1) + T(2) + T(3) + L(0) + L(1) + L(2) + L(3) + 4) >> 3;
		if (avail & 1) return (L(0) + L(1) + L(2) + L(3) + 2) >> 2;
		if (avail & 2) return (T(0) + T(1) + T(2) + T(3) + 2) >> 2;
		return 128;
	case 3:
		return (x == 3 && y == 3) ? (T(6) + 3 * T(7) + 2) >> 2 : (T(x + y) + 2 * T(x + y + 1) + T(x + y + 2) + 2) >> 2;
	case 4:
		if (x > y) return (T(x - y - 2) + 2 * T(x - y - 1) + T(x - y) + 2) >> 2;
		if (x < y) return (L(y - x - 2) + 2 * L(y - x - 1) + L(y - x) + 2) >> 2;
		return (T(0) + 2 * T(-1) + L(0) + 2) >> 2;
	case 5: {
		int z = 2 * x - y;
		if (z >= 0 && !(z & 1)) return (T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 1) >> 1;
		if (z >= 0) return (T(x - (y >> 1) - 2) + 2 * T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 2) >> 2;
		if (z == -1) return (L(0) + 2 * L(-1) + T(0) + 2) >> 2;
		return (L(y - 1) + 2 * L(y - 2) + L(y - 3) + 2) >> 2;
	}
	case 6: {
		int z = 2 * y - x;
		if (z >= 0 && !(z & 1)) return (L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 1) >> 1;
		if (z >= 0) return (L(y - (x >> 1) - 2) + 2 * L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 2) >> 2;
		if (z == -1) return (L(0) + 2 * L(-1) + T(0) + 2) >> 2;
		return (T(x - 1) + 2 * T(x - 2) + T(x - 3) + 2) >> 2;
	}
	case 7:
		if (!(y & 1)) return (T(x + (y >> 1)) + T(x + (y >> 1) + 1) + 1) >> 1;
		return (T(x + (y >> 1)) + 2 * T(x + (y >> 1) + 1) + T(x + (y >> 1) + 2) + 2) >> 2;
	default: {
		int z = x + 2 * y;
		if (z > 5) return L(3);
		if (z == 5) return (L(2) + 3 * L(3) + 2) >> 2;
		if (!(z & 1)) return (L(y + (x >> 1)) + L(y + (x >> 1) + 1) + 1) >> 1;
		return (L(y + (x >> 1)) + 2 * L(y + (x >> 1) + 1) + L(y + (x >> 1) + 2) + 2) >> 2;
	}
	}
#undef T
#undef L
}

// availability of 8x8 block q from the MB-level mask (luma_intra8x8_with_residual, h264.cpp:4093-4118)
__device__ __forceinline__ int blk8_avail(int q, int a) {
	return q == 0 ? (a & ~4) | ((a & 2) * 2) : (q == 1 ? (a & ~8) | ((a & 2) * 4) | 1 : (q == 2 ? 6 | ((a & 1) * 9) : 11));
}

// one sample of Intra8x8 prediction from the filtered reference samples (8.3.2.2.2 ff.)
__device__ int pred8x8_sample(const IntraTile &tl, int mode, int avail, int x, int y) {
#define PT(i) ((int)tl.f8t[1 + (i)])
#define PL(i) ((int)tl.f8l[1 + (i)])
	switch (mode) {
	case 0: return PT(x);
	case 1: return PL(y);
	case 2: {
		int st = 0, sl = 0;
		for (int i = 0; i < 8; ++i) { st += PT(i); sl += PL(i); }
		return (avail & 3) == 3 ? (st + sl + 8) >> 4 : ((avail & 1) ? (sl + 4) >> 3 : ((avail & 2) ? (st + 4) >> 3 : 128));
	}
	case 3: return (x == 7 && y == 7) ? (PT(14) + 3 * PT(15) + 2) >> 2 : (PT(x + y) + 2 * PT(x + y + 1) + PT(x + y + 2) + 2) >> 2;
	case 4:
		if (x > y) return (PT(x - y - 2) + 2 * PT(x - y - 1) + PT(x - y) + 2) >> 2;
		if (x < y) return (PL(y - x - 2) + 2 * PL(y - x - 1) + PL(y - x) + 2) >> 2;
		return (PT(0) + 2 * PT(-1) + PL(0) + 2) >> 2;
	case 5: {
		int z = 2 * x - y;
		if (z >= 0 && !(z & 1)) return (PT(x - (y >> 1) - 1) + PT(x - (y >> 1)) + 1) >> 1;
		if (z >= 0) return (PT(x - (y >> 1) - 2) + 2 * PT(x - (y >> 1) - 1) + PT(x - (y >> 1)) + 2) >> 2;
		if (z == -1) return (PL(0) + 2 * PT(-1) + PT(0) + 2) >> 2;
		return (PL(y - 2 * x - 1) + 2 * PL(y - 2 * x - 2) + PL(y - 2 * x - 3) + 2) >> 2;
	}
	case 6: {
		int z = 2 * y - x;
		if (z >= 0 && !(z & 1)) return (PL(y - (x >> 1) - 1) + PL(y - (x >> 1)) + 1) >> 1;
		if (z >= 0) return (PL(y - (x >> 1) - 2) + 2 * PL(y - (x >> 1) - 1) + PL(y - (x >> 1)) + 2) >> 2;
		if (z == -1) return (PL(0) + 2 * PT(-1) + PT(0) + 2) >> 2;
		return (PT(x - 2 * y - 1) + 2 * PT(x - 2 * y - 2) + PT(x - 2 * y - 3) + 2) >> 2;
	}
	case 7:
		if (!(y & 1)) return (PT(x + (y >> 1)) + PT(x + (y >> 1) + 1) + 1) >> 1;
		return (PT(x + (y >> 1)) + 2 * PT(x + (y >> 1) + 1) + PT(x + (y >> 1) + 2) + 2) >> 2;
	default: {
		int z = x + 2 * y;
		if (z > 13) return PL(7);
		if (z == 13) return (PL(6) + 3 * PL(7) + 2) >> 2;
		if (!(z & 1)) return (PL(y + (x >> 1)) + PL(y + (x >> 1) + 1) + 1) >> 1;
		return (PL(y + (x >> 1)) + 2 * PL(y + (x >> 1) + 1) + PL(y + (x >> 1) + 2) + 2) >> 2;
	}
	}
#undef PT
#undef PL
}

// The residual of an intra macroblock (every coded 4x4 / 8x8 block dequantised and transformed into tl.res, one
// block per lane).  It needs nothing from the neighbours, so the kernels run it BEFORE they wait for the
// macroblocks this one predicts from: the coefficient loads and the transforms are off the dependency chain
// between neighbouring intra macroblocks.
__device__ void intra_residual(const H4Job &job, const b2_h264_mb &m, IntraTile &tl, int lane) {
	if (m.type == B2_H4_IPCM) return;
	if (lane == 0) tl.has = 0;
	__syncwarp();
	if (m.flags & B2_H4_T8x8) {
		if (lane < 4) {
			if ((m.nz >> (1 + 4 * lane)) & 1) {
				block_residual8x8(job, m, lane, tl.res);
				atomicOr(&tl.has, 0xfu << (4 * lane));
			}
		} else if (lane >= 16 && lane < 24) {
			if (block_residual(job, m, lane, tl.res[lane])) atomicOr(&tl.has, 1u << lane);
		}
	} else if (lane < 24) {
		if (block_residual(job, m, lane, tl.res[lane])) atomicOr(&tl.has, 1u << lane);
	}
}

// Reconstruct one intra macroblock with one warp (intra_residual has run).  Pixels produced by other warps
// (rows above) are read through L2 (__ldcg) and results are written through to L2 (__stcg): the wavefront
// kernel orders warps with progress flags, and L1 is not coherent between SMs.
__device__ void intra_mb(const H4Job &job, const b2_h264_mb &m, int mbx, int mby, IntraTile &tl, int lane) {
	const int W = job.mb_w * 16, H = job.mb_h * 16;
	uint8_t *fy = job.frames + job.frame_bytes * job.dst;
	uint8_t *fc = fy + (size_t)W * H;
	uint8_t *py = fy + (size_t)mby * 16 * W + mbx * 16;
	uint8_t *pc = fc + (size_t)mby * 8 * W + mbx * 16;

	if (m.type == B2_H4_IPCM) {
		const uint8_t *src = reinterpret_cast<const uint8_t *>(job.coefs + (size_t)m.coef_idx * 16);
		for (int i = lane; i < 64; i += 32) {          // 64 luma words
			int row = i >> 2, xq = i & 3;
			__stcg(reinterpret_cast<uint32_t *>(py + (size_t)row * W + xq * 4), *reinterpret_cast<const uint32_t *>(src + row * 16 + xq * 4));
		}
		for (int i = lane; i < 128; i += 32) {         // 128 chroma bytes
			int row = i >> 4, xb = i & 15;
			__stcg(pc + (size_t)row * W + xb, src[256 + (xb & 1) * 64 + row * 8 + (xb >> 1)]);
		}
		return;
	}

	// ---- neighbours into the tile ------------------------------------------------------------
	const int a = m.intra_avail | m.intra_avail_c;
	if (lane < 25) {            // above row: x = -1 .. 23
		int x = lane - 1;
		// each part of the row by its own availability bit, as the reference's predictors see them
		// (the top-left bit survives constrained-intra masking of the MB above, h264.cpp:3271-3273)
		bool ok = x < 0 ? (a & 8) && mbx > 0 : (x < 16 ? (a & 2) != 0 : (a & 4) && mbx < job.mb_w - 1);
		tl.y[0][lane] = (ok && mby > 0) ? __ldcg(py - (ptrdiff_t)W + x) : 128;
	}
	if (lane < 16) tl.y[1 + lane][0] = ((a & 1) && mbx > 0) ? __ldcg(py + (size_t)lane * W - 1) : 128;
	if (lane < 18) {            // chroma above: byte x = -2 .. 15
		int xb = lane - 2;
		bool ok = (a & 2) && mby > 0 && !(xb < 0 && mbx == 0);
		tl.c[xb & 1][0][(xb + 2) >> 1] = ok ? __ldcg(pc - (ptrdiff_t)W + xb) : 128;
	}
	if (lane < 16) {            // chroma left: rows 0..7, two components
		int row = lane >> 1, comp = lane & 1;
		tl.c[comp][1 + row][0] = ((a & 1) && mbx > 0) ? __ldcg(pc + (size_t)row * W - 2 + comp) : 128;
	}
	__syncwarp();

	// ---- luma -----------------------------------------------------------------------------------
	if (m.type == B2_H4_I4x4) {
		// Ten steps instead of sixteen: blocks 2 | 4, 3 | 5, 6 | 8, 7 | 9, 10 | 12, 11 | 13 of the decoding order depend on
		// nothing of each other (a block's top-right neighbour counts only if it comes earlier in that order), so the two
		// halves of the warp take one each.
#pragma unroll 1
		for (int s = 0; s < 10; ++s) {
			const int k = c_i4_pair[lane >> 4][s];
			if (k >= 0) {
				const int bx = c_blk_x[k], by = c_blk_y[k];
				int x = lane & 3, y = (lane >> 2) & 3;
				int mode = (m.i4_modes[k >> 1] >> ((k & 1) * 4)) & 15;
				int v = pred4x4_sample(tl, bx, by, mode, blk_avail(k, m.intra_avail), x, y);
				if ((tl.has >> k) & 1) v = clip255(v + tl.res[k][y * 4 + x]);
				tl.y[by * 4 + 1 + y][bx * 4 + 1 + x] = (uint8_t)v;
			}
			__syncwarp();
		}
	} else if (m.type == B2_H4_I8x8) {
		for (int q = 0; q < 4; ++q) {
			const int bx = (q & 1) * 8, by = (q >> 1) * 8;
			const int av = blk8_avail(q, m.intra_avail);
			const int mode = (m.i4_modes[q >> 1] >> ((q & 1) * 4)) & 15;
			// reference sample filtering (8.3.2.2.1): lanes 0..15 the row above, 16 the corner, 17..24 the left column
			{
				const bool L = av & 1, T = av & 2, TR = av & 4, TL = av & 8;
				const uint8_t *top = &tl.y[by][bx + 1];          // top[i] = p[i, -1], top[-1] = p[-1, -1]
#define RT(i) ((int)top[(TR || (i) < 8) ? (i) : 7])
#define RL(i) ((int)tl.y[by + 1 + (i)][bx])
				if (lane < 16) {
					int v = 0;
					if (T) {
						if (lane == 0) v = TL ? (RT(-1) + 2 * RT(0) + RT(1) + 2) >> 2 : (3 * RT(0) + RT(1) + 2) >> 2;
						else if (lane == 15) v = (RT(14) + 3 * RT(15) + 2) >> 2;
						else v = (RT(lane - 1) + 2 * RT(lane) + RT(lane + 1) + 2) >> 2;
					}
					tl.f8t[1 + lane] = (uint8_t)v;
				} else if (lane == 16) {
					int v = 0;
					if (TL) {
						const int c = top[-1];
						v = (T && L) ? (RT(0) + 2 * c + RL(0) + 2) >> 2 : (T ? (3 * c + RT(0) + 2) >> 2 : (L ? (3 * c + RL(0) + 2) >> 2 : c));
					}
					tl.f8t[0] = tl.f8l[0] = (uint8_t)v;
				} else if (lane < 25) {
					const int yy = lane - 17;
					int v = 0;
					if (L) {
						if (yy == 0) v = TL ? ((int)top[-1] + 2 * RL(0) + RL(1) + 2) >> 2 : (3 * RL(0) + RL(1) + 2) >> 2;
						else if (yy == 7) v = (RL(6) + 3 * RL(7) + 2) >> 2;
						else v = (RL(yy - 1) + 2 * RL(yy) + RL(yy + 1) + 2) >> 2;
					}
					tl.f8l[1 + yy] = (uint8_t)v;
				}
#undef RT
#undef RL
			}
			__syncwarp();
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				const int x = lane & 7, y = (lane >> 3) + 4 * h;
				int v = pred8x8_sample(tl, mode, av, x, y);
				const int k = 4 * q + (y >> 2) * 2 + (x >> 2);
				if ((tl.has >> k) & 1) v = clip255(v + tl.res[k][(y & 3) * 4 + (x & 3)]);
				tl.y[by + 1 + y][bx + 1 + x] = (uint8_t)v;
			}
			__syncwarp();
		}
	} else {
		const int av = m.intra_avail;
		int dcv = 128, pa = 0, pb = 0, pcn = 0;
		if (m.i16_mode == 2) {
			int sum = 0;
			if (av & 1) for (int i = 0; i < 16; ++i) sum += tl.y[1 + i][0];
			if (av & 2) for (int i = 0; i < 16; ++i) sum += tl.y[0][1 + i];
			dcv = (av & 3) == 3 ? (sum + 16) >> 5 : ((av & 3) ? (sum + 8) >> 4 : 128);
		} else if (m.i16_mode == 3) {
			int Hh = 0, Vv = 0;
			for (int i = 0; i < 8; ++i) {
				Hh += (i + 1) * ((int)tl.y[0][1 + 8 + i] - (int)tl.y[0][1 + 6 - i]);
				Vv += (i + 1) * ((int)tl.y[1 + 8 + i][0] - (int)tl.y[1 + 6 - i][0]);
			}
			pa = 16 * ((int)tl.y[16][0] + (int)tl.y[0][16]);
			pb = (5 * Hh + 32) >> 6;
			pcn = (5 * Vv + 32) >> 6;
		}
		__syncwarp();
		for (int i = lane; i < 256; i += 32) {
			int x = i & 15, y = i >> 4;
			int v;
			switch (m.i16_mode) {
			case 0: v = tl.y[0][1 + x]; break;
			case 1: v = tl.y[1 + y][0]; break;
			case 2: v = dcv; break;
			default: v = clip255((pa + pb * (x - 7) + pcn * (y - 7) + 16) >> 5); break;
			}
			int k = c_r2b[(y >> 2) * 4 + (x >> 2)];
			if ((tl.has >> k) & 1) v = clip255(v + tl.res[k][(y & 3) * 4 + (x & 3)]);
			// rows/columns read by the predictors above are row 0 / column 0 only: safe to overwrite
			tl.y[1 + y][1 + x] = (uint8_t)v;
		}
	}
	// ---- chroma -----------------------------------------------------------------------------------
	{
		const int av = m.intra_avail_c;
		for (int i = lane; i < 128; i += 32) {
			int comp = i >> 6, x = i & 7, y = (i >> 3) & 7;
			const uint8_t (*c)[12] = tl.c[comp];
			int v;
			if (m.chroma_mode == 0) {
				int bx = x >> 2, by = y >> 2;
				int st = 0, sl = 0;
				for (int j = 0; j < 4; ++j) { st += c[0][1 + bx * 4 + j]; sl += c[1 + by * 4 + j][0]; }
				bool ht = av & 2, hl = av & 1;
				if (bx == by) v = (ht && hl) ? (st + sl + 4) >> 3 : (hl ? (sl + 2) >> 2 : (ht ? (st + 2) >> 2 : 128));
				else if (bx == 1) v = ht ? (st + 2) >> 2 : (hl ? (sl + 2) >> 2 : 128);
				else v = hl ? (sl + 2) >> 2 : (ht ? (st + 2) >> 2 : 128);
			} else if (m.chroma_mode == 1) {
				v = c[1 + y][0];
			} else if (m.chroma_mode == 2) {
				v = c[0][1 + x];
			} else {
				int Hh = 0, Vv = 0;
				for (int j = 0; j < 4; ++j) {
					Hh += (j + 1) * ((int)c[0][1 + 4 + j] - (int)c[0][1 + 2 - j]);
					Vv += (j + 1) * ((int)c[1 + 4 + j][0] - (int)c[1 + 2 - j][0]);
				}
				int pa = 16 * ((int)c[8][0] + (int)c[0][8]), pb = (34 * Hh + 32) >> 6, pcn = (34 * Vv + 32) >> 6;
				v = clip255((pa + pb * (x - 3) + pcn * (y - 3) + 16) >> 5);
			}
			int blk = 16 + comp * 4 + (y >> 2) * 2 + (x >> 2);
			if ((tl.has >> blk) & 1) v = clip255(v + tl.res[blk][(y & 3) * 4 + (x & 3)]);
			__stcg(pc + (size_t)y * W + 2 * x + comp, (uint8_t)v);
		}
	}
	__syncwarp();
	for (int i = lane; i < 64; i += 32) {
		int row = i >> 2, xq = i & 3;
		const uint8_t *s = &tl.y[1 + row][1 + xq * 4];
		uint32_t w = (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24);
		__stcg(reinterpret_cast<uint32_t *>(py + (size_t)row * W + xq * 4), w);
	}
}

// Deblock one macroblock with one warp, in two register passes.
//   V pass (vertical edges): lane l < 16 holds luma row l (columns -4..15) unpacked in registers,
//   lanes 16..23 hold the eight chroma rows (byte columns -4..15); the four edges of the row are
//   filtered in registers, then the row is parked in the shared-memory tile.  Lanes 24..29 bring in
//   the rows above the macroblock.
//   H pass (horizontal edges): lane c < 16 holds luma column c (rows -4..15), lanes 16..31 hold the
//   chroma byte columns (rows -2..7); same filtering down the column, written back to the tile.
//   The tile is then stored with 32-bit L2 stores.  Pixels of other warps are read through L2.
struct DbkTile {
	uint8_t y[20][36];      // rows -4..15 at index r+4, columns -4..15 at index c+4; 36-byte pitch: conflict-free columns
	uint8_t c[10][36];      // chroma rows -2..7 at index r+2, byte columns -4..15 at index c+4
};

// ------------------------------------------------------------------------------------------------
// deblocking, second form: one warp walks a macroblock row and keeps a sliding tile.
//   * all 32 lanes run one code path: lanes 0..15 hold a luma line of 20 samples, lanes 16..31 a
//     chroma line of ONE component (10 samples parked at v[2..11], so that the chroma edges sit at
//     the register positions of luma edges 0 and 1);
//   * the last four columns of a macroblock stay in shared memory for the next macroblock's left
//     edge instead of going through L2 (they are stored once, when final);
//   * the untouched pixels and the strength record of the next macroblock that needs work are
//     fetched while the current one is filtered.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire(const int *p) {
	int v;
	asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
	asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Spinning on a flag: every ld.acquire.gpu carries an L1 invalidation (CCTL.IVALL; 10 % of the stall samples of
// the round-1 deblocking kernel), so the loop polls with relaxed loads and one acquire fence follows the hit.
#ifndef B2_SPIN_RELAXED
#define B2_SPIN_RELAXED 1
#endif
__device__ __forceinline__ int ld_relaxed(const int *p) {
	int v;
	asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void fence_acquire() {
#if B2_SPIN_RELAXED != 2      // (2: timing experiment without the fence)
	asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}
#ifndef B2_GL_SPIN_NS
#define B2_GL_SPIN_NS 32
#endif
#ifndef B2_SM_SPIN_NS
#define B2_SM_SPIN_NS 64
#endif
// wait until *p >= want; returns the value seen
__device__ __forceinline__ int spin_until(const int *p, int want) {
	int v;
#if B2_SPIN_RELAXED
	while ((v = ld_relaxed(p)) < want) __nanosleep(B2_GL_SPIN_NS);
	fence_acquire();
#else
	while ((v = ld_acquire(p)) < want) __nanosleep(32);
#endif
	return v;
}

// Progress of a macroblock row ("macroblocks < n are final") as seen by the row below.  Rows of
// one CTA's group hand it over through shared memory with CTA-scope fences; only the first row of
// a group reads, and only the last one writes, the GPU-scope flag in global memory.
struct RowSync {
	volatile int *sm_above;   // row above is in this CTA (else null)
	const int *gl_above;      // row above belongs to another CTA (else null)
	volatile int *sm_my;
	int *gl_my;               // null: no other CTA reads this row's progress
	int seen;                 // progress of the row above known so far
	// Experiment knob (B2_DBK2_LEAD, default 0 = off): a row that has to wait resumes only once the row above is
	// `lead` macroblocks further (or complete, `limit` = macroblocks per row), so that rows running at the same speed
	// stay apart and stop polling right behind their producer.  Measured on B200 (256 lanes): every lead costs
	// throughput (packed deblocking 660 GB/s at 0, 577 at 4, 427 at 8, 296 at 16): a launch is one wave of all its
	// pictures, ~18 resident rows per picture pair, and each macroblock of extra distance is paid 68 times per
	// picture as rows that have not started yet.
	int lead, limit;

	__device__ __forceinline__ void wait(int want, int lane) {
#ifdef B2_DBK_NOWAIT      // timing experiment only (wrong pixels): what the kernels cost without the dependencies
		return;
#endif
		if (seen >= want) return;
		const int target = min(want + lead, limit);
		if (sm_above) {
			if (lane == 0) {
				if (lead) { while ((seen = *sm_above) < target) __nanosleep(B2_SM_SPIN_NS); }
				else { while ((seen = *sm_above) < target) { } }
				__threadfence_block();
			}
		} else if (gl_above) {
			if (lane == 0) seen = spin_until(gl_above, target);
		} else {
			seen = 1 << 30;
		}
		seen = __shfl_sync(0xffffffffu, seen, 0);
	}
	// all lanes' stores so far become visible before the new value (the caller has them behind a __syncwarp)
	__device__ __forceinline__ void publish(int v, int lane) {
		__syncwarp();
		if (lane == 0) {
			if (gl_my) st_release(gl_my, v);
			else __threadfence_block();
			*sm_my = v;
		}
	}
};

struct alignas(4) DbkTables { uint8_t alpha[52], beta[52], tc0[52][4]; };

// One edge of one line, without data-dependent branches for the common case (bS < 4): the filtered
// values are computed for every lane and selected by the line's own conditions, so lanes that differ
// in strength, thresholds or sample activity stay converged.  `on` = this lane filters this edge at all.
template <bool MAY_BS4>
__device__ __forceinline__ void dbk_filter(int *v, const int q, const bool on, const int bs, const int alpha, const int beta, const int tc0, const bool luma) {
	const int p0 = v[q - 1], p1 = v[q - 2], p2 = v[q - 3], q0 = v[q], q1 = v[q + 1], q2 = v[q + 2];
	// |d| < t as one unsigned comparison, (unsigned)(d + t - 1) < 2t - 1 (t >= 1; a zero threshold switches the edge off)
	const unsigned am = (unsigned)(alpha - 1), a2 = (unsigned)(2 * alpha - 1), bm = (unsigned)(beta - 1), b2 = (unsigned)(2 * beta - 1);
	const bool act = on && beta != 0 && (unsigned)(p0 - q0) + am < a2 && (unsigned)(p1 - p0) + bm < b2 && (unsigned)(q1 - q0) + bm < b2;
	const bool ap = luma && (unsigned)(p2 - p0) + bm < b2, aq = luma && (unsigned)(q2 - q0) + bm < b2;
	const bool normal = act && (!MAY_BS4 || bs < 4);
	{
		const int tc = luma ? tc0 + ap + aq : tc0 + 1, avg = (p0 + q0 + 1) >> 1;
		const int np1 = p1 + min(max((p2 + avg - 2 * p1) >> 1, -tc0), tc0);
		const int nq1 = q1 + min(max((q2 + avg - 2 * q1) >> 1, -tc0), tc0);
		const int delta = min(max(((q0 - p0) * 4 + (p1 - q1) + 4) >> 3, -tc), tc);
		if (normal && ap) v[q - 2] = np1;
		if (normal && aq) v[q + 1] = nq1;
		if (normal) { v[q - 1] = clip255(p0 + delta); v[q] = clip255(q0 - delta); }
	}
	// bS 4 exists on macroblock edges only (edge 0): the inner edges are straight-line code
	if (MAY_BS4 && __any_sync(__activemask(), act && bs == 4)) {
		if (act && bs == 4) {
			const bool strong = luma && (unsigned)(p0 - q0 + (alpha >> 2) + 1) < (unsigned)(2 * ((alpha >> 2) + 2) - 1);
			if (strong && ap) {
				const int p3 = v[q - 4];
				v[q - 1] = (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3;
				v[q - 2] = (p2 + p1 + p0 + q0 + 2) >> 2;
				v[q - 3] = (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3;
			} else {
				v[q - 1] = (2 * p1 + p0 + q1 + 2) >> 2;
			}
			if (strong && aq) {
				const int q3 = v[q + 3];
				v[q] = (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3;
				v[q + 1] = (p0 + q0 + q1 + q2 + 2) >> 2;
				v[q + 2] = (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3;
			} else {
				v[q] = (2 * q1 + q0 + p1 + 2) >> 2;
			}
		}
	}
}

// the (up to four) edges of one line held in v[]; bits = the 32-bit strength word of the
// direction, seg = 4-sample segment of this line, qpc / qpn = QP of this MB / of the MB across edge 0
__device__ __forceinline__ void dbk_line(int *v, const DbkTables &tb, uint32_t bits, int seg, bool luma, bool edge0_ok, bool edge0_bs4,
                                         int qpc, int qpn, int aoff, int boff) {
	// thresholds of the inner edges (this macroblock's QP) and of edge 0 (mean with the neighbour's)
	const int iai = min(max(qpc + aoff, 0), 51), ibi = min(max(qpc + boff, 0), 51);
	const int qp0 = (qpc + qpn + 1) >> 1;
	const int ia0 = min(max(qp0 + aoff, 0), 51), ib0 = min(max(qp0 + boff, 0), 51);
	const int alpha_i = tb.alpha[iai], beta_i = tb.beta[ibi], alpha_0 = tb.alpha[ia0], beta_0 = tb.beta[ib0];
	const uint32_t tc_i = *reinterpret_cast<const uint32_t *>(tb.tc0[iai]), tc_0 = *reinterpret_cast<const uint32_t *>(tb.tc0[ia0]);
	const uint32_t lbits = bits >> (2 * seg);       // this line's strength of edge e in bits 8e, 8e+1
	{
		int bs = lbits & 3;
		const bool on = bs != 0 && edge0_ok;
		if (edge0_bs4) bs = 4;
		if (__any_sync(__activemask(), on))
			dbk_filter<true>(v, 4, on && alpha_0 != 0, bs, alpha_0, beta_0, (tc_0 >> (8 * (bs < 4 ? (bs ? bs - 1 : 0) : 3))) & 255, luma);
	}
	if (!(bits & 0xffffff00u)) return;               // no inner edge of this macroblock has a strength
#pragma unroll
	for (int i = 1; i < 4; ++i) {
		const int e = luma ? i : 2 * i;                 // chroma edges 0 / 1 of the register layout are luma edges 0 / 2
		const int bs = e < 4 ? (lbits >> (8 * e)) & 3 : 0;
		const bool on = bs != 0 && (luma || i < 2);
		dbk_filter<false>(v, 4 + 4 * i, on && alpha_i != 0, bs, alpha_i, beta_i, (tc_i >> (8 * (bs ? bs - 1 : 0))) & 255, luma);
	}
}

// One row chunk: macroblocks x0 + k for the set bits k of `todo`, left to right.
__device__ void deblock_chunk(const H4Job &job, const DbkTables &tb, DbkTile &tl, int mby, int x0, unsigned todo, int lane,
                              RowSync &sy, int &done_to) {
	const int mb_w = job.mb_w, W = mb_w * 16, H = job.mb_h * 16;
	uint8_t *fy = job.frames + job.frame_bytes * job.dst;
	uint8_t *fc = fy + (size_t)W * H;
	const bool luma = lane < 16;
	const int comp = lane & 1;                              // chroma lanes: component
	const int line = luma ? lane : (lane - 16) >> 1;        // V pass: picture line inside the MB (chroma: 0..7)
	// V-pass source line of this lane inside the picture (bytes), without the MB x offset
	const uint8_t *line_base = luma ? fy + (size_t)(mby * 16 + line) * W : fc + (size_t)(mby * 8 + line) * W;
	uint32_t nw[4] = {0, 0, 0, 0};       // prefetched own words of the next macroblock
	uint4 nd = make_uint4(0, 0, 0, 0);   // its strength record
	uint32_t nql = 0, nqt = 0;           // QP words of its left / top neighbour records
	bool have_next = false, carry_in = false;
	while (todo) {
		const int k = __ffs(todo) - 1;
		todo &= todo - 1;
		const int x = x0 + k;
		const int kn = todo ? __ffs(todo) - 1 : -1;
		const bool carry_out = kn == k + 1;               // the next macroblock takes over columns 12..15
		if (x > done_to) sy.publish(x, lane);              // the stretch that needed nothing
		const b2_h264_dbk *dp = job.dbk + (size_t)mby * mb_w + x;
		uint32_t w[5];
		uint4 dr;
		uint32_t ql, qt;
		if (have_next) {
#pragma unroll
			for (int i = 0; i < 4; ++i) w[1 + i] = nw[i];
			dr = nd; ql = nql; qt = nqt;
		} else {
#pragma unroll
			for (int i = 0; i < 4; ++i) w[1 + i] = __ldcg(reinterpret_cast<const uint32_t *>(line_base + x * 16 + 4 * i));
			dr = __ldg(reinterpret_cast<const uint4 *>(dp));
			ql = x > 0 ? __ldg(reinterpret_cast<const uint32_t *>(dp - 1) + 2) : 0;
			qt = mby > 0 ? __ldg(reinterpret_cast<const uint32_t *>(dp - mb_w) + 2) : 0;
		}
		if (mby > 0) sy.wait(min(x + 2, mb_w), lane);
		// rows above the macroblock (final once the row above is two macroblocks ahead): lanes 0..15 one
		// word of luma rows -4..-1, lanes 16..23 one word of chroma rows -2..-1
		uint32_t tw = 0;
		if (mby > 0 && lane < 24) {
			const uint8_t *src = luma ? fy + (size_t)(mby * 16 - 4 + (lane >> 2)) * W : fc + (size_t)(mby * 8 - 2 + ((lane - 16) >> 2)) * W;
			tw = __ldcg(reinterpret_cast<const uint32_t *>(src + x * 16 + 4 * (lane & 3)));
		}
		// left four columns: from the tile when the previous macroblock left them there
		if (carry_in) w[0] = *reinterpret_cast<const uint32_t *>(luma ? &tl.y[4 + line][0] : &tl.c[2 + line][0]);
		else w[0] = x > 0 ? __ldcg(reinterpret_cast<const uint32_t *>(line_base + x * 16 - 4)) : 0;
		// prefetch for the next macroblock of this chunk
		have_next = kn >= 0;
		if (have_next) {
			const int xn = x0 + kn;
			const b2_h264_dbk *dn = job.dbk + (size_t)mby * mb_w + xn;
#pragma unroll
			for (int i = 0; i < 4; ++i) nw[i] = __ldcg(reinterpret_cast<const uint32_t *>(line_base + xn * 16 + 4 * i));
			nd = __ldg(reinterpret_cast<const uint4 *>(dn));
			nql = __ldg(reinterpret_cast<const uint32_t *>(dn - 1) + 2);      // xn > 0 always
			nqt = mby > 0 ? __ldg(reinterpret_cast<const uint32_t *>(dn - mb_w) + 2) : 0;
		}
		// ---- record fields ---------------------------------------------------------------------------
		const uint32_t bs_v = dr.x, bs_h = dr.y;
		const int flags = (dr.z >> 24) & 255;
		const int aoff = (int8_t)(dr.w & 255), boff = (int8_t)((dr.w >> 8) & 255);
		const int qsel = luma ? 0 : 8 + 8 * comp;            // byte of the QP word: qp_y, qp_cb, qp_cr
		const int qpc = (int8_t)((dr.z >> qsel) & 255), qpl = (int8_t)((ql >> qsel) & 255), qpt = (int8_t)((qt >> qsel) & 255);
		const bool left_ok = flags & B2_DBK_LEFT_EDGE, top_ok = flags & B2_DBK_TOP_EDGE;
		// ---- V pass ----------------------------------------------------------------------------------
		{
			int v[20];
#pragma unroll
			for (int i = 0; i < 20; ++i) v[i] = 0;
			if (luma) {
#pragma unroll
				for (int i = 0; i < 20; ++i) v[i] = (w[i >> 2] >> ((i & 3) * 8)) & 255;
			} else {
#pragma unroll
				for (int i = 0; i < 10; ++i) v[2 + i] = (w[i >> 1] >> (((i & 1) * 2 + comp) * 8)) & 255;
			}
			if (bs_v) dbk_line(v, tb, bs_v, luma ? line >> 2 : line >> 1, luma, left_ok, flags & B2_DBK_LEFT_BS4, qpc, qpl, aoff, boff);
			if (luma) {
				uint32_t *dstw = reinterpret_cast<uint32_t *>(&tl.y[4 + line][0]);
#pragma unroll
				for (int j = 0; j < 5; ++j)
					dstw[j] = (uint32_t)v[4 * j] | ((uint32_t)v[4 * j + 1] << 8) | ((uint32_t)v[4 * j + 2] << 16) | ((uint32_t)v[4 * j + 3] << 24);
			} else {
				uint8_t *dstb = &tl.c[2 + line][comp];
#pragma unroll
				for (int i = 0; i < 10; ++i) dstb[2 * i] = (uint8_t)v[2 + i];
			}
			if (lane < 24) {
				uint32_t *dstw = reinterpret_cast<uint32_t *>(luma ? &tl.y[lane >> 2][4] : &tl.c[(lane - 16) >> 2][4]);
				dstw[lane & 3] = tw;
			}
		}
		__syncwarp();
		// ---- H pass: lanes 0..15 luma column, lanes 16..31 chroma byte column ---------------------------
		if (bs_h) {
			const int col = lane & 15;
			int v[20];
#pragma unroll
			for (int i = 0; i < 20; ++i) v[i] = 0;
			if (luma) {
#pragma unroll
				for (int i = 0; i < 20; ++i) v[i] = tl.y[i][4 + col];
			} else {
#pragma unroll
				for (int i = 0; i < 10; ++i) v[2 + i] = tl.c[i][4 + col];
			}
			dbk_line(v, tb, bs_h, col >> 2, luma, top_ok, flags & B2_DBK_TOP_BS4, qpc, qpt, aoff, boff);
			if (luma) {
#pragma unroll
				for (int i = 1; i < 20; ++i) tl.y[i][4 + col] = (uint8_t)v[i];
			} else {
#pragma unroll
				for (int i = 1; i < 10; ++i) tl.c[i][4 + col] = (uint8_t)v[2 + i];
			}
			__syncwarp();
		}
		// ---- store what is final: columns -4..11 (or ..15 when nobody takes them over) of the lines, and
		//      the rows above (luma -3..-1 by lanes 24..26, chroma -1 by lane 27) ---------------------------
		if (lane < 24) {
			const int li = lane < 16 ? lane : lane - 16;
			const uint32_t *srcw = reinterpret_cast<const uint32_t *>(lane < 16 ? &tl.y[4 + li][0] : &tl.c[2 + li][0]);
			uint8_t *dst = (lane < 16 ? fy + (size_t)(mby * 16 + li) * W : fc + (size_t)(mby * 8 + li) * W) + x * 16;
			if (x > 0) __stcg(reinterpret_cast<uint32_t *>(dst - 4), srcw[0]);
#pragma unroll
			for (int j = 1; j < 4; ++j) __stcg(reinterpret_cast<uint32_t *>(dst + 4 * (j - 1)), srcw[j]);
			if (!carry_out) __stcg(reinterpret_cast<uint32_t *>(dst + 12), srcw[4]);
		} else if (lane < 28 && top_ok) {
			const bool tl_luma = lane < 27;
			const int r = tl_luma ? lane - 27 : -1;                 // luma rows -3..-1, chroma row -1
			const uint32_t *srcw = reinterpret_cast<const uint32_t *>(tl_luma ? &tl.y[4 + r][4] : &tl.c[2 + r][4]);
			uint8_t *dst = (tl_luma ? fy + (size_t)(mby * 16 + r) * W : fc + (size_t)(mby * 8 + r) * W) + x * 16;
#pragma unroll
			for (int j = 0; j < 4; ++j) __stcg(reinterpret_cast<uint32_t *>(dst + 4 * j), srcw[j]);
		}
		__syncwarp();
		if (carry_out && lane < 24) {
			uint32_t *rw = reinterpret_cast<uint32_t *>(lane < 16 ? &tl.y[4 + lane][0] : &tl.c[2 + lane - 16][0]);
			rw[0] = rw[4];
		}
		carry_in = carry_out;
		// progress: everything up to x is final, except columns 12..15 of x itself when they were carried
		// over -- which the row below does not touch before it has seen x + 2
		done_to = x + 1;
		sy.publish(x + 1, lane);
		__syncwarp();
	}
}

// ------------------------------------------------------------------------------------------------
// persistent wavefront: a warp owns one MB row of one picture and walks it left to right, staying
// two macroblocks behind the row above (left / top / top-right dependencies).  Rows are handed out
// through a ticket counter in (row, job) order, so the warp a row waits for always holds an earlier
// ticket and is already running: no deadlock whatever the CTA scheduling order.
// ------------------------------------------------------------------------------------------------

#ifndef B2_WAVE_WARPS
#define B2_WAVE_WARPS 4
#endif
constexpr int WAVE_WARPS = B2_WAVE_WARPS;
#ifndef B2_DBK_WARPS_PER_SM
#define B2_DBK_WARPS_PER_SM 24
#endif
#ifndef B2_INTRA_WARPS_PER_SM
#define B2_INTRA_WARPS_PER_SM 40
#endif
constexpr int WAVE_CTAS_DBK = B2_DBK_WARPS_PER_SM / WAVE_WARPS, WAVE_CTAS_INTRA = B2_INTRA_WARPS_PER_SM / WAVE_WARPS;    // resident CTAs per SM (register budget)

template <int PHASE>   // 1 intra, 2 deblock
__global__ void __launch_bounds__(32 * WAVE_WARPS, PHASE == 2 ? WAVE_CTAS_DBK : WAVE_CTAS_INTRA)
h264_wave_kernel(const H4Job *__restrict__ jobs, int n_jobs, int mb_w, int mb_h, int *__restrict__ sync_area) {
	__shared__ union { IntraTile intra; DbkTile dbk; } s_tile[WAVE_WARPS];
	__shared__ DbkTables s_tb;
	__shared__ int s_ticket;
	__shared__ int s_prog[WAVE_WARPS];
	if (PHASE == 2) {
		for (int i = threadIdx.x; i < 52; i += blockDim.x) {
			s_tb.alpha[i] = c_alpha[i];
			s_tb.beta[i] = c_beta[i];
			s_tb.tc0[i][0] = c_tc0[i][0]; s_tb.tc0[i][1] = c_tc0[i][1]; s_tb.tc0[i][2] = c_tc0[i][2]; s_tb.tc0[i][3] = 0;
		}
	}
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	int *ticket = sync_area;
	int *progress = sync_area + 32;           // [n_jobs][mb_h]: macroblocks of the row that are final (rows read across CTAs only)
	const int groups = (mb_h + WAVE_WARPS - 1) / WAVE_WARPS;
	const int total = n_jobs * groups;
	for (;;) {
		__syncthreads();                       // the previous group is complete: s_ticket / s_prog may be reused
		if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1);
		if (threadIdx.x < WAVE_WARPS) s_prog[threadIdx.x] = 0;
		__syncthreads();
		const int t = s_ticket;
		if (t >= total) return;
		// group-major ticket order: the group a CTA waits for (same picture, rows just above) always
		// holds an earlier ticket, so its CTA is resident or done -- no deadlock
		const int j = t % n_jobs, mby = (t / n_jobs) * WAVE_WARPS + warp;
		if (mby >= mb_h) continue;
		RowSync sy;
		sy.sm_above = warp > 0 ? &s_prog[warp - 1] : nullptr;
		sy.gl_above = (warp == 0 && mby > 0) ? progress + (size_t)j * mb_h + mby - 1 : nullptr;
		sy.sm_my = &s_prog[warp];
		sy.gl_my = (warp == WAVE_WARPS - 1 && mby + 1 < mb_h) ? progress + (size_t)j * mb_h + mby : nullptr;
		sy.seen = 0;
		sy.lead = 0;
		sy.limit = mb_w;
		const H4Job job = jobs[j];
		if (PHASE == 1 && h264_intra_sparse(job.n_intra, mb_w * mb_h)) continue;     // h264_intra_sparse_kernel's
		for (int x0 = 0; x0 < mb_w; x0 += 32) {
			// which of the next 32 macroblocks need work in this phase
			int mbx = x0 + lane;
			bool need = false;
			if (mbx < mb_w) {
				if (PHASE == 1) need = job.mbs[(size_t)mby * mb_w + mbx].type != B2_H4_INTER;
				else {
					const b2_h264_dbk *d = &job.dbk[(size_t)mby * mb_w + mbx];
					need = !(d->flags & B2_DBK_OFF) && (d->bs_v | d->bs_h);
				}
			}
			unsigned todo = __ballot_sync(0xffffffffu, need);
			int done_to = x0;                  // macroblocks < done_to are final
			if (PHASE == 2) {
				deblock_chunk(job, s_tb, s_tile[warp].dbk, mby, x0, todo, lane, sy, done_to);
				todo = 0;
			}
			while (todo) {
				int k = __ffs(todo) - 1;
				todo &= todo - 1;
				int x = x0 + k;
				if (x > done_to) sy.publish(x, lane);     // the stretch that needed nothing
				const b2_h264_mb m = job.mbs[(size_t)mby * mb_w + x];
				intra_residual(job, m, s_tile[warp].intra, lane);
				if (mby > 0) sy.wait(min(x + 2, mb_w), lane);
				intra_mb(job, m, x, mby, s_tile[warp].intra, lane);
				done_to = x + 1;
				sy.publish(done_to, lane);
			}
			int end = min(x0 + 32, mb_w);
			// a macroblock without work in this phase has nothing to wait for
			if (done_to < end) sy.publish(end, lane);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Deblocking of batched launches: helpers of the packed form.  The one-picture kernel above spends its
// instructions per sample and per macroblock; in h264_deblock3_kernel every register holds the same sample of
// TWO pictures of the launch (16-bit lanes: picture A low, B high), the filter runs on both with the packed
// 16x2 integer instructions of sm_100 (VIADD.16x2, VIMNMX.S16x2, VABSDIFF4) and plain 32-bit arithmetic on
// biased lanes, and all per-macroblock work (waiting, publishing, addressing, the transposes through shared
// memory) is done once for the pair.  Conditions are carried as bit 15 of each lane ((threshold - 1 + 0x8000) -
// |difference| has the bit set exactly when the test holds) and applied by masking the clipping bounds, so the
// filter has no branch except the bS 4 case.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t vneg2(uint32_t a) { return __vadd2(~a, 0x00010001u); }
__device__ __forceinline__ uint32_t bits_mask(uint32_t bits) { return bits - (bits >> 15); }     // 0x8000 per lane -> 0x7fff per lane
__device__ __forceinline__ uint32_t vsel(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }
// clip3(-t, t, x) per lane
__device__ __forceinline__ uint32_t vclamp2(uint32_t x, uint32_t t) { return __vmins2(__vmaxs2(x, vneg2(t)), t); }

struct DbkSide {            // what one picture of the pair contributes to a line: strengths, flags, QPs, offsets
	uint32_t bits;
	bool edge0_ok, edge0_bs4;
	int qpc, qpn, aoff, boff;
};

// bytes (s0, s1) of word a and of word b as two packed pairs: [a.s0, 0, b.s0, 0] and [a.s1, 0, b.s1, 0]
__device__ __forceinline__ void pair2(uint32_t a, uint32_t b, uint32_t sel, uint32_t &o0, uint32_t &o1) {
	const uint32_t t = __byte_perm(a, b, sel);            // [a.s0, a.s1, b.s0, b.s1]
	o0 = t & 0x00FF00FFu;
	o1 = (t >> 8) & 0x00FF00FFu;
}
// four consecutive packed pairs -> the word of picture A and the word of picture B
__device__ __forceinline__ void unpair4(const uint32_t *e, uint32_t &wa, uint32_t &wb) {
	const uint32_t x01 = __byte_perm(e[0], e[1], 0x6240), x23 = __byte_perm(e[2], e[3], 0x6240);     // [e0.A, e1.A, e0.B, e1.B]
	wa = __byte_perm(x01, x23, 0x5410);
	wb = __byte_perm(x01, x23, 0x7632);
}

// ------------------------------------------------------------------------------------------------
// Deblocking, two pictures AND two macroblock rows per warp.  In the kernel above the 16 chroma lanes of a
// warp ride along through all four luma edges of a pass (a quarter of the lane-instructions does nothing, and
// the chroma lanes pay for the luma-only half of the filter), and every row waits for the row above through a
// flag.  Here a warp owns rows 2g and 2g + 1 of a picture pair and walks them together, the lower row two
// macroblocks behind the upper one: step s filters macroblock (2g, s) in lanes 0..15 and (2g + 1, s - 2) in lanes
// 16..31 -- what the lower macroblock needs from the row above was finished by this very warp one step earlier,
// so half of the row-to-row dependencies need no flag at all.  Luma and chroma are separate passes, both on all
// 32 lanes: luma = one line (V pass) / column (H pass) of 20 samples per lane and four edges of the full filter;
// chroma = one line / byte column of 10 samples per lane and two edges of the chroma filter (no p1 / q1 terms).
// Tile: one packed pair per 32-bit entry (picture A low / B high 16 bits); columns 12..15 of a macroblock stay in
// the tile as columns -4..-1 of the next one (carry) when both have work; row progress through RowSync flags.
// ------------------------------------------------------------------------------------------------
struct alignas(16) DbkRow3 {    // per macroblock row of the warp
	uint32_t y[20][28];         // luma tile: rows -4..15 at r + 4, columns -4..15 at c + 4 (one packed pair per sample)
	uint32_t c[10][28];         // chroma tile: rows -2..7 at r + 2, byte columns -4..15 at c + 4
	uint4 rec[2][32];           // strength records of the current chunk of 32 macroblocks, [picture]
	uint32_t qp_top[2][32];     // QP word of the records above
	uint4 own[2][2][24];        // [slot][picture][line: luma 0..15, chroma 16..23]: the macroblock's own samples (cp.async ring)
};

template <bool MAY_BS4, bool LUMA>
__device__ __forceinline__ void dbk_filter3(uint32_t *v, const int q, const uint32_t onbits, const uint32_t bs4bits, const uint32_t alpha2, const uint32_t beta2,
                                            const uint32_t tc02) {
	const uint32_t p0 = v[q - 1], p1 = v[q - 2], q0 = v[q], q1 = v[q + 1];
	const uint32_t B15 = 0x80008000u;
	const uint32_t am = alpha2 + 0x7FFF7FFFu, bm = beta2 + 0x7FFF7FFFu;      // a zero threshold never passes
	const uint32_t d0 = __vabsdiffu4(p0, q0);
	const uint32_t act = (am - d0) & (bm - __vabsdiffu4(p1, p0)) & (bm - __vabsdiffu4(q1, q0)) & onbits;
	const uint32_t nb = MAY_BS4 ? act & ~bs4bits : act;
	const uint32_t mn = bits_mask(nb);
	const uint32_t X = ((q0 - p0) << 2) + (p1 - q1) + 0x08040804u;
	const uint32_t xs = (X >> 3) & 0x1FFF1FFFu;
	if (LUMA) {
		const uint32_t p2 = v[q - 3], q2 = v[q + 2];
		const uint32_t ap = (bm - __vabsdiffu4(p2, p0)) & B15, aq = (bm - __vabsdiffu4(q2, q0)) & B15;
		const uint32_t tc = (tc02 + (ap >> 15) + (aq >> 15)) & mn;
		const uint32_t delta = vclamp2(__vadd2(xs, 0xFF00FF00u), tc);
		uint32_t np0 = __vimin_s16x2_relu(__vadd2(p0, delta), 0x00FF00FFu);
		uint32_t nq0 = __vimin_s16x2_relu(__vadd2(q0, vneg2(delta)), 0x00FF00FFu);
		const uint32_t avg = ((p0 + q0 + 0x00010001u) >> 1) & 0x7FFF7FFFu;
		const uint32_t t0 = tc02 & mn;
		const uint32_t zp = __vadd2(((p2 + avg + 0x02000200u - (p1 << 1)) >> 1) & 0x7FFF7FFFu, 0xFF00FF00u);
		const uint32_t zq = __vadd2(((q2 + avg + 0x02000200u - (q1 << 1)) >> 1) & 0x7FFF7FFFu, 0xFF00FF00u);
		uint32_t np1 = __vadd2(p1, vclamp2(zp, t0 & bits_mask(ap)));
		uint32_t nq1 = __vadd2(q1, vclamp2(zq, t0 & bits_mask(aq)));
		if (MAY_BS4) {
			const uint32_t a4 = act & bs4bits;
			if (__any_sync(__activemask(), a4 != 0)) {
				const uint32_t p3 = v[q - 4], q3 = v[q + 3];
				const uint32_t sm = ((alpha2 >> 2) & 0x3FFF3FFFu) + 0x80018001u;          // (alpha >> 2) + 2, as a biased threshold
				const uint32_t st = (sm - d0) & a4;
				const uint32_t msp = bits_mask(st & ap), msq = bits_mask(st & aq), m4 = bits_mask(a4);
				const uint32_t pq = p0 + q0;
				const uint32_t sp0 = ((p2 + 2 * p1 + 2 * pq + q1 + 0x00040004u) >> 3) & 0x1FFF1FFFu;
				const uint32_t sp1 = ((p2 + p1 + pq + 0x00020002u) >> 2) & 0x3FFF3FFFu;
				const uint32_t sp2 = ((2 * p3 + 3 * p2 + p1 + pq + 0x00040004u) >> 3) & 0x1FFF1FFFu;
				const uint32_t wp0 = ((2 * p1 + p0 + q1 + 0x00020002u) >> 2) & 0x3FFF3FFFu;
				const uint32_t sq0 = ((q2 + 2 * q1 + 2 * pq + p1 + 0x00040004u) >> 3) & 0x1FFF1FFFu;
				const uint32_t sq1 = ((q2 + q1 + pq + 0x00020002u) >> 2) & 0x3FFF3FFFu;
				const uint32_t sq2 = ((2 * q3 + 3 * q2 + q1 + pq + 0x00040004u) >> 3) & 0x1FFF1FFFu;
				const uint32_t wq0 = ((2 * q1 + q0 + p1 + 0x00020002u) >> 2) & 0x3FFF3FFFu;
				np0 = vsel(msp, sp0, vsel(m4, wp0, np0));
				np1 = vsel(msp, sp1, np1);
				v[q - 3] = vsel(msp, sp2, p2);
				nq0 = vsel(msq, sq0, vsel(m4, wq0, nq0));
				nq1 = vsel(msq, sq1, nq1);
				v[q + 2] = vsel(msq, sq2, q2);
			}
		}
		v[q - 1] = np0;
		v[q] = nq0;
		v[q - 2] = np1;
		v[q + 1] = nq1;
	} else {
		// chroma: tc = tc0 + 1, only p0 / q0 change; bS 4: (2 p1 + p0 + q1 + 2) >> 2
		const uint32_t tc = (tc02 + 0x00010001u) & mn;
		const uint32_t delta = vclamp2(__vadd2(xs, 0xFF00FF00u), tc);
		uint32_t np0 = __vimin_s16x2_relu(__vadd2(p0, delta), 0x00FF00FFu);
		uint32_t nq0 = __vimin_s16x2_relu(__vadd2(q0, vneg2(delta)), 0x00FF00FFu);
		if (MAY_BS4) {
			const uint32_t m4 = bits_mask(act & bs4bits);
			const uint32_t wp0 = ((2 * p1 + p0 + q1 + 0x00020002u) >> 2) & 0x3FFF3FFFu;
			const uint32_t wq0 = ((2 * q1 + q0 + p1 + 0x00020002u) >> 2) & 0x3FFF3FFFu;
			np0 = vsel(m4, wp0, np0);
			nq0 = vsel(m4, wq0, nq0);
		}
		v[q - 1] = np0;
		v[q] = nq0;
	}
}

// The edges of one line: luma v[0..19] (edges at 4, 8, 12, 16), chroma v[0..9] (samples -2..7: edges at 2 and 6 = luma
// edges 0 and 2 of the strength words)
template <bool LUMA>
__device__ __forceinline__ void dbk_line3(uint32_t *v, const DbkTables &tb, const DbkSide &A, const DbkSide &B, int seg) {
	uint32_t alpha_i, beta_i, alpha_0, beta_0, tci[2], tc0[2];
	{
		const int ia = min(max(A.qpc + A.aoff, 0), 51), ib = min(max(A.qpc + A.boff, 0), 51), q0 = (A.qpc + A.qpn + 1) >> 1;
		const int ja = min(max(q0 + A.aoff, 0), 51), jb = min(max(q0 + A.boff, 0), 51);
		alpha_i = tb.alpha[ia]; beta_i = tb.beta[ib]; alpha_0 = tb.alpha[ja]; beta_0 = tb.beta[jb];
		tci[0] = *reinterpret_cast<const uint32_t *>(tb.tc0[ia]);
		tc0[0] = *reinterpret_cast<const uint32_t *>(tb.tc0[ja]);
	}
	{
		const int ia = min(max(B.qpc + B.aoff, 0), 51), ib = min(max(B.qpc + B.boff, 0), 51), q0 = (B.qpc + B.qpn + 1) >> 1;
		const int ja = min(max(q0 + B.aoff, 0), 51), jb = min(max(q0 + B.boff, 0), 51);
		alpha_i |= (uint32_t)tb.alpha[ia] << 16; beta_i |= (uint32_t)tb.beta[ib] << 16; alpha_0 |= (uint32_t)tb.alpha[ja] << 16; beta_0 |= (uint32_t)tb.beta[jb] << 16;
		tci[1] = *reinterpret_cast<const uint32_t *>(tb.tc0[ia]);
		tc0[1] = *reinterpret_cast<const uint32_t *>(tb.tc0[ja]);
	}
	const uint32_t la = A.bits >> (2 * seg), lb = B.bits >> (2 * seg);       // this line's strength of edge e in bits 8e, 8e + 1
	{
		const int bsa = la & 3, bsb = lb & 3;
		const bool ona = bsa != 0 && A.edge0_ok, onb = bsb != 0 && B.edge0_ok;
		const uint32_t onbits = (ona ? 0x8000u : 0u) | (onb ? 0x80000000u : 0u);
		const uint32_t bs4bits = (A.edge0_bs4 ? 0x8000u : 0u) | (B.edge0_bs4 ? 0x80000000u : 0u);
		const uint32_t tc = ((tc0[0] >> (8 * (bsa ? bsa - 1 : 0))) & 255u) | (((tc0[1] >> (8 * (bsb ? bsb - 1 : 0))) & 255u) << 16);
		if (__any_sync(__activemask(), onbits != 0)) dbk_filter3<true, LUMA>(v, LUMA ? 4 : 2, onbits, bs4bits, alpha_0, beta_0, tc);
	}
	if (LUMA) {
		if (!__any_sync(__activemask(), ((A.bits | B.bits) & 0xffffff00u) != 0)) return;          // no inner edge of any macroblock of the warp has a strength
#pragma unroll
		for (int e = 1; e < 4; ++e) {
			const int bsa = (la >> (8 * e)) & 3, bsb = (lb >> (8 * e)) & 3;
			const uint32_t onbits = (bsa ? 0x8000u : 0u) | (bsb ? 0x80000000u : 0u);
			const uint32_t tc = ((tci[0] >> (8 * (bsa ? bsa - 1 : 0))) & 255u) | (((tci[1] >> (8 * (bsb ? bsb - 1 : 0))) & 255u) << 16);
			dbk_filter3<false, true>(v, 4 + 4 * e, onbits, 0u, alpha_i, beta_i, tc);
		}
	} else {
		const int bsa = (la >> 16) & 3, bsb = (lb >> 16) & 3;
		const uint32_t onbits = (bsa ? 0x8000u : 0u) | (bsb ? 0x80000000u : 0u);
		if (!__any_sync(__activemask(), onbits != 0)) return;
		const uint32_t tc = ((tci[0] >> (8 * (bsa ? bsa - 1 : 0))) & 255u) | (((tci[1] >> (8 * (bsb ? bsb - 1 : 0))) & 255u) << 16);
		dbk_filter3<false, false>(v, 6, onbits, 0u, alpha_i, beta_i, tc);
	}
}

// the strength records of macroblocks xs .. xs + 31 of row mby (both pictures) and the QP words of the row above, into the row's staging
__device__ __forceinline__ void dbk3_load_chunk(DbkRow3 &R, const H4Job &job0, const H4Job &job1, bool valid1, int mby, int xs, int lane) {
	const int mb_w = job0.mb_w, mbx = xs + lane;
	uint4 ra = make_uint4(0, 0, 0, 0), rb = ra;
	uint32_t ta = 0, tb = 0;
	if (mbx < mb_w) {
		const size_t mbi = (size_t)mby * mb_w + mbx;
		ra = __ldg(reinterpret_cast<const uint4 *>(job0.dbk + mbi));
		if (valid1) rb = __ldg(reinterpret_cast<const uint4 *>(job1.dbk + mbi));
		if (mby > 0) {
			ta = __ldg(reinterpret_cast<const uint32_t *>(job0.dbk + mbi - mb_w) + 2);
			tb = __ldg(reinterpret_cast<const uint32_t *>(job1.dbk + mbi - mb_w) + 2);
		}
	}
	if ((ra.z >> 24) & B2_DBK_OFF) ra.x = ra.y = 0;
	if (((rb.z >> 24) & B2_DBK_OFF) || !valid1) rb.x = rb.y = 0;
	__syncwarp();                      // the previous chunk's readers are done
	R.rec[0][lane] = ra; R.rec[1][lane] = rb;
	R.qp_top[0][lane] = ta; R.qp_top[1][lane] = tb;
	__syncwarp();
}

#ifndef B2_DBK3_WARPS
#define B2_DBK3_WARPS 1
#endif
constexpr int DBK3_WARPS = B2_DBK3_WARPS;          // row pairs per CTA
#ifndef B2_DBK3_WARPS_PER_SM
#define B2_DBK3_WARPS_PER_SM 16
#endif
constexpr int DBK3_CTAS = B2_DBK3_WARPS_PER_SM / DBK3_WARPS;
#ifndef B2_DBK3_PUBLISH_EVERY
#define B2_DBK3_PUBLISH_EVERY 2
#endif

// Rows r0 and r0 + 1 (if has_r1) of a picture pair, all macroblocks, left to right.
__device__ void deblock_rows3(const H4Job &job0, const H4Job &job1, const bool valid1, const DbkTables &tb, DbkRow3 *rows, const int r0, const bool has_r1,
                              const int lane, RowSync &sy) {
	const int mb_w = job0.mb_w, W = mb_w * 16, H = job0.mb_h * 16;
	const int half = lane >> 4, l16 = lane & 15;
	DbkRow3 &R = rows[half];
	const int mby = r0 + half;
	const bool row_ok = half == 0 || has_r1;
	uint8_t *const fy0 = job0.frames + job0.frame_bytes * job0.dst, *const fy1 = job1.frames + job1.frame_bytes * job1.dst;
	uint8_t *const fc0 = fy0 + (size_t)W * H, *const fc1 = fy1 + (size_t)W * H;
	const size_t yline = (size_t)(mby * 16 + l16) * W;               // luma V pass: this lane's line
	const int cline = l16 >> 1, comp = l16 & 1;                      // chroma V pass: line 0..7, component
	const size_t cl_off = (size_t)(mby * 8 + cline) * W;
	const uint32_t csel = comp ? 0x7531u : 0x6420u;
	uint32_t *const trow = R.y[4 + l16];
	uint32_t *const crow = R.c[2 + cline];
	bool carry_in = false;
	const int steps = mb_w + (has_r1 ? 2 : 0);
	// own samples of the macroblocks of step `st` into ring slot st & 1: luma lanes their line of both pictures; the chroma
	// lines of a macroblock (8 lines x 2 pictures) one per lane
	auto fetch_own = [&](int st) {
		const int xx = half ? st - 2 : st;
		if (row_ok && xx >= 0 && xx < mb_w) {
			const int sl = st & 1;
			cp_async16(&R.own[sl][0][l16], fy0 + yline + xx * 16);
			if (valid1) cp_async16(&R.own[sl][1][l16], fy1 + yline + xx * 16);
			const int pic = l16 >> 3, cln = l16 & 7;
			if (pic == 0 || valid1) cp_async16(&R.own[sl][pic][16 + cln], (pic ? fc1 : fc0) + (size_t)(mby * 8 + cln) * W + xx * 16);
		}
		cp_async_commit();
	};
	fetch_own(0);
	for (int s = 0; s < steps; ++s) {
		const int x = half ? s - 2 : s;
		const bool in_row = row_ok && x >= 0 && x < mb_w;
		if (s < mb_w && (s & 31) == 0) dbk3_load_chunk(rows[0], job0, job1, valid1, r0, s, lane);
		if (has_r1 && s >= 2 && s - 2 < mb_w && ((s - 2) & 31) == 0) dbk3_load_chunk(rows[1], job0, job1, valid1, r0 + 1, s - 2, lane);
		fetch_own(s + 1);
		const int k = x & 31;
		uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0;
		bool next_need = false;
		if (in_row) {
			d0 = R.rec[0][k]; d1 = R.rec[1][k];
			if (k < 31 && x + 1 < mb_w) {
				const uint4 n0 = R.rec[0][k + 1], n1 = R.rec[1][k + 1];
				next_need = (n0.x | n0.y | n1.x | n1.y) != 0;
			}
		}
		const bool need = (d0.x | d0.y | d1.x | d1.y) != 0;        // (records of macroblocks with the filter off were cleared when staged)
		const bool carry_out = need && next_need;                  // the next macroblock takes over columns 12..15
		const unsigned need_bal = __ballot_sync(0xffffffffu, need);
		if (need_bal) {
			if ((need_bal & 1u) && r0 > 0) sy.wait(min(s + 2, mb_w), lane);
			uint32_t ql0 = 0, ql1 = 0, qt0 = 0, qt1 = 0;
			if (need) {
				if (k > 0) { ql0 = R.rec[0][k - 1].z; ql1 = R.rec[1][k - 1].z; }
				else if (x > 0) {
					const size_t mbi = (size_t)mby * mb_w + x;
					ql0 = __ldg(reinterpret_cast<const uint32_t *>(job0.dbk + mbi - 1) + 2);
					ql1 = __ldg(reinterpret_cast<const uint32_t *>(job1.dbk + mbi - 1) + 2);
				}
				qt0 = R.qp_top[0][k]; qt1 = R.qp_top[1][k];
			}
			const int flags0 = (d0.z >> 24) & 255, flags1 = (d1.z >> 24) & 255;
			DbkSide sA, sB;
			sA.aoff = (int8_t)(d0.w & 255); sA.boff = (int8_t)((d0.w >> 8) & 255);
			sB.aoff = (int8_t)(d1.w & 255); sB.boff = (int8_t)((d1.w >> 8) & 255);
			// rows above the macroblock, one word per lane and picture: luma rows -4..-1, chroma rows -2..-1 (lanes < 8 of the half)
			uint32_t tw0 = 0, tw1 = 0, tc0w = 0, tc1w = 0;
			if (need && mby > 0) {
				const size_t off = (size_t)(mby * 16 - 4 + (l16 >> 2)) * W + x * 16 + 4 * (l16 & 3);
				tw0 = __ldcg(reinterpret_cast<const uint32_t *>(fy0 + off));
				tw1 = __ldcg(reinterpret_cast<const uint32_t *>(fy1 + off));
				if (l16 < 8) {
					const size_t offc = (size_t)(mby * 8 - 2 + (l16 >> 2)) * W + x * 16 + 4 * (l16 & 3);
					tc0w = __ldcg(reinterpret_cast<const uint32_t *>(fc0 + offc));
					tc1w = __ldcg(reinterpret_cast<const uint32_t *>(fc1 + offc));
				}
			}
			// left four columns when the previous macroblock did not leave them in the tile
			uint32_t wl0 = 0, wl1 = 0, cl0 = 0, cl1 = 0;
			if (need && !carry_in && x > 0) {
				wl0 = __ldcg(reinterpret_cast<const uint32_t *>(fy0 + yline + x * 16 - 4));
				wl1 = __ldcg(reinterpret_cast<const uint32_t *>(fy1 + yline + x * 16 - 4));
				cl0 = __ldcg(reinterpret_cast<const uint32_t *>(fc0 + cl_off + x * 16 - 4));
				cl1 = __ldcg(reinterpret_cast<const uint32_t *>(fc1 + cl_off + x * 16 - 4));
			}
			cp_async_wait<1>();
			__syncwarp();
			const int sl = s & 1;
			// ---- luma V pass: lane = line ------------------------------------------------------------------
			{
				uint32_t v[20];
				const uint4 oa = R.own[sl][0][l16], ob = R.own[sl][1][l16];
				const uint32_t w0[5] = {wl0, oa.x, oa.y, oa.z, oa.w}, w1[5] = {wl1, ob.x, ob.y, ob.z, ob.w};
#pragma unroll
				for (int j = 0; j < 5; ++j) {
					pair2(w0[j], w1[j], 0x5410u, v[4 * j], v[4 * j + 1]);
					pair2(w0[j], w1[j], 0x7632u, v[4 * j + 2], v[4 * j + 3]);
				}
				if (carry_in) {
					const uint4 c4 = *reinterpret_cast<const uint4 *>(trow);
					v[0] = c4.x; v[1] = c4.y; v[2] = c4.z; v[3] = c4.w;
				}
				if (__any_sync(0xffffffffu, (d0.x | d1.x) != 0)) {
					sA.bits = d0.x; sB.bits = d1.x;
					sA.edge0_ok = flags0 & B2_DBK_LEFT_EDGE; sB.edge0_ok = flags1 & B2_DBK_LEFT_EDGE;
					sA.edge0_bs4 = flags0 & B2_DBK_LEFT_BS4; sB.edge0_bs4 = flags1 & B2_DBK_LEFT_BS4;
					sA.qpc = (int8_t)(d0.z & 255); sB.qpc = (int8_t)(d1.z & 255);
					sA.qpn = (int8_t)(ql0 & 255); sB.qpn = (int8_t)(ql1 & 255);
					dbk_line3<true>(v, tb, sA, sB, l16 >> 2);
				}
				if (need) {
#pragma unroll
					for (int j = 0; j < 5; ++j) reinterpret_cast<uint4 *>(trow)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
					uint4 e;
					pair2(tw0, tw1, 0x5410u, e.x, e.y);
					pair2(tw0, tw1, 0x7632u, e.z, e.w);
					*reinterpret_cast<uint4 *>(R.y[l16 >> 2] + 4 + 4 * (l16 & 3)) = e;
				}
			}
			// ---- chroma V pass: lane = line of one component (samples -2..7 in v[0..9]) ----------------------
			const int qsel = 8 + 8 * comp;
			{
				uint32_t v[10];
				const uint4 oa = R.own[sl][0][16 + cline], ob = R.own[sl][1][16 + cline];
				const uint32_t w0[5] = {cl0, oa.x, oa.y, oa.z, oa.w}, w1[5] = {cl1, ob.x, ob.y, ob.z, ob.w};
#pragma unroll
				for (int j = 0; j < 5; ++j) pair2(w0[j], w1[j], csel, v[2 * j], v[2 * j + 1]);
				if (carry_in) { v[0] = crow[comp]; v[1] = crow[2 + comp]; }
				if (__any_sync(0xffffffffu, (d0.x | d1.x) != 0)) {
					sA.bits = d0.x; sB.bits = d1.x;
					sA.edge0_ok = flags0 & B2_DBK_LEFT_EDGE; sB.edge0_ok = flags1 & B2_DBK_LEFT_EDGE;
					sA.edge0_bs4 = flags0 & B2_DBK_LEFT_BS4; sB.edge0_bs4 = flags1 & B2_DBK_LEFT_BS4;
					sA.qpc = (int8_t)((d0.z >> qsel) & 255); sB.qpc = (int8_t)((d1.z >> qsel) & 255);
					sA.qpn = (int8_t)((ql0 >> qsel) & 255); sB.qpn = (int8_t)((ql1 >> qsel) & 255);
					dbk_line3<false>(v, tb, sA, sB, cline >> 1);
				}
				if (need) {
#pragma unroll
					for (int i = 0; i < 10; ++i) crow[2 * i + comp] = v[i];
					if (l16 < 8) {
						uint4 e;
						pair2(tc0w, tc1w, 0x5410u, e.x, e.y);
						pair2(tc0w, tc1w, 0x7632u, e.z, e.w);
						*reinterpret_cast<uint4 *>(R.c[l16 >> 2] + 4 + 4 * (l16 & 3)) = e;
					}
				}
			}
			__syncwarp();
			// ---- H pass: lane = luma column, then chroma byte column -----------------------------------------
			if (__any_sync(0xffffffffu, (d0.y | d1.y) != 0)) {
				sA.bits = d0.y; sB.bits = d1.y;
				sA.edge0_ok = flags0 & B2_DBK_TOP_EDGE; sB.edge0_ok = flags1 & B2_DBK_TOP_EDGE;
				sA.edge0_bs4 = flags0 & B2_DBK_TOP_BS4; sB.edge0_bs4 = flags1 & B2_DBK_TOP_BS4;
				{
					uint32_t v[20];
#pragma unroll
					for (int i = 0; i < 20; ++i) v[i] = R.y[i][4 + l16];
					sA.qpc = (int8_t)(d0.z & 255); sB.qpc = (int8_t)(d1.z & 255);
					sA.qpn = (int8_t)(qt0 & 255); sB.qpn = (int8_t)(qt1 & 255);
					dbk_line3<true>(v, tb, sA, sB, l16 >> 2);
					if (need) {
#pragma unroll
						for (int i = 1; i < 20; ++i) R.y[i][4 + l16] = v[i];
					}
				}
				{
					uint32_t v[10];
#pragma unroll
					for (int i = 0; i < 10; ++i) v[i] = R.c[i][4 + l16];
					sA.qpc = (int8_t)((d0.z >> qsel) & 255); sB.qpc = (int8_t)((d1.z >> qsel) & 255);
					sA.qpn = (int8_t)((qt0 >> qsel) & 255); sB.qpn = (int8_t)((qt1 >> qsel) & 255);
					dbk_line3<false>(v, tb, sA, sB, l16 >> 2);
					if (need) {
#pragma unroll
						for (int i = 1; i < 10; ++i) R.c[i][4 + l16] = v[i];
					}
				}
				__syncwarp();
			}
			// ---- store what is final: columns -4..11 (or ..15 when nobody takes them over) of the macroblock's lines,
			//      and the rows above it (luma -3..-1, chroma -1: lanes 0..3 of the half) --------------------------------
			if (need) {
				{
					uint8_t *dA = fy0 + yline + x * 16, *dB = fy1 + yline + x * 16;
#pragma unroll
					for (int j = 0; j < 5; ++j) {
						if ((j == 0 && x == 0) || (j == 4 && carry_out)) continue;
						uint32_t wa, wb;
						const uint4 e4 = *reinterpret_cast<const uint4 *>(trow + 4 * j);
						const uint32_t ev[4] = {e4.x, e4.y, e4.z, e4.w};
						unpair4(ev, wa, wb);
						__stcg(reinterpret_cast<uint32_t *>(dA + 4 * (j - 1)), wa);
						if (valid1) __stcg(reinterpret_cast<uint32_t *>(dB + 4 * (j - 1)), wb);
					}
				}
				{
					// chroma: lane = (picture, line)
					const int pic = l16 >> 3, cln = l16 & 7;
					if (pic == 0 || valid1) {
						const uint32_t *src = R.c[2 + cln];
						uint8_t *d = (pic ? fc1 : fc0) + (size_t)(mby * 8 + cln) * W + x * 16;
#pragma unroll
						for (int j = 0; j < 5; ++j) {
							if ((j == 0 && x == 0) || (j == 4 && carry_out)) continue;
							uint32_t wa, wb;
							const uint4 e4 = *reinterpret_cast<const uint4 *>(src + 4 * j);
							const uint32_t ev[4] = {e4.x, e4.y, e4.z, e4.w};
							unpair4(ev, wa, wb);
							__stcg(reinterpret_cast<uint32_t *>(d + 4 * (j - 1)), pic ? wb : wa);
						}
					}
				}
				const bool top0 = flags0 & B2_DBK_TOP_EDGE, top1 = (flags1 & B2_DBK_TOP_EDGE) && valid1;
				if (l16 < 4 && (top0 || top1)) {
					const bool tl_luma = l16 < 3;
					const int r = tl_luma ? l16 - 3 : -1;                 // luma rows -3..-1, chroma row -1
					const uint32_t *src = (tl_luma ? R.y[4 + r] : R.c[2 + r]) + 4;
					const size_t off = (tl_luma ? (size_t)(mby * 16 + r) * W : (size_t)(mby * 8 + r) * W) + x * 16;
					uint8_t *dA = (tl_luma ? fy0 : fc0) + off, *dB = (tl_luma ? fy1 : fc1) + off;
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						uint32_t wa, wb;
						const uint4 e4 = *reinterpret_cast<const uint4 *>(src + 4 * j);
						const uint32_t ev[4] = {e4.x, e4.y, e4.z, e4.w};
						unpair4(ev, wa, wb);
						if (top0) __stcg(reinterpret_cast<uint32_t *>(dA + 4 * j), wa);
						if (top1) __stcg(reinterpret_cast<uint32_t *>(dB + 4 * j), wb);
					}
				}
			}
			__syncwarp();
			if (carry_out) {
				*reinterpret_cast<uint4 *>(trow) = *reinterpret_cast<const uint4 *>(trow + 16);
				if (l16 < 8) *reinterpret_cast<uint4 *>(R.c[2 + l16]) = *reinterpret_cast<const uint4 *>(R.c[2 + l16] + 16);
			}
		}
		carry_in = carry_out;
		// progress of the pair's last row: macroblocks < n are done
		const int n = has_r1 ? s - 1 : s + 1;
		if (n > 0 && ((n & (B2_DBK3_PUBLISH_EVERY - 1)) == 0 || n == mb_w)) sy.publish(n, lane);
		__syncwarp();
	}
	cp_async_wait<0>();
}

__global__ void __launch_bounds__(32 * DBK3_WARPS, DBK3_CTAS)
h264_deblock3_kernel(const H4Job *__restrict__ jobs, int n_jobs, int mb_w, int mb_h, int *__restrict__ sync_area) {
	__shared__ DbkRow3 s_rows[DBK3_WARPS][2];
	__shared__ DbkTables s_tb;
	__shared__ int s_ticket;
	__shared__ int s_prog[DBK3_WARPS];
	for (int i = threadIdx.x; i < 52; i += blockDim.x) {
		s_tb.alpha[i] = c_alpha[i];
		s_tb.beta[i] = c_beta[i];
		s_tb.tc0[i][0] = c_tc0[i][0]; s_tb.tc0[i][1] = c_tc0[i][1]; s_tb.tc0[i][2] = c_tc0[i][2]; s_tb.tc0[i][3] = 0;
	}
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	int *ticket = sync_area;
	int *progress = sync_area + 32;           // [pairs][mb_h]: macroblocks of the row that are final (rows read across CTAs only)
	const int row_pairs = (mb_h + 1) / 2, groups = (row_pairs + DBK3_WARPS - 1) / DBK3_WARPS, pairs = (n_jobs + 1) / 2;
	const int total = pairs * groups;
	for (;;) {
		__syncthreads();                       // the previous group is complete: s_ticket / s_prog may be reused
		if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1);
		if (threadIdx.x < DBK3_WARPS) s_prog[threadIdx.x] = 0;
		__syncthreads();
		const int t = s_ticket;
		if (t >= total) return;
		// group-major ticket order, as in h264_wave_kernel
		const int j = t % pairs, rp = (t / pairs) * DBK3_WARPS + warp;
		if (rp >= row_pairs) continue;
		const int r0 = 2 * rp;
		const bool has_r1 = r0 + 1 < mb_h;
		const int last = has_r1 ? r0 + 1 : r0;
		RowSync sy;
		sy.sm_above = warp > 0 ? &s_prog[warp - 1] : nullptr;
		sy.gl_above = (warp == 0 && r0 > 0) ? progress + (size_t)j * mb_h + r0 - 1 : nullptr;
		sy.sm_my = &s_prog[warp];
		sy.gl_my = (warp == DBK3_WARPS - 1 && last + 1 < mb_h) ? progress + (size_t)j * mb_h + last : nullptr;
		sy.seen = 0;
		sy.lead = 0;
		sy.limit = mb_w;
		const bool valid1 = 2 * j + 1 < n_jobs;
		const H4Job job0 = jobs[2 * j], job1 = jobs[valid1 ? 2 * j + 1 : 2 * j];
		deblock_rows3(job0, job1, valid1, s_tb, s_rows[warp], r0, has_r1, lane, sy);
	}
}

// Intra macroblocks of pictures that have few of them (P / B pictures).  A warp takes 32 consecutive
// macroblocks (raster order) at a time and reconstructs the intra ones among them; an intra
// macroblock waits only for those of its left / top-left / top / top-right neighbours that are intra
// themselves and that it may read (availability bits) -- inter neighbours were finished by the
// kernel before.  Chunks are handed out in raster order, so a neighbour always belongs to a chunk
// with an earlier ticket: no deadlock as long as every CTA of the grid is resident (persistent grid).
__global__ void __launch_bounds__(32 * WAVE_WARPS, WAVE_CTAS_INTRA)
h264_intra_sparse_kernel(const H4Job *__restrict__ jobs, int n_jobs, int mb_w, int mb_h, int *__restrict__ sync_area) {
	__shared__ IntraTile s_tile[WAVE_WARPS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	int *ticket = sync_area + 1;
	int *done = sync_area + 32;                // [n_jobs][n_mbs]
	const int n_mbs = mb_w * mb_h, chunks = (n_mbs + 31) / 32;
	const int total = n_jobs * chunks;
	for (;;) {
		int t = 0;
		if (lane == 0) t = atomicAdd(ticket, 1);
		t = __shfl_sync(0xffffffffu, t, 0);
		if (t >= total) return;
		const int j = t % n_jobs, c = t / n_jobs;
		const H4Job job = jobs[j];
		if (!h264_intra_sparse(job.n_intra, n_mbs) || job.n_intra == 0) continue;
		int *flags = done + (size_t)j * n_mbs;
		const int mb = c * 32 + lane;
		unsigned todo = __ballot_sync(0xffffffffu, mb < n_mbs && job.mbs[mb].type != B2_H4_INTER);
		while (todo) {
			const int k = __ffs(todo) - 1;
			todo &= todo - 1;
			const int m = c * 32 + k, mbx = m % mb_w, mby = m / mb_w;
			const b2_h264_mb hdr = job.mbs[m];
			intra_residual(job, hdr, s_tile[warp], lane);
			if (lane < 4) {
				// neighbour of lane: 0 left, 1 top, 2 top-right, 3 top-left (the availability bit order)
				const int a = hdr.intra_avail | hdr.intra_avail_c;
				const int nx = mbx + (lane == 0 || lane == 3 ? -1 : (lane == 2 ? 1 : 0)), ny = mby - (lane != 0);
				if (((a >> lane) & 1) && nx >= 0 && nx < mb_w && ny >= 0) {
					const int nb = ny * mb_w + nx;
					if (job.mbs[nb].type != B2_H4_INTER)
						spin_until(flags + nb, 1);
				}
			}
			__syncwarp();
			intra_mb(job, hdr, mbx, mby, s_tile[warp], lane);
			__syncwarp();
			if (lane == 0) st_release(flags + m, 1);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Motion vectors on the device (reference calc_mv* / P_Skip, h264.cpp:6669-6724, 7379-7604,
// standard 8.4.1.1, 8.4.1.3).  The parse of a Baseline-profile P picture hands over, per partition, the reference
// index and the motion vector DIFFERENCE; prediction needs the final vectors of the left, top, top-right and top-left
// neighbours, so the macroblocks are a wavefront like intra prediction: one warp per macroblock row (tickets in
// picture-major row order), a row waits until the row above is two macroblocks ahead.  The derivation itself is
// serial inside a macroblock (a partition predicts from the ones before it) and runs on lane 0; it is a few
// hundred instructions per macroblock and hidden behind the host's parse of the next picture.  Output: the
// b2_h264_mc record of every partition (default weights), the picture's motion field (for the rows below) and the
// macroblock each partition belongs to (for h264_items_kernel).
// ------------------------------------------------------------------------------------------------
struct MvNb { int ref; int mvx, mvy; };       // ref: -2 unavailable, -1 intra
__device__ __forceinline__ int median3(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }

__global__ void __launch_bounds__(32)
h264_mv_kernel(const H4Job *__restrict__ jobs, int n_jobs, int mb_w, int mb_h, int *__restrict__ sync_area) {
	// blocks (x, y), x = -1 .. 4, y = -1 .. 3 of the current macroblock's neighbourhood at [y + 1][x + 1]
	__shared__ int8_t s_ref[5][6];
	__shared__ uint32_t s_mv[5][6];
	// what the lanes fetch for lane 0 and store for it: the macroblock's partitions, the records made of them
	__shared__ b2_h264_part s_part[16];
	__shared__ uint4 s_rec[16][2];
	int *ticket = sync_area;
	int *progress = sync_area + 32;           // [n_jobs][mb_h]: macroblocks of the row whose motion is final
	const int lane = threadIdx.x;
	for (;;) {
		int t = 0;
		if (lane == 0) t = atomicAdd(ticket, 1);
		t = __shfl_sync(0xffffffffu, t, 0);
		if (t >= n_jobs * mb_h) return;
		const int j = t / mb_h, mby = t % mb_h;
		const H4Job job = jobs[j];
		if (!job.derive_mv) continue;
		int *prog_above = progress + (size_t)j * mb_h + mby - 1, *prog_mine = progress + (size_t)j * mb_h + mby;
		b2_h264_mc *mcs = const_cast<b2_h264_mc *>(job.mcs);
		int seen = 0;
		// The derivation is serial (lane 0); the other lanes turn its memory round trips into one per macroblock: the
		// header of the NEXT macroblock is fetched while this one is worked on, partitions and the row above come in
		// with one load per lane, records / owners / motion field go out with one store per lane.
		uint32_t hdr_next = 0, idx_next = 0, cnt_next = 0;
		if (lane == 0) {
			const b2_h264_mb *m0 = job.mbs + (size_t)mby * mb_w;
			hdr_next = __ldg(reinterpret_cast<const uint32_t *>(m0));             // type, flags, ...
			idx_next = __ldg(&m0->mc_idx);
			cnt_next = __ldg(&m0->mc_cnt);
		}
		for (int mbx = 0; mbx < mb_w; ++mbx) {
			const int mb = mby * mb_w + mbx;
			// header of this macroblock (lane 0 had asked for it), broadcast; then the request for the next one
			const uint32_t h0 = __shfl_sync(0xffffffffu, hdr_next, 0), mc_idx = __shfl_sync(0xffffffffu, idx_next, 0);
			const int cnt = min((int)__shfl_sync(0xffffffffu, cnt_next, 0), 16);
			if (lane == 0 && mbx + 1 < mb_w) {
				const b2_h264_mb *m1 = job.mbs + mb + 1;
				hdr_next = __ldg(reinterpret_cast<const uint32_t *>(m1));
				idx_next = __ldg(&m1->mc_idx);
				cnt_next = __ldg(&m1->mc_cnt);
			}
			const int type = h0 & 255, flags = (h0 >> 8) & 255, avail = flags >> B2_H4_NB_SHIFT;
			int8_t *fr = job.field_ref + (size_t)mb * 16;
			uint32_t *fm = job.field_mv + (size_t)mb * 16;
			const bool inter = type == B2_H4_INTER;
			// ---- fetch: partitions (lanes 0..15), the row above (lanes 16..21) ------------------------------------------
			if (inter && lane < cnt) s_part[lane] = job.parts[mc_idx + lane];
			if (inter && mby > 0 && (avail & 14)) {
				const int want = min(mbx + 2, mb_w);
				if (seen < want) {
					if (lane == 0) seen = spin_until(prog_above, want);
					seen = __shfl_sync(0xffffffffu, seen, 0);
				}
			}
			if (lane == 0) {
				// left column = the previous macroblock's right column (still in the tile); the rest is not decoded yet / outside
				for (int y = 0; y < 4; ++y) {
					if (inter && (avail & 1)) { s_ref[y + 1][0] = s_ref[y + 1][4]; s_mv[y + 1][0] = s_mv[y + 1][4]; }
					else s_ref[y + 1][0] = -2;
					for (int x = 0; x < 5; ++x) s_ref[y + 1][x + 1] = inter ? -2 : -1;
				}
			}
			if (inter && lane >= 16 && lane < 22) {
				const int x = lane - 17;                   // -1 .. 4
				const bool ok = x < 0 ? (avail & 8) != 0 : (x < 4 ? (avail & 2) != 0 : (avail & 4) != 0);
				int r = -2;
				uint32_t v = 0;
				if (ok) {
					const int nmb = mb - mb_w + (x < 0 ? -1 : (x < 4 ? 0 : 1)), blk = 12 + (x & 3);
					r = (int8_t)__ldcg(reinterpret_cast<const char *>(job.field_ref) + (size_t)nmb * 16 + blk);
					v = __ldcg(job.field_mv + (size_t)nmb * 16 + blk);
				}
				s_ref[0][x + 1] = (int8_t)r;
				s_mv[0][x + 1] = v;
			}
			__syncwarp();
			// ---- derive (lane 0) -------------------------------------------------------------------------------------------
			if (lane == 0 && inter) {
				auto nb = [&](int x, int y) {
					MvNb n;
					n.ref = s_ref[y + 1][x + 1];
					const uint32_t v = n.ref >= 0 ? s_mv[y + 1][x + 1] : 0u;
					n.mvx = (int16_t)(v & 0xffff);
					n.mvy = (int16_t)(v >> 16);
					return n;
				};
				for (int i = 0; i < cnt; ++i) {
					const b2_h264_part pt = s_part[i];
					const int x4 = pt.pos & 3, y4 = pt.pos >> 2, pw = pt.size & 15, ph = pt.size >> 4, refidx = pt.ref_idx;
					MvNb a = nb(x4 - 1, y4), b = nb(x4, y4 - 1), c = nb(x4 + pw, y4 - 1);
					if (c.ref == -2) c = nb(x4 - 1, y4 - 1);
					int px, py;
					bool done = false;
					if (flags & B2_H4_PSKIP) {
						if (a.ref == -2 || b.ref == -2 || (a.ref == 0 && a.mvx == 0 && a.mvy == 0) || (b.ref == 0 && b.mvx == 0 && b.mvy == 0)) { px = py = 0; done = true; }
					} else if (pw == 4 && ph == 2) {          // 16x8: the upper partition from above, the lower one from the left
						if (i == 0 && b.ref == refidx) { px = b.mvx; py = b.mvy; done = true; }
						if (i == 1 && a.ref == refidx) { px = a.mvx; py = a.mvy; done = true; }
					} else if (pw == 2 && ph == 4) {          // 8x16: the left partition from the left, the right one from above-right
						if (i == 0 && a.ref == refidx) { px = a.mvx; py = a.mvy; done = true; }
						if (i == 1 && c.ref == refidx) { px = c.mvx; py = c.mvy; done = true; }
					}
					if (!done) {
						if (b.ref == -2 && c.ref == -2 && a.ref != -2) { b = a; c = a; }
						const bool ma = a.ref == refidx, mb_ = b.ref == refidx, mc = c.ref == refidx;
						if ((int)ma + (int)mb_ + (int)mc == 1) {
							const MvNb &s1 = ma ? a : (mb_ ? b : c);
							px = s1.mvx; py = s1.mvy;
						} else {
							px = median3(a.mvx, b.mvx, c.mvx);
							py = median3(a.mvy, b.mvy, c.mvy);
						}
					}
					const int mvx = (int16_t)(px + pt.mvd[0]), mvy = (int16_t)(py + pt.mvd[1]);
					const uint32_t packed = (uint32_t)(uint16_t)mvx | ((uint32_t)(uint16_t)mvy << 16);
					for (int y = y4; y < y4 + ph; ++y)
						for (int x = x4; x < x4 + pw; ++x) { s_ref[y + 1][x + 1] = (int8_t)refidx; s_mv[y + 1][x + 1] = packed; }
					// the motion record: list 0 only, default weights (include/b2_recon.h b2_h264_mc)
					s_rec[i][0] = make_uint4((uint32_t)pt.pos | ((uint32_t)pt.size << 8) | ((uint32_t)(uint8_t)pt.ref_frame << 16) | 0xff000000u,
					                         (uint32_t)(uint8_t)refidx | 0x0000ff00u, packed, 0u);
					s_rec[i][1] = make_uint4(0x01010101u, 0x00000101u, 0u, 0u);
				}
			}
			__syncwarp();
			// ---- store: records and owners (one per lane), the macroblock's motion field (16 lanes) ------------------------
			if (inter && lane < cnt) {
				uint4 *rec = reinterpret_cast<uint4 *>(mcs + mc_idx + lane);
				rec[0] = s_rec[lane][0];
				rec[1] = s_rec[lane][1];
				job.owner[mc_idx + lane] = (uint32_t)mb;
			}
			if (lane < 16) {
				const int x = lane & 3, y = lane >> 2;
				fr[lane] = s_ref[y + 1][x + 1];
				fm[lane] = inter ? s_mv[y + 1][x + 1] : 0u;
			}
			__syncwarp();
			if (!inter && lane == 0)
				for (int y = 0; y < 4; ++y) s_mv[y + 1][4] = 0;          // (the next macroblock's left column: intra, no vector)
			// rows below read this macroblock's bottom row
			if (mby + 1 < mb_h) {
				__threadfence();
				__syncwarp();
				if (lane == 0) st_release(prog_mine, mbx + 1);
			}
		}
	}
}

// Work list of the inter kernel from the motion records (what the host front end's emit_mc does, h264_front.cpp):
// one CTA per picture counts the strips of every record per bucket (touches-the-edge, class, height), turns the
// counts into offsets and writes the items; the order inside a bucket is whatever the atomics give (every strip
// writes its own pixels, the order decides nothing).  Writes the picture's n_items into its job.
__global__ void __launch_bounds__(1024)
h264_items_kernel(H4Job *__restrict__ jobs, int mb_w, int mb_h) {
	__shared__ uint32_t s_cnt[72], s_at[72];
	H4Job &job = jobs[blockIdx.x];
	if (!job.derive_mv) return;
	for (int i = threadIdx.x; i < 72; i += blockDim.x) s_cnt[i] = 0;
	__syncthreads();
	const int W = mb_w * 16, H = mb_h * 16;
	const uint32_t n = job.n_parts;
	uint32_t *items = const_cast<uint32_t *>(job.items);
	for (int pass = 0; pass < 2; ++pass) {
		for (uint32_t r = threadIdx.x; r < n; r += blockDim.x) {
			const b2_h264_mc *mc = job.mcs + r;
			const uint32_t mb = job.owner[r];
			const int pos = mc->pos, size = mc->size, mvx = mc->mv[0][0], mvy = mc->mv[0][1];
			const int x4 = pos & 3, y4 = pos >> 2, pw = size & 15, ph = size >> 4;
			const int fc = frac_class(mvx & 3, mvy & 3);
			const int wide = pw >= 2 && fc < 4, cnt = wide ? pw >> 1 : pw, sw = wide ? 8 : 4;
			const int cls = fc + 6 * wide, hidx = ph == 4 ? 2 : ph - 1;
			const int mbx = mb % mb_w, mby = mb / mb_w;
			const int y0 = mby * 16 + y4 * 4 + (mvy >> 2);
			const bool yedge = y0 - 3 < 0 || y0 + ph * 4 + 4 > H;
			const uint32_t word = mb | ((r - job.mbs[mb].mc_idx) << 20) | ((uint32_t)cls << 26);
			for (int k = 0; k < cnt; ++k) {
				const int x0 = mbx * 16 + x4 * 4 + k * sw + (mvx >> 2);
				const bool edge = yedge || x0 - 6 < 0 || x0 + sw + 12 > W;
				const int b = (edge ? 36 : 0) + cls * 3 + hidx;
				if (pass == 0) atomicAdd(&s_cnt[b], 1u);
				else items[atomicAdd(&s_at[b], 1u)] = word | ((uint32_t)k << 24);
			}
		}
		__syncthreads();
		if (pass == 0) {
			if (threadIdx.x == 0) {
				uint32_t at = 0;
				for (int b = 0; b < 72; ++b) { s_at[b] = at; at += s_cnt[b]; }
				job.n_items = at;
				job.n_first = 0;
			}
			__syncthreads();
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Deblocking strengths on the device (reference: store_info_inter*, str_mv_calc*,
// str_previous_coef, h264.cpp:7119-7322, 7451-7604, 8796-9400, which run inside the parse).  One warp per
// macroblock: the 4x4 motion field of the macroblock, of the column to its left and of the row above it is
// rebuilt from the motion records (lanes 0..15 / 16..19 / 20..23 look their block up), then lane
// (direction, edge, segment) = (l >> 4, (l >> 2) & 3, l & 3) compares the two blocks its segment separates:
//   either block has coefficients                                          -> 2
//   P slice: other picture, or a vector component differs by >= 4          -> 1
//   B slice: other pictures or number of vectors -> 1; the vectors that point to the same pictures compared,
//            where both lists of both blocks name ONE picture "straight" and "crosswise" -- except on the edges the
//            reference compares only straight (b2_h264_dbk.bs_mode: a deviation from 8.7.2.1 that decides pixels)
//   intra macroblock: 3 everywhere, macroblock edges flagged bS 4
// and two ballots spread into the 2-bit fields of the record.  The host supplies what only the parse knows:
// nnz4x4, bs_mode, QPs, offsets and which macroblock edges are filtered at all.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool bs_far(uint32_t a, uint32_t b) {
	const int dx = (int)(int16_t)(a & 0xffff) - (int)(int16_t)(b & 0xffff), dy = (int)(int16_t)(a >> 16) - (int)(int16_t)(b >> 16);
	return (unsigned)(dx + 3) > 6u || (unsigned)(dy + 3) > 6u;
}
// Morton spread: bit i of the low 16 bits -> bit 2 i
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
	x &= 0xffffu;
	x = (x | (x << 8)) & 0x00ff00ffu;
	x = (x | (x << 4)) & 0x0f0f0f0fu;
	x = (x | (x << 2)) & 0x33333333u;
	x = (x | (x << 1)) & 0x55555555u;
	return x;
}
constexpr int BS_WARPS = 8;
struct BsField {            // blocks (x, y), x, y = -1 .. 3, at index (y + 1) * 5 + x + 1
	uint32_t mv[2][25];
	int16_t ref[25];        // list 0 picture in the low byte, list 1 in the high byte (frame-pool indices, -1 = unused)
};

__global__ void __launch_bounds__(32 * BS_WARPS)
h264_bs_kernel(const H4Job *__restrict__ jobs, int mb_w, int n_mbs) {
	__shared__ BsField s_f[BS_WARPS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int mb = blockIdx.x * BS_WARPS + warp;
	if (mb >= n_mbs) return;
	const H4Job job = jobs[blockIdx.y];
	if (!job.derive_bs) return;
	b2_h264_dbk *rec = const_cast<b2_h264_dbk *>(job.dbk) + mb;
	const uint4 d = *reinterpret_cast<const uint4 *>(rec);
	int flags = (d.z >> 24) & 255;
	if (flags & B2_DBK_OFF) return;
	const int mode = (d.w >> 16) & 255;
	const bool left_ok = flags & B2_DBK_LEFT_EDGE, top_ok = flags & B2_DBK_TOP_EDGE;
	const b2_h264_mb *cur = job.mbs + mb;
	const bool t8 = cur->flags & B2_H4_T8x8;
	uint32_t bsv, bsh;
	if (cur->type != B2_H4_INTER) {
		bsv = bsh = 0xffffff00u;
		if (left_ok) { bsv |= 0xffu; flags |= B2_DBK_LEFT_BS4; }
		if (top_ok) { bsh |= 0xffu; flags |= B2_DBK_TOP_BS4; }
	} else {
		const int mbx = mb % mb_w;
		const b2_h264_mb *left = left_ok ? cur - 1 : nullptr, *top = top_ok ? cur - mb_w : nullptr;
		const bool left_inter = left && left->type == B2_H4_INTER, top_inter = top && top->type == B2_H4_INTER;
		(void)mbx;
		BsField &f = s_f[warp];
		// ---- the motion field: every lane looks up the record that covers its block --------------------------
		{
			const b2_h264_mb *src = nullptr;
			int x = 0, y = 0, at = 0;
			if (lane < 16) { src = cur; x = lane & 3; y = lane >> 2; at = (y + 1) * 5 + x + 1; }
			else if (lane < 20) { if (left_inter) src = left; x = 3; y = lane - 16; at = (y + 1) * 5; }
			else if (lane < 24) { if (top_inter) src = top; x = lane - 20; y = 3; at = x + 1; }
			if (src) {
				const b2_h264_mc *mc = job.mcs + src->mc_idx;
				const int n = src->mc_cnt;
				uint32_t m0 = 0, m1 = 0;
				int r01 = -1;
				for (int i = 0; i < n; ++i) {
					const int pos = mc[i].pos, size = mc[i].size;
					const int x4 = pos & 3, y4 = pos >> 2, w4 = size & 15, h4 = size >> 4;
					if (x >= x4 && x < x4 + w4 && y >= y4 && y < y4 + h4) {
						const uint2 v = *reinterpret_cast<const uint2 *>(mc[i].mv);
						m0 = v.x; m1 = v.y;
						r01 = (int)(uint8_t)mc[i].ref_frame[0] | ((int)(uint8_t)mc[i].ref_frame[1] << 8);
						break;
					}
				}
				f.mv[0][at] = m0; f.mv[1][at] = m1; f.ref[at] = (int16_t)r01;
			}
		}
		__syncwarp();
		// ---- one pair of blocks per lane ------------------------------------------------------------------------
		const int dir = lane >> 4, e = (lane >> 2) & 3, sg = lane & 3;
		const int bx = dir ? sg : e, by = dir ? e : sg;                      // the block on the right of / below the edge
		const int ib = (by + 1) * 5 + bx + 1, ia = dir ? ib - 5 : ib - 1;
		const unsigned nzc = cur->nnz4x4;
		bool nzb = (nzc >> (by * 4 + bx)) & 1, nza;
		int bs;
		bool usable = true, other_intra = false;
		if (e == 0) {
			const b2_h264_mb *o = dir ? top : left;
			usable = o != nullptr;
			other_intra = usable && o->type != B2_H4_INTER;
			nza = usable && !other_intra && ((o->nnz4x4 >> (dir ? 12 + sg : sg * 4 + 3)) & 1);
		} else {
			nza = (nzc >> (dir ? (by - 1) * 4 + bx : by * 4 + bx - 1)) & 1;
		}
		if (!usable) bs = 0;
		else if (other_intra) bs = 3;
		else if (nza | nzb) bs = 2;
		else {
			const int ra = f.ref[ia], rb = f.ref[ib];
			const int a0 = (int8_t)(ra & 255), a1 = (int8_t)(ra >> 8), b0 = (int8_t)(rb & 255), b1 = (int8_t)(rb >> 8);
			const uint32_t ma0 = f.mv[0][ia], ma1 = f.mv[1][ia], mb0 = f.mv[0][ib], mb1 = f.mv[1][ib];
			if (!(mode & B2_BS_B_SLICE)) bs = (a0 != b0 || bs_far(ma0, mb0)) ? 1 : 0;
			else if (!((a0 == b0 && a1 == b1) || (a0 == b1 && a1 == b0))) bs = 1;
			else if (a0 != a1) {
				if (a0 == b0) bs = ((a0 >= 0 && bs_far(ma0, mb0)) || (a1 >= 0 && bs_far(ma1, mb1))) ? 1 : 0;
				else bs = ((a0 >= 0 && bs_far(ma0, mb1)) || (a1 >= 0 && bs_far(ma1, mb0))) ? 1 : 0;
			} else {
				const bool straight = bs_far(ma0, mb0) || bs_far(ma1, mb1), cross = bs_far(ma0, mb1) || bs_far(ma1, mb0);
				const bool reduced = mode & (e == 0 ? B2_BS_REDUCED_MB_EDGE : (e == 2 ? B2_BS_REDUCED_CENTRE : B2_BS_REDUCED_INNER));
				bs = reduced ? (straight ? 1 : 0) : ((straight && cross) ? 1 : 0);
			}
		}
		const unsigned lo = __ballot_sync(0xffffffffu, bs & 1), hi = __ballot_sync(0xffffffffu, bs & 2);
		bsv = spread16(lo) | (spread16(hi) << 1);
		bsh = spread16(lo >> 16) | (spread16(hi >> 16) << 1);
		if (left && !left_inter) flags |= B2_DBK_LEFT_BS4;
		if (top && !top_inter) flags |= B2_DBK_TOP_BS4;
	}
	if (t8) {               // transform_size_8x8_flag: the 4x4 edges inside the 8x8 blocks are not filtered (8.7)
		bsv &= 0x00ff00ffu;
		bsh &= 0x00ff00ffu;
	}
	if (lane == 0) *reinterpret_cast<uint4 *>(rec) = make_uint4(bsv, bsh, (d.z & 0x00ffffffu) | ((uint32_t)flags << 24), d.w);
}

}  // namespace

// ticket counters (32 ints) + row progress [n_jobs][mb_h] (wavefronts) or per-macroblock flags
// [n_jobs][mb_w * mb_h] (sparse intra); the caller owns one such area per CUDA stream, since kernels
// of different streams run concurrently
size_t h264_sync_bytes(int n_jobs, int mb_w, int mb_h) { return (32 + (size_t)n_jobs * mb_w * mb_h) * sizeof(int); }

int launch_h264_phase(cudaStream_t stream, int phase, const H4Job *d_jobs, int n_jobs, int mb_w, int mb_h, int *sync_area, uint64_t *launches,
                      uint32_t max_items, uint32_t max_first) {
	if (n_jobs <= 0) return 0;
	if (phase == 5) {
		B2_CUDA_OK(cudaMemsetAsync(sync_area, 0, (32 + (size_t)n_jobs * mb_h) * sizeof(int), stream));
		int ctas = n_jobs * mb_h;
		if (ctas > 148 * 16) ctas = 148 * 16;                       // persistent: every CTA resident
		h264_mv_kernel<<<ctas, 32, 0, stream>>>(d_jobs, n_jobs, mb_w, mb_h, sync_area);
		h264_items_kernel<<<n_jobs, 1024, 0, stream>>>(const_cast<H4Job *>(d_jobs), mb_w, mb_h);
		if (launches) *launches += 2;
		B2_CUDA_OK(cudaGetLastError());
		return 0;
	}
	if (phase == 4) {
		h264_bs_kernel<<<dim3((mb_w * mb_h + BS_WARPS - 1) / BS_WARPS, n_jobs), 32 * BS_WARPS, 0, stream>>>(d_jobs, mb_w, mb_w * mb_h);
		if (launches) ++*launches;
		B2_CUDA_OK(cudaGetLastError());
		return 0;
	}
	if (phase == 0) {
		dim3 grid5((max_items + I5_THREADS - 1) / I5_THREADS, n_jobs);
		constexpr size_t smem5 = sizeof(uint32_t) * 32 * I5_THREADS + sizeof(int16_t) * 66 * I5_THREADS + sizeof(uint16_t) * I5_QUEUE * (I5_THREADS / 32);
		static_assert(smem5 <= 48 * 1024, "above 48 KB the kernels need cudaFuncAttributeMaxDynamicSharedMemorySize, per device");
		if (max_first) {
			dim3 grid_first((max_first + I5_THREADS - 1) / I5_THREADS, n_jobs);
			h264_inter5_kernel<1><<<grid_first, I5_THREADS, smem5, stream>>>(d_jobs);
			if (launches) ++*launches;
			h264_inter5_kernel<2><<<grid5, I5_THREADS, smem5, stream>>>(d_jobs);
		} else {
			h264_inter5_kernel<0><<<grid5, I5_THREADS, smem5, stream>>>(d_jobs);
		}
		if (launches) ++*launches;
	} else {
		int *area = sync_area;
		if (phase == 3) {
			B2_CUDA_OK(cudaMemsetAsync(area, 0, h264_sync_bytes(n_jobs, mb_w, mb_h), stream));
			const int chunks = n_jobs * ((mb_w * mb_h + 31) / 32);
			int ctas = (chunks + WAVE_WARPS - 1) / WAVE_WARPS;
			if (ctas > 148 * WAVE_CTAS_INTRA) ctas = 148 * WAVE_CTAS_INTRA;      // persistent: every CTA resident
			h264_intra_sparse_kernel<<<ctas, 32 * WAVE_WARPS, 0, stream>>>(d_jobs, n_jobs, mb_w, mb_h, area);
			if (launches) ++*launches;
			B2_CUDA_OK(cudaGetLastError());
			return 0;
		}
		B2_CUDA_OK(cudaMemsetAsync(area, 0, (32 + (size_t)n_jobs * mb_h) * sizeof(int), stream));
		int ctas = n_jobs * ((mb_h + WAVE_WARPS - 1) / WAVE_WARPS);
		int cap = 148 * 2 * WAVE_CTAS_DBK;
		if (ctas > cap) ctas = cap;
		if (phase == 1) h264_wave_kernel<1><<<ctas, 32 * WAVE_WARPS, 0, stream>>>(d_jobs, n_jobs, mb_w, mb_h, area);
		else {
			// batched launches pair the pictures up (two pictures and two rows per warp); a single picture keeps the one-picture kernel
			static const bool one_picture = getenv("B2_DBK_ONE_PICTURE") != nullptr;     // (development: the scalar kernel for batched launches too)
			if (!one_picture && n_jobs > 1) {
				int ctas3 = ((n_jobs + 1) / 2) * (((mb_h + 1) / 2 + DBK3_WARPS - 1) / DBK3_WARPS);
				if (ctas3 > 148 * 2 * DBK3_CTAS) ctas3 = 148 * 2 * DBK3_CTAS;
				h264_deblock3_kernel<<<ctas3, 32 * DBK3_WARPS, 0, stream>>>(d_jobs, n_jobs, mb_w, mb_h, area);
			} else {
				h264_wave_kernel<2><<<ctas, 32 * WAVE_WARPS, 0, stream>>>(d_jobs, n_jobs, mb_w, mb_h, area);
			}
		}
		if (launches) ++*launches;
	}
	B2_CUDA_OK(cudaGetLastError());
	return 0;
}

int launch_h264(cudaStream_t stream, const H4Job *d_jobs, int n_jobs, int mb_w, int mb_h, uint32_t max_items, bool intra_dense, bool intra_sparse,
                bool deblock, int *sync_area, uint64_t *launches, uint32_t max_first, bool derive_bs, bool derive_mv) {
	if (derive_mv && max_items && launch_h264_phase(stream, 5, d_jobs, n_jobs, mb_w, mb_h, sync_area, launches) < 0) return -1;
	if (derive_bs && deblock && launch_h264_phase(stream, 4, d_jobs, n_jobs, mb_w, mb_h, sync_area, launches) < 0) return -1;
	if (max_items && launch_h264_phase(stream, 0, d_jobs, n_jobs, mb_w, mb_h, sync_area, launches, max_items, max_first) < 0) return -1;
	if (intra_dense && launch_h264_phase(stream, 1, d_jobs, n_jobs, mb_w, mb_h, sync_area, launches) < 0) return -1;
	if (intra_sparse && launch_h264_phase(stream, 3, d_jobs, n_jobs, mb_w, mb_h, sync_area, launches) < 0) return -1;
	if (deblock && launch_h264_phase(stream, 2, d_jobs, n_jobs, mb_w, mb_h, sync_area, launches) < 0) return -1;
	return 0;
}

}  // namespace b2
