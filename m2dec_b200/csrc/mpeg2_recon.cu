// mpeg2_recon.cu — K5+K6: MPEG-2 macroblock reconstruction on sm_100a.
//
// One fused kernel per batch of pictures: half-pel motion compensation (forward, backward,
// bidirectional average; frame or 2x field 16x8), the reference's own integer 8x8 IDCT
// (row pass rounded into int16, column pass), residual add / intra store with clip.
// Every output byte of the picture is produced and written exactly once.
//
// Arithmetic restated from reference src/lib/idct.cpp:144-235 (row pass, int16 write-back),
// :286-358 (column pass), :364-422 (ClipStore / AddStore) and src/lib/motioncomp.cpp:39-43
// (AVERAGE2), :314-438 (2x2 bilinear), :488-546 (address and half-pel mode selection);
// call order from src/lib/mpeg2.cpp:1277-1308, 1343-1394.
//
// Mapping: a CTA reconstructs MBS horizontally adjacent macroblocks (rows of 16*MBS
// contiguous bytes); 96 threads per macroblock: 48 do the IDCT passes (block, row/column),
// then each of the 96 produces one 4-byte word of output (64 luma words, 32 NV12 chroma).
// HBM-bound byte/integer work: no tensor cores.
#include "device_common.cuh"

namespace b2 {
namespace {

constexpr int MBS = 4;            // macroblocks per CTA
constexpr int TPM = 96;           // threads per macroblock

#define W1 2841
#define W2 2676
#define W3 2408
#define W5 1609
#define W6 1108
#define W7 565

__device__ __forceinline__ void idct_row(const int16_t *s, int16_t *o) {
	int x0 = s[0] * 2048 + 128, x1 = s[4] * 2048;
	int x4 = s[1], x5 = s[7], x3 = s[2], x7 = s[3], x6 = s[5], x2 = s[6];
	int t = x0;
	x0 = x0 - x1;
	x1 = x1 + t;
	t = W7 * (x4 + x5);
	x4 = t + (W1 - W7) * x4;
	x5 = t - (W1 + W7) * x5;
	t = W3 * (x6 + x7);
	x6 = t - (W3 - W5) * x6;
	x7 = t - (W3 + W5) * x7;
	t = x4; x4 -= x6; x6 += t;
	t = x5; x5 -= x7; x7 += t;
	t = x5;
	x5 = ((x4 + x5) * 181 + 128) >> 8;
	x4 = ((x4 - t) * 181 + 128) >> 8;
	t = W6 * (x3 + x2);
	x2 = t - (W2 + W6) * x2;
	x3 = t + (W2 - W6) * x3;
	t = x0; x0 = x0 - x2; x2 = t + x2;
	t = x1; x1 = x1 - x3; x3 = t + x3;
	o[0] = (int16_t)((x3 + x6) >> 8);
	o[1] = (int16_t)((x2 + x5) >> 8);
	o[2] = (int16_t)((x0 + x4) >> 8);
	o[3] = (int16_t)((x1 + x7) >> 8);
	o[4] = (int16_t)((x1 - x7) >> 8);
	o[5] = (int16_t)((x0 - x4) >> 8);
	o[6] = (int16_t)((x2 - x5) >> 8);
	o[7] = (int16_t)((x3 - x6) >> 8);
}

// column pass on v[r*8 + c] for one column c; writes the 8 (unclipped) sample values
__device__ __forceinline__ void idct_col(const int16_t *v, int16_t *o) {
	int x7 = v[8 * 3], x6 = v[8 * 5], x8;
	x8 = W3 * (x6 + x7) + 4;
	x6 = (x8 - (W3 - W5) * x6) >> 3;
	x7 = (x8 - (W3 + W5) * x7) >> 3;
	int x4 = v[8 * 1], x5 = v[8 * 7];
	x8 = W7 * (x4 + x5) + 4;
	x4 = (x8 + (W1 - W7) * x4) >> 3;
	x5 = (x8 - (W1 + W7) * x5) >> 3;
	int x3 = v[8 * 2], x2 = v[8 * 6];
	int x1 = W6 * (x3 + x2) + 4;
	x2 = (x1 - (W2 + W6) * x2) >> 3;
	x3 = (x1 + (W2 - W6) * x3) >> 3;
	x1 = x4 + x6;
	x4 = x4 - x6;
	x6 = x5 + x7;
	x5 = x5 - x7;
	int x0 = (int)v[0] * 256 + 8192;
	x7 = (int)v[8 * 4] * 256;
	x8 = x0 + x7;
	x0 = x0 - x7;
	x7 = x8 + x3;
	x8 -= x3;
	x3 = x0 + x2;
	x0 -= x2;
	x2 = ((x4 + x5) * 181 + 128) >> 8;
	x4 = ((x4 - x5) * 181 + 128) >> 8;
	o[8 * 0] = (int16_t)((x7 + x1) >> 14);
	o[8 * 1] = (int16_t)((x3 + x2) >> 14);
	o[8 * 2] = (int16_t)((x0 + x4) >> 14);
	o[8 * 3] = (int16_t)((x8 + x6) >> 14);
	o[8 * 4] = (int16_t)((x8 - x6) >> 14);
	o[8 * 5] = (int16_t)((x0 - x4) >> 14);
	o[8 * 6] = (int16_t)((x3 - x2) >> 14);
	o[8 * 7] = (int16_t)((x7 - x1) >> 14);
}

// Half-pel sample fetch. (x, y) are integer sample coordinates in the plane; xstep is the
// distance to the horizontal neighbour (1 luma, 2 NV12 chroma); ystride the distance, in
// rows, to the vertical neighbour (1 frame, 2 field).  The reference does not clamp motion
// vectors (SURVEY.md hazard 7: out-of-frame vectors read foreign memory); we clamp the
// coordinates to the plane so that such streams cannot fault the GPU.
__device__ __forceinline__ int fetch_hp(const uint8_t *__restrict__ plane, int stride, int rows,
                                        int x, int y, int hx, int hy, int xstep, int ystep) {
	int xmax = stride - 1, ymax = rows - 1;
	int x0 = min(max(x, 0), xmax), x1 = min(max(x + xstep, 0), xmax);
	int y0 = min(max(y, 0), ymax), y1 = min(max(y + ystep, 0), ymax);
	int a = __ldg(plane + (size_t)y0 * stride + x0);
	if (!hx && !hy) return a;
	if (hx && !hy) return (a + __ldg(plane + (size_t)y0 * stride + x1) + 1) >> 1;
	if (!hx) return (a + __ldg(plane + (size_t)y1 * stride + x0) + 1) >> 1;
	int b = __ldg(plane + (size_t)y0 * stride + x1);
	int c = __ldg(plane + (size_t)y1 * stride + x0);
	int d = __ldg(plane + (size_t)y1 * stride + x1);
	return (a + b + c + d + 2) >> 2;
}

// Four horizontally adjacent half-pel samples as one packed word.  Fast path: the row segment is read
// with aligned 32-bit loads and funnel shifts, the averages are done on packed bytes (`__vavgu4` is
// exactly AVERAGE2, (a + b + 1) >> 1) or, for the 2x2 case, in two sets of 16-bit lanes.  Positions
// whose words would leave the plane take the clamped byte path.
__device__ __forceinline__ uint32_t fetch_hp4(const uint8_t *__restrict__ plane, int stride, int rows,
                                              int x, int y, int hx, int hy, int xstep, int ystep) {
	const int y1 = y + (hy ? ystep : 0);
	const bool fast = x >= 0 && (x & ~3) + 12 <= stride && y >= 0 && y1 < rows;
	if (!fast) {
		uint32_t w = 0;
#pragma unroll
		for (int i = 0; i < 4; ++i) w |= (uint32_t)fetch_hp(plane, stride, rows, x + i, y, hx, hy, xstep, ystep) << (8 * i);
		return w;
	}
	const uint8_t *p0 = plane + (size_t)y * stride + x;
	const uint32_t *a0 = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p0) & ~(uintptr_t)3);
	const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(p0) & 3) * 8;
	const unsigned nsh = 8u * (unsigned)xstep;            // horizontal neighbour: 1 byte (luma) or 2 (NV12 chroma)
	uint32_t w0 = __ldg(a0), w1 = __ldg(a0 + 1), w2 = __ldg(a0 + 2);
	const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
	const uint32_t a = lo, b = __funnelshift_r(lo, hi, nsh);
	if (!hy) return hx ? __vavgu4(a, b) : a;
	const uint32_t *a1 = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(a0) + (size_t)ystep * stride);
	w0 = __ldg(a1); w1 = __ldg(a1 + 1); w2 = __ldg(a1 + 2);
	const uint32_t lo1 = __funnelshift_r(w0, w1, sh), hi1 = __funnelshift_r(w1, w2, sh);
	const uint32_t c = lo1, d = __funnelshift_r(lo1, hi1, nsh);
	if (!hx) return __vavgu4(a, c);
	const uint32_t m = 0x00FF00FFu;
	const uint32_t e = (a & m) + (b & m) + (c & m) + (d & m) + 0x00020002u;
	const uint32_t o = ((a >> 8) & m) + ((b >> 8) & m) + ((c >> 8) & m) + ((d >> 8) & m) + 0x00020002u;
	return ((e >> 2) & m) | (((o >> 2) & m) << 8);
}

__global__ void __launch_bounds__(MBS * TPM)
mpeg2_recon_kernel(const M2Job *__restrict__ jobs) {
	__shared__ __align__(16) int16_t s_tmp[MBS][6][64];
	__shared__ __align__(16) int16_t s_res[MBS][6][64];

	const M2Job job = jobs[blockIdx.z];
	const int lm = threadIdx.x / TPM;        // macroblock slot inside the CTA
	const int t = threadIdx.x % TPM;
	const int mbx = blockIdx.x * MBS + lm;
	const int mby = blockIdx.y;
	const bool live = mbx < job.mb_w;
	b2_m2_mb mb;
	mb.flags = 0;
	mb.cbp = 0;
	if (live) mb = job.mbs[(size_t)mby * job.mb_w + mbx];

	// ---- IDCT: rows, then columns -------------------------------------------------------
	const int blk = t >> 3, lane8 = t & 7;
	bool coded = false;
	int rec = 0;
	if (t < 48) {
		coded = (mb.cbp >> (5 - blk)) & 1;
		rec = __popc((unsigned)(mb.cbp >> (6 - blk)));     // coded blocks before this one
	}
	if (coded) {
		const int16_t *src = job.coefs + ((size_t)mb.coef_idx + rec) * 64 + lane8 * 8;
		int4 raw = __ldg(reinterpret_cast<const int4 *>(src));
		int16_t s[8];
		*reinterpret_cast<int4 *>(s) = raw;
		idct_row(s, &s_tmp[lm][blk][lane8 * 8]);
	}
	__syncthreads();
	if (coded) idct_col(&s_tmp[lm][blk][lane8], &s_res[lm][blk][lane8]);
	__syncthreads();
	if (!live || mb.flags == 0) return;

	// ---- prediction + residual: one 4-byte word per thread ------------------------------
	const int W = job.mb_w * 16;
	const int H = job.mb_h * 16;
	const bool is_chroma = t >= 64;
	const int w = is_chroma ? t - 64 : t;
	const int row = w >> 2;                  // 0..15 luma, 0..7 chroma
	const int xb = (w & 3) * 4;              // byte offset inside the 16-byte row
	const bool intra = mb.flags & B2_M2_INTRA;
	const bool field_mc = mb.flags & B2_M2_FIELD_MC;
	const bool field_dct = mb.flags & B2_M2_FIELD_DCT;

	int px[4];
	if (intra) {
		px[0] = px[1] = px[2] = px[3] = 0;
	} else {
		uint32_t acc = 0;
		int ndir = 0;
#pragma unroll
		for (int dir = 0; dir < 2; ++dir) {
			if (!(mb.flags & (dir ? B2_M2_BWD : B2_M2_FWD))) continue;
			const uint8_t *ref = is_chroma ? (dir ? job.bwd_c : job.fwd_c) : (dir ? job.bwd_y : job.fwd_y);
			const int f = field_mc ? (row & 1) : 0;
			int mvx = mb.mv[dir][f][0], mvy = mb.mv[dir][f][1];
			int xstep = 1;
			if (is_chroma) { mvx /= 2; mvy /= 2; xstep = 2; }
			const int hx = mvx & 1, hy = mvy & 1;
			int sx = mbx * 16 + xb + (is_chroma ? ((mvx >> 1) * 2) : (mvx >> 1));
			int sy, ystep;
			const int rows_plane = is_chroma ? H / 2 : H;
			const int mb_row0 = mby * (is_chroma ? 8 : 16);
			if (field_mc) {
				const int sel = (mb.field_sel >> (dir * 2 + f)) & 1;
				sy = mb_row0 + sel + 2 * ((row >> 1) + (mvy >> 1));
				ystep = 2;
			} else {
				sy = mb_row0 + row + (mvy >> 1);
				ystep = 1;
			}
			// NV12: bytes alternate Cb/Cr, so a byte's horizontal neighbour is 2 bytes away
			const uint32_t v = fetch_hp4(ref, W, rows_plane, sx, sy, hx, hy, xstep, ystep);
			acc = ndir ? __vavgu4(acc, v) : v;
			ndir++;
		}
#pragma unroll
		for (int i = 0; i < 4; ++i) px[i] = (acc >> (8 * i)) & 255;
	}

	// residual
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		int x = xb + i;
		int b, r, c;
		if (is_chroma) {
			b = 4 + (x & 1);
			r = row;
			c = x >> 1;
		} else if (field_dct) {
			b = (row & 1) * 2 + (x >> 3);
			r = row >> 1;
			c = x & 7;
		} else {
			b = (row >> 3) * 2 + (x >> 3);
			r = row & 7;
			c = x & 7;
		}
		if ((mb.cbp >> (5 - b)) & 1) px[i] = clip255(px[i] + s_res[lm][b][r * 8 + c]);
	}
	uint32_t word = (uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16) | ((uint32_t)px[3] << 24);
	uint8_t *dst = is_chroma ? job.dst_c + (size_t)(mby * 8 + row) * W : job.dst_y + (size_t)(mby * 16 + row) * W;
	*reinterpret_cast<uint32_t *>(dst + mbx * 16 + xb) = word;
}

}  // namespace

int launch_mpeg2(cudaStream_t stream, const M2Job *d_jobs, int n_jobs, int mb_w, int mb_h) {
	if (n_jobs <= 0) return 0;
	dim3 grid((mb_w + MBS - 1) / MBS, mb_h, n_jobs);
	mpeg2_recon_kernel<<<grid, MBS * TPM, 0, stream>>>(d_jobs);
	B2_CUDA_OK(cudaGetLastError());
	return 0;
}

}  // namespace b2
