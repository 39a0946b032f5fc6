/* recon_mpeg2.c — ORACLE (test infrastructure, never shipped or called by the product).
 *
 * CPU restatement of the reference's MPEG-2 reconstruction, driven by the same per-picture
 * SoA records (include/b2_recon.h) that the CUDA kernels consume.  Macroblocks are
 * processed in raster order in the reference's own order of operations: forward prediction,
 * backward prediction averaged in, then per coded block row-IDCT -> column-IDCT -> store/add.
 *
 * Follows (reference, /root/reference/src/lib):
 *   idct.cpp:144-235   m2d_idct_horizontal (C path; int16 write-back)
 *   idct.cpp:286-358   m2d_idct_vertical
 *   idct.cpp:364-422   ClipStore / AddStore and the four entry points
 *   motioncomp.cpp:39-43, 214-298, 314-438  AVERAGE2 / bilinear variants
 *   motioncomp.cpp:488-546                  address + half-pel mode selection, chroma mv = mv/2
 *   mpeg2.cpp:1120, 1136-1153, 1277-1308, 1343-1353  call sites (block offsets, field MC)
 *   m2d_macro.h:94-101 CLIP255C (defined for -256..767; we clip arithmetically)
 *
 * Parity pin: tests/test_mpeg2_* compare this file's output, frame by frame, with the MD5
 * lines of the compiled reference (oracle/_ref/h264dec -O) on the same streams.
 */
#include <stdint.h>
#include <string.h>
#include "b2_recon.h"

#define W1 2841
#define W2 2676
#define W3 2408
#define W5 1609
#define W6 1108
#define W7 565

static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* idct.cpp:144-235.  The reference skips rows whose AC part is zero (result is dc*8, which
 * the general path also yields), so the shortcut is not restated. */
static void idct_rows(int16_t *blk)
{
	for (int r = 0; r < 8; ++r) {
		int16_t *s = blk + r * 8;
		int32_t x0, x1, x2, x3, x4, x5, x6, x7, t;
		x0 = s[0] * 2048 + 128;
		x4 = s[1]; x5 = s[7]; x3 = s[2]; x7 = s[3];
		x1 = s[4] * 2048;
		t = x0; x0 = x0 - x1; x1 = x1 + t;
		t = W7 * (x4 + x5);
		x4 = t + (W1 - W7) * x4;
		x5 = t - (W1 + W7) * x5;
		x6 = s[5];
		t = W3 * (x6 + x7);
		x6 = t - (W3 - W5) * x6;
		x7 = t - (W3 + W5) * x7;
		t = x4; x4 -= x6; x6 += t;
		t = x5; x5 -= x7; x7 += t;
		t = x5;
		x5 = ((x4 + x5) * 181 + 128) >> 8;
		x4 = ((x4 - t) * 181 + 128) >> 8;
		x2 = s[6];
		t = W6 * (x3 + x2);
		x2 = t - (W2 + W6) * x2;
		x3 = t + (W2 - W6) * x3;
		t = x0; x0 = x0 - x2; x2 = t + x2;
		t = x1; x1 = x1 - x3; x3 = t + x3;
		s[0] = (int16_t)((x3 + x6) >> 8);
		s[1] = (int16_t)((x2 + x5) >> 8);
		s[2] = (int16_t)((x0 + x4) >> 8);
		s[3] = (int16_t)((x1 + x7) >> 8);
		s[4] = (int16_t)((x1 - x7) >> 8);
		s[5] = (int16_t)((x0 - x4) >> 8);
		s[6] = (int16_t)((x2 - x5) >> 8);
		s[7] = (int16_t)((x3 - x6) >> 8);
	}
}

/* idct.cpp:286-358 with ClipStore (add = 0) or AddStore (add = 1); step = 1 luma, 2 NV12 chroma */
static void idct_cols(uint8_t *dst, int stride, int step, const int16_t *blk, int add)
{
	for (int c = 0; c < 8; ++c) {
		const int16_t *v = blk + c;
		int32_t x0, x1, x2, x3, x4, x5, x6, x7, x8, out[8];
		x7 = v[24]; x6 = v[40];
		x8 = W3 * (x6 + x7) + 4;
		x6 = (x8 - (W3 - W5) * x6) >> 3;
		x7 = (x8 - (W3 + W5) * x7) >> 3;
		x4 = v[8]; x5 = v[56];
		x8 = W7 * (x4 + x5) + 4;
		x4 = (x8 + (W1 - W7) * x4) >> 3;
		x5 = (x8 - (W1 + W7) * x5) >> 3;
		x3 = v[16]; x2 = v[48];
		x1 = W6 * (x3 + x2) + 4;
		x2 = (x1 - (W2 + W6) * x2) >> 3;
		x3 = (x1 + (W2 - W6) * x3) >> 3;
		x1 = x4 + x6; x4 = x4 - x6; x6 = x5 + x7; x5 = x5 - x7;
		x0 = (int32_t)v[0] * 256 + 8192;
		x7 = (int32_t)v[32] * 256;
		x8 = x0 + x7; x0 = x0 - x7;
		x7 = x8 + x3; x8 -= x3;
		x3 = x0 + x2; x0 -= x2;
		x2 = ((x4 + x5) * 181 + 128) >> 8;
		x4 = ((x4 - x5) * 181 + 128) >> 8;
		out[0] = (x7 + x1) >> 14; out[1] = (x3 + x2) >> 14;
		out[2] = (x0 + x4) >> 14; out[3] = (x8 + x6) >> 14;
		out[4] = (x8 - x6) >> 14; out[5] = (x0 - x4) >> 14;
		out[6] = (x3 - x2) >> 14; out[7] = (x7 - x1) >> 14;
		uint8_t *d = dst + c * step;
		for (int r = 0; r < 8; ++r, d += stride)
			*d = (uint8_t)clip255(add ? *d + out[r] : out[r]);
	}
}

/* motioncomp.cpp: one 16-byte-wide block of `height` lines.  mode = (hy<<1)|hx; xstep = 1 for
 * luma, 2 for interleaved chroma; avg = 1 averages into dst (the *_add variants). */
static void predict16(const uint8_t *plane, int W, int rows, int x0, int y0, int ystep, int mode,
                      int xstep, uint8_t *dst, int dst_stride, int height, int avg)
{
	for (int y = 0; y < height; ++y) {
		for (int x = 0; x < 16; ++x) {
			/* the reference does not clamp (hazard 7); clamping only matters for streams whose
			 * vectors leave the picture, which the test streams avoid */
			int xa = x0 + x, xb = xa + xstep, ya = y0 + y * ystep, yb = ya + ystep;
			if (xa < 0) xa = 0; if (xa > W - 1) xa = W - 1;
			if (xb < 0) xb = 0; if (xb > W - 1) xb = W - 1;
			if (ya < 0) ya = 0; if (ya > rows - 1) ya = rows - 1;
			if (yb < 0) yb = 0; if (yb > rows - 1) yb = rows - 1;
			int a = plane[ya * W + xa], v;
			switch (mode) {
			case 0: v = a; break;
			case 1: v = (a + plane[ya * W + xb] + 1) >> 1; break;
			case 2: v = (a + plane[yb * W + xa] + 1) >> 1; break;
			default: v = (a + plane[ya * W + xb] + plane[yb * W + xa] + plane[yb * W + xb] + 2) >> 2; break;
			}
			uint8_t *d = dst + y * dst_stride + x;
			*d = (uint8_t)(avg ? (*d + v + 1) >> 1 : v);
		}
	}
}

/* m2d_motion_comp (mpeg2.cpp:1277-1308) for one direction */
static void motion_comp(const b2_m2_mb *mb, int dir, int avg, int mbx, int mby, int W, int H,
                        const uint8_t *ref_y, const uint8_t *ref_c, uint8_t *dst_y, uint8_t *dst_c)
{
	int nf = (mb->flags & B2_M2_FIELD_MC) ? 2 : 1;
	for (int f = 0; f < nf; ++f) {
		int mvx = mb->mv[dir][f][0], mvy = mb->mv[dir][f][1];
		int sel = (mb->field_sel >> (dir * 2 + f)) & 1;
		int ystep = nf, ly0, cy0, lh = 16 / nf, ch = 8 / nf;
		/* luma */
		ly0 = (nf == 2) ? mby * 16 + sel + 2 * (mvy >> 1) : mby * 16 + (mvy >> 1);
		predict16(ref_y, W, H, mbx * 16 + (mvx >> 1), ly0, ystep, ((mvy & 1) << 1) | (mvx & 1), 1,
		          dst_y + (mby * 16 + f) * W + mbx * 16, W * nf, lh, avg);
		/* chroma: vector halved with truncation toward zero (motioncomp.cpp:511-512) */
		int cx = mvx / 2, cy = mvy / 2;
		cy0 = (nf == 2) ? mby * 8 + sel + 2 * (cy >> 1) : mby * 8 + (cy >> 1);
		predict16(ref_c, W, H / 2, mbx * 16 + (cx & ~1), cy0, ystep, ((cy & 1) << 1) | (cx & 1), 2,
		          dst_c + (mby * 8 + f) * W + mbx * 16, W * nf, ch, avg);
	}
}

int oracle_mpeg2_picture(const b2_m2_picture *p, uint8_t *const *luma, uint8_t *const *chroma)
{
	int W = p->mb_w * 16, H = p->mb_h * 16;
	uint8_t *dy = luma[p->dst], *dc = chroma[p->dst];
	for (int mby = 0; mby < p->mb_h; ++mby) {
		for (int mbx = 0; mbx < p->mb_w; ++mbx) {
			const b2_m2_mb *mb = &p->mbs[mby * p->mb_w + mbx];
			if (mb->flags == 0) continue;
			int intra = mb->flags & B2_M2_INTRA;
			if (!intra) {
				int n = 0;
				if (mb->flags & B2_M2_FWD)
					motion_comp(mb, 0, n++, mbx, mby, W, H, luma[p->fwd], chroma[p->fwd], dy, dc);
				if (mb->flags & B2_M2_BWD)
					motion_comp(mb, 1, n++, mbx, mby, W, H, luma[p->bwd], chroma[p->bwd], dy, dc);
			}
			const int16_t *src = p->coefs + (size_t)mb->coef_idx * 64;
			int field = (mb->flags & B2_M2_FIELD_DCT) != 0;
			for (int k = 0; k < 6; ++k) {
				if (!(mb->cbp & (1 << (5 - k)))) continue;
				int16_t blk[64];
				memcpy(blk, src, sizeof(blk));
				src += 64;
				idct_rows(blk);
				if (k < 4) {
					/* LUMA_BLOCK_OFFSET (mpeg2.cpp:1120) */
					int off = field ? (k & 1) * 8 + ((k & 2) ? W : 0) : (k & 1) * 8 + ((k & 2) ? 8 * W : 0);
					idct_cols(dy + mby * 16 * W + mbx * 16 + off, W << field, 1, blk, !intra);
				} else {
					idct_cols(dc + mby * 8 * W + mbx * 16 + (k - 4), W, 2, blk, !intra);
				}
			}
		}
	}
	return 0;
}
