/* recon_h264.c — ORACLE (test infrastructure, never shipped or called by the product).
 *
 * CPU restatement of the reference's H.264 picture reconstruction, driven by the same
 * per-picture SoA records (include/b2_recon.h) that the CUDA kernels consume.  Macroblocks
 * are reconstructed in raster order exactly in the reference's order of operations
 * (prediction, then residual, block by block for Intra4x4), then the whole picture is
 * deblocked MB by MB in raster order (left edge, inner vertical edges, top edge, inner
 * horizontal edges), as deblock_pb does.
 *
 * Follows (reference, /root/reference/src/lib/h264.cpp unless noted):
 *   964-1054   qp_matrix (flat dequantisation tables)          2005-2022  coeff_writeback
 *   2113-2269  4x4 inverse transform + add (SSE2 path = exact clip), chroma variant, DC-only
 *   4309-4365  intra16x16_dc_transform ((x+2)>>2)              4387-4404  intra_chroma_dc_transform (>>1)
 *   2510-2997  Intra4x4 predictors                              3042, 2537, 2510, 4224-4304  Intra16x16
 *   4561-4705  chroma predictors (NV12)                         4736-4761  I_PCM
 *   3121-3254, 4407-4552, 2374-2461  per-MB drivers (block order, DC-only vs AC paths)
 *   5381-6406  luma 6-tap quarter-pel (incl. UMV edge replication 5932-6116)
 *   4859-5330  chroma eighth-pel bilinear
 *   6726-6749  inter_pred_basic
 *   10253-10663 deblock_pb and the edge filters; tables h264vld.h:4536-4627
 *   m2d.h:286-341 DC-only packed add (equals a clipped add for |dc| <= 255, the only range
 *   the test streams use — SURVEY.md hazard 2)
 *
 * Parity pin: tests/test_h264_* compare this file's output, frame by frame, with the MD5
 * lines of the compiled reference (oracle/_ref/h264dec -O) on writer-generated streams.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "b2_recon.h"

static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

static const uint8_t blk_x[16] = {0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3};
static const uint8_t blk_y[16] = {0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3};

/* ---- dequantisation (qp_matrix, h264.cpp:964-995) ------------------------------------------ */
static int level_scale4(int qp, int pos)
{
	static const int8_t na[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};
	int x = pos & 3, y = pos >> 2;
	int cls = (!(x & 1) && !(y & 1)) ? 0 : ((x & 1) && (y & 1)) ? 1 : 2;
	return na[qp % 6][cls] << (qp / 6);
}

/* 4x4 inverse transform of dequantised coefficients c[4*y+x], +32 on DC, >>6 (h264.cpp:2145-2197) */
static void idct4x4(const int *c, int *res)
{
	int t[16];
	for (int y = 0; y < 4; ++y) {          /* horizontal first, as the reference and 8.5.12.2 */
		int d0 = c[4 * y], d1 = c[4 * y + 1], d2 = c[4 * y + 2], d3 = c[4 * y + 3];
		int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
		t[4 * y] = e0 + e3; t[4 * y + 1] = e1 + e2; t[4 * y + 2] = e1 - e2; t[4 * y + 3] = e0 - e3;
	}
	for (int x = 0; x < 4; ++x) {
		int d0 = t[x], d1 = t[4 + x], d2 = t[8 + x], d3 = t[12 + x];
		int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
		res[x] = (e0 + e3) >> 6; res[4 + x] = (e1 + e2) >> 6; res[8 + x] = (e1 - e2) >> 6; res[12 + x] = (e0 - e3) >> 6;
	}
}

/* residual add of one 4x4 block; step = 1 luma, 2 NV12 chroma.  lv = levels (raster) or NULL,
 * dc_override: use *dc (already dequantised/transformed DC) in place of the level at position 0 */
static void add_block4x4(uint8_t *dst, int stride, int step, const int16_t *lv, int qp, const int *dc)
{
	int c[16], res[16];
	for (int i = 0; i < 16; ++i) c[i] = lv ? lv[i] * level_scale4(qp, i) : 0;
	if (dc) c[0] = *dc;
	c[0] += 32;
	idct4x4(c, res);
	for (int y = 0; y < 4; ++y)
		for (int x = 0; x < 4; ++x) {
			uint8_t *p = dst + y * stride + x * step;
			*p = (uint8_t)clip255(*p + res[4 * y + x]);
		}
}

/* 8x8 dequantisation (qp_matrix8x8, h264.cpp:997-1054): normAdjust8x8 << (qp/6 - 2); for qp < 12
 * the reference shifts the table entry RIGHT by (2 - qp/6), truncating it (a deviation from 8.5.9
 * that is part of the reference's observable behaviour and therefore restated). */
static int level_scale8(int qp, int pos)
{
	static const int8_t na[6][6] = {{20, 18, 32, 19, 25, 24}, {22, 19, 35, 21, 28, 26}, {26, 23, 42, 24, 33, 31},
	                                {28, 25, 45, 26, 35, 33}, {32, 28, 51, 30, 40, 38}, {36, 32, 58, 34, 46, 43}};
	int x = pos & 3, y = (pos >> 3) & 3, cls;   /* the class pattern repeats every 4 columns / rows */
	if (x == 0 && y == 0) cls = 0;
	else if ((x & 1) && (y & 1)) cls = 1;
	else if (x == 2 && y == 2) cls = 2;
	else if ((x == 0 && (y & 1)) || ((x & 1) && y == 0)) cls = 3;
	else if ((x == 0 && y == 2) || (x == 2 && y == 0)) cls = 4;
	else cls = 5;
	int shift = qp / 6 - 2;
	return shift >= 0 ? na[qp % 6][cls] << shift : na[qp % 6][cls] >> -shift;
}

/* one dimension of the 8x8 inverse transform (8.5.13; ac8x8transform_interim, h264.cpp:3942-3983) */
static void idct8_1d(const int *x, int step, int *o)
{
	int x0 = x[0], x1 = x[step], x2 = x[2 * step], x3 = x[3 * step], x4 = x[4 * step], x5 = x[5 * step], x6 = x[6 * step], x7 = x[7 * step];
	int a0 = x0 + x4, a4 = x0 - x4, a2 = (x2 >> 1) - x6, a6 = x2 + (x6 >> 1);
	int b0 = a0 + a6, b2 = a4 + a2, b4 = a4 - a2, b6 = a0 - a6;
	int a1 = x5 - x3 - x7 - (x7 >> 1), a3 = x1 + x7 - x3 - (x3 >> 1), a5 = x5 + (x5 >> 1) + x7 - x1, a7 = x3 + x5 + x1 + (x1 >> 1);
	int b1 = a1 + (a7 >> 2), b7 = a7 - (a1 >> 2), b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5;
	o[0] = b0 + b7; o[1] = b2 + b5; o[2] = b4 + b3; o[3] = b6 + b1;
	o[4] = b6 - b1; o[5] = b4 - b3; o[6] = b2 - b5; o[7] = b0 - b7;
}

/* residual add of one 8x8 luma block: lv = 64 levels, raster (ac8x8transform, h264.cpp:3985-4080) */
static void add_block8x8(uint8_t *dst, int stride, const int16_t *lv, int qp)
{
	int c[64], t[64], o[8];
	for (int i = 0; i < 64; ++i) c[i] = lv[i] * level_scale8(qp, i);
	c[0] += 32;
	for (int y = 0; y < 8; ++y) {          /* rows first */
		idct8_1d(c + 8 * y, 1, o);
		for (int x = 0; x < 8; ++x) t[8 * y + x] = o[x];
	}
	for (int x = 0; x < 8; ++x) {
		idct8_1d(t + x, 8, o);
		for (int y = 0; y < 8; ++y) {
			uint8_t *p = dst + y * stride + x;
			*p = (uint8_t)clip255(*p + (o[y] >> 6));
		}
	}
}

/* ---- intra prediction -------------------------------------------------------------------------- */
/* Intra4x4 (8.3.1.2; h264.cpp:2510-2985).  avail: bit0 left, bit1 top, bit2 top-right */
static void pred4x4(uint8_t *d, int s, int mode, int avail)
{
	int t[9], l[5];   /* t[0] = p[-1,-1], t[1..8] = p[0..7,-1]; l[0] = p[-1,-1], l[1..4] = p[-1,0..3] */
	uint8_t pr[16];
	for (int i = 0; i < 9; ++i) t[i] = 128;
	for (int i = 0; i < 5; ++i) l[i] = 128;
	if (avail & 2) {
		for (int i = 0; i < 4; ++i) t[1 + i] = d[-s + i];
		for (int i = 4; i < 8; ++i) t[1 + i] = (avail & 4) ? d[-s + i] : d[-s + 3];
	}
	if (avail & 1) for (int i = 0; i < 4; ++i) l[1 + i] = d[i * s - 1];
	if ((avail & 3) == 3) t[0] = l[0] = d[-s - 1];
#define T(i) t[(i) + 1]
#define L(i) l[(i) + 1]
	for (int y = 0; y < 4; ++y)
		for (int x = 0; x < 4; ++x) {
			int v;
			switch (mode) {
			case 0: v = T(x); break;
			case 1: v = L(y); break;
			case 2:
				if ((avail & 3) == 3) v = (T(0) + T(1) + T(2) + T(3) + L(0) + L(1) + L(2) + L(3) + 4) >> 3;
				else if (avail & 1) v = (L(0) + L(1) + L(2) + L(3) + 2) >> 2;
				else if (avail & 2) v = (T(0) + T(1) + T(2) + T(3) + 2) >> 2;
				else v = 128;
				break;
			case 3:
				v = (x == 3 && y == 3) ? (T(6) + 3 * T(7) + 2) >> 2 : (T(x + y) + 2 * T(x + y + 1) + T(x + y + 2) + 2) >> 2;
				break;
			case 4:
				if (x > y) v = (T(x - y - 2) + 2 * T(x - y - 1) + T(x - y) + 2) >> 2;
				else if (x < y) v = (L(y - x - 2) + 2 * L(y - x - 1) + L(y - x) + 2) >> 2;
				else v = (T(0) + 2 * T(-1) + L(0) + 2) >> 2;
				break;
			case 5: {
				int z = 2 * x - y;
				if (z >= 0 && !(z & 1)) v = (T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 1) >> 1;
				else if (z >= 0) v = (T(x - (y >> 1) - 2) + 2 * T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 2) >> 2;
				else if (z == -1) v = (L(0) + 2 * L(-1) + T(0) + 2) >> 2;
				else v = (L(y - 1) + 2 * L(y - 2) + L(y - 3) + 2) >> 2;
				break;
			}
			case 6: {
				int z = 2 * y - x;
				if (z >= 0 && !(z & 1)) v = (L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 1) >> 1;
				else if (z >= 0) v = (L(y - (x >> 1) - 2) + 2 * L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 2) >> 2;
				else if (z == -1) v = (L(0) + 2 * L(-1) + T(0) + 2) >> 2;
				else v = (T(x - 1) + 2 * T(x - 2) + T(x - 3) + 2) >> 2;
				break;
			}
			case 7:
				if (!(y & 1)) v = (T(x + (y >> 1)) + T(x + (y >> 1) + 1) + 1) >> 1;
				else v = (T(x + (y >> 1)) + 2 * T(x + (y >> 1) + 1) + T(x + (y >> 1) + 2) + 2) >> 2;
				break;
			default: {
				int z = x + 2 * y;
				if (z > 5) v = L(3);
				else if (z == 5) v = (L(2) + 3 * L(3) + 2) >> 2;
				else if (!(z & 1)) v = (L(y + (x >> 1)) + L(y + (x >> 1) + 1) + 1) >> 1;
				else v = (L(y + (x >> 1)) + 2 * L(y + (x >> 1) + 1) + L(y + (x >> 1) + 2) + 2) >> 2;
				break;
			}
			}
			pr[4 * y + x] = (uint8_t)v;
		}
#undef T
#undef L
	for (int y = 0; y < 4; ++y) memcpy(d + y * s, pr + 4 * y, 4);
}

/* per-block availability from the MB-level mask (h264.cpp:3132-3225) */
static int blk_avail(int k, int a)
{
	int x = blk_x[k], y = blk_y[k];
	int left = x > 0 || (a & 1), top = y > 0 || (a & 2), tr;
	if (y == 0) tr = x < 3 ? (a & 2) != 0 : (a & 4) != 0;
	else if (x == 3) tr = 0;
	else {
		static const uint8_t r2b[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};
		tr = r2b[(y - 1) * 4 + x + 1] < k;
	}
	return (left ? 1 : 0) | (top ? 2 : 0) | (tr ? 4 : 0);
}

/* Intra8x8 (8.3.2.2; h264.cpp:3315-3929).  avail: bit0 left, bit1 top, bit2 top-right, bit3
 * top-left, exactly as the reference's predictors receive it.  Reference samples are filtered
 * first (8.3.2.2.1); t[1 + x] = p'[x, -1] (x = 0..15), l[1 + y] = p'[-1, y], t[0] = l[0] = p'[-1, -1]. */
static void pred8x8l(uint8_t *d, int s, int mode, int avail)
{
	int L = avail & 1, T = avail & 2, TR = avail & 4, TL = avail & 8;
	int pt[16], pl[8], ptl = TL ? d[-s - 1] : 0;
	int t[17], l[9];
	memset(t, 0, sizeof(t));
	memset(l, 0, sizeof(l));
	if (T) {
		for (int x = 0; x < 8; ++x) pt[x] = d[-s + x];
		for (int x = 8; x < 16; ++x) pt[x] = TR ? d[-s + x] : pt[7];
		t[1] = TL ? (ptl + 2 * pt[0] + pt[1] + 2) >> 2 : (3 * pt[0] + pt[1] + 2) >> 2;
		for (int x = 1; x < 15; ++x) t[1 + x] = (pt[x - 1] + 2 * pt[x] + pt[x + 1] + 2) >> 2;
		t[16] = (pt[14] + 3 * pt[15] + 2) >> 2;
	}
	if (L) {
		for (int y = 0; y < 8; ++y) pl[y] = d[y * s - 1];
		l[1] = TL ? (ptl + 2 * pl[0] + pl[1] + 2) >> 2 : (3 * pl[0] + pl[1] + 2) >> 2;
		for (int y = 1; y < 7; ++y) l[1 + y] = (pl[y - 1] + 2 * pl[y] + pl[y + 1] + 2) >> 2;
		l[8] = (pl[6] + 3 * pl[7] + 2) >> 2;
	}
	if (TL) {
		int v;
		if (T && L) v = (pt[0] + 2 * ptl + pl[0] + 2) >> 2;
		else if (T) v = (3 * ptl + pt[0] + 2) >> 2;
		else if (L) v = (3 * ptl + pl[0] + 2) >> 2;
		else v = ptl;
		t[0] = l[0] = v;
	}
#define PT(x) t[1 + (x)]
#define PL(y) l[1 + (y)]
	for (int y = 0; y < 8; ++y)
		for (int x = 0; x < 8; ++x) {
			int v;
			switch (mode) {
			case 0: v = PT(x); break;
			case 1: v = PL(y); break;
			case 2: {
				int st = 0, sl = 0;
				for (int i = 0; i < 8; ++i) { st += PT(i); sl += PL(i); }
				v = (T && L) ? (st + sl + 8) >> 4 : (L ? (sl + 4) >> 3 : (T ? (st + 4) >> 3 : 128));
				break;
			}
			case 3: v = (x == 7 && y == 7) ? (PT(14) + 3 * PT(15) + 2) >> 2 : (PT(x + y) + 2 * PT(x + y + 1) + PT(x + y + 2) + 2) >> 2; break;
			case 4:
				if (x > y) v = (PT(x - y - 2) + 2 * PT(x - y - 1) + PT(x - y) + 2) >> 2;
				else if (x < y) v = (PL(y - x - 2) + 2 * PL(y - x - 1) + PL(y - x) + 2) >> 2;
				else v = (PT(0) + 2 * PT(-1) + PL(0) + 2) >> 2;
				break;
			case 5: {
				int z = 2 * x - y;
				if (z >= 0 && !(z & 1)) v = (PT(x - (y >> 1) - 1) + PT(x - (y >> 1)) + 1) >> 1;
				else if (z >= 0) v = (PT(x - (y >> 1) - 2) + 2 * PT(x - (y >> 1) - 1) + PT(x - (y >> 1)) + 2) >> 2;
				else if (z == -1) v = (PL(0) + 2 * PT(-1) + PT(0) + 2) >> 2;
				else v = (PL(y - 2 * x - 1) + 2 * PL(y - 2 * x - 2) + PL(y - 2 * x - 3) + 2) >> 2;
				break;
			}
			case 6: {
				int z = 2 * y - x;
				if (z >= 0 && !(z & 1)) v = (PL(y - (x >> 1) - 1) + PL(y - (x >> 1)) + 1) >> 1;
				else if (z >= 0) v = (PL(y - (x >> 1) - 2) + 2 * PL(y - (x >> 1) - 1) + PL(y - (x >> 1)) + 2) >> 2;
				else if (z == -1) v = (PL(0) + 2 * PT(-1) + PT(0) + 2) >> 2;
				else v = (PT(x - 2 * y - 1) + 2 * PT(x - 2 * y - 2) + PT(x - 2 * y - 3) + 2) >> 2;
				break;
			}
			case 7:
				v = !(y & 1) ? (PT(x + (y >> 1)) + PT(x + (y >> 1) + 1) + 1) >> 1
				             : (PT(x + (y >> 1)) + 2 * PT(x + (y >> 1) + 1) + PT(x + (y >> 1) + 2) + 2) >> 2;
				break;
			default: {
				int z = x + 2 * y;
				if (z > 13) v = PL(7);
				else if (z == 13) v = (PL(6) + 3 * PL(7) + 2) >> 2;
				else if (!(z & 1)) v = (PL(y + (x >> 1)) + PL(y + (x >> 1) + 1) + 1) >> 1;
				else v = (PL(y + (x >> 1)) + 2 * PL(y + (x >> 1) + 1) + PL(y + (x >> 1) + 2) + 2) >> 2;
				break;
			}
			}
			d[y * s + x] = (uint8_t)v;
		}
#undef PT
#undef PL
}

/* availability of 8x8 block q from the MB-level mask (luma_intra8x8_with_residual, h264.cpp:4093-4118) */
static int blk8_avail(int q, int a)
{
	switch (q) {
	case 0: return (a & ~4) | ((a & 2) * 2);
	case 1: return (a & ~8) | ((a & 2) * 4) | 1;
	case 2: return 6 | ((a & 1) * 9);
	default: return 11;
	}
}

/* Intra16x16 (8.3.3) */
static void pred16x16(uint8_t *d, int s, int mode, int avail)
{
	if (mode == 0) {
		for (int y = 0; y < 16; ++y) memcpy(d + y * s, d - s, 16);
	} else if (mode == 1) {
		for (int y = 0; y < 16; ++y) memset(d + y * s, d[y * s - 1], 16);
	} else if (mode == 2) {
		int sum = 0, v;
		if (avail & 1) for (int i = 0; i < 16; ++i) sum += d[i * s - 1];
		if (avail & 2) for (int i = 0; i < 16; ++i) sum += d[-s + i];
		if ((avail & 3) == 3) v = (sum + 16) >> 5;
		else if (avail & 3) v = (sum + 8) >> 4;
		else v = 128;
		for (int y = 0; y < 16; ++y) memset(d + y * s, v, 16);
	} else {
		int H = 0, V = 0;
		for (int i = 0; i < 8; ++i) {
			H += (i + 1) * (d[-s + 8 + i] - d[-s + 6 - i]);
			V += (i + 1) * (d[(8 + i) * s - 1] - d[(6 - i) * s - 1]);
		}
		int a = 16 * (d[15 * s - 1] + d[-s + 15]), b = (5 * H + 32) >> 6, c = (5 * V + 32) >> 6;
		uint8_t tmp[256];
		for (int y = 0; y < 16; ++y)
			for (int x = 0; x < 16; ++x) tmp[16 * y + x] = (uint8_t)clip255((a + b * (x - 7) + c * (y - 7) + 16) >> 5);
		for (int y = 0; y < 16; ++y) memcpy(d + y * s, tmp + 16 * y, 16);
	}
}

/* chroma 8x8 of one component inside the NV12 plane (8.3.4); d points at the component's first byte */
static void pred_chroma(uint8_t *d, int s, int mode, int avail)
{
	uint8_t tmp[64];
	if (mode == 0) {
		for (int by = 0; by < 2; ++by)
			for (int bx = 0; bx < 2; ++bx) {
				int st = 0, sl = 0, v;
				for (int i = 0; i < 4; ++i) {
					if (avail & 2) st += d[-s + (bx * 4 + i) * 2];
					if (avail & 1) sl += d[(by * 4 + i) * s - 2];
				}
				int has_t = (avail & 2) != 0, has_l = (avail & 1) != 0;
				if (bx == 0 && by == 0) v = (has_t && has_l) ? (st + sl + 4) >> 3 : has_l ? (sl + 2) >> 2 : has_t ? (st + 2) >> 2 : 128;
				else if (bx == 1 && by == 0) v = has_t ? (st + 2) >> 2 : has_l ? (sl + 2) >> 2 : 128;
				else if (bx == 0 && by == 1) v = has_l ? (sl + 2) >> 2 : has_t ? (st + 2) >> 2 : 128;
				else v = (has_t && has_l) ? (st + sl + 4) >> 3 : has_l ? (sl + 2) >> 2 : has_t ? (st + 2) >> 2 : 128;
				for (int y = 0; y < 4; ++y)
					for (int x = 0; x < 4; ++x) tmp[(by * 4 + y) * 8 + bx * 4 + x] = (uint8_t)v;
			}
	} else if (mode == 1) {
		for (int y = 0; y < 8; ++y) memset(tmp + 8 * y, d[y * s - 2], 8);
	} else if (mode == 2) {
		for (int y = 0; y < 8; ++y)
			for (int x = 0; x < 8; ++x) tmp[8 * y + x] = d[-s + 2 * x];
	} else {
		int H = 0, V = 0;
		for (int i = 0; i < 4; ++i) {
			H += (i + 1) * (d[-s + (4 + i) * 2] - d[-s + (2 - i) * 2]);
			V += (i + 1) * (d[(4 + i) * s - 2] - d[(2 - i) * s - 2]);
		}
		int a = 16 * (d[7 * s - 2] + d[-s + 14]), b = (34 * H + 32) >> 6, c = (34 * V + 32) >> 6;
		for (int y = 0; y < 8; ++y)
			for (int x = 0; x < 8; ++x) tmp[8 * y + x] = (uint8_t)clip255((a + b * (x - 3) + c * (y - 3) + 16) >> 5);
	}
	for (int y = 0; y < 8; ++y)
		for (int x = 0; x < 8; ++x) d[y * s + 2 * x] = tmp[8 * y + x];
}

/* ---- inter prediction ------------------------------------------------------------------------------ */
static inline int ref_px(const uint8_t *p, int W, int H, int x, int y)
{
	/* edge replication (inter_pred_luma_umv / fill_rect_umv_luma, h264.cpp:5932-6116) */
	return p[clip3(0, H - 1, y) * W + clip3(0, W - 1, x)];
}

static int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

/* one luma sample at integer position (x, y) + quarter-pel fraction (fx, fy) (8.4.2.2.1) */
static int luma_sample(const uint8_t *p, int W, int H, int x, int y, int fx, int fy)
{
#define P(dx, dy) ref_px(p, W, H, x + (dx), y + (dy))
#define HB(dx, dy) tap6(P((dx) - 2, dy), P((dx) - 1, dy), P(dx, dy), P((dx) + 1, dy), P((dx) + 2, dy), P((dx) + 3, dy))   /* b1 at (dx+1/2, dy) */
#define VH(dx, dy) tap6(P(dx, (dy) - 2), P(dx, (dy) - 1), P(dx, dy), P(dx, (dy) + 1), P(dx, (dy) + 2), P(dx, (dy) + 3))   /* h1 at (dx, dy+1/2) */
	if (!fx && !fy) return P(0, 0);
	int G = P(0, 0);
	int b = clip255((HB(0, 0) + 16) >> 5);
	int h = clip255((VH(0, 0) + 16) >> 5);
	if (fy == 0) return fx == 2 ? b : (fx == 1 ? (G + b + 1) >> 1 : (P(1, 0) + b + 1) >> 1);
	if (fx == 0) return fy == 2 ? h : (fy == 1 ? (G + h + 1) >> 1 : (P(0, 1) + h + 1) >> 1);
	/* j: 6-tap over the unrounded horizontal intermediates */
	int j = clip255((tap6(HB(0, -2), HB(0, -1), HB(0, 0), HB(0, 1), HB(0, 2), HB(0, 3)) + 512) >> 10);
	int s = clip255((HB(0, 1) + 16) >> 5);     /* half-pel below b */
	int m = clip255((VH(1, 0) + 16) >> 5);     /* half-pel right of h */
	if (fx == 2 && fy == 2) return j;
	if (fx == 2) return fy == 1 ? (b + j + 1) >> 1 : (s + j + 1) >> 1;        /* f, q */
	if (fy == 2) return fx == 1 ? (h + j + 1) >> 1 : (m + j + 1) >> 1;        /* i, k */
	/* diagonal quarter positions e, g, p, r */
	if (fx == 1 && fy == 1) return (b + h + 1) >> 1;
	if (fx == 3 && fy == 1) return (b + m + 1) >> 1;
	if (fx == 1 && fy == 3) return (h + s + 1) >> 1;
	return (s + m + 1) >> 1;
#undef P
#undef HB
#undef VH
}

/* weighted sample prediction (8.4.2.3; inter_pred_weighted_onedir / _bidir, weighted_copy,
 * add_bidir_weighted_type1/2: h264.cpp:6811-6977, 7027-7099).  n = number of lists used. */
static int weigh(const b2_h264_mc *r, int comp, int n, int first, int p0, int p1)
{
	int logwd = comp ? r->log2wd_c : r->log2wd_y;
	if (n == 1) {
		int w = r->w[first][comp], o = r->o[first][comp];
		int rnd = logwd ? 1 << (logwd - 1) : 0;
		return clip255(((p0 * w + rnd) >> logwd) + o);
	}
	return clip255(((p0 * r->w[0][comp] + p1 * r->w[1][comp] + (1 << logwd)) >> (logwd + 1)) + ((r->o[0][comp] + r->o[1][comp] + 1) >> 1));
}

static void inter_pred(const b2_h264_picture *p, const b2_h264_mc *r, int mbx, int mby, uint8_t *const *luma,
                       uint8_t *const *chroma)
{
	int W = p->mb_w * 16, H = p->mb_h * 16;
	int x0 = mbx * 16 + (r->pos & 3) * 4, y0 = mby * 16 + (r->pos >> 2) * 4;
	int w = (r->size & 15) * 4, h = (r->size >> 4) * 4;
	uint8_t *dy = luma[p->dst], *dc = chroma[p->dst];
	int n = (r->ref_frame[0] >= 0) + (r->ref_frame[1] >= 0);
	int first = r->ref_frame[0] >= 0 ? 0 : 1;
	int cw = W / 2, ch = H / 2;
	for (int y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x) {
			int pr[2] = {0, 0}, k = 0;
			for (int l = 0; l < 2; ++l) {
				if (r->ref_frame[l] < 0) continue;
				int mvx = r->mv[l][0], mvy = r->mv[l][1];
				pr[k++] = luma_sample(luma[r->ref_frame[l]], W, H, x0 + x + (mvx >> 2), y0 + y + (mvy >> 2), mvx & 3, mvy & 3);
			}
			dy[(y0 + y) * W + x0 + x] = (uint8_t)weigh(r, 0, n, first, pr[0], pr[1]);
		}
	/* chroma (8.4.2.2.2; h264.cpp:5267-5296): eighth-pel bilinear on each component */
	for (int y = 0; y < h / 2; ++y)
		for (int x = 0; x < w / 2; ++x)
			for (int c = 0; c < 2; ++c) {
				int pr[2] = {0, 0}, k = 0;
				for (int l = 0; l < 2; ++l) {
					if (r->ref_frame[l] < 0) continue;
					const uint8_t *rc = chroma[r->ref_frame[l]];
					int mvx = r->mv[l][0], mvy = r->mv[l][1], fx = mvx & 7, fy = mvy & 7;
					int cx = x0 / 2 + x + (mvx >> 3), cy = y0 / 2 + y + (mvy >> 3);
#define C(dx, dy) rc[clip3(0, ch - 1, cy + (dy)) * W + clip3(0, cw - 1, cx + (dx)) * 2 + c]
					pr[k++] = ((8 - fx) * (8 - fy) * C(0, 0) + fx * (8 - fy) * C(1, 0) + (8 - fx) * fy * C(0, 1) + fx * fy * C(1, 1) + 32) >> 6;
#undef C
				}
				dc[(y0 / 2 + y) * W + (x0 / 2 + x) * 2 + c] = (uint8_t)weigh(r, 1 + c, n, first, pr[0], pr[1]);
			}
}

/* ---- one macroblock ------------------------------------------------------------------------------------ */
static const int16_t *rec_of(const b2_h264_picture *p, const b2_h264_mb *m, int bit)
{
	if (!(m->nz & (1u << bit))) return 0;
	return p->coefs + ((size_t)m->coef_idx + __builtin_popcount(m->nz & ((1u << bit) - 1))) * 16;
}

static void chroma_residual(const b2_h264_picture *p, const b2_h264_mb *m, uint8_t *dc, int W)
{
	const int16_t *cdc = rec_of(p, m, 17);
	for (int c = 0; c < 2; ++c) {
		int qp = c ? m->qp_cr : m->qp_cb;
		int f[4] = {0, 0, 0, 0}, dcv[4];
		if (cdc) for (int i = 0; i < 4; ++i) f[i] = cdc[c * 4 + i] * level_scale4(qp, 0);
		/* intra_chroma_dc_transform (h264.cpp:4387-4404): 2x2 Hadamard, >> 1 */
		dcv[0] = (f[0] + f[1] + f[2] + f[3]) >> 1;
		dcv[1] = (f[0] - f[1] + f[2] - f[3]) >> 1;
		dcv[2] = (f[0] + f[1] - f[2] - f[3]) >> 1;
		dcv[3] = (f[0] - f[1] - f[2] + f[3]) >> 1;
		for (int b = 0; b < 4; ++b) {
			const int16_t *ac = rec_of(p, m, 18 + c * 4 + b);
			if (!ac && !dcv[b]) continue;
			add_block4x4(dc + (b >> 1) * 4 * W + (b & 1) * 8 + c, W, 2, ac, qp, &dcv[b]);
		}
	}
}

static void recon_mb(const b2_h264_picture *p, int mbx, int mby, uint8_t *const *luma, uint8_t *const *chroma)
{
	const b2_h264_mb *m = &p->mbs[mby * p->mb_w + mbx];
	int W = p->mb_w * 16;
	uint8_t *dy = luma[p->dst] + mby * 16 * W + mbx * 16;
	uint8_t *dc = chroma[p->dst] + mby * 8 * W + mbx * 16;
	if (m->type == B2_H4_IPCM) {
		const uint8_t *src = (const uint8_t *)(p->coefs + (size_t)m->coef_idx * 16);
		for (int y = 0; y < 16; ++y) memcpy(dy + y * W, src + 16 * y, 16);
		for (int c = 0; c < 2; ++c)
			for (int y = 0; y < 8; ++y)
				for (int x = 0; x < 8; ++x) dc[y * W + 2 * x + c] = src[256 + c * 64 + 8 * y + x];
		return;
	}
	if (m->type == B2_H4_INTER) {
		for (int i = 0; i < m->mc_cnt; ++i) inter_pred(p, &p->mcs[m->mc_idx + i], mbx, mby, luma, chroma);
		if (m->flags & B2_H4_T8x8) {
			for (int q = 0; q < 4; ++q) {
				const int16_t *lv = rec_of(p, m, 1 + 4 * q);
				if (lv) add_block8x8(dy + (q >> 1) * 8 * W + (q & 1) * 8, W, lv, m->qp_y);
			}
		} else
		for (int k = 0; k < 16; ++k) {
			const int16_t *lv = rec_of(p, m, 1 + k);
			if (lv) add_block4x4(dy + blk_y[k] * 4 * W + blk_x[k] * 4, W, 1, lv, m->qp_y, 0);
		}
		chroma_residual(p, m, dc, W);
		return;
	}
	if (m->type == B2_H4_I4x4) {
		for (int k = 0; k < 16; ++k) {
			uint8_t *d = dy + blk_y[k] * 4 * W + blk_x[k] * 4;
			int mode = (m->i4_modes[k >> 1] >> ((k & 1) * 4)) & 15;
			pred4x4(d, W, mode, blk_avail(k, m->intra_avail));
			const int16_t *lv = rec_of(p, m, 1 + k);
			if (lv) add_block4x4(d, W, 1, lv, m->qp_y, 0);
		}
	} else if (m->type == B2_H4_I8x8) {   /* luma_intra8x8_with_residual (h264.cpp:4083-4127) */
		for (int q = 0; q < 4; ++q) {
			uint8_t *d = dy + (q >> 1) * 8 * W + (q & 1) * 8;
			int mode = (m->i4_modes[q >> 1] >> ((q & 1) * 4)) & 15;
			pred8x8l(d, W, mode, blk8_avail(q, m->intra_avail));
			const int16_t *lv = rec_of(p, m, 1 + 4 * q);
			if (lv) add_block8x8(d, W, lv, m->qp_y);
		}
	} else {   /* Intra16x16 (h264.cpp:4407-4552) */
		pred16x16(dy, W, m->i16_mode, m->intra_avail);
		const int16_t *dcl = rec_of(p, m, 0);
		int dcv[16];
		memset(dcv, 0, sizeof(dcv));
		if (dcl) {
			/* intra16x16_dc_transform (h264.cpp:4309-4365): 4x4 Hadamard of dequantised DCs, (x+2)>>2 */
			int f[16], t[16];
			for (int i = 0; i < 16; ++i) f[i] = dcl[i] * level_scale4(m->qp_y, 0);
			for (int y = 0; y < 4; ++y) {
				int a = f[4 * y], b = f[4 * y + 1], c = f[4 * y + 2], d = f[4 * y + 3];
				t[4 * y] = a + b + c + d; t[4 * y + 1] = a + b - c - d; t[4 * y + 2] = a - b - c + d; t[4 * y + 3] = a - b + c - d;
			}
			for (int x = 0; x < 4; ++x) {
				int a = t[x], b = t[4 + x], c = t[8 + x], d = t[12 + x];
				f[x] = a + b + c + d; f[4 + x] = a + b - c - d; f[8 + x] = a - b - c + d; f[12 + x] = a - b + c - d;
			}
			for (int k = 0; k < 16; ++k) dcv[k] = (f[blk_y[k] * 4 + blk_x[k]] + 2) >> 2;
		}
		for (int k = 0; k < 16; ++k) {
			const int16_t *lv = rec_of(p, m, 1 + k);
			if (!lv && !dcv[k]) continue;
			add_block4x4(dy + blk_y[k] * 4 * W + blk_x[k] * 4, W, 1, lv, m->qp_y, &dcv[k]);
		}
	}
	pred_chroma(dc, W, m->chroma_mode, m->intra_avail_c);
	pred_chroma(dc + 1, W, m->chroma_mode, m->intra_avail_c);
	chroma_residual(p, m, dc, W);
}

/* ---- deblocking (8.7; h264.cpp:10253-10663) ------------------------------------------------------------- */
static const uint8_t alpha_tab[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20,
	22, 25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255};
static const uint8_t beta_tab[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 8, 8,
	9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18};
static const uint8_t tc0_tab[52][3] = {
	{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0},
	{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 1, 1},
	{0, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 2, 3}, {1, 2, 3},
	{2, 2, 3}, {2, 2, 4}, {2, 3, 4}, {2, 3, 4}, {3, 3, 5}, {3, 4, 6}, {3, 4, 6}, {4, 5, 7}, {4, 5, 8}, {4, 6, 9}, {5, 7, 10},
	{6, 8, 11}, {6, 8, 13}, {7, 10, 14}, {8, 11, 16}, {9, 12, 18}, {10, 13, 20}, {11, 15, 23}, {13, 17, 25}};

/* filter one line of samples across an edge; px points at q0, `st` is the step across the edge */
static void filter_line(uint8_t *px, int st, int bs, int alpha, int beta, int tc0, int luma)
{
	int p0 = px[-st], p1 = px[-2 * st], q0 = px[0], q1 = px[st];
	if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
	if (bs == 4) {
		if (luma) {
			int p2 = px[-3 * st], q2 = px[2 * st];
			int strong = abs(p0 - q0) < (alpha >> 2) + 2;
			if (strong && abs(p2 - p0) < beta) {
				int p3 = px[-4 * st];
				px[-st] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
				px[-2 * st] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
				px[-3 * st] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
			} else {
				px[-st] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
			}
			if (strong && abs(q2 - q0) < beta) {
				int q3 = px[3 * st];
				px[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
				px[st] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
				px[2 * st] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
			} else {
				px[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
			}
		} else {
			px[-st] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
			px[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
		}
		return;
	}
	int tc = tc0;
	if (luma) {
		int p2 = px[-3 * st], q2 = px[2 * st];
		int ap = abs(p2 - p0) < beta, aq = abs(q2 - q0) < beta;
		tc = tc0 + ap + aq;
		if (ap) px[-2 * st] = (uint8_t)(p1 + clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - 2 * p1) >> 1));
		if (aq) px[st] = (uint8_t)(q1 + clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - 2 * q1) >> 1));
	} else {
		tc = tc0 + 1;
	}
	int delta = clip3(-tc, tc, ((q0 - p0) * 4 + (p1 - q1) + 4) >> 3);
	px[-st] = (uint8_t)clip255(p0 + delta);
	px[0] = (uint8_t)clip255(q0 - delta);
}

/* one edge of 16 luma lines or 8 chroma lines; bs4: whole edge has bS 4; bits: 4 x 2-bit strengths */
static void filter_edge(uint8_t *px, int across, int along, int bits, int bs4, int qp, int aoff, int boff, int luma)
{
	int ia = clip3(0, 51, qp + aoff), ib = clip3(0, 51, qp + boff);
	int alpha = alpha_tab[ia], beta = beta_tab[ib];
	if (!alpha) return;
	int n = luma ? 16 : 8;
	for (int i = 0; i < n; ++i) {
		int seg = luma ? i >> 2 : i >> 1;
		int bs = (bits >> (seg * 2)) & 3;
		if (!bs) continue;
		if (bs4) bs = 4;
		filter_line(px + i * along, across, bs, alpha, beta, bs < 4 ? tc0_tab[ia][bs - 1] : 0, luma);
	}
}

static void deblock_mb(const b2_h264_picture *p, int mbx, int mby, uint8_t *luma, uint8_t *chroma)
{
	const b2_h264_dbk *d = &p->dbk[mby * p->mb_w + mbx];
	if (d->flags & B2_DBK_OFF) return;
	int W = p->mb_w * 16;
	uint8_t *y = luma + mby * 16 * W + mbx * 16, *c = chroma + mby * 8 * W + mbx * 16;
	int a = d->alpha_off, b = d->beta_off;
	for (int dir = 0; dir < 2; ++dir) {          /* 0: vertical edges (filter horizontally), 1: horizontal edges */
		uint32_t bits = dir ? d->bs_h : d->bs_v;
		int edge_ok = d->flags & (dir ? B2_DBK_TOP_EDGE : B2_DBK_LEFT_EDGE);
		int bs4 = (d->flags & (dir ? B2_DBK_TOP_BS4 : B2_DBK_LEFT_BS4)) != 0;
		const b2_h264_dbk *n = dir ? d - p->mb_w : d - 1;
		for (int e = 0; e < 4; ++e) {
			int eb = (bits >> (8 * e)) & 255;
			if (!eb || (e == 0 && !edge_ok)) continue;
			int qpy = e ? d->qp_y : (d->qp_y + n->qp_y + 1) >> 1;
			uint8_t *py = dir ? y + e * 4 * W : y + e * 4;
			filter_edge(py, dir ? W : 1, dir ? 1 : W, eb, e == 0 && bs4, qpy, a, b, 1);
			if (e & 1) continue;                 /* chroma: luma edges 0 and 2 only */
			for (int k = 0; k < 2; ++k) {
				int qc = k ? d->qp_cr : d->qp_cb, qn = k ? n->qp_cr : n->qp_cb;
				int qpc = e ? qc : (qc + qn + 1) >> 1;
				uint8_t *pc = dir ? c + (e / 2) * 4 * W + k : c + (e / 2) * 8 + k;
				filter_edge(pc, dir ? W : 2, dir ? 2 : W, eb, e == 0 && bs4, qpc, a, b, 0);
			}
		}
	}
}

int oracle_h264_picture(const b2_h264_picture *p, uint8_t *const *luma, uint8_t *const *chroma)
{
	for (int mby = 0; mby < p->mb_h; ++mby)
		for (int mbx = 0; mbx < p->mb_w; ++mbx) recon_mb(p, mbx, mby, luma, chroma);
	if (p->deblock)
		for (int mby = 0; mby < p->mb_h; ++mby)
			for (int mbx = 0; mbx < p->mb_w; ++mbx) deblock_mb(p, mbx, mby, luma[p->dst], chroma[p->dst]);
	return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Deblocking strengths from the records (test infrastructure like the rest of this file).
 * Restates what the reference's parse stores into deblock_info_t (store_info_inter*, str_mv_calc*,
 * str_previous_coef: h264.cpp:7119-7322, 7451-7604, 8796-9400; intra: 3 / bS 4 on macroblock edges, 8.7.2.1) from
 * the inputs the device kernel gets (motion records, b2_h264_mb.nnz4x4, b2_h264_dbk.bs_mode, the edge flags) and
 * compares the result with the strengths the host front end filled in.  Returns the number of macroblocks whose
 * bs_v / bs_h / BS4 flags differ (0 = the two derivations agree), -1 if the picture carries no host strengths.
 * tests/util.py runs it on every tapped picture, so every CPU test stream pins the derivation the CUDA kernel
 * (h264_bs_kernel) implements.
 * ------------------------------------------------------------------------------------------------ */
typedef struct { int8_t ref[2]; int16_t mv[2][2]; } bs_block;

static int bs_far(const int16_t *a, const int16_t *b)
{
	return (unsigned)(a[0] - b[0] + 3) > 6u || (unsigned)(a[1] - b[1] + 3) > 6u;
}

static int bs_pair(const bs_block *a, const bs_block *b, int nza, int nzb, int bslice, int reduced)
{
	if (nza | nzb) return 2;
	if (!bslice) return a->ref[0] != b->ref[0] || bs_far(a->mv[0], b->mv[0]);
	if (!((a->ref[0] == b->ref[0] && a->ref[1] == b->ref[1]) || (a->ref[0] == b->ref[1] && a->ref[1] == b->ref[0]))) return 1;
	if (a->ref[0] != a->ref[1]) {
		if (a->ref[0] == b->ref[0]) return (a->ref[0] >= 0 && bs_far(a->mv[0], b->mv[0])) || (a->ref[1] >= 0 && bs_far(a->mv[1], b->mv[1]));
		return (a->ref[0] >= 0 && bs_far(a->mv[0], b->mv[1])) || (a->ref[1] >= 0 && bs_far(a->mv[1], b->mv[0]));
	}
	int straight = bs_far(a->mv[0], b->mv[0]) || bs_far(a->mv[1], b->mv[1]), cross = bs_far(a->mv[0], b->mv[1]) || bs_far(a->mv[1], b->mv[0]);
	return reduced ? straight : (straight && cross);
}

int oracle_h264_bs_check(const b2_h264_picture *p)
{
	if (p->bs_on_device) return -1;
	const int w4 = p->mb_w * 4, n_mbs = p->mb_w * p->mb_h;
	bs_block *f = (bs_block *)calloc((size_t)n_mbs * 16, sizeof(bs_block));
	for (int mb = 0; mb < n_mbs; ++mb) {
		const b2_h264_mb *m = &p->mbs[mb];
		if (m->type != B2_H4_INTER) continue;
		int mbx = mb % p->mb_w, mby = mb / p->mb_w;
		for (int i = 0; i < m->mc_cnt; ++i) {
			const b2_h264_mc *r = &p->mcs[m->mc_idx + i];
			for (int y = r->pos >> 2; y < (r->pos >> 2) + (r->size >> 4); ++y)
				for (int x = r->pos & 3; x < (r->pos & 3) + (r->size & 15); ++x) {
					bs_block *b = &f[(size_t)(mby * 4 + y) * w4 + mbx * 4 + x];
					for (int l = 0; l < 2; ++l) { b->ref[l] = r->ref_frame[l]; b->mv[l][0] = r->mv[l][0]; b->mv[l][1] = r->mv[l][1]; }
				}
		}
	}
	int bad = 0;
	for (int mb = 0; mb < n_mbs; ++mb) {
		const b2_h264_mb *m = &p->mbs[mb];
		const b2_h264_dbk *d = &p->dbk[mb];
		if (d->flags & B2_DBK_OFF) continue;
		int mbx = mb % p->mb_w, mby = mb / p->mb_w;
		uint32_t bs[2] = {0, 0};
		int flags = d->flags & ~(B2_DBK_LEFT_BS4 | B2_DBK_TOP_BS4);
		for (int dir = 0; dir < 2; ++dir) {
			int edge_ok = d->flags & (dir ? B2_DBK_TOP_EDGE : B2_DBK_LEFT_EDGE);
			const b2_h264_mb *o = edge_ok ? (dir ? m - p->mb_w : m - 1) : NULL;
			for (int e = 0; e < 4; ++e)
				for (int sg = 0; sg < 4; ++sg) {
					int v;
					if (m->type != B2_H4_INTER) v = (e > 0 || o) ? 3 : 0;
					else if (e == 0 && !o) v = 0;
					else if (e == 0 && o->type != B2_H4_INTER) v = 3;
					else {
						int bx = dir ? sg : e, by = dir ? e : sg;
						int gx = mbx * 4 + bx, gy = mby * 4 + by;
						const bs_block *b = &f[(size_t)gy * w4 + gx], *a = dir ? b - w4 : b - 1;
						int nzb = (m->nnz4x4 >> (by * 4 + bx)) & 1, nza;
						if (e == 0) nza = (o->nnz4x4 >> (dir ? 12 + sg : sg * 4 + 3)) & 1;
						else nza = (m->nnz4x4 >> (dir ? (by - 1) * 4 + bx : by * 4 + bx - 1)) & 1;
						int red = d->bs_mode & (e == 0 ? B2_BS_REDUCED_MB_EDGE : (e == 2 ? B2_BS_REDUCED_CENTRE : B2_BS_REDUCED_INNER));
						v = bs_pair(a, b, nza, nzb, d->bs_mode & B2_BS_B_SLICE, red != 0);
					}
					bs[dir] |= (uint32_t)v << (e * 8 + sg * 2);
				}
			if (o && (m->type != B2_H4_INTER || o->type != B2_H4_INTER)) flags |= dir ? B2_DBK_TOP_BS4 : B2_DBK_LEFT_BS4;
		}
		if (m->flags & B2_H4_T8x8) { bs[0] &= 0x00ff00ffu; bs[1] &= 0x00ff00ffu; }
		if (bs[0] != d->bs_v || bs[1] != d->bs_h || flags != d->flags) ++bad;
	}
	free(f);
	return bad;
}

/* ------------------------------------------------------------------------------------------------
 * Motion vectors from the partitions (test infrastructure).  Restates motion vector prediction
 * for P macroblocks (standard 8.4.1.1 P_Skip, 8.4.1.3 median / directional prediction with the neighbour
 * availability of 6.4.8 and 6.4.11.7; reference calc_mv*, h264.cpp:6669-6724, 7379-7604) from the inputs the device
 * kernel (h264_mv_kernel) gets -- partitions with vector differences, b2_h264_mb.flags -- and compares vectors,
 * reference frames and geometry with the motion records the host front end derived.  Returns the number of
 * differing records (0 = the derivations agree), -1 if the picture carries no partitions next to its records.
 * ------------------------------------------------------------------------------------------------ */
typedef struct { int ref, mvx, mvy; } mv_nb;       /* ref: -2 unavailable, -1 intra */

static int med3(int a, int b, int c) { int mx = a > b ? a : b, mn = a < b ? a : b; return c > mx ? mx : (c < mn ? mn : c); }

int oracle_h264_mv_check(const b2_h264_picture *p)
{
	if (p->mv_on_device || !p->parts || p->n_parts != p->n_mcs) return -1;
	const int w4 = p->mb_w * 4, n_mbs = p->mb_w * p->mb_h;
	int8_t *fref = (int8_t *)malloc((size_t)n_mbs * 16);
	int16_t *fmv = (int16_t *)calloc((size_t)n_mbs * 32, sizeof(int16_t));
	memset(fref, -2, (size_t)n_mbs * 16);           /* nothing decoded yet */
	int bad = 0;
	for (int mb = 0; mb < n_mbs; ++mb) {
		const b2_h264_mb *m = &p->mbs[mb];
		const int mbx = mb % p->mb_w, mby = mb / p->mb_w, avail = m->flags >> B2_H4_NB_SHIFT;
		if (m->type != B2_H4_INTER) {
			for (int y = 0; y < 4; ++y) memset(&fref[(size_t)(mby * 4 + y) * w4 + mbx * 4], -1, 4);
			continue;
		}
		for (int i = 0; i < m->mc_cnt; ++i) {
			const b2_h264_part *pt = &p->parts[m->mc_idx + i];
			const b2_h264_mc *want = &p->mcs[m->mc_idx + i];
			const int x4 = pt->pos & 3, y4 = pt->pos >> 2, pw = pt->size & 15, ph = pt->size >> 4;
			mv_nb nb[4];      /* A left, B above, C above-right, D above-left */
			const int dx[4] = {-1, 0, pw, -1}, dy[4] = {0, -1, -1, -1};
			for (int k = 0; k < 4; ++k) {
				int x = x4 + dx[k], y = y4 + dy[k];
				int rx = x < 0 ? -1 : (x > 3 ? 1 : 0), ry = y < 0 ? -1 : 0;
				int ok = ry == 0 ? (rx == 0 || (rx < 0 && (avail & 1))) : (rx == 0 ? (avail & 2) : (rx < 0 ? (avail & 8) : (avail & 4)));
				nb[k].ref = -2; nb[k].mvx = nb[k].mvy = 0;
				if (!ok) continue;
				size_t g = (size_t)(mby * 4 + y) * w4 + mbx * 4 + x;
				nb[k].ref = fref[g];
				if (nb[k].ref >= 0) { nb[k].mvx = fmv[g * 2]; nb[k].mvy = fmv[g * 2 + 1]; }
			}
			mv_nb a = nb[0], b = nb[1], c = nb[2].ref != -2 ? nb[2] : nb[3];
			int px = 0, py = 0, done = 0;
			if (m->flags & B2_H4_PSKIP) {
				if (a.ref == -2 || b.ref == -2 || (a.ref == 0 && !a.mvx && !a.mvy) || (b.ref == 0 && !b.mvx && !b.mvy)) done = 1;
			} else if (pw == 4 && ph == 2) {
				if (i == 0 && b.ref == pt->ref_idx) { px = b.mvx; py = b.mvy; done = 1; }
				if (i == 1 && a.ref == pt->ref_idx) { px = a.mvx; py = a.mvy; done = 1; }
			} else if (pw == 2 && ph == 4) {
				if (i == 0 && a.ref == pt->ref_idx) { px = a.mvx; py = a.mvy; done = 1; }
				if (i == 1 && c.ref == pt->ref_idx) { px = c.mvx; py = c.mvy; done = 1; }
			}
			if (!done) {
				if (b.ref == -2 && c.ref == -2 && a.ref != -2) { b = a; c = a; }
				int ma = a.ref == pt->ref_idx, mb_ = b.ref == pt->ref_idx, mc = c.ref == pt->ref_idx;
				if (ma + mb_ + mc == 1) { const mv_nb *s = ma ? &a : (mb_ ? &b : &c); px = s->mvx; py = s->mvy; }
				else { px = med3(a.mvx, b.mvx, c.mvx); py = med3(a.mvy, b.mvy, c.mvy); }
			}
			int mvx = (int16_t)(px + pt->mvd[0]), mvy = (int16_t)(py + pt->mvd[1]);
			for (int y = y4; y < y4 + ph; ++y)
				for (int x = x4; x < x4 + pw; ++x) {
					size_t g = (size_t)(mby * 4 + y) * w4 + mbx * 4 + x;
					fref[g] = pt->ref_idx; fmv[g * 2] = (int16_t)mvx; fmv[g * 2 + 1] = (int16_t)mvy;
				}
			if (want->pos != pt->pos || want->size != pt->size || want->ref_frame[0] != pt->ref_frame || want->ref_frame[1] != -1 ||
			    want->mv[0][0] != mvx || want->mv[0][1] != mvy) ++bad;
		}
	}
	free(fref);
	free(fmv);
	return bad;
}
