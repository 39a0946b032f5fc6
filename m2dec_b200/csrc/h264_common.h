// h264_common.h — H.264 syntax-level context shared by the host front end (h264_front.cpp)
// and the test-only bitstream writer (tools/streamgen): scan tables, the coded_block_pattern
// mapping, and the neighbour derivations of clauses 6.4.11, 8.3.1.1, 8.4.1.3 and 9.2.1
// (availability, Intra4x4PredMode prediction, motion-vector prediction, CAVLC nC).
//
// The reference keeps this state as packed per-row words inside h264d_mb_current
// (left4x4pred/top4x4pred, left4x4coef/top4x4coef, left4x4inter/top4x4inter, h264.h:330-395)
// and derives neighbours with hand-unrolled code (calc_mv*, h264.cpp:6669-6724, 7379-8330).
// Here the picture keeps full-frame per-4x4 arrays, which makes every neighbour a lookup.
#pragma once
#include <stdint.h>
#include <string.h>
#include <memory>
#include <vector>

namespace b2 { namespace h4 {

// 4x4 luma block k (decode order) -> position inside the MB in 4-pel units
static const uint8_t blk_x[16] = {0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3};
static const uint8_t blk_y[16] = {0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3};
// raster position (y4 * 4 + x4) -> decode-order index
static const uint8_t raster_to_blk[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};

// zig-zag scan index -> raster position 4*y + x (frame scan, Table 8-2 / Figure 8-8)
static const uint8_t zigzag4x4[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
static const uint8_t zigzag8x8[64] = {
	0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20,
	13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59,
	52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// Table 9-4: codeNum -> coded_block_pattern for [0] Intra4x4/8x8, [1] Inter (ChromaArrayType 1)
static const uint8_t cbp_from_code[2][48] = {
	{47, 31, 15, 0, 23, 27, 29, 30, 7, 11, 13, 14, 39, 43, 45, 46, 16, 3, 5, 10, 12, 19, 21, 26,
	 28, 35, 37, 42, 44, 1, 2, 4, 8, 17, 18, 20, 24, 6, 9, 22, 25, 32, 33, 34, 36, 40, 38, 41},
	{0, 16, 1, 2, 4, 8, 32, 3, 5, 10, 12, 15, 47, 7, 11, 13, 14, 6, 9, 31, 35, 37, 42, 44,
	 33, 34, 36, 40, 39, 43, 45, 46, 17, 18, 20, 24, 19, 21, 26, 28, 23, 27, 29, 30, 22, 25, 38, 41}};

// Table 8-15: qPI -> QPC
static inline int chroma_qp(int qp_y, int offset) {
	static const uint8_t tab[22] = {29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39};
	int q = qp_y + offset;
	if (q < 0) q = 0;
	if (q > 51) q = 51;
	return q < 30 ? q : tab[q - 30];
}

enum MbKind : uint8_t { MB_NONE = 0, MB_INTER, MB_SKIP, MB_I4x4, MB_I8x8, MB_I16x16, MB_IPCM };
static inline bool kind_is_intra(uint8_t k) { return k >= MB_I4x4; }

struct MbInfo {
	uint16_t slice;     // slice number inside the picture, 0xffff = not decoded yet
	uint8_t kind;
	uint8_t qp_y;       // QP'Y after this MB (the predictor for the next mb_qp_delta)
	uint8_t cbp;
	uint8_t t8x8;
	uint8_t chroma_mode;
	uint8_t cflags;     // CABAC neighbour state: MBF_* bits
};
// MbInfo.cflags: coded_block_flag of the DC blocks, and "B_Skip or B_Direct_16x16"
enum { MBF_CBF_LUMA_DC = 1, MBF_CBF_CB_DC = 2, MBF_CBF_CR_DC = 4, MBF_DIRECT16 = 8, MBF_SHAPE_SHIFT = 4 };
// bits 4..5 of cflags: partition shape of the macroblock as the reference records it for a
// co-located picture (col_mbtype_t, h264.h:84-89): 0 16x16 (also intra), 1 16x8, 2 8x16, 3 8x8
enum { SHAPE_16x16 = 0, SHAPE_16x8 = 1, SHAPE_8x16 = 2, SHAPE_8x8 = 3 };

struct Neighbour {
	bool avail;
	int ref;            // -1: intra or list unused
	int16_t mv[2];
};

// Reference picture marking + list construction state (clauses 8.2.4, 8.2.5), shared so the
// writer can produce streams whose lists it knows.
struct RefPic {
	int frame_idx;     // frame-pool index
	int frame_num;     // FrameNum
	int long_idx;      // LongTermFrameIdx, -1 when short term
	int poc;
};

// Motion of a decoded reference picture as direct prediction needs it (co-located picture, 8.4.1.2.1)
struct ColBlk {
	int8_t ref;        // reference index of the list the block used (list 0 preferred), -1 = intra
	int8_t pad;
	int16_t mv[2];
	int32_t refpoc;    // POC of the picture that vector points to
};
struct ColPic {
	std::unique_ptr<ColBlk[]> b;  // n blocks (4x4 raster); plain storage: store_col writes every field of every block
	size_t n = 0;
	std::vector<uint8_t> shape;   // per macroblock, SHAPE_*
};
// result of direct prediction for one macroblock, per 4x4 block in raster order
struct DirectOut {
	int ref[16][2];
	int mv[16][2][2];
};

// Per-picture neighbour context.
class PicCtx {
public:
	int mb_w = 0, mb_h = 0, w4 = 0, h4 = 0;
	std::vector<MbInfo> mb;
	std::vector<uint8_t> nnz;        // [mb][24]: luma raster 0..15, Cb 16..19, Cr 20..23
	std::vector<int8_t> ipred;       // [mb][16] raster Intra4x4/8x8PredMode (2 when not I4x4/I8x8)
	std::vector<int16_t> mv[2];      // [(gy * w4 + gx) * 2]
	std::vector<int8_t> ref[2];      // per 4x4: -2 not decoded, -1 intra / unused list, >= 0 refIdx
	std::vector<int8_t> refpic[2];   // per 4x4: frame-pool index of the reference picture, -1 none
	// CABAC only (9.3.3.1.1.6 / .7): per 4x4 block |mvd| (x, y; saturated) and "refIdx > 0 and not
	// direct-predicted", both zero for everything that codes neither
	bool cabac = false;
	bool keep_motion = true;
	std::vector<uint8_t> mvd[2];
	std::vector<uint8_t> refgt0[2];

	void resize(int mbw, int mbh) {
		mb_w = mbw; mb_h = mbh; w4 = mbw * 4; h4 = mbh * 4;
		size_t n = (size_t)mbw * mbh;
		mb.assign(n, MbInfo());
		nnz.assign(n * 24, 0);
		ipred.assign(n * 16, 2);
		for (int l = 0; l < 2; ++l) {
			mv[l].assign(n * 16 * 2, 0);
			ref[l].assign(n * 16, -2);
			refpic[l].assign(n * 16, -1);
			mvd[l].assign(n * 16 * 2, 0);
			refgt0[l].assign(n * 16, 0);
		}
	}
	// motion: the picture's 4x4 motion field is kept here (not when the device predicts the vectors: nothing on the
	// host reads it then, and clearing it is 1.6 MB of stores per 1080p picture)
	void begin_picture(bool motion = true) {
		keep_motion = motion;
		for (auto &m : mb) { m = MbInfo(); m.slice = 0xffff; }
		std::fill(nnz.begin(), nnz.end(), 0);
		std::fill(ipred.begin(), ipred.end(), 2);
		for (int l = 0; l < 2 && motion; ++l) {
			std::fill(mv[l].begin(), mv[l].end(), 0);
			std::fill(ref[l].begin(), ref[l].end(), -2);
			std::fill(refpic[l].begin(), refpic[l].end(), -1);
		}
	}

	// --- macroblock-level availability (6.4.8): same slice and already decoded --------------
	inline bool mb_avail(int mbx, int mby, int cur_slice) const {
		if (mbx < 0 || mby < 0 || mbx >= mb_w || mby >= mb_h) return false;
		return mb[(size_t)mby * mb_w + mbx].slice == cur_slice;
	}
	// bit0 left, bit1 top, bit2 top-right, bit3 top-left
	inline int mb_avail_mask(int mbx, int mby, int cur_slice) const {
		return (mb_avail(mbx - 1, mby, cur_slice) ? 1 : 0) | (mb_avail(mbx, mby - 1, cur_slice) ? 2 : 0) |
		       (mb_avail(mbx + 1, mby - 1, cur_slice) ? 4 : 0) | (mb_avail(mbx - 1, mby - 1, cur_slice) ? 8 : 0);
	}

	// --- start of a macroblock: its 16 blocks become "not yet decoded" ------------------------
	// Also caches what every neighbour derivation of this macroblock needs (availability of the
	// four neighbouring MBs, pointers to their coefficient counts).
	struct Env {
		int mbx, mby;
		bool availL, availT, availTR, availTL;
		const uint8_t *nnz_left, *nnz_top;
		uint8_t *nnz_cur;
	} env;

	void open_mb(int mbx, int mby, int cur_slice) {
		MbInfo &m = mb[(size_t)mby * mb_w + mbx];
		m.slice = (uint16_t)cur_slice;
		for (int y = 0; y < 4 && keep_motion; ++y) {
			size_t g = (size_t)(mby * 4 + y) * w4 + mbx * 4;
			const uint32_t undecoded = 0xFEFEFEFEu;       // -2 in every block
			memcpy(&ref[0][g], &undecoded, 4);
			memcpy(&ref[1][g], &undecoded, 4);
			if (cabac)
				for (int l = 0; l < 2; ++l) {
					memset(&mvd[l][g * 2], 0, 8);
					memset(&refgt0[l][g], 0, 4);
				}
		}
		env.mbx = mbx;
		env.mby = mby;
		env.availL = mb_avail(mbx - 1, mby, cur_slice);
		env.availT = mb_avail(mbx, mby - 1, cur_slice);
		env.availTR = mb_avail(mbx + 1, mby - 1, cur_slice);
		env.availTL = mb_avail(mbx - 1, mby - 1, cur_slice);
		env.nnz_cur = &nnz[((size_t)mby * mb_w + mbx) * 24];
		env.nnz_left = env.availL ? env.nnz_cur - 24 : nullptr;
		env.nnz_top = env.availT ? env.nnz_cur - (size_t)mb_w * 24 : nullptr;
	}

	// fill a w4 x h4 partition (4-pel units) of the current MB
	void set_motion(int list, int mbx, int mby, int x4, int y4, int pw, int ph, int refidx, int pic, int mvx, int mvy) {
		const uint32_t packed = (uint32_t)(uint16_t)mvx | ((uint32_t)(uint16_t)mvy << 16);
		size_t g = (size_t)(mby * 4 + y4) * w4 + mbx * 4 + x4;
		int8_t *rp = &ref[list][g], *pp = &refpic[list][g];
		uint32_t *m = reinterpret_cast<uint32_t *>(&mv[list][g * 2]);
		// pw is 1, 2 or 4: fixed-size stores instead of memset calls
		if (pw == 4) {
			const uint32_t r4 = 0x01010101u * (uint8_t)refidx, p4 = 0x01010101u * (uint8_t)pic;
			for (int y = 0; y < ph; ++y, rp += w4, pp += w4, m += w4) {
				memcpy(rp, &r4, 4);
				memcpy(pp, &p4, 4);
				m[0] = m[1] = m[2] = m[3] = packed;
			}
		} else if (pw == 2) {
			const uint16_t r2 = (uint16_t)(0x0101u * (uint8_t)refidx), p2 = (uint16_t)(0x0101u * (uint8_t)pic);
			for (int y = 0; y < ph; ++y, rp += w4, pp += w4, m += w4) {
				memcpy(rp, &r2, 2);
				memcpy(pp, &p2, 2);
				m[0] = m[1] = packed;
			}
		} else {
			for (int y = 0; y < ph; ++y, rp += w4, pp += w4, m += w4) {
				for (int x = 0; x < pw; ++x) { rp[x] = (int8_t)refidx; pp[x] = (int8_t)pic; m[x] = packed; }
			}
		}
	}
	void set_intra_motion(int mbx, int mby) {
		for (int l = 0; l < 2; ++l) set_motion(l, mbx, mby, 0, 0, 4, 4, -1, -1, 0, 0);
	}

	// neighbouring 4x4 block at global 4-pel coordinates (gx, gy)
	inline Neighbour nb(int list, int gx, int gy, int /*cur_slice*/) const {
		Neighbour n = {false, -1, {0, 0}};
		// (gx, gy) lies in the current macroblock or in one of its four causal neighbours
		const int rx = (gx >> 2) - env.mbx, ry = (gy >> 2) - env.mby;
		bool ok;
		if (ry == 0) ok = rx == 0 || (rx < 0 && env.availL);
		else if (ry < 0) ok = rx == 0 ? env.availT : (rx < 0 ? env.availTL : env.availTR);
		else ok = false;
		if (!ok) return n;
		size_t g = (size_t)gy * w4 + gx;
		int r = ref[list][g];
		if (r == -2) return n;
		n.avail = true;
		n.ref = r;
		if (r >= 0) { n.mv[0] = mv[list][g * 2]; n.mv[1] = mv[list][g * 2 + 1]; }
		return n;
	}

	static inline int median3(int a, int b, int c) {
		int mx = a > b ? a : b, mn = a < b ? a : b;
		return c > mx ? mx : (c < mn ? mn : c);
	}

	// 8.4.1.3: motion vector prediction for a partition at (x4, y4) of size pw x ph (4-pel units)
	// shape: 0 generic/median, 1 = 16x8 partition (part = 0 top, 1 bottom), 2 = 8x16 (0 left, 1 right)
	void predict_mv(int list, int mbx, int mby, int x4, int y4, int pw, int refidx, int cur_slice,
	                int shape, int part, int16_t out[2]) const {
		int gx = mbx * 4 + x4, gy = mby * 4 + y4;
		Neighbour a = nb(list, gx - 1, gy, cur_slice);
		Neighbour b = nb(list, gx, gy - 1, cur_slice);
		Neighbour c = nb(list, gx + pw, gy - 1, cur_slice);
		if (!c.avail) c = nb(list, gx - 1, gy - 1, cur_slice);
		if (shape == 1) {
			if (part == 0 && b.avail && b.ref == refidx) { out[0] = b.mv[0]; out[1] = b.mv[1]; return; }
			if (part == 1 && a.avail && a.ref == refidx) { out[0] = a.mv[0]; out[1] = a.mv[1]; return; }
		} else if (shape == 2) {
			if (part == 0 && a.avail && a.ref == refidx) { out[0] = a.mv[0]; out[1] = a.mv[1]; return; }
			if (part == 1 && c.avail && c.ref == refidx) { out[0] = c.mv[0]; out[1] = c.mv[1]; return; }
		}
		if (!b.avail && !c.avail && a.avail) { b = a; c = a; }
		int ma = a.avail && a.ref == refidx, mb_ = b.avail && b.ref == refidx, mc = c.avail && c.ref == refidx;
		if (ma + mb_ + mc == 1) {
			const Neighbour &s = ma ? a : (mb_ ? b : c);
			out[0] = s.mv[0]; out[1] = s.mv[1];
			return;
		}
		out[0] = (int16_t)median3(a.mv[0], b.mv[0], c.mv[0]);
		out[1] = (int16_t)median3(a.mv[1], b.mv[1], c.mv[1]);
	}

	// 8.4.1.1: P_Skip motion vector
	void pskip_mv(int mbx, int mby, int cur_slice, int16_t out[2]) const {
		int gx = mbx * 4, gy = mby * 4;
		Neighbour a = nb(0, gx - 1, gy, cur_slice);
		Neighbour b = nb(0, gx, gy - 1, cur_slice);
		out[0] = out[1] = 0;
		if (!a.avail || !b.avail) return;
		if (a.ref == 0 && a.mv[0] == 0 && a.mv[1] == 0) return;
		if (b.ref == 0 && b.mv[0] == 0 && b.mv[1] == 0) return;
		predict_mv(0, mbx, mby, 0, 0, 4, 0, cur_slice, 0, 0, out);
	}

	// 9.2.1: nC for the block at (x, y) in units of blocks; plane 0 luma (4x4 grid), 1 Cb, 2 Cr (2x2 grid)
	inline int cavlc_nc(int /*mbx*/, int /*mby*/, int plane, int x, int y, int /*cur_slice*/) const {
		const int dim = plane ? 2 : 4, base = plane ? 12 + plane * 4 : 0;
		int na = -1, nb_ = -1;
		if (x > 0) na = env.nnz_cur[base + y * dim + x - 1];
		else if (env.nnz_left) na = env.nnz_left[base + y * dim + dim - 1];
		if (y > 0) nb_ = env.nnz_cur[base + (y - 1) * dim + x];
		else if (env.nnz_top) nb_ = env.nnz_top[base + (dim - 1) * dim + x];
		if (na >= 0 && nb_ >= 0) return (na + nb_ + 1) >> 1;
		if (na >= 0) return na;
		if (nb_ >= 0) return nb_;
		return 0;
	}
	inline uint8_t *nnz_of(int mbx, int mby) { return &nnz[((size_t)mby * mb_w + mbx) * 24]; }
	inline int8_t *ipred_of(int mbx, int mby) { return &ipred[((size_t)mby * mb_w + mbx) * 16]; }
	inline MbInfo &info(int mbx, int mby) { return mb[(size_t)mby * mb_w + mbx]; }
	inline const MbInfo &info(int mbx, int mby) const { return mb[(size_t)mby * mb_w + mbx]; }

	// snapshot of this picture's motion, taken when it becomes a reference picture
	template <class F>
	void store_col(ColPic &c, F poc_of_frame) const {
		if (c.n != (size_t)w4 * h4) {
			c.n = (size_t)w4 * h4;
			c.b.reset(new ColBlk[c.n]);
		}
		c.shape.resize(mb.size());
		for (size_t i = 0; i < mb.size(); ++i) c.shape[i] = (uint8_t)((mb[i].cflags >> MBF_SHAPE_SHIFT) & 3);
		for (size_t g = 0; g < c.n; ++g) {
			ColBlk &o = c.b[g];
			int l = ref[0][g] >= 0 ? 0 : (ref[1][g] >= 0 ? 1 : -1);
			o.pad = 0;
			if (l < 0) { o.ref = -1; o.mv[0] = o.mv[1] = 0; o.refpoc = 0; continue; }
			o.ref = ref[l][g];
			o.mv[0] = mv[l][g * 2];
			o.mv[1] = mv[l][g * 2 + 1];
			o.refpoc = poc_of_frame(refpic[l][g]);
		}
	}

	static inline int min_positive(int a, int b) { return (a >= 0 && b >= 0) ? (a < b ? a : b) : (a > b ? a : b); }

	// 8.4.1.2.2: spatial direct prediction of the whole macroblock (open_mb must have been called)
	// col_shape >= 0: take the co-located block per partition of that shape (what the reference
	// does for B_Skip / B_Direct_16x16, h264.cpp:9790-9850) instead of per corner / 4x4 block;
	// both agree unless the reference has overwritten the recorded shape (see derive_direct).
	// refs_out (optional): the two reference indices before the "both negative" substitution.
	void direct_spatial(int mbx, int mby, int cur_slice, const ColPic *col, bool col_short_term, bool inference8x8,
	                    DirectOut &out, int col_shape = -1, int *refs_out = nullptr) const {
		const int gx = mbx * 4, gy = mby * 4;
		int refs[2];
		int16_t mvp[2][2] = {{0, 0}, {0, 0}};
		for (int l = 0; l < 2; ++l) {
			Neighbour a = nb(l, gx - 1, gy, cur_slice), b = nb(l, gx, gy - 1, cur_slice), c = nb(l, gx + 4, gy - 1, cur_slice);
			if (!c.avail) c = nb(l, gx - 1, gy - 1, cur_slice);
			refs[l] = min_positive(a.avail ? a.ref : -1, min_positive(b.avail ? b.ref : -1, c.avail ? c.ref : -1));
		}
		bool zero = false;
		if (refs_out) { refs_out[0] = refs[0]; refs_out[1] = refs[1]; }
		if (refs[0] < 0 && refs[1] < 0) { refs[0] = refs[1] = 0; zero = true; }
		else
			for (int l = 0; l < 2; ++l)
				if (refs[l] >= 0) predict_mv(l, mbx, mby, 0, 0, 4, refs[l], cur_slice, 0, 0, mvp[l]);
		for (int y = 0; y < 4; ++y)
			for (int x = 0; x < 4; ++x) {
				int cx = inference8x8 ? (x >> 1) * 3 : x, cy = inference8x8 ? (y >> 1) * 3 : y;
				if (col_shape == SHAPE_16x16) cx = cy = 0;
				else if (col_shape == SHAPE_16x8) { cx = 0; cy = y & 2; }
				else if (col_shape == SHAPE_8x16) { cx = x & 2; cy = 0; }
				bool col_zero = false;
				if (col && col_short_term && col->n) {
					const ColBlk &cb = col->b[(size_t)(gy + cy) * w4 + gx + cx];
					col_zero = cb.ref == 0 && cb.mv[0] >= -1 && cb.mv[0] <= 1 && cb.mv[1] >= -1 && cb.mv[1] <= 1;
				}
				for (int l = 0; l < 2; ++l) {
					out.ref[y * 4 + x][l] = refs[l];
					bool z = zero || refs[l] < 0 || (refs[l] == 0 && col_zero);
					out.mv[y * 4 + x][l][0] = z ? 0 : mvp[l][0];
					out.mv[y * 4 + x][l][1] = z ? 0 : mvp[l][1];
				}
			}
	}

	// 8.4.1.2.3: temporal direct prediction
	void direct_temporal(int mbx, int mby, const ColPic *col, bool inference8x8, const RefPic *l0, int n0, const RefPic *l1,
	                     int cur_poc, DirectOut &out) const {
		const int gx = mbx * 4, gy = mby * 4;
		for (int y = 0; y < 4; ++y)
			for (int x = 0; x < 4; ++x) {
				int cx = inference8x8 ? (x >> 1) * 3 : x, cy = inference8x8 ? (y >> 1) * 3 : y;
				int i = y * 4 + x;
				out.ref[i][0] = 0; out.ref[i][1] = 0;
				out.mv[i][0][0] = out.mv[i][0][1] = out.mv[i][1][0] = out.mv[i][1][1] = 0;
				if (!col || !col->n) continue;
				const ColBlk &cb = col->b[(size_t)(gy + cy) * w4 + gx + cx];
				if (cb.ref < 0) continue;
				int r0 = 0;
				for (int k = 0; k < n0; ++k)
					if (l0[k].frame_idx >= 0 && l0[k].poc == cb.refpoc) { r0 = k; break; }
				out.ref[i][0] = r0;
				int td = l1[0].poc - l0[r0].poc, tb = cur_poc - l0[r0].poc;
				td = td < -128 ? -128 : (td > 127 ? 127 : td);
				tb = tb < -128 ? -128 : (tb > 127 ? 127 : tb);
				if (l0[r0].long_idx >= 0 || td == 0) {
					out.mv[i][0][0] = cb.mv[0]; out.mv[i][0][1] = cb.mv[1];
				} else {
					int tx = (16384 + (td < 0 ? -td : td) / 2) / td;
					int dsf = (tb * tx + 32) >> 6;
					dsf = dsf < -1024 ? -1024 : (dsf > 1023 ? 1023 : dsf);
					for (int k = 0; k < 2; ++k) {
						int m0 = (dsf * cb.mv[k] + 128) >> 8;
						out.mv[i][0][k] = m0;
						out.mv[i][1][k] = m0 - cb.mv[k];
					}
				}
			}
	}

	// --- CABAC ctxIdxInc derivations (9.3.3.1.1.x); open_mb must have been called ----------------
	inline const MbInfo *mb_left() const { return env.availL ? &mb[(size_t)env.mby * mb_w + env.mbx - 1] : nullptr; }
	inline const MbInfo *mb_top() const { return env.availT ? &mb[(size_t)(env.mby - 1) * mb_w + env.mbx] : nullptr; }
	int cabac_inc_skip() const {
		const MbInfo *l = mb_left(), *t = mb_top();
		return (l && l->kind != MB_SKIP) + (t && t->kind != MB_SKIP);
	}
	int cabac_inc_mbtype_i() const {
		const MbInfo *l = mb_left(), *t = mb_top();
		return (l && l->kind != MB_I4x4 && l->kind != MB_I8x8) + (t && t->kind != MB_I4x4 && t->kind != MB_I8x8);
	}
	int cabac_inc_mbtype_b() const {
		const MbInfo *l = mb_left(), *t = mb_top();
		return (l && !(l->cflags & MBF_DIRECT16)) + (t && !(t->cflags & MBF_DIRECT16));
	}
	int cabac_inc_chroma_mode() const {
		const MbInfo *l = mb_left(), *t = mb_top();
		return (l && kind_is_intra(l->kind) && l->kind != MB_IPCM && l->chroma_mode) +
		       (t && kind_is_intra(t->kind) && t->kind != MB_IPCM && t->chroma_mode);
	}
	int cabac_inc_t8x8() const {
		const MbInfo *l = mb_left(), *t = mb_top();
		return (l && l->t8x8) + (t && t->t8x8);
	}
	// bits = coded_block_pattern luma bins of the current MB decoded so far
	int cabac_inc_cbp_luma(int b8, int bits) const {
		const MbInfo *l = mb_left(), *t = mb_top();
		const int ca = l ? l->cbp : 0x0f, cb = t ? t->cbp : 0x0f;
		switch (b8) {
		case 0: return !(ca & 2) + 2 * !(cb & 4);
		case 1: return !(bits & 1) + 2 * !(cb & 8);
		case 2: return !(ca & 8) + 2 * !(bits & 1);
		default: return !(bits & 4) + 2 * !(bits & 2);
		}
	}
	int cabac_inc_cbp_chroma(int bin) const {
		const MbInfo *l = mb_left(), *t = mb_top();
		const int ca = l ? l->cbp >> 4 : 0, cb = t ? t->cbp >> 4 : 0;
		return bin == 0 ? (ca != 0) + 2 * (cb != 0) : (ca == 2) + 2 * (cb == 2);
	}
	int cabac_inc_ref(int list, int x4, int y4) const {
		const size_t g = (size_t)(env.mby * 4 + y4) * w4 + env.mbx * 4 + x4;
		const int a = (x4 > 0 || env.availL) ? refgt0[list][g - 1] : 0;
		const int b = (y4 > 0 || env.availT) ? refgt0[list][g - w4] : 0;
		return a + 2 * b;
	}
	int cabac_mvd_sum(int list, int x4, int y4, int comp) const {
		const size_t g = (size_t)(env.mby * 4 + y4) * w4 + env.mbx * 4 + x4;
		const int a = (x4 > 0 || env.availL) ? mvd[list][(g - 1) * 2 + comp] : 0;
		const int b = (y4 > 0 || env.availT) ? mvd[list][(g - w4) * 2 + comp] : 0;
		return a + b;
	}
	void set_mvd(int list, int x4, int y4, int pw, int ph, int dx, int dy) {
		dx = dx < 0 ? -dx : dx; dy = dy < 0 ? -dy : dy;
		const uint8_t ax = (uint8_t)(dx > 127 ? 127 : dx), ay = (uint8_t)(dy > 127 ? 127 : dy);
		for (int y = y4; y < y4 + ph; ++y) {
			size_t g = (size_t)(env.mby * 4 + y) * w4 + env.mbx * 4 + x4;
			for (int x = 0; x < pw; ++x) { mvd[list][(g + x) * 2] = ax; mvd[list][(g + x) * 2 + 1] = ay; }
		}
	}
	void set_refgt0(int list, int x4, int y4, int pw, int ph, int refidx) {
		for (int y = y4; y < y4 + ph; ++y)
			memset(&refgt0[list][(size_t)(env.mby * 4 + y) * w4 + env.mbx * 4 + x4], refidx > 0, pw);
	}
	// coded_block_flag: luma 4x4 block at (x4, y4), chroma AC block (x, y) of plane c, DC blocks
	int cabac_inc_cbf_luma(int x4, int y4, bool cur_intra) const {
		const int a = x4 > 0 ? env.nnz_cur[y4 * 4 + x4 - 1] != 0 : (env.nnz_left ? env.nnz_left[y4 * 4 + 3] != 0 : cur_intra);
		const int b = y4 > 0 ? env.nnz_cur[(y4 - 1) * 4 + x4] != 0 : (env.nnz_top ? env.nnz_top[12 + x4] != 0 : cur_intra);
		return a + 2 * b;
	}
	int cabac_inc_cbf_chroma_ac(int c, int x, int y, bool cur_intra) const {
		const int base = 16 + c * 4;
		const int a = x > 0 ? env.nnz_cur[base + y * 2] != 0 : (env.nnz_left ? env.nnz_left[base + y * 2 + 1] != 0 : cur_intra);
		const int b = y > 0 ? env.nnz_cur[base + x] != 0 : (env.nnz_top ? env.nnz_top[base + 2 + x] != 0 : cur_intra);
		return a + 2 * b;
	}
	int cabac_inc_cbf_dc(int flag, bool cur_intra) const {
		const MbInfo *l = mb_left(), *t = mb_top();
		const int a = l ? (l->cflags & flag) != 0 : cur_intra;
		const int b = t ? (t->cflags & flag) != 0 : cur_intra;
		return a + 2 * b;
	}

	// 8.3.1.1: predIntra4x4PredMode for the 4x4 block at raster (x4, y4) of the current MB.
	// constrained: constrained_intra_pred_flag (inter neighbours then count as unavailable).
	int predict_ipred(int mbx, int mby, int x4, int y4, int cur_slice, bool constrained) const {
		int m[2];
		for (int k = 0; k < 2; ++k) {
			int nx = x4 - (k == 0), ny = y4 - (k == 1);
			int nmbx = mbx, nmby = mby;
			if (nx < 0) { nx = 3; nmbx--; }
			if (ny < 0) { ny = 3; nmby--; }
			if (nmbx != mbx || nmby != mby) {
				if (!mb_avail(nmbx, nmby, cur_slice)) return 2;
				const MbInfo &o = info(nmbx, nmby);
				if (!kind_is_intra(o.kind) && constrained) return 2;
				m[k] = (o.kind == MB_I4x4 || o.kind == MB_I8x8) ? ipred[((size_t)nmby * mb_w + nmbx) * 16 + ny * 4 + nx] : 2;
			} else {
				m[k] = ipred[((size_t)mby * mb_w + mbx) * 16 + ny * 4 + nx];
			}
		}
		return m[0] < m[1] ? m[0] : m[1];
	}
};


}}  // namespace b2::h4
