// mpeg2_writer.cpp — TEST-ONLY MPEG-2 video elementary-stream writer (frame pictures, I/P/B).
// Like the H.264 writer it draws random but valid syntax: macroblock types, half-pel vectors
// that keep every referenced block inside the picture (the reference decoder does not clamp,
// SURVEY.md hazard 7), frame / field prediction and DCT (when `field` is set), skipped
// macroblocks, missing slices, quantiser changes, run/level codes of both tables and escapes.
// Coefficient amplitudes are bounded so that prediction + residual stays inside the reference's
// clip-table range (hazard 1).  VLC tables are the strings of m2dec_b200/csrc/mpeg2_tables.h;
// every stream is checked against the compiled reference decoder in tests/.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "bitwriter.h"
#include "mpeg2_tables.h"
#include "streamgen.h"

namespace m2t = b2::m2t;

namespace sg {
namespace {

struct Enc2 {
	VlcEnc mb_inc, type_i, type_p, type_b, cbp, motion, dc_luma, dc_chroma, dct[2];
	Enc2() {
		mb_inc.build(m2t::mb_addr_inc, B2_COUNT(m2t::mb_addr_inc));
		type_i.build(m2t::mb_type_i, B2_COUNT(m2t::mb_type_i));
		type_p.build(m2t::mb_type_p, B2_COUNT(m2t::mb_type_p));
		type_b.build(m2t::mb_type_b, B2_COUNT(m2t::mb_type_b));
		cbp.build(m2t::cbp, B2_COUNT(m2t::cbp));
		motion.build(m2t::motion_code, B2_COUNT(m2t::motion_code));
		dc_luma.build(m2t::dc_size_luma, B2_COUNT(m2t::dc_size_luma));
		dc_chroma.build(m2t::dc_size_chroma, B2_COUNT(m2t::dc_size_chroma));
		dct[0].build(m2t::dct_table0, B2_COUNT(m2t::dct_table0));
		dct[1].build(m2t::dct_table1, B2_COUNT(m2t::dct_table1));
	}
};

enum { I_PIC = 1, P_PIC = 2, B_PIC = 3 };

class Writer2 {
public:
	explicit Writer2(const Mpeg2Opts &o) : o_(o), rng_((uint64_t)o.seed) {
		W_ = o.width ? o.width : o.mbw * 16;
		H_ = o.height ? o.height : o.mbh * 16;
		mbw_ = (W_ + 15) / 16;
		mbh_ = (H_ + 15) / 16;
	}

	std::vector<uint8_t> run() {
		sequence_header();
		// display order -> coded order: anchors first, then the B pictures that precede them
		int m = o_.bframes + 1;
		int done = 0;
		bool first = true;
		while (done < o_.frames) {
			if (first) { gop_header(); picture(I_PIC, 0); done = 1; first = false; continue; }
			int remaining = o_.frames - done;
			int nb = remaining > m ? o_.bframes : remaining - 1;
			bool intra = o_.gop > 0 && ((done + nb) % o_.gop) == 0;
			picture(intra ? I_PIC : P_PIC, done + nb);
			for (int b = 0; b < nb; ++b) picture(B_PIC, done + b);
			done += nb + 1;
		}
		static const uint8_t end_code[4] = {0, 0, 1, 0xb7};
		out_.insert(out_.end(), end_code, end_code + 4);
		return out_;
	}

private:
	void start_code(int code) {
		out_.push_back(0); out_.push_back(0); out_.push_back(1); out_.push_back((uint8_t)code);
	}
	void flush(BitWriter &bw) {
		bw.align_zero();
		out_.insert(out_.end(), bw.bytes().begin(), bw.bytes().end());
	}

	void sequence_header() {
		BitWriter bw;
		start_code(0xb3);
		bw.put(W_ & 0xfff, 12); bw.put(H_ & 0xfff, 12);
		bw.put(1, 4); bw.put(3, 4);
		bw.put(0x3ffff, 18); bw.put1(1); bw.put(112, 10); bw.put1(0);
		bw.put1(0); bw.put1(0);            // default quantiser matrices
		flush(bw);
		BitWriter ex;
		start_code(0xb5);
		ex.put(1, 4);
		ex.put(mbw_ * mbh_ > 1620 ? 0x44 : 0x48, 8);
		ex.put1(o_.field ? 0 : 1);         // progressive_sequence
		ex.put(1, 2);                      // 4:2:0
		ex.put(W_ >> 12, 2); ex.put(H_ >> 12, 2);
		ex.put(0, 12); ex.put1(1); ex.put(0, 8); ex.put1(0); ex.put(0, 2); ex.put(0, 5);
		flush(ex);
	}

	void gop_header() {
		BitWriter bw;
		start_code(0xb8);
		bw.put(0, 25); bw.put1(1); bw.put1(0);
		flush(bw);
	}

	// ---- picture ---------------------------------------------------------------------------------
	void picture(int type, int temporal_ref) {
		type_ = type;
		BitWriter bw;
		start_code(0x00);
		bw.put(temporal_ref & 0x3ff, 10);
		bw.put(type, 3);
		bw.put(0xffff, 16);
		if (type != I_PIC) { bw.put1(0); bw.put(7, 3); }
		if (type == B_PIC) { bw.put1(0); bw.put(7, 3); }
		bw.put1(0);
		flush(bw);
		f_code_ = 3;                       // vectors within +-64 half samples
		frame_pred_frame_dct_ = o_.field ? 0 : 1;
		BitWriter ex;
		start_code(0xb5);
		ex.put(8, 4);
		ex.put(type != I_PIC ? f_code_ : 15, 4); ex.put(type != I_PIC ? f_code_ : 15, 4);
		ex.put(type == B_PIC ? f_code_ : 15, 4); ex.put(type == B_PIC ? f_code_ : 15, 4);
		ex.put(o_.dc_precision, 2);
		ex.put(3, 2);                      // frame picture
		ex.put1(1);                        // top_field_first
		ex.put1(frame_pred_frame_dct_);
		ex.put1(0);                        // concealment_motion_vectors
		ex.put1(o_.q_scale_type);
		ex.put1(o_.intra_vlc);
		ex.put1(o_.alt_scan);
		ex.put1(0); ex.put1(o_.field ? 0 : 1); ex.put1(o_.field ? 0 : 1); ex.put1(0);
		flush(ex);
		prev_type_ = 0;
		for (int row = 0; row < mbh_; ++row) {
			// a missing slice is copied from the forward reference by the decoder; never row 0
			// (which starts the picture) and not in I pictures of the very first GOP
			if (row > 0 && row < mbh_ - 1 && type == P_PIC && rng_.chance(o_.skip_pct / 3)) continue;
			slice(row);
		}
	}

	int qscale_of(int code) const { return o_.q_scale_type ? (code ? m2t::non_linear_qscale[code] : 1) : code * 2; }

	void slice(int row) {
		BitWriter bw;
		start_code(row + 1);
		qcode_ = o_.qscale;
		bw.put(qcode_, 5);
		bw.put1(0);
		reset_dc();
		memset(pmv_, 0, sizeof(pmv_));
		int x = 0, pending_inc = 1;
		bool first = true;
		while (x < mbw_) {
			bool last = x == mbw_ - 1;
			// skipped macroblocks: not first / last of the slice; B: previous MB must not be intra and its
			// vectors must keep this block inside the picture
			if (!first && !last && type_ != I_PIC && rng_.chance(o_.skip_pct) && can_skip(x, row)) {
				if (type_ == P_PIC) { reset_dc(); memset(pmv_, 0, sizeof(pmv_)); }
				pending_inc++;
				x++;
				continue;
			}
			while (pending_inc > 33) { enc_.mb_inc.put(bw, 0); pending_inc -= 33; }
			enc_.mb_inc.put(bw, pending_inc);
			pending_inc = 1;
			macroblock(bw, x, row);
			first = false;
			x++;
		}
		flush(bw);
	}

	bool can_skip(int x, int row) const {
		if (type_ == P_PIC) return true;
		if (prev_type_ & m2t::MT_INTRA) return false;
		for (int s = 0; s < 2; ++s)
			if (prev_type_ & (s ? m2t::MT_BWD : m2t::MT_FWD))
				if (!vector_ok(x, row, pmv_[s][0][0], pmv_[s][0][1], 16)) return false;
		return (prev_type_ & (m2t::MT_FWD | m2t::MT_BWD)) != 0;
	}

	// does the 16 x h block of MB (x, row) displaced by the half-pel vector stay inside the picture?
	bool vector_ok(int x, int row, int mvx, int mvy, int h) const {
		int px = x * 16 + (mvx >> 1), py = row * 16 + (mvy >> 1);
		return px >= 0 && py >= 0 && px + 16 + (mvx & 1) <= mbw_ * 16 && py + h + (mvy & 1) <= mbh_ * 16;
	}

	void reset_dc() { dc_pred_[0] = dc_pred_[1] = dc_pred_[2] = 1 << (o_.dc_precision + 7); }

	// ---- motion vectors (mirror of the decoder's PMV handling) ---------------------------------------
	void put_mv_component(BitWriter &bw, int16_t *pred, int value, int shift) {
		int r_size = f_code_ - 1, f = 1 << r_size;
		int delta = value - (*pred >> shift);
		int limit = 16 << r_size;
		if (delta < -limit) delta += 2 * limit;
		if (delta >= limit) delta -= 2 * limit;
		if (delta == 0) {
			enc_.motion.put(bw, 0);
		} else {
			int a = abs(delta) + f - 1;
			int code = a / f, res = a % f;
			enc_.motion.put(bw, delta > 0 ? code : -code);
			if (r_size) bw.put(res, r_size);
		}
		*pred = (int16_t)(value << shift);
	}

	// draw a vector for MB (x, row); frame vector (field = 0) or field vector for a 16x8 field block
	void draw_mv(int x, int row, bool field, int sel, int mv[2]) {
		for (int tries = 0; tries < 64; ++tries) {
			mv[0] = rng_.range(-o_.mv_range, o_.mv_range);
			mv[1] = rng_.range(-o_.mv_range, o_.mv_range);
			if (rng_.chance(20)) mv[0] = 0;
			if (rng_.chance(20)) mv[1] = 0;
			if (field) {
				// field block: rows (row*16 + sel + 2*(j + (mvy >> 1))), j < 8, plus one more field row when half-pel
				int px = x * 16 + (mv[0] >> 1), top = row * 16 + sel + 2 * (mv[1] >> 1);
				int bottom = top + 2 * (7 + (mv[1] & 1));
				if (px >= 0 && px + 16 + (mv[0] & 1) <= mbw_ * 16 && top >= 0 && bottom < mbh_ * 16) return;
			} else if (vector_ok(x, row, mv[0], mv[1], 16)) {
				return;
			}
		}
		mv[0] = mv[1] = 0;
	}

	// ---- coefficients ---------------------------------------------------------------------------------
	// run/level pairs for one block; first_pos = 0 (non-intra) or 1 (intra, after DC)
	void put_coefficients(BitWriter &bw, bool intra, int first_pos) {
		const VlcEnc &tab = enc_.dct[intra ? o_.intra_vlc : 0];
		int n = intra ? rng_.range(0, 3) : rng_.range(1, 5);
		int pos = first_pos;
		bool first = !intra;
		for (int i = 0; i < n; ++i) {
			int run = rng_.chance(60) ? 0 : rng_.range(1, intra ? 3 : 6);
			if (rng_.chance(3)) run = rng_.range(7, 20);
			if (pos + run > (intra ? 12 : 40)) break;      // stay at low frequencies: bounded amplitude
			int level = rng_.chance(70) ? 1 : rng_.range(2, 3);
			if (rng_.chance(2)) level = rng_.range(4, 9);
			bool neg = rng_.chance(50);
			pos += run;
			int key = run * 64 + level;
			if (first && run == 0 && level == 1) {
				bw.put1(1); bw.put1(neg);                   // "1s": first coefficient of a non-intra block
			} else if (level <= 40 && tab.has(key)) {
				tab.put(bw, key);
				bw.put1(neg);
			} else {
				tab.put(bw, m2t::DCT_ESC);
				bw.put(run, 6);
				bw.put((neg ? 4096 - level : level) & 0xfff, 12);
			}
			first = false;
			pos++;
		}
		if (first) { bw.put1(1); bw.put1(rng_.chance(50)); }   // a coded non-intra block needs one coefficient
		tab.put(bw, m2t::DCT_EOB);
	}

	void put_intra_dc(BitWriter &bw, int comp) {
		// keep the DC around mid grey so that DC + AC stays far from the clip-table limits
		int maxv = 255 << o_.dc_precision;
		int target = rng_.range(maxv / 4, maxv * 3 / 4);
		int diff = target - dc_pred_[comp];
		if (rng_.chance(40)) diff = 0;
		int size = 0;
		for (int a = abs(diff); a; a >>= 1) size++;
		(comp == 0 ? enc_.dc_luma : enc_.dc_chroma).put(bw, size);
		if (size) {
			bw.put(diff > 0 ? diff : diff + (1 << size) - 1, size);
			dc_pred_[comp] += diff;
		}
	}

	// ---- macroblock -------------------------------------------------------------------------------------
	void macroblock(BitWriter &bw, int x, int row) {
		bool prev_intra = (prev_type_ & m2t::MT_INTRA) != 0;
		int type;
		bool want_intra = type_ == I_PIC || rng_.chance(o_.intra_pct);
		bool quant = rng_.chance(10);
		if (want_intra) {
			type = m2t::MT_INTRA;
		} else if (type_ == P_PIC) {
			int r = rng_.range(0, 99);
			type = r < 55 ? (m2t::MT_FWD | m2t::MT_PATTERN) : (r < 80 ? m2t::MT_FWD : m2t::MT_PATTERN);
		} else {
			int d = rng_.range(0, 2);
			type = d == 0 ? m2t::MT_FWD : (d == 1 ? m2t::MT_BWD : (m2t::MT_FWD | m2t::MT_BWD));
			if (rng_.chance(o_.coded_pct)) type |= m2t::MT_PATTERN;
		}
		if (quant && (type & (m2t::MT_INTRA | m2t::MT_PATTERN))) {
			// MB_QUANT exists only in some combinations (Tables B.3 / B.4)
			bool ok = (type & m2t::MT_INTRA) || (type_ == P_PIC) || (type & m2t::MT_PATTERN);
			if (ok) type |= m2t::MT_QUANT;
		}
		const VlcEnc &tv = type_ == I_PIC ? enc_.type_i : (type_ == P_PIC ? enc_.type_p : enc_.type_b);
		if (!tv.has(type)) type &= ~m2t::MT_QUANT;
		tv.put(bw, type);
		prev_type_ = type;
		bool mc = (type & (m2t::MT_FWD | m2t::MT_BWD)) != 0;
		bool field_mc = false;
		if (mc && !frame_pred_frame_dct_) {
			field_mc = rng_.chance(40);
			bw.put(field_mc ? 1 : 2, 2);       // frame_motion_type: 01 field-based, 10 frame-based
		}
		if (!frame_pred_frame_dct_ && (type & (m2t::MT_INTRA | m2t::MT_PATTERN))) bw.put1(rng_.chance(40));
		if (type & m2t::MT_QUANT) {
			qcode_ = rng_.range(o_.qscale > 2 ? o_.qscale - 2 : 1, o_.qscale + 2 > 31 ? 31 : o_.qscale + 2);
			bw.put(qcode_, 5);
		}
		if (type & m2t::MT_INTRA) {
			if (!prev_intra) reset_dc();
			for (int k = 0; k < 6; ++k) {
				put_intra_dc(bw, k < 4 ? 0 : k - 3);
				put_coefficients(bw, true, 1);
			}
			return;
		}
		if (prev_intra) memset(pmv_, 0, sizeof(pmv_));
		if (mc) {
			for (int s = 0; s < 2; ++s) {
				if (!(type & (s ? m2t::MT_BWD : m2t::MT_FWD))) continue;
				if (!field_mc) {
					int mv[2];
					draw_mv(x, row, false, 0, mv);
					put_mv_component(bw, &pmv_[s][0][0], mv[0], 0);
					put_mv_component(bw, &pmv_[s][0][1], mv[1], 0);
					pmv_[s][1][0] = pmv_[s][0][0];
					pmv_[s][1][1] = pmv_[s][0][1];
				} else {
					for (int f = 0; f < 2; ++f) {
						int sel = rng_.range(0, 1), mv[2];
						draw_mv(x, row, true, sel, mv);
						bw.put1(sel);
						put_mv_component(bw, &pmv_[s][f][0], mv[0], 0);
						put_mv_component(bw, &pmv_[s][f][1], mv[1], 1);
					}
				}
			}
		} else {
			reset_dc();
			memset(pmv_, 0, sizeof(pmv_));
		}
		if (type & m2t::MT_PATTERN) {
			int cbp = rng_.range(1, 63);
			enc_.cbp.put(bw, cbp);
			for (int k = 0; k < 6; ++k)
				if (cbp & (1 << (5 - k))) put_coefficients(bw, false, 0);
		}
	}

	Mpeg2Opts o_;
	Rng rng_;
	Enc2 enc_;
	std::vector<uint8_t> out_;
	int W_, H_, mbw_, mbh_;
	int type_ = 0, f_code_ = 3, frame_pred_frame_dct_ = 1, qcode_ = 6, prev_type_ = 0;
	int dc_pred_[3] = {128, 128, 128};
	int16_t pmv_[2][2][2] = {};
};

}  // namespace

std::vector<uint8_t> write_mpeg2(const Mpeg2Opts &o) {
	Writer2 w(o);
	return w.run();
}

}  // namespace sg
