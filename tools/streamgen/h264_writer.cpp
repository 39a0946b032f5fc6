// h264_writer.cpp — TEST-ONLY H.264 elementary-stream writer (not an encoder: it does not look
// at pictures).  It draws random but *valid* syntax — macroblock types, partitions, motion
// vectors, intra modes, coefficient levels, QP changes, slices — and entropy-codes it with
// CAVLC or CABAC, so that every reconstruction path of a decoder is exercised with known-legal input.
// Streams stay inside the region where the reference decoder is well defined (SURVEY.md §8a
// hazards 1, 2, 5, 10): bounded residual amplitude, beta_offset >= alpha_offset,
// seq_parameter_set_id 0, disable_deblocking_filter_idc in {0, 1}.
//
// Neighbour-dependent syntax (mvd, prev_intra4x4_pred_mode, coeff_token nC, P_Skip motion)
// uses the same context class as the decoder front end (m2dec_b200/csrc/h264_common.h); that
// a shared mistake cannot hide is guaranteed by decoding every stream with the compiled
// reference (tests/, oracle/_ref).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <algorithm>
#include <vector>
#include "bitwriter.h"
#include "cabac_enc.h"
#include "h264_cavlc_tables.h"
#include "h264_common.h"
#include "streamgen.h"

using namespace b2::h4;
namespace h4t = b2::h4t;

namespace sg {
namespace {

struct Enc {
	VlcEnc coeff_token[5], total_zeros[15], total_zeros_cdc[3], run_before[7];
	uint8_t code_from_cbp[2][48];
	Enc() {
		for (int i = 0; i < 5; ++i) coeff_token[i].build(h4t::coeff_token_tabs[i].codes, h4t::coeff_token_tabs[i].n);
		for (int i = 0; i < 15; ++i) total_zeros[i].build(h4t::total_zeros_tabs[i].codes, h4t::total_zeros_tabs[i].n);
		for (int i = 0; i < 3; ++i) total_zeros_cdc[i].build(h4t::total_zeros_cdc_tabs[i].codes, h4t::total_zeros_cdc_tabs[i].n);
		for (int i = 0; i < 7; ++i) run_before[i].build(h4t::run_before_tabs[i].codes, h4t::run_before_tabs[i].n);
		for (int t = 0; t < 2; ++t)
			for (int c = 0; c < 48; ++c) code_from_cbp[t][cbp_from_code[t][c]] = (uint8_t)c;
	}
};

// dequantisation scale of position (x, y) of a 4x4 block (flat matrices)
int level_scale(int qp, int x, int y) {
	static const int na[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};
	int cls = ((x & 1) == 0 && (y & 1) == 0) ? 0 : (((x & 1) && (y & 1)) ? 1 : 2);
	return na[qp % 6][cls] << (qp / 6);
}

class Writer {
public:
	Writer(const H264Opts &o) : o_(o), rng_((uint64_t)o.seed) {
		// transform_size_8x8_flag and direct prediction: the reference reads the flag for
		// B_Direct_16x16 even without direct_8x8_inference_flag (h264.cpp:9420); stay where it
		// agrees with the standard
		if (o_.transform8x8 && o_.bframes > 0) o_.direct_8x8_inference = 1;
		ctx_.resize(o.mbw, o.mbh);
		ctx_.cabac = o_.cabac != 0;
	}

	std::vector<uint8_t> run() {
		write_sps();
		write_pps();
		// display index d -> POC 2d.  Coded order: anchor first, then the B pictures in front of it.
		int frame_num = 0, since_idr = 0, done = 0;
		while (done < o_.frames) {
			bool idr = done == 0 || (o_.bframes == 0 && (done % o_.gop) == 0);
			int nb = 0;
			if (!idr && o_.bframes > 0) {
				int remaining = o_.frames - done;
				nb = remaining > o_.bframes ? o_.bframes : remaining - 1;
			}
			if (idr) { frame_num = 0; since_idr = 0; idr_id_ = (idr_id_ + 1) & 0xff; poc_base_ = done; refs_.clear(); }
			picture(idr ? 2 : 0, idr, frame_num, since_idr, done + nb, true);
			frame_num = (frame_num + 1) & 15;
			since_idr++;
			for (int b = 0; b < nb; ++b) picture(1, false, frame_num, since_idr, done + b, false);
			done += nb + 1;
		}
		return out_;
	}

private:
	// ---- NAL ------------------------------------------------------------------------------------
	void emit_nal(int ref_idc, int type, BitWriter &bw) {
		static const uint8_t sc[4] = {0, 0, 0, 1};
		out_.insert(out_.end(), sc, sc + 4);
		out_.push_back((uint8_t)((ref_idc << 5) | type));
		int zeros = 0;
		for (uint8_t b : bw.bytes()) {
			if (zeros >= 2 && b <= 3) { out_.push_back(3); zeros = 0; }
			out_.push_back(b);
			zeros = b ? 0 : zeros + 1;
		}
	}

	void write_sps() {
		BitWriter bw;
		bw.put(o_.profile, 8);
		bw.put(0, 8);
		bw.put(40, 8);               // level 4.0
		bw.ue(0);                    // seq_parameter_set_id (hazard 10: must be 0)
		if (o_.profile == 100) { bw.ue(1); bw.ue(0); bw.ue(0); bw.put1(0); bw.put1(0); }   // 4:2:0, 8 bit, no scaling matrices
		bw.ue(0);                    // log2_max_frame_num_minus4
		if (o_.bframes > 0) { bw.ue(0); bw.ue(4); }   // pic_order_cnt_type 0, log2_max_pic_order_cnt_lsb = 8
		else bw.ue(2);               // pic_order_cnt_type 2
		bw.ue(o_.refs);              // num_ref_frames
		bw.put1(0);                  // gaps_in_frame_num_value_allowed_flag
		bw.ue(o_.mbw - 1);
		bw.ue(o_.mbh - 1);
		bw.put1(1);                  // frame_mbs_only_flag
		bw.put1(o_.direct_8x8_inference);   // direct_8x8_inference_flag
		if (o_.crop_bottom || o_.crop_right) {
			bw.put1(1);
			bw.ue(0); bw.ue(o_.crop_right / 2); bw.ue(0); bw.ue(o_.crop_bottom / 2);
		} else {
			bw.put1(0);
		}
		bw.put1(0);                  // vui_parameters_present_flag
		bw.rbsp_trailing();
		emit_nal(3, 7, bw);
	}

	void write_pps() {
		BitWriter bw;
		bw.ue(0); bw.ue(0);
		bw.put1(o_.cabac);           // entropy_coding_mode_flag
		bw.put1(0);
		bw.ue(0);                    // num_slice_groups_minus1
		bw.ue(o_.refs - 1);          // num_ref_idx_l0_default_active_minus1
		bw.ue(o_.refs - 1);          // num_ref_idx_l1_default_active_minus1
		bw.put1(o_.weighted_pred);   // weighted_pred_flag
		bw.put(o_.weighted_bipred, 2);
		bw.se(o_.qp - 26);           // pic_init_qp_minus26
		bw.se(0);
		bw.se(o_.chroma_qp_offset);
		bw.put1(1);                  // deblocking_filter_control_present_flag
		bw.put1(o_.constrained_intra);
		bw.put1(0);
		if (o_.transform8x8) {
			bw.put1(1);              // transform_8x8_mode_flag
			bw.put1(0);              // pic_scaling_matrix_present_flag
			bw.se(o_.chroma_qp_offset);   // second_chroma_qp_index_offset
		}
		bw.rbsp_trailing();
		emit_nal(3, 8, bw);
	}

	// ---- picture / slice ---------------------------------------------------------------------------
	// slice_kind: 0 P, 1 B, 2 I.  display = display index (POC = 2 * (display - poc_base_)).
	void picture(int slice_kind, bool idr, int frame_num, int since_idr, int display, bool is_ref) {
		(void)since_idr;
		ctx_.begin_picture();
		int total = o_.mbw * o_.mbh;
		int nslices = o_.slices < 1 ? 1 : o_.slices;
		int per = (total + nslices - 1) / nslices;
		qp_ = o_.qp;
		kind_ = slice_kind;
		cur_poc_ = 2 * (display - poc_base_);
		build_lists();
		for (int s = 0, first = 0; first < total; ++s, first += per) {
			int last = first + per < total ? first + per : total;
			slice_num_ = s;
			BitWriter bw;
			bw.ue(first);
			bw.ue(slice_kind == 2 ? 7 : (slice_kind == 0 ? 5 : 6));   // all slices of the picture have this type
			bw.ue(0);                // pic_parameter_set_id
			bw.put(frame_num, 4);
			if (idr) bw.ue(idr_id_);
			if (o_.bframes > 0) bw.put(cur_poc_ & 255, 8);            // pic_order_cnt_lsb
			if (slice_kind == 1) bw.put1(o_.direct_spatial);
			if (slice_kind != 2) {
				nref_[0] = (int)list_[0].size();
				nref_[1] = (int)list_[1].size();
				bool override_flag = nref_[0] != o_.refs || (slice_kind == 1 && nref_[1] != o_.refs);
				bw.put1(override_flag);
				if (override_flag) { bw.ue(nref_[0] - 1); if (slice_kind == 1) bw.ue(nref_[1] - 1); }
				bw.put1(0);          // ref_pic_list_modification_flag_l0
				if (slice_kind == 1) bw.put1(0);
				if ((slice_kind == 0 && o_.weighted_pred) || (slice_kind == 1 && o_.weighted_bipred == 1)) weight_table(bw, slice_kind == 1);
			}
			if (is_ref) {
				if (idr) { bw.put1(0); bw.put1(0); }   // no_output_of_prior_pics, long_term_reference_flag
				else bw.put1(0);                        // adaptive_ref_pic_marking_mode_flag
			}
			int init_idc = 0;
			if (o_.cabac && slice_kind != 2) { init_idc = o_.cabac_init_idc < 0 ? rng_.range(0, 2) : o_.cabac_init_idc; bw.ue(init_idc); }
			int slice_qp = o_.qp + (o_.vary_qp ? rng_.range(-2, 2) : 0);
			bw.se(slice_qp - o_.qp);
			qp_ = slice_qp;
			int idc = o_.deblock ? ((o_.deblock == 2 && (s & 1)) ? 1 : 0) : 1;
			bw.ue(idc);
			if (idc != 1) { bw.se(o_.alpha_div2); bw.se(o_.beta_div2); }
			if (o_.cabac) {
				while (!bw.aligned()) bw.put1(1);   // cabac_alignment_one_bit
				cs_.e.m.init(slice_qp, slice_kind == 2 ? 0 : 1 + init_idc);
				cs_.e.start(&bw);
				last_dqp_ = 0;
				slice_data_cabac(bw, slice_kind == 2, first, last);
				bw.align_zero();     // the flush wrote rbsp_stop_one_bit
			} else {
				slice_data(bw, slice_kind == 2, first, last);
				bw.rbsp_trailing();
			}
			emit_nal(idr ? 3 : (is_ref ? 2 : 0), idr ? 5 : 1, bw);
		}
		if (is_ref) {
			// sliding window marking (8.2.5.3) and the co-located motion of this picture
			int id = next_pic_id_++;
			for (auto &r : refs_) { r.frame_num_wrap = r.rp.frame_num > frame_num ? r.rp.frame_num - 16 : r.rp.frame_num; }
			if ((int)refs_.size() >= o_.refs) {
				int victim = 0;
				for (size_t i = 1; i < refs_.size(); ++i) if (refs_[i].frame_num_wrap < refs_[victim].frame_num_wrap) victim = (int)i;
				refs_.erase(refs_.begin() + victim);
			}
			WRef nr;
			nr.rp = RefPic{id, frame_num, -1, cur_poc_};
			nr.frame_num_wrap = frame_num;
			ctx_.store_col(nr.col, [&](int f) { for (auto &r : refs_) if (r.rp.frame_idx == f) return r.rp.poc; return 0; });
			refs_.push_back(std::move(nr));
		}
	}

	// reference picture lists as the decoder will build them (8.2.4.2.1 / 8.2.4.2.3), no modification
	void build_lists() {
		list_[0].clear(); list_[1].clear();
		if (kind_ == 2) return;
		std::vector<const WRef *> v;
		for (auto &r : refs_) v.push_back(&r);
		if (kind_ == 0) {
			std::sort(v.begin(), v.end(), [](const WRef *a, const WRef *b) { return a->frame_num_wrap > b->frame_num_wrap; });
			for (auto *r : v) list_[0].push_back(r);
		} else {
			std::vector<const WRef *> before, after;
			for (auto *r : v) (r->rp.poc < cur_poc_ ? before : after).push_back(r);
			std::sort(before.begin(), before.end(), [](const WRef *a, const WRef *b) { return a->rp.poc > b->rp.poc; });
			std::sort(after.begin(), after.end(), [](const WRef *a, const WRef *b) { return a->rp.poc < b->rp.poc; });
			list_[0] = before; list_[0].insert(list_[0].end(), after.begin(), after.end());
			list_[1] = after; list_[1].insert(list_[1].end(), before.begin(), before.end());
			if (list_[1].size() > 1 && list_[0] == list_[1]) std::swap(list_[1][0], list_[1][1]);
		}
		// A P picture that is full of references evicts its oldest one when it is marked; the B pictures
		// that follow could then not map a co-located vector that points there (temporal direct), so
		// such P pictures leave the oldest reference unused.
		if (kind_ == 0 && o_.bframes > 0 && (int)list_[0].size() >= o_.refs && o_.refs > 1) list_[0].resize(o_.refs - 1);
		for (int l = 0; l < 2; ++l) {
			if ((int)list_[l].size() > o_.refs) list_[l].resize(o_.refs);
			rp_[l].clear();
			for (auto *r : list_[l]) rp_[l].push_back(r->rp);
		}
	}

	// explicit weights: log2 denominators <= 6, weights <= 64 and small offsets (hazard 4)
	void weight_table(BitWriter &bw, bool bslice) {
		int dl = rng_.range(2, 6), dc = rng_.range(2, 6);
		bw.ue(dl); bw.ue(dc);
		for (int l = 0; l < (bslice ? 2 : 1); ++l)
			for (int i = 0; i < nref_[l]; ++i) {
				bool fl = rng_.chance(70), fc = rng_.chance(70);
				bw.put1(fl);
				// w0 + w1 <= 128 keeps the reference's 16-bit saturating bi-prediction sum exact (hazard 4)
				auto weight = [&](int d) { int w = (1 << d) + rng_.range(-(1 << d) / 4, (1 << d) / 4); return w > 64 ? 64 : w; };
				if (fl) { bw.se(weight(dl)); bw.se(rng_.range(-8, 8)); }
				bw.put1(fc);
				if (fc) for (int c = 0; c < 2; ++c) { bw.se(weight(dc)); bw.se(rng_.range(-8, 8)); }
			}
	}

	void slice_data(BitWriter &bw, bool idr, int first, int last) {
		int skip_run = 0;
		for (int addr = first; addr < last; ++addr) {
			int mbx = addr % o_.mbw, mby = addr / o_.mbw;
			ctx_.open_mb(mbx, mby, slice_num_);
			if (!idr && rng_.chance(o_.skip_pct)) {
				skip_motion(mbx, mby);
				skip_run++;
				continue;
			}
			if (!idr) { bw.ue(skip_run); skip_run = 0; }
			if (idr || rng_.chance(o_.intra_pct)) intra_mb(bw, mbx, mby, idr);
			else if (kind_ == 1) b_mb(bw, mbx, mby);
			else inter_mb(bw, mbx, mby);
		}
		if (!idr && skip_run) bw.ue(skip_run);
	}

	void skip_motion(int mbx, int mby) {
		if (kind_ == 1) {
			DirectOut d;
			derive_direct(mbx, mby, d);
			set_direct(mbx, mby, d, 0xf);
		} else {
			int16_t mv[2];
			ctx_.pskip_mv(mbx, mby, slice_num_, mv);
			ctx_.set_motion(0, mbx, mby, 0, 0, 4, 4, 0, pic_of(0, 0), mv[0], mv[1]);
			ctx_.set_motion(1, mbx, mby, 0, 0, 4, 4, -1, -1, 0, 0);
		}
		MbInfo &info = ctx_.info(mbx, mby);
		info.kind = MB_SKIP;
		info.cbp = 0;
		if (kind_ == 1) info.cflags |= MBF_DIRECT16;
		memset(ctx_.nnz_of(mbx, mby), 0, 24);
	}

	void slice_data_cabac(BitWriter &bw, bool idr, int first, int last) {
		for (int addr = first; addr < last; ++addr) {
			int mbx = addr % o_.mbw, mby = addr / o_.mbw;
			ctx_.open_mb(mbx, mby, slice_num_);
			bool skip = !idr && rng_.chance(o_.skip_pct);
			if (!idr) cs_.mb_skip(kind_ == 1, ctx_.cabac_inc_skip(), skip);
			if (skip) {
				skip_motion(mbx, mby);
				last_dqp_ = 0;
			} else if (idr || rng_.chance(o_.intra_pct)) intra_mb(bw, mbx, mby, idr);
			else if (kind_ == 1) b_mb(bw, mbx, mby);
			else inter_mb(bw, mbx, mby);
			cs_.end_of_slice(addr == last - 1);
		}
	}

	// ---- syntax elements, either entropy coder ----------------------------------------------------------
	void put_mb_type(BitWriter &bw, int v) {
		if (!o_.cabac) { bw.ue(v); return; }
		if (kind_ == 2) cs_.mb_type_intra(3, true, ctx_.cabac_inc_mbtype_i(), v);
		else if (kind_ == 0) cs_.mb_type_p(v);
		else cs_.mb_type_b(ctx_.cabac_inc_mbtype_b(), v);
	}
	void put_sub_mb_type(BitWriter &bw, int v) {
		if (!o_.cabac) { bw.ue(v); return; }
		if (kind_ == 1) cs_.sub_mb_type_b(v); else cs_.sub_mb_type_p(v);
	}
	void put_ref(BitWriter &bw, int l, int x4, int y4, int pw, int ph, int r) {
		if (nref_[l] > 1) {
			if (o_.cabac) cs_.ref_idx(ctx_.cabac_inc_ref(l, x4, y4), r);
			else if (nref_[l] == 2) bw.put1(r ^ 1);
			else bw.ue(r);
		}
		if (o_.cabac) ctx_.set_refgt0(l, x4, y4, pw, ph, r);
	}
	void put_mvd(BitWriter &bw, int l, int x4, int y4, int pw, int ph, int dx, int dy) {
		if (!o_.cabac) { bw.se(dx); bw.se(dy); return; }
		cs_.mvd(40, ctx_.cabac_mvd_sum(l, x4, y4, 0), dx);
		cs_.mvd(47, ctx_.cabac_mvd_sum(l, x4, y4, 1), dy);
		ctx_.set_mvd(l, x4, y4, pw, ph, dx, dy);
	}
	void put_cbp(BitWriter &bw, bool intra, int cbp) {
		if (o_.cabac) cs_.cbp(ctx_, cbp);
		else bw.ue(enc_.code_from_cbp[intra ? 0 : 1][cbp]);
	}
	void put_qp_delta(BitWriter &bw, int d) {
		if (o_.cabac) cs_.qp_delta(last_dqp_ != 0, d);
		else bw.se(d);
		last_dqp_ = d;
	}
	void put_chroma_mode(BitWriter &bw, int v) {
		if (o_.cabac) cs_.chroma_mode(ctx_.cabac_inc_chroma_mode(), v);
		else bw.ue(v);
	}
	void put_ipred(BitWriter &bw, int pred, int mode) {
		if (o_.cabac) { cs_.ipred_mode(pred, mode); return; }
		if (mode == pred) bw.put1(1);
		else { bw.put1(0); bw.put(mode < pred ? mode : mode - 1, 3); }
	}
	void put_t8x8(int v) { cs_.transform8x8(ctx_.cabac_inc_t8x8(), v); }

	// ---- residual generation / coding ---------------------------------------------------------------
	// Draw up to `maxc` coefficients in scan order for a 4x4 block whose dequantised L1 amplitude
	// (sum |level| * scale / 64) must stay below *budget; returns the number of non-zero levels.
	int draw_block(int16_t *lv, int maxc, int first_scan, int qp, int *budget) {
		memset(lv, 0, sizeof(int16_t) * 16);
		if (o_.dense_pct && rng_.chance(o_.dense_pct)) {
			// nearly full block (maxc - 3 .. maxc coefficients of magnitude 1-2): reaches the long coeff_token
			// codes and the total_zeros tables of 13-15 coefficients; draws nothing when the option is off
			int holes = rng_.range(0, 3), cnt = 0;
			for (int pos = 0; pos < maxc; ++pos) lv[pos] = (int16_t)((rng_.chance(85) ? 1 : 2) * (rng_.chance(50) ? 1 : -1));
			for (int i = 0; i < holes; ++i) lv[rng_.range(0, maxc - 1)] = 0;
			for (int pos = 0; pos < maxc; ++pos) cnt += lv[pos] != 0;
			*budget = 0;
			return cnt;
		}
		int n = 1 + (int)(rng_.next() % 4);
		if (rng_.chance(8)) n = rng_.range(5, maxc);
		int cnt = 0;
		for (int i = 0; i < n; ++i) {
			int pos = rng_.chance(70) ? rng_.range(0, maxc > 6 ? 5 : maxc - 1) : rng_.range(0, maxc - 1);
			if (lv[pos]) continue;
			int mag = rng_.chance(75) ? 1 : rng_.range(2, o_.max_level);
			if (rng_.chance(1) && o_.big_levels) mag = rng_.range(16, 300);
			int r = zigzag4x4[pos + first_scan];
			int cost = (mag * level_scale(qp, r & 3, r >> 2) + 63) / 64 + 1;
			if (cost > *budget) {
				mag = 1;
				cost = (level_scale(qp, r & 3, r >> 2) + 63) / 64 + 1;
				if (cost > *budget) continue;
			}
			*budget -= cost;
			lv[pos] = (int16_t)(rng_.chance(50) ? mag : -mag);
			cnt++;
		}
		return cnt;
	}

	// 9.2 encode one block; lv[] in scan order, maxc = 4, 15 or 16
	int code_block(BitWriter &bw, const int16_t *lv, int maxc, int nc) {
		int idx[16], total = 0;
		for (int i = 0; i < maxc; ++i) if (lv[i]) idx[total++] = i;
		int t1 = 0;
		for (int k = total - 1; k >= 0 && t1 < 3; --k) {
			if (lv[idx[k]] == 1 || lv[idx[k]] == -1) t1++;
			else break;
		}
		const VlcEnc &ct = nc < 0 ? enc_.coeff_token[4] : (nc < 2 ? enc_.coeff_token[0] : (nc < 4 ? enc_.coeff_token[1] : (nc < 8 ? enc_.coeff_token[2] : enc_.coeff_token[3])));
		ct.put(bw, total * 4 + t1);
		if (!total) return 0;
		for (int k = 0; k < t1; ++k) bw.put1(lv[idx[total - 1 - k]] < 0);
		int sl = (total > 10 && t1 < 3) ? 1 : 0;
		for (int k = t1; k < total; ++k) {
			int level = lv[idx[total - 1 - k]];
			int code = level > 0 ? 2 * level - 2 : -2 * level - 1;
			if (k == t1 && t1 < 3) code -= 2;
			if (sl == 0) {
				if (code < 14) { bw.put(0, code); bw.put1(1); }
				else if (code < 30) { bw.put(0, 14); bw.put1(1); bw.put(code - 14, 4); }
				else { bw.put(0, 15); bw.put1(1); bw.put(code - 30, 12); }
			} else {
				if (code < (15 << sl)) { bw.put(0, code >> sl); bw.put1(1); bw.put(code & ((1 << sl) - 1), sl); }
				else { bw.put(0, 15); bw.put1(1); bw.put(code - (15 << sl), 12); }
			}
			if (sl == 0) sl = 1;
			int a = level < 0 ? -level : level;
			if (sl < 6 && a > (3 << (sl - 1))) sl++;
		}
		int zeros = idx[total - 1] + 1 - total;
		if (total < maxc) (maxc == 4 ? enc_.total_zeros_cdc[total - 1] : enc_.total_zeros[total - 1]).put(bw, zeros);
		int left = zeros;
		for (int k = total - 1; k > 0 && left > 0; --k) {
			int run = idx[k] - idx[k - 1] - 1;
			enc_.run_before[(left > 7 ? 7 : left) - 1].put(bw, run);
			left -= run;
		}
		return total;
	}

	struct Residual {
		int16_t luma_dc[16];
		int16_t luma[16][16];
		bool luma_coded[16];
		int16_t cdc[2][4];
		int16_t cac[2][4][16];
		bool cac_coded[2][4];
		int16_t luma8[4][64];     // transform_size_8x8_flag: levels of the four 8x8 blocks in scan order
		int cbp;
	};

	// dequantisation scale of position (x, y) of an 8x8 block (flat matrices, 8.5.9 normAdjust8x8)
	static int level_scale8(int qp, int x, int y) {
		static const int na[6][6] = {{20, 18, 32, 19, 25, 24}, {22, 19, 35, 21, 28, 26}, {26, 23, 42, 24, 33, 31},
		                             {28, 25, 45, 26, 35, 33}, {32, 28, 51, 30, 40, 38}, {36, 32, 58, 34, 46, 43}};
		int cls;
		if ((x & 3) == 0 && (y & 3) == 0) cls = 0;
		else if ((x & 1) && (y & 1)) cls = 1;
		else if ((x & 3) == 2 && (y & 3) == 2) cls = 2;
		else if (((x & 3) == 0 && (y & 1)) || ((x & 1) && (y & 3) == 0)) cls = 3;
		else if (((x & 3) == 0 && (y & 3) == 2) || ((x & 3) == 2 && (y & 3) == 0)) cls = 4;
		else cls = 5;
		return (na[qp % 6][cls] << (qp / 6)) >> 2;
	}

	int draw_block8(int16_t *lv, int qp, int *budget) {
		memset(lv, 0, sizeof(int16_t) * 64);
		int n = 1 + (int)(rng_.next() % 6);
		if (rng_.chance(8)) n = rng_.range(7, 40);
		int cnt = 0;
		for (int i = 0; i < n; ++i) {
			int pos = rng_.chance(70) ? rng_.range(0, 9) : rng_.range(0, 63);
			if (lv[pos]) continue;
			int mag = rng_.chance(75) ? 1 : rng_.range(2, o_.max_level);
			int r = zigzag8x8[pos];
			int sc = level_scale8(qp, r & 7, r >> 3);
			int cost = (mag * sc + 63) / 64 + 1;
			if (cost > *budget) {
				mag = 1;
				cost = (sc + 63) / 64 + 1;
				if (cost > *budget) continue;
			}
			*budget -= cost;
			lv[pos] = (int16_t)(rng_.chance(50) ? mag : -mag);
			cnt++;
		}
		return cnt;
	}

	// decide the residual of one MB (after QP is final); i16 selects the Intra16x16 layout
	void draw_residual(Residual &r, int qp, int qpc0, int qpc1, bool i16, bool t8x8 = false) {
		memset(&r, 0, sizeof(r));
		int budget = o_.amp_budget;
		int cbp = 0;
		if (t8x8) {
			for (int q = 0; q < 4; ++q) {
				if (!rng_.chance(o_.coded_pct * 2 > 100 ? 100 : o_.coded_pct * 2)) continue;
				int b = budget;
				if (draw_block8(r.luma8[q], qp, &b) > 0) cbp |= 1 << q;
			}
		} else if (i16) {
			if (rng_.chance(70)) {
				int n = rng_.range(1, 4);
				for (int i = 0; i < n; ++i) {
					// Hadamard of the DC matrix spreads every level to all 16 blocks: |f| <= sum |c|
					int cost = (2 * level_scale(qp, 0, 0) / 4 + 63) / 64 + 1;
					if (cost > budget) break;
					budget -= cost;
					r.luma_dc[rng_.range(0, 15)] = (int16_t)(rng_.chance(50) ? rng_.range(1, 2) : -rng_.range(1, 2));
				}
			}
			if (rng_.chance(50)) {
				cbp |= 15;
				for (int k = 0; k < 16; ++k)
					if (rng_.chance(o_.coded_pct)) { int b = budget; r.luma_coded[k] = draw_block(r.luma[k], 15, 1, qp, &b) > 0; }
			}
		} else {
			for (int q = 0; q < 4; ++q) {
				if (!rng_.chance(o_.coded_pct * 2 > 100 ? 100 : o_.coded_pct * 2)) continue;
				bool any = false;
				for (int s = 0; s < 4; ++s) {
					int k = q * 4 + s;
					if (rng_.chance(50)) { int b = budget; r.luma_coded[k] = draw_block(r.luma[k], 16, 0, qp, &b) > 0; any |= r.luma_coded[k]; }
				}
				if (any) cbp |= 1 << q;
			}
		}
		int cmode = rng_.chance(o_.coded_pct) ? rng_.range(1, 2) : 0;
		if (cmode) {
			bool any_dc = false, any_ac = false;
			for (int c = 0; c < 2; ++c) {
				int q = c ? qpc1 : qpc0;
				int b = o_.amp_budget;
				int n = rng_.range(0, 3);
				for (int i = 0; i < n; ++i) {
					int cost = (2 * level_scale(q, 0, 0) / 2 + 63) / 64 + 1;
					if (cost > b) break;
					b -= cost;
					r.cdc[c][rng_.range(0, 3)] = (int16_t)(rng_.chance(50) ? rng_.range(1, 2) : -rng_.range(1, 2));
					any_dc = true;
				}
				if (cmode == 2)
					for (int k = 0; k < 4; ++k)
						if (rng_.chance(50)) { int bb = b; r.cac_coded[c][k] = draw_block(r.cac[c][k], 15, 1, q, &bb) > 0; any_ac |= r.cac_coded[c][k]; }
			}
			if (cmode == 2 && !any_ac) cmode = 1;
			if (cmode == 1 && !any_dc) cmode = 0;
			if (!cmode) { memset(r.cdc, 0, sizeof(r.cdc)); memset(r.cac_coded, 0, sizeof(r.cac_coded)); }
		}
		r.cbp = cbp | (cmode << 4);
	}

	void code_luma(BitWriter &bw, const Residual &r, int mbx, int mby, bool i16, bool intra = false, bool t8x8 = false) {
		uint8_t *nnz = ctx_.nnz_of(mbx, mby);
		static const int16_t zero[16] = {0};
		if (t8x8) {
			for (int q = 0; q < 4; ++q) {
				const int x4 = (q & 1) * 2, y4 = (q >> 1) * 2;
				int total = 0;
				if (r.cbp & (1 << q)) {
					int16_t lv[64];
					memcpy(lv, r.luma8[q], sizeof(lv));
					total = cs_.residual(5, -1, lv, 64);
					if (total > 15) total = 15;
				}
				nnz[y4 * 4 + x4] = nnz[y4 * 4 + x4 + 1] = nnz[(y4 + 1) * 4 + x4] = nnz[(y4 + 1) * 4 + x4 + 1] = (uint8_t)total;
			}
			return;
		}
		for (int k = 0; k < 16; ++k) {
			int x4 = blk_x[k], y4 = blk_y[k];
			if (!(r.cbp & (1 << (k >> 2)))) { nnz[y4 * 4 + x4] = 0; continue; }
			const int16_t *src = r.luma_coded[k] ? r.luma[k] : zero;
			if (o_.cabac) {
				int16_t lv[16];
				memcpy(lv, src, sizeof(lv));
				nnz[y4 * 4 + x4] = (uint8_t)cs_.residual(i16 ? 1 : 2, ctx_.cabac_inc_cbf_luma(x4, y4, intra), lv, i16 ? 15 : 16);
			} else {
				int nc = ctx_.cavlc_nc(mbx, mby, 0, x4, y4, slice_num_);
				nnz[y4 * 4 + x4] = (uint8_t)code_block(bw, src, i16 ? 15 : 16, nc);
			}
		}
	}

	void code_chroma(BitWriter &bw, const Residual &r, int mbx, int mby, bool intra = false) {
		uint8_t *nnz = ctx_.nnz_of(mbx, mby);
		MbInfo &info = ctx_.info(mbx, mby);
		memset(nnz + 16, 0, 8);
		int cm = r.cbp >> 4;
		if (!cm) return;
		static const int16_t zero[16] = {0};
		for (int c = 0; c < 2; ++c) {
			if (o_.cabac) {
				int16_t lv[4];
				memcpy(lv, r.cdc[c], sizeof(lv));
				if (cs_.residual(3, ctx_.cabac_inc_cbf_dc(MBF_CBF_CB_DC << c, intra), lv, 4)) info.cflags |= (uint8_t)(MBF_CBF_CB_DC << c);
			} else {
				code_block(bw, r.cdc[c], 4, -1);
			}
		}
		if (cm & 2)
			for (int c = 0; c < 2; ++c)
				for (int b = 0; b < 4; ++b) {
					const int16_t *src = r.cac_coded[c][b] ? r.cac[c][b] : zero;
					if (o_.cabac) {
						int16_t lv[16];
						memcpy(lv, src, sizeof(lv));
						nnz[16 + c * 4 + b] = (uint8_t)cs_.residual(4, ctx_.cabac_inc_cbf_chroma_ac(c, b & 1, b >> 1, intra), lv, 15);
					} else {
						int nc = ctx_.cavlc_nc(mbx, mby, 1 + c, b & 1, b >> 1, slice_num_);
						nnz[16 + c * 4 + b] = (uint8_t)code_block(bw, src, 15, nc);
					}
				}
	}

	// coded_block_pattern [+ transform_size_8x8_flag] + mb_qp_delta + residual of a non-Intra16x16 MB
	void code_cbp_residual(BitWriter &bw, Residual &r, int mbx, int mby, bool intra, bool t8x8, int qd, int qp) {
		MbInfo &info = ctx_.info(mbx, mby);
		put_cbp(bw, intra, r.cbp);
		info.cbp = (uint8_t)r.cbp;
		if (!intra && o_.transform8x8 && may_t8x8_ && (r.cbp & 15)) put_t8x8(t8x8);
		if (t8x8 && (intra || (r.cbp & 15))) info.t8x8 = 1;
		if (r.cbp) { put_qp_delta(bw, qd); qp_ = qp; }
		else last_dqp_ = 0;
		code_luma(bw, r, mbx, mby, false, intra, t8x8);
		code_chroma(bw, r, mbx, mby, intra);
	}

	int draw_qp_delta() {
		if (!o_.vary_qp || !rng_.chance(15)) return 0;
		int d = rng_.range(-2, 2);
		if (qp_ + d < o_.qp - 4 || qp_ + d > o_.qp + 4) d = 0;
		return d;
	}

	// ---- intra ------------------------------------------------------------------------------------------
	int pick(const int *cands, int n) { return cands[rng_.range(0, n - 1)]; }

	void intra_mb(BitWriter &bw, int mbx, int mby, bool islice) {
		int avail = ctx_.mb_avail_mask(mbx, mby, slice_num_);
		if (o_.constrained_intra) {
			static const int dx[4] = {-1, 0, 1, -1}, dy[4] = {0, -1, -1, -1};
			for (int b = 0; b < 4; ++b)
				if ((avail >> b) & 1)
					if (!kind_is_intra(ctx_.info(mbx + dx[b], mby + dy[b]).kind)) avail &= ~(1 << b);
		}
		bool L = avail & 1, T = avail & 2, TL = avail & 8;
		MbInfo &info = ctx_.info(mbx, mby);
		ctx_.set_intra_motion(mbx, mby);
		int8_t *ip = ctx_.ipred_of(mbx, mby);
		int ofs = islice ? 0 : (kind_ == 1 ? 23 : 5);
		// chroma mode candidates
		int cc[4], ncc = 0;
		cc[ncc++] = 0;
		if (L) cc[ncc++] = 1;
		if (T) cc[ncc++] = 2;
		if (L && T && TL) cc[ncc++] = 3;
		int chroma_mode = pick(cc, ncc);
		int r100 = rng_.range(0, 99);
		if (r100 < o_.ipcm_pct) {
			put_mb_type(bw, ofs + 25);
			bw.align_zero();
			for (int i = 0; i < 384; ++i) bw.put((uint32_t)rng_.range(0, 255), 8);
			if (o_.cabac) cs_.e.start(&bw);
			info.kind = MB_IPCM;
			info.cbp = 0x2f;
			info.cflags = MBF_CBF_LUMA_DC | MBF_CBF_CB_DC | MBF_CBF_CR_DC;
			last_dqp_ = 0;
			memset(ctx_.nnz_of(mbx, mby), 16, 24);
			memset(ip, 2, 16);
			return;
		}
		Residual r;
		may_t8x8_ = false;
		if (r100 < 55) {   // Intra NxN
			const bool t8x8 = o_.transform8x8 && rng_.chance(50);
			info.kind = t8x8 ? MB_I8x8 : MB_I4x4;
			put_mb_type(bw, ofs + 0);
			if (o_.transform8x8) put_t8x8(t8x8);
			for (int k = 0; k < (t8x8 ? 4 : 16); ++k) {
				int x4 = t8x8 ? (k & 1) * 2 : blk_x[k], y4 = t8x8 ? (k >> 1) * 2 : blk_y[k];
				bool l = x4 > 0 || L, t = y4 > 0 || T;
				bool tl = (x4 > 0 && y4 > 0) || (x4 == 0 && y4 > 0 && L) || (y4 == 0 && x4 > 0 && T) || (x4 == 0 && y4 == 0 && TL);
				int cand[9], n = 0;
				cand[n++] = 2;
				if (t) { cand[n++] = 0; cand[n++] = 3; cand[n++] = 7; }
				if (l) { cand[n++] = 1; cand[n++] = 8; }
				if (l && t && tl) { cand[n++] = 4; cand[n++] = 5; cand[n++] = 6; }
				int mode = pick(cand, n);
				int pred = ctx_.predict_ipred(mbx, mby, x4, y4, slice_num_, o_.constrained_intra != 0);
				if (rng_.chance(30)) mode = pred;       // pred is always legal here: it is DC or a neighbour's legal choice... not necessarily
				// a neighbour's mode may need samples this block lacks; fall back to DC then
				bool ok = false;
				for (int i = 0; i < n; ++i) ok |= cand[i] == mode;
				if (!ok) mode = 2;
				put_ipred(bw, pred, mode);
				ip[y4 * 4 + x4] = (int8_t)mode;
				if (t8x8) ip[y4 * 4 + x4 + 1] = ip[(y4 + 1) * 4 + x4] = ip[(y4 + 1) * 4 + x4 + 1] = (int8_t)mode;
			}
			put_chroma_mode(bw, chroma_mode);
			info.chroma_mode = (uint8_t)chroma_mode;
			int qd = draw_qp_delta();
			int qp = qp_ + qd;
			draw_residual(r, qp, chroma_qp(qp, o_.chroma_qp_offset), chroma_qp(qp, o_.chroma_qp_offset), false, t8x8);
			code_cbp_residual(bw, r, mbx, mby, true, t8x8, qd, qp);
			return;
		}
		// Intra16x16
		info.kind = MB_I16x16;
		memset(ip, 2, 16);
		int cand[4], n = 0;
		cand[n++] = 2;
		if (T) cand[n++] = 0;
		if (L) cand[n++] = 1;
		if (L && T && TL) cand[n++] = 3;
		int mode = pick(cand, n);
		int qd = draw_qp_delta();
		int qp = qp_ + qd;
		draw_residual(r, qp, chroma_qp(qp, o_.chroma_qp_offset), chroma_qp(qp, o_.chroma_qp_offset), true);
		int cl = (r.cbp & 15) ? 1 : 0, cchroma = r.cbp >> 4;
		put_mb_type(bw, ofs + 1 + mode + cchroma * 4 + cl * 12);
		put_chroma_mode(bw, chroma_mode);
		info.chroma_mode = (uint8_t)chroma_mode;
		info.cbp = (uint8_t)r.cbp;
		put_qp_delta(bw, qd);
		qp_ = qp;
		if (o_.cabac) {
			int16_t lv[16];
			memcpy(lv, r.luma_dc, sizeof(lv));
			if (cs_.residual(0, ctx_.cabac_inc_cbf_dc(MBF_CBF_LUMA_DC, true), lv, 16)) info.cflags |= MBF_CBF_LUMA_DC;
		} else {
			int nc = ctx_.cavlc_nc(mbx, mby, 0, 0, 0, slice_num_);
			code_block(bw, r.luma_dc, 16, nc);
		}
		code_luma(bw, r, mbx, mby, true, true, false);
		code_chroma(bw, r, mbx, mby, true);
	}

	// ---- inter -------------------------------------------------------------------------------------------
	void draw_mv(int mbx, int mby, const int16_t pred[2], int out[2]) {
		int mode = rng_.range(0, 99);
		if (mode < 25) { out[0] = pred[0]; out[1] = pred[1]; }
		else if (mode < 85) {
			out[0] = pred[0] + rng_.range(-o_.mv_step, o_.mv_step);
			out[1] = pred[1] + rng_.range(-o_.mv_step, o_.mv_step);
		} else {
			out[0] = rng_.range(-o_.mv_range, o_.mv_range);
			out[1] = rng_.range(-o_.mv_range, o_.mv_range);
		}
		// keep the block within `mv_out` pixels of the picture so that edge replication is
		// exercised but vectors stay inside level limits
		int px = mbx * 16 * 4, py = mby * 16 * 4;
		int lim = o_.mv_out * 4;
		int minx = -px - lim, maxx = (o_.mbw * 16 * 4 - px - 16 * 4) + lim;
		int miny = -py - lim, maxy = (o_.mbh * 16 * 4 - py - 16 * 4) + lim;
		if (out[0] < minx) out[0] = minx;
		if (out[0] > maxx) out[0] = maxx;
		if (out[1] < miny) out[1] = miny;
		if (out[1] > maxy) out[1] = maxy;
	}

	void inter_mb(BitWriter &bw, int mbx, int mby) {
		MbInfo &info = ctx_.info(mbx, mby);
		info.kind = MB_INTER;
		memset(ctx_.ipred_of(mbx, mby), 2, 16);
		ctx_.set_motion(1, mbx, mby, 0, 0, 4, 4, -1, -1, 0, 0);
		struct Part { int x4, y4, pw, ph, ref, shape, part; };
		std::vector<Part> parts;
		int t = rng_.range(0, 99);
		int mb_type, sub[4] = {0, 0, 0, 0}, refs[4];
		for (int i = 0; i < 4; ++i) refs[i] = nref_[0] > 1 ? rng_.range(0, nref_[0] - 1) : 0;
		if (t < 40) { mb_type = 0; parts.push_back({0, 0, 4, 4, refs[0], 0, 0}); }
		else if (t < 55) { mb_type = 1; parts.push_back({0, 0, 4, 2, refs[0], 1, 0}); parts.push_back({0, 2, 4, 2, refs[1], 1, 1}); }
		else if (t < 70) { mb_type = 2; parts.push_back({0, 0, 2, 4, refs[0], 2, 0}); parts.push_back({2, 0, 2, 4, refs[1], 2, 1}); }
		else {
			const bool plain = nref_[0] > 1 || rng_.chance(50);
			mb_type = (plain || o_.cabac) ? 3 : 4;   // P_8x8 / P_8x8ref0 (the latter does not exist under CABAC)
			if (mb_type == 4) for (int i = 0; i < 4; ++i) refs[i] = 0;
			for (int i = 0; i < 4; ++i) {
				sub[i] = rng_.range(0, 3);
				if (o_.transform8x8 && (t & 1)) sub[i] = 0;    // half of the P_8x8 MBs may use the 8x8 transform
				int bx = (i & 1) * 2, by = (i >> 1) * 2;
				switch (sub[i]) {
				case 0: parts.push_back({bx, by, 2, 2, refs[i], 0, 0}); break;
				case 1: parts.push_back({bx, by, 2, 1, refs[i], 0, 0}); parts.push_back({bx, by + 1, 2, 1, refs[i], 0, 0}); break;
				case 2: parts.push_back({bx, by, 1, 2, refs[i], 0, 0}); parts.push_back({bx + 1, by, 1, 2, refs[i], 0, 0}); break;
				default: for (int s = 0; s < 4; ++s) parts.push_back({bx + (s & 1), by + (s >> 1), 1, 1, refs[i], 0, 0}); break;
				}
			}
		}
		put_mb_type(bw, mb_type);
		may_t8x8_ = true;
		if (mb_type == 0) put_ref(bw, 0, 0, 0, 4, 4, refs[0]);
		else if (mb_type == 1) { put_ref(bw, 0, 0, 0, 4, 2, refs[0]); put_ref(bw, 0, 0, 2, 4, 2, refs[1]); }
		else if (mb_type == 2) { put_ref(bw, 0, 0, 0, 2, 4, refs[0]); put_ref(bw, 0, 2, 0, 2, 4, refs[1]); }
		else {
			for (int i = 0; i < 4; ++i) { put_sub_mb_type(bw, sub[i]); if (sub[i]) may_t8x8_ = false; }
			if (mb_type == 3) for (int i = 0; i < 4; ++i) put_ref(bw, 0, (i & 1) * 2, (i >> 1) * 2, 2, 2, refs[i]);
		}
		for (const Part &p : parts) {
			int16_t pred[2];
			ctx_.predict_mv(0, mbx, mby, p.x4, p.y4, p.pw, p.ref, slice_num_, p.shape, p.part, pred);
			int mv[2];
			draw_mv(mbx, mby, pred, mv);
			put_mvd(bw, 0, p.x4, p.y4, p.pw, p.ph, mv[0] - pred[0], mv[1] - pred[1]);
			ctx_.set_motion(0, mbx, mby, p.x4, p.y4, p.pw, p.ph, p.ref, pic_of(0, p.ref), mv[0], mv[1]);
		}
		Residual r;
		int qd = draw_qp_delta();
		int qp = qp_ + qd;
		const bool t8x8 = o_.transform8x8 && may_t8x8_ && rng_.chance(50);
		draw_residual(r, qp, chroma_qp(qp, o_.chroma_qp_offset), chroma_qp(qp, o_.chroma_qp_offset), false, t8x8);
		code_cbp_residual(bw, r, mbx, mby, false, t8x8, qd, qp);
	}

	// ---- B pictures ---------------------------------------------------------------------------------------
	int pic_of(int l, int refidx) const { return refidx >= 0 && refidx < (int)list_[l].size() ? list_[l][refidx]->rp.frame_idx : -1; }

	void derive_direct(int mbx, int mby, DirectOut &d) {
		const ColPic *col = list_[1].empty() ? nullptr : &list_[1][0]->col;
		if (o_.direct_spatial) ctx_.direct_spatial(mbx, mby, slice_num_, col, true, o_.direct_8x8_inference != 0, d);
		else ctx_.direct_temporal(mbx, mby, col, o_.direct_8x8_inference != 0, rp_[0].data(), (int)rp_[0].size(), rp_[1].data(), cur_poc_, d);
	}
	void set_direct(int mbx, int mby, const DirectOut &d, int quads) {
		for (int q = 0; q < 4; ++q) {
			if (!((quads >> q) & 1)) continue;
			for (int y = (q >> 1) * 2; y < (q >> 1) * 2 + 2; ++y)
				for (int x = (q & 1) * 2; x < (q & 1) * 2 + 2; ++x)
					for (int l = 0; l < 2; ++l)
						ctx_.set_motion(l, mbx, mby, x, y, 1, 1, d.ref[y * 4 + x][l], pic_of(l, d.ref[y * 4 + x][l]), d.mv[y * 4 + x][l][0], d.mv[y * 4 + x][l][1]);
		}
	}

	void b_mb(BitWriter &bw, int mbx, int mby) {
		MbInfo &info = ctx_.info(mbx, mby);
		info.kind = MB_INTER;
		memset(ctx_.ipred_of(mbx, mby), 2, 16);
		struct Part { int x4, y4, pw, ph, shape, part, use, sub; };
		std::vector<Part> parts;
		int sub_use[4] = {0, 0, 0, 0}, sub_type[4] = {0, 0, 0, 0};
		int mb_type, direct_quads = 0;
		int t = rng_.range(0, 99);
		auto mode = [&]() { return rng_.range(1, 3); };
		if (t < o_.direct_pct) {
			mb_type = 0;
			direct_quads = 0xf;
		} else if (t < o_.direct_pct + 25) {
			int u = mode();
			mb_type = u;
			parts.push_back({0, 0, 4, 4, 0, 0, u, 0});
			sub_use[0] = u;
		} else if (t < o_.direct_pct + 55) {
			static const uint8_t modes[22][2] = {{0, 0}, {1, 0}, {2, 0}, {3, 0}, {1, 1}, {1, 1}, {2, 2}, {2, 2}, {1, 2}, {1, 2}, {2, 1}, {2, 1},
			                                     {1, 3}, {1, 3}, {2, 3}, {2, 3}, {3, 1}, {3, 1}, {3, 2}, {3, 2}, {3, 3}, {3, 3}};
			mb_type = rng_.range(4, 21);
			if (mb_type & 1) { parts.push_back({0, 0, 2, 4, 2, 0, modes[mb_type][0], 0}); parts.push_back({2, 0, 2, 4, 2, 1, modes[mb_type][1], 1}); }
			else { parts.push_back({0, 0, 4, 2, 1, 0, modes[mb_type][0], 0}); parts.push_back({0, 2, 4, 2, 1, 1, modes[mb_type][1], 1}); }
			sub_use[0] = modes[mb_type][0];
			sub_use[1] = modes[mb_type][1];
		} else {
			static const uint8_t sub_mode[13] = {0, 1, 2, 3, 1, 1, 2, 2, 3, 3, 1, 2, 3};
			static const uint8_t sub_shape[13] = {0, 0, 0, 0, 1, 2, 1, 2, 1, 2, 3, 3, 3};
			mb_type = 22;
			for (int q = 0; q < 4; ++q) {
				// with transform_8x8_mode the reference reads transform_size_8x8_flag for every B_8x8
				// (need_transform_size_8x8, h264.cpp:1306) although sub-8x8 partitions exclude it
				int st = rng_.chance(o_.direct_pct) ? 0 : rng_.range(1, o_.transform8x8 ? 3 : 12);
				sub_type[q] = st;
				int bx = (q & 1) * 2, by = (q >> 1) * 2, use = sub_mode[st];
				sub_use[q] = use;
				if (use == 0) { direct_quads |= 1 << q; parts.push_back({bx, by, 2, 2, 0, 0, 0, q}); continue; }
				switch (sub_shape[st]) {
				case 0: parts.push_back({bx, by, 2, 2, 0, 0, use, q}); break;
				case 1: parts.push_back({bx, by, 2, 1, 0, 0, use, q}); parts.push_back({bx, by + 1, 2, 1, 0, 0, use, q}); break;
				case 2: parts.push_back({bx, by, 1, 2, 0, 0, use, q}); parts.push_back({bx + 1, by, 1, 2, 0, 0, use, q}); break;
				default: for (int k = 0; k < 4; ++k) parts.push_back({bx + (k & 1), by + (k >> 1), 1, 1, 0, 0, use, q}); break;
				}
			}
		}
		put_mb_type(bw, mb_type);
		may_t8x8_ = mb_type != 0 || o_.direct_8x8_inference;
		if (mb_type == 0) info.cflags |= MBF_DIRECT16;
		if (mb_type == 22)
			for (int q = 0; q < 4; ++q) {
				put_sub_mb_type(bw, sub_type[q]);
				if (sub_type[q] > 3 || (sub_type[q] == 0 && !o_.direct_8x8_inference)) may_t8x8_ = false;
			}
		DirectOut dir;
		if (direct_quads) derive_direct(mbx, mby, dir);
		int ref[4][2];
		const int nsub = mb_type == 22 ? 4 : (mb_type == 0 ? 0 : (mb_type <= 3 ? 1 : 2));
		for (int l = 0; l < 2; ++l)
			for (int q = 0; q < nsub; ++q) {
				ref[q][l] = -1;
				if (!(sub_use[q] & (1 << l))) continue;
				int r = nref_[l] > 1 ? rng_.range(0, nref_[l] - 1) : 0;
				ref[q][l] = r;
				if (mb_type == 22) put_ref(bw, l, (q & 1) * 2, (q >> 1) * 2, 2, 2, r);
				else put_ref(bw, l, parts[q].x4, parts[q].y4, parts[q].pw, parts[q].ph, r);
			}
		for (int l = 0; l < 2; ++l)
			for (const Part &p : parts) {
				if (p.use == 0) {
					for (int y = p.y4; y < p.y4 + 2; ++y)
						for (int x = p.x4; x < p.x4 + 2; ++x)
							ctx_.set_motion(l, mbx, mby, x, y, 1, 1, dir.ref[y * 4 + x][l], pic_of(l, dir.ref[y * 4 + x][l]), dir.mv[y * 4 + x][l][0], dir.mv[y * 4 + x][l][1]);
					continue;
				}
				if (!(p.use & (1 << l))) { ctx_.set_motion(l, mbx, mby, p.x4, p.y4, p.pw, p.ph, -1, -1, 0, 0); continue; }
				int16_t pred[2];
				int r = ref[p.sub][l];
				ctx_.predict_mv(l, mbx, mby, p.x4, p.y4, p.pw, r, slice_num_, p.shape, p.part, pred);
				int mv[2];
				draw_mv(mbx, mby, pred, mv);
				put_mvd(bw, l, p.x4, p.y4, p.pw, p.ph, mv[0] - pred[0], mv[1] - pred[1]);
				ctx_.set_motion(l, mbx, mby, p.x4, p.y4, p.pw, p.ph, r, pic_of(l, r), mv[0], mv[1]);
			}
		if (mb_type == 0) set_direct(mbx, mby, dir, 0xf);
		Residual r;
		int qd = draw_qp_delta();
		int qp = qp_ + qd;
		const bool t8x8 = o_.transform8x8 && may_t8x8_ && rng_.chance(50);
		draw_residual(r, qp, chroma_qp(qp, o_.chroma_qp_offset), chroma_qp(qp, o_.chroma_qp_offset), false, t8x8);
		code_cbp_residual(bw, r, mbx, mby, false, t8x8, qd, qp);
	}

	struct WRef { RefPic rp; int frame_num_wrap; ColPic col; };
	std::vector<WRef> refs_;
	std::vector<const WRef *> list_[2];
	std::vector<RefPic> rp_[2];
	int kind_ = 2, cur_poc_ = 0, poc_base_ = 0, next_pic_id_ = 0;
	int nref_[2] = {0, 0};
	H264Opts o_;
	Rng rng_;
	Enc enc_;
	PicCtx ctx_;
	std::vector<uint8_t> out_;
	int slice_num_ = 0, qp_ = 26, idr_id_ = 0;
	int last_dqp_ = 0;
	bool may_t8x8_ = false;
	CabacSyntax<CabacEnc> cs_;
};

}  // namespace

std::vector<uint8_t> write_h264(const H264Opts &o) {
	Writer w(o);
	return w.run();
}

}  // namespace sg
