// mpeg2_front.cpp — MPEG-2 video host front end: header / slice / macroblock parse that
// emits per-picture SoA records (b2_m2_mb + dequantised coefficient blocks) instead of
// touching pixels, exported behind the reference's plugin table as `m2d_gpu_func`.
//
// Behavioural model: reference src/lib/mpeg2.cpp (m2d_func, :1800-1811).  Written from the
// standard's syntax with the reference's observable behaviour preserved where it decides
// output: frame rotation and output order (mpeg2.cpp:130-194, 1543-1587), skipped-macroblock
// and missing-slice handling (715-808), predictor resets (1401-1416), dequantisation with
// saturation and mismatch control (945-1097).  Scope: MPEG-2 frame pictures (progressive or
// interlaced frame pictures with field prediction / field DCT).  Field pictures and MPEG-1
// streams are rejected (the reference handles field pictures only partially).
#include <exception>
#include <memory>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include "frontend.h"
#include "mpeg2_tables.h"

namespace b2 {
namespace {

enum { I_PIC = 1, P_PIC = 2, B_PIC = 3 };
enum { MAX_M2_FRAMES = 16 };

struct M2Tables {
	Vlc mb_inc, type_i, type_p, type_b, cbp, motion, dc_luma, dc_chroma, dct[2];
	M2Tables() {
		using namespace m2t;
		mb_inc.build(mb_addr_inc, B2_COUNT(mb_addr_inc), 8);
		type_i.build(mb_type_i, B2_COUNT(mb_type_i), 6);
		type_p.build(mb_type_p, B2_COUNT(mb_type_p), 6);
		type_b.build(mb_type_b, B2_COUNT(mb_type_b), 6);
		cbp.build(m2t::cbp, B2_COUNT(m2t::cbp), 9);
		motion.build(motion_code, B2_COUNT(motion_code), 8);
		dc_luma.build(dc_size_luma, B2_COUNT(dc_size_luma), 9);
		dc_chroma.build(dc_size_chroma, B2_COUNT(dc_size_chroma), 10);
		dct[0].build(dct_table0, B2_COUNT(dct_table0), 9);
		dct[1].build(dct_table1, B2_COUNT(dct_table1), 9);
	}
};
const M2Tables &tables() {
	static const M2Tables t;
	return t;
}

struct BitError {};

class Mpeg2Decoder {
public:
	Mpeg2Decoder(dec_bits *stream) : stream_(stream) { reset(nullptr, nullptr); }

	void reset(int (*cb)(void *, void *), void *arg) {
		header_cb_ = cb;
		header_arg_ = arg;
		width_ = height_ = 0;
		mb_w_ = mb_h_ = 0;
		disp_w_ = disp_h_ = 0;
		is_mpeg2_ = false;
		memcpy(qmat_[0], m2t::default_intra_qmat, 64);
		memset(qmat_[1], 16, 64);
		pic_type_ = open_type_ = 0;
		temporal_ref_ = 0;
		q_scale_ = 2;
		pic_fwd_ = pic_bwd_ = 0;
		memset(frames_, 0, sizeof(frames_));
		memset(&scratch_mb_, 0, sizeof(scratch_mb_));
		reset_dc_init();
		for (int i = 0; i < 4; ++i) f_code_[i >> 1][i & 1] = 15;
		intra_dc_precision_ = 0;
		picture_structure_ = 3;
		frame_pred_frame_dct_ = 1;
		concealment_mv_ = 0;
		q_scale_type_ = 0;
		intra_vlc_format_ = 0;
		alternate_scan_ = 0;
		num_frames_ = 0;
		frames_mb_w_ = frames_mb_h_ = 0;
		cur_idx_ = -1;
		ref_idx_[0] = ref_idx_[1] = 0;
		memset(lru_, 0, sizeof(lru_));
		out_state_ = 0;
		outq_n_ = 0;
		mb_x_ = -1;
		mb_y_ = 0;
		prev_type_ = 0;
		picture_open_ = false;
		memset(host_valid_, 0, sizeof(host_valid_));
	}

	Backend backend;

	// ---- m2d_func_table_t entry points -------------------------------------------------
	int get_info(m2d_info_t *info) const {
		if (!info) return -1;
		info->src_width = (int16_t)((width_ + 15) & ~15);
		info->src_height = (int16_t)((height_ + 15) & ~15);
		info->disp_width = (int16_t)disp_w_;
		info->disp_height = (int16_t)disp_h_;
		info->frame_num = 4;      // two references, the picture in flight and a B picture waiting to go out (peek_frame)
		info->crop[0] = 0;
		info->crop[1] = (int16_t)(info->src_width - width_);
		info->crop[2] = 0;
		info->crop[3] = (int16_t)(info->src_height - height_);
		info->additional_size = 0;
		return 0;
	}

	int set_frames(int num, const m2d_frame_t *frames) {
		if (num < 0 || num > MAX_M2_FRAMES || !frames) return -1;
		if (mb_w_ <= 0 || mb_h_ <= 0 || (long)mb_w_ * mb_h_ > 139264) return -1;   // no (usable) sequence header yet
		for (int i = 0; i < num; ++i) {
			if (!frames[i].luma || !frames[i].chroma) return -1;
			if (((uintptr_t)frames[i].luma | (uintptr_t)frames[i].chroma) & 15) return -1;
		}
		num_frames_ = num;
		memcpy(frames_, frames, sizeof(frames_[0]) * num);
		cur_idx_ = -1;
		outq_n_ = 0;
		memset(host_valid_, 0, sizeof(host_valid_));
		frames_mb_w_ = mb_w_;    // the application sized its buffers for this geometry (get_info); slices of
		frames_mb_h_ = mb_h_;    // any other are ignored, their pictures would overrun those buffers
		backend.bind_frames(frames_, num);
		// frames go out with width / height of the sequence header and the application's own (zero) crop fields: its
		// writer reads both planes with pitch `width`
		backend.set_md5_window(width_, 0, 0, width_, height_);
		return backend.configure(B2_CODEC_MPEG2, mb_w_ * 16, mb_h_ * 16, num);
	}

	int decode_picture() {
		pic_type_ = 0;
		for (;;) {
			if (backend.dead()) return -4;   // no CUDA device: nothing else can reconstruct
			Unit u;
			if (next_unit(stream_, &u) < 0) {
				// end of data: a picture whose last macroblock never came (truncated stream, lost last
				// slice) is still submitted, so that frame-out does not hand out stale pixels
				try {
					flush_picture();
				} catch (const std::exception &) {
				}
				return -1;
			}
			if (u.len == 0) continue;
			int code = u.data[0];
			br_.init(u.data + 1, u.len - 1);
			int r = 0;
			try {
				r = dispatch(code);
			} catch (const BitError &) {
				r = 0;   // the reference unwinds to m2d_dispatch_one_nal and carries on (mpeg2.cpp:666-668)
			} catch (const std::exception &e) {
				// allocation failure and the like: nothing may unwind through the extern "C" entry point
				fprintf(stderr, "m2dec_b200: MPEG-2 decode_picture: %s\n", e.what());
				picture_open_ = false;
				return -1;
			}
			if (r == 1) {
				note_due();
				return 1;
			}
		}
	}

	// Frame-out.  mpeg2.cpp:1543-1587 decides which frame is due after a picture; the order of frames is the reference's.
	// One thing is different: a frame is due the moment its picture (a B picture) or the picture after it (an anchor) has
	// been decoded, which here is the moment that picture's records were handed to the device, so returning it at once
	// would make the host wait for the copy, the kernel and the frame's way back, picture after picture.  With a fourth
	// frame buffer (get_info asks for it) every due frame goes out one decode_picture call later instead: the frames wait
	// in a queue whose last entry is held back until the next picture is done or the stream ends.  The held frame is
	// either a reference or the most recently written buffer, neither of which the next picture takes (pick_frame), and
	// a frame that has been handed out stays untouched until the next decode_picture call, as in the reference.
	int peek_frame(m2d_frame_t *out, int is_end) { return out_frame(out, is_end, false); }
	int get_frame(m2d_frame_t *out, int is_end) { return out_frame(out, is_end, true); }

private:
	bool can_defer() const { return num_frames_ >= 4; }
	// after a complete picture: what the reference's frame-out would report now joins the queue
	void note_due() {
		int idx = -1;
		if (pic_type_ == B_PIC) {
			if (out_state_ & 1) { idx = cur_idx_; out_state_ &= ~1; }
		} else if ((out_state_ >> 1) >= 2) {
			idx = ref_idx_[0];
			out_state_ -= 2;
		}
		if (idx < 0) return;
		if (outq_n_ == MAX_OUTQ) {            // an application that never collects: the oldest frame is dropped
			memmove(outq_, outq_ + 1, sizeof(outq_[0]) * (MAX_OUTQ - 1));
			outq_n_--;
		}
		outq_[outq_n_++] = idx;
	}
	int hand_out(m2d_frame_t *out, int idx, bool fetch) {
		*out = frames_[idx];
		out->width = (int16_t)width_;
		out->height = (int16_t)height_;
		if (fetch && !host_valid_[idx]) {
			if (backend.fetch(idx, frames_[idx].luma, frames_[idx].chroma) < 0) return -1;
			host_valid_[idx] = true;
		}
		return 0;
	}
	int out_frame(m2d_frame_t *out, int is_end, bool consume) {
		if (!out) return -1;
		const int hold = (!is_end && can_defer()) ? 1 : 0;
		if (outq_n_ > hold) {
			const int idx = outq_[0];
			if (hand_out(out, idx, true) < 0) return -1;
			if (consume) {
				outq_n_--;
				memmove(outq_, outq_ + 1, sizeof(outq_[0]) * outq_n_);
			}
			return 1;
		}
		// nothing queued (or only the held frame): at the end of the stream the last anchor is still due
		int idx, ret;
		if (pic_type_ == B_PIC) {
			idx = cur_idx_;
			ret = out_state_ & 1;
		} else {
			idx = (is_end && out_state_ > 0 && out_state_ < 4) ? ref_idx_[1] : ref_idx_[0];
			const int st = out_state_ >> 1;
			ret = (st == 0) ? 0 : (st == 1 ? (is_end != 0) : 1);
		}
		if (idx < 0) idx = 0;
		if (hand_out(out, idx, ret > 0) < 0) return -1;
		if (ret > 0 && consume) {
			if (pic_type_ == B_PIC) out_state_ &= ~1;
			else out_state_ -= 2;
		}
		return ret;
	}

	// ---- start-code dispatch (mpeg2.cpp:662-697) ---------------------------------------
	int dispatch(int code) {
		if (code == 0x00) return picture_header();
		if (code < 0xb0) return slice(code);
		switch (code) {
		case 0xb3: return sequence_header();
		case 0xb5: return extension();
		case 0xb8: return 0;   // group_of_pictures_header: nothing in it affects decoding here
		default: return 0;     // user data, sequence end, anything else
		}
	}

	void load_qmat(uint8_t *q, const uint8_t *scan) {
		for (int i = 0; i < 64; ++i) q[scan[i]] = (uint8_t)br_.get(8);
	}

	void set_size(int w, int h) {
		// a size change under an open picture: submit it first, its records have the old geometry
		if (picture_open_ && (((w + 15) >> 4) != mb_w_ || ((h + 15) >> 4) != mb_h_)) flush_picture();
		width_ = w;
		height_ = h;
		mb_w_ = (w + 15) >> 4;
		mb_h_ = (h + 15) >> 4;
	}

	int sequence_header() {
		int w = br_.get(12), h = br_.get(12);
		br_.skip(4 + 4);          // aspect_ratio_information, frame_rate_code
		br_.skip(18 + 1 + 10 + 1); // bit_rate, marker, vbv_buffer_size, constrained_parameters_flag
		// matrices loaded here are always stored through the zig-zag scan (mpeg2.cpp:266-277)
		if (br_.get1()) load_qmat(qmat_[0], m2t::scan[0]);
		else memcpy(qmat_[0], m2t::default_intra_qmat, 64);
		if (br_.get1()) load_qmat(qmat_[1], m2t::scan[0]);
		else memset(qmat_[1], 16, 64);
		set_size(w, h);
		if (header_cb_) header_cb_(header_arg_, stream_->id);
		return 0;
	}

	int extension() {
		int id = br_.get(4);
		switch (id) {
		case 1: {   // sequence_extension
			br_.skip(8);  // profile_and_level
			br_.skip(1);  // progressive_sequence
			int chroma_format = br_.get(2);
			int w = width_ | (int)(br_.get(2) << 12);
			int h = height_ | (int)(br_.get(2) << 12);
			(void)chroma_format;
			set_size(w, h);
			is_mpeg2_ = true;
			if (header_cb_) header_cb_(header_arg_, stream_->id);
			break;
		}
		case 2: {   // sequence_display_extension
			int fmt = br_.get(4);
			if (fmt & 1) br_.skip(24);
			disp_w_ = br_.get(14);
			br_.skip(1);
			disp_h_ = br_.get(14);
			break;
		}
		case 3: {   // quant_matrix_extension: stored through the *current picture's* scan,
			        // as the reference does (mpeg2.cpp:351-360)
			const uint8_t *scan = m2t::scan[alternate_scan_];
			uint8_t dummy[64];
			for (int i = 0; i < 4; ++i) {
				if (br_.get1()) load_qmat(i < 2 ? qmat_[i] : dummy, scan);
			}
			break;
		}
		case 8: {   // picture_coding_extension
			f_code_[0][0] = br_.get(4);
			f_code_[0][1] = br_.get(4);
			f_code_[1][0] = br_.get(4);
			f_code_[1][1] = br_.get(4);
			intra_dc_precision_ = br_.get(2);
			picture_structure_ = br_.get(2);
			br_.skip(1);  // top_field_first
			frame_pred_frame_dct_ = br_.get1();
			concealment_mv_ = br_.get1();
			q_scale_type_ = br_.get1();
			intra_vlc_format_ = br_.get1();
			alternate_scan_ = br_.get1();
			break;
		}
		default:
			break;
		}
		return 0;
	}

	int picture_header() {
		temporal_ref_ = br_.get(10);
		pic_type_ = br_.get(3);
		br_.skip(16);
		mb_x_ = -1;
		mb_y_ = 0;
		return 0;
	}

	// ---- frame rotation (mpeg2.cpp:130-194) --------------------------------------------
	int pick_frame(int r0, int r1) {
		int best = -1, best_age = -1;
		for (int i = 0; i < num_frames_; ++i) {
			if (i == r0 || i == r1) continue;
			int age = lru_[i]++;
			if (age > best_age) { best_age = age; best = i; }
		}
		if (best < 0) best = r0;
		lru_[best] = 0;
		return best;
	}

	void begin_picture() {
		flush_picture();
		backend.ensure(B2_CODEC_MPEG2, mb_w_ * 16, mb_h_ * 16, num_frames_);
		bool anchor = (pic_type_ == I_PIC || pic_type_ == P_PIC);
		if (cur_idx_ < 0) {
			out_state_ = anchor ? 2 : 0;
			cur_idx_ = 0;
		} else {
			int cur = pick_frame(ref_idx_[0], ref_idx_[1]);
			if (anchor) {
				ref_idx_[0] = ref_idx_[1];
				ref_idx_[1] = cur;
				if (out_state_ < 4) out_state_ += 2;
			} else {
				out_state_ |= 1;
			}
			cur_idx_ = cur;
			frames_[cur].cnt = temporal_ref_;
		}
		pic_fwd_ = ref_idx_[0];
		pic_bwd_ = ref_idx_[1];
		mbs_.assign((size_t)mb_w_ * mb_h_, b2_m2_mb());   // flags = 0: macroblock left untouched
		if (cap_blocks_ < 6 * mbs_.size()) {
			cap_blocks_ = 6 * mbs_.size();
			coefs_.reset(new int16_t[64 * cap_blocks_]);
		}
		n_blocks_ = 0;
		picture_open_ = true;
		host_valid_[cur_idx_] = false;
	}

	void flush_picture() {
		if (!picture_open_) return;
		picture_open_ = false;
		b2_m2_picture p;
		memset(&p, 0, sizeof(p));
		p.mb_w = mb_w_;
		p.mb_h = mb_h_;
		p.pic_type = open_type_;
		p.dst = cur_idx_;
		p.fwd = pic_fwd_;
		p.bwd = pic_bwd_;
		p.n_mbs = (uint32_t)mbs_.size();
		p.n_blocks = (uint32_t)n_blocks_;
		p.mbs = mbs_.data();
		p.coefs = coefs_.get();
		backend.submit(B2_CODEC_MPEG2, &p);
	}

	// ---- slice / macroblock layer ------------------------------------------------------
	inline int qscale_from_code(int code) const {
		if (q_scale_type_) return code ? m2t::non_linear_qscale[code] : 1;
		return code ? code * 2 : 2;
	}

	int slice(int code) {
		if (!is_mpeg2_ || picture_structure_ != 3 || mb_w_ == 0) return 0;
		if (pic_type_ < I_PIC || pic_type_ > B_PIC) return 0;  // unsupported: ignore slice data
		if (num_frames_ == 0) return 0;
		if (mb_w_ <= 0 || mb_h_ <= 0 || mb_w_ != frames_mb_w_ || mb_h_ != frames_mb_h_) return 0;
		q_scale_ = qscale_from_code(br_.get(5));
		int row = code - 1;
		if (row == 0) {
			open_type_ = pic_type_;
			begin_picture();
		}
		if (!picture_open_) return 0;
		if (row >= mb_h_) return 0;
		if (row - mb_y_ > 1) {
			// rows with no slice are copied from the forward reference (mpeg2.cpp:644-649, 715-733)
			for (int y = mb_y_ + 1; y < row; ++y)
				for (int x = 0; x < mb_w_; ++x) emit_copy(x, y);
		}
		mb_x_ = -1;
		mb_y_ = row;
		if (br_.get1()) {
			br_.skip(8);
			while (br_.get1()) br_.skip(8);
		}
		return macroblocks();
	}

	void advance_mb() {
		int x = mb_x_ + 1;
		if (x >= mb_w_) {
			do { x -= mb_w_; mb_y_++; } while (mb_w_ < x);
		}
		mb_x_ = x;
	}

	bool at_last_mb() const {
		return (mb_y_ == mb_h_ - 1 && mb_x_ >= mb_w_ - 1) || mb_y_ >= mb_h_;
	}

	b2_m2_mb *cur_mb() {
		if (mb_y_ < 0 || mb_y_ >= mb_h_ || mb_x_ < 0 || mb_x_ >= mb_w_) return &scratch_mb_;
		return &mbs_[(size_t)mb_y_ * mb_w_ + mb_x_];
	}

	void emit_copy(int x, int y) {
		if (x < 0 || x >= mb_w_ || y < 0 || y >= mb_h_) return;
		b2_m2_mb &m = mbs_[(size_t)y * mb_w_ + x];
		memset(&m, 0, sizeof(m));
		m.flags = B2_M2_FWD;
	}

	void reset_dc_init() { dc_pred_[0] = dc_pred_[1] = dc_pred_[2] = 128; memset(pmv_, 0, sizeof(pmv_)); }
	void reset_dc() {
		int dc = 1 << (intra_dc_precision_ + 7);
		dc_pred_[0] = dc_pred_[1] = dc_pred_[2] = dc;
	}
	void reset_pmv() { memset(pmv_, 0, sizeof(pmv_)); }

	// macroblocks skipped by macroblock_address_increment (mpeg2.cpp:740-763, 773-808)
	void skipped(int increment) {
		int n = increment - 1;
		if (open_type_ == B_PIC) {
			int dir = prev_type_ & (m2t::MT_FWD | m2t::MT_BWD);
			bool bi = dir == (m2t::MT_FWD | m2t::MT_BWD);
			int d = bi ? 0 : (dir >> 1);
			while (n-- > 0) {
				advance_mb();
				b2_m2_mb *m = cur_mb();
				memset(m, 0, sizeof(*m));
				if (bi) {
					m->flags = B2_M2_FWD | B2_M2_BWD;
					m->mv[0][0][0] = pmv_[0][0][0]; m->mv[0][0][1] = pmv_[0][0][1];
					m->mv[1][0][0] = pmv_[1][0][0]; m->mv[1][0][1] = pmv_[1][0][1];
				} else {
					m->flags = d ? B2_M2_BWD : B2_M2_FWD;
					m->mv[d][0][0] = pmv_[d][0][0]; m->mv[d][0][1] = pmv_[d][0][1];
				}
			}
		} else {
			while (n-- > 0) {
				advance_mb();
				emit_copy(mb_x_, mb_y_);
			}
			reset_dc();
			reset_pmv();
		}
	}

	int macroblocks() {
		const M2Tables &t = tables();
		reset_dc();
		reset_pmv();
		do {
			int inc = 0;
			for (;;) {
				int v = t.mb_inc.decode(br_);
				if (v == Vlc::INVALID) throw BitError();
				if (v == 0) { inc += 33; continue; }
				inc += v;
				break;
			}
			if (inc > 1) skipped(inc);
			advance_mb();
			macroblock();
			if (br_.overrun()) throw BitError();
			if (at_last_mb()) {
				mb_x_ = -1;
				mb_y_ = 0;
				flush_picture();
				return 1;
			}
		} while (br_.bits_left() > 0 && br_.peek(23) != 0);
		return 0;
	}

	// one motion vector component (mpeg2.cpp:1189-1212)
	int mv_component(int16_t *pred, int f_code, int shift) {
		if (f_code == 0) throw BitError();   // forbidden value (6.3.10): r_size would be a negative shift count
		int r_size = f_code - 1;
		int code = tables().motion.decode(br_);
		if (code == Vlc::INVALID) throw BitError();
		int mv;
		if (code != 0) {
			int residual = r_size > 0 ? 1 + (int)br_.get(r_size) : 1;
			mv = code > 0 ? ((code - 1) << r_size) + residual : -(((-code - 1) << r_size) + residual);
			mv += *pred >> shift;
			int limit = 16 << r_size;
			if (mv < -limit) mv += 2 * limit;
			else if (mv >= limit) mv -= 2 * limit;
		} else {
			mv = *pred >> shift;
		}
		*pred = (int16_t)(mv * (1 << shift));   // (not mv << shift: mv may be negative)
		return mv;
	}

	struct MotionType { int count; bool field_format; bool dmv; };

	// motion_vectors(s) (mpeg2.cpp:1245-1275); out: mv[f][x,y], sel[f]
	void motion_vectors(int s, const MotionType &mt, int16_t mv[2][2], int sel[2]) {
		sel[0] = sel[1] = 0;
		if (mt.count == 1) {
			int shift = mt.field_format ? 1 : 0;
			mv[0][0] = (int16_t)mv_component(&pmv_[s][0][0], f_code_[s][0], 0);
			if (mt.dmv && br_.get1()) br_.skip(1);
			mv[0][1] = (int16_t)mv_component(&pmv_[s][0][1], f_code_[s][1], shift);
			if (mt.dmv && br_.get1()) br_.skip(1);
			pmv_[s][1][0] = pmv_[s][0][0];
			pmv_[s][1][1] = pmv_[s][0][1];
			mv[1][0] = mv[1][1] = 0;
		} else {
			for (int f = 0; f < 2; ++f) {
				sel[f] = br_.get1();
				mv[f][0] = (int16_t)mv_component(&pmv_[s][f][0], f_code_[s][0], 0);
				mv[f][1] = (int16_t)mv_component(&pmv_[s][f][1], f_code_[s][1], 1);
			}
		}
	}

	// Coefficient records of a whole macroblock at once: a bitstream error in the middle of the macroblock
	// then leaves zero blocks behind, never a cbp that claims records which do not exist.
	int16_t *new_blocks(int count) {
		// (one allocation per picture size, sized for every macroblock coding all six blocks: growing a vector block by
		// block cost a fifth of the parse of a stream's first pictures in reallocation copies and element-wise fills)
		if (n_blocks_ + (size_t)count > cap_blocks_) throw BitError();   // more blocks than the picture can hold: overlapping slices
		int16_t *p = coefs_.get() + 64 * n_blocks_;
		memset(p, 0, 128 * (size_t)count);
		n_blocks_ += (size_t)count;
		return p;
	}

	// run/level decode + inverse quantisation (mpeg2.cpp:1020-1097); idx = first scan position
	void coefficients(int16_t *blk, bool inter, int idx) {
		const Vlc &tab = tables().dct[inter ? 0 : intra_vlc_format_];
		const uint8_t *qm = qmat_[inter ? 1 : 0];
		const uint8_t *scan = m2t::scan[alternate_scan_];
		int sum = idx ? blk[0] : 0;
		// One look at the stream per coefficient: 32 bits hold the longest code (16 bits) with its sign, and the escape
		// (6) with run (6) and level (12).  Past the end of the unit the reader feeds zeros, which are an undefined
		// code in both tables, so the overrun is looked at once per block.
		for (;; ++idx) {
			const uint32_t w = br_.peek(32);
			int len;
			int v = tab.lookup(w, &len);
			if (v == Vlc::INVALID) throw BitError();
			if (v == m2t::DCT_EOB) { br_.skip(len); break; }
			int level, neg;
			if (v == m2t::DCT_ESC) {
				idx += (int)((w >> 20) & 63);
				int l = (int)((w >> 8) & 4095);
				neg = l >> 11;
				level = neg ? 4096 - l : l;
				br_.skip(24);
			} else {
				idx += v >> 6;
				level = v & 63;
				neg = (int)((w >> (31 - len)) & 1);
				br_.skip(len + 1);
			}
			if (idx >= 64) break;
			int pos = scan[idx];
			int q = qm[pos] * q_scale_;
			int val = inter ? ((2 * level + 1) * q) >> 5 : (level * q) >> 4;
			if (neg) val = -val;
			if (val > 2047) val = 2047;
			else if (val < -2048) val = -2048;
			sum += val;
			blk[pos] = (int16_t)val;
		}
		if (br_.overrun()) throw BitError();
		if (!(sum & 1)) blk[63] ^= 1;   // mismatch control (mpeg2.cpp:974-983)
	}

	int intra_dc(int comp) {
		const M2Tables &t = tables();
		int size = (comp == 0 ? t.dc_luma : t.dc_chroma).decode(br_);
		if (size == Vlc::INVALID) throw BitError();
		int dc = dc_pred_[comp];
		if (size) {
			int diff = (int)br_.get(size);
			int half = 1 << (size - 1);
			if (!(diff & half)) diff = diff + 1 - 2 * half;
			dc += diff;
			dc_pred_[comp] = (int16_t)dc;
			int dc_max = (1 << (intra_dc_precision_ + 8)) - 1;
			if (dc < 0) dc = 0;
			else if (dc > dc_max) dc = dc_max;
		}
		return dc * (1 << (3 - intra_dc_precision_));   // (a damaged stream can drive the predictor negative: no shift)
	}

	void macroblock() {
		const M2Tables &t = tables();
		bool prev_intra = (prev_type_ & m2t::MT_INTRA) != 0;
		const Vlc &tv = (open_type_ == I_PIC) ? t.type_i : (open_type_ == P_PIC ? t.type_p : t.type_b);
		int type = tv.decode(br_);
		if (type == Vlc::INVALID) throw BitError();
		prev_type_ = type;
		// macroblock_modes (mpeg2.cpp:836-873)
		MotionType mt = {1, false, false};
		if (type & (m2t::MT_FWD | m2t::MT_BWD)) {
			if (!frame_pred_frame_dct_) {
				int fmt = (int)br_.get(2);
				if (fmt == 2) mt = MotionType{1, false, false};
				else if (fmt == 3) mt = MotionType{1, true, true};
				else mt = MotionType{2, true, false};
			}
		}
		int dct_type = 0;
		if (!frame_pred_frame_dct_ && (type & (m2t::MT_PATTERN | m2t::MT_INTRA))) dct_type = (int)br_.get1();

		b2_m2_mb *m = cur_mb();
		memset(m, 0, sizeof(*m));
		if (type & m2t::MT_INTRA) {
			if (!prev_intra) reset_dc();
			if (type & m2t::MT_QUANT) q_scale_ = qscale_from_code((int)br_.get(5));
			if (concealment_mv_) {
				int16_t mv[2][2];
				int sel[2];
				motion_vectors(0, mt, mv, sel);
				br_.skip(1);
			}
			const uint32_t first = (uint32_t)n_blocks_;
			int16_t *blocks = new_blocks(6);      // (may refuse: before the record claims any block)
			m->flags = B2_M2_INTRA | (dct_type ? B2_M2_FIELD_DCT : 0);
			m->cbp = 0x3f;
			m->coef_idx = first;
			for (int k = 0; k < 6; ++k) {
				int comp = k < 4 ? 0 : k - 3;
				int dc = intra_dc(comp);
				int16_t *blk = blocks + 64 * k;
				blk[0] = (int16_t)dc;
				coefficients(blk, false, 1);
			}
			return;
		}
		if (prev_intra) reset_pmv();
		if (type & m2t::MT_QUANT) q_scale_ = qscale_from_code((int)br_.get(5));
		uint8_t flags = dct_type ? B2_M2_FIELD_DCT : 0;
		if (type & (m2t::MT_FWD | m2t::MT_BWD)) {
			if (mt.count == 2) flags |= B2_M2_FIELD_MC;
			for (int s = 0; s < 2; ++s) {
				if (!(type & (s ? m2t::MT_BWD : m2t::MT_FWD))) continue;
				int16_t mv[2][2];
				int sel[2];
				motion_vectors(s, mt, mv, sel);
				flags |= s ? B2_M2_BWD : B2_M2_FWD;
				for (int f = 0; f < 2; ++f) {
					m->mv[s][f][0] = mv[f][0];
					m->mv[s][f][1] = mv[f][1];
					if (sel[f]) m->field_sel |= (uint8_t)(1 << (s * 2 + f));
				}
			}
		} else {
			// P macroblock without motion_forward: co-located copy, predictors reset (mpeg2.cpp:1378-1380, 762)
			flags |= B2_M2_FWD;
			reset_dc();
			reset_pmv();
		}
		m->flags = flags;
		if (type & m2t::MT_PATTERN) {
			int cbp = t.cbp.decode(br_);
			if (cbp == Vlc::INVALID) throw BitError();
			const uint32_t first = (uint32_t)n_blocks_;
			int16_t *blk = new_blocks(__builtin_popcount((unsigned)cbp & 63u)) - 64;
			m->cbp = (uint8_t)cbp;
			m->coef_idx = first;
			for (int k = 0; k < 6; ++k) {
				if (!(cbp & (1 << (5 - k)))) continue;
				blk += 64;
				int idx = 0;
				if (br_.peek(1)) {   // "1s": first coefficient, run 0 level +-1 (mpeg2.cpp:1318-1340)
					int neg = (int)(br_.get(2) & 1);
					int v = (3 * qmat_[1][0] * q_scale_) >> 5;
					blk[0] = (int16_t)(neg ? -v : v);
					idx = 1;
				}
				coefficients(blk, true, idx);
			}
		}
	}

	// ---- state ---------------------------------------------------------------------------
	dec_bits *stream_;
	BitReader br_;
	int (*header_cb_)(void *, void *);
	void *header_arg_;
	int width_, height_, mb_w_, mb_h_, disp_w_, disp_h_;
	bool is_mpeg2_;
	uint8_t qmat_[2][64];
	int pic_type_, open_type_, temporal_ref_;
	int f_code_[2][2];
	int intra_dc_precision_, picture_structure_, frame_pred_frame_dct_, concealment_mv_;
	int q_scale_type_, intra_vlc_format_, alternate_scan_;
	int q_scale_;
	m2d_frame_t frames_[MAX_M2_FRAMES];
	bool host_valid_[MAX_M2_FRAMES];
	int num_frames_, cur_idx_, ref_idx_[2], lru_[MAX_M2_FRAMES], out_state_;
	int pic_fwd_, pic_bwd_;
	enum { MAX_OUTQ = 8 };
	int outq_[MAX_OUTQ], outq_n_;             // frames due, in output order (peek_frame)
	int mb_x_, mb_y_, prev_type_;
	int16_t dc_pred_[3];
	int16_t pmv_[2][2][2];
	bool picture_open_;
	int frames_mb_w_ = 0, frames_mb_h_ = 0;   // geometry set_frames was called for
	std::vector<b2_m2_mb> mbs_;
	std::unique_ptr<int16_t[]> coefs_;        // 64 coefficients per coded block, in macroblock order
	size_t n_blocks_ = 0, cap_blocks_ = 0;
	b2_m2_mb scratch_mb_;
};

inline Mpeg2Decoder *impl(void *ctx) {
	CtxBlob *b = (CtxBlob *)ctx;
	if (!b || b->magic != kCtxMagic || b->codec != B2_CODEC_MPEG2) return nullptr;
	return (Mpeg2Decoder *)b->impl;
}

int m2_init(void *ctx, int, int (*cb)(void *, void *), void *arg) {
	if (!ctx) return -1;
	CtxBlob *b = (CtxBlob *)ctx;
	memset(b, 0, sizeof(*b));
	b->magic = kCtxMagic;
	b->codec = B2_CODEC_MPEG2;
	dec_bits_open(&b->stream, nullptr);
	Mpeg2Decoder *d = new (std::nothrow) Mpeg2Decoder(&b->stream);
	if (!d) return -1;
	d->reset(cb, arg);
	b->impl = d;
	return 0;
}
dec_bits *m2_stream_pos(void *ctx) { return ctx ? &((CtxBlob *)ctx)->stream : nullptr; }
int m2_get_info(void *ctx, m2d_info_t *info) { Mpeg2Decoder *d = impl(ctx); return d ? d->get_info(info) : -1; }
int m2_set_frames(void *ctx, int n, m2d_frame_t *f, uint8_t *, int) { Mpeg2Decoder *d = impl(ctx); return d ? d->set_frames(n, f) : -1; }
int m2_decode(void *ctx) { Mpeg2Decoder *d = impl(ctx); return d ? d->decode_picture() : -1; }
int m2_peek(void *ctx, m2d_frame_t *f, int e) { Mpeg2Decoder *d = impl(ctx); return d ? d->peek_frame(f, e) : -1; }
int m2_get(void *ctx, m2d_frame_t *f, int e) { Mpeg2Decoder *d = impl(ctx); return d ? d->get_frame(f, e) : -1; }

const m2d_func_table_t m2_table = {
	sizeof(CtxBlob), m2_init, m2_stream_pos, m2_get_info, m2_set_frames, m2_decode, m2_peek, m2_get,
};

}  // namespace

Backend *mpeg2_backend(void *ctx) {
	Mpeg2Decoder *d = impl(ctx);
	return d ? &d->backend : nullptr;
}
void mpeg2_destroy(void *ctx) {
	Mpeg2Decoder *d = impl(ctx);
	if (d) { delete d; ((CtxBlob *)ctx)->impl = nullptr; }
}

}  // namespace b2

extern "C" const m2d_func_table_t *const m2d_gpu_func = &b2::m2_table;
