// h264_front.cpp — H.264 host front end: NAL / parameter-set / slice-header parse, reference
// list and DPB management, CAVLC macroblock parse with motion-vector / intra-mode prediction
// and deblocking-strength derivation.  It never touches pixels: every reconstruction call of
// the reference becomes a record of the per-picture SoA (include/b2_recon.h) that is handed
// to the device at picture end.  Exported behind the reference's plugin table as
// `h264d_gpu_func`.
//
// Behavioural model: reference src/lib/h264.cpp (h264d_func, :12057-12068).  The syntax
// parse is written from ITU-T H.264; the reference's observable behaviour is preserved where
// it decides the output: DPB insertion / bumping / output order (h264.cpp:695-867), frame
// buffer choice (924-962), POC derivation quirks (1121-1215), slice-header order (1417-1579),
// picture-end processing (11022-11050).  Scope: frame pictures; CAVLC and CABAC; I, P and B
// slices; 4x4 transform, and the 8x8 transform / Intra8x8 under CABAC (the reference's CAVLC
// 8x8 path is not usable).  Interlaced coding and CAVLC + 8x8 are rejected with
// an error, not approximated.
#include <algorithm>
#include <exception>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include <emmintrin.h>
#include "frontend.h"
#include "h264_cabac.h"
#include "h264_cavlc_tables.h"
#include "h264_common.h"

namespace b2 {
namespace {

using namespace h4;

enum { NAL_SLICE = 1, NAL_IDR = 5, NAL_SEI = 6, NAL_SPS = 7, NAL_PPS = 8 };
enum { SLICE_P = 0, SLICE_B = 1, SLICE_I = 2 };
enum { MAX_H264_FRAMES = 64 };
enum { kMaxMbDim = 1024, kMaxMbs = 139264 };   // 16384 samples per side; MaxFS of level 6.2

struct Sps {
	bool valid;
	int profile_idc, level_idc, constraint;
	int log2_max_frame_num, poc_type, log2_max_poc_lsb;
	int delta_pic_order_always_zero, offset_for_non_ref_pic, offset_for_top_to_bottom;
	int num_ref_frames_in_cycle;
	int offset_for_ref_frame[256];
	int num_ref_frames, gaps_allowed;
	int width, height;       // coded size, multiples of 16
	int frame_mbs_only, mbaff, direct_8x8_inference;
	int crop[4];
	int max_dpb_mbs;
};

struct Pps {
	bool valid;
	int sps_id;
	int cabac, pic_order_present;
	int num_ref_idx_active[2];
	int weighted_pred, weighted_bipred_idc;
	int pic_init_qp, chroma_qp_offset[2];
	int deblocking_control, constrained_intra, redundant_pic_cnt;
	int transform_8x8;
};

struct H264Tables {
	VlcSmall coeff_token[5], total_zeros[15], total_zeros_cdc[3], run_before[7];
	H264Tables() {
		for (int i = 0; i < 5; ++i) coeff_token[i].build(h4t::coeff_token_tabs[i].codes, h4t::coeff_token_tabs[i].n, 8);
		for (int i = 0; i < 15; ++i) total_zeros[i].build(h4t::total_zeros_tabs[i].codes, h4t::total_zeros_tabs[i].n, 9);
		for (int i = 0; i < 3; ++i) total_zeros_cdc[i].build(h4t::total_zeros_cdc_tabs[i].codes, h4t::total_zeros_cdc_tabs[i].n, 3);
		for (int i = 0; i < 7; ++i) run_before[i].build(h4t::run_before_tabs[i].codes, h4t::run_before_tabs[i].n, 6);
	}
};
const H264Tables &tables() {
	static const H264Tables t;
	return t;
}

struct BitError {};

// DPB element / state: restated from h264.h:301-313 and h264.cpp:695-867
struct DpbElem { int poc, frame_idx; bool is_idr, is_terminal; };
struct Dpb {
	int size, max, output;
	bool is_ready;
	DpbElem d[16];
	void init(int mx) { size = 0; max = mx; output = -1; is_ready = false; }
	bool contains(int idx) const {
		for (int i = 0; i < size; ++i) if (d[i].frame_idx == idx) return true;
		return false;
	}
	void insert_non_idr(int poc, int idx) {
		int pos;
		if (size > 0) {
			pos = size;
			do { --pos; } while (pos != 0 && !d[pos].is_terminal && poc < d[pos].poc);
			if (size < max) {
				output = -1;
				if (d[pos].is_terminal || d[pos].poc < poc) ++pos;
				for (int i = size; i > pos; --i) d[i] = d[i - 1];
				size++;
			} else {
				output = d[0].frame_idx;
				if (d[0].is_terminal) is_ready = false;
				for (int i = 0; i < pos; ++i) d[i] = d[i + 1];
			}
		} else {
			size = 1;
			output = -1;
			pos = 0;
		}
		d[pos].poc = poc; d[pos].frame_idx = idx; d[pos].is_idr = false; d[pos].is_terminal = false;
	}
	void insert_idr(int idx) {
		int s = size;
		if (s < max) {
			size = s + 1;
		} else {
			s--;
			output = d[0].frame_idx;
			if (d[0].is_terminal) is_ready = false;
			for (int i = 0; i < s; ++i) d[i] = d[i + 1];
		}
		d[s].poc = 0; d[s].frame_idx = idx; d[s].is_idr = true; d[s].is_terminal = false;
		if (s > 0) { d[s - 1].is_terminal = true; is_ready = true; }
	}
	int force_peek() const { return output >= 0 ? output : (size == 0 ? -1 : d[0].frame_idx); }
	int force_pop() {
		if (output >= 0) { int r = output; output = -1; return r; }
		if (size == 0) return -1;
		size--;
		output = -1;
		if (d[0].is_terminal) is_ready = false;
		int r = d[0].frame_idx;
		for (int i = 0; i < size; ++i) d[i] = d[i + 1];
		return r;
	}
};

struct SliceHeader {
	int first_mb, type, pps_id, frame_num, idr_pic_id;
	bool idr;
	int nal_ref_idc;
	int poc;
	int num_ref_idx_active[2];
	bool direct_spatial;
	int cabac_init_idc;
	int qp;
	int disable_deblock, alpha_off, beta_off;
	// dec_ref_pic_marking
	bool long_term_reference_flag, adaptive_marking, has_mmco5;
	struct Mmco { int op; unsigned a, b; } mmco[32];
	int n_mmco;
};

class H264Decoder {
public:
	explicit H264Decoder(dec_bits *stream) : stream_(stream), tab_(&tables()) { reset(-1, nullptr, nullptr); }
	Backend backend;

	void reset(int dpb_max, int (*cb)(void *, void *), void *arg) {
		header_cb_ = cb;
		header_arg_ = arg;
		memset(sps_, 0, sizeof(sps_));
		memset(pps_, 0, sizeof(pps_));
		active_sps_ = 0;
		dpb_.init(dpb_max);
		num_frames_ = 0;
		memset(frames_, 0, sizeof(frames_));
		memset(lru_, 0, sizeof(lru_));
		memset(host_valid_, 0, sizeof(host_valid_));
		cur_frame_ = 0;
		pic_open_ = false;
		prev_first_mb_ = UINT32_MAX;
		refs_.clear();
		prev_frame_num_ = 0;
		poc_lsb_ = poc_msb_ = 0;
		frame_num_offset_ = 0;
		prev_mmco5_ = false;
		mb_w_ = mb_h_ = 0;
		frames_w_ = frames_h_ = 0;
		unsupported_ = false;
		memset(&hdr_, 0, sizeof(hdr_));
		nal_ref_idc_ = nal_type_ = 0;
		cur_poc_ = 0;
		mb_addr_ = slice_num_ = 0;
		qp_ = 26;
		n_intra_ = 0;
		any_deblock_ = false;
		stop_bit_pos_ = total_bits_ = 0;
	}

	// ---- m2d_func_table_t entry points (h264.cpp:505-525, 647-659, 663-693, 817-867) --------
	int get_info(m2d_info_t *info) const {
		if (!info) return -1;
		const Sps &s = sps_[active_sps_];
		info->src_width = (int16_t)s.width;
		info->src_height = (int16_t)s.height;
		info->disp_width = (int16_t)s.width;
		info->disp_height = (int16_t)s.height;
		info->frame_num = (int16_t)(s.num_ref_frames + 1);
		for (int i = 0; i < 4; ++i) info->crop[i] = (int16_t)s.crop[i];
		info->additional_size = 16;   // the reference's scratch area is not needed here
		return 0;
	}

	int set_frames(int num, const m2d_frame_t *frames) {
		if (num < 3 || num > MAX_H264_FRAMES || !frames) return -1;
		num_frames_ = num;
		memcpy(frames_, frames, sizeof(frames_[0]) * num);
		memset(lru_, 0, sizeof(lru_));
		memset(host_valid_, 0, sizeof(host_valid_));
		const Sps &s = sps_[active_sps_];
		if (!s.valid) return -1;
		frames_w_ = s.width;     // the application sized its buffers for this SPS (get_info): pictures of any
		frames_h_ = s.height;    // other size are refused in slice(), they would overrun those buffers
		backend.bind_frames(frames_, num);
		// what the application's writer hashes of a frame of this SPS: pitch m2d_frame_t.width, the cropping window
		backend.set_md5_window(s.width, s.crop[0], s.crop[2], s.width - s.crop[0] - s.crop[1], s.height - s.crop[2] - s.crop[3]);
		return backend.configure(B2_CODEC_H264, s.width, s.height, num);
	}

	int decode_picture() {
		prev_first_mb_ = UINT32_MAX;
		for (;;) {
			if (backend.dead()) return -4;
			Unit u;
			if (next_unit(stream_, &u) < 0) return -2;
			if (u.len == 0) continue;
			int err = 0;
			try {
				err = nal(u.data, u.len);
			} catch (const BitError &) {
				return -2;     // bitstream error: the reference longjmps to its setjmp and returns -2
			} catch (const std::exception &e) {
				// allocation failure and the like: nothing may unwind through the extern "C" entry point
				fprintf(stderr, "m2dec_b200: H.264 decode_picture: %s\n", e.what());
				pic_open_ = false;
				return -1;
			}
			if (err != 0) return err;
		}
	}

	int peek_frame(m2d_frame_t *out, int bypass) {
		if (!out) return -1;
		int idx = (!bypass && !dpb_.is_ready) ? dpb_.output : dpb_.force_peek();
		if (idx < 0) return 0;
		if (fetch(idx) < 0) return -1;
		*out = frames_[idx];
		return 1;
	}
	int get_frame(m2d_frame_t *out, int bypass) {
		if (!out) return -1;
		int idx;
		if (!bypass && !dpb_.is_ready) { idx = dpb_.output; dpb_.output = -1; }
		else idx = dpb_.force_pop();
		if (idx < 0) return 0;
		if (fetch(idx) < 0) return -1;
		*out = frames_[idx];
		return 1;
	}

private:
	int fetch(int idx) {
		if (host_valid_[idx]) return 0;
		if (backend.fetch(idx, frames_[idx].luma, frames_[idx].chroma) < 0) return -1;
		host_valid_[idx] = true;
		return 0;
	}

	// ---- NAL layer -------------------------------------------------------------------------
	int nal(const uint8_t *data, size_t len) {
		int hdr = data[0];
		int type = hdr & 31;
		if (type != NAL_SLICE && type != NAL_IDR && type != NAL_SPS && type != NAL_PPS) return 0;
		// strip emulation prevention bytes (7.4.1.1) and trailing zero bytes
		rbsp_.resize(len);
		{
			// copy runs between 00 00 03 sequences; a candidate needs a zero at every second byte
			uint8_t *dst = rbsp_.data();
			size_t i = 1, run = 1;
			while (i + 2 < len) {
				if (data[i + 1] != 0) { i += 2; continue; }
				if (data[i] == 0 && data[i + 2] == 3) {
					memcpy(dst, data + run, i + 2 - run);
					dst += i + 2 - run;
					i += 3;
					run = i;
				} else {
					i += 1;
				}
			}
			memcpy(dst, data + run, len - run);
			dst += len - run;
			rbsp_.resize((size_t)(dst - rbsp_.data()));
		}
		while (!rbsp_.empty() && rbsp_.back() == 0) rbsp_.pop_back();
		br_.init(rbsp_.data(), rbsp_.size());
		// position of the rbsp_stop_one_bit, for more_rbsp_data()
		stop_bit_pos_ = 0;
		if (!rbsp_.empty()) {
			uint8_t last = rbsp_.back();
			int tz = 0;
			while (!((last >> tz) & 1)) tz++;
			stop_bit_pos_ = (long)rbsp_.size() * 8 - tz - 1;
		}
		total_bits_ = (long)rbsp_.size() * 8;
		switch (type) {
		case NAL_SPS: {
			int id = parse_sps();
			if (id < 0) return 0;      // reference: a bad SPS ends decode_picture only through err<0 on slices
			active_sps_ = id;
			if (pic_open_ && (mb_w_ != sps_[id].width >> 4 || mb_h_ != sps_[id].height >> 4)) {
				// the size changes under an unfinished picture: drop it, its tables have the old geometry
				pic_open_ = false;
				prev_first_mb_ = UINT32_MAX;
			}
			mb_w_ = sps_[id].width >> 4;
			mb_h_ = sps_[id].height >> 4;
			if (header_cb_) header_cb_(header_arg_, stream_->id);
			return 0;
		}
		case NAL_PPS:
			parse_pps();
			return 0;
		default:
			nal_ref_idc_ = (hdr >> 5) & 3;
			nal_type_ = type;
			return slice();
		}
	}
	inline bool more_rbsp_data() { return total_bits_ - br_.bits_left() < stop_bit_pos_; }

	void skip_scaling_list(int size) {
		int last = 8, next = 8;
		for (int i = 0; i < size; ++i) {
			if (next != 0) next = (last + br_.se() + 256) & 255;
			last = next == 0 ? last : next;
		}
	}

	int parse_sps() {
		Sps s;
		memset(&s, 0, sizeof(s));
		s.profile_idc = br_.get(8);
		s.constraint = br_.get(8);
		s.level_idc = br_.get(8);
		unsigned id = br_.ue();
		if (id > 31) return -1;
		int p = s.profile_idc;
		if (p == 44 || p == 83 || p == 86 || p == 100 || p == 110 || p == 118 || p == 128 || p == 122 || p == 244) {
			unsigned chroma = br_.ue();
			if (chroma == 3) br_.skip(1);
			unsigned depth_y = br_.ue(), depth_c = br_.ue();
			// the reconstruction kernels are 4:2:0, 8 bit (as the reference, which ignores these fields and
			// would decode garbage)
			if (chroma != 1 || depth_y != 0 || depth_c != 0) return -1;
			br_.skip(1);
			if (br_.get1()) {
				// scaling lists are parsed and discarded, as in the reference (h264.cpp:254-305)
				int n8 = chroma != 3 ? 2 : 6;
				for (int i = 0; i < 6; ++i) if (br_.get1()) skip_scaling_list(16);
				for (int i = 0; i < n8; ++i) if (br_.get1()) skip_scaling_list(64);
			}
		}
		// ranges of 7.4.2.1.1: out-of-range values would become shift counts / read lengths further down
		unsigned l2fn = br_.ue(), poc_type = br_.ue();
		if (l2fn > 12 || poc_type > 2) return -1;
		s.log2_max_frame_num = (int)l2fn + 4;
		s.poc_type = (int)poc_type;
		if (s.poc_type == 0) {
			unsigned l2poc = br_.ue();
			if (l2poc > 12) return -1;
			s.log2_max_poc_lsb = (int)l2poc + 4;
		} else if (s.poc_type == 1) {
			s.delta_pic_order_always_zero = br_.get1();
			s.offset_for_non_ref_pic = br_.se();
			s.offset_for_top_to_bottom = br_.se();
			s.num_ref_frames_in_cycle = br_.ue() & 255;
			int acc = 0;
			for (int i = 0; i < s.num_ref_frames_in_cycle; ++i) { acc += br_.se(); s.offset_for_ref_frame[i] = acc; }
		}
		unsigned nref = br_.ue();
		if (nref > 16) return -1;
		s.num_ref_frames = (int)nref;
		s.gaps_allowed = br_.get1();
		// picture size: m2d_info_t carries int16 sizes (m2d.h:49-57), Table A-1 ends at MaxFS = 139264 macroblocks
		unsigned wmb = br_.ue(), hmb = br_.ue();
		if (wmb >= kMaxMbDim || hmb >= kMaxMbDim || (uint64_t)(wmb + 1) * (hmb + 1) > kMaxMbs) return -1;
		s.width = (int)(wmb + 1) * 16;
		s.height = (int)(hmb + 1) * 16;
		s.frame_mbs_only = br_.get1();
		if (!s.frame_mbs_only) s.mbaff = br_.get1();
		s.direct_8x8_inference = br_.get1();
		if (br_.get1()) {
			unsigned c[4];
			for (int i = 0; i < 4; ++i) c[i] = br_.ue();
			if ((uint64_t)c[0] + c[1] >= (unsigned)s.width / 2 || (uint64_t)c[2] + c[3] >= (unsigned)s.height / 2) return -1;
			for (int i = 0; i < 4; ++i) s.crop[i] = (int)c[i] * 2;
		}
		// VUI is not needed for reconstruction
		s.max_dpb_mbs = max_dpb_mbs(s.profile_idc, s.level_idc, s.constraint);
		if (br_.overrun()) return -1;
		s.valid = true;
		sps_[id] = s;
		return (int)id;
	}

	static int max_dpb_mbs(int profile, int level, int constraint) {
		// Table A-1 MaxDpbMbs, as tabulated by the reference (h264.cpp:191-246)
		if (profile == 100 && level == 9) level = 10;
		(void)constraint;
		switch (level) {
		case 10: return 396;
		case 11: return 900;
		case 12: case 13: case 20: return 2376;
		case 21: return 4752;
		case 22: case 30: return 8100;
		case 31: return 18000;
		case 32: return 20480;
		case 40: case 41: return 32768;
		case 42: return 34816;
		case 50: return 110400;
		case 51: return 184320;
		default: return -1;
		}
	}

	void parse_pps() {
		unsigned id = br_.ue();
		if (id > 255) return;
		Pps p;
		memset(&p, 0, sizeof(p));
		p.sps_id = br_.ue();
		if (p.sps_id > 31) return;
		p.cabac = br_.get1();
		p.pic_order_present = br_.get1();
		if (br_.ue() > 0) return;   // FMO: not supported by the reference either (h264.cpp:418-421)
		unsigned n0 = br_.ue(), n1 = br_.ue();
		if (n0 > 31 || n1 > 31) return;
		p.num_ref_idx_active[0] = (int)n0 + 1;
		p.num_ref_idx_active[1] = (int)n1 + 1;
		p.weighted_pred = br_.get1();
		p.weighted_bipred_idc = br_.get(2);
		if (p.weighted_bipred_idc > 2) return;
		p.pic_init_qp = br_.se() + 26;
		if (p.pic_init_qp < 0 || p.pic_init_qp > 51) return;
		br_.se();
		p.chroma_qp_offset[0] = p.chroma_qp_offset[1] = br_.se();
		if (p.chroma_qp_offset[0] < -12 || p.chroma_qp_offset[0] > 12) return;
		p.deblocking_control = br_.get1();
		p.constrained_intra = br_.get1();
		p.redundant_pic_cnt = br_.get1();
		if (more_rbsp_data()) {
			p.transform_8x8 = br_.get1();
			if (br_.get1()) { /* pic scaling matrix: flags only are consumed by the reference (h264.cpp:437-438) */ }
			p.chroma_qp_offset[1] = br_.se();
			if (p.chroma_qp_offset[1] < -12 || p.chroma_qp_offset[1] > 12) return;
		}
		p.valid = !br_.overrun();
		pps_[id] = p;
	}

	// ---- frame buffer choice (h264.cpp:924-962) ----------------------------------------------
	void find_empty_frame() {
		for (int i = 0; i < num_frames_; ++i) {
			if (dpb_.contains(i)) lru_[i] = 0;
			else lru_[i] = (int8_t)(lru_[i] + 1);
		}
		for (const RefPic &r : refs_) lru_[r.frame_idx] = 0;
		int best = 0, best_val = -1;
		for (int i = 0; i < num_frames_; ++i)
			if (lru_[i] > best_val) { best_val = lru_[i]; best = i; }
		lru_[best] = 0;
		cur_frame_ = best;
	}

	// ---- slice header (h264.cpp:1417-1579) -----------------------------------------------------
	int slice() {
		if (num_frames_ == 0 || unsupported_) return unsupported_ ? -3 : 0;
		SliceHeader &h = hdr_;
		uint32_t first_mb = br_.ue();
		bool new_picture = first_mb <= prev_first_mb_;
		if (new_picture) {
			if (prev_first_mb_ != UINT32_MAX) return -2;
		}
		int st = br_.ue();
		if (st > 9) return -1;
		if (st > 4) st -= 5;
		if (st > 2) return -1;
		unsigned pps_id = br_.ue();
		if (pps_id > 255 || !pps_[pps_id].valid) return -1;
		const Pps &pps = pps_[pps_id];
		const Sps &sps = sps_[pps.sps_id];
		if (!sps.valid) return -1;
		if ((pps.transform_8x8 && !pps.cabac) || !sps.frame_mbs_only) {
			fprintf(stderr, "m2dec_b200: H.264 feature outside this revision (CAVLC with 8x8 transform / interlace)\n");
			unsupported_ = true;
			return -3;
		}
		// the application's frame buffers (and the device pool) were sized by set_frames for frames_w_ x frames_h_
		if (sps.width != frames_w_ || sps.height != frames_h_) return -1;
		// first_mb_in_slice is ue(v): up to 2^32 - 2.  It becomes the macroblock address every table of the
		// picture is indexed with
		if (first_mb >= (uint32_t)((sps.width >> 4) * (sps.height >> 4))) return -1;
		if (new_picture) begin_picture(sps);
		prev_first_mb_ = first_mb;
		h.first_mb = (int)first_mb;
		h.type = st;
		h.pps_id = (int)pps_id;
		h.frame_num = br_.get(sps.log2_max_frame_num);
		h.idr = nal_type_ == NAL_IDR;
		h.nal_ref_idc = nal_ref_idc_;
		if (h.idr) h.idr_pic_id = br_.ue();
		// picture order count (h264.cpp:1121-1215)
		int delta_poc_bottom = 0, delta_poc[2] = {0, 0};
		int lsb = 0;
		if (sps.poc_type == 0) {
			lsb = br_.get(sps.log2_max_poc_lsb);
			if (pps.pic_order_present) delta_poc_bottom = br_.se();
		} else if (sps.poc_type == 1 && !sps.delta_pic_order_always_zero) {
			delta_poc[0] = br_.se();
			if (pps.pic_order_present) delta_poc[1] = br_.se();
		}
		(void)delta_poc_bottom;
		if (first_mb == 0) compute_poc(sps, h, lsb, delta_poc);
		h.poc = cur_poc_;
		frames_[cur_frame_].cnt = cur_poc_;
		if (pps.redundant_pic_cnt) br_.ue();
		h.num_ref_idx_active[0] = pps.num_ref_idx_active[0];
		h.num_ref_idx_active[1] = pps.num_ref_idx_active[1];
		h.direct_spatial = true;
		wt_.mode = 0;
		if (st == SLICE_B) h.direct_spatial = br_.get1() != 0;
		if (st != SLICE_I) {
			if (br_.get1()) {
				h.num_ref_idx_active[0] = br_.ue() + 1;
				if (st == SLICE_B) h.num_ref_idx_active[1] = br_.ue() + 1;
			}
			if (h.num_ref_idx_active[0] > 32 || h.num_ref_idx_active[1] > 32) return -1;
			build_lists(sps, h, st == SLICE_B);
			if ((st == SLICE_P && pps.weighted_pred) || (st == SLICE_B && pps.weighted_bipred_idc == 1)) pred_weight_table(h, st == SLICE_B);
			else if (st == SLICE_B && pps.weighted_bipred_idc == 2) wt_.mode = 2;
		}
		h.adaptive_marking = false;
		h.long_term_reference_flag = false;
		h.n_mmco = 0;
		h.has_mmco5 = false;
		if (nal_ref_idc_) {
			if (h.idr) {
				br_.skip(1);  // no_output_of_prior_pics_flag
				h.long_term_reference_flag = br_.get1();
			} else if ((h.adaptive_marking = br_.get1())) {
				for (;;) {
					int op = br_.ue();
					if (op == 0 || h.n_mmco >= 32) break;
					SliceHeader::Mmco &m = h.mmco[h.n_mmco++];
					m.op = op;
					m.a = m.b = 0;
					if (op == 1 || op == 3) m.a = br_.ue();
					if (op == 2) m.a = br_.ue();
					if (op == 3 || op == 6) m.b = br_.ue();
					if (op == 4) m.a = br_.ue();
					if (op == 5) h.has_mmco5 = true;
					if (br_.overrun()) throw BitError();
				}
			}
		}
		h.cabac_init_idc = 0;
		if (pps.cabac && st != SLICE_I) {
			h.cabac_init_idc = (int)br_.ue();
			if (h.cabac_init_idc > 2) return -1;
		}
		h.qp = pps.pic_init_qp + br_.se();
		if (h.qp < 0 || h.qp > 51) return -1;    // SliceQPY indexes the dequantisation and filter tables
		h.disable_deblock = 0;
		h.alpha_off = h.beta_off = 0;
		if (pps.deblocking_control) {
			h.disable_deblock = br_.ue();
			if (h.disable_deblock > 2) return -1;
			if (h.disable_deblock != 1) {
				int a = br_.se(), b = br_.se();
				if (a < -6 || a > 6 || b < -6 || b > 6) return -1;
				h.alpha_off = a * 2;
				h.beta_off = b * 2;
			}
		}
		if (br_.overrun()) throw BitError();
		slice_num_++;
		bslice_ = st == SLICE_B;
		// (tools the Baseline profile does not have: a stream that claims it and uses them is damaged)
		if (dev_mv_ && (bslice_ || pps.cabac || wt_.mode != 0)) return -1;
		direct_8x8_inference_ = sps.direct_8x8_inference != 0;
		cabac_ = pps.cabac != 0;
		ctx_.cabac = cabac_;
		if (cabac_) slice_data_cabac(pps);
		else slice_data(sps, pps);
		// picture end (h264.cpp:11022-11050)
		if (mb_addr_ >= mb_w_ * mb_h_) {
			end_picture(sps);
			return 1;
		}
		return 0;
	}

	void compute_poc(const Sps &sps, const SliceHeader &h, int lsb, const int delta_poc[2]) {
		bool reset = h.idr || prev_mmco5_;
		if (sps.poc_type == 0) {
			int prev_lsb = reset ? 0 : poc_lsb_, prev_msb = reset ? 0 : poc_msb_;
			int half = (1 << sps.log2_max_poc_lsb) >> 1, msb;
			if (lsb < prev_lsb && half <= prev_lsb - lsb) msb = prev_msb + half * 2;
			else if (prev_lsb < lsb && half < lsb - prev_lsb) msb = prev_msb - half * 2;
			else msb = prev_msb;
			poc_lsb_ = lsb;
			poc_msb_ = msb;
			cur_poc_ = msb + lsb;
		} else {
			if (reset) frame_num_offset_ = 0;
			else if (h.frame_num < prev_frame_num_) frame_num_offset_ += 1 << sps.log2_max_frame_num;
			if (sps.poc_type == 2) {
				cur_poc_ = (h.frame_num + frame_num_offset_) * 2 - (nal_ref_idc_ == 0);
			} else {
				int poc = 0;
				if (sps.num_ref_frames_in_cycle) {
					unsigned fn = (unsigned)(h.frame_num + frame_num_offset_);
					if (fn != 0) {
						int cycle_sum = sps.offset_for_ref_frame[sps.num_ref_frames_in_cycle - 1];
						int cnt = 0;
						fn--;
						if (fn != 0 && !nal_ref_idc_) fn--;
						while (cycle_sum > 0 && cycle_sum <= (int)fn) { fn -= cycle_sum; cnt++; }
						poc = cnt * cycle_sum + sps.offset_for_ref_frame[fn & 255];
					} else {
						poc = sps.offset_for_ref_frame[0];
					}
					if (!nal_ref_idc_) poc += sps.offset_for_non_ref_pic;
				}
				cur_poc_ = poc + delta_poc[0];
			}
		}
	}

	// initial lists (8.2.4.2.1 P, 8.2.4.2.3 B) + modification (8.2.4.3)
	void build_lists(const Sps &sps, SliceHeader &h, bool bslice) {
		const int max_fn = 1 << sps.log2_max_frame_num;
		std::vector<RefPic> st, lt;
		for (const RefPic &r : refs_) (r.long_idx < 0 ? st : lt).push_back(r);
		auto picnum = [&](const RefPic &r) { return r.frame_num > h.frame_num ? r.frame_num - max_fn : r.frame_num; };
		std::sort(lt.begin(), lt.end(), [](const RefPic &a, const RefPic &b) { return a.long_idx < b.long_idx; });
		const RefPic none = {-1, -1, -1, 0};
		for (int lx = 0; lx < (bslice ? 2 : 1); ++lx) {
			std::vector<RefPic> list;
			if (!bslice) {
				list = st;
				std::sort(list.begin(), list.end(), [&](const RefPic &a, const RefPic &b) { return picnum(a) > picnum(b); });
			} else {
				std::vector<RefPic> before, after;
				for (const RefPic &r : st) (r.poc < cur_poc_ ? before : after).push_back(r);
				std::sort(before.begin(), before.end(), [](const RefPic &a, const RefPic &b) { return a.poc > b.poc; });
				std::sort(after.begin(), after.end(), [](const RefPic &a, const RefPic &b) { return a.poc < b.poc; });
				if (lx == 0) { list = before; list.insert(list.end(), after.begin(), after.end()); }
				else { list = after; list.insert(list.end(), before.begin(), before.end()); }
			}
			list.insert(list.end(), lt.begin(), lt.end());
			if (lx == 1 && list.size() > 1 && init0_.size() == list.size()) {
				// identical initial lists: the first two entries of list 1 are switched (8.2.4.2.3)
				bool same = true;
				for (size_t i = 0; i < list.size(); ++i) same &= init0_[i].frame_idx == list[i].frame_idx;
				if (same) std::swap(list[0], list[1]);
			}
			if (lx == 0) init0_ = list;
			const int n = h.num_ref_idx_active[lx];
			if ((int)list.size() > n) list.resize(n);
			list.resize(n, none);
			if (br_.get1()) {   // ref_pic_list_modification_flag_lX
				int pred = h.frame_num, idx = 0;
				for (;;) {
					unsigned op = br_.ue();
					if (op == 3) break;
					if (op > 3 || idx >= n) throw BitError();
					unsigned v = br_.ue();
					RefPic target = none;
					if (op < 2) {
						int d = (int)v + 1;
						pred = op == 0 ? pred - d : pred + d;
						if (pred < 0) pred += max_fn;
						if (pred >= max_fn) pred -= max_fn;
						int pn = pred > h.frame_num ? pred - max_fn : pred;
						for (const RefPic &r : st) if (picnum(r) == pn) target = r;
					} else {
						for (const RefPic &r : lt) if (r.long_idx == (int)v) target = r;
					}
					list.insert(list.begin() + idx, target);
					idx++;
					int w = idx;
					for (int k = idx; k < (int)list.size(); ++k) {
						bool same = list[k].frame_idx == target.frame_idx && target.frame_idx >= 0;
						if (!same) list[w++] = list[k];
					}
					list.resize(w);
					list.resize(n, none);
				}
			}
			list_[lx] = list;
		}
		if (!bslice) list_[1].clear();
	}

	// pred_weight_table (7.3.3.2; reference h264.cpp:1668-1695): weights are int8 in the reference too
	void pred_weight_table(SliceHeader &h, bool bslice) {
		wt_.mode = 1;
		wt_.denom[0] = (int)br_.ue();
		wt_.denom[1] = (int)br_.ue();
		if (wt_.denom[0] > 7 || wt_.denom[1] > 7) throw BitError();
		for (int lx = 0; lx < (bslice ? 2 : 1); ++lx)
			for (int i = 0; i < h.num_ref_idx_active[lx]; ++i) {
				int8_t *w = wt_.w[lx][i], *o = wt_.o[lx][i];
				if (br_.get1()) { w[0] = (int8_t)br_.se(); o[0] = (int8_t)br_.se(); }
				else { w[0] = (int8_t)(1 << wt_.denom[0]); o[0] = 0; }
				if (br_.get1()) {
					for (int c = 1; c < 3; ++c) { w[c] = (int8_t)br_.se(); o[c] = (int8_t)br_.se(); }
				} else {
					w[1] = w[2] = (int8_t)(1 << wt_.denom[1]);
					o[1] = o[2] = 0;
				}
			}
	}

	static int dist_scale_factor(int poc0, int poc1, int cur) {
		if (poc1 == poc0) return 256;
		int td = poc1 - poc0, tb = cur - poc0;
		td = td < -128 ? -128 : (td > 127 ? 127 : td);
		tb = tb < -128 ? -128 : (tb > 127 ? 127 : tb);
		int tx = (16384 + abs(td / 2)) / td;
		return (tb * tx + 32) >> 6;
	}

	// ---- picture begin / end ------------------------------------------------------------------
	void begin_picture(const Sps &sps) {
		find_empty_frame();
		mb_w_ = sps.width >> 4;
		mb_h_ = sps.height >> 4;
		if (dpb_.max < 0) {
			int n = sps.max_dpb_mbs / (mb_w_ * mb_h_);
			dpb_.max = n > 16 ? 16 : n;
		}
		m2d_frame_t &f = frames_[cur_frame_];
		f.width = (int16_t)sps.width;
		f.height = (int16_t)sps.height;
		for (int i = 0; i < 4; ++i) f.crop[i] = (int16_t)sps.crop[i];
		backend.ensure(B2_CODEC_H264, sps.width, sps.height, num_frames_);
		// Deblocking strengths: derived on the device from the motion records unless someone looks at
		// the records on the host -- a picture tap (the CPU oracle of tests/, bench.py's capture for the kernel replay),
		// parse-only runs -- or the application asked for it (b2_context_set_host_strengths; B2_HOST_BS=1 for A/B runs).
		static const bool env_host_bs = getenv("B2_HOST_BS") != nullptr;
		static const bool env_as_device = getenv("B2_PARSE_AS_DEVICE") != nullptr;      // (tools/parsebench: time the parse of the product path)
		host_bs_ = env_host_bs || backend.host_bs() || ((backend.has_tap() || backend.device_off()) && !env_as_device);
		// Motion vectors: Baseline-profile streams (P and I slices only, no weighted prediction, no
		// co-located motion to keep for later direct prediction) hand the device motion vector DIFFERENCES; prediction,
		// motion records and the inter kernel's work list are made there.  Needs the device-side strengths.
		static const bool env_host_mv = getenv("B2_HOST_MV") != nullptr;
		dev_mv_ = !host_bs_ && !env_host_mv && sps.profile_idc == 66;
		keep_parts_ = !dev_mv_ && backend.has_tap() && sps.profile_idc == 66;
		parts_.clear();
		item_bound_ = 0;
		if (ctx_.mb_w != mb_w_ || ctx_.mb_h != mb_h_) ctx_.resize(mb_w_, mb_h_);
		ctx_.begin_picture(!dev_mv_);
		size_t n = (size_t)mb_w_ * mb_h_;
		mbs_.assign(n, b2_h264_mb());
		dbk_.assign(n, b2_h264_dbk());
		mb_seen_.assign(n, 0);
		mcs_.clear();
		for (auto &b : items_) b.clear();
		for (auto &b : items_first_) b.clear();
		coef_used_ = 0;
		n_intra_ = 0;
		any_deblock_ = false;
		slice_num_ = -1;
		mb_addr_ = 0;
		pic_open_ = true;
		host_valid_[cur_frame_] = false;
	}

	void end_picture(const Sps &sps) {
		const SliceHeader &h = hdr_;
		b2_h264_picture p;
		memset(&p, 0, sizeof(p));
		p.mb_w = mb_w_;
		p.mb_h = mb_h_;
		p.dst = cur_frame_;
		p.deblock = any_deblock_;
		p.n_mbs = (uint32_t)mbs_.size();
		p.n_mcs = (uint32_t)mcs_.size();
		p.n_coef_recs = (uint32_t)(coef_used_ / 16);
		p.n_intra = n_intra_;
		p.mbs = mbs_.data();
		p.dbk = dbk_.data();
		p.mcs = mcs_.data();
		p.coefs = coefs_.data();
		items_all_.clear();
		for (const auto &b : items_first_) items_all_.insert(items_all_.end(), b.begin(), b.end());
		p.n_items_first = (uint32_t)items_all_.size();
		for (const auto &b : items_) items_all_.insert(items_all_.end(), b.begin(), b.end());
		p.items = items_all_.data();
		p.n_items = (uint32_t)items_all_.size();
		p.bs_on_device = host_bs_ ? 0u : 1u;
		if (keep_parts_ && parts_.size() == mcs_.size()) {       // host derivation, partitions alongside (1:1 with the records)
			p.n_parts = (uint32_t)parts_.size();
			p.parts = parts_.data();
		}
		if (dev_mv_) {
			p.mv_on_device = 1;
			p.n_parts = (uint32_t)parts_.size();
			p.parts = parts_.data();
			p.mcs = nullptr;
			p.items = nullptr;
			p.n_mcs = 0;
			p.n_items = item_bound_;
			p.n_items_first = 0;
		}
		backend.submit(B2_CODEC_H264, &p);
		pic_open_ = false;
		// reference marking (8.2.5) and DPB insertion (h264.cpp:10840-10873, 11036-11045)
		int max_fn = 1 << sps.log2_max_frame_num;
		bool mmco5 = false;
		// co-located motion for later direct prediction; Baseline streams have no B slices to use it
		if (h.nal_ref_idc && sps.profile_idc != 66) ctx_.store_col(col_[cur_frame_], [&](int f) { return f >= 0 ? (int)frames_[f].cnt : 0; });
		if (h.nal_ref_idc) {
			if (h.idr) {
				refs_.clear();
				refs_.push_back(RefPic{cur_frame_, h.frame_num, h.long_term_reference_flag ? 0 : -1, cur_poc_});
			} else {
				if (h.adaptive_marking) mmco5 = apply_mmco(h, max_fn);
				bool cur_long = false;
				for (int i = 0; i < h.n_mmco; ++i) if (h.mmco[i].op == 6) cur_long = true;
				if (!cur_long) {
					// sliding window (8.2.5.3)
					int n_short = 0;
					for (const RefPic &r : refs_) if (r.long_idx < 0) n_short++;
					if ((int)refs_.size() >= std::max(sps.num_ref_frames, 1) && n_short > 0) {
						int victim = -1, best = INT32_MAX;
						for (size_t i = 0; i < refs_.size(); ++i) {
							if (refs_[i].long_idx >= 0) continue;
							int fnw = refs_[i].frame_num > h.frame_num ? refs_[i].frame_num - max_fn : refs_[i].frame_num;
							if (fnw < best) { best = fnw; victim = (int)i; }
						}
						if (victim >= 0) refs_.erase(refs_.begin() + victim);
					}
					refs_.push_back(RefPic{cur_frame_, mmco5 ? 0 : h.frame_num, -1, mmco5 ? 0 : cur_poc_});
				}
			}
			if (h.idr || mmco5) dpb_.insert_idr(cur_frame_);
			else dpb_.insert_non_idr(cur_poc_, cur_frame_);
		} else {
			dpb_.insert_non_idr(cur_poc_, cur_frame_);
		}
		prev_mmco5_ = mmco5;
		prev_frame_num_ = mmco5 ? 0 : h.frame_num;
		prev_first_mb_ = (uint32_t)(mb_w_ * mb_w_);   // sic: h264.cpp:11047
	}

	bool apply_mmco(const SliceHeader &h, int max_fn) {
		bool op5 = false;
		auto picnum = [&](const RefPic &r) { return r.frame_num > h.frame_num ? r.frame_num - max_fn : r.frame_num; };
		for (int i = 0; i < h.n_mmco; ++i) {
			const SliceHeader::Mmco &m = h.mmco[i];
			switch (m.op) {
			case 1: {
				int pn = h.frame_num - ((int)m.a + 1);
				for (size_t k = 0; k < refs_.size(); ++k)
					if (refs_[k].long_idx < 0 && picnum(refs_[k]) == pn) { refs_.erase(refs_.begin() + k); break; }
				break;
			}
			case 2:
				for (size_t k = 0; k < refs_.size(); ++k)
					if (refs_[k].long_idx == (int)m.a) { refs_.erase(refs_.begin() + k); break; }
				break;
			case 3: {
				int pn = h.frame_num - ((int)m.a + 1);
				for (size_t k = 0; k < refs_.size(); ++k)
					if (refs_[k].long_idx == (int)m.b) { refs_.erase(refs_.begin() + k); break; }
				for (RefPic &r : refs_)
					if (r.long_idx < 0 && picnum(r) == pn) { r.long_idx = (int)m.b; break; }
				break;
			}
			case 4:
				for (size_t k = 0; k < refs_.size();)
					if (refs_[k].long_idx >= (int)m.a) refs_.erase(refs_.begin() + k); else ++k;
				break;
			case 5:
				refs_.clear();
				op5 = true;
				break;
			case 6:
				for (size_t k = 0; k < refs_.size(); ++k)
					if (refs_[k].long_idx == (int)m.b) { refs_.erase(refs_.begin() + k); break; }
				refs_.push_back(RefPic{cur_frame_, h.frame_num, (int)m.b, cur_poc_});
				break;
			}
		}
		return op5;
	}

	// ---- slice data (CAVLC) ------------------------------------------------------------------------
	void slice_data(const Sps &sps, const Pps &pps) {
		(void)sps;
		const SliceHeader &h = hdr_;
		qp_ = h.qp;
		mb_addr_ = h.first_mb;
		int total = mb_w_ * mb_h_;
		bool more = true;
		while (more && mb_addr_ < total) {
			if (h.type != SLICE_I) {
				uint32_t run = br_.ue();
				while (run-- > 0 && mb_addr_ < total) { skip_mb(pps); mb_addr_++; }
				if (mb_addr_ >= total) break;
				if (!more_rbsp_data()) break;
			}
			macroblock(pps);
			mb_addr_++;
			if (br_.overrun()) throw BitError();
			more = more_rbsp_data();
		}
	}

	// ---- slice data (CABAC, 7.3.4 with entropy_coding_mode_flag = 1) -------------------------------
	void slice_data_cabac(const Pps &pps) {
		const SliceHeader &h = hdr_;
		qp_ = h.qp;
		mb_addr_ = h.first_mb;
		const int total = mb_w_ * mb_h_;
		br_.byte_align();     // cabac_alignment_one_bit
		cs_.e.m.init(h.qp, h.type == SLICE_I ? 0 : 1 + h.cabac_init_idc);
		cs_.e.start(&br_);
		last_dqp_ = 0;
		while (mb_addr_ < total) {
			bool skipped = false;
			if (h.type != SLICE_I) {
				ctx_.open_mb(mb_addr_ % mb_w_, mb_addr_ / mb_w_, slice_num_);
				skipped = cs_.mb_skip(h.type == SLICE_B, ctx_.cabac_inc_skip(), 0) != 0;
			}
			if (skipped) { skip_mb(pps); last_dqp_ = 0; }
			else macroblock(pps);
			mb_addr_++;
			if (cs_.e.overrun()) throw BitError();
			if (cs_.end_of_slice(0)) break;
		}
	}

	// ---- syntax elements, either entropy coder -------------------------------------------------------
	int rd_mb_type() {
		if (!cabac_) return (int)br_.ue();
		if (hdr_.type == SLICE_I) return cs_.mb_type_intra(3, true, ctx_.cabac_inc_mbtype_i(), 0);
		if (hdr_.type == SLICE_P) return cs_.mb_type_p(0);
		return cs_.mb_type_b(ctx_.cabac_inc_mbtype_b(), 0);
	}
	int rd_sub_mb_type() {
		if (!cabac_) return (int)br_.ue();
		return bslice_ ? cs_.sub_mb_type_b(0) : cs_.sub_mb_type_p(0);
	}
	// reference index of the partition at (x4, y4), pw x ph 4-pel units
	int rd_ref(int l, int x4, int y4, int pw, int ph, int nref) {
		int r = 0;
		if (nref > 1) {
			if (cabac_) r = cs_.ref_idx(ctx_.cabac_inc_ref(l, x4, y4), 0);
			else r = nref == 2 ? (int)(br_.get1() ^ 1) : (int)br_.ue();
			if (r >= nref) throw BitError();
		}
		if (cabac_) ctx_.set_refgt0(l, x4, y4, pw, ph, r);
		return r;
	}
	void rd_mvd(int l, int x4, int y4, int pw, int ph, int d[2]) {
		if (!cabac_) { d[0] = br_.se(); d[1] = br_.se(); return; }
		d[0] = cs_.mvd(40, ctx_.cabac_mvd_sum(l, x4, y4, 0), 0);
		d[1] = cs_.mvd(47, ctx_.cabac_mvd_sum(l, x4, y4, 1), 0);
		ctx_.set_mvd(l, x4, y4, pw, ph, d[0], d[1]);
	}
	int rd_cbp(bool intra) {
		if (cabac_) return cs_.cbp(ctx_, 0);
		unsigned code = br_.ue();
		if (code > 47) throw BitError();
		return cbp_from_code[intra ? 0 : 1][code];
	}
	int rd_chroma_mode() {
		int v = cabac_ ? cs_.chroma_mode(ctx_.cabac_inc_chroma_mode(), 0) : (int)br_.ue();
		if (v > 3) throw BitError();
		return v;
	}
	int rd_ipred(int pred) {
		if (cabac_) return cs_.ipred_mode(pred, 0);
		if (br_.get1()) return pred;
		int rem = (int)br_.get(3);
		return rem < pred ? rem : rem + 1;
	}
	bool rd_t8x8() { return cs_.transform8x8(ctx_.cabac_inc_t8x8(), 0) != 0; }

	void open_mb(int mbx, int mby) {
		cur_mbx_ = mbx;         // (kept for emit_mc: no division per motion record)
		cur_mby_ = mby;
		ctx_.open_mb(mbx, mby, slice_num_);
		// a macroblock address comes once per picture; a (damaged) slice that runs into macroblocks an earlier
		// slice already delivered would leave their motion records and work items behind
		uint8_t &seen = mb_seen_[(size_t)mby * mb_w_ + mbx];
		if (seen) throw BitError();
		seen = 1;
		b2_h264_mb &m = mbs_[(size_t)mby * mb_w_ + mbx];
		memset(&m, 0, sizeof(m));
		m.coef_idx = (uint32_t)(coef_used_ / 16);
		m.mc_idx = (uint32_t)(dev_mv_ ? parts_.size() : mcs_.size());
	}

	int ref_frame_of(int lx, int refidx) const {
		if (refidx < 0 || refidx >= (int)list_[lx].size()) return -1;
		return list_[lx][refidx].frame_idx;
	}

	// one mb->inter_pred call: ref[lx] < 0 = list unused
	void emit_mc(b2_h264_mb &m, int x4, int y4, int pw, int ph, const int ref[2], const int mv[2][2]) {
		// default prediction: weights 1, offsets 0, both lists unused
		static const b2_h264_mc proto = {0, 0, {-1, -1}, {-1, -1}, 0, 0, {{0, 0}, {0, 0}}, {{1, 1, 1}, {1, 1, 1}}, {{0, 0, 0}, {0, 0, 0}}, {0, 0, 0, 0}};
		b2_h264_mc r = proto;
		r.pos = (uint8_t)((y4 << 2) | x4);
		r.size = (uint8_t)((ph << 4) | pw);
		for (int lx = 0; lx < 2; ++lx) {
			r.ref_idx[lx] = (int8_t)ref[lx];
			if (ref[lx] < 0) continue;
			int f = ref_frame_of(lx, ref[lx]);
			r.ref_frame[lx] = (int8_t)(f < 0 ? cur_frame_ : f);
			r.mv[lx][0] = (int16_t)mv[lx][0];
			r.mv[lx][1] = (int16_t)mv[lx][1];
		}
		const bool bi = ref[0] >= 0 && ref[1] >= 0;
		if (wt_.mode == 1) {
			r.log2wd_y = (uint8_t)wt_.denom[0];
			r.log2wd_c = (uint8_t)wt_.denom[1];
			for (int lx = 0; lx < 2; ++lx)
				if (ref[lx] >= 0)
					for (int c = 0; c < 3; ++c) { r.w[lx][c] = wt_.w[lx][ref[lx] & 31][c]; r.o[lx][c] = wt_.o[lx][ref[lx] & 31][c]; }
		} else if (wt_.mode == 2 && bi) {
			// implicit weights (8.4.2.3.1; pred_weight_type2, h264.cpp:7001-7025)
			const RefPic &r0 = list_[0][ref[0]], &r1 = list_[1][ref[1]];
			int w0 = 32, w1 = 32;
			if (r0.poc != r1.poc && r0.long_idx < 0 && r1.long_idx < 0 && r0.frame_idx >= 0 && r1.frame_idx >= 0) {
				int t = dist_scale_factor(r0.poc, r1.poc, cur_poc_) >> 2;
				if (t >= -64 && t <= 128) { w1 = t; w0 = 64 - t; }
			}
			r.log2wd_y = r.log2wd_c = 5;
			for (int c = 0; c < 3; ++c) { r.w[0][c] = (int8_t)w0; r.w[1][c] = (int8_t)w1; }
		}
		// work items of the inter kernel (include/b2_recon.h): strips of this record, bucketed by class and height.
		// A bi-predicted strip has two: the first-list pass (class of list 0; a launch of its own) and the
		// finishing pass (class of list 1); a strip with one list only the latter.
		{
			auto frac_class = [](int fx, int fy) { return (!fx && !fy) ? 0 : (!fy ? 1 : (!fx ? 2 : (fx == 2 ? 4 : (fy == 2 ? 5 : 3)))); };
			const int first = ref[0] >= 0 ? 0 : 1, last = bi ? 1 : first;
			const int fc_first = frac_class(r.mv[first][0] & 3, r.mv[first][1] & 3), fc_last = frac_class(r.mv[last][0] & 3, r.mv[last][1] & 3);
			const bool centre = fc_first >= 4 || fc_last >= 4;      // the centre sample's filter state does not fit 8-wide strips
			const int wide = pw >= 2 && !centre;
			const int n = wide ? pw >> 1 : pw;
			// strips whose reference window (6-tap margins included) leaves the picture take the clamped loads: they go
			// to buckets of their own (the last 36), so that a warp of the kernel is all fast path or all clamped
			const int W = mb_w_ * 16, H = mb_h_ * 16, sw = wide ? 8 : 4;
			const int mbx = cur_mbx_, mby = cur_mby_;
			const int hidx = ph == 4 ? 2 : ph - 1;
			for (int k = 0; k < n; ++k) {
				const int px = mbx * 16 + x4 * 4 + k * sw, py = mby * 16 + y4 * 4;
				auto edge = [&](int lx) {
					const int x0 = px + (r.mv[lx][0] >> 2), y0 = py + (r.mv[lx][1] >> 2);
					return x0 - 6 < 0 || x0 + sw + 12 > W || y0 - 3 < 0 || y0 + ph * 4 + 4 > H;
				};
				const uint32_t word = (uint32_t)mb_addr_ | ((uint32_t)m.mc_cnt << 20) | ((uint32_t)k << 24);
				const int cls_last = fc_last + 6 * wide;
				items_[(edge(last) ? 36 : 0) + cls_last * 3 + hidx].push_back(word | ((uint32_t)cls_last << 26));
				if (bi) {
					const int cls_first = fc_first + 6 * wide;
					items_first_[(edge(0) ? 36 : 0) + cls_first * 3 + hidx].push_back(word | ((uint32_t)cls_first << 26));
				}
			}
		}
		mcs_.push_back(r);
		m.mc_cnt++;
	}
	void emit_mc_p(b2_h264_mb &m, int x4, int y4, int pw, int ph, int refidx, int mvx, int mvy) {
		const int ref[2] = {refidx, -1};
		const int mv[2][2] = {{mvx, mvy}, {0, 0}};
		emit_mc(m, x4, y4, pw, ph, ref, mv);
	}

	void skip_mb(const Pps &pps) {
		int mbx = mb_addr_ % mb_w_, mby = mb_addr_ / mb_w_;
		open_mb(mbx, mby);
		b2_h264_mb &m = mbs_[mb_addr_];
		m.type = B2_H4_INTER;
		set_mb_qp(m, pps);
		if (hdr_.type == SLICE_B) {
			DirectOut d;
			set_shape(mbx, mby, derive_direct(mbx, mby, d, true), false);
			apply_direct(m, mbx, mby, d, 0xf);
		} else {
			int16_t mv[2];
			set_shape(mbx, mby, SHAPE_16x16, false);
			if (dev_mv_) {
				const int f = ref_frame_of(0, 0);
				m.flags |= (uint8_t)(B2_H4_PSKIP | (ctx_.mb_avail_mask(mbx, mby, slice_num_) << B2_H4_NB_SHIFT));
				parts_.push_back(b2_h264_part{0, 0x44, 0, (int8_t)(f < 0 ? cur_frame_ : f), {0, 0}});
				m.mc_cnt++;
				item_bound_ += 4;
			} else {
			if (keep_parts_) {
				const int f = ref_frame_of(0, 0);
				m.flags |= (uint8_t)(B2_H4_PSKIP | (ctx_.mb_avail_mask(mbx, mby, slice_num_) << B2_H4_NB_SHIFT));
				parts_.push_back(b2_h264_part{0, 0x44, 0, (int8_t)(f < 0 ? cur_frame_ : f), {0, 0}});
			}
			ctx_.pskip_mv(mbx, mby, slice_num_, mv);
			ctx_.set_motion(0, mbx, mby, 0, 0, 4, 4, 0, ref_frame_of(0, 0), mv[0], mv[1]);
			emit_mc_p(m, 0, 0, 4, 4, 0, mv[0], mv[1]);
			}
		}
		MbInfo &info = ctx_.info(mbx, mby);
		info.kind = MB_SKIP;
		info.qp_y = (uint8_t)qp_;
		info.cbp = 0;
		if (hdr_.type == SLICE_B) info.cflags |= MBF_DIRECT16;
		memset(ctx_.nnz_of(mbx, mby), 0, 24);
		finish_mb(mbx, mby, pps, qp_);
	}

	// ---- direct prediction (B_Skip, B_Direct_16x16, direct 8x8 sub-macroblocks) ------------------
	// whole: B_Skip / B_Direct_16x16 (as opposed to a direct sub-macroblock of B_8x8).  Returns
	// the partition shape the reference then treats the macroblock as (its deblocking-strength
	// derivation runs on the co-located macroblock's shape, h264.cpp:9426, 10163).
	int derive_direct(int mbx, int mby, DirectOut &d, bool whole) {
		const RefPic &c1 = list_[1].empty() ? RefPic{-1, -1, -1, 0} : list_[1][0];
		ColPic *col = c1.frame_idx >= 0 ? &col_[c1.frame_idx] : nullptr;
		const size_t mbi = (size_t)mby * mb_w_ + mbx;
		const bool have_shape = col && col->shape.size() > mbi;
		int shape = have_shape ? col->shape[mbi] : SHAPE_16x16;
		if (!hdr_.direct_spatial) {
			ctx_.direct_temporal(mbx, mby, col, direct_8x8_inference_, list_[0].data(), (int)list_[0].size(), list_[1].data(), cur_poc_, d);
			return shape;
		}
		if (!whole) {
			ctx_.direct_spatial(mbx, mby, slice_num_, col, c1.long_idx < 0, direct_8x8_inference_, d);
			return shape;
		}
		// The reference hands the shape to its bookkeeping by overwriting the co-located
		// picture's record whenever the co-located motion is not consulted (h264.cpp:9786, 9968,
		// 9975); later pictures that share the co-located picture see the overwritten value.
		int refs[2];
		DirectOut probe;
		ctx_.direct_spatial(mbx, mby, slice_num_, col, c1.long_idx < 0, direct_8x8_inference_, probe, -1, refs);
		if ((refs[0] < 0 && refs[1] < 0) || c1.long_idx >= 0 || (refs[0] != 0 && refs[1] != 0)) {
			shape = SHAPE_16x16;
			if (have_shape) col->shape[mbi] = SHAPE_16x16;
		}
		ctx_.direct_spatial(mbx, mby, slice_num_, col, c1.long_idx < 0, direct_8x8_inference_, d, shape == SHAPE_8x8 ? -1 : shape);
		return shape;
	}

	// store the motion of the direct-predicted 8x8 quadrants in `quads` and emit their MC records
	void apply_direct(b2_h264_mb &m, int mbx, int mby, const DirectOut &d, int quads) {
		// one record for the whole MB when every block got the same motion (the common case)
		if (quads == 0xf) {
			bool same = true;
			for (int i = 1; i < 16 && same; ++i)
				same = memcmp(d.ref[i], d.ref[0], sizeof(d.ref[0])) == 0 && memcmp(d.mv[i], d.mv[0], sizeof(d.mv[0])) == 0;
			if (same) {
				for (int l = 0; l < 2; ++l)
					ctx_.set_motion(l, mbx, mby, 0, 0, 4, 4, d.ref[0][l], ref_frame_of(l, d.ref[0][l]), d.mv[0][l][0], d.mv[0][l][1]);
				emit_mc(m, 0, 0, 4, 4, d.ref[0], d.mv[0]);
				return;
			}
		}
		for (int q = 0; q < 4; ++q) {
			if (!((quads >> q) & 1)) continue;
			const int bx = (q & 1) * 2, by = (q >> 1) * 2;
			const int step = direct_8x8_inference_ ? 2 : 1;
			for (int y = by; y < by + 2; y += step)
				for (int x = bx; x < bx + 2; x += step) {
					const int i = y * 4 + x;
					for (int l = 0; l < 2; ++l)
						ctx_.set_motion(l, mbx, mby, x, y, step, step, d.ref[i][l], ref_frame_of(l, d.ref[i][l]), d.mv[i][l][0], d.mv[i][l][1]);
					emit_mc(m, x, y, step, step, d.ref[i], d.mv[i]);
				}
		}
	}

	// partition shape of the current MB (kept for co-located use) and the class of its
	// deblocking-strength derivation in the reference: coded8x8 = P_8x8 / B_8x8
	// (store_info_intermb8x8) as opposed to a direct-predicted MB on an 8x8 co-located shape
	void set_shape(int mbx, int mby, int shape, bool coded8x8) {
		MbInfo &info = ctx_.info(mbx, mby);
		info.cflags = (uint8_t)((info.cflags & ~(3 << MBF_SHAPE_SHIFT)) | (shape << MBF_SHAPE_SHIFT));
		// bit 0: MB edges, bit 1: centre edges, bit 2: the other inner edges use the reduced comparison
		bs_reduced_ = shape == SHAPE_16x16 ? 0 : (shape != SHAPE_8x8 ? 2 : (coded8x8 ? 3 : 6));
	}

	void set_mb_qp(b2_h264_mb &m, const Pps &pps) {
		m.qp_y = (uint8_t)qp_;
		m.qp_cb = (uint8_t)chroma_qp(qp_, pps.chroma_qp_offset[0]);
		m.qp_cr = (uint8_t)chroma_qp(qp_, pps.chroma_qp_offset[1]);
	}

	// coefficient pool: grows geometrically, records are handed out zeroed
	int16_t *new_record() {
		if (coef_used_ + 16 > coefs_.size()) coefs_.resize(coefs_.size() * 2 + 65536);
		int16_t *r = &coefs_[coef_used_];
		coef_used_ += 16;
		memset(r, 0, 32);
		return r;
	}

	// residual_block_cavlc (7.3.5.3.2 / 9.2): returns total_coeff; levels go to out[scan[...]]
	// out() is asked for the destination only once the block turns out to have coefficients, so that
	// empty blocks (most of them) cost neither a record nor its clearing
	template <class Out>
	inline __attribute__((always_inline)) int residual_block(int nc, int max_coeff, Out &&out_of, const uint8_t *scan, int scan_ofs) {
		const H264Tables &t = *tab_;
		// one 32-bit look serves coeff_token (at most 16 bits) and the trailing-ones signs (at most 3)
		uint32_t look = br_.peek(32);
		if (nc >= 0 && nc < 2 && (look >> 31)) { br_.skip(1); return 0; }      // "1": no coefficients (the most frequent code)
		const VlcSmall &ct = nc < 0 ? t.coeff_token[4] : (nc < 2 ? t.coeff_token[0] : (nc < 4 ? t.coeff_token[1] : (nc < 8 ? t.coeff_token[2] : t.coeff_token[3])));
		int len;
		const int tok = ct.lookup(look >> 16, &len);
		if (tok < 0) throw BitError();
		const int total = tok >> 2, t1 = tok & 3;
		if (total == 0) { br_.skip(len); return 0; }
		if (total > max_coeff) throw BitError();
		int16_t *out = out_of();
		int level[16];
		// trailing-ones signs: all three candidate bits (the loop's trip count is a coin the predictor loses)
		level[0] = 1 - 2 * (int)((look >> (31 - len)) & 1);
		level[1] = 1 - 2 * (int)((look >> (30 - len)) & 1);
		level[2] = 1 - 2 * (int)((look >> (29 - len)) & 1);
		br_.skip(len + t1);
		int suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
		for (int i = t1; i < total; ++i) {
			// level_prefix: count leading zeros, then the terminating one; the suffix comes from the same look
			look = br_.peek(32);
			if (!look) throw BitError();
			const int prefix = __builtin_clz(look);
			int code = (prefix < 15 ? prefix : 15) << suffix_len;
			int size = suffix_len;
			if (prefix == 14 && suffix_len == 0) size = 4;
			else if (prefix >= 15) size = prefix - 3;
			if (prefix + 1 + size <= 32) {
				if (size) code += (int)((look << (prefix + 1)) >> (32 - size));
				br_.skip(prefix + 1 + size);
			} else {
				br_.skip(prefix + 1);
				code += (int)br_.get(size);
			}
			if (prefix >= 15 && suffix_len == 0) code += 15;
			if (prefix >= 16) code += (1 << (prefix - 3)) - 4096;
			if (i == t1 && t1 < 3) code += 2;
			int lv = (code & 1) ? -((code + 1) >> 1) : ((code + 2) >> 1);
			level[i] = lv;
			if (suffix_len == 0) suffix_len = 1;
			int a = lv < 0 ? -lv : lv;
			if (suffix_len < 6 && a > (3 << (suffix_len - 1))) suffix_len++;
		}
		// total_zeros (at most 9 bits) and the run_before codes (at most 11 bits each) from shared looks
		look = br_.peek(32);
		int used = 0;
		int zeros_left = 0;
		if (total < max_coeff) {
			const VlcSmall &tz = max_coeff == 4 ? t.total_zeros_cdc[total - 1] : t.total_zeros[total - 1];
			zeros_left = tz.lookup(look >> 16, &len);
			if (zeros_left < 0) throw BitError();
			used = len;
		}
		// levels come highest frequency first
		int pos = total + zeros_left - 1;
		if (pos >= max_coeff) throw BitError();
		scan += scan_ofs;
		for (int i = 0; i < total; ++i) {
			int lv = level[i];
			if (lv > 32767) lv = 32767;
			if (lv < -32768) lv = -32768;
			out[scan[pos]] = (int16_t)lv;
			if (i < total - 1 && zeros_left > 0) {
				if (used > 16) { br_.skip(used); look = br_.peek(32); used = 0; }
				const int rb = t.run_before[(zeros_left > 7 ? 7 : zeros_left) - 1].lookup((look << used) >> 16, &len);
				if (rb < 0 || rb > zeros_left) throw BitError();
				used += len;
				zeros_left -= rb;
				pos -= rb;
			}
			pos--;
		}
		br_.skip(used);
		return total;
	}

	// luma residual of one MB for the 8x8 blocks selected by cbp; i16: AC-only blocks (cat 1);
	// t8x8: transform_size_8x8_flag (CABAC only: one 64-coefficient block per 8x8, cat 5)
	void residual_luma(b2_h264_mb &m, int mbx, int mby, int cbp, bool i16, bool intra, bool t8x8) {
		uint8_t *nnz = ctx_.nnz_of(mbx, mby);
		if (t8x8) {
			for (int q = 0; q < 4; ++q) {
				const int x4 = (q & 1) * 2, y4 = (q >> 1) * 2;
				int total = 0;
				if (cbp & (1 << q)) {
					const size_t at = coef_used_;
					for (int r = 0; r < 4; ++r) new_record();
					total = cs_.residual(5, -1, &coefs_[at], 64, zigzag8x8);      // 64 levels, raster 8x8
					m.nz |= 0xfu << (1 + 4 * q);
					if (total > 15) total = 15;
				}
				nnz[y4 * 4 + x4] = nnz[y4 * 4 + x4 + 1] = nnz[(y4 + 1) * 4 + x4] = nnz[(y4 + 1) * 4 + x4 + 1] = (uint8_t)total;
			}
			return;
		}
		if (!cabac_) {
			// 9.2.1 without branches: the counts of the macroblock, of the row above it and of the column to its left in one
			// 5 x 8 grid, 64 standing for "not available": nA + nB < 64 -> both there (rounded mean), else the one that is
			// there (or 0) is what is left after masking the 64s away.  (The same for the eight chroma blocks measured 2 %
			// slower, not faster; they keep cavlc_nc.)
			uint8_t g[5 * 8];
			const uint8_t *nl = ctx_.env.nnz_left, *nt = ctx_.env.nnz_top;
			for (int i = 0; i < 4; ++i) {
				g[1 + i] = nt ? nt[12 + i] : 64;
				g[8 * (1 + i)] = nl ? nl[4 * i + 3] : 64;
			}
			auto rec = [this]() { return new_record(); };
			for (int k = 0; k < 16; ++k) {
				const int x4 = blk_x[k], y4 = blk_y[k], gi = 8 * (1 + y4) + 1 + x4;
				int total = 0;
				if (cbp & (1 << (k >> 2))) {
					const int sum = g[gi - 1] + g[gi - 8];
					const int nc = sum < 64 ? (sum + 1) >> 1 : (sum & 63);
					total = i16 ? residual_block(nc, 15, rec, zigzag4x4, 1) : residual_block(nc, 16, rec, zigzag4x4, 0);
					if (total) m.nz |= 1u << (1 + k);
				}
				g[gi] = (uint8_t)total;
				nnz[y4 * 4 + x4] = (uint8_t)total;
			}
			return;
		}
		for (int k = 0; k < 16; ++k) {
			int x4 = blk_x[k], y4 = blk_y[k];
			if (!(cbp & (1 << (k >> 2)))) { nnz[y4 * 4 + x4] = 0; continue; }
			int total;
			if (cabac_) {
				// (the levels go straight into a record, which is handed back when the block turns out empty)
				total = cs_.residual(i16 ? 1 : 2, ctx_.cabac_inc_cbf_luma(x4, y4, intra), new_record(), i16 ? 15 : 16, zigzag4x4 + (i16 ? 1 : 0));
				if (!total) coef_used_ -= 16;
			} else {
				int nc = ctx_.cavlc_nc(mbx, mby, 0, x4, y4, slice_num_);
				auto rec = [this]() { return new_record(); };
				total = i16 ? residual_block(nc, 15, rec, zigzag4x4, 1) : residual_block(nc, 16, rec, zigzag4x4, 0);
			}
			nnz[y4 * 4 + x4] = (uint8_t)total;
			if (total) m.nz |= 1u << (1 + k);
		}
	}

	void residual_chroma(b2_h264_mb &m, int mbx, int mby, int cbp_chroma, bool intra) {
		uint8_t *nnz = ctx_.nnz_of(mbx, mby);
		MbInfo &info = ctx_.info(mbx, mby);
		memset(nnz + 16, 0, 8);
		if (!cbp_chroma) return;
		static const uint8_t scan_dc[4] = {0, 1, 2, 3};
		int16_t dc[16];
		memset(dc, 0, sizeof(dc));
		int any = 0;
		for (int c = 0; c < 2; ++c) {
			int n;
			if (cabac_) {
				n = cs_.residual(3, ctx_.cabac_inc_cbf_dc(MBF_CBF_CB_DC << c, intra), dc + c * 4, 4);
				if (n) info.cflags |= (uint8_t)(MBF_CBF_CB_DC << c);
			} else {
				int16_t *dst = dc + c * 4;
				n = residual_block(-1, 4, [dst]() { return dst; }, scan_dc, 0);
			}
			any |= n;
		}
		if (any) {
			memcpy(new_record(), dc, sizeof(dc));
			m.nz |= 1u << 17;
		}
		if (cbp_chroma & 2) {
			for (int c = 0; c < 2; ++c)
				for (int b = 0; b < 4; ++b) {
					int x = b & 1, y = b >> 1;
					int total;
					if (cabac_) {
						total = cs_.residual(4, ctx_.cabac_inc_cbf_chroma_ac(c, x, y, intra), new_record(), 15, zigzag4x4 + 1);
						if (!total) coef_used_ -= 16;
					} else {
						int nc = ctx_.cavlc_nc(mbx, mby, 1 + c, x, y, slice_num_);
						total = residual_block(nc, 15, [this]() { return new_record(); }, zigzag4x4, 1);
					}
					nnz[16 + c * 4 + b] = (uint8_t)total;
					if (total) m.nz |= 1u << (18 + c * 4 + b);
				}
		}
	}

	void read_qp_delta() {
		int d = cabac_ ? cs_.qp_delta(last_dqp_ != 0, 0) : br_.se();
		if (d < -26 || d > 25) throw BitError();   // 7.4.5: outside this range the QP leaves 0..51 (a table index on the device)
		last_dqp_ = d;
		qp_ += d;
		if (qp_ < 0) qp_ += 52;
		else if (qp_ >= 52) qp_ -= 52;
	}

	void macroblock(const Pps &pps) {
		const SliceHeader &h = hdr_;
		int mbx = mb_addr_ % mb_w_, mby = mb_addr_ / mb_w_;
		open_mb(mbx, mby);
		b2_h264_mb &m = mbs_[mb_addr_];
		MbInfo &info = ctx_.info(mbx, mby);
		int mb_type = rd_mb_type();
		if (h.type == SLICE_P) {
			if (mb_type >= 5) mb_type -= 5;          // intra types follow the five P types
			else { inter_mb(m, mbx, mby, mb_type, pps); return; }
		} else if (h.type == SLICE_B) {
			if (mb_type >= 23) mb_type -= 23;        // intra types follow the 23 B types
			else { b_mb(m, mbx, mby, mb_type, pps); return; }
		}
		if (mb_type > 25) throw BitError();
		int avail = ctx_.mb_avail_mask(mbx, mby, slice_num_);
		int avail_intra = avail;
		if (pps.constrained_intra) {
			// neighbours coded in inter mode do not feed intra prediction (8.3.1.2 etc.); the
			// reference masks left / top / top-right only, never top-left (h264.cpp:3271-3273)
			static const int dx[3] = {-1, 0, 1}, dy[3] = {0, -1, -1};
			for (int b = 0; b < 3; ++b)
				if ((avail >> b) & 1) {
					const MbInfo &o = ctx_.info(mbx + dx[b], mby + dy[b]);
					if (!kind_is_intra(o.kind)) avail_intra &= ~(1 << b);
				}
		}
		if (ctx_.keep_motion) ctx_.set_intra_motion(mbx, mby);
		n_intra_++;
		uint8_t *nnz = ctx_.nnz_of(mbx, mby);
		int8_t *ip = ctx_.ipred_of(mbx, mby);
		if (mb_type == 25) {   // I_PCM (h264.cpp:4736-4761)
			if (cabac_) cs_.e.before_pcm();
			else br_.byte_align();
			m.type = B2_H4_IPCM;
			m.nz = 0xfff;
			for (int r = 0; r < 12; ++r) {
				uint8_t *dst = (uint8_t *)new_record();
				for (int i = 0; i < 32; ++i) dst[i] = (uint8_t)br_.get(8);
			}
			if (cabac_) cs_.e.start(&br_);
			set_mb_qp(m, pps);
			memset(nnz, 16, 24);
			memset(ip, 2, 16);
			info.kind = MB_IPCM;
			info.qp_y = (uint8_t)qp_;
			info.cbp = 0x2f;
			info.cflags = MBF_CBF_LUMA_DC | MBF_CBF_CB_DC | MBF_CBF_CR_DC;
			last_dqp_ = 0;
			finish_mb(mbx, mby, pps, 0);
			return;
		}
		int cbp;
		if (mb_type == 0) {   // I_NxN
			const bool t8x8 = pps.transform_8x8 && rd_t8x8();
			if (t8x8) {
				m.type = B2_H4_I8x8;
				m.flags |= B2_H4_T8x8;
				info.kind = MB_I8x8;
				info.t8x8 = 1;
				for (int q = 0; q < 4; ++q) {
					const int x4 = (q & 1) * 2, y4 = (q >> 1) * 2;
					int pred = ctx_.predict_ipred(mbx, mby, x4, y4, slice_num_, pps.constrained_intra != 0);
					int mode = rd_ipred(pred);
					ip[y4 * 4 + x4] = ip[y4 * 4 + x4 + 1] = ip[(y4 + 1) * 4 + x4] = ip[(y4 + 1) * 4 + x4 + 1] = (int8_t)mode;
					m.i4_modes[q >> 1] |= (uint8_t)(mode << ((q & 1) * 4));
				}
			} else {
				m.type = B2_H4_I4x4;
				info.kind = MB_I4x4;
				for (int k = 0; k < 16; ++k) {
					int x4 = blk_x[k], y4 = blk_y[k];
					int pred = ctx_.predict_ipred(mbx, mby, x4, y4, slice_num_, pps.constrained_intra != 0);
					int mode = rd_ipred(pred);
					ip[y4 * 4 + x4] = (int8_t)mode;
					m.i4_modes[k >> 1] |= (uint8_t)(mode << ((k & 1) * 4));
				}
			}
			m.chroma_mode = (uint8_t)rd_chroma_mode();
			info.chroma_mode = m.chroma_mode;
			cbp = rd_cbp(true);
			// The reference feeds the no-residual Intra4x4 path the unmasked availability
			// (hazard 8, h264.cpp:3292); carried here so the device needs no special case.
			// Intra8x8 always gets the masked value (h264.cpp:4150-4163).
			m.intra_avail = (uint8_t)((t8x8 || (cbp & 15)) ? avail_intra : avail);
			m.intra_avail_c = (uint8_t)avail_intra;
			if (cbp) read_qp_delta();
			else last_dqp_ = 0;
			set_mb_qp(m, pps);
			residual_luma(m, mbx, mby, cbp & 15, false, true, t8x8);
		} else {               // Intra16x16
			m.type = B2_H4_I16x16;
			info.kind = MB_I16x16;
			memset(ip, 2, 16);
			int t = mb_type - 1;
			m.i16_mode = (uint8_t)(t & 3);
			int cbp_chroma = (t >> 2) % 3;
			int cbp_luma = t >= 12 ? 15 : 0;
			cbp = cbp_luma | (cbp_chroma << 4);
			m.intra_avail = m.intra_avail_c = (uint8_t)avail_intra;
			m.chroma_mode = (uint8_t)rd_chroma_mode();
			info.chroma_mode = m.chroma_mode;
			read_qp_delta();
			set_mb_qp(m, pps);
			int16_t dc[16];
			memset(dc, 0, sizeof(dc));
			int ndc;
			if (cabac_) {
				int16_t lv[16];
				memset(lv, 0, sizeof(lv));
				ndc = cs_.residual(0, ctx_.cabac_inc_cbf_dc(MBF_CBF_LUMA_DC, true), lv, 16);
				if (ndc) {
					info.cflags |= MBF_CBF_LUMA_DC;
					for (int i = 0; i < 16; ++i) dc[zigzag4x4[i]] = lv[i];
				}
			} else {
				int nc = ctx_.cavlc_nc(mbx, mby, 0, 0, 0, slice_num_);
				ndc = residual_block(nc, 16, [&dc]() { return dc; }, zigzag4x4, 0);
			}
			if (ndc) {
				memcpy(new_record(), dc, sizeof(dc));
				m.nz |= 1;
			}
			residual_luma(m, mbx, mby, cbp_luma, true, true, false);
		}
		residual_chroma(m, mbx, mby, cbp >> 4, true);
		info.qp_y = (uint8_t)qp_;
		info.cbp = (uint8_t)cbp;
		finish_mb(mbx, mby, pps, qp_);
	}

	// coded_block_pattern, transform_size_8x8_flag, mb_qp_delta and residual of an inter MB.
	// may_t8x8: no sub-macroblock partition smaller than 8x8 and, for direct prediction,
	// direct_8x8_inference_flag (7.3.5)
	void inter_residual(b2_h264_mb &m, int mbx, int mby, const Pps &pps, bool may_t8x8) {
		MbInfo &info = ctx_.info(mbx, mby);
		int cbp = rd_cbp(false);
		info.cbp = (uint8_t)cbp;      // visible to the chroma bins' own context? no: neighbours only
		bool t8x8 = false;
		if ((cbp & 15) && pps.transform_8x8 && may_t8x8) t8x8 = rd_t8x8();
		if (t8x8) { m.flags |= B2_H4_T8x8; info.t8x8 = 1; }
		if (cbp) read_qp_delta();
		else last_dqp_ = 0;
		set_mb_qp(m, pps);
		residual_luma(m, mbx, mby, cbp & 15, false, false, t8x8);
		residual_chroma(m, mbx, mby, cbp >> 4, false);
		info.qp_y = (uint8_t)qp_;
		finish_mb(mbx, mby, pps, qp_);
	}

	void inter_mb(b2_h264_mb &m, int mbx, int mby, int mb_type, const Pps &pps) {
		const SliceHeader &h = hdr_;
		MbInfo &info = ctx_.info(mbx, mby);
		m.type = B2_H4_INTER;
		info.kind = MB_INTER;
		memset(ctx_.ipred_of(mbx, mby), 2, 16);
		const int nref = h.num_ref_idx_active[0];
		if (dev_mv_ || keep_parts_) m.flags |= (uint8_t)(ctx_.mb_avail_mask(mbx, mby, slice_num_) << B2_H4_NB_SHIFT);
		auto do_part = [&](int x4, int y4, int pw, int ph, int ref, int shape, int part) {
			int16_t pred[2];
			int d[2];
			if (dev_mv_) {
				// the device predicts: hand over the difference as read
				rd_mvd(0, x4, y4, pw, ph, d);
				const int f = ref_frame_of(0, ref);
				parts_.push_back(b2_h264_part{(uint8_t)((y4 << 2) | x4), (uint8_t)((ph << 4) | pw), (int8_t)ref, (int8_t)(f < 0 ? cur_frame_ : f), {(int16_t)d[0], (int16_t)d[1]}});
				m.mc_cnt++;
				item_bound_ += (uint32_t)pw;
				return;
			}
			ctx_.predict_mv(0, mbx, mby, x4, y4, pw, ref, slice_num_, shape, part, pred);
			rd_mvd(0, x4, y4, pw, ph, d);
			if (keep_parts_) {           // (a picture tap sees both the partitions and the records made of them)
				const int f = ref_frame_of(0, ref);
				parts_.push_back(b2_h264_part{(uint8_t)((y4 << 2) | x4), (uint8_t)((ph << 4) | pw), (int8_t)ref, (int8_t)(f < 0 ? cur_frame_ : f), {(int16_t)d[0], (int16_t)d[1]}});
			}
			int mvx = pred[0] + d[0];
			int mvy = pred[1] + d[1];
			ctx_.set_motion(0, mbx, mby, x4, y4, pw, ph, ref, ref_frame_of(0, ref), mvx, mvy);
			emit_mc_p(m, x4, y4, pw, ph, ref, mvx, mvy);
		};
		bool may_t8x8 = true;
		set_shape(mbx, mby, mb_type > 3 ? SHAPE_8x8 : mb_type, true);
		if (mb_type == 0) {
			int r = rd_ref(0, 0, 0, 4, 4, nref);
			do_part(0, 0, 4, 4, r, 0, 0);
		} else if (mb_type == 1) {
			int r0 = rd_ref(0, 0, 0, 4, 2, nref), r1 = rd_ref(0, 0, 2, 4, 2, nref);
			do_part(0, 0, 4, 2, r0, 1, 0);
			do_part(0, 2, 4, 2, r1, 1, 1);
		} else if (mb_type == 2) {
			int r0 = rd_ref(0, 0, 0, 2, 4, nref), r1 = rd_ref(0, 2, 0, 2, 4, nref);
			do_part(0, 0, 2, 4, r0, 2, 0);
			do_part(2, 0, 2, 4, r1, 2, 1);
		} else {
			if (mb_type > 4) throw BitError();
			int sub[4], ref[4];
			for (int i = 0; i < 4; ++i) {
				sub[i] = rd_sub_mb_type();
				if (sub[i] > 3) throw BitError();
				if (sub[i]) may_t8x8 = false;
			}
			for (int i = 0; i < 4; ++i) ref[i] = mb_type == 4 ? 0 : rd_ref(0, (i & 1) * 2, (i >> 1) * 2, 2, 2, nref);   // P_8x8ref0
			for (int i = 0; i < 4; ++i) {
				int bx = (i & 1) * 2, by = (i >> 1) * 2;
				switch (sub[i]) {
				case 0: do_part(bx, by, 2, 2, ref[i], 0, 0); break;
				case 1: do_part(bx, by, 2, 1, ref[i], 0, 0); do_part(bx, by + 1, 2, 1, ref[i], 0, 0); break;
				case 2: do_part(bx, by, 1, 2, ref[i], 0, 0); do_part(bx + 1, by, 1, 2, ref[i], 0, 0); break;
				default:
					for (int s = 0; s < 4; ++s) do_part(bx + (s & 1), by + (s >> 1), 1, 1, ref[i], 0, 0);
					break;
				}
			}
		}
		inter_residual(m, mbx, mby, pps, may_t8x8);
	}

	// B macroblock (7.3.5.1, 7.3.5.2; Tables 7-14, 7-18)
	void b_mb(b2_h264_mb &m, int mbx, int mby, int mb_type, const Pps &pps) {
		const SliceHeader &h = hdr_;
		MbInfo &info = ctx_.info(mbx, mby);
		m.type = B2_H4_INTER;
		info.kind = MB_INTER;
		memset(ctx_.ipred_of(mbx, mby), 2, 16);
		struct Part { int8_t x4, y4, pw, ph, shape, part, use, sub; };   // use: 0 direct, 1 L0, 2 L1, 3 Bi
		Part parts[16];
		int np = 0;
		int sub_use[4] = {0, 0, 0, 0};
		DirectOut dir;
		int direct_quads = 0;
		bool may_t8x8 = true;
		if (mb_type > 22) throw BitError();
		if (mb_type == 0) {
			direct_quads = 0xf;
			info.cflags |= MBF_DIRECT16;
			may_t8x8 = direct_8x8_inference_;
		} else if (mb_type < 22) {
			// 1..3: 16x16 L0/L1/Bi; 4..21: pairs (16x8, 8x16) with the prediction modes of Table 7-14
			static const uint8_t modes[22][2] = {{0, 0}, {1, 0}, {2, 0}, {3, 0}, {1, 1}, {1, 1}, {2, 2}, {2, 2}, {1, 2}, {1, 2}, {2, 1}, {2, 1},
			                                     {1, 3}, {1, 3}, {2, 3}, {2, 3}, {3, 1}, {3, 1}, {3, 2}, {3, 2}, {3, 3}, {3, 3}};
			if (mb_type <= 3) parts[np++] = Part{0, 0, 4, 4, 0, 0, (int8_t)modes[mb_type][0], 0};
			else if (mb_type & 1) {   // odd: 8x16
				parts[np++] = Part{0, 0, 2, 4, 2, 0, (int8_t)modes[mb_type][0], 0};
				parts[np++] = Part{2, 0, 2, 4, 2, 1, (int8_t)modes[mb_type][1], 1};
			} else {                  // even: 16x8
				parts[np++] = Part{0, 0, 4, 2, 1, 0, (int8_t)modes[mb_type][0], 0};
				parts[np++] = Part{0, 2, 4, 2, 1, 1, (int8_t)modes[mb_type][1], 1};
			}
			sub_use[0] = parts[0].use;
			if (np > 1) sub_use[1] = parts[1].use;
		} else {
			// B_8x8: sub_mb_type per quadrant (Table 7-18)
			static const uint8_t sub_mode[13] = {0, 1, 2, 3, 1, 1, 2, 2, 3, 3, 1, 2, 3};
			static const uint8_t sub_shape[13] = {0, 0, 0, 0, 1, 2, 1, 2, 1, 2, 3, 3, 3};   // 0 8x8, 1 8x4, 2 4x8, 3 4x4
			for (int q = 0; q < 4; ++q) {
				unsigned st = (unsigned)rd_sub_mb_type();
				if (st > 12) throw BitError();
				int bx = (q & 1) * 2, by = (q >> 1) * 2, use = sub_mode[st];
				sub_use[q] = use;
				if (use == 0) {
					direct_quads |= 1 << q;
					if (!direct_8x8_inference_) may_t8x8 = false;
					parts[np++] = Part{(int8_t)bx, (int8_t)by, 2, 2, 0, 0, 0, (int8_t)q};
					continue;
				}
				if (sub_shape[st]) may_t8x8 = false;
				switch (sub_shape[st]) {
				case 0: parts[np++] = Part{(int8_t)bx, (int8_t)by, 2, 2, 0, 0, (int8_t)use, (int8_t)q}; break;
				case 1: parts[np++] = Part{(int8_t)bx, (int8_t)by, 2, 1, 0, 0, (int8_t)use, (int8_t)q};
				        parts[np++] = Part{(int8_t)bx, (int8_t)(by + 1), 2, 1, 0, 0, (int8_t)use, (int8_t)q}; break;
				case 2: parts[np++] = Part{(int8_t)bx, (int8_t)by, 1, 2, 0, 0, (int8_t)use, (int8_t)q};
				        parts[np++] = Part{(int8_t)(bx + 1), (int8_t)by, 1, 2, 0, 0, (int8_t)use, (int8_t)q}; break;
				default:
					for (int k = 0; k < 4; ++k) parts[np++] = Part{(int8_t)(bx + (k & 1)), (int8_t)(by + (k >> 1)), 1, 1, 0, 0, (int8_t)use, (int8_t)q};
					break;
				}
			}
		}
		if (mb_type == 0) set_shape(mbx, mby, derive_direct(mbx, mby, dir, true), false);
		else {
			if (direct_quads) derive_direct(mbx, mby, dir, false);
			set_shape(mbx, mby, mb_type == 22 ? SHAPE_8x8 : (mb_type <= 3 ? SHAPE_16x16 : ((mb_type & 1) ? SHAPE_8x16 : SHAPE_16x8)), true);
		}
		// reference indices: list 0 for every (sub-)macroblock partition that uses it, then list 1
		int ref[4][2];
		const int nsub = mb_type == 22 ? 4 : (mb_type == 0 ? 0 : (mb_type <= 3 ? 1 : 2));
		for (int q = 0; q < 4; ++q) ref[q][0] = ref[q][1] = -1;
		for (int l = 0; l < 2; ++l) {
			const int nref = h.num_ref_idx_active[l];
			for (int q = 0; q < nsub; ++q) {
				if (!(sub_use[q] & (1 << l))) continue;
				// area the index applies to: the partition (16x16 / 16x8 / 8x16) or the 8x8 quadrant
				int x4, y4, pw, ph;
				if (mb_type == 22) { x4 = (q & 1) * 2; y4 = (q >> 1) * 2; pw = ph = 2; }
				else { x4 = parts[q].x4; y4 = parts[q].y4; pw = parts[q].pw; ph = parts[q].ph; }
				ref[q][l] = rd_ref(l, x4, y4, pw, ph, nref);
			}
		}
		// motion vector differences: list 0 for all partitions in order, then list 1
		int pmv[16][2][2];
		memset(pmv, 0, sizeof(pmv));
		for (int l = 0; l < 2; ++l) {
			for (int i = 0; i < np; ++i) {
				const Part &p = parts[i];
				if (p.use == 0) {
					// direct sub-macroblock: motion was derived, make it visible to later predictions
					const int step = direct_8x8_inference_ ? 2 : 1;
					for (int y = p.y4; y < p.y4 + 2; y += step)
						for (int x = p.x4; x < p.x4 + 2; x += step) {
							const int k = y * 4 + x;
							ctx_.set_motion(l, mbx, mby, x, y, step, step, dir.ref[k][l], ref_frame_of(l, dir.ref[k][l]), dir.mv[k][l][0], dir.mv[k][l][1]);
						}
					continue;
				}
				if (!(p.use & (1 << l))) { ctx_.set_motion(l, mbx, mby, p.x4, p.y4, p.pw, p.ph, -1, -1, 0, 0); continue; }
				int16_t pred[2];
				int d[2];
				const int r = ref[p.sub][l];
				ctx_.predict_mv(l, mbx, mby, p.x4, p.y4, p.pw, r, slice_num_, p.shape, p.part, pred);
				rd_mvd(l, p.x4, p.y4, p.pw, p.ph, d);
				pmv[i][l][0] = pred[0] + d[0];
				pmv[i][l][1] = pred[1] + d[1];
				ctx_.set_motion(l, mbx, mby, p.x4, p.y4, p.pw, p.ph, r, ref_frame_of(l, r), pmv[i][l][0], pmv[i][l][1]);
			}
		}
		if (mb_type == 0) {
			apply_direct(m, mbx, mby, dir, 0xf);
		} else {
			for (int i = 0; i < np; ++i) {
				const Part &p = parts[i];
				if (p.use == 0) {
					// records of a direct quadrant (motion already stored above)
					const int step = direct_8x8_inference_ ? 2 : 1;
					for (int y = p.y4; y < p.y4 + 2; y += step)
						for (int x = p.x4; x < p.x4 + 2; x += step) emit_mc(m, x, y, step, step, dir.ref[y * 4 + x], dir.mv[y * 4 + x]);
					continue;
				}
				const int rr[2] = {(p.use & 1) ? ref[p.sub][0] : -1, (p.use & 2) ? ref[p.sub][1] : -1};
				emit_mc(m, p.x4, p.y4, p.pw, p.ph, rr, pmv[i]);
			}
		}
		inter_residual(m, mbx, mby, pps, may_t8x8);
	}

	// ---- deblocking strengths (8.7.2.1) for the MB just parsed ------------------------------------
	// bS of an inter/inter 4x4 block pair: 2 if either block has coefficients, else 1 if the
	// reference pictures differ or a motion vector component differs by >= 4 quarter samples
	static inline bool mv_far(const int16_t *a, const int16_t *b) {
		return (unsigned)(a[0] - b[0] + 3) > 6u || (unsigned)(a[1] - b[1] + 3) > 6u;
	}
	// reduced: where the reference's derivation for this kind of edge skips the crosswise
	// comparison when all four reference pictures are the same one (str_mv_calc8x8_edge_bidir,
	// str_mv_calc8x8_half_bidir, is_str_mv_calc16x8_center_bidir, str_mv_calc8x8_inner_bidir:
	// h264.cpp:8797, 8926, 7476, 9218) -- a deviation from 8.7.2.1 that decides pixels
	inline int bs_inter(size_t g0, size_t g1, bool nz0, bool nz1, bool reduced = false) const {
		if (nz0 | nz1) return 2;
		const int a0 = ctx_.refpic[0][g0], b0 = ctx_.refpic[0][g1];
		const int16_t *ma0 = &ctx_.mv[0][g0 * 2], *mb0 = &ctx_.mv[0][g1 * 2];
		if (!bslice_) return (a0 != b0 || mv_far(ma0, mb0)) ? 1 : 0;
		// two lists: different pictures or number of vectors -> 1; same pictures -> compare the vectors
		// that point to the same picture (8.7.2.1)
		const int a1 = ctx_.refpic[1][g0], b1 = ctx_.refpic[1][g1];
		const int16_t *ma1 = &ctx_.mv[1][g0 * 2], *mb1 = &ctx_.mv[1][g1 * 2];
		if (!((a0 == b0 && a1 == b1) || (a0 == b1 && a1 == b0))) return 1;
		if (a0 != a1) {
			if (a0 == b0) return ((a0 >= 0 && mv_far(ma0, mb0)) || (a1 >= 0 && mv_far(ma1, mb1))) ? 1 : 0;
			return ((a0 >= 0 && mv_far(ma0, mb1)) || (a1 >= 0 && mv_far(ma1, mb0))) ? 1 : 0;
		}
		bool straight = mv_far(ma0, mb0) || mv_far(ma1, mb1), cross = mv_far(ma0, mb1) || mv_far(ma1, mb0);
		if (reduced) return straight ? 1 : 0;
		return (straight && cross) ? 1 : 0;
	}

	void finish_mb(int mbx, int mby, const Pps &pps, int dbk_qp) {
		const SliceHeader &h = hdr_;
		b2_h264_dbk &d = dbk_[(size_t)mby * mb_w_ + mbx];
		memset(&d, 0, sizeof(d));
		const MbInfo &cur = ctx_.info(mbx, mby);
		bool intra = kind_is_intra(cur.kind);
		d.qp_y = (int8_t)dbk_qp;
		if (cur.kind == MB_IPCM) {
			// the reference stores qpc - qp here (h264.cpp:4750-4751), not QPc(0)
			d.qp_cb = (int8_t)(chroma_qp(qp_, pps.chroma_qp_offset[0]) - qp_);
			d.qp_cr = (int8_t)(chroma_qp(qp_, pps.chroma_qp_offset[1]) - qp_);
		} else {
			d.qp_cb = (int8_t)chroma_qp(dbk_qp, pps.chroma_qp_offset[0]);
			d.qp_cr = (int8_t)chroma_qp(dbk_qp, pps.chroma_qp_offset[1]);
		}
		d.alpha_off = (int8_t)h.alpha_off;
		d.beta_off = (int8_t)h.beta_off;
		// what the device needs to derive the strengths itself (b2_h264_picture.bs_on_device)
		const uint8_t *nnz = ctx_.nnz_of(mbx, mby);
		{
			// bit i = nnz[i] != 0: one compare of the sixteen counts with zero
			const unsigned nzm = ~(unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i *>(nnz)), _mm_setzero_si128())) & 0xffffu;
			mbs_[(size_t)mby * mb_w_ + mbx].nnz4x4 = (uint16_t)nzm;
			d.bs_mode = (uint8_t)((bs_reduced_ & 7) | (bslice_ ? B2_BS_B_SLICE : 0));
		}
		if (h.disable_deblock == 1) { d.flags = B2_DBK_OFF; return; }
		any_deblock_ = true;
		bool left_ok = mbx > 0 && (h.disable_deblock == 0 || ctx_.mb_avail(mbx - 1, mby, slice_num_));
		bool top_ok = mby > 0 && (h.disable_deblock == 0 || ctx_.mb_avail(mbx, mby - 1, slice_num_));
		if (!host_bs_) {
			if (left_ok) d.flags |= B2_DBK_LEFT_EDGE;
			if (top_ok) d.flags |= B2_DBK_TOP_EDGE;
			return;
		}
		uint32_t bsv = 0, bsh = 0;
		if (intra) {
			// inner edges bS 3; MB edges bS 4 (flagged, stored as 3)
			bsv = bsh = 0xffffff00u;
			if (left_ok) { bsv |= 0xff; d.flags |= B2_DBK_LEFT_BS4; }
			if (top_ok) { bsh |= 0xff; d.flags |= B2_DBK_TOP_BS4; }
		} else if (!bslice_) {
			// P slices (one list): every pair of 4x4 blocks across an edge is 2 when either block has
			// coefficients, else 1 when the reference pictures differ or a vector component differs by four
			// quarter samples or more.  Done for the 16 blocks at once: one 16-bit mask per rule and
			// direction (bit y * 4 + x = the edge on the left of / above block (x, y)), spread into the
			// 2-bit fields of the record through two 16-entry tables.
			const size_t w4 = (size_t)ctx_.w4;
			const size_t g00 = (size_t)mby * 4 * w4 + (size_t)mbx * 4;
			unsigned nzm = 0;
			for (int i = 0; i < 16; ++i) nzm |= (unsigned)(nnz[i] != 0) << i;
			const bool left_inter = left_ok && !kind_is_intra(ctx_.info(mbx - 1, mby).kind);
			const bool top_inter = top_ok && !kind_is_intra(ctx_.info(mbx, mby - 1).kind);
			unsigned two_v = (nzm | (nzm << 1)) & 0xEEEEu, two_h = (nzm | (nzm << 4)) & 0xFFF0u;
			if (left_inter) {
				const uint8_t *ln = ctx_.nnz_of(mbx - 1, mby);
				for (int sg = 0; sg < 4; ++sg) two_v |= (unsigned)((ln[sg * 4 + 3] | nnz[sg * 4]) != 0) << (sg * 4);
			}
			if (top_inter) {
				const uint8_t *tn = ctx_.nnz_of(mbx, mby - 1);
				for (int sg = 0; sg < 4; ++sg) two_h |= (unsigned)((tn[12 + sg] | nnz[sg]) != 0) << sg;
			}
			unsigned far_v = 0, far_h = 0;
			if (mbs_[(size_t)mby * mb_w_ + mbx].mc_cnt != 1 || left_inter || top_inter) {
				const int16_t *mvb = ctx_.mv[0].data();
				const int8_t *rpb = ctx_.refpic[0].data();
				const __m128i three = _mm_set1_epi16(3);
				for (int y = 0; y < 4; ++y) {
					const size_t g = g00 + y * w4;
					const __m128i cur = _mm_loadu_si128(reinterpret_cast<const __m128i *>(mvb + g * 2));
					int32_t rc;
					memcpy(&rc, rpb + g, 4);
					const __m128i rcv = _mm_cvtsi32_si128(rc);
					// the block on the left: same row, one block earlier
					{
						// (first column of the picture: the same row shifted, whatever lands in block 0 is masked below)
						const __m128i oth = mbx ? _mm_loadu_si128(reinterpret_cast<const __m128i *>(mvb + (g - 1) * 2)) : _mm_slli_si128(cur, 4);
						const __m128i d = _mm_max_epi16(_mm_subs_epi16(cur, oth), _mm_subs_epi16(oth, cur));
						const __m128i f = _mm_cmpgt_epi16(d, three);                          // per component
						const int mvbits = _mm_movemask_ps(_mm_castsi128_ps(_mm_cmpeq_epi32(f, _mm_setzero_si128()))) ^ 15;
						int32_t ro;
						if (mbx) memcpy(&ro, rpb + g - 1, 4);
						else ro = (int32_t)((uint32_t)rc << 8);
						const int rbits = (_mm_movemask_epi8(_mm_cmpeq_epi8(rcv, _mm_cvtsi32_si128(ro))) & 15) ^ 15;
						far_v |= (unsigned)(mvbits | rbits) << (y * 4);
					}
					// the block above
					if (y > 0 || mby > 0) {
						const size_t gu = g - w4;
						const __m128i oth = _mm_loadu_si128(reinterpret_cast<const __m128i *>(mvb + gu * 2));
						const __m128i d = _mm_max_epi16(_mm_subs_epi16(cur, oth), _mm_subs_epi16(oth, cur));
						const __m128i f = _mm_cmpgt_epi16(d, three);
						const int mvbits = _mm_movemask_ps(_mm_castsi128_ps(_mm_cmpeq_epi32(f, _mm_setzero_si128()))) ^ 15;
						int32_t ro;
						memcpy(&ro, rpb + gu, 4);
						const int rbits = (_mm_movemask_epi8(_mm_cmpeq_epi8(rcv, _mm_cvtsi32_si128(ro))) & 15) ^ 15;
						far_h |= (unsigned)(mvbits | rbits) << (y * 4);
					}
				}
			}
			const unsigned mask_v = left_inter ? 0xFFFFu : 0xEEEEu, mask_h = top_inter ? 0xFFFFu : 0xFFF0u;
			const unsigned one_v = far_v & ~two_v & mask_v, one_h = far_h & ~two_h & mask_h;
			// vertical edges: field of (edge x, segment y) at bit x * 8 + y * 2; horizontal: (edge y, segment x) at y * 8 + x * 2
			static const uint32_t spread_v[16] = {0x00000000u, 0x00000001u, 0x00000100u, 0x00000101u, 0x00010000u, 0x00010001u, 0x00010100u, 0x00010101u,
			                                      0x01000000u, 0x01000001u, 0x01000100u, 0x01000101u, 0x01010000u, 0x01010001u, 0x01010100u, 0x01010101u};
			static const uint32_t spread_h[16] = {0x00, 0x01, 0x04, 0x05, 0x10, 0x11, 0x14, 0x15, 0x40, 0x41, 0x44, 0x45, 0x50, 0x51, 0x54, 0x55};
			for (int y = 0; y < 4; ++y) {
				bsv |= ((spread_v[(two_v >> (4 * y)) & 15] << 1) | spread_v[(one_v >> (4 * y)) & 15]) << (2 * y);
				bsh |= ((spread_h[(two_h >> (4 * y)) & 15] << 1) | spread_h[(one_h >> (4 * y)) & 15]) << (8 * y);
			}
			if (left_ok && !left_inter) { bsv |= 0xff; d.flags |= B2_DBK_LEFT_BS4; }
			if (top_ok && !top_inter) { bsh |= 0xff; d.flags |= B2_DBK_TOP_BS4; }
#ifdef B2_CHECK_BS
			{
				uint32_t cv = 0, ch = 0;
				for (int e = 1; e < 4; ++e)
					for (int sg = 0; sg < 4; ++sg) {
						cv |= (uint32_t)bs_inter(g00 + sg * w4 + e - 1, g00 + sg * w4 + e, (nzm >> (sg * 4 + e - 1)) & 1, (nzm >> (sg * 4 + e)) & 1) << (e * 8 + sg * 2);
						ch |= (uint32_t)bs_inter(g00 + (e - 1) * w4 + sg, g00 + e * w4 + sg, (nzm >> ((e - 1) * 4 + sg)) & 1, (nzm >> (e * 4 + sg)) & 1) << (e * 8 + sg * 2);
					}
				if (left_ok) {
					if (!left_inter) cv |= 0xff;
					else { const uint8_t *ln = ctx_.nnz_of(mbx - 1, mby); for (int sg = 0; sg < 4; ++sg) cv |= (uint32_t)bs_inter(g00 + sg * w4 - 1, g00 + sg * w4, ln[sg * 4 + 3] != 0, nnz[sg * 4] != 0) << (sg * 2); }
				}
				if (top_ok) {
					if (!top_inter) ch |= 0xff;
					else { const uint8_t *tn = ctx_.nnz_of(mbx, mby - 1); for (int sg = 0; sg < 4; ++sg) ch |= (uint32_t)bs_inter(g00 - w4 + sg, g00 + sg, tn[12 + sg] != 0, nnz[sg] != 0) << (sg * 2); }
				}
				if (cv != bsv || ch != bsh) fprintf(stderr, "bs mismatch mb %d,%d: v %08x want %08x  h %08x want %08x  (nzm %04x far_v %04x far_h %04x mc %d)\n", mbx, mby, bsv, cv, bsh, ch, nzm, far_v, far_h, mbs_[(size_t)mby * mb_w_ + mbx].mc_cnt);
			}
#endif
		} else {
			const size_t w4 = (size_t)ctx_.w4;
			const size_t g00 = (size_t)mby * 4 * w4 + (size_t)mbx * 4;
			unsigned nzm = 0;
			for (int i = 0; i < 16; ++i) nzm |= (unsigned)(nnz[i] != 0) << i;
			if (bslice_ && mbs_[(size_t)mby * mb_w_ + mbx].mc_cnt != 1) {
				for (int e = 1; e < 4; ++e) {
					const bool red = (bs_reduced_ >> (e == 2 ? 1 : 2)) & 1;
					for (int sg = 0; sg < 4; ++sg) {
						bsv |= (uint32_t)bs_inter(g00 + sg * w4 + e - 1, g00 + sg * w4 + e, (nzm >> (sg * 4 + e - 1)) & 1, (nzm >> (sg * 4 + e)) & 1, red) << (e * 8 + sg * 2);
						bsh |= (uint32_t)bs_inter(g00 + (e - 1) * w4 + sg, g00 + e * w4 + sg, (nzm >> ((e - 1) * 4 + sg)) & 1, (nzm >> (e * 4 + sg)) & 1, red) << (e * 8 + sg * 2);
					}
				}
			} else if (mbs_[(size_t)mby * mb_w_ + mbx].mc_cnt == 1) {
				// one partition: no motion discontinuity inside the macroblock
				for (int e = 1; e < 4; ++e)
					for (int sg = 0; sg < 4; ++sg) {
						bsv |= (uint32_t)((((nzm >> (sg * 4 + e - 1)) | (nzm >> (sg * 4 + e))) & 1) * 2) << (e * 8 + sg * 2);
						bsh |= (uint32_t)((((nzm >> ((e - 1) * 4 + sg)) | (nzm >> (e * 4 + sg))) & 1) * 2) << (e * 8 + sg * 2);
					}
			} else {
				// local copy of the macroblock's motion (raster 4x4) keeps the 48 comparisons cheap
				int16_t mvx[16], mvy[16];
				int8_t rp[16];
				for (int yy = 0; yy < 4; ++yy)
					for (int xx = 0; xx < 4; ++xx) {
						size_t g = g00 + yy * w4 + xx;
						mvx[yy * 4 + xx] = ctx_.mv[0][g * 2];
						mvy[yy * 4 + xx] = ctx_.mv[0][g * 2 + 1];
						rp[yy * 4 + xx] = ctx_.refpic[0][g];
					}
				auto pair = [&](int i0, int i1) -> uint32_t {
					if (((nzm >> i0) | (nzm >> i1)) & 1) return 2;
					if (rp[i0] != rp[i1]) return 1;
					return ((unsigned)(mvx[i0] - mvx[i1] + 3) > 6u || (unsigned)(mvy[i0] - mvy[i1] + 3) > 6u) ? 1 : 0;
				};
				for (int e = 1; e < 4; ++e)
					for (int sg = 0; sg < 4; ++sg) {
						bsv |= pair(sg * 4 + e - 1, sg * 4 + e) << (e * 8 + sg * 2);
						bsh |= pair((e - 1) * 4 + sg, e * 4 + sg) << (e * 8 + sg * 2);
					}
			}
			if (left_ok) {
				if (kind_is_intra(ctx_.info(mbx - 1, mby).kind)) { bsv |= 0xff; d.flags |= B2_DBK_LEFT_BS4; }
				else {
					const uint8_t *ln = ctx_.nnz_of(mbx - 1, mby);
					for (int sg = 0; sg < 4; ++sg)
						bsv |= (uint32_t)bs_inter(g00 + sg * w4 - 1, g00 + sg * w4, ln[sg * 4 + 3] != 0, nnz[sg * 4] != 0, bs_reduced_ & 1) << (sg * 2);
				}
			}
			if (top_ok) {
				if (kind_is_intra(ctx_.info(mbx, mby - 1).kind)) { bsh |= 0xff; d.flags |= B2_DBK_TOP_BS4; }
				else {
					const uint8_t *tn = ctx_.nnz_of(mbx, mby - 1);
					for (int sg = 0; sg < 4; ++sg)
						bsh |= (uint32_t)bs_inter(g00 - w4 + sg, g00 + sg, tn[12 + sg] != 0, nnz[sg] != 0, bs_reduced_ & 1) << (sg * 2);
				}
			}
		}
		if (cur.t8x8) {
			// transform_size_8x8_flag: the 4x4 edges inside the 8x8 blocks are not filtered (8.7)
			bsv &= 0x00ff00ffu;
			bsh &= 0x00ff00ffu;
		}
		d.bs_v = bsv;
		d.bs_h = bsh;
		if (left_ok) d.flags |= B2_DBK_LEFT_EDGE;
		if (top_ok) d.flags |= B2_DBK_TOP_EDGE;
	}

	// ---- state --------------------------------------------------------------------------------------
	dec_bits *stream_;
	BitReader br_;
	std::vector<uint8_t> rbsp_;
	long stop_bit_pos_, total_bits_;
	int (*header_cb_)(void *, void *);
	void *header_arg_;
	Sps sps_[32];
	Pps pps_[256];
	int active_sps_;
	int nal_ref_idc_, nal_type_;
	Dpb dpb_;
	int num_frames_;
	m2d_frame_t frames_[MAX_H264_FRAMES];
	int8_t lru_[MAX_H264_FRAMES];
	bool host_valid_[MAX_H264_FRAMES];
	int cur_frame_;
	bool pic_open_;
	uint32_t prev_first_mb_;
	const H264Tables *tab_;             // the CAVLC tables (built once per process)
	int frames_w_ = 0, frames_h_ = 0;   // coded size set_frames was called for
	std::vector<RefPic> refs_;       // pictures marked "used for reference"
	std::vector<RefPic> list_[2], init0_;
	ColPic col_[MAX_H264_FRAMES];
	bool bslice_ = false, direct_8x8_inference_ = true;
	bool cabac_ = false;
	int bs_reduced_ = 0;        // see set_shape
	bool host_bs_ = true;       // this picture's deblocking strengths are derived here, not on the device (begin_picture)
	bool dev_mv_ = false;       // this picture's motion vectors are predicted on the device (begin_picture)
	bool keep_parts_ = false;   // host derivation, but the partitions are recorded too (what a tap's owner checks against)
	std::vector<b2_h264_part> parts_;
	uint32_t item_bound_ = 0;
	int last_dqp_ = 0;          // mb_qp_delta of the previous macroblock in decoding order (CABAC ctxIdxInc)
	CabacSyntax<CabacDec> cs_;
	struct { int mode, denom[2]; int8_t w[2][32][3], o[2][32][3]; } wt_;   // mode 0 default, 1 explicit, 2 implicit
	int prev_frame_num_, poc_lsb_, poc_msb_, frame_num_offset_, cur_poc_;
	bool prev_mmco5_;
	bool unsupported_;
	SliceHeader hdr_;
	int mb_w_, mb_h_, mb_addr_, slice_num_, qp_;
	PicCtx ctx_;
	std::vector<b2_h264_mb> mbs_;
	std::vector<b2_h264_dbk> dbk_;
	std::vector<uint8_t> mb_seen_;          // macroblock addresses delivered in the open picture
	std::vector<b2_h264_mc> mcs_;
	int cur_mbx_ = 0, cur_mby_ = 0;                           // macroblock being parsed (set by open_mb)
	std::vector<uint32_t> items_[72], items_first_[72], items_all_;      // inter work items per (class, height) bucket / concatenated
	std::vector<int16_t> coefs_;
	size_t coef_used_ = 0;
	uint32_t n_intra_;
	bool any_deblock_;
};

inline H264Decoder *impl(void *ctx) {
	CtxBlob *b = (CtxBlob *)ctx;
	if (!b || b->magic != kCtxMagic || b->codec != B2_CODEC_H264) return nullptr;
	return (H264Decoder *)b->impl;
}

int h4_init(void *ctx, int dpb_max, int (*cb)(void *, void *), void *arg) {
	if (!ctx) return -1;
	CtxBlob *b = (CtxBlob *)ctx;
	memset(b, 0, sizeof(*b));
	b->magic = kCtxMagic;
	b->codec = B2_CODEC_H264;
	dec_bits_open(&b->stream, nullptr);
	H264Decoder *d = new (std::nothrow) H264Decoder(&b->stream);
	if (!d) return -1;
	d->reset(dpb_max, cb, arg);
	b->impl = d;
	return 0;
}
dec_bits *h4_stream_pos(void *ctx) { return ctx ? &((CtxBlob *)ctx)->stream : nullptr; }
int h4_get_info(void *ctx, m2d_info_t *info) { H264Decoder *d = impl(ctx); return d ? d->get_info(info) : -1; }
int h4_set_frames(void *ctx, int n, m2d_frame_t *f, uint8_t *, int) { H264Decoder *d = impl(ctx); return d ? d->set_frames(n, f) : -1; }
int h4_decode(void *ctx) { H264Decoder *d = impl(ctx); return d ? d->decode_picture() : -1; }
int h4_peek(void *ctx, m2d_frame_t *f, int e) { H264Decoder *d = impl(ctx); return d ? d->peek_frame(f, e) : -1; }
int h4_get(void *ctx, m2d_frame_t *f, int e) { H264Decoder *d = impl(ctx); return d ? d->get_frame(f, e) : -1; }

const m2d_func_table_t h4_table = {
	sizeof(CtxBlob), h4_init, h4_stream_pos, h4_get_info, h4_set_frames, h4_decode, h4_peek, h4_get,
};

}  // namespace

Backend *h264_backend(void *ctx) {
	H264Decoder *d = impl(ctx);
	return d ? &d->backend : nullptr;
}
void h264_destroy(void *ctx) {
	H264Decoder *d = impl(ctx);
	if (d) { delete d; ((CtxBlob *)ctx)->impl = nullptr; }
}

}  // namespace b2

extern "C" const m2d_func_table_t *const h264d_gpu_func = &b2::h4_table;
