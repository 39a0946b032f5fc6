// validate.cpp — structural checks of a picture's SoA records before they are sent to the device.
// The kernels index frames, records and work items with what the records say and check nothing
// (an out-of-range index there is an illegal address that takes the whole CUDA context down, every
// stream of the process with it), so the C ABI checks here: every index inside its array, every
// enumerated field inside its range.  O(records) integer work per picture, no pixels.
#include <stdio.h>
#include "b2_recon.h"

namespace b2 {
void set_error(const char *fmt, ...);
}

extern "C" int b2_h264_picture_check(const b2_h264_picture *p, int nframes) {
	using b2::set_error;
	if (!p || p->mb_w <= 0 || p->mb_h <= 0 || (uint64_t)p->mb_w * p->mb_h != p->n_mbs || p->n_mbs > (1u << 20)) {
		set_error("H.264 picture: bad geometry");
		return -1;
	}
	if ((unsigned)p->dst >= (unsigned)nframes) { set_error("H.264 picture: destination frame %d outside the pool", p->dst); return -1; }
	if (!p->mbs || !p->dbk || (p->n_mcs && !p->mcs) || (p->n_coef_recs && !p->coefs) || (!p->mv_on_device && p->n_items && !p->items) || p->n_items_first > p->n_items) {
		set_error("H.264 picture: missing arrays");
		return -1;
	}
	if (p->mv_on_device && (!p->bs_on_device || p->n_mcs || p->n_items_first || (p->n_parts && !p->parts))) {
		set_error("H.264 picture: mv_on_device needs bs_on_device, partitions and no motion records");
		return -1;
	}
	const uint32_t n_recs = p->mv_on_device ? p->n_parts : p->n_mcs;
	uint32_t n_intra = 0;
	for (uint32_t i = 0; i < p->n_mbs; ++i) {
		const b2_h264_mb &m = p->mbs[i];
		const uint32_t nrec = (uint32_t)__builtin_popcount(m.nz);
		bool ok = m.type <= B2_H4_IPCM && m.qp_y <= 51 && m.qp_cb <= 51 && m.qp_cr <= 51 && m.i16_mode <= 3 && m.chroma_mode <= 3 &&
		          !(m.nz >> 26) && (uint64_t)m.coef_idx + nrec <= p->n_coef_recs && (uint64_t)m.mc_idx + m.mc_cnt <= n_recs && m.mc_cnt <= 16;
		if (ok && m.type == B2_H4_IPCM) ok = m.nz == 0xfff;
		if (ok && m.type != B2_H4_INTER) { ok = m.mc_cnt == 0; n_intra++; }
		if (ok && (m.flags & B2_H4_T8x8))
			for (int q = 0; q < 4; ++q) { const uint32_t g = (m.nz >> (1 + 4 * q)) & 15; ok &= g == 0 || g == 15; }   // 64 levels = four records
		if (ok && (m.type == B2_H4_I4x4 || m.type == B2_H4_I8x8))
			for (int k = 0; k < 8; ++k) ok &= (m.i4_modes[k] & 15) <= 8 && (m.i4_modes[k] >> 4) <= 8;
		if (!ok) { set_error("H.264 picture: macroblock %u has a field out of range", i); return -1; }
		// the kernels take a macroblock's neighbours from these flags: none may point outside the picture
		{
			const uint32_t mbx = i % (uint32_t)p->mb_w, mby = i / (uint32_t)p->mb_w;
			const int fl = p->dbk[i].flags, nb = p->mv_on_device ? m.flags >> B2_H4_NB_SHIFT : 0;
			if (((fl & B2_DBK_LEFT_EDGE) && mbx == 0) || ((fl & B2_DBK_TOP_EDGE) && mby == 0) ||
			    ((nb & 1) && mbx == 0) || ((nb & 2) && mby == 0) || ((nb & 4) && (mby == 0 || mbx + 1 >= (uint32_t)p->mb_w)) || ((nb & 8) && (mbx == 0 || mby == 0))) {
				set_error("H.264 picture: macroblock %u names a neighbour outside the picture", i);
				return -1;
			}
		}
		const b2_h264_dbk &d = p->dbk[i];
		if (d.qp_y < -51 || d.qp_y > 51 || d.qp_cb < -51 || d.qp_cb > 51 || d.qp_cr < -51 || d.qp_cr > 51 || d.alpha_off < -12 || d.alpha_off > 12 || d.beta_off < -12 || d.beta_off > 12) {
			set_error("H.264 picture: deblocking record %u out of range", i);
			return -1;
		}
	}
	if (n_intra != p->n_intra) { set_error("H.264 picture: n_intra does not match the macroblock types"); return -1; }
	if (p->mv_on_device) {
		uint64_t cols = 0;
		for (uint32_t i = 0; i < p->n_parts; ++i) {
			const b2_h264_part &c = p->parts[i];
			const int x4 = c.pos & 3, y4 = c.pos >> 2, w4 = c.size & 15, h4 = c.size >> 4;
			const bool size_ok = (w4 == 1 || w4 == 2 || w4 == 4) && (h4 == 1 || h4 == 2 || h4 == 4) && x4 + w4 <= 4 && y4 + h4 <= 4 && y4 <= 3;
			if (!size_ok || c.ref_idx < 0 || c.ref_frame < 0 || c.ref_frame >= nframes) {
				set_error("H.264 picture: partition %u out of range (frame %d of %d)", i, c.ref_frame, nframes);
				return -1;
			}
			cols += (uint64_t)w4;
		}
		if (cols > p->n_items) { set_error("H.264 picture: n_items is not an upper bound of the work items"); return -1; }
		return 0;
	}
	for (uint32_t i = 0; i < p->n_mcs; ++i) {
		const b2_h264_mc &c = p->mcs[i];
		const int x4 = c.pos & 3, y4 = c.pos >> 2, w4 = c.size & 15, h4 = c.size >> 4;
		const bool size_ok = (w4 == 1 || w4 == 2 || w4 == 4) && (h4 == 1 || h4 == 2 || h4 == 4) && x4 + w4 <= 4 && y4 + h4 <= 4 && y4 <= 3;
		const bool refs_ok = c.ref_frame[0] >= -1 && c.ref_frame[0] < nframes && c.ref_frame[1] >= -1 && c.ref_frame[1] < nframes && (c.ref_frame[0] >= 0 || c.ref_frame[1] >= 0);
		if (!size_ok || !refs_ok || c.log2wd_y > 7 || c.log2wd_c > 7) {
			set_error("H.264 picture: motion record %u out of range (frames %d / %d of %d)", i, c.ref_frame[0], c.ref_frame[1], nframes);
			return -1;
		}
	}
	for (uint32_t i = 0; i < p->n_items; ++i) {
		const uint32_t it = p->items[i], mb = it & 0xfffff, ri = (it >> 20) & 15, k = (it >> 24) & 3, cls = (it >> 26) & 15;
		bool ok = mb < p->n_mbs && cls < 12;
		if (ok) {
			const b2_h264_mb &m = p->mbs[mb];
			ok = ri < m.mc_cnt;
			if (ok) {
				const b2_h264_mc &c = p->mcs[m.mc_idx + ri];
				const int w4 = c.size & 15, strip_w4 = cls >= 6 ? 2 : 1;
				ok = (int)(k + 1) * strip_w4 <= w4 && (i >= p->n_items_first || (c.ref_frame[0] >= 0 && c.ref_frame[1] >= 0));
			}
		}
		if (!ok) { set_error("H.264 picture: work item %u out of range", i); return -1; }
	}
	return 0;
}

extern "C" int b2_mpeg2_picture_check(const b2_m2_picture *p, int nframes) {
	using b2::set_error;
	if (!p || p->mb_w <= 0 || p->mb_h <= 0 || (uint64_t)p->mb_w * p->mb_h != p->n_mbs || p->n_mbs > (1u << 20) || !p->mbs || (p->n_blocks && !p->coefs)) {
		set_error("MPEG-2 picture: bad geometry");
		return -1;
	}
	if ((unsigned)p->dst >= (unsigned)nframes || (unsigned)p->fwd >= (unsigned)nframes || (unsigned)p->bwd >= (unsigned)nframes) {
		set_error("MPEG-2 picture: frame index outside the pool");
		return -1;
	}
	for (uint32_t i = 0; i < p->n_mbs; ++i) {
		const b2_m2_mb &m = p->mbs[i];
		if ((m.cbp >> 6) || (uint64_t)m.coef_idx + (uint32_t)__builtin_popcount(m.cbp) > p->n_blocks) {
			set_error("MPEG-2 picture: macroblock %u points outside the coefficient records", i);
			return -1;
		}
	}
	return 0;
}
