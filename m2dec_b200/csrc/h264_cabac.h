// h264_cabac.h — CABAC (ITU-T H.264 clause 9.3): context models, the arithmetic decoding
// engine, and the binarisation / context selection of every slice-data syntax element the
// front end parses (frame pictures, 4:2:0).
//
// The syntax layer (CabacSyntax<E>) is written once over an engine type E with
//     int bin(int ctxIdx, int v);  int bypass(int v);  int term(int v);
// The decoding engine ignores `v` and returns the decoded bin; the test-only encoding engine
// (tools/streamgen/cabac_enc.h) codes `v` and returns it.  Every binarisation below steers its
// control flow by the *returned* bins only, so the same code parses a stream and writes one.
// That a shared mistake cannot hide is settled by decoding every generated stream with the
// compiled reference and with ffmpeg (tests/).
//
// Behavioural model: reference src/lib/m2d.h:130-279 (engine), src/lib/h264.cpp:11057-11600
// (syntax elements); both follow the standard, which is what is restated here.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "bits.h"
#include "h264_cabac_tables.h"
#include "h264_common.h"

namespace b2 { namespace h4 {

// 9.3.1.1: context variables, state = (pStateIdx << 1) | valMPS
struct CabacModels {
	uint8_t st[460];
	// table: 0 for I slices, 1 + cabac_init_idc otherwise
	void init(int slice_qp, int table) {
		int qp = slice_qp < 0 ? 0 : (slice_qp > 51 ? 51 : slice_qp);
		for (int i = 0; i < 460; ++i) {
			int m = h4t::cabac_mn[i][table][0], n = h4t::cabac_mn[i][table][1];
			int pre = ((m * qp) >> 4) + n;
			pre = pre < 1 ? 1 : (pre > 126 ? 126 : pre);
			st[i] = pre <= 63 ? (uint8_t)((63 - pre) << 1) : (uint8_t)(((pre - 64) << 1) | 1);
		}
	}
};

// 9.3.3.2.1.1 for both outcomes: t[s] after the most probable symbol, t[128 + s] after the least probable one
struct CabacNext {
	uint8_t t[256];
	constexpr CabacNext() : t() {
		for (int s = 0; s < 128; ++s) {
			const int p = s >> 1, mps = s & 1;
			t[s] = (uint8_t)(p < 62 ? s + 2 : s);
			t[128 + s] = (uint8_t)((h4t::trans_idx_lps[p] << 1) | (p == 0 ? mps ^ 1 : mps));
		}
	}
};
static constexpr CabacNext cabac_next{};

// 9.3.3.2: arithmetic decoding engine.  codIOffset is kept with `bits_` bits of the stream already behind it
// (value_ = codIOffset << bits_ | the next bits_ bits of the stream): renormalisation and bypass bins then only move
// bits_, and the stream is touched once per 32 bits instead of once per renormalisation.  The comparison
// codIOffset >= codIRange becomes value_ >= codIRange << bits_.  The engine reads the slice data itself, from the
// byte position the front end's bit reader stands at when start() is called; the bit reader is moved only for I_PCM.
class CabacDec {
public:
	CabacModels m;
	void start(BitReader *br) {          // (the reader is byte aligned: cabac_alignment_one_bit / after pcm samples)
		br_ = br;
		start_ = p_ = br->byte_ptr();
		end_ = br->end_ptr();
		range_ = 510;
		value_ = 0;
		bits_ = -9;                       // nine bits of the first fetch are codIOffset itself
		fetched_ = 0;
		refill();
	}
	// Straight-line: which of the two sub-intervals the offset falls into is a coin the branch predictor loses a
	// quarter of the time, so the interval, the offset and the next state are selected with a mask (the state transition
	// of both outcomes comes from one table, cabac_next.t[LPS ? 128 + s : s]).
	inline int bin(int ctx, int /*unused*/) {
		const uint32_t s = m.st[ctx];
		const uint32_t lps = h4t::range_tab_lps[s >> 1][(range_ >> 6) & 3];
		const uint32_t rmps = range_ - lps;
		const uint64_t scaled = (uint64_t)rmps << bits_;
		const uint64_t is_lps = (uint64_t)0 - (uint64_t)(value_ >= scaled);      // all ones: least probable symbol
		value_ -= scaled & is_lps;
		uint32_t r = rmps ^ ((rmps ^ lps) & (uint32_t)is_lps);
		m.st[ctx] = cabac_next.t[s + ((uint32_t)is_lps & 128u)];
		const int sh = __builtin_clz(r) - 23;              // 0 when the range is still >= 256; at most 6
		range_ = r << sh;
		bits_ -= sh;
		if (bits_ < 8) refill();
		return (int)((s ^ (uint32_t)is_lps) & 1);
	}
	inline int bypass(int /*unused*/) {          // (signs and escape suffixes: a fair coin, so no branch either)
		--bits_;
		const uint64_t scaled = (uint64_t)range_ << bits_;
		const uint64_t one = (uint64_t)0 - (uint64_t)(value_ >= scaled);
		value_ -= scaled & one;
		if (bits_ < 8) refill();
		return (int)(one & 1);
	}
	inline int term(int /*unused*/) {
		range_ -= 2;
		if (value_ >= ((uint64_t)range_ << bits_)) return 1;
		if (range_ < 256) {
			range_ <<= 1;
			if (--bits_ < 8) refill();
		}
		return 0;
	}
	// After term() == 1 for mb_type I_PCM.  The nine bits the offset register holds at that
	// point are exactly what the encoder's flush (9.3.4.5) wrote, its first-ever output bit
	// having been dropped (firstBitFlag), so the position after the bits this engine has consumed already sits after the
	// flush and only pcm_alignment_zero_bit remains (the reference does the same, h264.cpp:4741): the bit reader, which
	// still stands where start() found it, is moved there.
	void before_pcm() { br_->skip_bytes((size_t)((consumed() + 7) >> 3)); }
	// true once the engine has consumed bits beyond the end of the data (what it fetched ahead does not count)
	bool overrun() const { return consumed() > (long)(end_ - start_) * 8; }
private:
	long consumed() const { return fetched_ - bits_; }
	inline void refill() {
		uint32_t v;
		if (end_ - p_ >= 4) {
			memcpy(&v, p_, 4);
			v = __builtin_bswap32(v);
			p_ += 4;
		} else {                          // past the end: zeros (overrun() tells)
			v = 0;
			for (int i = 0; i < 4; ++i) v = (v << 8) | (p_ < end_ ? *p_++ : 0u);
		}
		value_ = (value_ << 32) | v;
		bits_ += 32;
		fetched_ += 32;
	}
	BitReader *br_ = nullptr;
	const uint8_t *start_ = nullptr, *p_ = nullptr, *end_ = nullptr;
	uint64_t value_ = 0;
	uint32_t range_ = 510;
	int bits_ = 0;
	long fetched_ = 0;
};

// Table 9-34 ctxIdxOffset per ctxBlockCat for coded_block_flag / significant_coeff_flag /
// last_significant_coeff_flag / coeff_abs_level_minus1 (frame coded); cat 5 = luma 8x8
static const int16_t cabac_cbf_base[5] = {85, 89, 93, 97, 101};
static const int16_t cabac_sig_base[6] = {105, 120, 134, 149, 152, 402};
static const int16_t cabac_last_base[6] = {166, 181, 195, 210, 213, 417};
static const int16_t cabac_abs_base[6] = {227, 237, 247, 257, 266, 426};
// Table 9-43: ctxIdxInc of significant_coeff_flag (frame coded) and last_significant_coeff_flag
// for ctxBlockCat 5, by levelListIdx
static const uint8_t cabac_sig8x8[63] = {
	0, 1, 2, 3, 4, 5, 5, 4, 4, 3, 3, 4, 4, 4, 5, 5, 4, 4, 4, 4, 3, 3, 6, 7, 7, 7, 8, 9, 10, 9, 8, 7,
	7, 6, 11, 12, 13, 11, 6, 7, 8, 9, 14, 10, 9, 8, 6, 11, 12, 13, 11, 6, 9, 14, 10, 9, 11, 12, 13, 11, 14, 10, 12};
static const uint8_t cabac_last8x8[63] = {
	0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2,
	3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8};

template <class E>
class CabacSyntax {
public:
	E e;

	int mb_skip(bool bslice, int inc, int v) { return e.bin((bslice ? 24 : 11) + inc, v); }
	int end_of_slice(int v) { return e.term(v); }

	// mb_type of an intra macroblock (0 I_NxN, 1..24 Intra16x16, 25 I_PCM).  base = 3 in I
	// slices (with the neighbour-dependent first bin), 17 / 32 as the suffix in P / B slices.
	int mb_type_intra(int base, bool islice, int inc, int v) {
		if (!e.bin(base + (islice ? inc : 0), v != 0)) return 0;
		if (e.term(v == 25)) return 25;
		const int t = v - 1, o = islice ? base + 2 : base, is = islice ? 1 : 0;
		const int want_c = (t >> 2) % 3;
		int r = 1 + 12 * e.bin(o + 1, t >= 12);
		if (e.bin(o + 2, want_c != 0)) r += 4 + 4 * e.bin(o + 2 + is, want_c == 2);
		r += 2 * e.bin(o + 3 + is, (t >> 1) & 1);
		r += e.bin(o + 3 + 2 * is, t & 1);
		return r;
	}
	// P slice: 0 16x16, 1 16x8, 2 8x16, 3 8x8, 5.. intra (numbering of Table 7-13 + 5)
	int mb_type_p(int v) {
		if (e.bin(14, v >= 5)) return 5 + mb_type_intra(17, false, 0, v - 5);
		if (e.bin(15, v == 1 || v == 2)) return e.bin(17, v == 1) ? 1 : 2;
		return e.bin(16, v == 3) ? 3 : 0;
	}
	// B slice: Table 7-14 numbering, 23.. intra
	int mb_type_b(int inc, int v) {
		if (!e.bin(27 + inc, v != 0)) return 0;
		if (!e.bin(30, v > 2)) return 1 + e.bin(32, v == 2);
		const int code = (v >= 3 && v <= 10) ? v - 3 : (v == 11 ? 14 : (v == 22 ? 15 : (v >= 23 ? 13 : (v + 4) >> 1)));
		int bits = e.bin(31, (code >> 3) & 1) << 3;
		bits |= e.bin(32, (code >> 2) & 1) << 2;
		bits |= e.bin(32, (code >> 1) & 1) << 1;
		bits |= e.bin(32, code & 1);
		if (bits < 8) return bits + 3;
		if (bits == 13) return 23 + mb_type_intra(32, false, 0, v - 23);
		if (bits == 14) return 11;
		if (bits == 15) return 22;
		bits = (bits << 1) | e.bin(32, (v + 4) & 1);
		return bits - 4;
	}
	int sub_mb_type_p(int v) {
		if (e.bin(21, v == 0)) return 0;
		if (!e.bin(22, v != 1)) return 1;
		return e.bin(23, v == 2) ? 2 : 3;
	}
	int sub_mb_type_b(int v) {
		if (!e.bin(36, v != 0)) return 0;
		if (!e.bin(37, v > 2)) return 1 + e.bin(39, v == 2);
		int t = 3;
		if (e.bin(38, v >= 7)) {
			if (e.bin(39, v >= 11)) return 11 + e.bin(39, v == 12);
			t = 7;
		}
		const int r = v - t;
		t += 2 * e.bin(39, (r >> 1) & 1);
		t += e.bin(39, r & 1);
		return t;
	}
	int ref_idx(int inc, int v) {
		int m = 0, c = 54 + inc;
		while (m < 32 && e.bin(c, v > m)) { m++; c = m == 1 ? 58 : 59; }
		return m;
	}
	// mvd component: base 40 (x) / 47 (y); sum = absMvdComp of the neighbours A + B (9.3.3.1.1.7)
	int mvd(int base, int sum, int v) {
		const int a = v < 0 ? -v : v;
		if (!e.bin(base + (sum < 3 ? 0 : (sum <= 32 ? 1 : 2)), a != 0)) return 0;
		int m = 1, c = base + 3;
		while (m < 9 && e.bin(c, a > m)) { if (m < 4) c++; m++; }
		if (m >= 9) m += exp_golomb(3, a - 9);
		return e.bypass(v < 0) ? -m : m;
	}
	int qp_delta(int inc, int v) {
		const int k = v > 0 ? 2 * v - 1 : -2 * v;
		if (!e.bin(60 + inc, k != 0)) return 0;
		int m = 1, c = 62;
		while (m < 104 && e.bin(c, k > m)) { m++; c = 63; }
		return (m & 1) ? (m + 1) >> 1 : -(m >> 1);
	}
	int chroma_mode(int inc, int v) {
		if (!e.bin(64 + inc, v != 0)) return 0;
		int m = 1;
		while (m < 3 && e.bin(67, v > m)) m++;
		return m;
	}
	// prev_intra{4x4,8x8}_pred_mode_flag + rem_intra*_pred_mode: returns the mode
	int ipred_mode(int pred, int mode) {
		if (e.bin(68, mode == pred)) return pred;
		const int rem = mode < pred ? mode : mode - 1;
		int r = e.bin(69, rem & 1);
		r |= e.bin(69, (rem >> 1) & 1) << 1;
		r |= e.bin(69, (rem >> 2) & 1) << 2;
		return r < pred ? r : r + 1;
	}
	int transform8x8(int inc, int v) { return e.bin(399 + inc, v); }
	int cbp(const PicCtx &c, int v) {
		int bits = 0;
		for (int b8 = 0; b8 < 4; ++b8) bits |= e.bin(73 + c.cabac_inc_cbp_luma(b8, bits), (v >> b8) & 1) << b8;
		int chroma = 0;
		if (e.bin(77 + c.cabac_inc_cbp_chroma(0), (v >> 4) != 0)) chroma = 1 + e.bin(81 + c.cabac_inc_cbp_chroma(1), (v >> 4) == 2);
		return bits | (chroma << 4);
	}

	// residual_block_cabac (7.3.5.3.3).  lv[0..maxc) holds the levels in scan order (written
	// when decoding, read when encoding).  cbf_inc < 0: coded_block_flag is not coded (cat 5).
	// Returns the number of non-zero levels.
	// scan (decoding only): level k of the scan goes to lv[scan[k]] instead of lv[k] -- the decoder hands over the
	// zeroed coefficient record and the raster position of every scan index, so only the non-zero levels are stored.
	int residual(int cat, int cbf_inc, int16_t *lv, int maxc, const uint8_t *scan = nullptr) {
		int last_nz = -1;
		for (int i = 0; i < maxc; ++i) if (lv[i]) last_nz = i;
		if (cbf_inc >= 0 && !e.bin(cabac_cbf_base[cat] + cbf_inc, last_nz >= 0)) return 0;
		uint8_t idx[64];
		int n = 0, i;
		for (i = 0; i < maxc - 1; ++i) {
			const int inc_s = cat == 5 ? cabac_sig8x8[i] : (cat == 3 ? (i < 2 ? i : 2) : i);
			if (!e.bin(cabac_sig_base[cat] + inc_s, lv[i] != 0)) continue;
			idx[n++] = (uint8_t)i;
			const int inc_l = cat == 5 ? cabac_last8x8[i] : inc_s;
			if (e.bin(cabac_last_base[cat] + inc_l, i == last_nz)) break;
		}
		if (i == maxc - 1) idx[n++] = (uint8_t)i;   // the last position needs no flags
		int eq1 = 0, gt1 = 0;
		const int abs_base = cabac_abs_base[cat], gt1_max = cat == 3 ? 3 : 4;
		for (int k = n - 1; k >= 0; --k) {
			const int at = scan ? scan[idx[k]] : idx[k];
			const int v = lv[at], a = v < 0 ? -v : v;
			int m = 1;
			if (e.bin(abs_base + (gt1 ? 0 : (eq1 < 3 ? 1 + eq1 : 4)), a > 1)) {
				const int c = abs_base + 5 + (gt1 < gt1_max ? gt1 : gt1_max);
				m = 2;
				while (m < 15 && e.bin(c, a > m)) m++;
				if (m == 15) m += exp_golomb(0, a - 15);
				gt1++;
			} else {
				eq1++;
			}
			if (m > 32767) m = 32767;
			lv[at] = (int16_t)(e.bypass(v < 0) ? -m : m);
		}
		return n;
	}

private:
	// k-th order Exp-Golomb suffix in bypass bins (9.3.2.3); returns the value
	int exp_golomb(int k, int rem) {
		int m = 0;
		while (k < 24 && e.bypass(rem >= (1 << k))) { rem -= 1 << k; m += 1 << k; k++; }
		while (k--) m += e.bypass((rem >> k) & 1) << k;
		return m;
	}
};

}}  // namespace b2::h4
