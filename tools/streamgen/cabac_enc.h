// cabac_enc.h — test-only: the CABAC arithmetic *encoding* engine of ITU-T H.264 9.3.4
// (EncodeDecision / EncodeBypass / EncodeTerminate / EncodeFlush, PutBit with outstanding
// bits).  It has the interface CabacSyntax<E> expects (m2dec_b200/csrc/h264_cabac.h), so the
// writer emits every syntax element through the very binarisation code the decoder parses with.
#pragma once
#include "bitwriter.h"
#include "h264_cabac.h"

namespace sg {

class CabacEnc {
public:
	b2::h4::CabacModels m;
	void start(BitWriter *bw) {
		bw_ = bw;
		low_ = 0;
		range_ = 510;
		first_ = true;
		outstanding_ = 0;
	}
	int bin(int ctx, int v) {
		v = v ? 1 : 0;
		uint8_t &s = m.st[ctx];
		const int p = s >> 1;
		int mps = s & 1;
		const uint32_t lps = b2::h4t::range_tab_lps[p][(range_ >> 6) & 3];
		range_ -= lps;
		if (v != mps) {
			low_ += range_;
			range_ = lps;
			if (p == 0) mps ^= 1;
			s = (uint8_t)((b2::h4t::trans_idx_lps[p] << 1) | mps);
		} else if (p < 62) {
			s = (uint8_t)(s + 2);
		}
		renorm();
		return v;
	}
	int bypass(int v) {
		v = v ? 1 : 0;
		low_ <<= 1;
		if (v) low_ += range_;
		if (low_ >= 1024) { put_bit(1); low_ -= 1024; }
		else if (low_ < 512) put_bit(0);
		else { low_ -= 512; outstanding_++; }
		return v;
	}
	// v = 1 also flushes (9.3.4.5): the last bit written is the rbsp_stop_one_bit at the end of
	// a slice, or precedes pcm_alignment_zero_bit for I_PCM
	int term(int v) {
		v = v ? 1 : 0;
		range_ -= 2;
		if (v) {
			low_ += range_;
			range_ = 2;
			renorm();
			put_bit((low_ >> 9) & 1);
			bw_->put(((low_ >> 7) & 3) | 1, 2);
		} else {
			renorm();
		}
		return v;
	}
private:
	void renorm() {
		while (range_ < 256) {
			if (low_ < 256) put_bit(0);
			else if (low_ >= 512) { low_ -= 512; put_bit(1); }
			else { low_ -= 256; outstanding_++; }
			range_ <<= 1;
			low_ <<= 1;
		}
	}
	void put_bit(int b) {
		if (first_) first_ = false;
		else bw_->put1(b);
		while (outstanding_ > 0) { bw_->put1(1 - b); outstanding_--; }
	}
	BitWriter *bw_ = nullptr;
	uint32_t low_ = 0, range_ = 510;
	bool first_ = true;
	int outstanding_ = 0;
};

}  // namespace sg
