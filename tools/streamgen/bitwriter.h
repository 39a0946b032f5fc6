// bitwriter.h — test-only: MSB-first bit writer, Exp-Golomb codes, and VLC encoding from the
// same string tables the decoder builds its lookups from (m2dec_b200/csrc/vlc.h).
#pragma once
#include <stdint.h>
#include <map>
#include <vector>
#include "vlc.h"

namespace sg {

class BitWriter {
public:
	BitWriter() : acc_(0), n_(0) {}
	void put(uint32_t v, int bits) {
		for (int i = bits - 1; i >= 0; --i) {
			acc_ = (acc_ << 1) | ((v >> i) & 1);
			if (++n_ == 8) { buf_.push_back((uint8_t)acc_); acc_ = 0; n_ = 0; }
		}
	}
	void put1(int b) { put((uint32_t)(b & 1), 1); }
	void ue(uint32_t v) {
		uint64_t x = (uint64_t)v + 1;
		int len = 0;
		while ((x >> len) > 1) len++;
		put(0, len);
		for (int i = len; i >= 0; --i) put1((int)((x >> i) & 1));
	}
	void se(int32_t v) { ue(v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * (int64_t)v)); }
	void code(const char *bits) {
		for (; *bits; ++bits)
			if (*bits == '0' || *bits == '1') put1(*bits - '0');
	}
	bool aligned() const { return n_ == 0; }
	void align_zero() { while (n_) put1(0); }
	void rbsp_trailing() { put1(1); align_zero(); }
	size_t bit_count() const { return buf_.size() * 8 + n_; }
	std::vector<uint8_t> &bytes() { return buf_; }
private:
	std::vector<uint8_t> buf_;
	uint32_t acc_;
	int n_;
};

// value -> code string
class VlcEnc {
public:
	void build(const b2::VlcCode *codes, int n) {
		for (int i = 0; i < n; ++i) map_[codes[i].value] = codes[i].bits;
	}
	bool has(int v) const { return map_.count(v) != 0; }
	void put(BitWriter &bw, int v) const { bw.code(map_.at(v)); }
private:
	std::map<int, const char *> map_;
};

// deterministic, platform independent generator (xorshift64*)
class Rng {
public:
	explicit Rng(uint64_t seed) : s_(seed * 0x9E3779B97F4A7C15ULL + 0x1234567ULL) { next(); next(); }
	uint64_t next() {
		s_ ^= s_ >> 12; s_ ^= s_ << 25; s_ ^= s_ >> 27;
		return s_ * 0x2545F4914F6CDD1DULL;
	}
	// uniform integer in [lo, hi]
	int range(int lo, int hi) { return lo + (int)(next() % (uint64_t)(hi - lo + 1)); }
	bool chance(int percent) { return (int)(next() % 100) < percent; }
private:
	uint64_t s_;
};

}  // namespace sg
