// vlc.h — variable-length-code tables written as bit strings (as printed in the standards'
// annexes) and a two-level lookup decoder built from them at start-up.  The same string
// tables drive the test-only bitstream writer (tools/streamgen), so a typo shows up as a
// mismatch against the reference decoder rather than cancelling out silently... it would
// cancel out, which is why tests/ also decode writer output with the compiled reference.
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>
#include "bits.h"

namespace b2 {

struct VlcCode {
	const char *bits;   // '0'/'1' characters, spaces ignored
	int32_t value;
};

inline int vlc_parse(const char *s, uint32_t *code) {
	int len = 0;
	uint32_t c = 0;
	for (; *s; ++s) {
		if (*s == '0' || *s == '1') { c = (c << 1) | (uint32_t)(*s - '0'); ++len; }
	}
	*code = c;
	return len;
}

class Vlc {
public:
	enum { INVALID = INT32_MIN };
	Vlc() : l1_(0), maxlen_(0) {}
	void build(const VlcCode *codes, int n, int l1bits) {
		l1_ = l1bits;
		maxlen_ = 0;
		std::vector<uint32_t> code(n);
		std::vector<int> len(n);
		for (int i = 0; i < n; ++i) {
			len[i] = vlc_parse(codes[i].bits, &code[i]);
			if (len[i] > maxlen_) maxlen_ = len[i];
		}
		if (l1_ > maxlen_) l1_ = maxlen_;
		int l2 = maxlen_ - l1_;
		tab_.assign((size_t)1 << l1_, Entry{INVALID, 0});
		for (int i = 0; i < n; ++i) {
			if (len[i] <= l1_) {
				uint32_t base = code[i] << (l1_ - len[i]);
				for (uint32_t k = 0; k < (1u << (l1_ - len[i])); ++k)
					tab_[base + k] = Entry{codes[i].value, (int8_t)len[i]};
			} else {
				uint32_t prefix = code[i] >> (len[i] - l1_);
				if (tab_[prefix].len != -1) {
					// open a sub-table
					tab_[prefix] = Entry{(int32_t)tab_.size(), -1};
					tab_.resize(tab_.size() + ((size_t)1 << l2), Entry{INVALID, 0});
				}
				uint32_t off = (uint32_t)tab_[prefix].value;
				int rem = len[i] - l1_;
				uint32_t sub = code[i] & ((1u << rem) - 1);
				uint32_t base = sub << (l2 - rem);
				for (uint32_t k = 0; k < (1u << (l2 - rem)); ++k)
					tab_[off + base + k] = Entry{codes[i].value, (int8_t)len[i]};
			}
		}
	}
	// Returns the decoded value, or INVALID (no bits consumed) for an undefined code.
	inline __attribute__((always_inline)) int32_t decode(BitReader &br) const {
		uint32_t v = br.peek(maxlen_);
		const Entry *e = &tab_[v >> (maxlen_ - l1_)];
		if (e->len == -1) e = &tab_[(uint32_t)e->value + (v & ((1u << (maxlen_ - l1_)) - 1))];
		if (e->len <= 0) return INVALID;
		br.skip(e->len);
		return e->value;
	}
	// The same lookup on 32 bits of the stream the caller holds (nothing is consumed): the value and its code length,
	// or INVALID.
	inline __attribute__((always_inline)) int32_t lookup(uint32_t look32, int *len) const {
		const Entry *e = &tab_[look32 >> (32 - l1_)];
		if (e->len == -1) e = &tab_[(uint32_t)e->value + ((look32 >> (32 - maxlen_)) & ((1u << (maxlen_ - l1_)) - 1))];
		*len = e->len;
		return e->len <= 0 ? (int32_t)INVALID : e->value;
	}
	int maxlen() const { return maxlen_; }
private:
	struct Entry { int32_t value; int8_t len; };
	std::vector<Entry> tab_;
	int l1_, maxlen_;
};

// Compact two-level table for codes of at most 16 bits and values 0..255 (the H.264 CAVLC tables): 16-bit
// entries, so that all tables of the residual parse together stay a few KB of L1.  Entry: value in bits 0..7,
// code length in bits 8..12 (0 = undefined code), bit 15 = link to a second-level table (value = its number).
class VlcSmall {
public:
	enum { INVALID = -1 };
	VlcSmall() : l1_(0), l2_(0) {}
	void build(const VlcCode *codes, int n, int l1bits) {
		int maxlen = 0;
		std::vector<uint32_t> code(n);
		std::vector<int> len(n);
		for (int i = 0; i < n; ++i) {
			len[i] = vlc_parse(codes[i].bits, &code[i]);
			if (len[i] > maxlen) maxlen = len[i];
		}
		l1_ = l1bits < maxlen ? l1bits : maxlen;
		l2_ = maxlen - l1_;
		tab_.assign((size_t)1 << l1_, 0);
		int n_sub = 0;
		for (int i = 0; i < n; ++i) {
			const uint16_t leaf = (uint16_t)((len[i] << 8) | (codes[i].value & 255));
			if (len[i] <= l1_) {
				const uint32_t base = code[i] << (l1_ - len[i]);
				for (uint32_t k = 0; k < (1u << (l1_ - len[i])); ++k) tab_[base + k] = leaf;
			} else {
				const uint32_t prefix = code[i] >> (len[i] - l1_);
				if (!(tab_[prefix] & 0x8000)) {
					tab_[prefix] = (uint16_t)(0x8000 | n_sub++);
					tab_.resize(tab_.size() + ((size_t)1 << l2_), 0);
				}
				const uint32_t off = ((uint32_t)1 << l1_) + ((uint32_t)(tab_[prefix] & 255) << l2_);
				const int rem = len[i] - l1_;
				const uint32_t base = (code[i] & ((1u << rem) - 1)) << (l2_ - rem);
				for (uint32_t k = 0; k < (1u << (l2_ - rem)); ++k) tab_[off + base + k] = leaf;
			}
		}
	}
	// look16 = the next 16 bits of the stream.  Returns the value and sets *len, or INVALID (undefined code).
	inline __attribute__((always_inline)) int lookup(uint32_t look16, int *len) const {
		uint32_t e = tab_[look16 >> (16 - l1_)];
		if (e & 0x8000) e = tab_[(1u << l1_) + ((e & 255) << l2_) + ((look16 >> (16 - l1_ - l2_)) & ((1u << l2_) - 1))];
		*len = (int)(e >> 8);
		return e ? (int)(e & 255) : (int)INVALID;
	}
private:
	std::vector<uint16_t> tab_;
	int l1_, l2_;
};

}  // namespace b2
