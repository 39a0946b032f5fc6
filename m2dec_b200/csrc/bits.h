// bits.h — byte-stream feeder (`dec_bits`), start-code unit assembler and the bit reader
// used by both front ends.
//
// Role in the reference: src/lib/bitio.{c,h} + m2d_find_mpeg_data (src/lib/m2d.cpp:130-155).
// Design here is different on purpose: the reference reads through a refillable cache and
// unwinds with longjmp at end of data; we first cut the stream into start-code delimited
// units (NAL units / MPEG-2 start-code payloads), then parse each unit from contiguous
// memory with a bounds-checked reader that latches an `overrun` flag instead of jumping.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#include "m2d_plugin.h"

struct dec_bits_t {
	const uint8_t *cur;       // next unread byte of the current buffer
	const uint8_t *head;
	const uint8_t *tail;
	int (*refill)(void *);    // returns <0 when the application has no more data
	void *refill_arg;
	void *id;
	std::vector<uint8_t> *pending;  // bytes carried over a refill (see next_unit, bits.cpp)
	bool unit_open;           // cur is inside a unit (behind its start code)
	bool pending_returned;    // the last unit handed out lives in `pending`
};

namespace b2 {

struct Unit {
	const uint8_t *data;   // first byte after the 00 00 01 prefix (start-code value / NAL header)
	size_t len;
};

// Fetch the next start-code delimited unit.  Returns 0 and fills `u`, or <0 at end of data.
int next_unit(dec_bits *s, Unit *u);

class BitReader {
public:
	BitReader() : p_(nullptr), end_(nullptr), cache_(0), n_(0), fake_(0) {}
	void init(const uint8_t *data, size_t len) {
		p_ = data; end_ = data + len; cache_ = 0; n_ = 0; fake_ = 0;
	}
	// look at the next n (1..32) bits, MSB first
	inline __attribute__((always_inline)) uint32_t peek(int n) {
		if (n_ < n) fill();
		return (uint32_t)(cache_ >> (64 - n));
	}
	inline __attribute__((always_inline)) void skip(int n) {
		if (n_ < n) fill();
		cache_ <<= n;
		n_ -= n;
	}
	inline __attribute__((always_inline)) uint32_t get(int n) {
		if (n == 0) return 0;
		uint32_t v = peek(n);
		cache_ <<= n;
		n_ -= n;
		return v;
	}
	inline __attribute__((always_inline)) uint32_t get1() {
		if (n_ < 1) fill();
		uint32_t v = (uint32_t)(cache_ >> 63);
		cache_ <<= 1;
		n_ -= 1;
		return v;
	}
	inline __attribute__((always_inline)) uint32_t ue() {
		uint32_t look = peek(32);
		if (look & 0x80000000u) { skip(1); return 0; }
		if (!look) { skip(32); return 0; }
		int zeros = __builtin_clz(look);
		skip(zeros + 1);
		return (1u << zeros) - 1 + get(zeros);
	}
	inline __attribute__((always_inline)) int32_t se() {
		uint32_t k = ue();
		int32_t t = (int32_t)((k + 1) >> 1);
		return (k & 1) ? t : -t;
	}
	inline int bit_misalign() const { return n_ & 7; }
	inline void byte_align() { int d = n_ & 7; cache_ <<= d; n_ -= d; }
	// bits still unread; negative once the parse has consumed zeros fed past the end
	inline long bits_left() const { return (long)(end_ - p_) * 8 + n_ - (long)fake_ * 8; }
	inline bool overrun() const { return bits_left() < 0; }
	const uint8_t *end_ptr() const { return end_; }
	// byte pointer of the current position (call when byte aligned)
	const uint8_t *byte_ptr() const {
		long back = (n_ >> 3) - fake_;
		const uint8_t *q = p_ - (back > 0 ? back : 0);
		return q;
	}
	void skip_bytes(size_t nbytes) {
		const uint8_t *q = byte_ptr();
		size_t avail = (size_t)(end_ - q);
		if (nbytes > avail) { fake_ += (int)(nbytes - avail) + 1; nbytes = avail; }
		p_ = q + nbytes; cache_ = 0; n_ = 0;
		if (fake_) { n_ = 0; }
	}
	// true when only zero bits (padding before the next start code) remain
	bool only_zero_left() {
		if (bits_left() <= 0) return true;
		if (n_ > 0 && (cache_ >> (64 - n_)) != 0) return false;
		for (const uint8_t *q = p_; q < end_; ++q)
			if (*q) return false;
		return true;
	}
private:
	inline void fill() {
		if (n_ <= 32 && p_ + 4 <= end_) {
			// common case: four bytes at once
			uint32_t v;
			memcpy(&v, p_, 4);
			cache_ |= (uint64_t)__builtin_bswap32(v) << (32 - n_);
			p_ += 4;
			n_ += 32;
			return;
		}
		while (n_ <= 56) {
			uint64_t b = 0;
			if (p_ < end_) b = *p_++; else fake_++;   // past the unit: feed zeros, count them
			cache_ |= b << (56 - n_);
			n_ += 8;
		}
	}
	const uint8_t *p_;
	const uint8_t *end_;
	uint64_t cache_;
	int n_;
	int fake_;
};

}  // namespace b2
