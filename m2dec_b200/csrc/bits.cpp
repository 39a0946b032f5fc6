// bits.cpp — dec_bits feeder entry points and the start-code unit assembler.
// Interface mirrors reference src/lib/bitio.h:57-61 (dec_bits_open/close/set_callback/
// set_data/current/tail) and src/lib/m2d.h:80 (m2d_next_start_code); implementation is ours.
#include "bits.h"
#include <string.h>

static int no_more_data(void *) { return -1; }

extern "C" int dec_bits_open(dec_bits *s, void (*)(dec_bits *, int)) {
	if (!s) return -1;
	memset(s, 0, sizeof(*s));
	s->refill = no_more_data;
	return 0;
}

extern "C" void dec_bits_close(dec_bits *s) {
	if (!s) return;
	delete s->pending;
	memset(s, 0, sizeof(*s));
}

extern "C" void dec_bits_set_callback(dec_bits *s, int (*refill)(void *), void *arg) {
	s->refill = refill ? refill : no_more_data;
	if (arg) s->refill_arg = arg;
}

extern "C" int dec_bits_set_data(dec_bits *s, const byte_t *buf, size_t len, void *id) {
	if (!s || !buf || (long)len < 0) return -1;
	s->cur = s->head = buf;
	s->tail = buf + len;
	s->id = id;
	return 0;
}

extern "C" const byte_t *dec_bits_current(dec_bits *s) { return s->cur; }
extern "C" const byte_t *dec_bits_tail(dec_bits *s) { return s->tail; }

// Offset just past the first 00 00 01 in [p, p+len), or -1.
static long find_start_code(const uint8_t *p, long len) {
	long i = 2;
	while (i < len) {
		uint8_t c = p[i];
		if (c > 1) { i += 3; continue; }        // this byte cannot be the 2nd zero or the 01
		if (c == 1) {
			if (p[i - 1] == 0 && p[i - 2] == 0) return i + 1;
			i += 3;
		} else {
			i += 1;
		}
	}
	return -1;
}

extern "C" int m2d_next_start_code(const byte_t *src, int len) {
	long r = find_start_code(src, len);
	return r < 0 ? -1 : (int)r;
}

namespace b2 {

int next_unit(dec_bits *s, Unit *u) {
	// 1. locate the start code that opens the unit
	for (;;) {
		long avail = s->cur ? (long)(s->tail - s->cur) : 0;
		long pos = avail >= 3 ? find_start_code(s->cur, avail) : -1;
		if (pos >= 0) { s->cur += pos; break; }
		// keep nothing: a start code split across application buffers before any unit
		// began is not supported (the reference handles it with prev_[], bitio.c:117-120)
		if (s->cur) s->cur = s->tail;
		if (s->refill(s->refill_arg) < 0) return -1;
	}
	// 2. find where it ends
	long avail = (long)(s->tail - s->cur);
	long nxt = find_start_code(s->cur, avail);
	if (nxt >= 0) {
		u->data = s->cur;
		u->len = (size_t)(nxt - 3);
		s->cur += nxt - 3;
		return 0;
	}
	// unit runs to the end of this buffer: ask for more; if none, the unit ends here
	const uint8_t *start = s->cur;
	size_t len = (size_t)avail;
	const uint8_t *old_tail = s->tail;
	s->cur = s->tail;
	if (s->refill(s->refill_arg) < 0 || s->tail == old_tail) {
		u->data = start;
		u->len = len;
		return 0;
	}
	if (!s->pending) s->pending = new std::vector<uint8_t>();
	std::vector<uint8_t> &pd = *s->pending;
	pd.assign(start, start + len);
	for (;;) {
		size_t from = pd.size() >= 2 ? pd.size() - 2 : 0;
		pd.insert(pd.end(), s->cur, s->tail);
		long q = find_start_code(pd.data() + from, (long)(pd.size() - from));
		if (q >= 0) {
			size_t unit_len = from + (size_t)q - 3;
			size_t consumed_new = unit_len > (pd.size() - (size_t)(s->tail - s->cur))
				? unit_len - (pd.size() - (size_t)(s->tail - s->cur)) : 0;
			s->cur += consumed_new;
			pd.resize(unit_len);
			break;
		}
		s->cur = s->tail;
		if (s->refill(s->refill_arg) < 0) break;
	}
	u->data = pd.data();
	u->len = pd.size();
	return 0;
}

}  // namespace b2
