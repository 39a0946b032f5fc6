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

// The application's buffer [cur, tail) is only valid until the next refill() (the reference's
// bitio.c keeps a few bytes of its own and the reference application refills one and the same
// buffer, src/app/m2decoder.cpp), so every byte that is still needed is copied into `pending`
// BEFORE refill() is called.  `pending` holds, while unit_open, the bytes of the unit read so far
// (they follow a start code); otherwise the last (at most two) bytes seen, which may be the
// beginning of a start code that straddles two buffers.

// A start code whose first byte(s) are the last bytes of `pd` and whose 01 lies in [cur, cur + avail):
// returns the number of its bytes inside pd (1 or 2) and sets *used to the number inside the buffer;
// 0 if there is none.
static int straddling_start_code(const std::vector<uint8_t> &pd, const uint8_t *cur, long avail, int *used) {
	int pn = pd.size() < 2 ? (int)pd.size() : 2;
	int bn = avail < 2 ? (int)avail : 2;
	uint8_t t[4];
	for (int i = 0; i < pn; ++i) t[i] = pd[pd.size() - pn + i];
	for (int i = 0; i < bn; ++i) t[pn + i] = cur[i];
	for (int i = 0; i < pn && i + 2 < pn + bn; ++i) {
		if (t[i] == 0 && t[i + 1] == 0 && t[i + 2] == 1) {
			*used = i + 3 - pn;
			return pn - i;
		}
	}
	return 0;
}

// refill() and report whether bytes arrived
static bool more_data(dec_bits *s) {
	s->cur = s->tail;
	if (s->refill(s->refill_arg) < 0) return false;
	return s->cur && s->cur < s->tail;
}

int next_unit(dec_bits *s, Unit *u) {
	if (!s->pending) s->pending = new std::vector<uint8_t>();
	std::vector<uint8_t> &pd = *s->pending;
	if (s->pending_returned) {       // the unit handed out by the previous call lived in `pending`
		pd.clear();
		s->pending_returned = false;
	}
	// 1. locate the start code that opens the unit
	while (!s->unit_open) {
		long avail = s->cur ? (long)(s->tail - s->cur) : 0;
		int used = 0;
		if (avail > 0 && straddling_start_code(pd, s->cur, avail, &used)) {
			s->cur += used;
			pd.clear();
			s->unit_open = true;
			break;
		}
		long pos = avail >= 3 ? find_start_code(s->cur, avail) : -1;
		if (pos >= 0) {
			s->cur += pos;
			pd.clear();
			s->unit_open = true;
			break;
		}
		// keep the last two bytes seen: they may begin a start code that ends in the next buffer
		for (long i = avail > 2 ? avail - 2 : 0; i < avail; ++i) pd.push_back(s->cur[i]);
		if (pd.size() > 2) pd.erase(pd.begin(), pd.end() - 2);
		if (!more_data(s)) return -1;
	}
	// 2. find where it ends: the unit is `pending` followed by [cur, ...)
	for (;;) {
		long avail = s->cur ? (long)(s->tail - s->cur) : 0;
		int used = 0, inside = avail > 0 ? straddling_start_code(pd, s->cur, avail, &used) : 0;
		if (inside) {
			// the next start code began in the previous buffer: its zeros are not part of this unit
			pd.resize(pd.size() - (size_t)inside);
			s->cur += used;
			u->data = pd.data();
			u->len = pd.size();
			s->pending_returned = true;
			return 0;            // unit_open stays set: cur is just behind the next start code
		}
		long nxt = avail >= 3 ? find_start_code(s->cur, avail) : -1;
		if (nxt >= 0) {
			if (pd.empty()) {
				u->data = s->cur;    // whole unit inside the application's buffer: no copy
				u->len = (size_t)(nxt - 3);
			} else {
				pd.insert(pd.end(), s->cur, s->cur + (nxt - 3));
				u->data = pd.data();
				u->len = pd.size();
				s->pending_returned = true;
			}
			s->cur += nxt;
			return 0;
		}
		// the unit runs to the end of this buffer: take the bytes along, then ask for more
		if (avail > 0) pd.insert(pd.end(), s->cur, s->tail);
		if (!more_data(s)) {
			// end of data: the unit ends here
			s->unit_open = false;
			u->data = pd.data();
			u->len = pd.size();
			s->pending_returned = true;
			return 0;
		}
	}
}

}  // namespace b2
