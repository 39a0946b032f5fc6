// e2ebench — end-to-end decode through the plugin table (m2d_func_table_t) from a native multi-threaded host:
// elementary streams in host memory -> parse -> H2D of the records -> kernels -> D2H of every output frame into
// page-locked application frame buffers.  One decoder context per stream, T host threads, the decode loop of the
// reference application (src/app/m2decoder.h:132-157, h264dec.cpp:251-257); a thread keeps two contexts and collects the
// last frames of a finished stream after the first pictures of its next one (start_stream / step_stream /
// finish_stream below), as a server multiplexing streams would.  Links libm2dec_b200.so; the Python
// bench drives it and checks its frame hashes against the frames the (MD5-verified) Python path produced.
//   e2ebench <h264|mpeg2> <threads> <repetitions> <cuda device> <md5 0|1|2> file...
// md5 = 2: digest mode (b2_context_set_md5): the frames are hashed ON THE DEVICE and only their 16-byte MD5s come
// back; a frame's digest is collected DIGEST_LAG output frames after the frame came out (a hash takes ~30 ms, the
// decoder does not reuse a frame buffer that soon: on_header hands it 16 frames more than it asks for).
// Output: one line "frames F seconds S h2d B d2h B launches L" (seconds = sum over the repetitions; per-repetition
// times go to stderr), then with md5=1 one line per stream: "stream i n m0 m1 ..." = the MD5 of every output frame
// as `h264dec -O` of the reference prints it (cropped luma rows, then cropped interleaved-CbCr rows;
// src/app/filewrite.h:11-29, 99-124).
#include <malloc.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "m2d_plugin.h"
#include "b2_recon.h"

// MD5 (RFC 1321), for the verification pass only
struct Md5 {
	uint32_t h[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
	uint64_t len = 0;
	uint8_t buf[64];
	size_t fill = 0;
	static uint32_t rol(uint32_t x, int c) { return (x << c) | (x >> (32 - c)); }
	void block(const uint8_t *p) {
		static const uint32_t K[64] = {
			0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
			0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
			0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
			0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
		static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
		                          4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
		uint32_t w[16];
		memcpy(w, p, 64);
		uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
		for (int i = 0; i < 64; ++i) {
			uint32_t f;
			int g;
			if (i < 16) { f = (b & c) | (~b & d); g = i; }
			else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
			else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
			else { f = c ^ (b | ~d); g = (7 * i) & 15; }
			const uint32_t t = d;
			d = c;
			c = b;
			b = b + rol(a + f + K[i] + w[g], S[i]);
			a = t;
		}
		h[0] += a; h[1] += b; h[2] += c; h[3] += d;
	}
	void update(const uint8_t *p, size_t n) {
		len += n;
		if (fill) {
			const size_t k = n < 64 - fill ? n : 64 - fill;
			memcpy(buf + fill, p, k);
			fill += k; p += k; n -= k;
			if (fill == 64) { block(buf); fill = 0; }
		}
		for (; n >= 64; n -= 64, p += 64) block(p);
		if (n) { memcpy(buf, p, n); fill = n; }
	}
	std::string hex() {
		const uint64_t bits = len * 8;
		const uint8_t pad = 0x80, zero = 0;
		update(&pad, 1);
		while (fill != 56) update(&zero, 1);
		uint8_t l[8];
		memcpy(l, &bits, 8);
		update(l, 8);
		char out[33];
		for (int i = 0; i < 16; ++i) snprintf(out + 2 * i, 3, "%02x", (unsigned)((h[i >> 2] >> (8 * (i & 3))) & 255));
		return std::string(out, 32);
	}
};

struct Stream {
	std::vector<uint8_t> data;
	std::vector<std::string> hashes;
	uint64_t stats[3] = {0, 0, 0};      // bytes host -> device, device -> host, kernel launches (of the last decode)
};

#ifdef E2EBENCH_SAMPLE
// SIGPROF sampler (no perf in the image): program counters + /proc/self/maps to e2e_pcs.txt / e2e_maps.txt;
// tools/e2ebench/report.py buckets them by mapping and, inside libm2dec_b200.so, by function
#include <signal.h>
#include <sys/time.h>
#include <ucontext.h>
static uint64_t g_pcs[1 << 21];
static std::atomic<int> g_npcs(0);
static void on_prof(int, siginfo_t *, void *uc) {
	int i = g_npcs.fetch_add(1);
	if (i < (1 << 21)) g_pcs[i] = (uint64_t)((ucontext_t *)uc)->uc_mcontext.gregs[REG_RIP];
}
static void sample_start() {
	struct sigaction sa;
	memset(&sa, 0, sizeof(sa));
	sa.sa_sigaction = on_prof;
	sa.sa_flags = SA_SIGINFO | SA_RESTART;
	sigaction(SIGPROF, &sa, nullptr);
	struct itimerval tv = {{0, 1000}, {0, 1000}};
	setitimer(ITIMER_PROF, &tv, nullptr);
}
static void sample_stop() {
	struct itimerval tv = {{0, 0}, {0, 0}};
	setitimer(ITIMER_PROF, &tv, nullptr);
	FILE *f = fopen("e2e_pcs.txt", "w");
	int n = g_npcs.load();
	if (n > (1 << 21)) n = 1 << 21;
	for (int i = 0; i < n; ++i) fprintf(f, "%llx\n", (unsigned long long)g_pcs[i]);
	fclose(f);
	FILE *m = fopen("/proc/self/maps", "r"), *o = fopen("e2e_maps.txt", "w");
	char line[1024];
	while (fgets(line, sizeof(line), m)) fputs(line, o);
	fclose(m);
	fclose(o);
}
#else
static void sample_start() {}
static void sample_stop() {}
#endif

struct FramePool {          // the application's frame buffers of one thread, reused from stream to stream
	uint8_t *block = nullptr;
	size_t bytes = 0;
	~FramePool() { if (block) b2_host_free(block); }
};

struct App {
	const m2d_func_table_t *f;
	bool h264;
	int device;
	bool hash;
	std::vector<uint8_t> ctx;
	const Stream *in = nullptr;
	Stream *out = nullptr;
	bool fed = false;
	FramePool *pool;
	std::vector<m2d_frame_t> frames;
	int frames_w = 0, frames_h = 0;
	long n_frames = 0;
	bool failed = false;
	bool digest = false;                 // md5 = 2
	std::deque<m2d_frame_t> pending;     // frames whose digest has not been collected yet
};
#ifndef B2_DIGEST_LAG
#define B2_DIGEST_LAG 10
#endif
enum { DIGEST_LAG = B2_DIGEST_LAG, OVERLAP_PICTURES = 3 };

static void collect_digest(App *a) {
	uint8_t d[16];
	if (b2_context_frame_md5(a->ctx.data(), &a->pending.front(), d) < 0) { a->failed = true; a->pending.pop_front(); return; }
	a->pending.pop_front();
	char out[33];
	for (int i = 0; i < 16; ++i) snprintf(out + 2 * i, 3, "%02x", d[i]);
	a->out->hashes.push_back(std::string(out, 32));
}

static int on_refill(void *arg) {
	App *a = (App *)arg;
	if (a->fed) return -1;
	a->fed = true;
	dec_bits_set_data(a->f->stream_pos(a->ctx.data()), a->in->data.data(), a->in->data.size(), nullptr);
	return 0;
}

static int on_header(void *arg, void *) {        // reference m2decoder.h:54-80
	App *a = (App *)arg;
	m2d_info_t info;
	a->f->get_info(a->ctx.data(), &info);
	const int w = (info.src_width + 15) & ~15, h = (info.src_height + 15) & ~15;
	int n = info.frame_num + (a->h264 ? 16 : 0);
	if (n > (a->h264 ? 64 : 16)) n = a->h264 ? 64 : 16;
	if (w == a->frames_w && h == a->frames_h && (int)a->frames.size() >= n) return 0;
	const size_t fb = (size_t)w * h * 3 / 2;
	if (a->pool->bytes < fb * n) {
		if (a->pool->block) b2_host_free(a->pool->block);
		a->pool->block = (uint8_t *)b2_host_alloc(fb * n, a->device);
		a->pool->bytes = a->pool->block ? fb * n : 0;
		if (!a->pool->block) { a->failed = true; return -1; }
	}
	a->frames.assign(n, m2d_frame_t());
	for (int i = 0; i < n; ++i) {
		a->frames[i].luma = a->pool->block + fb * i;
		a->frames[i].chroma = a->frames[i].luma + (size_t)w * h;
	}
	a->frames_w = w;
	a->frames_h = h;
	if (a->f->set_frames(a->ctx.data(), n, a->frames.data(), nullptr, 0) < 0) { a->failed = true; return -1; }
	return 0;
}

static void post(App *a, const m2d_frame_t &fr) {
	a->n_frames++;
	if (a->digest) {
		a->pending.push_back(fr);
		while (a->pending.size() > DIGEST_LAG) collect_digest(a);
		return;
	}
	if (!a->hash) return;
	const int stride = fr.width, width = stride - fr.crop[0] - fr.crop[1], height = fr.height - fr.crop[2] - fr.crop[3];
	Md5 m;
	auto rows = [&](const uint8_t *p, int first, int n) {
		for (int y = first; y < first + n; ++y) m.update(p + (size_t)y * stride + fr.crop[0], (size_t)width);
	};
	rows(fr.luma, fr.crop[2], height);
	rows(fr.chroma, fr.crop[2] >> 1, height >> 1);
	a->out->hashes.push_back(m.hex());
}

static int decode(App *a) {                       // reference m2decoder.h:132-157 (the flush at the end of data: finish_stream)
	void *c = a->ctx.data();
	m2d_frame_t fr;
	while (a->f->peek_decoded_frame(c, &fr, 0) <= 0) {
		int err = a->f->decode_picture(c);
		if (err < 0) return err;
	}
	a->f->get_decoded_frame(c, &fr, 0);
	post(a, fr);
	return a->f->decode_picture(c);
}

// A stream in three parts, so that a thread can have the end of one stream and the start of the next in flight at
// the same time: the frames still inside the decoder when the data ends (the last pictures' copies need the device
// another ~2 ms) are collected after the next stream's first pictures have been parsed, not waited for.
static bool start_stream(App *a, const Stream *in, Stream *out) {
	a->in = in;
	a->out = out;
	a->fed = false;
	a->failed = false;
	a->frames_w = a->frames_h = 0;
	a->frames.clear();
	if (a->hash || a->digest) out->hashes.clear();
	a->pending.clear();
	void *c = a->ctx.data();
	if (a->f->init(c, -1, on_header, a) < 0) return false;
	b2_context_set_device(c, a->device);
	if (a->digest) b2_context_set_md5(c, 1);
	dec_bits_set_callback(a->f->stream_pos(c), on_refill, a);
	return true;
}
// one step of the decode loop; false once the data has ended (or the stream failed)
static bool step_stream(App *a) {
	int err = decode(a);
	if (a->failed || err == -4) { a->failed = true; return false; }
	return err >= 0;
}
static bool finish_stream(App *a) {
	void *c = a->ctx.data();
	m2d_frame_t fr;
	if (!a->failed)
		while (a->f->peek_decoded_frame(c, &fr, 1) > 0) { post(a, fr); a->f->get_decoded_frame(c, &fr, 1); }
	while (!a->failed && !a->pending.empty()) collect_digest(a);
	if (!a->failed) b2_context_stats(c, a->out->stats);
	b2_context_release(c);
	return !a->failed;
}

int main(int argc, char **argv) {
	// A decoder context allocates a few MB of tables per stream; above glibc's mmap threshold every stream start maps and
	// every release unmaps them, and 16 threads then queue for the address-space lock and fault the pages in again
	// (+2 % end to end with the heap kept).
	mallopt(M_MMAP_THRESHOLD, 1 << 30);
	mallopt(M_TRIM_THRESHOLD, 1 << 30);
	if (argc < 7) { fprintf(stderr, "usage: e2ebench h264|mpeg2 threads repetitions device hash file...\n"); return 2; }
	const bool h264 = !strcmp(argv[1], "h264");
	const int threads = atoi(argv[2]), reps = atoi(argv[3]), device = atoi(argv[4]);
	const bool hash = atoi(argv[5]) == 1, digest = atoi(argv[5]) == 2;
	std::vector<Stream> streams(argc - 6);
	for (int i = 6; i < argc; ++i) {
		FILE *fp = fopen(argv[i], "rb");
		if (!fp) { perror(argv[i]); return 1; }
		fseek(fp, 0, SEEK_END);
		streams[i - 6].data.resize(ftell(fp));
		fseek(fp, 0, SEEK_SET);
		if (fread(streams[i - 6].data.data(), 1, streams[i - 6].data.size(), fp) != streams[i - 6].data.size()) return 1;
		fclose(fp);
	}
	if (b2_device_count() <= 0) { fprintf(stderr, "e2ebench: no CUDA device: %s\n", b2_last_error()); return 3; }
	const m2d_func_table_t *f = h264 ? h264d_gpu_func : m2d_gpu_func;
	std::vector<FramePool> pools(2 * threads);
	std::atomic<long> frames(0);
	std::atomic<bool> ok(true);
	double seconds = 0;
	for (int rep = 0; rep < reps; ++rep) {
		if (rep == 1) sample_start();       // (repetition 0 is the warm-up)
		std::atomic<int> next(0);
		auto t0 = std::chrono::steady_clock::now();
		std::vector<double> busy(threads, 0.0);       // when each thread ran out of streams
		std::vector<std::thread> ths;
		for (int t = 0; t < threads; ++t)
			ths.emplace_back([&, t]() {
				App slot[2];
				for (int k = 0; k < 2; ++k) {
					App &a = slot[k];
					a.f = f;
					a.h264 = h264;
					a.device = device;
					a.hash = hash;
					a.digest = digest;
					a.ctx.resize(f->context_size);
					a.pool = &pools[2 * t + k];
				}
				int cur = 0;
				bool waiting = false;            // slot[1 - cur] has reached the end of its data and awaits finish_stream
				for (;;) {
					const int i = next.fetch_add(1);
					if (i >= (int)streams.size()) break;
					App &a = slot[cur];
					if (!start_stream(&a, &streams[i], &streams[i])) { ok = false; continue; }
					int steps = 0;
					while (step_stream(&a))
						if (waiting && ++steps == OVERLAP_PICTURES) { if (!finish_stream(&slot[1 - cur])) ok = false; waiting = false; }
					if (waiting && !finish_stream(&slot[1 - cur])) ok = false;
					waiting = true;
					cur = 1 - cur;
				}
				if (waiting && !finish_stream(&slot[1 - cur])) ok = false;
				frames += slot[0].n_frames + slot[1].n_frames;
				busy[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			});
		for (auto &th : ths) th.join();
		const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		double sum = 0;
		for (double b : busy) sum += b;
		// (second figure: how much of the repetition the average thread had work for -- the rest is the tail)
		fprintf(stderr, "e2ebench threads busy %.0f %% of repetition %d\n", 100.0 * sum / (threads * dt), rep);
		fprintf(stderr, "e2ebench repetition %d: %.3f s\n", rep, dt);
		seconds += dt;
	}
	sample_stop();
	if (!ok) { fprintf(stderr, "e2ebench: a stream failed: %s\n", b2_last_error()); return 4; }
	uint64_t tot[3] = {0, 0, 0};
	for (const Stream &st : streams)
		for (int k = 0; k < 3; ++k) tot[k] += st.stats[k];
	// (traffic and launches of ONE pass over the streams)
	printf("frames %ld seconds %.6f h2d %llu d2h %llu launches %llu\n", frames.load(), seconds, (unsigned long long)tot[0], (unsigned long long)tot[1], (unsigned long long)tot[2]);
	if (hash || digest)
		for (size_t i = 0; i < streams.size(); ++i) {
			printf("stream %zu %zu", i, streams[i].hashes.size());
			for (const std::string &h : streams[i].hashes) printf(" %s", h.c_str());
			printf("\n");
		}
	return 0;
}
