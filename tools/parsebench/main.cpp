// parsebench — host front end alone (no device, no Python): parses an elementary stream through the plugin
// table with the device switched off and reports milliseconds per picture.  Development tool for the host
// parse (the end-to-end bottleneck); links the front-end sources directly, with stubs for the device session.
//   parsebench h264|mpeg2 file [repetitions]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>
#include "m2d_plugin.h"
#include "b2_recon.h"

// ---- stubs of the device-side C ABI (never reached: device_off) ----------------------------------------
extern "C" {
int b2_session_create(int, int, int, int, int, b2_session **) { return -1; }
void b2_session_destroy(b2_session *) {}
int b2_session_submit_mpeg2(b2_session *, const b2_m2_picture *) { return -1; }
int b2_session_submit_h264(b2_session *, const b2_h264_picture *) { return -1; }
int b2_session_bind_host_frame(b2_session *, int, uint8_t *, uint8_t *) { return -1; }
int b2_session_download(b2_session *, int, uint8_t *, uint8_t *) { return -1; }
int b2_session_traffic(const b2_session *, uint64_t *) { return -1; }
int b2_session_set_record_check(b2_session *, int) { return -1; }
int b2_session_set_md5(b2_session *, int, int, int, int, int, int) { return -1; }
int b2_session_md5(b2_session *, int, uint8_t *) { return -1; }
uint64_t b2_session_launches(const b2_session *) { return 0; }
const char *b2_last_error(void) { return ""; }
}
namespace b2 { void set_error(const char *, ...) {} }

#ifdef PARSEBENCH_SAMPLE
// poor man's sampling profiler (no perf in the image): SIGPROF every 250 us, program counters to pcs.txt;
// tools/parsebench/report.py turns them into per-function / per-line shares with addr2line
#include <signal.h>
#include <sys/time.h>
#include <ucontext.h>
static uint64_t g_pcs[1 << 20];
static volatile int g_npcs = 0;
static void on_prof(int, siginfo_t *, void *uc) {
	if (g_npcs < (1 << 20)) g_pcs[g_npcs++] = (uint64_t)((ucontext_t *)uc)->uc_mcontext.gregs[REG_RIP];
}
static void sample_start() {
	struct sigaction sa;
	memset(&sa, 0, sizeof(sa));
	sa.sa_sigaction = on_prof;
	sa.sa_flags = SA_SIGINFO | SA_RESTART;
	sigaction(SIGPROF, &sa, nullptr);
	struct itimerval tv = {{0, 250}, {0, 250}};
	setitimer(ITIMER_PROF, &tv, nullptr);
}
static void sample_stop() {
	struct itimerval tv = {{0, 0}, {0, 0}};
	setitimer(ITIMER_PROF, &tv, nullptr);
	FILE *f = fopen("pcs.txt", "w");
	for (int i = 0; i < g_npcs; ++i) fprintf(f, "%llx\n", (unsigned long long)g_pcs[i]);
	fclose(f);
}
#else
static void sample_start() {}
static void sample_stop() {}
#endif

struct App {
	const m2d_func_table_t *f;
	void *ctx;
	std::vector<uint8_t> data;
	bool fed = false;
	std::vector<std::vector<uint8_t>> bufs;
	std::vector<m2d_frame_t> frames;
	long pictures = 0;
};

static int on_refill(void *arg) {
	App *a = (App *)arg;
	if (a->fed) return -1;
	a->fed = true;
	dec_bits_set_data(a->f->stream_pos(a->ctx), a->data.data(), a->data.size(), nullptr);
	return 0;
}
static int on_header(void *arg, void *) {
	App *a = (App *)arg;
	m2d_info_t info;
	a->f->get_info(a->ctx, &info);
	int w = (info.src_width + 15) & ~15, h = (info.src_height + 15) & ~15, n = info.frame_num + 16;
	if (n > 16 && a->f == m2d_gpu_func) n = 16;
	// (allocated once: the sequence header comes back with every GOP / repetition, 50 MB of clearing and copying each
	// time would be timed as parse)
	const size_t fb = (size_t)w * h * 3 / 2 + 32;
	if ((int)a->bufs.size() != n || a->bufs[0].size() != fb) a->bufs.assign(n, std::vector<uint8_t>(fb));
	a->frames.assign(n, m2d_frame_t());
	for (int i = 0; i < n; ++i) {
		uint8_t *p = a->bufs[i].data();
		p += (16 - ((uintptr_t)p & 15)) & 15;
		a->frames[i].luma = p;
		a->frames[i].chroma = p + (size_t)w * h;
	}
	return a->f->set_frames(a->ctx, n, a->frames.data(), nullptr, 0);
}
static int on_tap(void *arg, int, const void *) { ((App *)arg)->pictures++; return 0; }

int main(int argc, char **argv) {
	if (argc < 3) { fprintf(stderr, "usage: parsebench h264|mpeg2 file [repetitions]\n"); return 2; }
	App a;
	a.f = !strcmp(argv[1], "mpeg2") ? m2d_gpu_func : h264d_gpu_func;
	FILE *fp = fopen(argv[2], "rb");
	if (!fp) { perror(argv[2]); return 1; }
	fseek(fp, 0, SEEK_END);
	a.data.resize(ftell(fp));
	fseek(fp, 0, SEEK_SET);
	if (fread(a.data.data(), 1, a.data.size(), fp) != a.data.size()) return 1;
	fclose(fp);
	int reps = argc > 3 ? atoi(argv[3]) : 5;
	double best = 1e30;
	long pics = 0;
	sample_start();
	for (int r = 0; r < reps; ++r) {
		std::vector<uint8_t> ctx(a.f->context_size);
		a.ctx = ctx.data();
		a.fed = false;
		a.pictures = 0;
		a.f->init(a.ctx, -1, on_header, &a);
		b2_context_set_tap(a.ctx, on_tap, &a, 1);
		dec_bits_set_callback(a.f->stream_pos(a.ctx), on_refill, &a);
		auto t0 = std::chrono::steady_clock::now();
		while (a.f->decode_picture(a.ctx) >= 0) {
			m2d_frame_t frm;
			while (a.f->get_decoded_frame(a.ctx, &frm, 0) > 0) {}
		}
		double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		b2_context_release(a.ctx);
		if (dt < best) best = dt;
		pics = a.pictures;
	}
	sample_stop();
	printf("%ld pictures, %.3f ms per picture (best of %d), %.1f Mbit/s\n", pics, 1000.0 * best / (pics ? pics : 1), reps,
	       a.data.size() * 8.0 / best / 1e6);
	return 0;
}
