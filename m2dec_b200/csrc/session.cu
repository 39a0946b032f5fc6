// session.cu — device-side C ABI (include/b2_recon.h): a reconstruction session owns one
// CUDA stream, a pool of device-resident NV12 frames and a pinned staging ring through which
// each picture's SoA travels host -> HBM in a single asynchronous copy.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "device_common.cuh"

namespace b2 {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

}  // namespace b2

using namespace b2;

enum { NSLOT = 4 };

// B2_TRACE=1: where the calling (parsing) thread's time goes inside this file, per session, printed at destroy.
// Development aid for the end-to-end path (every microsecond here is taken from the parse).
#include <chrono>
#include <unistd.h>
namespace {
const bool g_trace = getenv("B2_TRACE") != nullptr;
enum { T_CHECK, T_SLOT, T_PACK, T_H2D, T_LAUNCH, T_EVENTS, T_D2H_QUEUE, T_FETCH_WAIT, T_CREATE, T_GPU, T_N };
// ("device time" is not the calling thread's: copy + kernels of a picture on the session's stream, from CUDA events)
const char *const g_trace_names[T_N] = {"record check", "slot wait", "pack SoA", "H2D enqueue", "kernel launches", "event records", "D2H enqueue", "frame wait", "session create (per picture)", "device time"};
struct Trace {
	double us[T_N] = {0};
	uint64_t pictures = 0;
	std::chrono::steady_clock::time_point t;
	inline void begin() { if (g_trace) t = std::chrono::steady_clock::now(); }
	inline void lap(int k) {
		if (!g_trace) return;
		auto n = std::chrono::steady_clock::now();
		us[k] += std::chrono::duration<double, std::micro>(n - t).count();
		t = n;
	}
};
}  // namespace

// Process-wide free lists: cudaMalloc / cudaMallocHost / cudaFree synchronise the whole device, so a
// server that opens and closes streams all the time must not call them per stream.
namespace {
struct CachedBuf { int dev; bool pinned; size_t bytes; void *ptr; };
std::mutex g_cache_mu;
std::vector<CachedBuf> g_cache;

// sizes are rounded up to 1 MiB buckets so that buffers of closed streams fit the next ones
int cache_get(int dev, bool pinned, size_t bytes, void **out, size_t *actual = nullptr) {
	bytes = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
	if (actual) *actual = bytes;
	{
		std::lock_guard<std::mutex> lk(g_cache_mu);
		size_t best = (size_t)-1;
		for (size_t i = 0; i < g_cache.size(); ++i) {
			const CachedBuf &c = g_cache[i];
			if (c.dev == dev && c.pinned == pinned && c.bytes >= bytes && c.bytes <= bytes * 2 &&
			    (best == (size_t)-1 || c.bytes < g_cache[best].bytes)) best = i;
		}
		if (best != (size_t)-1) {
			*out = g_cache[best].ptr;
			if (actual) *actual = g_cache[best].bytes;
			g_cache.erase(g_cache.begin() + best);
			return 0;
		}
	}
	if (g_trace) fprintf(stderr, "m2dec_b200 trace: buffer cache miss: %zu MiB %s\n", bytes >> 20, pinned ? "page-locked" : "device");
	if (pinned) B2_CUDA_OK(cudaMallocHost(out, bytes));
	else B2_CUDA_OK(cudaMalloc(out, bytes));
	return 0;
}
void cache_put(int dev, bool pinned, size_t bytes, void *ptr) {
	if (!ptr) return;
	std::lock_guard<std::mutex> lk(g_cache_mu);
	g_cache.push_back(CachedBuf{dev, pinned, bytes, ptr});
}
// Free the oldest cached buffers until at most `keep` bytes of each kind (device / pinned) stay on the
// free lists: a long-running server that has seen many geometries does not keep the peak of all of them.
constexpr size_t kCacheKeepBytes = (size_t)4 << 30;
void cache_trim(size_t keep) {
	std::vector<CachedBuf> drop;
	{
		std::lock_guard<std::mutex> lk(g_cache_mu);
		for (int pinned = 0; pinned < 2; ++pinned) {
			size_t total = 0;
			for (const CachedBuf &c : g_cache) if (c.pinned == (pinned != 0)) total += c.bytes;
			for (size_t i = 0; i < g_cache.size() && total > keep;) {
				if (g_cache[i].pinned == (pinned != 0)) {
					total -= g_cache[i].bytes;
					drop.push_back(g_cache[i]);
					g_cache.erase(g_cache.begin() + i);
				} else ++i;
			}
		}
	}
	for (const CachedBuf &c : drop) {
		cudaSetDevice(c.dev);
		if (c.pinned) cudaFreeHost(c.ptr); else cudaFree(c.ptr);
	}
}
}  // namespace

// Release every buffer on the process-wide free lists (device and page-locked host memory of closed sessions).
static void pool_drain();
extern "C" void b2_cache_trim(void) { pool_drain(); cache_trim(0); }

struct StageSlot {
	uint8_t *host;
	uint8_t *dev;
	size_t cap, cap_host, cap_dev;
	cudaEvent_t done;      // recorded after the kernels that read the slot
	cudaEvent_t t0, t1;    // B2_TRACE: before the picture's copy / after its kernels, with timing
	bool pending;
};

struct b2_session {
	int dev, codec, width, height, nframes;
	size_t frame_bytes, frames_alloc, sync_alloc;
	cudaStream_t stream, copy_stream;
	uint8_t *frames;
	int *sync_area;                 // wavefront ticket / progress flags of this stream
	cudaEvent_t frame_done[64];     // recorded after the kernels that write frame i
	bool frame_written[64];
	// Caller-owned host frame buffers (m2d_frame_t of set_frames).  A frame whose host buffers could be
	// page-locked is copied out as soon as its picture is queued ("eager"), so that the application's
	// peek / get finds it there; otherwise it is copied when asked for.
	uint8_t *host_luma[64], *host_chroma[64];
	uint8_t host_state[64];         // 0 not looked at yet, 1 page-locked by the application, 2 pageable
	cudaEvent_t copy_done[64];      // recorded on the copy stream after the eager copy of frame i
	bool eager[64];                 // host buffers of frame i hold (or will hold, after copy_done) its current content
	StageSlot slots[NSLOT];
	int next_slot;
	uint64_t launches;
	uint64_t h2d_bytes, d2h_bytes;
	bool check_records;             // run b2_*_picture_check on every submitted picture (default)
	bool built;                     // session_build went through (only such a session is parked for the next stream)
	// Digest mode (b2_session_set_md5): every reconstructed picture is hashed on the device (md5.cu), on one of a
	// few side streams, and only its 16-byte MD5 comes back; the pixels stay in HBM.
	bool md5_on;
	int md5_stride, md5_left, md5_top, md5_w, md5_h;
	cudaStream_t md5_streams[12];   // one hash takes tens of milliseconds (a serial chain): several run side by side
	size_t md5_dev_alloc, md5_host_alloc;
	int md5_next;
	uint8_t *md5_dev, *md5_host;    // [nframes][16]
	cudaEvent_t md5_done[64];       // digest of frame i is in md5_host
	bool md5_pending[64];
	Trace trace;
};

extern "C" const char *b2_last_error(void) { return b2::g_err; }

// A server runs one session (two CUDA streams) per decoding thread.  CUDA maps streams onto 8 hardware queues by
// default; streams that share a queue wait for each other's work in issue order, and with 16 sessions the pictures
// of one stream sat behind another stream's kernels for milliseconds (16 host threads: 2.87 k -> 3.3 k frames/s
// with 16 queues; 32 queues: 3.1 k).  The variable is read when the process creates its CUDA context, so it is
// set when the library is loaded, unless the application chose a value itself.
namespace {
struct MoreQueues {
	MoreQueues() {
		// one queue per host core (a session per decoding thread is the expected use), within what CUDA accepts
		long n = sysconf(_SC_NPROCESSORS_ONLN);
		n = n < 8 ? 8 : (n > 32 ? 32 : n);
		char v[8];
		snprintf(v, sizeof(v), "%ld", n);
		setenv("CUDA_DEVICE_MAX_CONNECTIONS", v, 0);
	}
} g_more_queues;
}  // namespace

extern "C" int b2_device_count(void) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) {
		set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
		return -1;
	}
	return n;
}

// Gives a staging slot buffers for at least `bytes`.  Sized for a dense picture of this geometry (2.5 frames' worth: what
// a dense 1080p picture's records plus the device-side scratch of mv_on_device come to) so that it rarely has to
// grow, and in 4 MiB steps: slots of one geometry then nearly always ask for the same size, which the buffer cache
// has (an odd size is a cudaMalloc + cudaMallocHost in the middle of decoding, and those stall every stream of
// the device).
static int slot_grow(b2_session *s, StageSlot *sl, size_t bytes) {
	cache_put(s->dev, true, sl->cap_host, sl->host);
	cache_put(s->dev, false, sl->cap_dev, sl->dev);
	sl->host = sl->dev = nullptr;
	sl->cap = 0;
	size_t cap = bytes + bytes / 2 + 4096, typical = (size_t)s->width * s->height * 15 / 4, got_h = 0, got_d = 0;
	if (cap < typical) cap = typical;
	cap = (cap + (4u << 20) - 1) & ~(size_t)((4u << 20) - 1);
	if (cache_get(s->dev, true, cap, (void **)&sl->host, &got_h) < 0) return -1;
	if (cache_get(s->dev, false, cap, (void **)&sl->dev, &got_d) < 0) return -1;
	sl->cap_host = got_h;
	sl->cap_dev = got_d;
	sl->cap = got_h < got_d ? got_h : got_d;
	return 0;
}

// Parked sessions.  A stream that ends leaves its session -- two streams, ~50 events, frame pool, staging slots -- to the
// next stream of the same geometry instead of tearing it down: with closed GOPs or short clips as the unit of work a
// session lives for a dozen pictures, and creating its streams and events cost 0.07-0.1 ms per picture with sixteen
// threads doing the same (the driver serialises them).  b2_cache_trim() frees the parked sessions too.
namespace {
std::mutex g_pool_mu;
std::vector<b2_session *> g_pool;
constexpr size_t kPoolMax = 64;
const bool g_pool_on = getenv("B2_NO_SESSION_POOL") == nullptr;
}  // namespace

static int session_build(b2_session *s) {
	B2_CUDA_OK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
	B2_CUDA_OK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
	for (int i = 0; i < s->nframes; ++i) {
		B2_CUDA_OK(cudaEventCreateWithFlags(&s->frame_done[i], cudaEventDisableTiming));
		// waited for by host threads that have nothing else to do meanwhile: block, do not spin
		B2_CUDA_OK(cudaEventCreateWithFlags(&s->copy_done[i], cudaEventDisableTiming | cudaEventBlockingSync));
	}
	if (cache_get(s->dev, false, s->frame_bytes * s->nframes + H4_POOL_TAIL, (void **)&s->frames, &s->frames_alloc) < 0) return -1;
	B2_CUDA_OK(cudaMemsetAsync(s->frames, 0, s->frame_bytes * s->nframes, s->stream));
	if (cache_get(s->dev, false, h264_sync_bytes(1, s->width / 16, s->height / 16), (void **)&s->sync_area, &s->sync_alloc) < 0) return -1;
	for (int i = 0; i < NSLOT; ++i) {
		B2_CUDA_OK(cudaEventCreateWithFlags(&s->slots[i].done, cudaEventDisableTiming | cudaEventBlockingSync));
		// all slots at once: a process then reaches the number of buffers it will ever hold as soon as its threads have
		// their sessions open, not a few passes over the streams later
		if (slot_grow(s, &s->slots[i], 1) < 0) return -1;
	}
	return 0;
}

extern "C" void b2_session_destroy(b2_session *s);
static void session_free(b2_session *s);

extern "C" int b2_session_create(int cuda_device, int codec, int width, int height, int nframes, b2_session **out) {
	if (!out || width <= 0 || height <= 0 || (width & 15) || (height & 15) || nframes <= 0) {
		set_error("b2_session_create: bad arguments (%dx%d, %d frames)", width, height, nframes);
		return -1;
	}
	int n = b2_device_count();
	if (n <= 0) {
		if (n == 0) set_error("no CUDA device visible");
		return -1;
	}
	B2_CUDA_OK(cudaSetDevice(cuda_device));
	if (nframes > 64) { set_error("b2_session_create: at most 64 frames"); return -1; }
	{
		b2_session *parked = nullptr;
		{
			std::lock_guard<std::mutex> lk(g_pool_mu);
			for (size_t i = 0; i < g_pool.size(); ++i) {
				b2_session *c = g_pool[i];
				if (c->dev == cuda_device && c->codec == codec && c->width == width && c->height == height && c->nframes == nframes) {
					parked = c;
					g_pool.erase(g_pool.begin() + i);
					break;
				}
			}
		}
		if (parked) {
			// (session_park left it idle and in its initial state; the frames start out black like a new session's)
			parked->trace.begin();
			if (cudaMemsetAsync(parked->frames, 0, parked->frame_bytes * parked->nframes, parked->stream) == cudaSuccess) {
				parked->trace.lap(T_CREATE);
				*out = parked;
				return 0;
			}
			cudaGetLastError();
			session_free(parked);
		}
	}
	b2_session *s = new b2_session();
	memset(s, 0, sizeof(*s));
	s->dev = cuda_device;
	s->codec = codec;
	s->width = width;
	s->height = height;
	s->nframes = nframes;
	s->frame_bytes = (size_t)width * height * 3 / 2;
	s->check_records = true;
	// a failure anywhere below releases what was obtained so far (b2_session_destroy copes with a partly built
	// session: everything is zero until it is created)
	s->trace.begin();
	if (session_build(s) < 0) {
		b2_session_destroy(s);
		return -1;
	}
	s->built = true;
	s->trace.lap(T_CREATE);
	*out = s;
	return 0;
}

// Releases everything a (possibly partly built) session holds; its streams have been drained or never used.
static void session_free(b2_session *s) {
	cudaSetDevice(s->dev);
	for (int i = 0; i < NSLOT; ++i) {
		cache_put(s->dev, true, s->slots[i].cap_host, s->slots[i].host);
		cache_put(s->dev, false, s->slots[i].cap_dev, s->slots[i].dev);
		if (s->slots[i].done) cudaEventDestroy(s->slots[i].done);
		if (s->slots[i].t0) { cudaEventDestroy(s->slots[i].t0); cudaEventDestroy(s->slots[i].t1); }
	}
	for (int i = 0; i < 12; ++i)
		if (s->md5_streams[i]) cudaStreamDestroy(s->md5_streams[i]);
	// (through the free lists: cudaMalloc / cudaFree would wait for every hash in flight on the device)
	cache_put(s->dev, false, s->md5_dev_alloc, s->md5_dev);
	cache_put(s->dev, true, s->md5_host_alloc, s->md5_host);
	for (int i = 0; i < s->nframes; ++i) {
		if (s->frame_done[i]) cudaEventDestroy(s->frame_done[i]);
		if (s->copy_done[i]) cudaEventDestroy(s->copy_done[i]);
		if (s->md5_done[i]) cudaEventDestroy(s->md5_done[i]);
	}
	cache_put(s->dev, false, s->frames_alloc, s->frames);
	cache_put(s->dev, false, s->sync_alloc, s->sync_area);
	if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
	if (s->stream) cudaStreamDestroy(s->stream);
	cudaGetLastError();
	delete s;
}

// Back to the state b2_session_create hands out (the session is idle: its streams were synchronised), then parked.
static bool session_park(b2_session *s) {
	if (!g_pool_on || !s->built) return false;
	memset(s->frame_written, 0, sizeof(s->frame_written));
	memset(s->host_luma, 0, sizeof(s->host_luma));
	memset(s->host_chroma, 0, sizeof(s->host_chroma));
	memset(s->host_state, 0, sizeof(s->host_state));
	memset(s->eager, 0, sizeof(s->eager));
	memset(s->md5_pending, 0, sizeof(s->md5_pending));
	for (int i = 0; i < NSLOT; ++i) s->slots[i].pending = false;
	s->next_slot = 0;
	s->launches = s->h2d_bytes = s->d2h_bytes = 0;
	s->check_records = true;
	s->md5_on = false;
	s->md5_next = 0;
	s->md5_stride = s->md5_left = s->md5_top = s->md5_w = s->md5_h = 0;
	s->trace = Trace();
	std::lock_guard<std::mutex> lk(g_pool_mu);
	if (g_pool.size() >= kPoolMax) return false;
	g_pool.push_back(s);
	return true;
}

extern "C" void b2_session_destroy(b2_session *s) {
	if (!s) return;
	if (g_trace && s->trace.pictures) {
		char line[768];
		int n = snprintf(line, sizeof(line), "m2dec_b200 trace: %llu pictures, us per picture:", (unsigned long long)s->trace.pictures);
		for (int k = 0; k < T_N; ++k) n += snprintf(line + n, sizeof(line) - n, " %s %.1f,", g_trace_names[k], s->trace.us[k] / s->trace.pictures);
		fprintf(stderr, "%s\n", line);
	}
	cudaSetDevice(s->dev);
	bool idle = true;
	if (s->stream) idle &= cudaStreamSynchronize(s->stream) == cudaSuccess;
	if (s->copy_stream) idle &= cudaStreamSynchronize(s->copy_stream) == cudaSuccess;
	for (int i = 0; i < 12; ++i)
		if (s->md5_streams[i]) idle &= cudaStreamSynchronize(s->md5_streams[i]) == cudaSuccess;
	cudaGetLastError();
	if (idle && session_park(s)) return;       // (a session whose work ended in an error is not kept)
	session_free(s);
	cache_trim(kCacheKeepBytes);
}

static void pool_drain() {
	std::vector<b2_session *> all;
	{
		std::lock_guard<std::mutex> lk(g_pool_mu);
		all.swap(g_pool);
	}
	for (b2_session *s : all) session_free(s);
}

// Obtain a staging slot of at least `bytes`, waiting for its previous use to drain.
static int acquire_slot(b2_session *s, size_t bytes, StageSlot **out) {
	StageSlot *sl = &s->slots[s->next_slot];
	s->next_slot = (s->next_slot + 1) % NSLOT;
	if (sl->pending) {
		B2_CUDA_OK(cudaEventSynchronize(sl->done));
		sl->pending = false;
		if (g_trace && sl->t0) {
			float ms = 0;
			if (cudaEventElapsedTime(&ms, sl->t0, sl->t1) == cudaSuccess) s->trace.us[T_GPU] += 1000.0 * ms;
		}
	}
	if (g_trace && !sl->t0) { cudaEventCreate(&sl->t0); cudaEventCreate(&sl->t1); }
	if (sl->cap < bytes && slot_grow(s, sl, bytes) < 0) return -1;
	*out = sl;
	return 0;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Records into the page-locked staging slot with non-temporal stores: the slot is read by the copy engine only,
// so the 2 MB of a picture neither wait for the slot's old lines to be fetched (read-for-ownership) nor push the
// parse's working set out of the core's caches.  dst is 16-byte aligned (slot offsets are multiples of 256).
#include <emmintrin.h>
static void stream_copy(void *dst, const void *src, size_t n) {
	uint8_t *d = (uint8_t *)dst;
	const uint8_t *s = (const uint8_t *)src;
	if (((uintptr_t)d & 15) == 0) {
		for (; n >= 64; n -= 64, d += 64, s += 64) {
			const __m128i a = _mm_loadu_si128((const __m128i *)s), b = _mm_loadu_si128((const __m128i *)(s + 16));
			const __m128i c = _mm_loadu_si128((const __m128i *)(s + 32)), e = _mm_loadu_si128((const __m128i *)(s + 48));
			_mm_stream_si128((__m128i *)d, a);
			_mm_stream_si128((__m128i *)(d + 16), b);
			_mm_stream_si128((__m128i *)(d + 32), c);
			_mm_stream_si128((__m128i *)(d + 48), e);
		}
	}
	if (n) memcpy(d, s, n);
}

extern "C" int b2_session_bind_host_frame(b2_session *s, int idx, uint8_t *luma, uint8_t *chroma) {
	if (!s || (unsigned)idx >= (unsigned)s->nframes) return -1;
	if (s->eager[idx]) {
		cudaSetDevice(s->dev);
		cudaStreamSynchronize(s->copy_stream);
	}
	if (s->host_luma[idx] != luma || s->host_chroma[idx] != chroma) s->host_state[idx] = 0;
	s->host_luma[idx] = luma;
	s->host_chroma[idx] = chroma;
	s->eager[idx] = false;
	return 0;
}

extern "C" void *b2_host_alloc(size_t bytes, int cuda_device) {
	void *p = nullptr;
	if (b2_device_count() <= 0) return nullptr;
	if (cuda_device >= 0 && cudaSetDevice(cuda_device) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}
extern "C" void b2_host_free(void *p) {
	if (p) cudaFreeHost(p);
}

// Before the kernels of a picture that (re)writes frame `dst` are queued: they must not overtake an
// eager copy of the frame's previous content that is still in flight.
static int before_write(b2_session *s, int dst) {
	if (s->eager[dst]) B2_CUDA_OK(cudaStreamWaitEvent(s->stream, s->copy_done[dst], 0));
	s->eager[dst] = false;
	// ... nor a hash of the previous content that is still reading it
	if (s->md5_pending[dst]) B2_CUDA_OK(cudaStreamWaitEvent(s->stream, s->md5_done[dst], 0));
	return 0;
}

// After frame_done[dst] has been recorded: queue the copy into the caller's frame buffers when they are
// page-locked.  Pageable buffers are left to b2_session_download: a copy into pageable
// memory would block the parsing thread here until the kernels are done.
static int after_write(b2_session *s, int dst) {
	if (s->md5_on) {
		// digest mode: hash the frame behind its kernels, bring 16 bytes back; no pixels cross PCIe
		cudaStream_t ms = s->md5_streams[s->md5_next];
		s->md5_next = (s->md5_next + 1) % 12;
		B2_CUDA_OK(cudaStreamWaitEvent(ms, s->frame_done[dst], 0));
		if (md5_frame(ms, s->frames, s->frame_bytes, dst, s->md5_stride, (size_t)s->width * s->height, s->md5_left, s->md5_top, s->md5_w, s->md5_h,
		              s->md5_dev + 16 * dst) < 0) return -1;
		s->launches++;
		B2_CUDA_OK(cudaMemcpyAsync(s->md5_host + 16 * dst, s->md5_dev + 16 * dst, 16, cudaMemcpyDeviceToHost, ms));
		B2_CUDA_OK(cudaEventRecord(s->md5_done[dst], ms));
		s->md5_pending[dst] = true;
		s->d2h_bytes += 16;
		return 0;
	}
	static const bool off = getenv("B2_NO_EAGER_D2H") != nullptr;
	if (off || !s->host_luma[dst] || !s->host_chroma[dst]) return 0;
	const size_t ls = (size_t)s->width * s->height;
	if (s->host_state[dst] == 0) {
		// page-locked by the application (b2_host_alloc, cudaHostAlloc, cudaHostRegister)?  Page-locking here, per
		// stream, would cost more than it saves (milliseconds per buffer).
		cudaPointerAttributes al, ac;
		const bool pinned = cudaPointerGetAttributes(&al, s->host_luma[dst]) == cudaSuccess && al.type == cudaMemoryTypeHost &&
		                    cudaPointerGetAttributes(&ac, s->host_chroma[dst]) == cudaSuccess && ac.type == cudaMemoryTypeHost;
		cudaGetLastError();
		s->host_state[dst] = pinned ? 1 : 2;
	}
	if (s->host_state[dst] != 1) return 0;
	const uint8_t *src = s->frames + s->frame_bytes * dst;
	B2_CUDA_OK(cudaStreamWaitEvent(s->copy_stream, s->frame_done[dst], 0));
	B2_CUDA_OK(cudaMemcpyAsync(s->host_luma[dst], src, ls, cudaMemcpyDeviceToHost, s->copy_stream));
	B2_CUDA_OK(cudaMemcpyAsync(s->host_chroma[dst], src + ls, ls / 2, cudaMemcpyDeviceToHost, s->copy_stream));
	B2_CUDA_OK(cudaEventRecord(s->copy_done[dst], s->copy_stream));
	s->d2h_bytes += ls + ls / 2;
	s->eager[dst] = true;
	return 0;
}

extern "C" int b2_session_submit_mpeg2(b2_session *s, const b2_m2_picture *p) {
	if (!s || !p || s->codec != B2_CODEC_MPEG2) return -1;
	if (p->mb_w * 16 != s->width || p->mb_h * 16 != s->height) {
		set_error("picture %dx%d MB does not match session %dx%d", p->mb_w, p->mb_h, s->width, s->height);
		return -1;
	}
	s->trace.begin();
	if (s->check_records && b2_mpeg2_picture_check(p, s->nframes) < 0) return -1;
	s->trace.lap(T_CHECK);
	B2_CUDA_OK(cudaSetDevice(s->dev));
	size_t off_mbs = align_up(sizeof(M2Job), 256);
	size_t off_coef = align_up(off_mbs + sizeof(b2_m2_mb) * p->n_mbs, 256);
	size_t total = off_coef + (size_t)p->n_blocks * 128;
	StageSlot *sl;
	if (acquire_slot(s, total, &sl) < 0) return -1;
	s->trace.lap(T_SLOT);
	size_t luma = (size_t)s->width * s->height;
	M2Job job;
	job.dst_y = s->frames + s->frame_bytes * p->dst;
	job.dst_c = job.dst_y + luma;
	job.fwd_y = s->frames + s->frame_bytes * p->fwd;
	job.fwd_c = job.fwd_y + luma;
	job.bwd_y = s->frames + s->frame_bytes * p->bwd;
	job.bwd_c = job.bwd_y + luma;
	job.mbs = (const b2_m2_mb *)(sl->dev + off_mbs);
	job.coefs = (const int16_t *)(sl->dev + off_coef);
	job.mb_w = p->mb_w;
	job.mb_h = p->mb_h;
	memcpy(sl->host, &job, sizeof(job));
	// (streamed past the caches like the H.264 records: the slot is read by the copy engine only.  Writing the records
	// in place during the parse was tried: one thread gains 0.1 ms per picture, sixteen lose a third of their rate (slots
	// come back 0.85 ms late and the copy's enqueue blocks for 2 ms: the device falls behind, presumably because the
	// copy engine then reads lines that sit modified in sixteen cores' caches instead of memory).)
	stream_copy(sl->host + off_mbs, p->mbs, sizeof(b2_m2_mb) * p->n_mbs);
	if (p->n_blocks) stream_copy(sl->host + off_coef, p->coefs, (size_t)p->n_blocks * 128);
	_mm_sfence();
	s->trace.lap(T_PACK);
	B2_CUDA_OK(cudaMemcpyAsync(sl->dev, sl->host, total, cudaMemcpyHostToDevice, s->stream));
	s->trace.lap(T_H2D);
	s->h2d_bytes += total;
	if (before_write(s, p->dst) < 0) return -1;
	if (launch_mpeg2(s->stream, (const M2Job *)sl->dev, 1, p->mb_w, p->mb_h) < 0) return -1;
	s->launches++;
	s->trace.lap(T_LAUNCH);
	B2_CUDA_OK(cudaEventRecord(sl->done, s->stream));
	B2_CUDA_OK(cudaEventRecord(s->frame_done[p->dst], s->stream));
	s->frame_written[p->dst] = true;
	sl->pending = true;
	s->trace.lap(T_EVENTS);
	const int r = after_write(s, p->dst);
	s->trace.lap(T_D2H_QUEUE);
	s->trace.pictures++;
	return r;
}

extern "C" int b2_session_submit_h264(b2_session *s, const b2_h264_picture *p) {
	if (!s || !p || s->codec != B2_CODEC_H264) return -1;
	if (p->mb_w * 16 != s->width || p->mb_h * 16 != s->height) {
		set_error("picture %dx%d MB does not match session %dx%d", p->mb_w, p->mb_h, s->width, s->height);
		return -1;
	}
	s->trace.begin();
	if (s->check_records && b2_h264_picture_check(p, s->nframes) < 0) return -1;
	s->trace.lap(T_CHECK);
	B2_CUDA_OK(cudaSetDevice(s->dev));
	const bool dev_mv = p->mv_on_device != 0;
	if (dev_mv && (!p->bs_on_device || p->n_mcs || p->n_items_first)) { set_error("mv_on_device needs bs_on_device and carries no motion records"); return -1; }
	// motion records: copied from the host, or (mv_on_device) the partitions are copied and the records are device scratch
	const size_t n_recs = dev_mv ? p->n_parts : p->n_mcs, rec_bytes = dev_mv ? sizeof(b2_h264_part) * p->n_parts : sizeof(b2_h264_mc) * p->n_mcs;
	size_t off_mbs = align_up(sizeof(H4Job), 256);
	size_t off_dbk = align_up(off_mbs + sizeof(b2_h264_mb) * p->n_mbs, 256);
	size_t off_mcs = align_up(off_dbk + sizeof(b2_h264_dbk) * p->n_mbs, 256);
	size_t off_coef = align_up(off_mcs + rec_bytes, 256);
	size_t off_items = align_up(off_coef + (size_t)p->n_coef_recs * 32, 256);
	size_t total = off_items + (dev_mv ? 0 : (size_t)p->n_items * 4);
	// device only, behind what is copied (mv_on_device): motion records, work list, owner table, motion field
	size_t off_x_mcs = align_up(total, 256);
	size_t off_x_items = align_up(off_x_mcs + n_recs * sizeof(b2_h264_mc), 256);
	size_t off_x_owner = align_up(off_x_items + ((size_t)p->n_items + 64) * 4, 256);
	size_t off_x_fref = align_up(off_x_owner + n_recs * 4, 256);
	size_t off_x_fmv = align_up(off_x_fref + (size_t)p->n_mbs * 16, 256);
	size_t total_dev = dev_mv ? off_x_fmv + (size_t)p->n_mbs * 64 : total;
	StageSlot *sl;
	if (acquire_slot(s, total_dev, &sl) < 0) return -1;
	s->trace.lap(T_SLOT);
	H4Job job;
	memset(&job, 0, sizeof(job));
	job.frames = s->frames;
	job.frame_bytes = s->frame_bytes;
	job.mbs = (const b2_h264_mb *)(sl->dev + off_mbs);
	job.dbk = (const b2_h264_dbk *)(sl->dev + off_dbk);
	job.mcs = (const b2_h264_mc *)(sl->dev + (dev_mv ? off_x_mcs : off_mcs));
	job.coefs = (const int16_t *)(sl->dev + off_coef);
	job.dst = p->dst;
	job.mb_w = p->mb_w;
	job.mb_h = p->mb_h;
	job.n_intra = (int32_t)p->n_intra;
	job.items = (const uint32_t *)(sl->dev + (dev_mv ? off_x_items : off_items));
	job.n_items = dev_mv ? 0 : p->n_items;        // (mv_on_device: h264_items_kernel writes it)
	job.n_first = p->n_items_first;
	job.derive_bs = p->bs_on_device;
	job.derive_mv = p->mv_on_device;
	if (dev_mv) {
		job.parts = (const b2_h264_part *)(sl->dev + off_mcs);
		job.n_parts = p->n_parts;
		job.owner = (uint32_t *)(sl->dev + off_x_owner);
		job.field_ref = (int8_t *)(sl->dev + off_x_fref);
		job.field_mv = (uint32_t *)(sl->dev + off_x_fmv);
	}
	memcpy(sl->host, &job, sizeof(job));
	stream_copy(sl->host + off_mbs, p->mbs, sizeof(b2_h264_mb) * p->n_mbs);
	stream_copy(sl->host + off_dbk, p->dbk, sizeof(b2_h264_dbk) * p->n_mbs);
	if (rec_bytes) stream_copy(sl->host + off_mcs, dev_mv ? (const void *)p->parts : (const void *)p->mcs, rec_bytes);
	if (p->n_coef_recs) stream_copy(sl->host + off_coef, p->coefs, (size_t)p->n_coef_recs * 32);
	if (!dev_mv && p->n_items) stream_copy(sl->host + off_items, p->items, (size_t)p->n_items * 4);
	_mm_sfence();                 // the streamed stores are globally visible before the copy engine is told to read
	s->trace.lap(T_PACK);
	if (g_trace) cudaEventRecord(sl->t0, s->stream);
	B2_CUDA_OK(cudaMemcpyAsync(sl->dev, sl->host, total, cudaMemcpyHostToDevice, s->stream));
	s->h2d_bytes += total;
	if (before_write(s, p->dst) < 0) return -1;
	s->trace.lap(T_H2D);
	if (launch_h264(s->stream, (const H4Job *)sl->dev, 1, p->mb_w, p->mb_h, p->n_items,
	                p->n_intra > 0 && !h264_intra_sparse((int)p->n_intra, (int)p->n_mbs), p->n_intra > 0 && h264_intra_sparse((int)p->n_intra, (int)p->n_mbs),
	                p->deblock != 0, s->sync_area, &s->launches, p->n_items_first, p->bs_on_device != 0, dev_mv) < 0) return -1;
	s->trace.lap(T_LAUNCH);
	if (g_trace) cudaEventRecord(sl->t1, s->stream);
	B2_CUDA_OK(cudaEventRecord(sl->done, s->stream));
	B2_CUDA_OK(cudaEventRecord(s->frame_done[p->dst], s->stream));
	s->frame_written[p->dst] = true;
	sl->pending = true;
	s->trace.lap(T_EVENTS);
	const int rc = after_write(s, p->dst);
	s->trace.lap(T_D2H_QUEUE);
	s->trace.pictures++;
	return rc;
}

extern "C" int b2_session_set_md5(b2_session *s, int on, int stride, int crop_left, int crop_top, int width, int height) {
	if (!s) return -1;
	if (!on) { s->md5_on = false; return 0; }
	if (stride <= 0 || crop_left < 0 || crop_top < 0 || width <= 0 || height <= 0 || crop_left + width > stride ||
	    (size_t)(crop_top + height) * stride > (size_t)s->width * s->height) {
		set_error("b2_session_set_md5: window %d,%d %dx%d (pitch %d) outside the %dx%d frame", crop_left, crop_top, width, height, stride, s->width, s->height);
		return -1;
	}
	B2_CUDA_OK(cudaSetDevice(s->dev));
	if (!s->md5_dev) {
		if (cache_get(s->dev, false, 16 * 64, (void **)&s->md5_dev, &s->md5_dev_alloc) < 0) return -1;
		if (cache_get(s->dev, true, 16 * 64, (void **)&s->md5_host, &s->md5_host_alloc) < 0) return -1;
		for (int i = 0; i < 12; ++i) B2_CUDA_OK(cudaStreamCreateWithFlags(&s->md5_streams[i], cudaStreamNonBlocking));
		for (int i = 0; i < s->nframes; ++i) B2_CUDA_OK(cudaEventCreateWithFlags(&s->md5_done[i], cudaEventDisableTiming | cudaEventBlockingSync));
	}
	s->md5_stride = stride; s->md5_left = crop_left; s->md5_top = crop_top; s->md5_w = width; s->md5_h = height;
	s->md5_on = true;
	return 0;
}

extern "C" int b2_session_md5(b2_session *s, int idx, uint8_t digest[16]) {
	if (!s || (unsigned)idx >= (unsigned)s->nframes || !digest) return -1;
	if (!s->md5_on || !s->md5_pending[idx]) { set_error("b2_session_md5: frame %d has no digest (digest mode off, or nothing was decoded into it)", idx); return -1; }
	B2_CUDA_OK(cudaSetDevice(s->dev));
	if (cudaEventQuery(s->md5_done[idx]) != cudaSuccess) {
		cudaGetLastError();
		B2_CUDA_OK(cudaEventSynchronize(s->md5_done[idx]));
	}
	memcpy(digest, s->md5_host + 16 * idx, 16);
	return 0;
}

extern "C" int b2_session_set_record_check(b2_session *s, int on) {
	if (!s) return -1;
	s->check_records = on != 0;
	return 0;
}

extern "C" int b2_session_sync(b2_session *s) {
	if (!s) return -1;
	B2_CUDA_OK(cudaSetDevice(s->dev));
	B2_CUDA_OK(cudaStreamSynchronize(s->stream));
	return 0;
}

extern "C" int b2_session_download(b2_session *s, int idx, uint8_t *luma, uint8_t *chroma) {
	if (!s || (unsigned)idx >= (unsigned)s->nframes || !luma || !chroma) return -1;
	B2_CUDA_OK(cudaSetDevice(s->dev));
	size_t ls = (size_t)s->width * s->height;
	const uint8_t *src = s->frames + s->frame_bytes * idx;
	if (s->eager[idx] && luma == s->host_luma[idx] && chroma == s->host_chroma[idx]) {
		// already on its way (or there) since the picture was queued
		s->trace.begin();
		// (the copy usually finished pictures ago: ask before going to sleep on the event)
		if (cudaEventQuery(s->copy_done[idx]) != cudaSuccess) {
			cudaGetLastError();
			B2_CUDA_OK(cudaEventSynchronize(s->copy_done[idx]));
		}
		s->trace.lap(T_FETCH_WAIT);
		return 0;
	}
	// wait for the picture that last wrote this frame only, not for pictures queued behind it:
	// the host keeps parsing ahead while earlier frames are copied out
	if (s->frame_written[idx]) B2_CUDA_OK(cudaStreamWaitEvent(s->copy_stream, s->frame_done[idx], 0));
	B2_CUDA_OK(cudaMemcpyAsync(luma, src, ls, cudaMemcpyDeviceToHost, s->copy_stream));
	B2_CUDA_OK(cudaMemcpyAsync(chroma, src + ls, ls / 2, cudaMemcpyDeviceToHost, s->copy_stream));
	s->d2h_bytes += ls + ls / 2;
	B2_CUDA_OK(cudaStreamSynchronize(s->copy_stream));
	return 0;
}

extern "C" int b2_session_upload(b2_session *s, int idx, const uint8_t *luma, const uint8_t *chroma) {
	if (!s || (unsigned)idx >= (unsigned)s->nframes || !luma || !chroma) return -1;
	B2_CUDA_OK(cudaSetDevice(s->dev));
	size_t ls = (size_t)s->width * s->height;
	uint8_t *dst = s->frames + s->frame_bytes * idx;
	B2_CUDA_OK(cudaMemcpyAsync(dst, luma, ls, cudaMemcpyHostToDevice, s->stream));
	B2_CUDA_OK(cudaMemcpyAsync(dst + ls, chroma, ls / 2, cudaMemcpyHostToDevice, s->stream));
	B2_CUDA_OK(cudaStreamSynchronize(s->stream));
	return 0;
}

extern "C" uint64_t b2_session_launches(const b2_session *s) { return s ? s->launches : 0; }

extern "C" int b2_session_traffic(const b2_session *s, uint64_t bytes[2]) {
	if (!s || !bytes) return -1;
	bytes[0] = s->h2d_bytes;
	bytes[1] = s->d2h_bytes;
	return 0;
}
