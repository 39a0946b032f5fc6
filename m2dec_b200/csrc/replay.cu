// replay.cu — kernel replay (include/b2_recon.h, b2_replay_*): many independent streams'
// pictures resident in HBM, reconstructed step by step with one batched launch per kernel.
// This is how the reconstruction kernels are driven at throughput: a single 1080p picture is
// ~14 MB of traffic (~2 us at HBM speed, less than a launch), so frames of independent
// streams / closed GOPs are batched into each launch (SURVEY.md §7 "hard parts").
//
// The intra and deblocking kernels are wavefronts: one launch cannot finish faster than the
// dependency chain of a single picture (254 macroblock steps at 1080p), however many pictures it
// holds, and runs with few rows in flight while that chain ramps up and down.  The lanes are
// can therefore be split into groups ($B2_REPLAY_GROUPS), each with its own CUDA stream and chain of
// launches; the groups are independent, so their kernels overlap on the device.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "device_common.cuh"

using namespace b2;

namespace {
struct Blob { uint8_t *dev; size_t bytes; };
struct Step {
	std::vector<H4Job> h4;       // one job per lane (of the group) that has a picture at this step
	std::vector<M2Job> m2;
	void *d_jobs = nullptr;
	bool intra_dense = false, intra_sparse = false, deblock = false, derive_bs = false;
	uint32_t max_items = 0, max_first = 0;
};
struct Group {
	cudaStream_t stream = nullptr;
	int *sync_area = nullptr;
	std::vector<Step> steps;
	std::vector<cudaEvent_t> ev;     // 4 per step when timing
	cudaEvent_t done = nullptr;
};
}  // namespace

struct b2_replay {
	int dev, codec, width, height, lanes, pool;
	size_t frame_bytes;
	uint8_t *frames;
	cudaStream_t stream;             // time line: begin / end / mark events; the groups fork from and join it
	std::vector<Group> groups;       // lane l belongs to group l % groups.size()
	std::vector<Blob> blobs;
	std::vector<int> lane_len;
	bool dirty;
	b2_replay_stats stats;
	cudaEvent_t ev_begin, ev_end, ev_mark[2];
	bool timed_pending;
	float acc_ms[3];
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int b2_replay_create(int cuda_device, int codec, int width, int height, int lanes, int pool_frames, b2_replay **out) {
	if (!out || lanes <= 0 || pool_frames <= 0 || (width & 15) || (height & 15) || width <= 0 || height <= 0) return -1;
	if (b2_device_count() <= 0) return -1;
	B2_CUDA_OK(cudaSetDevice(cuda_device));
	b2_replay *r = new b2_replay();
	r->dev = cuda_device; r->codec = codec; r->width = width; r->height = height;
	r->lanes = lanes; r->pool = pool_frames;
	r->frame_bytes = (size_t)width * height * 3 / 2;
	r->frames = nullptr;
	r->dirty = true;
	r->timed_pending = false;
	memset(&r->stats, 0, sizeof(r->stats));
	r->lane_len.assign(lanes, 0);
	B2_CUDA_OK(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
	B2_CUDA_OK(cudaMalloc(&r->frames, r->frame_bytes * pool_frames * lanes + H4_POOL_TAIL));
	B2_CUDA_OK(cudaMemset(r->frames, 0, r->frame_bytes * pool_frames * lanes));
	// default: one group.  Measured on B200 (64 and 128 lanes of 1080p): 2 / 4 / 8 groups took
	// 1.04x / 1.15x / 1.19x and 0.99x / 1.09x / 1.26x the time of one group, and no better with the
	// wavefront kernels on high-priority streams: the inter kernel owns the whole register file of an
	// SM while it runs, so the chains of different groups hardly overlap.
	int ng = 1;
	if (const char *e = getenv("B2_REPLAY_GROUPS")) ng = atoi(e);
	if (ng < 1) ng = 1;
	if (ng > lanes) ng = lanes;
	r->groups.resize(ng);
	for (Group &g : r->groups) {
		B2_CUDA_OK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
		B2_CUDA_OK(cudaMalloc(&g.sync_area, h264_sync_bytes((lanes + ng - 1) / ng, width / 16, height / 16)));
		B2_CUDA_OK(cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming));
	}
	B2_CUDA_OK(cudaEventCreate(&r->ev_begin));
	B2_CUDA_OK(cudaEventCreate(&r->ev_end));
	B2_CUDA_OK(cudaEventCreate(&r->ev_mark[0]));
	B2_CUDA_OK(cudaEventCreate(&r->ev_mark[1]));
	*out = r;
	return 0;
}

extern "C" void b2_replay_destroy(b2_replay *r) {
	if (!r) return;
	cudaSetDevice(r->dev);
	cudaStreamSynchronize(r->stream);
	for (Blob &b : r->blobs) cudaFree(b.dev);
	for (Group &g : r->groups) {
		cudaStreamSynchronize(g.stream);
		for (Step &s : g.steps) if (s.d_jobs) cudaFree(s.d_jobs);
		for (cudaEvent_t e : g.ev) cudaEventDestroy(e);
		cudaEventDestroy(g.done);
		cudaFree(g.sync_area);
		cudaStreamDestroy(g.stream);
	}
	cudaEventDestroy(r->ev_begin);
	cudaEventDestroy(r->ev_end);
	cudaEventDestroy(r->ev_mark[0]);
	cudaEventDestroy(r->ev_mark[1]);
	cudaFree(r->frames);
	cudaStreamDestroy(r->stream);
	delete r;
}

static int upload(b2_replay *r, const std::vector<std::pair<const void *, size_t>> &parts, std::vector<size_t> *offs, uint8_t **dev) {
	size_t total = 0;
	offs->clear();
	for (auto &p : parts) { offs->push_back(total); total = align_up(total + p.second, 256); }
	if (total == 0) total = 256;
	uint8_t *d = nullptr;
	B2_CUDA_OK(cudaMalloc(&d, total));
	for (size_t i = 0; i < parts.size(); ++i)
		if (parts[i].second) B2_CUDA_OK(cudaMemcpy(d + (*offs)[i], parts[i].first, parts[i].second, cudaMemcpyHostToDevice));
	r->blobs.push_back(Blob{d, total});
	*dev = d;
	return 0;
}

extern "C" int b2_replay_add_h264(b2_replay *r, int lane, const b2_h264_picture *p) {
	if (!r || !p || r->codec != B2_CODEC_H264 || (unsigned)lane >= (unsigned)r->lanes) return -1;
	if (p->mb_w * 16 != r->width || p->mb_h * 16 != r->height || (unsigned)p->dst >= (unsigned)r->pool) return -1;
	if (b2_h264_picture_check(p, r->pool) < 0) return -1;
	if (p->mv_on_device) { set_error("kernel replay takes pictures with motion records (mv_on_device pictures go through a session)"); return -1; }
	B2_CUDA_OK(cudaSetDevice(r->dev));
	std::vector<size_t> off;
	uint8_t *d;
	if (upload(r, {{p->mbs, sizeof(b2_h264_mb) * p->n_mbs}, {p->dbk, sizeof(b2_h264_dbk) * p->n_mbs},
	               {p->mcs, sizeof(b2_h264_mc) * p->n_mcs}, {p->coefs, (size_t)p->n_coef_recs * 32}, {p->items, (size_t)p->n_items * 4}}, &off, &d) < 0) return -1;
	H4Job job;
	memset(&job, 0, sizeof(job));
	job.frames = r->frames + r->frame_bytes * r->pool * lane;
	job.frame_bytes = r->frame_bytes;
	job.mbs = (const b2_h264_mb *)(d + off[0]);
	job.dbk = (const b2_h264_dbk *)(d + off[1]);
	job.mcs = (const b2_h264_mc *)(d + off[2]);
	job.coefs = (const int16_t *)(d + off[3]);
	job.items = (const uint32_t *)(d + off[4]);
	job.n_items = p->n_items;
	job.n_first = p->n_items_first;
	job.derive_bs = p->bs_on_device;
	job.dst = p->dst; job.mb_w = p->mb_w; job.mb_h = p->mb_h; job.n_intra = (int32_t)p->n_intra;
	int t = r->lane_len[lane]++;
	std::vector<Step> &steps = r->groups[lane % r->groups.size()].steps;
	if ((int)steps.size() <= t) steps.resize(t + 1);
	Step &s = steps[t];
	s.h4.push_back(job);
	if (p->n_items > s.max_items) s.max_items = p->n_items;
	if (p->n_items_first > s.max_first) s.max_first = p->n_items_first;
	if (p->n_intra > 0) (h264_intra_sparse((int)p->n_intra, (int)p->n_mbs) ? s.intra_sparse : s.intra_dense) = true;
	s.deblock |= p->deblock != 0;
	s.derive_bs |= p->bs_on_device != 0 && p->deblock != 0;
	// algorithmic bytes: S(1 + 2 dbk) + R + C + M, split by kernel
	uint64_t n_inter = 0, c_inter = 0, c_intra = 0, rbytes = 0;
	for (uint32_t i = 0; i < p->n_mbs; ++i) {
		const b2_h264_mb &m = p->mbs[i];
		uint64_t c = (uint64_t)__builtin_popcount(m.nz) * 32;
		if (m.type == B2_H4_INTER) { n_inter++; c_inter += c; } else c_intra += c;
	}
	for (uint32_t i = 0; i < p->n_mcs; ++i) {
		int w = (p->mcs[i].size & 15) * 4, h = (p->mcs[i].size >> 4) * 4;
		rbytes += (uint64_t)w * h * 3 / 2 * ((p->mcs[i].ref_frame[0] >= 0) + (p->mcs[i].ref_frame[1] >= 0));
	}
	uint64_t S = (uint64_t)r->frame_bytes;
	r->stats.alg_bytes[0] += n_inter * 384 + rbytes + c_inter + (uint64_t)p->n_mbs * 32 + (uint64_t)p->n_mcs * sizeof(b2_h264_mc);
	if (p->n_intra) r->stats.alg_bytes[1] += (uint64_t)p->n_intra * 384 + c_intra + (uint64_t)p->n_mbs * 32;
	if (p->deblock) r->stats.alg_bytes[2] += 2 * S + (uint64_t)p->n_mbs * 16;
	r->stats.pictures++;
	r->dirty = true;
	return 0;
}

extern "C" int b2_replay_add_mpeg2(b2_replay *r, int lane, const b2_m2_picture *p) {
	if (!r || !p || r->codec != B2_CODEC_MPEG2 || (unsigned)lane >= (unsigned)r->lanes) return -1;
	if (p->mb_w * 16 != r->width || p->mb_h * 16 != r->height) return -1;
	if (b2_mpeg2_picture_check(p, r->pool) < 0) return -1;
	B2_CUDA_OK(cudaSetDevice(r->dev));
	std::vector<size_t> off;
	uint8_t *d;
	if (upload(r, {{p->mbs, sizeof(b2_m2_mb) * p->n_mbs}, {p->coefs, (size_t)p->n_blocks * 128}}, &off, &d) < 0) return -1;
	size_t luma = (size_t)r->width * r->height;
	uint8_t *base = r->frames + r->frame_bytes * r->pool * lane;
	M2Job job;
	job.dst_y = base + r->frame_bytes * p->dst; job.dst_c = job.dst_y + luma;
	job.fwd_y = base + r->frame_bytes * p->fwd; job.fwd_c = job.fwd_y + luma;
	job.bwd_y = base + r->frame_bytes * p->bwd; job.bwd_c = job.bwd_y + luma;
	job.mbs = (const b2_m2_mb *)(d + off[0]);
	job.coefs = (const int16_t *)(d + off[1]);
	job.mb_w = p->mb_w; job.mb_h = p->mb_h;
	int t = r->lane_len[lane]++;
	std::vector<Step> &steps = r->groups[lane % r->groups.size()].steps;
	if ((int)steps.size() <= t) steps.resize(t + 1);
	steps[t].m2.push_back(job);
	uint64_t rbytes = 0;
	for (uint32_t i = 0; i < p->n_mbs; ++i) {
		int f = p->mbs[i].flags;
		rbytes += ((f & B2_M2_FWD) ? 384 : 0) + ((f & B2_M2_BWD) ? 384 : 0);
	}
	r->stats.alg_bytes[0] += r->frame_bytes + rbytes + (uint64_t)p->n_blocks * 128 + (uint64_t)p->n_mbs * sizeof(b2_m2_mb);
	r->stats.pictures++;
	r->dirty = true;
	return 0;
}

static int finalize(b2_replay *r) {
	if (!r->dirty) return 0;
	for (Group &g : r->groups) {
		for (Step &s : g.steps) {
			if (s.d_jobs) { cudaFree(s.d_jobs); s.d_jobs = nullptr; }
			size_t bytes = r->codec == B2_CODEC_H264 ? s.h4.size() * sizeof(H4Job) : s.m2.size() * sizeof(M2Job);
			if (!bytes) continue;
			B2_CUDA_OK(cudaMalloc(&s.d_jobs, bytes));
			B2_CUDA_OK(cudaMemcpy(s.d_jobs, r->codec == B2_CODEC_H264 ? (const void *)s.h4.data() : (const void *)s.m2.data(), bytes, cudaMemcpyHostToDevice));
		}
		for (cudaEvent_t e : g.ev) cudaEventDestroy(e);
		g.ev.resize(g.steps.size() * 4);
		for (auto &e : g.ev) B2_CUDA_OK(cudaEventCreate(&e));
	}
	r->dirty = false;
	return 0;
}

extern "C" int b2_replay_run(b2_replay *r, int timed) {
	if (!r) return -1;
	B2_CUDA_OK(cudaSetDevice(r->dev));
	if (finalize(r) < 0) return -1;
	const int mb_w = r->width / 16, mb_h = r->height / 16;
	uint64_t launches = 0;
	B2_CUDA_OK(cudaEventRecord(r->ev_begin, r->stream));
	size_t n_steps = 0;
	for (Group &g : r->groups) {
		B2_CUDA_OK(cudaStreamWaitEvent(g.stream, r->ev_begin, 0));
		if (g.steps.size() > n_steps) n_steps = g.steps.size();
	}
	// step-major issue order: the groups' chains advance together
	for (size_t t = 0; t < n_steps; ++t) {
		for (Group &g : r->groups) {
			if (t >= g.steps.size()) continue;
			Step &s = g.steps[t];
			if (!s.d_jobs) continue;
			if (r->codec == B2_CODEC_MPEG2) {
				if (launch_mpeg2(g.stream, (const M2Job *)s.d_jobs, (int)s.m2.size(), mb_w, mb_h) < 0) return -1;
				launches++;
				continue;
			}
			const H4Job *jobs = (const H4Job *)s.d_jobs;
			int n = (int)s.h4.size();
			if (timed) B2_CUDA_OK(cudaEventRecord(g.ev[t * 4 + 0], g.stream));
			if (s.max_items && launch_h264_phase(g.stream, 0, jobs, n, mb_w, mb_h, g.sync_area, &launches, s.max_items, s.max_first) < 0) return -1;
			if (timed) B2_CUDA_OK(cudaEventRecord(g.ev[t * 4 + 1], g.stream));
			if (s.intra_dense && launch_h264_phase(g.stream, 1, jobs, n, mb_w, mb_h, g.sync_area, &launches) < 0) return -1;
			if (s.intra_sparse && launch_h264_phase(g.stream, 3, jobs, n, mb_w, mb_h, g.sync_area, &launches) < 0) return -1;
			if (timed) B2_CUDA_OK(cudaEventRecord(g.ev[t * 4 + 2], g.stream));
			if (s.derive_bs && launch_h264_phase(g.stream, 4, jobs, n, mb_w, mb_h, g.sync_area, &launches) < 0) return -1;
			if (s.deblock && launch_h264_phase(g.stream, 2, jobs, n, mb_w, mb_h, g.sync_area, &launches) < 0) return -1;
			if (timed) B2_CUDA_OK(cudaEventRecord(g.ev[t * 4 + 3], g.stream));

		}
	}
	for (Group &g : r->groups) {
		B2_CUDA_OK(cudaEventRecord(g.done, g.stream));
		B2_CUDA_OK(cudaStreamWaitEvent(r->stream, g.done, 0));
	}
	B2_CUDA_OK(cudaEventRecord(r->ev_end, r->stream));
	r->stats.launches = launches;
	r->timed_pending = timed != 0;
	return 0;
}

extern "C" int b2_replay_sync(b2_replay *r) {
	if (!r) return -1;
	B2_CUDA_OK(cudaSetDevice(r->dev));
	B2_CUDA_OK(cudaStreamSynchronize(r->stream));
	B2_CUDA_OK(cudaEventElapsedTime(&r->stats.total_ms, r->ev_begin, r->ev_end));
	if (r->timed_pending) {
		// Per-kernel device time.  With one group these are the launches' own durations.  With
		// several, launches of different groups overlap, so the event intervals are summed per
		// kernel and the run's wall time is attributed to the kernels in proportion to those sums.
		float acc[3] = {0, 0, 0};
		if (r->codec == B2_CODEC_H264) {
			for (Group &g : r->groups)
				for (size_t t = 0; t < g.steps.size(); ++t) {
					if (!g.steps[t].d_jobs) continue;
					for (int k = 0; k < 3; ++k) {
						float ms = 0;
						B2_CUDA_OK(cudaEventElapsedTime(&ms, g.ev[t * 4 + k], g.ev[t * 4 + k + 1]));
						acc[k] += ms;
					}
				}
			float sum = acc[0] + acc[1] + acc[2];
			if (r->groups.size() > 1 && sum > 0)
				for (int k = 0; k < 3; ++k) acc[k] = acc[k] / sum * r->stats.total_ms;
		} else {
			acc[0] = r->stats.total_ms;
		}
		for (int k = 0; k < 3; ++k) r->stats.kernel_ms[k] = acc[k];
		r->timed_pending = false;
	}
	return 0;
}

extern "C" int b2_replay_get_stats(b2_replay *r, b2_replay_stats *out) {
	if (!r || !out) return -1;
	*out = r->stats;
	return 0;
}

extern "C" int b2_replay_download(b2_replay *r, int lane, int idx, uint8_t *luma, uint8_t *chroma) {
	if (!r || (unsigned)lane >= (unsigned)r->lanes || (unsigned)idx >= (unsigned)r->pool || !luma || !chroma) return -1;
	B2_CUDA_OK(cudaSetDevice(r->dev));
	B2_CUDA_OK(cudaStreamSynchronize(r->stream));
	size_t ls = (size_t)r->width * r->height;
	const uint8_t *src = r->frames + r->frame_bytes * ((size_t)r->pool * lane + idx);
	B2_CUDA_OK(cudaMemcpy(luma, src, ls, cudaMemcpyDeviceToHost));
	B2_CUDA_OK(cudaMemcpy(chroma, src + ls, ls / 2, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int b2_replay_md5(b2_replay *r, int n, const int *lanes, const int *frame_idx, int crop_left, int crop_top, int width, int height, uint8_t *digests) {
	if (!r || n < 0 || !lanes || !frame_idx || !digests || crop_left < 0 || crop_top < 0 || width <= 0 || height <= 0 || (height & 1) || (crop_top & 1) ||
	    crop_left + width > r->width || crop_top + height > r->height) return -1;
	if (n == 0) return 0;
	std::vector<int> slots(n);
	for (int i = 0; i < n; ++i) {
		if ((unsigned)lanes[i] >= (unsigned)r->lanes || (unsigned)frame_idx[i] >= (unsigned)r->pool) return -1;
		slots[i] = r->pool * lanes[i] + frame_idx[i];
	}
	B2_CUDA_OK(cudaSetDevice(r->dev));
	int *d_slots = nullptr;
	uint8_t *d_out = nullptr;
	B2_CUDA_OK(cudaMalloc(&d_slots, sizeof(int) * n));
	B2_CUDA_OK(cudaMalloc(&d_out, (size_t)16 * n));
	int rc = 0;
	if (cudaMemcpyAsync(d_slots, slots.data(), sizeof(int) * n, cudaMemcpyHostToDevice, r->stream) != cudaSuccess ||
	    md5_pictures(r->stream, r->frames, r->frame_bytes, d_slots, n, r->width, r->height, crop_left, crop_top, width, height, d_out) < 0 ||
	    cudaMemcpyAsync(digests, d_out, (size_t)16 * n, cudaMemcpyDeviceToHost, r->stream) != cudaSuccess || cudaStreamSynchronize(r->stream) != cudaSuccess) {
		set_error("b2_replay_md5: %s", cudaGetErrorString(cudaGetLastError()));
		rc = -1;
	}
	cudaFree(d_slots);
	cudaFree(d_out);
	return rc;
}

extern "C" int b2_replay_mark(b2_replay *r, int which) {
	if (!r || (unsigned)which > 1) return -1;
	B2_CUDA_OK(cudaSetDevice(r->dev));
	B2_CUDA_OK(cudaEventRecord(r->ev_mark[which], r->stream));
	return 0;
}

extern "C" int b2_replay_elapsed_ms(b2_replay *r, float *ms) {
	if (!r || !ms) return -1;
	B2_CUDA_OK(cudaSetDevice(r->dev));
	B2_CUDA_OK(cudaEventSynchronize(r->ev_mark[1]));
	B2_CUDA_OK(cudaEventElapsedTime(ms, r->ev_mark[0], r->ev_mark[1]));
	return 0;
}
