// plugin.cpp — context-level helpers of the C ABI that are not part of the reference's
// m2d_func_table_t (tap, device selection, release).  See include/b2_recon.h.
#include "frontend.h"

namespace b2 {
Backend *mpeg2_backend(void *ctx);
void mpeg2_destroy(void *ctx);
Backend *h264_backend(void *ctx);
void h264_destroy(void *ctx);

static Backend *backend_of(void *ctx) {
	CtxBlob *b = (CtxBlob *)ctx;
	if (!b || b->magic != kCtxMagic) return nullptr;
	if (b->codec == B2_CODEC_MPEG2) return mpeg2_backend(ctx);
	if (b->codec == B2_CODEC_H264) return h264_backend(ctx);
	return nullptr;
}
}  // namespace b2

extern "C" int b2_context_set_tap(void *ctx, b2_picture_tap tap, void *arg, int device_off) {
	b2::Backend *be = b2::backend_of(ctx);
	if (!be) return -1;
	be->set_tap(tap, arg, device_off);
	return 0;
}

extern "C" int b2_context_set_device(void *ctx, int cuda_device) {
	b2::Backend *be = b2::backend_of(ctx);
	if (!be) return -1;
	be->set_device(cuda_device);
	return 0;
}

extern "C" int b2_context_set_md5(void *ctx, int on) {
	b2::Backend *be = b2::backend_of(ctx);
	if (!be) return -1;
	be->set_md5(on != 0);
	return 0;
}

extern "C" int b2_context_set_host_strengths(void *ctx, int on) {
	b2::Backend *be = b2::backend_of(ctx);
	if (!be) return -1;
	be->set_host_bs(on != 0);
	return 0;
}

extern "C" int b2_context_frame_md5(void *ctx, const m2d_frame_t *frame, uint8_t digest[16]) {
	b2::Backend *be = b2::backend_of(ctx);
	if (!be) return -1;
	return be->frame_md5(frame, digest);
}

extern "C" int b2_context_release(void *ctx) {
	b2::CtxBlob *b = (b2::CtxBlob *)ctx;
	if (!b || b->magic != b2::kCtxMagic) return -1;
	if (b->codec == B2_CODEC_MPEG2) b2::mpeg2_destroy(ctx);
	else if (b->codec == B2_CODEC_H264) b2::h264_destroy(ctx);
	dec_bits_close(&b->stream);
	b->magic = 0;
	return 0;
}

extern "C" int b2_context_stats(void *ctx, uint64_t counts[3]) {
	b2::Backend *be = b2::backend_of(ctx);
	if (!be || !counts) return -1;
	counts[0] = counts[1] = counts[2] = 0;
	if (be->session()) {
		b2_session_traffic(be->session(), counts);
		counts[2] = b2_session_launches(be->session());
	}
	return 0;
}
