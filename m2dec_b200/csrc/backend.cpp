// backend.cpp — picture hand-off from a front end to the device session / tap.
#include "frontend.h"
#include <stdlib.h>
#include <stdio.h>

namespace b2 {

Backend::Backend() : tap_(nullptr), tap_arg_(nullptr), device_off_(false), dead_(false), cuda_device_(0), sess_(nullptr), codec_(0), width_(0), height_(0), nframes_(0), n_host_(0),
                     md5_(false), md5_stride_(0), md5_left_(0), md5_top_(0), md5_w_(0), md5_h_(0) {
	const char *e = getenv("B2_CUDA_DEVICE");
	if (e) cuda_device_ = atoi(e);
}

Backend::~Backend() { release(); }

void Backend::release() {
	if (sess_) { b2_session_destroy(sess_); sess_ = nullptr; }
}

int Backend::configure(int codec, int width, int height, int nframes) {
	release();
	dead_ = false;
	codec_ = codec; width_ = width; height_ = height; nframes_ = nframes;
	if (device_off_) return 0;
	int err = b2_session_create(cuda_device_, codec, width, height, nframes, &sess_);
	if (err < 0) {
		// There is deliberately no CPU reconstruction to fall back to.
		fprintf(stderr, "m2dec_b200: cannot create CUDA reconstruction session: %s\n", b2_last_error());
		sess_ = nullptr;
		dead_ = true;
		return err;
	}
	// the front ends' records are sound by construction (and checked in tests/ on every stream, damaged ones included)
	static const bool keep_check = getenv("B2_CHECK_RECORDS") != nullptr;
	if (!keep_check) b2_session_set_record_check(sess_, 0);
	for (int i = 0; i < n_host_ && i < nframes; ++i) b2_session_bind_host_frame(sess_, i, host_luma_[i], host_chroma_[i]);
	if (md5_ && b2_session_set_md5(sess_, 1, md5_stride_, md5_left_, md5_top_, md5_w_, md5_h_) < 0) {
		fprintf(stderr, "m2dec_b200: digest mode: %s\n", b2_last_error());
		release();
		dead_ = true;
		return -1;
	}
	return err;
}

int Backend::submit(int codec, const void *picture) {
	int err = 0;
	if (tap_) err = tap_(tap_arg_, codec, picture);
	if (device_off_) return err;
	if (!sess_) return -1;
	int r = -1;
	if (codec == B2_CODEC_MPEG2) r = b2_session_submit_mpeg2(sess_, (const b2_m2_picture *)picture);
	else if (codec == B2_CODEC_H264) r = b2_session_submit_h264(sess_, (const b2_h264_picture *)picture);
	if (r < 0) fprintf(stderr, "m2dec_b200: submit failed: %s\n", b2_last_error());
	return r;
}

int Backend::frame_md5(const m2d_frame_t *frame, uint8_t digest[16]) {
	if (!frame || !digest || !sess_ || !md5_) return -1;
	for (int i = 0; i < n_host_; ++i)
		if (host_luma_[i] == frame->luma) return b2_session_md5(sess_, i, digest);
	return -1;
}

int Backend::fetch(int idx, uint8_t *luma, uint8_t *chroma) {
	if (device_off_ || md5_) return 0;
	if (!sess_) return -1;
	int r = b2_session_download(sess_, idx, luma, chroma);
	if (r < 0) fprintf(stderr, "m2dec_b200: frame download failed: %s\n", b2_last_error());
	return r;
}

}  // namespace b2
