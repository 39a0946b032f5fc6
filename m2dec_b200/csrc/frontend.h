// frontend.h — pieces shared by the two host front ends (mpeg2_front.cpp, h264_front.cpp):
// the caller-owned context blob of the m2d_func_table_t API and the hand-off of finished
// pictures to the reconstruction back end (device session and/or picture tap).
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>
#include "m2d_plugin.h"
#include "b2_recon.h"
#include "bits.h"

namespace b2 {

// What `new byte_t[context_size]` (reference src/app/m2decoder.h:184) holds.  The decoder
// proper lives on the heap so that it can own std::vectors; the reference API has no
// destroy hook, so b2_context_release() (or a second init) frees it.
struct CtxBlob {
	uint64_t magic;
	int codec;
	void *impl;
	dec_bits_t stream;
};
static const uint64_t kCtxMagic = 0x6232646563303031ULL;  // "b2dec001"

// Hand-off of pictures to the device (product path) and/or a tap (bench capture, tests).
class Backend {
public:
	Backend();
	~Backend();
	void set_tap(b2_picture_tap tap, void *arg, int device_off) { tap_ = tap; tap_arg_ = arg; device_off_ = device_off != 0; }
	void set_device(int dev) { cuda_device_ = dev; }
	// digest mode (b2_context_set_md5): frames are hashed on the device, fetch() copies nothing
	void set_md5(bool on) { md5_ = on; }
	bool md5() const { return md5_; }
	// the window and row pitch the application's MD5 writer would use for this stream's frames (the front end knows
	// them from the sequence header / SPS); must be called before configure()
	void set_md5_window(int stride, int left, int top, int w, int h) { md5_stride_ = stride; md5_left_ = left; md5_top_ = top; md5_w_ = w; md5_h_ = h; }
	int frame_md5(const m2d_frame_t *frame, uint8_t digest[16]);
	// the application's frame buffers (set_frames): frame i of the device pool is copied out to them
	void bind_frames(const m2d_frame_t *frames, int n) {
		n_host_ = n < 64 ? n : 64;
		for (int i = 0; i < n_host_; ++i) { host_luma_[i] = frames[i].luma; host_chroma_[i] = frames[i].chroma; }
	}
	// (re)create the device frame pool; called from set_frames (after bind_frames)
	int configure(int codec, int width, int height, int nframes);
	// same, but keeps the existing session when nothing changed (called at every picture start)
	int ensure(int codec, int width, int height, int nframes) {
		if (device_off_ || dead_ || (sess_ && codec == codec_ && width == width_ && height == height_ && nframes == nframes_)) return 0;
		return configure(codec, width, height, nframes);
	}
	int submit(int codec, const void *picture);
	// make frame idx visible in the caller's host buffers
	int fetch(int idx, uint8_t *luma, uint8_t *chroma);
	void release();
	bool device_off() const { return device_off_; }
	bool has_tap() const { return tap_ != nullptr; }
	// b2_context_set_host_strengths: the H.264 front end derives the deblocking strengths itself
	void set_host_bs(bool on) { host_bs_ = on; }
	bool host_bs() const { return host_bs_; }
	const b2_session *session() const { return sess_; }
	// true once a CUDA session could not be created: the decoder then refuses to go on
	bool dead() const { return dead_; }
private:
	b2_picture_tap tap_;
	void *tap_arg_;
	bool device_off_;
	bool dead_;
	int cuda_device_;
	b2_session *sess_;
	int codec_, width_, height_, nframes_;
	int n_host_;
	uint8_t *host_luma_[64], *host_chroma_[64];
	bool host_bs_ = false;
	bool md5_;
	int md5_stride_, md5_left_, md5_top_, md5_w_, md5_h_;
};

}  // namespace b2
