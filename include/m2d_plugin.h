/* m2d_plugin.h — the drop-in boundary of m2dec-b200.
 *
 * These are the ABI types of the reference's decoder plugin table, restated so that an
 * application written against the reference's `m2d_func_table_t` (reference
 * src/lib/m2d.h:35-42 m2d_frame_t, :58-64 m2d_info_t, :66-75 m2d_func_table_t) links
 * against libm2dec_b200.so unchanged.  Member order and types are the ABI; nothing else
 * in this header comes from the reference.
 *
 * The bit-stream feeder (`dec_bits`) is part of that boundary: the application obtains it
 * with `stream_pos()` and hands data over with `dec_bits_set_data()` /
 * `dec_bits_set_callback()` (reference src/lib/bitio.h:57-61, used by
 * src/app/m2decoder.h:185-192 and src/app/h264dec.cpp:67-84).  Here `dec_bits` is an
 * opaque handle; its layout is private to libm2dec_b200.
 */
#ifndef M2D_PLUGIN_H_
#define M2D_PLUGIN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char byte_t;

typedef struct {
	uint8_t *luma;      /* W*H bytes, stride = W (caller-owned, 16-byte aligned)            */
	uint8_t *chroma;    /* NV12 interleaved CbCr, W*H/2 bytes, stride = W                   */
	void *id;
	int32_t cnt;        /* POC (H.264) or temporal_reference (MPEG-2)                        */
	int16_t width, height;
	int16_t crop[4];    /* left, right, top, bottom                                          */
} m2d_frame_t;

typedef struct {
	int16_t src_width, src_height;
	int16_t disp_width, disp_height;
	int16_t frame_num;
	int16_t crop[4];
	int additional_size;
} m2d_info_t;

typedef struct dec_bits_t dec_bits;

typedef struct {
	size_t context_size;
	int (*init)(void *ctx, int dpb_max, int (*header_callback)(void *, void *), void *cb_arg);
	dec_bits *(*stream_pos)(void *ctx);
	int (*get_info)(void *ctx, m2d_info_t *info);
	int (*set_frames)(void *ctx, int num, m2d_frame_t *frames, uint8_t *second, int second_size);
	int (*decode_picture)(void *ctx);
	int (*peek_decoded_frame)(void *ctx, m2d_frame_t *frame, int bypass_dpb);
	int (*get_decoded_frame)(void *ctx, m2d_frame_t *frame, int bypass_dpb);
} m2d_func_table_t;

/* --- bit-stream feeder: same entry points as reference src/lib/bitio.h:57-61,72-74 --- */
int dec_bits_open(dec_bits *ths, void (*loadbytes_func)(dec_bits *, int bytes));
void dec_bits_close(dec_bits *ths);
void dec_bits_set_callback(dec_bits *ths, int (*refill_func)(void *), void *refill_arg);
int dec_bits_set_data(dec_bits *ths, const byte_t *buf, size_t buf_len, void *id);
const byte_t *dec_bits_current(dec_bits *ths);
const byte_t *dec_bits_tail(dec_bits *ths);
/* reference src/lib/m2d.h:80 (used by src/app/m2decoder.h:96-131) */
int m2d_next_start_code(const byte_t *org_src, int byte_len);

/* --- plugin tables: GPU-reconstruction replacements for the reference's
 *     m2d_func (src/lib/mpeg2.cpp:1800-1811) and h264d_func (src/lib/h264.cpp:12057-12068) --- */
extern const m2d_func_table_t *const m2d_gpu_func;
extern const m2d_func_table_t *const h264d_gpu_func;

#ifdef __cplusplus
}
#endif
#endif /* M2D_PLUGIN_H_ */
