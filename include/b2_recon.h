/* b2_recon.h — C ABI between the host bit-stream front end and the sm_100a reconstruction
 * kernels of m2dec-b200: the per-picture structure-of-arrays (SoA) records the parse emits,
 * and the device-side entry points that consume them.
 *
 * Nothing here exists in the reference: in wujs/m2dec reconstruction is called
 * synchronously, block by block, from inside the parse.  Each record type
 * below names the reference call site(s) whose arguments it carries, so a maintainer can see
 * which C call each field replaces.
 *
 * Plain C, plain pointers and sizes; no torch / CUDA types in any signature.
 */
#ifndef B2_RECON_H_
#define B2_RECON_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { B2_CODEC_MPEG2 = 1, B2_CODEC_H264 = 2 };

/* ------------------------------------------------------------------------------------ */
/* MPEG-2 (frame pictures).  One b2_m2_mb per macroblock of the picture, raster order.    */
/* ------------------------------------------------------------------------------------ */

/* b2_m2_mb.flags */
enum {
	B2_M2_FWD = 0x01,       /* predict from the forward reference   (mpeg2.h:175 MB_FORWARD)   */
	B2_M2_BWD = 0x02,       /* predict from the backward reference  (mpeg2.h:176 MB_BACKWARD)  */
	B2_M2_INTRA = 0x04,     /* intra: IDCT result is stored         (mpeg2.cpp:1136-1153)      */
	B2_M2_FIELD_MC = 0x08,  /* two 16x8 field predictions per direction (mpeg2.cpp:1291-1307)  */
	B2_M2_FIELD_DCT = 0x10  /* dct_type = 1: luma blocks interleave lines (mpeg2.cpp:1120)     */
};

/* Carries the arguments of m2d_motion_comp (mpeg2.cpp:1277-1308), m2d_skip_mb_P/_B
 * (mpeg2.cpp:740-808), m2d_copy_slice (mpeg2.cpp:715-733) and the six m2d_idct_* calls of
 * one macroblock (mpeg2.cpp:1136-1153, 1343-1353).  Skipped / not-coded / missing-slice
 * macroblocks are emitted by the host as ordinary predictions (flags = FWD [| BWD], mv as
 * the reference resolves them) with cbp = 0, so the device needs no special case.          */
typedef struct {
	uint8_t flags;
	uint8_t cbp;           /* bit (5 - k) set <=> block k (Y0 Y1 Y2 Y3 Cb Cr) has a coef record  */
	uint8_t field_sel;     /* bit (dir*2 + f): motion_vertical_field_select of prediction f      */
	uint8_t reserved;
	uint32_t coef_idx;     /* index, in 64-coefficient blocks, of this MB's first coef record    */
	int16_t mv[2][2][2];   /* [dir 0=fwd 1=bwd][field f][x,y] half-pel, as m2d_mv() returns them  */
} b2_m2_mb;                /* 24 bytes */

typedef struct {
	int32_t mb_w, mb_h;            /* picture size in macroblocks; frame stride = 16*mb_w          */
	int32_t pic_type;              /* 1 I, 2 P, 3 B                                                */
	int32_t dst, fwd, bwd;         /* frame-pool indices (diff_to_ref of mpeg2.cpp:153-194)        */
	uint32_t n_mbs;                /* = mb_w * mb_h                                                */
	uint32_t n_blocks;             /* number of 64-coefficient records                             */
	const b2_m2_mb *mbs;
	const int16_t *coefs;          /* dequantised, saturated, mismatch-corrected (mpeg2.cpp:1020-1097), raster 8x8 */
} b2_m2_picture;

/* ------------------------------------------------------------------------------------ */
/* Picture tap: a hook that sees every picture the front end emits, before it is queued   */
/* for the device.  Used by bench.py to capture SoA for kernel replay and by tests/ to    */
/* hand the same records to the CPU oracle.  It never produces pixels itself.             */
/* ------------------------------------------------------------------------------------ */
typedef int (*b2_picture_tap)(void *arg, int codec, const void *picture /* b2_m2_picture* or b2_h264_picture* */);

/* Attach a tap to a decoder context created through m2d_gpu_func / h264d_gpu_func.
 * If device_off != 0 the context runs parse-only: nothing is sent to a GPU and frame-out
 * returns the caller's buffers untouched (whatever the tap's owner put there).            */
int b2_context_set_tap(void *ctx, b2_picture_tap tap, void *arg, int device_off);
/* Select the CUDA device ordinal used by a context (default 0, or $B2_CUDA_DEVICE).       */
int b2_context_set_device(void *ctx, int cuda_device);
/* Release device resources held by a context (the reference API has no destroy hook).     */
int b2_context_release(void *ctx);

/* ------------------------------------------------------------------------------------ */
/* Device side.  A session owns one CUDA stream, a pool of device-resident NV12 frames    */
/* and a pinned staging ring for SoA uploads.                                             */
/* ------------------------------------------------------------------------------------ */
typedef struct b2_session b2_session;

int b2_device_count(void);                           /* <0: CUDA unavailable                    */
const char *b2_last_error(void);
int b2_session_create(int cuda_device, int codec, int width, int height, int nframes, b2_session **out);
void b2_session_destroy(b2_session *s);
/* Asynchronous: copies the picture's SoA into pinned staging, enqueues H2D + kernels.     */
int b2_session_submit_mpeg2(b2_session *s, const b2_m2_picture *pic);
/* Waits for everything queued so far, then copies frame `idx` to host memory.             */
int b2_session_download(b2_session *s, int idx, uint8_t *luma, uint8_t *chroma);
int b2_session_upload(b2_session *s, int idx, const uint8_t *luma, const uint8_t *chroma);
int b2_session_sync(b2_session *s);
/* Number of kernel launches issued by this session so far.                                */
uint64_t b2_session_launches(const b2_session *s);

#ifdef __cplusplus
}
#endif
#endif /* B2_RECON_H_ */
