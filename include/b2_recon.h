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
#include "m2d_plugin.h"

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
/* H.264 (frame pictures, 4:2:0, 8 bit).  Per picture: one b2_h264_mb and one            */
/* b2_h264_dbk per macroblock in raster order, a pool of motion-compensation records,     */
/* and a pool of coefficient records (16 int16 each, levels BEFORE dequantisation).       */
/* ------------------------------------------------------------------------------------ */

enum { B2_H4_INTER = 0, B2_H4_I4x4 = 1, B2_H4_I8x8 = 2, B2_H4_I16x16 = 3, B2_H4_IPCM = 4 };
/* b2_h264_mb.flags: transform_size_8x8_flag; with mv_on_device: P_Skip (the motion vector is inferred, 8.4.1.1) and,
 * in bits 4..7, which neighbouring macroblocks (bit 4 left, 5 top, 6 top-right, 7 top-left) are available to the
 * motion vector prediction (same slice, already decoded: 6.4.8)                                                   */
enum { B2_H4_T8x8 = 0x01, B2_H4_PSKIP = 0x02, B2_H4_NB_SHIFT = 4 };

/* b2_h264_mb.nz — which coefficient records exist, bit order == record order:
 *   bit 0        Intra16x16 luma DC (16 levels, raster 4x4 of the DC matrix)
 *   bits 1..16   luma 4x4 block k = bit-1 (k in decode order);
 *                an 8x8-transform block i owns bits 1+4i..4+4i (64 levels, raster 8x8)
 *   bit 17       chroma DC (one record: Cb c0..c3 then Cr c0..c3, raster 2x2)
 *   bits 18..21  Cb AC blocks 0..3 (raster 2x2 of 4x4 blocks), bits 22..25 Cr AC
 * I_PCM: 12 records hold the 384 raw sample bytes (256 Y, 64 Cb, 64 Cr), nz = 0xfff.
 * Record index of bit b = coef_idx + popcount(nz & ((1 << b) - 1)).
 * Levels are stored in raster order inside a record (position 4*y + x), i.e. the inverse
 * zig-zag (h264.cpp:1927-1994) is already applied; dequantisation (h264.cpp:2005-2022,
 * 964-1054) is done on the device from qp_y / qp_cb / qp_cr.                               */
typedef struct {
	uint8_t type;           /* B2_H4_*                                                            */
	uint8_t flags;
	uint8_t qp_y, qp_cb, qp_cr;  /* QPs in force when this MB's residual was dequantised (set_qp, h264.cpp:1092-1119) */
	uint8_t intra_avail;    /* bit0 left, bit1 top, bit2 top-right, bit3 top-left: MB-level availability as used by the predictors (h264.cpp:9704-9715, constrained-intra mask 3271-3273) */
	uint8_t i16_mode;       /* Intra16x16 prediction mode 0 V, 1 H, 2 DC, 3 plane                 */
	uint8_t chroma_mode;    /* intra_chroma_pred_mode 0 DC, 1 H, 2 V, 3 plane                     */
	uint32_t nz;
	uint32_t coef_idx;      /* first coefficient record of this MB                                */
	uint32_t mc_idx;        /* first motion-compensation record of this MB                        */
	uint8_t mc_cnt;
	uint8_t intra_avail_c;  /* availability mask the chroma predictor received (differs from intra_avail only for a no-residual Intra4x4 MB under constrained_intra_pred, h264.cpp:3281 vs 3292) */
	uint16_t nnz4x4;        /* bit 4 * y + x: luma 4x4 block (x, y) counts as "has coefficients" in the deblocking-strength derivation (the parse's total_coeff / coded_block_flag state as str_previous_coef sees it, h264.cpp:7119-7322); used when bs_on_device is set */
	uint8_t i4_modes[8];    /* Intra4x4 (16 nibbles) / Intra8x8 (4 nibbles) prediction modes, block k in nibble k */
} b2_h264_mb;               /* 32 bytes */

/* One call of mb->inter_pred (h264.h:396; 14 call sites), with the
 * weighted-prediction parameters the call would have looked up resolved by the host
 * (inter_pred_weighted1 / _weighted2, h264.cpp:6979-6999, 7101-7117):
 *   one list : pred = ((p * w + 2^(logWD-1)) >> logWD) + o        (logWD = 0: p * w + o)
 *   two lists: pred = ((p0*w0 + p1*w1 + 2^logWD) >> (logWD+1)) + ((o0 + o1 + 1) >> 1)
 * Default prediction is w = 1, o = 0, logWD = 0; implicit bi-prediction is w0 + w1 = 64, logWD = 5.  */
typedef struct {
	uint8_t pos;            /* (y4 << 2) | x4: top-left 4x4 block of the partition inside the MB  */
	uint8_t size;           /* (h4 << 4) | w4: partition size in 4-pel units (1, 2 or 4)          */
	int8_t ref_frame[2];    /* frame-pool index for list 0 / 1 (refs[lx][ref_idx].frame_idx, h264.cpp:6740); -1 = list unused */
	int8_t ref_idx[2];      /* reference index per list (-1 unused)                               */
	uint8_t log2wd_y, log2wd_c;
	int16_t mv[2][2];       /* [list][x,y] quarter-pel                                            */
	int8_t w[2][3];         /* [list][Y, Cb, Cr] weight                                           */
	int8_t o[2][3];         /* [list][Y, Cb, Cr] offset                                           */
	uint8_t reserved[4];
} b2_h264_mc;               /* 32 bytes */

/* mv_on_device (SURVEY.md 8f rank 3, P macroblocks of Baseline-profile streams): one partition of a P macroblock
 * as the parse read it -- the motion vector DIFFERENCE, not the vector.  The device predicts the vectors
 * (8.4.1.3; reference calc_mv*, h264.cpp:6669-6724, 7379-7604) in a macroblock wavefront, writes the
 * b2_h264_mc record of every partition (default weights) and builds the inter kernel's work list.
 * b2_h264_mb.mc_idx / mc_cnt index the partitions (and the motion records the device makes of them, 1:1). */
typedef struct {
	uint8_t pos;            /* (y4 << 2) | x4                                                        */
	uint8_t size;           /* (h4 << 4) | w4                                                        */
	int8_t ref_idx;         /* refIdxL0                                                              */
	int8_t ref_frame;       /* frame-pool index of RefPicList0[ref_idx]                              */
	int16_t mvd[2];         /* mvd_l0 (x, y), quarter samples; 0 for P_Skip                          */
} b2_h264_part;             /* 8 bytes */

/* b2_h264_dbk.flags */
enum {
	B2_DBK_LEFT_BS4 = 0x01,   /* left MB edge has bS 4 (str4_horiz, h264.h:344-348)               */
	B2_DBK_TOP_BS4 = 0x02,    /* top MB edge has bS 4 (str4_vert)                                 */
	B2_DBK_OFF = 0x04,        /* disable_deblocking_filter_idc == 1 for this MB's slice           */
	B2_DBK_LEFT_EDGE = 0x08,  /* left MB edge is filtered (exists and not excluded by idc == 2)   */
	B2_DBK_TOP_EDGE = 0x10    /* top MB edge is filtered                                          */
};

/* Deblocking input for one MB: the reference's deblock_info_t (h264.h:344-348) restated.
 * bs_v: the four vertical edges (filtered horizontally; reference str_horiz), byte e = edge
 * e (0 = MB edge), 2 bits per 4-line segment, top segment in the low bits: 0 none, 1..3 = bS
 * (3 also stands for bS 4 on an MB edge when the BS4 flag is set).  bs_h likewise for the
 * four horizontal edges, left segment in the low bits.                                      */
typedef struct {
	uint32_t bs_v;
	uint32_t bs_h;
	int8_t qp_y, qp_cb, qp_cr;   /* as the reference stores them for the filter (I_PCM: 0 and qpc - qp, h264.cpp:4749-4751) */
	uint8_t flags;
	int8_t alpha_off, beta_off;  /* FilterOffsetA / FilterOffsetB (= 2 * slice_*_offset_div2)          */
	uint8_t bs_mode;             /* bs_on_device: bit 3 = the macroblock belongs to a B slice; bits 0..2 = which of its edges (0 macroblock edges, 1 centre edges, 2 other inner edges) take the reference's reduced same-picture comparison (enum below) */
	uint8_t reserved;
} b2_h264_dbk;              /* 16 bytes */
enum { B2_BS_REDUCED_MB_EDGE = 0x01, B2_BS_REDUCED_CENTRE = 0x02, B2_BS_REDUCED_INNER = 0x04, B2_BS_B_SLICE = 0x08 };

typedef struct {
	int32_t mb_w, mb_h;
	int32_t dst;               /* frame-pool index being reconstructed                                */
	int32_t deblock;           /* != 0 if any MB of the picture has the loop filter enabled          */
	uint32_t n_mbs, n_mcs, n_coef_recs;
	uint32_t n_intra;          /* number of intra macroblocks (0: the intra wavefront is skipped)     */
	const b2_h264_mb *mbs;
	const b2_h264_dbk *dbk;
	const b2_h264_mc *mcs;
	const int16_t *coefs;
	/* Work list of the inter kernel: one item per vertical strip of a motion-compensation record (8
	 * luma samples wide, or 4 for 4-wide partitions and for records that need the centre half sample
	 * j; full height of the partition), sorted by (class, height) so that a warp runs one code path.
	 * bits 0..19 macroblock index, 20..23 record index inside the macroblock, 24..25 strip index inside
	 * the record, 26..29 class = interpolation class of the first used list (0 integer, 1 horizontal,
	 * 2 vertical, 3 diagonal, 4 / 5 centre column / row) + 6 if the strip is 8 samples wide.           */
	const uint32_t *items;
	uint32_t n_items;
	/* The first n_items_first items are the FIRST-list passes of bi-predicted strips (class = that of
	 * list 0; the kernel writes their prediction to the destination in a launch of its own); the rest
	 * is one item per strip for its last (or only) list, class = that list's, which combines with the
	 * first pass, adds the residual and stores.  Both parts are sorted by (class, height, edge).      */
	uint32_t n_items_first;
	/* SURVEY.md 8f rank 2.  != 0: bs_v / bs_h and the BS4 flags of the b2_h264_dbk records are NOT filled in; the
	 * device derives them (h264_bs_kernel) from the motion records, b2_h264_mb.nnz4x4 and b2_h264_dbk.bs_mode, as
	 * the reference's parse does (store_info_inter*, str_mv_calc*, str_previous_coef: h264.cpp:7119-7322, 7451-7604,
	 * 8796-9400).  qp_*, alpha_off / beta_off and the flags B2_DBK_OFF / LEFT_EDGE / TOP_EDGE always come from the host. */
	uint32_t bs_on_device;
	/* SURVEY.md 8f rank 3.  != 0: mcs and items are NULL (n_mcs = 0); `parts` holds n_parts partitions, the device
	 * derives vectors, motion records and work list itself (h264_mv_kernel, h264_items_kernel).  n_items is then an
	 * UPPER BOUND of the number of work items (one per 4 luma columns of every partition); requires bs_on_device.
	 * Only pictures of P / I slices without weighted prediction qualify (the front end uses it for Baseline streams). */
	uint32_t mv_on_device;
	uint32_t n_parts;
	const b2_h264_part *parts;
} b2_h264_picture;

/* ------------------------------------------------------------------------------------ */
/* Picture tap: a hook that sees every picture the front end emits, before it is queued   */
/* for the device.  Used by bench.py to capture SoA for kernel replay and by tests/ to    */
/* hand the same records to the CPU oracle.  It never produces pixels itself.             */
/* ------------------------------------------------------------------------------------ */
typedef int (*b2_picture_tap)(void *arg, int codec, const void *picture /* b2_m2_picture* or b2_h264_picture* */);

/* Structural check of a picture's records against a frame pool of `nframes` frames: every index inside
 * its array, every enumerated field inside its range (the kernels check nothing).  0 = sound, < 0 = not
 * (b2_last_error() says which record).  b2_session_submit_* and b2_replay_add_* run it on every picture;
 * it is exported for hosts that build the records themselves.                                          */
int b2_h264_picture_check(const b2_h264_picture *picture, int nframes);
int b2_mpeg2_picture_check(const b2_m2_picture *picture, int nframes);

/* Attach a tap to a decoder context created through m2d_gpu_func / h264d_gpu_func.
 * If device_off != 0 the context runs parse-only: nothing is sent to a GPU and frame-out
 * returns the caller's buffers untouched (whatever the tap's owner put there).            */
int b2_context_set_tap(void *ctx, b2_picture_tap tap, void *arg, int device_off);
/* Select the CUDA device ordinal used by a context (default 0, or $B2_CUDA_DEVICE).       */
int b2_context_set_device(void *ctx, int cuda_device);
/* Release device resources held by a context (the reference API has no destroy hook).     */
int b2_context_release(void *ctx);
/* Digest mode behind the plugin table (b2_session_set_md5 above): call after init(), before the first       */
/* decode_picture().  peek / get_decoded_frame then return frames whose host buffers are NOT filled;         */
/* b2_context_frame_md5() returns the MD5 the reference's `-O` writer would print for that frame (window and */
/* pitch taken from the stream's headers exactly as the front end fills m2d_frame_t).  The digest of a frame */
/* stays valid until the decoder reuses the frame buffer, as the pixels do in the reference; handing         */
/* set_frames() a few frames more than get_info() asks for lets the application collect digests a few        */
/* pictures late without ever waiting for a hash.                                           */
int b2_context_set_md5(void *ctx, int on);
/* H.264 deblocking strengths: by default the front end leaves them to the device        */
/* (b2_h264_picture.bs_on_device) whenever nobody reads the records on the host; on != 0 makes it derive them   */
/* itself, as it does for a picture tap.  Call after init(), before the first decode_picture().               */
int b2_context_set_host_strengths(void *ctx, int on);
int b2_context_frame_md5(void *ctx, const m2d_frame_t *frame, uint8_t digest[16]);

/* ------------------------------------------------------------------------------------ */
/* Device side.  A session owns one CUDA stream, a pool of device-resident NV12 frames    */
/* and a pinned staging ring for SoA uploads.                                             */
/* ------------------------------------------------------------------------------------ */
typedef struct b2_session b2_session;

int b2_device_count(void);                           /* <0: CUDA unavailable                    */
const char *b2_last_error(void);
int b2_session_create(int cuda_device, int codec, int width, int height, int nframes, b2_session **out);
void b2_session_destroy(b2_session *s);
/* A destroyed session is parked whole (streams, events, frame pool, staging slots; at most 64 sessions) for the
 * next b2_session_create with the same device, codec, size and frame count, which gets it back in its initial state
 * with black frames; what is not parked leaves its device / page-locked buffers on a process-wide free list (at
 * most 4 GiB of each kind is kept) so that the next session does not pay cudaMalloc.  This frees all of it.
 * B2_NO_SESSION_POOL=1 in the environment switches the parking off.                                             */
void b2_cache_trim(void);
/* Asynchronous: copies the picture's SoA into pinned staging, enqueues H2D + kernels.     */
int b2_session_submit_mpeg2(b2_session *s, const b2_m2_picture *pic);
int b2_session_submit_h264(b2_session *s, const b2_h264_picture *pic);
/* Every submitted picture goes through b2_*_picture_check (about 0.2 ms of host time for a 1080p picture).  A
 * producer whose records are sound by construction may switch it off: the library's own front ends do (their
 * records are checked by tests/ on every test stream and under fuzzing; B2_CHECK_RECORDS=1 keeps it on).        */
int b2_session_set_record_check(b2_session *s, int on);
/* Tell the session where the application keeps frame `idx` (the m2d_frame_t of set_frames,  */
/* reference m2d.h:66-75).  When these buffers are page-locked (b2_host_alloc below, or the    */
/* application's own cudaHostAlloc / cudaHostRegister) every picture is copied out as soon as  */
/* its kernels are queued and b2_session_download() only waits for that copy; pageable         */
/* buffers keep the copy-on-demand behaviour.  The plugin front ends call this from            */
/* set_frames().                                                                               */
int b2_session_bind_host_frame(b2_session *s, int idx, uint8_t *luma, uint8_t *chroma);
/* Page-locked host memory for the application's frame buffers (replaces the `new byte_t[]` of */
/* reference src/app/frames.h), usable from every device; cuda_device < 0: the calling         */
/* thread's current device.  NULL without a CUDA device.                                       */
void *b2_host_alloc(size_t bytes, int cuda_device);
void b2_host_free(void *p);
/* Makes frame `idx` visible in host memory: waits for the picture that wrote it (not for    */
/* pictures queued behind it) and copies it out unless it was bound and is there already.   */
int b2_session_download(b2_session *s, int idx, uint8_t *luma, uint8_t *chroma);
int b2_session_upload(b2_session *s, int idx, const uint8_t *luma, const uint8_t *chroma);
/* Digest mode: what the reference application's MD5 writer computes per output frame  */
/* (`h264dec -O`: src/app/filewrite.h:91-124 FileWriterMd5::writeframe over write_cropping, src/app/md5.c)   */
/* is computed on the device instead -- the cropped luma rows, then the cropped rows of the interleaved CbCr */
/* plane, both addressed with row pitch `stride` (the application's m2d_frame_t.width).  While it is on,     */
/* every submitted picture is hashed behind its kernels and only the 16-byte digest is copied to the host;   */
/* frames are no longer copied into the bound host buffers.  A frame is not overwritten by a later picture   */
/* before its hash has read it.  b2_session_md5() waits for the digest of the picture last decoded into      */
/* frame `idx` (a hash of a 1080p frame takes ~30 ms -- MD5 is one serial chain -- so an application that    */
/* wants throughput asks for a frame's digest a few pictures after the frame came out).                      */
int b2_session_set_md5(b2_session *s, int on, int stride, int crop_left, int crop_top, int width, int height);
int b2_session_md5(b2_session *s, int idx, uint8_t digest[16]);
int b2_session_sync(b2_session *s);
/* Number of kernel launches issued by this session so far.                                */
uint64_t b2_session_launches(const b2_session *s);

/* Host<->device traffic of a session so far: bytes[0] = H2D, bytes[1] = D2H.                 */
int b2_session_traffic(const b2_session *s, uint64_t bytes[2]);
/* Same, for the session owned by a decoder context (0 when the context is parse-only);
 * counts[2] = kernel launches.                                                              */
int b2_context_stats(void *ctx, uint64_t counts[3]);

/* ------------------------------------------------------------------------------------ */
/* Kernel replay: `lanes` independent streams (one frame pool each) whose pictures are    */
/* uploaded once and stay resident in HBM; one run reconstructs picture #t of every lane   */
/* in a single batched launch per kernel, for t = 0 .. steps-1.  This is the throughput    */
/* path (bench.py `value`, roofline): no host work and no PCIe traffic inside a run.       */
/* ------------------------------------------------------------------------------------ */
typedef struct b2_replay b2_replay;

typedef struct {
	uint64_t pictures;        /* pictures reconstructed per run                                  */
	uint64_t launches;        /* kernel launches per run                                         */
	/* algorithmic bytes per run, split by kernel: [0] prediction+residual of   */
	/* inter MBs (MPEG-2: the whole fused kernel), [1] intra wavefront, [2] deblock wavefront   */
	uint64_t alg_bytes[3];
	/* device time of the last timed run per kernel, milliseconds (CUDA events on the replay   */
	/* stream), and of the whole run                                                           */
	float kernel_ms[3];
	float total_ms;
} b2_replay_stats;

int b2_replay_create(int cuda_device, int codec, int width, int height, int lanes, int pool_frames, b2_replay **out);
void b2_replay_destroy(b2_replay *r);
/* Append a picture to a lane's sequence (SoA copied to the device now, untimed).            */
int b2_replay_add_h264(b2_replay *r, int lane, const b2_h264_picture *pic);
int b2_replay_add_mpeg2(b2_replay *r, int lane, const b2_m2_picture *pic);
/* Enqueue one full run on the replay stream; with timed != 0 per-kernel events are recorded. */
int b2_replay_run(b2_replay *r, int timed);
int b2_replay_sync(b2_replay *r);
/* Record a CUDA event on the replay stream (which = 0 begin, 1 end) / read the elapsed device
 * time between the two after a sync: brackets any number of runs for whole-job timing.       */
int b2_replay_mark(b2_replay *r, int which);
int b2_replay_elapsed_ms(b2_replay *r, float *ms);
int b2_replay_get_stats(b2_replay *r, b2_replay_stats *out);
int b2_replay_download(b2_replay *r, int lane, int idx, uint8_t *luma, uint8_t *chroma);
/* MD5 of n reconstructed pictures on the device, as the reference application hashes an output   */
/* frame for `h264dec -O` (src/app/filewrite.h:11-29, 91-124; src/app/md5.c): the cropped luma    */
/* rows (window crop_left, crop_top, width x height), then the cropped rows of the interleaved    */
/* CbCr plane.  Picture i is frame frame_idx[i] of lane lanes[i]; digests receives 16 bytes per   */
/* picture.  Waits for the runs queued so far.                                                    */
int b2_replay_md5(b2_replay *r, int n, const int *lanes, const int *frame_idx, int crop_left, int crop_top, int width, int height, uint8_t *digests);

#ifdef __cplusplus
}
#endif
#endif /* B2_RECON_H_ */
