"""ctypes view of libm2dec_b200.so (include/m2d_plugin.h, include/b2_recon.h)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B2_LIB: another build of the library (kernel experiments: tools/variants.py)
LIB_PATH = os.environ.get("B2_LIB") or os.path.join(HERE, "libm2dec_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "m2dec_b200: %s is missing - run `python -m m2dec_b200.build` (or __graft_entry__.build()). "
        "There is no fallback implementation." % LIB_PATH)

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


class FrameT(C.Structure):          # m2d_frame_t
    _fields_ = [("luma", C.c_void_p), ("chroma", C.c_void_p), ("id", C.c_void_p), ("cnt", C.c_int32),
                ("width", C.c_int16), ("height", C.c_int16), ("crop", C.c_int16 * 4)]


class InfoT(C.Structure):           # m2d_info_t
    _fields_ = [("src_width", C.c_int16), ("src_height", C.c_int16), ("disp_width", C.c_int16),
                ("disp_height", C.c_int16), ("frame_num", C.c_int16), ("crop", C.c_int16 * 4),
                ("additional_size", C.c_int)]


HEADER_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)
REFILL_CB = C.CFUNCTYPE(C.c_int, C.c_void_p)
TAP_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)


class FuncTable(C.Structure):       # m2d_func_table_t
    _fields_ = [("context_size", C.c_size_t),
                ("init", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, HEADER_CB, C.c_void_p)),
                ("stream_pos", C.CFUNCTYPE(C.c_void_p, C.c_void_p)),
                ("get_info", C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(InfoT))),
                ("set_frames", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(FrameT), C.c_void_p, C.c_int)),
                ("decode_picture", C.CFUNCTYPE(C.c_int, C.c_void_p)),
                ("peek_decoded_frame", C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(FrameT), C.c_int)),
                ("get_decoded_frame", C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(FrameT), C.c_int))]


class M2MB(C.Structure):            # b2_m2_mb
    _fields_ = [("flags", C.c_uint8), ("cbp", C.c_uint8), ("field_sel", C.c_uint8), ("reserved", C.c_uint8),
                ("coef_idx", C.c_uint32), ("mv", C.c_int16 * 8)]


class M2Picture(C.Structure):       # b2_m2_picture
    _fields_ = [("mb_w", C.c_int32), ("mb_h", C.c_int32), ("pic_type", C.c_int32),
                ("dst", C.c_int32), ("fwd", C.c_int32), ("bwd", C.c_int32),
                ("n_mbs", C.c_uint32), ("n_blocks", C.c_uint32),
                ("mbs", C.POINTER(M2MB)), ("coefs", C.POINTER(C.c_int16))]


def func_table(name):
    """`m2d_gpu_func` / `h264d_gpu_func`: exported `const m2d_func_table_t *const`."""
    p = C.POINTER(FuncTable).in_dll(lib, name)
    if not p:
        raise RuntimeError("%s is not available in this build" % name)
    return p.contents


lib.dec_bits_set_data.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
lib.dec_bits_set_data.restype = C.c_int
lib.dec_bits_set_callback.argtypes = [C.c_void_p, REFILL_CB, C.c_void_p]
lib.dec_bits_set_callback.restype = None
lib.b2_context_set_tap.argtypes = [C.c_void_p, TAP_CB, C.c_void_p, C.c_int]
lib.b2_context_set_device.argtypes = [C.c_void_p, C.c_int]
lib.b2_context_release.argtypes = [C.c_void_p]
lib.b2_context_set_md5.argtypes = [C.c_void_p, C.c_int]
lib.b2_context_set_host_strengths.argtypes = [C.c_void_p, C.c_int]
lib.b2_context_frame_md5.argtypes = [C.c_void_p, C.POINTER(FrameT), C.c_void_p]
lib.b2_device_count.restype = C.c_int
lib.b2_host_alloc.argtypes = [C.c_size_t, C.c_int]
lib.b2_host_alloc.restype = C.c_void_p
lib.b2_host_free.argtypes = [C.c_void_p]
lib.b2_host_free.restype = None
lib.b2_last_error.restype = C.c_char_p


class ReplayStats(C.Structure):     # b2_replay_stats
    _fields_ = [("pictures", C.c_uint64), ("launches", C.c_uint64), ("alg_bytes", C.c_uint64 * 3),
                ("kernel_ms", C.c_float * 3), ("total_ms", C.c_float)]


lib.b2_replay_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
lib.b2_replay_destroy.argtypes = [C.c_void_p]
lib.b2_replay_destroy.restype = None
lib.b2_replay_add_h264.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
lib.b2_replay_add_mpeg2.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
lib.b2_replay_run.argtypes = [C.c_void_p, C.c_int]
lib.b2_replay_sync.argtypes = [C.c_void_p]
lib.b2_replay_get_stats.argtypes = [C.c_void_p, C.POINTER(ReplayStats)]
lib.b2_replay_download.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.b2_replay_md5.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
lib.b2_context_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64 * 3)]
lib.b2_replay_mark.argtypes = [C.c_void_p, C.c_int]
lib.b2_replay_elapsed_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
lib.b2_h264_picture_check.argtypes = [C.c_void_p, C.c_int]
lib.b2_mpeg2_picture_check.argtypes = [C.c_void_p, C.c_int]
