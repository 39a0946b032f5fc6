"""Host-side driver mirroring the reference application classes on top of the plugin table.

`M2Decoder` follows reference src/app/m2decoder.h:22-223 (set_codec / SetFrames / decode /
decode_residual), `Frames` src/app/frames.h, and `frame_md5`/`crop_frame` the cropped NV12
writer of src/app/filewrite.h:11-29, 99-124 — so that tests read like `h264dec -O`.
"""
import ctypes as C
import hashlib

import numpy as np

from . import _native as L

MODE_MPEG2 = "mpeg2"
MODE_H264 = "h264"
_TABLES = {MODE_MPEG2: "m2d_gpu_func", MODE_H264: "h264d_gpu_func"}
MAX_FRAME_NUM = 16          # reference src/lib/mpeg2.h:112
H264D_MAX_FRAME_NUM = 64    # reference src/lib/h264.h


def _align16(a):
    off = (-a.ctypes.data) & 15
    return a[off:]


class Frames:
    """Caller-owned frame buffers (reference src/app/frames.h): 16-byte aligned luma W*H and
    NV12 chroma W*H/2."""

    def __init__(self, width, height, num, second_len, pinned=True, cuda_device=-1):
        self.width, self.height = width, height
        self.luma_len = width * height
        # one page-locked block for all frames (b2_host_alloc) when a CUDA device is there: finished
        # pictures are then copied straight into these buffers as soon as their kernels are queued
        fb = self.luma_len * 3 // 2
        self._pinned = L.lib.b2_host_alloc(fb * num, cuda_device) if pinned else None
        if self._pinned:
            blk = np.ctypeslib.as_array(C.cast(self._pinned, C.POINTER(C.c_uint8)), shape=(fb * num,))
            blk[:] = 0
            self.luma = [blk[i * fb:i * fb + self.luma_len] for i in range(num)]
            self.chroma = [blk[i * fb + self.luma_len:(i + 1) * fb] for i in range(num)]
        else:
            self.luma = [_align16(np.zeros(self.luma_len + 15, np.uint8)) for _ in range(num)]
            self.chroma = [_align16(np.zeros(self.luma_len // 2 + 15, np.uint8)) for _ in range(num)]
        self.second = np.zeros(max(second_len, 1), np.uint8)
        self.array = (L.FrameT * num)()
        for i in range(num):
            self.array[i].luma = self.luma[i].ctypes.data
            self.array[i].chroma = self.chroma[i].ctypes.data
        self.num = num

    def __del__(self):
        try:
            if getattr(self, "_pinned", None):
                self.luma = self.chroma = None
                L.lib.b2_host_free(self._pinned)
                self._pinned = None
        except Exception:
            pass

    def sufficient(self, num, luma_len, second_len):
        return num <= self.num and luma_len <= self.luma_len and second_len <= len(self.second)


# Frame buffers belong to the application (reference frames.h); an application that decodes many
# streams reuses them instead of paying page faults for fresh ones on every stream.
_frames_pool = {}
_frames_lock = __import__("threading").Lock()


def _frames_acquire(width, height, num, second_len, pinned=True, cuda_device=-1):
    key = (width, height, num)
    with _frames_lock:
        lst = _frames_pool.get(key)
        if lst:
            return lst.pop()
    return Frames(width, height, num, second_len, pinned, cuda_device)


def _frames_release(fr):
    key = (int(fr.width), int(fr.height), int(fr.num))
    with _frames_lock:
        _frames_pool.setdefault(key, []).append(fr)


def crop_frame(frm):
    """Cropped luma rows followed by cropped interleaved-CbCr rows (filewrite.h:11-29)."""
    stride = frm.width
    height = frm.height - frm.crop[2] - frm.crop[3]
    width = stride - frm.crop[0] - frm.crop[1]
    luma = np.ctypeslib.as_array(C.cast(frm.luma, C.POINTER(C.c_uint8)), shape=(stride * frm.height,))
    chroma = np.ctypeslib.as_array(C.cast(frm.chroma, C.POINTER(C.c_uint8)), shape=(stride * frm.height // 2,))
    y = luma.reshape(-1, stride)[frm.crop[2]:frm.crop[2] + height, frm.crop[0]:frm.crop[0] + width]
    c0 = frm.crop[2] >> 1
    c = chroma.reshape(-1, stride)[c0:c0 + (height >> 1), frm.crop[0]:frm.crop[0] + width]
    return y, c


def frame_md5(frm):
    y, c = crop_frame(frm)
    h = hashlib.md5()
    h.update(np.ascontiguousarray(y).tobytes())
    h.update(np.ascontiguousarray(c).tobytes())
    return h.hexdigest()


class M2Decoder:
    def __init__(self, codec, data, dpb=-1, outbuf=0, tap=None, device_off=False, cuda_device=None, chunk=0, device_md5=False, host_strengths=False):
        """chunk > 0: the stream is handed over `chunk` bytes at a time through ONE reused buffer, as the
        reference application refills its input buffer (bytes of an earlier refill do not stay valid).
        device_md5: digest mode (b2_context_set_md5) -- frames are hashed on the device, `frame_digest(frm)`
        returns what `frame_md5(frm)` would, the frame buffers themselves are not filled.
        host_strengths: the H.264 front end derives the deblocking strengths itself instead of leaving them to the
        device (b2_context_set_host_strengths)."""
        self.codec = codec
        self._chunk, self._pos = int(chunk), 0
        self._chunk_buf = np.empty(max(int(chunk), 1), np.uint8)
        self._device_off, self._cuda_device = device_off, cuda_device
        self.func = L.func_table(_TABLES[codec])
        self._data = np.frombuffer(bytes(data), np.uint8)
        self._fed = False
        self.frames = None
        self.outbuf = outbuf
        self.error = None
        self._ctx = (C.c_uint8 * self.func.context_size)()
        self.ctx = C.addressof(self._ctx)
        self._header_cb = L.HEADER_CB(self._on_header)
        self._refill_cb = L.REFILL_CB(self._on_refill)
        if self.func.init(self.ctx, dpb, self._header_cb, None) < 0:
            raise RuntimeError("init failed")
        self._tap_cb = None
        if tap is not None or device_off:
            self._tap = tap
            self._tap_cb = L.TAP_CB(self._on_tap if tap is not None else 0)
            L.lib.b2_context_set_tap(self.ctx, self._tap_cb, None, 1 if device_off else 0)
        if cuda_device is not None:
            L.lib.b2_context_set_device(self.ctx, cuda_device)
        if device_md5:
            L.lib.b2_context_set_md5(self.ctx, 1)
        if host_strengths:
            L.lib.b2_context_set_host_strengths(self.ctx, 1)
        self.stream = self.func.stream_pos(self.ctx)
        L.lib.dec_bits_set_callback(self.stream, self._refill_cb, None)

    def close(self):
        if self.ctx:
            L.lib.b2_context_release(self.ctx)
            self.ctx = None
            if self.frames is not None:
                _frames_release(self.frames)
                self.frames = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frame_digest(self, frm):
        """MD5 (hex) of an output frame as computed on the device (digest mode)."""
        out = (C.c_uint8 * 16)()
        if L.lib.b2_context_frame_md5(self.ctx, C.byref(frm), out) < 0:
            raise RuntimeError("b2_context_frame_md5: %s" % (L.lib.b2_last_error() or b"").decode())
        return bytes(out).hex()

    # reference h264dec.cpp:67-84: hand the whole file over once, then report end of data
    def _on_refill(self, _arg):
        if self._chunk > 0:
            n = min(self._chunk, len(self._data) - self._pos)
            if n <= 0:
                return -1
            self._chunk_buf[:] = 0xA5                       # whatever was there before is gone
            self._chunk_buf[:n] = self._data[self._pos:self._pos + n]
            self._pos += n
            L.lib.dec_bits_set_data(self.stream, self._chunk_buf.ctypes.data, n, None)
            return 0
        if self._fed:
            return -1
        self._fed = True
        L.lib.dec_bits_set_data(self.stream, self._data.ctypes.data, len(self._data), None)
        return 0

    def _on_tap(self, _arg, codec, pic):
        return int(self._tap(self, codec, pic) or 0)

    # reference m2decoder.h:54-80
    def _on_header(self, _arg, _id):
        info = L.InfoT()
        self.func.get_info(self.ctx, C.byref(info))
        width = (info.src_width + 15) & ~15
        height = (info.src_height + 15) & ~15
        bufnum = self.outbuf + info.frame_num + (16 if self.codec == MODE_H264 else 0)
        bufnum = min(bufnum, H264D_MAX_FRAME_NUM if self.codec == MODE_H264 else MAX_FRAME_NUM)
        if self.frames is not None and self.frames.sufficient(bufnum, width * height, info.additional_size):
            return 0
        self.frames = _frames_acquire(width, height, bufnum, info.additional_size, not self._device_off,
                                      -1 if self._cuda_device is None else self._cuda_device)
        self.info = info
        r = self.func.set_frames(self.ctx, bufnum, self.frames.array,
                                 self.frames.second.ctypes.data, info.additional_size)
        if r < 0:
            self.error = RuntimeError("set_frames failed: %s" % (L.lib.b2_last_error() or b"").decode())
            return -1
        return 0

    # reference m2decoder.h:132-157
    def decode(self, post, emptify=False):
        frm = L.FrameT()
        f = self.func
        while f.peek_decoded_frame(self.ctx, C.byref(frm), 0) <= 0:
            err = f.decode_picture(self.ctx)
            if err < 0:
                while f.peek_decoded_frame(self.ctx, C.byref(frm), 1) > 0:
                    post(frm)
                    f.get_decoded_frame(self.ctx, C.byref(frm), 1)
                return err
        while True:
            f.get_decoded_frame(self.ctx, C.byref(frm), 0)
            post(frm)
            if not (emptify and f.peek_decoded_frame(self.ctx, C.byref(frm), 0) > 0):
                break
        return f.decode_picture(self.ctx)

    def decode_residual(self, post):
        frm = L.FrameT()
        f = self.func
        while f.peek_decoded_frame(self.ctx, C.byref(frm), 1) > 0:
            post(frm)
            f.get_decoded_frame(self.ctx, C.byref(frm), 1)

    # reference h264dec.cpp:251-257
    def run(self, post, emptify=False):
        while True:
            err = self.decode(post, emptify)
            if self.error is not None or err == -4:
                raise self.error or RuntimeError("CUDA reconstruction unavailable")
            if err < 0:
                self.decode_residual(post)
                return err


def decode_file(codec, data, post=None, **kw):
    """Decode a whole elementary stream; returns the list of per-frame MD5 hex strings in
    output order (what `h264dec -O` writes, one per line)."""
    md5s = []

    def _post(frm):
        md5s.append(dec.frame_digest(frm) if kw.get("device_md5") else frame_md5(frm))
        if post is not None:
            post(frm)

    dec = M2Decoder(codec, data, **kw)
    try:
        dec.run(_post)
    finally:
        dec.close()
    return md5s
