"""Kernel replay (include/b2_recon.h, b2_replay_*): parse streams once on the host, keep every
picture's SoA resident in HBM, then reconstruct picture #t of all lanes per batched launch."""
import ctypes as C
import threading

import numpy as np

from . import _native as L
from .decoder import M2Decoder, MODE_H264, MODE_MPEG2

CODEC_ID = {MODE_MPEG2: 1, MODE_H264: 2}


class Replay:
    def __init__(self, codec, width, height, lanes, pool_frames, cuda_device=0):
        self.codec = codec
        self.width, self.height, self.lanes, self.pool = width, height, lanes, pool_frames
        h = C.c_void_p()
        if L.lib.b2_replay_create(cuda_device, CODEC_ID[codec], width, height, lanes, pool_frames, C.byref(h)) < 0:
            raise RuntimeError("b2_replay_create failed: %s" % (L.lib.b2_last_error() or b"").decode())
        self.h = h
        self._lock = threading.Lock()      # b2_replay_add_* is not re-entrant; parses may run on many threads
        self.outputs = [[] for _ in range(lanes)]     # per lane: output-order list of frame-pool indices... filled by load_stream

    def close(self):
        if self.h:
            L.lib.b2_replay_destroy(self.h)
            self.h = None

    def load_stream(self, lane, data):
        """Parse `data` on the host (product front end, device off) and upload every picture's
        SoA into lane `lane`.  Returns the number of pictures."""
        n = [0]
        dsts = []
        dst_field = 2 if self.codec == MODE_H264 else 3      # index of `dst` among the int32 header fields
        add = L.lib.b2_replay_add_h264 if self.codec == MODE_H264 else L.lib.b2_replay_add_mpeg2

        def tap(dec, codec, pic):
            with self._lock:
                rc = add(self.h, lane, pic)
            if rc < 0:
                raise RuntimeError("b2_replay_add failed: %s" % (L.lib.b2_last_error() or b"").decode())
            dsts.append(C.cast(pic, C.POINTER(C.c_int32))[dst_field])
            n[0] += 1
            return 0

        dec = M2Decoder(self.codec, data, tap=tap, device_off=True)
        try:
            dec.run(lambda frm: None)
        finally:
            dec.close()
        self.outputs[lane] = dsts          # frame-pool index written by each picture, decode order
        return n[0]

    def run(self, timed=False):
        if L.lib.b2_replay_run(self.h, 1 if timed else 0) < 0:
            raise RuntimeError("b2_replay_run failed: %s" % (L.lib.b2_last_error() or b"").decode())

    def sync(self):
        if L.lib.b2_replay_sync(self.h) < 0:
            raise RuntimeError("b2_replay_sync failed: %s" % (L.lib.b2_last_error() or b"").decode())

    def mark(self, which):
        L.lib.b2_replay_mark(self.h, which)

    def elapsed_ms(self):
        ms = C.c_float()
        if L.lib.b2_replay_elapsed_ms(self.h, C.byref(ms)) < 0:
            raise RuntimeError("b2_replay_elapsed_ms failed")
        return ms.value

    def stats(self):
        st = L.ReplayStats()
        L.lib.b2_replay_get_stats(self.h, C.byref(st))
        return st

    def md5(self, pictures, width=None, height=None, left=0, top=0):
        """MD5 hex digests of `pictures` = [(lane, frame index), ...], computed on the device over the
        cropped window, as `h264dec -O` would print them."""
        n = len(pictures)
        lanes = np.asarray([p[0] for p in pictures], np.int32)
        idxs = np.asarray([p[1] for p in pictures], np.int32)
        out = np.empty(16 * max(n, 1), np.uint8)
        if L.lib.b2_replay_md5(self.h, n, lanes.ctypes.data, idxs.ctypes.data, left, top, width or self.width, height or self.height, out.ctypes.data) < 0:
            raise RuntimeError("b2_replay_md5 failed: %s" % (L.lib.b2_last_error() or b"").decode())
        return [out[16 * i:16 * i + 16].tobytes().hex() for i in range(n)]

    def download(self, lane, idx):
        luma = np.empty(self.width * self.height, np.uint8)
        chroma = np.empty(self.width * self.height // 2, np.uint8)
        if L.lib.b2_replay_download(self.h, lane, idx, luma.ctypes.data, chroma.ctypes.data) < 0:
            raise RuntimeError("b2_replay_download failed")
        return luma.reshape(self.height, self.width), chroma.reshape(self.height // 2, self.width)
