"""Host logic of the inter kernel's work list (b2_h264_picture.items, include/b2_recon.h): what the
front end emits must cover every motion-compensation record exactly once per strip, keep the two
parts apart (first-list passes of bi-predicted strips / finishing passes) and be grouped by class."""
import ctypes as C

import pytest

import streams


class MC(C.Structure):          # b2_h264_mc
    _fields_ = [("pos", C.c_uint8), ("size", C.c_uint8), ("ref_frame", C.c_int8 * 2), ("ref_idx", C.c_int8 * 2),
                ("log2wd_y", C.c_uint8), ("log2wd_c", C.c_uint8), ("mv", C.c_int16 * 4), ("w", C.c_int8 * 6), ("o", C.c_int8 * 6),
                ("reserved", C.c_uint8 * 4)]


class MB(C.Structure):          # b2_h264_mb
    _fields_ = [("type", C.c_uint8), ("flags", C.c_uint8), ("qp", C.c_uint8 * 3), ("intra_avail", C.c_uint8), ("i16_mode", C.c_uint8),
                ("chroma_mode", C.c_uint8), ("nz", C.c_uint32), ("coef_idx", C.c_uint32), ("mc_idx", C.c_uint32), ("mc_cnt", C.c_uint8),
                ("intra_avail_c", C.c_uint8), ("reserved", C.c_uint8 * 2), ("i4_modes", C.c_uint8 * 8)]


class Picture(C.Structure):     # b2_h264_picture
    _fields_ = [("mb_w", C.c_int32), ("mb_h", C.c_int32), ("dst", C.c_int32), ("deblock", C.c_int32),
                ("n_mbs", C.c_uint32), ("n_mcs", C.c_uint32), ("n_coef_recs", C.c_uint32), ("n_intra", C.c_uint32),
                ("mbs", C.POINTER(MB)), ("dbk", C.c_void_p), ("mcs", C.POINTER(MC)), ("coefs", C.c_void_p),
                ("items", C.POINTER(C.c_uint32)), ("n_items", C.c_uint32), ("n_items_first", C.c_uint32)]


def frac_class(mvx, mvy):
    fx, fy = mvx & 3, mvy & 3
    return 0 if not fx and not fy else 1 if not fy else 2 if not fx else 4 if fx == 2 else 5 if fy == 2 else 3


@pytest.mark.parametrize("name", ["h264_basic", "h264_umv", "h264_b_bipred", "h264_b_mixed", "h264_cabac_b_mixed", "h264_high_b"])
def test_work_list_covers_every_strip_once(name):
    import m2dec_b200
    assert C.sizeof(MC) == 32 and C.sizeof(MB) == 32
    data = streams.generate("h264", streams.H264_CASES[name])
    seen = {"pictures": 0, "bi": 0}

    def tap(dec, codec, pic):
        p = C.cast(pic, C.POINTER(Picture)).contents
        seen["pictures"] += 1
        want_first, want_last = {}, {}
        for mb in range(p.n_mbs):
            m = p.mbs[mb]
            for ri in range(m.mc_cnt):
                r = p.mcs[m.mc_idx + ri]
                pw, bi = r.size & 15, r.ref_frame[0] >= 0 and r.ref_frame[1] >= 0
                first = 0 if r.ref_frame[0] >= 0 else 1
                last = 1 if bi else first
                c_first, c_last = frac_class(r.mv[2 * first], r.mv[2 * first + 1]), frac_class(r.mv[2 * last], r.mv[2 * last + 1])
                wide = pw >= 2 and c_first < 4 and c_last < 4
                for k in range(pw // 2 if wide else pw):
                    want_last[(mb, ri, k)] = c_last + 6 * wide
                    if bi:
                        want_first[(mb, ri, k)] = c_first + 6 * wide
        seen["bi"] += len(want_first)
        assert p.n_items_first == len(want_first) and p.n_items == len(want_first) + len(want_last)
        for lo, hi, want in ((0, p.n_items_first, want_first), (p.n_items_first, p.n_items, want_last)):
            got, runs, prev = {}, 0, None
            for i in range(lo, hi):
                it = p.items[i]
                key, cls = (it & 0xfffff, (it >> 20) & 15, (it >> 24) & 3), (it >> 26) & 15
                assert key not in got
                got[key] = cls
                runs += cls != prev
                prev = cls
            assert got == want
            assert runs <= 2 * 12          # grouped: each class at most once per edge / non-edge half
        return 0

    dec = m2dec_b200.M2Decoder("h264", data, tap=tap, device_off=True)
    try:
        dec.run(lambda frm: None)
    finally:
        dec.close()
    assert seen["pictures"] > 0
    if "_b_" in name:
        assert seen["bi"] > 0
