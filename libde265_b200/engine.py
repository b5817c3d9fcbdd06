"""Thin Python wrapper over b200_engine_* (include/b200hevc.h part 2).  CUDA only — no CPU fallback."""
import ctypes as C

import numpy as np

from . import capi


class Engine:
    def __init__(self, device=0):
        self.lib = capi.load()
        self._h = C.c_void_p()
        capi.check(self.lib.b200_engine_create(C.byref(self._h), device), "b200_engine_create")

    def close(self):
        if self._h:
            self.lib.b200_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def submit(self, pic):
        """pic: capi.Picture (or an object with a ``.c`` capi.Picture, e.g. synth.SynthPicture)."""
        cp = getattr(pic, "c", pic)
        capi.check(self.lib.b200_engine_submit_picture(self._h, C.byref(cp)), "b200_engine_submit_picture")

    def submit_async(self, pic):
        """Queues the picture (planned by the engine's planner threads, issued in submission order).  The picture's record arrays
        must stay alive until flush() / sync() / read_slot() returns."""
        cp = getattr(pic, "c", pic)
        capi.check(self.lib.b200_engine_submit_picture_async(self._h, C.byref(cp)), "b200_engine_submit_picture_async")

    def flush(self):
        capi.check(self.lib.b200_engine_flush(self._h), "b200_engine_flush")

    def sync(self):
        capi.check(self.lib.b200_engine_sync(self._h), "b200_engine_sync")

    def fill_slot(self, slot, params, vy, vc):
        capi.check(self.lib.b200_engine_fill_slot(self._h, slot, C.byref(params), vy, vc), "b200_engine_fill_slot")

    def upload_slot(self, slot, params, planes):
        """planes: 3 C-contiguous numpy arrays (uint8 or uint16)."""
        pa = capi.PlaneArray(*[p.ctypes.data for p in planes])
        sa = capi.StrideArray(*[p.strides[0] for p in planes])
        capi.check(self.lib.b200_engine_upload_slot(self._h, slot, C.byref(params), pa, sa), "b200_engine_upload_slot")

    def read_slot(self, slot, params):
        """Returns [Y, Cb, Cr] numpy arrays (uint8 for 8-bit, uint16 otherwise)."""
        dt = np.uint16 if params.bit_depth_luma > 8 else np.uint8
        shapes = [(params.height, params.width)]
        if params.chroma_format_idc:
            shapes += [(params.height // 2, params.width // 2)] * 2
        out = [np.empty(s, dt) for s in shapes]
        ptrs = [o.ctypes.data for o in out] + [None] * (3 - len(out))
        strides = [o.strides[0] for o in out] + [0] * (3 - len(out))
        capi.check(self.lib.b200_engine_read_slot(self._h, slot, capi.PlaneArray(*ptrs), capi.StrideArray(*strides)), "b200_engine_read_slot")
        return out

    def read_slot_into(self, slot, planes_ptrs, strides):
        capi.check(self.lib.b200_engine_read_slot(self._h, slot, capi.PlaneArray(*planes_ptrs), capi.StrideArray(*strides)), "b200_engine_read_slot")

    def enable_timing(self, on=True):
        capi.check(self.lib.b200_engine_enable_timing(self._h, int(on)), "b200_engine_enable_timing")

    def last_timing(self):
        ms = (C.c_float * 6)()
        capi.check(self.lib.b200_engine_last_timing(self._h, C.byref(ms)), "b200_engine_last_timing")
        return dict(zip(("h2d", "inter_pred", "recon", "deblock", "sao", "total"), list(ms)))

    def timing_sum(self, reset=True):
        ms = (C.c_float * 6)()
        n = C.c_int(0)
        capi.check(self.lib.b200_engine_timing_sum(self._h, C.byref(ms), C.byref(n), int(reset)), "b200_engine_timing_sum")
        return dict(zip(("h2d", "inter_pred", "recon", "deblock", "sao", "total"), list(ms))), n.value

    def prepare(self, pic):
        cp = getattr(pic, "c", pic)
        h = C.c_void_p()
        capi.check(self.lib.b200_engine_prepare_picture(self._h, C.byref(cp), C.byref(h)), "b200_engine_prepare_picture")
        return h

    def run_prepared(self, h):
        capi.check(self.lib.b200_engine_run_prepared(self._h, h), "b200_engine_run_prepared")

    def free_prepared(self, h):
        self.lib.b200_engine_free_prepared(self._h, h)

    def launch_count(self):
        return int(self.lib.b200_engine_launch_count(self._h))

    def stream(self):
        return self.lib.b200_engine_stream(self._h)

    def set_streams(self, n):
        """Number of CUDA streams pictures are pipelined over (results do not depend on it)."""
        capi.check(self.lib.b200_engine_set_streams(self._h, n), "b200_engine_set_streams")

    def join(self):
        """Stream 0 waits for everything issued so far on the other streams."""
        capi.check(self.lib.b200_engine_join(self._h), "b200_engine_join")
