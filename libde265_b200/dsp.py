"""Host-side mirror of boundary B1 (include/b200hevc_dsp.h): the entries of libde265's DSP table
(`struct acceleration_functions`, acceleration.h:29-231) as batched device-executed calls on numpy arrays.
Each method queues one command with the reference entry's argument meaning; run() executes the batch."""
import ctypes as C

import numpy as np

from . import capi


class DspTable:
    def __init__(self, device=0):
        self.lib = capi.load()
        self._h = C.c_void_p()
        capi.check(self.lib.b200_dsp_create(C.byref(self._h), device), "b200_dsp_create")
        self._cmds, self._keep = [], []

    def close(self):
        if self._h:
            self.lib.b200_dsp_destroy(self._h)
            self._h = None

    def _add(self, op, bd, dst, dststride, src=None, src2=None, srcstride=0, w=0, h=0, a=()):
        c = capi.DspCmd()
        c.op, c.bit_depth, c.dst, c.dststride = op, bd, dst, dststride
        c.src, c.src2, c.srcstride, c.w, c.h = src, src2, srcstride, w, h
        for i, v in enumerate(a):
            c.a[i] = int(v)
        self._cmds.append(c)

    @staticmethod
    def _ptr(arr, elem_offset=0):
        return arr.ctypes.data + elem_offset * arr.itemsize

    # put_hevc_qpel_{8,16}[xf][yf] / put_hevc_epel_*: src = 2-D pixel array, (x, y) = the PU's integer position in it
    def mc(self, luma, dst, src, x, y, w, h, fx, fy, bd):
        self._keep += [dst, src]
        self._add(capi.DSP_QPEL if luma else capi.DSP_EPEL, bd, self._ptr(dst), dst.strides[0] // 2, self._ptr(src, y * (src.strides[0] // src.itemsize) + x),
                  None, src.strides[0] // src.itemsize, w, h, (fx, fy))

    def pred(self, op, dst, s1, s2, w, h, bd, params=()):
        self._keep += [dst, s1, s2]
        self._add(op, bd, self._ptr(dst), dst.strides[0] // dst.itemsize, self._ptr(s1), None if s2 is None else self._ptr(s2), s1.strides[0] // 2, w, h, params)

    def transform_add(self, dst, coeffs, log2, bd, dst7=False):
        self._keep += [dst, coeffs]
        self._add(capi.DSP_DST_ADD if dst7 else capi.DSP_TRANSFORM_ADD, bd, self._ptr(dst), dst.strides[0] // dst.itemsize, self._ptr(coeffs), None, 0, 0, 0, (log2,))

    def intra(self, op, dst, border, nT, cidx, bd, mode=0, disable_boundary_filter=0):
        """border: 1-D array of 4nT+1 samples, element 2nT = border[0]."""
        self._keep += [dst, border]
        self._add(op, bd, self._ptr(dst), dst.strides[0] // dst.itemsize, self._ptr(border, 2 * nT), None, 0, 0, 0, (nT, cidx, mode, disable_boundary_filter))

    def deblock(self, luma, buf, x, y, vertical, bd, params):
        """buf: 2-D pixel array, (x, y) = q0 of line 0."""
        self._keep.append(buf)
        stride = buf.strides[0] // buf.itemsize
        self._add(capi.DSP_DEBLOCK_LUMA if luma else capi.DSP_DEBLOCK_CHROMA, bd, self._ptr(buf, y * stride + x), stride, None, None, 0, 0, 0, (vertical,) + tuple(params))

    def run(self):
        n = len(self._cmds)
        arr = (capi.DspCmd * n)(*self._cmds)
        rc = self.lib.b200_dsp_run_batch(self._h, arr, n)
        self._cmds, self._keep = [], []
        capi.check(rc, "b200_dsp_run_batch")
        return n
