"""Host-side mirror of libde265's public C API (libde265/de265.h:222-456) over ctypes.

The same class drives either an unmodified libde265 build (CPU reconstruction, the reference arm)
or a libde265 built with the B2 hook sites of INTEGRATION.md, in which case ``attach()`` re-points
the per-picture reconstruction at a sink (the B200 engine in production, the CPU oracle in tests).
Names, argument meaning and error behaviour follow de265.h.
"""
import ctypes as C

from . import capi

DE265_OK = 0
DE265_ERROR_IMAGE_BUFFER_FULL = 9
DE265_ERROR_WAITING_FOR_INPUT_DATA = 13

DE265_DECODER_PARAM_BOOL_SEI_CHECK_HASH = 0
DE265_DECODER_PARAM_ACCELERATION_CODE = 5
DE265_DECODER_PARAM_DISABLE_DEBLOCKING = 7
DE265_DECODER_PARAM_DISABLE_SAO = 8
de265_acceleration_SCALAR = 0
de265_acceleration_B200 = 200  # added by the reference-side binding (integration/libde265_hooks.h)
de265_acceleration_AUTO = 10000

SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(capi.Picture), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t))


class Image:
    """A decoded picture handed out by de265_get_next_picture (valid until the next de265_* call)."""

    def __init__(self, lib, ptr):
        self._lib, self._p = lib, ptr

    def width(self, c):
        return self._lib.de265_get_image_width(self._p, c)

    def height(self, c):
        return self._lib.de265_get_image_height(self._p, c)

    def bits_per_pixel(self, c):
        return self._lib.de265_get_bits_per_pixel(self._p, c)

    def chroma_format(self):
        return self._lib.de265_get_chroma_format(self._p)

    def plane_bytes(self, c):
        """Cropped plane as dec265 -o writes it (dec265.cc:121-170): rows of width*bytes, no padding."""
        stride = C.c_int(0)
        ptr = self._lib.de265_get_image_plane(self._p, c, C.byref(stride))
        w, h = self.width(c), self.height(c)
        bpp = (self.bits_per_pixel(c) + 7) // 8
        rows = []
        base = C.cast(ptr, C.c_void_p).value
        for y in range(h):
            rows.append(C.string_at(base + y * stride.value, w * bpp))
        return b"".join(rows)


class Decoder:
    def __init__(self, libpath):
        self.lib = lib = C.CDLL(libpath)  # RTLD_LOCAL: several libde265 builds may coexist in one process
        vp = C.c_void_p
        lib.de265_new_decoder.restype = vp
        lib.de265_free_decoder.argtypes = [vp]
        lib.de265_push_data.argtypes = [vp, C.c_char_p, C.c_int, C.c_int64, vp]
        lib.de265_flush_data.argtypes = [vp]
        lib.de265_decode.argtypes = [vp, C.POINTER(C.c_int)]
        lib.de265_get_next_picture.argtypes = [vp]
        lib.de265_get_next_picture.restype = vp
        lib.de265_release_next_picture.argtypes = [vp]
        lib.de265_release_next_picture.restype = None
        lib.de265_set_parameter_int.argtypes = [vp, C.c_int, C.c_int]
        lib.de265_set_parameter_int.restype = None
        lib.de265_set_parameter_bool.argtypes = [vp, C.c_int, C.c_int]
        lib.de265_set_parameter_bool.restype = None
        lib.de265_get_warning.argtypes = [vp]
        lib.de265_start_worker_threads.argtypes = [vp, C.c_int]
        for f in ("de265_get_image_width", "de265_get_image_height", "de265_get_bits_per_pixel"):
            getattr(lib, f).argtypes = [vp, C.c_int]
        lib.de265_get_chroma_format.argtypes = [vp]
        lib.de265_get_image_plane.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
        lib.de265_get_image_plane.restype = vp
        self.ctx = lib.de265_new_decoder()
        self._sink_ref = None

    # -- B2 boundary ---------------------------------------------------------------------------
    def attach(self, sink):
        """sink(pic: capi.Picture, planes: void*[3], strides: size_t[3]) -> int; None detaches."""
        if not hasattr(self.lib, "de265_b200_attach"):
            raise RuntimeError("this libde265 build has no B2 hook sites (see INTEGRATION.md)")
        self.lib.de265_b200_attach.argtypes = [C.c_void_p, SINK, C.c_void_p]
        self.lib.de265_b200_attach.restype = C.c_int
        if sink is None:
            self.lib.de265_b200_attach(self.ctx, SINK(), None)
            self._sink_ref = None
            return

        def tramp(_user, pic, planes, strides):
            return sink(pic.contents, planes, strides)

        self._sink_ref = SINK(tramp)
        rc = self.lib.de265_b200_attach(self.ctx, self._sink_ref, None)
        if rc:
            raise RuntimeError(f"de265_b200_attach failed: {rc}")

    def select_b200(self):
        """The drop-in switch: DE265_DECODER_PARAM_ACCELERATION_CODE = de265_acceleration_B200 (de265.h:416-427) — the decoder then
        owns a B200 engine, reconstructs every picture on the GPU and hands out pictures exactly as before."""
        if not hasattr(self.lib, "de265_b200_enable"):
            raise RuntimeError("this libde265 build has no B200 backend (see INTEGRATION.md)")
        self.set_parameter_int(DE265_DECODER_PARAM_ACCELERATION_CODE, de265_acceleration_B200)

    # -- de265.h ---------------------------------------------------------------------------------
    def set_parameter_int(self, param, value):
        self.lib.de265_set_parameter_int(self.ctx, param, value)

    def set_parameter_bool(self, param, value):
        self.lib.de265_set_parameter_bool(self.ctx, param, int(bool(value)))

    def push_data(self, data):
        return self.lib.de265_push_data(self.ctx, data, len(data), 0, None)

    def flush_data(self):
        return self.lib.de265_flush_data(self.ctx)

    def decode(self):
        more = C.c_int(0)
        err = self.lib.de265_decode(self.ctx, C.byref(more))
        return err, bool(more.value)

    def get_next_picture(self):
        p = self.lib.de265_get_next_picture(self.ctx)
        return Image(self.lib, p) if p else None

    def get_warning(self):
        return self.lib.de265_get_warning(self.ctx)

    def close(self):
        if self.ctx:
            if self._sink_ref is not None:
                self.attach(None)
            self.lib.de265_free_decoder(self.ctx)
            self.ctx = None

    def decode_stream(self, data, on_picture, chunk=40960, lag=0):
        """The dec265 main loop (dec265.cc:745-881): push 40 KiB chunks, decode, drain pictures.  lag=1: pictures are fetched one
        de265_decode call late (what a player with an output thread does): with the asynchronous B200 backend the host then
        parses picture N+1 while the GPU reconstructs picture N."""
        if lag:
            return self._decode_stream_lagged(data, on_picture, chunk)
        pos, n = 0, 0
        stop = False
        while not stop:
            buf = data[pos:pos + chunk]
            if buf:
                if self.push_data(buf) != DE265_OK:
                    break
            pos += len(buf)
            if pos >= len(data):
                self.flush_data()
                stop = True
            more = True
            while more:
                err, more = self.decode()
                if err != DE265_OK:
                    break
                img = self.get_next_picture()
                if img is not None:
                    on_picture(img)
                    n += 1
                    more = True
                while self.get_warning() != DE265_OK:
                    pass
        return n

    def _decode_stream_lagged(self, data, on_picture, chunk):
        pos, n, owed = 0, 0, 0  # owed: pictures decoded but not fetched yet
        stop = False
        while not stop:
            buf = data[pos:pos + chunk]
            if buf and self.push_data(buf) != DE265_OK:
                break
            pos += len(buf)
            if pos >= len(data):
                self.flush_data()
                stop = True
            more = True
            while more:
                err, more = self.decode()
                if err != DE265_OK:
                    break
                while self.get_warning() != DE265_OK:
                    pass
                if owed:  # the picture of the previous call: its read-back ran while this call parsed
                    img = self.get_next_picture()
                    if img is not None:
                        on_picture(img)
                        n += 1
                        continue
                owed = 1
        while True:  # drain
            img = self.get_next_picture()
            if img is None:
                break
            on_picture(img)
            n += 1
        return n
