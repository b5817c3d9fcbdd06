"""ctypes access to the CPU oracle (oracle/libhevc_oracle.so) and, when present, the real reference
(oracle/_ref/*.so).  TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py only."""
import ctypes as C
import os
import subprocess

from libde265_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libhevc_oracle.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

_orc = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def oracle():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        lib = C.CDLL(ORACLE_SO)
        vp = C.c_void_p
        lib.orc_create.restype = vp
        lib.orc_destroy.argtypes = [vp]
        lib.orc_destroy.restype = None
        lib.orc_reconstruct.argtypes = [vp, C.POINTER(capi.Picture)]
        lib.orc_fill_slot.argtypes = [vp, C.c_int, C.POINTER(capi.PicParams), C.c_int, C.c_int]
        lib.orc_upload_slot.argtypes = [vp, C.c_int, C.POINTER(capi.PicParams), capi.PlaneArray, capi.StrideArray]
        lib.orc_read_slot.argtypes = [vp, C.c_int, capi.PlaneArray, capi.StrideArray]
        _orc = lib
    return _orc


def ref_path(name):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None
