#!/usr/bin/env python3
"""Regenerates tests/golden/girlshy_records.bin.gz and girlshy_expected.json.

Runs the reference parser (hooked build, oracle/_ref/libde265_hooked.so, built from /root/reference by
`make -C oracle hooked`) over tests/golden/girlshy.h265, serialises every picture's command records
(b200_picture_serialize) and stores, per picture in decode order, the md5 of the three reconstructed
planes as produced by the CPU oracle.  The run is only accepted if the md5 of the decoder OUTPUT equals
the reference's golden md5 b81538fa33a67278e5263e231e43ca98 (scripts/ci-run.sh:91-92), which pins both
the record format and the oracle to the reference.  Needs /root/reference => dev container only.
"""
import ctypes as C
import gzip
import hashlib
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from libde265_b200 import capi, de265  # noqa: E402
import oracle_lib  # noqa: E402

GOLDEN_MD5 = "b81538fa33a67278e5263e231e43ca98"


def main():
    lib = capi.load()
    orc = oracle_lib.Oracle()
    dec = de265.Decoder(oracle_lib.ref_path("libde265_hooked.so"))
    blobs, md5s = [], []

    def sink(pic, planes, strides):
        n = lib.b200_picture_serialized_size(C.byref(pic))
        buf = C.create_string_buffer(n)
        assert lib.b200_picture_serialize(C.byref(pic), buf, n) == n
        blobs.append(buf.raw)
        orc.reconstruct(pic)
        orc.lib.orc_read_slot(orc.ctx, pic.params.dst_slot, capi.PlaneArray(planes[0], planes[1], planes[2]),
                              capi.StrideArray(strides[0], strides[1], strides[2]))
        md5s.append(hashlib.md5(b"".join(p.tobytes() for p in orc.read_slot(pic.params.dst_slot, pic.params))).hexdigest())
        return 0

    dec.attach(sink)
    data = open(os.path.join(HERE, "girlshy.h265"), "rb").read()
    md = hashlib.md5()
    n = dec.decode_stream(data, lambda img: [md.update(img.plane_bytes(c)) for c in range(3)])
    dec.close()
    assert md.hexdigest() == GOLDEN_MD5, md.hexdigest()
    with gzip.open(os.path.join(HERE, "girlshy_records.bin.gz"), "wb", compresslevel=9) as f:
        for b in blobs:
            f.write(struct.pack("<I", len(b)))
            f.write(b)
    json.dump({"output_md5": GOLDEN_MD5, "frames_output": n, "decode_order_plane_md5": md5s},
              open(os.path.join(HERE, "girlshy_expected.json"), "w"), indent=1)
    print("wrote", len(blobs), "pictures")


if __name__ == "__main__":
    main()
