#!/usr/bin/env python3
"""How tests/golden/intra1080.h265 and intra4k.h265 were made (needs /root/reference; run once, the streams are committed).

Real HEVC bitstreams at BASELINE config 2's size (1920x1080, intra-only, 2 pictures, QP 30, CTB 32) and at the 4K size of the
headline configs (3840x2160, intra-only, 1 picture, QP 32), produced by the
REFERENCE'S OWN ENCODER (enc265 + libde265/encoder, compiled from the sources under /root/reference with plain g++) from a
synthetic YUV sequence (smooth fields at three scales + flat boxes + noise).  enc265 crashes in this reference snapshot when
it allocates its input pictures (de265_image_get_buffer zero-fills through a null SPS, image.cc:164), so the encoder is
built against a SCRATCH copy of image.cc with that one call guarded (`if (img->has_sps())`); nothing of it is kept in the
repo, and only `--sop-structure intra` works (the low-delay mode aborts).  The stream's validity and its golden md5 come from
the UNMODIFIED reference decoder (oracle/_ref/libde265_ref.so): tests/golden/intra1080_expected.json, intra4k_expected.json.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
STREAMS = {  # name: (width, height, pictures, qp, seed, coarsest field scale, boxes, (min, max) box width, (min, max) box height)
    "intra1080": (1920, 1080, 2, 30, 7, 64, 12, (40, 200), (30, 120)),
    "intra4k": (3840, 2160, 1, 32, 11, 128, 30, (60, 400), (40, 240)),
}


def synth_yuv(path, W, H, N, seed, S, boxes, bw, bh):
    rng = np.random.default_rng(seed)

    def field(scale, amp):
        g = rng.standard_normal((H // scale + 3, W // scale + 3)).astype(np.float32)
        f = np.kron(g, np.ones((scale, scale), np.float32))[:H + scale, :W + scale]
        for _ in range(2):
            f = (f + np.roll(f, scale // 2, 0) + np.roll(f, scale // 2, 1) + np.roll(np.roll(f, scale // 2, 0), scale // 2, 1)) / 4
        return amp * f[:H, :W]

    with open(path, "wb") as f:
        for _ in range(N):
            y = 128 + field(S, 60) + field(S // 4, 25) + field(S // 16, 8)
            for _ in range(boxes):
                x0, y0 = rng.integers(0, W - bw[1]), rng.integers(0, H - bh[1])
                w, h = rng.integers(bw[0], bw[1]), rng.integers(bh[0], bh[1])
                y[y0:y0 + h, x0:x0 + w] = rng.integers(30, 220)
            y = np.clip(y + rng.integers(-2, 3, y.shape), 0, 255).astype(np.uint8)
            u = np.clip(128 + field(S // 2, 20)[::2, ::2], 0, 255).astype(np.uint8)
            v = np.clip(128 + field(S // 2, 20)[::2, ::2], 0, 255).astype(np.uint8)
            f.write(y.tobytes() + u.tobytes() + v.tobytes())


def main():
    tmp = tempfile.mkdtemp()
    patched = os.path.join(tmp, "image.cc")
    src = open(os.path.join(REF, "libde265/image.cc")).read()
    open(patched, "w").write(src.replace("  img->fill_image(0,0,0);", "  if (img->has_sps()) img->fill_image(0,0,0);"))
    gen = os.path.join(ROOT, "oracle/_ref/gen")
    srcs = [os.path.join(REF, "libde265", f) for f in os.listdir(os.path.join(REF, "libde265")) if f.endswith(".cc") and f != "image.cc"]
    for d in ("libde265/encoder", "libde265/encoder/algo"):
        srcs += [os.path.join(REF, d, f) for f in os.listdir(os.path.join(REF, d)) if f.endswith(".cc")]
    enc = os.path.join(tmp, "enc265")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-DLIBDE265_EXPORTS", "-DHAVE_POSIX_MEMALIGN", "-DHAVE_MALLOC_H", f"-I{gen}", f"-I{REF}",
                           f"-I{REF}/libde265"] + srcs + [patched, os.path.join(REF, "enc265/enc265.cc"), "-o", enc, "-lpthread"])
    sys.path.insert(0, ROOT)
    from libde265_b200 import de265
    for name, (W, H, N, qp, seed, S, boxes, bw, bh) in STREAMS.items():
        yuv = os.path.join(tmp, name + ".yuv")
        synth_yuv(yuv, W, H, N, seed, S, boxes, bw, bh)
        out = os.path.join(HERE, name + ".h265")
        subprocess.check_call([enc, "--input", yuv, "--width", str(W), "--height", str(H), "--frames", str(N), "--qp", str(qp), "--sop-structure", "intra",
                               "--TB-IntraPredMode", "min-residual", "--output", out], cwd=tmp)  # (the encoder drops a recon.yuv in its cwd)
        dec = de265.Decoder(os.path.join(ROOT, "oracle/_ref/libde265_ref.so"))
        dec.set_parameter_int(de265.DE265_DECODER_PARAM_ACCELERATION_CODE, de265.de265_acceleration_SCALAR)
        md = hashlib.md5()
        n = dec.decode_stream(open(out, "rb").read(), lambda img: [md.update(img.plane_bytes(c)) for c in range(3)])
        dec.close()
        json.dump({"stream": name + ".h265", "width": W, "height": H, "pictures": n, "md5_of_all_planes_in_output_order": md.hexdigest(),
                   "decoder": "unmodified reference, scalar table"}, open(os.path.join(HERE, name + "_expected.json"), "w"), indent=1)
        print(name, n, md.hexdigest())


if __name__ == "__main__":
    main()
