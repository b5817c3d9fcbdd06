"""CPU tests (-m "not gpu"): the oracle against the reference's golden vectors and real functions,
the recorder/serialiser, and that the C-ABI library loads and exports every declared symbol."""
import ctypes as C
import gzip
import hashlib
import json
import os
import struct

import numpy as np
import pytest

from libde265_b200 import capi, de265, synth
import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_MD5 = "b81538fa33a67278e5263e231e43ca98"  # scripts/ci-run.sh:91-92


def load_records(lib):
    raw = gzip.open(os.path.join(GOLDEN, "girlshy_records.bin.gz"), "rb").read()
    pics, keep, pos = [], [], 0
    while pos < len(raw):
        (n,) = struct.unpack_from("<I", raw, pos)
        pos += 4
        buf = C.create_string_buffer(raw[pos:pos + n], n)
        pos += n
        p = capi.Picture()
        assert lib.b200_picture_deserialize(buf, n, C.byref(p)) == n
        keep.append(buf)
        pics.append(p)
    return pics, keep


# ---- C-ABI boundary ---------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(b200lib):
    hdr = open(os.path.join(ROOT, "include", "b200hevc.h")).read() + open(os.path.join(ROOT, "include", "b200hevc_dsp.h")).read()
    import re
    declared = set(re.findall(r"B200_API\s+[\w\s\*]+?\s+\*?(b200_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    for name in declared:
        assert hasattr(b200lib, name), name
    assert b200lib.b200_abi_version() == 1


def test_engine_fails_loudly_without_gpu(b200lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = b200lib.b200_engine_create(C.byref(h), 0)
    assert rc == -5  # B200_ERR_NO_DEVICE: no CPU fallback
    assert b"no CPU fallback" in b200lib.b200_last_error()


def test_struct_sizes_match_header():
    assert C.sizeof(capi.PicParams) == 20 and C.sizeof(capi.PU) == 24 and C.sizeof(capi.TU) == 24
    assert C.sizeof(capi.WeightEntry) == 28 and C.sizeof(capi.CtbInfo) == 24 and C.sizeof(capi.SliceInfo) == 8


def test_recorder_and_serialisation_roundtrip(b200lib):
    lib = b200lib
    rec = C.c_void_p()
    assert lib.b200_rec_create(C.byref(rec)) == 0
    p = capi.PicParams(width=64, height=64, chroma_format_idc=1, bit_depth_luma=8, bit_depth_chroma=8, log2_ctb_size=6)
    assert lib.b200_rec_begin_picture(rec, C.byref(p)) == 0
    assert lib.b200_rec_add_slice(rec, C.byref(capi.SliceInfo(flags=capi.SLICE_SAO_LUMA))) == 0
    pu = capi.PU(x=0, y=0, w=64, h=64, flags=capi.PU_PRED_L0)
    assert lib.b200_rec_add_pu(rec, C.byref(pu)) == 0
    bad = capi.PU(x=32, y=0, w=64, h=64, flags=1)
    assert lib.b200_rec_add_pu(rec, C.byref(bad)) < 0  # outside the picture
    lv = (C.c_int16 * 3)(5, -3, 1)
    ps = (C.c_int16 * 3)(0, 1, 17)
    tu = capi.TU(x=8, y=8, log2_size=3, cidx=0, flags=capi.TU_CBF, qp=30)
    assert lib.b200_rec_add_tu(rec, C.byref(tu), lv, ps, 3) == 0
    ps_bad = (C.c_int16 * 1)(64)
    assert lib.b200_rec_add_tu(rec, C.byref(tu), lv, ps_bad, 1) < 0  # position outside the 8x8 block
    assert lib.b200_rec_set_ctb(rec, 0, 0, C.byref(capi.CtbInfo(slice_idx=0))) == 0
    lib.b200_rec_bs_map(rec)[3] = 2
    pic = capi.Picture()
    assert lib.b200_rec_end_picture(rec, C.byref(pic)) == 0
    assert (pic.n_pu, pic.n_tu, pic.n_coeff, pic.n_slices) == (1, 1, 3, 1)
    n = lib.b200_picture_serialized_size(C.byref(pic))
    buf = C.create_string_buffer(n)
    assert lib.b200_picture_serialize(C.byref(pic), buf, n) == n
    assert lib.b200_picture_serialize(C.byref(pic), buf, n - 1) == 0  # too small
    back = capi.Picture()
    assert lib.b200_picture_deserialize(buf, n, C.byref(back)) == n
    assert back.n_coeff == 3 and back.coeffs[2].pos == 17 and back.coeffs[1].level == -3 and back.bs_map[3] == 2
    assert lib.b200_picture_deserialize(buf, 10, C.byref(back)) == 0  # truncated / ragged input
    lib.b200_rec_destroy(rec)


# ---- the oracle against the reference's golden vector ------------------------------------------------
def test_oracle_replays_golden_records(b200lib, oracle_mod):
    """Committed fixture (made by tests/golden/make_girlshy_records.py from the reference parser): the oracle
    must reproduce every picture's md5; the fixture itself was only accepted with the reference's output md5."""
    exp = json.load(open(os.path.join(GOLDEN, "girlshy_expected.json")))
    assert exp["output_md5"] == GOLDEN_MD5
    pics, keep = load_records(b200lib)
    assert len(pics) == 75
    orc = oracle_mod.Oracle()
    for i, pic in enumerate(pics):
        orc.reconstruct(pic)
        md5 = hashlib.md5(b"".join(p.tobytes() for p in orc.read_slot(pic.params.dst_slot, pic.params))).hexdigest()
        assert md5 == exp["decode_order_plane_md5"][i], f"picture {i}"
    orc.close()


@pytest.mark.skipif(oracle_lib.ref_path("libde265_hooked.so") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_reference_parser_plus_oracle_reproduces_golden_md5(b200lib, oracle_mod):
    """End to end through the B2 boundary on the CPU: reference parser (hooked) -> records -> oracle replay."""
    orc = oracle_mod.Oracle()
    dec = de265.Decoder(oracle_lib.ref_path("libde265_hooked.so"))

    def sink(pic, planes, strides):
        orc.reconstruct(pic)
        orc.lib.orc_read_slot(orc.ctx, pic.params.dst_slot, capi.PlaneArray(planes[0], planes[1], planes[2]),
                              capi.StrideArray(strides[0], strides[1], strides[2]))
        return 0

    dec.attach(sink)
    md = hashlib.md5()
    n = dec.decode_stream(open(os.path.join(GOLDEN, "girlshy.h265"), "rb").read(), lambda img: [md.update(img.plane_bytes(c)) for c in range(3)])
    dec.close()
    assert n == 75 and md.hexdigest() == GOLDEN_MD5


@pytest.mark.skipif(oracle_lib.ref_path("libde265_ref.so") is None, reason="oracle/_ref not built")
def test_unmodified_reference_reproduces_golden_md5():
    dec = de265.Decoder(oracle_lib.ref_path("libde265_ref.so"))
    dec.set_parameter_int(de265.DE265_DECODER_PARAM_ACCELERATION_CODE, de265.de265_acceleration_SCALAR)
    md = hashlib.md5()
    n = dec.decode_stream(open(os.path.join(GOLDEN, "girlshy.h265"), "rb").read(), lambda img: [md.update(img.plane_bytes(c)) for c in range(3)])
    dec.close()
    assert n == 75 and md.hexdigest() == GOLDEN_MD5


# ---- second golden vector: a real 1080p intra-only stream made with the reference's own encoder (config 2's size) ----
REAL_STREAMS = {name: json.load(open(os.path.join(GOLDEN, name + "_expected.json"))) for name in ("intra1080", "intra4k")}


@pytest.mark.parametrize("name", sorted(REAL_STREAMS))
@pytest.mark.skipif(oracle_lib.ref_path("libde265_ref.so") is None, reason="oracle/_ref not built")
def test_unmodified_reference_reproduces_real_intra_stream_md5(name):
    exp = REAL_STREAMS[name]
    dec = de265.Decoder(oracle_lib.ref_path("libde265_ref.so"))
    md = hashlib.md5()
    n = dec.decode_stream(open(os.path.join(GOLDEN, name + ".h265"), "rb").read(), lambda img: [md.update(img.plane_bytes(c)) for c in range(3)])
    dec.close()
    assert n == exp["pictures"] and md.hexdigest() == exp["md5_of_all_planes_in_output_order"]


@pytest.mark.parametrize("name", sorted(REAL_STREAMS))
@pytest.mark.skipif(oracle_lib.ref_path("libde265_hooked.so") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_reference_parser_plus_oracle_reproduces_real_intra_stream_md5(b200lib, oracle_mod, name):
    """1920x1080 and 3840x2160 (33.75 / 67.5 CTB rows of 32: partial CTBs), every intra mode / partition / TU split a real encoder chose, DST."""
    exp = REAL_STREAMS[name]
    orc = oracle_mod.Oracle()
    dec = de265.Decoder(oracle_lib.ref_path("libde265_hooked.so"))
    stats = {"tus": 0, "pics": 0}

    def sink(pic, planes, strides):
        stats["tus"] += pic.n_tu
        stats["pics"] += 1
        orc.reconstruct(pic)
        orc.lib.orc_read_slot(orc.ctx, pic.params.dst_slot, capi.PlaneArray(planes[0], planes[1], planes[2]),
                              capi.StrideArray(strides[0], strides[1], strides[2]))
        return 0

    dec.attach(sink)
    md = hashlib.md5()
    n = dec.decode_stream(open(os.path.join(GOLDEN, name + ".h265"), "rb").read(), lambda img: [md.update(img.plane_bytes(c)) for c in range(3)])
    dec.close()
    orc.close()
    assert n == exp["pictures"] and stats["pics"] == n and stats["tus"] > 50000
    assert md.hexdigest() == exp["md5_of_all_planes_in_output_order"]


# ---- synthetic generator sanity (host logic) ---------------------------------------------------------
def test_synth_is_deterministic_and_legal():
    a = synth.make_picture(128, 72, "B", seed=7, dst_slot=1, ref_slots=(0,), weighted=True, n_slices=2)
    b = synth.make_picture(128, 72, "B", seed=7, dst_slot=1, ref_slots=(0,), weighted=True, n_slices=2)
    assert a.tus.tobytes() == b.tus.tobytes() and a.pus.tobytes() == b.pus.tobytes() and a.bs_map.tobytes() == b.bs_map.tobytes()
    pu = a.pus
    assert ((pu["x"].astype(int) + pu["w"]) <= 128).all() and ((pu["y"].astype(int) + pu["h"]) <= 72).all()
    # PUs and intra/PCM luma blocks tile the picture exactly once
    cover = np.zeros((72, 128), np.int32)
    for p in pu:
        cover[p["y"]:p["y"] + p["h"], p["x"]:p["x"] + p["w"]] += 1
    assert cover.max() <= 1
    assert a.algorithmic_mc_bytes() > 0


def test_oracle_synthetic_idempotent(oracle_mod):
    """Replaying the same records twice gives the same picture (no hidden state in the oracle)."""
    orc = oracle_mod.Oracle()
    p0 = synth.make_picture(128, 64, "I", seed=3, dst_slot=0)
    p1 = synth.make_picture(128, 64, "B", seed=4, dst_slot=1, ref_slots=(0,))
    orc.reconstruct(p0)
    orc.reconstruct(p1)
    first = [x.copy() for x in orc.read_slot(1, p1.params)]
    orc.reconstruct(p1)
    again = orc.read_slot(1, p1.params)
    assert all((x == y).all() for x, y in zip(first, again))


@pytest.mark.skipif(oracle_lib.ref_path("libde265_hooked.so") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_hook_availability_masks_match_the_reference_border_computer():
    """integration/libde265_hooks.cc derives b200_tu.avail without touching samples; with B200_HOOK_CHECK=1 every TU's mask is
    compared against the reference's own intra_border_computer (preproc + fill_from_image) and a mismatch aborts.  All three
    golden streams (inter pictures with intra blocks, 1080p and 4K intra pictures)."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import oracle_lib\n"
        "from libde265_b200 import de265\n"
        "for name, n_exp in (('girlshy', 75), ('intra1080', 2), ('intra4k', 1)):\n"
        "    dec = de265.Decoder(oracle_lib.ref_path('libde265_hooked.so'))\n"
        "    dec.attach(lambda pic, planes, strides: 0)\n"
        "    n = dec.decode_stream(open(%r + '/' + name + '.h265', 'rb').read(), lambda img: None)\n"
        "    dec.close()\n"
        "    assert n == n_exp, (name, n)\n"
        "print('masks ok')\n") % (ROOT, os.path.join(ROOT, "tests"), GOLDEN)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B200_HOOK_CHECK="1"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "masks ok" in r.stdout, r.stderr[-800:]
