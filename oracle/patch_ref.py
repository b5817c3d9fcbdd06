#!/usr/bin/env python3
"""Create the *hooked* variant of the reference for the tests.

Reads the five reference files that carry a B2 hook site from
/root/reference/libde265, inserts the one-line hook calls at anchored positions and writes the
patched copies to oracle/_ref/patched/libde265/ (git-ignored build output — reference sources are
never committed).  All other reference files are compiled from where they lie.

TEST INFRASTRUCTURE ONLY (see oracle/hevc_oracle.h).
"""
import os
import re
import sys


def sub_once(text, pattern, repl, what):
    new, n = re.subn(pattern, repl, text, count=1, flags=re.S)
    if n != 1:
        raise SystemExit(f"patch_ref.py: anchor not found for {what}")
    return new


def write_new(path, text):
    """Write a patched copy.  A stale symlink of the same name (left by an earlier run that did not patch this file yet) must be
    removed first: open(..., "w") would follow it and write into the reference tree."""
    if os.path.islink(path):
        os.remove(path)
    real = os.path.realpath(path)
    if real.startswith("/root/reference"):
        raise SystemExit(f"patch_ref.py: refusing to write into the reference tree ({real})")
    with open(path, "w") as f:
        f.write(text)


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "_ref", "patched")
    src = os.path.join(ref, "libde265")
    dst = os.path.join(out, "libde265")
    os.makedirs(dst, exist_ok=True)
    inc = '#include "libde265_hooks.h"\n'
    # Every other file is a symlink into the reference tree, so that `#include "decctx.h"` from any
    # translation unit resolves to the patched header (quote-includes search the includer's directory).
    patched = {"slice.cc", "motion.cc", "decctx.cc", "decctx.h", "de265.cc", "de265.h"}
    for name in os.listdir(src):
        if name in patched:
            continue
        link = os.path.join(dst, name)
        if os.path.islink(link) or os.path.exists(link):
            os.remove(link) if not os.path.isdir(link) or os.path.islink(link) else None
        if not os.path.exists(link):
            os.symlink(os.path.join(src, name), link)

    # --- slice.cc: decode_TU + PCM ---
    t = open(os.path.join(src, "slice.cc")).read()
    t = sub_once(t, r'(#include "slice.h"\n)', r"\1" + inc, "slice.cc include")
    t = sub_once(
        t,
        r"(static void decode_TU\(thread_context\* tctx,[^{]*\{\n)",
        r"\1  if (b200_hook_decode_TU(tctx, x0, y0, nT, cIdx, (int)cuPredMode, cbf)) return;\n",
        "decode_TU",
    )
    t = sub_once(
        t,
        r"(ptr\[y \* stride \+ x\] = value << shift;\n\s*\}\n)",
        r"\1  b200_hook_pcm(tctx, x0, y0, w, h, cIdx);\n",
        "read_pcm_samples_internal",
    )
    write_new(os.path.join(dst, "slice.cc"), t)

    # --- motion.cc: generate_inter_prediction_samples ---
    t = open(os.path.join(src, "motion.cc")).read()
    t = sub_once(t, r'(#include "motion.h"\n)', r"\1" + inc, "motion.cc include")
    t = sub_once(
        t,
        r"(void generate_inter_prediction_samples\(base_context\* ctx,[^{]*\{\n)",
        r"\1  if (b200_hook_inter_pred(ctx, shdr, img, xC + xB, yC + yB, nPbW, nPbH, vi)) return;\n",
        "generate_inter_prediction_samples",
    )
    write_new(os.path.join(dst, "motion.cc"), t)

    # --- decctx.cc: post-processing filters -> picture done ---
    t = open(os.path.join(src, "decctx.cc")).read()
    t = sub_once(t, r'(#include "decctx.h"\n)', r"\1" + inc, "decctx.cc include")
    t = sub_once(
        t,
        r"(\n\s*)(if \(img->decctx->num_worker_threads\)\s*\n\s*run_postprocessing_filters_parallel\(imgunit\);)",
        r"\1if (b200_hook_picture_done(this, imgunit->img)) { } else\1\2",
        "decode_some post-processing",
    )
    # set_acceleration_functions: the B200 level (DE265_DECODER_PARAM_ACCELERATION_CODE = de265_acceleration_B200)
    t = sub_once(
        t,
        r"(void base_context::set_acceleration_functions\(enum de265_acceleration l\)\n\{\n)",
        r"\1  b200_hook_set_acceleration(this, (int)l);\n",
        "set_acceleration_functions",
    )
    # synthesised reference pictures are mirrored into the backend
    t = sub_once(
        t,
        r"(img->integrity = INTEGRITY_UNAVAILABLE_REFERENCE;\n)(\n\s*return idx;)",
        r"\1  b200_hook_unavailable_reference(this, img);\n\2",
        "generate_unavailable_reference_picture",
    )
    write_new(os.path.join(dst, "decctx.cc"), t)

    # --- de265.cc: the deferred read-back is awaited when the picture is handed to the application ---
    t = open(os.path.join(src, "de265.cc")).read()
    t = sub_once(t, r'(#include "de265.h"\n)', r"\1" + inc, "de265.cc include")
    t = sub_once(
        t,
        r"(de265_image\* img = ctx->get_next_picture_in_output_queue\(\);\n)(\s*return img;)",
        r"\1    b200_hook_wait_image(ctx, img);\n\2",
        "de265_peek_next_picture",
    )
    write_new(os.path.join(dst, "de265.cc"), t)

    # --- de265.h: the acceleration level ---
    t = open(os.path.join(src, "de265.h")).read()
    t = sub_once(t, r"(  de265_acceleration_NEON = 80,\n)", r"\1  de265_acceleration_B200 = 200, // reconstruction on a B200 GPU (libde265_hooks.h)\n", "de265_acceleration enum")
    write_new(os.path.join(dst, "de265.h"), t)

    # --- decctx.h: per-context hook state ---
    t = open(os.path.join(src, "decctx.h")).read()
    t = sub_once(
        t,
        r"(struct acceleration_functions acceleration;[^\n]*\n)",
        r"\1  void* b200_state = nullptr; // B200 reconstruction backend (libde265_hooks.h)\n",
        "base_context member",
    )
    write_new(os.path.join(dst, "decctx.h"), t)
    print("patched:", dst)


if __name__ == "__main__":
    main()
