/*
 * b200hevc_dsp.h — boundary B1: the per-block DSP table of libde265 (`struct acceleration_functions`,
 * libde265/acceleration.h:29-231) as batched device-executed entry points.
 *
 * One b200_dsp_cmd = one call of one table entry, with the SAME argument meaning as the reference's
 * function pointer (host pointers, strides in ELEMENTS, caller owns every buffer, the callee may read the
 * same halo the reference functions read).  b200_dsp_run_batch stages the blocks of n commands into
 * device memory, executes them in ONE kernel launch (one CTA per command) and copies the results back;
 * it returns when the host buffers hold the results.  This is the parity / ABI boundary: the engine's
 * throughput path is the per-picture boundary B2 (b200hevc.h), because the table's callers run per 4x4..64x64
 * block on the parsing thread and read reconstructed host pixels.
 * `integration/accel_b200.cc` is the reference-side stub that fills a `struct acceleration_functions`
 * (init_acceleration_functions_b200) with wrappers issuing batches of one.
 *
 * Entry (acceleration.h line)                      op                         a[] / fields used
 *   put_hevc_qpel_{8,16}[xFrac][yFrac]     :100-120  B200_DSP_QPEL              a0=xFrac a1=yFrac; src = uint8/uint16 at the PU's
 *                                                                                integer position (halo -3..+4 read when the phase is
 *                                                                                fractional), dst = int16 w x h
 *   put_hevc_epel{,_h,_v,_hv}_{8,16}       :87-99    B200_DSP_EPEL              a0=mx a1=my (eighth-sample phases), halo -1..+2
 *   put_unweighted_pred_{8,16}             :35,53    B200_DSP_PRED_UNI          src = int16, dst = pixels
 *   put_weighted_pred_avg_{8,16}           :31,49    B200_DSP_PRED_AVG          src, src2
 *   put_weighted_pred_{8,16}               :39,57    B200_DSP_PRED_WEIGHTED     a0=w a1=o a2=log2WD
 *   put_weighted_bipred_{8,16}             :43,61    B200_DSP_PRED_WEIGHTED_BI  a0=w1 a1=o1 a2=w2 a3=o2 a4=log2WD
 *   transform_add_{8,16}[log2-2]           :153,158  B200_DSP_TRANSFORM_ADD     a0=log2(nT); src = int16 coeffs[nT*nT], dst += residual
 *   transform_4x4_dst_add_{8,16}           :152,157  B200_DSP_DST_ADD
 *   intra_pred_dc_{8,16}                   :205-206  B200_DSP_INTRA_DC          a0=nT a1=cIdx; src = border (centre element, valid [-2nT,2nT])
 *   intra_pred_planar_{8,16}               :207-208  B200_DSP_INTRA_PLANAR      a0=nT a1=cIdx
 *   intra_pred_angular_{8,16}              :209-212  B200_DSP_INTRA_ANGULAR     a0=nT a1=cIdx a2=mode a3=disableBoundaryFilter
 *   deblock_luma_8 (+ >8-bit kernel)       :184      B200_DSP_DEBLOCK_LUMA      a0=vertical a1=dE a2=dEp a3=dEq a4=tc a5=filterP a6=filterQ;
 *                                                                                dst = q0 of line 0
 *   deblock_chroma_8 (+ >8-bit kernel)     :186      B200_DSP_DEBLOCK_CHROMA    a0=vertical a1=tc a2=filterP a3=filterQ
 * Entries not listed (dequant, transform-skip/bypass/rdpcm, rotate, encoder-only) keep the scalar functions
 * init_acceleration_functions_fallback installed, exactly as the SSE/ARM initialisers do (decctx.cc:243-269).
 */
#ifndef B200HEVC_DSP_H
#define B200HEVC_DSP_H

#include "b200hevc.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
  B200_DSP_QPEL = 1,
  B200_DSP_EPEL = 2,
  B200_DSP_PRED_UNI = 3,
  B200_DSP_PRED_AVG = 4,
  B200_DSP_PRED_WEIGHTED = 5,
  B200_DSP_PRED_WEIGHTED_BI = 6,
  B200_DSP_TRANSFORM_ADD = 7,
  B200_DSP_DST_ADD = 8,
  B200_DSP_INTRA_DC = 9,
  B200_DSP_INTRA_PLANAR = 10,
  B200_DSP_INTRA_ANGULAR = 11,
  B200_DSP_DEBLOCK_LUMA = 12,
  B200_DSP_DEBLOCK_CHROMA = 13
};

typedef struct b200_dsp_cmd {
  int32_t op;          /* B200_DSP_* */
  int32_t bit_depth;   /* 8..12; selects the _8 / _16 flavour (pixel buffers are uint8_t at 8, uint16_t above) */
  void* dst;
  ptrdiff_t dststride; /* elements */
  const void* src;
  const void* src2;
  ptrdiff_t srcstride; /* elements */
  int32_t w, h;        /* block size where the entry takes one (MC, weighting) */
  int32_t a[8];        /* see the table above */
} b200_dsp_cmd;

typedef struct b200_dsp b200_dsp;

B200_API int  b200_dsp_create(b200_dsp** out, int device);
B200_API void b200_dsp_destroy(b200_dsp*);
/* Executes cmds[0..n) (independent blocks; overlapping destinations are the caller's bug, as with the table). */
B200_API int  b200_dsp_run_batch(b200_dsp*, const b200_dsp_cmd* cmds, int n);

#ifdef __cplusplus
}
#endif
#endif
