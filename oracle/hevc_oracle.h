/*
 * hevc_oracle.h — CPU restatement of libde265's scalar reconstruction path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (libde265_b200/, include/)
 * may call, link or import this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * Parity pinning: this restatement is validated (tests/test_oracle_vs_ref.py)
 *   (i)  function by function against the real reference functions compiled from
 *        /root/reference into oracle/_ref/libde265_ref.so (oracle/Makefile), and
 *   (ii) end to end: the reference parser (hooked build, oracle/_ref/libde265_hooked.so)
 *        records girlshy.h265, this file replays the records, and the md5 of the
 *        output equals the reference's golden b81538fa33a67278e5263e231e43ca98
 *        (scripts/ci-run.sh:91-92).
 *
 * Pixels are held as uint16_t for every bit depth (simplicity over speed).
 */
#ifndef HEVC_ORACLE_H
#define HEVC_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "b200hevc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t orc_pixel;

/* ---- B1-level restatements (each cites the reference function it follows) ---- */

/* fallback-dct.cc:1212-1220 + transform.cc:452-525 (scaling list: sclist != NULL) */
void orc_dequant(int16_t* coeff_buf, const int16_t* levels, const uint16_t* pos, int n,
                 int qP, int bit_depth, int log2_nT, const uint8_t* sclist);
/* fallback-dct.cc:550-691 */
void orc_idct_add(orc_pixel* dst, ptrdiff_t stride, int nT, const int16_t* coeffs, int bit_depth);
/* fallback-dct.cc:269-407 */
void orc_dst4_add(orc_pixel* dst, ptrdiff_t stride, const int16_t* coeffs, int bit_depth);
/* fallback-dct.cc:81-91, 161-225 + fallback-dct.h:65-73 ; mode: 0 none, 1 rdpcm_h, 2 rdpcm_v */
void orc_tskip_add(orc_pixel* dst, ptrdiff_t stride, int nT, const int16_t* coeffs, int bit_depth, int rdpcm);
void orc_bypass_add(orc_pixel* dst, ptrdiff_t stride, int nT, const int16_t* coeffs, int bit_depth, int rdpcm);

/* motion.cc:48-174 (incl. edge clamping) + fallback-motion.cc:431-636.  out stride = out_stride. */
void orc_mc_luma(int16_t* out, int out_stride, const orc_pixel* ref, ptrdiff_t ref_stride,
                 int pic_w, int pic_h, int xP, int yP, int mvx, int mvy, int w, int h, int bit_depth);
/* motion.cc:178-282 + fallback-motion.cc:262-415 ; sub_w/sub_h = SubWidthC/SubHeightC */
void orc_mc_chroma(int16_t* out, int out_stride, const orc_pixel* ref, ptrdiff_t ref_stride,
                   int pic_w, int pic_h, int sub_w, int sub_h,
                   int xP, int yP, int mvx, int mvy, int wC, int hC, int bit_depth);
/* fallback-motion.cc:33-256 */
void orc_put_unweighted(orc_pixel* dst, ptrdiff_t stride, const int16_t* src, int sstride, int w, int h, int bd);
void orc_put_avg(orc_pixel* dst, ptrdiff_t stride, const int16_t* s1, const int16_t* s2, int sstride, int w, int h, int bd);
void orc_put_weighted(orc_pixel* dst, ptrdiff_t stride, const int16_t* src, int sstride, int w, int h,
                      int wt, int o, int log2wd, int bd);
void orc_put_weighted_bi(orc_pixel* dst, ptrdiff_t stride, const int16_t* s1, const int16_t* s2, int sstride,
                         int w, int h, int w1, int o1, int w2, int o2, int log2wd, int bd);

/* intrapred.h:529-674 given the availability mask (b200hevc.h); border points at the centre
 * element of an array valid on [-2nT, 2nT]. */
void orc_intra_border(orc_pixel* border, const orc_pixel* plane, ptrdiff_t stride, int xB, int yB, int nT,
                      uint64_t avail, int bit_depth);
/* intrapred.h:185-258 */
void orc_intra_filter(orc_pixel* border, int nT, int cIdx, int mode, int strong, int bit_depth_luma);
/* intrapred.h:261-433 */
void orc_intra_pred(orc_pixel* dst, ptrdiff_t stride, int nT, int cIdx, int mode,
                    const orc_pixel* border, int bit_depth, int disable_boundary_filter);

/* fallback-deblk.h:32-124 */
void orc_deblock_luma_seg(orc_pixel* ptr, ptrdiff_t stride, int vertical, int dE, int dEp, int dEq,
                          int tc, int filterP, int filterQ, int bit_depth);
void orc_deblock_chroma_seg(orc_pixel* ptr, ptrdiff_t stride, int vertical, int tc, int filterP, int filterQ,
                            int bit_depth);

/* ---- B2-level: picture replay ---- */

typedef struct orc_surface {
  int width, height;             /* luma */
  int cw, ch;                    /* chroma */
  int chroma_format_idc;
  int bd_y, bd_c;
  orc_pixel* plane[3];
  ptrdiff_t stride[3];           /* in pixels */
} orc_surface;

typedef struct orc_ctx orc_ctx;

orc_ctx* orc_create(void);
void     orc_destroy(orc_ctx*);
/* Replays one picture into slot params.dst_slot.  Returns 0 or a negative B200_ERR_*. */
int      orc_reconstruct(orc_ctx*, const b200_picture*);
const orc_surface* orc_slot(orc_ctx*, int slot);
int      orc_fill_slot(orc_ctx*, int slot, const b200_pic_params*, int vy, int vc);
/* planes as 8-bit (bd<=8) or little-endian 16-bit samples; strides in bytes */
int      orc_upload_slot(orc_ctx*, int slot, const b200_pic_params*, const void* const planes[3], const size_t strides[3]);
int      orc_read_slot(orc_ctx*, int slot, void* const planes[3], const size_t strides[3]);

/* stage pieces exposed for stage-level tests */
void orc_deblock_picture(orc_surface* s, const b200_picture* pic, int vertical);
void orc_sao_picture(const orc_surface* in, orc_surface* out, const b200_picture* pic);

#ifdef __cplusplus
}
#endif
#endif
