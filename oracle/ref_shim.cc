// ref_shim.cc — C entry points onto the UNMODIFIED reference's DSP table for ctypes.
// TEST INFRASTRUCTURE ONLY.  Linked against oracle/_ref/libde265_ref.so (built from /root/reference
// by oracle/Makefile).  `simd`=0 selects the scalar fallback table (fallback.cc:28-140, the parity
// oracle), `simd`=1 the SSE4.1/AVX2/AVX-512 table (x86/sse.cc:46-172, the CPU baseline).
#include <stdint.h>
#include <string.h>

#include "libde265/acceleration.h"
#include "libde265/fallback.h"
#include "libde265/fallback-deblk.h"
#include "libde265/intrapred.h"
#include "libde265/sps.h"
#ifdef HAVE_SSE4_1
#include "libde265/x86/sse.h"
#endif

static acceleration_functions g_tab[2];
static bool g_init = false;

static const acceleration_functions& tab(int simd)
{
  if (!g_init) {
    init_acceleration_functions_fallback(&g_tab[0]);
    init_acceleration_functions_fallback(&g_tab[1]);
#ifdef HAVE_SSE4_1
    init_acceleration_functions_sse(&g_tab[1]);
#endif
    g_init = true;
  }
  return g_tab[simd ? 1 : 0];
}

#define EXPORT extern "C" __attribute__((visibility("default")))

EXPORT int ref_has_simd(void)
{
#ifdef HAVE_SSE4_1
  return 1;
#else
  return 0;
#endif
}

// ---- MC ----
EXPORT void ref_put_qpel_8(int simd, int xf, int yf, int16_t* dst, ptrdiff_t dststride, const uint8_t* src, ptrdiff_t srcstride, int w, int h)
{
  ALIGNED_16(int16_t) mcbuffer[64 * (64 + 7)];
  tab(simd).put_hevc_qpel_8[xf][yf](dst, dststride, src, srcstride, w, h, mcbuffer);
}
EXPORT void ref_put_qpel_16(int simd, int xf, int yf, int16_t* dst, ptrdiff_t dststride, const uint16_t* src, ptrdiff_t srcstride, int w, int h,
                            int bit_depth)
{
  ALIGNED_16(int16_t) mcbuffer[64 * (64 + 7)];
  tab(simd).put_hevc_qpel_16[xf][yf](dst, dststride, src, srcstride, w, h, mcbuffer, bit_depth);
}
EXPORT void ref_put_epel_8(int simd, int mx, int my, int16_t* dst, ptrdiff_t dststride, const uint8_t* src, ptrdiff_t srcstride, int w, int h)
{
  ALIGNED_16(int16_t) mcbuffer[64 * (64 + 7)];
  const acceleration_functions& t = tab(simd);
  if (mx && my) t.put_hevc_epel_hv_8(dst, dststride, src, srcstride, w, h, mx, my, mcbuffer, 8);
  else if (mx) t.put_hevc_epel_h_8(dst, dststride, src, srcstride, w, h, mx, my, mcbuffer, 8);
  else if (my) t.put_hevc_epel_v_8(dst, dststride, src, srcstride, w, h, mx, my, mcbuffer, 8);
  else t.put_hevc_epel_8(dst, dststride, src, srcstride, w, h, mx, my, mcbuffer);
}
EXPORT void ref_put_epel_16(int simd, int mx, int my, int16_t* dst, ptrdiff_t dststride, const uint16_t* src, ptrdiff_t srcstride, int w, int h,
                            int bit_depth)
{
  ALIGNED_16(int16_t) mcbuffer[64 * (64 + 7)];
  const acceleration_functions& t = tab(simd);
  if (mx && my) t.put_hevc_epel_hv_16(dst, dststride, src, srcstride, w, h, mx, my, mcbuffer, bit_depth);
  else if (mx) t.put_hevc_epel_h_16(dst, dststride, src, srcstride, w, h, mx, my, mcbuffer, bit_depth);
  else if (my) t.put_hevc_epel_v_16(dst, dststride, src, srcstride, w, h, mx, my, mcbuffer, bit_depth);
  else t.put_hevc_epel_16(dst, dststride, src, srcstride, w, h, mx, my, mcbuffer, bit_depth);
}

// ---- weighting ----
EXPORT void ref_put_unweighted_8(int simd, uint8_t* dst, ptrdiff_t ds, const int16_t* src, ptrdiff_t ss, int w, int h)
{ tab(simd).put_unweighted_pred_8(dst, ds, src, ss, w, h); }
EXPORT void ref_put_avg_8(int simd, uint8_t* dst, ptrdiff_t ds, const int16_t* s1, const int16_t* s2, ptrdiff_t ss, int w, int h)
{ tab(simd).put_weighted_pred_avg_8(dst, ds, s1, s2, ss, w, h); }
EXPORT void ref_put_weighted_8(int simd, uint8_t* dst, ptrdiff_t ds, const int16_t* src, ptrdiff_t ss, int w, int h, int wt, int o, int log2wd)
{ tab(simd).put_weighted_pred_8(dst, ds, src, ss, w, h, wt, o, log2wd); }
EXPORT void ref_put_bipred_8(int simd, uint8_t* dst, ptrdiff_t ds, const int16_t* s1, const int16_t* s2, ptrdiff_t ss, int w, int h, int w1, int o1,
                             int w2, int o2, int log2wd)
{ tab(simd).put_weighted_bipred_8(dst, ds, s1, s2, ss, w, h, w1, o1, w2, o2, log2wd); }
EXPORT void ref_put_unweighted_16(int simd, uint16_t* dst, ptrdiff_t ds, const int16_t* src, ptrdiff_t ss, int w, int h, int bd)
{ tab(simd).put_unweighted_pred_16(dst, ds, src, ss, w, h, bd); }
EXPORT void ref_put_avg_16(int simd, uint16_t* dst, ptrdiff_t ds, const int16_t* s1, const int16_t* s2, ptrdiff_t ss, int w, int h, int bd)
{ tab(simd).put_weighted_pred_avg_16(dst, ds, s1, s2, ss, w, h, bd); }
EXPORT void ref_put_weighted_16(int simd, uint16_t* dst, ptrdiff_t ds, const int16_t* src, ptrdiff_t ss, int w, int h, int wt, int o, int log2wd, int bd)
{ tab(simd).put_weighted_pred_16(dst, ds, src, ss, w, h, wt, o, log2wd, bd); }
EXPORT void ref_put_bipred_16(int simd, uint16_t* dst, ptrdiff_t ds, const int16_t* s1, const int16_t* s2, ptrdiff_t ss, int w, int h, int w1, int o1,
                              int w2, int o2, int log2wd, int bd)
{ tab(simd).put_weighted_bipred_16(dst, ds, s1, s2, ss, w, h, w1, o1, w2, o2, log2wd, bd); }

// ---- residual ----
EXPORT void ref_transform_add_8(int simd, int log2, uint8_t* dst, const int16_t* coeffs, ptrdiff_t stride)
{ tab(simd).transform_add_8[log2 - 2](dst, coeffs, stride); }
EXPORT void ref_transform_add_16(int simd, int log2, uint16_t* dst, const int16_t* coeffs, ptrdiff_t stride, int bd)
{ tab(simd).transform_add_16[log2 - 2](dst, coeffs, stride, bd); }
EXPORT void ref_dst_add_8(int simd, uint8_t* dst, const int16_t* coeffs, ptrdiff_t stride) { tab(simd).transform_4x4_dst_add_8(dst, coeffs, stride); }
EXPORT void ref_dst_add_16(int simd, uint16_t* dst, const int16_t* coeffs, ptrdiff_t stride, int bd)
{ tab(simd).transform_4x4_dst_add_16(dst, coeffs, stride, bd); }
EXPORT void ref_dequant(int simd, int16_t* buf, const int16_t* list, const int16_t* pos, int n, int fact, int offset, int bdshift)
{ tab(simd).dequant_coeff_block(buf, list, pos, n, fact, offset, bdshift); }
EXPORT void ref_tskip_residual(int simd, int32_t* r, const int16_t* c, int nT, int tsShift, int bdShift) { tab(simd).transform_skip_residual(r, c, nT, tsShift, bdShift); }
EXPORT void ref_rdpcm(int simd, int vertical, int32_t* r, const int16_t* c, int nT, int tsShift, int bdShift)
{ if (vertical) tab(simd).rdpcm_v(r, c, nT, tsShift, bdShift); else tab(simd).rdpcm_h(r, c, nT, tsShift, bdShift); }
EXPORT void ref_bypass(int simd, int mode, int32_t* r, const int16_t* c, int nT)
{
  if (mode == 0) tab(simd).transform_bypass(r, c, nT);
  else if (mode == 1) tab(simd).transform_bypass_rdpcm_h(r, c, nT);
  else tab(simd).transform_bypass_rdpcm_v(r, c, nT);
}
EXPORT void ref_add_residual_8(int simd, uint8_t* dst, ptrdiff_t stride, const int32_t* r, int nT) { tab(simd).add_residual_8(dst, stride, r, nT, 8); }
EXPORT void ref_add_residual_16(int simd, uint16_t* dst, ptrdiff_t stride, const int32_t* r, int nT, int bd) { tab(simd).add_residual_16(dst, stride, r, nT, bd); }
EXPORT void ref_rotate(int simd, int16_t* c, int nT) { tab(simd).rotate_coefficients(c, nT); }

// ---- intra ----  (border points at the centre element, valid on [-2nT, 2nT])
EXPORT void ref_intra_8(int simd, uint8_t* dst, int stride, int nT, int cIdx, int mode, uint8_t* border, int disable_filter)
{
  const acceleration_functions& t = tab(simd);
  if (mode == 0) t.intra_pred_planar_8(dst, stride, nT, cIdx, border);
  else if (mode == 1) t.intra_pred_dc_8(dst, stride, nT, cIdx, border);
  else t.intra_pred_angular_8(dst, stride, 8, disable_filter, 0, 0, (enum IntraPredMode)mode, nT, cIdx, border);
}
EXPORT void ref_intra_16(int simd, uint16_t* dst, int stride, int nT, int cIdx, int mode, uint16_t* border, int disable_filter, int bd)
{
  const acceleration_functions& t = tab(simd);
  if (mode == 0) t.intra_pred_planar_16(dst, stride, nT, cIdx, border);
  else if (mode == 1) t.intra_pred_dc_16(dst, stride, nT, cIdx, border);
  else t.intra_pred_angular_16(dst, stride, bd, disable_filter, 0, 0, (enum IntraPredMode)mode, nT, cIdx, border);
}
EXPORT void ref_intra_filter_8(uint8_t* border, int nT, int cIdx, int mode, int strong)
{
  seq_parameter_set sps;
  sps.strong_intra_smoothing_enable_flag = strong;
  sps.bit_depth_luma = 8;
  intra_prediction_sample_filtering<uint8_t>(sps, border, nT, cIdx, (enum IntraPredMode)mode);
}
EXPORT void ref_intra_filter_16(uint16_t* border, int nT, int cIdx, int mode, int strong, int bd)
{
  seq_parameter_set sps;
  sps.strong_intra_smoothing_enable_flag = strong;
  sps.bit_depth_luma = bd;
  intra_prediction_sample_filtering<uint16_t>(sps, border, nT, cIdx, (enum IntraPredMode)mode);
}

// ---- deblocking ----
EXPORT void ref_deblock_luma_8(int simd, uint8_t* ptr, ptrdiff_t stride, int vertical, int dE, int dEp, int dEq, int tc, int fP, int fQ)
{ tab(simd).deblock_luma_8(ptr, stride, vertical, dE, dEp, dEq, tc, fP, fQ); }
EXPORT void ref_deblock_chroma_8(int simd, uint8_t* ptr, ptrdiff_t stride, int vertical, int tc, int fP, int fQ)
{ tab(simd).deblock_chroma_8(ptr, stride, vertical, tc, fP, fQ); }
EXPORT void ref_deblock_luma_16(uint16_t* ptr, ptrdiff_t stride, int vertical, int dE, int dEp, int dEq, int tc, int fP, int fQ, int bd)
{ deblock_luma_kernel<uint16_t>(ptr, stride, vertical, dE, dEp, dEq, tc, fP, fQ, bd); }
EXPORT void ref_deblock_chroma_16(uint16_t* ptr, ptrdiff_t stride, int vertical, int tc, int fP, int fQ, int bd)
{ deblock_chroma_kernel<uint16_t>(ptr, stride, vertical, tc, fP, fQ, bd); }

// ---- picture-level post-filter DRIVERS of the reference on a synthetic de265_image --------------------------------
// Pins the oracle's deblocking driver (orc_deblock_picture) and SAO driver (orc_sao_picture) — the part no table entry
// reaches — to the reference's own edge_filtering_luma / edge_filtering_chroma (deblock.cc:608-617,764-774; V pass over
// the whole picture, then H, as apply_deblocking_filter deblock.cc:908-946 orders them) and
// apply_sample_adaptive_offset_sequential (sao.cc:327-382).  The image metadata is filled from the b200 records: bS from
// bs_map (what derive_boundaryStrength would have stored, image.h:832-842), QP_Y and the no-filter bit per 8x8
// (cu_transquant_bypass: deblock.cc:576-592, sao.cc:112-117), slice headers, SAO parameters, tiles (uniform spacing).
#include "b200hevc.h"
#include "libde265/deblock.h"
#include "libde265/image.h"
#include "libde265/pps.h"
#include "libde265/sao.h"
#include "libde265/slice.h"

#include <memory>
#include <cstdio>
#include <cstdlib>
#include <vector>

// non-static in deblock.cc but not declared in deblock.h
void edge_filtering_luma(de265_image* img, bool vertical, int yStart, int yEnd, int xStart, int xEnd);
void edge_filtering_chroma(de265_image* img, bool vertical, int yStart, int yEnd, int xStart, int xEnd);

#include "libde265/decctx.h"

// A decoder context only as the holder of the DSP table the drivers dispatch through (deblock.cc:595,736,752):
// simd=0 the scalar fallback table, simd=1 whatever de265_acceleration_AUTO selects on this CPU (SSE4.1 + AVX2 + AVX-512).
static decoder_context* filter_ctx(int simd)
{
  static decoder_context* ctx[2] = {nullptr, nullptr};
  decoder_context*& c = ctx[simd ? 1 : 0];
  if (!c) {
    c = new decoder_context();
    c->set_acceleration_functions(simd ? de265_acceleration_AUTO : de265_acceleration_SCALAR);
  }
  return c;
}

// planes: in/out, strides in BYTES.  stages: bit 0 deblock, bit 1 SAO, bit 2 use the SIMD table.
EXPORT int ref_postfilter(const b200_picture* pic, void* const planes[3], const size_t strides[3], int stages, int tile_cols, int tile_rows)
{
  const b200_pic_params& p = pic->params;
  auto sps = std::make_shared<seq_parameter_set>();
  sps->set_defaults();
  sps->pic_width_in_luma_samples = p.width;
  sps->pic_height_in_luma_samples = p.height;
  sps->chroma_format_idc = p.chroma_format_idc;
  sps->bit_depth_luma = p.bit_depth_luma;
  sps->bit_depth_chroma = p.bit_depth_chroma;
  sps->log2_min_luma_coding_block_size = 3;
  sps->log2_diff_max_min_luma_coding_block_size = p.log2_ctb_size - 3;
  sps->log2_min_transform_block_size = 2;
  sps->log2_diff_max_min_transform_block_size = (p.log2_ctb_size < 5 ? p.log2_ctb_size : 5) - 2;
  sps->max_transform_hierarchy_depth_inter = 1;
  sps->max_transform_hierarchy_depth_intra = 1;
  sps->sample_adaptive_offset_enabled_flag = (p.flags & B200_PIC_SAO_ENABLED) != 0;
  sps->pcm_loop_filter_disable_flag = 0;  // the records fold (pcm && pcm_loop_filter_disable) into the no-filter bit
  if (sps->compute_derived_values(true) != DE265_OK) return -1;
  auto pps = std::make_shared<pic_parameter_set>();
  pps->set_defaults();
  pps->pic_cb_qp_offset = p.pps_cb_qp_offset;
  pps->pic_cr_qp_offset = p.pps_cr_qp_offset;
  pps->loop_filter_across_tiles_enabled_flag = (p.flags & B200_PIC_LF_ACROSS_TILES) != 0;
  pps->tiles_enabled_flag = tile_cols * tile_rows > 1;
  pps->num_tile_columns = tile_cols;
  pps->num_tile_rows = tile_rows;
  pps->uniform_spacing_flag = 1;
  pps->set_derived_values(sps.get());

  de265_image img;
  const de265_chroma chroma = p.chroma_format_idc == 0 ? de265_chroma_mono : p.chroma_format_idc == 1 ? de265_chroma_420 : p.chroma_format_idc == 2 ? de265_chroma_422 : de265_chroma_444;
  if (img.alloc_image(p.width, p.height, chroma, sps, true, filter_ctx(stages & 4), 0, nullptr, false) != DE265_OK) return -2;
  img.set_headers(nullptr, sps, pps);
  img.clear_metadata();
  const int nc = p.chroma_format_idc ? 3 : 1;
  for (int c = 0; c < nc; c++) {
    const int bpp = (c ? p.bit_depth_chroma : p.bit_depth_luma) > 8 ? 2 : 1;
    for (int y = 0; y < img.get_height(c); y++)
      memcpy(img.get_image_plane(c) + (size_t)y * img.get_image_stride(c) * bpp, (const uint8_t*)planes[c] + y * strides[c], (size_t)img.get_width(c) * bpp);
  }
  std::vector<std::unique_ptr<slice_segment_header>> hdrs;
  for (uint32_t i = 0; i < pic->n_slices; i++) {
    const b200_slice_info& s = pic->slices[i];
    hdrs.emplace_back(new slice_segment_header());
    slice_segment_header* h = hdrs.back().get();
    h->SliceAddrRS = s.slice_addr_rs;
    h->slice_segment_address = s.slice_addr_rs;
    h->slice_beta_offset = s.beta_offset;
    h->slice_tc_offset = s.tc_offset;
    h->slice_deblocking_filter_disabled_flag = (s.flags & B200_SLICE_DEBLOCK_DISABLED) != 0;
    h->slice_loop_filter_across_slices_enabled_flag = (s.flags & B200_SLICE_LF_ACROSS_SLICES) != 0;
    h->slice_sao_luma_flag = (s.flags & B200_SLICE_SAO_LUMA) != 0;
    h->slice_sao_chroma_flag = (s.flags & B200_SLICE_SAO_CHROMA) != 0;
    img.add_slice_segment_header(h);
  }
  const int S = 1 << p.log2_ctb_size, wctb = sps->PicWidthInCtbsY, hctb = sps->PicHeightInCtbsY;
  const int w8 = (p.width + 7) / 8, h8 = (p.height + 7) / 8, w4 = (p.width + 3) / 4, h4 = (p.height + 3) / 4;
  for (int cy = 0; cy < hctb; cy++)
    for (int cx = 0; cx < wctb; cx++) {
      const b200_ctb_info& ci = pic->ctbs[cx + cy * wctb];
      if (ci.slice_idx >= pic->n_slices) return -3;
      if (ci.tile_id != pps->scan->TileIdRS[cx + cy * wctb]) return -4;  // the records' tile ids must be the uniform-spacing ones
      img.set_SliceHeaderIndex(cx * S, cy * S, ci.slice_idx);
      img.set_SliceAddrRS(cx, cy, pic->slices[ci.slice_idx].slice_addr_rs);
      sao_info sao;
      sao.SaoTypeIdx = ci.sao_type;
      sao.SaoEoClass = ci.sao_eo_class;
      for (int c = 0; c < 3; c++) {
        sao.sao_band_position[c] = ci.sao_band_pos[c];
        for (int k = 0; k < 4; k++) sao.saoOffsetVal[c][k] = ci.sao_offset[c][k];
      }
      img.set_sao_info(cx, cy, &sao);
    }
  for (int y = 0; y < h8; y++)
    for (int x = 0; x < w8; x++) {
      img.set_QPY(8 * x, 8 * y, 3, pic->qp_map[x + y * w8]);
      if (pic->nofilt_map[x + y * w8] & 1) img.set_cu_transquant_bypass(8 * x, 8 * y, 3, 1);
    }
  if ((stages & 1) && pic->bs_map) {
    for (int pass = 0; pass < 2; pass++) {
      const bool vertical = pass == 0;
      for (int y = 0; y < h4; y++)
        for (int x = 0; x < w4; x++) {
          const uint8_t b = pic->bs_map[x + y * w4];
          img.set_deblk_bS(4 * x, 4 * y, vertical ? B200_BS_V(b) : B200_BS_H(b));
        }
      edge_filtering_luma(&img, vertical, 0, img.get_deblk_height(), 0, img.get_deblk_width());
      if (p.chroma_format_idc) edge_filtering_chroma(&img, vertical, 0, img.get_deblk_height(), 0, img.get_deblk_width());
    }
  }
  if (stages & 2) apply_sample_adaptive_offset_sequential(&img);
  for (int c = 0; c < nc; c++) {
    const int bpp = (c ? p.bit_depth_chroma : p.bit_depth_luma) > 8 ? 2 : 1;
    for (int y = 0; y < img.get_height(c); y++)
      memcpy((uint8_t*)planes[c] + y * strides[c], img.get_image_plane(c) + (size_t)y * img.get_image_stride(c) * bpp, (size_t)img.get_width(c) * bpp);
  }
  img.slices.clear();  // owned by hdrs
  return 0;
}
