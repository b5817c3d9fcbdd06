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
