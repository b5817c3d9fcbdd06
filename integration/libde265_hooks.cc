// libde265_hooks.cc — reference-side binding: turns libde265's "reconstruct while parsing" calls
// into command records for the B200 engine (B2 boundary).
//
// This file is compiled INTO libde265 (it includes libde265's internal headers).  It contains no
// reconstruction arithmetic: it only reads the parser's state (thread_context, slice header, image
// metadata) at the moment the reference would have reconstructed, and forwards it through the C ABI
// in include/b200hevc.h.  Single decode thread only: attaching fails while worker threads run, and a picture decoded with
// worker threads started later is flagged as damaged instead of being recorded concurrently.
//
// Two ways to use it: (a) de265_b200_attach with an application-provided sink (tests, bench.py), or (b) the BUILT-IN backend
// (de265_b200_enable, selected by DE265_DECODER_PARAM_ACCELERATION_CODE = de265_acceleration_B200): it owns a B200 engine,
// submits every picture asynchronously at picture end, reads it back into page-locked picture planes (installed through
// libde265's image allocation plug-in) and waits for that transfer only when the picture is handed to the application, so the
// host parses picture N+1 while the GPU reconstructs picture N.

#include "libde265_hooks.h"

#include <cstdio>
#include <cstdlib>

#include <map>
#include <tuple>
#include <vector>

#include "libde265/decctx.h"
#include "libde265/image.h"
#include "libde265/intrapred.h"
#include "libde265/motion.h"
#include "libde265/slice.h"
#include "libde265/transform.h"

// non-static functions of deblock.cc that are not in deblock.h (deblock.cc:230,243)
bool derive_edgeFlags(de265_image* img);
void derive_boundaryStrength(de265_image* img, bool vertical, int yStart, int yEnd, int xStart, int xEnd);

namespace {

struct builtin_backend;
struct hook_state {
  de265_b200_sink sink = nullptr;
  de265_b200_wait wait = nullptr;
  de265_b200_fill fill = nullptr;
  void* user = nullptr;
  b200_recorder* rec = nullptr;  // the recorder of the picture being parsed (= ring[ring_idx] when the built-in backend runs)
  // The built-in backend submits asynchronously (the engine's planner threads read the records after the hook has returned), so
  // consecutive pictures record into a ring of recorders; a recorder is reused once the picture that used it last has been issued
  // (b200_engine_wait_ticket), which normally happened long before.
  std::vector<b200_recorder*> ring;
  std::vector<unsigned long long> ring_ticket;
  size_t ring_idx = 0;
  unsigned long long last_ticket = 0;  // set by the built-in sink
  const de265_image* cur_img = nullptr;
  uint32_t cur_id = 0;
  bool open = false;
  bool failed = false;  // a recorder call failed or the picture uses something the backend does not implement
  std::map<std::tuple<const void*, int, int, int>, int> weight_cache;
  std::map<const de265_image*, int> pending;  // pictures whose read-back is in flight -> DPB slot
  builtin_backend* builtin = nullptr;
};

int ring_wait(hook_state* st, unsigned long long ticket);  // built-in backend: b200_engine_wait_ticket

void fail_picture(hook_state* st, const char* why)
{
  if (!st->failed) fprintf(stderr, "b200 hook: picture not reconstructed: %s\n", why);
  st->failed = true;
}
#define REC(call) do { if ((call) < 0) fail_picture(st, #call); } while (0)

inline hook_state* state_of(base_context* ctx) { return static_cast<hook_state*>(ctx->b200_state); }

int dpb_slot_of(const base_context* ctx, const de265_image* img)
{
  for (int i = 0; i < B200_MAX_SLOTS; i++)
    if (ctx->get_image((uint16_t)i) == img) return i;
  return -1;
}

void begin_picture_if_needed(hook_state* st, base_context* ctx, de265_image* img)
{
  if (st->open && st->cur_img == img && st->cur_id == img->get_ID()) return;
  const seq_parameter_set& sps = img->get_sps();
  const pic_parameter_set& pps = img->get_pps();
  b200_pic_params p{};
  p.width = (uint16_t)sps.pic_width_in_luma_samples;
  p.height = (uint16_t)sps.pic_height_in_luma_samples;
  p.chroma_format_idc = (uint8_t)sps.chroma_format_idc;
  p.bit_depth_luma = (uint8_t)sps.BitDepth_Y;
  p.bit_depth_chroma = (uint8_t)sps.BitDepth_C;
  p.log2_ctb_size = sps.Log2CtbSizeY;
  p.flags = 0;
  if (sps.sample_adaptive_offset_enabled_flag) p.flags |= B200_PIC_SAO_ENABLED;
  if (sps.strong_intra_smoothing_enable_flag) p.flags |= B200_PIC_STRONG_INTRA_SMOOTHING;
  if (sps.pcm_loop_filter_disable_flag) p.flags |= B200_PIC_PCM_LF_DISABLE;
  if (pps.loop_filter_across_tiles_enabled_flag) p.flags |= B200_PIC_LF_ACROSS_TILES;
  if (sps.range_extension.intra_smoothing_disabled_flag) p.flags |= B200_PIC_INTRA_SMOOTHING_OFF;
  if (sps.scaling_list_enable_flag) p.flags |= B200_PIC_SCALING_LIST;
  p.pps_cb_qp_offset = (int8_t)pps.pic_cb_qp_offset;
  p.pps_cr_qp_offset = (int8_t)pps.pic_cr_qp_offset;
  st->failed = false;
  if (!st->ring.empty()) {
    st->rec = st->ring[st->ring_idx];
    if (st->ring_ticket[st->ring_idx] && ring_wait(st, st->ring_ticket[st->ring_idx]) < 0) fail_picture(st, "an earlier picture failed in the backend");
    st->ring_ticket[st->ring_idx] = 0;
  }
  const int slot = dpb_slot_of(ctx, img);
  if (slot < 0) fail_picture(st, "the picture is not in the DPB (or beyond B200_MAX_SLOTS)");
  p.dst_slot = (uint8_t)(slot < 0 ? 0 : slot);
  p.poc = img->PicOrderCntVal;
  if (pps.range_extension.cross_component_prediction_enabled_flag) fail_picture(st, "RExt cross-component prediction is not implemented");
  if (sps.chroma_format_idc > 1) fail_picture(st, "4:2:2 / 4:4:4 are not implemented on the device");
  if (decoder_context* dc = dynamic_cast<decoder_context*>(ctx))
    if (dc->get_num_worker_threads() > 0) fail_picture(st, "worker threads are running (the recorder is single-threaded)");
  REC(b200_rec_begin_picture(st->rec, &p));
  if (sps.scaling_list_enable_flag) {
    // pps.scaling_list holds the active factors (transform.cc:502-506)
    std::vector<uint8_t> f(B200_SCALING_FACTOR_BYTES);
    uint8_t* d = f.data();
    memcpy(d, pps.scaling_list.ScalingFactor_Size0, 6 * 16); d += 6 * 16;
    memcpy(d, pps.scaling_list.ScalingFactor_Size1, 6 * 64); d += 6 * 64;
    memcpy(d, pps.scaling_list.ScalingFactor_Size2, 6 * 256); d += 6 * 256;
    memcpy(d, pps.scaling_list.ScalingFactor_Size3, 6 * 1024);
    REC(b200_rec_set_scaling_factors(st->rec, f.data()));
  }
  st->cur_img = img;
  st->cur_id = img->get_ID();
  st->open = true;
  st->weight_cache.clear();
}

template <class pixel_t>
uint64_t intra_avail_mask_via_border_computer(const de265_image* img, int xB, int yB, int nT, int cIdx)
{
  // Runs the reference's own availability derivation (intrapred.h:436-633) and packs `available[]`.
  // The sample values it gathers from the (unreconstructed) host planes are ignored.  Slow (it copies 4nT+1 samples per
  // TU); kept as the cross-check of intra_avail_mask below (environment B200_HOOK_CHECK=1).
  pixel_t border_mem[4 * MAX_INTRA_PRED_BLOCK_SIZE + 1];
  intra_border_computer<pixel_t> c;
  c.init(&border_mem[2 * MAX_INTRA_PRED_BLOCK_SIZE], img, nT, cIdx, xB, yB);
  c.preproc();
  c.fill_from_image();
  uint64_t m = 0;
  for (int k = 0; k < nT / 2; k++) {
    if (c.available[-4 * k - 1]) m |= 1ull << k;
    if (c.available[4 * k + 1]) m |= 1ull << (B200_AVAIL_TOP_BIT0 + k);
  }
  if (c.available[0]) m |= 1ull << B200_AVAIL_CORNER_BIT;
  return m;
}

// The same decisions without touching a sample: the side flags of intra_border_computer::preproc (intrapred.h:436-526:
// picture border, slice address and tile id of the neighbouring CTBs) and, per group of 4 border samples, the tests of
// fill_from_image (intrapred.h:529-633: the neighbour's minimum-TB z-scan address must not exceed the current block's;
// with constrained_intra_pred the neighbour must be intra).  One group = one bit of b200_tu.avail.
uint64_t intra_avail_mask(const de265_image* img, int xB, int yB, int nT, int cIdx)
{
  const seq_parameter_set& sps = img->get_sps();
  const pic_parameter_set& pps = img->get_pps();
  const int SW = (cIdx == 0) ? 1 : sps.SubWidthC, SH = (cIdx == 0) ? 1 : sps.SubHeightC;
  const int xL = xB * SW, yL = yB * SH;
  bool aL = true, aT = true, aTR = true, aTL = true;
  if (xL == 0) aL = aTL = false;
  if (yL == 0) aT = aTL = aTR = false;
  if (xL + nT * SW >= sps.pic_width_in_luma_samples) aTR = false;
  const int l2c = sps.Log2CtbSizeY, wC = sps.PicWidthInCtbsY;
  const int xC = xL >> l2c, yC = yL >> l2c, xLc = (xL - 1) >> l2c, xRc = (xL + nT * SW) >> l2c, yTc = (yL - 1) >> l2c;
  const int sl = img->get_SliceAddrRS(xC, yC);
  const uint32_t tile = pps.scan->TileIdRS[xC + yC * wC];
  auto same = [&](int cx, int cy) { return img->get_SliceAddrRS(cx, cy) == sl && pps.scan->TileIdRS[cx + cy * wC] == tile; };
  if (aL && !same(xLc, yC)) aL = false;
  if (aT && !same(xC, yTc)) aT = false;
  if (aTL && !same(xLc, yTc)) aTL = false;
  if (aTR && !same(xRc, yTc)) aTR = false;
  int nBottom = (sps.pic_height_in_luma_samples - yL + SH - 1) / SH;
  if (nBottom > 2 * nT) nBottom = 2 * nT;
  int nRight = (sps.pic_width_in_luma_samples - xL + SW - 1) / SW;
  if (nRight > 2 * nT) nRight = 2 * nT;
  const int l2t = sps.Log2MinTrafoSize, wT = sps.PicWidthInTbsY;
  const int cur = pps.scan->MinTbAddrZS[(xL >> l2t) + (yL >> l2t) * wT];
  const bool cip = pps.constrained_intra_pred_flag;
  auto ok = [&](int xs, int ys) {  // component coordinates of a neighbouring sample
    const int xl = xs * SW, yl = ys * SH;
    if (pps.scan->MinTbAddrZS[(xl >> l2t) + (yl >> l2t) * wT] > cur) return false;
    return !cip || img->get_pred_mode(xl, yl) == MODE_INTRA;
  };
  uint64_t m = 0;
  if (aL)
    for (int y = nBottom - 1; y >= 0; y -= 4)
      if (ok(xB - 1, yB + y)) m |= 1ull << (y >> 2);
  if (aTL && ok(xB - 1, yB - 1)) m |= 1ull << B200_AVAIL_CORNER_BIT;
  for (int x = 0; x < nRight; x += 4)
    if (((x < nT) ? aT : aTR) && ok(xB + x, yB - 1)) m |= 1ull << (B200_AVAIL_TOP_BIT0 + (x >> 2));
  return m;
}

}  // namespace

extern "C" int de265_b200_attach(void* de265_decoder_ctx, de265_b200_sink sink, void* user)
{
  decoder_context* ctx = static_cast<decoder_context*>(de265_decoder_ctx);
  hook_state* st = state_of(ctx);
  if (!sink) {
    if (st) {
      if (st->ring.empty()) b200_rec_destroy(st->rec);
      for (b200_recorder* r : st->ring) b200_rec_destroy(r);
      delete st;
      ctx->b200_state = nullptr;
    }
    return B200_OK;
  }
  if (ctx->get_num_worker_threads() > 0) return B200_ERR_UNSUPPORTED;  // recording is single-threaded (slice.cc hook sites run under WPP tasks)
  if (!st) {
    st = new hook_state();
    if (b200_rec_create(&st->rec) < 0) { delete st; return B200_ERR_NOMEM; }
    ctx->b200_state = st;
  }
  st->sink = sink;
  st->user = user;
  return B200_OK;
}

extern "C" void de265_b200_set_callbacks(void* de265_decoder_ctx, de265_b200_wait wait, de265_b200_fill fill)
{
  hook_state* st = state_of(static_cast<decoder_context*>(de265_decoder_ctx));
  if (!st) return;
  st->wait = wait;
  st->fill = fill;
}

bool b200_hook_decode_TU(thread_context* tctx, int x0, int y0, int nT, int cIdx, int cuPredMode, bool cbf)
{
  hook_state* st = state_of(tctx->decctx);
  if (!st) return false;
  de265_image* img = tctx->img;
  begin_picture_if_needed(st, tctx->decctx, img);
  const seq_parameter_set& sps = img->get_sps();
  const bool intra = (cuPredMode == MODE_INTRA);
  if (!intra && !cbf) return true;  // nothing to do (cross-component prediction is not supported)

  b200_tu tu{};
  tu.x = (uint16_t)x0;
  tu.y = (uint16_t)y0;
  tu.log2_size = (uint8_t)Log2(nT);
  tu.cidx = (uint8_t)cIdx;
  int rdpcm = 0;
  if (intra) {
    // slice.cc:3471-3498
    int mode = (cIdx == 0) ? img->get_IntraPredMode(x0, y0) : img->get_IntraPredModeC(x0 * sps.SubWidthC, y0 * sps.SubHeightC);
    if (mode < 0 || mode >= 35) mode = INTRA_DC;
    tu.intra_mode = (uint8_t)mode;
    tu.flags |= B200_TU_INTRA;
    tu.avail = intra_avail_mask(img, x0, y0, nT, cIdx);
    static const bool check = getenv("B200_HOOK_CHECK") != nullptr;
    if (check) {
      const uint64_t ref = img->high_bit_depth(cIdx) ? intra_avail_mask_via_border_computer<uint16_t>(img, x0, y0, nT, cIdx)
                                                     : intra_avail_mask_via_border_computer<uint8_t>(img, x0, y0, nT, cIdx);
      if (ref != tu.avail) {
        fprintf(stderr, "b200 hook: availability mask mismatch at (%d,%d) nT %d cIdx %d: %llx vs reference %llx\n", x0, y0, nT, cIdx,
                (unsigned long long)tu.avail, (unsigned long long)ref);
        abort();
      }
    }
    if (sps.range_extension.implicit_rdpcm_enabled_flag && img->get_cu_transquant_bypass(x0, y0))  // intrapred.cc:308-310
      tu.flags |= B200_TU_NO_BOUNDARY_FILTER;
    if (sps.range_extension.implicit_rdpcm_enabled_flag && (tctx->cu_transquant_bypass_flag || tctx->transform_skip_flag[cIdx]) &&
        (mode == 10 || mode == 26))
      rdpcm = (mode == 26) ? 2 : 1;
  } else if (tctx->explicit_rdpcm_flag) {
    rdpcm = tctx->explicit_rdpcm_dir ? 2 : 1;
  }
  int n = 0;
  if (cbf) {
    // transform.cc:361-448: flags exactly as scale_coefficients_internal derives them
    tu.flags |= B200_TU_CBF;
    tu.qp = (uint8_t)(cIdx == 0 ? tctx->qPYPrime : cIdx == 1 ? tctx->qPCbPrime : tctx->qPCrPrime);
    const bool cu_intra_at_xT = (img->get_pred_mode(x0, y0) == MODE_INTRA);  // transform.cc:400 (component coordinates, literally)
    if (tctx->cu_transquant_bypass_flag) tu.flags |= B200_TU_BYPASS;
    else if (tctx->transform_skip_flag[cIdx]) tu.flags |= B200_TU_TSKIP;
    if (rdpcm == 1) tu.flags |= B200_TU_RDPCM_H;
    if (rdpcm == 2) tu.flags |= B200_TU_RDPCM_V;
    if (nT == 4 && cIdx == 0 && cu_intra_at_xT) tu.flags |= B200_TU_DST;
    if (sps.range_extension.transform_skip_rotation_enabled_flag && nT == 4 && cu_intra_at_xT) tu.flags |= B200_TU_ROTATE;
    if (sps.scaling_list_enable_flag && !tctx->cu_transquant_bypass_flag) {
      tu.flags |= B200_TU_SCALING_LIST;
      if (!intra) tu.flags |= B200_TU_INTER_MATRIX;
    }
    n = tctx->nCoeff[cIdx];
  }
  REC(b200_rec_add_tu(st->rec, &tu, tctx->coeffList[cIdx], tctx->coeffPos[cIdx], n));
  return true;
}

void b200_hook_pcm(thread_context* tctx, int x0, int y0, int w, int h, int cIdx)
{
  // called at the end of read_pcm_samples_internal (slice.cc:4211-4255) with component coordinates;
  // the shifted samples are taken back from the host plane the parser just wrote.
  hook_state* st = state_of(tctx->decctx);
  if (!st) return;
  de265_image* img = tctx->img;
  begin_picture_if_needed(st, tctx->decctx, img);
  std::vector<int16_t> lv((size_t)w * h), ps((size_t)w * h);
  const int stride = img->get_image_stride(cIdx);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int v = img->high_bit_depth(cIdx) ? img->get_image_plane_at_pos_NEW<uint16_t>(cIdx, x0, y0)[x + y * stride]
                                        : img->get_image_plane_at_pos_NEW<uint8_t>(cIdx, x0, y0)[x + y * stride];
      lv[x + y * w] = (int16_t)v;
      ps[x + y * w] = (int16_t)(x + y * w);
    }
  b200_tu tu{};
  tu.x = (uint16_t)x0;
  tu.y = (uint16_t)y0;
  tu.log2_size = (uint8_t)Log2(w);
  tu.cidx = (uint8_t)cIdx;
  tu.flags = B200_TU_PCM;
  REC(b200_rec_add_tu(st->rec, &tu, lv.data(), ps.data(), w * h));
}

bool b200_hook_inter_pred(base_context* ctx, const slice_segment_header* shdr, de265_image* img, int xP, int yP, int nPbW, int nPbH,
                          const PBMotion* vi)
{
  hook_state* st = state_of(ctx);
  if (!st) return false;
  begin_picture_if_needed(st, ctx, img);
  const pic_parameter_set* pps = shdr->pps.get();
  const seq_parameter_set* sps = pps->sps.get();
  // motion.cc:303-318: mismatching SPS -> the reference predicts nothing
  if (sps->BitDepth_Y != img->get_bit_depth(0) || sps->BitDepth_C != img->get_bit_depth(1) ||
      sps->chroma_format_idc != img->get_chroma_format())
    return true;

  int predFlag[2] = {vi->predFlag[0], vi->predFlag[1]};
  if (pps->weighted_pred_flag == 0 && predFlag[0] && predFlag[1] && vi->mv[0].x == vi->mv[1].x && vi->mv[0].y == vi->mv[1].y &&
      shdr->RefPicList[0][vi->refIdx[0]] == shdr->RefPicList[1][vi->refIdx[1]])
    predFlag[1] = 0;  // motion.cc:348-357

  b200_pu pu{};
  pu.x = (uint16_t)xP;
  pu.y = (uint16_t)yP;
  pu.w = (uint8_t)nPbW;
  pu.h = (uint8_t)nPbH;
  pu.ref_slot[0] = pu.ref_slot[1] = -1;
  for (int l = 0; l < 2; l++) {
    if (!predFlag[l]) continue;
    int idx = shdr->RefPicList[l][vi->refIdx[l]];
    const de265_image* ref = ctx->get_image((uint16_t)idx);
    // motion.cc:380-408: any of these -> mid-grey prediction
    bool ok = ref && ref->PicState != UnusedForReference && ref->get_width(0) == sps->pic_width_in_luma_samples &&
              ref->get_height(0) == sps->pic_height_in_luma_samples && img->get_chroma_format() == ref->get_chroma_format() &&
              img->get_bit_depth(0) == ref->get_bit_depth(0) && img->get_bit_depth(1) == ref->get_bit_depth(1);
    pu.ref_slot[l] = (ok && idx >= 0 && idx < B200_MAX_SLOTS) ? (int8_t)idx : (int8_t)-1;
    pu.mv[l][0] = vi->mv[l].x;
    pu.mv[l][1] = vi->mv[l].y;
    pu.flags |= (l == 0) ? B200_PU_PRED_L0 : B200_PU_PRED_L1;
  }
  bool weighted;
  if (shdr->slice_type == SLICE_TYPE_P) {
    if (!(predFlag[0] == 1 && predFlag[1] == 0)) return true;  // motion.cc:512-516: warning, nothing written
    weighted = pps->weighted_pred_flag;
  } else {
    if (!predFlag[0] && !predFlag[1]) return true;  // motion.cc:690-693
    weighted = pps->weighted_bipred_flag;
  }
  if (weighted) {
    // motion.cc:518-529, 585-600, 640-650
    const int r0 = predFlag[0] ? vi->refIdx[0] : 0, r1 = predFlag[1] ? vi->refIdx[1] : 0;
    auto key = std::make_tuple((const void*)shdr, predFlag[0] ? r0 : -1, predFlag[1] ? r1 : -1, 0);
    auto it = st->weight_cache.find(key);
    int wi;
    if (it != st->weight_cache.end()) {
      wi = it->second;
    } else {
      const int shift1_L = std::max(2, 14 - sps->BitDepth_Y), shift1_C = std::max(2, 14 - sps->BitDepth_C);
      b200_weight_entry we{};
      we.log2wd_luma = (uint8_t)(shdr->luma_log2_weight_denom + shift1_L);
      we.log2wd_chroma = (uint8_t)(shdr->ChromaLog2WeightDenom + shift1_C);
      const int ridx[2] = {r0, r1};
      for (int l = 0; l < 2; l++) {
        we.w[l][0] = shdr->LumaWeight[l][ridx[l]];
        we.o[l][0] = (int16_t)(shdr->luma_offset[l][ridx[l]] * (1 << sps->WpOffsetBdShiftY));
        for (int c = 0; c < 2; c++) {
          we.w[l][1 + c] = shdr->ChromaWeight[l][ridx[l]][c];
          we.o[l][1 + c] = (int16_t)(shdr->ChromaOffset[l][ridx[l]][c] * (1 << sps->WpOffsetBdShiftC));
        }
      }
      wi = b200_rec_add_weights(st->rec, &we);
      if (wi < 0) { fail_picture(st, "b200_rec_add_weights"); wi = 0; }
      st->weight_cache[key] = wi;
    }
    pu.flags |= B200_PU_WEIGHTED;
    pu.wt_idx = (uint16_t)wi;
  }
  REC(b200_rec_add_pu(st->rec, &pu));
  return true;
}

bool b200_hook_picture_done(decoder_context* ctx, de265_image* img)
{
  hook_state* st = state_of(ctx);
  if (!st) return false;
  begin_picture_if_needed(st, ctx, img);  // pictures without any record (cannot happen in valid streams)
  const seq_parameter_set& sps = img->get_sps();
  const pic_parameter_set& pps = img->get_pps();
  const int W = sps.pic_width_in_luma_samples, H = sps.pic_height_in_luma_samples;
  const int w4 = (W + 3) / 4, h4 = (H + 3) / 4, w8 = (W + 7) / 8, h8 = (H + 7) / 8;

  // slices, CTBs (image.h:160-170, slice.h:268-276)
  for (size_t i = 0; i < img->slices.size(); i++) {
    const slice_segment_header* sh = img->slices[i];
    b200_slice_info s{};
    s.slice_addr_rs = sh->SliceAddrRS;
    s.beta_offset = sh->slice_beta_offset;
    s.tc_offset = sh->slice_tc_offset;
    if (sh->slice_deblocking_filter_disabled_flag) s.flags |= B200_SLICE_DEBLOCK_DISABLED;
    if (sh->slice_loop_filter_across_slices_enabled_flag) s.flags |= B200_SLICE_LF_ACROSS_SLICES;
    if (sh->slice_sao_luma_flag) s.flags |= B200_SLICE_SAO_LUMA;
    if (sh->slice_sao_chroma_flag) s.flags |= B200_SLICE_SAO_CHROMA;
    REC(b200_rec_add_slice(st->rec, &s));
  }
  for (int cy = 0; cy < sps.PicHeightInCtbsY; cy++)
    for (int cx = 0; cx < sps.PicWidthInCtbsY; cx++) {
      b200_ctb_info c{};
      c.slice_idx = img->get_SliceHeaderIndexCtb(cx, cy);
      c.tile_id = (uint16_t)pps.scan->TileIdRS[cx + cy * sps.PicWidthInCtbsY];
      const sao_info* si = img->get_sao_info(cx, cy);
      c.sao_type = si->SaoTypeIdx;
      c.sao_eo_class = si->SaoEoClass;
      for (int k = 0; k < 3; k++) {
        c.sao_band_pos[k] = si->sao_band_position[k];
        for (int j = 0; j < 4; j++) c.sao_offset[k][j] = si->saoOffsetVal[k][j];
      }
      REC(b200_rec_set_ctb(st->rec, cx, cy, &c));
    }

  // QP / no-filter maps at 8x8 granularity (deblock.cc:513-515,576-592; sao.cc:112-117)
  int8_t* qp = b200_rec_qp_map(st->rec);
  uint8_t* nf = b200_rec_nofilt_map(st->rec);
  for (int y = 0; y < h8; y++)
    for (int x = 0; x < w8; x++) {
      qp[x + y * w8] = (int8_t)img->get_QPY(x * 8, y * 8);
      nf[x + y * w8] = ((sps.pcm_loop_filter_disable_flag && img->get_pcm_flag(x * 8, y * 8)) || img->get_cu_transquant_bypass(x * 8, y * 8)) ? 1 : 0;
    }

  // edge flags + boundary strength on the host (deblock.cc:132-383), both directions
  b200_picture pic{};
  bool deblock = !ctx->param_disable_deblocking && derive_edgeFlags(img);
  if (deblock) {
    uint8_t* bs = b200_rec_bs_map(st->rec);
    derive_boundaryStrength(img, true, 0, img->get_deblk_height(), 0, img->get_deblk_width());
    for (int y = 0; y < h4; y++)
      for (int x = 0; x < w4; x += 2) bs[x + y * w4] |= img->get_deblk_bS(x * 4, y * 4) & 3;
    derive_boundaryStrength(img, false, 0, img->get_deblk_height(), 0, img->get_deblk_width());
    for (int y = 0; y < h4; y += 2)
      for (int x = 0; x < w4; x++) bs[x + y * w4] |= (img->get_deblk_bS(x * 4, y * 4) & 3) << 2;
  }
  REC(b200_rec_end_picture(st->rec, &pic));
  st->open = false;
  if (st->failed) {  // nothing trustworthy to hand to the backend: the picture stays as it is and is flagged (decctx.cc:615)
    img->integrity = INTEGRITY_DECODING_ERRORS;
    return true;
  }
  if (!deblock) pic.params.flags |= B200_PIC_SKIP_DEBLOCK;
  if (ctx->param_disable_sao) pic.params.flags |= B200_PIC_SKIP_SAO;

  void* planes[3] = {img->get_image_plane(0), img->get_image_plane(1), img->get_image_plane(2)};
  size_t strides[3];
  for (int c = 0; c < 3; c++) strides[c] = (size_t)img->get_image_stride(c) * ((img->get_bit_depth(c) + 7) / 8);
  if (img->get_chroma_format() == de265_chroma_mono) planes[1] = planes[2] = nullptr;
  const int rc = st->sink(st->user, &pic, planes, strides);
  if (!st->ring.empty()) {  // the records stay untouched until this picture has been issued
    st->ring_ticket[st->ring_idx] = st->last_ticket;
    st->ring_idx = (st->ring_idx + 1) % st->ring.size();
    st->rec = st->ring[st->ring_idx];
  }
  if (rc < 0) img->integrity = INTEGRITY_DECODING_ERRORS;
  else if (rc == DE265_B200_SINK_PENDING) {
    st->pending[img] = pic.params.dst_slot;
    if (ctx->param_sei_check_hash) b200_hook_wait_image(ctx, img);  // the decoded-picture-hash check reads the host planes right away
  }
  return true;
}

void b200_hook_wait_image(decoder_context* ctx, const de265_image* img)
{
  hook_state* st = state_of(ctx);
  if (!st || !img) return;
  auto it = st->pending.find(img);
  if (it == st->pending.end()) return;
  if (st->wait && st->wait(st->user, it->second) < 0) const_cast<de265_image*>(img)->integrity = INTEGRITY_DECODING_ERRORS;
  st->pending.erase(it);
}

// libde265 fills a synthesised reference picture on the host (decctx.cc:1294-1318): mirror it into the backend's DPB slot,
// otherwise pictures predicted from it would read whatever the slot held before.
void b200_hook_unavailable_reference(decoder_context* ctx, de265_image* img)
{
  hook_state* st = state_of(ctx);
  if (!st || !st->fill || !img) return;
  const int slot = dpb_slot_of(ctx, img);
  if (slot < 0) return;
  const seq_parameter_set& sps = img->get_sps();
  b200_pic_params p{};
  p.width = (uint16_t)sps.pic_width_in_luma_samples;
  p.height = (uint16_t)sps.pic_height_in_luma_samples;
  p.chroma_format_idc = (uint8_t)sps.chroma_format_idc;
  p.bit_depth_luma = (uint8_t)sps.BitDepth_Y;
  p.bit_depth_chroma = (uint8_t)sps.BitDepth_C;
  p.log2_ctb_size = sps.Log2CtbSizeY;
  p.dst_slot = (uint8_t)slot;
  st->fill(st->user, slot, &p, 1 << (sps.BitDepth_Y - 1), 1 << (sps.BitDepth_C - 1));
}

// ---- the built-in backend -------------------------------------------------------------------------------------
namespace {
struct builtin_backend {
  b200_engine* eng = nullptr;
  hook_state* st = nullptr;
  bool sync_submit = true;
  std::multimap<size_t, void*> pool;  // free page-locked planes by size: libde265 allocates the planes of every new picture
  std::map<void*, size_t> live;
};

int ring_wait(hook_state* st, unsigned long long ticket)
{
  return (st->builtin && st->builtin->eng) ? b200_engine_wait_ticket(st->builtin->eng, ticket) : 0;
}

int builtin_sink(void* user, const b200_picture* pic, void* const planes[3], const size_t strides[3])
{
  builtin_backend* be = static_cast<builtin_backend*>(user);
  // B200_HOOK_ASYNC=1: validation / planning / packing run on the engine's planner threads while the parser goes on with the next picture
  // (default: the synchronous call, planned on the engine's pool before the hook returns; B200_HOOK_ASYNC=1: queued)
  int rc = be->sync_submit ? b200_engine_submit_picture(be->eng, pic) : b200_engine_submit_picture_async(be->eng, pic);
  if (rc < 0) { fprintf(stderr, "b200 backend: %s\n", b200_last_error()); return rc; }
  if (be->st) be->st->last_ticket = be->sync_submit ? 0 : b200_engine_last_ticket(be->eng);
  rc = b200_engine_read_slot_async(be->eng, pic->params.dst_slot, planes, strides);
  if (rc < 0) { fprintf(stderr, "b200 backend: %s\n", b200_last_error()); return rc; }
  return DE265_B200_SINK_PENDING;
}
int builtin_wait(void* user, int slot)
{
  const int rc = b200_engine_wait_slot(static_cast<builtin_backend*>(user)->eng, slot);
  if (rc < 0) fprintf(stderr, "b200 backend: %s\n", b200_last_error());
  return rc;
}
int builtin_fill(void* user, int slot, const b200_pic_params* p, int vy, int vc)
{
  return b200_engine_fill_slot(static_cast<builtin_backend*>(user)->eng, slot, p, vy, vc);
}

// de265_image_allocation plug-in (de265.h:350-365): page-locked planes with the reference's own layout rules (image.cc:110-160)
int builtin_get_buffer(de265_decoder_context*, de265_image_spec* spec, de265_image* img, void* userdata)
{
  builtin_backend* be = static_cast<builtin_backend*>(userdata);
  const int bpp_y = (img->BitDepth_Y + 7) / 8, bpp_c = (img->BitDepth_C + 7) / 8;
  const int align = spec->alignment > 0 ? spec->alignment : 16;
  const bool mono = img->get_chroma_format() == de265_chroma_mono;
  const int cw = mono ? 0 : spec->width / img->SubWidthC, chh = mono ? 0 : spec->height / img->SubHeightC;
  const int ls = (spec->width + align - 1) / align * align, cs = mono ? 0 : (cw + align - 1) / align * align;
  const size_t sizes[3] = {(size_t)ls * bpp_y * spec->height + 64, (size_t)cs * bpp_c * chh + 64, (size_t)cs * bpp_c * chh + 64};
  void* p[3] = {nullptr, nullptr, nullptr};
  for (int c = 0; c < (mono ? 1 : 3); c++) {
    auto it = be->pool.find(sizes[c]);
    if (it != be->pool.end()) { p[c] = it->second; be->pool.erase(it); }
    else p[c] = b200_host_alloc(sizes[c]);
    if (!p[c]) {
      for (int k = 0; k < c; k++) { be->live.erase(p[k]); be->pool.emplace(sizes[k], p[k]); }
      return 0;
    }
    be->live[p[c]] = sizes[c];
  }
  img->set_image_plane(0, (uint8_t*)p[0], ls, nullptr);
  img->set_image_plane(1, (uint8_t*)p[1], cs, nullptr);
  img->set_image_plane(2, (uint8_t*)p[2], cs, nullptr);
  return 1;
}
void builtin_release_buffer(de265_decoder_context*, de265_image* img, void* userdata)
{
  builtin_backend* be = static_cast<builtin_backend*>(userdata);
  for (int c = 0; c < 3; c++) {
    void* p = img->get_image_plane(c);
    auto it = be->live.find(p);
    if (it == be->live.end()) continue;
    be->pool.emplace(it->second, p);  // kept for the next picture: page-locking is expensive
    be->live.erase(it);
  }
}
}  // namespace

extern "C" int de265_b200_enable(void* de265_decoder_ctx, int device)
{
  decoder_context* ctx = static_cast<decoder_context*>(de265_decoder_ctx);
  if (hook_state* st = state_of(ctx))
    if (st->builtin) return B200_OK;
  builtin_backend* be = new builtin_backend();
  int rc = b200_engine_create(&be->eng, device);
  if (rc < 0) { delete be; return rc; }
  rc = de265_b200_attach(ctx, builtin_sink, be);
  if (rc < 0) { b200_engine_destroy(be->eng); delete be; return rc; }
  de265_b200_set_callbacks(ctx, builtin_wait, builtin_fill);
  {
    hook_state* st = state_of(ctx);
    st->builtin = be;
    be->st = st;
    // B200_HOOK_ASYNC=1: queue the pictures (b200_engine_submit_picture_async) and record into a ring of recorders; pays on long
    // streams (the parse thread no longer plans), costs three more recorders' worth of first-touch memory on short ones.  Measured on
    // 16 concatenated copies of the golden intra streams: 1080p 44.4 vs 41.7-46.7 frames/s, 4K 24.1 vs 27.4: the default stays synchronous
    if (const char* e = getenv("B200_HOOK_ASYNC")) be->sync_submit = atoi(e) == 0;
    st->ring.push_back(st->rec);  // the recorder de265_b200_attach created + three more
    for (int i = 0; i < (be->sync_submit ? 0 : 3); i++) {
      b200_recorder* r = nullptr;
      if (b200_rec_create(&r) < 0) break;
      st->ring.push_back(r);
    }
    st->ring_ticket.assign(st->ring.size(), 0);
    st->ring_idx = 0;
  }
  de265_image_allocation alloc = {builtin_get_buffer, builtin_release_buffer};
  ctx->set_image_allocation_functions(&alloc, be);
  return B200_OK;
}

extern "C" void de265_b200_disable(void* de265_decoder_ctx)
{
  decoder_context* ctx = static_cast<decoder_context*>(de265_decoder_ctx);
  hook_state* st = state_of(ctx);
  if (!st || !st->builtin) return;
  builtin_backend* be = st->builtin;
  b200_engine_sync(be->eng);
  be->st = nullptr;
  de265_b200_attach(ctx, nullptr, nullptr);
  b200_engine_destroy(be->eng);
  be->eng = nullptr;
  // The allocation plug-in and `be` stay: pictures still alive were allocated through it and are released through it
  // (image.cc release() passes the context's CURRENT userdata); their page-locked planes return to the pool.
}

// DE265_DECODER_PARAM_ACCELERATION_CODE = de265_acceleration_B200 (de265.cc:576-578 -> base_context::set_acceleration_functions)
void b200_hook_set_acceleration(base_context* bctx, int level)
{
  decoder_context* ctx = dynamic_cast<decoder_context*>(bctx);
  if (!ctx) return;
  if (level == DE265_ACCELERATION_B200) {
    const char* dev = getenv("B200_DEVICE");
    const int rc = de265_b200_enable(ctx, dev ? atoi(dev) : 0);
    if (rc < 0) fprintf(stderr, "libde265: de265_acceleration_B200 unavailable (%s); the host tables stay in place\n", b200_last_error());
  } else if (hook_state* st = state_of(ctx)) {
    if (st->builtin) de265_b200_disable(ctx);
  }
}
