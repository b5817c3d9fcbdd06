// kernels_mc8.cuh — 8-bit inter prediction with packed integer dot products (the roofline-graded kernel).
//
// Same contract as k_inter_pred<uint8_t> (kernels_mc.cuh): one warp per <=16x16 luma tile (+ its 8x8 Cb/Cr),
// luma 8-tap / chroma 4-tap separable interpolation (motion.cc:48-282, fallback-motion.cc:262-636) fused with the
// four weighting modes (fallback-motion.cc:33-256).  What changes is the arithmetic:
//
//   pass 1 (horizontal, on bytes): a lane produces 4 adjacent outputs of one row from 12 source bytes held in three
//     32-bit registers (aligned 16-byte load + funnel shifts); output j is  dp4a(b0,T[j][0]) + dp4a(b1,T[j][1]) +
//     dp4a(b2,T[j][2])  with the 8 taps pre-shifted by j bytes into T (11 dp4a per 4 outputs instead of 32 MACs).
//     Results (|v| < 2^15, no shift at 8 bit) go to a per-warp shared-memory strip stored COLUMN-major, so that two
//     vertically adjacent samples share a 32-bit word.
//   pass 2 (vertical, on int16 pairs): a lane produces 8 rows of one column from 8 words of the strip with dp2a
//     (4 dp2a for even rows, 5 for odd rows: the taps are pre-packed for both parities), then >> 6, int16 wrap
//     (SURVEY App. A.1), weighting, clip, byte stores.
//   Integer phases use the identity tap, so there is one code path; a zero vertical phase skips pass 2 and its 7 halo rows.
//   Tiles whose reference window crosses the left/right picture edge build their 12 bytes from clamped loads
//   (motion.cc:147-153); rows are always clamped.
//
// Tap tables live in constant memory and are built on the host from the HEVC filter taps (engine.cu init_tables).
#pragma once
#include "dev_common.cuh"
#include "kernels_mc.cuh"

#define MC8_CS 26   // luma strip: int16 column stride (>= 23 rows, even, spreads columns over banks)
#define MC8_CCS 12  // chroma strip column stride (>= 11 rows)
#define MC8_STRIP (16 * MC8_CS + 2 * 8 * MC8_CCS)

// [frac 0..3, 4 = integer position with gain 64][output j][word]
__constant__ uint32_t c_qh[5][4][3];
__constant__ uint32_t c_qv[4][5];      // [frac][A,B (even rows), C,D,E (odd rows)]
__constant__ uint32_t c_eh[9][4][2];   // [frac 0..7, 8 = integer with gain 64][output j][word]
__constant__ uint32_t c_ev[8][3];      // [frac][A (even), C,D (odd)]

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)
{
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_lo_ss(uint32_t a, uint32_t b, int c)
{
  int d;
  asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi_ss(uint32_t a, uint32_t b, int c)
{
  int d;
  asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Source bytes starting at column xb of row `row` (row pointer is 256-byte aligned; pitch >= width + 16).
template <int NW>  // NW = number of 32-bit words of source bytes wanted (3 for luma: 12 bytes, 2 for chroma: 8 bytes)
__device__ __forceinline__ void load_bytes(const uint8_t* row, int xb, int pw, bool clampx, uint32_t (&b)[3])
{
  if (!clampx) {
    const int xa = xb & ~3, sh = (xb & 3) * 8;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(row + xa);
    const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
    b[0] = __funnelshift_r(w0, w1, sh);
    b[1] = __funnelshift_r(w1, w2, sh);
    if (NW == 3) {
      const uint32_t w3 = p[3];
      b[2] = __funnelshift_r(w2, w3, sh);
    }
  } else {
#pragma unroll
    for (int k = 0; k < NW; k++) {
      uint32_t v = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) v |= (uint32_t)row[clip3i(0, pw - 1, xb + 4 * k + i)] << (8 * i);
      b[k] = v;
    }
  }
}

// pass 1 for one plane of one list: strip[col * CS + r] = sum_k taps[k] * ref[y][x + col + k - before]
template <int CS, bool LUMA>
__device__ __forceinline__ void mc8_hpass(int16_t* strip, const uint8_t* ref, int pitch, int pw, int ph, int x_int, int y_int, int tw, int nrows,
                                          int before_rows, const uint32_t* taps /* [4][LUMA ? 3 : 2] */, int r0, int rstep, int g, bool active)
{
  constexpr int BEFORE = LUMA ? 3 : 1;
  constexpr int NW = LUMA ? 3 : 2;
  const int ng = (tw + 3) >> 2;
  const int xb = x_int + 4 * g - BEFORE;
  // the whole window of the tile is inside the picture horizontally?
  const bool clampx = (x_int - BEFORE < 0) || (x_int + tw + (LUMA ? 4 : 2) > pw - 1);
  uint32_t t[4][NW];
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int k = 0; k < NW; k++) t[j][k] = taps[j * NW + k];
  if (!active || g >= ng) return;
  for (int r = r0; r < nrows; r += rstep) {
    const int ya = clip3i(0, ph - 1, y_int + r - before_rows);
    uint32_t b[3];
    load_bytes<NW>(ref + (size_t)ya * pitch, xb, pw, clampx, b);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int s = dp4a_us(b[0], t[j][0], 0);
      s = dp4a_us(b[1], t[j][1], s);
      if (LUMA && j > 0) s = dp4a_us(b[2], t[j][2], s);
      if (4 * g + j < tw) strip[(4 * g + j) * CS + r] = (int16_t)s;
    }
  }
}

// pass 2 for NOUT consecutive rows of one column; w[] = the strip words of that column starting at the first output row
// (row pair per word).  LUMA: NOUT = 8, 8 words; chroma: NOUT = 4, 4 words.  Returns int16-wrapped values.
template <bool LUMA>
__device__ __forceinline__ void mc8_vpass(const uint32_t* w, const uint32_t* tv, int shift, int* out)
{
  if (LUMA) {
    const uint32_t A = tv[0], B = tv[1], C = tv[2], D = tv[3], E = tv[4];
#pragma unroll
    for (int m = 0; m < 4; m++) {
      int e = dp2a_lo_ss(w[m], A, 0);
      e = dp2a_hi_ss(w[m + 1], A, e);
      e = dp2a_lo_ss(w[m + 2], B, e);
      e = dp2a_hi_ss(w[m + 3], B, e);
      int o = dp2a_lo_ss(w[m], C, 0);
      o = dp2a_hi_ss(w[m + 1], C, o);
      o = dp2a_lo_ss(w[m + 2], D, o);
      o = dp2a_hi_ss(w[m + 3], D, o);
      o = dp2a_lo_ss(w[m + 4], E, o);
      out[2 * m] = (int)(int16_t)(e >> shift);
      out[2 * m + 1] = (int)(int16_t)(o >> shift);
    }
  } else {
    const uint32_t A = tv[0], C = tv[1], D = tv[2];
#pragma unroll
    for (int m = 0; m < 2; m++) {
      int e = dp2a_lo_ss(w[m], A, 0);
      e = dp2a_hi_ss(w[m + 1], A, e);
      int o = dp2a_lo_ss(w[m], C, 0);
      o = dp2a_hi_ss(w[m + 1], C, o);
      o = dp2a_lo_ss(w[m + 2], D, o);
      out[2 * m] = (int)(int16_t)(e >> shift);
      out[2 * m + 1] = (int)(int16_t)(o >> shift);
    }
  }
}

__global__ void __launch_bounds__(128) k_inter_pred8(DevPic pic, RefTable refs, const b200_pu* __restrict__ pus,
                                                     const b200_weight_entry* __restrict__ wts, const uint32_t* __restrict__ tiles, int n_tiles)
{
  __shared__ __align__(16) int16_t s_strip[4][2][MC8_STRIP + 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * 4 + warp;
  if (tile >= n_tiles) return;
  const uint32_t t = tiles[tile];
  const b200_pu pu = pus[t & 0xFFFFF];
  const int tx = (t >> 20) & 3, ty = (t >> 22) & 3;
  const int x0 = pu.x + tx * MC_TILE, y0 = pu.y + ty * MC_TILE;
  const int tw = min(MC_TILE, pu.w - tx * MC_TILE), th = min(MC_TILE, pu.h - ty * MC_TILE);
  const bool use0 = pu.flags & B200_PU_PRED_L0, use1 = pu.flags & B200_PU_PRED_L1;
  const bool has_chroma = pic.chroma != 0;
  const int cwd = tw >> 1, chh = th >> 1;

  int yf[2] = {0, 0}, yfc[2] = {0, 0}, sh_l[2] = {0, 0}, sh_c[2] = {0, 0};
  bool missing[2] = {false, false};
#pragma unroll
  for (int l = 0; l < 2; l++) {
    if (!(l ? use1 : use0)) continue;
    const int slot = pu.ref_slot[l];
    const uint8_t* ry = (slot >= 0) ? refs.plane[slot][0] : nullptr;
    if (!ry) { missing[l] = true; continue; }
    const int mvx = pu.mv[l][0], mvy = pu.mv[l][1];
    int16_t* strip = s_strip[warp][l];
    {
      const int xf = mvx & 3;
      yf[l] = mvy & 3;
      const int hidx = (xf == 0 && yf[l] == 0) ? 4 : xf;  // full-sample: gain 64 (<< 6), no second pass
      sh_l[l] = (xf && yf[l]) ? 6 : 0;
      const int nrows = th + (yf[l] ? 7 : 0);
      mc8_hpass<MC8_CS, true>(strip, ry, pic.pitch[0], pic.w, pic.h, x0 + (mvx >> 2), y0 + (mvy >> 2), tw, nrows, yf[l] ? 3 : 0, &c_qh[hidx][0][0],
                              lane >> 2, 8, lane & 3, true);
    }
    if (has_chroma) {
      const int xfc = mvx & 7;
      yfc[l] = mvy & 7;
      const int hidx = (xfc == 0 && yfc[l] == 0) ? 8 : xfc;
      sh_c[l] = (xfc && yfc[l]) ? 6 : 0;
      const int nrows = chh + (yfc[l] ? 3 : 0);
      const int pl = (lane >> 1) & 1;
      mc8_hpass<MC8_CCS, false>(strip + 16 * MC8_CS + pl * 8 * MC8_CCS, refs.plane[slot][1 + pl], pic.pitch[1], pic.cw, pic.ch, (x0 >> 1) + (mvx >> 3),
                                (y0 >> 1) + (mvy >> 3), cwd, nrows, yfc[l] ? 1 : 0, &c_eh[hidx][0][0], lane >> 2, 8, lane & 1, true);
    }
  }
  __syncwarp();

  // weighting parameters per plane (motion.cc:493-688)
  const bool bi = use0 && use1;
  const int lu = use0 ? 0 : 1;
  const bool wgt = pu.flags & B200_PU_WEIGHTED;
  b200_weight_entry we;
  if (wgt) we = wts[pu.wt_idx];

  // ---- luma: lane -> (column, 8-row half) ----
  {
    const int c = lane & 15, h = lane >> 4;
    if (c < tw && 8 * h < th) {
      int v[2][8];
#pragma unroll
      for (int l = 0; l < 2; l++) {
        if (!(l ? use1 : use0)) continue;
        if (missing[l]) {
#pragma unroll
          for (int k = 0; k < 8; k++) v[l][k] = 1 << 13;
          continue;
        }
        const uint32_t* col = reinterpret_cast<const uint32_t*>(s_strip[warp][l] + c * MC8_CS + 8 * h);
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = col[k];
        if (yf[l]) {
          mc8_vpass<true>(w, c_qv[yf[l]], sh_l[l], v[l]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            v[l][2 * k] = (int)(int16_t)(w[k] & 0xffff);
            v[l][2 * k + 1] = (int)(int16_t)(w[k] >> 16);
          }
        }
      }
      WeightParams wp;
      wp.mode = (bi ? 1 : 0) + (wgt ? 2 : 0);
      if (wgt) {
        wp.log2wd = we.log2wd_luma;
        if (bi) { wp.w0 = we.w[0][0]; wp.o0 = we.o[0][0]; wp.w1 = we.w[1][0]; wp.o1 = we.o[1][0]; }
        else { wp.w0 = we.w[lu][0]; wp.o0 = we.o[lu][0]; wp.w1 = 0; wp.o1 = 0; }
      }
      uint8_t* dst = pic.cur[0] + (size_t)(y0 + 8 * h) * pic.pitch[0] + x0 + c;
      const int nr = min(8, th - 8 * h);
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (k < nr) dst[(size_t)k * pic.pitch[0]] = (uint8_t)weight_sample(bi ? v[0][k] : v[lu][k], bi ? v[1][k] : 0, wp, 8);
    }
  }
  // ---- chroma: lane -> (plane, column, 4-row half) ----
  if (has_chroma) {
    const int pl = lane >> 4, c = lane & 7, h = (lane >> 3) & 1;
    if (c < cwd && 4 * h < chh) {
      int v[2][4];
#pragma unroll
      for (int l = 0; l < 2; l++) {
        if (!(l ? use1 : use0)) continue;
        if (missing[l]) {
#pragma unroll
          for (int k = 0; k < 4; k++) v[l][k] = 1 << 13;
          continue;
        }
        const uint32_t* col = reinterpret_cast<const uint32_t*>(s_strip[warp][l] + 16 * MC8_CS + pl * 8 * MC8_CCS + c * MC8_CCS + 4 * h);
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) w[k] = col[k];
        if (yfc[l]) {
          mc8_vpass<false>(w, c_ev[yfc[l]], sh_c[l], v[l]);
        } else {
#pragma unroll
          for (int k = 0; k < 2; k++) {
            v[l][2 * k] = (int)(int16_t)(w[k] & 0xffff);
            v[l][2 * k + 1] = (int)(int16_t)(w[k] >> 16);
          }
        }
      }
      WeightParams wp;
      wp.mode = (bi ? 1 : 0) + (wgt ? 2 : 0);
      if (wgt) {
        wp.log2wd = we.log2wd_chroma;
        if (bi) { wp.w0 = we.w[0][1 + pl]; wp.o0 = we.o[0][1 + pl]; wp.w1 = we.w[1][1 + pl]; wp.o1 = we.o[1][1 + pl]; }
        else { wp.w0 = we.w[lu][1 + pl]; wp.o0 = we.o[lu][1 + pl]; wp.w1 = 0; wp.o1 = 0; }
      }
      uint8_t* dst = pic.cur[1 + pl] + (size_t)((y0 >> 1) + 4 * h) * pic.pitch[1 + pl] + (x0 >> 1) + c;
      const int nr = min(4, chh - 4 * h);
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (k < nr) dst[(size_t)k * pic.pitch[1 + pl]] = (uint8_t)weight_sample(bi ? v[0][k] : v[lu][k], bi ? v[1][k] : 0, wp, 8);
    }
  }
}
