// kernels_mc8.cuh — 8-bit inter prediction with packed integer dot products (the roofline-graded kernel).
//
// Replaces mc_luma / mc_chroma (motion.cc:48-282), every put_hevc_qpel/epel table entry
// (fallback-motion.cc:262-636) and the four put_*_pred functions (fallback-motion.cc:33-256) for one picture's
// worth of PUs in ONE launch.
//
// Work split: the host cuts every PU into UNITS of at most 8x16 luma samples (+ the co-located 4x8 Cb/Cr samples).
// A quarter-warp (8 lanes) owns a unit, so a warp works on four independent units at once and lanes stay busy
// for every PU size down to 8x8.  CTAs are persistent (grid-stride over units) and keep the packed tap tables in
// shared memory (each quarter-warp may need a different phase).
//
//   pass 1 (horizontal, on bytes): a lane produces 4 adjacent outputs of one row from 12 source bytes held in three
//     32-bit registers (four aligned 32-bit loads + funnel shifts); output j is dp4a(b0,T[j][0]) + dp4a(b1,T[j][1]) +
//     dp4a(b2,T[j][2]) with the 8 taps pre-shifted by j bytes into T (11 dp4a per 4 outputs instead of 32 MACs).
//     Results (|v| < 2^15, no shift at 8 bit) go to the unit's shared-memory strip stored COLUMN-major, so that two
//     vertically adjacent samples share a 32-bit word.
//   pass 2 (vertical, on int16 pairs): a lane owns one column and produces 8 rows at a time from 8 strip words with
//     dp2a (4 per even row, 5 per odd row: taps pre-packed for both parities), then >> 6 with int16 wrap
//      as one bit-field extract, the weighting (all four modes through one branch-free
//     multiply-add-shift-offset form), saturation and byte stores.
//   Integer phases use the identity tap, so there is one code path; a zero vertical phase skips pass 2 and its halo rows.
//   Units whose reference window crosses the left/right picture edge build their source bytes from clamped loads
//   (motion.cc:147-153); rows are always clamped.
//
// Tap tables are built on the host from the HEVC filter taps (engine.cu init_tables) into constant memory.
#pragma once
#include "dev_common.cuh"

#define MC8_UW 8      // unit width  (luma samples)
#define MC8_UH 16     // unit height
#define MC8_CS 24     // luma strip: int16 column stride (>= 23 rows, even)
#define MC8_CCS 12    // chroma strip column stride (>= 11 rows)
#define MC8_USTRIP (8 * MC8_CS + 2 * 4 * MC8_CCS + 8)  // one unit, one list (+8: bank skew between units)
#define MC8_WARPS 4
#define MC8_UNITS_PER_CTA (MC8_WARPS * 4)

// unit word: bits 0-19 PU index, 20-22 x offset / 8, 23-24 y offset / 16
#define MC8_UNIT(pu, ux, uy) ((uint32_t)(pu) | ((uint32_t)(ux) << 20) | ((uint32_t)(uy) << 23))

struct Mc8Tables {
  uint32_t qh[5][4][3];  // [frac 0..3, 4 = integer position with gain 64][output j][word]
  uint32_t qv[4][5];     // [frac][A,B (even rows), C,D,E (odd rows)]
  uint32_t eh[9][4][2];  // [frac 0..7, 8 = integer with gain 64][output j][word]
  uint32_t ev[8][3];     // [frac][A (even), C,D (odd)]
};
__constant__ Mc8Tables c_mc8;

// Builds the packed tap tables from the HEVC interpolation taps (host; engine.cu uploads them to c_mc8).
//   qh/eh[f][j][k]: the taps of phase f shifted right by j bytes across k words: dp4a operands for output j of a group of 4
//                   (f = 4 resp. 8: the full-sample position with gain 64); qv/ev: the taps packed for dp2a on int16 pairs,
//                   A,B for outputs on an even pair boundary, C,D,E (C,D for 4 taps) for the odd ones.
inline void mc8_build_tables(Mc8Tables& tb)
{
  static const int8_t q[4][8] = {{0, 0, 0, 1, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
  static const int8_t ep[8][4] = {{0, 1, 0, 0},     {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4},
                                  {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};
  auto pack = [](int b0, int b1, int b2, int b3) { return (uint32_t)(uint8_t)b0 | ((uint32_t)(uint8_t)b1 << 8) | ((uint32_t)(uint8_t)b2 << 16) | ((uint32_t)(uint8_t)b3 << 24); };
  auto& qh = tb.qh; auto& qv = tb.qv; auto& eh = tb.eh; auto& ev = tb.ev;
  for (int f = 0; f < 5; f++) {
    int8_t t[8];
    for (int i = 0; i < 8; i++) t[i] = (f == 4) ? (int8_t)(i == 3 ? 64 : 0) : q[f][i];
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 3; k++) {
        int b[4];
        for (int i = 0; i < 4; i++) { const int idx = 4 * k + i - j; b[i] = (idx >= 0 && idx < 8) ? t[idx] : 0; }
        qh[f][j][k] = pack(b[0], b[1], b[2], b[3]);
      }
    if (f < 4) {
      qv[f][0] = pack(t[0], t[1], t[2], t[3]); qv[f][1] = pack(t[4], t[5], t[6], t[7]);
      qv[f][2] = pack(0, t[0], t[1], t[2]);    qv[f][3] = pack(t[3], t[4], t[5], t[6]); qv[f][4] = pack(t[7], 0, 0, 0);
    }
  }
  for (int f = 0; f < 9; f++) {
    int8_t t[4];
    for (int i = 0; i < 4; i++) t[i] = (f == 8) ? (int8_t)(i == 1 ? 64 : 0) : ep[f][i];
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 2; k++) {
        int b[4];
        for (int i = 0; i < 4; i++) { const int idx = 4 * k + i - j; b[i] = (idx >= 0 && idx < 4) ? t[idx] : 0; }
        eh[f][j][k] = pack(b[0], b[1], b[2], b[3]);
      }
    if (f < 8) { ev[f][0] = pack(t[0], t[1], t[2], t[3]); ev[f][1] = pack(0, t[0], t[1], t[2]); ev[f][2] = pack(t[3], 0, 0, 0); }
  }
}

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
// (v >> shift) wrapped to int16: one signed bit-field extract of 16 bits at `shift`
__device__ __forceinline__ int shr_wrap16(int v, int shift)
{
  int d;
  asm("bfe.s32 %0, %1, %2, 16;" : "=r"(d) : "r"(v), "r"(shift));
  return d;
}
__device__ __forceinline__ int sat_u8(int v)
{
  int d;
  asm("cvt.sat.u8.s32 %0, %1;" : "=r"(d) : "r"(v));
  return d;
}

// Source bytes starting at column xb of `row` (row pointer 256-byte aligned; pitch >= width + 16).
template <int NW>  // 3 words (12 bytes) for luma, 2 words (8 bytes) for chroma
__device__ __forceinline__ void mc8_load(const uint8_t* row, int xb, int pw, bool clampx, uint32_t (&b)[3])
{
  if (!clampx) {
    const int sh = (xb & 3) * 8;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(row + (xb & ~3));
    const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
    b[0] = __funnelshift_r(w0, w1, sh);
    b[1] = __funnelshift_r(w1, w2, sh);
    if (NW == 3) b[2] = __funnelshift_r(w2, p[3], sh);
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

// Branch-free weighting: out = sat_u8(((a*w0 + b*w1 + rnd) >> shift) + off) covers fallback-motion.cc:33-256:
//   uni          (a + 32) >> 6                                   w0=1 w1=0 rnd=32            shift=6        off=0
//   bi average   (a + b + 64) >> 7                               w0=1 w1=1 rnd=64            shift=7        off=0
//   uni explicit ((a*w + 2^(wd-1)) >> wd) + o                    w0=w w1=0 rnd=2^(wd-1)      shift=wd       off=o
//   bi explicit  (a*w0 + b*w1 + ((o0+o1+1) << wd)) >> (wd+1)     w0,w1     rnd=(o0+o1+1)<<wd shift=wd+1     off=0
struct Mc8Weight {
  int w0, w1, rnd, shift, off;
};
__host__ __device__ __forceinline__ Mc8Weight mc8_weight(bool bi, bool wgt, int lu, const b200_weight_entry* __restrict__ wp_, int c)
{
  Mc8Weight r;
  if (!wgt) {
    r.w0 = 1; r.w1 = bi ? 1 : 0; r.rnd = bi ? 64 : 32; r.shift = bi ? 7 : 6; r.off = 0;
  } else {
    const int wd = c ? wp_->log2wd_chroma : wp_->log2wd_luma;
    if (bi) { r.w0 = wp_->w[0][c]; r.w1 = wp_->w[1][c]; r.rnd = (int)((unsigned)(wp_->o[0][c] + wp_->o[1][c] + 1) << wd); r.shift = wd + 1; r.off = 0; }
    else { r.w0 = wp_->w[lu][c]; r.w1 = 0; r.rnd = 1 << (wd - 1); r.shift = wd; r.off = wp_->o[lu][c]; }
  }
  return r;
}

__global__ void __launch_bounds__(MC8_WARPS * 32, 5)
k_inter_pred8(DevPic pic, RefTable refs, const b200_pu* __restrict__ pus, const b200_weight_entry* __restrict__ wts,
              const uint32_t* __restrict__ units, int n_units)
{
  __shared__ Mc8Tables s_tab;
  __shared__ __align__(16) int16_t s_strip[MC8_WARPS][4][2][MC8_USTRIP];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&c_mc8);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_tab);
    for (int i = threadIdx.x; i < (int)(sizeof(Mc8Tables) / 4); i += MC8_WARPS * 32) dst[i] = src[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane >> 3, l8 = lane & 7;
  const bool has_chroma = pic.chroma != 0;

  for (int base = blockIdx.x * MC8_UNITS_PER_CTA; base < n_units; base += gridDim.x * MC8_UNITS_PER_CTA) {
    const int u = base + warp * 4 + q;
    const bool valid = u < n_units;
    b200_pu pu;
    int x0 = 0, y0 = 0, tw = 0, th = 0;
    if (valid) {
      const uint32_t uw = units[u];
      pu = pus[uw & 0xFFFFF];
      const int ux = (uw >> 20) & 7, uy = (uw >> 23) & 3;
      x0 = pu.x + ux * MC8_UW; y0 = pu.y + uy * MC8_UH;
      tw = min(MC8_UW, pu.w - ux * MC8_UW); th = min(MC8_UH, pu.h - uy * MC8_UH);
    } else {
      pu.flags = 0;
    }
    const bool use0 = pu.flags & B200_PU_PRED_L0, use1 = pu.flags & B200_PU_PRED_L1;
    const int cwd = tw >> 1, chh = th >> 1;
    int yf[2] = {0, 0}, yfc[2] = {0, 0};
    bool missing[2] = {false, false};

    // ================= pass 1: horizontal, both lists =================
#pragma unroll
    for (int l = 0; l < 2; l++) {
      const bool used = l ? use1 : use0;
      const int slot = used ? pu.ref_slot[l] : -1;
      const uint8_t* ry = (slot >= 0) ? refs.plane[slot][0] : nullptr;
      missing[l] = used && !ry;
      if (!ry) continue;
      const int mvx = pu.mv[l][0], mvy = pu.mv[l][1];
      int16_t* strip = s_strip[warp][q][l];
      {  // ---- luma ----
        const int xf = mvx & 3;
        yf[l] = mvy & 3;
        const uint32_t* tp = &s_tab.qh[(xf == 0 && yf[l] == 0) ? 4 : xf][0][0];  // full-sample position: gain 64 (<< 6), no second pass
        uint32_t t[4][3];
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int k = 0; k < 3; k++) t[j][k] = tp[j * 3 + k];
        const int x_int = x0 + (mvx >> 2), y_int = y0 + (mvy >> 2) - (yf[l] ? 3 : 0);
        const int nrows = th + (yf[l] ? 7 : 0);
        const bool clampx = (x_int - 3 < 0) || (x_int + tw + 4 > pic.w - 1);
        // lanes of the quarter-warp: two 4-sample groups x 4 rows per iteration (8-wide unit) or 8 rows (4-wide unit)
        const bool two = tw > 4;
        const int g = two ? (l8 & 1) : 0;
        const int rstep = two ? 4 : 8;
        const int xb = x_int + 4 * g - 3;
        for (int r = two ? (l8 >> 1) : l8; r < nrows; r += rstep) {
          const int ya = clip3i(0, pic.h - 1, y_int + r);
          uint32_t b[3];
          mc8_load<3>(ry + (size_t)ya * pic.pitch[0], xb, pic.w, clampx, b);
          int16_t* o = strip + (4 * g) * MC8_CS + r;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            int s = dp4a_us(b[0], t[j][0], 0);
            s = dp4a_us(b[1], t[j][1], s);
            if (j > 0) s = dp4a_us(b[2], t[j][2], s);
            o[j * MC8_CS] = (int16_t)s;
          }
        }
      }
      if (has_chroma) {  // ---- chroma: lane -> (plane, row) ; 4:2:0: chroma mv in eighth samples = luma mv (motion.cc:196-206)
        const int xfc = mvx & 7;
        yfc[l] = mvy & 7;
        const uint32_t* tp = &s_tab.eh[(xfc == 0 && yfc[l] == 0) ? 8 : xfc][0][0];
        uint32_t t[4][2];
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int k = 0; k < 2; k++) t[j][k] = tp[j * 2 + k];
        const int x_int = (x0 >> 1) + (mvx >> 3), y_int = (y0 >> 1) + (mvy >> 3) - (yfc[l] ? 1 : 0);
        const int nrows = chh + (yfc[l] ? 3 : 0);
        const bool clampx = (x_int - 1 < 0) || (x_int + cwd + 2 > pic.cw - 1);
        const int pl = l8 & 1;
        const uint8_t* rc = refs.plane[slot][1 + pl];
        int16_t* cstrip = strip + 8 * MC8_CS + pl * 4 * MC8_CCS;
        for (int r = l8 >> 1; r < nrows; r += 4) {
          const int ya = clip3i(0, pic.ch - 1, y_int + r);
          uint32_t b[3];
          mc8_load<2>(rc + (size_t)ya * pic.pitch[1], x_int - 1, pic.cw, clampx, b);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            int s = dp4a_us(b[0], t[j][0], 0);
            if (j > 0) s = dp4a_us(b[1], t[j][1], s);
            cstrip[j * MC8_CCS + r] = (int16_t)s;
          }
        }
      }
    }
    __syncwarp();

    // ================= pass 2: vertical + weighting + store =================
    const bool bi = use0 && use1;
    const int lu = use0 ? 0 : 1;
    const bool wgt = pu.flags & B200_PU_WEIGHTED;
    const b200_weight_entry* we = wts + (wgt ? pu.wt_idx : 0);
    {  // ---- luma: lane -> column l8, two blocks of 8 rows ----
      const Mc8Weight wp = mc8_weight(bi, wgt, lu, we, 0);
      const int la = bi ? 0 : lu;  // list providing operand a
      uint32_t tv[2][5];
#pragma unroll
      for (int l = 0; l < 2; l++)
#pragma unroll
        for (int k = 0; k < 5; k++) tv[l][k] = s_tab.qv[yf[l]][k];
      const int shl[2] = {((pu.mv[0][0] & 3) && yf[0]) ? 6 : 0, ((pu.mv[1][0] & 3) && yf[1]) ? 6 : 0};
      const bool act = valid && (use0 || use1) && l8 < tw;
#pragma unroll
      for (int hb = 0; hb < 2; hb++) {
        if (!(act && 8 * hb < th)) continue;
        int v[2][8];
#pragma unroll
        for (int l = 0; l < 2; l++) {
          const bool need = (l == la) || (bi && l == 1);
          if (!need || missing[l]) {  // never-written slot: mid-grey intermediate (decctx.cc:1538-1583 conceals with grey)
#pragma unroll
            for (int k = 0; k < 8; k++) v[l][k] = need ? (1 << 13) : 0;
            continue;
          }
          const uint32_t* col = reinterpret_cast<const uint32_t*>(s_strip[warp][q][l] + l8 * MC8_CS + 8 * hb);
          uint32_t w[8];
#pragma unroll
          for (int k = 0; k < 8; k++) w[k] = col[k];
          if (yf[l]) {
            const uint32_t A = tv[l][0], B = tv[l][1], C = tv[l][2], D = tv[l][3], E = tv[l][4];
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
              v[l][2 * m] = shr_wrap16(e, shl[l]);
              v[l][2 * m + 1] = shr_wrap16(o, shl[l]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
              v[l][2 * k] = (int)(int16_t)(w[k] & 0xffff);
              v[l][2 * k + 1] = (int)w[k] >> 16;
            }
          }
        }
        uint8_t* dst = pic.cur[0] + (size_t)(y0 + 8 * hb) * pic.pitch[0] + x0 + l8;
        const int nr = th - 8 * hb;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int a = (la == 0) ? v[0][k] : v[1][k];
          const int out = sat_u8(((a * wp.w0 + v[1][k] * wp.w1 + wp.rnd) >> wp.shift) + wp.off);
          if (k < nr) dst[0] = (uint8_t)out;
          dst += pic.pitch[0];
        }
      }
    }
    if (has_chroma) {  // ---- chroma: lane -> (plane l8 >> 2, column l8 & 3), two blocks of 4 rows ----
      const int pl = l8 >> 2, c = l8 & 3;
      const Mc8Weight wp = mc8_weight(bi, wgt, lu, we, 1 + pl);
      const int la = bi ? 0 : lu;
      uint32_t tv[2][3];
#pragma unroll
      for (int l = 0; l < 2; l++)
#pragma unroll
        for (int k = 0; k < 3; k++) tv[l][k] = s_tab.ev[yfc[l]][k];
      const int shl[2] = {((pu.mv[0][0] & 7) && yfc[0]) ? 6 : 0, ((pu.mv[1][0] & 7) && yfc[1]) ? 6 : 0};
      const bool act = valid && (use0 || use1) && c < cwd;
#pragma unroll
      for (int hb = 0; hb < 2; hb++) {
        if (!(act && 4 * hb < chh)) continue;
        int v[2][4];
#pragma unroll
        for (int l = 0; l < 2; l++) {
          const bool need = (l == la) || (bi && l == 1);
          if (!need || missing[l]) {
#pragma unroll
            for (int k = 0; k < 4; k++) v[l][k] = need ? (1 << 13) : 0;
            continue;
          }
          const uint32_t* col = reinterpret_cast<const uint32_t*>(s_strip[warp][q][l] + 8 * MC8_CS + pl * 4 * MC8_CCS + c * MC8_CCS + 4 * hb);
          uint32_t w[4];
#pragma unroll
          for (int k = 0; k < 4; k++) w[k] = col[k];
          if (yfc[l]) {
            const uint32_t A = tv[l][0], C = tv[l][1], D = tv[l][2];
#pragma unroll
            for (int m = 0; m < 2; m++) {
              int e = dp2a_lo_ss(w[m], A, 0);
              e = dp2a_hi_ss(w[m + 1], A, e);
              int o = dp2a_lo_ss(w[m], C, 0);
              o = dp2a_hi_ss(w[m + 1], C, o);
              o = dp2a_lo_ss(w[m + 2], D, o);
              v[l][2 * m] = shr_wrap16(e, shl[l]);
              v[l][2 * m + 1] = shr_wrap16(o, shl[l]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 2; k++) {
              v[l][2 * k] = (int)(int16_t)(w[k] & 0xffff);
              v[l][2 * k + 1] = (int)w[k] >> 16;
            }
          }
        }
        uint8_t* dst = pic.cur[1 + pl] + (size_t)((y0 >> 1) + 4 * hb) * pic.pitch[1 + pl] + (x0 >> 1) + c;
        const int nr = chh - 4 * hb;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int a = (la == 0) ? v[0][k] : v[1][k];
          const int out = sat_u8(((a * wp.w0 + v[1][k] * wp.w1 + wp.rnd) >> wp.shift) + wp.off);
          if (k < nr) dst[0] = (uint8_t)out;
          dst += pic.pitch[1 + pl];
        }
      }
    }
    __syncwarp();  // the strips are reused by the next group of units
  }
}
