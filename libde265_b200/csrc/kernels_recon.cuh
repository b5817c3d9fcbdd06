// kernels_recon.cuh — intra prediction + residual (dequant, inverse DCT/DST, transform-skip, bypass,
// PCM) for one picture in ONE launch.
//
// One CTA per CTB.  The CTB's samples (the inter prediction written by k_inter_pred, or nothing yet
// for intra CUs) plus the neighbouring reconstructed row/column are staged in shared memory, every TU
// of the CTB is reconstructed there, and the finished CTB is written back with 16-byte row stores.
//   phase A  all TUs that are not intra (inter residual, PCM): independent -> spread over all warps
//   phase B  intra TUs: serially dependent inside a colour plane, independent between
//            planes -> warp c walks plane c's TUs in decode order
// CTBs that contain intra TUs wait (acquire-spin on a per-CTB flag) until their left, top-left, top
// and top-right neighbours are finished — the WPP dependency (slice.cc:4789-4795); CTBs are handed
// out in raster order through an atomic ticket so a waiting CTA only ever waits for CTBs that are
// already running or done.
//
// Replaces decode_TU (slice.cc:3460), decode_intra_prediction (intrapred.cc:277-345) incl. border
// fetch/substitution/smoothing (intrapred.h:185-258,529-674), scale_coefficients (transform.cc:361-642)
// and the transform_* / add_residual / dequant entries of the DSP table (fallback-dct.cc).
#pragma once
#include "dev_common.cuh"

#define RC_WARPS 8
#define RC_THREADS (RC_WARPS * 32)
#define RC_LSTRIDE 128  // luma tile row stride (samples): x = -16 .. 111
#define RC_CSTRIDE 64   // chroma tile row stride: x = -16 .. 47
#define RC_XOFF 16
#define RC_LROWS 65
#define RC_CROWS 33
#define RC_GSTRIDE 34   // int16 row stride of the first-stage buffer (conflict-free for 32 lanes)

struct ReconArgs {
  const b200_tu* tus;            // grouped by CTB, decode order inside
  const uint32_t* ctb_tu_start;  // [n_ctb + 1]; bit 31 of entry i+1 is NOT used; see ctb_has_intra
  const uint8_t* ctb_has_intra;  // [n_ctb]
  const b200_coeff* coeffs;
  const uint8_t* scaling;        // B200_SCALING_FACTOR_BYTES or null
  unsigned int* ticket;          // zeroed before launch
  unsigned int* ctb_done;        // [n_ctb], zeroed before launch
};

template <typename P>
struct ReconSmem {
  P luma[RC_LROWS * RC_LSTRIDE];
  P chroma[2][RC_CROWS * RC_CSTRIDE];
  int16_t coef[RC_WARPS][32 * 32];
  int16_t g[RC_WARPS][32 * RC_GSTRIDE];
  P border[RC_WARPS][2][4 * 32 + 4];  // [0] gathered/substituted, [1] filtered / angular ref
  int8_t dct[32][32];
  int ctb;
};

__device__ __forceinline__ int warp_max(int v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// -------------------------------------------------------------------------------------------------
// residual of one TU, executed by one warp; `dst` points at the TU's top-left sample in the smem tile
// -------------------------------------------------------------------------------------------------
template <typename P>
__device__ void tu_residual(const b200_tu& tu, const b200_coeff* __restrict__ coeffs, const uint8_t* __restrict__ scaling, P* dst,
                            int dstride, int bd, int16_t* coef, int16_t* g, const int8_t (*dct)[32], int lane)
{
  const int log2 = tu.log2_size, nT = 1 << log2, n = tu.n_coeff;
  const b200_coeff* co = coeffs + tu.coeff_off;
  const int flags = tu.flags;
  for (int i = lane; i < nT * nT; i += 32) coef[i] = 0;
  __syncwarp();
  // ---- dequant + scatter (transform.cc:452-525) ----
  int max_row = 0, max_col = 0;
  {
    const bool bypass = flags & B200_TU_BYPASS;
    const bool rotate = (flags & B200_TU_ROTATE) && (flags & (B200_TU_BYPASS | B200_TU_TSKIP));
    const uint8_t* scl = nullptr;
    if ((flags & B200_TU_SCALING_LIST) && scaling) {
      int m = (nT == 32) ? 0 : tu.cidx;
      if (flags & B200_TU_INTER_MATRIX) m += (nT < 32) ? 3 : 1;
      const int base = (nT == 4) ? 0 : (nT == 8) ? 6 * 16 : (nT == 16) ? 6 * 16 + 6 * 64 : 6 * 16 + 6 * 64 + 6 * 256;
      scl = scaling + base + m * nT * nT;
    }
    int bd_shift = bd + log2 - 5;
    if (!scl) bd_shift -= 4;
    const int qp = tu.qp;
    const int ls = (qp % 6 == 0) ? 40 : (qp % 6 == 1) ? 45 : (qp % 6 == 2) ? 51 : (qp % 6 == 3) ? 57 : (qp % 6 == 4) ? 64 : 72;
    for (int i = lane; i < n; i += 32) {
      const b200_coeff c = co[i];
      int v;
      if (bypass) {
        v = c.level;
      } else {
        const long long fact = (long long)((scl ? scl[c.pos] : 1) * ls) << (qp / 6);
        long long q = ((long long)c.level * fact + (1ll << (bd_shift - 1))) >> bd_shift;
        v = (int)max(-32768ll, min(32767ll, q));
      }
      const int pos = rotate ? (nT * nT - 1 - c.pos) : c.pos;
      coef[pos] = (int16_t)v;
      max_row = max(max_row, pos >> log2);
      max_col = max(max_col, pos & (nT - 1));
    }
    max_row = warp_max(max_row);
    max_col = warp_max(max_col);
  }
  __syncwarp();

  if (flags & (B200_TU_BYPASS | B200_TU_TSKIP)) {
    // transform.cc:408-448 / :548-596 with fallback-dct.cc:81-91,161-225
    const bool ts = !(flags & B200_TU_BYPASS);
    const int bd_shift = 20 - bd, ts_shift = 5 + log2, rnd = 1 << (bd_shift - 1);
    if (flags & (B200_TU_RDPCM_H | B200_TU_RDPCM_V)) {
      const bool vert = flags & B200_TU_RDPCM_V;
      if (lane < nT) {
        int sum = 0;
        for (int k = 0; k < nT; k++) {
          const int x = vert ? lane : k, y = vert ? k : lane;
          int c = coef[x + y * nT];
          if (ts) c = ((int)((unsigned)c << ts_shift) + rnd) >> bd_shift;
          sum += c;
          dst[x + y * dstride] = (P)clip_bd((int)dst[x + y * dstride] + sum, bd);
        }
      }
    } else {
      for (int i = lane; i < nT * nT; i += 32) {
        int c = coef[i];
        if (ts) c = ((int)((unsigned)c << ts_shift) + rnd) >> bd_shift;
        const int x = i & (nT - 1), y = i >> log2;
        dst[x + y * dstride] = (P)clip_bd((int)dst[x + y * dstride] + c, bd);
      }
    }
    __syncwarp();
    return;
  }

  const int post_shift = 20 - bd, rnd2 = 1 << (post_shift - 1);
  if (flags & B200_TU_DST) {
    // fallback-dct.cc:269-407 (mat_8_357 :260-265)
    const int m[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};
    if (lane < 16) {
      const int c = lane & 3, i = lane >> 2;
      int sum = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) sum += m[j][i] * coef[c + j * 4];
      g[i * RC_GSTRIDE + c] = (int16_t)clip3i(-32768, 32767, (sum + 64) >> 7);
    }
    __syncwarp();
    if (lane < 16) {
      const int i = lane & 3, y = lane >> 2;
      int sum = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) sum += m[j][i] * g[y * RC_GSTRIDE + j];
      const int out = clip3i(-32768, 32767, (sum + rnd2) >> post_shift);
      dst[i + y * dstride] = (P)clip_bd((int)dst[i + y * dstride] + out, bd);
    }
    __syncwarp();
    return;
  }

  // ---- inverse DCT (fallback-dct.cc:550-691); zero rows/columns beyond the last coefficient are skipped ----
  const int fact = 32 >> log2;
  for (int o = lane; o < nT * nT; o += 32) {  // pass 1: columns.  o -> (i = output row, c = column)
    const int c = o & (nT - 1), i = o >> log2;
    int sum = 0;
    if (c <= max_col)
      for (int j = 0; j <= max_row; j++) sum += (int)dct[fact * j][i] * (int)coef[c + j * nT];
    g[i * RC_GSTRIDE + c] = (int16_t)clip3i(-32768, 32767, (sum + 64) >> 7);
  }
  __syncwarp();
  for (int o = lane; o < nT * nT; o += 32) {  // pass 2: rows.  lanes along x: g broadcast, dct row contiguous, dst contiguous
    const int i = o & (nT - 1), y = o >> log2;
    int sum = 0;
    for (int j = 0; j <= max_col; j++) sum += (int)dct[fact * j][i] * (int)g[y * RC_GSTRIDE + j];
    const int out = (sum + rnd2) >> post_shift;
    dst[i + y * dstride] = (P)clip_bd((int)dst[i + y * dstride] + out, bd);
  }
  __syncwarp();
}

// -------------------------------------------------------------------------------------------------
// intra prediction of one TU by one warp.  tile(x,y) addresses the plane's smem tile relative to the
// TU's top-left sample.
// -------------------------------------------------------------------------------------------------
__constant__ int8_t k_intra_angle[35] = {0,   0,   32,  26,  21,  17, 13, 9,  5,  2,  0,  -2, -5, -9, -13, -17, -21, -26,
                                         -32, -26, -21, -17, -13, -9, -5, -2, 0,  2,  5,  9,  13, 17, 21,  26,  32};
__constant__ int16_t k_inv_angle[15] = {-4096, -1638, -910, -630, -482, -390, -315, -256, -315, -390, -482, -630, -910, -1638, -4096};

template <typename P>
__device__ void tu_intra(const b200_tu& tu, P* dst, int dstride, int bd, int bd_luma, uint32_t pic_flags, bool filter_plane, P* b0mem, P* b1mem,
                         int lane)
{
  const int log2 = tu.log2_size, nT = 1 << log2, mode = tu.intra_mode, cidx = tu.cidx;
  const uint64_t avail = tu.avail;
  P* b0 = b0mem + 2 * 32 + 2;  // centre element; valid [-2nT, 2nT]
  P* b1 = b1mem + 2 * 32 + 2;
  const int total = 4 * nT + 1;
  // ---- gather + substitution (intrapred.h:529-674); scan index s: 0 -> border[-2nT], 2nT -> border[0], 4nT -> border[2nT]
  {
    int carry = -1;  // value of the last sample of the previous chunk after substitution
    // first available sample in scan order (firstValue)
    int first_val = 1 << (bd - 1);
    bool any = false;
    for (int base = 0; base < total; base += 32) {
      const int s = base + lane, i = s - 2 * nT;
      bool av = false;
      int v = 0;
      if (s < total) {
        if (i < 0) { const int r = -i - 1; av = (avail >> (r >> 2)) & 1; if (av) v = dst[-1 + r * dstride]; }
        else if (i == 0) { av = (avail >> B200_AVAIL_CORNER_BIT) & 1; if (av) v = dst[-1 - dstride]; }
        else { const int c = i - 1; av = (avail >> (B200_AVAIL_TOP_BIT0 + (c >> 2))) & 1; if (av) v = dst[c - dstride]; }
      }
      const unsigned m = __ballot_sync(0xffffffffu, av);
      if (!any && m) { first_val = __shfl_sync(0xffffffffu, v, __ffs(m) - 1); any = true; }
    }
    if (!any) {
      for (int s = lane; s < total; s += 32) b0[s - 2 * nT] = (P)(1 << (bd - 1));
    } else {
      carry = first_val;
      for (int base = 0; base < total; base += 32) {
        const int s = base + lane, i = s - 2 * nT;
        bool av = false;
        int v = 0;
        if (s < total) {
          if (i < 0) { const int r = -i - 1; av = (avail >> (r >> 2)) & 1; if (av) v = dst[-1 + r * dstride]; }
          else if (i == 0) { av = (avail >> B200_AVAIL_CORNER_BIT) & 1; if (av) v = dst[-1 - dstride]; }
          else { const int c = i - 1; av = (avail >> (B200_AVAIL_TOP_BIT0 + (c >> 2))) & 1; if (av) v = dst[c - dstride]; }
        }
        const unsigned m = __ballot_sync(0xffffffffu, av);
        const unsigned below = m & ((2u << lane) - 1u);  // available lanes <= this one
        const int src = below ? 31 - __clz(below) : 0;
        const int sv = __shfl_sync(0xffffffffu, v, src);
        const int outv = below ? sv : carry;
        if (s < total) b0[i] = (P)outv;
        carry = __shfl_sync(0xffffffffu, outv, 31);
      }
    }
  }
  __syncwarp();
  // ---- smoothing (intrapred.h:185-258) ----
  const P* bsrc = b0;
  if (filter_plane && mode != 1 && nT != 4) {
    const int d = min(abs(mode - 26), abs(mode - 10));
    const bool filt = (nT == 8) ? (d > 7) : (nT == 16) ? (d > 1) : (d > 0);
    if (filt) {
      const bool strong = (pic_flags & B200_PIC_STRONG_INTRA_SMOOTHING) && cidx == 0 && nT == 32 &&
                          abs((int)b0[0] + (int)b0[64] - 2 * (int)b0[32]) < (1 << (bd_luma - 5)) &&
                          abs((int)b0[0] + (int)b0[-64] - 2 * (int)b0[-32]) < (1 << (bd_luma - 5));
      for (int s = lane; s < total; s += 32) {
        const int i = s - 2 * nT;
        int v;
        if (i == -2 * nT || i == 2 * nT) v = b0[i];
        else if (strong) {
          if (i == 0) v = b0[0];
          else if (i < 0) v = (int)b0[0] + (((-i) * ((int)b0[-64] - (int)b0[0]) + 32) >> 6);
          else v = (int)b0[0] + ((i * ((int)b0[64] - (int)b0[0]) + 32) >> 6);
        } else v = ((int)b0[i + 1] + 2 * (int)b0[i] + (int)b0[i - 1] + 2) >> 2;
        b1[i] = (P)v;
      }
      __syncwarp();
      bsrc = b1;
    }
  }
  P* bfree = (bsrc == b0) ? b1 : b0;  // scratch for the angular reference array
  // ---- prediction (intrapred.h:261-433) ----
  if (mode == 0) {
    for (int o = lane; o < nT * nT; o += 32) {
      const int x = o & (nT - 1), y = o >> log2;
      dst[x + y * dstride] = (P)(((nT - 1 - x) * (int)bsrc[-1 - y] + (x + 1) * (int)bsrc[1 + nT] + (nT - 1 - y) * (int)bsrc[1 + x] +
                                  (y + 1) * (int)bsrc[-1 - nT] + nT) >> (log2 + 1));
    }
  } else if (mode == 1) {
    int part = 0;
    for (int i = lane; i < nT; i += 32) part += (int)bsrc[i + 1] + (int)bsrc[-i - 1];
    const int dc = (warp_sum(part) + nT) >> (log2 + 1);
    const bool edge = (cidx == 0 && nT < 32);
    for (int o = lane; o < nT * nT; o += 32) {
      const int x = o & (nT - 1), y = o >> log2;
      int v = dc;
      if (edge) {
        if (x == 0 && y == 0) v = ((int)bsrc[-1] + 2 * dc + (int)bsrc[1] + 2) >> 2;
        else if (y == 0) v = ((int)bsrc[x + 1] + 3 * dc + 2) >> 2;
        else if (x == 0) v = ((int)bsrc[-y - 1] + 3 * dc + 2) >> 2;
      }
      dst[x + y * dstride] = (P)v;
    }
  } else {
    const int angle = k_intra_angle[mode];
    const bool vert = mode >= 18;
    const int sgn = vert ? 1 : -1;
    P* ref = bfree;  // ref[x] valid on [-nT, 2nT]
    const int last = (nT * angle) >> 5;
    for (int s = lane; s <= 3 * nT; s += 32) {
      const int x = s - nT;
      int v = 0;
      bool w = false;
      if (x >= 0 && x <= nT) { v = bsrc[sgn * x]; w = true; }
      else if (x > nT) { if (angle >= 0) { v = bsrc[sgn * x]; w = true; } }
      else if (angle < 0 && last < -1 && x >= last) { v = bsrc[-sgn * ((x * (int)k_inv_angle[mode - 11] + 128) >> 8)]; w = true; }
      if (w) ref[x] = (P)v;
    }
    __syncwarp();
    const bool bfilt = (cidx == 0 && nT < 32 && !(tu.flags & B200_TU_NO_BOUNDARY_FILTER) && (mode == 26 || mode == 10));
    for (int o = lane; o < nT * nT; o += 32) {
      const int x = o & (nT - 1), y = o >> log2;
      const int a = vert ? y : x, b = vert ? x : y;
      const int idx = ((a + 1) * angle) >> 5, fact = ((a + 1) * angle) & 31;
      int v = fact ? ((32 - fact) * (int)ref[b + idx + 1] + fact * (int)ref[b + idx + 2] + 16) >> 5 : (int)ref[b + idx + 1];
      if (bfilt) {
        if (mode == 26 && x == 0) v = clip_bd((int)bsrc[1] + (((int)bsrc[-1 - y] - (int)bsrc[0]) >> 1), bd);
        if (mode == 10 && y == 0) v = clip_bd((int)bsrc[-1] + (((int)bsrc[1 + x] - (int)bsrc[0]) >> 1), bd);
      }
      dst[x + y * dstride] = (P)v;
    }
  }
  __syncwarp();
}

// -------------------------------------------------------------------------------------------------
template <typename P>
__global__ void __launch_bounds__(RC_THREADS) k_recon(DevPic pic, ReconArgs args)
{
  extern __shared__ __align__(16) uint8_t smem_raw[];
  ReconSmem<P>& sm = *reinterpret_cast<ReconSmem<P>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) sm.ctb = (int)atomicAdd(args.ticket, 1u);
  for (int i = tid; i < 32 * 32; i += RC_THREADS) sm.dct[i >> 5][i & 31] = c_dct[i >> 5][i & 31];
  __syncthreads();
  const int ctb = sm.ctb;
  const int n_ctb = pic.wctb * pic.hctb;
  if (ctb >= n_ctb) return;
  const int cx = ctb % pic.wctb, cy = ctb / pic.wctb;
  const int S = 1 << pic.log2ctb;
  const int xC = cx << pic.log2ctb, yC = cy << pic.log2ctb;
  const int bw = min(S, pic.w - xC), bh = min(S, pic.h - yC);  // CTB part inside the picture
  const uint32_t t0 = args.ctb_tu_start[ctb], t1 = args.ctb_tu_start[ctb + 1];
  const bool has_intra = args.ctb_has_intra[ctb];
  const int nplanes = pic.chroma ? 3 : 1;

  if (t1 == t0) {  // nothing to do for this CTB (pure skip CUs): prediction already final
    if (tid == 0) { __threadfence(); atomicExch(&args.ctb_done[ctb], 1u); }
    return;
  }

  // ---- wait for the neighbours an intra CTB may read (left, top-left, top, top-right) ----
  if (has_intra) {
    if (tid < 4) {
      const int nx = (tid == 0) ? cx - 1 : (tid == 1) ? cx - 1 : (tid == 2) ? cx : cx + 1;
      const int ny = (tid == 0) ? cy : cy - 1;
      if (nx >= 0 && ny >= 0 && nx < pic.wctb) {
        const volatile unsigned int* f = args.ctb_done + (nx + ny * pic.wctb);
        while (*f == 0u) __nanosleep(20);
      }
      __threadfence();
    }
    __syncthreads();
  }

  // ---- stage the CTB (+ top row / left column when intra) into shared memory ----
  for (int c = 0; c < nplanes; c++) {
    const int sh = c ? 1 : 0;
    const int w = bw >> sh, h = bh >> sh, x0 = xC >> sh, y0 = yC >> sh;
    const int pw = c ? pic.cw : pic.w;
    P* tile = c ? sm.chroma[c - 1] : sm.luma;
    const int ts = c ? RC_CSTRIDE : RC_LSTRIDE;
    const uint8_t* src = pic.cur[c];
    const int pitch = pic.pitch[c];
    constexpr int VEC = 16 / sizeof(P);
    const int vpr = (w + VEC - 1) / VEC;  // 16-byte vectors per row (surface rows are padded)
    for (int i = tid; i < vpr * h; i += RC_THREADS) {
      const int y = i / vpr, v = i % vpr;
      // .cg loads: neighbour CTBs are written by other SMs during this launch, L1 must not serve them
      const uint4 d = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)(y0 + y) * pitch + (size_t)(x0 + v * VEC) * sizeof(P)));
      *reinterpret_cast<uint4*>(&tile[(y + 1) * ts + RC_XOFF + v * VEC]) = d;
    }
    if (has_intra) {
      const int tr = min(S >> sh, 32 >> sh);  // top-right reach = largest TU of the plane
      if (y0 > 0)
        for (int x = tid - 1; x < w + tr; x += RC_THREADS)
          if (x0 + x >= 0 && x0 + x < pw) tile[RC_XOFF + x] = __ldcg(row_ptr<P>(src, pitch, y0 - 1) + x0 + x);
      if (x0 > 0)
        for (int y = tid; y < h; y += RC_THREADS) tile[(y + 1) * ts + RC_XOFF - 1] = __ldcg(row_ptr<P>(src, pitch, y0 + y) + x0 - 1);
    }
  }
  __syncthreads();

  // ---- phase A: non-intra TUs, one warp each ----
  for (uint32_t t = t0 + warp; t < t1; t += RC_WARPS) {
    const b200_tu tu = args.tus[t];
    if (tu.flags & B200_TU_INTRA) continue;
    const int c = tu.cidx, sh = c ? 1 : 0;
    P* tile = c ? sm.chroma[c - 1] : sm.luma;
    const int ts = c ? RC_CSTRIDE : RC_LSTRIDE;
    P* dst = tile + (tu.y - (yC >> sh) + 1) * ts + RC_XOFF + (tu.x - (xC >> sh));
    if (tu.flags & B200_TU_PCM) {
      const int nT = 1 << tu.log2_size;
      for (int i = lane; i < tu.n_coeff; i += 32) {
        const b200_coeff co = args.coeffs[tu.coeff_off + i];
        dst[(co.pos & (nT - 1)) + (co.pos >> tu.log2_size) * ts] = (P)(uint16_t)co.level;
      }
      __syncwarp();
    } else if (tu.flags & B200_TU_CBF) {
      tu_residual<P>(tu, args.coeffs, args.scaling, dst, ts, c ? pic.bd_c : pic.bd_y, sm.coef[warp], sm.g[warp], sm.dct, lane);
    }
  }
  __syncthreads();

  // ---- phase B: intra TUs, warp c owns colour plane c ----
  if (has_intra && warp < nplanes) {
    const int c = warp, sh = c ? 1 : 0;
    P* tile = c ? sm.chroma[c - 1] : sm.luma;
    const int ts = c ? RC_CSTRIDE : RC_LSTRIDE;
    const int bd = c ? pic.bd_c : pic.bd_y;
    const bool filter_plane = !(pic.flags & B200_PIC_INTRA_SMOOTHING_OFF) && (c == 0 || pic.chroma == 3);
    for (uint32_t t = t0; t < t1; t++) {
      const b200_tu tu = args.tus[t];
      if (!(tu.flags & B200_TU_INTRA) || tu.cidx != c) continue;
      P* dst = tile + (tu.y - (yC >> sh) + 1) * ts + RC_XOFF + (tu.x - (xC >> sh));
      tu_intra<P>(tu, dst, ts, bd, pic.bd_y, pic.flags, filter_plane, sm.border[warp][0], sm.border[warp][1], lane);
      if (tu.flags & B200_TU_CBF) tu_residual<P>(tu, args.coeffs, args.scaling, dst, ts, bd, sm.coef[warp], sm.g[warp], sm.dct, lane);
    }
  }
  __syncthreads();

  // ---- write the CTB back (16-byte row stores) and publish it ----
  for (int c = 0; c < nplanes; c++) {
    const int sh = c ? 1 : 0;
    const int w = bw >> sh, h = bh >> sh, x0 = xC >> sh, y0 = yC >> sh;
    const P* tile = c ? sm.chroma[c - 1] : sm.luma;
    const int ts = c ? RC_CSTRIDE : RC_LSTRIDE;
    uint8_t* dstp = pic.cur[c];
    const int pitch = pic.pitch[c];
    constexpr int VEC = 16 / sizeof(P);
    const int vpr = (w + VEC - 1) / VEC;
    for (int i = tid; i < vpr * h; i += RC_THREADS) {
      const int y = i / vpr, v = i % vpr;
      *reinterpret_cast<uint4*>(dstp + (size_t)(y0 + y) * pitch + (size_t)(x0 + v * VEC) * sizeof(P)) =
          *reinterpret_cast<const uint4*>(&tile[(y + 1) * ts + RC_XOFF + v * VEC]);
    }
  }
  __syncthreads();
  if (tid == 0) { __threadfence(); atomicExch(&args.ctb_done[ctb], 1u); }
}
