// dsp_table.cuh — boundary B1 (include/b200hevc_dsp.h): the entries of libde265's DSP table as batched,
// device-executed calls.  One CTA per command; the device code reuses the MC pass / weighting helpers of the
// picture kernels (kernels_mc.cuh) and restates the small per-block transforms, intra predictors and deblocking
// filters on explicit block buffers.  Included by engine.cu (single translation unit: constant tables).
#pragma once
#include "b200hevc_dsp.h"
#include "dev_common.cuh"
#include "kernels_mc.cuh"

struct DspDevCmd {
  int op, bd, w, h;
  int a[8];
  uint32_t in0, in1, io;       // byte offsets into the arena
  int p_in0, p_in1, p_io;      // row pitches in bytes
};

// ---- device side --------------------------------------------------------------------------------
template <typename P>
__device__ void dsp_mc(const DspDevCmd& c, uint8_t* arena, int16_t* strip_all, bool luma)
{
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xf = c.a[0], yf = c.a[1];
  const int hl = luma ? 3 : 1;                                  // halo before the block
  const int ox = xf ? hl : 0, oy = yf ? hl : 0;                 // the staged window starts at (-ox, -oy)
  const int pw = c.w + (xf ? (luma ? 7 : 3) : 0), ph = c.h + (yf ? (luma ? 7 : 3) : 0);
  int16_t* strip = strip_all + warp * MC_STRIP;
  int16_t* out = reinterpret_cast<int16_t*>(arena + c.io);
  const int tiles_y = (c.h + MC_TILE - 1) / MC_TILE;
  // luma tiles are MC_TILE wide (the strip row stride of the luma helpers); the chroma helpers use MC_TILE / 2
  const int tile_w = luma ? MC_TILE : MC_TILE / 2;
  const int ntx = (c.w + tile_w - 1) / tile_w;
  for (int t = warp; t < ntx * tiles_y; t += 4) {
    const int tx = (t % ntx) * tile_w, ty = (t / ntx) * MC_TILE;
    const int tw = min(tile_w, c.w - tx), th = min(MC_TILE, c.h - ty);
    if (luma) mc_hpass<P, 8>(strip, arena + c.in0, c.p_in0, pw, ph, ox + tx, oy + ty, xf, yf, tw, th, c.bd, lane, k_qpel[xf]);
    else mc_hpass<P, 4>(strip, arena + c.in0, c.p_in0, pw, ph, ox + tx, oy + ty, xf, yf, tw, th, c.bd, lane, k_epel[xf]);
    __syncwarp();
    for (int o = lane; o < tw * th; o += 32) {
      const int x = o % tw, y = o / tw;
      const int v = luma ? mc_vsample<8>(strip, y, x, xf, yf, c.bd, k_qpel[yf]) : mc_vsample<4>(strip, y, x, xf, yf, c.bd, k_epel[yf]);
      out[(size_t)(ty + y) * (c.p_io / 2) + tx + x] = (int16_t)v;
    }
    __syncwarp();
  }
}

template <typename P>
__device__ void dsp_pred(const DspDevCmd& c, uint8_t* arena)
{
  WeightParams wp;
  wp.mode = c.op - B200_DSP_PRED_UNI;  // 0 uni, 1 avg, 2 weighted, 3 weighted bi
  wp.w0 = c.a[0]; wp.o0 = c.a[1]; wp.w1 = c.a[2]; wp.o1 = c.a[3];
  wp.log2wd = (c.op == B200_DSP_PRED_WEIGHTED) ? c.a[2] : c.a[4];
  const int16_t* s1 = reinterpret_cast<const int16_t*>(arena + c.in0);
  const int16_t* s2 = reinterpret_cast<const int16_t*>(arena + c.in1);
  const bool two = (c.op == B200_DSP_PRED_AVG || c.op == B200_DSP_PRED_WEIGHTED_BI);
  for (int o = threadIdx.x; o < c.w * c.h; o += blockDim.x) {
    const int x = o % c.w, y = o / c.w;
    const int a = s1[(size_t)y * (c.p_in0 / 2) + x], b = two ? s2[(size_t)y * (c.p_in1 / 2) + x] : 0;
    row_ptr<P>(arena + c.io, c.p_io, y)[x] = (P)weight_sample(a, b, wp, c.bd);
  }
}

template <typename P>
__device__ void dsp_transform(const DspDevCmd& c, uint8_t* arena, int16_t* g)
{
  // fallback-dct.cc:550-691 (DCT) / :269-407 (DST)
  const bool dst7 = c.op == B200_DSP_DST_ADD;
  const int log2 = dst7 ? 2 : c.a[0], nT = 1 << log2, fact = 32 >> log2;
  const int16_t* coef = reinterpret_cast<const int16_t*>(arena + c.in0);
  const int post_shift = 20 - c.bd, rnd2 = 1 << (post_shift - 1);
  auto m = [&](int j, int i) -> int {
    if (!dst7) return c_dct[fact * j][i];
    const int t[16] = {29, 55, 74, 84, 74, 74, 0, -74, 84, -29, -74, 55, 55, -84, 74, -29};
    int r = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) r = (k == j * 4 + i) ? t[k] : r;
    return r;
  };
  for (int o = threadIdx.x; o < nT * nT; o += blockDim.x) {
    const int cc = o & (nT - 1), i = o >> log2;
    int sum = 0;
    for (int j = 0; j < nT; j++) sum += m(j, i) * (int)coef[cc + j * nT];
    g[i * nT + cc] = (int16_t)clip3i(-32768, 32767, (sum + 64) >> 7);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < nT * nT; o += blockDim.x) {
    const int i = o & (nT - 1), y = o >> log2;
    int sum = 0;
    for (int j = 0; j < nT; j++) sum += m(j, i) * (int)g[y * nT + j];
    int r = (sum + rnd2) >> post_shift;
    if (dst7) r = clip3i(-32768, 32767, r);
    P* p = row_ptr<P>(arena + c.io, c.p_io, y) + i;
    *p = (P)clip_bd((int)*p + r, c.bd);
  }
}

template <typename P>
__device__ void dsp_intra(const DspDevCmd& c, uint8_t* arena, int* ref_s)
{
  const int nT = c.a[0], cidx = c.a[1], log2 = 31 - __clz(nT);
  const P* b = reinterpret_cast<const P*>(arena + c.in0) + 2 * nT;  // centre element
  auto put = [&](int x, int y, int v) { row_ptr<P>(arena + c.io, c.p_io, y)[x] = (P)v; };
  if (c.op == B200_DSP_INTRA_PLANAR) {  // intrapred.h:261-285
    for (int o = threadIdx.x; o < nT * nT; o += blockDim.x) {
      const int x = o & (nT - 1), y = o >> log2;
      put(x, y, ((nT - 1 - x) * (int)b[-1 - y] + (x + 1) * (int)b[1 + nT] + (nT - 1 - y) * (int)b[1 + x] + (y + 1) * (int)b[-1 - nT] + nT) >> (log2 + 1));
    }
  } else if (c.op == B200_DSP_INTRA_DC) {  // intrapred.h:288-322
    int dc = nT;
    for (int i = 0; i < nT; i++) dc += (int)b[i + 1] + (int)b[-i - 1];
    dc >>= log2 + 1;
    const bool edge = (cidx == 0 && nT < 32);
    for (int o = threadIdx.x; o < nT * nT; o += blockDim.x) {
      const int x = o & (nT - 1), y = o >> log2;
      int v = dc;
      if (edge) {
        if (x == 0 && y == 0) v = ((int)b[-1] + 2 * dc + (int)b[1] + 2) >> 2;
        else if (y == 0) v = ((int)b[x + 1] + 3 * dc + 2) >> 2;
        else if (x == 0) v = ((int)b[-y - 1] + 3 * dc + 2) >> 2;
      }
      put(x, y, v);
    }
  } else {  // angular, intrapred.h:330-433
    const int mode = c.a[2];
    const int angle = k_intra_angle[mode];
    const bool vert = mode >= 18;
    const int sgn = vert ? 1 : -1;
    int* ref = ref_s + 64;  // ref[x] valid on [-nT, 2nT]
    const int last = (nT * angle) >> 5;
    const int inv = (angle < 0) ? (int)k_inv_angle[mode - 11] : 0;
    const bool project = (angle < 0) && (last < -1);
    for (int s = threadIdx.x; s <= 3 * nT; s += blockDim.x) {
      const int x = s - nT;
      const bool w = (x >= 0) ? (x <= nT || angle >= 0) : (project && x >= last);
      const int idx = (x >= 0) ? sgn * x : -sgn * ((x * inv + 128) >> 8);
      if (w) ref[x] = b[idx];
    }
    __syncthreads();
    const bool bfilt = (cidx == 0 && nT < 32 && !c.a[3] && (mode == 26 || mode == 10));
    for (int o = threadIdx.x; o < nT * nT; o += blockDim.x) {
      const int x = o & (nT - 1), y = o >> log2;
      const int a = vert ? y : x, bb = vert ? x : y;
      const int idx = ((a + 1) * angle) >> 5, fact = ((a + 1) * angle) & 31;
      int v = fact ? ((32 - fact) * ref[bb + idx + 1] + fact * ref[bb + idx + 2] + 16) >> 5 : ref[bb + idx + 1];
      if (bfilt) {
        if (mode == 26 && x == 0) v = clip_bd((int)b[1] + (((int)b[-1 - y] - (int)b[0]) >> 1), c.bd);
        if (mode == 10 && y == 0) v = clip_bd((int)b[-1] + (((int)b[1 + x] - (int)b[0]) >> 1), c.bd);
      }
      put(x, y, v);
    }
  }
}

template <typename P>
__device__ void dsp_deblock(const DspDevCmd& c, uint8_t* arena)
{
  // fallback-deblk.h:32-124.  The staged window holds `side` samples each side of the edge for 4 lines:
  // vertical edge: 4 rows x 2*side samples, q0 at column `side`; horizontal: 2*side rows x 4 samples, q0 at row `side`.
  const int k = threadIdx.x;
  if (k >= 4) return;
  const bool luma = c.op == B200_DSP_DEBLOCK_LUMA;
  const int side = luma ? 4 : 2, vertical = c.a[0], bd = c.bd;
  P* base = reinterpret_cast<P*>(arena + c.io);
  const int pitch = c.p_io / (int)sizeof(P);
  const ptrdiff_t sa = vertical ? 1 : pitch, sb = vertical ? pitch : 1;
  P* e = base + side * sa + k * sb;  // q0 of line k
  if (luma) {
    const int dE = c.a[1], dEp = c.a[2], dEq = c.a[3], tc = c.a[4], fP = c.a[5], fQ = c.a[6];
    const int p0 = e[-sa], p1 = e[-2 * sa], p2 = e[-3 * sa], p3 = e[-4 * sa], q0 = e[0], q1 = e[sa], q2 = e[2 * sa], q3 = e[3 * sa];
    if (dE == 2) {
      if (fP) {
        e[-sa] = (P)clip3i(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
        e[-2 * sa] = (P)clip3i(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
        e[-3 * sa] = (P)clip3i(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
      }
      if (fQ) {
        e[0] = (P)clip3i(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
        e[sa] = (P)clip3i(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
        e[2 * sa] = (P)clip3i(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
      }
    } else {
      int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
      if (abs(delta) < tc * 10) {
        delta = clip3i(-tc, tc, delta);
        if (fP) e[-sa] = (P)clip_bd(p0 + delta, bd);
        if (fQ) e[0] = (P)clip_bd(q0 - delta, bd);
        if (dEp && fP) e[-2 * sa] = (P)clip_bd(p1 + clip3i(-(tc >> 1), tc >> 1, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1), bd);
        if (dEq && fQ) e[sa] = (P)clip_bd(q1 + clip3i(-(tc >> 1), tc >> 1, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1), bd);
      }
    }
  } else {
    const int tc = c.a[1], fP = c.a[2], fQ = c.a[3];
    const int p0 = e[-sa], p1 = e[-2 * sa], q0 = e[0], q1 = e[sa];
    const int delta = clip3i(-tc, tc, ((((q0 - p0) * 4) + p1 - q1 + 4) >> 3));
    if (fP) e[-sa] = (P)clip_bd(p0 + delta, bd);
    if (fQ) e[0] = (P)clip_bd(q0 - delta, bd);
  }
}

__global__ void __launch_bounds__(128) k_dsp_table(const DspDevCmd* __restrict__ cmds, uint8_t* arena)
{
  __shared__ int16_t s_buf[4 * MC_STRIP > 32 * 32 ? 4 * MC_STRIP : 32 * 32];
  __shared__ int s_ref[64 + 2 * 32 + 8];
  const DspDevCmd c = cmds[blockIdx.x];
  const bool wide = c.bd > 8;
  switch (c.op) {
    case B200_DSP_QPEL: wide ? dsp_mc<uint16_t>(c, arena, s_buf, true) : dsp_mc<uint8_t>(c, arena, s_buf, true); break;
    case B200_DSP_EPEL: wide ? dsp_mc<uint16_t>(c, arena, s_buf, false) : dsp_mc<uint8_t>(c, arena, s_buf, false); break;
    case B200_DSP_PRED_UNI: case B200_DSP_PRED_AVG: case B200_DSP_PRED_WEIGHTED: case B200_DSP_PRED_WEIGHTED_BI:
      wide ? dsp_pred<uint16_t>(c, arena) : dsp_pred<uint8_t>(c, arena); break;
    case B200_DSP_TRANSFORM_ADD: case B200_DSP_DST_ADD: wide ? dsp_transform<uint16_t>(c, arena, s_buf) : dsp_transform<uint8_t>(c, arena, s_buf); break;
    case B200_DSP_INTRA_DC: case B200_DSP_INTRA_PLANAR: case B200_DSP_INTRA_ANGULAR:
      wide ? dsp_intra<uint16_t>(c, arena, s_ref) : dsp_intra<uint8_t>(c, arena, s_ref); break;
    case B200_DSP_DEBLOCK_LUMA: case B200_DSP_DEBLOCK_CHROMA: wide ? dsp_deblock<uint16_t>(c, arena) : dsp_deblock<uint8_t>(c, arena); break;
    default: break;
  }
}

// ---- host side ----------------------------------------------------------------------------------
struct DspRegion {   // a 2-D block in host memory and its place in the arena
  uint8_t* host = nullptr;
  size_t row_bytes = 0, stride_bytes = 0;
  int rows = 0;
  uint32_t off = 0;
  int pitch = 0;
  bool in = false, out = false;
};

struct b200_dsp {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint8_t *h_arena = nullptr, *d_arena = nullptr;
  size_t cap = 0;
  DspDevCmd* d_cmds = nullptr;
  size_t cmd_cap = 0;
  std::vector<DspDevCmd> cmds;
  std::vector<DspRegion> regs;
};

extern "C" int b200_dsp_create(b200_dsp** out, int device)
{
  if (!out) return set_err(B200_ERR_INVALID, "null out");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return set_err(B200_ERR_NO_DEVICE, "no CUDA device available (%s); the DSP table has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= n) return set_err(B200_ERR_INVALID, "device %d out of range", device);
  CU(cudaSetDevice(device));
  int rc = init_tables(device);
  if (rc) return rc;
  b200_dsp* d = new (std::nothrow) b200_dsp();
  if (!d) return set_err(B200_ERR_NOMEM, "out of memory");
  d->device = device;
  CU(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  *out = d;
  return B200_OK;
}

extern "C" void b200_dsp_destroy(b200_dsp* d)
{
  if (!d) return;
  cudaSetDevice(d->device);
  if (d->stream) { cudaStreamSynchronize(d->stream); cudaStreamDestroy(d->stream); }
  if (d->h_arena) cudaFreeHost(d->h_arena);
  if (d->d_arena) cudaFree(d->d_arena);
  if (d->d_cmds) cudaFree(d->d_cmds);
  delete d;
}

extern "C" int b200_dsp_run_batch(b200_dsp* d, const b200_dsp_cmd* cmds, int n)
{
  if (!d || (!cmds && n)) return set_err(B200_ERR_INVALID, "null argument");
  if (n <= 0) return B200_OK;
  CU(cudaSetDevice(d->device));
  d->cmds.assign((size_t)n, DspDevCmd());
  d->regs.assign((size_t)n * 3, DspRegion());
  size_t total = 0;
  auto place = [&](DspRegion& r) {
    r.pitch = (int)align_up(r.row_bytes, 16);
    r.off = (uint32_t)total;
    total += align_up((size_t)r.pitch * r.rows, 256);
  };
  for (int i = 0; i < n; i++) {
    const b200_dsp_cmd& c = cmds[i];
    DspDevCmd& dc = d->cmds[i];
    DspRegion *r0 = &d->regs[3 * i], *r1 = r0 + 1, *rio = r0 + 2;
    if (c.bit_depth < 8 || c.bit_depth > 12 || !c.dst) return set_err(B200_ERR_INVALID, "dsp cmd %d: bit depth / dst", i);
    const size_t bps = c.bit_depth > 8 ? 2 : 1;
    dc.op = c.op; dc.bd = c.bit_depth; dc.w = c.w; dc.h = c.h;
    memcpy(dc.a, c.a, sizeof(dc.a));
    auto block = [&](DspRegion* r, const void* p, ptrdiff_t off_elems, size_t elem, int w, int rows, ptrdiff_t stride, bool in, bool out) {
      r->host = (uint8_t*)p + off_elems * (ptrdiff_t)elem;
      r->row_bytes = (size_t)w * elem; r->stride_bytes = (size_t)stride * elem; r->rows = rows; r->in = in; r->out = out;
    };
    switch (c.op) {
      case B200_DSP_QPEL: case B200_DSP_EPEL: {
        const bool luma = c.op == B200_DSP_QPEL;
        const int xf = c.a[0], yf = c.a[1], before = luma ? 3 : 1, extra = luma ? 7 : 3;
        if (!c.src || c.w < 2 || c.h < 2 || c.w > 64 || c.h > 64 || xf < 0 || yf < 0 || xf > (luma ? 3 : 7) || yf > (luma ? 3 : 7))
          return set_err(B200_ERR_INVALID, "dsp cmd %d: MC arguments", i);
        const int ox = xf ? before : 0, oy = yf ? before : 0;
        block(r0, c.src, -ox - oy * c.srcstride, bps, c.w + (xf ? extra : 0), c.h + (yf ? extra : 0), c.srcstride, true, false);
        block(rio, c.dst, 0, 2, c.w, c.h, c.dststride, false, true);
        break;
      }
      case B200_DSP_PRED_UNI: case B200_DSP_PRED_AVG: case B200_DSP_PRED_WEIGHTED: case B200_DSP_PRED_WEIGHTED_BI: {
        const bool two = c.op == B200_DSP_PRED_AVG || c.op == B200_DSP_PRED_WEIGHTED_BI;
        if (!c.src || (two && !c.src2) || c.w < 1 || c.h < 1 || c.w > 64 || c.h > 64) return set_err(B200_ERR_INVALID, "dsp cmd %d: weighting arguments", i);
        block(r0, c.src, 0, 2, c.w, c.h, c.srcstride, true, false);
        if (two) block(r1, c.src2, 0, 2, c.w, c.h, c.srcstride, true, false);
        block(rio, c.dst, 0, bps, c.w, c.h, c.dststride, false, true);
        break;
      }
      case B200_DSP_TRANSFORM_ADD: case B200_DSP_DST_ADD: {
        const int log2 = c.op == B200_DSP_DST_ADD ? 2 : c.a[0];
        if (!c.src || log2 < 2 || log2 > 5) return set_err(B200_ERR_INVALID, "dsp cmd %d: transform size", i);
        const int nT = 1 << log2;
        block(r0, c.src, 0, 2, nT * nT, 1, nT * nT, true, false);
        block(rio, c.dst, 0, bps, nT, nT, c.dststride, true, true);
        break;
      }
      case B200_DSP_INTRA_DC: case B200_DSP_INTRA_PLANAR: case B200_DSP_INTRA_ANGULAR: {
        const int nT = c.a[0];
        if (!c.src || (nT != 4 && nT != 8 && nT != 16 && nT != 32) || (c.op == B200_DSP_INTRA_ANGULAR && (c.a[2] < 2 || c.a[2] > 34)))
          return set_err(B200_ERR_INVALID, "dsp cmd %d: intra arguments", i);
        block(r0, c.src, -2 * nT, bps, 4 * nT + 1, 1, 4 * nT + 1, true, false);
        block(rio, c.dst, 0, bps, nT, nT, c.dststride, false, true);
        break;
      }
      case B200_DSP_DEBLOCK_LUMA: case B200_DSP_DEBLOCK_CHROMA: {
        const int side = c.op == B200_DSP_DEBLOCK_LUMA ? 4 : 2;
        if (c.a[0]) block(rio, c.dst, -side, bps, 2 * side, 4, c.dststride, true, true);
        else block(rio, c.dst, -side * c.dststride, bps, 4, 2 * side, c.dststride, true, true);
        break;
      }
      default: return set_err(B200_ERR_INVALID, "dsp cmd %d: unknown op %d", i, c.op);
    }
    if (r0->host) place(*r0);
    if (r1->host) place(*r1);
    place(*rio);
    dc.in0 = r0->off; dc.in1 = r1->off; dc.io = rio->off;
    dc.p_in0 = r0->pitch; dc.p_in1 = r1->pitch; dc.p_io = rio->pitch;
  }
  if (d->cap < total) {
    if (d->h_arena) cudaFreeHost(d->h_arena);
    if (d->d_arena) cudaFree(d->d_arena);
    d->h_arena = nullptr; d->d_arena = nullptr;
    d->cap = align_up(total * 2, 1 << 16);
    CU(cudaMallocHost(&d->h_arena, d->cap));
    CU(cudaMalloc(&d->d_arena, d->cap));
  }
  if (d->cmd_cap < (size_t)n) {
    if (d->d_cmds) cudaFree(d->d_cmds);
    d->d_cmds = nullptr;
    d->cmd_cap = (size_t)n * 2;
    CU(cudaMalloc(&d->d_cmds, d->cmd_cap * sizeof(DspDevCmd)));
  }
  for (const DspRegion& r : d->regs)
    if (r.host && r.in)
      for (int y = 0; y < r.rows; y++) memcpy(d->h_arena + r.off + (size_t)y * r.pitch, r.host + (size_t)y * r.stride_bytes, r.row_bytes);
  CU(cudaMemcpyAsync(d->d_arena, d->h_arena, total, cudaMemcpyHostToDevice, d->stream));
  CU(cudaMemcpyAsync(d->d_cmds, d->cmds.data(), (size_t)n * sizeof(DspDevCmd), cudaMemcpyHostToDevice, d->stream));
  k_dsp_table<<<n, 128, 0, d->stream>>>(d->d_cmds, d->d_arena);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(d->h_arena, d->d_arena, total, cudaMemcpyDeviceToHost, d->stream));
  CU(cudaStreamSynchronize(d->stream));
  for (const DspRegion& r : d->regs)
    if (r.host && r.out)
      for (int y = 0; y < r.rows; y++) memcpy(r.host + (size_t)y * r.stride_bytes, d->h_arena + r.off + (size_t)y * r.pitch, r.row_bytes);
  return B200_OK;
}
