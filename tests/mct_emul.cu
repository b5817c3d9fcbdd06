// mct_emul.cu — TEST INFRASTRUCTURE: runs the task bodies of the TMA-staged MC kernel (libde265_b200/csrc/kernels_mct.cuh, the
// very same __host__ __device__ functions the GPU executes) on the CPU, batch by batch, with the TMA box fetch replaced by a
// copy out of a padded reference surface.  tests/test_cpu_mct_emul.py compares the result with the oracle's inter-prediction
// stage, so that the kernel's arithmetic and index logic are checked without a GPU; staging (TMA, mbarrier) is device-only.
// Built by tests/test_cpu_mct_emul.py with nvcc as a host-only shared library (tests/libmct_emul.so); not part of the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "kernels_mct.cuh"

namespace {
struct Padded {  // one plane with a replicated border, as engine.cu builds them
  std::vector<uint8_t> mem;
  int pitch = 0, rows = 0, padx = 0, pady = 0;
  void build(const uint8_t* src, size_t stride, int w, int h, int px, int py)
  {
    padx = px; pady = py;
    pitch = (w + 2 * px + 255) / 256 * 256;
    rows = h + 2 * py;
    mem.assign((size_t)pitch * rows, 0);
    for (int y = -py; y < h + py; y++)
      for (int x = -px; x < w + px; x++) {
        const int ys = y < 0 ? 0 : y >= h ? h - 1 : y, xs = x < 0 ? 0 : x >= w ? w - 1 : x;
        mem[(size_t)(y + py) * pitch + x + px] = src[(size_t)ys * stride + xs];
      }
  }
  // what a TMA box load delivers: `bw` bytes x `bh` rows from allocation coordinates (x, y), zero outside
  void box(uint8_t* dst, int x, int y, int bw, int bh) const
  {
    for (int r = 0; r < bh; r++)
      for (int c = 0; c < bw; c++) {
        const int yy = y + r, xx = x + c;
        dst[r * bw + c] = (yy >= 0 && yy < rows && xx >= 0 && xx < pitch) ? mem[(size_t)yy * pitch + xx] : 0;
      }
  }
};
}  // namespace

// ref_planes[slot][c]: tight planes (stride = width) or null; dst planes tight.  tiles: the planner's class-sorted tile words.
extern "C" __attribute__((visibility("default"))) int mct_emulate(const b200_picture* pic, const uint8_t* const* ref_planes /* [32*3] */,
                                                                     uint8_t* const dst[3], const uint32_t* tiles, int n_tiles)
{
  const b200_pic_params& p = pic->params;
  const int W = p.width, H = p.height, CW = p.chroma_format_idc ? W / 2 : 0, CH = p.chroma_format_idc ? H / 2 : 0;
  DevPic dp{};
  dp.w = W; dp.h = H; dp.cw = CW; dp.ch = CH; dp.chroma = p.chroma_format_idc;
  dp.pitch[0] = W; dp.pitch[1] = dp.pitch[2] = CW;
  for (int c = 0; c < 3; c++) dp.cur[c] = dst[c];
  static Padded pad[B200_MAX_SLOTS][3];
  uint32_t valid = 0;
  for (int s = 0; s < B200_MAX_SLOTS; s++) {
    if (!ref_planes[s * 3]) continue;
    valid |= 1u << s;
    pad[s][0].build(ref_planes[s * 3], W, W, H, B200_PAD_X, B200_PAD_Y);
    for (int c = 1; c < 3 && CW; c++) pad[s][c].build(ref_planes[s * 3 + c], CW, CW, CH, B200_PAD_CX, B200_PAD_CY);
  }
  static MctShared sm;
  mc8_build_tables(sm.tab);
  int it = 0;
  for (int first = 0; first < n_tiles; it++) {
    const uint32_t w0 = tiles[first];
    if (w0 == MCT_INVALID) return -2;  // a batch starts with a real tile
    const int cls = (w0 >> 24) & 7;
    const MctGeom g = mct_geom(cls);
    if (g.ntiles != MCT_CLASS_TILES(cls) || first + g.ntiles > n_tiles) return -1;
    MctTile* info = sm.info[it & 1];
    memset(sm.win[0], 0xAB, sizeof(sm.win[0]));  // poison: stale data must never matter
    memset(sm.interm, 0xCD, sizeof(sm.interm));
    for (int tid = 0; tid < g.ntl; tid++) {  // the producer threads
      const int tile = g.nl == 2 ? (tid >> 1) : tid, s = g.nl == 2 ? (tid & 1) : 0;
      const MctBox bx = mct_decode_tile(tiles[first + tile], s, pic->pus, pic->weights, valid, dp, &info[tile]);
      if (!bx.active) continue;
      const int skew = tid & 3;
      pad[bx.slot][0].box(sm.win[0] + tid * g.lw_slot, bx.lx + B200_PAD_X, bx.ly + B200_PAD_Y - skew, g.lw_pitch, g.small ? MCT_LWS_ROWS : MCT_LWB_ROWS);
      if (CW)
        for (int c = 0; c < 2; c++)
          pad[bx.slot][1 + c].box(sm.win[0] + g.cw_off + tid * g.cw_slot + c * g.cw_plane, bx.cx + B200_PAD_CX, bx.cy + B200_PAD_CY - skew, g.cw_pitch,
                                  g.small ? MCT_CWS_ROWS : MCT_CWB_ROWS);
    }
    for (int t = 0; t < g.n1l; t++) mct_pass1_luma(t, g, info, sm.win[0], sm.interm, sm.tab);
    if (CW)
      for (int t = 0; t < g.n1c; t++) mct_pass1_chroma(t, g, info, sm.win[0], sm.interm, sm.tab);
    for (int t = 0; t < g.n2l; t++) mct_pass2_luma(t, g, info, sm.interm, sm.tab, dp.cur[0], dp.pitch[0]);
    if (CW)
      for (int t = 0; t < g.n2c; t++) mct_pass2_chroma(t, g, info, sm.interm, sm.tab, dp.cur[1], dp.cur[2], dp.pitch[1]);
    first += g.ntiles;
  }
  return 0;
}
