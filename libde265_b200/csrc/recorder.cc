// recorder.cc — host-side command recorder (b200hevc.h part 3) and picture (de)serialisation.
//
// The recorder is what the reference-side hooks call while libde265 parses a picture
// (INTEGRATION.md): decode_TU (slice.cc:3460) -> b200_rec_add_tu, generate_inter_prediction_samples
// (motion.cc:288) -> b200_rec_add_pu, picture completion (decctx.cc:605-650) -> b200_rec_end_picture.
// Plain C++ (no CUDA), part of libb200hevc.so.

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "b200hevc.h"

struct b200_recorder {
  b200_pic_params params{};
  int w4 = 0, h4 = 0, w8 = 0, h8 = 0, wctb = 0, hctb = 0;
  std::vector<b200_pu> pus;
  std::vector<b200_weight_entry> weights;
  std::vector<b200_tu> tus;
  std::vector<b200_coeff> coeffs;
  std::vector<b200_slice_info> slices;
  std::vector<b200_ctb_info> ctbs;
  std::vector<uint8_t> bs_map;
  std::vector<int8_t> qp_map;
  std::vector<uint8_t> nofilt_map;
  std::vector<uint8_t> scaling;
  bool has_scaling = false;
  bool open = false;
};

extern "C" {

int b200_abi_version(void) { return B200_ABI_VERSION; }

int b200_rec_create(b200_recorder** out)
{
  if (!out) return B200_ERR_INVALID;
  b200_recorder* r = new (std::nothrow) b200_recorder();
  if (!r) return B200_ERR_NOMEM;
  *out = r;
  return B200_OK;
}

void b200_rec_destroy(b200_recorder* r) { delete r; }

int b200_rec_begin_picture(b200_recorder* r, const b200_pic_params* p)
{
  if (!r || !p) return B200_ERR_INVALID;
  if (p->width == 0 || p->height == 0 || p->log2_ctb_size < 3 || p->log2_ctb_size > 6) return B200_ERR_INVALID;
  if (p->dst_slot >= B200_MAX_SLOTS) return B200_ERR_INVALID;
  r->params = *p;
  r->w4 = (p->width + 3) / 4;
  r->h4 = (p->height + 3) / 4;
  r->w8 = (p->width + 7) / 8;
  r->h8 = (p->height + 7) / 8;
  const int ctb = 1 << p->log2_ctb_size;
  r->wctb = (p->width + ctb - 1) / ctb;
  r->hctb = (p->height + ctb - 1) / ctb;
  try {
    r->pus.clear();
    r->weights.clear();
    r->tus.clear();
    r->coeffs.clear();
    r->slices.clear();
    r->ctbs.assign((size_t)r->wctb * r->hctb, b200_ctb_info{});
    r->bs_map.assign((size_t)r->w4 * r->h4, 0);
    r->qp_map.assign((size_t)r->w8 * r->h8, 0);
    r->nofilt_map.assign((size_t)r->w8 * r->h8, 0);
  } catch (const std::bad_alloc&) {
    return B200_ERR_NOMEM;
  }
  r->has_scaling = false;
  r->open = true;
  return B200_OK;
}

int b200_rec_add_slice(b200_recorder* r, const b200_slice_info* s)
{
  if (!r || !s || !r->open) return B200_ERR_INVALID;
  r->slices.push_back(*s);
  return (int)r->slices.size() - 1;
}

int b200_rec_add_weights(b200_recorder* r, const b200_weight_entry* w)
{
  if (!r || !w || !r->open) return B200_ERR_INVALID;
  if (r->weights.size() >= 65535) return B200_ERR_INVALID;
  r->weights.push_back(*w);
  return (int)r->weights.size() - 1;
}

int b200_rec_add_pu(b200_recorder* r, const b200_pu* pu)
{
  if (!r || !pu || !r->open) return B200_ERR_INVALID;
  if (pu->w == 0 || pu->h == 0 || pu->w > 64 || pu->h > 64) return B200_ERR_INVALID;
  if ((unsigned)pu->x + pu->w > r->params.width || (unsigned)pu->y + pu->h > r->params.height) return B200_ERR_INVALID;
  r->pus.push_back(*pu);
  return B200_OK;
}

int b200_rec_add_tu(b200_recorder* r, const b200_tu* tu, const int16_t* levels, const int16_t* positions, int n)
{
  if (!r || !tu || !r->open || n < 0) return B200_ERR_INVALID;
  if (tu->log2_size < 2 || tu->log2_size > 5 || tu->cidx > 2) return B200_ERR_INVALID;
  const int nT = 1 << tu->log2_size;
  if (n > nT * nT || (n && (!levels || !positions))) return B200_ERR_INVALID;
  b200_tu t = *tu;
  t.coeff_off = (uint32_t)r->coeffs.size();
  t.n_coeff = (uint16_t)n;
  for (int i = 0; i < n; i++) {
    if ((unsigned)positions[i] >= (unsigned)(nT * nT)) return B200_ERR_INVALID;
    b200_coeff c;
    c.pos = (uint16_t)positions[i];
    c.level = levels[i];
    r->coeffs.push_back(c);
  }
  r->tus.push_back(t);
  return B200_OK;
}

int b200_rec_set_ctb(b200_recorder* r, int ctb_x, int ctb_y, const b200_ctb_info* c)
{
  if (!r || !c || !r->open) return B200_ERR_INVALID;
  if (ctb_x < 0 || ctb_y < 0 || ctb_x >= r->wctb || ctb_y >= r->hctb) return B200_ERR_INVALID;
  r->ctbs[(size_t)ctb_x + (size_t)ctb_y * r->wctb] = *c;
  return B200_OK;
}

uint8_t* b200_rec_bs_map(b200_recorder* r) { return (r && r->open) ? r->bs_map.data() : nullptr; }
int8_t* b200_rec_qp_map(b200_recorder* r) { return (r && r->open) ? r->qp_map.data() : nullptr; }
uint8_t* b200_rec_nofilt_map(b200_recorder* r) { return (r && r->open) ? r->nofilt_map.data() : nullptr; }

int b200_rec_set_scaling_factors(b200_recorder* r, const uint8_t* f)
{
  if (!r || !f || !r->open) return B200_ERR_INVALID;
  r->scaling.assign(f, f + B200_SCALING_FACTOR_BYTES);
  r->has_scaling = true;
  return B200_OK;
}

int b200_rec_end_picture(b200_recorder* r, b200_picture* out)
{
  if (!r || !out || !r->open) return B200_ERR_INVALID;
  if (r->slices.empty()) return B200_ERR_INVALID;
  for (const b200_ctb_info& c : r->ctbs)
    if (c.slice_idx >= r->slices.size()) return B200_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  out->params = r->params;
  out->n_pu = (uint32_t)r->pus.size();
  out->n_weights = (uint32_t)r->weights.size();
  out->n_tu = (uint32_t)r->tus.size();
  out->n_coeff = (uint32_t)r->coeffs.size();
  out->n_slices = (uint32_t)r->slices.size();
  out->pus = r->pus.data();
  out->weights = r->weights.data();
  out->tus = r->tus.data();
  out->coeffs = r->coeffs.data();
  out->slices = r->slices.data();
  out->ctbs = r->ctbs.data();
  out->bs_map = r->bs_map.data();
  out->qp_map = r->qp_map.data();
  out->nofilt_map = r->nofilt_map.data();
  out->scaling_factors = r->has_scaling ? r->scaling.data() : nullptr;
  r->open = false;
  return B200_OK;
}

// ---- serialisation -----------------------------------------------------------------------------
// Layout: header {magic, version, params, counts[5], has_scaling, map sizes} followed by the arrays,
// each padded to 8 bytes.

struct ser_header {
  uint32_t magic;  // 'B2HV'
  uint32_t version;
  b200_pic_params params;
  uint32_t n_pu, n_weights, n_tu, n_coeff, n_slices, n_ctb, n_bs, n_q8, has_scaling;
};
static const uint32_t kMagic = 0x56483242u;

static size_t pad8(size_t n) { return (n + 7) & ~(size_t)7; }

static void geom(const b200_pic_params& p, uint32_t* n_ctb, uint32_t* n_bs, uint32_t* n_q8)
{
  const int ctb = 1 << p.log2_ctb_size;
  *n_ctb = (uint32_t)(((p.width + ctb - 1) / ctb) * ((p.height + ctb - 1) / ctb));
  *n_bs = (uint32_t)(((p.width + 3) / 4) * ((p.height + 3) / 4));
  *n_q8 = (uint32_t)(((p.width + 7) / 8) * ((p.height + 7) / 8));
}

size_t b200_picture_serialized_size(const b200_picture* pic)
{
  if (!pic) return 0;
  uint32_t n_ctb, n_bs, n_q8;
  geom(pic->params, &n_ctb, &n_bs, &n_q8);
  size_t n = pad8(sizeof(ser_header));
  n += pad8(sizeof(b200_pu) * pic->n_pu) + pad8(sizeof(b200_weight_entry) * pic->n_weights) + pad8(sizeof(b200_tu) * pic->n_tu) +
       pad8(sizeof(b200_coeff) * pic->n_coeff) + pad8(sizeof(b200_slice_info) * pic->n_slices) + pad8(sizeof(b200_ctb_info) * n_ctb) +
       pad8(pic->bs_map ? n_bs : 0) + pad8(n_q8) + pad8(n_q8) + pad8(pic->scaling_factors ? B200_SCALING_FACTOR_BYTES : 0);
  return n;
}

size_t b200_picture_serialize(const b200_picture* pic, void* buf, size_t cap)
{
  size_t need = b200_picture_serialized_size(pic);
  if (!need || !buf || cap < need) return 0;
  uint8_t* p = (uint8_t*)buf;
  memset(p, 0, need);
  ser_header h{};
  h.magic = kMagic;
  h.version = B200_ABI_VERSION;
  h.params = pic->params;
  h.n_pu = pic->n_pu; h.n_weights = pic->n_weights; h.n_tu = pic->n_tu; h.n_coeff = pic->n_coeff; h.n_slices = pic->n_slices;
  uint32_t n_bs_full;
  geom(pic->params, &h.n_ctb, &n_bs_full, &h.n_q8);
  h.n_bs = pic->bs_map ? n_bs_full : 0;
  h.has_scaling = pic->scaling_factors ? 1 : 0;
  memcpy(p, &h, sizeof(h));
  size_t off = pad8(sizeof(ser_header));
  auto put = [&](const void* src, size_t bytes) {
    if (bytes) memcpy(p + off, src, bytes);
    off += pad8(bytes);
  };
  put(pic->pus, sizeof(b200_pu) * pic->n_pu);
  put(pic->weights, sizeof(b200_weight_entry) * pic->n_weights);
  put(pic->tus, sizeof(b200_tu) * pic->n_tu);
  put(pic->coeffs, sizeof(b200_coeff) * pic->n_coeff);
  put(pic->slices, sizeof(b200_slice_info) * pic->n_slices);
  put(pic->ctbs, sizeof(b200_ctb_info) * h.n_ctb);
  put(pic->bs_map, h.n_bs);
  put(pic->qp_map, h.n_q8);
  put(pic->nofilt_map, h.n_q8);
  put(pic->scaling_factors, h.has_scaling ? B200_SCALING_FACTOR_BYTES : 0);
  return off;
}

size_t b200_picture_deserialize(const void* buf, size_t len, b200_picture* out)
{
  if (!buf || !out || len < sizeof(ser_header)) return 0;
  ser_header h;
  memcpy(&h, buf, sizeof(h));
  if (h.magic != kMagic || h.version != B200_ABI_VERSION) return 0;
  uint32_t n_ctb, n_bs, n_q8;
  if (h.params.width == 0 || h.params.height == 0 || h.params.log2_ctb_size < 3 || h.params.log2_ctb_size > 6) return 0;
  geom(h.params, &n_ctb, &n_bs, &n_q8);
  if (h.n_ctb != n_ctb || h.n_q8 != n_q8 || (h.n_bs != 0 && h.n_bs != n_bs)) return 0;
  const uint8_t* p = (const uint8_t*)buf;
  size_t off = pad8(sizeof(ser_header));
  bool ok = true;
  auto get = [&](size_t bytes) -> const void* {
    const void* r = bytes ? p + off : nullptr;
    if (off + pad8(bytes) > len) { ok = false; return nullptr; }
    off += pad8(bytes);
    return r;
  };
  memset(out, 0, sizeof(*out));
  out->params = h.params;
  out->n_pu = h.n_pu; out->n_weights = h.n_weights; out->n_tu = h.n_tu; out->n_coeff = h.n_coeff; out->n_slices = h.n_slices;
  out->pus = (const b200_pu*)get(sizeof(b200_pu) * (size_t)h.n_pu);
  out->weights = (const b200_weight_entry*)get(sizeof(b200_weight_entry) * (size_t)h.n_weights);
  out->tus = (const b200_tu*)get(sizeof(b200_tu) * (size_t)h.n_tu);
  out->coeffs = (const b200_coeff*)get(sizeof(b200_coeff) * (size_t)h.n_coeff);
  out->slices = (const b200_slice_info*)get(sizeof(b200_slice_info) * (size_t)h.n_slices);
  out->ctbs = (const b200_ctb_info*)get(sizeof(b200_ctb_info) * (size_t)h.n_ctb);
  out->bs_map = (const uint8_t*)get(h.n_bs);
  out->qp_map = (const int8_t*)get(h.n_q8);
  out->nofilt_map = (const uint8_t*)get(h.n_q8);
  out->scaling_factors = (const uint8_t*)get(h.has_scaling ? B200_SCALING_FACTOR_BYTES : 0);
  if (!ok) return 0;
  return off;
}

}  // extern "C"
