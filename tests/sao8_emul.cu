// sao8_emul.cu — TEST INFRASTRUCTURE: exhaustive CPU check of k_sao8's byte-parallel primitives (libde265_b200/csrc/sao8_swar.cuh, the
// same functions the GPU executes) against their per-byte definitions.  Built by tests/test_cpu_sao8_swar.py as a host-only library.
#include <cstdint>
#include "sao8_swar.cuh"

namespace {
inline uint32_t lane4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) { return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24); }
}

// number of mismatches (0 = all good)
extern "C" __attribute__((visibility("default"))) long sao8_check_lt()
{
  long bad = 0;
  for (uint32_t x = 0; x < 256; x++)
    for (uint32_t y = 0; y < 256; y++) {
      // the pair in byte 1, neighbours chosen to provoke borrows / carries across byte borders
      const uint32_t X = lane4(0x00, x, 0xFF, 0x80), Y = lane4(0xFF, y, 0x00, 0x80);
      const uint32_t m = sao8_lt(X, Y);
      const uint32_t want = lane4(0xFF, x < y ? 0xFF : 0, 0, 0);
      if (m != want) bad++;
    }
  return bad;
}

extern "C" __attribute__((visibility("default"))) long sao8_check_apply()
{
  long bad = 0;
  for (uint32_t s = 0; s < 256; s++)
    for (uint32_t p = 0; p < 128; p++)
      for (uint32_t n = 0; n < (p ? 1u : 128u); n++) {  // an offset is one-sided: positive part or negative part, never both
        const uint32_t S = lane4(0xFF, s, 0x00, s), P = lane4(127, p, 0, 0), N = lane4(0, n, 127, n);
        const uint32_t r = sao8_apply(S, P, N);
        auto clip = [](int v) { return (uint32_t)(v < 0 ? 0 : v > 255 ? 255 : v); };
        const uint32_t want = lane4(clip(255 + 127), clip((int)s + (int)p - (int)n), clip(0 - 127), clip((int)s - (int)n));
        if (r != want) bad++;
      }
  return bad;
}

extern "C" __attribute__((visibility("default"))) long sao8_check_eq_mask()
{
  long bad = 0;
  for (uint32_t k = 0; k < 64; k++)
    for (uint32_t c = 0; c < 4; c++) {
      const uint32_t K = lane4(k, 0, 63, c), Cst = c * 0x01010101u;
      const uint32_t want = lane4(k == c ? 0xFF : 0, c == 0 ? 0xFF : 0, 0, 0xFF);
      if (sao8_eq_small(K, Cst) != want) bad++;
    }
  for (uint32_t b = 0; b < 16; b++) {
    const uint32_t want = lane4(b & 1 ? 0xFF : 0, b & 2 ? 0xFF : 0, b & 4 ? 0xFF : 0, b & 8 ? 0xFF : 0);
    if (sao8_mask4(b) != want || sao8_mask4(b | 0xF0) != want) bad++;
  }
  return bad;
}
