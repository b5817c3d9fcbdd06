// sao8_swar.cuh — the byte-parallel (SIMD-within-a-register) primitives of k_sao8 (kernels_filter.cuh).  __host__ __device__ so that
// tests/sao8_emul.cu can check them exhaustively on the CPU (tests/test_cpu_sao8_swar.py); on the device sao8_rep is one PRMT.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define SAO8_HD __host__ __device__ __forceinline__
#else
#define SAO8_HD inline
#endif

SAO8_HD uint32_t sao8_rep(uint32_t x)  // 0xFF in every byte whose bit 7 is set
{
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(r) : "r"(x));
  return r;
#else
  return ((x >> 7) & 0x01010101u) * 0xFFu;
#endif
}
SAO8_HD uint32_t sao8_lt(uint32_t x, uint32_t y)  // byte mask: x < y (unsigned)
{
  const uint32_t d = (x | 0x80808080u) - (y & 0x7F7F7F7Fu);
  return sao8_rep((~x & y) | (~(x ^ y) & ~d));
}
SAO8_HD uint32_t sao8_eq_small(uint32_t k, uint32_t cst)  // byte mask: k == cst, for bytes < 0x80
{
  const uint32_t z = k ^ cst;
  return ~sao8_rep(z + 0x7F7F7F7Fu);  // bit 7 set <=> byte nonzero
}
SAO8_HD uint32_t sao8_apply(uint32_t s, uint32_t pos, uint32_t neg)  // clip(s + pos - neg, 0, 255) per byte; pos, neg < 128
{
  // saturating add: low 7 bits with carry into bit 7, then the true bit 7; overflow <=> s's bit 7 set and the sum's clear
  const uint32_t t = ((s & 0x7F7F7F7Fu) + pos) ^ (s & 0x80808080u);
  const uint32_t a = t | sao8_rep(s & ~t);
  // saturating subtract: (a | 0x80) - neg never borrows; bit 7 of u <=> low7(a) >= neg
  const uint32_t u = (a | 0x80808080u) - neg;
  const uint32_t hi = sao8_rep(a), ok = sao8_rep(u);
  return (hi & u) | (~hi & ok & u & 0x7F7F7F7Fu);
}
SAO8_HD uint32_t sao8_mask4(unsigned bits)  // 4 bits -> 4 byte masks
{
  return sao8_rep((((bits & 0xFu) * 0x00204081u) & 0x01010101u) * 0x80u);
}
