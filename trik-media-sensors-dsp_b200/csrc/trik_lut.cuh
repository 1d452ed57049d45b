// trik_lut.cuh -- entries of the chroma-indexed HSV table (built by chroma_table_kernel, trik_kernels_lut.cu) and the
// pass test of a pixel pair against them; shared by the WO / OO table kernels and the WO preview.
#pragma once
#include <cstdint>

namespace trikb200 {

// (lo, nhi = 255 - hi) codes with lo > hi: no luma satisfies lo <= Y <= hi, so the interval test fails by itself
constexpr uint32_t LUT_NEVER_LO = 255u, LUT_NEVER_NHI = 255u;        // hi = 0
constexpr uint32_t LUT_RAGGED_LO = 255u, LUT_RAGGED_NHI = 254u;      // hi = 1: marks "consult the mask"
constexpr uint32_t LUT_RAGGED_CODE = LUT_RAGGED_LO * 256u + LUT_RAGGED_NHI;     // lo * 256 + nhi
constexpr uint32_t LUT_STRIDE_PLAIN = 65536u, LUT_STRIDE_SKEW = 66560u;
constexpr uint32_t LUT_SKEW_IMAGE_OFS = 2u * LUT_STRIDE_PLAIN;                  // second image inside the table buffer

// Pass bits of a YUYV pixel pair (bit 15: pixel 0, bit 31: pixel 1) from its table entry: pass <=> lo <= Y <= hi.
// Guard bit 15 in every 16-bit lane keeps the two subtractions of one 32-bit operation apart (0x8000 + a - b with
// bytes a, b never borrows) and doubles as the result: it stays set <=> a >= b.  Y <= hi is tested as
// 255 - Y >= 255 - hi, so that both compares are one IMAD each (lane replication by 0x10001 included).
__device__ __forceinline__ uint32_t lut_pass_pair(uint32_t word, uint32_t lo, uint32_t nhi)
{
  const uint32_t yG  = (word & 0x00FF00FFu) | 0x80008000u;             // 0x8000 + Y per lane
  const uint32_t nyG = (~word & 0x00FF00FFu) | 0x80008000u;            // 0x8000 + 255 - Y per lane
  const uint32_t geLo = yG - lo * 0x00010001u;
  const uint32_t leHi = nyG - nhi * 0x00010001u;
  return geLo & leHi & 0x80008000u;
}

// pass bits of a pair whose entry is RAGGED: the luma masks decide
__device__ __forceinline__ uint32_t lut_pass_pair_masks(uint32_t word, uint32_t ci, const uint32_t* __restrict__ masks)
{
  const uint32_t y0 = word & 0xFFu, y1 = (word >> 16) & 0xFFu;
  const uint32_t m0 = __ldg(masks + (size_t)ci * 8u + (y0 >> 5));
  const uint32_t m1 = __ldg(masks + (size_t)ci * 8u + (y1 >> 5));
  return (((m0 >> (y0 & 31u)) & 1u) << 15) | (((m1 >> (y1 & 31u)) & 1u) << 31);
}

} // namespace trikb200
