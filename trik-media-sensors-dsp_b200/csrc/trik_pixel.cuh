// trik_pixel.cuh -- per-pixel arithmetic of the sensor pipeline for sm_100a.
//
// What is computed (bit-exact with the reference's C674x code):
//   a4  YUV -> RGB888  <sensor>/include/internal/cv_ball_detector_seqpass.hpp:181-205
//   a5  RGB888 -> HSV  <sensor>/include/internal/cv_ball_detector_seqpass.hpp:207-249 (+ LUTs :400-406)
//   a6  range test     <sensor>/include/internal/cv_ball_detector_seqpass.hpp:171-179
//
// How (B200-first, not a transliteration of the C6x lanes):
//   * Two pixels share one chroma pair, so they travel together in one 32-bit register as two
//     16-bit lanes.  Blackwell has native lane-isolated 16x2 integer ops (SASS VIADD.16x2,
//     VIMNMX.U16x2, VIMNMX3.U16x2) and a byte dot product (IDP.4A), which is everything the
//     colour matrix needs.
//   * The reference computes R,G,B as int16 sums that WRAP (only blue can: 129*U-17672+74*Y
//     exceeds 32767 for 106 of the 65536 (Y,U) pairs, all with U >= 245), then >>6 and saturates
//     to 0..255, so those pixels come out with B = 0 instead of 255.
//     Here every lane holds key = (value + 0x8000) mod 2^16, so unsigned lane order equals the
//     reference's signed order and the blue wrap is simply the lane wrap of VIADD.16x2.
//     Red and green keys cannot overflow a lane (max 63400 / 60334), so their "replicate the
//     chroma term into both lanes and add luma" is a single IMAD; blue needs PRMT + VIADD.16x2.
//   * >>6 and saturation are monotonic, so V = max(R,G,B) is taken on the keys:
//     V = sat8((max3(keys) - 0x8000) >> 6).  The line sensors threshold V only, so they never
//     leave the key domain: 16 integer instructions per two pixels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace trikb200 {

// ---- byte dot products (SASS IDP.4A) ---------------------------------------------------------
__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int32_t dp4a_us(uint32_t a, uint32_t b_signed_bytes, int32_t c)
{
  int32_t d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b_signed_bytes), "r"(c));
  return d;
}

// ---- chroma coefficient sets -----------------------------------------------------------------
// A "chroma word" is any 32-bit word that holds U and V of a pixel pair at fixed byte positions;
// the coefficient words put 102 (V->R), -52/-25 (V,U->G) and 129 (U->B) on those bytes and zero
// on the others, so the luma bytes of a YUYV word drop out of the dot product.
struct ChromaCoef { uint32_t r, g, b; };
// YUYV word  [Y0 U Y1 V]: U = byte 1, V = byte 3
__device__ __forceinline__ constexpr ChromaCoef coef_yuyv()  { return {0x66000000u, 0xCC00E700u, 0x00008100u}; }
// YUV422P chroma word [V0 U0 V1 U1] (ov7670/object_sensor/.../cv_ball_detector_seqpass.hpp:369-373): pair 0 / pair 1
__device__ __forceinline__ constexpr ChromaCoef coef_planar0() { return {0x00000066u, 0x0000E7CCu, 0x00008100u}; }
__device__ __forceinline__ constexpr ChromaCoef coef_planar1() { return {0x00660000u, 0xE7CC0000u, 0x81000000u}; }

constexpr uint32_t KEY_BIAS = 0x8000u;

// keys of R,G,B (before >>6 and saturation) for the two pixels whose lumas sit in the low bytes of
// the two 16-bit lanes of yy (0x00YY00YY) and whose chroma is addressed by (cw, coef).
__device__ __forceinline__ void rgb_keys(uint32_t yy, uint32_t cw, const ChromaCoef coef,
                                         uint32_t& kr, uint32_t& kg, uint32_t& kb)
{
  const uint32_t y74 = yy * 74u;                                            // lanes <= 18870
  const uint32_t cr = dp4a_uu(cw, coef.r, KEY_BIAS - 14248u);               // 102*V - 14248 + bias, 18520..44530
  const uint32_t cg = (uint32_t)dp4a_us(cw, coef.g, (int32_t)(KEY_BIAS + 8696u)); // -52*V - 25*U + 8696 + bias, 21829..41464
  const uint32_t cb = dp4a_uu(cw, coef.b, KEY_BIAS - 17672u);               // 129*U - 17672 + bias, 15096..47991
  kr = cr * 0x10001u + y74;                                                 // no lane carry possible
  kg = cg * 0x10001u + y74;
  kb = __vadd2(__byte_perm(cb, 0u, 0x1010), y74);                           // lane wrap == the reference's int16 wrap
}

// 8-bit channel lanes (0x00CC00CC) from key lanes: clamp to [0,16383] in the value domain, >>6.
__device__ __forceinline__ uint32_t chan8_from_key(uint32_t k)
{
  const uint32_t c = __vminu2(__vmaxu2(k, 0x80008000u), 0xBFFFBFFFu);
  return (c >> 6) & 0x00FF00FFu;   // bit 15 of each lane lands on bit 9 and is masked off with the spill of the upper lane
}

// ---- HSV of one pixel from 8-bit R,G,B, with the two division LUTs in shared memory ----------
// lut43[i] = 43*256/i, lut255[i] = 255*256/i, entry 0 = 0  (cv_ball_detector_seqpass.hpp:400-406)
__device__ __forceinline__ void fill_div_luts(uint16_t* lut43, uint16_t* lut255)
{
  for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x)
  {
    lut43[i]  = i ? (uint16_t)(11008u / i) : (uint16_t)0;
    lut255[i] = i ? (uint16_t)(65280u / i) : (uint16_t)0;
  }
}

// returns 0x00VVSSHH
__device__ __forceinline__ uint32_t hsv_from_rgb8(int32_t r, int32_t g, int32_t b,
                                                  const uint16_t* __restrict__ lut43,
                                                  const uint16_t* __restrict__ lut255)
{
  const int32_t mx = __vimax3_s32(r, g, b);
  const int32_t mn = __vimin3_s32(r, g, b);
  const int32_t d = mx - mn;
  const uint32_t sat_x256 = (uint32_t)lut255[mx] * (uint32_t)d;
  const int32_t t43 = (int32_t)lut43[d];
  const bool gmax = (mx == g), bmax = (mx == b);
  // _cmpeq2 result 0 -> red sector; 1 (blue only) -> blue sector; 2,3 -> green sector (:231-242)
  int32_t off, diff;
  if (gmax)      { off = 21845; diff = b - r; }
  else if (bmax) { off = 43690; diff = r - g; }
  else           { off = 0;     diff = g - b; }
  const uint32_t hue_x256 = (uint32_t)(off + t43 * diff);
  return ((uint32_t)mx << 16) | (sat_x256 & 0xFF00u) | ((hue_x256 >> 8) & 0xFFu);
}

// detectHsvPixel (:171-179): ((hsv <u4 from) | (hsv >u4 to)) == expected, bytes H,S,V (byte 3 is 0 everywhere).
__device__ __forceinline__ bool detect_hsv(uint32_t hsv, uint32_t from, uint32_t to, uint32_t expected)
{
  const uint32_t h = hsv & 0xFFu, s = (hsv >> 8) & 0xFFu, v = hsv >> 16;
  const bool hout = (h < (from & 0xFFu)) | (h > (to & 0xFFu));
  const bool sout = (s < ((from >> 8) & 0xFFu)) | (s > ((to >> 8) & 0xFFu));
  const bool vout = (v < ((from >> 16) & 0xFFu)) | (v > ((to >> 16) & 0xFFu));
  const uint32_t mask = (hout ? 1u : 0u) | (sout ? 2u : 0u) | (vout ? 4u : 0u);
  return mask == expected;
}

// ---- packed HSV of a pixel pair ----------------------------------------------------------------
// Everything that can stays in two 16-bit lanes, and the 8-bit channels are kept SCALED BY 64 with their
// low six bits cleared (c6 = clamp(value, 0, 16383) & ~63 == channel << 6): that drops the per-channel
// shift, and all derived quantities stay exact integers --
//   V6 = max6,  d6 = max6 - min6,  S = (lut255[V] * d6) >> 14,  H = ((off * 64 + lut43[d] * diff6) >> 14) & 255.
// The hue LUT entry carries {t, -16384 * t} so that the +16384 bias which keeps the packed channel
// differences positive (one IADD3 each, no lane borrow) costs nothing to remove.
struct HueLutEntry { int32_t t; int32_t tBias; };            // lut43[d], -16384 * lut43[d]

__device__ __forceinline__ void fill_hue_lut(HueLutEntry* lut)
{
  for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x)
  {
    const int32_t t = i ? (int32_t)(11008u / i) : 0;
    lut[i].t = t;
    lut[i].tBias = -16384 * t;
  }
}

// 64-scaled channel lanes from key lanes: clamp to [0, 16383] in the value domain, clear bias and low 6 bits
__device__ __forceinline__ uint32_t chan6_from_key(uint32_t k)
{
  return __vminu2(__vmaxu2(k, 0x80008000u), 0xBFFFBFFFu) & 0x3FC03FC0u;
}

// sx (S = (sx >> 14) & 255) of both pixels from the packed max6 / d6 lanes
__device__ __forceinline__ void sat_scaled(uint32_t mx6, uint32_t d6, const uint16_t* __restrict__ lut255,
                                           uint32_t& sx0, uint32_t& sx1)
{
  const uint16_t* base = lut255;
  const uint32_t t0 = *reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(base) + ((mx6 & 0xFFFFu) >> 5));
  const uint32_t t1 = *reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(base) + (mx6 >> 21));
  sx0 = t0 * (d6 & 0xFFFFu);
  sx1 = t1 * (d6 >> 16);
}

// hx (H = (hx >> 14) & 255) of both pixels.  Sector choice as the reference's _cmpeq2 (:230-242):
// max == G -> green (off 21845, B - R); else max == B -> blue (43690, R - G); else red (0, G - B).
__device__ __forceinline__ void hue_scaled(uint32_t r6, uint32_t g6, uint32_t b6, uint32_t mx6, uint32_t d6,
                                           const HueLutEntry* __restrict__ lutHue, int32_t& hx0, int32_t& hx1)
{
  // lane masks 0xFFFF where max != G / max != B
  const uint32_t notG = __vminu2(mx6 ^ g6, 0x00010001u) * 0xFFFFu;
  const uint32_t notB = __vminu2(mx6 ^ b6, 0x00010001u) * 0xFFFFu;
  // biased differences, lanes in [64, 32704]
  const uint32_t dG = b6 + 0x40004000u - r6;
  const uint32_t dB = r6 + 0x40004000u - g6;
  const uint32_t dR = g6 + 0x40004000u - b6;
  const uint32_t dsel = (~notG & dG) | (notG & ((~notB & dB) | (notB & dR)));
  const uint32_t osel = (~notG & 0x55555555u) | (notG & ~notB & 0xAAAAAAAAu);       // 21845 / 43690 / 0 per lane
  const HueLutEntry e0 = *reinterpret_cast<const HueLutEntry*>(reinterpret_cast<const uint8_t*>(lutHue) + ((d6 & 0xFFFFu) >> 3));
  const HueLutEntry e1 = *reinterpret_cast<const HueLutEntry*>(reinterpret_cast<const uint8_t*>(lutHue) + (d6 >> 19));
  hx0 = e0.t * (int32_t)(dsel & 0xFFFFu) + ((int32_t)(osel & 0xFFFFu) * 64 + e0.tBias);
  hx1 = e1.t * (int32_t)(dsel >> 16) + ((int32_t)(osel >> 16) * 64 + e1.tBias);
}

// HSV (0x00VVSSHH) of both pixels of a pair.
__device__ __forceinline__ void hsv_pair(uint32_t yy, uint32_t cw, const ChromaCoef coef,
                                         const HueLutEntry* __restrict__ lutHue, const uint16_t* __restrict__ lut255,
                                         uint32_t& hsv0, uint32_t& hsv1)
{
  uint32_t kr, kg, kb;
  rgb_keys(yy, cw, coef, kr, kg, kb);
  const uint32_t r6 = chan6_from_key(kr), g6 = chan6_from_key(kg), b6 = chan6_from_key(kb);
  const uint32_t mx6 = __vimax3_u16x2(r6, g6, b6), mn6 = __vimin3_u16x2(r6, g6, b6);
  const uint32_t d6 = mx6 - mn6;
  uint32_t sx0, sx1;
  int32_t hx0, hx1;
  sat_scaled(mx6, d6, lut255, sx0, sx1);
  hue_scaled(r6, g6, b6, mx6, d6, lutHue, hx0, hx1);
  hsv0 = ((mx6 & 0xFFFFu) << 10) | ((sx0 >> 6) & 0xFF00u) | (((uint32_t)hx0 >> 14) & 0xFFu);
  hsv1 = ((mx6 >> 16) << 10) | ((sx1 >> 6) & 0xFF00u) | (((uint32_t)hx1 >> 14) & 0xFFu);
}

// ---- threshold of a pixel pair with warp-level early out --------------------------------------
// det = ((hsv <u4 from) | (hsv >u4 to)) == expected with expected in {0,1}: S and V must always be inside
// their bounds, only the hue test can be inverted.  So V (a max3 on the keys) and S (one LUT fetch per
// pixel) are tested first, and the expensive hue is only computed when some lane of the warp still has a
// candidate -- on frames where the wanted colour is a small part of the picture most warps stop after the
// cheap tests.  All comparisons are made on the scaled forms (no shifts to bytes).  Bit 0 / 1 = pixel 0 / 1.
struct HsvBounds {                     // loop-invariant, derived once per thread from FrameParams.from/to
  uint32_t vAdd2, vSub2;               // guard-bit constants for the packed V test on 64-scaled lanes
  uint32_t sLo, sSpan;                 // sx - sLo <=u sSpan
  uint32_t hLo, hSpan;                 // ((hx - hLo) & 0x3FFFFF) <=u hSpan
  uint32_t hValid;                     // hue interval non-empty
};

__device__ __forceinline__ HsvBounds make_bounds(uint32_t from, uint32_t to)
{
  HsvBounds b;
  const uint32_t hf = from & 0xFFu, ht = to & 0xFFu, sf = (from >> 8) & 0xFFu, st = (to >> 8) & 0xFFu;
  const uint32_t vf = (from >> 16) & 0xFFu, vt = (to >> 16) & 0xFFu;
  // x6 >= vf*64  <=> bit 15 of (x6 + 0x8000 - vf*64);  x6 > vt*64 <=> bit 15 of (x6 + 0x7FFF - vt*64)
  b.vAdd2 = (0x8000u - vf * 64u) * 0x10001u;
  b.vSub2 = (0x7FFFu - vt * 64u) * 0x10001u;
  b.sLo = sf << 14;
  b.sSpan = st >= sf ? (((st - sf + 1u) << 14) - 1u) : 0u;
  if (st < sf) b.sLo = 0xFFFFFFFFu;                    // sx < 2^30, so sx - sLo wraps above any span: never inside
  b.hLo = hf << 14;
  b.hSpan = ht >= hf ? (((ht - hf + 1u) << 14) - 1u) : 0u;
  b.hValid = ht >= hf ? 1u : 0u;
  return b;
}

__device__ __forceinline__ uint32_t detect_pair_bits(uint32_t yy, uint32_t cw, const ChromaCoef coef,
                                                     const HueLutEntry* __restrict__ lutHue,
                                                     const uint16_t* __restrict__ lut255,
                                                     const HsvBounds& bd, uint32_t expected)
{
  uint32_t kr, kg, kb;
  rgb_keys(yy, cw, coef, kr, kg, kb);
  const uint32_t mx6 = chan6_from_key(__vimax3_u16x2(kr, kg, kb));     // V*64 of both pixels
  const uint32_t vIn = (mx6 + bd.vAdd2) & ~(mx6 + bd.vSub2) & 0x80008000u;
  const unsigned am = __activemask();
  if (!__any_sync(am, vIn != 0u))
    return 0u;
  const uint32_t mn6 = chan6_from_key(__vimin3_u16x2(kr, kg, kb));
  const uint32_t d6 = mx6 - mn6;                                       // lanes >= 0, no borrow
  uint32_t sx0, sx1;
  sat_scaled(mx6, d6, lut255, sx0, sx1);
  uint32_t cand = 0u;
  if ((vIn & 0x8000u) && (sx0 - bd.sLo) <= bd.sSpan) cand |= 1u;
  if ((vIn & 0x80000000u) && (sx1 - bd.sLo) <= bd.sSpan) cand |= 2u;
  if (!__any_sync(am, cand != 0u))
    return 0u;
  const uint32_t r6 = chan6_from_key(kr), g6 = chan6_from_key(kg), b6 = chan6_from_key(kb);
  int32_t hx0, hx1;
  hue_scaled(r6, g6, b6, mx6, d6, lutHue, hx0, hx1);
  const uint32_t in0 = (bd.hValid && (((uint32_t)hx0 - bd.hLo) & 0x3FFFFFu) <= bd.hSpan) ? 1u : 0u;
  const uint32_t in1 = (bd.hValid && (((uint32_t)hx1 - bd.hLo) & 0x3FFFFFu) <= bd.hSpan) ? 1u : 0u;
  // hout == expected: expected 0 -> hue inside, expected 1 -> hue outside
  return cand & (((in0 ^ expected) & 1u) | (((in1 ^ expected) & 1u) << 1));
}

// ---- 128-bit streaming load ------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream(const void* p)
{
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// ---- cp.async (LDGSTS): 16 bytes global -> shared, L2 only, no register staging ---------------
__device__ __forceinline__ void cp_async16(void* smemDst, const void* gmemSrc)
{
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmemSrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

} // namespace trikb200
