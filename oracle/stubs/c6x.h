/*
 * TEST INFRASTRUCTURE (oracle/): host emulation of the TI C674x packed-SIMD
 * intrinsics that the reference's hot loops use (<c6x.h> is a TI compiler builtin
 * header, not vendored in the reference).  Call sites: e.g.
 * trik/webcam/object_sensor/include/internal/cv_ball_detector_seqpass.hpp:171-249.
 *
 * Semantics follow TI's "TMS320C6000 Optimizing Compiler" intrinsic tables
 * (SURVEY.md Appendix A).  They are the root of trust of every parity claim in
 * this repository and are not pinned by any test inside the reference.
 *
 * Written as branch-free scalar bit manipulation (no per-lane loops) so that the
 * host build of the reference is a fair CPU baseline: gcc -O2 turns each of these
 * into a handful of integer instructions.  Notation: b0..b3 bytes (b0 least
 * significant), lo/hi 16-bit halves.
 */
#ifndef ORACLE_STUB_C6X_H_
#define ORACLE_STUB_C6X_H_

#include <stdint.h>

#ifndef restrict
#define restrict __restrict__
#endif

#define C6X_INLINE static inline __attribute__((always_inline))

C6X_INLINE uint32_t _loll(uint64_t x) { return (uint32_t)x; }
C6X_INLINE uint32_t _hill(uint64_t x) { return (uint32_t)(x >> 32); }
C6X_INLINE uint64_t _itoll(uint32_t h, uint32_t l) { return ((uint64_t)h << 32) | (uint64_t)l; }

C6X_INLINE uint32_t _pack2(uint32_t a, uint32_t b)   { return (a << 16) | (b & 0xffffu); }
C6X_INLINE uint32_t _packh2(uint32_t a, uint32_t b)  { return (a & 0xffff0000u) | (b >> 16); }
C6X_INLINE uint32_t _packlh2(uint32_t a, uint32_t b) { return (a << 16) | (b >> 16); }
C6X_INLINE uint32_t _packhl2(uint32_t a, uint32_t b) { return (a & 0xffff0000u) | (b & 0xffffu); }

/* bytes [a.b3, a.b1, b.b3, b.b1] from most to least significant */
C6X_INLINE uint32_t _packh4(uint32_t a, uint32_t b)
{
  return (a & 0xff000000u) | ((a & 0x0000ff00u) << 8) | ((b & 0xff000000u) >> 16) | ((b & 0x0000ff00u) >> 8);
}

/* two independent 16-bit adds, each wrapping modulo 2^16 */
C6X_INLINE uint32_t _add2(uint32_t a, uint32_t b)
{
  return ((a & 0xffff0000u) + (b & 0xffff0000u)) | ((a + b) & 0xffffu);
}

/* each half treated as signed 16-bit, arithmetic shift right */
C6X_INLINE uint32_t _shr2(uint32_t a, uint32_t n)
{
  const uint32_t hi = (uint32_t)((int32_t)a >> n) & 0xffff0000u;
  const uint32_t lo = (uint32_t)((int32_t)(int16_t)(a & 0xffffu) >> n) & 0xffffu;
  return hi | lo;
}

/* clear bits lo..hi inclusive */
C6X_INLINE uint32_t _clr(uint32_t a, uint32_t lo, uint32_t hi)
{
  const uint32_t width = hi - lo + 1u;
  const uint32_t mask = (width >= 32u ? 0xffffffffu : ((1u << width) - 1u)) << lo;
  return a & ~mask;
}

C6X_INLINE uint32_t c6x_sat_s16_to_u8(int32_t v) { return v < 0 ? 0u : (v > 255 ? 255u : (uint32_t)v); }

/* bytes [sat(a.hi), sat(a.lo), sat(b.hi), sat(b.lo)], sat = signed 16 -> unsigned 8 */
C6X_INLINE uint32_t _spacku4(uint32_t a, uint32_t b)
{
  return (c6x_sat_s16_to_u8((int16_t)(a >> 16)) << 24)
       | (c6x_sat_s16_to_u8((int16_t)(a & 0xffffu)) << 16)
       | (c6x_sat_s16_to_u8((int16_t)(b >> 16)) << 8)
       |  c6x_sat_s16_to_u8((int16_t)(b & 0xffffu));
}

/* four unsigned 8x8 products, product of byte i in bits 16i..16i+15 */
C6X_INLINE uint64_t _mpyu4ll(uint32_t a, uint32_t b)
{
  const uint64_t p0 = (uint64_t)((a      ) & 0xffu) * ((b      ) & 0xffu);
  const uint64_t p1 = (uint64_t)((a >>  8) & 0xffu) * ((b >>  8) & 0xffu);
  const uint64_t p2 = (uint64_t)((a >> 16) & 0xffu) * ((b >> 16) & 0xffu);
  const uint64_t p3 = (uint64_t)((a >> 24)        ) * ((b >> 24)        );
  return p0 | (p1 << 16) | (p2 << 32) | (p3 << 48);
}

/* sum_i unsigned byte i of u  x  signed byte i of s */
C6X_INLINE int32_t _dotpus4(uint32_t u, uint32_t s)
{
  return (int32_t)((u      ) & 0xffu) * (int32_t)(int8_t)(s      )
       + (int32_t)((u >>  8) & 0xffu) * (int32_t)(int8_t)(s >>  8)
       + (int32_t)((u >> 16) & 0xffu) * (int32_t)(int8_t)(s >> 16)
       + (int32_t)((u >> 24)        ) * (int32_t)(int8_t)(s >> 24);
}

/* a.hi*b.hi - a.lo*b.lo, halves signed 16-bit */
C6X_INLINE int32_t _dotpn2(uint32_t a, uint32_t b)
{
  return (int32_t)(int16_t)(a >> 16) * (int32_t)(int16_t)(b >> 16)
       - (int32_t)(int16_t)(a & 0xffffu) * (int32_t)(int16_t)(b & 0xffffu);
}

C6X_INLINE uint32_t _cmpeq2(uint32_t a, uint32_t b)
{
  const uint32_t x = a ^ b;
  return (((x >> 16) == 0u) ? 2u : 0u) | (((x & 0xffffu) == 0u) ? 1u : 0u);
}

C6X_INLINE uint32_t _cmpgtu4(uint32_t a, uint32_t b)
{
  return ((( a        & 0xffu) > ( b        & 0xffu)) ? 1u : 0u)
       | ((((a >>  8) & 0xffu) > ((b >>  8) & 0xffu)) ? 2u : 0u)
       | ((((a >> 16) & 0xffu) > ((b >> 16) & 0xffu)) ? 4u : 0u)
       | ((( a >> 24         ) > ( b >> 24         )) ? 8u : 0u);
}

C6X_INLINE uint32_t _cmpltu4(uint32_t a, uint32_t b) { return _cmpgtu4(b, a); }

C6X_INLINE uint32_t _maxu4(uint32_t a, uint32_t b)
{
  uint32_t r = 0;
  uint32_t x, y;
  x = a & 0x000000ffu; y = b & 0x000000ffu; r |= x > y ? x : y;
  x = a & 0x0000ff00u; y = b & 0x0000ff00u; r |= x > y ? x : y;
  x = a & 0x00ff0000u; y = b & 0x00ff0000u; r |= x > y ? x : y;
  x = a & 0xff000000u; y = b & 0xff000000u; r |= x > y ? x : y;
  return r;
}

C6X_INLINE uint32_t _minu4(uint32_t a, uint32_t b)
{
  uint32_t r = 0;
  uint32_t x, y;
  x = a & 0x000000ffu; y = b & 0x000000ffu; r |= x < y ? x : y;
  x = a & 0x0000ff00u; y = b & 0x0000ff00u; r |= x < y ? x : y;
  x = a & 0x00ff0000u; y = b & 0x00ff0000u; r |= x < y ? x : y;
  x = a & 0xff000000u; y = b & 0xff000000u; r |= x < y ? x : y;
  return r;
}

/* (b3 << 16) + b2 */
C6X_INLINE uint32_t _unpkhu4(uint32_t a) { return ((a >> 8) & 0x00ff0000u) | ((a >> 16) & 0xffu); }
/* (b1 << 16) + b0 */
C6X_INLINE uint32_t _unpklu4(uint32_t a) { return ((a << 8) & 0x00ff0000u) | (a & 0xffu); }
/* swap the bytes inside each half-word */
C6X_INLINE uint32_t _swap4(uint32_t a) { return ((a & 0x00ff00ffu) << 8) | ((a >> 8) & 0x00ff00ffu); }

#endif /* ORACLE_STUB_C6X_H_ */
