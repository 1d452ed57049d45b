// trik_line.cuh -- device helpers shared by the sum / line kernels (trik_kernels.cu, trik_kernels_line.cu):
// the per-pair V threshold in the key domain and the result tails.
#pragma once
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"

namespace trikb200 {

// ---------------------------------------------------------------------------------------------
// finalisation: raw sums -> OutArgs, exactly the integer / float steps of the reference tails
// ---------------------------------------------------------------------------------------------
template <int KIND>
__device__ void finalize_sum(const Geometry& g, const FrameParams& p, uint32_t fails, uint32_t sxFail,
                             uint32_t syFail, uint32_t crossFail, TargetOut* o, const TargetOut* outBase)
{
  const uint32_t W = (uint32_t)g.width, H = (uint32_t)g.height;
  TargetOut r;
  DrawInfo di;
  di.v[0] = 0; di.v[1] = 0; di.v[2] = 0; di.v[3] = 0;
  r.targetX = 0; r.targetY = 0; r.targetSize = 0; r.pad = 0;
  r.detectHue = r.detectHueTolerance = r.detectSat = r.detectSatTolerance = r.detectVal = r.detectValTolerance = 0;
  if (KIND == KIND_WO)
  {
    const uint32_t points = W * H - fails;
    const uint32_t sx = H * (W * (W - 1u) / 2u) - sxFail;
    const uint32_t sy = W * (H * (H - 1u) / 2u) - syFail;
    if (points > 0u)
    {
      const int32_t tx = (int32_t)(sx / points);
      const int32_t ty = (int32_t)(sy / points);
      const uint32_t radius = (uint32_t)ceilf(sqrtf((float)points / 3.1415927f));
      r.targetX = (int8_t)(((tx - (int32_t)W / 2) * 100 * 2) / (int32_t)W);
      r.targetY = (int8_t)(((ty - (int32_t)H / 2) * 100 * 2) / (int32_t)H);
      r.targetSize = (uint8_t)((uint32_t)(radius * 100u * 4u) / (uint32_t)(W + H));
      di.v[0] = 1; di.v[1] = tx; di.v[2] = ty; di.v[3] = (int32_t)radius;
    }
  }
  else
  {
    const bool ol = (KIND == KIND_OL);
    const uint32_t winW = ol ? (W - 9u) : W;                                   // OL counts columns 5..W-5 (:288)
    const uint32_t colSum = ol ? ((W - 5u) * (W - 4u) / 2u - 10u) : (W * (W - 1u) / 2u);
    const uint32_t points = winW * H - fails;
    const uint32_t sx = H * colSum - sxFail;
    uint32_t cross = 0;
    if (ol)
    {
      uint32_t nrows = 0;
      if (p.hStart <= p.hStop && p.hStart < H)
        nrows = (p.hStop < H - 1u ? p.hStop : H - 1u) - p.hStart + 1u;
      cross = winW * nrows - crossFail;
    }
    if (points > 10u)
    {
      const int32_t tx = (int32_t)(sx / points);
      r.targetX = (int8_t)(((tx - (int32_t)W / 2) * 100 * 2) / (int32_t)W);
      if (ol)
        r.targetY = (int8_t)(int32_t)((uint32_t)(cross * 100u) / (uint32_t)(W * 2u * 40u));
      r.targetSize = (uint8_t)((uint32_t)(points * 100u) / (uint32_t)(H * W));
      di.v[0] = 1; di.v[1] = tx;
    }
  }
  *o = r;
  if (g.drawInfo)
  {
    int32_t* dv = static_cast<DrawInfo*>(g.drawInfo)[o - outBase].v;
    dv[0] = di.v[0]; dv[1] = di.v[1]; dv[2] = di.v[2]; dv[3] = di.v[3];
  }
}

// {N, N+1} lanes (pass, fail) of a pixel pair from its key lanes.
// IMAD / IDP.4A issue on the FMA pipe and LOP3 / PRMT / VIADD / VIMNMX on the ALU pipe, both at one warp
// instruction per two cycles per scheduler, so the pair is split 7 : 6 between them:
//   FMA: 74*yy, three IDP.4A chroma terms, two replicate-and-add IMADs (red, green), one replicate IMAD (blue)
//   ALU: luma extraction, VIADD.16x2 (blue, wraps like the reference's int16), VIMNMX3, VIADDMNMX, VIMNMX, VIADD.16x2 (count)
__device__ __forceinline__ uint32_t vtest_lanes(uint32_t yy, uint32_t cw, const ChromaCoef cf,
                                                uint32_t negKlo2, uint32_t n2, uint32_t cap2)
{
  const uint32_t y74 = yy * 74u;
  const uint32_t cr = dp4a_uu(cw, cf.r, KEY_BIAS - 14248u);
  const uint32_t cg = (uint32_t)dp4a_us(cw, cf.g, (int32_t)(KEY_BIAS + 8696u));
  const uint32_t cb = dp4a_uu(cw, cf.b, KEY_BIAS - 17672u);
  const uint32_t kr = cr * 0x10001u + y74;
  const uint32_t kg = cg * 0x10001u + y74;
  const uint32_t kb = __vadd2(cb * 0x10001u, y74);
  const uint32_t km = __vimax3_u16x2(kr, kg, kb);
  return __vminu2(__vmaxu2(__vadd2(km, negKlo2), n2), cap2);
}

// YUYV: measured slightly faster with the luma term folded into the dot product -- one IDP.4A per pixel
// and channel yields the finished 32-bit key (74*Y + 102*V + c, ...), two keys are packed into lanes by
// one IMAD (red, green: cannot overflow a lane) or one PRMT (blue: truncation to 16 bits IS the wrap).
// 8 FMA-pipe + 5 ALU-pipe instructions per pair.
__device__ __forceinline__ uint32_t vtest_yuyv(uint32_t w, uint32_t negKlo2, uint32_t n2, uint32_t cap2)
{
  const uint32_t r0 = dp4a_uu(w, 0x6600004Au, KEY_BIAS - 14248u);                       // 74*Y0 + 102*V + c
  const uint32_t r1 = dp4a_uu(w, 0x664A0000u, KEY_BIAS - 14248u);                       // 74*Y1 + 102*V + c
  const uint32_t g0 = (uint32_t)dp4a_us(w, 0xCC00E74Au, (int32_t)(KEY_BIAS + 8696u));   // 74*Y0 - 25*U - 52*V + c
  const uint32_t g1 = (uint32_t)dp4a_us(w, 0xCC4AE700u, (int32_t)(KEY_BIAS + 8696u));
  const uint32_t b0 = dp4a_uu(w, 0x0000814Au, KEY_BIAS - 17672u);                       // 74*Y0 + 129*U + c (may exceed 16 bits)
  const uint32_t b1 = dp4a_uu(w, 0x004A8100u, KEY_BIAS - 17672u);
  const uint32_t kr = r1 * 65536u + r0;
  const uint32_t kg = g1 * 65536u + g0;
  const uint32_t kb = __byte_perm(b0, b1, 0x5410);
  const uint32_t km = __vimax3_u16x2(kr, kg, kb);
  return __vminu2(__vmaxu2(__vadd2(km, negKlo2), n2), cap2);
}

// pair J (0/1) of a luma word L and the chroma word Cw under it
template <int J>
__device__ __forceinline__ uint32_t vtest_planar(uint32_t L, uint32_t Cw, uint32_t negKlo2, uint32_t n2, uint32_t cap2)
{
  return vtest_lanes(__byte_perm(L, 0u, J ? 0x4342 : 0x4140), Cw, J ? coef_planar1() : coef_planar0(), negKlo2, n2, cap2);
}

__device__ __forceinline__ uint32_t lanes_sub(uint32_t a, uint32_t b)      // lane-wise (a - b) mod 2^16
{
  return __vadd2(a, __vadd2(~b, 0x00010001u));
}
__device__ __forceinline__ uint32_t lanes_total(uint32_t a) { return (a & 0xFFFFu) + (a >> 16); }


} // namespace trikb200
