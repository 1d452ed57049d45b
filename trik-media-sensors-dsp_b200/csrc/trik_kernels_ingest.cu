// trik_kernels_ingest.cu -- ingest front end: packed RGB565 camera frames -> the YUV422P layout
// the ov7670 sensors read (luma plane, then at lineLength * H a plane of interleaved chroma bytes V U V U ...,
// ov7670/object_sensor/include/internal/cv_ball_detector_seqpass.hpp:343-373).
//
// The reference has NO RGB565 input path: its ov7670 sensors accept YUV422P only (src/vidtranscode_cv.cpp:76-84) and
// RGB565X is the format of the preview image they WRITE (writeOutputPixel, webcam/object_sensor/include/internal/
// cv_ball_detector_seqpass.hpp:66-70).  So this conversion is defined here, not by the reference (parity unpinned: there
// is nothing to pin it to); it is the usual integer BT.601 studio-swing matrix, i.e. the inverse of the matrix the
// reference's own YUV -> RGB step uses (:181-205: 74/64 = 1.156 on Y-16, 102/64 = 1.594 on V-128, ...):
//     R8 = r5 << 3 | r5 >> 2,  G8 = g6 << 2 | g6 >> 4,  B8 = b5 << 3 | b5 >> 2
//     Y  = (( 66 R + 129 G +  25 B + 128) >> 8) + 16
//     chroma of a pixel pair from its summed channels Rs = R0 + R1, Gs, Bs (arithmetic shift = floor):
//     Cb = ((-38 Rs -  74 Gs + 112 Bs + 256) >> 9) + 128
//     Cr = ((112 Rs -  94 Gs -  18 Bs + 256) >> 9) + 128
// One thread converts 8 pixels: one 16-byte load, one 8-byte luma store, one 8-byte chroma store; HBM bound
// (2 bytes read + 2 bytes written per pixel).
#include <atomic>
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"

namespace trikb200 {

extern std::atomic<long long> g_launches_ingest;
std::atomic<long long> g_launches_ingest{0};

__device__ __forceinline__ int32_t dp2a_lo_us(uint32_t a_u16x2, uint32_t b_s8x4, int32_t c)
{
  int32_t d;
  asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u16x2), "r"(b_s8x4), "r"(c));
  return d;
}

// One pixel pair (two RGB565 words in one register) at a time, the two pixels in the two 16-bit lanes: no lane of
// any intermediate exceeds 16 bits (luma sum <= 60 324), so packed IMADs do the luma of both pixels at once and a
// two-way dot product (IDP.2A) sums a chroma row over the pair.  yy: 0xY1__Y0__ (bytes 3 and 1), vu = V | U << 8.
template <bool BGR>
__device__ __forceinline__ void pair_to_ycc(uint32_t w, uint32_t& yy, uint32_t& vu)
{
  const uint32_t hi5 = (w >> 11) & 0x001F001Fu, g6 = (w >> 5) & 0x003F003Fu, lo5 = w & 0x001F001Fu;
  const uint32_t r5 = BGR ? lo5 : hi5, b5 = BGR ? hi5 : lo5;
  const uint32_t R = ((r5 * 33u) >> 2) & 0x00FF00FFu;      // r5 << 3 | r5 >> 2
  const uint32_t G = ((g6 * 65u) >> 4) & 0x00FF00FFu;      // g6 << 2 | g6 >> 4
  const uint32_t B = ((b5 * 33u) >> 2) & 0x00FF00FFu;
  yy = R * 66u + G * 129u + B * 25u + 0x10801080u;         // + 128 + (16 << 8) per lane; Y = bits 15..8 of each lane
  // chroma of the pair from its summed RGB, biased by 128 << 9 so that the shift floors a positive number
  const int32_t cb = dp2a_lo_us(B, 0x7070u, dp2a_lo_us(G, 0xB6B6u, dp2a_lo_us(R, 0xDADAu, 256 + (128 << 9))));   // -38 -74 112
  const int32_t cr = dp2a_lo_us(B, 0xEEEEu, dp2a_lo_us(G, 0xA2A2u, dp2a_lo_us(R, 0x7070u, 256 + (128 << 9))));   // 112 -94 -18
  vu = ((uint32_t)cr >> 9) | (((uint32_t)cb >> 9) << 8);
}

template <bool BGR>
__global__ void __launch_bounds__(256)
ingest_rgb565_kernel(const uint8_t* __restrict__ src, const long long srcStride, const int srcLine,
                     uint8_t* __restrict__ dst, const long long dstStride, const int dstLine,
                     const int width, const int height, const int numFrames, const uint32_t cprMagic)
{
  const int cpr = width >> 3;                              // 8-pixel chunks per row
  const uint32_t rem = blockIdx.x * blockDim.x + threadIdx.x;
  if (rem >= (uint32_t)(cpr * height))
    return;
  const int r = (int)__umulhi(rem, cprMagic), c = (int)rem - r * cpr;       // rem / cpr, exact for rem * cpr < 2^32
  for (int frame = blockIdx.y; frame < numFrames; frame += gridDim.y)
  {
    const uint4 in = ld_stream(src + (long long)frame * srcStride + (size_t)r * srcLine + (size_t)c * 16);
    uint32_t y0, y1, y2, y3, c0, c1, c2, c3;
    pair_to_ycc<BGR>(in.x, y0, c0);
    pair_to_ycc<BGR>(in.y, y1, c1);
    pair_to_ycc<BGR>(in.z, y2, c2);
    pair_to_ycc<BGR>(in.w, y3, c3);
    uint8_t* out = dst + (long long)frame * dstStride + (size_t)r * dstLine + (size_t)c * 8;
    *reinterpret_cast<uint2*>(out) = make_uint2(__byte_perm(y0, y1, 0x7531), __byte_perm(y2, y3, 0x7531));
    *reinterpret_cast<uint2*>(out + (size_t)dstLine * height) = make_uint2(c0 | (c1 << 16), c2 | (c3 << 16));   // V first
  }
}

cudaError_t launch_ingest_rgb565(const uint8_t* src, long long srcStride, int srcLine, uint8_t* dst, long long dstStride,
                                 int dstLine, int width, int height, int numFrames, int bgr, int smCount, cudaStream_t stream)
{
  if (numFrames <= 0)
    return cudaSuccess;
  if (width <= 0 || height <= 0 || (width & 7) || (srcLine & 15) || (dstLine & 7) || (srcStride & 15) || (dstStride & 7)
      || srcLine < 2 * width || dstLine < width)
    return cudaErrorInvalidValue;
  const long long perFrame = (long long)(width >> 3) * height;
  if (perFrame * (width >> 3) >= (1ll << 32))
    return cudaErrorInvalidValue;
  const uint32_t cprMagic = (uint32_t)((1ull << 32) / (uint32_t)(width >> 3)) + 1u;
  (void)smCount;
  const dim3 grid((unsigned)((perFrame + 255) / 256), (unsigned)(numFrames < 65535 ? numFrames : 65535));
  if (bgr)
    ingest_rgb565_kernel<true><<<grid, 256, 0, stream>>>(src, srcStride, srcLine, dst, dstStride, dstLine, width, height, numFrames, cprMagic);
  else
    ingest_rgb565_kernel<false><<<grid, 256, 0, stream>>>(src, srcStride, srcLine, dst, dstStride, dstLine, width, height, numFrames, cprMagic);
  ++g_launches_ingest;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Scattered pinned host frames -> the handle's device batch buffer, in ONE launch.
// trikb200_processMixed hands over one pointer per frame (the reference's model is one codec instance per camera, so a time
// step of 1024 streams is 1024 separate buffers).  One cudaMemcpyAsync per frame costs about 2 us of host time each, more
// than the 2.8 us the 153.6 KB take on the PCIe link; this kernel reads the pinned buffers in place instead (the host
// pointers are device pointers under unified addressing) with 64 bytes per thread in flight, and writes the frames at
// i * dstStride as the sensor kernels expect them.  No arithmetic: bound by the PCIe read bandwidth.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_frames_kernel(const uint8_t* const* __restrict__ srcPtrs, uint8_t* __restrict__ dst, const long long dstStride,
                     const uint32_t chunks /* 16-byte units per frame */, const int numFrames)
{
  for (int frame = blockIdx.y; frame < numFrames; frame += gridDim.y)
  {
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(srcPtrs[frame]);
    uint4* __restrict__ out = reinterpret_cast<uint4*>(dst + (long long)frame * dstStride);
    const uint32_t step = gridDim.x * blockDim.x;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3u * step < chunks; i += 4u * step)
    {
      const uint4 a = ld_stream(reinterpret_cast<const uint8_t*>(src + i));
      const uint4 b = ld_stream(reinterpret_cast<const uint8_t*>(src + i + step));
      const uint4 c = ld_stream(reinterpret_cast<const uint8_t*>(src + i + 2u * step));
      const uint4 d = ld_stream(reinterpret_cast<const uint8_t*>(src + i + 3u * step));
      out[i] = a; out[i + step] = b; out[i + 2u * step] = c; out[i + 3u * step] = d;
    }
    for (; i < chunks; i += step)
      out[i] = ld_stream(reinterpret_cast<const uint8_t*>(src + i));
  }
}

cudaError_t launch_gather_frames(const uint8_t* const* dSrcPtrs, uint8_t* dst, long long dstStride, size_t frameBytes,
                                 int numFrames, cudaStream_t stream)
{
  if (numFrames <= 0)
    return cudaSuccess;
  if ((frameBytes & 15) || (dstStride & 15) || frameBytes / 16 >= (1ull << 30))
    return cudaErrorInvalidValue;
  const uint32_t chunks = (uint32_t)(frameBytes / 16);
  // about 4 KB per CTA pass: enough CTAs in flight to cover the link latency even for a handful of frames
  unsigned perFrame = (chunks + 1023u) / 1024u;
  if (perFrame < 1u) perFrame = 1u;
  if (perFrame > 8u) perFrame = 8u;
  const dim3 grid(perFrame, (unsigned)(numFrames < 65535 ? numFrames : 65535));
  gather_frames_kernel<<<grid, 256, 0, stream>>>(dSrcPtrs, dst, dstStride, chunks, numFrames);
  ++g_launches_ingest;
  return cudaGetLastError();
}

} // namespace trikb200
