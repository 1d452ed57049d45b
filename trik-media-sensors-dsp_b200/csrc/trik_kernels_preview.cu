// trik_kernels_preview.cu -- the RGB565X preview image with overlays:
// the other half of process()'s surface.
//
// Reference behaviour (e.g. webcam/object_sensor/include/internal/cv_ball_detector_seqpass.hpp):
//   * the shim zero-fills the preview buffer                          (src/vidtranscode_cv_fxns.c:234)
//   * pass 2 visits every SOURCE pixel in raster order and writes it -- cyan 0x00ffff if detected, else
//     its RGB888 -- to dst[hi2ho[row]][wi2wo[col]] where both index maps are i * min(outW/W, outH/H)
//     truncated (:371-387, writeOutputPixel :66-70, loop :330-353).  When several source pixels map to one
//     destination pixel the LAST one in raster order wins; when the preview is larger than the frame the
//     pixels never written stay 0.
//   * then the overlays are drawn in a fixed order, later writes winning (:471-494).
//
// Here the base layer is a GATHER: because the maps are separable and monotonic, the last writer of
// dst(r', c') is src(max row with hi2ho == r', max col with wi2wo == c'); the host precomputes those
// inverse maps with the reference's own double arithmetic and every destination pixel is produced exactly
// once, coalesced.  Overlays are replayed per frame by one warp: primitives in reference order, the pixels
// of one primitive (one colour) spread over the lanes.
#include <atomic>
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"

namespace trikb200 {

extern std::atomic<long long> g_launches_preview;
std::atomic<long long> g_launches_preview{0};

__device__ __forceinline__ uint16_t rgb565x(uint32_t rgb888)           // writeOutputPixel (:66-70)
{
  return (uint16_t)(((rgb888 >> 19) & 0x001Fu) | ((rgb888 >> 5) & 0x07E0u) | ((rgb888 << 8) & 0xF800u));
}

// colour the reference's pass 2 writes for source pixel (row, col)
template <int KIND>
__device__ __forceinline__ uint32_t source_colour(const Geometry& g, const uint8_t* __restrict__ frame,
                                                  const FrameParams& p, const uint16_t* __restrict__ bitmap,
                                                  const uint16_t* lut43, const uint16_t* lut255, int row, int col)
{
  constexpr bool PLANAR = (KIND == KIND_OO || KIND == KIND_OL || KIND == KIND_OM);
  const int pr = col >> 1, e = col & 1;
  uint32_t yy, cw;
  ChromaCoef coef = coef_yuyv();
  if (!PLANAR)
  {
    cw = *reinterpret_cast<const uint32_t*>(frame + (size_t)row * g.lineLength + (size_t)pr * 4u);
    yy = cw & 0x00FF00FFu;
  }
  else
  {
    const uint32_t l2 = *reinterpret_cast<const uint16_t*>(frame + (size_t)row * g.lineLength + (size_t)pr * 2u);
    cw = *reinterpret_cast<const uint16_t*>(frame + (size_t)(g.height + row) * g.lineLength + (size_t)pr * 2u);
    yy = (l2 & 0xFFu) | ((l2 >> 8) << 16);
    coef = coef_planar0();
  }
  uint32_t kr, kg, kb;
  rgb_keys(yy, cw, coef, kr, kg, kb);
  const uint32_t r2 = chan8_from_key(kr), g2 = chan8_from_key(kg), b2 = chan8_from_key(kb);
  const uint32_t r = e ? r2 >> 16 : r2 & 0xFFFFu, gch = e ? g2 >> 16 : g2 & 0xFFFFu, b = e ? b2 >> 16 : b2 & 0xFFFFu;
  const uint32_t rgb = (r << 16) | (gch << 8) | b;
  bool det = false;
  if (KIND == KIND_WO)
  {
    const uint32_t hsv = hsv_from_rgb8((int32_t)r, (int32_t)gch, (int32_t)b, lut43, lut255);
    det = detect_hsv(hsv, p.from, p.to, p.expected);
  }
  else if (KIND == KIND_WL || KIND == KIND_OL)
  {
    const uint32_t v = max(r, max(gch, b));
    det = v >= ((p.from >> 16) & 0xFFu) && v <= ((p.to >> 16) & 0xFFu);
  }
  else if (KIND == KIND_OO)
  {
    // det = the pixel's metapixel carries a label, i.e. it is "on" (popcount > 2), cv_ball_detector_seqpass.hpp:411-415
    det = __popc((unsigned)bitmap[(size_t)(row >> 2) * (g.width >> 2) + (col >> 2)]) > 2;
  }
  return det ? 0x00FFFFu : rgb;
}

struct PreviewGeom {
  int outW, outH, outLine;
  const int32_t* lastRow;     // [outH]  last source row mapping to this destination row, -1 if none
  const int32_t* lastCol;     // [outW]
  const int32_t* hi2ho;       // [H]
  const int32_t* wi2wo;       // [W]
};

template <int KIND>
__global__ void __launch_bounds__(256)
preview_base_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
                    const int paramStride, const uint16_t* __restrict__ bitmaps, const PreviewGeom pg,
                    uint8_t* __restrict__ previews, const long long previewStride)
{
  __shared__ uint16_t s_lut43[256];
  __shared__ uint16_t s_lut255[256];
  if (KIND == KIND_WO)
  {
    fill_div_luts(s_lut43, s_lut255);
    __syncthreads();
  }
  const int frame = blockIdx.x;
  const int halfW = (pg.outW + 1) >> 1;
  const long long item = (long long)blockIdx.y * blockDim.x + threadIdx.x;
  if (item >= (long long)halfW * pg.outH)
    return;
  const int drow = (int)(item / halfW), dc0 = (int)(item % halfW) * 2;
  const FrameParams p = params[(size_t)frame * paramStride];
  const uint8_t* fr = frames + (size_t)frame * g.frameStride;
  const uint16_t* bm = (KIND == KIND_OO) ? bitmaps + (size_t)frame * (g.width >> 2) * (g.height >> 2) : nullptr;
  uint8_t* dst = previews + (size_t)frame * previewStride + (size_t)drow * pg.outLine + (size_t)dc0 * 2u;
  const int srow = pg.lastRow[drow];
#pragma unroll
  for (int k = 0; k < 2; ++k)
  {
    if (dc0 + k >= pg.outW)
      break;
    const int scol = pg.lastCol[dc0 + k];
    uint16_t px = 0;
    if (srow >= 0 && scol >= 0)
      px = rgb565x(source_colour<KIND>(g, fr, p, bm, s_lut43, s_lut255, srow, scol));
    *reinterpret_cast<uint16_t*>(dst + 2 * k) = px;
  }
}

// ---------------------------------------------------------------------------------------------
// overlays
// ---------------------------------------------------------------------------------------------
struct Canvas {
  uint8_t* img;
  int W, H, outLine;
  const int32_t* hi2ho;
  const int32_t* wi2wo;
  __device__ __forceinline__ void put(int col, int row, uint32_t rgb) const      // drawOutputPixelBound (:72-89)
  {
    col = col < 0 ? 0 : (col > W - 1 ? W - 1 : col);
    row = row < 0 ? 0 : (row > H - 1 ? H - 1 : row);
    *reinterpret_cast<uint16_t*>(img + (size_t)hi2ho[row] * outLine + (size_t)wi2wo[col] * 2u) = rgb565x(rgb);
  }
};

// vertical / horizontal "target" lines of +-100 pixels (WO :136-168), one colour: lanes share the work
__device__ void centre_line(const Canvas& c, int lane, int col, int row, uint32_t rgb)
{
  for (int adj = lane; adj < 100; adj += 32) { c.put(col, row - adj, rgb); c.put(col, row + adj, rgb); }
}
__device__ void horizontal_centre_line(const Canvas& c, int lane, int col, int row, uint32_t rgb)
{
  for (int adj = lane; adj < 100; adj += 32) { c.put(col - adj, row, rgb); c.put(col + adj, row, rgb); }
}
// Bresenham circle, sequential (WO :91-134)
__device__ void circle(const Canvas& c, int col, int row, int radius, uint32_t rgb)
{
  int err = 1 - radius, errY = 1, errX = -2 * radius, x = radius, y = 0;
  c.put(col, row + radius, rgb); c.put(col, row - radius, rgb); c.put(col + radius, row, rgb); c.put(col - radius, row, rgb);
  while (y < x)
  {
    if (err >= 0) { x -= 1; errX += 2; err += errX; }
    y += 1; errY += 2; err += errY;
    c.put(col + x, row + y, rgb); c.put(col + x, row - y, rgb); c.put(col - x, row + y, rgb); c.put(col - x, row - y, rgb);
    c.put(col + y, row + x, rgb); c.put(col + y, row - x, rgb); c.put(col - y, row + x, rgb); c.put(col - y, row - x, rgb);
  }
}

template <int KIND>
__global__ void __launch_bounds__(32)
preview_overlay_kernel(const Geometry g, const FrameParams* __restrict__ params, const int paramStride,
                       const DrawInfo* __restrict__ draw, const int32_t* __restrict__ omColours, const PreviewGeom pg,
                       uint8_t* __restrict__ previews, const long long previewStride)
{
  const int frame = blockIdx.x, lane = threadIdx.x;
  const int W = g.width, H = g.height;
  Canvas c{previews + (size_t)frame * previewStride, W, H, pg.outLine, pg.hi2ho, pg.wi2wo};
  const int hW = W / 2, hH = H / 2;
  if (KIND == KIND_WO || KIND == KIND_OO)
  {
    const int step = H / 6;                                                     // m_detectZoneScale = 6
    centre_line(c, lane, hW - 2 * step, hH, 0xFF00FFu); centre_line(c, lane, hW - step, hH, 0xFF00FFu);
    centre_line(c, lane, hW + step, hH, 0xFF00FFu);     centre_line(c, lane, hW + 2 * step, hH, 0xFF00FFu);
    horizontal_centre_line(c, lane, hW, hH - 2 * step, 0xFF00FFu); horizontal_centre_line(c, lane, hW, hH - step, 0xFF00FFu);
    horizontal_centre_line(c, lane, hW, hH + step, 0xFF00FFu);     horizontal_centre_line(c, lane, hW, hH + 2 * step, 0xFF00FFu);
    __syncwarp();
    const DrawInfo d = draw[frame];
    if (KIND == KIND_WO)
    {
      if (lane == 0 && d.v[0])
        circle(c, d.v[1], d.v[2], d.v[3], 0xFFFF00u);                           // WO :494
    }
    else
    {
      // drawFatPixel (OO :99-119) for every reported target, all red: order irrelevant
      for (int i = 0; i < 8; ++i)
        if ((d.v[0] >> i) & 1)
          if (lane < 9)
            c.put(d.v[1 + 2 * i] + lane / 3 - 1, d.v[2 + 2 * i] + lane % 3 - 1, 0xFF0000u);
    }
  }
  else if (KIND == KIND_WL || KIND == KIND_OL)
  {
    // drawRgbThinLine: full-height verticals at hW -+ 40, -+ 80 (WL :384-387); drawY = 0
    for (int adj = lane; adj < H; adj += 32)
    {
      c.put(hW - 40, adj, 0xFF00FFu); c.put(hW + 40, adj, 0xFF00FFu);
      c.put(hW - 80, adj, 0xFF00FFu); c.put(hW + 80, adj, 0xFF00FFu);
    }
    __syncwarp();
    if (KIND == KIND_OL)
    {
      for (int adj = lane; adj < W; adj += 32)                                  // drawRgbHorizontalLine (OL :454-455)
      {
        c.put(adj, hH, 0xFF0000u);
        c.put(adj, hH + 80, 0xFF0000u);
      }
      __syncwarp();
    }
    const DrawInfo d = draw[frame];
    if (d.v[0])                                                                 // points > 10: 3-wide red line (WL :410)
      for (int adj = lane; adj < H; adj += 32)
      {
        c.put(d.v[1] - 1, adj, 0xFF0000u); c.put(d.v[1], adj, 0xFF0000u); c.put(d.v[1] + 1, adj, 0xFF0000u);
      }
  }
  else if (KIND == KIND_OM)
  {
    // fillImage (OM :355-366): a 20x20 square of the cell's colour at every cell origin, cells in order
    const FrameParams p = params[(size_t)frame * paramStride];
    const int M = (int)p.gridRows, N = (int)p.gridCols;
    if (M <= 0 || N <= 0)
      return;
    const int ws = (uint16_t)(W / N), hs = (uint16_t)(H / M);
    const int32_t* colours = omColours + (size_t)frame * 100;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j)
      {
        const uint32_t rgb = (uint32_t)colours[i * N + j];
        const int r0 = (uint16_t)(i * hs), c0 = (uint16_t)(j * ws);
        for (int k = lane; k < 400; k += 32)
          c.put(c0 + k % 20, r0 + k / 20, rgb);
        __syncwarp();
      }
  }
}

cudaError_t launch_preview(int kind, const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                           int paramStride, const uint16_t* bitmaps, const DrawInfo* draw, const int32_t* omColours,
                           int outW, int outH, int outLine, const int32_t* lastRow, const int32_t* lastCol,
                           const int32_t* hi2ho, const int32_t* wi2wo, uint8_t* previews, long long previewStride,
                           cudaStream_t stream)
{
  if (numFrames <= 0 || outW <= 0 || outH <= 0)
    return cudaSuccess;
  PreviewGeom pg{outW, outH, outLine, lastRow, lastCol, hi2ho, wi2wo};
  const long long items = (long long)((outW + 1) / 2) * outH;
  const unsigned gy = (unsigned)((items + 255) / 256);
  if (gy > 65535u)
    return cudaErrorInvalidValue;
  dim3 grid((unsigned)numFrames, gy);
#define TRIK_PREVIEW(K)                                                                                            \
  preview_base_kernel<K><<<grid, 256, 0, stream>>>(g, frames, params, paramStride, bitmaps, pg, previews, previewStride); \
  preview_overlay_kernel<K><<<(unsigned)numFrames, 32, 0, stream>>>(g, params, paramStride, draw, omColours, pg, previews, previewStride)
  switch (kind)
  {
    case KIND_WO: TRIK_PREVIEW(KIND_WO); break;
    case KIND_WL: TRIK_PREVIEW(KIND_WL); break;
    case KIND_OO: TRIK_PREVIEW(KIND_OO); break;
    case KIND_OL: TRIK_PREVIEW(KIND_OL); break;
    case KIND_OM: TRIK_PREVIEW(KIND_OM); break;
    default: return cudaErrorInvalidValue;
  }
#undef TRIK_PREVIEW
  g_launches_preview += 2;
  return cudaGetLastError();
}

} // namespace trikb200
