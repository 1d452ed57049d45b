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
#include "trik_lut.cuh"

namespace trikb200 {

extern std::atomic<long long> g_launches_preview;
std::atomic<long long> g_launches_preview{0};
static long long g_previewChunkBytes = 0;                // images of one sub-batch of a 1:1 preview pass (0 = whole batch at once, the default)
void set_preview_chunk_bytes(long long bytes) { g_previewChunkBytes = bytes; }
static int g_sectorOverlay = -1;                         // line sensors, 1:1 preview, where the overlays are drawn: -1 = inside the streaming
                                                         // kernel, CTA-local (default); 0 = generic overlay kernel; 1 = full-sector kernel;
                                                         // 2..8 = inside, counting (forced blocks per CTA); 9..15 = inside, CTA-local (forced)
void set_preview_sector_overlay(int on) { g_sectorOverlay = on; }
static int g_previewTable = 1;                           // WO, 1:1 preview: detection through the batch's chroma table when there is one
void set_preview_table(int on) { g_previewTable = on; }

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
// 1:1 previews (outputWidth x outputHeight == the frame: the reference's default for the object sensors) -- both index maps
// are the identity, so the gather is a plain stream: 8 pixels per thread, one 16-byte load (two 8-byte loads for the planar
// sensors) and one 16-byte store, the colour conversion on pixel PAIRS in 16-bit lanes as in the sum kernels.  The generic
// kernel above (one scalar load and a 64-bit division per destination pair) stays for every other geometry.
// colFirst..colLast: the source columns pass 2 visits at all (5..W-5 for the ov7670 line sensor, everything else 0..W-1).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rgb565x_pair(uint32_t r2, uint32_t g2, uint32_t b2)       // lanes 0x00CC00CC
{
  return ((r2 >> 3) & 0x001F001Fu) | ((g2 << 3) & 0x07E007E0u) | ((b2 << 8) & 0xF800F800u);
}

template <int KIND>
__device__ __forceinline__ void lines_sector_item(const Geometry& g, const int2 d, const int outLine, uint8_t* img, const int item,
                                                  const uint32_t chunkFirst = 0u, const uint32_t chunkEnd = 0xFFFFFFFFu);

// FUSE (line sensors): the frame's line overlays inside this kernel, as the same full-sector read-modify-writes as
// preview_lines_sector_kernel but on lines that have just been written, i.e. still in L2 (no sector fill from DRAM, one
// write-back per line), and without a second launch.  Either CTA-local (`local`, the default: every CTA patches its own rows
// after a barrier) or counting (the CTAs of a frame count themselves in at the frame's DrawInfo record, v[19], back at zero
// afterwards, and the last one to arrive patches the whole frame).  A CTA takes `iters` blocks of 256 items.
template <int KIND, bool FUSE>
__global__ void __launch_bounds__(256)
preview_identity_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
                        const int paramStride, const uint16_t* __restrict__ bitmaps, const int outLine,
                        uint8_t* __restrict__ previews, const long long previewStride, const int numFrames,
                        const uint32_t cprMagic, const int colFirst, const int colLast, DrawInfo* draw, const int iters, const int local,
                        const uint8_t* __restrict__ lutTable, const uint32_t* __restrict__ lutMasks)
{
  constexpr bool PLANAR = (KIND == KIND_OO || KIND == KIND_OL || KIND == KIND_OM);
  __shared__ uint16_t s_lut43[256];
  __shared__ uint16_t s_lut255[256];
  __shared__ HueLutEntry s_lutHue[KIND == KIND_WO ? 256 : 1];
  if (KIND == KIND_WO)
  {
    fill_div_luts(s_lut43, s_lut255);
    fill_hue_lut(s_lutHue);
    __syncthreads();
  }
  __shared__ int s_last;
  const int cpr = g.width >> 3;
  const uint32_t item0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (!FUSE && item0 >= (uint32_t)cpr * (uint32_t)g.height)
    return;
  const int row0 = (int)__umulhi(item0, cprMagic), cc0 = (int)item0 - row0 * cpr;
  for (int frame = blockIdx.y; frame < numFrames; frame += gridDim.y)
  {
    // FUSE: a CTA takes `iters` consecutive blocks of 256 items, so that the counting-in below is paid once per several rows
    for (int it = 0; it < (FUSE ? iters : 1); ++it)
    {
    int row = row0, c = cc0;
    if (FUSE)
    {
      const uint32_t item = (blockIdx.x * (uint32_t)iters + (uint32_t)it) * blockDim.x + threadIdx.x;
      if (item >= (uint32_t)cpr * (uint32_t)g.height)
        break;
      row = (int)__umulhi(item, cprMagic); c = (int)item - row * cpr;
    }
    const int col0 = c * 8;
    const uint8_t* fr = frames + (size_t)frame * g.frameStride;
    uint32_t yy[4], cw[4];
    if (!PLANAR)
    {
      const uint4 in = ld_stream(fr + (size_t)row * g.lineLength + (size_t)c * 16);
      cw[0] = in.x; cw[1] = in.y; cw[2] = in.z; cw[3] = in.w;
#pragma unroll
      for (int j = 0; j < 4; ++j) yy[j] = cw[j] & 0x00FF00FFu;
    }
    else
    {
      const uint2 l = *reinterpret_cast<const uint2*>(fr + (size_t)row * g.lineLength + (size_t)c * 8);
      const uint2 ch = *reinterpret_cast<const uint2*>(fr + (size_t)(g.height + row) * g.lineLength + (size_t)c * 8);
      yy[0] = __byte_perm(l.x, 0u, 0x4140); yy[1] = __byte_perm(l.x, 0u, 0x4342);
      yy[2] = __byte_perm(l.y, 0u, 0x4140); yy[3] = __byte_perm(l.y, 0u, 0x4342);
      cw[0] = ch.x; cw[1] = ch.x; cw[2] = ch.y; cw[3] = ch.y;
    }
    uint32_t from = 0, to = 0, expected = 0;
    if (KIND == KIND_WO || KIND == KIND_WL || KIND == KIND_OL)
    {
      const FrameParams& p = params[(size_t)frame * paramStride];
      from = p.from; to = p.to; expected = p.expected;
    }
    uint32_t on0 = 0, on1 = 0;                                   // OO: the two metapixels under these 8 pixels
    if (KIND == KIND_OO)
    {
      const uint32_t two = *reinterpret_cast<const uint32_t*>(bitmaps + ((size_t)frame * (g.height >> 2) + (size_t)(row >> 2)) * (g.width >> 2)
                                                              + (size_t)(col0 >> 2));
      on0 = __popc(two & 0xFFFFu) > 2 ? 0xFFFFFFFFu : 0u;
      on1 = __popc(two >> 16) > 2 ? 0xFFFFFFFFu : 0u;
    }
    // line sensors: V = max(R,G,B) inside vFrom..vTo, tested on the clamped keys (channel = bits 6..13 of a lane):
    // key - keyLo <= span (mod 2^16); an empty range never passes
    uint32_t keyLo2 = 0x7FFF7FFFu, span2 = 0u;
    if (KIND == KIND_WL || KIND == KIND_OL)
    {
      const uint32_t vf = (from >> 16) & 0xFFu, vt = (to >> 16) & 0xFFu;
      if (vf <= vt)
      {
        keyLo2 = (0x8000u + (vf << 6)) * 0x10001u;
        span2 = (((vt - vf) << 6) + 63u) * 0x10001u;
      }
    }
    HsvBounds bd{};
    if (KIND == KIND_WO)
      bd = make_bounds(from, to);
    uint32_t px[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      const ChromaCoef coef = !PLANAR ? coef_yuyv() : ((j & 1) ? coef_planar1() : coef_planar0());
      uint32_t kr, kg, kb;
      rgb_keys(yy[j], cw[j], coef, kr, kg, kb);
      // clamped keys, lanes 0x8000 + (channel << 6) + 6 low bits; the RGB565X fields are cut straight out of them
      // (R bits 9..13 -> 0..4, G bits 8..13 -> 5..10, B bits 9..13 -> 11..15), the shifts on the multiplier pipe
      const uint32_t cr = __vminu2(__vmaxu2(kr, 0x80008000u), 0xBFFFBFFFu);
      const uint32_t cg = __vminu2(__vmaxu2(kg, 0x80008000u), 0xBFFFBFFFu);
      const uint32_t cb = __vminu2(__vmaxu2(kb, 0x80008000u), 0xBFFFBFFFu);
      const uint32_t rgb = (__umulhi(cr, 1u << 23) & 0x001F001Fu) | (__umulhi(cg, 1u << 29) & 0x07E007E0u) | ((cb * 4u) & 0xF800F800u);
      uint32_t keep = 0xFFFFFFFFu;                               // 0xFFFF in the lanes that keep their own colour
      if (KIND == KIND_WO && lutTable)
      {
        // the batch's threshold set has its chroma table (the one the main pass has just used, trik_kernels_lut.cu): two
        // byte look-ups (L1 / L2) and two packed compares instead of the HSV arithmetic; ragged entries ask the luma masks
        const uint32_t ci = __byte_perm(cw[j], 0u, 0x4431);     // U | V << 8
        const uint32_t lo = __ldg(lutTable + ci), nhi = __ldg(lutTable + LUT_STRIDE_PLAIN + ci);
        const uint32_t bits = (lo * 256u + nhi == LUT_RAGGED_CODE) ? lut_pass_pair_masks(cw[j], ci, lutMasks)
                                                                   : lut_pass_pair(cw[j], lo, nhi);
        keep = ~(((bits >> 15) & 0x00010001u) * 0xFFFFu);
      }
      else if (KIND == KIND_WO)
      {
        // the pair-wise threshold of the sum kernel (V, then S, then hue, each with a warp-wide early out)
        const uint32_t det = detect_pair_bits(yy[j], cw[j], coef, s_lutHue, s_lut255, bd, expected);
        keep = ~(((det & 1u) * 0x0000FFFFu) | ((det >> 1) * 0xFFFF0000u));
      }
      else if (KIND == KIND_WL || KIND == KIND_OL)
      {
        const uint32_t d = __vsub2(__vmaxu2(cr, __vmaxu2(cg, cb)), keyLo2);
        const uint32_t over = __vmaxu2(d, span2) - span2;        // lane 0 <=> inside the range
        keep = __vminu2(over, 0x00010001u) * 0xFFFFu;
      }
      else if (KIND == KIND_OO)
        keep = j < 2 ? ~on0 : ~on1;
      uint32_t p2 = (rgb & keep) | (0xFFE0FFE0u & ~keep);        // detected: cyan 0x00ffff -> 0xFFE0
      if (KIND == KIND_OL)
      {
        const int cA = col0 + 2 * j;
        if (cA < colFirst || cA > colLast) p2 &= 0xFFFF0000u;
        if (cA + 1 < colFirst || cA + 1 > colLast) p2 &= 0x0000FFFFu;
      }
      px[j] = p2;
    }
    *reinterpret_cast<uint4*>(previews + (size_t)frame * previewStride + (size_t)row * outLine + (size_t)c * 16)
        = make_uint4(px[0], px[1], px[2], px[3]);
    }
    if (FUSE && local)
    {
      // CTA-local: a 32-byte sector is two consecutive items with an even first index and W/8 is even, so every sector lies
      // inside ONE CTA's blocks of 256 items: after a barrier the CTA patches the sectors the lines cross among the rows it has
      // just written itself (its own stores, read back from L2) -- no counting, no fence, nobody waits for anybody
      __syncthreads();
      const uint32_t total = (uint32_t)cpr * (uint32_t)g.height;
      const uint32_t a = blockIdx.x * (uint32_t)iters * blockDim.x, b = min(a + (uint32_t)iters * blockDim.x, total);
      if (a < b)
      {
        const int rowFirst = (int)__umulhi(a, cprMagic), rowLast = (int)__umulhi(b - 1u, cprMagic);
        const int2 d = __ldcg(reinterpret_cast<const int2*>(draw[frame].v));
        const int nLocal = (rowLast - rowFirst + 1) * 6;
        const int nAll = nLocal + (KIND == KIND_OL ? 2 * (g.width >> 4) : 0);
        uint8_t* img = previews + (size_t)frame * previewStride;
        for (int t = (int)threadIdx.x; t < nAll; t += (int)blockDim.x)
          lines_sector_item<KIND>(g, d, outLine, img, t < nLocal ? rowFirst * 6 + t : g.height * 6 + (t - nLocal), a, b);
      }
    }
    else if (FUSE)
    {
      // the CTA's pixels, then (barrier, cumulative fence of one thread: the grid-barrier idiom) its count
      __syncthreads();
      if (threadIdx.x == 0)
      {
        __threadfence();
        const uint32_t before = atomicAdd(reinterpret_cast<uint32_t*>(&draw[frame].v[19]), 1u);
        s_last = before == gridDim.x - 1u;
      }
      __syncthreads();
      if (s_last)
      {
        __threadfence();
        const int2 d = __ldcg(reinterpret_cast<const int2*>(draw[frame].v));
        const int items = g.height * 6 + (KIND == KIND_OL ? 2 * (g.width >> 4) : 0);
        uint8_t* img = previews + (size_t)frame * previewStride;
        for (int it = (int)threadIdx.x; it < items; it += (int)blockDim.x)
          lines_sector_item<KIND>(g, make_int2(d.x, d.y), outLine, img, it);
        if (threadIdx.x == 0)
          draw[frame].v[19] = 0;
      }
      __syncthreads();                                           // s_last is rewritten by the next frame of this CTA
    }
  }
}

// ---------------------------------------------------------------------------------------------
// overlays
// ---------------------------------------------------------------------------------------------
// One CTA of OVERLAY_THREADS threads per frame: primitives in reference order (a block barrier where a later primitive may
// overwrite an earlier one of another colour), the pixels of one primitive spread over the threads.
constexpr int OVERLAY_THREADS = 128;
struct Canvas {
  uint8_t* img;
  int W, H, outLine;
  const int32_t* hi2ho;
  const int32_t* wi2wo;
  __device__ __forceinline__ void put(int col, int row, uint32_t rgb) const      // drawOutputPixelBound (:72-89)
  {
    col = col < 0 ? 0 : (col > W - 1 ? W - 1 : col);
    row = row < 0 ? 0 : (row > H - 1 ? H - 1 : row);
    *reinterpret_cast<uint16_t*>(img + (size_t)__ldg(hi2ho + row) * outLine + (size_t)__ldg(wi2wo + col) * 2u) = rgb565x(rgb);
  }
};

// vertical / horizontal "target" lines of +-100 pixels (WO :136-168), one colour: lanes share the work
__device__ void centre_line(const Canvas& c, int lane, int col, int row, uint32_t rgb)
{
  for (int adj = lane; adj < 100; adj += OVERLAY_THREADS) { c.put(col, row - adj, rgb); c.put(col, row + adj, rgb); }
}
__device__ void horizontal_centre_line(const Canvas& c, int lane, int col, int row, uint32_t rgb)
{
  for (int adj = lane; adj < 100; adj += OVERLAY_THREADS) { c.put(col - adj, row, rgb); c.put(col + adj, row, rgb); }
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
__global__ void __launch_bounds__(OVERLAY_THREADS)
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
    __syncthreads();
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
    for (int adj = lane; adj < H; adj += OVERLAY_THREADS)
    {
      c.put(hW - 40, adj, 0xFF00FFu); c.put(hW + 40, adj, 0xFF00FFu);
      c.put(hW - 80, adj, 0xFF00FFu); c.put(hW + 80, adj, 0xFF00FFu);
    }
    __syncthreads();
    if (KIND == KIND_OL)
    {
      for (int adj = lane; adj < W; adj += OVERLAY_THREADS)                                  // drawRgbHorizontalLine (OL :454-455)
      {
        c.put(adj, hH, 0xFF0000u);
        c.put(adj, hH + 80, 0xFF0000u);
      }
      __syncthreads();
    }
    const DrawInfo d = draw[frame];
    if (d.v[0])                                                                 // points > 10: 3-wide red line (WL :410)
      for (int adj = lane; adj < H; adj += OVERLAY_THREADS)
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
        for (int k = lane; k < 400; k += OVERLAY_THREADS)
          c.put(c0 + k % 20, r0 + k / 20, rgb);
        __syncthreads();
      }
  }
}

// ---------------------------------------------------------------------------------------------
// Line sensors, 1:1 preview: the overlays as FULL-SECTOR read-modify-writes.
// Their overlays are full-height lines (four fixed columns, the 3-wide target line) and, for the ov7670 sensor, two full
// rows: 7 two-byte stores per row in the generic overlay kernel, nearly each in a 32-byte sector of its own, and a partial
// store to a sector that has left L2 costs a fill before the write-back (measured: 112 us per 1024 x 640x480 with 1 % of the
// issue slots in use).  Here a thread owns one 32-byte sector (16 pixels) that some line crosses: it loads the sector,
// applies every primitive in the reference's order to its 16 pixels, and stores the whole sector.  Items per frame:
// H rows x 6 candidate sectors (four fixed columns, first and last column of the target line; duplicates dropped), plus for
// the ov7670 sensor 2 rows x W/16 sectors.
// ---------------------------------------------------------------------------------------------
// chunkFirst .. chunkEnd: only sectors whose first 16-byte chunk (index row * W/8 + 2 * sector, as the streaming kernel counts
// its items) lies in this range -- the caller that has just written exactly these chunks
template <int KIND>
__device__ __forceinline__ void lines_sector_item(const Geometry& g, const int2 d, const int outLine, uint8_t* img, const int item,
                                                  const uint32_t chunkFirst, const uint32_t chunkEnd)
{
  const int W = g.width, H = g.height, hW = W >> 1, hH = H >> 1, spr = W >> 4;
  auto clampc = [&](int x) { return x < 0 ? 0 : (x > W - 1 ? W - 1 : x); };
  auto clampr = [&](int y) { return y < 0 ? 0 : (y > H - 1 ? H - 1 : y); };
  const int c0 = clampc(hW - 80), c1 = clampc(hW - 40), c2 = clampc(hW + 40), c3 = clampc(hW + 80);   // WL :384-387
  const int rA = clampr(hH), rB = clampr(hH + 80);                                                    // OL :454-455
  constexpr uint32_t MAGENTA = 0xF81Fu, RED = 0x001Fu;             // rgb565x(0xFF00FF), rgb565x(0xFF0000)
  // d: points > 10, target column (WL :410)
  const int L = clampc(d.y - 1), R = clampc(d.y + 1);
  int row, sec;
  bool wholeRow = false;
  if (item < H * 6)
  {
    row = item / 6;
    const int sIdx = item - row * 6;
    const int cand[6] = {c0 >> 4, c1 >> 4, c2 >> 4, c3 >> 4, d.x ? L >> 4 : -1, d.x ? R >> 4 : -1};
    sec = -1;
    bool dup = false;
#pragma unroll
    for (int j = 0; j < 6; ++j)
    {
      if (j == sIdx) sec = cand[j];
    }
#pragma unroll
    for (int j = 0; j < 6; ++j)
      if (j < sIdx && cand[j] == sec) dup = true;
    if (sec < 0 || dup)
      return;
    if (KIND == KIND_OL && (row == rA || row == rB))
      return;                                                    // the whole row is somebody else's
  }
  else if (KIND == KIND_OL && item < H * 6 + 2 * spr)
  {
    const int idx = item - H * 6;
    const int which = idx >= spr ? 1 : 0;
    if (which == 1 && rB == rA)
      return;
    row = which ? rB : rA;
    sec = idx - which * spr;
    wholeRow = true;
  }
  else
    return;
  const uint32_t chunk = (uint32_t)row * (uint32_t)(W >> 3) + 2u * (uint32_t)sec;
  if (chunk < chunkFirst || chunk >= chunkEnd)
    return;
  uint4* const p = reinterpret_cast<uint4*>(img + (size_t)row * outLine + (size_t)sec * 32);
  const uint4 a = __ldcg(p), b = __ldcg(p + 1);                  // L2: the fused caller reads what other SMs have just written
  uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int k = 0; k < 16; ++k)
  {
    const int col = sec * 16 + k;
    int colour = -1;
    if (col == c0 || col == c1 || col == c2 || col == c3) colour = (int)MAGENTA;
    if (wholeRow) colour = (int)RED;
    if (d.x && col >= L && col <= R) colour = (int)RED;
    if (colour >= 0)
      w[k >> 1] = (k & 1) ? (w[k >> 1] & 0x0000FFFFu) | ((uint32_t)colour << 16) : (w[k >> 1] & 0xFFFF0000u) | (uint32_t)colour;
  }
  p[0] = make_uint4(w[0], w[1], w[2], w[3]);
  p[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

template <int KIND>
__global__ void __launch_bounds__(128)
preview_lines_sector_kernel(const Geometry g, const DrawInfo* __restrict__ draw, const int outLine,
                            uint8_t* __restrict__ previews, const long long previewStride, const int numFrames)
{
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  for (int frame = blockIdx.y; frame < numFrames; frame += gridDim.y)
  {
    const int2 d = *reinterpret_cast<const int2*>(draw[frame].v);
    lines_sector_item<KIND>(g, d, outLine, previews + (size_t)frame * previewStride, item);
  }
}

cudaError_t launch_preview(int kind, const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                           int paramStride, const uint16_t* bitmaps, const DrawInfo* draw, const int32_t* omColours,
                           int outW, int outH, int outLine, const int32_t* lastRow, const int32_t* lastCol,
                           const int32_t* hi2ho, const int32_t* wi2wo, uint8_t* previews, long long previewStride,
                           cudaStream_t stream, const uint8_t* lutTable, const uint32_t* lutMasks)
{
  // WO: the chroma table of the batch's ONE threshold set, when the main pass went through it (else NULL: arithmetic)
  const uint8_t* pvTable = (kind == KIND_WO && g_previewTable) ? lutTable : nullptr;
  const uint32_t* pvMasks = pvTable ? lutMasks : nullptr;
  if (numFrames <= 0 || outW <= 0 || outH <= 0)
    return cudaSuccess;
  PreviewGeom pg{outW, outH, outLine, lastRow, lastCol, hi2ho, wi2wo};
  const long long items = (long long)((outW + 1) / 2) * outH;
  const unsigned gy = (unsigned)((items + 255) / 256);
  if (gy > 65535u)
    return cudaErrorInvalidValue;
  dim3 grid((unsigned)numFrames, gy);
  // identity maps (1:1 preview) and 16-byte aligned rows: the streaming kernel
  const bool identity = outW == g.width && outH == g.height && (g.width & 7) == 0 && (outLine & 15) == 0
                        && (previewStride & 15) == 0 && (((uintptr_t)previews) & 15) == 0 && (g.lineLength & 15) == 0
                        && (g.frameStride & 15) == 0 && (((uintptr_t)frames) & 15) == 0
                        && (long long)(g.width >> 3) * g.height * (g.width >> 3) < (1ll << 32);
  if (identity)
  {
    const int cpr = g.width >> 3;
    const uint32_t cprMagic = (uint32_t)((1ull << 32) / (uint32_t)cpr) + 1u;
    const dim3 igrid((unsigned)(((long long)cpr * g.height + 255) / 256), (unsigned)(numFrames < 65535 ? numFrames : 65535));
    const int colFirst = kind == KIND_OL ? 5 : 0, colLast = kind == KIND_OL ? g.width - 5 : g.width - 1;
    // The overlay kernel's stores are two bytes each, nearly every one in a sector of its own: on an image that has left
    // L2 each costs a sector fill and a write-back (measured: 112 us for 1024 x 640x480, 1 % of the issue slots used).
    // Sub-batches whose images fit in L2 together (so that the overlay lands on lines the base-layer kernel has just written)
    // were measured and are SLOWER at every size tried -- 1024 x 640x480 WL: 0.500 ms at once, 0.506 / 0.516 / 0.573 / 0.753 ms
    // with 96 / 64 / 40 / 20 MiB sub-batches (the kernel boundaries cost more than the L2 hits save) -- so the default is the
    // whole batch at once; the knob stays for measurements (profiles/r02m_sweep_preview_chunks.jsonl).
    const size_t perFrame = (size_t)outH * outLine;
    int chunk = numFrames;
    if (g_previewChunkBytes > 0 && perFrame > 0)
    {
      chunk = (int)(g_previewChunkBytes / perFrame);
      if (chunk < 8) chunk = 8;
    }
    const size_t cells = (size_t)(g.width >> 2) * (g.height >> 2);
    for (int f0 = 0; f0 < numFrames; f0 += chunk)
    {
      const int cnt = numFrames - f0 < chunk ? numFrames - f0 : chunk;
      const dim3 cgrid(igrid.x, (unsigned)(cnt < 65535 ? cnt : 65535));
      // Line sensors: the overlays inside the streaming kernel (one launch fewer, the sectors patched while still in L2).
      //  * default (-1) and 9..15, CTA-local: a CTA patches the sectors the lines cross among the rows it has just written
      //    itself, after one barrier.  Paid per CTA, so a CTA takes several blocks of 256 items: 1 / 4 / 8 / 16 / 32 blocks
      //    0.500 / 0.416 / 0.404 / 0.410 / 0.422 ms per 1024 x 640x480 WL against 0.477 ms with the separate sector kernel,
      //    OL 0.486 against 0.548 ms, 4096 x 320x240 0.424 against 0.474 ms, 1024 x 320x240 0.123 against 0.130 ms,
      //    64 x 320x240 (one block per CTA) 0.0172 against 0.0179 ms (profiles/r02z_*): by default up to 8 blocks, as long as
      //    about four waves of CTAs are left.
      //  * 2..8, counting: the CTAs of a frame count themselves in at the frame's DrawInfo record (every frame is visited by
      //    exactly gridDim.x CTAs) and the last one patches the whole frame: 0.438 ms at best for the same 1024 x 640x480 WL
      //    (16 blocks per CTA), and short passes cannot hide the last CTA's H * 6 / 256 dependent rounds (64 x 320x240: 0.034
      //    against 0.018 ms; profiles/r02w_*, r02x_*).  Kept for measurements.
      const bool fused = (g_sectorOverlay < 0 || g_sectorOverlay >= 2) && draw != nullptr && (kind == KIND_WL || kind == KIND_OL);
      const int fuseLocal = (g_sectorOverlay < 0 || g_sectorOverlay >= 9) ? 1 : 0;
      unsigned fuseIters = 1u;
      if (g_sectorOverlay >= 9)
        fuseIters = 1u << (g_sectorOverlay <= 15 ? g_sectorOverlay - 9 : 6);
      else if (g_sectorOverlay >= 2)
        fuseIters = 1u << (g_sectorOverlay - 2);
      else
        while (fuseIters < 8u && (unsigned long long)cgrid.x * cgrid.y / (fuseIters * 2u) >= 4ull * 148ull * 8ull)
          fuseIters *= 2u;
      const uint8_t* cf = frames + (size_t)f0 * g.frameStride;
      const FrameParams* cp = params + (size_t)f0 * paramStride;
      const uint16_t* cb = bitmaps ? bitmaps + (size_t)f0 * cells : nullptr;
      const DrawInfo* cd = draw ? draw + f0 : nullptr;
      const int32_t* co = omColours ? omColours + (size_t)f0 * 100 : nullptr;
      uint8_t* cpv = previews + (size_t)f0 * previewStride;
      int launched = 2;                                          // base layer + overlays; one when the overlays are inside
#define TRIK_PREVIEW_ID(K)                                                                                         \
  preview_identity_kernel<K, false><<<cgrid, 256, 0, stream>>>(g, cf, cp, paramStride, cb, outLine, cpv,               \
                                                        previewStride, cnt, cprMagic, colFirst, colLast, nullptr, 1, 0, pvTable, pvMasks);     \
  preview_overlay_kernel<K><<<(unsigned)cnt, OVERLAY_THREADS, 0, stream>>>(g, cp, paramStride, cd, co, pg, cpv, previewStride)
#define TRIK_PREVIEW_LINES(K)                                                                                      \
  preview_identity_kernel<K, false><<<cgrid, 256, 0, stream>>>(g, cf, cp, paramStride, cb, outLine, cpv,               \
                                                        previewStride, cnt, cprMagic, colFirst, colLast, nullptr, 1, 0, pvTable, pvMasks);     \
  preview_lines_sector_kernel<K><<<dim3((unsigned)((g.height * 6 + 2 * (g.width >> 4) + 127) / 128), cgrid.y), 128, 0, stream>>>( \
      g, cd, outLine, cpv, previewStride, cnt)
#define TRIK_PREVIEW_FUSED(K)                                                                                      \
  preview_identity_kernel<K, true><<<dim3((cgrid.x + fuseIters - 1) / fuseIters, cgrid.y), 256, 0, stream>>>(          \
      g, cf, cp, paramStride, cb, outLine, cpv, previewStride, cnt, cprMagic, colFirst, colLast, const_cast<DrawInfo*>(cd), (int)fuseIters, fuseLocal, nullptr, nullptr); \
  launched = 1
      switch (kind)
      {
        case KIND_WO: TRIK_PREVIEW_ID(KIND_WO); break;
        case KIND_WL: if (fused) { TRIK_PREVIEW_FUSED(KIND_WL); } else if (g_sectorOverlay) { TRIK_PREVIEW_LINES(KIND_WL); } else { TRIK_PREVIEW_ID(KIND_WL); } break;
        case KIND_OO: TRIK_PREVIEW_ID(KIND_OO); break;
        case KIND_OL: if (fused) { TRIK_PREVIEW_FUSED(KIND_OL); } else if (g_sectorOverlay) { TRIK_PREVIEW_LINES(KIND_OL); } else { TRIK_PREVIEW_ID(KIND_OL); } break;
        case KIND_OM: TRIK_PREVIEW_ID(KIND_OM); break;
        default: return cudaErrorInvalidValue;
      }
#undef TRIK_PREVIEW_FUSED
#undef TRIK_PREVIEW_LINES
#undef TRIK_PREVIEW_ID
      g_launches_preview += launched;
    }
    return cudaGetLastError();
  }
#define TRIK_PREVIEW(K)                                                                                            \
  preview_base_kernel<K><<<grid, 256, 0, stream>>>(g, frames, params, paramStride, bitmaps, pg, previews, previewStride); \
  preview_overlay_kernel<K><<<(unsigned)numFrames, OVERLAY_THREADS, 0, stream>>>(g, params, paramStride, draw, omColours, pg, previews, previewStride)
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
