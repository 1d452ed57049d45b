// trik_kernels.cu -- the fused per-frame kernels of libtrikb200 (sm_100a).
//
// Sum sensors (this file, sum_kernel<KIND>):
//   WL  webcam line sensor   pass 1+2 of webcam/line_sensor/include/internal/cv_line_detector_seqpass.hpp:197-269, tail :401-417
//   OL  ov7670 line sensor   ov7670/line_sensor/include/internal/cv_line_detector_seqpass.hpp:210-301, tail :449-473
//   WO  webcam object sensor webcam/object_sensor/include/internal/cv_ball_detector_seqpass.hpp:251-354, tail :486-505
// The reference converts the whole frame to an 8-byte-per-pixel RGB+HSV image (s_rgb888hsv, :24)
// and re-reads it; here one pass reads the 2 bytes per pixel once and nothing else touches HBM
// but 16 result bytes per frame.
//
// Work decomposition: a CTA owns a slab of consecutive rows of one frame.  A thread owns ONE
// 16-byte column chunk (8 pixels YUYV / 16 pixels YUV422P) and walks down the rows with a
// stride of rowsPerIter = blockDim / chunksPerRow, so its column never changes: the column
// sums collapse to (fail count) * column + (tiny packed in-chunk sums) at the very end, and a
// warp's 32 loads are one contiguous 512-byte run of a row.  All sums are integer and order
// independent, so shuffles + shared memory + one atomic per CTA per sum stay bit-exact.
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"

namespace trikb200 {

static long long g_launches = 0;
extern long long g_launches_grid;
extern long long g_launches_detect;
long long launch_count() { return g_launches + g_launches_grid + g_launches_detect; }

// ---------------------------------------------------------------------------------------------
// per-pair work
// ---------------------------------------------------------------------------------------------

// V-only threshold (line sensors: H and S bounds are fixed to 0..255, WL/.../cv_line_detector_seqpass.hpp:345-348).
// Returns the FAIL lanes (0/1 in each 16-bit lane) of the pixel pair.
__device__ __forceinline__ uint32_t vfail_pair(uint32_t yy, uint32_t cw, const ChromaCoef coef,
                                               uint32_t negKlo2, uint32_t n2)
{
  uint32_t kr, kg, kb;
  rgb_keys(yy, cw, coef, kr, kg, kb);
  const uint32_t km = __vimax3_u16x2(kr, kg, kb);         // key of max(R,G,B) before >>6/saturation
  const uint32_t d  = __vadd2(km, negKlo2);               // (key - klo) mod 2^16 per lane
  const uint32_t x  = __vmaxu2(d, n2);                    // == n2 lane-wise iff d <= N (pass)
  return __vminu2(x - n2, 0x00010001u);                   // 0 = pass, 1 = fail (no lane borrow: x >= n2)
}

__device__ __forceinline__ uint32_t hsvfail_pair(uint32_t yy, uint32_t cw, const ChromaCoef coef,
                                                 const uint16_t* lut43, const uint16_t* lut255,
                                                 uint32_t from, uint32_t to, uint32_t expected)
{
  uint32_t h0, h1;
  hsv_pair(yy, cw, coef, lut43, lut255, h0, h1);
  return (detect_hsv(h0, from, to, expected) ? 0u : 1u) | (detect_hsv(h1, from, to, expected) ? 0u : 0x10000u);
}

// ---------------------------------------------------------------------------------------------
// finalisation: raw sums -> OutArgs, exactly the integer / float steps of the reference tails
// ---------------------------------------------------------------------------------------------
template <int KIND>
__device__ void finalize_sum(const Geometry& g, const FrameParams& p, uint32_t fails, uint32_t sxFail,
                             uint32_t syFail, uint32_t crossFail, TargetOut* o)
{
  const uint32_t W = (uint32_t)g.width, H = (uint32_t)g.height;
  TargetOut r;
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
    }
  }
  *o = r;
}

// ---------------------------------------------------------------------------------------------
// the sum kernel
// ---------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(1024)
sum_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
           const int paramStride, SumAcc* __restrict__ acc, TargetOut* __restrict__ out,
           const int slabs, const int rowsPerSlab, const int cpr, const int rpi)
{
  constexpr bool PLANAR = (KIND == KIND_OL);
  __shared__ uint16_t s_lut43[256];
  __shared__ uint16_t s_lut255[256];
  __shared__ uint32_t s_red[32][4];

  const int frame = blockIdx.x / slabs;
  const int slab  = blockIdx.x - frame * slabs;
  const int t  = threadIdx.x;
  const int cc = t % cpr;                  // chunk column of this thread, constant over the loop
  const int rr = t / cpr;                  // row offset inside one iteration
  const FrameParams p = params[(size_t)frame * paramStride];

  if (KIND == KIND_WO)
  {
    fill_div_luts(s_lut43, s_lut255);
    __syncthreads();
  }

  const int r0 = slab * rowsPerSlab;
  const int r1 = min(r0 + rowsPerSlab, g.height);
  const uint8_t* base = frames + (size_t)frame * g.frameStride + (size_t)cc * 16u;
  const size_t rowStep = (size_t)rpi * g.lineLength;

  uint32_t S = 0, A = 0, SI = 0, SC = 0;   // packed 16x2 lane accumulators (even pixel | odd pixel << 16)

  // OL column window 5..W-5: chunk 0 loses pixels 0..4, the last chunk loses its pixels 12..15
  uint32_t m01 = 0x00010001u, m2 = 0x00010001u, m67 = 0x00010001u;
  if (KIND == KIND_OL)
  {
    if (cc == 0)       { m01 = 0u; m2 = 0x00010000u; }
    if (cc == cpr - 1) { m67 = 0u; }
  }

  int row = r0 + rr;
  const uint8_t* ptr = base + (size_t)row * g.lineLength;
  const size_t chromaOfs = (size_t)g.height * g.lineLength;
  uint4 cur = make_uint4(0, 0, 0, 0), curC = make_uint4(0, 0, 0, 0);
  if (row < r1)
  {
    cur = ld_stream(ptr);
    if (PLANAR) curC = ld_stream(ptr + chromaOfs);
  }
  for (uint32_t it = 0; row < r1; ++it)
  {
    // prefetch the next row's chunk before working on this one
    uint4 nxt = make_uint4(0, 0, 0, 0), nxtC = make_uint4(0, 0, 0, 0);
    const int nrow = row + rpi;
    if (nrow < r1)
    {
      nxt = ld_stream(ptr + rowStep);
      if (PLANAR) nxtC = ld_stream(ptr + rowStep + chromaOfs);
    }

    uint32_t Sc = 0;                       // fail lanes of this chunk
    if (!PLANAR)
    {
      const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
      {
        const uint32_t yy = w[k] & 0x00FF00FFu;
        uint32_t fl;
        if (KIND == KIND_WO)
          fl = hsvfail_pair(yy, w[k], coef_yuyv(), s_lut43, s_lut255, p.from, p.to, p.expected);
        else
          fl = vfail_pair(yy, w[k], coef_yuyv(), p.negKlo2, p.n2);
        Sc += fl;
        A  += (uint32_t)k * fl;
      }
    }
    else
    {
      const uint32_t L[4] = {cur.x, cur.y, cur.z, cur.w};
      const uint32_t Cw[4] = {curC.x, curC.y, curC.z, curC.w};
#pragma unroll
      for (int k = 0; k < 8; ++k)
      {
        const uint32_t yy = __byte_perm(L[k >> 1], 0u, (k & 1) ? 0x4342 : 0x4140);
        uint32_t fl = vfail_pair(yy, Cw[k >> 1], (k & 1) ? coef_planar1() : coef_planar0(), p.negKlo2, p.n2);
        if (k < 2)  fl &= m01;
        if (k == 2) fl &= m2;
        if (k >= 6) fl &= m67;
        Sc += fl;
        A  += (uint32_t)k * fl;
      }
    }
    S += Sc;
    if (KIND == KIND_WO)
      SI += it * Sc;
    if (KIND == KIND_OL)
      if ((uint32_t)row - p.hStart <= p.hStop - p.hStart && p.hStart <= p.hStop)
        SC += Sc;

    cur = nxt; curC = nxtC;
    row = nrow;
    ptr += rowStep;
  }

  // unpack the lanes
  constexpr uint32_t PXC = PLANAR ? 16u : 8u;
  uint32_t fails  = (S & 0xFFFFu) + (S >> 16);
  uint32_t inIdx  = 2u * ((A & 0xFFFFu) + (A >> 16)) + (S >> 16);   // sum of in-chunk pixel indices of the fails
  uint32_t sxFail = fails * ((uint32_t)cc * PXC) + inIdx;
  uint32_t syFail = 0, crossFail = 0;
  if (KIND == KIND_WO)
    syFail = fails * (uint32_t)(r0 + rr) + (uint32_t)rpi * ((SI & 0xFFFFu) + (SI >> 16));
  if (KIND == KIND_OL)
    crossFail = (SC & 0xFFFFu) + (SC >> 16);

  // CTA reduction
  __syncthreads();
  const unsigned am = __activemask();
  fails  = __reduce_add_sync(am, fails);
  sxFail = __reduce_add_sync(am, sxFail);
  if (KIND == KIND_WO) syFail = __reduce_add_sync(am, syFail);
  if (KIND == KIND_OL) crossFail = __reduce_add_sync(am, crossFail);
  const int warp = t >> 5, lane = t & 31, nwarps = (blockDim.x + 31) >> 5;
  if (lane == 0)
  {
    s_red[warp][0] = fails; s_red[warp][1] = sxFail; s_red[warp][2] = syFail; s_red[warp][3] = crossFail;
  }
  __syncthreads();
  if (warp == 0)
  {
    uint32_t a = 0, b = 0, c = 0, d = 0;
    if (lane < nwarps) { a = s_red[lane][0]; b = s_red[lane][1]; c = s_red[lane][2]; d = s_red[lane][3]; }
    const unsigned fm = __activemask();
    a = __reduce_add_sync(fm, a);
    b = __reduce_add_sync(fm, b);
    if (KIND == KIND_WO) c = __reduce_add_sync(fm, c);
    if (KIND == KIND_OL) d = __reduce_add_sync(fm, d);
    if (lane == 0)
    {
      SumAcc* fa = acc + frame;
      bool last = true;
      if (slabs > 1)
      {
        atomicAdd(&fa->fails, a);
        atomicAdd(&fa->sxFail, b);
        if (KIND == KIND_WO) atomicAdd(&fa->syFail, c);
        if (KIND == KIND_OL) atomicAdd(&fa->crossFail, d);
        __threadfence();
        last = (atomicAdd(&fa->done, 1u) == (uint32_t)slabs - 1u);
        if (last)
        {
          __threadfence();
          a = atomicExch(&fa->fails, 0u);
          b = atomicExch(&fa->sxFail, 0u);
          c = atomicExch(&fa->syFail, 0u);
          d = atomicExch(&fa->crossFail, 0u);
          atomicExch(&fa->done, 0u);
        }
      }
      if (last)
        finalize_sum<KIND>(g, p, a, b, c, d, out + frame);
    }
  }
}

int sum_sensor_block_threads(int kind, int width)
{
  const int cpr = width / (kind == KIND_OL ? 16 : 8);
  if (cpr <= 0 || cpr > 1024)
    return 0;
  int k = (256 + cpr - 1) / cpr;
  for (int j = 0; j < 16; ++j)
    if ((cpr * (k + j)) % 32 == 0 && cpr * (k + j) <= 1024)
    {
      k += j;
      break;
    }
  if (cpr * k > 1024)
    k = 1024 / cpr;
  return cpr * k;
}

cudaError_t launch_sum_sensor(int kind, const Geometry& g, int numFrames, const uint8_t* frames,
                              const FrameParams* params, int paramStride, SumAcc* acc, TargetOut* out,
                              int slabsPerFrame, cudaStream_t stream)
{
  if (numFrames <= 0)
    return cudaSuccess;
  const int threads = sum_sensor_block_threads(kind, g.width);
  if (threads <= 0)
    return cudaErrorInvalidValue;
  const int cpr = g.width / (kind == KIND_OL ? 16 : 8);
  const int rpi = threads / cpr;
  // SI lanes (WO) hold sum(it * fails_per_chunk) <= 8 * I*(I-1)/2: keep I <= 128 iterations per slab
  const int maxRowsPerSlab = 128 * rpi;
  int slabs = slabsPerFrame;
  if (slabs <= 0)
  {
    const int wantCtas = 148 * 16;
    slabs = (wantCtas + numFrames - 1) / numFrames;
    const int maxSlabs = g.height / (rpi * 4) > 0 ? g.height / (rpi * 4) : 1;
    if (slabs > maxSlabs) slabs = maxSlabs;
    if (slabs < 1) slabs = 1;
  }
  int rowsPerSlab = (g.height + slabs - 1) / slabs;
  rowsPerSlab = ((rowsPerSlab + rpi - 1) / rpi) * rpi;
  if (rowsPerSlab > maxRowsPerSlab)
    rowsPerSlab = maxRowsPerSlab;
  slabs = (g.height + rowsPerSlab - 1) / rowsPerSlab;
  if (slabs < 1) slabs = 1;
  const long long grid = (long long)numFrames * slabs;
  if (grid > 0x7FFFFFFFLL)
    return cudaErrorInvalidValue;
  switch (kind)
  {
    case KIND_WL:
      sum_kernel<KIND_WL><<<(unsigned)grid, threads, 0, stream>>>(g, frames, params, paramStride, acc, out, slabs, rowsPerSlab, cpr, rpi);
      break;
    case KIND_OL:
      sum_kernel<KIND_OL><<<(unsigned)grid, threads, 0, stream>>>(g, frames, params, paramStride, acc, out, slabs, rowsPerSlab, cpr, rpi);
      break;
    case KIND_WO:
      sum_kernel<KIND_WO><<<(unsigned)grid, threads, 0, stream>>>(g, frames, params, paramStride, acc, out, slabs, rowsPerSlab, cpr, rpi);
      break;
    default:
      return cudaErrorInvalidValue;
  }
  ++g_launches;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// exhaustive pixel-function probes (tests only call them through the C ABI)
// ---------------------------------------------------------------------------------------------
__global__ void probe_yuv2rgb_kernel(uint32_t first, uint32_t count, uint32_t* __restrict__ out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t idx = first + i;                     // idx = Y | U << 8 | V << 16
  const uint32_t y = idx & 0xFFu, u = (idx >> 8) & 0xFFu, v = (idx >> 16) & 0xFFu;
  // exercise both lanes and both layouts: lane 0 carries the probe, lane 1 the complementary luma
  const uint32_t word = y | (u << 8) | ((255u - y) << 16) | (v << 24);
  uint32_t kr, kg, kb;
  rgb_keys(word & 0x00FF00FFu, word, coef_yuyv(), kr, kg, kb);
  const uint32_t r2 = chan8_from_key(kr), g2 = chan8_from_key(kg), b2 = chan8_from_key(kb);
  uint32_t rgb = ((r2 & 0xFFu) << 16) | ((g2 & 0xFFu) << 8) | (b2 & 0xFFu);
  // the planar path must agree: chroma word [V U V U], luma in the high lane this time
  const uint32_t cw = v | (u << 8) | (v << 16) | (u << 24);
  uint32_t pr, pg, pb;
  rgb_keys(((255u - y) & 0xFFu) | (y << 16), cw, coef_planar1(), pr, pg, pb);
  const uint32_t r3 = chan8_from_key(pr), g3 = chan8_from_key(pg), b3 = chan8_from_key(pb);
  const uint32_t rgbP = ((r3 >> 16) << 16) | ((g3 >> 16) << 8) | (b3 >> 16);
  if (rgbP != rgb)
    rgb |= 0x80000000u;                               // flags a layout disagreement
  out[i] = rgb;
}

__global__ void probe_rgb2hsv_kernel(uint32_t first, uint32_t count, uint32_t* __restrict__ out)
{
  __shared__ uint16_t s_lut43[256];
  __shared__ uint16_t s_lut255[256];
  fill_div_luts(s_lut43, s_lut255);
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t rgb = first + i;
  out[i] = hsv_from_rgb8((int32_t)((rgb >> 16) & 0xFFu), (int32_t)((rgb >> 8) & 0xFFu), (int32_t)(rgb & 0xFFu),
                         s_lut43, s_lut255);
}

cudaError_t launch_probe_yuv2rgb(uint32_t first, uint32_t count, uint32_t* out, cudaStream_t stream)
{
  if (!count) return cudaSuccess;
  probe_yuv2rgb_kernel<<<(count + 255u) / 256u, 256, 0, stream>>>(first, count, out);
  ++g_launches;
  return cudaGetLastError();
}

cudaError_t launch_probe_rgb2hsv(uint32_t first, uint32_t count, uint32_t* out, cudaStream_t stream)
{
  if (!count) return cudaSuccess;
  probe_rgb2hsv_kernel<<<(count + 255u) / 256u, 256, 0, stream>>>(first, count, out);
  ++g_launches;
  return cudaGetLastError();
}

} // namespace trikb200
