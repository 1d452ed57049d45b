// trik_kernels_edge.cu -- ov7670/edge_line_sensor: Sobel 3x3 on the luma plane, threshold, column
// centroid of the edge pixels.
//
// Reference: ov7670/edge_line_sensor/include/internal/cv_ball_detector_seqpass.hpp
//   :176-182  IMG_sobel_3x3_8 + IMG_thr_gt2max_8(.., 50) on the luma plane          (TI IMGLIB: closed, not vendored)
//   :186-205  count / column-sum of the pixels == 0xFF over columns 16 .. W-16, per row in uint16_t
//   :386-414  targetX = centroid column, targetY from a sum that is never added to (-100 whenever anything is found),
//             targetSize = ceil(sqrt(points / pi)) * 400 / (W + H)
// PARITY UNPINNED for the two IMGLIB kernels: they are restated from TI's published natural-C models (oracle/imglib_open.c),
// and the reference's own sensor code is built against that restatement as the checker (oracle/_ref/libtrikref_oe.so).
//
// What the IMGLIB pair amounts to for the pixels the sensor looks at: IMG_sobel_3x3_8 treats the image as one raster line
// and writes output i + 1 from the 3x3 block whose top-left input is i, so output (r, c) is the Sobel magnitude centred on
// input (r + 1, c); its wrap-around garbage lands in columns 0 and W-1, which the sensor's 16-column margins never read,
// and the last two rows of the work buffer are never written (zero for ever).  "== 0xFF after gt2max(50)" is |H| + |V| > 50.
// So: points = #{(r, c): 0 <= r <= H-3, 16 <= c <= W-16, |H| + |V| > 50 at (r+1, c)} -- no intermediate image at all.
//
// Kernel: one CTA per frame, one thread per column walking down the rows with a three-row sliding window of the separable
// parts (A = l + 2m + r, D = r - l): one byte load per row and thread, the neighbours' bytes by shuffle.  A thread's
// column never changes, so its column sum is count * column -- except that the reference sums each ROW in uint16_t, which
// wraps from W = 384 on; then (WRAP) the row sums are formed exactly, per row, in shared memory.
#include <atomic>
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"

namespace trikb200 {

extern std::atomic<long long> g_launches_edge;
std::atomic<long long> g_launches_edge{0};

template <bool WRAP>
__global__ void __launch_bounds__(1024)
edge_line_kernel(const uint8_t* __restrict__ frames, const long long frameStride, const int lineLength,
                 const int W, const int H, TargetOut* __restrict__ out, const int outStride)
{
  extern __shared__ uint32_t s_rows[];                     // WRAP: column sum of every row
  __shared__ uint32_t s_cnt[32], s_sx[32];
  const int frame = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = (int)(blockDim.x >> 5);
  const int c = 16 + t;                                    // this thread's column
  const bool counted = c <= W - 16;
  const uint8_t* base = frames + (long long)frame * frameStride;
  if (WRAP)
  {
    for (int i = t; i < H; i += blockDim.x) s_rows[i] = 0u;
    __syncthreads();
  }
  // row k of this thread: its own byte, the left neighbour (lane - 1; lane 0 loads column c - 1 itself) and the right one
  // (lane + 1; lane 31 loads c + 1).  (Fetching eight rows ahead of the combine step was measured: slower, 0.60 against
  // 0.47 ms per 4096 x 320x240 -- the walk is bound by its ~25 instructions per row and warp, not by load latency.)
  auto load_row = [&](int k, int& A, int& D)
  {
    const uint8_t* rowp = base + (size_t)k * lineLength;
    const int m = c < W ? (int)__ldg(rowp + c) : 0;
    int l = __shfl_up_sync(0xFFFFFFFFu, m, 1);
    int r = __shfl_down_sync(0xFFFFFFFFu, m, 1);
    if (lane == 0) l = c - 1 < W ? (int)__ldg(rowp + c - 1) : 0;   // c >= 16
    if (lane == 31) r = c + 1 < W ? (int)__ldg(rowp + c + 1) : 0;
    A = l + 2 * m + r;
    D = r - l;
  };
  uint32_t cnt = 0u;
  if (H >= 3)
  {
    int A0, D0, A1, D1;
    load_row(0, A0, D0);
    load_row(1, A1, D1);
    for (int r = 0; r + 2 < H; ++r)
    {
      int A2, D2;
      load_row(r + 2, A2, D2);
      const int Hs = A2 - A0, Vs = D0 + 2 * D1 + D2;
      const bool det = counted && (abs(Hs) + abs(Vs) > 50);
      cnt += det ? 1u : 0u;
      if (WRAP)
      {
        const uint32_t rowPart = __reduce_add_sync(0xFFFFFFFFu, det ? (uint32_t)c : 0u);
        if (lane == 0 && rowPart) atomicAdd(&s_rows[r], rowPart);
      }
      A0 = A1; D0 = D1; A1 = A2; D1 = D2;
    }
  }
  uint32_t sx = cnt * (uint32_t)c;
  cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
  sx = __reduce_add_sync(0xFFFFFFFFu, sx);
  if (lane == 0) { s_cnt[warp] = cnt; s_sx[warp] = sx; }
  __syncthreads();
  if (warp == 0)
  {
    uint32_t points = lane < nwarps ? s_cnt[lane] : 0u;
    uint32_t tx = lane < nwarps ? s_sx[lane] : 0u;
    points = __reduce_add_sync(0xFFFFFFFFu, points);
    tx = __reduce_add_sync(0xFFFFFFFFu, tx);
    if (WRAP)
    {
      uint32_t w = 0u;
      for (int i = lane; i < H; i += 32) w += s_rows[i] & 0xFFFFu;            // each row's sum as the reference's uint16_t
      tx = __reduce_add_sync(0xFFFFFFFFu, w);
    }
    if (lane == 0)
    {
      TargetOut o;
      o.targetX = 0; o.targetY = 0; o.targetSize = 0; o.pad = 0;
      o.detectHue = o.detectHueTolerance = o.detectSat = o.detectSatTolerance = o.detectVal = o.detectValTolerance = 0;
      if (points > 0u)                                                        // :388-403
      {
        const int32_t targetX = (int32_t)(tx / points);
        const uint32_t radius = (uint32_t)ceilf(sqrtf((float)points / 3.1415927f));
        o.targetX = (int8_t)(((targetX - W / 2) * 100 * 2) / W);
        o.targetY = (int8_t)(((0 - H / 2) * 100 * 2) / H);                    // m_targetY is never accumulated
        o.targetSize = (uint8_t)((radius * 100u * 4u) / (uint32_t)(W + H));
      }
      *reinterpret_cast<TargetOut*>(reinterpret_cast<uint8_t*>(out) + (size_t)frame * outStride) = o;
    }
  }
}

cudaError_t launch_edge_line(const uint8_t* frames, long long frameStride, int lineLength, int width, int height,
                             int numFrames, TargetOut* out, int outStride, cudaStream_t stream)
{
  if (numFrames <= 0)
    return cudaSuccess;
  if (width < 32 || height < 4 || width % 32 != 0 || height % 4 != 0 || lineLength < width || outStride < (int)sizeof(TargetOut)
      || width - 31 > 1024)
    return cudaErrorInvalidValue;
  const int cols = width - 31;                              // columns 16 .. W-16
  const int threads = ((cols + 31) / 32) * 32;
  // the reference's uint16_t row sums can wrap once sum_{c=16}^{W-16} c exceeds 65535
  const long long maxRowSum = ((long long)(width - 16) * (width - 15) - 16 * 15) / 2;
  if (maxRowSum > 65535)
    edge_line_kernel<true><<<numFrames, threads, (size_t)height * sizeof(uint32_t), stream>>>(frames, frameStride, lineLength, width, height, out, outStride);
  else
    edge_line_kernel<false><<<numFrames, threads, 0, stream>>>(frames, frameStride, lineLength, width, height, out, outStride);
  ++g_launches_edge;
  return cudaGetLastError();
}

} // namespace trikb200
