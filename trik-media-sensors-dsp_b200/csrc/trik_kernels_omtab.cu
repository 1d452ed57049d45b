// trik_kernels_omtab.cu -- the ov7670 mxn grid-colour sensor (OM) through a full colour-bin table (sm_100a).
//
// ov7670/mxn_sensor/include/internal/cv_ball_detector_seqpass.hpp: pass 1 :252-296 (YUV422P -> RGB888 -> HSV),
// GetImgColor2 :411-452 (per cell a 32x4x4 histogram of (H>>3, S>>6, V>>6), first bin to reach the final maximum).
//
// om_kernel (trik_kernels_grid.cu) evaluates that arithmetic per pixel: 95 warp instructions per pixel, bound by
// instruction issue at 0.09 of the HBM roofline.  Unlike the threshold of the object sensors the 9-bit colour bin
// has no compact form per chroma pair (it changes 14 times on average along Y), but it does not
// depend on any argument either: it is ONE fixed function of three bytes.  So it is tabulated completely:
//
//   om_bin_table_kernel  runs hsv_pair() (the code om_kernel runs) on all 2^24 (Y,U,V) once per device and stores
//                        bin * 4 as uint16 at  [U][V][Y]  -- 32 MB, resident in the 126 MB L2.  Y is the fastest
//                        index, so the 16 luma values around a pixel's own share its 32-byte sector and a flat
//                        region of the picture keeps hitting the few sectors of its colour in L1.
//   om_table_kernel      one CTA per frame and cell row as om_kernel, a thread owns a 16-pixel column chunk and walks
//                        down the rows; the per-pixel arithmetic is replaced by one 16-bit gather (16 independent
//                        loads in flight per thread), followed by the update of the cell histograms in shared memory.
//
// Bit-identical to om_kernel by construction (the table IS that arithmetic); tests/test_omtab_gpu.py compares the
// table with the oracle on all 2^24 inputs and the sensor through both paths with the oracle on frames.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"

namespace trikb200 {

extern std::atomic<long long> g_launches_omtab;
std::atomic<long long> g_launches_omtab{0};

constexpr int OMT_BINS = 512;
constexpr int OMT_MAX_GROUP = 12;      // cells of one cell-row whose histograms live in shared memory at once

// ---------------------------------------------------------------------------------------------
// table construction
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t om_bin_of_hsv(uint32_t hsv)
{
  return ((hsv & 0xFFu) >> 3 << 4) | (((hsv >> 8) & 0xFFu) >> 6 << 2) | ((hsv >> 16) >> 6);
}

__global__ void __launch_bounds__(256)
om_bin_table_kernel(uint16_t* __restrict__ table)
{
  __shared__ HueLutEntry s_lutHue[256];
  __shared__ uint16_t s_lut43[256];
  __shared__ uint16_t s_lut255[256];
  fill_div_luts(s_lut43, s_lut255);
  fill_hue_lut(s_lutHue);
  __syncthreads();
  // one CTA per chroma pair, one thread per luma; index = (U << 16) | (V << 8) | Y
  const uint32_t half = blockIdx.x;                       // V | U << 8: the chroma bytes as they lie in the plane
  const uint32_t y = threadIdx.x;
  uint32_t h0, h1;
  hsv_pair(y | (y << 16), half | (half << 16), coef_planar0(), s_lutHue, s_lut255, h0, h1);
  table[((size_t)half << 8) + y] = (uint16_t)(om_bin_of_hsv(h0) << 2);
}

cudaError_t launch_om_bin_table(uint16_t* table, cudaStream_t stream)
{
  om_bin_table_kernel<<<65536, 256, 0, stream>>>(table);
  ++g_launches_omtab;
  return cudaGetLastError();
}

// probe for the parity test: out[i] = bin of (Y = i & 255, U = (i >> 8) & 255, V = i >> 16), read from the table
__global__ void om_table_probe_kernel(const uint16_t* __restrict__ table, uint32_t first, uint32_t count, uint32_t* __restrict__ out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t idx = first + i;
  const uint32_t y = idx & 0xFFu, u = (idx >> 8) & 0xFFu, v = (idx >> 16) & 0xFFu;
  out[i] = (uint32_t)table[((size_t)(v | (u << 8)) << 8) + y] >> 2;
}

cudaError_t launch_om_table_probe(const uint16_t* table, uint32_t first, uint32_t count, uint32_t* out, cudaStream_t stream)
{
  if (!count) return cudaSuccess;
  om_table_probe_kernel<<<(count + 255u) / 256u, 256, 0, stream>>>(table, first, count, out);
  ++g_launches_omtab;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// the sensor pass
// ---------------------------------------------------------------------------------------------
// table gather: read-only path, allocating in L1 (the frame stream does not: ld_stream)
__device__ __forceinline__ uint32_t ld_bin(const uint16_t* __restrict__ table, uint32_t index)
{
  uint16_t r;
  asm volatile("ld.global.nc.L1::evict_last.u16 %0, [%1];" : "=h"(r) : "l"(table + index));
  return (uint32_t)r;
}

__device__ __forceinline__ void hist_add(uint32_t addr, uint32_t cnt)
{
  asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(cnt) : "memory");
}

// The histogram update.  What limits this kernel is the L1 / shared-memory data pipe (table gathers, the frame stream
// and the shared-memory atomics all pass through it; ncu: 85 % busy with issue slots half idle), so
//   * a warp works on a compact tile (c chunks x 32/c rows, every row segment a whole 32-byte sector or more) that
//     mostly lies inside one cell: the lanes of a gather then hit the few table lines of that cell's colour;
//   * a thread keeps ONE pending {cell, bin} slot down its column and only counts matches in a register; any other
//     pixel is added on its own.  On a region of one colour with scattered outliers that is one atomic per outlier
//     (plain run-length coding pays two: the run it interrupts and the outlier itself).  A row without a single match
//     re-seeds the slot, so the thread follows the picture when the colour under it changes;
//   * the hot loop counts only (dense 32-bank count arrays).  The reference's "first bin to reach the final maximum"
//     (:434-440) needs the bins' last raster positions only to order bins whose FINAL counts tie at the maximum --
//     rare -- so they are recovered afterwards, by a second pass over the cell row, only for such cells and bins.
// one cell row of one frame, by the whole CTA
__device__ __forceinline__ void
om_table_item(const Geometry& g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
              const int paramStride, const uint16_t* __restrict__ table, const uint32_t* __restrict__ colorTable,
              int32_t* __restrict__ out, const int frame, const int cellRow, const int log2c)
{
  extern __shared__ uint32_t s_dyn[];      // [group + 1][512] counts (slot `group` is a sink), then [group][512] last positions
  __shared__ uint32_t s_tieMax[OMT_MAX_GROUP];            // per cell: the maximum count if several bins share it, else 0
  __shared__ uint32_t s_anyTie;

  const FrameParams p = params[(size_t)frame * paramStride];
  const int M = (int)p.gridRows, N = (int)p.gridCols;
  if (cellRow >= M || N <= 0)
    return;                                               // uniform per CTA

  const int W = g.width, H = g.height;
  const int ws = W / N, hs = H / M;                       // m_widthStep, m_heightStep (:587-588); remainders ignored
  const int t = threadIdx.x;
  const int r0 = cellRow * hs, r1 = r0 + hs;
  const int group = N < OMT_MAX_GROUP ? N : OMT_MAX_GROUP;
  uint32_t* s_cnt = s_dyn;
  uint32_t* s_pos = s_dyn + (group + 1) * OMT_BINS;
  const uint32_t sBase = (uint32_t)__cvta_generic_to_shared(s_cnt);

  const uint8_t* fbase = frames + (size_t)frame * g.frameStride;
  const size_t chromaOfs = (size_t)H * g.lineLength;
  int32_t* frameOut = out + (size_t)frame * 100;
  const int warp = t >> 5, lane = t & 31, nwarps = (int)(blockDim.x >> 5);

  // tiles: c = 2^log2c chunks of 16 pixels wide, 32 / c rows high; a warp keeps its strip of c chunks and walks down
  const int cpr = W >> 4;
  const int strips = cpr >> log2c;                        // the launcher picks c so that it divides cpr
  const int spp = nwarps < strips ? nwarps : strips;      // strips worked on at a time
  const int rowGroups = nwarps / spp;                     // warps sharing one strip take alternate tiles
  const int tileRows = 32 >> log2c;
  const int sx = warp % spp, rg = warp / spp;

  for (int g0 = 0; g0 < N; g0 += group)
  {
    const int g1 = min(g0 + group, N);
    __syncthreads();
    for (int i = t; i < (group + 1) * OMT_BINS / 4; i += blockDim.x)
      reinterpret_cast<uint4*>(s_cnt)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (t == 0) s_anyTie = 0u;
    __syncthreads();

    const int cLo = g0 * ws, cHi = g1 * ws;               // columns of cells g0..g1-1: [cLo, cHi)
    if (ws > 0 && rg < rowGroups)
    for (int strip = sx; strip < strips; strip += spp)
    {
      const int col0 = ((strip << log2c) + (lane & ((1 << log2c) - 1))) << 4;
      if (col0 >= cHi || col0 + 16 <= cLo)
        continue;                                         // this thread's 16-pixel chunk touches none of these cells
      // shared-memory address of the histogram each of the 16 pixel slots counts into (the sink for columns outside);
      // one division, then the cell boundaries are met walking right (ws >= 1: at most one per column)
      uint32_t hb[16];
      {
        const int cFirst = col0 > cLo ? col0 : cLo;
        int cell = cFirst / ws - g0;
        int next = (cell + g0 + 1) * ws;
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
          const int col = col0 + k;
          if (col >= next) { ++cell; next += ws; }
          const int c = (col >= cLo && col < cHi) ? cell : group;
          hb[k] = sBase + (uint32_t)c * (OMT_BINS * 4u);
        }
      }
      const uint8_t* base = fbase + (size_t)col0;
      const int rowStep = tileRows * rowGroups;
      int row = r0 + rg * tileRows + (lane >> log2c);
      uint4 lu = make_uint4(0u, 0u, 0u, 0u), ch = lu;
      if (row < r1)
      {
        const uint8_t* ptr = base + (size_t)row * g.lineLength;
        lu = ld_stream(ptr);
        ch = ld_stream(ptr + chromaOfs);
      }
      uint32_t pend = 0xFFFFFFFFu, cnt = 0u;
      while (row < r1)
      {
        const uint32_t L[4] = {lu.x, lu.y, lu.z, lu.w};
        const uint32_t Cw[4] = {ch.x, ch.y, ch.z, ch.w};
        uint32_t e[16];
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
          // chroma word [V0 U0 V1 U1]: the half-word of pair k is V | U << 8, the table's row; index = half << 8 | Y
          // built by one byte permute from the luma word and the isolated half-word
          const uint32_t half = (k & 1) ? (Cw[k >> 1] >> 16) : (Cw[k >> 1] & 0xFFFFu);
          const uint32_t i0 = __byte_perm(L[k >> 1], half, (k & 1) ? 0x6542 : 0x6540);
          const uint32_t i1 = __byte_perm(L[k >> 1], half, (k & 1) ? 0x6543 : 0x6541);
          e[2 * k] = ld_bin(table, i0);
          e[2 * k + 1] = ld_bin(table, i1);
        }
        // next tile's chunk while the gathers are in flight
        row += rowStep;
        if (row < r1)
        {
          const uint8_t* ptr = base + (size_t)row * g.lineLength;
          lu = ld_stream(ptr);
          ch = ld_stream(ptr + chromaOfs);
        }
        const uint32_t before = cnt;
        if (pend == 0xFFFFFFFFu) pend = hb[0] + e[0];
#pragma unroll
        for (int k = 0; k < 16; k += 2)
        {
          // the two pixels of a chroma pair: one test in the usual case that both match
          const uint32_t i0 = hb[k] + e[k], i1 = hb[k + 1] + e[k + 1];
          if (i0 == pend && i1 == pend)
            cnt += 2u;
          else
          {
            if (i0 == pend) ++cnt; else hist_add(i0, 1u);
            if (i1 == pend) ++cnt; else hist_add(i1, 1u);
          }
        }
        if (cnt == before)
        {
          // not one match in this row: the colour under this thread has changed, follow it.  The last pixel has been
          // counted already, so the new slot starts at 0.
          if (cnt) hist_add(pend, cnt);
          pend = hb[15] + e[15];
          cnt = 0u;
        }
      }
      if (cnt) hist_add(pend, cnt);
    }
    __syncthreads();

    // per cell: the bin with the maximum count.  One warp per cell.  Several bins with that count: note it for pass 2.
    for (int c = warp; c < g1 - g0; c += nwarps)
    {
      uint32_t best = 0u, bestBin = 0u, ties = 0u;
      for (int b = lane; b < OMT_BINS; b += 32)
      {
        const uint32_t n = s_cnt[c * OMT_BINS + b];
        if (n > best) { best = n; bestBin = (uint32_t)b; ties = 1u; }
        else if (n == best) ++ties;
      }
      const uint32_t m = __reduce_max_sync(0xFFFFFFFFu, best);
      const uint32_t nTied = __reduce_add_sync(0xFFFFFFFFu, best == m ? ties : 0u);
      const uint32_t bin = __reduce_min_sync(0xFFFFFFFFu, best == m ? bestBin : 0xFFFFFFFFu);
      if (lane == 0)
      {
        if (m == 0u || nTied == 1u)
        {
          // empty cell (every bin ties at 0 and none is ever "first": bin 0) or a single winner
          frameOut[cellRow * N + g0 + c] = (int32_t)colorTable[m == 0u ? 0u : bin];
          s_tieMax[c] = 0u;
        }
        else
        {
          s_tieMax[c] = m;
          s_anyTie = 1u;
        }
      }
    }
    __syncthreads();
    if (s_anyTie == 0u)
      continue;                                           // uniform per CTA

    // pass 2 (rare): last raster position of every bin that ties at the maximum of its cell
    for (int i = t; i < group * OMT_BINS; i += blockDim.x)
      s_pos[i] = 0u;
    __syncthreads();
    {
      const int span = cHi - cLo;
      for (int i = t; i < hs * span; i += blockDim.x)
      {
        const int row = r0 + i / span, col = cLo + i % span;
        const int c = col / ws - g0;
        const uint32_t m = s_tieMax[c];
        if (m == 0u) continue;
        const uint8_t* ptr = fbase + (size_t)row * g.lineLength;
        const uint32_t half = *reinterpret_cast<const uint16_t*>(ptr + chromaOfs + (col & ~1));
        const uint32_t bin = ld_bin(table, (half << 8) | ptr[col]) >> 2;
        if (s_cnt[c * OMT_BINS + bin] == m)
          atomicMax(&s_pos[c * OMT_BINS + bin], (uint32_t)row * (uint32_t)W + (uint32_t)col);
      }
    }
    __syncthreads();
    // among the tied bins the first to reach the maximum is the one whose last pixel comes first
    for (int c = warp; c < g1 - g0; c += nwarps)
    {
      const uint32_t m = s_tieMax[c];
      if (m == 0u) continue;                              // uniform per warp
      unsigned long long best = ~0ull;
      for (int b = lane; b < OMT_BINS; b += 32)
      {
        const unsigned long long key = ((unsigned long long)s_pos[c * OMT_BINS + b] << 32) | (unsigned long long)b;
        if (s_cnt[c * OMT_BINS + b] == m && key < best) best = key;
      }
      for (int off = 16; off > 0; off >>= 1)
      {
        const unsigned long long o = __shfl_down_sync(0xFFFFFFFFu, best, off);
        if (o < best) best = o;
      }
      if (lane == 0)
        frameOut[cellRow * N + g0 + c] = (int32_t)colorTable[(uint32_t)best & (OMT_BINS - 1)];
    }
  }
}

__global__ void __launch_bounds__(512, 2)
om_table_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
                const int paramStride, const uint16_t* __restrict__ table, const uint32_t* __restrict__ colorTable,
                int32_t* __restrict__ out, const int maxGridRows, const int log2c)
{
  const int frame = blockIdx.x / maxGridRows;
  om_table_item(g, frames, params, paramStride, table, colorTable, out, frame, blockIdx.x - frame * maxGridRows, log2c);
}

// The cell rows the majority pass (trik_kernels_ommaj.cu) could not decide: persistent CTAs take them from its list.
// (The list counters are two, used by alternate batches; the majority pass zeroes the one it does not use.)
__global__ void __launch_bounds__(512, 2)
om_table_list_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
                     const int paramStride, const uint16_t* __restrict__ table, const uint32_t* __restrict__ colorTable,
                     int32_t* __restrict__ out, const int maxGridRows, const int log2c,
                     const int* __restrict__ list, const int* __restrict__ listCount)
{
  const int n = *listCount;
  for (int i = blockIdx.x; i < n; i += gridDim.x)
  {
    const int item = list[i];
    const int frame = item / maxGridRows;
    om_table_item(g, frames, params, paramStride, table, colorTable, out, frame, item - frame * maxGridRows, log2c);
  }
}

static int g_omtWarps = 6;
void set_om_table_threads(int threads) { g_omtWarps = threads > 0 ? (threads + 31) / 32 : 6; }

cudaError_t launch_om_table(const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                            int paramStride, const uint16_t* table, const uint32_t* colorTable, int32_t* out,
                            int maxGridRows, int maxGridCols, cudaStream_t stream, const int* fbList, const int* fbCount, int smCount)
{
  if (numFrames <= 0 || maxGridRows <= 0)
    return cudaSuccess;
  const int cpr = g.width / 16;
  if (cpr <= 0 || (cpr & 1))
    return cudaErrorInvalidValue;
  const int target = g_omtWarps > 16 ? 16 : g_omtWarps;
  // tile width: the narrowest (at least two chunks = one 32-byte sector per row) that divides the row and leaves no
  // more strips than warps; else the widest that divides the row
  int log2c = 1;
  while (log2c < 5 && (cpr >> log2c) > target && cpr % (2 << log2c) == 0)
    ++log2c;
  const int strips = cpr >> log2c;
  int warps = strips;
  if (strips >= target)
    warps = target;
  else
  {
    const int tilesDown = (g.height + (32 >> log2c) - 1) / (32 >> log2c);
    int rowGroups = target / strips;
    if (rowGroups > tilesDown) rowGroups = tilesDown;
    warps = strips * (rowGroups < 1 ? 1 : rowGroups);
  }
  const int group = maxGridCols < OMT_MAX_GROUP ? maxGridCols : OMT_MAX_GROUP;
  const size_t smem = (size_t)(2 * group + 1) * OMT_BINS * sizeof(uint32_t);
  const long long grid = (long long)numFrames * maxGridRows;
  if (grid > 0x7FFFFFFFLL)
    return cudaErrorInvalidValue;
  if (smem > 48u * 1024u)
    cudaFuncSetAttribute(om_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (2 * OMT_MAX_GROUP + 1) * OMT_BINS * (int)sizeof(uint32_t));
  if (fbList)
  {
    // list mode: only the cell rows the majority pass left undecided, by at most two CTAs per SM
    if (smem > 48u * 1024u)
      cudaFuncSetAttribute(om_table_list_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (2 * OMT_MAX_GROUP + 1) * OMT_BINS * (int)sizeof(uint32_t));
    static const bool debugList = getenv("TRIKB200_OMJ_DEBUG") != nullptr;       // diagnostics: how long is the list?
    if (debugList)
    {
      int c = -1;
      cudaStreamSynchronize(stream);
      cudaMemcpy(&c, fbCount, sizeof(c), cudaMemcpyDeviceToHost);
      fprintf(stderr, "om list: %d of %lld cell rows undecided\n", c, grid);
    }
    // as many CTAs as are resident at once (the histogram kernel lives on residency: ~14 KB of shared memory each)
    int perSm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, om_table_list_kernel, warps * 32, smem) != cudaSuccess || perSm < 1)
      perSm = 2;
    const long long resident = (long long)perSm * smCount;
    const unsigned ctas = (unsigned)(grid < resident ? grid : resident);
    om_table_list_kernel<<<ctas, warps * 32, smem, stream>>>(g, frames, params, paramStride, table, colorTable, out,
                                                            maxGridRows, log2c, fbList, fbCount);
    ++g_launches_omtab;
    return cudaGetLastError();
  }
  om_table_kernel<<<(unsigned)grid, warps * 32, smem, stream>>>(g, frames, params, paramStride, table, colorTable, out,
                                                                maxGridRows, log2c);
  ++g_launches_omtab;
  return cudaGetLastError();
}

} // namespace trikb200
