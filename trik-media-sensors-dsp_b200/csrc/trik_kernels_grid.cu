// trik_kernels_grid.cu -- the ov7670 mxn grid-colour sensor (OM) and the ov7670 object sensor (OO).
//
// OM  ov7670/mxn_sensor/include/internal/cv_ball_detector_seqpass.hpp
//       pass 1 :252-296, GetImgColor2 :411-452, HSVtoRGB :480-517, grid loop :585-618
//     Per cell the reference builds a 32x4x4 histogram of (H>>3, S>>6, V>>6) in raster order and
//     keeps "the first bin to reach the final maximum".  For a pure counting histogram that is:
//     among the bins whose final count equals the maximum, the one whose LAST pixel comes first.
//     So the kernel keeps (count, last raster position) per bin -- atomicAdd + atomicMax in shared
//     memory -- and the order of the additions no longer matters.
//
// OO  ov7670/object_sensor/include/internal/
//       cv_bitmap_builder_reference.hpp:107-217  threshold -> 4x4 "metapixel" bitmap
//       cv_clusterizer_reference.hpp:38-202      raster labelling with a one-level equivalence table
//       cv_ball_detector_seqpass.hpp:569-597     the eight largest clusters -> targets
//     The labelling is order dependent by construction (non-transitive equivalences, the first
//     pixel of a label is not counted, merge in label order, std::sort tie order), so it is
//     replayed step for step by one lane per frame; frames run in parallel.
#include <atomic>
#include <cstdlib>
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"

namespace trikb200 {

extern std::atomic<long long> g_launches_grid;
std::atomic<long long> g_launches_grid{0};

// threads per CTA = chunksPerRow * rowsPerIteration, about `target` threads, a multiple of 32 when possible
static int rows_per_iteration(int cpr, int target)
{
  int k = (target + cpr - 1) / cpr;
  for (int j = 0; j < 16; ++j)
    if ((cpr * (k + j)) % 32 == 0 && cpr * (k + j) <= 1024)
      return k + j;
  if (cpr * k > 1024) k = 1024 / cpr;
  return k < 1 ? 1 : k;
}

// =============================================================================================
// OM
// =============================================================================================
constexpr int OM_BINS = 512;
constexpr int OM_MAX_GROUP = 12;      // cells of one cell-row whose histograms live in shared memory at once

__global__ void __launch_bounds__(1024)
om_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
          const int paramStride, const uint32_t* __restrict__ colorTable, int32_t* __restrict__ out,
          const int maxGridRows, const int cpr, const int rpi)
{
  extern __shared__ uint32_t s_dyn[];                     // [group][512] counts, then [group][512] last positions
  __shared__ HueLutEntry s_lutHue[256];
  __shared__ uint16_t s_lut43[256];
  __shared__ uint16_t s_lut255[256];
  __shared__ unsigned long long s_best[32];
  __shared__ uint32_t s_bestBin[32];

  const int frame = blockIdx.x / maxGridRows;
  const int cellRow = blockIdx.x - frame * maxGridRows;
  const FrameParams p = params[(size_t)frame * paramStride];
  const int M = (int)p.gridRows, N = (int)p.gridCols;
  if (cellRow >= M || N <= 0)
    return;                                               // uniform per CTA
  fill_div_luts(s_lut43, s_lut255);
  fill_hue_lut(s_lutHue);

  const int W = g.width, H = g.height;
  const int ws = W / N, hs = H / M;                       // m_widthStep, m_heightStep (:587-588); remainders ignored
  const int t = threadIdx.x;
  const int cc = t % cpr, rr = t / cpr;
  const int col0 = cc * 16;
  const int r0 = cellRow * hs, r1 = r0 + hs;
  const int group = N < OM_MAX_GROUP ? N : OM_MAX_GROUP;
  uint32_t* s_cnt = s_dyn;
  uint32_t* s_pos = s_dyn + group * OM_BINS;

  const uint8_t* base = frames + (size_t)frame * g.frameStride + (size_t)col0;
  const size_t chromaOfs = (size_t)H * g.lineLength;
  int32_t* frameOut = out + (size_t)frame * 100;

  for (int g0 = 0; g0 < N; g0 += group)
  {
    const int g1 = min(g0 + group, N);
    __syncthreads();
    for (int i = t; i < 2 * group * OM_BINS; i += blockDim.x)
      s_dyn[i] = 0u;
    __syncthreads();

    // does this thread's 16-pixel chunk touch the columns of cells g0..g1-1 ?
    const int cLo = g0 * ws, cHi = g1 * ws;               // [cLo, cHi)
    if (ws > 0 && col0 < cHi && col0 + 16 > cLo)
    {
      for (int row = r0 + rr; row < r1; row += rpi)
      {
        const uint8_t* ptr = base + (size_t)row * g.lineLength;
        const uint4 lu = ld_stream(ptr);
        const uint4 ch = ld_stream(ptr + chromaOfs);
        const uint32_t L[4] = {lu.x, lu.y, lu.z, lu.w};
        const uint32_t Cw[4] = {ch.x, ch.y, ch.z, ch.w};
        int pendIdx = -1;                                 // run-length compression of identical (cell, bin)
        uint32_t pendCnt = 0, pendPos = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
          const uint32_t yy = __byte_perm(L[k >> 1], 0u, (k & 1) ? 0x4342 : 0x4140);
          uint32_t hsv[2];
          hsv_pair(yy, Cw[k >> 1], (k & 1) ? coef_planar1() : coef_planar0(), s_lutHue, s_lut255, hsv[0], hsv[1]);
#pragma unroll
          for (int e = 0; e < 2; ++e)
          {
            const int col = col0 + 2 * k + e;
            if (col < cLo || col >= cHi)
              continue;
            const int cell = col / ws - g0;
            const uint32_t bin = ((hsv[e] & 0xFFu) >> 3 << 4) | (((hsv[e] >> 8) & 0xFFu) >> 6 << 2) | ((hsv[e] >> 16) >> 6);
            const int idx = cell * OM_BINS + (int)bin;
            const uint32_t pos = (uint32_t)row * (uint32_t)W + (uint32_t)col;
            if (idx == pendIdx)
            {
              ++pendCnt; pendPos = pos;
            }
            else
            {
              if (pendIdx >= 0)
              {
                atomicAdd(&s_cnt[pendIdx], pendCnt);
                atomicMax(&s_pos[pendIdx], pendPos);
              }
              pendIdx = idx; pendCnt = 1; pendPos = pos;
            }
          }
        }
        if (pendIdx >= 0)
        {
          atomicAdd(&s_cnt[pendIdx], pendCnt);
          atomicMax(&s_pos[pendIdx], pendPos);
        }
      }
    }
    __syncthreads();

    // per cell: max count, ties -> earliest last position, full tie (empty cell) -> bin 0
    for (int c = 0; c < g1 - g0; ++c)
    {
      unsigned long long best = 0ull;
      uint32_t bestBin = 0xFFFFFFFFu;
      for (int b = t; b < OM_BINS; b += blockDim.x)
      {
        const unsigned long long key = ((unsigned long long)s_cnt[c * OM_BINS + b] << 32)
                                     | (unsigned long long)(0xFFFFFFFFu - s_pos[c * OM_BINS + b]);
        if (key > best || (key == best && (uint32_t)b < bestBin))
        {
          best = key; bestBin = (uint32_t)b;
        }
      }
      const unsigned am = __activemask();
      for (int off = 16; off > 0; off >>= 1)
      {
        const unsigned long long ok = __shfl_down_sync(am, best, off);
        const uint32_t ob = __shfl_down_sync(am, bestBin, off);
        if (ok > best || (ok == best && ob < bestBin))
        {
          best = ok; bestBin = ob;
        }
      }
      const int warp = t >> 5, lane = t & 31, nwarps = (blockDim.x + 31) >> 5;
      if (lane == 0) { s_best[warp] = best; s_bestBin[warp] = bestBin; }
      __syncthreads();
      if (t == 0)
      {
        for (int w2 = 1; w2 < nwarps; ++w2)
          if (s_best[w2] > best || (s_best[w2] == best && s_bestBin[w2] < bestBin))
          {
            best = s_best[w2]; bestBin = s_bestBin[w2];
          }
        if (bestBin >= (uint32_t)OM_BINS) bestBin = 0;
        frameOut[cellRow * N + g0 + c] = (int32_t)colorTable[bestBin];
      }
      __syncthreads();
    }
  }
}

cudaError_t launch_om(const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                      int paramStride, const uint32_t* colorTable, int32_t* out, int maxGridRows, int maxGridCols,
                      cudaStream_t stream)
{
  if (numFrames <= 0 || maxGridRows <= 0)
    return cudaSuccess;
  const int cpr = g.width / 16;
  if (cpr <= 0 || cpr > 1024)
    return cudaErrorInvalidValue;
  const int k = rows_per_iteration(cpr, 192);
  const int threads = cpr * k;
  const int group = maxGridCols < OM_MAX_GROUP ? maxGridCols : OM_MAX_GROUP;
  const size_t smem = (size_t)2 * group * OM_BINS * sizeof(uint32_t);
  const long long grid = (long long)numFrames * maxGridRows;
  if (grid > 0x7FFFFFFFLL)
    return cudaErrorInvalidValue;
  cudaFuncSetAttribute(om_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * OM_MAX_GROUP * OM_BINS * (int)sizeof(uint32_t));
  om_kernel<<<(unsigned)grid, threads, smem, stream>>>(g, frames, params, paramStride, colorTable, out, maxGridRows, cpr, k);
  ++g_launches_grid;
  return cudaGetLastError();
}

// =============================================================================================
// OO, step 1: threshold -> metapixel bitmap
// =============================================================================================
// bitmap[(row/4) * (W/4) + col/4], bit (row%4)*4 + col%4 = det   (cv_bitmap_builder_reference.hpp:163-187)
__global__ void __launch_bounds__(1024)
oo_bitmap_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
                 const int paramStride, uint16_t* __restrict__ bitmaps, const int cpr, const int rpi)
{
  __shared__ HueLutEntry s_lutHue[256];
  __shared__ uint16_t s_lut43[256];
  __shared__ uint16_t s_lut255[256];
  fill_div_luts(s_lut43, s_lut255);
  fill_hue_lut(s_lutHue);
  __syncthreads();

  const int slabs = gridDim.y;
  const int frame = blockIdx.x;
  const FrameParams p = params[(size_t)frame * paramStride];
  const HsvBounds bd = make_bounds(p.from, p.to);
  const int t = threadIdx.x;
  const int cc = t % cpr, rr = t / cpr;
  const int bw = g.width / 4, bh = g.height / 4;
  const int mrPerSlab = (bh + slabs - 1) / slabs;
  const int mr0 = blockIdx.y * mrPerSlab, mr1 = min(mr0 + mrPerSlab, bh);
  const uint8_t* base = frames + (size_t)frame * g.frameStride + (size_t)cc * 16u;
  const size_t chromaOfs = (size_t)g.height * g.lineLength;
  uint16_t* bm = bitmaps + (size_t)frame * bw * bh;

  for (int mr = mr0 + rr; mr < mr1; mr += rpi)
  {
    uint32_t meta[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
      const uint8_t* ptr = base + (size_t)(mr * 4 + r) * g.lineLength;
      const uint4 lu = ld_stream(ptr);
      const uint4 ch = ld_stream(ptr + chromaOfs);
      const uint32_t L[4] = {lu.x, lu.y, lu.z, lu.w};
      const uint32_t Cw[4] = {ch.x, ch.y, ch.z, ch.w};
#pragma unroll
      for (int k = 0; k < 8; ++k)
      {
        const uint32_t yy = __byte_perm(L[k >> 1], 0u, (k & 1) ? 0x4342 : 0x4140);
        const uint32_t det = detect_pair_bits(yy, Cw[k >> 1], (k & 1) ? coef_planar1() : coef_planar0(),
                                              s_lutHue, s_lut255, bd, p.expected);
        meta[k >> 1] |= det << (r * 4 + (k & 1) * 2);
      }
    }
    uint2 v;
    v.x = meta[0] | (meta[1] << 16);
    v.y = meta[2] | (meta[3] << 16);
    *reinterpret_cast<uint2*>(bm + (size_t)mr * bw + cc * 4) = v;
  }
}

// =============================================================================================
// OO, step 2: labelling + ranking + targets, one warp per frame, lane 0 sequences
// =============================================================================================
struct OoCluster { int32_t x, y, size; };

// libstdc++ 13 std::sort (bits/stl_algo.h: __introsort_loop, __final_insertion_sort, _S_threshold = 16;
// bits/stl_heap.h for the depth-limit fallback) with compareTargetBySize (cv_clusterizer_reference.hpp:28-31).
// Restated because the order of equal sizes is observable in the eight reported targets.
__device__ __forceinline__ bool oo_before(const OoCluster& a, const OoCluster& b) { return a.size > b.size; }
__device__ __forceinline__ void oo_swap(OoCluster& a, OoCluster& b) { const OoCluster t = a; a = b; b = t; }

__device__ void oo_push_heap(OoCluster* first, int hole, int top, OoCluster value)
{
  int parent = (hole - 1) / 2;
  while (hole > top && oo_before(first[parent], value))
  {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
__device__ void oo_adjust_heap(OoCluster* first, int hole, int len, OoCluster value)
{
  const int top = hole;
  int second = hole;
  while (second < (len - 1) / 2)
  {
    second = 2 * (second + 1);
    if (oo_before(first[second], first[second - 1]))
      second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2)
  {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  oo_push_heap(first, hole, top, value);
}
__device__ void oo_heap_sort(OoCluster* first, int len)
{
  if (len >= 2)
    for (int parent = (len - 2) / 2;; --parent)
    {
      oo_adjust_heap(first, parent, len, first[parent]);
      if (parent == 0)
        break;
    }
  int last = len;
  while (last > 1)
  {
    --last;
    const OoCluster value = first[last];
    first[last] = first[0];
    oo_adjust_heap(first, 0, last, value);
  }
}
__device__ void oo_linear_insert(OoCluster* a, int last)
{
  const OoCluster val = a[last];
  int next = last - 1;
  while (oo_before(val, a[next]))
  {
    a[last] = a[next];
    last = next;
    --next;
  }
  a[last] = val;
}
__device__ void oo_insertion_sort(OoCluster* a, int first, int last)
{
  if (first == last) return;
  for (int i = first + 1; i != last; ++i)
  {
    if (oo_before(a[i], a[first]))
    {
      const OoCluster val = a[i];
      for (int j = i; j > first; --j)
        a[j] = a[j - 1];
      a[first] = val;
    }
    else
      oo_linear_insert(a, i);
  }
}
__device__ void oo_std_sort(OoCluster* a, int n)
{
  if (n <= 0) return;
  int lg = 0;
  while ((n >> (lg + 1)) != 0) ++lg;
  // __introsort_loop, recursion on the right part turned into an explicit stack
  int stackFirst[64], stackLast[64], stackDepth[64];
  int sp = 0;
  stackFirst[0] = 0; stackLast[0] = n; stackDepth[0] = 2 * lg; sp = 1;
  while (sp > 0)
  {
    --sp;
    const int first = stackFirst[sp];
    int last = stackLast[sp], depth = stackDepth[sp];
    // the reference recursion runs introsort_loop(cut,last) to completion BEFORE continuing with
    // [first,cut); the two ranges are disjoint, so the order of processing cannot change the result
    while (last - first > 16)
    {
      if (depth == 0)
      {
        oo_heap_sort(a + first, last - first);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      {                                                   // __move_median_to_first(first, first+1, mid, last-1)
        OoCluster &r = a[first], &x = a[first + 1], &y = a[mid], &z = a[last - 1];
        if (oo_before(x, y))
        {
          if (oo_before(y, z)) oo_swap(r, y);
          else if (oo_before(x, z)) oo_swap(r, z);
          else oo_swap(r, x);
        }
        else if (oo_before(x, z)) oo_swap(r, x);
        else if (oo_before(y, z)) oo_swap(r, z);
        else oo_swap(r, y);
      }
      int lo = first + 1, hi = last;                      // __unguarded_partition(first+1, last, pivot = first)
      for (;;)
      {
        while (oo_before(a[lo], a[first])) ++lo;
        --hi;
        while (oo_before(a[first], a[hi])) --hi;
        if (!(lo < hi))
          break;
        oo_swap(a[lo], a[hi]);
        ++lo;
      }
      const int cut = lo;
      if (sp < 64)
      {
        stackFirst[sp] = cut; stackLast[sp] = last; stackDepth[sp] = depth; ++sp;
      }
      last = cut;
    }
  }
  if (n > 16)                                             // __final_insertion_sort
  {
    oo_insertion_sort(a, 0, 16);
    for (int i = 16; i != n; ++i)
      oo_linear_insert(a, i);
  }
  else
    oo_insertion_sort(a, 0, n);
}

struct ObjOut {                                           // == TRIKB200_ObjOutArgsAlg (36 bytes)
  int8_t  t[24];                                          // XDAS_Target target[8] = {int8 x, int8 y, uint8 size}
  uint16_t detectHue, detectHueTolerance, detectSat, detectSatTolerance, detectVal, detectValTolerance;
};
static_assert(sizeof(ObjOut) == 36, "ObjOutArgsAlg layout");

// One warp per frame.  The 32 lanes test 32 metapixels at a time (popcount > 2) and vote; lane 0 then
// walks only the set bits in raster order and replays the reference's labelling on them.  The label
// tables live in shared memory when they fit (tablesInSmem), else in the global scratch.
// One warp per frame, OO_WARPS_PER_CTA independent warps per CTA (they share nothing but the launch): the walk is
// a chain of dependent shuffles and small table updates, so what hides its latency is the number of resident warps.
constexpr int OO_WARPS_PER_CTA = 2;
constexpr int OO_RING_ROWS = 8, OO_RING_AHEAD = 7;        // bitmap rows staged in shared memory / rows fetched ahead
__device__ __forceinline__ void red_add_global(int32_t* p, int v)
{
  asm volatile("red.global.add.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__global__ void __launch_bounds__(32 * OO_WARPS_PER_CTA)
oo_cluster_kernel(const Geometry g, const uint16_t* __restrict__ bitmaps, OoCluster* __restrict__ clustersAll,
                  uint16_t* __restrict__ equalAll, const int maxLabels, const int tablesInSmem,
                  ObjOut* __restrict__ out, int* __restrict__ labelCounts, const int numFrames, const int smemPerWarp,
                  const int tailCap, const int tailOfs, const int ringOfs)
{
  extern __shared__ __align__(16) uint8_t s_all[];
  const int frame = blockIdx.x * OO_WARPS_PER_CTA + (threadIdx.x >> 5);
  if (frame >= numFrames)
    return;
  uint8_t* const s_raw = s_all + (size_t)(threadIdx.x >> 5) * smemPerWarp;
  const int lane = threadIdx.x & 31;
  const int bw = g.width / 4, bh = g.height / 4;
  const uint16_t* bm = bitmaps + (size_t)frame * bw * bh;
  // shared layout: [clusters (12 B each)] [equal (2 B each)] [two label rows of bw+? u16]
  OoCluster* cl;
  uint16_t* eq;
  uint16_t* rows;
  if (tablesInSmem)
  {
    cl = reinterpret_cast<OoCluster*>(s_raw);
    eq = reinterpret_cast<uint16_t*>(s_raw + (size_t)maxLabels * sizeof(OoCluster));
    rows = eq + ((maxLabels + 7) & ~7);
  }
  else
  {
    cl = clustersAll + (size_t)frame * maxLabels;
    eq = equalAll + (size_t)frame * maxLabels;
    rows = reinterpret_cast<uint16_t*>(s_raw);
  }
  uint16_t* prev = rows;
  uint16_t* cur = rows + bw;
  for (int i = lane; i < 2 * bw; i += 32)
    rows[i] = 0;
  int ncl = 1;                                            // label 0 = background (:181-186); lane 0's copy is the truth
  if (lane == 0)
  {
    eq[0] = 0;
    cl[0] = OoCluster{0, 0, 0};
  }
  __syncwarp();

  // Raster replay, 32 metapixels per step.  Along a run of consecutive "on" cells the label is the running
  // minimum of the non-zero labels seen so far (the left neighbour carries it), so it comes out of a
  // segmented prefix-min scan over the lanes; a run that starts with no labelled neighbour opens a new
  // label (numbered in raster order by a ballot rank).  The masses are order independent and are added
  // per distinct label.  Only the cells where two DIFFERENT labels meet touch the equivalence table, and
  // those are replayed one by one, in order, exactly as the reference does (:92-102).
  const unsigned FULL = 0xFFFFFFFFu;
  // The metapixels do not depend on the walk, but the walk is a chain of short dependent steps: a bitmap word fetched
  // one group ahead arrives after ~500 cycles while an empty group takes ~100.  So whole bitmap rows are staged
  // OO_RING_AHEAD rows ahead into a small ring in shared memory with cp.async, one commit group per row.
  uint16_t* const ring = reinterpret_cast<uint16_t*>(s_raw + ringOfs);
  // Most rows of most frames are empty, and with ~28 walks resident per SM their cost is instruction issue (measured per
  // frame: an empty 320x240 frame took 66 k cycles, 121 instructions per row): so the row overhead is kept short -- up to
  // 256 metapixels per row (W <= 1024) one 16-byte copy per lane fetches a row, one 16-byte load per lane tells whether
  // it has any set bit at all, and a label row that is known to be zero is not zeroed again.
  const int chunks = bw >> 3;                                                // 16-byte pieces of a bitmap row (W % 32 == 0)
  const bool narrow = chunks <= 32;
  auto issue_row = [&](int r)
  {
    if (r < bh)
    {
      if (narrow)
      {
        if (lane < chunks)
          cp_async16(ring + (r & (OO_RING_ROWS - 1)) * bw + lane * 8, bm + (size_t)r * bw + lane * 8);
      }
      else
        for (int c = lane; c < chunks; c += 32)
          cp_async16(ring + (size_t)(r & (OO_RING_ROWS - 1)) * bw + c * 8, bm + (size_t)r * bw + c * 8);
    }
    cp_async_commit();                                                       // also when empty: the group count stays uniform
  };
  bool zeroPrev = true, zeroCur = true;                                      // the label rows start zeroed
  const bool vecRows = narrow && ((uint32_t)__cvta_generic_to_shared(rows) & 15u) == 0u;   // (behind the tables they may not be aligned)
  for (int r = 0; r < OO_RING_AHEAD; ++r)
    issue_row(r);
  for (int row = 0; row < bh; ++row)
  {
    issue_row(row + OO_RING_AHEAD);                                          // overwrites the slot of row - 1, fully read by now
    cp_async_wait<OO_RING_AHEAD>();                                          // all but the newest groups: this row has landed
    __syncwarp();
    const uint16_t* const bmRow = ring + (size_t)(row & (OO_RING_ROWS - 1)) * bw;
    {
      // a row without a single detected cell (most rows of most frames): its labels are all 0, nothing else happens
      bool anyOn = false;
      if (narrow)
      {
        uint4 wv = make_uint4(0u, 0u, 0u, 0u);
        if (lane < chunks)
          wv = *reinterpret_cast<const uint4*>(bmRow + lane * 8);
        anyOn = (wv.x | wv.y | wv.z | wv.w) != 0u;                           // any bit at all: cheap and conservative
      }
      else
      {
#pragma unroll 1
        for (int c = lane; c < bw; c += 32)
          anyOn |= __popc((unsigned)bmRow[c]) > 2;
      }
      // (narrow: a row with set bits but no cell above the threshold goes through the group loop, which skips its groups)
      const bool some = __any_sync(FULL, anyOn);
      if (!some)
      {
        if (!zeroCur)
        {
          if (vecRows)
          {
            if (lane < chunks)
              *reinterpret_cast<uint4*>(cur + lane * 8) = make_uint4(0u, 0u, 0u, 0u);
          }
          else
            for (int c = lane; c < bw; c += 32)
              cur[c] = 0;
          __syncwarp();
        }
        uint16_t* tmp = prev; prev = cur; cur = tmp;
        zeroCur = zeroPrev; zeroPrev = true;
        continue;
      }
    }
    for (int base = 0; base < bw; base += 32)
    {
      const int col = base + lane;
      const bool inside = col < bw;
      const unsigned bmCur = inside ? (unsigned)bmRow[col] : 0u;
      const bool on = inside && __popc(bmCur) > 2;                          // pop(...) > METAPIX_SIZE/2 (:192)
      // up-left, up, up-right labels (:69-83)
      uint32_t p0 = 0, p1 = 0, p2 = 0;
      if (on && row != 0)
      {
        p1 = prev[col];
        if (col != 0) p0 = prev[col - 1];
        if (col != bw - 1) p2 = prev[col + 1];
      }
      const uint32_t carry = (base != 0) ? (uint32_t)cur[base - 1] : 0u;     // label of the cell left of this group
      const unsigned onMask = __ballot_sync(FULL, on);
      if (onMask == 0u)                                                      // nothing detected in these 32 cells
      {
        if (inside) cur[col] = 0;
        __syncwarp();
        continue;
      }
      const bool leftOn = lane == 0 ? (carry != 0u) : ((onMask >> (lane - 1)) & 1u);
      // smallest non-zero up label: 0 - 1 wraps to the largest value, so it is a plain three-way minimum
      const uint32_t mu = min(min(p0 - 1u, p1 - 1u), p2 - 1u) + 1u;
      // a run start without any labelled neighbour opens a new label, numbered in raster order
      const bool opens = on && !leftOn && mu == 0u;
      const unsigned openMask = __ballot_sync(FULL, opens);
      uint32_t x = on ? mu : 0u;
      if (opens)
      {
        const int rank = __popc(openMask & ((1u << lane) - 1u));
        const int lab = ncl + rank;
        x = lab < maxLabels ? (uint32_t)lab : 0u;
        if (lab < maxLabels)
        {
          eq[lab] = (uint16_t)lab;                                           // zero mass, not counted (:104-111)
          cl[lab] = OoCluster{0, 0, 0};
        }
      }
      ncl = min(ncl + __popc(openMask), maxLabels);
      __syncwarp();                                                          // the zeroed records before anybody adds to them
      if (lane == 0 && on && carry != 0u)                                    // the run continues from the previous group
        x = (x == 0u || carry < x) ? carry : x;
      // segmented inclusive scan of "min over non-zero" along the runs.  Usually nothing smaller than the label a run
      // starts with comes in from above further along the run, and then every cell of the run simply takes that label:
      // one shuffle from the run's first lane and a vote instead of the five scan steps.
      bool flag = !on || !leftOn || lane == 0;
      const unsigned startsAll = __ballot_sync(FULL, flag);
      bool plainRuns;
      {
        const unsigned starts = startsAll & ((2u << lane) - 1u);                // segment starts at or below this lane (bit 0 always)
        const uint32_t xs = __shfl_sync(FULL, x, 31 - __clz((int)starts));
        const bool plain = !on || x == 0u || (xs != 0u && x >= xs);
        plainRuns = __all_sync(FULL, plain);
        if (plainRuns)
          x = on ? xs : 0u;
        else
        {
#pragma unroll
          for (int d = 1; d < 32; d <<= 1)
          {
            const uint32_t y = __shfl_up_sync(FULL, x, d);
            const bool g2 = __shfl_up_sync(FULL, (int)flag, d) != 0;
            if (lane >= d && !flag)
            {
              if (y != 0u && (x == 0u || y < x)) x = y;
              flag = g2;
            }
          }
        }
      }
      const uint32_t v = on ? x : 0u;
      if (inside)
        cur[col] = (uint16_t)v;
      // left label of every cell: a[0] of the reference
      uint32_t L = __shfl_up_sync(FULL, v, 1);
      if (lane == 0) L = carry;
      if (!on) L = 0u;
      // masses: every on cell except the openers adds (col, row, 1) to its label.  The cells of one label find each
      // other with one MATCH; the lowest lane of each group adds for all of them (sum of their columns from the bit
      // positions of the peer mask), all labels of the group at once.
      if (plainRuns)
      {
        // every run carries one label: its first lane adds for the whole run (lanes first..e1-1; an opener does not count
        // itself).  Two runs of a group may carry the same label: reductions / atomics.
        if (flag && on)
        {
          const unsigned above = startsAll & ~((2u << lane) - 1u);               // segment starts above this lane
          const int e1 = above ? __ffs((int)above) - 1 : 32;
          const int first = lane + (opens ? 1 : 0);
          const int cnt = e1 - first;
          if (cnt > 0 && v != 0u)
          {
            const int sumc = base * cnt + ((first + e1 - 1) * cnt) / 2;
            if (tablesInSmem)
            {
              atomicAdd(&cl[v].x, sumc); atomicAdd(&cl[v].y, row * cnt); atomicAdd(&cl[v].size, cnt);
            }
            else
            {
              red_add_global(&cl[v].x, sumc); red_add_global(&cl[v].y, row * cnt); red_add_global(&cl[v].size, cnt);
            }
          }
        }
      }
      else
      {
        const bool needAdd = on && !opens && v != 0u;
        const unsigned peers = __match_any_sync(FULL, needAdd ? v : 0xFFFFFFFFu);
        if (needAdd && lane == __ffs((int)peers) - 1)
        {
          const int cnt = __popc(peers);
          const int sumc = base * cnt + __popc(peers & 0xAAAAAAAAu) + 2 * __popc(peers & 0xCCCCCCCCu)
                         + 4 * __popc(peers & 0xF0F0F0F0u) + 8 * __popc(peers & 0xFF00FF00u) + 16 * __popc(peers & 0xFFFF0000u);
          // reductions without a return value: nothing in the walk waits for the table (read back only at the end)
          if (tablesInSmem)
          {
            cl[v].x += sumc; cl[v].y += row * cnt; cl[v].size += cnt;
          }
          else
          {
            red_add_global(&cl[v].x, sumc); red_add_global(&cl[v].y, row * cnt); red_add_global(&cl[v].size, cnt);
          }
        }
      }
      // equivalences: only where a different non-zero label touches the cell, replayed in raster order
      const bool meets = on && !opens && v != 0u &&
                         ((L && L != v) || (p0 && p0 != v) || (p1 && p1 != v) || (p2 && p2 != v));
      unsigned meetMask = __ballot_sync(FULL, meets);
      __syncwarp();
      while (meetMask)
      {
        const int src = __ffs((int)meetMask) - 1;
        meetMask &= meetMask - 1;
        const uint32_t a0 = __shfl_sync(FULL, L, src), a1 = __shfl_sync(FULL, p0, src);
        const uint32_t a2 = __shfl_sync(FULL, p1, src), a3 = __shfl_sync(FULL, p2, src);
        const uint32_t vv = __shfl_sync(FULL, v, src);
        if (lane == 0)
        {
          const uint32_t a[4] = {a0, a1, a2, a3};
#pragma unroll
          for (int n = 0; n < 4; ++n)
            if (a[n])
              if (!(a[n] == vv || eq[a[n]] == eq[vv]))
                eq[a[n]] = eq[vv];
        }
      }
      __syncwarp();
    }
    uint16_t* tmp = prev; prev = cur; cur = tmp;
    zeroCur = zeroPrev; zeroPrev = false;                                    // the row just written has labels in it
  }
  // The tail (merge in label order, std::sort, eight targets) is one lane's sequential work on the label tables.  In
  // global memory every step of it is an L2 round trip; frames with few labels (the usual case) first bring their
  // tables into this warp's corner of shared memory.
  __syncwarp();
  if (!tablesInSmem && ncl <= tailCap)
  {
    OoCluster* sc = reinterpret_cast<OoCluster*>(s_raw + tailOfs);
    uint16_t* se = reinterpret_cast<uint16_t*>(sc + tailCap);
    for (int i = lane; i < ncl; i += 32)
    {
      sc[i] = cl[i];
      se[i] = eq[i];
    }
    __syncwarp();
    cl = sc;
    eq = se;
  }
  ncl = __shfl_sync(FULL, ncl, 0);
  const bool tablesNear = tablesInSmem || ncl <= tailCap;  // the tables the tail works on are in shared memory
  if (lane == 0)
    for (int i = 0; i < ncl; ++i)                         // postProcessing (:115-124)
    {
      const int e = eq[i];
      if (i != e)
      {
        cl[e].x += cl[i].x; cl[e].y += cl[i].y; cl[e].size += cl[i].size;
        cl[i].size = 0;
      }
    }
  __syncwarp();
  // std::sort(..., compareTargetBySize) (:126), of which only the first eight records are ever read (:572-590).  The
  // order std::sort leaves EQUAL sizes in is observable, so it is restated exactly (oo_std_sort) -- but one lane sorting
  // 66 records takes 35 k cycles.  When the sizes decide by themselves, any correct sort gives the same first eight: the
  // warp selects the maximum eight times, and only if two non-empty clusters tie for one of those places (or the tables
  // are large) the sequential sort runs.  Records of size 0 never become targets, so ties among them do not matter.
  int sel[8];
  int nsel = 0;
  bool bySelection = tablesNear && ncl >= 32 && ncl <= 256;   // (a short table is sorted faster than it is searched eight times)
  if (bySelection)
  {
    uint32_t taken = 0u;                                  // bit j: record lane + 32 j
    bool done = false, tie = false;
#pragma unroll
    for (int r = 0; r < 8; ++r)
    {
      sel[r] = 0;
      if (!done)
      {
        int best = -1, bestJ = 0, nBest = 0;
        for (int j = 0; lane + 32 * j < ncl; ++j)
          if (!((taken >> j) & 1u))
          {
            const int sz = cl[lane + 32 * j].size;
            if (sz > best) { best = sz; bestJ = j; nBest = 1; }
            else if (sz == best) ++nBest;
          }
        const int m = __reduce_max_sync(FULL, best);
        if (m <= 0)
          done = true;                                    // only empty clusters left
        else if (__reduce_add_sync(FULL, best == m ? nBest : 0) > 1)
        {
          done = true; tie = true;                        // two clusters of this size: std::sort's own order decides
        }
        else
        {
          const int owner = __ffs((int)__ballot_sync(FULL, best == m)) - 1;
          sel[r] = __shfl_sync(FULL, lane + 32 * bestJ, owner);
          if (lane == owner) taken |= 1u << bestJ;
          nsel = r + 1;
        }
      }
    }
    bySelection = !tie;
  }
  if (lane != 0)
    return;
  if (!bySelection)
    oo_std_sort(cl, ncl);
  const int nTargets = bySelection ? nsel : (ncl < 8 ? ncl : 8);

  ObjOut r;
  for (int i = 0; i < 24; ++i) r.t[i] = 0;
  r.detectHue = r.detectHueTolerance = r.detectSat = r.detectSatTolerance = r.detectVal = r.detectValTolerance = 0;
  bool noObjects = true;
  const int W = g.width, H = g.height;
  DrawInfo di;
  for (int i = 0; i < 20; ++i) di.v[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)                             // cv_ball_detector_seqpass.hpp:572-590
  {
    // slots past the last label: the reference reads beyond its vector (undefined); defined here as empty, and an
    // empty slot (size 0 -> radius 0) never is a target: nothing to compute for it
    if (i >= nTargets)
      break;
    const OoCluster c = cl[bySelection ? sel[i] : i];
    int size = (int)sqrtf((float)(uint16_t)c.size);
    const uint32_t radius = (uint32_t)ceilf((float)size / 3.1415927f);
    size = (int)((uint32_t)(radius * 100u * 4u) / (uint32_t)(bw + bh));
    if (size > 4)
    {
      noObjects = false;
      const int x = (c.x / (c.size + 1)) * 4;
      const int y = (c.y / (c.size + 1)) * 4;
      r.t[3 * i + 2] = (int8_t)(uint8_t)size;
      r.t[3 * i + 0] = (int8_t)(((x - W / 2) * 100 * 2) / W);
      r.t[3 * i + 1] = (int8_t)(((y - H / 2) * 100 * 2) / H);
      di.v[0] |= 1 << i; di.v[1 + 2 * i] = x; di.v[2 + 2 * i] = y;
    }
  }
  if (g.drawInfo)
    static_cast<DrawInfo*>(g.drawInfo)[frame] = di;
  if (noObjects)
  {
    r.t[0] = 0; r.t[1] = 0; r.t[2] = 0;
  }
  out[frame] = r;
  if (labelCounts)
    labelCounts[frame] = ncl;
}

cudaError_t launch_oo(const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                      int paramStride, uint16_t* bitmaps, void* clusters, uint16_t* equal, int maxLabels,
                      void* out, int* labelCounts, cudaStream_t stream,
                      const uint8_t* lutTable, const uint32_t* lutMasks, int smCount,
                      const int* lutFrameList, const LutSets* lutSets)
{
  if (numFrames <= 0)
    return cudaSuccess;
  const int cpr = g.width / 16;
  if (cpr <= 0 || cpr > 1024)
    return cudaErrorInvalidValue;
  const bool viaTable = lutTable != nullptr && cpr <= 768;
  const int k = rows_per_iteration(cpr, 192);
  const int bh = g.height / 4;
  int slabs = (148 * 8 + numFrames - 1) / numFrames;
  const int maxSlabs = bh / (k * 2) > 0 ? bh / (k * 2) : 1;
  if (slabs > maxSlabs) slabs = maxSlabs;
  if (slabs < 1) slabs = 1;
  if (slabs > 65535) slabs = 65535;
  dim3 grid((unsigned)numFrames, (unsigned)slabs);
  cudaError_t e;
  if (lutSets && lutFrameList && cpr <= 768)
    e = launch_oo_bitmap_lut_sets(g, frames, bitmaps, smCount, stream, lutFrameList, *lutSets);
  else if (viaTable)
    e = launch_oo_bitmap_lut(g, numFrames, frames, lutTable, lutMasks, bitmaps, smCount, stream);
  else
  {
    oo_bitmap_kernel<<<grid, cpr * k, 0, stream>>>(g, frames, params, paramStride, bitmaps, cpr, k);
    ++g_launches_grid;
    e = cudaGetLastError();
  }
  if (e != cudaSuccess)
    return e;
  const size_t rowBytes = (size_t)2 * (g.width / 4) * sizeof(uint16_t);
  const size_t tableBytes = (size_t)maxLabels * sizeof(OoCluster) + (size_t)((maxLabels + 7) & ~7) * sizeof(uint16_t);
  // one warp per frame: residency is what hides its latency, so the tables only live in shared memory while
  // that still leaves ~10 CTAs per SM (env override for A/B measurements)
  static const long smemLimit = getenv("TRIKB200_OO_SMEM_LIMIT") ? atol(getenv("TRIKB200_OO_SMEM_LIMIT")) : 4 * 1024;
  const int tablesInSmem = ((long)(tableBytes + rowBytes) <= smemLimit) ? 1 : 0;
  // tables in global memory: room for the tail's copy of up to tailCap labels (12 + 2 bytes each)
  static const int tailCapEnv = getenv("TRIKB200_OO_TAIL_CAP") ? atoi(getenv("TRIKB200_OO_TAIL_CAP")) : 256;
  const int tailCap = tablesInSmem ? 0 : (tailCapEnv < maxLabels ? tailCapEnv : maxLabels);
  const size_t tailOfs = (rowBytes + 15) & ~(size_t)15;
  const size_t ringBytes = (size_t)OO_RING_ROWS * (g.width / 4) * sizeof(uint16_t);
  const size_t ringOfs = tablesInSmem ? ((rowBytes + tableBytes + 15) & ~(size_t)15)
                                      : ((tailOfs + (size_t)tailCap * (sizeof(OoCluster) + sizeof(uint16_t)) + 15) & ~(size_t)15);
  const size_t smemPerWarp = (ringOfs + ringBytes + 15) & ~(size_t)15;
  const size_t smem = smemPerWarp * OO_WARPS_PER_CTA;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(oo_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const unsigned ctas = (unsigned)((numFrames + OO_WARPS_PER_CTA - 1) / OO_WARPS_PER_CTA);
  oo_cluster_kernel<<<ctas, 32 * OO_WARPS_PER_CTA, smem, stream>>>(g, bitmaps, reinterpret_cast<OoCluster*>(clusters), equal,
                                                                   maxLabels, tablesInSmem, reinterpret_cast<ObjOut*>(out), labelCounts,
                                                                   numFrames, (int)smemPerWarp, tailCap, (int)tailOfs, (int)ringOfs);
  ++g_launches_grid;
  return cudaGetLastError();
}

} // namespace trikb200
