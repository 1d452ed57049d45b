// trik_kernels_ommaj.cu -- the ov7670 mxn grid-colour sensor (OM): majority pass (sm_100a).
//
// ov7670/mxn_sensor/include/internal/cv_ball_detector_seqpass.hpp: GetImgColor2 :411-452 builds, per cell, a 32x4x4
// histogram of (H>>3, S>>6, V>>6) in raster order and reports "the first bin to reach the final maximum" (:434-440).
//
// om_table_kernel (trik_kernels_omtab.cu) builds those histograms exactly: one table gather and one histogram update per
// pixel, bound by the L1 / shared-memory data pipe (an outlier pixel costs a gather wavefront of its own and a shared
// atomic).  But the answer is one bin per cell, and a cell painted one colour does not need its histogram:
//
//     if a bin holds MORE THAN HALF of a cell's pixels it is the unique maximum, hence "the first to reach the final
//     maximum", whatever the other pixels are and in whatever order they come.
//
// So this pass only PROVES a majority, with a lower bound that costs four instructions per pixel and no gather:
//   1. per cell, 16 sample pixels vote for a candidate chroma pair;
//   2. the table row of that chroma pair ([U][V][0..255], 512 bytes, one coalesced read -- the only table access of the
//      pass) gives the candidate bin (the one of a voter's luma) and the luma interval lo..hi around it on which the
//      row holds that bin;
//   3. the cell's pixels are streamed once (16 pixels per thread and row, cp.async ring as the line kernels) and those
//      with exactly the candidate's chroma and a luma inside lo..hi are counted -- each of them IS in the candidate's
//      bin, so the count is a lower bound of that bin's count;
//   4. 2 * count > cell pixels: the candidate is the cell's colour, exactly as the reference finds it.  Otherwise
//      nothing is concluded: the cell row is put on a list and om_table_list_kernel builds its histograms as before.
// Bit-identical by construction: the shortcut is only taken where it is a proof.  Frames it cannot serve (noise, chroma
// noise, cells of several colours, cells narrower than 16 pixels) cost the samples and then take the histogram path.
#include <atomic>
#include <cstdlib>
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"

namespace trikb200 {

extern std::atomic<long long> g_launches_omtab;

constexpr int OMJ_MAX_CELLS = 100;           // widthM * heightN <= 100 (outColor[100])
constexpr int OMJ_MIN_VOTES = 6;             // of 16 samples: below that a cell is not worth streaming for

struct OmjCand {
  uint32_t half;                             // chroma half-word V | U << 8 as it lies in the plane
  uint32_t lo;                               // luma interval lo..hi; lo = 256 (nothing passes) for a cell without candidate
  uint32_t nhi;                              // 255 - hi
  uint32_t bin4;                             // table entry of the candidate: bin * 4
};

// a * b + c on the FMA pipe (IMAD), also where b is 1: the ALU pipe is the busy one in this kernel
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

__device__ __forceinline__ uint32_t madhi_u32(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t d;
  asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

__device__ __forceinline__ void omj_push(int* __restrict__ list, int* __restrict__ count, int item)
{
  list[atomicAdd(count, 1)] = item;
}

template <int OMJ_STAGES, int MINB>
__global__ void __launch_bounds__(192, MINB)
om_major_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
                const int paramStride, const uint16_t* __restrict__ table, const uint32_t* __restrict__ colorTable,
                int32_t* __restrict__ out, const int maxGridRows, const int cellRowsPerCta, const int ctasPerFrame,
                const int cpr, const int rpi, int* __restrict__ fbList, int* __restrict__ fbCount, int* __restrict__ fbCountNext,
                const uint32_t one, const uint32_t minusOne, const uint32_t k17)     // 1, -1, 1 << 17: multipliers the compiler cannot fold
{
  __shared__ OmjCand s_cand[OMJ_MAX_CELLS];
  __shared__ uint32_t s_count[OMJ_MAX_CELLS];
  __shared__ uint32_t s_rowBad[OMJ_MAX_CELLS];
  extern __shared__ uint4 s_ring[];            // [STAGES][2][blockDim]: luma chunk, chroma chunk

  if (blockIdx.x == 0 && threadIdx.x == 0)
    *fbCountNext = 0;                                     // the counter of the previous batch (drained, by stream order)
  const int frame = blockIdx.x / ctasPerFrame;
  const int part = blockIdx.x - frame * ctasPerFrame;
  const FrameParams p = params[(size_t)frame * paramStride];
  const int M = (int)p.gridRows, N = (int)p.gridCols;
  const int cr0 = part * cellRowsPerCta;
  if (cr0 >= M || N <= 0)
    return;                                               // uniform per CTA
  const int cr1 = min(cr0 + cellRowsPerCta, M);
  const int W = g.width, H = g.height;
  const int ws = W / N, hs = H / M;                       // m_widthStep, m_heightStep (:587-588); remainders ignored
  const int t = threadIdx.x;
  if (ws < 16 || hs < 1)
  {
    // cells narrower than a thread's chunk (or no rows at all): the histogram path takes them
    if (t < cr1 - cr0)
      omj_push(fbList, fbCount, frame * maxGridRows + cr0 + t);
    return;
  }
  const int nc = (cr1 - cr0) * N;                         // cells of this CTA, <= 100
  const uint8_t* fbase = frames + (size_t)frame * g.frameStride;
  const size_t chromaOfs = (size_t)H * g.lineLength;

  // ---- the stream: rows startRow + rr + i * rpi of this thread's 16-pixel column, i = 0 .. iters-1 -----------------
  const int cc = t % cpr, rr = t / cpr;
  const int startRow = cr0 * hs, endRow = cr1 * hs;
  const int iters = (startRow + rr < endRow) ? (endRow - startRow - rr + rpi - 1) / rpi : 0;
  const uint8_t* fillPtr = fbase + (size_t)cc * 16u + (size_t)(startRow + rr) * g.lineLength;
  const size_t rowStep = (size_t)rpi * g.lineLength;
  const uint32_t planeBytes = blockDim.x * 16u, stageBytes = 2u * planeBytes;
  const uint32_t slotBase = (uint32_t)__cvta_generic_to_shared(s_ring) + (uint32_t)t * 16u;
  auto fill = [&](uint32_t slot)
  {
    const uint32_t dst = slotBase + slot * stageBytes;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(fillPtr) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + planeBytes), "l"(fillPtr + chromaOfs) : "memory");
  };
  // the first rows are requested before the candidates are known: they travel while the samples vote
  int fillIt = 0;
#pragma unroll
  for (int sIdx = 0; sIdx < OMJ_STAGES - 1; ++sIdx)
  {
    if (fillIt < iters)
      fill((uint32_t)sIdx);
    cp_async_commit();
    ++fillIt;
    fillPtr += rowStep;
  }

  // ---- 1 + 2: candidates.  Half a warp per cell: 16 samples on a 4 x 4 lattice inside the cell ---------------------
  for (int i = t; i < cr1 - cr0; i += blockDim.x)
    s_rowBad[i] = 0u;
  for (int i = t; i < nc; i += blockDim.x)
    s_count[i] = 0u;
  __syncthreads();
  {
    const int lane = t & 31, l16 = lane & 15;
    const unsigned hmask = 0xFFFFu << (lane & 16);
    const int nhw = (int)(blockDim.x >> 4);                 // (a trailing partial half-warp takes no cell)
    for (int cell = (t | 15) < (int)blockDim.x ? t >> 4 : nc; cell < nc; cell += nhw)      // whole half-warps only
    {
      const int cr = cr0 + cell / N, c = cell - (cell / N) * N;
      const int row = cr * hs + (hs * (2 * (l16 >> 2) + 1)) / 8;
      const int col = c * ws + (ws * (2 * (l16 & 3) + 1)) / 8;
      const uint8_t* ptr = fbase + (size_t)row * g.lineLength;
      const uint32_t y = ptr[col];
      const uint32_t half = *reinterpret_cast<const uint16_t*>(ptr + chromaOfs + (col & ~1));
      // the samples vote for a chroma pair; the luma of one of its voters and the table row of that pair then give the
      // candidate bin and its luma interval in ONE further round trip
      // (among the voters of the winning chroma pair, one whose luma is the most common: not an outlier's)
      const int votes = __popc(__match_any_sync(hmask, half) & hmask);
      const int score = votes * 32 + __popc(__match_any_sync(hmask, (half << 8) | y) & hmask);
      const int bestScore = __reduce_max_sync(hmask, score);
      const int best = bestScore >> 5;
      const int owner = __ffs((int)(__ballot_sync(hmask, score == bestScore) & hmask)) - 1;
      const uint32_t cHalf = __shfl_sync(hmask, half, owner);
      const int cY = (int)__shfl_sync(hmask, y, owner);
      OmjCand cand;
      cand.half = cHalf; cand.bin4 = 0u;
      if (best < OMJ_MIN_VOTES)
      {
        cand.lo = 256u; cand.nhi = 0u;                                     // lo = 256: no luma passes
        if (l16 == 0)
          s_rowBad[cell / N] = 1u;
      }
      else
      {
        // the luma interval around cY on which the candidate chroma's table row holds the candidate bin
        const uint4* rowp = reinterpret_cast<const uint4*>(table + ((size_t)cHalf << 8)) + 2 * l16;
        const uint4 v0 = __ldg(rowp), v1 = __ldg(rowp + 1);
        const uint32_t cE = __ldg(table + (((size_t)cHalf << 8) | (uint32_t)cY));
        cand.bin4 = cE;
        const uint32_t wv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t eq = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
          if ((wv[k] & 0xFFFFu) == cE) eq |= 1u << (2 * k);
          if ((wv[k] >> 16) == cE) eq |= 2u << (2 * k);
        }
        const uint32_t zeros = ~eq & 0xFFFFu;
        const int basePos = 16 * l16, rel = cY - basePos;                  // own positions basePos .. basePos + 15
        const uint32_t below = rel <= 0 ? 0u : (rel >= 16 ? 0xFFFFu : ((1u << rel) - 1u));
        const uint32_t above = rel >= 15 ? 0u : (rel < 0 ? 0xFFFFu : ((0xFFFFu << (rel + 1)) & 0xFFFFu));
        const uint32_t zb = zeros & below, za = zeros & above;
        const int zlo = zb ? basePos + 31 - __clz((int)zb) : -1;
        const int zhi = za ? basePos + __ffs((int)za) - 1 : 256;
        const int lo = __reduce_max_sync(hmask, zlo) + 1;
        const int hi = __reduce_min_sync(hmask, zhi) - 1;
        cand.lo = (uint32_t)lo;
        cand.nhi = (uint32_t)(255 - hi);
      }
      if (l16 == 0)
        s_cand[cell] = cand;
    }
  }
  __syncthreads();
  {
    // nothing to prove anywhere (noise, chroma noise): straight to the histogram path
    bool allBad = true;
    for (int i = 0; i < cr1 - cr0; ++i)
      allBad = allBad && s_rowBad[i] != 0u;
    if (allBad)
    {
      cp_async_wait<0>();
      if (t < cr1 - cr0)
        omj_push(fbList, fbCount, frame * maxGridRows + cr0 + t);
      return;
    }
  }

  // ---- 3: the stream.  This thread's 16 pixels lie in at most two cells (ws >= 16): A left of `split`, B from it on --
  const int col0 = cc * 16;
  const int cA = col0 / ws;                               // may be >= N: remainder columns, counted nowhere
  const int split = (cA + 1) * ws - col0;                 // >= 1; >= 16: all pixels in A
  int git = 0;                                            // global iteration (ring slot = git % STAGES)
  for (int cr = cr0; cr < cr1; ++cr)
  {
    auto first_it = [&](int r) { const int x = r - startRow - rr; return x <= 0 ? 0 : (x + rpi - 1) / rpi; };
    const int itEnd = min(first_it((cr + 1) * hs), iters);
    const int nIt = itEnd - git;
    if (nIt <= 0)
      continue;
    OmjCand a, b;
    a.half = 0u; a.lo = 256u; a.nhi = 0u; a.bin4 = 0u;
    b = a;
    if (cA < N) a = s_cand[(cr - cr0) * N + cA];
    if (cA + 1 < N) b = s_cand[(cr - cr0) * N + cA + 1];
    // per luma word w (pixels 4w .. 4w+3, chroma pairs 2w and 2w+1): the pair's candidate chroma is that of its EVEN
    // pixel's cell; an odd pixel whose cell differs from its pair's (odd cell width) is not counted -- a lower bound stays one.
    // With yG = 0x8000 + Y per lane:  yG - lo2  and  hiG - Y2 = (0x80008000 + hi2) - (yG - 0x80008000) = C - yG  with
    // C = 0x010000FF - nhi2 (mod 2^32; hi2 = 0x00FF00FF - nhi2, and twice the guard is 0x1_0001_0000) are one multiply-add
    // each, and exact: every RESULT lane lies in 0x7F00 .. 0x80FF, so no lane borrows from its neighbour whatever the
    // intermediate sums look like.
    uint32_t candW[4], mLoE[4], mHiE[4], mLoO[4], mHiO[4];
#pragma unroll
    for (int w = 0; w < 4; ++w)
    {
      const bool e0A = 4 * w < split, e1A = 4 * w + 2 < split;              // even pixels 4w, 4w+2
      const bool o0A = 4 * w + 1 < split, o1A = 4 * w + 3 < split;          // odd pixels 4w+1, 4w+3
      candW[w] = (e0A ? a.half : b.half) | ((e1A ? a.half : b.half) << 16);
      mLoE[w] = 0u - ((e0A ? a.lo : b.lo) | ((e1A ? a.lo : b.lo) << 16));
      mHiE[w] = 0x010000FFu - ((e0A ? a.nhi : b.nhi) | ((e1A ? a.nhi : b.nhi) << 16));
      const uint32_t lo0 = o0A == e0A ? (o0A ? a.lo : b.lo) : 256u, lo1 = o1A == e1A ? (o1A ? a.lo : b.lo) : 256u;
      const uint32_t h0 = o0A == e0A ? (o0A ? a.nhi : b.nhi) : 0u, h1 = o1A == e1A ? (o1A ? a.nhi : b.nhi) : 0u;
      mLoO[w] = 0u - (lo0 | (lo1 << 16));
      mHiO[w] = 0x010000FFu - (h0 | (h1 << 16));
    }
    uint32_t accE[4] = {0u, 0u, 0u, 0u}, accO[4] = {0u, 0u, 0u, 0u};        // passes per pixel position, two 16-bit lanes each
#pragma unroll 2
    for (int it = 0; it < nIt; ++it, ++git)
    {
      const uint32_t slot = (uint32_t)git % OMJ_STAGES;
      if (fillIt < iters)
        fill((slot + OMJ_STAGES - 1u) % OMJ_STAGES);
      cp_async_commit();
      ++fillIt;
      fillPtr += rowStep;
      cp_async_wait<OMJ_STAGES - 1>();
      const uint32_t src = slotBase + slot * stageBytes;
      uint32_t L[4], Cw[4];
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(L[0]), "=r"(L[1]), "=r"(L[2]), "=r"(L[3]) : "r"(src));
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(Cw[0]), "=r"(Cw[1]), "=r"(Cw[2]), "=r"(Cw[3]) : "r"(src + planeBytes));
#pragma unroll
      for (int w = 0; w < 4; ++w)
      {
        // The work is split between the two integer pipes (each issues one warp instruction per two cycles): byte
        // permutes and logic on the ALU pipe, the compares and the counting as multiply-adds on the FMA pipe.
        // eqG: guard bit 15 in the lane of a pair whose chroma IS the candidate's
        const uint32_t eqG = mad_u32(__vminu2(Cw[w] ^ candW[w], 0x00010001u), 0xFFFF8000u, 0x80008000u);
        const uint32_t yE = __byte_perm(L[w], 0x80808080u, 0x4240);         // 0x8000 + Y of pixels 4w, 4w+2
        const uint32_t yO = __byte_perm(L[w], 0x80808080u, 0x4341);         // ... 4w+1, 4w+3
        // bit 15 of a lane stays set <=> Y >= lo (guard + Y - lo), resp. Y <= hi (guard + hi - Y, as C - (guard + Y))
        const uint32_t xE = mad_u32(yE, one, mLoE[w]) & mad_u32(yE, minusOne, mHiE[w]) & eqG;
        const uint32_t xO = mad_u32(yO, one, mLoO[w]) & mad_u32(yO, minusOne, mHiO[w]) & eqG;
        accE[w] = madhi_u32(xE, k17, accE[w]);                              // bit 15 -> +1 in lane 0, bit 31 -> +1 in lane 1
        accO[w] = madhi_u32(xO, k17, accO[w]);
      }
    }
    // pixels left of `split` count for cell A, the others for B
    uint32_t cntA = 0u, cntB = 0u;
#pragma unroll
    for (int w = 0; w < 4; ++w)
    {
#pragma unroll
      for (int h = 0; h < 2; ++h)                                           // lane h: even pixel 4w + 2h, odd pixel 4w + 2h + 1
      {
        const uint32_t pE = (accE[w] >> (16 * h)) & 0xFFFFu, pO = (accO[w] >> (16 * h)) & 0xFFFFu;
        if (4 * w + 2 * h < split) cntA += pE; else cntB += pE;
        if (4 * w + 2 * h + 1 < split) cntA += pO; else cntB += pO;
      }
    }
    if (cntA && cA < N) atomicAdd(&s_count[(cr - cr0) * N + cA], cntA);
    if (cntB && cA + 1 < N) atomicAdd(&s_count[(cr - cr0) * N + cA + 1], cntB);
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- 4: the verdicts ----------------------------------------------------------------------------------------------
  int32_t* frameOut = out + (size_t)frame * 100;
  for (int cell = t; cell < nc; cell += blockDim.x)
  {
    const int crl = cell / N, c = cell - crl * N;
    if (2u * s_count[cell] > (uint32_t)(ws * hs) && s_rowBad[crl] == 0u)
      frameOut[(cr0 + crl) * N + c] = (int32_t)colorTable[s_cand[cell].bin4 >> 2];
    else
      s_rowBad[crl] = 1u;                                                   // (benign race: every writer stores 1)
  }
  __syncthreads();
  if (t < cr1 - cr0 && s_rowBad[t] != 0u)
    omj_push(fbList, fbCount, frame * maxGridRows + cr0 + t);               // om_table_list_kernel rewrites the whole cell row
}

cudaError_t launch_om_major(const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                            int paramStride, const uint16_t* table, const uint32_t* colorTable, int32_t* out,
                            int maxGridRows, int* fbList, int* fbCount, int* fbCountNext, cudaStream_t stream)
{
  if (numFrames <= 0 || maxGridRows <= 0)
    return cudaSuccess;
  const int cpr = g.width / 16;
  if (cpr <= 0 || cpr > 192)
    return cudaErrorInvalidValue;
  // ~160 threads, a whole number of rows per iteration, whole warps when that is possible
  static const int targetThreads = getenv("TRIKB200_OMJ_THREADS") ? atoi(getenv("TRIKB200_OMJ_THREADS")) : 160;
  int rpi = (targetThreads + cpr - 1) / cpr;
  for (int j = 0; j < 16; ++j)
    if ((cpr * (rpi + j)) % 32 == 0 && cpr * (rpi + j) <= 192)
    {
      rpi += j;
      break;
    }
  while (cpr * rpi > 192) --rpi;
  if (rpi < 1) rpi = 1;
  const int threads = cpr * rpi;
  // one CTA per frame once the frames alone fill the machine, else one per cell row
  // (whole frames amortise the candidate search better, but the grid should still be five or more waves of CTAs)
  static const int perFrameFrom = getenv("TRIKB200_OMJ_PER_FRAME_FROM") ? atoi(getenv("TRIKB200_OMJ_PER_FRAME_FROM")) : 148 * 4 * 5;
  const int cellRowsPerCta = numFrames >= perFrameFrom ? maxGridRows : 1;
  const int ctasPerFrame = (maxGridRows + cellRowsPerCta - 1) / cellRowsPerCta;
  const long long grid = (long long)numFrames * ctasPerFrame;
  if (grid > 0x7FFFFFFFLL)
    return cudaErrorInvalidValue;
  static const int stages = getenv("TRIKB200_OMJ_STAGES") ? atoi(getenv("TRIKB200_OMJ_STAGES")) : 4;
  static const int minb = getenv("TRIKB200_OMJ_MINB") ? atoi(getenv("TRIKB200_OMJ_MINB")) : 4;
  const size_t smem = (size_t)stages * 2 * 16 * threads;
#define OMJ_LAUNCH(S, B)                                                                                                     \
  do {                                                                                                                       \
    if (smem > 48u * 1024u)                                                                                                  \
      cudaFuncSetAttribute(om_major_kernel<S, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                   \
    om_major_kernel<S, B><<<(unsigned)grid, threads, smem, stream>>>(g, frames, params, paramStride, table, colorTable, out, \
                                                                     maxGridRows, cellRowsPerCta, ctasPerFrame, cpr, rpi,    \
                                                                     fbList, fbCount, fbCountNext, 1u, 0xFFFFFFFFu, 1u << 17);            \
  } while (0)
  if (stages == 8) { if (minb >= 5) OMJ_LAUNCH(8, 5); else if (minb == 4) OMJ_LAUNCH(8, 4); else OMJ_LAUNCH(8, 3); }
  else if (stages == 6) { if (minb >= 5) OMJ_LAUNCH(6, 5); else if (minb == 4) OMJ_LAUNCH(6, 4); else OMJ_LAUNCH(6, 3); }
  else { if (minb >= 5) OMJ_LAUNCH(4, 5); else if (minb == 4) OMJ_LAUNCH(4, 4); else OMJ_LAUNCH(4, 3); }
#undef OMJ_LAUNCH
  ++g_launches_omtab;
  return cudaGetLastError();
}

} // namespace trikb200
