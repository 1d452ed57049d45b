// trik_kernels_lut.cu -- HSV detection through a chroma-indexed table (sm_100a).
//
// The HSV threshold of the object sensors (webcam/object_sensor/include/internal/cv_ball_detector_seqpass.hpp:
// 171-284: YUV -> RGB888 -> 0x00VVSSHH -> range test) costs ~58 integer instructions per pixel, which
// binds sum_kernel<WO> to the ALU pipe at 0.13 of the HBM roofline.  But for FIXED thresholds the test is a
// pure function of the pixel's three bytes, and for a fixed chroma pair (U, V) the set of luma values that
// pass is almost always one interval: V = max(R,G,B) grows with Y, S falls, H barely moves.  So:
//
//   chroma_table_kernel   evaluates the exact arithmetic (detect_pair_bits, the same code the direct kernels
//                         run) on all 2^24 (Y,U,V) once per threshold set: 65 536 warps, one per chroma pair,
//                         each producing the 256-bit pass mask over Y.  It writes
//                           lo[U | V<<8], nhi[U | V<<8] = lo, 255 - hi   when the mask is exactly the interval lo..hi,
//                                                         NEVER          when no luma passes,
//                                                         RAGGED         otherwise (rounding jitter near a hue /
//                                                                        saturation bound: 1..10 % of the entries),
//                         and the 256-bit masks themselves (2 MB, L2 resident) for the RAGGED entries.
//   wo_lut_kernel         the WO pass with the two 64 KB byte tables in shared memory: per pixel pair two byte LDS
//                         and two packed compares instead of the HSV arithmetic; a warp that meets a RAGGED entry
//                         fetches the mask word of those pixels from L2.  The work is split between the ALU pipe
//                         (byte extraction, guard lanes, final AND) and the FMA pipe (the compares and all
//                         accumulation are IMAD / IMAD.HI), ~17 instructions per pair like the line kernels.  The CTA is persistent (one per SM, the table is loaded
//                         once) and is split into independent groups of threads, one frame per group at a
//                         time, synchronised by named barriers.
//
// Results are bit-identical to the arithmetic path by construction (the table IS that path, tabulated), and
// tests/test_lut_gpu.py checks it on all 2^24 inputs per threshold set and on frames.
#include <atomic>
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"
#include "trik_line.cuh"
#include "trik_lut.cuh"

namespace trikb200 {

std::atomic<long long> g_launches_lut{0};
static int g_lutSkew = 1;
void set_lut_skew(int on) { g_lutSkew = on; }
static int g_lutParts = 0;                    // bands per frame of the WO table kernel: 0 = chosen per launch, 1 / 2 / 4 / 8 = fixed
void set_lut_parts(int parts) { g_lutParts = parts; }

// ---------------------------------------------------------------------------------------------
// table construction: one warp per chroma pair
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
chroma_table_kernel(const uint32_t from, const uint32_t to, const uint32_t expected,
                    uint8_t* __restrict__ table, uint32_t* __restrict__ masks)
{
  __shared__ HueLutEntry s_lutHue[256];
  __shared__ uint16_t s_lut43[256];
  __shared__ uint16_t s_lut255[256];
  fill_div_luts(s_lut43, s_lut255);
  fill_hue_lut(s_lutHue);
  __syncthreads();
  const HsvBounds bd = make_bounds(from, to);

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // U | V << 8
  const uint32_t u = idx & 0xFFu, v = idx >> 8;
  uint32_t w[8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
  {
    // this lane's pixel pair: Y0 = 2*lane + 64*k, Y1 = Y0 + 1, as the YUYV word  Y0 U Y1 V
    const uint32_t y0 = 2u * lane + 64u * (uint32_t)k;
    const uint32_t word = y0 | (u << 8) | ((y0 + 1u) << 16) | (v << 24);
    const uint32_t det = detect_pair_bits(word & 0x00FF00FFu, word, coef_yuyv(), s_lutHue, s_lut255, bd, expected);
    const uint32_t bits = det << (2u * (lane & 15u));
    w[2 * k]     = __reduce_or_sync(0xFFFFFFFFu, lane < 16u ? bits : 0u);          // Y = 64k      .. 64k + 31
    w[2 * k + 1] = __reduce_or_sync(0xFFFFFFFFu, lane < 16u ? 0u : bits);          // Y = 64k + 32 .. 64k + 63
  }
  uint32_t count = 0u, rises = 0u, lo = 256u, hi = 0u, carry = 0u;
#pragma unroll
  for (int m = 0; m < 8; ++m)
  {
    count += (uint32_t)__popc(w[m]);
    rises += (uint32_t)__popc(w[m] & ~((w[m] << 1) | carry));
    carry = w[m] >> 31;
    if (w[m] != 0u)
    {
      if (lo == 256u) lo = 32u * (uint32_t)m + (uint32_t)__ffs((int)w[m]) - 1u;
      hi = 32u * (uint32_t)m + 31u - (uint32_t)__clz((int)w[m]);
    }
  }
  if (lane == 0u)
  {
    const uint8_t eLo  = (uint8_t)(count == 0u ? LUT_NEVER_LO : (rises == 1u ? lo : LUT_RAGGED_LO));
    const uint8_t eNhi = (uint8_t)(count == 0u ? LUT_NEVER_NHI : (rises == 1u ? 255u - hi : LUT_RAGGED_NHI));
    table[idx]          = eLo;
    table[65536u + idx] = eNhi;
    // the same entries as the skewed shared-memory image of wo_lut_kernel<., true> (rows 260 bytes apart), copied verbatim
    const uint32_t phys = idx + (idx >> 6);
    table[LUT_SKEW_IMAGE_OFS + phys]                   = eLo;
    table[LUT_SKEW_IMAGE_OFS + LUT_STRIDE_SKEW + phys] = eNhi;
  }
  if (lane < 8u)
  {
    uint32_t mine = w[0];
#pragma unroll
    for (int m = 1; m < 8; ++m)
      if (lane == (uint32_t)m) mine = w[m];
    masks[(size_t)idx * 8u + lane] = mine;
  }
}

cudaError_t launch_chroma_table(uint32_t from, uint32_t to, uint32_t expected, uint8_t* table, uint32_t* masks,
                                cudaStream_t stream)
{
  chroma_table_kernel<<<65536 / 8, 256, 0, stream>>>(from, to, expected, table, masks);
  ++g_launches_lut;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// WO through the table
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void group_barrier(int id, int count)
{
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}

// Persistent CTA: `groups` independent groups of `gthreads` threads, one frame per group at a time.
// Inside a group the decomposition is sum_kernel's: a thread owns a 16-byte column chunk and walks down the rows.
// SKEW: the table rows (one per V, 256 bytes apart) all start in the same shared-memory bank, so lanes whose chroma differs
// only in V collide (a camera with chroma noise: ~5 values of V per warp -> 5-way conflicts on both lookups).  With SKEW
// the entry of (U, V) lives at ci + (ci >> 6): rows 260 bytes apart (each starts one bank further), still injective
// (U + (U >> 6) <= 258), one IMAD.HI more per pixel pair.
template <bool SKEW>
__device__ __forceinline__ uint32_t lut_phys(uint32_t ci) { return SKEW ? ci + __umulhi(ci, 1u << 26) : ci; }

template <int STAGES, bool SKEW>
__device__ __forceinline__ void
wo_lut_body(const Geometry& g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
            const uint8_t* __restrict__ table, const uint32_t* __restrict__ masks, TargetOut* __restrict__ out,
            const int numFrames, const int groups, const int gthreads, const int cpr, const int rpi,
            const int parts, const int rowsPerPart, SumAcc* __restrict__ acc, const int* __restrict__ frameList,
            const int ctaIndex, const int ctaCount)
{
  constexpr uint32_t STRIDE = SKEW ? LUT_STRIDE_SKEW : LUT_STRIDE_PLAIN;
  extern __shared__ __align__(16) uint8_t s_raw[];
  uint8_t* const s_table = s_raw;                                                      // lo[STRIDE], nhi[STRIDE]
  uint4* const s_ring = reinterpret_cast<uint4*>(s_raw + 2u * STRIDE);                // [STAGES][groups * gthreads]
  uint32_t* const s_red = reinterpret_cast<uint32_t*>(s_raw + 2u * STRIDE + (size_t)STAGES * groups * gthreads * 16);   // [groups][32][4]

  // the table, once per CTA
  {
    const uint4* src = reinterpret_cast<const uint4*>(table + (SKEW ? LUT_SKEW_IMAGE_OFS : 0u));
    uint4* dst = reinterpret_cast<uint4*>(s_table);
    for (int i = threadIdx.x; i < (int)(2u * STRIDE / 16u); i += blockDim.x)
      dst[i] = __ldg(src + i);
  }
  __syncthreads();

  const int t = threadIdx.x;
  const int group = t / gthreads;
  if (group >= groups)
    return;                                              // spare threads of the CTA (they took part in the table load)
  const int tg = t - group * gthreads;
  const int cc = tg % cpr;
  const int rr = tg / cpr;
  const int warpInGroup = tg >> 5, lane = tg & 31, gwarps = gthreads >> 5;
  const size_t rowStep = (size_t)rpi * g.lineLength;
  const uint32_t tbl = (uint32_t)__cvta_generic_to_shared(s_table);
  const uint32_t stageBytes = (uint32_t)(groups * gthreads) * 16u;
  const uint32_t slotBase = (uint32_t)__cvta_generic_to_shared(s_ring) + (uint32_t)t * 16u;
  uint32_t* const myRed = s_red + group * 128;
  const FrameParams p = params[0];
  const uint32_t W = (uint32_t)g.width, H = (uint32_t)g.height;

  // A work item is one frame, or -- when whole frames would leave the last round of the persistent groups mostly empty
  // (1024 frames on 888 groups: two rounds for 1.15 rounds of work) -- one of `parts` bands of rowsPerPart rows; the bands of
  // a frame meet in its SumAcc record (zero between launches, as the sum kernels keep it) and the last one finalises.
  for (int item = ctaIndex * groups + group; item < numFrames * parts; item += ctaCount * groups)
  {
    const int slot = parts == 1 ? item : item / parts;
    const int row0 = parts == 1 ? 0 : (item - slot * parts) * rowsPerPart;
    const int frame = frameList ? frameList[slot] : slot;  // a batch under several threshold sets: this set's frames
    const int rowsHere = min(rowsPerPart, g.height - row0);
    const int itersAll = rr < rowsHere ? (rowsHere - rr + rpi - 1) / rpi : 0;
    const uint8_t* fillPtr = frames + (size_t)frame * g.frameStride + (size_t)cc * 16u + (size_t)(row0 + rr) * g.lineLength;
    uint32_t passes = 0u, inIdx = 0u, syPass = 0u;       // 32-bit totals of this thread
    int fillIt = 0;
#pragma unroll
    for (int sIdx = 0; sIdx < STAGES - 1; ++sIdx)
    {
      if (fillIt < itersAll)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(slotBase + (uint32_t)sIdx * stageBytes), "l"(fillPtr) : "memory");
      cp_async_commit();
      ++fillIt;
      fillPtr += rowStep;
    }
    uint32_t S = 0u, A = 0u, SI = 0u;                    // packed 16-bit lane sums of the current segment
    // one row chunk: ring slot `slot` holds it, `j` is its iteration number inside the segment
    auto body = [&](const uint32_t slot, const uint32_t j)
    {
      if (fillIt < itersAll)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(slotBase + ((slot + STAGES - 1u) % STAGES) * stageBytes), "l"(fillPtr) : "memory");
      cp_async_commit();
      ++fillIt;
      fillPtr += rowStep;
      cp_async_wait<STAGES - 1>();
      uint32_t wd[4];
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(wd[0]), "=r"(wd[1]), "=r"(wd[2]), "=r"(wd[3]) : "r"(slotBase + slot * stageBytes));
      uint32_t x[4];
      bool ragged = false;
#pragma unroll
      for (int k = 0; k < 4; ++k)
      {
        const uint32_t ci = lut_phys<SKEW>(__byte_perm(wd[k], 0u, 0x4431));    // U | V << 8 (skewed)
        uint32_t lo, nhi;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(lo) : "r"(tbl + ci));
        asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(nhi) : "r"(tbl + ci), "n"(STRIDE));
        x[k] = lut_pass_pair(wd[k], lo, nhi);
        ragged |= (lo * 256u + nhi == LUT_RAGGED_CODE);
      }
      if (__any_sync(__activemask(), ragged))
      {
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
          const uint32_t ci = __byte_perm(wd[k], 0u, 0x4431);
          if (s_table[lut_phys<SKEW>(ci)] == LUT_RAGGED_LO && s_table[STRIDE + lut_phys<SKEW>(ci)] == LUT_RAGGED_NHI)
            x[k] = lut_pass_pair_masks(wd[k], ci, masks);
        }
      }
      // pass lanes {0,1} are x >> 15; the shifts and all sums ride on IMAD.HI (x * 2^17 >> 32)
      uint32_t Pc = __umulhi(x[0], 1u << 17);
      Pc = __umulhi(x[1], 1u << 17) + Pc;
      Pc = __umulhi(x[2], 1u << 17) + Pc;
      Pc = __umulhi(x[3], 1u << 17) + Pc;
      A = __umulhi(x[1], 1u << 17) + A;
      A = __umulhi(x[2], 2u << 17) + A;
      A = __umulhi(x[3], 3u << 17) + A;
      S += Pc;
      SI += j * Pc;
    };
    // packed lane sums are flushed every 128 iterations (SI <= 4 * 127 * 128 / 2 < 2^16)
    for (int seg0 = 0; seg0 < itersAll; seg0 += 128)
    {
      const int segN = min(128, itersAll - seg0);
      int j = 0;
      for (; j + STAGES <= segN; j += STAGES)
      {
#pragma unroll
        for (int k = 0; k < STAGES; ++k)
          body((uint32_t)k, (uint32_t)(j + k));          // seg0 and j are multiples of STAGES (128 % STAGES == 0)
      }
      for (; j < segN; ++j)
        body((uint32_t)j % STAGES, (uint32_t)j);
      const uint32_t segPasses = (S & 0xFFFFu) + (S >> 16);
      passes += segPasses;
      inIdx  += 2u * ((A & 0xFFFFu) + (A >> 16)) + (S >> 16);
      syPass += segPasses * (uint32_t)(row0 + rr + seg0 * rpi) + (uint32_t)rpi * ((SI & 0xFFFFu) + (SI >> 16));
      S = 0u; A = 0u; SI = 0u;
    }
    uint32_t sxPass = passes * ((uint32_t)cc * 8u) + inIdx;

    // group reduction.  The row loop's trip count differs between the lanes of a warp when cpr is not a multiple of 32:
    // reconverge first, and name the lanes by the group layout (whole warps whenever gthreads % 32 == 0, the usual case)
    __syncwarp();
    const unsigned am = (gthreads & 31) == 0 ? 0xFFFFFFFFu : __activemask();
    passes = __reduce_add_sync(am, passes);
    sxPass = __reduce_add_sync(am, sxPass);
    syPass = __reduce_add_sync(am, syPass);
    if (lane == 0)
    {
      myRed[warpInGroup * 4 + 0] = passes; myRed[warpInGroup * 4 + 1] = sxPass; myRed[warpInGroup * 4 + 2] = syPass;
    }
    group_barrier(1 + group, gthreads);
    if (warpInGroup == 0)
    {
      uint32_t a = 0, b = 0, c = 0;
      if (lane < gwarps) { a = myRed[lane * 4 + 0]; b = myRed[lane * 4 + 1]; c = myRed[lane * 4 + 2]; }
      a = __reduce_add_sync(0xFFFFFFFFu, a);
      b = __reduce_add_sync(0xFFFFFFFFu, b);
      c = __reduce_add_sync(0xFFFFFFFFu, c);
      if (lane == 0)
      {
        bool last = true;
        if (parts > 1)
        {
          SumAcc* const fa = acc + frame;
          atomicAdd(&fa->fails, a); atomicAdd(&fa->sxFail, b); atomicAdd(&fa->syFail, c);      // (sums of the PASSING pixels here)
          __threadfence();
          last = atomicAdd(&fa->done, 1u) == (uint32_t)(parts - 1);
          if (last)
          {
            __threadfence();
            a = atomicExch(&fa->fails, 0u); b = atomicExch(&fa->sxFail, 0u); c = atomicExch(&fa->syFail, 0u);
            fa->done = 0u;
          }
        }
        if (last)         // finalize_sum works from the FAILING pixels' sums (all arithmetic modulo 2^32, as there)
          finalize_sum<KIND_WO>(g, p, W * H - a, H * (W * (W - 1u) / 2u) - b, W * (H * (H - 1u) / 2u) - c, 0u, out + frame, out);
      }
    }
    group_barrier(1 + group, gthreads);                  // myRed is reused by the next frame
  }
}

template <int STAGES, bool SKEW>
__global__ void __launch_bounds__(1024, 1)
wo_lut_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
              const uint8_t* __restrict__ table, const uint32_t* __restrict__ masks, TargetOut* __restrict__ out,
              const int numFrames, const int groups, const int gthreads, const int cpr, const int rpi,
              const int parts, const int rowsPerPart, SumAcc* __restrict__ acc, const int* __restrict__ frameList)
{
  wo_lut_body<STAGES, SKEW>(g, frames, params, table, masks, out, numFrames, groups, gthreads, cpr, rpi, parts, rowsPerPart,
                            acc, frameList, (int)blockIdx.x, (int)gridDim.x);
}

// One launch for a batch whose frames come under several threshold sets: the persistent CTAs are dealt out to the sets in
// proportion to their frames, each CTA loads the table of ITS set and works through that set's frame list.
template <int STAGES, bool SKEW>
__global__ void __launch_bounds__(1024, 1)
wo_lut_sets_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
                   TargetOut* __restrict__ out, const int groups, const int gthreads, const int cpr, const int rpi,
                   SumAcc* __restrict__ acc, const int* __restrict__ frameList, const LutSets sets)
{
  int k = 0;
  while (k + 1 < sets.numSets && (int)blockIdx.x >= sets.ctaStart[k + 1])
    ++k;
  wo_lut_body<STAGES, SKEW>(g, frames, params + sets.paramIndex[k], sets.table[k], sets.masks[k], out, sets.count[k], groups,
                            gthreads, cpr, rpi, sets.parts[k], sets.rowsPerPart[k], acc, frameList + sets.listOffset[k],
                            (int)blockIdx.x - sets.ctaStart[k], sets.ctaStart[k + 1] - sets.ctaStart[k]);
}

// ---------------------------------------------------------------------------------------------
// OO step 1 through the table: threshold -> metapixel bitmap (cv_bitmap_builder_reference.hpp:163-187)
// ---------------------------------------------------------------------------------------------
// bitmap[(row/4) * (W/4) + col/4], bit (row%4)*4 + col%4 = det.  A thread produces the four metapixels under one
// 16-pixel luma chunk: 4 rows x (16 luma + 16 chroma bytes), one 8-byte store.  Persistent CTAs (table in shared
// memory, one per SM), blockDim = cpr * rows: `rows` metapixel rows of the batch at a time, no barrier needed.
template <bool SKEW>
__device__ __forceinline__ void
oo_bitmap_lut_body(const Geometry& g, const uint8_t* __restrict__ frames, const uint8_t* __restrict__ table,
                   const uint32_t* __restrict__ masks, uint16_t* __restrict__ bitmaps,
                   const int numFrames, const int cpr, const int rows, const int* __restrict__ frameList,
                   const int ctaIndex, const int ctaCount)
{
  constexpr uint32_t STRIDE = SKEW ? LUT_STRIDE_SKEW : LUT_STRIDE_PLAIN;
  extern __shared__ __align__(16) uint8_t s_raw[];
  {
    const uint4* src = reinterpret_cast<const uint4*>(table + (SKEW ? LUT_SKEW_IMAGE_OFS : 0u));
    uint4* dst = reinterpret_cast<uint4*>(s_raw);
    for (int i = threadIdx.x; i < (int)(2u * STRIDE / 16u); i += blockDim.x)
      dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= cpr * rows)
    return;
  const int cc = t % cpr, rr = t / cpr;
  const int bw = g.width / 4, bh = g.height / 4;
  const uint32_t tbl = (uint32_t)__cvta_generic_to_shared(s_raw);
  const size_t chromaOfs = (size_t)g.height * g.lineLength;
  const long long units = (long long)numFrames * bh;                 // (frame, metapixel row)

  for (long long u = (long long)ctaIndex * rows + rr; u < units; u += (long long)ctaCount * rows)
  {
    const int slot = (int)(u / bh);
    const int mr = (int)(u - (long long)slot * bh);
    const int frame = frameList ? frameList[slot] : slot;            // a batch under several ranges: this range's frames
    const uint8_t* ptr = frames + (size_t)frame * g.frameStride + (size_t)cc * 16u + (size_t)(mr * 4) * g.lineLength;
    uint4 lu[4], ch[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
      lu[r] = ld_stream(ptr + (size_t)r * g.lineLength);
      ch[r] = ld_stream(ptr + (size_t)r * g.lineLength + chromaOfs);
    }
    uint32_t meta[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
      const uint32_t L[4] = {lu[r].x, lu[r].y, lu[r].z, lu[r].w};
      const uint32_t Cw[4] = {ch[r].x, ch[r].y, ch[r].z, ch[r].w};
      uint32_t x[8];
      bool ragged = false;
#pragma unroll
      for (int k = 0; k < 8; ++k)
      {
        const uint32_t yy = __byte_perm(L[k >> 1], 0u, (k & 1) ? 0x4342 : 0x4140);     // Y0 | Y1 << 16
        const uint32_t ci = lut_phys<SKEW>(__byte_perm(Cw[k >> 1], 0u, (k & 1) ? 0x4423 : 0x4401));   // U | V << 8 (plane bytes: V U V U)
        uint32_t lo, nhi;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(lo) : "r"(tbl + ci));
        asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(nhi) : "r"(tbl + ci), "n"(STRIDE));
        x[k] = lut_pass_pair(yy, lo, nhi);
        ragged |= (lo * 256u + nhi == LUT_RAGGED_CODE);
      }
      if (__any_sync(__activemask(), ragged))
      {
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
          const uint32_t ci = __byte_perm(Cw[k >> 1], 0u, (k & 1) ? 0x4423 : 0x4401);
          if (s_raw[lut_phys<SKEW>(ci)] == LUT_RAGGED_LO && s_raw[STRIDE + lut_phys<SKEW>(ci)] == LUT_RAGGED_NHI)
            x[k] = lut_pass_pair_masks(__byte_perm(L[k >> 1], 0u, (k & 1) ? 0x4342 : 0x4140), ci, masks);
        }
      }
      // bit 15 -> bit s, bit 31 -> bit s + 1 (s = 4r + 2(k&1) <= 14) in ONE IMAD.HI: with M = 2^(17+s) + 2^(s+2),
      // (x * M) >> 32 = a 2^s + b 2^(s+1) + b 2^(16+s) for x = a 2^15 + b 2^31; the stray bit is masked at the end
#pragma unroll
      for (int k = 0; k < 8; ++k)
      {
        const int sft = r * 4 + (k & 1) * 2;
        meta[k >> 1] = __umulhi(x[k], (1u << (17 + sft)) + (1u << (sft + 2))) + meta[k >> 1];
      }
    }
    uint2 v;
    v.x = (meta[0] & 0xFFFFu) | (meta[1] << 16);
    v.y = (meta[2] & 0xFFFFu) | (meta[3] << 16);
    *reinterpret_cast<uint2*>(bitmaps + ((size_t)frame * bh + mr) * bw + cc * 4) = v;
  }
}

template <bool SKEW>
__global__ void __launch_bounds__(768, 1)
oo_bitmap_lut_kernel(const Geometry g, const uint8_t* __restrict__ frames, const uint8_t* __restrict__ table,
                     const uint32_t* __restrict__ masks, uint16_t* __restrict__ bitmaps,
                     const int numFrames, const int cpr, const int rows)
{
  oo_bitmap_lut_body<SKEW>(g, frames, table, masks, bitmaps, numFrames, cpr, rows, nullptr, (int)blockIdx.x, (int)gridDim.x);
}

// frames under several ranges (object sensor instances gathered into one batch, each carrying its own range): the persistent
// CTAs dealt out to the ranges as in wo_lut_sets_kernel
template <bool SKEW>
__global__ void __launch_bounds__(768, 1)
oo_bitmap_lut_sets_kernel(const Geometry g, const uint8_t* __restrict__ frames, uint16_t* __restrict__ bitmaps,
                          const int cpr, const int rows, const int* __restrict__ frameList, const LutSets sets)
{
  int k = 0;
  while (k + 1 < sets.numSets && (int)blockIdx.x >= sets.ctaStart[k + 1])
    ++k;
  oo_bitmap_lut_body<SKEW>(g, frames, sets.table[k], sets.masks[k], bitmaps, sets.count[k], cpr, rows,
                           frameList + sets.listOffset[k], (int)blockIdx.x - sets.ctaStart[k],
                           sets.ctaStart[k + 1] - sets.ctaStart[k]);
}

cudaError_t launch_oo_bitmap_lut_sets(const Geometry& g, const uint8_t* frames, uint16_t* bitmaps, int smCount,
                                      cudaStream_t stream, const int* frameList, LutSets sets)
{
  if (sets.numSets <= 0 || sets.numSets > LUT_MAX_SETS || !frameList)
    return cudaErrorInvalidValue;
  const int cpr = g.width / 16;
  if (cpr <= 0 || cpr > 768)
    return cudaErrorInvalidValue;
  const int rows = 768 / cpr;
  const int threads = ((cpr * rows + 31) / 32) * 32;
  long long total = 0;
  for (int k = 0; k < sets.numSets; ++k) total += sets.count[k];
  if (total <= 0)
    return cudaSuccess;
  sets.ctaStart[0] = 0;
  int used = 0, biggest = 0, ctas[LUT_MAX_SETS];
  for (int k = 0; k < sets.numSets; ++k)
  {
    ctas[k] = (int)((long long)smCount * sets.count[k] / total);
    if (ctas[k] < 1) ctas[k] = 1;
    used += ctas[k];
    if (sets.count[k] > sets.count[biggest]) biggest = k;
  }
  if (used < smCount)
    ctas[biggest] += smCount - used;
  for (int k = 0; k < sets.numSets; ++k)
  {
    const long long need = ((long long)sets.count[k] * (g.height / 4) + rows - 1) / rows;
    if (ctas[k] > need) ctas[k] = (int)need;
    sets.ctaStart[k + 1] = sets.ctaStart[k] + ctas[k];
  }
  const int grid = sets.ctaStart[sets.numSets];
  const bool skew = g_lutSkew != 0;
  const int smem = (int)(2u * (skew ? LUT_STRIDE_SKEW : LUT_STRIDE_PLAIN));
  cudaError_t e = skew ? cudaFuncSetAttribute(oo_bitmap_lut_sets_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                       : cudaFuncSetAttribute(oo_bitmap_lut_sets_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess)
    return e;
  if (skew)
    oo_bitmap_lut_sets_kernel<true><<<grid, threads, smem, stream>>>(g, frames, bitmaps, cpr, rows, frameList, sets);
  else
    oo_bitmap_lut_sets_kernel<false><<<grid, threads, smem, stream>>>(g, frames, bitmaps, cpr, rows, frameList, sets);
  ++g_launches_lut;
  return cudaGetLastError();
}

cudaError_t launch_oo_bitmap_lut(const Geometry& g, int numFrames, const uint8_t* frames, const uint8_t* table,
                                 const uint32_t* masks, uint16_t* bitmaps, int smCount, cudaStream_t stream)
{
  if (numFrames <= 0)
    return cudaSuccess;
  const int cpr = g.width / 16;
  if (cpr <= 0 || cpr > 768)
    return cudaErrorInvalidValue;
  const int rows = 768 / cpr;
  const int threads = ((cpr * rows + 31) / 32) * 32;
  const long long units = (long long)numFrames * (g.height / 4);
  long long grid = (units + rows - 1) / rows;
  if (grid > smCount) grid = smCount;
  const bool skew = g_lutSkew != 0;
  const int smem = (int)(2u * (skew ? LUT_STRIDE_SKEW : LUT_STRIDE_PLAIN));
  cudaError_t e = skew ? cudaFuncSetAttribute(oo_bitmap_lut_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                       : cudaFuncSetAttribute(oo_bitmap_lut_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess)
    return e;
  if (skew)
    oo_bitmap_lut_kernel<true><<<(unsigned)grid, threads, smem, stream>>>(g, frames, table, masks, bitmaps, numFrames, cpr, rows);
  else
    oo_bitmap_lut_kernel<false><<<(unsigned)grid, threads, smem, stream>>>(g, frames, table, masks, bitmaps, numFrames, cpr, rows);
  ++g_launches_lut;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// self-check used by the parity tests: table-based detection against the arithmetic on all 2^24 inputs
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lut_check_kernel(const uint32_t from, const uint32_t to, const uint32_t expected,
                 const uint8_t* __restrict__ table, const uint32_t* __restrict__ masks,
                 unsigned long long* __restrict__ stats)     // [0] mismatching pixels, [1] never, [2] interval, [3] ragged entries, [4] passing pixels
{
  __shared__ HueLutEntry s_lutHue[256];
  __shared__ uint16_t s_lut43[256];
  __shared__ uint16_t s_lut255[256];
  fill_div_luts(s_lut43, s_lut255);
  fill_hue_lut(s_lutHue);
  __syncthreads();
  const HsvBounds bd = make_bounds(from, to);
  // one thread per (pair of luma values, chroma pair): index = pairY | U << 7 | V << 15
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t y0 = (i & 127u) * 2u, u = (i >> 7) & 0xFFu, v = i >> 15;
  const uint32_t word = y0 | (u << 8) | ((y0 + 1u) << 16) | (v << 24);
  const uint32_t det = detect_pair_bits(word & 0x00FF00FFu, word, coef_yuyv(), s_lutHue, s_lut255, bd, expected);
  const uint32_t ci = __byte_perm(word, 0u, 0x4431);
  const uint32_t lo = table[ci], nhi = table[65536u + ci];
  const bool isRagged = lo == LUT_RAGGED_LO && nhi == LUT_RAGGED_NHI;
  const bool isNever = lo == LUT_NEVER_LO && nhi == LUT_NEVER_NHI;
  const uint32_t x = isRagged ? lut_pass_pair_masks(word, ci, masks) : lut_pass_pair(word, lo, nhi);
  const uint32_t passLut = ((x >> 15) & 1u) | ((x >> 31) << 1);
  const uint32_t bad = (uint32_t)__popc(passLut ^ det);
  if (bad) atomicAdd(stats + 0, (unsigned long long)bad);
  if (det) atomicAdd(stats + 4, (unsigned long long)__popc(det));
  if ((i & 127u) == 0u)
    atomicAdd(stats + (isNever ? 1 : (isRagged ? 3 : 2)), 1ull);
}

cudaError_t launch_lut_check(uint32_t from, uint32_t to, uint32_t expected, const uint8_t* table, const uint32_t* masks,
                             unsigned long long* stats, cudaStream_t stream)
{
  lut_check_kernel<<<(1u << 23) / 256, 256, 0, stream>>>(from, to, expected, table, masks, stats);
  ++g_launches_lut;
  return cudaGetLastError();
}

// bands per frame: the fewest of 1, 2, 4, 8 that fill the last round of `slots` persistent groups to 90 % (a band is at least
// 2 * STAGES row iterations), else the best of them
static int lut_bands(int numFrames, long long slots, int height, int rpi, int stages)
{
  int parts = 1;
  if (g_lutParts == 1)
    return 1;
  if (g_lutParts > 1)                       // forced (measurements, tests): as asked, down to one row iteration per band
  {
    int forced = g_lutParts > 8 ? 8 : g_lutParts;
    while (forced > 1 && height / forced < rpi) forced >>= 1;
    return forced;
  }
  double best = 0.0;
  for (int cand = 1; cand <= 8; cand *= 2)
  {
    if (cand > 1 && height / cand < rpi * 2 * stages)
      break;
    const long long items = (long long)numFrames * cand;
    const double eff = items <= slots ? 1.0 : (double)items / slots / (double)((items + slots - 1) / slots);
    if (eff > best + 1e-9) { best = eff; parts = cand; }
    if (eff >= 0.9)
      break;
  }
  return parts;
}

cudaError_t launch_wo_lut(const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                          const uint8_t* table, const uint32_t* masks, TargetOut* out, int smCount, cudaStream_t stream,
                          SumAcc* acc, const int* frameList)
{
  if (numFrames <= 0)
    return cudaSuccess;
  const int gthreads = sum_sensor_block_threads(KIND_WO, g.width);
  if (gthreads <= 0 || gthreads > 1024)
    return cudaErrorInvalidValue;
  const int cpr = g.width / 8;
  const int rpi = gthreads / cpr;
  constexpr int STAGES = 4;
  int groups = 1024 / gthreads;
  if (groups > 15) groups = 15;                           // named barriers 1..15
  // shared memory: table + ring + reduction scratch
  const bool skew = g_lutSkew != 0;
  const size_t tableBytes = 2u * (skew ? LUT_STRIDE_SKEW : LUT_STRIDE_PLAIN);
  auto smem_for = [&](int ng) { return tableBytes + (size_t)STAGES * ng * gthreads * 16 + (size_t)ng * 128 * 4; };
  while (groups > 1 && smem_for(groups) > 220 * 1024) --groups;
  if (smem_for(groups) > 227 * 1024)
    return cudaErrorInvalidValue;
  const int threads = ((groups * gthreads + 31) / 32) * 32;
  // bands per frame: the fewest of 1, 2, 4, 8 that fill the last round of the persistent groups to 90 % (a band is at least
  // 2 * STAGES row iterations), else the best of them; needs the accumulator records
  const int parts = acc ? lut_bands(numFrames, (long long)smCount * groups, g.height, rpi, STAGES) : 1;
  const int rowsPerPart = (((g.height + parts - 1) / parts + rpi - 1) / rpi) * rpi;     // whole row iterations per band
  int grid = (int)(((long long)numFrames * parts + groups - 1) / groups);
  if (grid > smCount) grid = smCount;
  const size_t smem = smem_for(groups);
  cudaError_t e = skew ? cudaFuncSetAttribute(wo_lut_kernel<STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                       : cudaFuncSetAttribute(wo_lut_kernel<STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess)
    return e;
  if (skew)
    wo_lut_kernel<STAGES, true><<<grid, threads, smem, stream>>>(g, frames, params, table, masks, out, numFrames, groups, gthreads, cpr, rpi, parts, rowsPerPart, acc, frameList);
  else
    wo_lut_kernel<STAGES, false><<<grid, threads, smem, stream>>>(g, frames, params, table, masks, out, numFrames, groups, gthreads, cpr, rpi, parts, rowsPerPart, acc, frameList);
  ++g_launches_lut;
  return cudaGetLastError();
}

cudaError_t launch_wo_lut_sets(const Geometry& g, const uint8_t* frames, const FrameParams* params, TargetOut* out, int smCount,
                               cudaStream_t stream, SumAcc* acc, const int* frameList, LutSets sets)
{
  if (sets.numSets <= 0 || sets.numSets > LUT_MAX_SETS || !acc || !frameList)
    return cudaErrorInvalidValue;
  const int gthreads = sum_sensor_block_threads(KIND_WO, g.width);
  if (gthreads <= 0 || gthreads > 1024)
    return cudaErrorInvalidValue;
  const int cpr = g.width / 8;
  const int rpi = gthreads / cpr;
  constexpr int STAGES = 4;
  int groups = 1024 / gthreads;
  if (groups > 15) groups = 15;
  const bool skew = g_lutSkew != 0;
  const size_t tableBytes = 2u * (skew ? LUT_STRIDE_SKEW : LUT_STRIDE_PLAIN);
  auto smem_for = [&](int ng) { return tableBytes + (size_t)STAGES * ng * gthreads * 16 + (size_t)ng * 128 * 4; };
  while (groups > 1 && smem_for(groups) > 220 * 1024) --groups;
  if (smem_for(groups) > 227 * 1024)
    return cudaErrorInvalidValue;
  const int threads = ((groups * gthreads + 31) / 32) * 32;
  // the SMs (one persistent CTA each) dealt out in proportion to the sets' frames, at least one per set
  long long total = 0;
  for (int k = 0; k < sets.numSets; ++k) total += sets.count[k];
  if (total <= 0)
    return cudaSuccess;
  int ctas[LUT_MAX_SETS], used = 0, biggest = 0;
  for (int k = 0; k < sets.numSets; ++k)
  {
    ctas[k] = (int)((long long)smCount * sets.count[k] / total);
    if (ctas[k] < 1) ctas[k] = 1;
    const int need = (sets.count[k] + groups - 1) / groups;           // no more CTAs than whole-frame items
    if (ctas[k] > need * 8) ctas[k] = need * 8;
    used += ctas[k];
    if (sets.count[k] > sets.count[biggest]) biggest = k;
  }
  if (used < smCount)
    ctas[biggest] += smCount - used;
  sets.ctaStart[0] = 0;
  for (int k = 0; k < sets.numSets; ++k)
  {
    sets.parts[k] = lut_bands(sets.count[k], (long long)ctas[k] * groups, g.height, rpi, STAGES);
    sets.rowsPerPart[k] = (((g.height + sets.parts[k] - 1) / sets.parts[k] + rpi - 1) / rpi) * rpi;
    const int need = (int)(((long long)sets.count[k] * sets.parts[k] + groups - 1) / groups);
    if (ctas[k] > need) ctas[k] = need;
    sets.ctaStart[k + 1] = sets.ctaStart[k] + ctas[k];
  }
  const int grid = sets.ctaStart[sets.numSets];
  const size_t smem = smem_for(groups);
  cudaError_t e = skew ? cudaFuncSetAttribute(wo_lut_sets_kernel<STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                       : cudaFuncSetAttribute(wo_lut_sets_kernel<STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess)
    return e;
  if (skew)
    wo_lut_sets_kernel<STAGES, true><<<grid, threads, smem, stream>>>(g, frames, params, out, groups, gthreads, cpr, rpi, acc, frameList, sets);
  else
    wo_lut_sets_kernel<STAGES, false><<<grid, threads, smem, stream>>>(g, frames, params, out, groups, gthreads, cpr, rpi, acc, frameList, sets);
  ++g_launches_lut;
  return cudaGetLastError();
}

} // namespace trikb200
