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
//                           table[U | V<<8] = lo | hi<<8   when the mask is exactly the interval lo..hi,
//                                             NEVER        when no luma passes,
//                                             RAGGED       otherwise (rounding jitter near a hue / saturation
//                                                          bound: 1..10 % of the entries),
//                         and the 256-bit masks themselves (2 MB, L2 resident) for the RAGGED entries.
//   wo_lut_kernel         the WO pass with the 128 KB table in shared memory: per pixel pair one 16-bit LDS and
//                         two packed compares instead of the HSV arithmetic; a warp that meets a RAGGED entry
//                         fetches the mask word of those pixels from L2.  ~17 instructions per pair, the same
//                         budget as the line kernels.  The CTA is persistent (one per SM, the table is loaded
//                         once) and is split into independent groups of threads, one frame per group at a
//                         time, synchronised by named barriers.
//
// Results are bit-identical to the arithmetic path by construction (the table IS that path, tabulated), and
// tests/test_lut_gpu.py checks it on all 2^24 inputs per threshold set and on frames.
#include <atomic>
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"
#include "trik_line.cuh"

namespace trikb200 {

std::atomic<long long> g_launches_lut{0};

constexpr uint32_t LUT_NEVER  = 0x00FFu;      // lo = 255, hi = 0: no luma can satisfy lo <= Y <= hi
constexpr uint32_t LUT_RAGGED = 0x01FFu;      // lo = 255, hi = 1: fails the interval test too, marks "consult the mask"

// ---------------------------------------------------------------------------------------------
// table construction: one warp per chroma pair
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
chroma_table_kernel(const uint32_t from, const uint32_t to, const uint32_t expected,
                    uint16_t* __restrict__ table, uint32_t* __restrict__ masks)
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
    table[idx] = (uint16_t)(count == 0u ? LUT_NEVER : (rises == 1u ? (lo | (hi << 8)) : LUT_RAGGED));
  if (lane < 8u)
  {
    uint32_t mine = w[0];
#pragma unroll
    for (int m = 1; m < 8; ++m)
      if (lane == (uint32_t)m) mine = w[m];
    masks[(size_t)idx * 8u + lane] = mine;
  }
}

cudaError_t launch_chroma_table(uint32_t from, uint32_t to, uint32_t expected, uint16_t* table, uint32_t* masks,
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

// fail lanes {0,1} of a YUYV pixel pair from its table entry: pass <=> lo <= Y <= hi.
// Guard bit 15 in every lane keeps the two 16-bit subtractions of one 32-bit SUB apart (no borrow can leave a
// lane: 0x8000 + Y - lo >= 0x7F01) and doubles as the result: bit 15 stays set <=> the difference is >= 0.
__device__ __forceinline__ uint32_t lut_fail_pair(uint32_t word, uint32_t entry)
{
  const uint32_t yyG = (word & 0x00FF00FFu) | 0x80008000u;
  const uint32_t lo2 = __byte_perm(entry, 0u, 0x4040);                 // lo in both lanes
  const uint32_t hiG = __byte_perm(entry, 0x00000080u, 0x4141);        // 0x8000 + hi in both lanes
  const uint32_t geLo = yyG - lo2;                                     // bit 15 of a lane <=> Y >= lo
  const uint32_t leHi = hiG - (word & 0x00FF00FFu);                    // bit 15 of a lane <=> Y <= hi
  return (~(geLo & leHi) & 0x80008000u) >> 15;
}

// Persistent CTA: `groups` independent groups of `gthreads` threads, one frame per group at a time.
// Inside a group the decomposition is sum_kernel's: a thread owns a 16-byte column chunk and walks down the rows.
template <int STAGES>
__global__ void __launch_bounds__(1024, 1)
wo_lut_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
              const uint16_t* __restrict__ table, const uint32_t* __restrict__ masks, TargetOut* __restrict__ out,
              const int numFrames, const int groups, const int gthreads, const int cpr, const int rpi)
{
  extern __shared__ __align__(16) uint8_t s_raw[];
  uint16_t* const s_table = reinterpret_cast<uint16_t*>(s_raw);                       // 65 536 entries
  uint4* const s_ring = reinterpret_cast<uint4*>(s_raw + 131072);                     // [STAGES][groups * gthreads]
  uint32_t* const s_red = reinterpret_cast<uint32_t*>(s_raw + 131072 + (size_t)STAGES * groups * gthreads * 16);   // [groups][32][4]

  // the table, once per CTA
  {
    const uint4* src = reinterpret_cast<const uint4*>(table);
    uint4* dst = reinterpret_cast<uint4*>(s_table);
    for (int i = threadIdx.x; i < 131072 / 16; i += blockDim.x)
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
  const int itersAll = rr < g.height ? (g.height - rr + rpi - 1) / rpi : 0;
  const size_t rowStep = (size_t)rpi * g.lineLength;
  const int ringThreads = groups * gthreads;
  uint4* const mySlot = s_ring + t;
  uint32_t* const myRed = s_red + group * 128;
  const FrameParams p = params[0];

  for (int frame = blockIdx.x * groups + group; frame < numFrames; frame += gridDim.x * groups)
  {
    const uint8_t* fillPtr = frames + (size_t)frame * g.frameStride + (size_t)cc * 16u + (size_t)rr * g.lineLength;
    uint32_t fails = 0u, inIdx = 0u, syFail = 0u;        // 32-bit totals of this thread
    int fillIt = 0;
#pragma unroll
    for (int sIdx = 0; sIdx < STAGES - 1; ++sIdx)
    {
      if (fillIt < itersAll)
        cp_async16(mySlot + sIdx * ringThreads, fillPtr);
      cp_async_commit();
      ++fillIt;
      fillPtr += rowStep;
    }
    // packed 16-bit lane sums are flushed every 128 iterations (SI <= 8 * 128 * 127 / 2 < 2^16)
    for (int seg0 = 0; seg0 < itersAll; seg0 += 128)
    {
      const int segN = min(128, itersAll - seg0);
      uint32_t S = 0u, A = 0u, SI = 0u;
      for (int j = 0; j < segN; ++j)
      {
        const int it = seg0 + j;
        if (fillIt < itersAll)
          cp_async16(mySlot + ((it + STAGES - 1) % STAGES) * ringThreads, fillPtr);
        cp_async_commit();
        ++fillIt;
        fillPtr += rowStep;
        cp_async_wait<STAGES - 1>();
        const uint4 cur = mySlot[(it % STAGES) * ringThreads];
        const uint32_t wd[4] = {cur.x, cur.y, cur.z, cur.w};
        uint32_t fl[4];
        bool ragged = false;
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
          const uint32_t e = s_table[__byte_perm(wd[k], 0u, 0x4431)];          // U | V << 8
          fl[k] = lut_fail_pair(wd[k], e);
          ragged |= (e == LUT_RAGGED);
        }
        if (__any_sync(__activemask(), ragged))
        {
#pragma unroll
          for (int k = 0; k < 4; ++k)
          {
            const uint32_t ci = __byte_perm(wd[k], 0u, 0x4431);
            if (s_table[ci] == LUT_RAGGED)
            {
              const uint32_t y0 = wd[k] & 0xFFu, y1 = (wd[k] >> 16) & 0xFFu;
              const uint32_t m0 = __ldg(masks + (size_t)ci * 8u + (y0 >> 5));
              const uint32_t m1 = __ldg(masks + (size_t)ci * 8u + (y1 >> 5));
              const uint32_t pass = ((m0 >> (y0 & 31u)) & 1u) | (((m1 >> (y1 & 31u)) & 1u) << 16);
              fl[k] = 0x00010001u - pass;                                    // the interval test said "fail, fail"
            }
          }
        }
        const uint32_t Sc = (fl[0] + fl[1]) + (fl[2] + fl[3]);
        A  += fl[1] + 2u * fl[2] + 3u * fl[3];
        S  += Sc;
        SI += (uint32_t)j * Sc;
      }
      const uint32_t segFails = (S & 0xFFFFu) + (S >> 16);
      fails  += segFails;
      inIdx  += 2u * ((A & 0xFFFFu) + (A >> 16)) + (S >> 16);
      syFail += segFails * (uint32_t)(rr + seg0 * rpi) + (uint32_t)rpi * ((SI & 0xFFFFu) + (SI >> 16));
    }
    uint32_t sxFail = fails * ((uint32_t)cc * 8u) + inIdx;

    // group reduction
    const unsigned am = __activemask();
    fails  = __reduce_add_sync(am, fails);
    sxFail = __reduce_add_sync(am, sxFail);
    syFail = __reduce_add_sync(am, syFail);
    if (lane == 0)
    {
      myRed[warpInGroup * 4 + 0] = fails; myRed[warpInGroup * 4 + 1] = sxFail; myRed[warpInGroup * 4 + 2] = syFail;
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
        finalize_sum<KIND_WO>(g, p, a, b, c, 0u, out + frame, out);
    }
    group_barrier(1 + group, gthreads);                  // myRed is reused by the next frame
  }
}

// ---------------------------------------------------------------------------------------------
// self-check used by the parity tests: table-based detection against the arithmetic on all 2^24 inputs
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lut_check_kernel(const uint32_t from, const uint32_t to, const uint32_t expected,
                 const uint16_t* __restrict__ table, const uint32_t* __restrict__ masks,
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
  const uint32_t e = table[ci];
  uint32_t fl = lut_fail_pair(word, e);
  if (e == LUT_RAGGED)
  {
    const uint32_t y1 = y0 + 1u;
    const uint32_t m0 = masks[(size_t)ci * 8u + (y0 >> 5)], m1 = masks[(size_t)ci * 8u + (y1 >> 5)];
    fl = 0x00010001u - (((m0 >> (y0 & 31u)) & 1u) | (((m1 >> (y1 & 31u)) & 1u) << 16));
  }
  const uint32_t passLut = ((fl & 1u) ^ 1u) | ((((fl >> 16) & 1u) ^ 1u) << 1);
  const uint32_t bad = (uint32_t)__popc(passLut ^ det);
  if (bad) atomicAdd(stats + 0, (unsigned long long)bad);
  if (det) atomicAdd(stats + 4, (unsigned long long)__popc(det));
  if ((i & 127u) == 0u)
    atomicAdd(stats + (e == LUT_NEVER ? 1 : (e == LUT_RAGGED ? 3 : 2)), 1ull);
}

cudaError_t launch_lut_check(uint32_t from, uint32_t to, uint32_t expected, const uint16_t* table, const uint32_t* masks,
                             unsigned long long* stats, cudaStream_t stream)
{
  lut_check_kernel<<<(1u << 23) / 256, 256, 0, stream>>>(from, to, expected, table, masks, stats);
  ++g_launches_lut;
  return cudaGetLastError();
}

cudaError_t launch_wo_lut(const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                          const uint16_t* table, const uint32_t* masks, TargetOut* out, int smCount, cudaStream_t stream)
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
  auto smem_for = [&](int ng) { return (size_t)131072 + (size_t)STAGES * ng * gthreads * 16 + (size_t)ng * 128 * 4; };
  while (groups > 1 && smem_for(groups) > 220 * 1024) --groups;
  if (smem_for(groups) > 227 * 1024)
    return cudaErrorInvalidValue;
  const int threads = ((groups * gthreads + 31) / 32) * 32;
  int grid = (numFrames + groups - 1) / groups;
  if (grid > smCount) grid = smCount;
  const size_t smem = smem_for(groups);
  cudaError_t e = cudaFuncSetAttribute(wo_lut_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess)
    return e;
  wo_lut_kernel<STAGES><<<grid, threads, smem, stream>>>(g, frames, params, table, masks, out, numFrames, groups, gthreads, cpr, rpi);
  ++g_launches_lut;
  return cudaGetLastError();
}

} // namespace trikb200
