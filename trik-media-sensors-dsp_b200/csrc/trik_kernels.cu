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
#include <atomic>
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"
#include "trik_line.cuh"

namespace trikb200 {

static std::atomic<long long> g_launches{0};
// Tuning state.  Default (-1): the measured best per sensor on B200 --
//   WL: tuned line kernel, cp.async ring of 4;  OL: first-version kernel, ring of 2;  WO: register prefetch.
// trikb200_setLoadStages(v): v % 100 = load path (0 register prefetch, 2 / 4 ring depth),
//   v / 100 = 0 per-sensor default kernel, 1 force the first-version kernel, 2 force the tuned line kernel.
static int g_targetThreads = 0;             // 0 = default (about 256 threads per CTA)
void set_target_threads(int t) { g_targetThreads = t; }
static int g_framesPerCta = 0;               // wide OL kernel: 0 = heuristic, 1 = off, n = force
void set_frames_per_cta(int n) { g_framesPerCta = n; }
static int g_overlapLaunch = 1;              // programmatic dependent launch of the tuned line kernel
void set_overlap_launch(int on) { g_overlapLaunch = on ? 1 : 0; }
static int g_tuneStages = -1;
static int g_tuneKernel = 0;
// resolved per launch from the knobs above; thread_local so that host threads driving different handles do not race
static thread_local bool g_legacyLineKernel = false;
static thread_local bool g_widePlanarKernel = false;      // resolved per launch: OL through vsum16_kernel
static thread_local bool g_bulkLineKernel = false;        // resolved per launch: WL / OL through tsum_kernel (trik_kernels_line.cu)
static thread_local int g_sumStages = 0;
void set_sum_stages(int v)
{
  if (v < 0) { g_tuneStages = -1; g_tuneKernel = 0; return; }
  g_tuneStages = v % 100;
  g_tuneKernel = v / 100;
}
static void resolve_tuning(int kind, int width)
{
  // measured defaults (profiles/): WL -> vsum_kernel, ring of 4; OL -> vsum16_kernel, ring of 4, from 320 columns
  // up (narrower rows leave a thread too few iterations: first-version kernel, ring of 2); WO -> register prefetch
  int kernel = g_tuneKernel;
  if (kernel == 0)
    kernel = kind == KIND_WL ? 2 : (kind == KIND_OL ? (width >= 320 ? 3 : 1) : 1);
  g_legacyLineKernel = kernel == 1;
  g_widePlanarKernel = kind == KIND_OL && kernel == 3;
  g_bulkLineKernel = (kind == KIND_OL || kind == KIND_WL) && kernel == 4;
  if (g_tuneStages >= 0)
    g_sumStages = g_tuneStages;
  else
    g_sumStages = kind == KIND_WL ? 4 : (kind == KIND_OL ? (g_widePlanarKernel ? 4 : 2) : 0);
}
extern std::atomic<long long> g_launches_grid;
extern std::atomic<long long> g_launches_detect;
extern std::atomic<long long> g_launches_preview;
extern std::atomic<long long> g_launches_line;
extern std::atomic<long long> g_launches_lut;
extern std::atomic<long long> g_launches_anneal;
extern std::atomic<long long> g_launches_omtab;
extern std::atomic<long long> g_launches_ingest;
extern std::atomic<long long> g_launches_edge;
long long launch_count() { return g_launches + g_launches_grid + g_launches_detect + g_launches_preview + g_launches_line + g_launches_lut + g_launches_anneal + g_launches_omtab + g_launches_ingest + g_launches_edge; }

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
                                                 const HueLutEntry* lutHue, const uint16_t* lut255,
                                                 const HsvBounds& bd, uint32_t expected)
{
  const uint32_t det = detect_pair_bits(yy, cw, coef, lutHue, lut255, bd, expected);
  return ((det & 1u) ^ 1u) | (((det >> 1) ^ 1u) << 16);
}

// ---------------------------------------------------------------------------------------------
// the sum kernel
// ---------------------------------------------------------------------------------------------
// STAGES == 0: the next chunk is prefetched into registers (one load in flight per thread).
// STAGES >= 2: every thread keeps a PRIVATE ring of STAGES chunks in shared memory filled by
//   cp.async (LDGSTS, L2-only .cg): STAGES-1 loads in flight per thread with no register cost, and
//   because a thread only ever reads the slots it filled itself, no CTA barrier is needed --
//   cp.async.wait_group orders the thread's own copies.
template <int KIND, int STAGES>
__global__ void __launch_bounds__(1024)
sum_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
           const int paramStride, SumAcc* __restrict__ acc, TargetOut* __restrict__ out,
           const int slabs, const int rowsPerSlab, const int cpr, const int rpi)
{
  constexpr bool PLANAR = (KIND == KIND_OL);
  __shared__ HueLutEntry s_lutHue[256];
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
    fill_hue_lut(s_lutHue);
    __syncthreads();
  }
  const HsvBounds bd = make_bounds(p.from, p.to);

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
  extern __shared__ uint4 s_ring[];          // STAGES x blockDim (x2 planes for YUV422P), thread-private slots
  constexpr int PLANES = PLANAR ? 2 : 1;
  uint4 cur = make_uint4(0, 0, 0, 0), curC = make_uint4(0, 0, 0, 0);
  int fillRow = row;                         // next row to request (STAGES > 0)
  const uint8_t* fillPtr = ptr;
  if (STAGES == 0)
  {
    if (row < r1)
    {
      cur = ld_stream(ptr);
      if (PLANAR) curC = ld_stream(ptr + chromaOfs);
    }
  }
  else
  {
#pragma unroll
    for (int sIdx = 0; sIdx < (STAGES > 0 ? STAGES - 1 : 0); ++sIdx)
    {
      if (fillRow < r1)
      {
        cp_async16(&s_ring[(sIdx * PLANES) * blockDim.x + t], fillPtr);
        if (PLANAR) cp_async16(&s_ring[(sIdx * PLANES + 1) * blockDim.x + t], fillPtr + chromaOfs);
      }
      cp_async_commit();
      fillRow += rpi;
      fillPtr += rowStep;
    }
  }
  for (uint32_t it = 0; row < r1; ++it)
  {
    uint4 nxt = make_uint4(0, 0, 0, 0), nxtC = make_uint4(0, 0, 0, 0);
    const int nrow = row + rpi;
    if (STAGES == 0)
    {
      // prefetch the next row's chunk before working on this one
      if (nrow < r1)
      {
        nxt = ld_stream(ptr + rowStep);
        if (PLANAR) nxtC = ld_stream(ptr + rowStep + chromaOfs);
      }
    }
    else
    {
      constexpr int ST = STAGES > 0 ? STAGES : 1;
      const int fillSlot = (int)((it + ST - 1) % ST);
      if (fillRow < r1)
      {
        cp_async16(&s_ring[(fillSlot * PLANES) * blockDim.x + t], fillPtr);
        if (PLANAR) cp_async16(&s_ring[(fillSlot * PLANES + 1) * blockDim.x + t], fillPtr + chromaOfs);
      }
      cp_async_commit();
      fillRow += rpi;
      fillPtr += rowStep;
      cp_async_wait<STAGES - 1>();
      const int slot = (int)(it % ST);
      cur = s_ring[(slot * PLANES) * blockDim.x + t];
      if (PLANAR) curC = s_ring[(slot * PLANES + 1) * blockDim.x + t];
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
          fl = hsvfail_pair(yy, w[k], coef_yuyv(), s_lutHue, s_lut255, bd, p.expected);
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

    if (STAGES == 0)
    {
      cur = nxt; curC = nxtC;
    }
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
        finalize_sum<KIND>(g, p, a, b, c, d, out + frame, out);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the line-sensor kernel (WL, OL): V-only threshold, tuned instruction mix
// ---------------------------------------------------------------------------------------------
// ncu on the first version showed the ALU pipe (LOP3/PRMT/VIADD/VIMNMX, half rate) as the busiest
// unit with the FMA pipe (IMAD/IDP.4A) at half its load.  This version moves everything that can
// move onto the FMA pipe and trims the bookkeeping:
//   * the blue chroma term is replicated into both lanes by an IMAD (x 0x10001) instead of a PRMT.
//   * the fail test leaves {N, N+1} in each lane (pass, fail) instead of {0,1}; lane-isolated
//     VIADD.16x2 accumulates that per pair position, and the known pass contribution
//     iterations*N is subtracted modulo 2^16 at the end.  No per-word fix-up instructions.
//   * OL: the column window is folded into the per-lane cap constant, and the cross band (rows
//     hStart..hStop, a contiguous run of a thread's iterations) is the difference of two snapshots
//     of the accumulators -- nothing per pixel.
// Per two pixels: 7 FMA-pipe + 6 ALU-pipe instructions in both layouts.
__device__ __forceinline__ void cp_async8(void* smemDst, const void* gmemSrc)
{
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gmemSrc) : "memory");
}
__device__ __forceinline__ uint2 ld_stream8(const void* p)
{
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// A thread owns 8 pixels of a row in BOTH layouts: one 16-byte YUYV chunk, or 8 luma bytes + the 8
// chroma bytes under them (YUV422P); either way the chunk is four 32-bit "pair" slots in one uint4
// (YUV422P: x,y = luma words, z,w = chroma words), so the cp.async ring is identical.
// MAXT = 512 asks ptxas for three resident CTAs of up to 512 threads (<= 40 registers per thread), which
// is what the usual 256..512-thread configurations want; MAXT = 1024 is the catch-all for very wide rows.
template <bool PLANAR, int STAGES, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 512 ? 3 : 1)
vsum_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
            const int paramStride, SumAcc* __restrict__ acc, TargetOut* __restrict__ out,
            const int slabs, const int rowsPerSlab, const int cpr, const int rpi)
{
  __shared__ uint32_t s_red[32][4];
  extern __shared__ uint4 s_ring[];

  // Programmatic dependent launch: let the next kernel of the stream start placing its CTAs while this
  // grid drains, and (as such a successor) wait for everything before us in the stream to be complete
  // and visible before the first global access -- stream order as the caller sees it is unchanged.
  asm volatile("griddepcontrol.launch_dependents;");
  const int frame = blockIdx.x / slabs;
  const int slab  = blockIdx.x - frame * slabs;
  const int t  = threadIdx.x;
  const int cc = t % cpr;
  const int rr = t / cpr;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const FrameParams p = params[(size_t)frame * paramStride];

  const int r0 = slab * rowsPerSlab;
  const int r1 = min(r0 + rowsPerSlab, g.height);
  const int firstRow = r0 + rr;
  const int iters = firstRow < r1 ? (r1 - firstRow + rpi - 1) / rpi : 0;
  const uint8_t* ptr = frames + (size_t)frame * g.frameStride + (size_t)cc * (PLANAR ? 8u : 16u)
                     + (size_t)firstRow * g.lineLength;
  const size_t rowStep = (size_t)rpi * g.lineLength;
  const size_t chromaOfs = (size_t)g.height * g.lineLength;

  const uint32_t negKlo2 = p.negKlo2, n2 = p.n2;
  const uint32_t nLane = n2 & 0xFFFFu;
  const uint32_t np1 = n2 + 0x00010001u;               // N + 1 in both lanes (N <= 65534 by construction)
  // per-position cap: N+1 where the pixel counts, N (== "always pass") where OL's column window
  // (columns 5..W-5) excludes it: chunk 0 loses its pixels 0..4, the last chunk its pixels 4..7
  uint32_t cap01 = np1, cap2 = np1, cap3 = np1;
  if (PLANAR)
  {
    if (cc == 0)       { cap01 = n2; cap2 = (np1 & 0xFFFF0000u) | nLane; }
    if (cc == cpr - 1) { cap2 = n2; cap3 = n2; }
  }

  // OL cross band as a run of this thread's iterations [itA, itB)
  int itA = 0, itB = 0;
  if (PLANAR && p.hStart <= p.hStop)
  {
    // rows past the image do not exist: clamp the band to it first, then 32-bit arithmetic is enough
    const int a = (int)min(p.hStart, (uint32_t)g.height) - firstRow;
    const int b = (int)min(p.hStop, (uint32_t)g.height - 1u) + 1 - firstRow;
    itA = a <= 0 ? 0 : min(iters, (a + rpi - 1) / rpi);
    itB = b <= 0 ? 0 : min(iters, (b + rpi - 1) / rpi);
  }

  uint32_t S0 = 0u, S1 = 0u, S2 = 0u, S3 = 0u;
  uint32_t snapA = 0u, snapB = 0u;

  const uint8_t* fillPtr = ptr;
  int fillIt = 0;
  uint4 cur = make_uint4(0, 0, 0, 0);
  auto request = [&](uint4* slot, const uint8_t* src)
  {
    if (PLANAR)
    {
      cp_async8(slot, src);
      cp_async8(reinterpret_cast<uint8_t*>(slot) + 8, src + chromaOfs);
    }
    else
      cp_async16(slot, src);
  };
  auto load_now = [&](const uint8_t* src) -> uint4
  {
    if (PLANAR)
    {
      const uint2 l = ld_stream8(src), c = ld_stream8(src + chromaOfs);
      return make_uint4(l.x, l.y, c.x, c.y);
    }
    return ld_stream(src);
  };
  if (STAGES == 0)
  {
    if (iters > 0)
      cur = load_now(ptr);
  }
  else
  {
#pragma unroll
    for (int sIdx = 0; sIdx < (STAGES > 0 ? STAGES - 1 : 0); ++sIdx)
    {
      if (fillIt < iters)
        request(&s_ring[sIdx * blockDim.x + t], fillPtr);
      cp_async_commit();
      ++fillIt;
      fillPtr += rowStep;
    }
  }

  for (int it = 0; it < iters; ++it)
  {
    uint4 nxt = make_uint4(0, 0, 0, 0);
    if (STAGES == 0)
    {
      if (it + 1 < iters)
        nxt = load_now(ptr + rowStep);
      ptr += rowStep;
    }
    else
    {
      constexpr int ST = STAGES > 0 ? STAGES : 1;
      if (fillIt < iters)
        request(&s_ring[((it + ST - 1) % ST) * blockDim.x + t], fillPtr);
      cp_async_commit();
      ++fillIt;
      fillPtr += rowStep;
      cp_async_wait<STAGES - 1>();
      cur = s_ring[(it % ST) * blockDim.x + t];
    }

    if (PLANAR)
    {
      if (it == itA || it == itB)
      {
        const uint32_t tot = __vadd2(__vadd2(S0, S1), __vadd2(S2, S3));
        if (it == itA) snapA = tot;
        if (it == itB) snapB = tot;
      }
      S0 = __vadd2(S0, vtest_planar<0>(cur.x, cur.z, negKlo2, n2, cap01));
      S1 = __vadd2(S1, vtest_planar<1>(cur.x, cur.z, negKlo2, n2, cap01));
      S2 = __vadd2(S2, vtest_planar<0>(cur.y, cur.w, negKlo2, n2, cap2));
      S3 = __vadd2(S3, vtest_planar<1>(cur.y, cur.w, negKlo2, n2, cap3));
    }
    else
    {
      S0 = __vadd2(S0, vtest_yuyv(cur.x, negKlo2, n2, np1));
      S1 = __vadd2(S1, vtest_yuyv(cur.y, negKlo2, n2, np1));
      S2 = __vadd2(S2, vtest_yuyv(cur.z, negKlo2, n2, np1));
      S3 = __vadd2(S3, vtest_yuyv(cur.w, negKlo2, n2, np1));
    }
    if (STAGES == 0)
      cur = nxt;
  }

  // every lane holds (iterations * N + fails) mod 2^16; fails <= iterations <= 128 per lane
  const uint32_t total = __vadd2(__vadd2(S0, S1), __vadd2(S2, S3));
  if (PLANAR)
  {
    if (itA >= iters) snapA = total;
    if (itB >= iters) snapB = total;
  }
  const uint32_t passBias2 = (((uint32_t)iters * nLane) & 0xFFFFu) * 0x10001u;
  const uint32_t f0 = lanes_sub(S0, passBias2), f1 = lanes_sub(S1, passBias2);
  const uint32_t f2 = lanes_sub(S2, passBias2), f3 = lanes_sub(S3, passBias2);
  uint32_t fails = lanes_total(f0) + lanes_total(f1) + lanes_total(f2) + lanes_total(f3);
  // in-chunk pixel index of pair k, lane e is 2k + e
  const uint32_t inIdx = 2u * (lanes_total(f1) + 2u * lanes_total(f2) + 3u * lanes_total(f3))
                       + (f0 >> 16) + (f1 >> 16) + (f2 >> 16) + (f3 >> 16);
  uint32_t sxFail = fails * ((uint32_t)cc * 8u) + inIdx;
  uint32_t crossFail = 0u;
  if (PLANAR)
  {
    const uint32_t bandBias2 = (((uint32_t)(itB - itA) * 4u * nLane) & 0xFFFFu) * 0x10001u;
    crossFail = lanes_total(lanes_sub(lanes_sub(snapB, snapA), bandBias2));
  }

  __syncthreads();
  const unsigned am = __activemask();
  fails  = __reduce_add_sync(am, fails);
  sxFail = __reduce_add_sync(am, sxFail);
  if (PLANAR) crossFail = __reduce_add_sync(am, crossFail);
  const int warp = t >> 5, lane = t & 31, nwarps = (blockDim.x + 31) >> 5;
  if (lane == 0)
  {
    s_red[warp][0] = fails; s_red[warp][1] = sxFail; s_red[warp][3] = crossFail;
  }
  __syncthreads();
  if (warp == 0)
  {
    uint32_t a = 0, b = 0, d = 0;
    if (lane < nwarps) { a = s_red[lane][0]; b = s_red[lane][1]; d = s_red[lane][3]; }
    const unsigned fm = __activemask();
    a = __reduce_add_sync(fm, a);
    b = __reduce_add_sync(fm, b);
    if (PLANAR) d = __reduce_add_sync(fm, d);
    if (lane == 0)
    {
      SumAcc* fa = acc + frame;
      bool last = true;
      if (slabs > 1)
      {
        atomicAdd(&fa->fails, a);
        atomicAdd(&fa->sxFail, b);
        if (PLANAR) atomicAdd(&fa->crossFail, d);
        __threadfence();
        last = (atomicAdd(&fa->done, 1u) == (uint32_t)slabs - 1u);
        if (last)
        {
          __threadfence();
          a = atomicExch(&fa->fails, 0u);
          b = atomicExch(&fa->sxFail, 0u);
          d = atomicExch(&fa->crossFail, 0u);
          atomicExch(&fa->done, 0u);
        }
      }
      if (last)
      {
        if (PLANAR) finalize_sum<KIND_OL>(g, p, a, b, 0u, d, out + frame, out);
        else        finalize_sum<KIND_WL>(g, p, a, b, 0u, 0u, out + frame, out);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// YUV422P line sensor (OL), 16 pixels per thread and row: one 16-byte luma chunk and the 16 chroma
// bytes under it, i.e. two full-width cp.async per ring slot and eight pixel pairs of vtest_lanes work
// per iteration.  Same bookkeeping as vsum_kernel (per-position {N, N+1} lanes, window folded into
// the caps, cross band as a difference of snapshots).
// ---------------------------------------------------------------------------------------------
template <int STAGES, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
vsum16_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
              const int paramStride, SumAcc* __restrict__ acc, TargetOut* __restrict__ out,
              const int slabs, const int rowsPerSlab, const int cpr, const int rpiAll,
              const int framesPerCta, const int numFrames)
{
  static_assert(STAGES >= 2, "ring depth");
  __shared__ uint32_t s_red[32][4];
  __shared__ uint32_t s_sums[8][4];            // framesPerCta > 1: per-frame totals by shared atomics
  extern __shared__ uint4 s_ring[];            // [STAGES][2][blockDim]: luma chunk, chroma chunk

  asm volatile("griddepcontrol.launch_dependents;");
  // framesPerCta == 1: a CTA is one slab of one frame.  framesPerCta == F > 1 (then slabs == 1): the rows of one
  // iteration are dealt out to F frames, so a thread walks its frame with a stride of rpiAll / F rows and gets
  // F x the iterations to spread its prologue and epilogue over (what small frames lack).
  const int t  = threadIdx.x;
  const int cc = t % cpr;
  const int rrAll = t / cpr;
  const int rpi = rpiAll / framesPerCta;
  const int sub = rrAll / rpi;                 // which of this CTA's frames
  const int rr = rrAll - sub * rpi;
  const int frame = framesPerCta > 1 ? (int)blockIdx.x * framesPerCta + sub : (int)blockIdx.x / slabs;
  const int slab  = framesPerCta > 1 ? 0 : (int)blockIdx.x - frame * slabs;
  const bool live = frame < numFrames;
  if (framesPerCta > 1 && t < 32)
    s_sums[t >> 2][t & 3] = 0u;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const FrameParams p = params[live ? (size_t)frame * paramStride : 0];

  const int r0 = slab * rowsPerSlab;
  const int r1 = min(r0 + rowsPerSlab, g.height);
  const int firstRow = r0 + rr;
  const int iters = (live && firstRow < r1) ? (r1 - firstRow + rpi - 1) / rpi : 0;
  const uint8_t* fillPtr = frames + (size_t)(live ? frame : 0) * g.frameStride + (size_t)cc * 16u + (size_t)firstRow * g.lineLength;
  const size_t rowStep = (size_t)rpi * g.lineLength;
  const size_t chromaOfs = (size_t)g.height * g.lineLength;

  const uint32_t negKlo2 = p.negKlo2, n2 = p.n2;
  const uint32_t nLane = n2 & 0xFFFFu;
  const uint32_t np1 = n2 + 0x00010001u;
  // window columns 5..W-5: chunk 0 loses its pixels 0..4 (pairs 0, 1 and the even pixel of pair 2),
  // the last chunk its pixels 12..15 (pairs 6, 7)
  uint32_t cap01 = np1, cap2 = np1, cap67 = np1;
  if (cc == 0)       { cap01 = n2; cap2 = (np1 & 0xFFFF0000u) | nLane; }
  if (cc == cpr - 1) { cap67 = n2; }

  int itA = 0, itB = 0;
  if (p.hStart <= p.hStop)
  {
    // rows past the image do not exist: clamp the band to it first, then 32-bit arithmetic is enough
    const int a = (int)min(p.hStart, (uint32_t)g.height) - firstRow;
    const int b = (int)min(p.hStop, (uint32_t)g.height - 1u) + 1 - firstRow;
    itA = a <= 0 ? 0 : min(iters, (a + rpi - 1) / rpi);
    itB = b <= 0 ? 0 : min(iters, (b + rpi - 1) / rpi);
  }

  uint32_t S0 = 0u, S1 = 0u, S2 = 0u, S3 = 0u, S4 = 0u, S5 = 0u, S6 = 0u, S7 = 0u;
  uint32_t snapA = 0u, snapB = 0u;
  // explicit 32-bit shared addresses: slot s of this thread is slotBase + s * stageBytes (luma), + planeBytes (chroma)
  const uint32_t planeBytes = blockDim.x * 16u, stageBytes = 2u * planeBytes;
  const uint32_t slotBase = (uint32_t)__cvta_generic_to_shared(s_ring) + (uint32_t)t * 16u;
  auto fill = [&](uint32_t slot)
  {
    const uint32_t dst = slotBase + slot * stageBytes;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(fillPtr) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + planeBytes), "l"(fillPtr + chromaOfs) : "memory");
  };

  int fillIt = 0;
#pragma unroll
  for (int sIdx = 0; sIdx < STAGES - 1; ++sIdx)
  {
    if (fillIt < iters)
      fill((uint32_t)sIdx);
    cp_async_commit();
    ++fillIt;
    fillPtr += rowStep;
  }
  // consume the chunk in ring slot `slot` (== iteration % STAGES) and request the one STAGES-1 iterations ahead
  auto body = [&](const uint32_t slot)
  {
    if (fillIt < iters)
      fill((slot + STAGES - 1u) % STAGES);
    cp_async_commit();
    ++fillIt;
    fillPtr += rowStep;
    cp_async_wait<STAGES - 1>();
    const uint32_t src = slotBase + slot * stageBytes;
    uint4 L, Cw;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(L.x), "=r"(L.y), "=r"(L.z), "=r"(L.w) : "r"(src));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(Cw.x), "=r"(Cw.y), "=r"(Cw.z), "=r"(Cw.w) : "r"(src + planeBytes));
    S0 = __vadd2(S0, vtest_planar<0>(L.x, Cw.x, negKlo2, n2, cap01));
    S1 = __vadd2(S1, vtest_planar<1>(L.x, Cw.x, negKlo2, n2, cap01));
    S2 = __vadd2(S2, vtest_planar<0>(L.y, Cw.y, negKlo2, n2, cap2));
    S3 = __vadd2(S3, vtest_planar<1>(L.y, Cw.y, negKlo2, n2, np1));
    S4 = __vadd2(S4, vtest_planar<0>(L.z, Cw.z, negKlo2, n2, np1));
    S5 = __vadd2(S5, vtest_planar<1>(L.z, Cw.z, negKlo2, n2, np1));
    S6 = __vadd2(S6, vtest_planar<0>(L.w, Cw.w, negKlo2, n2, cap67));
    S7 = __vadd2(S7, vtest_planar<1>(L.w, Cw.w, negKlo2, n2, cap67));
  };
  auto lanes_sum = [&]() -> uint32_t
  {
    return __vadd2(__vadd2(__vadd2(S0, S1), __vadd2(S2, S3)), __vadd2(__vadd2(S4, S5), __vadd2(S6, S7)));
  };
  // three runs of iterations -- before, inside and after the cross band -- with a snapshot of the
  // accumulators between them, so the band costs nothing per iteration (three copies of the loop measured
  // faster than one copy driven by an outer loop, and than grouping iterations by ring depth)
  auto run = [&](const int from, const int to)
  {
#pragma unroll 4
    for (int it = from; it < to; ++it)
      body((uint32_t)it % STAGES);
  };
  run(0, itA);
  snapA = lanes_sum();
  run(itA, itB);
  snapB = lanes_sum();
  run(itB, iters);

  const uint32_t total = __vadd2(__vadd2(__vadd2(S0, S1), __vadd2(S2, S3)), __vadd2(__vadd2(S4, S5), __vadd2(S6, S7)));
  (void)total;
  const uint32_t passBias2 = (((uint32_t)iters * nLane) & 0xFFFFu) * 0x10001u;
  const uint32_t f[8] = {lanes_sub(S0, passBias2), lanes_sub(S1, passBias2), lanes_sub(S2, passBias2), lanes_sub(S3, passBias2),
                         lanes_sub(S4, passBias2), lanes_sub(S5, passBias2), lanes_sub(S6, passBias2), lanes_sub(S7, passBias2)};
  uint32_t fails = 0u, inIdx = 0u;
#pragma unroll
  for (int k = 0; k < 8; ++k)
  {
    const uint32_t tk = lanes_total(f[k]);
    fails += tk;
    inIdx += 2u * (uint32_t)k * tk + (f[k] >> 16);       // in-chunk pixel index of pair k, lane e: 2k + e
  }
  uint32_t sxFail = fails * ((uint32_t)cc * 16u) + inIdx;
  const uint32_t bandBias2 = (((uint32_t)(itB - itA) * 8u * nLane) & 0xFFFFu) * 0x10001u;
  uint32_t crossFail = lanes_total(lanes_sub(lanes_sub(snapB, snapA), bandBias2));

  if (framesPerCta > 1)
  {
    __syncthreads();                                     // s_sums zeroed
    if (live)
    {
      atomicAdd(&s_sums[sub][0], fails);
      atomicAdd(&s_sums[sub][1], sxFail);
      atomicAdd(&s_sums[sub][3], crossFail);
    }
    __syncthreads();
    if (t < framesPerCta && (int)blockIdx.x * framesPerCta + t < numFrames)
    {
      const int f = (int)blockIdx.x * framesPerCta + t;
      const FrameParams pf = params[(size_t)f * paramStride];
      finalize_sum<KIND_OL>(g, pf, s_sums[t][0], s_sums[t][1], 0u, s_sums[t][3], out + f, out);
    }
    return;
  }
  __syncthreads();
  const unsigned am = __activemask();
  fails  = __reduce_add_sync(am, fails);
  sxFail = __reduce_add_sync(am, sxFail);
  crossFail = __reduce_add_sync(am, crossFail);
  const int warp = t >> 5, lane = t & 31, nwarps = (blockDim.x + 31) >> 5;
  if (lane == 0)
  {
    s_red[warp][0] = fails; s_red[warp][1] = sxFail; s_red[warp][3] = crossFail;
  }
  __syncthreads();
  if (warp == 0)
  {
    uint32_t a = 0, b = 0, d = 0;
    if (lane < nwarps) { a = s_red[lane][0]; b = s_red[lane][1]; d = s_red[lane][3]; }
    const unsigned fm = __activemask();
    a = __reduce_add_sync(fm, a);
    b = __reduce_add_sync(fm, b);
    d = __reduce_add_sync(fm, d);
    if (lane == 0)
    {
      SumAcc* fa = acc + frame;
      bool last = true;
      if (slabs > 1)
      {
        atomicAdd(&fa->fails, a);
        atomicAdd(&fa->sxFail, b);
        atomicAdd(&fa->crossFail, d);
        __threadfence();
        last = (atomicAdd(&fa->done, 1u) == (uint32_t)slabs - 1u);
        if (last)
        {
          __threadfence();
          a = atomicExch(&fa->fails, 0u);
          b = atomicExch(&fa->sxFail, 0u);
          d = atomicExch(&fa->crossFail, 0u);
          atomicExch(&fa->done, 0u);
        }
      }
      if (last)
        finalize_sum<KIND_OL>(g, p, a, b, 0u, d, out + frame, out);
    }
  }
}

static int sum_chunk_pixels(int kind) { return (kind == KIND_OL && (g_legacyLineKernel || g_widePlanarKernel || g_bulkLineKernel)) ? 16 : 8; }

int sum_sensor_block_threads(int kind, int width)
{
  const int cpr = width / sum_chunk_pixels(kind);
  if (cpr <= 0 || cpr > 1024)
    return 0;
  // measured on B200 (profiles/r01e_tuning_cta_size.jsonl): CTAs of ~160 threads beat 320 at 160x120 and 320x240
  const int target = g_targetThreads > 0 ? g_targetThreads : 128;
  int k = (target + cpr - 1) / cpr;
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
  resolve_tuning(kind, g.width);
  const int threads = sum_sensor_block_threads(kind, g.width);
  if (threads <= 0)
    return cudaErrorInvalidValue;
  const int cpr = g.width / sum_chunk_pixels(kind);
  const int rpi = threads / cpr;
  // SI lanes (WO) hold sum(it * fails_per_chunk) <= 8 * I*(I-1)/2: keep I <= 128 iterations per slab
  const int maxRowsPerSlab = 128 * rpi;
  int slabs = slabsPerFrame;
  if (slabs <= 0)
  {
    // measured (profiles/r01k_tuning_slabs_small_batches.jsonl): one CTA per frame wins as soon as the frames
    // alone fill the machine (~1480 resident CTAs); the wide YUV422P kernel has the costlier prologue and
    // epilogue and wants >= 24 iterations per thread, the others >= 4
    const int wantCtas = 148 * 12;
    slabs = (wantCtas + numFrames - 1) / numFrames;
    const int minIters = g_widePlanarKernel ? 24 : 4;
    const int maxSlabs = g.height / (rpi * minIters) > 0 ? g.height / (rpi * minIters) : 1;
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
  const int stages = g_sumStages;
  const size_t ringBytes = (size_t)stages * threads * sizeof(uint4) * ((kind == KIND_OL && (g_legacyLineKernel || g_widePlanarKernel)) ? 2 : 1);
#define TRIK_LAUNCH_SUM(K, ST)                                                                                   \
  do {                                                                                                           \
    if (ringBytes > 48 * 1024)                                                                                   \
      cudaFuncSetAttribute(sum_kernel<K, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ringBytes);      \
    sum_kernel<K, ST><<<(unsigned)grid, threads, ringBytes, stream>>>(g, frames, params, paramStride, acc, out,  \
                                                                      slabs, rowsPerSlab, cpr, rpi);             \
  } while (0)
#define TRIK_LAUNCH_SUM_KIND(K)                                 \
  switch (stages)                                               \
  {                                                             \
    case 0: TRIK_LAUNCH_SUM(K, 0); break;                       \
    case 2: TRIK_LAUNCH_SUM(K, 2); break;                       \
    case 4: TRIK_LAUNCH_SUM(K, 4); break;                       \
    default: return cudaErrorInvalidValue;                      \
  }
#define TRIK_LAUNCH_V2(PL, ST, MT)                                                                               \
  do {                                                                                                           \
    if (ringBytes > 48 * 1024)                                                                                   \
      cudaFuncSetAttribute(vsum_kernel<PL, ST, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ringBytes);\
    cudaLaunchConfig_t cfg = {};                                                                                 \
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads);                                  \
    cfg.dynamicSmemBytes = ringBytes; cfg.stream = stream;                                                       \
    cudaLaunchAttribute attr[1];                                                                                 \
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                             \
    attr[0].val.programmaticStreamSerializationAllowed = 1;                                                      \
    cfg.attrs = attr; cfg.numAttrs = g_overlapLaunch ? 1u : 0u;                                                  \
    cudaLaunchKernelEx(&cfg, vsum_kernel<PL, ST, MT>, g, frames, params, paramStride, acc, out, slabs,           \
                       rowsPerSlab, cpr, rpi);                                                                   \
  } while (0)
#define TRIK_LAUNCH_V(PL, ST)                                   \
  do { if (threads <= 512) TRIK_LAUNCH_V2(PL, ST, 512); else TRIK_LAUNCH_V2(PL, ST, 1024); } while (0)
#define TRIK_LAUNCH_V_KIND(PL)                                  \
  switch (stages)                                               \
  {                                                             \
    case 0: TRIK_LAUNCH_V(PL, 0); break;                        \
    case 2: TRIK_LAUNCH_V(PL, 2); break;                        \
    case 4: TRIK_LAUNCH_V(PL, 4); break;                        \
    case 6: TRIK_LAUNCH_V(PL, 6); break;                        \
    case 8: TRIK_LAUNCH_V(PL, 8); break;                        \
    default: return cudaErrorInvalidValue;                      \
  }
  if (g_bulkLineKernel)
  {
    const cudaError_t e = launch_line_bulk(kind == KIND_OL, g, grid, threads, frames, params, paramStride, acc, out, slabs,
                                           rowsPerSlab, cpr, rpi, stages, g_overlapLaunch != 0, stream);
    if (e != cudaSuccess)
      return e;
    return cudaGetLastError();
  }
  switch (kind)
  {
    case KIND_WL:
      if (g_legacyLineKernel) { TRIK_LAUNCH_SUM_KIND(KIND_WL); } else { TRIK_LAUNCH_V_KIND(false); }
      break;
    case KIND_OL:
      if (g_widePlanarKernel)
      {
        // small frames: two frames per CTA, so that a thread has ~60 iterations to spread its prologue / epilogue over
        int fpc = 1;
        if (slabs == 1 && g_framesPerCta != 1)
        {
          const int itersOne = (g.height + rpi - 1) / rpi;
          const int want = g_framesPerCta > 1 ? g_framesPerCta : (itersOne >= 50 ? 1 : 2);     // measured: 2 gives +2 % at 320x240, 4 loses 6 %
          for (fpc = want > 8 ? 8 : want; fpc > 1 && (rpi % fpc != 0 || g.height % (rpi / fpc) != 0); --fpc) {}
        }
        const long long gridW = fpc > 1 ? (numFrames + fpc - 1) / fpc : grid;
#define TRIK_LAUNCH_W2(ST, MT, MB)                                                                                       \
  do {                                                                                                           \
    if (ringBytes > 48 * 1024)                                                                                   \
      cudaFuncSetAttribute(vsum16_kernel<ST, MT, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ringBytes);  \
    cudaLaunchConfig_t cfg = {};                                                                                 \
    cfg.gridDim = dim3((unsigned)gridW); cfg.blockDim = dim3((unsigned)threads);                                 \
    cfg.dynamicSmemBytes = ringBytes; cfg.stream = stream;                                                       \
    cudaLaunchAttribute attr[1];                                                                                 \
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                             \
    attr[0].val.programmaticStreamSerializationAllowed = 1;                                                      \
    cfg.attrs = attr; cfg.numAttrs = g_overlapLaunch ? 1u : 0u;                                                  \
    cudaLaunchKernelEx(&cfg, vsum16_kernel<ST, MT, MB>, g, frames, params, paramStride, acc, out, slabs, rowsPerSlab,\
                       cpr, rpi, fpc, numFrames);                                                                \
  } while (0)
#define TRIK_LAUNCH_W(ST, MB) do { if (threads <= 256) TRIK_LAUNCH_W2(ST, 256, MB); else TRIK_LAUNCH_W2(ST, 1024, 1); } while (0)
        switch (stages)
        {
          case 2: TRIK_LAUNCH_W(2, 4); break;
          case 4: TRIK_LAUNCH_W(4, 4); break;
          default: return cudaErrorInvalidValue;
        }
#undef TRIK_LAUNCH_W
#undef TRIK_LAUNCH_W2
      }
      else if (g_legacyLineKernel) { TRIK_LAUNCH_SUM_KIND(KIND_OL); } else { TRIK_LAUNCH_V_KIND(true); }
      break;
    case KIND_WO: TRIK_LAUNCH_SUM_KIND(KIND_WO); break;
    default:
      return cudaErrorInvalidValue;
  }
#undef TRIK_LAUNCH_V_KIND
#undef TRIK_LAUNCH_V
#undef TRIK_LAUNCH_V2
#undef TRIK_LAUNCH_SUM_KIND
#undef TRIK_LAUNCH_SUM
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

// packed pair path: index = Y | U << 8 | V << 16 -> 0x00VVSSHH through hsv_pair(); the probe pixel sits in
// lane 0 of a YUYV word and in lane 1 of a YUV422P pair, both must agree (bit 31 flags a disagreement)
__global__ void probe_yuv2hsv_kernel(uint32_t first, uint32_t count, uint32_t* __restrict__ out)
{
  __shared__ HueLutEntry s_lutHue[256];
  __shared__ uint16_t s_lut43[256];
  __shared__ uint16_t s_lut255[256];
  fill_div_luts(s_lut43, s_lut255);
  fill_hue_lut(s_lutHue);
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t idx = first + i;
  const uint32_t y = idx & 0xFFu, u = (idx >> 8) & 0xFFu, v = (idx >> 16) & 0xFFu;
  const uint32_t word = y | (u << 8) | ((255u - y) << 16) | (v << 24);
  uint32_t a0, a1, b0, b1;
  hsv_pair(word & 0x00FF00FFu, word, coef_yuyv(), s_lutHue, s_lut255, a0, a1);
  const uint32_t cw = v | (u << 8) | (v << 16) | (u << 24);
  hsv_pair(((255u - y) & 0xFFu) | (y << 16), cw, coef_planar1(), s_lutHue, s_lut255, b0, b1);
  out[i] = a0 | ((a0 != b1 || a1 != b0) ? 0x80000000u : 0u);
}

cudaError_t launch_probe_yuv2hsv(uint32_t first, uint32_t count, uint32_t* out, cudaStream_t stream)
{
  if (!count) return cudaSuccess;
  probe_yuv2hsv_kernel<<<(count + 255u) / 256u, 256, 0, stream>>>(first, count, out);
  ++g_launches;
  return cudaGetLastError();
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
