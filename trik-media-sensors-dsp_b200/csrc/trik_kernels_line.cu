// trik_kernels_line.cu -- line sensors (WL, OL) with bulk-copy staging (sm_100a).
//
// Same arithmetic and decomposition as vsum_kernel (trik_kernels.cu): a CTA owns a slab of rows of one
// frame, a thread owns one 16-byte column chunk and walks down the rows.  What changes is how the bytes
// arrive.  ncu showed both line kernels bound by instruction issue, and a third of the issued
// instructions were per-thread copy bookkeeping (address arithmetic, LDGSTS and its dummy LDS, commit /
// wait groups, ring indices).  Here the copy engine does that work: per iteration ONE warp issues
// `cp.async.bulk` row copies (one per plane when the rows are contiguous) into a ring of shared-memory
// stages, completion is counted in bytes on an mbarrier per stage, and every thread's share of the
// staging is one try_wait and one 16-byte LDS per plane.  Stages are handed back through a second
// mbarrier per stage on which each warp arrives once it holds its chunk in registers.
//
//   WL  webcam/line_sensor/include/internal/cv_line_detector_seqpass.hpp:197-269, tail :401-417
//   OL  ov7670/line_sensor/include/internal/cv_line_detector_seqpass.hpp:210-301, tail :449-473
#include <atomic>
#include "trik_kernels.cuh"
#include "trik_pixel.cuh"
#include "trik_line.cuh"

namespace trikb200 {

std::atomic<long long> g_launches_line{0};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  uint32_t done;
  do
  {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// PLANAR = false: YUYV, a chunk is 8 pixels (16 bytes of one plane), four pair accumulators.
// PLANAR = true : YUV422P, a chunk is 16 pixels (16 luma bytes + the 16 chroma bytes under them),
//                 eight pair accumulators, column window and cross band as in vsum_kernel.
template <bool PLANAR, int STAGES>
__global__ void __launch_bounds__(512, 2)
tsum_kernel(const Geometry g, const uint8_t* __restrict__ frames, const FrameParams* __restrict__ params,
            const int paramStride, SumAcc* __restrict__ acc, TargetOut* __restrict__ out,
            const int slabs, const int rowsPerSlab, const int cpr, const int rpi)
{
  constexpr int PLANES = PLANAR ? 2 : 1;
  constexpr int NACC = PLANAR ? 8 : 4;
  __shared__ uint32_t s_red[32][4];
  __shared__ __align__(8) uint64_t s_full[STAGES];
  __shared__ __align__(8) uint64_t s_empty[STAGES];
  extern __shared__ __align__(128) uint8_t s_stage[];   // [STAGES][PLANES][blockDim * 16]

  asm volatile("griddepcontrol.launch_dependents;");
  const int frame = blockIdx.x / slabs;
  const int slab  = blockIdx.x - frame * slabs;
  const int t  = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int nthreads = (int)blockDim.x;
  const int nwarps = (nthreads + 31) >> 5;
  const int cc = t % cpr;
  const int rr = t / cpr;

  const int r0 = slab * rowsPerSlab;
  const int r1 = min(r0 + rowsPerSlab, g.height);
  const int nIt = (r1 - r0 + rpi - 1) / rpi;                 // CTA-wide iterations (>= 1)
  const int firstRow = r0 + rr;
  const int iters = firstRow < r1 ? (r1 - firstRow + rpi - 1) / rpi : 0;   // this thread's share (nIt or nIt - 1)

  const uint32_t rowBytes = (uint32_t)(PLANAR ? g.width : g.width * 2);
  const uint32_t planeBytes = (uint32_t)nthreads * 16u;      // == rpi * rowBytes
  const uint32_t stageBytes = planeBytes * PLANES;
  const uint32_t stage0 = smem_u32(s_stage);
  const uint32_t full0 = smem_u32(s_full), empty0 = smem_u32(s_empty);
  const bool contiguous = (uint32_t)g.lineLength == rowBytes;
  const uint8_t* const frameBase = frames + (size_t)frame * g.frameStride;
  const size_t chromaOfs = (size_t)g.height * g.lineLength;

  if (t == 0)
  {
#pragma unroll
    for (int s = 0; s < STAGES; ++s)
    {
      mbar_init(full0 + 8u * s, 1u);
      mbar_init(empty0 + 8u * s, (uint32_t)nwarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // warp 0: request CTA iteration `j` (rows r0 + j*rpi ...) into stage `s`
  auto request = [&](int j, int s)
  {
    const int rowBase = r0 + j * rpi;
    const int valid = min(rpi, r1 - rowBase);
    const uint32_t bar = full0 + 8u * s;
    const uint32_t dst0 = stage0 + (uint32_t)s * stageBytes;
    if (lane == 0)
      mbar_arrive_expect_tx(bar, (uint32_t)valid * rowBytes * PLANES);
    __syncwarp();
    if (contiguous)
    {
      if (lane < PLANES)
        bulk_copy_g2s(dst0 + (uint32_t)lane * planeBytes,
                      frameBase + (size_t)lane * chromaOfs + (size_t)rowBase * g.lineLength,
                      (uint32_t)valid * rowBytes, bar);
    }
    else
    {
      for (int c = lane; c < valid * PLANES; c += 32)
      {
        const int plane = PLANES == 2 ? (c >= valid) : 0;
        const int row = c - plane * valid;
        bulk_copy_g2s(dst0 + (uint32_t)plane * planeBytes + (uint32_t)row * rowBytes,
                      frameBase + (size_t)plane * chromaOfs + (size_t)(rowBase + row) * g.lineLength, rowBytes, bar);
      }
    }
  };
  if (warp == 0)
  {
#pragma unroll 1
    for (int j = 0; j < STAGES && j < nIt; ++j)
      request(j, j);
  }

  const FrameParams p = params[(size_t)frame * paramStride];
  const uint32_t negKlo2 = p.negKlo2, n2 = p.n2;
  const uint32_t nLane = n2 & 0xFFFFu;
  const uint32_t np1 = n2 + 0x00010001u;               // N + 1 in both lanes (N <= 65534 by construction)
  // OL counts columns 5..W-5: chunk 0 loses its pixels 0..4 (pairs 0, 1 and the even pixel of pair 2),
  // the last chunk its pixels 12..15 (pairs 6, 7); an excluded position is capped to N ("always passes")
  uint32_t cap01 = np1, cap2 = np1, cap67 = np1;
  if (PLANAR)
  {
    if (cc == 0)       { cap01 = n2; cap2 = (np1 & 0xFFFF0000u) | nLane; }
    if (cc == cpr - 1) { cap67 = n2; }
  }
  // OL cross band as a run [itA, itB) of this thread's iterations
  int itA = 0, itB = 0;
  if (PLANAR && p.hStart <= p.hStop)
  {
    // rows past the image do not exist: clamp the band to it first, then 32-bit arithmetic is enough
    const int a = (int)min(p.hStart, (uint32_t)g.height) - firstRow;
    const int b = (int)min(p.hStop, (uint32_t)g.height - 1u) + 1 - firstRow;
    itA = a <= 0 ? 0 : min(iters, (a + rpi - 1) / rpi);
    itB = b <= 0 ? 0 : min(iters, (b + rpi - 1) / rpi);
  }

  uint32_t S[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) S[k] = 0u;
  uint32_t snapA = 0u, snapB = 0u;
  auto lanes_sum = [&]() -> uint32_t
  {
    uint32_t tot = S[0];
#pragma unroll
    for (int k = 1; k < NACC; ++k) tot = __vadd2(tot, S[k]);
    return tot;
  };

  uint32_t stage = 0u, phase = 0u;                       // ring position of CTA iteration `it`
  uint32_t slotAddr = stage0 + (uint32_t)t * 16u;
#pragma unroll 1
  for (int it = 0; it < nIt; ++it)
  {
    mbar_wait(full0 + 8u * stage, phase);
    uint4 L, Cw;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(L.x), "=r"(L.y), "=r"(L.z), "=r"(L.w) : "r"(slotAddr));
    if (PLANAR)
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(Cw.x), "=r"(Cw.y), "=r"(Cw.z), "=r"(Cw.w) : "r"(slotAddr + planeBytes));
    __syncwarp();
    if (lane == 0)
      mbar_arrive(empty0 + 8u * stage);
    // warp 0 refills the stage of the PREVIOUS iteration (its readers are done, or nearly)
    if (warp == 0 && it >= 1 && it - 1 + STAGES < nIt)
    {
      const uint32_t ps = stage == 0u ? (uint32_t)STAGES - 1u : stage - 1u;
      const uint32_t pp = stage == 0u ? phase ^ 1u : phase;
      mbar_wait(empty0 + 8u * ps, pp);
      request(it - 1 + STAGES, (int)ps);
    }
    if (it < iters)
    {
      if (PLANAR)
      {
        if (__any_sync(__activemask(), it == itA || it == itB))     // rare: keep the snapshots off the common path
        {
          if (it == itA) snapA = lanes_sum();
          if (it == itB) snapB = lanes_sum();
        }
        S[0] = __vadd2(S[0], vtest_planar<0>(L.x, Cw.x, negKlo2, n2, cap01));
        S[1] = __vadd2(S[1], vtest_planar<1>(L.x, Cw.x, negKlo2, n2, cap01));
        S[2] = __vadd2(S[2], vtest_planar<0>(L.y, Cw.y, negKlo2, n2, cap2));
        S[3] = __vadd2(S[3], vtest_planar<1>(L.y, Cw.y, negKlo2, n2, np1));
        S[4 % NACC] = __vadd2(S[4 % NACC], vtest_planar<0>(L.z, Cw.z, negKlo2, n2, np1));
        S[5 % NACC] = __vadd2(S[5 % NACC], vtest_planar<1>(L.z, Cw.z, negKlo2, n2, np1));
        S[6 % NACC] = __vadd2(S[6 % NACC], vtest_planar<0>(L.w, Cw.w, negKlo2, n2, cap67));
        S[7 % NACC] = __vadd2(S[7 % NACC], vtest_planar<1>(L.w, Cw.w, negKlo2, n2, cap67));
      }
      else
      {
        S[0] = __vadd2(S[0], vtest_yuyv(L.x, negKlo2, n2, np1));
        S[1] = __vadd2(S[1], vtest_yuyv(L.y, negKlo2, n2, np1));
        S[2] = __vadd2(S[2], vtest_yuyv(L.z, negKlo2, n2, np1));
        S[3] = __vadd2(S[3], vtest_yuyv(L.w, negKlo2, n2, np1));
      }
    }
    slotAddr += stageBytes;
    if (++stage == (uint32_t)STAGES)
    {
      stage = 0u; phase ^= 1u;
      slotAddr = stage0 + (uint32_t)t * 16u;
    }
  }

  // every lane holds (iterations * N + fails) mod 2^16; fails <= iterations <= 128 per lane
  const uint32_t total = lanes_sum();
  if (PLANAR)
  {
    if (itA >= iters) snapA = total;
    if (itB >= iters) snapB = total;
  }
  const uint32_t passBias2 = (((uint32_t)iters * nLane) & 0xFFFFu) * 0x10001u;
  uint32_t fails = 0u, inIdx = 0u;
#pragma unroll
  for (int k = 0; k < NACC; ++k)
  {
    const uint32_t fk = lanes_sub(S[k], passBias2);
    const uint32_t tk = lanes_total(fk);
    fails += tk;
    inIdx += 2u * (uint32_t)k * tk + (fk >> 16);          // in-chunk pixel index of pair k, lane e: 2k + e
  }
  uint32_t sxFail = fails * ((uint32_t)cc * (PLANAR ? 16u : 8u)) + inIdx;
  uint32_t crossFail = 0u;
  if (PLANAR)
  {
    const uint32_t bandBias2 = (((uint32_t)(itB - itA) * (uint32_t)NACC * nLane) & 0xFFFFu) * 0x10001u;
    crossFail = lanes_total(lanes_sub(lanes_sub(snapB, snapA), bandBias2));
  }

  const unsigned am = __activemask();
  fails  = __reduce_add_sync(am, fails);
  sxFail = __reduce_add_sync(am, sxFail);
  if (PLANAR) crossFail = __reduce_add_sync(am, crossFail);
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

template <bool PLANAR, int STAGES>
static cudaError_t launch_tsum(const Geometry& g, long long grid, int threads, const uint8_t* frames,
                               const FrameParams* params, int paramStride, SumAcc* acc, TargetOut* out,
                               int slabs, int rowsPerSlab, int cpr, int rpi, bool overlap, cudaStream_t stream)
{
  const size_t stageBytes = (size_t)STAGES * (PLANAR ? 2 : 1) * threads * 16u;
  if (stageBytes > 48 * 1024)
  {
    const cudaError_t e = cudaFuncSetAttribute(tsum_kernel<PLANAR, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stageBytes);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = stageBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = overlap ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, tsum_kernel<PLANAR, STAGES>, g, frames, params, paramStride, acc, out, slabs,
                            rowsPerSlab, cpr, rpi);
}

// Called by launch_sum_sensor (trik_kernels.cu) once the decomposition is fixed.
cudaError_t launch_line_bulk(bool planar, const Geometry& g, long long grid, int threads, const uint8_t* frames,
                             const FrameParams* params, int paramStride, SumAcc* acc, TargetOut* out,
                             int slabs, int rowsPerSlab, int cpr, int rpi, int stages, bool overlap, cudaStream_t stream)
{
  if (threads > 512 || threads != cpr * rpi)
    return cudaErrorInvalidValue;
  cudaError_t e = cudaErrorInvalidValue;
#define TRIK_TSUM(PL, ST) e = launch_tsum<PL, ST>(g, grid, threads, frames, params, paramStride, acc, out, slabs, rowsPerSlab, cpr, rpi, overlap, stream)
  if (planar)
  {
    if (stages == 2) TRIK_TSUM(true, 2); else if (stages == 3) TRIK_TSUM(true, 3); else if (stages == 4) TRIK_TSUM(true, 4);
  }
  else
  {
    if (stages == 2) TRIK_TSUM(false, 2); else if (stages == 3) TRIK_TSUM(false, 3); else if (stages == 4) TRIK_TSUM(false, 4);
    else if (stages == 6) TRIK_TSUM(false, 6);
  }
#undef TRIK_TSUM
  if (e == cudaSuccess)
    ++g_launches_line;
  return e;
}

} // namespace trikb200
