// trik_capi.cu -- the C ABI of libtrikb200 (include/trik_b200.h): the reference's codec surface
// (alloc / initObj / free / process / control, <sensor>/src/vidtranscode_cv_fxns.c:85-334 and the
// handle/dispatch layer <sensor>/src/vidtranscode_cv.cpp:22-320) plus the batch extension.
//
// There is NO CPU fallback: every pixel of every frame goes through the CUDA kernels, and any
// CUDA failure is reported as IVIDTRANSCODE_EFAIL / IALG_EFAIL with trikb200_lastError() set.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <climits>
#include <mutex>
#include <new>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "trik_b200.h"
#include "trik_host.hpp"
#include "trik_kernels.cuh"

using namespace trikb200;

namespace {

thread_local std::string t_lastError;
thread_local int t_device = -1;
int g_slabsPerFrame = 0;
int g_zeroCopyBytes = 1 << 20;     // synchronous host batches up to this many frame bytes go through the handle's pinned staging
int g_lutMode = 0;                 // 0 auto, 1 whenever the arguments are shared by the batch, -1 never
int g_gatherMode = 0;              // scattered pinned host frames: 0 auto (gather kernel from 8 frames), 1 always, -1 never (one copy per frame)
int g_mxnTableMode = 0;            // mxn sensor through the colour-bin table: 0 / 1 yes (default), -1 never (arithmetic kernel)

// The colour-bin table of the mxn sensor (trik_kernels_omtab.cu) depends on nothing but the device: built once per
// device and process, on first use, and kept until the process ends.
constexpr int MAX_DEVICES = 64;
std::mutex g_omTableMutex;
uint16_t* g_omTable[MAX_DEVICES] = {};

void set_error(const char* what, cudaError_t e = cudaSuccess)
{
  t_lastError = what;
  if (e != cudaSuccess)
  {
    t_lastError += ": ";
    t_lastError += cudaGetErrorString(e);
  }
}

#define CUDA_TRY(expr)                                   \
  do {                                                   \
    cudaError_t e__ = (expr);                            \
    if (e__ != cudaSuccess) { set_error(#expr, e__); return false; } \
  } while (0)

const size_t kFastRamSize = 0x1000;          // include/internal/vidtranscode_cv.h:28
const char   kVersion[]   = "1.00.00.00";    // src/vidtranscode_cv_fxns.c:75

// The handle the caller allocates from the alloc() table: IALG_Obj first (vidtranscode_cv.h:17-27).
struct TrikB200Handle {
  IALG_Obj                           alg;
  XDAS_Int32                         kind;
  TRIK_VIDTRANSCODE_CV_Params        params;
  TRIK_VIDTRANSCODE_CV_DynamicParams dynamicParams;
  void*                              impl;       // Instance*
  XDAS_Int8*                         fastRam;
  size_t                             fastRamSize;
};

size_t in_args_alg_size(int kind)
{
  switch (kind)
  {
    case KIND_OO: return sizeof(TRIKB200_ObjInArgsAlg);
    case KIND_OM: return sizeof(TRIKB200_MxnInArgsAlg);
    default:      return sizeof(TRIKB200_RangeInArgsAlg);
  }
}
size_t out_args_alg_size(int kind)
{
  switch (kind)
  {
    case KIND_OO: return sizeof(TRIKB200_ObjOutArgsAlg);
    case KIND_OM: return sizeof(TRIKB200_MxnOutArgsAlg);
    default:      return sizeof(TRIKB200_TargetOutArgsAlg);
  }
}
size_t in_args_size(int kind)
{
  switch (kind)
  {
    case KIND_OO: return sizeof(TRIKB200_ObjInArgs);
    case KIND_OM: return sizeof(TRIKB200_MxnInArgs);
    default:      return sizeof(TRIKB200_RangeInArgs);
  }
}
size_t out_args_size(int kind)
{
  switch (kind)
  {
    case KIND_OO: return sizeof(TRIKB200_ObjOutArgs);
    case KIND_OM: return sizeof(TRIKB200_MxnOutArgs);
    default:      return sizeof(TRIKB200_TargetOutArgs);
  }
}
XDAS_Int32 default_input_format(int kind)
{
  return kind_is_planar(kind) ? TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV422P : TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV422;
}

// ---------------------------------------------------------------------------------------------
// Instance: the "CVAlgorithm" object of a handle.  Re-created by every SETPARAMS, as the
// reference re-creates its algorithm object (src/vidtranscode_cv.cpp:52-66).
// ---------------------------------------------------------------------------------------------
struct Instance {
  int          kind = 0;
  int          device = 0;
  bool         valid = false;        // setup() accepted the geometry
  Geometry     geo{};
  int          outWidth = 0, outHeight = 0, outLineLength = 0;
  CarriedState state;
  std::vector<CarriedState> streamStates;   // logical streams of the batch extension (TRIKB200_Batch.streamIds)
  int64_t      seed = -1;
  cudaStream_t stream = nullptr;

  // device workspace, grown on demand
  uint8_t*     dFrames = nullptr;   size_t dFramesCap = 0;
  FrameParams* dParams = nullptr;   size_t dParamsCap = 0;     // entries
  SumAcc*      dAcc = nullptr;      size_t dAccCap = 0;        // entries
  uint8_t*     dOut = nullptr;      size_t dOutCap = 0;        // bytes
  uint32_t*    dMxnTable = nullptr;                            // 512 colours of the mxn sensor
  uint16_t*    dBitmaps = nullptr;  size_t dBitmapsCap = 0;    // OO metapixel bitmaps (uint16 entries)
  uint8_t*     dClusters = nullptr; size_t dClustersCap = 0;   // OO cluster records (bytes)
  uint16_t*    dEqual = nullptr;    size_t dEqualCap = 0;      // OO label equivalences
  int*         dFlagged = nullptr;  size_t dFlaggedCap = 0;    // indices of frames that auto-calibrate
  int*         dOmList = nullptr;   size_t dOmListCap = 0;     // OM: cell rows the majority pass left to the histogram kernel
  int*         dOmCount = nullptr;  size_t dOmCountCap = 0;    // ... their number: two counters, used by alternate batches
  int          omParity = 0;
  int32_t*     dHist = nullptr;     size_t dHistCap = 0;       // ordered +1/-2 histograms (int32 entries)
  // chroma-indexed detection table of the last threshold set (WO batches, trik_kernels_lut.cu)
  uint8_t*     dLutTable = nullptr; uint32_t* dLutMasks = nullptr;
  uint32_t     lutFrom = 0, lutTo = 0, lutExpected = 0; bool lutValid = false; cudaStream_t lutStream = nullptr;
  int          smCount = 0;
  // further tables for batches whose frames come under several threshold sets (up to LUT_CACHE of them, least recently
  // used replaced), and the frame lists of such a batch
  struct LutSlot { uint8_t* table = nullptr; uint32_t* masks = nullptr; uint32_t from = 0, to = 0, expected = 0;
                   bool valid = false; cudaStream_t stream = nullptr; unsigned long long lastUse = 0; };
  static constexpr int LUT_CACHE = LUT_MAX_SETS;
  LutSlot      lutCache[LUT_CACHE];  unsigned long long lutClock = 0;
  int*         hLutList = nullptr;  size_t hLutListCap = 0;
  int*         dLutList = nullptr;  size_t dLutListCap = 0;
  cudaEvent_t  hLutListFree = nullptr;
  // preview (RGB565X) support: index maps of this geometry, overlay inputs, staging image
  int32_t*     dHi2ho = nullptr;  int32_t* dWi2wo = nullptr;   // source row/col -> preview row/col
  int32_t*     dLastRow = nullptr; int32_t* dLastCol = nullptr; // preview row/col -> last source row/col (-1: none)
  DrawInfo*    dDraw = nullptr;     size_t dDrawCap = 0;
  uint8_t*     dPreview = nullptr;  size_t dPreviewCap = 0;
  // pinned host staging
  FrameParams* hParams = nullptr;   size_t hParamsCap = 0;
  cudaEvent_t  hParamsFree = nullptr;                          // completes when the last upload has read hParams
  std::vector<FrameParams> paramsScratch;                      // a batch's parameters before they are staged
  std::vector<int> argSetScratch;                              // ... and, for the webcam sensors, which distinct argument set each frame uses
  // the single broadcast record dParams[0] currently holds, and the stream that uploaded it
  FrameParams  dBroadcast = {};     bool dBroadcastValid = false;  cudaStream_t dBroadcastStream = nullptr;
  uint8_t*     hOut = nullptr;      size_t hOutCap = 0;
  uint8_t*     hFrames = nullptr;   size_t hFramesCap = 0;     // small host batches: frames staged here, read in place by the kernels
  uint8_t*     hPreview = nullptr;  size_t hPreviewCap = 0;    // ... and previews written here by the kernels
  int*         hFlagged = nullptr;  size_t hFlaggedCap = 0;
  const uint8_t** hPtrs = nullptr;  size_t hPtrsCap = 0;       // scattered frames: their device-visible addresses, staged ...
  const uint8_t** dPtrs = nullptr;  size_t dPtrsCap = 0;       // ... and on the device for the gather kernel
  cudaEvent_t  hPtrsFree = nullptr;                            // completes when the last upload has read hPtrs
  // a threshold set that keeps coming back in medium-sized batches is worth its chroma table
  uint32_t     lutSeenFrom = 0, lutSeenTo = 0, lutSeenExpected = 0; int lutSeenCount = 0;
  uint32_t*    hSeeds = nullptr;    size_t hSeedsCap = 0;      // device annealing tail: one seed per calibrating frame
  uint32_t*    dSeeds = nullptr;    size_t dSeedsCap = 0;
  cudaEvent_t  hStageFree = nullptr;                           // completes when hFlagged / hSeeds have been read
  int32_t*     hHist = nullptr;     size_t hHistCap = 0;

  ~Instance() { release(); }

  void release()
  {
    if (stream || dFrames || dParams || dAcc || dOut || hParams || hOut || dMxnTable)
      cudaSetDevice(device);
    if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); stream = nullptr; }
    cudaFree(dFrames); dFrames = nullptr; dFramesCap = 0;
    cudaFree(dParams); dParams = nullptr; dParamsCap = 0;
    cudaFree(dLutTable); dLutTable = nullptr; cudaFree(dLutMasks); dLutMasks = nullptr; lutValid = false;
    for (LutSlot& sl : lutCache) { cudaFree(sl.table); cudaFree(sl.masks); sl = LutSlot(); }
    cudaFreeHost(hLutList); hLutList = nullptr; hLutListCap = 0;
    cudaFree(dLutList); dLutList = nullptr; dLutListCap = 0;
    if (hLutListFree) { cudaEventDestroy(hLutListFree); hLutListFree = nullptr; }
    cudaFree(dAcc);    dAcc = nullptr;    dAccCap = 0;
    cudaFree(dOut);    dOut = nullptr;    dOutCap = 0;
    cudaFree(dMxnTable); dMxnTable = nullptr;
    free_maps();
    cudaFree(dDraw);    dDraw = nullptr;    dDrawCap = 0;
    cudaFree(dPreview); dPreview = nullptr; dPreviewCap = 0;
    cudaFree(dBitmaps);  dBitmaps = nullptr;  dBitmapsCap = 0;
    cudaFree(dClusters); dClusters = nullptr; dClustersCap = 0;
    cudaFree(dEqual);    dEqual = nullptr;    dEqualCap = 0;
    cudaFree(dFlagged);  dFlagged = nullptr;  dFlaggedCap = 0;
    cudaFree(dOmList);   dOmList = nullptr;   dOmListCap = 0;
    cudaFree(dOmCount);  dOmCount = nullptr;  dOmCountCap = 0;
    cudaFree(dHist);     dHist = nullptr;     dHistCap = 0;
    cudaFreeHost(hParams);  hParams = nullptr;  hParamsCap = 0;
    if (hParamsFree) { cudaEventDestroy(hParamsFree); hParamsFree = nullptr; }
    dBroadcastValid = false;
    cudaFreeHost(hOut);     hOut = nullptr;     hOutCap = 0;
    cudaFreeHost(hFrames);  hFrames = nullptr;  hFramesCap = 0;
    cudaFreeHost(hPreview); hPreview = nullptr; hPreviewCap = 0;
    cudaFreeHost(hFlagged); hFlagged = nullptr; hFlaggedCap = 0;
    cudaFreeHost(hPtrs);    hPtrs = nullptr;    hPtrsCap = 0;
    cudaFree(dPtrs);        dPtrs = nullptr;    dPtrsCap = 0;
    if (hPtrsFree) { cudaEventDestroy(hPtrsFree); hPtrsFree = nullptr; }
    cudaFreeHost(hHist);    hHist = nullptr;    hHistCap = 0;
    cudaFreeHost(hSeeds);   hSeeds = nullptr;   hSeedsCap = 0;
    cudaFree(dSeeds); dSeeds = nullptr; dSeedsCap = 0;
    if (hStageFree) { cudaEventDestroy(hStageFree); hStageFree = nullptr; }
  }

  void free_maps()
  {
    cudaFree(dHi2ho); cudaFree(dWi2wo); cudaFree(dLastRow); cudaFree(dLastCol);
    dHi2ho = dWi2wo = dLastRow = dLastCol = nullptr;
  }

  // index maps of the preview: i * min(outW/W, outH/H) truncated, in double as the reference computes them
  // (webcam/object_sensor/.../cv_ball_detector_seqpass.hpp:371-387), and their "last writer" inverses
  bool build_maps()
  {
    free_maps();
    const int W = geo.width, H = geo.height;
    if (W <= 0 || H <= 0 || outWidth <= 0 || outHeight <= 0)
      return true;
    const double a = static_cast<double>(outWidth) / W, b = static_cast<double>(outHeight) / H;
    const double shift = a < b ? a : b;
    std::vector<int32_t> wi(W), hi(H), lastCol(outWidth, -1), lastRow(outHeight, -1);
    for (int i = 0; i < W; ++i) wi[i] = (int32_t)(uint32_t)(i * shift);
    for (int i = 0; i < H; ++i) hi[i] = (int32_t)(uint32_t)(i * shift);
    // OL only writes source columns 5..W-5 (ov7670/line_sensor/.../cv_line_detector_seqpass.hpp:288)
    const int c0 = kind == KIND_OL ? 5 : 0, c1 = kind == KIND_OL ? W - 5 : W - 1;
    for (int i = c0; i <= c1; ++i) if (wi[i] < outWidth) lastCol[wi[i]] = i;
    for (int i = 0; i < H; ++i) if (hi[i] < outHeight) lastRow[hi[i]] = i;
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaMalloc(&dWi2wo, sizeof(int32_t) * W));
    CUDA_TRY(cudaMalloc(&dHi2ho, sizeof(int32_t) * H));
    CUDA_TRY(cudaMalloc(&dLastCol, sizeof(int32_t) * outWidth));
    CUDA_TRY(cudaMalloc(&dLastRow, sizeof(int32_t) * outHeight));
    CUDA_TRY(cudaMemcpy(dWi2wo, wi.data(), sizeof(int32_t) * W, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dHi2ho, hi.data(), sizeof(int32_t) * H, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dLastCol, lastCol.data(), sizeof(int32_t) * outWidth, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dLastRow, lastRow.data(), sizeof(int32_t) * outHeight, cudaMemcpyHostToDevice));
    return true;
  }

  bool init_device()
  {
    int dev = t_device;
    if (dev < 0)
      CUDA_TRY(cudaGetDevice(&dev));
    device = dev;
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (kind == KIND_OM)
    {
      uint32_t table[512];
      mxn_color_table(table);
      CUDA_TRY(cudaMalloc(&dMxnTable, sizeof(table)));
      CUDA_TRY(cudaMemcpy(dMxnTable, table, sizeof(table), cudaMemcpyHostToDevice));
    }
    return true;
  }

  size_t frame_bytes() const
  {
    return (size_t)geo.height * geo.lineLength * (kind_is_planar(kind) ? 2u : 1u);
  }

  template <typename T>
  bool grow_device(T*& p, size_t& cap, size_t want, bool zero)
  {
    if (want <= cap)
      return true;
    size_t ncap = cap ? cap : 1;
    while (ncap < want) ncap *= 2;
    T* np = nullptr;
    CUDA_TRY(cudaMalloc(&np, ncap * sizeof(T)));
    if (zero)
    {
      // complete before returning: the buffer may be used on a caller's stream next
      CUDA_TRY(cudaMemsetAsync(np, 0, ncap * sizeof(T), stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
    }
    if (p)
    {
      CUDA_TRY(cudaStreamSynchronize(stream));
      cudaFree(p);
    }
    p = np; cap = ncap;
    return true;
  }
  template <typename T>
  bool grow_pinned(T*& p, size_t& cap, size_t want)
  {
    if (want <= cap)
      return true;
    size_t ncap = cap ? cap : 1;
    while (ncap < want) ncap *= 2;
    T* np = nullptr;
    CUDA_TRY(cudaMallocHost(&np, ncap * sizeof(T)));
    if (p)
    {
      CUDA_TRY(cudaStreamSynchronize(stream));
      cudaFreeHost(p);
    }
    p = np; cap = ncap;
    return true;
  }
};

Instance* instance_of(TrikB200Handle* h) { return reinterpret_cast<Instance*>(h->impl); }

// CVAlgorithm::setup(): size rules of webcam/object_sensor/.../cv_ball_detector_seqpass.hpp:365-369
bool instance_setup(Instance* in, int width, int height, int lineLength, int outW, int outH, int outLine)
{
  in->valid = false;
  in->state = CarriedState();
  in->streamStates.clear();
  if (width < 0 || height < 0 || width % 32 != 0 || height % 4 != 0)
    return false;
  // OO: bitmap row offsets and labels are uint16 in the reference (cv_bitmap_builder_reference.hpp:93-96);
  // sizes where they would wrap are outside what it supports and are refused here
  if (in->kind == KIND_OO && (long long)(width / 4) * (height / 4) > 65536LL)
    return false;
  // The kernels read rows with 16-byte loads at row * lineLength + 16 * chunk: a stride that is not a multiple of 16 or
  // does not cover a row (2 * W bytes for YUV422, W per plane for YUV422P) cannot be served.  The reference accepts any
  // value here and fails (or reads out of bounds) later; this is the documented deviation (DESIGN.md "Boundary").
  if (width > 0 && height > 0
      && ((lineLength & 15) != 0 || lineLength < (kind_is_planar(in->kind) ? width : 2 * width)))
    return false;
  in->geo.width = width;
  in->geo.height = height;
  in->geo.lineLength = lineLength;
  in->geo.frameStride = 0;
  in->outWidth = outW; in->outHeight = outH; in->outLineLength = outLine;
  in->geo.drawInfo = nullptr;
  if (!in->build_maps())
    return false;
  in->valid = true;
  return true;
}

// ---------------------------------------------------------------------------------------------
// the batch core: n frames, host or device, through the kernels
// ---------------------------------------------------------------------------------------------
struct BatchView {
  int            n;
  const uint8_t* frames; int64_t frameStride; bool framesOnDevice;
  const uint8_t* inArgs; int inStride;
  uint8_t*       outArgs; int outStride; bool outOnDevice;
  const int64_t* seeds; bool seedsBroadcast;
  cudaStream_t   stream;
  bool           async;
  bool           deviceTail = false;     // TRIKB200_BATCH_DEVICE_TAIL
  // scattered form (trikb200_processMixed): one pointer per frame instead of base + stride
  const int32_t* streamIds = nullptr; int numStreams = 0;
  uint8_t* previews = nullptr; int64_t previewStride = 0; bool previewsOnDevice = false;
  const uint8_t* const* framePtrs = nullptr;
  const uint8_t* const* inPtrs = nullptr;
  uint8_t* const*       outPtrs = nullptr;
  CarriedState* const*  statePtrs = nullptr;   // per frame: the carried state of the handle the frame belongs to

  const uint8_t* in_args(int i) const { return inPtrs ? inPtrs[i] : inArgs + (size_t)i * inStride; }
  uint8_t* out_args(int i) const { return outPtrs ? outPtrs[i] : outArgs + (size_t)i * outStride; }
  bool broadcast_in() const { return !inPtrs && inStride == 0; }
};

// what enqueue_batch leaves for finish_batch
struct Pending {
  cudaStream_t s = nullptr;
  size_t recBytes = 0;
  int numFlagged = 0;
  bool hostTail = false;
  int histBins = 0;
  bool needsFinish = false;
  // small host batches: the preview sits in pinned memory after the stream has drained and is copied out by the CPU
  const uint8_t* previewSrc = nullptr; size_t previewBytes = 0;
};

size_t result_record_bytes(int kind)
{
  switch (kind)
  {
    case KIND_OO: return 36;
    case KIND_OM: return 400;
    default:      return sizeof(TargetOut);
  }
}

bool wants_autodetect(int kind, const void* inArgsAlg)
{
  switch (kind)
  {
    case KIND_OO: return reinterpret_cast<const TRIKB200_ObjInArgsAlg*>(inArgsAlg)->autoDetectHsv != 0;
    case KIND_OM: return false;
    default:      return reinterpret_cast<const TRIKB200_RangeInArgsAlg*>(inArgsAlg)->autoDetectHsv != 0;
  }
}

// merge a device result record into the caller's OutArgsAlg, touching only the fields the
// reference's run() assigns for this call
void merge_result(int kind, const void* inArgsAlg, const uint8_t* rec, const uint16_t* detect, uint8_t* dst)
{
  switch (kind)
  {
    case KIND_WO: case KIND_WL: case KIND_OL:
    {
      const TargetOut* r = reinterpret_cast<const TargetOut*>(rec);
      TRIKB200_TargetOutArgsAlg* o = reinterpret_cast<TRIKB200_TargetOutArgsAlg*>(dst);
      o->targetX = r->targetX; o->targetY = r->targetY; o->targetSize = r->targetSize;
      if (wants_autodetect(kind, inArgsAlg))
      {
        if (detect)
        {
          o->detectHue = detect[0]; o->detectHueTolerance = detect[1];
          o->detectSat = detect[2]; o->detectSatTolerance = detect[3];
          o->detectVal = detect[4]; o->detectValTolerance = detect[5];
        }
        else
        {
          o->detectHue = r->detectHue; o->detectHueTolerance = r->detectHueTolerance;
          o->detectSat = r->detectSat; o->detectSatTolerance = r->detectSatTolerance;
          o->detectVal = r->detectVal; o->detectValTolerance = r->detectValTolerance;
        }
      }
      break;
    }
    case KIND_OO:
    {
      TRIKB200_ObjOutArgsAlg* o = reinterpret_cast<TRIKB200_ObjOutArgsAlg*>(dst);
      memcpy(o->target, rec, sizeof(o->target));           // memset + fill of all eight (cv_ball_detector_seqpass.hpp:569-597)
      if (wants_autodetect(kind, inArgsAlg))
      {
        if (detect)                                         // host annealing tail
        {
          o->detectHue = detect[0]; o->detectHueTolerance = detect[1];
          o->detectSat = detect[2]; o->detectSatTolerance = detect[3];
          o->detectVal = detect[4]; o->detectValTolerance = detect[5];
        }
        else                                                // device tail: the record carries them
          memcpy(&o->detectHue, rec + sizeof(o->target), 6 * sizeof(uint16_t));
      }
      break;
    }
    case KIND_OM:
    {
      const TRIKB200_MxnInArgsAlg* ia = reinterpret_cast<const TRIKB200_MxnInArgsAlg*>(inArgsAlg);
      const int cells = ia->widthM * ia->heightN;          // validated by enqueue_batch: 1 <= cells <= 100
      memcpy(dst, rec, sizeof(int32_t) * (size_t)cells);   // only counter entries are assigned (:605-618)
      break;
    }
    default:
      break;
  }
}

// caller memory that the DMA engines can reach directly (cudaHostAlloc / cudaHostRegister): no staging needed
bool is_pinned_host(const void* p)
{
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess)
  {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeHost;
}

// Batches of 32..255 frames do not pay for a table on their own; the third one in a row under the same threshold set
// does (a camera rig calling once per time step).
bool lut_set_recurs(Instance* in, const FrameParams& fp)
{
  if (in->lutSeenCount > 0 && in->lutSeenFrom == fp.from && in->lutSeenTo == fp.to && in->lutSeenExpected == fp.expected)
    ++in->lutSeenCount;
  else
  {
    in->lutSeenFrom = fp.from; in->lutSeenTo = fp.to; in->lutSeenExpected = fp.expected;
    in->lutSeenCount = 1;
  }
  return in->lutSeenCount >= 3;
}

// build (or keep) the chroma table of a threshold set on stream s
bool ensure_lut(Instance* in, const FrameParams& fp, bool have, cudaStream_t s)
{
  if (have)
    return true;
  if (!in->dLutTable)
  {
    CUDA_TRY(cudaMalloc(&in->dLutTable, LUT_TABLE_BYTES));
    CUDA_TRY(cudaMalloc(&in->dLutMasks, LUT_MASK_BYTES));
    CUDA_TRY(cudaDeviceGetAttribute(&in->smCount, cudaDevAttrMultiProcessorCount, in->device));
  }
  in->lutValid = false;
  CUDA_TRY(launch_chroma_table(fp.from, fp.to, fp.expected, in->dLutTable, in->dLutMasks, s));
  in->lutFrom = fp.from; in->lutTo = fp.to; in->lutExpected = fp.expected;
  in->lutValid = true; in->lutStream = s;
  return true;
}

// the table of one of several threshold sets of a batch: cached per handle, built on stream s when missing
Instance::LutSlot* ensure_lut_slot(Instance* in, const FrameParams& fp, cudaStream_t s)
{
  Instance::LutSlot* victim = &in->lutCache[0];
  for (Instance::LutSlot& sl : in->lutCache)
  {
    if (sl.valid && sl.stream == s && sl.from == fp.from && sl.to == fp.to && sl.expected == fp.expected)
    {
      sl.lastUse = ++in->lutClock;
      return &sl;
    }
    if (sl.lastUse < victim->lastUse)
      victim = &sl;
  }
  if (!victim->table)
  {
    if (cudaMalloc(&victim->table, LUT_TABLE_BYTES) != cudaSuccess || cudaMalloc(&victim->masks, LUT_MASK_BYTES) != cudaSuccess)
    {
      set_error("chroma table allocation", cudaGetLastError());
      return nullptr;
    }
  }
  victim->valid = false;
  const cudaError_t e = launch_chroma_table(fp.from, fp.to, fp.expected, victim->table, victim->masks, s);
  if (e != cudaSuccess)
  {
    set_error("launch_chroma_table", e);
    return nullptr;
  }
  victim->from = fp.from; victim->to = fp.to; victim->expected = fp.expected;
  victim->valid = true; victim->stream = s; victim->lastUse = ++in->lutClock;
  return victim;
}

// the mxn colour-bin table of a device; built (and waited for, once) on first use so that every stream may read it
const uint16_t* ensure_om_table(int device, cudaStream_t s)
{
  if (device < 0 || device >= MAX_DEVICES)
    return nullptr;
  std::lock_guard<std::mutex> lock(g_omTableMutex);
  if (g_omTable[device])
    return g_omTable[device];
  uint16_t* t = nullptr;
  cudaError_t e = cudaMalloc(&t, OM_TABLE_BYTES);
  if (e == cudaSuccess) e = launch_om_bin_table(t, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess)
  {
    set_error("mxn colour-bin table", e);
    cudaFree(t);
    return nullptr;
  }
  g_omTable[device] = t;
  return t;
}

bool enqueue_batch(Instance* in, const BatchView& b, Pending& pend)
{
  if (!in->valid || in->geo.width <= 0 || in->geo.height <= 0)
  {
    set_error("algorithm not set up for a non-empty image");
    return false;
  }
  const int kind = in->kind;
  CUDA_TRY(cudaSetDevice(in->device));
  cudaStream_t s = b.stream ? b.stream : in->stream;
  const size_t fbytes = in->frame_bytes();
  const size_t recBytes = result_record_bytes(kind);
  const size_t inSize = in_args_alg_size(kind);

  // 0. argument checks that the reference leaves to undefined behaviour
  int maxRows = 0, maxCols = 0, numFlagged = 0;
  const size_t nargs = b.broadcast_in() ? 1 : (size_t)b.n;
  for (size_t i = 0; i < nargs; ++i)
  {
    const uint8_t* ia = b.in_args((int)i);
    if (kind == KIND_OM)
    {
      const TRIKB200_MxnInArgsAlg* a = reinterpret_cast<const TRIKB200_MxnInArgsAlg*>(ia);
      const int M = a->widthM, N = a->heightN;
      if (M < 1 || N < 1 || M > 100 || N > 100 || M * N > 100)
      {
        // the reference divides by zero / overruns outColor[100] (mxn_sensor/.../cv_ball_detector_seqpass.hpp:587-588,612)
        set_error("mxn sensor: widthM and heightN must be >= 1 with widthM*heightN <= 100");
        return false;
      }
      if (M > maxRows) maxRows = M;
      if (N > maxCols) maxCols = N;
    }
    else if (wants_autodetect(kind, ia))
      numFlagged += b.broadcast_in() ? b.n : 1;
  }
  const bool hostTail = numFlagged > 0 && kind != KIND_WO && !b.deviceTail;
  if (hostTail && (b.async || b.outOnDevice))
  {
    set_error("annealed auto-detect needs its host tail: not available with TRIKB200_BATCH_ASYNC or device results");
    return false;
  }
  (void)inSize;

  // 1. per-frame parameters (carried state advances frame by frame, as n process() calls would)
  bool broadcast = b.broadcast_in() && kind != KIND_OL && kind != KIND_OO;
  if (b.streamIds && (int)in->streamStates.size() < b.numStreams)
    in->streamStates.resize((size_t)b.numStreams);
  size_t np = broadcast ? 1 : (size_t)b.n;
  in->paramsScratch.resize(np);
  // The webcam sensors carry no state: their parameters are a function of the argument bytes alone, and a batch usually
  // holds a handful of distinct argument sets (per-camera thresholds).  The first LUT_MAX_SETS distinct ones are remembered
  // with their parameters, so a frame costs one short compare instead of the conversion, and its set number falls out for
  // the table path below (argSet[i], -1 once a batch has more sets than that).
  const bool stateless = kind == KIND_WO || kind == KIND_WL;
  struct ArgMemo { uint8_t bytes[sizeof(TRIKB200_RangeInArgsAlg)]; FrameParams fp; };
  ArgMemo memo[LUT_MAX_SETS];
  int numMemo = 0;
  bool memoComplete = stateless && !broadcast;           // every frame's set is known
  std::vector<int>& argSet = in->argSetScratch;
  if (memoComplete)
    argSet.resize(np);
  for (size_t i = 0; i < np; ++i)
  {
    const uint8_t* ia = b.in_args((int)i);
    if (stateless && !broadcast)
    {
      int k = 0;
      for (; k < numMemo; ++k)
        if (std::memcmp(memo[k].bytes, ia, sizeof(memo[k].bytes)) == 0)
          break;
      if (k < numMemo)
      {
        in->paramsScratch[i] = memo[k].fp;
        argSet[i] = k;
        continue;
      }
      prepare_frame_params(kind, in->geo, ia, in->state, in->paramsScratch[i]);
      if (numMemo < LUT_MAX_SETS)
      {
        std::memcpy(memo[numMemo].bytes, ia, sizeof(memo[numMemo].bytes));
        memo[numMemo].fp = in->paramsScratch[i];
        argSet[i] = numMemo++;
      }
      else
      {
        memoComplete = false;
        argSet[i] = -1;
      }
      continue;
    }
    prepare_frame_params(kind, in->geo, ia,
                         b.statePtrs ? *b.statePtrs[i] : b.streamIds ? in->streamStates[(size_t)b.streamIds[i]] : in->state,
                         in->paramsScratch[i]);
  }
  // A per-frame array (or pointer list) of arguments that all say the same thing is a broadcast: one record, and the
  // chroma-table path of the webcam object sensor stays available (a batch gathered from many handles looks like this).
  if (memoComplete && numMemo == 1 && np > 1)
  {
    broadcast = true;
    np = 1;
    in->paramsScratch.resize(1);
  }
  const FrameParams* const dParamsBefore = in->dParams;
  if (!in->grow_device(in->dParams, in->dParamsCap, np, false)) return false;
  if (in->dParams != dParamsBefore)
    in->dBroadcastValid = false;
  // A broadcast record that dParams[0] already holds (uploaded on this stream) is not sent again, so
  // back-to-back batches with unchanged arguments are kernel launches only.  Otherwise stage through
  // pinned memory; an asynchronous upload may still be reading it, hence the event.
  const bool resident = broadcast && in->dBroadcastValid && in->dBroadcastStream == s
                        && std::memcmp(&in->dBroadcast, &in->paramsScratch[0], sizeof(FrameParams)) == 0;
  if (!resident)
  {
    if (in->hParamsFree)
      CUDA_TRY(cudaEventSynchronize(in->hParamsFree));
    else
      CUDA_TRY(cudaEventCreateWithFlags(&in->hParamsFree, cudaEventDisableTiming));
    if (!in->grow_pinned(in->hParams, in->hParamsCap, np)) return false;
    std::memcpy(in->hParams, in->paramsScratch.data(), np * sizeof(FrameParams));
    CUDA_TRY(cudaMemcpyAsync(in->dParams, in->hParams, np * sizeof(FrameParams), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaEventRecord(in->hParamsFree, s));
    in->dBroadcastValid = broadcast;
    if (broadcast)
    {
      in->dBroadcast = in->paramsScratch[0];
      in->dBroadcastStream = s;
    }
  }

  // 2. frames
  const bool smallHost = !b.framesOnDevice && !b.framePtrs && !b.async && !b.outOnDevice && g_zeroCopyBytes > 0
                         && (size_t)b.n * fbytes <= (size_t)g_zeroCopyBytes;
  Geometry g = in->geo;
  const uint8_t* dFrames = b.frames;
  if (b.framePtrs)
  {
    // scattered frames: stage them one by one into the handle's device buffer
    const size_t stride = (fbytes + 15u) & ~(size_t)15u;
    if (!in->grow_device(in->dFrames, in->dFramesCap, stride * b.n, false)) return false;
    // pinned host frames (or device frames): one gather kernel reads them in place; anything else (pageable memory)
    // goes frame by frame through the copy engine, neighbours in memory merged into one copy
    bool gathered = false;
    if (g_gatherMode >= 0 && (g_gatherMode > 0 || b.n >= 8) && stride == fbytes)
    {
      if (in->hPtrsFree)
        CUDA_TRY(cudaEventSynchronize(in->hPtrsFree));
      else
        CUDA_TRY(cudaEventCreateWithFlags(&in->hPtrsFree, cudaEventDisableTiming));
      if (!in->grow_pinned(in->hPtrs, in->hPtrsCap, (size_t)b.n)) return false;
      if (!in->grow_device(in->dPtrs, in->dPtrsCap, (size_t)b.n, false)) return false;
      bool all = true;
      for (int i = 0; all && i < b.n; ++i)
      {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, b.framePtrs[i]) != cudaSuccess)
        {
          cudaGetLastError();
          all = false;
        }
        else if (b.framesOnDevice)
          all = (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) && attr.devicePointer != nullptr;
        else
          all = attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr;
        if (all)
          all = (((uintptr_t)attr.devicePointer) & 15u) == 0;
        if (all)
          in->hPtrs[i] = reinterpret_cast<const uint8_t*>(attr.devicePointer);
      }
      if (all)
      {
        CUDA_TRY(cudaMemcpyAsync(in->dPtrs, in->hPtrs, sizeof(uint8_t*) * b.n, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaEventRecord(in->hPtrsFree, s));
        CUDA_TRY(launch_gather_frames(in->dPtrs, in->dFrames, (long long)stride, fbytes, b.n, s));
        gathered = true;
      }
    }
    if (!gathered)
      for (int i = 0; i < b.n; )
      {
        int j = i + 1;
        while (j < b.n && stride == fbytes && b.framePtrs[j] == b.framePtrs[j - 1] + fbytes)
          ++j;
        CUDA_TRY(cudaMemcpyAsync(in->dFrames + (size_t)i * stride, b.framePtrs[i], fbytes * (size_t)(j - i),
                                 b.framesOnDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
        i = j;
      }
    dFrames = in->dFrames;
    g.frameStride = (int64_t)stride;
  }
  else if (b.framesOnDevice)
    g.frameStride = b.frameStride;
  else
  {
    const size_t stride = (fbytes + 15u) & ~(size_t)15u;
    if (smallHost && !is_pinned_host(b.frames))
    {
      // A few frames from (usually pageable) caller memory, synchronous call: the driver's staged copies cost more
      // than the kernels.  Stage through the handle's pinned buffers with plain CPU copies and let the DMA engine
      // move pinned <-> device; the small result records are written straight into pinned memory by the kernels.
      // (Letting the kernels read the frame in place over PCIe was measured slower: 152 vs 94 us per call.)
      if (!in->grow_pinned(in->hFrames, in->hFramesCap, stride * b.n)) return false;
      if (!in->grow_device(in->dFrames, in->dFramesCap, stride * b.n, false)) return false;
      for (int i = 0; i < b.n; ++i)
        std::memcpy(in->hFrames + (size_t)i * stride, b.frames + (size_t)i * b.frameStride, fbytes);
      CUDA_TRY(cudaMemcpyAsync(in->dFrames, in->hFrames, stride * b.n, cudaMemcpyHostToDevice, s));
      dFrames = in->dFrames;
    }
    else
    {
      if (!in->grow_device(in->dFrames, in->dFramesCap, stride * b.n, false)) return false;
      if (b.frameStride == (int64_t)stride || b.n == 1)
        CUDA_TRY(cudaMemcpyAsync(in->dFrames, b.frames, b.n == 1 ? fbytes : stride * b.n, cudaMemcpyHostToDevice, s));
      else
        CUDA_TRY(cudaMemcpy2DAsync(in->dFrames, stride, b.frames, (size_t)b.frameStride, fbytes, b.n, cudaMemcpyHostToDevice, s));
      dFrames = in->dFrames;
    }
    g.frameStride = (int64_t)stride;
  }

  // preview wanted: the result tails also leave their source-coordinate values for the overlays
  const bool wantPreview = b.previews != nullptr && in->outWidth > 0 && in->outHeight > 0 && in->dLastRow != nullptr;
  const size_t previewBytes = (size_t)in->outHeight * in->outLineLength;
  g.drawInfo = nullptr;
  if (wantPreview)
  {
    if (in->outLineLength < in->outWidth * 2)
    {
      set_error("outputLineLength smaller than outputWidth * 2");
      return false;
    }
    if (!in->grow_device(in->dDraw, in->dDrawCap, (size_t)b.n, true)) return false;
    g.drawInfo = in->dDraw;
  }

  // 3. the per-pixel kernels
  uint8_t* dOut;
  const bool directOut = b.outOnDevice && b.outStride == (int)recBytes && kind != KIND_OM;
  if (directOut)
    dOut = b.outArgs;
  else if (smallHost)
  {
    if (!in->grow_pinned(in->hOut, in->hOutCap, recBytes * b.n)) return false;
    dOut = in->hOut;                                   // the kernels write the records straight into pinned host memory
  }
  else
  {
    if (!in->grow_device(in->dOut, in->dOutCap, recBytes * b.n, false)) return false;
    dOut = in->dOut;
  }
  const int pstride = broadcast ? 0 : 1;
  const uint8_t* pvLutTable = nullptr;
  const uint32_t* pvLutMasks = nullptr;
  switch (kind)
  {
    case KIND_WO: case KIND_WL: case KIND_OL:
    {
      // WO batches that share one threshold set go through the chroma table (built once per set, kept with the
      // handle); below ~256 frames building it costs more than it saves unless it is already there
      bool useLut = false;
      pvLutTable = nullptr; pvLutMasks = nullptr;
      if (kind == KIND_WO && broadcast && g_lutMode >= 0 && g.width % 8 == 0)
      {
        const FrameParams& fp = in->paramsScratch[0];
        const bool have = in->lutValid && in->lutStream == s && in->lutFrom == fp.from && in->lutTo == fp.to
                          && in->lutExpected == fp.expected;
        useLut = g_lutMode > 0 || b.n >= 256 || (have && b.n >= 32) || (b.n >= 32 && lut_set_recurs(in, fp));
        if (useLut && !ensure_lut(in, fp, have, s)) return false;
      }
      // WO frames under SEVERAL threshold sets (per-stream thresholds gathered into one batch): partition the batch by set;
      // when there are at most LUT_CACHE sets of at least 32 frames each, every set goes through its own cached table
      bool multiLut = false;
      std::vector<int> setFirst, setCount;
      const std::vector<int>& setOf = argSet;
      if (kind == KIND_WO && !broadcast && memoComplete && g_lutMode >= 0 && g.width % 8 == 0 && (b.n >= 256 || g_lutMode > 0))
      {
        setFirst.assign((size_t)numMemo, -1); setCount.assign((size_t)numMemo, 0);
        for (int i = 0; i < b.n; ++i)
        {
          const size_t k = (size_t)setOf[(size_t)i];
          if (setFirst[k] < 0) setFirst[k] = i;
          ++setCount[k];
        }
        multiLut = true;
        for (size_t k = 0; multiLut && k < setCount.size(); ++k)
          multiLut = setCount[k] >= 32;
      }
      if (multiLut)
      {
        if (!in->smCount)
          CUDA_TRY(cudaDeviceGetAttribute(&in->smCount, cudaDevAttrMultiProcessorCount, in->device));
        if (in->hLutListFree)
          CUDA_TRY(cudaEventSynchronize(in->hLutListFree));
        else
          CUDA_TRY(cudaEventCreateWithFlags(&in->hLutListFree, cudaEventDisableTiming));
        if (!in->grow_pinned(in->hLutList, in->hLutListCap, (size_t)b.n)) return false;
        if (!in->grow_device(in->dLutList, in->dLutListCap, (size_t)b.n, false)) return false;
        if (!in->grow_device(in->dAcc, in->dAccCap, (size_t)b.n, true)) return false;
        std::vector<int> ofs(setCount.size() + 1, 0), fill(setCount.size(), 0);
        for (size_t k = 0; k < setCount.size(); ++k) ofs[k + 1] = ofs[k] + setCount[k];
        for (int i = 0; i < b.n; ++i)
        {
          const size_t k = (size_t)setOf[(size_t)i];
          in->hLutList[ofs[k] + fill[k]++] = i;
        }
        CUDA_TRY(cudaMemcpyAsync(in->dLutList, in->hLutList, sizeof(int) * b.n, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaEventRecord(in->hLutListFree, s));
        LutSets sets{};
        sets.numSets = (int)setCount.size();
        for (size_t k = 0; k < setCount.size(); ++k)
        {
          Instance::LutSlot* sl = ensure_lut_slot(in, in->paramsScratch[(size_t)setFirst[k]], s);
          if (!sl) return false;
          sets.table[k] = sl->table; sets.masks[k] = sl->masks;
          sets.listOffset[k] = ofs[k]; sets.count[k] = setCount[k]; sets.paramIndex[k] = setFirst[k];
        }
        CUDA_TRY(launch_wo_lut_sets(g, dFrames, in->dParams, reinterpret_cast<TargetOut*>(dOut), in->smCount, s, in->dAcc,
                                    in->dLutList, sets));
      }
      else if (useLut)
      {
        pvLutTable = in->dLutTable; pvLutMasks = in->dLutMasks;   // the preview pass of this batch detects through it too
        if (!in->grow_device(in->dAcc, in->dAccCap, (size_t)b.n, true)) return false;
        CUDA_TRY(launch_wo_lut(g, b.n, dFrames, in->dParams, in->dLutTable, in->dLutMasks,
                               reinterpret_cast<TargetOut*>(dOut), in->smCount, s, in->dAcc));
      }
      else
      {
        if (!in->grow_device(in->dAcc, in->dAccCap, (size_t)b.n, true)) return false;
        CUDA_TRY(launch_sum_sensor(kind, g, b.n, dFrames, in->dParams, pstride, in->dAcc,
                                   reinterpret_cast<TargetOut*>(dOut), g_slabsPerFrame, s));
      }
      break;
    }
    case KIND_OM:
      if (b.outOnDevice || b.async)                       // whole records reach the caller: entries past M*N are 0
        CUDA_TRY(cudaMemsetAsync(dOut, 0, recBytes * b.n, s));
      if (g_mxnTableMode >= 0)
      {
        // the colour bin is one fixed function of (Y,U,V): gather it from the device's 2^24-entry table (identical results)
        const uint16_t* binTable = ensure_om_table(in->device, s);
        if (!binTable) return false;
        if (g_mxnTableMode == 0)
        {
          // majority pass first (a cell whose colour bin provably holds more than half of its pixels needs no
          // histogram); the cell rows it leaves undecided go through the histogram kernel, taken from its list
          if (!in->smCount)
            CUDA_TRY(cudaDeviceGetAttribute(&in->smCount, cudaDevAttrMultiProcessorCount, in->device));
          if (!in->grow_device(in->dOmList, in->dOmListCap, (size_t)b.n * maxRows, false)) return false;
          if (!in->grow_device(in->dOmCount, in->dOmCountCap, 2, true)) return false;
          int* const cnt = in->dOmCount + in->omParity;                   // alternate batches use alternate counters:
          int* const cntNext = in->dOmCount + (in->omParity ^ 1);         // each batch zeroes the one the next will use
          in->omParity ^= 1;
          CUDA_TRY(launch_om_major(g, b.n, dFrames, in->dParams, pstride, binTable, in->dMxnTable,
                                   reinterpret_cast<int32_t*>(dOut), maxRows, in->dOmList, cnt, cntNext, s));
          CUDA_TRY(launch_om_table(g, b.n, dFrames, in->dParams, pstride, binTable, in->dMxnTable,
                                   reinterpret_cast<int32_t*>(dOut), maxRows, maxCols, s, in->dOmList, cnt, in->smCount));
        }
        else
          CUDA_TRY(launch_om_table(g, b.n, dFrames, in->dParams, pstride, binTable, in->dMxnTable,
                                   reinterpret_cast<int32_t*>(dOut), maxRows, maxCols, s));
      }
      else
        CUDA_TRY(launch_om(g, b.n, dFrames, in->dParams, pstride, in->dMxnTable, reinterpret_cast<int32_t*>(dOut), maxRows, maxCols, s));
      break;
    case KIND_OO:
    {
      const size_t cells = (size_t)(g.width / 4) * (g.height / 4);
      const int maxLabels = oo_max_labels(g.width, g.height);
      if (!in->grow_device(in->dBitmaps, in->dBitmapsCap, cells * b.n, false)) return false;
      if (!in->grow_device(in->dClusters, in->dClustersCap, (size_t)12 * maxLabels * b.n, false)) return false;
      if (!in->grow_device(in->dEqual, in->dEqualCap, (size_t)maxLabels * b.n, false)) return false;
      // all frames under one threshold set (the usual case: the range is carried state): step 1 through the chroma table
      bool useLut = g_lutMode >= 0;
      const FrameParams& fp0 = in->paramsScratch[0];
      for (size_t i = 1; useLut && i < np; ++i)
        useLut = in->paramsScratch[i].from == fp0.from && in->paramsScratch[i].to == fp0.to
                 && in->paramsScratch[i].expected == fp0.expected;
      if (useLut)
      {
        const bool have = in->lutValid && in->lutStream == s && in->lutFrom == fp0.from && in->lutTo == fp0.to
                          && in->lutExpected == fp0.expected;
        useLut = g_lutMode > 0 || b.n >= 256 || (have && b.n >= 32) || (b.n >= 32 && lut_set_recurs(in, fp0));
        if (useLut && !ensure_lut(in, fp0, have, s)) return false;
      }
      // frames under several ranges (object sensor instances gathered into one batch, each carrying its own range): up to
      // LUT_MAX_SETS ranges of at least 32 frames each go through their cached tables in one step-1 launch
      LutSets sets{};
      bool multiLut = false;
      if (!useLut && g_lutMode >= 0 && (b.n >= 256 || g_lutMode > 0) && g.width / 16 <= 768)
      {
        std::vector<int>& setOf = in->argSetScratch;
        setOf.resize((size_t)b.n);
        int setFirst[LUT_MAX_SETS], setCount[LUT_MAX_SETS], numSets = 0;
        multiLut = true;
        for (int i = 0; multiLut && i < b.n; ++i)
        {
          const FrameParams& fp = in->paramsScratch[(size_t)i];
          int k = 0;
          for (; k < numSets; ++k)
          {
            const FrameParams& f0 = in->paramsScratch[(size_t)setFirst[k]];
            if (f0.from == fp.from && f0.to == fp.to && f0.expected == fp.expected)
              break;
          }
          if (k == numSets)
          {
            if (numSets == LUT_MAX_SETS) { multiLut = false; break; }
            setFirst[numSets] = i; setCount[numSets] = 0; ++numSets;
          }
          setOf[(size_t)i] = k; ++setCount[k];
        }
        for (int k = 0; multiLut && k < numSets; ++k)
          multiLut = setCount[k] >= 32;
        if (multiLut)
        {
          if (!in->smCount)
            CUDA_TRY(cudaDeviceGetAttribute(&in->smCount, cudaDevAttrMultiProcessorCount, in->device));
          if (in->hLutListFree)
            CUDA_TRY(cudaEventSynchronize(in->hLutListFree));
          else
            CUDA_TRY(cudaEventCreateWithFlags(&in->hLutListFree, cudaEventDisableTiming));
          if (!in->grow_pinned(in->hLutList, in->hLutListCap, (size_t)b.n)) return false;
          if (!in->grow_device(in->dLutList, in->dLutListCap, (size_t)b.n, false)) return false;
          int ofs[LUT_MAX_SETS + 1] = {0}, fill[LUT_MAX_SETS] = {0};
          for (int k = 0; k < numSets; ++k) ofs[k + 1] = ofs[k] + setCount[k];
          for (int i = 0; i < b.n; ++i)
          {
            const int k = setOf[(size_t)i];
            in->hLutList[ofs[k] + fill[k]++] = i;
          }
          CUDA_TRY(cudaMemcpyAsync(in->dLutList, in->hLutList, sizeof(int) * b.n, cudaMemcpyHostToDevice, s));
          CUDA_TRY(cudaEventRecord(in->hLutListFree, s));
          sets.numSets = numSets;
          for (int k = 0; k < numSets; ++k)
          {
            Instance::LutSlot* sl = ensure_lut_slot(in, in->paramsScratch[(size_t)setFirst[k]], s);
            if (!sl) return false;
            sets.table[k] = sl->table; sets.masks[k] = sl->masks;
            sets.listOffset[k] = ofs[k]; sets.count[k] = setCount[k]; sets.paramIndex[k] = setFirst[k];
          }
        }
      }
      CUDA_TRY(launch_oo(g, b.n, dFrames, in->dParams, pstride, in->dBitmaps, in->dClusters, in->dEqual, maxLabels, dOut, nullptr, s,
                         useLut ? in->dLutTable : nullptr, in->dLutMasks, in->smCount,
                         multiLut ? in->dLutList : nullptr, multiLut ? &sets : nullptr));
      break;
    }
    default:
      set_error("unknown sensor kind");
      return false;
  }

  // 4. auto-calibration histograms for the frames that ask for it
  const int histBins = (kind == KIND_OO) ? 1024 : 256;
  if (numFlagged > 0)
  {
    // pinned staging of the frame list (and seeds): an earlier asynchronous batch may still be reading it
    if (in->hStageFree)
      CUDA_TRY(cudaEventSynchronize(in->hStageFree));
    else
      CUDA_TRY(cudaEventCreateWithFlags(&in->hStageFree, cudaEventDisableTiming));
    if (!in->grow_pinned(in->hFlagged, in->hFlaggedCap, (size_t)numFlagged)) return false;
    if (!in->grow_device(in->dFlagged, in->dFlaggedCap, (size_t)numFlagged, false)) return false;
    int k = 0;
    for (int i = 0; i < b.n; ++i)
      if (wants_autodetect(kind, b.in_args(i)))
        in->hFlagged[k++] = i;
    CUDA_TRY(cudaMemcpyAsync(in->dFlagged, in->hFlagged, sizeof(int) * numFlagged, cudaMemcpyHostToDevice, s));
    const bool deviceTail = b.deviceTail && kind != KIND_WO;
    if (deviceTail)
    {
      if (!in->grow_pinned(in->hSeeds, in->hSeedsCap, (size_t)numFlagged)) return false;
      if (!in->grow_device(in->dSeeds, in->dSeedsCap, (size_t)numFlagged, false)) return false;
      const unsigned now = (unsigned)time(NULL);                      // the reference's srand(time(NULL))
      for (int j = 0; j < numFlagged; ++j)
      {
        const int64_t sd = b.seeds ? b.seeds[b.seedsBroadcast ? 0 : in->hFlagged[j]] : -1;
        in->hSeeds[j] = sd >= 0 ? (uint32_t)sd : now;
      }
      CUDA_TRY(cudaMemcpyAsync(in->dSeeds, in->hSeeds, sizeof(uint32_t) * numFlagged, cudaMemcpyHostToDevice, s));
    }
    CUDA_TRY(cudaEventRecord(in->hStageFree, s));
    if (kind == KIND_WO)
      CUDA_TRY(launch_wo_detect(g, numFlagged, dFrames, in->dFlagged, reinterpret_cast<TargetOut*>(dOut), s));
    else
    {
      const size_t words = (size_t)(histBins + 2) * numFlagged;
      if (!in->grow_device(in->dHist, in->dHistCap, words, false)) return false;
      if (!in->grow_pinned(in->hHist, in->hHistCap, words)) return false;
      CUDA_TRY(launch_ordered_hist(kind, g, numFlagged, dFrames, in->dFlagged, in->dHist, s));
      if (deviceTail)
        CUDA_TRY(launch_anneal(kind, numFlagged, in->dFlagged, in->dHist, in->dSeeds, dOut, s));
      else
        CUDA_TRY(cudaMemcpyAsync(in->hHist, in->dHist, words * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
  }

  // 4b. preview image: base layer gather + overlays, then to the caller
  if (wantPreview)
  {
    uint8_t* dPrev = b.previews;
    long long pstrideBytes = b.previewStride;
    const bool previewInPlace = smallHost && !b.previewsOnDevice && !is_pinned_host(b.previews);
    if (previewInPlace)
    {
      if (!in->grow_pinned(in->hPreview, in->hPreviewCap, previewBytes * b.n)) return false;
      pend.previewSrc = in->hPreview; pend.previewBytes = previewBytes;
    }
    if (!b.previewsOnDevice)
    {
      if (!in->grow_device(in->dPreview, in->dPreviewCap, previewBytes * b.n, false)) return false;
      dPrev = in->dPreview;
      pstrideBytes = (long long)previewBytes;
    }
    if (in->outLineLength > in->outWidth * 2)      // row padding is not produced by the kernels: keep it zero
      CUDA_TRY(cudaMemset2DAsync(dPrev, (size_t)pstrideBytes, 0, previewBytes, b.n, s));
    CUDA_TRY(launch_preview(kind, g, b.n, dFrames, in->dParams, pstride, in->dBitmaps, in->dDraw,
                            reinterpret_cast<const int32_t*>(dOut), in->outWidth, in->outHeight, in->outLineLength,
                            in->dLastRow, in->dLastCol, in->dHi2ho, in->dWi2wo, dPrev, pstrideBytes, s, pvLutTable, pvLutMasks));
    if (previewInPlace)
      CUDA_TRY(cudaMemcpyAsync(in->hPreview, dPrev, previewBytes * b.n, cudaMemcpyDeviceToHost, s));
    else if (!b.previewsOnDevice)
      CUDA_TRY(cudaMemcpy2DAsync(b.previews, (size_t)b.previewStride, dPrev, previewBytes, previewBytes, b.n, cudaMemcpyDeviceToHost, s));
  }

  // 5. results
  if (b.outOnDevice)
  {
    if (dOut != b.outArgs)
      CUDA_TRY(cudaMemcpy2DAsync(b.outArgs, (size_t)b.outStride, dOut, recBytes, recBytes, b.n, cudaMemcpyDeviceToDevice, s));
    if (!b.async)
      CUDA_TRY(cudaStreamSynchronize(s));
    return true;
  }
  if (b.async)
  {
    // pinned host results, whole records (documented: fields the sensor does not produce are 0)
    CUDA_TRY(cudaMemcpy2DAsync(b.outArgs, (size_t)b.outStride, dOut, recBytes, recBytes, b.n, cudaMemcpyDeviceToHost, s));
    return true;
  }
  if (dOut != in->hOut)
  {
    if (!in->grow_pinned(in->hOut, in->hOutCap, recBytes * b.n)) return false;
    CUDA_TRY(cudaMemcpyAsync(in->hOut, dOut, recBytes * b.n, cudaMemcpyDeviceToHost, s));
  }
  pend.s = s; pend.recBytes = recBytes; pend.numFlagged = numFlagged; pend.hostTail = hostTail;
  pend.histBins = histBins; pend.needsFinish = true;
  return true;
}

// wait for the stream, run the host tail (annealing) and merge the records into the caller's structs
bool finish_batch(Instance* in, const BatchView& b, const Pending& pend)
{
  const int kind = in->kind;
  const int numFlagged = pend.numFlagged, histBins = pend.histBins;
  const bool hostTail = pend.hostTail;
  const size_t recBytes = pend.recBytes;
  CUDA_TRY(cudaSetDevice(in->device));
  CUDA_TRY(cudaStreamSynchronize(pend.s));
  if (pend.previewSrc)
    for (int i = 0; i < b.n; ++i)
      std::memcpy(b.previews + (size_t)i * b.previewStride, pend.previewSrc + (size_t)i * pend.previewBytes, pend.previewBytes);

  // 6. host tail: anneal the flagged frames (threads across frames), then merge
  std::vector<uint16_t> detect;
  if (hostTail)
  {
    detect.assign((size_t)6 * numFlagged, 0);
    const int nthreads = std::max(1, std::min<int>(numFlagged, (int)std::thread::hardware_concurrency()));
    auto work = [&](int tid)
    {
      for (int k = tid; k < numFlagged; k += nthreads)
      {
        const int32_t* rec = in->hHist + (size_t)k * (histBins + 2);
        const int frame = in->hFlagged[k];
        const int64_t sd = b.seeds ? b.seeds[b.seedsBroadcast ? 0 : frame] : -1;
        const unsigned seed = sd >= 0 ? (unsigned)sd : (unsigned)time(NULL);
        if (kind == KIND_OO)
          anneal_oo(rec, rec[histBins], seed, &detect[(size_t)6 * k]);
        else
          anneal_line(rec, rec[histBins], kind == KIND_OL, seed, &detect[(size_t)6 * k]);
      }
    };
    if (nthreads == 1)
      work(0);
    else
    {
      std::vector<std::thread> pool;
      for (int tIdx = 0; tIdx < nthreads; ++tIdx) pool.emplace_back(work, tIdx);
      for (auto& th : pool) th.join();
    }
  }
  int k = 0;
  for (int i = 0; i < b.n; ++i)
  {
    const uint8_t* ia = b.in_args(i);
    const uint16_t* det = nullptr;
    if (hostTail && wants_autodetect(kind, ia))
      det = &detect[(size_t)6 * k++];
    merge_result(kind, ia, in->hOut + (size_t)i * recBytes, det, b.out_args(i));
  }
  return true;
}

bool run_batch(Instance* in, const BatchView& b)
{
  Pending pend;
  if (!enqueue_batch(in, b, pend))
    return false;
  return pend.needsFinish ? finish_batch(in, b, pend) : true;
}

// ---------------------------------------------------------------------------------------------
// handle / dispatch layer: src/vidtranscode_cv.cpp
// ---------------------------------------------------------------------------------------------
XDAS_Int32 handle_setup_image_desc(TrikB200Handle* h)     // vidtranscode_cv.cpp:90-145
{
  Instance* in = instance_of(h);
  const IVIDTRANSCODE_Params& pb = h->params.base;
  const TRIK_VIDTRANSCODE_CV_DynamicParams& dp = h->dynamicParams;
  if (pb.numOutputStreams != 0 && pb.numOutputStreams != 1)
    return IALG_EFAIL;

  // TrikCvImageDimension is XDAS_Int16 (include/internal/vidtranscode_cv.h:32)
  const XDAS_Int16 inW = (XDAS_Int16)(dp.inputWidth  > 0 ? dp.inputWidth  : 0);
  const XDAS_Int16 inH = (XDAS_Int16)(dp.inputHeight > 0 ? dp.inputHeight : 0);
  const XDAS_Int32 inLine = dp.inputLineLength > 0 ? dp.inputLineLength : 0;
  XDAS_Int16 outW = 0, outH = 0;
  XDAS_Int32 outLine = 0, outFormat = TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_UNKNOWN;
  if (pb.numOutputStreams == 1)
  {
    outFormat = pb.formatOutput[0];
    outW = (XDAS_Int16)(dp.base.outputWidth[0]  > 0 ? dp.base.outputWidth[0]  : 0);
    outH = (XDAS_Int16)(dp.base.outputHeight[0] > 0 ? dp.base.outputHeight[0] : 0);
    outLine = dp.outputLineLength[0] > 0 ? dp.outputLineLength[0] : 0;
  }
  if (   (inW  < 0 || inW  > pb.maxWidthInput)
      || (inH  < 0 || inH  > pb.maxHeightInput)
      || (outW < 0 || outW > pb.maxWidthOutput[0])
      || (outH < 0 || outH > pb.maxHeightOutput[0]))
    return IALG_EFAIL;

  // only the sensor's own (input format, RGB565X) pair exists (vidtranscode_cv.cpp:76-84)
  if (pb.formatInput != default_input_format(h->kind) || outFormat != TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_RGB565X)
    return IALG_EFAIL;

  if (!instance_setup(in, inW, inH, inLine, outW, outH, outLine))
    return IALG_EFAIL;
  return IALG_EOK;
}

XDAS_Int32 handle_setup_params(TrikB200Handle* h, const TRIK_VIDTRANSCODE_CV_Params* params)   // :150-197
{
  TRIK_VIDTRANSCODE_CV_Params def;
  memset(&def, 0, sizeof(def));
  def.base.size = sizeof(TRIK_VIDTRANSCODE_CV_Params);
  def.base.numOutputStreams = 1;
  def.base.formatInput = default_input_format(h->kind);
  def.base.formatOutput[0] = TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_RGB565X;
  def.base.formatOutput[1] = TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_UNKNOWN;
  def.base.maxHeightInput = 480;
  def.base.maxWidthInput = 640;
  def.base.maxFrameRateInput = 60000;
  def.base.maxBitRateInput = -1;
  def.base.maxHeightOutput[0] = 480; def.base.maxHeightOutput[1] = -1;
  def.base.maxWidthOutput[0] = 640;  def.base.maxWidthOutput[1] = -1;
  def.base.maxFrameRateOutput[0] = def.base.maxFrameRateOutput[1] = -1;
  def.base.maxBitRateOutput[0] = def.base.maxBitRateOutput[1] = -1;
  def.base.dataEndianness = XDM_BYTE;
  h->params = params ? *params : def;
  return IALG_EOK;
}

XDAS_Int32 handle_setup_dynamic_params(TrikB200Handle* h, const TRIK_VIDTRANSCODE_CV_DynamicParams* dyn)  // :202-286
{
  TRIK_VIDTRANSCODE_CV_DynamicParams def;
  memset(&def, 0, sizeof(def));
  def.base.size = sizeof(TRIK_VIDTRANSCODE_CV_DynamicParams);
  def.base.readHeaderOnlyFlag = 0;
  def.base.keepInputResolutionFlag[0] = XDAS_FALSE; def.base.keepInputResolutionFlag[1] = XDAS_TRUE;
  // line sensors default to a 240 wide x 320 high preview (line_sensor/src/vidtranscode_cv.cpp:214,218)
  const bool line = (h->kind == KIND_WL || h->kind == KIND_OL);
  def.base.outputHeight[0] = line ? 320 : 240; def.base.outputHeight[1] = 0;
  def.base.outputWidth[0]  = line ? 240 : 320; def.base.outputWidth[1] = 0;
  def.base.keepInputFrameRateFlag[0] = def.base.keepInputFrameRateFlag[1] = XDAS_TRUE;
  def.base.inputFrameRate = -1;
  def.base.outputFrameRate[0] = def.base.outputFrameRate[1] = -1;
  def.base.targetBitRate[0] = def.base.targetBitRate[1] = -1;
  def.base.rateControl[0] = def.base.rateControl[1] = IVIDEO_NONE;
  def.base.keepInputGOPFlag[0] = def.base.keepInputGOPFlag[1] = XDAS_TRUE;
  def.base.intraFrameInterval[0] = def.base.intraFrameInterval[1] = 1;
  def.base.interFrameInterval[0] = def.base.interFrameInterval[1] = 0;
  def.base.forceFrame[0] = def.base.forceFrame[1] = IVIDEO_NA_FRAME;
  def.base.frameSkipTranscodeFlag[0] = def.base.frameSkipTranscodeFlag[1] = XDAS_FALSE;
  def.inputHeight = -1; def.inputWidth = -1; def.inputLineLength = -1;
  def.outputLineLength[0] = def.outputLineLength[1] = -1;
  h->dynamicParams = dyn ? *dyn : def;
  return handle_setup_image_desc(h);
}

// ---------------------------------------------------------------------------------------------
// IALG + IVIDTRANSCODE entry points, shared by the five kinds
// ---------------------------------------------------------------------------------------------
Int alg_alloc(const IALG_Params*, IALG_Fxns**, IALG_MemRec memTab[])          // fxns.c:85-102
{
  memTab[0].size = sizeof(TrikB200Handle);
  memTab[0].alignment = 0;
  memTab[0].space = IALG_EXTERNAL;
  memTab[0].attrs = IALG_PERSIST;
  memTab[1].size = kFastRamSize;
  memTab[1].alignment = 0;
  memTab[1].space = IALG_DARAM0;
  memTab[1].attrs = IALG_PERSIST;
  return 2;
}

Int alg_free(IALG_Handle algHandle, IALG_MemRec memTab[])                      // fxns.c:114-136
{
  TrikB200Handle* h = reinterpret_cast<TrikB200Handle*>(algHandle);
  delete instance_of(h);                                                      // trikCvHandleDestroy
  h->impl = nullptr;
  memTab[0].base = h;
  memTab[0].size = sizeof(TrikB200Handle);
  memTab[0].alignment = 0;
  memTab[0].space = IALG_EXTERNAL;
  memTab[0].attrs = IALG_PERSIST;
  memTab[1].base = h->fastRam;
  memTab[1].size = (Uns)h->fastRamSize;
  memTab[1].alignment = 0;
  memTab[1].space = IALG_DARAM0;
  memTab[1].attrs = IALG_PERSIST;
  return 2;
}

Int alg_init(int kind, IALG_Handle algHandle, const IALG_MemRec memTab[], const IALG_Params* algParams)  // fxns.c:146-166
{
  TrikB200Handle* h = reinterpret_cast<TrikB200Handle*>(algHandle);
  h->kind = kind;
  h->fastRam = reinterpret_cast<XDAS_Int8*>(memTab[1].base);
  h->fastRamSize = memTab[1].size;
  Instance* in = new (std::nothrow) Instance();                               // trikCvHandleInit
  if (!in)
    return IALG_EFAIL;
  in->kind = kind;
  h->impl = in;
  if (!in->init_device())
    return IALG_EFAIL;
  XDAS_Int32 res;
  if ((res = handle_setup_params(h, reinterpret_cast<const TRIK_VIDTRANSCODE_CV_Params*>(algParams))) != IALG_EOK)
    return res;
  if ((res = handle_setup_dynamic_params(h, nullptr)) != IALG_EOK)
    return res;
  return IALG_EOK;
}

XDAS_Int32 vid_process(int kind, IVIDTRANSCODE_Handle algHandle, XDM1_BufDesc* inBufs, XDM_BufDesc* outBufs,
                       IVIDTRANSCODE_InArgs* vidInArgs, IVIDTRANSCODE_OutArgs* vidOutArgs)   // fxns.c:174-264
{
  TrikB200Handle* h = reinterpret_cast<TrikB200Handle*>(algHandle);
  if (   vidInArgs->size  != (XDAS_Int32)in_args_size(kind)
      || vidOutArgs->size != (XDAS_Int32)out_args_size(kind))
  {
    XDM_SETUNSUPPORTEDPARAM(vidOutArgs->extendedError);
    return IVIDTRANSCODE_EUNSUPPORTED;
  }
  if (inBufs->numBufs != 1 || outBufs->numBufs < h->params.base.numOutputStreams)
  {
    XDM_SETUNSUPPORTEDPARAM(vidOutArgs->extendedError);
    return IVIDTRANSCODE_EFAIL;
  }
  XDM1_SingleBufDesc* inBuf = &inBufs->descs[0];
  if (inBuf->buf == NULL || vidInArgs->numBytes < 0 || vidInArgs->numBytes > inBuf->bufSize)
  {
    XDM_SETUNSUPPORTEDPARAM(vidOutArgs->extendedError);
    return IVIDTRANSCODE_EFAIL;
  }

  XDM_SETACCESSMODE_READ(inBuf->accessMask);
  vidOutArgs->bitsConsumed            = vidInArgs->numBytes * CHAR_BIT;
  vidOutArgs->decodedPictureType      = IVIDEO_NA_PICTURE;
  vidOutArgs->decodedPictureStructure = IVIDEO_CONTENTTYPE_NA;
  vidOutArgs->decodedHeight           = h->dynamicParams.inputHeight;
  vidOutArgs->decodedWidth            = h->dynamicParams.inputWidth;

  XDM1_SingleBufDesc* outBuf = NULL;
  XDAS_Int8* outPtr = NULL;
  XDAS_Int32 outSize = 0;
  if (h->params.base.numOutputStreams == 1)
  {
    outBuf = &vidOutArgs->encodedBuf[0];
    outBuf->buf = outBufs->bufs[0];
    outBuf->bufSize = outBufs->bufSizes[0];
    outBuf->accessMask = 0;
    outPtr = outBuf->buf;
    outSize = outBuf->bufSize;
    memset(outPtr, 0, (size_t)outSize);                                        // fxns.c:234
  }

  // trikCvProceedImage + CVAlgorithm::run() entry checks (cv_ball_detector_seqpass.hpp:415-419)
  Instance* in = instance_of(h);
  uint8_t* inAlg  = reinterpret_cast<uint8_t*>(vidInArgs)  + sizeof(IVIDTRANSCODE_InArgs);
  uint8_t* outAlg = reinterpret_cast<uint8_t*>(vidOutArgs) + sizeof(IVIDTRANSCODE_OutArgs);
  bool ok = (in != nullptr) && in->valid;
  if (ok)
  {
    // run() checks height*lineLength against numBytes (cv_ball_detector_seqpass.hpp:415-416) -- for
    // YUV422P that covers the luma plane only although both planes are read (:345-349).  Same check
    // here; additionally both planes must lie inside the caller's BUFFER (bufSize), the one case
    // where the reference would read out of bounds (documented deviation, DESIGN.md "Boundary").
    if ((long long)in->geo.height * in->geo.lineLength > (long long)vidInArgs->numBytes)
      ok = false;
    if ((long long)in->frame_bytes() > (long long)inBuf->bufSize)
      ok = false;
    if ((long long)in->outHeight * in->outLineLength > (long long)outSize)
      ok = false;
  }
  if (ok)
  {
    outSize = in->outHeight * in->outLineLength;
    if (in->geo.width > 0 && in->geo.height > 0)
    {
      BatchView b{};
      b.n = 1;
      b.frames = reinterpret_cast<const uint8_t*>(inBuf->buf); b.frameStride = 0; b.framesOnDevice = false;
      b.inArgs = inAlg; b.inStride = (int)in_args_alg_size(kind);
      b.outArgs = outAlg; b.outStride = (int)out_args_alg_size(kind); b.outOnDevice = false;
      int64_t seed = in->seed >= 0 ? in->seed : (int64_t)time(NULL);
      b.seeds = &seed; b.seedsBroadcast = true;
      b.stream = nullptr; b.async = false;
      if (outPtr != NULL && outSize > 0)
      {
        b.previews = reinterpret_cast<uint8_t*>(outPtr); b.previewStride = outSize; b.previewsOnDevice = false;
      }
      ok = run_batch(in, b);
    }
    else
      ok = false; /* the reference would run its tail on a 0x0 image; nothing sensible to produce */
  }
  if (!ok)
  {
    XDM_SETCORRUPTEDDATA(vidOutArgs->extendedError);
    return IVIDTRANSCODE_EFAIL;
  }

  if (outBuf)
  {
    outBuf->bufSize = outSize;
    XDM_SETACCESSMODE_WRITE(outBuf->accessMask);
    vidOutArgs->bitsGenerated[0]               = outBuf->bufSize * CHAR_BIT;
    vidOutArgs->encodedPictureType[0]          = vidOutArgs->decodedPictureType;
    vidOutArgs->encodedPictureStructure[0]     = vidOutArgs->decodedPictureStructure;
    vidOutArgs->outputID[0]                    = vidInArgs->inputID;
    vidOutArgs->inputFrameSkipTranscodeFlag[0] = XDAS_FALSE;
  }
  vidOutArgs->outBufsInUseFlag = XDAS_FALSE;
  return IVIDTRANSCODE_EOK;
}

XDAS_Int32 vid_control(int, IVIDTRANSCODE_Handle algHandle, IVIDTRANSCODE_Cmd cmd,
                       IVIDTRANSCODE_DynamicParams* dynParams, IVIDTRANSCODE_Status* status)   // fxns.c:272-334
{
  TrikB200Handle* h = reinterpret_cast<TrikB200Handle*>(algHandle);
  XDAS_Int32 ret = IVIDTRANSCODE_EFAIL;
  XDM_CLEARACCESSMODE_READ(status->data.accessMask);
  XDM_CLEARACCESSMODE_WRITE(status->data.accessMask);
  switch (cmd)
  {
    case XDM_GETSTATUS:
    case XDM_GETBUFINFO:
      status->extendedError = 0;
      status->bufInfo.minNumInBufs = 1;
      status->bufInfo.minNumOutBufs = 1;
      status->bufInfo.minInBufSize[0] = 0;
      status->bufInfo.minOutBufSize[0] = 0;
      XDM_SETACCESSMODE_WRITE(status->data.accessMask);
      ret = IVIDTRANSCODE_EOK;
      break;
    case XDM_SETPARAMS:
      if (dynParams->size == (XDAS_Int32)sizeof(TRIK_VIDTRANSCODE_CV_DynamicParams))
        ret = handle_setup_dynamic_params(h, reinterpret_cast<TRIK_VIDTRANSCODE_CV_DynamicParams*>(dynParams));
      else
        ret = IVIDTRANSCODE_EUNSUPPORTED;
      break;
    case XDM_RESET:
    case XDM_SETDEFAULT:
      ret = handle_setup_dynamic_params(h, nullptr);
      break;
    case XDM_FLUSH:
      ret = IVIDTRANSCODE_EOK;
      break;
    case XDM_GETVERSION:
      if (status->data.buf != NULL && status->data.bufSize >= (XDAS_Int32)(strlen(kVersion) + 1))
      {
        memcpy(status->data.buf, kVersion, strlen(kVersion) + 1);
        XDM_SETACCESSMODE_WRITE(status->data.accessMask);
        ret = IVIDTRANSCODE_EOK;
      }
      else
        ret = IVIDTRANSCODE_EFAIL;
      break;
    default:
      ret = IVIDTRANSCODE_EFAIL;
      break;
  }
  return ret;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// exported tables, one per sensor kind
// ---------------------------------------------------------------------------------------------
#define TRIKB200_DEFINE_KIND(NAME, KIND)                                                                        \
  static Int NAME##_init(IALG_Handle hh, const IALG_MemRec mt[], IALG_Handle, const IALG_Params* p)             \
  { return alg_init(KIND, hh, mt, p); }                                                                         \
  static XDAS_Int32 NAME##_process(IVIDTRANSCODE_Handle hh, XDM1_BufDesc* ib, XDM_BufDesc* ob,                  \
                                   IVIDTRANSCODE_InArgs* ia, IVIDTRANSCODE_OutArgs* oa)                         \
  { return vid_process(KIND, hh, ib, ob, ia, oa); }                                                             \
  static XDAS_Int32 NAME##_control(IVIDTRANSCODE_Handle hh, IVIDTRANSCODE_Cmd c,                                \
                                   IVIDTRANSCODE_DynamicParams* dp, IVIDTRANSCODE_Status* st)                   \
  { return vid_control(KIND, hh, c, dp, st); }                                                                  \
  extern "C" { IALG_Fxns TRIKB200_##NAME##_IALG = {                                                             \
    &TRIKB200_##NAME##_IALG, NULL, alg_alloc, NULL, NULL, alg_free, NAME##_init, NULL, NULL };                  \
  IVIDTRANSCODE_Fxns TRIKB200_##NAME##_FXNS = {                                                                 \
    { &TRIKB200_##NAME##_IALG, NULL, alg_alloc, NULL, NULL, alg_free, NAME##_init, NULL, NULL },                \
    NAME##_process, NAME##_control }; }

extern "C" {
extern IALG_Fxns TRIKB200_WO_IALG, TRIKB200_WL_IALG, TRIKB200_OO_IALG, TRIKB200_OL_IALG, TRIKB200_OM_IALG;
}
TRIKB200_DEFINE_KIND(WO, KIND_WO)
TRIKB200_DEFINE_KIND(WL, KIND_WL)
TRIKB200_DEFINE_KIND(OO, KIND_OO)
TRIKB200_DEFINE_KIND(OL, KIND_OL)
TRIKB200_DEFINE_KIND(OM, KIND_OM)

extern "C" {

IVIDTRANSCODE_Fxns* trikb200_fxns(XDAS_Int32 kind)
{
  switch (kind)
  {
    case KIND_WO: return &TRIKB200_WO_FXNS;
    case KIND_WL: return &TRIKB200_WL_FXNS;
    case KIND_OO: return &TRIKB200_OO_FXNS;
    case KIND_OL: return &TRIKB200_OL_FXNS;
    case KIND_OM: return &TRIKB200_OM_FXNS;
    default: return NULL;
  }
}

XDAS_Int32 trikb200_sizeofInArgsAlg(XDAS_Int32 kind)  { return (XDAS_Int32)in_args_alg_size(kind); }
XDAS_Int32 trikb200_sizeofOutArgsAlg(XDAS_Int32 kind) { return (XDAS_Int32)out_args_alg_size(kind); }
XDAS_Int32 trikb200_sizeofInArgs(XDAS_Int32 kind)     { return (XDAS_Int32)in_args_size(kind); }
XDAS_Int32 trikb200_sizeofOutArgs(XDAS_Int32 kind)    { return (XDAS_Int32)out_args_size(kind); }
XDAS_Int32 trikb200_sizeofHandle(void)                { return (XDAS_Int32)sizeof(TrikB200Handle); }

IVIDTRANSCODE_Handle trikb200_create(XDAS_Int32 kind, XDAS_Int32 width, XDAS_Int32 height,
                                     XDAS_Int32 lineLength, XDAS_Int32 outWidth, XDAS_Int32 outHeight)
{
  IVIDTRANSCODE_Fxns* fx = trikb200_fxns(kind);
  if (!fx)
  {
    set_error("unknown sensor kind");
    return NULL;
  }
  if (outWidth <= 0) outWidth = width;
  if (outHeight <= 0) outHeight = height;
  if (lineLength <= 0) lineLength = kind_is_planar(kind) ? width : 2 * width;

  TRIK_VIDTRANSCODE_CV_Params params;
  memset(&params, 0, sizeof(params));
  params.base.size = sizeof(params);
  params.base.numOutputStreams = 1;
  params.base.formatInput = default_input_format(kind);
  params.base.formatOutput[0] = TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_RGB565X;
  params.base.maxHeightInput = height > 480 ? height : 480;
  params.base.maxWidthInput = width > 640 ? width : 640;
  params.base.maxFrameRateInput = 60000;
  params.base.maxBitRateInput = -1;
  int outMax = outWidth > outHeight ? outWidth : outHeight;
  if (outMax < 640) outMax = 640;
  params.base.maxHeightOutput[0] = outMax; params.base.maxHeightOutput[1] = -1;
  params.base.maxWidthOutput[0] = outMax;  params.base.maxWidthOutput[1] = -1;
  params.base.maxFrameRateOutput[0] = params.base.maxFrameRateOutput[1] = -1;
  params.base.maxBitRateOutput[0] = params.base.maxBitRateOutput[1] = -1;
  params.base.dataEndianness = XDM_BYTE;

  IALG_MemRec memTab[IALG_DEFMEMRECS];
  memset(memTab, 0, sizeof(memTab));
  IALG_Fxns* parent = NULL;
  const int n = fx->ialg.algAlloc(reinterpret_cast<const IALG_Params*>(&params), &parent, memTab);
  for (int i = 0; i < n; ++i)
  {
    memTab[i].base = calloc(1, memTab[i].size ? memTab[i].size : 1);
    if (!memTab[i].base)
      return NULL;
  }
  IALG_Handle alg = reinterpret_cast<IALG_Handle>(memTab[0].base);
  alg->fxns = &fx->ialg;
  bool ok = fx->ialg.algInit(alg, memTab, NULL, reinterpret_cast<const IALG_Params*>(&params)) == IALG_EOK;
  if (ok)
  {
    TrikB200Handle* h = reinterpret_cast<TrikB200Handle*>(alg);
    TRIK_VIDTRANSCODE_CV_DynamicParams dyn = h->dynamicParams;   // the defaults initObj installed
    dyn.base.outputHeight[0] = outHeight;
    dyn.base.outputWidth[0] = outWidth;
    dyn.inputHeight = height;
    dyn.inputWidth = width;
    dyn.inputLineLength = lineLength;
    dyn.outputLineLength[0] = outWidth * 2;
    IVIDTRANSCODE_Status status;
    memset(&status, 0, sizeof(status));
    status.size = sizeof(status);
    ok = fx->control(reinterpret_cast<IVIDTRANSCODE_Handle>(alg), XDM_SETPARAMS,
                     reinterpret_cast<IVIDTRANSCODE_DynamicParams*>(&dyn), &status) == IVIDTRANSCODE_EOK;
    if (!ok)
      set_error("control(XDM_SETPARAMS) rejected the geometry");
  }
  if (!ok)
  {
    trikb200_delete(reinterpret_cast<IVIDTRANSCODE_Handle>(alg));
    return NULL;
  }
  return reinterpret_cast<IVIDTRANSCODE_Handle>(alg);
}

void trikb200_delete(IVIDTRANSCODE_Handle handle)
{
  if (!handle)
    return;
  IALG_MemRec memTab[IALG_DEFMEMRECS];
  memset(memTab, 0, sizeof(memTab));
  const int n = alg_free(reinterpret_cast<IALG_Handle>(handle), memTab);
  for (int i = n - 1; i >= 0; --i)
    free(memTab[i].base);
}

XDAS_Int32 trikb200_processBatch(IVIDTRANSCODE_Handle handle, const TRIKB200_Batch* batch)
{
  TrikB200Handle* h = reinterpret_cast<TrikB200Handle*>(handle);
  if (!h || !batch || batch->size != (XDAS_Int32)sizeof(TRIKB200_Batch))
  {
    set_error("bad handle or TRIKB200_Batch.size");
    return IVIDTRANSCODE_EFAIL;
  }
  Instance* in = instance_of(h);
  if (!in || batch->numFrames < 0 || (batch->numFrames > 0 && (!batch->frames || !batch->inArgsAlg || !batch->outArgsAlg)))
  {
    set_error("null batch pointers");
    return IVIDTRANSCODE_EFAIL;
  }
  if (batch->numFrames == 0)
    return IVIDTRANSCODE_EOK;
  const size_t fbytes = in->frame_bytes();
  if (batch->numFrames > 1 && (batch->frameStride < (int64_t)fbytes || (batch->frameStride & 15)))
  {
    set_error("frameStride must cover a frame and be a multiple of 16");
    return IVIDTRANSCODE_EFAIL;
  }
  if (batch->framesMem == TRIKB200_MEM_DEVICE && (((uintptr_t)batch->frames & 15) || (in->geo.lineLength & 15)))
  {
    set_error("device frames must be 16-byte aligned with a 16-byte multiple lineLength");
    return IVIDTRANSCODE_EFAIL;
  }
  if ((in->geo.lineLength & 15) != 0)
  {
    set_error("inputLineLength must be a multiple of 16");
    return IVIDTRANSCODE_EFAIL;
  }
  if (batch->outArgsStride < (XDAS_Int32)out_args_alg_size(h->kind)
      || (batch->inArgsStride != 0 && batch->inArgsStride < (XDAS_Int32)in_args_alg_size(h->kind)))
  {
    set_error("inArgsStride / outArgsStride smaller than the sensor's structs");
    return IVIDTRANSCODE_EFAIL;
  }
  BatchView b{};
  b.n = batch->numFrames;
  b.frames = reinterpret_cast<const uint8_t*>(batch->frames);
  b.frameStride = batch->frameStride;
  b.framesOnDevice = (batch->framesMem == TRIKB200_MEM_DEVICE);
  b.inArgs = reinterpret_cast<const uint8_t*>(batch->inArgsAlg);
  b.inStride = batch->inArgsStride;
  b.outArgs = reinterpret_cast<uint8_t*>(batch->outArgsAlg);
  b.outStride = batch->outArgsStride;
  b.outOnDevice = (batch->outArgsMem == TRIKB200_MEM_DEVICE);
  b.seeds = batch->seeds; b.seedsBroadcast = false;
  b.stream = reinterpret_cast<cudaStream_t>(batch->stream);
  b.async = (batch->flags & TRIKB200_BATCH_ASYNC) != 0;
  b.deviceTail = (batch->flags & TRIKB200_BATCH_DEVICE_TAIL) != 0;
  if (batch->previews)
  {
    const long long need = (long long)in->outHeight * in->outLineLength;
    if (batch->numFrames > 1 && batch->previewStride < need)
    {
      set_error("previewStride smaller than outputHeight * outputLineLength");
      return IVIDTRANSCODE_EFAIL;
    }
    b.previews = reinterpret_cast<uint8_t*>(batch->previews);
    b.previewStride = batch->previewStride > 0 ? batch->previewStride : need;
    b.previewsOnDevice = (batch->previewsMem == TRIKB200_MEM_DEVICE);
  }
  if (batch->streamIds)
  {
    if (batch->numStreams <= 0)
    {
      set_error("streamIds given but numStreams <= 0");
      return IVIDTRANSCODE_EFAIL;
    }
    for (int i = 0; i < batch->numFrames; ++i)
      if (batch->streamIds[i] < 0 || batch->streamIds[i] >= batch->numStreams)
      {
        set_error("streamIds entry out of range");
        return IVIDTRANSCODE_EFAIL;
      }
    b.streamIds = batch->streamIds;
    b.numStreams = batch->numStreams;
  }
  return run_batch(in, b) ? IVIDTRANSCODE_EOK : IVIDTRANSCODE_EFAIL;
}

// frames shard by batch across the GPUs of one box -- one host thread and one stream set per GPU, no
// collective on the per-pixel path, the results land in the caller's one array.
XDAS_Int32 trikb200_processBatchMulti(const IVIDTRANSCODE_Handle* handles, XDAS_Int32 numHandles, const TRIKB200_Batch* batch)
{
  if (!handles || numHandles < 1 || !batch || batch->size != (XDAS_Int32)sizeof(TRIKB200_Batch))
  {
    set_error("bad handle table or TRIKB200_Batch.size");
    return IVIDTRANSCODE_EFAIL;
  }
  if (numHandles == 1)
    return trikb200_processBatch(handles[0], batch);
  std::vector<Instance*> ins((size_t)numHandles);
  for (int d = 0; d < numHandles; ++d)
  {
    TrikB200Handle* h = reinterpret_cast<TrikB200Handle*>(handles[d]);
    Instance* in = h ? instance_of(h) : nullptr;
    const Instance* l = ins[0];
    if (!in || !in->valid
        || (d > 0 && (in == l || in->kind != l->kind || in->geo.width != l->geo.width || in->geo.height != l->geo.height
                      || in->geo.lineLength != l->geo.lineLength || in->outWidth != l->outWidth
                      || in->outHeight != l->outHeight || in->outLineLength != l->outLineLength)))
    {
      set_error("processBatchMulti: the handles must be distinct instances of one sensor kind and geometry");
      return IVIDTRANSCODE_EFAIL;
    }
    for (int e = 0; e < d; ++e)
      if (ins[(size_t)e] == in)
      {
        set_error("processBatchMulti: the same handle given twice");
        return IVIDTRANSCODE_EFAIL;
      }
    ins[(size_t)d] = in;
  }
  if (batch->framesMem != TRIKB200_MEM_HOST || batch->outArgsMem != TRIKB200_MEM_HOST
      || (batch->previews && batch->previewsMem != TRIKB200_MEM_HOST) || batch->stream || batch->streamIds
      || (batch->flags & TRIKB200_BATCH_ASYNC))
  {
    set_error("processBatchMulti: host frames, results and previews only; synchronous; no caller stream, no streamIds");
    return IVIDTRANSCODE_EFAIL;
  }
  const int n = batch->numFrames;
  if (n < 0 || (n > 0 && (!batch->frames || !batch->inArgsAlg || !batch->outArgsAlg)))
  {
    set_error("null batch pointers");
    return IVIDTRANSCODE_EFAIL;
  }
  if (n == 0)
    return IVIDTRANSCODE_EOK;
  const int kind = ins[0]->kind;
  if (batch->inArgsStride != 0 && batch->inArgsStride < (XDAS_Int32)in_args_alg_size(kind))
  {
    set_error("inArgsStride smaller than the sensor's struct");
    return IVIDTRANSCODE_EFAIL;
  }
  // contiguous frame ranges, the first n % D of them one frame longer
  std::vector<int> first((size_t)numHandles + 1, 0);
  for (int d = 0; d < numHandles; ++d)
    first[(size_t)d + 1] = first[(size_t)d] + n / numHandles + (d < n % numHandles ? 1 : 0);
  // Carried state (the ov7670 line sensor's lagging band, the object sensor's persisting range) is a function of the
  // arguments alone: walk it over the whole batch on the host so that every range starts from the state the frame before
  // it leaves behind -- the results are those of n sequential process() calls on handles[0].
  const bool carries = kind == KIND_OL || kind == KIND_OO;
  CarriedState walk = ins[0]->state;
  if (carries)
    for (int d = 0; d < numHandles; ++d)
    {
      ins[(size_t)d]->state = walk;
      FrameParams scratch;
      for (int i = first[(size_t)d]; i < first[(size_t)d + 1]; ++i)
        prepare_frame_params(kind, ins[0]->geo, reinterpret_cast<const uint8_t*>(batch->inArgsAlg) + (size_t)i * batch->inArgsStride,
                             walk, scratch);
    }
  std::vector<XDAS_Int32> rets((size_t)numHandles, IVIDTRANSCODE_EOK);
  std::vector<std::string> errors((size_t)numHandles);
  auto work = [&](int d)
  {
    const int lo = first[(size_t)d], cnt = first[(size_t)d + 1] - lo;
    if (cnt <= 0)
      return;
    TRIKB200_Batch sub = *batch;
    sub.numFrames = cnt;
    sub.frames = reinterpret_cast<const uint8_t*>(batch->frames) + (int64_t)lo * batch->frameStride;
    sub.inArgsAlg = reinterpret_cast<const uint8_t*>(batch->inArgsAlg) + (size_t)lo * batch->inArgsStride;
    sub.outArgsAlg = reinterpret_cast<uint8_t*>(batch->outArgsAlg) + (size_t)lo * batch->outArgsStride;
    if (batch->seeds) sub.seeds = batch->seeds + lo;
    if (batch->previews) sub.previews = reinterpret_cast<uint8_t*>(batch->previews) + (int64_t)lo * batch->previewStride;
    rets[(size_t)d] = trikb200_processBatch(handles[d], &sub);
    if (rets[(size_t)d] != IVIDTRANSCODE_EOK)
      errors[(size_t)d] = t_lastError;
  };
  std::vector<std::thread> pool;
  for (int d = 1; d < numHandles; ++d)
    pool.emplace_back(work, d);
  work(0);
  for (auto& th : pool)
    th.join();
  if (carries)
    for (int d = 0; d < numHandles; ++d)
      ins[(size_t)d]->state = walk;
  for (int d = 0; d < numHandles; ++d)
    if (rets[(size_t)d] != IVIDTRANSCODE_EOK)
    {
      t_lastError = errors[(size_t)d];
      return IVIDTRANSCODE_EFAIL;
    }
  return IVIDTRANSCODE_EOK;
}

XDAS_Int32 trikb200_processMixed(const TRIKB200_MixedEntry* entries, XDAS_Int32 numEntries)
{
  if (numEntries < 0 || (numEntries > 0 && !entries))
  {
    set_error("null entries");
    return IVIDTRANSCODE_EFAIL;
  }
  // Group by (sensor kind, device, geometry), not by handle: the reference's model is one codec instance per camera, so a
  // rig of 1024 cameras is 1024 handles, and all handles of one class go through ONE launch per kernel.  A class borrows
  // the workspace and stream of its first handle; each frame is judged with (and advances) the carried state of ITS OWN
  // handle, in entry order, so the results are those of per-handle process() calls.
  struct Group {
    Instance* lead;
    std::vector<const uint8_t*> frames, ins;
    std::vector<uint8_t*> outs;
    std::vector<int64_t> seeds;
    std::vector<CarriedState*> states;
    BatchView view{};
    Pending pend;
    bool enqueued = false;
  };
  std::vector<Group> groups;
  Group* last = nullptr;
  for (int i = 0; i < numEntries; ++i)
  {
    const TRIKB200_MixedEntry& e = entries[i];
    TrikB200Handle* h = reinterpret_cast<TrikB200Handle*>(e.handle);
    Instance* in = h ? instance_of(h) : nullptr;
    if (!in || !e.frame || !e.inArgsAlg || !e.outArgsAlg)
    {
      set_error("null pointer in a mixed entry");
      return IVIDTRANSCODE_EFAIL;
    }
    if (!in->valid || in->geo.width <= 0 || in->geo.height <= 0)
    {
      set_error("algorithm not set up for a non-empty image");
      return IVIDTRANSCODE_EFAIL;
    }
    auto same_class = [in](const Group& g) {
      const Instance* l = g.lead;
      return l == in || (l->kind == in->kind && l->device == in->device && l->geo.width == in->geo.width
                         && l->geo.height == in->geo.height && l->geo.lineLength == in->geo.lineLength);
    };
    Group* gptr = (last && same_class(*last)) ? last : nullptr;
    if (!gptr)
      for (Group& g : groups)
        if (same_class(g)) { gptr = &g; break; }
    if (!gptr)
    {
      groups.emplace_back();
      gptr = &groups.back();
      gptr->lead = in;
    }
    last = nullptr;                       // (groups may have been reallocated)
    gptr->frames.push_back(reinterpret_cast<const uint8_t*>(e.frame));
    gptr->ins.push_back(reinterpret_cast<const uint8_t*>(e.inArgsAlg));
    gptr->outs.push_back(reinterpret_cast<uint8_t*>(e.outArgsAlg));
    gptr->seeds.push_back(e.seed);
    gptr->states.push_back(&in->state);
    last = gptr;
  }
  // enqueue every class on its lead handle's stream, then finish them all: the classes overlap on the GPU
  bool ok = true;
  for (Group& g : groups)
  {
    BatchView& b = g.view;
    b.n = (int)g.frames.size();
    b.framePtrs = g.frames.data(); b.inPtrs = g.ins.data(); b.outPtrs = g.outs.data();
    b.statePtrs = g.states.data();
    b.framesOnDevice = false; b.outOnDevice = false;
    b.inStride = (int)in_args_alg_size(g.lead->kind); b.outStride = (int)out_args_alg_size(g.lead->kind);
    b.seeds = g.seeds.data(); b.seedsBroadcast = false;
    b.stream = nullptr; b.async = false;
    if ((g.lead->geo.lineLength & 15) != 0)
    {
      set_error("inputLineLength must be a multiple of 16");
      ok = false;
      break;
    }
    g.enqueued = enqueue_batch(g.lead, b, g.pend);
    if (!g.enqueued) { ok = false; break; }
  }
  for (Group& g : groups)
    if (g.enqueued && g.pend.needsFinish)
      ok = finish_batch(g.lead, g.view, g.pend) && ok;
  return ok ? IVIDTRANSCODE_EOK : IVIDTRANSCODE_EFAIL;
}

XDAS_Int32 trikb200_synchronize(IVIDTRANSCODE_Handle handle)
{
  TrikB200Handle* h = reinterpret_cast<TrikB200Handle*>(handle);
  Instance* in = h ? instance_of(h) : nullptr;
  if (!in)
    return IVIDTRANSCODE_EFAIL;
  cudaSetDevice(in->device);
  cudaError_t e = cudaStreamSynchronize(in->stream);
  if (e != cudaSuccess)
  {
    set_error("cudaStreamSynchronize", e);
    return IVIDTRANSCODE_EFAIL;
  }
  return IVIDTRANSCODE_EOK;
}

void trikb200_setSeed(IVIDTRANSCODE_Handle handle, int64_t seed)
{
  TrikB200Handle* h = reinterpret_cast<TrikB200Handle*>(handle);
  if (h && instance_of(h))
    instance_of(h)->seed = seed;
}

XDAS_Int32 trikb200_deviceCount(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
    return 0;
  return n;
}

XDAS_Int32 trikb200_setDevice(XDAS_Int32 device)
{
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess)
  {
    set_error("cudaSetDevice", e);
    return IVIDTRANSCODE_EFAIL;
  }
  t_device = device;
  return IVIDTRANSCODE_EOK;
}

int64_t trikb200_launchCount(void) { return launch_count(); }
void trikb200_setSlabsPerFrame(XDAS_Int32 slabs) { g_slabsPerFrame = slabs; }
void trikb200_setLoadStages(XDAS_Int32 stages) { set_sum_stages(stages); }
void trikb200_setBlockThreads(XDAS_Int32 threads) { set_target_threads(threads); }
void trikb200_setOverlapLaunch(XDAS_Int32 on) { set_overlap_launch(on); }
void trikb200_setLutMode(XDAS_Int32 mode) { g_lutMode = mode; }
void trikb200_setMxnTableMode(XDAS_Int32 mode) { g_mxnTableMode = mode; }
void trikb200_setGatherMode(XDAS_Int32 mode) { g_gatherMode = mode; }
void trikb200_setLutSkew(XDAS_Int32 on) { set_lut_skew(on); }
void trikb200_setLutParts(XDAS_Int32 parts) { set_lut_parts(parts); }
void trikb200_setPreviewChunkMB(XDAS_Int32 mb) { set_preview_chunk_bytes((long long)mb << 20); }
void trikb200_setPreviewSectorOverlay(XDAS_Int32 on) { set_preview_sector_overlay(on); }
void trikb200_setPreviewTable(XDAS_Int32 on) { set_preview_table(on); }
void trikb200_setEdgeLineVariant(XDAS_Int32 variant) { set_edge_variant(variant); }
void trikb200_setMxnTableThreads(XDAS_Int32 threads) { set_om_table_threads(threads); }
void trikb200_setZeroCopyBytes(XDAS_Int32 bytes) { g_zeroCopyBytes = bytes; }
void trikb200_setFramesPerCta(XDAS_Int32 n) { set_frames_per_cta(n); }
const char* trikb200_lastError(void) { return t_lastError.c_str(); }

/* test probes: exhaustive pixel functions straight from the device code (tests/test_pixel_gpu.py) */
XDAS_Int32 trikb200_probePixels(XDAS_Int32 which, uint32_t first, uint32_t count, uint32_t* hostOut)
{
  uint32_t* d = nullptr;
  if (cudaMalloc(&d, (size_t)count * 4) != cudaSuccess)
    return IVIDTRANSCODE_EFAIL;
  cudaError_t e;
  if (which == 3)
  {
    int dev = 0;
    cudaGetDevice(&dev);
    const uint16_t* binTable = ensure_om_table(dev, 0);
    if (!binTable) { cudaFree(d); return IVIDTRANSCODE_EFAIL; }
    e = launch_om_table_probe(binTable, first, count, d, 0);
  }
  else
    e = which == 0 ? launch_probe_yuv2rgb(first, count, d, 0)
      : (which == 1 ? launch_probe_rgb2hsv(first, count, d, 0) : launch_probe_yuv2hsv(first, count, d, 0));
  if (e == cudaSuccess)
    e = cudaMemcpy(hostOut, d, (size_t)count * 4, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess)
  {
    set_error("probe", e);
    return IVIDTRANSCODE_EFAIL;
  }
  return IVIDTRANSCODE_EOK;
}

XDAS_Int32 trikb200_ingestRgb565(const TRIKB200_Ingest* d)
{
  if (!d || d->size != (XDAS_Int32)sizeof(TRIKB200_Ingest) || d->numFrames < 0 || !d->src || !d->dst
      || (d->pixelFormat != TRIKB200_PIXEL_RGB565 && d->pixelFormat != TRIKB200_PIXEL_RGB565X))
  {
    set_error("ingestRgb565: bad descriptor");
    return IVIDTRANSCODE_EFAIL;
  }
  if (d->numFrames == 0)
    return IVIDTRANSCODE_EOK;
  cudaStream_t s = static_cast<cudaStream_t>(d->stream);
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t srcFrame = (size_t)d->srcLineLength * d->height, dstFrame = (size_t)2 * d->dstLineLength * d->height;
  const bool srcHost = d->srcMem != TRIKB200_MEM_DEVICE, dstHost = d->dstMem != TRIKB200_MEM_DEVICE;
  uint8_t *dSrc = nullptr, *dDst = nullptr;
  const uint8_t* src = static_cast<const uint8_t*>(d->src);
  uint8_t* dst = static_cast<uint8_t*>(d->dst);
  long long srcStride = d->srcStride, dstStride = d->dstStride;
  if (e == cudaSuccess && srcHost)
  {
    e = cudaMalloc(&dSrc, srcFrame * d->numFrames);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(dSrc, srcFrame, src, (size_t)d->srcStride, srcFrame, (size_t)d->numFrames, cudaMemcpyHostToDevice, s);
    src = dSrc; srcStride = (long long)srcFrame;
  }
  if (e == cudaSuccess && dstHost)
  {
    e = cudaMalloc(&dDst, dstFrame * d->numFrames);
    dst = dDst; dstStride = (long long)dstFrame;
  }
  if (e == cudaSuccess)
    e = launch_ingest_rgb565(src, srcStride, d->srcLineLength, dst, dstStride, d->dstLineLength, d->width, d->height,
                             d->numFrames, d->pixelFormat == TRIKB200_PIXEL_RGB565X ? 1 : 0, sms, s);
  if (e == cudaSuccess && dstHost)
    e = cudaMemcpy2DAsync(d->dst, (size_t)d->dstStride, dDst, dstFrame, dstFrame, (size_t)d->numFrames, cudaMemcpyDeviceToHost, s);
  if (srcHost || dstHost)
  {
    const cudaError_t e2 = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = e2;
    cudaFree(dSrc);
    cudaFree(dDst);
  }
  if (e != cudaSuccess)
  {
    cudaGetLastError();
    set_error("ingestRgb565", e);
    return IVIDTRANSCODE_EFAIL;
  }
  return IVIDTRANSCODE_EOK;
}

XDAS_Int32 trikb200_edgeLineBatch(const TRIKB200_EdgeLineBatch* d)
{
  if (!d || d->size != (XDAS_Int32)sizeof(TRIKB200_EdgeLineBatch) || d->numFrames < 0 || !d->frames || !d->outArgsAlg
      || d->outArgsStride < (XDAS_Int32)sizeof(TRIKB200_TargetOutArgsAlg))
  {
    set_error("edgeLineBatch: bad descriptor");
    return IVIDTRANSCODE_EFAIL;
  }
  if (d->numFrames == 0)
    return IVIDTRANSCODE_EOK;
  cudaStream_t s = static_cast<cudaStream_t>(d->stream);
  const bool srcHost = d->framesMem != TRIKB200_MEM_DEVICE, dstHost = d->outArgsMem != TRIKB200_MEM_DEVICE;
  const size_t plane = (size_t)d->lineLength * d->height;
  const uint8_t* frames = static_cast<const uint8_t*>(d->frames);
  long long frameStride = d->frameStride;
  uint8_t *dFrames = nullptr, *dOut = nullptr;
  uint8_t* out = static_cast<uint8_t*>(d->outArgsAlg);
  int outStride = d->outArgsStride;
  cudaError_t e = cudaSuccess;
  if (srcHost)
  {
    e = cudaMalloc(&dFrames, plane * d->numFrames);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(dFrames, plane, frames, (size_t)d->frameStride, plane, (size_t)d->numFrames, cudaMemcpyHostToDevice, s);
    frames = dFrames; frameStride = (long long)plane;
  }
  if (e == cudaSuccess && dstHost)
  {
    e = cudaMalloc(&dOut, sizeof(TargetOut) * (size_t)d->numFrames);
    out = dOut; outStride = (int)sizeof(TargetOut);
  }
  if (e == cudaSuccess)
    e = launch_edge_line(frames, frameStride, d->lineLength, d->width, d->height, d->numFrames, reinterpret_cast<TargetOut*>(out), outStride, s);
  if (e == cudaSuccess && dstHost)
    e = cudaMemcpy2DAsync(d->outArgsAlg, (size_t)d->outArgsStride, dOut, sizeof(TargetOut), sizeof(TargetOut), (size_t)d->numFrames,
                          cudaMemcpyDeviceToHost, s);
  if (srcHost || dstHost)
  {
    const cudaError_t e2 = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = e2;
    cudaFree(dFrames);
    cudaFree(dOut);
  }
  if (e != cudaSuccess)
  {
    cudaGetLastError();
    set_error("edgeLineBatch", e);
    return IVIDTRANSCODE_EFAIL;
  }
  return IVIDTRANSCODE_EOK;
}

XDAS_Int32 trikb200_probeLut(const TRIKB200_RangeInArgsAlg* inArgsAlg, uint64_t stats[5])
{
  Geometry g{};
  CarriedState st{};
  FrameParams fp;
  prepare_frame_params(KIND_WO, g, inArgsAlg, st, fp);
  uint8_t* table = nullptr; uint32_t* masks = nullptr; unsigned long long* dStats = nullptr;
  cudaError_t e = cudaMalloc(&table, LUT_TABLE_BYTES);
  if (e == cudaSuccess) e = cudaMalloc(&masks, LUT_MASK_BYTES);
  if (e == cudaSuccess) e = cudaMalloc(&dStats, 5 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(dStats, 0, 5 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = launch_chroma_table(fp.from, fp.to, fp.expected, table, masks, 0);
  if (e == cudaSuccess) e = launch_lut_check(fp.from, fp.to, fp.expected, table, masks, dStats, 0);
  if (e == cudaSuccess) e = cudaMemcpy(stats, dStats, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(table); cudaFree(masks); cudaFree(dStats);
  if (e != cudaSuccess)
  {
    set_error("probeLut", e);
    return IVIDTRANSCODE_EFAIL;
  }
  return IVIDTRANSCODE_EOK;
}

} // extern "C"
