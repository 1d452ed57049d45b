// trik_kernels.cuh -- device-side data layout and launcher declarations (internal to libtrikb200).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace trikb200 {

enum Kind : int { KIND_WO = 0, KIND_WL = 1, KIND_OO = 2, KIND_OL = 3, KIND_OM = 4, KIND_COUNT = 5 };

__host__ __device__ inline bool kind_is_planar(int kind) { return kind == KIND_OO || kind == KIND_OL || kind == KIND_OM; }

// Geometry of one sensor instance (what CVAlgorithm::setup() receives).
struct Geometry {
  int32_t width, height;      // pixels; width % 32 == 0, height % 4 == 0
  int32_t lineLength;         // byte stride of a row (of each plane for YUV422P)
  int64_t frameStride;        // bytes between frames of a batch
  void*   drawInfo;           // DrawInfo[numFrames] for the preview overlays, or NULL when no preview is wanted
};

// What the overlay pass needs from the result tails, in SOURCE image coordinates (the OutArgs hold the
// same quantities only after the lossy scaling to -100..100):
//   WO      v[0] = points > 0,  v[1] = targetX, v[2] = targetY, v[3] = radius
//   WL, OL  v[0] = points > 10, v[1] = targetX
//   OO      v[0] = bit i set if target i is reported, v[1+2i], v[2+2i] = its x, y
struct DrawInfo { int32_t v[20]; };

// Per-frame parameters, prepared on the host from InArgsAlg + the handle's carried state.
struct FrameParams {
  // line sensors (V-only threshold in the key domain, see trik_pixel.cuh)
  uint32_t negKlo2;           // both lanes: (-klo) mod 2^16
  uint32_t n2;                // both lanes: N = khi - klo;  pass <=> ((key - klo) mod 2^16) <= N
  uint32_t hStart, hStop;     // OL cross band of THIS frame (lags one frame, OL/.../cv_line_detector_seqpass.hpp:449-450)
  // full HSV threshold (WO, OO): packed 0x00VVSSHH bounds and expected compare mask
  uint32_t from, to, expected;
  // OM grid
  uint32_t gridRows, gridCols; // m_heightM (= inArgs.widthM), m_widthN (= inArgs.heightN)
  uint32_t flags;             // FP_* bits
  uint32_t pad[2];
};
constexpr uint32_t FP_AUTODETECT = 1u;

// Raw per-frame accumulators of the sum kernels; zero before the first launch and reset to zero
// by the CTA that finalises the frame, so back-to-back launches need no memset.
struct SumAcc {
  uint32_t fails;             // pixels inside the counting window that FAIL the threshold
  uint32_t sxFail;            // sum of their columns
  uint32_t syFail;            // sum of their rows (WO)
  uint32_t crossFail;         // fails inside rows hStart..hStop (OL)
  uint32_t done;              // CTAs of this frame that have contributed
  uint32_t pad[3];
};

// Device result record of WO / WL / OL == TRIKB200_TargetOutArgsAlg (16 bytes).
struct TargetOut {
  int8_t  targetX, targetY;
  uint8_t targetSize, pad;
  uint16_t detectHue, detectHueTolerance, detectSat, detectSatTolerance, detectVal, detectValTolerance;
};
static_assert(sizeof(TargetOut) == 16, "TargetOutArgsAlg layout");

struct LaunchStats { long long launches; };

// ---- launchers (trik_kernels.cu) ---------------------------------------------------------------
// params: numFrames entries, or one entry when paramStride == 0.
cudaError_t launch_sum_sensor(int kind, const Geometry& g, int numFrames, const uint8_t* frames,
                              const FrameParams* params, int paramStride, SumAcc* acc, TargetOut* out,
                              int slabsPerFrame, cudaStream_t stream);

// OM: out = numFrames x int32[100] (only gridRows*gridCols entries per frame are written)
cudaError_t launch_om(const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                      int paramStride, const uint32_t* colorTable, int32_t* out, int maxGridRows, int maxGridCols,
                      cudaStream_t stream);
// OO: bitmaps = numFrames x (W/4)(H/4) uint16; clusters = numFrames x maxLabels x 12 bytes;
// equal = numFrames x maxLabels uint16; out = numFrames x 36-byte ObjOutArgsAlg records
// lutTable != nullptr: all frames share one threshold set and step 1 goes through its chroma table;
// lutSets != nullptr: the frames come under several sets, each with its table and its frame list (see LutSets below)
struct LutSets;
cudaError_t launch_oo(const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                      int paramStride, uint16_t* bitmaps, void* clusters, uint16_t* equal, int maxLabels,
                      void* out, int* labelCounts, cudaStream_t stream,
                      const uint8_t* lutTable = nullptr, const uint32_t* lutMasks = nullptr, int smCount = 0,
                      const int* lutFrameList = nullptr, const LutSets* lutSets = nullptr);
inline int oo_max_labels(int width, int height) { return ((width / 4) / 2 + 1) * ((height / 4) / 2 + 1) + 2; }

// auto-calibration histograms over the frames listed in frameIdx (device array of numFlagged indices)
cudaError_t launch_wo_detect(const Geometry& g, int numFlagged, const uint8_t* frames, const int* frameIdx,
                             TargetOut* out, cudaStream_t stream);
// results: numFlagged records of (bins + 2) int32: the +1/-2 histogram, the seed bin, the seed value
// (bins = 256 for WL/OL, 1024 = 32x32 (H>>3, S>>3) for OO)
cudaError_t launch_ordered_hist(int kind, const Geometry& g, int numFlagged, const uint8_t* frames, const int* frameIdx,
                                int32_t* results, cudaStream_t stream);

// RGB565X preview with overlays (trik_kernels_preview.cu); maps are device arrays prepared by the host
cudaError_t launch_preview(int kind, const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                           int paramStride, const uint16_t* bitmaps, const DrawInfo* draw, const int32_t* omColours,
                           int outW, int outH, int outLine, const int32_t* lastRow, const int32_t* lastCol,
                           const int32_t* hi2ho, const int32_t* wi2wo, uint8_t* previews, long long previewStride,
                           cudaStream_t stream, const uint8_t* lutTable = nullptr, const uint32_t* lutMasks = nullptr);

// exhaustive pixel-function probes for the parity tests: out[i] for i = blockIdx*blockDim+threadIdx
cudaError_t launch_probe_yuv2rgb(uint32_t first, uint32_t count, uint32_t* out, cudaStream_t stream);
cudaError_t launch_probe_rgb2hsv(uint32_t first, uint32_t count, uint32_t* out, cudaStream_t stream);
cudaError_t launch_probe_yuv2hsv(uint32_t first, uint32_t count, uint32_t* out, cudaStream_t stream);

int sum_sensor_block_threads(int kind, int width);
void set_sum_stages(int stages);
void set_target_threads(int threads);
void set_overlap_launch(int on);
// annealing tail of the WL / OL / OO auto-calibration on the device (trik_kernels_anneal.cu)
cudaError_t launch_anneal(int kind, int numFlagged, const int* frameIdx, const int32_t* hist, const uint32_t* seeds,
                          void* out, cudaStream_t stream);
void set_frames_per_cta(int n);
// chroma-indexed detection table (trik_kernels_lut.cu): 2 x 65 536 bytes (+ the skewed copy) + 65 536 x 8 x uint32
constexpr size_t LUT_TABLE_BYTES = 2 * 65536 + 2 * 66560;    // plain image, then the skewed image (rows 260 bytes apart)
constexpr size_t LUT_MASK_BYTES  = (size_t)65536 * 8 * sizeof(uint32_t);
cudaError_t launch_chroma_table(uint32_t from, uint32_t to, uint32_t expected, uint8_t* table, uint32_t* masks,
                                cudaStream_t stream);
cudaError_t launch_oo_bitmap_lut(const Geometry& g, int numFrames, const uint8_t* frames, const uint8_t* table,
                                 const uint32_t* masks, uint16_t* bitmaps, int smCount, cudaStream_t stream);
cudaError_t launch_lut_check(uint32_t from, uint32_t to, uint32_t expected, const uint8_t* table, const uint32_t* masks,
                             unsigned long long* stats, cudaStream_t stream);
cudaError_t launch_wo_lut(const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                          const uint8_t* table, const uint32_t* masks, TargetOut* out, int smCount, cudaStream_t stream,
                          SumAcc* acc = nullptr, const int* frameList = nullptr);   // frameList: numFrames frame indices
void set_lut_parts(int parts);
void set_preview_chunk_bytes(long long bytes);
void set_preview_sector_overlay(int on);
void set_preview_table(int on);
// the same for a batch under several threshold sets, in one launch: set k has count[k] frames, their indices at
// frameList + listOffset[k], its FrameParams at params[paramIndex[k]] and its table / masks; the launcher fills the rest
constexpr int LUT_MAX_SETS = 8;
struct LutSets {
  int numSets;
  int ctaStart[LUT_MAX_SETS + 1];
  const uint8_t* table[LUT_MAX_SETS];
  const uint32_t* masks[LUT_MAX_SETS];
  int listOffset[LUT_MAX_SETS], count[LUT_MAX_SETS], paramIndex[LUT_MAX_SETS], parts[LUT_MAX_SETS], rowsPerPart[LUT_MAX_SETS];
};
cudaError_t launch_oo_bitmap_lut_sets(const Geometry& g, const uint8_t* frames, uint16_t* bitmaps, int smCount,
                                      cudaStream_t stream, const int* frameList, LutSets sets);
cudaError_t launch_wo_lut_sets(const Geometry& g, const uint8_t* frames, const FrameParams* params, TargetOut* out, int smCount,
                               cudaStream_t stream, SumAcc* acc, const int* frameList, LutSets sets);
cudaError_t launch_line_bulk(bool planar, const Geometry& g, long long grid, int threads, const uint8_t* frames,
                             const FrameParams* params, int paramStride, SumAcc* acc, TargetOut* out,
                             int slabs, int rowsPerSlab, int cpr, int rpi, int stages, bool overlap, cudaStream_t stream);
// OM through the full colour-bin table (trik_kernels_omtab.cu): 2^24 uint16 entries [U][V][Y] = bin * 4
constexpr size_t OM_TABLE_BYTES = ((size_t)1 << 24) * sizeof(uint16_t);
cudaError_t launch_om_bin_table(uint16_t* table, cudaStream_t stream);
cudaError_t launch_om_table_probe(const uint16_t* table, uint32_t first, uint32_t count, uint32_t* out, cudaStream_t stream);
// fbList != nullptr: only the cell rows listed there (items frame * maxGridRows + cellRow, *fbCount of them, written by
// launch_om_major on the same stream)
cudaError_t launch_om_table(const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                            int paramStride, const uint16_t* table, const uint32_t* colorTable, int32_t* out,
                            int maxGridRows, int maxGridCols, cudaStream_t stream,
                            const int* fbList = nullptr, const int* fbCount = nullptr, int smCount = 148);
// OM majority pass (trik_kernels_ommaj.cu): decides every cell in which one colour bin provably holds more than half of
// the pixels, and lists the cell rows with an undecided cell in fbList / fbCount for launch_om_table's list mode;
// *fbCount must be 0 at launch, *fbCountNext (the counter the NEXT batch will use) is zeroed by this launch
cudaError_t launch_om_major(const Geometry& g, int numFrames, const uint8_t* frames, const FrameParams* params,
                            int paramStride, const uint16_t* table, const uint32_t* colorTable, int32_t* out,
                            int maxGridRows, int* fbList, int* fbCount, int* fbCountNext, cudaStream_t stream);
void set_om_table_threads(int threads);
void set_lut_skew(int on);
// RGB565 -> YUV422P ingest front end (trik_kernels_ingest.cu)
cudaError_t launch_ingest_rgb565(const uint8_t* src, long long srcStride, int srcLine, uint8_t* dst, long long dstStride,
                                 int dstLine, int width, int height, int numFrames, int bgr, int smCount, cudaStream_t stream);
// scattered pinned host frames -> one device batch buffer (trik_kernels_ingest.cu); dSrcPtrs is a device array of the
// frames' device-visible addresses
cudaError_t launch_gather_frames(const uint8_t* const* dSrcPtrs, uint8_t* dst, long long dstStride, size_t frameBytes,
                                 int numFrames, cudaStream_t stream);
// ov7670/edge_line_sensor (trik_kernels_edge.cu): out record i at out + i * outStride bytes
cudaError_t launch_edge_line(const uint8_t* frames, long long frameStride, int lineLength, int width, int height,
                             int numFrames, TargetOut* out, int outStride, cudaStream_t stream);
void set_edge_variant(int v);
long long launch_count();

} // namespace trikb200
