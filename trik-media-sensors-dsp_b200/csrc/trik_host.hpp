// trik_host.hpp -- host-side pieces of the sensor pipeline that are not per-pixel work:
// argument scaling (InArgsAlg -> FrameParams), per-handle carried state, the annealing tail of the
// auto-calibration, and the private restatement of glibc's rand().
#pragma once
#include <cstdint>

#include "trik_b200.h"
#include "trik_kernels.cuh"

namespace trikb200 {

// State the reference carries between process() calls of one algorithm object:
//   OL  m_hStart / m_hStop are assigned after the pixel pass (ov7670/line_sensor/.../cv_line_detector_seqpass.hpp:449-450)
//       so a frame is judged with the previous frame's band; a fresh object is value-initialised
//       (new T(), src/vidtranscode_cv.cpp:58) -> 0,0.
//   OO  the packed HSV range persists while setHsvRange == 0 (cv_bitmap_builder_reference.hpp:110-130),
//       zero on a fresh object.
struct CarriedState {
  uint32_t olHStart = 0, olHStop = 0;
  uint32_t ooFrom = 0, ooTo = 0, ooExpected = 0;
};

// scaling of webcam/object_sensor/.../cv_ball_detector_seqpass.hpp:425-445
void pack_hsv_range(uint32_t hf, uint32_t ht, uint32_t sf, uint32_t st, uint32_t vf, uint32_t vt,
                    uint32_t& from, uint32_t& to, uint32_t& expected);

// Fill FrameParams of one frame and advance the carried state exactly as one run() would.
void prepare_frame_params(int kind, const Geometry& g, const void* inArgsAlg, CarriedState& st, FrameParams& fp);

// The 512 colours the mxn sensor can report: HSVtoRGB(h*8, s*64, v*64) for the 32x4x4 histogram bins
// (ov7670/mxn_sensor/include/internal/cv_ball_detector_seqpass.hpp:480-517), index = h<<4 | s<<2 | v.
void mxn_color_table(uint32_t table[512]);

// glibc TYPE_3 additive feedback generator == srand()/rand() of the reference's host build
// (stdlib/random_r.c).  Private state so that frames can be annealed concurrently.
struct GlibcRand {
  int32_t r[34];
  int f, b;
  void seed(unsigned s);
  int next();
};

// The annealing tails of the auto-calibration (host side; driven by GlibcRand and libm pow, the same
// two third-party functions the reference calls).  out = {hue, hueTol, sat, satTol, val, valTol}.
//   WL / OL  webcam/line_sensor/include/internal/cv_hsv_range_detector.hpp:242-303 (OL differs at :281)
//   OO       ov7670/object_sensor/include/internal/cv_hsv_range_detector.hpp:229-293
void anneal_line(const int32_t hist[256], int seedBin, bool isOL, unsigned seed, uint16_t out[6]);
void anneal_oo(const int32_t hist[1024], int seedBin, unsigned seed, uint16_t out[6]);

} // namespace trikb200
