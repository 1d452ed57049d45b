/*
 * TEST INFRASTRUCTURE (oracle/): CPU restatement, in plain C, of the per-frame pixel pipeline of
 * trikset/trik-media-sensors-dsp (the five sensors of SURVEY.md section 8).  It is the checker that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use.  Nothing in the product
 * (trik-media-sensors-dsp_b200/, include/) includes, links or calls it.
 *
 * Parity pinning: the reference holds no golden vectors or tests.  This
 * restatement is pinned against the reference's own sources compiled for the host
 * (oracle/_ref/libtrikref_*.so, oracle/build_ref.sh) by tests/test_oracle_vs_ref.py -- exhaustively
 * for the two pixel functions (2^24 inputs each) and on seeded frames for every sensor -- and
 * against the committed fixtures in tests/golden/, which were generated from that host build.
 * The C6x intrinsic semantics in oracle/stubs/c6x.h remain the unpinned root of trust.
 *
 * Unlike the reference it has no 640x480 limit and no file-scope state, so it also covers the
 * larger sizes of BASELINE.json config 5.
 */
#ifndef TRIK_ORACLE_H_
#define TRIK_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  TRIK_ORACLE_WO = 0, /* trik/webcam/object_sensor  */
  TRIK_ORACLE_WL = 1, /* trik/webcam/line_sensor    */
  TRIK_ORACLE_OO = 2, /* trik/ov7670/object_sensor  */
  TRIK_ORACLE_OL = 3, /* trik/ov7670/line_sensor    */
  TRIK_ORACLE_OM = 4  /* trik/ov7670/mxn_sensor     */
};

/* InArgsAlg / OutArgsAlg layouts, identical to the reference's per-sensor headers. */
typedef struct {            /* WO, WL, OL: webcam/line_sensor/trik_vidtranscode_cv.h:48-56 */
  uint16_t detectHueFrom, detectHueTo;
  uint8_t  detectSatFrom, detectSatTo, detectValFrom, detectValTo;
  uint8_t  autoDetectHsv;
} trik_oracle_range_in;

typedef struct {            /* WO, WL, OL: webcam/line_sensor/trik_vidtranscode_cv.h:64-74 */
  int8_t   targetX, targetY;
  uint8_t  targetSize;
  uint16_t detectHue, detectHueTolerance, detectSat, detectSatTolerance, detectVal, detectValTolerance;
} trik_oracle_target_out;

typedef struct {            /* OO: ov7670/object_sensor/trik_vidtranscode_cv.h:51-60 */
  uint8_t  setHsvRange;
  uint16_t detectHue, detectHueTol;
  uint8_t  detectSat, detectSatTol, detectVal, detectValTol;
  uint8_t  autoDetectHsv;
} trik_oracle_obj_in;

typedef struct { int8_t x, y; uint8_t size; } trik_oracle_target;

typedef struct {            /* OO: ov7670/object_sensor/trik_vidtranscode_cv.h:67-81 */
  trik_oracle_target target[8];
  uint16_t detectHue, detectHueTolerance, detectSat, detectSatTolerance, detectVal, detectValTolerance;
} trik_oracle_obj_out;

typedef struct { int32_t widthM, heightN; } trik_oracle_mxn_in;     /* OM: mxn_sensor/trik_vidtranscode_cv.h:49-52 */
typedef struct { int32_t outColor[100]; } trik_oracle_mxn_out;      /* OM: mxn_sensor/trik_vidtranscode_cv.h:60-62 */

typedef struct trik_oracle_sensor trik_oracle_sensor;

/* pixel functions (section 8 rows a4, a5, a6) */
uint32_t trik_oracle_yuv_to_rgb888(uint32_t y, uint32_t u, uint32_t v);
uint32_t trik_oracle_rgb888_to_hsv(uint32_t rgb888);
/* bulk forms: index = Y | U<<8 | V<<16 -> 0x00RRGGBB ; index = 0x00RRGGBB -> 0x00VVSSHH */
void     trik_oracle_yuv_to_rgb888_range(uint32_t first, uint32_t count, uint32_t* out);
void     trik_oracle_rgb888_to_hsv_range(uint32_t first, uint32_t count, uint32_t* out);
int      trik_oracle_detect(uint32_t hsv, uint32_t range_from, uint32_t range_to, uint32_t expected);
uint32_t trik_oracle_hsv_to_rgb_mxn(int h, int s, int v);

/* the codec's algorithm object: create == CVAlgorithm::setup(), run == CVAlgorithm::run() */
trik_oracle_sensor* trik_oracle_create(int kind, int width, int height, int lineLength);
void trik_oracle_destroy(trik_oracle_sensor* s);
/* returns 1 on success, 0 when run() would return false (input smaller than height*lineLength) */
int trik_oracle_run(trik_oracle_sensor* s, const uint8_t* frame, int numBytes,
                    const void* inArgsAlg, void* outArgsAlg, long long seed);

/* Where the reference's behaviour is undefined the oracle picks a definition and says so:
 *  LINE_SEED_INDETERMINATE  WL/OL auto-detect: no centre-band bin ever became positive, so the
 *      reference starts the annealing from an uninitialised stack byte
 *      (WL/inc/cv_hsv_range_detector.hpp:84,229-233,252-253); the oracle starts from v = 0.
 *  OO_FEWER_THAN_8  fewer than 8 labels: the reference reads past the end of its cluster vector
 *      (OO/inc/cv_ball_detector_seqpass.hpp:575-589); the oracle treats those slots as empty.
 * Frames that raise a flag are compared product-vs-oracle only, never against oracle/_ref. */
#define TRIK_ORACLE_FLAG_LINE_SEED_INDETERMINATE 1
#define TRIK_ORACLE_FLAG_OO_FEWER_THAN_8         2
int trik_oracle_last_flags(const trik_oracle_sensor* s);

/* whole-image HSV (0x00VVSSHH per pixel) of the last run, for kernel debugging */
const uint32_t* trik_oracle_last_hsv(const trik_oracle_sensor* s);

/* ov7670/edge_line_sensor, restated: Sobel 3x3 of the luma plane through the open IMGLIB restatement
 * (oracle/imglib_open.c -- the real IMGLIB is closed and absent: PARITY UNPINNED for those kernels), threshold 50, then the
 * sensor's own counting loop and tail (include/internal/cv_ball_detector_seqpass.hpp:186-205, :369-420).
 * frame: luma plane of height rows, lineLength bytes apart; out: trik_oracle_target_out (targetX, targetY, targetSize). */
void trik_oracle_edge_line(const uint8_t* frame, int width, int height, int lineLength, void* outArgsAlg);

/* glibc TYPE_3 rand() restated (used by tests to pin the product's private generator) */
void trik_oracle_srand(uint32_t* state34, unsigned seed);
int  trik_oracle_rand(uint32_t* state34);

#ifdef __cplusplus
}
#endif
#endif
