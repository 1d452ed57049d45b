/*
 * TEST INFRASTRUCTURE (oracle/): OPEN RESTATEMENT of the three TI IMGLIB kernels that
 * trik/ov7670/edge_line_sensor calls (include/internal/cv_ball_detector_seqpass.hpp:13-16, :176-184, :224-229).
 *
 * IMGLIB itself (TI "C64x+ IMGLIB", imglib_c64Px) is a closed, un-vendored dependency of the reference: not in
 * /root/reference, no version pinned by the reference's makefiles beyond the include paths.  What follows restates the
 * "natural C" models that TI publishes for these kernels in the IMGLIB Programmer's Guide (SPRUF30: IMG_sobel_3x3_8,
 * IMG_thr_gt2max_8, IMG_ycbcr422pl_to_rgb565) -- written from that published behaviour, not copied from any source
 * file.  PARITY UNPINNED: nothing in the reference tests these kernels, and the optimised library builds are only
 * documented to match their C models.
 */
#include <stdlib.h>

/* 3x3 Sobel over the image taken as ONE raster line: output i + 1 from the 3x3 neighbourhood whose top-left input is i,
 * for i < cols * (rows - 2) - 2; |H| + |V| clamped to 255.  out[0] and everything from cols * (rows - 2) - 1 on are
 * not written (the two last rows of an in-place rows x cols buffer keep what they held). */
void IMG_sobel_3x3_8(const unsigned char* in, unsigned char* out, short cols, short rows)
{
  const int w = cols;
  int i;
  for (i = 0; i < cols * (rows - 2) - 2; ++i)
  {
    const int i00 = in[i], i01 = in[i + 1], i02 = in[i + 2];
    const int i10 = in[i + w], i12 = in[i + w + 2];
    const int i20 = in[i + 2 * w], i21 = in[i + 2 * w + 1], i22 = in[i + 2 * w + 2];
    const int H = -i00 - 2 * i01 - i02 + i20 + 2 * i21 + i22;
    const int V = -i00 + i02 - 2 * i10 + 2 * i12 - i20 + i22;
    int O = abs(H) + abs(V);
    if (O > 255) O = 255;
    out[i + 1] = (unsigned char)O;
  }
}

/* pixels above the threshold become 255, the others pass unchanged */
void IMG_thr_gt2max_8(const unsigned char* in_data, unsigned char* out_data, short cols, short rows, unsigned char threshold)
{
  const int pixels = rows * cols;
  int i;
  for (i = 0; i < pixels; ++i)
    out_data[i] = in_data[i] > threshold ? 255 : in_data[i];
}

/* planar Y / Cb / Cr 4:2:2 -> RGB565, Q13 coefficients {luma, r_cr, g_cb, g_cr, b_cb}; pixels in pairs sharing chroma */
void IMG_ycbcr422pl_to_rgb565(const short coeff[5], const unsigned char* y_data, const unsigned char* cb_data,
                              const unsigned char* cr_data, unsigned short* rgb_data, unsigned num_pixels)
{
  const int luma = coeff[0], r_cr = coeff[1], g_cb = coeff[2], g_cr = coeff[3], b_cb = coeff[4];
  unsigned i;
  for (i = 0; i < num_pixels / 2; ++i)
  {
    const int y0 = y_data[2 * i] - 16, y1 = y_data[2 * i + 1] - 16;
    const int cb = cb_data[i] - 128, cr = cr_data[i] - 128;
    const int y0t = luma * y0, y1t = luma * y1;
    const int rt = r_cr * cr, gt = g_cb * cb + g_cr * cr, bt = b_cb * cb;
    int r0 = (y0t + rt) >> 16, g0 = (y0t + gt) >> 15, b0 = (y0t + bt) >> 16;
    int r1 = (y1t + rt) >> 16, g1 = (y1t + gt) >> 15, b1 = (y1t + bt) >> 16;
    r0 = r0 < 0 ? 0 : (r0 > 31 ? 31 : r0); g0 = g0 < 0 ? 0 : (g0 > 63 ? 63 : g0); b0 = b0 < 0 ? 0 : (b0 > 31 ? 31 : b0);
    r1 = r1 < 0 ? 0 : (r1 > 31 ? 31 : r1); g1 = g1 < 0 ? 0 : (g1 > 63 ? 63 : g1); b1 = b1 < 0 ? 0 : (b1 > 31 ? 31 : b1);
    rgb_data[2 * i]     = (unsigned short)((r0 << 11) + (g0 << 5) + b0);
    rgb_data[2 * i + 1] = (unsigned short)((r1 << 11) + (g1 << 5) + b1);
  }
}
