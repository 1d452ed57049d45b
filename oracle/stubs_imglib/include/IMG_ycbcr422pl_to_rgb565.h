/* TEST INFRASTRUCTURE: declaration stub for the un-vendored TI IMGLIB header of the same name (see oracle/imglib_open.c). */
void IMG_ycbcr422pl_to_rgb565(const short coeff[5], const unsigned char* y_data, const unsigned char* cb_data,
                              const unsigned char* cr_data, unsigned short* rgb_data, unsigned num_pixels);
