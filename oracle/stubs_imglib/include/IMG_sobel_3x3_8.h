/* TEST INFRASTRUCTURE: declaration stub for the un-vendored TI IMGLIB header of the same name (see oracle/imglib_open.c). */
void IMG_sobel_3x3_8(const unsigned char* in, unsigned char* out, short cols, short rows);
