/* TEST INFRASTRUCTURE: declaration stub for the un-vendored TI IMGLIB header of the same name (see oracle/imglib_open.c). */
void IMG_thr_gt2max_8(const unsigned char* in_data, unsigned char* out_data, short cols, short rows, unsigned char threshold);
