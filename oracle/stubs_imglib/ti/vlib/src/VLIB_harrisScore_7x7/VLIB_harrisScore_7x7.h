/* TEST INFRASTRUCTURE: empty stand-in for the un-vendored TI VLIB header; the reference only uses it under #ifdef CORNERS (off). */
