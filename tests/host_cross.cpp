// Test harness (CPU): the integer pre-test of the crossing count, compiled from the product
// header with g++ and called through ctypes by tests/test_cross_int.py.
#include "pcb_cross.h"

extern "C" void pcb_test_cross_int(int n, const int* rec, int* out) {
  // rec[i] = ax1, ay1, bx1, by1, scale1, ax2, ay2, bx2, by2, scale2 (coordinates already scaled)
  for (int i = 0; i < n; ++i) {
    const int* r = rec + 10 * i;
    out[i] = pcb::segments_cross_int(r[0], r[1], r[2], r[3], pcb::segment_scale_byte(r[4]), r[5], r[6],
                                     r[7], r[8], pcb::segment_scale_byte(r[9]));
  }
}
