// Exact integer form of the reference's segment intersection test (is_intersect,
// environment/dummy_env_rectangular_pin.py:687-739) -- host + device, no CUDA types, so that
// tests/test_cross_int.py can run it on the CPU against the oracle's float64 restatement.
#ifndef PCB_CROSS_H
#define PCB_CROSS_H
#if defined(__CUDACC__)
#define PCB_CROSS_HD __host__ __device__
#else
#define PCB_CROSS_HD
#endif

namespace pcb {

constexpr int kMaxExactScale = 16;  // pins per net up to which a pin -> centroid segment has an integer record
constexpr int kSegExact = 0x80;

// Scale byte of a segment whose end points are stored multiplied by `scale` (the pins of the
// net for a pin -> centroid segment, whose centroid is (sum / scale); 1 for a pin -> pin
// segment): 0 = no integer test (the coordinates would not fit 16 bits), kSegExact set when
// the end points are exact in float64 (scale 1, 2, 4, 8, 16: integers, halves, quarters ...).
PCB_CROSS_HD inline int segment_scale_byte(int scale) {
  if (scale < 1 || scale > kMaxExactScale) return 0;
  return scale | ((scale & (scale - 1)) == 0 ? kSegExact : 0);
}

// Segment 1 = (ax1, ay1) -> (bx1, by1), every coordinate multiplied by c1 = cs & 0x7f;
// segment 2 likewise with ct.  Returns 0 = the reference says no, 1 = yes, 2 = undecided
// (the float64 formula has to be evaluated).
//
// With o1, o2 = orientation of the end points of segment 2 against the line of segment 1 and
// o3, o4 = the reverse (integers: every coordinate is a multiple of 1 / scale),
//  * end points that are exact in float64: every product and difference of the reference's
//    formula is exact (|values| < 2^13 with at most 8 fractional bits), the two quotients are
//    correctly rounded and compared with representable bounds, so its result is the exact
//    closed test -- shared end point, or lines not parallel and the crossing point on both
//    closed segments (o1 o2 <= 0 and o3 o4 <= 0);
//  * otherwise (a centroid that is a rounded third, fifth, ...) the reference's result can
//    differ from exact geometry only where the geometry is degenerate: an end point on the
//    other line (some o == 0, which includes shared end points and parallel lines on one line),
//    or a proper crossing on an axis-parallel segment, whose bounding box has no interior for
//    the rounded quotient to fall into.  Those are left to the float64 formula.  A proper
//    crossing of two oblique segments and a strict separation keep a margin of at least
//    1 / (16^2 |D|) (D = the integer determinant) against a rounding error below
//    1e-9 * 16^2 / |D| on boards of at most 64 x 64.
PCB_CROSS_HD inline int segments_cross_int(int ax1, int ay1, int bx1, int by1, int cs, int ax2, int ay2,
                                     int bx2, int by2, int ct) {
  const int c1 = cs & 0x7f, c2 = ct & 0x7f;
  const int d1x = bx1 - ax1, d1y = by1 - ay1, d2x = bx2 - ax2, d2y = by2 - ay2;
  const int wx = ax2 * c1 - ax1 * c2, wy = ay2 * c1 - ay1 * c2;  // A2 - A1, times c1 c2
  const int det = d1x * d2y - d1y * d2x;
  const int o1 = d1x * wy - d1y * wx, o2 = o1 + c1 * det;
  const int o3 = d2y * wx - d2x * wy, o4 = o3 - c2 * det;
  int r;
  if ((cs & ct & kSegExact) != 0) {
    const bool closed = det != 0 && ((o1 ^ o2) < 0 || o1 == 0 || o2 == 0) &&
                        ((o3 ^ o4) < 0 || o3 == 0 || o4 == 0);
    r = closed ? 1 : 0;
    if (det == 0) {  // parallel: only a shared end point counts (first test of is_intersect)
      const int ex = wx + d2x * c1, ey = wy + d2y * c1;  // B2 - A1
      const int fx = d1x * c2, fy = d1y * c2;            // B1 - A1
      if ((wx | wy) == 0 || (ex | ey) == 0 || ((wx - fx) | (wy - fy)) == 0 ||
          ((ex - fx) | (ey - fy)) == 0)
        r = 1;
    }
  } else {
    const bool any_zero = o1 == 0 || o2 == 0 || o3 == 0 || o4 == 0;
    const bool strict = ((o1 ^ o2) & (o3 ^ o4)) < 0;  // both pairs of opposite sign
    const bool axis = d1x == 0 || d1y == 0 || d2x == 0 || d2y == 0;
    r = any_zero ? 2 : (strict ? (axis ? 2 : 1) : 0);
  }
  return (c1 == 0 || c2 == 0) ? 2 : r;
}

}  // namespace pcb
#endif
