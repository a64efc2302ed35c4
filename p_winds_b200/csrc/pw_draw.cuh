// pw_draw.cuh -- host-side span table of the transit-geometry rasteriser
// (reference: transit.draw_transit, transit.py:87-116, which delegates to the absent
// third-party package `flatstar`; its disks are Pillow ImageDraw.ellipse fills -- pinned
// by the reference's golden transit depth, tests/test_transit.py:37).
#pragma once
#include <cstdlib>

namespace pw {

// Pillow's filled-ellipse rasteriser (src/libImaging/Draw.c: quarter_init / quarter_next
// / ellipse_next, Pillow >= 8.2): integer Bresenham-like walk over a quarter arc in
// half-pixel units; (int) truncation of the float bounding box as in _imaging.c.
// Writes the inclusive column span of every row into lo[] / hi[] (clipped to [0, n)).
inline void pil_ellipse_rows(double fx0, double fy0, double fx1, double fy1, int n, int* lo, int* hi) {
  for (int i = 0; i < n; ++i) { lo[i] = 1; hi[i] = 0; }
  const long long x0 = (long long)fx0, y0 = (long long)fy0, x1 = (long long)fx1, y1 = (long long)fy1;
  const long long a = x1 - x0, b = y1 - y0;
  if (a < 0 || b < 0) return;
  const long long a2 = a * a, b2 = b * b, a2b2 = a2 * b2;
  auto delta = [&](long long x, long long y) { return std::llabs(a2 * y * y + b2 * x * x - a2b2); };
  auto emit = [&](long long y, long long xr) {
    for (int sgn = 0; sgn < 2; ++sgn) {
      const long long yy = sgn ? -y : y;
      if (sgn && y == 0) break;
      const long long row = y0 + (yy + b) / 2;
      if (row < 0 || row >= n) continue;
      long long l = x0 + (a - xr) / 2, h = x0 + (a + xr) / 2;
      if (l < 0) l = 0;
      if (h > n - 1) h = n - 1;
      lo[row] = (int)l; hi[row] = (int)h;
    }
  };
  long long cx = a, cy = b % 2;
  const long long ex = a % 2, ey = b;
  long long last_row = cy;
  emit(cy, cx);
  while (!(cx == ex && cy == ey)) {
    long long nx = cx, ny = cy + 2;
    long long nd = delta(nx, ny);
    if (nx > 1) {
      const long long d2 = delta(cx - 2, cy + 2);
      if (nd > d2) { nx = cx - 2; ny = cy + 2; nd = d2; }
      const long long d3 = delta(cx - 2, cy);
      if (nd > d3) { nx = cx - 2; ny = cy; }
    }
    cx = nx; cy = ny;
    if (cy != last_row) { emit(cy, cx); last_row = cy; }   // first (widest) point of a new row
  }
}



}  // namespace pw
