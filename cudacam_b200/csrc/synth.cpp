// synth.cpp -- deterministic synthetic BGR8 frames (host side), SURVEY.md 8(d).
//
// The reference reads frames from a webcam (src/io/webcam.cpp:65-83); there is no camera on a GPU box, so
// benches and tests use frames made from a counter-based integer hash (no RNG library, identical on every
// machine).  Three distributions:
//   0 "scene": smooth low-frequency background + filled rectangles / discs of random colour, 3x3 box
//              soften, +-4 hash noise  -> a few % edge pixels, weak chains, tens of reference hysteresis rounds
//   1 "noise": i.i.d. uniform bytes   -> worst case edge density
//   2 "steps": constant regions with 0<->255 steps -> exercises the Gaussian S%159==0 path and the
//              (unsigned char) wrap of the reference's NMS value (src/cvp/cannyEdgeD.cu:267)
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/b200canny.h"

namespace
{
inline uint64_t mix(uint64_t x)
{   // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline uint64_t h2(uint64_t seed, uint64_t a, uint64_t b) { return mix(seed ^ mix(a * 0x100000001B3ull + b)); }

void scene(uint64_t seed, int w, int h, uint8_t *out, size_t stride)
{
  std::vector<uint8_t> img((size_t)w * h * 3);
  // background: bilinear interpolation of a coarse hash lattice (cell 96 px), per channel
  const int cell = 96;
  const int gw = w / cell + 2, gh = h / cell + 2;
  std::vector<int> lat((size_t)gw * gh * 3);
  for (int cy = 0; cy < gh; ++cy)
    for (int cx = 0; cx < gw; ++cx)
      for (int c = 0; c < 3; ++c) lat[((size_t)cy * gw + cx) * 3 + c] = (int)(h2(seed, (uint64_t)cy * 4096 + cx, c) & 127) + 48;
  for (int y = 0; y < h; ++y) {
    const int cy = y / cell, fy = y % cell;
    for (int x = 0; x < w; ++x) {
      const int cx = x / cell, fx = x % cell;
      const int *l0 = &lat[((size_t)cy * gw + cx) * 3], *l1 = l0 + (size_t)gw * 3;
      for (int c = 0; c < 3; ++c) {
        const int top = l0[c] * (cell - fx) + l0[3 + c] * fx, bot = l1[c] * (cell - fx) + l1[3 + c] * fx;
        img[((size_t)y * w + x) * 3 + c] = (uint8_t)((top * (cell - fy) + bot * fy) / (cell * cell));
      }
    }
  }
  // shapes: ~60 per megapixel, at least 6
  const int nshapes = std::max(6, (int)((double)w * h * 60.0 / 1.0e6));
  for (int s = 0; s < nshapes; ++s) {
    const uint64_t r = h2(seed, 0xABCDEF, s), r2 = h2(seed, 0x123457, s);
    const int cx = (int)(r % (uint64_t)w), cy = (int)((r >> 20) % (uint64_t)h);
    const int maxr = std::max(4, std::min(std::min(w, h) / 4, 90));
    const int rx = 3 + (int)((r >> 40) % (uint64_t)maxr), ry = 3 + (int)((r >> 52) % (uint64_t)maxr);
    const bool disc = (r2 & 1) != 0;
    // contrast against the background varies: some shapes are faint (weak edges), some strong
    const int amp = (r2 >> 1) & 3;   // 0 faint .. 3 strong
    uint8_t col[3];
    for (int c = 0; c < 3; ++c) {
      const int base = img[((size_t)cy * w + cx) * 3 + c];
      const int delta = (int)((r2 >> (8 + 8 * c)) & 0xFF) - 128;
      const int d = amp == 0 ? delta / 10 : amp == 1 ? delta / 5 : amp == 2 ? delta / 2 : delta;
      col[c] = (uint8_t)std::min(255, std::max(0, base + d));
    }
    const int x0 = std::max(0, cx - rx), x1 = std::min(w - 1, cx + rx), y0 = std::max(0, cy - ry), y1 = std::min(h - 1, cy + ry);
    for (int y = y0; y <= y1; ++y)
      for (int x = x0; x <= x1; ++x) {
        if (disc) {
          const long long dx = x - cx, dy = y - cy;
          if (dx * dx * ry * ry + dy * dy * rx * rx > (long long)rx * rx * ry * ry) continue;
        }
        uint8_t *p = &img[((size_t)y * w + x) * 3];
        p[0] = col[0]; p[1] = col[1]; p[2] = col[2];
      }
  }
  // 3x3 box soften (clamped at the border) + +-4 noise
  for (int y = 0; y < h; ++y) {
    const int ya = std::max(0, y - 1), yb = std::min(h - 1, y + 1);
    uint8_t *o = out + (size_t)y * stride;
    for (int x = 0; x < w; ++x) {
      const int xa = std::max(0, x - 1), xb = std::min(w - 1, x + 1);
      const uint64_t nz = h2(seed, 0x5EED0000ull + (uint64_t)y, x);
      for (int c = 0; c < 3; ++c) {
        int acc = 0;
        const int ys[3] = { ya, y, yb }, xs[3] = { xa, x, xb };
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 3; ++i) acc += img[((size_t)ys[j] * w + xs[i]) * 3 + c];
        const int n = (int)((nz >> (10 * c)) % 9) - 4;
        o[3 * x + c] = (uint8_t)std::min(255, std::max(0, acc / 9 + n));
      }
    }
  }
}

void noise(uint64_t seed, int w, int h, uint8_t *out, size_t stride)
{
  for (int y = 0; y < h; ++y) {
    uint8_t *o = out + (size_t)y * stride;
    for (int x = 0; x < w; ++x) {
      const uint64_t r = h2(seed, y, x);
      o[3 * x] = (uint8_t)r;
      o[3 * x + 1] = (uint8_t)(r >> 8);
      o[3 * x + 2] = (uint8_t)(r >> 16);
    }
  }
}

void steps(uint64_t seed, int w, int h, uint8_t *out, size_t stride)
{
  // checker of random-sized constant blocks; levels from {0, 255, grey levels}
  const int cell = 24 + (int)(mix(seed) % 40);
  for (int y = 0; y < h; ++y) {
    uint8_t *o = out + (size_t)y * stride;
    for (int x = 0; x < w; ++x) {
      const uint64_t r = h2(seed, (uint64_t)(y / cell), (uint64_t)(x / cell));
      const int k = (int)(r % 5);
      const uint8_t v = k == 0 ? 0 : k == 1 ? 255 : (uint8_t)(r >> 8);
      const bool grey = ((r >> 40) & 3) != 0;
      o[3 * x] = v;
      o[3 * x + 1] = grey ? v : (uint8_t)(r >> 16);
      o[3 * x + 2] = grey ? v : (uint8_t)(r >> 24);
    }
  }
}
}// namespace

extern "C" int b2c_synth_frame(int kind, uint64_t seed, int w, int h, uint8_t *out, size_t row_stride)
{
  if (!out || w < 1 || h < 1 || row_stride < (size_t)w * 3) return B2C_ERR_INVALID;
  switch (kind) {
  case 0: scene(seed, w, h, out, row_stride); return B2C_OK;
  case 1: noise(seed, w, h, out, row_stride); return B2C_OK;
  case 2: steps(seed, w, h, out, row_stride); return B2C_OK;
  default: return B2C_ERR_INVALID;
  }
}
