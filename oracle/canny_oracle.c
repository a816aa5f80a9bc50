/*
 * canny_oracle.c -- CPU restatement of the reference Canny path (axoloto/CudaCam src/cvp).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path
 * (cudacam_b200/csrc) never links or calls anything in oracle/.
 *
 * Parity status: PINNED.  The reference ships no golden vectors (its test/ dir never touches src/cvp),
 * so the pin is the reference's own CUDA kernels compiled unmodified from /root/reference/src/cvp/
 * cannyEdgeD.cu into oracle/_ref/ (see oracle/Makefile, oracle/ref_harness.cu) and run on a B200;
 * their outputs on the seeded inputs of tests/golden/ are committed there together with the
 * generating script (oracle/make_golden.py) and this file is checked against them stage by stage.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference/).
 * Plain scalar C; float arithmetic is written so that no contraction / reassociation can happen
 * (explicit fmaf, compiled with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* src/cvp/cannyEdgeD.cu:14-19 -- integer luma weights, 64*w + 0.5 truncated */
enum { B_WT = 7, G_WT = 38, R_WT = 19 };
/* src/cvp/cannyEdgeD.cu:28 */
#define GRAD_COEFF 4.0f
/* src/cvp/cannyEdgeD.cu:31-33 */
enum { FINAL_EDGE = 255, CANDIDATE_EDGE = 128, NO_EDGE = 0 };

/* src/cvp/cannyEdgeH.cu:372-379 -- GK = k * (1/159.0f), product rounded to fp32 on the host */
ORACLE_API void oracle_gauss_kernel(float gk[25])
{
  static const float k[25] = { 2, 4, 5, 4, 2, 4, 9, 12, 9, 4, 5, 12, 15, 12, 5, 4, 9, 12, 9, 4, 2, 4, 5, 4, 2 };
  const float inv = 1 / 159.0f;
  for (int i = 0; i < 25; ++i) {
    volatile float w = k[i] * inv; /* volatile: force the single fp32 rounding of the host multiply */
    gk[i] = w;
  }
}

/* src/cvp/cannyEdgeD.cu:53-69 -- byte 0 weighted as B, byte 1 as G, byte 2 as R; >>6; min(255,.) */
ORACLE_API void oracle_rgb2mono(const uint8_t *bgr, size_t stride, int w, int h, uint8_t *mono)
{
  for (int y = 0; y < h; ++y) {
    const uint8_t *p = bgr + (size_t)y * stride;
    for (int x = 0; x < w; ++x) {
      int v = (p[3 * x] * B_WT + p[3 * x + 1] * G_WT + p[3 * x + 2] * R_WT) >> 6;
      mono[(size_t)y * w + x] = (uint8_t)(v > 255 ? 255 : v);
    }
  }
}

/* src/cvp/cannyEdgeD.cu:72-118 -- zero outside the image (:91-98); fSum starts at 0 and takes 25
 * fused multiply-adds in r-major, c-minor order (:102-111; nvcc -O3 default -fmad=true contracts
 * `fSum += GK*v` to fma.rn.f32); result truncated to u8 (:115). */
__attribute__((target_clones("fma", "default")))
ORACLE_API void oracle_gaussian(const uint8_t *mono, int w, int h, uint8_t *blur)
{
  float gk[25];
  oracle_gauss_kernel(gk);
  for (int y = 0; y < h; ++y) {
    for (int x = 0; x < w; ++x) {
      float s = 0.0f;
      for (int r = 0; r < 5; ++r) {
        const int yy = y + r - 2;
        for (int c = 0; c < 5; ++c) {
          const int xx = x + c - 2;
          const float v = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? (float)mono[(size_t)yy * w + xx] : 0.0f;
          s = fmaf(gk[r * 5 + c], v, s);
        }
      }
      blur[(size_t)y * w + x] = (uint8_t)s; /* 0 <= s < 256 always */
    }
  }
}

static inline int px(const uint8_t *img, int w, int h, int y, int x)
{
  return (y >= 0 && y < h && x >= 0 && x < w) ? img[(size_t)y * w + x] : 0;
}

/* src/cvp/cannyEdgeD.cu:121-172 -- blur is zero outside the image (:142-149); integer sums (:158-167),
 * stored as float(sum)/8.0f (exact).  sumx/sumy (optional) receive the raw integer sums. */
ORACLE_API void oracle_sobel(const uint8_t *blur, int w, int h, float *sx, float *sy, int16_t *sumx, int16_t *sumy)
{
  for (int y = 0; y < h; ++y) {
    for (int x = 0; x < w; ++x) {
      const int a00 = px(blur, w, h, y - 1, x - 1), a01 = px(blur, w, h, y - 1, x), a02 = px(blur, w, h, y - 1, x + 1);
      const int a10 = px(blur, w, h, y, x - 1), a12 = px(blur, w, h, y, x + 1);
      const int a20 = px(blur, w, h, y + 1, x - 1), a21 = px(blur, w, h, y + 1, x), a22 = px(blur, w, h, y + 1, x + 1);
      const int gx = (-a00 + a02) + (-2 * a10 + 2 * a12) + (-a20 + a22);
      const int gy = (a00 + 2 * a01 + a02) - (a20 + 2 * a21 + a22);
      const size_t i = (size_t)y * w + x;
      if (sx) sx[i] = (float)gx / 8.0f;
      if (sy) sy[i] = (float)gy / 8.0f;
      if (sumx) sumx[i] = (int16_t)gx;
      if (sumy) sumy[i] = (int16_t)gy;
    }
  }
}

/* src/cvp/cannyEdgeD.cu:195 -- grad = 4 * sqrtf(sX*sX + sY*sY); nvcc contracts to
 * fma.rn(sX, sX, sY*sY); every intermediate is exact in fp32 (N = sumx^2+sumy^2 < 2^24, /64),
 * sqrt.rn is correctly rounded, *4 exact. */
__attribute__((target_clones("fma", "default")))
ORACLE_API void oracle_grad(const float *sx, const float *sy, size_t n, float *grad)
{
  for (size_t i = 0; i < n; ++i) {
    const float t = sy[i] * sy[i];
    grad[i] = GRAD_COEFF * sqrtf(fmaf(sx[i], sx[i], t));
  }
}

/* src/cvp/cannyEdgeD.cu:196 + :239-264 -- slope = atan2f(sX, sY) (note the argument order), angle in
 * degrees = slope*180/pi, +180 if negative, then 4 sectors with the boundaries of :245-264.
 * CUDA's atan2f is libdevice code, not IEEE, so the float value cannot be restated portably; the
 * *sector* is a function of the integer pair (sumx, sumy) in [-1020,1020]^2 only.  The rule below is
 * the exact-geometry sector (tan 22.5 deg = sqrt2-1 is irrational, so no lattice point lies on a
 * boundary); oracle/_ref runs the reference's own gradSlope over the whole 2041^2 domain on the GPU
 * and tests/ checks this table against it (tests/golden/sector_table.sha256).
 *   a = |sumx|, b = |sumy|
 *   a == 0 or (a+b)^2 < 2 b^2   -> sector 0   (angle < 22.5 or > 157.5)
 *   a > b and (a-b)^2 > 2 b^2   -> sector 2   (67.5 < angle <= 112.5)
 *   else sumx*sumy > 0          -> sector 1   (22.5 <= angle <= 67.5)
 *   else                        -> sector 3   (112.5 < angle <= 157.5)
 */
ORACLE_API int oracle_sector(int sumx, int sumy)
{
  const long a = labs((long)sumx), b = labs((long)sumy);
  if (a == 0 || (a + b) * (a + b) < 2 * b * b) return 0;
  if (a > b && (a - b) * (a - b) > 2 * b * b) return 2;
  return ((sumx > 0) == (sumy > 0)) ? 1 : 3;
}

/* Whole finite domain of the gradient stage: for every (sumx, sumy) in [-1020,1020]^2 (row = sumy+1020,
 * col = sumx+1020) the sector and the NMS byte value (unsigned char)grad of a kept pixel
 * (src/cvp/cannyEdgeD.cu:195,267).  tests/ compares both tables with what the reference's own gradSlope /
 * nonMaxSuppr kernels produce on the GPU (tests/golden/tables.json holds the hashes of that run). */
ORACLE_API void oracle_domain_tables(uint8_t *sector_out, uint8_t *nmsval_out)
{
  for (int j = 0; j < 2041; ++j)
    for (int i = 0; i < 2041; ++i) {
      const int gx = i - 1020, gy = j - 1020;
      const float sx = (float)gx / 8.0f, sy = (float)gy / 8.0f;
      float g;
      oracle_grad(&sx, &sy, 1, &g);
      if (sector_out) sector_out[(size_t)j * 2041 + i] = (uint8_t)oracle_sector(gx, gy);
      if (nmsval_out) nmsval_out[(size_t)j * 2041 + i] = (uint8_t)((uint32_t)g & 0xFFu);
    }
}

/* Best-effort float slope (correctly rounded double atan2 -> float).  Informational only: CUDA's
 * atan2f may differ by an ulp or two; parity is asserted on sectors, never on these bits. */
ORACLE_API void oracle_slope(const float *sx, const float *sy, size_t n, float *slope)
{
  for (size_t i = 0; i < n; ++i) slope[i] = (float)atan2((double)sx[i], (double)sy[i]);
}

/* src/cvp/cannyEdgeD.cu:201-270 -- grad is zero outside the image (:222-229); neighbours per sector
 * (:245-264); keep iff q <= g && r <= g (ties kept, :267); value = (unsigned char)g, which nvcc 12.9
 * compiles to cvt.rzi.u32.f32 + byte store, i.e. trunc(g) mod 256 -- it WRAPS for g >= 256. */
ORACLE_API void oracle_nms(const float *grad, const int16_t *sumx, const int16_t *sumy, int w, int h, uint8_t *nms)
{
  for (int y = 0; y < h; ++y) {
    for (int x = 0; x < w; ++x) {
      const size_t i = (size_t)y * w + x;
      const float g = grad[i];
#define G(yy, xx) (((yy) >= 0 && (yy) < h && (xx) >= 0 && (xx) < w) ? grad[(size_t)(yy) * w + (xx)] : 0.0f)
      float q, r;
      switch (oracle_sector(sumx[i], sumy[i])) {
      case 0: q = G(y + 1, x); r = G(y - 1, x); break;
      case 1: q = G(y + 1, x - 1); r = G(y - 1, x + 1); break;
      case 2: q = G(y, x + 1); r = G(y, x - 1); break;
      default: q = G(y - 1, x - 1); r = G(y + 1, x + 1); break;
      }
#undef G
      nms[i] = (q <= g && r <= g) ? (uint8_t)((uint32_t)g & 0xFFu) : 0;
    }
  }
}

/* src/cvp/cannyEdgeD.cu:273-293 -- strict '>' on unsigned char */
ORACLE_API void oracle_threshold(const uint8_t *nms, size_t n, uint8_t low, uint8_t high, uint8_t *thresh)
{
  for (size_t i = 0; i < n; ++i)
    thresh[i] = nms[i] > high ? FINAL_EDGE : nms[i] > low ? CANDIDATE_EDGE : NO_EDGE;
}

/* src/cvp/cannyEdgeD.cu:295-377 + src/cvp/cannyEdgeH.cu:297-338 as a fixpoint: a CANDIDATE pixel
 * becomes FINAL iff it is 8-connected to a FINAL pixel through CANDIDATE/FINAL pixels; thresh is zero
 * outside the image (:322-329); then removeCandidates (:379-395) maps the remaining 128s to 0.
 * This is what the reference's launch loop converges to when it stops with flag 0 in < 100 rounds. */
ORACLE_API void oracle_hysteresis(const uint8_t *thresh, int w, int h, uint8_t *edges)
{
  const size_t n = (size_t)w * h;
  size_t *stack = (size_t *)malloc(n * sizeof(size_t));
  size_t top = 0;
  memcpy(edges, thresh, n);
  for (size_t i = 0; i < n; ++i)
    if (edges[i] == FINAL_EDGE) stack[top++] = i;
  while (top) {
    const size_t i = stack[--top];
    const int y = (int)(i / w), x = (int)(i % w);
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = y + dy, xx = x + dx;
        if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
        const size_t j = (size_t)yy * w + xx;
        if (edges[j] == CANDIDATE_EDGE) {
          edges[j] = FINAL_EDGE;
          stack[top++] = j;
        }
      }
  }
  for (size_t i = 0; i < n; ++i)
    if (edges[i] == CANDIDATE_EDGE) edges[i] = NO_EDGE;
  free(stack);
}

/* Launch-level emulation of the reference loop (src/cvp/cannyEdgeH.cu:303-324): each launch closes
 * every 30x30 tile over a static 1-px halo taken from the previous global state
 * (src/cvp/cannyEdgeD.cu:307-363) and raises the flag when a tile changed (:375-376).  Returns the
 * number of launches after the first one (the reference's nbIters); *flag_out gets the last flag.
 * Used to refuse inputs that hit the 100-launch cap (contract T10 in SURVEY.md). */
ORACLE_API int oracle_hysteresis_launches(const uint8_t *thresh, int w, int h, int max_iters, uint8_t *state_out, int *flag_out)
{
  const int T = 30;
  const size_t n = (size_t)w * h;
  uint8_t *cur = (uint8_t *)malloc(n), *nxt = (uint8_t *)malloc(n);
  uint8_t tile[32][32];
  memcpy(cur, thresh, n);
  int iters = -1, flag = 1;
  while (flag && iters < max_iters) {
    flag = 0;
    for (int ty = 0; ty < (h + T - 1) / T; ++ty)
      for (int tx = 0; tx < (w + T - 1) / T; ++tx) {
        for (int r = 0; r < 32; ++r)
          for (int c = 0; c < 32; ++c) tile[r][c] = (uint8_t)px(cur, w, h, ty * T + r - 1, tx * T + c - 1);
        int changed_tile = 0, changed = 1;
        while (changed) {
          changed = 0;
          for (int r = 1; r <= T; ++r)
            for (int c = 1; c <= T; ++c) {
              if (tile[r][c] != CANDIDATE_EDGE) continue;
              int hit = 0;
              for (int dy = -1; dy <= 1 && !hit; ++dy)
                for (int dx = -1; dx <= 1; ++dx)
                  if (tile[r + dy][c + dx] == FINAL_EDGE) { hit = 1; break; }
              if (hit) { tile[r][c] = FINAL_EDGE; changed = 1; changed_tile = 1; }
            }
        }
        for (int r = 1; r <= T; ++r)
          for (int c = 1; c <= T; ++c) {
            const int y = ty * T + r - 1, x = tx * T + c - 1;
            if (y < h && x < w) nxt[(size_t)y * w + x] = tile[r][c];
          }
        flag += changed_tile;
      }
    uint8_t *t = cur; cur = nxt; nxt = t;
    ++iters;
  }
  if (state_out) memcpy(state_out, cur, n);
  if (flag_out) *flag_out = flag;
  free(cur); free(nxt);
  return iters;
}

/* src/cvp/cannyEdgeD.cu:35-50 -- saturating display view of the gradient (stage GRADIENT) */
ORACLE_API void oracle_float2uchar(const float *in, size_t n, uint8_t *out)
{
  for (size_t i = 0; i < n; ++i) {
    const float a = fabsf(in[i]);
    out[i] = (uint8_t)(a < 255.0f ? a : 255.0f);
  }
}

/* src/cvp/cannyEdgeH.hpp:25-29 -- threshold setters clamp against each other */
ORACLE_API void oracle_set_low(uint8_t *low, const uint8_t *high, uint8_t v) { *low = v < *high ? v : *high; }
ORACLE_API void oracle_set_high(const uint8_t *low, uint8_t *high, uint8_t v) { *high = v > *low ? v : *low; }

/* Whole path, src/cvp/cannyEdgeH.cu:49-120 with finalStage = HYSTER.  Any output pointer may be NULL.
 * All outputs are tightly packed w*h. */
/* channels = bytes per pixel: 3 = BGR8 (the reference), 4 = BGRA8 (same weights on bytes 0..2, byte 3 ignored),
 * 1 = GRAY8 (mono = the byte; the reference accepts CV_8UC1 in cvPipeline.cpp:32 but its upload is overwritten,
 * cannyEdgeH.cu:140-146 + :60-64, so there is no reference output to match -- this is the evident intent) */
ORACLE_API int oracle_canny_ch(const uint8_t *pix, size_t stride, int w, int h, int channels, uint8_t low, uint8_t high,
                               uint8_t *mono_o, uint8_t *blur_o, float *grad_o, uint8_t *sector_o, uint8_t *nms_o,
                               uint8_t *thresh_o, uint8_t *edges_o);

ORACLE_API int oracle_canny(const uint8_t *bgr, size_t stride, int w, int h, uint8_t low, uint8_t high,
                            uint8_t *mono_o, uint8_t *blur_o, float *grad_o, uint8_t *sector_o, uint8_t *nms_o,
                            uint8_t *thresh_o, uint8_t *edges_o)
{
  return oracle_canny_ch(bgr, stride, w, h, 3, low, high, mono_o, blur_o, grad_o, sector_o, nms_o, thresh_o, edges_o);
}

ORACLE_API int oracle_canny_ch(const uint8_t *bgr, size_t stride, int w, int h, int channels, uint8_t low, uint8_t high,
                               uint8_t *mono_o, uint8_t *blur_o, float *grad_o, uint8_t *sector_o, uint8_t *nms_o,
                               uint8_t *thresh_o, uint8_t *edges_o)
{
  if (channels != 1 && channels != 3 && channels != 4) return -2;
  const size_t n = (size_t)w * h;
  uint8_t *mono = (uint8_t *)malloc(n), *blur = (uint8_t *)malloc(n), *nms = (uint8_t *)malloc(n), *th = (uint8_t *)malloc(n);
  float *sx = (float *)malloc(n * 4), *sy = (float *)malloc(n * 4), *grad = (float *)malloc(n * 4);
  int16_t *gx = (int16_t *)malloc(n * 2), *gy = (int16_t *)malloc(n * 2);
  if (!mono || !blur || !nms || !th || !sx || !sy || !grad || !gx || !gy) return -1;
  if (channels == 3) oracle_rgb2mono(bgr, stride, w, h, mono);
  else
    for (int y = 0; y < h; ++y) {
      const uint8_t *q = bgr + (size_t)y * stride;
      for (int x = 0; x < w; ++x) {
        int v = channels == 1 ? q[x] : (q[4 * x] * B_WT + q[4 * x + 1] * G_WT + q[4 * x + 2] * R_WT) >> 6;
        mono[(size_t)y * w + x] = (uint8_t)(v > 255 ? 255 : v);
      }
    }
  oracle_gaussian(mono, w, h, blur);
  oracle_sobel(blur, w, h, sx, sy, gx, gy);
  oracle_grad(sx, sy, n, grad);
  oracle_nms(grad, gx, gy, w, h, nms);
  oracle_threshold(nms, n, low, high, th);
  if (edges_o) oracle_hysteresis(th, w, h, edges_o);
  if (mono_o) memcpy(mono_o, mono, n);
  if (blur_o) memcpy(blur_o, blur, n);
  if (grad_o) memcpy(grad_o, grad, n * 4);
  if (sector_o) for (size_t i = 0; i < n; ++i) sector_o[i] = (uint8_t)oracle_sector(gx[i], gy[i]);
  if (nms_o) memcpy(nms_o, nms, n);
  if (thresh_o) memcpy(thresh_o, th, n);
  free(mono); free(blur); free(nms); free(th); free(sx); free(sy); free(grad); free(gx); free(gy);
  return 0;
}
