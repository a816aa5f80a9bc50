/*
 * ref_harness.cu -- C-ABI driver for the UNMODIFIED reference kernels (test infrastructure).
 *
 * The reference's device code (/root/reference/src/cvp/cannyEdgeD.cu, .hpp) compiles from its own two
 * files; its host class (cannyEdgeH.cu) cannot be built here (needs C++ OpenCV headers, a live GL
 * context, spdlog, Conan).  oracle/Makefile compiles cannyEdgeD.cu where it lies and links it with this
 * file into oracle/_ref/libcvpref.so (git-ignored, shipped to the GPU box).  Nothing from the
 * reference is copied into the repo.
 *
 * This file re-creates what src/cvp/cannyEdgeH.cu does around those kernels, without GL/OpenCV:
 *   - buffers:  11 pitched allocations + 1 int flag          (cannyEdgeH.cu:340-385)
 *   - GK:       k * (1/159.0f) on the host, to constant mem   (cannyEdgeH.cu:372-380)
 *   - launches: 32x32 blocks, "simple" / tile-28 / tile-30 grids (cannyEdgeH.cu:214-295)
 *   - hysteresis: 1 + up to 100 launches, each bracketed by two blocking 4-byte memcpys,
 *     ping-pong swap, then removeCandidates              (cannyEdgeH.cu:297-338)
 *   - profiling: one event pair, record/record/synchronize per stage, ON by default
 *                                                             (cannyEdgeH.cu:24,409-430)
 *   - upload:   blocking cudaMemcpy2D from pageable host memory (cannyEdgeH.cu:122-152)
 *   - output:   cudaMemcpy2D D->D of the selected stage into a tight W*H byte buffer that stands in
 *     for the GL PBO, float2uchar for GRADIENT               (cannyEdgeH.cu:154-212)
 * Used by tests (golden generation, oracle validation) and by `bench.py --impl reference`.
 */
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>

#include "cannyEdgeD.hpp" /* found via -I/root/reference/src/cvp at build time */

namespace rk = cvp::cuda;

namespace
{
struct Plane
{
  void *p = nullptr;
  size_t pitch = 0; /* bytes */
};

enum BufId { RGB, MONO, BLUR, SOBELX, SOBELY, GRAD, SLOPE, NMS, THRESH, HYST, HYST_TMP, NBUF };
const int kElem[NBUF] = { 1, 1, 1, 4, 4, 4, 4, 1, 1, 1, 1 };

struct Ref
{
  int w = 0, h = 0;
  Plane b[NBUF];
  int *flag = nullptr;
  uint8_t *pbo = nullptr; /* tight w*h */
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  unsigned char lo = 10, hi = 40; /* cannyEdgeH.cu:22-23 */
  bool profiling = true;          /* cannyEdgeH.cu:24 */
  int nb_iters = 0, last_flag = 0;
  float ms[6] = { 0, 0, 0, 0, 0, 0 };
  int hyst_is = HYST; /* which buffer currently holds the result (the reference swaps pointers) */
};

int fail(cudaError_t e, const char *what)
{
  fprintf(stderr, "[cvpref] %s: %s\n", what, cudaGetErrorString(e));
  return -(int)e - 1000;
}
#define CK(x)                                   \
  do {                                          \
    cudaError_t e_ = (x);                       \
    if (e_ != cudaSuccess) return fail(e_, #x); \
  } while (0)

dim3 grid_for(const Ref *r, int out_tile) { return dim3((r->w + out_tile - 1) / out_tile, (r->h + out_tile - 1) / out_tile, 1); }
const dim3 kBlock(rk::MAX_2D_BLOCK_SIDE, rk::MAX_2D_BLOCK_SIDE, 1);

template <class T> T *ptr(Ref *r, int id) { return static_cast<T *>(r->b[id].p); }
int pitchB(Ref *r, int id) { return (int)r->b[id].pitch; }
int pitchE(Ref *r, int id) { return (int)(r->b[id].pitch / sizeof(float)); }

void t_begin(Ref *r)
{
  if (r->profiling) cudaEventRecord(r->e0);
}
void t_end(Ref *r, int stage)
{
  if (!r->profiling) return;
  cudaEventRecord(r->e1);
  cudaEventSynchronize(r->e1);
  cudaEventElapsedTime(&r->ms[stage], r->e0, r->e1);
}
}// namespace

extern "C" {

__attribute__((visibility("default"))) int cvpref_create(void **out, int w, int h)
{
  Ref *r = new (std::nothrow) Ref;
  if (!r) return -1;
  r->w = w;
  r->h = h;
  for (int i = 0; i < NBUF; ++i) {
    const size_t row = (size_t)w * kElem[i] * (i == RGB ? 3 : 1);
    CK(cudaMallocPitch(&r->b[i].p, &r->b[i].pitch, row, h));
    CK(cudaMemset2D(r->b[i].p, r->b[i].pitch, 0, row, h));
  }
  CK(cudaMalloc(&r->flag, sizeof(int)));
  CK(cudaMalloc(&r->pbo, (size_t)w * h));
  float gk[25] = { 2, 4, 5, 4, 2, 4, 9, 12, 9, 4, 5, 12, 15, 12, 5, 4, 9, 12, 9, 4, 2, 4, 5, 4, 2 };
  for (float &v : gk) v *= 1 / 159.0f;
  CK(cudaMemcpyToSymbol(rk::GK, gk, sizeof(gk)));
  CK(cudaEventCreate(&r->e0));
  CK(cudaEventCreate(&r->e1));
  *out = r;
  return 0;
}

__attribute__((visibility("default"))) void cvpref_destroy(void *h)
{
  Ref *r = static_cast<Ref *>(h);
  if (!r) return;
  for (auto &pl : r->b) cudaFree(pl.p);
  cudaFree(r->flag);
  cudaFree(r->pbo);
  if (r->e0) cudaEventDestroy(r->e0);
  if (r->e1) cudaEventDestroy(r->e1);
  delete r;
}

__attribute__((visibility("default"))) void cvpref_set_thresholds(void *h, unsigned char lo, unsigned char hi)
{
  Ref *r = static_cast<Ref *>(h);
  r->lo = 0;
  r->hi = std::max(hi, r->lo);
  r->lo = std::min(lo, r->hi);
}
__attribute__((visibility("default"))) void cvpref_enable_profiling(void *h, int on) { static_cast<Ref *>(h)->profiling = on != 0; }

/* finalStage as in src/cvp/define.hpp:9-17 (0 MONO .. 5 HYSTER).  host_bgr is pageable or pinned. */
__attribute__((visibility("default"))) int cvpref_run(void *h, const uint8_t *host_bgr, size_t stride, int final_stage)
{
  Ref *r = static_cast<Ref *>(h);
  const int w = r->w, hh = r->h;
  if (final_stage < 0 || final_stage > 5) return -2;
  CK(cudaMemcpy2D(r->b[RGB].p, r->b[RGB].pitch, host_bgr, stride, (size_t)w * 3, hh, cudaMemcpyHostToDevice));

  t_begin(r);
  rk::rgb2mono<<<grid_for(r, 32), kBlock>>>(ptr<uint8_t>(r, RGB), ptr<uint8_t>(r, MONO), w, hh, pitchB(r, RGB), pitchB(r, MONO));
  t_end(r, 0);
  if (final_stage >= 1) {
    t_begin(r);
    rk::gaussianFilter5x5<<<grid_for(r, 28), kBlock>>>(ptr<uint8_t>(r, MONO), ptr<uint8_t>(r, BLUR), w, hh, pitchB(r, MONO), pitchB(r, BLUR));
    t_end(r, 1);
  }
  if (final_stage >= 2) {
    t_begin(r);
    rk::sobelXY<<<grid_for(r, 30), kBlock>>>(ptr<uint8_t>(r, BLUR), ptr<float>(r, SOBELX), ptr<float>(r, SOBELY), w, hh, pitchB(r, BLUR), pitchE(r, SOBELX), pitchE(r, SOBELY));
    rk::gradSlope<<<grid_for(r, 32), kBlock>>>(ptr<float>(r, SOBELX), ptr<float>(r, SOBELY), ptr<float>(r, GRAD), ptr<float>(r, SLOPE), w, hh, pitchE(r, SOBELX), pitchE(r, SOBELY), pitchE(r, GRAD), pitchE(r, SLOPE));
    t_end(r, 2);
  }
  if (final_stage >= 3) {
    t_begin(r);
    rk::nonMaxSuppr<<<grid_for(r, 30), kBlock>>>(ptr<float>(r, GRAD), ptr<float>(r, SLOPE), ptr<uint8_t>(r, NMS), w, hh, pitchE(r, GRAD), pitchE(r, SLOPE), pitchB(r, NMS));
    t_end(r, 3);
  }
  if (final_stage >= 4) {
    t_begin(r);
    rk::doubleThreshold<<<grid_for(r, 32), kBlock>>>(ptr<uint8_t>(r, NMS), ptr<uint8_t>(r, THRESH), w, hh, pitchB(r, NMS), pitchB(r, THRESH), r->lo, r->hi);
    t_end(r, 4);
  }
  if (final_stage >= 5) {
    t_begin(r);
    int cur = HYST, other = HYST_TMP, flag = 0, iters = 0;
    cudaMemcpy(r->flag, &flag, sizeof(int), cudaMemcpyHostToDevice);
    rk::hysteresis<<<grid_for(r, 30), kBlock>>>(ptr<uint8_t>(r, THRESH), ptr<uint8_t>(r, cur), r->flag, w, hh, pitchB(r, THRESH), pitchB(r, cur));
    cudaMemcpy(&flag, r->flag, sizeof(int), cudaMemcpyDeviceToHost);
    while (iters < 100 && flag) {
      std::swap(cur, other); /* `other` now holds the previous state, `cur` receives the new one */
      flag = 0;
      cudaMemcpy(r->flag, &flag, sizeof(int), cudaMemcpyHostToDevice);
      rk::hysteresis<<<grid_for(r, 30), kBlock>>>(ptr<uint8_t>(r, other), ptr<uint8_t>(r, cur), r->flag, w, hh, pitchB(r, other), pitchB(r, cur));
      cudaMemcpy(&flag, r->flag, sizeof(int), cudaMemcpyDeviceToHost);
      ++iters;
    }
    r->nb_iters = iters;
    r->last_flag = flag;
    std::swap(cur, other); /* latest state becomes the input of removeCandidates */
    rk::removeCandidates<<<grid_for(r, 32), kBlock>>>(ptr<uint8_t>(r, other), ptr<uint8_t>(r, cur), w, hh, pitchB(r, other), pitchB(r, cur));
    r->hyst_is = cur;
    t_end(r, 5);
  }

  /* stand-in for _sendOutputToOpenGL */
  const int view[6] = { MONO, BLUR, GRAD, NMS, THRESH, r->hyst_is };
  if (final_stage == 2)
    rk::float2uchar<<<grid_for(r, 32), kBlock>>>(ptr<float>(r, GRAD), r->pbo, w, hh, pitchE(r, GRAD), w);
  else
    CK(cudaMemcpy2D(r->pbo, w, r->b[view[final_stage]].p, r->b[view[final_stage]].pitch, w, hh, cudaMemcpyDeviceToDevice));
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  return 0;
}

/* which: 0..9 = mono, blur, sobelX, sobelY, grad, slope, nms, thresh, hyster(final), pbo.  Tight rows. */
__attribute__((visibility("default"))) int cvpref_download(void *h, int which, void *host)
{
  Ref *r = static_cast<Ref *>(h);
  const int map[9] = { MONO, BLUR, SOBELX, SOBELY, GRAD, SLOPE, NMS, THRESH, r->hyst_is };
  if (which == 9) {
    CK(cudaMemcpy(host, r->pbo, (size_t)r->w * r->h, cudaMemcpyDeviceToHost));
    return 0;
  }
  if (which < 0 || which > 9) return -2;
  const int id = map[which];
  const size_t row = (size_t)r->w * kElem[id];
  CK(cudaMemcpy2D(host, row, r->b[id].p, r->b[id].pitch, row, r->h, cudaMemcpyDeviceToHost));
  return 0;
}

__attribute__((visibility("default"))) int cvpref_info(void *h, int *nb_iters, int *last_flag, float *stage_ms6)
{
  Ref *r = static_cast<Ref *>(h);
  if (nb_iters) *nb_iters = r->nb_iters;
  if (last_flag) *last_flag = r->last_flag;
  if (stage_ms6) memcpy(stage_ms6, r->ms, sizeof(r->ms));
  return 0;
}

/* Runs the reference gradSlope over every (sumX, sumY) in [-1020,1020]^2 (sobel outputs are sum/8.0f,
 * cannyEdgeD.cu:163,169) and returns grad and slope, row = sumY + 1020, col = sumX + 1020.
 * Lets tests pin the sector rule of the oracle to the real libdevice atan2f on this toolkit/GPU. */
__attribute__((visibility("default"))) int cvpref_gradslope_table(float *grad_out, float *slope_out)
{
  const int n = 2041;
  const size_t cnt = (size_t)n * n;
  float *hx = new float[cnt], *hy = new float[cnt];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      hx[(size_t)j * n + i] = (float)(i - 1020) / 8.0f;
      hy[(size_t)j * n + i] = (float)(j - 1020) / 8.0f;
    }
  float *dx, *dy, *dg, *ds;
  CK(cudaMalloc(&dx, cnt * 4));
  CK(cudaMalloc(&dy, cnt * 4));
  CK(cudaMalloc(&dg, cnt * 4));
  CK(cudaMalloc(&ds, cnt * 4));
  CK(cudaMemcpy(dx, hx, cnt * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy, hy, cnt * 4, cudaMemcpyHostToDevice));
  rk::gradSlope<<<dim3((n + 31) / 32, (n + 31) / 32, 1), kBlock>>>(dx, dy, dg, ds, n, n, n, n, n, n);
  CK(cudaGetLastError());
  CK(cudaMemcpy(grad_out, dg, cnt * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(slope_out, ds, cnt * 4, cudaMemcpyDeviceToHost));
  cudaFree(dx); cudaFree(dy); cudaFree(dg); cudaFree(ds);
  delete[] hx;
  delete[] hy;
  return 0;
}

/* Runs the reference nonMaxSuppr on caller-supplied grad/slope planes (tight, w*h) -- used with the
 * table above to pin sector selection and the (unsigned char) cast through the real kernel. */
__attribute__((visibility("default"))) int cvpref_nms_raw(const float *grad, const float *slope, int w, int h, uint8_t *nms_out)
{
  const size_t cnt = (size_t)w * h;
  float *dg, *ds;
  uint8_t *dn;
  CK(cudaMalloc(&dg, cnt * 4));
  CK(cudaMalloc(&ds, cnt * 4));
  CK(cudaMalloc(&dn, cnt));
  CK(cudaMemcpy(dg, grad, cnt * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ds, slope, cnt * 4, cudaMemcpyHostToDevice));
  rk::nonMaxSuppr<<<dim3((w + 29) / 30, (h + 29) / 30, 1), kBlock>>>(dg, ds, dn, w, h, w, w, w);
  CK(cudaGetLastError());
  CK(cudaMemcpy(nms_out, dn, cnt, cudaMemcpyDeviceToHost));
  cudaFree(dg); cudaFree(ds); cudaFree(dn);
  return 0;
}

}// extern "C"
