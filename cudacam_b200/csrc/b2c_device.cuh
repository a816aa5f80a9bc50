// b2c_device.cuh -- shared definitions for the device code of the B200 Canny path.
//
// Data formats (all device-resident, see DESIGN.md "Data layout in HBM"):
//   BGR8 input      : interleaved bytes, byte 0 = B (reference: src/cvp/cannyEdgeD.cu:14-19,66-67)
//   bit planes S, C : one u32 per 32 pixels (bit i = pixel 32*k+i); S = strong, C = candidates (weak|strong): the two
//                     planes ARE the 2-bit weak/strong map, the equivalent of the reference's d_thresh {0,128,255}
//                     (src/cvp/cannyEdgeD.cu:273-293) at 2 bits per pixel, written by the stencil kernel.  Rows -1
//                     and h of every frame are zero ghost rows (= the reference's zero padding, cannyEdgeD.cu:322-329).
//   bit plane E     : edges = S | weak pixels promoted by the hysteresis (written by the resolve kernel).
//   2-bit map view  : one u32 per 16 pixels, bit i = strong flag of pixel 16*g+i, bit 16+i = weak-only flag
//                     (accessor format B2C_BUF_MAP2, made from the planes on demand).
//   edges           : u8 {0,255} per pixel = the reference's d_hyster after removeCandidates
//                     (src/cvp/cannyEdgeD.cu:379-395).
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef B2C_EMU
#include "cuda_emu.h"
#define B2C_DYN_SMEM(name) char *name = emu_dyn_smem()
#define B2C_GRID_SYNC() emu::grid_sync()
#else
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#define B2C_DYN_SMEM(name)                               \
  extern __shared__ __align__(128) char b2c_dyn_smem_[]; \
  char *name = b2c_dyn_smem_
#define B2C_GRID_SYNC() cooperative_groups::this_grid().sync()
#endif

#define B2C_FULL 0xffffffffu

// Stage ids == cvp::CannyStage (src/cvp/define.hpp:9-17)
enum { B2C_MONO = 0, B2C_GAUSSIAN = 1, B2C_GRADIENT = 2, B2C_NMS = 3, B2C_THRESH = 4, B2C_HYSTER = 5 };

struct B2cStencilParams {
  const uint8_t *bgr;        // frame 0, band row 0 (rows -4.. may be read in band mode)
  const uint8_t *zeros;      // >= 64 zero bytes, 16-byte aligned: what lanes outside the image read
  long long row_stride;      // bytes
  long long frame_stride;    // bytes
  int w, h;                  // width, rows produced by this launch
  int y0, h_glob;            // global row of row 0 and global image height (zero padding applies
                             // outside [0,h_glob), real neighbour rows are read inside it)
  int nframes;
  int channels;              // bytes per input pixel: 3 = BGR8, 4 = BGRA8 (alpha ignored), 1 = GRAY8
  long long plane_stride;    // 0 = interleaved pixels; else planar BGR8: bytes from the B plane to the G plane to the R plane
                             // (channels = 3, row_stride = bytes per plane row; tile kernel only)
  // out: the two bit planes of the 2-bit weak/strong map, as 16-bit halves (one u16 per 16 pixels: strips are 15
  // such groups wide): S = strong, C = weak | strong; row 0 of frame 0
  uint16_t *pl_S, *pl_C;
  int pl_pitch16;            // u16 per plane row
  long long pl_frame_stride16;
  unsigned lo, hi;           // thresholds (src/cvp/cannyEdgeH.cu:22-23, strict '>' cannyEdgeD.cu:290)
  float gk[25];              // k*(1/159.0f) rounded on the host (src/cvp/cannyEdgeH.cu:372-379)
  // double threshold in the N = sumX^2+sumY^2 domain.  The reference thresholds v = (unsigned char)grad
  // (cannyEdgeD.cu:267,290), grad = 0.5*sqrt(N); the cast wraps mod 256, and trunc(grad) >= m <=> N >= 4m^2, so
  //   v > T  <=>  N in [n[0], 4*256^2) or [n[1], 4*512^2) or [n[2], inf),  n[k] = 4*(256k + T + 1)^2
  // The marching kernel works on Sobel sums scaled by 2^-12 (fp16 normal numbers), so its N is scaled by 2^-24.
  float n_lo[3], n_hi[3], n_wrap[2];
  uint32_t n_pre;            // fp16x2: candidate pre-filter threshold on N' = fl16(gx'^2 + gy'^2), see b2c_fill_thresholds
  // optional per-stage outputs (frame 0 only; null = not written)
  uint8_t *mono, *blur, *nms, *thresh;
  float *grad;
  int pitch8, pitchf;        // in elements
  // row bands over peer memory (marching kernel only; null = off): the neighbours store the 4 halo rows above / below
  // the band while this launch already runs; the CTAs that read them first wait until the arrival counter (system
  // scope) reaches halo_need.  halo_err: set to 1 after a 2 s time-out.
  const uint32_t *halo_cnt_up, *halo_cnt_dn;
  int halo_need;
  int *halo_err;
};

struct B2cHystParams {
  const uint32_t *S, *C;     // planes written by the stencil: strong, weak | strong; row 0 of frame 0; rows -1 / h are zero ghost rows
  uint32_t *E;               // edges = S | promoted weak pixels (written by the resolve kernel), same geometry
  int plane_pitch;           // u32 per row
  long long plane_frame_stride;
  int w, h, nframes;
  uint8_t *edges;            // u8 {0,255} map; may be null (bit-plane consumers read E)
  long long edges_pitch, edges_frame_stride;
  int *flags;                // [3] on-device passes of the last run (= 1, out)
  int spread;                // tile kernel: 1 = deal the compacted work items round-robin to the warps, 0 = pack them
  int *parent;               // union-find parents, one int per pixel of the padded plane (only weak pixels are used)
  long long parent_frame_stride;
};

// Largest fp16 <= v (v >= 0, below the fp16 range limit), as its bit pattern.
static inline uint16_t b2c_f2h_rd(float v)
{
  if (!(v > 0.0f)) return 0;
  int e;
  const float m = frexpf(v, &e);   // v = m * 2^e, m in [0.5, 1)
  const int E = e - 1;             // v = (2m) * 2^E
  if (E < -14) return (uint16_t)floorf(ldexpf(v, 24));   // subnormal: multiples of 2^-24
  const uint32_t mant = (uint32_t)floorf(ldexpf(2.0f * m - 1.0f, 10));
  return (uint16_t)(((uint32_t)(E + 15) << 10) | mant);
}
// Thresholds of the marching kernel from low / high (p.lo, p.hi must be set): N = sumX^2 + sumY^2 scaled by 2^-24.
// The packed pre-filter computes N' = fl16(gy'^2 + fl16(gx'^2)) with gx' = sumX * 2^-12: each rounding is relative
// <= 2^-11 or, for subnormal results, absolute <= 2^-25, so N' >= N 2^-24 (1 - 2^-10) - 2^-24: the threshold below
// never rejects a pixel with N >= N_low.
static inline void b2c_fill_thresholds(B2cStencilParams &p)
{
  for (int k = 0; k < 3; ++k) {
    const float a = (float)(256 * k + p.lo + 1), b = (float)(256 * k + p.hi + 1);
    p.n_lo[k] = ldexpf(4.0f * a * a, -24);   // exact: < 2^24
    p.n_hi[k] = ldexpf(4.0f * b * b, -24);
  }
  p.n_wrap[0] = ldexpf(262144.0f, -24);    // 4*256^2
  p.n_wrap[1] = ldexpf(1048576.0f, -24);   // 4*512^2
  const float nlow = 4.0f * (float)(p.lo + 1) * (float)(p.lo + 1);
  const uint32_t h = b2c_f2h_rd(ldexpf(nlow * (1.0f - 1.0f / 1024.0f) - 1.0f, -24));
  p.n_pre = h | (h << 16);
}

// ---- packed fp16x2 helpers (operands are the raw 32-bit patterns) -------------------------------------------
// All values that pass through them are integers of magnitude <= 2048, which fp16 represents exactly, so
// this arithmetic is exact integer arithmetic on two pixels per instruction.
#ifdef B2C_EMU
static inline uint32_t emu_pack_h2(_Float16 lo, _Float16 hi) { uint16_t a, b; memcpy(&a, &lo, 2); memcpy(&b, &hi, 2); return (uint32_t)a | ((uint32_t)b << 16); }
static inline _Float16 emu_h_lo(uint32_t v) { uint16_t a = (uint16_t)v; _Float16 h; memcpy(&h, &a, 2); return h; }
static inline _Float16 emu_h_hi(uint32_t v) { uint16_t a = (uint16_t)(v >> 16); _Float16 h; memcpy(&h, &a, 2); return h; }
static inline uint32_t b2c_h2add(uint32_t a, uint32_t b) { return emu_pack_h2((_Float16)((float)emu_h_lo(a) + (float)emu_h_lo(b)), (_Float16)((float)emu_h_hi(a) + (float)emu_h_hi(b))); }
static inline uint32_t b2c_h2sub(uint32_t a, uint32_t b) { return emu_pack_h2((_Float16)((float)emu_h_lo(a) - (float)emu_h_lo(b)), (_Float16)((float)emu_h_hi(a) - (float)emu_h_hi(b))); }
static inline uint32_t b2c_h2fma2(uint32_t a, uint32_t c) { return emu_pack_h2((_Float16)(2.0f * (float)emu_h_lo(a) + (float)emu_h_lo(c)), (_Float16)(2.0f * (float)emu_h_hi(a) + (float)emu_h_hi(c))); }
static inline float b2c_fhfma_ll(uint32_t a, uint32_t b, float c) { return fmaf((float)emu_h_lo(a), (float)emu_h_lo(b), c); }
static inline float b2c_fhfma_hh(uint32_t a, uint32_t b, float c) { return fmaf((float)emu_h_hi(a), (float)emu_h_hi(b), c); }
static inline uint32_t b2c_u2h_bits(unsigned v) { _Float16 h = (_Float16)(float)v; uint16_t a; memcpy(&a, &h, 2); return a; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
#else
#include <cuda_fp16.h>
__device__ __forceinline__ uint32_t b2c_h2add(uint32_t a, uint32_t b) { uint32_t r; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t b2c_h2sub(uint32_t a, uint32_t b) { uint32_t r; asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// 2*a + c
__device__ __forceinline__ uint32_t b2c_h2fma2(uint32_t a, uint32_t c) { uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0x40004000u), "r"(c)); return r; }
// fp32 <- half * half + fp32 (sm_100 mixed-precision FMA, SASS FHFMA with free .H0/.H1 operand select)
__device__ __forceinline__ float b2c_fhfma_ll(uint32_t a, uint32_t b, float c)
{
  float r;
  asm("{.reg .f16 al, ah, bl, bh; mov.b32 {al, ah}, %1; mov.b32 {bl, bh}, %2; fma.rn.f32.f16 %0, al, bl, %3;}" : "=f"(r) : "r"(a), "r"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float b2c_fhfma_hh(uint32_t a, uint32_t b, float c)
{
  float r;
  asm("{.reg .f16 al, ah, bl, bh; mov.b32 {al, ah}, %1; mov.b32 {bl, bh}, %2; fma.rn.f32.f16 %0, ah, bh, %3;}" : "=f"(r) : "r"(a), "r"(b), "f"(c));
  return r;
}
__device__ __forceinline__ uint32_t b2c_u2h_bits(unsigned v) { return (uint32_t)__half_as_ushort(__uint2half_rn(v)); }
#endif

// Exact sector of the gradient direction from the integer Sobel sums -- same rule as
// oracle_sector() (pinned to the reference's atan2f path on the GPU, see tests/).
// Reference: src/cvp/cannyEdgeD.cu:196 (atan2(sX,sY)) and :239-264 (sector boundaries).
__device__ __forceinline__ int b2c_sector(int gx, int gy)
{
  const int a = gx < 0 ? -gx : gx, b = gy < 0 ? -gy : gy;
  const int apb = a + b, amb = a - b, b2 = 2 * b * b;
  if (a == 0 || apb * apb < b2) return 0;
  if (a > b && amb * amb > b2) return 2;
  return ((gx ^ gy) >= 0) ? 1 : 3;
}
