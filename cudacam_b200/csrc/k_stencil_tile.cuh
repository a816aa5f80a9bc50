// k_stencil_tile.cuh -- staged (shared-memory tile) version of the fused stencil.
//
// One CTA produces a 64x32 output tile: BGR8 -> mono -> 5x5 Gaussian -> Sobel -> N = gx^2+gy^2 and
// direction sector -> non-max suppression -> double threshold -> 2-bit map.  Every intermediate lives
// in shared memory; with EMIT the per-stage buffers the reference exposes through its `finalStage`
// switch (src/cvp/cannyEdgeH.cu:169-207) are also written (mono, blur, grad, nms, thresh).
// This is the all-stages / accessor path; the throughput path is k_stencil_fused.cuh.
//
// Arithmetic contract (bit-exact to the nvcc-compiled reference):
//   mono   (B*7 + G*38 + R*19) >> 6                              cannyEdgeD.cu:14-19,66-67
//   blur   trunc of the 25-step fp32 FMA chain with GK=k*(1/159.0f); equals S/159 for the integer
//          sum S unless S % 159 == 0, where the chain is replayed   cannyEdgeD.cu:102-115
//   zero padding per stage (mono, blur, grad) outside the image   cannyEdgeD.cu:91-98,142-149,222-229
//   grad   4*sqrtf(sX^2+sY^2) == 0.5f*sqrtf(N)                     cannyEdgeD.cu:195
//   nms    keep iff q<=g && r<=g (ties kept), value trunc(g) & 255 cannyEdgeD.cu:245-267
//   thresh v>high -> strong, v>low -> weak                         cannyEdgeD.cu:290
#pragma once
#include "b2c_device.cuh"

namespace b2c
{
constexpr int TILE_W = 64, TILE_H = 32, TILE_THREADS = 256;
constexpr int MW = TILE_W + 8, MH = TILE_H + 8;   // mono tile, halo 4
constexpr int BW = TILE_W + 4, BH = TILE_H + 4;   // blur tile, halo 2
constexpr int NW = TILE_W + 2, NH = TILE_H + 2;   // N tile, halo 1
constexpr int TILE_SMEM = MW * MH + BW * BH + NW * NH * 4 + NW * NH + 64;

template <bool EMIT>
__global__ void __launch_bounds__(TILE_THREADS) k_stencil_tile(const B2cStencilParams p)
{
  B2C_DYN_SMEM(smem);
  uint32_t *s_n = reinterpret_cast<uint32_t *>(smem);
  uint8_t *s_mono = reinterpret_cast<uint8_t *>(s_n + NW * NH);
  uint8_t *s_blur = s_mono + MW * MH;
  uint8_t *s_sec = s_blur + BW * BH;

  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TILE_W, y0t = blockIdx.y * TILE_H, frame = blockIdx.z;
  const uint8_t *src = p.bgr + (long long)frame * p.frame_stride;

  // 1. gray, halo 4; zero outside the (global) image
  for (int i = tid; i < MW * MH; i += TILE_THREADS) {
    const int r = i / MW, c = i - r * MW;
    const int y = y0t + r - 4, x = x0 + c - 4, yg = y + p.y0;
    unsigned v = 0;
    // (rows past the 4 halo rows below the band / frame feed no stored pixel and may lie outside the caller's buffer)
    if (x >= 0 && x < p.w && yg >= 0 && yg < p.h_glob && y < p.h + 4) {
      if (p.plane_stride) {   // planar BGR8
        const uint8_t *q = src + (long long)y * p.row_stride + x;
        v = (q[0] * 7u + q[p.plane_stride] * 38u + q[2 * p.plane_stride] * 19u) >> 6;
      } else {
        const uint8_t *q = src + (long long)y * p.row_stride + p.channels * x;
        v = p.channels == 1 ? q[0] : (q[0] * 7u + q[1] * 38u + q[2] * 19u) >> 6;   // BGR8 / BGRA8 (alpha ignored) / GRAY8
      }
    }
    s_mono[i] = (uint8_t)v;
  }
  __syncthreads();

  // 2. Gaussian, halo 2
  for (int i = tid; i < BW * BH; i += TILE_THREADS) {
    const int r = i / BW, c = i - r * BW;
    const int y = y0t + r - 2, x = x0 + c - 2, yg = y + p.y0;
    unsigned q = 0;
    if (x >= 0 && x < p.w && yg >= 0 && yg < p.h_glob) {
      const uint8_t *m = s_mono + r * MW + c;
      unsigned e0 = 0, e1 = 0, e2 = 0;   // column-symmetric row sums, weights rows (2,4,5),(4,9,12),(5,12,15)
#define ROW(k) (2u * (m[(k)*MW] + m[(k)*MW + 4]) + 4u * (m[(k)*MW + 1] + m[(k)*MW + 3]) + 5u * m[(k)*MW + 2])
#define ROW1(k) (4u * (m[(k)*MW] + m[(k)*MW + 4]) + 9u * (m[(k)*MW + 1] + m[(k)*MW + 3]) + 12u * m[(k)*MW + 2])
#define ROW2(k) (5u * (m[(k)*MW] + m[(k)*MW + 4]) + 12u * (m[(k)*MW + 1] + m[(k)*MW + 3]) + 15u * m[(k)*MW + 2])
      e0 = ROW(0) + ROW(4);
      e1 = ROW1(1) + ROW1(3);
      e2 = ROW2(2);
#undef ROW
#undef ROW1
#undef ROW2
      const unsigned S = e0 + e1 + e2;
      q = S / 159u;
      if (q * 159u == S) {   // the only case where the fp32 chain can land below the integer
        float f = 0.0f;
#pragma unroll
        for (int rr = 0; rr < 5; ++rr)
#pragma unroll
          for (int cc = 0; cc < 5; ++cc) f = __fmaf_rn(p.gk[rr * 5 + cc], (float)m[rr * MW + cc], f);
        q = (unsigned)f;
      }
    }
    s_blur[i] = (uint8_t)q;
  }
  __syncthreads();

  // 3. Sobel sums, N and sector, halo 1
  for (int i = tid; i < NW * NH; i += TILE_THREADS) {
    const int r = i / NW, c = i - r * NW;
    const int y = y0t + r - 1, x = x0 + c - 1, yg = y + p.y0;
    unsigned n = 0, sec = 0;
    if (x >= 0 && x < p.w && yg >= 0 && yg < p.h_glob) {
      const uint8_t *b = s_blur + r * BW + c;
      const int a00 = b[0], a01 = b[1], a02 = b[2];
      const int a10 = b[BW], a12 = b[BW + 2];
      const int a20 = b[2 * BW], a21 = b[2 * BW + 1], a22 = b[2 * BW + 2];
      const int gx = (a02 - a00) + 2 * (a12 - a10) + (a22 - a20);
      const int gy = (a00 + 2 * a01 + a02) - (a20 + 2 * a21 + a22);
      n = (unsigned)(gx * gx + gy * gy);
      sec = (unsigned)b2c_sector(gx, gy);
    }
    s_n[i] = n;
    s_sec[i] = (uint8_t)sec;
  }
  __syncthreads();

  // 4. NMS + double threshold + pack; a warp covers 32 consecutive pixels of one row
  const int lane = tid & 31;
  for (int it = 0; it < TILE_W * TILE_H / TILE_THREADS; ++it) {
    const int idx = it * TILE_THREADS + tid;
    const int r = idx / TILE_W, c = idx - r * TILE_W;
    const int y = y0t + r, x = x0 + c;
    const bool valid = (y < p.h) && (x < p.w);
    const uint32_t *n = s_n + (r + 1) * NW + (c + 1);
    const unsigned ng = n[0];
    unsigned nq, nr;
    switch (s_sec[(r + 1) * NW + (c + 1)]) {
    case 0: nq = n[NW]; nr = n[-NW]; break;
    case 1: nq = n[NW - 1]; nr = n[-NW + 1]; break;
    case 2: nq = n[1]; nr = n[-1]; break;
    default: nq = n[-NW - 1]; nr = n[NW + 1]; break;
    }
    const float g = 0.5f * __fsqrt_rn((float)ng);
    const unsigned v = (nq <= ng && nr <= ng) ? ((unsigned)g & 255u) : 0u;
    const bool strong = valid && v > p.hi;
    const bool weak = valid && !strong && v > p.lo;
    const unsigned bs = __ballot_sync(B2C_FULL, strong), bw = __ballot_sync(B2C_FULL, weak);
    if ((lane & 15) == 0 && valid) {
      const unsigned sh = lane & 16;
      const long long o = (long long)frame * p.pl_frame_stride16 + (long long)y * p.pl_pitch16 + (x >> 4);
      const unsigned s16 = (bs >> sh) & 0xFFFFu, w16 = (bw >> sh) & 0xFFFFu;
      p.pl_S[o] = (uint16_t)s16;
      p.pl_C[o] = (uint16_t)(s16 | w16);
    }
    if (EMIT && valid && frame == 0) {
      if (p.mono) p.mono[(long long)y * p.pitch8 + x] = s_mono[(r + 4) * MW + c + 4];
      if (p.blur) p.blur[(long long)y * p.pitch8 + x] = s_blur[(r + 2) * BW + c + 2];
      if (p.grad) p.grad[(long long)y * p.pitchf + x] = g;
      if (p.nms) p.nms[(long long)y * p.pitch8 + x] = (uint8_t)v;
      if (p.thresh) p.thresh[(long long)y * p.pitch8 + x] = strong ? 255 : weak ? 128 : 0;
    }
  }
}
}// namespace b2c
