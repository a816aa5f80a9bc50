// k_stencil_fused.cuh -- the throughput kernel: BGR8 -> 2-bit weak/strong map in ONE launch.
//
// Replaces six launches of the reference (rgb2mono, gaussianFilter5x5, sobelXY, gradSlope, nonMaxSuppr,
// doubleThreshold: src/cvp/cannyEdgeD.cu:53-293, launched at src/cvp/cannyEdgeH.cu:214-295) and their 22 B/pixel of
// intermediate global traffic.  HBM traffic here is the 3 B/pixel of input (+ halo re-reads that hit L2) and the
// 0.25 B/pixel map.  No tensor cores: nothing in this path is a dense contraction; the kernel is bounded by
// integer/half2 issue rate and HBM, so the design goal is FEW INSTRUCTIONS PER PIXEL:
//
//   * a CTA (8 warps) owns a 240 x 60 output tile; lane l of every warp owns the 8 pixel columns X0-8+8l..+7
//     (lanes 0 and 31 are halo lanes), so every shared-memory access is one aligned 128-bit word per lane and
//     horizontal neighbours come from the adjacent lane by shuffle;
//   * two pixels per 32-bit register everywhere: gray via dp4a (weights x4 so that >>6 becomes "take byte 1"),
//     the 5x5 Gaussian as packed 16-bit integer sums (vertical 3-output filter, then horizontal combine),
//     /159 by one multiply-high per pixel whose upper half *is* the quotient and whose low half tells whether
//     S % 159 == 0, Sobel in exact fp16x2 arithmetic (all values are integers <= 2048), N = gx^2+gy^2 with the
//     sm_100 mixed-precision FMA (fp32 <- half*half+fp32, SASS FHFMA, half-select for free);
//   * the two data-dependent rarities are deferred to dense work lists in shared memory instead of diverging
//     in the hot loops: (1) pixels with S % 159 == 0, where the reference's 25-step fp32 FMA chain can land
//     just below the integer -- replayed exactly; (2) pixels above the low threshold -- only
//     those get direction, non-maximum suppression and the double threshold.
//
// Arithmetic contract: see k_stencil_tile.cuh (same results, bit for bit; tests compare both with the oracle).
#pragma once
#include "b2c_device.cuh"

namespace b2c
{
#ifndef B2C_FT_Y
#define B2C_FT_Y 36
#endif
constexpr int FT_X = 240, FT_Y = B2C_FT_Y, FT_THREADS = 256, FT_WARPS = 8;
constexpr int FT_CTAS_PER_SM = FT_Y <= 36 ? 3 : 2;
constexpr int FT_MROWS = FT_Y + 8;   // gray rows   Y0-4 .. Y0+63
constexpr int FT_BROWS = FT_Y + 4;   // blur rows   Y0-2 .. Y0+61
constexpr int FT_GROWS = FT_Y + 2;   // gx/gy rows  Y0-1 .. Y0+60
constexpr int FT_ROWB = 512;         // bytes per tile row: 256 columns x 16 bit
constexpr int FT_OUTW = FT_X / 16;   // map words per tile row
// shared-memory map (bytes).  gx/gy alias gray + flags (dead once the Gaussian is final).
constexpr int FS_MONO = 0;
constexpr int FS_FLAG = FS_MONO + FT_MROWS * FT_ROWB;   // blur rows x 256 B: byte == 0 <=> S % 159 == 0
constexpr int FT_R1 = FT_BROWS / FT_WARPS;                    // blur rows per warp
constexpr int FT_R3 = (FT_GROWS + FT_WARPS - 1) / FT_WARPS;   // gx/gy rows per warp
static_assert(FT_BROWS % FT_WARPS == 0, "blur rows must split evenly over the warps");
constexpr int FS_GX = 0;
constexpr int FS_GY = FT_GROWS * FT_ROWB;
constexpr int FS_A_END = 2 * FT_GROWS * FT_ROWB;
static_assert(FS_A_END >= FS_FLAG + FT_BROWS * 256, "alias region too small");
constexpr int FS_BLUR = FS_A_END;
constexpr int FS_CAND = FS_BLUR + FT_BROWS * FT_ROWB;   // 62 rows x 32 lanes, 1 byte = 8 candidate bits
constexpr int FS_OUT = FS_CAND + ((FT_GROWS * 32 + 127) / 128) * 128;   // FT_Y x 15 map words
constexpr int FS_LIST = FS_OUT + ((FT_Y * FT_OUTW * 4 + 127) / 128) * 128;
constexpr int FT_LIST_CAP = 2048;
constexpr int FS_CNT = FS_LIST + FT_LIST_CAP * 2;
constexpr int FUSED_SMEM = FS_CNT + 16;

// (B*7 + G*38 + R*19) >> 6 for 4 pixels held in 3 words of interleaved BGR (src/cvp/cannyEdgeD.cu:14-19,66-67).
// Weights x4 = (28,152,76): the sum x4 fits 16 bits and ">> 6" becomes "byte 1 of the dp4a result".
__device__ __forceinline__ void mono4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t &p01, uint32_t &p23)
{
  const uint32_t t0 = __dp4a(w0, 0x004C981Cu, 0u);
  const uint32_t t1 = __dp4a(w1, 0x00004C98u, __dp4a(w0, 0x1C000000u, 0u));
  const uint32_t t2 = __dp4a(w2, 0x0000004Cu, __dp4a(w1, 0x981C0000u, 0u));
  const uint32_t t3 = __dp4a(w2, 0x4C981C00u, 0u);
  p01 = __byte_perm(t0, t1, 0x7531);   // (gray0, gray1) as two 16-bit lanes
  p23 = __byte_perm(t2, t3, 0x7531);
}

// The reference's Gaussian for one pixel, replayed exactly: 25 fp32 FMAs in r-major, c-minor order starting from
// 0, then truncation (src/cvp/cannyEdgeD.cu:102-115).  M = gray tile as u16, (br, col) = blur tile coordinates.
__device__ __forceinline__ void gauss_replay(const B2cStencilParams &p, char *smem, int br, int col)
{
  const uint16_t *M = reinterpret_cast<const uint16_t *>(smem + FS_MONO) + br * 256 + col - 2;
  float f = 0.0f;
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c < 5; ++c) f = __fmaf_rn(p.gk[r * 5 + c], (float)M[r * 256 + c], f);
  reinterpret_cast<uint16_t *>(smem + FS_BLUR)[br * 256 + col] = (uint16_t)(unsigned)f;   // fp16 subnormal = the integer itself
}

// Direction, non-maximum suppression and double threshold for one candidate pixel (src/cvp/cannyEdgeD.cu:196,
// 239-267, 290), in exact fp32 on the integer Sobel sums: sector from 2|gx*gy| vs |gy^2-gx^2| (== the atan2
// sectors, pinned exhaustively in tests), keep iff both neighbours along it have N <= N (ties kept).
__device__ __forceinline__ void nms_item(const B2cStencilParams &p, char *smem, int g, int col)
{
  const uint16_t *GX = reinterpret_cast<const uint16_t *>(smem + FS_GX), *GY = reinterpret_cast<const uint16_t *>(smem + FS_GY);
  const int i = g * 256 + col;
  const uint32_t cx = GX[i], cy = GY[i];
  const float nx = b2c_fhfma_ll(cx, cx, 0.0f), ny = b2c_fhfma_ll(cy, cy, 0.0f), pr = b2c_fhfma_ll(cx, cy, 0.0f);
  const float n = nx + ny, d = ny - nx, a2 = fabsf(pr) + fabsf(pr);
  int o;
  if (a2 < fabsf(d)) o = (d > 0.0f) ? 256 : 1;   // sector 0: rows +-1;  sector 2: columns +-1
  else o = (pr > 0.0f) ? 255 : 257;              // sector 1: (y+1,x-1),(y-1,x+1);  sector 3: (y-1,x-1),(y+1,x+1)
  const uint32_t qx = GX[i + o], qy = GY[i + o], rx = GX[i - o], ry = GY[i - o];
  const float nq = b2c_fhfma_ll(qx, qx, b2c_fhfma_ll(qy, qy, 0.0f)), nr = b2c_fhfma_ll(rx, rx, b2c_fhfma_ll(ry, ry, 0.0f));
  if (nq > n || nr > n) return;
  bool strong, weak;
  if (n < p.n_wrap[0]) {   // trunc(grad) < 256: no wrap of the (unsigned char) cast; n >= n_lo[0] is what made it a candidate
    strong = n >= p.n_hi[0];
    weak = !strong;
  } else {
    strong = (n >= p.n_hi[1] && n < p.n_wrap[1]) || n >= p.n_hi[2];
    weak = !strong && ((n >= p.n_lo[1] && n < p.n_wrap[1]) || n >= p.n_lo[2]);
  }
  if (strong || weak) {
    const int c = col - 8;
    atomicOr(reinterpret_cast<uint32_t *>(smem + FS_OUT) + (g - 1) * FT_OUTW + (c >> 4), (strong ? 1u : 0x10000u) << (c & 15));
  }
}

// append one 16-bit item to the CTA's work list (plain shared-memory atomic, no warp aggregation)
__device__ __forceinline__ void list_push(char *smem, int which, uint32_t item)
{
#ifdef B2C_EMU
  const int idx = atomicAdd(reinterpret_cast<int *>(smem + FS_CNT) + which, 1);
#else
  int idx;
  asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(idx) : "r"((unsigned)__cvta_generic_to_shared(smem + FS_CNT + 4 * which)) : "memory");
#endif
  if (idx < FT_LIST_CAP) reinterpret_cast<uint16_t *>(smem + FS_LIST)[idx] = (uint16_t)item;
}

__global__ void __launch_bounds__(FT_THREADS, FT_CTAS_PER_SM) k_stencil_fused(const B2cStencilParams p)
{
  B2C_DYN_SMEM(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int X0 = blockIdx.x * FT_X, Y0 = blockIdx.y * FT_Y, frame = blockIdx.z;
  const int xl = X0 - 8 + 8 * lane;              // first pixel column of this lane
  const bool lane_in = xl >= 0 && xl < p.w;      // w % 8 == 0: a lane is wholly inside or wholly outside
  // CTA-uniform: does the 256-column window leave the image on either side?  Only then do lanes need masking.
  const bool xborder = X0 == 0 || X0 + 248 > p.w;
  const uint32_t lane_mask = lane_in ? 0xFFFFFFFFu : 0u;
  const int yg0 = Y0 + p.y0;                     // global row of tile row 0
  int *cnt = reinterpret_cast<int *>(smem + FS_CNT);
  uint32_t *s_out = reinterpret_cast<uint32_t *>(smem + FS_OUT);
  const uint16_t *list = reinterpret_cast<const uint16_t *>(smem + FS_LIST);

  for (int i = tid; i < FT_Y * FT_OUTW; i += FT_THREADS) s_out[i] = 0;
  if (tid < 4) cnt[tid] = 0;

  // ---- stage 0: gray, rows Y0-4 .. Y0+63; zero outside the image (cannyEdgeD.cu:91-98) ----------------------
  // Lanes outside the image read a block of zeros with stride 0, rows outside are skipped (warp-uniform).
  {
    constexpr int NR = (FT_MROWS + FT_WARPS - 1) / FT_WARPS;
    const long long lstride = lane_in ? p.row_stride : 0;
    const uint8_t *lp = lane_in ? p.bgr + (long long)frame * p.frame_stride + 3 * xl + (long long)(Y0 - 4 + warp) * p.row_stride : p.zeros;
    uint2 ld[NR][3];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int ry = warp + FT_WARPS * i, yg = yg0 - 4 + ry;
      if (ry < FT_MROWS && yg >= 0 && yg < p.h_glob) {
        const uint2 *q = reinterpret_cast<const uint2 *>(lp + (long long)(FT_WARPS * i) * lstride);
        ld[i][0] = __ldg(q);
        ld[i][1] = __ldg(q + 1);
        ld[i][2] = __ldg(q + 2);
      } else {
        ld[i][0] = ld[i][1] = ld[i][2] = make_uint2(0u, 0u);
      }
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int ry = warp + FT_WARPS * i;
      if (ry < FT_MROWS) {
        uint4 m;
        mono4(ld[i][0].x, ld[i][0].y, ld[i][1].x, m.x, m.y);
        mono4(ld[i][1].y, ld[i][2].x, ld[i][2].y, m.z, m.w);
        *reinterpret_cast<uint4 *>(smem + FS_MONO + ry * FT_ROWB + lane * 16) = m;
      }
    }
  }
  __syncthreads();

  // ---- stage 1: 5x5 Gaussian, rows Y0-2 .. Y0+61 (8 per warp) --------------------------------------------------
  // S = sum k*gray (exact, <= 40545) per 16-bit lane; q = S/159 via multiply-high; blur = q unless S%159 == 0.
  {
    const uint4 *mt = reinterpret_cast<const uint4 *>(smem + FS_MONO) + lane;
    const int b0 = warp * FT_R1;
    uint4 r0 = mt[(b0 + 0) * 32], r1 = mt[(b0 + 1) * 32], r2 = mt[(b0 + 2) * 32], r3 = mt[(b0 + 3) * 32];
#pragma unroll
    for (int k = 0; k < FT_R1; ++k) {
      const int br = b0 + k;
      const uint4 r4 = mt[(br + 4) * 32];
      const uint32_t a0[4] = { r0.x, r0.y, r0.z, r0.w }, a1[4] = { r1.x, r1.y, r1.z, r1.w }, a2[4] = { r2.x, r2.y, r2.z, r2.w },
                     a3[4] = { r3.x, r3.y, r3.z, r3.w }, a4[4] = { r4.x, r4.y, r4.z, r4.w };
      uint32_t v0[6], v1[6], v2[4];   // index j+1 = pair j; [0] / [5] come from the neighbouring lanes
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // vertical pass, weights per column offset: v0 = 2p+4q+5c (|dx|=2), v1 = 4p+9q+12c (|dx|=1), v2 = 5p+12q+15c (dx=0)
        const uint32_t pp = a0[j] + a4[j], c = a2[j], a = a1[j] + a3[j] + c;
        const uint32_t b = pp + 2u * a;
        const uint32_t w0 = c + 2u * b, d = a + c;
        const uint32_t w1 = d + 2u * w0;
        v0[j + 1] = w0;
        v1[j + 1] = w1;
        v2[j] = w0 + w1 - (pp + d);
      }
      v0[0] = __shfl_up_sync(B2C_FULL, v0[4], 1);
      v1[0] = __shfl_up_sync(B2C_FULL, v1[4], 1);
      v0[5] = __shfl_down_sync(B2C_FULL, v0[1], 1);
      v1[5] = __shfl_down_sync(B2C_FULL, v1[1], 1);
      uint32_t o1[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) o1[j] = __byte_perm(v1[j], v1[j + 1], 0x5432);   // (v1[2j-1], v1[2j])
      uint32_t hq[4], uu[4], vv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t S = (v0[j] + v0[j + 2] + v2[j]) + (o1[j] + o1[j + 1]);
        // upper 16 bits of the product = lane / 159; byte 1 == 0 <=> lane % 159 == 0
        const uint32_t U = __umulhi(S, 27012373u);         // high lane (the low lane adds < 1/159)
        const uint32_t V = __umulhi(S << 16, 27012373u);   // low lane
        uu[j] = U;
        vv[j] = V;
        // (q_lo, q_hi) as 16-bit integers == fp16 SUBNORMALS q * 2^-24: every later fp16 value is an integer
        // multiple of 2^-24 below 2^-14, so the fp16x2 arithmetic stays exact and needs no int->half conversion
        hq[j] = __byte_perm(V, U, 0x7632);
      }
      uint32_t f01 = __byte_perm(__byte_perm(vv[0], uu[0], 0x5151), __byte_perm(vv[1], uu[1], 0x5151), 0x5410);
      uint32_t f23 = __byte_perm(__byte_perm(vv[2], uu[2], 0x5151), __byte_perm(vv[3], uu[3], 0x5151), 0x5410);
      const int yg = yg0 - 2 + br;
      uint4 *bdst = reinterpret_cast<uint4 *>(smem + FS_BLUR + br * FT_ROWB + lane * 16);
      uint2 *fdst = reinterpret_cast<uint2 *>(smem + FS_FLAG + br * 256 + lane * 8);
      if (yg >= 0 && yg < p.h_glob) {   // warp-uniform
        if (xborder) {                   // CTA-uniform: blur is zero outside the image (cannyEdgeD.cu:142-149)
#pragma unroll
          for (int j = 0; j < 4; ++j) hq[j] &= lane_mask;
          f01 |= ~lane_mask;
          f23 |= ~lane_mask;
        }
        *bdst = make_uint4(hq[0], hq[1], hq[2], hq[3]);
        *fdst = make_uint2(f01, f23);
      } else {
        *bdst = make_uint4(0u, 0u, 0u, 0u);
        *fdst = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
      }
      r0 = r1; r1 = r2; r2 = r3; r3 = r4;
    }
  }
  __syncthreads();

  // ---- stage 2: exact replay of the S % 159 == 0 pixels (columns 6..249 of the window feed later stages) --------
  {
    const uint4 *ft = reinterpret_cast<const uint4 *>(smem + FS_FLAG);
    constexpr uint32_t K1 = 0x01010101u, K8 = 0x80808080u;
#pragma unroll
    for (int j = 0; j < (FT_BROWS * 16 + FT_THREADS - 1) / FT_THREADS; ++j) {
      const int i = tid + FT_THREADS * j;
      if (i >= FT_BROWS * 16) break;
      const uint4 f = ft[i];
      const uint32_t z = (((f.x - K1) & ~f.x) | ((f.y - K1) & ~f.y) | ((f.z - K1) & ~f.z) | ((f.w - K1) & ~f.w)) & K8;
      if (z) {   // some byte of these 16 is zero (a byte above a zero byte may be a false positive: re-checked below)
        const uint32_t w[4] = { f.x, f.y, f.z, f.w };
        uint32_t m16 = 0u;   // bit (4q+b) <=> byte b of word q looks zero
#pragma unroll
        for (int q = 0; q < 4; ++q) m16 |= (((((w[q] - K1) & ~w[q] & K8) >> 7) * 0x00204081u) >> 21 & 0xFu) << (4 * q);
        const int row8 = (i >> 4) << 8, col0 = (i & 15) * 16;
        const uint8_t *fb = reinterpret_cast<const uint8_t *>(ft + i);
        while (m16) {
          const int k = __ffs((int)m16) - 1;
          m16 &= m16 - 1u;
          const int col = col0 + k;
          if (fb[k] == 0 && col >= 6 && col < 250) list_push(smem, 0, (uint32_t)(row8 | col));
        }
      }
    }
  }
  __syncthreads();
  {
    const int n = cnt[0];
    if (n <= FT_LIST_CAP) {
      for (int i = tid; i < n; i += FT_THREADS) gauss_replay(p, smem, list[i] >> 8, list[i] & 255);
    } else {   // flat pictures: (almost) every pixel takes the replay
      const uint8_t *fb = reinterpret_cast<const uint8_t *>(smem + FS_FLAG);
      for (int i = tid; i < FT_BROWS * 256; i += FT_THREADS) {
        const int col = i & 255;
        if (fb[i] == 0 && col >= 6 && col < 250) gauss_replay(p, smem, i >> 8, col);
      }
    }
  }
  __syncthreads();

  // ---- stage 3a: Sobel sums in exact fp16x2, rows Y0-1 .. Y0+60; N >= low threshold -> candidate bit -------------
  {
    const uint4 *bt = reinterpret_cast<const uint4 *>(smem + FS_BLUR) + lane;
    const int g0 = warp * FT_R3;
    uint32_t D0[4], T0[4], D1[4], T1[4], D2[4], T2[4];
    const float negl = -p.n_lo[0];
    auto dt = [&](int row, uint32_t (&D)[4], uint32_t (&T)[4]) {
      const uint4 b = bt[row * 32];
      const uint32_t B[6] = { __shfl_up_sync(B2C_FULL, b.w, 1), b.x, b.y, b.z, b.w, __shfl_down_sync(B2C_FULL, b.x, 1) };
      uint32_t O[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) O[j] = __byte_perm(B[j], B[j + 1], 0x5432);   // (blur[2j-1], blur[2j])
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        D[j] = b2c_h2sub(O[j + 1], O[j]);                            // blur(x+1) - blur(x-1)      (cannyEdgeD.cu:158-160)
        T[j] = b2c_h2add(b2c_h2fma2(B[j + 1], O[j]), O[j + 1]);      // blur(x-1)+2blur(x)+blur(x+1) (:164-166)
      }
    };
    if (g0 < FT_GROWS) {   // warp-uniform
      dt(g0, D0, T0);
      dt(g0 + 1, D1, T1);
#pragma unroll
      for (int k = 0; k < FT_R3; ++k) {
        const int g = g0 + k;
        if (g < FT_GROWS) {
          dt(g + 2, D2, T2);
          const int yg = yg0 - 1 + g;
          uint4 *gxd = reinterpret_cast<uint4 *>(smem + FS_GX + g * FT_ROWB + lane * 16);
          uint4 *gyd = reinterpret_cast<uint4 *>(smem + FS_GY + g * FT_ROWB + lane * 16);
          uint8_t *cd = reinterpret_cast<uint8_t *>(smem + FS_CAND) + g * 32 + lane;
          if (yg >= 0 && yg < p.h_glob) {   // warp-uniform
            uint32_t gx[4], gy[4], nm = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              gx[j] = b2c_h2add(b2c_h2fma2(D1[j], D0[j]), D2[j]);   // sumX = right - left
              gy[j] = b2c_h2sub(T0[j], T2[j]);                      // sumY = top - bottom
            }
            if (xborder) {   // CTA-uniform: grad is zero outside the image (cannyEdgeD.cu:222-229)
#pragma unroll
              for (int j = 0; j < 4; ++j) { gx[j] &= lane_mask; gy[j] &= lane_mask; }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // sign bit of N - Nlow, shifted into a mask: bit (7-k) set <=> pixel k is NOT a candidate
              nm = __funnelshift_l(__float_as_uint(b2c_fhfma_ll(gx[j], gx[j], b2c_fhfma_ll(gy[j], gy[j], negl))), nm, 1);
              nm = __funnelshift_l(__float_as_uint(b2c_fhfma_hh(gx[j], gx[j], b2c_fhfma_hh(gy[j], gy[j], negl))), nm, 1);
            }
            *gxd = make_uint4(gx[0], gx[1], gx[2], gx[3]);
            *gyd = make_uint4(gy[0], gy[1], gy[2], gy[3]);
            *cd = (uint8_t)~nm;
          } else {
            *gxd = make_uint4(0u, 0u, 0u, 0u);
            *gyd = make_uint4(0u, 0u, 0u, 0u);
            *cd = 0;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            D0[j] = D1[j]; D1[j] = D2[j];
            T0[j] = T1[j]; T1[j] = T2[j];
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- stage 3b: candidates of the 240 x 60 output region -> work list -> NMS + double threshold ----------------
  for (int wi = tid; wi < FT_GROWS * 8; wi += FT_THREADS) {
    uint32_t w = reinterpret_cast<const uint32_t *>(smem + FS_CAND)[wi];   // 4 lanes x 8 pixels, bit (7-k) of a byte = pixel k
    const int g = wi >> 3, l4 = (wi & 7) * 4;
    if (l4 == 0) w &= 0xFFFFFF00u;    // lane 0 and lane 31 are halo lanes
    if (l4 == 28) w &= 0x00FFFFFFu;
    if (g < 1 || g > FT_Y || Y0 + g - 1 >= p.h) w = 0u;
    while (w) {
      const int bit = __ffs((int)w) - 1;
      w &= w - 1u;
      list_push(smem, 1, (uint32_t)((g << 8) | ((l4 + (bit >> 3)) * 8 + 7 - (bit & 7))));
    }
  }
  __syncthreads();
  {
    const int n = cnt[1];
    if (n <= FT_LIST_CAP) {
      for (int i = tid; i < n; i += FT_THREADS) nms_item(p, smem, list[i] >> 8, list[i] & 255);
    } else {   // dense pictures (noise): walk the candidate bits directly
      const uint8_t *cb = reinterpret_cast<const uint8_t *>(smem + FS_CAND);
      for (int i = tid; i < FT_Y * 240; i += FT_THREADS) {
        const int g = 1 + i / 240, col = 8 + i % 240;
        if (Y0 + g - 1 < p.h && ((cb[g * 32 + (col >> 3)] >> (7 - (col & 7))) & 1u)) nms_item(p, smem, g, col);
      }
    }
  }
  __syncthreads();

  // ---- stage 4: the 2-bit map tile -> global (16 threads per row, 15 words) --------------------------------------
  {
    const int wi = tid & 15, gw = blockIdx.x * FT_OUTW + wi;
    uint32_t *dst = p.map2 + (long long)frame * p.map_frame_stride + (long long)Y0 * p.map_pitch + gw;
    if (wi < FT_OUTW && gw < p.map_pitch) {
#pragma unroll
      for (int r = tid >> 4; r < FT_Y; r += FT_THREADS / 16)
        if (Y0 + r < p.h) dst[(long long)r * p.map_pitch] = s_out[r * FT_OUTW + wi];
    }
  }
}

#ifdef B2C_EMU
inline int fused_emu_launch(const B2cStencilParams &p)
{
  if (p.w % 8 || p.row_stride % 8 || p.frame_stride % 8 || (reinterpret_cast<uintptr_t>(p.bgr) & 7)) return -2;
  dim3 grid((p.w + FT_X - 1) / FT_X, (p.h + FT_Y - 1) / FT_Y, p.nframes);
  emu::launch(grid, dim3(FT_THREADS), FUSED_SMEM, false, [p] { k_stencil_fused(p); });
  return 0;
}
#else
inline cudaError_t fused_configure() { return cudaFuncSetAttribute(k_stencil_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM); }
// 8-byte aligned rows and whole lanes (w % 8 == 0); anything else goes through the tile kernel
inline bool fused_supported(const B2cStencilParams &p)
{
  return p.w % 8 == 0 && p.w >= 8 && p.row_stride % 8 == 0 && p.frame_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.bgr) & 7) == 0;
}
inline cudaError_t fused_launch(const B2cStencilParams &p, int, cudaStream_t st)
{
  dim3 grid((p.w + FT_X - 1) / FT_X, (p.h + FT_Y - 1) / FT_Y, p.nframes);
  k_stencil_fused<<<grid, FT_THREADS, FUSED_SMEM, st>>>(p);
  return cudaGetLastError();
}
#endif
}// namespace b2c
