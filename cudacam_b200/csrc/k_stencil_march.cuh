// k_stencil_march.cuh -- the throughput kernel: BGR8 -> 2-bit weak/strong map in ONE launch, one WARP per strip.
//
// Replaces six launches of the reference (rgb2mono, gaussianFilter5x5, sobelXY, gradSlope, nonMaxSuppr,
// doubleThreshold: src/cvp/cannyEdgeD.cu:53-293, launched at src/cvp/cannyEdgeH.cu:214-295) and their 22 B/pixel of
// intermediate global traffic.  HBM traffic is the 3 B/pixel of input (halo re-reads hit L2) and the 0.25 B/pixel map.
// No tensor cores: nothing here is a dense contraction.  The kernel is bounded by integer / packed-half issue rate
// (alu and fma pipes each retire one warp instruction per two cycles), so the design goal is FEW INSTRUCTIONS PER
// PIXEL and NO BARRIERS:
//
//   * one warp owns a 240-column strip (lane l owns the 8 columns X0-8+8l .. +7; lanes 0 and 31 are halo lanes) and
//     MARCHES down a band of rows: every gray row is computed once (no vertical halo recompute inside a band), the
//     5-row gray window of the Gaussian and the Sobel row state live in registers, horizontal neighbours come from
//     the adjacent lane by shuffle.  Warps never synchronise with each other: a CTA is one warp;
//   * two pixels per 32-bit register everywhere: gray via dp4a (weights x4 so that >>6 becomes "take byte 1"), the
//     5x5 Gaussian as packed 16-bit integer sums S, S/159 by multiply-high, remainder S - 159*q by ONE packed
//     multiply-add (zero 16-bit lane <=> S % 159 == 0), Sobel in exact fp16x2 arithmetic on integers stored as fp16
//     subnormals, N = gx^2 + gy^2 with the mixed-precision FMA (fp32 <- half*half + fp32, SASS FHFMA);
//   * the two data-dependent rarities are deferred to dense per-warp work lists (ballot compaction, no atomics)
//     instead of diverging in the hot loops: (1) pixels with S % 159 == 0, where the reference's 25-step fp32 FMA
//     chain can land just below the integer -- replayed exactly once per block of 10 rows from a small
//     ring of gray rows; (2) pixels above the low threshold -- only those get direction, non-maximum suppression and
//     the double threshold, once per 2 rows from a 4-row ring of Sobel sums.
//
// Arithmetic contract: see k_stencil_tile.cuh (same results, bit for bit; tests compare both with the oracle).
#pragma once
#include "b2c_device.cuh"

namespace b2c
{
constexpr int MT_X = 240;            // output columns per strip
constexpr int MK = 10;               // blur rows per block (two unrolled groups of 5: the gray window has period 5)
constexpr int M_GRING = 16;          // gray ring rows   (bytes, 256 B per row)
constexpr int M_BRING = MK;          // blur rows of the current block (u16, 512 B per row); slot = row within the block
constexpr int M_SRING = 4;           // gx / gy ring rows (u16,  512 B per row each)
constexpr int M_OUTW = 16;           // map words per out-tile row (15 used)
constexpr int M_RCAP = 256;          // replay list capacity (per block of 10 rows)
constexpr int MS_GRAY = 0;
constexpr int MS_BLUR = MS_GRAY + M_GRING * 256;
constexpr int MS_GX = MS_BLUR + M_BRING * 512;
constexpr int MS_GY = MS_GX + M_SRING * 512;
constexpr int MS_OUT = MS_GY + M_SRING * 512;
constexpr int MS_LIST = MS_OUT + MK * M_OUTW * 4;
constexpr int MARCH_SMEM = MS_LIST + M_RCAP * 2;
constexpr int MARCH_CTAS_PER_SM = 15;

// (B*7 + G*38 + R*19) >> 6 for 4 pixels held in 3 words of interleaved BGR (src/cvp/cannyEdgeD.cu:14-19,66-67).
// Weights x4 = (28,152,76): the sum x4 fits 16 bits and ">> 6" becomes "byte 1 of the dp4a result".
__device__ __forceinline__ void m_mono4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t &p01, uint32_t &p23)
{
  const uint32_t t0 = __dp4a(w0, 0x004C981Cu, 0u);
  const uint32_t t1 = __dp4a(w1, 0x00004C98u, __dp4a(w0, 0x1C000000u, 0u));
  const uint32_t t2 = __dp4a(w2, 0x0000004Cu, __dp4a(w1, 0x981C0000u, 0u));
  const uint32_t t3 = __dp4a(w2, 0x4C981C00u, 0u);
  p01 = __byte_perm(t0, t1, 0x7531);   // (gray0, gray1) as two 16-bit lanes
  p23 = __byte_perm(t2, t3, 0x7531);
}

// The reference's Gaussian for one pixel, replayed exactly: 25 fp32 FMAs in r-major, c-minor order starting from
// 0, then truncation (src/cvp/cannyEdgeD.cu:102-115).  b = band-local blur row (rr = its row within the block),
// col = window column.
__device__ __forceinline__ void m_gauss_replay(const B2cStencilParams &p, char *smem, int b, int rr, int col)
{
  const uint8_t *G = reinterpret_cast<const uint8_t *>(smem + MS_GRAY) + col - 2;
  float f = 0.0f;
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const uint8_t *row = G + ((b - 2 + r + 16) & (M_GRING - 1)) * 256;
#pragma unroll
    for (int c = 0; c < 5; ++c) f = __fmaf_rn(p.gk[r * 5 + c], (float)row[c], f);
  }
  reinterpret_cast<uint16_t *>(smem + MS_BLUR)[rr * 256 + col] = (uint16_t)(unsigned)f;   // fp16 subnormal = the integer itself
}

// Direction, non-maximum suppression and double threshold for one pixel (src/cvp/cannyEdgeD.cu:196, 239-267, 290),
// in exact fp32 on the integer Sobel sums: sector from 2|gx*gy| vs |gy^2-gx^2| (== the atan2 sectors, pinned
// exhaustively in tests), keep iff both neighbours along it have N <= N (ties kept).  n = band-local row.
__device__ __forceinline__ void m_nms_item(const B2cStencilParams &p, char *smem, int n, int col, int out_row)
{
  const uint16_t *GX = reinterpret_cast<const uint16_t *>(smem + MS_GX), *GY = reinterpret_cast<const uint16_t *>(smem + MS_GY);
  const int sn = (n + 4) & (M_SRING - 1);
  const int i = sn * 256 + col;
  const uint32_t cx = GX[i], cy = GY[i];
  const float nx = b2c_fhfma_ll(cx, cx, 0.0f), ny = b2c_fhfma_ll(cy, cy, 0.0f), pr = b2c_fhfma_ll(cx, cy, 0.0f);
  const float nn = nx + ny, d = ny - nx, a2 = fabsf(pr) + fabsf(pr);
  int dy, dx;
  if (a2 < fabsf(d)) {   // sector 0: rows +-1;  sector 2: columns +-1
    dy = (d > 0.0f) ? 1 : 0;
    dx = 1 - dy;
  } else {               // sector 1: (y+1,x-1),(y-1,x+1);  sector 3: (y-1,x-1),(y+1,x+1)
    dy = 1;
    dx = (pr > 0.0f) ? -1 : 1;
  }
  const int iq = ((sn + dy) & (M_SRING - 1)) * 256 + col + dx, ir = ((sn - dy) & (M_SRING - 1)) * 256 + col - dx;
  const uint32_t qx = GX[iq], qy = GY[iq], rx = GX[ir], ry = GY[ir];
  const float nq = b2c_fhfma_ll(qx, qx, b2c_fhfma_ll(qy, qy, 0.0f)), nr = b2c_fhfma_ll(rx, rx, b2c_fhfma_ll(ry, ry, 0.0f));
  if (nq > nn || nr > nn) return;
  bool strong, weak;
  if (nn < p.n_wrap[0]) {   // trunc(grad) < 256: no wrap of the (unsigned char) cast; nn >= n_lo[0] is what made it a candidate
    strong = nn >= p.n_hi[0];
    weak = !strong;
  } else {
    strong = (nn >= p.n_hi[1] && nn < p.n_wrap[1]) || nn >= p.n_hi[2];
    weak = !strong && ((nn >= p.n_lo[1] && nn < p.n_wrap[1]) || nn >= p.n_lo[2]);
  }
  if (strong || weak) {
    const int c = col - 8;
    atomicOr(reinterpret_cast<uint32_t *>(smem + MS_OUT) + out_row * M_OUTW + (c >> 4), (strong ? 1u : 0x10000u) << (c & 15));
  }
}

// CH = bytes per input pixel: 3 = BGR8 (the reference's format), 4 = BGRA8 (alpha ignored), 1 = GRAY8 (gray = the byte;
// the reference's own CV_8UC1 path is broken, SURVEY T13 -- this is what it evidently meant to do)
template <int CH>
__global__ void __launch_bounds__(32, MARCH_CTAS_PER_SM) k_stencil_march(const B2cStencilParams p, const int rb)
{
  B2C_DYN_SMEM(smem);
  const int lane = threadIdx.x & 31;
  const int X0 = blockIdx.x * MT_X, Y0 = blockIdx.y * rb, frame = blockIdx.z;
  const int rows_out = min(rb, p.h - Y0);        // this band produces rows Y0 .. Y0+rows_out-1
  const int xl = X0 - 8 + 8 * lane;              // first pixel column of this lane
  const bool lane_in = xl >= 0 && xl < p.w;      // w % 8 == 0: a lane is wholly inside or wholly outside
  const int yg0 = Y0 + p.y0;                     // global row of band-local row 0
  const int ilim = p.h + 4 - Y0;                 // band-local gray rows >= ilim are not backed by memory
  const uint32_t lt_mask = (1u << lane) - 1u;
  const bool out_lane = lane_in && lane != 0 && lane != 31;   // lanes 0 and 31 are halo lanes
  uint32_t *s_out = reinterpret_cast<uint32_t *>(smem + MS_OUT);
  uint16_t *list = reinterpret_cast<uint16_t *>(smem + MS_LIST);
  // which of this lane's 8 columns can feed an output pixel: window columns 6..249 for the blur (lane 0: px 6,7;
  // lane 31: px 0,1), bits laid out like the z accumulators (bit 8*(k&3) + 4*(k>>2) + row, k = pixel)
  // A lane outside the image never stores to the blur / gx / gy rings, so its slots keep the zeros written here:
  // that IS the reference's per-stage zero padding left and right of the image (cannyEdgeD.cu:142-149, 222-229).
  const uint32_t zkeep = !lane_in ? 0u : lane == 0 ? 0xF0F00000u : lane == 31 ? 0x00000F0Fu : 0xFFFFFFFFu;

  for (int i = lane; i < (MARCH_SMEM - MS_BLUR) / 16; i += 32) reinterpret_cast<uint4 *>(smem + MS_BLUR)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();

  const long long lstride = lane_in ? p.row_stride : 0;
  const uint8_t *lp = lane_in ? p.bgr + (long long)frame * p.frame_stride + CH * xl + (long long)Y0 * p.row_stride : p.zeros;

  // raw pixels of band-local gray row i (CH x 8 bytes per lane); zero outside the image (cannyEdgeD.cu:91-98)
  auto load_row = [&](int i, uint2 (&d)[CH]) {
    const int yg = yg0 + i;
    if (yg >= 0 && yg < p.h_glob && i < ilim) {   // warp-uniform
      const uint2 *q = reinterpret_cast<const uint2 *>(lp + (long long)i * lstride);
#pragma unroll
      for (int k = 0; k < CH; ++k) d[k] = __ldg(q + k);
    } else {
#pragma unroll
      for (int k = 0; k < CH; ++k) d[k] = make_uint2(0u, 0u);
    }
  };
  // gray of one row as 4 words of two 16-bit pixels; also kept as bytes in the gray ring for the replay
  auto gray_row = [&](int i, const uint2 (&d)[CH], uint32_t (&m)[4]) {
    uint2 bytes;
    if constexpr (CH == 3) {
      m_mono4(d[0].x, d[0].y, d[1].x, m[0], m[1]);
      m_mono4(d[1].y, d[2].x, d[2].y, m[2], m[3]);
      bytes = make_uint2(__byte_perm(m[0], m[1], 0x6420), __byte_perm(m[2], m[3], 0x6420));
    } else if constexpr (CH == 4) {   // one dp4a per pixel: (B*28 + G*152 + R*76 + A*0), byte 1 of the sum = gray
      uint32_t t[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        t[2 * k] = __dp4a(d[k].x, 0x004C981Cu, 0u);
        t[2 * k + 1] = __dp4a(d[k].y, 0x004C981Cu, 0u);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) m[j] = __byte_perm(t[2 * j], t[2 * j + 1], 0x7531);
      bytes = make_uint2(__byte_perm(m[0], m[1], 0x6420), __byte_perm(m[2], m[3], 0x6420));
    } else {                          // gray input: the bytes are the gray values
      bytes = d[0];
      m[0] = __byte_perm(bytes.x, 0u, 0x4140);
      m[1] = __byte_perm(bytes.x, 0u, 0x4342);
      m[2] = __byte_perm(bytes.y, 0u, 0x4140);
      m[3] = __byte_perm(bytes.y, 0u, 0x4342);
    }
    *reinterpret_cast<uint2 *>(smem + MS_GRAY + ((i + 16) & (M_GRING - 1)) * 256 + lane * 8) = bytes;
  };

#if !defined(B2C_EMU)
  // Optional pseudo-random start delay (option march_stagger_ns, default 0).  Hypothesis: stage A is alu-pipe work,
  // stage C fma-pipe work, so warps that start together fight for one pipe.  Measured: no effect at any band height
  // (the warps are not phase-locked) -- kept only as an experiment knob.
  if (p.stagger_ns > 0) {
    const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    __nanosleep(((bid * 2654435761u) >> 24) * (unsigned)p.stagger_ns >> 8);
  }
#endif
  // ---- prologue: gray rows -4 .. -1 into window slots 1 .. 4 -----------------------------------------------
  uint32_t win[5][4];
  uint2 pre[CH];   // raw pixels of the next gray row, in flight while the current row is processed
  {
    uint2 d[CH];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      load_row(-4 + k, d);
      gray_row(-4 + k, d, win[k + 1]);
    }
    load_row(0, pre);
  }
  // Sobel row state (rows are consumed in order): Dp = D(r-1), PX = D(r-2) + 2 D(r-1), Ta = T(r-2), Tb = T(r-1)
  uint32_t Dp[4] = { 0, 0, 0, 0 }, PX[4] = { 0, 0, 0, 0 }, Ta[4] = { 0, 0, 0, 0 }, Tb[4] = { 0, 0, 0, 0 };
  uint32_t cand_carry = 0u;   // candidate bits of the last Sobel row of the previous chunk
  // -N_low through a shuffle: the value is warp-uniform, and left in a uniform register it is copied into a vector
  // register once per FHFMA chain (8 copies per row); a shuffle result lives in a vector register
  const float negl = __shfl_sync(B2C_FULL, -p.n_lo[0], 0);
  const uint32_t cmask = out_lane ? 0xFFu : 0u;

  const int nblocks = (rows_out + 4 + MK - 1) / MK;
  for (int t = 0; t < nblocks; ++t) {
    const int b0 = -2 + MK * t;   // first blur row of this block
    uint32_t zacc[4] = { 0u, 0u, 0u, 0u };   // replay flags: [2*half] rows 0-3 of the half, [2*half+1] row 4
#if !defined(B2C_EMU) && !defined(B2C_MARCH_NO_L2PF)
    // pull the gray rows of the NEXT block (768 B per row and strip = 7 lines of 128 B) into L2 while this block
    // computes: 4 rows per instruction (lanes 7q .. 7q+6 take row q), 3 instructions for 12 rows
    if (CH == 3 && lane < 28) {   // (only for the 3-byte format: 7 lines per strip row)
      const int q = lane / 7;
      const long long xoff = (long long)(X0 - 8) * 3 + (lane - 7 * q) * 128;
      if (xoff >= 0 && xoff < (long long)p.w * 3) {
        const uint8_t *pf = p.bgr + (long long)frame * p.frame_stride + (long long)Y0 * p.row_stride + xoff;
#pragma unroll
        for (int it = 0; it < 3; ++it) {
          const int i = b0 + MK + 3 + 4 * it + q, yg = yg0 + i;
          if (yg >= 0 && yg < p.h_glob && i < ilim) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (long long)i * p.row_stride));
        }
      }
    }
#endif

    // ---- stage A: gray row b+2, 5x5 Gaussian of blur row b, for the 10 rows of the block ---------------------
    // S = sum k*gray (exact, <= 40545) per 16-bit lane; q = S/159 via multiply-high; remainder lane == 0 -> replay.
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      uint32_t zA = 0u, zB = 0u;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int rr = half * 5 + k;   // row within the block
        const int b = b0 + rr, g = b + 2;
        uint2 cur[CH];
#pragma unroll
        for (int q = 0; q < CH; ++q) cur[q] = pre[q];
        load_row(g + 1, pre);
        gray_row(g, cur, win[k]);
        const uint32_t(&a0)[4] = win[(k + 1) % 5], (&a1)[4] = win[(k + 2) % 5], (&a2)[4] = win[(k + 3) % 5], (&a3)[4] = win[(k + 4) % 5], (&a4)[4] = win[k];
        uint32_t v0[6], v1[6], v2[4];   // index j+1 = pair j; [0] / [5] come from the neighbouring lanes
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // vertical pass, weights per column offset: v0 = 2p+4q+5c (|dx|=2), v1 = 4p+9q+12c (|dx|=1), v2 = 5p+12q+15c (dx=0)
          const uint32_t pp = a0[j] + a4[j], c = a2[j], a = a1[j] + a3[j] + c;
          const uint32_t bb = pp + 2u * a;
          const uint32_t w0 = c + 2u * bb, d = a + c;
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
        uint32_t hq[4], rem[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t S = (v0[j] + v0[j + 2] + v2[j]) + (o1[j] + o1[j + 1]);
          const uint32_t U = __umulhi(S, 27012373u);         // upper 16 bits = high lane / 159 (the low lane adds < 1/159)
          const uint32_t V = __umulhi(S << 16, 27012373u);   // low lane
          // (q_lo, q_hi) as 16-bit integers == fp16 SUBNORMALS q * 2^-24: every later fp16 value is an integer
          // multiple of 2^-24 below 2^-14, so the fp16x2 arithmetic stays exact and needs no int->half conversion
          hq[j] = __byte_perm(V, U, 0x7632);
          rem[j] = S - 159u * hq[j];                         // both lanes at once: 159*q <= 40545 never carries
        }
        // remainders are < 159: a zero BYTE 0 / 2 marks S % 159 == 0.  Bit 7 of each byte of z01 / z23 <=> pixel
        // k = byte (+4) needs the replay (a zero byte can set the bit of the byte above it too: harmless, the
        // replay is exact for every pixel)
        const uint32_t r01 = __byte_perm(rem[0], rem[1], 0x6420), r23 = __byte_perm(rem[2], rem[3], 0x6420);
        uint32_t z01 = (r01 - 0x01010101u) & ~r01 & 0x80808080u, z23 = (r23 - 0x01010101u) & ~r23 & 0x80808080u;
        const int yb = yg0 + b;
        uint4 *bdst = reinterpret_cast<uint4 *>(smem + MS_BLUR + rr * 512 + lane * 16);
        if (yb >= 0 && yb < p.h_glob) {   // warp-uniform
          if (lane_in) *bdst = make_uint4(hq[0], hq[1], hq[2], hq[3]);
          const uint32_t zb = (z01 >> 7) | (z23 >> 3);
          if (k < 4) zA |= zb << k;
          else zB = zb;
        } else {
          *bdst = make_uint4(0u, 0u, 0u, 0u);   // blur is zero above / below the image
        }
      }
      if (half == 0) { zacc[0] = zA; zacc[1] = zB; }
      else { zacc[2] = zA; zacc[3] = zB; }
    }
    __syncwarp();

    // ---- stage B: exact replay of the S % 159 == 0 pixels of the block ----------------------------------------
    {
      // every lane appends its own flagged pixels at its offset in the list (exclusive prefix sum over the lanes)
#pragma unroll
      for (int wi = 0; wi < 4; ++wi) zacc[wi] &= zkeep;
      const int mine = __popc(zacc[0]) + __popc(zacc[1]) + __popc(zacc[2]) + __popc(zacc[3]);
      int incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(B2C_FULL, incl, d);
        if (lane >= d) incl += v;
      }
      const int cnt = __shfl_sync(B2C_FULL, incl, 31);
      if (cnt <= M_RCAP) {
        int off = incl - mine;
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
          const int row0 = (wi >> 1) * 5 + (wi & 1) * 4;
          uint32_t bits = zacc[wi];
          while (bits) {
            const int pb = __ffs((int)bits) - 1;
            bits &= bits - 1u;
            // bit pb: row = row0 + (pb & 3), pixel = (pb >> 3) + 4 * ((pb >> 2) & 1)
            list[off++] = (uint16_t)(((row0 + (pb & 3)) << 8) | (lane * 8 + (pb >> 3) + 4 * ((pb >> 2) & 1)));
          }
        }
        __syncwarp();
        for (int i = lane; i < cnt; i += 32) m_gauss_replay(p, smem, b0 + (list[i] >> 8), list[i] >> 8, list[i] & 255);
      } else {   // flat pictures: (almost) every pixel takes the replay -- do all of them, the replay is exact everywhere
        for (int i = lane; i < MK * 244; i += 32) {
          const int rr = i / 244, col = 6 + i % 244;
          const int yb = yg0 + b0 + rr, xg = X0 - 8 + col;
          if (yb >= 0 && yb < p.h_glob && xg >= 0 && xg < p.w) m_gauss_replay(p, smem, b0 + rr, rr, col);
        }
      }
      __syncwarp();
    }

    // ---- stages C / D: 5 chunks of 2 Sobel rows, each followed by NMS + double threshold of 2 rows ------------
#pragma unroll 1   // (fully unrolled the kernel is 55 KB of SASS and runs 20 % slower: instruction-cache misses)
    for (int ch = 0; ch < 5; ++ch) {
      const int s0 = b0 - 1 + 2 * ch;   // Sobel rows s0, s0+1 (new blur rows s0+1, s0+2)
      uint32_t cand[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int s = s0 + e, r = s + 1;
        const uint4 bw = *reinterpret_cast<const uint4 *>(smem + MS_BLUR + (2 * ch + e) * 512 + lane * 16);   // blur row r = s+1 = b0 + 2ch + e
        const uint32_t B[6] = { __shfl_up_sync(B2C_FULL, bw.w, 1), bw.x, bw.y, bw.z, bw.w, __shfl_down_sync(B2C_FULL, bw.x, 1) };
        uint32_t O[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) O[j] = __byte_perm(B[j], B[j + 1], 0x5432);   // (blur[2j-1], blur[2j])
        uint32_t gx[4], gy[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t D = b2c_h2sub(O[j + 1], O[j]);                           // blur(x+1) - blur(x-1)        (cannyEdgeD.cu:158-160)
          const uint32_t T = b2c_h2add(b2c_h2fma2(B[j + 1], O[j]), O[j + 1]);     // blur(x-1)+2blur(x)+blur(x+1) (:164-166)
          gx[j] = b2c_h2add(PX[j], D);          // sumX(s) = D(s-1) + 2 D(s) + D(s+1): right - left
          gy[j] = b2c_h2sub(Ta[j], T);          // sumY(s) = T(s-1) - T(s+1): top - bottom
          PX[j] = b2c_h2fma2(D, Dp[j]);
          Dp[j] = D;
          Ta[j] = Tb[j];
          Tb[j] = T;
        }

        const int ys = yg0 + s;
        uint4 *gxd = reinterpret_cast<uint4 *>(smem + MS_GX + ((s + 4) & (M_SRING - 1)) * 512 + lane * 16);
        uint4 *gyd = reinterpret_cast<uint4 *>(smem + MS_GY + ((s + 4) & (M_SRING - 1)) * 512 + lane * 16);
        if (ys >= 0 && ys < p.h_glob && s >= -1) {   // warp-uniform
          uint32_t nm = 0u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // sign bit of N - Nlow, shifted into a mask: bit (7-k) set <=> pixel k is NOT a candidate
            nm = __funnelshift_l(__float_as_uint(b2c_fhfma_ll(gx[j], gx[j], b2c_fhfma_ll(gy[j], gy[j], negl))), nm, 1);
            nm = __funnelshift_l(__float_as_uint(b2c_fhfma_hh(gx[j], gx[j], b2c_fhfma_hh(gy[j], gy[j], negl))), nm, 1);
          }
          if (lane_in) {
            *gxd = make_uint4(gx[0], gx[1], gx[2], gx[3]);
            *gyd = make_uint4(gy[0], gy[1], gy[2], gy[3]);
          }
          cand[e] = ~nm & ((s >= 0 && s < rows_out) ? cmask : 0u);
        } else {
          *gxd = make_uint4(0u, 0u, 0u, 0u);
          *gyd = make_uint4(0u, 0u, 0u, 0u);
          cand[e] = 0u;
        }
      }
      __syncwarp();
      // NMS rows n = s0-1 (candidates carried over from the previous chunk) and n = s0.  Work items are the
      // (row, lane) groups of 8 pixels that hold any candidate (~9 % of them, about half of their pixels set):
      // two ballots build the list, then 8 lanes -- one per pixel -- take each item, four items per pass.
      {
        const int nbase = b0 - 2;   // out-tile row 0 of this block
        const uint32_t c0 = cand_carry, c1 = cand[0];
        cand_carry = cand[1];
        const uint32_t m0 = __ballot_sync(B2C_FULL, c0 != 0u), m1 = __ballot_sync(B2C_FULL, c1 != 0u);
        const int n0 = __popc(m0), cnt = n0 + __popc(m1);
        if (cnt) {   // warp-uniform
          if (c0) list[__popc(m0 & lt_mask)] = (uint16_t)((lane << 8) | c0);
          if (c1) list[n0 + __popc(m1 & lt_mask)] = (uint16_t)(0x2000u | (lane << 8) | c1);
          __syncwarp();
          const int px = lane & 7;
          for (int it = lane >> 3; it < cnt; it += 4) {
            const uint32_t e = list[it];
            if ((e >> (7 - px)) & 1u) {   // bit (7-k) of the mask = pixel k
              const int n = s0 - 1 + (int)(e >> 13);
              m_nms_item(p, smem, n, (int)((e >> 8) & 31u) * 8 + px, n - nbase);
            }
          }
          __syncwarp();
        }
      }
    }

    // ---- stage E: the finished map rows n = b0-2 .. b0+7 -> global (2 rows of 15 words per store) --------------
    {
      const int wi = lane & 15, gw = blockIdx.x * (MT_X / 16) + wi;
      uint32_t *dst = p.map2 + (long long)frame * p.map_frame_stride + (long long)Y0 * p.map_pitch + gw;
#pragma unroll
      for (int it = 0; it < MK / 2; ++it) {
        const int idx = 2 * it + (lane >> 4), n = b0 - 2 + idx;
        const uint32_t v = s_out[idx * M_OUTW + wi];
        s_out[idx * M_OUTW + wi] = 0u;
        if (wi < MT_X / 16 && gw < p.map_pitch && n >= 0 && n < rows_out) dst[(long long)n * p.map_pitch] = v;
      }
      __syncwarp();
    }
  }
}

#ifdef B2C_EMU
inline int march_emu_launch(const B2cStencilParams &p, int rb)
{
  if (p.w % 8 || p.row_stride % 8 || p.frame_stride % 8 || (reinterpret_cast<uintptr_t>(p.bgr) & 7)) return -2;
  dim3 grid((p.w + MT_X - 1) / MT_X, (p.h + rb - 1) / rb, p.nframes);
  if (p.channels == 1) emu::launch(grid, dim3(32), MARCH_SMEM, false, [p, rb] { k_stencil_march<1>(p, rb); });
  else if (p.channels == 4) emu::launch(grid, dim3(32), MARCH_SMEM, false, [p, rb] { k_stencil_march<4>(p, rb); });
  else emu::launch(grid, dim3(32), MARCH_SMEM, false, [p, rb] { k_stencil_march<3>(p, rb); });
  return 0;
}
#else
inline cudaError_t march_configure()
{
  cudaError_t e = cudaFuncSetAttribute(k_stencil_march<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, MARCH_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_stencil_march<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MARCH_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_stencil_march<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, MARCH_SMEM);
  return e;
}
// 8-byte aligned rows and whole lanes (w % 8 == 0); anything else goes through the tile kernel
inline bool march_supported(const B2cStencilParams &p)
{
  return p.w % 8 == 0 && p.w >= 8 && p.row_stride % 8 == 0 && p.frame_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.bgr) & 7) == 0;
}
// Rows per band.  Bands of 10k-4 rows waste no block (a band of rb rows runs ceil((rb+4)/10) blocks of 10 rows).
// Measured on B200 (tools/sweep_rb.py): the kernel wants every SM full of warps all the time, so MANY short CTAs
// (the block scheduler balances them; 36 rows = 11 % recompute) beat one wave of long ones (276 rows = 4 %
// recompute but a ragged tail): 289 us vs 336 us on 64 x 1080p.  Cost model: waves x rows marched per CTA, with a
// penalty for few waves.
inline int march_band_rows(int w, int h, int nframes, int sm_count)
{
  const double slots = (double)sm_count * MARCH_CTAS_PER_SM;
  const long long strips = (w + MT_X - 1) / MT_X;
  int best_rb = MK - 4;
  double best = 1e30;
  for (int k = 1; k <= (h + 4 + MK - 1) / MK; ++k) {
    const int rb = MK * k - 4;
    const long long nb = (h + rb - 1) / rb;
    const double waves = (double)(strips * nb * nframes) / slots;
    const double cost = ceil(waves) * (MK * k + 3) * (1.0 + 0.5 / (waves > 1.0 ? waves : 1.0));
    if (cost < best) { best = cost; best_rb = rb; }
  }
  return best_rb;
}
inline cudaError_t march_launch(const B2cStencilParams &p, int sm_count, int rb_override, cudaStream_t st)
{
  const int rb = rb_override > 0 ? rb_override : march_band_rows(p.w, p.h, p.nframes, sm_count);
  dim3 grid((p.w + MT_X - 1) / MT_X, (p.h + rb - 1) / rb, p.nframes);
  if (p.channels == 1) k_stencil_march<1><<<grid, 32, MARCH_SMEM, st>>>(p, rb);
  else if (p.channels == 4) k_stencil_march<4><<<grid, 32, MARCH_SMEM, st>>>(p, rb);
  else k_stencil_march<3><<<grid, 32, MARCH_SMEM, st>>>(p, rb);
  return cudaGetLastError();
}
#endif
}// namespace b2c
