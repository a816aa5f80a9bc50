// k_stencil_march.cuh -- the throughput kernel: BGR8 -> 2-bit weak/strong map (as the bit planes S and C) in ONE launch.
//
// Replaces six launches of the reference (rgb2mono, gaussianFilter5x5, sobelXY, gradSlope, nonMaxSuppr,
// doubleThreshold: src/cvp/cannyEdgeD.cu:53-293, launched at src/cvp/cannyEdgeH.cu:214-295) and their 22 B/pixel of
// intermediate global traffic.  HBM traffic is the 3 B/pixel of input (halo re-reads hit L2) and the 0.25 B/pixel map.
// No tensor cores: nothing here is a dense contraction.  The kernel is bounded by integer / packed-half issue rate
// (the alu pipe retires one warp instruction per two cycles), so the design goals are FEW INSTRUCTIONS PER PIXEL and
// MANY RESIDENT WARPS:
//
//   * a CTA is TWO warps that pipeline one 240-column strip of a band of rows through shared-memory rings:
//       warp A (producer): loads, gray, 5x5 Gaussian -> blur ring (+ the exact replay of the rare S % 159 == 0 pixels)
//       warp C (consumer): Sobel, candidate pre-filter, non-maximum suppression + double threshold, map rows -> global
//     They meet only at named barriers (bar.arrive / bar.sync, five ids): "block of 12 blur rows is final" (A -> C)
//     and "three blur rows consumed" (C -> A, four per block), so A fills rows of block t+1 right behind C reading
//     block t: one ring, no double buffer, both warps busy.  Same shared memory as one marching warp needed alone,
//     twice the resident warps, and neither warp carries the other's register state;
//   * lane l of either warp owns the 8 columns X0-8+8l .. +7 (lanes 0 and 31 are halo lanes); rows are marched top to
//     bottom so every gray / blur / Sobel row is computed once per band; horizontal neighbours come by shuffle;
//   * two pixels per 32-bit register everywhere: gray via dp4a (weights x4 so that >>6 becomes "take byte 1"); the
//     5x5 Gaussian as packed 16-bit integer sums: vertical pass from ROLLING 3-row box sums (kernel columns
//     (2,4,5,4,2) = 2t - c, (4,9,12,9,4) = 4t + q, (5,12,15,12,5) = 5t + 2q with t = box3(box3), q = g(y-1)+g(y+1):
//     six integer ops per pixel pair), horizontal combine, S/159 by multiply-high, remainder S - 159*q by ONE packed
//     multiply-add (zero 16-bit lane <=> S % 159 == 0); Sobel in exact fp16x2 arithmetic on integers stored as fp16
//     subnormals; candidates by the packed L1 norm |gx| + |gy| >= 2(low+1) (never misses a pixel with
//     gx^2 + gy^2 >= N_low; the exact test is redone per pixel in the sparse stage);
//   * the two data-dependent rarities are deferred to dense per-warp work lists instead of diverging in the hot loops:
//     (1) pixels with S % 159 == 0, where the reference's 25-step fp32 FMA chain can land just below the integer
//      -- replayed exactly once per block of 12 rows from a ring of gray rows; (2) 8-pixel groups that
//     hold a candidate -- only those get direction, non-maximum suppression and the double threshold, once per 2 rows
//     from a 4-row ring of Sobel sums;
//   * register roles rotate with period 3 in both warps (3 gray rows / box sums live in A, 3 rows of horizontal
//     differences / sums in C): the row loops are unrolled by 3 (A) and 6 (C), so no register-to-register moves;
//   * rows that touch the image border (zero padding is PER STAGE: gray, blur and gradient are each zero outside the
//     image, cannyEdgeD.cu:91-98,142-149,222-229) and strips whose last lane is only partly inside the image run the
//     CHECK instance of the row code; every other block of 12 rows runs without a single bounds test.
//
// Arithmetic contract: see k_stencil_tile.cuh (same results, bit for bit; tests compare both with the oracle).
#pragma once
#include "b2c_device.cuh"

#ifndef B2C_X
#define B2C_X 0   // experiment mask (profiling only): 1 = no NMS passes, 2 = no replay, 4 = warp C idles, 8 = warp A idles
#endif
namespace b2c
{
constexpr int MT_X = 240;            // output columns per strip
constexpr int MK = 12;               // blur rows per block
constexpr int M_GRING = 16;          // gray ring rows (bytes, 256 B per row): rows b0-2 .. b0+13 of a block
constexpr int M_SRING = 4;           // gx / gy ring rows (u16 pairs, 512 B per row each)
constexpr int M_OUTW = 16;           // map words per out-tile row (15 used)
constexpr int M_RCAP = 128;          // replay list capacity (per block of 12 rows; mean 19, flat pictures overflow)
constexpr int MS_GX = 0;                              // gx ring (2 KB; the slot arithmetic of the Sobel stage XORs bit 10)
constexpr int MS_GY = MS_GX + M_SRING * 512;
constexpr int MS_GRAY = MS_GY + M_SRING * 512;
constexpr int MS_BLUR = MS_GRAY + M_GRING * 256;      // blur ring: MK rows of 256 16-bit integers
constexpr int MS_OUT = MS_BLUR + MK * 512;
constexpr int MS_LISTA = MS_OUT + MK * M_OUTW * 4;
constexpr int MS_LISTC = MS_LISTA + M_RCAP * 2;
constexpr int MS_CNT = MS_LISTC + 64 * 2;
constexpr int MARCH_SMEM = MS_CNT + 16;
#ifndef B2C_MARCH_DUO
#define B2C_MARCH_DUO 0   // 1 = two warps per CTA (producer / consumer over named barriers), 0 = one warp does both roles block by block
#endif
constexpr bool MARCH_DUO = B2C_MARCH_DUO != 0;
constexpr int MARCH_THREADS = MARCH_DUO ? 64 : 32;
#ifndef B2C_MARCH_CTAS
#define B2C_MARCH_CTAS 14
#endif
constexpr int MARCH_CTAS_PER_SM = B2C_MARCH_CTAS;
// Hand-over between the two warps: THREE hardware named barriers (bar.arrive by the signalling warp, bar.sync by the
// waiting one: a blocked warp costs no issue slot and wakes within tens of cycles).  Measured alternatives: six named
// barriers limit the SM to 10 resident CTAs (an SM has 64; launch__occupancy_limit_barriers), mbarriers in shared memory
// cost either 25 % of all issue slots (try_wait returns after ~20 cycles, so waiting is a spin loop) or, with a
// nanosleep between polls, microseconds of wake-up latency per hand-over.  Hence only three ids:
//   MB_FULL (0, also the start-up barrier): "block t of the blur ring is final" (A -> C);
//   MB_EMPTY0 + h (1, 2): "rows 6h .. 6h+5 of the blur ring have been read" (C -> A).
constexpr int MB_FULL = 0, MB_EMPTY0 = 1;

#ifdef B2C_EMU
static inline void m_bar_sync(int id) { emu::named_bar_sync(id, MARCH_THREADS); }
static inline void m_bar_arrive(int id) { emu::named_bar_arrive(id, MARCH_THREADS); }
static inline uint32_t m_h2absadd(uint32_t a, uint32_t b)
{
  const float al = fabsf((float)emu_h_lo(a)), ah = fabsf((float)emu_h_hi(a)), bl = fabsf((float)emu_h_lo(b)), bh = fabsf((float)emu_h_hi(b));
  return emu_pack_h2((_Float16)(al + bl), (_Float16)(ah + bh));
}
static inline uint32_t m_h2max(uint32_t a, uint32_t b)
{
  return emu_pack_h2((float)emu_h_lo(a) > (float)emu_h_lo(b) ? emu_h_lo(a) : emu_h_lo(b), (float)emu_h_hi(a) > (float)emu_h_hi(b) ? emu_h_hi(a) : emu_h_hi(b));
}
static inline bool m_h2any_ge(uint32_t a, uint32_t b) { return (float)emu_h_lo(a) >= (float)emu_h_lo(b) || (float)emu_h_hi(a) >= (float)emu_h_hi(b); }
static inline uint32_t m_h2mul(uint32_t a, uint32_t b) { return emu_pack_h2((_Float16)((float)emu_h_lo(a) * (float)emu_h_lo(b)), (_Float16)((float)emu_h_hi(a) * (float)emu_h_hi(b))); }
static inline uint32_t m_h2sq2(uint32_t x, uint32_t y)   // fl(y*y + fl(x*x)) on both halves
{
  const uint32_t t = m_h2mul(x, x);
  return emu_pack_h2((_Float16)((double)emu_h_lo(y) * (double)emu_h_lo(y) + (double)emu_h_lo(t)), (_Float16)((double)emu_h_hi(y) * (double)emu_h_hi(y) + (double)emu_h_hi(t)));
}
#else
// (immediate barrier ids: with an id in a register ptxas reserves all 16 named barriers for the CTA)
__device__ __forceinline__ void m_bar_sync(int id)
{
  if (id == 0) asm volatile("bar.sync 0, %0;" ::"n"(MARCH_THREADS) : "memory");
  else if (id == 1) asm volatile("bar.sync 1, %0;" ::"n"(MARCH_THREADS) : "memory");
  else asm volatile("bar.sync 2, %0;" ::"n"(MARCH_THREADS) : "memory");
}
__device__ __forceinline__ void m_bar_arrive(int id)
{
  if (id == 0) asm volatile("bar.arrive 0, %0;" ::"n"(MARCH_THREADS) : "memory");
  else if (id == 1) asm volatile("bar.arrive 1, %0;" ::"n"(MARCH_THREADS) : "memory");
  else asm volatile("bar.arrive 2, %0;" ::"n"(MARCH_THREADS) : "memory");
}
// |a| + |b| on both halves: ONE instruction (HADD2 |a|, |b|)
__device__ __forceinline__ uint32_t m_h2absadd(uint32_t a, uint32_t b)
{
  uint32_t r;
  asm("{.reg .b32 x, y; abs.f16x2 x, %1; abs.f16x2 y, %2; add.rn.f16x2 %0, x, y;}" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t m_h2max(uint32_t a, uint32_t b) { uint32_t r; asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t m_h2mul(uint32_t a, uint32_t b) { uint32_t r; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t m_h2sq2(uint32_t x, uint32_t y)   // fl(y*y + fl(x*x)) on both halves
{
  uint32_t r;
  asm("{.reg .b32 t; mul.rn.f16x2 t, %1, %1; fma.rn.f16x2 %0, %2, %2, t;}" : "=r"(r) : "r"(x), "r"(y));
  return r;
}
__device__ __forceinline__ bool m_h2any_ge(uint32_t a, uint32_t b)
{
  uint32_t r;
  asm("{.reg .pred p, q; setp.ge.f16x2 p|q, %1, %2; or.pred p, p, q; selp.u32 %0, 1, 0, p;}" : "=r"(r) : "r"(a), "r"(b));
  return r != 0;
}
#endif

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
// 0, then truncation (src/cvp/cannyEdgeD.cu:102-115).  b = band-local blur row (gray row y lives in ring slot y & 15),
// rr = the blur row's slot in the blur ring, col = window column.
__device__ __forceinline__ void m_gauss_replay(const B2cStencilParams &p, char *smem, int b, int rr, int col)
{
  const uint8_t *G = reinterpret_cast<const uint8_t *>(smem + MS_GRAY) + col - 2;
  float f = 0.0f;
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const uint8_t *row = G + ((b - 2 + r) & (M_GRING - 1)) * 256;
#pragma unroll
    for (int c = 0; c < 5; ++c) f = __fmaf_rn(p.gk[r * 5 + c], (float)row[c], f);
  }
  reinterpret_cast<uint16_t *>(smem + MS_BLUR)[rr * 256 + col] = (uint16_t)(unsigned)f;   // fp16 subnormal = the integer itself
}

// Direction, non-maximum suppression and double threshold for one pixel (src/cvp/cannyEdgeD.cu:196, 239-267, 290),
// in exact fp32 on the integer Sobel sums: sector from 2|gx*gy| vs |gy^2-gx^2| (== the atan2 sectors, pinned
// exhaustively in tests), keep iff both neighbours along it have N <= N (ties kept).  BRANCH-FREE (every lane runs the
// whole body, the result is masked by `valid`): the body is a chain of shared-memory round trips and dependent fp32 ops,
// and straight-line code lets the compiler overlap several of these chains with each other and with the Sobel rows
// that follow in the same basic block.  gxr = the gx ring (the gy ring lies MS_GY - MS_GX behind it); cb / ub / db =
// byte offsets of the ring rows n, n-1, n+1; c2 = 2 * window column.  Returns 1 = strong, 2 = weak, 0 = neither.
__device__ __forceinline__ uint32_t m_nms_item(const B2cStencilParams &p, const char *gxr, const uint32_t cb, const uint32_t ub, const uint32_t db, const uint32_t c2, const bool valid)
{
  const uint32_t ctr = cb + c2;
  const uint32_t cx = *reinterpret_cast<const uint16_t *>(gxr + ctr), cy = *reinterpret_cast<const uint16_t *>(gxr + (MS_GY - MS_GX) + ctr);
  const float nx = b2c_fhfma_ll(cx, cx, 0.0f), ny = b2c_fhfma_ll(cy, cy, 0.0f), pr = b2c_fhfma_ll(cx, cy, 0.0f);
  const float nn = nx + ny, d = ny - nx, a2 = fabsf(pr) + fabsf(pr);
  // neighbours q / r: sector 0: rows +-1;  sector 2: columns +-1;  sector 1: (y+1,x-1),(y-1,x+1);  sector 3: (y-1,x-1),(y+1,x+1)
  const bool axis = a2 < fabsf(d), vert = d > 0.0f, diag1 = pr > 0.0f;
  const bool rows = !axis || vert;                                   // the neighbours lie in the rows n+1 / n-1
  const int dx2 = axis ? (vert ? 0 : 2) : (diag1 ? -2 : 2);          // 2 * column step of q
  const uint32_t oq = (rows ? db : cb) + c2 + dx2, orr = (rows ? ub : cb) + c2 - dx2;
  const uint32_t qx = *reinterpret_cast<const uint16_t *>(gxr + oq), qy = *reinterpret_cast<const uint16_t *>(gxr + (MS_GY - MS_GX) + oq);
  const uint32_t rx = *reinterpret_cast<const uint16_t *>(gxr + orr), ry = *reinterpret_cast<const uint16_t *>(gxr + (MS_GY - MS_GX) + orr);
  const float nq = b2c_fhfma_ll(qx, qx, b2c_fhfma_ll(qy, qy, 0.0f)), nr = b2c_fhfma_ll(rx, rx, b2c_fhfma_ll(ry, ry, 0.0f));
  // the packed pre-filter is conservative: nn >= n_lo[0] is the exact candidate test (N >= N_low)
  const bool keep = valid & (nn >= p.n_lo[0]) & (nq <= nn) & (nr <= nn);
  // double threshold on v = (unsigned char) trunc(grad): v > T  <=>  N in [n[0], wrap0) or [n[1], wrap1) or [n[2], inf)
  // (bitwise logic on purpose: && / ?: would come back as branches)
  const bool w0 = nn < p.n_wrap[0], w1 = nn < p.n_wrap[1];
  const bool strong = (w0 & (nn >= p.n_hi[0])) | (!w0 & (((nn >= p.n_hi[1]) & w1) | (nn >= p.n_hi[2])));
  const bool cand = w0 | ((nn >= p.n_lo[1]) & w1) | (nn >= p.n_lo[2]);
  return (keep & strong ? 1u : 0u) | (keep & !strong & cand ? 2u : 0u);
}

// One pass over (up to) 4 H entries of an NMS work list: 8 lanes -- one per pixel -- take an entry, every lane takes H
// entries (q0 + 4h + lane/8) with all H dependency chains in flight together.  K = row within the group
// of six at which the pass runs (static), sflip = (it << 10): the ring slot of Sobel row s is (K + 1 + 2 it) & 3, i.e.
// the byte offset ((K + 1) & 3) * 512 with bit 10 flipped for the second group of six.  The list holds the NMS rows
// n = s-2 + up.
template <int K, int H>
__device__ __forceinline__ void m_nms_pass(const B2cStencilParams &p, char *smem, const uint16_t *list, const int cnt, const int q0, const int rr, const uint32_t sflip)
{
  const int lane = threadIdx.x & 31, px = lane & 7, sub = lane >> 3;
  const char *gxr = smem + MS_GX;
  uint8_t *o8 = reinterpret_cast<uint8_t *>(smem + MS_OUT);
  // rows s-3, s-2, s-1, s
  const uint32_t R3 = (uint32_t)(((K + 2) & 3) * 512) ^ sflip, R2 = (uint32_t)(((K + 3) & 3) * 512) ^ sflip, R1 = (uint32_t)((K & 3) * 512) ^ sflip, R0 = (uint32_t)(((K + 1) & 3) * 512) ^ sflip;
  uint32_t res[H], e[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const int q = q0 + 4 * h + sub;
    const bool valid = q < cnt;
    e[h] = valid ? list[q] : 1u;   // (any in-range entry for the idle lanes)
    const bool up = (e[h] >> 8) != 0u;
    res[h] = m_nms_item(p, gxr, up ? R1 : R2, up ? R2 : R3, up ? R0 : R1, ((e[h] & 31u) * 8u + px) * 2u, valid);
  }
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const uint32_t bs = __ballot_sync(B2C_FULL, res[h] & 1u), bw = __ballot_sync(B2C_FULL, res[h] & 2u);
    if (px == 0 && q0 + 4 * h + sub < cnt) {   // one lane per entry stores the group's strong and weak byte: out-tile row rr-1 or rr, byte = lane - 1
      uint8_t *o = o8 + (rr - 1 + (int)(e[h] >> 8)) * 64 + (e[h] & 31u) - 1u;
      o[0] = (uint8_t)(bs >> (8 * sub));
      o[32] = (uint8_t)(bw >> (8 * sub));
    }
  }
}

struct MarchGeo {
  int lane, X0, Y0, frame, rows_out, xl, yg0, ilim, nblocks;
  bool lane_in, out_lane;
  uint32_t pm[4];   // pixel-pair masks of this lane (a lane that is only partly inside the image, w % 8 != 0)
  bool partial;     // this strip holds such a lane
};

// ---- warp A: gray + Gaussian + replay --------------------------------------------------------------------------------
// CH = bytes per input pixel: 3 = BGR8 (the reference's format), 4 = BGRA8 (alpha ignored), 1 = GRAY8 (gray = the byte;
// the reference's own CV_8UC1 path is broken, SURVEY T13 -- this is what it evidently meant to do)
template <int CH>
__device__ __forceinline__ void m_gray_row(const uint2 (&d)[CH], uint32_t (&m)[4], uint2 &bytes)
{
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
}

template <int CH>
struct MarchA {
  uint32_t G[3][4];    // gray rows i-2, i-1, i (slot = row % 3)
  uint32_t BX[3][4];   // 3-row box sums bx(r) = g(r-1) + g(r) + g(r+1), rows b-1, b, b+1
  uint2 pre[2][CH];    // raw pixels of the next TWO gray rows, in flight while the current row is processed (one row of
                       // this warp's work is shorter than an L2 round trip); slot = row parity within the group of 3 ... see m_a_rows3
  const uint8_t *np;   // address of the row after `pre` for this lane
  uint32_t gaddr;      // byte offset of this lane's 8 gray bytes in the ring slot of the next gray row
};

// three consecutive rows 3j .. 3j+2 of a block (j = 0..3 at run time, the row within the group is static):
// gray row b+2, Gaussian of blur row b.  zw collects the replay flags of the group: bit 7 + 8*k + px.
template <int CH, bool CHECK>
__device__ __forceinline__ void m_a_rows3(const B2cStencilParams &p, const MarchGeo &g, char *smem, MarchA<CH> &a, const long long lstride, const int b_first, char *blur_rows, uint32_t &zw)
{
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int b = b_first + k, i = b + 2;   // blur row, gray row (band-local)
    uint2 cur[CH];
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      cur[q] = a.pre[0][q];
      a.pre[0][q] = a.pre[1][q];   // (register renaming after unrolling: the group of 3 rows is followed by a swap-free rotation
    }                               //  only if the compiler sees through it; it costs at most CH moves per row)
    {   // raw pixels of gray row i+2; zero outside the image (cannyEdgeD.cu:91-98)
      const uint8_t *src = a.np;
      if (CHECK) {
        const int yg = g.yg0 + i + 2;
        if (!(yg >= 0 && yg < p.h_glob && i + 2 < g.ilim)) src = p.zeros;   // warp-uniform
      }
      const uint2 *q2 = reinterpret_cast<const uint2 *>(src);
#pragma unroll
      for (int q = 0; q < CH; ++q) a.pre[1][q] = (B2C_X & 16) ? make_uint2(a.pre[1][q].y * 3u + 1u, a.pre[1][q].x ^ 0x55u) : __ldg(q2 + q);   // (16: no loads, profiling only)
      a.np += lstride;
    }
    uint32_t(&gN)[4] = a.G[k];
    uint2 bytes;
    m_gray_row<CH>(cur, gN, bytes);
    if (CHECK && g.partial) {
#pragma unroll
      for (int j = 0; j < 4; ++j) gN[j] &= g.pm[j];
      bytes = make_uint2(__byte_perm(gN[0], gN[1], 0x6420), __byte_perm(gN[2], gN[3], 0x6420));
    }
    *reinterpret_cast<uint2 *>(smem + MS_GRAY + a.gaddr) = bytes;
    a.gaddr = (a.gaddr + 256u) & (M_GRING * 256 - 1);
    const uint32_t(&gA)[4] = a.G[(k + 1) % 3], (&gB)[4] = a.G[(k + 2) % 3];          // g(b), g(b+1)
    const uint32_t(&bxP)[4] = a.BX[(k + 1) % 3], (&bxC)[4] = a.BX[(k + 2) % 3];      // bx(b-1), bx(b)
    uint32_t(&bxN)[4] = a.BX[k];                                                    // bx(b+1)
    uint32_t v0[6], v1[6], v2[4];   // index j+1 = pair j; [0] / [5] come from the neighbouring lanes
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // vertical pass: t = g(b-2) + 2g(b-1) + 3g(b) + 2g(b+1) + g(b+2), q = g(b-1) + g(b+1), c = g(b):
      // kernel columns (2,4,5,4,2) = 2t - c, (4,9,12,9,4) = 4t + q, (5,12,15,12,5) = 5t + 2q
      bxN[j] = gA[j] + gB[j] + gN[j];
      const uint32_t tt = bxP[j] + bxC[j] + bxN[j];
      const uint32_t q = bxC[j] - gA[j];
      v0[j + 1] = 2u * tt - gA[j];
      v1[j + 1] = 4u * tt + q;
      v2[j] = v1[j + 1] + tt + q;
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
      // S = sum k*gray (exact, <= 40545) per 16-bit lane; q = S/159 via multiply-high; remainder lane == 0 -> replay
      const uint32_t S = (v0[j] + v0[j + 2] + v2[j]) + (o1[j] + o1[j + 1]);
      const uint32_t U = __umulhi(S, 27012373u);         // upper 16 bits = high lane / 159 (the low lane adds < 1/159)
      const uint32_t V = __umulhi(S << 16, 27012373u);   // low lane
      // (q_lo, q_hi) as 16-bit integers == fp16 SUBNORMALS q * 2^-24: every later fp16 value is an integer
      // multiple of 2^-24 below 2^-13, so the fp16x2 arithmetic stays exact and needs no int->half conversion
      hq[j] = __byte_perm(V, U, 0x7632);
      rem[j] = S - 159u * hq[j];                         // both lanes at once: 159*q <= 40545 never carries
    }
    if (CHECK && g.partial) {
#pragma unroll
      for (int j = 0; j < 4; ++j) hq[j] &= g.pm[j];      // blur is zero outside the image (cannyEdgeD.cu:142-149)
    }
    // remainders are < 159: a zero BYTE 0 / 2 marks S % 159 == 0.  Bit 7 of each byte of z01 / z23 <=> pixel
    // k = byte (+4) needs the replay (a zero byte can set the bit of the byte above it too: harmless, the
    // replay is exact for every pixel)
    const uint32_t r01 = __byte_perm(rem[0], rem[1], 0x6420), r23 = __byte_perm(rem[2], rem[3], 0x6420);
    const uint32_t z01 = (r01 - 0x01010101u) & ~r01 & 0x80808080u, z23 = (r23 - 0x01010101u) & ~r23 & 0x80808080u;
    // the 0x80 marker bytes -> the row's 8-bit pixel mask at bits 7..14 (two dp4a: fma pipe, no shifts)
    uint32_t zb = __dp4a(z23, 0x80402010u, __dp4a(z01, 0x08040201u, 0u));
    uint4 *bdst = reinterpret_cast<uint4 *>(blur_rows + k * 512);
    bool store = g.lane_in;
    if (CHECK) {
      const int yb = g.yg0 + b;
      if (!(yb >= 0 && yb < p.h_glob)) {   // warp-uniform: blur is zero above / below the image
        hq[0] = hq[1] = hq[2] = hq[3] = 0u;
        zb = 0u;
        store = true;
      }
    }
    if (store) *bdst = make_uint4(hq[0], hq[1], hq[2], hq[3]);
    zw += zb << (8 * k);   // rows k = 0, 1, 2 of the group at bits 7.., 15.., 23..
  }
}

template <int CH>
struct MarchAState {
  MarchA<CH> a;
  long long lstride;   // bytes between rows for this lane (0 for lanes outside the image: they read zeros)
  uint32_t zkeep;      // which replay flags of this lane count
};

// gray rows -4 .. -1 of the band, box sums, first row in flight
template <int CH>
__device__ __forceinline__ void m_a_init(const B2cStencilParams &p, const MarchGeo &g, char *smem, MarchAState<CH> &st)
{
  const int lane = g.lane;
  // which of this lane's 8 columns can feed an output pixel: window columns 6..249 for the blur (lane 0: px 6,7;
  // lane 31: px 0,1), bits laid out like the flag words of a group (bit 7 + 8*row + px)
  uint32_t zpat = !g.lane_in ? 0u : lane == 0 ? 0xC0u : lane == 31 ? 0x03u : 0xFFu;
  if (g.partial) {   // pixels of a partly covered lane that lie outside the image are never replayed
    uint32_t m = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if ((g.pm[k >> 1] >> (16 * (k & 1))) & 1u) m |= 1u << k;
    zpat &= m;
  }
  st.zkeep = (zpat | (zpat << 8) | (zpat << 16)) << 7;
  const long long lstride = st.lstride = g.lane_in ? p.row_stride : 0;
  const uint8_t *lp = g.lane_in ? p.bgr + (long long)g.frame * p.frame_stride + CH * g.xl + (long long)g.Y0 * p.row_stride : p.zeros;

  MarchA<CH> &a = st.a;
  // ---- prologue: gray rows -4 .. -1, box sums bx(-3), bx(-2); raw pixels of gray row 0 in flight -----------------
  {
    uint32_t gm[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = -4 + k, yg = g.yg0 + i;
      uint2 d[CH];
      if (yg >= 0 && yg < p.h_glob && i < g.ilim) {   // warp-uniform
        const uint2 *q = reinterpret_cast<const uint2 *>(lp + (long long)i * lstride);
#pragma unroll
        for (int c = 0; c < CH; ++c) d[c] = __ldg(q + c);
      } else {
#pragma unroll
        for (int c = 0; c < CH; ++c) d[c] = make_uint2(0u, 0u);
      }
      uint2 bytes;
      m_gray_row<CH>(d, gm[k], bytes);
      if (g.partial) {
#pragma unroll
        for (int j = 0; j < 4; ++j) gm[k][j] &= g.pm[j];
        bytes = make_uint2(__byte_perm(gm[k][0], gm[k][1], 0x6420), __byte_perm(gm[k][2], gm[k][3], 0x6420));
      }
      *reinterpret_cast<uint2 *>(smem + MS_GRAY + ((i + 16) & (M_GRING - 1)) * 256 + lane * 8) = bytes;
    }
    // block 0, row rr = 0 computes gray row i = 0 into G[0] and expects g(-2) in G[1], g(-1) in G[2], bx(-3) in BX[1], bx(-2) in BX[2]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a.G[0][j] = 0u;
      a.G[1][j] = gm[2][j];
      a.G[2][j] = gm[3][j];
      a.BX[0][j] = 0u;
      a.BX[1][j] = gm[0][j] + gm[1][j] + gm[2][j];
      a.BX[2][j] = gm[1][j] + gm[2][j] + gm[3][j];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {   // gray rows 0 and 1 in flight
      const int yg = g.yg0 + r;
      const uint8_t *src = (yg >= 0 && yg < p.h_glob && r < g.ilim) ? lp + r * lstride : p.zeros;
      const uint2 *q = reinterpret_cast<const uint2 *>(src);
#pragma unroll
      for (int c = 0; c < CH; ++c) a.pre[r][c] = __ldg(q + c);
    }
    a.np = lp + 2 * lstride;
    a.gaddr = (uint32_t)lane * 8u;   // gray row 0 -> ring slot 0
  }
}

// block t of warp A's work: 12 gray / blur rows into the rings, then the exact replay of the block's S % 159 == 0 pixels
template <int CH, bool DUO>
__device__ __forceinline__ void m_a_block(const B2cStencilParams &p, const MarchGeo &g, char *smem, MarchAState<CH> &st, const int t)
{
  const int lane = g.lane;
  uint16_t *list = reinterpret_cast<uint16_t *>(smem + MS_LISTA);
  int *cnt_s = reinterpret_cast<int *>(smem + MS_CNT);
  MarchA<CH> &a = st.a;
  const long long lstride = st.lstride;
  const uint32_t zkeep = st.zkeep;
  {
    const int b0 = -2 + MK * t;   // first blur row of this block
#if !defined(B2C_EMU) && !defined(B2C_MARCH_NO_L2PF)
    // pull the gray rows of the NEXT block (768 B per row and strip = 7 lines of 128 B) into L2 while this block
    // computes: 4 rows per instruction (lanes 7q .. 7q+6 take row q), 3 instructions for 12 rows
    if (CH == 3 && lane < 28) {   // (only for the 3-byte format: 7 lines per strip row)
      const int q = lane / 7;
      const long long xoff = (long long)(g.X0 - 8) * 3 + (lane - 7 * q) * 128;
      if (xoff >= 0 && xoff < (long long)p.w * 3) {
        const uint8_t *pf = p.bgr + (long long)g.frame * p.frame_stride + (long long)g.Y0 * p.row_stride + xoff;
#pragma unroll
        for (int it = 0; it < 3; ++it) {
          const int i = b0 + MK + 3 + 4 * it + q, yg = g.yg0 + i;
          if (yg >= 0 && yg < p.h_glob && i < g.ilim) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (long long)i * p.row_stride));
        }
      }
    }
#endif
    // this block touches only rows inside the image (gray rows up to i = b0+15 incl. the two in flight, blur rows
    // b0 .. b0+11) and the strip has no partly covered lane: no bounds test at all
    const bool fast = !g.partial && g.yg0 + b0 >= 0 && g.yg0 + b0 + MK + 3 < p.h_glob && b0 + MK + 3 < g.ilim;
    uint32_t zacc[4] = { 0u, 0u, 0u, 0u };   // replay flags of the four groups of 3 rows
#pragma unroll 1
    for (int j = 0; j < MK / 3; ++j) {
      // rows 3j .. 3j+5 of the blur ring are free once warp C has read them for block t-1
      if (DUO && t > 0 && (j & 1) == 0) m_bar_sync(MB_EMPTY0 + (j >> 1));
      uint32_t zw = 0u;
      char *blur_rows = smem + MS_BLUR + j * 1536 + lane * 16;
      if (fast) m_a_rows3<CH, false>(p, g, smem, a, lstride, b0 + 3 * j, blur_rows, zw);
      else m_a_rows3<CH, true>(p, g, smem, a, lstride, b0 + 3 * j, blur_rows, zw);
      if (j == 0) zacc[0] = zw;
      else if (j == 1) zacc[1] = zw;
      else if (j == 2) zacc[2] = zw;
      else zacc[3] = zw;
    }
    __syncwarp();

    // ---- exact replay of the S % 159 == 0 pixels of the block ------------------------------------------------------
    {
#pragma unroll
      for (int wi = 0; wi < 4; ++wi) zacc[wi] &= zkeep;
      const int mine = (B2C_X & 2) ? 0 : __popc(zacc[0]) + __popc(zacc[1]) + __popc(zacc[2]) + __popc(zacc[3]);
      int off = 0;
      if (mine) off = atomicAdd(cnt_s, mine);   // order within the list does not matter: every entry is replayed
      __syncwarp();
      const int cnt = *cnt_s;
      __syncwarp();
      if (lane == 0) *cnt_s = 0;
      if (cnt <= M_RCAP) {
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
          uint32_t bits = zacc[wi];
          while (bits) {
            const int pb = __ffs((int)bits) - 1;
            bits &= bits - 1u;
            // bit pb: row = 3 * wi + ((pb - 7) >> 3), pixel = (pb - 7) & 7
            list[off++] = (uint16_t)(((3 * wi + ((pb - 7) >> 3)) << 8) | (lane * 8 + ((pb - 7) & 7)));
          }
        }
        __syncwarp();
        for (int i = lane; i < cnt; i += 32) {
          const int e = list[i], rr = e >> 8;
          m_gauss_replay(p, smem, b0 + rr, rr, e & 255);
        }
      } else {   // flat pictures: (almost) every pixel takes the replay -- do all of them, the replay is exact everywhere
        for (int i = lane; i < MK * 244; i += 32) {
          const int rr = i / 244, col = 6 + i % 244;
          const int yb = g.yg0 + b0 + rr, xg = g.X0 - 8 + col;
          if (yb >= 0 && yb < p.h_glob && xg >= 0 && xg < p.w) m_gauss_replay(p, smem, b0 + rr, rr, col);
        }
      }
      __syncwarp();
    }
    if (DUO) m_bar_arrive(MB_FULL);   // block t of the blur ring is final
  }
}

// ---- warp C: Sobel + candidate pre-filter + NMS / double threshold + output -------------------------------------
struct MarchC {
  uint32_t D[3][4];   // horizontal differences D(r) = blur(x+1) - blur(x-1) of blur rows r-2, r-1, r (slot = row % 3)
  uint32_t T[3][4];   // horizontal sums T(r) = blur(x-1) + 2 blur(x) + blur(x+1)
};

// Row k (0..5, static) of a group of six blur rows 6it .. 6it+5 of a block: Sobel row s = b0-1+rr, rr = 6it+k.  After
// an odd row the NMS work list of the two NMS rows s-2, s-1 is built (two ballots) and worked off.  chk (warp-uniform):
// some row may lie outside the image or the strip has a partly covered lane.
struct MarchCList {
  uint32_t flag_carry;   // candidate flag of the last Sobel row of the previous pair
};
template <int K>
__device__ __forceinline__ void m_c_row(const B2cStencilParams &p, const MarchGeo &g, char *smem, MarchC &c, const int b0, const int it, const bool chk, const uint32_t thr2,
                                        MarchCList &L, uint32_t &flag_even, const bool signal)
{
  constexpr int k = K;
  const int lane = g.lane;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const char *blur_rows = smem + MS_BLUR + it * (6 * 512) + lane * 16;
  // ring slot of Sobel row s = (s + 4) & 3 = (rr + 1) & 3 = (k + 1 + 2 it) & 3: the slot's address bits 9-10 are
  // those of the first group XOR (it << 10) (MS_GX and MS_GY have these bits clear)
  static_assert((MS_GX & 0x600) == 0 && (MS_GY & 0x600) == 0, "gx / gy rings must start at a multiple of 2 KB");
  const uint32_t sflip = (uint32_t)it << 10;
  {
    const int rr = 6 * it + k;    // rr % 3 == k % 3
    const int s = b0 - 1 + rr;    // Sobel row (band-local); its newest blur row is r = s + 1 = b0 + rr
    // blur as 16-bit integers == fp16 subnormals q * 2^-24; x 4096 (exact) makes them the normal numbers q * 2^-12, on
    // which the packed-half arithmetic below is still exact integer arithmetic (|values| <= 2040 < 2^11) AND squares
    // do not underflow: the candidate pre-filter can use N' = gx'^2 + gy'^2 in fp16
    const uint4 bw = *reinterpret_cast<const uint4 *>(blur_rows + k * 512);
    const uint32_t b1 = m_h2mul(bw.x, 0x6C006C00u), b2 = m_h2mul(bw.y, 0x6C006C00u), b3 = m_h2mul(bw.z, 0x6C006C00u), b4 = m_h2mul(bw.w, 0x6C006C00u);
    const uint32_t B[6] = { __shfl_up_sync(B2C_FULL, b4, 1), b1, b2, b3, b4, __shfl_down_sync(B2C_FULL, b1, 1) };
    if (signal && k == 5) m_bar_arrive(MB_EMPTY0 + it);   // rows 6it .. 6it+5 have been read: warp A may refill them
    uint32_t O[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) O[j] = __byte_perm(B[j], B[j + 1], 0x5432);   // (blur[2j-1], blur[2j])
    uint32_t gx[4], gy[4];
    uint32_t(&Dn)[4] = c.D[k % 3], (&Tn)[4] = c.T[k % 3];
    const uint32_t(&D1)[4] = c.D[(k + 2) % 3], (&D2)[4] = c.D[(k + 1) % 3], (&T2)[4] = c.T[(k + 1) % 3];   // rows r-1, r-2
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      Dn[j] = b2c_h2sub(O[j + 1], O[j]);                           // blur(x+1) - blur(x-1)        (cannyEdgeD.cu:158-160)
      Tn[j] = b2c_h2add(b2c_h2fma2(B[j + 1], O[j]), O[j + 1]);     // blur(x-1)+2blur(x)+blur(x+1) (:164-166)
      gx[j] = b2c_h2add(b2c_h2fma2(D1[j], D2[j]), Dn[j]);          // sumX(s) = D(s-1) + 2 D(s) + D(s+1): right - left
      gy[j] = b2c_h2sub(T2[j], Tn[j]);                             // sumY(s) = T(s-1) - T(s+1): top - bottom
    }
    const uint32_t so = ((uint32_t)(MS_GX + ((k + 1) & (M_SRING - 1)) * 512) ^ sflip) + lane * 16;
    if (g.lane_in) {
      *reinterpret_cast<uint4 *>(smem + so) = make_uint4(gx[0], gx[1], gx[2], gx[3]);
      *reinterpret_cast<uint4 *>(smem + so + (MS_GY - MS_GX)) = make_uint4(gy[0], gy[1], gy[2], gy[3]);
    }
    // candidate pre-filter: N' = gx^2 + gy^2 in packed fp16 (relative error < 2^-10) against a threshold lowered by that
    // much: no pixel with N >= N_low is missed, almost none below it passes
    bool f = m_h2any_ge(m_h2max(m_h2max(m_h2sq2(gx[0], gy[0]), m_h2sq2(gx[1], gy[1])), m_h2max(m_h2sq2(gx[2], gy[2]), m_h2sq2(gx[3], gy[3]))), thr2);
    if (chk) {   // warp-uniform and rare: the stores above are simply redone (no register merge on the fast path)
      const int ys = g.yg0 + s;
      if (!(ys >= 0 && ys < p.h_glob)) {   // the gradient is zero outside the image (cannyEdgeD.cu:222-229)
        *reinterpret_cast<uint4 *>(smem + so) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(smem + so + (MS_GY - MS_GX)) = make_uint4(0u, 0u, 0u, 0u);
        f = false;
      } else if (g.partial) {
        uint32_t mx[4], my[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { mx[j] = gx[j] & g.pm[j]; my[j] = gy[j] & g.pm[j]; }
        if (g.lane_in) {
          *reinterpret_cast<uint4 *>(smem + so) = make_uint4(mx[0], mx[1], mx[2], mx[3]);
          *reinterpret_cast<uint4 *>(smem + so + (MS_GY - MS_GX)) = make_uint4(my[0], my[1], my[2], my[3]);
        }
        f = m_h2any_ge(m_h2max(m_h2max(m_h2sq2(mx[0], my[0]), m_h2sq2(mx[1], my[1])), m_h2max(m_h2sq2(mx[2], my[2]), m_h2sq2(mx[3], my[3]))), thr2);
      }
    }
    const uint32_t fl = (f && g.out_lane) ? 1u : 0u;
    if ((k & 1) == 0) {
      flag_even = fl;
    } else {
      // NMS rows n = s-2 (flag carried over from the previous pair of rows) and n = s-1.  Work items are the
      // (row, lane) groups of 8 pixels that may hold a candidate: two ballots build the list, then 8 lanes -- one per
      // pixel -- take each item, eight items per pass.  Out-tile row of NMS row n: n - (b0 - 2), i.e. rr-1 and rr.
      uint16_t *list = reinterpret_cast<uint16_t *>(smem + MS_LISTC);
      __syncwarp();
      const uint32_t m0 = __ballot_sync(B2C_FULL, L.flag_carry != 0u), m1 = __ballot_sync(B2C_FULL, flag_even != 0u);
      const int n0 = __popc(m0), cnt = (B2C_X & 1) ? 0 : n0 + __popc(m1);
      if (cnt) {   // warp-uniform
        if (L.flag_carry) list[__popc(m0 & lt_mask)] = (uint16_t)lane;
        if (flag_even) list[n0 + __popc(m1 & lt_mask)] = (uint16_t)(0x100u | lane);
        __syncwarp();
        if (!(B2C_X & 32)) {
          // 8 entries per pass, 2 per lane: both dependency chains of a lane are in flight together.  (4 per lane was
          // measured: 128 registers, 58 KB of code, 1.4x slower.)
          for (int q0 = 0; q0 < cnt; q0 += 8) m_nms_pass<K, 2>(p, smem, list, cnt, q0, rr, sflip);
        }
        __syncwarp();
      }
      L.flag_carry = fl;
    }
  }
}
__device__ __forceinline__ void m_c_rows6(const B2cStencilParams &p, const MarchGeo &g, char *smem, MarchC &c, const int b0, const int it, const bool chk, const uint32_t thr2,
                                          MarchCList &L, const bool signal)
{
  uint32_t fe = 0u;
  m_c_row<0>(p, g, smem, c, b0, it, chk, thr2, L, fe, signal);
  m_c_row<1>(p, g, smem, c, b0, it, chk, thr2, L, fe, signal);
  m_c_row<2>(p, g, smem, c, b0, it, chk, thr2, L, fe, signal);
  m_c_row<3>(p, g, smem, c, b0, it, chk, thr2, L, fe, signal);
  m_c_row<4>(p, g, smem, c, b0, it, chk, thr2, L, fe, signal);
  m_c_row<5>(p, g, smem, c, b0, it, chk, thr2, L, fe, signal);
}

struct MarchCState {
  MarchC c;
  MarchCList L;
  uint32_t thr2;
};
__device__ __forceinline__ void m_c_init(const B2cStencilParams &p, MarchCState &st)
{
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) st.c.D[k][j] = st.c.T[k][j] = 0u;
  st.L.flag_carry = 0u;
  st.thr2 = p.n_pre;   // fp16x2: the conservative candidate threshold on N' (see b2c_fill_thresholds)
}

// block t of warp C's work: Sobel rows b0-1 .. b0+10, NMS rows b0-2 .. b0+9, the 12 finished map rows -> global
template <bool DUO>
__device__ __forceinline__ void m_c_block(const B2cStencilParams &p, const MarchGeo &g, char *smem, MarchCState &st, const int t)
{
  const int lane = g.lane;
  {
    const int b0 = -2 + MK * t;
    if (DUO) m_bar_sync(MB_FULL);   // block t of the blur ring is final
    // Sobel rows b0-1 .. b0+10 all inside the image and no partly covered lane: nothing to test
    const bool chk = g.partial || g.yg0 + b0 - 1 < 0 || g.yg0 + b0 + MK - 2 >= p.h_glob;
    const bool signal = DUO && t + 1 < g.nblocks;
#pragma unroll 1
    for (int it = 0; it < 2; ++it) m_c_rows6(p, g, smem, st.c, b0, it, chk, st.thr2, st.L, signal);
    // ---- the finished map rows n = b0-2 .. b0+9 -> global (2 rows of 15 words per store) ----------------------------
    // out tile: 64 bytes per row = 32 strong bytes (one per 8-pixel group, byte g = output columns 8g .. 8g+7) + 32 weak
    {
      const int wi = lane & 15, gw = blockIdx.x * (MT_X / 16) + wi;
      const long long o0 = (long long)g.frame * p.pl_frame_stride16 + (long long)g.Y0 * p.pl_pitch16 + gw;
      uint16_t *s16 = reinterpret_cast<uint16_t *>(smem + MS_OUT);
      const int gmax = (p.w + 15) >> 4;   // 16-pixel groups per row
#pragma unroll
      for (int it = 0; it < MK / 2; ++it) {
        const int idx = 2 * it + (lane >> 4), n = b0 - 2 + idx;
        const uint32_t sv = s16[idx * 32 + wi], wv = s16[idx * 32 + 16 + wi];
        s16[idx * 32 + wi] = 0;
        s16[idx * 32 + 16 + wi] = 0;
        if (wi < MT_X / 16 && gw < gmax && n >= 0 && n < g.rows_out) {   // the two planes ARE the 2-bit map: S = strong, C = weak | strong
          p.pl_S[o0 + (long long)n * p.pl_pitch16] = (uint16_t)sv;
          p.pl_C[o0 + (long long)n * p.pl_pitch16] = (uint16_t)(sv | wv);
        }
      }
      __syncwarp();
    }
  }
}

// Row bands over peer memory: spin until the neighbour's halo rows have landed, 2 s time-out.  Every lane polls the same
// word and the exit condition goes through a shuffle: a loop the compiler can prove warp-uniform (a divergent one in
// front of the marching code would make it compile every later shuffle and vote with a divergence fallback).
__device__ __forceinline__ void m_halo_wait(const uint32_t *cnt, const int need, int *err)
{
#ifndef B2C_EMU
  unsigned long long t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    unsigned long long t;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    const int st = (int)v >= need ? 1 : (t - t0 > 2000000000ull ? 2 : 0);
    const int s0 = __shfl_sync(B2C_FULL, st, 0);
    if (s0) {
      if (s0 == 2) *err = 1;
      break;
    }
  }
  asm volatile("fence.acq_rel.sys;" ::: "memory");
#endif
}

// HALO: row bands over peer memory (the plain instance carries none of it)
template <int CH, bool HALO = false>
__global__ void __launch_bounds__(MARCH_THREADS, MARCH_CTAS_PER_SM) k_stencil_march(const B2cStencilParams p, const int rb)
{
  B2C_DYN_SMEM(smem);
  MarchGeo g;
  g.lane = threadIdx.x & 31;
  // (through a shuffle so that the compiler knows the role branch is warp-uniform: otherwise every shuffle and vote
  // below is compiled with a divergence fallback)
  const int warp = MARCH_DUO ? __shfl_sync(B2C_FULL, (int)(threadIdx.x >> 5), 0) : 0;
  g.X0 = blockIdx.x * MT_X;
  g.Y0 = blockIdx.y * rb;
  g.frame = blockIdx.z;
  g.rows_out = min(rb, p.h - g.Y0);           // this band produces rows Y0 .. Y0+rows_out-1
  g.xl = g.X0 - 8 + 8 * g.lane;               // first pixel column of this lane
  g.lane_in = g.xl >= 0 && g.xl < p.w;
  g.out_lane = g.lane_in && g.lane != 0 && g.lane != 31;   // lanes 0 and 31 are halo lanes
  g.yg0 = g.Y0 + p.y0;                        // global row of band-local row 0
  g.ilim = p.h + 4 - g.Y0;                    // band-local gray rows >= ilim are not backed by memory
  g.nblocks = (g.rows_out + 4 + MK - 1) / MK;
  // a lane that is only partly inside the image (w % 8 != 0): masks of its pixel pairs; `partial` is strip-uniform
  g.partial = (p.w & 7) != 0 && g.X0 - 8 + 256 > p.w;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = g.xl + 2 * j;
    g.pm[j] = (x < p.w ? 0xFFFFu : 0u) | (x + 1 < p.w ? 0xFFFF0000u : 0u);
  }
  // A lane outside the image never stores to the blur / gx / gy rings, so its slots keep the zeros written here:
  // that IS the reference's per-stage zero padding left and right of the image (cannyEdgeD.cu:142-149, 222-229).
  for (int i = threadIdx.x; i < MARCH_SMEM / 16; i += MARCH_THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  // halo rows of a row band that arrive while the launch runs: only the CTAs of the first / last band of rows read them,
  // and wait here (the neighbour's stores were issued before its own stencil: a few microseconds at most, while all other
  // CTAs of the launch are already at work)
  if (HALO) {
    if (p.halo_cnt_up != nullptr && blockIdx.y == 0) m_halo_wait(p.halo_cnt_up, p.halo_need, p.halo_err);
    if (p.halo_cnt_dn != nullptr && blockIdx.y == gridDim.y - 1) m_halo_wait(p.halo_cnt_dn, p.halo_need, p.halo_err);
  }
  if (MARCH_DUO) {
#if B2C_X & 12
    // profiling builds: one of the two warps only keeps the hand-over protocol alive
    if ((warp == 0 && (B2C_X & 8)) || (warp == 1 && (B2C_X & 4))) {
      for (int t = 0; t < g.nblocks; ++t) {
        if (warp == 0) {
          for (int j = 0; j < 2; ++j)
            if (t > 0) m_bar_sync(MB_EMPTY0 + j);
          m_bar_arrive(MB_FULL);
        } else {
          m_bar_sync(MB_FULL);
          if (t + 1 < g.nblocks)
            for (int j = 0; j < 2; ++j) m_bar_arrive(MB_EMPTY0 + j);
        }
      }
      return;
    }
#endif
    if (warp == 0) {
      MarchAState<CH> sa;
      m_a_init<CH>(p, g, smem, sa);
      for (int t = 0; t < g.nblocks; ++t) m_a_block<CH, true>(p, g, smem, sa, t);
    } else {
      MarchCState sc;
      m_c_init(p, sc);
      for (int t = 0; t < g.nblocks; ++t) m_c_block<true>(p, g, smem, sc, t);
    }
  } else {
    // one warp, both roles, block by block: 12 blur rows (+ replay), then their Sobel / NMS / map rows
    MarchAState<CH> sa;
    MarchCState sc;
    m_a_init<CH>(p, g, smem, sa);
    m_c_init(p, sc);
    for (int t = 0; t < g.nblocks; ++t) {
      m_a_block<CH, false>(p, g, smem, sa, t);
      if (!(B2C_X & 4)) m_c_block<false>(p, g, smem, sc, t);
    }
  }
}

#ifdef B2C_EMU
inline bool march_supported(const B2cStencilParams &p)
{
  return p.plane_stride == 0 && p.w >= 8 && p.row_stride % 8 == 0 && p.frame_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.bgr) & 7) == 0 &&
         p.row_stride >= (long long)((p.w + 7) / 8 * 8) * p.channels;
}
inline int march_emu_launch(const B2cStencilParams &p, int rb)
{
  if (!march_supported(p)) return -2;
  dim3 grid((p.w + MT_X - 1) / MT_X, (p.h + rb - 1) / rb, p.nframes);
  if (p.channels == 1) emu::launch(grid, dim3(MARCH_THREADS), MARCH_SMEM, false, [p, rb] { k_stencil_march<1>(p, rb); });
  else if (p.channels == 4) emu::launch(grid, dim3(MARCH_THREADS), MARCH_SMEM, false, [p, rb] { k_stencil_march<4>(p, rb); });
  else emu::launch(grid, dim3(MARCH_THREADS), MARCH_SMEM, false, [p, rb] { k_stencil_march<3>(p, rb); });
  return 0;
}
#else
inline cudaError_t march_configure(int *ctas_per_sm)
{
  cudaError_t e = cudaFuncSetAttribute(k_stencil_march<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_stencil_march<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MARCH_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_stencil_march<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, MARCH_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_stencil_march<3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_stencil_march<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_stencil_march<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_stencil_march<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_stencil_march<3, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e == cudaSuccess && ctas_per_sm) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_stencil_march<3>, MARCH_THREADS, MARCH_SMEM);
  return e;
}
// 8-byte aligned rows (64-bit loads: a lane's 8 pixels are 24 bytes); any width >= 8 (a last lane that is only partly
// inside the image is masked, but its 8 pixels must be backed by the row: row_stride >= ceil8(w) * channels).
// Anything else goes through the tile kernel.
inline bool march_supported(const B2cStencilParams &p)
{
  return p.plane_stride == 0 && p.w >= 8 && p.row_stride % 8 == 0 && p.frame_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.bgr) & 7) == 0 &&
         p.row_stride >= (long long)((p.w + 7) / 8 * 8) * p.channels;
}
// Rows per band.  Bands of 12k-4 rows waste no block (a band of rb rows runs ceil((rb+4)/12) blocks of 12 rows).
// Cost model: CTAs are dealt to `slots` resident CTAs; the launch takes ceil(waves) x (rows marched per CTA + a fixed
// prologue), so tall bands (little recompute) are good until the last wave is mostly empty.
inline int march_band_rows(int w, int h, int nframes, int sm_count, int ctas_per_sm)
{
  const double slots = (double)sm_count * (ctas_per_sm > 0 ? ctas_per_sm : MARCH_CTAS_PER_SM);
  const long long strips = (w + MT_X - 1) / MT_X;
  int best_rb = MK - 4;
  double best = 1e30;
  for (int k = 1; k <= (h + 4 + MK - 1) / MK; ++k) {
    const int rb = MK * k - 4;
    const long long nb = (h + rb - 1) / rb;
    const double waves = (double)(strips * nb * nframes) / slots;
    // a partly filled last wave still runs at full speed per CTA when the SM has fewer CTAs than slots only if the
    // kernel is latency-bound; it is issue-bound, so count a partial wave by its fill, but never less than one pass
    const double full = floor(waves), frac = waves - full;
    const double eff_waves = full + (frac > 0 ? (full >= 1 ? std::max(frac, 0.35) : 1.0) : 0.0);
    const double cost = eff_waves * (MK * k + 6);
    if (cost < best) { best = cost; best_rb = rb; }
  }
  return best_rb;
}
inline cudaError_t march_launch(const B2cStencilParams &p, int sm_count, int ctas_per_sm, int rb_override, cudaStream_t st, int extra_smem = 0)
{
  const int smem_bytes = MARCH_SMEM + extra_smem;   // (extra_smem: profiling knob that lowers the number of resident CTAs)
  const int rb = rb_override > 0 ? rb_override : march_band_rows(p.w, p.h, p.nframes, sm_count, ctas_per_sm);
  dim3 grid((p.w + MT_X - 1) / MT_X, (p.h + rb - 1) / rb, p.nframes);
  if (p.halo_cnt_up || p.halo_cnt_dn) k_stencil_march<3, true><<<grid, MARCH_THREADS, smem_bytes, st>>>(p, rb);   // (row bands are BGR8)
  else if (p.channels == 1) k_stencil_march<1><<<grid, MARCH_THREADS, smem_bytes, st>>>(p, rb);
  else if (p.channels == 4) k_stencil_march<4><<<grid, MARCH_THREADS, smem_bytes, st>>>(p, rb);
  else k_stencil_march<3><<<grid, MARCH_THREADS, smem_bytes, st>>>(p, rb);
  return cudaGetLastError();
}
#endif
}// namespace b2c
