// k_hysteresis.cuh -- on-device hysteresis: one cooperative launch, no host round trip.
//
// Replaces the reference's CPU-driven relaunch loop (src/cvp/cannyEdgeH.cu:297-338: up to 1+100 launches
// of `hysteresis`, each bracketed by two blocking 4-byte memcpys) and `removeCandidates`
// (src/cvp/cannyEdgeD.cu:379-395).  The reference promotes 128 -> 255 when any 8-neighbour is 255 until
// nothing changes (cannyEdgeD.cu:333-363); its fixpoint is "weak pixels 8-connected to a strong pixel
// through weak/strong pixels".  Any correct propagation reaches the same fixpoint, so we work on bit
// planes (32 pixels per word):
//     S |= C & dilate3x3(S)          S = strong-so-far, C = weak|strong
// Phases, separated by grid-wide barriers inside the one launch:
//   0. build S and C planes from the 2-bit map;
//   1. rounds: every warp owns tiles of 1024 px x tile_rows; it loads the tile plus a 1-px ring of S
//      from the global plane, runs scanline sweeps (down, up) in shared memory until the tile is
//      stable -- horizontal runs are closed in O(1) word ops, lanes hand run ends to their neighbours
//      by shuffle -- and writes S back if it changed.  A device-side flag says whether any tile
//      changed; the loop ends in the first round where none did (global fixpoint).  S words that other
//      CTAs may rewrite are read with ld.global.cg (L2), never through the non-coherent L1.
//   2. expand the S plane to the u8 {0,255} edge map.
#pragma once
#include "b2c_device.cuh"

namespace b2c
{
constexpr int HYST_THREADS = 256;
constexpr int HYST_SW = 34;   // tile row in smem: 32 words + one halo word each side

__host__ __device__ inline int hyst_smem_bytes(int tile_rows) { return (HYST_THREADS / 32) * ((tile_rows + 2) * HYST_SW + tile_rows * 32) * 4; }

// Flood the seed bits s along the runs of ones of c, both directions (Kogge-Stone occluded fill).
__device__ __forceinline__ uint32_t hfill(uint32_t s, uint32_t c)
{
  uint32_t m = c;
  s |= m & (s << 1); m &= m << 1;
  s |= m & (s << 2); m &= m << 2;
  s |= m & (s << 4); m &= m << 4;
  s |= m & (s << 8); m &= m << 8;
  s |= m & (s << 16);
  m = c;
  s |= m & (s >> 1); m &= m >> 1;
  s |= m & (s >> 2); m &= m >> 2;
  s |= m & (s >> 4); m &= m >> 4;
  s |= m & (s >> 8); m &= m >> 8;
  s |= m & (s >> 16);
  return s;
}

// One scanline step on smem row r (1..rows).  Returns (per lane) whether its word changed.
__device__ __forceinline__ bool hyst_row_step(uint32_t *sS, const uint32_t *sC, int r, int lane)
{
  const uint32_t *a = sS + (r - 1) * HYST_SW + lane, *m = a + HYST_SW, *b = m + HYST_SW;
  const uint32_t cur = m[1], curL = m[0], curR = m[2];
  const uint32_t nb = a[1] | b[1] | cur, nbL = a[0] | b[0] | curL, nbR = a[2] | b[2] | curR;
  const uint32_t c = sC[(r - 1) * 32 + lane];
  const uint32_t dil = nb | (nb << 1) | (nb >> 1) | (nbL >> 31) | (nbR << 31);
  uint32_t s = hfill(cur | (c & dil), c);
  for (;;) {   // hand run ends across lanes until the 1024-px row segment is closed
    uint32_t sl = __shfl_up_sync(B2C_FULL, s, 1), sr = __shfl_down_sync(B2C_FULL, s, 1);
    if (lane == 0) sl = curL;
    if (lane == 31) sr = curR;
    const uint32_t add = c & ~s & ((sl >> 31) | (sr << 31));
    if (!__any_sync(B2C_FULL, add != 0)) break;
    s = hfill(s | add, c);
  }
  const bool changed = s != cur;
  __syncwarp();   // everyone has read row r before anyone overwrites it
  if (changed) sS[r * HYST_SW + lane + 1] = s;
  __syncwarp();
  return changed;
}

__global__ void __launch_bounds__(HYST_THREADS) k_hysteresis(const B2cHystParams p)
{
  B2C_DYN_SMEM(smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (long long)gridDim.x * blockDim.x;
  const int gwarp = warp * gridDim.x + blockIdx.x, total_warps = gridDim.x * nwarps;   // CTA-major: spread tiles over SMs first
  const int wpr = (p.w + 31) >> 5;
  const int TR = p.tile_rows;

  // ---- phase 0: bit planes from the 2-bit map ----
  if (!p.skip_init) {
    const long long total = (long long)p.nframes * p.h * wpr;
    for (long long i = gtid; i < total; i += gthreads) {
      const int xw = (int)(i % wpr);
      const long long t = i / wpr;
      const int y = (int)(t % p.h), f = (int)(t / p.h);
      const uint32_t *mrow = p.map2 + f * p.map_frame_stride + (long long)y * p.map_pitch;
      const uint32_t m0 = mrow[2 * xw], m1 = (2 * xw + 1 < p.map_pitch) ? mrow[2 * xw + 1] : 0u;
      const uint32_t s = (m0 & 0xFFFFu) | (m1 << 16), wk = (m0 >> 16) | (m1 & 0xFFFF0000u);
      const long long o = f * p.plane_frame_stride + (long long)y * p.plane_pitch + xw;
      p.S[o] = s;
      p.C[o] = s | wk;
    }
  }
  if (gtid == 0) { p.flags[0] = 0; p.flags[4] = 0; }
  __threadfence();
  B2C_GRID_SYNC();

  // ---- phase 1: rounds to the global fixpoint ----
  uint32_t *sS = reinterpret_cast<uint32_t *>(smem) + warp * ((TR + 2) * HYST_SW + TR * 32);
  uint32_t *sC = sS + (TR + 2) * HYST_SW;
  const int tiles_x = (wpr + 31) >> 5, tiles_y = (p.h + TR - 1) / TR;
  const int ntiles = p.nframes * tiles_x * tiles_y;
  int round = 0;
  for (; round < p.max_rounds; ++round) {
    int *flag = p.flags + (round % 3);
    if (gtid == 0) p.flags[(round + 1) % 3] = 0;
    bool any_changed = false;
    for (int t = gwarp; t < ntiles; t += total_warps) {
      const int f = t / (tiles_x * tiles_y), rem = t - f * (tiles_x * tiles_y);
      const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
      const int y0 = ty * TR, rows = min(TR, p.h - y0), xw0 = tx * 32, xw = xw0 + lane;
      const bool in = xw < wpr;
      uint32_t *Sf = p.S + f * p.plane_frame_stride;
      const uint32_t *Cf = p.C + f * p.plane_frame_stride;
      __syncwarp();
      bool has_work = false;
      for (int r = 0; r < rows + 2; ++r) {
        const uint32_t *row = Sf + (long long)(y0 + r - 1) * p.plane_pitch;
        sS[r * HYST_SW + lane + 1] = in ? __ldcg(row + xw) : 0u;
        if (lane == 0) sS[r * HYST_SW] = xw0 > 0 ? __ldcg(row + xw0 - 1) : 0u;
        if (lane == 31) sS[r * HYST_SW + 33] = (xw0 + 32 < wpr) ? __ldcg(row + xw0 + 32) : 0u;
      }
      for (int r = 0; r < rows; ++r) {
        const uint32_t c = in ? Cf[(long long)(y0 + r) * p.plane_pitch + xw] : 0u;
        sC[r * 32 + lane] = c;
        has_work |= (c != 0);
      }
      __syncwarp();
      if (!__any_sync(B2C_FULL, has_work)) continue;   // no candidates at all in this tile
      bool tile_changed = false;
      for (;;) {
        bool ch = false;
        for (int r = 1; r <= rows; ++r) ch |= hyst_row_step(sS, sC, r, lane);
        for (int r = rows - 1; r >= 1; --r) ch |= hyst_row_step(sS, sC, r, lane);
        if (!__any_sync(B2C_FULL, ch)) break;
        tile_changed = true;
      }
      if (tile_changed) {
        for (int r = 1; r <= rows; ++r)
          if (in) __stcg(Sf + (long long)(y0 + r - 1) * p.plane_pitch + xw, sS[r * HYST_SW + lane + 1]);
        any_changed = true;
      }
    }
    if (any_changed && lane == 0) { atomicExch(flag, 1); atomicExch(p.flags + 4, 1); }
    __threadfence();
    B2C_GRID_SYNC();
    if (*reinterpret_cast<volatile int *>(flag) == 0) break;
  }
  if (gtid == 0) p.flags[3] = round;

  // ---- phase 2: S plane -> u8 {0,255} ----
  if (p.edges && !p.skip_expand) {
    const int gpr = (p.w + 15) >> 4;
    const long long total = (long long)p.nframes * p.h * gpr;
    for (long long i = gtid; i < total; i += gthreads) {
      const int g = (int)(i % gpr);
      const long long t = i / gpr;
      const int y = (int)(t % p.h), f = (int)(t / p.h);
      const uint32_t word = __ldcg(p.S + f * p.plane_frame_stride + (long long)y * p.plane_pitch + (g >> 1));
      const uint32_t bits = (word >> ((g & 1) * 16)) & 0xFFFFu;
      uint8_t *out = p.edges + f * p.edges_frame_stride + (long long)y * p.edges_pitch + g * 16;
      const int n = min(16, p.w - g * 16);
      if (n == 16 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
        uint4 v;
        v.x = (((bits & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
        v.y = ((((bits >> 4) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
        v.z = ((((bits >> 8) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
        v.w = ((((bits >> 12) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
        *reinterpret_cast<uint4 *>(out) = v;
      } else {
        for (int k = 0; k < n; ++k) out[k] = ((bits >> k) & 1u) ? 255 : 0;
      }
    }
  }
}
}// namespace b2c
