// k_hysteresis_uf.cuh -- on-device hysteresis as ONE cooperative launch with a constant number of phases.
//
// Replaces the reference's CPU-driven relaunch loop (src/cvp/cannyEdgeH.cu:297-338: 1 + up to 100 launches of
// `hysteresis`, two blocking 4-byte memcpys per launch) and `removeCandidates` (src/cvp/cannyEdgeD.cu:379-395).
// The reference iterates "a 128 becomes 255 if any 8-neighbour is 255" to a fixpoint (cannyEdgeD.cu:333-363); the
// fixpoint is "weak pixels that are 8-connected to a strong pixel through weak pixels survive".  That is a
// connected-components question, answered here without rounds by a lock-free union-find over the WEAK pixels only
// (~0.7 % of a frame), with one virtual node 0 = "touches a strong pixel":
//   A. bit planes S (strong) and C (weak|strong) from the 2-bit map, 32 pixels per word;
//   B. init: every weak pixel gets parent = first pixel of its horizontal run inside the word (runs are born
//      flat), or 0 if any pixel of the run has a strong 8-neighbour (3x3 dilation of S done on words);
//   C. union: every weak pixel is united with its weak W / NW / N / NE neighbours (atomicMin links the larger
//      root under the smaller one, so 0 always wins);
//   D. resolve: a weak pixel is an edge iff find() == 0; S |= those bits;
//   E. expand S to the u8 {0,255} map the reference hands to its PBO.
// Phases are separated by grid-wide barriers inside the launch; nothing returns to the host.
// Parent words that other CTAs may update are read with ld.global.cg (L2).
#pragma once
#include "b2c_device.cuh"

namespace b2c
{
constexpr int UF_THREADS = 256;

// find with path halving.  Parents only ever decrease (every write is an atomicMin with an ancestor), so a
// concurrent union can never be lost, and since all pixels of a chain halve their paths concurrently the
// depth collapses like pointer jumping (a vertical weak line is born as a chain: each run hangs under the
// run above it).
__device__ __forceinline__ int uf_find(int *P, int n)
{
  while (n != 0) {
    const int pn = __ldcg(P + n - 1);
    if (pn == n || pn == 0) return pn;
    const int gp = __ldcg(P + pn - 1);
    if (gp == pn) return pn;
    atomicMin(P + n - 1, gp);
    n = gp;
  }
  return n;
}

// find for the resolve phase: no union runs concurrently any more, so every value ever stored in P[n-1] is an
// ancestor of n and plain stores are enough for the compression (no atomic round trip on the critical path).
__device__ __forceinline__ int uf_find_final(int *P, int n)
{
  while (n != 0) {
    const int pn = __ldcg(P + n - 1);
    if (pn == n || pn == 0) return pn;
    const int gp = __ldcg(P + pn - 1);
    if (gp == pn) return pn;
    __stcg(P + n - 1, gp);
    n = gp;
  }
  return n;
}

__device__ __forceinline__ void uf_union(int *P, int a, int b)
{
  for (;;) {
    a = uf_find(P, a);
    b = uf_find(P, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }   // a > b >= 0: hang a under b
    const int old = atomicMin(P + a - 1, b);
    if (old == a) return;
    a = old;   // somebody re-parented a meanwhile: continue from there
  }
}

// ---- per-word bodies of the three sparse phases -------------------------------------------------------------------
// B: parents of the weak pixels of one word
__device__ __forceinline__ void uf_init_word(const B2cHystParams &p, int f, int y, int xw, uint32_t wd, uint32_t sM, int W32)
{
  const int pp = p.plane_pitch;
  const uint32_t *Sr = p.S + f * p.plane_frame_stride + (long long)y * pp + xw;
  const uint32_t v = __ldcg(Sr - pp) | sM | __ldcg(Sr + pp);
  const uint32_t vl = xw > 0 ? (__ldcg(Sr - pp - 1) | __ldcg(Sr - 1) | __ldcg(Sr + pp - 1)) : 0u;
  const uint32_t vr = xw + 1 < pp ? (__ldcg(Sr - pp + 1) | __ldcg(Sr + 1) | __ldcg(Sr + pp + 1)) : 0u;
  const uint32_t near = wd & (v | (v << 1) | (v >> 1) | (vl >> 31) | (vr << 31));
  int *P = p.parent + f * p.parent_frame_stride;
  const int base = y * W32 + xw * 32;
  uint32_t m = wd;
  while (m) {
    const uint32_t lo = m & (0u - m);
    const uint32_t run = m & ~(m + lo);   // the run of ones that starts at the lowest set bit
    m &= ~run;
    const int val = (run & near) ? 0 : base + __ffs((int)lo);
    uint32_t r = run;
    while (r) {
      const int b = __ffs((int)r) - 1;
      r &= r - 1u;
      P[base + b] = val;
    }
  }
}

// C: one union per (run, touching fragment of the row above / previous word): a run already shares one parent
__device__ __forceinline__ void uf_union_word(const B2cHystParams &p, int f, int y, int xw, int W32)
{
  const int pp = p.plane_pitch;
  const long long o = f * p.plane_frame_stride + (long long)y * pp + xw;
  const uint32_t wd = p.C[o] & ~p.S[o];
  if (wd == 0u) return;
  const uint32_t wl = xw > 0 ? (p.C[o - 1] & ~p.S[o - 1]) : 0u;
  uint32_t wu = 0u, wul = 0u, wur = 0u;
  if (y > 0) {
    wu = p.C[o - pp] & ~p.S[o - pp];
    if (xw > 0) wul = p.C[o - pp - 1] & ~p.S[o - pp - 1];
    if (xw + 1 < pp) wur = p.C[o - pp + 1] & ~p.S[o - pp + 1];
  }
  int *P = p.parent + f * p.parent_frame_stride;
  const int base = y * W32 + xw * 32 + 1;
  uint32_t m = wd;
  while (m) {
    const uint32_t lo = m & (0u - m);
    const uint32_t run = m & ~(m + lo);
    m &= ~run;
    const int n = base + __ffs((int)lo) - 1;
    if ((run & 1u) && (wl >> 31)) uf_union(P, n, n - 1);                 // W neighbour in the previous word
    uint32_t a = (run | (run << 1) | (run >> 1)) & wu;                   // N / NW / NE inside this word column
    while (a) {
      const uint32_t lo2 = a & (0u - a);
      a &= (a + lo2);                                                    // drop the fragment that starts at lo2
      uf_union(P, n, base - W32 + __ffs((int)lo2) - 1);
    }
    if ((run & 1u) && (wul >> 31)) uf_union(P, n, base - W32 - 1);       // NW across the word boundary
    if ((run >> 31) && (wur & 1u)) uf_union(P, n, base - W32 + 32);      // NE across the word boundary
  }
}

// D: a run survives iff its root is node 0; returns true if the word gained edge bits
__device__ __forceinline__ bool uf_resolve_word(const B2cHystParams &p, int f, int y, int xw, int W32)
{
  const long long o = f * p.plane_frame_stride + (long long)y * p.plane_pitch + xw;
  const uint32_t s = p.S[o];
  uint32_t m = p.C[o] & ~s;
  if (m == 0u) return false;
  int *P = p.parent + f * p.parent_frame_stride;
  const int base = y * W32 + xw * 32 + 1;
  uint32_t add = 0u;
  while (m) {
    const uint32_t lo = m & (0u - m);
    const uint32_t run = m & ~(m + lo);
    m &= ~run;
    if (uf_find_final(P, base + __ffs((int)lo) - 1) == 0) add |= run;
  }
  if (add) p.S[o] = s | add;
  return add != 0u;
}

#ifndef B2C_EMU
__device__ __forceinline__ unsigned b2c_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return (unsigned)t; }
#define B2C_STAMP(k) do { if (gtid == 0) p.flags[8 + (k)] = (int)b2c_gtime(); } while (0)
#else
#define B2C_STAMP(k) do { } while (0)
#endif

constexpr int UF_LIST_CAP = 512;   // words with weak pixels remembered per warp (2 KB of shared memory)
constexpr int UF_SMEM = (UF_THREADS / 32) * UF_LIST_CAP * 4;

__global__ void __launch_bounds__(UF_THREADS) k_hysteresis_uf(const B2cHystParams p)
{
  B2C_DYN_SMEM(smem);
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (long long)gridDim.x * blockDim.x;
  const int wpr = (p.w + 31) >> 5;
  const int W32 = p.plane_pitch * 32;   // node id = y * W32 + x + 1 (0 = "touches strong")
  // one warp per plane row: (frame, y) from a 32-bit division per row, lanes stride over the row's words
  const int lane = threadIdx.x & 31;
  const int gwarp = (int)(gtid >> 5), nwarps = (int)(gthreads >> 5);
  const int nrows = p.nframes * p.h;
  // Words that hold weak pixels are ~10 % of a plane: phase B remembers them per warp (shared memory survives
  // the grid barriers), so that the latency-bound phases C and D run with every lane busy.
  uint32_t *wlist = reinterpret_cast<uint32_t *>(smem) + (threadIdx.x >> 5) * UF_LIST_CAP;
  int wcnt = 0;
  bool wovf = wpr > 1024;
#define B2C_FOR_WORDS                                            \
  for (int row_ = gwarp; row_ < nrows; row_ += nwarps)           \
    for (int xw = lane, f = row_ / p.h, y = row_ - f * p.h; xw < wpr; xw += 32)

  B2C_STAMP(0);
  // ---- A: bit planes ----
  if (!p.skip_init) {
    B2C_FOR_WORDS {
      const uint32_t *mrow = p.map2 + f * p.map_frame_stride + (long long)y * p.map_pitch;
      const uint32_t m0 = mrow[2 * xw], m1 = (2 * xw + 1 < p.map_pitch) ? mrow[2 * xw + 1] : 0u;
      const uint32_t s = (m0 & 0xFFFFu) | (m1 << 16), wk = (m0 >> 16) | (m1 & 0xFFFF0000u);
      const long long o = f * p.plane_frame_stride + (long long)y * p.plane_pitch + xw;
      p.S[o] = s;
      p.C[o] = s | wk;
    }
  }
  if (gtid == 0) { p.flags[3] = 1; p.flags[4] = 0; }
  __threadfence();
  B2C_GRID_SYNC();

  B2C_STAMP(1);
  // ---- B: parents of the weak pixels; remember the words that have any ----
  for (int row_ = gwarp; row_ < nrows; row_ += nwarps) {
    const int f = row_ / p.h, y = row_ - f * p.h;
    for (int xw0 = 0; xw0 < wpr; xw0 += 32) {
      const int xw = xw0 + lane;
      uint32_t wd = 0u, sM = 0u;
      if (xw < wpr) {
        const long long o = f * p.plane_frame_stride + (long long)y * p.plane_pitch + xw;
        sM = __ldcg(p.S + o);
        wd = __ldcg(p.C + o) & ~sM;
      }
      if (wd) uf_init_word(p, f, y, xw, wd, sM, W32);
      const uint32_t mask = __ballot_sync(B2C_FULL, wd != 0u);
      if (!wovf) {
        const int add = __popc(mask);
        if (wcnt + add > UF_LIST_CAP) wovf = true;
        else {
          if (wd) wlist[wcnt + __popc(mask & ((1u << lane) - 1u))] = ((uint32_t)row_ << 10) | (uint32_t)xw;
          wcnt += add;
        }
      }
    }
  }
  __threadfence();
  B2C_GRID_SYNC();

  B2C_STAMP(2);
  // ---- C: unions ----
  if (!wovf) {
    __syncwarp();
    for (int i = lane; i < wcnt; i += 32) {
      const uint32_t e = wlist[i];
      const int row_ = (int)(e >> 10), f = row_ / p.h;
      uf_union_word(p, f, row_ - f * p.h, (int)(e & 1023u), W32);
    }
  } else {
    B2C_FOR_WORDS uf_union_word(p, f, y, xw, W32);
  }
  __threadfence();
  B2C_GRID_SYNC();

  B2C_STAMP(3);
  // ---- D: resolve ----
  bool changed = false;
  if (!wovf) {
    for (int i = lane; i < wcnt; i += 32) {
      const uint32_t e = wlist[i];
      const int row_ = (int)(e >> 10), f = row_ / p.h;
      changed |= uf_resolve_word(p, f, row_ - f * p.h, (int)(e & 1023u), W32);
    }
  } else {
    B2C_FOR_WORDS changed |= uf_resolve_word(p, f, y, xw, W32);
  }
  if (changed) atomicExch(p.flags + 4, 1);
  __threadfence();
  B2C_GRID_SYNC();

  B2C_STAMP(4);
  // ---- E: S plane -> u8 {0,255} ----
  if (p.edges && !p.skip_expand) {
    const int gpr = (p.w + 15) >> 4;
    for (int row_ = gwarp; row_ < nrows; row_ += nwarps)
      for (int g = lane, f = row_ / p.h, y = row_ - f * p.h; g < gpr; g += 32) {
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
  B2C_STAMP(5);
#undef B2C_FOR_WORDS
}
}// namespace b2c
