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

__device__ __forceinline__ int uf_find(int *P, int n)
{
  while (n != 0) {
    const int pn = __ldcg(P + n - 1);
    if (pn == n) break;
    n = pn;
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

__global__ void __launch_bounds__(UF_THREADS) k_hysteresis_uf(const B2cHystParams p)
{
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (long long)gridDim.x * blockDim.x;
  const int wpr = (p.w + 31) >> 5;
  const int W32 = p.plane_pitch * 32;   // node id = y * W32 + x + 1 (0 = "touches strong")
  const long long total = (long long)p.nframes * p.h * wpr;

  // ---- A: bit planes ----
  if (!p.skip_init) {
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
  if (gtid == 0) { p.flags[3] = 1; p.flags[4] = 0; }
  __threadfence();
  B2C_GRID_SYNC();

  // ---- B: parents of the weak pixels ----
  for (long long i = gtid; i < total; i += gthreads) {
    const int xw = (int)(i % wpr);
    const long long t = i / wpr;
    const int y = (int)(t % p.h), f = (int)(t / p.h);
    const uint32_t *Sr = p.S + f * p.plane_frame_stride + (long long)y * p.plane_pitch + xw;
    const uint32_t sM = __ldcg(Sr);
    const uint32_t wd = __ldcg(p.C + f * p.plane_frame_stride + (long long)y * p.plane_pitch + xw) & ~sM;
    if (wd == 0u) continue;
    const int pp = p.plane_pitch;
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
      const int start = __ffs((int)lo) - 1;
      const int val = (run & near) ? 0 : base + start + 1;
      uint32_t r = run;
      while (r) {
        const int b = __ffs((int)r) - 1;
        r &= r - 1u;
        P[base + b] = val;
      }
    }
  }
  __threadfence();
  B2C_GRID_SYNC();

  // ---- C: unions with the W / NW / N / NE weak neighbours ----
  for (long long i = gtid; i < total; i += gthreads) {
    const int xw = (int)(i % wpr);
    const long long t = i / wpr;
    const int y = (int)(t % p.h), f = (int)(t / p.h);
    const long long o = f * p.plane_frame_stride + (long long)y * p.plane_pitch + xw;
    const uint32_t wd = p.C[o] & ~p.S[o];
    if (wd == 0u) continue;
    const int pp = p.plane_pitch;
    const uint32_t wl = xw > 0 ? (p.C[o - 1] & ~p.S[o - 1]) : 0u;
    uint32_t wu = 0u, wul = 0u, wur = 0u;
    if (y > 0) {
      wu = p.C[o - pp] & ~p.S[o - pp];
      if (xw > 0) wul = p.C[o - pp - 1] & ~p.S[o - pp - 1];
      if (xw + 1 < pp) wur = p.C[o - pp + 1] & ~p.S[o - pp + 1];
    }
    const uint32_t left = wd & 1u & (wl >> 31);                  // only bit 0 can have a left neighbour in another word
    const uint32_t up = wd & wu, upl = wd & ((wu << 1) | (wul >> 31)), upr = wd & ((wu >> 1) | (wur << 31));
    uint32_t any = left | up | upl | upr;
    int *P = p.parent + f * p.parent_frame_stride;
    const int base = y * W32 + xw * 32 + 1;
    while (any) {
      const int b = __ffs((int)any) - 1;
      any &= any - 1u;
      const int n = base + b;
      if ((left >> b) & 1u) uf_union(P, n, n - 1);
      if ((up >> b) & 1u) uf_union(P, n, n - W32);
      if ((upl >> b) & 1u) uf_union(P, n, n - W32 - 1);
      if ((upr >> b) & 1u) uf_union(P, n, n - W32 + 1);
    }
  }
  __threadfence();
  B2C_GRID_SYNC();

  // ---- D: resolve ----
  bool changed = false;
  for (long long i = gtid; i < total; i += gthreads) {
    const int xw = (int)(i % wpr);
    const long long t = i / wpr;
    const int y = (int)(t % p.h), f = (int)(t / p.h);
    const long long o = f * p.plane_frame_stride + (long long)y * p.plane_pitch + xw;
    const uint32_t s = p.S[o];
    uint32_t m = p.C[o] & ~s;
    if (m == 0u) continue;
    int *P = p.parent + f * p.parent_frame_stride;
    const int base = y * W32 + xw * 32 + 1;
    uint32_t add = 0u;
    while (m) {
      const int b = __ffs((int)m) - 1;
      m &= m - 1u;
      if (uf_find(P, base + b) == 0) add |= 1u << b;
    }
    if (add) {
      p.S[o] = s | add;
      changed = true;
    }
  }
  if (changed) atomicExch(p.flags + 4, 1);
  __threadfence();
  B2C_GRID_SYNC();

  // ---- E: S plane -> u8 {0,255} ----
  if (p.edges && !p.skip_expand) {
    const int gpr = (p.w + 15) >> 4;
    const long long tot = (long long)p.nframes * p.h * gpr;
    for (long long i = gtid; i < tot; i += gthreads) {
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
