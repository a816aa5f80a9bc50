// k_hysteresis_uf.cuh -- on-device hysteresis: union-find over the weak pixels, three ordinary launches, no host round trip.
//
// Replaces the reference's CPU-driven relaunch loop (src/cvp/cannyEdgeH.cu:297-338: 1 + up to 100 launches of
// `hysteresis`, two blocking 4-byte memcpys per launch) and `removeCandidates` (src/cvp/cannyEdgeD.cu:379-395).
// The reference iterates "a 128 becomes 255 if any 8-neighbour is 255" to a fixpoint (cannyEdgeD.cu:333-363); the
// fixpoint is "weak pixels that are 8-connected to a strong pixel through weak pixels survive".  That is a
// connected-components question, answered here without rounds by a lock-free union-find over the WEAK pixels only
// (~0.7 % of a frame), with one virtual node 0 = "touches a strong pixel":
//   planes S (strong) and C (weak|strong), 32 pixels per word (written by the stencil kernel);
//   every horizontal run of weak pixels inside a word is a node; a run that touches a strong pixel (3x3 dilation of S
//   done on words) starts under node 0;
//   unions with the weak W / NW / N / NE neighbours (atomicMin links the larger root under the smaller one, so 0 wins);
//   resolve: a run is an edge iff find() == 0; E = S | those bits; E expands to the u8 {0,255} map of the reference's PBO.
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
// Row bands: k_seam_publish TAGS the roots of the components that reach the band's first or last row (sign bit of the
// root's own entry, UF_TAG); such a root is returned with its tag (a negative number), every other root as itself.
constexpr int UF_TAG = (int)0x80000000u;
__device__ __forceinline__ int uf_find_final(int *P, int n)
{
  while (n != 0) {
    const int pn = __ldcg(P + n - 1);
    if (pn == n || pn <= 0) return pn;
    const int gp = __ldcg(P + pn - 1);
    if (gp == pn) return pn;
    if (gp < 0) return gp;
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

// lowest bit of the run of ones of X that contains bit k (X has bit k set)
__device__ __forceinline__ int uf_run_head(uint32_t X, int k)
{
  const uint32_t inv = ~X & ((1u << k) - 1u);   // zeros below k
  return inv ? 32 - __clz((int)inv) : 0;
}

// rank of the run that starts at bit hb among the runs of word X (a 32-bit word holds at most 16 runs)
__device__ __forceinline__ int uf_run_rank(uint32_t X, int hb) { return __popc(X & ~(X << 1) & ((1u << hb) - 1u)); }

// ---- per-word bodies -----------------------------------------------------------------------------------------
// C: one union per (run, touching fragment of the row above / previous word): a run already shares one parent.
// doW / doN / doNW / doNE select which adjacencies are united here (the tile kernel has already united the ones that
// stay inside a tile).
__device__ __forceinline__ void uf_union_word(const B2cHystParams &p, int f, int y, int xw, int W32, bool doW = true, bool doN = true, bool doNW = true, bool doNE = true)
{
  const int pp = p.plane_pitch;
  const long long o = f * p.plane_frame_stride + (long long)y * pp + xw;
  const uint32_t wd = p.C[o] & ~p.S[o];
  if (wd == 0u) return;
  const uint32_t wl = (doW && xw > 0) ? (p.C[o - 1] & ~p.S[o - 1]) : 0u;
  uint32_t wu = 0u, wul = 0u, wur = 0u;
  if (y > 0) {
    if (doN) wu = p.C[o - pp] & ~p.S[o - pp];
    if (doNW && xw > 0) wul = p.C[o - pp - 1] & ~p.S[o - pp - 1];
    if (doNE && xw + 1 < pp) wur = p.C[o - pp + 1] & ~p.S[o - pp + 1];
  }
  int *P = p.parent + f * p.parent_frame_stride;
  const int base = y * W32 + xw * 32 + 1;
  uint32_t m = wd;
  while (m) {
    const uint32_t lo = m & (0u - m);
    const uint32_t run = m & ~(m + lo);
    m &= ~run;
    const int n = base + __ffs((int)lo) - 1;
    // every union partner is named by the HEAD of its run inside its word (the only pixels the tile kernel gives a parent)
    if ((run & 1u) && (wl >> 31)) uf_union(P, n, base - 32 + uf_run_head(wl, 31));      // W neighbour in the previous word
    uint32_t a = (run | (run << 1) | (run >> 1)) & wu;                                 // N / NW / NE inside this word column
    while (a) {
      const uint32_t lo2 = a & (0u - a);
      a &= (a + lo2);                                                                  // drop the fragment that starts at lo2
      uf_union(P, n, base - W32 + uf_run_head(wu, __ffs((int)lo2) - 1));
    }
    if ((run & 1u) && (wul >> 31)) uf_union(P, n, base - W32 - 32 + uf_run_head(wul, 31));   // NW across the word boundary
    if ((run >> 31) && (wur & 1u)) uf_union(P, n, base - W32 + 32);                         // NE across the word boundary (bit 0 is a head)
  }
}

// =====================================================================================================================
// THREE ordinary launches on one stream (no host round trip between them):
//   k_uf_tile (tile-local union-find in shared memory) -> k_uf_border -> k_uf_resolve (+ expansion to the u8 map).
// Kernel boundaries take the place of grid-wide barriers, every phase gets its own thread mapping (one thread per
// plane word; words without weak pixels leave at once) and the block scheduler balances the load.
// =====================================================================================================================
constexpr int UFK_THREADS = 256;
constexpr int UF_XW_BITS = 12;   // border-list entry = (row << 12) | plane word: images up to 131072 pixels wide, 2^20 rows

// ---- tile build: the union-find of a 32-row x 256-pixel tile entirely in shared memory ------------------------------
// Concurrent unions on a long weak line build a parent chain as long as the line (every run hooks under the run above
// it at the same time), and every later find() walks it with two dependent L2 round trips per step: measured 85 +
// 85 us for the union and resolve phases of 32 x 1080p however well the load was balanced.  So the chains are built
// and FLATTENED where a step costs ~30 cycles: one CTA per tile, one thread per plane word, parents in shared memory,
// local unions (W / N / NW / NE inside the tile), then every weak pixel's GLOBAL parent is written as the root of its
// tile-local tree (or 0 = "touches a strong pixel").  What remains for global memory are the unions across tile
// borders (k_uf_border) on trees whose depth is the number of tile crossings, not the number of pixels.
constexpr int UT_ROWS = 32, UT_WORDS = 8, UT_THREADS = UT_ROWS * UT_WORDS;
constexpr int UT_OFF_LW = (UT_ROWS * UT_WORDS * 16 + 4) * 4, UT_OFF_LS = UT_OFF_LW + UT_THREADS * 4, UT_OFF_IT = UT_OFF_LS + UT_THREADS * 4,
              UT_OFF_WC = UT_OFF_IT + UT_THREADS * 2, UT_SMEM = UT_OFF_WC + 64;

__device__ __forceinline__ int ut_find(int *P, int n)
{
  while (n != 0) {
    const int pn = P[n];
    if (pn == n || pn == 0) return pn;
    const int gp = P[pn];
    if (gp == pn) return pn;
    atomicMin(P + n, gp);
    n = gp;
  }
  return n;
}
__device__ __forceinline__ void ut_union(int *P, int a, int b)
{
  for (;;) {
    a = ut_find(P, a);
    b = ut_find(P, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(P + a, b);
    if (old == a) return;
    a = old;
  }
}

__global__ void __launch_bounds__(UT_THREADS) k_uf_tile(const B2cHystParams p, uint32_t *blist, int *bcount, const int bcap)
{
  B2C_DYN_SMEM(smem);
  int *LP = reinterpret_cast<int *>(smem);   // local parents: node 1 + 16 * word + (rank of the run inside its word); 0 = strong
  uint32_t *LW = reinterpret_cast<uint32_t *>(smem + UT_OFF_LW);             // weak words of the tile
  uint32_t *LS = reinterpret_cast<uint32_t *>(smem + UT_OFF_LS);             // strong words of the tile
  uint16_t *IT = reinterpret_cast<uint16_t *>(smem + UT_OFF_IT);             // compacted list: tile positions of the words with weak pixels
  int *WC = reinterpret_cast<int *>(smem + UT_OFF_WC);                       // per-warp counts
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wpr = (p.w + 31) >> 5;
  const int f = blockIdx.z, y0 = blockIdx.y * UT_ROWS, xw0 = blockIdx.x * UT_WORDS;
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0) p.flags[3] = 1;   // one on-device pass (b2c_last_timings [5])
  // ---- phase 0: S / C planes -> tile copies + compacted list of the words with weak pixels.  Two warps do it with
  // 128-bit loads (thread = 4 consecutive words of one tile row; rows are 16-byte aligned, words past the image are zero
  // padding), the other six go straight to the barrier: with one word per thread this phase alone was ~90 instructions
  // in each of the 8 warps of every tile, and the kernel is issue-bound on it.
  if (tid < 2 * UT_ROWS) {
    const int ly = tid >> 1, lw = 4 * (tid & 1), y = y0 + ly, xw = xw0 + lw;
    uint4 sv = make_uint4(0u, 0u, 0u, 0u), cv = sv;
    if (y < p.h && xw < wpr) {   // the planes were written by the stencil kernel: S = strong, C = weak | strong
      const long long o = f * p.plane_frame_stride + (long long)y * p.plane_pitch + xw;
      sv = *reinterpret_cast<const uint4 *>(p.S + o);
      cv = *reinterpret_cast<const uint4 *>(p.C + o);
    }
    const uint4 wv = make_uint4(cv.x & ~sv.x, cv.y & ~sv.y, cv.z & ~sv.z, cv.w & ~sv.w);
    const int t0 = ly * UT_WORDS + lw;   // tile position of the first of the 4 words
    *reinterpret_cast<uint4 *>(LW + t0) = wv;
    *reinterpret_cast<uint4 *>(LS + t0) = sv;
    // list position of (lane, k): words k of all lanes, then words k+1 ... (any order serves)
    const unsigned m0 = __ballot_sync(B2C_FULL, wv.x != 0u), m1 = __ballot_sync(B2C_FULL, wv.y != 0u), m2 = __ballot_sync(B2C_FULL, wv.z != 0u),
                   m3 = __ballot_sync(B2C_FULL, wv.w != 0u);
    const int n0 = __popc(m0), n1 = __popc(m1), n2 = __popc(m2), n3 = __popc(m3), mine = n0 + n1 + n2 + n3;
    if (lane == 0) WC[warp] = mine;
    // (both warps are past their ballots before either reads the other's count: named barrier over the two warps)
#ifdef B2C_EMU
    emu::named_bar_sync(1, 64);
#else
    asm volatile("bar.sync 1, 64;" ::: "memory");
#endif
    const int before = warp == 0 ? 0 : WC[0];
    const unsigned lt = (1u << lane) - 1u;
    if (wv.x) IT[before + __popc(m0 & lt)] = (uint16_t)t0;
    if (wv.y) IT[before + n0 + __popc(m1 & lt)] = (uint16_t)(t0 + 1);
    if (wv.z) IT[before + n0 + n1 + __popc(m2 & lt)] = (uint16_t)(t0 + 2);
    if (wv.w) IT[before + n0 + n1 + n2 + __popc(m3 & lt)] = (uint16_t)(t0 + 3);
    if (tid == 32) WC[UT_THREADS / 32] = before + mine;   // warp 1 knows the total
  }
  __syncthreads();
  const int nitems = WC[UT_THREADS / 32];
  // ---- the remaining phases run on the compacted list: thread k takes the k-th word that holds weak pixels, so the
  // divergent per-run loops fill whole warps instead of a few lanes of every warp
  // Items are dealt round-robin to W warps, not packed into warp 0: the per-run loops are dependent chains of
  // shared-memory atomics, so several warps with a few busy lanes each finish sooner than one full warp while the others
  // wait at the barrier -- but every warp that takes part issues the whole divergent code: W = 8 is fastest for one frame
  // (latency), W = 4 for big batches (issue slots).
  // p.spread = W in {1, 2, 4, 8}: consecutive items are dealt to W warps (W = 1: packed, item = tid)
  const int W = p.spread, item = (warp / W) * (32 * W) + lane * W + (warp % W);
  const bool act = item < nitems;
  const int t = act ? IT[item] : 0, ly = t >> 3, lw = t & 7, y = y0 + ly, xw = xw0 + lw;
  const uint32_t wd = act ? LW[t] : 0u;
  const int lbase = 1 + t * 16;   // local node of the first run of this word
  if (act) {
    // strong bits of the 3x3 neighbourhood, on words: from the tile copy, or from the S plane outside the tile (rows -1
    // and h are the zero ghost rows of the plane)
    const int pp = p.plane_pitch;
    const uint32_t *Sr = p.S + f * p.plane_frame_stride + (long long)y * pp + xw;
    auto srow = [&](int dy, int dx) -> uint32_t {   // S word (y + dy, xw + dx)
      const int l2 = ly + dy, w2 = lw + dx, x2 = xw + dx, yy = y + dy;
      if (x2 < 0 || x2 >= wpr) return 0u;
      if (l2 >= 0 && l2 < UT_ROWS && yy < p.h && w2 >= 0 && w2 < UT_WORDS) return LS[l2 * UT_WORDS + w2];
      return __ldg(Sr + (long long)dy * pp + dx);
    };
    uint32_t v = LS[t] | srow(-1, 0) | srow(1, 0), vl = 0u, vr = 0u;
    if (wd & 1u) vl = srow(-1, -1) | srow(0, -1) | srow(1, -1);
    if (wd >> 31) vr = srow(-1, 1) | srow(0, 1) | srow(1, 1);
    const uint32_t near = wd & (v | (v << 1) | (v >> 1) | (vl >> 31) | (vr << 31));
    uint32_t m = wd;
    while (m) {
      const uint32_t lo = m & (0u - m);
      const uint32_t run = m & ~(m + lo);   // the run of ones that starts at the lowest set bit
      m &= ~run;
      const int n = lbase + uf_run_rank(wd, __ffs((int)lo) - 1);
      LP[n] = (run & near) ? 0 : n;
    }
  }
  {   // words with weak pixels on the tile border have union partners in other tiles: per-frame list for k_uf_border
      // (one atomic per warp on the frame's own counter: no cross-frame contention)
    const bool bw = act && (ly == 0 || lw == 0 || lw == UT_WORDS - 1);
    const uint32_t bm = __ballot_sync(B2C_FULL, bw);
    if (bm) {
      const int leader = __ffs((int)bm) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(bcount + f, __popc(bm));
      base = __shfl_sync(B2C_FULL, base, leader);
      if (bw) blist[(long long)f * bcap + base + __popc(bm & ((1u << lane) - 1u))] = ((uint32_t)y << UF_XW_BITS) | (uint32_t)xw;
    }
  }
  __syncthreads();
  if (act) {   // local unions: W, and N / NW / NE with the row above, inside the tile
    const uint32_t wl = lw > 0 ? LW[t - 1] : 0u;
    uint32_t wu = 0u, wul = 0u, wur = 0u;
    if (ly > 0) {
      wu = LW[t - UT_WORDS];
      if (lw > 0) wul = LW[t - UT_WORDS - 1];
      if (lw + 1 < UT_WORDS) wur = LW[t - UT_WORDS + 1];
    }
    uint32_t m = wd;
    while (m) {
      const uint32_t lo = m & (0u - m);
      const uint32_t run = m & ~(m + lo);
      m &= ~run;
      const int n = lbase + uf_run_rank(wd, __ffs((int)lo) - 1);
      if ((run & 1u) && (wl >> 31)) ut_union(LP, n, lbase - 16 + uf_run_rank(wl, uf_run_head(wl, 31)));
      uint32_t a = (run | (run << 1) | (run >> 1)) & wu;
      while (a) {
        const uint32_t lo2 = a & (0u - a);
        a &= (a + lo2);
        ut_union(LP, n, lbase - UT_WORDS * 16 + uf_run_rank(wu, uf_run_head(wu, __ffs((int)lo2) - 1)));
      }
      if ((run & 1u) && (wul >> 31)) ut_union(LP, n, lbase - UT_WORDS * 16 - 16 + uf_run_rank(wul, uf_run_head(wul, 31)));
      if ((run >> 31) && (wur & 1u)) ut_union(LP, n, lbase - UT_WORDS * 16 + 16);   // bit 0 starts the first run
    }
  }
  __syncthreads();
  if (act) {   // global parents of the run heads = tile-local roots
    const int W32 = p.plane_pitch * 32;
    int *P = p.parent + f * p.parent_frame_stride;
    const int gbase = y * W32 + xw * 32;   // global node of bit b = gbase + b + 1, stored at P[gbase + b]
    uint32_t m = wd;
    while (m) {
      const uint32_t lo = m & (0u - m);
      const uint32_t run = m & ~(m + lo);
      m &= ~run;
      const int hb = __ffs((int)lo) - 1;
      int r = lbase + uf_run_rank(wd, hb);
      for (;;) {   // plain walk: the local trees are final now
        const int pr = LP[r];
        if (pr == r || pr == 0) { r = pr; break; }
        r = pr;
      }
      int val = 0;
      if (r != 0) {
        const int q = r - 1, tw = q >> 4;             // root = run number (q & 15) of tile word tw
        uint32_t st = LW[tw] & ~(LW[tw] << 1);        // its run starts
        for (int k = q & 15; k > 0; --k) st &= st - 1u;
        val = (y0 + (tw >> 3)) * W32 + (xw0 + (tw & 7)) * 32 + __ffs((int)st);   // global node = pixel index + 1
      }
      P[gbase + hb] = val;
    }
  }
}

// unions across tile borders (everything k_uf_tile could not see), from the per-frame lists written by k_uf_tile.
// Grid: x = blocks over a frame's list, z = frame.
__global__ void __launch_bounds__(UFK_THREADS) k_uf_border(const B2cHystParams p, const uint32_t *blist, const int *bcount, const int bcap)
{
  // Four threads per list entry, one per kind of neighbour (W, N, NW, NE): a union is a chain of dependent L2 round
  // trips (two finds + an atomicMin), and a word's unions done one after the other by ONE thread were the whole
  // duration of this kernel (22 us even for a single frame).
  const int f = blockIdx.z, n = 4 * bcount[f], W32 = p.plane_pitch * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t e = blist[(long long)f * bcap + (i >> 2)];
    const int kind = i & 3, y = (int)(e >> UF_XW_BITS), xw = (int)(e & ((1u << UF_XW_BITS) - 1u));
    const bool top = (y % UT_ROWS) == 0, left = (xw % UT_WORDS) == 0, right = (xw % UT_WORDS) == UT_WORDS - 1;
    const bool doW = kind == 0 && left, doN = kind == 1 && top, doNW = kind == 2 && (top || left), doNE = kind == 3 && (top || right);
    if (doW || doN || doNW || doNE) uf_union_word(p, f, y, xw, W32, doW, doN, doNW, doNE);
  }
}

// a run survives iff its root is node 0; then (EXPAND) the final S word goes out as 32 bytes of the u8 {0,255} edge map.
// Body for one plane word; returns true if the word gained edge bits.
// (s, c = the word's S and C values, already loaded: lets a caller keep many loads in flight)
template <bool EXPAND, bool ONLY_CHANGED = false>
__device__ __forceinline__ bool uf_resolve_expand_word_sc(const B2cHystParams &p, int f, int y, int xw, int W32, uint32_t s, const uint32_t c, int *left_root = nullptr)
{
  bool changed = false;
  int nleft = 0, lroot = 0;
  const long long o = f * p.plane_frame_stride + (long long)y * p.plane_pitch + xw;
  uint32_t m = c & ~s;
  if (m) {
    int *P = p.parent + f * p.parent_frame_stride;
    const int base = y * W32 + xw * 32 + 1;
    uint32_t add = 0u;
    while (m) {
      const uint32_t lo = m & (0u - m);
      const uint32_t run = m & ~(m + lo);
      m &= ~run;
      const int root = uf_find_final(P, base + __ffs((int)lo) - 1);
      if (root == 0) add |= run;
      else if (root < 0) { lroot = nleft ? -1 : (root & ~UF_TAG); ++nleft; }   // (an untagged root can never be promoted)
    }
    if (add) {
      s |= add;
      changed = true;
    }
  }
  if (left_root) *left_root = lroot;          // row bands, runs under TAGGED roots: 0 = none, > 0 = the root of the only one, -1 = several
  if (!ONLY_CHANGED || changed) p.E[o] = s;   // E = edges: strong | promoted weak (S itself stays what the stencil wrote)
  if (EXPAND && (!ONLY_CHANGED || changed)) {   // (ONLY_CHANGED: the map already holds the previous state of every word)
    uint8_t *out = p.edges + f * p.edges_frame_stride + (long long)y * p.edges_pitch + xw * 32;
    const int n = min(32, p.w - xw * 32);
    if (n == 32 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
      uint32_t v[8];   // 4 pixels per word: bit k -> byte k = 0 / 255
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = ((((s >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
#ifndef B2C_EMU
      if ((reinterpret_cast<uintptr_t>(out) & 31) == 0) {
        // one 256-bit store per word (sm_100 STG.256): a lane writes its whole 32-byte sector at once; as two 128-bit
        // stores every store instruction of a warp touched 32 half sectors (twice the L1 store transactions)
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                     : "memory");
      } else
#endif
      {
        reinterpret_cast<uint4 *>(out)[0] = make_uint4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<uint4 *>(out)[1] = make_uint4(v[4], v[5], v[6], v[7]);
      }
    } else {
      for (int k = 0; k < n; ++k) out[k] = ((s >> k) & 1u) ? 255 : 0;
    }
  }
  return changed;
}
// One thread per plane word of TWO rows (both rows' loads in flight before either is looked at): block =
// (blockDim.x words) x (2 * blockDim.y rows); grid: x = word blocks of a row, y = row blocks, z = frame.
// LIST (row bands, one frame): the words with unresolved weak runs whose component reaches the band's first or last row
// (tagged roots) are appended to ulist as (word index y * pitch + xw, root of the word's one such run -- or 0 if it has
// several); *ucount may exceed ucap = overflow.  Only these can still be promoted by the seam solve, and
// k_uf_resolve_list then visits these words only.
template <bool EXPAND, bool LIST = false>
__global__ void __launch_bounds__(UFK_THREADS) k_uf_resolve(const B2cHystParams p, int *bcount, uint2 *ulist = nullptr, int *ucount = nullptr, const int ucap = 0)
{
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) bcount[blockIdx.z] = 0;   // the border list of this frame is consumed
  const int wpr = (p.w + 31) >> 5, W32 = p.plane_pitch * 32;
  const int xw = blockIdx.x * blockDim.x + threadIdx.x, y = 2 * (blockIdx.y * blockDim.y + threadIdx.y), f = blockIdx.z;
  int left0 = 0, left1 = 0;
  if (xw < wpr && y < p.h) {
    const long long o = f * p.plane_frame_stride + (long long)y * p.plane_pitch + xw;
    const bool two = y + 1 < p.h;
    const uint32_t s0 = p.S[o], c0 = p.C[o], s1 = two ? p.S[o + p.plane_pitch] : 0u, c1 = two ? p.C[o + p.plane_pitch] : 0u;
    uf_resolve_expand_word_sc<EXPAND>(p, f, y, xw, W32, s0, c0, &left0);
    if (two) uf_resolve_expand_word_sc<EXPAND>(p, f, y + 1, xw, W32, s1, c1, &left1);
  }
  if (LIST) {   // warp-aggregated append (blockDim.x is a multiple of 32: a warp is 32 consecutive words of one row pair)
    const unsigned m0 = __ballot_sync(B2C_FULL, left0 != 0), m1 = __ballot_sync(B2C_FULL, left1 != 0);
    const int n0 = __popc(m0), n = n0 + __popc(m1);
    if (n) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(ucount, n);
      base = __shfl_sync(B2C_FULL, base, 0);
      const unsigned lt = (1u << lane) - 1u;
      const int i0 = base + __popc(m0 & lt), i1 = base + n0 + __popc(m1 & lt);
      if (left0 && i0 < ucap) ulist[i0] = make_uint2((uint32_t)(y * p.plane_pitch + xw), (uint32_t)max(left0, 0));
      if (left1 && i1 < ucap) ulist[i1] = make_uint2((uint32_t)((y + 1) * p.plane_pitch + xw), (uint32_t)max(left1, 0));
    }
  }
}

// Row bands, after the seam solve hung this band's promoted roots under node 0: the listed words again -- one load of
// the recorded root's parent decides a word with one unresolved run -- or, if the list overflowed, every word of the
// band; only words that gain edge pixels are rewritten (E plane and u8 map).
template <bool EXPAND>
__global__ void __launch_bounds__(UFK_THREADS) k_uf_resolve_list(const B2cHystParams p, const uint2 *ulist, const int *ucount, const int ucap, const int *need)
{
  if (need && __ldcg(need) == 0) return;
  const int wpr = (p.w + 31) >> 5, W32 = p.plane_pitch * 32, n = __ldcg(ucount);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  if (n <= ucap) {
    for (int i = tid; i < n; i += nt) {
      const uint2 e = ulist[i];
      if (e.y != 0u && __ldcg(p.parent + e.y - 1) != 0) continue;   // its component was not promoted
      const int wi = (int)e.x, y = wi / p.plane_pitch, xw = wi - y * p.plane_pitch;
      uf_resolve_expand_word_sc<EXPAND, true>(p, 0, y, xw, W32, p.E[wi], p.C[wi]);
    }
  } else {
    for (int i = tid; i < p.h * wpr; i += nt) {
      const int y = i / wpr, xw = i - y * wpr;
      const long long o = (long long)y * p.plane_pitch + xw;
      uf_resolve_expand_word_sc<EXPAND, true>(p, 0, y, xw, W32, p.E[o], p.C[o]);
    }
  }
}

}// namespace b2c
