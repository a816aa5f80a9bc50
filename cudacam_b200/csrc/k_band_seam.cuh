// k_band_seam.cuh -- cross-band hysteresis of the row-band mode in ONE exchange (no rounds).
//
// No reference counterpart (the reference is single-GPU).  After the band-local union-find (k_uf_tile / k_uf_border /
// k_uf_resolve) every band keeps its planes and its forest: S = edges so far, U = C & ~S = weak pixels whose component
// touches no strong pixel INSIDE the band.  Whether such a component survives is decided across the seams: it does iff
// it is 8-connected, through unresolved components of any bands, to an edge pixel of some band.  That is a connected-
// components question on a tiny graph -- the unresolved weak RUNS of the first and last row of every band:
//
//   k_seam_publish (one CTA per band): the band's SEAM RECORD = S and U words of its first and last row + for every
//       unresolved run of these rows the ordinal of the first run with the same root (runs of one component are thereby
//       already united; a component that reaches from the band's first to its last row links the two seams);
//   all-gather of the records (a few KB per band; NCCL / gloo, or peer stores into every rank's mailbox + flags);
//   k_seam_solve (one CTA per band, every band solves the same graph): union-find over all runs of all seams with one
//       virtual node 0 = "is an edge": a run is united with 0 if it touches an S pixel across its seam, and with every
//       unresolved run it touches across its seam; then the roots of MY runs that ended up under 0 are hung under node 0
//       of my band's forest;
//   one k_uf_resolve over the band promotes their components (only changed words of the u8 map are rewritten).
//
// Same fixpoint as the reference's iteration (src/cvp/cannyEdgeD.cu:295-377 driven by cannyEdgeH.cu:297-338) on the
// unsharded image, so the sharded edge map is bit-identical; the earlier protocol needed one globally synchronous round
// per seam crossing of the longest weak chain (7 rounds of ~50 us on the 16384^2 mosaic at 8 GPUs).
#pragma once
#include "b2c_device.cuh"
#include "k_hysteresis_uf.cuh"

namespace b2c
{
constexpr int SEAM_HDR = 8;           // header words: [0] run id, [1] runs in the first row, [2] runs in the last row, [3] wpr
constexpr int SEAM_THREADS = 1024;
constexpr int SEAM_MAXW = 16;         // bands
constexpr int SEAM_EMPTY = 0;         // hash key of an empty slot (roots are >= 1)
__host__ __device__ inline size_t seam_smem_bytes(int wpr) { return ((size_t)2 * wpr + 64) * sizeof(int); }   // dynamic shared memory of both kernels

__host__ __device__ inline int seam_cap(int wpr) { return wpr * 16; }                                   // runs per row, worst case
__host__ __device__ inline size_t seam_rec_words(int wpr) { return (size_t)SEAM_HDR + 4 * (size_t)wpr + 2 * (size_t)seam_cap(wpr); }
__host__ __device__ inline int seam_hash_size(int wpr)   // power of two >= 4 x the worst-case number of runs of two rows
{
  int n = 1024;
  while (n < 8 * seam_cap(wpr)) n <<= 1;
  return n;
}

struct B2cSeamBand {
  const uint32_t *S, *C;   // planes, row 0 of the band
  int plane_pitch, wpr, h; // u32 per plane row, used words per row, band rows
  int *parent;             // the band's union-find forest (node = y * plane_pitch * 32 + x + 1, 0 = edge)
  int *roots;              // [2 * cap] roots of my boundary runs (first row, then last row), band-private
  int *hkey, *hval;        // hash: root -> smallest run ordinal with that root
  int hsize;
  int *ctl;                // [0] "a root of mine was promoted" (out), [1] run id, [2] error (peer time-out), [3] runs promoted
};

// starts of the unresolved runs of one plane row inside word w: a run that continues from word w-1 has no start here
__device__ __forceinline__ uint32_t seam_starts(uint32_t u, uint32_t u_prev) { return u & ~((u << 1) | (u_prev >> 31)); }
// the contiguous run of ones of x that contains bit b (x has bit b set), restricted to the word
__device__ __forceinline__ uint32_t seam_frag(uint32_t x, int b)
{
  const uint32_t below = ~x & ((1u << b) - 1u);                       // zeros below b
  const int lo = below ? 32 - __clz((int)below) : 0;                   // first bit of the run
  const uint32_t above = ~x & ~((2u << b) - 1u) & (b == 31 ? 0u : 0xFFFFFFFFu);
  const int hi = above ? __ffs((int)above) - 1 : 32;                   // one past the last bit
  const uint32_t m_hi = hi == 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u);
  return m_hi & ~((1u << lo) - 1u);
}

// CTA-wide exclusive prefix sum of v[0..n) in place (n <= 2 * 4096), returns the total; `tmp` = SEAM_THREADS / 32 ints
__device__ inline int seam_scan(int *v, int n, int *tmp)
{
  const int tid = threadIdx.x, nt = blockDim.x;
  const int per = (n + nt - 1) / nt, b0 = tid * per;
  int sum = 0;
  for (int i = b0; i < min(b0 + per, n); ++i) sum += v[i];
  // scan of the per-thread sums: warp scan + scan of the warp totals
  const int lane = tid & 31, warp = tid >> 5;
  int incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(B2C_FULL, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) tmp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (nt >> 5) ? tmp[lane] : 0, wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(B2C_FULL, wi, d);
      if (lane >= d) wi += o;
    }
    tmp[lane] = wi - w;                                   // exclusive warp offsets
    if (lane == 31) tmp[32] = wi;                         // grand total
  }
  __syncthreads();
  int run = tmp[warp] + incl - sum;
  for (int i = b0; i < min(b0 + per, n); ++i) {
    const int x = v[i];
    v[i] = run;
    run += x;
  }
  __syncthreads();
  return tmp[32];
}

// ---- publish: one CTA -------------------------------------------------------------------------------------------
// rec = where the record is built (device memory of this band).  peers / npeers: if given, the finished record is also
// stored into these (peer-mapped) addresses and flag words are released afterwards (peer-to-peer all-gather).
__global__ void __launch_bounds__(SEAM_THREADS) k_seam_publish(const B2cSeamBand b, uint32_t *rec, const int run_id)
{
  B2C_DYN_SMEM(smem);
  int *cnt = reinterpret_cast<int *>(smem);               // [2 * wpr] run starts per word, then their exclusive prefix
  int *tmp = cnt + 2 * b.wpr;                             // [40]
  const int tid = threadIdx.x, nt = blockDim.x, wpr = b.wpr, pp = b.plane_pitch, cap = seam_cap(wpr);
  const int W32 = pp * 32;
  uint32_t *rS[2] = { rec + SEAM_HDR, rec + SEAM_HDR + 2 * wpr }, *rU[2] = { rec + SEAM_HDR + wpr, rec + SEAM_HDR + 3 * wpr };
  uint32_t *rep = rec + SEAM_HDR + 4 * wpr;
  for (int i = tid; i < b.hsize; i += nt) { b.hkey[i] = SEAM_EMPTY; b.hval[i] = 0x7FFFFFFF; }
  if (tid == 0) { b.ctl[0] = 0; b.ctl[3] = 0; }
  // 1. S / U words of the first and last row, run starts per word
  for (int i = tid; i < 2 * wpr; i += nt) {
    const int row = i >= wpr, w = i - row * wpr;
    const long long o = (long long)(row ? b.h - 1 : 0) * pp + w;
    const uint32_t s = __ldcg(b.S + o), u = __ldcg(b.C + o) & ~s;
    const uint32_t up = w > 0 ? (__ldcg(b.C + o - 1) & ~__ldcg(b.S + o - 1)) : 0u;
    rS[row][w] = s;
    rU[row][w] = u;
    cnt[i] = __popc(seam_starts(u, up));
  }
  __syncthreads();
  const int total = seam_scan(cnt, 2 * wpr, tmp);
  const int ntop = cnt[wpr];                              // prefix at the first word of the last row = runs of the first row
  // 2. root of every run -> roots[], smallest ordinal per root -> hash
  for (int i = tid; i < 2 * wpr; i += nt) {
    const int row = i >= wpr, w = i - row * wpr, y = row ? b.h - 1 : 0;
    const uint32_t u = rU[row][w], up = w > 0 ? rU[row][w - 1] : 0u;
    uint32_t st = seam_starts(u, up);
    int ord = cnt[i];
    while (st) {
      const int bit = __ffs((int)st) - 1;
      st &= st - 1u;
      const int root = uf_find(b.parent, y * W32 + w * 32 + bit + 1);   // != 0: resolved runs are in S, not in U
      b.roots[ord] = root;
      unsigned slot = ((unsigned)root * 2654435761u) & (unsigned)(b.hsize - 1);
      for (;;) {
        const int k = atomicCAS(b.hkey + slot, SEAM_EMPTY, root);
        if (k == SEAM_EMPTY || k == root) { atomicMin(b.hval + slot, ord); break; }
        slot = (slot + 1u) & (unsigned)(b.hsize - 1);
      }
      ++ord;
    }
  }
  __syncthreads();
  // 3. representative (first run with the same root) of every run
  for (int o = tid; o < total; o += nt) {
    const int root = b.roots[o];
    unsigned slot = ((unsigned)root * 2654435761u) & (unsigned)(b.hsize - 1);
    while (b.hkey[slot] != root) slot = (slot + 1u) & (unsigned)(b.hsize - 1);
    rep[o] = (uint32_t)b.hval[slot];
  }
  if (tid == 0) {
    rec[1] = (uint32_t)ntop;
    rec[2] = (uint32_t)(total - ntop);
    rec[3] = (uint32_t)wpr;
    rec[4] = (uint32_t)cap;
    b.ctl[1] = run_id;
  }
  __syncthreads();
  if (tid == 0) { __threadfence(); rec[0] = (uint32_t)run_id; }
}

// ---- solve: one CTA ---------------------------------------------------------------------------------------------
struct B2cSeamAll {
  const uint32_t *rec[SEAM_MAXW];   // record of every band, in band order (rec[rank] = my own)
  int world, rank;
  int *P;                            // scratch: union-find parents of all runs of all bands (1 + world * 2 * cap ints)
  int *pre;                          // scratch: exclusive run-start prefixes per (band, row, word): world * 2 * wpr ints
};

// run ordinal (within its band: first-row runs, then last-row runs) of the run that contains bit `bit` of word w of a
// record row; U = the row's U words, pre = the row's exclusive prefixes of run starts
__device__ __forceinline__ int seam_run_of(const uint32_t *U, const int *pre, int w, int bit)
{
  const uint32_t st = seam_starts(U[w], w > 0 ? U[w - 1] : 0u);
  const uint32_t upto = bit == 31 ? 0xFFFFFFFFu : ((2u << bit) - 1u);
  return pre[w] + __popc(st & upto) - 1;   // a run that came in from word w-1 started before this word: prefix - 1
}

__global__ void __launch_bounds__(SEAM_THREADS) k_seam_solve(const B2cSeamBand b, const B2cSeamAll a)
{
  B2C_DYN_SMEM(smem);
  int *cnt = reinterpret_cast<int *>(smem);   // [2 * wpr] scan workspace
  int *tmp = cnt + 2 * b.wpr;                 // [40]
  int *base = tmp + 40;                       // [SEAM_MAXW + 1] first node of every band
  int &changed_s = tmp[40 + SEAM_MAXW + 1];
  const int tid = threadIdx.x, nt = blockDim.x, wpr = b.wpr, world = a.world;
  if (tid == 0) {
    int acc = 0;
    for (int r = 0; r < world; ++r) {
      base[r] = acc;
      acc += (int)a.rec[r][1] + (int)a.rec[r][2];
    }
    base[world] = acc;
    changed_s = 0;
  }
  __syncthreads();
  const int total = base[world];
  int *P = a.P;   // node n (1 .. total) lives at P[n-1]; node 0 = "is an edge"
  // 1. every run starts under the first run of its component (runs of one component inside one band)
  for (int r = 0; r < world; ++r) {
    const uint32_t *rep = a.rec[r] + SEAM_HDR + 4 * wpr;
    const int n = base[r + 1] - base[r];
    for (int i = tid; i < n; i += nt) P[base[r] + i] = base[r] + (int)rep[i] + 1;
  }
  // 2. exclusive prefixes of the run starts of every record row (first row: row 0, last row: row 1 of the record)
  for (int r = 0; r < world; ++r) {
    const uint32_t *U0 = a.rec[r] + SEAM_HDR + wpr, *U1 = a.rec[r] + SEAM_HDR + 3 * wpr;
    for (int i = tid; i < 2 * wpr; i += nt) {
      const int row = i >= wpr, w = i - row * wpr;
      const uint32_t *U = row ? U1 : U0;
      cnt[i] = __popc(seam_starts(U[w], w > 0 ? U[w - 1] : 0u));
    }
    __syncthreads();
    seam_scan(cnt, 2 * wpr, tmp);
    for (int i = tid; i < 2 * wpr; i += nt) a.pre[(long long)r * 2 * wpr + i] = cnt[i];   // (the last-row prefixes already include ntop)
    __syncthreads();
  }
  __threadfence();
  __syncthreads();
  // 3. the seams: last row of band s against first row of band s+1, one thread per plane word
  for (int idx = tid; idx < (world - 1) * wpr; idx += nt) {
    const int s = idx / wpr, w = idx - s * wpr;
    const uint32_t *SA = a.rec[s] + SEAM_HDR + 2 * wpr, *UA = a.rec[s] + SEAM_HDR + 3 * wpr;          // last row of band s
    const uint32_t *SB = a.rec[s + 1] + SEAM_HDR, *UB = a.rec[s + 1] + SEAM_HDR + wpr;                 // first row of band s+1
    const int *preA = a.pre + (long long)s * 2 * wpr + wpr, *preB = a.pre + (long long)(s + 1) * 2 * wpr;
    const int nodeA0 = base[s] + 1, nodeB0 = base[s + 1] + 1;
    const uint32_t ua = UA[w], ub = UB[w];
    if ((ua | ub) == 0u) continue;
    auto dil = [&](const uint32_t *X) -> uint32_t {   // 3-dilation along the row of the OTHER side's word w
      const uint32_t x = X[w], xl = w > 0 ? X[w - 1] : 0u, xr = w + 1 < wpr ? X[w + 1] : 0u;
      return x | (x << 1) | (x >> 1) | (xl >> 31) | (xr << 31);
    };
    // unresolved runs that touch an edge pixel across the seam
    uint32_t m = ua & dil(SB);
    while (m) {
      const int bit = __ffs((int)m) - 1;
      m &= ~seam_frag(ua, bit);
      uf_union(P, nodeA0 + seam_run_of(UA, preA, w, bit), 0);
    }
    m = ub & dil(SA);
    while (m) {
      const int bit = __ffs((int)m) - 1;
      m &= ~seam_frag(ub, bit);
      uf_union(P, nodeB0 + seam_run_of(UB, preB, w, bit), 0);
    }
    // unresolved runs that touch each other across the seam (every touching pair is seen from the A side)
    m = ua & dil(UB);
    while (m) {
      const int bit = __ffs((int)m) - 1;
      const uint32_t fa = seam_frag(ua, bit);
      m &= ~fa;
      const int na = nodeA0 + seam_run_of(UA, preA, w, bit);
      uint32_t pb = (fa | (fa << 1) | (fa >> 1)) & ub;
      while (pb) {
        const int b2 = __ffs((int)pb) - 1;
        pb &= ~seam_frag(ub, b2);
        uf_union(P, na, nodeB0 + seam_run_of(UB, preB, w, b2));
      }
      if ((fa & 1u) && w > 0 && (UB[w - 1] >> 31)) uf_union(P, na, nodeB0 + seam_run_of(UB, preB, w - 1, 31));
      if ((fa >> 31) && w + 1 < wpr && (UB[w + 1] & 1u)) uf_union(P, na, nodeB0 + seam_run_of(UB, preB, w + 1, 0));
    }
  }
  __threadfence();
  __syncthreads();
  // 4. my runs whose component reached an edge: hang their band-local roots under node 0 of my forest
  const int mine = base[a.rank + 1] - base[a.rank];
  int promoted = 0;
  for (int i = tid; i < mine; i += nt) {
    if (uf_find(P, base[a.rank] + i + 1) == 0) {
      const int root = b.roots[i];
      if (root != 0 && atomicMin(b.parent + root - 1, 0) != 0) ++promoted;
    }
  }
  if (promoted) { atomicAdd(&changed_s, promoted); }
  __syncthreads();
  if (tid == 0) {
    b.ctl[3] = changed_s;
    __threadfence();
    b.ctl[0] = changed_s ? 1 : 0;
  }
}

// ---- peer-to-peer all-gather of the records (ranks of one box) ---------------------------------------------------
#ifndef B2C_EMU
__device__ __forceinline__ void seam_store_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t seam_load_acquire_sys(const uint32_t *p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long seam_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

struct B2cSeamPeers {
  uint32_t *slot[SEAM_MAXW];   // where MY record goes in the mailbox of every rank (slot[rank] = my own copy: the source)
  uint32_t *flag[SEAM_MAXW];   // my arrival flag in the mailbox of every rank
  int world, rank;
};
// my record (already built in my own mailbox) -> every other rank's mailbox, then the flags (release, system scope)
__global__ void __launch_bounds__(SEAM_THREADS) k_seam_push(const B2cSeamPeers q, const int wpr, const int run_id)
{
  const uint32_t *src = q.slot[q.rank];
  const int nruns = (int)(src[1] + src[2]);
  const int words = SEAM_HDR + 4 * wpr + nruns;   // only the used part of the record travels
  const int peer = blockIdx.x;
  if (peer != q.rank) {
    uint32_t *dst = q.slot[peer];
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) seam_store_release_sys(q.flag[peer], (uint32_t)run_id);
}
// waits until the records of run `run_id` of all ranks have landed in my mailbox (2 s time-out -> ctl[2] = 1)
__global__ void k_seam_wait(const uint32_t *flags, const int world, const int run_id, int *ctl)
{
  const unsigned long long t0 = seam_now();
  const int r = threadIdx.x;
  if (r < world) {
    while (seam_load_acquire_sys(flags + r) != (uint32_t)run_id) {
      if (seam_now() - t0 > 2000000000ull) { ctl[2] = 1; break; }
    }
  }
}
#endif
}// namespace b2c
