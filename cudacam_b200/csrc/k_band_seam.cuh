// k_band_seam.cuh -- cross-band hysteresis of the row-band mode in ONE exchange (no rounds).
//
// No reference counterpart (the reference is single-GPU).  After the band-local union-find (k_uf_tile / k_uf_border)
// every band keeps its planes and its forest: edges so far = strong pixels + weak pixels whose root is node 0, U = weak
// pixels whose component touches no strong pixel INSIDE the band.  Whether such a component survives is decided across the seams: it does iff
// it is 8-connected, through unresolved components of any bands, to an edge pixel of some band.  That is a connected-
// components question on a tiny graph -- the unresolved weak RUNS of the first and last row of every band:
//
//   k_seam_publish (one CTA per band): the band's SEAM RECORD = S and U words of its first and last row, the prefix
//       counts of the run starts per word (so that a reader finds a run's ordinal with one popc) + for every unresolved
//       run of these rows the ordinal of the first run with the same root (runs of one component are thereby already
//       united; a component that reaches from the band's first to its last row links the two seams).  With peer wiring
//       the same CTA stores the record into every rank's mailbox and releases the arrival flags;
//   all-gather of the records (a few KB per band; NCCL / gloo, or peer stores into every rank's mailbox + flags);
//   k_seam_solve (one CTA per band, every band solves the same graph; with peer wiring it first waits for the arrival
//       flags): union-find over all runs of all seams with one
//       virtual node 0 = "is an edge": a run is united with 0 if it touches an S pixel across its seam, and with every
//       unresolved run it touches across its seam; then the roots of MY runs that ended up under 0 are hung under node 0
//       of my band's forest;
//   k_uf_resolve_list promotes their components: it visits only the words the band-local resolve listed as still
//       unresolved and rewrites only the words that change (E plane and u8 map).
// The record is published (and pushed to the peers) BEFORE the band-local resolve runs, so that it travels meanwhile.
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
constexpr int SEAM_SHASH = 4096;      // publish: shared-memory hash while the band has <= SEAM_SHASH / 2 boundary runs, else the global one
constexpr int SEAM_SNODES = 12000;    // solve: shared-memory forest while all bands together have fewer runs, else the global one
constexpr int SEAM_TMP = 64;          // scan scratch + small shared scalars

__host__ __device__ inline int seam_cap(int wpr) { return wpr * 16; }                                   // runs per row, worst case
// record: header | S, U of the first row | S, U of the last row | run-start prefixes of both rows | representative per run
__host__ __device__ inline size_t seam_rec_words(int wpr) { return (size_t)SEAM_HDR + 6 * (size_t)wpr + 2 * (size_t)seam_cap(wpr); }
__host__ __device__ inline int seam_hash_size(int wpr)   // power of two >= 4 x the worst-case number of runs of two rows
{
  int n = 1024;
  while (n < 8 * seam_cap(wpr)) n <<= 1;
  return n;
}
// dynamic shared memory of the two kernels
__host__ __device__ inline size_t seam_publish_smem(int wpr) { return ((size_t)2 * wpr + SEAM_TMP + 2 * SEAM_SHASH) * sizeof(int); }
__host__ __device__ inline size_t seam_solve_smem() { return ((size_t)SEAM_TMP + SEAM_SNODES) * sizeof(int); }

struct B2cSeamBand {
  const uint32_t *S, *C;   // strong and candidate planes, row 0 of the band
  int plane_pitch, wpr, h; // u32 per plane row, used words per row, band rows
  int *parent;             // the band's union-find forest (node = y * plane_pitch * 32 + x + 1, 0 = edge)
  int *roots;              // [2 * cap] roots of my boundary runs (first row, then last row), band-private
  int *hkey, *hval;        // global hash: root -> smallest run ordinal with that root (bands with very many boundary runs)
  int hsize;
  int shash, snodes;       // SEAM_SHASH / SEAM_SNODES (tests pass 0 to force the global-memory paths)
  int *ctl;                // [0] "a root of mine was promoted" (out), [1] run id, [2] error (peer time-out), [3] runs promoted, [4] list length
};
struct B2cSeamPeers {
  uint32_t *slot[SEAM_MAXW];   // where MY record goes in the mailbox of every rank (slot[rank] = the record itself)
  uint32_t *flag[SEAM_MAXW];   // my arrival flag in the mailbox of every rank
  int world, rank;             // world < 2: no peers
};
struct B2cSeamAll {
  const uint32_t *rec[SEAM_MAXW];   // record of every band, in band order (rec[rank] = my own)
  int world, rank;
  int *P;                            // global scratch: union-find parents of all runs of all bands (1 + world * 2 * cap ints)
  const uint32_t *flags;             // peer wiring: arrival flags of all ranks in my mailbox (null: the records are there)
  int run_id;
};

#ifndef B2C_EMU
__device__ __forceinline__ void seam_store_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t seam_load_acquire_sys(const uint32_t *p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long seam_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif

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

// CTA-wide exclusive prefix sum of v[0..n) in place, returns the total; `tmp` = 33 ints
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

// union-find of the solve on a forest in shared (SH) or global memory
template <bool SH>
__device__ __forceinline__ int seam_ld(int *P, int i) { return SH ? reinterpret_cast<volatile int *>(P)[i] : __ldcg(P + i); }
template <bool SH>
__device__ __forceinline__ int seam_find(int *P, int n)
{
  while (n != 0) {
    const int pn = seam_ld<SH>(P, n - 1);
    if (pn == n || pn == 0) return pn;
    const int gp = seam_ld<SH>(P, pn - 1);
    if (gp == pn) return pn;
    atomicMin(P + n - 1, gp);
    n = gp;
  }
  return n;
}
template <bool SH>
__device__ __forceinline__ void seam_union(int *P, int a, int b)
{
  for (;;) {
    a = seam_find<SH>(P, a);
    b = seam_find<SH>(P, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }   // a > b >= 0: hang a under b
    const int old = atomicMin(P + a - 1, b);
    if (old == a) return;
    a = old;
  }
}

// ---- publish: one CTA ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SEAM_THREADS) k_seam_publish(const B2cSeamBand b, uint32_t *rec, const int run_id)
{
  B2C_DYN_SMEM(smem);
  int *cnt = reinterpret_cast<int *>(smem);               // [2 * wpr] run starts per word, then their exclusive prefix
  int *tmp = cnt + 2 * b.wpr;                             // [SEAM_TMP]
  int *skey = tmp + SEAM_TMP, *sval = skey + SEAM_SHASH;  // shared hash
  const int tid = threadIdx.x, nt = blockDim.x, wpr = b.wpr, pp = b.plane_pitch, cap = seam_cap(wpr);
  const int W32 = pp * 32;
  uint32_t *rS[2] = { rec + SEAM_HDR, rec + SEAM_HDR + 2 * wpr }, *rU[2] = { rec + SEAM_HDR + wpr, rec + SEAM_HDR + 3 * wpr };
  uint32_t *pre = rec + SEAM_HDR + 4 * wpr, *rep = rec + SEAM_HDR + 6 * wpr;
  if (tid == 0) { b.ctl[0] = 0; b.ctl[3] = 0; b.ctl[4] = 0; }   // ([4]: the unresolved-word list of the resolve pass that follows)
  // 1. edge / unresolved words of the first and last row.  The forest is complete but the band is not resolved yet (the
  // record travels while it is): a weak run is an edge iff its root is node 0.
  for (int i = tid; i < 2 * wpr; i += nt) {
    const int row = i >= wpr, w = i - row * wpr, y = row ? b.h - 1 : 0;
    const long long o = (long long)y * pp + w;
    const uint32_t s = __ldcg(b.S + o), c = __ldcg(b.C + o);
    uint32_t m = c & ~s, add = 0u;
    while (m) {
      const uint32_t lo = m & (0u - m), run = m & ~(m + lo);
      m &= ~run;
      if (uf_find(b.parent, y * W32 + w * 32 + __ffs((int)lo)) == 0) add |= run;
    }
    rS[row][w] = s | add;
    rU[row][w] = c & ~(s | add);
  }
  __syncthreads();
  for (int i = tid; i < 2 * wpr; i += nt) {
    const int row = i >= wpr, w = i - row * wpr;
    cnt[i] = __popc(seam_starts(rU[row][w], w > 0 ? rU[row][w - 1] : 0u));
  }
  __syncthreads();
  const int total = seam_scan(cnt, 2 * wpr, tmp);
  const int ntop = cnt[wpr];                              // prefix at the first word of the last row = runs of the first row
  for (int i = tid; i < 2 * wpr; i += nt) pre[i] = (uint32_t)cnt[i];
  // 2. root of every run -> roots[], smallest ordinal per root -> hash (shared memory unless the rows are very busy)
  const bool sh = 2 * total <= b.shash;
  int *hkey = sh ? skey : b.hkey, *hval = sh ? sval : b.hval;
  const unsigned hmask = (unsigned)((sh ? b.shash : b.hsize) - 1);
  if (total) {
    for (int i = tid; i <= (int)hmask; i += nt) { hkey[i] = SEAM_EMPTY; hval[i] = 0x7FFFFFFF; }
    __syncthreads();
    for (int i = tid; i < 2 * wpr; i += nt) {
      const int row = i >= wpr, w = i - row * wpr, y = row ? b.h - 1 : 0;
      uint32_t st = seam_starts(rU[row][w], w > 0 ? rU[row][w - 1] : 0u);
      int ord = cnt[i];
      while (st) {
        const int bit = __ffs((int)st) - 1;
        st &= st - 1u;
        const int root = uf_find(b.parent, y * W32 + w * 32 + bit + 1);   // != 0: resolved runs are not in U
        b.roots[ord] = root;
        unsigned slot = ((unsigned)root * 2654435761u) & hmask;
        for (;;) {
          const int k = atomicCAS(hkey + slot, SEAM_EMPTY, root);
          if (k == SEAM_EMPTY || k == root) { atomicMin(hval + slot, ord); break; }
          slot = (slot + 1u) & hmask;
        }
        ++ord;
      }
    }
    __threadfence();
    __syncthreads();
    // 3. representative (first run with the same root) of every run
    for (int o = tid; o < total; o += nt) {
      const int root = b.roots[o];
      unsigned slot = ((unsigned)root * 2654435761u) & hmask;
      while (reinterpret_cast<volatile int *>(hkey)[slot] != root) slot = (slot + 1u) & hmask;
      rep[o] = (uint32_t)reinterpret_cast<volatile int *>(hval)[slot];
      b.parent[root - 1] = root | UF_TAG;   // the resolve pass that follows lists the words under tagged roots (uf_find_final)
    }
  }
  if (tid == 0) {
    rec[0] = (uint32_t)run_id;
    rec[1] = (uint32_t)ntop;
    rec[2] = (uint32_t)(total - ntop);
    rec[3] = (uint32_t)wpr;
    rec[4] = (uint32_t)cap;
    b.ctl[1] = run_id;
  }
}

#ifndef B2C_EMU
// peer wiring: the used part of my record (in my own mailbox) -> every other rank's mailbox, then the arrival flags
// (release, system scope; my own flag too).  One CTA, on a side stream: the band resolves meanwhile.
__global__ void __launch_bounds__(SEAM_THREADS) k_seam_push(const B2cSeamPeers q, const int wpr, const int run_id)
{
  const uint32_t *rec = q.slot[q.rank];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int words = SEAM_HDR + 6 * wpr + (int)(__ldcg(rec + 1) + __ldcg(rec + 2));
  for (int i = tid; i < words * q.world; i += nt) {
    const int peer = i / words, k = i - peer * words;
    if (peer != q.rank) q.slot[peer][k] = __ldcg(rec + k);
  }
  __threadfence_system();
  __syncthreads();
  if (tid < q.world) seam_store_release_sys(q.flag[tid], (uint32_t)run_id);
}
#endif

// ---- solve: one CTA ---------------------------------------------------------------------------------------------
// run ordinal (within its band: first-row runs, then last-row runs) of the run that contains bit `bit` of word w of a
// record row; U = the row's U words, pre = the row's exclusive prefixes of run starts
__device__ __forceinline__ int seam_run_of(const uint32_t *U, const uint32_t *pre, int w, int bit)
{
  const uint32_t st = seam_starts(U[w], w > 0 ? U[w - 1] : 0u);
  const uint32_t upto = bit == 31 ? 0xFFFFFFFFu : ((2u << bit) - 1u);
  return (int)pre[w] + __popc(st & upto) - 1;   // a run that came in from word w-1 started before this word: prefix - 1
}

template <bool SH>
__device__ __forceinline__ int seam_solve_body(const B2cSeamBand &b, const B2cSeamAll &a, int *P, const int *base)
{
  const int tid = threadIdx.x, nt = blockDim.x, wpr = b.wpr, world = a.world;
  // 1. every run starts under the first run of its component (runs of one component inside one band)
  for (int r = 0; r < world; ++r) {
    const uint32_t *rep = a.rec[r] + SEAM_HDR + 6 * wpr;
    const int n = base[r + 1] - base[r];
    for (int i = tid; i < n; i += nt) P[base[r] + i] = base[r] + (int)rep[i] + 1;
  }
  if (!SH) __threadfence();
  __syncthreads();
  // 2. the seams: last row of band s against first row of band s+1, one thread per plane word
  for (int idx = tid; idx < (world - 1) * wpr; idx += nt) {
    const int s = idx / wpr, w = idx - s * wpr;
    const uint32_t *SA = a.rec[s] + SEAM_HDR + 2 * wpr, *UA = a.rec[s] + SEAM_HDR + 3 * wpr;          // last row of band s
    const uint32_t *SB = a.rec[s + 1] + SEAM_HDR, *UB = a.rec[s + 1] + SEAM_HDR + wpr;                 // first row of band s+1
    const uint32_t *preA = a.rec[s] + SEAM_HDR + 5 * wpr, *preB = a.rec[s + 1] + SEAM_HDR + 4 * wpr;
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
      seam_union<SH>(P, nodeA0 + seam_run_of(UA, preA, w, bit), 0);
    }
    m = ub & dil(SA);
    while (m) {
      const int bit = __ffs((int)m) - 1;
      m &= ~seam_frag(ub, bit);
      seam_union<SH>(P, nodeB0 + seam_run_of(UB, preB, w, bit), 0);
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
        seam_union<SH>(P, na, nodeB0 + seam_run_of(UB, preB, w, b2));
      }
      if ((fa & 1u) && w > 0 && (UB[w - 1] >> 31)) seam_union<SH>(P, na, nodeB0 + seam_run_of(UB, preB, w - 1, 31));
      if ((fa >> 31) && w + 1 < wpr && (UB[w + 1] & 1u)) seam_union<SH>(P, na, nodeB0 + seam_run_of(UB, preB, w + 1, 0));
    }
  }
  if (!SH) __threadfence();
  __syncthreads();
  // 3. my runs whose component reached an edge: hang their band-local roots under node 0 of my forest
  const int mine = base[a.rank + 1] - base[a.rank];
  int promoted = 0;
  for (int i = tid; i < mine; i += nt) {
    if (seam_find<SH>(P, base[a.rank] + i + 1) == 0) {
      const int root = b.roots[i];
      if (root != 0 && atomicExch(b.parent + root - 1, 0) != 0) ++promoted;   // (the entry held the tagged root itself)
    }
  }
  return promoted;
}

__global__ void __launch_bounds__(SEAM_THREADS) k_seam_solve(const B2cSeamBand b, const B2cSeamAll a)
{
  B2C_DYN_SMEM(smem);
  int *tmp = reinterpret_cast<int *>(smem);   // [SEAM_TMP]: [0..SEAM_MAXW] first node of every band, [32] promoted runs
  int *sP = tmp + SEAM_TMP;                   // [SEAM_SNODES] forest
  const int tid = threadIdx.x, world = a.world;
#ifndef B2C_EMU
  if (a.flags && tid < world) {   // peer wiring: wait until the records of this run of all ranks have landed in my mailbox
    const unsigned long long t0 = seam_now();
    while (seam_load_acquire_sys(a.flags + tid) != (uint32_t)a.run_id) {
      if (seam_now() - t0 > 2000000000ull) { b.ctl[2] = 1; break; }   // 2 s: a peer never arrived
    }
  }
  __syncthreads();
#endif
  if (tid == 0) {
    int acc = 0;
    for (int r = 0; r < world; ++r) {
      tmp[r] = acc;
      acc += (int)__ldcg(a.rec[r] + 1) + (int)__ldcg(a.rec[r] + 2);
    }
    tmp[world] = acc;
    tmp[32] = 0;
  }
  __syncthreads();
  const int total = tmp[world];
  // node n (1 .. total) lives at P[n-1]; node 0 = "is an edge"
  const int promoted = total < b.snodes ? seam_solve_body<true>(b, a, sP, tmp) : seam_solve_body<false>(b, a, a.P, tmp);
  if (promoted) atomicAdd(tmp + 32, promoted);
  __syncthreads();
  if (tid == 0) {
    b.ctl[3] = tmp[32];
    b.ctl[0] = tmp[32] ? 1 : 0;
  }
}
}// namespace b2c
