// k_band_p2p.cuh -- cross-band hysteresis rounds of the row-band mode WITHOUT the host in the loop.
//
// No reference counterpart (the reference is single-GPU).  After the band-local union-find every rank keeps its
// planes and forest; what remains is the exchange "my first / last row of the edge bit plane -> the neighbour's ghost
// row, re-seed, re-resolve" until no rank is seeded anything new.  Driven from the host with NCCL this costs two
// collective launches and one blocking read per round (~200 us per round at 2 GPUs, milliseconds at 8 ranks on a
// 16-core host); here every rank has the other ranks' MAILBOXES mapped (CUDA IPC, peer stores over NVLink):
//
//   k_band_rounds (one cooperative launch per run), per round r:
//     exchange  (block 0) my boundary rows -> the two neighbours' ghost slots of parity r&1, then my "my last seeding
//               found something" flag -> ALL ranks' flag slots of round r (release, system scope);
//     wait      (block 0) until the flags of round r of all ranks have arrived (acquire, system scope); if nobody was
//               seeded anything the rank is done -- every rank takes the same decision from the same flags;
//     seed      (grid) the weak runs of the boundary rows that touch a strong ghost pixel are hung under node 0;
//     resolve   (grid, only if something was seeded) promotes the components and rewrites the u8 map.
//
// The host reads the control words once per launch (done / error / rounds).
// Flag words carry the round number ((round << 2) | state), rounds are numbered by a device-side counter that never
// goes back, so no slot ever has to be cleared.  Every spin has a time-out (2 s): a rank that never arrives sets
// the error flag instead of hanging the GPU.
#pragma once
#include "b2c_device.cuh"
#include "k_hysteresis_uf.cuh"

namespace b2c
{
constexpr int BP_MAXW = 16;   // ranks of one box
constexpr int BP_MAXR = 32;   // flag slots (ranks are never more than one round apart)
// control ints (own device memory): round counter, done, current round, seeded-in-last-round, rounds of this run, error
enum { BP_ROUNDS = 0, BP_DONE = 1, BP_CUR = 2, BP_SEEDED = 3, BP_RUN_ROUNDS = 4, BP_ERROR = 5 };

struct B2cBandP2P {
  uint32_t *mail[BP_MAXW];   // mailbox of every rank (mail[rank] = own), see bp_* offsets
  int world, rank, wpr;
  int *ctl;
  // input halo exchange: the neighbours' band input buffers (rows 0..3 = halo above, 4..4+rows-1 = band, then 4 halo rows)
  uint8_t *in_up, *in_dn, *in_own;
  long long in_stride;
  int rows_own, rows_up;
  int row_bytes;
};
__host__ __device__ inline int bp_ghost(int wpr, int parity, int which) { return (parity * 2 + which) * wpr; }   // which: 0 top, 1 bottom
__host__ __device__ inline int bp_flag(int wpr, int round, int rank) { return 4 * wpr + (round % BP_MAXR) * BP_MAXW + rank; }
__host__ __device__ inline int bp_halo_flag(int wpr, int from_below) { return 4 * wpr + BP_MAXR * BP_MAXW + from_below; }   // blocks of halo rows received so far (never reset)
__host__ __device__ inline size_t bp_mailbox_words(int wpr) { return (size_t)4 * wpr + BP_MAXR * BP_MAXW + 4; }

#ifndef B2C_EMU
__device__ __forceinline__ void bp_store_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t bp_load_acquire_sys(const uint32_t *p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long bp_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// All rounds of one run in ONE cooperative launch: no launch gaps between exchange, seeding and resolving (three
// launches per round cost ~60 us per round at 8 ranks; here a round is a few grid barriers).  Block 0 does the
// exchange and the wait, the whole grid the seeding and the resolve pass.
template <bool EXPAND>
__global__ void __launch_bounds__(256) k_band_rounds(const B2cHystParams p, const B2cBandP2P q, const int max_rounds)
{
  __shared__ int r_s;
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int tid = threadIdx.x;
  const int wpr = q.wpr, W32 = p.plane_pitch * 32, pp = p.plane_pitch;
  const long long gtid = (long long)blockIdx.x * blockDim.x + tid, gthreads = (long long)gridDim.x * blockDim.x;
  uint32_t *mine = q.mail[q.rank];
  // time stamps of block 0 (ns, low 32 bits) for b2c_get_info("p2p_stamp<k>"): [0] start, then per round: flags sent,
  // flags of all ranks seen, seeding done, resolve done
  int *stamp = q.ctl + 8;
  int ns = 0;
#define BP_STAMP() do { if (blockIdx.x == 0 && tid == 0 && ns < 56) stamp[ns++] = (int)(unsigned)bp_now(); } while (0)
  BP_STAMP();
  for (int it = 0; it < max_rounds; ++it) {
    if (blockIdx.x == 0) {
      // ---- exchange: my boundary rows -> the neighbours' ghost slots, then my flag -> every rank
      if (tid == 0) {
        r_s = ++q.ctl[BP_ROUNDS];
        q.ctl[BP_CUR] = r_s;
        q.ctl[BP_RUN_ROUNDS] += 1;
      }
      __syncthreads();
      const int r = r_s, par = r & 1;
      const uint32_t *top = p.S, *bot = p.S + (long long)(p.h - 1) * pp;
      if (q.rank > 0) {
        uint32_t *dst = q.mail[q.rank - 1] + bp_ghost(wpr, par, 1);   // my first row is the upper neighbour's bottom ghost row
        for (int i = tid; i < wpr; i += blockDim.x) dst[i] = __ldcg(top + i);
      }
      if (q.rank + 1 < q.world) {
        uint32_t *dst = q.mail[q.rank + 1] + bp_ghost(wpr, par, 0);
        for (int i = tid; i < wpr; i += blockDim.x) dst[i] = __ldcg(bot + i);
      }
      __threadfence_system();
      __syncthreads();
      if (tid < q.world) bp_store_release_sys(q.mail[tid] + bp_flag(wpr, r, q.rank), ((uint32_t)r << 2) | (__ldcg(q.ctl + BP_SEEDED) ? 2u : 1u));
      __syncthreads();
      BP_STAMP();
      // ---- wait for the flags of round r of all ranks; any == 1: somebody was seeded something new in the last round
      if (tid == 0) {
        q.ctl[BP_SEEDED] = 0;
        int any = 0;
        const unsigned long long t0 = bp_now();
        for (int k = 0; k < q.world && any >= 0; ++k) {
          uint32_t v;
          while (((v = bp_load_acquire_sys(mine + bp_flag(wpr, r, k))) >> 2) != (uint32_t)r) {
            if (bp_now() - t0 > 2000000000ull) { q.ctl[BP_ERROR] = 1; any = -1; break; }
          }
          if (any >= 0) any |= (v & 3u) == 2u;
        }
        if (any == 0) q.ctl[BP_DONE] = 1;
        BP_STAMP();
        __stcg(q.ctl + 7, any);
        __threadfence();
      }
    }
    grid.sync();
    if (__ldcg(q.ctl + 7) <= 0) break;   // converged (or a peer never arrived)
    // ---- seeding: weak runs of the first / last band row that touch a strong ghost pixel hang their root under node 0
    const int par = __ldcg(q.ctl + BP_CUR) & 1;
    for (long long i = gtid; i < 2ll * wpr; i += gthreads) {
      const int which = i >= wpr, xw = (int)(i - (which ? wpr : 0));
      if ((which == 0 && q.rank == 0) || (which == 1 && q.rank + 1 == q.world)) continue;   // image border: no neighbour
      const int y = which ? p.h - 1 : 0;
      const long long o = (long long)y * pp + xw;
      const uint32_t wd = __ldcg(p.C + o) & ~__ldcg(p.S + o);
      if (wd == 0u) continue;
      const uint32_t *G = mine + bp_ghost(wpr, par, which) + xw;
      const uint32_t g = __ldcg(G), gl = xw > 0 ? __ldcg(G - 1) : 0u, gr = xw + 1 < wpr ? __ldcg(G + 1) : 0u;
      const uint32_t near = wd & (g | (g << 1) | (g >> 1) | (gl >> 31) | (gr << 31));
      if (near == 0u) continue;
      const int base = y * W32 + xw * 32 + 1;
      uint32_t m = wd;
      while (m) {
        const uint32_t lo = m & (0u - m);
        const uint32_t run = m & ~(m + lo);
        m &= ~run;
        if (run & near) {
          const int root = uf_find(p.parent, base + __ffs((int)lo) - 1);
          if (root != 0) {
            atomicMin(p.parent + root - 1, 0);
            if (__ldcg(q.ctl + BP_SEEDED) == 0) __stcg(q.ctl + BP_SEEDED, 1);
          }
        }
      }
    }
    __threadfence();
    grid.sync();
    BP_STAMP();
    // ---- resolve (and rewrite the u8 map) if this band was seeded anything new
    if (__ldcg(q.ctl + BP_SEEDED)) {
      // 8 rows per block and step, all 16 loads of a thread in flight before the first word is looked at (the pass is
      // latency-bound: with one word at a time it took 137 us for a 8192-row band); the first pass
      // (b2c_band_hysteresis) wrote the whole map, so only changed words are rewritten
      constexpr int RB = 8;
      for (int yb = blockIdx.x * RB; yb < p.h; yb += gridDim.x * RB)
        for (int xw = tid; xw < wpr; xw += blockDim.x) {
          uint32_t sv[RB], cv[RB];
#pragma unroll
          for (int k = 0; k < RB; ++k) {
            const bool ok = yb + k < p.h;
            sv[k] = ok ? __ldcg(p.S + (long long)(yb + k) * pp + xw) : 0u;
            cv[k] = ok ? __ldcg(p.C + (long long)(yb + k) * pp + xw) : 0u;
          }
#pragma unroll
          for (int k = 0; k < RB; ++k)
            if (cv[k] & ~sv[k]) uf_resolve_expand_word_sc<EXPAND, true>(p, 0, yb + k, xw, W32, sv[k], cv[k]);
        }
      __threadfence();
    }
    grid.sync();
    BP_STAMP();
  }
  if (blockIdx.x == 0 && tid == 0) stamp[63 - 8] = ns;
#undef BP_STAMP
}

// Input halo: my first 4 rows -> the upper neighbour's 4 halo rows below its band, my last 4 rows -> the lower
// neighbour's 4 halo rows above its band (128-bit peer stores).  Every block then adds 1 to the neighbour's halo
// counter (system-scope atomic after a system fence): the neighbour waits for run * blocks_per_direction.
// Grid: x = blocks over the 16-byte chunks of 4 rows, y = 0 (to the upper neighbour) / 1 (to the lower one).
__global__ void __launch_bounds__(256) k_band_push_halo(const B2cBandP2P q)
{
  const int up = blockIdx.y == 0;
  if ((up && q.rank == 0) || (!up && q.rank + 1 == q.world)) return;
  const int cpr = (q.row_bytes + 15) / 16;   // 16-byte chunks per row (rows are 16-byte aligned and padded)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * cpr) {
    const int r = i / cpr, c = i - r * cpr;
    const uint8_t *src = q.in_own + (long long)(up ? 4 + r : q.rows_own + r) * q.in_stride + 16 * c;
    uint8_t *dst = up ? q.in_up + (long long)(4 + q.rows_up + r) * q.in_stride + 16 * c : q.in_dn + (long long)r * q.in_stride + 16 * c;
    *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src);
  }
  __threadfence_system();
  __syncthreads();
  // my rows arrive "from below" at the upper neighbour (counter 1) and "from above" at the lower one (counter 0)
  if (threadIdx.x == 0) atomicAdd_system(reinterpret_cast<int *>((up ? q.mail[q.rank - 1] : q.mail[q.rank + 1]) + bp_halo_flag(q.wpr, up ? 1 : 0)), 1);
}

// waits until both neighbours' halo rows of run `run` have landed
__global__ void k_band_wait_halo(const B2cBandP2P q, const int run, const int nblocks)
{
  const unsigned long long t0 = bp_now();
  for (int d = 0; d < 2; ++d) {
    if ((d == 0 && q.rank == 0) || (d == 1 && q.rank + 1 == q.world)) continue;   // d = 0: rows from above, 1: from below
    const uint32_t *f = q.mail[q.rank] + bp_halo_flag(q.wpr, d);
    while ((int)bp_load_acquire_sys(f) < run * nblocks) {
      if (bp_now() - t0 > 2000000000ull) { q.ctl[BP_ERROR] = 1; return; }
    }
  }
}

#endif
}// namespace b2c
