// k_band_p2p.cuh -- peer-memory plumbing of the row-band mode for ranks of one box: mailbox layout and the exchange of
// the 4 input halo rows.
//
// No reference counterpart (the reference is single-GPU).  Every rank owns a MAILBOX in device memory that the other
// ranks map (CUDA IPC between processes, plain pointers inside one process) and write with ordinary stores over NVLink:
//   seam records : 2 parities x BP_MAXW slots of seam_rec_words() words -- rank r stores its record of run k into slot
//                  (k & 1, r) of every mailbox (k_band_seam.cuh); two parities suffice because a rank can publish run
//                  k+2 only after every rank has published run k+1, i.e. has finished reading run k;
//   seam flags   : 2 x BP_MAXW words, flag (k & 1, r) = k once the record has landed (release / acquire, system scope);
//   halo counters: 2 words counting the blocks of halo rows received from above / below (never reset).
// Every spin has a 2 s time-out that sets the error word instead of hanging the GPU; the error is cleared at the start
// of the next run (b2c_band_p2p_halo).
#pragma once
#include "b2c_device.cuh"
#include "k_band_seam.cuh"

namespace b2c
{
constexpr int BP_MAXW = SEAM_MAXW;   // ranks of one box

struct B2cBandP2P {
  uint32_t *mail[BP_MAXW];   // mailbox of every rank (mail[rank] = own)
  int world, rank, wpr;
  int *ctl;                  // own control ints: [2] = error (a peer never arrived)
  // input halo exchange: the neighbours' band input buffers (rows 0..3 = halo above, 4..4+rows-1 = band, then 4 halo rows)
  uint8_t *in_up, *in_dn, *in_own;
  long long in_stride;
  int rows_own, rows_up;
  int row_bytes;
};
__host__ __device__ inline size_t bp_seam_slot(int wpr, int parity, int rank) { return ((size_t)parity * BP_MAXW + rank) * seam_rec_words(wpr); }
__host__ __device__ inline size_t bp_seam_flag(int wpr, int parity, int rank) { return 2 * BP_MAXW * seam_rec_words(wpr) + (size_t)parity * BP_MAXW + rank; }
__host__ __device__ inline size_t bp_halo_flag(int wpr, int from_below) { return 2 * BP_MAXW * seam_rec_words(wpr) + 2 * BP_MAXW + from_below; }
__host__ __device__ inline size_t bp_mailbox_words(int wpr) { return 2 * BP_MAXW * seam_rec_words(wpr) + 2 * BP_MAXW + 4; }

#ifndef B2C_EMU
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
  const unsigned long long t0 = seam_now();
  for (int d = 0; d < 2; ++d) {
    if ((d == 0 && q.rank == 0) || (d == 1 && q.rank + 1 == q.world)) continue;   // d = 0: rows from above, 1: from below
    const uint32_t *f = q.mail[q.rank] + bp_halo_flag(q.wpr, d);
    while ((int)seam_load_acquire_sys(f) < run * nblocks) {
      if (seam_now() - t0 > 2000000000ull) { q.ctl[2] = 1; return; }
    }
  }
}
#endif
}// namespace b2c
