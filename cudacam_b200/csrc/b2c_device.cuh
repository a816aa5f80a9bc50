// b2c_device.cuh -- shared definitions for the device code of the B200 Canny path.
//
// Data formats (all device-resident, see DESIGN.md "Data layout in HBM"):
//   BGR8 input      : interleaved bytes, byte 0 = B (reference: src/cvp/cannyEdgeD.cu:14-19,66-67)
//   2-bit map       : one u32 per 16 horizontally adjacent pixels; bit i (0..15) = STRONG flag of pixel
//                     16*g+i, bit 16+i = WEAK-only flag.  Equivalent of the reference's d_thresh
//                     {0,128,255} (src/cvp/cannyEdgeD.cu:273-293) at 2 bits per pixel.
//   bit planes S, C : one u32 per 32 pixels (bit i = pixel 32*k+i); S = final edges so far,
//                     C = candidates (weak|strong).  Rows -1 and h of every frame are ghost rows
//                     (zero for a whole image = the reference's zero padding, cannyEdgeD.cu:322-329;
//                     the neighbour band's boundary row in row-band mode).
//   edges           : u8 {0,255} per pixel = the reference's d_hyster after removeCandidates
//                     (src/cvp/cannyEdgeD.cu:379-395).
#pragma once
#include <stdint.h>

#ifdef B2C_EMU
#include "cuda_emu.h"
#define B2C_DYN_SMEM(name) char *name = emu_dyn_smem()
#define B2C_GRID_SYNC() emu::grid_sync()
#else
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#define B2C_DYN_SMEM(name)                               \
  extern __shared__ __align__(128) char b2c_dyn_smem_[]; \
  char *name = b2c_dyn_smem_
#define B2C_GRID_SYNC() cooperative_groups::this_grid().sync()
#endif

#define B2C_FULL 0xffffffffu

// Stage ids == cvp::CannyStage (src/cvp/define.hpp:9-17)
enum { B2C_MONO = 0, B2C_GAUSSIAN = 1, B2C_GRADIENT = 2, B2C_NMS = 3, B2C_THRESH = 4, B2C_HYSTER = 5 };

struct B2cStencilParams {
  const uint8_t *bgr;        // frame 0, band row 0 (rows -4.. may be read in band mode)
  long long row_stride;      // bytes
  long long frame_stride;    // bytes
  int w, h;                  // width, rows produced by this launch
  int y0, h_glob;            // global row of row 0 and global image height (zero padding applies
                             // outside [0,h_glob), real neighbour rows are read inside it)
  int nframes;
  uint32_t *map2;            // 2-bit map out
  int map_pitch;             // u32 per row
  long long map_frame_stride;// u32 per frame
  unsigned lo, hi;           // thresholds (src/cvp/cannyEdgeH.cu:22-23, strict '>' cannyEdgeD.cu:290)
  float gk[25];              // k*(1/159.0f) rounded on the host (src/cvp/cannyEdgeH.cu:372-379)
  // optional per-stage outputs (frame 0 only; null = not written)
  uint8_t *mono, *blur, *nms, *thresh;
  float *grad;
  int pitch8, pitchf;        // in elements
};

struct B2cHystParams {
  const uint32_t *map2;
  int map_pitch;
  long long map_frame_stride;
  uint32_t *S, *C;           // row 0 of frame 0; row -1 / row h are ghost rows
  int plane_pitch;           // u32 per row
  long long plane_frame_stride;
  int w, h, nframes;
  uint8_t *edges;            // may be null (bit-plane consumers read S)
  long long edges_pitch, edges_frame_stride;
  int *flags;                // [0..2] round flags, [3] rounds used (out), [4] any-change-at-all (out)
  int max_rounds;
  int tile_rows;
  int skip_init;             // 1 = S/C planes already built (row-band mode re-entry)
  int skip_expand;           // 1 = do not write edges (row-band mode intermediate rounds)
};

// Exact sector of the gradient direction from the integer Sobel sums -- same rule as
// oracle_sector() (pinned to the reference's atan2f path on the GPU, see tests/).
// Reference: src/cvp/cannyEdgeD.cu:196 (atan2(sX,sY)) and :239-264 (sector boundaries).
__device__ __forceinline__ int b2c_sector(int gx, int gy)
{
  const int a = gx < 0 ? -gx : gx, b = gy < 0 ? -gy : gy;
  const int apb = a + b, amb = a - b, b2 = 2 * b * b;
  if (a == 0 || apb * apb < b2) return 0;
  if (a > b && amb * amb > b2) return 2;
  return ((gx ^ gy) >= 0) ? 1 : 3;
}
