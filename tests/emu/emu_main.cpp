// emu_main.cpp -- TEST-ONLY: runs the product's kernel sources on the CPU through tests/emu/cuda_emu.h
// (one OS thread per CUDA thread) so that kernel logic can be checked against the oracle without a GPU.
// Never part of libb200canny.so.
#define B2C_EMU 1
#include "cuda_emu.h"

#include "../../cudacam_b200/csrc/b2c_device.cuh"
#include "../../cudacam_b200/csrc/k_hysteresis.cuh"
#include "../../cudacam_b200/csrc/k_hysteresis_uf.cuh"
#include "../../cudacam_b200/csrc/k_stencil_tile.cuh"
#ifdef B2C_EMU_FUSED
#include "../../cudacam_b200/csrc/k_stencil_march.cuh"
#endif

static void fill_gk(float gk[25])
{
  static const float k[25] = { 2, 4, 5, 4, 2, 4, 9, 12, 9, 4, 5, 12, 15, 12, 5, 4, 9, 12, 9, 4, 2, 4, 5, 4, 2 };
  const float inv = 1 / 159.0f;
  for (int i = 0; i < 25; ++i) {
    volatile float v = k[i] * inv;
    gk[i] = v;
  }
}

static int g_emu_channels = 3;
extern "C" {
__attribute__((visibility("default"))) void emu_set_channels(int ch) { g_emu_channels = ch; }
// impl: 1 = tile kernel (EMIT when any stage pointer is given), 100 + rb = marching kernel with rb rows per band
__attribute__((visibility("default"))) int emu_stencil(int impl, const uint8_t *bgr, long long row_stride, long long frame_stride, int w, int h, int y0, int h_glob, int nframes,
                                                       unsigned lo, unsigned hi, uint32_t *map2, uint8_t *mono, uint8_t *blur, float *grad, uint8_t *nms, uint8_t *thresh)
{
  B2cStencilParams p;
  memset(&p, 0, sizeof(p));
  alignas(16) static const uint8_t zeros[256] = { 0 };
  p.zeros = zeros;
  p.bgr = bgr; p.row_stride = row_stride; p.frame_stride = frame_stride;
  p.w = w; p.h = h; p.y0 = y0; p.h_glob = h_glob; p.nframes = nframes; p.channels = g_emu_channels;
  p.map2 = map2; p.map_pitch = (w + 15) / 16; p.map_frame_stride = (long long)h * p.map_pitch;
  p.lo = lo; p.hi = hi;
  fill_gk(p.gk);
  b2c_fill_thresholds(p);
  p.mono = mono; p.blur = blur; p.grad = grad; p.nms = nms; p.thresh = thresh;
  p.pitch8 = w; p.pitchf = w;
  if (impl == 1) {
    dim3 grid((w + b2c::TILE_W - 1) / b2c::TILE_W, (h + b2c::TILE_H - 1) / b2c::TILE_H, nframes);
    if (mono || blur || grad || nms || thresh)
      emu::launch(grid, dim3(b2c::TILE_THREADS), b2c::TILE_SMEM, false, [p] { b2c::k_stencil_tile<true>(p); });
    else
      emu::launch(grid, dim3(b2c::TILE_THREADS), b2c::TILE_SMEM, false, [p] { b2c::k_stencil_tile<false>(p); });
    return 0;
  }
#ifdef B2C_EMU_FUSED
  if (impl >= 100) return b2c::march_emu_launch(p, impl - 100);   // impl = 100 + rows per band
  return -1;
#else
  return -1;
#endif
}

// S/C planes are allocated here (with ghost rows); ghost_top/ghost_bot (wpr words each, may be null) seed the
// ghost rows (row-band mode).  Returns rounds used; *changed = flags[4].
__attribute__((visibility("default"))) int emu_hysteresis(const uint32_t *map2, int w, int h, int nframes, int grid_blocks, int tile_rows, uint8_t *edges, uint32_t *bits_out,
                                                          const uint32_t *ghost_top, const uint32_t *ghost_bot, int *changed)
{
  const int wpr = (w + 31) / 32, pitch = (wpr + 3) / 4 * 4;
  const long long fs = (long long)(h + 2) * pitch;
  std::vector<uint32_t> S((size_t)fs * nframes, 0), Cc((size_t)fs * nframes, 0);
  if (ghost_top) memcpy(S.data(), ghost_top, wpr * 4);
  if (ghost_bot) memcpy(S.data() + (size_t)(h + 1) * pitch, ghost_bot, wpr * 4);
  int flags[16] = { 0 };
  B2cHystParams p;
  memset(&p, 0, sizeof(p));
  p.map2 = map2; p.map_pitch = (w + 15) / 16; p.map_frame_stride = (long long)h * p.map_pitch;
  p.S = S.data() + pitch; p.C = Cc.data() + pitch; p.plane_pitch = pitch; p.plane_frame_stride = fs;
  p.w = w; p.h = h; p.nframes = nframes;
  p.edges = edges; p.edges_pitch = w; p.edges_frame_stride = (long long)w * h;
  p.flags = flags; p.max_rounds = 1 << 20; p.tile_rows = tile_rows; p.spread = 1;
  std::vector<int> parent((size_t)nframes * h * pitch * 32, -12345);
  p.parent = parent.data(); p.parent_frame_stride = (long long)h * pitch * 32;
  if (tile_rows > 0)
    emu::launch(dim3(grid_blocks), dim3(b2c::HYST_THREADS), b2c::hyst_smem_bytes(tile_rows), true, [p] { b2c::k_hysteresis(p); });
  else if (tile_rows == 0)   // the cooperative union-find kernel
    emu::launch(dim3(grid_blocks), dim3(b2c::UF_THREADS), b2c::UF_SMEM, true, [p] { b2c::k_hysteresis_uf(p); });
  else {   // tile_rows < 0: union-find as four launches; ghost rows given => re-entry on planes built by a first pass
    const int reps = (ghost_top || ghost_bot) ? 2 : 1;   // second repetition exercises the REENTRY build on the retained planes
    const dim3 gt((wpr + b2c::UT_WORDS - 1) / b2c::UT_WORDS, (h + b2c::UT_ROWS - 1) / b2c::UT_ROWS, nframes);
    const int bcap = (int)gt.y * wpr + 2 * (int)gt.x * h;
    std::vector<uint32_t> bl((size_t)nframes * bcap + 1);
    std::vector<int> bc(nframes, 0);
    uint32_t *blp = bl.data();
    int *bcp = bc.data();
    for (int rep = 0; rep < reps; ++rep) {
      if (rep == 0) {
        // first pass WITHOUT the ghost rows when a re-entry follows (they arrive later in row-band mode)
        std::vector<uint32_t> keep_top, keep_bot;
        if (reps == 2) {
          keep_top.assign(S.begin(), S.begin() + pitch);
          keep_bot.assign(S.begin() + (size_t)(h + 1) * pitch, S.begin() + (size_t)(h + 2) * pitch);
          std::fill(S.begin(), S.begin() + pitch, 0u);
          std::fill(S.begin() + (size_t)(h + 1) * pitch, S.begin() + (size_t)(h + 2) * pitch, 0u);
        }
        emu::launch(gt, dim3(b2c::UT_THREADS), b2c::UT_SMEM, false, [=] { b2c::k_uf_tile(p, blp, bcp, bcap); });
        emu::launch(dim3(2, 1, nframes), dim3(b2c::UFK_THREADS), 0, false, [=] { b2c::k_uf_border(p, blp, bcp, bcap); });
        if (reps == 2) {
          std::copy(keep_top.begin(), keep_top.end(), S.begin());
          std::copy(keep_bot.begin(), keep_bot.end(), S.begin() + (size_t)(h + 1) * pitch);
        }
      } else {
        emu::launch(dim3((wpr + b2c::UFK_THREADS - 1) / b2c::UFK_THREADS, 2, nframes), dim3(b2c::UFK_THREADS), 0, false, [=] { b2c::k_uf_seed(p); });
      }
      const int tx = 32, ty = 4;
      const dim3 gr((wpr + tx - 1) / tx, (h + 2 * ty - 1) / (2 * ty), nframes), br(tx, ty);   // a thread takes 2 rows
      if (rep + 1 < reps) emu::launch(gr, br, 0, false, [=] { b2c::k_uf_resolve<false>(p, bcp); });
      else emu::launch(gr, br, 0, false, [=] { b2c::k_uf_resolve<true>(p, bcp); });
    }
  }
  if (bits_out)
    for (int f = 0; f < nframes; ++f)
      for (int y = 0; y < h; ++y) memcpy(bits_out + ((size_t)f * h + y) * wpr, p.S + f * fs + (long long)y * pitch, wpr * 4);
  if (changed) *changed = flags[4];
  return flags[3];
}
}
