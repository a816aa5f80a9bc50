// emu_main.cpp -- TEST-ONLY: runs the product's kernel sources on the CPU through tests/emu/cuda_emu.h
// (one OS thread per CUDA thread) so that kernel logic can be checked against the oracle without a GPU.
// Never part of libb200canny.so.
#define B2C_EMU 1
#include "cuda_emu.h"

#include "../../cudacam_b200/csrc/b2c_device.cuh"
#include "../../cudacam_b200/csrc/k_hysteresis_uf.cuh"
#include "../../cudacam_b200/csrc/k_band_seam.cuh"
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
static long long g_emu_plane_stride = 0;
static int g_emu_spread = 8;
extern "C" {
__attribute__((visibility("default"))) void emu_set_channels(int ch) { g_emu_channels = ch; }
__attribute__((visibility("default"))) void emu_set_plane_stride(long long s) { g_emu_plane_stride = s; }   // planar BGR8 (tile kernel)
__attribute__((visibility("default"))) void emu_set_spread(int w) { g_emu_spread = w; }   // k_uf_tile: warps the work items are dealt to
// impl: 1 = tile kernel (EMIT when any stage pointer is given), 100 + rb = marching kernel with rb rows per band
__attribute__((visibility("default"))) int emu_stencil(int impl, const uint8_t *bgr, long long row_stride, long long frame_stride, int w, int h, int y0, int h_glob, int nframes,
                                                       unsigned lo, unsigned hi, uint32_t *map2, uint8_t *mono, uint8_t *blur, float *grad, uint8_t *nms, uint8_t *thresh)
{
  B2cStencilParams p;
  memset(&p, 0, sizeof(p));
  alignas(16) static const uint8_t zeros[256] = { 0 };
  p.zeros = zeros;
  p.bgr = bgr; p.row_stride = row_stride; p.frame_stride = frame_stride;
  p.w = w; p.h = h; p.y0 = y0; p.h_glob = h_glob; p.nframes = nframes; p.channels = g_emu_channels; p.plane_stride = g_emu_plane_stride;
  // the kernels write the bit planes S and C; the tests look at the 2-bit map view
  const int gpr = (w + 15) / 16, pitch16 = ((w + 31) / 32 + 3) / 4 * 4 * 2;
  std::vector<uint16_t> pS((size_t)nframes * h * pitch16, 0), pC((size_t)nframes * h * pitch16, 0);
  p.pl_S = pS.data(); p.pl_C = pC.data(); p.pl_pitch16 = pitch16; p.pl_frame_stride16 = (long long)h * pitch16;
  auto to_map2 = [&] {
    for (int f = 0; f < nframes; ++f)
      for (int y = 0; y < h; ++y)
        for (int g = 0; g < gpr; ++g) {
          const uint32_t sv = pS[((size_t)f * h + y) * pitch16 + g], cv = pC[((size_t)f * h + y) * pitch16 + g];
          map2[((size_t)f * h + y) * gpr + g] = sv | ((cv & ~sv) << 16);
        }
  };
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
    to_map2();
    return 0;
  }
#ifdef B2C_EMU_FUSED
  if (impl >= 100) {   // impl = 100 + rows per band
    const int rc = b2c::march_emu_launch(p, impl - 100);
    to_map2();
    return rc;
  }
  return -1;
#else
  return -1;
#endif
}

}   // extern "C"

// ---- hysteresis: planes from the 2-bit map (what the stencil kernels write), then the product's three kernels ------
namespace
{
struct EmuPlanes {
  int w = 0, h = 0, n = 0, wpr = 0, pitch = 0;
  long long fs = 0;
  std::vector<uint32_t> S, C, E;
  std::vector<int> parent;
  std::vector<uint32_t> bl;
  std::vector<int> bc;
  int bcap = 0;
  int flags[16] = { 0 };
  void init(int w_, int h_, int n_)
  {
    w = w_; h = h_; n = n_;
    wpr = (w + 31) / 32;
    pitch = (wpr + 3) / 4 * 4;
    fs = (long long)(h + 2) * pitch;
    S.assign((size_t)fs * n, 0); C.assign((size_t)fs * n, 0); E.assign((size_t)fs * n, 0);
    parent.assign((size_t)n * h * pitch * 32, -12345);
    const int gy = (h + b2c::UT_ROWS - 1) / b2c::UT_ROWS, gx = (wpr + b2c::UT_WORDS - 1) / b2c::UT_WORDS;
    bcap = gy * wpr + 2 * gx * h;
    bl.assign((size_t)n * bcap + 1, 0);
    bc.assign(n, 0);
  }
  void from_map2(const uint32_t *map2)
  {
    const int gpr = (w + 15) / 16;
    std::fill(S.begin(), S.end(), 0u); std::fill(C.begin(), C.end(), 0u);
    for (int f = 0; f < n; ++f)
      for (int y = 0; y < h; ++y)
        for (int g = 0; g < gpr; ++g) {
          const uint32_t m = map2[((size_t)f * h + y) * gpr + g], sv = m & 0xFFFFu, cv = sv | (m >> 16);
          const size_t o = (size_t)f * fs + (size_t)(y + 1) * pitch + (g >> 1);
          S[o] |= sv << (16 * (g & 1));
          C[o] |= cv << (16 * (g & 1));
        }
  }
  B2cHystParams params(uint8_t *edges)
  {
    B2cHystParams p;
    memset(&p, 0, sizeof(p));
    p.S = S.data() + pitch; p.C = C.data() + pitch; p.E = E.data() + pitch; p.plane_pitch = pitch; p.plane_frame_stride = fs;
    p.w = w; p.h = h; p.nframes = n;
    p.edges = edges; p.edges_pitch = w; p.edges_frame_stride = (long long)w * h;
    p.flags = flags; p.spread = g_emu_spread;
    p.parent = parent.data(); p.parent_frame_stride = (long long)h * pitch * 32;
    return p;
  }
  void hysteresis(uint8_t *edges)
  {
    const B2cHystParams p = params(edges);
    const dim3 gt((wpr + b2c::UT_WORDS - 1) / b2c::UT_WORDS, (h + b2c::UT_ROWS - 1) / b2c::UT_ROWS, n);
    uint32_t *blp = bl.data();
    int *bcp = bc.data();
    const int cap = bcap;
    emu::launch(gt, dim3(b2c::UT_THREADS), b2c::UT_SMEM, false, [=] { b2c::k_uf_tile(p, blp, bcp, cap); });
    emu::launch(dim3(2, 1, n), dim3(b2c::UFK_THREADS), 0, false, [=] { b2c::k_uf_border(p, blp, bcp, cap); });
    const int tx = 32, ty = 4;
    const dim3 gr((wpr + tx - 1) / tx, (h + 2 * ty - 1) / (2 * ty), n), br(tx, ty);   // a thread takes 2 rows
    if (edges) emu::launch(gr, br, 0, false, [=] { b2c::k_uf_resolve<true>(p, bcp); });
    else emu::launch(gr, br, 0, false, [=] { b2c::k_uf_resolve<false>(p, bcp); });
  }
};
}// namespace

extern "C" {
// planes from a 2-bit map, then the product's three hysteresis launches; edges may be null (bit plane only)
__attribute__((visibility("default"))) int emu_hysteresis(const uint32_t *map2, int w, int h, int nframes, uint8_t *edges, uint32_t *bits_out)
{
  EmuPlanes P;
  P.init(w, h, nframes);
  P.from_map2(map2);
  P.hysteresis(edges);
  if (bits_out)
    for (int f = 0; f < nframes; ++f)
      for (int y = 0; y < h; ++y) memcpy(bits_out + ((size_t)f * h + y) * P.wpr, P.E.data() + f * P.fs + (long long)(y + 1) * P.pitch, P.wpr * 4);
  return 0;
}

// ---- one row band with retained planes and forest (row-band protocol on the CPU) ---------------------------------------
// Same launch order as b2c_band_hysteresis / b2c_band_seam_solve: tile, border, publish, resolve<LIST> | solve, list pass.
struct EmuBand {
  EmuPlanes P;
  std::vector<uint8_t> edges;
  std::vector<int> roots, hkey, hval, sP;
  std::vector<uint2> ulist;
  std::vector<uint32_t> rec;
  bool force_global = false;
  int ucap = 0;
  int ctl[8] = { 0 };
  int run = 0;
};
constexpr int EMU_SEAM_THREADS = 128;   // (the kernels take the CTA size from blockDim; 1024 OS threads per block would be slow)
static b2c::B2cSeamBand emu_seam_band(EmuBand *b)
{
  b2c::B2cSeamBand s;
  s.S = b->P.S.data() + b->P.pitch; s.C = b->P.C.data() + b->P.pitch;
  s.plane_pitch = b->P.pitch; s.wpr = b->P.wpr; s.h = b->P.h;
  s.parent = b->P.parent.data(); s.roots = b->roots.data(); s.hkey = b->hkey.data(); s.hval = b->hval.data();
  s.hsize = (int)b->hkey.size(); s.ctl = b->ctl;
  s.shash = b->force_global ? 0 : b2c::SEAM_SHASH; s.snodes = b->force_global ? 0 : b2c::SEAM_SNODES;
  return s;
}
// ucap: capacity of the unresolved-word list (<= 0: the product's sizing); a tiny one exercises the overflow path
// force_global: global-memory hash / forest in the seam kernels (the product takes them for very busy seams only)
__attribute__((visibility("default"))) void *emu_band_create(int w, int h, int ucap, int force_global)
{
  EmuBand *b = new EmuBand;
  b->force_global = force_global != 0;
  b->P.init(w, h, 1);
  b->edges.assign((size_t)w * h, 0);
  const int cap = b2c::seam_cap(b->P.wpr), hs = b2c::seam_hash_size(b->P.wpr);
  b->roots.assign(2 * cap, 0);
  b->hkey.assign(hs, 0);
  b->hval.assign(hs, 0);
  b->sP.assign((size_t)b2c::SEAM_MAXW * 2 * cap + 1, 0);
  b->rec.assign(b2c::seam_rec_words(b->P.wpr), 0u);
  b->ucap = ucap > 0 ? ucap : std::max(4096, b->P.wpr * h / 4);
  b->ulist.assign(b->ucap, uint2{ 0u, 0u });
  return b;
}
__attribute__((visibility("default"))) void emu_band_destroy(void *h) { delete static_cast<EmuBand *>(h); }
__attribute__((visibility("default"))) int emu_seam_words(int w) { return (int)b2c::seam_rec_words((w + 31) / 32); }
// band-local hysteresis; the seam record is copied to rec_out
__attribute__((visibility("default"))) void emu_band_hysteresis(void *h, const uint32_t *map2, uint32_t *rec_out)
{
  EmuBand *b = static_cast<EmuBand *>(h);
  EmuPlanes &P = b->P;
  P.from_map2(map2);
  const B2cHystParams p = P.params(b->edges.data());
  const dim3 gt((P.wpr + b2c::UT_WORDS - 1) / b2c::UT_WORDS, (P.h + b2c::UT_ROWS - 1) / b2c::UT_ROWS, 1);
  uint32_t *blp = P.bl.data();
  int *bcp = P.bc.data();
  const int cap = P.bcap;
  emu::launch(gt, dim3(b2c::UT_THREADS), b2c::UT_SMEM, false, [=] { b2c::k_uf_tile(p, blp, bcp, cap); });
  emu::launch(dim3(2, 1, 1), dim3(b2c::UFK_THREADS), 0, false, [=] { b2c::k_uf_border(p, blp, bcp, cap); });
  const b2c::B2cSeamBand s = emu_seam_band(b);
  const int run = ++b->run;
  uint32_t *rec = b->rec.data();
  emu::launch(dim3(1), dim3(EMU_SEAM_THREADS), b2c::seam_publish_smem(P.wpr), false, [=] { b2c::k_seam_publish(s, rec, run); });
  const int tx = 32, ty = 4;
  const dim3 gr((P.wpr + tx - 1) / tx, (P.h + 2 * ty - 1) / (2 * ty), 1), br(tx, ty);
  uint2 *ul = b->ulist.data();
  int *uc = b->ctl + 4;
  const int ucap = b->ucap;
  emu::launch(gr, br, 0, false, [=] { b2c::k_uf_resolve<true, true>(p, bcp, ul, uc, ucap); });
  memcpy(rec_out, rec, b->rec.size() * 4);
}
__attribute__((visibility("default"))) int emu_band_solve(void *h, const uint32_t *all_records, int world, int rank)
{
  EmuBand *b = static_cast<EmuBand *>(h);
  const b2c::B2cSeamBand s = emu_seam_band(b);
  b2c::B2cSeamAll a;
  memset(&a, 0, sizeof(a));
  const size_t stride = b2c::seam_rec_words(b->P.wpr);
  for (int r = 0; r < world; ++r) a.rec[r] = all_records + r * stride;
  a.world = world; a.rank = rank; a.P = b->sP.data();
  emu::launch(dim3(1), dim3(EMU_SEAM_THREADS), b2c::seam_solve_smem(), false, [=] { b2c::k_seam_solve(s, a); });
  const B2cHystParams p = b->P.params(b->edges.data());
  const uint2 *ul = b->ulist.data();
  const int *uc = b->ctl + 4, *need = b->ctl;
  const int ucap = b->ucap;
  emu::launch(dim3(3), dim3(b2c::UFK_THREADS), 0, false, [=] { b2c::k_uf_resolve_list<true>(p, ul, uc, ucap, need); });
  return b->ctl[3];
}
__attribute__((visibility("default"))) int emu_band_unresolved_words(void *h) { return static_cast<EmuBand *>(h)->ctl[4]; }
__attribute__((visibility("default"))) void emu_band_edges(void *h, uint8_t *out)
{
  EmuBand *b = static_cast<EmuBand *>(h);
  memcpy(out, b->edges.data(), b->edges.size());
}
}
