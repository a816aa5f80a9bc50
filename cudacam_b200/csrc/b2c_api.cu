// b2c_api.cu -- the C ABI of libb200canny.so (see include/b200canny.h) and the host driver behind it.
//
// Replaces the host class of the reference, cvp::cuda::CannyEdge (src/cvp/cannyEdgeH.{hpp,cu}):
//   ctor/_initAlloc (:16-38,:340-385)  -> b2c_create          (device buffers, pinned staging, streams)
//   run            (:49-120)           -> b2c_run             (one host frame, blocking, stage select)
//   _loadInputImage(:122-152)          -> async pitched H2D copy on the handle's stream
//   _run* x6 + CPU hysteresis loop (:214-338) -> 1 fused stencil launch (k_stencil_march) + 3 union-find launches
//                                                 (k_uf_tile, k_uf_border, k_uf_resolve), no host round trip
//   _sendOutputToOpenGL (:154-212)     -> the VIEW buffer (tight w*h bytes, what the PBO receives)
//   _start/_endCudaTimer (:409-430)    -> event pairs read back only by b2c_last_timings
//   checkCudaErrors -> exit (helper.hpp:4-17) -> status codes, never exit
// plus what the reference does not have: device-resident batches, a pinned double-buffered host
// pipeline, and row-band mode for one image split over several GPUs.
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/b200canny.h"
#include "b2c_device.cuh"
#include "k_hysteresis_uf.cuh"
#include "k_band_p2p.cuh"
#include "k_band_seam.cuh"
#include "k_stencil_march.cuh"
#include "k_stencil_tile.cuh"
#include "k_views.cuh"

namespace
{
constexpr int NSLOT = 8;        // slots of the host pipeline: the batch buffers (input, planes, forest, edge maps) are cut into up to 8 slots
constexpr int SLOT_FRAMES = 8;  // of at most 8 frames, so that only one small upload and one small download stay exposed

struct Ev {
  cudaEvent_t e = nullptr;
};

}// namespace

struct b2c_ctx {
  int dev = 0, w = 0, h = 0, ch = 3, max_batch = 1;
  int sm_count = 0;
  uint8_t lo = 10, hi = 40;   // src/cvp/cannyEdgeH.cu:22-23
  bool profiling = true;      // src/cvp/cannyEdgeH.cu:24
  int stencil_impl = 0;       // 0 marching two-warp pipeline kernel, 1 staged tile kernel (all-stages path)
  int march_rb = 0;           // rows per band of the marching kernel, 0 = automatic
  int march_ctas_per_sm = 0;  // resident CTAs per SM of the marching kernel (occupancy query at creation)
  int march_extra_smem = 0;   // profiling knob: extra dynamic shared memory per CTA (lowers the occupancy)
  int uf_spread = 0;   // tile kernel: warps the compacted work items are dealt to (1, 2, 4, 8); 0 = by batch size

  // geometry
  int map_pitch = 0;      // u32 per row of the 2-bit map
  int wpr = 0;            // u32 per row of a bit plane (used words)
  int plane_pitch = 0;    // u32 per row of a bit plane (allocated)
  int rows_alloc = 0;     // rows per frame of map / planes / edges (h, or band rows)
  size_t in_row_stride = 0, in_frame_stride = 0;   // own input buffer
  bool planar = false;      // B2C_PLANAR_BGR8: three planes (B, G, R) of `height` rows each instead of interleaved pixels
  size_t row_bytes() const { return planar ? (size_t)w : (size_t)w * ch; }   // payload bytes of one input row
  int rows_in() const { return planar ? 3 * h : h; }                         // input rows per frame
  size_t edges_pitch = 0, edges_frame_stride = 0;
  int pitch8 = 0, pitchf = 0;   // stage buffers, elements

  // device memory
  uint8_t *d_in = nullptr;
  uint32_t *d_map2 = nullptr;   // 2-bit map view of the planes (accessor format, made on demand)
  uint32_t *d_S_base = nullptr, *d_C_base = nullptr, *d_E_base = nullptr;   // bit planes incl. one zero ghost row above and below each frame
  uint8_t *d_edges = nullptr;
  uint8_t *d_mono = nullptr, *d_blur = nullptr, *d_nms = nullptr, *d_thresh = nullptr, *d_view = nullptr;
  float *d_grad = nullptr;
  int *d_flags = nullptr;
  int *d_parent = nullptr;
  uint32_t *d_blist = nullptr;   // per-frame lists of tile-border words with weak pixels
  int *d_bcount = nullptr;       // their lengths (zeroed again by the resolve kernel)
  int bcap = 0;                  // entries per frame
  uint8_t *d_zeros = nullptr;
  int *h_flags = nullptr;   // pinned mirror

  // last input (for on-demand stage buffers)
  const uint8_t *last_in = nullptr;
  int acc_frame = 0;   // batch slot (first frame) the accessors of "frame 0 of the last run" read
  size_t last_row_stride = 0;
  bool have_frame = false, stages_valid = false, edges_valid = false, map2_valid = false;
  int last_stage = -1;

  // host pipeline
  uint8_t *h_in[NSLOT] = {};
  uint8_t *h_out[NSLOT] = {};
  size_t h_in_bytes = 0, h_out_bytes = 0;
  cudaStream_t s_main = nullptr, s_h2d = nullptr, s_d2h = nullptr, s_side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_in[NSLOT] = {}, ev_k[NSLOT] = {}, ev_out[NSLOT] = {};
  cudaEvent_t ev_h[4] = {};   // hysteresis phase marks (only with the hyst_phase_timing option)
  bool hyst_phase_timing = false;
  cudaEvent_t ev_t[5] = {};   // timing marks: start, after upload, after stencil, after hysteresis, end
  bool timings_valid = false;

  // band mode
  bool band = false;
  int band_y0 = 0, h_glob = 0;
  // band mode: seam records (k_band_seam.cuh); peer-to-peer: own mailbox, the other ranks' mailboxes mapped through CUDA IPC
  uint32_t *d_seam_rec = nullptr;          // own record (NCCL / gloo all-gather path)
  int *d_seam_roots = nullptr, *d_seam_hkey = nullptr, *d_seam_hval = nullptr, *d_seam_P = nullptr;
  int *d_seam_ctl = nullptr, *h_seam_ctl = nullptr;   // [0] promoted flag, [2] peer time-out, [3] runs promoted, [4] length of ulist
  uint2 *d_ulist = nullptr;                // plane words that keep unresolved weak pixels after the band-local resolve
  int ucap = 0;
  cudaEvent_t ev_b[4] = {};                // phase marks of b2c_band_p2p_stencil: start, after the push, before the stencil, end
  cudaEvent_t ev_s[7] = {};                // phase marks of the band hysteresis / seam pass (hyst_phase_timing option)
  int seam_run = 0;
  bool seam_force_global = false;          // test option: global-memory hash / forest in the seam kernels
  uint32_t *d_mailbox = nullptr;
  void *peer_mail[b2c::BP_MAXW] = {};
  void *peer_in[2] = { nullptr, nullptr };   // band input buffers of the upper / lower neighbour
  bool peers_ipc = false;                  // peer pointers came from cudaIpcOpenMemHandle (to be closed)
  int peer_rows_up = 0;
  uint8_t *d_band_in = nullptr;            // own band input buffer: 4 halo rows, the band, 4 halo rows
  int p2p_world = 0, p2p_rank = 0, p2p_run = 0;

  long long launches = 0;
  std::string last_err;
};

namespace
{
int set_err(b2c_ctx *c, cudaError_t e, const char *what)
{
  if (c) {
    c->last_err = what;
    c->last_err += ": ";
    c->last_err += cudaGetErrorString(e);
  }
  (void)cudaGetLastError();
  return B2C_ERR_CUDA;
}
#define CK(c, x)                                      \
  do {                                                \
    cudaError_t e_ = (x);                             \
    if (e_ != cudaSuccess) return set_err(c, e_, #x); \
  } while (0)

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void fill_gk(float gk[25])
{
  // src/cvp/cannyEdgeH.cu:372-379: k * (1/159.0f), single fp32 rounding of the product, on the host
  static const float k[25] = { 2, 4, 5, 4, 2, 4, 9, 12, 9, 4, 5, 12, 15, 12, 5, 4, 9, 12, 9, 4, 2, 4, 5, 4, 2 };
  const float inv = 1 / 159.0f;
  for (int i = 0; i < 25; ++i) {
    volatile float v = k[i] * inv;
    gk[i] = v;
  }
}

uint32_t *S0(b2c_ctx *c) { return c->d_S_base + c->plane_pitch; }
uint32_t *C0(b2c_ctx *c) { return c->d_C_base + c->plane_pitch; }
uint32_t *E0(b2c_ctx *c) { return c->d_E_base + c->plane_pitch; }
long long plane_frame_stride(const b2c_ctx *c) { return (long long)(c->rows_alloc + 2) * c->plane_pitch; }

int alloc_common(b2c_ctx *c)
{
  const int w = c->w, rows = c->rows_alloc, nb = c->max_batch;
  c->map_pitch = (w + 15) / 16;
  c->wpr = (w + 31) / 32;
  c->plane_pitch = (int)round_up((size_t)c->wpr, 4);
  c->edges_pitch = (size_t)w;   // tight, like the reference's PBO (src/imgui/imguiApp.cpp:76)
  c->edges_frame_stride = round_up(c->edges_pitch * rows, 256);
  c->pitch8 = (int)round_up((size_t)w, 16);
  c->pitchf = (int)round_up((size_t)w, 4);
  cudaDeviceProp prop;
  CK(c, cudaGetDeviceProperties(&prop, c->dev));
  c->sm_count = prop.multiProcessorCount;
  if (c->wpr >= (1 << b2c::UF_XW_BITS) || rows >= (1 << (32 - b2c::UF_XW_BITS)) || (long long)rows * c->plane_pitch * 32 >= (1ll << 31)) {
    c->last_err = "image or band too large for 32-bit union-find node ids / border-list entries";   // (split it into more row bands)
    return B2C_ERR_UNSUPPORTED;
  }
  const size_t plane_bytes = (size_t)nb * plane_frame_stride(c) * 4;
  CK(c, cudaMalloc(&c->d_S_base, plane_bytes));
  CK(c, cudaMalloc(&c->d_C_base, plane_bytes));
  CK(c, cudaMalloc(&c->d_E_base, plane_bytes));
  CK(c, cudaMemset(c->d_S_base, 0, plane_bytes));
  CK(c, cudaMemset(c->d_C_base, 0, plane_bytes));
  CK(c, cudaMemset(c->d_E_base, 0, plane_bytes));
  CK(c, cudaMalloc(&c->d_edges, (size_t)nb * c->edges_frame_stride));
  CK(c, cudaMalloc(&c->d_parent, (size_t)nb * rows * c->plane_pitch * 32 * sizeof(int)));
  {
    const int ntr = (rows + b2c::UT_ROWS - 1) / b2c::UT_ROWS, ntc = (c->wpr + b2c::UT_WORDS - 1) / b2c::UT_WORDS;
    c->bcap = ntr * c->wpr + 2 * ntc * rows;   // every word of the top rows + both border columns of every tile
    CK(c, cudaMalloc(&c->d_blist, (size_t)nb * c->bcap * sizeof(uint32_t)));
    CK(c, cudaMalloc(&c->d_bcount, (size_t)nb * sizeof(int)));
    CK(c, cudaMemset(c->d_bcount, 0, (size_t)nb * sizeof(int)));
  }
  CK(c, cudaMalloc(&c->d_zeros, 256));
  CK(c, cudaMemset(c->d_zeros, 0, 256));
  CK(c, cudaMalloc(&c->d_flags, 16 * sizeof(int)));
  CK(c, cudaMemset(c->d_flags, 0, 16 * sizeof(int)));
  CK(c, cudaMallocHost(&c->h_flags, 16 * sizeof(int)));
  memset(c->h_flags, 0, 16 * sizeof(int));
  CK(c, cudaStreamCreateWithFlags(&c->s_main, cudaStreamNonBlocking));
  CK(c, cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
  CK(c, cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  if (c->band) {   // side stream of the peer-to-peer pushes
    CK(c, cudaStreamCreateWithFlags(&c->s_side, cudaStreamNonBlocking));
    CK(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  }
  for (int i = 0; i < NSLOT; ++i) {
    CK(c, cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
    CK(c, cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
    CK(c, cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
  }
  for (auto &e : c->ev_t) CK(c, cudaEventCreate(&e));

  CK(c, cudaFuncSetAttribute(b2c::k_uf_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, b2c::UT_SMEM));
  CK(c, cudaFuncSetAttribute(b2c::k_stencil_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b2c::TILE_SMEM));
  CK(c, cudaFuncSetAttribute(b2c::k_stencil_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, b2c::TILE_SMEM));
  if (b2c::march_configure(&c->march_ctas_per_sm) != cudaSuccess) return set_err(c, cudaGetLastError(), "march_configure");
  return B2C_OK;
}

// frame0: first frame slot of the handle's batch buffers (planes, forest) this launch works in
void fill_stencil_params(b2c_ctx *c, B2cStencilParams &p, const uint8_t *bgr, size_t row_stride, size_t frame_stride, int n, int frame0 = 0)
{
  memset(&p, 0, sizeof(p));
  p.bgr = bgr;
  p.zeros = c->d_zeros;
  p.row_stride = (long long)row_stride;
  p.frame_stride = (long long)frame_stride;
  p.w = c->w;
  p.h = c->rows_alloc;
  p.y0 = c->band ? c->band_y0 : 0;
  p.h_glob = c->band ? c->h_glob : c->h;
  p.nframes = n;
  p.channels = c->ch;
  p.plane_stride = c->planar ? (long long)row_stride * c->rows_alloc : 0;
  p.pl_S = reinterpret_cast<uint16_t *>(S0(c) + (long long)frame0 * plane_frame_stride(c));
  p.pl_C = reinterpret_cast<uint16_t *>(C0(c) + (long long)frame0 * plane_frame_stride(c));
  p.pl_pitch16 = c->plane_pitch * 2;
  p.pl_frame_stride16 = plane_frame_stride(c) * 2;
  p.lo = c->lo;
  p.hi = c->hi;
  fill_gk(p.gk);
  b2c_fill_thresholds(p);
  p.pitch8 = c->pitch8;
  p.pitchf = c->pitchf;
}

// Fused stencil: BGR8 -> 2-bit map for n frames.
int launch_stencil(b2c_ctx *c, const uint8_t *bgr, size_t row_stride, size_t frame_stride, int n, cudaStream_t st, int frame0 = 0)
{
  B2cStencilParams p;
  fill_stencil_params(c, p, bgr, row_stride, frame_stride, n, frame0);
  c->acc_frame = frame0;
  if (c->stencil_impl == 0 && b2c::march_supported(p)) {
    cudaError_t e = b2c::march_launch(p, c->sm_count, c->march_ctas_per_sm, c->march_rb, st, c->march_extra_smem);
    if (e != cudaSuccess) return set_err(c, e, "k_stencil_march launch");
  } else {
    dim3 grid((c->w + b2c::TILE_W - 1) / b2c::TILE_W, (c->rows_alloc + b2c::TILE_H - 1) / b2c::TILE_H, n);
    b2c::k_stencil_tile<false><<<grid, b2c::TILE_THREADS, b2c::TILE_SMEM, st>>>(p);
    CK(c, cudaGetLastError());
  }
  c->launches++;
  c->map2_valid = false;
  c->edges_valid = false;
  return B2C_OK;
}

// All-stages variant for frame 0 (mono, blur, grad, nms, thresh buffers of the reference).
int launch_stencil_emit(b2c_ctx *c, const uint8_t *bgr, size_t row_stride, cudaStream_t st)
{
  if (!c->d_mono) {
    const size_t n8 = (size_t)c->pitch8 * c->rows_alloc;
    CK(c, cudaMalloc(&c->d_mono, n8));
    CK(c, cudaMalloc(&c->d_blur, n8));
    CK(c, cudaMalloc(&c->d_nms, n8));
    CK(c, cudaMalloc(&c->d_thresh, n8));
    CK(c, cudaMalloc(&c->d_grad, (size_t)c->pitchf * c->rows_alloc * 4));
    CK(c, cudaMalloc(&c->d_view, (size_t)c->w * c->rows_alloc));
  }
  B2cStencilParams p;
  fill_stencil_params(c, p, bgr, row_stride, 0, 1);
  p.mono = c->d_mono;
  p.blur = c->d_blur;
  p.nms = c->d_nms;
  p.thresh = c->d_thresh;
  p.grad = c->d_grad;
  dim3 grid((c->w + b2c::TILE_W - 1) / b2c::TILE_W, (c->rows_alloc + b2c::TILE_H - 1) / b2c::TILE_H, 1);
  b2c::k_stencil_tile<true><<<grid, b2c::TILE_THREADS, b2c::TILE_SMEM, st>>>(p);
  CK(c, cudaGetLastError());
  c->launches++;
  c->stages_valid = true;
  return B2C_OK;
}

void fill_hyst_params(b2c_ctx *c, B2cHystParams &p, int n, uint8_t *edges, size_t edges_pitch, size_t edges_frame_stride, int frame0 = 0)
{
  memset(&p, 0, sizeof(p));
  const long long po = (long long)frame0 * plane_frame_stride(c);
  p.S = S0(c) + po;
  p.C = C0(c) + po;
  p.E = E0(c) + po;
  p.plane_pitch = c->plane_pitch;
  p.plane_frame_stride = plane_frame_stride(c);
  p.w = c->w;
  p.h = c->rows_alloc;
  p.nframes = n;
  p.edges = edges;
  p.edges_pitch = (long long)edges_pitch;
  p.edges_frame_stride = (long long)edges_frame_stride;
  p.flags = c->d_flags;
  p.parent_frame_stride = (long long)c->rows_alloc * c->plane_pitch * 32;
  p.parent = c->d_parent + frame0 * p.parent_frame_stride;
  // measured (tools/hyst_phases.py): one frame 17 us with 8 warps vs 19 with 4 (latency), 64 frames 93 us with 4 vs 97 with 8 (issue slots)
  p.spread = c->uf_spread ? c->uf_spread : (n >= 8 ? 4 : 8);
}

// Union-find hysteresis of n frames from the planes the stencil wrote: three ordinary launches (tile, border,
// resolve + expansion to the u8 map), no barrier inside, no host round trip between.  edges == null: bit plane E only.
int launch_hysteresis(b2c_ctx *c, int n, uint8_t *edges, size_t edges_pitch, size_t edges_frame_stride, cudaStream_t st, int frame0 = 0)
{
  B2cHystParams p;
  fill_hyst_params(c, p, n, edges, edges_pitch, edges_frame_stride, frame0);
  uint32_t *blist = c->d_blist + (size_t)frame0 * c->bcap;
  int *bcount = c->d_bcount + frame0;
  const dim3 gt((c->wpr + b2c::UT_WORDS - 1) / b2c::UT_WORDS, (c->rows_alloc + b2c::UT_ROWS - 1) / b2c::UT_ROWS, n);
  const bool pt = c->hyst_phase_timing;
  const int T = b2c::UFK_THREADS;
  if (pt) cudaEventRecord(c->ev_h[0], st);
  b2c::k_uf_tile<<<gt, b2c::UT_THREADS, b2c::UT_SMEM, st>>>(p, blist, bcount, c->bcap);
  if (pt) cudaEventRecord(c->ev_h[1], st);
  // border list: typically ~12 % of bcap entries; 4 threads per entry, a quarter of the worst case in blocks, grid-stride for the rest
  b2c::k_uf_border<<<dim3(std::max(1, (c->bcap + T - 1) / T), 1, n), T, 0, st>>>(p, blist, bcount, c->bcap);
  if (pt) cudaEventRecord(c->ev_h[2], st);
  const int tx = c->wpr >= 256 ? 256 : c->wpr > 32 ? 64 : 32, ty = 256 / tx;   // 256 threads = tx words x ty rows
  const dim3 gr((c->wpr + tx - 1) / tx, (c->rows_alloc + 2 * ty - 1) / (2 * ty), n), br(tx, ty);   // a thread takes 2 rows
  if (edges) b2c::k_uf_resolve<true><<<gr, br, 0, st>>>(p, bcount);
  else b2c::k_uf_resolve<false><<<gr, br, 0, st>>>(p, bcount);
  if (pt) cudaEventRecord(c->ev_h[3], st);
  CK(c, cudaGetLastError());
  c->launches += 3;
  c->edges_valid = true;
  return B2C_OK;
}

int check_handle(b2c_ctx *c) { return c ? B2C_OK : B2C_ERR_INVALID; }

struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev)
  {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DevGuard()
  {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
}// namespace

extern "C" {

int b2c_create(b2c_handle *out, int device, int width, int height, int channels, int max_batch)
{
  if (!out || width < 1 || height < 1 || max_batch < 1) return B2C_ERR_INVALID;
  *out = nullptr;
  const bool planar = channels == B2C_PLANAR_BGR8;
  if (planar) channels = 3;
  if (channels != 3 && channels != 1 && channels != 4) return B2C_ERR_UNSUPPORTED;   // BGR8 (interleaved or planar), GRAY8, BGRA8
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    (void)cudaGetLastError();
    return B2C_ERR_CUDA;
  }
  b2c_ctx *c = new (std::nothrow) b2c_ctx;
  if (!c) return B2C_ERR_NOMEM;
  c->dev = device;
  c->w = width;
  c->h = height;
  c->ch = channels;
  c->max_batch = max_batch;
  c->rows_alloc = height;
  DevGuard g(device);
  c->planar = planar;
  c->in_row_stride = planar ? round_up((size_t)width, 16) : round_up(round_up((size_t)width, 8) * channels, 16);   // whole 8-pixel lanes are backed by memory
  c->in_frame_stride = c->in_row_stride * c->rows_in();
  int rc = alloc_common(c);
  if (rc == B2C_OK && cudaMalloc(&c->d_in, (size_t)max_batch * c->in_frame_stride) != cudaSuccess) rc = set_err(c, cudaGetLastError(), "cudaMalloc(d_in)");
  if (rc != B2C_OK) {
    fprintf(stderr, "[b200canny] b2c_create failed: %s\n", c->last_err.c_str());
    b2c_destroy(c);
    return rc;
  }
  *out = c;
  return B2C_OK;
}

namespace
{
int seam_alloc(b2c_ctx *c);
}
int b2c_create_band(b2c_handle *out, int device, int width, int band_rows, int y0, int height_global)
{
  if (!out || width < 1 || band_rows < 1 || y0 < 0 || y0 + band_rows > height_global) return B2C_ERR_INVALID;
  // a band with a neighbour sends its first / last 4 rows as the neighbour's input halo: it must have them
  if (band_rows < 4 && !(y0 == 0 && band_rows == height_global)) return B2C_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    (void)cudaGetLastError();
    return B2C_ERR_CUDA;
  }
  b2c_ctx *c = new (std::nothrow) b2c_ctx;
  if (!c) return B2C_ERR_NOMEM;
  c->dev = device;
  c->w = width;
  c->h = band_rows;
  c->max_batch = 1;
  c->rows_alloc = band_rows;
  c->band = true;
  c->band_y0 = y0;
  c->h_glob = height_global;
  DevGuard g(device);
  int rc = alloc_common(c);
  if (rc == B2C_OK) rc = seam_alloc(c);
  if (rc != B2C_OK) {
    fprintf(stderr, "[b200canny] b2c_create_band failed: %s\n", c->last_err.c_str());
    b2c_destroy(c);
    return rc;
  }
  *out = c;
  return B2C_OK;
}

void b2c_destroy(b2c_handle c)
{
  if (!c) return;
  DevGuard g(c->dev);
  if (c->s_main) cudaStreamSynchronize(c->s_main);
  if (c->s_h2d) cudaStreamSynchronize(c->s_h2d);
  if (c->s_d2h) cudaStreamSynchronize(c->s_d2h);
  if (c->s_side) cudaStreamSynchronize(c->s_side);
  cudaFree(c->d_in);
  cudaFree(c->d_map2);
  cudaFree(c->d_S_base);
  cudaFree(c->d_C_base);
  cudaFree(c->d_E_base);
  cudaFree(c->d_edges);
  cudaFree(c->d_mono);
  cudaFree(c->d_blur);
  cudaFree(c->d_nms);
  cudaFree(c->d_thresh);
  cudaFree(c->d_view);
  cudaFree(c->d_grad);
  cudaFree(c->d_flags);
  cudaFree(c->d_parent);
  if (c->peers_ipc) {
    for (int k = 0; k < c->p2p_world; ++k)
      if (k != c->p2p_rank && c->peer_mail[k]) cudaIpcCloseMemHandle(c->peer_mail[k]);
    for (auto &q : c->peer_in)
      if (q) cudaIpcCloseMemHandle(q);
  }
  cudaFree(c->d_band_in);
  cudaFree(c->d_mailbox);
  cudaFree(c->d_seam_rec);
  cudaFree(c->d_seam_roots);
  cudaFree(c->d_seam_hkey);
  cudaFree(c->d_seam_hval);
  cudaFree(c->d_seam_P);
  cudaFree(c->d_seam_ctl);
  cudaFree(c->d_ulist);
  for (auto &e : c->ev_s)
    if (e) cudaEventDestroy(e);
  for (auto &e : c->ev_b)
    if (e) cudaEventDestroy(e);
  if (c->h_seam_ctl) cudaFreeHost(c->h_seam_ctl);
  cudaFree(c->d_blist);
  cudaFree(c->d_bcount);
  cudaFree(c->d_zeros);
  if (c->h_flags) cudaFreeHost(c->h_flags);
  for (int i = 0; i < NSLOT; ++i) {
    if (c->h_in[i]) cudaFreeHost(c->h_in[i]);
    if (c->h_out[i]) cudaFreeHost(c->h_out[i]);
    if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
    if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
  }
  for (auto &e : c->ev_t)
    if (e) cudaEventDestroy(e);
  for (auto &e : c->ev_h)
    if (e) cudaEventDestroy(e);
  if (c->s_main) cudaStreamDestroy(c->s_main);
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  if (c->s_side) cudaStreamDestroy(c->s_side);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  (void)cudaGetLastError();
  delete c;
}

// src/cvp/cannyEdgeH.hpp:25-29
int b2c_set_low_threshold(b2c_handle c, uint8_t low)
{
  if (!c) return B2C_ERR_INVALID;
  c->lo = std::min(low, c->hi);
  c->stages_valid = false;
  return B2C_OK;
}
int b2c_set_high_threshold(b2c_handle c, uint8_t high)
{
  if (!c) return B2C_ERR_INVALID;
  c->hi = std::max(high, c->lo);
  c->stages_valid = false;
  return B2C_OK;
}
int b2c_get_low_threshold(b2c_handle c) { return c ? c->lo : B2C_ERR_INVALID; }
int b2c_get_high_threshold(b2c_handle c) { return c ? c->hi : B2C_ERR_INVALID; }

int b2c_enable_profiling(b2c_handle c, int on)
{
  if (!c) return B2C_ERR_INVALID;
  c->profiling = on != 0;
  return B2C_OK;
}
int b2c_is_profiling_enabled(b2c_handle c) { return c ? (c->profiling ? 1 : 0) : B2C_ERR_INVALID; }

int b2c_last_timings(b2c_handle c, float *ms, int n)
{
  if (!c || !ms || n < 1) return B2C_ERR_INVALID;
  if (!c->timings_valid) return B2C_ERR_STATE;
  DevGuard g(c->dev);
  float v[6] = { 0, 0, 0, 0, 0, 0 };
  CK(c, cudaEventSynchronize(c->ev_t[4]));
  for (int i = 0; i < 4; ++i) CK(c, cudaEventElapsedTime(&v[i], c->ev_t[i], c->ev_t[i + 1]));
  CK(c, cudaEventElapsedTime(&v[4], c->ev_t[0], c->ev_t[4]));
  v[5] = (float)c->h_flags[3];
  for (int i = 0; i < n && i < 6; ++i) ms[i] = v[i];
  return B2C_OK;
}

int b2c_run(b2c_handle c, const uint8_t *host_bgr, size_t row_stride, int final_stage)
{
  if (!c || !host_bgr || c->band) return B2C_ERR_INVALID;
  if (final_stage < B2C_STAGE_MONO || final_stage > B2C_STAGE_HYSTER) return B2C_ERR_INVALID;
  if (row_stride < c->row_bytes()) return B2C_ERR_SIZE;
  DevGuard g(c->dev);
  cudaStream_t st = c->s_main;
  const bool prof = c->profiling;
  if (prof) CK(c, cudaEventRecord(c->ev_t[0], st));
  CK(c, cudaMemcpy2DAsync(c->d_in, c->in_row_stride, host_bgr, row_stride, c->row_bytes(), c->rows_in(), cudaMemcpyHostToDevice, st));
  if (prof) CK(c, cudaEventRecord(c->ev_t[1], st));
  c->last_in = c->d_in;
  c->last_row_stride = c->in_row_stride;
  c->have_frame = true;
  c->stages_valid = false;
  c->last_stage = final_stage;
  int rc;
  if (final_stage == B2C_STAGE_HYSTER) {
    if ((rc = launch_stencil(c, c->d_in, c->in_row_stride, c->in_frame_stride, 1, st)) != B2C_OK) return rc;
    if (prof) CK(c, cudaEventRecord(c->ev_t[2], st));
    if ((rc = launch_hysteresis(c, 1, c->d_edges, c->edges_pitch, c->edges_frame_stride, st)) != B2C_OK) return rc;
    if (prof) CK(c, cudaEventRecord(c->ev_t[3], st));
  } else {
    // stage views: the pipeline stops after the selected stage (src/cvp/cannyEdgeH.cu:58-115)
    if ((rc = launch_stencil_emit(c, c->d_in, c->in_row_stride, st)) != B2C_OK) return rc;
    c->edges_valid = false;   // the pipeline stops after the selected stage: no edge map of this frame exists
    c->map2_valid = false;
    if (prof) CK(c, cudaEventRecord(c->ev_t[2], st));
    if (prof) CK(c, cudaEventRecord(c->ev_t[3], st));
    const int blocks = c->sm_count * 4;
    if (final_stage == B2C_STAGE_GRADIENT) {
      b2c::k_grad_view<<<blocks, 256, 0, st>>>(c->d_grad, c->pitchf, c->d_view, c->w, c->w, c->h);
    } else {
      const uint8_t *src = final_stage == B2C_STAGE_MONO ? c->d_mono : final_stage == B2C_STAGE_GAUSSIAN ? c->d_blur : final_stage == B2C_STAGE_NMS ? c->d_nms : c->d_thresh;
      b2c::k_copy2d_u8<<<blocks, 256, 0, st>>>(src, c->pitch8, c->d_view, c->w, c->w, c->h);
    }
    CK(c, cudaGetLastError());
    c->launches++;
  }
  CK(c, cudaMemcpyAsync(c->h_flags, c->d_flags, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (prof) CK(c, cudaEventRecord(c->ev_t[4], st));
  CK(c, cudaStreamSynchronize(st));   // blocking, like the reference's run()
  c->timings_valid = prof;
  return B2C_OK;
}

int b2c_run_device(b2c_handle c, const uint8_t *dev_bgr, size_t row_stride, size_t frame_stride, int n, uint8_t *dev_edges, size_t edges_pitch, size_t edges_frame_stride, void *stream)
{
  if (!c || !dev_bgr || c->band || n < 1 || n > c->max_batch) return B2C_ERR_INVALID;
  if (row_stride < c->row_bytes()) return B2C_ERR_SIZE;
  DevGuard g(c->dev);
  cudaStream_t st = stream ? (cudaStream_t)stream : c->s_main;
  if (!dev_edges) {
    dev_edges = c->d_edges;
    edges_pitch = c->edges_pitch;
    edges_frame_stride = c->edges_frame_stride;
  } else if (edges_pitch < (size_t)c->w) {
    return B2C_ERR_SIZE;
  }
  const bool prof = c->profiling;
  if (prof) {
    CK(c, cudaEventRecord(c->ev_t[0], st));
    CK(c, cudaEventRecord(c->ev_t[1], st));
  }
  int rc;
  if ((rc = launch_stencil(c, dev_bgr, row_stride, frame_stride, n, st)) != B2C_OK) return rc;
  if (prof) CK(c, cudaEventRecord(c->ev_t[2], st));
  if ((rc = launch_hysteresis(c, n, dev_edges, edges_pitch, edges_frame_stride, st)) != B2C_OK) return rc;
  if (prof) {
    CK(c, cudaEventRecord(c->ev_t[3], st));
    CK(c, cudaEventRecord(c->ev_t[4], st));
  }
  c->last_in = nullptr;   // caller-owned memory: not retained (the stage accessors need a frame run through b2c_run)
  c->last_row_stride = row_stride;
  c->have_frame = true;
  c->stages_valid = false;
  c->last_stage = B2C_STAGE_HYSTER;
  c->timings_valid = prof;
  return B2C_OK;
}

int b2c_stencil_device(b2c_handle c, const uint8_t *dev_bgr, size_t row_stride, size_t frame_stride, int n, void *stream)
{
  if (!c || !dev_bgr || c->band || n < 1 || n > c->max_batch) return B2C_ERR_INVALID;
  if (row_stride < c->row_bytes()) return B2C_ERR_SIZE;
  DevGuard g(c->dev);
  return launch_stencil(c, dev_bgr, row_stride, frame_stride, n, stream ? (cudaStream_t)stream : c->s_main);
}

int b2c_hysteresis_device(b2c_handle c, int n, uint8_t *dev_edges, size_t edges_pitch, size_t edges_frame_stride, void *stream)
{
  if (!c || c->band || n < 1 || n > c->max_batch) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  if (!dev_edges) {
    dev_edges = c->d_edges;
    edges_pitch = c->edges_pitch;
    edges_frame_stride = c->edges_frame_stride;
  }
  return launch_hysteresis(c, n, dev_edges, edges_pitch, edges_frame_stride, stream ? (cudaStream_t)stream : c->s_main);
}

int b2c_run_batch_host(b2c_handle c, const uint8_t *frames, size_t row_stride, int n, uint8_t *edges_out, int packed_bits)
{
  if (!c || !frames || !edges_out || c->band || n < 1) return B2C_ERR_INVALID;
  if (row_stride < c->row_bytes()) return B2C_ERR_SIZE;
  DevGuard g(c->dev);
  const int w = c->w, h = c->h;
  // The batch buffers of the handle (input, planes, forest, edge maps: max_batch frames each) are cut into up to NSLOT
  // slots of at most SLOT_FRAMES frames.  Chunk k lives in slot k % nslots: upload on s_h2d, kernels on s_main,
  // download on s_d2h.  With 8-frame chunks only the first upload and the last download are not hidden.
  const int slot_frames = std::max(1, std::min(SLOT_FRAMES, (c->max_batch + 1) / 2));
  const int nslots = std::max(1, std::min(NSLOT, c->max_batch / slot_frames));
  const size_t in_frame_host = row_stride * c->rows_in();
  const size_t out_row = packed_bits ? (size_t)c->wpr * 4 : (size_t)w;
  const size_t out_frame_host = out_row * h;

  // Is the caller's memory pinned?  If not, stage through our own pinned ring.
  cudaPointerAttributes ai, ao;
  const bool in_pinned = cudaPointerGetAttributes(&ai, frames) == cudaSuccess && ai.type == cudaMemoryTypeHost;
  const bool out_pinned = cudaPointerGetAttributes(&ao, edges_out) == cudaSuccess && ao.type == cudaMemoryTypeHost;
  (void)cudaGetLastError();
  if (!in_pinned && c->h_in_bytes < (size_t)slot_frames * in_frame_host) {
    for (int s = 0; s < NSLOT; ++s) {
      if (c->h_in[s]) cudaFreeHost(c->h_in[s]);
      c->h_in[s] = nullptr;
    }
    for (int s = 0; s < nslots; ++s) CK(c, cudaMallocHost(&c->h_in[s], (size_t)slot_frames * in_frame_host));
    c->h_in_bytes = (size_t)slot_frames * in_frame_host;
  }
  if (!out_pinned && c->h_out_bytes < (size_t)slot_frames * out_frame_host) {
    for (int s = 0; s < NSLOT; ++s) {
      if (c->h_out[s]) cudaFreeHost(c->h_out[s]);
      c->h_out[s] = nullptr;
    }
    for (int s = 0; s < nslots; ++s) CK(c, cudaMallocHost(&c->h_out[s], (size_t)slot_frames * out_frame_host));
    c->h_out_bytes = (size_t)slot_frames * out_frame_host;
  }

  const int nchunks = (n + slot_frames - 1) / slot_frames;
  int pending_out[NSLOT];   // chunk whose D2H into the slot's host buffer still has to be waited for / copied out
  for (int &v : pending_out) v = -1;
  auto drain = [&](int slot) -> int {
    const int k = pending_out[slot];
    if (k < 0) return B2C_OK;
    CK(c, cudaEventSynchronize(c->ev_out[slot]));
    if (!out_pinned) {
      const int f0 = k * slot_frames, cnt = std::min(slot_frames, n - f0);
      memcpy(edges_out + (size_t)f0 * out_frame_host, c->h_out[slot], (size_t)cnt * out_frame_host);
    }
    pending_out[slot] = -1;
    return B2C_OK;
  };

  for (int k = 0; k < nchunks; ++k) {
    const int slot = k % nslots, fr0 = slot * slot_frames;
    const int f0 = k * slot_frames, cnt = std::min(slot_frames, n - f0);
    // the slot's previous chunk is completely through (its download has finished): every buffer of the slot is free
    int rc = drain(slot);
    if (rc != B2C_OK) return rc;
    uint8_t *din = c->d_in + (size_t)fr0 * c->in_frame_stride;
    const uint8_t *src = frames + (size_t)f0 * in_frame_host;
    if (!in_pinned) {
      memcpy(c->h_in[slot], src, (size_t)cnt * in_frame_host);
      src = c->h_in[slot];
    }
    if (row_stride == c->in_row_stride) {
      CK(c, cudaMemcpyAsync(din, src, (size_t)cnt * in_frame_host, cudaMemcpyHostToDevice, c->s_h2d));
    } else {
      for (int f = 0; f < cnt; ++f)
        CK(c, cudaMemcpy2DAsync(din + (size_t)f * c->in_frame_stride, c->in_row_stride, src + (size_t)f * in_frame_host, row_stride, c->row_bytes(), c->rows_in(), cudaMemcpyHostToDevice, c->s_h2d));
    }
    CK(c, cudaEventRecord(c->ev_in[slot], c->s_h2d));

    CK(c, cudaStreamWaitEvent(c->s_main, c->ev_in[slot], 0));
    if ((rc = launch_stencil(c, din, c->in_row_stride, c->in_frame_stride, cnt, c->s_main, fr0)) != B2C_OK) return rc;
    uint8_t *dedges = c->d_edges + (size_t)fr0 * c->edges_frame_stride;
    if ((rc = launch_hysteresis(c, cnt, packed_bits ? nullptr : dedges, c->edges_pitch, c->edges_frame_stride, c->s_main, fr0)) != B2C_OK) return rc;
    CK(c, cudaEventRecord(c->ev_k[slot], c->s_main));
    CK(c, cudaStreamWaitEvent(c->s_d2h, c->ev_k[slot], 0));
    uint8_t *hout = out_pinned ? edges_out + (size_t)f0 * out_frame_host : c->h_out[slot];
    if (packed_bits) {   // the edge bit planes of the slot (rows are plane_pitch words apart, frames carry two ghost rows)
      CK(c, cudaMemcpy2DAsync(hout, out_row, E0(c) + (long long)fr0 * plane_frame_stride(c), (size_t)c->plane_pitch * 4, out_row, (size_t)h, cudaMemcpyDeviceToHost, c->s_d2h));
      for (int f = 1; f < cnt; ++f)
        CK(c, cudaMemcpy2DAsync(hout + (size_t)f * out_frame_host, out_row, E0(c) + (long long)(fr0 + f) * plane_frame_stride(c), (size_t)c->plane_pitch * 4, out_row, (size_t)h, cudaMemcpyDeviceToHost,
                                c->s_d2h));
    } else if (c->edges_frame_stride == out_frame_host) {
      CK(c, cudaMemcpyAsync(hout, dedges, (size_t)cnt * out_frame_host, cudaMemcpyDeviceToHost, c->s_d2h));
    } else {
      for (int f = 0; f < cnt; ++f)
        CK(c, cudaMemcpyAsync(hout + (size_t)f * out_frame_host, dedges + (size_t)f * c->edges_frame_stride, out_frame_host, cudaMemcpyDeviceToHost, c->s_d2h));
    }
    CK(c, cudaEventRecord(c->ev_out[slot], c->s_d2h));
    pending_out[slot] = k;
  }
  for (int s = 0; s < nslots; ++s) {
    int rc = drain(s);
    if (rc != B2C_OK) return rc;
  }
  CK(c, cudaStreamSynchronize(c->s_main));
  c->last_in = c->d_in + (size_t)c->acc_frame * c->in_frame_stride;   // (acc_frame = the last chunk's slot)
  c->last_row_stride = c->in_row_stride;
  c->have_frame = true;
  c->stages_valid = false;
  c->last_stage = B2C_STAGE_HYSTER;
  c->timings_valid = false;
  return B2C_OK;
}

int b2c_get_buffer(b2c_handle c, int id, const void **dev_ptr, size_t *pitch_bytes, int *elem_size)
{
  if (!c || !dev_ptr) return B2C_ERR_INVALID;
  if (!c->have_frame) return B2C_ERR_STATE;
  DevGuard g(c->dev);
  const void *p = nullptr;
  size_t pitch = 0;
  int es = 1;
  if (id >= B2C_BUF_MONO && id <= B2C_BUF_THRESH) {
    if (c->band) return B2C_ERR_UNSUPPORTED;
    // the stage buffers are recomputed from the retained input: only frames that live in the handle's own input buffer
    // qualify (a caller-owned device batch may be gone by now)
    if (!c->last_in) return B2C_ERR_STATE;
    if (!c->stages_valid) {
      int rc = launch_stencil_emit(c, c->last_in, c->last_row_stride, c->s_main);
      if (rc != B2C_OK) return rc;
      CK(c, cudaStreamSynchronize(c->s_main));
    }
    switch (id) {
    case B2C_BUF_MONO: p = c->d_mono; pitch = c->pitch8; break;
    case B2C_BUF_BLUR: p = c->d_blur; pitch = c->pitch8; break;
    case B2C_BUF_NMS: p = c->d_nms; pitch = c->pitch8; break;
    case B2C_BUF_THRESH: p = c->d_thresh; pitch = c->pitch8; break;
    default: p = c->d_grad; pitch = (size_t)c->pitchf * 4; es = 4; break;
    }
  } else if (id == B2C_BUF_EDGES) {
    if (!c->edges_valid) return B2C_ERR_STATE;   // the last run stopped before the hysteresis
    p = c->d_edges + (size_t)c->acc_frame * c->edges_frame_stride; pitch = c->edges_pitch;
  } else if (id == B2C_BUF_MAP2) {
    if (!c->map2_valid) {   // the planes S and C are the map; this view is its accessor format
      if (!c->d_map2) CK(c, cudaMalloc(&c->d_map2, (size_t)c->rows_alloc * c->map_pitch * 4));
      b2c::k_planes_to_map2<<<c->sm_count * 2, 256, 0, c->s_main>>>(reinterpret_cast<const uint16_t *>(S0(c) + (long long)c->acc_frame * plane_frame_stride(c)),
                                                                   reinterpret_cast<const uint16_t *>(C0(c) + (long long)c->acc_frame * plane_frame_stride(c)), c->plane_pitch * 2, c->d_map2,
                                                                   c->map_pitch, c->rows_alloc);
      CK(c, cudaGetLastError());
      c->launches++;
      c->map2_valid = true;
    }
    p = c->d_map2; pitch = (size_t)c->map_pitch * 4; es = 4;
  } else if (id == B2C_BUF_BITS) {
    if (!c->edges_valid) return B2C_ERR_STATE;
    p = E0(c) + (long long)c->acc_frame * plane_frame_stride(c); pitch = (size_t)c->plane_pitch * 4; es = 4;
  } else if (id == B2C_BUF_VIEW) {
    if (c->last_stage == B2C_STAGE_HYSTER) { p = c->d_edges + (size_t)c->acc_frame * c->edges_frame_stride; pitch = c->edges_pitch; }
    else { p = c->d_view; pitch = (size_t)c->w; }
  } else {
    return B2C_ERR_INVALID;
  }
  *dev_ptr = p;
  if (pitch_bytes) *pitch_bytes = pitch;
  if (elem_size) *elem_size = es;
  return B2C_OK;
}

int b2c_download(b2c_handle c, int id, void *host, size_t host_pitch)
{
  if (!c || !host) return B2C_ERR_INVALID;
  const void *p;
  size_t pitch;
  int es;
  int rc = b2c_get_buffer(c, id, &p, &pitch, &es);
  if (rc != B2C_OK) return rc;
  DevGuard g(c->dev);
  size_t row;
  if (id == B2C_BUF_MAP2) row = (size_t)c->map_pitch * 4;
  else if (id == B2C_BUF_BITS) row = (size_t)c->wpr * 4;
  else row = (size_t)c->w * es;
  if (host_pitch == 0) host_pitch = row;
  if (host_pitch < row) return B2C_ERR_SIZE;
  CK(c, cudaStreamSynchronize(c->s_main));
  CK(c, cudaMemcpy2D(host, host_pitch, p, pitch, row, c->rows_alloc, cudaMemcpyDeviceToHost));
  return B2C_OK;
}

// The reference's _sendOutputToOpenGL (cannyEdgeH.cu:154-212) without the GL part: the selected stage's u8 picture of
// the last run (for GRADIENT the saturated float2uchar view, cannyEdgeD.cu:35-50) copied device-to-device into the
// caller's buffer -- the pointer cudaGraphicsResourceGetMappedPointer returned for the PBO (tight pitch = width there,
// cannyEdgeH.cu:172,188-207).  Asynchronous on `stream` (0 = the handle's compute stream).
int b2c_copy_view(b2c_handle c, void *dev_dst, size_t dst_pitch, void *stream)
{
  if (!c || !dev_dst || c->band) return B2C_ERR_INVALID;
  if (dst_pitch == 0) dst_pitch = (size_t)c->w;
  if (dst_pitch < (size_t)c->w) return B2C_ERR_SIZE;
  const void *p;
  size_t pitch;
  int rc = b2c_get_buffer(c, B2C_BUF_VIEW, &p, &pitch, nullptr);
  if (rc != B2C_OK) return rc;
  DevGuard g(c->dev);
  cudaStream_t st = stream ? (cudaStream_t)stream : c->s_main;
  if (st != c->s_main) {   // the view was produced on the handle's stream
    CK(c, cudaEventRecord(c->ev_k[0], c->s_main));   // (an event of the host pipeline, idle outside b2c_run_batch_host)
    CK(c, cudaStreamWaitEvent(st, c->ev_k[0], 0));
  }
  CK(c, cudaMemcpy2DAsync(dev_dst, dst_pitch, p, pitch, (size_t)c->w, c->rows_alloc, cudaMemcpyDeviceToDevice, st));
  return B2C_OK;
}

int b2c_dev_alloc(b2c_handle c, size_t bytes, void **dev_ptr)
{
  if (!c || !dev_ptr) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  CK(c, cudaMalloc(dev_ptr, bytes));
  return B2C_OK;
}
int b2c_dev_free(b2c_handle c, void *p)
{
  if (!c) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  CK(c, cudaFree(p));
  return B2C_OK;
}
int b2c_dev_upload(b2c_handle c, void *dst, const void *src, size_t bytes)
{
  if (!c || !dst || !src) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  CK(c, cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return B2C_OK;
}
int b2c_dev_download(b2c_handle c, void *dst, const void *src, size_t bytes)
{
  if (!c || !dst || !src) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  CK(c, cudaStreamSynchronize(c->s_main));
  CK(c, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return B2C_OK;
}
int b2c_host_alloc(size_t bytes, void **host_ptr)
{
  if (!host_ptr) return B2C_ERR_INVALID;
  if (cudaMallocHost(host_ptr, bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    return B2C_ERR_NOMEM;
  }
  return B2C_OK;
}
int b2c_host_free(void *p)
{
  if (cudaFreeHost(p) != cudaSuccess) {
    (void)cudaGetLastError();
    return B2C_ERR_CUDA;
  }
  return B2C_OK;
}
int b2c_sync(b2c_handle c)
{
  if (!c) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  if (c->band) {   // a band's work may sit on the caller's stream and on the side stream of the pushes
    CK(c, cudaDeviceSynchronize());
    return B2C_OK;
  }
  CK(c, cudaStreamSynchronize(c->s_main));
  CK(c, cudaStreamSynchronize(c->s_h2d));
  CK(c, cudaStreamSynchronize(c->s_d2h));
  return B2C_OK;
}
void *b2c_stream(b2c_handle c) { return c ? (void *)c->s_main : nullptr; }

// A thresholded map (0 / 128 / 255, the reference's d_threshImage: cannyEdgeD.cu:274-292) from host memory into the
// planes of frame 0, in place of a stencil run: hysteresis on maps produced elsewhere.  Blocking.
int b2c_load_thresh(b2c_handle c, const uint8_t *host_thresh, size_t row_stride)
{
  if (!c || !host_thresh) return B2C_ERR_INVALID;
  if (row_stride < (size_t)c->w) return B2C_ERR_SIZE;
  DevGuard g(c->dev);
  const int pp = c->plane_pitch, h = c->rows_alloc;
  std::vector<uint32_t> S((size_t)pp * h, 0u), C((size_t)pp * h, 0u);
  for (int y = 0; y < h; ++y) {
    const uint8_t *r = host_thresh + (size_t)y * row_stride;
    for (int x = 0; x < c->w; ++x) {
      if (r[x] == 255) S[(size_t)y * pp + (x >> 5)] |= 1u << (x & 31);
      if (r[x] >= 128) C[(size_t)y * pp + (x >> 5)] |= 1u << (x & 31);
    }
  }
  CK(c, cudaStreamSynchronize(c->s_main));
  CK(c, cudaMemcpy(S0(c), S.data(), S.size() * 4, cudaMemcpyHostToDevice));
  CK(c, cudaMemcpy(C0(c), C.data(), C.size() * 4, cudaMemcpyHostToDevice));
  c->have_frame = true;
  c->stages_valid = false;
  c->map2_valid = false;
  c->edges_valid = false;
  c->last_in = nullptr;
  return B2C_OK;
}

// ---- row-band mode --------------------------------------------------------------------------------
// Order of one band step (stream order; [p2p] = only with peer wiring):
//   [halo push] stencil (its CTAs next to a seam wait for the halo rows) | k_uf_tile, k_uf_border, k_seam_publish [+ push],
//   k_uf_resolve<LIST> | k_seam_solve [waits for the peers' records first], k_uf_resolve_list
// The halo rows travel while the interior rows are computed, the seam records while the band resolves.
namespace
{
// scratch of the seam kernels
int seam_alloc(b2c_ctx *c)
{
  if (c->d_seam_rec) return B2C_OK;
  const int cap = b2c::seam_cap(c->wpr), hs = b2c::seam_hash_size(c->wpr);
  CK(c, cudaMalloc(&c->d_seam_rec, b2c::seam_rec_words(c->wpr) * 4));
  CK(c, cudaMemset(c->d_seam_rec, 0, b2c::seam_rec_words(c->wpr) * 4));
  CK(c, cudaMalloc(&c->d_seam_roots, (size_t)2 * cap * sizeof(int)));
  CK(c, cudaMalloc(&c->d_seam_hkey, (size_t)hs * sizeof(int)));
  CK(c, cudaMalloc(&c->d_seam_hval, (size_t)hs * sizeof(int)));
  CK(c, cudaMemset(c->d_seam_hkey, 0, (size_t)hs * sizeof(int)));
  CK(c, cudaMalloc(&c->d_seam_P, ((size_t)b2c::SEAM_MAXW * 2 * cap + 1) * sizeof(int)));
  CK(c, cudaMalloc(&c->d_seam_ctl, 8 * sizeof(int)));
  CK(c, cudaMemset(c->d_seam_ctl, 0, 8 * sizeof(int)));
  CK(c, cudaMallocHost(&c->h_seam_ctl, 8 * sizeof(int)));
  memset(c->h_seam_ctl, 0, 8 * sizeof(int));
  c->ucap = (int)std::max<long long>(4096, (long long)c->wpr * c->rows_alloc / 4);
  CK(c, cudaMalloc(&c->d_ulist, (size_t)c->ucap * sizeof(uint2)));
  CK(c, cudaFuncSetAttribute(b2c::k_seam_publish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b2c::seam_publish_smem(4095)));
  CK(c, cudaFuncSetAttribute(b2c::k_seam_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b2c::seam_solve_smem()));
  for (auto &e : c->ev_s) CK(c, cudaEventCreate(&e));
  for (auto &e : c->ev_b) CK(c, cudaEventCreate(&e));
  return B2C_OK;
}
void fill_seam_band(b2c_ctx *c, b2c::B2cSeamBand &b)
{
  b.S = S0(c);
  b.C = C0(c);
  b.plane_pitch = c->plane_pitch;
  b.wpr = c->wpr;
  b.h = c->rows_alloc;
  b.parent = c->d_parent;
  b.roots = c->d_seam_roots;
  b.hkey = c->d_seam_hkey;
  b.hval = c->d_seam_hval;
  b.hsize = b2c::seam_hash_size(c->wpr);
  b.shash = c->seam_force_global ? 0 : b2c::SEAM_SHASH;
  b.snodes = c->seam_force_global ? 0 : b2c::SEAM_SNODES;
  b.ctl = c->d_seam_ctl;
}
uint32_t *own_mail(b2c_ctx *c) { return (uint32_t *)c->peer_mail[c->p2p_rank]; }

// rows [r0, r0 + nrows) of the band through the fused stencil
int launch_stencil_rows(b2c_ctx *c, const uint8_t *band_row0, size_t row_stride, int r0, int nrows, cudaStream_t st)
{
  if (nrows <= 0) return B2C_OK;
  B2cStencilParams p;
  fill_stencil_params(c, p, band_row0 + (size_t)r0 * row_stride, row_stride, 0, 1);
  p.h = nrows;
  p.y0 = c->band_y0 + r0;
  p.pl_S += (long long)r0 * p.pl_pitch16;
  p.pl_C += (long long)r0 * p.pl_pitch16;
  if (c->stencil_impl == 0 && b2c::march_supported(p)) {
    cudaError_t e = b2c::march_launch(p, c->sm_count, c->march_ctas_per_sm, r0 == 0 && nrows == c->rows_alloc ? c->march_rb : 0, st, c->march_extra_smem);
    if (e != cudaSuccess) return set_err(c, e, "k_stencil_march launch");
  } else {
    dim3 grid((c->w + b2c::TILE_W - 1) / b2c::TILE_W, (nrows + b2c::TILE_H - 1) / b2c::TILE_H, 1);
    b2c::k_stencil_tile<false><<<grid, b2c::TILE_THREADS, b2c::TILE_SMEM, st>>>(p);
    CK(c, cudaGetLastError());
  }
  c->launches++;
  c->map2_valid = false;
  c->edges_valid = false;
  c->have_frame = true;
  return B2C_OK;
}
}// namespace

int b2c_band_stencil(b2c_handle c, const uint8_t *dev_bgr_band_row0, size_t row_stride, void *stream)
{
  if (!c || !c->band || !dev_bgr_band_row0) return B2C_ERR_INVALID;
  if (row_stride < (size_t)c->w * 3) return B2C_ERR_SIZE;
  DevGuard g(c->dev);
  return launch_stencil_rows(c, dev_bgr_band_row0, row_stride, 0, c->rows_alloc, stream ? (cudaStream_t)stream : c->s_main);
}

// Band-local hysteresis.  The seam record is built as soon as the forest is complete (and, with peer wiring, pushed to
// every rank) so that it travels while the band resolves.
int b2c_band_hysteresis(b2c_handle c, void *stream)
{
  if (!c || !c->band) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  cudaStream_t st = stream ? (cudaStream_t)stream : c->s_main;
  B2cHystParams p;
  fill_hyst_params(c, p, 1, c->d_edges, c->edges_pitch, c->edges_frame_stride);
  const bool pt = c->hyst_phase_timing;
  if (pt) cudaEventRecord(c->ev_s[0], st);
  const dim3 gt((c->wpr + b2c::UT_WORDS - 1) / b2c::UT_WORDS, (c->rows_alloc + b2c::UT_ROWS - 1) / b2c::UT_ROWS, 1);
  const int T = b2c::UFK_THREADS;
  b2c::k_uf_tile<<<gt, b2c::UT_THREADS, b2c::UT_SMEM, st>>>(p, c->d_blist, c->d_bcount, c->bcap);
  b2c::k_uf_border<<<dim3(std::max(1, (c->bcap + T - 1) / T), 1, 1), T, 0, st>>>(p, c->d_blist, c->d_bcount, c->bcap);
  if (pt) cudaEventRecord(c->ev_s[1], st);
  // seam record
  b2c::B2cSeamBand b;
  fill_seam_band(c, b);
  c->seam_run += 1;
  const int world = c->p2p_world, rank = c->p2p_rank, par = c->seam_run & 1;
  const bool p2p = world >= 2 && c->d_mailbox;
  uint32_t *rec = p2p ? own_mail(c) + b2c::bp_seam_slot(c->wpr, par, rank) : c->d_seam_rec;
  b2c::k_seam_publish<<<1, b2c::SEAM_THREADS, b2c::seam_publish_smem(c->wpr), st>>>(b, rec, c->seam_run);
  c->launches += 3;
  if (p2p) {   // the record travels to every rank's mailbox on the side stream while this stream resolves the band
    b2c::B2cSeamPeers q;
    memset(&q, 0, sizeof(q));
    for (int k = 0; k < world; ++k) {
      q.slot[k] = (uint32_t *)c->peer_mail[k] + b2c::bp_seam_slot(c->wpr, par, rank);
      q.flag[k] = (uint32_t *)c->peer_mail[k] + b2c::bp_seam_flag(c->wpr, par, rank);
    }
    q.world = world;
    q.rank = rank;
    CK(c, cudaEventRecord(c->ev_fork, st));
    CK(c, cudaStreamWaitEvent(c->s_side, c->ev_fork, 0));
    b2c::k_seam_push<<<1, b2c::SEAM_THREADS, 0, c->s_side>>>(q, c->wpr, c->seam_run);
    CK(c, cudaEventRecord(c->ev_join, c->s_side));
    c->launches++;
  }
  if (pt) cudaEventRecord(c->ev_s[2], st);
  // resolve + expansion; the words that stay unresolved go to the list of the seam pass
  const int tx = c->wpr >= 256 ? 256 : c->wpr > 32 ? 64 : 32, ty = 256 / tx;
  const dim3 gr((c->wpr + tx - 1) / tx, (c->rows_alloc + 2 * ty - 1) / (2 * ty), 1), br(tx, ty);
  b2c::k_uf_resolve<true, true><<<gr, br, 0, st>>>(p, c->d_bcount, c->d_ulist, c->d_seam_ctl + 4, c->ucap);
  if (pt) cudaEventRecord(c->ev_s[3], st);
  CK(c, cudaGetLastError());
  c->launches++;
  c->edges_valid = true;
  return B2C_OK;
}

namespace
{
// solve on the gathered records + the pass over the still unresolved words that promotes what the solve hung under node 0
int seam_solve(b2c_ctx *c, const uint32_t *const *recs, int world, int rank, const uint32_t *flags, cudaStream_t st)
{
  b2c::B2cSeamBand b;
  fill_seam_band(c, b);
  b2c::B2cSeamAll a;
  memset(&a, 0, sizeof(a));
  for (int r = 0; r < world; ++r) a.rec[r] = recs[r];
  a.world = world;
  a.rank = rank;
  a.P = c->d_seam_P;
  a.flags = flags;
  a.run_id = c->seam_run;
  b2c::k_seam_solve<<<1, b2c::SEAM_THREADS, b2c::seam_solve_smem(), st>>>(b, a);
  if (c->hyst_phase_timing) cudaEventRecord(c->ev_s[5], st);
  B2cHystParams p;
  fill_hyst_params(c, p, 1, c->d_edges, c->edges_pitch, c->edges_frame_stride);
  b2c::k_uf_resolve_list<true><<<c->sm_count * 2, b2c::UFK_THREADS, 0, st>>>(p, c->d_ulist, c->d_seam_ctl + 4, c->ucap, c->d_seam_ctl);
  if (c->hyst_phase_timing) cudaEventRecord(c->ev_s[6], st);
  CK(c, cudaGetLastError());
  c->launches += 2;
  return B2C_OK;
}
}// namespace

int b2c_band_seam_bytes(b2c_handle c, size_t *bytes)
{
  if (!c || !c->band || !bytes) return B2C_ERR_INVALID;
  *bytes = b2c::seam_rec_words(c->wpr) * 4;
  return B2C_OK;
}

int b2c_band_seam_record(b2c_handle c, void **record_dev)
{
  if (!c || !c->band || !record_dev) return B2C_ERR_INVALID;
  *record_dev = c->d_seam_rec;
  return B2C_OK;
}

int b2c_band_seam_solve(b2c_handle c, const void *all_records_dev, int world, int rank, void *stream)
{
  if (!c || !c->band || !all_records_dev || world < 1 || world > b2c::SEAM_MAXW || rank < 0 || rank >= world) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  cudaStream_t st = stream ? (cudaStream_t)stream : c->s_main;
  if (c->hyst_phase_timing) { cudaEventRecord(c->ev_s[3], st); cudaEventRecord(c->ev_s[4], st); }
  const uint32_t *recs[b2c::SEAM_MAXW];
  const size_t stride = b2c::seam_rec_words(c->wpr);
  for (int r = 0; r < world; ++r) recs[r] = (const uint32_t *)all_records_dev + (size_t)r * stride;
  return seam_solve(c, recs, world, rank, nullptr, st);
}

int b2c_band_status(b2c_handle c, int *promoted_runs, int *error)
{
  if (!c || !c->band) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  CK(c, cudaDeviceSynchronize());   // the band's work may be on any stream
  CK(c, cudaMemcpy(c->h_seam_ctl, c->d_seam_ctl, 8 * sizeof(int), cudaMemcpyDeviceToHost));
  CK(c, cudaMemset(c->d_seam_ctl + 2, 0, sizeof(int)));   // a time-out is reported once
  if (promoted_runs) *promoted_runs = c->h_seam_ctl[3];
  if (error) *error = c->h_seam_ctl[2];
  return B2C_OK;
}

// ---- peer-to-peer exchange for ranks of one box (see k_band_p2p.cuh / k_band_seam.cuh) -------------------------------
int b2c_band_input(b2c_handle c, void **dev_ptr, size_t *row_stride)
{
  if (!c || !c->band || !dev_ptr) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  const size_t stride = round_up((size_t)c->w * 3, 16);
  if (!c->d_band_in) {
    CK(c, cudaMalloc(&c->d_band_in, stride * (c->rows_alloc + 8)));
    CK(c, cudaMemset(c->d_band_in, 0, stride * (c->rows_alloc + 8)));
  }
  *dev_ptr = c->d_band_in;
  if (row_stride) *row_stride = stride;
  return B2C_OK;
}

namespace
{
int p2p_alloc(b2c_ctx *c)
{
  if (c->d_mailbox) return B2C_OK;
  const size_t bytes = b2c::bp_mailbox_words(c->wpr) * sizeof(uint32_t);
  CK(c, cudaMalloc(&c->d_mailbox, bytes));
  CK(c, cudaMemset(c->d_mailbox, 0, bytes));
  void *in;
  return b2c_band_input(c, &in, nullptr);
}
}// namespace

int b2c_band_p2p_export(b2c_handle c, void *blob_144)
{
  if (!c || !c->band || !blob_144) return B2C_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DevGuard g(c->dev);
  int rc = p2p_alloc(c);
  if (rc != B2C_OK) return rc;
  cudaIpcMemHandle_t h[2];
  CK(c, cudaIpcGetMemHandle(&h[0], c->d_mailbox));
  CK(c, cudaIpcGetMemHandle(&h[1], c->d_band_in));
  memset(blob_144, 0, 144);
  memcpy(blob_144, h, 128);
  const int rows = c->rows_alloc;
  memcpy((char *)blob_144 + 128, &rows, sizeof(int));
  return B2C_OK;
}

int b2c_band_p2p_open(b2c_handle c, const void *all_blobs, int world, int rank)
{
  if (!c || !c->band || !all_blobs || !c->d_mailbox || world < 2 || world > b2c::BP_MAXW || rank < 0 || rank >= world) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  for (int k = 0; k < world; ++k) {
    const char *blob = (const char *)all_blobs + 144 * k;
    if (k == rank) {
      c->peer_mail[k] = c->d_mailbox;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, blob, 64);
    CK(c, cudaIpcOpenMemHandle(&c->peer_mail[k], h, cudaIpcMemLazyEnablePeerAccess));
    if (k == rank - 1 || k == rank + 1) {
      memcpy(&h, blob + 64, 64);
      CK(c, cudaIpcOpenMemHandle(&c->peer_in[k == rank - 1 ? 0 : 1], h, cudaIpcMemLazyEnablePeerAccess));
      if (k == rank - 1) memcpy(&c->peer_rows_up, blob + 128, sizeof(int));
    }
  }
  c->peers_ipc = true;
  c->p2p_world = world;
  c->p2p_rank = rank;
  return B2C_OK;
}

// the same wiring for bands that live in ONE process (any devices with peer access, or one device): the "peers" are
// the other handles themselves.  Used by single-process drivers and by the single-GPU test of the peer-to-peer kernels.
int b2c_band_p2p_open_local(b2c_handle c, const b2c_handle *all_handles, int world, int rank)
{
  if (!c || !c->band || !all_handles || world < 2 || world > b2c::BP_MAXW || rank < 0 || rank >= world || all_handles[rank] != c) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  for (int k = 0; k < world; ++k) {
    b2c_ctx *o = all_handles[k];
    if (!o || !o->band || o->w != c->w) return B2C_ERR_INVALID;
    {
      DevGuard g2(o->dev);
      int rc = p2p_alloc(o);
      if (rc != B2C_OK) return rc;
    }
    if (o->dev != c->dev) {
      int can = 0;
      CK(c, cudaDeviceCanAccessPeer(&can, c->dev, o->dev));
      if (!can) return B2C_ERR_UNSUPPORTED;
      cudaError_t e = cudaDeviceEnablePeerAccess(o->dev, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return set_err(c, e, "cudaDeviceEnablePeerAccess");
      (void)cudaGetLastError();
    }
    c->peer_mail[k] = o->d_mailbox;
    if (k == rank - 1) { c->peer_in[0] = o->d_band_in; c->peer_rows_up = o->rows_alloc; }
    if (k == rank + 1) c->peer_in[1] = o->d_band_in;
  }
  c->peers_ipc = false;
  c->p2p_world = world;
  c->p2p_rank = rank;
  return B2C_OK;
}

namespace
{
void fill_p2p(b2c_ctx *c, b2c::B2cBandP2P &q)
{
  memset(&q, 0, sizeof(q));
  for (int k = 0; k < c->p2p_world; ++k) q.mail[k] = (uint32_t *)c->peer_mail[k];
  q.world = c->p2p_world;
  q.rank = c->p2p_rank;
  q.wpr = c->wpr;
  q.ctl = c->d_seam_ctl;
  q.in_up = (uint8_t *)c->peer_in[0];
  q.in_dn = (uint8_t *)c->peer_in[1];
  q.in_own = c->d_band_in;
  q.in_stride = (long long)round_up((size_t)c->w * 3, 16);
  q.rows_own = c->rows_alloc;
  q.rows_up = c->peer_rows_up;
  q.row_bytes = c->w * 3;
}
}// namespace

// Stencil of the band in the handle's own input buffer (b2c_band_input) with the halo exchange over peer memory hidden
// behind it: my first / last 4 rows are stored into the neighbours' buffers, then ONE stencil launch over the whole band
// whose CTAs next to a seam wait -- on the device, when they get there -- for the neighbour's arrival counter (the
// marching kernel; geometries it does not take wait in a kernel of their own before the launch).
// phase: B2C_P2P_ALL, or B2C_P2P_PUSH (the stores) then B2C_P2P_WAIT (the stencil): a single-process driver of several
// bands issues ALL pushes before the first stencil -- a device-side wait must never be queued ahead of the store it
// waits for.
int b2c_band_p2p_stencil(b2c_handle c, void *stream, int phase)
{
  if (!c || !c->band || c->p2p_world < 2 || !c->d_band_in || phase < B2C_P2P_ALL || phase > B2C_P2P_WAIT) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  cudaStream_t st = stream ? (cudaStream_t)stream : c->s_main;
  b2c::B2cBandP2P q;
  fill_p2p(c, q);
  const size_t stride = (size_t)q.in_stride;
  const uint8_t *row0 = c->d_band_in + 4 * stride;
  const int cpr = (c->w * 3 + 15) / 16, nblocks = (4 * cpr + 255) / 256;
  const bool up = c->p2p_rank > 0, dn = c->p2p_rank < c->p2p_world - 1;
  const bool pt = c->hyst_phase_timing;
  if (phase != B2C_P2P_WAIT) {
    if (pt) cudaEventRecord(c->ev_b[0], st);
    c->p2p_run += 1;
    // on the side stream, behind whatever filled the input buffer on `stream`: the stencil does not wait for the stores
    CK(c, cudaEventRecord(c->ev_fork, st));
    CK(c, cudaStreamWaitEvent(c->s_side, c->ev_fork, 0));
    b2c::k_band_push_halo<<<dim3(nblocks, 2), 256, 0, c->s_side>>>(q);
    CK(c, cudaEventRecord(c->ev_join, c->s_side));
    c->launches++;
    if (pt) cudaEventRecord(c->ev_b[1], st);
  }
  if (phase != B2C_P2P_PUSH) {
    B2cStencilParams p;
    fill_stencil_params(c, p, row0, stride, 0, 1);
    if (c->stencil_impl == 0 && b2c::march_supported(p)) {
      const uint32_t *mine = own_mail(c);
      p.halo_cnt_up = up ? mine + b2c::bp_halo_flag(c->wpr, 0) : nullptr;   // rows from above
      p.halo_cnt_dn = dn ? mine + b2c::bp_halo_flag(c->wpr, 1) : nullptr;   // rows from below
      p.halo_need = c->p2p_run * nblocks;
      p.halo_err = c->d_seam_ctl + 2;
      if (pt) cudaEventRecord(c->ev_b[2], st);
      cudaError_t e = b2c::march_launch(p, c->sm_count, c->march_ctas_per_sm, c->march_rb, st, c->march_extra_smem);
      if (e != cudaSuccess) return set_err(c, e, "k_stencil_march launch");
      c->launches++;
      c->map2_valid = false;
      c->edges_valid = false;
      c->have_frame = true;
    } else {
      b2c::k_band_wait_halo<<<1, 1, 0, st>>>(q, c->p2p_run, nblocks);
      c->launches++;
      if (pt) cudaEventRecord(c->ev_b[2], st);
      int rc = launch_stencil_rows(c, row0, stride, 0, c->rows_alloc, st);
      if (rc != B2C_OK) return rc;
    }
    CK(c, cudaStreamWaitEvent(st, c->ev_join, 0));   // later work on `stream` (a new image in the input buffer) follows my stores
    if (pt) cudaEventRecord(c->ev_b[3], st);
  }
  CK(c, cudaGetLastError());
  return B2C_OK;
}

// cross-band hysteresis over peer memory, after b2c_band_hysteresis (which published and pushed my record): wait for all
// records, solve, promote.  Asynchronous on `stream`; b2c_band_status reports a peer time-out.
int b2c_band_p2p_seam(b2c_handle c, void *stream)
{
  if (!c || !c->band || c->p2p_world < 2 || !c->d_mailbox) return B2C_ERR_INVALID;
  DevGuard g(c->dev);
  cudaStream_t st = stream ? (cudaStream_t)stream : c->s_main;
  const int world = c->p2p_world, rank = c->p2p_rank, par = c->seam_run & 1;
  uint32_t *mine = own_mail(c);
  CK(c, cudaStreamWaitEvent(st, c->ev_join, 0));   // my own push (long finished: the band resolved meanwhile)
  if (c->hyst_phase_timing) cudaEventRecord(c->ev_s[4], st);
  const uint32_t *recs[b2c::SEAM_MAXW];
  for (int r = 0; r < world; ++r) recs[r] = mine + b2c::bp_seam_slot(c->wpr, par, r);
  return seam_solve(c, recs, world, rank, mine + b2c::bp_seam_flag(c->wpr, par, 0), st);   // (the solve kernel waits for the flags)
}

// ---- misc -----------------------------------------------------------------------------------------
const char *b2c_strerror(int s)
{
  switch (s) {
  case B2C_OK: return "ok";
  case B2C_ERR_INVALID: return "invalid argument";
  case B2C_ERR_CUDA: return "CUDA error";
  case B2C_ERR_NOMEM: return "out of memory";
  case B2C_ERR_SIZE: return "frame geometry mismatch";
  case B2C_ERR_UNSUPPORTED: return "unsupported";
  case B2C_ERR_STATE: return "no frame has been run yet";
  default: return "unknown status";
  }
}
const char *b2c_last_cuda_error(b2c_handle c) { return c ? c->last_err.c_str() : ""; }
const char *b2c_version(void) { return "b200canny 0.1 (sm_100a)"; }
int b2c_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}
long long b2c_launch_count(b2c_handle c) { return c ? c->launches : 0; }

// Binds the calling thread (and the threads it starts later) to the CPUs of the NUMA node the GPU hangs on, so that the
// pinned frame rings allocated afterwards (first touch) and the staging memcpys are local to the GPU's PCIe root.  Linux
// sysfs; returns the node, or -1 if the platform does not say (nothing is changed then).
int b2c_bind_host_to_device(int device)
{
  char bus[32] = { 0 };
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1;
  }
  for (char *q = bus; *q; ++q) *q = (char)tolower((unsigned char)*q);
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE *f = fopen(path, "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  if (node < 0) return -1;
  snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  f = fopen(path, "r");
  if (!f) return -1;
  char list[4096] = { 0 };
  const bool ok = fgets(list, sizeof(list), f) != nullptr;
  fclose(f);
  if (!ok) return -1;
  cpu_set_t set;
  CPU_ZERO(&set);
  int count = 0;
  for (char *q = list; *q;) {   // "0-55,112-167"
    char *e;
    const long a = strtol(q, &e, 10);
    if (e == q) break;
    long b = a;
    if (*e == '-') b = strtol(e + 1, &e, 10);
    for (long k = a; k <= b && k < CPU_SETSIZE; ++k) { CPU_SET((int)k, &set); ++count; }
    q = (*e == ',') ? e + 1 : e;
    if (*e != ',') break;
  }
  if (count == 0 || sched_setaffinity(0, sizeof(set), &set) != 0) return -1;
  return node;
}

int b2c_set_option(b2c_handle c, const char *name, int value)
{
  if (!c || !name) return B2C_ERR_INVALID;
  if (!strcmp(name, "stencil_impl")) {
    if (value < 0 || value > 1) return B2C_ERR_INVALID;
    c->stencil_impl = value;
    return B2C_OK;
  }
  if (!strcmp(name, "uf_spread")) {
    if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8) return B2C_ERR_INVALID;
    c->uf_spread = value;
    return B2C_OK;
  }
  if (!strcmp(name, "hyst_phase_timing")) {
    c->hyst_phase_timing = value != 0;
    if (c->hyst_phase_timing && !c->ev_h[0])
      for (auto &e : c->ev_h) CK(c, cudaEventCreate(&e));
    return B2C_OK;
  }
  if (!strcmp(name, "seam_force_global")) {
    c->seam_force_global = value != 0;
    return B2C_OK;
  }
  if (!strcmp(name, "march_extra_smem")) {
    if (value < 0 || value > 64 * 1024) return B2C_ERR_INVALID;
    c->march_extra_smem = value;
    return B2C_OK;
  }
  if (!strcmp(name, "march_rb")) {
    if (value < 0) return B2C_ERR_INVALID;
    c->march_rb = value;
    return B2C_OK;
  }
  return B2C_ERR_INVALID;
}
int b2c_get_info(b2c_handle c, const char *name)
{
  if (!c || !name) return B2C_ERR_INVALID;
  if (!strncmp(name, "band_stencil_us", 15)) {   // b2c_band_p2p_stencil with "hyst_phase_timing": halo push, gap, stencil incl. its waits (us)
    const int k = name[15] - '0';
    if (k < 0 || k > 2 || !c->ev_b[0]) return B2C_ERR_INVALID;
    float ms = 0;
    if (cudaEventSynchronize(c->ev_b[3]) != cudaSuccess || cudaEventElapsedTime(&ms, c->ev_b[k], c->ev_b[k + 1]) != cudaSuccess) return B2C_ERR_CUDA;
    return (int)(ms * 1000.0f + 0.5f);
  }
  if (!strncmp(name, "seam_phase_us", 13)) {   // band mode with "hyst_phase_timing": tile+border, publish(+push), resolve, wait, solve, list pass (us)
    const int k = name[13] - '0';
    if (k < 0 || k > 5 || !c->ev_s[0]) return B2C_ERR_INVALID;
    float ms = 0;
    if (cudaEventSynchronize(c->ev_s[6]) != cudaSuccess || cudaEventElapsedTime(&ms, c->ev_s[k], c->ev_s[k + 1]) != cudaSuccess) return B2C_ERR_CUDA;
    return (int)(ms * 1000.0f + 0.5f);
  }
  if (!strncmp(name, "hyst_phase_us", 13)) {   // phase times of the last union-find hysteresis run with "hyst_phase_timing" on (us)
    const int k = name[13] - '0';
    if (k < 0 || k > 2 || !c->ev_h[0]) return B2C_ERR_INVALID;
    float ms = 0;
    if (cudaEventSynchronize(c->ev_h[3]) != cudaSuccess || cudaEventElapsedTime(&ms, c->ev_h[k], c->ev_h[k + 1]) != cudaSuccess) return B2C_ERR_CUDA;
    return (int)(ms * 1000.0f + 0.5f);
  }
  if (!strcmp(name, "sm_count")) return c->sm_count;
  if (!strcmp(name, "stencil_impl")) return c->stencil_impl;
  if (!strcmp(name, "march_ctas_per_sm")) return c->march_ctas_per_sm;
  if (!strcmp(name, "march_band_rows")) return c->march_rb > 0 ? c->march_rb : b2c::march_band_rows(c->w, c->rows_alloc, c->max_batch, c->sm_count, c->march_ctas_per_sm);
  if (!strcmp(name, "in_row_stride")) return (int)c->in_row_stride;
  if (!strcmp(name, "plane_pitch_words")) return c->plane_pitch;
  if (!strcmp(name, "map_pitch_words")) return c->map_pitch;
  return B2C_ERR_INVALID;
}

}// extern "C"
