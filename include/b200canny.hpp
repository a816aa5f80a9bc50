// b200canny.hpp -- header-only C++17 host classes over the C ABI of libb200canny.so (include/b200canny.h).
//
// They rebuild the class surface of the reference's src/cvp so that its call sites keep compiling:
//   cvp::CannyStage          <- src/cvp/define.hpp:9-17 (same values), CANNY_STAGES names :27-34
//   cvp::cuda::CannyEdge     <- src/cvp/cannyEdgeH.hpp:17-110: ctor for given frame dims, run(frame, finalStage),
//                               set/getLow/HighThreshold (:25-29, same clamping), enable/isKernelProfilingEnabled (:31-32)
//                               + explicit accessors mono()/blur()/gradient()/nms()/thresh()/edges() for the buffers the
//                               reference only exposes through the finalStage switch (cannyEdgeH.cu:169-207)
//   cvp::cvPipeline          <- src/cvp/cvPipeline.hpp:20-39: ctor(pbo, cols, rows, nbChannels), bool process(frame, stage),
//                               threshold + profiling forwarding; process() returns false where the reference does
//                               (cvPipeline.cpp:19-41)
//   b2c::TimerManager        <- src/utils/timer.hpp:13-67: createTimer / addTime / getAverageTime / begin/endTimerList,
//                               the sink CannyEdge::run feeds per stage name when kernel profiling is on
//                               (cannyEdgeH.cu:35-37, 409-430), read by the UI table (imguiApp.cpp:357-376)
//   b2c::BandRunner          <- no reference counterpart: one row band of a large image per GPU (BASELINE config 5),
//                               the C++ driver over the b2c_band_* ABI (cudacam_b200/bands.py is its Python mirror)
// No CUDA, OpenCV or GL header is needed to include this file.  Frames are passed either as b2c::FrameView or as
// anything shaped like cv::Mat (members data, rows, cols, step, channels(), empty()), so `process(cv::Mat, stage)`
// compiles unchanged when OpenCV is present.  Errors: the reference exits the process on any CUDA error
// (helper.hpp:4-17); here the classes throw b2c::Error carrying the C ABI status.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "b200canny.h"

namespace b2c
{
struct Error : std::runtime_error {
  int status;
  Error(int st, const std::string &what) : std::runtime_error(what + ": " + b2c_strerror(st)), status(st) {}
};

// What the classes need to know about a frame (the fields of cv::Mat that cannyEdgeH.cu:122-152 reads).
struct FrameView {
  const uint8_t *data = nullptr;
  int rows = 0, cols = 0, nbChannels = 3;
  size_t step = 0;   // bytes per row
  bool empty() const { return !data || rows <= 0 || cols <= 0; }
  int channels() const { return nbChannels; }
};

// cv::Mat::type() values used by the reference's input check (cvPipeline.cpp:32): depth in the low 3 bits (CV_8U = 0),
// channels - 1 above them
constexpr int kCvDepthMask = 7, kCv8U = 0;
template <class M, class = void> struct has_type : std::false_type {};
template <class M> struct has_type<M, std::void_t<decltype(std::declval<const M &>().type())>> : std::true_type {};

// true unless the frame says it is not 8-bit (a cv::Mat of CV_16U / CV_32F ... must be refused, cvPipeline.cpp:32-36)
template <class M> inline bool is_8bit(const M &m)
{
  if constexpr (has_type<M>::value) return (m.type() & kCvDepthMask) == kCv8U;
  else return true;
}

template <class M> inline FrameView view_of(const M &m)
{
  if constexpr (std::is_same_v<M, FrameView>) return m;
  else {
    FrameView v;
    v.data = reinterpret_cast<const uint8_t *>(m.data);
    v.rows = m.rows;
    v.cols = m.cols;
    v.nbChannels = m.channels();
    v.step = static_cast<size_t>(m.step);
    return v;
  }
}

// src/utils/timer.hpp:6-11
struct Timer {
  double totalTime = 0.0;
  size_t nbCount = 0;
  float averageTime() const { return nbCount > 0 ? (float)(totalTime / nbCount) : 0.0f; }
};

// src/utils/timer.hpp:13-67: running per-name totals.  The reference's is a process-wide singleton (Get()) that
// cvp::cuda::CannyEdge feeds and the UI reads; the same here.  Unknown names are ignored by addTime and read as 0
// (the reference logs an error, timer.hpp:36-39, 49-52).
class TimerManager
{
public:
  static TimerManager &Get()
  {
    static TimerManager manager;
    return manager;
  }
  void createTimer(const std::string &name) { m_timers.insert(std::make_pair(name, Timer())); }
  void addTime(const std::string &name, double time)
  {
    auto it = m_timers.find(name);
    if (it == m_timers.end()) return;
    it->second.totalTime += time;
    it->second.nbCount++;
  }
  double getAverageTime(const std::string &name) const
  {
    auto it = m_timers.find(name);
    return it != m_timers.end() && it->second.nbCount > 0 ? it->second.totalTime / it->second.nbCount : 0.0;
  }
  std::map<std::string, Timer>::const_iterator beginTimerList() const { return m_timers.cbegin(); }
  std::map<std::string, Timer>::const_iterator endTimerList() const { return m_timers.cend(); }
  void reset()
  {
    for (auto &kv : m_timers) kv.second = Timer();
  }

private:
  TimerManager() = default;
  TimerManager(const TimerManager &) = delete;
  TimerManager &operator=(const TimerManager &) = delete;
  std::map<std::string, Timer> m_timers;
};
}// namespace b2c

namespace cvp
{
// src/cvp/define.hpp:9-17
enum CannyStage { MONO = 0, GAUSSIAN = 1, GRADIENT = 2, NMS = 3, THRESH = 4, HYSTER = 5 };

// src/cvp/define.hpp:27-34
inline const std::map<CannyStage, std::string> &cannyStages()
{
  static const std::map<CannyStage, std::string> m = { { MONO, "1/6 Mono Conversion" }, { GAUSSIAN, "2/6 Gaussian Noise Removal" },
    { GRADIENT, "3/6 Gradient Computation" }, { NMS, "4/6 Non Maximum Suppression" }, { THRESH, "5/6 Double Threshold" }, { HYSTER, "6/6 Hysteresis" } };
  return m;
}

namespace cuda
{
class CannyEdge
{
public:
  // Reference: CannyEdge(unsigned pbo, unsigned w, unsigned h, unsigned nbChannels) (cannyEdgeH.hpp:20).  The GL
  // buffer id is kept only so call sites compile; the bytes the reference copies into the PBO are served by view().
  CannyEdge(unsigned int pbo, unsigned int inputWidth, unsigned int inputHeight, unsigned int inputNbChannels, int device = 0, int maxBatch = 1)
      : m_pbo(pbo), m_w((int)inputWidth), m_h((int)inputHeight), m_ch((int)inputNbChannels)
  {
    const int rc = b2c_create(&m_h_, device, m_w, m_h, (int)inputNbChannels, maxBatch);
    if (rc != B2C_OK) throw b2c::Error(rc, "b2c_create");
    for (const auto &stage : cannyStages()) b2c::TimerManager::Get().createTimer(stage.second);   // cannyEdgeH.cu:35-37
  }
  CannyEdge(unsigned int w, unsigned int h) : CannyEdge(0u, w, h, 3u) {}
  ~CannyEdge() { b2c_destroy(m_h_); }
  CannyEdge(const CannyEdge &) = delete;
  CannyEdge &operator=(const CannyEdge &) = delete;

  // cannyEdgeH.cu:49-120
  template <class Mat> void run(const Mat &input, CannyStage finalStage = HYSTER)
  {
    const b2c::FrameView f = b2c::view_of(input);
    if (f.rows != m_h || f.cols != m_w) throw b2c::Error(B2C_ERR_SIZE, "CannyEdge::run");   // reference: logs and carries on (cannyEdgeH.cu:124-130)
    check(b2c_run(m_h_, f.data, f.step, (int)finalStage), "b2c_run");
    if (isKernelProfilingEnabled()) feedTimers(finalStage);
  }
  // The reference times every stage kernel with its own event pair and a host sync (cannyEdgeH.cu:409-430).  Here stages
  // 1..5 are ONE fused launch, so its time is booked on the first stage name and the stages 2..min(finalStage, 5) that
  // ran inside it get a 0 ms sample (their counters advance like the reference's); "6/6 Hysteresis" gets the on-device
  // hysteresis.  The UI's "total up to the selected stage" (imguiApp.cpp:363-376) therefore stays the true GPU time.
  void feedTimers(CannyStage finalStage)
  {
    const std::vector<float> t = lastTimings();
    auto &tm = b2c::TimerManager::Get();
    const auto &names = cannyStages();
    tm.addTime(names.at(MONO), t[1]);
    for (int s = GAUSSIAN; s <= (int)finalStage && s <= THRESH; ++s) tm.addTime(names.at((CannyStage)s), 0.0);
    if (finalStage == HYSTER) tm.addTime(names.at(HYSTER), t[2]);
  }
  // the GL-free half of _sendOutputToOpenGL (cannyEdgeH.cu:154-212): the stage picture of the last run, device to
  // device, into the mapped PBO pointer (pitch 0 = width, the layout of imguiApp.cpp:76)
  void copyViewTo(void *devDst, size_t pitch = 0, void *stream = nullptr) { check(b2c_copy_view(m_h_, devDst, pitch, stream), "b2c_copy_view"); }
  // n contiguous host frames -> n tightly packed u8 edge maps, through the pinned async pipeline
  void runBatch(const uint8_t *frames, size_t rowStride, int n, uint8_t *edgesOut, bool packedBits = false)
  {
    check(b2c_run_batch_host(m_h_, frames, rowStride, n, edgesOut, packedBits ? 1 : 0), "b2c_run_batch_host");
  }

  // cannyEdgeH.hpp:25-29
  void setLowThreshold(unsigned char low) { check(b2c_set_low_threshold(m_h_, low), "setLowThreshold"); }
  void setHighThreshold(unsigned char high) { check(b2c_set_high_threshold(m_h_, high), "setHighThreshold"); }
  unsigned char getLowThreshold() const { return (unsigned char)b2c_get_low_threshold(m_h_); }
  unsigned char getHighThreshold() const { return (unsigned char)b2c_get_high_threshold(m_h_); }
  // cannyEdgeH.hpp:31-32
  void enableKernelProfiling(bool enable) { check(b2c_enable_profiling(m_h_, enable ? 1 : 0), "enableKernelProfiling"); }
  bool isKernelProfilingEnabled() const { return b2c_is_profiling_enabled(m_h_) == 1; }
  // ms: upload, fused stencil, hysteresis, output, total; [5] = 1 (one on-device hysteresis pass)
  std::vector<float> lastTimings() const
  {
    std::vector<float> v(6, 0.0f);
    check(b2c_last_timings(m_h_, v.data(), 6), "b2c_last_timings");
    return v;
  }

  // accessors for the intermediate buffers (frame 0 of the last run), host copies
  std::vector<uint8_t> mono() const { return get8(B2C_BUF_MONO); }
  std::vector<uint8_t> blur() const { return get8(B2C_BUF_BLUR); }
  std::vector<uint8_t> nms() const { return get8(B2C_BUF_NMS); }
  std::vector<uint8_t> thresh() const { return get8(B2C_BUF_THRESH); }
  std::vector<uint8_t> edges() const { return get8(B2C_BUF_EDGES); }
  std::vector<uint8_t> view() const { return get8(B2C_BUF_VIEW); }   // what the reference's PBO would hold
  std::vector<float> gradient() const
  {
    std::vector<float> v((size_t)m_w * m_h);
    check(b2c_download(m_h_, B2C_BUF_GRAD, v.data(), (size_t)m_w * sizeof(float)), "b2c_download");
    return v;
  }
  // device pointers (no copy) for consumers that stay on the GPU
  const void *deviceBuffer(int bufferId, size_t *pitchBytes = nullptr, int *elemSize = nullptr) const
  {
    const void *p = nullptr;
    check(b2c_get_buffer(m_h_, bufferId, &p, pitchBytes, elemSize), "b2c_get_buffer");
    return p;
  }
  b2c_handle handle() const { return m_h_; }
  int width() const { return m_w; }
  int height() const { return m_h; }
  int channels() const { return m_ch; }

private:
  void check(int rc, const char *what) const
  {
    if (rc != B2C_OK) throw b2c::Error(rc, std::string(what) + (rc == B2C_ERR_CUDA ? std::string(" [") + b2c_last_cuda_error(m_h_) + "]" : std::string()));
  }
  std::vector<uint8_t> get8(int id) const
  {
    std::vector<uint8_t> v((size_t)m_w * m_h);
    check(b2c_download(m_h_, id, v.data(), (size_t)m_w), "b2c_download");
    return v;
  }
  unsigned int m_pbo;
  int m_w, m_h, m_ch;
  b2c_handle m_h_ = nullptr;
};
}// namespace cuda

// src/cvp/cvPipeline.hpp:20-39
class cvPipeline
{
public:
  cvPipeline(const unsigned int pbo, const unsigned int inputImageCols, const unsigned int inputImageRows, const int inputImageNbChannels)
      : m_cudaCannyEdge(new cuda::CannyEdge(pbo, inputImageCols, inputImageRows, (unsigned)inputImageNbChannels))
  {
  }
  ~cvPipeline() { delete m_cudaCannyEdge; }   // the reference leaks it (cvPipeline.cpp:14-17 calls release())
  cvPipeline(const cvPipeline &) = delete;
  cvPipeline &operator=(const cvPipeline &) = delete;

  // cvPipeline.cpp:19-41: false for a null implementation, an empty frame or a type other than 8-bit 1/3 channels.
  // The frame must have the channel count given to the constructor: 3 (BGR8), 1 (GRAY8 -- works here; upstream the
  // 1-channel upload is overwritten) or 4 (BGRA8, an addition).
  template <class Mat> bool process(const Mat &inputImage, CannyStage finalStage)
  {
    if (!m_cudaCannyEdge) return false;
    const b2c::FrameView f = b2c::view_of(inputImage);
    if (f.empty()) return false;
    if (!b2c::is_8bit(inputImage)) return false;   // "Only supporting CV_8UC3 and CV_8UC1" (cvPipeline.cpp:32-36; CV_8UC4 added)
    if (f.channels() != m_cudaCannyEdge->channels()) return false;
    m_cudaCannyEdge->run(f, finalStage);
    return true;
  }
  std::vector<uint8_t> output() const { return m_cudaCannyEdge->view(); }

  void setLowThreshold(unsigned char low) { if (m_cudaCannyEdge) m_cudaCannyEdge->setLowThreshold(low); }
  unsigned char getLowThreshold() const { return m_cudaCannyEdge ? m_cudaCannyEdge->getLowThreshold() : 0; }
  void setHighThreshold(unsigned char high) { if (m_cudaCannyEdge) m_cudaCannyEdge->setHighThreshold(high); }
  unsigned char getHighThreshold() const { return m_cudaCannyEdge ? m_cudaCannyEdge->getHighThreshold() : 255; }   // 255 when not ready, like cvPipeline.cpp:73-81
  void enableCudaProfiling(bool enable) { if (m_cudaCannyEdge) m_cudaCannyEdge->enableKernelProfiling(enable); }
  bool isCudaProfilingEnabled() const { return m_cudaCannyEdge && m_cudaCannyEdge->isKernelProfilingEnabled(); }
  cuda::CannyEdge *impl() { return m_cudaCannyEdge; }

private:
  cuda::CannyEdge *m_cudaCannyEdge;
};
}// namespace cvp

namespace b2c
{
// One row band of a height_global image on one GPU (BASELINE config 5).  No reference counterpart (the reference is
// single-GPU); host logic of the b2c_band_* ABI in C++: 4-row input halo exchange, stencil, band-local hysteresis, ONE
// exchange of seam records, solve.  Transports:
//   * collective: the caller supplies the halo exchange and the all-gather of the records as callbacks (MPI, NCCL, a
//     memcpy between bands of one process ...) working on DEVICE pointers;
//   * peer memory (ranks of one box): wirePeers() / wireLocal() map the other bands' buffers, then halo rows and records
//     travel as stores over NVLink hidden behind the stencil and the resolve pass.
class BandRunner
{
public:
  // all-gather: `bytes` from send (device) of every band, in band order, into recv (device, world * bytes)
  using AllGather = std::function<void(const void *sendDev, void *recvDev, size_t bytes)>;
  // halo exchange: sendUp goes to band rank-1 (arrives as its recvDown), sendDown to rank+1 (its recvUp); null = no neighbour
  using HaloExchange = std::function<void(const void *sendUp, void *recvUp, const void *sendDown, void *recvDown, size_t bytes)>;

  static void bandRows(int heightGlobal, int world, int rank, int *y0, int *rows)
  {
    const int base = heightGlobal / world, extra = heightGlobal % world;
    *rows = base + (rank < extra ? 1 : 0);
    *y0 = rank * base + (rank < extra ? rank : extra);
  }

  BandRunner(int device, int width, int heightGlobal, int world, int rank) : m_w(width), m_world(world), m_rank(rank)
  {
    bandRows(heightGlobal, world, rank, &m_y0, &m_rows);
    check(b2c_create_band(&m_h, device, width, m_rows, m_y0, heightGlobal), "b2c_create_band");
    void *p = nullptr;
    check(b2c_band_input(m_h, &p, &m_stride), "b2c_band_input");
    m_in = static_cast<uint8_t *>(p);
    check(b2c_band_seam_bytes(m_h, &m_seamBytes), "b2c_band_seam_bytes");
  }
  ~BandRunner()
  {
    if (m_all) b2c_dev_free(m_h, m_all);
    b2c_destroy(m_h);
  }
  BandRunner(const BandRunner &) = delete;
  BandRunner &operator=(const BandRunner &) = delete;

  int y0() const { return m_y0; }
  int rows() const { return m_rows; }
  size_t rowStride() const { return m_stride; }
  size_t seamBytes() const { return m_seamBytes; }
  b2c_handle handle() const { return m_h; }
  // device address of buffer row r (rows 0..3: halo above, 4..4+rows-1: the band, then 4 halo rows)
  uint8_t *inputRow(int r) const { return m_in + (size_t)r * m_stride; }
  // the band's pixels from host memory (rows hostStride bytes apart)
  void upload(const uint8_t *hostBand, size_t hostStride)
  {
    if (hostStride == m_stride) { check(b2c_dev_upload(m_h, inputRow(4), hostBand, m_stride * m_rows), "b2c_dev_upload"); return; }
    for (int r = 0; r < m_rows; ++r) check(b2c_dev_upload(m_h, inputRow(4 + r), hostBand + (size_t)r * hostStride, (size_t)m_w * 3), "b2c_dev_upload");
  }
  // peer memory between processes: exportBlob() of every rank, all-gathered by the caller in rank order, then wirePeers()
  std::vector<uint8_t> exportBlob()
  {
    std::vector<uint8_t> b(144);
    check(b2c_band_p2p_export(m_h, b.data()), "b2c_band_p2p_export");
    return b;
  }
  void wirePeers(const void *allBlobs)
  {
    check(b2c_band_p2p_open(m_h, allBlobs, m_world, m_rank), "b2c_band_p2p_open");
    m_p2p = true;
  }
  // peer memory between the bands of ONE process
  static void wireLocal(const std::vector<BandRunner *> &bands)
  {
    std::vector<b2c_handle> hs;
    for (auto *b : bands) hs.push_back(b->m_h);
    for (size_t r = 0; r < bands.size(); ++r) {
      bands[r]->check(b2c_band_p2p_open_local(bands[r]->m_h, hs.data(), (int)bands.size(), (int)r), "b2c_band_p2p_open_local");
      bands[r]->m_p2p = true;
    }
  }

  // one band per process.  Returns the number of cross-band exchanges (0 or 1).  Asynchronous on `stream`, except that the
  // callbacks of the collective transport are called in between (they see device memory that is ready once `stream`
  // has drained: sync() is called before them).
  int run(void *stream = nullptr, const HaloExchange &halo = nullptr, const AllGather &gather = nullptr)
  {
    if (m_world == 1) {
      check(b2c_band_stencil(m_h, inputRow(4), m_stride, stream), "b2c_band_stencil");
      check(b2c_band_hysteresis(m_h, stream), "b2c_band_hysteresis");
      return 0;
    }
    if (m_p2p) {
      check(b2c_band_p2p_stencil(m_h, stream, B2C_P2P_ALL), "b2c_band_p2p_stencil");
      check(b2c_band_hysteresis(m_h, stream), "b2c_band_hysteresis");
      check(b2c_band_p2p_seam(m_h, stream), "b2c_band_p2p_seam");
      return 1;
    }
    if (!halo || !gather) throw Error(B2C_ERR_INVALID, "BandRunner::run: no transport (wire peers or pass the callbacks)");
    const bool up = m_rank > 0, dn = m_rank + 1 < m_world;
    halo(up ? inputRow(4) : nullptr, up ? inputRow(0) : nullptr, dn ? inputRow(m_rows) : nullptr, dn ? inputRow(4 + m_rows) : nullptr, 4 * m_stride);
    check(b2c_band_stencil(m_h, inputRow(4), m_stride, stream), "b2c_band_stencil");
    check(b2c_band_hysteresis(m_h, stream), "b2c_band_hysteresis");
    void *rec = nullptr;
    check(b2c_band_seam_record(m_h, &rec), "b2c_band_seam_record");
    if (!m_all) check(b2c_dev_alloc(m_h, m_seamBytes * m_world, &m_all), "b2c_dev_alloc");
    sync();
    gather(rec, m_all, m_seamBytes);
    check(b2c_band_seam_solve(m_h, m_all, m_world, m_rank, stream), "b2c_band_seam_solve");
    return 1;
  }
  // all bands of one process wired with wireLocal(): every band on its own stream, every store issued before the
  // first device-side wait for it
  static int runLocal(const std::vector<BandRunner *> &bands)
  {
    if (bands.size() == 1) return bands[0]->run();
    for (auto *b : bands) b->check(b2c_band_p2p_stencil(b->m_h, nullptr, B2C_P2P_PUSH), "b2c_band_p2p_stencil");
    for (auto *b : bands) b->check(b2c_band_p2p_stencil(b->m_h, nullptr, B2C_P2P_WAIT), "b2c_band_p2p_stencil");
    for (auto *b : bands) b->check(b2c_band_hysteresis(b->m_h, nullptr), "b2c_band_hysteresis");
    for (auto *b : bands) b->check(b2c_band_p2p_seam(b->m_h, nullptr), "b2c_band_p2p_seam");
    return 1;
  }
  void sync() { check(b2c_sync(m_h), "b2c_sync"); }
  // promoted weak runs of the last solve; throws if a peer never arrived
  int status()
  {
    int n = 0, err = 0;
    check(b2c_band_status(m_h, &n, &err), "b2c_band_status");
    if (err) throw Error(B2C_ERR_STATE, "BandRunner: a peer did not arrive within 2 s");
    return n;
  }
  // the band's u8 edge map (rows x width, tight) to host memory; blocking
  void edges(uint8_t *hostOut)
  {
    sync();
    check(b2c_download(m_h, B2C_BUF_EDGES, hostOut, (size_t)m_w), "b2c_download");
  }

private:
  void check(int rc, const char *what) const
  {
    if (rc != B2C_OK) throw Error(rc, std::string(what) + (rc == B2C_ERR_CUDA ? std::string(" [") + b2c_last_cuda_error(m_h) + "]" : std::string()));
  }
  b2c_handle m_h = nullptr;
  int m_w, m_world, m_rank, m_y0 = 0, m_rows = 0;
  size_t m_stride = 0, m_seamBytes = 0;
  uint8_t *m_in = nullptr;
  void *m_all = nullptr;
  bool m_p2p = false;
};
}// namespace b2c
