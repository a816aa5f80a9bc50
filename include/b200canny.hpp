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
// No CUDA, OpenCV or GL header is needed to include this file.  Frames are passed either as b2c::FrameView or as
// anything shaped like cv::Mat (members data, rows, cols, step, channels(), empty()), so `process(cv::Mat, stage)`
// compiles unchanged when OpenCV is present.  Errors: the reference exits the process on any CUDA error
// (helper.hpp:4-17); here the classes throw b2c::Error carrying the C ABI status.
#pragma once
#include <cstddef>
#include <cstdint>
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
  }
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
  // ms: upload, fused stencil, hysteresis, output, total; [5] = hysteresis rounds (replaces the timerManager sink, cannyEdgeH.cu:415-430)
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
