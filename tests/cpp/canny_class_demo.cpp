// canny_class_demo.cpp -- TEST: C++ host code written against the reference's class surface (cvp::cvPipeline /
// cvp::cuda::CannyEdge as rebuilt by include/b200canny.hpp) -- the shape of src/imgui/imguiApp.cpp:102,328-348,515.
// usage: canny_class_demo <kind> <seed> <w> <h> <low> <high> <out_prefix>
// Writes <out_prefix>.edges / .blur / .nms / .grad (raw) for the Python test to compare with the oracle.
// Prints the b2c::TimerManager table (the reference's profiling widget, imguiApp.cpp:357-376) as "timer <name> <count>
// <average ms>" lines and the last run's phase times as "timings ..." for the a14 test.
// Without a CUDA device the constructor throws (no CPU fallback): exit code 3.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "b200canny.hpp"

// What cvPipeline::process sees of a cv::Mat (cvPipeline.cpp:19-41): data / rows / cols / step / channels() / type() / empty()
struct FakeMat {
  uint8_t *data;
  int rows, cols;
  size_t step;
  int cvType;   // CV_MAKETYPE(depth, channels): depth | (channels - 1) << 3
  int type() const { return cvType; }
  int channels() const { return (cvType >> 3) + 1; }
  bool empty() const { return !data || rows * cols == 0; }
};

static void dump(const std::string &path, const void *p, size_t n)
{
  FILE *f = fopen(path.c_str(), "wb");
  if (!f || fwrite(p, 1, n, f) != n) { perror(path.c_str()); exit(2); }
  fclose(f);
}

int main(int argc, char **argv)
{
  if (argc < 8) return 1;
  const int kind = atoi(argv[1]);
  const uint64_t seed = strtoull(argv[2], nullptr, 0);
  const int w = atoi(argv[3]), h = atoi(argv[4]), lo = atoi(argv[5]), hi = atoi(argv[6]);
  const std::string out = argv[7];
  // a strided frame, like a cv::Mat ROI
  const size_t step = (size_t)w * 3 + 40;
  std::vector<uint8_t> buf(step * h);
  if (b2c_synth_frame(kind, seed, w, h, buf.data(), step) != B2C_OK) return 1;
  b2c::FrameView frame;
  frame.data = buf.data(); frame.rows = h; frame.cols = w; frame.step = step; frame.nbChannels = 3;
  try {
    cvp::cvPipeline pipe(/*pbo*/ 0u, (unsigned)w, (unsigned)h, 3);
    pipe.setHighThreshold((unsigned char)hi);
    pipe.setLowThreshold((unsigned char)lo);
    if (pipe.getLowThreshold() != lo || pipe.getHighThreshold() != hi) return 4;
    b2c::FrameView empty;
    if (pipe.process(empty, cvp::HYSTER)) return 5;            // blank frame -> false (cvPipeline.cpp:27-31)
    // a cv::Mat-shaped frame: CV_8UC3 = 16 passes; CV_32FC3 = 21, CV_16UC3 = 18 are refused like cvPipeline.cpp:32-36
    FakeMat mat{ buf.data(), h, w, step, 16 };
    FakeMat matf = mat, mat16 = mat;
    matf.cvType = 21;
    mat16.cvType = 18;
    if (pipe.process(matf, cvp::HYSTER) || pipe.process(mat16, cvp::HYSTER)) return 7;
    b2c::TimerManager::Get().reset();
    if (!pipe.process(mat, cvp::HYSTER)) return 6;
    if (!pipe.process(frame, cvp::HYSTER)) return 6;
    const auto edges = pipe.output();
    {   // the PBO copy without GL: device-to-device into a caller-owned buffer with its own pitch
      auto *ce = pipe.impl();
      void *dev = nullptr;
      const size_t pitch = (size_t)w + 24;
      if (b2c_dev_alloc(ce->handle(), pitch * h, &dev) != B2C_OK) return 8;
      ce->copyViewTo(dev, pitch);
      b2c_sync(ce->handle());
      std::vector<uint8_t> back(pitch * h);
      if (b2c_dev_download(ce->handle(), back.data(), dev, pitch * h) != B2C_OK) return 8;
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
          if (back[y * pitch + x] != edges[(size_t)y * w + x]) return 9;
      b2c_dev_free(ce->handle(), dev);
    }
    dump(out + ".edges", edges.data(), edges.size());
    auto *ce = pipe.impl();
    const auto blur = ce->blur(), nms = ce->nms();
    const auto grad = ce->gradient();
    dump(out + ".blur", blur.data(), blur.size());
    dump(out + ".nms", nms.data(), nms.size());
    dump(out + ".grad", grad.data(), grad.size() * sizeof(float));
    if (!pipe.process(frame, cvp::GRADIENT)) return 6;           // stage select: saturated gradient view (cannyEdgeH.cu:181-186)
    const auto gview = pipe.output();
    dump(out + ".gview", gview.data(), gview.size());
    const auto t = ce->lastTimings();
    printf("ok %dx%d total %.3f ms\n", w, h, t[4]);
    printf("timings upload %.4f stencil %.4f hysteresis %.4f output %.4f total %.4f\n", t[0], t[1], t[2], t[3], t[4]);
    const auto &tm = b2c::TimerManager::Get();
    for (auto it = tm.beginTimerList(); it != tm.endTimerList(); ++it) printf("timer %s | %zu | %.5f\n", it->first.c_str(), it->second.nbCount, it->second.averageTime());
    printf("avg_hyster %.5f\n", tm.getAverageTime(cvp::cannyStages().at(cvp::HYSTER)));
  } catch (const b2c::Error &e) {
    fprintf(stderr, "b2c::Error %d: %s\n", e.status, e.what());
    return 3;
  }
  return 0;
}
