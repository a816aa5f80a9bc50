// canny_class_demo.cpp -- TEST: C++ host code written against the reference's class surface (cvp::cvPipeline /
// cvp::cuda::CannyEdge as rebuilt by include/b200canny.hpp) -- the shape of src/imgui/imguiApp.cpp:102,328-348,515.
// usage: canny_class_demo <kind> <seed> <w> <h> <low> <high> <out_prefix>
// Writes <out_prefix>.edges / .blur / .nms / .grad (raw) for the Python test to compare with the oracle.
// Without a CUDA device the constructor throws (no CPU fallback): exit code 3.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "b200canny.hpp"

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
    if (!pipe.process(frame, cvp::HYSTER)) return 6;
    const auto edges = pipe.output();
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
  } catch (const b2c::Error &e) {
    fprintf(stderr, "b2c::Error %d: %s\n", e.status, e.what());
    return 3;
  }
  return 0;
}
