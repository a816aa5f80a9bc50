// band_runner_demo.cpp -- TEST: one image in row bands through b2c::BandRunner (include/b200canny.hpp), all bands on
// device 0 of this process: (a) collective transport with host-staged callbacks, (b) peer-memory transport
// (BandRunner::wireLocal + runLocal).  Compares both with the unsharded cvp::cuda::CannyEdge run and writes the
// assembled edge map for the Python test (oracle comparison).
// usage: band_runner_demo <kind> <seed> <w> <h> <bands> <out_prefix>      exit 3 = no device (no CPU fallback)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>

#include "b200canny.hpp"

int main(int argc, char **argv)
{
  if (argc < 7) return 1;
  const int kind = atoi(argv[1]);
  const uint64_t seed = strtoull(argv[2], nullptr, 0);
  const int w = atoi(argv[3]), h = atoi(argv[4]), nb = atoi(argv[5]);
  const std::string out = argv[6];
  const size_t step = (size_t)w * 3;
  std::vector<uint8_t> img(step * h);
  if (b2c_synth_frame(kind, seed, w, h, img.data(), step) != B2C_OK) return 1;
  try {
    std::vector<uint8_t> want;
    {
      cvp::cuda::CannyEdge whole((unsigned)w, (unsigned)h);
      b2c::FrameView f;
      f.data = img.data(); f.rows = h; f.cols = w; f.step = step;
      whole.run(f);
      want = whole.edges();
    }
    std::vector<std::unique_ptr<b2c::BandRunner>> own;
    std::vector<b2c::BandRunner *> bands;
    for (int r = 0; r < nb; ++r) {
      own.emplace_back(new b2c::BandRunner(0, w, h, nb, r));
      bands.push_back(own.back().get());
      bands[r]->upload(img.data() + (size_t)bands[r]->y0() * step, step);
    }
    auto assemble = [&](std::vector<uint8_t> &got) {
      got.assign((size_t)w * h, 0);
      for (auto *b : bands) b->edges(got.data() + (size_t)b->y0() * w);
    };
    // (a) collective transport, driven band by band the way one rank per band would: the callbacks move the bytes
    // through host memory with the C ABI's own copies.  All bands' stencil + hysteresis first, then the exchange.
    const size_t hb = 4 * bands[0]->rowStride();
    std::vector<uint8_t> tmp(hb);
    for (int r = 0; r + 1 < nb; ++r) {   // halo rows across seam r | r+1
      b2c_dev_download(bands[r]->handle(), tmp.data(), bands[r]->inputRow(bands[r]->rows()), hb);
      b2c_dev_upload(bands[r + 1]->handle(), bands[r + 1]->inputRow(0), tmp.data(), hb);
      b2c_dev_download(bands[r + 1]->handle(), tmp.data(), bands[r + 1]->inputRow(4), hb);
      b2c_dev_upload(bands[r]->handle(), bands[r]->inputRow(4 + bands[r]->rows()), tmp.data(), hb);
    }
    const size_t sb = bands[0]->seamBytes();
    std::vector<uint8_t> all(sb * nb);
    b2c::BandRunner::HaloExchange halo = [](const void *, void *, const void *, void *, size_t) {};   // done above
    // every band's run() reaches its gather callback after its record is ready; with one thread per band the callback
    // would be an MPI_Allgather -- here the bands run one after the other, so the records are collected in a first
    // pass and the solve happens in a second one
    for (int r = 0; r < nb; ++r) {
      b2c_band_stencil(bands[r]->handle(), bands[r]->inputRow(4), bands[r]->rowStride(), nullptr);
      b2c_band_hysteresis(bands[r]->handle(), nullptr);
      void *rec = nullptr;
      b2c_band_seam_record(bands[r]->handle(), &rec);
      bands[r]->sync();
      b2c_dev_download(bands[r]->handle(), all.data() + sb * r, rec, sb);
    }
    b2c::BandRunner::AllGather gather = [&](const void *, void *recvDev, size_t bytes) {
      if (bytes != sb) abort();
      for (auto *b : bands)
        if (b2c_dev_upload(b->handle(), recvDev, all.data(), sb * nb) == B2C_OK) break;
    };
    int ex = 0;
    for (int r = 0; r < nb; ++r) ex += bands[r]->run(nullptr, halo, gather);   // (recomputes the band; the records are identical)
    std::vector<uint8_t> got;
    assemble(got);
    if (got != want) { fprintf(stderr, "collective transport differs from the unsharded run\n"); return 4; }
    if (nb > 1 && ex != nb) return 5;
    // (b) peer memory
    if (nb > 1) {
      b2c::BandRunner::wireLocal(bands);
      for (int rep = 0; rep < 2; ++rep) {
        b2c::BandRunner::runLocal(bands);
        for (auto *b : bands) b->status();
        assemble(got);
        if (got != want) { fprintf(stderr, "peer-memory transport differs from the unsharded run (rep %d)\n", rep); return 6; }
      }
    }
    FILE *f = fopen((out + ".edges").c_str(), "wb");
    if (!f || fwrite(got.data(), 1, got.size(), f) != got.size()) return 2;
    fclose(f);
    printf("ok %dx%d in %d bands\n", w, h, nb);
  } catch (const b2c::Error &e) {
    fprintf(stderr, "b2c::Error %d: %s\n", e.status, e.what());
    return 3;
  }
  return 0;
}
