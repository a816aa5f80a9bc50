/*
 * b200canny.h -- C ABI of the B200-native Canny path (libb200canny.so).
 *
 * This is the drop-in boundary for the hot path of axoloto/CudaCam's src/cvp: everything the
 * reference's host class cvp::cuda::CannyEdge does around its kernels is reachable through these
 * plain-C entry points (no CUDA, OpenCV, GL or torch types in any signature).  Each entry cites the
 * reference interface it replaces (paths relative to the reference tree).  The header-only C++ class
 * in b200canny.hpp rebuilds the reference's class surface (cvp::cuda::CannyEdge / cvp::cvPipeline)
 * on top of it; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * All functions return B2C_OK (0) or a negative status; nothing ever calls exit() (the reference's
 * checkCudaErrors does: src/cvp/helper.hpp:4-17).  There is no CPU fallback: without a CUDA device
 * b2c_create() fails with B2C_ERR_CUDA.
 */
#ifndef B200CANNY_H
#define B200CANNY_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B2C_API __attribute__((visibility("default")))
#else
#define B2C_API
#endif

typedef struct b2c_ctx *b2c_handle;

/* status codes */
enum {
  B2C_OK = 0,
  B2C_ERR_INVALID = -1,     /* bad argument (null handle, stage out of range, n > max_batch ...)        */
  B2C_ERR_CUDA = -2,        /* a CUDA runtime call or a kernel launch failed; see b2c_last_cuda_error() */
  B2C_ERR_NOMEM = -3,
  B2C_ERR_SIZE = -4,        /* frame geometry differs from the one given at creation
                               (reference: logs and carries on with stale data, cannyEdgeH.cu:124-130)  */
  B2C_ERR_UNSUPPORTED = -5, /* e.g. channels not in {1, 3, 4}, or an option that a mode does not support          */
  B2C_ERR_STATE = -6        /* accessor used before any frame was run                                   */
};

/* stage ids == enum cvp::CannyStage (src/cvp/define.hpp:9-17) */
enum { B2C_STAGE_MONO = 0, B2C_STAGE_GAUSSIAN = 1, B2C_STAGE_GRADIENT = 2, B2C_STAGE_NMS = 3, B2C_STAGE_THRESH = 4, B2C_STAGE_HYSTER = 5 };

/* buffer ids for b2c_get_buffer / b2c_download (device buffers owned by the handle) */
enum {
  B2C_BUF_MONO = 0,   /* u8  w*h   == d_mono   (cannyEdgeH.hpp:58)                         */
  B2C_BUF_BLUR = 1,   /* u8  w*h   == d_blurr  (:61)                                       */
  B2C_BUF_GRAD = 2,   /* f32 w*h   == d_grad   (:70)                                       */
  B2C_BUF_NMS = 3,    /* u8  w*h   == d_nms    (:76)                                       */
  B2C_BUF_THRESH = 4, /* u8  w*h   == d_thresh (:79) {0,128,255}                           */
  B2C_BUF_EDGES = 5,  /* u8  w*h   == d_hyster after removeCandidates (:84) {0,255}        */
  B2C_BUF_MAP2 = 6,   /* u32 per 16 px: bits 0-15 strong, 16-31 weak (view of the 2-bit map) */
  B2C_BUF_BITS = 7,   /* u32 per 32 px: final edges, 1 bit per pixel                       */
  B2C_BUF_VIEW = 8    /* u8  w*h tight: what the reference copies into its GL PBO
                         (cannyEdgeH.cu:154-212) for the stage passed to the last b2c_run  */
};

/* ---- lifetime: cvp::cuda::CannyEdge::CannyEdge / ~CannyEdge (cannyEdgeH.cu:16-47), _initAlloc/_endAlloc (:340-407)
 * channels = bytes per pixel of the input frames: 3 = BGR8 (what the reference processes), 4 = BGRA8 (alpha ignored),
 * 1 = GRAY8 (the gray value is the byte: what the reference's CV_8UC1 path evidently meant to do -- upstream it
 * uploads the frame to d_mono and then overwrites it with rgb2mono of a stale d_rgb, cannyEdgeH.cu:140-146 + :60-64);
 * NV12 surfaces use it on their luma plane), B2C_PLANAR_BGR8 = three planes B, G, R of `height` rows each instead of
 * interleaved pixels (row_stride = bytes per plane row, frame = 3 * height rows; served by the staged tile kernel, not
 * by the fast marching kernel). */
#define B2C_PLANAR_BGR8 0x103
B2C_API int b2c_create(b2c_handle *out, int device, int width, int height, int channels, int max_batch);
B2C_API void b2c_destroy(b2c_handle h);

/* ---- thresholds: setLow/HighThreshold, getLow/HighThreshold (cannyEdgeH.hpp:25-29), same clamping */
B2C_API int b2c_set_low_threshold(b2c_handle h, uint8_t low);
B2C_API int b2c_set_high_threshold(b2c_handle h, uint8_t high);
B2C_API int b2c_get_low_threshold(b2c_handle h);
B2C_API int b2c_get_high_threshold(b2c_handle h);

/* ---- profiling toggle: enableKernelProfiling / isKernelProfilingEnabled (cannyEdgeH.hpp:31-32).
 * Timings are recorded with events and only read back by b2c_last_timings (no sync in the hot path,
 * unlike cannyEdgeH.cu:415-430).  ms[0]=upload, [1]=fused stencil, [2]=hysteresis, [3]=output, [4]=total,
 * [5] = on-device hysteresis passes of that run (1: the union-find needs no rounds; the reference needs 16-30 launches). */
B2C_API int b2c_enable_profiling(b2c_handle h, int on);
B2C_API int b2c_is_profiling_enabled(b2c_handle h);
B2C_API int b2c_last_timings(b2c_handle h, float *ms, int n);

/* ---- one frame from host memory: CannyEdge::run(cv::Mat, CannyStage) (cannyEdgeH.cu:49-120) incl.
 * _loadInputImage (:122-152) and _sendOutputToOpenGL (:154-212; the view buffer stands in for the PBO).
 * Blocking, like the reference.  bgr = interleaved BGR8, row_stride in bytes (cv::Mat::step). */
B2C_API int b2c_run(b2c_handle h, const uint8_t *host_bgr, size_t row_stride, int final_stage);

/* ---- batch of device-resident frames, asynchronous on `stream` (a cudaStream_t passed as void*, 0 =
 * default stream).  Fused stencil + on-device hysteresis; edges (u8 {0,255}) are written to dev_edges
 * if non-null, else to the handle's own edge buffer.  n <= max_batch.  No reference counterpart (the
 * reference is one frame at a time, cannyEdgeH.cu:49); this is the throughput entry of BASELINE config 2/4. */
B2C_API int b2c_run_device(b2c_handle h, const uint8_t *dev_bgr, size_t row_stride, size_t frame_stride, int n,
                           uint8_t *dev_edges, size_t edges_pitch, size_t edges_frame_stride, void *stream);

/* the two halves of b2c_run_device, separately (bench / profiling of the fused stencil alone) */
B2C_API int b2c_stencil_device(b2c_handle h, const uint8_t *dev_bgr, size_t row_stride, size_t frame_stride, int n, void *stream);
B2C_API int b2c_hysteresis_device(b2c_handle h, int n, uint8_t *dev_edges, size_t edges_pitch, size_t edges_frame_stride, void *stream);
/* hysteresis on a map produced elsewhere: a thresholded image (0 / 128 / 255 = the reference's d_threshImage,
 * cannyEdgeD.cu:274-292) from HOST memory becomes frame 0's state in place of a stencil run; follow with
 * b2c_hysteresis_device(h, 1, ...) or, on a band handle, b2c_band_hysteresis.  Blocking. */
B2C_API int b2c_load_thresh(b2c_handle h, const uint8_t *host_thresh, size_t row_stride);

/* ---- batch of host frames through the pinned async pipeline: the handle's batch buffers are cut into up to 8 slots of
 * at most 8 frames; upload, kernels and download of successive chunks overlap on three streams, so only the first
 * small upload and the last small download are exposed.  frames = n contiguous frames of row_stride*height bytes;
 * edges_out receives n tightly packed w*h maps (or, with packed_bits != 0, n bit maps of
 * ceil(w/32)*4 bytes per row).  Blocking until the last frame is back.  Replaces the per-frame
 * blocking upload + PBO copy of cannyEdgeH.cu:122-212 for streams of frames. */
B2C_API int b2c_run_batch_host(b2c_handle h, const uint8_t *frames, size_t row_stride, int n, uint8_t *edges_out, int packed_bits);

/* ---- accessors for the intermediate buffers (the reference exposes them only through the finalStage
 * switch of run(), cannyEdgeH.cu:169-207).  Valid for frame 0 of the last run; stage buffers that the
 * fused path did not write are produced on demand from the retained input. */
B2C_API int b2c_get_buffer(b2c_handle h, int buffer_id, const void **dev_ptr, size_t *pitch_bytes, int *elem_size);
B2C_API int b2c_download(b2c_handle h, int buffer_id, void *host, size_t host_pitch_bytes);
/* The GL-free half of _sendOutputToOpenGL (cannyEdgeH.cu:154-212): the u8 picture of the stage the last run stopped at
 * (GRADIENT: the saturated float2uchar view) copied device-to-device into dev_dst, rows dst_pitch bytes apart (0 = width,
 * the PBO layout of imguiApp.cpp:76).  dev_dst is what cudaGraphicsResourceGetMappedPointer gave the caller for its PBO.
 * Asynchronous on `stream` (0 = the handle's compute stream). */
B2C_API int b2c_copy_view(b2c_handle h, void *dev_dst, size_t dst_pitch, void *stream);

/* ---- device memory helpers so that non-CUDA hosts can keep frames resident */
B2C_API int b2c_dev_alloc(b2c_handle h, size_t bytes, void **dev_ptr);
B2C_API int b2c_dev_free(b2c_handle h, void *dev_ptr);
B2C_API int b2c_dev_upload(b2c_handle h, void *dev_dst, const void *host_src, size_t bytes);
B2C_API int b2c_dev_download(b2c_handle h, void *host_dst, const void *dev_src, size_t bytes);
B2C_API int b2c_sync(b2c_handle h);
/* pinned host memory for the frame ring of a caller (b2c_run_batch_host copies straight out of / into it) */
B2C_API int b2c_host_alloc(size_t bytes, void **host_ptr);
B2C_API int b2c_host_free(void *host_ptr);
/* the handle's compute stream as a cudaStream_t (for callers that enqueue their own work behind a run) */
B2C_API void *b2c_stream(b2c_handle h);

/* ---- row-band mode for one image split over several GPUs (BASELINE config 5; no reference counterpart).
 * The band covers global rows [y0, y0+band_rows) of a height_global image (a band with a neighbour has >= 4 rows).
 * Input passed to b2c_band_stencil points at the band's first row inside a buffer that also holds the 4 rows above it
 * (unless the band starts at global row 0) and the 4 rows below it (unless it ends at the last global row): the
 * stencil needs 2 (Gaussian) + 1 (Sobel) + 1 (NMS) neighbour rows, and uses the reference's zero padding only outside
 * the global image.  Per image: [halo exchange] -> b2c_band_stencil -> b2c_band_hysteresis (band-local fixpoint, planes
 * and union-find forest are kept, the band's seam record is built) -> ONE exchange of seam records -> solve; the result
 * equals the unsharded run bit for bit.  All calls are asynchronous on `stream`. */
B2C_API int b2c_create_band(b2c_handle *out, int device, int width, int band_rows, int y0, int height_global);
B2C_API int b2c_band_stencil(b2c_handle h, const uint8_t *dev_bgr_band_row0, size_t row_stride, void *stream);
B2C_API int b2c_band_hysteresis(b2c_handle h, void *stream);
/* Cross-band hysteresis in one step.  b2c_band_hysteresis leaves the band's SEAM RECORD (b2c_band_seam_bytes bytes: the
 * edge and unresolved-weak bit rows of its first and last row + a component label per unresolved run) in device memory;
 * the caller all-gathers the records of all bands in band order (any transport: NCCL, MPI, a memcpy), and every band
 * solves the same small connected-components problem over all seams and promotes its own components that reach an edge
 * pixel of any band (only the plane words that were still unresolved are visited again).
 * b2c_band_seam_record: *record_dev = device address of this band's record (owned by the handle, rewritten by every
 * b2c_band_hysteresis on its stream);
 * b2c_band_seam_solve: all_records_dev = world records, b2c_band_seam_bytes apart. */
B2C_API int b2c_band_seam_bytes(b2c_handle h, size_t *bytes);
B2C_API int b2c_band_seam_record(b2c_handle h, void **record_dev);
B2C_API int b2c_band_seam_solve(b2c_handle h, const void *all_records_dev, int world, int rank, void *stream);
/* blocking (synchronises the device): weak runs of this band promoted by the last solve, and whether a peer-to-peer
 * wait timed out */
B2C_API int b2c_band_status(b2c_handle h, int *promoted_runs, int *error);

/* Peer-to-peer transport for ranks of ONE box: every rank maps the other ranks' mailboxes and band input buffers
 * (CUDA IPC between processes: export a 144-byte blob, all-gather the blobs, open; or b2c_band_p2p_open_local for
 * bands of one process) and the halo rows and seam records travel as ordinary stores over NVLink, with flag words
 * instead of collectives, hidden behind the band's own work:
 * b2c_band_input: the band's input buffer owned by the handle (so that it can be shared): rows 0..3 = halo rows above
 *   the band, rows 4..4+band_rows-1 = the band, then 4 halo rows; rows are row_stride bytes apart;
 * b2c_band_p2p_stencil: stencil of the band in that buffer: my first / last 4 rows -> the neighbours' buffers, then one
 *   stencil launch whose thread blocks next to a seam wait on the device for the neighbour's rows when they need them;
 * b2c_band_hysteresis (on a wired handle): also stores the seam record into every rank's mailbox before it resolves;
 * b2c_band_p2p_seam: waits for all records, solves, promotes (replaces the gather + b2c_band_seam_solve above).
 * phase: B2C_P2P_ALL for a rank that drives one band.  A process that drives SEVERAL bands (b2c_band_p2p_open_local)
 * gives every band its own stream and issues, band by band, b2c_band_p2p_stencil(PUSH = the stores), then (WAIT = the
 * stencil), then
 * b2c_band_hysteresis, then b2c_band_p2p_seam: the waits spin on the device and must never be queued ahead of the
 * stores they wait for.  Every wait has a 2 s time-out (b2c_band_status reports it once).  The stores run on a side stream
 * of the handle; work enqueued on `stream` after these calls is ordered behind them. */
enum { B2C_P2P_ALL = 0, B2C_P2P_PUSH = 1, B2C_P2P_WAIT = 2 };
B2C_API int b2c_band_input(b2c_handle h, void **dev_ptr, size_t *row_stride);
B2C_API int b2c_band_p2p_export(b2c_handle h, void *blob_144);
B2C_API int b2c_band_p2p_open(b2c_handle h, const void *all_blobs, int world, int rank);
B2C_API int b2c_band_p2p_open_local(b2c_handle h, const b2c_handle *all_handles, int world, int rank);
B2C_API int b2c_band_p2p_stencil(b2c_handle h, void *stream, int phase);
B2C_API int b2c_band_p2p_seam(b2c_handle h, void *stream);

/* ---- misc */
B2C_API const char *b2c_strerror(int status);
B2C_API const char *b2c_last_cuda_error(b2c_handle h);
B2C_API const char *b2c_version(void);
B2C_API int b2c_device_count(void);
/* Host placement for one process per GPU on multi-socket boxes: binds the calling thread to the CPUs of the NUMA node
 * the GPU hangs on (Linux sysfs), so that pinned rings allocated afterwards and the staging copies are local to the
 * GPU's PCIe root.  Returns the node, or -1 if unknown (nothing changed). */
B2C_API int b2c_bind_host_to_device(int device);
/* kernels launched by this handle since creation (bench.py reports it as gpu_launches) */
B2C_API long long b2c_launch_count(b2c_handle h);
/* options: "stencil_impl" 0 = marching kernel (default), 1 = staged tile kernel (the all-stages path of the
 * accessors); "march_rb" rows per band of the marching kernel (0 = automatic); "hyst_phase_timing"; "uf_spread" (warps per tile the
 * hysteresis work items are dealt to: 1, 2, 4, 8; 0 = chosen by batch size);
 * "seam_force_global" (tests: the large-seam code paths of the seam kernels) */
B2C_API int b2c_set_option(b2c_handle h, const char *name, int value);
/* read-only facts: "sm_count", "stencil_impl", "march_ctas_per_sm", "march_band_rows", "in_row_stride",
 * "plane_pitch_words", "map_pitch_words", "hyst_phase_us0..2", "seam_phase_us0..5" and "band_stencil_us0..2"
 * (band handles) */
B2C_API int b2c_get_info(b2c_handle h, const char *name);

/* ---- deterministic synthetic frames (host side; identical to cudacam_b200/synth.py).
 * kind: 0 "scene", 1 "noise", 2 "steps".  Writes w*h BGR8 pixels with the given row stride. */
B2C_API int b2c_synth_frame(int kind, uint64_t seed, int w, int h, uint8_t *out, size_t row_stride);

#ifdef __cplusplus
}
#endif
#endif /* B200CANNY_H */
