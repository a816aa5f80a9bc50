// Saturated throughput of the dense row code of the stencil kernel (warp A: gray + Gaussian; warp C: Sobel + pre-filter)
// without barriers, sparse stages or DRAM: W warps per CTA all run the same role on aliased shared-memory rings.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../cudacam_b200/csrc/k_stencil_march.cuh"
using namespace b2c;
template <int ROLE>
__global__ void __launch_bounds__(128, 7) k_rows(const B2cStencilParams p, int blocks, uint32_t *out)
{
  B2C_DYN_SMEM(smem);
  MarchGeo g;
  g.lane = threadIdx.x & 31; g.X0 = 240; g.Y0 = 0; g.frame = 0; g.rows_out = 1 << 20; g.xl = 240 - 8 + 8 * g.lane; g.yg0 = 100; g.ilim = 1 << 20; g.nblocks = blocks;
  g.lane_in = true; g.out_lane = g.lane != 0 && g.lane != 31; g.partial = false;
  for (int j = 0; j < 4; ++j) g.pm[j] = 0xFFFFFFFFu;
  uint32_t acc = 0;
  if (ROLE == 0) {
    MarchA<3> a;
    for (int k = 0; k < 3; ++k) for (int j = 0; j < 4; ++j) { a.G[k][j] = threadIdx.x * 3 + k + j; a.BX[k][j] = a.G[k][j] * 3; }
    const uint8_t *base = p.bgr + 24 * g.lane;
    for (int q = 0; q < 3; ++q) a.pre[q] = make_uint2(threadIdx.x, q);
    a.np = base; a.gaddr = g.lane * 8;
    for (int t = 0; t < blocks; ++t) {
      a.np = base + (t & 7) * 768;   // stays in L1/L2
      uint32_t zw = 0;
#pragma unroll 1
      for (int j = 0; j < 4; ++j) m_a_rows3<3, false>(p, g, smem, a, 768, 12 * t + 3 * j, smem + MS_BLUR + j * 1536 + g.lane * 16, zw);
      acc += zw;
    }
  } else {
    MarchC c;
    for (int k = 0; k < 3; ++k) for (int j = 0; j < 4; ++j) { c.D[k][j] = threadIdx.x + k; c.T[k][j] = j; }
    uint32_t fc = 0;
    for (int t = 0; t < blocks; ++t) {
#pragma unroll 1
      for (int it = 0; it < 2; ++it) m_c_rows6(p, g, smem, c, 12 * t, it, false, 0x7bff7bffu, fc, false);
    }
    acc = fc + c.D[0][0];
  }
  if (acc == 0x12345678u) out[threadIdx.x] = acc;
}
int main()
{
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  uint8_t *buf; cudaMalloc(&buf, 1 << 20); cudaMemset(buf, 7, 1 << 20);
  uint32_t *out; cudaMalloc(&out, 4096);
  B2cStencilParams p; memset(&p, 0, sizeof(p));
  p.bgr = buf; p.zeros = buf; p.row_stride = 768; p.w = 1 << 20; p.h = 1 << 20; p.h_glob = 1 << 20; p.lo = 10; p.hi = 40;
  for (int i = 0; i < 3; ++i) { p.n_lo[i] = 1e30f; p.n_hi[i] = 1e30f; }
  cudaFuncSetAttribute(k_rows<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, MARCH_SMEM);
  cudaFuncSetAttribute(k_rows<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MARCH_SMEM);
  const int blocks = 200;
  for (int role = 0; role < 2; ++role)
    for (int ctas_per_sm = 1; ctas_per_sm <= 7; ++ctas_per_sm) {
      int occ = 0;
      if (role == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rows<0>, 128, MARCH_SMEM); else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rows<1>, 128, MARCH_SMEM);
      if (ctas_per_sm > occ) break;
      const int grid = sms * ctas_per_sm;
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        if (role == 0) k_rows<0><<<grid, 128, MARCH_SMEM>>>(p, blocks, out); else k_rows<1><<<grid, 128, MARCH_SMEM>>>(p, blocks, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
      }
      float ms; cudaEventElapsedTime(&ms, a, b);
      const double rows = (double)grid * 4 * blocks * 12;
      const double cyc_per_row_smsp = ms * 1e-3 * 1.965e9 / (rows / (sms * 4.0));
      printf("role %c  %d warps/SMSP: %.3f ms, %.1f cycles per row per SMSP (%s)\n", role ? 'C' : 'A', ctas_per_sm, ms, cyc_per_row_smsp, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
