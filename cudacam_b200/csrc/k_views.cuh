// k_views.cuh -- small export kernels around the hot path.
//
//   k_grad_view : saturating u8 view of the fp32 gradient, what the reference shows for stage GRADIENT
//                 (`float2uchar`, src/cvp/cannyEdgeD.cu:35-50, launched at cannyEdgeH.cu:181-186).
//   k_planes_to_map2 : the 2-bit map in its accessor format (one u32 per 16 pixels: strong bits | weak-only bits << 16) from
//                 the bit planes S and C the stencil kernel writes.
//   k_copy2d_u8 : pitched u8 -> tight u8 copy, the device-to-device copy of `_sendOutputToOpenGL`
//                 (src/cvp/cannyEdgeH.cu:188-207) for the other stages.
#pragma once
#include "b2c_device.cuh"

namespace b2c
{
__global__ void __launch_bounds__(256) k_grad_view(const float *__restrict__ grad, int pitchf, uint8_t *__restrict__ out, int out_pitch, int w, int h)
{
  const long long total = (long long)w * h;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / w), x = (int)(i - (long long)y * w);
    const float a = fabsf(grad[(long long)y * pitchf + x]);
    out[(long long)y * out_pitch + x] = (uint8_t)(a < 255.0f ? a : 255.0f);
  }
}

__global__ void __launch_bounds__(256) k_copy2d_u8(const uint8_t *__restrict__ in, int in_pitch, uint8_t *__restrict__ out, int out_pitch, int w, int h)
{
  const long long total = (long long)w * h;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / w), x = (int)(i - (long long)y * w);
    out[(long long)y * out_pitch + x] = in[(long long)y * in_pitch + x];
  }
}
__global__ void __launch_bounds__(256) k_planes_to_map2(const uint16_t *__restrict__ S16, const uint16_t *__restrict__ C16, int pitch16, uint32_t *__restrict__ map2, int map_pitch, int h)
{
  const long long total = (long long)map_pitch * h;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / map_pitch), g = (int)(i - (long long)y * map_pitch);
    const uint32_t s = S16[(long long)y * pitch16 + g], c = C16[(long long)y * pitch16 + g];
    map2[i] = s | ((c & ~s) << 16);
  }
}
}// namespace b2c
