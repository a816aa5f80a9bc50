// k_stencil_fused.cuh -- throughput version of the fused stencil (placeholder until the packed-integer kernel lands).
#pragma once
#include "b2c_device.cuh"

namespace b2c
{
#ifdef B2C_EMU
inline int fused_emu_launch(const B2cStencilParams &) { return -1; }
#else
inline cudaError_t fused_configure() { return cudaSuccess; }
inline bool fused_supported(const B2cStencilParams &) { return false; }
inline cudaError_t fused_launch(const B2cStencilParams &, int, cudaStream_t) { return cudaErrorNotSupported; }
#endif
}// namespace b2c
