// Throughput of the instruction classes the stencil kernel is made of, on the GPU at hand (warp-instructions per cycle
// per SM sub-partition).  8 independent chains per thread, 16 warps per SM sub-partition... build: nvcc -arch=sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 512
#define DEF(name, body)                                                                         \
  __global__ void __launch_bounds__(256) k_##name(uint32_t *out, uint32_t s0, uint32_t s1)       \
  {                                                                                             \
    uint32_t a[8];                                                                              \
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 17u + i * s0 + blockIdx.x;                 \
    uint32_t b = s1, c = s0 | 1u;                                                               \
    for (int it = 0; it < ITER; ++it) {                                                         \
      _Pragma("unroll") for (int i = 0; i < 8; ++i) { body; }                                   \
    }                                                                                           \
    uint32_t r = 0;                                                                             \
    for (int i = 0; i < 8; ++i) r ^= a[i];                                                      \
    if (r == 0x12345678u) out[threadIdx.x] = r;                                                 \
  }
DEF(iadd3, a[i] = a[i] + b + c)
DEF(imad, a[i] = a[i] * c + b)
DEF(imadhi, a[i] = __umulhi(a[i], c) + b)
DEF(idp4a, a[i] = __dp4a(a[i], c, b))
DEF(prmt, a[i] = __byte_perm(a[i], b, 0x5432))
DEF(lop3, a[i] = (a[i] & b) ^ c)
DEF(shf, a[i] = __funnelshift_l(a[i], b, 3))
DEF(hfma2, asm("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(b)))
DEF(hadd2, asm("add.rn.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b)))
DEF(hmnmx2, asm("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b)))
DEF(ffma, { float f = __uint_as_float(a[i]); f = __fmaf_rn(f, __uint_as_float(c), __uint_as_float(b)); a[i] = __float_as_uint(f); })
DEF(fhfma, { float f; asm("{.reg .f16 l, h; mov.b32 {l, h}, %1; fma.rn.f32.f16 %0, l, l, %2;}" : "=f"(f) : "r"(a[i]), "f"(__uint_as_float(b))); a[i] = __float_as_uint(f); })
DEF(shfl, a[i] = __shfl_up_sync(0xffffffffu, a[i], 1))
DEF(i2f, a[i] = __float_as_uint((float)a[i]))
DEF(popc, a[i] = __popc(a[i]) + b)
DEF(imad_shl, a[i] = (a[i] << 16) + b)
DEF(isetp_sel, a[i] = (a[i] > b) ? c : a[i] + 1u)
DEF(vote, a[i] = __ballot_sync(0xffffffffu, a[i] & 1u) + b)
__global__ void __launch_bounds__(256) k_lds(uint32_t *out, uint32_t s0, uint32_t s1)
{
  __shared__ uint4 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = make_uint4(i, s0, s1, 1);
  __syncthreads();
  uint32_t r = 0, idx = threadIdx.x;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { uint4 v = sm[(idx + i * 32) & 1023]; r += v.x + v.w; }
    idx += s0;
  }
  if (r == 0x12345678u) out[threadIdx.x] = r;
}
__global__ void __launch_bounds__(256) k_sts(uint32_t *out, uint32_t s0, uint32_t s1)
{
  __shared__ uint4 sm[1024];
  uint32_t idx = threadIdx.x;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[(idx + i * 32) & 1023] = make_uint4(idx, s0, s1, i);
    idx += s0;
  }
  __syncthreads();
  if (sm[threadIdx.x].x == 0x12345678u) out[threadIdx.x] = 1;
}
template <class K> void run(const char *name, K k, uint32_t *d, double ops_per_iter, int sms, double ghz_hint)
{
  const int blocks = sms * 8;   // 8 CTAs of 8 warps per SM = 16 warps per sub-partition
  k<<<blocks, 256>>>(d, 3, 5);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) k<<<blocks, 256>>>(d, 3, 5);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double warp_inst = 5.0 * blocks * 8 * ITER * ops_per_iter;
  const double per_smsp_per_cycle = warp_inst / (sms * 4.0) / (ms * 1e-3 * ghz_hint * 1e9);
  printf("%-10s %8.3f ms  %.3f warp-inst/cycle/SMSP (at %.2f GHz)  -> %.2f cycles per instruction\n", name, ms, per_smsp_per_cycle, ghz_hint, 1.0 / per_smsp_per_cycle);
}
int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
  uint32_t *d; cudaMalloc(&d, 4096);
  const int sms = p.multiProcessorCount;
#define R(n, ops) run(#n, k_##n, d, ops, sms, ghz)
  R(iadd3, 8); R(imad, 8); R(imadhi, 8); R(idp4a, 8); R(prmt, 8); R(lop3, 8); R(shf, 8); R(hfma2, 8); R(hadd2, 8); R(hmnmx2, 8); R(ffma, 8); R(fhfma, 8);
  R(shfl, 8); R(i2f, 8); R(popc, 8); R(imad_shl, 8); R(isetp_sel, 8); R(vote, 8); R(lds, 8); R(sts, 8);
  return 0;
}
