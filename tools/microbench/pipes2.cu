// Second round: 3-operand integer adds / logic, and whether instructions of different pipes overlap (B200).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 512
#define DEF(name, body)                                                                         \
  __global__ void __launch_bounds__(256) k_##name(uint32_t *out, uint32_t s0, uint32_t s1)       \
  {                                                                                             \
    uint32_t a[8], b[8];                                                                        \
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 17u + i * s0 + blockIdx.x; b[i] = a[i] * s1 + 7u; } \
    uint32_t c = s0 | 1u, d = s1 * 3u;                                                          \
    for (int it = 0; it < ITER; ++it) {                                                         \
      _Pragma("unroll") for (int i = 0; i < 8; ++i) { body; }                                   \
    }                                                                                           \
    uint32_t r = 0;                                                                             \
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ b[i];                                               \
    if (r == 0x12345678u) out[threadIdx.x] = r;                                                 \
  }
#define IADD3(x, y, z) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x) : "r"(y), "r"(z))
#define LOP3(x, y, z) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z))
#define PRMT(x, y) asm volatile("prmt.b32 %0, %0, %1, 0x5432;" : "+r"(x) : "r"(y))
#define IMAD(x, y, z) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define IMADI(x, z) asm volatile("mad.lo.u32 %0, %0, 5, %1;" : "+r"(x) : "r"(z))
#define HFMA2(x, y, z) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define FFMA(x, y, z) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(y), "f"(z))
#define IDP(x, y, z) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
DEF(iadd3, IADD3(a[i], b[i], c))
DEF(iadd3_same, IADD3(a[i], c, d))
DEF(lop3, LOP3(a[i], b[i], c))
DEF(prmt, PRMT(a[i], b[i]))
DEF(imad3, IMAD(a[i], b[i], c))
DEF(imad_imm, IMADI(a[i], c))
DEF(prmt_imad, { PRMT(a[i], c); IMAD(b[i], c, d); })
DEF(prmt_hfma2, { PRMT(a[i], c); HFMA2(b[i], c, d); })
DEF(iadd3_imad, { IADD3(a[i], c, d); IMAD(b[i], c, d); })
DEF(iadd3_idp, { IADD3(a[i], c, d); IDP(b[i], c, d); })
DEF(prmt_iadd3, { PRMT(a[i], c); IADD3(b[i], c, d); })
DEF(imad_hfma2, { IMAD(a[i], c, d); HFMA2(b[i], c, d); })
DEF(prmt_ffma, { PRMT(a[i], c); float f = __uint_as_float(b[i]); FFMA(f, __uint_as_float(c), __uint_as_float(d)); b[i] = __float_as_uint(f); })
DEF(imad_ffma, { IMAD(a[i], c, d); float f = __uint_as_float(b[i]); FFMA(f, __uint_as_float(c), __uint_as_float(d)); b[i] = __float_as_uint(f); })
template <class K> void run(const char *name, K k, uint32_t *d, double ops_per_iter, int sms, double ghz)
{
  const int blocks = sms * 8;
  k<<<blocks, 256>>>(d, 3, 5);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) k<<<blocks, 256>>>(d, 3, 5);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double warp_inst = 5.0 * blocks * 8 * ITER * ops_per_iter;
  const double ipc = warp_inst / (sms * 4.0) / (ms * 1e-3 * ghz * 1e9);
  printf("%-12s %8.3f ms  IPC %.3f per SMSP -> %.2f cycles per instruction\n", name, ms, ipc, 1.0 / ipc);
}
int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  uint32_t *d; cudaMalloc(&d, 4096);
  const int sms = p.multiProcessorCount;
#define R(n, ops) run(#n, k_##n, d, ops, sms, ghz)
  R(iadd3, 8); R(iadd3_same, 8); R(lop3, 8); R(prmt, 8); R(imad3, 8); R(imad_imm, 8);
  R(prmt_imad, 16); R(prmt_hfma2, 16); R(iadd3_imad, 16); R(iadd3_idp, 16); R(prmt_iadd3, 16); R(imad_hfma2, 16); R(prmt_ffma, 16); R(imad_ffma, 16);
  return 0;
}
