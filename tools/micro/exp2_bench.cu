// Micro-benchmark: throughput of the exp2 variants a softmax inner loop can use on sm_100a (elements / clk / SM).
//   A  ex2.approx.ftz.f32            (MUFU, one element per lane-op)
//   B  ex2.approx.ftz.bf16x2         (two elements per lane-op, bf16 in / bf16 out)
//   C  ex2.approx.f16x2              (two elements per lane-op, fp16)
//   D  Cody-Waite + degree-3 polynomial on packed fp32x2 FMA pipes
//   E  mixed: 3 of 4 elements on MUFU (A), 1 of 4 on the polynomial (D)
//   F  mixed: 1 of 2 each
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ecad_b200/csrc -o tools/micro/exp2_bench tools/micro/exp2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ecadk;

__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) {
  uint32_t y;
  asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
  uint32_t y;
  asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
// exp2 of two fp32 values on the FMA pipes: n = round(x), f = x - n in [-0.5, 0.5], 2^f by a cubic, exponent add
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& y0, float& y1) {
  const uint64_t magic = pack_f2(12582912.f, 12582912.f);
  const uint64_t nmagic = pack_f2(-12582912.f, -12582912.f);
  const uint64_t x = pack_f2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const uint64_t t = fadd2(x, magic);
  const uint64_t n = fadd2(t, nmagic);
  const uint64_t f = ffma2(n, pack_f2(-1.f, -1.f), x);
  uint64_t p = ffma2(pack_f2(0.05550410866f, 0.05550410866f), f, pack_f2(0.2402265070f, 0.2402265070f));
  p = ffma2(p, f, pack_f2(0.6931471806f, 0.6931471806f));
  p = ffma2(p, f, pack_f2(1.0f, 1.0f));
  float p0, p1, t0, t1;
  unpack_f2(p, p0, p1);
  unpack_f2(t, t0, t1);
  y0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  y1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, long long* clocks, int iters) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -0.01f * (threadIdx.x + i);
  float acc = 0.f;
  uint32_t acc2 = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      if constexpr (MODE == 0) {
        acc += fast_exp2(v[i]) + fast_exp2(v[i + 1]) + fast_exp2(v[i + 2]) + fast_exp2(v[i + 3]);
      } else if constexpr (MODE == 1) {
        acc2 ^= ex2_bf16x2(pack_bf16x2(v[i], v[i + 1])) + ex2_bf16x2(pack_bf16x2(v[i + 2], v[i + 3]));
      } else if constexpr (MODE == 2) {
        uint32_t h0, h1;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(v[i + 1]), "f"(v[i]));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(v[i + 3]), "f"(v[i + 2]));
        acc2 ^= ex2_f16x2(h0) + ex2_f16x2(h1);
      } else if constexpr (MODE == 3) {
        float a, b, c, d;
        exp2_poly2(v[i], v[i + 1], a, b);
        exp2_poly2(v[i + 2], v[i + 3], c, d);
        acc += a + b + c + d;
      } else if constexpr (MODE == 4) {
        float a = fast_exp2(v[i]), b = fast_exp2(v[i + 1]), c, d;
        if (i & 4) {
          c = fast_exp2(v[i + 2]);
          d = fast_exp2(v[i + 3]);
        } else {
          exp2_poly2(v[i + 2], v[i + 3], c, d);
        }
        acc += a + b + c + d;
      } else if constexpr (MODE == 5) {
        float a = fast_exp2(v[i]), b = fast_exp2(v[i + 1]), c, d;
        exp2_poly2(v[i + 2], v[i + 3], c, d);
        acc += a + b + c + d;
      } else if constexpr (MODE == 6) {  // what the softmax loop does: MUFU + cvt.rn.bf16x2 pack
        const float a = fast_exp2(v[i]), b = fast_exp2(v[i + 1]), c = fast_exp2(v[i + 2]), d = fast_exp2(v[i + 3]);
        acc += a + b + c + d;
        acc2 ^= pack_bf16x2(a, b) ^ pack_bf16x2(c, d);
      } else {  // MUFU + integer round-to-nearest pack (IADD, IADD, PRMT) instead of F2FP
        const float a = fast_exp2(v[i]), b = fast_exp2(v[i + 1]), c = fast_exp2(v[i + 2]), d = fast_exp2(v[i + 3]);
        acc += a + b + c + d;
        acc2 ^= __byte_perm(__float_as_uint(a) + 0x8000u, __float_as_uint(b) + 0x8000u, 0x7632) ^
                __byte_perm(__float_as_uint(c) + 0x8000u, __float_as_uint(d) + 0x8000u, 0x7632);
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += 1e-6f * acc + __uint_as_float(acc2 & 1u);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(acc2);
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int blocks_per_sm) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * blocks_per_sm, iters = 2000;
  float* out;
  long long* clk;
  cudaMalloc(&out, blocks * 256 * sizeof(float));
  cudaMalloc(&clk, blocks * sizeof(long long));
  bench<MODE><<<blocks, 256>>>(out, clk, 10);
  bench<MODE><<<blocks, 256>>>(out, clk, iters);
  cudaDeviceSynchronize();
  long long h[4096];
  cudaMemcpy(h, clk, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < blocks; ++i) mean += h[i];
  mean /= blocks;
  const double elems_per_sm = double(blocks_per_sm) * 256 * 32 * iters;
  printf("%-34s warps/SMSP=%d  %.2f elements/clk/SM (incl. the loop-carried update: 32 FFMA per 32 elements)\n", name,
         blocks_per_sm * 2, elems_per_sm / mean);
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  // 1 block of 256 threads per SM = 2 warps per SMSP, the occupancy of the softmax warps in the attention kernels
  for (int bps : {1, 2}) {
    run<0>("A ex2.approx.ftz.f32", bps);
    run<1>("B ex2.approx.ftz.bf16x2 (+cvt)", bps);
    run<2>("C ex2.approx.f16x2 (+cvt)", bps);
    run<3>("D polynomial fp32x2", bps);
    run<4>("E 3/4 MUFU + 1/4 polynomial", bps);
    run<5>("F 1/2 MUFU + 1/2 polynomial", bps);
    run<6>("G MUFU + cvt.rn.bf16x2 pack", bps);
    run<7>("H MUFU + integer-rounded PRMT pack", bps);
  }
  return 0;
}
