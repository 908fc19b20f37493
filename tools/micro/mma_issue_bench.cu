// Cost of one tcgen05.mma (cta_group::1, M = 128, K = 16, bf16) as seen by the issuing thread, for the instruction
// shapes the attention kernels use: dependent chains (every MMA accumulates into the SAME TMEM tile - what a Q K^T or
// a P V of one query tile is) against chains that alternate between 2 / 4 accumulators, A from shared memory (SS) or
// from TMEM (TS), N = 64 / 80 / 128 / 256.  One CTA per SM, one issuing thread, operands are zeros in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/mma_issue_bench tools/micro/mma_issue_bench.cu
//   ./tools/micro/mma_issue_bench            (prints clocks per MMA, mean over the CTAs)
#include <cstdio>
#include <vector>

#include "../../ecad_b200/csrc/ptx.cuh"

using namespace ecadk;

struct Case {
  int n;        // MMA N
  int nacc;     // accumulators the chain cycles through
  int ts;       // 1 = A operand from TMEM
  int count;    // MMAs per measurement
  int second_issuer;  // 1 = a second thread (another warp) issues the same chain on its own accumulators
};

__global__ void __launch_bounds__(128, 1) bench(Case c, unsigned int* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const int warp = threadIdx.x >> 5;
  if ((warp == 0 || (warp == 1 && c.second_issuer)) && (threadIdx.x & 31) == 0) {
    const uint32_t sb = smem_u32(smem);
    const uint64_t da = make_smem_desc(sb, 16, 1024, kLayoutSW128);               // A: 128 rows x 64 columns
    const uint64_t db = make_smem_desc(sb + 16 * 1024, 16, 1024, kLayoutSW128);   // B: up to 256 rows x 64 columns
    const uint32_t idesc = make_idesc_bf16(128, c.n);
    // accumulators: nacc tiles of n columns; with a second issuer each thread owns half of TMEM; TS: A at column 448
    const uint32_t base = tmem + (warp == 1 ? 256 : 0);
    const uint32_t stride = c.n <= 128 ? 128 : 256;
    const uint32_t a_tmem = tmem + 448;
    // warm-up
    for (int i = 0; i < 8; ++i) umma_bf16_ss(base, da, db, idesc, 0);
    umma_commit(&bars[warp]);
    mbar_wait(&bars[warp], 0);
    tc_fence_after();
    const unsigned int t0 = clock();
    for (int i = 0; i < c.count; ++i) {
      const uint32_t d = base + (i % c.nacc) * stride;
      const int k = i & 3;
      if (c.ts) umma_bf16_ts(d, a_tmem + k * 8, db + 2 * k, idesc, 1);
      else umma_bf16_ss(d, da + 2 * k, db + 2 * k, idesc, 1);
    }
    const unsigned int t1 = clock();
    umma_commit(&bars[warp]);
    mbar_wait(&bars[warp], 1);
    const unsigned int t2 = clock();
    if (warp == 0) {
      out[blockIdx.x * 2] = t1 - t0;      // issue loop alone
      out[blockIdx.x * 2 + 1] = t2 - t0;  // until the last MMA has retired
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  const int grid = 148, smem = 48 * 1024 + 1024 + 256;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  unsigned int* d_out;
  cudaMalloc(&d_out, grid * 2 * sizeof(unsigned int));
  std::vector<unsigned int> h(grid * 2);
  const Case cases[] = {
      {128, 1, 0, 64, 0}, {128, 2, 0, 64, 0}, {128, 4, 0, 64, 0}, {128, 1, 0, 64, 1}, {128, 2, 0, 64, 1},
      {256, 1, 0, 64, 0}, {256, 2, 0, 64, 0}, {64, 1, 0, 64, 0},  {64, 2, 0, 64, 0},  {64, 4, 0, 64, 0},
      {80, 1, 0, 64, 0},  {80, 2, 0, 64, 0},  {80, 1, 1, 64, 0},  {80, 2, 1, 64, 0},  {128, 1, 1, 64, 0},
      {128, 2, 1, 64, 0}, {128, 1, 0, 8, 0},  {128, 1, 0, 16, 0}, {128, 1, 0, 256, 0}, {80, 1, 1, 8, 0},
  };
  printf("# tcgen05.mma cta_group::1 M=128 K=16 bf16, clocks per MMA (mean over %d CTAs); ideal = N/2\n", grid);
  printf("# %4s %5s %3s %6s %8s | %10s %12s\n", "N", "accs", "A", "count", "issuers", "issue/MMA", "retired/MMA");
  for (const Case& c : cases) {
    bench<<<grid, 128, smem>>>(c, d_out);
    if (cudaDeviceSynchronize() != cudaSuccess) {
      printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      return 1;
    }
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost);
    double a = 0, b = 0;
    for (int i = 0; i < grid; ++i) {
      a += h[2 * i];
      b += h[2 * i + 1];
    }
    printf("  %4d %5d %3s %6d %8d | %10.1f %12.1f\n", c.n, c.nacc, c.ts ? "TS" : "SS", c.count, 1 + c.second_issuer,
           a / grid / c.count, b / grid / c.count);
  }
  return 0;
}
