// Micro-benchmarks (GPU box only): (1) tcgen05.mma issue/throughput, SS and TS operands, N=128/256, one or two
// accumulators; (2) TMA 16 KB box loads per SM with no consumer work.  Prints cycles per operation.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../two-towers-overlords_b200/csrc/tt_ptx.cuh"
#include "../../two-towers-overlords_b200/csrc/tt_tma.cuh"
namespace tt { void set_error(const char*, ...) {} int check_cuda(cudaError_t e, const char*, const char*, int) { return e != cudaSuccess; } void note_launch() {} }
using namespace tt::ptx;

__device__ __forceinline__ void mbar_wait_test(uint64_t* bar, uint32_t parity) {  // non-suspending poll
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) mma_bench(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t da = make_smem_desc_sw128(smem_u32(smem)), db = make_smem_desc_sw128(smem_u32(smem) + 16384);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t acc = tb + (uint32_t)(((i * 4 + k) % NACC) * N);
        if (TS) mma_bf16_ts(acc, tb + 448 + k * 8, db + 2 * k, idesc, 1);
        else mma_bf16(acc, da + 2 * k, db + 2 * k, idesc, 1);
      }
    }
    long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

__global__ void __launch_bounds__(64, 1) tma_bench(const __grid_constant__ CUtensorMap map, int iters, int stages, int kb,
                                                   long long rows_per_cta, long long* out, int box_rows = 128, int poll = 0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t full[16], empty[16];
  if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } fence_barrier_init(); }
  __syncthreads();
  const long long row0 = (long long)blockIdx.x * rows_per_cta;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      if (poll) mbar_wait_test(&empty[s], ph ^ 1); else mbar_wait(&empty[s], ph ^ 1);
      mbar_arrive_expect_tx(&full[s], box_rows * 128);
      tma_load_2d(smem + s * box_rows * 128, &map, &full[s], (it % kb) * 64, (int)(row0 + (long long)(it / kb) * box_rows));
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      if (poll) mbar_wait_test(&full[s], ph); else mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

// pure issue rate: `n` loads into `n` distinct slots, no slot reuse; t_issue = after the last issue, t_done = all landed
__global__ void __launch_bounds__(64, 1) tma_issue_bench(const __grid_constant__ CUtensorMap map, int n, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) { for (int s = 0; s < n; ++s) mbar_init(&full[s], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int s = 0; s < n; ++s) {
      mbar_arrive_expect_tx(&full[s], 16384);
      tma_load_2d(smem + s * 16384, &map, &full[s], (s % 6) * 64, (s / 6) * 128 + blockIdx.x * 1024);
    }
    long long t1 = clock64();
    long long tl[16];
    for (int s = 0; s < n; ++s) { mbar_wait(&full[s], 0); tl[s] = clock64(); }
    if (blockIdx.x == 0) { out[0] = t1 - t0; for (int s = 0; s < n; ++s) out[1 + s] = tl[s] - t0; }
  }
}

// cluster of 2: each CTA issues half of every 16 KB tile as a multicast into both CTAs
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1)
    tma_mc_bench(const __grid_constant__ CUtensorMap map_half, int iters, int stages, int kb, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t full[16], empty[16];
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); } fence_barrier_init(); }
  __syncthreads();
  cluster_sync();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(&empty[s], ph ^ 1);   // both CTAs' consumers have released the slot
      mbar_arrive_expect_tx(&full[s], 16384);
      tma_load_2d_mc(smem + s * 16384 + rank * 8192, &map_half, &full[s], (it % kb) * 64, (it / kb) * 128 + (int)rank * 64, 3);
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(&full[s], ph);
      mbar_arrive_cluster(&empty[s], 0);
      mbar_arrive_cluster(&empty[s], 1);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  cluster_sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* out; cudaMallocManaged(&out, 64);
  const int iters = 4096;
#define RUN_MMA(N, TS, NACC)                                                                            \
  cudaFuncSetAttribute(mma_bench<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);  \
  mma_bench<N, TS, NACC><<<148, 128, 70 * 1024>>>(iters, out); cudaDeviceSynchronize();                 \
  mma_bench<N, TS, NACC><<<148, 128, 70 * 1024>>>(iters, out);                                          \
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("mma_bench failed\n"); return 1; }                \
  printf("MMA M=128 N=%d %s acc=%d: issue %.1f clk/MMA, complete %.1f clk/MMA\n", N, TS ? "TS" : "SS", NACC,   \
         (double)out[0] / (iters * 4), (double)out[1] / (iters * 4));
  RUN_MMA(128, false, 1) RUN_MMA(128, true, 1) RUN_MMA(128, false, 2) RUN_MMA(256, false, 1) RUN_MMA(256, true, 1)
  // TMA: 2M x 384 bf16 docs
  const long long N = 2000000; const int P = 384;
  void* d; cudaMalloc(&d, N * P * 2); cudaMemset(d, 0, N * P * 2);
  CUtensorMap map;
  if (tt::make_map_bf16_kmajor(&map, d, N, P, P, 128)) { printf("map failed\n"); return 1; }
  cudaFuncSetAttribute(tma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int mode = 0; mode < 2; ++mode)
    for (int stages : {2, 4, 8, 12}) {
      const int it2 = 6 * 512;  // 512 tiles of 128 docs
      const long long rows_per_cta = mode == 0 ? 0 : 512 * 128;  // 0: every CTA reads the same rows (L2), 1: own rows
      tma_bench<<<148, 64, stages * 16384 + 2048>>>(map, it2, stages, 6, rows_per_cta, out); cudaDeviceSynchronize();
      tma_bench<<<148, 64, stages * 16384 + 2048>>>(map, it2, stages, 6, rows_per_cta, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("tma_bench failed\n"); return 1; }
      printf("TMA %s stages=%2d: %.0f clk per 16 KB load -> %.1f B/clk/SM\n", mode ? "distinct rows" : "shared rows  ", stages,
             (double)out[0] / it2, 16384.0 * it2 / out[0]);
    }
  for (int ctas : {16, 37, 74}) {
    const int it2 = 6 * 512;
    tma_bench<<<ctas, 64, 8 * 16384 + 2048>>>(map, it2, 8, 6, 0, out); cudaDeviceSynchronize();
    tma_bench<<<ctas, 64, 8 * 16384 + 2048>>>(map, it2, 8, 6, 0, out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("tma_bench failed\n"); return 1; }
    printf("TMA shared rows, %3d CTAs, 8 stages: %.0f clk per 16 KB load -> %.1f B/clk/SM\n", ctas, (double)out[0] / it2, 16384.0 * it2 / out[0]);
  }
  {
    typedef tt::EncodeTiledFn Fn;
    Fn fn = tt::encode_tiled_fn();
    void* d2; cudaMalloc(&d2, (size_t)200000 * 4096 * 2); cudaMemset(d2, 0, (size_t)200000 * 4096 * 2);
    struct Cfg { const char* name; void* base; uint64_t rows, cols; uint32_t box_rows; CUtensorMapL2promotion prom; int kb; };
    Cfg cfgs[] = {
      {"pitch 768 B, 128 rows, L2 promo 256", d, 2000000, 384, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 6},
      {"pitch 768 B, 128 rows, L2 promo 128", d, 2000000, 384, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 6},
      {"pitch 768 B, 128 rows, L2 promo none", d, 2000000, 384, 128, CU_TENSOR_MAP_L2_PROMOTION_NONE, 6},
      {"pitch 768 B,  64 rows, L2 promo 128", d, 2000000, 384, 64, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 6},
      {"pitch 768 B, 256 rows, L2 promo 128", d, 2000000, 384, 256, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 6},
      {"pitch 8 KB, 128 rows, L2 promo 128", d2, 200000, 4096, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 64},
      {"pitch 8 KB, 128 rows, L2 promo 256", d2, 200000, 4096, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 64},
    };
    for (auto& c : cfgs) {
      CUtensorMap m;
      const cuuint64_t dims[2] = {c.cols, c.rows}; const cuuint64_t strides[1] = {c.cols * 2};
      const cuuint32_t box[2] = {64, c.box_rows}; const cuuint32_t es[2] = {1, 1};
      if (fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c.base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, c.prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
      const int it2 = 3072; const int stages = 4;
      tma_bench<<<148, 64, stages * c.box_rows * 128 + 2048>>>(m, it2, stages, c.kb, 0, out, c.box_rows); cudaDeviceSynchronize();
      tma_bench<<<148, 64, stages * c.box_rows * 128 + 2048>>>(m, it2, stages, c.kb, 0, out, c.box_rows);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("tma_bench failed\n"); return 1; }
      printf("TMA %-40s: %.0f clk per load -> %.1f B/clk/SM\n", c.name, (double)out[0] / it2, c.box_rows * 128.0 * it2 / out[0]);
    }
  }
  for (int poll = 0; poll < 2; ++poll)
    for (int stages : {4, 8}) {
      const int it2 = 3072;
      tma_bench<<<148, 64, stages * 16384 + 2048>>>(map, it2, stages, 6, 0, out, 128, poll); cudaDeviceSynchronize();
      tma_bench<<<148, 64, stages * 16384 + 2048>>>(map, it2, stages, 6, 0, out, 128, poll);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("tma_bench failed\n"); return 1; }
      printf("TMA loop %s stages=%d: %.0f clk per 16 KB load\n", poll ? "test_wait poll" : "try_wait      ", stages, (double)out[0] / it2);
    }
  cudaFuncSetAttribute(tma_issue_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    tma_issue_bench<<<148, 64, 12 * 16384 + 2048>>>(map, 12, out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("issue bench failed\n"); return 1; }
  }
  printf("TMA issue of 12 loads: %lld clk total; landed at:", out[0]);
  for (int s = 0; s < 12; ++s) printf(" %lld", out[1 + s]);
  printf("\n");
  CUtensorMap map_half;
  if (tt::make_map_bf16_kmajor(&map_half, d, N, P, P, 64)) { printf("map failed\n"); return 1; }
  cudaFuncSetAttribute(tma_mc_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int stages : {4, 8}) {
    const int it2 = 6 * 512;
    tma_mc_bench<<<148, 64, stages * 16384 + 2048>>>(map_half, it2, stages, 6, out); cudaDeviceSynchronize();
    tma_mc_bench<<<148, 64, stages * 16384 + 2048>>>(map_half, it2, stages, 6, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tma_mc_bench failed: %s\n", cudaGetErrorString(e)); return 1; }
    printf("TMA multicast x2 stages=%2d: %.0f clk per 16 KB tile received -> %.1f B/clk/SM\n", stages, (double)out[0] / it2, 16384.0 * it2 / out[0]);
  }
  return 0;
}
