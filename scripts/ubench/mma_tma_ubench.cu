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

// MMA loop shaped like a pipelined consumer: groups of 4 MMAs, optional fence before and commit after each group
template <bool TS, int COMMIT_EVERY, bool FENCE>
__global__ void __launch_bounds__(128, 1) mma_group_bench(int groups, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bars[8];
  __shared__ uint64_t done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); mbar_init(&done, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 128);
    const uint64_t da = make_smem_desc_sw128(smem_u32(smem));
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (FENCE) tc_fence_after();
      const uint64_t db = make_smem_desc_sw128(smem_u32(smem) + 16384 + (g & 1) * 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (TS) mma_bf16_ts(tb, tb + 256 + (g % 6) * 32 + k * 8, db + 2 * k, idesc, 1);
        else mma_bf16(tb, da + 2 * k, db + 2 * k, idesc, 1);
      }
      if (COMMIT_EVERY > 0 && (g % COMMIT_EVERY) == COMMIT_EVERY - 1) mma_commit(&bars[g & 7]);
    }
    mma_commit(&done);
    mbar_wait(&done, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) out[0] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

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

// division-free streaming loop: every CTA walks `tiles` 128-row tiles (6 k-blocks each) starting at row0, `reps` times
__global__ void __launch_bounds__(64, 1) tma_stream_bench(const __grid_constant__ CUtensorMap map, int tiles, int reps,
                                                          int stages, long long rows_per_cta, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t full[16], empty[16];
  if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } fence_barrier_init(); }
  __syncthreads();
  const int row0 = (int)(blockIdx.x * rows_per_cta);
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    for (int r = 0; r < reps; ++r)
      for (int t = 0; t < tiles; ++t)
        for (int kb = 0; kb < 6; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], 16384);
          tma_load_2d(smem + s * 16384, &map, &full[s], kb * 64, row0 + t * 128);
          if (++s == stages) { s = 0; ph ^= 1; }
        }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < reps * tiles * 6; ++i) {
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

// producer + MMA consumer over a ring, like the scan main loop but with no epilogue: clk per 64-wide k-block
template <bool TS>
__global__ void __launch_bounds__(128, 1) pipe_bench(const __grid_constant__ CUtensorMap map, int tiles, int stages,
                                                     int commit_every, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t full[16], empty[16], done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } mbar_init(&done, 1); fence_barrier_init(); }
  if (warp == 2) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  uint8_t* ring = smem + 16384;  // [0,16K): an A tile for SS mode
  long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int t = 0; t < tiles; ++t)
      for (int kb = 0; kb < 6; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], 16384);
        tma_load_2d(ring + s * 16384, &map, &full[s], kb * 64, t * 128);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 128);
    const uint64_t da = make_smem_desc_sw128(smem_u32(smem));
    int s = 0; uint32_t ph = 0; int pend = 0; int pend_slot[8];
    for (int t = 0; t < tiles; ++t)
      for (int kb = 0; kb < 6; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t db = make_smem_desc_sw128(smem_u32(ring) + s * 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (TS) mma_bf16_ts(tb + (t & 1) * 128, tb + 256 + kb * 32 + k * 8, db + 2 * k, idesc, (kb | k) != 0);
          else mma_bf16(tb + (t & 1) * 128, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        pend_slot[pend++] = s;
        if (pend == commit_every) { for (int i = 0; i < pend; ++i) mma_commit(&empty[pend_slot[i]]); pend = 0; }
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    for (int i = 0; i < pend; ++i) mma_commit(&empty[pend_slot[i]]);
    mma_commit(&done);
    mbar_wait(&done, 0);
  }
  __syncthreads();
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 2) tmem_dealloc(tb, 512);
}

__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {  // lean wait: no timeout bookkeeping
  while (!mbar_try_wait(bar, parity)) {}
}
// lean consumer: ring of exactly 6 slots = the 6 k-blocks of a tile, fully unrolled, descriptors precomputed
template <int NISSUE>
__global__ void __launch_bounds__(160, 1) pipe_lean_bench(const __grid_constant__ CUtensorMap map, int tiles, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t full[12], empty[12], done[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NS = 6 * NISSUE;  // one ring section per issuing thread
  if (threadIdx.x == 0) { for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } mbar_init(&done[0], 1); mbar_init(&done[1], 1); fence_barrier_init(); }
  if (warp == 4) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  uint8_t* ring = smem;
  long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int t = 0; t < tiles; ++t)
      for (int kb = 0; kb < 6; ++kb) {
        mbar_spin(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], 16384);
        tma_load_2d(ring + s * 16384, &map, &full[s], kb * 64, t * 128);
        if (++s == NS) { s = 0; ph ^= 1; }
      }
  } else if ((warp == 1 || (NISSUE == 2 && warp == 2)) && lane == 0) {
    const int me = warp - 1;  // issuer `me` handles tiles t = me, me + NISSUE, ... in ring section me
    const uint32_t idesc = make_idesc_bf16(128, 128);
    uint64_t db[6];
#pragma unroll
    for (int kb = 0; kb < 6; ++kb) db[kb] = make_smem_desc_sw128(smem_u32(ring) + (me * 6 + kb) * 16384);
    uint32_t ph = 0;
    for (int t = me; t < tiles; t += NISSUE) {
#pragma unroll
      for (int kb = 0; kb < 6; ++kb) {
        mbar_spin(&full[me * 6 + kb], ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_bf16_ts(tb + me * 128, tb + 256 + kb * 32 + k * 8, db[kb] + 2 * k, idesc, (kb | k) != 0);
        mma_commit(&empty[me * 6 + kb]);
      }
      ph ^= 1;
    }
    mma_commit(&done[me]);
    mbar_wait(&done[me], 0);
  }
  __syncthreads();
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 4) tmem_dealloc(tb, 512);
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
#define RUN_GRP(TS, CE, FE)                                                                                  \
  cudaFuncSetAttribute(mma_group_bench<TS, CE, FE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);   \
  mma_group_bench<TS, CE, FE><<<148, 128, 70 * 1024>>>(4096, out); cudaDeviceSynchronize();                    \
  mma_group_bench<TS, CE, FE><<<148, 128, 70 * 1024>>>(4096, out);                                             \
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("mma_group_bench failed\n"); return 1; }                \
  printf("MMA groups of 4, %s, commit every %d groups, fence %d: %.1f clk/MMA\n", TS ? "TS" : "SS", CE, (int)FE, \
         (double)out[0] / (4096 * 4));
  RUN_GRP(true, 0, false) RUN_GRP(true, 1, false) RUN_GRP(true, 1, true) RUN_GRP(true, 3, true) RUN_GRP(false, 1, true)
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
  cudaFuncSetAttribute(pipe_bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(pipe_bench<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int ts = 0; ts < 2; ++ts)
    for (int stages : {3, 7, 12})
      for (int ce : {1, 3}) {
        const int tiles = 512;
        for (int rep = 0; rep < 2; ++rep) {
          if (ts) pipe_bench<true><<<148, 128, (stages + 1) * 16384 + 2048>>>(map, tiles, stages, ce, out);
          else pipe_bench<false><<<148, 128, (stages + 1) * 16384 + 2048>>>(map, tiles, stages, ce, out);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("pipe bench failed\n"); return 1; }
        }
        printf("PIPE %s stages=%2d commit_every=%d: %.0f clk per k-block (MMA floor 256)\n", ts ? "TS" : "SS", stages, ce, (double)out[0] / (tiles * 6));
      }
  cudaFuncSetAttribute(pipe_lean_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(pipe_lean_bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int rep = 0; rep < 2; ++rep) { pipe_lean_bench<1><<<148, 160, 6 * 16384 + 2048>>>(map, 512, out); if (cudaDeviceSynchronize() != cudaSuccess) { printf("lean failed\n"); return 1; } }
  printf("PIPE lean, 1 issuer : %.0f clk per k-block (MMA floor 256)\n", (double)out[0] / (512 * 6));
  for (int rep = 0; rep < 2; ++rep) { pipe_lean_bench<2><<<148, 160, 12 * 16384 + 2048>>>(map, 512, out); if (cudaDeviceSynchronize() != cudaSuccess) { printf("lean2 failed\n"); return 1; } }
  printf("PIPE lean, 2 issuers: %.0f clk per k-block (MMA floor 256)\n", (double)out[0] / (512 * 6));
  cudaFuncSetAttribute(tma_stream_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int mode = 0; mode < 3; ++mode)
    for (int stages : {4, 8}) {
      // mode 0: all CTAs walk the same 512 tiles (48 MB, L2-resident); mode 1: each CTA its own 5 tiles (0.5 MB; 71 MB total,
      // L2-resident) over and over; mode 2: each CTA its own 100 tiles (9.6 MB; 1.4 GB total: HBM)
      const int tiles = mode == 0 ? 512 : (mode == 1 ? 5 : 100), reps = mode == 0 ? 1 : (mode == 1 ? 100 : 5);
      const long long rpc = mode == 0 ? 0 : tiles * 128;
      tma_stream_bench<<<148, 64, stages * 16384 + 2048>>>(map, tiles, reps, stages, rpc, out); cudaDeviceSynchronize();
      tma_stream_bench<<<148, 64, stages * 16384 + 2048>>>(map, tiles, reps, stages, rpc, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("stream bench failed\n"); return 1; }
      const double n = (double)tiles * reps * 6;
      printf("TMA stream mode %d stages=%d: %.0f clk per 16 KB load -> %.1f B/clk/SM\n", mode, stages, out[0] / n, 16384.0 * n / out[0]);
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
