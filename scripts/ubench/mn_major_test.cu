// Correctness probe (GPU box only): tcgen05.mma with MN-MAJOR shared-memory operands.  The weight-gradient contractions
// dW = dY^T X contract over the batch rows, i.e. over the OUTER dimension of the row-major activations; fed K-major they
// need transposed copies of every activation.  Here both operands are loaded as they lie in memory ([k][m] and [k][n]
// row-major, TMA boxes of 64 k-rows x 64 elements, 128-byte swizzle) and described to the MMA as MN-major.
// Tries the candidate (leading, stride) byte offsets of the shared-memory descriptor and prints the max error of
// D[m][n] = sum_k A[k][m] B[k][n] against the host.    mn_major_test [K]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../two-towers-overlords_b200/csrc/tt_ptx.cuh"
#include "../../two-towers-overlords_b200/csrc/tt_tma.cuh"
namespace tt { void set_error(const char* f, ...) { fprintf(stderr, "%s\n", f); } int check_cuda(cudaError_t e, const char*, const char*, int) { return e != cudaSuccess; } void note_launch() {} }
using namespace tt::ptx;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// one CTA: M = N = 128, K = 64 * kblocks; operands [K][128] row-major
__global__ void __launch_bounds__(128, 1) mn_major_kernel(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb,
                                                          int kblocks, uint32_t lbo, uint32_t sbo, uint32_t kadv, float* D) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t full, done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&full, 1); mbar_init(&done, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 128); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  uint8_t* sa = smem;            // A: two boxes [64 k][64 m] (m 0..63 | m 64..127), 8 KB each
  uint8_t* sb = smem + 16384;    // B likewise
  // instruction descriptor: bf16 x bf16 -> f32, A and B MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
  const uint32_t idesc = make_idesc_bf16(128, 128) | (1u << 15) | (1u << 16);
  uint32_t ph = 0;
  for (int kb = 0; kb < kblocks; ++kb) {
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(&full, 32768);
      tma_load_2d(sa, &ma, &full, 0, kb * 64);
      tma_load_2d(sa + 8192, &ma, &full, 64, kb * 64);
      tma_load_2d(sb, &mb, &full, 0, kb * 64);
      tma_load_2d(sb + 8192, &mb, &full, 64, kb * 64);
      mbar_wait(&full, ph);
      tc_fence_after();
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = desc_sw128(smem_u32(sa) + kk * kadv, lbo, sbo), db = desc_sw128(smem_u32(sb) + kk * kadv, lbo, sbo);
        mma_bf16(tb, da, db, idesc, (kb | kk) != 0);
      }
      mma_commit(&done);
      mbar_wait(&done, ph);
    }
    ph ^= 1u;
    __syncthreads();
  }
  tc_fence_after();
  for (int ch = 0; ch < 8; ++ch) {
    float v[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + ch * 16, v);
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 128 + ch * 16 + j] = v[j];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 128);
}

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 128;
  std::vector<__nv_bfloat16> A((size_t)K * 128), B((size_t)K * 128);
  std::vector<float> Af(A.size()), Bf(B.size());
  srand(1);
  for (size_t i = 0; i < A.size(); ++i) { A[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); Af[i] = __bfloat162float(A[i]); }
  for (size_t i = 0; i < B.size(); ++i) { B[i] = __float2bfloat16((rand() % 13 - 6) / 4.f); Bf[i] = __bfloat162float(B[i]); }
  std::vector<float> ref(128 * 128, 0.f);
  for (int k = 0; k < K; ++k)
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) ref[m * 128 + n] += Af[(size_t)k * 128 + m] * Bf[(size_t)k * 128 + n];
  __nv_bfloat16 *dA, *dB;
  float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * 128 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  if (tt::make_map_bf16_kmajor(&ma, dA, K, 128, 128, 64) || tt::make_map_bf16_kmajor(&mb, dB, K, 128, 128, 64)) return 1;
  cudaFuncSetAttribute(mn_major_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  const uint32_t cand[][3] = {{8192, 1024, 2048}, {1024, 8192, 2048}, {8192, 1024, 32}, {1024, 8192, 32}, {16, 1024, 2048},
                              {8192, 128, 2048}, {128, 8192, 2048}};
  for (auto& c : cand) {
    cudaMemset(dD, 0, 128 * 128 * 4);
    mn_major_kernel<<<1, 128, 40 * 1024>>>(ma, mb, K / 64, c[0], c[1], c[2], dD);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(128 * 128);
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    for (size_t i = 0; i < out.size(); ++i) err = fmax(err, fabs(out[i] - ref[i]));
    printf("LBO %5u SBO %5u k-advance %4u: %s max|err| %.4g\n", c[0], c[1], c[2], e == cudaSuccess ? "ok" : cudaGetErrorString(e), err);
    if (e != cudaSuccess) return 2;
  }
  return 0;
}
