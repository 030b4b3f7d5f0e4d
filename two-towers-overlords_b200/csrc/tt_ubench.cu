// Measurement probes exported beside the hot path (bench.py's roofline denominators that MEASURED_PEAKS.json does
// not carry).  tt_ubench_l2_read: every SM streams a cache-resident buffer with coalesced 16-byte ld.global.cg loads
// at full occupancy — the L2 -> SM delivery ceiling that bounds the pooled gather, whose two 46.9 MB token tables
// stay resident in the 126 MB L2 (the reference's counterpart of the gather is nn.Embedding inside
// backend/model.py:51-52).
#include "tt_ptx.cuh"
#include "tt_tma.cuh"

namespace tt {
namespace {

__global__ void __launch_bounds__(256) l2_read_kernel(const uint4* __restrict__ buf, size_t n16, int iters,
                                                      unsigned* __restrict__ sink) {
  const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x, stride = (size_t)gridDim.x * 256;
  uint4 acc = make_uint4(0u, 0u, 0u, 0u);
  for (int it = 0; it < iters; ++it) {
    size_t i = tid;
    for (; i + 3 * stride < n16; i += 4 * stride) {  // four independent 16-byte loads in flight per thread
      const uint4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride),
                  d = __ldcg(buf + i + 3 * stride);
      acc.x ^= a.x ^ b.x ^ c.x ^ d.x;
      acc.y ^= a.y ^ b.y ^ c.y ^ d.y;
      acc.z ^= a.z ^ b.z ^ c.z ^ d.z;
      acc.w ^= a.w ^ b.w ^ c.w ^ d.w;
    }
    for (; i < n16; i += stride) {
      const uint4 a = __ldcg(buf + i);
      acc.x ^= a.x; acc.y ^= a.y; acc.z ^= a.z; acc.w ^= a.w;
    }
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) *sink = acc.x;  // keeps the loads alive; practically never taken
}

// Self-test of the MN-major operand path the persistent chain kernel relies on (its weight gradients contract over
// the OUTER dimension of row-major activations): one CTA, D[128,128] = sum_k A[k][m] B[k][n] with both operands loaded
// as they lie in memory ([k][128] row-major, TMA boxes of 64 k-rows x 64 elements, 128-byte swizzle) and described
// to tcgen05.mma as MN-major.  The test compares D with a host product: exact for small-integer inputs.
__global__ void __launch_bounds__(128, 1) mn_major_selftest_kernel(const __grid_constant__ CUtensorMap ma,
                                                                   const __grid_constant__ CUtensorMap mb, int kblocks,
                                                                   float* __restrict__ D) {
  using namespace ptx;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t full, done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full, 1);
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  uint8_t* sa = smem;          // A: boxes [64 k][64 m] for m 0..63 | 64..127, 8 KB each
  uint8_t* sb = smem + 16384;  // B likewise
  const uint32_t idesc = make_idesc_bf16(128, 128) | (1u << 15) | (1u << 16);  // both operands MN-major
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
      const uint64_t da = make_smem_desc_sw128_mn(smem_u32(sa), 8192), db = make_smem_desc_sw128_mn(smem_u32(sb), 8192);
      for (int kk = 0; kk < 4; ++kk) mma_bf16(tb, da + 128 * kk, db + 128 * kk, idesc, (kb | kk) != 0);
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
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 128);
}

}  // namespace
}  // namespace tt

extern "C" int tt_selftest_mn_major(const void* A, const void* B, int K, float* D, tt_stream_t stream) {
  TT_REQUIRE(A && B && D && K >= 64 && K % 64 == 0, "tt_selftest_mn_major: A, B [K,128] bf16 with K a multiple of 64");
  CUtensorMap ma, mb;
  int rc;
  if ((rc = tt::make_map_bf16_kmajor(&ma, A, (uint64_t)K, 128, 128, 64))) return rc;
  if ((rc = tt::make_map_bf16_kmajor(&mb, B, (uint64_t)K, 128, 128, 64))) return rc;
  TT_CUDA(cudaFuncSetAttribute(tt::mn_major_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
  tt::mn_major_selftest_kernel<<<1, 128, 40 * 1024, tt::as_stream(stream)>>>(ma, mb, K / 64, D);
  TT_LAUNCH_CHECK();
  return 0;
}

extern "C" int tt_ubench_l2_read(const void* buf, size_t bytes, int iters, int ctas_per_sm, void* sink,
                                 tt_stream_t stream) {
  TT_REQUIRE(buf && sink && bytes >= 16 && iters >= 1, "tt_ubench_l2_read: bad arguments");
  TT_REQUIRE((reinterpret_cast<uintptr_t>(buf) & 15) == 0, "tt_ubench_l2_read: buffer must be 16-byte aligned");
  const int per_sm = ctas_per_sm > 0 ? ctas_per_sm : 8;  // 8 x 256 threads = a full SM
  tt::l2_read_kernel<<<tt::sm_count() * per_sm, 256, 0, tt::as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(buf), bytes / 16, iters, reinterpret_cast<unsigned*>(sink));
  TT_LAUNCH_CHECK();
  return 0;
}
