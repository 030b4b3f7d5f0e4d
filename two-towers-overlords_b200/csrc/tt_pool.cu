// Pooled gather: token-row gather + attention-masked mean + L2 normalise, one CTA per sequence.
// Replaces backend/model.py:48-56,63-72 of the reference (see include/tt_b200.h).
//
// HBM-bound byte work: each unmasked token costs one H-wide row read (H*4 B fp32 / H*2 B bf16).
// A warp reads one row per step as fully coalesced 16 B (fp32) / 8 B (bf16) per-lane vectors,
// UNROLL rows in flight per warp, 4 warps per sequence, fp32 accumulation in registers, cross-warp
// reduction staged in shared memory, warp-shuffle reduction for the norm.
#include <stdlib.h>

#include "tt_pool.cuh"

namespace tt {

namespace {

constexpr int kPoolThreads = 128;
constexpr int kPoolWarps = kPoolThreads / 32;
constexpr int kMaxL = 512;  // tokenizer max_length at backend/model.py:44

template <typename TE>
struct RowVec;
template <>
struct RowVec<float> {
  using V = float4;
  static __device__ __forceinline__ V zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ V load(const float* row, int vec) {
    return __ldg(reinterpret_cast<const float4*>(row) + vec);
  }
  static __device__ __forceinline__ void fma(float w, const V& v, float* acc) {
    acc[0] = fmaf(w, v.x, acc[0]);
    acc[1] = fmaf(w, v.y, acc[1]);
    acc[2] = fmaf(w, v.z, acc[2]);
    acc[3] = fmaf(w, v.w, acc[3]);
  }
};
template <>
struct RowVec<__nv_bfloat16> {
  using V = uint2;  // 4 bf16
  static __device__ __forceinline__ V zero() { return make_uint2(0u, 0u); }
  static __device__ __forceinline__ V load(const __nv_bfloat16* row, int vec) {
    return __ldg(reinterpret_cast<const uint2*>(row) + vec);
  }
  static __device__ __forceinline__ void fma(float w, const V& v, float* acc) {
    acc[0] = fmaf(w, __uint_as_float(v.x << 16), acc[0]);
    acc[1] = fmaf(w, __uint_as_float(v.x & 0xffff0000u), acc[1]);
    acc[2] = fmaf(w, __uint_as_float(v.y << 16), acc[2]);
    acc[3] = fmaf(w, __uint_as_float(v.y & 0xffff0000u), acc[3]);
  }
};

template <typename TE, int NV, int UNROLL, bool PIPE>
__global__ void __launch_bounds__(kPoolThreads) pool_fwd_kernel(const PoolParams p) {
  constexpr int H = NV * 128;
  __shared__ unsigned s_row[kMaxL];
  __shared__ float s_w[kMaxL];
  __shared__ __align__(16) float s_red[kPoolWarps][H];
  __shared__ int s_wc[kPoolWarps];
  __shared__ float s_fred[kPoolWarps];

  // a programmatic dependent (the persistent chain kernel of a whole-step call) may start filling SMs as soon as the
  // last wave of this grid is resident; it waits for this grid's completion before it reads xhat
  pdl_launch();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // block -> (segment, sequence); host orders segments longest-first so short ones fill the tail
  int b = blockIdx.x, s = 0;
  while (s + 1 < p.nseg && b >= p.seg[s].B) {
    b -= p.seg[s].B;
    ++s;
  }
  const PoolSegDev sg = p.seg[s];
  const int L = sg.L;
  const size_t tok0 = (size_t)b * L;
  const TE* __restrict__ table = reinterpret_cast<const TE*>(sg.table);

  // ---- 1. compact the unmasked tokens of this sequence (order kept -> deterministic sums) ----
  int n_valid = 0;
  float cnt_local = 0.f;
  for (int base = 0; base < L; base += kPoolThreads) {
    const int t = base + tid;
    long long id = 0;
    float w = 0.f;
    if (t < L) {
      w = (float)load_index(sg.mask, p.mask_dtype, tok0 + t);
      if (w != 0.f) {
        id = load_index(sg.ids, p.ids_dtype, tok0 + t);
        if (id < 0 || id >= p.vocab) {
          if (p.err) atomicExch(p.err, 1);
          w = 0.f;
          id = 0;
        }
      }
    }
    cnt_local += w;
    const bool valid = (w != 0.f);
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) s_wc[warp] = __popc(bal);
    __syncthreads();
    int off = n_valid, tot = 0;
#pragma unroll
    for (int i = 0; i < kPoolWarps; ++i) {
      const int c = s_wc[i];
      if (i < warp) off += c;
      tot += c;
    }
    if (valid) {
      const int pos = off + __popc(bal & ((1u << lane) - 1u));
      s_row[pos] = (unsigned)id;
      s_w[pos] = w;
    }
    n_valid += tot;
    __syncthreads();
  }
  // count = sum of mask values (model.py:71 `input_mask_expanded.sum(1)`)
  cnt_local = warp_sum(cnt_local);
  if (lane == 0) s_fred[warp] = cnt_local;

  // ---- 2. stream the rows: warp w takes entries w, w+4, ...; UNROLL rows in flight ----------
  float acc[NV * 4];
#pragma unroll
  for (int i = 0; i < NV * 4; ++i) acc[i] = 0.f;

  using RV = RowVec<TE>;
  // Software-pipelined in registers: the UNROLL*NV row loads of batch i+1 are issued before the FMAs of
  // batch i, so a warp always has one whole batch in flight and the FMAs only ever wait on loads issued a
  // full iteration earlier.  A padding entry re-reads row s_row[0] with weight 0 instead of predicating
  // the load, which keeps each batch one unconditional run of loads.
  typename RV::V v[UNROLL][NV];
  float wt[UNROLL];
  auto issue = [&](int e0, typename RV::V (&dst)[UNROLL][NV], float (&w)[UNROLL]) {
    const TE* rows[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int e = e0 + u * kPoolWarps;
      const bool ok = e < n_valid;
      w[u] = ok ? s_w[e] : 0.f;
      rows[u] = table + (size_t)s_row[ok ? e : 0] * H;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int j = 0; j < NV; ++j) dst[u][j] = RV::load(rows[u], j * 32 + lane);
  };
  if (PIPE) {
    if (warp < n_valid) issue(warp, v, wt);
    for (int e0 = warp; e0 < n_valid; e0 += kPoolWarps * UNROLL) {
      typename RV::V vn[UNROLL][NV];
      float wn[UNROLL];
      const int e1 = e0 + kPoolWarps * UNROLL;
      const bool more = e1 < n_valid;
      if (more) issue(e1, vn, wn);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int j = 0; j < NV; ++j) RV::fma(wt[u], v[u][j], acc + j * 4);
      if (more) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          wt[u] = wn[u];
#pragma unroll
          for (int j = 0; j < NV; ++j) v[u][j] = vn[u][j];
        }
      }
    }
  } else {  // one batch of UNROLL rows in flight per warp; latency is hidden by the other resident warps
    for (int e0 = warp; e0 < n_valid; e0 += kPoolWarps * UNROLL) {
      issue(e0, v, wt);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int j = 0; j < NV; ++j) RV::fma(wt[u], v[u][j], acc + j * 4);
    }
  }

  // ---- 3. cross-warp reduction staged in shared memory --------------------------------------
#pragma unroll
  for (int j = 0; j < NV; ++j)
    *reinterpret_cast<float4*>(&s_red[warp][(j * 32 + lane) * 4]) =
        make_float4(acc[j * 4 + 0], acc[j * 4 + 1], acc[j * 4 + 2], acc[j * 4 + 3]);
  __syncthreads();

  float cnt = 0.f;
#pragma unroll
  for (int i = 0; i < kPoolWarps; ++i) cnt += s_fred[i];
  const float denom = fmaxf(cnt, 1e-9f);  // clamp(min=1e-9), model.py:70-72

  float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
  float sq = 0.f;
  const bool owner = tid < NV * 32;  // thread owns columns [4*tid, 4*tid+4)
  if (owner) {
#pragma unroll
    for (int i = 0; i < kPoolWarps; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(&s_red[i][tid * 4]);
      m.x += t.x; m.y += t.y; m.z += t.z; m.w += t.w;
    }
    m.x = m.x / denom; m.y = m.y / denom; m.z = m.z / denom; m.w = m.w / denom;
    sq = m.x * m.x + m.y * m.y + m.z * m.z + m.w * m.w;
  }
  __syncthreads();  // s_fred is reused below
  sq = warp_sum(sq);
  if (lane == 0) s_fred[warp] = sq;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < kPoolWarps; ++i) tot += s_fred[i];
  const float nrm = sqrtf(tot);
  const float dn = fmaxf(nrm, 1e-12f);  // F.normalize eps, model.py:56

  const int row_out = sg.row0 + b;
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  if (owner) x = make_float4(m.x / dn, m.y / dn, m.z / dn, m.w / dn);
  // every output of one pooled row (owner threads: 4 columns each; thread 0: the count and the norm)
  auto emit_row = [&](int row) {
    if (owner) {
      *reinterpret_cast<float4*>(p.xhat + (size_t)row * H + tid * 4) = x;
      if (p.x_hi) {  // operands of the tensor-core projection: bf16 hi/lo split, row-major + transposed
        const float xv[4] = {x.x, x.y, x.z, x.w};
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) split_bf16(xv[c], hi[c], lo[c]);
        *reinterpret_cast<uint2*>(p.x_hi + (size_t)row * H + tid * 4) = *reinterpret_cast<uint2*>(hi);
        *reinterpret_cast<uint2*>(p.x_lo + (size_t)row * H + tid * 4) = *reinterpret_cast<uint2*>(lo);
        if (p.x_lo2) {
          __nv_bfloat16 l2[4];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            l2[c] = __float2bfloat16_rn((xv[c] - __bfloat162float(hi[c])) - __bfloat162float(lo[c]));
          *reinterpret_cast<uint2*>(p.x_lo2 + (size_t)row * H + tid * 4) = *reinterpret_cast<uint2*>(l2);
        }
        if (p.xt_hi) {
          const int tcol = row + (row >= p.t_split_row ? p.t_shift : 0);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            p.xt_hi[(size_t)(tid * 4 + c) * p.ldt + tcol] = hi[c];
            p.xt_lo[(size_t)(tid * 4 + c) * p.ldt + tcol] = lo[c];
          }
        }
      }
    }
    if (tid == 0) {
      if (p.cnt) p.cnt[row] = cnt;
      if (p.nrm) p.nrm[row] = nrm;
    }
  };
  emit_row(row_out);
  if (p.alias && (int)blockIdx.x < p.alias_n && tid == 0) {  // every alias index is range-checked by some CTA
    const int a = __ldg(p.alias + blockIdx.x);
    if ((a < 0 || a >= p.alias_n) && p.err) atomicExch(p.err, 1);
  }
  // In-batch negatives (backend/data.py:113-137: a negative IS another item's positive document): rows that alias
  // this one get the same bits instead of a second gather of the same tokens.
  if (p.alias && row_out >= p.alias_src_row0 && row_out < p.alias_src_row0 + p.alias_n) {
    constexpr int kMatchCap = 32;
    __shared__ int s_match[kMatchCap];
    __shared__ int s_nmatch;
    const int j = row_out - p.alias_src_row0;
    if (tid == 0) s_nmatch = 0;
    __syncthreads();
    for (int i = tid; i < p.alias_n; i += kPoolThreads)
      if (__ldg(p.alias + i) == j) {
        const int k = atomicAdd(&s_nmatch, 1);
        if (k < kMatchCap) s_match[k] = i;
      }
    __syncthreads();
    const int nm = s_nmatch;
    if (nm <= kMatchCap) {
      for (int k = 0; k < nm; ++k) emit_row(p.alias_dst_row0 + s_match[k]);
    } else {  // (a document drawn as the negative of more than 32 items of one batch: plain rescan, CTA-uniform)
      for (int i = 0; i < p.alias_n; ++i)
        if (__ldg(p.alias + i) == j) emit_row(p.alias_dst_row0 + i);
    }
  }
}

// Rows in flight per warp.  TT_POOL_VARIANT (tuning hook): 0 = register-pipelined x4, 1 = x4 (36 warps per SM),
// 2 = x8, 3 = pipelined x2.  Default: x8 for fp32 tables when the gather has the SMs to itself (whole-step calls:
// 100.5 vs 103 us per launch, 0.2373 vs 0.2407 ms per configs[1] step), x4 when it shares them with other kernels
// (more, lighter warps interleave better) and for bf16 tables.
int pool_variant(bool share_sm, bool fp32_table) {
  static int env = -2;
  if (env == -2) {
    const char* e = getenv("TT_POOL_VARIANT");
    env = e ? atoi(e) : -1;
  }
  if (env >= 0) return env;
  return (!share_sm && fp32_table) ? 2 : 1;
}

// Unused dynamic shared memory that caps how many gather CTAs an SM holds.  When the gather of step i+1 runs beside the
// tensor-core chain of step i (p.share_sm), an SM full of gather CTAs (36 warps, ~62 K registers) leaves no room for
// a GEMM CTA until a whole wave of gathers has retired: the chain — the critical path — then stalls ~60 us per step
// (torch.profiler timeline).  Fewer resident gather CTAs cost the gather some latency hiding, which does not matter
// because it runs in the chain's shadow.
size_t pool_pad_smem(bool share_sm) {
  static int pad = -1;
  if (pad < 0) {
    const char* e = getenv("TT_POOL_PAD_SMEM");  // tuning hook (bytes, <= 40960)
    pad = e ? atoi(e) : kPoolSharePadBytes;
    if (pad < 0) pad = 0;
    if (pad > 40960) pad = 40960;
  }
  return share_sm ? (size_t)pad : 0;
}

template <typename TE, int NV>
int launch_pool(const PoolParams& p, int total, cudaStream_t st) {
  constexpr int UNROLL = sizeof(TE) == 4 ? 4 : 8;
  const size_t pad = pool_pad_smem(p.share_sm != 0);
  switch (pool_variant(p.share_sm != 0, sizeof(TE) == 4)) {
    case 1: pool_fwd_kernel<TE, NV, UNROLL, false><<<total, kPoolThreads, pad, st>>>(p); break;
    case 2: pool_fwd_kernel<TE, NV, UNROLL * 2, false><<<total, kPoolThreads, pad, st>>>(p); break;
    case 3: pool_fwd_kernel<TE, NV, UNROLL / 2, true><<<total, kPoolThreads, pad, st>>>(p); break;
    default: pool_fwd_kernel<TE, NV, UNROLL, true><<<total, kPoolThreads, pad, st>>>(p); break;
  }
  TT_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int pool_fwd_launch(PoolParams p, int table_dtype, int H, cudaStream_t st) {
  TT_REQUIRE(p.nseg >= 1 && p.nseg <= 4, "tt_pool_fwd: nseg must be in [1,4], got %d", p.nseg);
  TT_REQUIRE(H == 128 || H == 256 || H == 384 || H == 768,
             "tt_pool_fwd: hidden size %d not supported (128, 256, 384, 768)", H);
  TT_REQUIRE(table_dtype == TT_F32 || table_dtype == TT_BF16, "tt_pool_fwd: table dtype must be f32 or bf16");
  TT_REQUIRE(p.ids_dtype == TT_I64 || p.ids_dtype == TT_I32 || p.ids_dtype == TT_U16,
             "tt_pool_fwd: ids dtype must be i64, i32 or u16");
  TT_REQUIRE(p.mask_dtype == TT_I64 || p.mask_dtype == TT_I32 || p.mask_dtype == TT_U8,
             "tt_pool_fwd: mask dtype must be i64, i32 or u8");
  // longest segments first (insertion sort, <= 4 entries)
  for (int i = 1; i < p.nseg; ++i)
    for (int j = i; j > 0 && p.seg[j].L > p.seg[j - 1].L; --j) {
      PoolSegDev t = p.seg[j];
      p.seg[j] = p.seg[j - 1];
      p.seg[j - 1] = t;
    }
  long long total = 0;
  for (int i = 0; i < p.nseg; ++i) {
    TT_REQUIRE(p.seg[i].L >= 1 && p.seg[i].L <= kMaxL, "tt_pool_fwd: L=%d outside [1,%d]", p.seg[i].L, kMaxL);
    TT_REQUIRE(p.seg[i].B >= 0, "tt_pool_fwd: negative batch");
    total += p.seg[i].B;
  }
  if (total == 0) return 0;
  const int NV = H / 128;
#define TT_POOL_CASE(TE)                                        \
  switch (NV) {                                                 \
    case 1: return launch_pool<TE, 1>(p, (int)total, st);       \
    case 2: return launch_pool<TE, 2>(p, (int)total, st);       \
    case 3: return launch_pool<TE, 3>(p, (int)total, st);       \
    default: return launch_pool<TE, 6>(p, (int)total, st);      \
  }
  if (table_dtype == TT_F32) { TT_POOL_CASE(float) }
  TT_POOL_CASE(__nv_bfloat16)
#undef TT_POOL_CASE
}

}  // namespace tt

extern "C" int tt_pool_fwd_multi(const tt_pool_seg* segs, int nseg, int table_dtype, int vocab, int H,
                                 int ids_dtype, int mask_dtype, float* xhat, float* cnt, float* nrm,
                                 int* err_flag, tt_stream_t stream) {
  TT_REQUIRE(segs != nullptr && nseg >= 1 && nseg <= 4, "tt_pool_fwd_multi: bad segment list");
  tt::PoolParams p{};
  p.nseg = nseg;
  for (int i = 0; i < nseg; ++i) {
    p.seg[i].table = segs[i].table;
    p.seg[i].ids = segs[i].ids;
    p.seg[i].mask = segs[i].mask;
    p.seg[i].B = segs[i].B;
    p.seg[i].L = segs[i].L;
    p.seg[i].row0 = segs[i].row0;
  }
  p.ids_dtype = ids_dtype;
  p.mask_dtype = mask_dtype;
  p.vocab = vocab;
  p.xhat = xhat;
  p.cnt = cnt;
  p.nrm = nrm;
  p.err = err_flag;
  return tt::pool_fwd_launch(p, table_dtype, H, tt::as_stream(stream));
}

extern "C" int tt_pool_fwd(const void* table, int table_dtype, int vocab, int H, const void* ids,
                           int ids_dtype, const void* mask, int mask_dtype, int B, int L, float* xhat,
                           float* cnt, float* nrm, int* err_flag, tt_stream_t stream) {
  tt_pool_seg s{table, ids, mask, B, L, 0, 0};
  return tt_pool_fwd_multi(&s, 1, table_dtype, vocab, H, ids_dtype, mask_dtype, xhat, cnt, nrm, err_flag,
                           stream);
}
