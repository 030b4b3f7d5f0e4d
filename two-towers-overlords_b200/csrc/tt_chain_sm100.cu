// The projection chain of one training step as ONE persistent kernel (north_star (2): the tower MLP on tcgen05 with
// cosine distance + margin loss fused into the second layer's epilogue, forward and backward).  Replaces
// backend/model.py:33-38,59,132-145 and what autograd does for them inside backend/training.py:40-51.
//
// One CTA per SM walks a static queue of tile tasks in dependency order; kernel boundaries of the per-kernel chain
// (tt_gemm_sm100.cu: split -> fwd1 -> fwd2 -> loss -> dY^T -> dz1 -> dW2 -> colsums -> dW1 -> reduce) become
// per-row-tile ready counters in global memory:
//
//   S    weights -> bf16 terms (+ transposes)                            SIMT, one task per CTA
//   F1   h  = relu(x W1^T + b1)        -> h fp32, (hi,lo), transposed     tile (row tile, col tile)
//   F2L  y  = h W2^T + b2 for the q | p | n rows of 128 triplets (three accumulators), then in the SAME epilogue:
//        partial |q|^2,|p|^2,|n|^2,q.p,q.n over this tile's columns -> exchanged with the sibling column tiles through
//        global memory -> cosines, hinge, closed-form dY straight from TMEM -> dY (hi,lo), transposed, db2 partials
//   DZ   dz1 = (dY W2) * (h > 0)       -> transposed (hi,lo), db1 partials [, (hi,lo) for DX]
//   DX   dxhat = dz1 W1                (only when the token tables train)
//   DW2  dW2 partial = dY^T h  over a chunk of batch rows                 raw fp32 partial tile
//   DW1  dW1 partial = dz1^T x over a chunk of batch rows
//   G    gradients = fixed-order sums of the partials; loss; Adam step-count advance
//
// Warp roles per CTA: warp 0 = TMA producer (also waits for a task's dependencies), warp 1 = TMEM owner + single-thread
// tcgen05.mma issuer, warps 2-9 = epilogue (two warps per TMEM lane quarter, 64 columns each).  The accumulators form
// a ring of four 128-column TMEM buffers, so an epilogue overlaps the following main loops.  All sums have a fixed
// order: results are bit-reproducible run to run.
#include <stdlib.h>

#include "tt_ptx.cuh"
#include "tt_step_ws.cuh"
#include "tt_tma.cuh"

namespace tt {

namespace {

using namespace ptx;

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kMaxStages = 6;
constexpr int kAcc = 4;  // TMEM accumulator ring
constexpr int kEpiThreads = 256;
constexpr int kThreads = 64 + kEpiThreads;
constexpr uint32_t kABytes = BM * BK * 2, kBBytes = BN * BK * 2, kStageBytes = kABytes + kBBytes;
constexpr int kStagingBytes = 2 * 8 * 64 * 4 /*column sums*/ + 2 * 32 * 33 * 2 /*split tile*/ + 64;
constexpr size_t chain_smem(int stages) { return (size_t)stages * kStageBytes + 1024 + 256 + kStagingBytes; }
constexpr float kCosEps = 1e-8f;  // torch.cosine_similarity eps (per-vector clamp), backend/model.py:134
constexpr long long kSpinLimit = 4000000000ll;

enum { T_S = 0, T_F1, T_F2L, T_DZ, T_DX, T_DW2, T_DW1, T_G, T_COUNT };

struct TowerMaps {  // K-major bf16 operand terms of one tower (box 128 rows x 64 k)
  CUtensorMap x[3], w1[3];   // F1:  [rows,H] x [P,H]
  CUtensorMap h[2], w2[2];   // F2:  [rows,P] x [P,P]
  CUtensorMap dy[2], w2t[2];  // DZ:  [rows,P] x [P,P]^T
  CUtensorMap dyt[2], ht[2];  // DW2: [P,rows] x [P,rows]
  CUtensorMap dzt[2], xt[2];  // DW1: [P,rows] x [H,rows]
  CUtensorMap dz[2], w1t[2];  // DX:  [rows,P] x [H,P]
};

struct SplitJobC {
  const float* X;
  bf16 *hi, *lo, *lo2, *thi, *tlo;
  int R, C, ldt, tile0, tiles_c;
};

struct alignas(64) ChainParams {
  TowerMaps tm[2];
  int B, H, P, RTB, NC, NCH, ldt, dcol;
  int pairs, terms, pairs1, terms1;
  int kcb, nch[2], kbt[2];
  int off[T_COUNT + 1];
  int stages, n_sj, n_split_tiles;
  SplitJobC sj[6];
  float margin, inv_batch, grad_scale;
  const float *b1[2], *b2[2];
  float *h, *y, *dy, *dz1, *dxhat, *stats, *loss;
  bf16 *h_hi, *h_lo, *ht_hi, *ht_lo, *dy_hi, *dy_lo, *dyt_hi, *dyt_lo, *dz_hi, *dz_lo, *dzt_hi, *dzt_lo;
  float *part2, *part1;
  float *dW1[2], *db1[2], *dW2[2], *db2[2];
  float *stat_part, *cs1, *cs2, *hinge_part;
  unsigned* ctr;
  // optional fused optimiser: torch.optim.Adam on the flat parameter buffer the 8 tensors are slices of
  double* adam_state;  // {t, beta1^t, beta2^t, -}: advanced once per launch (by the first task), read by the tail
  float *adam_p, *adam_g, *adam_m, *adam_v;
  float lr, beta1, beta2, eps;
  unsigned total_signals;
};

// ---- cross-CTA hand-off ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void signal(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned target) {
  if (ld_acquire(p) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire(p) < target) {
    __nanosleep(64);
    if (clock64() - t0 > kSpinLimit) __trap();  // a protocol bug fails the launch instead of hanging the GPU
  }
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct Task {
  int type, i;
};
__device__ __forceinline__ Task decode(const ChainParams& p, int idx) {
  int t = 0;
#pragma unroll
  for (int k = 1; k < T_COUNT; ++k)
    if (idx >= p.off[k]) t = k;
  return Task{t, idx - p.off[t]};
}

struct RowTile {
  int r, seg, t, trow, grow, valid;
};
// row tile u = 3 r + seg: 128 rows of segment seg (q | p | n) starting at triplet 128 r
__device__ __forceinline__ RowTile row_tile(const ChainParams& p, int u) {
  RowTile rt;
  rt.r = u / 3;
  rt.seg = u - 3 * rt.r;
  rt.t = rt.seg != 0;
  rt.trow = (rt.seg == 2 ? p.B : 0) + rt.r * BM;  // row inside the tower's operand (document tower: p rows, n rows)
  rt.grow = rt.seg * p.B + rt.r * BM;             // row inside the [3B, .] activations
  rt.valid = min(BM, p.B - rt.r * BM);
  return rt;
}

struct Sub {
  const CUtensorMap *a, *b;
  int terms, pairs, m0, n0, kb0, nkb;
};
__device__ __forceinline__ int n_subs(int type) { return type == T_F2L ? 3 : ((type == T_S || type == T_G) ? 0 : 1); }

__device__ __forceinline__ void dw_chunk(const ChainParams& p, int ci, int& t, int& j) {
  t = ci >= p.nch[0];
  j = ci - (t ? p.nch[0] : 0);
}

__device__ __forceinline__ Sub get_sub(const ChainParams& p, Task tk, int k) {
  Sub s;
  s.terms = p.terms;
  s.pairs = p.pairs;
  s.kb0 = 0;
  switch (tk.type) {
    case T_F1: {
      const RowTile rt = row_tile(p, tk.i / p.NC);
      s.a = p.tm[rt.t].x; s.b = p.tm[rt.t].w1;
      s.terms = p.terms1; s.pairs = p.pairs1;
      s.m0 = rt.trow; s.n0 = (tk.i % p.NC) * BN; s.nkb = p.H / BK;
    } break;
    case T_F2L: {
      const int r = tk.i / p.NC, t = k != 0;
      s.a = p.tm[t].h; s.b = p.tm[t].w2;
      s.m0 = (k == 2 ? p.B : 0) + r * BM; s.n0 = (tk.i % p.NC) * BN; s.nkb = p.P / BK;
    } break;
    case T_DZ: {
      const RowTile rt = row_tile(p, tk.i / p.NC);
      s.a = p.tm[rt.t].dy; s.b = p.tm[rt.t].w2t;
      s.m0 = rt.trow; s.n0 = (tk.i % p.NC) * BN; s.nkb = p.P / BK;
    } break;
    case T_DX: {
      const RowTile rt = row_tile(p, tk.i / p.NCH);
      s.a = p.tm[rt.t].dz; s.b = p.tm[rt.t].w1t;
      s.m0 = rt.trow; s.n0 = (tk.i % p.NCH) * BN; s.nkb = p.P / BK;
    } break;
    case T_DW2: {
      const int nt_ = p.NC * p.NC, tile = tk.i % nt_;
      int t, j;
      dw_chunk(p, tk.i / nt_, t, j);
      s.a = p.tm[t].dyt; s.b = p.tm[t].ht;
      s.m0 = (tile / p.NC) * BM; s.n0 = (tile % p.NC) * BN;
      s.kb0 = j * p.kcb; s.nkb = min(p.kcb, p.kbt[t] - s.kb0);
    } break;
    default: {  // T_DW1
      const int nt_ = p.NC * p.NCH, tile = tk.i % nt_;
      int t, j;
      dw_chunk(p, tk.i / nt_, t, j);
      s.a = p.tm[t].dzt; s.b = p.tm[t].xt;
      s.m0 = (tile / p.NCH) * BM; s.n0 = (tile % p.NCH) * BN;
      s.kb0 = j * p.kcb; s.nkb = min(p.kcb, p.kbt[t] - s.kb0);
    } break;
  }
  return s;
}

__device__ __forceinline__ unsigned* h_ready(const ChainParams& p) { return p.ctr + 8; }
__device__ __forceinline__ unsigned* dy_ready(const ChainParams& p) { return p.ctr + 8 + 3 * p.RTB; }
__device__ __forceinline__ unsigned* dz_ready(const ChainParams& p) { return p.ctr + 8 + 6 * p.RTB; }
__device__ __forceinline__ unsigned* stat_ready(const ChainParams& p) { return p.ctr + 8 + 9 * p.RTB; }

// every row tile [seg][r] that holds one of the batch rows [k0, k1) of tower t must have all NC column tiles done
__device__ __forceinline__ void wait_rows(const ChainParams& p, const unsigned* ready, int t, int k0, int k1) {
  if (t == 0) {
    for (int r = k0 / BM; r <= (k1 - 1) / BM; ++r) wait_counter(ready + r, p.NC);
    return;
  }
  if (k0 < p.B)
    for (int r = k0 / BM; r <= (min(k1, p.B) - 1) / BM; ++r) wait_counter(ready + p.RTB + r, p.NC);
  if (k1 > p.B)
    for (int r = (max(k0, p.B) - p.B) / BM; r <= (k1 - p.B - 1) / BM; ++r) wait_counter(ready + 2 * p.RTB + r, p.NC);
}

__device__ __forceinline__ void wait_deps(const ChainParams& p, Task tk) {
  switch (tk.type) {
    case T_F1: wait_counter(p.ctr + 0, (unsigned)(p.off[T_F1] - p.off[T_S])); break;
    case T_F2L: {
      const int r = tk.i / p.NC;
      for (int seg = 0; seg < 3; ++seg) wait_counter(h_ready(p) + seg * p.RTB + r, p.NC);
    } break;
    case T_DZ: {
      const RowTile rt = row_tile(p, tk.i / p.NC);
      wait_counter(dy_ready(p) + rt.seg * p.RTB + rt.r, p.NC);
    } break;
    case T_DX: {
      const RowTile rt = row_tile(p, tk.i / p.NCH);
      wait_counter(dz_ready(p) + rt.seg * p.RTB + rt.r, p.NC);
    } break;
    case T_DW2:
    case T_DW1: {
      const int nt_ = tk.type == T_DW2 ? p.NC * p.NC : p.NC * p.NCH;
      int t, j;
      dw_chunk(p, tk.i / nt_, t, j);
      const int k0 = j * p.kcb * BK, k1 = min(k0 + p.kcb * BK, t ? 2 * p.B : p.B);
      wait_rows(p, tk.type == T_DW2 ? dy_ready(p) : dz_ready(p), t, k0, k1);
    } break;
    default: break;
  }
}

// ---- epilogue helpers --------------------------------------------------------------------------------------
// sum over the 32 lanes of each of 16 per-lane values; every lane returns the total of column (lane >> 1) & 15
__device__ __forceinline__ float colsum16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float keep = h16 ? v[8 + j] : v[j], send = h16 ? v[j] : v[8 + j];
    a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float keep = h8 ? a[4 + j] : a[j], send = h8 ? a[j] : a[4 + j];
    b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float keep = h4 ? b[2 + j] : b[j], send = h4 ? b[j] : b[2 + j];
    c[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float keep = h2 ? c[1] : c[0], send = h2 ? c[0] : c[1];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

__device__ __forceinline__ void store16(float* dst, const float (&v)[16]) {
  float4* d = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int j = 0; j < 4; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

// (hi, lo) bf16 terms of 16 values: row-major [.., ld] at `o` and / or transposed [n + j][tcol]
__device__ __forceinline__ void store_terms(const float (&v)[16], bf16* c_hi, bf16* c_lo, size_t o, bf16* t_hi,
                                            bf16* t_lo, int n, int ldt, int tcol) {
  alignas(16) bf16 hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) split_bf16(v[j], hi[j], lo[j]);
  if (c_hi) {
    uint4* dh = reinterpret_cast<uint4*>(c_hi + o);
    uint4* dl = reinterpret_cast<uint4*>(c_lo + o);
    dh[0] = reinterpret_cast<const uint4*>(hi)[0];
    dh[1] = reinterpret_cast<const uint4*>(hi)[1];
    dl[0] = reinterpret_cast<const uint4*>(lo)[0];
    dl[1] = reinterpret_cast<const uint4*>(lo)[1];
  }
  if (t_hi) {  // lanes hold consecutive rows: each store instruction is one coalesced 64 B run per column
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      t_hi[(size_t)(n + j) * ldt + tcol] = hi[j];
      t_lo[(size_t)(n + j) * ldt + tcol] = lo[j];
    }
  }
}

struct EpiCtx {
  uint32_t tmem_base;
  uint64_t *acc_full, *acc_empty;
  float* cs_s;  // [2][8][64]
  bf16* sp_s;   // [2][32][33]
  float* hs_s;  // [4]
  int we, q, hf, lane, tid;  // epilogue warp 0..7, TMEM lane quarter, column half, lane, epilogue thread 0..255
  int ab;
  uint32_t aph;
};

__device__ __forceinline__ uint32_t acc_addr(const EpiCtx& e, int buf, int chunk) {
  return e.tmem_base + ((uint32_t)(e.q * 32) << 16) + (uint32_t)(buf * BN + e.hf * 64 + chunk * 16);
}
__device__ __forceinline__ void acc_advance(EpiCtx& e, int n) {
  e.ab += n;
  if (e.ab >= kAcc) {
    e.ab -= kAcc;
    e.aph ^= 1u;
  }
}
// all writes of this CTA's epilogue visible device-wide (also to TMA reads of other CTAs), then one signal per counter
__device__ __forceinline__ void publish(const EpiCtx& e) {
  fence_proxy_async();
  __threadfence();
  epi_bar();
}

// per-warp column sums of one chunk -> staging; reduce_cs() then adds the four lane quarters in a fixed order
__device__ __forceinline__ void stage_cs(const EpiCtx& e, int slot, int chunk, float d) {
  if ((e.lane & 1) == 0) e.cs_s[(slot * 8 + e.we) * 64 + chunk * 16 + (e.lane >> 1)] = d;
}
__device__ __forceinline__ float reduce_cs(const EpiCtx& e, int slot, int col) {  // col 0..127 of the tile
  const int hf = col >> 6, cc = col & 63;
  float s = 0.f;
#pragma unroll
  for (int quarter = 0; quarter < 4; ++quarter) s += e.cs_s[(slot * 8 + hf * 4 + ((quarter - 2) & 3)) * 64 + cc];
  return s;
}

__device__ void epi_split(const ChainParams& p, EpiCtx& e, int task_i) {
  const int tx = e.tid & 31, ty = e.tid >> 5;  // 32 x 8
  bf16(*s_hi)[33] = reinterpret_cast<bf16(*)[33]>(e.sp_s);
  bf16(*s_lo)[33] = reinterpret_cast<bf16(*)[33]>(e.sp_s + 32 * 33);
  const int stride = p.off[T_F1] - p.off[T_S];
  if (task_i == 0 && e.tid == 0 && p.adam_state) {
    // torch.optim.Adam's step count and beta powers live on the device so that a graph replay advances them; the
    // tail (epi_grad) reads them through the S -> ... -> G dependency chain
    const double t = p.adam_state[0];
    p.adam_state[0] = t + 1.0;
    p.adam_state[1] = (t == 0.0 ? 1.0 : p.adam_state[1]) * (double)p.beta1;
    p.adam_state[2] = (t == 0.0 ? 1.0 : p.adam_state[2]) * (double)p.beta2;
  }
  for (int tile = task_i; tile < p.n_split_tiles; tile += stride) {
    int ji = 0;
    for (int k = 1; k < p.n_sj; ++k)
      if (tile >= p.sj[k].tile0) ji = k;
    const SplitJobC& J = p.sj[ji];
    const int lt = tile - J.tile0;
    const int r0 = (lt / J.tiles_c) * 32, c0 = (lt % J.tiles_c) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + i * 8, c = c0 + tx;
      bf16 h = __float2bfloat16_rn(0.f), l = h;
      if (r < J.R && c < J.C) {
        const float xv = J.X[(size_t)r * J.C + c];
        split_bf16(xv, h, l);
        if (J.hi) {
          J.hi[(size_t)r * J.C + c] = h;
          J.lo[(size_t)r * J.C + c] = l;
        }
        if (J.lo2) J.lo2[(size_t)r * J.C + c] = __float2bfloat16_rn((xv - __bfloat162float(h)) - __bfloat162float(l));
      }
      s_hi[ty + i * 8][tx] = h;
      s_lo[ty + i * 8][tx] = l;
    }
    epi_bar();
    if (J.thi) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + i * 8, r = r0 + tx;
        if (r < J.R && c < J.C) {
          J.thi[(size_t)c * J.ldt + r] = s_hi[tx][ty + i * 8];
          J.tlo[(size_t)c * J.ldt + r] = s_lo[tx][ty + i * 8];
        }
      }
    }
    epi_bar();
  }
  publish(e);
  if (e.tid == 0) signal(p.ctr + 0);
}

__device__ void epi_f1(const ChainParams& p, EpiCtx& e, int task_i) {
  const RowTile rt = row_tile(p, task_i / p.NC);
  const int n0 = (task_i % p.NC) * BN;
  mbar_wait(&e.acc_full[e.ab], e.aph);
  tc_fence_after();
  const int row = e.q * 32 + e.lane;
  const bool ok = row < rt.valid;
  const size_t grow = (size_t)rt.grow + row;
  const int tcol = (rt.t ? p.dcol : 0) + rt.trow + row;
  const float* bias = p.b1[rt.t];
  for (int ch = 0; ch < 4; ++ch) {
    const int n = n0 + e.hf * 64 + ch * 16;
    if (n >= p.P) break;  // warp-uniform
    float v[16];
    tmem_ld16(acc_addr(e, e.ab, ch), v);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + __ldg(bias + n + j), 0.f);
    if (!ok) continue;
    const size_t o = grow * p.P + n;
    store16(p.h + o, v);
    store_terms(v, p.h_hi, p.h_lo, o, p.ht_hi, p.ht_lo, n, p.ldt, tcol);
  }
  tc_fence_before();
  mbar_arrive(&e.acc_empty[e.ab]);
  acc_advance(e, 1);
  publish(e);
  if (e.tid == 0) signal(h_ready(p) + rt.seg * p.RTB + rt.r);
}

__device__ void epi_f2l(const ChainParams& p, EpiCtx& e, int task_i) {
  const int r = task_i / p.NC, c = task_i % p.NC, n0 = c * BN;
  const int valid = min(BM, p.B - r * BM);
  const int row = e.q * 32 + e.lane;
  const bool ok = row < valid;
  const int trip = r * BM + row;
  int buf[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int idx = e.ab + k;
    buf[k] = idx & (kAcc - 1);
    mbar_wait(&e.acc_full[buf[k]], e.aph ^ (uint32_t)(idx >> 2));
  }
  tc_fence_after();
  const float *bq = p.b2[0], *bd = p.b2[1];
  // ---- pass 1: y, partial dot products over this thread's 64 columns ---------------------------------------
  float qq = 0.f, pp = 0.f, nn = 0.f, qp = 0.f, qn = 0.f;
  for (int ch = 0; ch < 4; ++ch) {
    const int n = n0 + e.hf * 64 + ch * 16;
    if (n >= p.P) break;
    float vq[16], vp[16], vn[16];
    tmem_ld16(acc_addr(e, buf[0], ch), vq);
    tmem_ld16(acc_addr(e, buf[1], ch), vp);
    tmem_ld16(acc_addr(e, buf[2], ch), vn);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = vq[j] + __ldg(bq + n + j), b = vp[j] + __ldg(bd + n + j), d = vn[j] + __ldg(bd + n + j);
      vq[j] = a; vp[j] = b; vn[j] = d;
      qq = fmaf(a, a, qq);
      pp = fmaf(b, b, pp);
      nn = fmaf(d, d, nn);
      qp = fmaf(a, b, qp);
      qn = fmaf(a, d, qn);
    }
    if (p.y && ok) {
      store16(p.y + (size_t)trip * p.P + n, vq);
      store16(p.y + (size_t)(p.B + trip) * p.P + n, vp);
      store16(p.y + (size_t)(2 * p.B + trip) * p.P + n, vn);
    }
  }
  const int nslots = 2 * p.NC;
  {
    float* sp = p.stat_part + ((size_t)(r * nslots + c * 2 + e.hf) * 5) * 128 + row;
    sp[0] = qq; sp[128] = pp; sp[256] = nn; sp[384] = qp; sp[512] = qn;
  }
  __threadfence();
  epi_bar();
  if (e.tid == 0) {
    signal(stat_ready(p) + r);
    wait_counter(stat_ready(p) + r, (unsigned)p.NC);
  }
  epi_bar();
  // ---- the whole row's sums, slot order fixed (identical in every sibling tile) ------------------------------
  qq = pp = nn = qp = qn = 0.f;
  for (int s = 0; s < nslots; ++s) {
    const float* sp = p.stat_part + ((size_t)(r * nslots + s) * 5) * 128 + row;
    qq += __ldcg(sp); pp += __ldcg(sp + 128); nn += __ldcg(sp + 256); qp += __ldcg(sp + 384); qn += __ldcg(sp + 512);
  }
  const float nq = sqrtf(qq), np_ = sqrtf(pp), nn_ = sqrtf(nn);
  const float cq = fmaxf(nq, kCosEps), cp = fmaxf(np_, kCosEps), cn = fmaxf(nn_, kCosEps);
  const float cos_p = qp / (cq * cp), cos_n = qn / (cq * cn);
  // relu(pos_dist - neg_dist + margin), model.py:140-143
  const float hinge = ok ? fmaxf((1.f - cos_p) - (1.f - cos_n) + p.margin, 0.f) : 0.f;
  const float gh = (hinge > 0.f) ? p.grad_scale * p.inv_batch : 0.f;
  const float a_qp = -gh / (cq * cp), a_qn = gh / (cq * cn);
  const float kq = (nq > kCosEps) ? (-gh * cos_p + gh * cos_n) / (cq * nq) : 0.f;
  const float kp = (np_ > kCosEps) ? (-gh * cos_p) / (cp * np_) : 0.f;
  const float kn = (nn_ > kCosEps) ? (gh * cos_n) / (cn * nn_) : 0.f;
  if (c == 0 && e.hf == 0) {
    if (ok) {
      float* s = p.stats + (size_t)trip * 8;
      s[0] = cos_p; s[1] = cos_n; s[2] = hinge; s[3] = nq; s[4] = np_; s[5] = nn_; s[6] = qp; s[7] = qn;
    }
    const float hsum = warp_sum(hinge);
    if (e.lane == 0) e.hs_s[e.q] = hsum;
  }
  // ---- pass 2: dY from the accumulators still in TMEM ---------------------------------------------------------
  for (int ch = 0; ch < 4; ++ch) {
    const int n = n0 + e.hf * 64 + ch * 16;
    if (n >= p.P) break;
    float vq[16], vp[16], vn[16];
    tmem_ld16(acc_addr(e, buf[0], ch), vq);
    tmem_ld16(acc_addr(e, buf[1], ch), vp);
    tmem_ld16(acc_addr(e, buf[2], ch), vn);
    float gq[16], gp[16], gn[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = vq[j] + __ldg(bq + n + j), b = vp[j] + __ldg(bd + n + j), d = vn[j] + __ldg(bd + n + j);
      gq[j] = ok ? (a_qp * b + a_qn * d - kq * a) : 0.f;
      gp[j] = ok ? (a_qp * a - kp * b) : 0.f;
      gn[j] = ok ? (a_qn * a - kn * d) : 0.f;
    }
    if (ok) {
      const size_t oq = (size_t)trip * p.P + n, op = (size_t)(p.B + trip) * p.P + n,
                   on = (size_t)(2 * p.B + trip) * p.P + n;
      if (p.dy) {
        store16(p.dy + oq, gq);
        store16(p.dy + op, gp);
        store16(p.dy + on, gn);
      }
      store_terms(gq, p.dy_hi, p.dy_lo, oq, p.dyt_hi, p.dyt_lo, n, p.ldt, trip);
      store_terms(gp, p.dy_hi, p.dy_lo, op, p.dyt_hi, p.dyt_lo, n, p.ldt, p.dcol + trip);
      store_terms(gn, p.dy_hi, p.dy_lo, on, p.dyt_hi, p.dyt_lo, n, p.ldt, p.dcol + p.B + trip);
    }
    stage_cs(e, 0, ch, colsum16(gq, e.lane));
#pragma unroll
    for (int j = 0; j < 16; ++j) gp[j] += gn[j];
    stage_cs(e, 1, ch, colsum16(gp, e.lane));
  }
  tc_fence_before();
#pragma unroll
  for (int k = 0; k < 3; ++k) mbar_arrive(&e.acc_empty[buf[k]]);
  acc_advance(e, 3);
  epi_bar();
  if (e.tid < BN && n0 + e.tid < p.P) {  // db2 partials of this triplet tile: query tower | document tower (p + n)
    p.cs2[(size_t)r * p.P + n0 + e.tid] = reduce_cs(e, 0, e.tid);
    p.cs2[(size_t)(p.RTB + r) * p.P + n0 + e.tid] = reduce_cs(e, 1, e.tid);
  }
  if (c == 0 && e.tid == 0) p.hinge_part[r] = (e.hs_s[0] + e.hs_s[1]) + (e.hs_s[2] + e.hs_s[3]);
  publish(e);
  if (e.tid == 0) {
    for (int seg = 0; seg < 3; ++seg) signal(dy_ready(p) + seg * p.RTB + r);
    signal(p.ctr + 1);
  }
}

__device__ void epi_dz(const ChainParams& p, EpiCtx& e, int task_i) {
  const RowTile rt = row_tile(p, task_i / p.NC);
  const int n0 = (task_i % p.NC) * BN;
  mbar_wait(&e.acc_full[e.ab], e.aph);
  tc_fence_after();
  const int row = e.q * 32 + e.lane;
  const bool ok = row < rt.valid;
  const size_t grow = (size_t)rt.grow + row;
  const int tcol = (rt.t ? p.dcol : 0) + rt.trow + row;
  for (int ch = 0; ch < 4; ++ch) {
    const int n = n0 + e.hf * 64 + ch * 16;
    if (n >= p.P) break;
    float v[16];
    tmem_ld16(acc_addr(e, e.ab, ch), v);
    const size_t o = grow * p.P + n;
    if (ok) {
      const float4* gp = reinterpret_cast<const float4*>(p.h + o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = __ldcg(gp + j);
        v[4 * j + 0] = t.x > 0.f ? v[4 * j + 0] : 0.f;
        v[4 * j + 1] = t.y > 0.f ? v[4 * j + 1] : 0.f;
        v[4 * j + 2] = t.z > 0.f ? v[4 * j + 2] : 0.f;
        v[4 * j + 3] = t.w > 0.f ? v[4 * j + 3] : 0.f;
      }
      if (p.dz1) store16(p.dz1 + o, v);
      store_terms(v, p.dz_hi, p.dz_lo, o, p.dzt_hi, p.dzt_lo, n, p.ldt, tcol);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
    stage_cs(e, 0, ch, colsum16(v, e.lane));
  }
  tc_fence_before();
  mbar_arrive(&e.acc_empty[e.ab]);
  acc_advance(e, 1);
  epi_bar();
  if (e.tid < BN && n0 + e.tid < p.P)
    p.cs1[(size_t)(rt.seg * p.RTB + rt.r) * p.P + n0 + e.tid] = reduce_cs(e, 0, e.tid);
  publish(e);
  if (e.tid == 0) {
    signal(dz_ready(p) + rt.seg * p.RTB + rt.r);
    signal(p.ctr + 1);
  }
}

__device__ void epi_dx(const ChainParams& p, EpiCtx& e, int task_i) {
  const RowTile rt = row_tile(p, task_i / p.NCH);
  const int n0 = (task_i % p.NCH) * BN;
  mbar_wait(&e.acc_full[e.ab], e.aph);
  tc_fence_after();
  const int row = e.q * 32 + e.lane;
  const bool ok = row < rt.valid;
  for (int ch = 0; ch < 4; ++ch) {
    const int n = n0 + e.hf * 64 + ch * 16;
    if (n >= p.H) break;
    float v[16];
    tmem_ld16(acc_addr(e, e.ab, ch), v);
    if (ok) store16(p.dxhat + ((size_t)rt.grow + row) * p.H + n, v);
  }
  tc_fence_before();
  mbar_arrive(&e.acc_empty[e.ab]);
  acc_advance(e, 1);
}

// raw partial sums of one (chunk, tile) of a weight gradient: part[chunk][M][N]
__device__ void epi_dw(const ChainParams& p, EpiCtx& e, Task tk) {
  const bool two = tk.type == T_DW2;
  const int N = two ? p.P : p.H, ntn = two ? p.NC : p.NCH, nt_ = p.NC * ntn;
  const int ci = tk.i / nt_, tile = tk.i % nt_;
  const int m = (tile / ntn) * BM + e.q * 32 + e.lane, n0 = (tile % ntn) * BN;
  float* part = (two ? p.part2 : p.part1) + (size_t)ci * p.P * N;
  mbar_wait(&e.acc_full[e.ab], e.aph);
  tc_fence_after();
  for (int ch = 0; ch < 4; ++ch) {
    const int n = n0 + e.hf * 64 + ch * 16;
    if (n >= N) break;
    float v[16];
    tmem_ld16(acc_addr(e, e.ab, ch), v);
    if (m < p.P) store16(part + (size_t)m * N + n, v);
  }
  tc_fence_before();
  mbar_arrive(&e.acc_empty[e.ab]);
  acc_advance(e, 1);
  __threadfence();
  epi_bar();
  if (e.tid == 0) signal(p.ctr + 1);
}

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

struct AdamCoef {
  float step_size, bc2_sqrt;
};
// torch.optim.Adam (non-fused, no weight decay / amsgrad): lerp, mul + addcmul, sqrt / bc2_sqrt + eps, addcdiv
__device__ __forceinline__ float adam1(const ChainParams& p, const AdamCoef& k, float g, float& m, float& v, float w) {
  m = m + (g - m) * (1.f - p.beta1);
  v = v * p.beta2 + (1.f - p.beta2) * g * g;
  const float denom = sqrtf(v) / k.bc2_sqrt + p.eps;
  return w - k.step_size * (m / denom);
}
// gradient element(s) at dst (a slice of the flat gradient buffer) -> memory, and the Adam update of the parameter,
// exp_avg and exp_avg_sq elements at the same offset of their flat buffers
__device__ __forceinline__ void emit4(const ChainParams& p, const AdamCoef& k, float4* dst, float4 g) {
  *dst = g;
  if (!p.adam_p) return;
  const size_t off = reinterpret_cast<float*>(dst) - p.adam_g;
  float4 m = *reinterpret_cast<float4*>(p.adam_m + off), v = *reinterpret_cast<float4*>(p.adam_v + off);
  float4 w = *reinterpret_cast<float4*>(p.adam_p + off);
  w.x = adam1(p, k, g.x, m.x, v.x, w.x);
  w.y = adam1(p, k, g.y, m.y, v.y, w.y);
  w.z = adam1(p, k, g.z, m.z, v.z, w.z);
  w.w = adam1(p, k, g.w, m.w, v.w, w.w);
  *reinterpret_cast<float4*>(p.adam_m + off) = m;
  *reinterpret_cast<float4*>(p.adam_v + off) = v;
  *reinterpret_cast<float4*>(p.adam_p + off) = w;
}
__device__ __forceinline__ void emit1(const ChainParams& p, const AdamCoef& k, float* dst, float g) {
  *dst = g;
  if (!p.adam_p) return;
  const size_t off = dst - p.adam_g;
  float m = p.adam_m[off], v = p.adam_v[off];
  p.adam_p[off] = adam1(p, k, g, m, v, p.adam_p[off]);
  p.adam_m[off] = m;
  p.adam_v[off] = v;
}

// gradients = fixed-order sums of the partials (chunk order / row-tile order) [-> Adam]; task 0 also finishes the loss.
// Nothing else reads the fp32 parameters any more at this point (the weights were split into bf16 terms by S, the
// biases were last read by the F1 / F2L epilogues, all of which have signalled), so they are updated in place.
__device__ void epi_grad(const ChainParams& p, EpiCtx& e, int task_i) {
  if (e.tid == 0) wait_counter(p.ctr + 1, p.total_signals);
  epi_bar();
  AdamCoef k{0.f, 1.f};
  if (p.adam_p) {
    const volatile double* st = p.adam_state;
    k.step_size = (float)((double)p.lr / (1.0 - st[1]));
    k.bc2_sqrt = (float)sqrt(1.0 - st[2]);
  }
  const int nG = p.off[T_COUNT] - p.off[T_G];
  const size_t gid = (size_t)task_i * kEpiThreads + e.tid, gstride = (size_t)nG * kEpiThreads;
  for (int t = 0; t < 2; ++t) {
    const int c0 = t ? p.nch[0] : 0;
    {
      const size_t n4 = (size_t)p.P * p.P / 4;
      const float4* src = reinterpret_cast<const float4*>(p.part2) + (size_t)c0 * n4;
      float4* dst = reinterpret_cast<float4*>(p.dW2[t]);
      for (size_t i = gid; i < n4; i += gstride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < p.nch[t]; ++j) acc = add4(acc, __ldcg(src + (size_t)j * n4 + i));
        emit4(p, k, dst + i, acc);
      }
    }
    {
      const size_t n4 = (size_t)p.P * p.H / 4;
      const float4* src = reinterpret_cast<const float4*>(p.part1) + (size_t)c0 * n4;
      float4* dst = reinterpret_cast<float4*>(p.dW1[t]);
      for (size_t i = gid; i < n4; i += gstride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < p.nch[t]; ++j) acc = add4(acc, __ldcg(src + (size_t)j * n4 + i));
        emit4(p, k, dst + i, acc);
      }
    }
    for (size_t n = gid; n < (size_t)p.P; n += gstride) {
      float s2 = 0.f, s1 = 0.f;
      for (int r = 0; r < p.RTB; ++r) s2 += __ldcg(p.cs2 + (size_t)(t * p.RTB + r) * p.P + n);
      if (t == 0) {
        for (int r = 0; r < p.RTB; ++r) s1 += __ldcg(p.cs1 + (size_t)r * p.P + n);
      } else {
        for (int r = 0; r < 2 * p.RTB; ++r) s1 += __ldcg(p.cs1 + (size_t)(p.RTB + r) * p.P + n);
      }
      emit1(p, k, p.db2[t] + n, s2);
      emit1(p, k, p.db1[t] + n, s1);
    }
  }
  if (task_i == 0 && e.tid == 0) {
    float s = 0.f;
    for (int r = 0; r < p.RTB; ++r) s += __ldcg(p.hinge_part + r);
    *p.loss = s * p.inv_batch;
  }
}

__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int NS = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NS * kStageBytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* acc_full = empty + kMaxStages;
  uint64_t* acc_empty = acc_full + kAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAcc);
  uint8_t* staging = smem + NS * kStageBytes + 256;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < kAcc; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kAcc * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tasks = p.off[T_COUNT];

  if (warp == 0) {
    if (elect_one()) {  // ---- TMA producer: dependencies, then the operand tiles of every sub-tile ------------
      int s = 0;
      uint32_t ph = 0;
      for (int idx = blockIdx.x; idx < n_tasks; idx += gridDim.x) {
        const Task tk = decode(p, idx);
        const int ns = n_subs(tk.type);
        if (ns == 0) continue;
        wait_deps(p, tk);
        fence_proxy_async();  // other CTAs' generic-proxy writes (acquired above) before this thread's TMA reads
        for (int k = 0; k < ns; ++k) {
          const Sub sb = get_sub(p, tk, k);
          for (int kbi = 0; kbi < sb.nkb; ++kbi) {
            for (int j = 0; j < sb.terms; ++j) {
              mbar_wait(&empty[s], ph ^ 1u);
              mbar_arrive_expect_tx(&full[s], kStageBytes);
              tma_load_2d(smem + s * kStageBytes, &sb.a[j], &full[s], (sb.kb0 + kbi) * BK, sb.m0);
              tma_load_2d(smem + s * kStageBytes + kABytes, &sb.b[j], &full[s], (sb.kb0 + kbi) * BK, sb.n0);
              if (++s == NS) {
                s = 0;
                ph ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // ---- MMA issuer ----------------------------------------------------------------------
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      const uint32_t ring = smem_u32(smem);
      int s = 0, ab = 0;
      uint32_t ph = 0, aph = 0;
      for (int idx = blockIdx.x; idx < n_tasks; idx += gridDim.x) {
        const Task tk = decode(p, idx);
        const int ns = n_subs(tk.type);
        for (int k = 0; k < ns; ++k) {
          const Sub sb = get_sub(p, tk, k);
          mbar_wait(&acc_empty[ab], aph ^ 1u);
          tc_fence_after();
          const uint32_t acc = tmem_base + (uint32_t)(ab * BN);
          for (int kbi = 0; kbi < sb.nkb; ++kbi) {
            uint32_t slot_addr[3];
            int slot_id[3];
            for (int j = 0; j < sb.terms; ++j) {
              mbar_wait(&full[s], ph);
              slot_id[j] = s;
              slot_addr[j] = ring + (uint32_t)s * kStageBytes;
              if (++s == NS) {
                s = 0;
                ph ^= 1u;
              }
            }
            tc_fence_after();
            for (int pair = 0; pair < sb.pairs; ++pair) {
              const int ai = (0x201100 >> (4 * pair)) & 3, bi = (0x021010 >> (4 * pair)) & 3;
              const uint64_t da = make_smem_desc_sw128(slot_addr[ai]),
                             db = make_smem_desc_sw128(slot_addr[bi] + kABytes);
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk)
                mma_bf16(acc, da + 2 * kk, db + 2 * kk, idesc, (kbi | pair | kk) != 0);
            }
            for (int j = 0; j < sb.terms; ++j) mma_commit(&empty[slot_id[j]]);
          }
          mma_commit(&acc_full[ab]);
          if (++ab == kAcc) {
            ab = 0;
            aph ^= 1u;
          }
        }
      }
    }
  } else {  // ---- epilogue warps ------------------------------------------------------------------------------
    EpiCtx e;
    e.tmem_base = tmem_base;
    e.acc_full = acc_full;
    e.acc_empty = acc_empty;
    e.cs_s = reinterpret_cast<float*>(staging);
    e.sp_s = reinterpret_cast<bf16*>(staging + 2 * 8 * 64 * 4);
    e.hs_s = reinterpret_cast<float*>(staging + 2 * 8 * 64 * 4 + 2 * 32 * 33 * 2);
    e.we = warp - 2;
    e.q = warp & 3;
    e.hf = e.we >> 2;
    e.lane = lane;
    e.tid = threadIdx.x - 64;
    e.ab = 0;
    e.aph = 0;
    for (int idx = blockIdx.x; idx < n_tasks; idx += gridDim.x) {
      const Task tk = decode(p, idx);
      switch (tk.type) {
        case T_S: epi_split(p, e, tk.i); break;
        case T_F1: epi_f1(p, e, tk.i); break;
        case T_F2L: epi_f2l(p, e, tk.i); break;
        case T_DZ: epi_dz(p, e, tk.i); break;
        case T_DX: epi_dx(p, e, tk.i); break;
        case T_DW2:
        case T_DW1: epi_dw(p, e, tk); break;
        default: epi_grad(p, e, tk.i); break;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kAcc * BN);
}

int chain_stages() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TT_CHAIN_STAGES");
    v = e ? atoi(e) : 0;
    if (v < 3 || v > kMaxStages) v = 5;
  }
  return v;
}

int chain_debug_out() {  // TT_CHAIN_FP32_OUT=1: also write y, dY and dz1 in fp32 (diagnostics)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TT_CHAIN_FP32_OUT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

int map_terms(CUtensorMap* m, int n, const bf16* const* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
  for (int i = 0; i < n; ++i) {
    const bf16* base = ptr[i] ? ptr[i] : ptr[0];  // unused terms alias the first (never loaded)
    int rc = make_map_bf16_kmajor(&m[i], base, rows, cols, ld, BM);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace

bool chain_enabled() {
  const char* e = getenv("TT_CHAIN");  // 0 = the per-kernel chain of tt_gemm_sm100.cu; read per call (tests toggle it)
  return e ? atoi(e) != 0 : true;
}

// Everything of the step after the pooled gather.  s.adam (optional): torch.optim.Adam on the flat parameter buffer in
// the kernel's tail, step count and beta powers advanced on the device.
int chain_sm100(const StepSm100& s, cudaStream_t st) {
  const int B = s.B, H = s.H, P = s.P;
  StepWs w;
  carve_step(reinterpret_cast<char*>(s.ws), B, H, P, s.dxhat != nullptr, &w);
  static bool attr_done = false;
  if (!attr_done) {
    TT_CUDA(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem(kMaxStages)));
    attr_done = true;
  }
  static thread_local ChainParams p;  // ~8 KB: kept off the stack; copied into the launch before this returns
  p = ChainParams{};
  int rc;
  const float* W1[2] = {s.Wq1, s.Wd1};
  const float* W2[2] = {s.Wq2, s.Wd2};
  const int row0[2] = {0, B}, rows[2] = {B, 2 * B}, tcol[2] = {0, w.dcol};
  const bool x3 = s.n_split == 3;
  for (int t = 0; t < 2; ++t) {
    TowerMaps& m = p.tm[t];
    const size_t oH = (size_t)row0[t] * H, oP = (size_t)row0[t] * P;
    const bf16* x[3] = {w.x_hi + oH, x3 ? w.x_lo + oH : nullptr, x3 ? w.x_lo2 + oH : nullptr};
    const bf16* w1[3] = {w.w1_hi[t], w.w1_lo[t], w.w1_lo2[t]};
    const bf16* hh[2] = {w.h_hi + oP, w.h_lo + oP};
    const bf16* w2[2] = {w.w2_hi[t], w.w2_lo[t]};
    const bf16* dy[2] = {w.dy_hi + oP, w.dy_lo + oP};
    const bf16* w2t[2] = {w.w2t_hi[t], w.w2t_lo[t]};
    const bf16* dyt[2] = {w.dyt_hi + tcol[t], w.dyt_lo + tcol[t]};
    const bf16* ht[2] = {w.ht_hi + tcol[t], w.ht_lo + tcol[t]};
    const bf16* dzt[2] = {w.dzt_hi + tcol[t], w.dzt_lo + tcol[t]};
    const bf16* xt[2] = {w.xt_hi + tcol[t], w.xt_lo + tcol[t]};
    const bf16* dz[2] = {w.dz_hi + oP, w.dz_lo + oP};
    const bf16* w1t[2] = {w.w1t_hi[t], w.w1t_lo[t]};
    if ((rc = map_terms(m.x, 3, x, rows[t], H, H))) return rc;
    if ((rc = map_terms(m.w1, 3, w1, P, H, H))) return rc;
    if ((rc = map_terms(m.h, 2, hh, rows[t], P, P))) return rc;
    if ((rc = map_terms(m.w2, 2, w2, P, P, P))) return rc;
    if ((rc = map_terms(m.dy, 2, dy, rows[t], P, P))) return rc;
    if ((rc = map_terms(m.w2t, 2, w2t, P, P, P))) return rc;
    if ((rc = map_terms(m.dyt, 2, dyt, P, rows[t], w.ldt))) return rc;
    if ((rc = map_terms(m.ht, 2, ht, P, rows[t], w.ldt))) return rc;
    if ((rc = map_terms(m.dzt, 2, dzt, P, rows[t], w.ldt))) return rc;
    if ((rc = map_terms(m.xt, 2, xt, H, rows[t], w.ldt))) return rc;
    if ((rc = map_terms(m.dz, 2, dz, rows[t], P, P))) return rc;
    if ((rc = map_terms(m.w1t, 2, w1t, H, P, P))) return rc;
  }
  p.B = B; p.H = H; p.P = P;
  p.RTB = (B + BM - 1) / BM; p.NC = (P + BN - 1) / BN; p.NCH = (H + BN - 1) / BN;
  p.ldt = w.ldt; p.dcol = w.dcol;
  p.pairs = x3 ? 3 : 1; p.terms = x3 ? 2 : 1;
  {
    const char* e = getenv("TT_FWD1_PRODUCTS");  // tuning hook shared with the per-kernel chain
    const int six = !(e && atoi(e) == 3);
    p.pairs1 = x3 ? (six ? 6 : 3) : 1;
    p.terms1 = x3 ? (six ? 3 : 2) : 1;
  }
  for (int t = 0; t < 2; ++t) p.kbt[t] = (rows[t] + BK - 1) / BK;
  p.kcb = 8;
  while ((p.kbt[0] + p.kcb - 1) / p.kcb + (p.kbt[1] + p.kcb - 1) / p.kcb > 48) p.kcb += 8;
  for (int t = 0; t < 2; ++t) p.nch[t] = (p.kbt[t] + p.kcb - 1) / p.kcb;
  const int grid = sm_count();
  const int n_rt = 3 * p.RTB, nchs = p.nch[0] + p.nch[1];
  const int count[T_COUNT] = {grid,
                              n_rt * p.NC,
                              p.RTB * p.NC,
                              n_rt * p.NC,
                              s.dxhat ? n_rt * p.NCH : 0,
                              nchs * p.NC * p.NC,
                              nchs * p.NC * p.NCH,
                              grid};
  p.off[0] = 0;
  for (int k = 0; k < T_COUNT; ++k) p.off[k + 1] = p.off[k] + count[k];
  p.total_signals = (unsigned)(count[T_F2L] + count[T_DZ] + count[T_DW2] + count[T_DW1]);
  p.stages = chain_stages();
  // weights of both towers -> bf16 terms (+ transposes for the backward contractions)
  int tile0 = 0;
  p.n_sj = 0;
  for (int t = 0; t < 2; ++t) {
    SplitJobC a{W1[t], w.w1_hi[t], w.w1_lo[t], x3 ? w.w1_lo2[t] : nullptr, s.dxhat ? w.w1t_hi[t] : nullptr,
                s.dxhat ? w.w1t_lo[t] : nullptr, P, H, P, tile0, (H + 31) / 32};
    tile0 += ((P + 31) / 32) * a.tiles_c;
    p.sj[p.n_sj++] = a;
    SplitJobC b{W2[t], w.w2_hi[t], w.w2_lo[t], nullptr, w.w2t_hi[t], w.w2t_lo[t], P, P, P, tile0, (P + 31) / 32};
    tile0 += ((P + 31) / 32) * b.tiles_c;
    p.sj[p.n_sj++] = b;
  }
  p.n_split_tiles = tile0;
  p.margin = s.margin; p.inv_batch = s.inv_batch; p.grad_scale = s.grad_scale;
  p.b1[0] = s.bq1; p.b1[1] = s.bd1; p.b2[0] = s.bq2; p.b2[1] = s.bd2;
  const bool dbg = chain_debug_out() != 0;
  p.h = s.h; p.y = dbg ? s.y : nullptr; p.dy = dbg ? s.dy : nullptr; p.dz1 = dbg ? w.dz1 : nullptr;
  p.dxhat = s.dxhat; p.stats = s.stats; p.loss = s.loss;
  p.h_hi = w.h_hi; p.h_lo = w.h_lo; p.ht_hi = w.ht_hi; p.ht_lo = w.ht_lo;
  p.dy_hi = w.dy_hi; p.dy_lo = w.dy_lo; p.dyt_hi = w.dyt_hi; p.dyt_lo = w.dyt_lo;
  p.dz_hi = s.dxhat ? w.dz_hi : nullptr; p.dz_lo = s.dxhat ? w.dz_lo : nullptr;
  p.dzt_hi = w.dzt_hi; p.dzt_lo = w.dzt_lo;
  p.part2 = w.partial2; p.part1 = w.partial;
  p.dW1[0] = s.dWq1; p.dW1[1] = s.dWd1; p.db1[0] = s.dbq1; p.db1[1] = s.dbd1;
  p.dW2[0] = s.dWq2; p.dW2[1] = s.dWd2; p.db2[0] = s.dbq2; p.db2[1] = s.dbd2;
  p.stat_part = w.stat_part; p.cs1 = w.cs1; p.cs2 = w.cs2; p.hinge_part = w.hinge_part;
  p.ctr = w.counters;
  if (s.adam.param) {
    const FusedAdam& a = s.adam;
    TT_REQUIRE(a.state && a.grad && a.exp_avg && a.exp_avg_sq, "tt_triplet_step: fused Adam needs state, grad and both moments");
    const float* W[8] = {s.Wq1, s.bq1, s.Wq2, s.bq2, s.Wd1, s.bd1, s.Wd2, s.bd2};
    float* G[8] = {s.dWq1, s.dbq1, s.dWq2, s.dbq2, s.dWd1, s.dbd1, s.dWd2, s.dbd2};
    const size_t len[8] = {(size_t)P * H, (size_t)P, (size_t)P * P, (size_t)P, (size_t)P * H, (size_t)P, (size_t)P * P, (size_t)P};
    for (int i = 0; i < 8; ++i) {
      const ptrdiff_t off = G[i] - a.grad;
      TT_REQUIRE(off >= 0 && (size_t)off + len[i] <= a.n && W[i] == a.param + off && off % 4 == 0,
                 "tt_triplet_step: fused Adam wants the 8 projection tensors and their gradients to be slices of the "
                 "flat buffers at equal, 16-byte aligned offsets (tensor %d)", i);
    }
    p.adam_state = a.state; p.adam_p = a.param; p.adam_g = a.grad; p.adam_m = a.exp_avg; p.adam_v = a.exp_avg_sq;
    p.lr = a.lr; p.beta1 = a.beta1; p.beta2 = a.beta2; p.eps = a.eps;
  }
  TT_CUDA(cudaMemsetAsync(w.counters, 0, (size_t)w.n_counters * sizeof(unsigned), st));
  chain_kernel<<<grid, kThreads, chain_smem(p.stages), st>>>(p);
  TT_LAUNCH_CHECK();
  return 0;
}

}  // namespace tt
