// pf_conv_tc.cu — convolution forward / dgrad on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// The one genuine dense contraction of the step (tf.nn.conv2d re-created on the
// quantized weight, /root/reference/learners/uniform_quantization/utils.py:92-104, and its dgrad).
// tcgen05 has no fp32 x fp32 MMA, and the parity bar is fp32 (1e-5 on losses), so operands are split
//     x = hi + lo,  hi = bf16(x), lo = bf16(x - hi)         (representation error 2^-18)
// and every k-slice issues three bf16 MMAs into ONE fp32 TMEM accumulator:
//     D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi               (the lo*lo term, 2^-18 relative, is dropped)
//
// Implicit GEMM, both operands K-major:  D[M x N] = A[M x K] * B[N x K]^T
//   fwd  : M = N*P*Q pixels, N = Cout, K = R*S*Cin;  A = im2col(x) gathered on the fly; B = w^T
//   dgrad: M = N*H*W pixels, N = Cin,  K = R*S*Cout; A = gathered dy;                  B = w as [Cin][(r,s,cout)]
// B is pre-split / pre-transposed once per step by pf_conv2d_tc_prep_weight (weights change every
// step; 20 B per weight).  A rows are gathered from NHWC fp32 by 128 producer threads (one GEMM row
// each: 64 contiguous channels of one tap = 256 B), split, and written to 128B-swizzled K-major smem
// tiles; one elected thread issues the MMAs; accumulators live in TMEM; the producer warps then run
// the epilogue (tcgen05.ld -> global NHWC fp32).  mbarrier ring of kStages smem stages.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "pf_common.cuh"
#include "pf_tc_common.cuh"

namespace {
using namespace pftc;

constexpr int TM = 128;      // GEMM rows per CTA (= TMEM lanes)
constexpr int BK = 64;       // bf16 elements per k-stage (= one 128-byte swizzled row)
constexpr int kProducerThreads = 256;  // two threads per GEMM row: 8 x 16 B of A and 8 x 16 B of B each per stage
constexpr int kThreads = 288;          // warps 0-7: producers + epilogue, warp 8: MMA issuer + TMEM alloc
constexpr int kMaxStages = 4;

struct TcGeom {
  int N, H, W, C, K, R, S, P, Q, sh, sw, pt, pl;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  const __nv_bfloat16 hx = __float2bfloat16_rn(v.x), hy = __float2bfloat16_rn(v.y);
  const __nv_bfloat16 hz = __float2bfloat16_rn(v.z), hw = __float2bfloat16_rn(v.w);
  hi.x = (uint32_t)__bfloat16_as_ushort(hx) | ((uint32_t)__bfloat16_as_ushort(hy) << 16);
  hi.y = (uint32_t)__bfloat16_as_ushort(hz) | ((uint32_t)__bfloat16_as_ushort(hw) << 16);
  lo.x = pack_bf16(v.x - __bfloat162float(hx), v.y - __bfloat162float(hy));
  lo.y = pack_bf16(v.z - __bfloat162float(hz), v.w - __bfloat162float(hw));
}

// MODE 0 = fwd (gather x), 1 = dgrad (gather dy).
// Producers (8 warps): 8 lanes cover one row's 64-channel chunk (256 contiguous bytes of NHWC fp32):
// thread t owns bytes [32*(t&7), +32) of rows (t>>3) + 32*i, i < 4 — a warp load touches 4 rows x 256 B
// of fully used lines (the first version's one-row-per-thread mapping saturated the L1 tag stage:
// ncu l1tex 79 %, 32 tag lookups per load instruction).  8 channels -> one 16-byte bf16 chunk each for
// the hi and the lo tile.  A is double-buffered in registers one k-stage ahead (ping-pong sets, no
// copies); B (pre-split weights, no conversion) goes global -> shared with cp.async straight into the
// swizzled tile, overlapping the A conversion.
// NPW = producer warps: 8 (1 CTA/SM, A ping-pong in registers; large K) or 4 (2-3 CTAs/SM so that one CTA's
// loads / epilogue overlap another's MMAs; the small-K, output-bound 1x1 layers of the early stages).
template <int MODE, int NPW>
__global__ void __launch_bounds__(NPW * 32 + 32, NPW == 8 ? 1 : 2)
conv_tc_kernel(const float* __restrict__ src, const __nv_bfloat16* __restrict__ b_hi,
               const __nv_bfloat16* __restrict__ b_lo, float* __restrict__ out, TcGeom g, int M, int Ng,
               int Kdim, int Kpad, int BN, int n_stages, int accumulate, const float* __restrict__ bias,
               int relu, const float* __restrict__ residual) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = TM * 128, b_bytes = (uint32_t)BN * 128;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], accum_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * BN;
  const int nk = Kpad / BK;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < BN) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full_bar[s], NPW * 32);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  constexpr int kProd = NPW * 32;          // producer threads
  constexpr int ROWS = TM * 8 / kProd;     // A rows per thread: 4 (NPW = 8) or 8 (NPW = 4)
  constexpr int RSTEP = kProd / 8;         // row stride between a thread's rows: 32 or 16 (multiples of 8)
  if (warp == NPW) tmem_alloc(&tmem_base_s, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < NPW) {
    const int l8 = tid & 7, rgrp = tid >> 3;          // 32-byte slice of the row, row group
    const int CC = (MODE == 0) ? g.C : g.K;           // channels of the gathered tensor
    int pn[ROWS], yx[ROWS];                           // per row: image index (< 0: beyond M), packed (y0, x0)
    {
      const int hw = (MODE == 0) ? g.P * g.Q : g.H * g.W;
      const int wq = (MODE == 0) ? g.Q : g.W;
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        const int m = m0 + rgrp + RSTEP * i;
        pn[i] = -1;
        yx[i] = 0;
        if (m < M) {
          const int n_ = m / hw;
          const int rem = m - n_ * hw;
          const int y = rem / wq, x = rem - y * wq;
          pn[i] = n_;
          const int y0 = (MODE == 0) ? y * g.sh - g.pt : y + g.pt;
          const int x0 = (MODE == 0) ? x * g.sw - g.pl : x + g.pl;
          yx[i] = (int)(((uint32_t)(y0 + 32768) << 16) | (uint32_t)(x0 + 32768));
        }
      }
    }
    const bool unit_stride = g.sh == 1 && g.sw == 1;
    const uint32_t sw = (uint32_t)(rgrp & 7);         // (row & 7) for every row of this thread (rows differ by 32)
    const uint32_t chunk_off = (((uint32_t)l8) ^ sw) << 4;
    auto issue_loads_a = [&](int ks, float4 (&av)[2 * ROWS]) {
      const int kk = ks * BK + (l8 >> 1) * 16;        // this lane's 16-channel sub-chunk: inside one filter tap
      const bool kok = kk < Kdim;
      const int tap = kk / CC, c = kk - tap * CC + (l8 & 1) * 8;
      const int r = tap / g.S, q = tap - r * g.S;
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        const float* p = nullptr;
        if (kok && pn[i] >= 0) {
          const int y0 = (int)((uint32_t)yx[i] >> 16) - 32768, x0 = (int)((uint32_t)yx[i] & 0xFFFFu) - 32768;
          if (MODE == 0) {
            const int ih = y0 + r, iw = x0 + q;
            if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) p = src + (((size_t)pn[i] * g.H + ih) * g.W + iw) * g.C + c;
          } else {
            const int th = y0 - r, tw = x0 - q;
            if (th >= 0 && tw >= 0) {
              if (unit_stride) {
                if (th < g.P && tw < g.Q) p = src + (((size_t)pn[i] * g.P + th) * g.Q + tw) * g.K + c;
              } else {
                const int oh = th / g.sh, ow = tw / g.sw;
                if (oh * g.sh == th && ow * g.sw == tw && oh < g.P && ow < g.Q)
                  p = src + (((size_t)pn[i] * g.P + oh) * g.Q + ow) * g.K + c;
              }
            }
          }
        }
        av[2 * i] = p ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
        av[2 * i + 1] = p ? __ldg(reinterpret_cast<const float4*>(p) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto issue_b_async = [&](int ks, uint8_t* st) {
      // rows rgrp + RSTEP*i of the BN x 128-byte tile, 16-byte chunk l8, hi and lo; zero-fill out-of-range rows
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        const int br = rgrp + RSTEP * i;
        if (br < BN) {
          const bool ok = n0 + br < Ng;
          const size_t off = (size_t)(ok ? n0 + br : 0) * Kpad + (size_t)ks * BK + l8 * 8;
          const uint32_t dst = smem_u32(st + 2 * a_bytes + (size_t)br * 128 + chunk_off);
          const uint32_t nbytes = ok ? 16u : 0u;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(b_hi + off), "r"(nbytes) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + b_bytes), "l"(b_lo + off), "r"(nbytes) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto store_a = [&](uint8_t* st, const float4 (&av)[2 * ROWS]) {
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        uint2 h0, l0, h1, l1;
        split4(av[2 * i], h0, l0);
        split4(av[2 * i + 1], h1, l1);
        uint8_t* rowp = st + (size_t)(rgrp + RSTEP * i) * 128 + chunk_off;
        *reinterpret_cast<uint4*>(rowp) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4*>(rowp + a_bytes) = make_uint4(l0.x, l0.y, l1.x, l1.y);
      }
    };
    auto do_stage = [&](int ks, const float4 (&av)[2 * ROWS]) {
      const int s = ks % n_stages;
      mbar_wait(&empty_bar[s], (((uint32_t)(ks / n_stages)) & 1u) ^ 1u);   // slot free?
      uint8_t* st = smem + (size_t)s * stage_bytes;
      issue_b_async(ks, st);
      store_a(st, av);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async_smem();   // generic-proxy + cp.async writes -> visible to the tensor core (async proxy)
      mbar_arrive(&full_bar[s]);
    };
    if (NPW == 8) {
      float4 a0[2 * ROWS], a1[2 * ROWS];
      issue_loads_a(0, a0);
      for (int ks = 0; ks < nk; ks += 2) {
        if (ks + 1 < nk) issue_loads_a(ks + 1, a1);
        do_stage(ks, a0);
        if (ks + 1 < nk) {
          if (ks + 2 < nk) issue_loads_a(ks + 2, a0);
          do_stage(ks + 1, a1);
        }
      }
    } else {
      for (int ks = 0; ks < nk; ++ks) {
        float4 a0[2 * ROWS];
        issue_loads_a(ks, a0);
        do_stage(ks, a0);
      }
    }
    // ======================= epilogue: TMEM -> registers -> smem -> coalesced global =======================
    // warp w owns TMEM lanes [32*(w%4), +32); warps 0-3 take the low half of the columns, 4-7 the high half.
    // The tile is staged row-major in the (now idle) stage memory so that global stores are full rows.
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    float* stile = reinterpret_cast<float*>(smem);
    const int pitch = BN + 4;                       // +16 B: conflict-free 128-bit row writes
    {
      const int erow = (warp & 3) * 32 + (tid & 31);
      const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
      const int cbeg = (NPW == 8 && BN >= 64) ? (warp >> 2) * (BN / 2) : 0;
      const int cend = (NPW == 8) ? ((BN >= 64) ? cbeg + BN / 2 : ((warp >> 2) == 0 ? BN : 0)) : BN;
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + lane_base + (uint32_t)c0, r);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (c0 + j < BN)
            *reinterpret_cast<uint4*>(stile + (size_t)erow * pitch + c0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NPW * 32) : "memory");   // the epilogue (= producer) warps only
    {
      // full rows to global, 4 independent 16-byte pieces per thread per round so that the residual /
      // accumulate loads of a round are all in flight before the first one is consumed
      const int vec_per_row = BN >> 2;
      const int total = TM * vec_per_row;
      const float* extra = residual ? residual : (accumulate ? out : nullptr);
      for (int e0 = tid; e0 < total; e0 += 4 * kProd) {
        float4 xv[4];
        size_t goff[4];
        int sidx[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * kProd;
          const int rr = e / vec_per_row, cv = (e - rr * vec_per_row) << 2;
          const int mm = m0 + rr, nn = n0 + cv;
          ok[u] = e < total && mm < M && nn + 3 < Ng;
          goff[u] = (size_t)mm * Ng + nn;
          sidx[u] = rr * pitch + cv;
          xv[u] = (ok[u] && extra) ? *reinterpret_cast<const float4*>(extra + goff[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!ok[u]) continue;
          float4 v = *reinterpret_cast<const float4*>(stile + sidx[u]);
          if (bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + (goff[u] % (size_t)Ng)));
            v.x = __fadd_rn(v.x, bb.x); v.y = __fadd_rn(v.y, bb.y); v.z = __fadd_rn(v.z, bb.z); v.w = __fadd_rn(v.w, bb.w);
          }
          if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          if (extra) {   // fused residual add (resnet_model.py:199,314) or dx += (gradient accumulation)
            v.x = __fadd_rn(v.x, xv[u].x); v.y = __fadd_rn(v.y, xv[u].y); v.z = __fadd_rn(v.z, xv[u].z); v.w = __fadd_rn(v.w, xv[u].w);
          }
          *reinterpret_cast<float4*>(out + goff[u]) = v;
        }
      }
    }
  } else if (warp == NPW) {
    // ======================= MMA issuer: one elected thread =======================
    if ((tid & 31) == 0) {
      const uint32_t idesc = make_idesc_bf16(TM, BN, 0, 0);
      for (int ks = 0; ks < nk; ++ks) {
        const int s = ks % n_stages;
        mbar_wait(&full_bar[s], ((uint32_t)(ks / n_stages)) & 1u);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_hi = base, a_lo = base + a_bytes, bh = base + 2 * a_bytes, bl = bh + b_bytes;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t dah = make_smem_desc(a_hi + kk * 32, 16, 1024);
          const uint64_t dal = make_smem_desc(a_lo + kk * 32, 16, 1024);
          const uint64_t dbh = make_smem_desc(bh + kk * 32, 16, 1024);
          const uint64_t dbl = make_smem_desc(bl + kk * 32, 16, 1024);
          umma_bf16(tmem_base, dah, dbh, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
          umma_bf16(tmem_base, dah, dbl, idesc, 1u);
          umma_bf16(tmem_base, dal, dbh, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);   // frees the smem stage when these MMAs have completed
      }
      umma_commit(&accum_bar);        // accumulator complete
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NPW) tmem_dealloc(tmem_base, tmem_cols);
}



// =============================================================================================
// v4: PERSISTENT, warp-specialised fwd / dgrad kernel.  One CTA per SM loops over output tiles;
//   warps 0-3   epilogue  (TMEM -> registers -> per-warp smem transpose -> coalesced global; bias / ReLU /
//               residual / accumulate), overlapped with the next tile's main loop through a DOUBLE-BUFFERED
//               TMEM accumulator (2 x BN columns);
//   warps 4-11  producers (A gather + fp32 -> split-bf16 conversion with a register ping-pong that runs
//               across tile boundaries; B via cp.async);
//   warp  12    MMA issuer (one elected thread) + TMEM allocation.
// BN goes up to 256 (A is read once for 256 output channels).  When one n-tile covers all output channels and
// the whole split weight matrix fits next to >= 2 A stages, B is loaded ONCE per CTA and stays resident
// ("B-stationary": every 1x1 layer of the early stages, K <= 256).
// MODE 2 = dgrad of a strided convolution decomposed into stride_h*stride_w pixel-parity classes: the rows of
// a tile all belong to one class (h = ph + sh*h', w = pw + sw*w'), and only the filter taps that can reach that
// class are visited, so no MMA multiplies structural zeros (the gather-with-divisibility-test formulation of
// MODE 1 wastes 3/4 of the tensor-core work of a 3x3 stride-2 layer).
struct FastDiv {
  uint32_t mul, shift;
};
inline FastDiv make_fastdiv(uint32_t d) {   // exact for 0 <= n < 2^31 (Granlund-Montgomery round-up method)
  FastDiv f;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.shift = s;
  f.mul = (uint32_t)((((1ull << 32) * ((1ull << s) - d)) / d) + 1);
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, FastDiv f) { return (__umulhi(n, f.mul) + n) >> f.shift; }

constexpr int kMaxClasses = 4;
constexpr int kMaxTaps = 9;
struct TcClass {
  int tile_begin, Mc, Hc, Wc, ph, pw, ntaps, pad_;
  FastDiv d_hw, d_w;
  int8_t eh[kMaxTaps + 3], ew[kMaxTaps + 3], tap[kMaxTaps + 3];
};
struct TcP {
  TcGeom g;
  int M, Ng, Kdim, Kpad, BN, nk, n_stages, n_bslots, b_stationary, m_tiles, n_tiles, total_tiles, acc_cols;
  int accumulate, relu, ncls, cblocks;
  FastDiv d_hw, d_w, d_cc, d_s, d_ntiles, d_cblocks;
  TcClass cls[kMaxClasses];
};

constexpr int kEpiWarps = 4, kProdWarps = 8;
constexpr int kThreadsP = (kEpiWarps + kProdWarps + 1) * 32;   // 416
constexpr int kStagePitch = 36;                                // floats per staged row (32 + 4: conflict-free)

template <int MODE>
__device__ __forceinline__ int tile_class(const TcP& p, int mt) {
  int ci = 0;
  if (MODE == 2) {
#pragma unroll
    for (int j = 1; j < kMaxClasses; ++j)
      if (j < p.ncls && mt >= p.cls[j].tile_begin) ci = j;
  }
  return ci;
}

template <int MODE>
__global__ void __launch_bounds__(kThreadsP, 1)
conv_tc_persist_kernel(const float* __restrict__ src, const __nv_bfloat16* __restrict__ b_hi,
                       const __nv_bfloat16* __restrict__ b_lo, float* __restrict__ out,
                       const float* __restrict__ bias, const float* __restrict__ residual,
                       const __grid_constant__ TcP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const TcGeom& g = p.g;
  const int BN = p.BN;
  const uint32_t a_bytes = TM * 128, b_bytes = (uint32_t)BN * 128;
  uint8_t* smem_a = smem;                                           // n_stages x (hi, lo)
  uint8_t* smem_b = smem + (size_t)p.n_stages * 2 * a_bytes;        // n_bslots x (hi, lo)
  float* stage_all = reinterpret_cast<float*>(smem_b + (size_t)p.n_bslots * 2 * b_bytes);
  long long* rowoff_all = reinterpret_cast<long long*>(stage_all + kEpiWarps * 32 * kStagePitch);
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(&full_bar[s], kProdWarps * 32);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], kEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarps + kProdWarps) tmem_alloc(&tmem_base_s, (uint32_t)(2 * p.acc_cols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int first_tile = blockIdx.x, tile_step = gridDim.x;

  if (warp >= kEpiWarps && warp < kEpiWarps + kProdWarps) {
    // =================================== producers ===================================
    const int pt_ = tid - kEpiWarps * 32;
    const int l8 = pt_ & 7, rgrp = pt_ >> 3;          // 32-byte slice of the row, row group (rows rgrp + 32*i)
    const int CC = (MODE == 0) ? g.C : g.K;           // channels of the gathered tensor
    const uint32_t chunk_off = (((uint32_t)l8) ^ (uint32_t)(rgrp & 7)) << 4;
    auto tile_nk = [&](int tile) -> int {
      if (MODE != 2) return p.nk;
      const int mt = (int)fdiv((uint32_t)tile, p.d_ntiles);
      return p.cls[tile_class<MODE>(p, mt)].ntaps * p.cblocks;
    };
    auto issue_loads_a = [&](int tile, int ks, float4 (&av)[8]) {
      const int mt = (int)fdiv((uint32_t)tile, p.d_ntiles);
      const float* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
      if (MODE != 2) {
        const int kk = ks * BK + (l8 >> 1) * 16;      // this lane's 16-channel sub-chunk: inside one filter tap
        const int tap = (int)fdiv((uint32_t)kk, p.d_cc);
        const int c = kk - tap * CC + (l8 & 1) * 8;
        const int r = (int)fdiv((uint32_t)tap, p.d_s), q = tap - r * g.S;
        const int hw = (MODE == 0) ? g.P * g.Q : g.H * g.W;
        const int wq = (MODE == 0) ? g.Q : g.W;
        const bool unit_stride = g.sh == 1 && g.sw == 1;
        if (kk < p.Kdim) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int m = mt * TM + rgrp + 32 * i;
            if (m < p.M) {
              const int n_ = (int)fdiv((uint32_t)m, p.d_hw);
              const int rem = m - n_ * hw;
              const int y = (int)fdiv((uint32_t)rem, p.d_w), x = rem - y * wq;
              if (MODE == 0) {
                const int ih = y * g.sh - g.pt + r, iw = x * g.sw - g.pl + q;
                if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) ptr[i] = src + (((size_t)n_ * g.H + ih) * g.W + iw) * g.C + c;
              } else {
                const int th = y + g.pt - r, tw = x + g.pl - q;
                if (th >= 0 && tw >= 0) {
                  if (unit_stride) {
                    if (th < g.P && tw < g.Q) ptr[i] = src + (((size_t)n_ * g.P + th) * g.Q + tw) * g.K + c;
                  } else {
                    const int oh = th / g.sh, ow = tw / g.sw;
                    if (oh * g.sh == th && ow * g.sw == tw && oh < g.P && ow < g.Q)
                      ptr[i] = src + (((size_t)n_ * g.P + oh) * g.Q + ow) * g.K + c;
                  }
                }
              }
            }
          }
        }
      } else {
        const TcClass& k = p.cls[tile_class<MODE>(p, mt)];
        const int tap_i = (int)fdiv((uint32_t)ks, p.d_cblocks);
        const int c = (ks - tap_i * p.cblocks) * BK + l8 * 8;
        const int eh = k.eh[tap_i], ew = k.ew[tap_i];
        const int hwc = k.Hc * k.Wc;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int mc = (mt - k.tile_begin) * TM + rgrp + 32 * i;
          if (mc < k.Mc) {
            const int n_ = (int)fdiv((uint32_t)mc, k.d_hw);
            const int rem = mc - n_ * hwc;
            const int y = (int)fdiv((uint32_t)rem, k.d_w), x = rem - y * k.Wc;
            const int oh = y + eh, ow = x + ew;
            if (oh >= 0 && oh < g.P && ow >= 0 && ow < g.Q) ptr[i] = src + (((size_t)n_ * g.P + oh) * g.Q + ow) * g.K + c;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        av[2 * i] = ptr[i] ? __ldg(reinterpret_cast<const float4*>(ptr[i])) : make_float4(0.f, 0.f, 0.f, 0.f);
        av[2 * i + 1] = ptr[i] ? __ldg(reinterpret_cast<const float4*>(ptr[i]) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    uint32_t it = 0;                                   // global k-stage counter of this CTA (ring position)
    auto do_stage = [&](int tile, int ks, bool first, const float4 (&av)[8]) {
      const uint32_t s = it % (uint32_t)p.n_stages;
      mbar_wait(&empty_bar[s], ((it / (uint32_t)p.n_stages) & 1u) ^ 1u);   // slot free?
      uint8_t* sa = smem_a + (size_t)s * 2 * a_bytes;
      if (!p.b_stationary || first) {
        const int mt = (int)fdiv((uint32_t)tile, p.d_ntiles);
        const int n0 = (tile - mt * p.n_tiles) * BN;
        uint8_t* sb = smem_b + (size_t)(p.b_stationary ? ks : (int)s) * 2 * b_bytes;
        size_t kcol;
        if (MODE == 2) {
          const TcClass& k = p.cls[tile_class<MODE>(p, mt)];
          const int tap_i = (int)fdiv((uint32_t)ks, p.d_cblocks);
          kcol = (size_t)k.tap[tap_i] * g.K + (size_t)(ks - tap_i * p.cblocks) * BK + l8 * 8;
        } else {
          kcol = (size_t)ks * BK + l8 * 8;
        }
        for (int br = rgrp; br < BN; br += 32) {
          const bool ok = n0 + br < p.Ng;
          const size_t off = (size_t)(ok ? n0 + br : 0) * p.Kpad + kcol;
          const uint32_t dst = smem_u32(sb + (size_t)br * 128 + chunk_off);
          const uint32_t nbytes = ok ? 16u : 0u;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(b_hi + off), "r"(nbytes) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + b_bytes), "l"(b_lo + off), "r"(nbytes) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint2 h0, l0, h1, l1;
        split4(av[2 * i], h0, l0);
        split4(av[2 * i + 1], h1, l1);
        uint8_t* rowp = sa + (size_t)(rgrp + 32 * i) * 128 + chunk_off;
        *reinterpret_cast<uint4*>(rowp) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4*>(rowp + a_bytes) = make_uint4(l0.x, l0.y, l1.x, l1.y);
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async_smem();   // generic-proxy + cp.async writes -> visible to the tensor core (async proxy)
      mbar_arrive(&full_bar[s]);
      ++it;
    };
    // flat (tile, k-stage) iteration with a two-deep register ping-pong
    int t0 = first_tile, k0 = 0, nk0 = (t0 < p.total_tiles) ? tile_nk(t0) : 0;
    while (t0 < p.total_tiles && nk0 == 0) { t0 += tile_step; nk0 = (t0 < p.total_tiles) ? tile_nk(t0) : 0; }
    auto advance = [&](int& t, int& k, int& nkt) {
      if (++k >= nkt) {
        k = 0;
        do {
          t += tile_step;
          nkt = (t < p.total_tiles) ? tile_nk(t) : 0;
        } while (t < p.total_tiles && nkt == 0);     // zero-tap classes have no main loop
      }
    };
    if (t0 < p.total_tiles) {
      float4 a0[8], a1[8];
      issue_loads_a(t0, k0, a0);
      while (true) {
        int t1 = t0, k1 = k0, nk1 = nk0;
        advance(t1, k1, nk1);
        const bool v1 = t1 < p.total_tiles;
        if (v1) issue_loads_a(t1, k1, a1);
        do_stage(t0, k0, t0 == first_tile, a0);
        if (!v1) break;
        int t2 = t1, k2 = k1, nk2 = nk1;
        advance(t2, k2, nk2);
        const bool v2 = t2 < p.total_tiles;
        if (v2) issue_loads_a(t2, k2, a0);
        do_stage(t1, k1, t1 == first_tile, a1);
        if (!v2) break;
        t0 = t2; k0 = k2; nk0 = nk2;
      }
    }
  } else if (warp == kEpiWarps + kProdWarps) {
    // =================================== MMA issuer ===================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(TM, BN, 0, 0);
      uint32_t it = 0, tcount = 0;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++tcount) {
        int nk = p.nk;
        if (MODE == 2) nk = p.cls[tile_class<MODE>(p, (int)fdiv((uint32_t)tile, p.d_ntiles))].ntaps * p.cblocks;
        const uint32_t buf = tcount & 1u;
        mbar_wait(&tempty_bar[buf], ((tcount >> 1) & 1u) ^ 1u);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)p.acc_cols;
        for (int ks = 0; ks < nk; ++ks, ++it) {
          const uint32_t s = it % (uint32_t)p.n_stages;
          mbar_wait(&full_bar[s], (it / (uint32_t)p.n_stages) & 1u);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem_a + (size_t)s * 2 * a_bytes), a_lo = a_hi + a_bytes;
          const uint32_t bh = smem_u32(smem_b + (size_t)(p.b_stationary ? ks : (int)s) * 2 * b_bytes), bl = bh + b_bytes;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t dah = make_smem_desc(a_hi + kk * 32, 16, 1024);
            const uint64_t dal = make_smem_desc(a_lo + kk * 32, 16, 1024);
            const uint64_t dbh = make_smem_desc(bh + kk * 32, 16, 1024);
            const uint64_t dbl = make_smem_desc(bl + kk * 32, 16, 1024);
            umma_bf16(d_tmem, dah, dbh, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
            umma_bf16(d_tmem, dah, dbl, idesc, 1u);
            umma_bf16(d_tmem, dal, dbh, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);   // frees the A (and ring B) stage when these MMAs have completed
        }
        if (nk > 0) umma_commit(&tfull_bar[buf]);   // accumulator of this tile complete
        else mbar_arrive(&tfull_bar[buf]);          // zero-tap class: nothing was issued, the epilogue writes zeros
      }
    }
  } else {
    // =================================== epilogue (warps 0-3) ===================================
    float* stg = stage_all + (size_t)warp * 32 * kStagePitch;
    long long* rowoff = rowoff_all + warp * 32;
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    const float* extra = residual ? residual : (p.accumulate ? out : nullptr);
    uint32_t tcount = 0;
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++tcount) {
      const int mt = (int)fdiv((uint32_t)tile, p.d_ntiles);
      const int n0 = (tile - mt * p.n_tiles) * BN;
      bool zero_tile = false;
      {   // global element offset of this lane's row (-1: beyond the problem)
        long long off = -1;
        if (MODE != 2) {
          const int m = mt * TM + warp * 32 + lane;
          if (m < p.M) off = (long long)m * p.Ng;
        } else {
          const TcClass& k = p.cls[tile_class<MODE>(p, mt)];
          zero_tile = k.ntaps == 0;
          const int mc = (mt - k.tile_begin) * TM + warp * 32 + lane;
          if (mc < k.Mc) {
            const int hwc = k.Hc * k.Wc;
            const int n_ = (int)fdiv((uint32_t)mc, k.d_hw);
            const int rem = mc - n_ * hwc;
            const int y = (int)fdiv((uint32_t)rem, k.d_w), x = rem - y * k.Wc;
            off = (((long long)n_ * g.H + (k.ph + y * g.sh)) * g.W + (k.pw + x * g.sw)) * g.C;
          }
        }
        rowoff[lane] = off;
      }
      const uint32_t buf = tcount & 1u;
      mbar_wait(&tfull_bar[buf], (tcount >> 1) & 1u);
      tc_fence_after();
      __syncwarp();
      const uint32_t t_addr = tmem_base + buf * (uint32_t)p.acc_cols + lane_base;
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        if (!zero_tile) {
          tmem_ld_32x32(t_addr + (uint32_t)c0, r);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (c0 + 32 >= BN) {                         // last read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(&tempty_bar[buf]);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(stg + lane * kStagePitch + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        __syncwarp();
        // rows 4*j + (lane >> 3), 16-byte chunk (lane & 7): 8 lanes write one row's 128 contiguous bytes
        const int cv = c0 + (lane & 7) * 4;
        const bool cok = cv < BN && n0 + cv + 3 < p.Ng;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias && cok) bb = __ldg(reinterpret_cast<const float4*>(bias + n0 + cv));
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4 xv[4];
          long long ro[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int rr = 4 * (half * 4 + u) + (lane >> 3);
            ro[u] = cok ? rowoff[rr] : -1;
            xv[u] = (extra && ro[u] >= 0) ? *reinterpret_cast<const float4*>(extra + ro[u] + n0 + cv)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (ro[u] < 0) continue;
            const int rr = 4 * (half * 4 + u) + (lane >> 3);
            float4 v = *reinterpret_cast<const float4*>(stg + rr * kStagePitch + (lane & 7) * 4);
            if (bias) { v.x = __fadd_rn(v.x, bb.x); v.y = __fadd_rn(v.y, bb.y); v.z = __fadd_rn(v.z, bb.z); v.w = __fadd_rn(v.w, bb.w); }
            if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            if (extra) {   // fused residual add (resnet_model.py:199,314) or dx += (gradient accumulation)
              v.x = __fadd_rn(v.x, xv[u].x); v.y = __fadd_rn(v.y, xv[u].y); v.z = __fadd_rn(v.z, xv[u].z); v.w = __fadd_rn(v.w, xv[u].w);
            }
            *reinterpret_cast<float4*>(out + ro[u] + n0 + cv) = v;
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + kProdWarps) tmem_dealloc(tmem_base, (uint32_t)(2 * p.acc_cols));
}

// ---------------------------------------------------------------------------------------------
// wgrad on tensor cores:  dW[kf][co] = sum_pix im2col(x)[pix][kf] * dy[pix][co]
// GEMM with M = kf = (r,s,c) tile of 128, N = cout tile, K = pixels; BOTH operands are MN-major
// (channels are the contiguous axis of NHWC): smem tiles are [k/8][mn/64][k%8] rows of 64 bf16
// (128 B, SWIZZLE_128B), descriptor LBO = 1024 (next 64-wide MN block), SBO = (tile_mn/64)*1024
// (next group of 8 pixels), one K=16 MMA step = 2*SBO — conventions pinned by tests/test_tc_gpu.py.
// Split-K over pixel ranges (grid.z); partials go to a workspace and are reduced in fixed order.
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ partial,
                     TcGeom g, int Mtot, int Npix, int pix_per_split, int BN, int n_stages) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = BK * TM * 2, b_bytes = (uint32_t)BK * BN * 2;   // one bf16 tile
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], accum_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * BN;
  const int pbeg = blockIdx.z * pix_per_split;
  const int pend = min(Npix, pbeg + pix_per_split);
  const int nk = (pend - pbeg + BK - 1) / BK;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < BN) tmem_cols <<= 1;
  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full_bar[s], kProducerThreads);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(&tmem_base_s, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int mbA = TM / 64, mbB = BN / 64;

  if (warp < 8) {
    // producers: 16 lanes cover one pixel's 128-wide MN extent (512 contiguous bytes of NHWC fp32), thread t
    // owns elements [8*(t&15), +8) of pixels (t>>4) + 16*i, i < 4, of BOTH operands: 8 fp32 -> one 16-byte
    // bf16 chunk each for the hi and lo tiles.
    const int l16 = tid & 15, pgrp = tid >> 4;
    // A: kf = m0 + 8*l16 .. +7 lies inside one filter tap (C % 16 == 0): decode is stage-invariant
    const int kf = m0 + l16 * 8;
    const bool a_ok = kf < Mtot;
    const int a_tap = kf / g.C, a_c = kf - a_tap * g.C;
    const int a_r = a_tap / g.S, a_q = a_tap - a_r * g.S;
    // B: couts n0 + 8*l16 .. +7
    const int bco = n0 + l16 * 8;
    const bool b_okc = l16 * 8 < BN && bco < g.K;
    const uint32_t e = (uint32_t)(l16 * 8);                 // element offset inside the 128-wide tile
    const uint32_t mblk = e >> 6, chunk = (e & 63) >> 3;
    auto issue_loads = [&](int ks, float4 (&av)[8], float4 (&bv)[8]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pix = pbeg + ks * BK + pgrp + 16 * i;
        const float* pa = nullptr;
        const float* pb = nullptr;
        if (pix < pend) {
          const int pq = g.P * g.Q;
          const int pn = pix / pq;
          const int rem = pix - pn * pq;
          const int oh = rem / g.Q, ow = rem - oh * g.Q;
          if (a_ok) {
            const int ih = oh * g.sh - g.pt + a_r, iw = ow * g.sw - g.pl + a_q;
            if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) pa = x + (((size_t)pn * g.H + ih) * g.W + iw) * g.C + a_c;
          }
          if (b_okc) pb = dy + (size_t)pix * g.K + bco;
        }
        av[2 * i] = pa ? __ldg(reinterpret_cast<const float4*>(pa)) : make_float4(0.f, 0.f, 0.f, 0.f);
        av[2 * i + 1] = pa ? __ldg(reinterpret_cast<const float4*>(pa) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        bv[2 * i] = pb ? __ldg(reinterpret_cast<const float4*>(pb)) : make_float4(0.f, 0.f, 0.f, 0.f);
        bv[2 * i + 1] = pb ? __ldg(reinterpret_cast<const float4*>(pb) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store_stage = [&](int s, const float4 (&av)[8], const float4 (&bv)[8]) {
      uint8_t* st = smem + (size_t)s * stage_bytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t k = (uint32_t)(pgrp + 16 * i), k8 = k & 7, kb = k >> 3;
        uint2 h0, l0, h1, l1;
        split4(av[2 * i], h0, l0);
        split4(av[2 * i + 1], h1, l1);
        uint8_t* hi = st + (size_t)(kb * (uint32_t)(mbA * 8) + mblk * 8 + k8) * 128 + ((chunk ^ k8) << 4);
        *reinterpret_cast<uint4*>(hi) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4*>(hi + a_bytes) = make_uint4(l0.x, l0.y, l1.x, l1.y);
        if (l16 * 8 < BN) {
          split4(bv[2 * i], h0, l0);
          split4(bv[2 * i + 1], h1, l1);
          uint8_t* bh = st + 2 * a_bytes + (size_t)(kb * (uint32_t)(mbB * 8) + mblk * 8 + k8) * 128 + ((chunk ^ k8) << 4);
          *reinterpret_cast<uint4*>(bh) = make_uint4(h0.x, h0.y, h1.x, h1.y);
          *reinterpret_cast<uint4*>(bh + b_bytes) = make_uint4(l0.x, l0.y, l1.x, l1.y);
        }
      }
    };
    for (int ks = 0; ks < nk; ++ks) {
      const int s = ks % n_stages;
      float4 av[8], bv[8];
      issue_loads(ks, av, bv);
      mbar_wait(&empty_bar[s], (((uint32_t)(ks / n_stages)) & 1u) ^ 1u);
      store_stage(s, av, bv);
      fence_proxy_async_smem();
      mbar_arrive(&full_bar[s]);
    }
    // ---- epilogue: D rows = kf, columns = cout -> partial[z][kf][cout]
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const int erow = (warp & 3) * 32 + (tid & 31);
    const int em = m0 + erow;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
    const int cbeg = (warp >> 2) * (BN / 2), cend = cbeg + BN / 2;
    float* outp = partial + (size_t)blockIdx.z * Mtot * g.K;
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + lane_base + (uint32_t)c0, r);
      if (em < Mtot) {
        float* o = outp + (size_t)em * g.K + n0 + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (n0 + c0 + j + 3 < g.K)
            *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                            __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
    }
  } else if (warp == 8) {
    if ((tid & 31) == 0) {
      const uint32_t idesc = make_idesc_bf16(TM, BN, 1, 1);
      const uint32_t sbo_a = (uint32_t)mbA * 1024u, sbo_b = (uint32_t)mbB * 1024u;
      for (int ks = 0; ks < nk; ++ks) {
        const int s = ks % n_stages;
        mbar_wait(&full_bar[s], ((uint32_t)(ks / n_stages)) & 1u);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_hi = base, a_lo = base + a_bytes, bh = base + 2 * a_bytes, bl = bh + b_bytes;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t dah = make_smem_desc(a_hi + kk * 2 * sbo_a, 1024, sbo_a);
          const uint64_t dal = make_smem_desc(a_lo + kk * 2 * sbo_a, 1024, sbo_a);
          const uint64_t dbh = make_smem_desc(bh + kk * 2 * sbo_b, 1024, sbo_b);
          const uint64_t dbl = make_smem_desc(bl + kk * 2 * sbo_b, 1024, sbo_b);
          umma_bf16(tmem_base, dah, dbh, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
          umma_bf16(tmem_base, dah, dbl, idesc, 1u);
          umma_bf16(tmem_base, dal, dbh, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&accum_bar);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, tmem_cols);
}

__global__ void __launch_bounds__(256)
tc_splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int64_t n, int splits) {
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i >= n) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int z = 0; z < splits; ++z) {
    const float4 v = *reinterpret_cast<const float4*>(partial + (size_t)z * n + i);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  *reinterpret_cast<float4*>(out + i) = s;
}

inline int tc_wgrad_splits(const TcGeom& g, int BN, int* pix_per_split) {
  const int Mtot = g.R * g.S * g.C, Npix = g.N * g.P * g.Q;
  const int tiles = ((Mtot + TM - 1) / TM) * ((g.K + BN - 1) / BN);
  int splits = (2 * PF_NUM_SMS + tiles - 1) / tiles;
  const int max_by_k = (Npix + 8 * BK - 1) / (8 * BK);
  if (splits > max_by_k) splits = max_by_k;
  if (splits > PF_CONV_TC_WGRAD_MAX_SPLITS) splits = PF_CONV_TC_WGRAD_MAX_SPLITS;
  if (splits < 1) splits = 1;
  int pps = (Npix + splits - 1) / splits;
  pps = (pps + BK - 1) / BK * BK;
  *pix_per_split = pps;
  return (Npix + pps - 1) / pps;
}

// ---- weight preparation: fp32 HWIO [R,S,C,K] -> split bf16, K-major for both passes
//   fwd  : [K (cout)][Kpad_f],  k = (r*S + s)*C + c
//   dgrad: [C (cin) ][Kpad_d],  k = (r*S + s)*K + co
__global__ void __launch_bounds__(256)
tc_prep_weight_kernel(const float* __restrict__ w, int RS, int C, int K, int kpad_f, int kpad_d,
                      __nv_bfloat16* __restrict__ f_hi, __nv_bfloat16* __restrict__ f_lo,
                      __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo) {
  const int64_t total = (int64_t)RS * C * K;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += stride) {
    const int co = (int)(i % K);
    const int64_t t = i / K;
    const int c = (int)(t % C);
    const int rs = (int)(t / C);
    const float v = __ldg(w + i);
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    const size_t fo = (size_t)co * kpad_f + (size_t)rs * C + c;
    f_hi[fo] = h;
    f_lo[fo] = l;
    if (d_hi) {
      const size_t dof = (size_t)c * kpad_d + (size_t)rs * K + co;
      d_hi[dof] = h;
      d_lo[dof] = l;
    }
  }
}

int tc_geom(const pf_conv_desc* d, TcGeom* g, const char* who) {
  PF_REQUIRE(d != nullptr, "%s: null descriptor", who);
  PF_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->c > 0 && d->k > 0 && d->r > 0 && d->s > 0 && d->p > 0 &&
                 d->q > 0 && d->stride_h > 0 && d->stride_w > 0 && d->pad_t >= 0 && d->pad_l >= 0,
             "%s: non-positive dimension in conv descriptor", who);
  *g = TcGeom{d->n, d->h, d->w, d->c, d->k, d->r, d->s, d->p, d->q, d->stride_h, d->stride_w, d->pad_t, d->pad_l};
  return PF_OK;
}

inline int pad64(int64_t k) { return (int)((k + 63) / 64 * 64); }

template <int MODE>
int launch_tc_v3(const TcGeom& g, const float* src, const void* b_hi, const void* b_lo, float* out, int accumulate,
                 const float* bias, int relu, const float* residual, cudaStream_t st, const char* who) {
  const int64_t M64 = (MODE == 0) ? (int64_t)g.N * g.P * g.Q : (int64_t)g.N * g.H * g.W;
  PF_REQUIRE(M64 < (1ll << 31), "%s: too many rows", who);
  const int M = (int)M64;
  const int Ng = (MODE == 0) ? g.K : g.C;
  const int Kdim = (MODE == 0) ? g.R * g.S * g.C : g.R * g.S * g.K;
  const int Kpad = pad64(Kdim);
  int BN = Ng >= 128 ? 128 : (Ng >= 64 ? 64 : (Ng >= 32 ? 32 : 16));
  const int nk = Kpad / BK;
  int stages = nk < 3 ? (nk < 2 ? 1 : 2) : 3;
  if (BN <= 64 && nk >= 4) stages = 4;
  size_t smem = (size_t)stages * (2 * TM * 128 + 2 * BN * 128);
  const size_t epi = (size_t)TM * (BN + 4) * 4;          // the epilogue stages the fp32 tile in the same memory
  if (smem < epi) smem = epi;
  smem += 1024;
  dim3 grid((M + TM - 1) / TM, (Ng + BN - 1) / BN);
  if (nk <= 1) {
    auto kern = conv_tc_kernel<MODE, 4>;
    PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kern<<<grid, 4 * 32 + 32, smem, st>>>(src, (const __nv_bfloat16*)b_hi, (const __nv_bfloat16*)b_lo, out, g, M, Ng,
                                          Kdim, Kpad, BN, stages, accumulate, bias, relu, residual);
  } else {
    auto kern = conv_tc_kernel<MODE, 8>;
    PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kern<<<grid, 8 * 32 + 32, smem, st>>>(src, (const __nv_bfloat16*)b_hi, (const __nv_bfloat16*)b_lo, out, g, M, Ng,
                                          Kdim, Kpad, BN, stages, accumulate, bias, relu, residual);
  }
  PF_CHECK_LAUNCH(who);
  return PF_OK;
}

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

constexpr int kSmemLimit = 232448;   // 227 KB opt-in maximum of dynamic shared memory per CTA on sm_100

template <int MODE>
int launch_persist(const TcGeom& g, TcP& p, const float* src, const void* b_hi, const void* b_lo, float* out,
                   const float* bias, const float* residual, cudaStream_t st, const char* who) {
  const int Ng = p.Ng;
  // ---- tile width: wide tiles read A once per 256 channels, but quantise worse over the 148 SMs
  int BN;
  if (Ng >= 256) {
    const int64_t t256 = (int64_t)p.m_tiles * ((Ng + 255) / 256), t128 = (int64_t)p.m_tiles * ((Ng + 127) / 128);
    const double c256 = (double)((t256 + PF_NUM_SMS - 1) / PF_NUM_SMS) * 1.8;
    const double c128 = (double)((t128 + PF_NUM_SMS - 1) / PF_NUM_SMS);
    BN = c256 <= c128 ? 256 : 128;
  } else {
    BN = Ng >= 128 ? 128 : (Ng >= 64 ? 64 : (Ng >= 32 ? 32 : 16));
  }
  const int forced = env_int("PF_TC_BN", 0);           // development knob
  if (forced >= 16 && forced <= 256 && forced <= ((Ng + 15) / 16) * 16 && (forced & (forced - 1)) == 0) BN = forced;
  p.BN = BN;
  p.n_tiles = (Ng + BN - 1) / BN;
  p.total_tiles = p.m_tiles * p.n_tiles;
  p.d_ntiles = make_fastdiv((uint32_t)p.n_tiles);
  p.acc_cols = 32;
  while (p.acc_cols < BN) p.acc_cols <<= 1;
  // ---- shared memory plan
  const int a_stage = 2 * TM * 128, b_slot = 2 * BN * 128;
  const int fixed = 1024 + kEpiWarps * 32 * kStagePitch * 4 + kEpiWarps * 32 * 8 + 256;
  const int budget = kSmemLimit - fixed;
  int max_nk = p.nk;
  if (MODE == 2) {
    max_nk = 0;
    for (int c = 0; c < p.ncls; ++c) max_nk = std::max(max_nk, p.cls[c].ntaps * p.cblocks);
  }
  p.b_stationary = 0;
  if (MODE != 2 && p.n_tiles == 1 && p.nk <= kMaxStages * 4 && (int64_t)p.nk * b_slot + 2 * a_stage <= budget &&
      env_int("PF_TC_STATIONARY", 1)) {
    p.b_stationary = 1;
    p.n_bslots = p.nk;
    p.n_stages = std::min(kMaxStages, (budget - p.nk * b_slot) / a_stage);
  } else {
    p.n_stages = std::min(kMaxStages, budget / (a_stage + b_slot));
    p.n_bslots = p.n_stages;
  }
  PF_REQUIRE(p.n_stages >= 2 || max_nk <= 1, "%s: shared-memory plan failed (BN %d)", who, BN);
  if (p.n_stages < 1) p.n_stages = 1;
  const size_t smem = (size_t)p.n_stages * a_stage + (size_t)p.n_bslots * b_slot + fixed;
  auto kern = conv_tc_persist_kernel<MODE>;
  PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(p.total_tiles, PF_NUM_SMS);
  kern<<<grid, kThreadsP, smem, st>>>(src, (const __nv_bfloat16*)b_hi, (const __nv_bfloat16*)b_lo, out, bias, residual, p);
  PF_CHECK_LAUNCH(who);
  return PF_OK;
}

template <int MODE>
int launch_tc(const TcGeom& g, const float* src, const void* b_hi, const void* b_lo, float* out, int accumulate,
              const float* bias, int relu, const float* residual, cudaStream_t st, const char* who) {
  static const int impl = env_int("PF_TC_IMPL", 4);
  if (impl == 3) return launch_tc_v3<MODE>(g, src, b_hi, b_lo, out, accumulate, bias, relu, residual, st, who);
  const int64_t M64 = (MODE == 0) ? (int64_t)g.N * g.P * g.Q : (int64_t)g.N * g.H * g.W;
  PF_REQUIRE(M64 < (1ll << 31), "%s: too many rows", who);
  TcP p;
  memset(&p, 0, sizeof(p));
  p.g = g;
  p.M = (int)M64;
  p.Ng = (MODE == 0) ? g.K : g.C;
  p.Kdim = (MODE == 0) ? g.R * g.S * g.C : g.R * g.S * g.K;
  p.Kpad = pad64(p.Kdim);
  p.nk = p.Kpad / BK;
  p.accumulate = accumulate;
  p.relu = relu;
  const int CC = (MODE == 0) ? g.C : g.K;
  const int hw = (MODE == 0) ? g.P * g.Q : g.H * g.W, wq = (MODE == 0) ? g.Q : g.W;
  p.d_hw = make_fastdiv((uint32_t)hw);
  p.d_w = make_fastdiv((uint32_t)wq);
  p.d_cc = make_fastdiv((uint32_t)CC);
  p.d_s = make_fastdiv((uint32_t)g.S);
  p.cblocks = std::max(1, g.K / BK);
  p.d_cblocks = make_fastdiv((uint32_t)p.cblocks);
  p.m_tiles = (p.M + TM - 1) / TM;
  if (MODE == 1 && (g.sh > 1 || g.sw > 1) && g.sh * g.sw <= kMaxClasses && g.R * g.S <= kMaxTaps && g.K % BK == 0 &&
      env_int("PF_TC_CLASSES", 1)) {
    // pixel-parity classes of the strided dgrad
    int tiles = 0;
    for (int ph = 0; ph < g.sh; ++ph)
      for (int pw = 0; pw < g.sw; ++pw) {
        if (ph >= g.H || pw >= g.W) continue;
        TcClass& k = p.cls[p.ncls];
        k.ph = ph;
        k.pw = pw;
        k.Hc = (g.H - ph + g.sh - 1) / g.sh;
        k.Wc = (g.W - pw + g.sw - 1) / g.sw;
        k.Mc = g.N * k.Hc * k.Wc;
        k.tile_begin = tiles;
        k.d_hw = make_fastdiv((uint32_t)(k.Hc * k.Wc));
        k.d_w = make_fastdiv((uint32_t)k.Wc);
        k.ntaps = 0;
        for (int r = 0; r < g.R; ++r) {
          const int nh = ph + g.pt - r;
          if (((nh % g.sh) + g.sh) % g.sh != 0) continue;
          for (int q = 0; q < g.S; ++q) {
            const int nw = pw + g.pl - q;
            if (((nw % g.sw) + g.sw) % g.sw != 0) continue;
            k.eh[k.ntaps] = (int8_t)(nh / g.sh);
            k.ew[k.ntaps] = (int8_t)(nw / g.sw);
            k.tap[k.ntaps] = (int8_t)(r * g.S + q);
            ++k.ntaps;
          }
        }
        tiles += (k.Mc + TM - 1) / TM;
        ++p.ncls;
      }
    p.m_tiles = tiles;
    return launch_persist<2>(g, p, src, b_hi, b_lo, out, bias, residual, st, who);
  }
  return launch_persist<MODE>(g, p, src, b_hi, b_lo, out, bias, residual, st, who);
}

}  // namespace

extern "C" {

int pf_conv2d_tc_supported(const pf_conv_desc* d) {
  if (!d) return 0;
  return (d->c % 16 == 0) && (d->k % 16 == 0) && d->c >= 16 && d->k >= 16;
}

int64_t pf_conv2d_tc_weight_elems(const pf_conv_desc* d, int dgrad) {
  if (!d) return 0;
  if (dgrad) return (int64_t)d->c * pad64((int64_t)d->r * d->s * d->k);
  return (int64_t)d->k * pad64((int64_t)d->r * d->s * d->c);
}

int pf_conv2d_tc_prep_weight(const pf_conv_desc* d, const float* w_dev, void* fwd_hi_dev, void* fwd_lo_dev,
                             void* dgrad_hi_dev, void* dgrad_lo_dev, void* stream) {
  TcGeom g;
  int rc = tc_geom(d, &g, "pf_conv2d_tc_prep_weight");
  if (rc) return rc;
  PF_REQUIRE(w_dev && fwd_hi_dev && fwd_lo_dev, "pf_conv2d_tc_prep_weight: null pointer");
  PF_REQUIRE((dgrad_hi_dev == nullptr) == (dgrad_lo_dev == nullptr), "pf_conv2d_tc_prep_weight: dgrad buffers come in pairs");
  const int kpf = pad64((int64_t)g.R * g.S * g.C), kpd = pad64((int64_t)g.R * g.S * g.K);
  cudaStream_t st = (cudaStream_t)stream;
  // padding columns must be zero: the buffers are zero-filled once by the caller at allocation time
  const int64_t total = (int64_t)g.R * g.S * g.C * g.K;
  int64_t blocks = (total + 255) / 256;
  if (blocks > PF_NUM_SMS * 8) blocks = PF_NUM_SMS * 8;
  tc_prep_weight_kernel<<<(unsigned)blocks, 256, 0, st>>>(w_dev, g.R * g.S, g.C, g.K, kpf, kpd,
                                                         (__nv_bfloat16*)fwd_hi_dev, (__nv_bfloat16*)fwd_lo_dev,
                                                         (__nv_bfloat16*)dgrad_hi_dev, (__nv_bfloat16*)dgrad_lo_dev);
  PF_CHECK_LAUNCH("pf_conv2d_tc_prep_weight");
  return PF_OK;
}

int pf_conv2d_tc_fwd(const pf_conv_desc* d, const float* x_dev, const void* w_hi_dev, const void* w_lo_dev,
                     const float* bias_dev, int relu, const float* residual_dev, float* y_dev, void* stream) {
  TcGeom g;
  int rc = tc_geom(d, &g, "pf_conv2d_tc_fwd");
  if (rc) return rc;
  PF_REQUIRE(pf_conv2d_tc_supported(d), "pf_conv2d_tc_fwd: Cin and Cout must be multiples of 16");
  PF_REQUIRE(x_dev && w_hi_dev && w_lo_dev && y_dev, "pf_conv2d_tc_fwd: null pointer");
  PF_REQUIRE((((uintptr_t)x_dev | (uintptr_t)y_dev | (uintptr_t)w_hi_dev | (uintptr_t)w_lo_dev) & 15) == 0,
             "pf_conv2d_tc_fwd: 16-byte alignment required");
  PF_REQUIRE(((uintptr_t)residual_dev & 15) == 0, "pf_conv2d_tc_fwd: residual must be 16-byte aligned");
  return launch_tc<0>(g, x_dev, w_hi_dev, w_lo_dev, y_dev, 0, bias_dev, relu, residual_dev, (cudaStream_t)stream,
                      "pf_conv2d_tc_fwd");
}

int pf_conv2d_tc_dgrad(const pf_conv_desc* d, const float* dy_dev, const void* wd_hi_dev, const void* wd_lo_dev,
                       int accumulate, float* dx_dev, void* stream) {
  TcGeom g;
  int rc = tc_geom(d, &g, "pf_conv2d_tc_dgrad");
  if (rc) return rc;
  PF_REQUIRE(pf_conv2d_tc_supported(d), "pf_conv2d_tc_dgrad: Cin and Cout must be multiples of 16");
  PF_REQUIRE(dy_dev && wd_hi_dev && wd_lo_dev && dx_dev, "pf_conv2d_tc_dgrad: null pointer");
  PF_REQUIRE((((uintptr_t)dy_dev | (uintptr_t)dx_dev | (uintptr_t)wd_hi_dev | (uintptr_t)wd_lo_dev) & 15) == 0,
             "pf_conv2d_tc_dgrad: 16-byte alignment required");
  return launch_tc<1>(g, dy_dev, wd_hi_dev, wd_lo_dev, dx_dev, accumulate, nullptr, 0, nullptr, (cudaStream_t)stream,
                      "pf_conv2d_tc_dgrad");
}

int pf_conv2d_tc_wgrad_supported(const pf_conv_desc* d) {
  if (!d) return 0;
  return (d->c % 16 == 0) && (d->k % 64 == 0) && d->c >= 16;
}

int64_t pf_conv2d_tc_wgrad_workspace_bytes(const pf_conv_desc* d) {
  TcGeom g;
  if (!d || tc_geom(d, &g, "pf_conv2d_tc_wgrad_workspace_bytes")) return 0;
  const int BN = g.K >= 128 ? 128 : 64;
  int pps;
  const int splits = tc_wgrad_splits(g, BN, &pps);
  return (int64_t)splits * g.R * g.S * g.C * g.K * 4;
}

int pf_conv2d_tc_wgrad(const pf_conv_desc* d, const float* x_dev, const float* dy_dev, float* ws_dev,
                       float* dw_dev, void* stream) {
  TcGeom g;
  int rc = tc_geom(d, &g, "pf_conv2d_tc_wgrad");
  if (rc) return rc;
  PF_REQUIRE(pf_conv2d_tc_wgrad_supported(d), "pf_conv2d_tc_wgrad: needs Cin %% 16 == 0 and Cout %% 64 == 0");
  PF_REQUIRE(x_dev && dy_dev && ws_dev && dw_dev, "pf_conv2d_tc_wgrad: null pointer");
  PF_REQUIRE((((uintptr_t)x_dev | (uintptr_t)dy_dev | (uintptr_t)ws_dev | (uintptr_t)dw_dev) & 15) == 0,
             "pf_conv2d_tc_wgrad: 16-byte alignment required");
  const int64_t np64 = (int64_t)g.N * g.P * g.Q;
  PF_REQUIRE(np64 < (1ll << 31), "pf_conv2d_tc_wgrad: too many pixels");
  const int Mtot = g.R * g.S * g.C, Npix = (int)np64;
  const int BN = g.K >= 128 ? 128 : 64;
  int pps;
  const int splits = tc_wgrad_splits(g, BN, &pps);
  const int nk = pps / BK;
  const int stages = nk < 3 ? (nk < 2 ? 1 : 2) : 3;
  const size_t smem = (size_t)stages * (2 * BK * TM * 2 + 2 * BK * BN * 2) + 1024;
  PF_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((Mtot + TM - 1) / TM, (g.K + BN - 1) / BN, splits);
  cudaStream_t st = (cudaStream_t)stream;
  conv_tc_wgrad_kernel<<<grid, kThreads, smem, st>>>(x_dev, dy_dev, splits == 1 ? dw_dev : ws_dev, g, Mtot, Npix,
                                                    pps, BN, stages);
  PF_CHECK_LAUNCH("pf_conv2d_tc_wgrad");
  if (splits > 1) {
    const int64_t n = (int64_t)Mtot * g.K;
    tc_splitk_reduce_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(ws_dev, dw_dev, n, splits);
    PF_CHECK_LAUNCH("pf_conv2d_tc_wgrad/reduce");
  }
  return PF_OK;
}

}  // extern "C"
