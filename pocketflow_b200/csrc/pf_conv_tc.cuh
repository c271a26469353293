// pf_conv_tc.cuh — pieces shared by the tcgen05 convolution kernels (pf_conv_tc.cu: cp.async-fed, any channel
// count that is a multiple of 16; pf_conv_tma.cu: TMA-fed, channel counts that are multiples of 64):
// tile constants, geometry, exact division by runtime constants, and the TMEM -> global epilogue.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "pf_common.cuh"
#include "pf_tc_common.cuh"

namespace pfconv {
using namespace pftc;

constexpr int TM = 128;      // GEMM rows per CTA (= TMEM lanes)
constexpr int BK = 64;       // bf16 elements per k-stage (= one 128-byte swizzled row)
constexpr int kSmemLimit = 232448;   // 227 KB opt-in maximum of dynamic shared memory per CTA on sm_100

struct TcGeom {
  int N, H, W, C, K, R, S, P, Q, sh, sw, pt, pl;
};

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

inline int pad64(int64_t k) { return (int)((k + 63) / 64 * 64); }
inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
int tc_geom(const pf_conv_desc* d, TcGeom* g, const char* who);

// ---------------------------------------------------------------------------------------------------------
// Epilogue of one 128 x BN accumulator tile by 4 epilogue warps (the warp with quarter index q = warp % 4 owns TMEM
// lanes [32q, 32q+32)): TMEM -> registers (thread = row, 32 columns) -> per-warp smem transpose -> 128-byte row
// segments to global.
// `extra` (residual / accumulate operand, same indexing as `out`) is prefetched one 32-column chunk ahead,
// and the first chunk is requested BEFORE waiting for the accumulator, so its latency hides behind the main loop.
// EXTRA: 0 = none; 1 = extra operand prefetched one chunk ahead in registers; 2 = extra operand streamed through a
// per-warp cp.async ring in shared memory, kRingDepth chunks (4 KB each) in flight per warp.
// AFF: the accumulator holds a product of INTEGER quantizer levels (SURVEY §7 hard part 1b): with
//   qw[k,c] = s_c (n[k,c] - centre) + o_c,   qa[m,k] = s_a j[m,k]          (s_c = alpha_c / k_w, o_c = beta_c + centre s_c)
// the convolution of the fake-quantized tensors is   s_a s_c * sum_k j (n - centre)  +  s_a o_c * J[m],
// J[m] = sum of the activation levels under the filter window of output row m (computed by the row's thread from
// per-pixel channel sums and handed in as `my_j`).  AFF 1: only the scalar s_a (weight gradient: j (x) dy).
constexpr int kEpiWarps = 4;
constexpr int kStagePitch = 36;                                // floats per staged row (32 + 4: conflict-free)
constexpr int kRingDepth = 4;
constexpr int kRingSlotBytes = 32 * 32 * 4;
constexpr int kEpiFixedBytes = 1024 + kEpiWarps * 32 * kStagePitch * 4 + kEpiWarps * 32 * 8 + kEpiWarps * 32 * 4 + 256;

// TMA-fed kernels: 8 epilogue warps (two per TMEM lane quarter, even / odd 32-column chunks), each with its own staging
// tile, row-offset table, J table and (optional) residual ring
constexpr int kTmaEpiWarps = 8;
constexpr int kTmaEpiFixedBytes = 1024 + kTmaEpiWarps * (32 * kStagePitch * 4 + 32 * 8 + 32 * 4) + 256;

struct EpiAff {
  const float* w_alpha;    // per-bucket alpha = (max - min) + 1e-10 of the weight quantizer (device)
  const float* w_beta;     // per-bucket beta = min
  const float* a_scale;    // device scalar: value of one activation level (or 1 for split-bf16 planes); null = 1
  int per_channel;         // 1: bucket = output channel, 0: one bucket per layer
  float w_rk, w_centre;    // 1 / (2^bits - 1), level subtracted from the stored weight levels
};

template <int EXTRA, int AFF, int RD = kRingDepth>
__device__ __forceinline__ void epilogue_tile_t(uint32_t t_acc, uint64_t* tfull, uint64_t* tempty, uint32_t parity,
                                                bool zero_tile, long long my_row_off, long long* __restrict__ rowoff,
                                                float* __restrict__ stg, float* __restrict__ out,
                                                const float* __restrict__ extra, const float* __restrict__ bias,
                                                int relu, int n0, int BN, int Ng, int q, int lane, uint8_t* ring,
                                                const EpiAff& aff, float my_j, float* __restrict__ jrow,
                                                int c_begin = 0, int c_step = 32,
                                                const float* __restrict__ aff_tab = nullptr) {
  // aff_tab (AFF == 2, TMA-fed kernels): the tile's per-column epilogue constants e1[c] (at c) and e2[c] (at 256 + c)
  // in shared memory, computed once per tile column range instead of being re-derived from global memory per chunk
  // c_begin / c_step: this warp handles the 32-column chunks c_begin, c_begin + c_step, ... (two warps of the same TMEM
  // lane quarter split a tile's columns between them in the TMA-fed kernels: c_step = 64)
  rowoff[lane] = my_row_off;
  if (AFF == 2) jrow[lane] = my_j;
  __syncwarp();
  const int csub = (lane & 7) * 4, rsub = lane >> 3;
  int ro[8];                                     // this lane's 8 rows (4*u + rsub) of the warp's 32, in float4 units
  float jr[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const long long o = rowoff[4 * u + rsub];
    ro[u] = o < 0 ? -1 : (int)(o >> 2);          // row offsets are multiples of 4 elements (channel counts % 16 == 0)
    jr[u] = (AFF == 2) ? jrow[4 * u + rsub] : 0.f;
  }
  const float a_s = (AFF && aff.a_scale && !(AFF == 2 && aff_tab)) ? __ldg(aff.a_scale) : 1.f;
  auto load_extra = [&](int c0, float4 (&xv)[8]) {
    const int cv = c0 + csub;
    const bool cok = cv < BN && n0 + cv + 3 < Ng;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      xv[u] = (cok && ro[u] >= 0) ? *reinterpret_cast<const float4*>(extra + ((size_t)ro[u] << 2) + n0 + cv)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  // ring: every lane copies exactly the 16-byte pieces it will read back itself (no cross-lane hand-off); one
  // commit group per chunk, empty groups past the end keep the wait_group count uniform
  const uint32_t ring_u32 = (EXTRA == 2) ? smem_u32(ring) : 0u;
  auto ring_issue = [&](int c0, int it) {
    if (c0 < BN) {
      const int cv = c0 + csub;
      const bool cok = cv < BN && n0 + cv + 3 < Ng;
      const uint32_t slot = ring_u32 + (uint32_t)(it & (RD - 1)) * kRingSlotBytes;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool ok = cok && ro[u] >= 0;
        const float* src = ok ? extra + ((size_t)ro[u] << 2) + n0 + cv : extra;
        const uint32_t dst = slot + (uint32_t)(((4 * u + rsub) * 32 + csub) * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float4 xa[8], xb[8];                           // dead (eliminated) unless EXTRA == 1
  if (EXTRA == 1) load_extra(c_begin, xa);
  if (EXTRA == 2) {
#pragma unroll
    for (int c = 0; c < RD; ++c) ring_issue(c_begin + c_step * c, c);
  }
  mbar_wait_bounded(tfull, parity);
  tc_fence_after();
  const uint32_t t_addr = t_acc + (((uint32_t)(q * 32)) << 16);
  int it = 0;
  for (int c0 = c_begin; c0 < BN; c0 += c_step, ++it) {
    // per-column constants first: their global loads overlap the TMEM read below
    // rows 4*u + (lane >> 3), 16-byte chunk (lane & 7): 8 lanes write one row's 128 contiguous bytes
    const int cv = c0 + csub;
    const bool cok = cv < BN && n0 + cv + 3 < Ng;
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias && cok) bb = __ldg(reinterpret_cast<const float4*>(bias + n0 + cv));
    float4 e1 = make_float4(a_s, a_s, a_s, a_s), e2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (AFF == 2 && aff_tab) {
      if (cv < BN) {
        e1 = *reinterpret_cast<const float4*>(aff_tab + cv);
        e2 = *reinterpret_cast<const float4*>(aff_tab + 256 + cv);
      }
    } else if (AFF == 2) {
      float4 al, be;
      if (aff.per_channel) {
        al = cok ? __ldg(reinterpret_cast<const float4*>(aff.w_alpha + n0 + cv)) : make_float4(0.f, 0.f, 0.f, 0.f);
        be = cok ? __ldg(reinterpret_cast<const float4*>(aff.w_beta + n0 + cv)) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        const float a0 = __ldg(aff.w_alpha), b0 = __ldg(aff.w_beta);
        al = make_float4(a0, a0, a0, a0);
        be = make_float4(b0, b0, b0, b0);
      }
      const float sx = al.x * aff.w_rk, sy = al.y * aff.w_rk, sz = al.z * aff.w_rk, sw_ = al.w * aff.w_rk;
      e1 = make_float4(sx * a_s, sy * a_s, sz * a_s, sw_ * a_s);
      e2 = make_float4(fmaf(aff.w_centre, sx, be.x) * a_s, fmaf(aff.w_centre, sy, be.y) * a_s,
                       fmaf(aff.w_centre, sz, be.z) * a_s, fmaf(aff.w_centre, sw_, be.w) * a_s);
    }
    // (issuing the NEXT chunk's TMEM read right after the staging stores, to run its latency under this chunk's
    // read-back and global stores, keeps 32 more registers live through that phase: 2-3 KB of spills at the 168
    // registers 10 warps allow — measured at compile time, not pursued)
    uint32_t r[32];
    if (!zero_tile) {
      tmem_ld_32x32(t_addr + (uint32_t)c0, r);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = 0u;
    }
    if (c0 + c_step >= BN) {                     // last read of this accumulator: hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(tempty);
    }
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<uint4*>(stg + lane * kStagePitch + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
    __syncwarp();
    if (EXTRA == 1 && c0 + c_step < BN) load_extra(c0 + c_step, xb);
    if (EXTRA == 2) asm volatile("cp.async.wait_group %0;" ::"n"(RD - 1) : "memory");   // chunk c0 has landed
    const float* slot = reinterpret_cast<const float*>(ring + (size_t)(it & (RD - 1)) * kRingSlotBytes);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 v[4], xr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = *reinterpret_cast<const float4*>(stg + (4 * (4 * half + u) + rsub) * kStagePitch + csub);
        if (EXTRA == 2) xr[u] = *reinterpret_cast<const float4*>(slot + (4 * (4 * half + u) + rsub) * 32 + csub);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int uu = 4 * half + u;
        if (!cok || ro[uu] < 0) continue;
        float4 w = v[u];
        if (AFF == 1) { w.x *= e1.x; w.y *= e1.y; w.z *= e1.z; w.w *= e1.w; }
        if (AFF == 2) {
          const float j = jr[uu];
          w.x = fmaf(w.x, e1.x, j * e2.x); w.y = fmaf(w.y, e1.y, j * e2.y);
          w.z = fmaf(w.z, e1.z, j * e2.z); w.w = fmaf(w.w, e1.w, j * e2.w);
        }
        if (bias) { w.x = __fadd_rn(w.x, bb.x); w.y = __fadd_rn(w.y, bb.y); w.z = __fadd_rn(w.z, bb.z); w.w = __fadd_rn(w.w, bb.w); }
        if (relu) { w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f); }
        if (EXTRA) {   // fused residual add (resnet_model.py:199,314) or dx += (gradient accumulation)
          const float4 x = (EXTRA == 2) ? xr[u] : xa[uu];
          w.x = __fadd_rn(w.x, x.x); w.y = __fadd_rn(w.y, x.y); w.z = __fadd_rn(w.z, x.z); w.w = __fadd_rn(w.w, x.w);
        }
        *reinterpret_cast<float4*>(out + ((size_t)ro[uu] << 2) + n0 + cv) = w;
      }
    }
    __syncwarp();                                // the staging buffer is overwritten by the next chunk
    if (EXTRA == 1) {
#pragma unroll
      for (int u = 0; u < 8; ++u) xa[u] = xb[u];
    }
    if (EXTRA == 2) ring_issue(c0 + c_step * RD, it + RD);   // refill the slot just consumed
  }
  if (EXTRA == 2) asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (c_begin >= BN) {                           // no chunk for this warp (BN < 64): release the accumulator all the same
    tc_fence_before();
    mbar_arrive(tempty);
  }
}

template <int AFF>
__device__ __forceinline__ void epilogue_tile_a(uint32_t t_acc, uint64_t* tfull, uint64_t* tempty, uint32_t parity,
                                                bool zero_tile, long long my_row_off, long long* rowoff, float* stg,
                                                float* __restrict__ out, const float* __restrict__ extra,
                                                const float* __restrict__ bias, int relu, int n0, int BN, int Ng,
                                                int q, int lane, uint8_t* ring, const EpiAff& aff, float my_j,
                                                float* jrow, int c_begin = 0, int c_step = 32,
                                                const float* aff_tab = nullptr, int ring_depth = kRingDepth) {
  if (extra && ring && ring_depth == 2) epilogue_tile_t<2, AFF, 2>(t_acc, tfull, tempty, parity, zero_tile, my_row_off, rowoff, stg, out, extra, bias, relu, n0, BN, Ng, q, lane, ring, aff, my_j, jrow, c_begin, c_step, aff_tab);
  else if (extra && ring) epilogue_tile_t<2, AFF>(t_acc, tfull, tempty, parity, zero_tile, my_row_off, rowoff, stg, out, extra, bias, relu, n0, BN, Ng, q, lane, ring, aff, my_j, jrow, c_begin, c_step, aff_tab);
  else if (extra) epilogue_tile_t<1, AFF>(t_acc, tfull, tempty, parity, zero_tile, my_row_off, rowoff, stg, out, extra, bias, relu, n0, BN, Ng, q, lane, nullptr, aff, my_j, jrow, c_begin, c_step, aff_tab);
  else epilogue_tile_t<0, AFF>(t_acc, tfull, tempty, parity, zero_tile, my_row_off, rowoff, stg, out, nullptr, bias, relu, n0, BN, Ng, q, lane, nullptr, aff, my_j, jrow, c_begin, c_step, aff_tab);
}
__device__ __forceinline__ void epilogue_tile(uint32_t t_acc, uint64_t* tfull, uint64_t* tempty, uint32_t parity,
                                              bool zero_tile, long long my_row_off, long long* rowoff, float* stg,
                                              float* __restrict__ out, const float* __restrict__ extra,
                                              const float* __restrict__ bias, int relu, int n0, int BN, int Ng,
                                              int q, int lane, uint8_t* ring = nullptr) {
  const EpiAff none{};
  epilogue_tile_a<0>(t_acc, tfull, tempty, parity, zero_tile, my_row_off, rowoff, stg, out, extra, bias, relu, n0, BN, Ng,
                     q, lane, ring, none, 0.f, nullptr);
}

// ---- TMA-fed kernels (pf_conv_tma.cu)
void conv_tma_set_feed(int mode);
bool conv_tma_eligible(int pass, const TcGeom& g);      // pass: 0 fwd, 1 dgrad, 2 wgrad
int conv_tma_launch(int pass, const TcGeom& g, const pf_tc_act& a, const pf_tc_wt& w, float* out, int accumulate,
                    const float* bias, int relu, const float* residual, cudaStream_t st, const char* who);
int conv_tma_wgrad_launch(const TcGeom& g, const pf_tc_act& x, const pf_tc_act& dy, int BN, int pps, int splits,
                          float* partial, cudaStream_t st, const char* who);

}  // namespace pfconv
