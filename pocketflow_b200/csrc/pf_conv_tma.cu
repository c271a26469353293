// pf_conv_tma.cu — the tcgen05 convolution kernels fed by the Tensor Memory Accelerator (sm_100a).
//
// Same contraction as pf_conv_tc.cu (tf.nn.conv2d re-created on the quantized weight,
// /root/reference/learners/uniform_quantization/utils.py:92-104, its dgrad and wgrad), for channel counts that are
// multiples of 64.  What changes is how the operands reach shared memory and how many MMAs a product costs:
//   * the NHWC operand (x in fwd / wgrad, dy in dgrad) is fetched by ONE im2col-mode TMA load per k-stage and plane
//     (128 filter-window positions x 64 channels of one tap; the padding is TMA's out-of-bounds zero fill), the
//     K-major weight matrix / the dy matrix by tiled TMA loads — one elected thread issues them, completion is counted
//     in bytes on the stage's mbarrier.  No LSU instruction touches an operand; producer warps are gone
//     (6 warps: TMA, MMA, 4 x epilogue instead of 13), stages are as deep as shared memory allows (up to 8);
//   * an operand that is a <= 8-bit fake-quantized tensor arrives as its INTEGER LEVELS (exact in bf16): one plane
//     instead of hi + lo.  MMAs per k-slice: levels x levels 1, levels x split 2, split x split 3; the per-channel
//     scale, and the rank-1 term that the weight offset contributes, are applied by the epilogue (pf_conv_tc.cuh).
#include <cuda.h>

#include "pf_conv_tc.cuh"
#include "pf_tma.cuh"

namespace pfconv {
using namespace pftma;

constexpr int kTmaMaxStages = 8;
constexpr int kTmaThreads = (2 + kTmaEpiWarps) * 32;     // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr uint32_t kATileBytes = TM * 128;

struct TmaP {
  int M, Ng, BN, nk, n_tiles, total_tiles, acc_cols;
  int cblocks, R, S;                       // k-stage ks -> tap = ks / cblocks (r = tap / S, q = tap % S), channel block
  int rows_hw, rows_w;                     // GEMM row m -> (image, y, x)
  int src_h, src_w;                        // gathered tensor
  int base_w, base_h, str_w, str_h, flip;  // window origin of row (y, x): (base + x * str); flip: tap offsets mirrored
  int na, nb;                              // operand planes (na: upper bound when a_hdr decides)
  int accumulate, relu, ring, stage_budget, epi_warps;     // ring: 0 = none, else the residual ring's depth (2 or 4)
  FastDiv d_hw, d_w, d_ntiles, d_cblocks, d_s;
  EpiAff aff;
  const pf_tc_act_hdr* a_hdr;
  const float* csum;
  int nseg;
};

// ---------------------------------------------------------------------------------------------------------
// fwd / dgrad (unit stride): D[M x Ng] = A[M x K] * B[Ng x K]^T, both operands K-major, 128 x BN output tiles,
// persistent CTAs, double-buffered TMEM accumulator (the epilogue of tile i overlaps the main loop of tile i+1).
template <int AFF>
__global__ void __launch_bounds__(kTmaThreads, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                float* __restrict__ out, const float* __restrict__ bias, const float* __restrict__ residual,
                const __grid_constant__ TmaP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[kTmaMaxStages], empty_bar[kTmaMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN;
  int na = p.na;
  if (p.a_hdr) na = (__ldg(&p.a_hdr->nplanes) == 2) ? 2 : 1;
  const int nb = p.nb;
  const uint32_t b_bytes = (uint32_t)BN * 128u;
  const uint32_t stage_bytes = (uint32_t)na * kATileBytes + (uint32_t)nb * b_bytes;
  const uint32_t n_stages = min((uint32_t)kTmaMaxStages, (uint32_t)p.stage_budget / stage_bytes);
  uint8_t* epi = smem + p.stage_budget;
  float* stage_all = reinterpret_cast<float*>(epi);
  long long* rowoff_all = reinterpret_cast<long long*>(stage_all + p.epi_warps * 32 * kStagePitch);
  float* jrow_all = reinterpret_cast<float*>(rowoff_all + p.epi_warps * 32);
  float* aff_tab = jrow_all + p.epi_warps * 32;               // AFF == 2: e1[256], e2[256] of the current tile columns
  uint8_t* ring_all = reinterpret_cast<uint8_t*>(aff_tab + (AFF == 2 ? 2 * 256 : 0));

  if (tid == 0) {
    for (int s = 0; s < kTmaMaxStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], p.epi_warps * 32);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)(2 * p.acc_cols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int first_tile = blockIdx.x, tile_step = gridDim.x;

  if (warp == 0) {
    // =================================== TMA producer (one thread) ===================================
    if (lane == 0) {
      prefetch_map(&tmA0);
      prefetch_map(&tmB0);
      if (na == 2) prefetch_map(&tmA1);
      if (nb == 2) prefetch_map(&tmB1);
      uint32_t s = 0, ph = 0;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_step) {
        const int mt = (int)fdiv((uint32_t)tile, p.d_ntiles);
        const int n0 = (tile - mt * p.n_tiles) * BN;
        const int m0 = mt * TM;
        const int img = (int)fdiv((uint32_t)m0, p.d_hw);
        const int rem = m0 - img * p.rows_hw;
        const int y = (int)fdiv((uint32_t)rem, p.d_w), x = rem - y * p.rows_w;
        const int cw = p.base_w + x * p.str_w, ch = p.base_h + y * p.str_h;
        for (int ks = 0; ks < p.nk; ++ks) {
          mbar_wait_bounded(&empty_bar[s], ph ^ 1u);                     // slot free?
          const int tap = (int)fdiv((uint32_t)ks, p.d_cblocks);
          const int cb = ks - tap * p.cblocks;
          const int r = (int)fdiv((uint32_t)tap, p.d_s), q = tap - r * p.S;
          const uint32_t ow = (uint32_t)(p.flip ? p.S - 1 - q : q), oh = (uint32_t)(p.flip ? p.R - 1 - r : r);
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t sb = sa + (uint32_t)na * kATileBytes;
          mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
          load_im2col(sa, &tmA0, &full_bar[s], cb * BK, cw, ch, img, ow, oh);
          if (na == 2) load_im2col(sa + kATileBytes, &tmA1, &full_bar[s], cb * BK, cw, ch, img, ow, oh);
          load_2d(sb, &tmB0, &full_bar[s], ks * BK, n0);
          if (nb == 2) load_2d(sb + b_bytes, &tmB1, &full_bar[s], ks * BK, n0);
          if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =================================== MMA issuer (one thread) ===================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(TM, BN, 0, 0);
      uint32_t s = 0, ph = 0, tcount = 0;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++tcount) {
        const uint32_t buf = tcount & 1u;
        mbar_wait_bounded(&tempty_bar[buf], ((tcount >> 1) & 1u) ^ 1u);    // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)p.acc_cols;
        for (int ks = 0; ks < p.nk; ++ks) {
          mbar_wait_bounded(&full_bar[s], ph);                             // TMA bytes have landed
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + (size_t)s * stage_bytes), a1 = a0 + kATileBytes;
          const uint32_t b0 = a0 + (uint32_t)na * kATileBytes, b1 = b0 + b_bytes;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t da0 = make_smem_desc(a0 + kk * 32, 16, 1024);
            const uint64_t db0 = make_smem_desc(b0 + kk * 32, 16, 1024);
            umma_bf16(d_tmem, da0, db0, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
            if (nb == 2) umma_bf16(d_tmem, da0, make_smem_desc(b1 + kk * 32, 16, 1024), idesc, 1u);
            if (na == 2) umma_bf16(d_tmem, make_smem_desc(a1 + kk * 32, 16, 1024), db0, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);        // frees the stage when these MMAs have completed
          if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
        umma_commit(&tfull_bar[buf]);        // accumulator of this tile complete
      }
    }
  } else if (warp < 2 + p.epi_warps) {
    // =================================== epilogue (warps 2-9, or 2-5 when shared memory is short) =================
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int ew = warp - 2, half = ew >> 2; // two warps per quarter: even / odd 32-column chunks
    float* stg = stage_all + (size_t)ew * 32 * kStagePitch;
    long long* rowoff = rowoff_all + ew * 32;
    float* jrow = jrow_all + ew * 32;
    const float* extra = residual ? residual : (p.accumulate ? out : nullptr);
    uint32_t tcount = 0;
    int tab_n0 = -1;
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++tcount) {
      const int mt = (int)fdiv((uint32_t)tile, p.d_ntiles);
      const int n0 = (tile - mt * p.n_tiles) * BN;
      const int m = mt * TM + q * 32 + lane;
      const long long off = m < p.M ? (long long)m * p.Ng : -1;
      if (AFF == 2 && n0 != tab_n0) {
        // per-column constants of this tile's columns, once (every epilogue warp walks the same tile sequence, so the
        // named barrier below is reached by all of them; with one n-tile per row of tiles this runs once per CTA):
        //   e1[c] = s_a * alpha_c / k_w ,  e2[c] = s_a * (centre * alpha_c / k_w + beta_c)
        const int nthr = p.epi_warps * 32;
        asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");          // readers of the previous table are done
        const float a_s = p.aff.a_scale ? __ldg(p.aff.a_scale) : 1.f;
        for (int c = ew * 32 + lane; c < BN; c += nthr) {
          const bool cok = n0 + c < p.Ng;
          const int bi = p.aff.per_channel ? n0 + c : 0;
          const float al = cok ? __ldg(p.aff.w_alpha + bi) : 0.f, be = cok ? __ldg(p.aff.w_beta + bi) : 0.f;
          const float sx = al * p.aff.w_rk;
          aff_tab[c] = sx * a_s;
          aff_tab[256 + c] = fmaf(p.aff.w_centre, sx, be) * a_s;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
        tab_n0 = n0;
      }
      float my_j = 0.f;
      if (AFF == 2 && p.csum && m < p.M) {
        // sum of the stored activation levels under this row's filter window, from the per-pixel channel sums: the
        // (tap, segment) terms are independent loads, issued four at a time (branch-free; the terms are integers below
        // 2^24, so the order of the additions does not change the result)
        const int img = (int)fdiv((uint32_t)m, p.d_hw);
        const int rem = m - img * p.rows_hw;
        const int y = (int)fdiv((uint32_t)rem, p.d_w), x = rem - y * p.rows_w;
        const int ow0 = p.base_w + x * p.str_w, oh0 = p.base_h + y * p.str_h;
        const int total = p.R * p.S * p.nseg;
        const float* cbase = p.csum + (size_t)img * p.src_h * p.src_w * p.nseg;
        auto term = [&](int u) -> float {
          if (u >= total) return 0.f;
          const int t = u / p.nseg, g = u - t * p.nseg;
          const int r = (int)fdiv((uint32_t)t, p.d_s), s_ = t - r * p.S;
          const int ih = oh0 + r, iw = ow0 + s_;
          const bool ok = (unsigned)ih < (unsigned)p.src_h && (unsigned)iw < (unsigned)p.src_w;
          return ok ? __ldg(cbase + ((size_t)ih * p.src_w + iw) * p.nseg + g) : 0.f;
        };
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int u = 0; u < total; u += 4) {
          const float v0 = term(u), v1 = term(u + 1), v2 = term(u + 2), v3 = term(u + 3);
          a0 += v0; a1 += v1; a2 += v2; a3 += v3;
        }
        my_j = (a0 + a1) + (a2 + a3);
      }
      epilogue_tile_a<AFF>(tmem_base + (tcount & 1u) * (uint32_t)p.acc_cols, &tfull_bar[tcount & 1u],
                           &tempty_bar[tcount & 1u], (tcount >> 1) & 1u, false, off, rowoff, stg, out, extra, bias, p.relu,
                           n0, BN, p.Ng, q, lane, p.ring ? ring_all + (size_t)ew * p.ring * kRingSlotBytes : nullptr,
                           p.aff, my_j, jrow, 32 * half, 8 * p.epi_warps, AFF == 2 ? aff_tab : nullptr, p.ring);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)(2 * p.acc_cols));
}

// ---------------------------------------------------------------------------------------------------------
// wgrad: dW[kf x cout] = X[pixels x kf]^T * dY[pixels x cout], both operands MN-major in shared memory
// ([64-wide MN block][pixel][128 B], exactly what a 64-channel x 64-pixel TMA box with SWIZZLE_128B writes).
// Work unit = (kf tile of 128 = two 64-channel blocks, cout tile of BN, pixel range); split-K partials as in
// pf_conv_tc.cu.  x arrives through an im2col-mode map (64 window positions x 64 channels of the tap the block
// belongs to), dy through a tiled map.
struct WgTmaP {
  TcGeom g;
  int Mtot, Npix, pps, splits, BN, n_tiles, tiles, total_units, acc_cols;
  int na, nb, stage_budget, epi_warps;
  FastDiv d_pq, d_q, d_c, d_s, d_tiles, d_ntiles;
  EpiAff aff;
  const pf_tc_act_hdr* x_hdr;
};
constexpr uint32_t kWgBlockBytes = BK * 128;   // one 64 (MN) x 64 (pixels) block

template <int AFF>
__global__ void __launch_bounds__(kTmaThreads, 1)
conv_tma_wgrad_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                      const __grid_constant__ CUtensorMap tmY0, const __grid_constant__ CUtensorMap tmY1,
                      float* __restrict__ partial, const __grid_constant__ WgTmaP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[kTmaMaxStages], empty_bar[kTmaMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TcGeom& g = p.g;
  const int BN = p.BN, nblkB = BN / 64;
  int na = p.na;
  if (p.x_hdr) na = (__ldg(&p.x_hdr->nplanes) == 2) ? 2 : 1;
  const int nb = p.nb;
  const uint32_t a_bytes = 2 * kWgBlockBytes, b_bytes = (uint32_t)nblkB * kWgBlockBytes;
  const uint32_t stage_bytes = (uint32_t)na * a_bytes + (uint32_t)nb * b_bytes;
  const uint32_t n_stages = min((uint32_t)kTmaMaxStages, (uint32_t)p.stage_budget / stage_bytes);
  uint8_t* epi = smem + p.stage_budget;
  float* stage_all = reinterpret_cast<float*>(epi);
  long long* rowoff_all = reinterpret_cast<long long*>(stage_all + p.epi_warps * 32 * kStagePitch);
  float* jrow_all = reinterpret_cast<float*>(rowoff_all + p.epi_warps * 32);
  if (tid == 0) {
    for (int s = 0; s < kTmaMaxStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], p.epi_warps * 32);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)(2 * p.acc_cols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  struct Unit {
    int split, m0, n0, pbeg, nk;
  };
  auto decode = [&](int u) -> Unit {
    Unit r;
    r.split = (int)fdiv((uint32_t)u, p.d_tiles);
    const int t = u - r.split * p.tiles;
    const int mt = (int)fdiv((uint32_t)t, p.d_ntiles);
    r.m0 = mt * TM;
    r.n0 = (t - mt * p.n_tiles) * BN;
    r.pbeg = r.split * p.pps;
    const int pend = min(p.Npix, r.pbeg + p.pps);
    r.nk = (pend - r.pbeg + BK - 1) / BK;
    return r;
  };

  if (warp == 0) {
    if (lane == 0) {
      prefetch_map(&tmX0);
      prefetch_map(&tmY0);
      if (na == 2) prefetch_map(&tmX1);
      if (nb == 2) prefetch_map(&tmY1);
      const int pq = g.P * g.Q;
      uint32_t s = 0, ph = 0;
      for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
        const Unit un = decode(u);
        // the two 64-row blocks of the kf tile: each lies inside one filter tap (Cin % 64 == 0)
        int c0[2], tr[2], tq[2], nvalid = 0;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int kf = un.m0 + 64 * b;
          const int tap = (int)fdiv((uint32_t)kf, p.d_c);
          c0[b] = kf - tap * g.C;
          tr[b] = (int)fdiv((uint32_t)tap, p.d_s);
          tq[b] = tap - tr[b] * g.S;
          if (kf < p.Mtot) nvalid = b + 1;
        }
        const uint32_t tx = (uint32_t)na * (uint32_t)nvalid * kWgBlockBytes + (uint32_t)nb * b_bytes;
        for (int ks = 0; ks < un.nk; ++ks) {
          mbar_wait_bounded(&empty_bar[s], ph ^ 1u);
          const int pix0 = un.pbeg + ks * BK;
          const int pn = (int)fdiv((uint32_t)pix0, p.d_pq);
          const int rem = pix0 - pn * pq;
          const int oh = (int)fdiv((uint32_t)rem, p.d_q), ow = rem - oh * g.Q;
          const int cw = ow * g.sw - g.pl, ch = oh * g.sh - g.pt;
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t sb = sa + (uint32_t)na * a_bytes;
          mbar_arrive_expect_tx(&full_bar[s], tx);
          for (int b = 0; b < nvalid; ++b) {
            load_im2col(sa + b * kWgBlockBytes, &tmX0, &full_bar[s], c0[b], cw, ch, pn, (uint32_t)tq[b], (uint32_t)tr[b]);
            if (na == 2)
              load_im2col(sa + a_bytes + b * kWgBlockBytes, &tmX1, &full_bar[s], c0[b], cw, ch, pn, (uint32_t)tq[b],
                          (uint32_t)tr[b]);
          }
          for (int j = 0; j < nblkB; ++j) {
            load_2d(sb + j * kWgBlockBytes, &tmY0, &full_bar[s], un.n0 + 64 * j, pix0);
            if (nb == 2) load_2d(sb + b_bytes + j * kWgBlockBytes, &tmY1, &full_bar[s], un.n0 + 64 * j, pix0);
          }
          if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(TM, BN, 1, 1);
      uint32_t s = 0, ph = 0, tcount = 0;
      for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++tcount) {
        const Unit un = decode(u);
        const uint32_t buf = tcount & 1u;
        mbar_wait_bounded(&tempty_bar[buf], ((tcount >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)p.acc_cols;
        for (int ks = 0; ks < un.nk; ++ks) {
          mbar_wait_bounded(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + (size_t)s * stage_bytes), a1 = a0 + a_bytes;
          const uint32_t b0 = a0 + (uint32_t)na * a_bytes, b1 = b0 + b_bytes;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            // MN-major: LBO = stride between 64-wide MN blocks (8 KB), SBO = stride between 8-pixel groups (1 KB);
            // one MMA consumes 16 pixels = 2 KB
            const uint64_t da0 = make_smem_desc(a0 + kk * 2048, kWgBlockBytes, 1024);
            const uint64_t db0 = make_smem_desc(b0 + kk * 2048, kWgBlockBytes, 1024);
            umma_bf16(d_tmem, da0, db0, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
            if (nb == 2) umma_bf16(d_tmem, da0, make_smem_desc(b1 + kk * 2048, kWgBlockBytes, 1024), idesc, 1u);
            if (na == 2) umma_bf16(d_tmem, make_smem_desc(a1 + kk * 2048, kWgBlockBytes, 1024), db0, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
          if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
        if (un.nk > 0) umma_commit(&tfull_bar[buf]);
        else mbar_arrive(&tfull_bar[buf]);
      }
    }
  } else if (warp < 2 + p.epi_warps) {
    const int q = warp & 3;
    const int ew = warp - 2, half = ew >> 2;
    float* stg = stage_all + (size_t)ew * 32 * kStagePitch;
    long long* rowoff = rowoff_all + ew * 32;
    float* jrow = jrow_all + ew * 32;
    uint32_t tcount = 0;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++tcount) {
      const Unit un = decode(u);
      const int em = un.m0 + q * 32 + lane;
      const long long off = em < p.Mtot ? ((long long)un.split * p.Mtot + em) * g.K : -1;
      epilogue_tile_a<AFF>(tmem_base + (tcount & 1u) * (uint32_t)p.acc_cols, &tfull_bar[tcount & 1u],
                           &tempty_bar[tcount & 1u], (tcount >> 1) & 1u, un.nk == 0, off, rowoff, stg, partial, nullptr,
                           nullptr, 0, un.n0, BN, g.K, q, lane, nullptr, p.aff, 0.f, jrow, 32 * half, 8 * p.epi_warps);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)(2 * p.acc_cols));
}

// ---------------------------------------------------------------------------------------------------------
// host side
static int g_feed_override = -1;          // pf_conv2d_tc_set_feed: 0 = cp.async kernels, 1 = TMA kernels, -1 = PF_TC_FEED / default
void conv_tma_set_feed(int mode) { g_feed_override = mode < 0 ? -1 : (mode ? 1 : 0); }
static bool tma_enabled() {
  if (g_feed_override >= 0) return g_feed_override == 1;
  static int on = -1;
  if (on < 0) {
    const char* v = getenv("PF_TC_FEED");
    on = !(v && strcmp(v, "lsu") == 0);
  }
  return on == 1;
}

bool conv_tma_eligible(int pass, const TcGeom& g) {
  if (!tma_enabled()) return false;
  if (pass == 0) return g.C % 64 == 0 && g.K % 16 == 0 && g.sh <= 8 && g.sw <= 8 && g.R <= 16 && g.S <= 16;
  if (pass == 1) return g.K % 64 == 0 && g.C % 16 == 0 && g.sh == 1 && g.sw == 1 && g.R <= 16 && g.S <= 16;
  return g.C % 64 == 0 && g.K % 64 == 0 && g.sh <= 8 && g.sw <= 8;
}

static int pick_bn(int Ng, int m_tiles, bool has_extra, bool split_planes, int nk) {
  int BN;
  if (Ng >= 256 && has_extra && (split_planes || nk >= 2)) {
    // an epilogue that also streams a residual / accumulate operand: 128-wide tiles leave shared memory for the ring
    // (and for 8 epilogue warps) that 256-wide stages take away.  Measured on the conv3 + shortcut layers of ResNet-50
    // (batch 256, ms at BN 256 -> 128): split planes 64->256 0.516 -> 0.319, 128->512 0.267 -> 0.265; levels 128->512
    // 0.299 -> 0.223, 256->1024 0.176 -> 0.133 — but levels 64->256 (one k-stage per tile) 0.354 -> 0.407: stays 256.
    BN = 128;
  } else if (Ng >= 256) {
    const int64_t t256 = (int64_t)m_tiles * ((Ng + 255) / 256), t128 = (int64_t)m_tiles * ((Ng + 127) / 128);
    const double c256 = (double)((t256 + PF_NUM_SMS - 1) / PF_NUM_SMS) * 1.3;   // a 256-wide tile costs ~1.3x a 128-wide one
    const double c128 = (double)((t128 + PF_NUM_SMS - 1) / PF_NUM_SMS);
    BN = c256 <= c128 ? 256 : 128;
  } else {
    BN = Ng >= 128 ? 128 : (Ng >= 64 ? 64 : (Ng >= 32 ? 32 : 16));
  }
  const int forced = env_int("PF_TC_BN", 0);
  if (forced >= 16 && forced <= 256 && forced <= ((Ng + 15) / 16) * 16 && (forced & (forced - 1)) == 0) BN = forced;
  return BN;
}

#define PF_TMA_ENCODE(call, who)                                                               \
  do {                                                                                         \
    const int e__ = (call);                                                                    \
    if (e__ != 0) {                                                                            \
      pf_set_error("%s: tensor-map encoding failed (%d): %s", who, e__, #call);                \
      return PF_ERR_INVALID_ARG;                                                               \
    }                                                                                          \
  } while (0)

// pass 0: fwd (a = x planes, b = [Cout][Kpad] weights); pass 1: unit-stride dgrad (a = dy planes, b = [Cin][Kpad_d])
int conv_tma_launch(int pass, const TcGeom& g, const pf_tc_act& a, const pf_tc_wt& w, float* out, int accumulate,
                    const float* bias, int relu, const float* residual, cudaStream_t st, const char* who) {
  TmaP p;
  memset(&p, 0, sizeof(p));
  const int CC = pass == 0 ? g.C : g.K;
  const int Hs = pass == 0 ? g.H : g.P, Ws = pass == 0 ? g.W : g.Q;      // gathered tensor
  const int Ho = pass == 0 ? g.P : g.H, Wo = pass == 0 ? g.Q : g.W;      // GEMM rows
  const int64_t M64 = (int64_t)g.N * Ho * Wo;
  PF_REQUIRE(M64 < (1ll << 31), "%s: too many rows", who);
  p.M = (int)M64;
  p.Ng = pass == 0 ? g.K : g.C;
  const int Kdim = g.R * g.S * CC;
  p.nk = Kdim / BK;
  p.cblocks = CC / BK;
  p.R = g.R;
  p.S = g.S;
  p.rows_hw = Ho * Wo;
  p.rows_w = Wo;
  p.src_h = Hs;
  p.src_w = Ws;
  if (pass == 0) {
    p.base_w = -g.pl; p.base_h = -g.pt; p.str_w = g.sw; p.str_h = g.sh; p.flip = 0;
  } else {
    p.base_w = g.pl - (g.S - 1); p.base_h = g.pt - (g.R - 1); p.str_w = 1; p.str_h = 1; p.flip = 1;
  }
  p.accumulate = accumulate;
  p.relu = relu;
  const int m_tiles = (p.M + TM - 1) / TM;
  const int BN = pick_bn(p.Ng, m_tiles, residual != nullptr || accumulate, a.plane1 != nullptr && a.hdr == nullptr, p.nk);
  p.BN = BN;
  p.n_tiles = (p.Ng + BN - 1) / BN;
  p.total_tiles = m_tiles * p.n_tiles;
  p.acc_cols = 32;
  while (p.acc_cols < BN) p.acc_cols <<= 1;
  p.d_hw = make_fastdiv((uint32_t)p.rows_hw);
  p.d_w = make_fastdiv((uint32_t)p.rows_w);
  p.d_ntiles = make_fastdiv((uint32_t)p.n_tiles);
  p.d_cblocks = make_fastdiv((uint32_t)p.cblocks);
  p.d_s = make_fastdiv((uint32_t)g.S);
  p.na = a.plane1 ? 2 : 1;
  p.nb = w.plane1 ? 2 : 1;
  p.a_hdr = a.hdr;
  PF_REQUIRE(a.hdr == nullptr || a.plane1 != nullptr, "%s: an operand with a device header needs both planes", who);
  int aff = 0;
  if (w.alpha) {
    PF_REQUIRE(w.beta != nullptr && w.bits >= 1 && w.bits <= 8, "%s: weight levels need alpha, beta and 1..8 bits", who);
    PF_REQUIRE(a.csum != nullptr && a.nseg >= 1, "%s: weight levels need the operand's channel sums", who);
    aff = 2;
    p.aff.w_alpha = w.alpha;
    p.aff.w_beta = w.beta;
    p.aff.per_channel = w.per_channel;
    p.aff.w_rk = 1.f / (float)((1 << w.bits) - 1);
    p.aff.w_centre = (float)(1 << (w.bits - 1));
    p.aff.a_scale = a.hdr ? &a.hdr->scale : nullptr;
    p.csum = a.csum;
    p.nseg = a.nseg;
  } else if (a.hdr) {
    aff = 1;
    p.aff.a_scale = &a.hdr->scale;
  }
  // ---- shared memory: [stages][epilogue staging, row offsets, J][residual ring]; 8 epilogue warps when at least two
  // (three with a residual ring) stages still fit beside their staging tiles, else 4
  const int stage_max = p.na * (int)kATileBytes + p.nb * BN * 128;
  const bool has_extra = residual != nullptr || accumulate;
  const int aff_tab_bytes = aff == 2 ? 2 * 256 * 4 : 0;       // the tile's per-column epilogue constants
  auto epi_bytes = [aff_tab_bytes](int warps) {
    return 1024 + warps * (32 * kStagePitch * 4 + 32 * 8 + 32 * 4) + aff_tab_bytes + 256;
  };
  // 8 epilogue warps unless their staging tiles cost a pipeline stage that 4 warps would leave (below 4 stages)
  const int st8 = (kSmemLimit - epi_bytes(kTmaEpiWarps)) / 1024 * 1024 / stage_max, st4 = (kSmemLimit - epi_bytes(4)) / 1024 * 1024 / stage_max;
  p.epi_warps = (BN >= 64 && st8 >= 2 && (st8 >= 4 || st8 == st4)) ? kTmaEpiWarps : 4;
  p.epi_warps = env_int("PF_TC_EPI_WARPS", p.epi_warps) == 4 ? 4 : p.epi_warps;
  // the residual / accumulate operand streams through a per-warp cp.async ring of 4 (else 2) 4 KB chunks when the
  // pipeline keeps enough stages beside it: 3, or nk + 1 for the short reductions of the 1x1 layers (a 64 -> 256 layer
  // has ONE k-stage per tile: two stages already let the next tile's loads fly during this tile's MMAs)
  int budget = (kSmemLimit - epi_bytes(p.epi_warps)) / 1024 * 1024;
  int ring_bytes = 0;
  p.ring = 0;
  if (has_extra && env_int("PF_TC_RING", 1) && BN >= 64) {
    const int need = std::min(3, p.nk + 1);
    for (int depth = kRingDepth; depth >= 2 && !p.ring; depth >>= 1) {
      const int rb = p.epi_warps * depth * kRingSlotBytes;
      if ((budget - rb) / stage_max >= need) {
        p.ring = depth;
        ring_bytes = rb;
      }
    }
    if (p.ring) budget = (kSmemLimit - epi_bytes(p.epi_warps) - ring_bytes) / 1024 * 1024;
  }
  PF_REQUIRE(budget / stage_max >= 2 || p.nk <= 1, "%s: shared-memory plan failed (BN %d)", who, BN);
  p.stage_budget = budget;
  const size_t smem = 1024 + (size_t)budget + (epi_bytes(p.epi_warps) - 1024) + (p.ring ? ring_bytes : 0);
  if (p.total_tiles == 0) return PF_OK;
  // ---- tensor maps
  alignas(64) CUtensorMap tA0, tA1, tB0, tB1;
  PF_TMA_ENCODE(encode_im2col_bf16(&tA0, a.plane0, g.N, Hs, Ws, CC, p.base_w, p.base_h, Wo, Ho, p.str_w, p.str_h, BK, TM), who);
  if (a.plane1)
    PF_TMA_ENCODE(encode_im2col_bf16(&tA1, a.plane1, g.N, Hs, Ws, CC, p.base_w, p.base_h, Wo, Ho, p.str_w, p.str_h, BK, TM), who);
  else
    tA1 = tA0;
  const int Kpad = pad64(Kdim);
  PF_TMA_ENCODE(encode_2d_bf16(&tB0, w.plane0, (uint64_t)Kpad, (uint64_t)p.Ng, (uint64_t)Kpad, BK, (uint32_t)BN), who);
  if (w.plane1)
    PF_TMA_ENCODE(encode_2d_bf16(&tB1, w.plane1, (uint64_t)Kpad, (uint64_t)p.Ng, (uint64_t)Kpad, BK, (uint32_t)BN), who);
  else
    tB1 = tB0;
  const int grid = std::min(p.total_tiles, PF_NUM_SMS);
#define PF_TMA_LAUNCH(AFFV)                                                                                         \
  do {                                                                                                              \
    auto kern = conv_tma_kernel<AFFV>;                                                                              \
    PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
    kern<<<grid, kTmaThreads, smem, st>>>(tA0, tA1, tB0, tB1, out, bias, residual, p);                              \
  } while (0)
  if (aff == 2) PF_TMA_LAUNCH(2);
  else if (aff == 1) PF_TMA_LAUNCH(1);
  else PF_TMA_LAUNCH(0);
#undef PF_TMA_LAUNCH
  PF_CHECK_LAUNCH(who);
  return PF_OK;
}

int conv_tma_wgrad_launch(const TcGeom& g, const pf_tc_act& x, const pf_tc_act& dy, int BN, int pps, int splits,
                          float* partial, cudaStream_t st, const char* who) {
  WgTmaP p;
  memset(&p, 0, sizeof(p));
  p.g = g;
  p.Mtot = g.R * g.S * g.C;
  p.Npix = g.N * g.P * g.Q;
  p.BN = BN;
  p.pps = pps;
  p.splits = splits;
  const int m_tiles = (p.Mtot + TM - 1) / TM;
  p.n_tiles = (g.K + BN - 1) / BN;
  p.tiles = m_tiles * p.n_tiles;
  p.total_units = p.tiles * p.splits;
  p.acc_cols = 32;
  while (p.acc_cols < BN) p.acc_cols <<= 1;
  p.d_pq = make_fastdiv((uint32_t)(g.P * g.Q));
  p.d_q = make_fastdiv((uint32_t)g.Q);
  p.d_c = make_fastdiv((uint32_t)g.C);
  p.d_s = make_fastdiv((uint32_t)g.S);
  p.d_tiles = make_fastdiv((uint32_t)p.tiles);
  p.d_ntiles = make_fastdiv((uint32_t)p.n_tiles);
  p.na = x.plane1 ? 2 : 1;
  p.nb = dy.plane1 ? 2 : 1;
  p.x_hdr = x.hdr;
  PF_REQUIRE(x.hdr == nullptr || x.plane1 != nullptr, "%s: an operand with a device header needs both planes", who);
  PF_REQUIRE(dy.hdr == nullptr, "%s: the gradient operand is always split-bf16", who);
  int aff = 0;
  if (x.hdr) {
    aff = 1;
    p.aff.a_scale = &x.hdr->scale;
  }
  const int stage_max = p.na * 2 * (int)kWgBlockBytes + p.nb * (BN / 64) * (int)kWgBlockBytes;
  auto epi_bytes = [](int warps) { return 1024 + warps * (32 * kStagePitch * 4 + 32 * 8 + 32 * 4) + 256; };
  const int st8 = (kSmemLimit - epi_bytes(kTmaEpiWarps)) / 1024 * 1024 / stage_max, st4 = (kSmemLimit - epi_bytes(4)) / 1024 * 1024 / stage_max;
  p.epi_warps = (st8 >= 2 && (st8 >= 4 || st8 == st4)) ? kTmaEpiWarps : 4;
  const int budget = (kSmemLimit - epi_bytes(p.epi_warps)) / 1024 * 1024;
  PF_REQUIRE(budget / stage_max >= 2, "%s: shared-memory plan failed (BN %d)", who, BN);
  p.stage_budget = budget;
  const size_t smem = 1024 + (size_t)budget + (epi_bytes(p.epi_warps) - 1024);
  if (p.total_units == 0) return PF_OK;
  alignas(64) CUtensorMap tX0, tX1, tY0, tY1;
  PF_TMA_ENCODE(encode_im2col_bf16(&tX0, x.plane0, g.N, g.H, g.W, g.C, -g.pl, -g.pt, g.Q, g.P, g.sw, g.sh, BK, BK), who);
  if (x.plane1)
    PF_TMA_ENCODE(encode_im2col_bf16(&tX1, x.plane1, g.N, g.H, g.W, g.C, -g.pl, -g.pt, g.Q, g.P, g.sw, g.sh, BK, BK), who);
  else
    tX1 = tX0;
  PF_TMA_ENCODE(encode_2d_bf16(&tY0, dy.plane0, (uint64_t)g.K, (uint64_t)p.Npix, (uint64_t)g.K, BK, BK), who);
  if (dy.plane1)
    PF_TMA_ENCODE(encode_2d_bf16(&tY1, dy.plane1, (uint64_t)g.K, (uint64_t)p.Npix, (uint64_t)g.K, BK, BK), who);
  else
    tY1 = tY0;
  const int grid = std::min(p.total_units, PF_NUM_SMS);
  if (aff == 1) {
    auto kern = conv_tma_wgrad_kernel<1>;
    PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kTmaThreads, smem, st>>>(tX0, tX1, tY0, tY1, partial, p);
  } else {
    auto kern = conv_tma_wgrad_kernel<0>;
    PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kTmaThreads, smem, st>>>(tX0, tX1, tY0, tY1, partial, p);
  }
  PF_CHECK_LAUNCH(who);
  return PF_OK;
}

}  // namespace pfconv
