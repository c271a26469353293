// pf_ws.cu — magnitude-threshold mask build (exact radix select) and order statistics.
//
// Replaces WeightSparseLearner.__build_masks
// (/root/reference/learners/weight_sparsification/learner.py:260-294): per maskable variable
//   bkup = where(mask > 0.5, w, bkup)
//   thr  = tf.contrib.distributions.percentile(|bkup|, 100*s)   -- a FULL top_k sort in the reference
//   mask = float(|bkup| > thr) ; w = bkup*mask
// Here: one multi-tensor launch refreshes bkup and histograms the top byte of |bkup|'s IEEE bit
// pattern, three more multi-tensor histogram passes narrow the 32-bit key of the rank-th element
// exactly (non-negative floats order like their bit patterns), and one multi-tensor launch writes
// mask and w.  No sort, no approximation: masks are bit-exact with the reference's definition.
#include "pf_common.cuh"

namespace {

constexpr int kThreads = 256;

// workspace layout per segment (uint32): [0..255] histogram, [256] prefix, [257] remaining rank
// (ascending, 0-based), [258] mask of the already-fixed high bits, [259..263] spare
constexpr int kWs = PF_WS_WORKSPACE_U32_PER_SEG;

__device__ __forceinline__ void hist_add(uint32_t* sh, uint32_t digit, bool valid) {
  // warp-aggregated shared-memory atomics: weight magnitudes crowd into 2-3 exponent bins
  const uint32_t active = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  const uint32_t peers = __match_any_sync(active, digit);
  if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sh[digit], __popc(peers));
}

// KEYMODE 0: key = bits(|v|) (mask build);  1: key = ordered encoding of v (plain order statistic)
template <int KEYMODE>
__device__ __forceinline__ uint32_t make_key(float v) {
  return KEYMODE == 0 ? (__float_as_uint(v) & 0x7FFFFFFFu) : pf_enc(v);
}

// Pass `shift` in {24,16,8,0}.  FIRST additionally performs bkup = where(mask>0.5, w, bkup).
template <int KEYMODE, bool FIRST>
__global__ void __launch_bounds__(kThreads)
ws_hist_kernel(const pf_ws_seg* __restrict__ segs, const int32_t* __restrict__ qseg,
               const pf_work* __restrict__ work, uint32_t* __restrict__ ws, int shift) {
  __shared__ uint32_t sh[256];
  const pf_work w = work[blockIdx.x];
  const pf_ws_seg s = segs[qseg ? qseg[w.seg] : w.seg];
  uint32_t* st = ws + (size_t)w.seg * kWs;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t prefix = st[256], fixed = st[258];
  const int64_t end = w.start + w.count;
  // warp-uniform trip count: hist_add uses full-warp ballots
  for (int64_t base = w.start; base < end; base += kThreads * 4) {
    const int64_t i = base + (int64_t)threadIdx.x * 4;
    float v[4];
    int nv = 4;
    if (i >= end) {
      nv = 0;
    } else if (i + 3 < end) {
      float4 b = pf_ld4(s.bkup + i);
      if (FIRST) {
        const float4 m = pf_ld_stream(s.mask + i);
        const float4 x = pf_ld4(s.w + i);
        b.x = m.x > 0.5f ? x.x : b.x;
        b.y = m.y > 0.5f ? x.y : b.y;
        b.z = m.z > 0.5f ? x.z : b.z;
        b.w = m.w > 0.5f ? x.w : b.w;
        *reinterpret_cast<float4*>(s.bkup + i) = b;
      }
      v[0] = b.x; v[1] = b.y; v[2] = b.z; v[3] = b.w;
    } else {
      nv = (int)(end - i);
      for (int j = 0; j < nv; ++j) {
        float b = s.bkup[i + j];
        if (FIRST) {
          b = s.mask[i + j] > 0.5f ? s.w[i + j] : b;
          s.bkup[i + j] = b;
        }
        v[j] = b;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t key = j < nv ? make_key<KEYMODE>(v[j]) : 0u;
      const bool valid = j < nv && ((key & fixed) == prefix);
      hist_add(sh, (key >> shift) & 255u, valid);
    }
  }
  __syncthreads();
  const uint32_t c = sh[threadIdx.x];
  if (c) atomicAdd(&st[threadIdx.x], c);
}

// One CTA per segment/query: pick the digit that holds the remaining rank, extend the prefix.
__global__ void __launch_bounds__(256)
ws_scan_kernel(uint32_t* __restrict__ ws, int shift) {
  __shared__ uint32_t sh[256];
  uint32_t* st = ws + (size_t)blockIdx.x * kWs;
  sh[threadIdx.x] = st[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t r = st[257], cum = 0, d = 0;
    for (; d < 256; ++d) {
      const uint32_t c = sh[d];
      if (r < cum + c) break;
      cum += c;
    }
    if (d == 256) d = 255;  // unreachable when rank < numel
    st[256] |= d << shift;
    st[257] = r - cum;
    st[258] |= 255u << shift;
  }
  __syncthreads();
  st[threadIdx.x] = 0;  // histogram ready for the next pass
}

__global__ void ws_init_kernel(const pf_ws_seg* __restrict__ segs, const int32_t* __restrict__ qseg,
                               int n, const int64_t* __restrict__ ranks_desc, uint32_t* __restrict__ ws) {
  const int q = blockIdx.x;
  if (q >= n) return;
  uint32_t* st = ws + (size_t)q * kWs;
  for (int t = threadIdx.x; t < kWs; t += blockDim.x) st[t] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int64_t numel = segs[qseg ? qseg[q] : q].numel;
    int64_t rd = ranks_desc[q];
    rd = rd < 0 ? 0 : (rd > numel - 1 ? numel - 1 : rd);
    st[257] = (uint32_t)(numel - 1 - rd);  // ascending 0-based rank
  }
}

__global__ void __launch_bounds__(kThreads)
ws_apply_kernel(const pf_ws_seg* __restrict__ segs, const pf_work* __restrict__ work,
                const uint32_t* __restrict__ ws, float* __restrict__ thr_out) {
  const pf_work w = work[blockIdx.x];
  const pf_ws_seg s = segs[w.seg];
  const float thr = __uint_as_float(ws[(size_t)w.seg * kWs + 256]);
  if (thr_out && w.start == 0 && threadIdx.x == 0) thr_out[w.seg] = thr;
  const int64_t end = w.start + w.count;
  for (int64_t i = w.start + (int64_t)threadIdx.x * 4; i < end; i += kThreads * 4) {
    if (i + 3 < end) {
      const float4 b = pf_ld_stream(s.bkup + i);
      float4 m, x;
      m.x = fabsf(b.x) > thr ? 1.f : 0.f;
      m.y = fabsf(b.y) > thr ? 1.f : 0.f;
      m.z = fabsf(b.z) > thr ? 1.f : 0.f;
      m.w = fabsf(b.w) > thr ? 1.f : 0.f;
      x.x = __fmul_rn(b.x, m.x);
      x.y = __fmul_rn(b.y, m.y);
      x.z = __fmul_rn(b.z, m.z);
      x.w = __fmul_rn(b.w, m.w);
      pf_st_stream(s.mask + i, m);
      pf_st_stream(s.w + i, x);
    } else {
      for (int64_t j = i; j < end; ++j) {
        const float b = s.bkup[j];
        const float m = fabsf(b) > thr ? 1.f : 0.f;
        s.mask[j] = m;
        s.w[j] = __fmul_rn(b, m);
      }
    }
  }
}

__global__ void select_out_kernel(const uint32_t* __restrict__ ws, int n, float* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) out[q] = pf_dec(ws[(size_t)q * kWs + 256]);
}

}  // namespace

extern "C" {

int pf_ws_mask_build(const pf_ws_seg* segs_dev, int n_seg, const pf_work* work_dev, int n_work,
                     const int64_t* ranks_desc_dev, uint32_t* workspace_dev, float* thr_out_dev,
                     void* stream) {
  PF_REQUIRE(n_seg >= 0 && n_work >= 0, "pf_ws_mask_build: negative count");
  if (n_seg == 0 || n_work == 0) return PF_OK;
  PF_REQUIRE(segs_dev && work_dev && ranks_desc_dev && workspace_dev, "pf_ws_mask_build: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  ws_init_kernel<<<n_seg, 64, 0, st>>>(segs_dev, nullptr, n_seg, ranks_desc_dev, workspace_dev);
  PF_CHECK_LAUNCH("pf_ws_mask_build/init");
  ws_hist_kernel<0, true><<<n_work, kThreads, 0, st>>>(segs_dev, nullptr, work_dev, workspace_dev, 24);
  PF_CHECK_LAUNCH("pf_ws_mask_build/hist0");
  ws_scan_kernel<<<n_seg, 256, 0, st>>>(workspace_dev, 24);
  PF_CHECK_LAUNCH("pf_ws_mask_build/scan0");
  for (int shift = 16; shift >= 0; shift -= 8) {
    ws_hist_kernel<0, false><<<n_work, kThreads, 0, st>>>(segs_dev, nullptr, work_dev, workspace_dev, shift);
    PF_CHECK_LAUNCH("pf_ws_mask_build/hist");
    ws_scan_kernel<<<n_seg, 256, 0, st>>>(workspace_dev, shift);
    PF_CHECK_LAUNCH("pf_ws_mask_build/scan");
  }
  ws_apply_kernel<<<n_work, kThreads, 0, st>>>(segs_dev, work_dev, workspace_dev, thr_out_dev);
  PF_CHECK_LAUNCH("pf_ws_mask_build/apply");
  return PF_OK;
}

int pf_select_desc(const pf_ws_seg* segs_dev, const int32_t* qseg_dev, int n_query,
                   const pf_work* work_dev, int n_work, const int64_t* ranks_desc_dev,
                   uint32_t* workspace_dev, float* out_dev, void* stream) {
  PF_REQUIRE(n_query >= 0 && n_work >= 0, "pf_select_desc: negative count");
  if (n_query == 0 || n_work == 0) return PF_OK;
  PF_REQUIRE(segs_dev && qseg_dev && work_dev && ranks_desc_dev && workspace_dev && out_dev,
             "pf_select_desc: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  ws_init_kernel<<<n_query, 64, 0, st>>>(segs_dev, qseg_dev, n_query, ranks_desc_dev, workspace_dev);
  PF_CHECK_LAUNCH("pf_select_desc/init");
  for (int shift = 24; shift >= 0; shift -= 8) {
    ws_hist_kernel<1, false><<<n_work, kThreads, 0, st>>>(segs_dev, qseg_dev, work_dev, workspace_dev, shift);
    PF_CHECK_LAUNCH("pf_select_desc/hist");
    ws_scan_kernel<<<n_query, 256, 0, st>>>(workspace_dev, shift);
    PF_CHECK_LAUNCH("pf_select_desc/scan");
  }
  select_out_kernel<<<(n_query + 127) / 128, 128, 0, st>>>(workspace_dev, n_query, out_dev);
  PF_CHECK_LAUNCH("pf_select_desc/out");
  return PF_OK;
}

}  // extern "C"
