// pf_cpg.cu — the layer-wise channel-selection phase of the channel-pruning learner.
//
// Reference: /root/reference/learners/channel_pruning_gpu/learner.py
//   :339-354  reg_loss_i = tf.nn.l2_loss(conv_i(full model) - conv_i(pruned model))
//   :375-383  proximal step on the kernel W [R,S,Cin,Cout] of conv_i:
//               W' = W - lr * dreg/dW ; n_c = sqrt(sum_{r,s,k} W'[r,s,c,k]^2) ; t = percentile(n, q)
//               W  = W' * max(1 - t / n_c, 0)                      (group soft-threshold over INPUT channels)
//   :250-260  mask[r,s,c,k] = (n_c > 0)
//   :404-443  masked gradients g * mask for the layer-wise Adam fine-tuning
// All of it is HBM-bound elementwise / small-reduction work on one layer's output (diff + loss) or one kernel
// tensor (norms, shrink, mask): 128-bit loads, one pass each, fixed-order reductions (deterministic).
// The weight gradient itself is the ordinary conv wgrad kernel fed with dY = pruned - full.
#include "pf_common.cuh"

namespace {
constexpr int kThreads = 256;

inline unsigned flat_grid(int64_t nvec) {
  int64_t want = (nvec + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)PF_NUM_SMS * 8;
  if (want < 1) want = 1;
  return (unsigned)(want < cap ? want : cap);
}

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = pf_warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < (kThreads >> 5) ? sh[lane] : 0.f;
    t = pf_warp_sum(t);
  }
  return t;   // valid in warp 0
}

// diff = a - b, partial[block] = sum(diff^2) over the block's grid-stride share (fixed order)
__global__ void __launch_bounds__(kThreads)
diff_l2_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                       float* __restrict__ diff, float* __restrict__ partial) {
  __shared__ float sh[kThreads / 32];
  const int64_t nvec = n >> 2, stride = (int64_t)gridDim.x * kThreads;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += stride) {
    const float4 x = pf_ld_stream(a + (i << 2)), y = pf_ld_stream(b + (i << 2));
    const float4 d = make_float4(__fsub_rn(x.x, y.x), __fsub_rn(x.y, y.y), __fsub_rn(x.z, y.z), __fsub_rn(x.w, y.w));
    pf_st_stream(diff + (i << 2), d);
    acc = fmaf(d.x, d.x, acc);
    acc = fmaf(d.y, d.y, acc);
    acc = fmaf(d.z, d.z, acc);
    acc = fmaf(d.w, d.w, acc);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t j = (nvec << 2) + threadIdx.x;
    const float d = __fsub_rn(a[j], b[j]);
    diff[j] = d;
    acc = fmaf(d, d, acc);
  }
  const float t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kThreads)
l2_final_kernel(const float* __restrict__ partial, int n_partial, float* __restrict__ out) {
  __shared__ float sh[kThreads / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_partial; i += kThreads) acc += partial[i];
  const float t = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = __fmul_rn(t, 0.5f);     // tf.nn.l2_loss = sum(t^2) / 2
}

// one block per input channel c: n_c = sqrt(sum_{rs,k} (w - lr*g)^2); g may be null (norm of w itself)
__global__ void __launch_bounds__(kThreads)
group_norm_kernel(const float* __restrict__ w, const float* __restrict__ g, float lr, int rs, int cin, int cout,
                  float* __restrict__ norms) {
  __shared__ float sh[kThreads / 32];
  const int c = blockIdx.x;
  float acc = 0.f;
  for (int t = 0; t < rs; ++t) {
    const size_t base = ((size_t)t * cin + c) * cout;
    for (int k = threadIdx.x; k < cout; k += kThreads) {
      float v = w[base + k];
      if (g) v = __fsub_rn(v, __fmul_rn(lr, g[base + k]));
      acc = fmaf(v, v, acc);
    }
  }
  const float t = block_sum(acc, sh);
  if (threadIdx.x == 0) norms[c] = __fsqrt_rn(t);
}

// w = (w - lr*g) * max(1 - thr / n_c, 0)
__global__ void __launch_bounds__(kThreads)
prox_apply_kernel(float* __restrict__ w, const float* __restrict__ g, float lr, const float* __restrict__ norms,
                  const float* __restrict__ thr, int cin, int cout, int64_t n) {
  const float t = __ldg(thr);
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const int c = (int)((i / cout) % cin);
    const float v = __fsub_rn(w[i], __fmul_rn(lr, g[i]));
    const float shrk = fmaxf(__fsub_rn(1.f, __fdiv_rn(t, __ldg(norms + c))), 0.f);
    w[i] = __fmul_rn(v, shrk);
  }
}

__global__ void __launch_bounds__(kThreads)
channel_mask_kernel(const float* __restrict__ norms, int cin, int cout, int64_t n, float* __restrict__ mask) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const int c = (int)((i / cout) % cin);
    mask[i] = __ldg(norms + c) > 0.f ? 1.f : 0.f;
  }
}

__global__ void __launch_bounds__(kThreads)
mul_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ out) {
  const int64_t nvec = n >> 2, stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += stride) {
    const float4 x = pf_ld4(a + (i << 2)), y = pf_ld_stream(b + (i << 2));
    pf_st_stream(out + (i << 2), make_float4(__fmul_rn(x.x, y.x), __fmul_rn(x.y, y.y), __fmul_rn(x.z, y.z),
                                              __fmul_rn(x.w, y.w)));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t j = (nvec << 2) + threadIdx.x;
    out[j] = __fmul_rn(a[j], b[j]);
  }
}
}  // namespace

extern "C" {

int pf_cpg_diff_l2(const float* a_dev, const float* b_dev, int64_t n, float* diff_dev, float* loss_dev,
                   float* partial_ws_dev, void* stream) {
  PF_REQUIRE(n >= 0, "pf_cpg_diff_l2: n < 0");
  PF_REQUIRE(a_dev && b_dev && diff_dev && loss_dev && partial_ws_dev, "pf_cpg_diff_l2: null pointer");
  PF_REQUIRE((((uintptr_t)a_dev | (uintptr_t)b_dev | (uintptr_t)diff_dev) & 15) == 0,
             "pf_cpg_diff_l2: buffers must be 16-byte aligned");
  unsigned grid = flat_grid(n >> 2);
  if (grid > PF_L2_PARTIALS) grid = PF_L2_PARTIALS;
  cudaStream_t st = (cudaStream_t)stream;
  diff_l2_partial_kernel<<<grid, kThreads, 0, st>>>(a_dev, b_dev, n, diff_dev, partial_ws_dev);
  PF_CHECK_LAUNCH("pf_cpg_diff_l2");
  l2_final_kernel<<<1, kThreads, 0, st>>>(partial_ws_dev, (int)grid, loss_dev);
  PF_CHECK_LAUNCH("pf_cpg_diff_l2(final)");
  return PF_OK;
}

int pf_cpg_group_norms(const float* w_dev, const float* g_dev, float lr, int rs, int cin, int cout,
                       float* norms_dev, void* stream) {
  PF_REQUIRE(rs >= 1 && cin >= 1 && cout >= 1, "pf_cpg_group_norms: bad shape %d x %d x %d", rs, cin, cout);
  PF_REQUIRE(w_dev && norms_dev, "pf_cpg_group_norms: null pointer");
  group_norm_kernel<<<cin, kThreads, 0, (cudaStream_t)stream>>>(w_dev, g_dev, lr, rs, cin, cout, norms_dev);
  PF_CHECK_LAUNCH("pf_cpg_group_norms");
  return PF_OK;
}

int pf_cpg_prox_apply(float* w_dev, const float* g_dev, float lr, const float* norms_dev, const float* thr_dev,
                      int rs, int cin, int cout, void* stream) {
  PF_REQUIRE(rs >= 1 && cin >= 1 && cout >= 1, "pf_cpg_prox_apply: bad shape %d x %d x %d", rs, cin, cout);
  PF_REQUIRE(w_dev && g_dev && norms_dev && thr_dev, "pf_cpg_prox_apply: null pointer");
  const int64_t n = (int64_t)rs * cin * cout;
  prox_apply_kernel<<<flat_grid(n), kThreads, 0, (cudaStream_t)stream>>>(w_dev, g_dev, lr, norms_dev, thr_dev, cin, cout, n);
  PF_CHECK_LAUNCH("pf_cpg_prox_apply");
  return PF_OK;
}

int pf_cpg_channel_mask(const float* norms_dev, int rs, int cin, int cout, float* mask_dev, void* stream) {
  PF_REQUIRE(rs >= 1 && cin >= 1 && cout >= 1, "pf_cpg_channel_mask: bad shape %d x %d x %d", rs, cin, cout);
  PF_REQUIRE(norms_dev && mask_dev, "pf_cpg_channel_mask: null pointer");
  const int64_t n = (int64_t)rs * cin * cout;
  channel_mask_kernel<<<flat_grid(n), kThreads, 0, (cudaStream_t)stream>>>(norms_dev, cin, cout, n, mask_dev);
  PF_CHECK_LAUNCH("pf_cpg_channel_mask");
  return PF_OK;
}

int pf_mul(const float* a_dev, const float* b_dev, int64_t n, float* out_dev, void* stream) {
  PF_REQUIRE(n >= 0, "pf_mul: n < 0");
  if (n == 0) return PF_OK;
  PF_REQUIRE(a_dev && b_dev && out_dev, "pf_mul: null pointer");
  PF_REQUIRE((((uintptr_t)a_dev | (uintptr_t)b_dev | (uintptr_t)out_dev) & 15) == 0, "pf_mul: buffers must be 16-byte aligned");
  mul_kernel<<<flat_grid(n >> 2), kThreads, 0, (cudaStream_t)stream>>>(a_dev, b_dev, n, out_dev);
  PF_CHECK_LAUNCH("pf_mul");
  return PF_OK;
}

}  // extern "C"
