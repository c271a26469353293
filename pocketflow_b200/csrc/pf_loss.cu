// pf_loss.cu — softmax cross-entropy (hard + temperature-scaled soft labels) and l2_loss.
//
// pf_softmax_ce_fwd_bwd replaces, in one pass over the N x K logits, the 5-8 small TF kernels of
//   tf.losses.softmax_cross_entropy(labels, logits)        (/root/reference/nets/resnet_at_cifar10.py:104)
//   DistillationHelper.calc_loss: softmax(t/T), s/T, softmax_cross_entropy, * loss_w_dst
//                                                         (/root/reference/learners/distillation_helper.py:98-100)
// and their autodiff, plus the accuracy metrics (resnet_at_cifar10.py:108-110).
// Row arithmetic follows TF's xent functor: z = x - max; lse = log(sum exp z);
// loss = sum(labels*(lse - z)); backprop = exp(z)/sum - labels.  One warp per row; deterministic.
#include "pf_common.cuh"

namespace {

__device__ __forceinline__ float row_max(const float* __restrict__ r, int k, float inv_t, int lane) {
  float m = -INFINITY;
  for (int j = lane; j < k; j += 32) m = fmaxf(m, __fdiv_rn(__ldg(r + j), inv_t));
  return pf_warp_max(m);
}

// inv_t is the temperature T itself (the reference DIVIDES by T); name kept short.
__device__ __forceinline__ float row_sumexp(const float* __restrict__ r, int k, float T, float mx, int lane) {
  float s = 0.f;
  for (int j = lane; j < k; j += 32) s += expf(__fsub_rn(__fdiv_rn(__ldg(r + j), T), mx));
  return pf_warp_sum(s);
}

__global__ void __launch_bounds__(256)
softmax_ce_rows_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                       const float* __restrict__ teacher, int n, int k, float T, float w_dst,
                       float* __restrict__ dlogits, float* __restrict__ row_ws) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* s = logits + (size_t)row * k;
  const float* lab = labels + (size_t)row * k;
  const float inv_n = (float)n;  // divisor (the batch mean)

  // ---- hard CE
  const float mx = row_max(s, k, 1.f, lane);
  const float sum = row_sumexp(s, k, 1.f, mx, lane);
  const float lse = logf(sum);
  // ---- soft CE (distillation)
  float mx2 = 0.f, sum2 = 1.f, lse2 = 0.f, mxt = 0.f, sumt = 1.f;
  const float* t = teacher ? teacher + (size_t)row * k : nullptr;
  if (t) {
    mx2 = row_max(s, k, T, lane);
    sum2 = row_sumexp(s, k, T, mx2, lane);
    lse2 = logf(sum2);
    mxt = row_max(t, k, T, lane);
    sumt = row_sumexp(t, k, T, mxt, lane);
  }
  float loss_h = 0.f, loss_d = 0.f;
  float best_s = -INFINITY, best_l = -INFINITY;
  int arg_s = 0x7fffffff, arg_l = 0x7fffffff;
  for (int j = lane; j < k; j += 32) {
    const float sj = __ldg(s + j), lj = __ldg(lab + j);
    const float z = __fsub_rn(sj, mx);
    loss_h += __fmul_rn(lj, __fsub_rn(lse, z));
    float g = __fdiv_rn(__fsub_rn(__fdiv_rn(expf(z), sum), lj), inv_n);
    if (t) {
      const float z2 = __fsub_rn(__fdiv_rn(sj, T), mx2);
      const float pt = __fdiv_rn(expf(__fsub_rn(__fdiv_rn(__ldg(t + j), T), mxt)), sumt);
      loss_d += __fmul_rn(pt, __fsub_rn(lse2, z2));
      const float gd = __fdiv_rn(__fsub_rn(__fdiv_rn(expf(z2), sum2), pt), inv_n);
      g = __fadd_rn(g, __fdiv_rn(__fmul_rn(gd, w_dst), T));
    }
    dlogits[(size_t)row * k + j] = g;
    if (sj > best_s) { best_s = sj; arg_s = j; }   // strict: first index wins within a lane
    if (lj > best_l) { best_l = lj; arg_l = j; }
  }
  loss_h = pf_warp_sum(loss_h);
  loss_d = pf_warp_sum(loss_d);
  // argmax across lanes, lowest index on ties (np.argmax / tf.argmax)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, best_s, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg_s, o);
    if (os > best_s || (os == best_s && oa < arg_s)) { best_s = os; arg_s = oa; }
    const float ol = __shfl_xor_sync(0xffffffffu, best_l, o);
    const int ob = __shfl_xor_sync(0xffffffffu, arg_l, o);
    if (ol > best_l || (ol == best_l && ob < arg_l)) { best_l = ol; arg_l = ob; }
  }
  // top-5 (tf.nn.in_top_k): the target is in the top 5 iff fewer than 5 classes score strictly higher
  const float target = __ldg(s + arg_l);
  int higher = 0;
  for (int j = lane; j < k; j += 32) higher += (__ldg(s + j) > target) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) higher += __shfl_xor_sync(0xffffffffu, higher, o);
  if (lane == 0) {
    row_ws[row] = loss_h;
    row_ws[n + row] = loss_d;
    row_ws[2 * n + row] = (arg_s == arg_l) ? 1.f : 0.f;
    row_ws[3 * n + row] = (higher < 5) ? 1.f : 0.f;
  }
}

// Fixed-order block reduction: deterministic regardless of scheduling.
__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  v = pf_warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < 8 ? sh[threadIdx.x] : 0.f;
    r = pf_warp_sum(r);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

__global__ void __launch_bounds__(256)
softmax_ce_final_kernel(const float* __restrict__ row_ws, int n, float w_dst, int has_teacher,
                        float* __restrict__ out) {
  __shared__ float sh[8];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < n; i += 256) {
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] += row_ws[q * n + i];
  }
  float r[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) r[q] = block_sum_256(acc[q], sh);
  if (threadIdx.x == 0) {
    const float fn = (float)n;
    out[0] = __fdiv_rn(r[0], fn);
    out[1] = has_teacher ? __fmul_rn(w_dst, __fdiv_rn(r[1], fn)) : 0.f;
    out[2] = __fdiv_rn(r[2], fn);
    out[3] = __fdiv_rn(r[3], fn);
  }
}

__global__ void __launch_bounds__(256)
l2_partial_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ partial) {
  __shared__ float sh[8];
  float acc = 0.f;
  const int64_t nvec = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += stride) {
    const float4 x = pf_ld_stream(v + (i << 2));
    acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float x = __ldg(v + (nvec << 2) + threadIdx.x);
    acc += x * x;
  }
  const float r = block_sum_256(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void __launch_bounds__(256)
l2_final_kernel(const float* __restrict__ partial, int nparts, float scale, int accumulate,
                float* __restrict__ out) {
  __shared__ float sh[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 256) acc += partial[i];
  const float r = block_sum_256(acc, sh);
  if (threadIdx.x == 0) {
    const float val = __fmul_rn(scale, __fdiv_rn(r, 2.f));
    out[0] = accumulate ? __fadd_rn(out[0], val) : val;
  }
}

}  // namespace

extern "C" {

int pf_softmax_ce_fwd_bwd(const float* logits_dev, const float* labels_dev,
                          const float* teacher_dev, int n, int k, float tempr, float w_dst,
                          float* dlogits_dev, float* out_dev, float* row_ws_dev, void* stream) {
  PF_REQUIRE(n > 0 && k > 0, "pf_softmax_ce_fwd_bwd: n and k must be positive (n=%d k=%d)", n, k);
  PF_REQUIRE(logits_dev && labels_dev && dlogits_dev && out_dev && row_ws_dev,
             "pf_softmax_ce_fwd_bwd: null pointer");
  PF_REQUIRE(teacher_dev == nullptr || tempr > 0.f, "pf_softmax_ce_fwd_bwd: temperature must be > 0");
  const int rows_per_cta = 8;
  softmax_ce_rows_kernel<<<(n + rows_per_cta - 1) / rows_per_cta, 256, 0, (cudaStream_t)stream>>>(
      logits_dev, labels_dev, teacher_dev, n, k, tempr, w_dst, dlogits_dev, row_ws_dev);
  PF_CHECK_LAUNCH("pf_softmax_ce_fwd_bwd/rows");
  softmax_ce_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(row_ws_dev, n, w_dst,
                                                              teacher_dev != nullptr, out_dev);
  PF_CHECK_LAUNCH("pf_softmax_ce_fwd_bwd/final");
  return PF_OK;
}

int pf_l2_loss(const float* v_dev, int64_t n, float scale, int accumulate, float* out_dev,
               float* partial_ws_dev, void* stream) {
  PF_REQUIRE(n >= 0, "pf_l2_loss: n < 0");
  PF_REQUIRE(out_dev && partial_ws_dev && (v_dev || n == 0), "pf_l2_loss: null pointer");
  PF_REQUIRE(((uintptr_t)v_dev & 15) == 0, "pf_l2_loss: v must be 16-byte aligned");
  int64_t want = ((n >> 2) + 256 * 4 - 1) / (256 * 4);
  if (want < 1) want = 1;
  const int parts = (int)(want < PF_L2_PARTIALS ? want : PF_L2_PARTIALS);
  l2_partial_kernel<<<parts, 256, 0, (cudaStream_t)stream>>>(v_dev, n, partial_ws_dev);
  PF_CHECK_LAUNCH("pf_l2_loss/partial");
  l2_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partial_ws_dev, parts, scale, accumulate, out_dev);
  PF_CHECK_LAUNCH("pf_l2_loss/final");
  return PF_OK;
}

}  // extern "C"
