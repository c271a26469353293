// pf_optim.cu — fused optimizer steps over flat fp32 ranges (HBM-bound).
//
// Replaces, per trainable variable in the reference, one `grad * mask` kernel
// (/root/reference/learners/weight_sparsification/learner.py:314-332), the AddN that merges the
// l2_loss gradient, Horovod's post-allreduce division (utils/multi_gpu_wrapper.py:82-89) and one
// ApplyMomentum / ApplyAdam kernel (weight_sparsification/learner.py:201;
// uniform_quantization/learner.py:244) — ~110 launches per step on ResNet-50 — by ONE launch over
// the flat parameter / gradient / slot buffers.  Algorithmic bytes: momentum 24 B/elem with a mask
// (20 without), Adam 28 B/elem.
#include "pf_common.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kUnroll = 2;

inline unsigned flat_grid(int64_t n) {
  int64_t want = ((n >> 2) + kThreads * kUnroll - 1) / (kThreads * kUnroll);
  const int64_t cap = (int64_t)PF_NUM_SMS * 8;
  if (want < 1) want = 1;
  return (unsigned)(want < cap ? want : cap);
}

__device__ __forceinline__ float grad_total(float g, float w, float gs, float wd) {
  // g*grad_scale (Horovod average) then + wd*w (gradient of loss_w_dcy*l2_loss), separately rounded
  float t = __fmul_rn(g, gs);
  return wd != 0.f ? __fadd_rn(t, __fmul_rn(wd, w)) : t;
}

template <bool MASKED>
__device__ __forceinline__ void mom1(float& w, float& a, float g, float m, float lr, float mom,
                                     float wd, float gs) {
  float gt = grad_total(g, w, gs, wd);
  if (MASKED) gt = __fmul_rn(gt, m);
  a = __fadd_rn(__fmul_rn(a, mom), gt);
  w = __fsub_rn(w, __fmul_rn(a, lr));
}

template <bool MASKED>
__global__ void __launch_bounds__(kThreads)
momentum_kernel(float* __restrict__ w, float* __restrict__ acc, const float* __restrict__ g,
                const float* __restrict__ mask, int64_t n, const float* __restrict__ hp, float mom,
                float wd, float gs) {
  const float lr = __ldg(hp);
  const int64_t nvec = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += stride) {
    const int64_t o = i << 2;
    float4 vw = pf_ld4(w + o), va = pf_ld4(acc + o);
    const float4 vg = pf_ld_stream(g + o);
    float4 vm = make_float4(1.f, 1.f, 1.f, 1.f);
    if (MASKED) vm = pf_ld_stream(mask + o);
    mom1<MASKED>(vw.x, va.x, vg.x, vm.x, lr, mom, wd, gs);
    mom1<MASKED>(vw.y, va.y, vg.y, vm.y, lr, mom, wd, gs);
    mom1<MASKED>(vw.z, va.z, vg.z, vm.z, lr, mom, wd, gs);
    mom1<MASKED>(vw.w, va.w, vg.w, vm.w, lr, mom, wd, gs);
    pf_st_stream(acc + o, va);
    pf_st_stream(w + o, vw);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t j = (nvec << 2) + threadIdx.x;
    float vw = w[j], va = acc[j];
    mom1<MASKED>(vw, va, g[j], MASKED ? mask[j] : 1.f, lr, mom, wd, gs);
    acc[j] = va;
    w[j] = vw;
  }
}

__device__ __forceinline__ void adam1(float& w, float& m, float& v, float g, float alpha, float omb1,
                                      float omb2, float eps, float wd, float gs) {
  const float gt = grad_total(g, w, gs, wd);
  m = __fadd_rn(m, __fmul_rn(__fsub_rn(gt, m), omb1));
  v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(gt, gt), v), omb2));
  w = __fsub_rn(w, __fdiv_rn(__fmul_rn(m, alpha), __fadd_rn(__fsqrt_rn(v), eps)));
}

__global__ void __launch_bounds__(kThreads)
adam_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
            const float* __restrict__ g, int64_t n, const float* __restrict__ hp, float beta1,
            float beta2, float eps, float wd, float gs) {
  const float lr = __ldg(hp), b1p = __ldg(hp + 1), b2p = __ldg(hp + 2);
  // alpha = lr * sqrt(1 - beta2_power) / (1 - beta1_power)   (TF ApplyAdam)
  const float alpha = __fdiv_rn(__fmul_rn(lr, __fsqrt_rn(__fsub_rn(1.f, b2p))), __fsub_rn(1.f, b1p));
  const float omb1 = __fsub_rn(1.f, beta1), omb2 = __fsub_rn(1.f, beta2);
  const int64_t nvec = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += stride) {
    const int64_t o = i << 2;
    float4 vw = pf_ld4(w + o), vm = pf_ld4(m + o), vv = pf_ld4(v + o);
    const float4 vg = pf_ld_stream(g + o);
    adam1(vw.x, vm.x, vv.x, vg.x, alpha, omb1, omb2, eps, wd, gs);
    adam1(vw.y, vm.y, vv.y, vg.y, alpha, omb1, omb2, eps, wd, gs);
    adam1(vw.z, vm.z, vv.z, vg.z, alpha, omb1, omb2, eps, wd, gs);
    adam1(vw.w, vm.w, vv.w, vg.w, alpha, omb1, omb2, eps, wd, gs);
    pf_st_stream(m + o, vm);
    pf_st_stream(v + o, vv);
    pf_st_stream(w + o, vw);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t j = (nvec << 2) + threadIdx.x;
    float vw = w[j], vm = m[j], vv = v[j];
    adam1(vw, vm, vv, g[j], alpha, omb1, omb2, eps, wd, gs);
    m[j] = vm;
    v[j] = vv;
    w[j] = vw;
  }
}
}  // namespace

extern "C" {

int pf_momentum_step(float* w_dev, float* acc_dev, const float* g_dev, const float* mask_dev,
                     int64_t n, const float* hp_dev, float momentum, float wd, float grad_scale,
                     void* stream) {
  PF_REQUIRE(n >= 0, "pf_momentum_step: n < 0");
  if (n == 0) return PF_OK;
  PF_REQUIRE(w_dev && acc_dev && g_dev && hp_dev, "pf_momentum_step: null pointer");
  PF_REQUIRE((((uintptr_t)w_dev | (uintptr_t)acc_dev | (uintptr_t)g_dev | (uintptr_t)mask_dev) & 15) == 0,
             "pf_momentum_step: buffers must be 16-byte aligned");
  if (mask_dev)
    momentum_kernel<true><<<flat_grid(n), kThreads, 0, (cudaStream_t)stream>>>(
        w_dev, acc_dev, g_dev, mask_dev, n, hp_dev, momentum, wd, grad_scale);
  else
    momentum_kernel<false><<<flat_grid(n), kThreads, 0, (cudaStream_t)stream>>>(
        w_dev, acc_dev, g_dev, nullptr, n, hp_dev, momentum, wd, grad_scale);
  PF_CHECK_LAUNCH("pf_momentum_step");
  return PF_OK;
}

int pf_adam_step(float* w_dev, float* m_dev, float* v_dev, const float* g_dev, int64_t n,
                 const float* hp_dev, float beta1, float beta2, float eps, float wd,
                 float grad_scale, void* stream) {
  PF_REQUIRE(n >= 0, "pf_adam_step: n < 0");
  if (n == 0) return PF_OK;
  PF_REQUIRE(w_dev && m_dev && v_dev && g_dev && hp_dev, "pf_adam_step: null pointer");
  PF_REQUIRE((((uintptr_t)w_dev | (uintptr_t)m_dev | (uintptr_t)v_dev | (uintptr_t)g_dev) & 15) == 0,
             "pf_adam_step: buffers must be 16-byte aligned");
  adam_kernel<<<flat_grid(n), kThreads, 0, (cudaStream_t)stream>>>(w_dev, m_dev, v_dev, g_dev, n,
                                                                   hp_dev, beta1, beta2, eps, wd,
                                                                   grad_scale);
  PF_CHECK_LAUNCH("pf_adam_step");
  return PF_OK;
}

}  // extern "C"
