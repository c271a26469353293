// pf_preproc.cu — ILSVRC-12 image preprocessing on the device: what utils/external/imagenet_preprocessing.py:225-260
// does after JPEG decoding (bilinear resize, optional left-right flip, central crop, mean subtraction), for a whole
// mini-batch of variable-sized uint8 crops in ONE launch.  HBM-bound byte work: ~1 MB per image (uint8 gathers through
// L1/L2, one coalesced fp32 store per output value).
//
// The resize is TF1's resize_bilinear with align_corners=False and no half-pixel centres:
//   src = dst_index * (in / out)   (fp32),  lo = floor(src), hi = min(ceil(src), in - 1), l = src - lo
//   top = tl + (tr - tl) * lx ; bot = bl + (br - bl) * lx ; out = top + (bot - top) * ly
// every operation individually rounded (no FMA contraction), so the result is bit-identical to the host
// restatement in pocketflow_b200/datasets/ilsvrc12_dataset.py.  The flip is applied to the SOURCE column index
// (training flips the crop before resizing it; with this asymmetric resize that differs from flipping afterwards).
//
// The per-value arithmetic lives in two __host__ __device__ functions so that a test can run the very same source on
// the host: tests/test_ilsvrc12_cpu.py compiles this file with -DPF_PREPROC_HOST_TEST into a scratch library that
// exposes a plain loop over them.  That loop is never part of libpf_b200.so — the product has no CPU path.
#ifndef PF_PREPROC_HOST_TEST
#include "pf_common.cuh"
#else
#include <math.h>
#include <stdint.h>
#include "pf_b200.h"
#endif

#ifdef __CUDA_ARCH__
#define PF_MUL(a, b) __fmul_rn((a), (b))
#define PF_ADD(a, b) __fadd_rn((a), (b))
#define PF_SUB(a, b) __fsub_rn((a), (b))
#define PF_DIV(a, b) __fdiv_rn((a), (b))
#else  // host pass (the scratch test build adds -ffp-contract=off)
#define PF_MUL(a, b) ((a) * (b))
#define PF_ADD(a, b) ((a) + (b))
#define PF_SUB(a, b) ((a) - (b))
#define PF_DIV(a, b) ((a) / (b))
#endif

namespace {

struct PixelIndex {
  int i, y, x, c;
};

// flat index into fp32 [n, out_h, out_w, 3]  ->  (image, row, column, channel)
__host__ __device__ inline PixelIndex preproc_decompose(int64_t idx, int64_t per_image, int out_w) {
  PixelIndex p;
  p.i = (int)(idx / per_image);
  int r = (int)(idx - (int64_t)p.i * per_image);
  p.c = r % 3;
  r /= 3;
  p.x = r % out_w;
  p.y = r / out_w;
  return p;
}

__host__ __device__ inline float preproc_value(const uint8_t* crops, const pf_img_desc& d, int y, int x, int c,
                                               float mean) {
  const float sy = PF_MUL((float)(y + d.top), PF_DIV((float)d.h, (float)d.rh));
  const float sx = PF_MUL((float)(x + d.left), PF_DIV((float)d.w, (float)d.rw));
  const int y0 = (int)floorf(sy), x0 = (int)floorf(sx);
  int y1 = (int)ceilf(sy), x1 = (int)ceilf(sx);
  y1 = y1 < d.h - 1 ? y1 : d.h - 1;
  x1 = x1 < d.w - 1 ? x1 : d.w - 1;
  const float ly = PF_SUB(sy, (float)y0), lx = PF_SUB(sx, (float)x0);
  const int xa = d.flip ? d.w - 1 - x0 : x0, xb = d.flip ? d.w - 1 - x1 : x1;
  const uint8_t* src = crops + d.offset;
  const int64_t row0 = (int64_t)y0 * d.w, row1 = (int64_t)y1 * d.w;
  const float tl = (float)src[(row0 + xa) * 3 + c], tr = (float)src[(row0 + xb) * 3 + c];
  const float bl = (float)src[(row1 + xa) * 3 + c], br = (float)src[(row1 + xb) * 3 + c];
  const float top = PF_ADD(tl, PF_MUL(PF_SUB(tr, tl), lx));
  const float bot = PF_ADD(bl, PF_MUL(PF_SUB(br, bl), lx));
  return PF_SUB(PF_ADD(top, PF_MUL(PF_SUB(bot, top), ly)), mean);
}

#ifndef PF_PREPROC_HOST_TEST
__global__ void __launch_bounds__(256)
preprocess_images_kernel(const uint8_t* __restrict__ crops, const pf_img_desc* __restrict__ desc, int n, int out_h,
                         int out_w, float mean_r, float mean_g, float mean_b, float* __restrict__ dst) {
  const int64_t per_image = (int64_t)out_h * out_w * 3;
  const int64_t total = per_image * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const PixelIndex p = preproc_decompose(idx, per_image, out_w);
    dst[idx] = preproc_value(crops, desc[p.i], p.y, p.x, p.c, p.c == 0 ? mean_r : (p.c == 1 ? mean_g : mean_b));
  }
}
#endif

}  // namespace

#ifndef PF_PREPROC_HOST_TEST
extern "C" int pf_preprocess_images(const uint8_t* crops_dev, const pf_img_desc* desc_dev, int n, int out_h, int out_w,
                                    float mean_r, float mean_g, float mean_b, float* dst_dev, void* stream) {
  PF_REQUIRE(n >= 0 && out_h > 0 && out_w > 0, "pf_preprocess_images: n=%d out=%dx%d", n, out_h, out_w);
  if (n == 0) return PF_OK;
  PF_REQUIRE(crops_dev && desc_dev && dst_dev, "pf_preprocess_images: null pointer");
  const int64_t total = (int64_t)n * out_h * out_w * 3;
  const int64_t want = (total + 255) / 256;
  const int blocks = (int)(want < (int64_t)PF_NUM_SMS * 16 ? want : (int64_t)PF_NUM_SMS * 16);
  preprocess_images_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(crops_dev, desc_dev, n, out_h, out_w, mean_r,
                                                                      mean_g, mean_b, dst_dev);
  PF_CHECK_LAUNCH("pf_preprocess_images");
  return PF_OK;
}
#else
// test-only: the kernel's loop body on the host, same indexing, same per-value function
extern "C" void pf_test_preprocess_host(const uint8_t* crops, const pf_img_desc* desc, int n, int out_h, int out_w,
                                        float mean_r, float mean_g, float mean_b, float* dst) {
  const int64_t per_image = (int64_t)out_h * out_w * 3;
  for (int64_t idx = 0; idx < per_image * n; ++idx) {
    const PixelIndex p = preproc_decompose(idx, per_image, out_w);
    dst[idx] = preproc_value(crops, desc[p.i], p.y, p.x, p.c, p.c == 0 ? mean_r : (p.c == 1 ? mean_g : mean_b));
  }
}
#endif
