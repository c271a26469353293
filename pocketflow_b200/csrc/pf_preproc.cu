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
#include "pf_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
preprocess_images_kernel(const uint8_t* __restrict__ crops, const pf_img_desc* __restrict__ desc, int n, int out_h,
                         int out_w, float mean_r, float mean_g, float mean_b, float* __restrict__ dst) {
  const int64_t per_image = (int64_t)out_h * out_w * 3;
  const int64_t total = per_image * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / per_image);
    int r = (int)(idx - (int64_t)i * per_image);
    const int c = r % 3;
    r /= 3;
    const int x = r % out_w, y = r / out_w;
    const pf_img_desc d = desc[i];
    const float sy = __fmul_rn((float)(y + d.top), __fdiv_rn((float)d.h, (float)d.rh));
    const float sx = __fmul_rn((float)(x + d.left), __fdiv_rn((float)d.w, (float)d.rw));
    const int y0 = (int)floorf(sy), x0 = (int)floorf(sx);
    const int y1 = min((int)ceilf(sy), d.h - 1), x1 = min((int)ceilf(sx), d.w - 1);
    const float ly = __fsub_rn(sy, (float)y0), lx = __fsub_rn(sx, (float)x0);
    const int xa = d.flip ? d.w - 1 - x0 : x0, xb = d.flip ? d.w - 1 - x1 : x1;
    const uint8_t* src = crops + d.offset;
    const int64_t row0 = (int64_t)y0 * d.w, row1 = (int64_t)y1 * d.w;
    const float tl = (float)src[(row0 + xa) * 3 + c], tr = (float)src[(row0 + xb) * 3 + c];
    const float bl = (float)src[(row1 + xa) * 3 + c], br = (float)src[(row1 + xb) * 3 + c];
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
    const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
    const float v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
    dst[idx] = __fsub_rn(v, c == 0 ? mean_r : (c == 1 ? mean_g : mean_b));
  }
}

}  // namespace

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
