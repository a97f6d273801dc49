// CLIP image preprocessing on the GPU, bit-exact with what the reference's drivers run on the CPU before model.generate:
//   image_processor.preprocess(image, return_tensors='pt')['pixel_values'][0]
// (reference experiments/eval/sampling/llava_sampling.py:82, eval/calibrate/llava_calibrate.py:149; processor = transformers-4.31
// CLIPImageProcessor of openai/clip-vit-large-patch14-336, builder.py:137-141) and, optionally, the expand2square padding of
// experiments/llava/mm_utils.py:152-164.  Sequence: [pad to a square with the mean colour] -> resize the shortest edge to S with
// Pillow's BICUBIC (ImagingResample: separable, anti-aliased, 8-bit fixed point with 22 fractional bits, horizontal pass first,
// each pass rounded to uint8) -> center crop S x S -> * (1/255) (float64, rounded to float32) -> (x - mean) / std in float32.
//
// Two kernels: the horizontal pass over exactly the input rows and output columns the crop needs (uint8 out), then the vertical
// pass fused with crop, rescale, normalise and the HWC -> CHW transpose.  The tap tables are computed on the host in double
// precision with Pillow's formula (precompute_coeffs + normalize_coeffs_8bpc) and uploaded; all device arithmetic is integer or
// explicitly rounded fp32 / fp64, so the result equals the CPU pipeline bit for bit (oracle/clip_preprocess_oracle.py, pinned to
// Pillow itself by tests/test_clip_preprocess.py).  Pure byte / integer work: HBM-bound, no tensor cores.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <vector>

#include "kernels.h"

namespace vdd {

constexpr int PP_PRECISION_BITS = 32 - 8 - 2;

static double pp_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for output positions [o0, o1) of an axis resized in_size -> out_size.
// bounds: (first input index, tap count) per position; taps: [o1 - o0][ksize] fixed-point weights.
static int pp_coeffs(int in_size, int out_size, int o0, int o1, std::vector<int32_t>& bounds, std::vector<int32_t>& taps) {
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  const double ss = 1.0 / filterscale;
  bounds.assign((size_t)(o1 - o0) * 2, 0);
  taps.assign((size_t)(o1 - o0) * ksize, 0);
  std::vector<double> w(ksize);
  for (int xx = o0; xx < o1; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = pp_bicubic((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      const double k = ww != 0.0 ? w[x] / ww : w[x];
      taps[(size_t)(xx - o0) * ksize + x] =
          k < 0 ? (int32_t)(-0.5 + k * (double)(1 << PP_PRECISION_BITS)) : (int32_t)(0.5 + k * (double)(1 << PP_PRECISION_BITS));
    }
    bounds[(size_t)(xx - o0) * 2] = xmin;
    bounds[(size_t)(xx - o0) * 2 + 1] = xmax;
  }
  return ksize;
}

struct PpGeom {
  int H, W;            // source image
  int side_h, side_w;  // canvas (after the optional square padding)
  int pad_top, pad_left;
  uint8_t bg[3];
  int S;               // output edge
  int row0, nrows;     // canvas rows the vertical pass reads: [row0, row0 + nrows)
};

__device__ __forceinline__ uint8_t pp_src(const uint8_t* __restrict__ img, const PpGeom& g, int y, int x, int c) {
  y -= g.pad_top;
  x -= g.pad_left;
  if (y < 0 || y >= g.H || x < 0 || x >= g.W) return g.bg[c];
  return img[((size_t)y * g.W + x) * 3 + c];
}
__device__ __forceinline__ uint8_t pp_clip8(int v) {
  v >>= PP_PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[r][c][3], r in [0, nrows), c in [0, S): horizontal pass of canvas row row0 + r at output column left + c (identity when the
// width is unchanged: bounds == null)
__global__ void __launch_bounds__(256) pp_horizontal_kernel(const uint8_t* __restrict__ img, PpGeom g, const int32_t* __restrict__ bounds,
                                                           const int32_t* __restrict__ taps, int ksize, int left, uint8_t* __restrict__ tmp) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= g.S) return;
  const int y = g.row0 + r;
  uint8_t* o = tmp + ((size_t)r * g.S + c) * 3;
  if (!bounds) {
    for (int ch = 0; ch < 3; ++ch) o[ch] = pp_src(img, g, y, left + c, ch);
    return;
  }
  const int x0 = bounds[2 * c], n = bounds[2 * c + 1];
  const int32_t* k = taps + (size_t)c * ksize;
  int a0 = 1 << (PP_PRECISION_BITS - 1), a1 = a0, a2 = a0;
  for (int t = 0; t < n; ++t) {
    const int kv = k[t];
    a0 += (int)pp_src(img, g, y, x0 + t, 0) * kv;
    a1 += (int)pp_src(img, g, y, x0 + t, 1) * kv;
    a2 += (int)pp_src(img, g, y, x0 + t, 2) * kv;
  }
  o[0] = pp_clip8(a0);
  o[1] = pp_clip8(a1);
  o[2] = pp_clip8(a2);
}

template <typename T>
__device__ __forceinline__ T pp_cast(float v);
template <> __device__ __forceinline__ float pp_cast<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 pp_cast<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half pp_cast<__half>(float v) { return __float2half_rn(v); }

// out[ch][r][c]: vertical pass at output row top + r over tmp (identity when the height is unchanged), then rescale + normalise
template <typename T>
__global__ void __launch_bounds__(256) pp_vertical_kernel(const uint8_t* __restrict__ tmp, PpGeom g, const int32_t* __restrict__ bounds,
                                                         const int32_t* __restrict__ taps, int ksize, float m0, float m1, float m2, float s0,
                                                         float s1, float s2, T* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= g.S) return;
  uint8_t px[3];
  if (!bounds) {
    const uint8_t* p = tmp + ((size_t)r * g.S + c) * 3;  // tmp row r IS output row r (row0 = top)
    px[0] = p[0]; px[1] = p[1]; px[2] = p[2];
  } else {
    const int y0 = bounds[2 * r] - g.row0, n = bounds[2 * r + 1];
    const int32_t* k = taps + (size_t)r * ksize;
    int a0 = 1 << (PP_PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < n; ++t) {
      const uint8_t* p = tmp + ((size_t)(y0 + t) * g.S + c) * 3;
      const int kv = k[t];
      a0 += (int)p[0] * kv;
      a1 += (int)p[1] * kv;
      a2 += (int)p[2] * kv;
    }
    px[0] = pp_clip8(a0); px[1] = pp_clip8(a1); px[2] = pp_clip8(a2);
  }
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    // numpy: uint8 * (1/255) is a float64 product, cast to float32; then (x - mean) / std in float32
    const float x = (float)((double)px[ch] * (1.0 / 255.0));
    const float v = __fdiv_rn(__fsub_rn(x, mean[ch]), sd[ch]);
    out[((size_t)ch * g.S + r) * g.S + c] = pp_cast<T>(v);
  }
}

// scratch shared by the calls of one process (single host thread by contract, like the engine)
static uint8_t* g_pp_img = nullptr;
static size_t g_pp_img_bytes = 0;
static uint8_t* g_pp_tmp = nullptr;
static size_t g_pp_tmp_bytes = 0;
static int32_t* g_pp_tab = nullptr;
static size_t g_pp_tab_bytes = 0;
static int32_t* g_pp_tab_host = nullptr;  // pinned staging of the tap tables

static cudaError_t pp_reserve(void** p, size_t* have, size_t need, bool pinned = false) {
  if (*have >= need) return cudaSuccess;
  cudaError_t e = cudaDeviceSynchronize();  // earlier calls may still read the old buffer
  if (e != cudaSuccess) return e;
  if (*p) pinned ? cudaFreeHost(*p) : cudaFree(*p);
  *p = nullptr;
  *have = 0;
  need = need + need / 4 + 4096;
  e = pinned ? cudaMallocHost(p, need) : cudaMalloc(p, need);
  if (e == cudaSuccess) *have = need;
  return e;
}

cudaError_t launch_clip_preprocess(const uint8_t* rgb, int H, int W, bool on_device, int S, const uint8_t* pad_rgb, const float* mean,
                                   const float* stdv, void* out, int out_dtype, cudaStream_t st) {
  const bool pad_to_square = pad_rgb != nullptr;
  if (!rgb || !out || H <= 0 || W <= 0 || S <= 0 || out_dtype < 0 || out_dtype > 2) return cudaErrorInvalidValue;
  PpGeom g{};
  g.H = H; g.W = W; g.S = S;
  g.side_h = H; g.side_w = W;
  if (pad_to_square && H != W) {
    const int side = H > W ? H : W;
    g.side_h = g.side_w = side;
    g.pad_top = (side - H) / 2;
    g.pad_left = (side - W) / 2;
    for (int c = 0; c < 3; ++c) g.bg[c] = pad_rgb[c];  // the caller's tuple(int(x * 255) for x in image_mean)
  }
  // transformers 4.31 get_resize_output_image_size(shortest_edge = S, default_to_square = False)
  const int ch = g.side_h, cw = g.side_w;
  int out_h, out_w;
  if (cw <= ch) { out_w = S; out_h = (int)((double)S * (double)ch / (double)cw); }
  else { out_h = S; out_w = (int)((double)S * (double)cw / (double)ch); }
  const int top = (out_h - S) / 2, left = (out_w - S) / 2;
  if (out_h < S || out_w < S) return cudaErrorInvalidValue;

  std::vector<int32_t> hb, ht, vb, vt;
  int hk = 0, vk = 0;
  const bool need_h = out_w != cw, need_v = out_h != ch;
  if (need_h) hk = pp_coeffs(cw, out_w, left, left + S, hb, ht);
  if (need_v) vk = pp_coeffs(ch, out_h, top, top + S, vb, vt);
  g.row0 = top;
  g.nrows = S;
  if (need_v) {
    g.row0 = vb[0];
    g.nrows = vb[(size_t)(S - 1) * 2] + vb[(size_t)(S - 1) * 2 + 1] - g.row0;
  }
  const size_t tab_ints = hb.size() + ht.size() + vb.size() + vt.size() + 4;
  cudaError_t e;
  static size_t host_bytes = 0;
  if ((e = pp_reserve((void**)&g_pp_tab, &g_pp_tab_bytes, tab_ints * 4)) != cudaSuccess) return e;
  if ((e = pp_reserve((void**)&g_pp_tab_host, &host_bytes, tab_ints * 4, true)) != cudaSuccess) return e;
  if ((e = pp_reserve((void**)&g_pp_tmp, &g_pp_tmp_bytes, (size_t)g.nrows * S * 3)) != cudaSuccess) return e;
  const uint8_t* img = rgb;
  if (!on_device) {
    if ((e = pp_reserve((void**)&g_pp_img, &g_pp_img_bytes, (size_t)H * W * 3)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(g_pp_img, rgb, (size_t)H * W * 3, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    img = g_pp_img;
  }
  // the previous call's table copy must have left the pinned staging before it is overwritten
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
  size_t o = 0;
  auto put = [&](const std::vector<int32_t>& v) {
    const size_t at = o;
    for (size_t i = 0; i < v.size(); ++i) g_pp_tab_host[o++] = v[i];
    return at;
  };
  const size_t o_hb = put(hb), o_ht = put(ht), o_vb = put(vb), o_vt = put(vt);
  if (o && (e = cudaMemcpyAsync(g_pp_tab, g_pp_tab_host, o * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;

  dim3 blk(256);
  pp_horizontal_kernel<<<dim3((S + 255) / 256, g.nrows), blk, 0, st>>>(img, g, need_h ? g_pp_tab + o_hb : nullptr, g_pp_tab + o_ht, hk, left,
                                                                        g_pp_tmp);
  const int32_t* vbp = need_v ? g_pp_tab + o_vb : nullptr;
  dim3 grid((S + 255) / 256, S);
  if (out_dtype == 0)
    pp_vertical_kernel<float><<<grid, blk, 0, st>>>(g_pp_tmp, g, vbp, g_pp_tab + o_vt, vk, mean[0], mean[1], mean[2], stdv[0], stdv[1],
                                                    stdv[2], (float*)out);
  else if (out_dtype == 1)
    pp_vertical_kernel<__nv_bfloat16><<<grid, blk, 0, st>>>(g_pp_tmp, g, vbp, g_pp_tab + o_vt, vk, mean[0], mean[1], mean[2], stdv[0],
                                                            stdv[1], stdv[2], (__nv_bfloat16*)out);
  else
    pp_vertical_kernel<__half><<<grid, blk, 0, st>>>(g_pp_tmp, g, vbp, g_pp_tab + o_vt, vk, mean[0], mean[1], mean[2], stdv[0], stdv[1],
                                                     stdv[2], (__half*)out);
  return cudaGetLastError();
}

}  // namespace vdd
