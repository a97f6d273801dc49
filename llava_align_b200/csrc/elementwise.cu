// Glue kernels of the VDD path: norms, embedding gather + image splice, RoPE + KV-cache write for prefill,
// patch extraction for the ViT, row casts.  All are pure HBM streams (one pass over their rows, 16-byte accesses).
#include "common.cuh"
#include "kernels.h"

namespace vdd {

// ------------------------------------------------------------------------------------------------ RMSNorm
// HF LlamaRMSNorm (fp32 statistics): out = w * (x * rsqrt(mean(x^2) + eps)), rounded to bf16 for the GEMM.
__global__ void __launch_bounds__(256) rmsnorm_bf16_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                          float eps, __nv_bfloat16* __restrict__ out, int H,
                                                          const int32_t* __restrict__ row_idx) {
  __shared__ float red[33];
  pdl_wait();  // no-op unless launched with programmatic stream serialization (decode chain)
  pdl_launch_dependents();
  const int row = blockIdx.x;
  const float* xr = x + (size_t)(row_idx ? row_idx[row] : row) * H;
  float ss = 0.f;
  for (int k = threadIdx.x * 4; k < H; k += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(xr + k);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  const float rstd = rsqrtf(block_reduce<false>(ss, red) / (float)H + eps);
  for (int k = threadIdx.x * 4; k < H; k += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(xr + k);
    const uint2 wv = *reinterpret_cast<const uint2*>(w + k);
    uint2 o;
    o.x = pack_bf16(v.x * rstd * bf16lo(wv.x), v.y * rstd * bf16hi(wv.x));
    o.y = pack_bf16(v.z * rstd * bf16lo(wv.y), v.w * rstd * bf16hi(wv.y));
    *reinterpret_cast<uint2*>(out + (size_t)row * H + k) = o;
  }
}
cudaError_t launch_rmsnorm_bf16(const float* x, const void* w, float eps, void* out, int rows, int H,
                                const int32_t* row_idx, cudaStream_t stream, bool pdl) {
  if (rows <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(rows);
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, rmsnorm_bf16_kernel, x, (const __nv_bfloat16*)w, eps, (__nv_bfloat16*)out, H, row_idx);
}

// Decode with many rows: the split-K planes of a residual GEMM are summed into the residual stream (plane order =>
// deterministic) and the NEXT RMSNorm is applied in the same pass: resid[r] += sum_z ws[z][r];  out[r] = bf16(norm(resid[r]) * w).
// splits == 0: the GEMM already accumulated into resid, only the norm remains.
// 1024 threads per row: one float4 of the row per thread (H = 4096), so the residual and all split planes of a row are
// requested at once -- the pass is a handful of dependent L2 round trips, not a bandwidth problem.
constexpr int RR_THREADS = 1024;
// bounded spin on a cross-GPU arrival counter (peer ranks bump it with red.sys after a system fence)
__device__ __forceinline__ void tp_counter_wait(const unsigned* counter, unsigned per_epoch, const int32_t* epoch) {
  const unsigned target = (unsigned)(*epoch + 1) * per_epoch;
  unsigned spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v >= target) break;
    if ((++spins & 0x3ffu) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      if (!t0) t0 = now;
      else if (now - t0 > 30000000000ull) __trap();  // a peer never arrived: error return instead of a hung box
    }
  }
  __threadfence_system();
}
__global__ void __launch_bounds__(RR_THREADS) reduce_rmsnorm_kernel(float* __restrict__ resid, const float* __restrict__ ws, int splits,
                                                                    int64_t plane_stride, const __nv_bfloat16* __restrict__ w, float eps,
                                                                    __nv_bfloat16* __restrict__ out, int H, const unsigned* wait,
                                                                    unsigned wait_per_epoch, const int32_t* epoch) {
  __shared__ float red[33];
  pdl_wait();
  pdl_launch_dependents();
  if (wait) {  // tensor parallel: the planes are the partial sums of all ranks, pushed into this rank's inbox by their GEMM epilogues
    if (threadIdx.x == 0) tp_counter_wait(wait, wait_per_epoch, epoch);
    __syncthreads();
  }
  const int row = blockIdx.x;
  float* xr = resid + (size_t)row * H;
  const float* wr = ws + (size_t)row * H;
  float ss = 0.f;
  for (int k = threadIdx.x * 4; k < H; k += RR_THREADS * 4) {
    float4 a = *reinterpret_cast<const float4*>(xr + k);
    int z = 0;
    for (; z + 4 <= splits; z += 4) {  // four planes in flight, added in plane order
      const float4 b0 = __ldcg(reinterpret_cast<const float4*>(wr + (int64_t)z * plane_stride + k));
      const float4 b1 = __ldcg(reinterpret_cast<const float4*>(wr + (int64_t)(z + 1) * plane_stride + k));
      const float4 b2 = __ldcg(reinterpret_cast<const float4*>(wr + (int64_t)(z + 2) * plane_stride + k));
      const float4 b3 = __ldcg(reinterpret_cast<const float4*>(wr + (int64_t)(z + 3) * plane_stride + k));
      a.x += b0.x; a.y += b0.y; a.z += b0.z; a.w += b0.w;
      a.x += b1.x; a.y += b1.y; a.z += b1.z; a.w += b1.w;
      a.x += b2.x; a.y += b2.y; a.z += b2.z; a.w += b2.w;
      a.x += b3.x; a.y += b3.y; a.z += b3.z; a.w += b3.w;
    }
    for (; z < splits; ++z) {
      const float4 b = __ldcg(reinterpret_cast<const float4*>(wr + (int64_t)z * plane_stride + k));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (splits > 0) *reinterpret_cast<float4*>(xr + k) = a;
    ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  const float rstd = rsqrtf(block_reduce<false>(ss, red) / (float)H + eps);
  for (int k = threadIdx.x * 4; k < H; k += RR_THREADS * 4) {
    const float4 v = *reinterpret_cast<const float4*>(xr + k);  // this thread's own stores
    const uint2 wv = *reinterpret_cast<const uint2*>(w + k);
    uint2 o;
    o.x = pack_bf16(v.x * rstd * bf16lo(wv.x), v.y * rstd * bf16hi(wv.x));
    o.y = pack_bf16(v.z * rstd * bf16lo(wv.y), v.w * rstd * bf16hi(wv.y));
    *reinterpret_cast<uint2*>(out + (size_t)row * H + k) = o;
  }
}
__global__ void tp_wait_kernel(const unsigned* counter, unsigned per_epoch, const int32_t* epoch) {
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) tp_counter_wait(counter, per_epoch, epoch);
}
cudaError_t launch_tp_wait(const unsigned* counter, unsigned per_epoch, const int32_t* epoch, cudaStream_t stream) {
  tp_wait_kernel<<<1, 32, 0, stream>>>(counter, per_epoch, epoch);
  return cudaGetLastError();
}
cudaError_t launch_reduce_rmsnorm(float* resid, const float* ws, int splits, int rows, int H, const void* w, float eps, void* out,
                                  cudaStream_t stream, bool pdl, const unsigned* wait, unsigned wait_per_epoch, const int32_t* epoch) {
  if (rows <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(rows);
  cfg.blockDim = dim3(RR_THREADS);
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, reduce_rmsnorm_kernel, resid, ws, splits, (int64_t)rows * H, (const __nv_bfloat16*)w, eps,
                            (__nv_bfloat16*)out, H, wait, wait_per_epoch, epoch);
}

// ------------------------------------------------------------------------------------------------ LayerNorm (CLIP)
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                       const __nv_bfloat16* __restrict__ b, float eps,
                                                       __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32,
                                                       int H, const int32_t* __restrict__ row_idx) {
  __shared__ float red[33];
  const int row = blockIdx.x;
  const float* xr = x + (size_t)(row_idx ? row_idx[row] : row) * H;
  float s = 0.f;
  for (int k = threadIdx.x * 4; k < H; k += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(xr + k);
    s += v.x + v.y + v.z + v.w;
  }
  const float mean = block_reduce<false>(s, red) / (float)H;
  float ss = 0.f;
  for (int k = threadIdx.x * 4; k < H; k += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(xr + k);
    const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
    ss += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
  }
  const float rstd = rsqrtf(block_reduce<false>(ss, red) / (float)H + eps);
  for (int k = threadIdx.x * 4; k < H; k += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(xr + k);
    const uint2 wv = *reinterpret_cast<const uint2*>(w + k);
    const uint2 bv = *reinterpret_cast<const uint2*>(b + k);
    float4 y;
    y.x = (v.x - mean) * rstd * bf16lo(wv.x) + bf16lo(bv.x);
    y.y = (v.y - mean) * rstd * bf16hi(wv.x) + bf16hi(bv.x);
    y.z = (v.z - mean) * rstd * bf16lo(wv.y) + bf16lo(bv.y);
    y.w = (v.w - mean) * rstd * bf16hi(wv.y) + bf16hi(bv.y);
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)row * H + k) = y;  // may alias x (same thread, same k)
    if (out_bf16) {
      uint2 o;
      o.x = pack_bf16(y.x, y.y);
      o.y = pack_bf16(y.z, y.w);
      *reinterpret_cast<uint2*>(out_bf16 + (size_t)row * H + k) = o;
    }
  }
}
cudaError_t launch_layernorm(const float* x, const void* w, const void* b, float eps, void* out_bf16, float* out_f32,
                             int rows, int H, const int32_t* row_idx, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  layernorm_kernel<<<rows, 256, 0, stream>>>(x, (const __nv_bfloat16*)w, (const __nv_bfloat16*)b, eps,
                                             (__nv_bfloat16*)out_bf16, out_f32, H, row_idx);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ embedding + image splice
// reference llava_arch.py:122-163: text positions take embed_tokens rows, the sentinel position is replaced by the
// 576 projector rows of that sequence's image.  tok[row] >= 0: embedding row; else img_row[row]: image feature row.
__global__ void __launch_bounds__(256) embed_rows_kernel(const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ tok,
                                                        const int32_t* __restrict__ img_row, const float* __restrict__ img,
                                                        float* __restrict__ resid, int H) {
  const int row = blockIdx.x;
  const int t = tok[row];
  float* dst = resid + (size_t)row * H;
  if (t >= 0) {
    const __nv_bfloat16* src = table + (size_t)t * H;
    for (int k = threadIdx.x * 8; k < H; k += 256 * 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + k);
      *reinterpret_cast<float4*>(dst + k) = make_float4(bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y));
      *reinterpret_cast<float4*>(dst + k + 4) = make_float4(bf16lo(v.z), bf16hi(v.z), bf16lo(v.w), bf16hi(v.w));
    }
  } else {
    const float* src = img + (size_t)img_row[row] * H;
    for (int k = threadIdx.x * 4; k < H; k += 256 * 4)
      *reinterpret_cast<float4*>(dst + k) = *reinterpret_cast<const float4*>(src + k);
  }
}
cudaError_t launch_embed_rows(const void* table, const int32_t* tok, const int32_t* img_row, const float* img,
                              float* resid, int rows, int H, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  embed_rows_kernel<<<rows, 256, 0, stream>>>((const __nv_bfloat16*)table, tok, img_row, img, resid, H);
  return cudaGetLastError();
}

// decode step: every stream's row b consumes the same new token (vcd_sample.py:263 appends one token to all streams)
__global__ void __launch_bounds__(256) decode_embed_kernel(const __nv_bfloat16* __restrict__ table,
                                                          const int64_t* __restrict__ tokens, float* __restrict__ resid,
                                                          int B, int H) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x;
  const int64_t t = tokens[row % B];
  const __nv_bfloat16* src = table + (size_t)t * H;
  float* dst = resid + (size_t)row * H;
  for (int k = threadIdx.x * 8; k < H; k += 256 * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(src + k);
    *reinterpret_cast<float4*>(dst + k) = make_float4(bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y));
    *reinterpret_cast<float4*>(dst + k + 4) = make_float4(bf16lo(v.z), bf16hi(v.z), bf16lo(v.w), bf16hi(v.w));
  }
}
__global__ void advance_kernel(int32_t* pos, int R, int32_t* step, int32_t* epoch) {
  pdl_wait();
  pdl_launch_dependents();
  const int i = threadIdx.x + blockIdx.x * blockDim.x;
  if (pos && i < R) pos[i] += 1;
  if (step && i == 0) *step += 1;
  if (epoch && i == 0) *epoch += 1;
}
cudaError_t launch_decode_embed(const void* table, const int64_t* tokens, float* resid, int B, int S, int H,
                                cudaStream_t stream) {
  decode_embed_kernel<<<B * S, 256, 0, stream>>>((const __nv_bfloat16*)table, tokens, resid, B, H);
  return cudaGetLastError();
}
cudaError_t launch_advance(int32_t* pos, int R, int32_t* step, cudaStream_t stream, int32_t* epoch) {
  const int n = R > 1 ? R : 1;
  advance_kernel<<<(n + 127) / 128, 128, 0, stream>>>(pos, R, step, epoch);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ RoPE + KV write (prefill)
// qkv fp32 [rows, 3*Hl] -> q bf16 [rows, Hl] (roped), K/V bf16 into the head-major cache of the row's sequence.
__global__ void __launch_bounds__(128) rope_kv_prefill_kernel(const float* __restrict__ qkv, int heads,
                                                             const int32_t* __restrict__ row_seq,
                                                             const int32_t* __restrict__ row_pos,
                                                             const int64_t* __restrict__ kv_off,
                                                             const int32_t* __restrict__ kv_cap,
                                                             const float* __restrict__ rc, const float* __restrict__ rs,
                                                             __nv_bfloat16* __restrict__ q_out, __nv_bfloat16* __restrict__ kc,
                                                             __nv_bfloat16* __restrict__ vc) {
  constexpr int D = 128;
  const int row = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int Hl = heads * D;
  const int seq = row_seq[row], pos = row_pos[row];
  const int i = tid & 63;
  const float c = rc[(size_t)pos * (D / 2) + i], s = rs[(size_t)pos * (D / 2) + i];
  const float* base = qkv + (size_t)row * 3 * Hl + h * D;
  const float q1 = base[i], q2 = base[i + 64];
  const float k1 = base[Hl + i], k2 = base[Hl + i + 64];
  const float qr = tid < 64 ? rope_first(q1, q2, c, s) : rope_second(q1, q2, c, s);
  const float kr = tid < 64 ? rope_first(k1, k2, c, s) : rope_second(k1, k2, c, s);
  q_out[(size_t)row * Hl + h * D + tid] = __float2bfloat16_rn(qr);
  const size_t o = kv_off[seq] + ((size_t)h * kv_cap[seq] + pos) * D + tid;
  kc[o] = __float2bfloat16_rn(kr);
  vc[o] = __float2bfloat16_rn(base[2 * Hl + tid]);
}
cudaError_t launch_rope_kv_prefill(const float* qkv, int rows, int heads, int D, const int32_t* row_seq,
                                   const int32_t* row_pos, const int64_t* kv_off, const int32_t* kv_cap,
                                   const float* rc, const float* rs, void* q_out, void* kc, void* vc, cudaStream_t stream) {
  if (D != 128) return cudaErrorInvalidValue;
  if (rows <= 0) return cudaSuccess;
  rope_kv_prefill_kernel<<<dim3(rows, heads), 128, 0, stream>>>(qkv, heads, row_seq, row_pos, kv_off, kv_cap, rc, rs,
                                                               (__nv_bfloat16*)q_out, (__nv_bfloat16*)kc, (__nv_bfloat16*)vc);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ ViT patch extraction
// pixels bf16 [B,3,S,S] -> A bf16 [B*NP, Kpad], column c*P*P + py*P + px (the conv weight's own flattening), zero pad.
__global__ void __launch_bounds__(256) im2col_kernel(const __nv_bfloat16* __restrict__ px, __nv_bfloat16* __restrict__ out,
                                                    int S, int P, int Kpad) {
  const int G = S / P;
  const int patch = blockIdx.x;  // b*G*G + gy*G + gx
  const int b = patch / (G * G), g = patch - b * G * G, gy = g / G, gx = g - gy * G;
  const int K = 3 * P * P;
  for (int k = threadIdx.x; k < Kpad; k += 256) {
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (k < K) {
      const int c = k / (P * P), rem = k - c * P * P, py = rem / P, pxx = rem - py * P;
      v = px[(((size_t)b * 3 + c) * S + gy * P + py) * S + gx * P + pxx];
    }
    out[(size_t)patch * Kpad + k] = v;
  }
}
cudaError_t launch_im2col(const void* pixels, void* out, int B, int S, int P, int Kpad, cudaStream_t stream) {
  const int G = S / P;
  im2col_kernel<<<B * G * G, 256, 0, stream>>>((const __nv_bfloat16*)pixels, (__nv_bfloat16*)out, S, P, Kpad);
  return cudaGetLastError();
}

// CLIPVisionEmbeddings: [cls ; patches] + position_embedding  -> fp32 residual [B, NP+1, H]
__global__ void __launch_bounds__(256) vit_embed_kernel(const float* __restrict__ patch_out, const __nv_bfloat16* __restrict__ cls,
                                                       const __nv_bfloat16* __restrict__ pos, float* __restrict__ resid,
                                                       int NP, int H) {
  const int row = blockIdx.x;  // b*(NP+1) + t
  const int b = row / (NP + 1), t = row - b * (NP + 1);
  for (int k = threadIdx.x; k < H; k += 256) {
    const float base = t == 0 ? __bfloat162float(cls[k]) : patch_out[((size_t)b * NP + t - 1) * H + k];
    resid[(size_t)row * H + k] = base + __bfloat162float(pos[(size_t)t * H + k]);
  }
}
cudaError_t launch_vit_embed(const float* patch_out, const void* cls, const void* pos, float* resid, int B, int NP, int H,
                             cudaStream_t stream) {
  vit_embed_kernel<<<B * (NP + 1), 256, 0, stream>>>(patch_out, (const __nv_bfloat16*)cls, (const __nv_bfloat16*)pos, resid, NP, H);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ misc
__global__ void add_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n4) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    const float4 b = reinterpret_cast<const float4*>(src)[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
}
cudaError_t launch_add_rows(float* dst, const float* src, int64_t n, cudaStream_t stream) {
  const int64_t n4 = n / 4;
  int grid = (int)((n4 + 255) / 256);
  if (grid > 1184) grid = 1184;
  if (grid < 1) grid = 1;
  add_rows_kernel<<<grid, 256, 0, stream>>>(dst, src, n4);
  return cudaGetLastError();
}

// split-K partials of a residual GEMM, summed in split order (deterministic) and added to the residual stream
__global__ void reduce_splits_kernel(float* __restrict__ dst, const float* __restrict__ ws, int64_t n4, int splits) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    for (int z = 0; z < splits; ++z) {
      const float4 b = reinterpret_cast<const float4*>(ws)[(int64_t)z * n4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    reinterpret_cast<float4*>(dst)[i] = a;
  }
}
cudaError_t launch_reduce_splits(float* dst, const float* ws, int64_t n, int splits, cudaStream_t stream) {
  const int64_t n4 = n / 4;
  int grid = (int)((n4 + 255) / 256);
  if (grid > 1184) grid = 1184;
  if (grid < 1) grid = 1;
  reduce_splits_kernel<<<grid, 256, 0, stream>>>(dst, ws, n4, splits);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) cast_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int H,
                                                       const int32_t* __restrict__ row_idx) {
  const int row = blockIdx.x;
  const float* s = src + (size_t)(row_idx ? row_idx[row] : row) * H;
  for (int k = threadIdx.x * 4; k < H; k += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(s + k);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + (size_t)row * H + k) = o;
  }
}
cudaError_t launch_cast_rows_bf16(const float* src, void* dst, int rows, int H, const int32_t* row_idx, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  cast_rows_kernel<<<rows, 256, 0, stream>>>(src, (__nv_bfloat16*)dst, H, row_idx);
  return cudaGetLastError();
}

}  // namespace vdd
