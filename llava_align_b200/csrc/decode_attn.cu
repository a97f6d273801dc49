// Decode-step attention for all R fused rows (streams x sequences) in one launch: RoPE of the new q/k, KV-cache
// append, split-KV attention over each row's own context length, and the split merge -- no separate kernels.
//
// Replaces, per layer and per stream, HF LlamaAttention as driven by the reference at
// experiments/llava/model/language_model/llava_llama.py:91-100 (apply_rotary_pos_emb, torch.cat KV append,
// eager softmax(QK^T/sqrt(d))V).  Rows carry their own position: main stream T+575+k,
// unk stream T+k, none stream T-1+k.
//
// Grid (R*heads, max_splits), 128 threads.  Split s of row r covers tokens [s*128, min(ctx, (s+1)*128)); the
// split that contains the new token also writes it to the cache.  Partials (m, l, o[128]) go to a workspace and
// the last CTA to arrive for a (row, head) merges them (threadfence + counter), so the launch is self-contained.
// HBM bytes: ctx*2*D*2 per (row, head) -- the kernel is a pure KV stream; K/V are read as 16-byte pieces, 16
// lanes per token (256 B contiguous), 8 tokens per CTA pass.
#include "common.cuh"
#include "kernels.h"

namespace vdd {

constexpr int DA_THREADS = 128;
constexpr int DA_CHUNK = 128;
constexpr int DA_D = 128;

__global__ void __launch_bounds__(DA_THREADS) decode_attn_kernel(const DecodeAttnArgs a) {
  __shared__ float q_s[DA_D];
  __shared__ float sc[DA_CHUNK];
  __shared__ float o_s[8][DA_D];
  __shared__ float red[33];
  __shared__ int is_last;

  const int rh = blockIdx.x, split = blockIdx.y;
  const int r = rh / a.heads, h = rh - r * a.heads;
  const int tid = threadIdx.x;
  const int Hl = a.heads * DA_D;
  pdl_wait();
  pdl_launch_dependents();
  const int pos = a.pos[r];
  const int ctx = pos + 1;
  const int nsplit = (ctx + DA_CHUNK - 1) / DA_CHUNK;
  if (split >= nsplit) return;
  const int t0 = split * DA_CHUNK, t1 = min(ctx, t0 + DA_CHUNK), nt = t1 - t0;
  const int cap = a.kv_cap[r];
  __nv_bfloat16* Kc = reinterpret_cast<__nv_bfloat16*>(a.kcache) + a.kv_off[r] + (size_t)h * cap * DA_D;
  __nv_bfloat16* Vc = reinterpret_cast<__nv_bfloat16*>(a.vcache) + a.kv_off[r] + (size_t)h * cap * DA_D;
  const float* qkv = a.qkv + (size_t)r * 3 * Hl;

  // ---- RoPE (rotate-half) of q; the split holding the new token also ropes k and appends k, v to the cache
  {
    const int i = tid & 63;
    const float c = a.rope_cos[(size_t)pos * (DA_D / 2) + i], s = a.rope_sin[(size_t)pos * (DA_D / 2) + i];
    const float* qh = qkv + h * DA_D;
    const float x1 = qh[i], x2 = qh[i + 64];
    q_s[tid] = tid < 64 ? (x1 * c - x2 * s) : (x2 * c + x1 * s);
    if (split == nsplit - 1) {
      const float* kh = qkv + Hl + h * DA_D;
      const float k1 = kh[i], k2 = kh[i + 64];
      const float kr = tid < 64 ? (k1 * c - k2 * s) : (k2 * c + k1 * s);
      Kc[(size_t)pos * DA_D + tid] = __float2bfloat16_rn(kr);
      Vc[(size_t)pos * DA_D + tid] = __float2bfloat16_rn(qkv[2 * Hl + h * DA_D + tid]);
    }
  }
  __syncthreads();

  // ---- scores: 16 lanes per token, 8 tokens per pass
  const int grp = tid >> 4, gl = tid & 15;
  float qr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qr[i] = q_s[gl * 8 + i];
#pragma unroll 4
  for (int tt = grp; tt < DA_CHUNK; tt += 8) {
    float s = 0.f;
    if (tt < nt) {
      const uint4 kv = *reinterpret_cast<const uint4*>(Kc + (size_t)(t0 + tt) * DA_D + gl * 8);
      s = bf16lo(kv.x) * qr[0] + bf16hi(kv.x) * qr[1] + bf16lo(kv.y) * qr[2] + bf16hi(kv.y) * qr[3] +
          bf16lo(kv.z) * qr[4] + bf16hi(kv.z) * qr[5] + bf16lo(kv.w) * qr[6] + bf16hi(kv.w) * qr[7];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (gl == 0) sc[tt] = tt < nt ? s * a.scale : -INFINITY;
  }
  __syncthreads();

  // ---- chunk softmax statistics
  const float sv = sc[tid];
  const float m = block_reduce<true>(sv, red);
  const float pv = expf(sv - m);  // 0 for the padding
  const float l = block_reduce<false>(pv, red);
  sc[tid] = pv;
  __syncthreads();

  // ---- P.V: lane holds 8 dims, 8 token slots
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll 4
  for (int tt = grp; tt < DA_CHUNK; tt += 8) {
    if (tt < nt) {
      const uint4 vv = *reinterpret_cast<const uint4*>(Vc + (size_t)(t0 + tt) * DA_D + gl * 8);
      const float p = sc[tt];
      o[0] = fmaf(p, bf16lo(vv.x), o[0]); o[1] = fmaf(p, bf16hi(vv.x), o[1]);
      o[2] = fmaf(p, bf16lo(vv.y), o[2]); o[3] = fmaf(p, bf16hi(vv.y), o[3]);
      o[4] = fmaf(p, bf16lo(vv.z), o[4]); o[5] = fmaf(p, bf16hi(vv.z), o[5]);
      o[6] = fmaf(p, bf16lo(vv.w), o[6]); o[7] = fmaf(p, bf16hi(vv.w), o[7]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) o_s[grp][gl * 8 + i] = o[i];
  __syncthreads();
  float od = 0.f;
#pragma unroll
  for (int g = 0; g < 8; ++g) od += o_s[g][tid];

  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out_bf16) + (size_t)r * Hl + h * DA_D;
  if (nsplit == 1) {
    out[tid] = __float2bfloat16_rn(od / l);
    return;
  }
  // ---- park the partial, last arriver merges (fixed order over splits => deterministic)
  const size_t slot = (size_t)rh * a.max_splits + split;
  a.part_o[slot * DA_D + tid] = od;
  if (tid == 0) { a.part_ml[slot * 2] = m; a.part_ml[slot * 2 + 1] = l; }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(&a.counters[rh], 1);
    is_last = prev == nsplit - 1;
    if (is_last) a.counters[rh] = 0;  // ready for the next launch
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, __ldcg(&a.part_ml[((size_t)rh * a.max_splits + s) * 2]));
  float L = 0.f, acc = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const size_t sl = (size_t)rh * a.max_splits + s;
    const float w = expf(__ldcg(&a.part_ml[sl * 2]) - M);
    L += __ldcg(&a.part_ml[sl * 2 + 1]) * w;
    acc += __ldcg(&a.part_o[sl * DA_D + tid]) * w;
  }
  out[tid] = __float2bfloat16_rn(acc / L);
}

cudaError_t launch_decode_attn(const DecodeAttnArgs& a, cudaStream_t stream) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.R * a.heads, a.max_splits);
  cfg.blockDim = dim3(DA_THREADS);
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, decode_attn_kernel, a);
}

}  // namespace vdd
