// Varlen multi-head attention for the GEMM-shaped phases: causal over each stream's prompt (LLM prefill, head_dim
// 128, keys read from the head-major KV cache just written) and non-causal over the 577 ViT tokens (head_dim 64).
//
// Replaces HF eager attention (softmax(QK^T*scale + mask) V with fp32 softmax) as reached from the reference at
// llava_llama.py:91-100 (LLM, step 0 of vcd_sample.py:109/163/178) and clip_encoder.py:48 (vision tower).
//
// Flash-style: one CTA = 64 queries of one (sequence, head); 4 warps x 16 query rows; K/V tiles of 64 keys staged in
// padded shared memory, S = QK^T and O += PV on mma.sync m16n8k16 (bf16 in, fp32 accumulate), online softmax in the
// exp2 domain.  Attention is <2% of the prefill FLOPs, so this kernel stays on the legacy MMA path in
// round 1; the GEMMs that dominate are tcgen05 (gemm_tc.cu).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vdd {

constexpr int PA_BQ = 64, PA_BK = 64, PA_THREADS = 128;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int D, bool CAUSAL>
__global__ void __launch_bounds__(PA_THREADS) prefill_attn_kernel(const PrefillAttnArgs a) {
  constexpr int LDS = D + 8;  // padded row (bf16 elements): 16-byte rows land in distinct bank groups
  extern __shared__ __align__(16) unsigned char smem_pa[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_pa);
  __nv_bfloat16* sK = sQ + PA_BQ * LDS;
  __nv_bfloat16* sV = sK + PA_BK * LDS;

  const int seq = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
  const int len = a.seqlen[seq];
  const int q0 = qt * PA_BQ;
  if (q0 >= len) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  const __nv_bfloat16* Q = reinterpret_cast<const __nv_bfloat16*>(a.q) + (size_t)a.q_row0[seq] * a.q_ld + h * D;
  const __nv_bfloat16* Kp = reinterpret_cast<const __nv_bfloat16*>(a.k) + a.k_off[seq] + (size_t)h * a.k_head_stride[seq];
  const __nv_bfloat16* Vp = reinterpret_cast<const __nv_bfloat16*>(a.v) + a.k_off[seq] + (size_t)h * a.k_head_stride[seq];

  // ---- stage Q
  constexpr int CPR = D / 8;  // 16-byte chunks per row
  for (int i = tid; i < PA_BQ * CPR; i += PA_THREADS) {
    const int r = i / CPR, c = i - r * CPR;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (q0 + r < len) v = *reinterpret_cast<const uint4*>(Q + (size_t)(q0 + r) * a.q_ld + c * 8);
    *reinterpret_cast<uint4*>(sQ + r * LDS + c * 8) = v;
  }
  __syncthreads();
  uint32_t qf[D / 16][4];
  {
    const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int k = 0; k < D / 16; ++k) ldsm_x4(qf[k], smem_u32(sQ + row * LDS + k * 16 + (lane >> 4) * 8));
  }

  float o[D / 8][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const float sl2 = a.scale * 1.4426950408889634f;
  const int qrow0 = q0 + warp * 16 + g;  // this thread's rows: qrow0 and qrow0 + 8

  const int kend = CAUSAL ? min(len, q0 + PA_BQ) : len;
  for (int k0 = 0; k0 < kend; k0 += PA_BK) {
    __syncthreads();  // previous tile fully consumed
    for (int i = tid; i < PA_BK * CPR; i += PA_THREADS) {
      const int r = i / CPR, c = i - r * CPR;
      uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
      if (k0 + r < len) {
        kv = *reinterpret_cast<const uint4*>(Kp + (size_t)(k0 + r) * a.k_tok_stride + c * 8);
        vv = *reinterpret_cast<const uint4*>(Vp + (size_t)(k0 + r) * a.k_tok_stride + c * 8);
      }
      *reinterpret_cast<uint4*>(sK + r * LDS + c * 8) = kv;
      *reinterpret_cast<uint4*>(sV + r * LDS + c * 8) = vv;
    }
    __syncthreads();

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[PA_BK / 8][4];
#pragma unroll
    for (int j = 0; j < PA_BK / 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int j = 0; j < PA_BK / 8; ++j) {
#pragma unroll
      for (int k = 0; k < D / 16; k += 2) {
        uint32_t bf[4];
        ldsm_x4(bf, smem_u32(sK + (j * 8 + (lane & 7)) * LDS + k * 16 + (lane >> 3) * 8));
        mma_bf16_16816(s[j], qf[k], bf[0], bf[1]);
        mma_bf16_16816(s[j], qf[k + 1], bf[2], bf[3]);
      }
    }
    // ---- mask + online softmax (exp2 domain)
    float mnew[2] = {mrow[0], mrow[1]};
#pragma unroll
    for (int j = 0; j < PA_BK / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = k0 + j * 8 + 2 * t + (e & 1);
        const int qr = qrow0 + (e >> 1) * 8;
        const bool ok = key < len && (!CAUSAL || key <= qr);
        s[j][e] = ok ? s[j][e] * sl2 : -INFINITY;
        mnew[e >> 1] = fmaxf(mnew[e >> 1], s[j][e]);
      }
    }
    float corr[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mnew[i] = fmaxf(mnew[i], __shfl_xor_sync(0xffffffffu, mnew[i], 1));
      mnew[i] = fmaxf(mnew[i], __shfl_xor_sync(0xffffffffu, mnew[i], 2));
      corr[i] = mnew[i] == -INFINITY ? 1.f : exp2f(mrow[i] - mnew[i]);
      mrow[i] = mnew[i];
      lrow[i] *= corr[i];
    }
    const float mu0 = mrow[0] == -INFINITY ? 0.f : mrow[0], mu1 = mrow[1] == -INFINITY ? 0.f : mrow[1];
#pragma unroll
    for (int j = 0; j < PA_BK / 8; ++j) {
      s[j][0] = exp2f(s[j][0] - mu0); s[j][1] = exp2f(s[j][1] - mu0);
      s[j][2] = exp2f(s[j][2] - mu1); s[j][3] = exp2f(s[j][3] - mu1);
      lrow[0] += s[j][0] + s[j][1];
      lrow[1] += s[j][2] + s[j][3];
    }
#pragma unroll
    for (int j = 0; j < D / 8; ++j) { o[j][0] *= corr[0]; o[j][1] *= corr[0]; o[j][2] *= corr[1]; o[j][3] *= corr[1]; }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < PA_BK / 16; ++kk) {
      uint32_t pf[4];
      pf[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int jd = 0; jd < D / 8; jd += 2) {
        uint32_t vf[4];
        ldsm_x4_t(vf, smem_u32(sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + (jd + (lane >> 4)) * 8));
        mma_bf16_16816(o[jd], pf, vf[0], vf[1]);
        mma_bf16_16816(o[jd + 1], pf, vf[2], vf[3]);
      }
    }
  }
  // ---- finalize: quad-reduce the row sums, normalise, store bf16
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    lrow[i] += __shfl_xor_sync(0xffffffffu, lrow[i], 1);
    lrow[i] += __shfl_xor_sync(0xffffffffu, lrow[i], 2);
  }
  __nv_bfloat16* O = reinterpret_cast<__nv_bfloat16*>(a.out) + (size_t)a.q_row0[seq] * a.out_ld + h * D;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int qr = qrow0 + i * 8;
    if (qr < len) {
      const float inv = 1.f / lrow[i];
#pragma unroll
      for (int j = 0; j < D / 8; ++j)
        *reinterpret_cast<uint32_t*>(O + (size_t)qr * a.out_ld + j * 8 + 2 * t) = pack_bf16(o[j][2 * i] * inv, o[j][2 * i + 1] * inv);
    }
  }
}

template <int D, bool CAUSAL>
static cudaError_t launch_pa(const PrefillAttnArgs& a, cudaStream_t stream) {
  const size_t smem = (size_t)(PA_BQ + 2 * PA_BK) * (D + 8) * 2;
  auto kern = prefill_attn_kernel<D, CAUSAL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((a.max_seqlen + PA_BQ - 1) / PA_BQ, a.heads, a.nseq);
  kern<<<grid, PA_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_prefill_attn(const PrefillAttnArgs& a, cudaStream_t stream) {
  static const bool legacy = getenv("VDD_PREFILL_ATTN_MMA") != nullptr;
  if (!legacy && prefill_attn_tc_supported(a)) return launch_prefill_attn_tc(a, stream);
  return launch_prefill_attn_mma(a, stream);
}

cudaError_t launch_prefill_attn_mma(const PrefillAttnArgs& a, cudaStream_t stream) {
  if (a.nseq <= 0 || a.max_seqlen <= 0) return cudaSuccess;
  if (a.head_dim == 128) return a.causal ? launch_pa<128, true>(a, stream) : launch_pa<128, false>(a, stream);
  if (a.head_dim == 64) return a.causal ? launch_pa<64, true>(a, stream) : launch_pa<64, false>(a, stream);
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------ attention maps
// The `attentions[0][-1]` the reference's post-hoc-debias drivers read from generate(output_attentions=True)
// (llava_calibrate.py:180-182, vcd_sample.py:241-246): the LAST layer's softmax(q k^T / sqrt(d)) of the main stream at
// step 0, [B, heads, len, len] fp32 with zeros above the diagonal (HF eager attention returns the post-softmax weights).
// Computed after the prefill from what it leaves behind -- the roped bf16 queries of the last layer and that layer's K
// cache -- so the fused prefill attention itself never materialises a len x len matrix.  One CTA per (query, head, seq).
constexpr int AP_THREADS = 128;
__global__ void __launch_bounds__(AP_THREADS) attn_probs_kernel(const AttnProbsArgs a) {
  extern __shared__ float ap_sc[];  // [len] scores of this query
  __shared__ float q_s[128];
  __shared__ float red[33];
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(a.q) + (size_t)(a.q_row0[b] + i) * a.q_ld + h * 128;
  const __nv_bfloat16* K = reinterpret_cast<const __nv_bfloat16*>(a.k) + a.k_off[b] + (size_t)h * a.kv_cap[b] * 128;
  q_s[tid] = __bfloat162float(q[tid]);
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j <= i; j += AP_THREADS) {
    const uint4* kr = reinterpret_cast<const uint4*>(K + (size_t)j * 128);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint4 w = kr[c];
      const float* qq = q_s + c * 8;
      s += bf16lo(w.x) * qq[0] + bf16hi(w.x) * qq[1] + bf16lo(w.y) * qq[2] + bf16hi(w.y) * qq[3] + bf16lo(w.z) * qq[4] +
           bf16hi(w.z) * qq[5] + bf16lo(w.w) * qq[6] + bf16hi(w.w) * qq[7];
    }
    s *= a.scale;
    ap_sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_reduce<true>(mx, red);
  float sum = 0.f;
  for (int j = tid; j <= i; j += AP_THREADS) {
    const float e = expf(ap_sc[j] - mx);
    ap_sc[j] = e;
    sum += e;
  }
  sum = block_reduce<false>(sum, red);
  const float inv = 1.f / sum;
  float* o = a.out + (((size_t)b * a.heads + h) * a.len + i) * a.len;
  for (int j = tid; j < a.len; j += AP_THREADS) o[j] = j <= i ? ap_sc[j] * inv : 0.f;
}

// The same for a decode step (`attentions[step][layer]`, step >= 1: [B, heads, 1, len + step]): the query of the new token is roped
// here from the step's fp32 qkv row; the layer's cache already holds the new key (decode_attn_kernel ran).  One CTA per (head, row).
__global__ void __launch_bounds__(AP_THREADS) attn_probs_decode_kernel(const AttnProbsDecodeArgs a) {
  extern __shared__ float ap_sc[];  // [W]
  __shared__ float q_s[128];
  __shared__ float red[33];
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int pos = a.pos[b], ctx = pos + 1;
  {
    const float* qh = a.qkv + (size_t)b * 3 * a.heads * 128 + h * 128;
    const int i = tid & 63;
    const float c = a.rope_cos[(size_t)pos * 64 + i], s = a.rope_sin[(size_t)pos * 64 + i];
    const float x1 = qh[i], x2 = qh[i + 64];
    q_s[tid] = tid < 64 ? (x1 * c - x2 * s) : (x2 * c + x1 * s);
  }
  __syncthreads();
  const __nv_bfloat16* K = reinterpret_cast<const __nv_bfloat16*>(a.k) + a.k_off[b] + (size_t)h * a.kv_cap[b] * 128;
  float mx = -INFINITY;
  for (int j = tid; j < ctx; j += AP_THREADS) {
    const uint4* kr = reinterpret_cast<const uint4*>(K + (size_t)j * 128);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint4 w = kr[c];
      const float* qq = q_s + c * 8;
      s += bf16lo(w.x) * qq[0] + bf16hi(w.x) * qq[1] + bf16lo(w.y) * qq[2] + bf16hi(w.y) * qq[3] + bf16lo(w.z) * qq[4] +
           bf16hi(w.z) * qq[5] + bf16lo(w.w) * qq[6] + bf16hi(w.w) * qq[7];
    }
    s *= a.scale;
    ap_sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_reduce<true>(mx, red);
  float sum = 0.f;
  for (int j = tid; j < ctx; j += AP_THREADS) {
    const float e = expf(ap_sc[j] - mx);
    ap_sc[j] = e;
    sum += e;
  }
  sum = block_reduce<false>(sum, red);
  const float inv = 1.f / sum;
  float* o = a.out + ((((size_t)(*a.step - 1) * a.layers + a.layer) * a.B + b) * a.heads + h) * a.W;
  for (int j = tid; j < a.W; j += AP_THREADS) o[j] = j < ctx ? ap_sc[j] * inv : 0.f;
}

cudaError_t launch_attn_probs_decode(const AttnProbsDecodeArgs& a, cudaStream_t stream) {
  if (a.B <= 0 || a.W <= 0) return cudaSuccess;
  const size_t smem = (size_t)a.W * 4;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(attn_probs_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  attn_probs_decode_kernel<<<dim3(a.heads, a.B), AP_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_attn_probs(const AttnProbsArgs& a, cudaStream_t stream) {
  if (a.B <= 0 || a.len <= 0) return cudaSuccess;
  const size_t smem = (size_t)a.len * 4;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(attn_probs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  attn_probs_kernel<<<dim3(a.len, a.heads, a.B), AP_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace vdd
