// Decode-step attention for all R fused rows (streams x sequences) in one launch: RoPE of the new q/k, KV-cache
// append, split-KV attention over each row's own context length, and the split merge -- no separate kernels.
//
// Replaces, per layer and per stream, HF LlamaAttention as driven by the reference at
// experiments/llava/model/language_model/llava_llama.py:91-100 (apply_rotary_pos_emb, torch.cat KV append,
// eager softmax(QK^T/sqrt(d))V).  Rows carry their own position: main stream T+575+k,
// unk stream T+k, none stream T-1+k.
//
// Grid (R*heads, max_splits), 128 threads.  Split s of row r covers tokens [s*64, min(ctx, (s+1)*64)); the
// split that contains the new token also writes it to the cache.  Partials (m, l, o[128]) go to a workspace and
// the last CTA to arrive for a (row, head) merges them (threadfence + counter), so the launch is self-contained.
// HBM bytes: ctx*2*D*2 per (row, head) -- the kernel is a pure KV stream; K/V are read as 16-byte pieces, 16
// lanes per token (256 B contiguous); all 16 loads of a lane are issued before any is consumed.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vdd {

constexpr int DA_THREADS = 128;
constexpr int DA_CHUNK = 64;   // tokens per split: 8 groups of 16 lanes x 8 tokens each
constexpr int DA_D = 128;
constexpr int DA_TPG = DA_CHUNK / 8;  // tokens per 16-lane group

__global__ void __launch_bounds__(DA_THREADS) decode_attn_kernel(const DecodeAttnArgs a) {
  __shared__ float q_s[DA_D];
  __shared__ __align__(16) __nv_bfloat16 knew_s[DA_D];
  __shared__ __align__(16) __nv_bfloat16 vnew_s[DA_D];
  __shared__ float o_s[8][DA_D];
  __shared__ float ml_s[8][2];
  __shared__ int is_last;

  const int rh = blockIdx.x, split = blockIdx.y;
  const int r = rh / a.heads, h = rh - r * a.heads;
  const int tid = threadIdx.x;
  const int Hl = a.heads * DA_D;
  const int grp = tid >> 4, gl = tid & 15;
  const int cap = a.kv_cap[r];  // written at prefill, constant during decoding: safe before the PDL wait
  __nv_bfloat16* Kc = reinterpret_cast<__nv_bfloat16*>(a.kcache) + a.kv_off[r] + (size_t)h * cap * DA_D;
  __nv_bfloat16* Vc = reinterpret_cast<__nv_bfloat16*>(a.vcache) + a.kv_off[r] + (size_t)h * cap * DA_D;
  // pos and the cached tokens < pos were written by kernels that finished before the immediate predecessor (the QKV
  // projection) could even start, so they may be read before the programmatic-dependency wait: the KV stream of this
  // kernel overlaps the tail of the projection.
  const int pos = a.pos[r];
  const int ctx = pos + 1;
  const int nsplit = (ctx + DA_CHUNK - 1) / DA_CHUNK;
  if (split >= nsplit) return;
  const int t0 = split * DA_CHUNK;

  // ---- issue every K/V load of this split up front (16 x 16 B per lane in flight); the new token comes from registers
  uint4 kr[DA_TPG], vr[DA_TPG];
#pragma unroll
  for (int j = 0; j < DA_TPG; ++j) {
    const int t = t0 + grp + 8 * j;
    kr[j] = make_uint4(0u, 0u, 0u, 0u);
    vr[j] = kr[j];
    if (t < pos) {
      kr[j] = *reinterpret_cast<const uint4*>(Kc + (size_t)t * DA_D + gl * 8);
      vr[j] = *reinterpret_cast<const uint4*>(Vc + (size_t)t * DA_D + gl * 8);
    }
  }
  pdl_wait();
  pdl_launch_dependents();
  // ---- RoPE (rotate-half) of q; the split holding the new token also ropes k and appends k, v to the cache
  {
    const float* qkv = a.qkv + (size_t)r * 3 * Hl;
    const int i = tid & 63;
    const float c = a.rope_cos[(size_t)pos * (DA_D / 2) + i], s = a.rope_sin[(size_t)pos * (DA_D / 2) + i];
    const float* qh = qkv + h * DA_D;
    const float x1 = qh[i], x2 = qh[i + 64];
    q_s[tid] = tid < 64 ? (x1 * c - x2 * s) : (x2 * c + x1 * s);
    if (split == nsplit - 1) {
      const float* kh = qkv + Hl + h * DA_D;
      const float k1 = kh[i], k2 = kh[i + 64];
      const __nv_bfloat16 kb = __float2bfloat16_rn(tid < 64 ? (k1 * c - k2 * s) : (k2 * c + k1 * s));
      const __nv_bfloat16 vb = __float2bfloat16_rn(qkv[2 * Hl + h * DA_D + tid]);
      Kc[(size_t)pos * DA_D + tid] = kb;
      Vc[(size_t)pos * DA_D + tid] = vb;
      knew_s[tid] = kb;
      vnew_s[tid] = vb;
    }
  }
  __syncthreads();
  float qr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qr[i] = q_s[gl * 8 + i];

  // ---- per-group scores, softmax statistics and P.V entirely in registers
  float sc[DA_TPG];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < DA_TPG; ++j) {
    const int t = t0 + grp + 8 * j;
    if (t == pos) {
      kr[j] = *reinterpret_cast<const uint4*>(knew_s + gl * 8);
      vr[j] = *reinterpret_cast<const uint4*>(vnew_s + gl * 8);
    }
    float s = bf16lo(kr[j].x) * qr[0] + bf16hi(kr[j].x) * qr[1] + bf16lo(kr[j].y) * qr[2] + bf16hi(kr[j].y) * qr[3] +
              bf16lo(kr[j].z) * qr[4] + bf16hi(kr[j].z) * qr[5] + bf16lo(kr[j].w) * qr[6] + bf16hi(kr[j].w) * qr[7];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    sc[j] = t <= pos ? s * a.scale : -INFINITY;
    m = fmaxf(m, sc[j]);
  }
  float l = 0.f, o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
  for (int j = 0; j < DA_TPG; ++j) {
    const float p = m == -INFINITY ? 0.f : expf(sc[j] - m);
    l += p;
    o[0] = fmaf(p, bf16lo(vr[j].x), o[0]); o[1] = fmaf(p, bf16hi(vr[j].x), o[1]);
    o[2] = fmaf(p, bf16lo(vr[j].y), o[2]); o[3] = fmaf(p, bf16hi(vr[j].y), o[3]);
    o[4] = fmaf(p, bf16lo(vr[j].z), o[4]); o[5] = fmaf(p, bf16hi(vr[j].z), o[5]);
    o[6] = fmaf(p, bf16lo(vr[j].w), o[6]); o[7] = fmaf(p, bf16hi(vr[j].w), o[7]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) o_s[grp][gl * 8 + i] = o[i];
  if (gl == 0) { ml_s[grp][0] = m; ml_s[grp][1] = l; }
  __syncthreads();
  // ---- merge the 8 groups (fixed order): thread = output dim
  float M = -INFINITY;
#pragma unroll
  for (int g = 0; g < 8; ++g) M = fmaxf(M, ml_s[g][0]);
  float L = 0.f, od = 0.f;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float w = ml_s[g][0] == -INFINITY ? 0.f : expf(ml_s[g][0] - M);
    L += ml_s[g][1] * w;
    od += o_s[g][tid] * w;
  }

  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out_bf16) + (size_t)r * Hl + h * DA_D;
  if (nsplit == 1) {
    out[tid] = __float2bfloat16_rn(od / L);
    return;
  }
  // ---- park the partial, last arriver merges (fixed order over splits => deterministic)
  const size_t slot = (size_t)rh * a.max_splits + split;
  a.part_o[slot * DA_D + tid] = od;
  if (tid == 0) { a.part_ml[slot * 2] = M; a.part_ml[slot * 2 + 1] = L; }
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
  float MM = -INFINITY;
  for (int s = 0; s < nsplit; ++s) MM = fmaxf(MM, __ldcg(&a.part_ml[((size_t)rh * a.max_splits + s) * 2]));
  float LL = 0.f, acc = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const size_t sl = (size_t)rh * a.max_splits + s;
    const float w = expf(__ldcg(&a.part_ml[sl * 2]) - MM);
    LL += __ldcg(&a.part_ml[sl * 2 + 1]) * w;
    acc += __ldcg(&a.part_o[sl * DA_D + tid]) * w;
  }
  out[tid] = __float2bfloat16_rn(acc / LL);
}

// ------------------------------------------------------------------------------------------------ cluster variant
// One thread-block CLUSTER of 8 CTAs per (row, head): every CTA takes one eighth of the context, and the eight partial
// (m, l, o[128]) are merged through distributed shared memory after one cluster barrier -- no global partials, no
// fence + atomic + "last CTA" round trips, no idle CTAs for the short debias rows.  Same per-token arithmetic as above.
namespace cg = cooperative_groups;
constexpr int DC_CL = 8;
constexpr int DC_TPG = 8;   // tokens per 16-lane group per pass => 64 tokens per CTA per pass (registers: 6 CTAs per SM, one wave)

__global__ void __cluster_dims__(1, DC_CL, 1) __launch_bounds__(DA_THREADS, 6) decode_attn_cluster_kernel(const DecodeAttnArgs a) {
  __shared__ float q_s[DA_D];
  __shared__ __align__(16) __nv_bfloat16 knew_s[DA_D];
  __shared__ __align__(16) __nv_bfloat16 vnew_s[DA_D];
  __shared__ float o_s[8][DA_D];
  __shared__ float ml_s[8][2];
  __shared__ float pub_o[DA_D];  // this CTA's partial, read by its cluster peers
  __shared__ float pub_ml[2];
  cg::cluster_group cl = cg::this_cluster();
  const int rh = blockIdx.x, rank = blockIdx.y;
  const int r = rh / a.heads, h = rh - r * a.heads;
  const int tid = threadIdx.x;
  const int Hl = a.heads * DA_D;
  const int grp = tid >> 4, gl = tid & 15;
  const int cap = a.kv_cap[r];
  __nv_bfloat16* Kc = reinterpret_cast<__nv_bfloat16*>(a.kcache) + a.kv_off[r] + (size_t)h * cap * DA_D;
  __nv_bfloat16* Vc = reinterpret_cast<__nv_bfloat16*>(a.vcache) + a.kv_off[r] + (size_t)h * cap * DA_D;
  // pos and the cached tokens < pos are older than the immediate predecessor kernel: safe before the dependency wait
  const int pos = a.pos[r];
  const int ctx = pos + 1;
  const int span = ((ctx + DC_CL - 1) / DC_CL + 7) & ~7;  // tokens per CTA, a multiple of the 8 groups
  const int t_begin = rank * span, t_end = min(ctx, t_begin + span);
  const bool owns_new = pos >= t_begin && pos < t_end;

  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  bool waited = false;
  float qr[8];
  for (int p0 = t_begin; p0 < t_end || !waited; p0 += 8 * DC_TPG) {
    uint4 kr[DC_TPG], vr[DC_TPG];
#pragma unroll
    for (int j = 0; j < DC_TPG; ++j) {
      const int t = p0 + grp + 8 * j;
      kr[j] = make_uint4(0u, 0u, 0u, 0u);
      vr[j] = kr[j];
      if (t < t_end && t < pos) {
        kr[j] = *reinterpret_cast<const uint4*>(Kc + (size_t)t * DA_D + gl * 8);
        vr[j] = *reinterpret_cast<const uint4*>(Vc + (size_t)t * DA_D + gl * 8);
      }
    }
    if (!waited) {
      waited = true;
      pdl_wait();
      pdl_launch_dependents();
      const float* qkv = a.qkv + (size_t)r * 3 * Hl;
      const int i = tid & 63;
      const float c = a.rope_cos[(size_t)pos * (DA_D / 2) + i], s = a.rope_sin[(size_t)pos * (DA_D / 2) + i];
      const float* qh = qkv + h * DA_D;
      const float x1 = qh[i], x2 = qh[i + 64];
      q_s[tid] = tid < 64 ? (x1 * c - x2 * s) : (x2 * c + x1 * s);
      if (owns_new) {
        const float* kh = qkv + Hl + h * DA_D;
        const float k1 = kh[i], k2 = kh[i + 64];
        const __nv_bfloat16 kb = __float2bfloat16_rn(tid < 64 ? (k1 * c - k2 * s) : (k2 * c + k1 * s));
        const __nv_bfloat16 vb = __float2bfloat16_rn(qkv[2 * Hl + h * DA_D + tid]);
        Kc[(size_t)pos * DA_D + tid] = kb;
        Vc[(size_t)pos * DA_D + tid] = vb;
        knew_s[tid] = kb;
        vnew_s[tid] = vb;
      }
      __syncthreads();
#pragma unroll
      for (int i2 = 0; i2 < 8; ++i2) qr[i2] = q_s[gl * 8 + i2];
    }
    float sc[DC_TPG];
    float mp = -INFINITY;
#pragma unroll
    for (int j = 0; j < DC_TPG; ++j) {
      const int t = p0 + grp + 8 * j;
      if (t == pos && owns_new) {
        kr[j] = *reinterpret_cast<const uint4*>(knew_s + gl * 8);
        vr[j] = *reinterpret_cast<const uint4*>(vnew_s + gl * 8);
      }
      float s = bf16lo(kr[j].x) * qr[0] + bf16hi(kr[j].x) * qr[1] + bf16lo(kr[j].y) * qr[2] + bf16hi(kr[j].y) * qr[3] +
                bf16lo(kr[j].z) * qr[4] + bf16hi(kr[j].z) * qr[5] + bf16lo(kr[j].w) * qr[6] + bf16hi(kr[j].w) * qr[7];
#pragma unroll
      for (int x = 8; x > 0; x >>= 1) s += __shfl_xor_sync(0xffffffffu, s, x);
      sc[j] = t < t_end ? s * a.scale : -INFINITY;
      mp = fmaxf(mp, sc[j]);
    }
    const float mn = fmaxf(m, mp);
    const float corr = mn == -INFINITY ? 1.f : expf(m - mn);  // exp(-inf) = 0 on the first pass
    l *= corr;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] *= corr;
    m = mn;
#pragma unroll
    for (int j = 0; j < DC_TPG; ++j) {
      const float p = m == -INFINITY ? 0.f : expf(sc[j] - m);
      l += p;
      o[0] = fmaf(p, bf16lo(vr[j].x), o[0]); o[1] = fmaf(p, bf16hi(vr[j].x), o[1]);
      o[2] = fmaf(p, bf16lo(vr[j].y), o[2]); o[3] = fmaf(p, bf16hi(vr[j].y), o[3]);
      o[4] = fmaf(p, bf16lo(vr[j].z), o[4]); o[5] = fmaf(p, bf16hi(vr[j].z), o[5]);
      o[6] = fmaf(p, bf16lo(vr[j].w), o[6]); o[7] = fmaf(p, bf16hi(vr[j].w), o[7]);
    }
  }
  // ---- merge the 8 groups of this CTA (fixed order), publish, merge the 8 CTAs of the cluster
#pragma unroll
  for (int i = 0; i < 8; ++i) o_s[grp][gl * 8 + i] = o[i];
  if (gl == 0) { ml_s[grp][0] = m; ml_s[grp][1] = l; }
  __syncthreads();
  {
    float M = -INFINITY;
#pragma unroll
    for (int g = 0; g < 8; ++g) M = fmaxf(M, ml_s[g][0]);
    float L = 0.f, od = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float w = ml_s[g][0] == -INFINITY ? 0.f : expf(ml_s[g][0] - M);
      L += ml_s[g][1] * w;
      od += o_s[g][tid] * w;
    }
    pub_o[tid] = od;
    if (tid == 0) { pub_ml[0] = M; pub_ml[1] = L; }
  }
  cl.sync();
  if (tid < DA_D / DC_CL) {  // this CTA finishes 16 of the 128 output dims
    const int dim = rank * (DA_D / DC_CL) + tid;
    float mk[DC_CL], MM = -INFINITY;
#pragma unroll
    for (int k = 0; k < DC_CL; ++k) {
      mk[k] = cl.map_shared_rank(pub_ml, k)[0];
      MM = fmaxf(MM, mk[k]);
    }
    float LL = 0.f, acc = 0.f;
#pragma unroll
    for (int k = 0; k < DC_CL; ++k) {
      const float w = mk[k] == -INFINITY ? 0.f : expf(mk[k] - MM);
      LL += cl.map_shared_rank(pub_ml, k)[1] * w;
      acc += cl.map_shared_rank(pub_o, k)[dim] * w;
    }
    reinterpret_cast<__nv_bfloat16*>(a.out_bf16)[(size_t)r * Hl + h * DA_D + dim] = __float2bfloat16_rn(acc / LL);
  }
  cl.sync();  // nobody leaves while a peer may still read its shared memory
}

// ------------------------------------------------------------------------------------------------ streaming variant
// Many fused rows (R > 8: BASELINE configs 3-5, 96-192 rows whose KV outweighs the weights).  The 64-token kernel above
// launches R*heads*max_splits CTAs, most of which belong to short debias rows and exit at once, and each live CTA moves
// only 32 KB.  Here a CTA owns up to `chunk` (a multiple of 64, default 256) tokens of ONE (row, head) taken from a work
// table built at prefill (row, split) -- rows only get the splits their capacity can ever need -- and streams them
// through a 3-stage shared-memory ring filled by 1-D bulk copies (cp.async.bulk + mbarrier; a head's tokens are
// contiguous, 64 tokens = 16 KB of K and 16 KB of V per stage), so ~190 KB per SM are in flight with no register
// staging.  Per-token arithmetic, the group merge and the split merge are those of decode_attn_kernel.
constexpr int DS_TOK = 64;                         // tokens per stage
constexpr int DS_STAGE_BYTES = DS_TOK * DA_D * 2;  // 16 KB (K) and the same again for V
constexpr int ds_smem_bytes(int stages) { return stages * 2 * DS_STAGE_BYTES + 128; }

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// DS_STAGES = 3: 96 KB ring, two CTAs per SM; DS_STAGES = 2: 64 KB ring, three CTAs per SM (more CTAs to overlap each other's
// start-up and merge latencies; same bytes in flight per SM).
template <int DS_STAGES>
__global__ void __launch_bounds__(DA_THREADS, DS_STAGES == 2 ? 3 : 2) decode_attn_stream_kernel(const DecodeAttnArgs a) {
  extern __shared__ __align__(128) unsigned char ds_smem[];
  __shared__ float q_s[DA_D];
  __shared__ __align__(16) __nv_bfloat16 knew_s[DA_D];
  __shared__ __align__(16) __nv_bfloat16 vnew_s[DA_D];
  __shared__ float o_s[8][DA_D];
  __shared__ float ml_s[8][2];
  __shared__ int is_last;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ds_smem + DS_STAGES * 2 * DS_STAGE_BYTES);

  const int h = blockIdx.x, item = blockIdx.y;
  const int r = a.item_row[item], split = a.item_split[item];
  const int rh = r * a.heads + h;
  const int tid = threadIdx.x;
  const int Hl = a.heads * DA_D;
  const int grp = tid >> 4, gl = tid & 15;
  const int cap = a.kv_cap[r];
  __nv_bfloat16* Kc = reinterpret_cast<__nv_bfloat16*>(a.kcache) + a.kv_off[r] + (size_t)h * cap * DA_D;
  __nv_bfloat16* Vc = reinterpret_cast<__nv_bfloat16*>(a.vcache) + a.kv_off[r] + (size_t)h * cap * DA_D;
  const int pos = a.pos[r];  // older than the immediate predecessor kernel (see decode_attn_kernel)
  const int ctx = pos + 1;
  const int t0 = split * a.chunk;
  if (t0 >= ctx) return;  // this row has not grown into this split yet
  const int nsplit = (ctx + a.chunk - 1) / a.chunk;
  const int t_old_end = min(pos, t0 + a.chunk);             // cached tokens of this split: [t0, t_old_end)
  const bool owns_new = pos < t0 + a.chunk;                 // the split that contains the new token
  const int nst = (t_old_end - t0 + DS_TOK - 1) / DS_TOK;   // ring stages to stream (0 when the split holds only the new token)

  if (tid == 0) {
    for (int s = 0; s < DS_STAGES; ++s) mbar_init(&full_bar[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int c) {  // thread 0: stage c of this split
    const int s = c % DS_STAGES;
    const int tb = t0 + c * DS_TOK;
    const uint32_t bytes = (uint32_t)(min(DS_TOK, t_old_end - tb) * DA_D * 2);
    mbar_arrive_expect_tx(&full_bar[s], 2 * bytes);
    bulk_load_1d(ds_smem + (size_t)s * 2 * DS_STAGE_BYTES, Kc + (size_t)tb * DA_D, bytes, &full_bar[s]);
    bulk_load_1d(ds_smem + (size_t)s * 2 * DS_STAGE_BYTES + DS_STAGE_BYTES, Vc + (size_t)tb * DA_D, bytes, &full_bar[s]);
  };
  if (tid == 0)
    for (int c = 0; c < nst && c < DS_STAGES; ++c) issue(c);
  pdl_wait();
  // No early griddepcontrol.launch_dependents here: this grid runs ~10 waves of CTAs, and dependents released early would sit
  // in their own dependency wait holding shared memory the later waves need (measured: 7.03 -> 6.91 ms per step, config 3).
  {
    const float* qkv = a.qkv + (size_t)r * 3 * Hl;
    const int i = tid & 63;
    const float c = a.rope_cos[(size_t)pos * (DA_D / 2) + i], s = a.rope_sin[(size_t)pos * (DA_D / 2) + i];
    const float* qh = qkv + h * DA_D;
    const float x1 = qh[i], x2 = qh[i + 64];
    q_s[tid] = tid < 64 ? (x1 * c - x2 * s) : (x2 * c + x1 * s);
    if (owns_new) {
      const float* kh = qkv + Hl + h * DA_D;
      const float k1 = kh[i], k2 = kh[i + 64];
      const __nv_bfloat16 kb = __float2bfloat16_rn(tid < 64 ? (k1 * c - k2 * s) : (k2 * c + k1 * s));
      const __nv_bfloat16 vb = __float2bfloat16_rn(qkv[2 * Hl + h * DA_D + tid]);
      Kc[(size_t)pos * DA_D + tid] = kb;
      Vc[(size_t)pos * DA_D + tid] = vb;
      knew_s[tid] = kb;
      vnew_s[tid] = vb;
    }
  }
  __syncthreads();
  float qr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qr[i] = q_s[gl * 8 + i];

  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  auto fold = [&](const uint4 (&kr)[DA_TPG], const uint4 (&vr)[DA_TPG], const bool (&ok)[DA_TPG]) {
    float sc[DA_TPG];
    float mp = -INFINITY;
#pragma unroll
    for (int j = 0; j < DA_TPG; ++j) {
      float s = bf16lo(kr[j].x) * qr[0] + bf16hi(kr[j].x) * qr[1] + bf16lo(kr[j].y) * qr[2] + bf16hi(kr[j].y) * qr[3] +
                bf16lo(kr[j].z) * qr[4] + bf16hi(kr[j].z) * qr[5] + bf16lo(kr[j].w) * qr[6] + bf16hi(kr[j].w) * qr[7];
#pragma unroll
      for (int x = 8; x > 0; x >>= 1) s += __shfl_xor_sync(0xffffffffu, s, x);
      sc[j] = ok[j] ? s * a.scale : -INFINITY;
      mp = fmaxf(mp, sc[j]);
    }
    const float mn = fmaxf(m, mp);
    const float corr = mn == -INFINITY ? 1.f : expf(m - mn);
    l *= corr;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] *= corr;
    m = mn;
#pragma unroll
    for (int j = 0; j < DA_TPG; ++j) {
      const float p = m == -INFINITY ? 0.f : expf(sc[j] - m);
      l += p;
      o[0] = fmaf(p, bf16lo(vr[j].x), o[0]); o[1] = fmaf(p, bf16hi(vr[j].x), o[1]);
      o[2] = fmaf(p, bf16lo(vr[j].y), o[2]); o[3] = fmaf(p, bf16hi(vr[j].y), o[3]);
      o[4] = fmaf(p, bf16lo(vr[j].z), o[4]); o[5] = fmaf(p, bf16hi(vr[j].z), o[5]);
      o[6] = fmaf(p, bf16lo(vr[j].w), o[6]); o[7] = fmaf(p, bf16hi(vr[j].w), o[7]);
    }
  };
  for (int c = 0; c < nst; ++c) {
    const int s = c % DS_STAGES;
    mbar_wait(&full_bar[s], (c / DS_STAGES) & 1);
    const unsigned char* ks = ds_smem + (size_t)s * 2 * DS_STAGE_BYTES;
    const unsigned char* vs = ks + DS_STAGE_BYTES;
    const int tb = t0 + c * DS_TOK;
    uint4 kr[DA_TPG], vr[DA_TPG];
    bool ok[DA_TPG];
#pragma unroll
    for (int j = 0; j < DA_TPG; ++j) {
      const int tl = grp + 8 * j;
      ok[j] = tb + tl < t_old_end;
      kr[j] = make_uint4(0u, 0u, 0u, 0u);
      vr[j] = kr[j];
      if (ok[j]) {
        kr[j] = *reinterpret_cast<const uint4*>(ks + (size_t)tl * (DA_D * 2) + gl * 16);
        vr[j] = *reinterpret_cast<const uint4*>(vs + (size_t)tl * (DA_D * 2) + gl * 16);
      }
    }
    fold(kr, vr, ok);
    __syncthreads();  // every thread is done with stage s
    if (tid == 0 && c + DS_STAGES < nst) issue(c + DS_STAGES);
  }
  if (owns_new) {  // the new token: group 0 takes it, the other groups fold an empty set (no effect)
    uint4 kr[DA_TPG], vr[DA_TPG];
    bool ok[DA_TPG];
#pragma unroll
    for (int j = 0; j < DA_TPG; ++j) {
      ok[j] = false;
      kr[j] = make_uint4(0u, 0u, 0u, 0u);
      vr[j] = kr[j];
    }
    if (grp == 0) {
      ok[0] = true;
      kr[0] = *reinterpret_cast<const uint4*>(knew_s + gl * 8);
      vr[0] = *reinterpret_cast<const uint4*>(vnew_s + gl * 8);
    }
    fold(kr, vr, ok);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) o_s[grp][gl * 8 + i] = o[i];
  if (gl == 0) { ml_s[grp][0] = m; ml_s[grp][1] = l; }
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int g = 0; g < 8; ++g) M = fmaxf(M, ml_s[g][0]);
  float L = 0.f, od = 0.f;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float w = ml_s[g][0] == -INFINITY ? 0.f : expf(ml_s[g][0] - M);
    L += ml_s[g][1] * w;
    od += o_s[g][tid] * w;
  }
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out_bf16) + (size_t)r * Hl + h * DA_D;
  if (nsplit == 1) {
    out[tid] = __float2bfloat16_rn(od / L);
    return;
  }
  const size_t slot = (size_t)rh * a.max_splits + split;
  a.part_o[slot * DA_D + tid] = od;
  if (tid == 0) { a.part_ml[slot * 2] = M; a.part_ml[slot * 2 + 1] = L; }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(&a.counters[rh], 1);
    is_last = prev == nsplit - 1;
    if (is_last) a.counters[rh] = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float MM = -INFINITY;
  for (int s = 0; s < nsplit; ++s) MM = fmaxf(MM, __ldcg(&a.part_ml[((size_t)rh * a.max_splits + s) * 2]));
  float LL = 0.f, acc = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const size_t sl = (size_t)rh * a.max_splits + s;
    const float w = expf(__ldcg(&a.part_ml[sl * 2]) - MM);
    LL += __ldcg(&a.part_ml[sl * 2 + 1]) * w;
    acc += __ldcg(&a.part_o[sl * DA_D + tid]) * w;
  }
  out[tid] = __float2bfloat16_rn(acc / LL);
}

cudaError_t launch_decode_attn(const DecodeAttnArgs& a, cudaStream_t stream) {
  if (a.item_row && a.n_items > 0) {  // streaming kernel over the (row, split) work table
    static const int stages = (getenv("VDD_ATTN_STAGES") && atoi(getenv("VDD_ATTN_STAGES")) == 3) ? 3 : 2;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t ce = cudaFuncSetAttribute(decode_attn_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ds_smem_bytes(2));
      if (ce == cudaSuccess) ce = cudaFuncSetAttribute(decode_attn_stream_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ds_smem_bytes(3));
      if (ce != cudaSuccess) return ce;
      attr_set = true;
    }
    if (a.chunk <= 0 || a.chunk % DS_TOK) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.heads, a.n_items);
    cfg.blockDim = dim3(DA_THREADS);
    cfg.dynamicSmemBytes = ds_smem_bytes(stages);
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return stages == 3 ? cudaLaunchKernelEx(&cfg, decode_attn_stream_kernel<3>, a) : cudaLaunchKernelEx(&cfg, decode_attn_stream_kernel<2>, a);
  }

  // Measured on B200 (bench.py, config 2): global-partials kernel 2.853 ms per decode step, cluster kernel 3.085 ms --
  // cluster scheduling + two cluster barriers cost more than fence + atomic + last-CTA merge at these sizes.  The
  // cluster variant stays available for experiments.
  static const bool cluster_kernel = getenv("VDD_ATTN_CLUSTER") != nullptr;
  if (cluster_kernel) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.R * a.heads, DC_CL);
    cfg.blockDim = dim3(DA_THREADS);
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, decode_attn_cluster_kernel, a);
  }
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
