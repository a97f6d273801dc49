// Varlen multi-head attention for the GEMM-shaped phases on the 5th-generation tensor cores: causal over each stream's
// prompt (LLM prefill, head_dim 128, keys read from the head-major KV cache just written) and non-causal over the 577 ViT
// tokens (head_dim 64).  Same contract as prefill_attn.cu (the mma.sync kernel it replaces on the product path).
//
// Replaces HF eager attention (softmax(QK^T*scale + mask) V with fp32 softmax) as reached from the reference at
// llava_llama.py:91-100 (LLM, step 0 of vcd_sample.py:109/163/178) and clip_encoder.py:48 (vision tower).
//
// One CTA = 128 queries of one (sequence, head), 192 threads:
//   warp 0      TMA producer: Q once, then K / V tiles of 64 keys into two-slot rings (cp.async.bulk.tensor, 128B swizzle)
//   warp 1      TMEM allocator + the single thread that issues tcgen05.mma:
//                 S[128 x 64]  = Q . K^T          A, B from shared memory (both K-major), fp32 accumulator in TMEM
//                 O[128 x D]  += P . V            A = P read from TMEM (bf16 pairs), B = the V tile as it lies in the cache
//                                                 (keys x D, i.e. MN-major: no transposed copy of V is ever made)
//   warps 2..5  softmax: thread = one query row = one TMEM lane.  tcgen05.ld S, scale + mask, running max / sum in the exp2
//               domain, rescale O in TMEM when the row maximum of any lane of the warp moved, tcgen05.st P, signal the issuer.
// TMEM columns: S fp32 [0, 64) | P bf16 pairs [64, 96) | O fp32 [128, 128 + D): 256 allocated, 96 KB of shared memory at
// D = 128, so two CTAs share an SM and one's softmax overlaps the other's MMAs (inside a CTA the steps are serial).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vdd {

constexpr int FA_BM = 128, FA_BN = 64, FA_THREADS = 192, FA_STAGES = 2;
constexpr int FA_TMEM_COLS = 256, FA_COL_S = 0, FA_COL_P = 64, FA_COL_O = 128;
constexpr int fa_smem(int D) { return FA_BM * D * 2 + 2 * FA_STAGES * FA_BN * D * 2 + 1024 /*align*/ + 256 /*barriers*/; }

struct FaArgs {
  const int64_t* k_off;
  const int64_t* k_head_stride;
  int64_t k_tok_stride;
  const int32_t* q_row0;
  const int32_t* seqlen;
  __nv_bfloat16* out;
  int64_t out_ld;
  float scale_log2;  // softmax scale * log2(e)
};

// D[tmem] (+)= A[tmem, bf16 pairs, lane = row] * B[smem desc]; issued by ONE thread
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// MN-major, 128-byte-swizzled operand tile: rows of 64 bf16 along N (128 B) are the K index, 8-row groups 1024 B apart (SBO),
// the next 64 columns of N `lbo_bytes` away (LBO)
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | (64ull << 32) | (1ull << 46) |
         (2ull << 61);
}

template <int D, bool CAUSAL>
__global__ void __launch_bounds__(FA_THREADS, 2)
prefill_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                       const __grid_constant__ CUtensorMap tm_v, const FaArgs a) {
  constexpr int KB = D / 64;                    // 64-column (128-byte) boxes per row of Q / K / V
  constexpr int Q_BYTES = FA_BM * D * 2, KV_BYTES = FA_BN * D * 2;
  const int seq = blockIdx.z, h = blockIdx.y;
  const int qb = gridDim.x - 1 - blockIdx.x;    // longest (last) causal query blocks first
  const int len = a.seqlen[seq];
  const int i0 = qb * FA_BM;
  if (i0 >= len) return;

  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + Q_BYTES;
  unsigned char* sV = sK + FA_STAGES * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FA_STAGES * KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;                  // [FA_STAGES]
  uint64_t* v_full = k_full + FA_STAGES;
  uint64_t* k_empty = v_full + FA_STAGES;
  uint64_t* v_empty = k_empty + FA_STAGES;
  uint64_t* s_full = v_empty + FA_STAGES;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_end = min(i0 + FA_BM, len);
  const int nblk = CAUSAL ? (q_end + FA_BN - 1) / FA_BN : (len + FA_BN - 1) / FA_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<FA_TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer
      const int64_t e = a.k_off[seq] + (int64_t)h * a.k_head_stride[seq];
      const int krow = (int)(e / a.k_tok_stride), kcol = (int)(e % a.k_tok_stride);
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sQ + kb * FA_BM * 128, &tm_q, q_full, h * D + kb * 64, a.q_row0[seq] + i0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % FA_STAGES;
        const uint32_t ph = ((j / FA_STAGES) & 1) ^ 1;
        mbar_wait(&k_empty[s], ph);
        mbar_arrive_expect_tx(&k_full[s], KV_BYTES);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(sK + s * KV_BYTES + kb * FA_BN * 128, &tm_k, &k_full[s], kcol + kb * 64, krow + j * FA_BN);
        mbar_wait(&v_empty[s], ph);
        mbar_arrive_expect_tx(&v_full[s], KV_BYTES);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(sV + s * KV_BYTES + kb * FA_BN * 128, &tm_v, &v_full[s], kcol + kb * 64, krow + j * FA_BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer (single thread).  Order: QK(0) | PV(0) QK(1) | PV(1) QK(2) | ... | PV(n-1)
      const uint32_t idesc_qk = umma_idesc_bf16(FA_BM, FA_BN);
      const uint32_t idesc_pv = umma_idesc_bf16(FA_BM, D) | (1u << 16);  // B (= V) is MN-major
      const uint32_t tS = tmem_base + FA_COL_S, tP = tmem_base + FA_COL_P, tO = tmem_base + FA_COL_O;
      auto issue_qk = [&](int j) {
        const int s = j % FA_STAGES;
        mbar_wait(&k_full[s], (j / FA_STAGES) & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t dq = umma_desc_sw128(smem_u32(sQ + kb * FA_BM * 128));
          const uint64_t dk = umma_desc_sw128(smem_u32(sK + s * KV_BYTES + kb * FA_BN * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS, dq + 2 * k, dk + 2 * k, idesc_qk, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&k_empty[s]);
        umma_commit(s_full);  // also covers every earlier MMA: PV(j-1) is complete when the softmax warps see S(j)
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % FA_STAGES;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[s], (j / FA_STAGES) & 1);
        tcgen05_fence_after();
        const uint32_t sv = smem_u32(sV + s * KV_BYTES);
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k)  // 16 keys = 16 rows of 128 B per step
          umma_bf16_ts(tO, tP + k * 8, umma_desc_sw128_mn(sv + k * 2048, FA_BN * 128), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&v_empty[s]);
        if (j + 1 < nblk) issue_qk(j + 1);
      }
      umma_commit(o_full);
    }
  } else {
    // ===== softmax: warp w owns TMEM lanes 32*(w%4) .. +31; thread = query row
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int i = i0 + row;
    const int lim = CAUSAL ? min(i, len - 1) : len - 1;  // last visible key (rows past the sequence end compute, never store)
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tcgen05_fence_after();
      uint32_t su[FA_BN];  // scores as raw bits (tcgen05.ld writes b32 registers), then scaled + masked in place
#pragma unroll
      for (int c = 0; c < FA_BN / 16; ++c) tmem_ld16(lane_base + FA_COL_S + c * 16, su + c * 16);
      tmem_ld_wait();
      const int t0 = j * FA_BN;
      float mx = m;
#pragma unroll
      for (int c = 0; c < FA_BN; ++c) {
        const float v = (t0 + c <= lim) ? __uint_as_float(su[c]) * a.scale_log2 : -INFINITY;
        su[c] = __float_as_uint(v);
        mx = fmaxf(mx, v);
      }
      const float alpha = exp2f(m - mx);  // 0 on the first block (m = -inf, mx finite: key 0 is always visible)
      if (j > 0 && __any_sync(0xffffffffu, mx > m)) {
#pragma unroll 1
        for (int c = 0; c < D / 16; ++c) {
          uint32_t rr[16];
          tmem_ld16(lane_base + FA_COL_O + c * 16, rr);
          tmem_ld_wait();
#pragma unroll
          for (int x = 0; x < 16; ++x) rr[x] = __float_as_uint(__uint_as_float(rr[x]) * alpha);
          tmem_st16(lane_base + FA_COL_O + c * 16, rr);
        }
      }
      m = mx;
      float sum = 0.f;
      uint32_t pk[FA_BN / 2];
#pragma unroll
      for (int c = 0; c < FA_BN / 2; ++c) {
        const float p0 = exp2f(__uint_as_float(su[2 * c]) - mx), p1 = exp2f(__uint_as_float(su[2 * c + 1]) - mx);
        sum += p0 + p1;
        pk[c] = pack_bf16(p0, p1);
      }
      l = l * alpha + sum;
      tmem_st16(lane_base + FA_COL_P, pk);
      tmem_st16(lane_base + FA_COL_P + 16, pk + 16);
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(p_full);
    }
    mbar_wait(o_full, 0);
    tcgen05_fence_after();
    const float inv = 1.f / l;
    __nv_bfloat16* o = a.out + (int64_t)(a.q_row0[seq] + i) * a.out_ld + h * D;
#pragma unroll 1
    for (int c = 0; c < D / 16; ++c) {
      uint32_t rr[16];
      tmem_ld16(lane_base + FA_COL_O + c * 16, rr);
      tmem_ld_wait();
      if (i < len) {
        uint32_t w[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) w[x] = pack_bf16(__uint_as_float(rr[2 * x]) * inv, __uint_as_float(rr[2 * x + 1]) * inv);
        uint4* dst = reinterpret_cast<uint4*>(o + c * 16);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
      __syncwarp();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<FA_TMEM_COLS>(tmem_base);
}

template <int D, bool CAUSAL>
static cudaError_t launch_fa(const PrefillAttnArgs& a, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(prefill_attn_tc_kernel<D, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, fa_smem(D));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  CUtensorMap tq, tk, tv;
  const int64_t q_cols = (int64_t)a.heads * D;
  if (!make_tmap_bf16_2d(&tq, a.q, a.q_rows, q_cols, a.q_ld, FA_BM) ||
      !make_tmap_bf16_2d(&tk, a.k, a.kv_rows, a.k_tok_stride, a.k_tok_stride, FA_BN) ||
      !make_tmap_bf16_2d(&tv, a.v, a.kv_rows, a.k_tok_stride, a.k_tok_stride, FA_BN))
    return cudaErrorInvalidValue;
  FaArgs fa{};
  fa.k_off = a.k_off; fa.k_head_stride = a.k_head_stride; fa.k_tok_stride = a.k_tok_stride; fa.q_row0 = a.q_row0; fa.seqlen = a.seqlen;
  fa.out = reinterpret_cast<__nv_bfloat16*>(a.out); fa.out_ld = a.out_ld; fa.scale_log2 = a.scale * 1.4426950408889634f;
  const dim3 grid((a.max_seqlen + FA_BM - 1) / FA_BM, a.heads, a.nseq);
  prefill_attn_tc_kernel<D, CAUSAL><<<grid, FA_THREADS, fa_smem(D), stream>>>(tq, tk, tv, fa);
  return cudaGetLastError();
}

bool prefill_attn_tc_supported(const PrefillAttnArgs& a) {
  return (a.head_dim == 128 || a.head_dim == 64) && a.q_rows > 0 && a.kv_rows > 0 && a.k_tok_stride % 8 == 0 && a.q_ld % 8 == 0 &&
         a.out_ld % 8 == 0 && !((uintptr_t)a.q & 15) && !((uintptr_t)a.k & 15) && !((uintptr_t)a.v & 15) && !((uintptr_t)a.out & 15);
}

cudaError_t launch_prefill_attn_tc(const PrefillAttnArgs& a, cudaStream_t stream) {
  if (a.nseq <= 0 || a.max_seqlen <= 0) return cudaSuccess;
  if (!prefill_attn_tc_supported(a)) return cudaErrorInvalidValue;
  if (a.head_dim == 128) return a.causal ? launch_fa<128, true>(a, stream) : launch_fa<128, false>(a, stream);
  return a.causal ? launch_fa<64, true>(a, stream) : launch_fa<64, false>(a, stream);
}

}  // namespace vdd
