// Tensor-core GEMM for the GEMM-shaped parts of the VDD path (LLM prefill of all streams, CLIP ViT, mm_projector,
// and decode steps with more than 8 fused rows):   C[M, N] = A[M, K] . W[N, K]^T   (bf16 in, fp32 accumulate).
//
// Replaces the cuBLAS fp16 GEMMs behind HF LlamaModel / CLIPVisionModel / the projector MLP that the reference
// triggers from llava_llama.py:91-103, clip_encoder.py:48 and llava_arch.py:84.
//
// sm_100a design: one 128x128 output tile per CTA; operands arrive by TMA (cp.async.bulk.tensor, 128B swizzle,
// out-of-bounds rows/columns zero-filled so M, N, K need not be tile multiples) into a 3-stage shared-memory ring
// guarded by full/empty mbarriers; ONE thread issues tcgen05.mma (UMMA 128x128x16, kind::f16, bf16 inputs) with the
// fp32 accumulator in 128 TMEM columns; tcgen05.commit releases ring slots and finally signals the epilogue warps,
// which read TMEM with tcgen05.ld (32 lanes x 16 columns) and apply the fused epilogue (bias, quick_gelu / gelu,
// residual add, SwiGLU over interleaved gate/up columns, row remap for the image splice) straight to global memory.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue (one TMEM lane quadrant each).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vdd {

constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 64;
// Ring depth (template parameter of the kernel):
//   3 x 32 KB per CTA => two CTAs per SM, one's epilogue overlaps the other's main loop (prefill, ViT: many waves of tiles);
//   6 x 32 KB, one CTA per SM, for grids that fit the machine in one wave anyway (decode with > 8 rows: 96-128 CTAs, each a
//   pure weight stream) -- twice the bytes in flight per SM (measured 3.0 TB/s with 3 stages: latency-bound, not HBM-bound).
//   The deep variant also takes its N-tile width at run time (32..256 weight rows, a multiple of 16), chosen so that the
//   tiles of a decode projection fill the SMs in ONE even wave: gate_up (22016 rows) as 138 tiles of 160 rows instead of
//   172 tiles of 128 (which leaves 24 SMs with two tiles and the kernel waiting on them), qkv as 128 tiles of 96.
constexpr int GM_THREADS = 192;
constexpr int GM_STAGE_BYTES = (GM_BM + GM_BN) * GM_BK * 2;  // 32 KB
constexpr int GM_DEEP_MAX_STAGES = 8, GM_DEEP_TMEM_COLS = 256;
constexpr int GM_SMEM_MAX = 227 * 1024;
constexpr int gm_smem(int stages, int stage_bytes = GM_STAGE_BYTES) { return stages * stage_bytes + 1024 /*align*/ + 256 /*barriers*/; }

struct GemmEpi {
  int M, N, K, mode, splits;
  int bn, stages;  // deep variant: N-tile width and ring depth (the shallow variant is compiled for 128 / 3)
  const __nv_bfloat16* bias;
  const int32_t* row_map;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  int64_t ldc;
  QkvRopeEpi rope;      // GEMM_OUT_QKV_ROPE
  int tp;               // > 1: GEMM_OUT_F32 tiles go to every rank (peer memory over NVLink) and every CTA signals every rank
  float* tp_out[8];
  unsigned* tp_cnt[8];
};
// end of a CTA's work under tensor parallelism: publish its peer stores, then bump the arrival counter on every rank
__device__ __forceinline__ void gm_tp_arrive(const GemmEpi& ep) {
  if (ep.tp > 1 && threadIdx.x == 0) {
    __threadfence_system();
    for (int k = 0; k < ep.tp; ++k) asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(ep.tp_cnt[k]) : "memory");
  }
}

__device__ __forceinline__ float quick_gelu(float x) { return x / (1.0f + expf(-1.702f * x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Fused epilogue of 16 consecutive output columns of one row (shared by all variants of the kernel).
__device__ __forceinline__ void gm_store16(const GemmEpi& ep, int64_t orow, int col, float (&v)[16], int zplane) {
  if (ep.bias && zplane == 0) {
    const uint4 b0 = *reinterpret_cast<const uint4*>(ep.bias + col);
    const uint4 b1 = *reinterpret_cast<const uint4*>(ep.bias + col + 8);
    const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[2 * i] += bf16lo(bw[i]); v[2 * i + 1] += bf16hi(bw[i]); }
  }
  if (ep.tp > 1) {
    // tensor parallel (GEMM_OUT_F32): the tile goes straight into every rank's buffer -- this rank's plane of the all-reduce
    // inbox, or its column block of the gathered logits; the bases are pre-offset by the host
    for (int k = 0; k < ep.tp; ++k) {
      float4* o = reinterpret_cast<float4*>(ep.tp_out[k] + orow * ep.ldc + col);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  } else if (ep.mode == GEMM_OUT_F32 || ep.mode == GEMM_OUT_ADD_F32) {
    // split-K: slice z writes its partial to scratch plane z; the planes are added in order by a later pass
    float4* o = reinterpret_cast<float4*>(ep.out_f32 + (ep.splits > 1 ? (int64_t)zplane * ep.M * ep.ldc : 0) + orow * ep.ldc + col);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 t = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      if (ep.mode == GEMM_OUT_ADD_F32 && ep.splits == 1) {
        const float4 old = o[i];
        t.x += old.x; t.y += old.y; t.z += old.z; t.w += old.w;
      }
      o[i] = t;
    }
  } else if (ep.mode == GEMM_OUT_SWIGLU) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float g0 = v[4 * i], u0 = v[4 * i + 1], g1 = v[4 * i + 2], u1 = v[4 * i + 3];
      pk[i] = pack_bf16(g0 / (1.0f + expf(-g0)) * u0, g1 / (1.0f + expf(-g1)) * u1);
    }
    *reinterpret_cast<uint4*>(ep.out_bf16 + orow * ep.ldc + (col >> 1)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  } else {
    if (ep.mode == GEMM_OUT_BF16_QGELU) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = quick_gelu(v[i]);
    } else if (ep.mode == GEMM_OUT_BF16_GELU) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
    }
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
    uint4* o = reinterpret_cast<uint4*>(ep.out_bf16 + orow * ep.ldc + col);
    o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// GEMM_OUT_QKV_ROPE epilogue of one row of a 256-column tile (= two heads of ONE of the q / k / v sections; Hl % 256 == 0):
// rotate-half pairs (d, d + 64) are four 16-column chunks apart in TMEM.  fp32 in, the same arithmetic as rope_kv_prefill_kernel.
__device__ __forceinline__ void gm_qkv_rope_row(const QkvRopeEpi& r, uint32_t tmem_row, int row, int n0, bool store, int hh0, int hh1) {
  const int sec = n0 / r.Hl, h0 = (n0 - sec * r.Hl) >> 7;
  const int seq = r.row_seq[row], pos = r.row_pos[row];
  const float* cs = r.rope_cos + (size_t)pos * 64;
  const float* sn = r.rope_sin + (size_t)pos * 64;
#pragma unroll 1
  for (int hh = hh0; hh < hh1; ++hh) {
    __nv_bfloat16* dst;
    if (sec == 0) dst = reinterpret_cast<__nv_bfloat16*>(r.q_out) + (size_t)row * r.Hl + (h0 + hh) * 128;
    else dst = reinterpret_cast<__nv_bfloat16*>(sec == 1 ? r.kc : r.vc) + r.kv_off[seq] + ((size_t)(h0 + hh) * r.kv_cap[seq] + pos) * 128;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t a[16], b[16];
      tmem_ld16(tmem_row + hh * 128 + cc * 16, a);
      tmem_ld16(tmem_row + hh * 128 + 64 + cc * 16, b);
      tmem_ld_wait();
      uint32_t lo[8], hi[8];
      if (sec < 2) {
        float c[16], s[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 c4 = *reinterpret_cast<const float4*>(cs + cc * 16 + 4 * i), s4 = *reinterpret_cast<const float4*>(sn + cc * 16 + 4 * i);
          c[4 * i] = c4.x; c[4 * i + 1] = c4.y; c[4 * i + 2] = c4.z; c[4 * i + 3] = c4.w;
          s[4 * i] = s4.x; s[4 * i + 1] = s4.y; s[4 * i + 2] = s4.z; s[4 * i + 3] = s4.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x1a = __uint_as_float(a[2 * i]), x1b = __uint_as_float(a[2 * i + 1]);
          const float x2a = __uint_as_float(b[2 * i]), x2b = __uint_as_float(b[2 * i + 1]);
          lo[i] = pack_bf16(rope_first(x1a, x2a, c[2 * i], s[2 * i]), rope_first(x1b, x2b, c[2 * i + 1], s[2 * i + 1]));
          hi[i] = pack_bf16(rope_second(x1a, x2a, c[2 * i], s[2 * i]), rope_second(x1b, x2b, c[2 * i + 1], s[2 * i + 1]));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          lo[i] = pack_bf16(__uint_as_float(a[2 * i]), __uint_as_float(a[2 * i + 1]));
          hi[i] = pack_bf16(__uint_as_float(b[2 * i]), __uint_as_float(b[2 * i + 1]));
        }
      }
      if (!store) continue;
      uint4* d0 = reinterpret_cast<uint4*>(dst + cc * 16);
      uint4* d1 = reinterpret_cast<uint4*>(dst + 64 + cc * 16);
      d0[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      d0[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      d1[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      d1[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    }
  }
}

template <bool DEEP>
__global__ void __launch_bounds__(GM_THREADS, DEEP ? 1 : 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmEpi ep) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int GM_STAGES = DEEP ? ep.stages : 3;
  const int BN = DEEP ? ep.bn : GM_BN;
  const int STAGE_BYTES = (GM_BM + BN) * GM_BK * 2;
  constexpr int TMEM_COLS = DEEP ? GM_DEEP_TMEM_COLS : GM_BN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GM_STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GM_STAGES;
  uint64_t* tmem_full = empty_bar + GM_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * GM_BM;
  // split-K (skinny decode GEMMs whose N/128 tiles cannot fill 148 SMs): this CTA owns k-blocks [kb0, kb1)
  const int nkb_all = (ep.K + GM_BK - 1) / GM_BK;
  const int kb0 = (int)(((long long)blockIdx.z * nkb_all) / ep.splits), kb1 = (int)(((long long)(blockIdx.z + 1) * nkb_all) / ep.splits);
  const int nkb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < GM_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Set-up above overlaps the tail of the previous kernel when launched with programmatic stream serialization (decode);
  // everything below may read what that kernel wrote.  No-op for ordinary launches.
  // (the producer thread first requests the W tiles of the first ring stages: weights never depend on the previous kernel)
  if (!(warp == 0 && lane == 0)) {
    pdl_wait();
    pdl_launch_dependents();
  }

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer
      const int npre = nkb < GM_STAGES ? nkb : GM_STAGES;
      for (int kb = 0; kb < npre; ++kb) {
        mbar_arrive_expect_tx(&full_bar[kb], STAGE_BYTES);
        tma_load_2d(smem + kb * STAGE_BYTES + GM_BM * GM_BK * 2, &tmap_b, &full_bar[kb], (kb0 + kb) * GM_BK, n0);
      }
      pdl_wait();
      pdl_launch_dependents();
      for (int kb = 0; kb < npre; ++kb) tma_load_2d(smem + kb * STAGE_BYTES, &tmap_a, &full_bar[kb], (kb0 + kb) * GM_BK, m0);
      for (int kb = npre; kb < nkb; ++kb) {
        const int s = kb % GM_STAGES;
        mbar_wait(&empty_bar[s], ((kb / GM_STAGES) & 1) ^ 1);
        unsigned char* sa = smem + s * STAGE_BYTES;
        unsigned char* sb = sa + GM_BM * GM_BK * 2;
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_2d(sa, &tmap_a, &full_bar[s], (kb0 + kb) * GM_BK, m0);
        tma_load_2d(sb, &tmap_b, &full_bar[s], (kb0 + kb) * GM_BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer (single thread)
      const uint32_t idesc = umma_idesc_bf16(GM_BM, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GM_STAGES;
        mbar_wait(&full_bar[s], (kb / GM_STAGES) & 1);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + GM_BM * GM_BK * 2;
        const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < GM_BK / 16; ++k)
          umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);  // +32 B along K
        umma_commit(&empty_bar[s]);  // frees the ring slot once these MMAs have read it
      }
      umma_commit(tmem_full);  // accumulator complete
    }
  } else {
    // ===== epilogue: warp w owns TMEM lanes 32*(w%4) .. +31  (rows of the tile)
    const int quad = warp & 3;
    const int row = m0 + quad * 32 + lane;
    mbar_wait(tmem_full, 0);
    tcgen05_fence_after();
    int64_t orow = -1;
    if (row < ep.M) orow = ep.row_map ? ep.row_map[row] : row;
#pragma unroll 1
    for (int c = 0; c < BN / 16; ++c) {
      uint32_t rr[16];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + c * 16, rr);
      tmem_ld_wait();
      const int col = n0 + c * 16;
      if (orow >= 0 && col < ep.N) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[i]);
        gm_store16(ep, orow, col, v, (int)blockIdx.z);
      }
      __syncwarp();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  gm_tp_arrive(ep);
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}


// ------------------------------------------------------------------------------------------------ wide persistent variant
// Prefill / ViT / projector GEMMs (M > 128 rows): the 128x128 tile above reads 8 KB of shared memory per UMMA of 64 clocks --
// exactly the 128 B/clk an SM can deliver -- and measured 38-52 % tensor-pipe activity (profiles/r1_ncu_prefill.json).  Here a
// tile is 128 x 256 (UMMA 128x256x16: 12 KB per 128 clocks = 96 B/clk), CTAs are persistent (one per SM) and walk the tiles
// in groups (gemm_tile_coords below) so that concurrently running CTAs share operand panels in L2, and the fp32 accumulator is double-buffered in TMEM (2 x 256
// columns) so the four epilogue warps drain tile i while the tensor core works on tile i+1.
constexpr int GW_BN = 256, GW_STAGES = 4;
// Persistent tile walk, grouped: GROUP m-tiles x all n-tiles at a time, m fastest inside the group.  The CTAs working at the same
// moment then share GROUP A panels and (CTAs / GROUP) B panels instead of one B panel and as many A panels as there are CTAs --
// with 24 512 prefill rows (BASELINE configs[2]) the plain m-fastest walk re-read all of A (200 MB, more than L2) once per n-tile:
// ~9.6 GB of DRAM traffic for the qkv projection, 4 TB/s, i.e. a DRAM-bound GEMM.  Small M (fewer m-tiles than GROUP) degenerates
// to the m-fastest walk.
constexpr int GEMM_TILE_GROUP = 16;
__host__ __device__ __forceinline__ void gemm_tile_coords(int tile, int m_tiles, int n_tiles, int& m_idx, int& n_idx) {
  const int per_group = GEMM_TILE_GROUP * n_tiles;
  const int g = tile / per_group, r = tile - g * per_group;
  const int first_m = g * GEMM_TILE_GROUP;
  const int gm = GEMM_TILE_GROUP < m_tiles - first_m ? GEMM_TILE_GROUP : m_tiles - first_m;
  n_idx = r / gm;
  m_idx = first_m + (r - n_idx * gm);
}
constexpr int GW_STAGE_BYTES = (GM_BM + GW_BN) * GM_BK * 2;  // 48 KB
constexpr int GW_SMEM = GW_STAGES * GW_STAGE_BYTES + 1024 + 256;

__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_tc_wide_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmEpi ep,
                    int m_tiles, int n_tiles) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GW_STAGES * GW_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GW_STAGES;
  uint64_t* tfull_bar = empty_bar + GW_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (ep.K + GM_BK - 1) / GM_BK;
  const int tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < GW_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      int gi = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int mi, ni;
        gemm_tile_coords(tile, m_tiles, n_tiles, mi, ni);
        const int m0 = mi * GM_BM, n0 = ni * GW_BN;
        for (int kb = 0; kb < nkb; ++kb, ++gi) {
          const int s = gi % GW_STAGES;
          mbar_wait(&empty_bar[s], ((gi / GW_STAGES) & 1) ^ 1);
          unsigned char* sa = smem + s * GW_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], GW_STAGE_BYTES);
          tma_load_2d(sa, &tmap_a, &full_bar[s], kb * GM_BK, m0);
          tma_load_2d(sa + GM_BM * GM_BK * 2, &tmap_b, &full_bar[s], kb * GM_BK, n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GM_BM, GW_BN);
      int gi = 0, it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)buf * GW_BN;
        for (int kb = 0; kb < nkb; ++kb, ++gi) {
          const int s = gi % GW_STAGES;
          mbar_wait(&full_bar[s], (gi / GW_STAGES) & 1);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + s * GW_STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + GM_BM * GM_BK * 2);
#pragma unroll
          for (int k = 0; k < GM_BK / 16; ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      int mi, ni;
      gemm_tile_coords(tile, m_tiles, n_tiles, mi, ni);
      const int m0 = mi * GM_BM, n0 = ni * GW_BN;
      const int buf = it & 1;
      const int row = m0 + quad * 32 + lane;
      int64_t orow = -1;
      if (row < ep.M) orow = ep.row_map ? ep.row_map[row] : row;
      mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < GW_BN / 16; c += 2) {  // two 16-column chunks per TMEM round trip
        uint32_t r0[16], r1[16];
        const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * GW_BN + c * 16;
        tmem_ld16(ta, r0);
        tmem_ld16(ta + 16, r1);
        tmem_ld_wait();
        const int col = n0 + c * 16;
        if (orow >= 0 && col < ep.N) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r0[i]);
          gm_store16(ep, orow, col, v, 0);
        }
        if (orow >= 0 && col + 16 < ep.N) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r1[i]);
          gm_store16(ep, orow, col + 16, v, 0);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  gm_tp_arrive(ep);
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ CTA-pair variant
// Large prefill GEMMs (at least one wave of 256x256 tiles over the 74 SM pairs): the 128x256 kernel above is bound by shared-memory
// bandwidth -- per UMMA of 128 clocks an SM reads 12 KB of operands and TMA writes 12 KB of the next ones, 192 B/clk against the
// 128 B/clk an SM has -- and measured 1030-1170 TFLOP/s (profiles/r2_launch_list_config2_prefill.json).  Here two CTAs of one
// cluster (the two SMs of a TPC) share a 256x256 tile through tcgen05.mma.cta_group::2: CTA r stages rows [128r, 128r+128) of A
// and rows [128r, 128r+128) of the B tile (8 KB + 8 KB per UMMA instead of 12 + 12), the LEADER CTA's single thread issues the
// UMMA 256x256x16 for both, and each CTA's TMEM receives the accumulator of its own 128 rows.  Barriers: both CTAs' TMA loads
// complete_tx on the leader's `full` barrier; the leader's tcgen05.commit multicasts `empty` / `tmem full` to both CTAs; the
// epilogue warps of both CTAs arrive on the leader's `tmem empty` barrier.  Everything else (persistent m-fastest tile walk, TMEM
// double buffer, fused epilogues) is the wide kernel's.
constexpr int GP_BM = 256, GP_BN = 256, GP_STAGES = 6;
// warps: 0 = TMA producer, 1 = TMEM allocator + MMA issuer (leader CTA), 2..9 = epilogue -- TWO warps per TMEM lane quadrant, each
// draining one 128-column half of the accumulator (with four warps the QuickGELU / SwiGLU / RoPE epilogues of short-K GEMMs, e.g.
// the ViT's fc1 at K = 1024, took longer than the tile's MMAs and the tensor core waited for a free accumulator)
constexpr int GP_THREADS = 320, GP_EPI_WARPS = 8;
constexpr int GP_STAGE_BYTES = (GM_BM + GP_BN / 2) * GM_BK * 2;  // 32 KB per CTA
constexpr int GP_SMEM = GP_STAGES * GP_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GP_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmEpi ep,
                    int m_tiles, int n_tiles) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GP_STAGES * GP_STAGE_BYTES);  // the leader's are the ones in use
  uint64_t* empty_bar = full_bar + GP_STAGES;
  uint64_t* tfull_bar = empty_bar + GP_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;                                                 // the leader's are the ones in use
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nkb = (ep.K + GM_BK - 1) / GM_BK;
  const int tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < GP_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 2 * GP_EPI_WARPS);  // the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {  // one warp of EACH CTA, same shared-memory slot: the allocation is collective over the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();  // the peer's barriers exist before anything signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      int gi = 0;
      for (int tile = pair; tile < tiles; tile += npairs) {
        int mi, ni;
        gemm_tile_coords(tile, m_tiles, n_tiles, mi, ni);
        const int m0 = mi * GP_BM + (int)rank * GM_BM, n0 = ni * GP_BN + (int)rank * (GP_BN / 2);
        for (int kb = 0; kb < nkb; ++kb, ++gi) {
          const int s = gi % GP_STAGES;
          mbar_wait(&empty_bar[s], ((gi / GP_STAGES) & 1) ^ 1);
          unsigned char* sa = smem + s * GP_STAGE_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * GP_STAGE_BYTES);  // both CTAs' bytes land on the leader's barrier
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
          tma_load_2d_pair(sa, &tmap_a, fb, kb * GM_BK, m0);
          tma_load_2d_pair(sa + GM_BM * GM_BK * 2, &tmap_b, fb, kb * GM_BK, n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GP_BM, GP_BN);
      int gi = 0, it = 0;
      for (int tile = pair; tile < tiles; tile += npairs, ++it) {
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator
        tcgen05_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)buf * GP_BN;
        for (int kb = 0; kb < nkb; ++kb, ++gi) {
          const int s = gi % GP_STAGES;
          mbar_wait(&full_bar[s], (gi / GP_STAGES) & 1);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + s * GP_STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + GM_BM * GM_BK * 2);
#pragma unroll
          for (int k = 0; k < GM_BK / 16; ++k) umma_bf16_pair(acc, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(&empty_bar[s]);
        }
        umma_commit_pair(&tfull_bar[buf]);
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3, half = (warp - 2) >> 2;  // TMEM lanes 32*quad.., accumulator columns [128*half, 128*half + 128)
    int it = 0;
    for (int tile = pair; tile < tiles; tile += npairs, ++it) {
      int mi, ni;
      gemm_tile_coords(tile, m_tiles, n_tiles, mi, ni);
      const int m0 = mi * GP_BM + (int)rank * GM_BM, n0 = ni * GP_BN;
      const int buf = it & 1;
      const int row = m0 + quad * 32 + lane;
      int64_t orow = -1;
      if (row < ep.M) orow = ep.row_map ? ep.row_map[row] : row;
      mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
      tcgen05_fence_after();
      if (ep.mode == GEMM_OUT_QKV_ROPE) {
        // every lane runs the (warp-collective) TMEM loads; lanes past M read row M-1's metadata and store nothing
        gm_qkv_rope_row(ep.rope, tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * GP_BN, min(row, ep.M - 1), n0, row < ep.M,
                        half, half + 1);
      } else
#pragma unroll 1
      for (int c = half * (GP_BN / 32); c < (half + 1) * (GP_BN / 32); c += 2) {
        uint32_t r0[16], r1[16];
        const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * GP_BN + c * 16;
        tmem_ld16(ta, r0);
        tmem_ld16(ta + 16, r1);
        tmem_ld_wait();
        const int col = n0 + c * 16;
        if (orow >= 0 && col < ep.N) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r0[i]);
          gm_store16(ep, orow, col, v, 0);
        }
        if (orow >= 0 && col + 16 < ep.N) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r1[i]);
          gm_store16(ep, orow, col + 16, v, 0);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[buf]), 0));
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while its peer can still signal its barriers or read its operands
  gm_tp_arrive(ep);
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static const char* g_init_err = nullptr;

static bool gemm_init() {
  static bool done = false;
  if (done) return g_encode != nullptr;
  done = true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    g_init_err = "cuTensorMapEncodeTiled not available from the driver";
    return false;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  e = cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gm_smem(3));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_MAX);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GW_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_SMEM);
  if (e != cudaSuccess) {
    g_init_err = "cudaFuncSetAttribute(gemm_tc_kernel) failed";
    g_encode = nullptr;
    return false;
  }
  return true;
}
const char* gemm_init_error() { return g_init_err; }

// 2D bf16 row-major [rows, K] tensor map with a {64, box_rows} box and 128-byte swizzle
static bool make_tmap(CUtensorMap* m, const void* ptr, int rows, int K, int box_rows = GM_BM) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {GM_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t pitch_elems, int box_rows) {
  if (!gemm_init() || rows <= 0 || cols < GM_BK || box_rows <= 0 || box_rows > 256 || (pitch_elems % 8) || ((uintptr_t)ptr & 15)) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {GM_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool gemm_uses_pair_kernel(int M, int N) {
  static const bool off = getenv("VDD_GEMM_NO_PAIR") != nullptr || getenv("VDD_GEMM_NARROW") != nullptr || getenv("VDD_GEMM_SHALLOW") != nullptr;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int pair_m = (M + GP_BM - 1) / GP_BM, pair_n = (N + GP_BN - 1) / GP_BN;
  return !off && M > GM_BM && (num_sms % 2) == 0 && pair_m * pair_n >= num_sms / 2;
}
void gemm_tile_coords_host(int tile, int m_tiles, int n_tiles, int* m_idx, int* n_idx) { gemm_tile_coords(tile, m_tiles, n_tiles, *m_idx, *n_idx); }
bool gemm_qkv_rope_fused(int M, int N, int K) {
  static const bool off = getenv("VDD_NO_ROPE_FUSION") != nullptr;
  return !off && gemm_init() && gemm_uses_pair_kernel(M, N) && N % (3 * GP_BN) == 0 && K % 8 == 0;
}

cudaError_t launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  if (!gemm_init()) return cudaErrorNotSupported;
  if (a.M <= 0) return cudaSuccess;
  if ((a.K % 8) != 0 || (a.N % 16) != 0 || ((uintptr_t)a.a & 15) || ((uintptr_t)a.w & 15)) return cudaErrorInvalidValue;
  CUtensorMap ta, tb;
  GemmEpi ep;
  ep.M = a.M; ep.N = a.N; ep.K = a.K; ep.mode = a.epilogue;
  ep.bias = reinterpret_cast<const __nv_bfloat16*>(a.bias);
  ep.row_map = a.row_map;
  ep.out_f32 = a.out_f32;
  ep.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
  ep.ldc = a.ldc;
  ep.tp = a.tp_n > 1 ? a.tp_n : 0;
  for (int k = 0; k < 8; ++k) { ep.tp_out[k] = a.tp_out[k]; ep.tp_cnt[k] = a.tp_cnt[k]; }
  if (ep.tp && (a.epilogue != GEMM_OUT_F32 || a.row_map || a.bias || a.tp_n > 8)) return cudaErrorInvalidValue;
  dim3 grid((a.N + GM_BN - 1) / GM_BN, (a.M + GM_BM - 1) / GM_BM);
  ep.splits = 1;
  if (a.epilogue == GEMM_OUT_ADD_F32 && a.M <= GM_BM && a.split_ws && !a.row_map && !a.bias) {
    // decode-sized residual GEMM: N/128 tiles cannot fill 148 SMs, so spread K over the idle ones
    const int tiles = grid.x * grid.y, nkb = (a.K + GM_BK - 1) / GM_BK;
    int sp = 148 / (tiles > 0 ? tiles : 1);
    if (sp > nkb / 4) sp = nkb / 4;
    if (sp > 16) sp = 16;
    while (sp > 1 && (size_t)sp * a.M * a.ldc * 4 > a.split_ws_bytes) --sp;
    if (sp > 1) ep.splits = sp;
  }
  grid.z = ep.splits;
  if (ep.splits > 1) ep.out_f32 = a.split_ws;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(GM_THREADS);
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  static const bool shallow_only = getenv("VDD_GEMM_SHALLOW") != nullptr, fixed_bn = getenv("VDD_GEMM_BN128") != nullptr,
                    no_wide = getenv("VDD_GEMM_NARROW") != nullptr;
  const int pair_m = (a.M + GP_BM - 1) / GP_BM, pair_n = (a.N + GP_BN - 1) / GP_BN;
  const bool use_pair = gemm_uses_pair_kernel(a.M, a.N) && ep.splits == 1;
  if (a.epilogue == GEMM_OUT_QKV_ROPE) {
    if (!use_pair || !a.qkv_rope || a.qkv_rope->Hl % GP_BN || a.N != 3 * a.qkv_rope->Hl || a.bias || a.row_map || ep.tp) return cudaErrorInvalidValue;
    ep.rope = *a.qkv_rope;
  }
  if (use_pair) {
    // at least one wave of 256x256 tiles over the SM pairs: cta_group::2
    ep.bn = GP_BN;
    ep.stages = GP_STAGES;
    if (!make_tmap(&ta, a.a, a.M, a.K) || !make_tmap(&tb, a.w, a.N, a.K, GP_BN / 2)) return cudaErrorInvalidValue;
    cfg.gridDim = dim3(num_sms);  // 74 clusters of 2 (the kernel carries __cluster_dims__(2, 1, 1))
    cfg.blockDim = dim3(GP_THREADS);
    cfg.dynamicSmemBytes = GP_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute atp[1];
    atp[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    atp[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = atp;
    cfg.numAttrs = a.pdl ? 1 : 0;
    if (a.defer_reduce) *a.defer_reduce = 0;
    if (a.tp_ctas) *a.tp_ctas = (int)cfg.gridDim.x;
    return cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel, ta, tb, ep, pair_m, pair_n);
  }
  const int wide_tiles = ((a.M + GM_BM - 1) / GM_BM) * ((a.N + GW_BN - 1) / GW_BN);
  if (a.M > GM_BM && ep.splits == 1 && !no_wide && !shallow_only && wide_tiles >= num_sms) {
    // GEMM-shaped work (prefill of all streams, ViT, projector): persistent 128x256 tiles, TMEM double buffer.  Problems with
    // less than one wave of such tiles keep the 128x128 tiles below (measured at 670 rows: o_proj / down_proj = 96 wide tiles
    // 49.6 / 91.2 us against 42.3 / 83.6 us; qkv / gate_up = 288 / 516 tiles 66 / 116 us against 83 / 159 us).
    const int m_tiles = (a.M + GM_BM - 1) / GM_BM, n_tiles = (a.N + GW_BN - 1) / GW_BN;
    ep.bn = GW_BN;
    ep.stages = GW_STAGES;
    if (!make_tmap(&ta, a.a, a.M, a.K) || !make_tmap(&tb, a.w, a.N, a.K, GW_BN)) return cudaErrorInvalidValue;
    const int tiles = m_tiles * n_tiles;
    cfg.gridDim = dim3(tiles < num_sms ? tiles : num_sms);
    cfg.dynamicSmemBytes = GW_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute atw[1];
    atw[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    atw[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = atw;
    cfg.numAttrs = a.pdl ? 1 : 0;
    if (a.defer_reduce) *a.defer_reduce = 0;
    if (a.tp_ctas) *a.tp_ctas = (int)cfg.gridDim.x;
    return cudaLaunchKernelEx(&cfg, gemm_tc_wide_kernel, ta, tb, ep, m_tiles, n_tiles);
  }
  ep.bn = GM_BN;
  ep.stages = 3;
  bool deep = false;
  if (!shallow_only) {
    if (a.M <= GM_BM && ep.splits == 1 && !fixed_bn) {
      // decode-sized projection without split-K: widen / narrow the N tile until the tiles form one even wave over the SMs
      int bn = ((a.N + num_sms - 1) / num_sms + 15) / 16 * 16;
      if (bn < 32) bn = 32;
      if (bn <= 256) {
        ep.bn = bn;
        grid.x = (a.N + bn - 1) / bn;
      }
    }
    deep = (int)(grid.x * grid.y * grid.z) <= num_sms;  // one wave at one CTA per SM
  }
  if (deep) {
    const int stage_bytes = (GM_BM + ep.bn) * GM_BK * 2;
    ep.stages = (GM_SMEM_MAX - 1024 - 256) / stage_bytes;
    if (ep.stages > GM_DEEP_MAX_STAGES) ep.stages = GM_DEEP_MAX_STAGES;
    cfg.dynamicSmemBytes = gm_smem(ep.stages, stage_bytes);
  } else {
    ep.bn = GM_BN;
    grid.x = (a.N + GM_BN - 1) / GM_BN;
    cfg.dynamicSmemBytes = gm_smem(3);
  }
  cfg.gridDim = grid;
  if (a.tp_ctas) *a.tp_ctas = (int)(grid.x * grid.y * grid.z);
  if (!make_tmap(&ta, a.a, a.M, a.K) || !make_tmap(&tb, a.w, a.N, a.K, ep.bn)) return cudaErrorInvalidValue;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = a.pdl ? 1 : 0;
  cudaError_t le = deep ? cudaLaunchKernelEx(&cfg, gemm_tc_kernel<true>, ta, tb, ep) : cudaLaunchKernelEx(&cfg, gemm_tc_kernel<false>, ta, tb, ep);
  if (a.defer_reduce) *a.defer_reduce = ep.splits > 1 ? ep.splits : 0;
  if (le != cudaSuccess || ep.splits == 1 || a.defer_reduce) return le;
  return launch_reduce_splits(a.out_f32, a.split_ws, (int64_t)a.M * a.ldc, ep.splits, stream);
}

}  // namespace vdd
