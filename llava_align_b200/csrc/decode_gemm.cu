// Decode projections for 8 < R <= 256 fused rows (BASELINE configs 3-5: 16-192 rows of x against W[N, K]).
//
// Replaces the cuBLAS fp16 GEMMs HF LlamaDecoderLayer / lm_head issue once per stream per step when the reference decodes
// (llava_llama.py:91-103 entered from vcd_sample.py:108-113,161-185); here all streams of all sequences are the rows of
// ONE skinny GEMM per projection, so the weights cross HBM once per step.
//
// At these row counts the projection is still a weight stream (R < 212 = the bf16 ridge of a B200), so the kernel is
// built like a copy engine with a tensor core attached, not like a square GEMM:
//   * roles are SWAPPED: the weight tile is the UMMA A operand (M = 128 output features, two of them per stage = 256
//     features), the activations are the B operand (N = R rounded up to 16), so no MMA work or TMEM is spent on padding
//     rows and a 32 KB weight tile needs only R*128 bytes of activations from L2 beside it;
//   * stream-K: the (feature tile, k-block) space is cut into gridDim.x equal contiguous ranges, one persistent CTA per
//     SM, so every SM streams the same number of bytes whatever N/256 is (o_proj has 16 tiles, gate_up 86, lm_head 125);
//   * TMA (cp.async.bulk.tensor, 128B swizzle, weights with an evict-first L2 policy) fills a 4-6 stage ring, one
//     thread issues tcgen05.mma into a double-buffered TMEM accumulator, four epilogue warps drain the previous
//     segment with tcgen05.ld while the next one streams;
//   * a tile whose k-range is shared by several CTAs is finished without a second launch: every sharer parks its fp32
//     partial in an L2-resident workspace and bumps the tile's counter; once all have arrived each sharer sums its share
//     of the tile over the segments IN SEGMENT ORDER (bit-identical run to run, independent of arrival order) and applies
//     the epilogue (store / residual add / SwiGLU over interleaved gate-up features);
//   * the weight loads of the first ring stages are issued BEFORE griddepcontrol.wait (weights never depend on the
//     previous kernel), the activation loads after it.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue (TMEM lane quadrant = warp % 4).
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace vdd {

constexpr int SK_BM = 256, SK_BK = 64, SK_THREADS = 192;
constexpr int SK_A_BYTES = SK_BM * SK_BK * 2;  // 32 KB of weights per stage
constexpr int SK_MAX_STAGES = 8;
constexpr int SK_SMEM_LIMIT = 227 * 1024;
constexpr int SK_TMEM_COLS = 512;

struct SkinnyParams {
  int R, RP, N, K, mode;
  int nkb, tiles, stages, nbuf, maxseg;
  long long units;  // tiles * nkb
  float* out_f32;
  __nv_bfloat16* out_bf16;
  int64_t ld;
  float* ws;      // [tiles][maxseg][2][RP][128] fp32 partials
  int* counters;  // [tiles], zero on entry
  unsigned long long* trace;  // debug: [grid][8] globaltimer stamps, or null
};

__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ unsigned long long sk_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SK_STAMP(slot)                                                              \
  do {                                                                              \
    if (p.trace) p.trace[(size_t)blockIdx.x * 8 + (slot)] = sk_now();               \
  } while (0)
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Values v[0..15] of output feature n (one feature per lane, 32 consecutive features per warp) for rows r0..r0+15.
// Every lane of the warp must call this (SwiGLU pairs features across lanes).
__device__ __forceinline__ void sk_emit16(const SkinnyParams& p, int n, int r0, const float (&v)[16]) {
  const int lane = threadIdx.x & 31;
  if (p.mode == GEMM_OUT_SWIGLU) {
    // features interleave gate (even) / up (odd); out[r, n/2] = silu(gate) * up, two results packed per 32-bit store
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float up = __shfl_xor_sync(0xffffffffu, v[i], 1);
      const float hval = v[i] / (1.0f + expf(-v[i])) * up;  // meaningful on even lanes
      const float hnext = __shfl_down_sync(0xffffffffu, hval, 2);
      const int r = r0 + i;
      if ((lane & 3) == 0 && r < p.R) {
        if (n + 3 < p.N) {
          *reinterpret_cast<uint32_t*>(p.out_bf16 + (int64_t)r * p.ld + (n >> 1)) = pack_bf16(hval, hnext);
        } else if (n + 1 < p.N) {
          p.out_bf16[(int64_t)r * p.ld + (n >> 1)] = __float2bfloat16_rn(hval);
        }
      }
    }
    return;
  }
  if (n >= p.N) return;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = r0 + i;
    if (r < p.R) {
      float* o = p.out_f32 + (int64_t)r * p.ld + n;
      *o = p.mode == GEMM_OUT_ADD_F32 ? *o + v[i] : v[i];
    }
  }
}

__global__ void __launch_bounds__(SK_THREADS, 1)
skinny_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, const SkinnyParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int stage_bytes = SK_A_BYTES + p.RP * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + SK_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + SK_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long G = gridDim.x, c = blockIdx.x;
  const long long u0 = c * p.units / G, u1 = (c + 1) * p.units / G;
  const int n_units = (int)(u1 - u0);

  if (threadIdx.x == 0) SK_STAMP(0);  // kernel entry
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<SK_TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: one continuous stream over this CTA's unit range (segments only matter to the MMA side)
      const uint64_t pol = policy_evict_first();
      const int npre = n_units < p.stages ? n_units : p.stages;
      for (int i = 0; i < npre; ++i) {  // weights first: they do not depend on the previous kernel
        const long long u = u0 + i;
        const int t = (int)(u / p.nkb), kb = (int)(u % p.nkb);
        mbar_arrive_expect_tx(&full_bar[i], stage_bytes);
        tma_load_2d_hint(smem + (size_t)i * stage_bytes, &tmap_w, &full_bar[i], kb * SK_BK, t * SK_BM, pol);
      }
      SK_STAMP(1);  // weights of the first stages requested
      pdl_wait();
      SK_STAMP(2);  // previous kernel complete
      for (int i = 0; i < npre; ++i) {
        const int kb = (int)((u0 + i) % p.nkb);
        tma_load_2d(smem + (size_t)i * stage_bytes + SK_A_BYTES, &tmap_x, &full_bar[i], kb * SK_BK, 0);
      }
      for (int i = npre; i < n_units; ++i) {
        const int s = i % p.stages;
        const long long u = u0 + i;
        const int t = (int)(u / p.nkb), kb = (int)(u % p.nkb);
        mbar_wait(&empty_bar[s], ((i / p.stages) & 1) ^ 1);
        unsigned char* sa = smem + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        tma_load_2d_hint(sa, &tmap_w, &full_bar[s], kb * SK_BK, t * SK_BM, pol);
        tma_load_2d(sa + SK_A_BYTES, &tmap_x, &full_bar[s], kb * SK_BK, 0);
      }
    }
    __syncwarp();  // reconverge before the CTA-wide barrier at the end
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: per stage 4 k-steps x 2 feature halves of UMMA 128 x RP x 16
      const uint32_t idesc = umma_idesc_bf16(128, p.RP);
      int seg = 0, buf = 0;
      for (int i = 0; i < n_units; ++i) {
        const int kb = (int)((u0 + i) % p.nkb);
        const bool seg_start = i == 0 || kb == 0;
        if (seg_start) {
          buf = seg % p.nbuf;
          mbar_wait(&tempty_bar[buf], ((seg / p.nbuf) & 1) ^ 1);  // the epilogue has drained this accumulator
          tcgen05_fence_after();
        }
        const int s = i % p.stages;
        mbar_wait(&full_bar[s], (i / p.stages) & 1);
        if (i == 0) SK_STAMP(3);  // first stage landed
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
        const uint64_t da0 = umma_desc_sw128(sa), da1 = umma_desc_sw128(sa + 128 * SK_BK * 2), db = umma_desc_sw128(sa + SK_A_BYTES);
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * 2) * p.RP, acc1 = acc0 + p.RP;
#pragma unroll
        for (int k = 0; k < SK_BK / 16; ++k) {
          const uint32_t accum = (seg_start && k == 0) ? 0u : 1u;
          umma_bf16(acc0, da0 + 2 * k, db + 2 * k, idesc, accum);
          umma_bf16(acc1, da1 + 2 * k, db + 2 * k, idesc, accum);
        }
        umma_commit(&empty_bar[s]);
        if (i == n_units - 1 || kb == p.nkb - 1) {
          umma_commit(&tfull_bar[buf]);
          ++seg;
        }
      }
      SK_STAMP(4);  // last MMA issued
    }
    __syncwarp();
  } else {
    // ===== epilogue warps
    const int quad = warp & 3;
    const int nch = p.RP >> 4;
    const size_t slot_floats = (size_t)2 * p.RP * 128;
    pdl_wait();  // outputs (and the residual for ADD) belong to earlier kernels until here
    pdl_launch_dependents();
    // ---- phase 1: drain every segment of this CTA; whole tiles go straight out, shared tiles park a partial
    {
      int seg = 0, i = 0;
      while (i < n_units) {
        const long long u = u0 + i;
        const int t = (int)(u / p.nkb), kb0 = (int)(u % p.nkb);
        const int len = min(n_units - i, p.nkb - kb0);
        const bool whole = kb0 == 0 && len == p.nkb;
        const int buf = seg % p.nbuf;
        mbar_wait(&tfull_bar[buf], (seg / p.nbuf) & 1);
        tcgen05_fence_after();
        int myseg = 0;
        if (!whole) {
          const long long s_lo = (long long)t * p.nkb;
          const long long c_first = ((s_lo + 1) * G + p.units - 1) / p.units - 1;
          myseg = (int)(c - c_first);
        }
        float* wsp = p.ws + ((size_t)t * p.maxseg + myseg) * slot_floats;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int n = t * SK_BM + h * 128 + quad * 32 + lane;
#pragma unroll 1
          for (int cc = 0; cc < nch; ++cc) {
            uint32_t rr[16];
            tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 2 + h) * p.RP + cc * 16, rr);
            tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
            if (whole) {
              sk_emit16(p, n, cc * 16, v);
            } else {
              float* w = wsp + ((size_t)h * p.RP + cc * 16) * 128 + quad * 32 + lane;
#pragma unroll
              for (int j = 0; j < 16; ++j) w[(size_t)j * 128] = v[j];
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        if (!whole) {
          __threadfence();
          __syncwarp();
          if (lane == 0) atomicAdd(&p.counters[t], 1);
        }
        i += len;
        ++seg;
      }
    }
    if (warp == 2 && lane == 0) SK_STAMP(5);  // phase 1 done (all segments drained)
    // ---- phase 2: finish the shared tiles.  Sharer `myseg` of `nseg` takes items myseg, myseg + nseg, ... of the tile's
    //      8 * RP/16 items (32 features x 16 rows), its four warps interleaved; sums run over segments 0..nseg-1 in order.
    {
      int i = 0;
      while (i < n_units) {
        const long long u = u0 + i;
        const int t = (int)(u / p.nkb), kb0 = (int)(u % p.nkb);
        const int len = min(n_units - i, p.nkb - kb0);
        i += len;
        if (kb0 == 0 && len == p.nkb) continue;
        const long long s_lo = (long long)t * p.nkb, s_hi = s_lo + p.nkb;
        const long long c_first = ((s_lo + 1) * G + p.units - 1) / p.units - 1;
        const long long c_last = (s_hi * G + p.units - 1) / p.units - 1;
        const int nseg = (int)(c_last - c_first + 1), myseg = (int)(c - c_first);
        if (lane == 0) {
          uint32_t spins = 0;
          while (ld_acquire_gpu(&p.counters[t]) < 4 * nseg) {
            if (++spins > VDD_SPIN_LIMIT) __trap();
          }
        }
        __syncwarp();
        if (warp == 2 && lane == 0) SK_STAMP(6);  // (last) shared tile complete
        const float* wst = p.ws + (size_t)t * p.maxseg * slot_floats;
        const int items = 8 * nch;
        for (int j = myseg + nseg * (warp - 2); j < items; j += nseg * 4) {
          const int hq = j / nch, cc = j - hq * nch;  // hq = h * 4 + quadrant
          const float* src = wst + ((size_t)(hq >> 2) * p.RP + cc * 16) * 128 + (hq & 3) * 32 + lane;
          float v[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = 0.f;
          for (int sgm = 0; sgm < nseg; ++sgm) {
            const float* q = src + (size_t)sgm * slot_floats;
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] += __ldcg(q + (size_t)k * 128);
          }
          sk_emit16(p, t * SK_BM + hq * 32 + lane, cc * 16, v);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SK_STAMP(7);  // exit
  if (warp == 1) tmem_dealloc<SK_TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*SkEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static SkEncodeTiledFn g_sk_encode = nullptr;

static bool skinny_init() {
  static bool done = false;
  if (done) return g_sk_encode != nullptr;
  done = true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return false;
  if (cudaFuncSetAttribute(skinny_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_LIMIT) != cudaSuccess) return false;
  g_sk_encode = reinterpret_cast<SkEncodeTiledFn>(fn);
  return true;
}

static bool sk_tmap(CUtensorMap* m, const void* ptr, int rows, int K, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {SK_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return g_sk_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Geometry shared by the launcher and the workspace sizing in the engine.
struct SkGeom {
  int RP, nkb, tiles, stages, nbuf, maxseg, grid, smem;
  long long units;
  size_t ws_bytes;
};
static SkGeom sk_geom(int R, int N, int K, int num_sms) {
  SkGeom g{};
  g.RP = (R + 15) / 16 * 16;
  if (g.RP < 16) g.RP = 16;
  g.nkb = (K + SK_BK - 1) / SK_BK;
  g.tiles = (N + SK_BM - 1) / SK_BM;
  g.units = (long long)g.tiles * g.nkb;
  g.grid = (int)(g.units < num_sms ? g.units : num_sms);
  const int stage_bytes = SK_A_BYTES + g.RP * 128;
  g.stages = (SK_SMEM_LIMIT - 1024 - 512) / stage_bytes;
  if (g.stages > SK_MAX_STAGES) g.stages = SK_MAX_STAGES;
  g.smem = g.stages * stage_bytes + 1024 + 512;
  g.nbuf = 4 * g.RP <= SK_TMEM_COLS ? 2 : 1;
  const long long min_len = g.units / g.grid;  // shortest CTA range
  g.maxseg = (int)(g.nkb / (min_len > 0 ? min_len : 1)) + 2;
  g.ws_bytes = (size_t)g.tiles * g.maxseg * 2 * g.RP * 128 * 4;
  return g;
}
size_t skinny_gemm_ws_bytes(int R, int N, int K, int num_sms) { return sk_geom(R, N, K, num_sms).ws_bytes; }
int skinny_gemm_tiles(int N) { return (N + SK_BM - 1) / SK_BM; }

cudaError_t launch_skinny_gemm(const SkinnyArgs& a, int num_sms, bool pdl, cudaStream_t stream) {
  if (!skinny_init()) return cudaErrorNotSupported;
  if (a.R <= 0 || a.R > 256 || (a.K % 8) != 0 || ((uintptr_t)a.x & 15) || ((uintptr_t)a.w & 15)) return cudaErrorInvalidValue;
  if (a.epilogue != GEMM_OUT_F32 && a.epilogue != GEMM_OUT_ADD_F32 && a.epilogue != GEMM_OUT_SWIGLU) return cudaErrorInvalidValue;
  if (a.epilogue == GEMM_OUT_SWIGLU && ((a.N & 3) || (a.ld & 1) || ((uintptr_t)a.out_bf16 & 3))) return cudaErrorInvalidValue;
  const SkGeom g = sk_geom(a.R, a.N, a.K, num_sms);
  if (g.stages < 2 || g.ws_bytes > a.ws_bytes || !a.ws || !a.counters) return cudaErrorInvalidValue;
  CUtensorMap tw, tx;
  if (!sk_tmap(&tw, a.w, a.N, a.K, SK_BM) || !sk_tmap(&tx, a.x, a.R, a.K, g.RP)) return cudaErrorInvalidValue;
  SkinnyParams p{};
  p.R = a.R; p.RP = g.RP; p.N = a.N; p.K = a.K; p.mode = a.epilogue;
  p.nkb = g.nkb; p.tiles = g.tiles; p.stages = g.stages; p.nbuf = g.nbuf; p.maxseg = g.maxseg; p.units = g.units;
  p.out_f32 = a.out_f32; p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a.out_bf16); p.ld = a.ld;
  p.ws = a.ws; p.counters = a.counters; p.trace = a.trace;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(g.grid);
  cfg.blockDim = dim3(SK_THREADS);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, skinny_gemm_kernel, tw, tx, p);
}

}  // namespace vdd
