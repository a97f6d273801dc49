// Decode projections for 8 < R <= 256 fused rows (BASELINE configs 3-5: 16-192 rows of x against W[N, K]).
//
// Replaces the cuBLAS fp16 GEMMs HF LlamaDecoderLayer / lm_head issue once per stream per step when the reference decodes
// (llava_llama.py:91-103 entered from vcd_sample.py:108-113,161-185); here all streams of all sequences are the rows of
// ONE skinny GEMM per projection, so the weights cross HBM once per step.
//
// At these row counts a projection is still a weight stream (R < 212 = the bf16 ridge of a B200), so the kernel is built
// like a copy engine with a tensor core attached, not like a square GEMM:
//   * roles are SWAPPED: the weight tile is the UMMA A operand (M = 128 output features, two of them per stage = 256
//     features), the activations are the B operand (N = R rounded up to 16), so no MMA work or TMEM is spent on padding
//     rows and a 32 KB weight tile needs only R*128 bytes of activations from L2 beside it;
//   * stream-K: the (feature tile, k-block) space of a projection is cut into equal contiguous ranges, one persistent CTA
//     per SM, so every SM streams the same number of bytes whatever N/256 is (o_proj has 16 tiles, gate_up 86);
//   * TMA (cp.async.bulk.tensor, 128B swizzle, weights with an evict-first L2 policy) fills a 3-6 stage ring, one thread
//     issues tcgen05.mma into a double-buffered TMEM accumulator, four epilogue warps drain a segment with tcgen05.ld
//     while the next one streams;
//   * a tile whose k-range is shared by several CTAs is finished without a second launch: every sharer parks its fp32
//     partial in an L2-resident workspace and bumps the tile's counter; once all have arrived each sharer sums its share
//     of the tile over the segments IN SEGMENT ORDER (bit-identical run to run, independent of arrival order) and applies
//     the epilogue;
//   * up to FOUR projections run in one launch (o_proj -> gate_up -> down_proj -> next layer's qkv / lm_head), separated
//     by grid barriers: set-up, TMEM allocation and the pipeline fill are paid once, and the weights of the next
//     projection are already streaming into the ring while the previous one is being fixed up (measured stand-alone:
//     set-up 1.3 us + first fill 3 us + drain 2.6 us + fix-up 6-8 us against a 15 us stream, profiles/r1_trace_skinny.json);
//   * RMSNorm never gets a pass of its own: the residual epilogue (RESID_XG) stores resid += acc AND bf16(resid * g) as
//     the next projection's B operand, the row statistic 1/rms(resid) is taken by the otherwise idle epilogue warps while
//     the weights stream, and the consumer's epilogue multiplies its accumulator by it (W (g.x/rms) == (W (g.x)) / rms);
//   * the weight loads of the first ring stages are issued BEFORE griddepcontrol.wait (weights never depend on the
//     previous kernel), the activation loads after it.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..9 = epilogue (TMEM lane quadrant = warp % 4, two warps
// per quadrant taking alternate 16-column chunks).
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace vdd {

constexpr int SK_BM = 256, SK_BK = 64;
constexpr int SK_EPI_WARPS = 8;                       // two per TMEM lane quadrant: they split a segment's column chunks
constexpr int SK_THREADS = 64 + 32 * SK_EPI_WARPS;  // warp 0 = TMA, warp 1 = MMA, then the epilogue warps
constexpr int SK_A_BYTES = SK_BM * SK_BK * 2;  // 32 KB of weights per stage
constexpr int SK_MAX_STAGES = 8;
constexpr int SK_SMEM_LIMIT = 226 * 1024;  // dynamic part; the kernel also holds a few bytes of static shared memory
constexpr int SK_TMEM_COLS = 512;
constexpr int SK_MAX_PHASES = 4;

struct SkPhase {
  int N, K, mode, nkb, tiles, maxseg, active;  // active = CTAs that own units of this projection (min(grid, units))
  long long units;                             // tiles * nkb
  float* out_f32;                              // F32: store; ADD_F32 / RESID_XG: +=
  __nv_bfloat16* out_bf16;                     // SWIGLU: [R, N/2]; RESID_XG: bf16(resid * g) [R, N]
  int64_t ld, ld_bf16;
  const __nv_bfloat16* g;                      // RESID_XG: weight of the NEXT RMSNorm
  const float* rs_x;                           // row scale: rows [R, rs_K] whose 1/rms multiplies this projection's result
  int rs_K;
  float rs_eps;
  float* inv_rms;                              // [R] scratch the row scales are published in
  int* norm_counter;                           // zero on entry; counts published rows
  int* counters;                               // [tiles] zero on entry
};
struct SkChainParams {
  int R, RP, stages, nbuf, nph;
  SkPhase ph[SK_MAX_PHASES];
  float* ws;                  // fp32 partials [tiles][maxseg][2][RP][128], reused by every phase
  unsigned* barrier;          // [nph - 1] zero on entry
  unsigned long long* trace;  // debug: [grid][16] globaltimer stamps, or null
};
struct SkChainMaps {
  CUtensorMap w[SK_MAX_PHASES], x[SK_MAX_PHASES];
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
#define SK_STAMP(slot)                                                \
  do {                                                                \
    if (p.trace) p.trace[(size_t)blockIdx.x * 16 + (slot)] = sk_now(); \
  } while (0)
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * SK_EPI_WARPS) : "memory"); }  // the epilogue warps
// spin (one lane) until *p >= target; bounded so that a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void sk_wait_ge(const int* p, int target) {
  uint32_t spins = 0;
  while (ld_acquire_gpu(p) < target) {
    __nanosleep(40);  // a flag is polled by ONE thread per CTA, politely: pollers share an L2 line with the arriving atomics
    if (++spins > (VDD_SPIN_LIMIT >> 3)) __trap();
  }
}

// This CTA's unit range of a phase.
struct SkRange {
  long long u0;
  int n;
};
__device__ __forceinline__ SkRange sk_range(const SkPhase& ph, int c) {
  SkRange r{0, 0};
  if (c < ph.active) {
    r.u0 = (long long)c * ph.units / ph.active;
    r.n = (int)((long long)(c + 1) * ph.units / ph.active - r.u0);
  }
  return r;
}

// Values v[0..NR) of output feature n (one feature per lane, 32 consecutive features per warp) for rows r0..r0+NR-1.
// Every lane of the warp must call this (SwiGLU pairs features across lanes).
// `old` (when have_old) holds the prefetched current contents of out_f32 for the accumulate modes.
template <int NR>
__device__ __forceinline__ void sk_emit(const SkPhase& ph, int R, int n, int r0, float (&v)[NR], const float (&old)[NR], bool have_old) {
  const int lane = threadIdx.x & 31;
  if (ph.rs_x) {
#pragma unroll
    for (int i = 0; i < NR; ++i)
      if (r0 + i < R) v[i] *= __ldcg(ph.inv_rms + r0 + i);
  }
  if (ph.mode == GEMM_OUT_SWIGLU) {
    // features interleave gate (even) / up (odd); out[r, n/2] = silu(gate) * up, two results packed per 32-bit store
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const float up = __shfl_xor_sync(0xffffffffu, v[i], 1);
      const float hval = v[i] / (1.0f + expf(-v[i])) * up;  // meaningful on even lanes
      const float hnext = __shfl_down_sync(0xffffffffu, hval, 2);
      const int r = r0 + i;
      if ((lane & 3) == 0 && r < R) {
        if (n + 3 < ph.N) {
          *reinterpret_cast<uint32_t*>(ph.out_bf16 + (int64_t)r * ph.ld_bf16 + (n >> 1)) = pack_bf16(hval, hnext);
        } else if (n + 1 < ph.N) {
          ph.out_bf16[(int64_t)r * ph.ld_bf16 + (n >> 1)] = __float2bfloat16_rn(hval);
        }
      }
    }
    return;
  }
  if (n >= ph.N) return;
  if (ph.mode == GEMM_OUT_RESID_XG) {
    const float gn = __bfloat162float(ph.g[n]);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = r0 + i;
      if (r < R) {
        float* o = ph.out_f32 + (int64_t)r * ph.ld + n;
        const float nv = (have_old ? old[i] : __ldcg(o)) + v[i];
        *o = nv;
        ph.out_bf16[(int64_t)r * ph.ld_bf16 + n] = __float2bfloat16_rn(nv * gn);
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int r = r0 + i;
    if (r < R) {
      float* o = ph.out_f32 + (int64_t)r * ph.ld + n;
      *o = ph.mode == GEMM_OUT_ADD_F32 ? (have_old ? old[i] : __ldcg(o)) + v[i] : v[i];
    }
  }
}

// Fix-up of this CTA's share of shared tile t: items of 32 features x NR rows, item j taken when j = myseg + nseg * (ew + 8 i).
// The fix-up is a chain of L2 round trips, so an item is shaped to need ONE: NR shrinks as the number of sharers grows
// (16 rows for <= 4 segments, 8 for <= 8, 4 beyond: 64 partial loads per lane in flight either way), and the current output
// of the accumulate modes is fetched in the same round.  Sums run over segments 0..nseg-1 in order whatever NR is.
template <int NR>
__device__ __forceinline__ void sk_fix_tile(const SkPhase& ph, int R, int RP, int t, int nseg, int myseg, int ew, int lane,
                                            const float* wst, size_t slot_floats) {
  constexpr int PER_ROUND = 64 / NR;
  const int ncr = RP / NR, items = 8 * ncr;
  const bool accum_mode = ph.mode == GEMM_OUT_ADD_F32 || ph.mode == GEMM_OUT_RESID_XG;
  for (int j = myseg + nseg * ew; j < items; j += nseg * SK_EPI_WARPS) {
    const int hq = j / ncr, r0 = (j - hq * ncr) * NR;  // hq = h * 4 + quadrant
    const float* src = wst + ((size_t)(hq >> 2) * RP + r0) * 128 + (hq & 3) * 32 + lane;
    const int n = t * SK_BM + hq * 32 + lane;
    float v[NR], old[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      v[k] = 0.f;
      old[k] = 0.f;
    }
    if (accum_mode && n < ph.N) {
#pragma unroll
      for (int k = 0; k < NR; ++k)
        if (r0 + k < R) old[k] = __ldcg(ph.out_f32 + (int64_t)(r0 + k) * ph.ld + n);
    }
    for (int sgm = 0; sgm < nseg; sgm += PER_ROUND) {
      float a[PER_ROUND][NR];
#pragma unroll
      for (int z = 0; z < PER_ROUND; ++z) {
        if (sgm + z < nseg) {
#pragma unroll
          for (int k = 0; k < NR; ++k) a[z][k] = __ldcg(src + (size_t)(sgm + z) * slot_floats + (size_t)k * 128);
        } else {
#pragma unroll
          for (int k = 0; k < NR; ++k) a[z][k] = 0.f;
        }
      }
#pragma unroll
      for (int z = 0; z < PER_ROUND; ++z)
#pragma unroll
        for (int k = 0; k < NR; ++k) v[k] += a[z][k];
    }
    sk_emit<NR>(ph, R, n, r0, v, old, accum_mode);
  }
}

__global__ void __launch_bounds__(SK_THREADS, 1)
skinny_chain_kernel(const __grid_constant__ SkChainMaps maps, const __grid_constant__ SkChainParams p) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ float red_s[SK_EPI_WARPS];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int stage_bytes = SK_A_BYTES + p.RP * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + SK_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + SK_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* phase_bar = tempty_bar + 2;  // [SK_MAX_PHASES] grid barrier q-1 observed (producer -> epilogue warps), used once each
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(phase_bar + SK_MAX_PHASES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, c = blockIdx.x;

  if (threadIdx.x == 0) SK_STAMP(0);  // kernel entry
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], SK_EPI_WARPS);  // one arrival per epilogue warp
    }
    for (int q = 0; q < SK_MAX_PHASES; ++q) mbar_init(&phase_bar[q], 1);
    fence_barrier_init();
    // The first projection's first weight tiles are requested right here, before TMEM allocation and the CTA barrier:
    // they depend on nothing (not on the previous kernel either), and 1-2 us of set-up would otherwise delay the stream.
    const SkPhase& ph = p.ph[0];
    const SkRange rg = sk_range(ph, c);
    if (rg.n > 0) tma_prefetch_desc(&maps.w[0]);
    const uint64_t pol = policy_evict_first();
    const int npre = rg.n < p.stages ? rg.n : p.stages;
    for (int i = 0; i < npre; ++i) {
      const long long u = rg.u0 + i;
      const int t = (int)(u / ph.nkb), kb = (int)(u % ph.nkb);
      mbar_arrive_expect_tx(&full_bar[i], stage_bytes);
      tma_load_2d_hint(smem + (size_t)i * stage_bytes, &maps.w[0], &full_bar[i], kb * SK_BK, t * SK_BM, pol);
    }
    SK_STAMP(1);  // weights of the first stages requested
  }
  if (warp == 1) tmem_alloc<SK_TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: one continuous stream over the unit ranges of all phases; `gi` = ring position
      const uint64_t pol = policy_evict_first();
      int gi = 0;
      for (int q = 0; q < p.nph; ++q) {
        const SkPhase& ph = p.ph[q];
        const SkRange rg = sk_range(ph, c);
        const CUtensorMap* tw = &maps.w[q];
        const CUtensorMap* tx = &maps.x[q];
        if (rg.n > 0) {
          if (q > 0) tma_prefetch_desc(tw);
          tma_prefetch_desc(tx);
        }
        // weights first: they depend neither on the previous kernel nor on the previous phase (phase 0: already requested)
        const int npre = rg.n < p.stages ? rg.n : p.stages;
        for (int i = 0; i < npre && q > 0; ++i) {
          const int s = (gi + i) % p.stages;
          const long long u = rg.u0 + i;
          const int t = (int)(u / ph.nkb), kb = (int)(u % ph.nkb);
          mbar_wait(&empty_bar[s], (((gi + i) / p.stages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
          tma_load_2d_hint(smem + (size_t)s * stage_bytes, tw, &full_bar[s], kb * SK_BK, t * SK_BM, pol);
        }
        if (q == 0) {
          pdl_wait();
          SK_STAMP(2);  // previous kernel complete
        } else {
          sk_wait_ge(reinterpret_cast<const int*>(p.barrier + (q - 1)), G);  // every CTA has finished phase q-1
          if (q == 1) SK_STAMP(7);  // grid barrier after the first projection observed
          mbar_arrive(&phase_bar[q]);  // tell this CTA's epilogue warps (they read the residual for the row scales)
          fence_proxy_async_all();  // its generic-proxy stores (our B operand) -> visible to the TMA reads below
        }
        for (int i = 0; i < npre; ++i) {
          const int s = (gi + i) % p.stages;
          const int kb = (int)((rg.u0 + i) % ph.nkb);
          tma_load_2d(smem + (size_t)s * stage_bytes + SK_A_BYTES, tx, &full_bar[s], kb * SK_BK, 0);
        }
        for (int i = npre; i < rg.n; ++i) {
          const int s = (gi + i) % p.stages;
          const long long u = rg.u0 + i;
          const int t = (int)(u / ph.nkb), kb = (int)(u % ph.nkb);
          mbar_wait(&empty_bar[s], (((gi + i) / p.stages) & 1) ^ 1);
          unsigned char* sa = smem + (size_t)s * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
          tma_load_2d_hint(sa, tw, &full_bar[s], kb * SK_BK, t * SK_BM, pol);
          tma_load_2d(sa + SK_A_BYTES, tx, &full_bar[s], kb * SK_BK, 0);
        }
        gi += rg.n;
      }
    }
    __syncwarp();  // reconverge before the CTA-wide barrier at the end
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: per stage 4 k-steps x 2 feature halves of UMMA 128 x RP x 16
      const uint32_t idesc = umma_idesc_bf16(128, p.RP);
      int gi = 0, seg = 0, buf = 0;
      for (int q = 0; q < p.nph; ++q) {
        const SkPhase& ph = p.ph[q];
        const SkRange rg = sk_range(ph, c);
        for (int i = 0; i < rg.n; ++i, ++gi) {
          const int kb = (int)((rg.u0 + i) % ph.nkb);
          const bool seg_start = i == 0 || kb == 0;
          if (seg_start) {
            buf = seg % p.nbuf;
            mbar_wait(&tempty_bar[buf], ((seg / p.nbuf) & 1) ^ 1);  // the epilogue has drained this accumulator
            tcgen05_fence_after();
          }
          const int s = gi % p.stages;
          mbar_wait(&full_bar[s], (gi / p.stages) & 1);
          if (gi == 0) SK_STAMP(3);  // first stage landed
          if (i == 0) SK_STAMP(8 + 2 * q);  // phase q: first stage (weights AND activations) landed
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
          if (i == rg.n - 1) SK_STAMP(9 + 2 * q);  // phase q: last MMA issued
          if (i == rg.n - 1 || kb == ph.nkb - 1) {
            umma_commit(&tfull_bar[buf]);
            ++seg;
          }
        }
      }
      SK_STAMP(4);  // last MMA issued
    }
    __syncwarp();
  } else {
    // ===== epilogue warps
    const int quad = warp & 3, ew = warp - 2, et = threadIdx.x - 64;  // ew / et: warp / thread index within the epilogue group
    const int chalf = ew >> 2;  // which of the quadrant's two warps: takes column chunks cc = chalf, chalf + 2, ...
    const int nch = p.RP >> 4;
    const size_t slot_floats = (size_t)2 * p.RP * 128;
    pdl_wait();  // outputs (and the residual) belong to earlier kernels until here
    pdl_launch_dependents();
    int seg = 0;
    for (int q = 0; q < p.nph; ++q) {
      const SkPhase& ph = p.ph[q];
      const SkRange rg = sk_range(ph, c);
      if (q > 0 && ph.rs_x) {
        // the rows whose 1/rms scales this projection were completed by phase q-1
        mbar_wait(&phase_bar[q], 0);
      }
      if (ph.rs_x) {
        // ---- row statistics, while the weights stream: rows c, c+G, ... of rs_x -> inv_rms (HF LlamaRMSNorm, fp32)
        for (int r = c; r < p.R; r += G) {
          const float* xr = ph.rs_x + (size_t)r * ph.rs_K;
          float ss = 0.f;
          for (int k = et * 4; k < ph.rs_K; k += 32 * SK_EPI_WARPS * 4) {
            const float4 v4 = __ldcg(reinterpret_cast<const float4*>(xr + k));
            ss += v4.x * v4.x + v4.y * v4.y + v4.z * v4.z + v4.w * v4.w;
          }
          ss = warp_sum(ss);
          epi_bar_sync();  // red_s free again
          if (lane == 0) red_s[ew] = ss;
          epi_bar_sync();
          if (et == 0) {
            float tot = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < SK_EPI_WARPS; ++w8) tot += red_s[w8];
            ph.inv_rms[r] = rsqrtf(tot / (float)ph.rs_K + ph.rs_eps);
            __threadfence();
            atomicAdd(ph.norm_counter, 1);
          }
        }
      }
      if (ph.rs_x) {  // every row scale has been published (they were taken while the weights streamed)
        if (et == 0) sk_wait_ge(ph.norm_counter, p.R);
        epi_bar_sync();
      }
      // ---- phase 1: drain every segment of this CTA; whole tiles go straight out, shared tiles park a partial
      {
        int i = 0;
        while (i < rg.n) {
          const long long u = rg.u0 + i;
          const int t = (int)(u / ph.nkb), kb0 = (int)(u % ph.nkb);
          const int len = min(rg.n - i, ph.nkb - kb0);
          const bool whole = kb0 == 0 && len == ph.nkb;
          const int buf = seg % p.nbuf;
          mbar_wait(&tfull_bar[buf], (seg / p.nbuf) & 1);
          tcgen05_fence_after();
          int myseg = 0;
          if (!whole) {
            const long long s_lo = (long long)t * ph.nkb;
            const long long c_first = ((s_lo + 1) * ph.active + ph.units - 1) / ph.units - 1;
            myseg = (int)(c - c_first);
          }
          float* wsp = p.ws + ((size_t)t * ph.maxseg + myseg) * slot_floats;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            const int n = t * SK_BM + h * 128 + quad * 32 + lane;
            const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 2 + h) * p.RP;
            constexpr int CSTEP = SK_EPI_WARPS / 4;
#pragma unroll 1
            for (int cc0 = chalf; cc0 < nch; cc0 += 4 * CSTEP) {  // up to four 16-column chunks per TMEM round trip
              uint32_t rr[4][16];
#pragma unroll
              for (int z = 0; z < 4; ++z)
                if (cc0 + z * CSTEP < nch) tmem_ld16(tcol + (cc0 + z * CSTEP) * 16, rr[z]);
              tmem_ld_wait();
#pragma unroll
              for (int z = 0; z < 4; ++z) {
                const int cc = cc0 + z * CSTEP;
                if (cc < nch) {
                  float v[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[z][j]);
                  if (whole) {
                    sk_emit<16>(ph, p.R, n, cc * 16, v, v, false);
                  } else {
                    float* w = wsp + ((size_t)h * p.RP + cc * 16) * 128 + quad * 32 + lane;
#pragma unroll
                    for (int j = 0; j < 16; ++j) w[(size_t)j * 128] = v[j];
                  }
                }
              }
            }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[buf]);
          if (!whole) {
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(&ph.counters[t], 1);
          }
          i += len;
          ++seg;
        }
      }
      if (q == 0 && warp == 2 && lane == 0) SK_STAMP(5);  // first projection: all segments drained
      // ---- phase 2: finish the shared tiles.  Sharer `myseg` of `nseg` takes items myseg, myseg + nseg, ... of the tile's
      //      8 * RP/16 items (32 features x 16 rows), its epilogue warps interleaved; sums run over segments 0..nseg-1 in order.
      {
        int i = 0;
        while (i < rg.n) {
          const long long u = rg.u0 + i;
          const int t = (int)(u / ph.nkb), kb0 = (int)(u % ph.nkb);
          const int len = min(rg.n - i, ph.nkb - kb0);
          i += len;
          if (kb0 == 0 && len == ph.nkb) continue;
          const long long s_lo = (long long)t * ph.nkb, s_hi = s_lo + ph.nkb;
          const long long c_first = ((s_lo + 1) * ph.active + ph.units - 1) / ph.units - 1;
          const long long c_last = (s_hi * ph.active + ph.units - 1) / ph.units - 1;
          const int nseg = (int)(c_last - c_first + 1), myseg = (int)(c - c_first);
          if (et == 0) sk_wait_ge(&ph.counters[t], SK_EPI_WARPS * nseg);
          epi_bar_sync();
          const float* wst = p.ws + (size_t)t * ph.maxseg * slot_floats;
          if (nseg <= 4) sk_fix_tile<16>(ph, p.R, p.RP, t, nseg, myseg, ew, lane, wst, slot_floats);
          else if (nseg <= 8) sk_fix_tile<8>(ph, p.R, p.RP, t, nseg, myseg, ew, lane, wst, slot_floats);
          else sk_fix_tile<4>(ph, p.R, p.RP, t, nseg, myseg, ew, lane, wst, slot_floats);
        }
      }
      // ---- this CTA's part of projection q is in memory: release the next phase on every CTA
      if (q + 1 < p.nph) {
        fence_proxy_async_all();  // our generic stores are the next phase's TMA (async proxy) operand
        epi_bar_sync();
        if (et == 0) {
          __threadfence();
          if (q == 0) SK_STAMP(6);  // first projection: this CTA's share fixed up, arriving at the grid barrier
          atomicAdd(p.barrier + q, 1u);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
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
  if (cudaFuncSetAttribute(skinny_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_LIMIT) != cudaSuccess) return false;
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

// Geometry of one projection on `grid` persistent CTAs (shared by the launcher and the workspace sizing in the engine).
struct SkGeom {
  int RP, nkb, tiles, maxseg, active;
  long long units;
  size_t ws_bytes;
};
static SkGeom sk_geom(int R, int N, int K, int grid) {
  SkGeom g{};
  g.RP = (R + 15) / 16 * 16;
  if (g.RP < 16) g.RP = 16;
  g.nkb = (K + SK_BK - 1) / SK_BK;
  g.tiles = (N + SK_BM - 1) / SK_BM;
  g.units = (long long)g.tiles * g.nkb;
  g.active = (int)(g.units < grid ? g.units : grid);
  const long long min_len = g.units / g.active;  // shortest CTA range
  g.maxseg = (int)(g.nkb / (min_len > 0 ? min_len : 1)) + 2;
  g.ws_bytes = (size_t)g.tiles * g.maxseg * 2 * g.RP * 128 * 4;
  return g;
}
size_t skinny_gemm_ws_bytes(int R, int N, int K, int num_sms) {
  // the grid of a chain is min(num_sms, largest projection's units); a smaller grid only lowers maxseg
  return sk_geom(R, N, K, num_sms).ws_bytes;
}
int skinny_gemm_tiles(int N) { return (N + SK_BM - 1) / SK_BM; }

cudaError_t launch_skinny_chain(const SkinnyChain& ch, int num_sms, bool pdl, cudaStream_t stream) {
  if (!skinny_init()) return cudaErrorNotSupported;
  if (ch.n < 1 || ch.n > SK_MAX_PHASES || ch.R <= 0 || ch.R > 256 || !ch.ws) return cudaErrorInvalidValue;
  if (ch.n > 1 && !ch.barrier) return cudaErrorInvalidValue;
  SkChainParams p{};
  SkChainMaps maps{};
  p.R = ch.R;
  p.RP = (ch.R + 15) / 16 * 16;
  const int stage_bytes = SK_A_BYTES + p.RP * 128;
  p.stages = (SK_SMEM_LIMIT - 1024 - 512) / stage_bytes;
  if (p.stages > SK_MAX_STAGES) p.stages = SK_MAX_STAGES;
  if (p.stages < 2) return cudaErrorInvalidValue;
  p.nbuf = 4 * p.RP <= SK_TMEM_COLS ? 2 : 1;
  p.nph = ch.n;
  long long max_units = 0;
  for (int q = 0; q < ch.n; ++q) {
    const SkinnyArgs& a = ch.ph[q];
    const long long u = (long long)((a.N + SK_BM - 1) / SK_BM) * ((a.K + SK_BK - 1) / SK_BK);
    if (u > max_units) max_units = u;
  }
  const int grid = (int)(max_units < num_sms ? max_units : num_sms);
  for (int q = 0; q < ch.n; ++q) {
    const SkinnyArgs& a = ch.ph[q];
    if ((a.K % 8) != 0 || ((uintptr_t)a.x & 15) || ((uintptr_t)a.w & 15) || !a.counters) return cudaErrorInvalidValue;
    if (a.epilogue != GEMM_OUT_F32 && a.epilogue != GEMM_OUT_ADD_F32 && a.epilogue != GEMM_OUT_SWIGLU && a.epilogue != GEMM_OUT_RESID_XG)
      return cudaErrorInvalidValue;
    if (a.epilogue == GEMM_OUT_SWIGLU && ((a.N & 3) || (a.ld_bf16 & 1) || ((uintptr_t)a.out_bf16 & 3))) return cudaErrorInvalidValue;
    if (a.epilogue == GEMM_OUT_RESID_XG && (!a.g || !a.out_bf16 || !a.out_f32)) return cudaErrorInvalidValue;
    if (a.rs_x && (!a.inv_rms || !a.norm_counter || (a.rs_K & 3) || ((uintptr_t)a.rs_x & 15))) return cudaErrorInvalidValue;
    const SkGeom g = sk_geom(ch.R, a.N, a.K, grid);
    if (g.ws_bytes > ch.ws_bytes) return cudaErrorInvalidValue;
    if (!sk_tmap(&maps.w[q], a.w, a.N, a.K, SK_BM) || !sk_tmap(&maps.x[q], a.x, ch.R, a.K, p.RP)) return cudaErrorInvalidValue;
    SkPhase& ph = p.ph[q];
    ph.N = a.N; ph.K = a.K; ph.mode = a.epilogue; ph.nkb = g.nkb; ph.tiles = g.tiles; ph.maxseg = g.maxseg; ph.active = g.active;
    ph.units = g.units;
    ph.out_f32 = a.out_f32; ph.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a.out_bf16); ph.ld = a.ld; ph.ld_bf16 = a.ld_bf16;
    ph.g = reinterpret_cast<const __nv_bfloat16*>(a.g);
    ph.rs_x = a.rs_x; ph.rs_K = a.rs_K; ph.rs_eps = a.rs_eps; ph.inv_rms = a.inv_rms; ph.norm_counter = a.norm_counter;
    ph.counters = a.counters;
  }
  p.ws = ch.ws; p.barrier = ch.barrier; p.trace = ch.trace;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(SK_THREADS);
  cfg.dynamicSmemBytes = p.stages * stage_bytes + 1024 + 512;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, skinny_chain_kernel, maps, p);
}

}  // namespace vdd
