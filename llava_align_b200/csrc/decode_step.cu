// ONE persistent launch for a whole decode step of R <= 4 fused rows (all debias streams of a sequence):
//
//   qkv_0 | for every layer: [ attention | o_proj | gate_up | down_proj | qkv of the next layer (or lm_head) ]
//
// i.e. what HF LlamaModel + lm_head do for one token (reference: entered once PER STREAM per token from
// experiments/llava/model/language_model/llava_llama.py:91-103 by the sampling loop vcd_utils/vcd_sample.py:109,163,178),
// for all streams at once, with the weights streamed ONCE.  161 dependent phases (7B) separated by grid barriers.
//
// Why one launch: the step is a chain of small HBM-bound stages; what the previous design (one gemv_kernel launch of four
// projections per layer + one attention launch per layer) lost was the time between stages -- the weight stream stopped
// while CTAs drained, met at a barrier and staged the next activations (profiles/r1_trace_fused_chain.json: ~17 us of a
// 79 us layer), and the attention launch moved 17 MB in 11 us.  Here:
//   * a dedicated PRODUCER WARP per CTA streams the weights of ALL phases, in order, with cp.async.bulk (TMA bulk copies:
//     8 x 512 contiguous bytes per 4 KB step, completion on an mbarrier) into per-warp rings of shared memory.  It never
//     touches activations, so it never waits for a phase boundary: while the consumer warps reduce, meet the other CTAs
//     and normalise the next input, the rings fill with the next projection's weights (up to ~140 KB per SM in flight or
//     landed), and they fill across the attention phase and across layers as well.  The bulk copies also keep the LSU /
//     L1 request queue free: the activation loads of a prologue no longer queue behind 128 KB of cp.async requests.
//   * CONSUMER WARPS (mma.sync m16n8k16 with swapped roles exactly as in decode_gemv.cu: the R activation rows are the M
//     side, 8 weight rows the N side, a lane's 16-byte piece of the ring IS its B fragment) wait on a slot's "full"
//     mbarrier, multiply, release the slot through its "empty" mbarrier.  Task -> warp assignment is static round-robin;
//     the K split of a projection depends only on (N, K, grid), and the RMSNorm statistics are reduced over a fixed set of
//     256 virtual lanes, so a row's bits do not depend on the ring geometry or on which other rows are fused beside it.
//   * ATTENTION is a phase of the same launch: (row, head, 64-token split) items are spread over groups of four consumer
//     warps, RoPE + KV append + split merge as in decode_attn.cu; the producer keeps prefetching o_proj meanwhile.
//   * Tensor parallel: the row-parallel epilogues push their partial band to every rank's inbox over NVLink and the next
//     phase reduces it after a cross-GPU arrival counter (same protocol as decode_gemv.cu), all inside this launch.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vdd {

constexpr int ST_RG = 8;                       // weight rows per step (the N side of m16n8k16)
constexpr int ST_UNIT = 256;                   // k elements per step
// A slot holds 8 weight rows x 512 B.  Even rows sit at 0, 512, 1024, 1536, odd rows 64 B further along behind them (2112 + ...):
// the two rows a quarter-warp reads (g = 2j, 2j + 1) then differ by 64 B modulo 128, which makes the 16-byte fragment reads
// bank-conflict free without a swizzle (a bulk copy cannot swizzle) at 64 B of padding per slot.
constexpr int ST_SLOT_BYTES = 4160;
__host__ __device__ __forceinline__ int st_row_off(int g) { return (g >> 1) * 512 + (g & 1) * 2112; }
constexpr int ST_MAX_NW = 8;                   // consumer warps (the producer is warp index nw)
constexpr int ST_MAX_D = 8;                    // ring slots per consumer warp
constexpr int ST_XPAD = 32;                    // bf16 elements of padding per x row when K % 64 == 0 (rows land 64 B apart modulo 128)
constexpr int ST_VL = 256;                     // virtual lanes of the RMSNorm reduction (independent of the thread count)
constexpr int ST_SMEM_BUDGET = 222 * 1024;     // 227 KB per CTA minus the kernel's static shared memory
constexpr int ST_ATTN_GROUP = 128;             // threads per attention item
constexpr int ST_ATTN_SCRATCH = 128 * 4 + 128 * 2 * 2 + 8 * 128 * 4 + 8 * 2 * 4 + 16;  // per group: q, knew, vnew, o_s, ml_s, flag

struct StepPlan {
  int n0, n1, ngroups, ks, ntasks, nunits;
};
// identical to decode_gemv.cu's make_plan: the K split (ks) is a function of N, K and the grid only
__host__ __device__ __forceinline__ StepPlan step_plan(int N, int K, int cta, int nctas) {
  StepPlan p;
  const int pairs = N >> 1;
  p.n0 = 2 * (int)(((long long)cta * pairs) / nctas);
  p.n1 = 2 * (int)(((long long)(cta + 1) * pairs) / nctas);
  const int nrows = p.n1 - p.n0;
  p.ngroups = (nrows + ST_RG - 1) / ST_RG;
  p.nunits = (K + ST_UNIT - 1) / ST_UNIT;
  int ks = p.ngroups > 0 ? (8 * 16 + p.ngroups - 1) / p.ngroups : 1;
  const int ksmax = p.nunits / 2 > 1 ? p.nunits / 2 : 1;
  ks = ks < 1 ? 1 : (ks > ksmax ? ksmax : ks);
  p.ks = ks;
  p.ntasks = p.ngroups * ks;
  return p;
}

// ------------------------------------------------------------------------------------------------ small PTX helpers
__device__ __forceinline__ void st_mma(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%5}, {%7,%8}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(0u), "r"(a2), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(bar), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void mbar_init_a(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned long long st_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Bounded waits: a protocol bug becomes a trap (error return to the host) within seconds instead of a hung GPU box.
struct SpinGuard {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  __device__ __forceinline__ void tick(unsigned long long limit_ns) {
    if ((++spins & 0x3ffu) == 0) {
      const unsigned long long now = st_gtime();
      if (!t0) t0 = now;
      else if (now - t0 > limit_ns) __trap();
    }
  }
};
constexpr unsigned long long ST_LOCAL_LIMIT_NS = 4000000000ull;   // one GPU: nothing legitimately takes seconds
constexpr unsigned long long ST_PEER_LIMIT_NS = 30000000000ull;   // a peer rank may still be instantiating its graph
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity, unsigned long long limit_ns) {
  SpinGuard sg;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    sg.tick(limit_ns);
  }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ unsigned st_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned st_ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------ kernel state
struct StepShared {
  uint64_t full[ST_MAX_NW][ST_MAX_D];
  uint64_t empty[ST_MAX_NW][ST_MAX_D];
  uint64_t xbar;          // completion of the bf16 activation bulk copies of a prologue
  float rstd[4];
  int attn_last[2];
};

struct ConsumerCtx {
  int nw, nthreads, tid, warp, lane, g, t, D;
  const unsigned char* ring;  // this warp's slot 0
  uint32_t full0, empty0;  // shared addresses of this warp's barrier arrays
  int slot;             // next slot to consume
  uint32_t parity;      // its phase parity
  uint32_t xparity;     // phase parity of xbar
  unsigned long long limit;  // spin bound of the local waits
};

#define ST_TRACE(ev)                                                                                         \
  do {                                                                                                       \
    if (trace && c.tid == 0 && ph < 64) trace[((size_t)blockIdx.x * 64 + ph) * 8 + (ev)] = st_gtime();    \
  } while (0)

// ------------------------------------------------------------------------------------------------ producer warp
// Streams the weights of every projection phase, in phase order, task order, unit order -- exactly the order in which the
// consumer warps will ask for them (task i belongs to consumer warp i % nw).  Lane q < 8 copies row q of a step.
__device__ __forceinline__ void producer_loop(const StepArgs& sa, StepShared* sh, uint32_t ring_base, int nw, int D) {
  const int lane = threadIdx.x & 31;
  uint32_t cnt = 0;  // lane w < nw: steps issued so far into ring w
  const uint64_t policy = policy_evict_first();
  const unsigned long long limit = sa.tpc.tp > 1 ? ST_PEER_LIMIT_NS : ST_LOCAL_LIMIT_NS;  // consumers may be waiting for a peer GPU
  for (int ph = 0; ph < sa.nph; ++ph) {
    const StepPhase& P = sa.phases[ph];
    if (P.kind != STEP_GEMV) continue;
    const int N = P.g.N, K = P.g.K;
    const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(P.g.w);
    const StepPlan p = step_plan(N, K, blockIdx.x, gridDim.x);
    int gi = 0, ji = 0;
    for (int task = 0; task < p.ntasks; ++task) {
      const int w = task % nw;
      const int u0 = (ji * p.nunits) / p.ks, u1 = ((ji + 1) * p.nunits) / p.ks;
      const int row = p.n0 + gi * ST_RG + lane;  // lanes >= 8 unused
      const int nvalid = min(ST_RG, p.n1 - (p.n0 + gi * ST_RG));
      for (int u = u0; u < u1; ++u) {
        const uint32_t cw = __shfl_sync(0xffffffffu, cnt, w);
        const int slot = (int)(cw % (uint32_t)D);
        const uint32_t par = (cw / (uint32_t)D) & 1u;
        const uint32_t fb = smem_u32(&sh->full[w][slot]), eb = smem_u32(&sh->empty[w][slot]);
        mbar_wait_a(eb, par ^ 1u, limit);  // first pass over the ring: a fresh barrier's "previous phase" counts as complete
        const int kb = K - u * ST_UNIT;
        const uint32_t bytes = (uint32_t)(kb >= ST_UNIT ? ST_UNIT * 2 : kb * 2);
        if (lane == 0) mbar_expect_a(fb, bytes * (uint32_t)nvalid);
        __syncwarp();
        if (lane < nvalid)
          bulk_g2s_hint(ring_base + (uint32_t)((w * D + slot) * ST_SLOT_BYTES + st_row_off(lane)),
                        W + (size_t)row * K + (size_t)u * ST_UNIT, bytes, fb, policy);
        if (lane == w) cnt++;
      }
      if (++ji == p.ks) { ji = 0; ++gi; }
    }
  }
}

// ------------------------------------------------------------------------------------------------ consumer: barriers
__device__ __forceinline__ void cons_sync(const ConsumerCtx& c) { named_sync(1, c.nthreads); }

__device__ __forceinline__ void grid_barrier(const ConsumerCtx& c, unsigned* counter, unsigned target) {
  cons_sync(c);  // every consumer thread's global stores of this phase are issued
  if (c.tid == 0) {
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    SpinGuard sg;
    while (st_ld_acquire(counter) < target) sg.tick(c.limit);
  }
  cons_sync(c);
}
__device__ __forceinline__ void tp_arrive_all(const ConsumerCtx& c, const TpComm& t, int cidx) {
  cons_sync(c);  // every thread's peer stores of this phase are issued
  if (c.tid == 0) {
    __threadfence_system();
    for (int k = 0; k < t.tp; ++k) {
      unsigned* p = reinterpret_cast<unsigned*>(t.peers[k]) + cidx;
      asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
    }
  }
}
__device__ __forceinline__ void tp_wait_all(const ConsumerCtx& c, const TpComm& t, int cidx, unsigned target) {
  if (c.tid == 0) {
    const unsigned* p = reinterpret_cast<const unsigned*>(t.peers[t.rank]) + cidx;
    SpinGuard sg;
    while (st_ld_acquire_sys(p) < target) sg.tick(ST_PEER_LIMIT_NS);
    __threadfence_system();
  }
  cons_sync(c);
}

// ------------------------------------------------------------------------------------------------ consumer: one projection
template <int R>
__device__ __forceinline__ void gemv_phase(const StepArgs& sa, const StepPhase& P, ConsumerCtx& c, StepShared* sh, unsigned char* smem_x,
                                           float* part, unsigned long long* trace, int ph) {
  const GemvArgs& a = P.g;
  const int K = a.K;
  const int XS = K + ((K & 63) == 0 ? ST_XPAD : 0);  // row pitch = 64 B modulo 128 either way (K % 32 == 0)
  const int tid = c.tid, nthreads = c.nthreads;
  const int g = c.g, t = c.t;
  const int epi = a.epilogue;
  const StepPlan p = step_plan(a.N, K, blockIdx.x, gridDim.x);
  const int nrows = p.n1 - p.n0;
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_x);

  // residual values this thread will add to (requested now, needed after the stream loop)
  float resid_old = 0.f;
  if (epi == GEMV_OUT_ADD_RESID && tid < nrows * R) {
    const int row = tid / R, r = tid - row * R;
    resid_old = __ldcg(a.out_f32 + (size_t)r * a.ld_out + p.n0 + row);
  }
  // ---- prologue: x (bf16) into shared memory
  if (a.x_resid) {
    // RMSNorm over 256 VIRTUAL lanes: lane v owns the float4 chunks v, v + 256, ...; a thread serves lanes tid, tid + nthreads, ...
    // Partials meet in shared memory and are summed in one fixed order, so the statistics do not depend on the thread count.
    const __nv_bfloat16* nw = reinterpret_cast<const __nv_bfloat16*>(a.norm_w);
    const int nchunk = K >> 2;
    float* red = part;  // [ST_VL][4] partials; the partial-sum area is idle until the stream loop
    constexpr int PASS = 5;  // float4 per row a thread can keep in registers (K <= 5120 with 256 threads)
    if (nthreads == ST_VL && nchunk <= PASS * ST_VL) {
      // one virtual lane per thread: the row pieces stay in registers between the statistics and the normalisation
      float4 xv[R][PASS];
#pragma unroll
      for (int q = 0; q < PASS; ++q) {
        const int ch = tid + q * ST_VL;
#pragma unroll
        for (int r = 0; r < R; ++r)
          xv[r][q] = ch < nchunk ? __ldcg(reinterpret_cast<const float4*>(a.x_resid + (size_t)r * K) + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < PASS; ++q) ss += xv[r][q].x * xv[r][q].x + xv[r][q].y * xv[r][q].y + xv[r][q].z * xv[r][q].z + xv[r][q].w * xv[r][q].w;
        red[tid * 4 + r] = ss;
      }
      cons_sync(c);
      if (c.warp == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < ST_VL / 32; ++j) s += red[(c.lane + 32 * j) * 4 + r];
          s = warp_sum(s);
          if (c.lane == 0) sh->rstd[r] = rsqrtf(s / (float)K + a.eps);
        }
      }
      cons_sync(c);
      float rstd[R];
#pragma unroll
      for (int r = 0; r < R; ++r) rstd[r] = sh->rstd[r];
#pragma unroll
      for (int q = 0; q < PASS; ++q) {
        const int ch = tid + q * ST_VL;
        if (ch < nchunk) {
          const uint2 wv = __ldg(reinterpret_cast<const uint2*>(nw) + ch);
          const float w0 = bf16lo(wv.x), w1 = bf16hi(wv.x), w2 = bf16lo(wv.y), w3 = bf16hi(wv.y);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            uint2 o;
            o.x = pack_bf16(xv[r][q].x * rstd[r] * w0, xv[r][q].y * rstd[r] * w1);
            o.y = pack_bf16(xv[r][q].z * rstd[r] * w2, xv[r][q].w * rstd[r] * w3);
            *reinterpret_cast<uint2*>(xs + (size_t)r * XS + ch * 4) = o;
          }
        }
      }
    } else {
      // any thread count / row width: same virtual lanes, same summation order (=> same bits), two passes over L2
      for (int v = tid; v < ST_VL; v += nthreads) {
        float ss[R];
#pragma unroll
        for (int r = 0; r < R; ++r) ss[r] = 0.f;
        for (int ch = v; ch < nchunk; ch += ST_VL) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float4 xv = __ldcg(reinterpret_cast<const float4*>(a.x_resid + (size_t)r * K) + ch);
            ss[r] += xv.x * xv.x + xv.y * xv.y + xv.z * xv.z + xv.w * xv.w;
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) red[v * 4 + r] = ss[r];
      }
      cons_sync(c);
      if (c.warp == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < ST_VL / 32; ++j) s += red[(c.lane + 32 * j) * 4 + r];
          s = warp_sum(s);
          if (c.lane == 0) sh->rstd[r] = rsqrtf(s / (float)K + a.eps);
        }
      }
      cons_sync(c);
      float rstd[R];
#pragma unroll
      for (int r = 0; r < R; ++r) rstd[r] = sh->rstd[r];
      for (int ch = tid; ch < nchunk; ch += nthreads) {
        const uint2 wv = __ldg(reinterpret_cast<const uint2*>(nw) + ch);
        const float w0 = bf16lo(wv.x), w1 = bf16hi(wv.x), w2 = bf16lo(wv.y), w3 = bf16hi(wv.y);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 xv = __ldcg(reinterpret_cast<const float4*>(a.x_resid + (size_t)r * K) + ch);
          uint2 o;
          o.x = pack_bf16(xv.x * rstd[r] * w0, xv.y * rstd[r] * w1);
          o.y = pack_bf16(xv.z * rstd[r] * w2, xv.w * rstd[r] * w3);
          *reinterpret_cast<uint2*>(xs + (size_t)r * XS + ch * 4) = o;
        }
      }
    }
  } else {
    // bf16 x produced by other CTAs of this launch: one bulk copy per row, completion on xbar
    if (tid == 0) {
      fence_proxy_async_all();  // the rows were written through the generic proxy (other SMs), read here by the async proxy
      const uint32_t xb = smem_u32(&sh->xbar);
      mbar_expect_a(xb, (uint32_t)(R * K * 2));
      const __nv_bfloat16* xg = reinterpret_cast<const __nv_bfloat16*>(a.x_bf16);
#pragma unroll
      for (int r = 0; r < R; ++r) bulk_g2s(smem_u32(xs + (size_t)r * XS), xg + (size_t)r * K, (uint32_t)(K * 2), xb);
    }
    mbar_wait_a(smem_u32(&sh->xbar), c.xparity, c.limit);
    c.xparity ^= 1u;
  }
  cons_sync(c);
  ST_TRACE(2);

  // ---- stream loop: this warp's tasks (task i -> warp i % nw), lane (g, t) accumulates y[r = g][features 2t, 2t+1 of the group]
  {
    const __nv_bfloat16* xrow = xs + (size_t)(g < R ? g : 0) * XS + t * 8;
    const uint32_t frag_off = (uint32_t)st_row_off(g) + (uint32_t)t * 16;
    for (int task = c.warp; task < p.ntasks; task += c.nw) {
      const int gi = task / p.ks, ji = task - gi * p.ks;
      const int u0 = (ji * p.nunits) / p.ks, u1 = ((ji + 1) * p.nunits) / p.ks;
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
      for (int u = u0; u < u1; ++u) {
        mbar_wait_a(c.full0 + (uint32_t)c.slot * 8u, c.parity, c.limit);
        const unsigned char* wq = c.ring + (size_t)c.slot * ST_SLOT_BYTES + frag_off;
        const __nv_bfloat16* xp = xrow + u * ST_UNIT;
        const int npiece = min(8, (K - u * ST_UNIT) >> 5);  // K % 32 == 0: a 32-k piece is all or nothing
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q < npiece) {
            const uint4 wv = *reinterpret_cast<const uint4*>(wq + q * 64);
            const uint4 xv = *reinterpret_cast<const uint4*>(xp + q * 32);
            st_mma(acc, xv.x, xv.y, wv.x, wv.y);
            st_mma(acc2, xv.z, xv.w, wv.z, wv.w);
          }
        }
        __syncwarp();  // every lane is done reading the slot
        if (c.lane == 0) mbar_arrive_a(c.empty0 + (uint32_t)c.slot * 8u);
        if (++c.slot == c.D) { c.slot = 0; c.parity ^= 1u; }
      }
      if (g < R) {  // park the task's partial sums
        const int row_local = gi * ST_RG + 2 * t;
        if (p.n0 + row_local < p.n1) part[((size_t)row_local * p.ks + ji) * R + g] = acc[0] + acc2[0];
        if (p.n0 + row_local + 1 < p.n1) part[((size_t)(row_local + 1) * p.ks + ji) * R + g] = acc[1] + acc2[1];
      }
    }
  }
  ST_TRACE(3);
  cons_sync(c);
  ST_TRACE(4);

  // ---- epilogue: fixed-order sum over the K slices
  if (epi == GEMV_OUT_SWIGLU_BF16) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
    for (int i = tid; i < (nrows / 2) * R; i += nthreads) {
      const int pr = i / R, r = i - pr * R;
      float gsum = 0.f, usum = 0.f;
      for (int s = 0; s < p.ks; ++s) {
        gsum += part[((size_t)(2 * pr) * p.ks + s) * R + r];
        usum += part[((size_t)(2 * pr + 1) * p.ks + s) * R + r];
      }
      const float h = gsum / (1.0f + expf(-gsum)) * usum;
      out[(size_t)r * a.ld_out + (p.n0 >> 1) + pr] = __float2bfloat16_rn(h);
    }
  } else {
    const TpComm& tpc = sa.tpc;
    for (int i = tid; i < nrows * R; i += nthreads) {
      const int row = i / R, r = i - row * R;
      float s = 0.f;
      for (int q = 0; q < p.ks; ++q) s += part[((size_t)row * p.ks + q) * R + r];
      if (epi == GEMV_OUT_TP_PARTIAL) {
        const size_t idx = ((size_t)(a.tp_slot * tpc.tp + tpc.rank) * 8 + r) * tpc.H + p.n0 + row;
        for (int k = 0; k < tpc.tp; ++k) reinterpret_cast<float*>(tpc.peers[k] + tpc.off_inbox)[idx] = s;
      } else if (epi == GEMV_OUT_TP_LOGITS) {
        const size_t idx = (size_t)r * tpc.V + (size_t)tpc.rank * tpc.Vl + p.n0 + row;
        for (int k = 0; k < tpc.tp; ++k) reinterpret_cast<float*>(tpc.peers[k] + tpc.off_logits)[idx] = s;
      } else {
        float* o = a.out_f32 + (size_t)r * a.ld_out + p.n0 + row;
        if (epi == GEMV_OUT_ADD_RESID) *o = (i < nthreads ? resid_old : __ldcg(o)) + s;
        else *o = s;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ consumer: attention phase
// (row, head, 64-token split) items over groups of four consumer warps; arithmetic of decode_attn_kernel.
template <int R>
__device__ __forceinline__ void attn_phase(const StepArgs& sa, const StepPhase& P, const ConsumerCtx& c, unsigned char* scratch) {
  constexpr int D = 128, TPG = 8, CHUNK = 64;
  const int ngroups_cta = c.nw >> 2;             // groups of 4 warps
  const int grp_id = c.warp >> 2;
  if (grp_id >= ngroups_cta) return;             // warps beyond the last full group idle through the phase
  unsigned char* sc = scratch + (size_t)grp_id * ST_ATTN_SCRATCH;
  float* q_s = reinterpret_cast<float*>(sc);
  __nv_bfloat16* knew_s = reinterpret_cast<__nv_bfloat16*>(sc + 512);
  __nv_bfloat16* vnew_s = reinterpret_cast<__nv_bfloat16*>(sc + 768);
  float (*o_s)[D] = reinterpret_cast<float (*)[D]>(sc + 1024);
  float (*ml_s)[2] = reinterpret_cast<float (*)[2]>(sc + 1024 + 8 * D * 4);
  int* is_last = reinterpret_cast<int*>(sc + 1024 + 8 * D * 4 + 64);
  const StepAttn& A = sa.attn;
  const int tid = c.tid & (ST_ATTN_GROUP - 1);
  const int grp = tid >> 4, gl = tid & 15;
  const int Hl = A.heads * D;
  const int bar_id = 2 + grp_id;
  const int n_items = R * A.heads * A.max_splits;
  const int gstride = gridDim.x * ngroups_cta;
  for (int item = blockIdx.x * ngroups_cta + grp_id; item < n_items; item += gstride) {
    // split-major order: the live splits of all (row, head) pairs come first and spread evenly over the groups
    const int split = item / (R * A.heads);
    const int rh = item - split * (R * A.heads);
    const int r = rh / A.heads, h = rh - r * A.heads;
    const int pos = A.pos[r];
    const int ctx = pos + 1;
    const int nsplit = (ctx + CHUNK - 1) / CHUNK;
    if (split >= nsplit) continue;  // uniform over the group
    const int cap = A.kv_cap[r];
    __nv_bfloat16* Kc = reinterpret_cast<__nv_bfloat16*>(P.kcache) + A.kv_off[r] + (size_t)h * cap * D;
    __nv_bfloat16* Vc = reinterpret_cast<__nv_bfloat16*>(P.vcache) + A.kv_off[r] + (size_t)h * cap * D;
    const int t0 = split * CHUNK;
    uint4 kr[TPG], vr[TPG];
#pragma unroll
    for (int j = 0; j < TPG; ++j) {
      const int tk = t0 + grp + 8 * j;
      kr[j] = make_uint4(0u, 0u, 0u, 0u);
      vr[j] = kr[j];
      if (tk < pos) {
        kr[j] = __ldcg(reinterpret_cast<const uint4*>(Kc + (size_t)tk * D + gl * 8));
        vr[j] = __ldcg(reinterpret_cast<const uint4*>(Vc + (size_t)tk * D + gl * 8));
      }
    }
    {  // RoPE (rotate-half) of q; the split holding the new token also ropes k and appends k, v to the cache
      const float* qkv = A.qkv + (size_t)r * 3 * Hl;
      const int i = tid & 63;
      const float cs = A.rope_cos[(size_t)pos * (D / 2) + i], sn = A.rope_sin[(size_t)pos * (D / 2) + i];
      const float* qh = qkv + h * D;
      const float x1 = __ldcg(qh + i), x2 = __ldcg(qh + i + 64);
      q_s[tid] = tid < 64 ? (x1 * cs - x2 * sn) : (x2 * cs + x1 * sn);
      if (split == nsplit - 1) {
        const float* kh = qkv + Hl + h * D;
        const float k1 = __ldcg(kh + i), k2 = __ldcg(kh + i + 64);
        const __nv_bfloat16 kb = __float2bfloat16_rn(tid < 64 ? (k1 * cs - k2 * sn) : (k2 * cs + k1 * sn));
        const __nv_bfloat16 vb = __float2bfloat16_rn(__ldcg(qkv + 2 * Hl + h * D + tid));
        Kc[(size_t)pos * D + tid] = kb;
        Vc[(size_t)pos * D + tid] = vb;
        knew_s[tid] = kb;
        vnew_s[tid] = vb;
      }
    }
    named_sync(bar_id, ST_ATTN_GROUP);
    float qr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qr[i] = q_s[gl * 8 + i];
    float scv[TPG];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < TPG; ++j) {
      const int tk = t0 + grp + 8 * j;
      if (tk == pos) {
        kr[j] = *reinterpret_cast<const uint4*>(knew_s + gl * 8);
        vr[j] = *reinterpret_cast<const uint4*>(vnew_s + gl * 8);
      }
      float s = bf16lo(kr[j].x) * qr[0] + bf16hi(kr[j].x) * qr[1] + bf16lo(kr[j].y) * qr[2] + bf16hi(kr[j].y) * qr[3] +
                bf16lo(kr[j].z) * qr[4] + bf16hi(kr[j].z) * qr[5] + bf16lo(kr[j].w) * qr[6] + bf16hi(kr[j].w) * qr[7];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      scv[j] = tk <= pos ? s * A.scale : -INFINITY;
      m = fmaxf(m, scv[j]);
    }
    float l = 0.f, o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TPG; ++j) {
      const float pj = m == -INFINITY ? 0.f : expf(scv[j] - m);
      l += pj;
      o[0] = fmaf(pj, bf16lo(vr[j].x), o[0]); o[1] = fmaf(pj, bf16hi(vr[j].x), o[1]);
      o[2] = fmaf(pj, bf16lo(vr[j].y), o[2]); o[3] = fmaf(pj, bf16hi(vr[j].y), o[3]);
      o[4] = fmaf(pj, bf16lo(vr[j].z), o[4]); o[5] = fmaf(pj, bf16hi(vr[j].z), o[5]);
      o[6] = fmaf(pj, bf16lo(vr[j].w), o[6]); o[7] = fmaf(pj, bf16hi(vr[j].w), o[7]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) o_s[grp][gl * 8 + i] = o[i];
    if (gl == 0) { ml_s[grp][0] = m; ml_s[grp][1] = l; }
    named_sync(bar_id, ST_ATTN_GROUP);
    float M = -INFINITY;
#pragma unroll
    for (int gg = 0; gg < 8; ++gg) M = fmaxf(M, ml_s[gg][0]);
    float L = 0.f, od = 0.f;
#pragma unroll
    for (int gg = 0; gg < 8; ++gg) {
      const float w = ml_s[gg][0] == -INFINITY ? 0.f : expf(ml_s[gg][0] - M);
      L += ml_s[gg][1] * w;
      od += o_s[gg][tid] * w;
    }
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(A.out_bf16) + (size_t)r * Hl + h * D;
    if (nsplit == 1) {
      out[tid] = __float2bfloat16_rn(od / L);
    } else {
      // park the partial; the last split to arrive merges all of them in split order (deterministic)
      const size_t slot = (size_t)rh * A.max_splits + split;
      A.part_o[slot * D + tid] = od;
      if (tid == 0) { A.part_ml[slot * 2] = M; A.part_ml[slot * 2 + 1] = L; }
      __threadfence();
      named_sync(bar_id, ST_ATTN_GROUP);
      if (tid == 0) {
        const int prev = atomicAdd(&A.counters[rh], 1);
        *is_last = prev == nsplit - 1;
        if (prev == nsplit - 1) A.counters[rh] = 0;  // ready for the next layer
      }
      named_sync(bar_id, ST_ATTN_GROUP);
      if (*is_last) {
        __threadfence();
        float MM = -INFINITY;
        for (int s = 0; s < nsplit; ++s) MM = fmaxf(MM, __ldcg(&A.part_ml[((size_t)rh * A.max_splits + s) * 2]));
        float LL = 0.f, acc = 0.f;
        for (int s = 0; s < nsplit; ++s) {
          const size_t sl = (size_t)rh * A.max_splits + s;
          const float w = expf(__ldcg(&A.part_ml[sl * 2]) - MM);
          LL += __ldcg(&A.part_ml[sl * 2 + 1]) * w;
          acc += __ldcg(&A.part_o[sl * D + tid]) * w;
        }
        out[tid] = __float2bfloat16_rn(acc / LL);
      }
    }
    named_sync(bar_id, ST_ATTN_GROUP);  // scratch is reused by the group's next item
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int R>
__global__ void __launch_bounds__((ST_MAX_NW + 1) * 32, 1) decode_step_kernel(const StepArgs sa) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ StepShared sh;
  const int nw = (blockDim.x >> 5) - 1;
  const int D = sa.ring_depth;
  const int warp = threadIdx.x >> 5;
  // dynamic shared memory: [rings: nw * D slots][x / attention scratch][partials]
  const uint32_t ring_base = smem_u32(smem);
  unsigned char* smem_x = smem + (size_t)nw * D * ST_SLOT_BYTES;
  float* part = reinterpret_cast<float*>(smem_x + sa.x_bytes);

  if (threadIdx.x == 0) {
    for (int w = 0; w < nw; ++w)
      for (int s = 0; s < D; ++s) {
        mbar_init_a(smem_u32(&sh.full[w][s]), 1);
        mbar_init_a(smem_u32(&sh.empty[w][s]), 1);
      }
    mbar_init_a(smem_u32(&sh.xbar), 1);
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == nw) {
    // weights never depend on an earlier kernel: the stream starts before the programmatic-dependency wait of the consumers
    producer_loop(sa, &sh, ring_base, nw, D);
    return;
  }

  ConsumerCtx c;
  c.nw = nw;
  c.nthreads = nw * 32;
  c.tid = threadIdx.x;
  c.warp = warp;
  c.lane = threadIdx.x & 31;
  c.g = c.lane >> 2;
  c.t = c.lane & 3;
  c.D = D;
  c.ring = smem + (size_t)(warp * D) * ST_SLOT_BYTES;
  c.full0 = smem_u32(&sh.full[warp][0]);
  c.empty0 = smem_u32(&sh.empty[warp][0]);
  c.slot = 0;
  c.parity = 0;
  c.xparity = 0;
  c.limit = sa.tpc.tp > 1 ? ST_PEER_LIMIT_NS : ST_LOCAL_LIMIT_NS;
  unsigned long long* trace = sa.trace;

  pdl_wait();               // activations (and positions) written by earlier kernels
  pdl_launch_dependents();
  unsigned nbar = 0;        // grid barriers passed so far (one monotonic counter, zeroed before the launch)
#pragma unroll 1
  for (int ph = 0; ph < sa.nph; ++ph) {
    const StepPhase& P = sa.phases[ph];
    ST_TRACE(1);
    if (P.kind == STEP_ATTN) {
      if (!sa.skip_attn) attn_phase<R>(sa, P, c, smem_x);
      ++nbar;
      grid_barrier(c, sa.barrier, nbar * gridDim.x);
      continue;
    }
    gemv_phase<R>(sa, P, c, &sh, smem_x, part, trace, ph);
    ST_TRACE(5);
    const int epi = P.g.epilogue;
    const bool tp_epi = epi == GEMV_OUT_TP_PARTIAL || epi == GEMV_OUT_TP_LOGITS;
    const unsigned tp_target = tp_epi ? (unsigned)(*sa.tpc.epoch + 1) * (unsigned)sa.tpc.tp * gridDim.x : 0u;
    if (tp_epi) tp_arrive_all(c, sa.tpc, P.g.tp_counter);
    if (ph + 1 < sa.nph) {
      if (epi == GEMV_OUT_TP_PARTIAL) {
        // all-reduce inside the launch: every rank's partials are in every inbox once the cross-GPU counter is full; each CTA
        // sums ITS band over the source ranks in rank order into the replicated residual; a grid barrier publishes it
        const StepPlan p = step_plan(P.g.N, P.g.K, blockIdx.x, gridDim.x);
        tp_wait_all(c, sa.tpc, P.g.tp_counter, tp_target);
        const float* inbox = reinterpret_cast<const float*>(sa.tpc.peers[sa.tpc.rank] + sa.tpc.off_inbox);
        for (int i = c.tid; i < (p.n1 - p.n0) * R; i += c.nthreads) {
          const int row = i / R, r = i - row * R;
          float* o = P.g.out_f32 + (size_t)r * P.g.ld_out + p.n0 + row;
          float v = __ldcg(o);
          for (int k = 0; k < sa.tpc.tp; ++k)
            v += __ldcg(inbox + ((size_t)(P.g.tp_slot * sa.tpc.tp + k) * 8 + r) * sa.tpc.H + p.n0 + row);
          *o = v;
        }
      }
      ++nbar;
      grid_barrier(c, sa.barrier, nbar * gridDim.x);
    } else if (tp_epi) {
      tp_wait_all(c, sa.tpc, P.g.tp_counter, tp_target);  // the launch ends only when every rank's logits have landed here
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static size_t step_part_bytes(const GemvArgs& a, int grid) {
  size_t entries = 0;
  for (int cta = 0; cta < grid; ++cta) {
    const StepPlan p = step_plan(a.N, a.K, cta, grid);
    const size_t e = (size_t)p.ngroups * ST_RG * p.ks;
    entries = e > entries ? e : entries;
  }
  return entries * a.R * 4;
}

bool step_kernel_supports(const StepPhase* host_phases, int nph, int R) {
  if (R < 1 || R > 4 || nph < 1) return false;
  for (int i = 0; i < nph; ++i) {
    if (host_phases[i].kind != STEP_GEMV) continue;
    const GemvArgs& a = host_phases[i].g;
    if (a.R != R || (a.K % 32) != 0 || (a.N % 2) != 0) return false;
    if (a.x_resid && a.epilogue == GEMV_OUT_ADD_RESID) return false;
  }
  return true;
}

cudaError_t plan_decode_step(const StepPhase* host_phases, int nph, int R, int num_sms, StepLaunch* out) {
  if (!step_kernel_supports(host_phases, nph, R)) return cudaErrorInvalidValue;
  int grid = num_sms;
  int Kmax = 0;
  for (int i = 0; i < nph; ++i)
    if (host_phases[i].kind == STEP_GEMV) {
      if (host_phases[i].g.N / 2 < grid) grid = host_phases[i].g.N / 2;
      Kmax = host_phases[i].g.K > Kmax ? host_phases[i].g.K : Kmax;
    }
  size_t part = 0;
  for (int i = 0; i < nph; ++i)
    if (host_phases[i].kind == STEP_GEMV) {
      const size_t pb = step_part_bytes(host_phases[i].g, grid);
      part = pb > part ? pb : part;
    }
  size_t xb = (size_t)R * (Kmax + ST_XPAD) * 2;
  if (xb < 2 * (size_t)ST_ATTN_SCRATCH) xb = 2 * ST_ATTN_SCRATCH;
  if (part < (size_t)ST_VL * 4 * 4) part = (size_t)ST_VL * 4 * 4;  // the RMSNorm partials borrow this area during the prologue
  xb = (xb + 127) / 128 * 128;
  part = (part + 127) / 128 * 128;
  const long long avail = (long long)ST_SMEM_BUDGET - (long long)xb - (long long)part - 256;
  // ring geometry: as many 4.5 KB slots as fit; prefer more consumer warps at equal capacity (at least one attention group)
  int best_nw = 0, best_d = 0;
  int force_nw = 0, force_d = 0;
  if (const char* ev = getenv("VDD_STEP_WARPS")) force_nw = atoi(ev);
  if (const char* ev = getenv("VDD_STEP_DEPTH")) force_d = atoi(ev);
  // eight consumer warps (two attention groups, one RMSNorm lane per thread) unless that leaves fewer than three slots per warp
  for (int nw : {8, 4})
    for (int d = ST_MAX_D; d >= 2; --d) {
      if ((long long)nw * d * ST_SLOT_BYTES > avail) continue;
      if (force_nw && nw != force_nw) continue;
      if (force_d && d != force_d) continue;
      const bool better = !best_nw || (best_nw == 8 && best_d < 3 && nw * d > best_nw * best_d);
      if (better) { best_nw = nw; best_d = d; }
    }
  if (!best_nw) return cudaErrorInvalidValue;
  out->grid = grid;
  out->nw = best_nw;
  out->depth = best_d;
  out->x_bytes = (int)xb;
  out->smem = (size_t)best_nw * best_d * ST_SLOT_BYTES + xb + part + 128;
  return cudaSuccess;
}

cudaError_t launch_decode_step(const StepArgs& sa_in, const StepLaunch& L, int R, bool pdl, cudaStream_t stream) {
  void (*kern)(const StepArgs) = nullptr;
  switch (R) {
    case 1: kern = decode_step_kernel<1>; break;
    case 2: kern = decode_step_kernel<2>; break;
    case 3: kern = decode_step_kernel<3>; break;
    case 4: kern = decode_step_kernel<4>; break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM_BUDGET);
  if (e != cudaSuccess) return e;
  // the grid barriers need every CTA resident at once: one CTA per SM, grid <= #SMs, and the kernel must fit
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (L.nw + 1) * 32, L.smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  StepArgs sa = sa_in;
  sa.ring_depth = L.depth;
  sa.x_bytes = L.x_bytes;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(L.grid);
  cfg.blockDim = dim3((L.nw + 1) * 32);
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, sa);
}

}  // namespace vdd
