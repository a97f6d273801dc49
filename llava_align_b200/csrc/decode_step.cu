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
//   * the weights of ALL phases flow through per-warp rings of shared memory filled by TMA (cp.async.bulk.tensor: one 3-D tile
//     of 8 weight rows x 256 k = 4 KB per instruction, 128-byte swizzled, completion on an mbarrier).  Every warp feeds its own
//     ring: right after it has consumed a slot it requests the step that will occupy the slot next, D steps ahead in ITS OWN
//     (phase, task, unit) sequence -- a sequence that runs on across phase boundaries, attention phases and layers, because
//     weights never depend on activations.  So while the warps reduce, meet the other CTAs at the grid barrier and normalise the
//     next input, the rings (up to ~130 KB per SM) already hold or await the next projection's first steps.  TMA requests also
//     keep the LSU / L1 request queue free: the activation loads of a prologue no longer queue behind the weight requests.
//     (Measured dead ends: 512-byte 1-D bulk copies, 8 per step: 1.2 TB/s; one dedicated producer warp issuing every ring's
//     4 KB tiles in order: 2.1 TB/s -- a single thread's wait / expect_tx / issue chain is ~500 clocks per step, three times
//     the 175 clocks per step an SM needs.)
//   * the warps multiply with mma.sync m16n8k16 with swapped roles exactly as in decode_gemv.cu: the R activation rows are the M
//     side, 8 weight rows the N side, a lane's 16-byte piece of the ring IS its B fragment.  Task -> warp assignment is static
//     round-robin; the K split of a projection depends only on (N, K, grid), and the RMSNorm statistics are reduced over a
//     fixed set of 256 virtual lanes, so a row's bits do not depend on the ring geometry or on which rows are fused beside it.
//   * ATTENTION is a phase of the same launch: one WARP per (row, head, 32-token split) item, all of a split's K/V requested at
//     once (16 x 16 B per lane), RoPE + KV append in registers, no shared memory; the last split of a (row, head) to arrive
//     merges the partials warp-parallel.  The rings keep filling with o_proj meanwhile.
//   * Tensor parallel: the row-parallel epilogues push their partial band to every rank's inbox over NVLink and the next
//     phase reduces it after a cross-GPU arrival counter (same protocol as decode_gemv.cu), all inside this launch.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace vdd {

constexpr int ST_RG = 8;                       // weight rows per step (the N side of m16n8k16)
constexpr int ST_TILE = 256;                   // k elements per TMA tile (8 rows x 256 k = 4 KB)
constexpr int ST_SUB = 2;                      // tiles per step: a step is what a warp waits for, multiplies and re-requests at once
constexpr int ST_UNIT = ST_TILE * ST_SUB;      // k elements per step (8 KB of weights).  Per-step bookkeeping (barrier wait, expect_tx,
                                               // TMA issue through the uniform datapath, the two sequence iterators) is ~200 of the
                                               // ~350 instructions a warp spent on a 4 KB step (ncu: profiles/r2_ncu_step_4k.json, issue-
                                               // bound at 31 % with 8 warps of 255 registers); 8 KB steps halve it per byte
// A slot holds one TMA tile: 8 weight rows x 256 k, stored as 32 lines of 128 B (line = row * 4 + k / 64) with the hardware's
// 128-byte swizzle (16-byte chunk c of line L sits at chunk c ^ (L & 7)); slots are 1024-byte aligned so L & 7 is the line index
// inside the slot.  The two rows a quarter-warp reads (g = 2j, 2j + 1) land in different halves of the 128 bytes => the 16-byte
// fragment reads are bank-conflict free.
constexpr int ST_TILE_BYTES = 4096;
constexpr int ST_SLOT_BYTES = ST_SUB * ST_TILE_BYTES;
constexpr int ST_MAX_NW = 8;                   // warps per CTA
constexpr int ST_MAX_D = 8;                    // ring slots per warp
constexpr int ST_XPAD = 32;                    // bf16 elements of padding per x row when K % 64 == 0 (rows land 64 B apart modulo 128)
constexpr int ST_VL = 256;                     // virtual lanes of the RMSNorm reduction (independent of the thread count)
constexpr int ST_SMEM_BUDGET = 222 * 1024;     // 227 KB per CTA minus the kernel's static shared memory

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
__device__ __forceinline__ void tma_tile_3d(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_tile_2d(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// HBM -> L2 only (no shared memory, no completion): the same tile, requested well ahead of the load that will consume it
__device__ __forceinline__ void tma_prefetch_3d(const void* tmap, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
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
  uint64_t full[ST_MAX_NW][ST_MAX_D];   // per (warp, slot): the TMA tile of the slot has landed
  uint64_t xbar;                        // completion of the bf16 activation bulk copies of a prologue
  float rstd[4];
};

// Position of a warp in ITS OWN sequence of weight steps: (projection phase, task, 256-k unit).  Task i of a projection belongs
// to warp i % nw.  The same walk is done twice per warp: `pi` requests tiles (D steps ahead), `ci` consumes them.
struct WIter {
  int ph;                       // index of a STEP_GEMV phase, or nph when the step's weights are exhausted
  int n0, n1, ks, nunits, ntasks, plain;
  int task, gi, ji, u, u1;
};
__device__ __forceinline__ void it_set_task(WIter& it) {
  it.gi = it.task / it.ks;
  it.ji = it.task - it.gi * it.ks;
  it.u = (it.ji * it.nunits) / it.ks;
  it.u1 = ((it.ji + 1) * it.nunits) / it.ks;
}
// first projection phase >= ph_from in which this warp owns a task
__device__ __forceinline__ void it_enter(const StepArgs& sa, WIter& it, int ph_from, int warp) {
  for (int ph = ph_from; ph < sa.nph; ++ph) {
    const StepPhase& P = sa.phases[ph];
    if (P.kind != STEP_GEMV) continue;
    const StepPlan p = step_plan(P.g.N, P.g.K, blockIdx.x, gridDim.x);
    if (warp >= p.ntasks) continue;
    it.ph = ph; it.n0 = p.n0; it.n1 = p.n1; it.ks = p.ks; it.nunits = p.nunits; it.ntasks = p.ntasks; it.plain = P.tmap_plain;
    it.task = warp;
    it_set_task(it);
    return;
  }
  it.ph = sa.nph;
}
__device__ __forceinline__ void it_next(const StepArgs& sa, WIter& it, int warp, int nw) {
  if (++it.u < it.u1) return;
  it.task += nw;
  if (it.task < it.ntasks) { it_set_task(it); return; }
  it_enter(sa, it, it.ph + 1, warp);
}

struct ConsumerCtx {
  int nw, nthreads, tid, warp, lane, g, t, D;
  const unsigned char* ring;  // this warp's slot 0
  uint32_t ring_s;            // ... and its shared-space address
  uint32_t full0;             // shared address of this warp's full barriers
  int slot;                   // next slot to consume (= the slot the next request refills)
  uint32_t parity;            // its phase parity
  uint32_t xparity;           // phase parity of xbar
  unsigned long long limit;   // spin bound of the local waits
  uint64_t policy;            // L2 evict-first for the weight stream
  WIter pi, ci, qi;           // request (D steps ahead of ci), consume, L2 prefetch (sa.l2_ahead steps ahead of pi)
};

#define ST_TRACE(ev)                                                                                         \
  do {                                                                                                       \
    if (trace && c.tid == 0 && ph < 64) trace[((size_t)blockIdx.x * 64 + ph) * 8 + (ev)] = st_gtime();    \
  } while (0)

// pull the tile of step `it` from HBM into L2 (one lane).  The shared-memory rings hold ~130 KB per SM, three microseconds of
// the HBM stream -- less than a phase boundary (epilogue + grid barrier + prologue) lasts, and barely HBM's latency x bandwidth
// product under load.  L2 (126 MB) is the deeper buffer: every warp keeps `l2_ahead` further steps of its own sequence in flight
// towards L2, so the ring loads hit L2 (~1 us) and HBM keeps streaming through the boundaries.
__device__ __forceinline__ void l2_request(const StepArgs& sa, const ConsumerCtx& c, const WIter& it) {
  if (c.lane == 0) {
    const int row0 = it.n0 + it.gi * ST_RG;
    const int nvalid = min(ST_RG, it.n1 - row0);
    const void* tm = &sa.phases[it.ph].tmap[(nvalid >> 1) - 1][0];
#pragma unroll
    for (int sub = 0; sub < ST_SUB; ++sub) {
      if (it.plain) tma_prefetch_2d(tm, it.u * ST_UNIT + sub * ST_TILE, row0);
      else tma_prefetch_3d(tm, 0, (it.u * ST_SUB + sub) * 4, row0);
    }
  }
}
// request the tile of step `it` into slot `slot` of this warp's ring (one lane; the others go on)
__device__ __forceinline__ void ring_request(const StepArgs& sa, const ConsumerCtx& c, const WIter& it, int slot) {
  if (c.lane == 0) {
    const int row0 = it.n0 + it.gi * ST_RG;
    const int nvalid = min(ST_RG, it.n1 - row0);  // even: bands are pairs of rows
    const uint32_t fb = c.full0 + (uint32_t)slot * 8u;
    const uint32_t dst = c.ring_s + (uint32_t)slot * ST_SLOT_BYTES;
    const void* tm = &sa.phases[it.ph].tmap[(nvalid >> 1) - 1][0];
    // box = 64 k x 4 lines x nvalid rows (or 256 k x nvalid rows); parts beyond the matrix are zero-filled and counted
    mbar_expect_a(fb, (uint32_t)(nvalid * ST_UNIT * 2));
#pragma unroll
    for (int sub = 0; sub < ST_SUB; ++sub) {  // tiles wholly beyond K arrive as zeros (and are skipped by the consumer)
      if (it.plain) tma_tile_2d(dst + sub * ST_TILE_BYTES, tm, fb, it.u * ST_UNIT + sub * ST_TILE, row0, c.policy);
      else tma_tile_3d(dst + sub * ST_TILE_BYTES, tm, fb, 0, (it.u * ST_SUB + sub) * 4, row0, c.policy);
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
      uint2 nwv[PASS];  // the norm weights are static: requested together with the rows, not after the reduction
#pragma unroll
      for (int q = 0; q < PASS; ++q) {
        const int ch = tid + q * ST_VL;
#pragma unroll
        for (int r = 0; r < R; ++r)
          xv[r][q] = ch < nchunk ? __ldcg(reinterpret_cast<const float4*>(a.x_resid + (size_t)r * K) + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
        nwv[q] = ch < nchunk ? __ldg(reinterpret_cast<const uint2*>(nw) + ch) : make_uint2(0u, 0u);
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
          const uint2 wv = nwv[q];
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

  // ---- stream loop: this warp's steps of this phase; lane (g, t) accumulates y[r = g][features 2t, 2t+1 of the group]
  {
    const __nv_bfloat16* xrow = xs + (size_t)(g < R ? g : 0) * XS + t * 8;
    // lane (g, t), piece q: logical 16-byte chunk q * 4 + t of weight row g = line g * 4 + (q >> 1), chunk (q & 1) * 4 + t, swizzled
    uint32_t frag_off[8];
    const int plain = P.tmap_plain;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int line = g * 4 + (q >> 1);
      frag_off[q] = plain ? (uint32_t)(g * 512 + q * 64 + t * 16) : (uint32_t)(line * 128 + (((((q & 1) << 2) | t) ^ (line & 7)) << 4));
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f}, acc3[4] = {0.f, 0.f, 0.f, 0.f}, acc4[4] = {0.f, 0.f, 0.f, 0.f};
    while (c.ci.ph == ph) {
      mbar_wait_a(c.full0 + (uint32_t)c.slot * 8u, c.parity, c.limit);
      const unsigned char* wq = c.ring + (size_t)c.slot * ST_SLOT_BYTES;
      const int u = c.ci.u;
      const __nv_bfloat16* xp = xrow + u * ST_UNIT;
      const int npiece = min(8 * ST_SUB, (K - u * ST_UNIT) >> 5);  // K % 32 == 0: a 32-k piece is all or nothing
      uint4 wv[8 * ST_SUB];
#pragma unroll
      for (int sub = 0; sub < ST_SUB; ++sub)
#pragma unroll
        for (int q = 0; q < 8; ++q)  // stale / zero bytes beyond npiece are never used
          wv[sub * 8 + q] = *reinterpret_cast<const uint4*>(wq + sub * ST_TILE_BYTES + frag_off[q]);
      __syncwarp();  // every lane holds its fragments: the slot may be refilled
      if (c.pi.ph < sa.nph) {
        ring_request(sa, c, c.pi, c.slot);  // the step D ahead of this one (possibly of a later phase / layer) takes the slot over
        it_next(sa, c.pi, c.warp, c.nw);
      }
      if (c.qi.ph < sa.nph) {
        l2_request(sa, c, c.qi);
        it_next(sa, c.qi, c.warp, c.nw);
      }
#pragma unroll
      for (int q = 0; q < 8 * ST_SUB; q += 2) {  // four independent accumulation chains
        if (q < npiece) {
          const uint4 xv = *reinterpret_cast<const uint4*>(xp + q * 32);
          st_mma(acc, xv.x, xv.y, wv[q].x, wv[q].y);
          st_mma(acc2, xv.z, xv.w, wv[q].z, wv[q].w);
        }
        if (q + 1 < npiece) {
          const uint4 xv = *reinterpret_cast<const uint4*>(xp + (q + 1) * 32);
          st_mma(acc3, xv.x, xv.y, wv[q + 1].x, wv[q + 1].y);
          st_mma(acc4, xv.z, xv.w, wv[q + 1].z, wv[q + 1].w);
        }
      }
      if (++c.slot == c.D) { c.slot = 0; c.parity ^= 1u; }
      if (u + 1 >= c.ci.u1) {  // task complete: park its partial sums
        if (g < R) {
          const int row_local = c.ci.gi * ST_RG + 2 * t;
          if (p.n0 + row_local < p.n1) part[((size_t)row_local * p.ks + c.ci.ji) * R + g] = (acc[0] + acc2[0]) + (acc3[0] + acc4[0]);
          if (p.n0 + row_local + 1 < p.n1)
            part[((size_t)(row_local + 1) * p.ks + c.ci.ji) * R + g] = (acc[1] + acc2[1]) + (acc3[1] + acc4[1]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = acc2[e] = acc3[e] = acc4[e] = 0.f;
      }
      it_next(sa, c.ci, c.warp, c.nw);
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

// ------------------------------------------------------------------------------------------------ attention phase
// One WARP per (row, head, 32-token split).  Half-warp hw takes tokens t0 + hw + 2j (j = 0..15), lane gl of a half 8 of the 128
// dims; all 32 K/V loads of a lane are issued before anything is consumed; RoPE of q (and of the new k) in registers; the two
// halves merge by shuffles; the last split of a (row, head) to arrive merges the partials warp-parallel, in split order.
constexpr int AT_CHUNK = 32, AT_PASS = 16;
__device__ __forceinline__ void rope8(const float* base, const float* cs, const float* sn, int gl, float (&out)[8]) {
  // dims d = gl*8 .. +7 of a 128-dim head, rotate-half: d < 64: x[d] c[d] - x[d+64] s[d];  d >= 64: x[d] c[d-64] + x[d-64] s[d-64]
  const int d0 = gl * 8;
  const int i0 = d0 & 63;
  const float4 a0 = __ldcg(reinterpret_cast<const float4*>(base + d0)), a1 = __ldcg(reinterpret_cast<const float4*>(base + d0 + 4));
  const int o0 = d0 < 64 ? d0 + 64 : d0 - 64;
  const float4 b0 = __ldcg(reinterpret_cast<const float4*>(base + o0)), b1 = __ldcg(reinterpret_cast<const float4*>(base + o0 + 4));
  const float4 c0 = __ldg(reinterpret_cast<const float4*>(cs + i0)), c1 = __ldg(reinterpret_cast<const float4*>(cs + i0 + 4));
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(sn + i0)), s1 = __ldg(reinterpret_cast<const float4*>(sn + i0 + 4));
  const float x[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, y[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w}, ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) out[e] = d0 < 64 ? (x[e] * cc[e] - y[e] * ss[e]) : (x[e] * cc[e] + y[e] * ss[e]);
}
__device__ __forceinline__ float dot8(const uint4& k, const float (&q)[8]) {
  return bf16lo(k.x) * q[0] + bf16hi(k.x) * q[1] + bf16lo(k.y) * q[2] + bf16hi(k.y) * q[3] + bf16lo(k.z) * q[4] + bf16hi(k.z) * q[5] +
         bf16lo(k.w) * q[6] + bf16hi(k.w) * q[7];
}
__device__ __forceinline__ void axpy8(float p, const uint4& v, float (&o)[8]) {
  o[0] = fmaf(p, bf16lo(v.x), o[0]); o[1] = fmaf(p, bf16hi(v.x), o[1]);
  o[2] = fmaf(p, bf16lo(v.y), o[2]); o[3] = fmaf(p, bf16hi(v.y), o[3]);
  o[4] = fmaf(p, bf16lo(v.z), o[4]); o[5] = fmaf(p, bf16hi(v.z), o[5]);
  o[6] = fmaf(p, bf16lo(v.w), o[6]); o[7] = fmaf(p, bf16hi(v.w), o[7]);
}
template <int R>
__device__ __forceinline__ void attn_phase(const StepArgs& sa, const StepPhase& P, const ConsumerCtx& c) {
  constexpr int D = 128;
  const StepAttn& A = sa.attn;
  const int lane = c.lane, hw = lane >> 4, gl = lane & 15;
  const int Hl = A.heads * D;
  // only LIVE items are enumerated (rows differ sevenfold in length): item -> (row, split, head), head fastest
  int rpos[R], rns[R], n_items = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    rpos[r] = A.pos[r];
    rns[r] = (rpos[r] + AT_CHUNK) / AT_CHUNK;  // ceil((pos + 1) / 32)
    n_items += rns[r] * A.heads;
  }
  const int wstride = gridDim.x * c.nw;
  for (int item = blockIdx.x * c.nw + c.warp; item < n_items; item += wstride) {
    int r = 0, rem = item, pos = rpos[0], nsplit = rns[0];
#pragma unroll
    for (int rr = 1; rr < R; ++rr) {
      if (r == rr - 1 && rem >= rns[rr - 1] * A.heads) { rem -= rns[rr - 1] * A.heads; r = rr; pos = rpos[rr]; nsplit = rns[rr]; }
    }
    const int split = rem / A.heads, h = rem - split * A.heads;
    const int rh = r * A.heads + h;
    const int cap = A.kv_cap[r];
    __nv_bfloat16* Kc = reinterpret_cast<__nv_bfloat16*>(P.kcache) + A.kv_off[r] + (size_t)h * cap * D;
    __nv_bfloat16* Vc = reinterpret_cast<__nv_bfloat16*>(P.vcache) + A.kv_off[r] + (size_t)h * cap * D;
    const int t0 = split * AT_CHUNK;
    uint4 kr[AT_PASS], vr[AT_PASS];
#pragma unroll
    for (int j = 0; j < AT_PASS; ++j) {
      const int tk = t0 + hw + 2 * j;
      kr[j] = make_uint4(0u, 0u, 0u, 0u);
      vr[j] = kr[j];
      if (tk < pos) {
        kr[j] = __ldcg(reinterpret_cast<const uint4*>(Kc + (size_t)tk * D + gl * 8));
        vr[j] = __ldcg(reinterpret_cast<const uint4*>(Vc + (size_t)tk * D + gl * 8));
      }
    }
    const float* qkv = A.qkv + (size_t)r * 3 * Hl;
    const float* cs = A.rope_cos + (size_t)pos * (D / 2);
    const float* sn = A.rope_sin + (size_t)pos * (D / 2);
    float q[8];
    rope8(qkv + h * D, cs, sn, gl, q);
    const bool owns_new = split == nsplit - 1;
    uint4 knew = make_uint4(0u, 0u, 0u, 0u), vnew = knew;
    if (owns_new) {  // rope the new k, round k / v to bf16, append to the cache (half-warp 0 writes), keep them for the fold below
      float kf[8];
      rope8(qkv + Hl + h * D, cs, sn, gl, kf);
      const float4 v0 = __ldcg(reinterpret_cast<const float4*>(qkv + 2 * Hl + h * D + gl * 8));
      const float4 v1 = __ldcg(reinterpret_cast<const float4*>(qkv + 2 * Hl + h * D + gl * 8 + 4));
      knew = make_uint4(pack_bf16(kf[0], kf[1]), pack_bf16(kf[2], kf[3]), pack_bf16(kf[4], kf[5]), pack_bf16(kf[6], kf[7]));
      vnew = make_uint4(pack_bf16(v0.x, v0.y), pack_bf16(v0.z, v0.w), pack_bf16(v1.x, v1.y), pack_bf16(v1.z, v1.w));
      if (hw == 0) {
        *reinterpret_cast<uint4*>(Kc + (size_t)pos * D + gl * 8) = knew;
        *reinterpret_cast<uint4*>(Vc + (size_t)pos * D + gl * 8) = vnew;
      }
    }
    // scores of this half-warp's tokens (16-lane reductions), then one softmax pass
    float sc[AT_PASS];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < AT_PASS; ++j) {
      const int tk = t0 + hw + 2 * j;
      if (tk == pos) { kr[j] = knew; vr[j] = vnew; }  // only in the owning split
      float sdot = dot8(kr[j], q);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
      sc[j] = tk <= pos ? sdot * A.scale : -INFINITY;
      m = fmaxf(m, sc[j]);
    }
    float l = 0.f, o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
    for (int j = 0; j < AT_PASS; ++j) {
      const float pj = m == -INFINITY ? 0.f : expf(sc[j] - m);
      l += pj;
      axpy8(pj, vr[j], o);
    }
    // merge the two half-warps (half 0 first: fixed order)
    {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, 16), l2 = __shfl_xor_sync(0xffffffffu, l, 16);
      const float M = fmaxf(m, m2);
      const float wa = m == -INFINITY ? 0.f : expf(m - M), wb = m2 == -INFINITY ? 0.f : expf(m2 - M);
      float oo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) oo[e] = __shfl_xor_sync(0xffffffffu, o[e], 16);
      // lanes of half 0 hold (own = half 0, other = half 1); lanes of half 1 compute the mirrored sum and are ignored
      l = l * wa + l2 * wb;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = o[e] * wa + oo[e] * wb;
      m = M;
    }
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(A.out_bf16) + (size_t)r * Hl + h * D;
    if (nsplit == 1) {
      if (hw == 0) {
        const float inv = 1.0f / l;
        *reinterpret_cast<uint4*>(out + gl * 8) = make_uint4(pack_bf16(o[0] * inv, o[1] * inv), pack_bf16(o[2] * inv, o[3] * inv),
                                                             pack_bf16(o[4] * inv, o[5] * inv), pack_bf16(o[6] * inv, o[7] * inv));
      }
      continue;
    }
    // park the partial; the last split to arrive merges all of them in split order (deterministic)
    const size_t slot0 = (size_t)rh * A.max_splits;
    if (hw == 0) {
      float* po = A.part_o + (slot0 + split) * D + gl * 8;
      *reinterpret_cast<float4*>(po) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(po + 4) = make_float4(o[4], o[5], o[6], o[7]);
      if (gl == 0) { A.part_ml[(slot0 + split) * 2] = m; A.part_ml[(slot0 + split) * 2 + 1] = l; }
    }
    __threadfence();
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      const int prev = atomicAdd(&A.counters[rh], 1);
      last = prev == nsplit - 1;
      if (last) A.counters[rh] = 0;  // ready for the next layer
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) continue;
    __threadfence();
    float MM = -INFINITY;
    for (int s0 = 0; s0 < nsplit; s0 += 32) {
      const int sidx = s0 + lane;
      MM = fmaxf(MM, sidx < nsplit ? __ldcg(&A.part_ml[(slot0 + sidx) * 2]) : -INFINITY);
    }
    MM = warp_max(MM);
    float LL = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);  // lane owns dims lane*4 .. +3
    for (int s0 = 0; s0 < nsplit; s0 += 32) {
      const int sidx = s0 + lane;
      const float ms = sidx < nsplit ? __ldcg(&A.part_ml[(slot0 + sidx) * 2]) : -INFINITY;
      const float ls = sidx < nsplit ? __ldcg(&A.part_ml[(slot0 + sidx) * 2 + 1]) : 0.f;
      const float ws = ms == -INFINITY ? 0.f : expf(ms - MM);
      const int nb = min(32, nsplit - s0);
      float lsum = ls * ws;  // fixed order: lane 0 .. 31 through the shuffle tree, blocks in order
      lsum = warp_sum(lsum);
      LL += lsum;
#pragma unroll 8
      for (int k = 0; k < nb; ++k) {
        const float wk = __shfl_sync(0xffffffffu, ws, k);
        const float4 pv = __ldcg(reinterpret_cast<const float4*>(A.part_o + (slot0 + s0 + k) * D) + lane);
        acc.x = fmaf(wk, pv.x, acc.x); acc.y = fmaf(wk, pv.y, acc.y); acc.z = fmaf(wk, pv.z, acc.z); acc.w = fmaf(wk, pv.w, acc.w);
      }
    }
    const float inv = 1.0f / LL;
    *reinterpret_cast<uint2*>(out + lane * 4) = make_uint2(pack_bf16(acc.x * inv, acc.y * inv), pack_bf16(acc.z * inv, acc.w * inv));
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int R>
__global__ void __launch_bounds__(ST_MAX_NW * 32, 1) decode_step_kernel(const StepArgs sa) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ StepShared sh;
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzled TMA tiles want 1024-byte alignment
  const int nw = blockDim.x >> 5;
  const int D = sa.ring_depth;
  const int warp = threadIdx.x >> 5;
  // dynamic shared memory: [rings: nw * D slots][x][partials]
  unsigned char* smem_x = smem + (size_t)nw * D * ST_SLOT_BYTES;
  float* part = reinterpret_cast<float*>(smem_x + sa.x_bytes);

  if (threadIdx.x == 0) {
    for (int w = 0; w < nw; ++w)
      for (int s = 0; s < D; ++s) mbar_init_a(smem_u32(&sh.full[w][s]), 1);
    mbar_init_a(smem_u32(&sh.xbar), 1);
    fence_barrier_init();
  }
  __syncthreads();

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
  c.ring_s = smem_u32(c.ring);
  c.full0 = smem_u32(&sh.full[warp][0]);
  c.slot = 0;
  c.parity = 0;
  c.xparity = 0;
  c.limit = sa.tpc.tp > 1 ? ST_PEER_LIMIT_NS : ST_LOCAL_LIMIT_NS;
  c.policy = sa.l2_hint ? policy_evict_first() : policy_evict_normal();
  unsigned long long* trace = sa.trace;

  // Fill this warp's ring before anything else: weights never depend on an earlier kernel, so the stream starts ahead of the
  // programmatic-dependency wait.
  it_enter(sa, c.pi, 0, warp);
  c.ci = c.pi;
  for (int s = 0; s < D; ++s) {
    if (c.pi.ph >= sa.nph) break;
    ring_request(sa, c, c.pi, s);
    it_next(sa, c.pi, warp, nw);
  }
  c.qi = c.pi;
  for (int s = 0; s < sa.l2_ahead; ++s) {
    if (c.qi.ph >= sa.nph) break;
    l2_request(sa, c, c.qi);
    it_next(sa, c.qi, warp, nw);
  }
  if (sa.l2_ahead <= 0) c.qi.ph = sa.nph;  // prefetching off
  pdl_wait();               // activations (and positions) written by earlier kernels
  pdl_launch_dependents();
  unsigned nbar = 0;        // grid barriers passed so far (one monotonic counter, zeroed before the launch)
#pragma unroll 1
  for (int ph = 0; ph < sa.nph; ++ph) {
    const StepPhase& P = sa.phases[ph];
    ST_TRACE(1);
    if (P.kind == STEP_ATTN) {
      if (!sa.skip_attn) attn_phase<R>(sa, P, c);
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

typedef CUresult (*StEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
cudaError_t step_build_tmaps(StepPhase* ph, int nph) {
  static StEncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
    enc = reinterpret_cast<StEncodeTiledFn>(fn);
  }
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  for (int i = 0; i < nph; ++i) {
    if (ph[i].kind != STEP_GEMV) continue;
    const GemvArgs& a = ph[i].g;
    // W [N, K] row-major seen as [N][ceil(K/64)][64]: the middle dimension walks a row in 128-byte lines
    cuuint64_t dims[3] = {64, (cuuint64_t)(a.K / 64), (cuuint64_t)a.N};
    cuuint64_t strides[2] = {128, (cuuint64_t)a.K * 2};
    cuuint32_t estr[3] = {1, 1, 1};
    ph[i].tmap_plain = (a.K % 64) != 0 || getenv("VDD_STEP_PLAIN_TMA") != nullptr;
    for (int b = 0; b < 4; ++b) {
      CUtensorMap m;
      CUresult r;
      if (!ph[i].tmap_plain) {
        cuuint32_t box[3] = {64, 4, (cuuint32_t)(2 * (b + 1))};
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else {  // [N][K], box 256 k x rows, rows 512 B apart in shared memory, columns beyond K zero-filled
        cuuint64_t d2[2] = {(cuuint64_t)a.K, (cuuint64_t)a.N};
        cuuint64_t s2[1] = {(cuuint64_t)a.K * 2};
        cuuint32_t box[2] = {256, (cuuint32_t)(2 * (b + 1))};
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a.w), d2, s2, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
      memcpy(ph[i].tmap[b], &m, 128);
    }
  }
  return cudaSuccess;
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
  if (part < (size_t)ST_VL * 4 * 4) part = (size_t)ST_VL * 4 * 4;  // the RMSNorm partials borrow this area during the prologue
  xb = (xb + 127) / 128 * 128;
  part = (part + 127) / 128 * 128;
  const long long avail = (long long)ST_SMEM_BUDGET - (long long)xb - (long long)part - 1024 - 256;
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
  out->smem = (size_t)best_nw * best_d * ST_SLOT_BYTES + xb + part + 1024 + 128;
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
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, L.nw * 32, L.smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  StepArgs sa = sa_in;
  sa.ring_depth = L.depth;
  sa.x_bytes = L.x_bytes;
  // measured: 8 / 16 / 32 steps ahead slow the step down by 20-30 % (947 -> 1141 / 1156 / 1260 ms per 256-token generate): off
  static const int l2_ahead = getenv("VDD_STEP_L2_AHEAD") ? atoi(getenv("VDD_STEP_L2_AHEAD")) : 0;
  sa.l2_ahead = l2_ahead < 0 ? 0 : (l2_ahead > 64 ? 64 : l2_ahead);
  static const int l2_hint = getenv("VDD_STEP_NOHINT") ? 0 : 1;
  sa.l2_hint = l2_hint;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(L.grid);
  cfg.blockDim = dim3(L.nw * 32);
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
