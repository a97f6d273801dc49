// Decode-step projection for R <= 8 fused rows (all debias streams of all sequences): y[r, n] = sum_k x[r,k] W[n,k].
//
// This is the HBM-bound heart of the decode step (reference: one cuBLAS GEMV per stream per projection inside HF
// LlamaModel, entered from experiments/llava/model/language_model/llava_llama.py:91-103 once PER STREAM per token;
// here the weights are streamed ONCE for all R rows).  Algorithmic bytes = N*K*2.
//
// Layout: W bf16 [N, K] row-major.  Each CTA owns a contiguous band of output features; inside the CTA a warp
// task is (8 consecutive features) x (a K-slice).  One warp step covers 8 rows x 256 k; each of its 8 cp.async
// instructions (16 B per lane, L1 bypass) moves 512 CONTIGUOUS bytes of one weight row into a PRIVATE shared-memory
// ring of 3 steps per warp (two steps = 8 KB per warp in flight while the third is consumed).  The ring is warp
// private, so cp.async.wait_group + __syncwarp is the only synchronisation in the stream loop; a 16-byte-chunk XOR
// swizzle makes the fragment reads (lane (g, t) = (lane/4, lane%4) reads row g) bank-conflict free.
// The arithmetic runs on the tensor cores as mma.sync m16n8k16 (bf16 x bf16 -> fp32) with the ROLES SWAPPED: the R<=8
// activation rows are the M side (rows R..15 are zero registers) and the 8 weight rows are the N side, so a lane's
// 16-byte weight piece IS its B fragment for two MMAs and its 16-byte piece of x (bf16 in shared memory) IS the A
// fragment -- the k index is permuted identically on both sides (slot pairs {2t,2t+1 | 2t+8,2t+9} <- elements
// t*8+{0,1 | 2,3} then {4,5 | 6,7}), which a dot product does not notice.  No unpacking, no shuffles: ~6 instructions
// per 512 weight bytes instead of ~300 on the CUDA cores, which is what moved this kernel from issue-bound to
// HBM-bound.  (tcgen05 needs M>=64 weight rows per instruction and TMEM round trips; at 3 rows of x the legacy MMA
// shape wastes nothing that matters -- the tensor pipe is <2% busy -- and keeps the weight stream in registers.)
// K-slice partial sums are parked in shared memory and summed in a fixed order => deterministic.  Prologues: plain
// bf16 x, or RMSNorm(residual) computed on the fly; epilogues: fp32 store, residual +=, SwiGLU over interleaved
// gate/up rows.  With programmatic dependent launch the first TWO steps of every warp (128 KB per SM) are requested
// BEFORE griddepcontrol.wait: weights do not depend on the previous kernel, so HBM stays busy while the previous
// kernel drains and while this one normalises / stages x.
//
// FUSED CHAIN: one launch can run up to four projections back to back (o -> gate_up -> down -> next layer's qkv), with a
// grid-wide barrier (one atomic counter per boundary, all 148 CTAs co-resident at one CTA per SM) where a kernel
// boundary used to be.  The next projection's first ring steps are requested BEFORE the barrier -- weights never
// depend on activations -- so the weight stream does not stop while CTAs wait for each other; that removes two thirds
// of the launch boundaries of a decode step (the probe in tools/stream_probe.cu shows a chain of 129 separate
// read-only kernels cannot exceed ~0.87 of the copy bandwidth on this part, whatever the kernel does).
// Activations produced by other CTAs are read with ld.global.cg (L2) because L1 is not coherent across a grid barrier.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vdd {

constexpr int GV_RG = 8;      // weight rows per warp task (the N side of m16n8k16)
constexpr int GV_UNIT = 256;  // k elements per warp step (8 pieces of 32 k; lane t takes 8 of each 32)
constexpr int GV_D = 3;       // ring depth in steps per warp (GV_D - 1 steps in flight)
constexpr int GV_STEP_BYTES = GV_RG * GV_UNIT * 2;  // 4 KB
constexpr int GV_MAX_WARPS = 16;
constexpr int GV_MAX_TASKS = 256;
constexpr int GV_XPAD = 32;   // bf16 elements of padding per x row: rows land 64 B apart modulo 128 (no LDS conflicts)
constexpr int GV_SMEM_BUDGET = 222 * 1024;  // 227 KB per CTA minus the kernel's static shared memory

struct GemvPlan {
  int n0, n1;  // CTA's feature band
  int ngroups, ks, ntasks, nunits;
};
__host__ __device__ __forceinline__ GemvPlan make_plan(int N, int K, int cta, int nctas, int ksmul) {
  GemvPlan p;
  const int pairs = N >> 1;
  p.n0 = 2 * (int)(((long long)cta * pairs) / nctas);
  p.n1 = 2 * (int)(((long long)(cta + 1) * pairs) / nctas);
  const int nrows = p.n1 - p.n0;
  p.ngroups = (nrows + GV_RG - 1) / GV_RG;
  p.nunits = (K + GV_UNIT - 1) / GV_UNIT;
  // the K-split must not depend on the warp count (which follows R through the shared-memory budget): a row's
  // result has to be bit-identical whatever other rows are fused beside it
  int ks = p.ngroups > 0 ? (ksmul * GV_MAX_WARPS + p.ngroups - 1) / p.ngroups : 1;
  const int ksmax = p.nunits / 2 > 1 ? p.nunits / 2 : 1;
  ks = ks < 1 ? 1 : (ks > ksmax ? ksmax : ks);
  p.ks = ks;
  p.ntasks = p.ngroups * ks;
  return p;
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  // rows 8..15 of A (registers a1, a3) are zero: only R <= 8 activation rows exist
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%5}, {%7,%8}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(0u), "r"(a2), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Grid barrier, split in two so that the next projection's weight requests are issued between arrive and wait.
// All CTAs of the grid are co-resident (grid <= #SMs, one CTA per SM); bounded spin => a bug traps instead of hanging.
__device__ __forceinline__ void grid_arrive(unsigned* counter) {
  __syncthreads();  // every thread's global stores of this phase are issued
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
  }
}
__device__ __forceinline__ void grid_wait(const unsigned* counter, unsigned target) {
  if (threadIdx.x == 0) {
    unsigned spins = 0;
    while (ld_acquire_u32(counter) < target) {
      if (++spins > (1u << 22)) __trap();
    }
  }
  __syncthreads();
}

// ---- cross-GPU barrier over peer memory (tensor parallel): every CTA of every rank bumps the counter on EVERY rank
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void tp_arrive(const TpComm& t, int cidx) {
  __syncthreads();  // every thread's peer stores of this phase are issued
  if (threadIdx.x == 0) {
    __threadfence_system();
    for (int k = 0; k < t.tp; ++k) {
      unsigned* c = reinterpret_cast<unsigned*>(t.peers[k]) + cidx;
      asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
    }
  }
}
__device__ __forceinline__ void tp_wait(const TpComm& t, int cidx, unsigned target) {
  if (threadIdx.x == 0) {
    const unsigned* c = reinterpret_cast<const unsigned*>(t.peers[t.rank]) + cidx;
    unsigned spins = 0;
    while (ld_acquire_sys_u32(c) < target) {
      if (++spins > (1u << 23)) __trap();
    }
    __threadfence_system();
  }
  __syncthreads();
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define GV_TRACE(ev)                                                                                  \
  do {                                                                                                \
    if (trace && threadIdx.x == 0) trace[((size_t)blockIdx.x * 4 + ph) * 8 + (ev)] = gtime();         \
  } while (0)

// position of a warp inside its (task, step) sequence
struct GemvIter {
  int task, tg, u, u_end, j, seq;
};
struct GemvCtx {
  GemvPlan p;
  GemvIter pi, ci;
  int4* task_tab;
  int* task_next;   // shared counter: tasks are handed to warps on demand (balance does not depend on ntasks % nwarps)
  int* task_fifo;   // [nwarps][8]: ids grabbed by a warp's request iterator, replayed by its consume iterator
  uint32_t ring_warp;  // shared-space address of slot 0 of this warp's ring
  int nwarps, g, t, ksmul;
  uint2 nwv[4];  // RMSNorm weights of this thread's columns, fetched before the grid barrier (they are static)
};
__device__ __forceinline__ void it_load(const GemvCtx& c, GemvIter& it, int task) {
  it.task = task;
  if (task < c.p.ntasks) {
    const int4 tk = c.task_tab[task];
    it.tg = tk.x; it.u = tk.y; it.u_end = tk.z; it.j = tk.w;
  }
}
// the request iterator grabs the next task of the CTA and remembers it for the consume iterator of the same warp
__device__ __forceinline__ void it_grab(const GemvCtx& c, GemvIter& it) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int task = 0;
  if (lane == 0) {
    task = atomicAdd(c.task_next, 1);
    c.task_fifo[warp * 8 + (it.seq & 7)] = task;
  }
  task = __shfl_sync(0xffffffffu, task, 0);
  it.seq++;
  it_load(c, it, task);
}
__device__ __forceinline__ void it_replay(const GemvCtx& c, GemvIter& it) {
  __syncwarp();
  const int task = c.task_fifo[(threadIdx.x >> 5) * 8 + (it.seq & 7)];
  it.seq++;
  it_load(c, it, task);
}
template <bool PRODUCER>
__device__ __forceinline__ void it_advance(const GemvCtx& c, GemvIter& it) {
  if (++it.u >= it.u_end) {
    if (PRODUCER) it_grab(c, it); else it_replay(c, it);
  }
}
// request one step into ring slot `slot`: instruction q copies 512 CONTIGUOUS bytes of weight row q of the group (lane l
// takes bytes [16 l, 16 l + 16)) -- the best shape HBM can be asked for.  The destination is swizzled (16-byte chunk
// index ^ 4 on odd rows) so that the fragment reads below are bank-conflict free.  Always commits a group so the
// counting stays uniform.
__device__ __forceinline__ void ring_request(const GemvCtx& c, const __nv_bfloat16* W, int K, const GemvIter& it, int slot) {
  if (it.task < c.p.ntasks) {
    const int lane = threadIdx.x & 31;
    const int k0 = it.u * GV_UNIT + lane * 8;
    const bool ok = k0 < K;  // ragged K (tensor-parallel slices): zero-fill
    const int row0 = c.p.n0 + it.tg * GV_RG;
    const uint32_t dst = c.ring_warp + (uint32_t)slot * GV_STEP_BYTES;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int row = row0 + q;
      row = row < c.p.n1 ? row : c.p.n1 - 1;
      const __nv_bfloat16* src = W + (size_t)row * K + k0;
      cp_async16(dst + q * 512 + ((lane ^ ((q & 1) << 2)) << 4), ok ? (const void*)src : (const void*)W, ok ? 16u : 0u);
    }
  }
  cp_async_commit();
}

// Part 1 of a projection: plan, task table, first GV_D-1 ring steps.  Touches only weights (never activations).
__device__ __forceinline__ void phase_begin(const GemvArgs& a, GemvCtx& c) {
  const int tid = threadIdx.x, nthreads = blockDim.x, warp = tid >> 5;
  c.p = make_plan(a.N, a.K, blockIdx.x, gridDim.x, c.ksmul);
  for (int i = tid; i < c.p.ntasks; i += nthreads) {
    const int gi = i / c.p.ks, ji = i - gi * c.p.ks;
    c.task_tab[i] = make_int4(gi, (ji * c.p.nunits) / c.p.ks, ((ji + 1) * c.p.nunits) / c.p.ks, ji);
  }
  if (tid == 0) *c.task_next = 0;
  if (a.x_resid) {
    const __nv_bfloat16* nw = reinterpret_cast<const __nv_bfloat16*>(a.norm_w);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = (q * nthreads + tid) * 4;
      c.nwv[q] = k < a.K ? __ldg(reinterpret_cast<const uint2*>(nw + k)) : make_uint2(0u, 0u);
    }
  }
  __syncthreads();
  (void)warp;
  c.pi.seq = 0;
  c.ci.seq = 0;
  it_grab(c, c.pi);
  it_replay(c, c.ci);
  const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(a.w);
#pragma unroll
  for (int s = 0; s < GV_D - 1; ++s) {
    ring_request(c, W, a.K, c.pi, s);
    if (c.pi.task < c.p.ntasks) it_advance<true>(c, c.pi);
  }
}

// Part 2: stage x, stream the rest of the weights, reduce, epilogue.  Prologue / epilogue kinds are runtime switches
// around ONE copy of the stream loop (a fused chain runs four projections through the same instructions).
template <int R>
__device__ __forceinline__ void phase_rest(const GemvArgs& a, GemvCtx& c, unsigned char* smem, float* red,
                                           unsigned long long* trace, int ph, const TpComm& tpc) {
  const int K = a.K;
  const int XS = K + GV_XPAD;  // x row stride (elements)
  const int nthreads = blockDim.x;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = c.g, t = c.t;
  const GemvPlan& p = c.p;
  const int epi = a.epilogue;
  unsigned char* ring = smem;                                                                              // [nwarps][GV_D][4 KB]
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem + (size_t)c.nwarps * GV_D * GV_STEP_BYTES);  // [R][XS]
  float* part = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(xs) + (size_t)R * XS * 2);       // [nrows][ks][R]
  const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(a.w);
  const int nrows = p.n1 - p.n0;

  // residual value this thread will add to (fetched now, needed after the stream loop)
  float resid_old = 0.f;
  if (epi == GEMV_OUT_ADD_RESID && tid < nrows * R) {
    const int row = tid / R, r = tid - row * R;
    resid_old = __ldcg(a.out_f32 + (size_t)r * a.ld_out + p.n0 + row);
  }
  // ---- prologue: stage x (bf16) in shared memory.  Loads go to L2 (.cg): the producer may be another CTA of this launch.
  if (a.x_resid) {
    // RMSNorm.  (The launcher picks the thread count from K and the row CLASS (<=4 / <=8 rows), never from R itself,
    // so the reduction order below -- and with it every output bit of a row -- is the same whatever is fused beside it.)
    const __nv_bfloat16* nw = reinterpret_cast<const __nv_bfloat16*>(a.norm_w);
    constexpr int PASS = R <= 3 ? 4 : (R <= 4 ? 3 : 1);  // float4 per thread per row held in registers
    if (K <= PASS * 4 * nthreads) {
      float4 xv[R][PASS];
      float ss[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        ss[r] = 0.f;
#pragma unroll
        for (int q = 0; q < PASS; ++q) {
          const int k = (q * nthreads + tid) * 4;
          xv[r][q] = k < K ? __ldcg(reinterpret_cast<const float4*>(a.x_resid + (size_t)r * K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int q = 0; q < PASS; ++q)
          ss[r] += xv[r][q].x * xv[r][q].x + xv[r][q].y * xv[r][q].y + xv[r][q].z * xv[r][q].z + xv[r][q].w * xv[r][q].w;
      float rstd[R];
#pragma unroll
      for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(block_reduce<false>(ss[r], red) / (float)K + a.eps);
#pragma unroll
      for (int q = 0; q < PASS; ++q) {
        const int k = (q * nthreads + tid) * 4;
        if (k < K) {
          const uint2 wv = c.nwv[q];
          const float w0 = bf16lo(wv.x), w1 = bf16hi(wv.x), w2 = bf16lo(wv.y), w3 = bf16hi(wv.y);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            uint2 o;
            o.x = pack_bf16(xv[r][q].x * rstd[r] * w0, xv[r][q].y * rstd[r] * w1);
            o.y = pack_bf16(xv[r][q].z * rstd[r] * w2, xv[r][q].w * rstd[r] * w3);
            *reinterpret_cast<uint2*>(xs + (size_t)r * XS + k) = o;
          }
        }
      }
    } else {  // wide rows: two passes over L2
      float ss[R];
#pragma unroll
      for (int r = 0; r < R; ++r) ss[r] = 0.f;
      for (int k = tid * 4; k < K; k += nthreads * 4) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 xv = __ldcg(reinterpret_cast<const float4*>(a.x_resid + (size_t)r * K + k));
          ss[r] += xv.x * xv.x + xv.y * xv.y + xv.z * xv.z + xv.w * xv.w;
        }
      }
      float rstd[R];
#pragma unroll
      for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(block_reduce<false>(ss[r], red) / (float)K + a.eps);
      for (int k = tid * 4; k < K; k += nthreads * 4) {
        const uint2 wv = *reinterpret_cast<const uint2*>(nw + k);
        const float w0 = bf16lo(wv.x), w1 = bf16hi(wv.x), w2 = bf16lo(wv.y), w3 = bf16hi(wv.y);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 xv = __ldcg(reinterpret_cast<const float4*>(a.x_resid + (size_t)r * K + k));
          uint2 o;
          o.x = pack_bf16(xv.x * rstd[r] * w0, xv.y * rstd[r] * w1);
          o.y = pack_bf16(xv.z * rstd[r] * w2, xv.w * rstd[r] * w3);
          *reinterpret_cast<uint2*>(xs + (size_t)r * XS + k) = o;
        }
      }
    }
  } else {
    // bf16 x: asynchronous 16-byte copies straight into shared memory, all in flight at once
    const int cpr = K / 8;  // 16-byte chunks per row
    const __nv_bfloat16* xg = reinterpret_cast<const __nv_bfloat16*>(a.x_bf16);
    for (int i = tid; i < R * cpr; i += nthreads) {
      const int r = i / cpr, q = i - r * cpr;
      cp_async16(smem_u32(xs + (size_t)r * XS + q * 8), xg + (size_t)r * K + q * 8, 16u);
    }
    cp_async_commit();
    cp_async_wait<0>();  // (also drains the ring steps requested in phase_begin, which are needed next anyway)
  }
  __syncthreads();
  GV_TRACE(2);

  // ---- stream loop; lane (g, t) accumulates y[r = g][features 2t, 2t+1 of the group]
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};  // two independent MMA chains
  // lanes with g >= R read x row 0: their products land in accumulator rows nobody stores
  const __nv_bfloat16* xrow = xs + (size_t)(g < R ? g : 0) * XS + t * 8;
  int cslot = 0, pslot = GV_D - 1;
  // lane (g, t) reads weight row g: logical 16-byte chunk qq*4+t of the row's 512 bytes, at its swizzled position
  // (row parity flips bit 2 of the chunk index = the low bit of the 64-byte piece index q)
  const uint32_t frag_off = (uint32_t)g * 512 + (uint32_t)t * 16;
  const int godd = g & 1;
  while (c.ci.task < p.ntasks) {
    __syncwarp();  // every lane is done reading the slot that is about to be refilled
    ring_request(c, W, K, c.pi, pslot);
    if (c.pi.task < p.ntasks) it_advance<true>(c, c.pi);
    cp_async_wait<GV_D - 1>();  // everything but the newest GV_D-1 groups has landed: the step to consume is there
    __syncwarp();               // ... for every lane of the warp (fragments are assembled from other lanes' copies)
    {
      const unsigned char* wq = ring + (size_t)(warp * GV_D + cslot) * GV_STEP_BYTES + frag_off;
      const __nv_bfloat16* xp = xrow + c.ci.u * GV_UNIT;
      const bool full = (c.ci.u + 1) * GV_UNIT <= K;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 wv = *reinterpret_cast<const uint4*>(wq + ((q ^ godd) << 6));
        uint4 xv = make_uint4(0u, 0u, 0u, 0u);
        if (full || c.ci.u * GV_UNIT + q * 32 + t * 8 < K) xv = *reinterpret_cast<const uint4*>(xp + q * 32);
        mma_bf16_16816(acc, xv.x, xv.y, wv.x, wv.y);
        mma_bf16_16816(acc2, xv.z, xv.w, wv.z, wv.w);
      }
    }
    if (c.ci.u + 1 >= c.ci.u_end) {  // task complete: park its partial sums
      if (g < R) {
        const int row_local = c.ci.tg * GV_RG + 2 * t;
        if (p.n0 + row_local < p.n1) part[((size_t)row_local * p.ks + c.ci.j) * R + g] = acc[0] + acc2[0];
        if (p.n0 + row_local + 1 < p.n1) part[((size_t)(row_local + 1) * p.ks + c.ci.j) * R + g] = acc[1] + acc2[1];
      }
      acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
      acc2[0] = acc2[1] = acc2[2] = acc2[3] = 0.f;
    }
    it_advance<false>(c, c.ci);
    cslot = cslot + 1 == GV_D ? 0 : cslot + 1;
    pslot = pslot + 1 == GV_D ? 0 : pslot + 1;
  }
  cp_async_wait<0>();
  GV_TRACE(3);
  __syncthreads();
  GV_TRACE(4);

  // ---- epilogue: fixed-order sum over the K-slices
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
    for (int i = tid; i < nrows * R; i += nthreads) {
      const int row = i / R, r = i - row * R;
      float s = 0.f;
      for (int q = 0; q < p.ks; ++q) s += part[((size_t)row * p.ks + q) * R + r];
      GV_TRACE(7);
      if (epi == GEMV_OUT_TP_PARTIAL) {
        // this rank's K-slice of the projection: hand the partial to every rank (its own inbox plane included)
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

template <int R>
__global__ void __launch_bounds__(GV_MAX_WARPS * 32, 1) gemv_kernel(const GemvChain ch) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ float red[33];
  __shared__ int4 task_tab[GV_MAX_TASKS];
  __shared__ int task_next;
  __shared__ int task_fifo[GV_MAX_WARPS * 8];
  GemvCtx c;
  c.task_tab = task_tab;
  c.task_next = &task_next;
  c.task_fifo = task_fifo;
  c.nwarps = blockDim.x >> 5;
  c.ksmul = ch.ksmul;
  c.g = (threadIdx.x & 31) >> 2;
  c.t = threadIdx.x & 3;
  c.ring_warp = smem_u32(smem) + (uint32_t)(threadIdx.x >> 5) * GV_D * GV_STEP_BYTES;

  unsigned long long* trace = ch.trace;
  {
    const int ph = 0;
    GV_TRACE(0);
    if (trace && threadIdx.x == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      trace[((size_t)blockIdx.x * 4 + 0) * 8 + 6] = smid;
    }
  }
  phase_begin(ch.ph[0], c);  // weights do not depend on the previous kernel: request before the PDL wait
  pdl_wait();
  pdl_launch_dependents();
#pragma unroll 1
  for (int ph = 0; ph < ch.n; ++ph) {
    const GemvArgs& a = ch.ph[ph];
    GV_TRACE(1);
    phase_rest<R>(a, c, smem, red, trace, ph, ch.tpc);
    GV_TRACE(5);
    const bool tp_epi = a.epilogue == GEMV_OUT_TP_PARTIAL || a.epilogue == GEMV_OUT_TP_LOGITS;
    const unsigned tp_target = tp_epi ? (unsigned)(*ch.tpc.epoch + 1) * (unsigned)ch.tpc.tp * gridDim.x : 0u;
    if (tp_epi) tp_arrive(ch.tpc, a.tp_counter);
    if (ph + 1 < ch.n) {
      if (a.epilogue == GEMV_OUT_TP_PARTIAL) {
        // all-reduce fused into the chain: (1) every rank's partials are in every inbox once the cross-GPU counter is
        // full, (2) each CTA sums ITS band over the source ranks in rank order into the replicated residual, (3) a
        // local grid barrier publishes the residual to the next projection's RMSNorm.
        const int bn0 = c.p.n0, bn1 = c.p.n1;
        phase_begin(ch.ph[ph + 1], c);  // next weights are requested before any waiting
        tp_wait(ch.tpc, a.tp_counter, tp_target);
        const float* inbox = reinterpret_cast<const float*>(ch.tpc.peers[ch.tpc.rank] + ch.tpc.off_inbox);
        for (int i = threadIdx.x; i < (bn1 - bn0) * R; i += blockDim.x) {
          const int row = i / R, r = i - row * R;
          float* o = a.out_f32 + (size_t)r * a.ld_out + bn0 + row;
          float v = __ldcg(o);
          for (int k = 0; k < ch.tpc.tp; ++k)
            v += __ldcg(inbox + ((size_t)(a.tp_slot * ch.tpc.tp + k) * 8 + r) * ch.tpc.H + bn0 + row);
          *o = v;
        }
        grid_arrive(ch.barrier + ph);
        grid_wait(ch.barrier + ph, gridDim.x);
      } else {
        grid_arrive(ch.barrier + ph);
        phase_begin(ch.ph[ph + 1], c);  // the next projection's weights start flowing while we wait for the other CTAs
        grid_wait(ch.barrier + ph, gridDim.x);
      }
    } else if (tp_epi) {
      tp_wait(ch.tpc, a.tp_counter, tp_target);  // the kernel ends only when every rank's contribution has landed here
    }
  }
}

static size_t gemv_smem(const GemvArgs& a, int grid, int nwarps, int ksmul) {
  const int Rc = a.R <= 4 ? 4 : 8;  // row class: the warp count must not depend on R itself (see the RMSNorm prologue)
  size_t entries = 0;  // max over CTAs of (8-row groups) * 8 * ks  >=  rows * ks
  for (int cta = 0; cta < grid; ++cta) {
    const GemvPlan p = make_plan(a.N, a.K, cta, grid, ksmul);
    const size_t e = (size_t)p.ngroups * GV_RG * p.ks;
    entries = e > entries ? e : entries;
  }
  const size_t xbytes = (size_t)Rc * (a.K + GV_XPAD) * 2;
  const size_t part = entries * Rc * 4;
  return (size_t)nwarps * GV_D * GV_STEP_BYTES + xbytes + part + 64;
}

cudaError_t launch_gemv_chain(const GemvChain& ch, int num_sms, bool pdl, cudaStream_t stream) {
  if (ch.n < 1 || ch.n > 4) return cudaErrorInvalidValue;
  const int R = ch.ph[0].R;
  int grid = num_sms;
  for (int i = 0; i < ch.n; ++i) {
    const GemvArgs& a = ch.ph[i];
    if (a.R != R || a.R < 1 || a.R > 8 || (a.K % 8) != 0 || (a.N % 2) != 0) return cudaErrorInvalidValue;
    if (a.x_resid && a.epilogue == GEMV_OUT_ADD_RESID) return cudaErrorInvalidValue;
    if ((a.epilogue == GEMV_OUT_TP_PARTIAL || a.epilogue == GEMV_OUT_TP_LOGITS) && (ch.tpc.tp < 2 || !ch.tpc.epoch)) return cudaErrorInvalidValue;
    if (a.epilogue == GEMV_OUT_TP_PARTIAL && i + 1 == ch.n) return cudaErrorInvalidValue;  // the reduce lives in the next phase
    if (a.N / 2 < grid) grid = a.N / 2;
  }
  if (ch.n > 1 && (!ch.barrier || grid > num_sms)) return cudaErrorInvalidValue;
  int ksmul = 8;
  if (const char* ev = getenv("VDD_GV_KSMUL")) {  // tuning knob: ~ksmul*16 warp tasks per CTA
    const int v = atoi(ev);
    if (v >= 1 && v <= 8) ksmul = v;
  }
  // one CTA per SM; as many warps (each with its private 3-step ring) as shared memory allows next to x
  int nwarps = GV_MAX_WARPS;
  if (const char* ev = getenv("VDD_GV_WARPS")) {  // tuning knob
    const int w = atoi(ev);
    if (w >= 1 && w <= GV_MAX_WARPS) nwarps = w;
  }
  size_t smem = 0;
  for (; nwarps >= 1; nwarps -= (nwarps > 4 ? 2 : 1)) {
    smem = 0;
    for (int i = 0; i < ch.n; ++i) {
      const size_t s = gemv_smem(ch.ph[i], grid, nwarps, ksmul);
      smem = s > smem ? s : smem;
    }
    if (smem <= (size_t)GV_SMEM_BUDGET) break;
  }
  if (nwarps < 1) return cudaErrorInvalidValue;
  void (*kern)(const GemvChain) = nullptr;
  switch (R) {
    case 1: kern = gemv_kernel<1>; break;
    case 2: kern = gemv_kernel<2>; break;
    case 3: kern = gemv_kernel<3>; break;
    case 4: kern = gemv_kernel<4>; break;
    case 5: kern = gemv_kernel<5>; break;
    case 6: kern = gemv_kernel<6>; break;
    case 7: kern = gemv_kernel<7>; break;
    case 8: kern = gemv_kernel<8>; break;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GV_SMEM_BUDGET);
  if (e != cudaSuccess) return e;
  if (ch.n > 1) {
    // The split grid barrier between chained projections needs every CTA of the grid resident at once.  Checked against the
    // occupancy calculator for this kernel / block / shared-memory combination (once per combination); what it cannot see is another
    // stream's kernel holding SMs for the whole launch -- the barrier's bounded spin turns that into a trap, not a hang.
    static int ok_R = 0, ok_threads = 0, ok_grid = 0;
    static size_t ok_smem = 0;
    if (!(ok_R == R && ok_threads == nwarps * 32 && ok_smem == smem && ok_grid == grid)) {
      int per_sm = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nwarps * 32, smem);
      if (e != cudaSuccess) return e;
      if (per_sm < 1 || grid > per_sm * num_sms) return cudaErrorCooperativeLaunchTooLarge;
      ok_R = R; ok_threads = nwarps * 32; ok_smem = smem; ok_grid = grid;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(nwarps * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  GemvChain chl = ch;
  chl.ksmul = ksmul;
  return cudaLaunchKernelEx(&cfg, kern, chl);
}

cudaError_t launch_gemv(const GemvArgs& a, int num_sms, bool pdl, cudaStream_t stream) {
  GemvChain ch{};
  ch.ph[0] = a;
  ch.n = 1;
  return launch_gemv_chain(ch, num_sms, pdl, stream);
}

}  // namespace vdd
