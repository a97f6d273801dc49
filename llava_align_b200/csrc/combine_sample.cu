// K8: contrastive logit combine + HF warpers + softmax + multinomial + EOS bookkeeping in ONE launch.
//
// Replaces reference vcd_utils/vcd_sample.py:185 (mean of the two debias streams), :188-194 (plausibility
// cutoff on the raw logits and (1+a)*l - a*cd), :197-198 (Temperature -> TopK -> TopP warpers of
// transformers/generation/logits_process.py), :201-202 (softmax + torch.multinomial == argmax(p/q), q~Exp(1)),
// :257-260 (finished rows emit pad) and :285-288 (unfinished &= tok != eos).
//
// One CTA of 1024 threads per sequence; the V<=32768 scores live in registers (32 per thread), the survivors of
// the cutoff are compacted to shared memory and bitonic-sorted there only when top-k/top-p need an order.
// HBM-bound: algorithmic bytes per sequence = S*V*4 read (+V*4 q, +V*4 scores when requested).
// Every fp32 operation that decides set membership is written with explicit round-to-nearest intrinsics in
// the reference's operation order (no FMA contraction), so scores are bit-identical to torch fp32 on the same inputs.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vdd {

// ------------------------------------------------------------------------------------------------ torch's exponential_ stream
// Philox4x32-10 exactly as cuRAND runs it for ATen (curand_init(seed, subsequence = thread index, offset) + curand_uniform4), the
// element -> (thread, call, component) mapping of ATen's distribution_elementwise_grid_stride_kernel (block 256, unroll 4,
// grid = min(#SM * maxThreadsPerSM / 256, ceil(numel / 256))) and transformation::exponential (-log(u), log clamped at 1).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float torch_exponential_from_bits(unsigned x) {
  const float u = __fmaf_rn(__uint2float_rn(x), 2.3283064e-10f, 1.1641532e-10f);  // curand_uniform: (0, 1]; the product is exact
  // at::transformation::exponential, CUDA branch: log clamped near 1, and the FAST logarithm -- measured on the B200 box: __logf
  // reproduces torch.empty(n).exponential_(1) bit for bit, logf differs in 85 % of the elements (tools/philox_probe.py)
  const float lg = u >= 1.0f - 5.9604645e-08f ? -5.9604645e-08f : __logf(u);
  return -lg;
}
__device__ __forceinline__ float philox_q(const SampleArgs& a, int step, long long li) {
  const unsigned T = a.philox_threads;
  const unsigned idx = (unsigned)(li % T);
  const unsigned j = (unsigned)(li / T);
  const unsigned long long ctr = (a.philox_offset + (unsigned long long)step * a.philox_stride) / 4ull + (j >> 2);
  const uint4 r = philox4x32_10(make_uint4((unsigned)ctr, (unsigned)(ctr >> 32), idx, 0u),
                                make_uint2((unsigned)a.philox_seed, (unsigned)(a.philox_seed >> 32)));
  const unsigned comp = j & 3u;
  return torch_exponential_from_bits(comp == 0 ? r.x : (comp == 1 ? r.y : (comp == 2 ? r.z : r.w)));
}
void sample_philox_geometry(SampleArgs& a, long long numel, int num_sms, int max_threads_per_sm) {
  unsigned long long grid = (unsigned long long)((numel + 255) / 256);
  const unsigned long long cap = (unsigned long long)num_sms * (unsigned long long)(max_threads_per_sm / 256);
  if (grid > cap) grid = cap;
  a.philox_threads = (unsigned)(256ull * grid);
  a.philox_stride = ((unsigned long long)(numel - 1) / (256ull * grid * 4ull) + 1ull) * 4ull;
}
__global__ void philox_exponential_kernel(SampleArgs a, long long numel, float* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x)
    out[i] = philox_q(a, 0, i);
}
cudaError_t launch_philox_exponential(unsigned long long seed, unsigned long long offset, long long numel, int num_sms,
                                      int max_threads_per_sm, float* out, cudaStream_t stream) {
  SampleArgs a{};
  a.philox_seed = seed;
  a.philox_offset = offset;
  sample_philox_geometry(a, numel, num_sms, max_threads_per_sm);
  philox_exponential_kernel<<<296, 256, 0, stream>>>(a, numel, out);
  return cudaGetLastError();
}

constexpr int CS_THREADS = 1024;
constexpr int CS_ITEMS = 32;  // V <= 32768

__device__ __forceinline__ int block_sum_int(int v, int* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    int t = scratch[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
__device__ __forceinline__ int block_min_int(int v, int* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    int t = scratch[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = min(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
// exclusive prefix sum of one int per thread over the 1024-thread block; *total gets the block sum
__device__ __forceinline__ int block_exscan_int(int v, int* scratch, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) scratch[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = scratch[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    scratch[lane] = winc - w;
    if (lane == 31) scratch[32] = winc;
  }
  __syncthreads();
  *total = scratch[32];
  return scratch[wid] + inc - v;
}
// inclusive SUFFIX sum of one float per thread (thread 1023 first): returns sum over threads >= tid
__device__ __forceinline__ float block_suffix_float(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(0xffffffffu, inc, o);
    if (lane + o < 32) inc += t;
  }
  __syncthreads();
  if (lane == 0) scratch[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    float w = scratch[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_down_sync(0xffffffffu, winc, o);
      if (lane + o < 32) winc += t;
    }
    scratch[lane] = winc - w;  // sum over warps strictly after this one
  }
  __syncthreads();
  return scratch[wid] + inc;
}

__global__ void __launch_bounds__(CS_THREADS, 1) combine_sample_kernel(const SampleArgs byval, const SampleArgs* dptr) {
  extern __shared__ float sortbuf[];  // next_pow2(V) floats
  __shared__ float redf[33];
  __shared__ int redi[33];
  __shared__ unsigned long long best_s;

  const SampleArgs& a = dptr ? *dptr : byval;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int V = a.V;
  const int step = a.step_ptr ? *a.step_ptr : 0;
  const float* l = a.l + (int64_t)b * a.ld;
  const float* u = a.u ? a.u + (int64_t)b * a.ld : nullptr;
  const float* d = a.d ? a.d + (int64_t)b * a.ld : nullptr;
  const int64_t so = ((int64_t)step * a.B + b) * V;
  const float* q = a.q ? a.q + so : nullptr;
  float* scores_out = a.scores_out ? a.scores_out + so : nullptr;

  float v[CS_ITEMS];
  // ---- pass 1: max of the raw main-stream logits (vcd_sample.py:191)
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {
    const int e = j * CS_THREADS + tid;
    v[j] = e < V ? __ldg(l + e) : -INFINITY;
    mx = fmaxf(mx, v[j]);
  }
  mx = block_reduce<true>(mx, redf);

  // ---- pass 2: combine + plausibility mask (vcd_sample.py:185,193-194), then temperature
  if (a.use_contrast && (u || d)) {
    const float cutoff = __fadd_rn(a.log_beta, mx);
    const float a1 = a.one_plus_alpha, al = a.alpha;
#pragma unroll
    for (int j = 0; j < CS_ITEMS; ++j) {
      const int e = j * CS_THREADS + tid;
      if (e < V) {
        float cd;
        if (u && d) cd = __fmul_rn(__fadd_rn(__ldg(u + e), __ldg(d + e)), 0.5f);
        else cd = u ? __ldg(u + e) : __ldg(d + e);
        const float s = __fsub_rn(__fmul_rn(a1, v[j]), __fmul_rn(al, cd));
        v[j] = v[j] < cutoff ? -INFINITY : s;
      }
    }
  }
  if (a.temperature != 1.0f) {
#pragma unroll
    for (int j = 0; j < CS_ITEMS; ++j) v[j] = __fdiv_rn(v[j], a.temperature);
  }

  // ---- warpers that need an order: TopK then TopP over the surviving (finite) scores
  const bool want_k = a.top_k > 0 && a.top_k < V;
  const bool want_p = a.top_p < 1.0f;
  if (want_k || want_p) {
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < CS_ITEMS; ++j) cnt += (v[j] > -INFINITY) ? 1 : 0;
    int n_fin;
    int off = block_exscan_int(cnt, redi, &n_fin);
    int P = 1;
    while (P < n_fin) P <<= 1;
#pragma unroll
    for (int j = 0; j < CS_ITEMS; ++j)
      if (v[j] > -INFINITY) sortbuf[off++] = v[j];
    for (int i = n_fin + tid; i < P; i += CS_THREADS) sortbuf[i] = -INFINITY;
    __syncthreads();
    // bitonic sort, descending
    for (int k = 2; k <= P; k <<= 1) {
      for (int jj = k >> 1; jj > 0; jj >>= 1) {
        for (int i = tid; i < P; i += CS_THREADS) {
          const int x = i ^ jj;
          if (x > i) {
            const float A = sortbuf[i], Bv = sortbuf[x];
            const bool desc = (i & k) == 0;
            if ((A < Bv) == desc) { sortbuf[i] = Bv; sortbuf[x] = A; }
          }
        }
        __syncthreads();
      }
    }
    int n_keep = n_fin;
    float thr = -INFINITY;  // keep score >= thr
    if (want_k && a.top_k < n_fin) {
      // TopKLogitsWarper: remove scores < k-th largest (ties with the k-th stay)
      thr = sortbuf[a.top_k - 1];
      int c = 0;
      for (int i = tid; i < n_fin; i += CS_THREADS) c += sortbuf[i] >= thr ? 1 : 0;
      n_keep = block_sum_int(c, redi);
    }
    int kept_ties = -1;  // >=0: only that many of the scores == thr survive (those with the highest index)
    if (want_p && n_keep > 1) {
      // TopPLogitsWarper: ascending cumulative softmax mass <= 1 - top_p is removed, the largest always stays.
      const float m0 = sortbuf[0];
      const int per = (n_keep + CS_THREADS - 1) / CS_THREADS;
      const int i0 = min(tid * per, n_keep), i1 = min(i0 + per, n_keep);
      float loc = 0.f;
      for (int i = i1 - 1; i >= i0; --i) loc += expf(sortbuf[i] - m0);
      const float suf_incl = block_suffix_float(loc, redf);  // sum over my chunk and all later ones
      const float Z = block_reduce<false>(tid == 0 ? suf_incl : 0.f, redf);
      const float lim = 1.0f - a.top_p;
      float run = suf_incl - loc;  // mass strictly after my chunk
      int first_removed = n_keep;
      for (int i = i1 - 1; i >= i0; --i) {
        run += expf(sortbuf[i] - m0);
        if (i >= 1 && __fdiv_rn(run, Z) <= lim) first_removed = i;
      }
      const int n3 = block_min_int(first_removed, redi);
      if (n3 < n_keep) {
        thr = sortbuf[n3 - 1];
        if (sortbuf[n3] == thr) {  // the boundary splits a run of equal scores
          int c = 0;
          for (int i = tid; i < n3; i += CS_THREADS) c += sortbuf[i] > thr ? 1 : 0;
          kept_ties = n3 - block_sum_int(c, redi);
        }
        n_keep = n3;
      }
    }
    __syncthreads();
    if (kept_ties >= 0) {
      // stable ascending sort puts the lower index first among equals, so the removed ties are the lowest indices
      unsigned* bits = reinterpret_cast<unsigned*>(sortbuf);
      for (int i = tid; i < 1024; i += CS_THREADS) bits[i] = 0u;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < CS_ITEMS; ++j) {
        const int e = j * CS_THREADS + tid;
        if (v[j] == thr) atomicOr(&bits[e >> 5], 1u << (e & 31));
      }
      __syncthreads();
      const float above = block_suffix_float((float)__popc(bits[tid]), redf) - (float)__popc(bits[tid]);
      __syncthreads();
      float* above_s = sortbuf + 1024;
      above_s[tid] = above;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < CS_ITEMS; ++j) {
        const int e = j * CS_THREADS + tid;
        if (v[j] == thr) {
          const unsigned w = bits[e >> 5];
          const unsigned hi = (e & 31) == 31 ? 0u : (w >> ((e & 31) + 1));
          const int rank = __popc(hi) + (int)above_s[e >> 5];
          if (rank >= kept_ties) v[j] = -INFINITY;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CS_ITEMS; ++j)
      if (v[j] < thr) v[j] = -INFINITY;
  }

  // ---- scores (vcd_sample.py:200), softmax + multinomial (:201-202)
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {
    const int e = j * CS_THREADS + tid;
    if (scores_out && e < V) scores_out[e] = v[j];
  }
  if (q || a.use_philox) {
    // torch.multinomial(softmax(s), 1) == argmax_e softmax(s)_e / q_e == argmax_e (s_e - log q_e): the normaliser and the
    // row maximum cancel in the ORDER, so neither the sum of exponentials nor a second row reduction is needed.
#pragma unroll
    for (int j = 0; j < CS_ITEMS; ++j) {
      const int e = j * CS_THREADS + tid;
      // masked entries lose the race whatever their variate: skip their Philox rounds
      if (e < V) v[j] = q ? v[j] - logf(__ldg(q + e)) : (v[j] > -INFINITY ? v[j] - logf(philox_q(a, step, (long long)b * V + e)) : v[j]);
      else v[j] = -INFINITY;
    }
  }
  // argmax, lowest index on ties: pack (orderable key bits, ~index) into 64 bits
  unsigned long long best = 0ull;
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {
    const int e = j * CS_THREADS + tid;
    if (e < V) {
      unsigned kb = __float_as_uint(v[j]);
      kb = (kb & 0x80000000u) ? ~kb : (kb | 0x80000000u);
      const unsigned long long p = ((unsigned long long)kb << 32) | (unsigned)(0x7fffffff - e);
      best = p > best ? p : best;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t > best ? t : best;
  }
  if (tid == 0) best_s = 0ull;
  __syncthreads();
  if ((tid & 31) == 0) atomicMax(&best_s, best);
  __syncthreads();
  if (tid == 0) {
    long long tok = 0x7fffffff - (int)(best_s & 0xffffffffull);
    // finished rows emit pad (vcd_sample.py:257-260), then the unfinished mask drops rows that hit EOS (:285-288)
    if (a.eos >= 0 && a.unfinished) {
      const int unf = a.unfinished[b];
      tok = tok * unf + a.pad * (1 - unf);
      a.unfinished[b] = (unf != 0 && tok != a.eos) ? 1 : 0;
    }
    if (a.tokens_out) a.tokens_out[b] = tok;
    if (a.seq_out) a.seq_out[(int64_t)b * a.seq_ld + step] = tok;
  }
}

// ------------------------------------------------------------------------------------------------ cluster fast path
// When no ordering is needed (no top-p, and top-k absent or == 1) a row is split over a thread-block CLUSTER of 8 CTAs
// (256 threads, 16 values per thread): eight times the loads in flight per row and several rows resident per SM, with
// the four row-wide reductions (max l, max s, sum exp, arg-max of the race) done through distributed shared memory.
// Same arithmetic, same operation order per element as the single-CTA kernel => bit-identical scores and tokens.
namespace cg = cooperative_groups;
constexpr int CC_CL = 8, CC_THREADS = 256, CC_ITEMS = 16;  // V <= 8 * 256 * 16 = 32768

__device__ __forceinline__ float cluster_reduce_f(cg::cluster_group& cl, float v, float* slots, float* red, bool is_max) {
  v = is_max ? block_reduce<true>(v, red) : block_reduce<false>(v, red);
  if (threadIdx.x == 0) slots[0] = v;  // slots: this CTA's published value for this reduction
  cl.sync();
  float r = is_max ? -INFINITY : 0.f;
#pragma unroll
  for (int k = 0; k < CC_CL; ++k) {  // fixed order => every CTA computes the same bits
    const float o = *cl.map_shared_rank(slots, k);
    r = is_max ? fmaxf(r, o) : r + o;
  }
  return r;
}

__global__ void __cluster_dims__(CC_CL, 1, 1) __launch_bounds__(CC_THREADS, 4)
combine_sample_cluster_kernel(const SampleArgs byval, const SampleArgs* dptr) {
  __shared__ float redf[33];
  __shared__ float pub[4];                 // one published value per reduction (no slot is reused => no extra syncs)
  __shared__ unsigned long long pub_best;
  __shared__ unsigned long long best_s;
  cg::cluster_group cl = cg::this_cluster();
  const SampleArgs& a = dptr ? *dptr : byval;
  const int rank = (int)cl.block_rank();
  const int b = blockIdx.x / CC_CL, tid = threadIdx.x;
  const int V = a.V;
  const int step = a.step_ptr ? *a.step_ptr : 0;
  const float* l = a.l + (int64_t)b * a.ld;
  const float* u = a.u ? a.u + (int64_t)b * a.ld : nullptr;
  const float* d = a.d ? a.d + (int64_t)b * a.ld : nullptr;
  const int64_t so = ((int64_t)step * a.B + b) * V;
  const float* q = a.q ? a.q + so : nullptr;
  float* scores_out = a.scores_out ? a.scores_out + so : nullptr;
  const bool contrast = a.use_contrast && (u || d);
  const int e0 = rank * (CC_THREADS * CC_ITEMS) + tid;  // element j of this thread: e0 + j*256

  float v[CC_ITEMS], cu[CC_ITEMS], cd[CC_ITEMS];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < CC_ITEMS; ++j) {  // every load of the row slice is issued before anything is consumed
    const int e = e0 + j * CC_THREADS;
    v[j] = e < V ? __ldg(l + e) : -INFINITY;
    cu[j] = (contrast && u && e < V) ? __ldg(u + e) : 0.f;
    cd[j] = (contrast && d && e < V) ? __ldg(d + e) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < CC_ITEMS; ++j) mx = fmaxf(mx, v[j]);
  mx = cluster_reduce_f(cl, mx, &pub[0], redf, true);
  if (contrast) {
    const float cutoff = __fadd_rn(a.log_beta, mx);
    const float a1 = a.one_plus_alpha, al = a.alpha;
#pragma unroll
    for (int j = 0; j < CC_ITEMS; ++j) {
      const int e = e0 + j * CC_THREADS;
      if (e < V) {
        const float c = (u && d) ? __fmul_rn(__fadd_rn(cu[j], cd[j]), 0.5f) : (u ? cu[j] : cd[j]);
        const float s = __fsub_rn(__fmul_rn(a1, v[j]), __fmul_rn(al, c));
        v[j] = v[j] < cutoff ? -INFINITY : s;
      }
    }
  }
  if (a.temperature != 1.0f) {
#pragma unroll
    for (int j = 0; j < CC_ITEMS; ++j) v[j] = __fdiv_rn(v[j], a.temperature);
  }
  if (a.top_k == 1) {  // TopK(1): everything below the row maximum goes (ties stay)
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < CC_ITEMS; ++j) m = fmaxf(m, v[j]);
    m = cluster_reduce_f(cl, m, &pub[1], redf, true);
#pragma unroll
    for (int j = 0; j < CC_ITEMS; ++j)
      if (v[j] < m) v[j] = -INFINITY;
  }
#pragma unroll
  for (int j = 0; j < CC_ITEMS; ++j) {
    const int e = e0 + j * CC_THREADS;
    if (scores_out && e < V) scores_out[e] = v[j];
  }
  if (q || a.use_philox) {  // argmax_e softmax(s)_e / q_e == argmax_e (s_e - log q_e): no normaliser, no second row reduction
#pragma unroll
    for (int j = 0; j < CC_ITEMS; ++j) {
      const int e = e0 + j * CC_THREADS;
      if (e < V) v[j] = q ? v[j] - logf(__ldg(q + e)) : (v[j] > -INFINITY ? v[j] - logf(philox_q(a, step, (long long)b * V + e)) : v[j]);
      else v[j] = -INFINITY;
    }
  }
  unsigned long long best = 0ull;
#pragma unroll
  for (int j = 0; j < CC_ITEMS; ++j) {
    const int e = e0 + j * CC_THREADS;
    if (e < V) {
      unsigned kb = __float_as_uint(v[j]);
      kb = (kb & 0x80000000u) ? ~kb : (kb | 0x80000000u);
      const unsigned long long p = ((unsigned long long)kb << 32) | (unsigned)(0x7fffffff - e);
      best = p > best ? p : best;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t > best ? t : best;
  }
  if (tid == 0) best_s = 0ull;
  __syncthreads();
  if ((tid & 31) == 0) atomicMax(&best_s, best);
  __syncthreads();
  if (tid == 0) pub_best = best_s;
  cl.sync();
  if (rank == 0 && tid == 0) {
    unsigned long long bb = 0ull;
    for (int k = 0; k < CC_CL; ++k) {
      const unsigned long long o = *cl.map_shared_rank(&pub_best, k);
      bb = o > bb ? o : bb;
    }
    long long tok = 0x7fffffff - (int)(bb & 0xffffffffull);
    if (a.eos >= 0 && a.unfinished) {
      const int unf = a.unfinished[b];
      tok = tok * unf + a.pad * (1 - unf);
      a.unfinished[b] = (unf != 0 && tok != a.eos) ? 1 : 0;
    }
    if (a.tokens_out) a.tokens_out[b] = tok;
    if (a.seq_out) a.seq_out[(int64_t)b * a.seq_ld + step] = tok;
  }
  cl.sync();  // nobody leaves while a peer may still read its shared memory
}

// ------------------------------------------------------------------------------------------------ row-streaming fast path
// Many rows (B >= 8) and no ordering needed: one PERSISTENT CTA per SM walks the rows.  A row's main-stream logits are bulk-copied
// (cp.async.bulk, one 128 KB request) into shared memory, where they stay for the row maximum, are overwritten in place by the
// combined scores and serve the scores_out write; the debias streams and the race variates are streamed ONCE with 16-byte loads,
// eighteen per thread in flight (~150 KB per SM), and consumed on the fly.  Two block reductions per row (three with top_k = 1)
// instead of four cluster-wide ones, no DSMEM, no 14 waves of 8-CTA clusters.  Per-element arithmetic and its order are those of
// the kernels above => bit-identical scores and tokens.
constexpr int CR_THREADS = 512, CR_UNROLL = 6;

__device__ __forceinline__ unsigned long long race_key(float v, int e) {
  unsigned kb = __float_as_uint(v);
  kb = (kb & 0x80000000u) ? ~kb : (kb | 0x80000000u);
  return ((unsigned long long)kb << 32) | (unsigned)(0x7fffffff - e);
}

__global__ void __launch_bounds__(CR_THREADS, 1) combine_sample_rows_kernel(const SampleArgs byval, const SampleArgs* dptr) {
  extern __shared__ __align__(16) float srow[];  // V floats
  __shared__ float redf[33];
  __shared__ unsigned long long best_s;
  __shared__ __align__(8) uint64_t bar;
  const SampleArgs& a = dptr ? *dptr : byval;
  const int tid = threadIdx.x;
  const int V = a.V, nv4 = V >> 2;
  const int step = a.step_ptr ? *a.step_ptr : 0;
  const bool topk1 = a.top_k == 1;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  uint32_t parity = 0;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const float* l = a.l + (int64_t)b * a.ld;
    const float* u = a.u ? a.u + (int64_t)b * a.ld : nullptr;
    const float* d = a.d ? a.d + (int64_t)b * a.ld : nullptr;
    const int64_t so = ((int64_t)step * a.B + b) * V;
    const float* q = a.q ? a.q + so : nullptr;
    float* scores_out = a.scores_out ? a.scores_out + so : nullptr;
    const bool contrast = a.use_contrast && (u || d);
    if (tid == 0) {
      fence_proxy_async();  // the previous row's generic accesses to srow are ordered before this async write
      mbar_arrive_expect_tx(&bar, (uint32_t)V * 4u);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(srow)), "l"(reinterpret_cast<uint64_t>(l)), "r"((uint32_t)V * 4u), "r"(smem_u32(&bar)) : "memory");
    }
    // the debias streams and the variates do not depend on the row in shared memory: their first chunk is requested before the wait
    const float4* u4 = reinterpret_cast<const float4*>(u);
    const float4* d4 = reinterpret_cast<const float4*>(d);
    const float4* q4 = reinterpret_cast<const float4*>(q);
    float4 uu[CR_UNROLL], dd[CR_UNROLL], qq[CR_UNROLL];
    auto load_chunk = [&](int i0) {
#pragma unroll
      for (int j = 0; j < CR_UNROLL; ++j) {  // every load of the chunk is issued before anything is consumed
        const int i = i0 + j * CR_THREADS;
        const bool ok = i < nv4;
        uu[j] = (contrast && u && ok) ? __ldg(u4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        dd[j] = (contrast && d && ok) ? __ldg(d4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        qq[j] = (q && !topk1 && ok) ? __ldg(q4 + i) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
    };
    load_chunk(tid);
    mbar_wait(&bar, parity);
    parity ^= 1u;
    float4* s4 = reinterpret_cast<float4*>(srow);
    // ---- row maximum of the raw main-stream logits (vcd_sample.py:191)
    float mx = -INFINITY;
    for (int i = tid; i < nv4; i += CR_THREADS) {
      const float4 x = s4[i];
      mx = fmaxf(fmaxf(mx, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
    }
    mx = block_reduce<true>(mx, redf);
    const float cutoff = __fadd_rn(a.log_beta, mx);
    const float a1 = a.one_plus_alpha, al = a.alpha, temp = a.temperature;
    // ---- combine + mask (+ temperature) in place; the race rides along unless top_k = 1 has to mask first
    unsigned long long best = 0ull;
    float m2 = -INFINITY;
    for (int i0 = tid; i0 < nv4; i0 += CR_THREADS * CR_UNROLL) {
      if (i0 != tid) load_chunk(i0);
#pragma unroll
      for (int j = 0; j < CR_UNROLL; ++j) {
        const int i = i0 + j * CR_THREADS;
        if (i < nv4) {
          const float4 lv = s4[i];
          float v[4] = {lv.x, lv.y, lv.z, lv.w};
          const float cu[4] = {uu[j].x, uu[j].y, uu[j].z, uu[j].w}, cdv[4] = {dd[j].x, dd[j].y, dd[j].z, dd[j].w};
          const float qv[4] = {qq[j].x, qq[j].y, qq[j].z, qq[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (contrast) {
              const float c = (u && d) ? __fmul_rn(__fadd_rn(cu[e], cdv[e]), 0.5f) : (u ? cu[e] : cdv[e]);
              const float sc = __fsub_rn(__fmul_rn(a1, v[e]), __fmul_rn(al, c));
              v[e] = v[e] < cutoff ? -INFINITY : sc;
            }
            if (temp != 1.0f) v[e] = __fdiv_rn(v[e], temp);
            if (topk1) m2 = fmaxf(m2, v[e]);
            else {
              float rv = v[e];
              if (q) rv = v[e] - logf(qv[e]);
              else if (a.use_philox && v[e] > -INFINITY) rv = v[e] - logf(philox_q(a, step, (long long)b * V + i * 4 + e));
              best = max(best, race_key(rv, i * 4 + e));
            }
          }
          s4[i] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
    if (topk1) {  // TopK(1): everything below the row maximum goes (ties stay), then the race among the survivors
      m2 = block_reduce<true>(m2, redf);
      for (int i = tid; i < nv4; i += CR_THREADS) {
        const float4 sv = s4[i];
        float v[4] = {sv.x, sv.y, sv.z, sv.w};
        bool any = false;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (v[e] < m2) v[e] = -INFINITY;
          else any = true;
        }
        s4[i] = make_float4(v[0], v[1], v[2], v[3]);
        if (any) {
          float4 qv4 = q ? __ldg(q4 + i) : make_float4(1.f, 1.f, 1.f, 1.f);
          const float qv[4] = {qv4.x, qv4.y, qv4.z, qv4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float rv = v[e];
            if (q) rv = v[e] - logf(qv[e]);
            else if (a.use_philox && v[e] > -INFINITY) rv = v[e] - logf(philox_q(a, step, (long long)b * V + i * 4 + e));
            best = max(best, race_key(rv, i * 4 + e));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) best = max(best, race_key(-INFINITY, i * 4 + e));
        }
      }
    }
    // ---- scores (vcd_sample.py:200): every thread re-reads what it wrote itself
    if (scores_out) {
      float4* o4 = reinterpret_cast<float4*>(scores_out);
      for (int i = tid; i < nv4; i += CR_THREADS) o4[i] = s4[i];
    }
    // ---- arg-max of the race, lowest index on ties
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
      best = t > best ? t : best;
    }
    if (tid == 0) best_s = 0ull;
    __syncthreads();
    if ((tid & 31) == 0) atomicMax(&best_s, best);
    __syncthreads();
    if (tid == 0) {
      long long tok = 0x7fffffff - (int)(best_s & 0xffffffffull);
      if (a.eos >= 0 && a.unfinished) {
        const int unf = a.unfinished[b];
        tok = tok * unf + a.pad * (1 - unf);
        a.unfinished[b] = (unf != 0 && tok != a.eos) ? 1 : 0;
      }
      if (a.tokens_out) a.tokens_out[b] = tok;
      if (a.seq_out) a.seq_out[(int64_t)b * a.seq_ld + step] = tok;
    }
    __syncthreads();  // srow and best_s are reused by the next row
  }
}

cudaError_t launch_combine_sample(const SampleArgs& args, const SampleArgs* dptr, int B, int V, cudaStream_t stream) {
  // host copy of the warper settings decides the kernel (the device struct, when used, holds the same values)
  const bool needs_order = args.top_p < 1.0f || (args.top_k > 1 && args.top_k < V);
  // measured (us per launch, 3 streams + q): B = 32: cluster 10.9 / rows 15.3;  B = 64: 13.6 / 15.5;  B = 1024: 268 / 123
  static const int rows_min = getenv("VDD_K8_ROWS_MIN") ? atoi(getenv("VDD_K8_ROWS_MIN")) : 96;
  const bool aligned = (V % 4) == 0 && (args.ld % 4) == 0 && !((uintptr_t)args.l & 15) && !((uintptr_t)args.u & 15) &&
                       !((uintptr_t)args.d & 15) && !((uintptr_t)args.q & 15) && !((uintptr_t)args.scores_out & 15);
  if (!needs_order && aligned && B >= rows_min && V <= 32768) {
    static int num_sms = 0;
    static bool attr = false;
    if (!attr) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
      cudaError_t e = cudaFuncSetAttribute(combine_sample_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4);
      if (e != cudaSuccess) return e;
      attr = true;
    }
    combine_sample_rows_kernel<<<B < num_sms ? B : num_sms, CR_THREADS, (size_t)V * 4, stream>>>(args, dptr);
    return cudaGetLastError();
  }
  if (!needs_order && V <= CC_CL * CC_THREADS * CC_ITEMS && getenv("VDD_K8_SINGLE_CTA") == nullptr) {
    combine_sample_cluster_kernel<<<B * CC_CL, CC_THREADS, 0, stream>>>(args, dptr);
    return cudaGetLastError();
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(combine_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int P = 1;
  while (P < V) P <<= 1;
  if (P < 2048) P = 2048;  // the tie-split path uses 2048 words
  combine_sample_kernel<<<B, CS_THREADS, (size_t)P * 4, stream>>>(args, dptr);
  return cudaGetLastError();
}

}  // namespace vdd
