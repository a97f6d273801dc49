// Decode-step projection for R <= 8 fused rows (all debias streams of all sequences): y[r, n] = sum_k x[r,k] W[n,k].
//
// This is the HBM-bound heart of the decode step (reference: one cuBLAS GEMV per stream per projection inside HF
// LlamaModel, entered from experiments/llava/model/language_model/llava_llama.py:91-103 once PER STREAM per token;
// here the weights are streamed ONCE for all R rows).  Algorithmic bytes = N*K*2.
//
// Layout: W bf16 [N, K] row-major.  Each CTA owns a contiguous band of output features; inside the CTA a warp
// task is (8 consecutive features) x (a K-slice).  Lane (g, t) = (lane/4, lane%4) streams 16-byte pieces of weight
// row g straight from global memory (L1 no-allocate): one warp step covers 8 rows x 256 k, every row contributing
// 512 contiguous bytes, with the next step's 8 loads per lane in flight while the current 8 are consumed.
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
// gate/up rows.  With programmatic dependent launch the first weight loads are issued BEFORE griddepcontrol.wait:
// weights do not depend on the previous kernel, so the stream of weight bytes continues across kernel boundaries.
#include "common.cuh"
#include "kernels.h"

namespace vdd {

constexpr int GV_THREADS = 256;
constexpr int GV_WARPS = GV_THREADS / 32;
constexpr int GV_RG = 8;     // weight rows per warp task (the N side of m16n8k16)
constexpr int GV_UNIT = 256; // k elements per warp step (8 pieces of 32 k; lane t takes 8 of each 32)
constexpr int GV_MAX_TASKS = 256;
constexpr int GV_XPAD = 32;  // bf16 elements of padding per x row: rows land 64 B apart modulo 128 (no LDS conflicts)

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

struct GemvPlan {
  int n0, n1;      // CTA's feature band
  int ngroups, ks, ntasks, nunits;
};
__device__ __forceinline__ GemvPlan make_plan(int N, int K, int cta, int nctas) {
  GemvPlan p;
  const int pairs = N >> 1;
  p.n0 = 2 * (int)(((long long)cta * pairs) / nctas);
  p.n1 = 2 * (int)(((long long)(cta + 1) * pairs) / nctas);
  const int nrows = p.n1 - p.n0;
  p.ngroups = (nrows + GV_RG - 1) / GV_RG;
  p.nunits = (K + GV_UNIT - 1) / GV_UNIT;
  int ks = p.ngroups > 0 ? (4 * GV_WARPS + p.ngroups - 1) / p.ngroups : 1;
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

template <int R, int PRO, int EPI>
__global__ void __launch_bounds__(GV_THREADS) gemv_kernel(const GemvArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int K = a.K;
  const int XS = K + GV_XPAD;                                                 // x row stride (elements)
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem);                 // [R][XS]
  float* part = reinterpret_cast<float*>(smem + (size_t)R * XS * 2);          // [nrows][ks][R]
  __shared__ float red[33];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const GemvPlan p = make_plan(a.N, K, blockIdx.x, gridDim.x);
  const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(a.w);

  // ---- task table: (row group, first unit, end unit, k-slice) per task, computed once (the divisions stay out of the
  //      streaming loop, whose instruction count per 4 KB step is what bounds a warp's bandwidth)
  __shared__ int4 task_tab[GV_MAX_TASKS];
  for (int i = tid; i < p.ntasks; i += GV_THREADS) {
    const int gi = i / p.ks, ji = i - gi * p.ks;
    task_tab[i] = make_int4(gi, (ji * p.nunits) / p.ks, ((ji + 1) * p.nunits) / p.ks, ji);
  }
  __syncthreads();
  int task = warp;
  int tg = 0, j = 0, u = 0, u_end = 0;
  auto load_step = [&](uint4 (&w)[8], int ltg, int lu) {
    int row = p.n0 + ltg * GV_RG + g;
    row = row < p.n1 ? row : p.n1 - 1;
    const __nv_bfloat16* wr = W + (size_t)row * K + lu * GV_UNIT + t * 8;
    if ((lu + 1) * GV_UNIT <= K) {
#pragma unroll
      for (int c = 0; c < 8; ++c) w[c] = ldg_stream(wr + c * 32);
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const bool ok = lu * GV_UNIT + c * 32 + t * 8 < K;
        w[c] = ok ? ldg_stream(wr + c * 32) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };

  uint4 wc[8];
  if (task < p.ntasks) {
    const int4 tk = task_tab[task];
    tg = tk.x; u = tk.y; u_end = tk.z; j = tk.w;
    load_step(wc, tg, u);  // weights do not depend on the previous kernel: issue before the PDL wait
  }
  pdl_wait();
  pdl_launch_dependents();

  // ---- prologue: stage x (bf16) in shared memory
  if constexpr (PRO == GEMV_X_RMSNORM) {
    const __nv_bfloat16* nw = reinterpret_cast<const __nv_bfloat16*>(a.norm_w);
    float ss[R];
#pragma unroll
    for (int r = 0; r < R; ++r) ss[r] = 0.f;
    for (int k = tid * 4; k < K; k += GV_THREADS * 4) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(a.x_resid + (size_t)r * K + k);
        ss[r] += xv.x * xv.x + xv.y * xv.y + xv.z * xv.z + xv.w * xv.w;
      }
    }
    float rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(block_reduce<false>(ss[r], red) / (float)K + a.eps);
    for (int k = tid * 4; k < K; k += GV_THREADS * 4) {
      const uint2 wv = *reinterpret_cast<const uint2*>(nw + k);
      const float w0 = bf16lo(wv.x), w1 = bf16hi(wv.x), w2 = bf16lo(wv.y), w3 = bf16hi(wv.y);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(a.x_resid + (size_t)r * K + k);
        uint2 o;
        o.x = pack_bf16(xv.x * rstd[r] * w0, xv.y * rstd[r] * w1);
        o.y = pack_bf16(xv.z * rstd[r] * w2, xv.w * rstd[r] * w3);
        *reinterpret_cast<uint2*>(xs + (size_t)r * XS + k) = o;
      }
    }
  } else {
    const int cpr = K / 8;  // 16-byte chunks per row
    for (int i = tid; i < R * cpr; i += GV_THREADS) {
      const int r = i / cpr, c = i - r * cpr;
      *reinterpret_cast<uint4*>(xs + (size_t)r * XS + c * 8) =
          *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.x_bf16) + (size_t)r * K + c * 8);
    }
  }
  __syncthreads();

  // ---- main loop: software-pipelined over (task, step); lane (g, t) accumulates y[r = g][features 2t, 2t+1 of the group]
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};  // two independent MMA chains
  const __nv_bfloat16* xrow = xs + (size_t)(g < R ? g : 0) * XS + t * 8;
  while (task < p.ntasks) {
    int ntask = task, ntg = tg, nu = u + 1, nu_end = u_end, nj = j;
    bool task_done = false;
    if (nu >= u_end) {
      task_done = true;
      ntask = task + GV_WARPS;
      if (ntask < p.ntasks) {
        const int4 tk = task_tab[ntask];
        ntg = tk.x; nu = tk.y; nu_end = tk.z; nj = tk.w;
      }
    }
    uint4 wn[8];
    if (ntask < p.ntasks) load_step(wn, ntg, nu);

    {
      const __nv_bfloat16* xp = xrow + u * GV_UNIT;
      const bool full = (u + 1) * GV_UNIT <= K;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        // lanes with g >= R read row 0: their products land in accumulator rows nobody stores
        uint4 xv = make_uint4(0u, 0u, 0u, 0u);
        if (full || u * GV_UNIT + c * 32 + t * 8 < K) xv = *reinterpret_cast<const uint4*>(xp + c * 32);
        mma_bf16_16816(acc, xv.x, xv.y, wc[c].x, wc[c].y);
        mma_bf16_16816(acc2, xv.z, xv.w, wc[c].z, wc[c].w);
      }
    }
    if (task_done) {
      if (g < R) {
        const int row_local = tg * GV_RG + 2 * t;
        if (p.n0 + row_local < p.n1) part[((size_t)row_local * p.ks + j) * R + g] = acc[0] + acc2[0];
        if (p.n0 + row_local + 1 < p.n1) part[((size_t)(row_local + 1) * p.ks + j) * R + g] = acc[1] + acc2[1];
      }
      acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
      acc2[0] = acc2[1] = acc2[2] = acc2[3] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) wc[c] = wn[c];
    task = ntask; tg = ntg; u = nu; u_end = nu_end; j = nj;
  }
  __syncthreads();

  // ---- epilogue: fixed-order sum over the K-slices
  const int nrows = p.n1 - p.n0;
  if constexpr (EPI == GEMV_OUT_SWIGLU_BF16) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
    for (int i = tid; i < (nrows / 2) * R; i += GV_THREADS) {
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
    for (int i = tid; i < nrows * R; i += GV_THREADS) {
      const int row = i / R, r = i - row * R;
      float s = 0.f;
      for (int q = 0; q < p.ks; ++q) s += part[((size_t)row * p.ks + q) * R + r];
      float* o = a.out_f32 + (size_t)r * a.ld_out + p.n0 + row;
      if constexpr (EPI == GEMV_OUT_ADD_RESID) *o += s; else *o = s;
    }
  }
}

template <int R, int PRO, int EPI>
static cudaError_t launch_r(const GemvArgs& a, int grid, size_t smem, bool pdl, cudaStream_t stream) {
  auto kern = gemv_kernel<R, PRO, EPI>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GV_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

template <int R>
static cudaError_t launch_pe(const GemvArgs& a, int pro, int grid, size_t smem, bool pdl, cudaStream_t s) {
  if (pro == GEMV_X_RMSNORM) {
    if (a.epilogue == GEMV_OUT_F32) return launch_r<R, GEMV_X_RMSNORM, GEMV_OUT_F32>(a, grid, smem, pdl, s);
    if (a.epilogue == GEMV_OUT_SWIGLU_BF16) return launch_r<R, GEMV_X_RMSNORM, GEMV_OUT_SWIGLU_BF16>(a, grid, smem, pdl, s);
  } else {
    if (a.epilogue == GEMV_OUT_F32) return launch_r<R, GEMV_X_BF16, GEMV_OUT_F32>(a, grid, smem, pdl, s);
    if (a.epilogue == GEMV_OUT_ADD_RESID) return launch_r<R, GEMV_X_BF16, GEMV_OUT_ADD_RESID>(a, grid, smem, pdl, s);
    if (a.epilogue == GEMV_OUT_SWIGLU_BF16) return launch_r<R, GEMV_X_BF16, GEMV_OUT_SWIGLU_BF16>(a, grid, smem, pdl, s);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_gemv(const GemvArgs& a, int num_sms, bool pdl, cudaStream_t stream) {
  if (a.R < 1 || a.R > 8 || (a.K % 8) != 0 || (a.N % 2) != 0) return cudaErrorInvalidValue;
  const int pro = a.x_resid ? GEMV_X_RMSNORM : GEMV_X_BF16;
  // two CTAs per SM when x fits twice, else one
  const size_t xbytes = (size_t)a.R * (a.K + GV_XPAD) * 2;
  int per_sm = xbytes <= 100 * 1024 ? 2 : 1;
  int grid = num_sms * per_sm;
  const int pairs = a.N / 2;
  if (grid > pairs) grid = pairs;
  const int nrows_max = 2 * ((pairs + grid - 1) / grid) + 2;
  const size_t part = (size_t)(4 * GV_WARPS * GV_RG + 2 * nrows_max + 64) * a.R * 4;  // >= rows*ks of any CTA
  const size_t smem = xbytes + part;
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  switch (a.R) {
    case 1: return launch_pe<1>(a, pro, grid, smem, pdl, stream);
    case 2: return launch_pe<2>(a, pro, grid, smem, pdl, stream);
    case 3: return launch_pe<3>(a, pro, grid, smem, pdl, stream);
    case 4: return launch_pe<4>(a, pro, grid, smem, pdl, stream);
    case 5: return launch_pe<5>(a, pro, grid, smem, pdl, stream);
    case 6: return launch_pe<6>(a, pro, grid, smem, pdl, stream);
    case 7: return launch_pe<7>(a, pro, grid, smem, pdl, stream);
    case 8: return launch_pe<8>(a, pro, grid, smem, pdl, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace vdd
