// Decode-step projection for R <= 8 fused rows (all debias streams of all sequences): y[r, n] = sum_k x[r,k] W[n,k].
//
// This is the HBM-bound heart of the decode step (reference: one cuBLAS GEMV per stream per projection inside HF
// LlamaModel, entered from experiments/llava/model/language_model/llava_llama.py:91-103 once PER STREAM per token;
// here the weights are streamed ONCE for all R rows).  Algorithmic bytes = N*K*2.
//
// Layout: W bf16 [N, K] row-major.  Each CTA owns a contiguous band of output features; inside the CTA a warp
// task is (4 consecutive features) x (a K-slice); a lane streams 16-byte pieces of 4 weight rows straight from
// global memory (coalesced 512 B per warp per row, L1 no-allocate) with the next step's 8 loads in flight while
// the current 8 are consumed; x (bf16, R rows) sits in shared memory and is reused by the 4 rows.  Partial sums
// are reduced with a halving shuffle network, parked in shared memory and summed in a fixed order => results are
// deterministic run to run.  Prologue options: plain bf16 x, or RMSNorm(residual) computed on the fly;
// epilogues: fp32 store, residual +=, SwiGLU over interleaved gate/up rows.
// With programmatic dependent launch the first weight loads are issued BEFORE griddepcontrol.wait: weights do
// not depend on the previous kernel, so the stream of weight bytes continues across kernel boundaries.
#include "common.cuh"
#include "kernels.h"

namespace vdd {

constexpr int GV_THREADS = 256;
constexpr int GV_WARPS = GV_THREADS / 32;
constexpr int GV_RG = 4;     // weight rows per warp task
constexpr int GV_UNIT = 256; // k elements per warp step-unit (32 lanes x 8 bf16)

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <int N, int O>
struct HalvingReduce {
  static __device__ __forceinline__ void run(float* v, int lane) {
    if constexpr (O >= 1) {
      if constexpr (N > 1) {
        constexpr int h = N / 2;
        const bool upper = (lane & O) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
          const float send = upper ? v[i] : v[i + h];
          const float keep = upper ? v[i + h] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
        }
        HalvingReduce<h, O / 2>::run(v, lane);
      } else {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], O);
        HalvingReduce<1, O / 2>::run(v, lane);
      }
    }
  }
};
__host__ __device__ constexpr int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

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

template <int R, int UU>
struct StepLoads {
  uint4 w[UU][GV_RG];
};

template <int R, int PRO, int EPI, int UU>
__global__ void __launch_bounds__(GV_THREADS) gemv_kernel(const GemvArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem);                 // [R][K]
  float* part = reinterpret_cast<float*>(smem + (size_t)R * a.K * 2);         // [nrows][ks][R]
  __shared__ float red[33];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K;
  const GemvPlan p = make_plan(a.N, K, blockIdx.x, gridDim.x);
  const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(a.w);

  // ---- iterator over this warp's (task, step) sequence
  int task = warp;
  int g = 0, j = 0, u = 0, u_end = 0;
  auto open_task = [&](int t) {
    g = t / p.ks;
    j = t - g * p.ks;
    u = (int)(((long long)j * p.nunits) / p.ks);
    u_end = (int)(((long long)(j + 1) * p.nunits) / p.ks);
  };
  auto load_step = [&](uint4 (&w)[UU][GV_RG], int tg, int tu, int tu_end) {
#pragma unroll
    for (int uu = 0; uu < UU; ++uu) {
      const int k = (tu + uu) * GV_UNIT + lane * 8;
      const bool ok = (tu + uu) < tu_end && k < K;
#pragma unroll
      for (int i = 0; i < GV_RG; ++i) {
        int row = p.n0 + tg * GV_RG + i;
        row = row < p.n1 ? row : p.n1 - 1;
        w[uu][i] = ok ? ldg_stream(W + (size_t)row * K + k) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };

  uint4 wc[UU][GV_RG];
  const bool have = task < p.ntasks;
  if (have) {
    open_task(task);
    load_step(wc, g, u, u_end);  // weights do not depend on the previous kernel: issue before the PDL wait
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
        *reinterpret_cast<uint2*>(xs + (size_t)r * K + k) = o;
      }
    }
  } else {
    const uint4* src = reinterpret_cast<const uint4*>(a.x_bf16);
    uint4* dst = reinterpret_cast<uint4*>(xs);
    for (int i = tid; i < R * K / 8; i += GV_THREADS) dst[i] = src[i];
  }
  __syncthreads();

  // ---- main loop: software-pipelined over (task, step)
  constexpr int NV = next_pow2(GV_RG * R);
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;

  while (task < p.ntasks) {
    // where does the NEXT step live?
    int ntask = task, ng = g, nu = u + UU, nu_end = u_end;
    bool task_done = false;
    if (nu >= u_end) {
      task_done = true;
      ntask = task + GV_WARPS;
      if (ntask < p.ntasks) {
        ng = ntask / p.ks;
        const int nj = ntask - ng * p.ks;
        nu = (int)(((long long)nj * p.nunits) / p.ks);
        nu_end = (int)(((long long)(nj + 1) * p.nunits) / p.ks);
      }
    }
    uint4 wn[UU][GV_RG];
    if (ntask < p.ntasks) load_step(wn, ng, nu, nu_end);

    // consume the current step
#pragma unroll
    for (int uu = 0; uu < UU; ++uu) {
      const int k = (u + uu) * GV_UNIT + lane * 8;
      if ((u + uu) < u_end && k < K) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)r * K + k);
          const float x0 = bf16lo(xv.x), x1 = bf16hi(xv.x), x2 = bf16lo(xv.y), x3 = bf16hi(xv.y);
          const float x4 = bf16lo(xv.z), x5 = bf16hi(xv.z), x6 = bf16lo(xv.w), x7 = bf16hi(xv.w);
#pragma unroll
          for (int i = 0; i < GV_RG; ++i) {
            const uint4 wv = wc[uu][i];
            float s = acc[i * R + r];
            s = fmaf(bf16lo(wv.x), x0, s);
            s = fmaf(bf16hi(wv.x), x1, s);
            s = fmaf(bf16lo(wv.y), x2, s);
            s = fmaf(bf16hi(wv.y), x3, s);
            s = fmaf(bf16lo(wv.z), x4, s);
            s = fmaf(bf16hi(wv.z), x5, s);
            s = fmaf(bf16lo(wv.w), x6, s);
            s = fmaf(bf16hi(wv.w), x7, s);
            acc[i * R + r] = s;
          }
        }
      }
    }
    if (task_done) {
      HalvingReduce<NV, 16>::run(acc, lane);
      constexpr int LPV = 32 / NV;  // lanes per value
      const int idx = lane / LPV;
      if ((lane % LPV) == 0 && idx < GV_RG * R) {
        const int i = idx / R, r = idx - i * R;
        const int row_local = g * GV_RG + i;
        if (p.n0 + row_local < p.n1) part[((size_t)row_local * p.ks + j) * R + r] = acc[0];
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    }
#pragma unroll
    for (int uu = 0; uu < UU; ++uu)
#pragma unroll
      for (int i = 0; i < GV_RG; ++i) wc[uu][i] = wn[uu][i];
    task = ntask; g = ng; u = nu; u_end = nu_end;
    if (task_done && task < p.ntasks) j = task - g * p.ks;
  }
  __syncthreads();

  // ---- epilogue: fixed-order sum over the K-slices
  const int nrows = p.n1 - p.n0;
  if constexpr (EPI == GEMV_OUT_SWIGLU_BF16) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
    for (int t = tid; t < (nrows / 2) * R; t += GV_THREADS) {
      const int pr = t / R, r = t - pr * R;
      float gsum = 0.f, usum = 0.f;
      for (int s = 0; s < p.ks; ++s) {
        gsum += part[((size_t)(2 * pr) * p.ks + s) * R + r];
        usum += part[((size_t)(2 * pr + 1) * p.ks + s) * R + r];
      }
      const float h = gsum / (1.0f + expf(-gsum)) * usum;
      out[(size_t)r * a.ld_out + (p.n0 >> 1) + pr] = __float2bfloat16_rn(h);
    }
  } else {
    for (int t = tid; t < nrows * R; t += GV_THREADS) {
      const int row = t / R, r = t - row * R;
      float s = 0.f;
      for (int q = 0; q < p.ks; ++q) s += part[((size_t)row * p.ks + q) * R + r];
      float* o = a.out_f32 + (size_t)r * a.ld_out + p.n0 + row;
      if constexpr (EPI == GEMV_OUT_ADD_RESID) *o += s; else *o = s;
    }
  }
}

template <int R, int PRO, int EPI>
static cudaError_t launch_r(const GemvArgs& a, int grid, size_t smem, bool pdl, cudaStream_t stream) {
  auto kern = gemv_kernel<R, PRO, EPI, (R <= 4 ? 2 : 1)>;
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
  const size_t xbytes = (size_t)a.R * a.K * 2;
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
