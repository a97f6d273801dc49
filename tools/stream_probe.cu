// Micro-benchmark (not part of the product): what does a chain of read-only kernels over the decode step's weight
// matrices cost on this GPU, as a function of access pattern and launch mode?  Establishes the attainable ceiling for
// gemv_kernel's chain.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// pattern 0: every CTA sweeps a contiguous band, all its warps side by side (512 B per warp per load, 8 loads in flight)
// pattern 1: GEMV-like: warp task = 8 rows x 256 k pieces (lane (g,t): row g, 16 B at t*16 + c*64)
template <int PATTERN, bool PDL>
__global__ void __launch_bounds__(512) read_kernel(const uint4* __restrict__ w, size_t n16, int K, unsigned* sink) {
  if (PDL) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned acc = 0;
  const size_t per_cta = (n16 + gridDim.x - 1) / gridDim.x;
  const size_t b0 = (size_t)blockIdx.x * per_cta, b1 = min(n16, b0 + per_cta);
  if (PATTERN == 0) {
    for (size_t i = b0 + (size_t)warp * 32 * 8 + lane; i < b1; i += (size_t)nw * 32 * 8) {
      uint4 v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = (i + c * 32 < b1) ? ldg_stream(w + i + c * 32) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc ^= v[c].x ^ v[c].y ^ v[c].z ^ v[c].w;
    }
  } else {
    // band of rows; K in bf16 elements; row = K/8 uint4
    const int rowq = K / 8;
    const size_t rows0 = b0 / rowq, rows1 = b1 / rowq;
    const int g = lane >> 2, t = lane & 3;
    const int units = rowq / 32;  // 256-element units per row
    const int ngroups = (int)((rows1 - rows0) / 8);
    const int ntasks = ngroups * units;
    for (int task = warp; task < ntasks; task += nw) {
      const int gi = task / units, u = task - gi * units;
      const uint4* p = w + (rows0 + gi * 8 + g) * rowq + u * 32 + t;
      uint4 v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = ldg_stream(p + c * 4);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc ^= v[c].x ^ v[c].y ^ v[c].z ^ v[c].w;
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <int PATTERN, bool PDL>
float run_chain(const uint4* buf, const std::vector<size_t>& offs, const std::vector<size_t>& sizes, const std::vector<int>& Ks,
                int grid, int threads, unsigned* sink, int reps) {
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps + 1; ++r) {
    CK(cudaEventRecord(e0, st));
    for (size_t i = 0; i < offs.size(); ++i) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
      CK(cudaLaunchKernelEx(&cfg, read_kernel<PATTERN, PDL>, buf + offs[i] / 16, sizes[i] / 16, Ks[i], sink));
    }
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaStreamDestroy(st);
  return best;
}

int main() {
  // one decode step of LLaVA-1.5-7B: per layer qkv 100.7 MB, o 33.6, gate_up 180.4, down 90.2; lm_head 262.1
  const size_t H = 4096, I = 11008, V = 32000;
  std::vector<size_t> sizes, offs; std::vector<int> Ks;
  size_t off = 0;
  for (int l = 0; l < 32; ++l) {
    size_t s[4] = {3 * H * H * 2, H * H * 2, 2 * I * H * 2, H * I * 2};
    int k[4] = {(int)H, (int)H, (int)H, (int)I};
    for (int i = 0; i < 4; ++i) { sizes.push_back(s[i]); offs.push_back(off); Ks.push_back(k[i]); off += s[i]; }
  }
  sizes.push_back(V * H * 2); offs.push_back(off); Ks.push_back((int)H); off += V * H * 2;
  uint4* buf; unsigned* sink;
  CK(cudaMalloc(&buf, off)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 1, off));
  const double gb = off / 1e9;
  printf("bytes per step %.3f GB, %zu launches\n", gb, sizes.size());
  int cfgs[][2] = {{148, 512}, {296, 256}, {444, 256}, {592, 128}, {148, 1024}};
  for (auto& c : cfgs) {
    float a = run_chain<0, false>(buf, offs, sizes, Ks, c[0], c[1], sink, 3);
    float b = run_chain<0, true>(buf, offs, sizes, Ks, c[0], c[1], sink, 3);
    float d = run_chain<1, false>(buf, offs, sizes, Ks, c[0], c[1], sink, 3);
    float e = run_chain<1, true>(buf, offs, sizes, Ks, c[0], c[1], sink, 3);
    printf("grid %4d x %4d thr | contiguous: %.3f ms (%.0f GB/s)  +PDL %.3f ms (%.0f GB/s) | gemv-pattern: %.3f ms (%.0f GB/s)  +PDL %.3f ms (%.0f GB/s)\n",
           c[0], c[1], a, gb / a * 1e3, b, gb / b * 1e3, d, gb / d * 1e3, e, gb / e * 1e3);
  }
  // one big kernel over everything = no boundaries at all
  {
    std::vector<size_t> o1{0}, s1{off}; std::vector<int> k1{4096};
    float a = run_chain<0, false>(buf, o1, s1, k1, 296, 256, sink, 3);
    printf("single launch over %.2f GB: %.3f ms (%.0f GB/s)\n", gb, a, gb / a * 1e3);
  }
  return 0;
}
