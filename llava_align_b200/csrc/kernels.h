// Host-visible launch wrappers of the VDD kernels (internal to libvdd; the public surface is include/vdd.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct CUtensorMap_st;  // <cuda.h>'s CUtensorMap, for files that do not include it

namespace vdd {

// ---------------------------------------------------------------- K8 combine + sample
struct SampleArgs {
  const float *l, *u, *d;  // [B, V] with row pitch ld
  const float* q;          // [steps, B, V] Exp(1) variates or null
  float* scores_out;       // [steps, B, V] or null
  int64_t* tokens_out;     // [B] or null
  int64_t* seq_out;        // [B, seq_ld]: column `step` receives the token, or null
  int32_t* unfinished;     // [B] or null
  const int32_t* step_ptr; // device step counter (graph replay) or null => step 0
  int64_t ld, seq_ld, eos, pad;
  int B, V, top_k, use_contrast;
  float one_plus_alpha, alpha, log_beta, temperature, top_p;
  // q == null and use_philox: the Exp(1) variates are generated in the kernel, bit for bit the ones the reference draws --
  // torch.multinomial(probs, 1) on CUDA = probs.div(empty_like(probs).exponential_(1)).argmax: call number `step` of
  // exponential_ on a [B, V] fp32 tensor with the generator at (philox_seed, philox_offset + step * philox_stride)
  int use_philox;
  unsigned philox_threads;             // threads of torch's distribution kernel for B*V elements (256 * grid.x)
  unsigned long long philox_seed, philox_offset, philox_stride;
};
// fill philox_threads / philox_stride the way ATen's calc_execution_policy does for numel elements on this device
void sample_philox_geometry(SampleArgs& a, long long numel, int num_sms, int max_threads_per_sm);
cudaError_t launch_philox_exponential(unsigned long long seed, unsigned long long offset, long long numel, int num_sms,
                                      int max_threads_per_sm, float* out, cudaStream_t stream);  // debug / tests: q as torch draws it
cudaError_t launch_combine_sample(const SampleArgs& args, const SampleArgs* dptr, int B, int V, cudaStream_t stream);

// ---------------------------------------------------------------- decode GEMV (R <= 8 rows of x against W[N,K])
enum GemvPrologue { GEMV_X_BF16 = 0, GEMV_X_RMSNORM = 1 };
enum GemvEpilogue {
  GEMV_OUT_F32 = 0,
  GEMV_OUT_ADD_RESID = 1,
  GEMV_OUT_SWIGLU_BF16 = 2,
  GEMV_OUT_TP_PARTIAL = 3,  // row-parallel projection under tensor parallelism: partial sums go to EVERY rank's inbox over NVLink;
                            // the next phase of the chain waits for all ranks, reduces into out_f32 (the residual) and goes on
  GEMV_OUT_TP_LOGITS = 4    // vocab-parallel lm_head: this rank's logit columns are stored into every rank's full logits
};
struct GemvArgs {
  const void* w;        // bf16 [N, K] row-major
  const void* x_bf16;   // bf16 [R, K] (GEMV_X_BF16)
  const float* x_resid; // fp32 [R, K] residual stream (GEMV_X_RMSNORM)
  const void* norm_w;   // bf16 [K]
  float eps;
  int R, N, K;
  int epilogue;
  float* out_f32;       // [R, N] (OUT_F32: store; ADD_RESID: +=)
  void* out_bf16;       // [R, N/2] (SWIGLU)
  int64_t ld_out;       // row pitch of the output in elements
  int tp_slot = 0;      // TP epilogues: inbox slot (0/1) ...
  int tp_counter = 0;   // ... and index of the cross-GPU arrival counter
};
cudaError_t launch_gemv(const GemvArgs& a, int num_sms, bool pdl, cudaStream_t stream);
// Peer-memory view of a tensor-parallel group: every rank owns one communication buffer
//   [0, 4096)            arrival counters (u32), monotonically increasing: target = (epoch + 1) * tp * grid
//   [off_inbox, ...)     inbox  [2 slots][tp sources][8 rows][H] fp32
//   [off_logits, ...)    logits [8 rows][V] fp32
// and can address all of them (cudaIpc-mapped) at peers[rank].
struct TpComm {
  int tp = 1, rank = 0, H = 0, V = 0, Vl = 0;
  unsigned long long peers[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long off_inbox = 0, off_logits = 0;
  const int32_t* epoch = nullptr;
};
// up to four projections in ONE launch, separated by grid barriers (barrier: n-1 counters, zero on entry)
struct GemvChain {
  GemvArgs ph[4];
  int n;
  int ksmul = 8;  // set by the launcher
  unsigned* barrier;
  TpComm tpc;
  unsigned long long* trace;  // debug: [grid][4 phases][8 slots] globaltimer stamps of thread 0, or null
};
cudaError_t launch_gemv_chain(const GemvChain& ch, int num_sms, bool pdl, cudaStream_t stream);

// ---------------------------------------------------------------- whole decode step in ONE persistent launch (R <= 4 rows)
// decode_step.cu: a producer warp per CTA streams the weights of every projection with cp.async.bulk into shared-memory rings
// while the consumer warps run  qkv_0 | per layer: attention | o | gate_up | down | next qkv (or lm_head)  behind grid barriers.
enum StepPhaseKind { STEP_GEMV = 0, STEP_ATTN = 1 };
struct StepPhase {
  int kind;
  GemvArgs g;            // STEP_GEMV
  void* kcache;          // STEP_ATTN: this layer's K / V planes
  void* vcache;
  // STEP_GEMV: TMA descriptors (CUtensorMap, 128 B each) of the weight matrix seen as [N][K/64][64] bf16 with 128-byte swizzle and
  // a box of 64 x 4 x {2,4,6,8} rows: ONE cp.async.bulk.tensor per 4 KB step (8 rows x 256 k), the last group of a CTA's band
  // (bands are pairs of rows) fetched with the narrower box instead of over-fetching the neighbour's rows
  alignas(64) unsigned char tmap[4][128];
  int tmap_plain;        // 1: K is not a multiple of 64 (7B at TP=8): plain 2-D box of 256 k x rows, no swizzle (2-way bank conflicts)
};
struct StepAttn {        // layer-independent part of the attention phases (see DecodeAttnArgs)
  const float* qkv;
  const int64_t* kv_off;
  const int32_t* kv_cap;
  const int32_t* pos;
  const float* rope_cos;
  const float* rope_sin;
  void* out_bf16;
  float* part_o;
  float* part_ml;
  int32_t* counters;
  int heads, max_splits;
  float scale;
};
struct StepArgs {
  const StepPhase* phases;  // DEVICE array [nph]
  int nph;
  int skip_attn;            // bench: projections only (the barriers of the attention phases stay)
  StepAttn attn;
  unsigned* barrier;        // one monotonic grid-barrier counter, zero on entry
  TpComm tpc;
  unsigned long long* trace;  // debug: [grid][64 phases][8] globaltimer stamps of consumer thread 0, or null
  int ring_depth, x_bytes;    // filled by launch_decode_step from the StepLaunch
  int l2_ahead;               // steps per warp kept in flight towards L2 beyond the shared-memory ring
  int l2_hint;                // 1: weights are loaded with the L2 evict-first policy
};
struct StepLaunch {
  int grid, nw, depth, x_bytes;
  size_t smem;
};
bool step_kernel_supports(const StepPhase* host_phases, int nph, int R);
cudaError_t step_build_tmaps(StepPhase* host_phases, int nph);  // fills tmap[] of every STEP_GEMV phase (host side, before upload)
cudaError_t plan_decode_step(const StepPhase* host_phases, int nph, int R, int num_sms, StepLaunch* out);
cudaError_t launch_decode_step(const StepArgs& sa, const StepLaunch& L, int R, bool pdl, cudaStream_t stream);

// ---------------------------------------------------------------- tcgen05 GEMM  C[M,N] = A[M,K] * W[N,K]^T
enum GemmEpilogue {
  GEMM_OUT_F32 = 0,        // out_f32[row, col] = acc (+bias)
  GEMM_OUT_ADD_F32 = 1,    // out_f32[row, col] += acc (+bias)
  GEMM_OUT_BF16 = 2,       // out_bf16 = acc (+bias)
  GEMM_OUT_BF16_QGELU = 3, // quick_gelu(acc+bias)   (CLIP MLP)
  GEMM_OUT_BF16_GELU = 4,  // gelu_erf(acc+bias)     (mm_projector)
  GEMM_OUT_SWIGLU = 5,     // out_bf16[row, col/2] = silu(acc[2i]) * acc[2i+1]  (interleaved gate/up)
  GEMM_OUT_QKV_ROPE = 6    // prefill qkv: RoPE + bf16 q / KV-cache writes in the epilogue (GemmArgs::qkv_rope; SM-pair kernel only)
};
struct QkvRopeEpi {
  const int32_t* row_seq;  // [M] sequence of the row
  const int32_t* row_pos;  // [M] position of the row inside its sequence
  const int64_t* kv_off;   // [nseq] element offset of the sequence's block in a cache layer
  const int32_t* kv_cap;   // [nseq]
  const float* rope_cos;   // [max_pos, 64]
  const float* rope_sin;
  void* q_out;             // bf16 [M, Hl]
  void* kc;                // bf16 K cache of the layer: (seq, head h, token t) at kc + kv_off[seq] + (h*kv_cap[seq] + t)*128
  void* vc;
  int Hl;                  // heads_local * 128, a multiple of 256
};
struct GemmArgs {
  const void* a;  // bf16 [M, K] row-major (lda = K)
  const void* w;  // bf16 [N, K] row-major
  int M, N, K;
  int epilogue;
  const void* bias;        // bf16 [N] or null
  const int32_t* row_map;  // out row = row_map[row] (or row when null); <0 skips the row
  float* out_f32;
  void* out_bf16;
  int64_t ldc;
  float* split_ws = nullptr;  // optional [splits, M, ldc] fp32 scratch: lets a decode-sized ADD GEMM split K deterministically
  size_t split_ws_bytes = 0;
  bool pdl = false;           // launch with programmatic stream serialization (the kernel waits before touching dependent data)
  int* defer_reduce = nullptr;  // non-null: leave the split-K planes in split_ws and report their number here (0 = accumulated
                                // directly); the caller folds the reduction into its next pass (launch_reduce_rmsnorm)
  // Tensor parallel, GEMM_OUT_F32 only: the epilogue stores every output tile into ALL ranks' buffers over NVLink peer memory
  // (tp_out[k] = where this rank's tile lands on rank k: its partial plane of the all-reduce inbox, or its column block of the
  // gathered logits), then every CTA bumps the arrival counter tp_cnt[k] on every rank (system fence + red.sys).  The consumer
  // kernel on each rank waits for tp_n * (*tp_ctas) arrivals per decode step.  No NCCL call, no extra launch.
  int tp_n = 0;
  float* tp_out[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  unsigned* tp_cnt[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int* tp_ctas = nullptr;       // out: CTAs of the launch (arrivals this rank contributes per peer)
  // GEMM_OUT_QKV_ROPE (prefill qkv projection, N = 3*Hl, head_dim 128): the epilogue ropes q and k in fp32 and writes q bf16 to
  // [M, Hl], k / v bf16 into the head-major KV cache of the row's sequence -- no fp32 [M, 3*Hl] round trip, no separate pass
  const QkvRopeEpi* qkv_rope = nullptr;
};
cudaError_t launch_gemm(const GemmArgs& a, cudaStream_t stream);
// true when launch_gemm would run GEMM_OUT_QKV_ROPE for this shape (the SM-pair kernel); otherwise the caller keeps the fp32
// projection + launch_rope_kv_prefill
bool gemm_qkv_rope_fused(int M, int N, int K);
// host copy of the persistent kernels' tile walk (tile index -> m-tile, n-tile), for tests
void gemm_tile_coords_host(int tile, int m_tiles, int n_tiles, int* m_idx, int* n_idx);
const char* gemm_init_error();

// ---------------------------------------------------------------- attention
struct DecodeAttnArgs {
  const float* qkv;     // fp32 [R, 3*Hl] (q | k | v) of the new token, pre-RoPE
  void* kcache;         // bf16, this layer: row r, head h at kcache + kv_off[r] + h*cap[r]*D
  void* vcache;
  const int64_t* kv_off;   // [R] element offset of the row's block inside the layer
  const int32_t* kv_cap;   // [R] capacity (tokens) of the row
  const int32_t* pos;      // [R] position of the new token = number of cached tokens
  const float* rope_cos;   // [max_pos, D/2]
  const float* rope_sin;
  void* out_bf16;       // [R, Hl]
  float* part_o;        // workspace [R*heads*max_splits, D]
  float* part_ml;       // workspace [R*heads*max_splits, 2]
  int32_t* counters;    // [R*heads] zero on entry, zero on exit
  int R, heads, max_splits;
  float scale;
  // streaming variant (many rows): work table of (row, split) items with `chunk` tokens per split; null => 64-token kernel
  const int32_t* item_row = nullptr;
  const int32_t* item_split = nullptr;
  int n_items = 0, chunk = 0;
};
cudaError_t launch_decode_attn(const DecodeAttnArgs& a, cudaStream_t stream);

struct PrefillAttnArgs {
  const void* q;  // bf16; query i of sequence s, head h at q + (q_row0[s] + i)*q_ld + h*D
  int64_t q_ld;
  const void* k;  // bf16; key t of (s, h) at k + k_off[s] + h*k_head_stride[s] + t*k_tok_stride
  const void* v;
  const int64_t* k_off;          // [nseq]
  const int64_t* k_head_stride;  // [nseq]
  int64_t k_tok_stride;
  const int32_t* q_row0;  // [nseq]
  const int32_t* seqlen;  // [nseq]
  void* out;              // bf16, same addressing as q with out_ld
  int64_t out_ld;
  int nseq, heads, max_seqlen, causal, head_dim;
  float scale;
  // tcgen05 kernel only (0 => the mma.sync kernel): extents of the buffers behind q and k / v for the TMA tensor maps --
  // q is [q_rows, q_ld], k and v are [kv_rows, k_tok_stride] (k_off + h*k_head_stride must put a head's 64-column boxes inside a row)
  int64_t q_rows = 0, kv_rows = 0;
};
// tcgen05 + TMA kernel (prefill_attn_tc.cu) when the layout allows it, else / with VDD_PREFILL_ATTN_MMA=1 the mma.sync kernel
cudaError_t launch_prefill_attn(const PrefillAttnArgs& a, cudaStream_t stream);
cudaError_t launch_prefill_attn_mma(const PrefillAttnArgs& a, cudaStream_t stream);
cudaError_t launch_prefill_attn_tc(const PrefillAttnArgs& a, cudaStream_t stream);
bool prefill_attn_tc_supported(const PrefillAttnArgs& a);
// 2-D bf16 row-major tensor map [rows, cols] with row pitch `pitch_elems`, box {64, box_rows}, 128-byte swizzle (gemm_tc.cu)
bool make_tmap_bf16_2d(::CUtensorMap_st* m, const void* ptr, int64_t rows, int64_t cols, int64_t pitch_elems, int box_rows);

// step-0 attention map of one layer for the first B sequences (all of length len): out [B, heads, len, len] fp32
struct AttnProbsArgs {
  const void* q;   // bf16 roped queries, query i of sequence b at q + (q_row0[b] + i)*q_ld + h*128
  int64_t q_ld;
  const void* k;   // bf16 K cache of the layer: key t of (b, h) at k + k_off[b] + h*kv_cap[b]*128 + t*128
  const int64_t* k_off;
  const int32_t* kv_cap;
  const int32_t* q_row0;
  float* out;
  int B, heads, len;
  float scale;
};
cudaError_t launch_attn_probs(const AttnProbsArgs& a, cudaStream_t stream);

// attention row of one decode step and one layer for the first B rows (the main stream): out[step-1][layer][b][h][0..W), the first
// pos[b]+1 entries are softmax(q k^T / sqrt(d)) over the row's cache (the new key included), the rest zero
struct AttnProbsDecodeArgs {
  const float* qkv;   // fp32 [R, 3*heads*128] of the step (q | k | v), pre-RoPE
  const void* k;      // bf16 K cache of the layer, as in AttnProbsArgs
  const int64_t* k_off;
  const int32_t* kv_cap;
  const int32_t* pos;       // [R] position of the new token
  const float* rope_cos;    // [max_pos, 64]
  const float* rope_sin;
  const int32_t* step;      // device step counter (>= 1 during decode)
  float* out;
  int B, heads, layers, layer, W;
  float scale;
};
cudaError_t launch_attn_probs_decode(const AttnProbsDecodeArgs& a, cudaStream_t stream);

// ---------------------------------------------------------------- element-wise / glue
cudaError_t launch_rmsnorm_bf16(const float* x, const void* w_bf16, float eps, void* out_bf16, int rows, int H,
                                const int32_t* row_idx, cudaStream_t stream, bool pdl = false);
// resid[r] += sum_z ws[z][r] (split-K planes of a residual GEMM, plane pitch rows*H), then out[r] = bf16(rmsnorm(resid[r]) * w)
// wait != null: the planes are the tensor-parallel inbox; the kernel first waits until *wait >= (*epoch + 1) * wait_per_epoch
cudaError_t launch_reduce_rmsnorm(float* resid, const float* ws, int splits, int rows, int H, const void* w_bf16, float eps,
                                  void* out_bf16, cudaStream_t stream, bool pdl, const unsigned* wait = nullptr,
                                  unsigned wait_per_epoch = 0, const int32_t* epoch = nullptr);
// one thread spins until *counter >= (*epoch + 1) * per_epoch (all ranks' logit blocks have landed), then the stream goes on
cudaError_t launch_tp_wait(const unsigned* counter, unsigned per_epoch, const int32_t* epoch, cudaStream_t stream);
cudaError_t launch_layernorm(const float* x, const void* w_bf16, const void* b_bf16, float eps, void* out_bf16,
                             float* out_f32, int rows, int H, const int32_t* row_idx, cudaStream_t stream);
cudaError_t launch_embed_rows(const void* table_bf16, const int32_t* tok, const int32_t* img_row,
                              const float* img_feats, float* resid, int rows, int H, cudaStream_t stream);
cudaError_t launch_rope_kv_prefill(const float* qkv, int rows, int heads, int D, const int32_t* row_seq,
                                   const int32_t* row_pos, const int64_t* kv_off, const int32_t* kv_cap,
                                   const float* rope_cos, const float* rope_sin, void* q_out_bf16, void* kcache,
                                   void* vcache, cudaStream_t stream);
cudaError_t launch_im2col(const void* pixels_bf16, void* out_bf16, int B, int S, int P, int Kpad, cudaStream_t stream);
cudaError_t launch_vit_embed(const float* patch_out, const void* cls_bf16, const void* pos_bf16, float* resid, int B,
                             int NP, int H, cudaStream_t stream);
cudaError_t launch_decode_embed(const void* table_bf16, const int64_t* tokens, float* resid, int B, int S, int H,
                                cudaStream_t stream);
cudaError_t launch_advance(int32_t* pos, int R, int32_t* step, cudaStream_t stream, int32_t* epoch = nullptr);  // pos[r] += 1, *step += 1, *epoch += 1
cudaError_t launch_add_rows(float* dst, const float* src, int64_t n, cudaStream_t stream);
cudaError_t launch_reduce_splits(float* dst, const float* ws, int64_t n, int splits, cudaStream_t stream);  // dst += sum_z ws[z]
cudaError_t launch_cast_rows_bf16(const float* src, void* dst_bf16, int rows, int H, const int32_t* row_idx,
                                  cudaStream_t stream);

// ---------------------------------------------------------------- CLIP image preprocessing (preprocess.cu)
// uint8 RGB [H, W, 3] (host or device) -> [3, S, S] fp32 / bf16 / fp16 (out_dtype 0 / 1 / 2): optional square padding with pad_rgb,
// Pillow-exact bicubic shortest-edge resize, center crop, * 1/255, (x - mean) / std
cudaError_t launch_clip_preprocess(const uint8_t* rgb, int H, int W, bool on_device, int S, const uint8_t* pad_rgb, const float* mean,
                                   const float* stdv, void* out, int out_dtype, cudaStream_t st);

}  // namespace vdd
