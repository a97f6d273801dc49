/* libvdd -- C ABI of the B200-native Visual-Debias-Decoding (VDD) hot path.
 *
 * The reference (yfzhang114/LLaVA-Align) is pure Python and has no FFI of its own; its boundary for this path is
 *   - evolve_vcd_sampling()                              reference vcd_utils/vcd_sample.py:325-326
 *   - model.generate(input_ids, images=, use_dd=, ...)   reference experiments/eval/sampling/llava_sampling.py:96-109
 * The Python package llava_align_b200 keeps that surface and binds the entry points below with ctypes
 * (llava_align_b200/_lib.py).  INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions: every call returns 0 on success and a negative code on failure (text via vdd_last_error());
 * pointers are DEVICE pointers unless the name ends in _host; all work is enqueued on the caller's stream
 * (a cudaStream_t passed as void*), no hidden synchronisation unless stated; one engine is used from one
 * host thread; tensor-parallel runs use one engine per rank.  No torch types appear here.
 *
 * Threading and memory: an engine is NOT thread-safe.  Its work buffers grow on demand inside vdd_prefill* / vdd_generate*
 * (cudaFree + cudaMalloc of the buffer that is too small); the CUDA graph of the decode step is keyed on every pointer and shape
 * it bakes in and is re-captured after such a growth, so a caller never sees a stale graph -- but two host threads driving one
 * engine, or a second stream still running the engine's kernels while the next prefill re-allocates, are outside the contract.
 * The chained decode kernels synchronise their CTAs with a grid barrier that needs the whole grid resident (checked against the
 * occupancy calculator at launch); a co-tenant kernel that holds SMs for the whole launch makes the barrier's bounded spin trap
 * (error return), it does not hang.
 */
#ifndef VDD_H_
#define VDD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VDD_API __attribute__((visibility("default")))
#else
#define VDD_API
#endif

#define VDD_OK 0
#define VDD_ERR_ARG (-1)
#define VDD_ERR_CUDA (-2)
#define VDD_ERR_STATE (-3)
#define VDD_ERR_UNSUPPORTED (-4)

#define VDD_IMAGE_TOKEN_INDEX (-200) /* reference experiments/llava/constants.py:9 */

typedef struct vdd_engine vdd_engine;

/* Model dimensions: the HF config.json values of the checkpoints the reference loads
 * (reference experiments/llava/model/builder.py:97-103, multimodal_encoder/clip_encoder.py:22-27). */
typedef struct vdd_config {
  int32_t hidden_size, intermediate_size, num_layers, num_heads, vocab_size, max_position;
  float rms_eps, rope_theta;
  int32_t v_hidden, v_intermediate, v_layers, v_layers_used, v_heads, image_size, patch_size;
  float v_ln_eps;
  int32_t tp_size, tp_rank; /* tensor parallel: heads / FFN columns / vocab are split tp_size ways */
} vdd_config;

/* Sampler parameters = the generate() kwargs the patched sampler reads (reference vcd_sample.py:122-124,
 * 188-189) plus the HF warper settings (temperature/top_k/top_p). */
typedef struct vdd_sample_params {
  float cd_alpha;     /* reference default 0.5 (vcd_sample.py:188) */
  float log_cd_beta;  /* log(cd_beta) evaluated by the host in fp32, as torch.log(torch.tensor(beta)) does */
  float temperature;  /* 1.0 = off */
  float top_p;        /* >= 1.0 = off */
  int32_t top_k;      /* <= 0 = off */
  int32_t use_contrast; /* 0: plain sampling of l (no VDD stream present) */
  int64_t eos_token_id; /* < 0: no EOS handling (reference: eos_token_id=None) */
  int64_t pad_token_id;
} vdd_sample_params;

VDD_API const char* vdd_last_error(void);
VDD_API int vdd_version(void);

/* ---- K8: contrastive combine + warp + softmax + sample + EOS bookkeeping, one launch -------------------------
 * Replaces reference vcd_sample.py:185,188-202 and :257-260,285-288.
 *   l, u, d       fp32 [B, V] (row pitch ld elements); u and/or d may be NULL.  both present: cd=(u+d)/2.
 *   q_exp         fp32 [B, V] Exp(1) variates (torch.multinomial's own draw, see tests) or NULL => argmax.
 *   scores_out    fp32 [B, V] post-warper scores (the reference's `scores` tuple entry) or NULL.
 *   tokens_out    int64 [B];  unfinished int32 [B] in/out or NULL.                                              */
VDD_API int vdd_logit_combine_sample(const float* l, const float* u, const float* d, int B, int V, int64_t ld,
                             const vdd_sample_params* params_host, const float* q_exp, float* scores_out,
                             int64_t* tokens_out, int32_t* unfinished, void* stream);

/* Measurement aid (bench.py's logit-combine GB/s): `reps` launches of K8 on the same inputs captured in one CUDA graph,
 * replayed four times; *ms_per_launch_out = best replay / reps.  The stream must not be the legacy default stream.  q = NULL: the
 * variates are drawn inside the kernel (Philox, fixed seed), as vdd_generate* does when no q_exp is passed.                 */
VDD_API int vdd_bench_combine(const float* l, const float* u, const float* d, int B, int V, int64_t ld,
                      const vdd_sample_params* params_host, const float* q_exp, int64_t* tokens_out, int reps,
                      float* ms_per_launch_out, void* stream);

/* ---- engine ---------------------------------------------------------------------------------------------------*/
VDD_API int vdd_create(const vdd_config* cfg, vdd_engine** out);
VDD_API void vdd_destroy(vdd_engine* e);

/* Copy one checkpoint tensor (bf16, HF LLaVA-1.5 state-dict name, already sliced for this TP rank) into the
 * engine's packed weight arena.  Replaces from_pretrained at reference llava/model/builder.py:102-103. */
VDD_API int vdd_load_tensor(vdd_engine* e, const char* name, const void* src_bf16, const int64_t* shape, int ndim,
                    void* stream);
/* 0 when every tensor of the configured model has been loaded, else the number still missing. */
VDD_API int vdd_missing_tensors(vdd_engine* e, char* names_out, size_t cap);

/* ViT + projector (reference clip_encoder.py:39-51, multimodal_projector/builder.py:39-46):
 * pixels bf16 [B,3,S,S] -> engine-held image features [B, P, H] fp32.  */
VDD_API int vdd_encode_images(vdd_engine* e, const void* pixels_bf16, int B, void* stream);
VDD_API const float* vdd_image_features(vdd_engine* e);
/* The VCD baseline's noised images (`images_cd`, reference vcd_utils/vcd_add_noise.py:3-28, llava_llama.py:170):
 * encoded into a second feature plane; call after vdd_encode_images with the same B. */
VDD_API int vdd_encode_images_cd(vdd_engine* e, const void* pixels_bf16, int B, void* stream);

/* Prefill of all streams in one varlen batch (reference vcd_sample.py:106-119,147-185 at step 0 +
 * llava_arch.py:87-204).  ids_host int64 [B,T], each row holding at most one -200 sentinel.
 * Stream order in every [S,B,...] result: main, then unk (if use_unk), then none (if use_none).
 * with_images=0: text-only main stream (sentinel rows are an error).  Leaves last-position logits in
 * vdd_logits().  */
VDD_API int vdd_prefill(vdd_engine* e, const int64_t* ids_host, int B, int T, int with_images, int use_unk, int use_none,
                int max_new_tokens, void* stream);
/* Prefill of the main stream plus the VCD contrast stream (same prompt, images_cd features spliced in): step 0 of
 * reference vcd_sample.py:148-150,161-169.  Streams: main, cd. */
VDD_API int vdd_prefill_cd(vdd_engine* e, const int64_t* ids_host, int B, int T, int max_new_tokens, void* stream);
/* Arms the NEXT prefill (vdd_prefill / vdd_generate*) to also write the step-0 attention map of the LAST layer of the main
 * stream -- the `attentions[0][-1]` of generate(output_attentions=True) that the reference's post-hoc-debias drivers read
 * (llava_calibrate.py:180-182, vcd_sample.py:241-246) -- into out: fp32 [B, heads, len, len], len = prompt length after
 * the image splice, zeros above the diagonal.  elems must equal B*heads*len*len.  out = NULL disarms.  Not under TP. */
VDD_API int vdd_set_attention_capture(vdd_engine* e, float* out, int64_t elems);
/* The whole `attentions` tuple of generate(output_attentions=True) (vcd_sample.py:241-246: the MAIN stream's per-layer maps, one entry
 * per step), for the NEXT vdd_generate* only.  step0: fp32 [L, B, heads, len, len] (attentions[0][l]); later: fp32
 * [max_new_tokens - 1, L, B, heads, len + max_new_tokens - 1], row (s-1, l, b, h) = attentions[s][l][b, h, 0, :len + s] followed by
 * zeros.  Either may be NULL.  The decode steps of such a call run as plain launches (no CUDA graph).  Not under TP. */
VDD_API int vdd_set_attention_capture_all(vdd_engine* e, float* step0, int64_t step0_elems, float* later, int64_t later_elems);
/* One decode step for all S*B rows: row (s,b) consumes tokens[b].  Leaves logits in vdd_logits(). */
/* Sampling variates from the caller's torch CUDA generator state instead of an explicit q_exp: the NEXT vdd_generate* with
 * q_exp == NULL draws, inside K8, exactly the Exp(1) numbers torch.multinomial(probs, 1) (reference vcd_sample.py:202) would draw
 * on this device -- call `step` of `empty([B, V]).exponential_(1)` with the generator at (seed, offset + step * stride): Philox4x32-10
 * as cuRAND runs it for ATen, ATen's element -> (thread, call, component) mapping, -log(u).  After the call the caller advances
 * its generator by vdd_sampler_rng_advance() (= steps * stride) to stay in step with the reference.  offset % 4 == 0.            */
VDD_API int vdd_set_sampler_rng(vdd_engine* e, uint64_t seed, uint64_t offset);
VDD_API uint64_t vdd_sampler_rng_advance(vdd_engine* e);
/* Test aid: out_dev[i] = element i of `torch.empty(numel, device='cuda').exponential_(1)` under (seed, offset). */
VDD_API int vdd_debug_philox_exponential(uint64_t seed, uint64_t offset, int64_t numel, float* out_dev, void* stream);
/* Test aid (host only, no GPU needed): the persistent GEMM kernels' grouped tile walk -- which (m-tile, n-tile) the tile-th unit of
 * work is.  A bijection onto [0, m_tiles) x [0, n_tiles); consecutive tiles share operand panels (gemm_tc.cu, gemm_tile_coords). */
VDD_API int vdd_debug_gemm_tile_coords(int tile, int m_tiles, int n_tiles, int* m_idx_host, int* n_idx_host);
/* Test aid: one launch of the varlen prefill attention (llava_llama.py:91-100 at step 0, clip_encoder.py:48) on caller-provided device
 * buffers.  q is bf16 [q_rows, q_ld]; k / v are bf16 [kv_rows, k_tok_stride]; key t of (sequence s, head h) starts at element
 * k_off[s] + h*k_head_stride[s] + t*k_tok_stride; out is bf16 addressed like q with out_ld.  use_tc: 1 = tcgen05 + TMA kernel, 0 = the
 * mma.sync kernel. */
VDD_API int vdd_debug_prefill_attn(const void* q, int64_t q_ld, int64_t q_rows, const void* k, const void* v, int64_t kv_rows,
                                   int64_t k_tok_stride, const int64_t* k_off, const int64_t* k_head_stride, const int32_t* q_row0,
                                   const int32_t* seqlen, void* out, int64_t out_ld, int nseq, int heads, int max_seqlen, int causal,
                                   int head_dim, float scale, int use_tc, void* stream);
/* Ragged batches: mask_host [B, T] bytes, 1 = token, 0 = padding (left or right), applies to the NEXT prefill / generate only (NULL
 * disarms).  Masked positions do not exist for any stream: a padded row computes exactly what the same prompt computes alone.  The
 * reference has no working equivalent (its decode step rebuilds the mask as all ones, llava_arch.py:91-94). */
VDD_API int vdd_set_prompt_mask(vdd_engine* e, const uint8_t* mask_host, int B, int T);
/* Arm the capture of the debias streams' own first-step scores for the NEXT vdd_generate / vdd_generate_cd: out receives
 * [streams - 1][B][V] fp32 post-warper scores (same temperature / top-k / top-p, no contrast) of the unk / none / cd streams in
 * stream order -- what the reference's post-hoc-debias drivers obtain from two extra generates per question
 * (experiments/eval/calibrate/llava_calibrate.py:192-194 -> :62-82, `scores[0]`).  elems must equal (streams - 1) * B * V. */
VDD_API int vdd_set_stream_scores_capture(vdd_engine* e, float* out_dev, int64_t elems);
VDD_API int vdd_decode_step(vdd_engine* e, const int64_t* tokens, void* stream);
VDD_API float* vdd_logits(vdd_engine* e); /* fp32 [S, B, V] */
VDD_API int vdd_num_streams(vdd_engine* e);

/* Whole generate() with HOST buffers: encode + prefill + graph-replayed decode steps with the combine/sample
 * kernel inside the graph.  pixels_host bf16 [B,3,S,S] (NULL => text only); q_exp [steps,B,V] device or NULL;
 * scores_out [steps,B,V] device or NULL; tokens_out_host int64 [B, max_new_tokens]; returns the number of steps
 * taken in *steps_out (stops early when every row has emitted EOS, reference vcd_sample.py:285-292).
 * Synchronises the stream before returning.  */
VDD_API int vdd_generate(vdd_engine* e, const int64_t* ids_host, int B, int T, const void* pixels_host, int use_unk,
                 int use_none, const vdd_sample_params* params_host, int max_new_tokens, const float* q_exp,
                 float* scores_out, int64_t* tokens_out_host, int* steps_out, void* stream);

/* Same with the VCD contrast stream: pixels_cd_host bf16 [B,3,S,S] (NULL => identical to vdd_generate).  As in the
 * reference (vcd_sample.py:148-150) only step 0 contrasts against the noised image; from step 1
 * the contrast logits are the main logits.  With use_unk / use_none the reference's precedence applies
 * (vcd_sample.py:148,171-185): the cd forward replaces the unk stream, and the 'none' stream joins (cd = (cd + none)/2)
 * only when BOTH flags are set. */
VDD_API int vdd_generate_cd(vdd_engine* e, const int64_t* ids_host, int B, int T, const void* pixels_host,
                    const void* pixels_cd_host, int use_unk, int use_none, const vdd_sample_params* params_host,
                    int max_new_tokens, const float* q_exp, float* scores_out, int64_t* tokens_out_host, int* steps_out,
                    void* stream);

/* Tensor parallel: 128-byte NCCL unique id made on rank 0, then every rank joins. */
VDD_API int vdd_tp_unique_id(void* id128_host);
VDD_API int vdd_tp_init(vdd_engine* e, const void* id128_host);
/* In-kernel all-reduce over NVLink peer memory (decode with <= 8 fused rows): every rank exports its 64-byte cudaIpc
 * handle, the host gathers them ([tp][64], rank order) and every rank maps its peers.  Without these two calls tensor
 * parallel decode falls back to NCCL all-reduce launches. */
VDD_API int vdd_tp_comm_handle(vdd_engine* e, void* handle64_host);
VDD_API int vdd_tp_comm_open(vdd_engine* e, const void* handles_host);

/* Introspection for bench.py: kernels launched since the last reset; per-phase device times of the last generate. */
VDD_API int64_t vdd_launch_count(vdd_engine* e, int reset);
VDD_API int vdd_last_timings(vdd_engine* e, float* vision_ms, float* prefill_ms, float* decode_ms);

/* Debug (VDD_TRACE=1): globaltimer stamps [148 CTAs][4 phases][6 events] of the middle layer's fused projection chain. */
VDD_API int vdd_debug_trace(vdd_engine* e, unsigned long long* out_host, int n);
/* Enqueue one decode step's projection kernels only (4 projections per layer + lm_head, with the norm passes of the
 * many-row path), no attention / embedding / sampling: bench.py's roofline leg. */
VDD_API int vdd_bench_projection_chain(vdd_engine* e, void* stream);

/* Stand-alone kernel entry points used by the unit tests and micro-benchmarks (same kernels the engine runs). */
VDD_API int vdd_gemv(const void* w_bf16, const void* x_bf16, const float* x_f32_resid, const void* norm_w_bf16, float eps,
             int R, int N, int K, int epilogue, float* out_f32, void* out_bf16, void* stream);
VDD_API int vdd_gemm(const void* a_bf16, const void* w_bf16, int M, int N, int K, int epilogue, const void* bias_bf16,
             const int32_t* row_map, float* out_f32, void* out_bf16, int64_t ldc, void* stream);

/* ---- host-side pipeline of the callers on the GPU -----------------------------------------------
 * CLIP preprocessing of ONE image, replacing `image_processor.preprocess(image, return_tensors='pt')['pixel_values'][0]`
 * (reference experiments/eval/sampling/llava_sampling.py:82, eval/calibrate/llava_calibrate.py:149; transformers-4.31
 * CLIPImageProcessor + Pillow BICUBIC) and, with pad_rgb != NULL, the expand2square padding of reference
 * experiments/llava/mm_utils.py:152-164 (pad_rgb = tuple(int(x * 255) for x in image_mean)).
 *   rgb        uint8 [H, W, 3], host memory (copied inside) or device memory (rgb_on_device != 0)
 *   out_dev    device [3, size, size]; out_dtype 0 = fp32 (bit-exact with the CPU pipeline), 1 = bf16, 2 = fp16            */
VDD_API int vdd_clip_preprocess(const uint8_t* rgb, int H, int W, int rgb_on_device, int size, const uint8_t* pad_rgb,
                        const float* mean3, const float* std3, void* out_dev, int out_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VDD_H_ */
