// libvdd engine: weight arena, per-generation row/stream plan, prefill / decode orchestration, CUDA-graph replay of
// the decode step with the combine+sample kernel inside, and the C ABI of include/vdd.h.
//
// What it stands in for in the reference (all Python there):
//   sample() loop + per-stream model_kwargs bookkeeping     vcd_utils/vcd_sample.py:93-299
//   LlavaLlamaForCausalLM.forward / prepare_inputs_*         experiments/llava/model/language_model/llava_llama.py:58-174
//   prepare_inputs_labels_for_multimodal (image splice)      experiments/llava/model/llava_arch.py:87-204
//   CLIPVisionTower.forward + mm_projector                   multimodal_encoder/clip_encoder.py:39-51, multimodal_projector/builder.py:39-46
// The reference runs S = 1 + {1,2} separate forwards per token with separate KV caches; here every stream of every
// sequence is a ROW of one fused batch (row = stream*B + b), so each weight byte is read once per token.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/vdd.h"
#include "kernels.h"

using namespace vdd;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t _e = (call);                                                                              \
    if (_e != cudaSuccess) return fail(VDD_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
  } while (0)

extern "C" const char* vdd_last_error(void) { return g_err; }
extern "C" int vdd_version(void) { return 100; }

// ------------------------------------------------------------------------------------------------ NCCL (loaded lazily; TP only)
typedef struct ncclComm* ncclComm_t;
struct NcclId { char b[128]; };  // ncclUniqueId, passed by value
struct NcclApi {
  void* lib = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static bool nccl_load() {
  if (g_nccl.lib) return true;
  const char* names[] = {getenv("VDD_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return false;
  *(void**)&g_nccl.GetUniqueId = dlsym(g_nccl.lib, "ncclGetUniqueId");
  *(void**)&g_nccl.CommInitRank = dlsym(g_nccl.lib, "ncclCommInitRank");
  *(void**)&g_nccl.AllReduce = dlsym(g_nccl.lib, "ncclAllReduce");
  *(void**)&g_nccl.AllGather = dlsym(g_nccl.lib, "ncclAllGather");
  *(void**)&g_nccl.CommDestroy = dlsym(g_nccl.lib, "ncclCommDestroy");
  *(void**)&g_nccl.GetErrorString = dlsym(g_nccl.lib, "ncclGetErrorString");
  return g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.AllGather;
}

// ------------------------------------------------------------------------------------------------ engine state
struct Slot {
  void* ptr = nullptr;
  int64_t rows = 0, cols = 0;     // expected source shape (flattened 2D)
  int64_t dst_pitch = 0;          // destination row pitch in elements (>= cols)
  bool loaded = false, ignore = false;
};
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};
struct LayerW {
  void *wqkv, *wo, *wgu, *wdown, *ln1, *ln2;
};
struct VLayerW {
  void *ln1w, *ln1b, *wqkv, *bqkv, *wo, *bo, *ln2w, *ln2b, *fc1, *b1, *fc2, *b2;
};

struct vdd_engine {
  vdd_config c;
  int D = 128, heads_l = 0, Hl = 0, Il = 0, Vl = 0, NP = 0, Kpad = 0, VD = 64;
  int num_sms = 148;
  bool use_pdl = true, use_graph = true, use_fused = true;
  // weights
  void* arena = nullptr;
  size_t arena_bytes = 0;
  std::map<std::string, Slot> slots;
  std::map<std::string, int> parts;  // arrivals per fused slot
  std::vector<LayerW> lw;
  std::vector<VLayerW> vw;
  void *embed, *final_norm, *lm_head;
  void *patch_w, *cls, *vpos, *preln_w, *preln_b, *pj_w1, *pj_b1, *pj_w2, *pj_b2;
  float *rope_cos = nullptr, *rope_sin = nullptr;
  // workspaces
  DevBuf resid, xn, qkv, qb, attn, hbuf, meta, kv, logits, logits_tp, dec_f32, dec_bf16, part, img_feats, vis, pix, misc, tp_tmp, split_ws;
  // per-generation plan
  int B = 0, T = 0, S = 0, R = 0, Mtot = 0, max_new = 0, max_len = 0, max_splits = 1;
  int attn_items = 0, attn_chunk = 0;  // streaming decode attention (R > 8): (row, split) work table
  int32_t *d_item_row = nullptr, *d_item_split = nullptr;
  bool rng_armed = false;            // armed by vdd_set_sampler_rng: the next generate draws its variates in the kernel
  unsigned long long rng_seed = 0, rng_offset = 0, rng_advance = 0;
  int max_threads_per_sm = 2048;
  std::vector<uint8_t> prompt_mask;  // armed by vdd_set_prompt_mask: [B*T] 1 = real token, 0 = padding (next prefill only)
  int prompt_mask_B = 0, prompt_mask_T = 0;
  float* stream_scores = nullptr;  // armed by vdd_set_stream_scores_capture: step-0 post-warper scores of the debias streams
  int64_t stream_scores_elems = 0;
  float* attn_capture = nullptr;  // armed by vdd_set_attention_capture: the next prefill writes the last layer's step-0 map here
  int64_t attn_capture_elems = 0;
  // armed by vdd_set_attention_capture_all for the next generate: every layer's step-0 map and every later step's rows
  float* attn_all0 = nullptr;   // [L, B, heads, len, len]
  float* attn_alln = nullptr;   // [max_new - 1, L, B, heads, len + max_new - 1]
  int64_t attn_all0_elems = 0, attn_alln_elems = 0;
  int attn_all_w = 0;           // row width of attn_alln, set by the generate that consumes it
  int main_len = 0;  // length of the main-stream rows after the image splice (0: rows differ)
  std::vector<int32_t> h_item_row;
  int64_t kv_layer_stride = 0;
  int32_t *d_tok = nullptr, *d_img_row = nullptr, *d_row_seq = nullptr, *d_row_pos = nullptr, *d_kv_cap = nullptr,
          *d_pos = nullptr, *d_last_rows = nullptr, *d_seqlen = nullptr, *d_q_row0 = nullptr, *d_step = nullptr,
          *d_unfinished = nullptr, *d_counters = nullptr;
  int64_t *d_kv_off = nullptr, *d_head_stride = nullptr, *d_tokens = nullptr, *d_seq_out = nullptr;
  SampleArgs* d_sargs = nullptr;
  unsigned long long* d_trace = nullptr;  // VDD_TRACE=1: time stamps of the middle layer's fused chain
  unsigned* d_bar = nullptr;  // grid-barrier counters of the fused projection chains, [layers + 2][4]
  bool prefilled = false;
  int image_B = 0, image_cd_B = 0;
  // decode buffers (views into dec_f32 / dec_bf16)
  float *resid_d = nullptr, *qkv_d = nullptr, *part_o = nullptr, *part_ml = nullptr, *tp_part = nullptr;
  void *attn_d = nullptr, *h_d = nullptr, *xn_d = nullptr;
  // whole-step persistent kernel (decode_step.cu): device table of phases for the current plan
  bool use_step = true, step_ready = false;
  int step_R = 0;
  DevBuf step_tab;
  StepLaunch step_launch{};
  // graph cache
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_key = 0;
  int graph_nodes = 0;
  // TP
  ncclComm_t comm = nullptr;
  void* comm_buf = nullptr;        // this rank's peer-visible communication buffer (see TpComm in kernels.h)
  size_t comm_bytes = 0;
  TpComm tpc;                      // filled by vdd_tp_comm_open
  bool tp_fused_ready = false;     // peers mapped: decode uses the in-kernel all-reduce instead of NCCL
  int tp_max_rows = 0;             // rows the peer-visible inbox / logits areas were sized for (many-row decode path)
  int32_t* d_epoch = nullptr;      // decode steps since the last prefill (targets of the monotonic arrival counters)
  float* logits_cur = nullptr;     // where the last prefill / decode step left its logits
  // stats
  int64_t launches = 0;
  float t_vision = 0, t_prefill = 0, t_decode = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  void* pinned = nullptr;  // small pinned scratch
  size_t pinned_bytes = 0;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int ensure(vdd_engine* e, DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes) return VDD_OK;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr;
  b.bytes = 0;
  bytes = align_up(bytes + bytes / 8, 1 << 20);
  CK(cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  if (e->graph_exec) {  // captured pointers are stale now
    cudaGraphExecDestroy(e->graph_exec);
    e->graph_exec = nullptr;
  }
  return VDD_OK;
}

// ------------------------------------------------------------------------------------------------ create
static void add_slot(vdd_engine* e, std::vector<std::pair<std::string, size_t>>& order, const std::string& name,
                     int64_t rows, int64_t cols, int64_t pitch = 0) {
  Slot s;
  s.rows = rows;
  s.cols = cols;
  s.dst_pitch = pitch ? pitch : cols;
  e->slots[name] = s;
  order.push_back({name, (size_t)rows * s.dst_pitch * 2});
}

extern "C" int vdd_create(const vdd_config* cfg, vdd_engine** out) {
  if (!cfg || !out) return fail(VDD_ERR_ARG, "null argument");
  int dev_count = 0;
  cudaError_t ce = cudaGetDeviceCount(&dev_count);
  if (ce != cudaSuccess || dev_count == 0)
    return fail(VDD_ERR_CUDA, "libvdd needs a CUDA device (sm_100a); none visible: %s", cudaGetErrorString(ce));
  const vdd_config& c = *cfg;
  if (c.num_heads <= 0 || c.hidden_size % c.num_heads || c.hidden_size / c.num_heads != 128)
    return fail(VDD_ERR_UNSUPPORTED, "head_dim must be 128 (LLaMA-family LLaVA-1.5)");
  if (c.v_hidden / c.v_heads != 64) return fail(VDD_ERR_UNSUPPORTED, "vision head_dim must be 64 (CLIP ViT-L/14)");
  const int tp = c.tp_size < 1 ? 1 : c.tp_size;
  if (c.num_heads % tp || c.intermediate_size % tp || c.vocab_size % tp)
    return fail(VDD_ERR_ARG, "heads/intermediate/vocab must divide by tp_size");
  if ((c.intermediate_size / tp) % 8 || (c.vocab_size / tp) % 16 || c.hidden_size % 64)
    return fail(VDD_ERR_UNSUPPORTED, "local dims must be multiples of 8/16/64");
  vdd_engine* e = new vdd_engine();
  e->c = c;
  e->c.tp_size = tp;
  e->heads_l = c.num_heads / tp;
  e->Hl = e->heads_l * 128;
  e->Il = c.intermediate_size / tp;
  e->Vl = c.vocab_size / tp;
  e->NP = (c.image_size / c.patch_size) * (c.image_size / c.patch_size);
  e->Kpad = (int)align_up((size_t)3 * c.patch_size * c.patch_size, 64);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&e->max_threads_per_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
  e->use_pdl = getenv("VDD_NO_PDL") == nullptr;
  e->use_graph = getenv("VDD_NO_GRAPH") == nullptr;
  e->use_fused = getenv("VDD_NO_FUSED") == nullptr;
  // Whole-step persistent kernel (decode_step.cu) vs one gemv_kernel chain + one attention launch per layer (decode_gemv.cu):
  // measured (profiles/r2_step_kernel.md): the single launch streams the weights at HBM speed inside a projection, but its phase
  // boundaries (8 warps of 255 registers for prologue / epilogue / attention) cost more than the launches they replace --
  // 100 us per layer against 90 us on one GPU, 930 ms against 666 ms per 256-token generate at TP=2.  Opt-in: VDD_STEP=1.
  e->use_step = getenv("VDD_STEP") ? atoi(getenv("VDD_STEP")) != 0 : false;
  if (getenv("VDD_TRACE")) {
    cudaMalloc(&e->d_trace, (size_t)256 * 64 * 8 * 8);
    cudaMemset(e->d_trace, 0, (size_t)256 * 64 * 8 * 8);
  }

  const int64_t H = c.hidden_size, VH = c.v_hidden, VI = c.v_intermediate;
  std::vector<std::pair<std::string, size_t>> order;
  add_slot(e, order, "model.embed_tokens.weight", c.vocab_size, H);
  add_slot(e, order, "model.norm.weight", 1, H);
  add_slot(e, order, "lm_head.weight", e->Vl, H);
  for (int i = 0; i < c.num_layers; ++i) {
    const std::string p = "model.layers." + std::to_string(i) + ".";
    add_slot(e, order, p + "qkv", 3 * e->Hl, H);
    add_slot(e, order, p + "self_attn.o_proj.weight", H, e->Hl);
    add_slot(e, order, p + "gate_up", 2 * e->Il, H);
    add_slot(e, order, p + "mlp.down_proj.weight", H, e->Il);
    add_slot(e, order, p + "input_layernorm.weight", 1, H);
    add_slot(e, order, p + "post_attention_layernorm.weight", 1, H);
  }
  add_slot(e, order, "model.mm_projector.0.weight", H, VH);
  add_slot(e, order, "model.mm_projector.0.bias", 1, H);
  add_slot(e, order, "model.mm_projector.2.weight", H, H);
  add_slot(e, order, "model.mm_projector.2.bias", 1, H);
  const std::string vt = "model.vision_tower.vision_tower.vision_model.";
  add_slot(e, order, vt + "embeddings.class_embedding", 1, VH);
  add_slot(e, order, vt + "embeddings.patch_embedding.weight", VH, 3 * c.patch_size * c.patch_size, e->Kpad);
  add_slot(e, order, vt + "embeddings.position_embedding.weight", e->NP + 1, VH);
  add_slot(e, order, vt + "pre_layrnorm.weight", 1, VH);
  add_slot(e, order, vt + "pre_layrnorm.bias", 1, VH);
  for (int i = 0; i < c.v_layers_used; ++i) {
    const std::string p = vt + "encoder.layers." + std::to_string(i) + ".";
    add_slot(e, order, p + "qkv.weight", 3 * VH, VH);
    add_slot(e, order, p + "qkv.bias", 1, 3 * VH);
    add_slot(e, order, p + "self_attn.out_proj.weight", VH, VH);
    add_slot(e, order, p + "self_attn.out_proj.bias", 1, VH);
    add_slot(e, order, p + "layer_norm1.weight", 1, VH);
    add_slot(e, order, p + "layer_norm1.bias", 1, VH);
    add_slot(e, order, p + "layer_norm2.weight", 1, VH);
    add_slot(e, order, p + "layer_norm2.bias", 1, VH);
    add_slot(e, order, p + "mlp.fc1.weight", VI, VH);
    add_slot(e, order, p + "mlp.fc1.bias", 1, VI);
    add_slot(e, order, p + "mlp.fc2.weight", VH, VI);
    add_slot(e, order, p + "mlp.fc2.bias", 1, VH);
  }
  size_t total = 0;
  for (auto& o : order) total += align_up(o.second, 256);
  cudaError_t me = cudaMalloc(&e->arena, total);
  if (me != cudaSuccess) {
    delete e;
    return fail(VDD_ERR_CUDA, "cudaMalloc(%zu bytes of weights): %s", total, cudaGetErrorString(me));
  }
  cudaMemset(e->arena, 0, total);
  e->arena_bytes = total;
  size_t off = 0;
  for (auto& o : order) {
    e->slots[o.first].ptr = (char*)e->arena + off;
    off += align_up(o.second, 256);
  }
  auto P = [&](const std::string& n) { return e->slots[n].ptr; };
  e->embed = P("model.embed_tokens.weight");
  e->final_norm = P("model.norm.weight");
  e->lm_head = P("lm_head.weight");
  e->lw.resize(c.num_layers);
  for (int i = 0; i < c.num_layers; ++i) {
    const std::string p = "model.layers." + std::to_string(i) + ".";
    e->lw[i] = {P(p + "qkv"), P(p + "self_attn.o_proj.weight"), P(p + "gate_up"), P(p + "mlp.down_proj.weight"),
                P(p + "input_layernorm.weight"), P(p + "post_attention_layernorm.weight")};
  }
  e->pj_w1 = P("model.mm_projector.0.weight");
  e->pj_b1 = P("model.mm_projector.0.bias");
  e->pj_w2 = P("model.mm_projector.2.weight");
  e->pj_b2 = P("model.mm_projector.2.bias");
  e->cls = P(vt + "embeddings.class_embedding");
  e->patch_w = P(vt + "embeddings.patch_embedding.weight");
  e->vpos = P(vt + "embeddings.position_embedding.weight");
  e->preln_w = P(vt + "pre_layrnorm.weight");
  e->preln_b = P(vt + "pre_layrnorm.bias");
  e->vw.resize(c.v_layers_used);
  for (int i = 0; i < c.v_layers_used; ++i) {
    const std::string p = vt + "encoder.layers." + std::to_string(i) + ".";
    e->vw[i] = {P(p + "layer_norm1.weight"), P(p + "layer_norm1.bias"), P(p + "qkv.weight"), P(p + "qkv.bias"),
                P(p + "self_attn.out_proj.weight"), P(p + "self_attn.out_proj.bias"), P(p + "layer_norm2.weight"),
                P(p + "layer_norm2.bias"), P(p + "mlp.fc1.weight"), P(p + "mlp.fc1.bias"), P(p + "mlp.fc2.weight"),
                P(p + "mlp.fc2.bias")};
  }
  // RoPE table, HF LlamaRotaryEmbedding arithmetic: inv_freq fp32, angle = pos * inv_freq in fp32, cos/sin of it
  {
    const int half = 64, mp = c.max_position;
    std::vector<float> hc((size_t)mp * half), hs((size_t)mp * half);
    for (int i = 0; i < half; ++i) {
      const float inv = 1.0f / powf(c.rope_theta, (float)(2 * i) / 128.0f);
      for (int p = 0; p < mp; ++p) {
        const float ang = (float)p * inv;
        hc[(size_t)p * half + i] = cosf(ang);
        hs[(size_t)p * half + i] = sinf(ang);
      }
    }
    cudaMalloc(&e->rope_cos, hc.size() * 4);
    cudaMalloc(&e->rope_sin, hs.size() * 4);
    cudaMemcpy(e->rope_cos, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(e->rope_sin, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice);
  }
  for (int i = 0; i < 4; ++i) cudaEventCreate(&e->ev[i]);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) {
    delete e;
    return fail(VDD_ERR_CUDA, "vdd_create: %s", cudaGetErrorString(le));
  }
  *out = e;
  return VDD_OK;
}

extern "C" void vdd_destroy(vdd_engine* e) {
  if (!e) return;
  cudaDeviceSynchronize();
  if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
  DevBuf* bufs[] = {&e->resid, &e->xn, &e->qkv, &e->qb, &e->attn, &e->hbuf, &e->meta, &e->kv, &e->logits, &e->logits_tp,
                    &e->dec_f32, &e->dec_bf16, &e->part, &e->img_feats, &e->vis, &e->pix, &e->misc, &e->tp_tmp, &e->split_ws, &e->step_tab};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (e->arena) cudaFree(e->arena);
  if (e->rope_cos) cudaFree(e->rope_cos);
  if (e->rope_sin) cudaFree(e->rope_sin);
  if (e->pinned) cudaFreeHost(e->pinned);
  for (int i = 0; i < 4; ++i)
    if (e->ev[i]) cudaEventDestroy(e->ev[i]);
  if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
  for (int k = 0; k < 8; ++k)
    if (e->tp_fused_ready && k != e->c.tp_rank && e->tpc.peers[k]) cudaIpcCloseMemHandle((void*)e->tpc.peers[k]);
  if (e->comm_buf) cudaFree(e->comm_buf);
  if (e->d_epoch) cudaFree(e->d_epoch);
  delete e;
}

// ------------------------------------------------------------------------------------------------ weights
static bool ends_with(const std::string& s, const char* suf) {
  const size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

extern "C" int vdd_load_tensor(vdd_engine* e, const char* name_c, const void* src, const int64_t* shape, int ndim,
                               void* stream_v) {
  if (!e || !name_c || !src || !shape || ndim < 1) return fail(VDD_ERR_ARG, "null argument");
  cudaStream_t st = (cudaStream_t)stream_v;
  std::string name(name_c);
  int64_t rows = ndim == 1 ? 1 : shape[0], cols = 1;
  for (int i = (ndim == 1 ? 0 : 1); i < ndim; ++i) cols *= shape[i];
  // fused / repacked destinations
  std::string key = name;
  int64_t row_off = 0, row_step = 1, col_off = 0, exp_rows = -1, exp_cols = -1;
  int parts_needed = 0;
  const std::string vt = "model.vision_tower.vision_tower.vision_model.";
  if (name.compare(0, 13, "model.layers.") == 0) {
    const std::string pre = name.substr(0, name.find('.', 13) + 1);
    const char* qkvn[3] = {"self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight"};
    for (int i = 0; i < 3; ++i)
      if (ends_with(name, qkvn[i])) {
        key = pre + "qkv"; row_off = (int64_t)i * e->Hl; exp_rows = e->Hl; exp_cols = e->c.hidden_size; parts_needed = 3;
      }
    // gate and up rows are interleaved (2j = gate_j, 2j+1 = up_j) so SwiGLU is an epilogue of one projection
    if (ends_with(name, "mlp.gate_proj.weight")) { key = pre + "gate_up"; row_off = 0; row_step = 2; exp_rows = e->Il; exp_cols = e->c.hidden_size; parts_needed = 2; }
    if (ends_with(name, "mlp.up_proj.weight")) { key = pre + "gate_up"; row_off = 1; row_step = 2; exp_rows = e->Il; exp_cols = e->c.hidden_size; parts_needed = 2; }
  } else if (name.compare(0, vt.size(), vt) == 0 && name.find("encoder.layers.") != std::string::npos) {
    const size_t lp = name.find("encoder.layers.") + 15;
    const int li = atoi(name.c_str() + lp);
    if (li >= e->c.v_layers_used) return VDD_OK;  // layer never contributes (select_layer=-2, clip_encoder.py:30)
    const std::string pre = name.substr(0, name.find('.', lp) + 1);
    const char* n3[3] = {"self_attn.q_proj.", "self_attn.k_proj.", "self_attn.v_proj."};
    for (int i = 0; i < 3; ++i) {
      if (name.find(n3[i]) != std::string::npos) {
        parts_needed = 3;
        if (ends_with(name, "weight")) { key = pre + "qkv.weight"; row_off = (int64_t)i * e->c.v_hidden; exp_rows = e->c.v_hidden; exp_cols = e->c.v_hidden; }
        else { key = pre + "qkv.bias"; col_off = (int64_t)i * e->c.v_hidden; exp_rows = 1; exp_cols = e->c.v_hidden; }
      }
    }
  } else if (name.compare(0, vt.size(), vt) == 0 && name.find("post_layernorm") != std::string::npos) {
    return VDD_OK;  // unused by feature_select (clip_encoder.py:29-37)
  }
  auto it = e->slots.find(key);
  if (it == e->slots.end()) return fail(VDD_ERR_ARG, "unknown tensor '%s'", name_c);
  Slot& s = it->second;
  if (exp_rows < 0) { exp_rows = s.rows; exp_cols = s.cols; }
  if (rows * cols != exp_rows * exp_cols)
    return fail(VDD_ERR_ARG, "tensor '%s': got %lld elements, expected %lld x %lld", name_c, (long long)(rows * cols),
                (long long)exp_rows, (long long)exp_cols);
  char* dst = (char*)s.ptr + (row_off * s.dst_pitch + col_off) * 2;
  CK(cudaMemcpy2DAsync(dst, (size_t)s.dst_pitch * row_step * 2, src, (size_t)exp_cols * 2, (size_t)exp_cols * 2, (size_t)exp_rows,
                       cudaMemcpyDeviceToDevice, st));
  if (parts_needed == 0) s.loaded = true;
  else if (++e->parts[key] >= parts_needed) s.loaded = true;
  return VDD_OK;
}

extern "C" int vdd_missing_tensors(vdd_engine* e, char* names_out, size_t cap) {
  int n = 0;
  std::string acc;
  for (auto& kv : e->slots)
    if (!kv.second.loaded) {
      ++n;
      if (acc.size() < 900) acc += kv.first + ";";
    }
  if (names_out && cap) snprintf(names_out, cap, "%s", acc.c_str());
  return n;
}

// ------------------------------------------------------------------------------------------------ launch helpers
#define LK(e, call)                                                                                        \
  do {                                                                                                     \
    cudaError_t _e = (call);                                                                               \
    (e)->launches++;                                                                                       \
    if (_e != cudaSuccess) return fail(VDD_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
  } while (0)

static int tp_allreduce_add(vdd_engine* e, float* partial, float* resid, int64_t n, cudaStream_t st) {
  // row-parallel projection: sum the per-rank partials over NVLink, then resid += sum
  int r = g_nccl.AllReduce(partial, partial, (size_t)n, /*ncclFloat*/ 7, /*ncclSum*/ 0, e->comm, st);
  if (r != 0) return fail(VDD_ERR_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  e->launches++;
  LK(e, launch_add_rows(resid, partial, n, st));
  return VDD_OK;
}

static int do_gemm(vdd_engine* e, const void* a, const void* w, int M, int N, int K, int mode, const void* bias,
                   const int32_t* row_map, float* of32, void* obf16, int64_t ldc, cudaStream_t st, bool pdl = false,
                   int* defer_reduce = nullptr) {
  GemmArgs g{a, w, M, N, K, mode, bias, row_map, of32, obf16, ldc};
  g.pdl = pdl;
  g.defer_reduce = defer_reduce;
  if (mode == GEMM_OUT_ADD_F32 && M <= 128 && e->split_ws.p) {  // decode with > 8 fused rows: deterministic split-K
    g.split_ws = (float*)e->split_ws.p;
    g.split_ws_bytes = e->split_ws.bytes;
    if (!defer_reduce) e->launches++;
  }
  cudaError_t ce = launch_gemm(g, st);
  e->launches++;
  if (ce != cudaSuccess)
    return fail(VDD_ERR_CUDA, "gemm M=%d N=%d K=%d: %s %s", M, N, K, cudaGetErrorString(ce), gemm_init_error() ? gemm_init_error() : "");
  return VDD_OK;
}
static bool tp_push_pays(int R, int tp) {
  if (getenv("VDD_TP_MANY_NCCL")) return false;
  if (getenv("VDD_TP_MANY_PUSH")) return true;
  return R * tp <= 256;
}
// Row-parallel / vocab-parallel projection under tensor parallelism, many-row decode path: the tcgen05 GEMM's epilogue stores its
// tiles into every rank's peer-visible buffer at byte offset `dst_off` (plus nothing else: the caller folds rank / slot / column
// offsets into it) and every CTA bumps arrival counter `counter` on every rank.  *nctas = CTAs of the launch.
static int gemm_tp_push(vdd_engine* e, const void* x, const void* w, int M, int N, int K, size_t dst_off, int64_t ldc, int counter,
                        int* nctas, cudaStream_t st, bool pdl) {
  GemmArgs g{x, w, M, N, K, GEMM_OUT_F32, nullptr, nullptr, nullptr, nullptr, ldc};
  g.pdl = pdl;
  g.tp_n = e->tpc.tp;
  for (int k = 0; k < e->tpc.tp; ++k) {
    g.tp_out[k] = (float*)(uintptr_t)(e->tpc.peers[k] + dst_off);
    g.tp_cnt[k] = (unsigned*)(uintptr_t)e->tpc.peers[k] + counter;
  }
  g.tp_ctas = nctas;
  cudaError_t ce = launch_gemm(g, st);
  e->launches++;
  if (ce != cudaSuccess)
    return fail(VDD_ERR_CUDA, "tp gemm M=%d N=%d K=%d: %s %s", M, N, K, cudaGetErrorString(ce), gemm_init_error() ? gemm_init_error() : "");
  return VDD_OK;
}
#define RC(x)               \
  do {                      \
    int _r = (x);           \
    if (_r != VDD_OK) return _r; \
  } while (0)

// ------------------------------------------------------------------------------------------------ vision
static int encode_images_slot(vdd_engine* e, const void* pixels, int B, int slot, void* stream_v) {
  if (!e || !pixels || B <= 0) return fail(VDD_ERR_ARG, "bad argument");
  if (vdd_missing_tensors(e, nullptr, 0)) {
    char buf[512];
    vdd_missing_tensors(e, buf, sizeof buf);
    return fail(VDD_ERR_STATE, "weights missing: %s", buf);
  }
  cudaStream_t st = (cudaStream_t)stream_v;
  const vdd_config& c = e->c;
  const int NP = e->NP, VH = c.v_hidden, VI = c.v_intermediate, H = c.hidden_size, NT = NP + 1;
  const int Mp = B * NP, Mv = B * NT;
  // carve the vision workspace
  size_t o = 0;
  auto carve = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  const size_t o_patch = carve((size_t)Mp * e->Kpad * 2), o_pout = carve((size_t)Mp * VH * 4), o_res = carve((size_t)Mv * VH * 4),
               o_xn = carve((size_t)Mv * VH * 2), o_qkv = carve((size_t)Mv * 3 * VH * 2), o_att = carve((size_t)Mv * VH * 2),
               o_h = carve((size_t)Mv * VI * 2), o_pin = carve((size_t)Mp * VH * 2), o_ph = carve((size_t)Mp * H * 2),
               o_idx = carve((size_t)Mp * 4), o_meta = carve((size_t)B * 32);
  RC(ensure(e, e->vis, o));
  {  // two planes: [0] the real images, [1] the VCD-noised ones (images_cd); growing must keep plane 0
    const size_t need = (size_t)2 * Mp * H * 4;
    if (e->img_feats.bytes < need) {
      if (slot == 1 && e->image_B == B) return fail(VDD_ERR_STATE, "internal: feature buffer too small for the cd plane");
      RC(ensure(e, e->img_feats, need));
    }
  }
  char* base = (char*)e->vis.p;
  void* patches = base + o_patch; float* pout = (float*)(base + o_pout); float* res = (float*)(base + o_res);
  void* xn = base + o_xn; void* qkv = base + o_qkv; void* att = base + o_att; void* hb = base + o_h;
  void* pin = base + o_pin; void* ph = base + o_ph; int32_t* idx = (int32_t*)(base + o_idx);
  // per-image attention metadata + feature-select row list (drop CLS: clip_encoder.py:33-34)
  std::vector<int32_t> hidx(Mp), hq0(B), hlen(B);
  std::vector<int64_t> hoff(B), hhs(B);
  for (int b = 0; b < B; ++b) {
    hq0[b] = b * NT; hlen[b] = NT; hoff[b] = (int64_t)b * NT * 3 * VH; hhs[b] = 64;
    for (int p = 0; p < NP; ++p) hidx[(size_t)b * NP + p] = b * NT + 1 + p;
  }
  char* meta = base + o_meta;
  int32_t* d_q0 = (int32_t*)meta; int32_t* d_len = d_q0 + B; int64_t* d_off = (int64_t*)(meta + (size_t)B * 8);
  int64_t* d_hs = d_off + B;
  CK(cudaMemcpyAsync(idx, hidx.data(), (size_t)Mp * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_q0, hq0.data(), (size_t)B * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_len, hlen.data(), (size_t)B * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_off, hoff.data(), (size_t)B * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_hs, hhs.data(), (size_t)B * 8, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));  // host vectors go out of scope

  LK(e, launch_im2col(pixels, patches, B, c.image_size, c.patch_size, e->Kpad, st));
  RC(do_gemm(e, patches, e->patch_w, Mp, VH, e->Kpad, GEMM_OUT_F32, nullptr, nullptr, pout, nullptr, VH, st));
  LK(e, launch_vit_embed(pout, e->cls, e->vpos, res, B, NP, VH, st));
  LK(e, launch_layernorm(res, e->preln_w, e->preln_b, c.v_ln_eps, nullptr, res, Mv, VH, nullptr, st));
  for (int l = 0; l < c.v_layers_used; ++l) {
    const VLayerW& w = e->vw[l];
    LK(e, launch_layernorm(res, w.ln1w, w.ln1b, c.v_ln_eps, xn, nullptr, Mv, VH, nullptr, st));
    RC(do_gemm(e, xn, w.wqkv, Mv, 3 * VH, VH, GEMM_OUT_BF16, w.bqkv, nullptr, nullptr, qkv, 3 * VH, st));
    PrefillAttnArgs pa{};
    pa.q = qkv; pa.q_ld = 3 * VH;
    pa.k = (char*)qkv + (size_t)VH * 2; pa.v = (char*)qkv + (size_t)2 * VH * 2;
    pa.k_off = d_off; pa.k_head_stride = d_hs; pa.k_tok_stride = 3 * VH;
    pa.q_row0 = d_q0; pa.seqlen = d_len; pa.out = att; pa.out_ld = VH;
    pa.nseq = B; pa.heads = c.v_heads; pa.max_seqlen = NT; pa.causal = 0; pa.head_dim = 64; pa.scale = 0.125f;
    pa.q_rows = Mv; pa.kv_rows = Mv;
    LK(e, launch_prefill_attn(pa, st));
    RC(do_gemm(e, att, w.wo, Mv, VH, VH, GEMM_OUT_ADD_F32, w.bo, nullptr, res, nullptr, VH, st));
    LK(e, launch_layernorm(res, w.ln2w, w.ln2b, c.v_ln_eps, xn, nullptr, Mv, VH, nullptr, st));
    RC(do_gemm(e, xn, w.fc1, Mv, VI, VH, GEMM_OUT_BF16_QGELU, w.b1, nullptr, nullptr, hb, VI, st));
    RC(do_gemm(e, hb, w.fc2, Mv, VH, VI, GEMM_OUT_ADD_F32, w.b2, nullptr, res, nullptr, VH, st));
  }
  // feature_select('patch') then mm_projector (mlp2x_gelu)
  LK(e, launch_cast_rows_bf16(res, pin, Mp, VH, idx, st));
  RC(do_gemm(e, pin, e->pj_w1, Mp, H, VH, GEMM_OUT_BF16_GELU, e->pj_b1, nullptr, nullptr, ph, H, st));
  RC(do_gemm(e, ph, e->pj_w2, Mp, H, H, GEMM_OUT_F32, e->pj_b2, nullptr, (float*)e->img_feats.p + (size_t)slot * Mp * H, nullptr, H, st));
  if (slot == 0) e->image_B = B; else e->image_cd_B = B;
  return VDD_OK;
}
extern "C" int vdd_encode_images(vdd_engine* e, const void* pixels, int B, void* stream_v) {
  if (e) e->image_cd_B = 0;
  return encode_images_slot(e, pixels, B, 0, stream_v);
}
// images_cd of the VCD baseline (reference vcd_utils/vcd_add_noise.py + llava_llama.py:170): call after vdd_encode_images
extern "C" int vdd_encode_images_cd(vdd_engine* e, const void* pixels, int B, void* stream_v) {
  if (!e || e->image_B != B) return fail(VDD_ERR_STATE, "vdd_encode_images(B=%d) must precede vdd_encode_images_cd", B);
  return encode_images_slot(e, pixels, B, 1, stream_v);
}
extern "C" const float* vdd_image_features(vdd_engine* e) { return e ? (const float*)e->img_feats.p : nullptr; }

// ------------------------------------------------------------------------------------------------ prefill
static int tp_gather_logits(vdd_engine* e, cudaStream_t st);
static int prepare_step(vdd_engine* e, cudaStream_t st);

static int prefill_impl(vdd_engine* e, const int64_t* ids, int B, int T, int with_images, int use_unk, int use_none, int use_cd,
                        int max_new, void* stream_v);
extern "C" int vdd_prefill(vdd_engine* e, const int64_t* ids, int B, int T, int with_images, int use_unk, int use_none,
                           int max_new, void* stream_v) {
  return prefill_impl(e, ids, B, T, with_images, use_unk, use_none, 0, max_new, stream_v);
}
// main + VCD contrast stream (the same prompt with images_cd spliced in; reference vcd_sample.py:148-150 at step 0)
extern "C" int vdd_prefill_cd(vdd_engine* e, const int64_t* ids, int B, int T, int max_new, void* stream_v) {
  return prefill_impl(e, ids, B, T, 1, 0, 0, 1, max_new, stream_v);
}
static int prefill_impl(vdd_engine* e, const int64_t* ids, int B, int T, int with_images, int use_unk, int use_none, int use_cd,
                        int max_new, void* stream_v) {
  if (!e || !ids || B <= 0 || T <= 0 || max_new < 1) return fail(VDD_ERR_ARG, "bad argument");
  if (vdd_missing_tensors(e, nullptr, 0)) {
    char buf[512];
    vdd_missing_tensors(e, buf, sizeof buf);
    return fail(VDD_ERR_STATE, "weights missing: %s", buf);
  }
  cudaStream_t st = (cudaStream_t)stream_v;
  const vdd_config& c = e->c;
  const int H = c.hidden_size, NP = e->NP, L = c.num_layers, Hl = e->Hl, Il = e->Il;
  const int S = 1 + (use_unk ? 1 : 0) + (use_none ? 1 : 0) + (use_cd ? 1 : 0), R = S * B;
  if (use_cd && e->image_cd_B != B) return fail(VDD_ERR_STATE, "vdd_encode_images_cd(B=%d) must precede the cd prefill", B);
  if (with_images && e->image_B != B) return fail(VDD_ERR_STATE, "vdd_encode_images(B=%d) must precede prefill(B=%d)", e->image_B, B);

  // ---- row plan (vcd_sample.py:152-160: unk = sentinel -> token 0; none = sentinel deleted; main = 576 image rows)
  std::vector<int> kinds;  // 0 main, 1 unk, 2 none, 3 cd (main prompt with the noised image's features)
  kinds.push_back(0);
  if (use_unk) kinds.push_back(1);
  if (use_none) kinds.push_back(2);
  if (use_cd) kinds.push_back(3);  // last: its rows are dropped after step 0 (vdd_generate_cd)
  // ragged batches: positions masked out by the caller (padding, left or right) simply do not exist for any stream, so a padded
  // row behaves exactly like the same prompt submitted alone (the reference cannot do this: its decode step rebuilds the mask
  // as all ones, llava_arch.py:91-94)
  std::vector<uint8_t> pmask;
  pmask.swap(e->prompt_mask);  // one-shot
  if (!pmask.empty() && (e->prompt_mask_B != B || e->prompt_mask_T != T))
    return fail(VDD_ERR_ARG, "prompt mask is %dx%d, the prompt %dx%d", e->prompt_mask_B, e->prompt_mask_T, B, T);
  std::vector<int32_t> tok, img_row, row_seq, row_pos, seqlen(R), q_row0(R), kv_cap(R), last_rows(R);
  std::vector<int64_t> kv_off(R), head_stride(R);
  int64_t kv_total = 0;
  int max_len = 0;
  for (int s = 0; s < S; ++s) {
    for (int b = 0; b < B; ++b) {
      const int q = s * B + b;
      q_row0[q] = (int32_t)tok.size();
      int nsent = 0, pos = 0;
      for (int t = 0; t < T; ++t) {
        if (!pmask.empty() && !pmask[(size_t)b * T + t]) continue;  // padding
        const int64_t id = ids[(size_t)b * T + t];
        if (id == VDD_IMAGE_TOKEN_INDEX) {
          ++nsent;
          if (nsent > 1) return fail(VDD_ERR_UNSUPPORTED, "row %d holds more than one image sentinel", b);
          if (kinds[s] == 0 || kinds[s] == 3) {
            if (!with_images) return fail(VDD_ERR_ARG, "row %d has an image sentinel but no images were given", b);
            const int plane = kinds[s] == 3 ? B * NP : 0;
            for (int p = 0; p < NP; ++p) {
              tok.push_back(-1); img_row.push_back(plane + b * NP + p); row_seq.push_back(q); row_pos.push_back(pos++);
            }
          } else if (kinds[s] == 1) {
            tok.push_back(0); img_row.push_back(-1); row_seq.push_back(q); row_pos.push_back(pos++);
          }  // kind 2: dropped
        } else {
          if (id < 0 || id >= c.vocab_size) return fail(VDD_ERR_ARG, "token id %lld out of range", (long long)id);
          tok.push_back((int32_t)id); img_row.push_back(-1); row_seq.push_back(q); row_pos.push_back(pos++);
        }
      }
      if (pos == 0) return fail(VDD_ERR_ARG, "empty sequence");
      seqlen[q] = pos;
      last_rows[q] = q_row0[q] + pos - 1;
      kv_cap[q] = (int32_t)align_up((size_t)pos + max_new, 8);
      if (kv_cap[q] > c.max_position) return fail(VDD_ERR_ARG, "prompt %d + %d new tokens exceeds max_position %d", pos, max_new, c.max_position);
      kv_off[q] = kv_total;
      head_stride[q] = (int64_t)kv_cap[q] * 128;
      kv_total += (int64_t)kv_cap[q] * Hl;
      if (pos > max_len) max_len = pos;
    }
  }
  const int Mtot = (int)tok.size();
  e->B = B; e->T = T; e->S = S; e->R = R; e->Mtot = Mtot; e->max_new = max_new; e->max_len = max_len;
  e->kv_layer_stride = kv_total;
  e->main_len = seqlen[0];
  for (int b = 1; b < B; ++b)
    if (seqlen[b] != seqlen[0]) e->main_len = 0;
  e->max_splits = (max_len + max_new + 63) / 64;
  // streaming decode attention for many rows: each row gets the splits its capacity can ever need, long rows first
  std::vector<int32_t> item_row, item_split;
  e->attn_items = 0;
  e->attn_chunk = 0;
  const bool stream_small = getenv("VDD_ATTN_STREAM_SMALL") != nullptr;  // experiment: streaming kernel for R <= 8 too
  if ((R > 8 || stream_small) && !getenv("VDD_ATTN_LEGACY")) {
    // measured on B200 (config 3, 96 rows, ms per decode step): 64-token kernel 8.51; chunk 128: 9.14, 512: 7.85, 1024: 7.40 --
    // one CTA per (row, head) up to 1024 tokens beats finer splits + merges even though main rows are 7x the debias rows
    // under tensor parallelism a rank holds 1/tp of the heads: keep the CTA count (and the per-CTA serial KV stream) by splitting finer
    int chunk = getenv("VDD_ATTN_CHUNK") ? atoi(getenv("VDD_ATTN_CHUNK")) : (R > 8 ? std::max(256, 1024 / c.tp_size) : 128);
    if (chunk < 64 || chunk % 64) chunk = 1024;
    for (int q = 0; q < R; ++q)
      for (int sp = 0; sp * chunk < kv_cap[q]; ++sp) { item_row.push_back(q); item_split.push_back(sp); }
    e->attn_items = (int)item_row.size();
    e->attn_chunk = chunk;
  }
  e->h_item_row = item_row;

  // ---- buffers
  RC(ensure(e, e->resid, (size_t)Mtot * H * 4));
  RC(ensure(e, e->xn, (size_t)Mtot * H * 2));
  if (!gemm_qkv_rope_fused(Mtot, 3 * Hl, H)) RC(ensure(e, e->qkv, (size_t)Mtot * 3 * Hl * 4));  // fp32 projection output (unfused RoPE)
  RC(ensure(e, e->qb, (size_t)Mtot * Hl * 2));
  RC(ensure(e, e->attn, (size_t)Mtot * Hl * 2));
  RC(ensure(e, e->hbuf, (size_t)Mtot * Il * 2));
  {
    // The tcgen05 prefill attention multiplies P = 0 with the V rows between a sequence's end and the end of its last 64-key
    // tile: a fresh allocation may hold NaN bit patterns there, data of earlier generates is finite.  Zero once per allocation.
    const void* before = e->kv.p;
    RC(ensure(e, e->kv, (size_t)2 * L * kv_total * 2));
    if (e->kv.p != before) CK(cudaMemsetAsync(e->kv.p, 0, e->kv.bytes, st));
  }
  RC(ensure(e, e->logits, (size_t)R * c.vocab_size * 4));
  if (R > 8 && R <= 128) RC(ensure(e, e->split_ws, (size_t)8 * R * H * 4));
  if (c.tp_size > 1) {
    RC(ensure(e, e->logits_tp, (size_t)R * c.vocab_size * 4));
    RC(ensure(e, e->tp_tmp, (size_t)(Mtot > R ? Mtot : R) * H * 4));
    RC(ensure(e, e->misc, (size_t)R * c.vocab_size * 4));
  }
  {
    size_t o = 0;
    auto carve = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
    const size_t o_res = carve((size_t)R * H * 4), o_qkv = carve((size_t)R * 3 * Hl * 4),
                 // split partials: 64-token splits for decode_attn_kernel, 32-token splits for the whole-step kernel (twice as many)
                 o_po = carve((size_t)R * e->heads_l * (2 * e->max_splits + 1) * 128 * 4),
                 o_pm = carve((size_t)R * e->heads_l * (2 * e->max_splits + 1) * 8);
    RC(ensure(e, e->dec_f32, o));
    char* b = (char*)e->dec_f32.p;
    e->resid_d = (float*)(b + o_res); e->qkv_d = (float*)(b + o_qkv); e->part_o = (float*)(b + o_po); e->part_ml = (float*)(b + o_pm);
    o = 0;
    const size_t o_at = carve((size_t)R * Hl * 2), o_h = carve((size_t)R * Il * 2), o_xn = carve((size_t)R * H * 2);
    RC(ensure(e, e->dec_bf16, o));
    b = (char*)e->dec_bf16.p;
    e->attn_d = b + o_at; e->h_d = b + o_h; e->xn_d = b + o_xn;
  }
  {
    size_t o = 0;
    auto carve = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
    const size_t o_tok = carve((size_t)Mtot * 4), o_img = carve((size_t)Mtot * 4), o_rs = carve((size_t)Mtot * 4), o_rp = carve((size_t)Mtot * 4),
                 o_cap = carve((size_t)R * 4), o_pos = carve((size_t)R * 4), o_last = carve((size_t)R * 4), o_len = carve((size_t)R * 4),
                 o_q0 = carve((size_t)R * 4), o_step = carve(256), o_unf = carve((size_t)B * 4), o_cnt = carve((size_t)R * e->heads_l * 4),
                 o_off = carve((size_t)R * 8), o_hs = carve((size_t)R * 8), o_tokens = carve((size_t)B * 8),
                 o_seq = carve((size_t)B * max_new * 8), o_sa = carve(sizeof(SampleArgs)), o_bar = carve((size_t)(L + 2) * 4 * 4),
                 o_ir = carve((size_t)(e->attn_items + 1) * 4), o_is = carve((size_t)(e->attn_items + 1) * 4);
    RC(ensure(e, e->meta, o));
    char* b = (char*)e->meta.p;
    e->d_tok = (int32_t*)(b + o_tok); e->d_img_row = (int32_t*)(b + o_img); e->d_row_seq = (int32_t*)(b + o_rs);
    e->d_row_pos = (int32_t*)(b + o_rp); e->d_kv_cap = (int32_t*)(b + o_cap); e->d_pos = (int32_t*)(b + o_pos);
    e->d_last_rows = (int32_t*)(b + o_last); e->d_seqlen = (int32_t*)(b + o_len); e->d_q_row0 = (int32_t*)(b + o_q0);
    e->d_step = (int32_t*)(b + o_step); e->d_unfinished = (int32_t*)(b + o_unf); e->d_counters = (int32_t*)(b + o_cnt);
    e->d_kv_off = (int64_t*)(b + o_off); e->d_head_stride = (int64_t*)(b + o_hs); e->d_tokens = (int64_t*)(b + o_tokens);
    e->d_seq_out = (int64_t*)(b + o_seq); e->d_sargs = (SampleArgs*)(b + o_sa); e->d_bar = (unsigned*)(b + o_bar);
    e->d_item_row = (int32_t*)(b + o_ir); e->d_item_split = (int32_t*)(b + o_is);
  }
  auto up = [&](void* d, const void* h, size_t n) { return cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, st); };
  CK(up(e->d_tok, tok.data(), (size_t)Mtot * 4));
  CK(up(e->d_img_row, img_row.data(), (size_t)Mtot * 4));
  CK(up(e->d_row_seq, row_seq.data(), (size_t)Mtot * 4));
  CK(up(e->d_row_pos, row_pos.data(), (size_t)Mtot * 4));
  CK(up(e->d_kv_cap, kv_cap.data(), (size_t)R * 4));
  CK(up(e->d_pos, seqlen.data(), (size_t)R * 4));
  CK(up(e->d_last_rows, last_rows.data(), (size_t)R * 4));
  CK(up(e->d_seqlen, seqlen.data(), (size_t)R * 4));
  CK(up(e->d_q_row0, q_row0.data(), (size_t)R * 4));
  CK(up(e->d_kv_off, kv_off.data(), (size_t)R * 8));
  CK(up(e->d_head_stride, head_stride.data(), (size_t)R * 8));
  if (e->attn_items) {
    CK(up(e->d_item_row, item_row.data(), (size_t)e->attn_items * 4));
    CK(up(e->d_item_split, item_split.data(), (size_t)e->attn_items * 4));
  }
  CK(cudaMemsetAsync(e->d_counters, 0, (size_t)R * e->heads_l * 4, st));
  CK(cudaMemsetAsync(e->d_step, 0, 4, st));
  if (e->tp_fused_ready) {
    // arrival counters and the epoch restart together; peers cannot reach a decode step (and touch these counters)
    // before this rank has run the prefill collectives that are enqueued after this memset
    CK(cudaMemsetAsync(e->comm_buf, 0, 4096, st));
    CK(cudaMemsetAsync(e->d_epoch, 0, 4, st));
  }
  CK(cudaStreamSynchronize(st));  // host vectors go out of scope

  // ---- layers over the concatenated rows of all streams
  float* resid = (float*)e->resid.p;
  LK(e, launch_embed_rows(e->embed, e->d_tok, e->d_img_row, (const float*)e->img_feats.p, resid, Mtot, H, st));
  __nv_bfloat16* kbase = (__nv_bfloat16*)e->kv.p;
  for (int l = 0; l < L; ++l) {
    const LayerW& w = e->lw[l];
    void* kc = kbase + (size_t)l * kv_total;
    void* vc = kbase + (size_t)(L + l) * kv_total;
    LK(e, launch_rmsnorm_bf16(resid, w.ln1, c.rms_eps, e->xn.p, Mtot, H, nullptr, st));
    if (gemm_qkv_rope_fused(Mtot, 3 * Hl, H)) {
      // RoPE, the bf16 casts and the KV-cache scatter ride in the projection's epilogue: no fp32 [Mtot, 3*Hl] round trip
      QkvRopeEpi qr{e->d_row_seq, e->d_row_pos, e->d_kv_off, e->d_kv_cap, e->rope_cos, e->rope_sin, e->qb.p, kc, vc, Hl};
      GemmArgs g{e->xn.p, w.wqkv, Mtot, 3 * Hl, H, GEMM_OUT_QKV_ROPE, nullptr, nullptr, nullptr, nullptr, 3 * Hl};
      g.qkv_rope = &qr;
      LK(e, launch_gemm(g, st));
    } else {
      RC(do_gemm(e, e->xn.p, w.wqkv, Mtot, 3 * Hl, H, GEMM_OUT_F32, nullptr, nullptr, (float*)e->qkv.p, nullptr, 3 * Hl, st));
      LK(e, launch_rope_kv_prefill((const float*)e->qkv.p, Mtot, e->heads_l, 128, e->d_row_seq, e->d_row_pos, e->d_kv_off, e->d_kv_cap,
                                   e->rope_cos, e->rope_sin, e->qb.p, kc, vc, st));
    }
    PrefillAttnArgs pa{};
    pa.q = e->qb.p; pa.q_ld = Hl; pa.k = kc; pa.v = vc; pa.k_off = e->d_kv_off; pa.k_head_stride = e->d_head_stride;
    pa.k_tok_stride = 128; pa.q_row0 = e->d_q_row0; pa.seqlen = e->d_seqlen; pa.out = e->attn.p; pa.out_ld = Hl;
    pa.nseq = R; pa.heads = e->heads_l; pa.max_seqlen = max_len; pa.causal = 1; pa.head_dim = 128;
    pa.scale = 0.08838834764831845f;  // 1/sqrt(128)
    pa.q_rows = Mtot; pa.kv_rows = kv_total / 128;
    LK(e, launch_prefill_attn(pa, st));
    if (e->attn_all0) {  // attentions[0][l]: this layer's roped queries are in qb, its keys in the cache
      if (c.tp_size > 1) return fail(VDD_ERR_UNSUPPORTED, "attention maps are not gathered across tensor-parallel ranks");
      if (!e->main_len) return fail(VDD_ERR_UNSUPPORTED, "attention maps need main-stream rows of equal length");
      const int64_t per_layer = (int64_t)B * e->heads_l * e->main_len * e->main_len;
      if (per_layer * L != e->attn_all0_elems)
        return fail(VDD_ERR_ARG, "step-0 attention buffer holds %lld floats, %d layers need %lld", (long long)e->attn_all0_elems, L,
                    (long long)(per_layer * L));
      AttnProbsArgs ap{};
      ap.q = e->qb.p; ap.q_ld = Hl; ap.k = kc; ap.k_off = e->d_kv_off; ap.kv_cap = e->d_kv_cap; ap.q_row0 = e->d_q_row0;
      ap.out = e->attn_all0 + (size_t)l * per_layer; ap.B = B; ap.heads = e->heads_l; ap.len = e->main_len; ap.scale = 0.08838834764831845f;
      LK(e, launch_attn_probs(ap, st));
    }
    if (c.tp_size > 1) {
      RC(do_gemm(e, e->attn.p, w.wo, Mtot, H, Hl, GEMM_OUT_F32, nullptr, nullptr, (float*)e->tp_tmp.p, nullptr, H, st));
      RC(tp_allreduce_add(e, (float*)e->tp_tmp.p, resid, (int64_t)Mtot * H, st));
    } else {
      RC(do_gemm(e, e->attn.p, w.wo, Mtot, H, Hl, GEMM_OUT_ADD_F32, nullptr, nullptr, resid, nullptr, H, st));
    }
    LK(e, launch_rmsnorm_bf16(resid, w.ln2, c.rms_eps, e->xn.p, Mtot, H, nullptr, st));
    RC(do_gemm(e, e->xn.p, w.wgu, Mtot, 2 * Il, H, GEMM_OUT_SWIGLU, nullptr, nullptr, nullptr, e->hbuf.p, Il, st));
    if (c.tp_size > 1) {
      RC(do_gemm(e, e->hbuf.p, w.wdown, Mtot, H, Il, GEMM_OUT_F32, nullptr, nullptr, (float*)e->tp_tmp.p, nullptr, H, st));
      RC(tp_allreduce_add(e, (float*)e->tp_tmp.p, resid, (int64_t)Mtot * H, st));
    } else {
      RC(do_gemm(e, e->hbuf.p, w.wdown, Mtot, H, Il, GEMM_OUT_ADD_F32, nullptr, nullptr, resid, nullptr, H, st));
    }
  }
  // ---- lm_head on the LAST row of each sequence only (the reference computes all rows, llava_llama.py:103, and reads one, vcd_sample.py:119)
  LK(e, launch_rmsnorm_bf16(resid, e->final_norm, c.rms_eps, e->xn_d, R, H, e->d_last_rows, st));
  float* lg = c.tp_size > 1 ? (float*)e->logits_tp.p : (float*)e->logits.p;
  if (R <= 8) {
    GemvArgs g{};
    g.w = e->lm_head; g.x_bf16 = e->xn_d; g.R = R; g.N = e->Vl; g.K = H; g.epilogue = GEMV_OUT_F32; g.out_f32 = lg; g.ld_out = e->Vl;
    LK(e, launch_gemv(g, e->num_sms, false, st));
  } else {
    RC(do_gemm(e, e->xn_d, e->lm_head, R, e->Vl, H, GEMM_OUT_F32, nullptr, nullptr, lg, nullptr, e->Vl, st));
  }
  if (c.tp_size > 1) RC(tp_gather_logits(e, st));
  e->logits_cur = (float*)e->logits.p;
  e->prefilled = true;
  RC(prepare_step(e, st));
  if (e->attn_capture) {
    // the roped queries of the LAST layer are still in qb and its keys are in the cache: materialise the main stream's map
    float* out = e->attn_capture;
    e->attn_capture = nullptr;
    if (c.tp_size > 1) return fail(VDD_ERR_UNSUPPORTED, "attention maps are not gathered across tensor-parallel ranks");
    if (!e->main_len) return fail(VDD_ERR_UNSUPPORTED, "attention maps need main-stream rows of equal length");
    const int64_t need = (int64_t)B * e->heads_l * e->main_len * e->main_len;
    if (need != e->attn_capture_elems)
      return fail(VDD_ERR_ARG, "attention capture buffer holds %lld floats, the map needs %lld", (long long)e->attn_capture_elems, (long long)need);
    AttnProbsArgs ap{};
    ap.q = e->qb.p; ap.q_ld = Hl; ap.k = kbase + (size_t)(L - 1) * kv_total; ap.k_off = e->d_kv_off; ap.kv_cap = e->d_kv_cap;
    ap.q_row0 = e->d_q_row0; ap.out = out; ap.B = B; ap.heads = e->heads_l; ap.len = e->main_len; ap.scale = 0.08838834764831845f;
    LK(e, launch_attn_probs(ap, st));
  }
  return VDD_OK;
}

// ------------------------------------------------------------------------------------------------ decode step
// [tp, R, Vl] (all-gather order) -> [R, V]
__global__ void tp_logits_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int tp, int R, int Vl) {
  const int64_t n = (int64_t)tp * R * Vl;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % Vl);
    const int r = (int)((i / Vl) % R);
    const int k = (int)(i / ((int64_t)Vl * R));
    dst[(int64_t)r * tp * Vl + (int64_t)k * Vl + v] = src[i];
  }
}
static int tp_gather_logits(vdd_engine* e, cudaStream_t st) {
  const int tp = e->c.tp_size, R = e->R, Vl = e->Vl;
  float* tmp = (float*)e->misc.p;  // sized in vdd_prefill (no allocation inside graph capture)
  int r = g_nccl.AllGather(e->logits_tp.p, tmp, (size_t)R * Vl, 7, e->comm, st);
  if (r != 0) return fail(VDD_ERR_CUDA, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  e->launches++;
  tp_logits_transpose_kernel<<<296, 256, 0, st>>>(tmp, (float*)e->logits.p, tp, R, Vl);
  e->launches++;
  CK(cudaGetLastError());
  return VDD_OK;
}

static GemvArgs gv_args(const void* w, const void* x_bf16, const float* x_resid, const void* norm_w, float eps, int R, int N, int K,
                        int epi, float* of32, void* obf16, int64_t ld) {
  GemvArgs g{};
  g.w = w; g.x_bf16 = x_bf16; g.x_resid = x_resid; g.norm_w = norm_w; g.eps = eps; g.R = R; g.N = N; g.K = K;
  g.epilogue = epi; g.out_f32 = of32; g.out_bf16 = obf16; g.ld_out = ld;
  return g;
}

// Phase table of the whole-step kernel for the CURRENT plan (R rows): qkv_0 | per layer: attention, o, gate_up, down, next qkv /
// lm_head.  Built outside graph capture (prefill; again when the VCD contrast rows are dropped after step 0).
static int prepare_step(vdd_engine* e, cudaStream_t st) {
  e->step_ready = false;
  const vdd_config& c = e->c;
  const int H = c.hidden_size, L = c.num_layers, Hl = e->Hl, Il = e->Il, R = e->R;
  const bool tp = c.tp_size > 1;
  const bool tpf = tp && e->tp_fused_ready;
  if (!e->use_step || !e->use_fused || R > 4 || (tp && !tpf)) return VDD_OK;
  __nv_bfloat16* kbase = (__nv_bfloat16*)e->kv.p;
  float* lg = tp ? (float*)e->logits_tp.p : (float*)e->logits.p;
  std::vector<StepPhase> ph;
  auto gemv = [&](const GemvArgs& g) { StepPhase p{}; p.kind = STEP_GEMV; p.g = g; ph.push_back(p); };
  auto qkv = [&](int l) {
    return gv_args(e->lw[l].wqkv, nullptr, e->resid_d, e->lw[l].ln1, c.rms_eps, R, 3 * Hl, H, GEMV_OUT_F32, e->qkv_d, nullptr, 3 * Hl);
  };
  const int res_epi = tpf ? GEMV_OUT_TP_PARTIAL : GEMV_OUT_ADD_RESID;
  gemv(qkv(0));
  for (int l = 0; l < L; ++l) {
    const LayerW& w = e->lw[l];
    StepPhase a{};
    a.kind = STEP_ATTN;
    a.kcache = kbase + (size_t)l * e->kv_layer_stride;
    a.vcache = kbase + (size_t)(L + l) * e->kv_layer_stride;
    ph.push_back(a);
    GemvArgs o = gv_args(w.wo, e->attn_d, nullptr, nullptr, 0.f, R, H, Hl, res_epi, e->resid_d, nullptr, H);
    GemvArgs gu = gv_args(w.wgu, nullptr, e->resid_d, w.ln2, c.rms_eps, R, 2 * Il, H, GEMV_OUT_SWIGLU_BF16, nullptr, e->h_d, Il);
    GemvArgs dn = gv_args(w.wdown, e->h_d, nullptr, nullptr, 0.f, R, H, Il, res_epi, e->resid_d, nullptr, H);
    if (tpf) {
      o.tp_slot = 0; o.tp_counter = 2 * l;
      dn.tp_slot = 1; dn.tp_counter = 2 * l + 1;
    }
    gemv(o); gemv(gu); gemv(dn);
    if (l + 1 < L) {
      gemv(qkv(l + 1));
    } else {
      GemvArgs hd = gv_args(e->lm_head, nullptr, e->resid_d, e->final_norm, c.rms_eps, R, e->Vl, H, GEMV_OUT_F32, lg, nullptr, e->Vl);
      if (tpf) { hd.epilogue = GEMV_OUT_TP_LOGITS; hd.tp_counter = 2 * L; }
      gemv(hd);
    }
  }
  if (!step_kernel_supports(ph.data(), (int)ph.size(), R)) return VDD_OK;  // odd shapes keep the per-layer launches
  if (plan_decode_step(ph.data(), (int)ph.size(), R, e->num_sms, &e->step_launch) != cudaSuccess) return VDD_OK;
  if (step_build_tmaps(ph.data(), (int)ph.size()) != cudaSuccess) return VDD_OK;
  RC(ensure(e, e->step_tab, ph.size() * sizeof(StepPhase)));
  CK(cudaMemcpyAsync(e->step_tab.p, ph.data(), ph.size() * sizeof(StepPhase), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));  // the host vector goes out of scope
  e->step_R = R;
  e->step_ready = true;
  return VDD_OK;
}

static int decode_step_enqueue(vdd_engine* e, const int64_t* tokens, cudaStream_t st, bool projections_only = false) {
  const vdd_config& c = e->c;
  const int H = c.hidden_size, L = c.num_layers, Hl = e->Hl, Il = e->Il, R = e->R;
  const bool pdl = e->use_pdl;
  const bool tp = c.tp_size > 1;
  const bool tpf = tp && e->tp_fused_ready;             // all-reduce / all-gather inside the chain, over peer memory
  const bool fused = R <= 8 && (!tp || tpf) && e->use_fused;  // o -> gate_up -> down -> next qkv in one launch per layer
  __nv_bfloat16* kbase = (__nv_bfloat16*)e->kv.p;
  float* lg = tp ? (float*)e->logits_tp.p : (float*)e->logits.p;
  auto qkv_args = [&](int l) {
    return gv_args(e->lw[l].wqkv, nullptr, e->resid_d, e->lw[l].ln1, c.rms_eps, R, 3 * Hl, H, GEMV_OUT_F32, e->qkv_d, nullptr, 3 * Hl);
  };
  auto head_args = [&]() {
    GemvArgs g = gv_args(e->lm_head, nullptr, e->resid_d, e->final_norm, c.rms_eps, R, e->Vl, H, GEMV_OUT_F32, lg, nullptr, e->Vl);
    if (tpf && fused) { g.epilogue = GEMV_OUT_TP_LOGITS; g.tp_counter = 2 * L; }
    return g;
  };
  if (fused) CK(cudaMemsetAsync(e->d_bar, 0, (size_t)(L + 2) * 4 * 4, st));
  if (fused && e->step_ready && e->step_R == R && !e->attn_alln) {
    // ONE persistent launch for the whole step (decode_step.cu): weights streamed by a producer warp across all phases
    if (!projections_only) LK(e, launch_decode_embed(e->embed, tokens, e->resid_d, e->B, e->S, H, st));
    StepArgs sa{};
    sa.phases = (const StepPhase*)e->step_tab.p;
    sa.nph = 1 + 5 * L;
    sa.skip_attn = projections_only ? 1 : 0;
    sa.attn.qkv = e->qkv_d; sa.attn.kv_off = e->d_kv_off; sa.attn.kv_cap = e->d_kv_cap; sa.attn.pos = e->d_pos;
    sa.attn.rope_cos = e->rope_cos; sa.attn.rope_sin = e->rope_sin; sa.attn.out_bf16 = e->attn_d; sa.attn.part_o = e->part_o;
    sa.attn.part_ml = e->part_ml; sa.attn.counters = e->d_counters; sa.attn.heads = e->heads_l;
    sa.attn.max_splits = (e->max_len + e->max_new + 31) / 32;  // 32-token splits (<= 2 * max_splits: the partial buffers are sized for it)
    sa.attn.scale = 0.08838834764831845f;
    sa.barrier = e->d_bar;
    if (tpf) sa.tpc = e->tpc;
    sa.trace = e->d_trace;
    cudaError_t ce = launch_decode_step(sa, e->step_launch, R, pdl, st);
    e->launches++;
    if (ce != cudaSuccess) return fail(VDD_ERR_CUDA, "decode step kernel (R=%d, grid %d, %d+1 warps, ring depth %d, %zu B smem): %s", R,
                                       e->step_launch.grid, e->step_launch.nw, e->step_launch.depth, e->step_launch.smem, cudaGetErrorString(ce));
    if (!projections_only) e->logits_cur = tpf ? (float*)((char*)e->comm_buf + e->tpc.off_logits) : (float*)e->logits.p;
    return VDD_OK;
  }
  auto gemm = [&](const void* x, const void* w, int N, int K, int mode, float* of32, void* obf16, int64_t ld) -> int {
    return do_gemm(e, x, w, R, N, K, mode, nullptr, nullptr, of32, obf16, ld, st);
  };
  if (!projections_only) LK(e, launch_decode_embed(e->embed, tokens, e->resid_d, e->B, e->S, H, st));
  if (fused) LK(e, launch_gemv(qkv_args(0), e->num_sms, pdl, st));
  // 8 < R on one GPU, square-tile GEMMs: every launch of the layer is programmatically serialised (set-up overlaps the
  // predecessor's tail), and the split-K reduction of o_proj / down_proj is folded into the RMSNorm pass that follows it
  // ... and under tensor parallelism with mapped peers: same launches, the all-reduce rides in the o_proj / down_proj epilogues
  // (peer stores + arrival counters) and in the reduce + RMSNorm pass that follows them; the logits are gathered the same way
  // Push-to-all costs tp x the message over NVLink and tp stores per accumulator in the epilogue: it wins while the message is
  // small (configs[2] shape at TP=2: 6.48 vs 6.67 ms per step with NCCL launches; configs[3]: 16 rows at TP=4) and loses badly when
  // it is not (configs[4]: 192 rows x 8 ranks = 31 MB per all-reduce, 23.2 ms per step against 11.4 ms with NCCL): rows x ranks <= 256.
  const bool tpm = tp && tpf && R > 8 && R <= e->tp_max_rows && tp_push_pays(R, c.tp_size);
  const bool wide = R > 8 && (!tp || tpm);
  const size_t slot_bytes = (size_t)c.tp_size * e->tp_max_rows * H * 4;   // one inbox slot on every rank
  const size_t plane_off = (size_t)c.tp_rank * R * H * 4;                 // this rank's plane inside a slot
  const float* inbox_local = tpm ? (const float*)((const char*)e->comm_buf + e->tpc.off_inbox) : nullptr;
  const unsigned* cnt_local = tpm ? (const unsigned*)e->comm_buf : nullptr;
  if (wide) LK(e, launch_rmsnorm_bf16(e->resid_d, e->lw[0].ln1, c.rms_eps, e->xn_d, R, H, nullptr, st, pdl));
  for (int l = 0; l < L; ++l) {
    const LayerW& w = e->lw[l];
    if (wide) {
      RC(do_gemm(e, e->xn_d, w.wqkv, R, 3 * Hl, H, GEMM_OUT_F32, nullptr, nullptr, e->qkv_d, nullptr, 3 * Hl, st, pdl));
    } else if (!fused) {
      if (R <= 8) {
        LK(e, launch_gemv(qkv_args(l), e->num_sms, pdl, st));
      } else {
        LK(e, launch_rmsnorm_bf16(e->resid_d, w.ln1, c.rms_eps, e->xn_d, R, H, nullptr, st));
        RC(gemm(e->xn_d, w.wqkv, 3 * Hl, H, GEMM_OUT_F32, e->qkv_d, nullptr, 3 * Hl));
      }
    }
    DecodeAttnArgs da{};
    da.qkv = e->qkv_d; da.kcache = kbase + (size_t)l * e->kv_layer_stride; da.vcache = kbase + (size_t)(L + l) * e->kv_layer_stride;
    da.kv_off = e->d_kv_off; da.kv_cap = e->d_kv_cap; da.pos = e->d_pos; da.rope_cos = e->rope_cos; da.rope_sin = e->rope_sin;
    da.out_bf16 = e->attn_d; da.part_o = e->part_o; da.part_ml = e->part_ml; da.counters = e->d_counters;
    da.R = R; da.heads = e->heads_l; da.max_splits = e->max_splits; da.scale = 0.08838834764831845f;
    if (e->attn_items) { da.item_row = e->d_item_row; da.item_split = e->d_item_split; da.n_items = e->attn_items; da.chunk = e->attn_chunk; }
    if (!projections_only) LK(e, launch_decode_attn(da, st));
    if (e->attn_alln && !projections_only) {  // attentions[step][l] of the main rows; qkv_d still holds this layer's query
      AttnProbsDecodeArgs ap{};
      ap.qkv = e->qkv_d; ap.k = da.kcache; ap.k_off = e->d_kv_off; ap.kv_cap = e->d_kv_cap; ap.pos = e->d_pos; ap.rope_cos = e->rope_cos;
      ap.rope_sin = e->rope_sin; ap.step = e->d_step; ap.out = e->attn_alln; ap.B = e->B; ap.heads = e->heads_l; ap.layers = L; ap.layer = l;
      ap.W = e->attn_all_w; ap.scale = 0.08838834764831845f;
      LK(e, launch_attn_probs_decode(ap, st));
    }
    if (fused) {
      GemvChain ch{};
      const int res_epi = tpf ? GEMV_OUT_TP_PARTIAL : GEMV_OUT_ADD_RESID;
      ch.ph[0] = gv_args(w.wo, e->attn_d, nullptr, nullptr, 0.f, R, H, Hl, res_epi, e->resid_d, nullptr, H);
      ch.ph[1] = gv_args(w.wgu, nullptr, e->resid_d, w.ln2, c.rms_eps, R, 2 * Il, H, GEMV_OUT_SWIGLU_BF16, nullptr, e->h_d, Il);
      ch.ph[2] = gv_args(w.wdown, e->h_d, nullptr, nullptr, 0.f, R, H, Il, res_epi, e->resid_d, nullptr, H);
      if (tpf) {
        ch.ph[0].tp_slot = 0; ch.ph[0].tp_counter = 2 * l;
        ch.ph[2].tp_slot = 1; ch.ph[2].tp_counter = 2 * l + 1;
        ch.tpc = e->tpc;
      }
      ch.ph[3] = l + 1 < L ? qkv_args(l + 1) : head_args();
      ch.n = 4;
      ch.barrier = e->d_bar + (size_t)l * 4;
      ch.trace = (l == L / 2) ? e->d_trace : nullptr;
      LK(e, launch_gemv_chain(ch, e->num_sms, pdl, st));
      continue;
    }
    if (wide && tpm) {
      int nct = 0;
      RC(gemm_tp_push(e, e->attn_d, w.wo, R, H, Hl, (size_t)e->tpc.off_inbox + plane_off, H, 2 * l, &nct, st, pdl));
      LK(e, launch_reduce_rmsnorm(e->resid_d, inbox_local, c.tp_size, R, H, w.ln2, c.rms_eps, e->xn_d, st, pdl, cnt_local + 2 * l,
                                  (unsigned)(c.tp_size * nct), e->d_epoch));
      RC(do_gemm(e, e->xn_d, w.wgu, R, 2 * Il, H, GEMM_OUT_SWIGLU, nullptr, nullptr, nullptr, e->h_d, Il, st, pdl));
      RC(gemm_tp_push(e, e->h_d, w.wdown, R, H, Il, (size_t)e->tpc.off_inbox + slot_bytes + plane_off, H, 2 * l + 1, &nct, st, pdl));
      LK(e, launch_reduce_rmsnorm(e->resid_d, (const float*)((const char*)inbox_local + slot_bytes), c.tp_size, R, H,
                                  l + 1 < L ? e->lw[l + 1].ln1 : e->final_norm, c.rms_eps, e->xn_d, st, pdl, cnt_local + 2 * l + 1,
                                  (unsigned)(c.tp_size * nct), e->d_epoch));
      continue;
    }
    if (wide) {
      int sp = 0;
      RC(do_gemm(e, e->attn_d, w.wo, R, H, Hl, GEMM_OUT_ADD_F32, nullptr, nullptr, e->resid_d, nullptr, H, st, pdl, &sp));
      LK(e, launch_reduce_rmsnorm(e->resid_d, (const float*)e->split_ws.p, sp, R, H, w.ln2, c.rms_eps, e->xn_d, st, pdl));
      RC(do_gemm(e, e->xn_d, w.wgu, R, 2 * Il, H, GEMM_OUT_SWIGLU, nullptr, nullptr, nullptr, e->h_d, Il, st, pdl));
      RC(do_gemm(e, e->h_d, w.wdown, R, H, Il, GEMM_OUT_ADD_F32, nullptr, nullptr, e->resid_d, nullptr, H, st, pdl, &sp));
      LK(e, launch_reduce_rmsnorm(e->resid_d, (const float*)e->split_ws.p, sp, R, H, l + 1 < L ? e->lw[l + 1].ln1 : e->final_norm,
                                  c.rms_eps, e->xn_d, st, pdl));
      continue;
    }
    float* o_dst = tp ? (float*)e->tp_tmp.p : e->resid_d;
    const int o_mode_v = tp ? GEMV_OUT_F32 : GEMV_OUT_ADD_RESID;
    const int o_mode_m = tp ? GEMM_OUT_F32 : GEMM_OUT_ADD_F32;
    if (R <= 8) {
      LK(e, launch_gemv(gv_args(w.wo, e->attn_d, nullptr, nullptr, 0.f, R, H, Hl, o_mode_v, o_dst, nullptr, H), e->num_sms, pdl, st));
    } else {
      RC(gemm(e->attn_d, w.wo, H, Hl, o_mode_m, o_dst, nullptr, H));
    }
    if (tp) RC(tp_allreduce_add(e, o_dst, e->resid_d, (int64_t)R * H, st));
    if (R <= 8) {
      LK(e, launch_gemv(gv_args(w.wgu, nullptr, e->resid_d, w.ln2, c.rms_eps, R, 2 * Il, H, GEMV_OUT_SWIGLU_BF16, nullptr, e->h_d, Il),
                        e->num_sms, pdl, st));
      LK(e, launch_gemv(gv_args(w.wdown, e->h_d, nullptr, nullptr, 0.f, R, H, Il, o_mode_v, o_dst, nullptr, H), e->num_sms, pdl, st));
    } else {
      LK(e, launch_rmsnorm_bf16(e->resid_d, w.ln2, c.rms_eps, e->xn_d, R, H, nullptr, st));
      RC(gemm(e->xn_d, w.wgu, 2 * Il, H, GEMM_OUT_SWIGLU, nullptr, e->h_d, Il));
      RC(gemm(e->h_d, w.wdown, H, Il, o_mode_m, o_dst, nullptr, H));
    }
    if (tp) RC(tp_allreduce_add(e, o_dst, e->resid_d, (int64_t)R * H, st));
  }
  if (!fused) {
    if (R <= 8) {
      LK(e, launch_gemv(head_args(), e->num_sms, pdl, st));
    } else if (wide && tpm) {
      // vocab-parallel lm_head: this rank's column block lands in every rank's gathered logits; one waiting thread holds the
      // stream until all blocks are there
      int nct = 0;
      RC(gemm_tp_push(e, e->xn_d, e->lm_head, R, e->Vl, H, (size_t)e->tpc.off_logits + (size_t)c.tp_rank * e->Vl * 4, c.vocab_size, 2 * L,
                      &nct, st, pdl));
      LK(e, launch_tp_wait(cnt_local + 2 * L, (unsigned)(c.tp_size * nct), e->d_epoch, st));
    } else if (wide) {
      RC(do_gemm(e, e->xn_d, e->lm_head, R, e->Vl, H, GEMM_OUT_F32, nullptr, nullptr, lg, nullptr, e->Vl, st, pdl));
    } else {
      LK(e, launch_rmsnorm_bf16(e->resid_d, e->final_norm, c.rms_eps, e->xn_d, R, H, nullptr, st));
      RC(gemm(e->xn_d, e->lm_head, e->Vl, H, GEMM_OUT_F32, lg, nullptr, e->Vl));
    }
  }
  if (tp && !(tpf && fused) && !tpm) RC(tp_gather_logits(e, st));
  if (!projections_only) e->logits_cur = ((tpf && fused) || tpm) ? (float*)((char*)e->comm_buf + e->tpc.off_logits) : (float*)e->logits.p;
  return VDD_OK;
}

extern "C" int vdd_decode_step(vdd_engine* e, const int64_t* tokens, void* stream_v) {
  if (!e || !tokens) return fail(VDD_ERR_ARG, "null argument");
  if (!e->prefilled) return fail(VDD_ERR_STATE, "vdd_prefill must run first");
  cudaStream_t st = (cudaStream_t)stream_v;
  RC(decode_step_enqueue(e, tokens, st));
  LK(e, launch_advance(e->d_pos, e->R, nullptr, st, e->d_epoch));  // pos += 1 (and the TP epoch)
  return VDD_OK;
}
extern "C" int vdd_debug_gemm_tile_coords(int tile, int m_tiles, int n_tiles, int* m_idx_host, int* n_idx_host) {
  if (!m_idx_host || !n_idx_host || m_tiles < 1 || n_tiles < 1 || tile < 0 || tile >= m_tiles * n_tiles) return fail(VDD_ERR_ARG, "bad argument");
  gemm_tile_coords_host(tile, m_tiles, n_tiles, m_idx_host, n_idx_host);
  return VDD_OK;
}
extern "C" int vdd_debug_prefill_attn(const void* q, int64_t q_ld, int64_t q_rows, const void* k, const void* v, int64_t kv_rows,
                                      int64_t k_tok_stride, const int64_t* k_off, const int64_t* k_head_stride, const int32_t* q_row0,
                                      const int32_t* seqlen, void* out, int64_t out_ld, int nseq, int heads, int max_seqlen, int causal,
                                      int head_dim, float scale, int use_tc, void* stream) {
  if (!q || !k || !v || !k_off || !k_head_stride || !q_row0 || !seqlen || !out) return fail(VDD_ERR_ARG, "null argument");
  PrefillAttnArgs pa{};
  pa.q = q; pa.q_ld = q_ld; pa.k = k; pa.v = v; pa.k_off = k_off; pa.k_head_stride = k_head_stride; pa.k_tok_stride = k_tok_stride;
  pa.q_row0 = q_row0; pa.seqlen = seqlen; pa.out = out; pa.out_ld = out_ld; pa.nseq = nseq; pa.heads = heads; pa.max_seqlen = max_seqlen;
  pa.causal = causal; pa.head_dim = head_dim; pa.scale = scale; pa.q_rows = q_rows; pa.kv_rows = kv_rows;
  const cudaError_t ce = use_tc ? launch_prefill_attn_tc(pa, (cudaStream_t)stream) : launch_prefill_attn_mma(pa, (cudaStream_t)stream);
  if (ce != cudaSuccess) return fail(VDD_ERR_CUDA, "prefill attention (%s): %s", use_tc ? "tcgen05" : "mma.sync", cudaGetErrorString(ce));
  return VDD_OK;
}
extern "C" int vdd_set_attention_capture(vdd_engine* e, float* out, int64_t elems) {
  if (!e) return fail(VDD_ERR_ARG, "null engine");
  e->attn_capture = out;
  e->attn_capture_elems = out ? elems : 0;
  return VDD_OK;
}
extern "C" int vdd_set_attention_capture_all(vdd_engine* e, float* step0, int64_t step0_elems, float* later, int64_t later_elems) {
  if (!e) return fail(VDD_ERR_ARG, "null engine");
  e->attn_all0 = step0;
  e->attn_all0_elems = step0 ? step0_elems : 0;
  e->attn_alln = later;
  e->attn_alln_elems = later ? later_elems : 0;
  return VDD_OK;
}
extern "C" int vdd_set_sampler_rng(vdd_engine* e, uint64_t seed, uint64_t offset) {
  if (!e) return fail(VDD_ERR_ARG, "null engine");
  if (offset % 4) return fail(VDD_ERR_ARG, "a torch CUDA generator offset is a multiple of 4");
  e->rng_armed = true;
  e->rng_seed = seed;
  e->rng_offset = offset;
  return VDD_OK;
}
extern "C" uint64_t vdd_sampler_rng_advance(vdd_engine* e) { return e ? e->rng_advance : 0; }
extern "C" int vdd_debug_philox_exponential(uint64_t seed, uint64_t offset, int64_t numel, float* out_dev, void* stream_v) {
  if (!out_dev || numel <= 0) return fail(VDD_ERR_ARG, "bad argument");
  int dev = 0, sms = 148, mt = 2048;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&mt, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
  cudaError_t ce = launch_philox_exponential(seed, offset, numel, sms, mt, out_dev, (cudaStream_t)stream_v);
  if (ce != cudaSuccess) return fail(VDD_ERR_CUDA, "philox_exponential: %s", cudaGetErrorString(ce));
  return VDD_OK;
}
extern "C" int vdd_set_prompt_mask(vdd_engine* e, const uint8_t* mask_host, int B, int T) {
  if (!e) return fail(VDD_ERR_ARG, "null engine");
  if (!mask_host) { e->prompt_mask.clear(); e->prompt_mask_B = e->prompt_mask_T = 0; return VDD_OK; }
  if (B <= 0 || T <= 0) return fail(VDD_ERR_ARG, "bad mask shape");
  e->prompt_mask.assign(mask_host, mask_host + (size_t)B * T);
  e->prompt_mask_B = B;
  e->prompt_mask_T = T;
  return VDD_OK;
}
extern "C" int vdd_set_stream_scores_capture(vdd_engine* e, float* out, int64_t elems) {
  if (!e) return fail(VDD_ERR_ARG, "null engine");
  e->stream_scores = out;
  e->stream_scores_elems = out ? elems : 0;
  return VDD_OK;
}
extern "C" float* vdd_logits(vdd_engine* e) { return e ? (e->logits_cur ? e->logits_cur : (float*)e->logits.p) : nullptr; }
extern "C" int vdd_num_streams(vdd_engine* e) { return e ? e->S : 0; }

// ------------------------------------------------------------------------------------------------ K8 stand-alone
static void fill_sample_args(SampleArgs& a, const vdd_sample_params& p) {
  a.one_plus_alpha = (float)(1.0 + (double)p.cd_alpha);
  a.alpha = p.cd_alpha;
  a.log_beta = p.log_cd_beta;
  a.temperature = p.temperature;
  a.top_p = p.top_p;
  a.top_k = p.top_k;
  a.use_contrast = p.use_contrast;
  a.eos = p.eos_token_id;
  a.pad = p.pad_token_id;
}

extern "C" int vdd_logit_combine_sample(const float* l, const float* u, const float* d, int B, int V, int64_t ld,
                                        const vdd_sample_params* p, const float* q, float* scores_out, int64_t* tokens_out,
                                        int32_t* unfinished, void* stream_v) {
  if (!l || !p || B <= 0 || V <= 0) return fail(VDD_ERR_ARG, "bad argument");
  if (V > 32768) return fail(VDD_ERR_UNSUPPORTED, "vocab %d > 32768", V);
  if (!(p->temperature > 0.f)) return fail(VDD_ERR_ARG, "temperature must be > 0");
  SampleArgs a{};
  a.l = l; a.u = u; a.d = d; a.q = q; a.scores_out = scores_out; a.tokens_out = tokens_out; a.unfinished = unfinished;
  a.ld = ld; a.B = B; a.V = V;
  fill_sample_args(a, *p);
  cudaError_t ce = launch_combine_sample(a, nullptr, B, V, (cudaStream_t)stream_v);
  if (ce != cudaSuccess) return fail(VDD_ERR_CUDA, "combine_sample: %s", cudaGetErrorString(ce));
  return VDD_OK;
}

// K8 alone, `reps` launches captured in ONE CUDA graph and replayed: the per-launch time is the kernel's, not the host's
// (bench.py's logit-combine GB/s; at B <= 64 a Python-driven loop measures ctypes + allocator overhead instead).
extern "C" int vdd_bench_combine(const float* l, const float* u, const float* d, int B, int V, int64_t ld, const vdd_sample_params* p,
                                 const float* q, int64_t* tokens_out, int reps, float* ms_out, void* stream_v) {
  if (!l || !p || !ms_out || !tokens_out || B <= 0 || V <= 0 || reps < 1) return fail(VDD_ERR_ARG, "bad argument");
  if (V > 32768) return fail(VDD_ERR_UNSUPPORTED, "vocab %d > 32768", V);
  cudaStream_t st = (cudaStream_t)stream_v;
  if (!st) return fail(VDD_ERR_ARG, "graph capture needs a non-default stream");
  SampleArgs a{};
  a.l = l; a.u = u; a.d = d; a.q = q; a.tokens_out = tokens_out; a.ld = ld; a.B = B; a.V = V;
  fill_sample_args(a, *p);
  if (!q) {  // the product path of generate(): the Exp(1) variates come from Philox inside the kernel, no [B, V] read
    int dev = 0, sms = 0, mt = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&mt, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
    sample_philox_geometry(a, (long long)B * V, sms, mt);
    a.philox_seed = 0x5eed;
    a.philox_offset = 0;
    a.use_philox = 1;
  }
  cudaGraph_t g = nullptr;
  cudaGraphExec_t ge = nullptr;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  cudaError_t ce = cudaSuccess;
  for (int i = 0; i < reps && ce == cudaSuccess; ++i) ce = launch_combine_sample(a, nullptr, B, V, st);
  cudaError_t ee = cudaStreamEndCapture(st, &g);
  if (ce != cudaSuccess || ee != cudaSuccess) {
    if (g) cudaGraphDestroy(g);
    return fail(VDD_ERR_CUDA, "combine graph capture: %s", cudaGetErrorString(ce != cudaSuccess ? ce : ee));
  }
  ce = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (ce != cudaSuccess) return fail(VDD_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int it = 0; it < 4; ++it) {  // first replay warms up; keep the best of the rest
    cudaEventRecord(e0, st);
    cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaGraphExecDestroy(ge);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(VDD_ERR_CUDA, "combine graph replay: %s", cudaGetErrorString(ce));
  *ms_out = best / (float)reps;
  return VDD_OK;
}

// ------------------------------------------------------------------------------------------------ generate
static int upload_and_encode(vdd_engine* e, const void* pixels, int B, int slot, cudaStream_t st) {
  const vdd_config& c = e->c;
  const size_t nb = (size_t)B * 3 * c.image_size * c.image_size * 2;
  cudaPointerAttributes pa{};
  const bool on_device = cudaPointerGetAttributes(&pa, pixels) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  if (on_device) return encode_images_slot(e, pixels, B, slot, st);  // already resident (bench.py's device-only leg)
  RC(ensure(e, e->pix, nb));
  CK(cudaMemcpyAsync(e->pix.p, pixels, nb, cudaMemcpyHostToDevice, st));
  return encode_images_slot(e, e->pix.p, B, slot, st);
}

extern "C" int vdd_generate(vdd_engine* e, const int64_t* ids, int B, int T, const void* pixels_host, int use_unk, int use_none,
                            const vdd_sample_params* p, int max_new, const float* q_exp, float* scores_out,
                            int64_t* tokens_out_host, int* steps_out, void* stream_v) {
  return vdd_generate_cd(e, ids, B, T, pixels_host, nullptr, use_unk, use_none, p, max_new, q_exp, scores_out, tokens_out_host,
                         steps_out, stream_v);
}

extern "C" int vdd_generate_cd(vdd_engine* e, const int64_t* ids, int B, int T, const void* pixels_host, const void* pixels_cd_host,
                               int use_unk, int use_none, const vdd_sample_params* p, int max_new, const float* q_exp,
                               float* scores_out, int64_t* tokens_out_host, int* steps_out, void* stream_v) {
  if (!e || !ids || !p || !tokens_out_host || max_new < 1) return fail(VDD_ERR_ARG, "bad argument");
  if (e->c.vocab_size > 32768) return fail(VDD_ERR_UNSUPPORTED, "vocab > 32768");
  if (!(p->temperature > 0.f)) return fail(VDD_ERR_ARG, "temperature must be > 0");
  cudaStream_t st = (cudaStream_t)stream_v;
  const vdd_config& c = e->c;
  CK(cudaEventRecord(e->ev[0], st));
  const bool use_cd = pixels_cd_host != nullptr;
  if (use_cd && !pixels_host) return fail(VDD_ERR_ARG, "images_cd needs images");
  if (use_cd) {
    // reference vcd_sample.py:148: `if use_cd:` wins the first contrast forward, so the unk stream never runs; the 'none'
    // stream is added by :171-185 only when BOTH use_dd and use_dd_unk are set
    use_none = (use_unk && use_none) ? 1 : 0;
    use_unk = 0;
  }
  if (pixels_host) {
    e->image_cd_B = 0;
    RC(upload_and_encode(e, pixels_host, B, 0, st));
    if (use_cd) RC(upload_and_encode(e, pixels_cd_host, B, 1, st));
  }
  CK(cudaEventRecord(e->ev[1], st));
  RC(prefill_impl(e, ids, B, T, pixels_host != nullptr, use_unk, use_none, use_cd ? 1 : 0, max_new, st));
  e->attn_all0 = nullptr;  // consumed by the prefill
  if (e->attn_alln) {      // rows of the later steps: [max_new - 1, L, B, heads, len + max_new - 1]
    if (max_new < 2) {
      e->attn_alln = nullptr;
    } else {
      if (c.tp_size > 1 || !e->main_len) { e->attn_alln = nullptr; return fail(VDD_ERR_UNSUPPORTED, "attention maps: one rank, main-stream rows of equal length"); }
      e->attn_all_w = e->main_len + max_new - 1;
      const int64_t need = (int64_t)(max_new - 1) * c.num_layers * B * e->heads_l * e->attn_all_w;
      if (need != e->attn_alln_elems) {
        e->attn_alln = nullptr;
        return fail(VDD_ERR_ARG, "decode attention buffer holds %lld floats, %d steps need %lld", (long long)e->attn_alln_elems, max_new - 1,
                    (long long)need);
      }
    }
  }
  const int V = c.vocab_size;
  int S = e->S;
  // device-resident sampler arguments: the graph reads them through a pointer, so one graph serves every call
  SampleArgs a{};
  float* lg = (float*)e->logits.p;
  a.l = lg;
  int si = 1;  // stream order: main, unk, none, cd
  a.u = use_unk ? lg + (size_t)(si++) * B * V : nullptr;
  a.d = use_none ? lg + (size_t)(si++) * B * V : nullptr;
  if (use_cd) a.u = lg + (size_t)(si++) * B * V;
  a.q = q_exp; a.scores_out = scores_out; a.tokens_out = e->d_tokens; a.seq_out = e->d_seq_out; a.seq_ld = max_new;
  a.unfinished = e->d_unfinished; a.step_ptr = e->d_step; a.ld = V; a.B = B; a.V = V;
  fill_sample_args(a, *p);
  a.use_contrast = (p->use_contrast && S > 1) ? 1 : 0;
  // Variates: explicit q_exp (tests, callers that drew them), or generated inside K8 from the caller's torch generator state so that
  // the tokens are the ones the reference's torch.multinomial would draw on this device under the same seed -- one exponential_
  // call on [B, V] per step; the generator is advanced by the caller afterwards (vdd_sampler_rng_advance).
  e->rng_advance = 0;
  const bool rng = e->rng_armed;
  e->rng_armed = false;
  if (rng) {
    sample_philox_geometry(a, (long long)B * V, e->num_sms, e->max_threads_per_sm);
    a.philox_seed = e->rng_seed;
    a.philox_offset = e->rng_offset;
    a.use_philox = q_exp == nullptr ? 1 : 0;
  }
  if (e->pinned_bytes < sizeof(SampleArgs) + (size_t)B * 4 + 64) {
    if (e->pinned) cudaFreeHost(e->pinned);
    e->pinned_bytes = sizeof(SampleArgs) + (size_t)B * 4 + 4096;
    CK(cudaMallocHost(&e->pinned, e->pinned_bytes));
  }
  memcpy(e->pinned, &a, sizeof a);
  int32_t* h_unf = (int32_t*)((char*)e->pinned + align_up(sizeof(SampleArgs), 64));
  for (int b = 0; b < B; ++b) h_unf[b] = 1;
  CK(cudaMemcpyAsync(e->d_sargs, e->pinned, sizeof a, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(e->d_unfinished, h_unf, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  LK(e, launch_combine_sample(a, e->d_sargs, B, V, st));  // step 0 on the prefill logits
  if (e->stream_scores) {
    // The post-hoc-debias drivers run two MORE generates per question on the image-free prompts only to read their first-step
    // scores (reference llava_calibrate.py:192-194 -> :62-82, scores[0]).  Those prompts are this call's debias streams: their
    // step-0 logits already sit beside the main stream's, so the same warpers without a contrast (alpha = 0) give the scores.
    float* out = e->stream_scores;
    e->stream_scores = nullptr;
    if (e->stream_scores_elems != (int64_t)(S - 1) * B * V)
      return fail(VDD_ERR_ARG, "stream-scores buffer holds %lld floats, %d debias streams need %lld", (long long)e->stream_scores_elems,
                  S - 1, (long long)(S - 1) * B * V);
    for (int sidx = 1; sidx < S; ++sidx) {
      SampleArgs b{};
      b.l = lg + (size_t)sidx * B * V;
      b.scores_out = out + (size_t)(sidx - 1) * B * V;
      b.ld = V; b.B = B; b.V = V;
      fill_sample_args(b, *p);
      b.use_contrast = 0;
      b.eos = -1;
      LK(e, launch_combine_sample(b, nullptr, B, V, st));
    }
  }
  LK(e, launch_advance(nullptr, 0, e->d_step, st));  // step = 1
  if (use_cd) {
    // reference vcd_sample.py:148-150: model_kwargs_cd is re-copied from the MAIN stream's kwargs every step, so from
    // step 1 on the "contrast" forward runs on the main KV cache and its logits equal the main logits (SURVEY Appendix A
    // Q7).  The cd rows are therefore dropped after step 0 and the combine reads l as its own contrast.
    CK(cudaStreamSynchronize(st));
    e->S = S - 1;
    e->R = e->S * B;
    S = e->S;
    if (e->attn_items) {  // the work table is ordered by row: keep the surviving rows' items
      int n = 0;
      while (n < (int)e->h_item_row.size() && e->h_item_row[n] < e->R) ++n;
      e->attn_items = n;
    }
    a.u = a.l;
    memcpy(e->pinned, &a, sizeof a);
    CK(cudaMemcpyAsync(e->d_sargs, e->pinned, sizeof a, cudaMemcpyHostToDevice, st));
    RC(prepare_step(e, st));  // the phase table is per row count
  }
  const bool tp_many = e->tp_fused_ready && e->R > 8 && e->R <= e->tp_max_rows && tp_push_pays(e->R, c.tp_size);
  if (e->tp_fused_ready && ((e->R <= 8 && e->use_fused) || tp_many) && max_new > 1) {
    // decode steps leave their (all-gathered) logits in the peer-visible buffer: re-point the sampler there
    CK(cudaStreamSynchronize(st));
    float* cl = (float*)((char*)e->comm_buf + e->tpc.off_logits);
    const bool alias = a.u == a.l;
    a.l = cl;
    int sj = 1;
    a.u = alias ? cl : (a.u ? cl + (size_t)(sj++) * B * V : nullptr);
    a.d = a.d ? cl + (size_t)(sj++) * B * V : nullptr;
    memcpy(e->pinned, &a, sizeof a);
    CK(cudaMemcpyAsync(e->d_sargs, e->pinned, sizeof a, cudaMemcpyHostToDevice, st));
  }
  CK(cudaEventRecord(e->ev[2], st));

  int steps = 1;
  if (max_new > 1) {
    // everything the captured nodes bake in: shapes, kernel choices and the buffers their arguments point into
    uint64_t key = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { key = (key ^ v) * 1099511628211ull; };
    mix((uint64_t)B); mix((uint64_t)S); mix((uint64_t)e->R); mix((uint64_t)e->Mtot); mix((uint64_t)e->max_splits);
    mix((uint64_t)e->attn_items); mix((uint64_t)e->attn_chunk);
    mix((uint64_t)(e->use_pdl ? 1 : 0) | (uint64_t)(e->use_fused ? 2 : 0) | (uint64_t)(e->tp_fused_ready ? 8 : 0) |
        (uint64_t)((a.top_p < 1.0f || (a.top_k > 1 && a.top_k < V)) ? 4 : 0));  // bit 2: which K8 kernel the graph holds
    mix((uint64_t)(uintptr_t)e->kv.p); mix((uint64_t)(uintptr_t)e->logits.p); mix((uint64_t)e->kv_layer_stride);
    mix((uint64_t)(uintptr_t)e->meta.p); mix((uint64_t)(uintptr_t)e->dec_f32.p); mix((uint64_t)(uintptr_t)e->dec_bf16.p);
    mix((uint64_t)(uintptr_t)e->split_ws.p);
    // the meta carve-out places the sampler arguments, barriers and work table BEHIND the [B, max_new] sequence buffer, so
    // their addresses move with max_new even when every shape above is unchanged
    mix((uint64_t)(e->step_ready ? 1 + e->step_launch.nw * 16 + e->step_launch.depth : 0)); mix((uint64_t)(uintptr_t)e->step_tab.p);
    mix((uint64_t)e->max_new); mix((uint64_t)(uintptr_t)e->d_sargs); mix((uint64_t)(uintptr_t)e->d_bar);
    mix((uint64_t)(uintptr_t)e->d_item_row); mix((uint64_t)(uintptr_t)e->d_seq_out);
    int per_step_launches = 0;
    // the legacy default stream cannot be captured; a full attention capture adds per-layer launches: those steps run eagerly
    const bool graph = e->use_graph && st != nullptr && !e->attn_alln;
    if (graph && (!e->graph_exec || e->graph_key != key)) {
      if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
      cudaGraph_t g = nullptr;
      const int64_t before = e->launches;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      int rc = decode_step_enqueue(e, e->d_tokens, st);
      if (rc == VDD_OK) {
        cudaError_t ce = launch_combine_sample(a, e->d_sargs, B, V, st);
        e->launches++;
        if (ce == cudaSuccess) ce = launch_advance(e->d_pos, e->R, e->d_step, st, e->d_epoch);
        e->launches++;
        if (ce != cudaSuccess) rc = fail(VDD_ERR_CUDA, "capture: %s", cudaGetErrorString(ce));
      }
      cudaError_t ce = cudaStreamEndCapture(st, &g);
      if (rc != VDD_OK) { if (g) cudaGraphDestroy(g); return rc; }
      if (ce != cudaSuccess) return fail(VDD_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ce));
      e->graph_nodes = (int)(e->launches - before);
      e->launches = before;
      ce = cudaGraphInstantiate(&e->graph_exec, g, 0);
      cudaGraphDestroy(g);
      if (ce != cudaSuccess) { e->graph_exec = nullptr; return fail(VDD_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce)); }
      e->graph_key = key;
    }
    per_step_launches = e->graph_nodes;
    for (int i = 1; i < max_new; ++i) {
      if (graph) {
        CK(cudaGraphLaunch(e->graph_exec, st));
        e->launches += per_step_launches;
      } else {
        RC(decode_step_enqueue(e, e->d_tokens, st));
        LK(e, launch_combine_sample(a, e->d_sargs, B, V, st));
        LK(e, launch_advance(e->d_pos, e->R, e->d_step, st, e->d_epoch));
      }
      steps = i + 1;
      // all rows finished? (vcd_sample.py:291: one host sync per step there; here every 16 steps)
      if (p->eos_token_id >= 0 && (i % 16) == 15 && i + 1 < max_new) {
        CK(cudaMemcpyAsync(h_unf, e->d_unfinished, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        int any = 0;
        for (int b = 0; b < B; ++b) any |= h_unf[b];
        if (!any) break;
      }
    }
  }
  CK(cudaEventRecord(e->ev[3], st));
  CK(cudaMemcpy2DAsync(tokens_out_host, (size_t)max_new * 8, e->d_seq_out, (size_t)max_new * 8, (size_t)steps * 8, B,
                       cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  // trim to the step at which the reference would have left its loop (all rows finished)
  if (p->eos_token_id >= 0) {
    std::vector<int> unf(B, 1);
    for (int i = 0; i < steps; ++i) {
      int any = 0;
      for (int b = 0; b < B; ++b) {
        if (unf[b] && tokens_out_host[(size_t)b * max_new + i] == p->eos_token_id) unf[b] = 0;
        any |= unf[b];
      }
      if (!any) { steps = i + 1; break; }
    }
  }
  e->attn_alln = nullptr;  // the capture applies to one generate
  if (steps_out) *steps_out = steps;
  if (rng) e->rng_advance = (unsigned long long)steps * a.philox_stride;  // `steps` exponential_ calls, as the reference's loop makes
  cudaEventElapsedTime(&e->t_vision, e->ev[0], e->ev[1]);
  cudaEventElapsedTime(&e->t_prefill, e->ev[1], e->ev[2]);
  cudaEventElapsedTime(&e->t_decode, e->ev[2], e->ev[3]);
  return VDD_OK;
}

extern "C" int vdd_debug_trace(vdd_engine* e, unsigned long long* out_host, int n) {
  if (!e || !e->d_trace) return fail(VDD_ERR_STATE, "set VDD_TRACE=1 before vdd_create");
  CK(cudaMemcpy(out_host, e->d_trace, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return VDD_OK;
}
extern "C" int64_t vdd_launch_count(vdd_engine* e, int reset) {
  if (!e) return 0;
  const int64_t n = e->launches;
  if (reset) e->launches = 0;
  return n;
}
extern "C" int vdd_last_timings(vdd_engine* e, float* v, float* p, float* d) {
  if (!e) return VDD_ERR_ARG;
  if (v) *v = e->t_vision;
  if (p) *p = e->t_prefill;
  if (d) *d = e->t_decode;
  return VDD_OK;
}

// One decode step's projection launches only (qkv, o, gate_up, down per layer + lm_head), back to back exactly as in
// the real step: bench.py times this chain with CUDA events to get the GEMV kernel's average launch duration.
extern "C" int vdd_bench_projection_chain(vdd_engine* e, void* stream_v) {
  if (!e || !e->prefilled) return fail(VDD_ERR_STATE, "vdd_prefill must run first");
  return decode_step_enqueue(e, nullptr, (cudaStream_t)stream_v, true);
}

// ------------------------------------------------------------------------------------------------ TP
extern "C" int vdd_tp_unique_id(void* id128) {
  if (!nccl_load()) return fail(VDD_ERR_UNSUPPORTED, "libnccl.so.2 not loadable (set VDD_NCCL_LIB): %s", dlerror());
  int r = g_nccl.GetUniqueId(id128);
  if (r != 0) return fail(VDD_ERR_CUDA, "ncclGetUniqueId failed (%d)", r);
  return VDD_OK;
}
extern "C" int vdd_tp_init(vdd_engine* e, const void* id128) {
  if (!e || !id128) return fail(VDD_ERR_ARG, "null argument");
  if (e->c.tp_size <= 1) return VDD_OK;
  if (!nccl_load()) return fail(VDD_ERR_UNSUPPORTED, "libnccl.so.2 not loadable (set VDD_NCCL_LIB): %s", dlerror());
  NcclId nid;
  memcpy(nid.b, id128, 128);
  int r = g_nccl.CommInitRank(&e->comm, e->c.tp_size, nid, e->c.tp_rank);
  if (r != 0) return fail(VDD_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return VDD_OK;
}

// Peer-memory communication buffer for the in-kernel all-reduce (one per rank, exchanged as cudaIpc handles).
extern "C" int vdd_tp_comm_handle(vdd_engine* e, void* handle64) {
  if (!e || !handle64) return fail(VDD_ERR_ARG, "null argument");
  if (e->c.tp_size <= 1 || e->c.tp_size > 8) return fail(VDD_ERR_UNSUPPORTED, "peer communication supports 2..8 ranks");
  if (!e->comm_buf) {
    // [2 slots][tp source planes][rows][H] fp32 partial sums + [rows][V] fp32 gathered logits; rows = 8 for the GEMV / whole-step
    // kernels, up to VDD_TP_MAX_ROWS (default 256) for the tensor-core decode path of BASELINE configs[3..4]
    int mr = getenv("VDD_TP_MAX_ROWS") ? atoi(getenv("VDD_TP_MAX_ROWS")) : 256;
    if (mr < 8) mr = 8;
    e->tp_max_rows = mr;
    const size_t inbox = (size_t)2 * e->c.tp_size * mr * e->c.hidden_size * 4;
    const size_t logits = (size_t)mr * e->c.vocab_size * 4;
    e->tpc.off_inbox = 4096;
    e->tpc.off_logits = 4096 + (long long)align_up(inbox, 256);
    e->comm_bytes = (size_t)e->tpc.off_logits + logits;
    CK(cudaMalloc(&e->comm_buf, e->comm_bytes));
    CK(cudaMemset(e->comm_buf, 0, e->comm_bytes));
    CK(cudaMalloc(&e->d_epoch, 4));
    CK(cudaMemset(e->d_epoch, 0, 4));
  }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, e->comm_buf));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  memcpy(handle64, &h, 64);
  return VDD_OK;
}
extern "C" int vdd_tp_comm_open(vdd_engine* e, const void* handles) {
  if (!e || !handles || !e->comm_buf) return fail(VDD_ERR_STATE, "vdd_tp_comm_handle must run first");
  const int tp = e->c.tp_size;
  for (int k = 0; k < tp; ++k) {
    if (k == e->c.tp_rank) {
      e->tpc.peers[k] = (unsigned long long)(uintptr_t)e->comm_buf;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)k * 64, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    e->tpc.peers[k] = (unsigned long long)(uintptr_t)p;
  }
  e->tpc.tp = tp;
  e->tpc.rank = e->c.tp_rank;
  e->tpc.H = e->c.hidden_size;
  e->tpc.V = e->c.vocab_size;
  e->tpc.Vl = e->Vl;
  e->tpc.epoch = e->d_epoch;
  e->tp_fused_ready = getenv("VDD_TP_NCCL_ONLY") == nullptr;
  return VDD_OK;
}

// ------------------------------------------------------------------------------------------------ stand-alone kernel entries
extern "C" int vdd_gemv(const void* w, const void* x_bf16, const float* x_resid, const void* norm_w, float eps, int R, int N,
                        int K, int epilogue, float* out_f32, void* out_bf16, void* stream_v) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  GemvArgs g{};
  g.w = w; g.x_bf16 = x_bf16; g.x_resid = x_resid; g.norm_w = norm_w; g.eps = eps; g.R = R; g.N = N; g.K = K;
  g.epilogue = epilogue; g.out_f32 = out_f32; g.out_bf16 = out_bf16;
  g.ld_out = epilogue == GEMV_OUT_SWIGLU_BF16 ? N / 2 : N;
  cudaError_t ce = launch_gemv(g, sms, getenv("VDD_NO_PDL") == nullptr, (cudaStream_t)stream_v);
  if (ce != cudaSuccess) return fail(VDD_ERR_CUDA, "gemv R=%d N=%d K=%d: %s", R, N, K, cudaGetErrorString(ce));
  return VDD_OK;
}
extern "C" int vdd_clip_preprocess(const uint8_t* rgb, int H, int W, int rgb_on_device, int size, const uint8_t* pad_rgb, const float* mean3,
                                   const float* std3, void* out_dev, int out_dtype, void* stream_v) {
  if (!rgb || !mean3 || !std3 || !out_dev) return fail(VDD_ERR_ARG, "null argument");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
    return fail(VDD_ERR_CUDA, "libvdd needs a CUDA device (sm_100a); none visible");
  cudaError_t ce = launch_clip_preprocess(rgb, H, W, rgb_on_device != 0, size, pad_rgb, mean3, std3, out_dev, out_dtype, (cudaStream_t)stream_v);
  if (ce != cudaSuccess) return fail(VDD_ERR_CUDA, "clip_preprocess %dx%d -> %d: %s", H, W, size, cudaGetErrorString(ce));
  return VDD_OK;
}
extern "C" int vdd_gemm(const void* a, const void* w, int M, int N, int K, int epilogue, const void* bias, const int32_t* row_map,
                        float* out_f32, void* out_bf16, int64_t ldc, void* stream_v) {
  GemmArgs g{a, w, M, N, K, epilogue, bias, row_map, out_f32, out_bf16, ldc};
  cudaError_t ce = launch_gemm(g, (cudaStream_t)stream_v);
  if (ce != cudaSuccess)
    return fail(VDD_ERR_CUDA, "gemm M=%d N=%d K=%d: %s %s", M, N, K, cudaGetErrorString(ce), gemm_init_error() ? gemm_init_error() : "");
  return VDD_OK;
}
