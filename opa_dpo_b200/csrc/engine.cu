// Host orchestration of the OPA-DPO hot path (see engine.h). Every arithmetic step is a kernel from gemm.cu,
// attention.cu or elementwise.cu; nothing here computes on the CPU.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "attention.h"
#include "common.h"
#include "decode_gemm.h"
#include "gemm.h"

namespace opadpo {

// The C ABI entry points clear g_last_error on entry, so an empty message here means the callee set none.
#define EN_CHECK(expr)                                                                                 \
  do {                                                                                                 \
    int _rc = (expr);                                                                                  \
    if (_rc != 0) {                                                                                    \
      if (g_last_error[0] == 0) set_error("%s failed (%d) at %s:%d", #expr, _rc, __FILE__, __LINE__);  \
      return (_rc >= OPADPO_ENOENT) ? _rc : OPADPO_ECUDA;                                              \
    }                                                                                                  \
  } while (0)

static int gemm_checked(const GemmArgs& a, cudaStream_t s) {
  int rc = gemm_launch(a, s);
  if (rc != 0) set_error("%s", gemm_last_error());
  return rc;
}
static int attn_fwd_checked(const AttnArgs& a, cudaStream_t s) {
  int rc = attn_forward(a, s);
  if (rc != 0) set_error("%s", attn_last_error());
  return rc;
}
static int attn_bwd_checked(const AttnArgs& a, cudaStream_t s) {
  int rc = attn_backward(a, s);
  if (rc != 0) set_error("%s", attn_last_error());
  return rc;
}

// Stream-ordered temporaries: allocated from the CUDA memory pool, released at scope exit on the same stream.
struct TempPool {
  cudaStream_t s;
  std::vector<void*> ptrs;
  explicit TempPool(cudaStream_t st) : s(st) {}
  ~TempPool() {
    for (void* p : ptrs) cudaFreeAsync(p, s);
  }
};

template <typename T>
static int pool_alloc(T** p, int64_t count, cudaStream_t s, std::vector<void*>* track) {
  if (count <= 0) count = 1;
  void* q = nullptr;
  cudaError_t e = cudaMallocAsync(&q, static_cast<size_t>(count) * sizeof(T), s);
  if (e != cudaSuccess) {
    set_error("cudaMallocAsync(%lld bytes): %s", static_cast<long long>(count * sizeof(T)), cudaGetErrorString(e));
    cudaGetLastError();
    return OPADPO_ENOMEM;
  }
  track->push_back(q);
  *p = static_cast<T*>(q);
  return 0;
}

Engine::Engine(const opadpo_config& c) : cfg(c) {}

Engine::~Engine() {
  cudaDeviceSynchronize();
  for (int i = 0; i < 2; ++i) release_slot(i, nullptr);
  drop_frozen_merged();
  for (void* p : allocs) cudaFree(p);
  if (ev_u) cudaEventDestroy(ev_u);
  for (int k = 0; k < 4; ++k)
    if (ev_side[k]) cudaEventDestroy(ev_side[k]);
  if (wg_stream) cudaStreamDestroy(wg_stream);
  cudaDeviceSynchronize();
}

int Engine::ensure_wgrad_stream() {
  if (wg_stream) return 0;
  cudaError_t e = cudaStreamCreateWithFlags(&wg_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_u, cudaEventDisableTiming);
  for (int k = 0; k < 4 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ev_side[k], cudaEventDisableTiming);
  if (e != cudaSuccess) {
    set_error("wgrad side stream: %s", cudaGetErrorString(e));
    return OPADPO_ECUDA;
  }
  return 0;
}

void Engine::wait_side(int kind, cudaStream_t s) {
  if (kind >= 0 && side_pending[kind]) {
    // a failure here would surface as a wrong gradient, not as an error: remember it for the end-of-backward check
    if (cudaStreamWaitEvent(s, ev_side[kind], 0) != cudaSuccess) side_error = true;
    side_pending[kind] = false;
  }
}

template <typename T>
int Engine::dalloc(T** p, int64_t count, std::vector<void*>* track) {
  void* q = nullptr;
  if (count <= 0) count = 1;
  cudaError_t e = cudaMalloc(&q, static_cast<size_t>(count) * sizeof(T));
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%lld bytes): %s", static_cast<long long>(count * sizeof(T)), cudaGetErrorString(e));
    cudaGetLastError();
    return OPADPO_ENOMEM;
  }
  cudaMemset(q, 0, static_cast<size_t>(count) * sizeof(T));
  track->push_back(q);
  *p = static_cast<T*>(q);
  return 0;
}

int Engine::alloc_linear(LinearW& L, int in, int out, int groups, bool bias, bool lora) {
  L.in = in;
  L.out = out;
  L.groups = groups;
  EN_CHECK(dalloc(&L.W, static_cast<int64_t>(in) * out, &allocs));
  if (bias) EN_CHECK(dalloc(&L.bias, out, &allocs));
  if (lora && r > 0) {
    LoraSlot sl;
    sl.in = in;
    sl.out = out;
    sl.groups = groups;
    sl.offA = n_total;
    n_total += static_cast<int64_t>(groups) * r * in;
    sl.offB = n_total;
    n_total += static_cast<int64_t>(out) * r;
    L.slot = static_cast<int>(slots.size());
    slots.push_back(sl);
  }
  return 0;
}

int Engine::init() {
  const opadpo_config& c = cfg;
  if (c.d_model <= 0 || c.n_heads <= 0 || c.d_model % c.n_heads != 0) {
    set_error("config: d_model %d / n_heads %d", c.d_model, c.n_heads);
    return OPADPO_EINVAL;
  }
  hd = c.d_model / c.n_heads;
  v_hd = c.v_heads > 0 ? c.v_d / c.v_heads : 0;
  if (hd != 128) {
    set_error("config: LLaMA head_dim %d unsupported (128)", hd);
    return OPADPO_EINVAL;
  }
  if (v_hd != 64 && v_hd != 128) {
    set_error("config: CLIP head_dim %d unsupported (64 or 128)", v_hd);
    return OPADPO_EINVAL;
  }
  if (c.d_model % 8 || c.d_ffn % 8 || c.v_d % 8 || c.v_ffn % 8 || c.vocab % 8 || c.lora_r % 8) {
    set_error("config: d_model, d_ffn, vocab, v_d, v_ffn and lora_r must be multiples of 8");
    return OPADPO_EINVAL;
  }
  if (c.v_image % c.v_patch != 0) {
    set_error("config: image %d not a multiple of patch %d", c.v_image, c.v_patch);
    return OPADPO_EINVAL;
  }
  const int g = c.v_image / c.v_patch;
  P = g * g;
  Kp = ((3 * c.v_patch * c.v_patch + 7) / 8) * 8;
  r = c.lora_r;
  lora_scale = r > 0 ? c.lora_alpha / static_cast<float>(r) : 0.f;
  const int d = c.d_model, F = c.d_ffn;

  EN_CHECK(dalloc(&embed, static_cast<int64_t>(c.vocab) * d, &allocs));
  EN_CHECK(dalloc(&final_norm, d, &allocs));
  EN_CHECK(dalloc(&lm_head, static_cast<int64_t>(c.vocab) * d, &allocs));
  layers.resize(c.n_layers);
  for (auto& L : layers) {
    EN_CHECK(alloc_linear(L.qkv, d, 3 * d, 3, false, true));
    EN_CHECK(alloc_linear(L.o, d, d, 1, false, true));
    EN_CHECK(alloc_linear(L.gu, d, 2 * F, 2, false, true));
    EN_CHECK(alloc_linear(L.down, F, d, 1, false, true));
    EN_CHECK(dalloc(&L.ln1, d, &allocs));
    EN_CHECK(dalloc(&L.ln2, d, &allocs));
  }
  EN_CHECK(alloc_linear(proj0, c.v_d, d, 1, true, true));
  EN_CHECK(alloc_linear(proj2, d, d, 1, true, true));
  n_train = n_total;  // the CLIP tower runs under no_grad: its LoRA tensors never receive gradients
  EN_CHECK(dalloc(&v_patch_w, static_cast<int64_t>(c.v_d) * Kp, &allocs));
  EN_CHECK(dalloc(&v_cls, c.v_d, &allocs));
  EN_CHECK(dalloc(&v_pos, static_cast<int64_t>(P + 1) * c.v_d, &allocs));
  EN_CHECK(dalloc(&v_pre_w, c.v_d, &allocs));
  EN_CHECK(dalloc(&v_pre_b, c.v_d, &allocs));
  vlayers.resize(c.v_layers);
  for (auto& L : vlayers) {
    EN_CHECK(alloc_linear(L.qkv, c.v_d, 3 * c.v_d, 3, true, true));
    EN_CHECK(alloc_linear(L.out, c.v_d, c.v_d, 1, true, true));
    EN_CHECK(alloc_linear(L.fc1, c.v_d, c.v_ffn, 1, true, true));
    EN_CHECK(alloc_linear(L.fc2, c.v_ffn, c.v_d, 1, true, true));
    EN_CHECK(dalloc(&L.ln1w, c.v_d, &allocs));
    EN_CHECK(dalloc(&L.ln1b, c.v_d, &allocs));
    EN_CHECK(dalloc(&L.ln2w, c.v_d, &allocs));
    EN_CHECK(dalloc(&L.ln2b, c.v_d, &allocs));
  }
  for (int a = 0; a < 2; ++a) EN_CHECK(dalloc(&adapters[a].shadow, n_total, &allocs));
  EN_CHECK(dalloc(&sumsq, 4, &allocs));
  EN_CHECK(dalloc(&rope_cs, static_cast<int64_t>(c.max_positions) * (hd / 2), &allocs));
  EN_CHECK(rope_table(rope_cs, c.max_positions, hd, c.rope_theta, nullptr));
  // keep freed temporaries in the pool instead of returning them to the OS
  int dev = 0;
  cudaGetDevice(&dev);
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("engine init: %s", cudaGetErrorString(e));
    return OPADPO_ECUDA;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ name resolution
static bool ends_with(const std::string& s, const char* suf) {
  const size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}
static int parse_layer(const std::string& s, const char* key) {
  const size_t p = s.find(key);
  if (p == std::string::npos) return -1;
  return atoi(s.c_str() + p + strlen(key));
}

// Resolves a module path ("...layers.3.self_attn.q_proj", "...mm_projector.0", CLIP "...encoder.layers.5.mlp.fc1").
bool Engine::resolve_module(const char* name, ModuleRef* ref) {
  const std::string s(name);
  if (s.find("mm_projector.0") != std::string::npos) {
    *ref = {&proj0, 0};
    return true;
  }
  if (s.find("mm_projector.2") != std::string::npos) {
    *ref = {&proj2, 0};
    return true;
  }
  if (s.find("encoder.layers.") != std::string::npos) {
    const int l = parse_layer(s, "encoder.layers.");
    if (l < 0 || l >= static_cast<int>(vlayers.size())) return false;
    ClipLayer& L = vlayers[l];
    if (ends_with(s, "self_attn.q_proj")) { *ref = {&L.qkv, 0}; return true; }
    if (ends_with(s, "self_attn.k_proj")) { *ref = {&L.qkv, 1}; return true; }
    if (ends_with(s, "self_attn.v_proj")) { *ref = {&L.qkv, 2}; return true; }
    if (ends_with(s, "self_attn.out_proj")) { *ref = {&L.out, 0}; return true; }
    if (ends_with(s, "mlp.fc1")) { *ref = {&L.fc1, 0}; return true; }
    if (ends_with(s, "mlp.fc2")) { *ref = {&L.fc2, 0}; return true; }
    return false;
  }
  if (s.find("layers.") != std::string::npos) {
    const int l = parse_layer(s, "layers.");
    if (l < 0 || l >= static_cast<int>(layers.size())) return false;
    LlamaLayer& L = layers[l];
    if (ends_with(s, "self_attn.q_proj")) { *ref = {&L.qkv, 0}; return true; }
    if (ends_with(s, "self_attn.k_proj")) { *ref = {&L.qkv, 1}; return true; }
    if (ends_with(s, "self_attn.v_proj")) { *ref = {&L.qkv, 2}; return true; }
    if (ends_with(s, "self_attn.o_proj")) { *ref = {&L.o, 0}; return true; }
    if (ends_with(s, "mlp.gate_proj")) { *ref = {&L.gu, 0}; return true; }
    if (ends_with(s, "mlp.up_proj")) { *ref = {&L.gu, 1}; return true; }
    if (ends_with(s, "mlp.down_proj")) { *ref = {&L.down, 0}; return true; }
    return false;
  }
  return false;
}

int Engine::copy_in(bf16* dst, const void* src, int dtype, int64_t n, cudaStream_t s) {
  if (dtype == OPADPO_DT_BF16) {
    cudaError_t e = cudaMemcpyAsync(dst, src, static_cast<size_t>(n) * 2, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) {
      set_error("copy_in: %s", cudaGetErrorString(e));
      return OPADPO_ECUDA;
    }
    return 0;
  }
  if (dtype == OPADPO_DT_F32) return f32_to_bf16(static_cast<const float*>(src), dst, n, s);
  set_error("unknown dtype %d", dtype);
  return OPADPO_EINVAL;
}

int Engine::load_tensor(const char* name, const void* src, int dtype, int64_t rows, int64_t cols, cudaStream_t s) {
  const std::string n(name);
  const opadpo_config& c = cfg;
  drop_frozen_merged();  // base weights change under any merged copy
  auto expect = [&](int64_t er, int64_t ec) -> int {
    if (rows != er || cols != ec) {
      set_error("load_tensor(%s): shape [%lld,%lld], expected [%lld,%lld]", name, (long long)rows, (long long)cols,
                (long long)er, (long long)ec);
      return OPADPO_EINVAL;
    }
    return 0;
  };
  const bool is_w = ends_with(n, ".weight"), is_b = ends_with(n, ".bias");
  // ---- CLIP tower
  if (n.find("vision_model.") != std::string::npos || n.find("vision_tower.") != std::string::npos) {
    if (ends_with(n, "embeddings.class_embedding")) {
      EN_CHECK(expect(1, c.v_d));
      return copy_in(v_cls, src, dtype, c.v_d, s);
    }
    if (ends_with(n, "embeddings.patch_embedding.weight")) {
      const int k = 3 * c.v_patch * c.v_patch;
      EN_CHECK(expect(c.v_d, k));
      if (k == Kp) return copy_in(v_patch_w, src, dtype, static_cast<int64_t>(c.v_d) * k, s);
      // repack rows to the padded pitch Kp (pad columns stay zero)
      TempPool tp(s);
      bf16* tmp = nullptr;
      EN_CHECK(pool_alloc(&tmp, static_cast<int64_t>(c.v_d) * k, s, &tp.ptrs));
      EN_CHECK(copy_in(tmp, src, dtype, static_cast<int64_t>(c.v_d) * k, s));
      cudaError_t e = cudaMemcpy2DAsync(v_patch_w, Kp * 2, tmp, k * 2, k * 2, c.v_d, cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) {
        set_error("patch repack: %s", cudaGetErrorString(e));
        return OPADPO_ECUDA;
      }
      return 0;
    }
    if (ends_with(n, "embeddings.position_embedding.weight")) {
      EN_CHECK(expect(P + 1, c.v_d));
      return copy_in(v_pos, src, dtype, static_cast<int64_t>(P + 1) * c.v_d, s);
    }
    if (n.find("pre_layrnorm.") != std::string::npos) {
      EN_CHECK(expect(1, c.v_d));
      return copy_in(is_w ? v_pre_w : v_pre_b, src, dtype, c.v_d, s);
    }
    if (n.find("encoder.layers.") != std::string::npos) {
      const int l = parse_layer(n, "encoder.layers.");
      if (l < 0) {
        set_error("load_tensor(%s): no layer index", name);
        return OPADPO_EINVAL;
      }
      if (l >= static_cast<int>(vlayers.size())) return 0;  // layers past the selected one never run
      ClipLayer& L = vlayers[l];
      if (n.find("layer_norm1.") != std::string::npos) { EN_CHECK(expect(1, c.v_d)); return copy_in(is_w ? L.ln1w : L.ln1b, src, dtype, c.v_d, s); }
      if (n.find("layer_norm2.") != std::string::npos) { EN_CHECK(expect(1, c.v_d)); return copy_in(is_w ? L.ln2w : L.ln2b, src, dtype, c.v_d, s); }
    }
    if (n.find("post_layernorm") != std::string::npos) return 0;  // unused by feature_select
  }
  // ---- linear modules (LLaMA, CLIP, projector)
  if (is_w || is_b) {
    const std::string mod = n.substr(0, n.size() - (is_w ? 7 : 5));
    ModuleRef ref;
    if (resolve_module(mod.c_str(), &ref)) {
      LinearW& L = *ref.lin;
      const int out_g = L.out / L.groups;
      if (is_w) {
        EN_CHECK(expect(out_g, L.in));
        return copy_in(L.W + static_cast<int64_t>(ref.group) * out_g * L.in, src, dtype, static_cast<int64_t>(out_g) * L.in, s);
      }
      if (!L.bias) {
        set_error("load_tensor(%s): module has no bias", name);
        return OPADPO_EINVAL;
      }
      EN_CHECK(expect(1, out_g));
      return copy_in(L.bias + static_cast<int64_t>(ref.group) * out_g, src, dtype, out_g, s);
    }
  }
  if (ends_with(n, "embed_tokens.weight")) {
    EN_CHECK(expect(c.vocab, c.d_model));
    return copy_in(embed, src, dtype, static_cast<int64_t>(c.vocab) * c.d_model, s);
  }
  if (ends_with(n, "lm_head.weight")) {
    EN_CHECK(expect(c.vocab, c.d_model));
    return copy_in(lm_head, src, dtype, static_cast<int64_t>(c.vocab) * c.d_model, s);
  }
  if (n.find("layers.") != std::string::npos && n.find("encoder.") == std::string::npos) {
    const int l = parse_layer(n, "layers.");
    if (l >= 0 && l < static_cast<int>(layers.size())) {
      if (ends_with(n, "input_layernorm.weight")) { EN_CHECK(expect(1, c.d_model)); return copy_in(layers[l].ln1, src, dtype, c.d_model, s); }
      if (ends_with(n, "post_attention_layernorm.weight")) { EN_CHECK(expect(1, c.d_model)); return copy_in(layers[l].ln2, src, dtype, c.d_model, s); }
    }
    if (ends_with(n, "rotary_emb.inv_freq")) return 0;
  }
  if (ends_with(n, "model.norm.weight") || n == "norm.weight") {
    EN_CHECK(expect(1, c.d_model));
    return copy_in(final_norm, src, dtype, c.d_model, s);
  }
  set_error("load_tensor: unknown tensor name '%s'", name);
  return OPADPO_ENOENT;
}

int Engine::lora_shape(const char* module, int64_t* in_f, int64_t* out_f) {
  ModuleRef ref;
  if (!resolve_module(module, &ref) || ref.lin->slot < 0) {
    set_error("'%s' is not a LoRA target of this engine", module);
    return OPADPO_EINVAL;
  }
  *in_f = ref.lin->in;
  *out_f = ref.lin->out / ref.lin->groups;
  return 0;
}

int Engine::lora_offsets(const char* module, int64_t* offA, int64_t* offB) {
  ModuleRef ref;
  if (!resolve_module(module, &ref) || ref.lin->slot < 0) {
    set_error("'%s' is not a LoRA target of this engine", module);
    return OPADPO_EINVAL;
  }
  const LoraSlot& sl = slots[ref.lin->slot];
  *offA = sl.offA + static_cast<int64_t>(ref.group) * r * sl.in;
  *offB = sl.offB + static_cast<int64_t>(ref.group) * (sl.out / sl.groups) * r;
  return 0;
}

int Engine::load_lora(int adapter, const char* module, const void* A, const void* B, int dtype, cudaStream_t s) {
  if (adapter < 0 || adapter > 1) {
    set_error("adapter %d out of range", adapter);
    return OPADPO_EINVAL;
  }
  int64_t oa, ob, in_f, out_f;
  EN_CHECK(lora_offsets(module, &oa, &ob));
  EN_CHECK(lora_shape(module, &in_f, &out_f));
  Adapter& ad = adapters[adapter];
  if (frozen_merged_adapter == adapter) drop_frozen_merged();  // the merged copy is stale
  EN_CHECK(copy_in(ad.shadow + oa, A, dtype, static_cast<int64_t>(r) * in_f, s));
  EN_CHECK(copy_in(ad.shadow + ob, B, dtype, out_f * r, s));
  if (ad.trainable) {
    if (dtype == OPADPO_DT_F32) {
      cudaMemcpyAsync(ad.master + oa, A, static_cast<size_t>(r) * in_f * 4, cudaMemcpyDeviceToDevice, s);
      cudaMemcpyAsync(ad.master + ob, B, static_cast<size_t>(out_f) * r * 4, cudaMemcpyDeviceToDevice, s);
    } else {
      EN_CHECK(bf16_to_f32(ad.shadow + oa, ad.master + oa, static_cast<int64_t>(r) * in_f, s));
      EN_CHECK(bf16_to_f32(ad.shadow + ob, ad.master + ob, out_f * r, s));
    }
  }
  return 0;
}

int Engine::export_lora(int adapter, const char* module, float* A_out, float* B_out, cudaStream_t s) {
  if (adapter < 0 || adapter > 1) {
    set_error("adapter %d out of range", adapter);
    return OPADPO_EINVAL;
  }
  int64_t oa, ob, in_f, out_f;
  EN_CHECK(lora_offsets(module, &oa, &ob));
  EN_CHECK(lora_shape(module, &in_f, &out_f));
  Adapter& ad = adapters[adapter];
  if (ad.trainable) {
    cudaMemcpyAsync(A_out, ad.master + oa, static_cast<size_t>(r) * in_f * 4, cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(B_out, ad.master + ob, static_cast<size_t>(out_f) * r * 4, cudaMemcpyDeviceToDevice, s);
  } else {
    EN_CHECK(bf16_to_f32(ad.shadow + oa, A_out, static_cast<int64_t>(r) * in_f, s));
    EN_CHECK(bf16_to_f32(ad.shadow + ob, B_out, out_f * r, s));
  }
  return 0;
}

int Engine::set_trainable(int adapter) {
  if (adapter < 0 || adapter > 1) {
    set_error("adapter %d out of range", adapter);
    return OPADPO_EINVAL;
  }
  Adapter& ad = adapters[adapter];
  if (ad.trainable) return 0;
  if (frozen_merged_adapter == adapter) drop_frozen_merged();
  if (trainable_adapter >= 0) {
    set_error("adapter %d is already the trainable one", trainable_adapter);
    return OPADPO_ESTATE;
  }
  EN_CHECK(dalloc(&ad.master, n_total, &allocs));
  EN_CHECK(dalloc(&ad.grad, n_train, &allocs));
  EN_CHECK(dalloc(&ad.m, n_train, &allocs));
  EN_CHECK(dalloc(&ad.v, n_train, &allocs));
  EN_CHECK(bf16_to_f32(ad.shadow, ad.master, n_total, nullptr));
  cudaDeviceSynchronize();
  ad.trainable = true;
  trainable_adapter = adapter;
  return 0;
}

int Engine::buffers(float** master, float** grad, float** m, float** v, int64_t* nt, int64_t* ntr) {
  if (trainable_adapter < 0) {
    set_error("no trainable adapter (call opadpo_engine_set_trainable first)");
    return OPADPO_ESTATE;
  }
  Adapter& ad = adapters[trainable_adapter];
  *master = ad.master;
  *grad = ad.grad;
  *m = ad.m;
  *v = ad.v;
  *nt = n_total;
  *ntr = n_train;
  return 0;
}

int Engine::sync_shadow(cudaStream_t s) {
  if (trainable_adapter < 0) return 0;
  Adapter& ad = adapters[trainable_adapter];
  return f32_to_bf16(ad.master, ad.shadow, n_total, s);
}

void Engine::release_slot(int slot, cudaStream_t s) {
  SaveCtx& c = ctx[slot];
  for (void* p : c.owned) cudaFreeAsync(p, s);
  c.owned.clear();
  c.layers.clear();
  c.valid = false;
}

// ------------------------------------------------------------------------------------------------ linear layers
// y = act(x W^T + bias + s*(x A^T) B^T) + residual. The LoRA-A product t = s*x*A^T is a skinny pre-GEMM whose
// output is consumed as a second K segment of the base GEMM ([x | t] [W ; B]^T into one TMEM accumulator).
int Engine::linear_fwd(const LinearW& L, int adapter, const bf16* x, int64_t ldx, int64_t M, bf16* t, bf16* y, int64_t ldy,
                       int act, const bf16* residual, int64_t ldr, cudaStream_t s, RopeFuse* rope) {
  const bool lora = adapter >= 0 && L.slot >= 0 && r > 0;
  const int G = L.groups, out_g = L.out / G;
  if (lora) {
    GemmArgs a;
    a.M = static_cast<int>(M);
    a.N = G * r;
    a.K1 = L.in;
    a.A1 = {x, ldx};
    a.B1 = {loraA(adapter, L), L.in};
    a.alpha = lora_scale;
    a.C = t;
    a.ldc = G * r;
    EN_CHECK(gemm_checked(a, s));
  }
  const bool fused = G == 1 || (out_g % 256 == 0);
  const int ncalls = fused ? 1 : G;
  for (int g = 0; g < ncalls; ++g) {
    GemmArgs a;
    a.M = static_cast<int>(M);
    a.N = fused ? L.out : out_g;
    a.K1 = L.in;
    a.A1 = {x, ldx};
    a.B1 = {L.W + static_cast<int64_t>(g) * out_g * L.in, L.in};
    if (lora) {
      a.K2 = r;
      a.A2 = {t + static_cast<int64_t>(g) * r, static_cast<int64_t>(G) * r};
      a.B2 = {loraB(adapter, L) + static_cast<int64_t>(g) * out_g * r, r};
      if (fused && G > 1) {
        a.a2_group_n = out_g;
        a.a2_group_stride = r;
      }
    }
    a.C = y + static_cast<int64_t>(g) * out_g;
    a.ldc = ldy;
    if (rope && fused && !L.bias && act == ACT_NONE && !residual && hd == 128 && L.out % 128 == 0 && rope->cols % 128 == 0) {
      a.epilogue = EPI_BF16_ROPE;
      a.rope_pos = rope->pos;
      a.rope_cs = rope_cs;
      a.rope_cols = rope->cols;
      a.rope_S = rope->S;
      rope->applied = true;
    }
    a.bias = L.bias ? L.bias + static_cast<int64_t>(g) * out_g : nullptr;
    a.act = act;
    if (residual) {
      a.residual = residual + static_cast<int64_t>(g) * out_g;
      a.ldr = ldr;
    }
    EN_CHECK(gemm_checked(a, s));
  }
  return 0;
}

static int pick_splits(int Mo, int No, int64_t Ktok) {
  const int tiles = ((Mo + 127) / 128) * ((No + 255) / 256);
  const int kb = static_cast<int>((Ktok + 63) / 64);
  int best = 1;
  double best_cost = 1e30;
  for (int sp = 1; sp <= 16; ++sp) {
    if (sp > 1 && kb / sp < 4) break;
    const int waves = (tiles * sp + 147) / 148;
    const double cost = static_cast<double>(waves) / sp + 0.02 * sp;  // small penalty: more atomics traffic
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = sp;
    }
  }
  return best;
}

// Backward of linear_fwd for a frozen W and trainable (A, B):
//   u  = s * dy B            (per group)          dB += dy^T t      (t = s x A^T saved by the forward)
//   dA += u^T x                                   dx  = [dy | u] [W ; A]
int Engine::linear_bwd(const LinearW& L, int adapter, const bf16* x, int64_t ldx, const bf16* t, const bf16* dy, int64_t ldy,
                       int64_t M, bf16* u, bf16* dx, int64_t lddx, const bf16* dres, int64_t lddres, cudaStream_t s, int side_kind,
                       const bf16* swiglu_gu) {
  const int G = L.groups, out_g = L.out / G;
  const bool lora = L.slot >= 0 && r > 0;
  Adapter& ad = adapters[adapter];
  if (lora) {
    const LoraSlot& sl = slots[L.slot];
    float* gA = ad.grad + sl.offA;
    float* gB = ad.grad + sl.offB;
    // side_kind >= 0: the weight gradients go to the side stream. `u` then belongs to this linear kind alone, and the caller
    // has made the main stream wait for the previous reader of it (wait_side) before getting here.
    const bool side = side_kind >= 0 && wgrad_stream && wg_stream != nullptr;
    cudaStream_t ws = side ? wg_stream : s;
    for (int g = 0; g < G; ++g) {
      GemmArgs a;  // u_g = s * dy_g B_g
      a.M = static_cast<int>(M);
      a.N = r;
      a.K1 = out_g;
      a.A1 = {dy + static_cast<int64_t>(g) * out_g, ldy};
      a.b_mn_major = true;
      a.B1 = {loraB(adapter, L) + static_cast<int64_t>(g) * out_g * r, r};
      a.alpha = lora_scale;
      a.C = u + static_cast<int64_t>(g) * r;
      a.ldc = static_cast<int64_t>(G) * r;
      EN_CHECK(gemm_checked(a, s));
    }
    if (side) {  // everything the weight gradients read (dy, t, x, u) is complete in main-stream order at this point
      if (cudaEventRecord(ev_u, s) != cudaSuccess || cudaStreamWaitEvent(ws, ev_u, 0) != cudaSuccess) side_error = true;
    }
    for (int g = 0; g < G; ++g) {
      GemmArgs w;  // dB_g += dy_g^T t_g
      w.M = out_g;
      w.N = r;
      w.K1 = static_cast<int>(M);
      w.a_mn_major = w.b_mn_major = true;
      w.A1 = {dy + static_cast<int64_t>(g) * out_g, ldy};
      w.B1 = {t + static_cast<int64_t>(g) * r, static_cast<int64_t>(G) * r};
      w.epilogue = EPI_F32_ATOMIC;
      w.splits = pick_splits(out_g, r, M);
      w.C = gB + static_cast<int64_t>(g) * out_g * r;
      w.ldc = r;
      EN_CHECK(gemm_checked(w, ws));
    }
    GemmArgs w;  // dA += u^T x
    w.M = G * r;
    w.N = L.in;
    w.K1 = static_cast<int>(M);
    w.a_mn_major = w.b_mn_major = true;
    w.A1 = {u, static_cast<int64_t>(G) * r};
    w.B1 = {x, ldx};
    w.epilogue = EPI_F32_ATOMIC;
    w.splits = pick_splits(G * r, L.in, M);
    w.C = gA;
    w.ldc = L.in;
    EN_CHECK(gemm_checked(w, ws));
    if (side) {
      if (cudaEventRecord(ev_side[side_kind], ws) != cudaSuccess) side_error = true;
      side_pending[side_kind] = true;
    }
  }
  if (dx) {
    GemmArgs a;
    a.M = static_cast<int>(M);
    a.N = L.in;
    a.K1 = L.out;
    a.A1 = {dy, ldy};
    a.b_mn_major = true;
    a.B1 = {L.W, L.in};
    if (lora) {
      a.K2 = G * r;
      a.A2 = {u, static_cast<int64_t>(G) * r};
      a.B2 = {loraA(adapter, L), L.in};
    }
    a.C = dx;
    a.ldc = lddx;
    if (dres) {
      a.residual = dres;
      a.ldr = lddres;
    }
    if (swiglu_gu) {  // this is the down projection: dx is d(gate|up) [M, 2 in], produced from the d(act) tile in the epilogue
      a.epilogue = EPI_SWIGLU_BWD;
      a.residual = swiglu_gu;
      a.ldr = 2 * static_cast<int64_t>(L.in);
    }
    EN_CHECK(gemm_checked(a, s));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ CLIP + projector
// CLIPVisionTower.forward under no_grad: hidden_states[select_layer][:, 1:] (llava_modifications.patch:15185-15217)
int Engine::clip_forward(int adapter, const float* images, int n, bf16* feats_out, cudaStream_t s, std::vector<void*>* tmp) {
  const opadpo_config& c = cfg;
  const int dv = c.v_d, Sv = P + 1;
  const int64_t Tv = static_cast<int64_t>(n) * Sv;
  bf16 *patches, *pe, *x, *x2, *h, *qkv, *o, *f, *t;
  EN_CHECK(pool_alloc(&patches, static_cast<int64_t>(n) * P * Kp, s, tmp));
  EN_CHECK(pool_alloc(&pe, static_cast<int64_t>(n) * P * dv, s, tmp));
  EN_CHECK(pool_alloc(&x, Tv * dv, s, tmp));
  EN_CHECK(pool_alloc(&x2, Tv * dv, s, tmp));
  EN_CHECK(pool_alloc(&h, Tv * dv, s, tmp));
  EN_CHECK(pool_alloc(&qkv, Tv * 3 * dv, s, tmp));
  EN_CHECK(pool_alloc(&o, Tv * dv, s, tmp));
  EN_CHECK(pool_alloc(&f, Tv * c.v_ffn, s, tmp));
  EN_CHECK(pool_alloc(&t, Tv * 3 * (r > 0 ? r : 8), s, tmp));
  EN_CHECK(clip_im2col(images, patches, n, c.v_image, c.v_patch, Kp, s));
  {
    GemmArgs a;  // patch embedding conv (no bias) as a GEMM over the im2col matrix
    a.M = n * P;
    a.N = dv;
    a.K1 = Kp;
    a.A1 = {patches, Kp};
    a.B1 = {v_patch_w, Kp};
    a.C = pe;
    a.ldc = dv;
    EN_CHECK(gemm_checked(a, s));
  }
  EN_CHECK(clip_assemble(pe, v_cls, v_pos, h, n, P, dv, s));
  EN_CHECK(layernorm_fwd(h, v_pre_w, v_pre_b, x, Tv, dv, c.v_eps, s));
  const float scale = 1.f / sqrtf(static_cast<float>(v_hd));
  for (size_t l = 0; l < vlayers.size(); ++l) {
    ClipLayer& L = vlayers[l];
    EN_CHECK(layernorm_fwd(x, L.ln1w, L.ln1b, h, Tv, dv, c.v_eps, s));
    EN_CHECK(linear_fwd(L.qkv, adapter, h, dv, Tv, t, qkv, 3 * dv, ACT_NONE, nullptr, 0, s));
    AttnArgs at;
    at.N = n;
    at.S = Sv;
    at.H = c.v_heads;
    at.HD = v_hd;
    at.causal = false;
    at.scale = scale;
    at.q = qkv;
    at.k = qkv + dv;
    at.v = qkv + 2 * dv;
    at.ld_qkv = 3 * dv;
    at.o = o;
    at.ld_o = dv;
    EN_CHECK(attn_fwd_checked(at, s));
    EN_CHECK(linear_fwd(L.out, adapter, o, dv, Tv, t, x2, dv, ACT_NONE, x, dv, s));
    EN_CHECK(layernorm_fwd(x2, L.ln2w, L.ln2b, h, Tv, dv, c.v_eps, s));
    EN_CHECK(linear_fwd(L.fc1, adapter, h, dv, Tv, t, f, c.v_ffn, ACT_QUICK_GELU, nullptr, 0, s));
    EN_CHECK(linear_fwd(L.fc2, adapter, f, c.v_ffn, Tv, t, x, dv, ACT_NONE, x2, dv, s));
  }
  EN_CHECK(clip_drop_cls(x, feats_out, n, P, dv, s));
  return 0;
}

// mm_projector mlp2x_gelu (llava_modifications.patch:15354-15367)
int Engine::projector_forward(int adapter, const bf16* clip_feats, int64_t rows, bf16* pre, bf16* act, bf16* t0, bf16* t2,
                              bf16* feats, cudaStream_t s) {
  const int d = cfg.d_model;
  EN_CHECK(linear_fwd(proj0, adapter, clip_feats, cfg.v_d, rows, t0, pre, d, ACT_NONE, nullptr, 0, s));
  EN_CHECK(gelu_fwd(pre, act, rows * d, s));
  EN_CHECK(linear_fwd(proj2, adapter, act, d, rows, t2, feats, d, ACT_NONE, nullptr, 0, s));
  return 0;
}

// ------------------------------------------------------------------------------------------------ LLaMA layer
int Engine::llama_layer_fwd(const LlamaLayer& L, int adapter, const SeqLayout& lay, const LayerActs& a, bf16* x_out,
                            cudaStream_t s) {
  const int64_t T = static_cast<int64_t>(lay.N) * lay.S;
  const int N = lay.N, S = lay.S;
  const opadpo_config& c = cfg;
  const int d = c.d_model, F = c.d_ffn;
  EN_CHECK(rmsnorm_fwd(a.x, L.ln1, a.h1, T, d, c.rms_eps, nullptr, s));
  RopeFuse rf;
  rf.pos = lay.pos;
  rf.S = S;
  rf.cols = 2 * d;
  EN_CHECK(linear_fwd(L.qkv, adapter, a.h1, d, T, a.tqkv, a.qkv, 3 * d, ACT_NONE, nullptr, 0, s, fuse_rope ? &rf : nullptr));
  if (!rf.applied) EN_CHECK(rope_inplace(a.qkv, 3 * d, T, S, lay.pos, 0, nullptr, 2 * c.n_heads, hd, rope_cs, 0, s));
  AttnArgs at;
  at.N = N;
  at.S = S;
  at.H = c.n_heads;
  at.HD = hd;
  at.causal = true;
  at.scale = 1.f / sqrtf(static_cast<float>(hd));
  at.q = a.qkv;
  at.k = a.qkv + d;
  at.v = a.qkv + 2 * d;
  at.ld_qkv = 3 * d;
  at.o = a.o;
  at.ld_o = d;
  at.lse = a.lse;
  at.kbits = lay.kbits;
  at.prefix_len = lay.S0;
  at.segment_len = lay.segR;
  EN_CHECK(attn_fwd_checked(at, s));
  EN_CHECK(linear_fwd(L.o, adapter, a.o, d, T, a.to, a.x1, d, ACT_NONE, a.x, d, s));
  EN_CHECK(rmsnorm_fwd(a.x1, L.ln2, a.h2, T, d, c.rms_eps, nullptr, s));
  if (adapter < 0 && L.gu.W_il) {  // merged frozen adapter: SwiGLU inside the projection's epilogue, `gu` never materialised
    GemmArgs g;
    g.M = static_cast<int>(T);
    g.N = 2 * F;
    g.K1 = d;
    g.A1 = {a.h2, d};
    g.B1 = {L.gu.W_il, d};
    g.epilogue = EPI_SWIGLU_BF16;
    g.C = a.act;
    g.ldc = F;
    EN_CHECK(gemm_checked(g, s));
  } else {
    EN_CHECK(linear_fwd(L.gu, adapter, a.h2, d, T, a.tgu, a.gu, 2 * F, ACT_NONE, nullptr, 0, s));
    EN_CHECK(swiglu_fwd(a.gu, a.act, T, F, s));
  }
  EN_CHECK(linear_fwd(L.down, adapter, a.act, F, T, a.td, x_out, d, ACT_NONE, a.x1, d, s));
  return 0;
}

static int alloc_layer_acts(LayerActs& a, int64_t T, int d, int F, int H, int r, bool want_x, cudaStream_t s,
                            std::vector<void*>* track) {
  const int rr = r > 0 ? r : 8;
  if (want_x) EN_CHECK(pool_alloc(&a.x, T * d, s, track));
  EN_CHECK(pool_alloc(&a.qkv, T * 3 * d, s, track));
  EN_CHECK(pool_alloc(&a.o, T * d, s, track));
  EN_CHECK(pool_alloc(&a.x1, T * d, s, track));
  EN_CHECK(pool_alloc(&a.gu, T * 2 * F, s, track));
  EN_CHECK(pool_alloc(&a.tqkv, T * 3 * rr, s, track));
  EN_CHECK(pool_alloc(&a.to, T * rr, s, track));
  EN_CHECK(pool_alloc(&a.tgu, T * 2 * rr, s, track));
  EN_CHECK(pool_alloc(&a.td, T * rr, s, track));
  EN_CHECK(pool_alloc(&a.h1, T * d, s, track));
  EN_CHECK(pool_alloc(&a.h2, T * d, s, track));
  EN_CHECK(pool_alloc(&a.act, T * F, s, track));
  EN_CHECK(pool_alloc(&a.lse, T * H, s, track));
  return 0;
}

// ------------------------------------------------------------------------------------------------ forward
int Engine::forward_logprobs(int adapter, const float* images, int b, const int64_t* queries, const uint8_t* qmask,
                             const uint8_t* img_mask, int Q, int K, const int64_t* const* responses, int R, float temperature,
                             float* out_lp, float* out_ent, int save_slot, cudaStream_t s) {
  const opadpo_config& c = cfg;
  if (adapter < -1 || adapter > 1 || b <= 0 || K <= 0 || K > 8 || Q <= 0 || R <= 0 || !(temperature > 0.f)) {
    set_error("forward_logprobs: bad arguments (adapter %d b %d K %d Q %d R %d T %g)", adapter, b, K, Q, R, temperature);
    return OPADPO_EINVAL;
  }
  const bool save = save_slot >= 0;
  if (save_slot > 1) {
    set_error("forward_logprobs: save_slot %d out of range", save_slot);
    return OPADPO_EINVAL;
  }
  if (save && (adapter < 0 || !adapters[adapter].trainable)) {
    set_error("forward_logprobs: save_slot needs the trainable adapter");
    return OPADPO_ESTATE;
  }
  const int d = c.d_model, F = c.d_ffn, H = c.n_heads;
  // Shared-prefix packing: the image+query prefix of a sample is computed once for its K responses (identical
  // results). Rows: packed b x [query | resp_0 .. resp_{K-1}], unpacked K*b x [query | resp_k].
  const int packed = (share_prefix && K > 1) ? 1 : 0;
  const int N = packed ? b : K * b, L = packed ? Q + K * R : Q + R, S = L - 1 + P, S0 = Q - 1 + P;
  if (S > c.max_positions) {
    set_error("forward_logprobs: sequence %d exceeds max_positions %d", S, c.max_positions);
    return OPADPO_EINVAL;
  }
  const int64_t T = static_cast<int64_t>(N) * S, NR = static_cast<int64_t>(K) * b * R;
  TempPool tp(s);
  std::vector<void*>* tmp = &tp.ptrs;
  if (save) release_slot(save_slot, s);
  SaveCtx* cx = save ? &ctx[save_slot] : nullptr;
  std::vector<void*>* keep = save ? &cx->owned : tmp;

  int32_t *ids, *labels, *img_of_row, *img_pos, *pos, *src_idx, *dst_idx;
  uint8_t *tokmask, *mask;
  EN_CHECK(pool_alloc(&ids, static_cast<int64_t>(N) * L, s, tmp));
  EN_CHECK(pool_alloc(&tokmask, static_cast<int64_t>(N) * L, s, tmp));
  EN_CHECK(pool_alloc(&labels, NR, s, keep));
  EN_CHECK(pool_alloc(&img_of_row, N, s, keep));
  EN_CHECK(pool_alloc(&img_pos, N, s, keep));
  EN_CHECK(pool_alloc(&pos, T, s, keep));
  EN_CHECK(pool_alloc(&src_idx, NR, s, keep));
  EN_CHECK(pool_alloc(&dst_idx, NR, s, keep));
  EN_CHECK(pool_alloc(&mask, T, s, tmp));
  EN_CHECK(stack_inputs(queries, qmask, responses, K, b, Q, R, c.pad_id, packed, ids, tokmask, labels, s));
  EN_CHECK(build_positions(pos, N, S, S0, R, packed, s));
  EN_CHECK(build_row_index(src_idx, dst_idx, K, b, R, S, S0, packed, static_cast<int>(T), s));
  EN_CHECK(fill_mod(img_of_row, N, b, s));  // images are shared by the K responses of a sample (rl_models.py:107-111)
  // vision tower + projector, once per unique image
  const int64_t PR = static_cast<int64_t>(b) * P;
  const int rr = r > 0 ? r : 8;
  bf16 *clip_feats, *p_pre, *p_act, *t_p0, *t_p2, *feats;
  EN_CHECK(pool_alloc(&clip_feats, PR * c.v_d, s, keep));
  EN_CHECK(pool_alloc(&p_pre, PR * d, s, keep));
  EN_CHECK(pool_alloc(&p_act, PR * d, s, keep));
  EN_CHECK(pool_alloc(&t_p0, PR * rr, s, keep));
  EN_CHECK(pool_alloc(&t_p2, PR * rr, s, keep));
  EN_CHECK(pool_alloc(&feats, PR * d, s, tmp));
  {
    TempPool clip_tp(s);
    EN_CHECK(clip_forward(adapter, images, b, clip_feats, s, &clip_tp.ptrs));
  }
  EN_CHECK(projector_forward(adapter, clip_feats, PR, p_pre, p_act, t_p0, t_p2, feats, s));

  // embedding gather + image splice
  bf16 *x0, *x1buf = nullptr;
  EN_CHECK(pool_alloc(&x0, T * d, s, keep));
  EN_CHECK(embed_splice(ids, tokmask, img_mask, img_of_row, embed, feats, x0, mask, img_pos, N, L, P, d, c.vocab, s));
  uint64_t* kbits;
  const int nwords = (S + 63) / 64;
  EN_CHECK(pool_alloc(&kbits, static_cast<int64_t>(N) * nwords, s, keep));
  EN_CHECK(attn_pack_mask(mask, kbits, N, S, s));
  SeqLayout lay;
  lay.N = N;
  lay.S = S;
  lay.S0 = packed ? S0 : 0;
  lay.segR = packed ? R : 0;
  lay.pos = pos;
  lay.kbits = kbits;

  // decoder weights of this pass: a frozen adapter's inference forward may run on the merged copy (no LoRA GEMMs)
  const std::vector<LlamaLayer>* LL = &layers;
  int llm_adapter = adapter;
  if (!save && merge_frozen && adapter >= 0 && r > 0 && !adapters[adapter].trainable && c.n_layers > 0) {
    EN_CHECK(ensure_frozen_merged(adapter, s));
    LL = &frozen_merged;
    llm_adapter = -1;
  }
  bf16* x_final = nullptr;
  if (save) {
    cx->layers.resize(c.n_layers);
    for (int l = 0; l < c.n_layers; ++l) {
      EN_CHECK(alloc_layer_acts(cx->layers[l], T, d, F, H, r, l > 0, s, keep));
    }
    if (c.n_layers > 0) {
      cx->layers[0].x = x0;
      EN_CHECK(pool_alloc(&x_final, T * d, s, keep));
    } else {
      x_final = x0;
    }
    for (int l = 0; l < c.n_layers; ++l) {
      bf16* xo = l + 1 < c.n_layers ? cx->layers[l + 1].x : x_final;
      EN_CHECK(llama_layer_fwd(layers[l], adapter, lay, cx->layers[l], xo, s));
    }
  } else {
    LayerActs sc;
    EN_CHECK(alloc_layer_acts(sc, T, d, F, H, r, false, s, tmp));
    EN_CHECK(pool_alloc(&x1buf, T * d, s, tmp));
    bf16* cur = x0;
    bf16* nxt = x1buf;
    for (int l = 0; l < c.n_layers; ++l) {
      sc.x = cur;
      EN_CHECK(llama_layer_fwd((*LL)[l], llm_adapter, lay, sc, nxt, s));
      bf16* t = cur;
      cur = nxt;
      nxt = t;
    }
    x_final = cur;
  }

  // final norm on the R response-predicting rows only, then the fused lm_head + log-sum-exp + label gather
  bf16* hn;
  EN_CHECK(pool_alloc(&hn, NR * d, s, keep));
  EN_CHECK(rmsnorm_fwd(x_final, final_norm, hn, NR, d, c.rms_eps, src_idx, s));
  const int ntiles = gemm_num_n_tiles(c.vocab);
  float4* partials;
  float *label_logit, *lse_head, *ent_head = nullptr;
  EN_CHECK(pool_alloc(&partials, NR * ntiles, s, tmp));
  EN_CHECK(pool_alloc(&label_logit, NR, s, tmp));
  EN_CHECK(pool_alloc(&lse_head, NR, s, keep));
  if (save) EN_CHECK(pool_alloc(&ent_head, NR, s, keep));
  cudaMemsetAsync(label_logit, 0, NR * sizeof(float), s);
  {
    GemmArgs a;
    a.M = static_cast<int>(NR);
    a.N = c.vocab;
    a.K1 = d;
    a.A1 = {hn, d};
    a.B1 = {lm_head, d};
    a.alpha = 1.f / temperature;
    a.epilogue = EPI_LSE;
    a.labels = labels;
    a.lse_partials = partials;
    a.label_logit = label_logit;
    EN_CHECK(gemm_checked(a, s));
  }
  EN_CHECK(logprob_finalize(partials, ntiles, label_logit, labels, out_lp, out_ent, lse_head, ent_head, NR, s));

  if (save) {
    cx->valid = true;
    cx->adapter = adapter;
    cx->b = b;
    cx->K = K;
    cx->N = N;
    cx->Q = Q;
    cx->R = R;
    cx->S = S;
    cx->inv_temp = 1.f / temperature;
    cx->x_final = x_final;
    cx->hn = hn;
    cx->labels = labels;
    cx->lse_head = lse_head;
    cx->ent_head = ent_head;
    cx->kbits = kbits;
    cx->lay = lay;
    cx->src_idx = src_idx;
    cx->dst_idx = dst_idx;
    cx->img_pos = img_pos;
    cx->img_of_row = img_of_row;
    cx->clip_feats = clip_feats;
    cx->p_pre = p_pre;
    cx->p_act = p_act;
    cx->t_p0 = t_p0;
    cx->t_p2 = t_p2;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("forward_logprobs: %s", cudaGetErrorString(e));
    return OPADPO_ECUDA;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ backward
int Engine::backward(int save_slot, const float* dlp, const float* dent, cudaStream_t s) {
  const opadpo_config& c = cfg;
  if (save_slot < 0 || save_slot > 1 || !ctx[save_slot].valid) {
    set_error("backward: slot %d holds no saved forward", save_slot);
    return OPADPO_ESTATE;
  }
  SaveCtx& cx = ctx[save_slot];
  const int adapter = cx.adapter;
  const int d = c.d_model, F = c.d_ffn, H = c.n_heads, N = cx.N, S = cx.S, R = cx.R;
  const int64_t T = static_cast<int64_t>(N) * S, NR = static_cast<int64_t>(cx.K) * cx.b * R;
  const int64_t T_scratch = T + static_cast<int64_t>(cx.K) * cx.b;  // + shared last-prefix-row scratch (packed rows)
  const int rr = r > 0 ? r : 8;
  TempPool tp(s);
  std::vector<void*>* tmp = &tp.ptrs;
  bf16 *dlogits, *dhn, *dxa, *dxb, *dact, *dgu, *dqkv, *dh, *dob, *u;
  float* delta;
  EN_CHECK(pool_alloc(&dlogits, NR * c.vocab, s, tmp));
  EN_CHECK(pool_alloc(&dhn, NR * d, s, tmp));
  EN_CHECK(pool_alloc(&dxa, T_scratch * d, s, tmp));
  EN_CHECK(pool_alloc(&dxb, T * d, s, tmp));
  EN_CHECK(pool_alloc(&dact, T * F, s, tmp));
  EN_CHECK(pool_alloc(&dgu, T * 2 * F, s, tmp));
  EN_CHECK(pool_alloc(&dqkv, T * 3 * d, s, tmp));
  EN_CHECK(pool_alloc(&dh, T * d, s, tmp));
  EN_CHECK(pool_alloc(&dob, T * d, s, tmp));
  EN_CHECK(pool_alloc(&u, T * 3 * rr, s, tmp));   // u of q|k|v (and, after the decoder, of the projector)
  // one u buffer per linear kind: the side-stream dA GEMM of a layer reads it while the main stream already runs the next linear
  bf16 *u_down, *u_gu, *u_o;
  EN_CHECK(pool_alloc(&u_down, T * rr, s, tmp));
  EN_CHECK(pool_alloc(&u_gu, T * 2 * rr, s, tmp));
  EN_CHECK(pool_alloc(&u_o, T * rr, s, tmp));
  const bool side = wgrad_stream && r > 0 && c.n_layers > 0;
  if (side) {
    EN_CHECK(ensure_wgrad_stream());
    for (int k = 0; k < 4; ++k) side_pending[k] = false;
  }
  enum { K_DOWN = 0, K_GU = 1, K_O = 2, K_QKV = 3 };
  const int kd = side ? K_DOWN : -1, kg = side ? K_GU : -1, ko = side ? K_O : -1, kq = side ? K_QKV : -1;
  int pending_cb = -1;  // a layer whose slice becomes final once its side-stream weight gradients are done
  EN_CHECK(pool_alloc(&delta, T * H, s, tmp));

  {  // d logits = dlp/T * (onehot - softmax) [- dent/T * p (ln p + H)], recomputed tile-wise from the saved LSE (never stored
     // in fp32)
    GemmArgs a;
    a.M = static_cast<int>(NR);
    a.N = c.vocab;
    a.K1 = d;
    a.A1 = {cx.hn, d};
    a.B1 = {lm_head, d};
    a.alpha = cx.inv_temp;
    a.epilogue = EPI_CEGRAD;
    a.labels = cx.labels;
    a.lse = cx.lse_head;
    a.dlp = dlp;
    a.dent = dent;  // entropy head (OPA entropy regulariser); null for the DPO / SFT losses
    a.ent = dent ? cx.ent_head : nullptr;
    a.C = dlogits;
    a.ldc = c.vocab;
    EN_CHECK(gemm_checked(a, s));
    GemmArgs g;  // d hn = dlogits W_lm
    g.M = static_cast<int>(NR);
    g.N = d;
    g.K1 = c.vocab;
    g.A1 = {dlogits, c.vocab};
    g.b_mn_major = true;
    g.B1 = {lm_head, d};
    g.C = dhn;
    g.ldc = d;
    EN_CHECK(gemm_checked(g, s));
  }
  cudaMemsetAsync(dxa, 0, static_cast<size_t>(T_scratch) * d * 2, s);
  EN_CHECK(rmsnorm_bwd(dhn, cx.x_final, final_norm, nullptr, dxa, NR, d, c.rms_eps, cx.src_idx, cx.dst_idx, s));
  if (cx.lay.segR > 0) EN_CHECK(fold_rows(dxa, cx.K, cx.b, S, cx.lay.S0, d, static_cast<int>(T), s));

  bf16* dx = dxa;   // gradient w.r.t. the layer output
  bf16* dx1 = dxb;  // gradient w.r.t. the post-attention residual stream
  for (int l = c.n_layers - 1; l >= 0; --l) {
    LlamaLayer& L = layers[l];
    const LayerActs& a = cx.layers[l];
    // MLP
    // Main-stream writers wait for the side stream's readers of the buffer they are about to overwrite (wait_side): u_k and
    // the dy of linear kind k were last read by that kind's weight gradients one layer ago — normally long done.
    wait_side(kd, s);  // u_down (previous layer's dA)
    wait_side(kg, s);  // dgu and u_gu (previous layer's gate|up weight gradients)
    if (fuse_swiglu_bwd && F % 32 == 0) {
      EN_CHECK(linear_bwd(L.down, adapter, a.act, F, a.td, dx, d, T, u_down, dgu, 2 * F, nullptr, 0, s, kd, a.gu));
    } else {
      EN_CHECK(linear_bwd(L.down, adapter, a.act, F, a.td, dx, d, T, u_down, dact, F, nullptr, 0, s, kd));
      EN_CHECK(swiglu_bwd(dact, a.gu, dgu, T, F, s));
    }
    EN_CHECK(linear_bwd(L.gu, adapter, a.h2, d, a.tgu, dgu, 2 * F, T, u_gu, dh, d, nullptr, 0, s, kg));
    wait_side(ko, s);  // dx1 and u_o (previous layer's o-projection weight gradients)
    EN_CHECK(rmsnorm_bwd(dh, a.x1, L.ln2, dx, dx1, T, d, c.rms_eps, nullptr, nullptr, s));
    // attention
    EN_CHECK(linear_bwd(L.o, adapter, a.o, d, a.to, dx1, d, T, u_o, dob, d, nullptr, 0, s, ko));
    AttnArgs at;
    at.N = N;
    at.S = S;
    at.H = H;
    at.HD = hd;
    at.causal = true;
    at.scale = 1.f / sqrtf(static_cast<float>(hd));
    at.q = a.qkv;
    at.k = a.qkv + d;
    at.v = a.qkv + 2 * d;
    at.ld_qkv = 3 * d;
    at.o = a.o;
    at.ld_o = d;
    at.lse = a.lse;
    at.kbits = cx.kbits;
    at.prefix_len = cx.lay.S0;
    at.segment_len = cx.lay.segR;
    at.dO = dob;
    at.ld_do = d;
    at.delta = delta;
    at.dq = dqkv;
    at.dk = dqkv + d;
    at.dv = dqkv + 2 * d;
    at.ld_dqkv = 3 * d;
    wait_side(kq, s);  // dqkv and u (previous layer's q|k|v weight gradients: the last side-stream work of that layer)
    if (pending_cb >= 0) {  // ... so the previous layer's LoRA-grad slice is now complete in main-stream order
      if (layer_cb) layer_cb(pending_cb, layer_cb_user);
      pending_cb = -1;
    }
    if (fuse_rope && fuse_rope_bwd && hd == 128) {  // inverse rotary embedding of dq / dk inside the attention backward's epilogues
      at.rope_cs = rope_cs;
      at.rope_pos = cx.lay.pos;
    }
    EN_CHECK(attn_bwd_checked(at, s));
    if (!at.rope_cs) EN_CHECK(rope_inplace(dqkv, 3 * d, T, S, cx.lay.pos, 0, nullptr, 2 * H, hd, rope_cs, 1, s));
    EN_CHECK(linear_bwd(L.qkv, adapter, a.h1, d, a.tqkv, dqkv, 3 * d, T, u, dh, d, nullptr, 0, s, kq));
    wait_side(kd, s);  // dx is this layer's down-projection dy: its weight gradients must have read it before it is rewritten
    EN_CHECK(rmsnorm_bwd(dh, a.x, L.ln1, dx1, dx, T, d, c.rms_eps, nullptr, nullptr, s));
    if (side)
      pending_cb = l;  // reported once the main stream has waited for this layer's side-stream work
    else if (layer_cb)
      layer_cb(l, layer_cb_user);  // this layer's LoRA-grad slice is complete in stream order
  }
  if (side) {  // join: every weight gradient is in the bucket before anything downstream (projector, all-reduce, optimizer)
    for (int k = 0; k < 4; ++k) wait_side(k, s);
    if (side_error) {
      side_error = false;
      cudaStreamSynchronize(wg_stream);
      set_error("backward: ordering the weight-gradient side stream failed (%s)", cudaGetErrorString(cudaGetLastError()));
      return OPADPO_ECUDA;
    }
    if (pending_cb >= 0 && layer_cb) layer_cb(pending_cb, layer_cb_user);
  }
  // image rows -> projector (its LoRA adapters are trainable; the CLIP tower is not differentiated)
  const int64_t PR = static_cast<int64_t>(cx.b) * P;
  bf16 *dfeats, *dpact, *dppre;
  EN_CHECK(pool_alloc(&dfeats, PR * d, s, tmp));
  EN_CHECK(pool_alloc(&dpact, PR * d, s, tmp));
  EN_CHECK(pool_alloc(&dppre, PR * d, s, tmp));
  EN_CHECK(embed_splice_bwd(dx, cx.img_of_row, cx.img_pos, dfeats, N, S, cx.b, P, d, s));
  EN_CHECK(linear_bwd(proj2, adapter, cx.p_act, d, cx.t_p2, dfeats, d, PR, u, dpact, d, nullptr, 0, s));
  EN_CHECK(gelu_bwd(dpact, cx.p_pre, dppre, PR * d, s));
  EN_CHECK(linear_bwd(proj0, adapter, cx.clip_feats, c.v_d, cx.t_p0, dppre, d, PR, u, nullptr, 0, nullptr, 0, s));
  if (layer_cb) layer_cb(c.n_layers, layer_cb_user);
  release_slot(save_slot, s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("backward: %s", cudaGetErrorString(e));
    return OPADPO_ECUDA;
  }
  return 0;
}

// Merged decoder weights for rollout: per group g, W'_g = W_g + (alpha/r) B_g A_g as one GEMM with W_g as the residual
// (A operand = B_g [out_g, r] K-major, B operand = A_g [r, in] MN-major).
int Engine::merge_lora(int adapter, std::vector<LlamaLayer>* merged, std::vector<void*>* track, cudaStream_t s) {
  merged->assign(layers.begin(), layers.end());
  for (size_t l = 0; l < layers.size(); ++l) {
    LinearW* src[4] = {&layers[l].qkv, &layers[l].o, &layers[l].gu, &layers[l].down};
    LinearW* dst[4] = {&(*merged)[l].qkv, &(*merged)[l].o, &(*merged)[l].gu, &(*merged)[l].down};
    for (int i = 0; i < 4; ++i) {
      const LinearW& L = *src[i];
      if (L.slot < 0) continue;
      bf16* Wm = nullptr;
      EN_CHECK(pool_alloc(&Wm, static_cast<int64_t>(L.out) * L.in, s, track));
      const int G = L.groups, out_g = L.out / G;
      for (int g = 0; g < G; ++g) {
        GemmArgs a;
        a.M = out_g;
        a.N = L.in;
        a.K1 = r;
        a.A1 = {loraB(adapter, L) + static_cast<int64_t>(g) * out_g * r, r};
        a.b_mn_major = true;
        a.B1 = {loraA(adapter, L) + static_cast<int64_t>(g) * r * L.in, L.in};
        a.alpha = lora_scale;
        a.residual = L.W + static_cast<int64_t>(g) * out_g * L.in;
        a.ldr = L.in;
        a.C = Wm + static_cast<int64_t>(g) * out_g * L.in;
        a.ldc = L.in;
        EN_CHECK(gemm_checked(a, s));
      }
      dst[i]->W = Wm;
      dst[i]->slot = -1;
    }
  }
  return 0;
}

void Engine::drop_frozen_merged() {
  if (frozen_merged_allocs.empty() && frozen_merged_adapter < 0) return;
  cudaDeviceSynchronize();
  for (void* p : frozen_merged_allocs) cudaFree(p);
  frozen_merged_allocs.clear();
  frozen_merged.clear();
  frozen_merged_adapter = -1;
}

int Engine::ensure_frozen_merged(int adapter, cudaStream_t s) {
  if (frozen_merged_adapter == adapter) return 0;
  drop_frozen_merged();
  int rc = merge_lora(adapter, &frozen_merged, &frozen_merged_allocs, s);
  if (rc == 0 && fuse_swiglu && cfg.d_ffn % 128 == 0) {
    for (auto& L : frozen_merged) {
      if (L.gu.slot >= 0 || L.gu.bias) continue;  // only projections that really were merged
      bf16* il = nullptr;
      rc = pool_alloc(&il, static_cast<int64_t>(L.gu.out) * L.gu.in, s, &frozen_merged_allocs);
      if (rc == 0) rc = interleave_gate_up(L.gu.W, il, cfg.d_ffn, L.gu.in, s) == 0 ? 0 : OPADPO_ECUDA;
      if (rc != 0) break;
      L.gu.W_il = il;
    }
  }
  if (rc != 0) {
    drop_frozen_merged();
    return rc;
  }
  frozen_merged_adapter = adapter;
  return 0;
}

int Engine::set_option(const char* name, int64_t value) {
  if (strcmp(name, "merge_frozen") == 0) {
    merge_frozen = value != 0;
    if (!merge_frozen) drop_frozen_merged();
    return 0;
  }
  if (strcmp(name, "fuse_swiglu") == 0) {
    if (fuse_swiglu != (value != 0)) drop_frozen_merged();  // the merged copy is built with or without the interleaved weight
    fuse_swiglu = value != 0;
    return 0;
  }
  if (strcmp(name, "fuse_rope_bwd") == 0) {
    fuse_rope_bwd = value != 0;
    return 0;
  }
  if (strcmp(name, "fuse_swiglu_bwd") == 0) {
    fuse_swiglu_bwd = value != 0;
    return 0;
  }
  if (strcmp(name, "wgrad_stream") == 0) {
    wgrad_stream = value != 0;
    return 0;
  }
  if (strcmp(name, "fuse_rope") == 0) {
    fuse_rope = value != 0;
    return 0;
  }
  if (strcmp(name, "share_prefix") == 0) {
    share_prefix = value != 0;
    return 0;
  }
  if (strcmp(name, "decode_swap_ab") == 0) {
    decode_swap_ab = value != 0;
    return 0;
  }
  if (strcmp(name, "merge_rollout") == 0) {
    merge_rollout = value != 0;
    return 0;
  }
  if (strcmp(name, "adam_step") == 0) {  // optimizer-state restore (resume_training, dpo_trainer.py:901-931)
    if (trainable_adapter < 0 || value < 0) {
      set_error("adam_step: no trainable adapter or negative step");
      return OPADPO_ESTATE;
    }
    adapters[trainable_adapter].step = static_cast<int>(value);
    return 0;
  }
  set_error("unknown option '%s'", name);
  return OPADPO_EINVAL;
}

int Engine::set_backward_callback(opadpo_layer_cb cb, void* user) {
  layer_cb = cb;
  layer_cb_user = user;
  return 0;
}

// Gradient-bucket slice of decoder layer `layer` (0..n_layers-1) or of the projector (layer == n_layers): the flat
// bucket is ordered layer by layer, so each slice is contiguous and final as soon as that layer's backward has run.
int Engine::layer_grad_range(int layer, int64_t* offset, int64_t* count) {
  if (layer < 0 || layer > cfg.n_layers) {
    set_error("layer_grad_range: layer %d out of range", layer);
    return OPADPO_EINVAL;
  }
  const int64_t begin = layer < cfg.n_layers ? slots[layers[layer].qkv.slot].offA : slots[proj0.slot].offA;
  const int64_t end = layer + 1 < cfg.n_layers ? slots[layers[layer + 1].qkv.slot].offA
                                               : (layer + 1 == cfg.n_layers ? slots[proj0.slot].offA : n_train);
  *offset = begin;
  *count = end - begin;
  return 0;
}

int Engine::adopt_grad_buffer(float* buffer, int64_t count) {
  if (trainable_adapter < 0) {
    set_error("adopt_grad_buffer: no trainable adapter");
    return OPADPO_ESTATE;
  }
  if (!buffer || count < n_train || (reinterpret_cast<uintptr_t>(buffer) & 255) != 0) {
    set_error("adopt_grad_buffer: need a 256 B-aligned device buffer of at least %lld floats", static_cast<long long>(n_train));
    return OPADPO_EINVAL;
  }
  Adapter& ad = adapters[trainable_adapter];
  if (ad.grad == buffer) return 0;
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(buffer, ad.grad, static_cast<size_t>(n_train) * 4, cudaMemcpyDeviceToDevice);
  if (e != cudaSuccess) {
    set_error("adopt_grad_buffer: %s", cudaGetErrorString(e));
    return OPADPO_ECUDA;
  }
  for (size_t i = 0; i < allocs.size(); ++i)
    if (allocs[i] == ad.grad) {  // the engine's own bucket (an adopted one is the caller's to free)
      cudaFree(ad.grad);
      allocs.erase(allocs.begin() + i);
      break;
    }
  ad.grad = buffer;
  return 0;
}

int Engine::zero_grad(cudaStream_t s) {
  if (trainable_adapter < 0) return 0;
  cudaError_t e = cudaMemsetAsync(adapters[trainable_adapter].grad, 0, static_cast<size_t>(n_train) * 4, s);
  if (e != cudaSuccess) {
    set_error("zero_grad: %s", cudaGetErrorString(e));
    return OPADPO_ECUDA;
  }
  return 0;
}

int Engine::optimizer_step(float lr, float b1, float b2, float eps, float wd, float max_norm, float gscale, float* norm_out,
                           cudaStream_t s) {
  if (trainable_adapter < 0) {
    set_error("optimizer_step: no trainable adapter");
    return OPADPO_ESTATE;
  }
  Adapter& ad = adapters[trainable_adapter];
  cudaMemsetAsync(sumsq, 0, sizeof(float), s);
  EN_CHECK(sumsq_accum(ad.grad, n_train, sumsq, s));
  if (norm_out) EN_CHECK(norm_from_sumsq(sumsq, gscale, norm_out, s));
  ad.step += 1;
  EN_CHECK(adamw_step(ad.master, ad.grad, ad.m, ad.v, ad.shadow, n_train, lr, b1, b2, eps, wd, ad.step, gscale, sumsq,
                      max_norm, s));
  return 0;
}

// ------------------------------------------------------------------------------------------------ greedy rollout
// HF GenerationMixin.greedy_search semantics [UPSTREAM] as driven by online_generator.py:292-309 with do_sample=False:
// multimodal prefill, then one token per step; finished rows emit pad; stops early when every row is finished
// (remaining columns are pad) — here the loop always runs max_new steps on device and pads, which yields the same ids.
int Engine::generate(int adapter, const float* images, int b, const int64_t* queries, const uint8_t* qmask, int Q,
                     int max_new, const opadpo_sampler* sampler, int64_t* out_ids, cudaStream_t s) {
  const opadpo_config& c = cfg;
  if (adapter < -1 || adapter > 1 || b <= 0 || Q <= 0 || max_new <= 0) {
    set_error("generate: bad arguments");
    return OPADPO_EINVAL;
  }
  const bool do_sample = sampler && sampler->do_sample;
  if (do_sample && (!(sampler->temperature > 0.f) || sampler->top_k < 0 || !(sampler->top_p > 0.f) || sampler->top_p > 1.f)) {
    set_error("generate: bad sampler (temperature %g top_k %d top_p %g): need temperature > 0, top_k >= 0, top_p in (0, 1]",
              sampler->temperature, sampler->top_k, sampler->top_p);
    return OPADPO_EINVAL;
  }
  const int n_sup = sampler ? sampler->n_suppress : 0;
  const int32_t* sup = sampler ? sampler->suppress_tokens : nullptr;
  if (n_sup < 0 || n_sup > OPADPO_MAX_SUPPRESS) {
    set_error("generate: n_suppress %d out of range [0, %d]", n_sup, OPADPO_MAX_SUPPRESS);
    return OPADPO_EINVAL;
  }
  if (do_sample && c.vocab > 40960) {
    set_error("generate: the sampling kernel holds at most 40960 vocabulary entries (vocab %d)", c.vocab);
    return OPADPO_EINVAL;
  }
  const int d = c.d_model, F = c.d_ffn, H = c.n_heads, V = c.vocab;
  const int N = b, S0 = Q - 1 + P, S_max = S0 + max_new;
  if (S_max > c.max_positions) {
    set_error("generate: %d positions exceed max_positions %d", S_max, c.max_positions);
    return OPADPO_EINVAL;
  }
  const int64_t T = static_cast<int64_t>(N) * S0;
  const int rr = r > 0 ? r : 8;
  TempPool tp(s);
  std::vector<void*>* tmp = &tp.ptrs;
  // Decoder weights used by this rollout: with an adapter and merge_rollout, W' = W + (alpha/r) B A is built once
  // (~3 TFLOP) so every decode linear is ONE weight-streaming GEMM; the vision tower / projector stay unmerged.
  std::vector<LlamaLayer> merged;
  const std::vector<LlamaLayer>* LL = &layers;
  int llm_adapter = adapter;
  if (adapter >= 0 && r > 0 && merge_rollout) {
    if (merge_frozen && !adapters[adapter].trainable) {  // a frozen adapter's merged copy is cached across calls
      EN_CHECK(ensure_frozen_merged(adapter, s));
      LL = &frozen_merged;
    } else {
      EN_CHECK(merge_lora(adapter, &merged, tmp, s));
      LL = &merged;
    }
    llm_adapter = -1;
  }

  // ---- prefill (same kernels as the training forward; K/V copied into the static cache)
  int32_t *ids, *img_of_row, *img_pos, *next_ids;
  uint8_t *mask, *kmask, *done;
  uint8_t* tokmask;
  EN_CHECK(pool_alloc(&ids, static_cast<int64_t>(N) * Q, s, tmp));
  EN_CHECK(pool_alloc(&tokmask, static_cast<int64_t>(N) * Q, s, tmp));
  EN_CHECK(pool_alloc(&img_of_row, N, s, tmp));
  EN_CHECK(pool_alloc(&img_pos, N, s, tmp));
  EN_CHECK(pool_alloc(&next_ids, N, s, tmp));
  EN_CHECK(pool_alloc(&mask, T, s, tmp));
  EN_CHECK(pool_alloc(&kmask, static_cast<int64_t>(N) * S_max, s, tmp));
  EN_CHECK(pool_alloc(&done, N, s, tmp));
  EN_CHECK(queries_to_i32(queries, qmask, ids, tokmask, out_ids, static_cast<int64_t>(Q) + max_new, N, Q, s));
  EN_CHECK(fill_mod(img_of_row, N, N, s));
  const int64_t PR = static_cast<int64_t>(N) * P;
  bf16 *clip_feats, *p_pre, *p_act, *t_p0, *t_p2, *feats, *x0, *x1buf;
  EN_CHECK(pool_alloc(&clip_feats, PR * c.v_d, s, tmp));
  EN_CHECK(pool_alloc(&p_pre, PR * d, s, tmp));
  EN_CHECK(pool_alloc(&p_act, PR * d, s, tmp));
  EN_CHECK(pool_alloc(&t_p0, PR * rr, s, tmp));
  EN_CHECK(pool_alloc(&t_p2, PR * rr, s, tmp));
  EN_CHECK(pool_alloc(&feats, PR * d, s, tmp));
  {
    TempPool clip_tp(s);
    EN_CHECK(clip_forward(adapter, images, N, clip_feats, s, &clip_tp.ptrs));
  }
  EN_CHECK(projector_forward(adapter, clip_feats, PR, p_pre, p_act, t_p0, t_p2, feats, s));
  EN_CHECK(pool_alloc(&x0, T * d, s, tmp));
  EN_CHECK(pool_alloc(&x1buf, T * d, s, tmp));
  EN_CHECK(embed_splice(ids, tokmask, nullptr, img_of_row, embed, feats, x0, mask, img_pos, N, Q, P, d, V, s));
  uint64_t* kbits;
  EN_CHECK(pool_alloc(&kbits, static_cast<int64_t>(N) * ((S0 + 63) / 64), s, tmp));
  EN_CHECK(attn_pack_mask(mask, kbits, N, S0, s));
  // key mask for decode: prefill mask, then ones for generated positions (llava_modifications.patch:12161-12175)
  cudaMemsetAsync(kmask, 1, static_cast<size_t>(N) * S_max, s);
  cudaMemcpy2DAsync(kmask, S_max, mask, S0, S0, N, cudaMemcpyDeviceToDevice, s);
  cudaMemsetAsync(done, 0, N, s);

  std::vector<bf16*> kc(c.n_layers), vc(c.n_layers);
  const int64_t cache_elems = static_cast<int64_t>(N) * S_max * d;
  for (int l = 0; l < c.n_layers; ++l) {
    EN_CHECK(pool_alloc(&kc[l], cache_elems, s, tmp));
    EN_CHECK(pool_alloc(&vc[l], cache_elems, s, tmp));
  }
  LayerActs sc;
  EN_CHECK(alloc_layer_acts(sc, T, d, F, H, r, false, s, tmp));
  SeqLayout lay;
  lay.N = N;
  lay.S = S0;
  lay.kbits = kbits;
  int32_t* last_rows;
  EN_CHECK(pool_alloc(&last_rows, N, s, tmp));
  EN_CHECK(fill_affine(last_rows, N, S0, S0 - 1, s));
  bf16* x_prefill_out = x0;
  {
    bf16* nxt = x1buf;
    for (int l = 0; l < c.n_layers; ++l) {
      sc.x = x_prefill_out;
      EN_CHECK(llama_layer_fwd((*LL)[l], llm_adapter, lay, sc, nxt, s));
      EN_CHECK(kv_store_prefill(sc.qkv, kc[l], vc[l], N, S0, S_max, d, s));
      bf16* t = x_prefill_out;
      x_prefill_out = nxt;
      nxt = t;
    }
  }
  // ---- decode: one row per sequence. Every linear is a weight-streaming tcgen05 GEMM split over K so all SMs pull
  // weights (EPI_F32_PARTIAL slices, reduced in a fixed order by the kernel that consumes them); the step reads its
  // position from a device counter, so steps 2.. replay ONE captured CUDA graph.
  const int64_t Nl = N;
  bf16 *xa, *xb, *x1d, *hdn, *qkvd, *od, *actd, *td;
  float* part;
  int32_t* cur;
  EN_CHECK(pool_alloc(&xa, Nl * d, s, tmp));
  EN_CHECK(pool_alloc(&xb, Nl * d, s, tmp));
  EN_CHECK(pool_alloc(&x1d, Nl * d, s, tmp));
  EN_CHECK(pool_alloc(&hdn, Nl * d, s, tmp));
  EN_CHECK(pool_alloc(&qkvd, Nl * 3 * d, s, tmp));
  EN_CHECK(pool_alloc(&od, Nl * d, s, tmp));
  EN_CHECK(pool_alloc(&actd, Nl * F, s, tmp));
  EN_CHECK(pool_alloc(&td, Nl * 3 * rr, s, tmp));
  EN_CHECK(pool_alloc(&cur, 1, s, tmp));
  const int kMaxSplits = 32;
  int64_t widest = std::max<int64_t>(std::max<int64_t>(3 * d, 2 * F), V);
  EN_CHECK(pool_alloc(&part, static_cast<int64_t>(kMaxSplits) * Nl * widest, s, tmp));
  cudaMemsetAsync(cur, 0, sizeof(int32_t), s);

  const bool lora = llm_adapter >= 0 && r > 0;
  // K-split count: every SM should stream weights. cost = waves of CTAs x k-blocks per split.
  auto splits_for = [&](int n_out, int K1, int K2) {
    const int tiles = ((N + 127) / 128) * ((n_out + 255) / 256);
    const int kb = (K1 + 63) / 64 + (K2 + 63) / 64;
    int best = 1;
    long best_cost = 1L << 60;
    for (int sp = 1; sp <= kMaxSplits && sp <= kb; ++sp) {
      const long waves = (static_cast<long>(tiles) * sp + 147) / 148;
      const long cost = waves * ((kb + sp - 1) / sp) * 16 + sp;  // small bias towards fewer slices to reduce
      if (cost < best_cost) {
        best_cost = cost;
        best = sp;
      }
    }
    return gemm_effective_splits(K1, K2, best);
  };
  // partial sums of y = x W^T + (s x A^T) B^T into `part` [splits, N, out]; returns the split count
  auto linear_partial = [&](const LinearW& L, const bf16* x, int max_sp, int* nsplit) -> int {
    const int G = L.groups;
    if (decode_swap_ab && !(lora && L.slot >= 0) && N <= 128) {
      // merged (or adapter-free) weights: weight-streaming GEMM with the weight rows on the MMA's M (decode_gemm.cu)
      DecodeGemmArgs a;
      a.Mb = N;
      a.N = L.out;
      a.K = L.in;
      a.x = x;
      a.ldx = L.in;
      a.W = L.W;
      a.ldw = L.in;
      a.part = part;
      a.ldc = L.out;
      a.splits = decode_gemm_pick_splits(N, L.out, L.in, max_sp > 0 ? std::min(max_sp, kMaxSplits) : kMaxSplits);
      int rc = decode_gemm_launch(a, s);
      if (rc != 0) {
        set_error("%s", decode_gemm_last_error());
        return rc <= -10 ? OPADPO_ECUDA : OPADPO_EINVAL;
      }
      *nsplit = a.splits;
      return 0;
    }
    if (lora && L.slot >= 0) {
      const int spA = splits_for(G * r, L.in, 0);
      GemmArgs a;
      a.M = N;
      a.N = G * r;
      a.K1 = L.in;
      a.A1 = {x, L.in};
      a.B1 = {loraA(llm_adapter, L), L.in};
      a.alpha = lora_scale;
      a.epilogue = EPI_F32_PARTIAL;
      a.splits = spA;
      a.C = part;
      a.ldc = G * r;
      EN_CHECK(gemm_checked(a, s));
      EN_CHECK(partial_to_bf16(part, spA, Nl * G * r, td, s));
    }
    const int K2 = (lora && L.slot >= 0) ? r : 0;
    int sp = splits_for(L.out, L.in, K2);
    if (max_sp > 0 && sp > max_sp) sp = gemm_effective_splits(L.in, K2, max_sp);
    GemmArgs a;
    a.M = N;
    a.N = L.out;
    a.K1 = L.in;
    a.A1 = {x, L.in};
    a.B1 = {L.W, L.in};
    if (K2 > 0) {
      a.K2 = r;
      a.A2 = {td, static_cast<int64_t>(G) * r};
      a.B2 = {loraB(llm_adapter, L), r};
      if (G > 1) {
        if ((L.out / G) % 256 != 0) {
          set_error("generate: fused group width %d must be a multiple of 256", L.out / G);
          return OPADPO_EINVAL;
        }
        a.a2_group_n = L.out / G;
        a.a2_group_stride = r;
      }
    }
    a.epilogue = EPI_F32_PARTIAL;
    a.splits = sp;
    a.C = part;
    a.ldc = L.out;
    EN_CHECK(gemm_checked(a, s));
    *nsplit = sp;
    return 0;
  };
  // fp32 logits with at most 2 K-splits: a + b is order independent, so argmax ties break exactly as torch.argmax
  auto head_and_pick = [&](const bf16* hrows) -> int {
    int sp = 1;
    if (decode_swap_ab && N <= 128) {
      DecodeGemmArgs a;
      a.Mb = N;
      a.N = V;
      a.K = d;
      a.x = hrows;
      a.ldx = d;
      a.W = lm_head;
      a.ldw = d;
      a.part = part;
      a.ldc = V;
      a.splits = sp = decode_gemm_pick_splits(N, V, d, 2);   // at most two slices: a + b is order independent (see above)
      int rc = decode_gemm_launch(a, s);
      if (rc != 0) {
        set_error("%s", decode_gemm_last_error());
        return rc <= -10 ? OPADPO_ECUDA : OPADPO_EINVAL;
      }
    } else {
      GemmArgs a;
      a.M = N;
      a.N = V;
      a.K1 = d;
      a.A1 = {hrows, d};
      a.B1 = {lm_head, d};
      a.epilogue = EPI_F32_PARTIAL;
      sp = gemm_effective_splits(d, 0, 2);
      a.splits = sp;
      a.C = part;
      a.ldc = V;
      EN_CHECK(gemm_checked(a, s));
    }
    if (do_sample) {
      EN_CHECK(sample_pick(part, sp, N, V, sampler->top_k, sampler->top_p, sampler->temperature, sampler->seed, c.eos_id,
                           c.pad_id, sup, n_sup, done, next_ids, out_ids + Q, static_cast<int64_t>(Q) + max_new, cur, s));
    } else {
      EN_CHECK(greedy_pick(part, sp, N, V, c.eos_id, c.pad_id, sup, n_sup, done, next_ids, out_ids + Q,
                           static_cast<int64_t>(Q) + max_new, cur, s));
    }
    EN_CHECK(advance_counter(cur, s));
    return 0;
  };
  // first token from the prefill's last position
  EN_CHECK(rmsnorm_fwd(x_prefill_out, final_norm, hdn, N, d, c.rms_eps, last_rows, s));
  EN_CHECK(head_and_pick(hdn));

  const float scale = 1.f / sqrtf(static_cast<float>(hd));
  auto decode_step = [&]() -> int {
    // token fed at this step was written at column cur-1; its absolute position is S0 + cur - 1
    EN_CHECK(embed_rows(next_ids, embed, xa, N, d, V, s));
    if (c.n_layers > 0) EN_CHECK(rmsnorm_fwd(xa, (*LL)[0].ln1, hdn, N, d, c.rms_eps, nullptr, s));
    bf16* xc = xa;
    bf16* xn = xb;
    for (int l = 0; l < c.n_layers; ++l) {
      const LlamaLayer& L = (*LL)[l];
      int sp = 1;
      EN_CHECK(linear_partial(L.qkv, hdn, 0, &sp));
      EN_CHECK(partial_qkv_rope(part, sp, qkvd, Nl, d, hd, rope_cs, S0 - 1, cur, s));
      EN_CHECK(decode_attention(qkvd, kc[l], vc[l], kmask, od, N, H, hd, S0 - 1, cur, S_max, scale, s));
      EN_CHECK(linear_partial(L.o, od, 0, &sp));
      EN_CHECK(partial_residual_rmsnorm(part, sp, xc, x1d, L.ln2, hdn, N, d, c.rms_eps, s));
      EN_CHECK(linear_partial(L.gu, hdn, 0, &sp));
      EN_CHECK(partial_swiglu(part, sp, actd, N, F, s));
      EN_CHECK(linear_partial(L.down, actd, 0, &sp));
      const bf16* next_w = l + 1 < c.n_layers ? (*LL)[l + 1].ln1 : final_norm;
      EN_CHECK(partial_residual_rmsnorm(part, sp, x1d, xn, next_w, hdn, N, d, c.rms_eps, s));
      bf16* t = xc;
      xc = xn;
      xn = t;
    }
    if (c.n_layers == 0) EN_CHECK(rmsnorm_fwd(xa, final_norm, hdn, N, d, c.rms_eps, nullptr, s));
    EN_CHECK(head_and_pick(hdn));
    return 0;
  };
  if (max_new > 1) {
    EN_CHECK(decode_step());  // step 1 eagerly (first-use kernel attributes are set outside of capture)
    const int remaining = max_new - 2;
    if (remaining > 0) {
      cudaStream_t cs = nullptr;
      cudaEvent_t ev_in = nullptr, ev_out = nullptr;
      if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming) != cudaSuccess) {
        set_error("generate: stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return OPADPO_ECUDA;
      }
      cudaEventRecord(ev_in, s);
      cudaStreamWaitEvent(cs, ev_in, 0);
      cudaStream_t caller = s;
      s = cs;  // the lambdas above capture `s` by reference
      cudaGraph_t graph = nullptr;
      cudaGraphExec_t exec = nullptr;
      int rc = 0;
      cudaError_t ce = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
      if (ce == cudaSuccess) {
        rc = decode_step();
        ce = cudaStreamEndCapture(cs, &graph);
      }
      if (rc == 0 && ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, graph, 0);
      if (rc == 0 && ce == cudaSuccess) {
        for (int i = 0; i < remaining && ce == cudaSuccess; ++i) ce = cudaGraphLaunch(exec, cs);
        count_launch(0);
      }
      cudaEventRecord(ev_out, cs);
      cudaStreamWaitEvent(caller, ev_out, 0);
      s = caller;
      if (exec) cudaGraphExecDestroy(exec);
      if (graph) cudaGraphDestroy(graph);
      cudaStreamSynchronize(cs);
      cudaEventDestroy(ev_in);
      cudaEventDestroy(ev_out);
      cudaStreamDestroy(cs);
      if (rc != 0) return rc;
      if (ce != cudaSuccess) {
        set_error("generate: CUDA graph decode failed: %s", cudaGetErrorString(ce));
        cudaGetLastError();
        return OPADPO_ECUDA;
      }
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("generate: %s", cudaGetErrorString(e));
    return OPADPO_ECUDA;
  }
  return 0;
}

}  // namespace opadpo
