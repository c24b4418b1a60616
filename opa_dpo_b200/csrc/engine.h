// The OPA-DPO engine: owns base weights, two LoRA adapter sets, the trainable adapter's flat fp32 bucket
// (masters | gradients | Adam moments) and all activations; runs the LLaVA-1.5 forward / LoRA backward / rollout
// entirely through the kernels in gemm.cu, attention.cu and elementwise.cu. Host logic only — no device code here.
//
// Reference being replaced (behaviour, not structure): AutoregressivePolicy.forward
// (opadpo/dpo_models/rl_models.py:75-144) down through LlavaLlamaForCausalLM.forward
// (llava_setup/llava_modifications.patch:12127-12202, 15022-15055) and the [UPSTREAM] HF LLaMA / CLIP / PEFT LoRA
// arithmetic; autograd backward of rl_trainer.py:162; optimiser of utils/trainer_utils.py:9-38.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/opadpo_b200.h"
#include "elementwise.h"

namespace opadpo {

struct LinearW {
  int in = 0, out = 0, groups = 1;
  bf16* W = nullptr;     // [out, in] (groups stacked along out)
  bf16* W_il = nullptr;  // gate|up only, merged frozen adapter + fuse_swiglu: the same rows interleaved in blocks of 128
  bf16* bias = nullptr;  // [out] or null
  int slot = -1;         // LoRA slot index, -1 = not a LoRA target
};

struct LoraSlot {
  int64_t offA = 0, offB = 0;  // element offsets into the adapter's flat buffer: A [groups*r, in], B [out, r]
  int in = 0, out = 0, groups = 1;
};

struct Adapter {
  bf16* shadow = nullptr;  // [n_total] bf16 values the GEMMs read
  float *master = nullptr, *grad = nullptr, *m = nullptr, *v = nullptr;  // trainable adapter only, [n_total]
  bool trainable = false;
  int step = 0;
};

struct LlamaLayer {
  LinearW qkv, o, gu, down;
  bf16 *ln1 = nullptr, *ln2 = nullptr;
};
struct ClipLayer {
  LinearW qkv, out, fc1, fc2;
  bf16 *ln1w = nullptr, *ln1b = nullptr, *ln2w = nullptr, *ln2b = nullptr;
};

struct LayerActs {
  bf16 *x = nullptr, *qkv = nullptr, *o = nullptr, *x1 = nullptr, *gu = nullptr;
  bf16 *tqkv = nullptr, *to = nullptr, *tgu = nullptr, *td = nullptr;
  bf16 *h1 = nullptr, *h2 = nullptr, *act = nullptr;  // normed inputs of qkv / gate|up and the SwiGLU output
  float* lse = nullptr;
};

// How token rows are laid out for one forward: N rows of S padded positions; packed rows carry K response segments of
// segR tokens behind a shared prefix of S0 positions (segR == 0: one response per row).
struct SeqLayout {
  int N = 0, S = 0, S0 = 0, segR = 0;
  const int32_t* pos = nullptr;     // RoPE position per row slot [N*S] (null: slot index)
  const uint64_t* kbits = nullptr;  // key validity bits [N, ceil(S/64)]
};

struct SaveCtx {
  bool valid = false;
  int adapter = 0, b = 0, K = 0, N = 0, Q = 0, R = 0, S = 0;
  float inv_temp = 1.f;
  std::vector<void*> owned;  // freed when the slot is released
  std::vector<LayerActs> layers;
  bf16* x_final = nullptr;  // input of the final norm [T, d]
  bf16* hn = nullptr;       // normed response rows [N*R, d]
  int32_t* labels = nullptr;
  float* lse_head = nullptr;
  float* ent_head = nullptr;  // unmasked per-row entropy of the head (backward of the entropy output)
  uint64_t* kbits = nullptr;
  SeqLayout lay;
  int32_t *src_idx = nullptr, *dst_idx = nullptr;  // response-predicting hidden rows (gather / gradient scatter)
  int32_t *img_pos = nullptr, *img_of_row = nullptr;
  bf16 *clip_feats = nullptr, *p_pre = nullptr, *p_act = nullptr, *t_p0 = nullptr, *t_p2 = nullptr;
};

struct ModuleRef {
  LinearW* lin = nullptr;
  int group = 0;
};

class Engine {
 public:
  explicit Engine(const opadpo_config& c);
  ~Engine();
  int init();

  int load_tensor(const char* name, const void* src, int dtype, int64_t rows, int64_t cols, cudaStream_t s);
  int load_lora(int adapter, const char* module, const void* A, const void* B, int dtype, cudaStream_t s);
  int export_lora(int adapter, const char* module, float* A_out, float* B_out, cudaStream_t s);
  int lora_shape(const char* module, int64_t* in_f, int64_t* out_f);
  int lora_offsets(const char* module, int64_t* offA, int64_t* offB);
  int set_trainable(int adapter);
  int buffers(float** master, float** grad, float** m, float** v, int64_t* n_total, int64_t* n_train);
  int sync_shadow(cudaStream_t s);

  int forward_logprobs(int adapter, const float* images, int b, const int64_t* queries, const uint8_t* qmask,
                       const uint8_t* img_mask, int Q, int K, const int64_t* const* responses, int R, float temperature,
                       float* out_lp, float* out_ent, int save_slot, cudaStream_t s);
  int backward(int save_slot, const float* dlogprobs, const float* dentropies, cudaStream_t s);
  int zero_grad(cudaStream_t s);
  int optimizer_step(float lr, float b1, float b2, float eps, float wd, float max_norm, float gscale, float* norm_out,
                     cudaStream_t s);
  int generate(int adapter, const float* images, int b, const int64_t* queries, const uint8_t* qmask, int Q, int max_new,
               const opadpo_sampler* sampler, int64_t* out_ids, cudaStream_t s);
  int set_option(const char* name, int64_t value);
  int set_backward_callback(opadpo_layer_cb cb, void* user);
  int layer_grad_range(int layer, int64_t* offset, int64_t* count);
  int adopt_grad_buffer(float* buffer, int64_t count);

 private:
  opadpo_config cfg;
  int hd = 0, v_hd = 0, P = 0, Kp = 0, r = 0;
  float lora_scale = 1.f;
  bool share_prefix = true;  // compute the image+query prefix once per sample for its K responses
  // decode-step linears on merged / adapter-free weights go through the weight-streaming GEMM (decode_gemm.cu); 0 = the
  // general kernel with split-K (A/B comparison, and batches above 128 sequences always take it)
  bool decode_swap_ab = true;
  bool merge_rollout = true;  // rollout runs on W + (alpha/r) B A merged once per generate() call
  // Inference forwards of a FROZEN adapter (the reference pass of the DPO step, dpo_trainer.py:333-335) run on decoder
  // weights merged once, W' = W + (alpha/r) B A (what peft's merge_adapter does), so they carry no LoRA GEMMs at all.
  // The cache is rebuilt when that adapter's tensors are reloaded. Costs one extra copy of the decoder weights.
  bool merge_frozen = false;
  // q|k|v projection with RoPE applied in the GEMM epilogue (EPI_BF16_ROPE) instead of a separate in-place pass
  bool fuse_rope = true;
  bool fuse_rope_bwd = true;  // with fuse_rope: inverse rotation of dq / dk inside the attention backward (option "fuse_rope_bwd")
  // LoRA weight-gradient GEMMs (dB += dy^T t, dA += u^T x) are off the backward's critical path: they run on a side stream,
  // where their tensor-bound tiles overlap the main stream's HBM-bound kernels (SwiGLU / RMSNorm backward, inverse RoPE) and
  // the tails of its persistent dgrad launches. Option "wgrad_stream" (default 1).
  // SwiGLU backward inside the down-projection dgrad's epilogue (option "fuse_swiglu_bwd"). Off by default: bit-compatible and
  // tested, but measured no faster (profiles/r02_bench_swiglu_bwd_epilogue_ab.log): 256 columns x (2 MUFU + ~20 instructions)
  // on the 4 epilogue warps outlast the 68 k-blocks of the mainloop, so the GEMM turns epilogue-bound and gives back what
  // the removed HBM pass saved.
  bool fuse_swiglu_bwd = false;
  bool wgrad_stream = true;
  cudaStream_t wg_stream = nullptr;
  cudaEvent_t ev_u = nullptr;         // main: this linear's u = dy B is enqueued (and, in stream order, its dy / t / x)
  cudaEvent_t ev_side[4] = {};        // side: the weight gradients of linear kind k (0 down, 1 gate|up, 2 o, 3 q|k|v) are done
  bool side_error = false;            // an event record / wait of the side-stream plumbing failed during this backward
  bool side_pending[4] = {};          // ev_side[k] carries a record nobody on the main stream has waited for since
  int ensure_wgrad_stream();
  void wait_side(int kind, cudaStream_t s);   // main stream waits for the side stream's readers of kind k's dy / u buffers
  // merged frozen adapter only: gate|up projection with SwiGLU in the GEMM epilogue (EPI_SWIGLU_BF16) over an interleaved
  // copy of the merged weight; bit-identical to projection + swiglu_fwd
  bool fuse_swiglu = true;
  int frozen_merged_adapter = -1;
  std::vector<LlamaLayer> frozen_merged;
  std::vector<void*> frozen_merged_allocs;
  void drop_frozen_merged();
  int ensure_frozen_merged(int adapter, cudaStream_t s);
  opadpo_layer_cb layer_cb = nullptr;  // host callback after each layer's backward kernels are enqueued (comm overlap)
  void* layer_cb_user = nullptr;
  // base weights
  bf16 *embed = nullptr, *final_norm = nullptr, *lm_head = nullptr;
  std::vector<LlamaLayer> layers;
  LinearW proj0, proj2;
  bf16 *v_patch_w = nullptr, *v_cls = nullptr, *v_pos = nullptr, *v_pre_w = nullptr, *v_pre_b = nullptr;
  std::vector<ClipLayer> vlayers;
  float2* rope_cs = nullptr;
  // LoRA
  std::vector<LoraSlot> slots;
  int64_t n_total = 0, n_train = 0;
  Adapter adapters[2];
  int trainable_adapter = -1;
  float* sumsq = nullptr;
  SaveCtx ctx[2];
  std::vector<void*> allocs;  // weight allocations

  // helpers
  template <typename T>
  int dalloc(T** p, int64_t count, std::vector<void*>* track);
  int alloc_linear(LinearW& L, int in, int out, int groups, bool bias, bool lora);
  bool resolve_module(const char* name, ModuleRef* ref);
  int copy_in(bf16* dst, const void* src, int dtype, int64_t n, cudaStream_t s);
  void release_slot(int slot, cudaStream_t s);

  const bf16* loraA(int adapter, const LinearW& L) const { return adapters[adapter].shadow + slots[L.slot].offA; }
  const bf16* loraB(int adapter, const LinearW& L) const { return adapters[adapter].shadow + slots[L.slot].offB; }

  struct RopeFuse {
    const int32_t* pos = nullptr;  // position of every row (null: row % S)
    int S = 0, cols = 0;           // cols: leading output columns that are rotated (q and k)
    bool applied = false;          // set by linear_fwd when the projection ran as ONE GEMM with the rotary epilogue
  };
  int linear_fwd(const LinearW& L, int adapter, const bf16* x, int64_t ldx, int64_t M, bf16* t, bf16* y, int64_t ldy, int act,
                 const bf16* residual, int64_t ldr, cudaStream_t s, RopeFuse* rope = nullptr);
  int linear_bwd(const LinearW& L, int adapter, const bf16* x, int64_t ldx, const bf16* t, const bf16* dy, int64_t ldy,
                 int64_t M, bf16* u, bf16* dx, int64_t lddx, const bf16* dres, int64_t lddres, cudaStream_t s, int side_kind = -1,
                 const bf16* swiglu_gu = nullptr);
  int clip_forward(int adapter, const float* images, int n, bf16* feats_out, cudaStream_t s, std::vector<void*>* tmp);
  int projector_forward(int adapter, const bf16* clip_feats, int64_t rows, bf16* pre, bf16* act, bf16* t0, bf16* t2,
                        bf16* feats, cudaStream_t s);
  int llama_layer_fwd(const LlamaLayer& L, int adapter, const SeqLayout& lay, const LayerActs& a, bf16* x_out,
                      cudaStream_t s);
  // W' = W + (alpha/r) B A for every decoder linear (rollout only: one plain weight-streaming GEMM per linear)
  int merge_lora(int adapter, std::vector<LlamaLayer>* merged, std::vector<void*>* track, cudaStream_t s);
};

}  // namespace opadpo
