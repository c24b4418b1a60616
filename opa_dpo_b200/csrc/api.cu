// C ABI (include/opadpo_b200.h) — operator-level entry points.
#include <cstdio>
#include <cstring>

#include "../../include/opadpo_b200.h"
#include "common.h"
#include "attention.h"
#include "gemm.h"

namespace opadpo {
thread_local char g_last_error[1024] = {0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
}  // namespace opadpo

using namespace opadpo;

extern "C" {

const char* opadpo_last_error(void) { return g_last_error; }
int opadpo_abi_version(void) { return 1; }

int opadpo_gemm_n_tiles(int32_t N) { return gemm_num_n_tiles(N); }

int opadpo_op_gemm(const opadpo_gemm_desc* d, void* stream) {
  if (!d) {
    set_error("opadpo_op_gemm: null descriptor");
    return OPADPO_EINVAL;
  }
  GemmArgs a;
  a.M = d->M;
  a.N = d->N;
  a.K1 = d->K1;
  a.K2 = d->K2;
  a.a_mn_major = d->a_mn_major != 0;
  a.b_mn_major = d->b_mn_major != 0;
  a.A1 = {static_cast<const __nv_bfloat16*>(d->A1), d->lda1};
  a.B1 = {static_cast<const __nv_bfloat16*>(d->B1), d->ldb1};
  a.A2 = {static_cast<const __nv_bfloat16*>(d->A2), d->lda2};
  a.B2 = {static_cast<const __nv_bfloat16*>(d->B2), d->ldb2};
  a.a2_group_n = d->a2_group_n;
  a.a2_group_stride = d->a2_group_stride;
  a.splits = d->splits;
  a.epilogue = d->epilogue;
  a.alpha = d->alpha;
  a.C = d->C;
  a.ldc = d->ldc;
  a.bias = static_cast<const __nv_bfloat16*>(d->bias);
  a.residual = static_cast<const __nv_bfloat16*>(d->residual);
  a.ldr = d->ldr;
  a.act = d->act;
  a.labels = d->labels;
  a.lse_partials = static_cast<float4*>(d->lse_partials);
  a.label_logit = d->label_logit;
  a.lse = d->lse;
  a.dlp = d->dlp;
  a.max_ctas = d->max_ctas;
  int rc = gemm_launch(a, static_cast<cudaStream_t>(stream));
  if (rc != 0) {
    set_error("%s", gemm_last_error());
    return rc <= -10 ? OPADPO_ECUDA : OPADPO_EINVAL;
  }
  return OPADPO_OK;
}


static AttnArgs to_attn(const opadpo_attn_desc* d) {
  AttnArgs a;
  a.N = d->N; a.S = d->S; a.H = d->H; a.HD = d->HD;
  a.causal = d->causal != 0;
  a.scale = d->scale;
  a.q = static_cast<const __nv_bfloat16*>(d->q);
  a.k = static_cast<const __nv_bfloat16*>(d->k);
  a.v = static_cast<const __nv_bfloat16*>(d->v);
  a.ld_qkv = d->ld_qkv;
  a.o = static_cast<__nv_bfloat16*>(d->o);
  a.ld_o = d->ld_o;
  a.lse = d->lse;
  a.kbits = d->kbits;
  a.dO = static_cast<const __nv_bfloat16*>(d->dO);
  a.ld_do = d->ld_do;
  a.delta = d->delta;
  a.dq = static_cast<__nv_bfloat16*>(d->dq);
  a.dk = static_cast<__nv_bfloat16*>(d->dk);
  a.dv = static_cast<__nv_bfloat16*>(d->dv);
  a.ld_dqkv = d->ld_dqkv;
  return a;
}

int opadpo_op_attn_fwd(const opadpo_attn_desc* d, void* stream) {
  if (!d) { set_error("opadpo_op_attn_fwd: null descriptor"); return OPADPO_EINVAL; }
  int rc = attn_forward(to_attn(d), static_cast<cudaStream_t>(stream));
  if (rc != 0) { set_error("%s", attn_last_error()); return rc <= -10 ? OPADPO_ECUDA : OPADPO_EINVAL; }
  return OPADPO_OK;
}

int opadpo_op_attn_bwd(const opadpo_attn_desc* d, void* stream) {
  if (!d) { set_error("opadpo_op_attn_bwd: null descriptor"); return OPADPO_EINVAL; }
  int rc = attn_backward(to_attn(d), static_cast<cudaStream_t>(stream));
  if (rc != 0) { set_error("%s", attn_last_error()); return rc <= -10 ? OPADPO_ECUDA : OPADPO_EINVAL; }
  return OPADPO_OK;
}

int opadpo_op_pack_mask(const uint8_t* mask, uint64_t* bits, int32_t N, int32_t S, void* stream) {
  int rc = attn_pack_mask(mask, bits, N, S, static_cast<cudaStream_t>(stream));
  if (rc != 0) { set_error("pack_mask launch failed"); return OPADPO_ECUDA; }
  return OPADPO_OK;
}

}  // extern "C"
