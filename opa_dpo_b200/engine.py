"""Python handle on the C-ABI engine (include/opadpo_b200.h, engine level).

Device memory, streams and tensors come from PyTorch; every computation is a call into libopadpo_b200.so.
There is no fallback: constructing an Engine without the library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, asdict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import Config, OpadpoError

POLICY, REF, BASE = 0, 1, -1          # adapter ids: lora_policy / lora_ref_policy (qlora_model.py:21) / bare base model
DT_BF16, DT_F32 = 0, 1
VT = "model.vision_tower.vision_tower.vision_model."


@dataclass
class EngineConfig:
    """Mirrors opadpo_config. Defaults are LLaVA-1.5-7B + CLIP-ViT-L/14-336 + LoRA r=256 a=512 (train_opa_dpo.sh:35-37)."""
    d_model: int = 4096
    n_layers: int = 32
    n_heads: int = 32
    d_ffn: int = 11008
    vocab: int = 32000
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_positions: int = 4096
    v_d: int = 1024
    v_layers: int = 23          # executed layers: mm_vision_select_layer=-2 of 24
    v_heads: int = 16
    v_ffn: int = 4096
    v_image: int = 336
    v_patch: int = 14
    v_eps: float = 1e-5
    lora_r: int = 256
    lora_alpha: float = 512.0
    pad_id: int = 0
    eos_id: int = 2
    image_token: int = -200

    @property
    def n_patches(self) -> int:
        return (self.v_image // self.v_patch) ** 2

    def to_c(self) -> Config:
        c = Config()
        for k, v in asdict(self).items():
            setattr(c, k, v)
        return c

    @staticmethod
    def llava_13b() -> "EngineConfig":
        return EngineConfig(d_model=5120, n_layers=40, n_heads=40, d_ffn=13824)


class _DevArray:
    """Zero-copy view of engine-owned device memory for torch.as_tensor."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def linear_modules(cfg: EngineConfig) -> Dict[str, Tuple[int, int]]:
    """LoRA target module path -> (out_features, in_features): every nn.Linear but lm_head (opa_train.py:177-190)."""
    d, f, vd, vf = cfg.d_model, cfg.d_ffn, cfg.v_d, cfg.v_ffn
    s: Dict[str, Tuple[int, int]] = {}
    for i in range(cfg.n_layers):
        p = f"model.layers.{i}."
        for m in ("q_proj", "k_proj", "v_proj", "o_proj"):
            s[p + "self_attn." + m] = (d, d)
        s[p + "mlp.gate_proj"] = (f, d)
        s[p + "mlp.up_proj"] = (f, d)
        s[p + "mlp.down_proj"] = (d, f)
    s["model.mm_projector.0"] = (d, vd)
    s["model.mm_projector.2"] = (d, d)
    for i in range(cfg.v_layers):
        p = f"{VT}encoder.layers.{i}."
        for m in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[p + "self_attn." + m] = (vd, vd)
        s[p + "mlp.fc1"] = (vf, vd)
        s[p + "mlp.fc2"] = (vd, vf)
    return s


def required_tensor_names(cfg: EngineConfig) -> List[str]:
    """Every base tensor the engine executes (what a base checkpoint has to provide): embeddings, norms, every linear of
    linear_modules (+ biases of the projector and the CLIP tower), the CLIP embeddings and layer norms of the executed layers."""
    names = ["model.embed_tokens.weight", "lm_head.weight", "model.norm.weight"]
    for i in range(cfg.n_layers):
        names += [f"model.layers.{i}.input_layernorm.weight", f"model.layers.{i}.post_attention_layernorm.weight"]
    for m in linear_modules(cfg):
        names.append(m + ".weight")
        if "vision_model" in m or "mm_projector" in m:
            names.append(m + ".bias")
    names += [VT + "embeddings.class_embedding", VT + "embeddings.patch_embedding.weight",
              VT + "embeddings.position_embedding.weight", VT + "pre_layrnorm.weight", VT + "pre_layrnorm.bias"]
    for i in range(cfg.v_layers):
        for ln in ("layer_norm1", "layer_norm2"):
            names += [f"{VT}encoder.layers.{i}.{ln}.weight", f"{VT}encoder.layers.{i}.{ln}.bias"]
    return names


def check_call_shapes(cfg: EngineConfig, images, queries, qmask, responses: Sequence = (), img_mask=None, what: str = "forward_logprobs"):
    """The C ABI takes plain pointers + sizes, so an array of the wrong extent would be read out of bounds on the device:
    every array's shape is checked against (b, Q, R) and the configured image size before its pointer crosses the boundary.
    (The reference fails in the first torch op that meets the mismatch; here it is an OpadpoError naming the argument.)"""
    if queries.dim() != 2:
        raise OpadpoError(f"{what}: queries must be [b, Q], got {tuple(queries.shape)}")
    b, Q = queries.shape
    if b < 1 or Q < 1:
        raise OpadpoError(f"{what}: empty batch (queries {tuple(queries.shape)})")
    want_img = (b, 3, cfg.v_image, cfg.v_image)
    if tuple(images.shape) != want_img:
        raise OpadpoError(f"{what}: images must be {want_img} (one {cfg.v_image}x{cfg.v_image} image per query), got {tuple(images.shape)}")
    if tuple(qmask.shape) != (b, Q):
        raise OpadpoError(f"{what}: the query attention mask must be [{b}, {Q}], got {tuple(qmask.shape)}")
    if responses:
        R = responses[0].shape[-1]
        for i, r in enumerate(responses):
            if tuple(r.shape) != (b, R):
                raise OpadpoError(f"{what}: response {i} must be [{b}, {R}] like the first one, got {tuple(r.shape)}")
    if img_mask is not None and tuple(img_mask.shape) != (b, cfg.n_patches):
        raise OpadpoError(f"{what}: the image-token mask must be [{b}, {cfg.n_patches}], got {tuple(img_mask.shape)}")


class Engine:
    def __init__(self, cfg: EngineConfig, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise OpadpoError("opa_dpo_b200.Engine needs a CUDA device: the hot path has no CPU implementation")
        self.cfg = cfg
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.lib = _lib.load()
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            c = cfg.to_c()
            _lib.check(self.lib.opadpo_engine_create(ctypes.byref(c), ctypes.byref(self._h)), "opadpo_engine_create")
        self._views = None
        self._resp_ptrs = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.opadpo_engine_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_tensor(self, name: str, t: torch.Tensor):
        t = t.detach()
        if t.dtype not in (torch.bfloat16, torch.float32):
            t = t.float()
        t = t.to(self.device).contiguous()
        if t.dim() == 1:
            rows, cols = 1, t.numel()
        else:
            rows, cols = t.shape[0], t.numel() // t.shape[0]
        dt = DT_BF16 if t.dtype == torch.bfloat16 else DT_F32
        _lib.check(self.lib.opadpo_engine_load_tensor(self._h, name.encode(), t.data_ptr(), dt, rows, cols, _stream()),
                   f"load_tensor({name})")
        torch.cuda.current_stream().synchronize()   # t may be a temporary

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        for name, t in sd.items():
            try:
                self.load_tensor(name, t)
            except _lib.OpadpoUnknownName:      # shape / dtype errors on tensors the engine holds always raise
                if strict:
                    raise

    def load_lora(self, adapter: int, module: str, A: torch.Tensor, B: torch.Tensor):
        A = A.detach().to(self.device, torch.float32).contiguous()
        B = B.detach().to(self.device, torch.float32).contiguous()
        r = self.cfg.lora_r
        if A.shape[0] != r or B.shape[1] != r:
            raise OpadpoError(f"load_lora({module}): rank mismatch A{tuple(A.shape)} B{tuple(B.shape)} r={r}")
        _lib.check(self.lib.opadpo_engine_load_lora(self._h, adapter, module.encode(), A.data_ptr(), B.data_ptr(), DT_F32,
                                                    _stream()), f"load_lora({module})")
        torch.cuda.current_stream().synchronize()

    def load_adapter(self, adapter: int, lora: Dict[str, Tuple[torch.Tensor, torch.Tensor]]):
        for m, (A, B) in lora.items():
            self.load_lora(adapter, m, A, B)

    def export_lora(self, adapter: int, module: str) -> Tuple[torch.Tensor, torch.Tensor]:
        o, i = self.lora_shape(module)
        A = torch.empty((self.cfg.lora_r, i), device=self.device, dtype=torch.float32)
        B = torch.empty((o, self.cfg.lora_r), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.opadpo_engine_export_lora(self._h, adapter, module.encode(), A.data_ptr(), B.data_ptr(),
                                                      _stream()), f"export_lora({module})")
        return A, B

    def lora_shape(self, module: str) -> Tuple[int, int]:
        i, o = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.opadpo_engine_lora_shape(self._h, module.encode(), ctypes.byref(i), ctypes.byref(o)),
                   f"lora_shape({module})")
        return o.value, i.value

    def lora_offsets(self, module: str) -> Tuple[int, int]:
        a, b = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.opadpo_engine_lora_offsets(self._h, module.encode(), ctypes.byref(a), ctypes.byref(b)),
                   f"lora_offsets({module})")
        return a.value, b.value

    def set_trainable(self, adapter: int = POLICY):
        _lib.check(self.lib.opadpo_engine_set_trainable(self._h, adapter), "set_trainable")
        self._views = None

    def buffers(self) -> Dict[str, torch.Tensor]:
        """Zero-copy fp32 views of the trainable adapter's flat bucket: master [n_total]; grad/exp_avg/exp_avg_sq
        [n_train] (the CLIP-tower LoRA tail never receives gradients)."""
        if self._views is None:
            ptrs = [ctypes.c_void_p() for _ in range(4)]
            nt, ntr = ctypes.c_int64(), ctypes.c_int64()
            _lib.check(self.lib.opadpo_engine_buffers(self._h, *[ctypes.byref(p) for p in ptrs], ctypes.byref(nt),
                                                      ctypes.byref(ntr)), "engine_buffers")
            with torch.cuda.device(self.device):
                mk = lambda p, n: torch.as_tensor(_DevArray(p.value, n, "<f4"), device=self.device)
                self._views = dict(master=mk(ptrs[0], nt.value), grad=mk(ptrs[1], ntr.value),
                                   exp_avg=mk(ptrs[2], ntr.value), exp_avg_sq=mk(ptrs[3], ntr.value),
                                   n_total=nt.value, n_train=ntr.value)
        return self._views

    def set_backward_callback(self, fn):
        """fn(layer:int) is called on the host inside backward() after each layer's kernels are enqueued (None: off)."""
        if fn is None:
            self._cb = _lib.LAYER_CB()   # NULL function pointer
        else:
            self._cb = _lib.LAYER_CB(lambda layer, _user: fn(int(layer)))
        _lib.check(self.lib.opadpo_engine_set_backward_callback(self._h, self._cb, None), "set_backward_callback")

    def adopt_grad_buffer(self, buf: torch.Tensor):
        """Accumulate the trainable adapter's gradients in `buf` (fp32, >= n_train elements, e.g. a symmetric-memory tensor
        every data-parallel rank maps) instead of the engine's own bucket; `buffers()["grad"]` then views `buf`."""
        if buf.dtype != torch.float32 or not buf.is_cuda or not buf.is_contiguous():
            raise _lib.OpadpoError("adopt_grad_buffer: a contiguous fp32 CUDA tensor is required")
        _lib.check(self.lib.opadpo_engine_adopt_grad_buffer(self._h, buf.data_ptr(), buf.numel()), "adopt_grad_buffer")
        self._grad_owner = buf      # keeps the allocation alive as long as the engine uses it
        self._views = None

    def layer_grad_range(self, layer: int) -> Tuple[int, int]:
        off, cnt = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.opadpo_engine_layer_grad_range(self._h, layer, ctypes.byref(off), ctypes.byref(cnt)),
                   "layer_grad_range")
        return off.value, cnt.value

    def set_option(self, name: str, value: int):
        _lib.check(self.lib.opadpo_engine_set_option(self._h, name.encode(), int(value)), f"set_option({name})")

    def sync_shadow(self):
        _lib.check(self.lib.opadpo_engine_sync_shadow(self._h, _stream()), "sync_shadow")

    # ------------------------------------------------------------------ hot path
    def forward_logprobs(self, adapter: int, images: torch.Tensor, queries: torch.Tensor, qmask: torch.Tensor,
                         responses: Sequence[torch.Tensor], temperature: float = 1.0, img_mask: Optional[torch.Tensor] = None,
                         save_slot: int = -1, want_entropy: bool = True):
        """-> (logprobs [K,b,R] f32, entropies [K,b,R] f32 or None). All inputs must already be on the device."""
        if len(responses) < 1:
            raise OpadpoError("forward_logprobs: at least one response array")
        check_call_shapes(self.cfg, images, queries, qmask, responses, img_mask)
        b, Q = queries.shape
        K, R = len(responses), responses[0].shape[1]
        images = images.to(torch.float32).contiguous()
        queries = queries.to(torch.int64).contiguous()
        qmask = qmask.to(torch.uint8).contiguous()
        resp = [r.to(torch.int64).contiguous() for r in responses]
        for t in [images, queries, qmask] + resp:
            if t.device != self.device:
                raise OpadpoError("forward_logprobs: inputs must live on the engine's device")
        if img_mask is not None:
            img_mask = img_mask.to(torch.uint8).contiguous()
        lp = torch.empty((K, b, R), device=self.device, dtype=torch.float32)
        ent = torch.empty((K, b, R), device=self.device, dtype=torch.float32) if want_entropy else None
        arr = (ctypes.c_void_p * K)(*[r.data_ptr() for r in resp])
        _lib.check(self.lib.opadpo_forward_logprobs(
            self._h, adapter, images.data_ptr(), b, queries.data_ptr(), qmask.data_ptr(),
            None if img_mask is None else img_mask.data_ptr(), Q, K, arr, R, float(temperature), lp.data_ptr(),
            None if ent is None else ent.data_ptr(), save_slot, _stream()), "opadpo_forward_logprobs")
        return lp, ent

    def backward(self, save_slot: int, dlogprobs: Optional[torch.Tensor], dentropies: Optional[torch.Tensor] = None):
        """d(loss)/d(logprobs) and/or d(loss)/d(entropies), both [K,b,R] -> LoRA gradients (accumulated)."""
        d = None if dlogprobs is None else dlogprobs.to(self.device, torch.float32).contiguous()
        if dentropies is None:
            if d is None:
                raise OpadpoError("backward: no incoming gradient")
            _lib.check(self.lib.opadpo_backward(self._h, save_slot, d.data_ptr(), _stream()), "opadpo_backward")
            return
        g = dentropies.to(self.device, torch.float32).contiguous()
        _lib.check(self.lib.opadpo_backward_ent(self._h, save_slot, None if d is None else d.data_ptr(), g.data_ptr(),
                                                _stream()), "opadpo_backward_ent")

    def zero_grad(self):
        _lib.check(self.lib.opadpo_zero_grad(self._h, _stream()), "opadpo_zero_grad")

    def optimizer_step(self, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                       max_grad_norm: float = 1.0, grad_scale: float = 1.0) -> torch.Tensor:
        norm = torch.empty((1,), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.opadpo_optimizer_step(self._h, lr, betas[0], betas[1], eps, weight_decay, max_grad_norm,
                                                  grad_scale, norm.data_ptr(), _stream()), "opadpo_optimizer_step")
        return norm

    def generate(self, adapter: int, images: torch.Tensor, queries: torch.Tensor, qmask: torch.Tensor, max_new_tokens: int,
                 do_sample: bool = False, top_k: int = 0, top_p: float = 1.0, temperature: float = 1.0, seed: int = 0,
                 suppress_tokens=None):
        check_call_shapes(self.cfg, images, queries, qmask, what="generate")
        b, Q = queries.shape
        sup = [int(t) for t in (suppress_tokens or [])]
        if len(sup) > 8:
            raise OpadpoError(f"at most 8 suppress_tokens (got {len(sup)})")
        sm = _lib.Sampler(int(do_sample), int(top_k), float(top_p), float(temperature), int(seed) & (2 ** 64 - 1), len(sup),
                          (ctypes.c_int32 * 8)(*sup))
        images = images.to(self.device, torch.float32).contiguous()
        queries = queries.to(self.device, torch.int64).contiguous()
        qmask = qmask.to(self.device, torch.uint8).contiguous()
        out = torch.empty((b, Q + max_new_tokens), device=self.device, dtype=torch.int64)
        _lib.check(self.lib.opadpo_generate(self._h, adapter, images.data_ptr(), b, queries.data_ptr(), qmask.data_ptr(), Q,
                                            max_new_tokens, ctypes.byref(sm), out.data_ptr(), _stream()), "opadpo_generate")
        return out

    def required_tensor_names(self) -> List[str]:
        return required_tensor_names(self.cfg)

    def launch_count(self) -> int:
        return int(self.lib.opadpo_launch_count())

    # ------------------------------------------------------------------ synthetic initialisation (bench / smoke)
    def random_init(self, seed: int = 1234, std: float = 0.02, lora_seeds=(11, 12)):
        """Seeded random-init weights of the configured architecture, generated on the device tensor by tensor
        (no checkpoints exist offline): N(0, std^2) linears/embeddings, norms = 1, LoRA A and B N(0, std^2)."""
        cfg = self.cfg
        g = torch.Generator(device=self.device).manual_seed(seed)

        def rnd(*shape):
            return (torch.randn(*shape, generator=g, device=self.device, dtype=torch.float32) * std).to(torch.bfloat16)

        ones = lambda n: torch.ones(n, device=self.device, dtype=torch.bfloat16)
        zeros = lambda n: torch.zeros(n, device=self.device, dtype=torch.bfloat16)
        self.load_tensor("model.embed_tokens.weight", rnd(cfg.vocab, cfg.d_model))
        self.load_tensor("lm_head.weight", rnd(cfg.vocab, cfg.d_model))
        self.load_tensor("model.norm.weight", ones(cfg.d_model))
        for name, (o, i) in linear_modules(cfg).items():
            self.load_tensor(name + ".weight", rnd(o, i))
            if "vision_model" in name or "mm_projector" in name:
                self.load_tensor(name + ".bias", rnd(o))
        for i in range(cfg.n_layers):
            self.load_tensor(f"model.layers.{i}.input_layernorm.weight", ones(cfg.d_model))
            self.load_tensor(f"model.layers.{i}.post_attention_layernorm.weight", ones(cfg.d_model))
        self.load_tensor(VT + "embeddings.class_embedding", rnd(cfg.v_d))
        self.load_tensor(VT + "embeddings.patch_embedding.weight", rnd(cfg.v_d, 3 * cfg.v_patch * cfg.v_patch))
        self.load_tensor(VT + "embeddings.position_embedding.weight", rnd(cfg.n_patches + 1, cfg.v_d))
        self.load_tensor(VT + "pre_layrnorm.weight", ones(cfg.v_d))
        self.load_tensor(VT + "pre_layrnorm.bias", zeros(cfg.v_d))
        for i in range(cfg.v_layers):
            for ln in ("layer_norm1", "layer_norm2"):
                self.load_tensor(f"{VT}encoder.layers.{i}.{ln}.weight", ones(cfg.v_d))
                self.load_tensor(f"{VT}encoder.layers.{i}.{ln}.bias", zeros(cfg.v_d))
        for adapter, s in zip((POLICY, REF), lora_seeds):
            ga = torch.Generator(device=self.device).manual_seed(s)
            for name, (o, i) in linear_modules(cfg).items():
                A = torch.randn(cfg.lora_r, i, generator=ga, device=self.device) * std
                B = torch.randn(o, cfg.lora_r, generator=ga, device=self.device) * std
                self.load_lora(adapter, name, A, B)
