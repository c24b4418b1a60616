"""opa_dpo_b200 — Blackwell-native (sm_100a) OPA-DPO hot path behind the reference's policy boundary.

csrc/      hand-written CUDA kernels + the C ABI (include/opadpo_b200.h) -> lib/libopadpo_b200.so
engine.py  ctypes handle on the engine-level C ABI
policy.py  mirror of opadpo/dpo_models/rl_models.py (AutoregressivePolicy) and the `base_model` surface
dpo.py     mirror of DPOTrainer.dpo_loss / compute_policy_loss / the optimisation step
ops.py     operator-level bindings used by the kernel parity tests
"""
from ._lib import OpadpoError, build_library, load  # noqa: F401
from .engine import BASE, POLICY, REF, Engine, EngineConfig  # noqa: F401

__all__ = ["OpadpoError", "build_library", "load", "Engine", "EngineConfig", "POLICY", "REF", "BASE"]
