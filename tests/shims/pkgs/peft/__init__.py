"""Minimal `peft` stand-in (tests only; see tests/shims/README.md): LoRA on nn.Linear with PEFT's published semantics.

  y = base_layer(x) + lora_B(lora_A(dropout(x.to(lora_A.weight.dtype)))) * (lora_alpha / r)
  keys: <module>.base_layer.{weight,bias}, <module>.lora_A.<adapter>.weight [r, in], <module>.lora_B.<adapter>.weight [out, r]
  init_lora_weights: True -> kaiming_uniform_(A, a=sqrt(5)), B = 0;  "gaussian" -> normal_(A, std=1/r), B = 0
  target_modules: list -> `key in targets or key.endswith("." + t)`;  str -> re.fullmatch(target_modules, key)
  injected (not get_peft_model) adapters take the base layer's dtype and device (no fp32 autocast of the adapter).
"""
__version__ = "0.0.0"
from .tuners.lora import Linear, LoraConfig, LoraLayer, inject_adapter_in_model  # noqa: F401
from .utils import get_peft_model_state_dict, set_peft_model_state_dict  # noqa: F401
