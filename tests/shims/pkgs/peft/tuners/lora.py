import math
import re
from dataclasses import dataclass, field

import torch
import torch.nn as nn


@dataclass
class LoraConfig:
    """peft.LoraConfig: the fields BaseTrainer.add_lora_adapter sets (r, lora_alpha, init_lora_weights, target_modules) + peft's defaults
    for the ones the injection reads (lora_dropout 0, bias "none"); scaling = lora_alpha / r (no rank-stabilised variant)."""
    r: int = 8
    target_modules: list | str | None = None
    lora_alpha: float = 8
    lora_dropout: float = 0.0
    bias: str = "none"
    init_lora_weights: bool | str = True
    rank_pattern: dict = field(default_factory=dict)
    alpha_pattern: dict = field(default_factory=dict)
    peft_type: str = "LORA"

    def to_dict(self):
        return dict(self.__dict__)


class LoraLayer:
    """marker base class (peft.tuners.lora.LoraLayer)"""
    adapter_layer_names = ("lora_A", "lora_B")


class Linear(nn.Module, LoraLayer):
    """peft.tuners.lora.Linear: wraps `base_layer`; forward = base_layer(x) + lora_B(lora_A(dropout(x))) * scaling with the low-rank branch
    computed in the adapter weights' dtype and added to the base result (the bf16 rounding points of PEFT's eager forward);
    lora_A init: "gaussian" -> normal with std 1/r, default -> kaiming_uniform(a=sqrt(5)); lora_B = 0."""
    def __init__(self, base_layer: nn.Linear, adapter_name: str, r: int, lora_alpha: float, lora_dropout: float = 0.0,
                 init_lora_weights=True):
        super().__init__()
        self.base_layer = base_layer
        self.in_features, self.out_features = base_layer.in_features, base_layer.out_features
        self.r, self.lora_alpha, self.scaling = {}, {}, {}
        self.lora_dropout, self.lora_A, self.lora_B = nn.ModuleDict(), nn.ModuleDict(), nn.ModuleDict()
        self.active_adapters = [adapter_name]
        self.update_layer(adapter_name, r, lora_alpha, lora_dropout, init_lora_weights)

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def update_layer(self, adapter_name, r, lora_alpha, lora_dropout, init_lora_weights):
        assert r > 0
        self.r[adapter_name], self.lora_alpha[adapter_name] = r, lora_alpha
        self.lora_dropout[adapter_name] = nn.Dropout(p=lora_dropout) if lora_dropout > 0.0 else nn.Identity()
        self.lora_A[adapter_name] = nn.Linear(self.in_features, r, bias=False)
        self.lora_B[adapter_name] = nn.Linear(r, self.out_features, bias=False)
        self.scaling[adapter_name] = lora_alpha / r
        if init_lora_weights is True:
            nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
        elif isinstance(init_lora_weights, str) and init_lora_weights.lower() == "gaussian":
            nn.init.normal_(self.lora_A[adapter_name].weight, std=1 / r)
        elif init_lora_weights is not False:
            raise ValueError(f"shim: init_lora_weights={init_lora_weights!r} not restated")
        if init_lora_weights is not False:
            nn.init.zeros_(self.lora_B[adapter_name].weight)
        w = self.base_layer.weight  # _move_adapter_to_device_of_base_layer: device always, dtype when floating point
        for md in (self.lora_A, self.lora_B):
            md[adapter_name].to(w.device, dtype=w.dtype if w.dtype.is_floating_point else None)

    def set_adapter(self, names):
        names = [names] if isinstance(names, str) else list(names)
        for md in (self.lora_A, self.lora_B):
            for k, layer in md.items():
                layer.requires_grad_(k in names)
        self.active_adapters = names

    def scale_layer(self, scale):
        for a in self.active_adapters:
            self.scaling[a] *= scale

    def unscale_layer(self, scale=None):
        for a in self.active_adapters:
            self.scaling[a] = self.lora_alpha[a] / self.r[a] if scale is None else self.scaling[a] / scale

    def forward(self, x, *args, **kwargs):
        result = self.base_layer(x, *args, **kwargs)
        torch_result_dtype = result.dtype
        for a in self.active_adapters:
            if a not in self.lora_A:
                continue
            lora_A, lora_B = self.lora_A[a], self.lora_B[a]
            xx = x.to(lora_A.weight.dtype)
            result = result + lora_B(lora_A(self.lora_dropout[a](xx))) * self.scaling[a]
        return result.to(torch_result_dtype)


def _matches(target_modules, key):
    if isinstance(target_modules, str):
        return re.fullmatch(target_modules, key) is not None
    return key in target_modules or any(key.endswith(f".{t}") for t in target_modules)


def inject_adapter_in_model(peft_config: LoraConfig, model: nn.Module, adapter_name: str = "default", low_cpu_mem_usage=False):
    """Replace every matching nn.Linear by a LoRA Linear; freeze everything that is not an adapter weight."""
    found = False
    for key, mod in list(model.named_modules()):
        if not key or not _matches(peft_config.target_modules, key):
            continue
        if isinstance(mod, Linear):
            mod.update_layer(adapter_name, peft_config.rank_pattern.get(key, peft_config.r),
                             peft_config.alpha_pattern.get(key, peft_config.lora_alpha), peft_config.lora_dropout,
                             peft_config.init_lora_weights)
            found = True
            continue
        if not isinstance(mod, nn.Linear):
            continue
        parent_name, _, child = key.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        new = Linear(mod, adapter_name, peft_config.rank_pattern.get(key, peft_config.r),
                     peft_config.alpha_pattern.get(key, peft_config.lora_alpha), peft_config.lora_dropout, peft_config.init_lora_weights)
        setattr(parent, child, new)
        found = True
    if not found:
        raise ValueError(f"Target modules {peft_config.target_modules} not found in the base model.")
    for n, p in model.named_parameters():  # _mark_only_adapters_as_trainable
        if "lora_" not in n:
            p.requires_grad = False
    if not hasattr(model, "peft_config"):
        model.peft_config = {}
    model.peft_config[adapter_name] = peft_config
    return model
