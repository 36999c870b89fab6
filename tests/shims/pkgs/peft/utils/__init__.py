"""peft.utils.get_peft_model_state_dict / set_peft_model_state_dict for LoRA adapters (bias="none")."""


def get_peft_model_state_dict(model, state_dict=None, adapter_name="default", **kwargs):
    """Adapter tensors only, adapter name removed from the keys: `...lora_A.<adapter>.weight` -> `...lora_A.weight`."""
    if state_dict is None:
        state_dict = model.state_dict()
    out = {k: v for k, v in state_dict.items() if "lora_" in k and f".{adapter_name}" in k}
    return {k.replace(f".{adapter_name}", ""): v for k, v in out.items()}


def set_peft_model_state_dict(model, peft_model_state_dict, adapter_name="default", **kwargs):
    sd = {}
    for k, v in peft_model_state_dict.items():
        if "lora_" in k:
            head, _, tail = k.rpartition(".")  # ...lora_A | weight
            if not head.endswith(f".{adapter_name}"):
                k = f"{head}.{adapter_name}.{tail}"
        sd[k] = v
    res = model.load_state_dict(sd, strict=False)
    unexpected = list(res[1] if isinstance(res, tuple) else res.unexpected_keys)
    if unexpected:
        raise ValueError(f"unexpected adapter keys: {unexpected[:4]}")
    return res
