"""Minimal `accelerate` stand-in (tests only; see tests/shims/README.md): the single-process behaviour of the Accelerator
methods the reference's step body touches (base_trainer.py:449-455, 518-536, 858-866)."""
import contextlib

import torch

__version__ = "0.0.0"


class Accelerator:
    """accelerate.Accelerator on ONE process without mixed precision: the methods the reference's step body and checkpointing touch
    (accumulate, backward, clip_grad_norm_, unwrap_model, prepare, gather, wait_for_everyone) with their single-process semantics."""
    def __init__(self, *args, gradient_accumulation_steps=1, **kwargs):
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.gradient_accumulation_steps = gradient_accumulation_steps
        self.sync_gradients = True
        self.is_main_process = True
        self.num_processes = 1
        self.process_index = 0

    @contextlib.contextmanager
    def accumulate(self, *models):
        yield

    def backward(self, loss, **kwargs):
        """Accelerator.backward: loss / gradient_accumulation_steps, then .backward()."""
        (loss / self.gradient_accumulation_steps).backward(**kwargs)

    def clip_grad_norm_(self, parameters, max_norm, norm_type=2):
        """Accelerator.clip_grad_norm_ (no mixed-precision scaler, no FSDP): torch.nn.utils.clip_grad_norm_."""
        return torch.nn.utils.clip_grad_norm_(parameters, max_norm, norm_type=norm_type)

    def unwrap_model(self, model, keep_fp32_wrapper=True):
        return getattr(model, "module", model)

    def wait_for_everyone(self):
        pass

    def gather(self, tensor):
        """Accelerator.gather on one process: the tensor itself, with a leading process dimension for 0-d inputs."""
        return tensor.reshape(1) if tensor.dim() == 0 else tensor

    def prepare(self, *objs):
        return objs if len(objs) != 1 else objs[0]
