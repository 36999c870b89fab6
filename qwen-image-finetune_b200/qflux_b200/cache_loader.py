"""Cached-embedding loader feeding the fused training step.

Reads the cache the reference's `EmbeddingCacheManager` writes (/root/reference/src/qflux/data/cache_manager.py:46-125):

    <cache_root>/metadata/<main_hash>.json      {"version": .., "<embedding key>": "<hash>", ..., "img_shapes": [[C,H,W], ...]}
    <cache_root>/<embedding key>/<hash>.pt      one fp16 tensor per sample and key (torch.save)

and collates like `collate_fn` / `pad_to_max_shape` (data/dataset.py:641-695, utils/tools.py:399-425): tensors are right-padded
with zeros to the batch maximum and stacked.  What is B200-first here: samples are read by a background thread, collated
straight into PINNED staging buffers (two, ping-pong) and uploaded with non-blocking copies on a side stream, so the step never
waits on the file system or on pageable-memory copies; `img_shapes` come back converted to the latent-patch units the step
classes take (convert_img_shapes_to_latent, trainer/qwen_image_edit_trainer.py:557-577).  Masks needed by the multi-resolution
recipes (`prompt_embeds_mask`) are derived from the un-padded lengths, not stored.
"""
from __future__ import annotations

import glob
import json
import os
import queue
import threading

import torch


def img_shapes_to_latent(shapes_px, vae_scale: int = 8, patch: int = 2):
    """[(C, H, W), ...] pixel shapes of one sample -> [(1, H/16, W/16), ...] latent-patch shapes."""
    return [(1, int(s[1]) // vae_scale // patch, int(s[2]) // vae_scale // patch) for s in shapes_px]


def pad_stack(tensors, out=None):
    """pad_to_max_shape: right-pad every dimension with zeros to the maximum and stack; `out` (pinned) is reused when it fits."""
    shape = [max(s) for s in zip(*[t.shape for t in tensors])]
    full = (len(tensors), *shape)
    if out is None or tuple(out.shape) != full or out.dtype != tensors[0].dtype:
        out = torch.zeros(full, dtype=tensors[0].dtype)
        if torch.cuda.is_available():
            out = out.pin_memory()
    else:
        out.zero_()
    for i, t in enumerate(tensors):
        out[(i, *[slice(0, d) for d in t.shape])] = t
    return out


class CachedEmbeddingLoader:
    def __init__(self, cache_root: str, batch_size: int, keys=("image_latents", "control_latents", "prompt_embeds"),
                 device="cuda", shuffle: bool = True, seed: int = 1234, drop_last: bool = True, prefetch: int = 2,
                 rank: int = 0, world_size: int = 1):
        self.root, self.bs, self.keys, self.device = str(cache_root), batch_size, tuple(keys), torch.device(device)
        self.shuffle, self.seed, self.drop_last, self.prefetch = shuffle, seed, drop_last, prefetch
        metas = sorted(glob.glob(os.path.join(self.root, "metadata", "*.json")))
        if not metas:
            raise FileNotFoundError(f"no cache metadata under {self.root}/metadata (EmbeddingCacheManager.exist would be False)")
        # data parallel: disjoint strided shards.  Every rank must yield the SAME number of batches (a rank with one batch more would
        # block forever in the gradient all-reduce), so the global list is first cut (drop_last) or wrapped around (like accelerate's
        # even_batches) to a multiple of world_size * batch_size
        unit = world_size * batch_size
        if world_size > 1 and len(metas) % unit:
            if drop_last:
                metas = metas[: len(metas) - len(metas) % unit]
            else:
                metas = metas + metas[: unit - len(metas) % unit]
        if not metas:
            raise ValueError(f"{self.root}: fewer cached samples than world_size * batch_size = {unit}")
        self.samples = metas[rank::world_size]
        self.epoch = 0
        self._stage = [dict(), dict()]  # two sets of pinned staging buffers: one being filled while the other uploads

    def __len__(self):
        n = len(self.samples)
        return n // self.bs if self.drop_last else -(-n // self.bs)

    # ---------------------------------------------------------------------------------------------------- host side
    def _load_sample(self, meta_path):
        with open(meta_path) as f:
            meta = json.load(f)
        out = {}
        for k in self.keys:
            out[k] = torch.load(os.path.join(self.root, k, f"{meta[k]}.pt"), map_location="cpu", weights_only=False)
        if "img_shapes" in meta:
            out["img_shapes"] = img_shapes_to_latent(meta["img_shapes"])
        return out

    def _collate(self, samples, stage):
        batch = {}
        for k in self.keys:
            ts = [s[k][0] if s[k].ndim == 3 and s[k].shape[0] == 1 else s[k] for s in samples]  # cached with a leading batch dim of 1
            stage[k] = pad_stack(ts, stage.get(k))
            batch[k] = stage[k]
        if "prompt_embeds" in self.keys:
            lens = torch.tensor([(s["prompt_embeds"][0] if s["prompt_embeds"].ndim == 3 else s["prompt_embeds"]).shape[0] for s in samples])
            batch["prompt_embeds_mask"] = (torch.arange(batch["prompt_embeds"].shape[1])[None, :] < lens[:, None]).to(torch.int64)
        if "img_shapes" in samples[0]:
            batch["img_shapes"] = [s["img_shapes"] for s in samples]
        return batch

    def _order(self):
        idx = list(range(len(self.samples)))
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            idx = torch.randperm(len(idx), generator=g).tolist()
        return idx

    def _producer(self, q, stop):
        idx = self._order()
        nb = len(self)
        for b in range(nb):
            if stop.is_set():
                break
            chunk = idx[b * self.bs:(b + 1) * self.bs]
            samples = [self._load_sample(self.samples[i]) for i in chunk]
            q.put((b, samples))
        q.put(None)

    # ---------------------------------------------------------------------------------------------------- iteration
    def __iter__(self):
        q, stop = queue.Queue(maxsize=self.prefetch), threading.Event()
        th = threading.Thread(target=self._producer, args=(q, stop), daemon=True)
        th.start()
        cuda = self.device.type == "cuda"
        side = torch.cuda.Stream(self.device) if cuda else None
        done = [None, None]  # upload-finished events guarding the re-use of each staging set
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                b, samples = item
                slot = b & 1
                if cuda and done[slot] is not None:
                    done[slot].synchronize()  # the copies out of this staging set (two batches ago) have finished
                host = self._collate(samples, self._stage[slot])
                if not cuda:
                    yield {k: (v.clone() if torch.is_tensor(v) else v) for k, v in host.items()}
                    continue
                dev = {}
                with torch.cuda.stream(side):
                    for k, v in host.items():
                        dev[k] = v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v
                    done[slot] = torch.cuda.Event()
                    done[slot].record(side)
                torch.cuda.current_stream().wait_event(done[slot])  # consumers on the current stream see complete tensors
                for v in dev.values():
                    if torch.is_tensor(v):
                        v.record_stream(torch.cuda.current_stream())
                yield dev
        finally:
            stop.set()
            self.epoch += 1
