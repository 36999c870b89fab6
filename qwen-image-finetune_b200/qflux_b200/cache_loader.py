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

Packed shards (`pack_cache`): the reference's layout costs one `torch.load` (un-pickling) and one small file per sample and key.
`pack_cache(cache_root)` rewrites it once into `<cache_root>/packed/<key>.<n>.bin` — the fp16 payloads back to back, sample by sample —
plus `packed/index.json` (per sample: shard, element offset and shape per key, `img_shapes`); the loader then maps the shards
(`numpy.memmap`) and copies each sample straight from the page cache into the pinned staging buffer.  The reference-format cache stays
the source of truth (the loader falls back to it when no pack is present; `pack_cache` is idempotent and keyed on the metadata list).
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


PACK_DIR, PACK_VERSION = "packed", 1


def pack_cache(cache_root: str, keys=("image_latents", "control_latents", "prompt_embeds"), shard_bytes: int = 1 << 30) -> str:
    """Rewrite a reference-format cache into packed fp16 shards (see the module docstring).  Returns the pack directory."""
    import numpy as np
    root = str(cache_root)
    metas = sorted(glob.glob(os.path.join(root, "metadata", "*.json")))
    if not metas:
        raise FileNotFoundError(f"no cache metadata under {root}/metadata")
    out_dir = os.path.join(root, PACK_DIR)
    os.makedirs(out_dir, exist_ok=True)
    files, shard_of, fill = {}, {k: 0 for k in keys}, {k: 0 for k in keys}
    open_shard = lambda k: open(os.path.join(out_dir, f"{k}.{shard_of[k]}.bin"), "wb")
    for k in keys:
        files[k] = open_shard(k)
    samples = []
    for mp in metas:
        with open(mp) as f:
            meta = json.load(f)
        entry = {"meta": os.path.basename(mp), "img_shapes": meta.get("img_shapes")}
        for k in keys:
            t = torch.load(os.path.join(root, k, f"{meta[k]}.pt"), map_location="cpu", weights_only=False).to(torch.float16).contiguous()
            nbytes = t.numel() * 2
            if fill[k] and fill[k] + nbytes > shard_bytes:
                files[k].close()
                shard_of[k], fill[k] = shard_of[k] + 1, 0
                files[k] = open_shard(k)
            entry[k] = {"shard": shard_of[k], "offset": fill[k] // 2, "shape": list(t.shape)}
            files[k].write(np.ascontiguousarray(t.numpy()).tobytes())
            fill[k] += nbytes
        samples.append(entry)
    for f in files.values():
        f.close()
    with open(os.path.join(out_dir, "index.json"), "w") as f:
        json.dump({"version": PACK_VERSION, "dtype": "float16", "keys": list(keys), "samples": samples}, f)
    return out_dir


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
                 rank: int = 0, world_size: int = 1, packed="auto", reference_batch_format: bool = False):
        """reference_batch_format: yield what the reference's own DataLoader yields for a cached sample — `img_shapes` in PIXEL units
        (C, H, W) and `cached = [True] * B` — so that the loader can feed `BaseTrainer.train_epoch` unchanged (its `training_step` ->
        `prepare_cached_embeddings` converts the shapes itself, base_trainer.py:457-466); the default hands out latent-patch shapes,
        the form the step classes take directly.
        packed: "auto" uses `<cache_root>/packed` when `pack_cache` has written it for exactly these samples and keys, True requires
        it, False always reads the reference's per-sample files."""
        self.root, self.bs, self.keys, self.device = str(cache_root), batch_size, tuple(keys), torch.device(device)
        self.shuffle, self.seed, self.drop_last, self.prefetch = shuffle, seed, drop_last, prefetch
        metas = sorted(glob.glob(os.path.join(self.root, "metadata", "*.json")))
        self._pack = self._open_pack(metas, packed)
        self.reference_batch_format = bool(reference_batch_format)
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
    def _open_pack(self, metas, packed):
        """{metadata file name: index entry} + lazily mapped shards, or None when the reference-format files are to be read."""
        idx_path = os.path.join(self.root, PACK_DIR, "index.json")
        if packed is False or not os.path.exists(idx_path):
            if packed is True:
                raise FileNotFoundError(f"{idx_path} is missing: run qflux_b200.cache_loader.pack_cache({self.root!r}) first")
            return None
        with open(idx_path) as f:
            idx = json.load(f)
        by_meta = {e["meta"]: e for e in idx["samples"]}
        ok = idx.get("version") == PACK_VERSION and set(self.keys) <= set(idx["keys"]) and all(os.path.basename(m) in by_meta for m in metas)
        if not ok:
            if packed is True:
                raise ValueError(f"{idx_path} does not cover this cache (stale pack, or other keys): re-run pack_cache")
            return None
        return {"by_meta": by_meta, "maps": {}}

    def _packed_tensor(self, key, ent):
        import numpy as np
        name = (key, ent["shard"])
        if name not in self._pack["maps"]:
            self._pack["maps"][name] = np.memmap(os.path.join(self.root, PACK_DIR, f"{key}.{ent['shard']}.bin"), dtype=np.float16, mode="r")
        n = 1
        for d in ent["shape"]:
            n *= d
        return torch.from_numpy(np.asarray(self._pack["maps"][name][ent["offset"]: ent["offset"] + n])).view(ent["shape"])

    def _load_sample(self, meta_path):
        if self._pack is not None:
            e = self._pack["by_meta"][os.path.basename(meta_path)]
            out = {k: self._packed_tensor(k, e[k]) for k in self.keys}
            if e.get("img_shapes") is not None:
                out["img_shapes"] = self._shapes(e["img_shapes"])
            return out
        with open(meta_path) as f:
            meta = json.load(f)
        out = {}
        for k in self.keys:
            out[k] = torch.load(os.path.join(self.root, k, f"{meta[k]}.pt"), map_location="cpu", weights_only=False)
        if "img_shapes" in meta:
            out["img_shapes"] = self._shapes(meta["img_shapes"])
        return out

    def _shapes(self, shapes_px):
        return [tuple(int(v) for v in s) for s in shapes_px] if self.reference_batch_format else img_shapes_to_latent(shapes_px)

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
        if self.reference_batch_format:
            batch["cached"] = [True] * len(samples)
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
