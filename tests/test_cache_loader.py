"""CachedEmbeddingLoader against a cache directory written in the reference's EmbeddingCacheManager layout
(cache_manager.py:46-93): per-key fp16 .pt files + metadata json with img_shapes; collate = pad_to_max_shape (tools.py:399-425)."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "qwen-image-finetune_b200"))


def _write_cache(root, n=5):
    g = torch.Generator().manual_seed(0)
    os.makedirs(os.path.join(root, "metadata"))
    sizes = [(512, 512), (320, 640), (640, 640), (512, 512), (320, 320)]
    truth = []
    for i in range(n):
        H, W = sizes[i % len(sizes)]
        L, T = (H // 16) * (W // 16), 5 + 3 * i
        data = dict(image_latents=torch.randn(1, L, 64, generator=g), control_latents=torch.randn(1, L, 64, generator=g),
                    prompt_embeds=torch.randn(1, T, 32, generator=g))
        meta = {"version": "1.0", "img_shapes": [[3, H, W], [3, H, W]]}
        for k, v in data.items():
            os.makedirs(os.path.join(root, k), exist_ok=True)
            h = f"{k[:2]}{i:04d}"
            torch.save(v.to(torch.float16), os.path.join(root, k, h + ".pt"))  # the reference caches fp16 (cache_manager.py:78)
            meta[k] = h
        with open(os.path.join(root, "metadata", f"main{i:04d}.json"), "w") as f:
            json.dump(meta, f)
        truth.append({k: v[0].to(torch.float16) for k, v in data.items()} | {"hw": (H // 16, W // 16)})
    return truth


def test_loader_pads_like_the_reference_collate(tmp_path):
    from qflux_b200.cache_loader import CachedEmbeddingLoader, img_shapes_to_latent
    truth = _write_cache(str(tmp_path))
    assert img_shapes_to_latent([[3, 512, 512], [3, 320, 640]]) == [(1, 32, 32), (1, 20, 40)]
    ld = CachedEmbeddingLoader(str(tmp_path), batch_size=2, device="cpu", shuffle=False, drop_last=False)
    assert len(ld) == 3
    batches = list(ld)
    assert [b["image_latents"].shape[0] for b in batches] == [2, 2, 1]
    b0 = batches[0]  # samples 0 (32x32 = 1024 tokens) and 1 (20x40 = 800 tokens)
    assert b0["image_latents"].shape == (2, 1024, 64) and b0["image_latents"].dtype == torch.float16
    assert torch.equal(b0["image_latents"][1, :800], truth[1]["image_latents"]) and b0["image_latents"][1, 800:].abs().max() == 0
    assert b0["prompt_embeds"].shape == (2, 8, 32) and torch.equal(b0["prompt_embeds"][0, :5], truth[0]["prompt_embeds"])
    assert b0["prompt_embeds_mask"].tolist() == [[1] * 5 + [0] * 3, [1] * 8]
    assert b0["img_shapes"] == [[(1, 32, 32), (1, 32, 32)], [(1, 20, 40), (1, 20, 40)]]
    # staging buffers are re-used every other batch: earlier batches must not be clobbered (CPU path hands out copies)
    assert torch.equal(batches[0]["control_latents"][0], truth[0]["control_latents"])
    # data-parallel sharding is disjoint and complete; shuffling is per-epoch deterministic
    r0 = CachedEmbeddingLoader(str(tmp_path), 1, device="cpu", shuffle=False, drop_last=False, rank=0, world_size=2)
    r1 = CachedEmbeddingLoader(str(tmp_path), 1, device="cpu", shuffle=False, drop_last=False, rank=1, world_size=2)
    # ... and EVEN: every rank yields the same number of batches (a rank with one batch more would hang in the gradient all-reduce);
    # without drop_last the global list wraps around (like accelerate's even_batches), with it the remainder is cut
    assert len(r0) == len(r1) == 3 and set(r0.samples) | set(r1.samples) == set(ld.samples) and len(set(r0.samples) & set(r1.samples)) == 1
    d0 = CachedEmbeddingLoader(str(tmp_path), 1, device="cpu", shuffle=False, drop_last=True, rank=0, world_size=2)
    d1 = CachedEmbeddingLoader(str(tmp_path), 1, device="cpu", shuffle=False, drop_last=True, rank=1, world_size=2)
    assert len(d0) == len(d1) == 2 and not set(d0.samples) & set(d1.samples)
    lens = {len(CachedEmbeddingLoader(str(tmp_path), 1, device="cpu", drop_last=dl, rank=r, world_size=4)) for r in range(4) for dl in (True,)}
    assert lens == {1}
    s = CachedEmbeddingLoader(str(tmp_path), 1, device="cpu", shuffle=True, drop_last=False, seed=3)
    e0 = [b["prompt_embeds"].shape[1] for b in s]
    e1 = [b["prompt_embeds"].shape[1] for b in s]
    assert sorted(e0) == sorted(e1) == [5, 8, 11, 14, 17] and e0 != e1


def test_loader_feeds_the_multi_resolution_step(tmp_path):
    """The batch dict is what QwenImageEditStep takes: a mixed-resolution batch goes through the pad-to-max recipe (emulated kernels)."""
    sys.path.insert(0, HERE)
    import emu_lib
    from qflux_b200 import lib
    restore = emu_lib.install(lib)
    try:
        from qflux_b200.cache_loader import CachedEmbeddingLoader
        from qflux_b200.qwen_model import QwenB200Config, QwenImageB200
        from qflux_b200.train_step import QwenImageEditStep
        _write_cache(str(tmp_path), n=2)
        m = QwenImageB200(QwenB200Config(num_layers=1, num_attention_heads=2, joint_attention_dim=32), device="cpu", _host_only=True)
        m.add_adapter(4, 4, b_std=0.05)
        batch = next(iter(CachedEmbeddingLoader(str(tmp_path), 2, device="cpu", shuffle=False)))
        loss = QwenImageEditStep(m, "attention_mask").train_step(batch)
        assert torch.isfinite(loss).all() and m._ws["kv_len"].tolist() == [5 + 2 * 1024 + 3, 8 + 2 * 800]
    finally:
        restore()


def test_packed_shards_reproduce_the_per_sample_cache(tmp_path):
    """`pack_cache` rewrites the reference-format cache into fp16 shards + an index; the loader must hand out exactly the same batches
    from the pack (memory-mapped) as from the per-sample `.pt` files, across shard boundaries, and must ignore a stale pack."""
    import warnings
    from qflux_b200.cache_loader import CachedEmbeddingLoader, pack_cache
    truth = _write_cache(str(tmp_path), n=7)
    plain = list(CachedEmbeddingLoader(str(tmp_path), batch_size=3, device="cpu", shuffle=True, seed=5, drop_last=False, packed=False))
    out_dir = pack_cache(str(tmp_path), shard_bytes=200_000)  # small shards: several files per key
    assert len([f for f in os.listdir(out_dir) if f.startswith("image_latents.")]) > 1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # torch.from_numpy on a read-only memmap
        ld = CachedEmbeddingLoader(str(tmp_path), batch_size=3, device="cpu", shuffle=True, seed=5, drop_last=False, packed=True)
        packed = list(ld)
    assert ld._pack is not None and len(packed) == len(plain) == 3
    for a, b in zip(plain, packed):
        assert a["img_shapes"] == b["img_shapes"]
        for k in ("image_latents", "control_latents", "prompt_embeds", "prompt_embeds_mask"):
            assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
    # a sample added after packing: "auto" falls back to the per-sample files, packed=True refuses
    g = torch.Generator().manual_seed(9)
    for k, shape in (("image_latents", (1, 64, 64)), ("control_latents", (1, 64, 64)), ("prompt_embeds", (1, 4, 32))):
        torch.save(torch.randn(*shape, generator=g).to(torch.float16), os.path.join(str(tmp_path), k, "new.pt"))
    with open(os.path.join(str(tmp_path), "metadata", "zzz_new.json"), "w") as f:
        json.dump({"version": "1.0", "img_shapes": [[3, 128, 128], [3, 128, 128]], "image_latents": "new", "control_latents": "new",
                   "prompt_embeds": "new"}, f)
    assert CachedEmbeddingLoader(str(tmp_path), 1, device="cpu", packed="auto")._pack is None
    import pytest
    with pytest.raises(ValueError, match="re-run pack_cache"):
        CachedEmbeddingLoader(str(tmp_path), 1, device="cpu", packed=True)
