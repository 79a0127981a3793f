"""CPU tests of the data-parallel plumbing: world_size-2 gloo process groups
exercise the weight broadcast, the batch sharding + image gather and the tile sharding + blend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rdeic_b200 import parallel


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            chunks = [parallel.shard_range(n, r, world) for r in range(world)]
            assert chunks[0][0] == 0 and chunks[-1][1] == n
            for (a, b), (c, d) in zip(chunks, chunks[1:]):
                assert b == c
            sizes = [b - a for a, b in chunks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def test_plan_tiles_cover_and_overlap():
    for h, w in ((256, 176), (64, 64), (96, 200), (97, 131)):
        plan = parallel.plan_tiles(h, w, tile=96, overlap=16)
        cover = np.zeros((h, w), int)
        for y0, x0, th, tw in plan:
            assert 0 <= y0 and y0 + th <= h and 0 <= x0 and x0 + tw <= w
            cover[y0:y0 + th, x0:x0 + tw] += 1
        assert cover.min() >= 1
        ys = sorted({p[0] for p in plan})
        for a, b in zip(ys, ys[1:]):
            assert a + min(96, h) - b >= 16          # neighbouring rows of tiles overlap by >= 16


def test_plan_tiles_balanced():
    """Tile counts are multiples of the world size, tiles are equal, aligned, cover the latent and overlap."""
    for h, w in ((176, 256), (64, 64), (96, 200), (128, 128)):
        for world in (1, 2, 4, 8):
            plan = parallel.plan_tiles_balanced(h, w, world, overlap=16)
            assert len(plan) % world == 0
            th, tw = plan[0][2], plan[0][3]
            assert all(p[2] == th and p[3] == tw for p in plan) and th % 8 == 0 and tw % 8 == 0 and th * tw <= 96 * 96
            cover = np.zeros((h, w), int)
            for y0, x0, _, _ in plan:
                assert 0 <= y0 and y0 + th <= h and 0 <= x0 and x0 + tw <= w
                cover[y0:y0 + th, x0:x0 + tw] += 1
            assert cover.min() >= 1
            ys, xs = sorted({p[0] for p in plan}), sorted({p[1] for p in plan})
            assert all(a + th - b >= 16 for a, b in zip(ys, ys[1:])) and all(a + tw - b >= 16 for a, b in zip(xs, xs[1:]))
    # BASELINE config 4 on 8 GPUs: two tiles per rank instead of the 3/3/3/3/2/2/2/2 deal of 20 tiles
    plan = parallel.plan_tiles_balanced(176, 256, 8, overlap=16)
    assert len(plan) in (8, 16) and (plan[0][2] * plan[0][3]) % 128 == 0
    assert len(parallel.plan_tiles_balanced(176, 256, 8, overlap=16, max_tile_area=80 * 80)) == 16


def _torch_blend_u8(tiles, origins, ov, H, W):
    """CPU stand-in for ops.blend_tiles_u8 (same ramp, same rounding) for the gloo plumbing test."""
    T, th, tw, _ = tiles.shape
    ramp = lambda n: torch.tensor([(i + 1) / (ov + 1) if i < ov else ((n - i) / (ov + 1) if i >= n - ov else 1.0)
                                   for i in range(n)])
    wgt = ramp(th)[:, None] * ramp(tw)[None, :]
    acc, ws = torch.zeros(H, W, 3), torch.zeros(H, W, 1)
    for t in range(T):
        y0, x0 = int(origins[t, 0]), int(origins[t, 1])
        acc[y0:y0 + th, x0:x0 + tw] += tiles[t].float() * wgt[..., None]
        ws[y0:y0 + th, x0:x0 + tw] += wgt[..., None]
    return torch.clamp(torch.round(acc / ws), 0, 255).to(torch.uint8)


def test_blend_of_consistent_tiles_is_identity():
    h, w, s = 40, 56, 2
    full = torch.randn(3, h * s, w * s)
    plan = parallel.plan_tiles(h, w, tile=24, overlap=8)
    tiles = [full[:, y0 * s:(y0 + th) * s, x0 * s:(x0 + tw) * s] for y0, x0, th, tw in plan]
    out = parallel.blend_tiles(tiles, plan, h, w, overlap=8, scale=s)
    assert torch.allclose(out, full, atol=1e-5)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. weight broadcast: only rank 0 holds the tensors
        spec = [("a.weight", (5, 3, 3, 3)), ("a.bias", (5,)), ("b.weight", (7, 5)), ("scale_list", (4,))]
        g = torch.Generator().manual_seed(0)
        full = {k: torch.randn(s, generator=g) for k, s in spec}
        got = parallel.broadcast_state_dict(full if rank == 0 else None, spec, "cpu", src=0, bucket_bytes=300)
        ok_bcast = all(torch.equal(got[k], full[k]) for k, _ in spec)
        # 2. batch sharding + gather of "decoded images" (uneven: 5 images over 2 ranks)
        n = 5
        imgs = torch.arange(n * 4 * 6 * 3, dtype=torch.uint8).reshape(n, 4, 6, 3)
        cond = {"c_latent": [torch.arange(n).float().view(n, 1, 1, 1)], "c_crossattn": [torch.zeros(n, 77, 8)],
                "guide_hint": torch.arange(n).float().view(n, 1, 1, 1)}
        mine = parallel.shard_cond(cond, rank, world)
        lo, hi = parallel.shard_range(n, rank, world)
        ok_shard = mine["guide_hint"].flatten().tolist() == list(range(lo, hi)) and \
            mine["c_crossattn"][0].shape[0] == hi - lo
        counts = [parallel.shard_range(n, r, world)[1] - parallel.shard_range(n, r, world)[0] for r in range(world)]
        out = parallel.gather_images(imgs[lo:hi].clone(), counts, dst=0)
        ok_gather = (out is None) if rank != 0 else torch.equal(out, imgs)
        # 3. tiles of one large latent dealt across ranks, decoded by a stand-in, blended on rank 0
        h, w, s = 40, 56, 2
        lat = torch.randn(1, 4, h, w, generator=torch.Generator().manual_seed(1))
        big = {"c_latent": [lat], "c_crossattn": [torch.zeros(1, 77, 8)], "guide_hint": torch.zeros(1, 2, h, w)}
        seen = []

        def fake_decode(c, i):   # "decode" = nearest upsample of the first 3 latent channels
            seen.append(i)
            return torch.nn.functional.interpolate(c["c_latent"][0][:, :3], scale_factor=s, mode="nearest")

        img = parallel.decode_tiled(fake_decode, big, tile=24, overlap=8, scale=s)
        ref = torch.nn.functional.interpolate(lat[:, :3], scale_factor=s, mode="nearest")[0]
        ok_tiles = (img is None) if rank != 0 else torch.allclose(img, ref, atol=1e-5)
        ok_deal = all(i % world == rank for i in seen)
        # same through the batched form: a rank's tiles stacked along the batch axis, one call

        def fake_decode_batched(c, idx):
            assert c["c_latent"][0].shape[0] == len(idx) == c["c_crossattn"][0].shape[0] == c["guide_hint"].shape[0]
            return torch.nn.functional.interpolate(c["c_latent"][0][:, :3], scale_factor=s, mode="nearest")

        img_b = parallel.decode_tiled(fake_decode_batched, big, tile=24, overlap=8, scale=s, batched=True)
        ok_tiles = ok_tiles and ((img_b is None) if rank != 0 else torch.allclose(img_b, ref, atol=1e-5))
        # the serving form: balanced plan, uint8 tiles, gather to rank 0 only, one blend call
        plan = parallel.plan_tiles_balanced(h, w, world, overlap=8, max_tile_area=24 * 32)
        full_u8 = (torch.rand(h * s, w * s, 3, generator=torch.Generator().manual_seed(2)) * 255).to(torch.uint8)

        def fake_decode_u8(c, idx):           # every tile is the matching crop of one consistent image
            return torch.stack([full_u8[plan[i][0] * s:(plan[i][0] + plan[i][2]) * s,
                                        plan[i][1] * s:(plan[i][1] + plan[i][3]) * s] for i in idx])

        img_u8 = parallel.decode_tiled_u8(fake_decode_u8, big, plan, overlap=8, scale=s, blend=_torch_blend_u8)
        ok_tiles = ok_tiles and ((img_u8 is None) if rank != 0 else torch.equal(img_u8, full_u8))
        ret[rank] = (ok_bcast, ok_shard, ok_gather, ok_tiles, ok_deal)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_broadcast_shard_gather_tiles():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r] == (True, True, True, True, True), (r, ret[r])


def test_pad_and_size_grouping():
    """Caller glue of SURVEY §8(f) rank 4: utils/image/common.py:251 pad, inference_partition.py:438-452 grouping."""
    import numpy as np

    from rdeic_b200.utils import group_by_padded_size, pad

    img = np.arange(5 * 70 * 3, dtype=np.uint8).reshape(5, 70, 3)
    p = pad(img, 64)
    assert p.shape == (64, 128, 3) and np.array_equal(p[:5, :70], img) and p[5:].sum() == 0 and p[:, 70:].sum() == 0
    assert pad(np.zeros((64, 128, 3), np.uint8), 64).shape == (64, 128, 3)
    sizes = [(500, 700), (512, 768), (100, 100), (510, 705), (128, 128), (65, 1)]
    groups = group_by_padded_size(sizes, batch_size=2)
    assert groups == [((128, 64), [5]), ((128, 128), [2, 4]), ((512, 704), [0]), ((512, 768), [1, 3])]
    assert group_by_padded_size([(512, 768)] * 3, batch_size=2) == [((512, 768), [0, 1]), ((512, 768), [2])]
    assert group_by_padded_size([], 4) == []
