"""Data-parallel plumbing for the relay decode: one process per GPU, weights
replicated, independent images (or independent latent tiles of one large image) dealt across
ranks.  There is no collective inside the denoising loop; torch.distributed (NCCL over NVLink on
the B200 box, gloo in CPU tests) is used for exactly two things: the one-time broadcast of the
checkpoint from rank 0, and the gather of the decoded uint8 images.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous chunk [lo, hi) of `n_items` owned by `rank`; the first n % world chunks get the
    extra item, so chunk sizes differ by at most one and concatenation in rank order restores the
    original order."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_cond(cond: Dict, rank: int, world_size: int) -> Dict:
    """Slice every batch-first tensor of the reference's cond dict (inference.py:57-61)."""
    n = cond["guide_hint"].shape[0]
    lo, hi = shard_range(n, rank, world_size)
    return {"c_latent": [t[lo:hi] for t in cond["c_latent"]],
            "c_crossattn": [t[lo:hi] if t.shape[0] == n else t for t in cond["c_crossattn"]],
            "guide_hint": cond["guide_hint"][lo:hi]}


def broadcast_state_dict(sd: Optional[Dict[str, torch.Tensor]], spec: Sequence[Tuple[str, Tuple[int, ...]]],
                         device, src: int = 0, bucket_bytes: int = 256 << 20) -> Dict[str, torch.Tensor]:
    """One-time weight broadcast.  Every rank knows (key, shape) from the config; rank `src` holds
    the tensors.  Tensors are packed into ~256 MB flat fp32 buckets (launch latency, not link count,
    is what matters on NVSwitch) and broadcast bucket by bucket."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert sd is not None
        return {k: sd[k].to(device) for k, _ in spec}
    rank = dist.get_rank()
    out: Dict[str, torch.Tensor] = {}
    bucket: List[Tuple[str, Tuple[int, ...]]] = []
    size = 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        numel = sum(int(torch.Size(s).numel()) for _, s in bucket)
        flat = torch.empty(numel, dtype=torch.float32, device=device)
        if rank == src:
            off = 0
            for k, s in bucket:
                n = int(torch.Size(s).numel())
                flat[off:off + n].copy_(sd[k].reshape(-1).to(device, torch.float32))
                off += n
        dist.broadcast(flat, src=src)
        off = 0
        for k, s in bucket:
            n = int(torch.Size(s).numel())
            out[k] = flat[off:off + n].view(*s).clone()
            off += n
        bucket, size = [], 0

    for k, s in spec:
        nbytes = int(torch.Size(s).numel()) * 4
        if size and size + nbytes > bucket_bytes:
            flush()
        bucket.append((k, tuple(s)))
        size += nbytes
    flush()
    return out


def gather_images(local: torch.Tensor, counts: Sequence[int], dst: int = 0) -> Optional[torch.Tensor]:
    """Gather per-rank uint8 image batches [b_r,H,W,3] to rank `dst` in rank order: a real gather (only `dst`
    receives; every other rank just sends its chunk).  Chunks may differ by one image (shard_range), so every
    rank pads to the largest chunk first."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    m = max(counts)
    if local.shape[0] == m:
        pad = local.contiguous()
    else:
        pad = torch.zeros((m, *local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    if all(c == m for c in counts):
        return torch.cat(bufs, 0)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


# ---- overlapping latent tiles of one large image (BASELINE config 4) --------------------------
def plan_tiles(h: int, w: int, tile: int = 96, overlap: int = 16) -> List[Tuple[int, int, int, int]]:
    """Cover an h x w latent with tile x tile windows (smaller at the far edge is avoided by
    shifting the last window back) overlapping by >= `overlap`; returns (y0, x0, th, tw)."""
    def starts(n):
        if n <= tile:
            return [0]
        step = tile - overlap
        s = list(range(0, n - tile, step)) + [n - tile]
        return sorted(set(s))
    th, tw = min(tile, h), min(tile, w)
    return [(y, x, th, tw) for y in starts(h) for x in starts(w)]


def plan_tiles_balanced(h: int, w: int, world_size: int, overlap: int = 16, max_tile_area: int = 96 * 96,
                        align: int = 8) -> List[Tuple[int, int, int, int]]:
    """Cover an h x w latent with ny x nx equal tiles whose COUNT is a multiple of `world_size`, so every rank
    decodes the same number of tiles (20 tiles on 8 GPUs keep four ranks idle for a third of the time).  Among
    the grids with ny * nx % world_size == 0, tile area <= `max_tile_area` and a token count that is a multiple of
    128 (the tcgen05 attention tile) pick the one with the least work per rank — convolutions / linears are linear
    in the tile area, self-attention (about a fifth of a 64x64 tile's FLOPs) quadratic; tile sides are multiples of
    `align` (three stride-2 levels), neighbours overlap by >= `overlap`.  Returns (y0, x0, th, tw) like plan_tiles."""
    def side(n: int, k: int) -> int:              # smallest aligned tile side so that k tiles cover n with the overlap
        t = -(-(n + (k - 1) * overlap) // k)
        t = -(-t // align) * align
        return min(t, n)

    def starts(n: int, k: int, t: int) -> List[int]:
        if k == 1 or t >= n:
            return [0]
        return [round(i * (n - t) / (k - 1)) for i in range(k)]

    best = None
    for ny in range(1, h // align + 1):
        for nx in range(1, w // align + 1):
            if (ny * nx) % world_size:
                continue
            th, tw = side(h, ny), side(w, nx)
            if th * tw > max_tile_area or (ny > 1 and 2 * overlap > th) or (nx > 1 and 2 * overlap > tw):
                continue
            if ny > 1 and (h - th) / (ny - 1) > th - overlap or nx > 1 and (w - tw) / (nx - 1) > tw - overlap:
                continue
            area = th * tw
            if area % 128 and (ny * nx > 1):
                continue
            cost = (ny * nx // world_size) * area * (1.0 + 0.2 * area / 4096.0)
            if area % 2048:                           # levels 1 / 2 leave the tcgen05 attention tile too
                cost *= 1.05
            if best is None or cost < best[0]:
                best = (cost, ny, nx, th, tw)
    if best is None:
        raise ValueError(f"plan_tiles_balanced: no {world_size}-way balanced tiling of {h}x{w} with tiles <= {max_tile_area}")
    _, ny, nx, th, tw = best
    return [(y, x, th, tw) for y in starts(h, ny, th) for x in starts(w, nx, tw)]


def tile_weights(th: int, tw: int, overlap: int, scale: int, device) -> torch.Tensor:
    """Separable linear ramp over the overlap band (in pixels = latent * scale) used to blend tiles."""
    def ramp(n, ov):
        r = torch.ones(n)
        if ov > 0:
            e = torch.linspace(1.0 / (ov + 1), ov / (ov + 1.0), ov)
            r[:ov] = e
            r[-ov:] = e.flip(0)
        return r
    wy, wx = ramp(th * scale, overlap * scale), ramp(tw * scale, overlap * scale)
    return (wy[:, None] * wx[None, :]).to(device)


def blend_tiles(tiles: Sequence[torch.Tensor], plan: Sequence[Tuple[int, int, int, int]], h: int, w: int,
                overlap: int, scale: int = 8) -> torch.Tensor:
    """Weighted average of decoded float tiles [3, th*scale, tw*scale] into one [3, h*scale, w*scale]."""
    dev = tiles[0].device
    acc = torch.zeros((3, h * scale, w * scale), dtype=torch.float32, device=dev)
    wsum = torch.zeros((1, h * scale, w * scale), dtype=torch.float32, device=dev)
    for t, (y0, x0, th, tw) in zip(tiles, plan):
        wgt = tile_weights(th, tw, overlap, scale, dev)
        ys, xs = y0 * scale, x0 * scale
        acc[:, ys:ys + th * scale, xs:xs + tw * scale] += t.float() * wgt
        wsum[:, ys:ys + th * scale, xs:xs + tw * scale] += wgt
    return acc / wsum


def crop_cond(cond: Dict, y0: int, x0: int, th: int, tw: int) -> Dict:
    """Latent-space crop of the reference cond dict (c_latent / guide_hint are at H/8)."""
    sl = (slice(None), slice(None), slice(y0, y0 + th), slice(x0, x0 + tw))
    return {"c_latent": [t[sl].contiguous() for t in cond["c_latent"]], "c_crossattn": cond["c_crossattn"],
            "guide_hint": cond["guide_hint"][sl].contiguous()}


def decode_tiled(decode_fn, cond: Dict, tile: int = 96, overlap: int = 16, scale: int = 8,
                 rank: Optional[int] = None, world_size: Optional[int] = None, dst: int = 0, batched: bool = False):
    """Decode ONE large image (batch 1) as overlapping latent tiles dealt round-robin across ranks
    (BASELINE config 4: 2048x1365 -> latent 256x176).  `decode_fn(cond_tile, tile_index) ->
    float tensor [1,3,th*scale,tw*scale]` is the per-tile relay decode (each tile is an independent
    decode unit: GroupNorm / attention statistics are per tile, so the result is NOT the full-frame
    decode — the reference never tiles).  With `batched`, a rank's tiles (all the same
    size) are stacked along the batch axis and decoded by ONE call `decode_fn(cond_tiles, indices) ->
    [n,3,th*scale,tw*scale]`.  Returns the blended image on `dst`."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    lat = cond["c_latent"][0]
    assert lat.shape[0] == 1, "decode_tiled handles one image at a time"
    h, w = lat.shape[-2:]
    plan = plan_tiles(h, w, tile, overlap)
    mine = list(range(rank, len(plan), world_size))
    if batched and mine:
        crops = [crop_cond(cond, *plan[i]) for i in mine]
        ctx = cond["c_crossattn"][0]
        stacked = {"c_latent": [torch.cat([c["c_latent"][0] for c in crops], 0)],
                   "c_crossattn": [ctx.expand(len(mine), -1, -1).contiguous()],
                   "guide_hint": torch.cat([c["guide_hint"] for c in crops], 0)}
        outs = list(decode_fn(stacked, mine).float().unbind(0))
    else:
        outs = [decode_fn(crop_cond(cond, *plan[i]), i)[0].float() for i in mine]
    th, tw = plan[0][2], plan[0][3]
    dev = lat.device
    if world_size == 1:
        return blend_tiles(outs, plan, h, w, overlap, scale)
    per = (len(plan) + world_size - 1) // world_size
    buf = torch.zeros((per, 3, th * scale, tw * scale), dtype=torch.float32, device=dev)
    for j, o in enumerate(outs):
        buf[j] = o
    gathered = [torch.empty_like(buf) for _ in range(world_size)]
    dist.all_gather(gathered, buf)
    if rank != dst:
        return None
    tiles = [None] * len(plan)
    for r in range(world_size):
        for j, i in enumerate(range(r, len(plan), world_size)):
            tiles[i] = gathered[r][j]
    return blend_tiles(tiles, plan, h, w, overlap, scale)


def decode_tiled_u8(decode_fn, cond: Dict, plan: Sequence[Tuple[int, int, int, int]], overlap: int = 16, scale: int = 8,
                    rank: Optional[int] = None, world_size: Optional[int] = None, dst: int = 0, blend=None):
    """The serving form of `decode_tiled`: tiles leave the VAE as uint8 HWC (the fused tail kernel), a rank's tiles
    are one batch, the gather moves uint8 (a quarter of the fp32 bytes) to `dst` only, and the blend is ONE kernel
    (`ops.blend_tiles_u8`: weighted mean of the covering tiles with the linear ramp of `tile_weights`, rounded to
    uint8).  `plan` comes from plan_tiles / plan_tiles_balanced (equal tile shapes); tiles are dealt round-robin.
    `decode_fn(cond_tiles, indices) -> uint8 [n, th*scale, tw*scale, 3]`.  Returns uint8 [H*scale, W*scale, 3] on
    `dst`, None elsewhere."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    lat = cond["c_latent"][0]
    assert lat.shape[0] == 1, "decode_tiled_u8 handles one image at a time"
    h, w = lat.shape[-2:]
    th, tw = plan[0][2], plan[0][3]
    assert all(p[2] == th and p[3] == tw for p in plan), "decode_tiled_u8 needs equal tile shapes"
    mine = list(range(rank, len(plan), world_size))
    per = -(-len(plan) // world_size)
    dev = lat.device
    buf = torch.zeros((per, th * scale, tw * scale, 3), dtype=torch.uint8, device=dev)
    if mine:
        crops = [crop_cond(cond, *plan[i]) for i in mine]
        ctx = cond["c_crossattn"][0]
        stacked = {"c_latent": [torch.cat([c["c_latent"][0] for c in crops], 0)],
                   "c_crossattn": [ctx.expand(len(mine), -1, -1).contiguous()],
                   "guide_hint": torch.cat([c["guide_hint"] for c in crops], 0)}
        out = decode_fn(stacked, mine)
        assert out.dtype == torch.uint8 and tuple(out.shape) == (len(mine), th * scale, tw * scale, 3)
        buf[:len(mine)] = out
    if world_size > 1:
        bufs = [torch.empty_like(buf) for _ in range(world_size)] if rank == dst else None
        dist.gather(buf, bufs, dst=dst)
        if rank != dst:
            return None
        order = [i for r in range(world_size) for i in range(r, len(plan), world_size)]
        tiles = torch.cat([b[:len(range(r, len(plan), world_size))] for r, b in enumerate(bufs)], 0)
    else:
        order, tiles = mine, buf
    origins = torch.tensor([[plan[i][0] * scale, plan[i][1] * scale] for i in order], dtype=torch.int32, device=dev)
    if blend is None:
        from . import ops
        blend = ops.blend_tiles_u8
    return blend(tiles, origins, overlap * scale, h * scale, w * scale)
