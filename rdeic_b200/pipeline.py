"""The decode half of the reference driver `process()` (inference.py:57-87,
inference_partition.py:280-310): from decompressed conditioning to uint8 images.

    cond = {"c_latent": [c_latent], "c_crossattn": [ctx], "guide_hint": guide_hint}
    imgs = relay_decode(model, cond, steps=5)          # uint8 [B,H,W,3] on the GPU
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence

import torch

from .ddim_sampler_relay import DDIMSampler
from .spaced_sampler_relay import SpacedSampler


@torch.no_grad()
def relay_decode(model, cond: Dict, steps: int, sampler: str = "ddpm", guidance_scale: float = 1.0,
                 start_noise: Optional[torch.Tensor] = None, step_noises: Optional[Sequence[torch.Tensor]] = None,
                 as_uint8: bool = True) -> torch.Tensor:
    """inference.py:63-87.  `start_noise` / `step_noises` replace the reference's device-side
    torch.randn draws (inference.py:65, spaced_sampler_relay.py:378) when reproducibility across
    devices is needed.  By default noise is drawn on the GPU in the reference's order: process() first draws an
    `x_T = randn(shape)` it never uses (inference.py:64) and only then the start noise (:65), so one draw of the
    same shape is made and discarded here to keep the generator stream aligned with the reference's."""
    c_latent = cond["c_latent"][0].to(model.device, torch.float32)
    n, _, h, w = c_latent.shape
    shape = (n, 4, h, w)
    if start_noise is None:
        torch.randn(shape, device=model.device, dtype=torch.float32)          # the reference's dead x_T draw
        noise = torch.randn(shape, device=model.device, dtype=torch.float32)
    else:
        noise = start_noise
    # inference.py:66-67: t = used_timesteps - 1 for every sample (host list: no device round trip)
    x_T = model.q_sample(x_start=c_latent, t=[model.used_timesteps - 1] * n, noise=noise)
    noise_fn: Optional[Callable] = None
    if step_noises is not None:
        noise_fn = lambda i, like: step_noises[i]
    if sampler == "ddpm":
        s = SpacedSampler(model, var_type="fixed_small")
        s.noise_fn = noise_fn
        samples = s.sample(steps, shape, cond, unconditional_guidance_scale=guidance_scale,
                           unconditional_conditioning=None, cond_fn=None, x_T=x_T)
    else:
        s = DDIMSampler(model)
        s.noise_fn = noise_fn
        samples, _ = s.sample(S=steps, batch_size=n, shape=shape[1:], conditioning=cond,
                              unconditional_conditioning=None, unconditional_guidance_scale=guidance_scale,
                              x_T=x_T, eta=0, verbose=False)
    if as_uint8:
        return model.decode_first_stage_u8(samples)
    return model.decode_first_stage(samples)


@torch.no_grad()
def decode_streams(model, stream_paths: Sequence[str], c_crossattn: List[torch.Tensor], steps: int,
                   sizes: Optional[Sequence] = None, batch_size: int = 8, sampler: str = "ddpm",
                   guidance_scale: float = 1.0) -> List[torch.Tensor]:
    """The receiver side of inference.py:57-87 / inference_partition.py:438-520 for a folder of
    bitstreams: decompress every stream (learned compressor on the GPU, byte coders on the host),
    bucket the conditionings by latent size, relay-decode each bucket in batches and crop the
    padding (inference.py:156) when the original (h, w) `sizes` are given.
    `c_crossattn[0]` is the [1,77,1024] embedding of the empty prompt (inference.py:134), repeated per
    batch.  Returns uint8 HWC tensors on the GPU, in input order."""
    from .utils import group_by_padded_size

    conds = [model.apply_condition_decompress(p) for p in stream_paths]
    latent_sizes = [(c.shape[-2] * 8, c.shape[-1] * 8) for c, _ in conds]
    out: List[Optional[torch.Tensor]] = [None] * len(conds)
    ctx = c_crossattn[0]
    for _, members in group_by_padded_size(latent_sizes, batch_size, scale=8):
        c_latent = torch.cat([conds[i][0] for i in members], 0)
        hint = torch.cat([conds[i][1] for i in members], 0)
        n = c_latent.shape[0]
        cond = {"c_latent": [c_latent], "c_crossattn": [ctx.expand(n, -1, -1).contiguous() if ctx.shape[0] == 1 else ctx[:n]],
                "guide_hint": hint}
        imgs = relay_decode(model, cond, steps, sampler=sampler, guidance_scale=guidance_scale)
        k = 0
        for i in members:
            b = conds[i][0].shape[0]
            img = imgs[k:k + b]
            k += b
            if sizes is not None:
                h, w = sizes[i]
                img = img[:, :h, :w]
            out[i] = img[0] if b == 1 else img
    return out


@torch.no_grad()
def decode_host_batches(model, batches: Iterable[Dict], steps: int, sampler: str = "ddpm",
                        guidance_scale: float = 1.0, ring: int = 3) -> Iterator[torch.Tensor]:
    """The serving loop around `relay_decode` for conditioning that lives on the HOST (what
    `apply_condition_decompress` leaves behind a byte coder, or a loader thread): a generator over
    `batches` of host tensors that yields, per batch, the uint8 [B,H,W,3] images in pinned HOST memory.

    Each batch is a dict with "c_latent" [B,4,h,w], "guide_hint" [B,Ch,h,w], "c_crossattn" [B,77,D] and,
    optionally, "start_noise" and "step_noises" (list of [B,4,h,w]); pinned tensors copy asynchronously.

    Double buffered on two copy streams: the host->device copy of batch i+1 and the device->host read of
    batch i-1 run while batch i decodes, so per batch the GPU sees only the decode.  Every batch's inputs
    are still copied host->device and its images device->host; a yielded tensor is complete (its copy
    event has been synchronised) and stays valid until `ring - 1` further batches have been yielded."""
    dev = model.device
    cur = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    slots: List[Optional[Dict]] = [None, None]                  # device-side input buffers
    slot_free = [torch.cuda.Event(), torch.cuda.Event()]        # recorded when the decode reading a slot is queued
    outs: List[Optional[torch.Tensor]] = [None] * max(ring, 2)  # pinned host images

    def flat(hb: Dict) -> Dict[str, torch.Tensor]:
        t = {"c_latent": hb["c_latent"], "guide_hint": hb["guide_hint"], "c_crossattn": hb["c_crossattn"]}
        if hb.get("start_noise") is not None:
            t["start_noise"] = hb["start_noise"]
        for j, n in enumerate(hb.get("step_noises") or ()):
            t[f"step_noise{j}"] = n
        return t

    def stage(i: int, hb: Dict) -> torch.cuda.Event:
        src, s = flat(hb), i & 1
        buf = slots[s]
        if buf is None or buf.keys() != src.keys() or any(buf[k].shape != v.shape for k, v in src.items()):
            buf = slots[s] = {k: torch.empty(v.shape, dtype=torch.float32, device=dev) for k, v in src.items()}
            # fresh buffers come from the current stream's pool: whatever the caller still has queued there (and the
            # previous owner of that memory) must be ordered before the copy stream writes them
            s_in.wait_stream(cur)
        with torch.cuda.stream(s_in):
            s_in.wait_event(slot_free[s])                       # the decode two batches back has consumed the slot
            for k, v in src.items():
                buf[k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_in)
        return ev

    it = iter(batches)
    nxt = next(it, None)
    ev_nxt = stage(0, nxt) if nxt is not None else None
    pending: Optional[tuple] = None
    i = 0
    while nxt is not None:
        ev_i, buf = ev_nxt, slots[i & 1]
        nxt = next(it, None)
        if nxt is not None:
            ev_nxt = stage(i + 1, nxt)                          # prefetch: queued before this batch's kernels
        cur.wait_event(ev_i)
        cond = {"c_latent": [buf["c_latent"]], "c_crossattn": [buf["c_crossattn"]], "guide_hint": buf["guide_hint"]}
        noises = [buf[k] for k in sorted((k for k in buf if k.startswith("step_noise")), key=lambda k: int(k[10:]))]
        img = relay_decode(model, cond, steps, sampler=sampler, guidance_scale=guidance_scale,
                           start_noise=buf.get("start_noise"), step_noises=noises or None)
        slot_free[i & 1].record(cur)
        done = torch.cuda.Event()
        done.record(cur)
        o = i % len(outs)
        if outs[o] is None or outs[o].shape != img.shape:
            outs[o] = torch.empty(img.shape, dtype=torch.uint8, pin_memory=True)
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            outs[o].copy_(img, non_blocking=True)
            ev_out = torch.cuda.Event()
            ev_out.record(s_out)
        img.record_stream(s_out)
        if pending is not None:                                 # the host runs at most one batch ahead of the GPU
            pending[1].synchronize()
            yield pending[0]
        pending = (outs[o], ev_out)
        i += 1
    if pending is not None:
        pending[1].synchronize()
        yield pending[0]
