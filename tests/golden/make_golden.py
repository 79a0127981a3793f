"""Generate the golden fixtures under tests/golden/ by running the REFERENCE ITSELF
(/root/reference, imported through tests/ref_harness.py) on seeded synthetic inputs and the
seeded synthetic checkpoint of rdeic_b200.synthetic (loaded into the reference modules with
their own `load_state_dict`, which also proves the checkpoint-layout contract).

Run in the build container only:   python tests/golden/make_golden.py
The fixtures travel to the GPU box; /root/reference does not.

Files written
  state_dict_keys.json     decode-path keys + shapes of the reference model (full rdeic.yaml)
  full_unet_step.npz       reference apply_model / apply_model_unconditional, full width, 256x256
  full_vae_decode.npz      reference decode_first_stage, full width, 16x16 latent
  full_sampler.npz         reference SpacedSampler.sample (2, 3, 5 steps, CFG) and DDIMSampler.sample (2, 5 steps)
  full_unet_step_64x64.npz, full_unet_step_64x96.npz   one full-width relay step at the latent shapes of BASELINE
                           configs 2/5 and 3
  small_*.npz              the same on the reduced-width config (fast CPU tests of the oracle)
  full_c1_decode.npz       BASELINE config[0] end to end: the reference's q_sample -> SpacedSampler.sample (2 relay
                           steps) -> decode_first_stage on one 256x256 image, full width (latent fp32, image fp16)
  {small,full}_compression.npz       reference model/compression.py Compression.compress -> .decompress
  {small,full}_compression_keys.json its state_dict keys + shapes
  {small,full}_vae_encode.npz   reference AutoencoderKL.encode_hc feature map; vae_encoder_keys.json its keys
  bitstream_ref.bin        bytes written by the reference utils/utils.py write_body (shape (8,12), 2 strings)
  entropy_ref.npz          reference utils/ckbd.py checkerboard ops, utils/func.py scale table and
                           model/compression_modules.py VectorQuantiser.quant / get_codebook_entry
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import ref_harness as rh  # noqa: E402
from rdeic_b200 import synthetic  # noqa: E402

DECODE_PREFIXES = ("model.diffusion_model.", "control_model.", "first_stage_model.decoder.",
                   "first_stage_model.post_quant_conv.")
WEIGHT_SEED = 231


def inputs(B, h, w, hint_c, ctx_dim, n_noise):
    g = lambda s: torch.Generator().manual_seed(s)
    c_latent = torch.randn(B, 4, h, w, generator=g(7))
    hint = torch.randn(B, hint_c, h, w, generator=g(8))
    ctx = torch.randn(B, 77, ctx_dim, generator=g(9))
    gn = g(231)
    noises = [torch.randn(B, 4, h, w, generator=gn) for _ in range(n_noise)]
    return c_latent, hint, ctx, noises


def load_synthetic_into_reference(model, params):
    sd = synthetic.make_state_dict(params, seed=WEIGHT_SEED)
    ref_keys = {k for k in model.state_dict() if k.startswith(DECODE_PREFIXES)}
    missing = ref_keys - set(sd)
    extra = set(sd) - ref_keys
    assert not missing and not extra, (sorted(missing)[:5], sorted(extra)[:5])
    for k, v in model.state_dict().items():
        if k in sd:
            assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
    res = model.load_state_dict(sd, strict=False)
    assert not [k for k in res.unexpected_keys], res.unexpected_keys
    return sd


def run_config(tag, overrides, unet_hw, vae_hw, samp_hw):
    model, mods = rh.build_reference_model(overrides)
    cfg = rh.load_config(overrides)
    params = cfg["params"]
    if tag == "full":
        keys = {k: list(v.shape) for k, v in model.state_dict().items() if k.startswith(DECODE_PREFIXES)}
        (HERE / "state_dict_keys.json").write_text(json.dumps(keys, indent=0, sort_keys=True))
    load_synthetic_into_reference(model, params)
    hint_c = params["control_stage_config"]["params"]["hint_channels"]
    ctx_dim = params["unet_config"]["params"]["context_dim"]

    with torch.no_grad():
        # ---- one relay step ----
        B, (h, w) = 1, unet_hw
        c_latent, hint, ctx, noises = inputs(B, h, w, hint_c, ctx_dim, 1)
        x = model.q_sample(c_latent, torch.full((B,), 299, dtype=torch.long), noises[0])
        cond = {"c_latent": [c_latent], "c_crossattn": [ctx], "guide_hint": hint}
        t = torch.full((B,), 224, dtype=torch.long)
        eps = model.apply_model(x, t, cond)
        eps_u = model.apply_model_unconditional(x, t, cond)
        np.savez_compressed(HERE / f"{tag}_unet_step.npz", x=x.numpy(), t=t.numpy(), eps=eps.numpy(),
                            eps_uncond=eps_u.numpy(), hw=np.array([h, w]))
        print(tag, "unet step", eps.abs().max().item(), eps_u.abs().max().item())

        # ---- VAE decode ----
        h, w = vae_hw
        z = torch.randn(1, 4, h, w, generator=torch.Generator().manual_seed(11))
        img = model.decode_first_stage(z)
        np.savez_compressed(HERE / f"{tag}_vae_decode.npz", z=z.numpy(), img=img.numpy())
        print(tag, "vae", img.abs().max().item())

        # ---- samplers (noise injected instead of torch.randn_like / noise_like) ----
        h, w = samp_hw
        B = 2 if tag == "small" else 1
        c_latent, hint, ctx, noises = inputs(B, h, w, hint_c, ctx_dim, 8)
        cond = {"c_latent": [c_latent], "c_crossattn": [ctx], "guide_hint": hint}
        x_T = model.q_sample(c_latent, torch.full((B,), model.used_timesteps - 1, dtype=torch.long), noises[0])
        out = {"x_T": x_T.numpy()}
        pool = []
        orig_randn_like = torch.randn_like
        torch.randn_like = lambda t_, **k: pool.pop(0)
        try:
            for steps in (2, 3, 5):
                pool[:] = [n.clone() for n in noises[1:1 + steps]]
                s = mods["spaced"].SpacedSampler(model, var_type="fixed_small")
                out[f"spaced_{steps}"] = s.sample(steps, (B, 4, h, w), cond, x_T=x_T.clone()).numpy()
            pool[:] = [n.clone() for n in noises[1:3]]
            s = mods["spaced"].SpacedSampler(model, var_type="fixed_small")
            out["spaced_2_cfg"] = s.sample(2, (B, 4, h, w), cond, x_T=x_T.clone(),
                                           unconditional_guidance_scale=1.5).numpy()
        finally:
            torch.randn_like = orig_randn_like
        rh.patch_ddim_register_buffer(mods["ddim"])
        pool[:] = [n.clone() for n in noises[1:3]]
        mods["ddim"].noise_like = lambda shape, device, repeat=False: pool.pop(0)
        d = mods["ddim"].DDIMSampler(model)
        samples, _ = d.sample(S=2, batch_size=B, shape=(4, h, w), conditioning=cond, x_T=x_T.clone(), eta=0.0,
                              verbose=False)
        out["ddim_2"] = samples.numpy()
        pool[:] = [n.clone() for n in noises[1:6]]
        samples, _ = d.sample(S=5, batch_size=B, shape=(4, h, w), conditioning=cond, x_T=x_T.clone(), eta=0.0,
                              verbose=False)
        out["ddim_5"] = samples.numpy()
        np.savez_compressed(HERE / f"{tag}_sampler.npz", **out)
        print(tag, "samplers", {k: float(np.abs(v).max()) for k, v in out.items()})


def run_big_steps():
    """One full-width relay step at the latent shapes of BASELINE configs 2/5 (64x64) and 3 (64x96) from the
    reference itself (full-width goldens existed only at 32x32 / 16x16)."""
    model, mods = rh.build_reference_model(None)
    params = rh.load_config(None)["params"]
    load_synthetic_into_reference(model, params)
    hint_c = params["control_stage_config"]["params"]["hint_channels"]
    ctx_dim = params["unet_config"]["params"]["context_dim"]
    with torch.no_grad():
        for h, w in ((64, 64), (64, 96)):
            c_latent, hint, ctx, noises = inputs(1, h, w, hint_c, ctx_dim, 1)
            x = model.q_sample(c_latent, torch.full((1,), 299, dtype=torch.long), noises[0])
            cond = {"c_latent": [c_latent], "c_crossattn": [ctx], "guide_hint": hint}
            t = torch.full((1,), 224, dtype=torch.long)
            eps = model.apply_model(x, t, cond)
            np.savez_compressed(HERE / f"full_unet_step_{h}x{w}.npz", x=x.numpy(), t=t.numpy(), eps=eps.numpy(),
                                hw=np.array([h, w]))
            print("full unet step", h, w, float(eps.abs().max()))


def run_c1():
    """BASELINE config[0]: one synthetic 256x256 image, 2 relay steps, full-width model, fp32 on the CPU,
    through the reference's own driver sequence (inference.py:63-87): q_sample at t = used_timesteps - 1,
    SpacedSampler.sample, decode_first_stage.  Noise is injected where the reference draws it."""
    model, mods = rh.build_reference_model(None)
    params = rh.load_config(None)["params"]
    load_synthetic_into_reference(model, params)
    hint_c = params["control_stage_config"]["params"]["hint_channels"]
    ctx_dim = params["unet_config"]["params"]["context_dim"]
    B, h, w, steps = 1, 32, 32, 2
    c_latent, hint, ctx, noises = inputs(B, h, w, hint_c, ctx_dim, 1 + steps)
    cond = {"c_latent": [c_latent], "c_crossattn": [ctx], "guide_hint": hint}
    pool = [n.clone() for n in noises[1:]]
    orig_randn_like = torch.randn_like
    torch.randn_like = lambda t_, **k: pool.pop(0)
    try:
        with torch.no_grad():
            x_T = model.q_sample(c_latent, torch.full((B,), model.used_timesteps - 1, dtype=torch.long), noises[0])
            s = mods["spaced"].SpacedSampler(model, var_type="fixed_small")
            z = s.sample(steps, (B, 4, h, w), cond, unconditional_guidance_scale=1.0, unconditional_conditioning=None,
                         cond_fn=None, x_T=x_T)
            img = model.decode_first_stage(z)
    finally:
        torch.randn_like = orig_randn_like
    np.savez_compressed(HERE / "full_c1_decode.npz", z=z.numpy(), img=img.numpy().astype(np.float16),
                        hw=np.array([h, w]), steps=np.array(steps))
    print("c1 decode", tuple(img.shape), float(img.abs().max()), float(z.abs().max()))


def entropy_inputs():
    g = torch.Generator().manual_seed(41)
    y = torch.randn(2, 8, 6, 10, generator=g) * 6
    y.view(-1)[3] = -0.0
    K, D = 512, 256
    cb = (torch.rand(K, D, generator=g) * 2 - 1) / K
    cb[7] = cb[300]                                      # exact tie: first index wins
    pick = torch.randint(0, K, (2 * 3 * 5,), generator=g)
    pick[0] = 300
    z = cb[pick] + 1e-5 * torch.randn(pick.numel(), D, generator=g)
    z[0] = cb[300]
    z = z.reshape(2, 3, 5, D).permute(0, 3, 1, 2).contiguous()
    return y, cb, z


def run_entropy():
    """a9 / a11: run the reference's own utils/ckbd.py checkerboard ops and
    model/compression_modules.py VectorQuantiser on seeded inputs."""
    import importlib

    rh.import_reference()
    ck = importlib.import_module("utils.ckbd")
    cm = importlib.import_module("model.compression_modules")
    func = importlib.import_module("utils.func")
    y, cb, z = entropy_inputs()
    out = {"anchor": ck.ckbd_anchor(y), "nonanchor": ck.ckbd_nonanchor(y),
           "anchor_sq": ck.ckbd_anchor_sequeeze(y), "nonanchor_sq": ck.ckbd_nonanchor_sequeeze(y)}
    out["anchor_unsq"] = ck.ckbd_anchor_unsequeeze(out["anchor_sq"])
    out["nonanchor_unsq"] = ck.ckbd_nonanchor_unsequeeze(out["nonanchor_sq"])
    a, n = ck.ckbd_split(y)
    out["merge"] = ck.ckbd_merge(a, n)
    out["scale_table"] = func.get_scale_table()
    vq = cm.VectorQuantiser(cb.shape[0], cb.shape[1])
    vq.eval()
    with torch.no_grad():
        vq.embedding.weight.copy_(cb)
        zq, idx = vq.quant(z)
        out["vq_zq"], out["vq_idx"] = zq, idx
        out["vq_entry"] = vq.get_codebook_entry(idx)
    np.savez_compressed(HERE / "entropy_ref.npz", **{k: v.numpy() for k, v in out.items()})
    print("entropy goldens", {k: tuple(v.shape) for k, v in out.items()})


def compression_inputs(pp, B, h, w):
    """Feature map fed to `Compression.compress` (the VAE encoder's 512-channel output at H/8)."""
    g = torch.Generator().manual_seed(53)
    return torch.randn(B, int(pp["in_nc"]), h, w, generator=g)


def run_compression(tag, params, B, h, w):
    """§8(f) rank 1/3: run the reference `Compression.compress` -> `.decompress` (model/compression.py:
    151-273) with its own conv stacks and orchestration.  Absent third-party pieces are replaced as
    SURVEY.md §8c prescribes: compressai's GaussianConditional arithmetic by the oracle restatement,
    the rANS coder by the loopback hand-off, torchac by an identity hand-off of the VQ indices."""
    import importlib

    rh.import_reference()
    comp = importlib.import_module("model.compression")
    from oracle import compression as oc, entropy as oe

    pp = params["preprocess_config"]["params"]
    m = comp.Compression(**pp).eval()
    sd = synthetic.make_compression_state_dict(pp, seed=232)
    res = m.load_state_dict({k[len("preprocess_model."):]: v for k, v in sd.items()}, strict=True)
    keys = {k: list(v.shape) for k, v in m.state_dict().items()}
    (HERE / f"{tag}_compression_keys.json").write_text(json.dumps(keys, indent=0, sort_keys=True))

    table = oe.get_scale_table()

    class GC:            # compressai 1.2.4 entry points used by utils/ckbd.py:81-82,92-93,102,111
        quantized_cdf = torch.zeros(1, 1)
        cdf_length = torch.zeros(1)
        offset = torch.zeros(1)

        def build_indexes(self, scales):
            return torch.from_numpy(oe.build_indexes(scales.numpy(), table))

        def quantize(self, x, mode, means=None):
            assert mode == "symbols"
            return torch.from_numpy(oe.quantize_symbols(x.numpy(), None if means is None else means.numpy()))

    del m.gaussian_conditional          # the shim registered it as a child module
    m.__dict__["gaussian_conditional"] = GC()
    coder = oc.LoopbackCoder()

    class Enc:
        def encode_with_indexes(self, symbols, indexes, *a):
            coder.encode_with_indexes(symbols, indexes)

        def flush(self):
            return b"loopback"

    class Dec:
        def set_stream(self, s):
            coder.pos = 0

        def decode_stream(self, indexes, *a):
            return coder.decode_stream(indexes)

    comp.BufferedRansEncoder, comp.RansDecoder = Enc, Dec
    held = {}
    comp.compress_hyper_latent = lambda idx, K: held.setdefault("z_idx", idx.clone())
    comp.decompress_hyper_latent = lambda s, shape, codebook_size: s

    x = compression_inputs(pp, B, h, w)
    with torch.no_grad():
        y = m.encoder(x)
        z = m.hyper_enc(y)
        out = m.compress(x)
        c_latent, guide_hint = m.decompress(out["strings"], out["shape"])
        hyper = m.hyper_dec(m.quantize.get_codebook_entry(held["z_idx"].long()))
    sym, idx = np.asarray(coder.symbols, dtype=np.int32), np.asarray(coder.indexes, dtype=np.int32)
    print(tag, "compression: y", tuple(y.shape), "std %.2f" % float(y.std()), "symbols nonzero %.2f" % float((sym != 0).mean()),
          "max|sym|", int(np.abs(sym).max()), "index range", int(idx.min()), int(idx.max()),
          "distinct indexes", len(np.unique(idx)))
    np.savez_compressed(HERE / f"{tag}_compression.npz", x=x.numpy(), y=y.numpy(), z=z.numpy(),
                        z_idx=held["z_idx"].numpy().astype(np.int64), hyper_params=hyper.numpy(), symbols=sym,
                        indexes=idx, c_latent=c_latent.numpy(), guide_hint=guide_hint.numpy())


def run_vae_encode(tag, overrides, hw):
    """§8(f) rank 3: the reference AutoencoderKL.encode_hc (autoencoder.py:91-95) feature map `c`."""
    model, mods = rh.build_reference_model(overrides)
    params = rh.load_config(overrides)["params"]
    sd = synthetic.make_state_dict(params, seed=WEIGHT_SEED, encoder=True)
    enc_keys = {k: list(v.shape) for k, v in model.state_dict().items()
                if k.startswith(("first_stage_model.encoder.", "first_stage_model.quant_conv."))}
    assert enc_keys == {k: list(v.shape) for k, v in sd.items() if k in enc_keys} and len(enc_keys) == 108, len(enc_keys)
    if tag == "full":
        (HERE / "vae_encoder_keys.json").write_text(json.dumps(enc_keys, indent=0, sort_keys=True))
    model.load_state_dict(sd, strict=False)
    H, W = hw
    x = torch.rand(1, 3, H, W, generator=torch.Generator().manual_seed(71)) * 2 - 1
    with torch.no_grad():
        _, c = model.encode_first_stage(x)
    print(tag, "vae encode_hc", tuple(c.shape), float(c.abs().max()))
    np.savez_compressed(HERE / f"{tag}_vae_encode.npz", x=x.numpy(), c=c.numpy())


def run_bitstream():
    """§8(f) rank 2: bytes written by the reference's own utils/utils.py write_body."""
    import importlib
    import io

    rh.import_reference()
    u = importlib.import_module("utils.utils")
    g = torch.Generator().manual_seed(5)
    y_string = bytes(torch.randint(0, 256, (1500,), generator=g).tolist())
    z_string = bytes(torch.randint(0, 256, (41,), generator=g).tolist())
    f = io.BytesIO()
    n = u.write_body(f, (8, 12), [[y_string], [z_string]])
    data = f.getvalue()
    assert n == len(data)
    f.seek(0)
    strings, shape = u.read_body(f)
    assert strings == [[y_string], [z_string]] and tuple(shape) == (8, 12)
    (HERE / "bitstream_ref.bin").write_bytes(data)
    print("bitstream golden", len(data), "bytes")


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["entropy", "bitstream", "vae_encode", "compression", "small", "full", "c1", "big"]
    if "big" in which:
        run_big_steps()
    if "entropy" in which:
        run_entropy()
    if "bitstream" in which:
        run_bitstream()
    if "vae_encode" in which:
        run_vae_encode("small", rh.SMALL_OVERRIDES, (64, 96))
        run_vae_encode("full", None, (64, 64))
    if "compression" in which:
        from rdeic_b200 import configs

        run_compression("small", configs.small_params(), 2, 16, 24)
        run_compression("full", configs.default_params(), 1, 16, 16)
    if "small" in which:
        run_config("small", rh.SMALL_OVERRIDES, (16, 16), (8, 8), (8, 16))
    if "full" in which:
        run_config("full", None, (32, 32), (16, 16), (16, 16))
    if "c1" in which:
        run_c1()
