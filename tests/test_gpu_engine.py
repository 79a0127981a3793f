"""GPU parity of the assembled decode path (UNet+control step, VAE decoder, relay samplers)
through the drop-in API, against golden vectors produced by the reference itself
(tests/golden/, full-width and reduced-width) and against the CPU oracle.

Tolerances (BASELINE.json north_star): per-step UNet output rel-L2 <= 1e-2 in bf16; decoded
image PSNR >= 40 dB against the reference's fp32 output (peak-to-peak 2.0 for [-1,1] images)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from helpers import inputs, psnr, rel_l2
from rdeic_b200 import configs, synthetic

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
UNET_TOL = 1e-2
PSNR_MIN = 40.0


def _model(params, cuda, graph=True):
    from rdeic_b200 import RDEIC

    sd = synthetic.make_state_dict(params, seed=231)
    m = RDEIC.from_config({"params": params}, device=cuda, use_cuda_graph=graph)
    m.load_state_dict(sd)
    return m


@pytest.fixture(scope="module")
def small_model(cuda):
    return _model(configs.small_params(), cuda)


@pytest.fixture(scope="module")
def full_model(cuda):
    return _model(configs.default_params(), cuda)


def _check_unet(model, tag, hint_c, ctx_dim, cuda):
    gold = np.load(GOLD / f"{tag}_unet_step.npz")
    h, w = gold["hw"]
    c_latent, hint, ctx, _ = inputs(1, h, w, hint_c, ctx_dim, 1)
    cond = {"c_latent": [c_latent.to(cuda)], "c_crossattn": [ctx.to(cuda)], "guide_hint": hint.to(cuda)}
    x, t = torch.from_numpy(gold["x"]).to(cuda), torch.from_numpy(gold["t"]).to(cuda)
    eps = model.apply_model(x, t, cond).cpu().numpy()
    eps_u = model.apply_model_unconditional(x, t, cond).cpu().numpy()
    e1, e2 = rel_l2(eps, gold["eps"]), rel_l2(eps_u, gold["eps_uncond"])
    print(f"[{tag}] unet step rel-L2 cond {e1:.3e} uncond {e2:.3e}")
    assert e1 <= UNET_TOL and e2 <= UNET_TOL
    # second call replays the captured CUDA graph: identical result
    eps2 = model.apply_model(x, t, cond).cpu().numpy()
    assert np.array_equal(eps, eps2)


def _check_vae(model, tag, cuda):
    gold = np.load(GOLD / f"{tag}_vae_decode.npz")
    img = model.decode_first_stage(torch.from_numpy(gold["z"]).to(cuda)).cpu().numpy()
    p = psnr(img, gold["img"], 2.0)
    print(f"[{tag}] vae decode PSNR {p:.1f} dB rel-L2 {rel_l2(img, gold['img']):.3e}")
    assert p >= PSNR_MIN
    u8 = model.decode_first_stage_u8(torch.from_numpy(gold["z"]).to(cuda)).cpu().numpy()
    ref_u8 = (((gold["img"] + 1) / 2).clip(0, 1).transpose(0, 2, 3, 1) * 255).clip(0, 255).astype(np.uint8)
    assert psnr(u8, ref_u8, 255.0) >= PSNR_MIN - 1.0     # +quantisation noise of the uint8 grid


def _check_samplers(model, tag, B, h, w, hint_c, ctx_dim, cuda):
    from rdeic_b200 import SpacedSampler, DDIMSampler

    gold = np.load(GOLD / f"{tag}_sampler.npz")
    c_latent, hint, ctx, noises = inputs(B, h, w, hint_c, ctx_dim, 8)
    cond = {"c_latent": [c_latent.to(cuda)], "c_crossattn": [ctx.to(cuda)], "guide_hint": hint.to(cuda)}
    t = torch.full((B,), model.used_timesteps - 1, dtype=torch.long, device=cuda)
    x_T = model.q_sample(c_latent.to(cuda), t, noises[0].to(cuda))
    assert np.array_equal(x_T.cpu().numpy(), gold["x_T"])          # elementwise fp32: bit exact
    for steps in (2, 3, 5):
        s = SpacedSampler(model, var_type="fixed_small")
        s.noise_fn = lambda i, like: noises[1 + i]
        out = s.sample(steps, (B, 4, h, w), cond, x_T=x_T).cpu().numpy()
        e = rel_l2(out, gold[f"spaced_{steps}"])
        print(f"[{tag}] spaced {steps} steps rel-L2 {e:.3e}")
        assert e <= UNET_TOL
    s = SpacedSampler(model)
    s.noise_fn = lambda i, like: noises[1 + i]
    out = s.sample(2, (B, 4, h, w), cond, x_T=x_T, unconditional_guidance_scale=1.5).cpu().numpy()
    assert rel_l2(out, gold["spaced_2_cfg"]) <= UNET_TOL
    d = DDIMSampler(model)
    d.noise_fn = lambda i, like: noises[1 + i]
    out, inter = d.sample(S=2, batch_size=B, shape=(4, h, w), conditioning=cond, x_T=x_T, eta=0.0, verbose=False)
    assert rel_l2(out.cpu().numpy(), gold["ddim_2"]) <= UNET_TOL
    assert set(inter) == {"x_inter", "pred_x0"}
    d = DDIMSampler(model)
    d.noise_fn = lambda i, like: noises[1 + i]
    out, _ = d.sample(S=5, batch_size=B, shape=(4, h, w), conditioning=cond, x_T=x_T, eta=0.0, verbose=False)
    e = rel_l2(out.cpu().numpy(), gold["ddim_5"])
    print(f"[{tag}] ddim 5 steps rel-L2 {e:.3e}")
    assert e <= UNET_TOL


def test_small_unet_step_vs_reference_golden(small_model, cuda):
    _check_unet(small_model, "small", 32, 64, cuda)


def test_small_vae_vs_reference_golden(small_model, cuda):
    _check_vae(small_model, "small", cuda)


def test_small_samplers_vs_reference_golden(small_model, cuda):
    _check_samplers(small_model, "small", 2, 8, 16, 32, 64, cuda)


def test_small_step_vs_oracle_batched_odd_shape(small_model, cuda):
    """Batch > 1 with different per-sample timesteps and a non-square, non-power-of-two latent."""
    from oracle import nn as onn

    p = configs.small_params()
    sd = synthetic.make_state_dict(p, seed=231)
    B, h, w = 3, 8, 24
    c_latent, hint, ctx, noises = inputs(B, h, w, 32, 64, 1)
    t = torch.tensor([299, 150, 0], dtype=torch.long)
    with torch.no_grad():
        ref = onn.noise_estimator_forward(sd, noises[0], hint, t, ctx, model_channels=64, base_d_head=16, ctrl_d_head=16)
    cond = {"c_latent": [c_latent.to(cuda)], "c_crossattn": [ctx.to(cuda)], "guide_hint": hint.to(cuda)}
    eps = small_model.apply_model(noises[0].to(cuda), t.to(cuda), cond).cpu().numpy()
    e = rel_l2(eps, ref.numpy())
    print(f"[small] B=3 8x24 rel-L2 {e:.3e}")
    assert e <= UNET_TOL


def test_graph_and_eager_agree(small_model, cuda):
    B, h, w = 1, 16, 16
    c_latent, hint, ctx, noises = inputs(B, h, w, 32, 64, 1)
    cond = {"c_latent": [c_latent.to(cuda)], "c_crossattn": [ctx.to(cuda)], "guide_hint": hint.to(cuda)}
    t = torch.full((B,), 75, dtype=torch.long, device=cuda)
    a = small_model.apply_model(noises[0].to(cuda), t, cond)
    small_model.use_cuda_graph = False
    try:
        b = small_model.apply_model(noises[0].to(cuda), t, cond)
    finally:
        small_model.use_cuda_graph = True
    assert torch.equal(a, b)


def test_vae_decode_graph_and_eager_agree(small_model, cuda):
    """The graphed VAE decode (RDEIC._graphed_decode) replays the same kernels as the eager one: bit-identical
    images, new latents picked up on every replay, and latents above VAE_GRAPH_MAX_POSITIONS stay eager."""
    g = torch.Generator().manual_seed(11)
    zs = [torch.randn(2, 4, 8, 16, generator=g).to(cuda) for _ in range(2)]
    got = [(small_model.decode_first_stage_u8(z), small_model.decode_first_stage(z)) for z in zs]
    for kind in (True, False):                                      # one graph per (shape, output kind)
        assert ("vae", (2, 4, 8, 16), kind) in small_model._graphs
    small_model.use_cuda_graph = False
    try:
        want = [(small_model.decode_first_stage_u8(z), small_model.decode_first_stage(z)) for z in zs]
    finally:
        small_model.use_cuda_graph = True
    for (a8, af), (b8, bf) in zip(got, want):
        assert a8.dtype == torch.uint8 and tuple(a8.shape) == (2, 64, 128, 3) and torch.equal(a8, b8)
        assert tuple(af.shape) == (2, 3, 64, 128) and torch.equal(af, bf)
    assert not torch.equal(got[0][0], got[1][0])
    limit, type(small_model).VAE_GRAPH_MAX_POSITIONS = type(small_model).VAE_GRAPH_MAX_POSITIONS, 2 * 8 * 16 - 1
    try:
        small_model._graphs.pop(("vae", (2, 4, 8, 16), True))
        assert torch.equal(small_model.decode_first_stage_u8(zs[0]), want[0][0])
        assert ("vae", (2, 4, 8, 16), True) not in small_model._graphs   # ran eager: no graph was captured
    finally:
        type(small_model).VAE_GRAPH_MAX_POSITIONS = limit


def test_model_raises_without_weights_and_on_cpu(cuda):
    from rdeic_b200 import RDEIC, _lib

    m = RDEIC.from_config({"params": configs.small_params()}, device=cuda)
    with pytest.raises(RuntimeError):
        m.decode_first_stage(torch.zeros(1, 4, 8, 8, device=cuda))
    with pytest.raises(_lib.RdeicLibraryError):
        m.to("cpu")
    with pytest.raises(KeyError):
        m.load_state_dict({"model.diffusion_model.time_embed.0.weight": torch.zeros(256, 64)})


def test_full_unet_step_vs_reference_golden(full_model, cuda):
    _check_unet(full_model, "full", 256, 1024, cuda)


@pytest.mark.parametrize("hw", ["64x64", "64x96"])
def test_full_unet_step_at_baseline_latent_shapes(full_model, cuda, hw):
    """One full-width relay step at the latent shapes of BASELINE configs 2/5 (64x64) and 3 (64x96) against the
    reference's own output (tests/golden/full_unet_step_{hw}.npz, written by make_golden.py `big`)."""
    gold = np.load(GOLD / f"full_unet_step_{hw}.npz")
    h, w = (int(v) for v in gold["hw"])
    c_latent, hint, ctx, _ = inputs(1, h, w, 256, 1024, 1)
    cond = {"c_latent": [c_latent.to(cuda)], "c_crossattn": [ctx.to(cuda)], "guide_hint": hint.to(cuda)}
    x, t = torch.from_numpy(gold["x"]).to(cuda), torch.from_numpy(gold["t"]).to(cuda)
    e = rel_l2(full_model.apply_model(x, t, cond).cpu().numpy(), gold["eps"])
    print(f"[full {hw}] unet step rel-L2 {e:.3e}")
    assert e <= UNET_TOL


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_full_unet_step_parity_over_seeds(cuda, seed):
    """The bf16 margin under the 1e-2 bar must not hinge on one weight / input draw: five more seeds at full
    width and the BASELINE latent shape, against the fp32 kernel mode (itself 1.1e-6 from the reference's own
    output, test_fp32_kernel_mode_unet_step) as the on-GPU oracle."""
    from rdeic_b200 import RDEIC

    params = configs.default_params()
    sd = synthetic.make_state_dict(params, seed=seed, device=cuda)
    m16 = RDEIC.from_config({"params": params}, device=cuda, use_cuda_graph=False).load_state_dict(sd)
    m32 = RDEIC.from_config({"params": params}, device=cuda, precision="fp32").load_state_dict(sd)
    g = torch.Generator(device=cuda).manual_seed(seed + 100)
    B, h, w = 1, 64, 64
    x = torch.randn(B, 4, h, w, generator=g, device=cuda)
    cond = {"c_latent": [x], "c_crossattn": [torch.randn(B, 77, 1024, generator=g, device=cuda)],
            "guide_hint": torch.randn(B, 256, h, w, generator=g, device=cuda)}
    t = torch.full((B,), (37 * seed) % 300, dtype=torch.long, device=cuda)
    ref = m32.apply_model(x, t, cond).cpu().numpy()
    e = rel_l2(m16.apply_model(x, t, cond).cpu().numpy(), ref)
    print(f"[seed {seed}] full-width 64x64 step, t={int(t[0])}: bf16 vs fp32 mode rel-L2 {e:.3e}")
    assert e <= UNET_TOL
    del m16, m32, sd
    torch.cuda.empty_cache()


def test_full_vae_vs_reference_golden(full_model, cuda):
    _check_vae(full_model, "full", cuda)


def test_full_samplers_vs_reference_golden(full_model, cuda):
    _check_samplers(full_model, "full", 1, 16, 16, 256, 1024, cuda)


def test_c1_decode_vs_reference_golden(full_model, cuda):
    """BASELINE config[0] (one 256x256 image, 2 relay steps) end to end through relay_decode, against the
    image the reference's own q_sample -> SpacedSampler.sample -> decode_first_stage produced on the CPU in
    fp32 (tests/golden/full_c1_decode.npz): final latent rel-L2 <= 1e-2, image PSNR >= 40 dB."""
    from rdeic_b200.pipeline import relay_decode

    gold = np.load(GOLD / "full_c1_decode.npz")
    h, w = (int(v) for v in gold["hw"])
    steps = int(gold["steps"])
    c_latent, hint, ctx, noises = inputs(1, h, w, 256, 1024, 1 + steps)
    cond = {"c_latent": [c_latent.to(cuda)], "c_crossattn": [ctx.to(cuda)], "guide_hint": hint.to(cuda)}
    nz = [n.to(cuda) for n in noises]
    img = relay_decode(full_model, cond, steps, start_noise=nz[0], step_noises=nz[1:], as_uint8=False).cpu().numpy()
    ref = gold["img"].astype(np.float32)
    p = psnr(img, ref, 2.0)
    # the latent the samplers hand to the VAE, through the sampler API the reference driver uses
    from rdeic_b200 import SpacedSampler

    x_T = full_model.q_sample(c_latent.to(cuda), [full_model.used_timesteps - 1], nz[0])
    s = SpacedSampler(full_model, var_type="fixed_small")
    s.noise_fn = lambda i, like: nz[1 + i]
    z = s.sample(steps, (1, 4, h, w), cond, x_T=x_T).cpu().numpy()
    e = rel_l2(z, gold["z"])
    print(f"[c1 256^2, {steps} steps] vs reference: latent rel-L2 {e:.3e}, image PSNR {p:.1f} dB")
    assert e <= UNET_TOL and p >= PSNR_MIN
    u8 = relay_decode(full_model, cond, steps, start_noise=nz[0], step_noises=nz[1:]).cpu().numpy()
    ref_u8 = (((ref + 1) / 2).clip(0, 1).transpose(0, 2, 3, 1) * 255).clip(0, 255).astype(np.uint8)
    assert psnr(u8, ref_u8, 255.0) >= PSNR_MIN - 1.0


def test_tiled_decode_matches_per_tile_oracle(small_model, cuda):
    """BASELINE config 4 in miniature: one large latent decoded as overlapping tiles.  The oracle is
    the reference arithmetic applied per tile + the identical blend (the reference never tiles and
    a full-frame decode differs by construction: GroupNorm/attention are global — SURVEY.md §5)."""
    from oracle import nn as onn
    from oracle import sampler as osamp
    from rdeic_b200 import parallel
    from rdeic_b200.pipeline import relay_decode

    p = configs.small_params()
    sd = synthetic.make_state_dict(p, seed=231)
    h, w, tile, ov = 24, 40, 16, 8
    c_latent, hint, ctx, noises = inputs(1, h, w, 32, 64, 3)
    plan = parallel.plan_tiles(h, w, tile, ov)
    assert len(plan) > 2
    kw = dict(model_channels=64, base_d_head=16, ctrl_d_head=16)
    crop = lambda t, y0, x0, th, tw: t[:, :, y0:y0 + th, x0:x0 + tw].contiguous()

    ref_tiles = []
    with torch.no_grad():
        for (y0, x0, th, tw) in plan:
            cl, hn = crop(c_latent, y0, x0, th, tw), crop(hint, y0, x0, th, tw)
            ns = [crop(n, y0, x0, th, tw) for n in noises]
            x_T = osamp.q_sample(cl, 299, ns[0])
            z = osamp.spaced_sample(lambda x, t: onn.noise_estimator_forward(sd, x, hn, t, ctx, **kw), x_T, 2, ns[1:])
            ref_tiles.append(onn.vae_decode(sd, z)[0])
    ref = parallel.blend_tiles(ref_tiles, plan, h, w, ov, 8)

    cond = {"c_latent": [c_latent.to(cuda)], "c_crossattn": [ctx.to(cuda)], "guide_hint": hint.to(cuda)}
    dn = [n.to(cuda) for n in noises]

    def decode_fn(c, i):
        y0, x0, th, tw = plan[i]
        ns = [crop(n, y0, x0, th, tw) for n in dn]
        return relay_decode(small_model, c, 2, start_noise=ns[0], step_noises=ns[1:], as_uint8=False)

    img = parallel.decode_tiled(decode_fn, cond, tile=tile, overlap=ov).cpu()
    pq = psnr(img.numpy(), ref.numpy(), 2.0)
    print(f"[small] tiled decode ({len(plan)} tiles) PSNR vs per-tile oracle {pq:.1f} dB")
    assert pq >= PSNR_MIN

    def decode_batched(c, idx):         # all tiles of this rank in one batch (samples are independent)
        ns = [torch.cat([crop(n, *plan[i]) for i in idx], 0) for n in dn]
        return relay_decode(small_model, c, 2, start_noise=ns[0], step_noises=ns[1:], as_uint8=False)

    img_b = parallel.decode_tiled(decode_batched, cond, tile=tile, overlap=ov, batched=True).cpu()
    assert psnr(img_b.numpy(), ref.numpy(), 2.0) >= PSNR_MIN


def test_decode_host_batches_matches_relay_decode(small_model, cuda):
    """pipeline.decode_host_batches (double-buffered host -> device -> host serving loop): every yielded
    pinned-host image equals relay_decode of the same batch, in order, across slot and ring reuse, with a
    shape change in the middle of the stream and with an un-pinned batch."""
    from rdeic_b200.pipeline import decode_host_batches, relay_decode

    def batch(seed, B, h, w, pinned=True):
        g = torch.Generator().manual_seed(seed)
        t = {"c_latent": torch.randn(B, 4, h, w, generator=g), "guide_hint": torch.randn(B, 32, h, w, generator=g),
             "c_crossattn": torch.randn(B, 77, 64, generator=g), "start_noise": torch.randn(B, 4, h, w, generator=g),
             "step_noises": [torch.randn(B, 4, h, w, generator=g) for _ in range(2)]}
        if pinned:
            t = {k: [n.pin_memory() for n in v] if isinstance(v, list) else v.pin_memory() for k, v in t.items()}
        return t

    hbs = [batch(1, 2, 8, 8), batch(2, 2, 8, 8), batch(3, 2, 8, 8), batch(4, 1, 8, 16), batch(5, 1, 8, 16, pinned=False),
           batch(6, 2, 8, 8)]
    got = []
    for out in decode_host_batches(small_model, iter(hbs), 2):
        assert out.device.type == "cpu" and out.is_pinned() and out.dtype == torch.uint8
        got.append(out.clone())
    assert len(got) == len(hbs)
    for hb, img in zip(hbs, got):
        cond = {"c_latent": [hb["c_latent"].to(cuda)], "c_crossattn": [hb["c_crossattn"].to(cuda)],
                "guide_hint": hb["guide_hint"].to(cuda)}
        want = relay_decode(small_model, cond, 2, start_noise=hb["start_noise"].to(cuda),
                            step_noises=[n.to(cuda) for n in hb["step_noises"]]).cpu()
        assert tuple(img.shape) == tuple(want.shape) and torch.equal(img, want)
    assert not torch.equal(got[0], got[1])
    assert list(decode_host_batches(small_model, iter([]), 2)) == []


def test_baseline_size_batch8_512(full_model, cuda):
    """BASELINE config[1] at full size (512x512, batch 8, 5 relay steps) through size-independent
    properties plus one full-size oracle comparison:
      * run-to-run determinism (bit-identical uint8 images and latents);
      * batch independence: sample i of the batch-8 decode == the same sample decoded alone
        (GroupNorm / attention are per sample, util.py:224; tile shapes and split-K differ, a changed
        fp32 summation order flips bf16 roundings downstream, so the two agree at the bf16 noise level:
        >= 45 dB on the uint8 grid, measured 49.7);
      * one UNet+control step and one VAE decode of a full-size sample against the CPU oracle."""
    from oracle import nn as onn
    from rdeic_b200.pipeline import relay_decode

    B, h, w, steps = 8, 64, 64, 5
    c_latent, hint, ctx, noises = inputs(B, h, w, 256, 1024, steps + 1)
    d = lambda t: t.to(cuda)
    cond = {"c_latent": [d(c_latent)], "c_crossattn": [d(ctx)], "guide_hint": d(hint)}
    nz = [d(n) for n in noises]
    img1 = relay_decode(full_model, cond, steps, start_noise=nz[0], step_noises=nz[1:])
    img2 = relay_decode(full_model, cond, steps, start_noise=nz[0], step_noises=nz[1:])
    assert tuple(img1.shape) == (B, 512, 512, 3) and img1.dtype == torch.uint8
    assert torch.equal(img1, img2)
    i = 5
    solo = {"c_latent": [d(c_latent[i:i + 1])], "c_crossattn": [d(ctx[i:i + 1])], "guide_hint": d(hint[i:i + 1])}
    img_solo = relay_decode(full_model, solo, steps, start_noise=nz[0][i:i + 1], step_noises=[n[i:i + 1] for n in nz[1:]])
    p = psnr(img1[i].cpu().numpy(), img_solo[0].cpu().numpy(), 255.0)
    print(f"[full 512^2] batch-8 sample vs solo decode PSNR {p:.1f} dB")
    assert p >= 45.0
    # full-size oracle comparison on one sample (CPU fp32, a few seconds)
    p_ = configs.default_params()
    sd = synthetic.make_state_dict(p_, seed=231)
    kw = dict(model_channels=320, base_d_head=64, ctrl_d_head=16)
    t = torch.full((1,), 224, dtype=torch.long)
    with torch.no_grad():
        eps_ref = onn.noise_estimator_forward(sd, noises[0][i:i + 1], hint[i:i + 1], t, ctx[i:i + 1], **kw)
        img_ref = onn.vae_decode(sd, c_latent[i:i + 1])
    eps = full_model.apply_model(nz[0][i:i + 1], t.to(cuda), solo).cpu()
    e = rel_l2(eps.numpy(), eps_ref.numpy())
    img = full_model.decode_first_stage(d(c_latent[i:i + 1])).cpu()
    q = psnr(img.numpy(), img_ref.numpy(), 2.0)
    print(f"[full 512^2] unet step rel-L2 {e:.3e}; vae decode PSNR {q:.1f} dB")
    assert e <= UNET_TOL and q >= PSNR_MIN


FP32_TOL = 1e-5


@pytest.mark.parametrize("tag", ["small", "full"])
def test_fp32_kernel_mode_unet_step(cuda, tag):
    """north_star: per-step UNet output rel-L2 <= 1e-5 for the fp32 kernel mode, against the reference's
    own fp32 output (golden), conditional and unconditional branches."""
    from rdeic_b200 import RDEIC

    params = configs.small_params() if tag == "small" else configs.default_params()
    sd = synthetic.make_state_dict(params, seed=231)
    model = RDEIC.from_config({"params": params}, device=cuda, precision="fp32").load_state_dict(sd)
    gold = np.load(GOLD / f"{tag}_unet_step.npz")
    h, w = gold["hw"]
    hint_c = params["control_stage_config"]["params"]["hint_channels"]
    ctx_dim = params["unet_config"]["params"]["context_dim"]
    c_latent, hint, ctx, _ = inputs(1, h, w, hint_c, ctx_dim, 1)
    cond = {"c_latent": [c_latent.to(cuda)], "c_crossattn": [ctx.to(cuda)], "guide_hint": hint.to(cuda)}
    x, t = torch.from_numpy(gold["x"]).to(cuda), torch.from_numpy(gold["t"]).to(cuda)
    e1 = rel_l2(model.apply_model(x, t, cond).cpu().numpy(), gold["eps"])
    e2 = rel_l2(model.apply_model_unconditional(x, t, cond).cpu().numpy(), gold["eps_uncond"])
    print(f"[{tag}] fp32 kernel mode: unet step rel-L2 cond {e1:.3e} uncond {e2:.3e}")
    assert e1 <= FP32_TOL and e2 <= FP32_TOL


def test_fp32_kernel_mode_sampler(cuda):
    """The relay sampler on top of the fp32 step (small config, 2 and 3 steps) stays within 1e-5."""
    from rdeic_b200 import RDEIC, SpacedSampler

    params = configs.small_params()
    model = RDEIC.from_config({"params": params}, device=cuda, precision="fp32").load_state_dict(
        synthetic.make_state_dict(params, seed=231))
    gold = np.load(GOLD / "small_sampler.npz")
    B, h, w = 2, 8, 16
    c_latent, hint, ctx, noises = inputs(B, h, w, 32, 64, 8)
    cond = {"c_latent": [c_latent.to(cuda)], "c_crossattn": [ctx.to(cuda)], "guide_hint": hint.to(cuda)}
    t = torch.full((B,), model.used_timesteps - 1, dtype=torch.long, device=cuda)
    x_T = model.q_sample(c_latent.to(cuda), t, noises[0].to(cuda))
    for steps in (2, 3, 5):
        s = SpacedSampler(model, var_type="fixed_small")
        s.noise_fn = lambda i, like: noises[1 + i]
        e = rel_l2(s.sample(steps, (B, 4, h, w), cond, x_T=x_T).cpu().numpy(), gold[f"spaced_{steps}"])
        print(f"[small] fp32 kernel mode: spaced {steps} steps rel-L2 {e:.3e}")
        assert e <= FP32_TOL


@pytest.mark.parametrize("tag", ["small", "full"])
def test_fp32_kernel_mode_vae_and_decode(cuda, tag):
    """fp32 kernel mode of decode_first_stage against the reference golden, and the uint8 post-process."""
    from rdeic_b200 import RDEIC

    params = configs.small_params() if tag == "small" else configs.default_params()
    model = RDEIC.from_config({"params": params}, device=cuda, precision="fp32").load_state_dict(
        synthetic.make_state_dict(params, seed=231))
    gold = np.load(GOLD / f"{tag}_vae_decode.npz")
    z = torch.from_numpy(gold["z"]).to(cuda)
    img = model.decode_first_stage(z).cpu().numpy()
    e = rel_l2(img, gold["img"])
    print(f"[{tag}] fp32 kernel mode: vae decode rel-L2 {e:.3e} PSNR {psnr(img, gold['img'], 2.0):.1f} dB")
    assert e <= FP32_TOL
    u8 = model.decode_first_stage_u8(z).cpu().numpy()
    ref_u8 = (((gold["img"] + 1) / 2).clip(0, 1).transpose(0, 2, 3, 1) * 255).clip(0, 255).astype(np.uint8)
    assert float((u8 != ref_u8).mean()) < 1e-3          # a value within 1e-6 of an integer boundary may flip


def test_full_size_decode_bf16_vs_fp32_mode(full_model, cuda):
    """BASELINE config[1] end to end at full size: the 5-step 512x512 decode of the bf16 throughput mode
    against the same decode in the fp32 kernel mode (itself within 1e-6 of the reference per step):
    final image PSNR >= 40 dB (north_star), i.e. the bf16 error does not blow up over the relay steps."""
    from rdeic_b200 import RDEIC
    from rdeic_b200.pipeline import relay_decode

    params = configs.default_params()
    ref_model = RDEIC.from_config({"params": params}, device=cuda, precision="fp32").load_state_dict(
        synthetic.make_state_dict(params, seed=231))
    B, h, w, steps = 2, 64, 64, 5
    c_latent, hint, ctx, noises = inputs(B, h, w, 256, 1024, steps + 1)
    d = lambda t: t.to(cuda)
    cond = {"c_latent": [d(c_latent)], "c_crossattn": [d(ctx)], "guide_hint": d(hint)}
    nz = [d(n) for n in noises]
    img16 = relay_decode(full_model, cond, steps, start_noise=nz[0], step_noises=nz[1:], as_uint8=False).cpu().numpy()
    img32 = relay_decode(ref_model, cond, steps, start_noise=nz[0], step_noises=nz[1:], as_uint8=False).cpu().numpy()
    p = psnr(img16, img32, 2.0)
    print(f"[full 512^2, 5 steps] bf16 mode vs fp32 kernel mode: image PSNR {p:.1f} dB rel-L2 {rel_l2(img16, img32):.3e}")
    assert p >= PSNR_MIN
