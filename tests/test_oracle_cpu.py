"""CPU suite (`-m "not gpu"`): pins the oracle against the golden vectors produced by the
reference itself (tests/golden/make_golden.py), checks host-side logic (schedules, config and
checkpoint-layout handling) and that the C-ABI library loads and exports every declared symbol."""
import json
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import entropy as oe
from oracle import nn as onn
from oracle import sampler as osamp
from rdeic_b200 import configs, synthetic

GOLD = Path(__file__).resolve().parent / "golden"
ROOT = Path(__file__).resolve().parents[1]


def _unet_kw(params):
    up, cp = params["unet_config"]["params"], params["control_stage_config"]["params"]
    return dict(model_channels=up["model_channels"], base_d_head=up["num_head_channels"],
                ctrl_d_head=cp["num_head_channels"], control_scale=cp["control_scale"])


def _inputs(B, h, w, hint_c, ctx_dim, n_noise):
    g = lambda s: torch.Generator().manual_seed(s)
    c_latent = torch.randn(B, 4, h, w, generator=g(7))
    hint = torch.randn(B, hint_c, h, w, generator=g(8))
    ctx = torch.randn(B, 77, ctx_dim, generator=g(9))
    gn = g(231)
    return c_latent, hint, ctx, [torch.randn(B, 4, h, w, generator=gn) for _ in range(n_noise)]


# ---- known answers derived from the reference code alone -----------------------
def test_spaced_schedule_known_answers():
    s2, s5, s10 = (osamp.make_spaced_schedule(n) for n in (2, 5, 10))
    assert s2.timesteps.tolist() == [0, 299]
    assert s5.timesteps.tolist() == [0, 75, 150, 224, 299]
    assert s10.timesteps.tolist() == [0, 33, 66, 100, 133, 166, 199, 233, 266, 299]
    np.testing.assert_allclose(s2.alphas_cumprod, [0.99915, 0.5921831224], rtol=1e-9)
    np.testing.assert_allclose(s2.sqrt_recip_alphas_cumprod, [1.0004252711, 1.2994871436], rtol=1e-9)
    np.testing.assert_allclose(s2.sqrt_recipm1_alphas_cumprod, [0.0291671582, 0.8298595282], rtol=1e-8)
    np.testing.assert_allclose(s2.posterior_variance, [0, 0.00084895], atol=1e-8)
    np.testing.assert_allclose(s5.alphas_cumprod, [0.99915, 0.9240925093, 0.8278116322, 0.7159236019, 0.5921831224], rtol=1e-9)
    np.testing.assert_allclose(s5.posterior_variance, [0, 0.0008411969, 0.0459309485, 0.0819258131, 0.1203968355], atol=1e-9)
    np.testing.assert_allclose(s5.posterior_mean_coef1, [1, 0.9892226682, 0.581672462, 0.4328951433, 0.3586024557], rtol=1e-9)
    np.testing.assert_allclose(s5.posterior_mean_coef2, [0, 0.010769032, 0.4172428117, 0.5636846833, 0.6335262732], atol=1e-9)


def test_ddim_and_qsample_known_answers():
    assert osamp.make_ddim_timesteps(2, 300).tolist() == [1, 151]
    assert osamp.make_ddim_timesteps(5, 300).tolist() == [1, 61, 121, 181, 241]
    assert osamp.make_ddim_timesteps(10, 300).tolist()[-1] == 271
    b = osamp.ddpm_buffers()
    assert abs(float(b["sqrt_alphas_cumprod"][299]) - 0.7695343543) < 1e-7
    assert abs(float(b["sqrt_one_minus_alphas_cumprod"][299]) - 0.6386054162) < 1e-7
    assert abs(float(b["sqrt_recipm1_alphas_cumprod"][299]) - 0.8298595) < 1e-6


def test_scale_table_known_answers():
    t = oe.get_scale_table()
    assert t.shape == (64,) and t.dtype == np.float32
    np.testing.assert_allclose(t[[0, 1, 2, 3, 62, 63]], [0.11, 0.1244, 0.1407, 0.1591, 226.3591, 256.0], rtol=5e-4)


def test_product_samplers_match_oracle_schedule():
    from rdeic_b200.spaced_sampler_relay import SpacedSampler, space_timesteps
    from rdeic_b200.ddim_sampler_relay import make_ddim_timesteps

    class M:
        num_timesteps, used_timesteps, linear_start, linear_end = 1000, 300, 0.00085, 0.012

    for n in (1, 2, 3, 5, 10, 50, 300):
        assert space_timesteps(300, str(n)) == osamp.space_timesteps(300, str(n))
        s = SpacedSampler(M())
        s.make_schedule(n)
        o = osamp.make_spaced_schedule(n)
        for name in ("betas", "alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                     "posterior_variance", "posterior_mean_coef1", "posterior_mean_coef2"):
            assert np.array_equal(getattr(s, name), getattr(o, name)), (n, name)
        assert np.array_equal(s.timesteps, o.timesteps)
    assert space_timesteps(300, "ddim10") == osamp.space_timesteps(300, "ddim10")
    with pytest.raises(ValueError):
        space_timesteps(10, "20")
    for S in (2, 5, 10):
        assert np.array_equal(make_ddim_timesteps("uniform", S, 300), osamp.make_ddim_timesteps(S, 300))


# ---- oracle vs golden vectors from the reference (reduced-width config: fast) -----------------
@pytest.fixture(scope="module")
def small_sd():
    return synthetic.make_state_dict(configs.small_params(), seed=231)


def test_oracle_unet_step_matches_reference_golden(small_sd):
    p = configs.small_params()
    gold = np.load(GOLD / "small_unet_step.npz")
    h, w = gold["hw"]
    _, hint, ctx, _ = _inputs(1, h, w, 32, 64, 1)
    x, t = torch.from_numpy(gold["x"]), torch.from_numpy(gold["t"])
    with torch.no_grad():
        eps = onn.noise_estimator_forward(small_sd, x, hint, t, ctx, **_unet_kw(p))
        eps_u = onn.noise_estimator_forward(small_sd, x, hint, t, ctx, unconditional=True, **_unet_kw(p))
    assert np.abs(eps.numpy() - gold["eps"]).max() < 2e-5
    assert np.abs(eps_u.numpy() - gold["eps_uncond"]).max() < 2e-5


def test_oracle_vae_matches_reference_golden(small_sd):
    gold = np.load(GOLD / "small_vae_decode.npz")
    with torch.no_grad():
        img = onn.vae_decode(small_sd, torch.from_numpy(gold["z"]))
    assert np.abs(img.numpy() - gold["img"]).max() < 2e-5


def test_oracle_samplers_match_reference_golden(small_sd):
    p = configs.small_params()
    gold = np.load(GOLD / "small_sampler.npz")
    B, h, w = 2, 8, 16
    c_latent, hint, ctx, noises = _inputs(B, h, w, 32, 64, 8)
    x_T = osamp.q_sample(c_latent, 299, noises[0])
    assert np.array_equal(x_T.numpy(), gold["x_T"])
    kw = _unet_kw(p)
    am = lambda x, t: onn.noise_estimator_forward(small_sd, x, hint, t, ctx, **kw)
    amu = lambda x, t: onn.noise_estimator_forward(small_sd, x, hint, t, ctx, unconditional=True, **kw)
    with torch.no_grad():
        for steps in (2, 3):
            out = osamp.spaced_sample(am, x_T, steps, noises[1:1 + steps])
            assert np.abs(out.numpy() - gold[f"spaced_{steps}"]).max() < 5e-5, steps
        out = osamp.spaced_sample(am, x_T, 2, noises[1:3], apply_model_uncond=amu, guidance_scale=1.5)
        assert np.abs(out.numpy() - gold["spaced_2_cfg"]).max() < 5e-5
        out = osamp.ddim_sample(am, x_T, 2, noises[1:3])
        assert np.abs(out.numpy() - gold["ddim_2"]).max() < 5e-5


# ---- checkpoint layout contract ----------------------------------------------------------------
def test_oracle_c1_decode_matches_reference_golden():
    """BASELINE config[0] end to end at full width (one 256x256 image, 2 relay steps, fp32 CPU): the oracle's
    q_sample -> spaced_sample -> vae_decode against what the reference's own q_sample / SpacedSampler.sample /
    decode_first_stage produced (tests/golden/make_golden.py run_c1)."""
    p = configs.default_params()
    sd = synthetic.make_state_dict(p, seed=231)
    gold = np.load(GOLD / "full_c1_decode.npz")
    h, w = (int(v) for v in gold["hw"])
    steps = int(gold["steps"])
    c_latent, hint, ctx, noises = _inputs(1, h, w, 256, 1024, 1 + steps)
    kw = _unet_kw(p)
    with torch.no_grad():
        x_T = osamp.q_sample(c_latent, 299, noises[0])
        z = osamp.spaced_sample(lambda x, t: onn.noise_estimator_forward(sd, x, hint, t, ctx, **kw), x_T, steps, noises[1:])
        img = onn.vae_decode(sd, z)
    assert np.abs(z.numpy() - gold["z"]).max() < 1e-4
    err = np.abs(img.numpy() - gold["img"].astype(np.float32)).max()
    assert err < 2e-3, err                              # the fixture stores the image in fp16 (|x| < 2: ulp 1e-3)


def test_state_dict_spec_matches_reference_keys():
    ref = json.loads((GOLD / "state_dict_keys.json").read_text())
    spec = {k: list(s) for k, s, _ in synthetic.state_dict_spec(configs.default_params())}
    assert set(spec) == set(ref)
    for k in ref:
        assert spec[k] == ref[k], k
    counts = lambda pre: sum(k.startswith(pre) for k in spec)
    assert counts("model.diffusion_model.") == 686
    assert counts("control_model.control_model.") == 298
    assert counts("first_stage_model.decoder.") == 138


def test_normalise_state_dict_prefixes():
    from rdeic_b200.model import normalise_state_dict

    sd = {"state_dict": {"module.a.weight": torch.zeros(1), "b": torch.ones(1)}}
    out = normalise_state_dict(sd)
    assert set(out) == {"a.weight", "b"}


# ---- entropy oracle self-consistency (the arithmetic itself is trivial integer work) ------------
def test_entropy_oracle_roundtrip_properties():
    g = np.random.default_rng(0)
    y = g.normal(0, 6, (2, 8, 6, 10)).astype(np.float32)
    mu = g.normal(0, 2, y.shape).astype(np.float32)
    sc = np.exp(g.uniform(np.log(0.05), np.log(300), y.shape)).astype(np.float32)
    a, n = oe.ckbd_split(y)
    assert np.array_equal(oe.ckbd_merge(a, n), y)
    assert np.array_equal(oe.ckbd_anchor_unsequeeze(oe.ckbd_anchor_sequeeze(y)), a)
    assert np.array_equal(oe.ckbd_nonanchor_unsequeeze(oe.ckbd_nonanchor_sequeeze(y)), n)
    table = oe.get_scale_table()
    idx = oe.build_indexes(sc, table)
    assert idx.min() >= 0 and idx.max() <= 63
    assert np.array_equal(oe.build_indexes(table.copy(), table), np.arange(64).clip(max=63))
    assert oe.quantize_symbols(np.array([0.5, 1.5, 2.5, -0.5, -1.5], np.float32), None).tolist() == [0, 2, 2, 0, -2]
    for which in (0, 1):
        sym, i1, yhat = oe.compress_phase(y, sc, mu, table, which)
        msq, i2 = oe.decompress_phase_pre(sc, mu, table, which)
        assert np.array_equal(i1, i2)
        assert np.array_equal(oe.decompress_phase_post(sym, msq, which), yhat)


# ---- the C-ABI library -----------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    from rdeic_b200 import _lib, build

    build.build()
    lib = _lib.load()
    header = (ROOT / "include" / "rdeic_b200.h").read_text()
    declared = set(re.findall(r"\b(rdeic_[a-z0-9_]+)\s*\(", header))
    declared -= {"rdeic_conv_params", "rdeic_stream_t"}
    assert declared, "no symbols parsed from the header"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rdeic_abi_version() == 5


def test_no_cpu_fallback():
    from rdeic_b200 import _lib, ops

    with pytest.raises(_lib.RdeicLibraryError):
        ops.ckbd_mask(torch.zeros(1, 1, 2, 2), 0)


def test_product_does_not_import_oracle():
    for py in (ROOT / "rdeic_b200").glob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), py


def test_entropy_oracle_matches_reference_golden():
    """a9 / a11 pinned: oracle/entropy.py vs outputs of the reference's own utils/ckbd.py,
    utils/func.py and VectorQuantiser (tests/golden/entropy_ref.npz), bit for bit."""
    from helpers import entropy_golden_inputs

    gold = np.load(GOLD / "entropy_ref.npz")
    y, cb, z = entropy_golden_inputs()
    yn = y.numpy()
    b = lambda a: np.ascontiguousarray(a).view(np.uint32)
    assert np.array_equal(b(oe.ckbd_anchor(yn)), b(gold["anchor"]))
    assert np.array_equal(b(oe.ckbd_nonanchor(yn)), b(gold["nonanchor"]))
    assert np.array_equal(b(oe.ckbd_anchor_sequeeze(yn)), b(gold["anchor_sq"]))
    assert np.array_equal(b(oe.ckbd_nonanchor_sequeeze(yn)), b(gold["nonanchor_sq"]))
    assert np.array_equal(b(oe.ckbd_anchor_unsequeeze(gold["anchor_sq"])), b(gold["anchor_unsq"]))
    assert np.array_equal(b(oe.ckbd_nonanchor_unsequeeze(gold["nonanchor_sq"])), b(gold["nonanchor_unsq"]))
    assert np.array_equal(b(oe.ckbd_merge(*oe.ckbd_split(yn))), b(gold["merge"]))
    assert np.array_equal(b(oe.get_scale_table()), b(gold["scale_table"]))
    zq, idx = oe.vq_quant(z.numpy(), cb.numpy())
    assert np.array_equal(idx, gold["vq_idx"]) and idx.reshape(-1)[0] == 7
    assert np.array_equal(b(zq), b(gold["vq_zq"]))
    assert np.array_equal(b(oe.vq_lookup(idx, cb.numpy())), b(gold["vq_entry"]))


# ---- §8(f): learned compressor and bitstream container ----------------------------------------------
@pytest.mark.parametrize("tag", ["small", "full"])
def test_compression_spec_matches_reference_keys(tag):
    params = configs.small_params() if tag == "small" else configs.default_params()
    pp = params["preprocess_config"]["params"]
    ref = json.loads((GOLD / f"{tag}_compression_keys.json").read_text())
    spec = {k[len("preprocess_model."):]: list(s) for k, s, _ in synthetic.compression_state_dict_spec(pp)}
    assert spec == ref
    if tag == "full":
        assert len(spec) == 316


@pytest.mark.parametrize("tag", ["small", "full"])
def test_compression_oracle_matches_reference_golden(tag):
    """oracle/compression_nets.py vs the reference's own Compression.compress -> .decompress
    (tests/golden/*_compression.npz): integer outputs identical, float outputs to fp32 round-off."""
    from oracle import compression_nets as ocn

    params = configs.small_params() if tag == "small" else configs.default_params()
    pp = params["preprocess_config"]["params"]
    sd = synthetic.make_compression_state_dict(pp, seed=232)
    g = np.load(GOLD / f"{tag}_compression.npz")
    r = ocn.compress(sd, torch.from_numpy(g["x"]), pp["slice_ch"])
    close = lambda a, b: float(np.abs(np.asarray(a) - b).max()) <= 2e-4 * float(np.abs(b).max())
    assert close(r["y"], g["y"]) and close(r["hyper_params"], g["hyper_params"])
    assert np.array_equal(r["z_idx"].numpy(), g["z_idx"])
    # a symbol / index may flip only where fp32 summation order moved a value across a rounding boundary
    assert float((np.asarray(r["symbols"]) == g["symbols"]).mean()) > 0.999
    assert float((np.asarray(r["indexes"]) == g["indexes"]).mean()) > 0.999
    c, gh, y_hat, _ = ocn.decompress(sd, g["z_idx"], g["symbols"].tolist(), g["indexes"].tolist(), pp["slice_ch"])
    assert close(c, g["c_latent"]) and close(gh, g["guide_hint"])
    assert float((y_hat - torch.from_numpy(g["y"])).abs().max()) <= 0.5 + 1e-3
    assert len(np.unique(g["indexes"])) > 20 and int(np.abs(g["symbols"]).max()) > 8     # non-degenerate data


def test_bitstream_container_matches_reference_bytes(tmp_path):
    """rdeic_b200.utils read_body / write_body vs bytes written by the reference's utils/utils.py."""
    import io
    import struct

    from rdeic_b200 import utils

    data = (GOLD / "bitstream_ref.bin").read_bytes()
    strings, shape = utils.read_body(io.BytesIO(data))
    assert tuple(shape) == (8, 12) and [len(s[0]) for s in strings] == [1500, 41]
    f = io.BytesIO()
    assert utils.write_body(f, shape, strings) == len(data)
    assert f.getvalue() == data
    # edge cases: empty string list, zero-length string, truncated stream
    f = io.BytesIO()
    utils.write_body(f, (1, 2), [])
    assert f.getvalue() == struct.pack(">3I", 1, 2, 0) and utils.read_body(io.BytesIO(f.getvalue())) == ([], (1, 2))
    f = io.BytesIO()
    utils.write_body(f, (3, 3), [[b""], [b"x"]])
    assert utils.read_body(io.BytesIO(f.getvalue())) == ([[b""], [b"x"]], (3, 3))
    with pytest.raises(struct.error):
        utils.read_body(io.BytesIO(data[:100]))
    p = tmp_path / "a.bin"
    p.write_bytes(data)
    assert utils.filesize(str(p)) == len(data)
    with pytest.raises(ValueError):
        utils.filesize(str(tmp_path / "missing.bin"))


@pytest.mark.parametrize("tag", ["small", "full"])
def test_oracle_vae_encoder_matches_reference_golden(tag):
    params = configs.small_params() if tag == "small" else configs.default_params()
    sd = synthetic.make_state_dict(params, seed=231, encoder=True)
    g = np.load(GOLD / f"{tag}_vae_encode.npz")
    c = onn.vae_encode_hc(sd, torch.from_numpy(g["x"]))
    assert float((c - torch.from_numpy(g["c"])).abs().max()) < 1e-4
    if tag == "full":
        ref = json.loads((GOLD / "vae_encoder_keys.json").read_text())
        spec = {k: list(s) for k, s, _ in synthetic.state_dict_spec(params, encoder=True) if k in ref}
        assert spec == ref and len(ref) == 108                      # encoder 106 + quant_conv 2
        # appending the encoder leaves every decode-path tensor of the seeded checkpoint unchanged
        base = synthetic.make_state_dict(configs.small_params(), seed=231)
        ext = synthetic.make_state_dict(configs.small_params(), seed=231, encoder=True)
        assert all(torch.equal(base[k], ext[k]) for k in base)


def test_bench_arms_share_config():
    """bench.py prints the same `config` object from both arms (the driver compares them) and the reference arm's
    step is a real bounded sample: one image of the batch decoded completely by the oracle port."""
    import importlib
    import inspect

    bench = importlib.import_module("bench")
    c1, c2 = bench.bench_config(1), bench.bench_config(8)
    assert set(c1) == {"workload", "global_batch", "parallelism", "l2"} and c1["global_batch"] * 8 == c2["global_batch"]
    assert "bf16" not in c1["workload"] and "fp32" not in c1["workload"]          # the arithmetic type lives in `dtype`
    src = inspect.getsource(bench.run_reference) + inspect.getsource(bench.run_b200)
    assert src.count("bench_config(") == 2
    assert "s_per_image\"] * 1e3," in inspect.getsource(bench.run_reference)        # measured step time, not x BATCH
