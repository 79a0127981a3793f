"""GPU parity tests of the individual kernels, called through the C ABI (rdeic_b200.ops ->
librdeic_b200.so) and compared with the CPU oracle (oracle/) on the same seeded inputs.
Integer / byte / index work is bit-exact; bf16 tensor-core work states its tolerance."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import entropy as oe
from oracle import sampler as osamp
from oracle import nn as onn

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _rand_y(shape, seed, scale=6.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).float()


# ------------------------------------------------------------------------------------------
# entropy front end — bit exact
# ------------------------------------------------------------------------------------------
# squeezed widths 16 / 24 (256-bit loads), 12 (128-bit loads), 5 / 3 / 1 / 2 (scalar)
CK_SHAPES = [(1, 8, 32, 32), (2, 16, 7, 10), (1, 3, 5, 6), (3, 64, 32, 48), (1, 1, 1, 2), (1, 2, 9, 4), (2, 4, 6, 24)]


@pytest.mark.parametrize("shape", CK_SHAPES)
def test_ckbd_ops_bit_exact(cuda, shape):
    from rdeic_b200 import ops

    y = _rand_y(shape, 1)
    y.view(-1)[0] = float("nan")
    y.view(-1)[-1] = -0.0
    yn = y.numpy()
    yc = y.to(cuda)
    a, n = ops.ckbd_split(yc)
    assert np.array_equal(_bits(a.cpu().numpy()), _bits(oe.ckbd_anchor(yn)))
    assert np.array_equal(_bits(n.cpu().numpy()), _bits(oe.ckbd_nonanchor(yn)))
    assert np.array_equal(_bits(ops.ckbd_mask(yc, 0).cpu().numpy()), _bits(oe.ckbd_anchor(yn)))
    assert np.array_equal(_bits(ops.ckbd_mask(yc, 1).cpu().numpy()), _bits(oe.ckbd_nonanchor(yn)))
    # merge is a float add: NaN + 0 yields a NaN whose payload is platform-defined (x86 keeps the
    # operand's, CUDA returns the canonical 0x7fffffff, as torch's own CUDA add does) -> compare
    # NaN positions, and bits everywhere else.
    m = ops.ckbd_merge(a, n).cpu().numpy()
    rm = oe.ckbd_merge(oe.ckbd_anchor(yn), oe.ckbd_nonanchor(yn))
    assert np.array_equal(np.isnan(m), np.isnan(rm))
    assert np.array_equal(_bits(m)[~np.isnan(rm)], _bits(rm)[~np.isnan(rm)])
    sa = ops.ckbd_squeeze(yc, 0)
    sn = ops.ckbd_squeeze(yc, 1)
    assert np.array_equal(_bits(sa.cpu().numpy()), _bits(oe.ckbd_anchor_sequeeze(yn)))
    assert np.array_equal(_bits(sn.cpu().numpy()), _bits(oe.ckbd_nonanchor_sequeeze(yn)))
    ua = ops.ckbd_unsqueeze(sa, 0).cpu().numpy()
    un = ops.ckbd_unsqueeze(sn, 1).cpu().numpy()
    assert np.array_equal(_bits(ua), _bits(oe.ckbd_anchor_unsequeeze(oe.ckbd_anchor_sequeeze(yn))))
    assert np.array_equal(_bits(un), _bits(oe.ckbd_nonanchor_unsequeeze(oe.ckbd_nonanchor_sequeeze(yn))))


def test_ckbd_odd_width_raises(cuda):
    from rdeic_b200 import ops, _lib

    with pytest.raises(_lib.RdeicLibraryError):
        ops.ckbd_squeeze(torch.zeros(1, 1, 4, 5, device=cuda), 0)


def test_ckbd_empty(cuda):
    from rdeic_b200 import ops

    y = torch.zeros(0, 4, 8, 8, device=cuda)
    assert ops.ckbd_mask(y, 0).shape == y.shape
    assert ops.ckbd_squeeze(y, 1).shape == (0, 4, 8, 4)


def _entropy_inputs(shape, seed):
    g = torch.Generator().manual_seed(seed)
    y = torch.randn(shape, generator=g) * 6.0
    mu = torch.randn(shape, generator=g) * 2.0
    lo, hi = math.log(0.05), math.log(300.0)
    sc = torch.exp(torch.rand(shape, generator=g) * (hi - lo) + lo)
    # exact .5 ties for round-half-even, table hits, lower-bound edge, NaN scale
    flat_y, flat_mu, flat_sc = y.view(-1), mu.view(-1), sc.view(-1)
    n = flat_y.numel()
    k = min(n // 4, 64)
    flat_mu[:k] = torch.round(flat_mu[:k] * 4) / 4
    flat_y[:k] = flat_mu[:k] + (torch.arange(k).float() - k // 2) + 0.5
    table = torch.from_numpy(oe.get_scale_table())
    kk = min(n - k, 64)
    flat_sc[k:k + kk] = table[:kk]
    if n > k + kk + 4:
        flat_sc[k + kk] = 0.11
        flat_sc[k + kk + 1] = 0.0
        flat_sc[k + kk + 2] = float("nan")
        flat_sc[k + kk + 3] = 1e9
    return y.float(), mu.float(), sc.float(), table


@pytest.mark.parametrize("numel", [1, 5, 64, 1000, 1005, 262144 + 3])
def test_quantize_dequantize_indexes_bit_exact(cuda, numel):
    from rdeic_b200 import ops

    y, mu, sc, table = _entropy_inputs((numel,), 3)
    sym = ops.quantize_symbols(y.to(cuda), mu.to(cuda))
    ref = oe.quantize_symbols(y.numpy(), mu.numpy())
    assert np.array_equal(sym.cpu().numpy(), ref)
    sym0 = ops.quantize_symbols(y.to(cuda), None)
    assert np.array_equal(sym0.cpu().numpy(), oe.quantize_symbols(y.numpy(), None))
    deq = ops.dequantize(sym, mu.to(cuda))
    assert np.array_equal(_bits(deq.cpu().numpy()), _bits(oe.dequantize(ref, mu.numpy())))
    idx = ops.build_indexes(sc.to(cuda), table.to(cuda), 0.11)
    assert np.array_equal(idx.cpu().numpy(), oe.build_indexes(sc.numpy(), table.numpy()))
    if numel > 4:       # scales 16- but not 32-byte aligned: the 128-bit path
        shifted = torch.cat([sc[:4], sc]).to(cuda)[4:]
        assert shifted.data_ptr() % 32 == 16 and shifted.is_contiguous()
        idx = ops.build_indexes(shifted, table.to(cuda), 0.11)
        assert np.array_equal(idx.cpu().numpy(), oe.build_indexes(sc.numpy(), table.numpy()))


def test_build_indexes_unsorted_table_falls_back(cuda):
    from rdeic_b200 import ops

    table = torch.from_numpy(oe.get_scale_table()).clone()
    table[[5, 9]] = table[[9, 5]]
    _, _, sc, _ = _entropy_inputs((4096,), 5)
    idx = ops.build_indexes(sc.to(cuda), table.to(cuda), 0.11)
    assert np.array_equal(idx.cpu().numpy(), oe.build_indexes(sc.numpy(), table.numpy()))


@pytest.mark.parametrize("shape", [(1, 8, 32, 32), (2, 16, 7, 10), (1, 64, 32, 48), (1, 32, 128, 88)])
@pytest.mark.parametrize("which", [0, 1])
def test_fused_phase_kernels_bit_exact(cuda, shape, which):
    from rdeic_b200 import ops

    y, mu, sc, table = _entropy_inputs(shape, 11)
    tc = table.to(cuda)
    sym, idx, yhat = ops.ckbd_encode_phase(y.to(cuda), sc.to(cuda), mu.to(cuda), tc, 0.11, which)
    rsym, ridx, ryhat = oe.compress_phase(y.numpy(), sc.numpy(), mu.numpy(), table.numpy(), which)
    assert np.array_equal(sym.cpu().numpy(), rsym)
    assert np.array_equal(idx.cpu().numpy(), ridx)
    assert np.array_equal(_bits(yhat.cpu().numpy()), _bits(ryhat))
    msq, idx2 = ops.ckbd_squeeze_indexes(sc.to(cuda), mu.to(cuda), tc, 0.11, which)
    rmsq, ridx2 = oe.decompress_phase_pre(sc.numpy(), mu.numpy(), table.numpy(), which)
    assert np.array_equal(_bits(msq.cpu().numpy()), _bits(rmsq))
    assert np.array_equal(idx2.cpu().numpy(), ridx2)
    # decode side reproduces the encoder's y_hat from (symbols, means): the round trip
    yhat2 = ops.ckbd_decode_phase(sym, msq, which)
    assert np.array_equal(_bits(yhat2.cpu().numpy()), _bits(ryhat))
    assert np.array_equal(_bits(yhat2.cpu().numpy()), _bits(oe.decompress_phase_post(rsym, rmsq, which)))


def test_vq_quant_and_lookup(cuda):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(5)
    K, D = 1024, 256
    cb = (torch.rand(K, D, generator=g) * 2 - 1).float()
    cb[17] = cb[900]                       # exact tie -> first index must win
    pick = torch.randint(0, K, (2 * 3 * 5,), generator=g)
    pick[0] = 900
    z = cb[pick] + 0.01 * torch.randn(pick.numel(), D, generator=g)
    z[0] = cb[900]
    z = z.reshape(2, 3, 5, D).permute(0, 3, 1, 2).contiguous()
    zq, idx = ops.vq_quant(z.to(cuda), cb.to(cuda))
    rzq, ridx = oe.vq_quant(z.numpy(), cb.numpy())
    assert np.array_equal(idx.cpu().numpy(), ridx)
    assert idx.view(-1)[0].item() == 17
    assert np.array_equal(_bits(zq.cpu().numpy()), _bits(rzq))
    out = ops.vq_lookup(idx, cb.to(cuda))
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(oe.vq_lookup(ridx, cb.numpy())))


def test_vq_full_codebook_size(cuda):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(6)
    K, D = 16384, 256
    cb = ((torch.rand(K, D, generator=g) * 2 - 1) / 4).float()
    pick = torch.randint(0, K, (64,), generator=g)
    z = (cb[pick] + 0.002 * torch.randn(64, D, generator=g)).reshape(1, 8, 8, D).permute(0, 3, 1, 2).contiguous()
    zq, idx = ops.vq_quant(z.to(cuda), cb.to(cuda))
    rzq, ridx = oe.vq_quant(z.numpy(), cb.numpy())
    assert np.array_equal(idx.cpu().numpy(), ridx)
    assert np.array_equal(idx.cpu().numpy().reshape(-1), pick.numpy())
    assert np.array_equal(_bits(zq.cpu().numpy()), _bits(rzq))


# ------------------------------------------------------------------------------------------
# sampler updates — bit exact (same fp32 op order as the reference)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("steps", [2, 5, 10])
def test_relay_update_bit_exact(cuda, steps):
    from rdeic_b200 import ops

    sch = osamp.make_spaced_schedule(steps)
    g = torch.Generator().manual_seed(9)
    x, e, n = (torch.randn(2, 4, 16, 24, generator=g) for _ in range(3))
    for index in range(steps):
        ref = osamp.relay_update(x, e, n, sch, index)
        f = lambda v: float(np.float32(v))
        sigma = 0.0 if index == 0 else float(np.sqrt(np.float32(sch.posterior_variance[index])))
        out = ops.relay_update(x.to(cuda), e.to(cuda), n.to(cuda), f(sch.sqrt_recip_alphas_cumprod[index]),
                               f(sch.sqrt_recipm1_alphas_cumprod[index]), f(sch.posterior_mean_coef1[index]),
                               f(sch.posterior_mean_coef2[index]), sigma)
        assert np.array_equal(_bits(out.cpu().numpy()), _bits(ref.numpy())), index


def test_q_sample_bit_exact(cuda):
    from rdeic_b200 import ops

    b = osamp.ddpm_buffers()
    g = torch.Generator().manual_seed(10)
    x0, n = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    ref = osamp.q_sample(x0, 299, n, b)
    out = ops.q_sample(x0.to(cuda), n.to(cuda), float(b["sqrt_alphas_cumprod"][299]),
                       float(b["sqrt_one_minus_alphas_cumprod"][299]))
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(ref.numpy()))


# ------------------------------------------------------------------------------------------
# glue kernels
# ------------------------------------------------------------------------------------------
def _bf(x):
    return x.to(torch.bfloat16).float()


def test_layout_roundtrip_and_window(cuda):
    from rdeic_b200 import ops

    x = _rand_y((2, 5, 7, 9), 2, 1.0)
    nh = ops.nchw_to_nhwc_bf16(x.to(cuda), ldc=8)
    assert torch.equal(nh[..., :5].float().cpu(), _bf(x).permute(0, 2, 3, 1))
    assert torch.count_nonzero(nh[..., 5:]) == 0
    back = ops.nhwc_to_nchw_f32(nh, 5)
    assert torch.equal(back.cpu(), _bf(x))


def test_timestep_embedding(cuda):
    from rdeic_b200 import ops

    t = torch.tensor([0, 75, 150, 224, 299, 999], dtype=torch.int64)
    out = ops.timestep_embedding(t.to(cuda), 320).float().cpu()
    ref = onn.timestep_embedding(t, 320)
    assert torch.allclose(out, ref, atol=8e-3)  # bf16 output, |v| <= 1


def test_geglu_upsample_im2col_transpose_softmax(cuda):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(4)
    x = _bf(torch.randn(37, 2 * 64, generator=g))
    a, gate = x.chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    out = ops.geglu(x.to(cuda).bfloat16()).float().cpu()
    assert torch.allclose(out, ref, atol=2e-2, rtol=1e-2)

    u = _bf(torch.randn(2, 3, 5, 16, generator=g))
    up = ops.upsample2x(u.to(cuda).bfloat16()).float().cpu()
    ref = F.interpolate(u.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up, ref)

    c = _bf(torch.randn(2, 6, 8, 24, generator=g))
    col = ops.im2col_3x3_s2(c.to(cuda).bfloat16()).float().cpu()
    unf = F.unfold(c.permute(0, 3, 1, 2), 3, padding=1, stride=2)          # [B, C*9, L]
    unf = unf.reshape(2, 24, 9, -1).permute(0, 3, 2, 1)                    # [B, L, tap, C]
    ref = torch.zeros(2, 12, 9, 64)
    ref[..., :24] = unf
    assert torch.equal(col, ref.reshape(24, 9 * 64))

    t = _bf(torch.randn(3, 33, 50, generator=g))
    assert torch.equal(ops.transpose_bf16(t.to(cuda).bfloat16()).float().cpu(), t.transpose(1, 2))

    for n in (777, 64, 1024, 1100, 4096, 8192, 8196):      # scalar fallback and every register-resident width
        s = torch.randn(5, n, generator=g) * 3
        sm = ops.softmax_rows(s.to(cuda), 0.7).float().cpu()
        assert torch.allclose(sm, F.softmax(s * 0.7, dim=-1), atol=2e-3), n
        assert torch.allclose(sm.sum(-1), torch.ones(5), atol=2e-2), n


def test_image_to_u8(cuda):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 6, 5, 4, generator=g) * 0.8
    out = ops.image_to_u8(x.to(cuda)).cpu()
    ref = onn.to_uint8(x[..., :3].permute(0, 3, 1, 2))
    assert torch.equal(out, ref)


# ------------------------------------------------------------------------------------------
# normalisation (bf16 in/out, fp32 statistics): abs tol 3e-2 on O(1) outputs (bf16 eps 2^-8)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,C1,C2,silu,eps", [
    (2, 16, 16, 320, 0, True, 1e-5), (1, 8, 8, 1280, 1280, True, 1e-5), (2, 32, 32, 640, 320, True, 1e-5),
    (1, 12, 20, 64, 0, False, 1e-6), (1, 64, 64, 128, 0, True, 1e-6), (3, 4, 4, 256, 0, False, 1e-6),
    (1, 5, 3, 960, 0, True, 1e-5),
    # UNet level-1 sizes (groups of 20 480 elements: fold / statistics + apply)
    (2, 32, 32, 640, 0, True, 1e-5), (8, 32, 32, 640, 0, False, 1e-5), (1, 32, 32, 320, 320, True, 1e-5),
    # > 4 M elements or larger groups: the statistics + apply pair (smaller ones run as one kernel, one CTA per sample and group)
    (2, 128, 128, 320, 0, True, 1e-5), (1, 96, 96, 320, 320, True, 1e-5), (5, 64, 64, 256, 0, False, 1e-6)])
def test_groupnorm(cuda, B, H, W, C1, C2, silu, eps):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(12)
    C = C1 + C2
    x = _bf(torch.randn(B, H, W, C, generator=g) * 2 + 0.5)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = F.group_norm(x.permute(0, 3, 1, 2), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    for dt in (torch.bfloat16, torch.float32):      # bf16 activations, or the fp32 residual master
        x1 = x[..., :C1].contiguous().to(cuda).to(dt)
        x2 = x[..., C1:].contiguous().to(cuda).to(dt) if C2 else None
        out = ops.groupnorm(x1, gamma.to(cuda), beta.to(cuda), 32, eps, silu, x2=x2).float().cpu()
        assert torch.allclose(out, ref, atol=4e-2, rtol=2e-2), (dt, (out - ref).abs().max())


@pytest.mark.parametrize("rows,C", [(77, 320), (4096, 640), (5, 1280), (130, 64), (9, 128), (33, 256)])
def test_layernorm(cuda, rows, C):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(13)
    x = _bf(torch.randn(rows, C, generator=g) * 3 - 1)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = F.layer_norm(x, (C,), gamma, beta, 1e-5)
    for dt in (torch.bfloat16, torch.float32):
        out = ops.layernorm(x.to(cuda).to(dt), gamma.to(cuda), beta.to(cuda)).float().cpu()
        assert torch.allclose(out, ref, atol=4e-2, rtol=2e-2), (dt, (out - ref).abs().max())


# ------------------------------------------------------------------------------------------
# tensor-core contractions: inputs are exact bf16 values, accumulation fp32 -> fp32 outputs
# agree with an fp32 conv to 1e-3 relative (summation order only).
# ------------------------------------------------------------------------------------------
def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (256, 320, 320), (77 * 2, 1024, 640), (4096, 1280, 2560),
                                   (5, 320, 1280), (1000, 96, 4), (128, 2560, 160), (300, 64, 72)])
def test_linear_tc(cuda, M, K, N):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(21)
    x = _bf(torch.randn(M, K, generator=g))
    w = _bf(torch.randn(N, K, generator=g) / math.sqrt(K))
    b = torch.randn(N, generator=g)
    ref = F.linear(x, w, b)
    wp = ops.pack_conv_weight(w.to(cuda))
    out = ops.linear(x.to(cuda).bfloat16(), wp, N, bias=b.to(cuda), out_f32=True).cpu()
    assert _rel(out, ref) < 1e-3
    out16 = ops.linear(x.to(cuda).bfloat16(), wp, N, bias=b.to(cuda)).float().cpu()
    assert _rel(out16, ref) < 6e-3


@pytest.mark.parametrize("tile_n", [32, 64, 128, 160, 256])
def test_linear_tc_all_tile_widths(cuda, tile_n):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(22)
    M, K, N = 384, 448, 328
    x = _bf(torch.randn(M, K, generator=g))
    w = _bf(torch.randn(N, K, generator=g) / math.sqrt(K))
    ref = F.linear(x, w)
    wp = ops.pack_conv_weight(w.to(cuda))
    out = ops.linear(x.to(cuda).bfloat16(), wp, N, out_f32=True, tile_n=tile_n).cpu()
    assert _rel(out, ref) < 1e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 8, 64, 64), (2, 16, 16, 320, 320), (8, 8, 8, 128, 256),
                                            (1, 64, 64, 8, 320), (1, 32, 32, 264, 64), (3, 4, 4, 1280, 640),
                                            (1, 12, 24, 64, 32), (1, 6, 10, 64, 4), (1, 128, 128, 128, 3),
                                            (2, 5, 7, 72, 40)])
def test_conv3x3_tc(cuda, B, H, W, Cin, Cout):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(23)
    x = _bf(torch.randn(B, Cin, H, W, generator=g))
    w = _bf(torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin))
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x, w, b, padding=1).permute(0, 2, 3, 1)
    a = x.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16()
    wp = ops.pack_conv_weight(w.to(cuda))
    out = ops.conv_gemm(a, wp, Cout, 9, bias=b.to(cuda), out_f32=True).cpu()
    assert _rel(out, ref) < 1e-3


def test_conv_two_sources_epilogue(cuda):
    """concat input as two K segments + per-sample bias + SiLU-free residual/alpha epilogue."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(24)
    B, H, W, C1, C2, Cout = 2, 16, 16, 640, 320, 320
    x1 = _bf(torch.randn(B, C1, H, W, generator=g))
    x2 = _bf(torch.randn(B, C2, H, W, generator=g))
    for k in (3, 1):
        w = _bf(torch.randn(Cout, C1 + C2, k, k, generator=g) / math.sqrt(k * k * (C1 + C2)))
        bias = torch.randn(Cout, generator=g)
        rb = torch.randn(B, Cout, generator=g)
        resid = _bf(torch.randn(B, H, W, Cout, generator=g))
        conv = F.conv2d(torch.cat([x1, x2], 1), w, bias, padding=k // 2) + rb[:, :, None, None]
        ref = resid + 0.5 * conv.permute(0, 2, 3, 1)
        wp = ops.pack_conv_weight(w.to(cuda), c1=C1)
        out = ops.conv_gemm(x1.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16(), wp, Cout, k * k,
                            a2=x2.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16(), bias=bias.to(cuda),
                            row_bias=rb.to(cuda), resid=resid.to(cuda).bfloat16(), alpha=0.5, out_f32=True).cpu()
        assert _rel(out, ref) < 1e-3, k
        # fp32 residual master + dual (fp32, bf16) outputs written by one epilogue
        of, oh = ops.conv_gemm(x1.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16(), wp, Cout, k * k,
                               a2=x2.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16(), bias=bias.to(cuda),
                               row_bias=rb.to(cuda), resid=resid.to(cuda), alpha=0.5, dual=True)
        assert _rel(of.cpu(), ref) < 1e-3, k
        assert torch.equal(oh.float().cpu(), of.cpu().bfloat16().float())


def test_linear_silu_epilogue_and_batched(cuda):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(25)
    x = _bf(torch.randn(8, 320, generator=g))
    w = _bf(torch.randn(1280, 320, generator=g) / 18)
    b = torch.randn(1280, generator=g)
    ref = F.silu(F.linear(x, w, b))
    out = ops.linear(x.to(cuda).bfloat16(), ops.pack_conv_weight(w.to(cuda)), 1280, bias=b.to(cuda), act=1,
                     out_f32=True).cpu()
    assert _rel(out, ref) < 1e-3
    # batched: S_b = Q_b K_b^T (VAE mid attention, model.py:192)
    q = _bf(torch.randn(2, 256, 512, generator=g))
    k = _bf(torch.randn(2, 256, 512, generator=g))
    ref = torch.bmm(q, k.transpose(1, 2))
    kc = k.to(cuda).bfloat16().contiguous()
    out = ops.conv_gemm(q.to(cuda).bfloat16().reshape(2, 1, 256, 512), kc, 256, 1, w_batch_stride=256 * 512,
                        out_f32=True).cpu().reshape(2, 256, 256)
    assert _rel(out, ref) < 1e-3


@pytest.mark.parametrize("M,C", [(4096, 320), (300, 64), (77, 1280)])
def test_linear_geglu_epilogue(cuda, M, C):
    """ff.net.0.proj + GEGLU (attention.py:49-56) fused: interleaved weight rows, act = 2."""
    from rdeic_b200 import ops
    from rdeic_b200.engine import Conv

    g = torch.Generator().manual_seed(26)
    F_ = 4 * C
    x = _bf(torch.randn(M, C, generator=g))
    w = _bf(torch.randn(2 * F_, C, generator=g) / math.sqrt(C))
    b = torch.randn(2 * F_, generator=g) * 0.5
    a, gate = F.linear(x, w, b).chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    conv = Conv.load({"p.weight": w, "p.bias": b}, "p", cuda, geglu=True)
    out = ops.linear(x.to(cuda).bfloat16(), conv.w, conv.n_out, bias=conv.b, act=2).float().cpu()
    assert out.shape == ref.shape
    assert _rel(out, ref) < 6e-3


@pytest.mark.parametrize("M,K,N", [(512, 11520, 1280), (64, 4096, 256), (2048, 2560, 640)])
def test_split_k_matches_single_pass(cuda, M, K, N):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(27)
    x = _bf(torch.randn(M, K, generator=g))
    w = _bf(torch.randn(N, K, generator=g) / math.sqrt(K))
    b = torch.randn(N, generator=g)
    resid = torch.randn(M, N, generator=g)
    ref = resid + 0.75 * F.silu(F.linear(x, w, b))
    wp = ops.pack_conv_weight(w.to(cuda))
    kw = dict(bias=b.to(cuda), resid=resid.to(cuda), alpha=0.75, act=1, dual=True)
    of, oh = ops.linear(x.to(cuda).bfloat16(), wp, N, **kw)
    of1, _ = ops.linear(x.to(cuda).bfloat16(), wp, N, split_k=False, **kw)
    assert _rel(of.cpu(), ref) < 1e-3 and _rel(of1.cpu(), ref) < 1e-3
    assert torch.equal(oh.float().cpu(), of.cpu().bfloat16().float())
    # deterministic: a second split-K run is bit-identical
    of2, _ = ops.linear(x.to(cuda).bfloat16(), wp, N, **kw)
    assert torch.equal(of, of2)


# ------------------------------------------------------------------------------------------
# attention: bf16 P, fp32 softmax: abs tol 2e-2 on O(1) outputs.  The last two shapes have enough
# CTAs (>= 2 per SM at 256 query rows each) to take the two-query-tile "ping-pong" tcgen05 kernel.
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,heads,Nq,Nk,d", [(2, 5, 256, 256, 64), (1, 10, 100, 77, 64), (1, 4, 1024, 1024, 16),
                                             (2, 16, 64, 77, 16), (1, 20, 64, 64, 64), (1, 2, 4096, 4096, 64),
                                             (1, 3, 70, 130, 16), (8, 10, 1024, 1024, 64), (5, 15, 1024, 384, 64),
                                             # one KV tile (cross-attention on the text keys): tcgen05 kernel without a key loop
                                             (8, 5, 4096, 77, 64), (2, 20, 64, 77, 64), (1, 3, 200, 128, 64), (2, 2, 130, 1, 64),
                                             (1, 4, 256, 16, 64), (3, 1, 5, 100, 64)])
def test_attention(cuda, B, heads, Nq, Nk, d):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(31)
    q = _bf(torch.randn(B, Nq, heads * d, generator=g))
    k = _bf(torch.randn(B, Nk, heads * d, generator=g))
    v = _bf(torch.randn(B, Nk, heads * d, generator=g))
    sp = lambda t: t.reshape(B, t.shape[1], heads, d).permute(0, 2, 1, 3)
    sim = torch.einsum("bhid,bhjd->bhij", sp(q), sp(k)) * d ** -0.5
    ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), sp(v)).permute(0, 2, 1, 3).reshape(B, Nq, heads * d)
    out = ops.attention(q.to(cuda).bfloat16(), k.to(cuda).bfloat16(), v.to(cuda).bfloat16(), heads, d,
                        d ** -0.5).float().cpu()
    assert (out - ref).abs().max() < 2e-2
    assert _rel(out, ref) < 1e-2


@pytest.mark.parametrize("B,heads,N", [(8, 10, 1024), (2, 5, 4096)])
def test_attention_growing_logits(cuda, B, heads, N):
    """Logits that keep growing along the key axis (and differently per row): the two-query-tile kernel keeps O
    in TMEM and rescales it only when a row's maximum has grown by 2^8 — every row must rescale several
    times here, some lanes of a warp without the others."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(33)
    d = 64
    q = _bf(torch.randn(B, N, heads * d, generator=g) * (0.5 + 3.0 * torch.rand(B, N, 1, generator=g)))
    k = _bf(torch.randn(B, N, heads * d, generator=g) * torch.linspace(0.3, 4.0, N)[None, :, None])
    v = _bf(torch.randn(B, N, heads * d, generator=g))
    sp = lambda t: t.reshape(B, N, heads, d).permute(0, 2, 1, 3)
    sim = torch.einsum("bhid,bhjd->bhij", sp(q), sp(k)) * d ** -0.5
    assert float(sim.max()) > 30.0
    ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), sp(v)).permute(0, 2, 1, 3).reshape(B, N, heads * d)
    out = ops.attention(q.to(cuda).bfloat16(), k.to(cuda).bfloat16(), v.to(cuda).bfloat16(), heads, d,
                        d ** -0.5).float().cpu()
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max() < 6e-2 and _rel(out, ref) < 1e-2


@pytest.mark.parametrize("B,heads,Nq,Nk,d,grow", [(2, 1, 256, 384, 512, False), (1, 1, 1024, 1024, 512, False),
                                                  (2, 2, 128, 256, 256, False), (1, 1, 384, 4096, 512, True),
                                                  (3, 1, 128, 128, 512, False)])
def test_attention_wide_heads(cuda, B, heads, Nq, Nk, d, grow):
    """The VAE mid-block attention (ldm/modules/diffusionmodules/model.py:181-205: one head, d = C = 512) as a flash
    kernel: value dimension split over two CTAs, O resident in tensor memory with lazy rescaling (grow = logits that
    keep rising along the key axis, so every row rescales several times)."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(35)
    q = _bf(torch.randn(B, Nq, heads * d, generator=g))
    k = _bf(torch.randn(B, Nk, heads * d, generator=g))
    if grow:
        q = _bf(q * (0.5 + 2.0 * torch.rand(B, Nq, 1, generator=g)))
        k = _bf(k * torch.linspace(0.3, 4.0, Nk)[None, :, None])
    v = _bf(torch.randn(B, Nk, heads * d, generator=g))
    sp = lambda t: t.reshape(B, t.shape[1], heads, d).permute(0, 2, 1, 3)
    sim = torch.einsum("bhid,bhjd->bhij", sp(q), sp(k)) * d ** -0.5
    if grow:
        assert float(sim.max()) > 30.0
    ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), sp(v)).permute(0, 2, 1, 3).reshape(B, Nq, heads * d)
    out = ops.attention(q.to(cuda).bfloat16(), k.to(cuda).bfloat16(), v.to(cuda).bfloat16(), heads, d,
                        d ** -0.5).float().cpu()
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max() < (6e-2 if grow else 2e-2)
    assert _rel(out, ref) < 1e-2


def test_attention_wide_rejects_ragged(cuda):
    from rdeic_b200 import ops
    from rdeic_b200._lib import RdeicLibraryError

    t = torch.zeros(1, 100, 512, device=cuda, dtype=torch.bfloat16)
    with pytest.raises(RdeicLibraryError):
        ops.attention(t, t, t, 1, 512, 512 ** -0.5)


def test_attention_fused_qkv_slices(cuda):
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(32)
    B, N, heads, d = 2, 128, 5, 64
    C = heads * d
    qkv = _bf(torch.randn(B, N, 3 * C, generator=g))
    q, k, v = qkv.chunk(3, -1)
    sp = lambda t: t.reshape(B, N, heads, d).permute(0, 2, 1, 3)
    sim = torch.einsum("bhid,bhjd->bhij", sp(q), sp(k)) * d ** -0.5
    ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), sp(v)).permute(0, 2, 1, 3).reshape(B, N, C)
    t = qkv.to(cuda).bfloat16()
    out = ops.attention(t[..., :C], t[..., C:2 * C], t[..., 2 * C:], heads, d, d ** -0.5).float().cpu()
    assert (out - ref).abs().max() < 2e-2


# ---------------------------------------------------------------------------------------------
# compressor layers: 5x5 taps, LeakyReLU / GELU epilogues, channel-slice operands,
# sub-pixel shuffle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 32, 32, 8, 16), (2, 8, 12, 224, 128), (1, 16, 16, 128, 64),
                                            (1, 5, 7, 16, 224), (1, 64, 96, 8, 224)])
def test_conv5x5_tc(cuda, B, H, W, Cin, Cout):
    """25-tap implicit GEMM == F.conv2d(kernel 5, padding 2) (model/compression.py:23, compression_modules.py:80-84)."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(61)
    x = _bf(torch.randn(B, Cin, H, W, generator=g))
    w = _bf(torch.randn(Cout, Cin, 5, 5, generator=g) / math.sqrt(25 * Cin))
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x, w, b, padding=2).permute(0, 2, 3, 1)
    a = x.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16()
    out = ops.conv_gemm(a, ops.pack_conv_weight(w.to(cuda)), Cout, 25, bias=b.to(cuda), out_f32=True).cpu()
    assert _rel(out, ref) < 1e-3


@pytest.mark.parametrize("act,slope", [(3, 0.01), (3, 0.1), (4, 0.0)])
def test_conv_leaky_relu_gelu_epilogues(cuda, act, slope):
    """act(conv + bias) + residual, the order of res_blk.py:30-36,57-62,88-95; also through split-K."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(62)
    for (B, H, W, Cin, Cout, k) in [(1, 16, 16, 256, 384, 3), (2, 8, 8, 528, 26, 1), (1, 4, 4, 512, 512, 3)]:
        x = _bf(torch.randn(B, Cin, H, W, generator=g))
        w = _bf(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(k * k * Cin))
        b = torch.randn(Cout, generator=g)
        resid = _bf(torch.randn(B, H, W, Cout, generator=g))
        pre = F.conv2d(x, w, b, padding=k // 2)
        ref = (F.leaky_relu(pre, slope) if act == 3 else F.gelu(pre)).permute(0, 2, 3, 1) + resid
        a = x.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16()
        n_pad = (Cout + 7) // 8 * 8
        wpad = torch.cat([w, torch.zeros(n_pad - Cout, Cin, k, k)], 0)
        bpad = torch.cat([b, torch.zeros(n_pad - Cout)], 0)
        rpad = torch.cat([resid, torch.zeros(B, H, W, n_pad - Cout)], -1)
        out = ops.conv_gemm(a, ops.pack_conv_weight(wpad.to(cuda)), n_pad, k * k, bias=bpad.to(cuda), act=act,
                            act_param=slope, resid=rpad.to(cuda).bfloat16(), out_f32=True).cpu()
        assert _rel(out[..., :Cout], ref) < 1e-3, (Cin, Cout, k)
        assert torch.count_nonzero(out[..., Cout:]) == 0           # padded channels stay exactly 0


def test_conv_channel_slice_operands(cuda):
    """A, a2 and the output as channel windows of wider NHWC buffers (a_ld / a2_ld / ldo): the
    torch.cat inputs of compression.py:170-190 without materialising them."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(63)
    B, H, W = 2, 8, 12
    wide = _bf(torch.randn(B, H, W, 96, generator=g))               # use channels [16, 48)
    hyper = _bf(torch.randn(B, H, W, 64, generator=g))
    c1, c2, n = 32, 64, 16
    w = _bf(torch.randn(n, c1 + c2, 1, 1, generator=g) / math.sqrt(c1 + c2))
    b = torch.randn(n, generator=g)
    x = torch.cat([wide[..., 16:48], hyper], -1).permute(0, 3, 1, 2)
    ref = F.conv2d(x, w, b).permute(0, 2, 3, 1)
    wd, hd = wide.to(cuda).bfloat16(), hyper.to(cuda).bfloat16()
    dst = torch.full((B, H, W, 40), 7.0, device=cuda, dtype=torch.bfloat16)
    ops.conv_gemm(wd[..., 16:48], ops.pack_conv_weight(w.to(cuda), c1=c1), n, 1, a2=hd, bias=b.to(cuda), out=dst[..., 8:24])
    got = dst.float().cpu()
    assert _rel(got[..., 8:24], ref) < 6e-3                         # bf16 output rounding
    assert torch.all(got[..., :8] == 7.0) and torch.all(got[..., 24:] == 7.0)
    # 5x5 over a channel prefix (channel_context reads y_hat[..., :sum(slice_ch[:i])] in place)
    w5 = _bf(torch.randn(24, 40, 5, 5, generator=g) / math.sqrt(25 * 40))
    ref5 = F.conv2d(wide[..., :40].permute(0, 3, 1, 2), w5, None, padding=2).permute(0, 2, 3, 1)
    out5 = ops.conv_gemm(wd[..., :40], ops.pack_conv_weight(w5.to(cuda)), 24, 25, out_f32=True).cpu()
    assert _rel(out5, ref5) < 1e-3
    with pytest.raises(TypeError):
        ops.conv_gemm(wd[:, ::2], ops.pack_conv_weight(w5.to(cuda)), 24, 25)        # rows skipped: not one pixel stride


def test_pixel_shuffle2_bit_exact(cuda):
    """sub-pixel conv tail (model/layers/conv.py:7-10): (i, j, c)-ordered channels -> nn.PixelShuffle(2)."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(64)
    B, H, W, C = 2, 3, 5, 24
    x = _bf(torch.randn(B, 4 * C, H, W, generator=g))                # torch order: c*4 + 2i + j
    ref = F.pixel_shuffle(x, 2).permute(0, 2, 3, 1)
    xp = x.view(B, C, 4, H, W).transpose(1, 2).reshape(B, 4 * C, H, W)      # (2i + j)*C + c
    out = ops.pixel_shuffle2(xp.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16()).float().cpu()
    assert torch.equal(out, ref)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,f32", [(16, 32, 32, 128, 128, 3, False), (8, 8, 8, 64, 320, 1, True),
                                                    (1, 64, 64, 320, 320, 1, True), (16, 32, 32, 256, 64, 3, False),
                                                    (1, 128, 128, 128, 128, 3, False), (2, 64, 64, 320, 320, 3, True)])
def test_groupnorm_from_fused_conv_stats(cuda, B, H, W, Cin, Cout, k, f32):
    """GroupNorm statistics emitted by the producing GEMM's epilogue == a statistics pass over the
    tensor it wrote (util.py:224 / model.py:48 semantics unchanged), single and two-source."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(71)
    x = _bf(torch.randn(B, H, W, Cin, generator=g)).to(cuda).bfloat16()
    w = ops.pack_conv_weight((_bf(torch.randn(Cout, Cin, k, k, generator=g)) / math.sqrt(k * k * Cin)).to(cuda))
    b = torch.randn(Cout, generator=g).to(cuda)
    resid = torch.randn(B, H, W, Cout, generator=g).to(cuda)
    resid = resid if f32 else resid.bfloat16()
    assert ops.conv_stats_supported(B, H, W, Cout, k * k * ((Cin + 63) // 64))
    if f32:
        of, oh, st = ops.conv_gemm(x, w, Cout, k * k, bias=b, resid=resid, dual=True, stats=True)
        y = of
    else:
        y, st = ops.conv_gemm(x, w, Cout, k * k, bias=b, resid=resid, stats=True)
    assert st is not None and tuple(st.shape) == (B * H * W // 32, Cout, 2)
    # slab sums against torch on the stored tensor (fp32 master exactly; bf16 output to rounding)
    ref = y.float().view(B * H * W // 32, 32, Cout)
    tol = 1e-4 if f32 else 2e-2
    assert _rel(st[..., 0].cpu(), ref.sum(1).cpu()) < tol and _rel(st[..., 1].cpu(), (ref * ref).sum(1).cpu()) < tol
    gamma, beta = (1 + 0.1 * torch.randn(Cout, generator=g)).to(cuda), (0.1 * torch.randn(Cout, generator=g)).to(cuda)
    a = ops.groupnorm(y, gamma, beta, 32, 1e-5, True)
    f = ops.groupnorm(y, gamma, beta, 32, 1e-5, True, stats1=st)
    assert _rel(f.float().cpu(), a.float().cpu()) < (2e-3 if f32 else 6e-3)
    # two sources (decoder concat, openaimodel.py:804): same tensor twice
    g2, b2 = torch.cat([gamma, gamma]), torch.cat([beta, beta])
    a2 = ops.groupnorm(y, g2, b2, 32, 1e-5, False, x2=y)
    f2 = ops.groupnorm(y, g2, b2, 32, 1e-5, False, x2=y, stats1=st, stats2=st)
    assert _rel(f2.float().cpu(), a2.float().cpu()) < (2e-3 if f32 else 6e-3)


def test_conv_stats_unsupported_grid_falls_back(cuda):
    from rdeic_b200 import ops

    assert not ops.conv_stats_supported(1, 6, 10)            # ragged tiles
    assert not ops.conv_stats_supported(8, 8, 8, 1280, 9 * 20)     # a split-K layer keeps split-K
    assert ops.conv_stats_supported(8, 8, 8, 1280, 4)
    x = torch.randn(1, 6, 10, 64, device=cuda).bfloat16()
    w = ops.pack_conv_weight(torch.randn(64, 64, 3, 3, device=cuda) / 24)
    y, st = ops.conv_gemm(x, w, 64, 9, stats=True)
    assert st is None and tuple(y.shape) == (1, 6, 10, 64)


@pytest.mark.parametrize("B,H,W,C1,C2,Cout,stats", [(2, 288, 288, 128, 0, 128, True), (1, 1297, 128, 128, 0, 128, False),
                                                     (1, 1297, 128, 64, 64, 128, True), (2, 288, 288, 256, 0, 96, False)])
def test_conv3x3_two_m_tiles_per_item(cuda, B, H, W, C1, C2, Cout, stats):
    """N <= 128 convs with many tiles run two M tiles per work item against one weight stage (even and
    odd tile counts, two K segments, fused statistics); the result must not depend on the pairing."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(91)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        x = _bf(torch.randn(B, C1 + C2, H, W, generator=g)).to(cuda)
        w = _bf(torch.randn(Cout, C1 + C2, 3, 3, generator=g) / math.sqrt(9 * (C1 + C2))).to(cuda)
        b = torch.randn(Cout, generator=g).to(cuda)
        resid = _bf(torch.randn(B, H, W, Cout, generator=g)).to(cuda)
        ref = resid + F.conv2d(x, w, b, padding=1).permute(0, 2, 3, 1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    a = x.permute(0, 2, 3, 1).contiguous().bfloat16()
    a1 = a[..., :C1].contiguous()
    a2 = a[..., C1:].contiguous() if C2 else None
    wp = ops.pack_conv_weight(w, c1=C1) if C2 else ops.pack_conv_weight(w)
    res = ops.conv_gemm(a1, wp, Cout, 9, a2=a2, bias=b, resid=resid.bfloat16(), out_f32=not stats, stats=stats)
    if stats:
        y, st = res
        assert st is not None
        ref_b = y.float().view(B * H * W // 32, 32, Cout)
        assert _rel(st[..., 0].cpu(), ref_b.sum(1).cpu()) < 2e-2 and _rel(st[..., 1].cpu(), (ref_b * ref_b).sum(1).cpu()) < 2e-2
        assert _rel(y.float().cpu(), ref.cpu()) < 4e-3            # bf16 output rounding
    else:
        assert _rel(res.cpu(), ref.cpu()) < 1e-3


@pytest.mark.parametrize("B,H,W,C1,C2,Cout,stats", [(8, 64, 64, 128, 0, 320, True), (5, 64, 64, 128, 0, 320, False),
                                                     (3, 64, 96, 64, 64, 640, True), (1, 129, 128, 192, 0, 320, False),
                                                     (8, 32, 32, 128, 0, 640, True)])
def test_conv3x3_two_issuers_n160(cuda, B, H, W, C1, C2, Cout, stats):
    """N = 160 tiles with a long reduction and >= 2 rounds of tiles (UNet levels 0 / 1: openaimodel.py:203,229 at 320 /
    640 channels): even and odd tile counts, two K segments, fused statistics, fp32 + bf16 outputs, bit-identical run to
    run.  These are the shapes the paired-tile variants (kNT = 2 / kIss = 2, `test_conv_variants_behind_switches`) take
    over when their switch is on; each accumulator there sees its own in-order instruction stream."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(92)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        x = _bf(torch.randn(B, C1 + C2, H, W, generator=g)).to(cuda)
        w = _bf(torch.randn(Cout, C1 + C2, 3, 3, generator=g) / math.sqrt(9 * (C1 + C2))).to(cuda)
        b = torch.randn(Cout, generator=g).to(cuda)
        resid = torch.randn(B, H, W, Cout, generator=g).to(cuda)
        ref = resid + F.conv2d(x, w, b, padding=1).permute(0, 2, 3, 1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    a = x.permute(0, 2, 3, 1).contiguous().bfloat16()
    a1 = a[..., :C1].contiguous()
    a2 = a[..., C1:].contiguous() if C2 else None
    wp = ops.pack_conv_weight(w, c1=C1) if C2 else ops.pack_conv_weight(w)
    run = lambda: ops.conv_gemm(a1, wp, Cout, 9, a2=a2, bias=b, resid=resid, dual=True, stats=stats)
    res = run()
    yf, yb = res[0], res[1]
    assert _rel(yf.cpu(), ref.cpu()) < 1e-3
    assert torch.equal(yb.float(), yf.bfloat16().float())
    if stats:
        st = res[2]
        assert st is not None
        ref_b = yf.view(B * H * W // 32, 32, Cout)
        assert _rel(st[..., 0].cpu(), ref_b.sum(1).cpu()) < 1e-3 and _rel(st[..., 1].cpu(), (ref_b * ref_b).sum(1).cpu()) < 1e-3
    res2 = run()
    assert torch.equal(res2[0], yf)


@pytest.mark.parametrize("env", ["RDEIC_PAIR160", "RDEIC_DUAL160", "RDEIC_M_MAJOR"])
def test_conv_variants_behind_switches(cuda, env):
    """The conv variants that lost their A/B stay in the library behind environment switches (two N tiles per item sharing
    the activation stage; two M tiles per item with one issuing warp each; activation-major item order).  The switches are
    read once per process, so the N = 160 cases above are re-run in a child process with the switch on."""
    import os
    import subprocess
    import sys

    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider", __file__, "-k",
                        "two_issuers_n160 or conv3x3_tc or conv_two_sources"], env=dict(os.environ, **{env: "1"}),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (1, 40, 48), (1, 8, 100)])
def test_vae_tail_fused_matches_unfused(cuda, B, H, W):
    """norm_out + swish + conv_out (+ uint8) in one kernel == GroupNorm kernel -> tcgen05 conv -> image_to_u8
    (model.py:683-686, inference.py:85-87): same bf16 operand rounding, fp32 sums in another order."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(93)
    C = 128
    xin = _bf(torch.randn(B, H, W, 64, generator=g)).to(cuda).bfloat16()
    wprod = ops.pack_conv_weight((_bf(torch.randn(C, 64, 3, 3, generator=g)) / 24).to(cuda))
    x, st = ops.conv_gemm(xin, wprod, C, 9, stats=True)
    if st is None:       # ragged pixel grid: statistics by a pass over the tensor, emitted in the slab layout
        xs = x.float().view(B * H * W // 32, 32, C)
        st = torch.stack([xs.sum(1), (xs * xs).sum(1)], -1).contiguous()
    gamma, beta = (1 + 0.1 * torch.randn(C, generator=g)).to(cuda), (0.1 * torch.randn(C, generator=g)).to(cuda)
    w = (torch.randn(3, C, 3, 3, generator=g) / 34).to(cuda)
    b = (0.1 * torch.randn(3, generator=g)).to(cuda)
    w4 = torch.cat([w, torch.zeros(1, C, 3, 3, device=cuda)], 0)
    b4 = torch.cat([b, torch.zeros(1, device=cuda)], 0)
    xn = ops.groupnorm(x, gamma, beta, 32, 1e-6, True, stats1=st)
    ref = ops.conv_gemm(xn, ops.pack_conv_weight(w4), 4, 9, bias=b4, out_f32=True)
    # and against torch on the same normalised bf16 activations
    tref = F.conv2d(xn.float().permute(0, 3, 1, 2), _bf(w.cpu()).to(cuda), b, padding=1).permute(0, 2, 3, 1)
    wt = ops.pack_tail_weight(w)
    out = ops.gn_silu_conv3x3_tail(x, st, gamma, beta, 32, 1e-6, wt, b, 3, False)
    assert tuple(out.shape) == (B, H, W, 4) and float(out[..., 3].abs().max()) == 0.0
    assert _rel(out[..., :3].cpu(), ref[..., :3].cpu()) < 1e-3         # approx-intrinsic SiLU forms differ in the last bf16 bit of a few operands
    assert _rel(out[..., :3].cpu(), tref.cpu()) < 1e-3
    u8 = ops.gn_silu_conv3x3_tail(x, st, gamma, beta, 32, 1e-6, wt, b, 3, True)
    ref8 = ops.image_to_u8(ref)
    assert u8.dtype == torch.uint8 and tuple(u8.shape) == (B, H, W, 3)
    d = (u8.int() - ref8.int()).abs()
    assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 1e-2     # a sum landing on a truncation boundary


@pytest.mark.parametrize("variant", ["silu", "f32_resid_dual", "leaky_f32", "gelu_alpha"])
def test_conv3x3_exchanged_roles_epilogues(cuda, variant):
    """The exchanged-operand form (weights as the M operand, 256 pixels as N; N <= 128 convs with many tiles)
    through its activation / residual / output variants, against torch on the same bf16 operands."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(95)
    B, H, W, Cin, Cout = 2, 288, 288, 64, 128
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        x = _bf(torch.randn(B, Cin, H, W, generator=g)).to(cuda)
        w = _bf(torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)).to(cuda)
        b = torch.randn(Cout, generator=g).to(cuda)
        conv = F.conv2d(x, w, b, padding=1).permute(0, 2, 3, 1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    a = x.permute(0, 2, 3, 1).contiguous().bfloat16()
    wp = ops.pack_conv_weight(w)
    if variant == "silu":
        out = ops.conv_gemm(a, wp, Cout, 9, bias=b, act=1).float()
        assert _rel(out.cpu(), F.silu(conv).cpu()) < 4e-3
    elif variant == "f32_resid_dual":
        r = torch.randn(B, H, W, Cout, generator=g).to(cuda)
        of, oh, st = ops.conv_gemm(a, wp, Cout, 9, bias=b, resid=r, alpha=0.5, dual=True, stats=True)
        ref = r + 0.5 * conv
        assert _rel(of.cpu(), ref.cpu()) < 1e-3 and torch.equal(oh.float(), of.bfloat16().float())
        blk = of.view(B * H * W // 32, 32, Cout)
        assert _rel(st[..., 0].cpu(), blk.sum(1).cpu()) < 1e-4
    elif variant == "leaky_f32":
        out = ops.conv_gemm(a, wp, Cout, 9, bias=b, act=3, act_param=0.1, out_f32=True)
        assert _rel(out.cpu(), F.leaky_relu(conv, 0.1).cpu()) < 1e-3
    else:
        # alpha != 1 without a residual is outside the exchanged form's fast path: the library must fall back
        out = ops.conv_gemm(a, wp, Cout, 9, bias=b, act=4, alpha=0.25, out_f32=True)
        assert _rel(out.cpu(), (0.25 * F.gelu(conv)).cpu()) < 1e-3


# ------------------------------------------------------------------------------------------
# round 2: upsample / stride-2 / zero-conv injection folded into the implicit GEMM
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 32, 32, 64, 96), (1, 64, 64, 128, 256), (2, 16, 16, 128, 64), (1, 8, 8, 72, 320),
                                            (1, 8, 40, 64, 32), (3, 12, 20, 16, 48)])
def test_conv3x3_after_upsample_folded(cuda, B, H, W, Cin, Cout):
    """nearest x2 + conv3x3 (openaimodel.py:106-113, model.py:63-67) as four 2x2 parity convs.  The folded
    weights are sums of bf16-exact taps rounded to bf16 once, so the tolerance is bf16 operand rounding."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(61)
    x = _bf(torch.randn(B, Cin, H, W, generator=g))
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1).permute(0, 2, 3, 1)
    wp = ops.pack_up2_weight(w.to(cuda))
    a = x.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16()
    of, oh, st = ops.conv_gemm(a, wp, Cout, 4, bias=b.to(cuda), dual=True, stats=True, up2=True, w_batch_stride=wp.stride(0))
    assert tuple(of.shape) == (B, 2 * H, 2 * W, Cout)
    assert _rel(of.cpu(), ref) < 4e-3, _rel(of.cpu(), ref)
    assert torch.equal(oh, of.bfloat16())
    assert (st is not None) == (W % 32 == 0 and Cout % 32 == 0 and ops.conv_stats_supported(B, H, W, Cout, 4 * ((Cin + 63) // 64)))
    if st is not None:
        # the slab table is ordered [sample][parity][h][w]; a consumer only needs per-sample sums
        gamma, beta = torch.randn(Cout, generator=g), torch.randn(Cout, generator=g)
        gn = ops.groupnorm(of, gamma.to(cuda), beta.to(cuda), 32, 1e-5, True, stats1=st).float().cpu()
        gref = F.silu(F.group_norm(of.cpu().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
        assert torch.allclose(gn, gref, atol=4e-2, rtol=2e-2)
        tot = st.view(B, -1, Cout, 2).sum(1).cpu()
        assert torch.allclose(tot[..., 0], of.cpu().sum((1, 2)), rtol=1e-3, atol=1e-2)
    # bf16-only output, no stats (the VAE form)
    o2 = ops.conv_gemm(a, wp, Cout, 4, bias=b.to(cuda), up2=True, w_batch_stride=wp.stride(0))
    assert torch.equal(o2, oh)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,pad_lo", [(2, 32, 32, 64, 96, 3, 1), (1, 64, 64, 320, 320, 3, 1), (2, 16, 16, 128, 64, 3, 0),
                                                     (1, 8, 8, 72, 320, 3, 1), (1, 12, 40, 64, 32, 3, 0), (2, 16, 24, 128, 64, 1, 1),
                                                     (8, 64, 64, 64, 64, 3, 1)])
def test_conv_stride2_strided_tensor_map(cuda, B, H, W, Cin, Cout, k, pad_lo):
    """Stride-2 conv (openaimodel.py:150-152; model.py:82-84 pads bottom/right only; res_blk.py:15-25 incl. the
    1x1 shortcut) read through a tensor map with element strides 2, against torch."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(62)
    x = _bf(torch.randn(B, Cin, H, W, generator=g))
    w = _bf(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(k * k * Cin))
    b = torch.randn(Cout, generator=g)
    if k == 1:
        ref = F.conv2d(x, w, b, stride=2)
    elif pad_lo:
        ref = F.conv2d(x, w, b, stride=2, padding=1)
    else:
        ref = F.conv2d(F.pad(x, (0, 1, 0, 1)), w, b, stride=2)
    ref = ref.permute(0, 2, 3, 1)
    wp = ops.pack_conv_weight(w.to(cuda))
    a = x.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16()
    of, oh, st = ops.conv_gemm(a, wp, Cout, k * k, bias=b.to(cuda), dual=True, stats=True, stride2=True, pad_lo=pad_lo)
    assert tuple(of.shape) == (B, H // 2, W // 2, Cout)
    assert _rel(of.cpu(), ref) < 1e-3, _rel(of.cpu(), ref)
    if st is not None:
        tot = st.view(B, -1, Cout, 2).sum(1).cpu()
        assert torch.allclose(tot[..., 0], of.cpu().sum((1, 2)), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("B,H,W,Cin,Cc,Cout,taps,stride2", [(2, 16, 16, 320, 64, 320, 9, False), (1, 32, 32, 64, 128, 96, 9, False),
                                                           (2, 8, 8, 1280, 256, 1280, 1, False), (1, 64, 64, 8, 64, 320, 9, False),
                                                           (2, 32, 32, 128, 64, 128, 9, True), (8, 8, 8, 1280, 256, 1280, 9, False)])
def test_conv_fused_zero_conv_injection(cuda, B, H, W, Cin, Cc, Cout, taps, stride2):
    """`h = conv(x) + resid; h = h + zero_conv(h_ctr) * scale` (rdeic.py:194,203,207) in one GEMM: the control
    tensor is a centre-tap-only second K segment with its own (pre-scaled) weights."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(63)
    k = 3 if taps == 9 else 1
    scale = 0.37
    x = _bf(torch.randn(B, Cin, H, W, generator=g))
    oh_, ow_ = (H // 2, W // 2) if stride2 else (H, W)
    hc = _bf(torch.randn(B, Cc, oh_, ow_, generator=g))
    w = _bf(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(k * k * Cin))
    wz = _bf(torch.randn(Cout, Cc, 1, 1, generator=g) / math.sqrt(Cc))
    b, bz = torch.randn(Cout, generator=g), torch.randn(Cout, generator=g)
    resid = torch.randn(B, oh_, ow_, Cout, generator=g)
    ref = F.conv2d(x, w, b, stride=2 if stride2 else 1, padding=k // 2) + F.conv2d(hc, wz, bz)
    ref = ref.permute(0, 2, 3, 1) + resid
    wp, w2 = ops.pack_conv_weight(w.to(cuda)), ops.pack_conv_weight(wz.to(cuda))
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16()
    of, oh, st = ops.conv_gemm(nhwc(x), wp, Cout, taps, a2=nhwc(hc), w2=w2, bias=(b + bz).to(cuda), resid=resid.to(cuda),
                               dual=True, stats=True, stride2=stride2)
    assert _rel(of.cpu(), ref) < 1e-3, _rel(of.cpu(), ref)
    # the scale is folded into w2 / the bias by the caller: check that route too (bf16 rounding of the scaled weights)
    w2s = ops.pack_conv_weight((wz * scale).to(cuda))
    of2 = ops.conv_gemm(nhwc(x), wp, Cout, taps, a2=nhwc(hc), w2=w2s, bias=(b + scale * bz).to(cuda), resid=resid.to(cuda),
                        out_f32=True, stride2=stride2)
    ref2 = (F.conv2d(x, w, b, stride=2 if stride2 else 1, padding=k // 2) + scale * F.conv2d(hc, wz, bz)).permute(0, 2, 3, 1) + resid
    assert _rel(of2.cpu(), ref2) < 3e-3, _rel(of2.cpu(), ref2)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride2", [(2, 16, 16, 128, 128, 3, False), (1, 32, 32, 3, 128, 3, False),
                                                      (1, 32, 32, 128, 256, 1, False), (2, 32, 32, 128, 128, 3, True)])
def test_conv_three_pass_split_bf16(cuda, B, H, W, Cin, Cout, k, stride2):
    """fp32-equivalent products on the tensor cores: activations as [hi | lo | hi] (`ops.split3`) against weights packed as
    [w_hi | w_hi | w_lo] (`Conv.load(split3=True)`) in ONE tcgen05 GEMM over 3C channels -- the convs of the VAE encoder's
    precision="high" mode (ldm/modules/diffusionmodules/model.py:82-84,128-151 at fp32 fidelity).  The plain bf16 conv of
    the same fp32 data is two orders of magnitude further from the fp32 result."""
    from rdeic_b200 import ops
    from rdeic_b200.engine import Conv

    g = torch.Generator().manual_seed(77)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(k * k * Cin)
    b = torch.randn(Cout, generator=g)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        if stride2:          # model.py:82-84: pad (0,1,0,1), stride 2, no padding
            ref = F.conv2d(F.pad(x.to(cuda).double(), (0, 1, 0, 1)), w.to(cuda).double(), b.to(cuda).double(), stride=2)
        else:
            ref = F.conv2d(x.to(cuda).double(), w.to(cuda).double(), b.to(cuda).double(), padding=k // 2)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    ref = ref.permute(0, 2, 3, 1).float().cpu()
    sd = {"c.weight": w, "c.bias": b}
    a32 = x.to(cuda).permute(0, 2, 3, 1).contiguous()
    kw = dict(stride2=True, pad_lo=0) if stride2 else {}
    c3 = Conv.load(sd, "c", cuda, split3=True)
    out3 = ops.conv_gemm(ops.split3(a32), c3.w, c3.n_out, k * k, bias=c3.b, out_f32=True, **kw).cpu()
    if Cin % 8 == 0:
        c1 = Conv.load(sd, "c", cuda)
        out1 = ops.conv_gemm(ops.f32_to_bf16(a32), c1.w, c1.n_out, k * k, bias=c1.b, out_f32=True, **kw).cpu()
        assert _rel(out1, ref) > 20 * _rel(out3, ref)
    assert _rel(out3, ref) < 3e-5, _rel(out3, ref)


def test_split_hilo(cuda):
    """hi = bf16(x), lo = bf16(x - hi): hi + lo reproduces x to ~2^-17 relative, hi is torch's own rounding."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(71)
    x = torch.randn(37, 5, 12, generator=g) * 3
    out = ops.split_hilo(x.to(cuda)).cpu()
    assert tuple(out.shape) == (37, 5, 24)
    hi, lo = out[..., :12].float(), out[..., 12:].float()
    assert torch.equal(hi, x.bfloat16().float())
    assert torch.equal(lo, (x - hi).bfloat16().float())
    assert ((hi + lo - x).abs() <= x.abs() * 2.0 ** -16 + 1e-30).all()
    # windows inside a wider, pre-zeroed row
    dst = torch.zeros(37 * 5, 40, dtype=torch.bfloat16, device=cuda)
    ops.split_hilo(x.view(-1, 12).to(cuda), dst, off_hi=16, off_lo=4)
    assert torch.equal(dst[:, 16:28].float().cpu(), hi.view(-1, 12)) and torch.equal(dst[:, 4:16].float().cpu(), lo.view(-1, 12))
    assert float(dst[:, :4].abs().sum()) == 0 and float(dst[:, 28:].abs().sum()) == 0


@pytest.mark.parametrize("h,w,tile,ov,s", [(40, 56, 24, 8, 2), (24, 24, 24, 4, 8), (64, 96, 32, 16, 4)])
def test_blend_tiles_u8(cuda, h, w, tile, ov, s):
    """One-kernel uint8 tile blend against the float blend of rdeic_b200.parallel (same ramp), +-1 LSB for the
    uint8 rounding; tiles cut from one image blend back to exactly that image."""
    from rdeic_b200 import ops, parallel

    g = torch.Generator().manual_seed(72)
    plan = parallel.plan_tiles(h, w, tile, ov)
    th, tw = plan[0][2], plan[0][3]
    tiles = torch.randint(0, 256, (len(plan), th * s, tw * s, 3), generator=g, dtype=torch.uint8)
    origins = torch.tensor([[p[0] * s, p[1] * s] for p in plan], dtype=torch.int32)
    out = ops.blend_tiles_u8(tiles.to(cuda), origins.to(cuda), ov * s, h * s, w * s).cpu()
    ref = parallel.blend_tiles([t.permute(2, 0, 1).float() for t in tiles], plan, h, w, ov, s).permute(1, 2, 0)
    assert (out.float() - ref).abs().max() <= 0.5 + 1e-3
    full = torch.randint(0, 256, (h * s, w * s, 3), generator=g, dtype=torch.uint8)
    cut = torch.stack([full[p[0] * s:(p[0] + th) * s, p[1] * s:(p[1] + tw) * s] for p in plan])
    assert torch.equal(ops.blend_tiles_u8(cut.to(cuda), origins.to(cuda), ov * s, h * s, w * s).cpu(), full)


@pytest.mark.parametrize("B,H,W,Cin,Cc,Cout", [(2, 32, 32, 64, 64, 96), (1, 16, 16, 128, 256, 320), (8, 8, 8, 1280, 256, 1280)])
def test_conv_upsample_folded_with_injection(cuda, B, H, W, Cin, Cc, Cout):
    """The decoder's `h = Upsample.conv(nearest2x(h)); h = h + zero_conv(h_ctr) * scale` (openaimodel.py:106-113 then
    rdeic.py:207 before the next block) in one GEMM: the control tensor lives on the 2H x 2W output grid and every
    parity class of the folded conv reads every other pixel of it."""
    from rdeic_b200 import ops

    g = torch.Generator().manual_seed(64)
    x = _bf(torch.randn(B, Cin, H, W, generator=g))
    hc = _bf(torch.randn(B, Cc, 2 * H, 2 * W, generator=g))
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)
    wz = _bf(torch.randn(Cout, Cc, 1, 1, generator=g) / math.sqrt(Cc))
    b, bz = torch.randn(Cout, generator=g), torch.randn(Cout, generator=g)
    ref = (F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1) + F.conv2d(hc, wz, bz)).permute(0, 2, 3, 1)
    wp, w2 = ops.pack_up2_weight(w.to(cuda)), ops.pack_conv_weight(wz.to(cuda))
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16()
    of, oh, st = ops.conv_gemm(nhwc(x), wp, Cout, 4, a2=nhwc(hc), w2=w2, bias=(b + bz).to(cuda), dual=True, stats=True, up2=True,
                               w_batch_stride=wp.stride(0))
    assert tuple(of.shape) == (B, 2 * H, 2 * W, Cout)
    assert _rel(of.cpu(), ref) < 4e-3, _rel(of.cpu(), ref)
