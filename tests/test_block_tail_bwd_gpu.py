"""Backward of the ExpertBlock tail (leaky_relu(residual) -> layer_scale -> Dropout2d -> BatchNorm -> LeakyReLU;
reference lunar_evaluator.py:241-258, 260-275 under autograd) through the two C ABI entry points, EVERY call variant:
stored upstream gradient / broadcast pooled gradient, with and without the block output, dpre, Dropout2d mask, layer
scale and bias sums. The common variants run the register-resident `_fast` kernels, the rest the generic ones
(csrc/teacher_elem.cu); both are held to a plain torch fp32 restatement of the same arithmetic."""
import ctypes

import pytest
import torch


def _bf(t):
    return t.to(torch.bfloat16).float()


def _ref_reduce(dout, gpool, out, a, mean, rstd, m2, slope_out, want_dpre):
    B, HW, C = a.shape
    d = dout.float() if dout is not None else gpool[:, None, :].expand(B, HW, C)
    if out is not None:
        d = torch.where(out.float() > 0, d, d * slope_out)
    dpre = None
    if want_dpre:
        dpre = d.to(torch.bfloat16)
        d = dpre.float()
    g = d * (m2[:, None, :] if m2 is not None else 1.0)
    t1 = g.double().sum((0, 1))
    t2 = (g * (a.float() - mean) * rstd).double().sum((0, 1))
    return dpre, t1.float(), t2.float()


def _ref_apply(dpre, gpool, out, a, mean, rstd, gamma, ls, m2, t1, t2, slope_out, slope_a):
    B, HW, C = a.shape
    if dpre is not None:
        d = dpre.float()
    else:
        d = gpool[:, None, :].expand(B, HW, C)
        if out is not None:
            d = torch.where(out.float() > 0, d, d * slope_out)
    lsv = ls if ls is not None else torch.ones_like(mean)
    inv_n = 1.0 / (B * HW)
    k0 = gamma * rstd
    s1, s2 = lsv * t1 * inv_n, lsv * t2 * inv_n
    p1 = (k0 * lsv)[None, None, :] * (m2[:, None, :] if m2 is not None else 1.0)
    z = d * p1 + a.float() * (-k0 * s2 * rstd) + k0 * (s2 * rstd * mean - s1)
    z = torch.where(a.float() <= 0, z * slope_a, z)
    return z, _bf(z).double().sum((0, 1)).float()


@pytest.mark.gpu
@pytest.mark.parametrize("C,B,HW", [(64, 3, 1000), (192, 3, 1000), (512, 4, 16384)])   # ragged small cases + the C3 shape
@pytest.mark.parametrize("variant", ["dout", "gpool", "dout_no_out", "gpool_no_dpre", "dout_no_mask_no_ls"])
def test_block_tail_backward_variants_match_fp32_restatement(cuda_dev, C, B, HW, variant):
    from lunaris_orion_b200 import _capi
    lib = _capi.lib()
    dev = cuda_dev
    torch.manual_seed(7)
    a = torch.randn(B, HW, C, device=dev).to(torch.bfloat16)
    out = torch.randn(B, HW, C, device=dev).to(torch.bfloat16)
    dout = (torch.randn(B, HW, C, device=dev) * 0.1).to(torch.bfloat16)
    gpool = torch.randn(B, C, device=dev) * 0.1
    mean, rstd = torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev) + 0.5
    gamma, ls = torch.rand(C, device=dev) + 0.5, torch.rand(C, device=dev) * 0.2
    m2 = (torch.rand(B, C, device=dev) > 0.1).float() * 1.109375
    use_dout = variant.startswith("dout")
    use_out = variant not in ("dout_no_out",)
    want_dpre = variant != "gpool_no_dpre"
    if variant == "dout_no_mask_no_ls":
        m2 = ls = None
    slope_out = 0.2 if use_out else 1.0
    p = lambda t: None if t is None else t.data_ptr()
    s = torch.cuda.current_stream().cuda_stream
    t = torch.zeros(2, C, device=dev)
    dpre = torch.empty_like(a) if want_dpre else None
    _capi.check(lib.lun_block_bwd_reduce_bf16(p(dout if use_dout else None), p(None if use_dout else gpool),
                                              p(out if use_out else None), a.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                              p(m2), p(dpre), t[0].data_ptr(), t[1].data_ptr(), B, HW, C,
                                              ctypes.c_float(slope_out), s), "lun_block_bwd_reduce_bf16")
    r_dpre, r_t1, r_t2 = _ref_reduce(dout if use_dout else None, gpool, out if use_out else None, a, mean, rstd, m2,
                                     slope_out, want_dpre)
    if want_dpre:
        assert torch.equal(dpre, r_dpre)              # one select and one rounding: bit-exact
    scale1, scale2 = r_t1.abs().max().item(), r_t2.abs().max().item()
    assert (t[0] - r_t1).abs().max().item() <= 2e-4 * scale1 + 1e-4
    assert (t[1] - r_t2).abs().max().item() <= 2e-4 * scale2 + 1e-4

    dz = torch.empty_like(a)
    dbias = torch.zeros(C, device=dev)
    _capi.check(lib.lun_block_bwd_apply_bf16(p(dpre), p(None if dpre is not None else gpool),
                                             p(None if dpre is not None else (out if use_out else None)), a.data_ptr(),
                                             mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), p(ls), p(m2),
                                             t[0].data_ptr(), t[1].data_ptr(), dz.data_ptr(), dbias.data_ptr(), B, HW, C,
                                             ctypes.c_float(slope_out), ctypes.c_float(0.2), s), "lun_block_bwd_apply_bf16")
    z, r_db = _ref_apply(dpre, gpool, out if use_out else None, a, mean, rstd, gamma, ls, m2, t[0], t[1], slope_out, 0.2)
    err = (dz.float() - z).abs()
    assert bool((err <= 2.0 ** -7 * z.abs() + 1e-5).all()), err.max().item()      # bf16 rounding of an fp32 result
    assert (dbias - r_db).abs().max().item() <= 2e-3 * r_db.abs().max().item() + 1e-3
