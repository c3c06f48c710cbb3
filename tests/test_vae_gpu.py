"""GPU parity of the VAE path (lunaris_orion_b200.lunar_generate) vs the CPU oracle, calibrated against a bf16-autocast
execution of the same op sequence (see test_teacher_gpu.py for the tolerance rationale)."""
import pytest

import vae_cases as vc


@pytest.mark.gpu
def test_vae_forward_backward_and_sampling(cuda_dev):
    rep = vc.vae_report(cuda_dev)
    assert rep["all_grads_present"]                                   # 72/72 tensors, as in the reference
    assert rep["recon"] <= 3 * rep["cal_recon"] + 5e-3
    assert rep["mu"] <= 3 * rep["cal_mu"] + 5e-3
    assert rep["logvar"] <= 3 * rep["cal_mu"] + 5e-3
    assert rep["grad_rel_max"] <= 3 * rep["cal_grad_rel_max"] + 0.01, rep["grad_worst"]
    assert rep["ratio_worst"][0][1] < 4.0, rep["ratio_worst"]
    assert rep["sample"] < 0.05
