"""GPU parity: tcgen05 implicit-GEMM conv / convT / linear (fprop, dgrad, wgrad) through the C ABI vs torch fp32."""
import pytest

from conv_cases import CASES


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_conv_case(cuda_dev, name):
    fn, kw = CASES[name]
    err, tol = fn(cuda_dev, **kw)
    assert err <= tol, f"{name}: max abs err {err:.4g} > tol {tol:.4g}"
