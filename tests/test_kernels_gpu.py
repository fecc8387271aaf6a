"""GPU parity tests (pytest -m gpu): every kernel through the C ABI vs a PyTorch fp32 reference
of the same op on the same seeded inputs.  See tests/kernel_checks.py for the cases."""
import pytest

import kernel_checks as kc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [n for n in kc.CHECKS if n != "diag_gemm"])
def test_kernel(name):
    fn, tol = kc.CHECKS[name]
    err = fn()
    assert err == err and err <= tol, f"{name}: rel err {err:.3e} > {tol}"
