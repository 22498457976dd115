"""-m gpu: every libi2it kernel path, called through the C ABI, against plain torch fp32 on the same 16-bit-rounded inputs.
Tolerance = the output rounding of the dtype (fp16: 2^-11, bf16: 2^-8 relative) — stated in tests/gpu_diag.py::report."""
import pytest

pytestmark = pytest.mark.gpu


def _cases():
    import gpu_diag
    return list(gpu_diag.CASES)


@pytest.mark.parametrize("name", _cases())
def test_kernel_case(name):
    import gpu_diag
    assert gpu_diag.CASES[name](), name
