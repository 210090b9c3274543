"""-m gpu: exhaustive proofs that the strict kernel's shortcuts are bit-identical to IEEE division /
the full-range shared exp (every float32 in the relevant ranges, on the device)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", [0, 1, 2])
def test_exhaustive_selftests(pkg, which):
    lib = pkg._lib.load()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    pkg._lib.check(lib.cstr_selftest(which, bad.data_ptr(), None), "cstr_selftest")
    torch.cuda.synchronize()
    assert int(bad.item()) == 0
