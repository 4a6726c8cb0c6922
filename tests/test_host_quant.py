"""host-side qparam rules that need no GPU and no oracle (reference: quantization/utils/helpers.py)"""
import pytest
import torch


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32, torch.float64])
@pytest.mark.parametrize("shape", [(), (7, 1), (3, 5)])
@pytest.mark.parametrize("symmetric", [True, False])
def test_calculate_qparams_scalar_divisor_semantics(dtype, shape, symmetric):
    """the tensor divisor used for CUDA exactness keeps the Python-scalar dtype rules of helpers.py:50-137 (0-dim min/max of the
    TENSOR strategy included) and, on CPU, the same values"""
    from compressed_tensors_b200.quantization.quant_args import QuantizationArgs
    from compressed_tensors_b200.quantization.utils.helpers import calculate_qparams, calculate_range
    torch.manual_seed(5)
    lo = -torch.rand(shape, dtype=torch.float32).to(dtype)
    hi = torch.rand(shape, dtype=torch.float32).to(dtype)
    args = QuantizationArgs(num_bits=8, symmetric=symmetric)
    scale, zp = calculate_qparams(lo, hi, args)
    assert scale.dtype == dtype and scale.shape == (lo.shape or (1,))  # 0-dim qparams become shape (1,) (helpers.py:133-135)
    bmin, bmax = calculate_range(args, lo.device)
    rng = float(bmax - bmin)
    want = torch.max(lo.abs(), hi.abs()) / (rng / 2) if symmetric else (hi - lo) / rng
    eps = torch.finfo(dtype).eps
    want = torch.where(want == 0, torch.tensor(eps, dtype=dtype), want)
    assert torch.equal(scale, want.reshape(scale.shape))
