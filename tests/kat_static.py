"""
Known answers of the reference's own lifecycle tests for fake_quantize, strategy by strategy:
/root/reference/tests/test_quantization/lifecycle/test_static_lifecycle.py:18-129 (weights), :166-216 (activations), :272-336
(attention heads).  Only the inputs (an arange) and the expected values are restated here; `run` re-creates the flow of those tests
(memoryless min/max observer -> calculate_qparams -> fake_quantize) around a caller-supplied fake_quantize, so the same table pins
the CPU oracle (tests/test_oracle_golden.py) and the CUDA kernels (tests/test_gpu_reference_suite.py).
The reference compares with torch.allclose at default tolerances after casting the 4-decimal literals to bf16, i.e. exactly.
"""
import torch

from compressed_tensors_b200.quantization import QuantizationArgs
from compressed_tensors_b200.quantization.utils.helpers import calculate_qparams, compute_dynamic_scales_and_zp, generate_gparam

FP8 = torch.float8_e4m3fn
W = (4, 6)        # weight  = arange(24).reshape(4, 6)
A = (1, 2, 6)     # input   = arange(12).reshape(1, 2, 6)
K = (1, 2, 3, 4)  # k state = arange(24).reshape(batch 1, heads 2, seq 3, head_dim 4)

CASES = [
    ("weight-tensor", dict(num_bits=4, type="int", symmetric=True, strategy="tensor"), W,
     [[0.0000, 0.0000, 3.0625, 3.0625, 3.0625, 6.1250], [6.1250, 6.1250, 9.1875, 9.1875, 9.1875, 12.2500],
      [12.2500, 12.2500, 15.3125, 15.3125, 15.3125, 18.3750], [18.3750, 18.3750, 21.5000, 21.5000, 21.5000, 21.5000]]),
    ("weight-channel", dict(num_bits=4, type="int", symmetric=True, strategy="channel"), W,
     [[0.0000, 1.3359, 2.0000, 2.6719, 4.0000, 4.6875], [5.8750, 7.3438, 7.3438, 8.8125, 10.2500, 10.2500],
      [11.3125, 13.6250, 13.6250, 15.8750, 15.8750, 15.8750], [18.3750, 18.3750, 21.5000, 21.5000, 21.5000, 21.5000]]),
    ("weight-group3", dict(num_bits=4, type="int", symmetric=True, strategy="group", group_size=3), W,
     [[0.0000, 1.0703, 1.8750, 2.6719, 4.0000, 4.6875], [6.4375, 7.5000, 7.5000, 8.8125, 10.2500, 10.2500],
      [11.1875, 13.0625, 13.0625, 15.8750, 15.8750, 15.8750], [18.7500, 18.7500, 18.7500, 21.5000, 21.5000, 21.5000]]),
    ("weight-tensor_group3-fp4", dict(num_bits=4, type="float", symmetric=True, strategy="tensor_group", group_size=3, scale_dtype=FP8, zp_dtype=FP8), W,
     [[0.0000, 1.0234, 2.0469, 3.2812, 3.2812, 4.9375], [5.4688, 8.1875, 8.1875, 10.6875, 10.6875, 10.6875],
      [9.8750, 14.7500, 14.7500, 16.3750, 16.3750, 16.3750], [19.7500, 19.7500, 19.7500, 23.0000, 23.0000, 23.0000]]),
    ("weight-block2x3", dict(num_bits=4, type="int", symmetric=True, strategy="block", block_structure=[2, 3]), W,
     [[0.0000, 1.0703, 2.1406, 2.9375, 4.4062, 4.4062], [6.4375, 7.5000, 7.5000, 8.8125, 10.2500, 10.2500],
      [10.6875, 13.3750, 13.3750, 15.3125, 15.3125, 18.3750], [18.7500, 18.7500, 18.7500, 21.5000, 21.5000, 21.5000]]),
    ("input-tensor", dict(num_bits=4, type="int", symmetric=True, strategy="tensor"), A,
     [[[0.0000, 1.4688, 1.4688, 2.9375, 4.4062, 4.4062], [5.8750, 7.3438, 7.3438, 8.8125, 10.2500, 10.2500]]]),
    ("input-tensor_group3-fp4-local", dict(num_bits=4, type="float", symmetric=True, strategy="tensor_group", dynamic="local", group_size=3, scale_dtype=FP8, zp_dtype=FP8), A,
     [[[0.0000, 0.9844, 1.9688, 3.4062, 3.4062, 5.1250], [5.2500, 7.8750, 7.8750, 7.3438, 11.0000, 11.0000]]]),
    ("k-tensor", dict(num_bits=4, type="int", symmetric=True, strategy="tensor"), K,
     [[[[0.0000, 0.0000, 3.0625, 3.0625], [3.0625, 6.1250, 6.1250, 6.1250], [9.1875, 9.1875, 9.1875, 12.2500]],
       [[12.2500, 12.2500, 15.3125, 15.3125], [15.3125, 18.3750, 18.3750, 18.3750], [21.5000, 21.5000, 21.5000, 21.5000]]]]),
    ("k-attn_head", dict(num_bits=4, type="int", symmetric=True, strategy="attn_head"), K,
     [[[[0.0000, 1.4688, 1.4688, 2.9375], [4.4062, 4.4062, 5.8750, 7.3438], [7.3438, 8.8125, 10.2500, 10.2500]],
       [[12.2500, 12.2500, 15.3125, 15.3125], [15.3125, 18.3750, 18.3750, 18.3750], [21.5000, 21.5000, 21.5000, 21.5000]]]]),
]


def _observe(x: torch.Tensor, args: QuantizationArgs):
    """min / max over what one scale covers (the reduction of the reference's memoryless observer for each strategy)"""
    s = args.strategy
    if s == "tensor":
        return x.amin().reshape(1), x.amax().reshape(1)
    if s == "channel":
        return x.amin(-1, keepdim=True), x.amax(-1, keepdim=True)
    if s in ("group", "tensor_group"):
        g = x.unflatten(-1, (x.shape[-1] // args.group_size, args.group_size))
        return g.amin(-1), g.amax(-1)
    if s == "block":
        bh, bw = args.block_structure
        b = x.reshape(x.shape[0] // bh, bh, x.shape[1] // bw, bw)
        return b.amin((1, 3)), b.amax((1, 3))
    if s == "attn_head":  # one scale per head of [batch, heads, seq, head_dim]
        return x.amin((0, 2, 3)).reshape(-1, 1, 1), x.amax((0, 2, 3)).reshape(-1, 1, 1)
    raise ValueError(s)


def run(case, fake_quantize, device="cpu"):
    """-> (output, expected); fake_quantize(x, scale, zero_point, args, global_scale) is the implementation under test"""
    name, kw, shape, want = case
    args = QuantizationArgs(**kw)
    n = 1
    for d in shape:
        n *= d
    x = torch.arange(n, dtype=torch.bfloat16, device=device).reshape(shape)
    gscale = generate_gparam(x.amin(), x.amax()) if args.strategy == "tensor_group" else None
    if args.dynamic in (True, "local"):
        scale, zp = compute_dynamic_scales_and_zp(x, args, global_scale=gscale)
    else:
        lo, hi = _observe(x, args)
        scale, zp = calculate_qparams(lo, hi, args, global_scale=gscale)
    return fake_quantize(x, scale, zp, args, gscale), torch.tensor(want).to(torch.bfloat16)
