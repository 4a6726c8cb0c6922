"""the oracle's calculate_qparams and the product's host mirror against the reference's outputs"""
import pytest
import torch

from compressed_tensors_b200.quantization import QuantizationArgs, calculate_qparams
from oracle.qparams import calculate_qparams as orc_qparams
from tests.golden import load
from tests.util import bits_equal, diff_report

_G = load("qparams")


@pytest.mark.parametrize("i", range(len(_G)))
def test_qparams_golden(i):
    c = _G[i]
    a = c["args"]
    s, z = orc_qparams(c["min"], c["max"], num_bits=a["num_bits"], qtype=a["type"], symmetric=a["symmetric"])
    assert bits_equal(s, c["scale"]), "oracle scale: " + diff_report(s, c["scale"])
    assert z.dtype == c["zp"].dtype
    zc, zw = (z.view(torch.uint8), c["zp"].view(torch.uint8)) if z.dtype == torch.float8_e4m3fn else (z, c["zp"])
    assert torch.equal(zc, zw), "oracle zero point: " + diff_report(zc, zw)
    # the product's host-side mirror (torch ops on qparam-sized tensors) must agree as well
    args = QuantizationArgs(num_bits=a["num_bits"], type=a["type"], symmetric=a["symmetric"], strategy="channel")
    s2, z2 = calculate_qparams(c["min"], c["max"], args)
    assert bits_equal(s2, c["scale"]), "mirror scale: " + diff_report(s2, c["scale"])
    z2c = z2.view(torch.uint8) if z2.dtype == torch.float8_e4m3fn else z2
    assert z2.dtype == c["zp"].dtype and torch.equal(z2c, zw)
