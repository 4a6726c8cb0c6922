"""shared helpers for the parity tests"""
import torch


def okw(args: dict) -> dict:
    """reference QuantizationArgs.model_dump() -> oracle keyword arguments"""
    return dict(
        strategy=args["strategy"],
        group_size=args.get("group_size"),
        block_structure=args.get("block_structure"),
        num_bits=args["num_bits"],
        qtype=args["type"],
    )


def bits_equal(a: torch.Tensor, b: torch.Tensor) -> bool:
    """bit-for-bit equality (distinguishes -0.0 from 0.0, compares NaN payloads)"""
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    a = a.contiguous()
    b = b.contiguous()
    if a.dtype.is_floating_point:
        width = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[a.element_size()]
        return torch.equal(a.view(width), b.view(width))
    return torch.equal(a, b)


def diff_report(a: torch.Tensor, b: torch.Tensor) -> str:
    if a.dtype != b.dtype or a.shape != b.shape:
        return f"dtype/shape {a.dtype}{tuple(a.shape)} vs {b.dtype}{tuple(b.shape)}"
    af, bf = a.float().flatten(), b.float().flatten()
    bad = (af != bf) & ~(af.isnan() & bf.isnan())
    n = int(bad.sum())
    idx = bad.nonzero().flatten()[:5].tolist()
    return f"{n}/{af.numel()} differ; first {[(i, af[i].item(), bf[i].item()) for i in idx]}"


def same(a: torch.Tensor, b: torch.Tensor, what: str = "") -> None:
    """bit-exact comparison with a compact failure message (no tensor dumps)"""
    import pytest

    if not bits_equal(a, b):
        pytest.fail(f"{what}: {diff_report(a, b)}", pytrace=False)


def same_values(a: torch.Tensor, b: torch.Tensor, what: str = "") -> None:
    """torch.equal semantics (the reference's own assertion: -0.0 == +0.0)"""
    import pytest

    if a.dtype != b.dtype or a.shape != b.shape or not torch.equal(a, b):
        pytest.fail(f"{what}: {diff_report(a, b)}", pytrace=False)


def same_nan(a: torch.Tensor, b: torch.Tensor, what: str = "") -> None:
    """bit-exact (sign of zero included) except that any NaN matches any NaN (payloads are not part of the contract)"""
    import pytest

    if a.dtype != b.dtype or a.shape != b.shape:
        pytest.fail(f"{what}: {diff_report(a, b)}", pytrace=False)
    na, nb = a.isnan(), b.isnan()
    if not torch.equal(na, nb):
        pytest.fail(f"{what}: NaN positions differ ({int(na.sum())} vs {int(nb.sum())})", pytrace=False)
    same(torch.where(na, torch.zeros_like(a), a), torch.where(nb, torch.zeros_like(b), b), what)


def free_port() -> str:
    """a TCP port nobody listens on right now, for torchrun's rendezvous on 127.0.0.1 (fixed ports collide when two suites run at once)"""
    import socket

    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sock:
        sock.bind(("127.0.0.1", 0))
        return str(sock.getsockname()[1])
