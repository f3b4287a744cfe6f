"""Helpers shared by the -m gpu parity tests."""
import torch


def no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def rel_err(x, ref):
    """max |x - ref| / max |ref|"""
    x, ref = x.float(), ref.float()
    return ((x - ref).abs().max() / (ref.abs().max() + 1e-20)).item()


def rel_rms(x, ref):
    x, ref = x.double(), ref.double()
    return ((x - ref).pow(2).mean().sqrt() / (ref.pow(2).mean().sqrt() + 1e-30)).item()


def psnr(x, ref, peak=None):
    x, ref = x.double(), ref.double()
    mse = (x - ref).pow(2).mean().item()
    peak = ref.abs().max().item() if peak is None else peak
    return float("inf") if mse == 0 else 10 * torch.log10(torch.tensor(peak * peak / mse)).item()


def to_dev(sd, dev="cuda"):
    return {k: v.to(dev) for k, v in sd.items()}
