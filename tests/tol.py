"""Tolerances of the decode parity tests (north_star: 1e-5 relative in fp32, 1e-3 in fp16).

Scores use the plain relative bound.  Box outputs come out of ``x2 - x1`` / ``(x1 + x2) / 2``
where the corners ``anchor -/+ distance`` were already rounded in the tensor dtype, so two
correct implementations whose expectation differs in the last ulp can differ by one ulp
*of the corner* -- an absolute amount, however small the box is.  The absolute term below
is therefore a few ulps of the largest corner coordinate of the level (in pixels); the
relative term is the north_star's.
"""
import math

import torch


def _ulp(x, mant_bits):
    return 2.0 ** (math.floor(math.log2(x)) - mant_bits)


def box_atol_per_level(level_hw, strides, dtype):
    out = []
    for (h, w), s in zip(level_hw, strides):
        gmax = max(h, w) + 16.0  # anchor + largest DFL distance, grid units
        if dtype == torch.float16:
            out.append(s * 1.0 * _ulp(gmax, 10))
        else:
            out.append(s * 4.0 * _ulp(gmax, 23))
    return out


def check_decode(y, ref, level_hw, strides, nc):
    """Returns (ok, message).  y/ref: [B, 4+nc, A] same dtype."""
    dtype = ref.dtype
    rtol = 1e-3 if dtype == torch.float16 else 1e-5
    yf, rf = y.float().cpu(), ref.float().cpu()
    if not torch.isfinite(yf).all():
        return False, "non-finite output"
    atols = box_atol_per_level(level_hw, strides, dtype)
    off = 0
    for (h, w), atol in zip(level_hw, atols):
        n = h * w
        d = (yf[:, :4, off : off + n] - rf[:, :4, off : off + n]).abs()
        bound = rtol * rf[:, :4, off : off + n].abs() + atol
        if (d > bound).any():
            i = (d - bound).argmax()
            return False, f"box mismatch level@{off}: worst excess {(d - bound).max().item():.3e} (flat {i.item()})"
        off += n
    d = (yf[:, 4:] - rf[:, 4:]).abs()
    bound = rtol * rf[:, 4:].abs() + (6e-8 if dtype == torch.float16 else 1e-12)
    if (d > bound).any():
        return False, f"score mismatch: worst excess {(d - bound).max().item():.3e}"
    if dtype == torch.float16:
        # bit-identical almost everywhere; the rest are 1-ulp flips of an intermediate (the
        # reference's own CPU kernels flip the same way between their vector and tail paths)
        nbox = (y[:, :4].cpu() != ref[:, :4].cpu()).sum().item()
        ncls = (y[:, 4:].cpu() != ref[:, 4:].cpu()).sum().item()
        if nbox > 0.02 * ref[:, :4].numel() + 8 or ncls > 0.01 * ref[:, 4:].numel() + 8:
            return False, f"too many 1-ulp flips: box {nbox}/{ref[:, :4].numel()}, scores {ncls}/{ref[:, 4:].numel()}"
    return True, "ok"


def check_bbox_decode(out, ref, grad=False, grad_out=None, gmax=None):
    """Training-time decode (Loss.bbox_decode).  Forward: a corner is ``anchor -/+ distance`` (grid units), so like
    the inference boxes its absolute error is a few ulps of the OPERANDS (``gmax`` = largest anchor coordinate + 16),
    however small the corner itself is: the north_star's relative bound plus 1 ulp(gmax) in fp16 / 4 ulp(gmax) in fp32.
    Backward (``grad=True``): the gradient of a bin is ``p_j * (g*j - g*d)`` -- a difference of terms of size
    ``16 |g|`` -- so the absolute term scales with the incoming gradient of that side: ``rtol * 16 * |g|``."""
    dtype = ref.dtype
    rtol = 1e-3 if dtype == torch.float16 else 1e-5
    of, rf = out.float().cpu(), ref.float().cpu()
    if not torch.isfinite(of).all():
        return False, "non-finite output"
    if not grad:
        atol = 1.0 * _ulp(gmax, 10) if dtype == torch.float16 else 4.0 * _ulp(gmax, 23)
        bound = rtol * rf.abs() + atol
    else:
        g = grad_out.float().cpu().abs()  # [B, A, 4] -> per side, broadcast over its 16 bins
        bound = rtol * rf.abs() + rtol * 16.0 * g.repeat_interleave(16, dim=-1) + (1e-7 if dtype == torch.float16 else 1e-12)
    d = (of - rf).abs()
    if (d > bound).any():
        return False, f"mismatch: worst excess {(d - bound).max().item():.3e} at flat {(d - bound).argmax().item()}"
    if dtype == torch.float16:
        n = (out.cpu() != ref.cpu()).sum().item()
        if n > 0.02 * ref.numel() + 8:
            return False, f"too many 1-ulp flips: {n}/{ref.numel()}"
    return True, "ok"
