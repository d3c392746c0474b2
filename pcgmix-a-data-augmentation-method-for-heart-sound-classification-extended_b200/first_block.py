"""Forward pass of the first block of the reference's ResNet9-1D on the device the cycles are on.

The consumer of ``augment``'s output in the reference's training loop is ``model(data)`` (train_model.py:536); for
the 1D models that is ``ResNet9_myrtle.forward`` (models.py:534-552), whose first layer ``conv1`` is
``conv_block(in_channels, filters[0])`` = ``nn.Sequential(nn.Conv1d(C, F, kernel_size=3, padding=1),
nn.BatchNorm1d(F), nn.ReLU(inplace=True))`` (models.py:468-473, :523).  SURVEY section 8(f)4 lists it as the second
downstream consumer of the augmented cycles.  :func:`first_conv_block` takes that very module (so its parameters,
its ``training`` flag and its running statistics are the module's own) and returns what ``block(data)`` returns,
computed by ``pcgmix_first_conv_block``: in training mode the batch statistics come from second moments of the input
patches, so the 16x larger output is written once instead of written, read for the statistics, and rewritten.

Forward only: the result carries no autograd graph (inference, feature extraction, or a frozen first block).  There
is no CPU path.
"""
from __future__ import annotations

import torch
from torch import nn

from . import native

__all__ = ["first_conv_block", "check_block"]


def check_block(block) -> tuple:
    """``(conv, bn)`` of a reference ``conv_block`` without pooling, or ``ValueError`` naming what differs."""
    layers = list(block) if isinstance(block, nn.Sequential) else None
    if not layers or len(layers) != 3:
        raise ValueError("first_conv_block expects the reference's conv_block: Sequential(Conv1d, BatchNorm1d, ReLU)")
    conv, bn, act = layers
    if not isinstance(conv, nn.Conv1d) or not isinstance(bn, nn.BatchNorm1d) or not isinstance(act, nn.ReLU):
        raise ValueError("first_conv_block expects Sequential(Conv1d, BatchNorm1d, ReLU)")
    if (conv.kernel_size != (3,) or conv.stride != (1,) or conv.padding != (1,) or conv.dilation != (1,)
            or conv.groups != 1 or conv.padding_mode != "zeros"):
        raise ValueError("first_conv_block: the convolution must be kernel_size=3, stride=1, padding=1 (zeros), groups=1")
    if bn.num_features != conv.out_channels:
        raise ValueError("first_conv_block: BatchNorm1d width differs from the convolution's")
    return conv, bn


def first_conv_block(block: nn.Sequential, data: torch.Tensor, out: torch.Tensor = None,
                     return_statistics: bool = False):
    """``block(data)`` for ``block = model.conv1`` of the reference's ResNet9-1D and ``data`` (B, C, L) float32 on a
    CUDA device.  Honours ``block.training`` like the modules do: in training mode (or without running statistics)
    the batch statistics normalise and ``running_mean`` / ``running_var`` / ``num_batches_tracked`` are updated in
    place; in evaluation mode the running statistics normalise.  ``return_statistics=True`` also returns the mean
    and inverse standard deviation used (what ``torch.batch_norm`` saves for its backward)."""
    conv, bn = check_block(block)
    if not isinstance(data, torch.Tensor) or not data.is_cuda:
        raise RuntimeError("first_conv_block: data must be a CUDA tensor (there is no CPU fallback)")
    if data.dim() != 3 or data.dtype != torch.float32 or data.shape[1] != conv.in_channels:
        raise ValueError(f"first_conv_block: data must be (B, {conv.in_channels}, L) float32")
    data = data.contiguous()
    B, C, L = data.shape
    F = conv.out_channels
    batch_stats = bn.training or bn.running_mean is None
    if batch_stats and B * L < 2:
        raise ValueError("Expected more than 1 value per channel when training")      # torch's own check
    momentum = 0.0
    track = bn.training and bn.track_running_stats and bn.running_mean is not None
    if track:
        bn.num_batches_tracked += 1                       # device-side increment, like the module's forward
        # momentum=None means a cumulative average; that factor needs the counter's value on the host
        momentum = float(bn.momentum) if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked.item())
    with torch.no_grad():
        params = [t.detach().contiguous() if t is not None else None
                  for t in (conv.weight, conv.bias, bn.weight, bn.bias)]
        if out is None:
            out = torch.empty((B, F, L), dtype=torch.float32, device=data.device)
        workspace = torch.empty(native.first_conv_block_workspace(C, F), dtype=torch.uint8, device=data.device)
        stats = (torch.empty(F, dtype=torch.float32, device=data.device),
                 torch.empty(F, dtype=torch.float32, device=data.device)) if return_statistics else (None, None)
        native.first_conv_block(data, params[0], params[1], params[2], params[3],
                                bn.running_mean if (track or not batch_stats) else None,
                                bn.running_var if (track or not batch_stats) else None,
                                out, workspace, batch_stats, float(bn.eps), momentum, stats[0], stats[1])
    return (out, stats[0], stats[1]) if return_statistics else out
