"""Single-head self-attention over the points of a cloud (reference models/attention.py:10-48).
Eval mode: softmax((q / sqrt(d))^T k) v as one streaming kernel; the (N, N) map is never stored."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class SelfAttention(nn.Module):
    def __init__(self, in_channel, out_channel=None, attn_dropout=0.1):
        super().__init__()
        self.in_channel = in_channel
        self.out_channel = out_channel if out_channel is not None else in_channel
        self.temperature = self.out_channel ** 0.5
        self.q_map = nn.Conv1d(in_channel, self.out_channel, 1, bias=False)
        self.k_map = nn.Conv1d(in_channel, self.out_channel, 1, bias=False)
        self.v_map = nn.Conv1d(in_channel, self.out_channel, 1, bias=False)
        self.dropout = nn.Dropout(attn_dropout)

    def forward(self, x):
        """x (B, in_channel, N) -> (B, out_channel, N)"""
        if self.training:
            raise NotImplementedError(
                "r3dfsseg_b200: no stand-alone training-mode forward; attention dropout and its "
                "backward run inside MPTI_SelfAtten.forward(..., train=True) — call .eval()")
        if self.out_channel != 64:
            raise NotImplementedError("the attention kernel is built for out_channel = 64")
        wqkv = torch.cat([self.q_map.weight, self.k_map.weight, self.v_map.weight], 0)
        y = ops.op.attention(x.transpose(1, 2), wqkv.reshape(192, -1))
        return y.transpose(1, 2)
