"""Host-side mirror of the reference's ConvPredictor (rnnt/predictor.py:189-229, rnnt/causalconv.py:9-39).

Not a kernel target of this round: its output is an input of the hot path.  It exists so that the
drop-in RNNTModel and the batched greedy decode can be built and tested without the reference tree; parameter names
match the reference's state_dict.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


class CausalConv1d(torch.nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, dilation, additional_context: int = 0):
        super().__init__()
        self.conv = torch.nn.Conv1d(in_channels, out_channels, kernel_size, stride, dilation=dilation)
        self.padding = (kernel_size - 1) * dilation - stride + 1
        if additional_context < 0:
            raise ValueError("additional_context must be non-negative")
        if additional_context > self.padding:
            raise ValueError("additional_context can't be greater than the padding")
        self.additional_context = additional_context
        self.left_padding = self.padding - additional_context

    def forward(self, x):
        return self.conv(F.pad(x, (self.left_padding, 0)))


class ConvPredictor(torch.nn.Module):
    RECEPTIVE_FIELD = 7   # 1 + (3-1) + (5-1) tokens

    def __init__(self, num_symbols: int, output_dim: int, symbol_embedding_dim: int, dropout: float) -> None:
        super().__init__()
        self.embedding = torch.nn.Embedding(num_symbols, symbol_embedding_dim)
        self.input_layer_norm = torch.nn.LayerNorm(symbol_embedding_dim)
        self.conv1 = CausalConv1d(symbol_embedding_dim, symbol_embedding_dim, kernel_size=3, stride=1, dilation=1)
        self.conv2 = CausalConv1d(symbol_embedding_dim, symbol_embedding_dim, kernel_size=5, stride=1, dilation=1)
        self.linear = torch.nn.Linear(symbol_embedding_dim, output_dim)
        self.output_layer_norm = torch.nn.LayerNorm(output_dim)
        self.dropout = torch.nn.Dropout(p=dropout)

    def forward(self, input):
        x = self.embedding(input)
        x = self.input_layer_norm(x)
        x = x.permute(0, 2, 1)
        x = self.dropout(F.gelu(self.conv1(x)))
        x = self.dropout(F.gelu(self.conv2(x)))
        x = x.permute(0, 2, 1)
        return self.output_layer_norm(self.linear(x))

    @torch.no_grad()
    def last_step(self, windows):
        """Predictor output for the LAST position of each row of `windows` (N, L<=7 or exactly 7 tokens).

        The output at position i depends only on tokens i-6..i, so a 7-token window reproduces what the reference's
        full-history re-run (rnnt/model.py:122-123) yields at the last position; shorter histories are passed
        whole so the left zero padding matches."""
        return self.forward(windows)[:, -1, :]


class ConvPredictorStepper:
    """Incremental evaluation of ConvPredictor for greedy decode: one new token per utterance per call.

    The predictor is causal with a receptive field of 7 tokens, so the reference's full-history re-run
    (rnnt/model.py:119-123) equals a streaming update that keeps, per utterance, the last 2 layer-normed embeddings
    (conv1, k=3) and the last 4 conv1 outputs (conv2, k=5).  Zero-initialised state IS the left zero padding of
    rnnt/causalconv.py:29, so short histories need no special case.  All tensors stay on the device; `advance`
    updates only the rows selected by `emit`.
    """

    def __init__(self, predictor: ConvPredictor, batch: int, device):
        p = predictor
        E = p.embedding.embedding_dim
        self.p = p
        self.E = E
        self.w1 = torch.empty(E, 3 * E, device=device)
        self.w2 = torch.empty(E, 5 * E, device=device)
        self.b1 = p.conv1.conv.bias.detach()
        self.b2 = p.conv2.conv.bias.detach()
        self.xs = torch.zeros(batch, 2, E, device=device)
        self.ys = torch.zeros(batch, 4, E, device=device)
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        """Re-read the conv weights (in place, so captured CUDA graphs stay valid) and zero the streaming state."""
        p, E = self.p, self.E
        # conv weight (Co, Ci, k) -> (Co, k*Ci) matching the [oldest ... newest] concatenation of the taps
        self.w1.copy_(p.conv1.conv.weight.detach().permute(0, 2, 1).reshape(E, -1))
        self.w2.copy_(p.conv2.conv.weight.detach().permute(0, 2, 1).reshape(E, -1))
        self.xs.zero_()
        self.ys.zero_()

    @torch.no_grad()
    def advance(self, tokens, emit):
        """tokens (B,) int64, emit (B,) bool -> predictor features (B, D) for rows with emit (others: garbage)."""
        p = self.p
        x = p.input_layer_norm(p.embedding(tokens))                                   # (B,E)
        y = F.gelu(F.linear(torch.cat([self.xs.flatten(1), x], 1), self.w1, self.b1))   # conv1 at the new position
        z = F.gelu(F.linear(torch.cat([self.ys.flatten(1), y], 1), self.w2, self.b2))   # conv2 at the new position
        out = p.output_layer_norm(p.linear(z))
        m = emit.view(-1, 1, 1)
        self.xs.copy_(torch.where(m, torch.cat([self.xs[:, 1:], x[:, None]], 1), self.xs))
        self.ys.copy_(torch.where(m, torch.cat([self.ys[:, 1:], y[:, None]], 1), self.ys))
        return out
