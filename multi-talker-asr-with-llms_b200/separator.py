"""Drop-in `Separator` for ref:models/separator.py (same constructor, parameter names and forward contract).

(1) pre_proj + act + LayerNorm  -> tcgen05 GEMM with fused ReLU epilogue + row LN kernel
(2) StackedCustomLSTM           -> per layer: one batched input GEMM + the persistent recurrent kernel (csrc/lstm.cu)
(3) post LayerNorm              -> row kernel
(4) N branches Linear-ReLU-Linear-ReLU-LN -> GEMMs with fused ReLU + row LN
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kernels as K
from . import ops
from . import precise

BF, F32 = torch.bfloat16, torch.float32


class CustomLSTMCell(nn.Module):
    """Parameter container: W = Linear(input+hidden -> 4*hidden), gate order i,f,g,o (ref:models/separator.py:6-24)."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size
        self.W = nn.Linear(input_size + hidden_size, 4 * hidden_size)


class StackedCustomLSTM(nn.Module):
    def __init__(self, input_size, hidden_size, num_layers, dropout=0.0, use_layernorm=False):
        super().__init__()
        self.num_layers = num_layers
        self.hidden_size = hidden_size
        self.cells = nn.ModuleList()
        self.norms = nn.ModuleList() if use_layernorm else None
        self.dropout = nn.Dropout(dropout)
        for i in range(num_layers):
            self.cells.append(CustomLSTMCell(input_size if i == 0 else hidden_size, hidden_size))
            if use_layernorm:
                self.norms.append(nn.LayerNorm(hidden_size))

    def forward(self, x):
        # The reference feeds layer l+1 with dropout(norm(h_l[t])) while h_l itself recurs un-normalised and
        # un-dropped (ref:models/separator.py:50-58), so norm/dropout separate cleanly from the recurrence.
        y = x
        for l, cell in enumerate(self.cells):
            y = ops.LSTMLayerFn.apply(y, cell.W.weight, cell.W.bias)
            if self.norms:
                y = ops.layer_norm(y, self.norms[l].weight, self.norms[l].bias, self.norms[l].eps, F32)
            y = self.dropout(y)
        return y


class Separator(nn.Module):
    def __init__(self, in_dim: int, hidden_size: int, talker_numbers: int, *, num_layers: int = 2, dropout: float = 0.2,
                 use_lstm_layernorm: bool = False, proj_activation: str | None = "relu", use_branch_ln: bool = True,
                 branch_dropout: float = 0.0, break_symmetry_eps: float = 1e-3):
        super().__init__()
        assert talker_numbers >= 2, "talker_numbers must be >= 2"
        if proj_activation not in ("relu", "gelu", None):
            raise KeyError(proj_activation)
        self.talker_numbers = talker_numbers
        self.hidden_size = hidden_size
        self.in_dim = in_dim
        self.proj_activation = proj_activation
        self.pre_proj = nn.Linear(in_dim, hidden_size, bias=True)
        self.pre_act = {"relu": nn.ReLU(), "gelu": nn.GELU(), None: nn.Identity()}[proj_activation]
        self.pre_ln = nn.LayerNorm(hidden_size)
        self.lstm = StackedCustomLSTM(hidden_size, hidden_size, num_layers, dropout=dropout, use_layernorm=use_lstm_layernorm)
        self.post_ln = nn.LayerNorm(hidden_size)

        def make_branch():
            layers = [nn.Linear(hidden_size, hidden_size), nn.ReLU()]
            if branch_dropout and branch_dropout > 0:
                layers.append(nn.Dropout(branch_dropout))
            layers += [nn.Linear(hidden_size, in_dim), nn.ReLU()]
            if use_branch_ln:
                layers.append(nn.LayerNorm(in_dim))
            return nn.Sequential(*layers)

        self.sep_branches = nn.ModuleList([make_branch() for _ in range(talker_numbers)])
        nn.init.xavier_uniform_(self.pre_proj.weight)
        nn.init.zeros_(self.pre_proj.bias)
        for bi, m in enumerate(self.sep_branches):
            lin1 = m[0]
            lin2 = m[3] if isinstance(m[2], nn.Dropout) else m[2]
            nn.init.xavier_uniform_(lin1.weight); nn.init.zeros_(lin1.bias)
            nn.init.xavier_uniform_(lin2.weight); nn.init.zeros_(lin2.bias)
            if break_symmetry_eps and break_symmetry_eps > 0:
                lin2.bias.data += break_symmetry_eps * bi

    def _branch(self, branch: nn.Sequential, y_bf16: torch.Tensor) -> torch.Tensor:
        h = y_bf16
        mods = list(branch)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear):
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                last_linear = not any(isinstance(x, nn.Linear) for x in mods[i + 1:])
                h = ops.linear(h, m.weight, m.bias, act=K.ACT_RELU if relu else K.ACT_NONE, out_dtype=F32 if last_linear else BF)
                i += 2 if relu else 1
            elif isinstance(m, nn.Dropout):
                h = m(h)
                i += 1
            elif isinstance(m, nn.LayerNorm):
                h = ops.layer_norm(h, m.weight, m.bias, m.eps, F32)
                i += 1
            else:
                raise NotImplementedError(type(m))
        return h

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        if precise.get_precision() == "fp32":      # fp32-class validation mode (forward + backward), see precise.py
            return precise.separator(self, x)
        act = {"relu": K.ACT_RELU, "gelu": K.ACT_GELU, None: K.ACT_NONE}[self.proj_activation]
        y = ops.linear(x, self.pre_proj.weight, self.pre_proj.bias, act=act, out_dtype=F32)
        y = ops.layer_norm(y, self.pre_ln.weight, self.pre_ln.bias, self.pre_ln.eps, BF)
        y = self.lstm(y)
        y = ops.layer_norm(y, self.post_ln.weight, self.post_ln.bias, self.post_ln.eps, BF)
        return [self._branch(br, y) for br in self.sep_branches]
