"""FNN1d / FNN2d: vectors -> vectors (reference: models/vec_to_vec1d.py, vec_to_vec2d.py)."""
import torch.nn as nn

from .. import ops
from ..shared import (MLP, LinearDecoder1d, LinearDecoder2d, LinearFunctionals1d, LinearFunctionals2d,
                      SpectralConv1d, SpectralConv2d, _get_act)
from ._trunk import Pending, functional_end, functional_end_is_fused, run_layers


class FNN2d(nn.Module):
    """Fourier Neural Network; same interface / state_dict as vec_to_vec2d.py:7-125."""

    def __init__(self, d_in, d_out, s_latentspace=(64, 64), modes1=12, modes2=12, width=32, width_initial=128,
                 width_final=128, width_ldec=None, width_lfunc=None, act="gelu", n_layers=4, nonlinear_first=True):
        super().__init__()
        self.d_physical = 2
        self.d_in, self.d_out = d_in, d_out
        self.modes1, self.modes2, self.width = modes1, modes2, width
        self.width_initial, self.width_final = width_initial, width_final
        self.width_ldec = width if width_ldec is None else width_ldec
        self.width_lfunc = width if width_lfunc is None else width_lfunc
        self.act_name, self.act = act, _get_act(act)
        self.n_layers = 4 if n_layers is None else n_layers
        if self.n_layers < 2:
            raise ValueError("n_layers for vec-to-vec models must be greater than or equal to 2")
        self.nonlinear_first = nonlinear_first
        self.set_latentspace_resolution(s_latentspace)
        self.mlp0 = MLP(d_in, width_initial, self.width_ldec, act) if nonlinear_first else \
            nn.Linear(d_in, self.width_ldec)
        self.ldec0 = LinearDecoder2d(self.width_ldec, width, modes1, modes2)
        self.speconvs = nn.ModuleList([SpectralConv2d(width, width, modes1, modes2) for _ in range(self.n_layers - 2)])
        self.ws = nn.ModuleList([nn.Conv2d(width, width, 1) for _ in range(self.n_layers - 2)])
        self.lfunc0 = LinearFunctionals2d(width, self.width_lfunc, modes1, modes2)
        self.mlpfunc0 = MLP(width, width_final, self.width_lfunc, act)
        self.mlp1 = MLP(2 * self.width_lfunc, 2 * width_final, d_out, act)

    def forward(self, x):
        """(batch, d_in) -> (batch, d_out)"""
        v = self.mlp0(x)
        y = ops.linear_decoder(v, self.ldec0.weights, self.ldec0.tables(self.s_latentspace))
        state = run_layers(Pending(y), self.speconvs, self.ws, self.act_name, [True] * len(self.ws), False,
                           end_takes_grad=functional_end_is_fused(y, self.mlpfunc0, False))
        return functional_end(state, self.lfunc0, self.mlpfunc0, self.mlp1, one_d=False)

    def set_latentspace_resolution(self, s):
        self.s_latentspace = s


class FNN1d(nn.Module):
    """Same interface / state_dict as vec_to_vec1d.py:7-122."""

    def __init__(self, d_in, d_out, s_latentspace=1024, modes1=16, width=64, width_initial=128, width_final=128,
                 width_ldec=None, width_lfunc=None, act="gelu", n_layers=4, nonlinear_first=True):
        super().__init__()
        self.d_physical = 1
        self.d_in, self.d_out = d_in, d_out
        self.modes1, self.width = modes1, width
        self.width_initial, self.width_final = width_initial, width_final
        self.width_ldec = width if width_ldec is None else width_ldec
        self.width_lfunc = width if width_lfunc is None else width_lfunc
        self.act_name, self.act = act, _get_act(act)
        self.n_layers = 4 if n_layers is None else n_layers
        if self.n_layers < 2:
            raise ValueError("n_layers for vec-to-vec models must be greater than or equal to 2")
        self.nonlinear_first = nonlinear_first
        self.set_latentspace_resolution(s_latentspace)
        self.mlp0 = MLP(d_in, width_initial, self.width_ldec, act) if nonlinear_first else \
            nn.Linear(d_in, self.width_ldec)
        self.ldec0 = LinearDecoder1d(self.width_ldec, width, modes1)
        self.speconvs = nn.ModuleList([SpectralConv1d(width, width, modes1) for _ in range(self.n_layers - 2)])
        self.ws = nn.ModuleList([nn.Conv1d(width, width, 1) for _ in range(self.n_layers - 2)])
        self.lfunc0 = LinearFunctionals1d(width, self.width_lfunc, modes1)
        self.mlpfunc0 = MLP(width, width_final, self.width_lfunc, act)
        self.mlp1 = MLP(2 * self.width_lfunc, 2 * width_final, d_out, act)

    def forward(self, x):
        """(batch, d_in) -> (batch, d_out)"""
        v = self.mlp0(x)
        y = ops.linear_decoder(v, self.ldec0.weights, self.ldec0.tables(self.s_latentspace))
        state = run_layers(Pending(y), self.speconvs, self.ws, self.act_name, [True] * len(self.ws), True,
                           end_takes_grad=functional_end_is_fused(y, self.mlpfunc0, True))
        return functional_end(state, self.lfunc0, self.mlpfunc0, self.mlp1, one_d=True)

    def set_latentspace_resolution(self, s):
        self.s_latentspace = s
