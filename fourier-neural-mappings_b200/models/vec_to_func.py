"""FND1d / FND2d: vectors -> functions (reference: models/vec_to_func1d.py, vec_to_func2d.py)."""
import torch.nn as nn

from .. import ops
from ..shared import MLP, LinearDecoder1d, LinearDecoder2d, SpectralConv1d, SpectralConv2d, _get_act
from ._trunk import Pending, project, run_layers


class FND2d(nn.Module):
    """Fourier Neural Decoder; same interface / state_dict as vec_to_func2d.py:6-111."""

    def __init__(self, d_in, s_outputspace, modes1=12, modes2=12, width=32, width_initial=128, width_final=128,
                 padding=8, d_out=1, width_ldec=None, act="gelu", n_layers=4, nonlinear_first=True):
        super().__init__()
        self.d_physical = 2
        self.d_in = d_in
        self.modes1, self.modes2, self.width = modes1, modes2, width
        self.width_initial, self.width_final = width_initial, width_final
        self.padding, self.d_out = padding, d_out
        self.width_ldec = width if width_ldec is None else width_ldec
        self.act_name, self.act = act, _get_act(act)
        self.n_layers = 4 if n_layers is None else n_layers
        self.nonlinear_first = nonlinear_first
        self.set_outputspace_resolution(s_outputspace)
        self.mlp0 = MLP(d_in, width_initial, self.width_ldec, act) if nonlinear_first else \
            nn.Linear(d_in, self.width_ldec)
        self.ldec0 = LinearDecoder2d(self.width_ldec, width, modes1, modes2)
        self.speconvs = nn.ModuleList([SpectralConv2d(width, width, modes1, modes2) for _ in range(self.n_layers - 1)])
        self.ws = nn.ModuleList([nn.Conv2d(width, width, 1) for _ in range(self.n_layers - 1)])
        self.mlp1 = MLP(width, width_final, d_out, act)

    def forward(self, x):
        """(batch, d_in) -> (batch, d_out, nx_out, ny_out)"""
        v = self.mlp0(x)
        y = ops.linear_decoder(v, self.ldec0.weights, self.ldec0.tables(self.s_outputspace))
        n = len(self.ws)
        state = run_layers(Pending(y), self.speconvs, self.ws, self.act_name, [True] * (n - 1) + [False], False)
        return project(state.value(), self.num_pad_outputspace, self.mlp1, one_d=False).permute(0, 3, 1, 2)

    def set_outputspace_resolution(self, s):
        self.s_outputspace = tuple(r + r // self.padding for r in list(s))
        self.num_pad_outputspace = tuple(r // self.padding for r in list(s))


class FND1d(nn.Module):
    """Same interface / state_dict as vec_to_func1d.py:6-107."""

    def __init__(self, d_in, s_outputspace, modes1=16, width=64, width_initial=128, width_final=128, padding=8,
                 d_out=1, width_ldec=None, act="gelu", n_layers=4, nonlinear_first=True):
        super().__init__()
        self.d_physical = 1
        self.d_in = d_in
        self.modes1, self.width = modes1, width
        self.width_initial, self.width_final = width_initial, width_final
        self.padding, self.d_out = padding, d_out
        self.width_ldec = width if width_ldec is None else width_ldec
        self.act_name, self.act = act, _get_act(act)
        self.n_layers = 4 if n_layers is None else n_layers
        self.nonlinear_first = nonlinear_first
        self.set_outputspace_resolution(s_outputspace)
        self.mlp0 = MLP(d_in, width_initial, self.width_ldec, act) if nonlinear_first else \
            nn.Linear(d_in, self.width_ldec)
        self.ldec0 = LinearDecoder1d(self.width_ldec, width, modes1)
        self.speconvs = nn.ModuleList([SpectralConv1d(width, width, modes1) for _ in range(self.n_layers - 1)])
        self.ws = nn.ModuleList([nn.Conv1d(width, width, 1) for _ in range(self.n_layers - 1)])
        self.mlp1 = MLP(width, width_final, d_out, act)

    def forward(self, x):
        """(batch, d_in) -> (batch, d_out, nx_out)"""
        v = self.mlp0(x)
        y = ops.linear_decoder(v, self.ldec0.weights, self.ldec0.tables(self.s_outputspace))
        n = len(self.ws)
        state = run_layers(Pending(y), self.speconvs, self.ws, self.act_name, [True] * (n - 1) + [False], True)
        return project(state.value(), (self.num_pad_outputspace,), self.mlp1, one_d=True).squeeze(1).permute(0, 2, 1)

    def set_outputspace_resolution(self, s):
        self.s_outputspace = s + s // self.padding
        self.num_pad_outputspace = s // self.padding
