"""FNF1d / FNF2d: functions -> vectors (reference: models/func_to_vec1d.py, func_to_vec2d.py)."""
import torch.nn as nn

from ..shared import MLP, LinearFunctionals1d, LinearFunctionals2d, SpectralConv1d, SpectralConv2d, _get_act
from ._trunk import Pending, functional_end, functional_end_is_fused, lift, run_layers


class FNF2d(nn.Module):
    """Fourier Neural Functional; same interface / state_dict as func_to_vec2d.py:8-112."""

    def __init__(self, modes1=12, modes2=12, width=32, width_final=128, padding=8, d_in=1, d_out=1,
                 width_lfunc=None, act="gelu", n_layers=4, get_grid=True):
        super().__init__()
        self.d_physical = 2
        self.modes1, self.modes2, self.width, self.width_final = modes1, modes2, width, width_final
        self.padding, self.d_in, self.d_out = padding, d_in, d_out
        self.width_lfunc = width if width_lfunc is None else width_lfunc
        self.act_name, self.act = act, _get_act(act)
        self.n_layers = 4 if n_layers is None else n_layers
        self.get_grid = get_grid
        self.fc0 = nn.Linear(d_in + self.d_physical if get_grid else d_in, width)
        self.speconvs = nn.ModuleList([SpectralConv2d(width, width, modes1, modes2) for _ in range(self.n_layers - 1)])
        self.ws = nn.ModuleList([nn.Conv2d(width, width, 1) for _ in range(self.n_layers - 1)])
        self.lfunc0 = LinearFunctionals2d(width, self.width_lfunc, modes1, modes2)
        self.mlpfunc0 = MLP(width, width_final, self.width_lfunc, act)
        self.mlp0 = MLP(2 * self.width_lfunc, 2 * width_final, d_out, act)

    def forward(self, x):
        """(batch, d_in, nx, ny) -> (batch, d_out)"""
        x, _ = lift(x, self.fc0, self.padding, self.get_grid, one_d=False)
        state = run_layers(Pending(x), self.speconvs, self.ws, self.act_name, [True] * len(self.ws), False,
                           end_takes_grad=functional_end_is_fused(x, self.mlpfunc0, False))
        return functional_end(state, self.lfunc0, self.mlpfunc0, self.mlp0, one_d=False)


class FNF1d(nn.Module):
    """Same interface / state_dict as func_to_vec1d.py:8-110."""

    def __init__(self, modes1=16, width=64, width_final=128, padding=8, d_in=1, d_out=1, width_lfunc=None,
                 act="gelu", n_layers=4, get_grid=True):
        super().__init__()
        self.d_physical = 1
        self.modes1, self.width, self.width_final = modes1, width, width_final
        self.padding, self.d_in, self.d_out = padding, d_in, d_out
        self.width_lfunc = width if width_lfunc is None else width_lfunc
        self.act_name, self.act = act, _get_act(act)
        self.n_layers = 4 if n_layers is None else n_layers
        self.get_grid = get_grid
        self.fc0 = nn.Linear(d_in + self.d_physical if get_grid else d_in, width)
        self.speconvs = nn.ModuleList([SpectralConv1d(width, width, modes1) for _ in range(self.n_layers - 1)])
        self.ws = nn.ModuleList([nn.Conv1d(width, width, 1) for _ in range(self.n_layers - 1)])
        self.lfunc0 = LinearFunctionals1d(width, self.width_lfunc, modes1)
        self.mlpfunc0 = MLP(width, width_final, self.width_lfunc, act)
        self.mlp0 = MLP(2 * self.width_lfunc, 2 * width_final, d_out, act)

    def forward(self, x):
        """(batch, d_in, nx) -> (batch, d_out)"""
        x, _ = lift(x, self.fc0, self.padding, self.get_grid, one_d=True)
        state = run_layers(Pending(x), self.speconvs, self.ws, self.act_name, [True] * len(self.ws), True,
                           end_takes_grad=functional_end_is_fused(x, self.mlpfunc0, True))
        return functional_end(state, self.lfunc0, self.mlpfunc0, self.mlp0, one_d=True)
