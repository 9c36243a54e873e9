"""FNO1d / FNO2d: functions -> functions (reference: models/func_to_func1d.py, func_to_func2d.py)."""
import torch.nn as nn

from ..shared import MLP, SpectralConv1d, SpectralConv2d, _get_act
from ._trunk import Pending, lift, project, run_layers


class FNO2d(nn.Module):
    """Same constructor, forward signature and state_dict as the reference FNO2d
    (func_to_func2d.py:8-118)."""

    def __init__(self, modes1=12, modes2=12, width=32, s_outputspace=None, width_final=128, padding=8,
                 d_in=1, d_out=1, act="gelu", n_layers=4, get_grid=True):
        super().__init__()
        self.d_physical = 2
        self.modes1, self.modes2, self.width, self.width_final = modes1, modes2, width, width_final
        self.padding, self.d_in, self.d_out = padding, d_in, d_out
        self.act_name, self.act = act, _get_act(act)
        self.n_layers = 4 if n_layers is None else n_layers
        self.get_grid = get_grid
        self.set_outputspace_resolution(s_outputspace)
        self.fc0 = nn.Linear(d_in + self.d_physical if get_grid else d_in, width)
        self.speconvs = nn.ModuleList([SpectralConv2d(width, width, modes1, modes2) for _ in range(self.n_layers)])
        self.ws = nn.ModuleList([nn.Conv2d(width, width, 1) for _ in range(self.n_layers)])
        self.mlp0 = MLP(width, width_final, d_out, act)

    def forward(self, x):
        """(batch, d_in, nx, ny) -> (batch, d_out, nx_out, ny_out)"""
        x, res = lift(x, self.fc0, self.padding, self.get_grid, one_d=False)
        flags = [True] * (self.n_layers - 1) + [False]
        state = run_layers(Pending(x), self.speconvs, self.ws, self.act_name, flags, False, self.s_outputspace)
        cut = self.num_pad_outputspace if self.s_outputspace is not None else tuple(r // self.padding for r in res)
        return project(state.value(), cut, self.mlp0, one_d=False).permute(0, 3, 1, 2)

    def set_outputspace_resolution(self, s=None):
        if s is None:
            self.s_outputspace, self.num_pad_outputspace = None, None
        else:
            self.s_outputspace = tuple(r + r // self.padding for r in list(s))
            self.num_pad_outputspace = tuple(r // self.padding for r in list(s))


class FNO1d(nn.Module):
    """Same constructor, forward signature and state_dict as the reference FNO1d
    (func_to_func1d.py:8-114)."""

    def __init__(self, modes1=16, width=64, s_outputspace=None, width_final=128, padding=8, d_in=1, d_out=1,
                 act="gelu", n_layers=4, get_grid=True):
        super().__init__()
        self.d_physical = 1
        self.modes1, self.width, self.width_final = modes1, width, width_final
        self.padding, self.d_in, self.d_out = padding, d_in, d_out
        self.act_name, self.act = act, _get_act(act)
        self.n_layers = 4 if n_layers is None else n_layers
        self.get_grid = get_grid
        self.set_outputspace_resolution(s_outputspace)
        self.fc0 = nn.Linear(d_in + self.d_physical if get_grid else d_in, width)
        self.speconvs = nn.ModuleList([SpectralConv1d(width, width, modes1) for _ in range(self.n_layers)])
        self.ws = nn.ModuleList([nn.Conv1d(width, width, 1) for _ in range(self.n_layers)])
        self.mlp0 = MLP(width, width_final, d_out, act)

    def forward(self, x):
        """(batch, d_in, nx) -> (batch, d_out, nx_out)"""
        x, res = lift(x, self.fc0, self.padding, self.get_grid, one_d=True)
        flags = [True] * (self.n_layers - 1) + [False]
        state = run_layers(Pending(x), self.speconvs, self.ws, self.act_name, flags, True, self.s_outputspace)
        cut = (self.num_pad_outputspace,) if self.s_outputspace is not None else (res[0] // self.padding,)
        return project(state.value(), cut, self.mlp0, one_d=True).squeeze(1).permute(0, 2, 1)

    def set_outputspace_resolution(self, s=None):
        if s is None:
            self.s_outputspace, self.num_pad_outputspace = None, None
        else:
            self.s_outputspace = s + s // self.padding
            self.num_pad_outputspace = s // self.padding
