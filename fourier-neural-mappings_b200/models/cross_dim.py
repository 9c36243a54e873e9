"""FNO1d2 / FNO2d1: maps between 1-D and 2-D functions (reference: models/func1d_to_func2d.py,
func2d_to_func1d.py).  They reuse the 2-D trunk and the 1-D F2V / V2F ends with the extra grid axis
folded into the batch: the reference gets the same effect from the ``...`` of its einsum
(shared.py:53), applied to a (B, C, nx) vector field or a (B, C, nx, ny) function of the last axis.
On channels-last tensors that fold is a view, not a permute."""
import torch
import torch.nn as nn

from .. import ops
from ..shared import MLP, LinearDecoder1d, LinearFunctionals1d, SpectralConv2d, _get_act
from ._trunk import Pending, _fused_ends_ok, functional_end, functional_end_is_fused, lift, project, run_layers


class FNO1d2(nn.Module):
    """1-D functions -> 2-D functions; same constructor, forward and state_dict as
    func1d_to_func2d.py:8-134."""

    def __init__(self, modes1d=16, width1d=96, modes1=12, modes2=12, width=32, s_outputspace=None,
                 width_final=128, padding=8, d_in=1, d_out=1, act="gelu", n_layers=4, get_grid=True):
        super().__init__()
        self.d_physical = 1
        self.modes1d, self.width1d = modes1d, width1d
        self.modes1, self.modes2, self.width, self.width_final = modes1, modes2, width, width_final
        self.padding, self.d_in, self.d_out = padding, d_in, d_out
        self.act_name, self.act = act, _get_act(act)
        self.n_layers = 4 if n_layers is None else n_layers
        self.get_grid = get_grid
        self.set_outputspace_resolution(s_outputspace)
        self.fc0 = nn.Linear(d_in + self.d_physical if get_grid else d_in, width1d)
        self.ldec0 = LinearDecoder1d(width1d, width, modes1d)
        self.speconvs = nn.ModuleList([SpectralConv2d(width, width, modes1, modes2) for _ in range(self.n_layers - 1)])
        self.ws = nn.ModuleList([nn.Conv2d(width, width, 1) for _ in range(self.n_layers - 1)])
        self.mlp0 = MLP(width, width_final, d_out, act)

    def forward(self, x):
        """(batch, d_in, nx) -> (batch, d_out, nx_out, ny_out).  Like the reference, the decoded second axis has
        the UNPADDED input length (``self.ldec0(x, x_res)``, func1d_to_func2d.py:97) and is cropped anyway."""
        x, (n,) = lift(x, self.fc0, self.padding, self.get_grid, one_d=True)        # (B, 1, P, width1d)
        B, _, P, W1 = x.shape
        # every grid point x carries a vector of width1d channels that is decoded into a function of y
        y = ops.linear_decoder(x.reshape(B * P, W1), self.ldec0.weights, self.ldec0.tables(n))
        y = y.view(B, P, n, self.width)
        nl = len(self.ws)
        state = run_layers(Pending(y), self.speconvs, self.ws, self.act_name, [True] * (nl - 1) + [False], False,
                           self.s_outputspace)
        cut = self.num_pad_outputspace if self.s_outputspace is not None else (n // self.padding, n // self.padding)
        return project(state.value(), cut, self.mlp0, one_d=False).permute(0, 3, 1, 2)

    def set_outputspace_resolution(self, s=None):
        if s is None:
            self.s_outputspace, self.num_pad_outputspace = None, None
        else:
            self.s_outputspace = tuple(r + r // self.padding for r in list(s))
            self.num_pad_outputspace = tuple(r // self.padding for r in list(s))


class FNO2d1(nn.Module):
    """2-D functions -> 1-D functions; same constructor, forward and state_dict as
    func2d_to_func1d.py:8-141.  (The reference never stores ``get_grid`` and its forward raises
    AttributeError until the attribute is set by hand, func2d_to_func1d.py:90; here the constructor
    argument is kept.)"""

    def __init__(self, modes1d=16, width1d=96, modes1=12, modes2=12, width=32, s_outputspace=None,
                 width_final=128, padding=8, d_in=1, d_out=1, act="gelu", n_layers=4, get_grid=True):
        super().__init__()
        self.d_physical = 2
        self.modes1d, self.width1d = modes1d, width1d
        self.modes1, self.modes2, self.width, self.width_final = modes1, modes2, width, width_final
        self.padding, self.d_in, self.d_out = padding, d_in, d_out
        self.act_name, self.act = act, _get_act(act)
        self.n_layers = 4 if n_layers is None else n_layers
        self.get_grid = get_grid
        self.set_outputspace_resolution(s_outputspace)
        self.fc0 = nn.Linear(d_in + self.d_physical if get_grid else d_in, width)
        self.speconvs = nn.ModuleList([SpectralConv2d(width, width, modes1, modes2) for _ in range(self.n_layers - 1)])
        self.ws = nn.ModuleList([nn.Conv2d(width, width, 1) for _ in range(self.n_layers - 1)])
        self.lfunc0 = LinearFunctionals1d(width, width1d, modes1d)
        self.mlpfunc0 = MLP(width, width_final, width1d, act)
        self.mlp0 = MLP(2 * width1d, 2 * width_final, d_out, act)

    def forward(self, x):
        """(batch, d_in, nx, ny) -> (batch, d_out, nx_out)"""
        x, (n1, n2) = lift(x, self.fc0, self.padding, self.get_grid, one_d=False)
        state = run_layers(Pending(x), self.speconvs, self.ws, self.act_name, [True] * len(self.ws), False,
                           end_takes_grad=functional_end_is_fused(x, self.mlpfunc0, True, rows=x.shape[0] * x.shape[1]))
        B, P1, P2, Cw = state.z.shape
        # functionals of the LAST axis for every x: rows (b, x) are a batch of 1-D functions
        rows = Pending(state.z.view(B * P1, 1, P2, Cw), state.act,
                       None if state.x is None else state.x.view(B * P1, 1, P2, Cw), state.zgrad)
        feats = functional_end(rows, self.lfunc0, self.mlpfunc0, lambda f: f, one_d=True)    # (B*P1, 2*width1d)
        f = feats.view(B, 1, P1, 2 * self.width1d)
        if self.s_outputspace is not None and self.s_outputspace != P1:
            if _fused_ends_ok(f):
                f = ops.project_resolution(f, self.s_outputspace, True)
            else:
                from ..shared import projector1d
                f = projector1d(f.squeeze(1).permute(0, 2, 1), self.s_outputspace).permute(0, 2, 1).unsqueeze(1)
        cut = (self.num_pad_outputspace,) if self.s_outputspace is not None else (n2 // self.padding,)
        return project(f, cut, self.mlp0, one_d=True).squeeze(1).permute(0, 2, 1)

    def set_outputspace_resolution(self, s=None):
        if s is None:
            self.s_outputspace, self.num_pad_outputspace = None, None
        else:
            self.s_outputspace = s + s // self.padding
            self.num_pad_outputspace = s // self.padding
