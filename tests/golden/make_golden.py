"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference ships no tests or known-answer vectors, so these files pin the
oracle (and through it the CUDA path) to outputs of the reference implementation itself:
  * models/shared.py blocks  : forward outputs + autograd gradients at small / edge shapes
  * models/*.py              : tiny model forward, gradients, and parameters after 3 steps of the
                               reference's util/Adam.py
  * util/Adam.py             : complex + real parameters, weight decay, amsgrad on/off
Everything is seeded; inputs, weights and outputs are stored so no RNG stream has to match.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.modules.setdefault("hdf5storage", types.ModuleType("hdf5storage"))   # utilities_module.py:12

import models                                                            # noqa: E402
from models import shared                                                # noqa: E402


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


RefAdam = _load(f"{REF}/util/Adam.py", "ref_adam").Adam
LpLoss = _load(f"{REF}/util/utilities_module.py", "ref_utils").LpLoss


def npz(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(f"{name}: {sum(a.nbytes for a in out.values()) / 1e3:.1f} kB")


def randn(*shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def block_case(tag, mod, x, fwd_kwargs):
    """forward + autograd grads of one shared.py block with upstream gradient g ~ N(0,1)."""
    x = x.clone().requires_grad_(True)
    y = mod(x, **fwd_kwargs)
    g = randn(*y.shape, seed=99)
    y.backward(g)
    d = {f"{tag}.x": x, f"{tag}.y": y, f"{tag}.g": g, f"{tag}.dx": x.grad}
    for n, p in mod.named_parameters():
        d[f"{tag}.{n}"] = p
        d[f"{tag}.d{n}"] = p.grad
    return d


def blocks():
    torch.manual_seed(0)
    d = {}
    # SpectralConv2d: odd grid, Nyquist-touching modes, non-square, resolution change (down & up)
    for tag, (ci, co, p1, p2, m1, m2, s) in {
        "sc2_odd": (3, 4, 9, 9, 2, 5, None),
        "sc2_nyq": (2, 3, 8, 8, 4, 5, None),
        "sc2_rect": (4, 2, 12, 10, 3, 4, None),
        "sc2_down": (3, 3, 16, 12, 3, 3, (8, 10)),
        "sc2_up": (2, 2, 8, 10, 2, 3, (12, 15)),
        "sc2_cut": (2, 2, 12, 12, 4, 5, (6, 6)),
    }.items():
        mod = shared.SpectralConv2d(ci, co, m1, m2)
        d.update(block_case(tag, mod, randn(2, ci, p1, p2, seed=1), {"s": s}))
        d[f"{tag}.s"] = np.array([-1, -1] if s is None else s)
    for tag, (ci, co, p, m, s) in {
        "sc1_even": (3, 4, 16, 5, None),
        "sc1_odd": (2, 2, 9, 4, None),
        "sc1_nyq": (2, 3, 8, 5, None),
        "sc1_down": (2, 2, 16, 4, 10),
        "sc1_up": (2, 2, 9, 3, 14),
    }.items():
        mod = shared.SpectralConv1d(ci, co, m)
        d.update(block_case(tag, mod, randn(2, ci, p, seed=2), {"s": s}))
        d[f"{tag}.s"] = np.array([-1 if s is None else s])
    # F2V: normal, odd grid, zero-padded (2*m1 >= P1 and m2+1 > P2//2+1)
    for tag, (ci, co, p1, p2, m1, m2) in {
        "lf2": (3, 4, 12, 10, 3, 3),
        "lf2_odd": (2, 3, 9, 7, 2, 2),
        "lf2_pad": (2, 2, 6, 6, 4, 4),
    }.items():
        mod = shared.LinearFunctionals2d(ci, co, m1, m2)
        d.update(block_case(tag, mod, randn(2, ci, p1, p2, seed=3), {}))
    for tag, (ci, co, p, m) in {"lf1": (3, 4, 16, 5), "lf1_odd": (2, 2, 9, 3), "lf1_pad": (2, 2, 6, 5)}.items():
        mod = shared.LinearFunctionals1d(ci, co, m)
        d.update(block_case(tag, mod, randn(2, ci, p, seed=4), {}))
    # V2F: normal, odd, truncating (s small)
    for tag, (ci, co, m1, m2, s) in {
        "ld2": (3, 4, 3, 3, (12, 10)),
        "ld2_odd": (2, 3, 2, 2, (9, 7)),
        "ld2_cut": (2, 2, 4, 4, (6, 5)),
        "ld2_cut_odd": (2, 2, 4, 3, (5, 8)),
    }.items():
        mod = shared.LinearDecoder2d(ci, co, m1, m2)
        d.update(block_case(tag, mod, randn(2, ci, seed=5), {"s": s}))
        d[f"{tag}.s"] = np.array(s)
    for tag, (ci, co, m, s) in {"ld1": (3, 4, 5, 16), "ld1_odd": (2, 2, 3, 9), "ld1_cut": (2, 2, 6, 8)}.items():
        mod = shared.LinearDecoder1d(ci, co, m)
        d.update(block_case(tag, mod, randn(2, ci, seed=6), {"s": s}))
        d[f"{tag}.s"] = np.array([s])
    # projectors
    xp = randn(2, 2, 12, 10, seed=7)
    d["proj2.x"] = xp
    d["proj2.down"] = shared.projector2d(xp, (8, 6))
    d["proj2.up"] = shared.projector2d(xp, (16, 14))
    xp1 = randn(2, 2, 12, seed=8)
    d["proj1.x"] = xp1
    d["proj1.down"] = shared.projector1d(xp1, 7)
    d["proj1.up"] = shared.projector1d(xp1, 20)
    npz("blocks.npz", **d)


def model_case(name, ctor, x_shape, steps=3):
    """Tiny model: weights, forward, loss, grads, and parameters after `steps` reference-Adam steps."""
    torch.manual_seed(0)
    model = ctor()
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    x = randn(*x_shape, seed=1234)
    with torch.no_grad():
        y_shape = model(x).shape
    y = randn(*y_shape, seed=1235)
    loss_fn = LpLoss(size_average=False)
    opt = RefAdam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    d = {"x": x, "y": y}
    d.update({f"w0.{k}": v for k, v in sd0.items()})
    for t in range(steps):
        opt.zero_grad()
        out = model(x)
        loss = loss_fn(out, y)
        loss.backward()
        if t == 0:
            d["out0"] = out
            d["loss0"] = loss
            for k, p in model.named_parameters():
                d[f"g0.{k}"] = p.grad if p.grad is not None else torch.zeros_like(p)
        opt.step()
    d["loss_last"] = loss
    d.update({f"w{steps}.{k}": v for k, v in model.state_dict().items()})
    npz(f"model_{name}.npz", **d)


def model_cases():
    model_case("fno2d", lambda: models.FNO2d(3, 3, 8, n_layers=3, width_final=16), (2, 1, 16, 16))
    model_case("fno2d_res", lambda: models.FNO2d(3, 3, 8, s_outputspace=(8, 8), n_layers=2, width_final=16, d_in=2, d_out=2),
               (2, 2, 16, 16))
    model_case("fno1d", lambda: models.FNO1d(5, 8, n_layers=3, width_final=16), (3, 1, 32))
    model_case("fnf2d", lambda: models.FNF2d(3, 3, 8, width_final=16, d_in=2, d_out=2, n_layers=3), (2, 2, 17, 16))
    model_case("fnf1d", lambda: models.FNF1d(5, 8, width_final=16, d_out=3, n_layers=3), (2, 1, 32))
    model_case("fnd2d", lambda: models.FND2d(5, (16, 16), modes1=3, modes2=3, width=8, width_initial=16,
                                             width_final=16, n_layers=3), (2, 5))
    model_case("fnd1d", lambda: models.FND1d(5, 32, modes1=5, width=8, width_initial=16, width_final=16,
                                             n_layers=3), (2, 5))
    model_case("fnn2d", lambda: models.FNN2d(5, 3, s_latentspace=(16, 16), modes1=3, modes2=3, width=8,
                                             width_initial=16, width_final=16, n_layers=4), (2, 5))
    model_case("fnn1d", lambda: models.FNN1d(5, 3, s_latentspace=32, modes1=5, width=8, width_initial=16,
                                             width_final=16, n_layers=3), (2, 5))


def _fno2d1(*a, **kw):
    """The reference's FNO2d1 never stores ``get_grid`` (func2d_to_func1d.py:8-75) and its forward reads it
    (:90), so the stock class raises AttributeError; the attribute is set here, nothing else is touched."""
    m = models.FNO2d1(*a, **kw)
    m.get_grid = True
    return m


def cross_dim_cases():
    """Cross-dimension models and extra resolution-change cases; kept in their own
    files so that the fixtures above stay byte-identical."""
    model_case("fno1d2", lambda: models.FNO1d2(4, 12, 3, 3, 8, width_final=16, n_layers=3), (2, 1, 32))
    model_case("fno1d2_res", lambda: models.FNO1d2(4, 12, 3, 3, 8, s_outputspace=(16, 11), width_final=16, n_layers=3,
                                                   d_in=2, d_out=2), (2, 2, 32))
    model_case("fno2d1", lambda: _fno2d1(4, 12, 3, 3, 8, width_final=16, n_layers=3), (2, 1, 16, 16))
    model_case("fno2d1_res", lambda: _fno2d1(4, 12, 3, 3, 8, s_outputspace=24, width_final=16, n_layers=3, d_in=2,
                                             d_out=2), (2, 2, 16, 14))
    model_case("fno1d_res", lambda: models.FNO1d(5, 8, s_outputspace=21, n_layers=3, width_final=16), (3, 1, 32))
    model_case("fno2d_res_odd", lambda: models.FNO2d(3, 3, 8, s_outputspace=(11, 23), n_layers=2, width_final=16),
               (2, 1, 17, 16))
    d = {}
    for tag, (shape, s) in {"p2_odd_up": ((2, 3, 9, 7), (14, 11)), "p2_odd_down": ((2, 3, 19, 17), (5, 7)),
                            "p2_even_down": ((2, 2, 16, 16), (10, 16)), "p2_mixed": ((2, 2, 8, 9), (16, 4))}.items():
        x = randn(*shape, seed=11)
        d[f"{tag}.x"], d[f"{tag}.y"], d[f"{tag}.s"] = x, shared.projector2d(x, s), np.array(s)
    for tag, (shape, s) in {"p1_odd_up": ((2, 3, 9), 14), "p1_odd_down": ((2, 3, 19), 5)}.items():
        x = randn(*shape, seed=12)
        d[f"{tag}.x"], d[f"{tag}.y"], d[f"{tag}.s"] = x, shared.projector1d(x, s), np.array([s])
    npz("projectors.npz", **d)
    import json
    lay = {}
    for name, c in {"FNO1d2": lambda: models.FNO1d2(), "FNO2d1": lambda: models.FNO2d1()}.items():
        lay[name] = {k: [list(v.shape), str(v.dtype)] for k, v in c().state_dict().items()}
    with open(os.path.join(HERE, "state_dict_layouts_cross_dim.json"), "w") as f:
        json.dump(lay, f, indent=1)


def adam_cases():
    """NB: the reference's amsgrad branch calls torch.maximum on the complex second moment and
    raises for complex parameters (Adam.py:43), so the amsgrad case holds a real parameter only."""
    d = {}
    for tag, kw in {"plain": dict(lr=1e-3), "wd": dict(lr=2e-3, weight_decay=1e-2),
                    "ams": dict(lr=1e-3, weight_decay=1e-4, amsgrad=True)}.items():
        torch.manual_seed(7)
        pc = torch.nn.Parameter(torch.randn(4, 3, 5, dtype=torch.cfloat))
        pr = torch.nn.Parameter(torch.randn(6, 7))
        use_c = not kw.get("amsgrad", False)
        opt = RefAdam([pc, pr] if use_c else [pr], **kw)
        d[f"{tag}.pc0"], d[f"{tag}.pr0"] = pc.detach().clone(), pr.detach().clone()
        for t in range(4):
            gc = torch.randn(4, 3, 5, dtype=torch.cfloat, generator=torch.Generator().manual_seed(100 + t))
            gr = randn(6, 7, seed=200 + t) * (3.0 if t == 1 else 1.0)     # make the running max matter
            pc.grad, pr.grad = (gc.clone() if use_c else None), gr.clone()
            opt.step()
            d[f"{tag}.gc{t}"], d[f"{tag}.gr{t}"] = gc, gr
            d[f"{tag}.pc{t + 1}"], d[f"{tag}.pr{t + 1}"] = pc.detach().clone(), pr.detach().clone()
        if use_c:
            st = opt.state[pc]
            d[f"{tag}.mc"], d[f"{tag}.vc"] = st["exp_avg"], st["exp_avg_sq"]
    npz("adam.npz", **d)


def state_dict_layouts():
    """Key / shape / dtype listing of the reference's state_dict for every model (the persistence
    contract)."""
    import json
    ctors = {
        "FNO2d": lambda: models.FNO2d(12, 12, 32),
        "FNO1d": lambda: models.FNO1d(16, 64),
        "FNF2d": lambda: models.FNF2d(12, 12, 32, d_in=2, d_out=2),
        "FNF1d": lambda: models.FNF1d(16, 64),
        "FND2d": lambda: models.FND2d(20, (128, 128), modes1=16, modes2=16, width=64),
        "FND2d_linear_first": lambda: models.FND2d(20, (32, 32), nonlinear_first=False),
        "FND1d": lambda: models.FND1d(20, 1024),
        "FNN2d": lambda: models.FNN2d(20, 4),
        "FNN1d": lambda: models.FNN1d(20, 4),
    }
    out = {}
    for name, c in ctors.items():
        out[name] = {k: [list(v.shape), str(v.dtype)] for k, v in c().state_dict().items()}
    with open(os.path.join(HERE, "state_dict_layouts.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("state_dict_layouts.json")


if __name__ == "__main__":
    torch.set_num_threads(4)
    if sys.argv[1:] == ["cross_dim"]:
        cross_dim_cases()
        sys.exit(0)
    blocks()
    model_cases()
    adam_cases()
    state_dict_layouts()
    cross_dim_cases()
