"""GPU: fused prune / cat bookkeeping against the reference's torch formulation
(scene/gaussian_model.py:424-442, :461-482 restated inline), including Adam state and `step`."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _model(P, seed):
    g = torch.Generator().manual_seed(seed)
    def t(*s):
        return nn.Parameter(torch.randn(*s, generator=g).cuda())
    ps = dict(xyz=t(P, 3), f_dc=t(P, 1, 3), f_rest=t(P, 15, 3), opacity=t(P, 1), scaling=t(P, 3), rotation=t(P, 4))
    multi = [t(8, 8), t(8)]
    from b200gs.adam import FusedAdam
    groups = [{"params": [ps["xyz"]], "lr": 1e-3, "name": "xyz"}, {"params": multi, "lr": 1e-3, "name": "deformation"}]
    groups += [{"params": [ps[k]], "lr": 1e-3, "name": k} for k in ("f_dc", "f_rest", "opacity", "scaling", "rotation")]
    opt = FusedAdam(groups, lr=0.0, eps=1e-15)
    for q in list(ps.values()) + multi:
        q.grad = torch.randn(q.shape, generator=g).cuda()
    opt.step(); opt.step()
    return ps, opt


def _ref_prune(opt, mask):
    out = {}
    for group in opt.param_groups:
        if len(group["params"]) > 1:
            continue
        st = opt.state.get(group["params"][0], None)
        st["exp_avg"] = st["exp_avg"][mask]; st["exp_avg_sq"] = st["exp_avg_sq"][mask]
        del opt.state[group["params"][0]]
        group["params"][0] = nn.Parameter(group["params"][0][mask].requires_grad_(True))
        opt.state[group["params"][0]] = st
        out[group["name"]] = group["params"][0]
    return out


def _ref_cat(opt, d):
    out = {}
    for group in opt.param_groups:
        if len(group["params"]) > 1:
            continue
        ext = d[group["name"]]
        st = opt.state.get(group["params"][0], None)
        st["exp_avg"] = torch.cat((st["exp_avg"], torch.zeros_like(ext)), dim=0)
        st["exp_avg_sq"] = torch.cat((st["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
        del opt.state[group["params"][0]]
        group["params"][0] = nn.Parameter(torch.cat((group["params"][0], ext), dim=0).requires_grad_(True))
        opt.state[group["params"][0]] = st
        out[group["name"]] = group["params"][0]
    return out


def _same(a, b, oa, ob):
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]) and a[k].requires_grad, k
        sa, sb = oa.state[a[k]], ob.state[b[k]]
        assert float(sa["step"]) == float(sb["step"]) == 2.0
        assert torch.equal(sa["exp_avg"], sb["exp_avg"]) and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"]), k


@pytest.mark.parametrize("P", [1000, 77777])
def test_prune_and_cat_match_reference_formulation(P):
    from b200gs import densify
    (pa, oa), (pb, ob) = _model(P, 1), _model(P, 1)
    g = torch.Generator().manual_seed(2)
    mask = (torch.rand(P, generator=g) > 0.3).cuda()
    _same(densify.prune_optimizer(oa, mask), _ref_prune(ob, mask), oa, ob)
    n_new = P // 7
    ext = {k: torch.randn((n_new,) + tuple(v.shape[1:]), generator=g).cuda() for k, v in pa.items()}
    _same(densify.cat_tensors_to_optimizer(oa, ext), _ref_cat(ob, {k: v.clone() for k, v in ext.items()}), oa, ob)
    # the optimiser keeps working on the new tensors
    for grp in oa.param_groups:
        for q in grp["params"]:
            q.grad = torch.ones_like(q)
    oa.step()
    # degenerate events
    none = torch.zeros(oa.param_groups[0]["params"][0].shape[0], dtype=torch.bool, device="cuda")
    out = densify.prune_optimizer(oa, none)
    assert out["xyz"].shape[0] == 0 and out["f_rest"].shape == (0, 15, 3)


# ---- whole events against the REAL reference GaussianModel (tests/golden/densify.pt, oracle/gen_golden_densify.py) -----------------
import os
import types

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "densify.pt")
_ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling", "rotation": "_rotation"}
_AUX = ("xyz_gradient_accum", "denom", "max_radii2D", "_deformation_accum", "_deformation_table", "_scene_flow")


def _load_model(snap):
    """A GaussianModel-shaped object (same attribute names, FusedAdam with the snapshot's moments and step counters)."""
    from b200gs.adam import FusedAdam
    m = types.SimpleNamespace(percent_dense=0.01)
    groups = []
    for name, attr in _ATTR.items():
        p = nn.Parameter(snap[name].clone().cuda())
        setattr(m, attr, p)
        groups.append({"params": [p], "lr": 1e-3, "name": name})
    multi = [nn.Parameter(torch.zeros(4, 4).cuda()), nn.Parameter(torch.zeros(4).cuda())]          # a multi-tensor group is left alone
    groups.insert(1, {"params": multi, "lr": 1e-3, "name": "deformation"})
    m.optimizer = FusedAdam(groups, lr=0.0, eps=1e-15)
    for name, attr in _ATTR.items():
        p = getattr(m, attr)
        m.optimizer.state[p] = {"step": torch.tensor(snap[name + ".step"]), "exp_avg": snap[name + ".exp_avg"].clone().cuda(),
                                "exp_avg_sq": snap[name + ".exp_avg_sq"].clone().cuda()}
    for a in _AUX:
        setattr(m, a, snap[a].clone().cuda())
    return m


def _check(m, snap, tag, exact_xyz_rows=None):
    n = snap["xyz"].shape[0]
    for name, attr in _ATTR.items():
        p = getattr(m, attr)
        assert p.shape == snap[name].shape and p.requires_grad and isinstance(p, nn.Parameter), (tag, name, tuple(p.shape), tuple(snap[name].shape))
        st = m.optimizer.state[p]
        assert float(st["step"]) == snap[name + ".step"], (tag, name)
        assert any(p is g["params"][0] for g in m.optimizer.param_groups), (tag, name)
        got, want = p.detach().cpu(), snap[name]
        if name in ("xyz", "scaling") and exact_xyz_rows is not None:
            # split children: R * sample + xyz through cuBLAS bmm / log(exp(s) / 1.6) on another device -> last-bit differences
            assert torch.equal(got[:exact_xyz_rows], want[:exact_xyz_rows]), (tag, name)
            torch.testing.assert_close(got[exact_xyz_rows:], want[exact_xyz_rows:], rtol=2e-6, atol=2e-7)
        else:
            assert torch.equal(got, want), (tag, name)
        assert torch.equal(st["exp_avg"].cpu(), snap[name + ".exp_avg"]) and torch.equal(st["exp_avg_sq"].cpu(), snap[name + ".exp_avg_sq"]), (tag, name)
    for a in _AUX:
        got = getattr(m, a).cpu()
        assert got.shape == snap[a].shape, (tag, a, tuple(got.shape), tuple(snap[a].shape))
        if got.dtype == torch.bool:
            assert torch.equal(got, snap[a]), (tag, a)
        else:
            torch.testing.assert_close(got.float(), snap[a].float(), rtol=1e-6, atol=1e-9, msg=f"{tag} {a}")
    assert n == m._xyz.shape[0]


def test_fused_densify_prune_stats_reset_match_the_real_gaussian_model(monkeypatch):
    from b200gs import densify
    g = torch.load(GOLD)
    # ---- add_densification_stats: replay the three views on the pre-stats state (accum = denom = 0) ----
    m = _load_model(g["before_densify"])
    m.xyz_gradient_accum.zero_(); m.denom.zero_()
    for vg, flt in g["stats_in"]:
        densify.add_densification_stats(m, vg.cuda(), flt.cuda())
    torch.testing.assert_close(m.xyz_gradient_accum.cpu(), g["before_densify"]["xyz_gradient_accum"], rtol=1e-6, atol=1e-12)
    assert torch.equal(m.denom.cpu(), g["before_densify"]["denom"])
    # ---- densify: same selection, same final order, same moments; the split offsets are the ones the reference drew ----
    m = _load_model(g["before_densify"])
    samples = [s.cuda() for s in g["normal_samples"]]
    calls = []

    def replay_normal(mean=None, std=None, **k):
        calls.append((tuple(mean.shape), tuple(std.shape)))
        return samples[len(calls) - 1]
    monkeypatch.setattr(torch, "normal", replay_normal)
    densify.densify(m, g["max_grad"], g["min_opacity"], g["extent"], None, 5, 5, None, 1, "fine")
    monkeypatch.undo()
    assert calls == [(tuple(samples[0].shape), tuple(samples[0].shape))]          # one torch.normal call on the reference's shapes
    n_children = samples[0].shape[0]
    _check(m, g["after_densify"], "densify", exact_xyz_rows=g["after_densify"]["xyz"].shape[0] - n_children)
    # ---- prune with the screen / world size criteria, then the opacity criterion alone ----
    m = _load_model(g["before_prune"])
    densify.prune(m, g["max_grad"], g["min_opacity"], g["extent"], 20)
    _check(m, g["after_prune"], "prune")
    densify.prune(m, g["max_grad"], 0.3, g["extent"], None)
    _check(m, g["after_prune_opacity_only"], "prune(opacity)")
    # ---- reset_opacity ----
    densify.reset_opacity(m)
    want = g["after_reset_opacity"]
    torch.testing.assert_close(m._opacity.detach().cpu(), want["opacity"], rtol=2e-6, atol=1e-6)
    st = m.optimizer.state[m._opacity]
    assert float(st["step"]) == want["opacity.step"] and float(st["exp_avg"].abs().max()) == 0.0 and float(st["exp_avg_sq"].abs().max()) == 0.0
    # the optimiser keeps working on the new tensors
    for grp in m.optimizer.param_groups:
        for q in grp["params"]:
            q.grad = torch.ones_like(q)
    m.optimizer.step()
