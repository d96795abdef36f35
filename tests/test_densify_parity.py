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
