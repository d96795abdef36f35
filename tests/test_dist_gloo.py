"""CPU, world_size 2 over gloo: the view-parallel step (sharding of the view batch, flat
gradient arena, one SUM all-reduce + one MAX all-reduce, identical optimiser step on every rank)
gives the same parameters as the single-process run over the whole batch.  The CUDA kernels are
replaced by a small differentiable stand-in for `render`, the optimiser by torch.optim.Adam, so
only the host-side multi-GPU logic is under test here (the kernels have their own GPU tests)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class TinyModel(nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self._xyz = nn.Parameter(torch.randn(50, 3, generator=g))
        self._opacity = nn.Parameter(torch.randn(50, 1, generator=g))
        plane = torch.rand(1, 4, 5, 6, generator=g).contiguous(memory_format=torch.channels_last)
        self._deformation = nn.ParameterDict({"plane": nn.Parameter(plane), "timenet_w": nn.Parameter(torch.randn(3, 3, generator=g))})
        self.optimizer = None

    @property
    def get_xyz(self):
        return self._xyz


def fake_render(cam, model, bg, stage):
    """Differentiable stand-in: an 'image' that depends on xyz, opacity and the plane."""
    sp = torch.zeros_like(model._xyz, requires_grad=True)
    w = torch.sigmoid(model._opacity) * (model._xyz + sp).mul(cam).sum(1, keepdim=True)        # [50,1]
    img = (w.sum() * model._deformation["plane"].mean(dim=(0, 1))).unsqueeze(0).expand(3, 5, 6) + bg[:, None, None]
    radii = (model._xyz[:, 0].detach() * cam * 10).to(torch.int32).clamp_min(0)
    return {"render": img, "viewspace_points": sp, "radii": radii}


def _run(rank, world, port, ret):
    sys.path.insert(0, os.path.join(ROOT, "iclr2025_3d-mom_b200"))
    from b200gs.engine import ViewParallelTrainer
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    model = TinyModel()
    model.optimizer = torch.optim.Adam([{"params": [model._xyz], "lr": 1e-2}, {"params": [model._opacity], "lr": 5e-2},
                                        {"params": list(model._deformation.parameters()), "lr": 1e-3}], lr=0.0, eps=1e-15)
    tr = ViewParallelTrainer(model, torch.tensor([0.1, 0.2, 0.3]), stage="fine", world_size=world, rank=rank,
                             render_fn=fake_render)
    cams = [0.5 + 0.25 * b for b in range(4)]
    g = torch.Generator().manual_seed(1)
    gts = [torch.rand(3, 5, 6, generator=g) for _ in range(4)]
    for _ in range(3):
        mine = tr.local_views(4)
        tr.step([cams[i] for i in mine], [gts[i] for i in mine], global_batch=4)
    out = {k: v.detach().clone() for k, v in model.state_dict().items()}
    out["viewspace"] = tr.viewspace_grad.clone(); out["max_radii"] = tr.max_radii.clone()
    out["timenet_grad_is_none"] = model._deformation["timenet_w"].grad is None
    # utils/image_utils.py:psnr of the LAST step's local views, as train_4DGS.py:212 logs it (evaluated before that step's update)
    out["psnr_local"] = tr.psnr().detach().clone()
    out["psnr_views"] = torch.tensor(mine)
    ret[rank if world > 1 else -1] = out
    if world > 1:
        dist.destroy_process_group()


def test_two_rank_step_equals_single_process():
    mgr = mp.Manager()
    ret = mgr.dict()
    _run(0, 1, 0, ret)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_run, args=(2, port, ret), nprocs=2, join=True)
    single, r0, r1 = ret[-1], ret[0], ret[1]
    assert single["timenet_grad_is_none"] and r0["timenet_grad_is_none"]
    # the per-rank PSNR means combine to the single-process mean over the batch (2 views each)
    assert torch.allclose((r0["psnr_local"] + r1["psnr_local"]) / 2, single["psnr_local"], rtol=1e-5)
    assert r0["psnr_views"].tolist() == [0, 2] and r1["psnr_views"].tolist() == [1, 3]
    for k in single:
        if k in ("timenet_grad_is_none", "psnr_local", "psnr_views"):
            continue
        assert torch.equal(r0[k], r1[k]), f"ranks diverged on {k}"                  # replicas stay identical
        assert torch.allclose(single[k].float(), r0[k].float(), rtol=1e-5, atol=1e-7), k


class TinyModelSH(TinyModel):
    """+ the two SH parameters the trainer concatenates once per step when it shares the SH tensor between views."""
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self._features_dc = nn.Parameter(torch.randn(50, 1, 3, generator=g))
        self._features_rest = nn.Parameter(torch.randn(50, 15, 3, generator=g) * 0.1)


def fake_render_shs(cam, model, bg, stage, shs=None):
    """As fake_render, with a colour term that depends on the step's shared SH tensor."""
    pkg = fake_render(cam, model, bg, stage)
    feats = shs if shs is not None else torch.cat((model._features_dc, model._features_rest), dim=1)
    colour = (feats * torch.linspace(0.5, 1.5, 16)[None, :, None]).sum(1).mean(0) * cam          # [3]
    pkg["render"] = pkg["render"] + colour[:, None, None]
    return pkg


def _run_sh(rank, world, port, ret, overlap):
    sys.path.insert(0, os.path.join(ROOT, "iclr2025_3d-mom_b200"))
    from b200gs.engine import ViewParallelTrainer
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    model = TinyModelSH()
    model.optimizer = torch.optim.Adam([{"params": [model._xyz], "lr": 1e-2}, {"params": [model._opacity], "lr": 5e-2},
                                        {"params": [model._features_dc], "lr": 2.5e-3}, {"params": [model._features_rest], "lr": 1.25e-4},
                                        {"params": list(model._deformation.parameters()), "lr": 1e-3}], lr=0.0, eps=1e-15)
    tr = ViewParallelTrainer(model, torch.tensor([0.1, 0.2, 0.3]), stage="fine", world_size=world, rank=rank,
                             render_fn=fake_render_shs, shared_shs=True, overlap_sh_reduce=overlap)
    # shared SH tensor: the two SH parameters live outside the flat arena, their gradient is ONE [P,16,3] buffer reduced by its
    # own collective (the "SH tail"; on CUDA it runs on a side stream from inside the last view's backward, on CPU after the loop)
    assert not tr.overlap_sh_reduce                      # CPU: no side stream
    assert len(tr.sh_params) == 2 and all(not any(p is q for q in tr.arena_params) for p in tr.sh_params)
    assert sum(p.numel() for p in tr.arena_params) + 3 * 50 <= tr.arena.numel()
    cams = [0.5 + 0.25 * b for b in range(4)]
    g = torch.Generator().manual_seed(1)
    gts = [torch.rand(3, 5, 6, generator=g) for _ in range(4)]
    for _ in range(3):
        mine = tr.local_views(4)
        tr.step([cams[i] for i in mine], [gts[i] for i in mine], global_batch=4)
    out = {k: v.detach().clone() for k, v in model.state_dict().items()}
    out["viewspace"] = tr.viewspace_grad.clone(); out["max_radii"] = tr.max_radii.clone()
    ret[(rank if world > 1 else -1, overlap)] = out
    if world > 1:
        dist.destroy_process_group()


def test_overlapped_sh_reduce_equals_single_process():
    """The SH gradient reduced by its own all-reduce and the rest of the arena by the post-loop one give the same parameters
    as the single-process run, whatever the overlap flag asks for."""
    mgr = mp.Manager()
    ret = mgr.dict()
    _run_sh(0, 1, 0, ret, False)
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_run_sh, args=(2, port, ret, True), nprocs=2, join=True)
    mp.spawn(_run_sh, args=(2, port + 1, ret, False), nprocs=2, join=True)
    single, o0, o1, p0 = ret[(-1, False)], ret[(0, True)], ret[(1, True)], ret[(0, False)]
    for k in single:
        assert torch.equal(o0[k], o1[k]), f"ranks diverged on {k}"
        assert torch.allclose(single[k].float(), o0[k].float(), rtol=1e-5, atol=1e-7), k
        assert torch.allclose(p0[k].float(), o0[k].float(), rtol=1e-6, atol=1e-8), k
