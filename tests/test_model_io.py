"""CPU: b200gs.modelio against the reference's own GaussianModel.save_ply / load_ply (scene/gaussian_model.py:342-360, 367-407),
run in the build container by oracle/gen_golden_model_io.py -> tests/golden/model_io.{ply,pt}: the file we write is byte-identical
to the reference's, what we read from the reference's file equals what the reference reads, and save_deformation / load_model
round-trip the field's state_dict under the reference's file names."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _model(inputs):
    from b200gs import engine
    raw = {"xyz": inputs["xyz"], "shs": torch.cat((inputs["f_dc"], inputs["f_rest"]), dim=1), "log_scale": inputs["scaling"],
           "rot": inputs["rotation"], "opacity_logit": inputs["opacity"], "scene_flow": torch.zeros_like(inputs["xyz"])}
    return engine.GaussianState(raw)


def test_save_ply_is_byte_identical_to_the_reference(tmp_path):
    gold = torch.load(os.path.join(GOLD, "model_io.pt"), weights_only=True)
    m = _model(gold["inputs"])
    out = tmp_path / "sub" / "point_cloud.ply"                     # save_ply creates the directory (mkdir_p, :343)
    m.save_ply(str(out))
    with open(out, "rb") as a, open(os.path.join(GOLD, "model_io.ply"), "rb") as b:
        assert a.read() == b.read()


def test_load_ply_equals_the_reference(tmp_path):
    gold = torch.load(os.path.join(GOLD, "model_io.pt"), weights_only=True)
    zeros = {k: torch.zeros_like(v) for k, v in gold["inputs"].items()}
    m = _model(zeros)
    m.active_sh_degree = 0
    m.load_ply(os.path.join(GOLD, "model_io.ply"))
    got = {"xyz": m._xyz, "f_dc": m._features_dc, "f_rest": m._features_rest, "opacity": m._opacity, "scaling": m._scaling,
           "rotation": m._rotation}
    for k, want in gold["loaded"].items():
        assert got[k].shape == want.shape and got[k].is_contiguous() and got[k].requires_grad, k
        assert torch.equal(got[k].detach(), want), k
        assert torch.equal(got[k].detach(), gold["inputs"][k]), k          # f4 on disk: the round trip is exact
    assert m.active_sh_degree == gold["active_sh_degree"] == 3
    assert m.get_features.shape == (11, 16, 3)


def test_load_ply_rejects_bad_files(tmp_path):
    from b200gs import modelio
    gold = torch.load(os.path.join(GOLD, "model_io.pt"), weights_only=True)
    m = _model(gold["inputs"])
    m.max_sh_degree = 2                                             # the file holds degree-3 coefficients (assert at :384)
    with pytest.raises(ValueError, match="f_rest"):
        m.load_ply(os.path.join(GOLD, "model_io.ply"))
    raw = open(os.path.join(GOLD, "model_io.ply"), "rb").read()
    cut = tmp_path / "cut.ply"
    cut.write_bytes(raw[:-5])
    with pytest.raises(ValueError, match="truncated"):
        modelio.read_ply_vertices(str(cut))
    asc = tmp_path / "ascii.ply"
    asc.write_bytes(raw.replace(b"binary_little_endian", b"ascii"))
    with pytest.raises(ValueError, match="little-endian"):
        modelio.read_ply_vertices(str(asc))


def test_ragged_and_empty_point_clouds(tmp_path):
    from b200gs import engine
    for P in (0, 1):
        raw = {"xyz": torch.randn(P, 3), "shs": torch.randn(P, 16, 3), "log_scale": torch.randn(P, 3), "rot": torch.randn(P, 4),
               "opacity_logit": torch.randn(P, 1), "scene_flow": torch.zeros(P, 3)}
        a = engine.GaussianState(raw)
        a.save_ply(str(tmp_path / f"p{P}.ply"))
        b = engine.GaussianState({k: torch.zeros_like(v) for k, v in raw.items()})
        b.load_ply(str(tmp_path / f"p{P}.ply"))
        for n in ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation"):
            assert torch.equal(getattr(a, n).detach(), getattr(b, n).detach()), (P, n)


def test_deformation_checkpoint_round_trip(tmp_path):
    gold = torch.load(os.path.join(GOLD, "model_io.pt"), weights_only=True)
    torch.manual_seed(1)
    a = _model(gold["inputs"])
    with torch.no_grad():
        for p in a._deformation.parameters():
            if p.requires_grad:
                p.add_(torch.randn_like(p) * 0.01)
        a._scene_flow.copy_(torch.randn_like(a._scene_flow))
    a.densification_setup()
    a._deformation_accum += 1.5
    a.save_deformation(str(tmp_path / "ckpt"))
    assert sorted(os.listdir(tmp_path / "ckpt")) == ["deformation.pth", "deformation_accum.pth", "deformation_table.pth", "scene_flow.pth"]
    torch.manual_seed(2)
    b = _model(gold["inputs"])
    b.load_model(str(tmp_path / "ckpt"))
    sa, sb = a._deformation.state_dict(), b._deformation.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert torch.equal(a._scene_flow, b._scene_flow)
    assert torch.equal(a._deformation_accum, b._deformation_accum) and torch.equal(a._deformation_table, b._deformation_table)
    assert b.max_radii2D.shape == (11,) and float(b.max_radii2D.abs().max()) == 0.0
