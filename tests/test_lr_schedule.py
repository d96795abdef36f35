"""CPU: engine.expon_lr / GaussianState.update_learning_rate against the reference's get_expon_lr_func
(utils/general_utils.py:35-68) and update_learning_rate (scene/gaussian_model.py:284-298)."""
import importlib.util
import math
import os

import pytest

REF = "/root/reference"


def test_expon_lr_known_values():
    from b200gs.engine import expon_lr
    assert expon_lr(0, 1.6e-4, 1.6e-6, 20000) == pytest.approx(1.6e-4, rel=1e-12)
    assert expon_lr(20000, 1.6e-4, 1.6e-6, 20000) == pytest.approx(1.6e-6, rel=1e-12)
    assert expon_lr(10000, 1.6e-4, 1.6e-6, 20000) == pytest.approx(1.6e-5, rel=1e-12)        # geometric mean at the midpoint
    assert expon_lr(50000, 1.6e-4, 1.6e-6, 20000) == pytest.approx(1.6e-6, rel=1e-12)        # clipped past max_steps
    assert expon_lr(-1, 1.6e-4, 1.6e-6, 20000) == 0.0 and expon_lr(5, 0.0, 0.0, 100) == 0.0
    assert expon_lr(0, 1e-3, 1e-5, 1000, lr_delay_steps=100, lr_delay_mult=0.01) == pytest.approx(1e-5, rel=1e-12)


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "utils", "general_utils.py")), reason="reference tree not present")
def test_expon_lr_matches_reference_function():
    from b200gs.engine import expon_lr
    spec = importlib.util.spec_from_file_location("ref_general_utils", os.path.join(REF, "utils", "general_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for args in [(1.6e-4, 1.6e-6, 0, 1.0, 20000), (1.6e-3, 1.6e-5, 0, 0.01, 20000), (1e-3, 1e-5, 500, 0.01, 3000)]:
        lr_init, lr_final, delay_steps, delay_mult, max_steps = args
        f = mod.get_expon_lr_func(lr_init, lr_final, lr_delay_steps=delay_steps, lr_delay_mult=delay_mult, max_steps=max_steps)
        for step in (0, 1, 17, 499, 500, 2999, 3000, 19999, 20000, 40000):
            assert expon_lr(step, lr_init, lr_final, max_steps, delay_steps, delay_mult) == pytest.approx(float(f(step)), rel=1e-12)


def test_update_learning_rate_touches_the_scheduled_groups_only():
    import torch
    from b200gs import engine, synthetic as syn

    class _Opt:            # stands in for FusedAdam (which needs CUDA): update_learning_rate only touches param_groups
        def __init__(self, groups):
            self.param_groups = [dict(g) for g in groups]
    torch.manual_seed(0)
    st = engine.GaussianState.__new__(engine.GaussianState)
    o = engine.default_opt()
    names = ["xyz", "deformation", "grid", "f_dc", "f_rest", "opacity", "scaling", "rotation"]
    st.optimizer = _Opt([{"name": n, "lr": 1.0, "params": []} for n in names])
    st._lr_args = {"xyz": (o.position_lr_init, o.position_lr_final), "deformation": (o.deformation_lr_init, o.deformation_lr_final),
                   "grid": (o.grid_lr_init, o.grid_lr_final)}
    st._lr_max_steps = o.position_lr_max_steps
    engine.GaussianState.update_learning_rate(st, 10000)
    lr = {g["name"]: g["lr"] for g in st.optimizer.param_groups}
    assert lr["xyz"] == pytest.approx(math.sqrt(o.position_lr_init * o.position_lr_final), rel=1e-9)
    assert lr["grid"] == pytest.approx(math.sqrt(o.grid_lr_init * o.grid_lr_final), rel=1e-9)
    assert lr["deformation"] == pytest.approx(math.sqrt(o.deformation_lr_init * o.deformation_lr_final), rel=1e-9)
    assert all(lr[n] == 1.0 for n in ("f_dc", "f_rest", "opacity", "scaling", "rotation"))
