"""GPU parity of the NON-default kernel variants (include/b200gs.h: b200gs_set_option): the field / split / whole-step parity
tests re-run in a child process with the first-generation kernels, with the minimal second-generation variants, and with the
round-1 defaults, selected through the environment.  The defaults (mlp_bwd_v2 87, mlp_fwd_elect 2, hexplane_time_* 2,
lookback_parallel 1, sort_ballot_rank 1) are what every other GPU test runs; they were validated against the first-generation kernels by
tools/native/{mlp_variant_check,hexplane_time_check,sort_check,rast_check} (profiles/r2a_*.txt)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {"first_generation": {"B200GS_MLP_BWD_V2": "0", "B200GS_MLP_FWD_ELECT": "0"},
            "minimal": {"B200GS_MLP_BWD_V2": "1", "B200GS_MLP_FWD_ELECT": "1"},
            "single_dy_without_tma": {"B200GS_MLP_BWD_V2": "55"},
            "round1_defaults": {"B200GS_MLP_BWD_V2": "7", "B200GS_MLP_FWD_ELECT": "2", "B200GS_HEXPLANE_TIME_FWD": "0",
                                "B200GS_HEXPLANE_TIME_BWD": "0", "B200GS_LOOKBACK_PARALLEL": "0", "B200GS_COMPOSITE_PAIRS": "0",
                                "B200GS_SORT_BALLOT_RANK": "0"}}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant_passes_the_field_and_step_parity_tests(name):
    env = dict(os.environ, **VARIANTS[name])
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_field_parity.py",
                        "tests/test_hexplane_split_parity.py", "tests/test_train_step_parity.py", "tests/test_raster_parity.py",
                        "tests/test_raster_parity_configs.py", "-k", "not 1000000 and not 5000000"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
