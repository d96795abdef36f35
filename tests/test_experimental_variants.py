"""GPU parity of the NON-default deformation-MLP kernel variants (include/b200gs.h: b200gs_set_option): the field / split /
whole-step parity tests re-run in a child process with the first-generation kernels (0 / 0) and with the minimal variants
(1 / 1) selected through the environment.  The defaults (7 / 2) are what every other GPU test runs; they were validated
against the first-generation kernels bit for bit by tools/native/mlp_variant_check (profiles/r1l_mlp_variant_check_b200.txt).
Three extra pytest sessions take a few minutes, so this only runs when B200GS_TEST_EXPERIMENTAL=1 is set."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {"first_generation": {"B200GS_MLP_BWD_V2": "0", "B200GS_MLP_FWD_ELECT": "0"},
            "minimal": {"B200GS_MLP_BWD_V2": "1", "B200GS_MLP_FWD_ELECT": "1"},
            "split_dfeature_store": {"B200GS_MLP_BWD_V2": "15", "B200GS_MLP_FWD_ELECT": "2"}}


@pytest.mark.skipif(os.environ.get("B200GS_TEST_EXPERIMENTAL") != "1", reason="opt-in: set B200GS_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant_passes_the_field_and_step_parity_tests(name):
    env = dict(os.environ, **VARIANTS[name])
    env.pop("B200GS_TEST_EXPERIMENTAL", None)           # the child must not recurse into this file
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_field_parity.py",
                        "tests/test_hexplane_split_parity.py", "tests/test_train_step_parity.py"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
