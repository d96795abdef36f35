"""GPU parity of the opt-in kernel variants (not the default path): the field / whole-step parity tests re-run in a child
process with the variants switched on (the library reads the switches once per process):
  B200GS_MLP_BWD_V2=1      deform_mlp_bwd_tc5_kernel<true>   (alternating weight slots, elected MMA issuer, coalesced flush)
  B200GS_MLP_FWD_ELECT=1   deform_mlp_fwd_tc5v2_kernel<64, true> (elected MMA issuer)
They were written without GPU access at the end of round 1, so this test only runs when B200GS_TEST_EXPERIMENTAL=1 is set;
once it has passed on a B200 the variants can become the default and this file folds into the ordinary parity tests."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {"mlp_bwd_v2": {"B200GS_MLP_BWD_V2": "1"}, "mlp_fwd_elect": {"B200GS_MLP_FWD_ELECT": "1"},
            "both": {"B200GS_MLP_BWD_V2": "1", "B200GS_MLP_FWD_ELECT": "1"}}


@pytest.mark.skipif(os.environ.get("B200GS_TEST_EXPERIMENTAL") != "1", reason="opt-in: set B200GS_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant_passes_the_field_and_step_parity_tests(name):
    env = dict(os.environ, **VARIANTS[name])
    env.pop("B200GS_TEST_EXPERIMENTAL", None)           # the child must not recurse into this file
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_field_parity.py",
                        "tests/test_hexplane_split_parity.py", "tests/test_train_step_parity.py"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
