"""Every MDC_NO_* fallback switch (the previous form of one optimisation each) runs the same parity
checks as the default build.  The switches are read once per process, so each combination runs in a subprocess
(tests/switch_case.py: UNet / decoder tapes forward + backward against the fp32 oracle, and a 5-step pipeline call)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

CASES = {
    "two_pass_groupnorm__unfused_attention__single_cta": ["MDC_NO_GNFUSE", "MDC_NO_FLASH", "MDC_NO_PAIR"],
    "no_splitk__materialised_upsample__plain_stores__no_alias": ["MDC_NO_SPLITK", "MDC_NO_UPCONV", "MDC_NO_TMASTORE", "MDC_NO_ALIAS"],
    "no_graph__no_pdl__no_side_stream": ["MDC_NO_GRAPH", "MDC_NO_PDL", "MDC_NO_SIDE"],
    "unfused_cross_attention__tap_by_tap_conv__epilogue_groupnorm_stats": ["MDC_NO_XFUSE", "MDC_NO_ROWSHARE", "MDC_GNEPI"],
}


@pytest.mark.parametrize("name", list(CASES))
def test_fallback_switches(cuda, name):
    env = dict(os.environ)
    for k in CASES[name]:
        env[k] = "1"
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "switch_case.py")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SWITCH_CASE_OK" in r.stdout, f"{name}: rc {r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}"
