"""GPU parity (-m gpu) on the MSA the reference's OWN pipeline produced (BASELINE.json configs[0]: DataSimulator ->
ReadCutter -> InitialAligner -> PW_ReAligner from the unmodified sources; tests/golden/real_pipeline_msareal.*, 312 reads x
18 557 columns, 7.7e7 pair tests): the same checks as every golden case of tests/test_gpu_parity.py - oracle parity for
every variant, pruning off, the general first-break path, and the host-finalised text byte-identical to the unmodified
MaxCorrelation.c's output.  The fixture was added after the round's last GPU call; the file sorts last so that a first
failure here cannot hide the other tests under `-x`."""
import pytest

from conftest import GOLDEN_CASES
from test_gpu_parity import LATE_CASES, VARIANTS, run_golden_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name,cov", [c for c in GOLDEN_CASES if c[0] in LATE_CASES])
def test_golden_case_from_the_reference_pipeline(name, cov, variant, tmp_path):
    run_golden_case(name, cov, variant, tmp_path)
