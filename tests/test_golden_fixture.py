"""The frozen fixture tests/golden/hotpath_cases.json.gz (BASELINE layouts, 300 read groups each):
CPU tier -- the oracle and the host run of the device functions still reproduce it;
GPU tier -- the CUDA path reproduces it through the C ABI."""
import numpy as np
import pytest

from precellar_b200.pipeline import assemble, render
from tests import util as U
from tests.golden_util import load

CASES = ["10x_rna_v3", "10x_atac", "10x_multiome_atac", "sci_rna_seq3"]


@pytest.fixture(scope="module")
def golden():
    return load()


def check(layout, infos, full, results, counts, c):
    g = assemble(layout, infos, full, results, correct=True)
    for i, want in enumerate(c["lines"]):
        assert render(layout, full, g, i, correct=True) == want, i
    for gi, (cnt, mm, tot) in c["counts"].items():
        np.testing.assert_array_equal(np.asarray(counts[gi][0], np.uint64), cnt)
        assert (int(counts[gi][1]), int(counts[gi][2])) == (mm, tot)


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_fixture(golden, name):
    c = golden[name]
    ores, ocounts = U.run_oracle(c["layout"], c["full"], correct=True, threshold=0.975)
    assert ores.lines == c["lines"]
    np.testing.assert_array_equal(ores.qc_cat, c["qc_cat"])
    np.testing.assert_array_equal(ores.qc_per_file, c["qc_per_file"])
    for gi, (cnt, mm, tot) in c["counts"].items():
        np.testing.assert_array_equal(ocounts[gi][0], cnt)


@pytest.mark.parametrize("name", CASES)
def test_device_functions_reproduce_fixture(golden, name):
    c = golden[name]
    infos, strides = U.emul_infos(c["layout"])
    infos, results, counts, _ = U.run_emul(c["layout"], U.narrow(c["full"], strides, infos), correct=True, threshold=0.975)
    check(c["layout"], infos, c["full"], results, counts, c)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_reproduces_fixture(golden, name):
    from tests.test_gpu_parity import run_gpu
    c = golden[name]
    infos, strides = U.emul_infos(c["layout"])
    ginfos, results, counts, defects, _ = run_gpu(c["layout"], U.narrow(c["full"], strides, infos), correct=True, threshold=0.975)
    check(c["layout"], ginfos, c["full"], results, counts, c)
    np.testing.assert_array_equal(defects, c["qc_per_file"][:, 1])
