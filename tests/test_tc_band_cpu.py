"""CPU check of the tensor-core graph filter's error band (constants as in prep_kernel, graph_build.cu):
the FP32 evaluation of d = t^2 - 2 beta^2 s + beta^4 and band = kap (t^2 + beta^4) + c0 is emulated in numpy on squared
norms perturbed by the worst-case tensor-core error; every DECIDED pair must agree with the exact FP64 predicate of the
reference (registration.cc:427-443).  The GPU counterpart is debug flag 2 of tests/test_gpu_round2.py."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
synth = importlib.import_module("teaser-plusplus_b200.synth")
band = importlib.import_module("tc_band_check")


@pytest.mark.parametrize("cfg,n", [("C2", 700), ("C2cube", 500), ("C3", 800), ("C5", 600)])
@pytest.mark.parametrize("mode", ["rand", "pp", "pm", "mm"])
def test_decided_pairs_agree_with_the_exact_predicate(cfg, n, mode):
    pr = synth.config_problem(cfg, 21, n=n)
    wrong, undecided, _ = band.check(pr["src"], pr["dst"], pr["noise_bound"], mode, cfg, verbose=False)
    assert wrong == 0
    assert undecided < 2e-3  # the exact path is the exception (3.5e-5 on C2, 6e-4 on the 15 %-dense graphs)


def test_short_tims_and_duplicates_never_decide_wrongly():
    """Pairs with s = a + b <= beta^2 (both TIMs shorter than beta; duplicate correspondences) are where the sign of the
    polynomial says nothing: the beta^4 floor of the band must keep them out of the 'surely not an edge' class."""
    pr = synth.config_problem("C2cube", 8, n=500)
    src, dst = pr["src"].copy(), pr["dst"].copy()
    for k in range(0, 60, 3):
        src[k + 1] = src[k]
    for k in range(100, 160, 3):
        dst[k + 1] = dst[k]
        src[k + 1] = src[k] + 1e-9
    for mode in ("rand", "pp", "pm"):
        assert band.check(src, dst, pr["noise_bound"], mode, "dups", verbose=False)[0] == 0
    # noise bounds up to a quarter of the extent (far beyond the use_tc guard): still never wrong, only more undecided
    pr = synth.config_problem("C2", 3, n=400)
    for nb in (0.05, 0.2, 0.5):
        wrong, _, use_tc = band.check(pr["src"], pr["dst"], nb, "pm", "big", verbose=False)
        assert wrong == 0
    assert not use_tc  # beta = 1.0 on a unit cube: prep_kernel would route this problem to the CUDA-core kernel
