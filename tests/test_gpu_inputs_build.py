"""CPU-collected guard for the `-m gpu` tests: every input they use can be generated here (no GPU), and no GPU test module
builds inputs outside the registry (tests/gpu_cases.py)."""
import glob
import os
import re

import numpy as np
import pytest

import gpu_cases as GC

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", sorted(GC.SYNTHETIC))
def test_synthetic_case_builds(name):
    kw = GC.SYNTHETIC[name]
    D = GC.synthetic(name)
    assert D["Y"].shape == (kw["d1"], kw["d2"], kw["T"]) and D["Y"].dtype == np.uint16
    assert D["A0"].shape == (kw["d1"] * kw["d2"], kw["K"]) and D["C0"].shape == (kw["K"], kw["T"])
    assert D["IND"].shape == D["A0"].shape and D["IND"].nnz > 0
    assert (np.asarray(D["A0"].sum(axis=0)).ravel() > 0).all(), "a neuron of %s has an empty initial footprint" % name


@pytest.mark.parametrize("name", sorted(n for n in GC.TRACES if GC.TRACES[n][2] <= 20011))
def test_trace_case_builds(name):
    Y, truth, spikes = GC.traces(name)
    g, noise, T, fr, rate, b, N, seed = GC.TRACES[name]
    assert Y.shape == (N, T) and np.isfinite(Y).all() and spikes.sum() > 0


def test_gpu_tests_only_use_registered_inputs():
    bad = []
    for path in sorted(glob.glob(os.path.join(HERE, "test_gpu_*.py"))):
        if os.path.basename(path) == "test_gpu_inputs_build.py":
            continue
        src = open(path).read()
        if re.search(r"make_synthetic\(|gen_data\(", src):
            bad.append(os.path.basename(path))
        for m in re.finditer(r"GC\.synthetic\(\"([^\"]+)\"\)", src):
            assert m.group(1) in GC.SYNTHETIC, "%s uses unregistered case %s" % (path, m.group(1))
        for m in re.finditer(r"_traces\(\"([^\"%]+)\"\)|GC\.traces\(\"([^\"]+)\"\)", src):
            nm = m.group(1) or m.group(2)
            assert nm in GC.TRACES, "%s uses unregistered traces %s" % (path, nm)
    assert not bad, "GPU test modules generating inputs outside tests/gpu_cases.py: %s" % bad
