"""The oracle against its committed golden digests (tests/golden/oracle_golden.json, made by make_golden.py), against the
digests of the outputs of THE REFERENCE ITSELF (tests/golden/reference_golden.json, made by make_reference_golden.py from
oracle/_ref/libsuma_ref_full.so where /root/reference exists), and -- on a GPU -- the CUDA path against the oracle's."""
import json
import os

import numpy as np
import pytest

from golden import make_golden as G

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_golden.json")))


def test_oracle_matches_golden():
    assert G.compute() == GOLD


def test_reference_generated_golden_is_what_the_reference_produces():
    """the committed digests are reproduced by running the reference itself (oracle/_ref/libsuma_ref_full.so) again"""
    from golden import make_reference_golden as RG
    from oracle import ref as R
    if not R.full_available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    ref = json.load(open(os.path.join(HERE, "golden", "reference_golden.json")))
    assert RG.compute(RG.ReferenceEngine) == ref


def test_oracle_matches_the_reference_generated_golden():
    """preprocessing, map update / rendering through a paging tour, and whole processScan runs (48 ICP values added the GL
    way): the oracle reproduces the digests the reference's own classes and shaders produced in the build container"""
    from golden import make_reference_golden as RG
    from oracle import oracle as O
    ref = json.load(open(os.path.join(HERE, "golden", "reference_golden.json")))
    old = O.gl_sums(1)
    try:
        got = RG.compute(RG.OracleEngine)
    finally:
        O.gl_sums(old)
    assert sorted(got) == sorted(ref)
    for k in ref:
        assert got[k] == ref[k], k


@pytest.mark.gpu
def test_cuda_matches_golden():
    from semantic_suma_b200 import api
    from helpers import both_params, scans, sized
    for semantic in (False, True):
        tag = "semantic" if semantic else "geometric"
        po, pp = both_params(**sized(900))
        sc, _ = scans(900, n=4, semantic=semantic)
        sl = api.SurfelMapping(pp)
        for s in sc:
            sl.processScan(*s)
        g = GOLD["slam_%s" % tag]
        assert G.digest(sl.getCurrentPose()) == g["pose"]
        assert sl.getMap().size() == g["surfels"]
        assert G.surfel_digest(sl.getMap().getAllSurfels()) == g["surfel_digest"]
        assert [G.digest(x) for x in sl.getLastModelFrame().maps()] == g["frame"]
        sl.ctx.close()

