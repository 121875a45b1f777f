"""The oracle against its committed golden digests (tests/golden/oracle_golden.json, made by make_golden.py), and --
on a GPU -- the CUDA path against the same digests."""
import json
import os

import numpy as np
import pytest

from golden import make_golden as G

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_golden.json")))


def test_oracle_matches_golden():
    assert G.compute() == GOLD


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
