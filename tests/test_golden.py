"""Committed golden fixtures (tests/golden/): the oracle must keep reproducing them (CPU), the CUDA engine must match them bit
for bit (GPU)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_binding as ob
import sabc_b200 as sb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)
import make_trajectory_fixtures as mtf  # noqa: E402

FIXTURES = json.load(open(os.path.join(GOLDEN, "trajectories.json")))


def test_known_answers_are_reproducible():
    r = subprocess.run([sys.executable, os.path.join(GOLDEN, "make_known_answers.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("fx", FIXTURES, ids=lambda f: "-".join(str(x) for x in f["case"]))
def test_oracle_reproduces_trajectory_fixture(fx):
    assert mtf.run_case(ob.OracleEngine, *fx["case"]) == fx


@pytest.mark.gpu
@pytest.mark.parametrize("fx", FIXTURES, ids=lambda f: "-".join(str(x) for x in f["case"]))
def test_engine_matches_trajectory_fixture(gpu, fx):
    got = mtf.run_case(sb.Engine, *fx["case"])
    for key in fx:
        assert got[key] == fx[key], key
