"""A bounded slice of the randomised differential campaign (scripts/fuzz_parity.py): random systems, frame
counts, boxes, mapping centres, tuning knobs and unaligned device slices, each checked against the oracle."""

import importlib.util
import pathlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_random_cases_against_the_oracle():
    path = pathlib.Path(__file__).resolve().parent.parent / "scripts" / "fuzz_parity.py"
    spec = importlib.util.spec_from_file_location("fuzz_parity", path)
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    rng = np.random.default_rng(20241019)
    kinds = set()
    for case in range(60):
        kind, n_frames, box_mode, n_terms = fuzz.check_case(rng, case)      # raises AssertionError on a mismatch
        kinds.add((str(kind), str(box_mode)))
    assert len(kinds) >= 8          # membranes, solvated membranes and chains x large / tight / triclinic / no box
