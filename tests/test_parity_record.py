"""The full-size parity record (tests/parity_record.py) at reduced sizes, so that every `pytest -m gpu` run re-measures what
profiles/r02_parity.json records at the sizes of BASELINE.json's configs, and prints the figures (terminal summary)."""
import json
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def test_parity_record_reduced(lb, figure):
    import parity_record as pr
    rec = pr.run(clips=300, chunk=100, sweep_clips=150, sweep_queries=40, db_clips=300, db_queries=24, log=lambda *_: None)
    for name, part in rec.items():
        if not isinstance(part, dict) or name == "tolerances":
            continue
        figure("%s: %s" % (name, json.dumps(part)))
    c2 = [v for k, v in rec.items() if k.startswith("config2")][0]
    assert c2["booleans"]["mismatch_rate"] <= 1e-3
    assert c2["band_energies"]["elements_beyond_test_tolerance (1e-4 rel + 1e-9 of image max)"] == 0
    assert c2["band_energies"]["normwise_error"] < 1e-6
    assert c2["haar"]["max_error_relative_to_image_max_coefficient"] < 1e-4
    c1 = [v for k, v in rec.items() if k.startswith("config1")][0]
    assert c1["score"]["abs_diff_on_identical_bits"] == 0.0 and c1["booleans"]["mismatch_rate"] <= 1e-3
    c4 = [v for k, v in rec.items() if k.startswith("config4 shape (")][0]
    assert c4["scores_bit_identical"] and c4["top10_indices_identical"] and c4["top10_scores_identical"]
    c5 = [v for k, v in rec.items() if k.startswith("config5")][0]
    for k, v in c5.items():
        if "stages" in k:
            assert v["haar"]["max_error_relative_to_image_max_coefficient"] < 1e-4, k
        else:
            assert v["mismatch_rate"] <= 2e-3 and v["score_max_abs_diff_identical_bits"] == 0.0, k
