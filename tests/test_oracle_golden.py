"""The oracle reproduces every golden fixture the real reference generated (tests/golden/).
Runs anywhere (no reference tree, no GPU needed)."""
import os
import sys

import pytest

import helpers
from oracle import shannon_oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_matches_golden(workdir, name):
    helpers.check_against_golden(shannon_oracle.extension_correction,
                                 shannon_oracle.kmers_for_component, name, workdir, "oracle")


def test_chunked_flush_equals_single_chunk(workdir):
    """NR (the reference's 10M-read flush interval, kmers_for_component.py:322) only changes
    when files are appended, never their content."""
    s1, s2 = helpers.synthetic_seqs(12, 1500, 1)
    case = helpers.make_case(workdir, 24, s1, s2)
    a = helpers.run_frontend(shannon_oracle.extension_correction,
                             shannon_oracle.kmers_for_component, case, "a")
    b = helpers.run_frontend(shannon_oracle.extension_correction,
                             shannon_oracle.kmers_for_component, case, "b",
                             extra_kfc={"NR": 97})
    helpers.assert_same_run(a, b, "NR")
