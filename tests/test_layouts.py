"""Record layouts: numpy dtypes == reference dtypes (golden) == C header sizes.  (reference: src/struct_types.py)"""
import json
import os
import re

from clive2_b200 import struct_types as T
from tests.common import GOLDEN

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_offsets_match_reference_dtypes():
    golden = json.load(open(os.path.join(GOLDEN, "layouts.json")))
    for name, g in golden.items():
        dt = getattr(T, name)
        assert dt.itemsize == g["itemsize"], name
        assert {k: v[1] for k, v in dt.fields.items()} == g["offsets"], name


def test_weight_aggregator_layout():
    wa = T.WeightAggregator
    assert wa.itemsize == 80
    assert wa.fields["weights"][1] == 0 and wa.fields["total_contribution"][1] == 48
    assert wa.fields["contrib_weight_sum"][1] == 64


def test_header_declares_same_sizes():
    text = open(os.path.join(REPO, "include", "clive_b200.h")).read()
    for name, size in (("Ray", 128), ("Path", 1040), ("Box", 48), ("Triangle", 128), ("Material", 48),
                       ("Camera", 112), ("WeightAggregator", 80)):
        m = re.search(r"typedef struct clv_%s \{ /\*[^*]*?(\d+) bytes" % name, text)
        assert m and int(m.group(1)) == size, name
