"""Tile producer host logic against the reference's own splitter
(DOTA_devkit/SplitOnlyImage_multi_process.py:51-87, run by tests/golden/gen_golden_split.py)."""
import json
import os

import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "split_tiles.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: "%dx%d" % (c["w"], c["h"]))
def test_tile_names_and_order_match_reference(case):
    from orientedreppoints_b200.dota.split_tiles import tile_names, tile_origins
    org = tile_origins(case["w"], case["h"], GOLD["subsize"], GOLD["gap"])
    names = tile_names("P%04dx%04d" % (case["w"], case["h"]), 1, org)
    assert names == [t[0] for t in case["tiles"]]
    for (l, u) in org:
        assert 0 <= l and 0 <= u
        assert l + GOLD["subsize"] <= max(case["w"], GOLD["subsize"]) and u + GOLD["subsize"] <= max(case["h"], GOLD["subsize"])


def test_tile_names_parse_back_in_result_merge():
    """the names are what ResultMerge splits on (`__<rate>__<left>___<up>`, ResultMerge_multi_process.py:196-205)"""
    import re
    from orientedreppoints_b200.dota.split_tiles import tile_names, tile_origins
    org = tile_origins(4000, 3000)
    for n, (l, u) in zip(tile_names("P0007", 1, org), org):
        x, y = re.findall(r"\d+", re.findall(r"__\d+___\d+", n)[0])
        assert (int(x), int(y)) == (l, u)
        assert re.findall(r"__([\d+\.]+)__\d+___", n)[0] == "1"


def test_bad_gap():
    from orientedreppoints_b200.dota.split_tiles import tile_origins
    with pytest.raises(ValueError):
        tile_origins(100, 100, subsize=64, gap=64)
