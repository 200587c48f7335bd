"""The oracle reproduces the committed fixtures of tests/golden/oracle_fixtures.json (regression pin; see the header of
tests/golden/make_oracle_fixtures.py for what the fixtures are and are not)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def test_oracle_reproduces_fixtures(oracle):
    import make_oracle_fixtures as mk
    with open(os.path.join(HERE, "golden", "oracle_fixtures.json")) as f:
        want = json.load(f)
    got = json.loads(json.dumps(mk.make()))
    assert got.keys() == want.keys()
    for key in want:
        assert got[key] == want[key], key
