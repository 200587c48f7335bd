"""Compare fixtures dumped from the real Rust crate (rust/parity_dump) with tests/golden/oracle_fixtures.json.

    python tests/golden/compare_with_crate.py /tmp/crate_fixtures.json

Every key the two files share is compared bit for bit; the crate dump leaves out the rows of reads shorter than k (the
reference panics on them) and says which rows it kept under "rows".  All equal => the oracle's restatement of the
probminhash / rand arithmetic is pinned and "parity unpinned" can be struck from DESIGN.md; a difference names the first
sketch family to look at."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main(path):
    crate = json.load(open(path))
    orc = json.load(open(os.path.join(HERE, "oracle_fixtures.json")))
    rows = crate.get("rows", {})
    bad = 0

    def kept(name):
        for tag in ("k21", "k16", "k8"):
            if name.startswith(tag + "_"):
                return rows.get(tag)
        return None

    for fam in ("s80_kmers", "pmh3a", "superminhash", "setsketch"):
        for name, want in orc[fam].items():
            if name == "params" or name not in crate.get(fam, {}):
                continue
            got = crate[fam][name]
            if isinstance(want, list) and kept(name) is not None:
                want = [want[i] for i in kept(name)]
            ok = got == want
            bad += not ok
            print(f"{'ok  ' if ok else 'DIFF'} {fam}.{name}")
    print("oracle pinned against the crate" if not bad else f"{bad} fixture(s) differ")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
