#!/usr/bin/env python
"""How far is "bit-exact against the oracle" from "bit-exact against the Rust crate" for the floating-point primitives
that the oracle (and the kernels) evaluate with their own deterministic ln / exp / expm1 instead of the platform libm
the reference calls?  Draws >= 1e9 arguments exactly as SetSketch / ExpRestricted01 draw them and counts the
evaluations whose OUTCOME (register value, accept decision) differs (oracle/kmer_oracle.cpp: orc_libm_divergence).

  python scripts/libm_divergence.py [nkeys]        -> one JSON line (committed as profiles/r2_libm_divergence.json)
"""
import json
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import get_oracle  # noqa: E402


def main():
    nkeys = int(sys.argv[1]) if len(sys.argv) > 1 else 260_000_000
    orc = get_oracle()
    t0 = time.time()
    res = {"setsketch_default": orc.libm_divergence(nkeys, seed=11, params=(1.001, 4096, 20.0, 65534), points=4, m_pmh=2)}
    res["setsketch_b2_m256_and_pmh3a_m3"] = orc.libm_divergence(nkeys // 8, seed=12, params=(2.0, 256, 20.0, 62), points=4, m_pmh=3)
    res["seconds"] = round(time.time() - t0, 1)
    res["libc"] = " ".join(platform.libc_ver())
    res["nkeys"] = nkeys
    print(json.dumps(res))


if __name__ == "__main__":
    main()
