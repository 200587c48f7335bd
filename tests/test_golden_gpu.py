"""GPU parity against the committed fixtures (tests/golden/oracle_fixtures.json): the CUDA path reproduces them bit for
bit without the oracle being present in the comparison."""
import json
import os

import numpy as np
import pytest

import kmerutils_b200 as kb

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def fx():
    with open(os.path.join(HERE, "golden", "oracle_fixtures.json")) as f:
        return json.load(f)


def rows(hexrows, dtype):
    return np.stack([np.frombuffer(bytes.fromhex(h), dtype=dtype) for h in hexrows])


def test_s80_kmers_and_nthash(engine, fx):
    b, _ = engine.batch_from_ascii([fx["s80"].encode()])
    v, _ = engine.generate_kmers(b, 8, kb.KMER32, kb.HASH_CANON_INVHASH)
    assert v.tobytes().hex() == fx["s80_kmers"]["k8_kmer32_canon_invhash"]
    v, _ = engine.generate_kmers(b, 16, kb.KMER16B32)
    assert v.tobytes().hex() == fx["s80_kmers"]["k16_kmer16b32_raw"]
    v, _ = engine.generate_kmers(b, 31, kb.KMER64, kb.HASH_CANON_RAW)
    assert v.tobytes().hex() == fx["s80_kmers"]["k31_kmer64_canon"]
    h, strand = engine.nthash_canonical(b, 16)
    for i, (hx, s) in enumerate(fx["s80_nthash_k16_first8"]):
        assert f"{int(h[i, 0]):016x}" == hx and int(strand[i]) == s


def test_signatures(engine, fx):
    nb = np.array(fx["lengths"], dtype=np.uint64)
    b = engine.batch_synth(fx["seed"], nb)
    p = fx["pmh3a"]
    assert np.array_equal(engine.sketch_pmh3a(b, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 64), rows(p["k8_kmer32_m64"], np.uint32))
    assert np.array_equal(engine.sketch_pmh3a(b, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 64), rows(p["k16_kmer16b32_m64"], np.uint32))
    assert np.array_equal(engine.sketch_pmh3a(b, 21, kb.KMER64, kb.HASH_CANON_INVHASH, 64), rows(p["k21_kmer64_m64"], np.uint64))
    assert engine.sketch_pmh3a_whole(b, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 64).astype(np.uint32).tobytes().hex() == p["whole_k8_kmer32_m64"]
    s = fx["superminhash"]
    got = engine.sketch_superminhash(b, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 64, kb.HASHER_NOHASH, np.float64)
    assert np.array_equal(got.view(np.uint64), rows(s["k8_kmer32_m64_f64_nohash"], np.uint64))
    got = engine.sketch_superminhash(b, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH, 64, kb.HASHER_FNV, np.float32)
    assert np.array_equal(got.view(np.uint32), rows(s["k16_kmer16b32_m64_f32_fnv"], np.uint32))
    t = fx["setsketch"]
    prm = (t["params"][0], int(t["params"][1]), t["params"][2], int(t["params"][3]))
    assert np.array_equal(engine.sketch_setsketch(b, 8, kb.KMER32, kb.HASH_CANON_INVHASH, prm, np.uint16), rows(t["k8_kmer32_u16"], np.uint16))
    assert engine.sketch_setsketch(b, 21, kb.KMER64, kb.HASH_CANON_INVHASH, prm, np.uint16, whole=True).tobytes().hex() == t["whole_k21_kmer64_u16"]


def test_counts(engine, fx):
    nb = np.array(fx["lengths"], dtype=np.uint64)
    b = engine.batch_synth(fx["seed"], nb)
    c = fx["count"]["k31"]
    ctr = engine.counter(31, kb.KMER64, capacity=20000)
    ctr.insert_seqs(b, canonical=True)
    st = ctr.stats()
    assert (st["nb_distinct"], st["nb_unique"]) == (c["nb_distinct"], c["nb_unique"])
    keys, _ = ctr.export(1)
    assert f"{int(np.bitwise_xor.reduce(keys.astype(np.uint64))):016x}" == c["xor_keys"]
    ctr.destroy()
    c = fx["count"]["k8"]
    ctr = engine.counter(8, kb.KMER32, capacity=70000)
    ctr.insert_seqs(b, canonical=True)
    st = ctr.stats()
    assert (st["nb_distinct"], st["nb_unique"]) == (c["nb_distinct"], c["nb_unique"])
    keys, cnts = ctr.export(1)
    assert int(cnts.max()) == c["max_count"]
    assert f"{int((keys.astype(np.uint64) * cnts.astype(np.uint64)).sum(dtype=np.uint64)):016x}" == c["sum_count_times_key_mod64"]
    ctr.destroy()
