"""Thread safety of the C ABI (the reference's entry points are `&self`, `Send + Sync` and are called from many
threads -- rayon inside, gsearch outside): calls on one context from many threads are serialised by the library and give
the single-thread results; different contexts run concurrently; the error message is per thread."""
import threading

import numpy as np
import pytest

import kmerutils_b200 as kb

pytestmark = pytest.mark.gpu


def work_items(engine):
    rng = np.random.default_rng(23)
    items = []
    for i in range(6):
        nb = rng.integers(200, 30000, 200).astype(np.uint64)
        items.append((engine.batch_synth(100 + i, nb), nb))
    return items


def run_all(engine, batch):
    out = {}
    out["pmh3a"] = engine.sketch_pmh3a(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 200)
    out["pmh3a64"] = engine.sketch_pmh3a(batch, 21, kb.KMER64, kb.HASH_CANON_INVHASH, 64)
    out["kmers"], _ = engine.generate_kmers(batch, 16, kb.KMER16B32, kb.HASH_CANON_INVHASH)
    out["smh"] = engine.sketch_superminhash(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 100)
    out["hll"] = engine.sketch_setsketch(batch, 12, kb.KMER32, kb.HASH_CANON_INVHASH, (1.001, 128, 20.0, 65534), np.uint16)
    ctr = engine.counter(31, kb.KMER64, capacity=int(batch.kmer_count(31)) + 16)
    ctr.insert_seqs(batch, canonical=True)
    st = ctr.stats()
    out["count"] = np.array([st["nb_distinct"], st["nb_unique"], st["nb_inserted"]], dtype=np.uint64)
    ctr.destroy()
    return out


def same(a, b):
    return a.keys() == b.keys() and all(np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)) for k in a)


def test_one_context_many_threads(engine):
    items = work_items(engine)
    want = [run_all(engine, b) for b, _ in items]
    got = [None] * len(items)
    errors = []

    def worker(i):
        try:
            for _ in range(3):
                got[i] = run_all(engine, items[i][0])
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(items))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i in range(len(items)):
        assert same(got[i], want[i]), f"thread {i} got different results"
    for b, _ in items:
        b.destroy()


def test_two_contexts_concurrently(engine):
    other = kb.Engine(0)
    engines = [engine, other]
    batches = [e.batch_synth(7, np.full(300, 5000, dtype=np.uint64)) for e in engines]
    want = run_all(engine, batches[0])
    got = [None, None]
    errors = []

    def worker(i):
        try:
            for _ in range(3):
                got[i] = run_all(engines[i], batches[i])
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert same(got[0], want) and same(got[1], want)
    for b in batches:
        b.destroy()
    other.close()


def test_error_message_is_per_thread(engine):
    batch = engine.batch_synth(9, np.array([1000], dtype=np.uint64))
    seen = {}

    def bad(name, fn):
        try:
            fn()
        except kb.KmuInvalid as e:
            seen[name] = str(e)

    t1 = threading.Thread(target=bad, args=("k", lambda: engine.generate_kmers(batch, 15, kb.KMER32)))
    t2 = threading.Thread(target=bad, args=("m", lambda: engine.sketch_pmh3a(batch, 8, kb.KMER32, kb.HASH_CANON_INVHASH, 1)))
    t1.start(); t2.start(); t1.join(); t2.join()
    assert "kmer size" in seen["k"] and "at least 2 hash values" in seen["m"]
    batch.destroy()
