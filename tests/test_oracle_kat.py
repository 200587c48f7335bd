"""Pins the CPU oracle on every known-answer value the reference's own unit tests hold
(tests/golden/reference_kats.json, transcribed with file:line) and on the derived vectors of
SURVEY.md Appendix C.  CPU only."""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
S80 = GOLD["s80"].encode()


def kmer_from_str(oracle, s, ktype):
    """Kmer32bit::from_str / Kmer64bit::from_str (kmer32bit.rs:186-208): push every base into new(k)."""
    k = len(s)
    w = oracle.L.orc_kmer_build(0, k, ktype)
    for c in s:
        w = oracle.L.orc_kmer_push(w, k, ktype, "ACGT".index(c))
    return w


def kmer_to_str(word, k):
    return "".join("ACGT"[(int(word) >> (2 * (k - 1 - j))) & 3] for j in range(k))


@pytest.mark.parametrize("case", GOLD["pack_2bit"])
def test_pack_2bit(oracle, case):
    packed = oracle.pack_2bit(case["ascii"].encode())
    want = bytes.fromhex(case["bytes_hex"])
    assert packed.tobytes() == want
    assert oracle.unpack_2bit(packed, len(case["ascii"])) == case["ascii"].encode()
    for i, c in enumerate(case["ascii"]):  # Sequence::get_base, sequence.rs:901-928
        assert oracle.L.orc_get_base(packed.ctypes.data_as(ol.u8p), i) == "ACGT".index(c)


def test_pack_2bit_rejects_non_acgt(oracle):
    with pytest.raises(ValueError):  # Alphabet2b::encode panics, alphabet.rs:125
        oracle.pack_2bit(b"ACGNT")
    assert oracle.count_non_acgt(b"ACGNTxacgt") == 2  # alphabet.rs:28-31, case-insensitive :157-159


@pytest.mark.parametrize("case", GOLD["encode_and_add"])
def test_encode_and_add(oracle, case):
    packed, kept = oracle.encode_and_add(case["ascii"].encode())
    assert kept == len(case["kept"])
    assert oracle.unpack_2bit(packed, kept) == case["kept"].encode()


@pytest.mark.parametrize("case", GOLD["seq_revcomp"])
def test_sequence_revcomp(oracle, case):
    packed = oracle.pack_2bit(case["seq"].encode())
    rc = oracle.seq_revcomp(packed, len(case["seq"]))
    assert oracle.unpack_2bit(rc, len(case["seq"])) == case["revcomp"].encode()


@pytest.mark.parametrize("case", GOLD["kmer16b32_revcomp"])
def test_kmer16b32_revcomp(oracle, case):
    w, want = int(case["word"], 2), int(case["revcomp"], 2)
    assert oracle.L.orc_kmer_revcomp(w, 16, ol.KMER16B32) == want
    assert oracle.L.orc_kmer_revcomp(want, 16, ol.KMER16B32) == w


@pytest.mark.parametrize("ktype", [ol.KMER32, ol.KMER64])
@pytest.mark.parametrize("case", GOLD["kmer_str_revcomp"])
def test_kmer_str_revcomp(oracle, case, ktype):
    k = len(case["kmer"])
    w = kmer_from_str(oracle, case["kmer"], ktype)
    want = kmer_from_str(oracle, case["revcomp"], ktype)
    got = oracle.L.orc_kmer_revcomp(w, k, ktype)
    assert got == want
    if ktype == ol.KMER32:  # the word carries k in its top four bits (kmer32bit.rs:68-76)
        assert got >> 28 == k
        assert oracle.L.orc_kmer_compressed_value(got, k, ktype) == got & 0x0FFFFFFF


def test_kmer32_order(oracle):
    c = GOLD["kmer32_order"]
    a, b = kmer_from_str(oracle, c["a"], ol.KMER32), kmer_from_str(oracle, c["b"], ol.KMER32)
    assert oracle.L.orc_kmer_cmp(a, a, 12, ol.KMER32) == 0
    assert (oracle.L.orc_kmer_cmp(a, b, 12, ol.KMER32) > 0) == c["a_gt_b"]
    # a longer k-mer is greater whatever the value (header compared first, kmer32bit.rs:47-55)
    short = kmer_from_str(oracle, "TTTTT", ol.KMER32)
    long_ = kmer_from_str(oracle, "AAAAAA", ol.KMER32)
    assert oracle.L.orc_kmer_cmp(long_, short, 0, ol.KMER32) > 0


@pytest.mark.parametrize("k,ktype", [(16, ol.KMER16B32), (11, ol.KMER32), (21, ol.KMER64), (8, ol.KMER32),
                                     (31, ol.KMER64), (32, ol.KMER64)])
def test_generate_kmers_strings(oracle, k, ktype):
    # kmergenerator.rs:596-699 (16-mers), :703-732 (11-mers), :943-972 (21-mers): every k-mer
    # decompresses to seq[i..i+k]; count = L - k + 1 (:897-939)
    packed = oracle.pack_2bit(S80)
    words = oracle.generate_kmers(packed, 80, k, ktype)
    assert len(words) == 80 - k + 1
    for i, w in enumerate(words):
        v = oracle.L.orc_kmer_compressed_value(int(w), k, ktype)
        assert kmer_to_str(v, k) == S80[i:i + k].decode()


def test_generate_kmers_first_words_and_range(oracle):
    packed = oracle.pack_2bit(S80)
    words = oracle.generate_kmers(packed, 80, 16, ol.KMER16B32)
    assert [int(w) for w in words[:3]] == [int(x, 16) for x in GOLD["kmer16_first_words"]["words"]]
    # range 3..25 (kmergenerator.rs:640-699): 25 - 3 - 16 + 1 = 7 k-mers, the first one at base 3
    sub = oracle.generate_kmers(packed, 80, 16, ol.KMER16B32, 3, 25)
    assert len(sub) == 7 and np.array_equal(sub, words[3:10])
    # a sequence shorter than k yields nothing (kmergenerator.rs:93-100)
    assert len(oracle.generate_kmers(packed, 15, 16, ol.KMER16B32)) == 0


def test_kmer_type_guards(oracle):
    packed = oracle.pack_2bit(S80)
    for k, t in [(15, ol.KMER32), (16, ol.KMER32), (15, ol.KMER16B32), (17, ol.KMER16B32), (33, ol.KMER64)]:
        with pytest.raises(ValueError):  # the reference panics (kmergenerator.rs:218,311,415)
            oracle.generate_kmers(packed, 80, k, t)


def test_weighted_3mers(oracle):
    g = GOLD["weighted_3mers"]
    seq = g["seq"].encode()
    words = oracle.generate_kmers(oracle.pack_2bit(seq), len(seq), 3, ol.KMER32)
    counts = {}
    for w in words:
        s = kmer_to_str(int(w) & 0x0FFFFFFF, 3)
        counts[s] = counts.get(s, 0) + 1
    assert counts == g["counts"]
    assert sum(counts.values()) == 46


def test_weighted_15mers(oracle):
    seq = GOLD["weighted_15mers"]["seq"]
    words = oracle.generate_kmers(oracle.pack_2bit(seq.encode()), len(seq), 15, ol.KMER64)
    vals, cnt = np.unique(words, return_counts=True)
    for v, c in zip(vals, cnt):
        s = kmer_to_str(v, 15)
        assert seq.count(s) >= 1 and sum(1 for i in range(len(seq) - 14) if seq[i:i + 15] == s) == c
    assert sorted(set(cnt.tolist())) == [1, 2]


# ---- SURVEY Appendix C (derived by hand from the reference formulas) -----------------------------
def test_appendix_c_kmer32_k8(oracle):
    packed = oracle.pack_2bit(S80)
    w0 = int(oracle.generate_kmers(packed, 80, 8, ol.KMER32)[0])
    assert w0 == 0x8000d02a
    rc = oracle.L.orc_kmer_revcomp(w0, 8, ol.KMER32)
    assert rc == 0x800057f8
    assert oracle.L.orc_apply_hash(w0, 8, ol.KMER32, ol.HASH_CANON_RAW) == 0x800057f8
    assert oracle.L.orc_int32_hash(0x800057f8) == 0x8360d6a4
    assert oracle.L.orc_apply_hash(w0, 8, ol.KMER32, ol.HASH_CANON_INVHASH) == 0x8360d6a4
    assert oracle.L.orc_nohash_seed(0x8360d6a4, 4) == 0xa4d66083  # NoHashHasher, nohasher.rs:22-48
    assert oracle.L.orc_apply_hash(w0, 8, ol.KMER32, ol.HASH_MASKED_VALUE) == 0xd02a
    assert oracle.L.orc_apply_hash(w0, 8, ol.KMER32, ol.HASH_IDENTITY_RAW) == 0x8000d02a


def test_appendix_c_nthash(oracle):
    packed = oracle.pack_2bit(S80)
    w16 = oracle.generate_kmers(packed, 80, 16, ol.KMER16B32)
    want = [(0x9840eab169670ddf, 0x684a2ec1114d51c5, 0x684a2ec1114d51c5, 1),
            (0x45ff6533035e369e, 0x0e9a4f48606ebe72, 0x0e9a4f48606ebe72, 1),
            (0x76f05375b83658db, 0xb3ae9d3d5337da01, 0x76f05375b83658db, 0)]
    for w, exp in zip(w16[:3], want):
        assert oracle.nthash_canonical(int(w), 16, ol.KMER16B32) == exp
    w8 = oracle.generate_kmers(packed, 80, 8, ol.KMER32)
    f8 = [0x935533199c1dfb81, 0x4f6868cb4fb9a55e, 0x319aaf47aa9e02f9]
    for w, f in zip(w8[:3], f8):
        got = oracle.nthash_canonical(int(w) & 0x0FFFFFFF, 8, ol.KMER32)
        assert got[0] == f and got[3] == 0
        assert oracle.L.orc_nthash_init(int(w) & 0x0FFFFFFF, 8, ol.KMER32) == f
    mult = oracle.nthash_mult(0x684a2ec1114d51c5, 16, 4)  # nthash.rs:63-72
    assert [int(x) for x in mult] == [0x684a2ec1114d51c5, 0x9f8aa4cb1f1ddcf8, 0x07d4d399050cea95, 0x701f02551303a00d]


def test_nthash_roll_consistency(oracle):
    # nthash.rs:303-335: the rolled value equals the re-initialised one for every window
    # (the reference checks its 8-bit functions; here the same property on the 2-bit formula)
    packed = oracle.pack_2bit(S80)
    k = 11
    words = oracle.generate_kmers(packed, 80, k, ol.KMER32)
    rotl = lambda x, r: ((x << (r % 64)) | (x >> (64 - r % 64))) & (2**64 - 1) if r % 64 else x
    seeds = [0x3c8bfbb395c60474, 0x3193c18562a02b4c, 0x20323ed082572324, 0x295549f54be24456]
    h = oracle.L.orc_nthash_init(int(words[0]) & 0x0FFFFFFF, k, ol.KMER32)
    for i in range(1, len(words)):
        old, new = "ACGT".index(chr(S80[i - 1])), "ACGT".index(chr(S80[i + k - 1]))
        h = rotl(h, 1) ^ rotl(seeds[old], k) ^ seeds[new]
        assert h == oracle.L.orc_nthash_init(int(words[i]) & 0x0FFFFFFF, k, ol.KMER32)


# ---- generators the sketch arithmetic is built from ---------------------------------------------
def test_splitmix_xoshiro_known_answers(oracle):
    # SplitMix64 reference outputs for seed 0 (Vigna's splitmix64.c test vector, also used by rand_xoshiro's tests)
    s = np.zeros(4, dtype=np.uint64)
    oracle.L.orc_xoshiro_seed(0, s.ctypes.data_as(ol.u64p))
    assert [int(x) for x in s] == [0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4, 0x06c45d188009454f, 0xf88bb8a8724c81ec]
    # xoshiro256++ from state [1,2,3,4] (rand_xoshiro's xoshiro256plusplus reference test)
    st = np.array([1, 2, 3, 4], dtype=np.uint64)
    out = [oracle.L.orc_xoshiro_next(st.ctypes.data_as(ol.u64p)) for _ in range(4)]
    assert out == [41943041, 58720359, 3588806011781223, 3591011842654386]


def test_invhash_is_a_bijection_sample(oracle):
    # int32_hash / int64_hash are invertible (probminhash::invhash): no collisions on a sample
    xs = np.arange(0, 1 << 16, dtype=np.uint64)
    h32 = {oracle.L.orc_int32_hash(int(x)) for x in xs}
    h64 = {oracle.L.orc_int64_hash(int(x)) for x in xs}
    assert len(h32) == len(xs) and len(h64) == len(xs)


# ---- statistical tests of the reference re-run on the oracle (seqsketchjaccard.rs:742-944) --------
def _sig(oracle, seq, k, ktype, kind, m):
    return oracle.sketch_pmh3a_seq(oracle.pack_2bit(seq), len(seq), k, ktype, kind, m)


def _revcomp_str(s):
    return "".join("TGCA"["ACGT".index(chr(c))] for c in reversed(s)).encode()


@pytest.mark.parametrize("k,ktype,m", [(5, ol.KMER32, 4000), (16, ol.KMER16B32, 50), (16, ol.KMER64, 50)])
def test_pmh3a_reference_inequalities(oracle, k, ktype, m):
    a = _sig(oracle, S80, k, ktype, ol.HASH_CANON_INVHASH, m)
    b = _sig(oracle, S80[:40], k, ktype, ol.HASH_CANON_INVHASH, m)
    j = float(np.mean(a == b))
    assert j >= 0.75 * (40 - k) / (80 - k)  # seqsketchjaccard.rs:784-785, 902-903
    rc = _sig(oracle, _revcomp_str(S80), k, ktype, ol.HASH_CANON_INVHASH, m)
    assert float(np.mean(a == rc)) >= 1.0  # :791, 908-909, 942-943
    ia = _sig(oracle, S80, k, ktype, ol.HASH_IDENTITY_RAW, m)
    irc = _sig(oracle, _revcomp_str(S80), k, ktype, ol.HASH_IDENTITY_RAW, m)
    assert float(np.mean(ia == irc)) <= 0.1  # :850


def test_pmh3a_short_sequence_is_all_zero(oracle):
    # 0 < L < k: empty multiplicity map, signature = m copies of Val::default()
    sig = _sig(oracle, b"ACGTA", 8, ol.KMER32, ol.HASH_CANON_INVHASH, 200)
    assert not sig.any()


# ---- SuperMinHash on the oracle: reference inequalities (seqsketchjaccard.rs:947-1005, seqminhash.rs:193-250) ----
def _smh(oracle, seq, k, ktype, kind, m, hasher, dtype=np.float64):
    packed = oracle.pack_2bit(seq)
    return oracle.sketch_superminhash_batch(packed, np.zeros(1, np.uint64), np.array([len(seq)], np.uint64), k, ktype,
                                            kind, m, hasher, dtype)[0]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("hasher", [0, 1])
def test_superminhash_reference_inequalities(oracle, hasher, dtype):
    k, ktype, m = 16, ol.KMER16B32, 50
    a = _smh(oracle, S80, k, ktype, ol.HASH_CANON_INVHASH, m, hasher, dtype)
    half = _smh(oracle, S80[:40], k, ktype, ol.HASH_CANON_INVHASH, m, hasher, dtype)
    rc = _smh(oracle, _revcomp_str(S80), k, ktype, ol.HASH_CANON_INVHASH, m, hasher, dtype)
    assert np.mean(a == half) >= 0.75 * (40 - k) / (80 - k)
    assert np.mean(a == rc) >= 1.0
    ia = _smh(oracle, S80, k, ktype, ol.HASH_IDENTITY_RAW, m, hasher, dtype)
    irc = _smh(oracle, _revcomp_str(S80), k, ktype, ol.HASH_IDENTITY_RAW, m, hasher, dtype)
    assert np.mean(ia == irc) <= 0.1


def test_superminhash_properties(oracle):
    # idempotent (a duplicated sequence changes nothing), order independent, mergeable by min
    s1, s2 = S80[:50], S80[30:]
    p1, p2 = oracle.pack_2bit(s1), oracle.pack_2bit(s2)
    buf = np.zeros(64, np.uint8)
    buf[:len(p1)] = p1
    buf[32:32 + len(p2)] = p2
    off = np.array([0, 32], np.uint64)
    nb = np.array([len(s1), len(s2)], np.uint64)
    both = oracle.sketch_superminhash_seqs(buf, off, nb, 8, ol.KMER32, ol.HASH_CANON_INVHASH, 64)
    rev = oracle.sketch_superminhash_seqs(buf, off[::-1].copy(), nb[::-1].copy(), 8, ol.KMER32, ol.HASH_CANON_INVHASH, 64)
    each = oracle.sketch_superminhash_batch(buf, off, nb, 8, ol.KMER32, ol.HASH_CANON_INVHASH, 64)
    assert np.array_equal(both, rev)
    assert np.array_equal(both, each.min(axis=0))
    assert (both < 64).all() and (both >= 0).all()
    # no k-mer at all: the initial value F::from(u32::MAX)
    empty = _smh(oracle, b"ACG", 8, ol.KMER32, ol.HASH_CANON_INVHASH, 16, 0)
    assert (empty == 4294967295.0).all()


# ---- amino-acid k-mers (src/aautils/kmeraa.rs:920-1021) ---------------------------------------------
AA_CODES = {c: v for c, v in zip("ACDEFGHIKLMNP", range(1, 14))}
AA_CODES.update({c: v for c, v in zip("QRSTVWY", range(15, 22))})
AA_DECODE = {v: c for c, v in AA_CODES.items()}


def aa_to_str(v, k):
    return "".join(AA_DECODE[(int(v) >> (5 * (k - 1 - j))) & 31] for j in range(k))


@pytest.mark.parametrize("ktype", [ol.KMERAA32, ol.KMERAA64])
def test_aa_iterator_range(oracle, ktype):
    g = GOLD["aa"]
    prot = np.frombuffer(g["protein"].encode(), dtype=np.uint8)
    got = oracle.generate_kmers(prot, len(prot), 4, ktype, 3, 10)
    assert [aa_to_str(v, 4) for v in got] == g["range_3_10_4mers"]


def test_aa_iterator_end_and_guards(oracle):
    g = GOLD["aa"]
    prot = np.frombuffer(g["protein"][:32].encode(), dtype=np.uint8)
    got = oracle.generate_kmers(prot, 32, 8, ol.KMERAA64)
    assert len(got) == 32 - 8 + 1 and aa_to_str(got[-1], 8) == g["last_8mer"]
    for i, v in enumerate(got):
        assert aa_to_str(v, 8) == g["protein"][i:i + 8]
    with pytest.raises(ValueError):  # KmerAA32bit holds at most 6 residues (kmeraa.rs:212-214)
        oracle.generate_kmers(prot, 32, 7, ol.KMERAA32)
    with pytest.raises(ValueError):
        oracle.generate_kmers(prot, 32, 13, ol.KMERAA64)
    assert len(oracle.generate_kmers(prot, 32, 12, ol.KMERAA64)) == 21  # k = 12 works through KmerBuilder (:622-623)
    bad = np.frombuffer(b"MTEQIELIKLYSTRILALAAQMPHVGXLDNPD", dtype=np.uint8)
    with pytest.raises(ValueError):  # Alphabet::encode panics on 'X' (kmeraa.rs:106)
        oracle.generate_kmers(bad, 32, 8, ol.KMERAA64)
    assert oracle.aa_filter(b"MTxEQ*IB") == b"MTEQI"  # new_filtered (kmeraa.rs:447-456); B, x, * are not in the alphabet


AA_STR1 = b"MTEQIELIKLYSTRILALAAQMPHVGSLDNPDASAMKRSPLCGSKVTVDVIMQNGKITFDGFEVLAPASEYKNRHASILLSLDATAEACASIAAQNSA"
AA_STR2 = b"MTEQIELIKLYSTRILALAAQMPHVGSLDNPDASAMKRSPLCGSKVMTEQIELIKLYSTRILALAAQMPHVGSLDNPDASAMKRSPLCGSKV"


@pytest.mark.parametrize("ktype,m", [(ol.KMERAA64, 400), (ol.KMERAA32, 800)])
def test_aa_probminhash_reference_inequality(oracle, ktype, m):
    # aautils/setsketchert.rs:1217-1265 (64 bit, m = 400) and :1267-1317 (32 bit): the second string is the
    # first half of the first one repeated; k = 5, masked value, |J - 0.5| < 0.1
    sigs = []
    for s_ in (AA_STR1, AA_STR2):
        a = np.frombuffer(s_, dtype=np.uint8)
        sigs.append(oracle.sketch_pmh3a_seq(a, len(a), 5, ktype, ol.HASH_MASKED_VALUE, m))
    assert abs(float(np.mean(sigs[0] == sigs[1])) - 0.5) < 0.1


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_aa_superminhash_reference_inequality(oracle, dtype):
    # aautils/setsketchert.rs:1319-1391: SuperMinHash f64 / f32 on the same strings, |J - 0.5| < 0.1
    sigs = []
    for s_ in (AA_STR1, AA_STR2):
        a = np.frombuffer(s_, dtype=np.uint8)
        sigs.append(oracle.sketch_superminhash_batch(a, np.zeros(1, np.uint64), np.array([len(a)], np.uint64), 5,
                                                     ol.KMERAA64, ol.HASH_MASKED_VALUE, 800, 0, dtype)[0])
    assert abs(float(np.mean(sigs[0] == sigs[1])) - 0.5) < 0.1


# ---- SetSketch building blocks and properties (the reference has no test for HLL) ---------
def test_det_log_exp_accuracy(oracle):
    import math
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.exp(rng.uniform(-700, 700, 2000)), rng.uniform(0.5, 2.0, 2000), 1 + rng.uniform(-1e-6, 1e-6, 500)])
    for x in xs:
        assert abs(oracle.L.orc_det_log(float(x)) - math.log(x)) <= 2.3e-16 * max(abs(math.log(x)), 1e-300) + 1e-320
    for y in rng.uniform(-30, 30, 2000):
        assert abs(oracle.L.orc_det_exp(float(y)) - math.exp(y)) <= 2.3e-16 * math.exp(y)
    assert oracle.L.orc_det_log(1.0) == 0.0 and oracle.L.orc_det_exp(0.0) == 1.0


def test_ziggurat_exp1_moments(oracle):
    v = np.array([oracle.L.orc_exp1_from_seed(s, 0) for s in range(40000)])
    assert (v > 0).all()
    assert abs(v.mean() - 1.0) < 0.02 and abs(v.var() - 1.0) < 0.06  # Exp(1)
    assert abs((v > 1.0).mean() - np.exp(-1.0)) < 0.01


def _hll(oracle, seqs, k=12, params=(1.001, 512, 20.0, 65534)):
    packed = [oracle.pack_2bit(s) for s in seqs]
    off, cur = [], 0
    for p in packed:
        off.append(cur)
        cur += (len(p) + 15) // 16 * 16
    buf = np.zeros(cur + 16, np.uint8)
    for o, p in zip(off, packed):
        buf[o:o + len(p)] = p
    nb = np.array([len(s) for s in seqs], np.uint64)
    off = np.array(off, np.uint64)
    each = oracle.sketch_setsketch_batch(buf, off, nb, k, ol.KMER32, ol.HASH_CANON_INVHASH, params)
    both = oracle.sketch_setsketch_seqs(buf, off, nb, k, ol.KMER32, ol.HASH_CANON_INVHASH, params)
    return each, both


def test_setsketch_properties(oracle):
    a = oracle.synth_ascii(90, 0, 30000)
    b = oracle.synth_ascii(90, 20000, 30000)  # overlaps the last third of a
    each, both = _hll(oracle, [a, b])
    assert np.array_equal(both, each.max(axis=0))  # mergeable by max (SetSketcher::merge)
    each_rev, both_rev = _hll(oracle, [b, a])
    assert np.array_equal(both, both_rev)  # order independent
    dup_each, dup_both = _hll(oracle, [a, a])
    assert np.array_equal(dup_both, each[0])  # idempotent
    assert (each > 0).all() and each.max() <= 65535
    # more distinct k-mers -> larger registers on average; Jaccard by equal registers is about |A n B| / |A u B|
    small, _ = _hll(oracle, [a[:3000]])
    assert small.mean() < each[0].mean()
    j = float(np.mean(each[0] == each[1]))
    assert 0.1 < j < 0.4  # true Jaccard = 10000 / 50000 = 0.2
