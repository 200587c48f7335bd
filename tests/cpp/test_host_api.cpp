// Parity tests of the C++ host layer (include/kmerutils_b200.hpp), written the way the reference's own unit tests
// read (same inputs, same assertions), plus bit-exact comparison with the CPU oracle (test infrastructure).
// Needs a GPU: run by tests/test_host_api_cpp.py under `-m gpu`; the CPU suite only compiles and links it.
//
// Reference tests mirrored:
//   src/base/kmergenerator.rs:596-621   test_gen_kmer16b32bit_80bases (+ the first three words)
//   src/base/kmergenerator.rs:661-700   test_gen_kmer16b32bit_50bases_range_iterator
//   src/base/kmer32bit.rs:228-312       reverse complement / ordering of Kmer32bit
//   src/base/kmercount.rs:1524-1565     test_kmer_counter
//   src/base/kmergenerator.rs:776-850   test_generate_weighted_kmer32bit
//   src/base/kmer.rs:45-145, nthash.rs:76-120   NtHash methods of the two u32 k-mer types
//   src/sketching/seqsketchjaccard.rs:742-791   test_pminhasha_kmer_smallb
//   src/sketching/seqsketchjaccard.rs:947-1004  test_superminhash_kmer_16b32bit_serial
//   src/sketching/seqsketchjaccard.rs:1015-...  test_reload_sketch_file
//   src/sketching/seqblocksketch.rs:458-496     test_block_32bit_sketch
//   src/aautils/kmeraa.rs:920-1021              test_seqaa_32bit_iterator_range, test_seqaa_iterator_end
//   src/aautils/setsketchert.rs:1218-1266       test_seqaa_probminhash_64bit
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/kmerutils_b200.hpp"
#include "../../oracle/kmer_oracle.hpp"

using namespace kmerutils;
using namespace kmerutils::base;
using namespace kmerutils::sketching;

static int failures = 0;
#define EXPECT(cond)                                                          \
    do {                                                                      \
        if (!(cond)) {                                                        \
            std::fprintf(stderr, "FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                       \
        }                                                                     \
    } while (0)

static const std::string S80 = "TCAAAGGGAAACATTCAAAATCAGTATGCGCCCGTTCAGTTACGTATTGCTCTCGCTAATGAGATGGGCTGGGTACAGAG";

static std::string synth(uint64_t seed, size_t n) {
    std::string s(n, 'A');
    orc_synth_ascii(seed, 0, n, (uint8_t*)s.data());
    return s;
}

static void test_gen_kmer16b32bit_80bases() {
    Sequence seq(S80, 2);
    EXPECT(seq.size() == 80);
    const std::vector<Kmer16b32bit> v = KmerGenerator<Kmer16b32bit>(16).generate_kmer(seq);
    EXPECT(v.size() == 80 - 16 + 1);
    for (size_t i = 0; i < v.size(); ++i) {
        const auto u = v[i].get_uncompressed_kmer();
        EXPECT(std::string(u.begin(), u.end()) == S80.substr(i, 16));
    }
    EXPECT(v[0].v == 0xd02a013du && v[1].v == 0x40a804f4u && v[2].v == 0x02a013d0u);  // kmergenerator.rs:612-621
}

static void test_gen_kmer_range_and_types() {
    Sequence seq(S80.substr(0, 50), 2);
    // range iterator: k-mers starting in [3, 25 - 16] (kmergenerator.rs:661-700)
    const auto r = KmerGenerator<Kmer16b32bit>(16).generate_kmer_in_range(seq, 3, 25);
    EXPECT(r.size() == 25 - 3 - 16 + 1);
    for (size_t i = 0; i < r.size(); ++i) {
        const auto u = r[i].get_uncompressed_kmer();
        EXPECT(std::string(u.begin(), u.end()) == S80.substr(3 + i, 16));
    }
    const auto k11 = KmerGenerator<Kmer32bit>(11).generate_kmer(seq);
    EXPECT(k11.size() == 40);
    for (size_t i = 0; i < k11.size(); ++i) {
        EXPECT(k11[i].get_nb_base() == 11);
        const auto u = k11[i].get_uncompressed_kmer();
        EXPECT(std::string(u.begin(), u.end()) == S80.substr(i, 11));
    }
    const auto k31 = KmerGenerator<Kmer64bit>(31).generate_kmer(seq);
    EXPECT(k31.size() == 20);
    for (size_t i = 0; i < k31.size(); ++i) {
        const auto u = k31[i].get_uncompressed_kmer();
        EXPECT(std::string(u.begin(), u.end()) == S80.substr(i, 31));
        // the host value type agrees with the oracle's word arithmetic
        EXPECT(k31[i].reverse_complement().v == orc_kmer_revcomp(k31[i].v, 31, ORC_KMER64));
        EXPECT(k31[i].push(2).v == orc_kmer_push(k31[i].v, 31, ORC_KMER64, 2));
    }
    for (const Kmer32bit& km : k11) {
        EXPECT(km.reverse_complement().v == (uint32_t)orc_kmer_revcomp(km.v, 11, ORC_KMER32));
        EXPECT(km.push(3).v == (uint32_t)orc_kmer_push(km.v, 11, ORC_KMER32, 3));
    }
    // bad k for the type: the reference panics (kmergenerator.rs:311)
    bool threw = false;
    try {
        KmerGenerator<Kmer32bit>(15).generate_kmer(seq);
    } catch (const Panic&) {
        threw = true;
    }
    EXPECT(threw);
    // non-ACGT character: Sequence::new panics (alphabet.rs:125)
    threw = false;
    try {
        Sequence bad(std::string("ACGTNACGT"), 2);
    } catch (const Panic&) {
        threw = true;
    }
    EXPECT(threw);
}

static void test_kmer32bit_revcomp_order() {
    Sequence a(std::string("TACGAGTAGGAT"), 2), b(std::string("ACTTGGAACGTT"), 2);
    const Kmer32bit ka = KmerGenerator<Kmer32bit>(12).generate_kmer(a)[0];
    const Kmer32bit kb = KmerGenerator<Kmer32bit>(12).generate_kmer(b)[0];
    auto str = [](const Kmer32bit& k) {
        const auto u = k.get_uncompressed_kmer();
        return std::string(u.begin(), u.end());
    };
    EXPECT(str(ka.reverse_complement()) == "ATCCTACTCGTA");  // kmer32bit.rs:228-259
    EXPECT(str(kb.reverse_complement()) == "AACGTTCCAAGT");
    EXPECT(kb < ka);                                           // kmer32bit.rs:294-312
    const auto d = a.get_reverse_complement().decompress();
    EXPECT(std::string(d.begin(), d.end()) == "ATCCTACTCGTA");  // sequence.rs:947-961
}

static void test_kmer_counter() {
    KmerCounter<Kmer16b32bit> kmer_counter(0.03f, 10000000, 8, 16);
    Sequence seq(S80, 2);
    const std::vector<Kmer16b32bit> vkmer = KmerGenerator<Kmer16b32bit>(16).generate_kmer(seq);
    kmer_counter.insert_kmer(vkmer[0]);
    kmer_counter.insert_kmer(vkmer[1]);
    uint64_t x = 12345;
    std::vector<uint32_t> want(vkmer.size(), 0);
    want[0] = 1;
    want[1] = 2;
    for (int i = 0; i < 1000000; ++i) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        const size_t xsi = 2 + (size_t)((x >> 33) % (vkmer.size() - 2));
        kmer_counter.insert_kmer(vkmer[xsi]);
        ++want[xsi];
    }
    kmer_counter.insert_kmer(vkmer[1]);
    EXPECT(kmer_counter.get_count(vkmer[0]) == 1);  // kmercount.rs:1558
    EXPECT(kmer_counter.get_count(vkmer[1]) == 2);  // kmercount.rs:1559
    const auto got = kmer_counter.get_counts(vkmer);
    for (size_t i = 0; i < vkmer.size(); ++i) EXPECT(got[i] == (want[i] > 255 ? 255u : want[i]));  // saturation :1615
    EXPECT(kmer_counter.get_count(Kmer16b32bit::from_word(0x12345678u, 16)) == 0);                // never inserted :1612
    EXPECT(kmer_counter.get_nb_distinct() == vkmer.size());
    EXPECT(kmer_counter.get_nb_unique() == 1);

    // count_kmer_threaded_one_to_many against the oracle's exact multiset
    std::vector<std::string> reads;
    for (int i = 0; i < 300; ++i) reads.push_back(synth(77, 20000).substr((size_t)i * 50, 400));
    const std::vector<Sequence> seqvec = Sequence::new_batch(reads, 2);
    auto pool = count_kmer_threaded_one_to_many<Kmer64bit>(seqvec, 4, 8, 31);  // 4 counters, 8 bits per count (:881-893)
    EXPECT(pool->counters.size() == 4 && pool->get_count_nb_bits() == 8);
    std::vector<uint8_t> packed;
    std::vector<uint64_t> off, nb;
    for (const Sequence& s : seqvec) {
        off.push_back(packed.size());
        nb.push_back(s.size());
        packed.insert(packed.end(), s.packed().begin(), s.packed().end());
    }
    std::vector<uint64_t> keys(200000), cnts(200000);
    const uint64_t nd = orc_count_kmers(packed.data(), off.data(), nb.data(), seqvec.size(), 31, ORC_KMER64, 1, keys.data(), cnts.data(), keys.size());
    EXPECT(nd <= keys.size());
    uint64_t nu = 0;
    std::vector<Kmer64bit> probes;
    for (uint64_t i = 0; i < nd; ++i) {
        nu += cnts[i] == 1;
        probes.push_back(Kmer64bit::from_word(keys[i], 31));
    }
    EXPECT(pool->get_nb_distinct() == nd);
    EXPECT(pool->get_nb_unique() == nu);
    const auto pc = pool->get_counts(probes);
    for (uint64_t i = 0; i < nd; ++i) EXPECT(pc[i] == (cnts[i] > 255 ? 255 : cnts[i]));
    // every k-mer sits in the counter its dispatch names (DispatchableT, kmercount.rs:382-420), nowhere else
    for (uint64_t i = 0; i < nd; i += 97) {
        const size_t loc = probes[i].dispatch(4);
        EXPECT(loc == (size_t)(orc_int64_hash(keys[i]) % 4));
        for (size_t c = 0; c < 4; ++c) EXPECT(pool->counters[c]->get_count(probes[i]) == (c == loc ? pc[i] : 0u));
        EXPECT(pool->get_above2_count(probes[i]) == (pc[i] >= 2 ? pc[i] : 0u));
    }
    // count_kmer_thread_independant (kmercount.rs:797-867): the same pool, 8-bit counts
    auto pool2 = count_kmer_thread_independant<Kmer64bit>(seqvec, 3, 31);
    EXPECT(pool2->counters.size() == 3 && pool2->get_nb_distinct() == nd && pool2->get_nb_unique() == nu);
}

// the advisor's round-1 finding: Kmer32bit::get_compressed_value() is the value WITHOUT the length header
// (kmer32bit.rs:173-178); insert_kmer / get_count and insert_sequences must key the table the same way
static void test_kmer32bit_counter_keys() {
    Sequence seq(S80, 2);
    const std::vector<Kmer32bit> v = KmerGenerator<Kmer32bit>(11).generate_kmer(seq);
    for (const Kmer32bit& km : v) {
        EXPECT(km.get_compressed_value() == (km.v & 0x0FFFFFFFu));
        EXPECT((km.v >> 28) == 11);
        EXPECT(km.get_compressed_value() == (uint32_t)orc_kmer_compressed_value(km.v, 11, ORC_KMER32));
        EXPECT(Kmer32bit::build(km.get_compressed_value(), 11) == km);
    }
    std::vector<std::string> reads;
    for (int i = 0; i < 200; ++i) reads.push_back(synth(91, 30000).substr((size_t)i * 70, 300));
    const std::vector<Sequence> seqvec = Sequence::new_batch(reads, 2);
    std::vector<uint8_t> packed;
    std::vector<uint64_t> off, nb;
    for (const Sequence& s : seqvec) {
        off.push_back(packed.size());
        nb.push_back(s.size());
        packed.insert(packed.end(), s.packed().begin(), s.packed().end());
    }
    {
        std::vector<uint64_t> keys(100000), cnts(100000);
        const uint64_t nd = orc_count_kmers(packed.data(), off.data(), nb.data(), seqvec.size(), 11, ORC_KMER32, 1, keys.data(), cnts.data(), keys.size());
        KmerCounter<Kmer32bit> counter(0.03f, 100000, 8, 11);
        counter.insert_sequences(as_refs(seqvec), true);
        std::vector<Kmer32bit> probes;
        for (uint64_t i = 0; i < nd; ++i) probes.push_back(Kmer32bit::build((uint32_t)keys[i], 11));
        auto got = counter.get_counts(probes);
        for (uint64_t i = 0; i < nd; ++i) EXPECT(got[i] == (cnts[i] > 255 ? 255 : cnts[i]));
        EXPECT(counter.get_count(probes[0]) == got[0]);
        // the same k-mers once more one by one: every count grows by one, no new distinct key appears
        for (const Kmer32bit& km : probes) counter.insert_kmer(km);
        got = counter.get_counts(probes);
        for (uint64_t i = 0; i < nd; ++i) EXPECT(got[i] == (cnts[i] + 1 > 255 ? 255 : cnts[i] + 1));
        EXPECT(counter.get_nb_distinct() == nd);
    }
    {
        std::vector<uint64_t> keys(100000), cnts(100000);
        const uint64_t nd = orc_count_kmers(packed.data(), off.data(), nb.data(), seqvec.size(), 23, ORC_KMER64, 1, keys.data(), cnts.data(), keys.size());
        KmerCounter<Kmer64bit> counter(0.03f, 100000, 8, 23);
        std::vector<Kmer64bit> probes;
        for (uint64_t i = 0; i < nd; ++i) probes.push_back(Kmer64bit::build(keys[i], 23));
        for (const Kmer64bit& km : probes) counter.insert_kmer(km);
        counter.insert_sequences(as_refs(seqvec), true);
        const auto got = counter.get_counts(probes);
        for (uint64_t i = 0; i < nd; ++i) EXPECT(got[i] == (cnts[i] + 1 > 255 ? 255 : cnts[i] + 1));
        EXPECT(counter.get_nb_distinct() == nd && counter.get_nb_unique() == 0);
    }
}

// NtHash trait (nthash.rs:76-120, kmer.rs:45-145) on the host value types against the batch kernel and the oracle
static void test_nthash_trait() {
    Sequence seq(S80, 2);
    const std::vector<Kmer16b32bit> v16 = KmerGenerator<Kmer16b32bit>(16).generate_kmer(seq);
    DeviceBatch b(std::vector<const Sequence*>{&seq});
    std::vector<uint64_t> h(v16.size() * 4);
    std::vector<uint8_t> strand(v16.size());
    check(kmu_nthash_canonical(Context::global().get(), b.get(), 16, 4, h.data(), strand.data(), 0), "kmu_nthash_canonical");
    for (size_t i = 0; i < v16.size(); ++i) {
        uint64_t f = 0, r = 0;
        const auto res = v16[i].nthash_canonical_init(f, r);
        EXPECT(res.first == h[4 * i] && res.second == strand[i]);
        EXPECT(v16[i].nthash_init() == f && f == orc_nthash_init(v16[i].v, 16, ORC_KMER16B32));
        std::vector<uint64_t> multi(4);
        EXPECT(v16[i].nthash_mult_canonical_init(f, r, multi) == strand[i]);
        for (int j = 0; j < 4; ++j) EXPECT(multi[j] == h[4 * i + j]);
    }
    // SURVEY Appendix C (derived from kmer.rs:74-94): window 0 of the 80-base string
    uint64_t f = 0, r = 0;
    EXPECT(v16[0].nthash_canonical_init(f, r).first == 0x684a2ec1114d51c5ull && f == 0x9840eab169670ddfull && r == 0x684a2ec1114d51c5ull);
    const std::vector<Kmer32bit> v8 = KmerGenerator<Kmer32bit>(8).generate_kmer(seq);
    EXPECT(v8[0].nthash_init() == 0x935533199c1dfb81ull && v8[1].nthash_init() == 0x4f6868cb4fb9a55eull);
    // the *_cycle methods are the reference's single-step shims, bug for bug: same values as the oracle's restatement
    for (size_t i = 0; i + 1 < v8.size(); i += 7) {
        Kmer32bit km = v8[i];
        const uint8_t nbase = (uint8_t)(v8[i + 1].v & 3);
        EXPECT(km.nthash_cycle(km.nthash_init(), nbase) == orc_nthash_cycle(km.v, 8, ORC_KMER32, orc_nthash_init(km.v, 8, ORC_KMER32), nbase));
        EXPECT(km == v8[i]);  // push(new_base) is dropped: self does not advance (kmer.rs:69)
        uint64_t f1 = 1, r1 = 2, f2 = 1, r2 = 2, canon = 0;
        const auto got = km.nthash_canonical_cycle(nbase, f1, r1);
        const int st = orc_nthash_canonical_cycle(km.v, 8, ORC_KMER32, nbase, &f2, &r2, &canon);
        EXPECT(f1 == f2 && r1 == r2 && got.first == canon && got.second == st);
    }
}

// KmerSeqIterator (kmergenerator.rs:30-107) and generate_weighted_kmer (:155-186) on the host layer
static void test_kmer_seq_iterator_and_distribution() {
    Sequence seq(S80, 2);
    const std::vector<Kmer16b32bit> all = KmerGenerator<Kmer16b32bit>(16).generate_kmer(seq);
    KmerSeqIterator<Kmer16b32bit> it(16, seq);
    size_t n = 0;
    while (auto km = it.next()) {
        EXPECT(n < all.size() && *km == all[n]);
        ++n;
    }
    EXPECT(n == all.size() && !it.next());
    // test_gen_kmer16b32bit_50bases_range_iterator (kmergenerator.rs:661-700): range 3..25
    EXPECT(it.set_range(3, 25));
    n = 0;
    while (auto km = it.next()) {
        const auto u = km->get_uncompressed_kmer();
        EXPECT(std::string(u.begin(), u.end()) == S80.substr(3 + n, 16));
        ++n;
    }
    EXPECT(n == 25 - 3 - 16 + 1);
    EXPECT(!it.set_range(10, 10) && !it.set_range(5, 81));  // Err(()) (sequence.rs:563-565)
    bool threw = false;
    try {
        KmerSeqIterator<Kmer32bit> bad(15, seq);  // kmergenerator.rs:48-53
    } catch (const Panic&) {
        threw = true;
    }
    EXPECT(threw);
    // a sequence longer than the iterator's window: every k-mer, in order, across the refills
    const std::string big = synth(5, (1u << 20) + 5000);
    Sequence sbig(big, 2);
    KmerSeqIterator<Kmer32bit> itb(11, sbig);
    const std::vector<Kmer32bit> wantb = KmerGenerator<Kmer32bit>(11).generate_kmer(sbig);
    n = 0;
    bool same = true;
    while (auto km = itb.next()) {
        same &= n < wantb.size() && *km == wantb[n];
        ++n;
    }
    EXPECT(same && n == wantb.size() && n == big.size() - 10);
    // test_generate_weighted_kmer32bit (kmergenerator.rs:776-850): 3-mers of the first 48 bases with their multiplicities
    Sequence s48(S80.substr(0, 48), 2);
    const auto dist = KmerGenerator<Kmer32bit>(3).generate_weighted_kmer(s48);
    const std::pair<const char*, uint32_t> want[] = {
        {"TCA", 4}, {"CAA", 2}, {"AAA", 4}, {"AAG", 1}, {"AGG", 1}, {"GGG", 1}, {"GGA", 1}, {"GAA", 1}, {"AAC", 1}, {"ACA", 1}, {"CAT", 1},
        {"ATT", 2}, {"TTC", 2}, {"AAT", 1}, {"ATC", 1}, {"CAG", 2}, {"AGT", 2}, {"GTA", 2}, {"TAT", 2}, {"ATG", 1}, {"TGC", 1}, {"GCG", 1},
        {"CGC", 1}, {"GCC", 1}, {"CCC", 1}, {"CCG", 1}, {"CGT", 2}, {"GTT", 2}, {"TTA", 1}, {"TAC", 1}, {"ACG", 1}};
    EXPECT(dist.size() == sizeof(want) / sizeof(want[0]));
    uint32_t total = 0;
    for (const auto& w : want) {
        Sequence one(std::string(w.first), 2);
        const Kmer32bit km = KmerGenerator<Kmer32bit>(3).generate_kmer(one)[0];
        const auto f = dist.find(km);
        EXPECT(f != dist.end() && f->second == w.second);
        total += w.second;
    }
    EXPECT(total == 46);
    EXPECT(hashmap_count_to_vec_count(dist).size() == dist.size());
    // the other two types: multiplicities sum to the number of k-mers
    uint64_t t16 = 0, t64 = 0;
    for (const auto& kv : KmerGenerator<Kmer16b32bit>(16).generate_weighted_kmer(seq)) t16 += kv.second;
    for (const auto& kv : KmerGenerator<Kmer64bit>(21).generate_kmer_distribution(seq)) t64 += kv.second;
    EXPECT(t16 == 65 && t64 == 60);
}

// sketch_seqrange_superminhash (seqminhash.rs:19-62) and the amino-acid SuperHashSketch (aautils/setsketchert.rs:203-329)
static void test_seqrange_and_aa_superminhash() {
    const std::string a = synth(17, 6000);
    Sequence seq(a, 2);
    for (const int k : {12, 16}) {
        const size_t lo = 500, hi = 4100, m = 300;
        const std::vector<double> got = sketch_seqrange_superminhash(seq, lo, hi, k, m);
        Sequence sub(a.substr(lo, hi - lo), 2);
        const uint64_t off0 = 0, nb0 = hi - lo;
        std::vector<double> want(m);
        orc_sketch_superminhash_batch(sub.packed().data(), &off0, &nb0, 1, k, k == 16 ? ORC_KMER16B32 : ORC_KMER32, ORC_HASH_CANON_INVHASH, m, 0, 8,
                                      want.data(), 1);
        EXPECT(got == want);
    }
    bool threw = false;
    try {
        sketch_seqrange_superminhash(seq, 0, 100, 8, 100);  // "unimplemented kmer_size" (seqminhash.rs:55-60)
    } catch (const Panic&) {
        threw = true;
    }
    EXPECT(threw);
    using namespace kmerutils::aautils;
    const std::string p1 = "MTEQIELIKLYSTRILALAAQMPHVGSLDNPDASAMKRSPLCGSKVTVDVIMQNGKITFDGFEVLAPASEYKNRHASILLSLDATAEACASIAAQNSA";
    const std::string p2 = "MTEQIELIKLYSTRILALAAQMPHVGSLDNPDASAMKRSPLCGSKVMTEQIELIKLYSTRILALAAQMPHVGSLDNPDASAMKRSPLCGSKV";
    SequenceAA s1(p1), s2(p2);
    aautils::SuperHashSketch<KmerAA64bit, double> sh({5, 200});
    const SeqSketcherAAT<KmerAA64bit, double>& tr = sh;
    const auto sig = tr.sketch_compressedkmeraa({&s1, &s2}, KmerHash::masked_value());
    const SequenceAA* both[2] = {&s1, &s2};
    for (int i = 0; i < 2; ++i) {
        const uint64_t off0 = 0, n0 = both[i]->size();
        std::vector<double> want(200);
        orc_sketch_superminhash_batch(both[i]->residues().data(), &off0, &n0, 1, 5, ORC_KMERAA64, ORC_HASH_MASKED_VALUE, 200, 0, 8, want.data(), 1);
        EXPECT(sig[i] == want);
    }
    // one signature for the collection = element-wise minimum of the per-sequence ones
    const auto whole = sh.sketch_compressedkmeraa_seqs({&s1, &s2}, KmerHash::masked_value());
    EXPECT(whole.size() == 1);
    for (size_t j = 0; j < 200 && whole.size() == 1; ++j) EXPECT(whole[0][j] == std::min(sig[0][j], sig[1][j]));
}

static void test_pminhasha_kmer_smallb() {
    const size_t kmer_size = 5, sketch_size = 4000;
    Sequence seqa(S80, 2);
    std::vector<Sequence> vecseqb;
    vecseqb.push_back(Sequence(S80.substr(0, 40), 2));
    vecseqb.push_back(seqa.get_reverse_complement());
    const double jac_theo_0 = double(40 - kmer_size) / double(80 - kmer_size);
    auto vecsig = jaccard_index_probminhash3a<Kmer32bit>(seqa, vecseqb, sketch_size, kmer_size, KmerHash::canonical_invhash());
    EXPECT(vecsig[0] >= 0.75 * jac_theo_0);  // seqsketchjaccard.rs:784
    EXPECT(vecsig[1] >= 1.);                 // :785
    vecsig = jaccard_index_probminhash3a<Kmer32bit>(seqa, vecseqb, sketch_size, kmer_size, KmerHash::identity());
    EXPECT(vecsig[0] >= 0.75 * jac_theo_0);  // :791

    // bit-exact signatures against the oracle, ragged lengths (shorter than k included), three k-mer types
    const std::string big = synth(11, 60000);
    std::vector<std::string> reads = {big.substr(0, 1000), big.substr(1000, 37), big.substr(2000, 7), big.substr(3000, 5000),
                                      big.substr(9000, 8), big.substr(10000, 20000), big.substr(30000, 999), S80};
    const std::vector<Sequence> seqs = Sequence::new_batch(reads, 2);
    struct Case { int k, type, hash; };
    for (const Case c : {Case{8, ORC_KMER32, ORC_HASH_CANON_INVHASH}, Case{16, ORC_KMER16B32, ORC_HASH_CANON_INVHASH}, Case{21, ORC_KMER64, ORC_HASH_CANON_INVHASH}}) {
        const uint32_t m = 200;
        SeqSketcher sk(c.k, m);
        for (size_t i = 0; i < seqs.size(); ++i) {
            std::vector<uint64_t> want(m);
            orc_sketch_pmh3a_seq(seqs[i].packed().data(), seqs[i].size(), c.k, c.type, c.hash, m, want.data());
            std::vector<uint64_t> got(m);
            if (c.type == ORC_KMER32) {
                auto s = sk.sketch_probminhash3a<Kmer32bit>({&seqs[i]}, KmerHash::canonical_invhash())[0];
                got.assign(s.begin(), s.end());
            } else if (c.type == ORC_KMER16B32) {
                auto s = sk.sketch_probminhash3a<Kmer16b32bit>({&seqs[i]}, KmerHash::canonical_invhash())[0];
                got.assign(s.begin(), s.end());
            } else {
                got = sk.sketch_probminhash3a<Kmer64bit>({&seqs[i]}, KmerHash::canonical_invhash())[0];
            }
            EXPECT(got == want);
        }
    }
    // whole-file form through the trait object
    ProbHash3aSketch<Kmer32bit> ph({8, 200});
    const SeqSketcherT<Kmer32bit, uint32_t>& tr = ph;
    const auto whole = tr.sketch_compressedkmer_seqs(as_refs(seqs), KmerHash::canonical_invhash());
    std::vector<uint8_t> packed;
    std::vector<uint64_t> off, nb;
    for (const Sequence& s : seqs) {
        while (packed.size() % 16) packed.push_back(0);
        off.push_back(packed.size());
        nb.push_back(s.size());
        packed.insert(packed.end(), s.packed().begin(), s.packed().end());
    }
    std::vector<uint64_t> want(200);
    orc_sketch_pmh3a_seqs(packed.data(), off.data(), nb.data(), seqs.size(), 8, ORC_KMER32, ORC_HASH_CANON_INVHASH, 200, want.data());
    EXPECT(whole.size() == 1 && std::vector<uint64_t>(whole[0].begin(), whole[0].end()) == want);
    // an empty sequence panics in the reference (set_range(..).unwrap(), seqsketchjaccard.rs:230)
    bool threw = false;
    try {
        Sequence empty = Sequence::from_packed({}, 0);
        SeqSketcher(8, 200).sketch_probminhash3a<Kmer32bit>({&empty}, KmerHash::identity());
    } catch (const Panic&) {
        threw = true;
    }
    EXPECT(threw);
}

static void test_superminhash_and_hll() {
    // test_superminhash_kmer_16b32bit_serial (seqsketchjaccard.rs:947-1004): J(a, first half of a) and J(a, a)
    const std::string a = synth(21, 4000);
    const std::vector<Sequence> seqs = Sequence::new_batch({a, a.substr(0, 2000), a}, 2);
    const size_t m = 800;
    SeqSketcher sk(16, m);
    const auto sig = sk.sketch_superminhash<Kmer16b32bit, double>(as_refs(seqs), KmerHash::canonical_invhash());
    const double j_half = compute_probminhash_jaccard(sig[0], sig[1]), j_same = compute_probminhash_jaccard(sig[0], sig[2]);
    const double theo = double(2000 - 16 + 1) / double(4000 - 16 + 1);
    EXPECT(j_same == 1.0);
    EXPECT(j_half > 0.75 * theo && j_half < 1.25 * theo);
    // bit-exact against the oracle (FNV key hasher here, NoHashHasher through SuperHashSketch)
    std::vector<uint8_t> packed;
    std::vector<uint64_t> off, nb;
    for (const Sequence& s : seqs) {
        while (packed.size() % 16) packed.push_back(0);
        off.push_back(packed.size());
        nb.push_back(s.size());
        packed.insert(packed.end(), s.packed().begin(), s.packed().end());
    }
    std::vector<double> want(seqs.size() * m);
    orc_sketch_superminhash_batch(packed.data(), off.data(), nb.data(), seqs.size(), 16, ORC_KMER16B32, ORC_HASH_CANON_INVHASH, m, 1, 8, want.data(), 2);
    for (size_t i = 0; i < seqs.size(); ++i) EXPECT(std::vector<double>(want.begin() + i * m, want.begin() + (i + 1) * m) == sig[i]);
    SuperHashSketch<Kmer16b32bit, float> sh({16, m});
    const auto sigf = sh.sketch_compressedkmer(as_refs(seqs), KmerHash::canonical_invhash());
    std::vector<float> wantf(seqs.size() * m);
    orc_sketch_superminhash_batch(packed.data(), off.data(), nb.data(), seqs.size(), 16, ORC_KMER16B32, ORC_HASH_CANON_INVHASH, m, 0, 4, wantf.data(), 2);
    for (size_t i = 0; i < seqs.size(); ++i) EXPECT(std::vector<float>(wantf.begin() + i * m, wantf.begin() + (i + 1) * m) == sigf[i]);
    // HyperLogLogSketch: per sequence and whole-file registers
    SetSketchParams prm;
    HyperLogLogSketch<Kmer32bit, uint16_t> hll({12, 256}, prm);
    const auto regs = hll.sketch_compressedkmer(as_refs(seqs), KmerHash::canonical_invhash());
    std::vector<uint16_t> wr(seqs.size() * 256);
    orc_sketch_setsketch_batch(packed.data(), off.data(), nb.data(), seqs.size(), 12, ORC_KMER32, ORC_HASH_CANON_INVHASH, prm.b, 256, prm.a, prm.q, 2, wr.data(), 2);
    for (size_t i = 0; i < seqs.size(); ++i) EXPECT(std::vector<uint16_t>(wr.begin() + i * 256, wr.begin() + (i + 1) * 256) == regs[i]);
    const auto whole = hll.sketch_compressedkmer_seqs(as_refs(seqs), KmerHash::canonical_invhash());
    std::vector<uint16_t> ww(256);
    orc_sketch_setsketch(packed.data(), off.data(), nb.data(), seqs.size(), 12, ORC_KMER32, ORC_HASH_CANON_INVHASH, prm.b, 256, prm.a, prm.q, 2, ww.data());
    EXPECT(whole.size() == 1 && whole[0] == ww);
}

static void test_reload_sketch_file(const std::string& dir) {
    const std::string big = synth(31, 12000);
    const std::vector<Sequence> seqs = Sequence::new_batch({big.substr(0, 3000), big.substr(3000, 4000), big.substr(7000, 5000)}, 2);
    SeqSketcher sk(8, 200);
    const auto sigs = sk.sketch_probminhash3a<Kmer32bit>(as_refs(seqs), KmerHash::canonical_invhash());
    const std::string fname = dir + "/sigs.bin";
    kmu_sigdump* out = sk.create_signature_dump(fname);
    dump_signatures_block_u32(sigs, out);
    check(kmu_sigdump_close(out), "close");
    SigSketchFileReader reader(fname);
    EXPECT(reader.get_kmer_size() == 8 && reader.get_signature_length() == 200 && reader.get_signature_size() == 4);
    size_t n = 0;
    while (auto s = reader.next()) {
        EXPECT(n < sigs.size() && *s == sigs[n]);
        ++n;
    }
    EXPECT(n == sigs.size());
}

static void test_block_32bit_sketch() {
    // seqblocksketch.rs:458-496
    Sequence seqa(std::string("TCAAAGGGAAACATTCAAAATCAGTATGCGCCCGTTCAGTTACGTATTGCTCTCGCCGTAGGCCTAATGAGATGGGCTGGGTACAGAG"), 2);
    Sequence seqb(std::string("TCAAAGGGAAATTTTTTTCATTCAAAATCAGTATGCGCCCGTTCAGTTACGTATTGCTCTCGCCGTAGGCCTAATGATTTTTTTGATGGGCTGGGTACAGAG"), 2);
    const size_t block_size = 10, kmer_size = 3, sketch_size = 6;
    BlockSeqSketcher sketcher(block_size, kmer_size, sketch_size);
    const BlockSketchedSeq sketcha = sketcher.blocksketch_sequence(1, seqa, KmerHash::canonical_invhash());
    const BlockSketchedSeq sketchb = sketcher.blocksketch_sequence(2, seqb, KmerHash::canonical_invhash());
    EXPECT(sketcha.sketch.size() == (seqa.size() + 9) / 10 && sketchb.sketch.size() == (seqb.size() + 9) / 10);
    DistBlockSketched mydist;
    EXPECT(mydist.eval(sketcha.sketch[0], sketcha.sketch[0]) == 1.f);  // same sequence: 1 (:487)
    const float dist_1 = mydist.eval(sketcha.sketch[0], sketchb.sketch[0]);  // the reference prints these two
    const float dist_2 = mydist.eval(sketcha.sketch[1], sketchb.sketch[1]);
    EXPECT(dist_1 >= 0.f && dist_1 <= 1.f && dist_2 >= 0.f && dist_2 <= 1.f);
    // blocks 3 and 4 of seqa (bases 30..49) are blocks 4 and 5 of seqb shifted by the 7 inserted T's: not aligned, so no
    // equality is expected there; the last full blocks before the first insertion agree only up to position 9's k-mer
    // every block against the oracle (blocks of block_size k-mers; the oracle takes the packed sequence)
    std::vector<uint32_t> want(sketcha.sketch.size() * sketch_size);
    const uint64_t nb = orc_blocksketch_seq(seqa.packed().data(), seqa.size(), (int)kmer_size, (uint32_t)sketch_size, block_size, want.data(),
                                            sketcha.sketch.size());
    EXPECT(nb == sketcha.sketch.size());
    for (size_t b = 0; b < sketcha.sketch.size() && b < nb; ++b) {
        EXPECT(sketcha.sketch[b].size() == 1 && sketcha.sketch[b][0].numblock == b && sketcha.sketch[b][0].numseq == 1);
        EXPECT(std::vector<uint32_t>(want.begin() + b * sketch_size, want.begin() + (b + 1) * sketch_size) == sketcha.sketch[b][0].get_skech_slice());
    }
}

static void test_amino_acids() {
    using namespace kmerutils::aautils;
    const std::string prot =
        "MTEQIELIKLYSTRILALAAQMPHVGSLDNPDASAMKRSPLCGSKVTVDVIMQNGKITEFAQNVKACALGQAAASVAAQNIIGRTAEEVVRARDELAAMLKSGGPPPGPPFDGFEVLAPASEYKNRHASILLSLDATAEACASIAAQNSA";
    SequenceAA seqaa(prot);
    // test_seqaa_32bit_iterator_range (kmeraa.rs:920-956): 4-mers of the range 3..10 are QIEL IELI ELIK LIKL
    const auto k4 = aautils::KmerGenerator<KmerAA32bit>(4).generate_kmer_in_range(seqaa, 3, 10);
    EXPECT(k4.size() == 4);
    const char* want4[4] = {"QIEL", "IELI", "ELIK", "LIKL"};
    for (size_t i = 0; i < k4.size() && i < 4; ++i) {
        const auto u = k4[i].get_uncompressed_kmer();
        EXPECT(std::string(u.begin(), u.end()) == want4[i]);
    }
    // test_seqaa_iterator_end (kmeraa.rs:997-1021): the last 8-mer of the first 32 residues is VGSLDNPD
    SequenceAA head(prot.substr(0, 32));
    const auto k8 = aautils::KmerGenerator<KmerAA64bit>(8).generate_kmer(head);
    EXPECT(k8.size() == 25);
    const auto last = k8.back().get_uncompressed_kmer();
    EXPECT(std::string(last.begin(), last.end()) == "VGSLDNPD");
    // the value type agrees with the generated words: push the next residue by hand
    KmerAA64bit km = k8[0];
    for (size_t i = 1; i < k8.size(); ++i) {
        km = km.push((uint8_t)prot[7 + i]);
        EXPECT(km == k8[i]);
    }
    // an invalid residue panics (Alphabet::encode, kmeraa.rs:106); new_filtered drops it (:447-456)
    bool threw = false;
    try {
        SequenceAA bad(std::string("MTEQXIELIK"));
        aautils::KmerGenerator<KmerAA32bit>(4).generate_kmer(bad);
    } catch (const Panic&) {
        threw = true;
    }
    EXPECT(threw);
    EXPECT(SequenceAA::new_filtered("MTxEQ*IBLKZ", Alphabet()).to_string() == "MTEQILK");
    // test_seqaa_probminhash_64bit (setsketchert.rs:1218-1266): the second string is the first half of the first, twice
    const std::string str1 = "MTEQIELIKLYSTRILALAAQMPHVGSLDNPDASAMKRSPLCGSKVTVDVIMQNGKITFDGFEVLAPASEYKNRHASILLSLDATAEACASIAAQNSA";
    const std::string str2 = "MTEQIELIKLYSTRILALAAQMPHVGSLDNPDASAMKRSPLCGSKVMTEQIELIKLYSTRILALAAQMPHVGSLDNPDASAMKRSPLCGSKV";
    SequenceAA seq1(str1), seq2(str2);
    aautils::SeqSketcher sketcher(5, 400);
    const auto sig = sketcher.sketch_probminhash3a<KmerAA64bit>({&seq1, &seq2}, KmerHash::masked_value());
    const double dist = sketching::compute_probminhash_jaccard(sig[0], sig[1]);
    EXPECT(std::abs(dist - 0.5) < 1. / 10.);  // setsketchert.rs:1264
    // bit-exact against the oracle (the oracle takes the residues as ASCII), trait object, 32-bit k-mers, SetSketch
    for (const SequenceAA* sq : {&seq1, &seq2, &seqaa}) {
        std::vector<uint64_t> want(400);
        orc_sketch_pmh3a_seq(sq->residues().data(), sq->size(), 5, ORC_KMERAA64, ORC_HASH_MASKED_VALUE, 400, want.data());
        aautils::ProbHash3aSketch<KmerAA64bit> ph({5, 400});
        const aautils::SeqSketcherAAT<KmerAA64bit, uint64_t>& tr = ph;
        EXPECT(tr.sketch_compressedkmeraa({sq}, KmerHash::masked_value())[0] == want);
        std::vector<uint64_t> want32(100);
        orc_sketch_pmh3a_seq(sq->residues().data(), sq->size(), 6, ORC_KMERAA32, ORC_HASH_MASKED_VALUE, 100, want32.data());
        const auto got32 = aautils::SeqSketcher(6, 100).sketch_probminhash3a<KmerAA32bit>({sq}, KmerHash::masked_value())[0];
        EXPECT(std::vector<uint64_t>(got32.begin(), got32.end()) == want32);
    }
    sketching::SetSketchParams prm;
    aautils::HyperLogLogSketch<KmerAA64bit, uint16_t> hll({12, 128}, prm);
    const auto regs = hll.sketch_compressedkmeraa({&seqaa}, KmerHash::masked_value());
    const uint64_t off0 = 0, nres = seqaa.size();
    std::vector<uint16_t> wr(128);
    orc_sketch_setsketch_batch(seqaa.residues().data(), &off0, &nres, 1, 12, ORC_KMERAA64, ORC_HASH_MASKED_VALUE, prm.b, 128, prm.a, prm.q, 2, wr.data(), 1);
    EXPECT(regs[0] == wr);
}

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    try {
        test_gen_kmer16b32bit_80bases();
        test_gen_kmer_range_and_types();
        test_kmer32bit_revcomp_order();
        test_kmer_counter();
        test_kmer32bit_counter_keys();
        test_nthash_trait();
        test_kmer_seq_iterator_and_distribution();
        test_seqrange_and_aa_superminhash();
        test_pminhasha_kmer_smallb();
        test_superminhash_and_hll();
        test_reload_sketch_file(dir);
        test_amino_acids();
        test_block_32bit_sketch();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
    if (failures) {
        std::fprintf(stderr, "%d expectation(s) failed\n", failures);
        return 1;
    }
    std::printf("host api ok: %llu kernel launches\n", (unsigned long long)Context::global().launch_count());
    return 0;
}
