//! Fixtures of the real crate, in the layout of tests/golden/oracle_fixtures.json (little-endian bytes as hex).
//! Inputs: the 80-base test string of the reference's unit tests and six synthetic reads cut from the SplitMix64
//! stream of seed 1 (base i = top two bits of output i), exactly what kmu_seqbatch_synth builds.
use kmerutils::base::kmer16b32bit::Kmer16b32bit;
use kmerutils::base::kmer32bit::Kmer32bit;
use kmerutils::base::kmer64bit::Kmer64bit;
use kmerutils::base::kmergenerator::{KmerGenerationPattern, KmerGenerator};
use kmerutils::base::kmertraits::{CompressedKmerT, KmerT};
use kmerutils::base::sequence::Sequence;
use kmerutils::sketcharg::{DataType, SeqSketcherParams, SketchAlgo};
use kmerutils::sketching::seqsketchjaccard::SeqSketcher;
use kmerutils::sketching::setsketchert::{
    HllSeqsThreading, HyperLogLogSketch, ProbHash3aSketch, SeqSketcherT, SuperHashSketch,
};
use probminhash::invhash::{int32_hash, int64_hash};
use probminhash::setsketcher::SetSketchParams;
use serde_json::{json, Map, Value};

const S80: &str = "TCAAAGGGAAACATTCAAAATCAGTATGCGCCCGTTCAGTTACGTATTGCTCTCGCTAATGAGATGGGCTGGGTACAGAG";
const LENGTHS: [usize; 6] = [1000, 37, 8, 150, 5000, 2500];
const SEED: u64 = 1;

fn synth_z(seed: u64, i: u64) -> u64 {
    let mut z = seed.wrapping_add((i + 1).wrapping_mul(0x9E3779B97F4A7C15));
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
    z ^ (z >> 31)
}

fn synth_reads() -> Vec<Sequence> {
    let mut first = 0u64;
    let mut out = Vec::new();
    for len in LENGTHS {
        let ascii: Vec<u8> = (0..len as u64).map(|j| b"ACGT"[(synth_z(SEED, first + j) >> 62) as usize]).collect();
        out.push(Sequence::new(&ascii, 2));
        first += len as u64;
    }
    out
}

fn hex<T: Copy, const N: usize>(vals: &[T], to: fn(T) -> [u8; N]) -> String {
    vals.iter().flat_map(|v| to(*v)).map(|b| format!("{:02x}", b)).collect()
}
fn rows_u32(sig: &[Vec<u32>]) -> Value { json!(sig.iter().map(|r| hex(r, u32::to_le_bytes)).collect::<Vec<_>>()) }
fn rows_u64(sig: &[Vec<u64>]) -> Value { json!(sig.iter().map(|r| hex(r, u64::to_le_bytes)).collect::<Vec<_>>()) }
fn rows_u16(sig: &[Vec<u16>]) -> Value { json!(sig.iter().map(|r| hex(r, u16::to_le_bytes)).collect::<Vec<_>>()) }
fn rows_f64(sig: &[Vec<f64>]) -> Value { json!(sig.iter().map(|r| hex(r, f64::to_le_bytes)).collect::<Vec<_>>()) }
fn rows_f32(sig: &[Vec<f32>]) -> Value { json!(sig.iter().map(|r| hex(r, f32::to_le_bytes)).collect::<Vec<_>>()) }

fn main() {
    let reads = synth_reads();
    // the reference unwraps set_range(0, size) on every sequence: leave the reads shorter than k out per configuration
    let refs = |k: usize| -> (Vec<&Sequence>, Vec<usize>) {
        let idx: Vec<usize> = (0..reads.len()).filter(|&i| reads[i].size() >= k).collect();
        (idx.iter().map(|&i| &reads[i]).collect(), idx)
    };
    let canon32 = |kmer: &Kmer32bit| -> u32 { int32_hash(kmer.reverse_complement().min(*kmer).0) };
    let canon16 = |kmer: &Kmer16b32bit| -> u32 { int32_hash(kmer.reverse_complement().min(*kmer).0) };
    let canon64 = |kmer: &Kmer64bit| -> u64 { int64_hash(kmer.reverse_complement().min(*kmer).0) };
    let mut out = Map::new();
    out.insert("seed".into(), json!(SEED));
    out.insert("lengths".into(), json!(LENGTHS));

    // ---- k-mer words of the 80-base string ----
    let s80 = Sequence::new(S80.as_bytes(), 2);
    let k8: Vec<u32> = KmerGenerator::<Kmer32bit>::new(8).generate_kmer(&s80).iter().map(canon32).collect();
    let k16: Vec<u32> = KmerGenerator::<Kmer16b32bit>::new(16).generate_kmer(&s80).iter().map(|k| k.0).collect();
    let k31: Vec<u64> = KmerGenerator::<Kmer64bit>::new(31).generate_kmer(&s80).iter().map(|k| k.reverse_complement().min(*k).0).collect();
    out.insert("s80_kmers".into(), json!({
        "k8_kmer32_canon_invhash": hex(&k8, u32::to_le_bytes),
        "k16_kmer16b32_raw": hex(&k16, u32::to_le_bytes),
        "k31_kmer64_canon": hex(&k31, u64::to_le_bytes),
    }));

    // ---- ProbMinHash3a (rows of reads shorter than k are absent: see "rows") ----
    let mut pmh = Map::new();
    let (v8, i8_) = refs(8);
    pmh.insert("k8_kmer32_m64".into(), rows_u32(&SeqSketcher::new(8, 64).sketch_probminhash3a(&v8, canon32)));
    let (v16, i16_) = refs(16);
    pmh.insert("k16_kmer16b32_m64".into(), rows_u32(&SeqSketcher::new(16, 64).sketch_probminhash3a(&v16, canon16)));
    let (v21, i21) = refs(21);
    pmh.insert("k21_kmer64_m64".into(), rows_u64(&SeqSketcher::new(21, 64).sketch_probminhash3a(&v21, canon64)));
    let p3a = SeqSketcherParams::new(8, 64, SketchAlgo::PROB3A, DataType::DNA);
    let whole = ProbHash3aSketch::<Kmer32bit>::new(&p3a).sketch_compressedkmer_seqs(&v8, canon32);
    pmh.insert("whole_k8_kmer32_m64".into(), json!(hex(&whole[0], u32::to_le_bytes)));
    out.insert("pmh3a".into(), Value::Object(pmh));

    // ---- SuperMinHash: NoHashHasher through SuperHashSketch, FnvHasher through SeqSketcher ----
    let psup = SeqSketcherParams::new(8, 64, SketchAlgo::SUPER, DataType::DNA);
    let smh = SuperHashSketch::<Kmer32bit, f64>::new(&psup).sketch_compressedkmer(&v8, canon32);
    let smh32 = SeqSketcher::new(16, 64).sketch_superminhash::<Kmer16b32bit, f32, _>(&v16, canon16);
    out.insert("superminhash".into(), json!({
        "k8_kmer32_m64_f64_nohash": rows_f64(&smh),
        "k16_kmer16b32_m64_f32_fnv": rows_f32(&smh32),
    }));

    // ---- SetSketch: b = 1.001, m = 64, a = 20, q = 2^16 - 2 ----
    // SetSketchParams lives in probminhash (not vendored with the reference): constructor assumed (b, m, a, q)
    let hllp = SetSketchParams::new(1.001, 64, 20., 65534);
    let phll = SeqSketcherParams::new(8, 64, SketchAlgo::HLL, DataType::DNA);
    let hll = HyperLogLogSketch::<Kmer32bit, u16>::new(&phll, hllp, HllSeqsThreading::default()).sketch_compressedkmer(&v8, canon32);
    let phll21 = SeqSketcherParams::new(21, 64, SketchAlgo::HLL, DataType::DNA);
    let hllp21 = SetSketchParams::new(1.001, 64, 20., 65534);
    let hll21 = HyperLogLogSketch::<Kmer64bit, u16>::new(&phll21, hllp21, HllSeqsThreading::default())
        .sketch_compressedkmer_seqs(&v21, canon64);
    out.insert("setsketch".into(), json!({
        "params": [1.001, 64, 20.0, 65534],
        "k8_kmer32_u16": rows_u16(&hll),
        "whole_k21_kmer64_u16": hex(&hll21[0], u16::to_le_bytes),
    }));
    out.insert("rows".into(), json!({ "k8": i8_, "k16": i16_, "k21": i21 }));
    println!("{}", serde_json::to_string_pretty(&Value::Object(out)).unwrap());
}
