//! Reference `src/aautils/kmeraa.rs`: `Alphabet` (:29-135), `KmerAA32bit` (:147-277), `KmerAA64bit` (:281-400),
//! `SequenceAA` (:404-486), `KmerSeqIterator` (:498-629), `KmerGenerator` (:646-916).  Residues are encoded on the GPU
//! (`kmu_seqbatch_from_aa`, one byte per residue in HBM) and k-mers come from `kmu_generate_kmers` with `KMU_KMERAA32/64`.
use crate::base::kmertraits::*;
use crate::devhash::RawWord;
use crate::ffi;
use fnv::FnvHashMap;
use std::io;
use std::marker::PhantomData;
use std::os::raw::c_void;
use std::str::FromStr;

/// codes 1..=21 in alphabetical order of the 20 residues, 0b01110 left out (Q = 0b01111, :44-46)
const RESIDUES: &[u8; 20] = b"ACDEFGHIKLMNPQRSTVWY";

pub struct Alphabet {
    pub bases: String,
}
impl Default for Alphabet { fn default() -> Self { Self::new() } }
impl Alphabet {
    pub fn new() -> Alphabet { Alphabet { bases: String::from_utf8(RESIDUES.to_vec()).unwrap() } }
    pub fn len(&self) -> u8 { RESIDUES.len() as u8 }
    pub fn is_valid_base(&self, c: u8) -> bool { RESIDUES.contains(&c) }
    pub fn get_nb_bits(&self) -> u8 { 5 }
    pub(crate) fn encode(&self, c: u8) -> u8 {
        match RESIDUES.iter().position(|r| *r == c) {
            Some(i) => (i as u8 + 1) + (i >= 13) as u8,
            None => panic!("encode: not a code in alpahabet for amino acid: {:x}", c),
        }
    }
    pub(crate) fn decode(&self, c: u8) -> u8 {
        match c {
            1..=13 => RESIDUES[c as usize - 1],
            15..=21 => RESIDUES[c as usize - 2],
            _ => panic!("decode : pattern not a code in alpahabet for Amino Acid got : {:#b}", c & 0b11111),
        }
    }
}

macro_rules! kmer_aa {
    ($name:ident, $val:ty, $maxb:expr, $kmu:expr, $msg:expr) => {
        #[derive(Copy, Clone, Hash, Debug)]
        pub struct $name {
            aa: $val,
            nb_base: u8,
        }
        impl $name {
            pub fn new(nb_base: u8) -> Self {
                if nb_base as usize >= $maxb { panic!($msg) } // `>=` as in the reference (:155, :288)
                $name { aa: 0, nb_base }
            }
        }
        impl KmerT for $name {
            fn get_nb_base(&self) -> u8 { self.nb_base }
            fn push(&self, c: u8) -> Self {
                let value_mask: $val = ((1 as $val) << (5 * self.nb_base)) - 1;
                $name { aa: ((self.aa << 5) & value_mask) | (Alphabet::new().encode(c) as $val & 0b11111), nb_base: self.nb_base }
            }
            fn reverse_complement(&self) -> Self { panic!(concat!(stringify!($name), " reverse_complement not yet implemented")) }
            fn dump(&self, bufw: &mut dyn io::Write) -> io::Result<usize> {
                bufw.write_all(&[self.nb_base])?;
                bufw.write(&self.aa.to_ne_bytes())
            }
        }
        impl PartialEq for $name { fn eq(&self, o: &Self) -> bool { self.aa == o.aa && self.nb_base == o.nb_base } }
        impl Eq for $name {}
        impl Ord for $name { fn cmp(&self, o: &Self) -> std::cmp::Ordering { self.nb_base.cmp(&o.nb_base).then(self.aa.cmp(&o.aa)) } }
        impl PartialOrd for $name { fn partial_cmp(&self, o: &Self) -> Option<std::cmp::Ordering> { Some(self.cmp(o)) } }
        impl CompressedKmerT for $name {
            type Val = $val;
            const KMU_TYPE: i32 = $kmu;
            fn get_nb_base_max() -> usize { <$val>::BITS as usize / 5 }
            fn get_compressed_value(&self) -> $val { self.aa }
            fn get_uncompressed_kmer(&self) -> Vec<u8> {
                let alphabet = Alphabet::new();
                (0..self.nb_base).rev().map(|i| alphabet.decode(((self.aa >> (5 * i)) & 0b11111) as u8)).collect()
            }
            fn get_bitsize(&self) -> usize { <$val>::BITS as usize }
        }
        impl KmerBuilder<$name> for $name { fn build(val: $val, nb_base: u8) -> $name { $name { aa: val, nb_base } } }
        impl RawWord for $name {
            fn raw(&self) -> $val { self.aa }
            fn invhash(v: $val) -> $val { <$name as AaInvHash>::h(v) }
            fn value_mask(&self) -> $val { ((1 as $val) << (5 * self.nb_base)) - 1 }
        }
    };
}
trait AaInvHash: CompressedKmerT { fn h(v: Self::Val) -> Self::Val; }
kmer_aa!(KmerAA32bit, u32, 6, ffi::KMU_KMERAA32, "For KmerAA32bit nb_base must be less or equal to 6");
kmer_aa!(KmerAA64bit, u64, 12, ffi::KMU_KMERAA64, "For KmerAA64bit nb_base must be less or equal to 12");
impl AaInvHash for KmerAA32bit { fn h(v: u32) -> u32 { crate::devhash::int32_hash(v) } }
impl AaInvHash for KmerAA64bit { fn h(v: u64) -> u64 { crate::devhash::int64_hash(v) } }

/// one residue per byte, ASCII (:404-486)
pub struct SequenceAA {
    seq: Vec<u8>,
}
impl SequenceAA {
    pub fn new(str: &[u8]) -> Self { SequenceAA { seq: str.to_vec() } } // the reference's validity check is a lazy iterator never run (:412-418)
    pub fn len(&self) -> usize { self.seq.len() }
    pub fn is_empty(&self) -> bool { self.seq.is_empty() }
    pub fn size(&self) -> usize { self.seq.len() }
    pub fn get_base(&self, pos: usize) -> u8 { self.seq[pos] }
    pub fn new_filtered(buf: &[u8], alphabet: &Alphabet) -> Self { SequenceAA { seq: buf.iter().copied().filter(|b| alphabet.is_valid_base(*b)).collect() } }
    pub fn as_bytes(&self) -> &[u8] { &self.seq }
}
impl FromStr for SequenceAA {
    type Err = std::convert::Infallible;
    fn from_str(s: &str) -> Result<Self, Self::Err> { Ok(SequenceAA { seq: s.as_bytes().to_vec() }) }
}
impl ToString for SequenceAA { fn to_string(&self) -> String { String::from_utf8(self.seq.clone()).unwrap() } }

/// `&[&SequenceAA]` -> device batch of 5-bit codes (kmu_seqbatch_from_aa; a residue outside the alphabet is an error there)
pub fn device_batch_aa(vseq: &[&SequenceAA]) -> ffi::DeviceBatch {
    let mut off = vec![0u64; vseq.len() + 1];
    let mut ascii = Vec::new();
    for (i, s) in vseq.iter().enumerate() {
        ascii.extend_from_slice(s.as_bytes());
        off[i + 1] = ascii.len() as u64;
    }
    ascii.push(0);
    let mut b = std::ptr::null_mut();
    ffi::check(unsafe { ffi::kmu_seqbatch_from_aa(ffi::ctx(), ascii.as_ptr(), off.as_ptr(), vseq.len() as u64, 0, std::ptr::null_mut(), &mut b) }, "SequenceAA");
    ffi::DeviceBatch(b)
}

fn kmers_in_range<T: CompressedKmerT + KmerBuilder<T>>(seq: &SequenceAA, k: usize, begin: usize, end: usize) -> Vec<T> {
    if end < begin + k { return Vec::new(); }
    let part = SequenceAA { seq: seq.seq[begin..end].to_vec() };
    let b = device_batch_aa(&[&part]);
    let n = unsafe { ffi::kmu_kmer_count(b.0, k as u32) } as usize;
    let mut vals = vec![T::Val::default(); n];
    ffi::check(unsafe { ffi::kmu_generate_kmers(ffi::ctx(), b.0, k as u32, T::KMU_TYPE, ffi::KMU_HASH_IDENTITY_RAW, vals.as_mut_ptr() as *mut c_void,
                                                std::ptr::null_mut(), 0) }, "aautils::KmerGenerator::generate_kmer");
    vals.into_iter().map(|v| T::build(v, k as u8)).collect()
}

pub trait KmerSeqIteratorT {
    type KmerVal;
    fn next(&mut self) -> Option<Self::KmerVal>;
}

/// (:498-629) streams the k-mers of a range; the range is generated on the GPU at the first `next`
pub struct KmerSeqIterator<'a, T: CompressedKmerT> {
    nb_base: usize,
    sequence: &'a SequenceAA,
    range: std::ops::Range<usize>,
    window: Option<std::vec::IntoIter<T>>,
}
impl<'a, T: CompressedKmerT + KmerBuilder<T>> KmerSeqIterator<'a, T> {
    pub fn new(kmer_size: usize, seq: &'a SequenceAA) -> Self { KmerSeqIterator { nb_base: kmer_size, sequence: seq, range: 0..seq.len(), window: None } }
    pub fn set_range(&mut self, first: usize, last: usize) -> Result<(), String> {
        if last <= first || last > self.sequence.len() { return Err("bad range for iterator".to_string()); }
        self.range = first..last;
        self.window = None;
        Ok(())
    }
}
impl<'a, T: CompressedKmerT + KmerBuilder<T>> KmerSeqIteratorT for KmerSeqIterator<'a, T> {
    type KmerVal = T;
    fn next(&mut self) -> Option<T> {
        if self.window.is_none() { self.window = Some(kmers_in_range::<T>(self.sequence, self.nb_base, self.range.start, self.range.end).into_iter()); }
        self.window.as_mut().unwrap().next()
    }
}

pub trait KmerGenerationPattern<T: KmerT> {
    fn generate_kmer_pattern(&self, seq: &SequenceAA) -> Vec<T>;
    fn generate_kmer_pattern_in_range(&self, seq: &SequenceAA, begin: usize, end: usize) -> Vec<T>;
    fn generate_kmer_distribution(&self, seq: &SequenceAA) -> FnvHashMap<T, usize>;
}

pub struct KmerGenerator<T: KmerT> {
    pub kmer_size: u8,
    t_marker: PhantomData<T>,
}
impl<T: KmerT> KmerGenerator<T> {
    pub fn new(ksize: u8) -> Self { KmerGenerator { kmer_size: ksize, t_marker: PhantomData } }
    pub fn generate_kmer(&self, seq: &SequenceAA) -> Vec<T> where Self: KmerGenerationPattern<T> { self.generate_kmer_pattern(seq) }
    pub fn generate_kmer_in_range(&self, seq: &SequenceAA, begin: usize, end: usize) -> Vec<T> where Self: KmerGenerationPattern<T> {
        self.generate_kmer_pattern_in_range(seq, begin, end)
    }
    pub fn generate_weighted_kmer(&self, seq: &SequenceAA) -> FnvHashMap<T, usize> where Self: KmerGenerationPattern<T> { self.generate_kmer_distribution(seq) }
    pub fn get_kmer_size(&self) -> usize { self.kmer_size as usize }
}
impl<T: CompressedKmerT + KmerBuilder<T> + std::hash::Hash + Eq> KmerGenerationPattern<T> for KmerGenerator<T> {
    fn generate_kmer_pattern(&self, seq: &SequenceAA) -> Vec<T> { kmers_in_range(seq, self.kmer_size as usize, 0, seq.size()) }
    fn generate_kmer_pattern_in_range(&self, seq: &SequenceAA, begin: usize, end: usize) -> Vec<T> { kmers_in_range(seq, self.kmer_size as usize, begin, end) }
    fn generate_kmer_distribution(&self, seq: &SequenceAA) -> FnvHashMap<T, usize> {
        let mut map = FnvHashMap::default();
        for kmer in self.generate_kmer_pattern(seq) { *map.entry(kmer).or_insert(0usize) += 1; } // a protein: hundreds of k-mers
        map
    }
}

pub fn hashmap_count_to_vec_count<T: CompressedKmerT + std::hash::Hash + Eq>(kmer_distribution: &FnvHashMap<T, usize>) -> Vec<(T, usize)> {
    kmer_distribution.iter().map(|(k, w)| (*k, *w)).collect()
}
