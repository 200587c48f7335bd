//! Reference `src/base/kmergenerator.rs`: `KmerSeqIterator` (:30-107), `KmerGenerationPattern` / `KmerGenerator` (:117-186),
//! `hashmap_count_to_vec_count` (:189-203).  All k-mers come from the GPU (kmu_generate_kmers); the iterator streams them
//! in windows.
use super::kmertraits::*;
use super::sequence::{device_batch, Sequence};
use crate::ffi;
use fnv::FnvHashMap;
use std::marker::PhantomData;
use std::os::raw::c_void;

pub trait KmerSeqIteratorT {
    type KmerVal;
    fn next(&mut self) -> Option<Self::KmerVal>;
}

const FETCH: usize = 1 << 20;

pub struct KmerSeqIterator<'a, T: CompressedKmerT + KmerBuilder<T>> {
    nb_base: u8,
    seq: &'a Sequence,
    batch: ffi::DeviceBatch,
    end: usize,
    pos: usize,
    window: Vec<T::Val>,
    wpos: usize,
}

impl<'a, T: CompressedKmerT + KmerBuilder<T>> KmerSeqIterator<'a, T> {
    pub fn new(ksize: u8, sequence: &'a Sequence) -> Self {
        if ksize as usize > T::get_nb_base_max() {
            panic!("\n KmerSeqIterator cannot support so many bases for given kmer type, kmer size  {}", ksize);
        }
        assert!(sequence.size() > 0, "IterSequence::new on an empty sequence"); // sequence.rs:531 underflows
        KmerSeqIterator { nb_base: ksize, seq: sequence, batch: device_batch(&[sequence]), end: sequence.size(), pos: 0, window: Vec::new(), wpos: 0 }
    }
    pub fn set_range(&mut self, begin: usize, end: usize) -> std::result::Result<(), ()> {
        if end <= begin || end > self.seq.size() { return Err(()); }
        self.pos = begin;
        self.end = end;
        self.window.clear();
        self.wpos = 0;
        Ok(())
    }
    fn refill(&mut self) -> bool {
        if self.pos + self.nb_base as usize > self.end { return false; }
        let (idx, b, e) = (0u64, self.pos as u64, (self.end.min(self.pos + FETCH + self.nb_base as usize - 1)) as u64);
        let mut part = std::ptr::null_mut();
        ffi::check(unsafe { ffi::kmu_seqbatch_slices(ffi::ctx(), self.batch.0, &idx, &b, &e, 1, &mut part) }, "KmerSeqIterator::next");
        let part = ffi::DeviceBatch(part);
        let n = unsafe { ffi::kmu_kmer_count(part.0, self.nb_base as u32) } as usize;
        self.window = vec![T::Val::default(); n];
        ffi::check(unsafe { ffi::kmu_generate_kmers(ffi::ctx(), part.0, self.nb_base as u32, T::KMU_TYPE, ffi::KMU_HASH_MASKED_VALUE,
                                                    self.window.as_mut_ptr() as *mut c_void, std::ptr::null_mut(), 0) }, "KmerSeqIterator::next");
        self.wpos = 0;
        self.pos += n;
        n > 0
    }
}

impl<'a, T: CompressedKmerT + KmerBuilder<T>> KmerSeqIteratorT for KmerSeqIterator<'a, T> {
    type KmerVal = T;
    fn next(&mut self) -> Option<T> {
        if self.wpos >= self.window.len() && !self.refill() { return None; }
        self.wpos += 1;
        Some(<T as KmerBuilder<T>>::build(self.window[self.wpos - 1], self.nb_base)) // MASKED_VALUE + build restores the word
    }
}

pub trait KmerGenerationPattern<T: KmerT> {
    fn generate_kmer_pattern(&self, seq: &Sequence) -> Vec<T>;
    fn generate_kmer_pattern_in_range(&self, seq: &Sequence, begin: usize, end: usize) -> Vec<T>;
    fn generate_kmer_distribution(&self, seq: &Sequence) -> FnvHashMap<T, u32>;
}

pub struct KmerGenerator<T: KmerT> {
    pub kmer_size: u8,
    t_marker: PhantomData<T>,
}

impl<T: KmerT> KmerGenerator<T> {
    pub fn new(ksize: u8) -> Self { KmerGenerator { kmer_size: ksize, t_marker: PhantomData } }
    pub fn generate_kmer(&self, seq: &Sequence) -> Vec<T> where Self: KmerGenerationPattern<T> { self.generate_kmer_pattern(seq) }
    pub fn generate_kmer_in_range(&self, seq: &Sequence, begin: usize, end: usize) -> Vec<T> where Self: KmerGenerationPattern<T> {
        self.generate_kmer_pattern_in_range(seq, begin, end)
    }
    pub fn generate_weighted_kmer(&self, seq: &Sequence) -> FnvHashMap<T, u32> where Self: KmerGenerationPattern<T> { self.generate_kmer_distribution(seq) }
    pub fn get_kmer_size(&self) -> usize { self.kmer_size as usize }
}

/// one implementation for the three DNA k-mer types
impl<T> KmerGenerationPattern<T> for KmerGenerator<T>
where
    T: CompressedKmerT + KmerBuilder<T> + std::hash::Hash + Eq,
{
    fn generate_kmer_pattern(&self, seq: &Sequence) -> Vec<T> { self.generate_kmer_pattern_in_range(seq, 0, seq.size()) }
    fn generate_kmer_pattern_in_range(&self, seq: &Sequence, begin: usize, end: usize) -> Vec<T> {
        let whole = device_batch(&[seq]);
        let (idx, b, e) = (0u64, begin as u64, end as u64);
        let mut part = std::ptr::null_mut();
        ffi::check(unsafe { ffi::kmu_seqbatch_slices(ffi::ctx(), whole.0, &idx, &b, &e, 1, &mut part) }, "KmerSeqIterator::set_range");
        let part = ffi::DeviceBatch(part);
        let n = unsafe { ffi::kmu_kmer_count(part.0, self.kmer_size as u32) } as usize;
        let mut vals = vec![T::Val::default(); n];
        ffi::check(unsafe { ffi::kmu_generate_kmers(ffi::ctx(), part.0, self.kmer_size as u32, T::KMU_TYPE, ffi::KMU_HASH_MASKED_VALUE,
                                                    vals.as_mut_ptr() as *mut c_void, std::ptr::null_mut(), 0) }, "KmerGenerator::generate_kmer");
        vals.into_iter().map(|v| <T as KmerBuilder<T>>::build(v, self.kmer_size)).collect()
    }
    fn generate_kmer_distribution(&self, seq: &Sequence) -> FnvHashMap<T, u32> {
        // counted in an exact table on the GPU (forward k-mers, 32-bit counts), read back once
        let b = device_batch(&[seq]);
        let nk = unsafe { ffi::kmu_kmer_count(b.0, self.kmer_size as u32) };
        let mut c = std::ptr::null_mut();
        ffi::check(unsafe { ffi::kmu_count_create(ffi::ctx(), self.kmer_size as u32, T::KMU_TYPE, 32, nk.max(16), &mut c) }, "generate_kmer_distribution");
        let (mut keys, mut counts, mut n) = (vec![T::Val::default(); nk as usize + 1], vec![0u32; nk as usize + 1], 0u64);
        let mut rc = unsafe { ffi::kmu_count_insert_seqs(ffi::ctx(), c, b.0, 0) };
        if rc == 0 { rc = unsafe { ffi::kmu_count_export(ffi::ctx(), c, 1, keys.as_mut_ptr() as *mut c_void, counts.as_mut_ptr(), nk + 1, &mut n) }; }
        unsafe { ffi::kmu_count_destroy(c) };
        ffi::check(rc, "generate_kmer_distribution");
        (0..n as usize).map(|i| (<T as KmerBuilder<T>>::build(keys[i], self.kmer_size), counts[i])).collect()
    }
}

pub fn hashmap_count_to_vec_count<T: CompressedKmerT + std::hash::Hash>(kmer_distribution: &FnvHashMap<T, u32>) -> Vec<(T, u32)> {
    kmer_distribution.iter().map(|(k, w)| (*k, *w)).collect()
}
