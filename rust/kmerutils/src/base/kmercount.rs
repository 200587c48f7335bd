//! Reference `src/base/kmercount.rs`: `KmerCountT` (:48-59), `KmerCounter` (:70-123, 241-288), `DispatchableT` (:382-420),
//! `KmerCounterPool` (:424-565), `count_kmer_threaded_one_to_many` (:881-974), `count_kmer_thread_independant` (:797-867).
//! The cuckoo + counting-Bloom pair is ONE exact table in HBM per counter: the reference's answers with zero filter false
//! positives (counts saturate at 2^nb_bits - 1).
use super::kmertraits::*;
use super::sequence::{device_batch, Sequence};
use super::{Kmer16b32bit, Kmer32bit, Kmer64bit};
use crate::devhash::{int32_hash, int64_hash};
use crate::ffi;
use std::marker::PhantomData;
use std::os::raw::c_void;

pub const COUNTER_UNIQUE: u32 = 0xcea2bbdd;
pub const COUNTER_MULTIPLE: u32 = 0xcea2bbff;

pub trait KmerCountT {
    type Kmer;
    fn insert_kmer(&mut self, kmer: Self::Kmer);
    fn get_count(&self, kmer: Self::Kmer) -> u32;
    fn get_nb_distinct(&self) -> u64;
    fn get_nb_unique(&self) -> u64;
}

pub struct KmerCounter<Kmer: CompressedKmerT> {
    handle: *mut ffi::kmu_counter,
    bloom_f_nb_bits: u8,
    pending: std::cell::RefCell<Vec<Kmer::Val>>,
    _kmertype: PhantomData<Kmer>,
}
unsafe impl<Kmer: CompressedKmerT> Send for KmerCounter<Kmer> {}

impl<Kmer: CompressedKmerT> KmerCounter<Kmer> {
    /// `KmerCounter::new(fpr, capacity, nb_bits)` (:88-98) + the k-mer size the table is typed with; fpr has no meaning here
    pub fn new(_fpr: f32, capacity: usize, nb_bits: usize, kmer_size: u8) -> KmerCounter<Kmer> {
        let mut h = std::ptr::null_mut();
        ffi::check(unsafe { ffi::kmu_count_create(ffi::ctx(), kmer_size as u32, Kmer::KMU_TYPE, nb_bits as u32, capacity as u64, &mut h) }, "KmerCounter::new");
        KmerCounter { handle: h, bloom_f_nb_bits: nb_bits as u8, pending: Default::default(), _kmertype: PhantomData }
    }
    fn flush(&self) {
        let mut p = self.pending.borrow_mut();
        if p.is_empty() { return; }
        ffi::check(unsafe { ffi::kmu_count_insert_kmers(ffi::ctx(), self.handle, p.as_ptr() as *const c_void, p.len() as u64, 0) }, "KmerCounter::insert_kmer");
        p.clear();
    }
    /// every k-mer of every sequence, canonical as count_kmer does (:313)
    pub fn insert_sequences(&mut self, vseq: &[&Sequence], canonical: bool) {
        self.flush();
        let b = device_batch(vseq);
        ffi::check(unsafe { ffi::kmu_count_insert_seqs(ffi::ctx(), self.handle, b.0, canonical as i32) }, "count_kmer");
    }
    pub fn get_above2_count(&self, kmer: Kmer) -> u32 { let c = self.get_count(kmer); if c >= 2 { c } else { 0 } }
    pub fn get_count_nb_bits(&self) -> u8 { self.bloom_f_nb_bits }
    pub(crate) fn raw(&self) -> *mut ffi::kmu_counter { self.flush(); self.handle }
}
impl<Kmer: CompressedKmerT> Drop for KmerCounter<Kmer> { fn drop(&mut self) { unsafe { ffi::kmu_count_destroy(self.handle) } } }

impl<Kmer: CompressedKmerT> KmerCountT for KmerCounter<Kmer> {
    type Kmer = Kmer;
    fn insert_kmer(&mut self, kmer: Kmer) {
        self.pending.borrow_mut().push(kmer.get_compressed_value());
        if self.pending.borrow().len() >= 1 << 20 { self.flush(); }
    }
    fn get_count(&self, kmer: Kmer) -> u32 {
        self.flush();
        let (key, mut cnt) = (kmer.get_compressed_value(), 0u32);
        ffi::check(unsafe { ffi::kmu_count_query(ffi::ctx(), self.handle, &key as *const Kmer::Val as *const c_void, 1, &mut cnt, 0) }, "get_count");
        cnt
    }
    fn get_nb_distinct(&self) -> u64 { self.stats().0 }
    fn get_nb_unique(&self) -> u64 { self.stats().1 }
}
impl<Kmer: CompressedKmerT> KmerCounter<Kmer> {
    fn stats(&self) -> (u64, u64) {
        self.flush();
        let (mut d, mut u) = (0u64, 0u64);
        ffi::check(unsafe { ffi::kmu_count_stats(ffi::ctx(), self.handle, &mut d, &mut u, std::ptr::null_mut(), std::ptr::null_mut()) }, "KmerCounter stats");
        (d, u)
    }
}

pub trait DispatchableT {
    type ToDispatch;
    fn dispatch(&self, nb_receiver: usize) -> usize;
}
impl DispatchableT for Kmer16b32bit { type ToDispatch = u32; fn dispatch(&self, n: usize) -> usize { (int32_hash(self.get_compressed_value()) % n as u32) as usize } }
impl DispatchableT for Kmer32bit { type ToDispatch = u32; fn dispatch(&self, n: usize) -> usize { (int32_hash(self.get_compressed_value()) % n as u32) as usize } }
impl DispatchableT for Kmer64bit { type ToDispatch = u64; fn dispatch(&self, n: usize) -> usize { (int64_hash(self.get_compressed_value()) % n as u64) as usize } }

pub struct KmerCounterPool<Kmer: CompressedKmerT> {
    pub counters: Vec<Box<KmerCounter<Kmer>>>,
}
impl<Kmer: CompressedKmerT + DispatchableT> KmerCounterPool<Kmer> {
    pub fn new(counters: Vec<Box<KmerCounter<Kmer>>>) -> Self { KmerCounterPool { counters } }
    pub fn get_above2_count(&self, kmer: Kmer) -> u32 { self.counters[kmer.dispatch(self.counters.len())].get_above2_count(kmer) }
    pub fn get_count_nb_bits(&self) -> u8 { self.counters.first().map_or(0, |c| c.get_count_nb_bits()) }
}
impl<Kmer: CompressedKmerT + DispatchableT> KmerCountT for KmerCounterPool<Kmer> {
    type Kmer = Kmer;
    fn insert_kmer(&mut self, kmer: Kmer) { let loc = kmer.dispatch(self.counters.len()); self.counters[loc].insert_kmer(kmer); }
    fn get_count(&self, kmer: Kmer) -> u32 { self.counters[kmer.dispatch(self.counters.len())].get_count(kmer) }
    fn get_nb_distinct(&self) -> u64 { self.counters.iter().map(|v| v.get_nb_distinct()).sum() }
    fn get_nb_unique(&self) -> u64 { self.counters.iter().map(|v| v.get_nb_unique()).sum() }
}

fn count_into_pool<Kmer: CompressedKmerT + DispatchableT>(seqvec: &Vec<Sequence>, nb_threads: usize, nb_bits: usize, kmer_size: usize) -> KmerCounterPool<Kmer> {
    // the canonical k-mers of all sequences bucketed by DispatchableT::dispatch on the GPU, bucket i into counter i
    let refs: Vec<&Sequence> = seqvec.iter().collect();
    let b = device_batch(&refs);
    let nk = unsafe { ffi::kmu_kmer_count(b.0, kmer_size as u32) } as usize;
    let counters: Vec<Box<KmerCounter<Kmer>>> =
        (0..nb_threads).map(|_| Box::new(KmerCounter::new(0.03, nk / nb_threads * 13 / 10 + 1024, nb_bits, kmer_size as u8))).collect();
    if nk > 0 {
        let mut keys = vec![Kmer::Val::default(); nk];
        let mut part = vec![0u64; nb_threads];
        ffi::check(unsafe { ffi::kmu_count_partition(ffi::ctx(), b.0, kmer_size as u32, Kmer::KMU_TYPE, 1, nb_threads as u32,
                                                     keys.as_mut_ptr() as *mut c_void, part.as_mut_ptr(), 0) }, "count_kmer");
        let mut off = 0usize;
        for (i, c) in counters.iter().enumerate() {
            ffi::check(unsafe { ffi::kmu_count_insert_kmers(ffi::ctx(), c.raw(), keys[off..].as_ptr() as *const c_void, part[i], 0) }, "count_kmer");
            off += part[i] as usize;
        }
    }
    KmerCounterPool::new(counters)
}

/// `count_size` is the number of BITS of a count (:893)
pub fn count_kmer_threaded_one_to_many<Kmer: CompressedKmerT + DispatchableT + Send>(seqvec: &Vec<Sequence>, nb_threads: usize, count_size: usize,
                                                                                     kmer_size: usize) -> KmerCounterPool<Kmer> {
    count_into_pool(seqvec, nb_threads, count_size, kmer_size)
}
pub fn count_kmer_thread_independant<Kmer: CompressedKmerT + DispatchableT + Send>(seqvec: &Vec<Sequence>, nb_threads: usize, kmer_size: usize) -> KmerCounterPool<Kmer> {
    count_into_pool(seqvec, nb_threads, 8, kmer_size)
}
