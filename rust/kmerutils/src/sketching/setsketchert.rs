//! Reference `src/sketching/setsketchert.rs`: the trait `SeqSketcherT` (:54-79) and its implementations on the path —
//! `ProbHash3aSketch` (:85-203), `SuperHashSketch` (:211-335), `HyperLogLogSketch` (:648-896) with `HllSeqsThreading`
//! (:602-635).  `sketch_compressedkmer` = one signature per sequence; `sketch_compressedkmer_seqs` = ONE signature for
//! the whole vector (a genome in several contigs), returned as a vector of length 1 like the reference does.
//! `SUPER2 / OPTDENS / REVOPTDENS` are not on the GPU path.
use crate::base::kmergenerator::{KmerGenerationPattern, KmerGenerator};
use crate::base::kmertraits::*;
use crate::base::sequence::{device_batch, Sequence};
use crate::devhash::DeviceKmerHash;
use crate::ffi;
use crate::sketcharg::{SeqSketcherParams, SketchAlgo};
use crate::sketching::seqsketchjaccard::{rows, SeqSketcher, SigFloat};
use serde::{Deserialize, Serialize};
use std::marker::PhantomData;
use std::os::raw::c_void;

pub trait SeqSketcherT<Kmer>
where
    Kmer: CompressedKmerT + KmerBuilder<Kmer>,
    KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
{
    type Sig: Clone + Send + Sync;
    fn get_kmer_size(&self) -> usize;
    fn get_sketch_size(&self) -> usize;
    fn get_algo(&self) -> SketchAlgo;
    fn sketch_compressedkmer<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&Sequence], fhash: H) -> Vec<Vec<Self::Sig>>;
    fn sketch_compressedkmer_seqs<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&Sequence], fhash: H) -> Vec<Vec<Self::Sig>>;
}

#[derive(Serialize, Deserialize, Copy, Clone)]
pub struct ProbHash3aSketch<Kmer> {
    _kmer_marker: PhantomData<Kmer>,
    params: SeqSketcherParams,
}
impl<Kmer> ProbHash3aSketch<Kmer> {
    pub fn new(params: &SeqSketcherParams) -> Self { ProbHash3aSketch { _kmer_marker: PhantomData, params: *params } }
}
impl<Kmer> ProbHash3aSketch<Kmer>
where
    Kmer: CompressedKmerT + KmerBuilder<Kmer>,
{
    /// many genomes in one call: `group_sizes[g]` consecutive sequences of `vseq` form genome g (gsearch's file loop
    /// without a host round trip between files)
    pub fn sketch_compressedkmer_groups<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&Sequence], group_sizes: &[u64], _fhash: H) -> Vec<Vec<Kmer::Val>> {
        let b = device_batch(vseq);
        let m = self.params.get_sketch_size();
        let mut flat = vec![Kmer::Val::default(); group_sizes.len() * m];
        ffi::check(unsafe { ffi::kmu_sketch_pmh3a_groups(ffi::ctx(), b.0, group_sizes.as_ptr(), group_sizes.len() as u64, self.params.get_kmer_size() as u32,
                                                         Kmer::KMU_TYPE, H::KIND, m as u32, flat.as_mut_ptr() as *mut c_void, 0) },
                   "ProbHash3aSketch::sketch_compressedkmer_seqs");
        rows(flat, m)
    }
}
impl<Kmer> SeqSketcherT<Kmer> for ProbHash3aSketch<Kmer>
where
    Kmer: CompressedKmerT + KmerBuilder<Kmer> + Send + Sync,
    Kmer::Val: Send + Sync,
    KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
{
    type Sig = Kmer::Val;
    fn get_kmer_size(&self) -> usize { self.params.get_kmer_size() }
    fn get_sketch_size(&self) -> usize { self.params.get_sketch_size() }
    fn get_algo(&self) -> SketchAlgo { SketchAlgo::PROB3A }
    fn sketch_compressedkmer<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&Sequence], fhash: H) -> Vec<Vec<Kmer::Val>> {
        SeqSketcher::new(self.get_kmer_size(), self.get_sketch_size()).sketch_probminhash3a::<Kmer, H>(vseq, fhash)
    }
    fn sketch_compressedkmer_seqs<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&Sequence], _fhash: H) -> Vec<Vec<Kmer::Val>> {
        let b = device_batch(vseq);
        let mut sig = vec![Kmer::Val::default(); self.get_sketch_size()];
        ffi::check(unsafe { ffi::kmu_sketch_pmh3a_whole(ffi::ctx(), b.0, self.get_kmer_size() as u32, Kmer::KMU_TYPE, H::KIND, self.get_sketch_size() as u32,
                                                        sig.as_mut_ptr() as *mut c_void, 0) }, "ProbHash3aSketch::sketch_compressedkmer_seqs");
        vec![sig]
    }
}

#[derive(Serialize, Deserialize, Copy, Clone)]
pub struct SuperHashSketch<Kmer, S> {
    _kmer_marker: PhantomData<Kmer>,
    _sig_marker: PhantomData<S>,
    params: SeqSketcherParams,
}
impl<Kmer, S> SuperHashSketch<Kmer, S> {
    pub fn new(params: &SeqSketcherParams) -> Self { SuperHashSketch { _kmer_marker: PhantomData, _sig_marker: PhantomData, params: *params } }
}
impl<Kmer, S> SeqSketcherT<Kmer> for SuperHashSketch<Kmer, S>
where
    Kmer: CompressedKmerT + KmerBuilder<Kmer> + Send + Sync,
    KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
    S: SigFloat,
{
    type Sig = S;
    fn get_kmer_size(&self) -> usize { self.params.get_kmer_size() }
    fn get_sketch_size(&self) -> usize { self.params.get_sketch_size() }
    fn get_algo(&self) -> SketchAlgo { SketchAlgo::SUPER }
    /// (:255-294) NoHashHasher on the k-mer values (:266-269) — `SeqSketcher::sketch_superminhash` uses FNV instead
    fn sketch_compressedkmer<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&Sequence], _fhash: H) -> Vec<Vec<S>> {
        let b = device_batch(vseq);
        let mut flat = vec![S::default(); vseq.len() * self.get_sketch_size()];
        ffi::check(unsafe { ffi::kmu_sketch_superminhash(ffi::ctx(), b.0, self.get_kmer_size() as u32, Kmer::KMU_TYPE, H::KIND, self.get_sketch_size() as u32,
                                                         ffi::KMU_HASHER_NOHASH, S::BYTES, flat.as_mut_ptr() as *mut c_void, 0) },
                   "SuperHashSketch::sketch_compressedkmer");
        rows(flat, self.get_sketch_size())
    }
    /// (:296-335) one signature for the whole vector, NoHashHasher (:308-311)
    fn sketch_compressedkmer_seqs<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&Sequence], _fhash: H) -> Vec<Vec<S>> {
        let b = device_batch(vseq);
        let mut sig = vec![S::default(); self.get_sketch_size()];
        ffi::check(unsafe { ffi::kmu_sketch_superminhash_whole(ffi::ctx(), b.0, self.get_kmer_size() as u32, Kmer::KMU_TYPE, H::KIND,
                                                               self.get_sketch_size() as u32, ffi::KMU_HASHER_NOHASH, S::BYTES, sig.as_mut_ptr() as *mut c_void, 0) },
                   "SuperHashSketch::sketch_compressedkmer_seqs");
        vec![sig]
    }
}

/// the reference bounds its rayon blocks with this; on the GPU the split over CTAs is the kernel's, the type is kept for
/// signature compatibility
#[derive(Serialize, Deserialize, Copy, Clone, Debug)]
pub struct HllSeqsThreading {
    nb_iter_thread: usize,
    thread_threshold: usize,
}
impl HllSeqsThreading {
    pub fn new(nb_iter_thread: usize, thread_threshold: usize) -> Self { HllSeqsThreading { nb_iter_thread, thread_threshold } }
    pub fn get_nb_iter_threads(&self) -> usize { self.nb_iter_thread }
    pub fn get_thread_threshold(&self) -> usize { self.thread_threshold }
}
impl Default for HllSeqsThreading {
    fn default() -> Self { HllSeqsThreading { nb_iter_thread: 4, thread_threshold: 10_000_000 } }
}

/// probminhash::setsketcher::SetSketchParams (b, m, a, q); `Default` = (1.001, 4096, 20., 2^16 - 2) as in the crate
#[derive(Serialize, Deserialize, Copy, Clone, Debug)]
pub struct SetSketchParams {
    b: f64,
    m: u64,
    a: f64,
    q: u64,
}
impl SetSketchParams {
    pub fn new(b: f64, m: u64, a: f64, q: u64) -> Self { SetSketchParams { b, m, a, q } }
    pub fn get_m(&self) -> u64 { self.m }
    pub(crate) fn as_ffi(&self) -> ffi::kmu_setsketch_params { ffi::kmu_setsketch_params { b: self.b, m: self.m, a: self.a, q: self.q } }
}
impl Default for SetSketchParams {
    fn default() -> Self { SetSketchParams { b: 1.001, m: 4096, a: 20., q: (1u64 << 16) - 2 } }
}

/// the register types of a SetSketch signature
pub trait SigReg: Copy + Default + Send + Sync { const BYTES: i32; }
impl SigReg for u16 { const BYTES: i32 = 2; }
impl SigReg for u32 { const BYTES: i32 = 4; }
impl SigReg for u64 { const BYTES: i32 = 8; }

#[derive(Serialize, Deserialize, Copy, Clone)]
pub struct HyperLogLogSketch<Kmer, S> {
    params: SeqSketcherParams,
    hll_params: SetSketchParams,
    hll_threads: HllSeqsThreading,
    _kmer_marker: PhantomData<Kmer>,
    _sig_marker: PhantomData<S>,
}
impl<Kmer, S> HyperLogLogSketch<Kmer, S> {
    pub fn new(seq_params: &SeqSketcherParams, hll_params: SetSketchParams, hll_threads: HllSeqsThreading) -> Self {
        HyperLogLogSketch { params: *seq_params, hll_params, hll_threads, _kmer_marker: PhantomData, _sig_marker: PhantomData }
    }
}
impl<Kmer, S> HyperLogLogSketch<Kmer, S>
where
    Kmer: CompressedKmerT + KmerBuilder<Kmer>,
    S: SigReg,
{
    fn run<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&Sequence], whole: bool) -> Vec<Vec<S>> {
        let b = device_batch(vseq);
        let m = self.hll_params.get_m() as usize;
        let mut flat = vec![S::default(); if whole { m } else { m * vseq.len() }];
        let p = self.hll_params.as_ffi();
        ffi::check(unsafe { ffi::kmu_sketch_setsketch(ffi::ctx(), b.0, self.params.get_kmer_size() as u32, Kmer::KMU_TYPE, H::KIND, &p, S::BYTES,
                                                      whole as i32, flat.as_mut_ptr() as *mut c_void, 0) }, "HyperLogLogSketch");
        rows(flat, m)
    }
    /// (:677-724) the registers of a block of sequences, to be merged by the caller with an element-wise max
    pub fn sketch_compressedkmer_seqs_block<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&Sequence], _fhash: H) -> Vec<S> {
        self.run::<H>(vseq, true).pop().unwrap()
    }
}
impl<Kmer, S> SeqSketcherT<Kmer> for HyperLogLogSketch<Kmer, S>
where
    Kmer: CompressedKmerT + KmerBuilder<Kmer> + Send + Sync,
    KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
    S: SigReg,
{
    type Sig = S;
    fn get_kmer_size(&self) -> usize { self.params.get_kmer_size() }
    fn get_sketch_size(&self) -> usize { self.params.get_sketch_size() }
    fn get_algo(&self) -> SketchAlgo { SketchAlgo::HLL }
    fn sketch_compressedkmer<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&Sequence], _fhash: H) -> Vec<Vec<S>> { self.run::<H>(vseq, false) }
    /// (:811-895) blocks of sequences sketched apart and merged by max in the reference; per-CTA partial registers merged
    /// by `atomicMax` here — the same registers, max being associative and commutative
    fn sketch_compressedkmer_seqs<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&Sequence], _fhash: H) -> Vec<Vec<S>> { self.run::<H>(vseq, true) }
}
