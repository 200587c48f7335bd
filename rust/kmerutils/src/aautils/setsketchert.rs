//! Reference `src/aautils/setsketchert.rs`: trait `SeqSketcherAAT` (:42-72), `ProbHash3aSketch` (:78-196),
//! `SuperHashSketch` (:203-329), `HyperLogLogSketch` (:790-1012), `SeqSketcher` (:1020-1200) — the amino-acid twins of
//! `sketching::setsketchert`, over `SequenceAA` and the 5-bit k-mers.
use super::kmeraa::*;
use crate::base::kmertraits::*;
use crate::devhash::DeviceKmerHash;
use crate::ffi;
use crate::sketcharg::{SeqSketcherParams, SketchAlgo};
use crate::sketching::seqsketchjaccard::{rows, SigFloat};
pub use crate::sketching::setsketchert::{HllSeqsThreading, SetSketchParams, SigReg};
use serde::{Deserialize, Serialize};
use std::fs::File;
use std::io::{BufReader, BufWriter};
use std::marker::PhantomData;
use std::os::raw::c_void;
use std::path::Path;

pub trait SeqSketcherAAT<Kmer>
where
    Kmer: CompressedKmerT + KmerBuilder<Kmer>,
    KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
{
    type Sig: Clone + Send + Sync;
    fn get_kmer_size(&self) -> usize;
    fn get_sketch_size(&self) -> usize;
    fn get_algo(&self) -> SketchAlgo;
    fn sketch_compressedkmeraa<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&SequenceAA], fhash: H) -> Vec<Vec<Self::Sig>>;
    fn sketch_compressedkmeraa_seqs<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&SequenceAA], fhash: H) -> Vec<Vec<Self::Sig>>;
}

fn pmh3a<Kmer: CompressedKmerT, H: DeviceKmerHash<Kmer>>(vseq: &[&SequenceAA], k: usize, m: usize, whole: bool, what: &str) -> Vec<Vec<Kmer::Val>> {
    let b = device_batch_aa(vseq);
    let mut flat = vec![Kmer::Val::default(); if whole { m } else { m * vseq.len() }];
    let rc = unsafe {
        if whole { ffi::kmu_sketch_pmh3a_whole(ffi::ctx(), b.0, k as u32, Kmer::KMU_TYPE, H::KIND, m as u32, flat.as_mut_ptr() as *mut c_void, 0) }
        else { ffi::kmu_sketch_pmh3a(ffi::ctx(), b.0, k as u32, Kmer::KMU_TYPE, H::KIND, m as u32, flat.as_mut_ptr() as *mut c_void, 0) }
    };
    ffi::check(rc, what);
    rows(flat, m)
}

fn smh<Kmer: CompressedKmerT, S: SigFloat, H: DeviceKmerHash<Kmer>>(vseq: &[&SequenceAA], k: usize, m: usize, whole: bool, hasher: i32, what: &str) -> Vec<Vec<S>> {
    let b = device_batch_aa(vseq);
    let mut flat = vec![S::default(); if whole { m } else { m * vseq.len() }];
    let rc = unsafe {
        if whole { ffi::kmu_sketch_superminhash_whole(ffi::ctx(), b.0, k as u32, Kmer::KMU_TYPE, H::KIND, m as u32, hasher, S::BYTES, flat.as_mut_ptr() as *mut c_void, 0) }
        else { ffi::kmu_sketch_superminhash(ffi::ctx(), b.0, k as u32, Kmer::KMU_TYPE, H::KIND, m as u32, hasher, S::BYTES, flat.as_mut_ptr() as *mut c_void, 0) }
    };
    ffi::check(rc, what);
    rows(flat, m)
}

#[derive(Serialize, Deserialize, Copy, Clone)]
pub struct ProbHash3aSketch<Kmer> {
    _kmer_marker: PhantomData<Kmer>,
    params: SeqSketcherParams,
}
impl<Kmer> ProbHash3aSketch<Kmer> {
    pub fn new(params: &SeqSketcherParams) -> Self { ProbHash3aSketch { _kmer_marker: PhantomData, params: *params } }
}
impl<Kmer> SeqSketcherAAT<Kmer> for ProbHash3aSketch<Kmer>
where
    Kmer: CompressedKmerT + KmerBuilder<Kmer> + Send + Sync,
    Kmer::Val: Send + Sync,
    KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
{
    type Sig = Kmer::Val;
    fn get_kmer_size(&self) -> usize { self.params.get_kmer_size() }
    fn get_sketch_size(&self) -> usize { self.params.get_sketch_size() }
    fn get_algo(&self) -> SketchAlgo { SketchAlgo::PROB3A }
    fn sketch_compressedkmeraa<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&SequenceAA], _fhash: H) -> Vec<Vec<Kmer::Val>> {
        pmh3a::<Kmer, H>(vseq, self.get_kmer_size(), self.get_sketch_size(), false, "aautils::ProbHash3aSketch::sketch_compressedkmeraa")
    }
    fn sketch_compressedkmeraa_seqs<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&SequenceAA], _fhash: H) -> Vec<Vec<Kmer::Val>> {
        pmh3a::<Kmer, H>(vseq, self.get_kmer_size(), self.get_sketch_size(), true, "aautils::ProbHash3aSketch::sketch_compressedkmeraa_seqs")
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
impl<Kmer, S> SeqSketcherAAT<Kmer> for SuperHashSketch<Kmer, S>
where
    Kmer: CompressedKmerT + KmerBuilder<Kmer> + Send + Sync,
    KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
    S: SigFloat,
{
    type Sig = S;
    fn get_kmer_size(&self) -> usize { self.params.get_kmer_size() }
    fn get_sketch_size(&self) -> usize { self.params.get_sketch_size() }
    fn get_algo(&self) -> SketchAlgo { SketchAlgo::SUPER }
    /// NoHashHasher on the k-mer values (:258-261, :300-303)
    fn sketch_compressedkmeraa<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&SequenceAA], _fhash: H) -> Vec<Vec<S>> {
        smh::<Kmer, S, H>(vseq, self.get_kmer_size(), self.get_sketch_size(), false, ffi::KMU_HASHER_NOHASH, "aautils::SuperHashSketch::sketch_compressedkmeraa")
    }
    fn sketch_compressedkmeraa_seqs<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&SequenceAA], _fhash: H) -> Vec<Vec<S>> {
        smh::<Kmer, S, H>(vseq, self.get_kmer_size(), self.get_sketch_size(), true, ffi::KMU_HASHER_NOHASH, "aautils::SuperHashSketch::sketch_compressedkmeraa_seqs")
    }
}

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
impl<Kmer: CompressedKmerT + KmerBuilder<Kmer>, S: SigReg> HyperLogLogSketch<Kmer, S> {
    fn run<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&SequenceAA], whole: bool) -> Vec<Vec<S>> {
        let b = device_batch_aa(vseq);
        let m = self.hll_params.get_m() as usize;
        let mut flat = vec![S::default(); if whole { m } else { m * vseq.len() }];
        let p = self.hll_params.as_ffi();
        ffi::check(unsafe { ffi::kmu_sketch_setsketch(ffi::ctx(), b.0, self.params.get_kmer_size() as u32, Kmer::KMU_TYPE, H::KIND, &p, S::BYTES,
                                                      whole as i32, flat.as_mut_ptr() as *mut c_void, 0) }, "aautils::HyperLogLogSketch");
        rows(flat, m)
    }
    pub fn sketch_compressedkmer_seqs_block<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&SequenceAA], _fhash: H) -> Vec<S> { self.run::<H>(vseq, true).pop().unwrap() }
}
impl<Kmer, S> SeqSketcherAAT<Kmer> for HyperLogLogSketch<Kmer, S>
where
    Kmer: CompressedKmerT + KmerBuilder<Kmer> + Send + Sync,
    KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
    S: SigReg,
{
    type Sig = S;
    fn get_kmer_size(&self) -> usize { self.params.get_kmer_size() }
    fn get_sketch_size(&self) -> usize { self.params.get_sketch_size() }
    fn get_algo(&self) -> SketchAlgo { SketchAlgo::HLL }
    fn sketch_compressedkmeraa<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&SequenceAA], _fhash: H) -> Vec<Vec<S>> { self.run::<H>(vseq, false) }
    fn sketch_compressedkmeraa_seqs<H: DeviceKmerHash<Kmer>>(&self, vseq: &[&SequenceAA], _fhash: H) -> Vec<Vec<S>> { self.run::<H>(vseq, true) }
}

/// (:1020-1200) the non-trait entry points on proteins
#[derive(Copy, Clone, Serialize, Deserialize)]
pub struct SeqSketcher {
    kmer_size: usize,
    sketch_size: usize,
}
impl SeqSketcher {
    pub fn new(kmer_size: usize, sketch_size: usize) -> Self { SeqSketcher { kmer_size, sketch_size } }
    pub fn get_kmer_size(&self) -> usize { self.kmer_size }
    pub fn get_sketch_size(&self) -> usize { self.sketch_size }
    pub fn dump_json(&self, filename: &String) -> Result<(), String> {
        let file = File::create(filename).map_err(|_| "SeqSketcher dump failed".to_string())?;
        serde_json::to_writer(BufWriter::new(file), self).map_err(|e| e.to_string())
    }
    pub fn reload_json(dirpath: &Path) -> Result<SeqSketcher, String> {
        let file = File::open(dirpath.join("sketchparams_dump.json")).map_err(|_| "SeqSketcher reload_json could not open file".to_string())?;
        serde_json::from_reader(BufReader::new(file)).map_err(|e| e.to_string())
    }
    pub fn sketch_probminhash3a<Kmer, H>(&self, vseq: &[&SequenceAA], _fhash: H) -> Vec<Vec<Kmer::Val>>
    where
        Kmer: CompressedKmerT + KmerBuilder<Kmer>,
        H: DeviceKmerHash<Kmer>,
        KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
    {
        pmh3a::<Kmer, H>(vseq, self.kmer_size, self.sketch_size, false, "aautils::SeqSketcher::sketch_probminhash3a")
    }
    /// f64 signatures, FNV-hashed keys (:1170-1173)
    pub fn sketch_superminhash<Kmer, H>(&self, vseq: &[&SequenceAA], _fhash: H) -> Vec<Vec<f64>>
    where
        Kmer: CompressedKmerT + KmerBuilder<Kmer>,
        H: DeviceKmerHash<Kmer>,
        KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
    {
        smh::<Kmer, f64, H>(vseq, self.kmer_size, self.sketch_size, false, ffi::KMU_HASHER_FNV, "aautils::SeqSketcher::sketch_superminhash")
    }
}
