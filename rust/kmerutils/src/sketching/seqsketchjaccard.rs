//! Reference `src/sketching/seqsketchjaccard.rs`: `compute_probminhash3a_jaccard` / `probminhash_get_jaccard_objects`
//! (:58-108), `SeqSketcher` (:117-414), `jaccard_index_probminhash3a` (:423-495), `dump_signatures_block_u32` (:577-585),
//! `SigSketchFileReader` (:588-712).  One call sketches the whole `&[&Sequence]`: the rayon loop over sequences of the
//! reference is the grid of the CUDA kernel.
use crate::base::kmergenerator::{KmerGenerationPattern, KmerGenerator};
use crate::base::kmertraits::*;
use crate::base::sequence::{device_batch, Sequence};
use crate::devhash::DeviceKmerHash;
use crate::ffi;
use serde::{Deserialize, Serialize};
use std::ffi::CString;
use std::fs::{self, File};
use std::io::{self, BufReader, BufWriter, Read, Write};
use std::os::raw::c_void;
use std::path::Path;

pub const MAGIC_SIG_DUMP: u32 = 0xceabeadd;

/// fraction of equal slots of two signatures (:86-108)
pub fn probminhash_get_jaccard_objects<D: Eq + Copy>(siga: &[D], sigb: &[D]) -> f64 {
    assert_eq!(siga.len(), sigb.len());
    siga.iter().zip(sigb).filter(|(a, b)| a == b).count() as f64 / siga.len() as f64
}

/// rows of `flat` (nseq x m) as the reference's `Vec<Vec<_>>`
pub(crate) fn rows<V: Copy>(flat: Vec<V>, m: usize) -> Vec<Vec<V>> { flat.chunks(m).map(|r| r.to_vec()).collect() }

/// the floating types a SuperMinHash signature comes in
pub trait SigFloat: Copy + Default + Send + Sync { const BYTES: i32; }
impl SigFloat for f32 { const BYTES: i32 = 4; }
impl SigFloat for f64 { const BYTES: i32 = 8; }

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

    /// (:211-260) one ProbMinHash3a signature of `sketch_size` k-mer values per sequence, k-mer multiplicities as weights.
    /// Host sequences go through the chunked three-stream pipeline (upload of chunk c+1 behind the kernels of chunk c).
    pub fn sketch_probminhash3a<Kmer, H>(&self, vseq: &[&Sequence], _fhash: H) -> Vec<Vec<Kmer::Val>>
    where
        Kmer: CompressedKmerT + KmerBuilder<Kmer>,
        H: DeviceKmerHash<Kmer>,
        KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
    {
        let ptrs: Vec<*const u8> = vseq.iter().map(|s| s.packed().as_ptr()).collect();
        let nb: Vec<u64> = vseq.iter().map(|s| s.size() as u64).collect();
        let mut flat = vec![Kmer::Val::default(); vseq.len() * self.sketch_size];
        ffi::check(unsafe { ffi::kmu_sketch_pmh3a_host_ptrs(ffi::ctx(), ptrs.as_ptr(), nb.as_ptr(), vseq.len() as u64, self.kmer_size as u32,
                                                            Kmer::KMU_TYPE, H::KIND, self.sketch_size as u32, flat.as_mut_ptr() as *mut c_void) },
                   "sketch_probminhash3a");
        rows(flat, self.sketch_size)
    }

    /// (:328-380) SuperMinHash, `S` = f32 or f64; the k-mer values go through `fnv::FnvHasher` as in the reference (:346-349)
    pub fn sketch_superminhash<Kmer, S, H>(&self, vseq: &[&Sequence], _fhash: H) -> Vec<Vec<S>>
    where
        Kmer: CompressedKmerT + KmerBuilder<Kmer>,
        S: SigFloat,
        H: DeviceKmerHash<Kmer>,
        KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
    {
        let b = device_batch(vseq);
        let mut flat = vec![S::default(); vseq.len() * self.sketch_size];
        ffi::check(unsafe { ffi::kmu_sketch_superminhash(ffi::ctx(), b.0, self.kmer_size as u32, Kmer::KMU_TYPE, H::KIND, self.sketch_size as u32,
                                                         ffi::KMU_HASHER_FNV, S::BYTES, flat.as_mut_ptr() as *mut c_void, 0) },
                   "sketch_superminhash");
        rows(flat, self.sketch_size)
    }

    /// (:390-414) header of a signature dump: magic, sig_size = 4, sketch_size, kmer_size as little-endian u32
    pub fn create_signature_dump(&self, dumpfname: &String) -> io::BufWriter<fs::File> {
        let file = File::create(dumpfname).unwrap_or_else(|_| { println!("cannot open {}", dumpfname); std::process::exit(1) });
        let mut sigbuf = io::BufWriter::with_capacity(1 << 26, file);
        for w in [MAGIC_SIG_DUMP, 4u32, self.sketch_size as u32, self.kmer_size as u32] { sigbuf.write_all(&w.to_le_bytes()).unwrap(); }
        sigbuf
    }
}

/// (:423-495) J_p(seqa, seqb) for every seqb: all sequences sketched in one call, the slot comparison on the GPU
pub fn jaccard_index_probminhash3a<Kmer, H>(seqa: &Sequence, vseqb: &[Sequence], sketch_size: usize, kmer_size: u8, fhash: H) -> Vec<f64>
where
    Kmer: CompressedKmerT + KmerBuilder<Kmer>,
    H: DeviceKmerHash<Kmer>,
    KmerGenerator<Kmer>: KmerGenerationPattern<Kmer>,
{
    let mut all: Vec<&Sequence> = vec![seqa];
    all.extend(vseqb.iter());
    let sigs = SeqSketcher::new(kmer_size as usize, sketch_size).sketch_probminhash3a::<Kmer, H>(&all, fhash);
    let flat_b: Vec<Kmer::Val> = sigs[1..].iter().flatten().copied().collect();
    let mut out = vec![0f64; vseqb.len()];
    ffi::check(unsafe { ffi::kmu_signature_jaccard(ffi::ctx(), sigs[0].as_ptr() as *const c_void, 1, flat_b.as_ptr() as *const c_void, vseqb.len() as u64,
                                                   sketch_size as u32, std::mem::size_of::<Kmer::Val>() as i32, out.as_mut_ptr(), 0) },
               "jaccard_index_probminhash3a");
    out
}

pub fn dump_signatures_block_u32(signatures: &[Vec<u32>], out: &mut dyn Write) -> io::Result<()> {
    for sig in signatures {
        let bytes: Vec<u8> = sig.iter().flat_map(|v| v.to_le_bytes()).collect();
        out.write_all(&bytes)?;
    }
    Ok(())
}

/// streams signatures back from a dump (:588-712)
pub struct SigSketchFileReader {
    _fname: String,
    sig_size: u8,
    sketch_size: usize,
    kmer_size: u8,
    signature_buf: io::BufReader<fs::File>,
}

impl SigSketchFileReader {
    pub fn new(fname: &String) -> Result<SigSketchFileReader, String> {
        let _ = CString::new(fname.as_str()).map_err(|_| "bad file name".to_string())?;
        let mut rd = BufReader::new(File::open(fname).map_err(|_| format!("SigSketchFileReader could not open file {}", fname))?);
        let mut word = || -> Result<u32, String> {
            let mut b = [0u8; 4];
            rd.read_exact(&mut b).map_err(|_| "SigSketchFileReader could no read header".to_string())?;
            Ok(u32::from_le_bytes(b))
        };
        if word()? != MAGIC_SIG_DUMP { return Err("file is not a dump of signature".to_string()); }
        let (sig_size, sketch_size, kmer_size) = (word()?, word()?, word()?);
        Ok(SigSketchFileReader { _fname: fname.clone(), sig_size: sig_size as u8, sketch_size: sketch_size as usize, kmer_size: kmer_size as u8, signature_buf: rd })
    }
    pub fn get_kmer_size(&self) -> u8 { self.kmer_size }
    pub fn get_signature_length(&self) -> usize { self.sketch_size }
    pub fn get_signature_size(&self) -> usize { self.sig_size as usize }
    pub fn next(&mut self) -> Option<Vec<u32>> {
        let mut bytes = vec![0u8; 4 * self.sketch_size];
        self.signature_buf.read_exact(&mut bytes).ok()?;
        Some(bytes.chunks_exact(4).map(|c| u32::from_le_bytes(c.try_into().unwrap())).collect())
    }
}
