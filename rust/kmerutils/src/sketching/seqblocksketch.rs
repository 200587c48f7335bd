//! Reference `src/sketching/seqblocksketch.rs`: `BlockSketched` (:38-65), `BlockSketchedSeq` (:73-76), `BlockSeqSketcher`
//! (:79-226), `DistBlockSketched` (:417-440).  A sequence is cut into runs of `block_size` consecutive Kmer32bit k-mers; the
//! number of blocks comes from the BASES (ceil(size / block_size)), so the last blocks may hold no k-mer at all.  All
//! blocks of all sequences of a pack are cut on the device (`kmu_seqbatch_slices`) and sketched by one call.
use crate::base::sequence::{device_batch, Sequence};
use crate::base::Kmer32bit;
use crate::devhash::DeviceKmerHash;
use crate::ffi;
use crate::sketching::seqsketchjaccard::probminhash_get_jaccard_objects;
use serde::{Deserialize, Serialize};
use std::fs::{self, File};
use std::io::{self, Write};
use std::os::raw::c_void;

pub const MAGIC_BLOCKSIG_DUMP: u32 = 0xceabbadd;

#[derive(Clone, Serialize, Deserialize)]
pub struct BlockSketched {
    numseq: u32,
    numblock: u32,
    sketch: Vec<u32>,
}

impl BlockSketched {
    pub fn new(numseq: u32, numblock: u32, sketch_size: u32) -> BlockSketched {
        BlockSketched { numseq, numblock, sketch: Vec::with_capacity(sketch_size as usize) }
    }
    pub fn get_skech_slice(&self) -> &[u32] { &self.sketch }
    fn dump(&self, out: &mut dyn Write) {
        let mut bytes = Vec::with_capacity(8 + 4 * self.sketch.len());
        bytes.extend(self.numseq.to_le_bytes());
        bytes.extend(self.numblock.to_le_bytes());
        bytes.extend(self.sketch.iter().flat_map(|v| v.to_le_bytes()));
        out.write_all(&bytes).unwrap();
    }
}

pub struct BlockSketchedSeq {
    numseq: usize,
    /// one vector of length 1 per block, the shape hnsw_rs wants
    pub sketch: Vec<Vec<BlockSketched>>,
}

pub struct BlockSeqSketcher {
    sig_size: u8,
    block_size: usize,
    kmer_size: usize,
    sketch_size: usize,
}

impl BlockSeqSketcher {
    pub fn new(block_size: usize, kmer_size: usize, sketch_size: usize) -> BlockSeqSketcher {
        BlockSeqSketcher { sig_size: 4, block_size, kmer_size, sketch_size }
    }

    pub fn blocksketch_sequence<H: DeviceKmerHash<Kmer32bit>>(&self, numseq: usize, seq: &Sequence, fhash: &H) -> BlockSketchedSeq {
        self.blocksketch_sequences(&[(numseq as u32, seq)], fhash).pop().unwrap()
    }

    pub fn blocksketch_sequences<H: DeviceKmerHash<Kmer32bit>>(&self, pack_seq: &[(u32, &Sequence)], _fhash: &H) -> Vec<BlockSketchedSeq> {
        let (mut idx, mut begin, mut end, mut nblocks) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for (s, (_, seq)) in pack_seq.iter().enumerate() {
            assert!(seq.size() > 0);
            let nb = (seq.size() + self.block_size - 1) / self.block_size;
            nblocks.push(nb);
            for b in 0..nb {
                idx.push(s as u64);
                begin.push((b * self.block_size) as u64);
                end.push((b * self.block_size + self.block_size + self.kmer_size - 1) as u64); // block_size k-mers; clamped to the sequence
            }
        }
        let vseq: Vec<&Sequence> = pack_seq.iter().map(|p| p.1).collect();
        let whole = device_batch(&vseq);
        let mut blocks = std::ptr::null_mut();
        ffi::check(unsafe { ffi::kmu_seqbatch_slices(ffi::ctx(), whole.0, idx.as_ptr(), begin.as_ptr(), end.as_ptr(), idx.len() as u64, &mut blocks) },
                   "BlockSeqSketcher");
        let blocks = ffi::DeviceBatch(blocks);
        let mut flat = vec![0u32; idx.len() * self.sketch_size];
        ffi::check(unsafe { ffi::kmu_sketch_pmh3a(ffi::ctx(), blocks.0, self.kmer_size as u32, ffi::KMU_KMER32, H::KIND, self.sketch_size as u32,
                                                  flat.as_mut_ptr() as *mut c_void, 0) }, "BlockSeqSketcher");
        let mut row = flat.chunks(self.sketch_size);
        pack_seq.iter().zip(&nblocks).map(|((numseq, _), nb)| BlockSketchedSeq {
            numseq: *numseq as usize,
            sketch: (0..*nb).map(|b| vec![BlockSketched { numseq: *numseq, numblock: b as u32, sketch: row.next().unwrap().to_vec() }]).collect(),
        }).collect()
    }

    /// (:172-196) per sequence: numseq, number of blocks, then every block (numseq, numblock, slots)
    pub fn dump_blocks(&self, out: &mut dyn Write, seqblocks: &[BlockSketchedSeq]) {
        for seqblock in seqblocks {
            assert!(!seqblock.sketch.is_empty());
            out.write_all(&(seqblock.numseq as u32).to_le_bytes()).unwrap();
            out.write_all(&(seqblock.sketch.len() as u32).to_le_bytes()).unwrap();
            seqblock.sketch.iter().for_each(|b| b[0].dump(out));
        }
    }

    /// (:198-226) header: magic, sig_size as ONE byte, sketch_size, kmer_size, block_size as little-endian u32
    pub fn create_signature_dump(&self, dumpfname: &String) -> io::BufWriter<fs::File> {
        let file = File::create(dumpfname).unwrap_or_else(|_| { println!("cannot open {}", dumpfname); std::process::exit(1) });
        let mut sigbuf = io::BufWriter::with_capacity(1 << 26, file);
        sigbuf.write_all(&MAGIC_BLOCKSIG_DUMP.to_le_bytes()).unwrap();
        sigbuf.write_all(&self.sig_size.to_le_bytes()).unwrap();
        for w in [self.sketch_size as u32, self.kmer_size as u32, self.block_size as u32] { sigbuf.write_all(&w.to_le_bytes()).unwrap(); }
        sigbuf
    }
}

/// the distance hnsw_rs is given (:417-440): 1 inside a sequence, else the fraction of differing slots
pub struct DistBlockSketched {}

impl DistBlockSketched {
    pub fn eval(&self, va: &[BlockSketched], vb: &[BlockSketched]) -> f32 {
        assert!(va.len() == 1 && vb.len() == 1);
        if va[0].numseq == vb[0].numseq { return 1.; }
        assert_eq!(va[0].sketch.len(), vb[0].sketch.len());
        (1. - probminhash_get_jaccard_objects(&va[0].sketch, &vb[0].sketch)) as f32
    }
}
