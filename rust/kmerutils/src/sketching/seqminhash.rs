//! Reference `src/sketching/seqminhash.rs:19-62` (`sketch_seqrange_superminhash`): SuperMinHash (f64) of the canonical,
//! `int32_hash`ed k-mers of a range of one sequence; Kmer16b32bit for k = 16, Kmer32bit for 9 <= k <= 15, a panic otherwise.
use crate::base::sequence::{device_batch, Sequence};
use crate::ffi;
use std::ops::Range;
use std::os::raw::c_void;

pub fn sketch_seqrange_superminhash(seq: &Sequence, range: &Range<usize>, kmer_size: usize, sketch_size: usize) -> Vec<f64> {
    let kmer_type = match kmer_size {
        16 => ffi::KMU_KMER16B32,
        9..=15 => ffi::KMU_KMER32,
        _ => panic!("sketch_sequence_superminhash , unimplemented kmer_size {} {} {} ", kmer_size, file!(), line!()),
    };
    let whole = device_batch(&[seq]);
    let (idx, b, e) = (0u64, range.start as u64, range.end as u64);
    let mut part = std::ptr::null_mut();
    ffi::check(unsafe { ffi::kmu_seqbatch_slices(ffi::ctx(), whole.0, &idx, &b, &e, 1, &mut part) }, "KmerSeqIterator::set_range");
    let part = ffi::DeviceBatch(part);
    let mut sig = vec![0f64; sketch_size];
    ffi::check(unsafe { ffi::kmu_sketch_superminhash(ffi::ctx(), part.0, kmer_size as u32, kmer_type, ffi::KMU_HASH_CANON_INVHASH, sketch_size as u32,
                                                     ffi::KMU_HASHER_NOHASH, 8, sig.as_mut_ptr() as *mut c_void, 0) },
               "sketch_seqrange_superminhash");
    sig
}
