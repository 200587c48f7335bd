//! Reference `src/base/nthash.rs` (seeds :17-20, multi-hash :63-72, trait :76-120) and the impl of `src/base/kmer.rs:45-145`
//! for the two u32 k-mer types.  The per-k-mer `*_init` methods are host-side single-word arithmetic (like `push`); the
//! data path is `nthash_canonical_batch`, one kernel over all k-mers of a batch.  The `*_cycle` methods keep the
//! reference's behaviour bug for bug.
use super::kmertraits::*;
use super::sequence::{device_batch, Sequence};
use super::{Kmer16b32bit, Kmer32bit};
use crate::ffi;

pub const MULTISHIFT: usize = 27;
pub const MULTISEED: u64 = 0x90b45d39fb6da1fa;
pub const BASE_MAPPING_2B: [u64; 8] = [0x3c8bfbb395c60474, 0x3193c18562a02b4c, 0x20323ed082572324, 0x295549f54be24456,
                                       0x295549f54be24456, 0x20323ed082572324, 0x3193c18562a02b4c, 0x3c8bfbb395c60474];

pub fn from_one_hash_val_to_mult_hash(ksize: u64, hashed: &mut [u64]) {
    for i in 1..hashed.len() {
        let mut t = hashed[0].wrapping_mul(i as u64 ^ ksize.wrapping_mul(MULTISEED));
        t ^= t >> MULTISHIFT;
        hashed[i] = t;
    }
}

pub trait NtHash {
    fn nthash_init(&self) -> u64;
    fn nthash_cycle(&mut self, hashval: u64, new_base: u8) -> u64;
    fn nthash_canonical_init(&self, fhash: &mut u64, rhash: &mut u64) -> (u64, u8);
    fn nthash_canonical_cycle(&mut self, new_base: u8, fhash: &mut u64, rhash: &mut u64) -> (u64, u8);
    fn nthash_mult_canonical_init(&self, fhash: &mut u64, rhash: &mut u64, hashed: &mut [u64]) -> u8;
    fn nthash_mult_canonical_cycle(&mut self, new_base: u8, fhash: &mut u64, rhash: &mut u64, hashed: &mut [u64]) -> u8;
}

macro_rules! implement_nthash_for {
    ($ty:ty) => {
        impl NtHash for $ty {
            fn nthash_init(&self) -> u64 {
                let (mut f, mut r) = (0u64, 0u64);
                self.nthash_canonical_init(&mut f, &mut r);
                f
            }
            fn nthash_cycle(&mut self, hashval: u64, new_base: u8) -> u64 {
                let k = self.get_nb_base() as u32;
                let old = ((self.0 >> (2 * (k - 1))) & 3) as usize;
                let _ = self.push(new_base); // dropped, as in the reference (kmer.rs:69)
                hashval.rotate_left(1) ^ BASE_MAPPING_2B[old].rotate_left(k) ^ BASE_MAPPING_2B[(new_base & 3) as usize]
            }
            fn nthash_canonical_init(&self, fhash: &mut u64, rhash: &mut u64) -> (u64, u8) {
                *fhash = 0;
                *rhash = 0;
                let k = self.get_nb_base() as u32;
                for i in 0..k {
                    let base = ((self.0 >> (2 * (k - 1 - i))) & 3) as usize;
                    *fhash ^= BASE_MAPPING_2B[base].rotate_left(k - i - 1);
                    *rhash ^= BASE_MAPPING_2B[4 + base].rotate_left(i);
                }
                if *fhash <= *rhash { (*fhash, 0) } else { (*rhash, 1) }
            }
            fn nthash_canonical_cycle(&mut self, new_base: u8, fhash: &mut u64, rhash: &mut u64) -> (u64, u8) {
                *fhash = 0; // kmer.rs:97-98
                *rhash = 0;
                let k = self.get_nb_base() as u32;
                let old = ((self.0 >> (2 * (k - 1))) & 3) as usize;
                *fhash = fhash.rotate_left(1) ^ BASE_MAPPING_2B[old].rotate_left(k) ^ BASE_MAPPING_2B[(new_base & 3) as usize];
                *rhash = rhash.rotate_right(1) ^ BASE_MAPPING_2B[4 + old].rotate_left(k) ^ BASE_MAPPING_2B[4 + (new_base & 3) as usize].rotate_left(k - 1);
                let _ = self.push(new_base);
                if *fhash <= *rhash { (*fhash, 0) } else { (*rhash, 1) }
            }
            fn nthash_mult_canonical_init(&self, fhash: &mut u64, rhash: &mut u64, hashed: &mut [u64]) -> u8 {
                let res = self.nthash_canonical_init(fhash, rhash);
                hashed[0] = res.0;
                from_one_hash_val_to_mult_hash(self.get_nb_base() as u64, hashed);
                res.1
            }
            fn nthash_mult_canonical_cycle(&mut self, new_base: u8, fhash: &mut u64, rhash: &mut u64, hashed: &mut [u64]) -> u8 {
                let res = self.nthash_canonical_cycle(new_base, fhash, rhash);
                hashed[0] = res.0;
                from_one_hash_val_to_mult_hash(self.get_nb_base() as u64, hashed);
                res.1
            }
        }
    };
}
implement_nthash_for!(Kmer32bit);
implement_nthash_for!(Kmer16b32bit);

/// canonical ntHash (and `n_multi - 1` derived hashes) of every k-mer of every sequence, in order: (hashes k-mer major, strands)
pub fn nthash_canonical_batch(vseq: &[&Sequence], k: u8, n_multi: u32) -> (Vec<u64>, Vec<u8>) {
    let b = device_batch(vseq);
    let n = unsafe { ffi::kmu_kmer_count(b.0, k as u32) } as usize;
    let (mut h, mut s) = (vec![0u64; n * n_multi as usize], vec![0u8; n]);
    ffi::check(unsafe { ffi::kmu_nthash_canonical(ffi::ctx(), b.0, k as u32, n_multi, h.as_mut_ptr(), s.as_mut_ptr(), 0) }, "nthash");
    (h, s)
}
