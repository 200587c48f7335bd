//! Reference `src/base/kmer16b32bit.rs`: exactly 16 bases in a u32.
use super::kmertraits::*;
use crate::devhash::{int32_hash, RawWord};
use crate::ffi;
use std::io;

#[derive(Clone, Copy, Debug, Hash, PartialEq, Eq, PartialOrd, Ord)]
pub struct Kmer16b32bit(pub u32);

impl KmerT for Kmer16b32bit {
    fn get_nb_base(&self) -> u8 { 16 }
    fn reverse_complement(&self) -> Kmer16b32bit {
        let r = (!self.0).reverse_bits();
        Kmer16b32bit(((r & 0x5555_5555) << 1) | ((r & 0xAAAA_AAAA) >> 1))
    }
    fn push(&self, base: u8) -> Kmer16b32bit { Kmer16b32bit((self.0 << 2) | (base as u32 & 3)) }
    fn dump(&self, bufw: &mut dyn io::Write) -> io::Result<usize> { bufw.write(&self.0.to_ne_bytes()) }
}
impl CompressedKmerT for Kmer16b32bit {
    type Val = u32;
    const KMU_TYPE: i32 = ffi::KMU_KMER16B32;
    fn get_nb_base_max() -> usize { 16 }
    fn get_compressed_value(&self) -> u32 { self.0 }
    fn get_uncompressed_kmer(&self) -> Vec<u8> { (0..16).map(|i| b"ACGT"[((self.0 >> (2 * (15 - i))) & 3) as usize]).collect() }
    fn get_bitsize(&self) -> usize { 32 }
}
impl KmerBuilder<Kmer16b32bit> for Kmer16b32bit {
    fn build(val: u32, kmer_size: u8) -> Kmer16b32bit {
        if kmer_size != 16 { panic!("Kmer16b32bit has 16 bases!!"); }
        Kmer16b32bit(val)
    }
}
impl RawWord for Kmer16b32bit {
    fn raw(&self) -> u32 { self.0 }
    fn invhash(v: u32) -> u32 { int32_hash(v) }
    fn value_mask(&self) -> u32 { u32::MAX }
}
