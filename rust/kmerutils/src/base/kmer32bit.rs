//! Reference `src/base/kmer32bit.rs`: up to 14 bases in a u32, the number of bases in the top 4 bits.
use super::kmertraits::*;
use crate::devhash::{int32_hash, RawWord};
use crate::ffi;
use std::io;

#[derive(Clone, Copy, Debug, Hash, PartialEq, Eq)]
pub struct Kmer32bit(pub u32);

impl Kmer32bit {
    pub fn new(nb_bases: u8) -> Kmer32bit {
        if nb_bases >= 15 { panic!("Kmer32bit cannot store more than 14 bases"); }
        Kmer32bit((nb_bases as u32) << 28)
    }
}
impl PartialOrd for Kmer32bit { fn partial_cmp(&self, o: &Self) -> Option<std::cmp::Ordering> { Some(self.cmp(o)) } }
impl Ord for Kmer32bit {  // number of bases first, then the value field (:47-55)
    fn cmp(&self, o: &Self) -> std::cmp::Ordering {
        (self.0 & 0xF000_0000).cmp(&(o.0 & 0xF000_0000)).then((self.0 & 0x0FFF_FFFF).cmp(&(o.0 & 0x0FFF_FFFF)))
    }
}
impl KmerT for Kmer32bit {
    fn get_nb_base(&self) -> u8 { (self.0 >> 28) as u8 }
    fn reverse_complement(&self) -> Kmer32bit {
        let nb = self.0 >> 28;
        let r = (!self.0).reverse_bits();
        let r = ((r & 0x5555_5555) << 1) | ((r & 0xAAAA_AAAA) >> 1);
        let r = if nb > 0 { r >> (32 - 2 * nb) } else { 0 };
        Kmer32bit((r & 0x0FFF_FFFF) | (self.0 & 0xF000_0000))
    }
    fn push(&self, base: u8) -> Kmer32bit {
        let mask = (1u32 << (2 * self.get_nb_base())) - 1;
        Kmer32bit((((self.0 << 2) & mask) | (base as u32 & 3)) | (self.0 & 0xF000_0000))
    }
    fn dump(&self, bufw: &mut dyn io::Write) -> io::Result<usize> { bufw.write(&self.0.to_ne_bytes()) }
}
impl CompressedKmerT for Kmer32bit {
    type Val = u32;
    const KMU_TYPE: i32 = ffi::KMU_KMER32;
    fn get_nb_base_max() -> usize { 14 }
    fn get_compressed_value(&self) -> u32 { self.0 & 0x0FFF_FFFF }  // :173-178
    fn get_uncompressed_kmer(&self) -> Vec<u8> {
        let nb = self.get_nb_base() as u32;
        (0..nb).map(|i| b"ACGT"[((self.0 >> (2 * (nb - 1 - i))) & 3) as usize]).collect()
    }
    fn get_bitsize(&self) -> usize { 32 }
}
impl KmerBuilder<Kmer32bit> for Kmer32bit {
    fn build(val: u32, kmer_size: u8) -> Kmer32bit { Kmer32bit(Kmer32bit::new(kmer_size).0 | (val & 0x0FFF_FFFF)) }
}
impl RawWord for Kmer32bit {
    fn raw(&self) -> u32 { self.0 }
    fn invhash(v: u32) -> u32 { int32_hash(v) }
    fn value_mask(&self) -> u32 { (1u32 << (2 * self.get_nb_base())) - 1 }
}
