//! Reference `src/base/kmer64bit.rs`: up to 32 bases in a u64, the number of bases beside it.
use super::kmertraits::*;
use crate::devhash::{int64_hash, RawWord};
use crate::ffi;
use std::io;

#[derive(Clone, Copy, Debug, Hash, PartialEq, Eq)]
pub struct Kmer64bit(pub u64, pub u8);

impl PartialOrd for Kmer64bit { fn partial_cmp(&self, o: &Self) -> Option<std::cmp::Ordering> { Some(self.cmp(o)) } }
impl Ord for Kmer64bit { fn cmp(&self, o: &Self) -> std::cmp::Ordering { self.1.cmp(&o.1).then(self.0.cmp(&o.0)) } }  // :45-53
impl KmerT for Kmer64bit {
    fn get_nb_base(&self) -> u8 { self.1 }
    fn reverse_complement(&self) -> Kmer64bit {
        let r = (!self.0).reverse_bits();
        let r = ((r & 0x5555_5555_5555_5555) << 1) | ((r & 0xAAAA_AAAA_AAAA_AAAA) >> 1);
        Kmer64bit(if self.1 > 0 { r >> (64 - 2 * self.1 as u32) } else { 0 }, self.1)
    }
    fn push(&self, base: u8) -> Kmer64bit {
        let mask = if self.1 >= 32 { u64::MAX } else { (1u64 << (2 * self.1)) - 1 };  // k = 32 is defined here (the reference wraps, :75)
        Kmer64bit(((self.0 << 2) & mask) | (base as u64 & 3), self.1)
    }
    fn dump(&self, bufw: &mut dyn io::Write) -> io::Result<usize> {  // 1 byte k + 8 bytes value (:98-104)
        bufw.write_all(&[self.1])?;
        bufw.write_all(&self.0.to_ne_bytes())?;
        Ok(9)
    }
}
impl CompressedKmerT for Kmer64bit {
    type Val = u64;
    const KMU_TYPE: i32 = ffi::KMU_KMER64;
    fn get_nb_base_max() -> usize { 32 }
    fn get_compressed_value(&self) -> u64 { self.0 }
    fn get_uncompressed_kmer(&self) -> Vec<u8> {
        let nb = self.1 as u32;
        (0..nb).map(|i| b"ACGT"[((self.0 >> (2 * (nb - 1 - i))) & 3) as usize]).collect()
    }
    fn get_bitsize(&self) -> usize { 64 }
}
impl KmerBuilder<Kmer64bit> for Kmer64bit {
    fn build(val: u64, kmer_size: u8) -> Kmer64bit { Kmer64bit(val, kmer_size) }
}
impl RawWord for Kmer64bit {
    fn raw(&self) -> u64 { self.0 }
    fn invhash(v: u64) -> u64 { int64_hash(v) }
    fn value_mask(&self) -> u64 { if self.1 >= 32 { u64::MAX } else { (1u64 << (2 * self.1)) - 1 } }
}
