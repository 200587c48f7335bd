//! Reference `src/base/kmertraits.rs:14-52`, unchanged contract.
use std::cmp::Ord;
use std::hash::Hash;
use std::io;

pub trait KmerT {
    fn get_nb_base(&self) -> u8;
    fn reverse_complement(&self) -> Self;
    fn push(&self, base: u8) -> Self;
    fn dump(&self, bufw: &mut dyn io::Write) -> io::Result<usize>;
}

pub trait CompressedKmerT: KmerT + Ord + Copy
where
    Self::Val: Hash + Ord + Copy + Default + std::ops::BitAnd<Output = Self::Val>,
{
    type Val;
    /// the word type of the engine: KMU_KMER32 / KMU_KMER16B32 / KMU_KMER64 / KMU_KMERAA32 / KMU_KMERAA64
    const KMU_TYPE: i32;
    fn get_nb_base_max() -> usize;
    fn get_compressed_value(&self) -> Self::Val;
    fn get_uncompressed_kmer(&self) -> Vec<u8>;
    fn get_bitsize(&self) -> usize;
}

pub trait KmerBuilder<Kmer: CompressedKmerT> {
    fn build(val: <Kmer as CompressedKmerT>::Val, kmer_size: u8) -> Kmer;
}
