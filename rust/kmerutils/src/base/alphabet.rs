//! Reference `src/base/alphabet.rs`: the 2-bit alphabet (A=0 C=1 G=2 T=3, case-insensitive; :117-169) and the helpers the
//! feeders use (:28-31).  Packing itself runs on the GPU (`Sequence::new` -> kmu_seqbatch_from_ascii).
pub fn is_acgt(c: u8) -> bool { matches!(c & 0xDF, b'A' | b'C' | b'G' | b'T') }
pub fn count_non_acgt(seq: &[u8]) -> usize { seq.iter().filter(|c| !is_acgt(**c)).count() }

pub trait BaseCompress {
    fn encode(&self, c: u8) -> u8;
    fn decode(&self, c: u8) -> u8;
    fn get_nb_bits(&self) -> u8;
    fn is_valid_base(&self, c: u8) -> bool;
}

#[derive(Default, Clone, Copy)]
pub struct Alphabet2b;
impl Alphabet2b { pub fn new() -> Self { Alphabet2b } }
impl BaseCompress for Alphabet2b {
    fn encode(&self, c: u8) -> u8 {
        match c & 0xDF { b'A' => 0, b'C' => 1, b'G' => 2, b'T' => 3, _ => panic!("pattern not a code in alphabet_2b") }
    }
    fn decode(&self, c: u8) -> u8 { b"ACGT"[(c & 3) as usize] }
    fn get_nb_bits(&self) -> u8 { 2 }
    fn is_valid_base(&self, c: u8) -> bool { is_acgt(c) }
}
