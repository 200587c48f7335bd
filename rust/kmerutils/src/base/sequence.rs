//! Reference `src/base/sequence.rs:14-229`: bases packed 4 per byte, first base in the two most significant bits.
use crate::ffi;

#[derive(Clone, Debug)]
pub struct Sequence {
    seq: Vec<u8>,
    nb_base: usize,
}

impl Sequence {
    /// `Sequence::new(raw, 2)` (:25-106): packed on the GPU; panics on a non-ACGT character (alphabet.rs:125)
    pub fn new(raw: &[u8], nb_bits: u8) -> Sequence {
        Sequence::new_batch(&[raw], nb_bits).pop().unwrap()
    }
    /// many reads with one upload and one kernel (what a feeder wants)
    pub fn new_batch(raws: &[&[u8]], nb_bits: u8) -> Vec<Sequence> {
        assert_eq!(nb_bits, 2, "only the 2-bit alphabet is on the GPU path");
        let mut off = vec![0u64; raws.len() + 1];
        let mut ascii = Vec::new();
        for (i, r) in raws.iter().enumerate() {
            ascii.extend_from_slice(r);
            off[i + 1] = ascii.len() as u64;
        }
        ascii.push(0);
        let mut b = std::ptr::null_mut();
        ffi::check(unsafe { ffi::kmu_seqbatch_from_ascii(ffi::ctx(), ascii.as_ptr(), off.as_ptr(), raws.len() as u64, 0, std::ptr::null_mut(), &mut b) },
                   "Sequence::new");
        let batch = ffi::DeviceBatch(b);
        let mut packed = vec![0u8; unsafe { ffi::kmu_seqbatch_packed_bytes(batch.0) } as usize];
        let (mut boff, mut nb) = (vec![0u64; raws.len()], vec![0u64; raws.len()]);
        ffi::check(unsafe { ffi::kmu_seqbatch_download(ffi::ctx(), batch.0, packed.as_mut_ptr(), boff.as_mut_ptr(), nb.as_mut_ptr()) }, "Sequence::new");
        (0..raws.len())
            .map(|i| Sequence { seq: packed[boff[i] as usize..boff[i] as usize + (nb[i] as usize + 3) / 4].to_vec(), nb_base: nb[i] as usize })
            .collect()
    }
    pub fn from_packed(seq: Vec<u8>, nb_base: usize) -> Sequence { Sequence { seq, nb_base } }
    pub fn size(&self) -> usize { self.nb_base }
    pub fn nb_bits_by_base(&self) -> u8 { 2 }
    pub fn packed(&self) -> &[u8] { &self.seq }
    pub fn get_base(&self, pos: usize) -> u8 {
        assert!(pos < self.nb_base);
        (self.seq[pos >> 2] >> (6 - 2 * (pos & 3))) & 3
    }
    pub fn decompress(&self) -> Vec<u8> { (0..self.nb_base).map(|i| b"ACGT"[self.get_base(i) as usize]).collect() }
}

/// `&[&Sequence]` -> device batch
pub fn device_batch(vseq: &[&Sequence]) -> ffi::DeviceBatch {
    let ptrs: Vec<*const u8> = vseq.iter().map(|s| s.packed().as_ptr()).collect();
    let nb: Vec<u64> = vseq.iter().map(|s| s.size() as u64).collect();
    ffi::DeviceBatch::from_packed(&ptrs, &nb)
}
