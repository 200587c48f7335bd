//! Reference `src/base/kmer.rs`: re-exports (:22-24), `KmerCoord` (:30-37); the `NtHash` implementations of :45-145 live
//! in `nthash.rs` beside the trait.
pub use super::nthash::*;
pub use super::{kmer16b32bit::*, kmer32bit::*, kmer64bit::*};

/// position of a k-mer in a read set
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct KmerCoord {
    pub read_num: u32,
    pub pos: u32,
}
