//! Reference `src/base/mod.rs:10-29`: same sub-modules, same re-exports.
pub mod nthash;

pub use alphabet::*;
pub use kmer16b32bit::*;
pub use kmer32bit::*;
pub use kmer64bit::*;
pub use kmer::*;
pub use kmertraits::*;
pub use sequence::*;

pub mod alphabet;
pub mod kmer;
pub mod kmer16b32bit;
pub mod kmer32bit;
pub mod kmer64bit;
pub mod kmercount;
pub mod kmergenerator;
pub mod kmertraits;
pub mod sequence;
