//! Reference `src/aautils/mod.rs`: amino-acid k-mers (5 bits per residue) and their sketchers.
pub mod kmeraa;
pub mod setsketchert;
