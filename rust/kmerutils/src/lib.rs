//! kmerutils over the B200 engine: same module paths, type names and method signatures as the reference crate for its
//! data-parallel hot path (reference `src/lib.rs:10-38`); every entry point that computes calls the CUDA library through
//! `ffi` -- there is no CPU implementation behind the sketchers, the generators or the counters.
//!
//! The one visible difference: the hash closure `fhash: Fn(&Kmer) -> Kmer::Val` of the sketcher entry points cannot cross
//! into CUDA.  The entry points take `impl DeviceKmerHash<Kmer>` instead: a sealed trait implemented by five zero-sized
//! markers, one per closure the reference itself passes (`devhash`).  Each marker also carries the closure (`call`) so
//! that host-side code and tests can evaluate it.
pub mod ffi;
pub mod devhash;

pub mod base;
pub mod aautils;
pub mod io;
pub mod sketcharg;
pub mod nohasher;
pub mod sketching;

pub mod prelude {
    pub use crate::base::kmergenerator::*;
    pub use crate::base::*;
    pub use crate::devhash::*;
}
