//! Reference `src/sketching/mod.rs`: the sketcher modules of the hot path.  `minhash` (bottom-k, only used by the anchors)
//! and `nbkmerguess` are outside the path and not mirrored.
pub mod seqblocksketch;
pub mod seqminhash;
pub mod seqsketchjaccard;
pub mod setsketchert;

pub use setsketchert::SeqSketcherT;
