//! The five hash closures the reference passes to its sketchers as zero-sized markers.
//! `fhash: F where F: Fn(&Kmer) -> Kmer::Val` becomes `fhash: H where H: DeviceKmerHash<Kmer>`: the kind crosses the FFI,
//! `call` is the closure itself for host-side use.  The trait is sealed: an arbitrary closure cannot run on the GPU and
//! there is no CPU fallback.
use crate::base::kmertraits::{CompressedKmerT, KmerT};
use crate::ffi;

mod sealed { pub trait Sealed {} }

pub trait DeviceKmerHash<Kmer: CompressedKmerT>: sealed::Sealed + Copy + Send + Sync {
    const KIND: i32;
    fn call(&self, kmer: &Kmer) -> Kmer::Val;
}

/// `|k| k.0` (seqsketchjaccard.rs:775)
#[derive(Clone, Copy, Default)] pub struct IdentityRaw;
/// `|k| k.get_compressed_value() & mask` (setsketchert.rs:1098-1104)
#[derive(Clone, Copy, Default)] pub struct MaskedValue;
/// `|k| intNN_hash(k.reverse_complement().min(*k).0)` (datasketcher.rs:222-226)
#[derive(Clone, Copy, Default)] pub struct CanonicalInvHash;
/// `|k| k.reverse_complement().min(*k).0` (kmercount.rs:313)
#[derive(Clone, Copy, Default)] pub struct CanonicalRaw;
/// `|k| intNN_hash(k.0)` (minhash.rs:226)
#[derive(Clone, Copy, Default)] pub struct InvHash;

impl sealed::Sealed for IdentityRaw {}
impl sealed::Sealed for MaskedValue {}
impl sealed::Sealed for CanonicalInvHash {}
impl sealed::Sealed for CanonicalRaw {}
impl sealed::Sealed for InvHash {}

/// word `.0` of a k-mer and the invertible hash of its width
pub trait RawWord: CompressedKmerT {
    fn raw(&self) -> Self::Val;
    fn invhash(v: Self::Val) -> Self::Val;
    fn value_mask(&self) -> Self::Val;
}

impl<K: RawWord + KmerT> DeviceKmerHash<K> for IdentityRaw {
    const KIND: i32 = ffi::KMU_HASH_IDENTITY_RAW;
    fn call(&self, k: &K) -> K::Val { k.raw() }
}
impl<K: RawWord + KmerT> DeviceKmerHash<K> for MaskedValue {
    const KIND: i32 = ffi::KMU_HASH_MASKED_VALUE;
    fn call(&self, k: &K) -> K::Val { k.get_compressed_value() & k.value_mask() }
}
impl<K: RawWord + KmerT> DeviceKmerHash<K> for CanonicalInvHash {
    const KIND: i32 = ffi::KMU_HASH_CANON_INVHASH;
    fn call(&self, k: &K) -> K::Val { K::invhash(k.reverse_complement().min(*k).raw()) }
}
impl<K: RawWord + KmerT> DeviceKmerHash<K> for CanonicalRaw {
    const KIND: i32 = ffi::KMU_HASH_CANON_RAW;
    fn call(&self, k: &K) -> K::Val { k.reverse_complement().min(*k).raw() }
}
impl<K: RawWord + KmerT> DeviceKmerHash<K> for InvHash {
    const KIND: i32 = ffi::KMU_HASH_INVHASH;
    fn call(&self, k: &K) -> K::Val { K::invhash(k.raw()) }
}

/// probminhash::invhash (Thomas Wang's invertible mixes)
pub fn int32_hash(mut key: u32) -> u32 {
    key = key.wrapping_add(!(key << 15)); key ^= key >> 10; key = key.wrapping_add(key << 3);
    key ^= key >> 6; key = key.wrapping_add(!(key << 11)); key ^= key >> 16; key
}
pub fn int64_hash(mut key: u64) -> u64 {
    key = (!key).wrapping_add(key << 21); key ^= key >> 24; key = key.wrapping_add(key << 3).wrapping_add(key << 8);
    key ^= key >> 14; key = key.wrapping_add(key << 2).wrapping_add(key << 4); key ^= key >> 28; key.wrapping_add(key << 31)
}
