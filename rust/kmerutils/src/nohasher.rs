//! Reference `src/nohasher.rs:11-48`: the hasher of already-hashed k-mers.  `write` keeps the reference's seed rule: the
//! bytes of a 4- or 8-byte key are read most-significant first into the u64 the sketchers seed their generators with
//! (KAT 0xa4d66083).  The device does the same in `nohash_seed()` (kmu_device.cuh); this type exists
//! for host-side maps keyed by k-mers.
use std::hash::Hasher;

#[derive(Default)]
pub struct NoHashHasher(u64);

impl Hasher for NoHashHasher {
    #[inline]
    fn write(&mut self, bytes: &[u8]) {
        self.0 = match bytes.len() {
            4 => u32::from_be_bytes(bytes.try_into().unwrap()) as u64,
            8 => u64::from_be_bytes(bytes.try_into().unwrap()),
            _ => panic!("bad slice len in NoHashHasher write"),
        };
    }
    fn finish(&self) -> u64 { self.0 }
}
