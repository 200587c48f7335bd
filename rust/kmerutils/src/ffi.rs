//! `extern "C"` declarations of include/kmerutils_b200.h (the subset the shim binds) + the process-wide context.
#![allow(non_camel_case_types, dead_code)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_void};
use std::sync::OnceLock;

#[repr(C)] pub struct kmu_ctx { _p: [u8; 0] }
#[repr(C)] pub struct kmu_seqbatch { _p: [u8; 0] }
#[repr(C)] pub struct kmu_counter { _p: [u8; 0] }
#[repr(C)] pub struct kmu_sigdump { _p: [u8; 0] }
#[repr(C)] pub struct kmu_ingest { _p: [u8; 0] }
#[repr(C)] pub struct kmu_setsketch_params { pub b: f64, pub m: u64, pub a: f64, pub q: u64 }

pub const KMU_OK: i32 = 0;
pub const KMU_KMER32: i32 = 0; pub const KMU_KMER16B32: i32 = 1; pub const KMU_KMER64: i32 = 2;
pub const KMU_KMERAA32: i32 = 3; pub const KMU_KMERAA64: i32 = 4;
pub const KMU_HASH_IDENTITY_RAW: i32 = 0; pub const KMU_HASH_MASKED_VALUE: i32 = 1; pub const KMU_HASH_CANON_INVHASH: i32 = 2;
pub const KMU_HASH_CANON_RAW: i32 = 3; pub const KMU_HASH_INVHASH: i32 = 4;
pub const KMU_HASHER_NOHASH: i32 = 0; pub const KMU_HASHER_FNV: i32 = 1;

extern "C" {
    pub fn kmu_ctx_create(device: i32, ctx: *mut *mut kmu_ctx) -> i32;
    pub fn kmu_ctx_destroy(ctx: *mut kmu_ctx);
    pub fn kmu_last_error() -> *const c_char;
    pub fn kmu_seqbatch_from_ptrs(ctx: *mut kmu_ctx, seq_ptrs: *const *const u8, nbases: *const u64, nseq: u64, batch: *mut *mut kmu_seqbatch) -> i32;
    pub fn kmu_seqbatch_from_ascii(ctx: *mut kmu_ctx, ascii: *const u8, ascii_off: *const u64, nseq: u64, drop_invalid: i32,
                                   invalid_counts: *mut u64, batch: *mut *mut kmu_seqbatch) -> i32;
    pub fn kmu_seqbatch_from_aa(ctx: *mut kmu_ctx, ascii: *const u8, ascii_off: *const u64, nseq: u64, drop_invalid: i32,
                                invalid_counts: *mut u64, batch: *mut *mut kmu_seqbatch) -> i32;
    pub fn kmu_seqbatch_slices(ctx: *mut kmu_ctx, src: *const kmu_seqbatch, seq_idx: *const u64, begin: *const u64, end: *const u64,
                               nslices: u64, batch: *mut *mut kmu_seqbatch) -> i32;
    pub fn kmu_seqbatch_download(ctx: *mut kmu_ctx, batch: *const kmu_seqbatch, packed: *mut u8, byte_off: *mut u64, nbases: *mut u64) -> i32;
    pub fn kmu_seqbatch_packed_bytes(batch: *const kmu_seqbatch) -> u64;
    pub fn kmu_seqbatch_destroy(batch: *mut kmu_seqbatch);
    pub fn kmu_kmer_count(batch: *const kmu_seqbatch, k: u32) -> u64;
    pub fn kmu_generate_kmers(ctx: *mut kmu_ctx, batch: *const kmu_seqbatch, k: u32, kmer_type: i32, hash_kind: i32, out: *mut c_void,
                              out_off: *mut u64, out_on_device: i32) -> i32;
    pub fn kmu_nthash_canonical(ctx: *mut kmu_ctx, batch: *const kmu_seqbatch, k: u32, n_multi: u32, out_hash: *mut u64,
                                out_strand: *mut u8, out_on_device: i32) -> i32;
    pub fn kmu_sketch_pmh3a(ctx: *mut kmu_ctx, batch: *const kmu_seqbatch, k: u32, kmer_type: i32, hash_kind: i32, m: u32,
                            sig: *mut c_void, sig_on_device: i32) -> i32;
    pub fn kmu_sketch_pmh3a_host_ptrs(ctx: *mut kmu_ctx, seq_ptrs: *const *const u8, nbases: *const u64, nseq: u64, k: u32,
                                      kmer_type: i32, hash_kind: i32, m: u32, sig: *mut c_void) -> i32;
    pub fn kmu_sketch_pmh3a_whole(ctx: *mut kmu_ctx, batch: *const kmu_seqbatch, k: u32, kmer_type: i32, hash_kind: i32, m: u32,
                                  sig: *mut c_void, sig_on_device: i32) -> i32;
    pub fn kmu_sketch_superminhash(ctx: *mut kmu_ctx, batch: *const kmu_seqbatch, k: u32, kmer_type: i32, hash_kind: i32, m: u32,
                                   key_hasher: i32, sig_bytes: i32, sig: *mut c_void, sig_on_device: i32) -> i32;
    pub fn kmu_sketch_superminhash_whole(ctx: *mut kmu_ctx, batch: *const kmu_seqbatch, k: u32, kmer_type: i32, hash_kind: i32, m: u32,
                                         key_hasher: i32, sig_bytes: i32, sig: *mut c_void, sig_on_device: i32) -> i32;
    pub fn kmu_sketch_setsketch(ctx: *mut kmu_ctx, batch: *const kmu_seqbatch, k: u32, kmer_type: i32, hash_kind: i32,
                                params: *const kmu_setsketch_params, sig_bytes: i32, whole: i32, sig: *mut c_void, sig_on_device: i32) -> i32;
    pub fn kmu_signature_jaccard(ctx: *mut kmu_ctx, sig_a: *const c_void, na: u64, sig_b: *const c_void, nb: u64, m: u32,
                                 slot_bytes: i32, out: *mut f64, on_device: i32) -> i32;
    pub fn kmu_count_create(ctx: *mut kmu_ctx, k: u32, kmer_type: i32, count_bits: u32, capacity: u64, counter: *mut *mut kmu_counter) -> i32;
    pub fn kmu_count_destroy(counter: *mut kmu_counter);
    pub fn kmu_count_insert_seqs(ctx: *mut kmu_ctx, counter: *mut kmu_counter, batch: *const kmu_seqbatch, canonical: i32) -> i32;
    pub fn kmu_count_insert_kmers(ctx: *mut kmu_ctx, counter: *mut kmu_counter, kmers: *const c_void, n: u64, on_device: i32) -> i32;
    pub fn kmu_count_query(ctx: *mut kmu_ctx, counter: *const kmu_counter, kmers: *const c_void, n: u64, counts: *mut u32, on_device: i32) -> i32;
    pub fn kmu_count_stats(ctx: *mut kmu_ctx, counter: *const kmu_counter, nb_distinct: *mut u64, nb_unique: *mut u64,
                           nb_inserted: *mut u64, hist256: *mut u64) -> i32;
    pub fn kmu_count_export(ctx: *mut kmu_ctx, counter: *const kmu_counter, min_count: u32, kmers: *mut c_void, counts: *mut u32,
                            cap: u64, n_out: *mut u64) -> i32;
    pub fn kmu_count_partition(ctx: *mut kmu_ctx, batch: *const kmu_seqbatch, k: u32, kmer_type: i32, canonical: i32, nparts: u32,
                               kmers_out: *mut c_void, part_counts: *mut u64, out_on_device: i32) -> i32;
    pub fn kmu_count_dump_multiple(ctx: *mut kmu_ctx, counter: *const kmu_counter, path: *const c_char, count_bytes: i32, nb_dumped: *mut u64) -> i32;
    pub fn kmu_sigdump_create(path: *const c_char, sketch_size: u32, kmer_size: u32, dump: *mut *mut kmu_sigdump) -> i32;
    pub fn kmu_sigdump_write(dump: *mut kmu_sigdump, sig: *const u32, nseq: u64) -> i32;
    pub fn kmu_sigdump_close(dump: *mut kmu_sigdump) -> i32;
    pub fn kmu_ingest_open(path: *const c_char, nthreads: u32, block_bytes: u64, reader: *mut *mut kmu_ingest) -> i32;
    pub fn kmu_ingest_next(reader: *mut kmu_ingest, ascii: *mut *const u8, ascii_off: *mut *const u64, nseq: *mut u64, token: *mut *mut c_void) -> i32;
    pub fn kmu_ingest_release(reader: *mut kmu_ingest, token: *mut c_void) -> i32;
    pub fn kmu_ingest_close(reader: *mut kmu_ingest);
}

struct CtxPtr(*mut kmu_ctx);
unsafe impl Send for CtxPtr {}
unsafe impl Sync for CtxPtr {}
static CTX: OnceLock<CtxPtr> = OnceLock::new();

/// The context of GPU `KMERUTILS_DEVICE` (default 0); the C layer serialises calls on a context.
pub fn ctx() -> *mut kmu_ctx {
    CTX.get_or_init(|| {
        let dev = std::env::var("KMERUTILS_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut c = std::ptr::null_mut();
        check(unsafe { kmu_ctx_create(dev, &mut c) }, "kmu_ctx_create");
        CtxPtr(c)
    }).0
}

/// A non-zero status is what the reference turns into a panic (bad k for the type, non-ACGT base, empty sequence ...).
pub fn check(rc: i32, what: &str) {
    if rc != KMU_OK {
        let msg = unsafe { CStr::from_ptr(kmu_last_error()) }.to_string_lossy().into_owned();
        panic!("{what}: {msg}");
    }
}

/// RAII device batch made of `&[&Sequence]` (kmu_seqbatch_from_ptrs).
pub struct DeviceBatch(pub *mut kmu_seqbatch);
impl DeviceBatch {
    pub fn from_packed(ptrs: &[*const u8], nbases: &[u64]) -> Self {
        let mut b = std::ptr::null_mut();
        check(unsafe { kmu_seqbatch_from_ptrs(ctx(), ptrs.as_ptr(), nbases.as_ptr(), ptrs.len() as u64, &mut b) }, "kmu_seqbatch_from_ptrs");
        DeviceBatch(b)
    }
}
impl Drop for DeviceBatch {
    fn drop(&mut self) { unsafe { kmu_seqbatch_destroy(self.0) } }
}

extern "C" {
    pub fn kmu_sketch_pmh3a_groups(ctx: *mut kmu_ctx, batch: *const kmu_seqbatch, group_sizes: *const u64, ngroups: u64, k: u32, kmer_type: i32,
                                   hash_kind: i32, m: u32, sig: *mut c_void, sig_on_device: i32) -> i32;
    pub fn kmu_blockdump_create(path: *const c_char, sketch_size: u32, kmer_size: u32, block_size: u32, dump: *mut *mut kmu_sigdump) -> i32;
    pub fn kmu_blockdump_write(dump: *mut kmu_sigdump, sig: *const u32, numseq: *const u32, numblock: *const u32, nblocks: u64) -> i32;
}
