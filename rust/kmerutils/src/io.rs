//! Reference `src/io.rs:12-72` (`parse_with_needletail`): a FASTA/FASTQ file (plain or gzip) into `Vec<Sequence>`, reads
//! holding a non-ACGT character dropped and counted.  Here the records come from the multi-threaded feeder of the
//! library (`kmu_ingest_*`: reader thread + parser threads, ordered pinned packs) and every pack is 2-bit packed on the
//! GPU in one call; `for_each_pack` is the streaming form `datasketcher` wants (reference `src/bin/datasketcher.rs:358-388`).
use crate::base::sequence::Sequence;
use crate::ffi;
use std::ffi::CString;

/// the one field of `parsearg::ParseFastqArgs` this path reads besides the file name
pub struct ParseFastqArgs {
    pub filename: String,
    pub nb_bits_by_base: u8,
}

/// one pack of records as the feeder hands it out: concatenated ASCII + offsets (nseq + 1)
pub struct Pack<'a> {
    pub ascii: &'a [u8],
    pub offsets: &'a [u64],
}

/// streams the file pack by pack; `f` may upload the pack (`Sequence::new_batch`, `kmu_seqbatch_from_ascii`) while the
/// parser threads fill the next ones
pub fn for_each_pack<F: FnMut(Pack)>(filename: &str, nthreads: u32, mut f: F) -> Result<(u64, u64), &'static str> {
    let cpath = CString::new(filename).map_err(|_| "bad file name")?;
    let mut rd = std::ptr::null_mut();
    if unsafe { ffi::kmu_ingest_open(cpath.as_ptr(), nthreads, 0, &mut rd) } != ffi::KMU_OK { return Err("file does not exist"); }
    let (mut nseq_total, mut nbases) = (0u64, 0u64);
    loop {
        let (mut ascii, mut off, mut nseq, mut token) = (std::ptr::null(), std::ptr::null(), 0u64, std::ptr::null_mut());
        let rc = unsafe { ffi::kmu_ingest_next(rd, &mut ascii, &mut off, &mut nseq, &mut token) };
        if rc != ffi::KMU_OK { unsafe { ffi::kmu_ingest_close(rd) }; return Err("invalid record"); }
        if nseq == 0 { break; }
        let offsets = unsafe { std::slice::from_raw_parts(off, nseq as usize + 1) };
        let bytes = unsafe { std::slice::from_raw_parts(ascii, offsets[nseq as usize] as usize) };
        nseq_total += nseq;
        nbases += offsets[nseq as usize];
        f(Pack { ascii: bytes, offsets });
        unsafe { ffi::kmu_ingest_release(rd, token) };
    }
    unsafe { ffi::kmu_ingest_close(rd) };
    Ok((nseq_total, nbases))
}

pub fn parse_with_needletail(parsed_args: ParseFastqArgs) -> std::result::Result<Vec<Sequence>, &'static str> {
    let mut seq_array: Vec<Sequence> = Vec::new();
    let nb_bits = parsed_args.nb_bits_by_base;
    for_each_pack(&parsed_args.filename, 0, |p| {
        let reads: Vec<&[u8]> = p.offsets.windows(2).map(|w| &p.ascii[w[0] as usize..w[1] as usize]).collect();
        seq_array.extend(Sequence::new_batch(&reads, nb_bits)); // the feeder has already dropped reads with non-ACGT characters
    })?;
    seq_array.shrink_to_fit();
    Ok(seq_array)
}
