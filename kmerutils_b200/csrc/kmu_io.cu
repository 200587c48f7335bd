// kmu_io.cu -- host-side feeders and writers around the GPU path (no device code):
//   * FASTA / FASTQ reader that hands out packs of ACCEPTED reads as one ASCII buffer + offsets, ready
//     for kmu_seqbatch_from_ascii: src/io.rs:12-72 (parse_with_needletail) and readblockseq
//     (src/bin/datasketcher.rs:358-388) drop every read that holds a non-ACGT character;
//   * signature dump writer / reader: SeqSketcher::create_signature_dump + dump_signatures_block_u32 +
//     SigSketchFileReader (src/sketching/seqsketchjaccard.rs:390-414, 572-712);
//   * block signature dump: BlockSeqSketcher::create_signature_dump / dump_blocks
//     (src/sketching/seqblocksketch.rs:59-65, 172-226).
// All integers are little-endian as written by the reference's to_le_bytes().
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "kmu_host.h"

namespace {

const uint32_t MAGIC_SIG_DUMP = 0xceabeadd;       // seqsketchjaccard.rs:572
const uint32_t MAGIC_BLOCKSIG_DUMP = 0xceabbadd;  // seqblocksketch.rs:33

void put_u32(std::vector<uint8_t>& b, uint32_t v) {
    for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i)));
}
uint32_t get_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

inline bool is_acgt(uint8_t c) {  // Alphabet2b::is_valid_base, case-insensitive (alphabet.rs:157-159)
    c &= 0xDF;
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

}  // namespace

struct kmu_sigdump {
    FILE* f = nullptr;
    int kind = 0;  // 0 sequence signatures, 1 block signatures
    uint32_t sketch_size = 0;
    uint64_t rows = 0;
};

struct kmu_fastx {
    gzFile f = nullptr;  // zlib reads plain and gzip-compressed files alike (needletail does the same, io.rs:20-24)
    std::vector<char> buf;
    size_t pos = 0, len = 0;
    bool eof = false, bad_stream = false;
    std::string pending;  // an accepted record that did not fit the previous pack
    bool has_pending = false;
    uint64_t nb_read = 0, nb_bad_read = 0, nb_bases = 0, nb_bad_bases = 0;
    bool fill() {
        if (eof) return false;
        if (pos > 0 && pos < len) std::memmove(buf.data(), buf.data() + pos, len - pos);
        len -= pos;
        pos = 0;
        if (len == buf.size()) buf.resize(buf.size() * 2);
        const int got = gzread(f, buf.data() + len, (unsigned)std::min<size_t>(buf.size() - len, 1u << 30));
        const size_t n = got > 0 ? (size_t)got : 0;
        if (got < 0) bad_stream = true;
        if (n == 0) eof = true;
        len += n;
        return n > 0;
    }
    // next line without its end-of-line characters; false at end of file
    bool line(std::string& out) {
        out.clear();
        for (;;) {
            char* p = (char*)std::memchr(buf.data() + pos, '\n', len - pos);
            if (p) {
                out.append(buf.data() + pos, p - (buf.data() + pos));
                pos = (p - buf.data()) + 1;
                if (!out.empty() && out.back() == '\r') out.pop_back();
                return true;
            }
            out.append(buf.data() + pos, len - pos);
            pos = len;
            if (!fill()) {
                if (!out.empty() && out.back() == '\r') out.pop_back();
                return !out.empty();
            }
        }
    }
    int peek() {
        if (pos >= len && !fill()) return -1;
        return (unsigned char)buf[pos];
    }
};

extern "C" {

// ---- FASTA / FASTQ ---------------------------------------------------------------------------
int32_t kmu_fastx_open(const char* path, kmu_fastx** out) {
    if (!path || !out) return fail(KMU_EINVAL, "null argument");
    gzFile f = gzopen(path, "rb");
    if (!f) return fail(KMU_EINVAL, "file does not exist: %s", path);
    gzbuffer(f, 1u << 20);
    auto* h = new kmu_fastx();
    h->f = f;
    h->buf.resize(1 << 22);
    *out = h;
    return KMU_OK;
}

void kmu_fastx_close(kmu_fastx* h) {
    if (!h) return;
    if (h->f) gzclose(h->f);
    delete h;
}

// Reads up to max_seqs ACCEPTED records (at most ascii_cap bytes of sequence): their bases go into `ascii` back to
// back, ascii_off[0..n] delimits them.  Records holding any non-ACGT character are skipped and counted
// (io.rs:41-48, datasketcher.rs:367-371).  *nseq_out == 0 at end of file.
int32_t kmu_fastx_next_pack(kmu_fastx* h, uint64_t max_seqs, uint8_t* ascii, uint64_t ascii_cap, uint64_t* ascii_off,
                            uint64_t* nseq_out) {
    if (!h || !ascii || !ascii_off || !nseq_out) return fail(KMU_EINVAL, "null argument");
    uint64_t n = 0, used = 0;
    ascii_off[0] = 0;
    std::string l, seq;
    if (h->has_pending) {
        if (h->pending.size() > ascii_cap) return fail(KMU_EOVERFLOW, "a record of %zu bases does not fit the pack buffer", h->pending.size());
        std::memcpy(ascii, h->pending.data(), h->pending.size());
        used = h->pending.size();
        ascii_off[++n] = used;
        h->has_pending = false;
        h->pending.clear();
    }
    while (n < max_seqs) {
        int c = h->peek();
        while (c == '\n' || c == '\r') {  // blank lines between records
            h->pos++;
            c = h->peek();
        }
        if (c < 0) break;
        if (c != '>' && c != '@') return fail(KMU_EINVAL, "invalid record: expected '>' or '@', got '%c'", c);
        h->line(l);  // header
        seq.clear();
        if (c == '>') {  // FASTA: sequence lines up to the next header
            for (;;) {
                int d = h->peek();
                if (d < 0 || d == '>') break;
                h->line(l);
                seq += l;
            }
        } else {  // FASTQ: sequence lines up to '+', then as many quality characters
            for (;;) {
                int d = h->peek();
                if (d < 0) return fail(KMU_EINVAL, "invalid record: truncated FASTQ");
                if (d == '+') break;
                h->line(l);
                seq += l;
            }
            h->line(l);  // '+' line
            size_t q = 0;
            while (q < seq.size()) {
                if (!h->line(l)) return fail(KMU_EINVAL, "invalid record: truncated FASTQ qualities");
                q += l.size();
            }
        }
        h->nb_read++;
        h->nb_bases += seq.size();
        uint64_t bad = 0;
        for (char ch : seq) bad += !is_acgt((uint8_t)ch);
        h->nb_bad_bases += bad;
        if (bad) {
            h->nb_bad_read++;
            continue;
        }
        if (used + seq.size() > ascii_cap) {
            if (n == 0) return fail(KMU_EOVERFLOW, "a record of %zu bases does not fit the pack buffer", seq.size());
            h->pending.swap(seq);  // first record of the next pack
            h->has_pending = true;
            break;
        }
        std::memcpy(ascii + used, seq.data(), seq.size());
        used += seq.size();
        ascii_off[++n] = used;
    }
    *nseq_out = n;
    if (h->bad_stream) return fail(KMU_EINVAL, "corrupt or truncated compressed stream");
    return KMU_OK;
}

// nb records read, records dropped for a non-ACGT character, bases read, non-ACGT bases seen (io.rs:66-69)
void kmu_fastx_stats(const kmu_fastx* h, uint64_t* nb_read, uint64_t* nb_bad_read, uint64_t* nb_bases, uint64_t* nb_bad_bases) {
    if (!h) return;
    if (nb_read) *nb_read = h->nb_read;
    if (nb_bad_read) *nb_bad_read = h->nb_bad_read;
    if (nb_bases) *nb_bases = h->nb_bases;
    if (nb_bad_bases) *nb_bad_bases = h->nb_bad_bases;
}

// ---- signature dumps ---------------------------------------------------------------------------
int32_t kmu_sigdump_create(const char* path, uint32_t sketch_size, uint32_t kmer_size, kmu_sigdump** out) {
    if (!path || !out) return fail(KMU_EINVAL, "null argument");
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(KMU_EINVAL, "cannot open %s", path);
    std::vector<uint8_t> hdr;
    put_u32(hdr, MAGIC_SIG_DUMP);
    put_u32(hdr, 4);  // sig_size: Vec<u32> signatures (seqsketchjaccard.rs:402)
    put_u32(hdr, sketch_size);
    put_u32(hdr, kmer_size);
    std::fwrite(hdr.data(), 1, hdr.size(), f);
    auto* h = new kmu_sigdump();
    h->f = f;
    h->sketch_size = sketch_size;
    *out = h;
    return KMU_OK;
}

// dump_signatures_block_u32 (seqsketchjaccard.rs:577-585): nseq rows of sketch_size u32, in order
int32_t kmu_sigdump_write(kmu_sigdump* h, const uint32_t* sig, uint64_t nseq) {
    if (!h || h->kind != 0 || (nseq && !sig)) return fail(KMU_EINVAL, "bad signature dump handle / buffer");
    const size_t n = (size_t)nseq * h->sketch_size;
    if (std::fwrite(sig, 4, n, h->f) != n) return fail(KMU_EINVAL, "short write to the signature dump");  // host is little-endian
    h->rows += nseq;
    return KMU_OK;
}

int32_t kmu_blockdump_create(const char* path, uint32_t sketch_size, uint32_t kmer_size, uint32_t block_size,
                             kmu_sigdump** out) {
    if (!path || !out) return fail(KMU_EINVAL, "null argument");
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(KMU_EINVAL, "cannot open %s", path);
    std::vector<uint8_t> hdr;
    put_u32(hdr, MAGIC_BLOCKSIG_DUMP);
    hdr.push_back(4);  // sig_size is ONE byte in the writer (seqblocksketch.rs:80,220: 17-byte header)
    put_u32(hdr, sketch_size);
    put_u32(hdr, kmer_size);
    put_u32(hdr, block_size);
    std::fwrite(hdr.data(), 1, hdr.size(), f);
    auto* h = new kmu_sigdump();
    h->f = f;
    h->kind = 1;
    h->sketch_size = sketch_size;
    *out = h;
    return KMU_OK;
}

// dump_blocks (seqblocksketch.rs:172-187): rows are the blocks of consecutive sequences (numseq non-decreasing);
// per sequence `numseq, nbblocks`, then per block `numseq, numblock, sketch`
int32_t kmu_blockdump_write(kmu_sigdump* h, const uint32_t* sig, const uint32_t* numseq, const uint32_t* numblock,
                            uint64_t nblocks) {
    if (!h || h->kind != 1 || (nblocks && (!sig || !numseq || !numblock))) return fail(KMU_EINVAL, "bad block dump handle / buffer");
    std::vector<uint8_t> rec;
    for (uint64_t i = 0; i < nblocks;) {
        uint64_t j = i;
        while (j < nblocks && numseq[j] == numseq[i]) ++j;
        rec.clear();
        put_u32(rec, numseq[i]);
        put_u32(rec, (uint32_t)(j - i));
        for (uint64_t b = i; b < j; ++b) {
            put_u32(rec, numseq[b]);
            put_u32(rec, numblock[b]);
            const uint32_t* s = sig + (size_t)b * h->sketch_size;
            for (uint32_t t = 0; t < h->sketch_size; ++t) put_u32(rec, s[t]);
        }
        std::fwrite(rec.data(), 1, rec.size(), h->f);
        i = j;
    }
    h->rows += nblocks;
    return KMU_OK;
}

int32_t kmu_sigdump_close(kmu_sigdump* h) {
    if (!h) return KMU_OK;
    int rc = h->f ? std::fclose(h->f) : 0;
    delete h;
    return rc ? fail(KMU_EINVAL, "closing the dump failed") : KMU_OK;
}

// SigSketchFileReader::new + next (seqsketchjaccard.rs:588-712).  The reference's next() reads the bytes and
// returns an EMPTY Vec (:708-709); this reader returns the values.
// Pass sig == NULL to get the header and the number of signatures only.
int32_t kmu_sigdump_read(const char* path, uint32_t* sig_size, uint32_t* sketch_size, uint32_t* kmer_size, uint64_t* nsig,
                         uint32_t* sig, uint64_t first, uint64_t count) {
    if (!path) return fail(KMU_EINVAL, "null argument");
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(KMU_EINVAL, "SigSketchFileReader : could not open dumpfile");
    uint8_t hdr[16];
    if (std::fread(hdr, 1, 16, f) != 16) {
        std::fclose(f);
        return fail(KMU_EINVAL, "SigSketchFileReader could no read magic");
    }
    if (get_u32(hdr) != MAGIC_SIG_DUMP) {
        std::fclose(f);
        return fail(KMU_EINVAL, "file is not a dump of signature");
    }
    const uint32_t ss = get_u32(hdr + 4), sk = get_u32(hdr + 8), ks = get_u32(hdr + 12);
    if (ss != 4) {
        std::fclose(f);
        return fail(KMU_EINVAL, "SigSketchFileReader , sig_size != 4 not yet implemented");
    }
    std::fseek(f, 0, SEEK_END);
    const uint64_t bytes = (uint64_t)std::ftell(f) - 16;
    const uint64_t n = sk ? bytes / (4ull * sk) : 0;
    if (sig_size) *sig_size = ss;
    if (sketch_size) *sketch_size = sk;
    if (kmer_size) *kmer_size = ks;
    if (nsig) *nsig = n;
    int32_t rc = KMU_OK;
    if (sig && count) {
        if (first + count > n) rc = fail(KMU_EINVAL, "signatures %llu..%llu requested, the dump holds %llu",
                                         (unsigned long long)first, (unsigned long long)(first + count), (unsigned long long)n);
        else {
            std::fseek(f, (long)(16 + first * 4ull * sk), SEEK_SET);
            if (std::fread(sig, 4, (size_t)count * sk, f) != (size_t)count * sk) rc = fail(KMU_EINVAL, "short read");
        }
    }
    std::fclose(f);
    return rc;
}

// KmerCountReload::load_multiple_kmers_from_file (src/base/kmercount.rs:1209-1351): header `u32 0xcea2bbff | u8 kmer_size |
// u8 nb_bytes_by_count | u64 nb_kmer`, then records `kmer.dump() | count` until end of file (the reference does not trust
// the declared number either).  kmer.dump() is the 4-byte word for the u32 types (kmer_size <= 16) and `u8 k | u64 value`
// for Kmer64bit (kmer64bit.rs:98-104) -- the reference's reader only knows the 4-byte form.
int32_t kmu_count_reload_multiple(const char* path, uint32_t* kmer_size, uint32_t* count_bytes, uint64_t* nb_declared,
                                  uint64_t* kmers, uint32_t* counts, uint64_t cap, uint64_t* n_read) {
    if (!path) return fail(KMU_EINVAL, "null argument");
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(KMU_EINVAL, "KmerCountReload::load_multiple_kmers_from_file cannot open file %s", path);
    uint8_t hdr[14];
    if (std::fread(hdr, 1, 14, f) != 14) {
        std::fclose(f);
        return fail(KMU_EINVAL, "KmerCountReload::load_multiple_kmers_from_file could no read magic");
    }
    if (get_u32(hdr) != 0xcea2bbffu) {
        std::fclose(f);
        return fail(KMU_EINVAL, "KmerCountReload::load_multiple_kmers_from_file unknow magic %x", get_u32(hdr));
    }
    const uint32_t ksz = hdr[4], cb = hdr[5];
    uint64_t declared;
    std::memcpy(&declared, hdr + 6, 8);
    if (cb != 1 && cb != 2) {
        std::fclose(f);
        return fail(KMU_EINVAL, "load_multiple_kmers, kmer count on more than 2 bytes not yet implemented");
    }
    const size_t ksize = ksz > 16 ? 9 : 4, rec = ksize + cb;
    std::fseek(f, 0, SEEK_END);
    const uint64_t nrec = ((uint64_t)std::ftell(f) - 14) / rec;
    if (kmer_size) *kmer_size = ksz;
    if (count_bytes) *count_bytes = cb;
    if (nb_declared) *nb_declared = declared;
    if (n_read) *n_read = nrec;
    int32_t rc = KMU_OK;
    if (kmers && counts) {
        if (nrec > cap) rc = fail(KMU_EOVERFLOW, "the dump holds %llu k-mers, the buffers %llu", (unsigned long long)nrec, (unsigned long long)cap);
        else {
            std::fseek(f, 14, SEEK_SET);
            std::vector<uint8_t> buf(nrec * rec);
            if (nrec && std::fread(buf.data(), rec, nrec, f) != nrec) rc = fail(KMU_EINVAL, "short read");
            for (uint64_t i = 0; rc == KMU_OK && i < nrec; ++i) {
                const uint8_t* r = buf.data() + i * rec;
                uint64_t v = 0;
                if (ksize == 9) std::memcpy(&v, r + 1, 8);
                else v = get_u32(r);
                kmers[i] = v;
                counts[i] = cb == 1 ? r[ksize] : (uint32_t)r[ksize] | ((uint32_t)r[ksize + 1] << 8);
            }
        }
    }
    std::fclose(f);
    return rc;
}

}  // extern "C"
