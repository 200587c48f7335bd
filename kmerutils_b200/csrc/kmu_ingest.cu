// kmu_ingest.cu -- multi-threaded FASTA / FASTQ feeder (host code only).
//
// Replaces the read loop of the reference's sketching binary -- `readblockseq` (src/bin/datasketcher.rs:358-388) feeding
// packs of sequences to the sketcher (:236-300) -- and `parse_with_needletail` (src/io.rs:12-72): records holding any
// non-ACGT character are dropped and counted, the accepted ones keep the order of the file.
//
//   plain files     no serial stage at all: the file is cut into blocks of `block_bytes` by OFFSET, block i owns the records that
//                   start inside it; every parser thread finds the first record start at or after each end of its block (a
//                   line starting with '@' whose record validates -- '+' two lines later, as many quality characters as
//                   bases; a quality line that happens to start with '@' does not -- two threads looking at the same offset
//                   find the same start), preads the range and parses it;
//   gzip files      a reader thread inflates (zlib, serial by nature of the format) and cuts blocks at record boundaries;
//   parser threads  find the records, check the bases through a 256-entry table, copy the bases of the accepted records
//                   back to back into a PINNED pack buffer + offsets;
//   consumer        kmu_ingest_next hands the packs out in file order, ready for kmu_seqbatch_from_ascii (one H2D copy from
//                   pinned memory + the 2-bit pack kernel): the parsers work on the following blocks while the GPU sketches.
//
// The general single-threaded reader (kmu_fastx_*, kmu_io.cu) stays for inputs this one refuses: FASTQ records whose sequence
// spans several lines.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "kmu_host.h"

namespace {

struct Block {  // raw text between two record boundaries
    uint64_t index = 0;
    std::vector<char> text;
};

struct Pack {  // accepted reads of one block
    uint64_t index = 0;
    uint8_t* ascii = nullptr;  // pinned (or malloc'ed without a device), `cap` bytes
    uint64_t cap = 0;
    bool pinned = false;
    std::vector<uint64_t> off;  // nseq + 1
    uint64_t nb_read = 0, nb_bad_read = 0, nb_bases = 0, nb_bad_bases = 0;
    int32_t error = 0;  // KMU_EINVAL: malformed record
    std::string what;
};

struct ValidTable {
    uint8_t ok[256];
    ValidTable() {
        std::memset(ok, 0, sizeof(ok));
        for (const char* p = "ACGTacgt"; *p; ++p) ok[(uint8_t)*p] = 1;  // Alphabet2b::is_valid_base (alphabet.rs:157-159)
    }
};
const ValidTable VALID;

// bases that are not ACGT, case-insensitive (Alphabet2b::is_valid_base, alphabet.rs:157-159); branch-free so that the
// compiler vectorises it
inline uint64_t count_invalid(const uint8_t* s, size_t len) {
    uint64_t bad = 0;
    size_t i = 0;
    for (; i + 64 <= len; i += 64) {
        uint8_t acc = 0;
        for (size_t j = 0; j < 64; ++j) {
            const uint8_t c = s[i + j] & 0xDF;
            acc += (uint8_t)((c != 'A') & (c != 'C') & (c != 'G') & (c != 'T'));
        }
        bad += acc;
    }
    for (; i < len; ++i) bad += !VALID.ok[s[i]];
    return bad;
}

inline const char* line_end(const char* p, const char* end) {
    const char* q = (const char*)std::memchr(p, '\n', (size_t)(end - p));
    return q ? q : end;
}

// does a 4-line FASTQ record start at p?  (header '@', sequence, '+', as many quality characters; then '@' or the end)
// 1 yes, 0 no, -1 the buffer ends before one can tell (at_eof: the buffer ends with the file, a cut-off record counts)
int fastq_record_state(const char* p, const char* end, bool at_eof) {
    if (p >= end) return at_eof ? 0 : -1;
    if (*p != '@') return 0;
    const char* e1 = line_end(p, end);
    if (e1 >= end) return at_eof ? 0 : -1;
    const char* s = e1 + 1;
    const char* e2 = line_end(s, end);
    if (e2 >= end) return at_eof ? 0 : -1;
    const char* plus = e2 + 1;
    if (plus >= end) return at_eof ? 0 : -1;
    if (*plus != '+') return 0;
    const char* e3 = line_end(plus, end);
    if (e3 >= end) return at_eof ? 0 : -1;
    const char* q = e3 + 1;
    const char* e4 = line_end(q, end);
    size_t ls = (size_t)(e2 - s), lq = (size_t)(e4 - q);
    if (ls && s[ls - 1] == '\r') --ls;
    if (e4 >= end) {
        if (!at_eof) return -1;
        return lq == ls ? 1 : 0;  // last record of the file without a final newline
    }
    if (lq && q[lq - 1] == '\r') --lq;
    if (ls != lq) return 0;
    if (e4 + 1 >= end) return at_eof ? 1 : -1;
    return (e4[1] == '@' || e4[1] == '\n' || e4[1] == '\r') ? 1 : 0;
}
bool fastq_record_at(const char* p, const char* end) { return fastq_record_state(p, end, true) == 1; }

}  // namespace

struct kmu_ingest {
    gzFile f = nullptr;
    int fd = -1;                 // plain files: pread by the parser threads, no reader thread
    uint64_t file_size = 0;
    uint64_t next_block = 0;     // plain files: next block index to hand out
    bool fastq = false;
    uint64_t block_bytes = 0;
    bool pinned = false;
    std::thread reader;
    std::vector<std::thread> parsers;
    std::mutex mu;
    std::condition_variable cv_blocks, cv_packs, cv_free;
    std::deque<Block*> blocks;          // reader -> parsers
    std::map<uint64_t, Pack*> ready;    // parsers -> consumer, keyed by block index
    std::vector<Pack*> free_packs;      // pack buffers not in use
    std::vector<Pack*> all_packs;
    uint64_t next_out = 0;              // block index the consumer waits for
    uint64_t nblocks = 0;               // blocks produced by the reader so far
    bool reader_done = false, stop = false;
    int32_t reader_error = 0;
    std::string reader_what;
    uint64_t nb_read = 0, nb_bad_read = 0, nb_bases = 0, nb_bad_bases = 0;
    size_t max_queued = 0;

    void reader_main();
    void parser_main();
    void parser_main_plain();
    uint64_t record_start_at_or_after(uint64_t off, std::vector<char>& win) const;
    void parse_block(const Block& b, Pack& p) const;
    bool grow_pack(Pack& p, uint64_t need) const;
};

// a block larger than the pack buffers (one record longer than a block): a larger buffer for this pack
bool kmu_ingest::grow_pack(Pack& p, uint64_t need) const {
    if (need <= p.cap) return true;
    void* mem = nullptr;
    bool pin = pinned;
    if (pin && cudaHostAlloc(&mem, need, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        pin = false;
    }
    if (!pin) mem = std::malloc(need);
    if (!mem) return false;
    if (p.pinned) cudaFreeHost(p.ascii);
    else std::free(p.ascii);
    p.ascii = (uint8_t*)mem;
    p.cap = need;
    p.pinned = pin;
    return true;
}

void kmu_ingest::reader_main() {
    std::vector<char> carry;  // bytes after the last record boundary of the previous read
    bool eof = false;
    while (!eof) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_blocks.wait(lk, [&] { return stop || blocks.size() < max_queued; });
            if (stop) break;
        }
        Block* b = new Block();
        b->text.reserve(carry.size() + block_bytes + 1);
        b->text.assign(carry.begin(), carry.end());
        carry.clear();
        const size_t have = b->text.size();
        b->text.resize(have + block_bytes);
        size_t got_total = 0;
        while (got_total < block_bytes) {
            const int got = gzread(f, b->text.data() + have + got_total, (unsigned)std::min<uint64_t>(block_bytes - got_total, 1u << 30));
            if (got < 0) {
                std::lock_guard<std::mutex> lk(mu);
                reader_error = KMU_EINVAL;
                reader_what = "corrupt or truncated compressed stream";
                eof = true;
                break;
            }
            if (got == 0) {
                eof = true;
                break;
            }
            got_total += (size_t)got;
        }
        b->text.resize(have + got_total);
        if (!eof) {
            // cut at the last record boundary; what follows starts the next block
            const char* base = b->text.data();
            const char* end = base + b->text.size();
            const char* cutp = nullptr;
            const char* p = end;
            size_t looked = 0;
            while (p > base && looked < (64u << 20)) {
                const char* nl = (const char*)memrchr(base, '\n', (size_t)(p - base));
                if (!nl) break;
                looked += (size_t)(p - nl);
                const char* cand = nl + 1;
                if (cand < end && ((fastq && fastq_record_at(cand, end)) || (!fastq && *cand == '>'))) {
                    cutp = cand;
                    break;
                }
                p = nl;
            }
            if (!cutp) {
                if (b->text.size() > (1ull << 31)) {
                    std::lock_guard<std::mutex> lk(mu);
                    reader_error = KMU_EINVAL;
                    reader_what = fastq ? "no 4-line FASTQ record boundary found (multi-line FASTQ? use kmu_fastx_open)" : "no FASTA record boundary found";
                    delete b;
                    break;
                }
                carry.swap(b->text);  // one record larger than a block: keep reading
                delete b;
                continue;
            }
            carry.assign(cutp, end);
            b->text.resize((size_t)(cutp - base));
        }
        if (b->text.empty()) {
            delete b;
            continue;
        }
        std::lock_guard<std::mutex> lk(mu);
        b->index = nblocks++;
        blocks.push_back(b);
        cv_blocks.notify_all();
    }
    std::lock_guard<std::mutex> lk(mu);
    reader_done = true;
    cv_blocks.notify_all();
    cv_packs.notify_all();
}

void kmu_ingest::parse_block(const Block& b, Pack& p) const {
    const char* cur = b.text.data();
    const char* end = cur + b.text.size();
    uint64_t used = 0;
    p.off.clear();
    p.off.push_back(0);
    auto bad_record = [&](const char* why) {
        p.error = KMU_EINVAL;
        p.what = why;
    };
    if (!grow_pack(p, b.text.size() + 64)) {
        p.error = KMU_ENOMEM;
        p.what = "pack buffer of the reader";
        return;
    }
    while (cur < end) {
        while (cur < end && (*cur == '\n' || *cur == '\r')) ++cur;  // blank lines between records
        if (cur >= end) break;
        if (*cur != (fastq ? '@' : '>')) return bad_record(fastq ? "invalid record: expected '@'" : "invalid record: expected '>'");
        cur = line_end(cur, end);  // header
        if (cur < end) ++cur;
        const uint64_t start = used;
        uint64_t bad = 0, n = 0;
        auto take_line = [&](const char* s, const char* e) {
            if (e > s && e[-1] == '\r') --e;
            const size_t len = (size_t)(e - s);
            if (used + len > p.cap) return false;
            std::memcpy(p.ascii + used, s, len);
            bad += count_invalid((const uint8_t*)s, len);
            used += len;
            n += len;
            return true;
        };
        if (fastq) {
            const char* e2 = line_end(cur, end);
            if (!take_line(cur, e2)) return bad_record("pack buffer too small");
            cur = e2 < end ? e2 + 1 : end;
            if (cur >= end || *cur != '+') return bad_record("invalid record: truncated FASTQ or a sequence on several lines (use kmu_fastx_open)");
            cur = line_end(cur, end);
            if (cur < end) ++cur;
            const char* e4 = line_end(cur, end);  // qualities
            size_t lq = (size_t)(e4 - cur);
            if (lq && cur[lq - 1] == '\r') --lq;
            if (lq != n) return bad_record("invalid record: truncated FASTQ qualities");
            cur = e4 < end ? e4 + 1 : end;
        } else {
            while (cur < end && *cur != '>') {  // sequence lines up to the next header
                const char* e = line_end(cur, end);
                if (!take_line(cur, e)) return bad_record("pack buffer too small");
                cur = e < end ? e + 1 : end;
            }
        }
        p.nb_read++;
        p.nb_bases += n;
        p.nb_bad_bases += bad;
        if (bad) {  // io.rs:41-48: the whole read is dropped
            p.nb_bad_read++;
            used = start;
        } else {
            p.off.push_back(used);
        }
    }
}

void kmu_ingest::parser_main() {
    for (;;) {
        Block* b = nullptr;
        Pack* p = nullptr;
        {
            std::unique_lock<std::mutex> lk(mu);
            // a pack buffer is taken only together with a block, and the block the consumer waits for is never starved:
            // blocks come out of the queue in index order and every parser that holds one holds a buffer
            cv_blocks.wait(lk, [&] { return stop || (!blocks.empty() && !free_packs.empty()) || (reader_done && blocks.empty()); });
            if (stop || (blocks.empty() && reader_done)) return;
            b = blocks.front();
            blocks.pop_front();
            p = free_packs.back();
            free_packs.pop_back();
            cv_blocks.notify_all();  // the reader may queue another block
        }
        p->index = b->index;
        p->nb_read = p->nb_bad_read = p->nb_bases = p->nb_bad_bases = 0;
        p->error = 0;
        p->what.clear();
        parse_block(*b, *p);
        delete b;
        std::lock_guard<std::mutex> lk(mu);
        ready[p->index] = p;
        cv_packs.notify_all();
    }
}

// plain files: first record start at or after byte `off` (file_size when there is none); deterministic in the file content, so
// the two threads that look at a block border agree
uint64_t kmu_ingest::record_start_at_or_after(uint64_t off, std::vector<char>& win) const {
    if (off == 0) return 0;
    if (off >= file_size) return file_size;
    uint64_t span = 1u << 16;
    for (;;) {
        const uint64_t from = off - 1;  // the byte before decides whether `off` itself starts a line
        const uint64_t len = std::min<uint64_t>(span, file_size - from);
        win.resize(len);
        uint64_t got = 0;
        while (got < len) {
            const ssize_t r = pread(fd, win.data() + got, len - got, (off_t)(from + got));
            if (r <= 0) break;
            got += (uint64_t)r;
        }
        const bool at_eof = from + got >= file_size;
        const char* base = win.data();
        const char* end = base + got;
        const char* p = base;
        bool need_more = false;
        while (p < end) {
            const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
            if (!nl) break;
            const char* cand = nl + 1;
            if (cand >= end) {
                need_more = !at_eof;
                break;
            }
            if (fastq) {
                const int st = fastq_record_state(cand, end, at_eof);
                if (st == 1) return from + (uint64_t)(cand - base);
                if (st < 0) {
                    need_more = true;
                    break;
                }
            } else if (*cand == '>') {
                return from + (uint64_t)(cand - base);
            }
            p = cand;
        }
        if (at_eof && !need_more) return file_size;
        if (!need_more && got == len && !at_eof) need_more = true;  // no line start in the window
        if (!need_more) return file_size;
        if (span >= (1ull << 33)) return file_size;
        span *= 4;
    }
}

void kmu_ingest::parser_main_plain() {
    std::vector<char> win;
    Block blk;
    for (;;) {
        Pack* p = nullptr;
        uint64_t idx = 0;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_blocks.wait(lk, [&] { return stop || next_block >= nblocks || !free_packs.empty(); });
            if (stop || next_block >= nblocks) return;
            idx = next_block++;
            p = free_packs.back();
            free_packs.pop_back();
        }
        const uint64_t b0 = record_start_at_or_after(idx * block_bytes, win);
        const uint64_t b1 = record_start_at_or_after(std::min(file_size, (idx + 1) * block_bytes), win);
        p->index = idx;
        p->nb_read = p->nb_bad_read = p->nb_bases = p->nb_bad_bases = 0;
        p->error = 0;
        p->what.clear();
        p->off.assign(1, 0);
        if (b1 > b0) {
            blk.text.resize(b1 - b0);
            uint64_t got = 0;
            while (got < b1 - b0) {
                const ssize_t r = pread(fd, blk.text.data() + got, b1 - b0 - got, (off_t)(b0 + got));
                if (r <= 0) break;
                got += (uint64_t)r;
            }
            if (got != b1 - b0) {
                p->error = KMU_EINVAL;
                p->what = "short read from the input file";
            } else {
                parse_block(blk, *p);
            }
        }
        std::lock_guard<std::mutex> lk(mu);
        ready[p->index] = p;
        cv_packs.notify_all();
    }
}

extern "C" {

int32_t kmu_ingest_open(const char* path, uint32_t nthreads, uint64_t block_bytes, kmu_ingest** out) {
    if (!path || !out) return fail(KMU_EINVAL, "null argument");
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(KMU_EINVAL, "file does not exist: %s", path);
    unsigned char magic[2] = {0, 0};
    const ssize_t nm = pread(fd, magic, 2, 0);
    const bool gz = nm == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    struct stat sb;
    fstat(fd, &sb);
    auto* h = new kmu_ingest();
    int c = -1;
    if (gz) {
        close(fd);
        gzFile f = gzopen(path, "rb");
        if (!f) {
            delete h;
            return fail(KMU_EINVAL, "file does not exist: %s", path);
        }
        gzbuffer(f, 1u << 22);
        c = gzgetc(f);
        while (c == '\n' || c == '\r') c = gzgetc(f);
        if (c != -1) gzungetc(c, f);
        h->f = f;
    } else {
        h->fd = fd;
        h->file_size = (uint64_t)sb.st_size;
        char first[4096];
        const ssize_t nf = pread(fd, first, sizeof(first), 0);
        for (ssize_t i = 0; i < nf; ++i)
            if (first[i] != '\n' && first[i] != '\r') {
                c = (unsigned char)first[i];
                break;
            }
    }
    if (c != '>' && c != '@' && c != -1) {
        kmu_ingest_close(h);
        return fail(KMU_EINVAL, "invalid record: expected '>' or '@', got '%c'", c);
    }
    h->fastq = c == '@';
    if (nthreads == 0) nthreads = std::max(1u, std::thread::hardware_concurrency());
    nthreads = std::min(nthreads, 64u);
    h->block_bytes = block_bytes ? std::max<uint64_t>(block_bytes, 1u << 16) : (32ull << 20);
    h->max_queued = nthreads + 2;
    // pack buffers: a block holds at most its own size in bases (+ what one oversized record may add: grown on demand)
    int ndev = 0;
    h->pinned = cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0;
    const size_t npacks = nthreads + 3;
    for (size_t i = 0; i < npacks; ++i) {
        Pack* p = new Pack();
        p->cap = h->block_bytes + (h->block_bytes >> 2) + 4096;
        void* mem = nullptr;
        p->pinned = h->pinned;
        if (p->pinned && cudaHostAlloc(&mem, p->cap, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            p->pinned = false;
        }
        if (!p->pinned) mem = std::malloc(p->cap);
        if (!mem) {
            delete p;
            kmu_ingest_close(h);
            return fail(KMU_ENOMEM, "pack buffers of the reader");
        }
        p->ascii = (uint8_t*)mem;
        h->all_packs.push_back(p);
        h->free_packs.push_back(p);
    }
    if (h->fd >= 0) {  // plain file: the blocks are known in advance, the parser threads read them themselves
        h->nblocks = (h->file_size + h->block_bytes - 1) / h->block_bytes;
        h->reader_done = true;
        for (uint32_t t = 0; t < nthreads; ++t) h->parsers.emplace_back(&kmu_ingest::parser_main_plain, h);
    } else {
        h->reader = std::thread(&kmu_ingest::reader_main, h);
        for (uint32_t t = 0; t < nthreads; ++t) h->parsers.emplace_back(&kmu_ingest::parser_main, h);
    }
    *out = h;
    return KMU_OK;
}

// The next pack in file order: *ascii (pinned host memory) holds the bases of its *nseq accepted reads back to back,
// (*ascii_off)[0..nseq] delimits them; valid until kmu_ingest_release(token).  End of file: *nseq = 0, *ascii = NULL.
int32_t kmu_ingest_next(kmu_ingest* h, const uint8_t** ascii, const uint64_t** ascii_off, uint64_t* nseq, void** token) {
    if (!h || !ascii || !ascii_off || !nseq || !token) return fail(KMU_EINVAL, "null argument");
    *ascii = nullptr;
    *ascii_off = nullptr;
    *nseq = 0;
    *token = nullptr;
    for (;;) {
        Pack* p = nullptr;
        {
            std::unique_lock<std::mutex> lk(h->mu);
            h->cv_packs.wait(lk, [&] { return h->ready.count(h->next_out) || (h->reader_done && h->next_out >= h->nblocks); });
            if (!h->ready.count(h->next_out)) {
                if (h->reader_error) return fail(h->reader_error, "%s", h->reader_what.c_str());
                return KMU_OK;  // end of file
            }
            p = h->ready[h->next_out];
            h->ready.erase(h->next_out);
            ++h->next_out;
            h->nb_read += p->nb_read;
            h->nb_bad_read += p->nb_bad_read;
            h->nb_bases += p->nb_bases;
            h->nb_bad_bases += p->nb_bad_bases;
        }
        if (p->error) {
            const int32_t rc = fail(p->error, "%s", p->what.c_str());
            std::lock_guard<std::mutex> lk(h->mu);
            h->free_packs.push_back(p);
            h->cv_blocks.notify_all();
            return rc;
        }
        if (p->off.size() <= 1) {  // every read of the block was dropped: next one
            std::lock_guard<std::mutex> lk(h->mu);
            h->free_packs.push_back(p);
            h->cv_blocks.notify_all();
            continue;
        }
        *ascii = p->ascii;
        *ascii_off = p->off.data();
        *nseq = p->off.size() - 1;
        *token = p;
        return KMU_OK;
    }
}

int32_t kmu_ingest_release(kmu_ingest* h, void* token) {
    if (!h || !token) return fail(KMU_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    h->free_packs.push_back((Pack*)token);
    h->cv_blocks.notify_all();
    return KMU_OK;
}

void kmu_ingest_stats(const kmu_ingest* h, uint64_t* nb_read, uint64_t* nb_bad_read, uint64_t* nb_bases, uint64_t* nb_bad_bases) {
    if (!h) return;
    if (nb_read) *nb_read = h->nb_read;
    if (nb_bad_read) *nb_bad_read = h->nb_bad_read;
    if (nb_bases) *nb_bases = h->nb_bases;
    if (nb_bad_bases) *nb_bad_bases = h->nb_bad_bases;
}

void kmu_ingest_close(kmu_ingest* h) {
    if (!h) return;
    {
        std::lock_guard<std::mutex> lk(h->mu);
        h->stop = true;
        h->cv_blocks.notify_all();
        h->cv_packs.notify_all();
    }
    if (h->reader.joinable()) h->reader.join();
    for (auto& t : h->parsers)
        if (t.joinable()) t.join();
    for (Block* b : h->blocks) delete b;
    for (Pack* p : h->all_packs) {
        if (p->ascii) {
            if (p->pinned) cudaFreeHost(p->ascii);
            else std::free(p->ascii);
        }
        delete p;
    }
    if (h->f) gzclose(h->f);
    if (h->fd >= 0) close(h->fd);
    delete h;
}

}  // extern "C"
