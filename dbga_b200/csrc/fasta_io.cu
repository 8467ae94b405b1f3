// FASTA on the wire: the reference loads one pair per file through cogent3
// (src/dbga/utils.py:95-96 load_unaligned_seqs) and writes aln.to_fasta() (src/dbga/cli.py:135-136,
// 60-column blocks).  A batch run wants many pairs per file: the reader parses straight into the
// C ABI's layout -- one pinned ASCII buffer + offsets, and optionally the 2-bit packed form that
// dbga_align_batch_packed uploads -- and the writer produces the wrapped records of every aligned
// pair from a result, expanding a COMPACT_ROWS result on the fly (the columns inside a run of
// anchors are the input's own bytes).  Host code only; all passes are split over host threads.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include <chrono>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "common.cuh"

namespace dbga {
bool host_pack_seq(const uint8_t *src, int64_t len, uint32_t *dst);                                // hostpack.cpp
int64_t host_fasta_body(const uint8_t *p, const uint8_t *end, uint8_t *dst, int64_t room);         // hostpack.cpp
}

namespace {

template <class Fn> void par_for(int64_t n, int threads, int64_t grain, Fn fn) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<int64_t>(threads, std::max<int64_t>(1, n / std::max<int64_t>(grain, 1)));
    if (threads <= 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back([=]() { fn(n * t / threads, n * (t + 1) / threads); });
    for (auto &x : th) x.join();
}

struct FastaImpl {
    dbga_fasta pub{};
    std::vector<int64_t> offsets, name_off;
    std::vector<char> names;
    uint8_t *seqs = nullptr;       // pinned (plain memory when no CUDA device is present: parsing needs none)
    uint32_t *packed = nullptr;
    bool pinned = true;
};

// Pinned buffers are kept for the next file when a batch is freed: page-locking costs ~0.4 ms per MB (150-190 ms
// for the 375 MB of a 100 000-pair file, ten times the parsing itself), and a tool that streams files through the
// aligner would pay it on every one.  A freed buffer is reused by a request of up to its size and at least half of
// it; at most PIN_CACHE_BYTES stay parked (the rest is released).
struct PinCache {
    std::mutex mu;
    std::vector<std::pair<void *, size_t>> free_list;
    size_t parked = 0;
    std::map<void *, size_t> size_of;          // live pinned buffers handed out by host_buf
};
constexpr size_t PIN_CACHE_BYTES = (size_t)4 << 30;
PinCache &pin_cache() { static PinCache *c = new PinCache(); return *c; }     // (never destroyed: no CUDA calls at exit)

void *host_buf(size_t bytes, bool &pinned) {
    void *p = nullptr;
    if (pinned) {
        PinCache &c = pin_cache();
        {
            std::lock_guard<std::mutex> lk(c.mu);
            size_t best = (size_t)-1, bi = 0;
            for (size_t i = 0; i < c.free_list.size(); ++i)
                if (c.free_list[i].second >= bytes && c.free_list[i].second / 2 <= bytes && c.free_list[i].second < best) { best = c.free_list[i].second; bi = i; }
            if (best != (size_t)-1) {
                p = c.free_list[bi].first;
                c.parked -= best;
                c.free_list.erase(c.free_list.begin() + (long)bi);
                c.size_of[p] = best;
                return p;
            }
        }
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) {
            std::lock_guard<std::mutex> lk(c.mu);
            c.size_of[p] = bytes;
            return p;
        }
    }
    cudaGetLastError();
    pinned = false;
    return malloc(bytes);
}
void host_release(void *p, bool) {
    if (!p) return;
    PinCache &c = pin_cache();
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.size_of.find(p);
        if (it != c.size_of.end()) { bytes = it->second; c.size_of.erase(it); }
        if (bytes && c.parked + bytes <= PIN_CACHE_BYTES) { c.free_list.emplace_back(p, bytes); c.parked += bytes; return; }
    }
    if (bytes) cudaFreeHost(p); else free(p);          // (a buffer host_buf did not pin came from malloc)
}

inline bool is_space(uint8_t c) { return c == '\n' || c == '\r' || c == ' ' || c == '\t'; }

}  // namespace

extern "C" {

int dbga_fasta_read(const char *path, int want_packed, int threads, dbga_fasta **out) {
    if (!path || !out) return DBGA_ERR_ARG;
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return DBGA_ERR_ARG;
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); return DBGA_ERR_ARG; }
    const int64_t size = (int64_t)sb.st_size;
    const uint8_t *data = nullptr;
    if (size > 0) {
        data = (const uint8_t *)mmap(nullptr, (size_t)size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (data == MAP_FAILED) { close(fd); return DBGA_ERR_NOMEM; }
        madvise((void *)data, (size_t)size, MADV_SEQUENTIAL);
    }
    close(fd);
    const auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *w) { if (getenv("DBGA_FASTA_TRACE")) fprintf(stderr, "[fasta] %s %.1f ms\n", w, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count()); };
    FastaImpl *f = new FastaImpl();
    // ---- pass 1: record starts ('>' at the start of a line), found in parallel slices
    std::vector<std::vector<int64_t>> found;
    {
        int T = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
        T = (int)std::max<int64_t>(1, std::min<int64_t>(T, size / (1 << 20)));
        found.resize((size_t)T);
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
                const int64_t lo = size * t / T, hi = size * (t + 1) / T;
                const uint8_t *p = data + lo, *end = data + hi;
                while (p < end) {
                    const uint8_t *q = (const uint8_t *)memchr(p, '>', (size_t)(end - p));
                    if (!q) break;
                    if (q == data || q[-1] == '\n') found[(size_t)t].push_back(q - data);
                    p = q + 1;
                }
            });
        for (auto &x : th) x.join();
    }
    lap("pass1 find");
    std::vector<int64_t> starts;
    for (auto &v : found) starts.insert(starts.end(), v.begin(), v.end());
    const int64_t R = (int64_t)starts.size();
    starts.push_back(size);
    // ---- pass 2: per record, where the header ends and how many sequence bytes follow
    std::vector<int64_t> body((size_t)R), nlen((size_t)R + 1, 0), nstart((size_t)R, 0);
    f->offsets.assign((size_t)R + 1, 0);
    par_for(R, threads, 64, [&](int64_t lo, int64_t hi) {
        for (int64_t r = lo; r < hi; ++r) {
            const uint8_t *p = data + starts[(size_t)r] + 1, *end = data + starts[(size_t)r + 1];
            const uint8_t *eol = (const uint8_t *)memchr(p, '\n', (size_t)(end - p));
            if (!eol) eol = end;
            // label = first word of the header line
            const uint8_t *a = p;
            while (a < eol && (*a == ' ' || *a == '\t')) ++a;
            const uint8_t *b = a;
            while (b < eol && !is_space(*b)) ++b;
            nlen[(size_t)r + 1] = b - a;
            nstart[(size_t)r] = a - data;
            body[(size_t)r] = (eol < end ? eol + 1 : end) - data;
            f->offsets[(size_t)r + 1] = dbga::host_fasta_body(data + body[(size_t)r], end, nullptr, 0);      // bases of the record
        }
    });
    lap("pass2 count");
    for (int64_t r = 0; r < R; ++r) { f->offsets[(size_t)r + 1] += f->offsets[(size_t)r]; nlen[(size_t)r + 1] += nlen[(size_t)r]; }
    const int64_t total = f->offsets[(size_t)R];
    f->name_off = nlen;
    f->names.resize((size_t)nlen[(size_t)R] + 1);
    f->seqs = (uint8_t *)host_buf((size_t)std::max<int64_t>(total + 64, 64), f->pinned);
    int64_t words = 0;
    if (want_packed) {
        words = (total >> 4) + R + 1;
        f->packed = (uint32_t *)(f->pinned ? host_buf((size_t)(4 * words + 64), f->pinned) : malloc((size_t)(4 * words + 64)));
    }
    if (!f->seqs || (want_packed && !f->packed)) {
        host_release(f->seqs, f->pinned); host_release(f->packed, f->pinned);
        if (size > 0) munmap((void *)data, (size_t)size);
        delete f;
        return DBGA_ERR_NOMEM;
    }
    lap("alloc");
    // ---- pass 3: labels, upper-cased bases, and (optionally) the canonical 2-bit words
    std::atomic<int64_t> bad{0};
    par_for(R, threads, 64, [&](int64_t lo, int64_t hi) {
        int64_t nbad = 0;
        for (int64_t r = lo; r < hi; ++r) {
            const uint8_t *end = data + starts[(size_t)r + 1];
            memcpy(f->names.data() + nlen[(size_t)r], data + nstart[(size_t)r], (size_t)(nlen[(size_t)r + 1] - nlen[(size_t)r]));
            uint8_t *dst = f->seqs + f->offsets[(size_t)r];
            const int64_t len = f->offsets[(size_t)r + 1] - f->offsets[(size_t)r];
            dbga::host_fasta_body(data + body[(size_t)r], end, dst, len);       // blanks dropped, upper-cased, 32 bytes per step
            if (f->packed) {
                uint32_t *pw = f->packed + (f->offsets[(size_t)r] >> 4) + r;
                nbad += dbga::host_pack_seq(dst, len, pw);          // hostpack.cpp: 32 bases per AVX2 step
            }
        }
        bad += nbad;
    });
    lap("pass3 copy+pack");
    if (size > 0) munmap((void *)data, (size_t)size);
    lap("munmap");
    f->pub.num_records = R;
    f->pub.offsets = f->offsets.data();
    f->pub.seqs = f->seqs;
    f->pub.packed = f->packed;
    f->pub.packed_words = words;
    f->pub.name_off = f->name_off.data();
    f->pub.names = f->names.data();
    f->pub.bad_records = bad.load();
    *out = &f->pub;
    return DBGA_SUCCESS;
}

void dbga_fasta_free(dbga_fasta *pub) {
    if (!pub) return;
    FastaImpl *f = reinterpret_cast<FastaImpl *>(pub);       // pub is the first member
    host_release(f->seqs, f->pinned);
    host_release(f->packed, f->pinned);
    delete f;
}

int64_t dbga_fasta_write(const char *path, const dbga_fasta *in, const dbga_result *res, int64_t first_record,
                         int width, int append, int threads) {
    if (!path || !in || !res || width < 1) return -1;
    const int64_t P = res->num_pairs;
    if (first_record < 0 || first_record + 2 * P > in->num_records) return -1;
    const bool compact = res->seg_cols != nullptr;
    const auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *w) { if (getenv("DBGA_FASTA_TRACE")) fprintf(stderr, "[fasta write] %s %.1f ms\n", w, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count()); };
    if (P > 0 && (!res->status || !res->aln_off || !res->aln0 || !res->aln1)) return -1;
    if (compact && P > 0 && (!res->runs || !res->run_off)) return -1;
    // ---- bytes of every pair's two records
    std::vector<int64_t> pos((size_t)P + 1, 0), full((size_t)P, 0);
    par_for(P, threads, 256, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            if (res->status[i] != DBGA_OK) continue;
            int64_t len = res->aln_off[i + 1] - res->aln_off[i];
            if (compact) {
                const int64_t r0 = res->run_off[i], Rn = res->run_off[i + 1] - r0;
                for (int64_t r = 0; r < Rn; ++r) len += res->runs[3 * (r0 + r) + 2];
                if (Rn > 0) len -= 1;
            }
            full[(size_t)i] = len;
            const int64_t rec = first_record + 2 * i;
            const int64_t lines = (len + width - 1) / width;
            const int64_t n0 = in->name_off[rec + 1] - in->name_off[rec], n1 = in->name_off[rec + 2] - in->name_off[rec + 1];
            // ">label\n", then the row in blocks of `width` columns, every block followed by "\n"
            pos[(size_t)i + 1] = (2 + n0 + len + lines) + (2 + n1 + len + lines);
        }
    });
    lap("sizes");
    int64_t written = 0;
    for (int64_t i = 0; i < P; ++i) { written += res->status[i] == DBGA_OK; pos[(size_t)i + 1] += pos[(size_t)i]; }
    const int64_t total = pos[(size_t)P];
    // the records are laid out in place in a mapping of the output file (no staging copy, no write())
    const int fd = open(path, O_RDWR | O_CREAT | (append ? 0 : O_TRUNC), 0644);
    if (fd < 0) return -1;
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); return -1; }
    const int64_t at = append ? (int64_t)sb.st_size : 0;
    if (total == 0) { close(fd); return written; }
    if (ftruncate(fd, (off_t)(at + total)) != 0) { close(fd); return -1; }
    const int64_t page = sysconf(_SC_PAGESIZE), map_lo = at / page * page;
    uint8_t *map = (uint8_t *)mmap(nullptr, (size_t)(at + total - map_lo), PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)map_lo);
    if (map == MAP_FAILED) { close(fd); return -1; }
    uint8_t *out_base = map + (at - map_lo);
    lap("open + truncate + map");
    par_for(P, threads, 256, [&](int64_t lo, int64_t hi) {
        std::vector<uint8_t> row;
        for (int64_t i = lo; i < hi; ++i) {
            if (res->status[i] != DBGA_OK) continue;
            const int64_t len = full[(size_t)i];
            uint8_t *o = out_base + pos[(size_t)i];
            for (int s = 0; s < 2; ++s) {
                const int64_t rec = first_record + 2 * i + s;
                const int64_t nn = in->name_off[rec + 1] - in->name_off[rec];
                *o++ = '>';
                memcpy(o, in->names + in->name_off[rec], (size_t)nn); o += nn;
                *o++ = '\n';
                const uint8_t *src = (s ? res->aln1 : res->aln0) + res->aln_off[i];
                if (compact) {
                    // rebuild the row: segment columns from the result, run bodies from the input (seq1's bytes
                    // on both rows: inside a run the two sequences are identical, debruijn_pairwise.py:384-388)
                    row.resize((size_t)len);
                    const int64_t r0 = res->run_off[i], Rn = res->run_off[i + 1] - r0;
                    const int32_t *sc = res->seg_cols + r0 + i;
                    const uint8_t *s0 = in->seqs + in->offsets[first_record + 2 * i];
                    uint8_t *w = row.data();
                    for (int64_t t = 0; t <= Rn; ++t) {
                        memcpy(w, src, (size_t)sc[t]); w += sc[t]; src += sc[t];
                        if (t < Rn) {
                            const int32_t a = res->runs[3 * (r0 + t)], l = res->runs[3 * (r0 + t) + 2] - (t == Rn - 1 ? 1 : 0);
                            memcpy(w, s0 + a, (size_t)l); w += l;
                        }
                    }
                    src = row.data();
                }
                for (int64_t c = 0; c < len; c += width) {
                    const int64_t n = std::min<int64_t>(width, len - c);
                    memcpy(o, src + c, (size_t)n); o += n;
                    *o++ = '\n';
                }
            }
        }
    });
    lap("records");
    munmap(map, (size_t)(at + total - map_lo));
    close(fd);
    lap("unmap + close");
    return written;
}

}  // extern "C"
