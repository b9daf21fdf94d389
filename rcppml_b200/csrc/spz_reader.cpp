// spz_reader.cpp — see spz_reader.hpp. Host code only (compiled by g++ with -ffp-contract=off: the QUANT8 value map is
// `offset + scale * q` with separately rounded multiply and add, as the reference's -O2 build computes it,
// sparsepress_v2.hpp:1064-1068).
#include "spz_reader.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

namespace b200 {
namespace spz {

namespace {

constexpr uint32_t kHeaderBytes = 128;       // header_v2.hpp:38
constexpr uint32_t kDescBytes = 48;          // header_v2.hpp:236
constexpr uint32_t kFooterBytes = 16;        // header_v2.hpp:41
constexpr int      kProbBits = 14;           // codec/rans.hpp:24
constexpr uint32_t kProbScale = 1u << kProbBits;
constexpr uint32_t kRansLow = 1u << 23;      // codec/rans.hpp:186
constexpr uint32_t kEscape = 255;            // sparsepress_v2.hpp:113-114

[[noreturn]] void corrupt(const char* what) { throw Error(kCorrupt, std::string("spz: ") + what); }

template <typename T>
inline T load(const uint8_t* p) { T v; std::memcpy(&v, p, sizeof(T)); return v; }

// Bounds-checked forward cursor over one stream.
struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    size_t left() const { return static_cast<size_t>(end - p); }
    void need(size_t n, const char* what) const { if (left() < n) corrupt(what); }
    uint32_t u32(const char* what) { need(4, what); uint32_t v = load<uint32_t>(p); p += 4; return v; }
    uint16_t u16(const char* what) { need(2, what); uint16_t v = load<uint16_t>(p); p += 2; return v; }
    uint8_t  u8(const char* what)  { need(1, what); return *p++; }
    // 7-bit continuation code, least significant group first (codec/varint.hpp:52-61)
    uint64_t varint(const char* what) {
        uint64_t v = 0; int shift = 0;
        for (;;) {
            if (p == end || shift > 63) corrupt(what);
            const uint8_t b = *p++;
            v |= static_cast<uint64_t>(b & 0x7F) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
        }
    }
};

// Decoding model of one rANS stream: slot (state mod 2^14) -> symbol, symbol -> (freq | cum << 16).
// Wire form (codec/rans.hpp:142-171): u16 n_symbols, then n_symbols u16 frequencies summing to 2^14.
struct Model {
    // slot -> symbol as BYTES when the alphabet has at most 256 symbols (every stream the writer produces: MAX_SYM = 255,
    // sparsepress_v2.hpp:113): 16 KB, so that the uniformly distributed slot look-ups of a few streams stay in L1;
    // a wider alphabet (legal on the wire, u16 n_symbols) falls back to 16-bit entries.
    uint8_t slot8[kProbScale];
    std::vector<uint16_t> slot16;
    std::vector<uint32_t> fc;
    bool wide = false;
    void parse(Cursor& c) {
        const uint32_t ns = c.u16("rANS table truncated");
        c.need(2u * ns, "rANS table truncated");
        fc.resize(ns);
        wide = ns > 256;
        if (wide) slot16.assign(kProbScale, 0);
        uint32_t cum = 0;
        for (uint32_t s = 0; s < ns; ++s) {
            const uint32_t f = load<uint16_t>(c.p + 2u * s);
            if (cum + f > kProbScale) corrupt("rANS frequencies exceed 2^14");
            fc[s] = f | (cum << 16);
            if (wide) for (uint32_t k = 0; k < f; ++k) slot16[cum + k] = static_cast<uint16_t>(s);
            else std::memset(slot8 + cum, static_cast<int>(s), f);
            cum += f;
        }
        c.p += 2u * ns;
        if (cum != kProbScale) corrupt("rANS frequencies do not sum to 2^14");
    }
};

// N byte-renormalised rANS streams decoded in lock-step (codec/rans.hpp:204-248 for one): the initial state is
// the first four bytes, most significant first; a symbol is the slot's owner; the state shrinks to
// freq * (x >> 14) + slot - cum and refills bytewise below 2^23 while input remains.
template <int N, bool WIDE, typename Emit>
void rans_decode_impl(const Model* const (&mdl)[N], const uint8_t* const (&src)[N], const size_t (&len)[N], uint64_t count,
                      Emit emit) {
    uint32_t x[N];
    const uint8_t* p[N];
    const uint8_t* e[N];
    const uint8_t* s8[N];
    const uint16_t* s16[N];
    const uint32_t* fcs[N];
    for (int s = 0; s < N; ++s) {
        if (len[s] < 4) corrupt("rANS stream shorter than its initial state");
        p[s] = src[s]; e[s] = src[s] + len[s];
        x[s] = (static_cast<uint32_t>(p[s][0]) << 24) | (static_cast<uint32_t>(p[s][1]) << 16) |
               (static_cast<uint32_t>(p[s][2]) << 8) | p[s][3];
        p[s] += 4;
        s8[s] = mdl[s]->slot8; s16[s] = mdl[s]->slot16.data(); fcs[s] = mdl[s]->fc.data();
    }
    for (uint64_t k = 0; k < count; ++k) {
        uint32_t sym[N];
#pragma GCC unroll 9
        for (int s = 0; s < N; ++s) {
            const uint32_t slot = x[s] & (kProbScale - 1);
            const uint32_t y = WIDE ? (mdl[s]->wide ? s16[s][slot] : s8[s][slot]) : s8[s][slot];
            const uint32_t fc = fcs[s][y];
            uint32_t nx = (fc & 0xFFFFu) * (x[s] >> kProbBits) + slot - (fc >> 16);
            const uint8_t* q = p[s];
            if (__builtin_expect(e[s] - q >= 2 && nx >= (1u << 7), 1)) {
                // A state of a sound stream is >= 2^9 here (x >= 2^23, freq >= 1), so the refill is 0, 1 or 2 bytes:
                // taken without a data-dependent branch (the bytewise loop mispredicts on every other symbol).
                const uint32_t take = static_cast<uint32_t>(nx < kRansLow) + static_cast<uint32_t>(nx < (1u << 15));
                const uint32_t two = (static_cast<uint32_t>(q[0]) << 8) | q[1];
                nx = (nx << (8 * take)) | (two >> (16 - 8 * take));
                q += take;
            } else {
                while (nx < kRansLow && q < e[s]) nx = (nx << 8) | *q++;
            }
            p[s] = q; x[s] = nx; sym[s] = y;
        }
        emit(k, sym);
    }
}

template <int N, typename Emit>
void rans_decode(const Model* const (&mdl)[N], const uint8_t* const (&src)[N], const size_t (&len)[N], uint64_t count,
                 Emit emit) {
    bool wide = false;
    for (int s = 0; s < N; ++s) wide = wide || mdl[s]->wide;
    if (wide) rans_decode_impl<N, true>(mdl, src, len, count, emit);
    else rans_decode_impl<N, false>(mdl, src, len, count, emit);
}

inline float half_to_float(uint16_t h) {
    const uint32_t sign = static_cast<uint32_t>(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu, frac = h & 0x3FFu, bits;
    if (exp == 0) {
        if (frac == 0) bits = sign;
        else {   // subnormal half: normalise
            int sh = 0;
            while (!(frac & 0x400u)) { frac <<= 1; ++sh; }
            bits = sign | ((127u - 15u + 1u - sh) << 23) | ((frac & 0x3FFu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (frac << 13);
    else bits = sign | ((exp + 127u - 15u) << 23) | (frac << 13);
    float f; std::memcpy(&f, &bits, 4); return f;
}

// One rANS stream ready to decode.
struct Plane {
    std::unique_ptr<Model> mdl;
    const uint8_t* src = nullptr;
    size_t len = 0;
};

// [model][u32 enc bytes][enc][u32 overflow bytes][overflow varints] (sparsepress_v2.hpp:404-438): the header of an
// "escaped" stream — a symbol 255 stands for the next overflow varint when the stream carries an overflow section.
// The caller decodes the plane and substitutes the varints in order (a partial read has to walk the ones it skips).
struct EscapedStream {
    Plane plane;
    Cursor overflow{nullptr, nullptr};
    bool has_overflow = false;
};

EscapedStream open_escaped(const uint8_t* data, size_t size) {
    EscapedStream r;
    Cursor c{data, data + size};
    r.plane.mdl = std::make_unique<Model>();
    r.plane.mdl->parse(c);
    const uint32_t enc = c.u32("stream truncated at its encoded size");
    c.need(enc, "stream truncated inside its rANS payload");
    r.plane.src = c.p; r.plane.len = enc;
    c.p += enc;
    if (c.left() >= 4) {
        const uint32_t ov = c.u32("");
        c.need(ov, "overflow section exceeds its stream");
        r.overflow = Cursor{c.p, c.p + ov};
        r.has_overflow = ov > 0;
    }
    return r;
}

// Lock-step decode of planes[0 .. N): emit(k, sym[N]).
template <int N, typename Emit>
void decode_planes(const Plane* const (&pl)[N], uint64_t count, Emit emit) {
    const Model* mdl[N]; const uint8_t* src[N]; size_t len[N];
    for (int s = 0; s < N; ++s) { mdl[s] = pl[s]->mdl.get(); src[s] = pl[s]->src; len[s] = pl[s]->len; }
    const Model* const (&m)[N] = mdl; const uint8_t* const (&q)[N] = src; const size_t (&l)[N] = len;
    rans_decode<N>(m, q, l, count, emit);
}

void parallel_for(size_t n, int threads, const std::function<void(size_t)>& fn) {
    if (n == 0) return;
    unsigned t = threads > 0 ? static_cast<unsigned>(threads) : std::max(1u, std::thread::hardware_concurrency());
    t = static_cast<unsigned>(std::min<size_t>(t, n));
    std::atomic<size_t> next{0};
    std::exception_ptr first;
    std::mutex mu;
    auto body = [&] {
        for (;;) {
            const size_t k = next.fetch_add(1, std::memory_order_relaxed);
            if (k >= n) return;
            try { fn(k); }
            catch (...) {
                std::lock_guard<std::mutex> g(mu);
                if (!first) first = std::current_exception();
                next.store(n, std::memory_order_relaxed);
                return;
            }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned k = 1; k < t; ++k) {
        try { pool.emplace_back(body); } catch (...) { break; }   // fewer threads, same result
    }
    body();
    for (auto& th : pool) th.join();
    if (first) std::rethrow_exception(first);
}

uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[8][256];
    static std::once_flag once;
    std::call_once(once, [] {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int j = 0; j < 8; ++j) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int t = 1; t < 8; ++t) table[t][i] = table[0][table[t - 1][i] & 0xFF] ^ (table[t - 1][i] >> 8);
    });
    while (n >= 8) {   // slicing-by-8
        const uint32_t a = load<uint32_t>(p) ^ crc, b = load<uint32_t>(p + 4);
        crc = table[7][a & 0xFF] ^ table[6][(a >> 8) & 0xFF] ^ table[5][(a >> 16) & 0xFF] ^ table[4][a >> 24] ^
              table[3][b & 0xFF] ^ table[2][(b >> 8) & 0xFF] ^ table[1][(b >> 16) & 0xFF] ^ table[0][b >> 24];
        p += 8; n -= 8;
    }
    while (n--) crc = table[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
    return crc;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------

File::File(const char* path) {
    fd_ = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd_ < 0) throw Error(kCannotOpen, std::string("spz: cannot open ") + path);
    struct stat st;
    if (::fstat(fd_, &st) != 0 || !S_ISREG(st.st_mode)) { ::close(fd_); fd_ = -1; throw Error(kReadFailed, "spz: cannot stat the file"); }
    size_ = static_cast<size_t>(st.st_size);
    if (size_ < 6) { ::close(fd_); fd_ = -1; throw Error(kTooSmall, "spz: file too small"); }
    void* map = ::mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (map == MAP_FAILED) { ::close(fd_); fd_ = -1; throw Error(kReadFailed, "spz: cannot map the file"); }
    base_ = static_cast<const uint8_t*>(map);
    ::madvise(map, size_, MADV_WILLNEED);
    try { parse(); }
    catch (...) {
        ::munmap(map, size_); ::close(fd_); fd_ = -1; base_ = nullptr;
        throw;
    }
}

File::~File() {
    if (base_) ::munmap(const_cast<uint8_t*>(base_), size_);
    if (fd_ >= 0) ::close(fd_);
}

void File::parse() {
    if (std::memcmp(base_, "SPRZ", 4) != 0) corrupt("bad magic bytes (not a StreamPress file)");
    const uint16_t version = load<uint16_t>(base_ + 4);
    if (version != 2) throw Error(kNotV2, "spz: only the v2 sparse container is read here (got v" + std::to_string(version) + ")");
    if (size_ < kHeaderBytes + kFooterBytes) corrupt("file smaller than header + footer");
    const uint8_t* h = base_;
    info_.m = load<uint32_t>(h + 8);
    info_.n = load<uint32_t>(h + 12);
    info_.nnz = load<uint64_t>(h + 16);
    info_.chunk_cols = load<uint32_t>(h + 24);
    info_.num_chunks = load<uint32_t>(h + 28);
    info_.value_type = h[40];
    info_.row_sorted = h[42];
    const uint64_t index_offset = load<uint64_t>(h + 48);
    info_.data_offset = load<uint64_t>(h + 64);
    info_.transpose_offset = load<uint64_t>(h + 72);
    info_.metadata_offset = load<uint64_t>(h + 80);
    info_.density = load<float>(h + 92);
    info_.obs_table_offset = load<uint64_t>(h + 96);
    info_.var_table_offset = load<uint64_t>(h + 104);
    info_.transp_chunk_cols = load<uint32_t>(h + 112);
    info_.file_bytes = size_;
    if (info_.value_type > kF64) corrupt("unknown value type");
    if (info_.m > 0x7FFFFFFFu || info_.n > 0x7FFFFFFFu || info_.nnz > 0x7FFFFFFFull)
        throw Error(kBadArgument, "spz: dimensions or nnz beyond the engine's 32-bit indices");

    const size_t body_end = size_ - kFooterBytes;
    auto read_descs = [&](uint64_t off, uint32_t count, uint64_t data_base, uint64_t data_end, uint32_t ncols,
                          std::vector<Chunk>& out, uint64_t& total_nnz) {
        if (off > body_end || static_cast<uint64_t>(count) * kDescBytes > body_end - off) corrupt("chunk index truncated");
        out.resize(count);
        uint64_t before = 0; uint32_t next_col = 0;
        for (uint32_t c = 0; c < count; ++c) {
            const uint8_t* d = base_ + off + static_cast<uint64_t>(c) * kDescBytes;
            Chunk& ck = out[c];
            ck.col_start = load<uint32_t>(d); ck.num_cols = load<uint32_t>(d + 4); ck.nnz = load<uint32_t>(d + 8);
            const uint64_t go = load<uint32_t>(d + 12), vo = load<uint32_t>(d + 16);
            ck.gap_bytes = load<uint32_t>(d + 20); ck.value_bytes = load<uint32_t>(d + 24);
            ck.quant_scale = load<float>(d + 36); ck.quant_offset = load<float>(d + 40);
            if (ck.col_start != next_col || ck.num_cols > ncols - next_col) corrupt("chunk index does not tile the columns");
            next_col += ck.num_cols;
            if (data_base + go + ck.gap_bytes > data_end || data_base + vo + ck.value_bytes > data_end)
                corrupt("chunk stream outside the file");
            // every column owns at least one varint byte of its chunk's gap stream (sparsepress_v2.hpp:87-94), so the
            // column count of a section is bounded by the file size — and with it every allocation sized from the header
            if (ck.gap_bytes < ck.num_cols) corrupt("gap stream shorter than its column table");
            ck.gaps = base_ + data_base + go; ck.values = base_ + data_base + vo;
            ck.nnz_before = before; before += ck.nnz;
        }
        if (next_col != ncols) corrupt("chunk index does not cover every column");
        total_nnz = before;
    };

    if (info_.data_offset > body_end) corrupt("data section outside the file");
    uint64_t total = 0;
    read_descs(index_offset, info_.num_chunks, info_.data_offset, body_end, info_.n, main_, total);
    if (total != info_.nnz) corrupt("chunk non-zeros do not add up to the header's nnz");

    if (info_.transpose_offset != 0) {
        // [u32 chunk count][descriptors][streams, offsets relative to the end of the descriptors]
        // (sparsepress_v2.hpp:777-797, read back at :1339-1361)
        if (info_.transpose_offset > body_end || body_end - info_.transpose_offset < 4) corrupt("transpose section truncated");
        const uint64_t t_end = (info_.metadata_offset > info_.transpose_offset && info_.metadata_offset <= body_end)
                                   ? info_.metadata_offset : body_end;
        const uint32_t tc = load<uint32_t>(base_ + info_.transpose_offset);
        const uint64_t t_index = info_.transpose_offset + 4;
        if (static_cast<uint64_t>(tc) * kDescBytes > t_end - t_index) corrupt("transpose index truncated");
        read_descs(t_index, tc, t_index + static_cast<uint64_t>(tc) * kDescBytes, t_end, info_.m, transpose_,
                   info_.transpose_nnz);
        info_.transpose_chunks = tc;
    }

    const uint8_t* f = base_ + body_end;
    if (std::memcmp(f + 12, "SPEN", 4) == 0) {
        info_.metadata_bytes = load<uint32_t>(f);
        info_.footer_crc32 = load<uint32_t>(f + 4);
    }
    const auto perm = metadata_record(2);
    info_.row_permutation_len = static_cast<uint32_t>(perm.size() / 4);
}

std::vector<uint8_t> File::metadata_record(uint8_t key) const {
    // [u32 entries] then per entry [u8 key][u32 bytes][bytes] (header_v2.hpp:372-425); a damaged tail ends the scan.
    std::vector<uint8_t> out;
    const size_t body_end = size_ - kFooterBytes;
    if (info_.metadata_offset == 0 || info_.metadata_offset >= body_end) return out;
    const uint8_t* p = base_ + info_.metadata_offset;
    const uint8_t* end = base_ + body_end;
    if (end - p < 4) return out;
    const uint32_t entries = load<uint32_t>(p); p += 4;
    for (uint32_t k = 0; k < entries && p < end; ++k) {
        const uint8_t kk = *p++;
        if (end - p < 4) break;
        const uint32_t len = load<uint32_t>(p); p += 4;
        if (static_cast<size_t>(end - p) < len) break;
        if (kk == key) { out.assign(p, p + len); return out; }
        p += len;
    }
    return out;
}

uint32_t File::compute_crc32() const {
    return crc32_update(0xFFFFFFFFu, base_, size_ - kFooterBytes) ^ 0xFFFFFFFFu;
}

const std::vector<Chunk>& File::chunks(int section) const {
    if (section == 0) return main_;
    if (section == 1) {
        if (info_.transpose_offset == 0) throw Error(kNoTranspose, "spz: the file does not contain a pre-stored transpose");
        return transpose_;
    }
    throw Error(kBadArgument, "spz: section must be 0 (A) or 1 (stored transpose)");
}

uint32_t File::section_cols(int section) const { return section == 0 ? info_.n : info_.m; }
uint32_t File::section_rows(int section) const { return section == 0 ? info_.m : info_.n; }

namespace {

// Column counts of a chunk: the varint table that follows the u32 size prefix of a non-empty gap stream
// (sparsepress_v2.hpp:143-151). An empty chunk is written WITHOUT the prefix (:94, the early return), so its
// stream is just num_cols zero bytes; every count is 0 by the descriptor and the stream is not consulted.
// Returns the cursor positioned at the rANS part of the stream.
Cursor chunk_counts(const Chunk& ck, std::vector<uint32_t>& counts) {
    counts.assign(ck.num_cols, 0);
    if (ck.nnz == 0) return Cursor{nullptr, nullptr};
    Cursor c{ck.gaps, ck.gaps + ck.gap_bytes};
    const uint32_t cc = c.u32("gap stream truncated");
    c.need(cc, "column-count table exceeds the gap stream");
    Cursor t{c.p, c.p + cc};
    uint64_t sum = 0;
    for (uint32_t j = 0; j < ck.num_cols; ++j) {
        const uint64_t v = t.varint("column-count table truncated");
        if (v > ck.nnz) corrupt("column count exceeds the chunk's nnz");
        counts[j] = static_cast<uint32_t>(v); sum += v;
    }
    if (sum != ck.nnz) corrupt("column counts do not add up to the chunk's nnz");
    c.p += cc;
    return c;
}

struct Piece {            // the part of one chunk a decode call needs
    const Chunk* ck;
    uint32_t lo, hi;      // local column range [lo, hi) inside the chunk
    uint64_t skip, take;  // entries of the chunk before column lo / inside [lo, hi)
    uint64_t out;         // first output entry
    uint32_t out_col;     // first output column (relative to c0)
};

template <typename V> inline V from_u64_bits(uint64_t b) { double d; std::memcpy(&d, &b, 8); return static_cast<V>(d); }

}  // namespace

uint64_t File::range_nnz(int section, uint32_t c0, uint32_t c1) const {
    const auto& cks = chunks(section);
    if (c0 > c1 || c1 > section_cols(section)) throw Error(kBadArgument, "spz: column range outside the matrix");
    uint64_t total = 0;
    std::vector<uint32_t> counts;
    for (const Chunk& ck : cks) {
        const uint32_t a = std::max(c0, ck.col_start), b = std::min(c1, ck.col_start + ck.num_cols);
        if (a >= b) continue;
        if (b - a == ck.num_cols) { total += ck.nnz; continue; }
        chunk_counts(ck, counts);
        for (uint32_t j = a - ck.col_start; j < b - ck.col_start; ++j) total += counts[j];
    }
    return total;
}

void File::col_counts(int section, int32_t* out, int threads) const {
    const auto& cks = chunks(section);
    parallel_for(cks.size(), threads, [&](size_t k) {
        std::vector<uint32_t> counts;
        chunk_counts(cks[k], counts);
        for (uint32_t j = 0; j < cks[k].num_cols; ++j) out[cks[k].col_start + j] = static_cast<int32_t>(counts[j]);
    });
}

template <typename V>
void File::decode(int section, uint32_t c0, uint32_t c1, int32_t* p, int32_t* i, V* x, bool reorder, int threads) const {
    const auto& cks = chunks(section);
    if (c0 > c1 || c1 > section_cols(section)) throw Error(kBadArgument, "spz: column range outside the matrix");
    const uint32_t rows = section_rows(section);
    const ValueType vt = static_cast<ValueType>(info_.value_type);

    // ---- plan: the chunks the range touches, their column counts into p[] (as counts), then one prefix sum ----
    std::vector<Piece> pieces;
    for (const Chunk& ck : cks) {
        const uint32_t a = std::max(c0, ck.col_start), b = std::min(c1, ck.col_start + ck.num_cols);
        if (a >= b) continue;
        pieces.push_back(Piece{&ck, a - ck.col_start, b - ck.col_start, 0, 0, 0, a - c0});
    }
    p[0] = 0;
    parallel_for(pieces.size(), threads, [&](size_t k) {
        Piece& pc = pieces[k];
        std::vector<uint32_t> counts;
        chunk_counts(*pc.ck, counts);
        uint64_t skip = 0, take = 0;
        for (uint32_t j = 0; j < pc.lo; ++j) skip += counts[j];
        for (uint32_t j = pc.lo; j < pc.hi; ++j) { take += counts[j]; p[pc.out_col + (j - pc.lo) + 1] = static_cast<int32_t>(counts[j]); }
        pc.skip = skip; pc.take = take;
    });
    {
        uint64_t run = 0;
        for (Piece& pc : pieces) { pc.out = run; run += pc.take; }
        if (run > 0x7FFFFFFFull) throw Error(kBadArgument, "spz: range holds more than 2^31 - 1 non-zeros");
        int32_t acc = 0;
        for (uint32_t j = 0; j < c1 - c0; ++j) { acc += p[j + 1]; p[j + 1] = acc; }
    }

    // ---- one task per chunk, largest first. Inside a task the gap stream and the value stream(s) of the chunk are decoded
    // in LOCK-STEP (1 + planes independent rANS states in one loop): a single state is a serial chain of two dependent
    // table look-ups, a multiply and the refill (~10 ns per symbol); interleaved chains overlap. ----
    std::vector<uint32_t> order;
    for (uint32_t k = 0; k < pieces.size(); ++k) if (pieces[k].take > 0) order.push_back(k);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return pieces[a].ck->gap_bytes + pieces[a].ck->value_bytes > pieces[b].ck->gap_bytes + pieces[b].ck->value_bytes;
    });

    auto chunk_task = [&](const Piece& pc) {
        const Chunk& ck = *pc.ck;
        const uint64_t need = pc.skip + pc.take, skip = pc.skip;
        std::vector<uint32_t> counts;
        Cursor c = chunk_counts(ck, counts);
        EscapedStream gs = open_escaped(c.p, c.left());
        std::vector<uint32_t> gtmp;
        uint32_t* gsym;
        if (skip == 0) gsym = reinterpret_cast<uint32_t*>(i + pc.out);
        else { gtmp.resize(need); gsym = gtmp.data(); }
        V* dst = x + pc.out;

        if (vt == kU8 || vt == kU16 || vt == kU32 || vt == kQuant8) {
            EscapedStream vs = open_escaped(ck.values, ck.value_bytes);
            std::unique_ptr<uint32_t[]> vsym(new uint32_t[need]);
            uint32_t* vv = vsym.get();
            const Plane* const pl[2] = {&gs.plane, &vs.plane};
            decode_planes<2>(pl, need, [&](uint64_t k, const uint32_t (&y)[2]) { gsym[k] = y[0]; vv[k] = y[1]; });
            if (vt == kQuant8) {
                // the writer emits no overflow section for QUANT8 and 255 is an ordinary level (sparsepress_v2.hpp:352-392)
                const float scale = ck.quant_scale, offset = ck.quant_offset;
                for (uint64_t k = 0; k < need; ++k) {
                    uint32_t q = vv[k];
                    if (q == kEscape && vs.has_overflow) q = static_cast<uint32_t>(vs.overflow.varint("overflow section truncated"));
                    if (k < skip) continue;
                    const float prod = scale * static_cast<float>(q);
                    dst[k - skip] = static_cast<V>(offset + prod);
                }
            } else if (!vs.has_overflow) {
                for (uint64_t k = skip; k < need; ++k) dst[k - skip] = static_cast<V>(vv[k]);
            } else {
                for (uint64_t k = 0; k < need; ++k) {
                    uint32_t v = vv[k];
                    if (v == kEscape) v = static_cast<uint32_t>(vs.overflow.varint("overflow section truncated"));
                    if (k >= skip) dst[k - skip] = static_cast<V>(v);
                }
            }
        } else {
            // byte planes: [u8 planes] then per plane [u32 table bytes][table][u32 enc bytes][enc] (sparsepress_v2.hpp:442-473)
            Cursor vc{ck.values, ck.values + ck.value_bytes};
            const uint32_t planes = vc.u8("value stream truncated");
            const uint32_t want = vt == kF32 ? 4u : vt == kF16 ? 2u : 8u;
            if (planes != want) corrupt("byte-plane count does not match the value type");
            Plane vp[8];
            for (uint32_t s = 0; s < planes; ++s) {
                const uint32_t tb = vc.u32("value stream truncated");
                vc.need(tb, "value stream truncated");
                Cursor t{vc.p, vc.p + tb};
                vp[s].mdl = std::make_unique<Model>();
                vp[s].mdl->parse(t);
                vc.p += tb;
                vp[s].len = vc.u32("value stream truncated");
                vc.need(vp[s].len, "value stream truncated");
                vp[s].src = vc.p; vc.p += vp[s].len;
            }
            if (vt == kF32) {
                const Plane* const pl[5] = {&gs.plane, &vp[0], &vp[1], &vp[2], &vp[3]};
                decode_planes<5>(pl, need, [&](uint64_t k, const uint32_t (&y)[5]) {
                    gsym[k] = y[0];
                    if (k < skip) return;
                    const uint32_t bits = y[1] | (y[2] << 8) | (y[3] << 16) | (y[4] << 24);
                    float f; std::memcpy(&f, &bits, 4);
                    dst[k - skip] = static_cast<V>(f);
                });
            } else if (vt == kF16) {
                const Plane* const pl[3] = {&gs.plane, &vp[0], &vp[1]};
                decode_planes<3>(pl, need, [&](uint64_t k, const uint32_t (&y)[3]) {
                    gsym[k] = y[0];
                    if (k < skip) return;
                    dst[k - skip] = static_cast<V>(half_to_float(static_cast<uint16_t>(y[1] | (y[2] << 8))));
                });
            } else {
                const Plane* const pl[9] = {&gs.plane, &vp[0], &vp[1], &vp[2], &vp[3], &vp[4], &vp[5], &vp[6], &vp[7]};
                decode_planes<9>(pl, need, [&](uint64_t k, const uint32_t (&y)[9]) {
                    gsym[k] = y[0];
                    if (k < skip) return;
                    uint64_t bits = 0;
                    for (int s = 0; s < 8; ++s) bits |= static_cast<uint64_t>(y[1 + s]) << (8 * s);
                    dst[k - skip] = from_u64_bits<V>(bits);
                });
            }
        }

        // gaps -> row indices, the running row restarting at every column (sparsepress_v2.hpp:1016-1027)
        if (gs.has_overflow)
            for (uint64_t k = 0; k < skip; ++k) if (gsym[k] == kEscape) gs.overflow.varint("overflow section truncated");
        const uint32_t* g = gsym + skip;
        int32_t* out_i = i + pc.out;
        uint64_t k = 0;
        for (uint32_t j = pc.lo; j < pc.hi; ++j) {
            // 32-bit wrap-around on purpose: a row-sorted file stores the PERMUTED rows of a column in their original
            // order, so a "gap" may be negative modulo 2^32 (the writer subtracts in uint32, sparsepress_v2.hpp:99-107)
            uint32_t row = 0;   // next admissible row
            for (uint32_t t = 0; t < counts[j]; ++t, ++k) {
                uint32_t gap = g[k];
                if (gap == kEscape && gs.has_overflow) gap = static_cast<uint32_t>(gs.overflow.varint("overflow section truncated"));
                row += gap;
                if (row >= rows) corrupt("row index outside the matrix");
                out_i[k] = static_cast<int32_t>(row);
                ++row;
            }
        }
    };

    parallel_for(order.size(), threads, [&](size_t k) { chunk_task(pieces[order[k]]); });

    // ---- stored row permutation, applied the way decompress_v2 applies it (sparsepress_v2.hpp:1093-1104): the record
    // is used as a plain map on the decoded indices; indices at or beyond its length stay. Rows inside a column are
    // NOT re-sorted afterwards (nor does the reference). ----
    if (section == 0 && reorder && info_.row_sorted) {
        const auto rec = metadata_record(2);
        if (!rec.empty()) {
            const size_t len = rec.size() / 4;
            const uint8_t* raw = rec.data();
            for (size_t k = 0; k < len; ++k) if (load<uint32_t>(raw + 4 * k) >= rows) corrupt("row permutation outside the matrix");
            const uint64_t total = pieces.empty() ? 0 : pieces.back().out + pieces.back().take;
            const size_t blocks = static_cast<size_t>((total + (1u << 16) - 1) >> 16);
            parallel_for(blocks, threads, [&](size_t b) {
                const uint64_t lo = static_cast<uint64_t>(b) << 16, hi = std::min<uint64_t>(total, lo + (1u << 16));
                for (uint64_t k = lo; k < hi; ++k) {
                    const uint32_t r = static_cast<uint32_t>(i[k]);
                    if (r < len) i[k] = static_cast<int32_t>(load<uint32_t>(raw + 4 * static_cast<size_t>(r)));
                }
            });
        }
    }
}

template void File::decode<float>(int, uint32_t, uint32_t, int32_t*, int32_t*, float*, bool, int) const;
template void File::decode<double>(int, uint32_t, uint32_t, int32_t*, int32_t*, double*, bool, int) const;

}  // namespace spz
}  // namespace b200
