// spz_reader.hpp — host-side reader of the reference's StreamPress v2 sparse container (`.spz`), the on-disk ingest row
// of SURVEY.md §8f-4. The FORMAT is the reference's (inst/include/streampress/format/header_v2.hpp:1-17 file layout,
// :119-176 header, :205-225 chunk descriptor, :233-266 footer; sparsepress_v2.hpp:79-171 gap stream, :179-392 value
// streams; codec/rans.hpp:24-25,186-262 byte-renormalised rANS with 14 probability bits; codec/varint.hpp:20-61);
// the reader is written for this engine:
//   * the file is mapped, not copied; every read is bounds-checked (a damaged file is an error, never a wild pointer:
//     row indices are checked against the dimension, column counts against the chunk's nnz, tables against 2^14);
//   * the unit of work is a chunk STREAM, not a chunk: the gap stream and the value stream of a chunk are separate
//     tasks on a pool of host threads, and the byte-shuffled planes of a float column block (4 for fp32, 2 for fp16,
//     8 for fp64) are decoded in lock-step — N independent rANS states in one loop, whole values stored once;
//   * it decodes straight into the engine's types (int32 pointers / indices, float or double values) in caller-owned
//     buffers (pinned staging in the ingest entries), and it can decode any column range [c0, c1) of A or of the
//     pre-stored transpose — so every GPU of a sharded fit decodes only its own column block and row block.
// No CUDA in this file: the device side of the ingest is rcppml_sp_read_gpu / rcppml_b200_set_matrix_spz (abi_spz.cu).
#pragma once

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200 {
namespace spz {

// Status codes of rcppml_sp_read_gpu (src/sp_gpu_bridge.cu:57-116): 1 cannot open, 2 read failed, 3 file too small,
// 4 not a v2 file, 5 decode error. 6+ are this reader's own.
enum Status : int {
    kOk = 0, kCannotOpen = 1, kReadFailed = 2, kTooSmall = 3, kNotV2 = 4, kCorrupt = 5, kNoTranspose = 6, kBadArgument = 7
};

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};

// header_v2.hpp:45-53
enum ValueType : uint8_t { kU8 = 0, kU16 = 1, kU32 = 2, kF32 = 3, kF16 = 4, kQuant8 = 5, kF64 = 6 };

struct Info {
    uint32_t m = 0, n = 0;
    uint64_t nnz = 0;
    uint32_t chunk_cols = 0, num_chunks = 0;
    uint8_t  value_type = 0;
    uint8_t  row_sorted = 0;
    float    density = 0.f;
    uint64_t data_offset = 0, transpose_offset = 0, metadata_offset = 0;
    uint64_t obs_table_offset = 0, var_table_offset = 0;
    uint32_t transp_chunk_cols = 0;
    uint32_t transpose_chunks = 0;      // 0 when the file has no pre-stored transpose
    uint64_t transpose_nnz = 0;
    uint64_t file_bytes = 0;
    uint32_t metadata_bytes = 0;        // footer.metadata_size
    uint32_t footer_crc32 = 0;
    uint32_t row_permutation_len = 0;   // entries of the ROW_PERMUTATION metadata record (0 = none)
};

// One chunk of a section as the descriptor table states it (header_v2.hpp:205-225), offsets made absolute.
struct Chunk {
    uint32_t col_start, num_cols, nnz;
    const uint8_t* gaps;   size_t gap_bytes;
    const uint8_t* values; size_t value_bytes;
    float quant_scale, quant_offset;
    uint64_t nnz_before;   // non-zeros of the section in earlier chunks
};

class File {
public:
    explicit File(const char* path);   // throws Error
    ~File();
    File(const File&) = delete;
    File& operator=(const File&) = delete;

    const Info& info() const { return info_; }
    const uint8_t* data() const { return base_; }
    size_t size() const { return size_; }

    // section 0: A (m x n, n columns); section 1: the pre-stored CSC(A^T) (n x m, m columns).
    uint32_t section_cols(int section) const;
    uint32_t section_rows(int section) const;
    // Non-zeros of columns [c0, c1) — chunk-aligned ends cost nothing, others parse one column-count table.
    uint64_t range_nnz(int section, uint32_t c0, uint32_t c1) const;
    // Non-zeros of every column of a section (the varint tables only — no entropy decoding): what a work-balanced
    // partition of a sharded fit needs (columns of A from section 0, rows of A from section 1).
    void col_counts(int section, int32_t* out, int threads) const;

    // Decode columns [c0, c1) of a section: p[0 .. c1-c0] (rebased to 0), i[], x[] sized by range_nnz.
    // reorder: apply the stored row permutation exactly as the reference's decompress_v2 does
    // (sparsepress_v2.hpp:1093-1104; section 0 only — decompress_v2_transpose never reorders). threads <= 0: all cores.
    template <typename V>
    void decode(int section, uint32_t c0, uint32_t c1, int32_t* p, int32_t* i, V* x, bool reorder, int threads) const;

    // CRC-32 (zlib polynomial, format/checksum.hpp:18-70) of everything before the footer, as the writer computed it.
    uint32_t compute_crc32() const;

    // Raw bytes of a metadata record (header_v2.hpp:108-113: 0 rownames, 1 colnames, 2 row permutation); empty if absent.
    std::vector<uint8_t> metadata_record(uint8_t key) const;

private:
    const std::vector<Chunk>& chunks(int section) const;
    void parse();

    int fd_ = -1;
    const uint8_t* base_ = nullptr;
    size_t size_ = 0;
    Info info_;
    std::vector<Chunk> main_, transpose_;
};

}  // namespace spz
}  // namespace b200
