// oracle/ref_fixture/spz_ref_tool.cpp — TEST INFRASTRUCTURE. Drives the REFERENCE's own StreamPress v2 codec
// (inst/include/streampress/sparsepress_v2.hpp: compress_v2 :480, decompress_v2 :897, decompress_v2_transpose :1318;
// compiled from /root/reference where it lies — no reference source is copied into this repo) so that the tests can
//   * write .spz files with the reference's writer (every value type, row sorting, pre-stored transpose, chunk sizes),
//   * decode any .spz file with the reference's reader,
// and compare both with rcppml_b200/csrc/spz_reader.cpp. Raw little-endian exchange format ("csc.bin"):
//   int32 m, n ; int64 nnz ; int32 p[n+1] ; int32 i[nnz] ; double x[nnz]
//
//   spz_ref_tool encode  in.bin out.spz <precision> <row_sort 0|1> <include_transpose 0|1> <chunk_cols> [obs_bytes var_bytes]
//                        (obs_bytes / var_bytes: sizes of opaque stand-ins for the serialized obs / var tables the
//                         writer places between the transpose section and the metadata, sparsepress_v2.hpp:810-818)
//   spz_ref_tool decode  in.spz out.bin [reorder 0|1] [col_start col_end]
//   spz_ref_tool decodet in.spz out.bin
//   spz_ref_tool time    in.spz <repeats> <threads>      (best wall-clock ms of decompress_v2 on the bytes in memory)
#include <streampress/sparsepress_v2.hpp>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

namespace {

std::vector<uint8_t> slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(3); }
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

streampress::CSCMatrix read_bin(const char* path) {
    auto buf = slurp(path);
    const uint8_t* q = buf.data();
    int32_t m, n; int64_t nnz;
    std::memcpy(&m, q, 4); std::memcpy(&n, q + 4, 4); std::memcpy(&nnz, q + 8, 8); q += 16;
    streampress::CSCMatrix A(static_cast<uint32_t>(m), static_cast<uint32_t>(n), static_cast<uint64_t>(nnz));
    std::memcpy(A.p.data(), q, 4u * (A.n + 1)); q += 4u * (A.n + 1);
    std::memcpy(A.i.data(), q, 4u * A.nnz); q += 4u * A.nnz;
    std::memcpy(A.x.data(), q, 8u * A.nnz);
    return A;
}

void write_bin(const char* path, const streampress::CSCMatrix& A) {
    FILE* o = std::fopen(path, "wb");
    if (!o) { std::fprintf(stderr, "cannot write %s\n", path); std::exit(3); }
    const int32_t m = static_cast<int32_t>(A.m), n = static_cast<int32_t>(A.n);
    const int64_t nnz = static_cast<int64_t>(A.nnz);
    std::fwrite(&m, 4, 1, o); std::fwrite(&n, 4, 1, o); std::fwrite(&nnz, 8, 1, o);
    std::fwrite(A.p.data(), 4, A.p.size(), o);
    std::fwrite(A.i.data(), 4, A.i.size(), o);
    std::fwrite(A.x.data(), 8, A.x.size(), o);
    std::fclose(o);
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: see the header of spz_ref_tool.cpp\n"); return 2; }
    const std::string cmd = argv[1];
    try {
        if (cmd == "encode" && (argc == 8 || argc == 10)) {
            auto A = read_bin(argv[2]);
            streampress::v2::CompressConfig_v2 cfg;
            cfg.precision = argv[4];
            cfg.row_sort = std::atoi(argv[5]) != 0;
            cfg.include_transpose = std::atoi(argv[6]) != 0;
            cfg.chunk_cols = static_cast<uint32_t>(std::atoi(argv[7]));
            if (argc == 10) {
                cfg.obs_buf.assign(static_cast<size_t>(std::atoi(argv[8])), 0xA5);
                cfg.var_buf.assign(static_cast<size_t>(std::atoi(argv[9])), 0x5A);
            }
            auto bytes = streampress::v2::compress_v2(A, cfg);
            streampress::v2::write_v2(argv[3], bytes);
            return 0;
        }
        if (cmd == "decode" && (argc == 4 || argc == 5 || argc == 7)) {
            auto buf = slurp(argv[2]);
            streampress::v2::DecompressConfig_v2 cfg;
            if (argc >= 5) cfg.reorder = std::atoi(argv[4]) != 0;
            if (argc == 7) { cfg.col_start = std::atoi(argv[5]); cfg.col_end = std::atoi(argv[6]); }
            write_bin(argv[3], streampress::v2::decompress_v2(buf.data(), buf.size(), cfg));
            return 0;
        }
        if (cmd == "time" && argc == 5) {
            auto buf = slurp(argv[2]);
            streampress::v2::DecompressConfig_v2 cfg;
            cfg.num_threads = std::atoi(argv[4]);
            double best = 1e300;
            for (int r = 0; r < std::atoi(argv[3]); ++r) {
                const auto t0 = std::chrono::steady_clock::now();
                auto A = streampress::v2::decompress_v2(buf.data(), buf.size(), cfg);
                const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                if (A.nnz > 0 && ms < best) best = ms;
            }
            std::printf("%.3f\n", best);
            return 0;
        }
        if (cmd == "decodet" && argc == 4) {
            auto buf = slurp(argv[2]);
            write_bin(argv[3], streampress::v2::decompress_v2_transpose(buf.data(), buf.size()));
            return 0;
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "spz_ref_tool: %s\n", e.what());
        return 1;
    }
    std::fprintf(stderr, "usage: see the header of spz_ref_tool.cpp\n");
    return 2;
}
