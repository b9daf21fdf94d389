// oracle/ref_fixture/spz_to_csc.cpp — TEST INFRASTRUCTURE. Reads a StreamPress .spz file with the REFERENCE's
// own decoder (inst/include/streampress/sparsepress_v2.hpp:897, compiled from /root/reference where it lies —
// no reference source is copied into this repo) and dumps the CSC matrix as raw little-endian arrays:
//   int32 m, n ; int64 nnz ; int32 p[n+1] ; int32 i[nnz] ; float x[nnz]
// Used only to materialise the reference's real dataset (inst/extdata/pbmc3k.spz) as a fixture under
// oracle/_ref/ (git-ignored, travels to the GPU box). The product's own reader of this format is rcppml_b200/csrc/spz_reader.cpp.
#include <streampress/sparsepress_v2.hpp>

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iterator>
#include <vector>

int main(int argc, char** argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: %s in.spz out.bin\n", argv[0]); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto A = streampress::v2::decompress_v2(buf.data(), buf.size());
    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 3;
    const int32_t m = static_cast<int32_t>(A.m), n = static_cast<int32_t>(A.n);
    const int64_t nnz = static_cast<int64_t>(A.nnz);
    std::fwrite(&m, 4, 1, o); std::fwrite(&n, 4, 1, o); std::fwrite(&nnz, 8, 1, o);
    std::vector<int32_t> p(A.p.begin(), A.p.end()), i(A.i.begin(), A.i.end());
    std::vector<float> x(A.x.begin(), A.x.end());
    std::fwrite(p.data(), 4, p.size(), o); std::fwrite(i.data(), 4, i.size(), o); std::fwrite(x.data(), 4, x.size(), o);
    std::fclose(o);
    std::printf("%d x %d nnz %lld\n", m, n, static_cast<long long>(nnz));
    return 0;
}
