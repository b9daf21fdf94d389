// tools/spz_fuzz.cpp — mutation fuzzing of the .spz reader (rcppml_b200/csrc/spz_reader.cpp) under ASan + UBSan:
// byte flips, truncations and 32-bit overwrites of valid files; every outcome must be a decoded matrix or a
// b200::spz::Error — never a sanitizer report. The run cited in DESIGN.md §6c: 1500 mutations of each golden file.
//
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -Ircppml_b200/csrc -pthread \
//       -o /tmp/spz_fuzz tools/spz_fuzz.cpp rcppml_b200/csrc/spz_reader.cpp && /tmp/spz_fuzz tests/golden/spz/*.spz
#include "spz_reader.hpp"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <random>
#include <vector>
int main(int argc, char** argv) {
    int ok = 0, err = 0;
    std::mt19937_64 rng(123);
    for (int a = 1; a < argc; ++a) {
        std::ifstream f(argv[a], std::ios::binary);
        std::vector<uint8_t> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        for (int trial = 0; trial < 1500; ++trial) {
            auto data = raw;
            int kind = trial % 3;
            if (kind == 0) { int nm = 1 + rng() % 4; for (int k = 0; k < nm; ++k) data[rng() % data.size()] = (uint8_t)rng(); }
            else if (kind == 1) { data.resize(rng() % data.size()); }
            else { size_t p = rng() % data.size(); uint32_t v = (uint32_t)rng(); if (rng() & 1) v = 0xFFFFFFFFu >> (rng() % 32); for (int k = 0; k < 4 && p + k < data.size(); ++k) data[p + k] = (uint8_t)(v >> (8 * k)); }
            FILE* o = fopen("/tmp/spz_fuzz_case.spz", "wb"); fwrite(data.data(), 1, data.size(), o); fclose(o);
            try {
                b200::spz::File file("/tmp/spz_fuzz_case.spz");
                for (int sec = 0; sec < 2; ++sec) {
                    if (sec == 1 && file.info().transpose_chunks == 0) break;
                    uint32_t nc = file.section_cols(sec);
                    uint32_t c0 = nc ? rng() % (nc + 1) : 0, c1 = nc ? rng() % (nc + 1) : 0; if (c0 > c1) std::swap(c0, c1);
                    if (trial % 2) { c0 = 0; c1 = nc; }
                    uint64_t nnz = file.range_nnz(sec, c0, c1);
                    std::vector<int32_t> p(c1 - c0 + 1), i(nnz + 1); std::vector<double> x(nnz + 1);
                    file.decode<double>(sec, c0, c1, p.data(), i.data(), x.data(), true, 1 + trial % 3);
                    std::vector<int32_t> cc(nc + 1); file.col_counts(sec, cc.data(), 2);
                }
                file.compute_crc32(); file.metadata_record(0);
                ++ok;
            } catch (const b200::spz::Error&) { ++err; }
        }
    }
    printf("ok %d err %d\n", ok, err);
}
