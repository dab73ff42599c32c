// Test helper: run the CLI's loader (libfm_b200/host/sparse_data.h) on a file and dump
// the CSR it produced as raw little-endian arrays, for comparison with the reference's
// Data::load (oracle/_ref).  Built on the fly by tests/test_host_cpu.py.
#include <cstdio>
#include <iostream>
#include <sstream>
#include "sparse_data.h"

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  host::SparseData d;
  std::ostringstream sink;
  std::streambuf* saved = std::cout.rdbuf(sink.rdbuf());
  try {
    d.load(argv[1]);
  } catch (std::string& e) {
    std::cout.rdbuf(saved);
    std::cerr << "ERROR: " << e << std::endl;
    return 1;
  }
  std::cout.rdbuf(saved);
  FILE* f = fopen(argv[2], "wb");
  uint64_t n = d.num_cases(), nnz = d.num_values();
  int64_t nf = d.num_feature;
  fwrite(&n, 8, 1, f);
  fwrite(&nnz, 8, 1, f);
  fwrite(&nf, 8, 1, f);
  fwrite(&d.min_target, 4, 1, f);
  fwrite(&d.max_target, 4, 1, f);
  fwrite(d.row_ptr.data(), 8, n + 1, f);
  fwrite(d.col.data(), 4, nnz, f);
  fwrite(d.val.data(), 4, nnz, f);
  fwrite(d.target.data(), 4, n, f);
  fclose(f);
  return 0;
}
