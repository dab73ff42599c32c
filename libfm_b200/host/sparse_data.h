// sparse_data.h -- the host-side sparse loader of the drop-in command line.
//
// Produces the SoA CSR that libfmb200 consumes (row offsets, column ids, values,
// targets) from the inputs the reference's Data::load accepts (reference
// src/libfm/src/Data.h:113-290):
//   * libfm text:  `target id:value id:value ...`  (# comments, blank lines)
//   * binary pair: <file>.x + <file>.y (or .data + .target), the format the
//     reference's `convert` tool writes (src/libfm/tools/convert.cpp:143-198,
//     util/fmatrix.h:44-50, util/matrix.h:364-380)
// Ordering (rows in file order, entries in line order) and the derived numbers
// (num_feature = max id + 1, min/max target) are bit-exact contracts.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

namespace host {

struct SparseData {
  std::vector<uint64_t> row_ptr{0};
  std::vector<uint32_t> col;
  std::vector<float> val;
  std::vector<float> target;
  int num_feature = 0;
  float min_target = +std::numeric_limits<float>::max();
  float max_target = -std::numeric_limits<float>::max();

  uint64_t num_cases() const { return row_ptr.size() - 1; }
  uint64_t num_values() const { return row_ptr.back(); }

  static bool file_exists(const std::string& f) {
    std::ifstream in(f.c_str());
    return in.is_open();
  }

  // Data::load (Data.h:113-290).  Prints the same progress lines.
  void load(const std::string& filename) {
    std::cout << "has x = " << 1 << std::endl;
    std::cout << "has xt = " << 0 << std::endl;
    if (file_exists(filename + ".data") && file_exists(filename + ".target")) {
      load_binary(filename + ".data", filename + ".target");
    } else if (file_exists(filename + ".x") && file_exists(filename + ".y")) {
      load_binary(filename + ".x", filename + ".y");
    } else {
      load_text(filename);
    }
  }

  // libfm.cpp:302-303
  void binarize_targets() {
    for (auto& t : target) t = (t <= 0.0f) ? -1.0f : 1.0f;
  }

 private:
  static std::string parse_error(const std::string& line, char at) {
    return "cannot parse line \"" + line + "\" at character " + at;
  }

  void load_text(const std::string& filename) {
    std::ifstream in(filename.c_str());
    if (!in.is_open()) throw "unable to open " + filename;
    bool has_feature = false;
    int max_id = 0;
    std::string line;
    while (std::getline(in, line)) {
      const char* p = line.c_str();
      while (*p == ' ' || *p == '\t') p++;
      if (*p == 0 || *p == '#') continue;  // Data.h:200-201
      char* end = nullptr;
      float y = strtof(p, &end);  // "%f"
      if (end == p) throw parse_error(line, p[0]);
      p = end;
      target.push_back(y);
      if (y < min_target) min_target = y;
      if (y > max_target) max_target = y;
      for (;;) {
        // "%d:%f" -- %d skips white space, ':' must follow the digits directly,
        // %f skips white space again
        const char* q = p;
        while (*q == ' ' || *q == '\t') q++;
        char* e1 = nullptr;
        long id = strtol(q, &e1, 10);
        if (e1 == q || *e1 != ':') break;
        char* e2 = nullptr;
        float x = strtof(e1 + 1, &e2);
        if (e2 == e1 + 1) break;
        col.push_back((uint32_t)(int)id);
        val.push_back(x);
        if ((int)id > max_id) max_id = (int)id;
        has_feature = true;
        p = e2;
      }
      while (*p == ' ' || *p == '\t') p++;
      if (*p != 0 && *p != '#') throw parse_error(line, p[0]);  // Data.h:218-220
      row_ptr.push_back(col.size());
    }
    num_feature = has_feature ? max_id + 1 : 0;  // Data.h:227-229
    std::cout << "num_rows=" << num_cases() << "\tnum_values=" << num_values()
              << "\tnum_features=" << num_feature << "\tmin_target=" << min_target
              << "\tmax_target=" << max_target << std::endl;
  }

  void load_binary(const std::string& fx, const std::string& fy) {
    // target vector: {uint version=1, uint type_size=4, uint n} + float[n]  (matrix.h:364-380)
    {
      std::ifstream in(fy.c_str(), std::ios::binary);
      uint32_t hdr[3];
      in.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
      if (!in || hdr[0] != 1 || hdr[1] != sizeof(float)) throw "could not read " + fy;
      target.resize(hdr[2]);
      in.read(reinterpret_cast<char*>(target.data()), sizeof(float) * (size_t)hdr[2]);
      if (!in) throw "could not read " + fy;
    }
    // matrix: file_header (24 B) then per row {uint size; size x {uint id; float value}}
    {
      std::cout << "data... ";
      std::ifstream in(fx.c_str(), std::ios::binary);
      if (!in.is_open()) throw "could not open " + fx;
      struct {
        uint32_t id, float_size;
        uint64_t num_values;
        uint32_t num_rows, num_cols;
      } fh;
      static_assert(sizeof(fh) == 24, "file_header layout (util/fmatrix.h:44-50)");
      in.read(reinterpret_cast<char*>(&fh), sizeof(fh));
      if (!in || fh.id != 2 || fh.float_size != sizeof(float)) throw "could not read " + fx;
      if (fh.num_rows != target.size()) throw "row count of " + fx + " and " + fy + " differ";
      col.resize(fh.num_values);
      val.resize(fh.num_values);
      row_ptr.assign(1, 0);
      row_ptr.reserve((size_t)fh.num_rows + 1);
      std::vector<char> buf;
      uint64_t pos = 0;
      for (uint32_t r = 0; r < fh.num_rows; r++) {
        uint32_t size = 0;
        in.read(reinterpret_cast<char*>(&size), sizeof(size));
        if (!in || pos + size > fh.num_values) throw "could not read " + fx;
        buf.resize((size_t)size * 8);
        in.read(buf.data(), buf.size());
        if (!in) throw "could not read " + fx;
        for (uint32_t j = 0; j < size; j++) {
          memcpy(&col[pos + j], buf.data() + 8 * (size_t)j, 4);
          memcpy(&val[pos + j], buf.data() + 8 * (size_t)j + 4, 4);
        }
        pos += size;
        row_ptr.push_back(pos);
      }
      num_feature = (int)fh.num_cols;
    }
    for (float y : target) {  // Data.h:166-171
      if (y < min_target) min_target = y;
      if (y > max_target) max_target = y;
    }
    std::cout << "num_cases=" << num_cases() << "\tnum_values=" << num_values()
              << "\tnum_features=" << num_feature << "\tmin_target=" << min_target
              << "\tmax_target=" << max_target << std::endl;
  }
};

}  // namespace host
