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
// (num_feature = max id + 1, min/max target) are bit-exact contracts.  Text files are
// parsed by all host cores (the reference: two sscanf passes on one core).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <functional>
#include <string>
#include <thread>
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

  // text input only (the `convert` tool never looks for binary siblings of its input)
  void load_text_file(const std::string& filename) { load_text(filename); }

  // libfm.cpp:302-303
  void binarize_targets() {
    for (auto& t : target) t = (t <= 0.0f) ? -1.0f : 1.0f;
  }

 private:
  static std::string parse_error(const std::string& line, char at) {
    return "cannot parse line \"" + line + "\" at character " + at;
  }

  // One contiguous piece of the file, parsed by one thread.
  struct TextChunk {
    std::vector<uint64_t> row_end;  // entries so far after each row (chunk-local)
    std::vector<uint32_t> col;
    std::vector<float> val, target;
    int max_id = 0;
    bool has_feature = false;
    float min_target = +std::numeric_limits<float>::max();
    float max_target = -std::numeric_limits<float>::max();
    std::string error;  // first parse error of the chunk, in file order
  };

  // Parse the lines in [begin, end) (each made NUL-terminated in place).  Same grammar
  // as Data::load (Data.h:192-225): "%f" then repeated "%d:%f", '#' comments.
  static void parse_chunk(char* begin, char* end, TextChunk& out) {
    char* line = begin;
    while (line < end) {
      char* nl = static_cast<char*>(memchr(line, '\n', (size_t)(end - line)));
      char* line_end = nl ? nl : end;
      *line_end = 0;
      const char* p = line;
      while (*p == ' ' || *p == '\t') p++;
      if (*p != 0 && *p != '#') {  // Data.h:200-201: blank and comment lines are skipped
        char* e0 = nullptr;
        const float y = strtof(p, &e0);  // "%f"
        if (e0 == p) {
          out.error = parse_error(line, p[0]);
          return;
        }
        p = e0;
        out.target.push_back(y);
        if (y < out.min_target) out.min_target = y;
        if (y > out.max_target) out.max_target = y;
        for (;;) {
          // "%d:%f" -- %d skips white space, ':' must follow the digits directly,
          // %f skips white space again
          const char* q = p;
          while (*q == ' ' || *q == '\t') q++;
          char* e1 = nullptr;
          const long id = strtol(q, &e1, 10);
          if (e1 == q || *e1 != ':') break;
          char* e2 = nullptr;
          const float x = strtof(e1 + 1, &e2);
          if (e2 == e1 + 1) break;
          out.col.push_back((uint32_t)(int)id);
          out.val.push_back(x);
          if ((int)id > out.max_id) out.max_id = (int)id;
          out.has_feature = true;
          p = e2;
        }
        while (*p == ' ' || *p == '\t') p++;
        if (*p != 0 && *p != '#') {  // Data.h:218-220
          out.error = parse_error(line, p[0]);
          return;
        }
        out.row_end.push_back(out.col.size());
      }
      line = line_end + 1;
    }
  }

  // The reference parses the file twice with sscanf on one core (Data.h:180-290) and
  // that dominates its wall time on large inputs.  Here the
  // file is read once, cut at line boundaries and parsed by all host cores; the chunks
  // are concatenated in file order, so the CSR is identical to the sequential result.
  void load_text(const std::string& filename) {
    std::string buf;
    {
      std::ifstream in(filename.c_str(), std::ios::binary);
      if (!in.is_open()) throw "unable to open " + filename;
      in.seekg(0, std::ios::end);
      const std::streamoff len = in.tellg();
      in.seekg(0, std::ios::beg);
      buf.resize((size_t)(len > 0 ? len : 0) + 1);
      if (len > 0) in.read(&buf[0], len);
      buf[buf.size() - 1] = 0;
    }
    char* base = &buf[0];
    char* stop = base + buf.size() - 1;
    unsigned hw = std::thread::hardware_concurrency();
    size_t n_chunks = hw ? hw : 4;
    if (n_chunks > 32) n_chunks = 32;
    if ((size_t)(stop - base) < (1u << 20)) n_chunks = 1;  // small files: not worth a thread
    std::vector<char*> cut(n_chunks + 1);
    cut[0] = base;
    cut[n_chunks] = stop;
    for (size_t i = 1; i < n_chunks; i++) {
      char* guess = base + (size_t)(stop - base) * i / n_chunks;
      if (guess < cut[i - 1]) guess = cut[i - 1];
      char* nl = static_cast<char*>(memchr(guess, '\n', (size_t)(stop - guess)));
      cut[i] = nl ? nl + 1 : stop;
    }
    std::vector<TextChunk> chunks(n_chunks);
    std::vector<std::thread> workers;
    for (size_t i = 1; i < n_chunks; i++)
      workers.emplace_back(parse_chunk, cut[i], cut[i + 1], std::ref(chunks[i]));
    parse_chunk(cut[0], cut[1], chunks[0]);
    for (auto& w : workers) w.join();

    uint64_t rows = 0, entries = 0;
    for (const auto& c : chunks) {
      if (!c.error.empty()) throw c.error;  // the first error in file order
      rows += c.target.size();
      entries += c.col.size();
    }
    row_ptr.assign(1, 0);
    row_ptr.reserve(rows + 1);
    col.resize(entries);
    val.resize(entries);
    target.resize(rows);
    bool has_feature = false;
    int max_id = 0;
    uint64_t r0 = 0, e0 = 0;
    for (const auto& c : chunks) {
      for (uint64_t re : c.row_end) row_ptr.push_back(e0 + re);
      if (!c.col.empty()) {
        memcpy(&col[e0], c.col.data(), c.col.size() * sizeof(uint32_t));
        memcpy(&val[e0], c.val.data(), c.val.size() * sizeof(float));
      }
      if (!c.target.empty()) memcpy(&target[r0], c.target.data(), c.target.size() * sizeof(float));
      r0 += c.target.size();
      e0 += c.col.size();
      has_feature |= c.has_feature;
      if (c.max_id > max_id) max_id = c.max_id;
      if (c.min_target < min_target) min_target = c.min_target;
      if (c.max_target > max_target) max_target = c.max_target;
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
      {  // a truncated / corrupt header must not size the arrays: bound num_values by the file
        in.seekg(0, std::ios::end);
        const uint64_t fsize = (uint64_t)in.tellg();
        in.seekg(sizeof(fh), std::ios::beg);
        const uint64_t fixed = sizeof(fh) + 4ull * fh.num_rows;
        if (!in || fsize < fixed || fh.num_values > (fsize - fixed) / 8) throw "could not read " + fx;
      }
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
      if (pos != fh.num_values) throw "could not read " + fx;
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
