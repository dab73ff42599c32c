// convert_main.cpp -- drop-in for the reference's `convert` tool
// (reference src/libfm/tools/convert.cpp:62-205): libfm text -> the binary pair
//   <ofilex>: file_header {uint id=2; uint float_size=4; uint64 num_values; uint num_rows;
//             uint num_cols} (util/fmatrix.h:44-50) + per row {uint size; size x {uint id; float value}}
//   <ofiley>: {uint version=1; uint type_size=4; uint num_rows} + float[num_rows]
//             (util/matrix.h:346-362)
// Same flags (-ifile, -ofilex, -ofiley), byte-identical output files; the text is parsed
// by all host cores (host/sparse_data.h) instead of two sscanf passes on one.
#include <cstdint>
#include <fstream>
#include <iostream>
#include <string>

#include "cli_flags.h"
#include "sparse_data.h"

int main(int argc, char** argv) {
  try {
    host::CmdLine cmd(argc, argv);
    const std::string p_ifile = cmd.add("ifile", "input file name, file has to be in binary sparse format [MANDATORY]");
    const std::string p_ofilex = cmd.add("ofilex", "output file name for x [MANDATORY]");
    const std::string p_ofiley = cmd.add("ofiley", "output file name for y [MANDATORY]");
    const std::string p_help = cmd.add("help", "this screen");
    if (cmd.has(p_help) || argc == 1) {
      cmd.print_help();
      return 0;
    }
    cmd.check();
    host::SparseData d;
    d.load_text_file(cmd.str(p_ifile));

    std::ofstream out_x(cmd.str(p_ofilex).c_str(), std::ios::out | std::ios::binary);
    if (!out_x.is_open()) throw "unable to open " + cmd.str(p_ofilex);
    std::ofstream out_y(cmd.str(p_ofiley).c_str(), std::ios::out | std::ios::binary);
    if (!out_y.is_open()) throw "unable to open " + cmd.str(p_ofiley);

    struct {
      uint32_t id, float_size;
      uint64_t num_values;
      uint32_t num_rows, num_cols;
    } fh = {2u, (uint32_t)sizeof(float), d.num_values(), (uint32_t)d.num_cases(), (uint32_t)d.num_feature};
    static_assert(sizeof(fh) == 24, "file_header layout");
    out_x.write(reinterpret_cast<const char*>(&fh), sizeof(fh));
    std::string row;
    for (uint64_t r = 0; r < d.num_cases(); r++) {
      const uint64_t a = d.row_ptr[r], b = d.row_ptr[r + 1];
      const uint32_t size = (uint32_t)(b - a);
      row.resize(4 + (size_t)size * 8);
      memcpy(&row[0], &size, 4);
      for (uint32_t j = 0; j < size; j++) {
        memcpy(&row[4 + 8 * (size_t)j], &d.col[a + j], 4);
        memcpy(&row[8 + 8 * (size_t)j], &d.val[a + j], 4);
      }
      out_x.write(row.data(), (std::streamsize)row.size());
    }
    const uint32_t yh[3] = {1u, (uint32_t)sizeof(float), (uint32_t)d.num_cases()};
    out_y.write(reinterpret_cast<const char*>(yh), sizeof(yh));
    out_y.write(reinterpret_cast<const char*>(d.target.data()), (std::streamsize)(sizeof(float) * d.target.size()));
    return 0;
  } catch (std::string& e) {
    std::cerr << e << std::endl;  // the reference prints the bare message (convert.cpp:200-202)
  } catch (char const*& e) {
    std::cerr << e << std::endl;
  } catch (const std::exception& e) {  // e.g. bad_alloc on a corrupt size field
    std::cerr << std::endl << "ERROR: " << e.what() << std::endl;
  }
  return 1;
}
