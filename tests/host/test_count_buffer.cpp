// tests/host/test_count_buffer.cpp — CPU: the observable text output of the host adapter is byte-identical to the
// reference's formats (src4/count_buffer.cpp:30-61,77-100): expected strings below are what
// `out << time << " " << value` and the .gdat writer print for these values.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <iomanip>
#include "../../mcell_b200/host/mcx_host.h"
using namespace MCell;

static std::string slurp(const std::string& path) { std::ifstream f(path); std::stringstream ss; ss << f.rdbuf(); return ss.str(); }

// what the reference's stream formatting prints (std::ostream defaults / scientific with precision 8)
static std::string ref_dat(double t, double v) { std::ostringstream o; o << t << " " << v << "\n"; return o.str(); }
static std::string ref_gdat(double d) {
  std::ostringstream o; o << std::scientific << std::setprecision(8) << d; return o.str();
}

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : "/tmp";
  const double vals[][2] = {{0, 500000}, {1e-05, 499873}, {0.00012, 123456789}, {2.5e-06, 0}, {1, 1e21}, {3.3333333e-4, 7}};
  for (auto& tv : vals) {
    if (CountBuffer::format_dat_row(CountItem{tv[0], tv[1]}) != ref_dat(tv[0], tv[1])) { printf("dat mismatch for %g %g\n", tv[0], tv[1]); return 1; }
    if (CountBuffer::format_gdat_value(tv[1]) != ref_gdat(tv[1])) { printf("gdat mismatch %g: %s vs %s\n", tv[1], CountBuffer::format_gdat_value(tv[1]).c_str(), ref_gdat(tv[1]).c_str()); return 1; }
  }
  if (CountBuffer::format_dat_row(CountItem{1e-05, 499873}) != "1e-05 499873\n") return 1;
  if (CountBuffer::format_gdat_value(499873) != "4.99873000e+05") return 1;
  {
    CountBuffer b(dir + "/A.World.dat", {}, 2, CountOutputFormat::DAT);
    b.add(0, CountItem{0, 500000}); b.add(0, CountItem{1e-05, 499873}); b.add(0, CountItem{2e-05, 499741});
  }
  if (slurp(dir + "/A.World.dat") != "0 500000\n1e-05 499873\n2e-05 499741\n") { printf("dat file mismatch\n"); return 1; }
  {
    CountBuffer g(dir + "/counts.gdat", {"A", "a_rather_long_observable_name"}, 10, CountOutputFormat::GDAT);
    g.add(0, CountItem{0, 10}); g.add(1, CountItem{0, 0});
    g.add(0, CountItem{1e-05, 9}); g.add(1, CountItem{1e-05, 1});
  }
  const std::string want = "#          time               A a_rather_long_observable_name\n"
                           " 0.00000000e+00  1.00000000e+01  0.00000000e+00\n"
                           " 1.00000000e-05  9.00000000e+00  1.00000000e+00\n";
  if (slurp(dir + "/counts.gdat") != want) { printf("gdat file mismatch:\n%s", slurp(dir + "/counts.gdat").c_str()); return 1; }
  printf("count buffer ok\n");
  return 0;
}
