// tests/host/test_viz_output.cpp — CPU: the molecule dumps of the host adapter against byte strings built here from the
// reference's format description (src4/viz_output_event.cpp:65-265): ASCII lines with %.9g, CellBlender binary v1/v2
// with species grouping, float32 positions in micrometres and normals for surface species only.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include "../../mcell_b200/host/mcx_host.h"
using namespace MCell;

static std::string slurp(const std::string& path) { std::ifstream f(path, std::ios::binary); std::stringstream ss; ss << f.rdbuf(); return ss.str(); }
template <class T> static void put(std::string& s, T v) { s.append(reinterpret_cast<const char*>(&v), sizeof(T)); }

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : "/tmp";
  if (VizOutputWriter::iterations_to_string(7, 1000) != "0007") return 1;
  if (VizOutputWriter::iterations_to_string(7, 999) != "007") return 1;
  if (VizOutputWriter::iterations_to_string(0, 0) != "0") return 1;
  if (VizOutputWriter::iterations_to_string(10, 10) != "10") return 1;
  if (VizOutputWriter::file_name("viz_data/seed_00001/Scene", ASCII_MODE, 20, 100) != "viz_data/seed_00001/Scene.ascii.020.dat") return 1;
  if (VizOutputWriter::file_name("Scene", CELLBLENDER_MODE_V1, 5, 5) != "Scene.cellbin.5.dat") return 1;

  const double length_unit = 0.01;
  std::vector<VizSpeciesInfo> species = {{"A", false}, {"unused", false}, {"R", true}};
  // one wall in the plane z = 12.5 lu: v0 = (10, 20, 12.5), unit_u = x, unit_v = y, normal = z
  std::vector<VizWallFrame> walls = {{Vec3{10, 20, 12.5}, Vec3{1, 0, 0}, Vec3{0, 1, 0}, Vec3{0, 0, 1}}};
  std::vector<Molecule> mols;
  mols.push_back(Molecule(0, 0, Vec3{1.0 / 3.0, -250.125, 1e-7}, 0));
  Molecule r; r.id = 1; r.species_id = 2; r.flags = MOLECULE_FLAG_SURF; r.s.pos = Vec2{0.5, 0.25}; r.s.orientation = -1; r.s.wall_index = 0; r.s.grid_tile_index = 3;
  mols.push_back(r);
  Molecule dead(2, 0, Vec3{1, 2, 3}, 0); dead.flags |= MOLECULE_FLAG_DEFUNCT;
  mols.push_back(dead);
  mols.push_back(Molecule(3, 0, Vec3{49.999999999, 0, -50}, 0));

  // ASCII: molecules in container order, defunct ones skipped
  const std::string pa = dir + "/" + VizOutputWriter::file_name("Scene", ASCII_MODE, 1, 10);
  if (!VizOutputWriter::write(pa, ASCII_MODE, mols, species, &walls, length_unit)) return 2;
  char want[512];
  snprintf(want, sizeof(want),
           // orientation * normal component: -1 * 0 = -0, printed "-0" by %.9g exactly like the reference's fprintf
           "A 0 %.9g %.9g %.9g 0 0 0\nR 1 %.9g %.9g %.9g -0 -0 -1\nA 3 %.9g %.9g %.9g 0 0 0\n",
           (1.0 / 3.0) * length_unit, -250.125 * length_unit, 1e-7 * length_unit,
           (0.5 * 1.0 + 0.25 * 0.0 + 10.0) * length_unit, (0.5 * 0.0 + 0.25 * 1.0 + 20.0) * length_unit, 12.5 * length_unit,
           49.999999999 * length_unit, 0.0, -50 * length_unit);
  if (slurp(pa) != want) { printf("ascii mismatch:\n%s---\n%s", slurp(pa).c_str(), want); return 3; }
  if (slurp(pa).find("A 0 0.00333333333 -2.50125 1e-09 0 0 0\n") != 0) { printf("ascii literal mismatch\n"); return 3; }

  // CellBlender v1 and v2
  for (int ver = 1; ver <= 2; ver++) {
    const viz_mode_t mode = ver == 1 ? CELLBLENDER_MODE_V1 : CELLBLENDER_MODE_V2;
    const std::string pb = dir + "/" + VizOutputWriter::file_name("Scene", mode, 1, 10) + (ver == 1 ? ".v1" : ".v2");
    if (!VizOutputWriter::write(pb, mode, mols, species, &walls, length_unit)) return 4;
    std::string w;
    put<uint32_t>(w, (uint32_t)ver);
    // species A: two live molecules (ids 0 and 3)
    if (ver == 1) put<unsigned char>(w, 1); else put<uint32_t>(w, 1);
    w += "A"; put<unsigned char>(w, 0);
    put<uint32_t>(w, ver == 1 ? 6u : 2u);
    if (ver == 2) { put<uint32_t>(w, 0u); put<uint32_t>(w, 3u); }
    put<float>(w, (float)((1.0 / 3.0) * length_unit)); put<float>(w, (float)(-250.125 * length_unit)); put<float>(w, (float)(1e-7 * length_unit));
    put<float>(w, (float)(49.999999999 * length_unit)); put<float>(w, 0.0f); put<float>(w, (float)(-50 * length_unit));
    // species "unused" has no molecules: skipped.  Species R: surface, one molecule, position then normal
    if (ver == 1) put<unsigned char>(w, 1); else put<uint32_t>(w, 1);
    w += "R"; put<unsigned char>(w, 1);
    put<uint32_t>(w, ver == 1 ? 3u : 1u);
    if (ver == 2) put<uint32_t>(w, 1u);
    put<float>(w, (float)(10.5 * length_unit)); put<float>(w, (float)(20.25 * length_unit)); put<float>(w, (float)(12.5 * length_unit));
    put<float>(w, -0.0f); put<float>(w, -0.0f); put<float>(w, -1.0f);
    if (slurp(pb) != w) { printf("cellblender v%d mismatch (%zu vs %zu bytes)\n", ver, slurp(pb).size(), w.size()); return 5; }
  }
  // species filter: only R
  const std::string pf = dir + "/only_R.dat";
  if (!VizOutputWriter::write(pf, ASCII_MODE, mols, species, &walls, length_unit, {2})) return 6;
  if (slurp(pf).find("R 1 ") != 0 || slurp(pf).find("A ") != std::string::npos) return 6;
  // the event: no device attached (diffuse == nullptr), dumps the host container as it is
  PartitionMolecules part; part.molecules = mols; part.rebuild_mapping();
  GpuVizOutputEvent ev(nullptr, &part, ASCII_MODE, dir + "/Ev", 1000, length_unit);
  ev.species = species; ev.walls = walls; ev.event_time = 40.0;
  ev.step();
  if (ev.last_file != dir + "/Ev.ascii.0040.dat" || slurp(ev.last_file) != want) { printf("event mismatch %s\n", ev.last_file.c_str()); return 7; }
  if (!ev.is_barrier() || ev.type_index != 300) return 8;
  printf("viz output ok\n");
  return 0;
}
