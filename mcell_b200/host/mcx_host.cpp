// mcx_host.cpp — see mcx_host.h.  Pure marshalling; every computation happens behind the C ABI on the GPU.
#include "mcx_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>

namespace MCell {

Molecule& PartitionMolecules::add_volume_molecule(species_id_t species, const Vec3& pos, double birthday) {
  Molecule m(next_molecule_id++, species, pos, birthday);
  m.flags |= MOLECULE_FLAG_SCHEDULE_UNIMOL_RXN;  // partition.h:588-590: lifetime is drawn when first diffused
  m.diffusion_time = birthday;
  molecules.push_back(m);
  if (molecule_id_to_index_mapping.size() <= m.id) molecule_id_to_index_mapping.resize(m.id + 1, INDEX_INVALID32);
  molecule_id_to_index_mapping[m.id] = (uint32_t)molecules.size() - 1;
  return molecules.back();
}

void PartitionMolecules::rebuild_mapping() {
  molecule_id_t max_id = 0;
  for (const Molecule& m : molecules) max_id = std::max(max_id, m.id);
  molecule_id_to_index_mapping.assign(molecules.empty() ? 0 : (size_t)max_id + 1, INDEX_INVALID32);
  for (size_t i = 0; i < molecules.size(); i++) molecule_id_to_index_mapping[molecules[i].id] = (uint32_t)i;
  if (!molecules.empty()) next_molecule_id = std::max(next_molecule_id, max_id + 1);
}

void GpuDiffuseReactEvent::check(int rc, const char* what) {
  if (rc == MCX_OK) return;
  std::string msg = std::string(what) + ": " + mcx_last_error(h);
  throw McxFatalError(rc, msg);
}

GpuDiffuseReactEvent::GpuDiffuseReactEvent(const GpuModelTables& t, PartitionMolecules* partition)
    : BaseEvent(EVENT_TYPE_INDEX_DIFFUSE_REACT), p(partition), n_species(t.species.size()), n_rules(0),
      time_up_to_next_barrier(DIFFUSE_REACT_EVENT_PERIODICITY) {
  periodicity_interval = DIFFUSE_REACT_EVENT_PERIODICITY;  // repeat this event each iteration
  int rc = mcx_create(&t.cfg, &h);
  if (rc != MCX_OK) throw McxFatalError(rc, std::string("mcx_create: ") + mcx_last_error(nullptr));
  check(mcx_set_species(h, t.species.data(), (uint32_t)t.species.size()), "mcx_set_species");
  check(mcx_set_reactions(h, t.rxn_classes.data(), (uint32_t)t.rxn_classes.size(), t.pathways.data(),
                          (uint32_t)t.pathways.size()), "mcx_set_reactions");
  check(mcx_set_surface_classes(h, t.surf_class_rxns.data(), (uint32_t)t.surf_class_rxns.size()), "mcx_set_surface_classes");
  check(mcx_set_geometry(h, t.vertices.data(), t.vertices.size() / 3, t.wall_vertex_indices.data(),
                         t.wall_vertex_indices.size() / 3, t.wall_surf_class.empty() ? nullptr : t.wall_surf_class.data(),
                         nullptr), "mcx_set_geometry");
  n_walls = t.wall_vertex_indices.size() / 3;
  for (const mcx_pathway& pw : t.pathways) n_rules = std::max<size_t>(n_rules, pw.rxn_rule_id + 1);
  for (const mcx_species& sp : t.species) has_surface_species = has_surface_species || !(sp.flags & MCX_SP_VOL);
  if (!t.wall_cv_front.empty()) {
    check(mcx_set_counted_volumes(h, t.n_counted_volumes, t.wall_cv_front.data(), t.wall_cv_back.data()), "mcx_set_counted_volumes");
    n_cv = t.n_counted_volumes;
    if (!t.cv_object_mask.empty()) {
      check(mcx_set_counted_volume_objects(h, t.cv_object_mask.data(), t.intersecting_objects), "mcx_set_counted_volume_objects");
      has_cv_masks = true;
    }
  }
  if (!t.wall_region_set.empty()) {
    check(mcx_set_surface_regions(h, t.n_region_sets, t.wall_region_set.data()), "mcx_set_surface_regions");
    n_rs = t.n_region_sets;
  }
}

GpuDiffuseReactEvent::~GpuDiffuseReactEvent() { mcx_destroy(h); }

void GpuDiffuseReactEvent::set_barrier_time_for_next_execution(const double t) {
  // "Diffusion must advance even if a little bit" / "expected to be a whole number" (diffuse_react_event.h:142-150)
  if (!(t > 0)) throw McxFatalError(MCX_ERR_INVALID_ARG, "time up to the next barrier must be positive");
  if (std::fabs(t - std::round(t)) >= 1e-12) throw McxFatalError(MCX_ERR_INVALID_ARG, "time up to the next barrier must be a whole number");
  time_up_to_next_barrier = t;
}

void GpuDiffuseReactEvent::mark_host_modified() {
  if (device_dirty)
    throw McxFatalError(MCX_ERR_STATE, "mark_host_modified(): the device holds newer molecule state than the host container; "
                                       "call sync_to_host() before editing molecules on the host");
  host_dirty = true;
}

void GpuDiffuseReactEvent::upload_from_host() {
  const size_t n = p->molecules.size();
  x.resize(n); y.resize(n); z.resize(n); tdiff.resize(n); tuni.resize(n); id.resize(n); species.resize(n); flags.resize(n);
  cvi.resize(n);
  if (has_surface_species) { su.resize(n); sv.resize(n); swall.resize(n); stile.resize(n); sorient.resize(n); }
  size_t k = 0;
  for (const Molecule& m : p->molecules) {
    if (m.is_defunct()) continue;
    if (!m.is_vol() && !has_surface_species)
      throw McxFatalError(MCX_ERR_INVALID_ARG, "surface molecule of a model without surface species");
    id[k] = m.id; species[k] = m.species_id;
    if (m.is_vol()) {
      x[k] = m.v.pos.x; y[k] = m.v.pos.y; z[k] = m.v.pos.z;
      cvi[k] = m.v.counted_volume_index == INDEX_INVALID32 ? 0u : m.v.counted_volume_index;
      if (has_surface_species) { swall[k] = MCX_NONE; stile[k] = MCX_NONE; sorient[k] = 0; su[k] = sv[k] = 0; }
    } else {  // Molecule::s; the device derives the 3-D position from (wall, uv) like Partition::add_surface_molecule
      x[k] = y[k] = z[k] = 0; cvi[k] = 0;
      swall[k] = m.s.wall_index; stile[k] = m.s.grid_tile_index; sorient[k] = m.s.orientation; su[k] = m.s.pos.u; sv[k] = m.s.pos.v;
    }
    uint32_t f = 0;
    if (m.flags & MOLECULE_FLAG_SCHEDULE_UNIMOL_RXN) f |= MCX_MOL_SCHEDULE_UNIMOL;
    // COUNTED_VOLUME_INDEX_INVALID: Partition::add_volume_molecule computes it (partition.h:572-576) — here the device does,
    // by a ray cast when the molecule is first evaluated
    if (m.is_vol() && has_cv_masks && m.v.counted_volume_index == INDEX_INVALID32) f |= MCX_MOL_CVI_PENDING;
    const bool partial = m.diffusion_time != TIME_INVALID && m.diffusion_time > event_time + 1e-12;
    if (partial) f |= MCX_MOL_PARTIAL;
    flags[k] = f;
    tdiff[k] = partial ? m.diffusion_time : event_time;
    tuni[k] = m.unimol_rxn_time;  // TIME_INVALID = not drawn yet, TIME_FOREVER = never (same sentinels in the ABI)
    k++;
  }
  mcx_mol_soa v{};
  v.n = k; v.x = x.data(); v.y = y.data(); v.z = z.data(); v.id = id.data(); v.species = species.data(); v.flags = flags.data();
  v.diffusion_time = tdiff.data(); v.unimol_rxn_time = tuni.data();
  if (n_cv > 1) v.counted_volume = cvi.data();
  if (has_surface_species) { v.wall = swall.data(); v.tile = stile.data(); v.orientation = sorient.data(); v.u = su.data(); v.v = sv.data(); }
  check(mcx_upload_molecules(h, &v), "mcx_upload_molecules");
  check(mcx_set_next_molecule_id(h, p->next_molecule_id), "mcx_set_next_molecule_id");   // e.g. restored from a checkpoint
  if (has_surface_species && n_walls && p->wall_has_grid.size() == n_walls)              // likewise: walls keep their grids
    check(mcx_set_wall_grids(h, p->wall_has_grid.data(), n_walls), "mcx_set_wall_grids");
  host_dirty = false;
}

void GpuDiffuseReactEvent::step() {
  if (host_dirty) upload_from_host();
  const double window = std::min(time_up_to_next_barrier, DIFFUSION_TIME_UPPER_LIMIT);
  const uint32_t n_it = (uint32_t)std::max(1.0, std::round(window));
  check(mcx_step(h, n_it, &stats), "mcx_step");
  iterations_last_step = n_it;
  device_dirty = true;
}

void GpuDiffuseReactEvent::sync_to_host() {
  if (!device_dirty) return;
  const size_t cap = (size_t)mcx_num_molecules(h) + 16;
  x.resize(cap); y.resize(cap); z.resize(cap); tdiff.resize(cap); tuni.resize(cap); id.resize(cap); species.resize(cap); flags.resize(cap);
  cvi.resize(cap);
  if (has_surface_species) { su.resize(cap); sv.resize(cap); swall.resize(cap); stile.resize(cap); sorient.resize(cap); }
  mcx_mol_soa v{};
  v.x = x.data(); v.y = y.data(); v.z = z.data(); v.id = id.data(); v.species = species.data(); v.flags = flags.data();
  v.diffusion_time = tdiff.data(); v.unimol_rxn_time = tuni.data();
  if (n_cv > 1) v.counted_volume = cvi.data();
  if (has_surface_species) { v.wall = swall.data(); v.tile = stile.data(); v.orientation = sorient.data(); v.u = su.data(); v.v = sv.data(); }
  check(mcx_download_molecules(h, &v, cap), "mcx_download_molecules");
  std::vector<Molecule> keep;
  keep.reserve(v.n);
  for (uint64_t k = 0; k < v.n; k++) {
    Molecule m(id[k], species[k], Vec3{x[k], y[k], z[k]}, TIME_INVALID);
    if (has_surface_species && swall[k] != MCX_NONE) {  // Molecule::s
      m.flags = MOLECULE_FLAG_SURF;
      m.s.pos = Vec2{su[k], sv[k]}; m.s.orientation = sorient[k]; m.s.wall_index = swall[k]; m.s.grid_tile_index = stile[k];
    } else if (n_cv > 1) m.v.counted_volume_index = cvi[k];
    if (flags[k] & MCX_MOL_SCHEDULE_UNIMOL) m.flags |= MOLECULE_FLAG_SCHEDULE_UNIMOL_RXN;
    m.diffusion_time = tdiff[k];
    m.unimol_rxn_time = tuni[k];
    keep.push_back(m);
  }
  p->molecules.swap(keep);
  p->rebuild_mapping();
  {  // Partition::next_molecule_id follows the device (products took fresh ids there); a checkpoint stores it
    uint32_t next = 0;
    check(mcx_get_next_molecule_id(h, &next), "mcx_get_next_molecule_id");
    if (next > p->next_molecule_id) p->next_molecule_id = next;
  }
  if (has_surface_species && n_walls) {  // Wall::has_initialized_grid follows the device too
    p->wall_has_grid.resize(n_walls);
    check(mcx_get_wall_grids(h, p->wall_has_grid.data(), n_walls), "mcx_get_wall_grids");
  }
  device_dirty = false;
}

molecule_id_t GpuDiffuseReactEvent::release_volume_molecules(species_id_t species, uint64_t number, uint32_t shape,
                                                             const Vec3& location, const Vec3& diameter, double release_time,
                                                             uint32_t counted_volume_index, uint32_t region_in, uint32_t region_out,
                                                             const std::vector<uint8_t>* region_expr) {
  if (host_dirty) upload_from_host();   // the device must hold the current population before molecules are added to it
  mcx_release r{};
  r.species = species; r.shape = shape; r.number = number;
  r.location[0] = location.x; r.location[1] = location.y; r.location[2] = location.z;
  r.diameter[0] = diameter.x; r.diameter[1] = diameter.y; r.diameter[2] = diameter.z;
  r.release_time = release_time; r.counted_volume_index = counted_volume_index;
  r.region_in = region_in; r.region_out = region_out;
  if (region_expr) {
    if (region_expr->size() > sizeof(r.region_expr)) throw McxFatalError(MCX_ERR_INVALID_ARG, "region expression longer than 28 bytes");
    r.region_expr_len = (uint32_t)region_expr->size();
    memcpy(r.region_expr, region_expr->data(), region_expr->size());
  }
  uint32_t first = 0;
  check(mcx_release_volume_molecules(h, &r, &first), "mcx_release_volume_molecules");
  device_dirty = true;
  p->next_molecule_id = first + (molecule_id_t)number;
  return first;
}

molecule_id_t GpuDiffuseReactEvent::release_list(const std::vector<species_id_t>& sp, const std::vector<Vec3>& positions,
                                                 const std::vector<uint32_t>* counted_volume, double release_time) {
  if (sp.size() != positions.size() || (counted_volume && counted_volume->size() != sp.size()))
    throw McxFatalError(MCX_ERR_INVALID_ARG, "release_list: array lengths differ");
  if (host_dirty) upload_from_host();
  const size_t n = sp.size();
  std::vector<double> lx(n), ly(n), lz(n);
  for (size_t k = 0; k < n; k++) { lx[k] = positions[k].x; ly[k] = positions[k].y; lz[k] = positions[k].z; }
  uint32_t first = 0;
  check(mcx_release_list(h, n, sp.data(), lx.data(), ly.data(), lz.data(), counted_volume ? counted_volume->data() : nullptr,
                         release_time, &first), "mcx_release_list");
  device_dirty = true;
  p->next_molecule_id = first + (molecule_id_t)n;
  return first;
}

molecule_id_t GpuDiffuseReactEvent::release_surface_molecules(species_id_t species, uint64_t number, const std::vector<wall_index_t>& walls,
                                                              int32_t orientation, double release_time, bool randomize_pos) {
  if (host_dirty) upload_from_host();
  mcx_surface_release r{};
  r.species = species; r.orientation = orientation; r.number = number; r.release_time = release_time;
  r.walls = walls.data(); r.n_walls = walls.size(); r.randomize_pos = randomize_pos ? 1u : 0u;
  uint32_t first = 0;
  check(mcx_release_surface_molecules(h, &r, &first), "mcx_release_surface_molecules");
  device_dirty = true;
  p->next_molecule_id = first + (molecule_id_t)number;
  return first;
}

void GpuDiffuseReactEvent::get_counts(std::vector<uint64_t>& per_species, std::vector<uint64_t>& per_rxn_rule) {
  per_species.assign(n_species, 0);
  per_rxn_rule.assign(n_rules, 0);
  check(mcx_counts(h, per_species.data(), (uint32_t)n_species, per_rxn_rule.empty() ? nullptr : per_rxn_rule.data(),
                   (uint32_t)n_rules), "mcx_counts");
}

void GpuDiffuseReactEvent::get_counts_by_volume(std::vector<uint64_t>& per_species, std::vector<uint64_t>& per_rxn_rule) {
  per_species.assign(n_species * n_cv, 0);
  per_rxn_rule.assign(std::max<size_t>(n_rules, 1) * n_cv, 0);
  check(mcx_counts_by_volume(h, per_species.data(), per_rxn_rule.data()), "mcx_counts_by_volume");
}

void GpuDiffuseReactEvent::get_counts_by_surface_region(std::vector<uint64_t>& per_species, std::vector<uint64_t>& per_rxn_rule) {
  per_species.assign(n_species * n_rs, 0);
  per_rxn_rule.assign(std::max<size_t>(n_rules, 1) * n_rs, 0);
  check(mcx_counts_by_surface_region(h, per_species.data(), per_rxn_rule.data()), "mcx_counts_by_surface_region");
}

// ---- CountBuffer --------------------------------------------------------------------------------------------
std::string CountBuffer::format_dat_row(const CountItem& item) {
  char buf[96];
  snprintf(buf, sizeof(buf), "%g %g\n", item.time, item.value);  // == operator<<(double) with default precision 6
  return buf;
}

std::string CountBuffer::format_gdat_value(double d) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.8e", d);  // 14-wide column: d.dddddddde+XX, exponent padded to two digits by printf
  return buf;
}

std::string CountBuffer::format_gdat_header() const {
  const size_t width = 14;
  std::string line = "#";
  line += std::string(width - 4, ' ') + "time";
  for (const std::string& name : column_names) {
    if (name.size() < width) line += std::string(width - name.size() + 2, ' ') + name;
    else line += " " + name;
  }
  line += "\n";
  return line;
}

bool CountBuffer::open() {
  if (fout) return true;
  fout = fopen(filename.c_str(), append ? "a" : "w");
  if (!fout) return false;
  if (output_format == CountOutputFormat::GDAT && !append) fputs(format_gdat_header().c_str(), (FILE*)fout);
  return true;
}

void CountBuffer::add(size_t column_index, const CountItem& item) {
  columns.at(column_index).push_back(item);
  if (columns[column_index].size() >= buffer_size && column_index + 1 == columns.size()) flush();
}

void CountBuffer::flush() {
  if (!open()) throw McxFatalError(MCX_ERR_STATE, "Could not open file " + filename + " for writing.");
  FILE* f = (FILE*)fout;
  if (output_format == CountOutputFormat::DAT) {
    for (const CountItem& it : columns[0]) fputs(format_dat_row(it).c_str(), f);
  } else {
    const size_t rows = columns[0].size();
    for (size_t r = 0; r < rows; r++) {
      std::string line = " " + format_gdat_value(columns[0][r].time);
      for (const auto& col : columns) line += "  " + format_gdat_value(col.at(r).value);
      line += "\n";
      fputs(line.c_str(), f);
    }
  }
  fflush(f);
  for (auto& col : columns) col.clear();
}

void CountBuffer::flush_and_close() {
  if (!columns.empty() && !columns[0].empty()) flush();
  if (fout) { fclose((FILE*)fout); fout = nullptr; }
}

void GpuMolOrRxnCountEvent::step() {
  std::vector<uint64_t> per_species, per_rxn, cv_species, cv_rxn, rs_species, rs_rxn;
  diffuse->get_counts(per_species, per_rxn);
  bool need_cv = false, need_rs = false;
  for (const MolOrRxnCountItem& item : items)
    for (const MolOrRxnCountTerm& t : item.terms) {
      need_cv = need_cv || t.where == CountWhere::VolumeRegion;
      need_rs = need_rs || t.where == CountWhere::SurfaceRegion;
    }
  if (need_cv) diffuse->get_counts_by_volume(cv_species, cv_rxn);
  if (need_rs) diffuse->get_counts_by_surface_region(rs_species, rs_rxn);
  for (const MolOrRxnCountItem& item : items) {
    double v = 0;
    for (const MolOrRxnCountTerm& t : item.terms) {
      if (t.where == CountWhere::World) {
        const std::vector<uint64_t>& src = t.is_rxn ? per_rxn : per_species;
        if (t.index < src.size()) v += t.multiplier * (double)src[t.index];
        continue;
      }
      // a term restricted to a region: sum over the counted volumes / region sets its expression holds for
      const bool vol = t.where == CountWhere::VolumeRegion;
      const std::vector<uint64_t>& src = vol ? (t.is_rxn ? cv_rxn : cv_species) : (t.is_rxn ? rs_rxn : rs_species);
      const size_t n_sets = vol ? diffuse->num_counted_volumes() : diffuse->num_region_sets();
      for (uint32_t set : t.sets) {
        const size_t at = (size_t)t.index * n_sets + set;
        if (set < n_sets && at < src.size()) v += t.multiplier * (double)src[at];
      }
    }
    buffers.at(item.buffer)->add(item.column, CountItem{event_time * time_unit, v});
  }
}

// ---- VizOutputWriter / GpuVizOutputEvent (src4/viz_output_event.cpp) ---------------------------------------------
std::string VizOutputWriter::iterations_to_string(uint64_t current_iteration, uint64_t total_iterations) {
  // :65-76 — as many digits as total_iterations has, zero padded
  uint64_t lli = 10;
  int ndigits;
  for (ndigits = 1; lli <= total_iterations && ndigits < 20; ndigits++) lli *= 10;
  char buf[48];
  snprintf(buf, sizeof(buf), "%0*llu", ndigits, (unsigned long long)current_iteration);
  return buf;
}

std::string VizOutputWriter::file_name(const std::string& prefix, viz_mode_t mode, uint64_t current_iteration,
                                       uint64_t total_iterations) {
  return prefix + "." + (mode == ASCII_MODE ? "ascii" : "cellbin") + "." +
         iterations_to_string(current_iteration, total_iterations) + ".dat";  // :79-101
}

void VizOutputWriter::compute_where_and_norm(const Molecule& m, const VizSpeciesInfo& sp, const std::vector<VizWallFrame>* walls,
                                             double length_unit, Vec3& where, Vec3& norm) {
  if (!sp.is_surf) {
    where = m.v.pos;
    norm = Vec3{0, 0, 0};
  } else {
    if (!walls || m.s.wall_index >= walls->size())
      throw McxFatalError(MCX_ERR_INVALID_ARG, "viz output: surface molecule on an unknown wall");
    const VizWallFrame& w = (*walls)[m.s.wall_index];
    // GeometryUtils::uv2xyz (geometry_utils.h:29-31): u * unit_u + v * unit_v + vert0, in this order
    where = Vec3{m.s.pos.u * w.unit_u.x + m.s.pos.v * w.unit_v.x + w.v0.x,
                 m.s.pos.u * w.unit_u.y + m.s.pos.v * w.unit_v.y + w.v0.y,
                 m.s.pos.u * w.unit_u.z + m.s.pos.v * w.unit_v.z + w.v0.z};
    const double o = (double)m.s.orientation;
    norm = Vec3{o * w.normal.x, o * w.normal.y, o * w.normal.z};
  }
  where.x *= length_unit; where.y *= length_unit; where.z *= length_unit;
}

bool VizOutputWriter::write(const std::string& path, viz_mode_t mode, const std::vector<Molecule>& molecules,
                            const std::vector<VizSpeciesInfo>& species, const std::vector<VizWallFrame>* walls, double length_unit,
                            const std::vector<species_id_t>& species_to_visualize) {
  if (mode == NO_VIZ_MODE) return true;
  std::vector<char> shown(species.size(), species_to_visualize.empty() ? 1 : 0);
  for (species_id_t s : species_to_visualize) if (s < shown.size()) shown[s] = 1;
  FILE* f = fopen(path.c_str(), mode == ASCII_MODE ? "w" : "wb");
  if (!f) return false;
  Vec3 where, norm;
  if (mode == ASCII_MODE) {  // output_ascii_molecules :132-170
    for (const Molecule& m : molecules) {
      if (m.is_defunct() || m.species_id >= species.size() || !shown[m.species_id]) continue;
      compute_where_and_norm(m, species[m.species_id], walls, length_unit, where, norm);
      fprintf(f, "%s %u %.9g %.9g %.9g %.9g %.9g %.9g\n", species[m.species_id].name.c_str(), m.id, where.x, where.y, where.z,
              norm.x, norm.y, norm.z);
    }
  } else {  // output_cellblender_molecules :173-265
    std::vector<std::vector<const Molecule*>> by_species(species.size());
    for (const Molecule& m : molecules) {
      if (m.is_defunct() || m.species_id >= species.size() || !shown[m.species_id]) continue;
      by_species[m.species_id].push_back(&m);
    }
    const uint32_t ver = mode == CELLBLENDER_MODE_V2 ? 2 : 1;
    fwrite(&ver, sizeof(uint32_t), 1, f);
    std::vector<float> pos, nrm;
    for (size_t si = 0; si < species.size(); si++) {
      const std::vector<const Molecule*>& mols = by_species[si];
      if (mols.empty()) continue;
      const std::string& name = species[si].name;
      if (ver == 1) { const unsigned char len = (unsigned char)name.size(); fwrite(&len, 1, 1, f); }
      else { const uint32_t len = (uint32_t)name.size(); fwrite(&len, sizeof(uint32_t), 1, f); }
      fwrite(name.data(), 1, name.size(), f);
      const unsigned char type = species[si].is_surf ? 1 : 0;
      fwrite(&type, 1, 1, f);
      const uint32_t count = (uint32_t)(ver == 1 ? 3 * mols.size() : mols.size());
      fwrite(&count, sizeof(uint32_t), 1, f);
      if (ver == 2) for (const Molecule* m : mols) fwrite(&m->id, sizeof(uint32_t), 1, f);
      pos.clear(); nrm.clear();
      for (const Molecule* m : mols) {
        compute_where_and_norm(*m, species[si], walls, length_unit, where, norm);
        pos.push_back((float)where.x); pos.push_back((float)where.y); pos.push_back((float)where.z);
        if (species[si].is_surf) { nrm.push_back((float)norm.x); nrm.push_back((float)norm.y); nrm.push_back((float)norm.z); }
      }
      fwrite(pos.data(), sizeof(float), pos.size(), f);
      fwrite(nrm.data(), sizeof(float), nrm.size(), f);
    }
  }
  const bool ok = ferror(f) == 0;
  return fclose(f) == 0 && ok;
}

void GpuVizOutputEvent::step() {
  if (viz_mode == NO_VIZ_MODE) return;
  if (diffuse) diffuse->sync_to_host();
  const uint64_t it = (uint64_t)std::llround(event_time);
  last_file = VizOutputWriter::file_name(file_prefix_name, viz_mode, it, total_iterations);
  if (!VizOutputWriter::write(last_file, viz_mode, p->molecules, species, walls.empty() ? nullptr : &walls, length_unit,
                              species_ids_to_visualize))
    throw McxFatalError(MCX_ERR_STATE, "Could not write viz output file " + last_file);
}

}  // namespace MCell
