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
  for (const mcx_pathway& pw : t.pathways) n_rules = std::max<size_t>(n_rules, pw.rxn_rule_id + 1);
}

GpuDiffuseReactEvent::~GpuDiffuseReactEvent() { mcx_destroy(h); }

void GpuDiffuseReactEvent::set_barrier_time_for_next_execution(const double t) {
  // "Diffusion must advance even if a little bit" / "expected to be a whole number" (diffuse_react_event.h:142-150)
  if (!(t > 0)) throw McxFatalError(MCX_ERR_INVALID_ARG, "time up to the next barrier must be positive");
  if (std::fabs(t - std::round(t)) >= 1e-12) throw McxFatalError(MCX_ERR_INVALID_ARG, "time up to the next barrier must be a whole number");
  time_up_to_next_barrier = t;
}

void GpuDiffuseReactEvent::upload_from_host() {
  const size_t n = p->molecules.size();
  x.resize(n); y.resize(n); z.resize(n); tdiff.resize(n); tuni.resize(n); id.resize(n); species.resize(n); flags.resize(n);
  size_t k = 0;
  for (const Molecule& m : p->molecules) {
    if (m.is_defunct() || !m.is_vol()) continue;  // surface molecules: not on the device path yet
    x[k] = m.v.pos.x; y[k] = m.v.pos.y; z[k] = m.v.pos.z; id[k] = m.id; species[k] = m.species_id;
    uint32_t f = 0;
    if (m.flags & MOLECULE_FLAG_SCHEDULE_UNIMOL_RXN) f |= MCX_MOL_SCHEDULE_UNIMOL;
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
  check(mcx_upload_molecules(h, &v), "mcx_upload_molecules");
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
  mcx_mol_soa v{};
  v.x = x.data(); v.y = y.data(); v.z = z.data(); v.id = id.data(); v.species = species.data(); v.flags = flags.data();
  v.diffusion_time = tdiff.data(); v.unimol_rxn_time = tuni.data();
  check(mcx_download_molecules(h, &v, cap), "mcx_download_molecules");
  std::vector<Molecule> keep;
  keep.reserve(v.n);
  for (const Molecule& m : p->molecules)
    if (!m.is_vol() && !m.is_defunct()) keep.push_back(m);  // host-resident (surface) molecules stay as they are
  for (uint64_t k = 0; k < v.n; k++) {
    Molecule m(id[k], species[k], Vec3{x[k], y[k], z[k]}, TIME_INVALID);
    if (flags[k] & MCX_MOL_SCHEDULE_UNIMOL) m.flags |= MOLECULE_FLAG_SCHEDULE_UNIMOL_RXN;
    m.diffusion_time = tdiff[k];
    m.unimol_rxn_time = tuni[k];
    keep.push_back(m);
  }
  p->molecules.swap(keep);
  p->rebuild_mapping();
  device_dirty = false;
}

void GpuDiffuseReactEvent::get_counts(std::vector<uint64_t>& per_species, std::vector<uint64_t>& per_rxn_rule) {
  per_species.assign(n_species, 0);
  per_rxn_rule.assign(n_rules, 0);
  check(mcx_counts(h, per_species.data(), (uint32_t)n_species, per_rxn_rule.empty() ? nullptr : per_rxn_rule.data(),
                   (uint32_t)n_rules), "mcx_counts");
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
  std::vector<uint64_t> per_species, per_rxn;
  diffuse->get_counts(per_species, per_rxn);
  for (const MolOrRxnCountItem& item : items) {
    double v = 0;
    for (const MolOrRxnCountTerm& t : item.terms) {
      const std::vector<uint64_t>& src = t.is_rxn ? per_rxn : per_species;
      if (t.index < src.size()) v += t.multiplier * (double)src[t.index];
    }
    buffers.at(item.buffer)->add(item.column, CountItem{event_time * time_unit, v});
  }
}

}  // namespace MCell
