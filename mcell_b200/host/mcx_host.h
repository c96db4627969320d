// mcx_host.h — C++ host side of libmcx: the reference-shaped adapter above the C ABI (include/mcx.h).
//
// It mirrors the part of MCell4's engine interface that DiffuseReactEvent lives behind, with the reference's
// names, argument meaning and error behaviour, so that a maintainer can drop GpuDiffuseReactEvent into
// World::init_simulation (src4/world.cpp:279-282) in place of DiffuseReactEvent:
//
//   MCell::BaseEvent                src4/base_event.h:63-139    (same virtuals, fields, type_index ordering)
//   MCell::Molecule                 src4/molecule.h:52-260      (80-byte AoS record, flag values frozen :30-47)
//   MCell::PartitionMolecules       src4/partition.h:648-666,1125-1143  (molecules vector + id->index map)
//   MCell::GpuDiffuseReactEvent     src4/diffuse_react_event.h:108-154  (step, barrier contract, type_index 500)
//
// The adapter owns no arithmetic: it marshals AoS Molecule records <-> the ABI's SoA view and calls mcx_*.
// The device owns the population between calls; the host vector is a cache with a dirty flag in each direction.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mcx.h"

namespace MCell {

typedef uint32_t molecule_id_t;
typedef uint32_t species_id_t;
typedef uint32_t subpart_index_t;
typedef uint32_t wall_index_t;
typedef uint32_t counted_volume_index_t;
typedef uint16_t event_type_index_t;

const double TIME_INVALID = -256;        // src4/defines.h:178
const double TIME_FOREVER = 1e20;        // src4/defines.h:179
const molecule_id_t MOLECULE_ID_INVALID = 0xFFFFFFFFu;
const uint32_t INDEX_INVALID32 = 0xFFFFFFFFu;
const event_type_index_t EVENT_TYPE_INDEX_DIFFUSE_REACT = 500;  // src4/base_event.h:33-56
const double DIFFUSE_REACT_EVENT_PERIODICITY = 1.0;             // src4/diffuse_react_event.h:36
const double DIFFUSION_TIME_UPPER_LIMIT = 100.0;                // src4/diffuse_react_event.h:37

// src4/molecule.h:30-47 — values are frozen (checkpoints store them)
enum molecule_flag_t {
  MOLECULE_FLAG_SURF = 1 << 0,
  MOLECULE_FLAG_VOL = 1 << 1,
  MOLECULE_FLAG_MATURE = 1 << 2,
  MOLECULE_FLAG_SCHEDULE_UNIMOL_RXN = 1 << 4,
  MOLECULE_FLAG_NO_NEED_TO_SCHEDULE = 1 << 14,
  MOLECULE_FLAG_DEFUNCT = 1 << 15,
};

struct Vec3 { double x, y, z; };
struct Vec2 { double u, v; };

// src4/molecule.h:52-260: "data is ordered to avoid alignment holes"
struct Molecule {
  molecule_id_t id;
  species_id_t species_id;
  uint32_t flags;
  uint32_t pad_;
  double diffusion_time;
  double unimol_rxn_time;
  double birthday;
  union {
    struct {
      Vec3 pos;
      subpart_index_t subpart_index;
      subpart_index_t reactant_subpart_index;
      counted_volume_index_t counted_volume_index;
      wall_index_t previous_wall_index;
    } v;
    struct {
      Vec2 pos;
      int32_t orientation;
      wall_index_t wall_index;
      uint32_t grid_tile_index;
    } s;
  };
  Molecule() { memset((void*)this, 0, sizeof(*this)); id = MOLECULE_ID_INVALID; diffusion_time = TIME_INVALID; unimol_rxn_time = TIME_FOREVER; birthday = TIME_INVALID; }
  Molecule(molecule_id_t id_, species_id_t sp, const Vec3& pos_, double birthday_) {
    memset((void*)this, 0, sizeof(*this));
    id = id_; species_id = sp; flags = MOLECULE_FLAG_VOL;
    diffusion_time = TIME_INVALID; unimol_rxn_time = TIME_INVALID; birthday = birthday_;
    v.pos = pos_; v.subpart_index = v.reactant_subpart_index = v.counted_volume_index = v.previous_wall_index = INDEX_INVALID32;
  }
  bool is_vol() const { return (flags & MOLECULE_FLAG_VOL) != 0; }
  bool is_defunct() const { return (flags & MOLECULE_FLAG_DEFUNCT) != 0; }
};
static_assert(sizeof(Molecule) == 80, "sizeof(Molecule) must stay 80 bytes (SURVEY A.5)");

// src4/base_event.h:63-139
class BaseEvent {
public:
  explicit BaseEvent(event_type_index_t t) : event_time(TIME_INVALID), periodicity_interval(0), type_index(t) {}
  virtual ~BaseEvent() {}
  virtual void step() = 0;
  virtual bool update_event_time_for_next_scheduled_time() {
    if (periodicity_interval == 0) return false;
    event_time = event_time + periodicity_interval;
    return true;
  }
  virtual bool is_barrier() const { return false; }
  virtual bool may_be_blocked_by_barrier_and_needs_set_time_step() const { return false; }
  virtual double get_max_time_up_to_next_barrier() const { return 0; }
  virtual void set_barrier_time_for_next_execution(const double) {}
  double event_time;
  double periodicity_interval;
  event_type_index_t type_index;
};

// The molecule containers of Partition that the event reads and writes (partition.h:1125-1143)
struct PartitionMolecules {
  std::vector<Molecule> molecules;
  std::vector<uint32_t> molecule_id_to_index_mapping;
  molecule_id_t next_molecule_id = 0;
  // Wall::has_initialized_grid of every wall (src4/wall.h:339-346), one byte each; empty = not tracked.  Part of a checkpoint
  // of a model with surface-surface reactions: sync_to_host() fills it, the adapter's upload restores it
  std::vector<uint8_t> wall_has_grid;

  // Partition::add_volume_molecule (partition.h:555-611): assigns the id, marks a newborn
  Molecule& add_volume_molecule(species_id_t species, const Vec3& pos, double birthday);
  // Partition::get_m (partition.h:210-219)
  Molecule& get_m(molecule_id_t id) { return molecules[molecule_id_to_index_mapping[id]]; }
  void rebuild_mapping();
};

// World::fatal_error equivalent for this seam: the ABI never exits; the adapter throws and the host decides
struct McxFatalError : std::runtime_error {
  int code;
  McxFatalError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// Immutable model tables in ABI form (filled by the host from World/BNGEngine/Partition)
struct GpuModelTables {
  mcx_config cfg{};
  std::vector<mcx_species> species;
  std::vector<mcx_rxn_class> rxn_classes;
  std::vector<mcx_pathway> pathways;
  std::vector<mcx_surf_class_rxn> surf_class_rxns;
  std::vector<double> vertices;        // 3 per vertex, length units
  std::vector<uint32_t> wall_vertex_indices;  // 3 per wall
  std::vector<uint32_t> wall_surf_class;      // per wall or empty
  // counted volumes (World::init_counted_volumes): per wall the counted-volume index in front of / behind it; empty = none
  uint32_t n_counted_volumes = 1;
  std::vector<uint8_t> wall_cv_front, wall_cv_back;
  // counted objects that intersect (Partition waypoints): per counted volume the set of enclosing objects (bit k = object
  // k), and the objects whose walls toggle membership instead of naming a pair of volumes; empty = none intersect
  std::vector<uint32_t> cv_object_mask;
  uint32_t intersecting_objects = 0;
  // counted surface regions: per wall the index of the set of counted regions it belongs to (set 0 = none); empty = none
  uint32_t n_region_sets = 1;
  std::vector<uint8_t> wall_region_set;
};

// Drop-in for DiffuseReactEvent (src4/diffuse_react_event.h:108-154) running on the GPU through libmcx.
class GpuDiffuseReactEvent : public BaseEvent {
public:
  GpuDiffuseReactEvent(const GpuModelTables& tables, PartitionMolecules* partition);
  ~GpuDiffuseReactEvent() override;

  // one scheduler slot: runs round(time_up_to_next_barrier) iterations on the device (the barrier contract of
  // diffuse_react_event.h:142-150 passes whole numbers <= DIFFUSION_TIME_UPPER_LIMIT)
  void step() override;
  bool update_event_time_for_next_scheduled_time() override {
    event_time = event_time + iterations_last_step;  // = min(time_up_to_next_barrier, window) whole iterations
    return true;
  }
  bool may_be_blocked_by_barrier_and_needs_set_time_step() const override { return true; }
  double get_max_time_up_to_next_barrier() const override { return DIFFUSION_TIME_UPPER_LIMIT; }
  void set_barrier_time_for_next_execution(const double t) override;

  // host <-> device population sync.  The two dirty flags are mutually exclusive: after a step() the device holds
  // the newer state, so a host-side edit (ReleaseEvent::step with a region / list / surface release, Model API
  // edits) goes  sync_to_host(); <edit Partition>; mark_host_modified();  — mark_host_modified() on a stale host
  // container throws instead of silently replacing the device population with it.  viz / checkpoint / introspection
  // paths call sync_to_host() before reading Partition::get_molecules().
  void mark_host_modified();
  void sync_to_host();
  // ReleaseEvent::release_ellipsoid_or_rectcuboid on the device (release_event.cpp:953-1003):
  // `number` molecules of a volume species in a cuboid / sphere / spherical shell (MCX_RELEASE_*) of the given centre
  // and diameter (internal length units) at event_time; returns the first of the consecutive new ids.  The host
  // container is stale afterwards until sync_to_host().
  molecule_id_t release_volume_molecules(species_id_t species, uint64_t number, uint32_t shape, const Vec3& location,
                                         const Vec3& diameter, double release_time = 0, uint32_t counted_volume_index = 0,
                                         uint32_t region_in = 0, uint32_t region_out = 0,
                                         const std::vector<uint8_t>* region_expr = nullptr);
  // ReleaseEvent::release_list (release_event.cpp:1008-1040) for volume molecules, on the device: one molecule of
  // species[k] at positions[k] (internal length units); returns the first of the consecutive new ids
  molecule_id_t release_list(const std::vector<species_id_t>& species, const std::vector<Vec3>& positions,
                             const std::vector<uint32_t>* counted_volume = nullptr, double release_time = 0);
  // ReleaseEvent::release_onto_regions (release_event.cpp:640-760) on the device: `number` molecules of a surface
  // species on vacant tiles of the listed walls (the walls of the release's surface regions); orientation 0 = random
  molecule_id_t release_surface_molecules(species_id_t species, uint64_t number, const std::vector<wall_index_t>& walls,
                                          int32_t orientation, double release_time = 0, bool randomize_pos = true);
  // MolOrRxnCountEvent world-count fast path (mol_or_rxn_count_event.cpp:622-653)
  void get_counts(std::vector<uint64_t>& per_species, std::vector<uint64_t>& per_rxn_rule);
  // count terms restricted to a volume / a surface region (mol_or_rxn_count_event.cpp:519-534, 571-600):
  // per_species[species * n_sets + set], per_rxn_rule[rule * n_sets + set]
  void get_counts_by_volume(std::vector<uint64_t>& per_species, std::vector<uint64_t>& per_rxn_rule);
  void get_counts_by_surface_region(std::vector<uint64_t>& per_species, std::vector<uint64_t>& per_rxn_rule);
  uint32_t num_counted_volumes() const { return n_cv; }
  uint32_t num_region_sets() const { return n_rs; }
  const mcx_step_stats& last_stats() const { return stats; }

private:
  void upload_from_host();
  void check(int rc, const char* what);
  mcx_handle* h = nullptr;
  PartitionMolecules* p;
  size_t n_species, n_rules;
  uint32_t n_cv = 1, n_rs = 1;
  bool has_cv_masks = false;   // the device can compute a counted volume by a ray cast (MCX_MOL_CVI_PENDING)
  bool has_surface_species = false;
  uint64_t n_walls = 0;
  double time_up_to_next_barrier;
  double iterations_last_step = 1;
  bool host_dirty = true, device_dirty = false;
  mcx_step_stats stats{};
  // SoA staging (host side of the ABI)
  std::vector<double> x, y, z, tdiff, tuni;
  std::vector<uint32_t> id, species, flags;
  // Molecule::s and v.counted_volume_index
  std::vector<double> su, sv;
  std::vector<uint32_t> swall, stile, cvi;
  std::vector<int32_t> sorient;
};

// ---- observables output: CountBuffer / CountItem (src4/count_buffer.h:24-123, count_buffer.cpp:30-129) -----------
// Rows of (time, value) per column, buffered and flushed as text that is byte-identical to the reference's
// react_data files: ".dat" = one column, "time value" per line in default stream formatting (%g, 6 significant
// digits); ".gdat" = '#' header with 14-wide right-aligned names, then " t  v1  v2 ..." in scientific notation
// with 8 fractional digits and an exponent of at least two digits.
enum class CountOutputFormat { DAT, GDAT };

struct CountItem {
  double time;    // seconds: iteration * time_unit (mol_or_rxn_count_event.cpp:707-716)
  double value;
};

class CountBuffer {
public:
  CountBuffer(const std::string& filename_, const std::vector<std::string>& column_names_, size_t buffer_size_,
              CountOutputFormat fmt, bool append_ = false)
    : filename(filename_), column_names(column_names_), buffer_size(buffer_size_), output_format(fmt), append(append_),
      columns(column_names_.empty() ? 1 : column_names_.size()) {}
  ~CountBuffer() { flush_and_close(); }
  // CountBuffer::add (count_buffer.h:70-76): flushes when a column holds buffer_size rows
  void add(size_t column_index, const CountItem& item);
  void flush();
  void flush_and_close();
  static std::string format_dat_row(const CountItem& item);
  static std::string format_gdat_value(double d);
  std::string format_gdat_header() const;
private:
  bool open();
  std::string filename;
  std::vector<std::string> column_names;
  size_t buffer_size;
  CountOutputFormat output_format;
  bool append;
  std::vector<std::vector<CountItem>> columns;
  void* fout = nullptr;  // FILE*
};

// The world-count part of MolOrRxnCountEvent (mol_or_rxn_count_event.cpp:622-653): every `periodicity` iterations
// one row per observable; species counts and reaction counts come from the device (mcx_counts), so a count
// iteration costs no molecule download.
// CountType of MolOrRxnCountTerm (mol_or_rxn_count_event.h): where the molecules / reactions of a term are counted
enum class CountWhere { World, VolumeRegion, SurfaceRegion };
// index: species id or rxn rule id.  sets (VolumeRegion: counted-volume indices, SurfaceRegion: region-set indices):
// the sets for which the term's region expression holds, evaluated once by the host
// (counted_volume_matches_region_expr_recursively / wall_matches_region_expr_recursively)
struct MolOrRxnCountTerm {
  bool is_rxn; uint32_t index; double multiplier;
  CountWhere where = CountWhere::World;
  std::vector<uint32_t> sets;
};
struct MolOrRxnCountItem { size_t buffer, column; std::vector<MolOrRxnCountTerm> terms; };

class GpuMolOrRxnCountEvent : public BaseEvent {
public:
  GpuMolOrRxnCountEvent(GpuDiffuseReactEvent* diffuse_, double time_unit_)
    : BaseEvent(290 /* EVENT_TYPE_INDEX_MOL_OR_RXN_COUNT, base_event.h:33-56 */), diffuse(diffuse_), time_unit(time_unit_) {}
  bool is_barrier() const override { return true; }
  void step() override;
  std::vector<CountBuffer*> buffers;
  std::vector<MolOrRxnCountItem> items;
private:
  GpuDiffuseReactEvent* diffuse;
  double time_unit;
};

// ---- releases: the device-capable part of ReleaseEvent (src4/release_event.h, release_event.cpp:953-1003) -------------
// One release of `release_number` molecules of a volume species in a cuboid, sphere or spherical shell at event_time
// (EVENT_TYPE_INDEX_RELEASE = 200, so that it runs before the counts and the diffusion of its iteration, base_event.h:
// 33-56).  release_shape MCX_RELEASE_REGION releases inside closed objects (release_inside_regions, :904-951; location /
// diameter = the region's bounding box, region_in / region_out = object masks); lists go through
// GpuDiffuseReactEvent::release_list.  Surface releases stay with the host's ReleaseEvent and reach the device through
// Partition::add_surface_molecule + mark_host_modified().
class GpuReleaseEvent : public BaseEvent {
public:
  GpuReleaseEvent(GpuDiffuseReactEvent* diffuse_, species_id_t species_id_, uint64_t release_number_, uint32_t shape_,
                  const Vec3& location_, const Vec3& diameter_, uint32_t counted_volume_index_ = 0)
    : BaseEvent(200), species_id(species_id_), release_number(release_number_), release_shape(shape_), location(location_),
      diameter(diameter_), counted_volume_index(counted_volume_index_), diffuse(diffuse_) {}
  bool is_barrier() const override { return true; }   // the diffuse event must stop at the release time
  void step() override {
    first_released_id = diffuse->release_volume_molecules(species_id, release_number, release_shape, location, diameter,
                                                          event_time, counted_volume_index, region_in, region_out,
                                                          region_expr.empty() ? nullptr : &region_expr);
  }
  uint32_t region_in = 0, region_out = 0;   // MCX_RELEASE_REGION: objects the molecules must be inside / outside of
  std::vector<uint8_t> region_expr;         // ... or the release's RegionExprNode tree in postfix (include/mcx.h)
  species_id_t species_id;
  uint64_t release_number;
  uint32_t release_shape;   // MCX_RELEASE_CUBIC / _SPHERICAL / _SPHERICAL_SHELL / _REGION
  Vec3 location, diameter;  // internal length units
  uint32_t counted_volume_index;
  molecule_id_t first_released_id = MOLECULE_ID_INVALID;
private:
  GpuDiffuseReactEvent* diffuse;
};

// ---- visualization output: VizOutputEvent (src4/viz_output_event.cpp:65-280) -------------------------------------
// Molecule dumps in the reference's two formats, byte for byte:
//   ASCII       "<species name> <id> <x> <y> <z> <nx> <ny> <nz>\n" with %.9g doubles (viz_output_event.cpp:132-170),
//               molecules in Partition::get_molecules() order;
//   CELLBLENDER binary: u32 version (1 or 2), then per species with at least one molecule, in species-id order:
//               name length (u8 in v1, u32 in v2), name bytes, u8 species type (1 = surface), u32 count (3 * n floats
//               in v1, n molecules in v2), [v2: n u32 ids], n * 3 float32 positions, and for surface species
//               n * 3 float32 normals (viz_output_event.cpp:173-265).
// Positions are in micrometres (internal position * length_unit); a volume molecule's normal is 0, a surface
// molecule's is orientation * wall normal and its position uv2xyz(s.pos) (compute_where_and_norm :104-129).
// File name: <prefix>.<ascii|cellbin>.<iteration, zero-padded to the digits of total_iterations>.dat (:65-101).
enum viz_mode_t { NO_VIZ_MODE = 0, ASCII_MODE = 1, CELLBLENDER_MODE_V1 = 2, CELLBLENDER_MODE_V2 = 3 };

struct VizSpeciesInfo { std::string name; bool is_surf; };
// per wall: vertex 0, unit_u, unit_v, normal (3 doubles each, internal length units) — what uv2xyz and the normal need
struct VizWallFrame { Vec3 v0, unit_u, unit_v, normal; };

class VizOutputWriter {
public:
  static std::string iterations_to_string(uint64_t current_iteration, uint64_t total_iterations);
  static std::string file_name(const std::string& prefix, viz_mode_t mode, uint64_t current_iteration, uint64_t total_iterations);
  // where (micrometres) and normal of one molecule; walls may be null when the model has no surface molecules
  static void compute_where_and_norm(const Molecule& m, const VizSpeciesInfo& sp, const std::vector<VizWallFrame>* walls,
                                     double length_unit, Vec3& where, Vec3& norm);
  // writes every non-defunct molecule of the listed species (all when species_to_visualize is empty); false on I/O error
  static bool write(const std::string& path, viz_mode_t mode, const std::vector<Molecule>& molecules,
                    const std::vector<VizSpeciesInfo>& species, const std::vector<VizWallFrame>* walls, double length_unit,
                    const std::vector<species_id_t>& species_to_visualize = std::vector<species_id_t>());
};

// Drop-in for VizOutputEvent: a barrier event (the diffuse event stops at it), pulls the population from the device
// only when it fires.
class GpuVizOutputEvent : public BaseEvent {
public:
  GpuVizOutputEvent(GpuDiffuseReactEvent* diffuse_, PartitionMolecules* partition, viz_mode_t mode, const std::string& prefix,
                    uint64_t total_iterations_, double length_unit_)
    : BaseEvent(300 /* EVENT_TYPE_INDEX_VIZ_OUTPUT, base_event.h:33-56 */), viz_mode(mode), file_prefix_name(prefix),
      total_iterations(total_iterations_), length_unit(length_unit_), diffuse(diffuse_), p(partition) {}
  bool is_barrier() const override { return true; }
  void step() override;
  viz_mode_t viz_mode;
  std::string file_prefix_name;
  uint64_t total_iterations;
  double length_unit;
  std::vector<VizSpeciesInfo> species;
  std::vector<VizWallFrame> walls;                  // empty: volume molecules only
  std::vector<species_id_t> species_ids_to_visualize;  // empty: all
  std::string last_file;
private:
  GpuDiffuseReactEvent* diffuse;
  PartitionMolecules* p;
};

}  // namespace MCell
