// tests/host/test_host_adapter.cpp — drives libmcx the way MCell4's scheduler would, through the C++ adapter
// (mcell_b200/host/mcx_host.h): BASELINE config 1 shape (free diffusion in a reflective 1 um cube) plus an
// A + B -> C run, with a count "barrier" every 10 iterations.
// Without a CUDA device it must fail loudly (McxFatalError, MCX_ERR_CUDA) — exit code 3.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../mcell_b200/host/mcx_host.h"

using namespace MCell;

static GpuModelTables box_model(double half_lu, bool reactive, uint64_t max_mols) {
  GpuModelTables t;
  t.cfg.abi_version = MCX_ABI_VERSION;
  t.cfg.device = 0;
  t.cfg.seed = 1;
  t.cfg.origin[0] = t.cfg.origin[1] = t.cfg.origin[2] = -500;
  t.cfg.partition_edge_length = 1000;
  t.cfg.num_subparts_per_edge = 20;
  t.cfg.use_expanded_list = reactive ? 1 : 0;
  t.cfg.rxn_radius_3d = 0.5641895835477563;
  for (int k = 0; k < 3; k++) { t.cfg.active_llf[k] = -half_lu; t.cfg.active_urb[k] = half_lu; }
  t.cfg.max_molecules = max_mols;
  t.cfg.rng_mode = MCX_RNG_PHILOX;
  t.cfg.world_size = 1;
  const double h = half_lu;
  const double v[8][3] = {{-h, -h, -h}, {-h, -h, h}, {-h, h, -h}, {-h, h, h}, {h, -h, -h}, {h, -h, h}, {h, h, -h}, {h, h, h}};
  const uint32_t f[12][3] = {{1, 2, 0}, {3, 6, 2}, {7, 4, 6}, {5, 0, 4}, {6, 0, 2}, {3, 5, 7},
                             {1, 3, 2}, {3, 7, 6}, {7, 5, 4}, {5, 1, 0}, {6, 4, 0}, {3, 1, 5}};
  for (auto& p : v) for (double c : p) t.vertices.push_back(c);
  for (auto& q : f) for (uint32_t c : q) t.wall_vertex_indices.push_back(c);
  const int ns = reactive ? 3 : 1;
  for (int i = 0; i < ns; i++) t.species.push_back(mcx_species{2.0, 1.0, MCX_SP_VOL | MCX_SP_CAN_DIFFUSE, 0});
  if (reactive) {
    mcx_rxn_class rc{};
    rc.kind = MCX_RXN_BIMOL_VOLVOL; rc.reactants[0] = 0; rc.reactants[1] = 1; rc.first_pathway = 0; rc.n_pathways = 1;
    rc.max_fixed_p = 0.3;
    mcx_pathway pw{};
    pw.cum_prob = 0.3; pw.n_products = 1; pw.products[0] = 2; pw.keep_reactant_mask = 0; pw.rxn_rule_id = 0;
    t.rxn_classes.push_back(rc);
    t.pathways.push_back(pw);
  }
  return t;
}

int main() {
  const int n = 20000;
  const double half = 50.0;
  std::mt19937_64 gen(7);
  std::uniform_real_distribution<double> U(-half * 0.999, half * 0.999);
  try {
    // ---- config 1 shape: free diffusion, barrier every 10 iterations
    PartitionMolecules part;
    for (int i = 0; i < n; i++) part.add_volume_molecule(0, Vec3{U(gen), U(gen), U(gen)}, 0.0);
    std::vector<Molecule> before = part.molecules;
    GpuModelTables t = box_model(half, false, 2 * n);
    GpuDiffuseReactEvent ev(t, &part);
    ev.event_time = 0;
    if (ev.type_index != 500 || !ev.may_be_blocked_by_barrier_and_needs_set_time_step()) return 1;
    // one iteration: MSD = 3*space_step^2/2 for molecules away from the walls
    ev.set_barrier_time_for_next_execution(1);
    ev.step();
    ev.update_event_time_for_next_scheduled_time();
    ev.sync_to_host();
    if (part.molecules.size() != (size_t)n) { printf("lost molecules: %zu\n", part.molecules.size()); return 1; }
    double msd = 0; int cnt = 0;
    for (const Molecule& b : before) {
      if (std::fabs(b.v.pos.x) > 30 || std::fabs(b.v.pos.y) > 30 || std::fabs(b.v.pos.z) > 30) continue;
      const Molecule& a = part.get_m(b.id);
      double dx = a.v.pos.x - b.v.pos.x, dy = a.v.pos.y - b.v.pos.y, dz = a.v.pos.z - b.v.pos.z;
      msd += dx * dx + dy * dy + dz * dz; cnt++;
    }
    msd /= cnt;
    if (std::fabs(msd - 6.0) > 6.0 * 5 * std::sqrt(2.0 / 3.0 / cnt)) { printf("MSD %g, expected 6\n", msd); return 1; }
    for (int window = 0; window < 3; window++) {
      ev.set_barrier_time_for_next_execution(10);   // a count event every 10 iterations
      ev.step();
      ev.update_event_time_for_next_scheduled_time();
    }
    if (ev.event_time != 31) { printf("event_time %g\n", ev.event_time); return 1; }
    ev.sync_to_host();
    for (const Molecule& m : part.molecules)
      if (std::fabs(m.v.pos.x) > half || std::fabs(m.v.pos.y) > half || std::fabs(m.v.pos.z) > half) { printf("escaped\n"); return 1; }
    if (ev.last_stats().molecule_steps != (uint64_t)10 * n) return 1;

    // ---- A + B -> C with counts at every barrier
    PartitionMolecules p2;
    for (int i = 0; i < n; i++) p2.add_volume_molecule(i & 1, Vec3{U(gen) * 0.3, U(gen) * 0.3, U(gen) * 0.3}, 0.0);
    GpuModelTables t2 = box_model(half * 0.3, true, 2 * n);
    GpuDiffuseReactEvent ev2(t2, &p2);
    ev2.event_time = 0;
    std::vector<uint64_t> sp, rx;
    uint64_t last_c = 0;
    for (int window = 0; window < 3; window++) {
      ev2.set_barrier_time_for_next_execution(10);
      ev2.step();
      ev2.update_event_time_for_next_scheduled_time();
      ev2.get_counts(sp, rx);
      if (sp[0] + sp[2] != (uint64_t)n / 2 || sp[1] + sp[2] != (uint64_t)n / 2 || rx[0] != sp[2] || sp[2] < last_c) { printf("count identity broken\n"); return 1; }
      last_c = sp[2];
    }
    if (last_c < 100) { printf("too few reactions: %llu\n", (unsigned long long)last_c); return 1; }
    // a host edit on a stale container must be refused, not silently uploaded over the newer device state
    {
      bool refused = false;
      try { ev2.mark_host_modified(); } catch (const McxFatalError& e) { refused = e.code == MCX_ERR_STATE; }
      if (!refused) { printf("mark_host_modified() accepted a stale host container\n"); return 1; }
    }
    ev2.sync_to_host();
    if (p2.molecules.size() != sp[0] + sp[1] + sp[2]) return 1;
    // host edits the population (a "release"), marks it dirty, continues
    const uint64_t rx_before = rx[0], c_before = sp[2];
    const molecule_id_t next_before = p2.next_molecule_id;
    for (int i = 0; i < 100; i++) p2.add_volume_molecule(0, Vec3{0, 0, 0}, ev2.event_time);
    ev2.mark_host_modified();
    ev2.set_barrier_time_for_next_execution(1);
    ev2.step();
    ev2.get_counts(sp, rx);
    if (sp[0] + sp[2] != (uint64_t)n / 2 + 100) { printf("release not seen\n"); return 1; }
    // reaction counts are cumulative over the run (MolOrRxnCountEvent): the re-upload must not restart them
    if (rx[0] < rx_before || rx[0] - rx_before != sp[2] - c_before) {
      printf("reaction count restarted by the re-upload: %llu -> %llu\n", (unsigned long long)rx_before, (unsigned long long)rx[0]); return 1;
    }
    ev2.sync_to_host();
    if (p2.next_molecule_id < next_before + 100) { printf("next id went backwards\n"); return 1; }
    // viz dump at a barrier: the event pulls the population from the device itself (no sync_to_host by the caller)
    {
      GpuVizOutputEvent viz(&ev2, &p2, CELLBLENDER_MODE_V2, "/tmp/mcx_host_adapter_viz", 100, 0.01);
      viz.species = {{"A", false}, {"B", false}, {"C", false}};
      viz.event_time = ev2.event_time;
      viz.step();
      FILE* f = fopen(viz.last_file.c_str(), "rb");
      if (!f) { printf("viz file missing\n"); return 1; }
      uint32_t ver = 0; size_t got = fread(&ver, 4, 1, f);
      uint64_t total = 0;
      for (int k = 0; k < 3 && got == 1; k++) {   // per species: name length, name, type, count, ids, positions
        uint32_t len = 0, cnt = 0; unsigned char type = 9; char name[8] = {0};
        if (fread(&len, 4, 1, f) != 1 || len != 1 || fread(name, 1, 1, f) != 1 || fread(&type, 1, 1, f) != 1 || fread(&cnt, 4, 1, f) != 1) break;
        if (name[0] != "ABC"[k] || type != 0 || cnt != sp[k]) { printf("viz species block %d wrong: %s %u\n", k, name, cnt); return 1; }
        fseek(f, (long)cnt * 16, SEEK_CUR);
        total += cnt;
      }
      const long end = ftell(f); fseek(f, 0, SEEK_END);
      if (ver != 2 || total != sp[0] + sp[1] + sp[2] || end != ftell(f)) { printf("viz file wrong\n"); return 1; }
      fclose(f); remove(viz.last_file.c_str());
    }
    // ---- surface molecules through the adapter, counts restricted to counted volumes and to surface regions:
    // L (volume) + R (surface, on the box faces) -> LR;  LR -> R (initiated by a surface molecule: counted on its wall)
    {
      const double hb = half * 0.3;
      GpuModelTables t3 = box_model(hb, false, 4 * n);
      t3.species.clear();
      t3.species.push_back(mcx_species{2.0, 1.0, MCX_SP_VOL | MCX_SP_CAN_DIFFUSE, 0});  // L
      t3.species.push_back(mcx_species{0.0, 1.0, 0, 0});                                  // R  (surface, static)
      t3.species.push_back(mcx_species{0.0, 1.0, 0, 0});                                  // LR (surface, static)
      mcx_rxn_class bind{}; bind.kind = MCX_RXN_BIMOL_VOLSURF; bind.reactants[0] = 0; bind.reactants[1] = 1;
      bind.first_pathway = 0; bind.n_pathways = 1; bind.max_fixed_p = 0.5;
      mcx_pathway pb{}; pb.cum_prob = 0.5; pb.n_products = 1; pb.products[0] = 2; pb.product_orientation[0] = 1; pb.rxn_rule_id = 0;
      mcx_rxn_class unb{}; unb.kind = MCX_RXN_UNIMOL; unb.reactants[0] = 2; unb.reactants[1] = MCX_NONE;
      unb.first_pathway = 1; unb.n_pathways = 1; unb.max_fixed_p = 0.2;
      mcx_pathway pu{}; pu.cum_prob = 0.2; pu.n_products = 1; pu.products[0] = 1; pu.product_orientation[0] = 1; pu.rxn_rule_id = 1;
      t3.rxn_classes = {bind, unb};
      t3.pathways = {pb, pu};
      // the box is a counted volume (index 1 inside = behind its outward-facing walls); walls 0-5 are region "low",
      // walls 6-11 region "high": region sets {}, {low}, {high}
      t3.n_counted_volumes = 2; t3.wall_cv_front.assign(12, 0); t3.wall_cv_back.assign(12, 1);
      t3.n_region_sets = 3; t3.wall_region_set.resize(12);
      for (int w = 0; w < 12; w++) t3.wall_region_set[w] = w < 6 ? 1 : 2;
      PartitionMolecules p3;
      for (int i = 0; i < n; i++) {
        Molecule& m = p3.add_volume_molecule(0, Vec3{U(gen) * 0.3, U(gen) * 0.3, U(gen) * 0.3}, 0.0);
        m.v.counted_volume_index = 1;
      }
      // receptors: every third tile of every wall, at the tile centre (Partition::add_surface_molecule)
      uint64_t n_rec = 0, n_low = 0;
      for (uint32_t w = 0; w < 12; w++) {
        double v9[9];
        for (int c = 0; c < 3; c++) for (int k = 0; k < 3; k++) v9[3 * c + k] = t3.vertices[3 * t3.wall_vertex_indices[3 * w + c] + k];
        const uint32_t nt = mcx_grid_num_tiles(v9);
        for (uint32_t tile = 0; tile < nt; tile += 3) {
          double uv[2]; mcx_grid2uv(v9, tile, uv);
          Molecule m; m.id = p3.next_molecule_id++; m.species_id = 1; m.flags = MOLECULE_FLAG_SURF | MOLECULE_FLAG_SCHEDULE_UNIMOL_RXN;
          m.diffusion_time = 0; m.unimol_rxn_time = TIME_INVALID; m.birthday = 0;
          m.s.pos = Vec2{uv[0], uv[1]}; m.s.orientation = 1; m.s.wall_index = w; m.s.grid_tile_index = tile;
          p3.molecules.push_back(m);
          n_rec++; if (w < 6) n_low++;
        }
      }
      p3.rebuild_mapping();
      GpuDiffuseReactEvent ev3(t3, &p3);
      ev3.event_time = 0;
      GpuMolOrRxnCountEvent cnt3(&ev3, 1e-6);
      CountBuffer buf("/tmp/mcx_host_adapter_counts.gdat", {"LR", "LR_low", "LR_high", "unbind_low", "unbind_high", "bind_in_box", "L_in_box"},
                      100, CountOutputFormat::GDAT);
      cnt3.buffers = {&buf};
      cnt3.items = {
        {0, 0, {{false, 2, 1.0}}},
        {0, 1, {{false, 2, 1.0, CountWhere::SurfaceRegion, {1}}}},
        {0, 2, {{false, 2, 1.0, CountWhere::SurfaceRegion, {2}}}},
        {0, 3, {{true, 1, 1.0, CountWhere::SurfaceRegion, {1}}}},
        {0, 4, {{true, 1, 1.0, CountWhere::SurfaceRegion, {2}}}},
        {0, 5, {{true, 0, 1.0, CountWhere::VolumeRegion, {1}}}},
        {0, 6, {{false, 0, 1.0, CountWhere::VolumeRegion, {1}}}}};
      std::vector<uint64_t> rs_sp, rs_rx, cv_sp, cv_rx;
      for (int window = 0; window < 3; window++) {
        ev3.set_barrier_time_for_next_execution(10);
        ev3.step();
        ev3.update_event_time_for_next_scheduled_time();
        cnt3.event_time = ev3.event_time;
        cnt3.step();
        ev3.get_counts(sp, rx);
        ev3.get_counts_by_surface_region(rs_sp, rs_rx);
        ev3.get_counts_by_volume(cv_sp, cv_rx);
        if (sp[1] + sp[2] != n_rec || sp[0] + sp[2] + rx[1] != (uint64_t)n) { printf("surface count identity broken\n"); return 1; }
        if (rs_sp[2 * 3 + 1] + rs_sp[2 * 3 + 2] != sp[2] || rs_sp[2 * 3 + 0] != 0) { printf("LR per region != LR\n"); return 1; }
        if (rs_sp[1 * 3 + 1] + rs_sp[2 * 3 + 1] != n_low) { printf("receptors of region low moved\n"); return 1; }
        if (rs_rx[1 * 3 + 1] + rs_rx[1 * 3 + 2] != rx[1] || rs_rx[0 * 3 + 1] + rs_rx[0 * 3 + 2] != 0) { printf("unbinding per region != unbinding\n"); return 1; }
        if (cv_rx[0 * 2 + 1] != rx[0] || cv_rx[1 * 2 + 0] + cv_rx[1 * 2 + 1] != 0) { printf("binding in the box != binding\n"); return 1; }
        if (cv_sp[0 * 2 + 1] != sp[0]) { printf("L in the box != L\n"); return 1; }
      }
      if (rx[0] < 50 || rx[1] < 5) { printf("too few surface reactions: %llu %llu\n", (unsigned long long)rx[0], (unsigned long long)rx[1]); return 1; }
      buf.flush_and_close();
      ev3.sync_to_host();
      uint64_t surf_back = 0;
      for (const Molecule& m : p3.molecules) if (!m.is_vol()) { surf_back++; if (m.s.wall_index >= 12) return 1; }
      if (surf_back != n_rec) { printf("surface molecules lost in the download: %llu of %llu\n", (unsigned long long)surf_back, (unsigned long long)n_rec); return 1; }
      FILE* f = fopen("/tmp/mcx_host_adapter_counts.gdat", "r");
      if (!f) { printf("count file missing\n"); return 1; }
      int lines = 0; char line[512];
      while (fgets(line, sizeof(line), f)) lines++;
      fclose(f); remove("/tmp/mcx_host_adapter_counts.gdat");
      if (lines != 4) { printf("count file has %d lines\n", lines); return 1; }
    }
    printf("host adapter ok: MSD %.4f, C after 30 iterations %llu\n", msd, (unsigned long long)last_c);
    return 0;
  } catch (const McxFatalError& e) {
    printf("McxFatalError %d: %s\n", e.code, e.what());
    return e.code == MCX_ERR_CUDA ? 3 : 2;
  }
}
