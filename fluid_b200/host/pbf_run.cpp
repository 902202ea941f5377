// pbf_run — windowless driver mirroring the reference's `pathtracer -p particles.xml -d seconds`
// simulation loop (main.cpp:110-117,159-171; pathtracer.cpp:444-480) without the renderer:
//   load_particles -> estimateDensities -> while (simulate_time < T) timeStep()   (63 steps / second, Q18)
// Optional --dump writes the per-step state in the same PBFDUMP1 format as oracle/ref_harness.
//   pbf_run -p particles.xml [-d seconds | --steps N] [--dump out.bin] [--quiet] [--parse-only] [--iterations I]
//           [--sphere cx cy cz r]...      obstacle spheres of the collision scene (the CBspheres scenes hold two)
//           [--tris file]                 obstacle triangles (int64 count, then p1 p2 p3 n1 n2 n3 as 18 doubles each)
//           [--surface file]              after the last step: updateSurface() (marching cubes on the GPU), int64 count + 18 doubles
//                                         per triangle (p1 p2 p3 n1 n2 n3), the layout of oracle/ref_harness --surface
//           [--gpus N | --devices a,b,c]  several GPUs behind the same Particles object: x-slabs, halos by peer stores over NVLink
//                                         (pbf_create_multi); bit-identical to one GPU.  --box x0 y0 z0 x1 y1 z1 replaces the Cornell box
//           [--block nx ny nz [--rho0 r]] instead of -p: a pgen.py-style lattice block (spacing 0.1 from (0.1, 0.1, 0.1), v = (0, -1, 0), index order
//                                         x outer / z inner, particles/pgen.py:54-61) generated in memory -- a 128M-particle XML file is 10 GB of text
//           [--lazy-mirror]               read the particles back only after the last step (Particles::mirror_each_step = false); ignored with --dump
//           [--save-state f] [--load-state f]   restart files (PBFCKPT1, particles_b200.h); a continued run is bit-identical
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "particles_b200.h"

using namespace pbfhost;

int main(int argc, char** argv) {
  const char* pfile = nullptr; const char* dump = nullptr; const char* save_state = nullptr; const char* load_state = nullptr;
  double seconds = -1; int steps = -1, iterations = -1; bool quiet = false, parse_only = false, lazy = false;
  std::vector<double> spheres, box; long long blk[3] = {0, 0, 0}; double blk_rho0 = 700.0;
  std::vector<int> devices;
  const char* trisfile = nullptr; const char* surffile = nullptr;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "-p" && i + 1 < argc) pfile = argv[++i];
    else if (a == "-d" && i + 1 < argc) seconds = atof(argv[++i]);
    else if (a == "--steps" && i + 1 < argc) steps = atoi(argv[++i]);
    else if (a == "--iterations" && i + 1 < argc) iterations = atoi(argv[++i]);
    else if (a == "--dump" && i + 1 < argc) dump = argv[++i];
    else if (a == "--sphere" && i + 4 < argc) { for (int k = 0; k < 4; k++) spheres.push_back(atof(argv[++i])); }   // obstacle sphere cx cy cz r, repeatable
    else if (a == "--gpus" && i + 1 < argc) { const int g = atoi(argv[++i]); devices.clear(); for (int d = 0; d < g; d++) devices.push_back(d); }
    else if (a == "--devices" && i + 1 < argc) { devices.clear(); for (char* t = strtok(argv[++i], ","); t; t = strtok(nullptr, ",")) devices.push_back(atoi(t)); }
    else if (a == "--box" && i + 6 < argc) { for (int k = 0; k < 6; k++) box.push_back(atof(argv[++i])); }   // simulation box instead of the Cornell box
    else if (a == "--block" && i + 3 < argc) { for (int k = 0; k < 3; k++) blk[k] = atoll(argv[++i]); }
    else if (a == "--rho0" && i + 1 < argc) blk_rho0 = atof(argv[++i]);
    else if (a == "--tris" && i + 1 < argc) trisfile = argv[++i];            // obstacle triangles: int64 count + 18 doubles each
    else if (a == "--surface" && i + 1 < argc) surffile = argv[++i];          // marching-cubes surface of the final state
    else if (a == "--save-state" && i + 1 < argc) save_state = argv[++i];   // restart file written after the last step
    else if (a == "--load-state" && i + 1 < argc) load_state = argv[++i];   // continue from a restart file instead of -p
    else if (a == "--quiet") quiet = true;
    else if (a == "--lazy-mirror") lazy = true;
    else if (a == "--parse-only") parse_only = true;
    else { fprintf(stderr, "usage: pbf_run -p particles.xml [-d seconds | --steps N] [--dump out.bin] [--quiet] [--parse-only] [--sphere cx cy cz r]...\n"); return 2; }
  }
  const bool have_block = blk[0] > 0 && blk[1] > 0 && blk[2] > 0;
  if (!pfile && !load_state && !have_block) { printf("[Warning] Particle file not passed in or not found\n[Warning] use -p <particle_file_path>\n"); return 2; }
  std::string err;
  if (parse_only && !pfile) { printf("[ERROR] --parse-only needs -p <particle_file_path>\n"); return 2; }
  if (load_state && iterations >= 0) { printf("[ERROR] --iterations cannot be combined with --load-state (the checkpoint carries its parameters)\n"); return 2; }
  if (parse_only) {
    std::vector<double> pos, vel; double rho0 = 0;
    if (!parse_particles_xml(pfile, pos, vel, rho0, &err)) { printf("[ERROR] XML error: %s\n", err.c_str()); return 1; }
    double sp = 0, sv = 0;
    for (double v : pos) sp += v;
    for (double v : vel) sv += v;
    printf("{\"n\": %zu, \"rho0\": %.17g, \"sum_pos\": %.17g, \"sum_vel\": %.17g}\n", pos.size() / 3, rho0, sp, sv);
    return 0;
  }
  PbfParams prm; pbf_default_params(&prm);
  if (iterations >= 0) prm.iterations = iterations;
  if (box.size() == 6) { for (int k = 0; k < 3; k++) { prm.box_min[k] = box[k]; prm.box_max[k] = box[3 + k]; } prm.y_light = box[4]; prm.z_front = box[5]; }
  const std::vector<int>* devs = devices.empty() ? nullptr : &devices;
  printf("[Fluid Simulation] Loading particle file...");
  Particles* ps = nullptr;
  if (have_block && !pfile && !load_state) {                   // what Application::load_particles does (application.cpp:302-344), from a generator
    ps = new Particles(blk_rho0, &prm, 0, quiet);
    if (devs) ps->setDevices(*devs);
    for (long long i = 0; i < blk[0]; i++)
      for (long long j = 0; j < blk[1]; j++)
        for (long long k = 0; k < blk[2]; k++) ps->addParticle(Vector3D(0.1 + 0.1 * i, 0.1 + 0.1 * j, 0.1 + 0.1 * k), Vector3D(0.0, -1.0, 0.0));
    ps->estimateDensities();
  } else
    ps = load_state ? load_checkpoint(load_state, &err, 0, quiet, devs) : load_particles_xml(pfile, &err, &prm, 0, quiet, devs);
  if (!ps) { printf("[ERROR] %s: %s\n", load_state ? "checkpoint error" : "XML error", err.c_str()); return EXIT_FAILURE; }   // application.cpp:313-317
  printf("Done!\n");
  ps->quiet = quiet;
  if (box.size() == 6) { ps->surface_min = Vector3D(box[0], box[1], box[2]); ps->surface_max = Vector3D(box[3], box[4], box[5]); }   // the surfacer's lattice follows the box
  if (!spheres.empty()) ps->setObstacleSpheres(spheres);
  if (trisfile) {
    FILE* tf = fopen(trisfile, "rb");
    int64_t nt = 0;
    if (!tf || fread(&nt, 8, 1, tf) != 1 || nt < 0) { printf("[ERROR] cannot read %s\n", trisfile); return EXIT_FAILURE; }
    std::vector<double> tv(18 * (size_t)nt);
    if (fread(tv.data(), 8, tv.size(), tf) != tv.size()) { printf("[ERROR] short %s\n", trisfile); return EXIT_FAILURE; }
    fclose(tf);
    ps->setObstacleTriangles(tv);
  }
  const int64_t n = (int64_t)ps->ps.size();
  if (steps < 0 && seconds < 0) steps = 1;
  if (lazy && !dump) ps->mirror_each_step = false;
  FILE* f = dump ? fopen(dump, "wb") : nullptr;
  std::vector<std::vector<double>> frames;
  auto t0 = std::chrono::steady_clock::now();
  int done = 0;
  const double start = ps->simulate_time;
  while (steps >= 0 ? done < steps : ps->simulate_time < start + seconds) {   // pathtracer.cpp:454,475
    ps->timeStep();
    done++;
    if (f) {
      std::vector<double> st(8 * n);
      for (int64_t i = 0; i < n; i++) {
        const Particle* p = ps->ps[i];
        Vector3D x = p->getPosition();
        st[8*i] = x.x; st[8*i+1] = x.y; st[8*i+2] = x.z;
        st[8*i+3] = p->velocity.x; st[8*i+4] = p->velocity.y; st[8*i+5] = p->velocity.z;
        st[8*i+6] = p->getLatestDensityEstimate(); st[8*i+7] = 0.0;
      }
      frames.push_back(std::move(st));
    }
  }
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  ps->syncMirror();                                              // --lazy-mirror: the one read-back of the run
  if (f) {
    int64_t hdr[2] = {n, done};
    fwrite("PBFDUMP1", 1, 8, f); fwrite(hdr, 8, 2, f);
    std::vector<int32_t> zero(n, 0);
    double per = done ? secs / done : 0.0;
    for (auto& st : frames) { fwrite(st.data(), 8, st.size(), f); fwrite(zero.data(), 4, n, f); fwrite(&per, 8, 1, f); }
    fclose(f);
  }
  if (surffile) {
    ps->updateSurface();
    FILE* sf = fopen(surffile, "wb");
    if (!sf) { printf("[ERROR] cannot open %s\n", surffile); return EXIT_FAILURE; }
    const int64_t nt = (int64_t)ps->surface.size();
    fwrite(&nt, 8, 1, sf);
    for (const Particles::SurfaceTriangle& t : ps->surface) fwrite(&t, sizeof(double), 18, sf);
    fclose(sf);
    fprintf(stderr, "including %lld marching cube surfacing triangles\n", (long long)nt);   // pathtracer.cpp:250
  }
  if (save_state && !ps->saveCheckpoint(save_state, &err)) { printf("[ERROR] %s\n", err.c_str()); return EXIT_FAILURE; }
  fprintf(stderr, "{\"n\": %lld, \"steps\": %d, \"devices\": %d, \"seconds_total\": %.6f, \"ms_per_step_incl_readback\": %.4f}\n", (long long)n, done,
          ps->numDevices(), secs, done ? 1e3 * secs / done : 0.0);
  delete ps;
  return 0;
}
