// particles_b200.cpp — see particles_b200.h.  Host C++ only; the GPU is reached through the C ABI.
#include "particles_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

namespace pbfhost {

Particles::Particles(double rho0, const PbfParams* params, int device, bool quiet_ctor)
    : simulate_time(0.0), rest_density(rho0), quiet(quiet_ctor), device_(device) {
  if (params) params_ = *params; else pbf_default_params(&params_);
  params_.rest_density = rho0;
  if (!quiet) { fprintf(stdout, "%s", paramsString().c_str()); fflush(stdout); }   // particles.h:114-116
}

Particles::~Particles() {
  if (scratch_) pbf_destroy(scratch_);
  if (multi_) pbf_multi_destroy(multi_);
  if (handle_) pbf_destroy(handle_);
}

const char* Particles::lastError() const { return multi_ ? pbf_multi_last_error(multi_) : (handle_ ? pbf_last_error(handle_) : "no device handle"); }

void Particles::setDevices(const std::vector<int>& ids) {
  if (uploaded_) { std::cerr << "[pbf_b200] setDevices after the first step is not supported" << std::endl; std::exit(EXIT_FAILURE); }
  devices_ = ids;
  if (devices_.size() == 1) { device_ = devices_[0]; devices_.clear(); }
}

// Multi-device case: the surfacer and the density field run on one device, on a copy of the mirror (fp32 state widened,
// so the copy IS the state), refreshed when a step has been taken since.
pbf_handle* Particles::scratchHandle() {
  if (!scratch_) {
    if (pbf_create(&params_, devices_[0], &scratch_) != PBF_OK) { std::cerr << "[pbf_b200] pbf_create failed for the surfacer" << std::endl; std::exit(EXIT_FAILURE); }
    scratch_step_ = -1;
  }
  if (scratch_step_ != steps_taken) {
    if (pbf_upload(scratch_, ps.size(), pos_.data(), vel_.data()) != PBF_OK) { std::cerr << "[pbf_b200] " << pbf_last_error(scratch_) << std::endl; std::exit(EXIT_FAILURE); }
    scratch_step_ = steps_taken;
  }
  return scratch_;
}

// (Re)create the views after the mirror arrays moved: all four containers grow together, so this runs O(log n) times.
void Particles::rebind() {
  const size_t n = rho_.size();
  const size_t cap = std::max<size_t>(1024, 2 * n);
  pos_.reserve(3 * cap); vel_.reserve(3 * cap); rho_.reserve(cap);
  storage_.clear(); storage_.reserve(cap);
  ps.resize(n);
  for (size_t i = 0; i < n; i++) {
    storage_.emplace_back(reinterpret_cast<Vector3D*>(&pos_[3 * i]), reinterpret_cast<Vector3D*>(&vel_[3 * i]), &rho_[i], rest_density);
    ps[i] = &storage_[i];
  }
}

void Particles::addParticle(Vector3D pos, Vector3D v) {
  if (uploaded_) { std::cerr << "[pbf_b200] addParticle after the first step is not supported" << std::endl; std::exit(EXIT_FAILURE); }
  const bool grow = rho_.size() == rho_.capacity() || storage_.size() == storage_.capacity();
  const double p3[3] = {pos.x, pos.y, pos.z}, v3[3] = {v.x, v.y, v.z};
  pos_.insert(pos_.end(), p3, p3 + 3); vel_.insert(vel_.end(), v3, v3 + 3); rho_.push_back(0.0);
  if (grow) { rebind(); return; }
  const size_t i = rho_.size() - 1;
  storage_.emplace_back(reinterpret_cast<Vector3D*>(&pos_[3 * i]), reinterpret_cast<Vector3D*>(&vel_[3 * i]), &rho_[i], rest_density);
  ps.push_back(&storage_.back());
}

void Particles::ensureUploaded() {
  if (uploaded_) return;
  if (multi()) {
    int rc = pbf_create_multi(&params_, (int)devices_.size(), devices_.data(), &multi_);
    if (rc != PBF_OK) { std::cerr << "[pbf_b200] pbf_create_multi failed (code " << rc << "): no CUDA device / bad device list?" << std::endl; std::exit(EXIT_FAILURE); }
    if ((!spheres_.empty() && pbf_multi_set_obstacle_spheres(multi_, spheres_.size() / 4, spheres_.data()) != PBF_OK) ||
        (!tris_.empty() && pbf_multi_set_obstacle_triangles(multi_, tris_.size() / 18, tris_.data()) != PBF_OK) ||
        pbf_multi_upload(multi_, ps.size(), pos_.data(), vel_.data()) != PBF_OK) {
      std::cerr << "[pbf_b200] upload failed: " << lastError() << std::endl; std::exit(EXIT_FAILURE);
    }
    uploaded_ = true;
    return;
  }
  int rc = pbf_create(&params_, device_, &handle_);
  if (rc != PBF_OK) {   // the reference's error style is exit() (application.cpp:313-317); there is no CPU fallback
    std::cerr << "[pbf_b200] pbf_create failed (code " << rc << "): no CUDA device?" << std::endl;
    std::exit(EXIT_FAILURE);
  }
  const size_t n = ps.size();
  // the mirror arrays live as long as the handle (addParticle is refused from here on): page-lock them once so every
  // step's read-back is a direct DMA
  if (n) { pbf_host_register(handle_, pos_.data(), 3 * n * sizeof(double)); pbf_host_register(handle_, vel_.data(), 3 * n * sizeof(double));
           pbf_host_register(handle_, rho_.data(), n * sizeof(double)); }
  if (!spheres_.empty() && pbf_set_obstacle_spheres(handle_, spheres_.size() / 4, spheres_.data()) != PBF_OK) {
    std::cerr << "[pbf_b200] " << lastError() << std::endl; std::exit(EXIT_FAILURE);
  }
  if (!tris_.empty() && pbf_set_obstacle_triangles(handle_, tris_.size() / 18, tris_.data()) != PBF_OK) {
    std::cerr << "[pbf_b200] " << lastError() << std::endl; std::exit(EXIT_FAILURE);
  }
  rc = pbf_upload(handle_, n, pos_.data(), vel_.data());
  if (rc != PBF_OK) { std::cerr << "[pbf_b200] upload failed: " << lastError() << std::endl; std::exit(EXIT_FAILURE); }
  if (n && mirror_each_step) { pbf_set_readback(handle_, pos_.data(), vel_.data(), rho_.data()); readback_on_ = true; }   // every step streams its result into the mirror
  if (n && neighbor_alert_threshold > 0) pbf_set_neighbor_alert(handle_, neighbor_alert_threshold, std::max<size_t>(neighbor_alert_max_lines, 1));
  uploaded_ = true;
}

// `ps` are views into pos_ / vel_ / rho_, so completing the transfer IS the refresh
void Particles::refreshMirror(bool already_streamed) {
  // after a step the streaming read-back has the data on its way: pbf_sync completes it
  int rc = multi_ ? pbf_multi_download(multi_, pos_.data(), vel_.data(), rho_.data())
                  : (already_streamed ? pbf_sync(handle_) : pbf_download(handle_, pos_.data(), vel_.data(), rho_.data()));
  if (rc != PBF_OK) { std::cerr << "[pbf_b200] step failed: " << lastError() << std::endl; std::exit(EXIT_FAILURE); }
}

// particles.cpp:165-173: `cerr << *this << " only has " << neighbors.size() << " neighbors."` with
// operator<<(Particle) = "P(p" << new_position << ",v" << velocity << ')' (particles.cpp:153-156)
void Particles::reportNeighborAlerts() {
  neighbor_alerts_last_step = 0;
  if (quiet || neighbor_alert_threshold <= 0 || !handle_) return;
  const size_t cap = std::max<size_t>(neighbor_alert_max_lines, 1);
  std::vector<uint32_t> ids(cap), cnt(cap); std::vector<double> xp(3 * cap), vp(3 * cap);
  size_t total = 0, shown = 0;
  if (pbf_get_neighbor_alerts(handle_, cap, ids.data(), cnt.data(), xp.data(), vp.data(), &shown, &total) != PBF_OK) return;
  neighbor_alerts_last_step = total;
  for (size_t k = 0; k < shown; k++)
    std::cerr << "P(p(" << xp[3*k] << "," << xp[3*k+1] << "," << xp[3*k+2] << "),v(" << vp[3*k] << "," << vp[3*k+1] << "," << vp[3*k+2] << "))"
              << " only has " << cnt[k] << " neighbors." << std::endl;
  if (total > shown) std::cerr << "[pbf_b200] ... and " << (total - shown) << " more particles with fewer than " << neighbor_alert_threshold << " neighbors." << std::endl;
}

void Particles::estimateDensities() {
  ensureUploaded();
  if ((multi_ ? pbf_multi_estimate_densities(multi_) : pbf_estimate_densities(handle_)) != PBF_OK) { std::cerr << "[pbf_b200] " << lastError() << std::endl; std::exit(EXIT_FAILURE); }
  refreshMirror();
}

void Particles::setObstacleSpheres(const std::vector<double>& s) {
  spheres_ = s;
  if (scratch_) pbf_set_obstacle_spheres(scratch_, spheres_.size() / 4, spheres_.data());
  if ((multi_ && pbf_multi_set_obstacle_spheres(multi_, spheres_.size() / 4, spheres_.data()) != PBF_OK) ||
      (handle_ && pbf_set_obstacle_spheres(handle_, spheres_.size() / 4, spheres_.data()) != PBF_OK)) {
    std::cerr << "[pbf_b200] " << lastError() << std::endl; std::exit(EXIT_FAILURE);
  }
}

void Particles::setObstacleTriangles(const std::vector<double>& t) {
  tris_ = t;
  if ((multi_ && pbf_multi_set_obstacle_triangles(multi_, tris_.size() / 18, tris_.data()) != PBF_OK) ||
      (handle_ && pbf_set_obstacle_triangles(handle_, tris_.size() / 18, tris_.data()) != PBF_OK)) {
    std::cerr << "[pbf_b200] " << lastError() << std::endl; std::exit(EXIT_FAILURE);
  }
}

void Particles::timeStep(double delta_t) {
  if (std::fabs(delta_t - params_.dt) > 1e-15) {
    std::cerr << "[pbf_b200] timeStep(dt): dt is fixed at construction (PbfParams.dt = " << params_.dt << ")" << std::endl;
    std::exit(EXIT_FAILURE);
  }
  ensureUploaded();
  if (!quiet) std::cerr << "Time: " << simulate_time;          // particles.cpp:251-253
  simulate_time += delta_t;
  if (!quiet) std::cerr << " => " << simulate_time << std::endl;
  if (handle_ && ps.size() > 0 && readback_on_ != mirror_each_step) {       // the switch was flipped between steps
    if (mirror_each_step) pbf_set_readback(handle_, pos_.data(), vel_.data(), rho_.data()); else pbf_set_readback(handle_, nullptr, nullptr, nullptr);
    readback_on_ = mirror_each_step;
  }
  if ((multi_ ? pbf_multi_step(multi_, 1) : pbf_step(handle_, 1)) != PBF_OK) { std::cerr << "[pbf_b200] " << lastError() << std::endl; std::exit(EXIT_FAILURE); }
  if (mirror_each_step) { refreshMirror(/*already_streamed=*/ps.size() > 0 && readback_on_); mirror_stale_ = false; }
  else {                                                                   // errors of the step still surface here, `ps` is refreshed on demand
    mirror_stale_ = true;
    if ((multi_ ? pbf_multi_sync(multi_) : pbf_sync(handle_)) != PBF_OK) { std::cerr << "[pbf_b200] step failed: " << lastError() << std::endl; std::exit(EXIT_FAILURE); }
  }
  reportNeighborAlerts();                                        // particles.cpp:268-270 (initializeWithNewNeighbors); single device only
  double ms = 0;
  if (multi_) pbf_multi_stats(multi_, &avg_rho_first_iter, &avg_rho_final, &ms);
  else pbf_stats(handle_, &avg_rho_first_iter, &avg_rho_final, &ms);
  if (!quiet) std::cout << "avg rho: " << avg_rho_first_iter << " => " << avg_rho_final << std::endl;   // particles.cpp:267,279,295
  surfaceUpToTimestep = false;                                   // particles.cpp:296
  steps_taken++;
}

void Particles::syncMirror() {
  if (!uploaded_ || !mirror_stale_) return;
  refreshMirror(false);
  mirror_stale_ = false;
}

void Particles::timeStep() { timeStep(params_.dt); }             // DEFAULT_DELTA_T, particles.cpp:299-301

std::vector<Particles::SurfaceTriangle> Particles::getSurfacePrims(double isolevel, double fStepSize) {
  ensureUploaded();
  const double lo[3] = {surface_min.x, surface_min.y, surface_min.z}, hi[3] = {surface_max.x, surface_max.y, surface_max.z};
  const double grad_eps = 0.001;                                 // GRADIENT_EPS, particles.cpp:16
  size_t nt = 0;
  pbf_handle* hs = multi_ ? scratchHandle() : handle_;
  if (pbf_extract_surface(hs, lo, hi, isolevel, fStepSize, grad_eps, 0, nullptr, &nt) != PBF_OK) { std::cerr << "[pbf_b200] " << pbf_last_error(hs) << std::endl; std::exit(EXIT_FAILURE); }
  std::vector<double> buf(18 * nt);
  if (nt && pbf_extract_surface(hs, lo, hi, isolevel, fStepSize, grad_eps, nt, buf.data(), &nt) != PBF_OK) { std::cerr << "[pbf_b200] " << pbf_last_error(hs) << std::endl; std::exit(EXIT_FAILURE); }
  std::vector<SurfaceTriangle> out(nt);
  for (size_t t = 0; t < nt; t++) {
    const double* q = &buf[18 * t];
    out[t] = SurfaceTriangle{Vector3D(q[0], q[1], q[2]), Vector3D(q[3], q[4], q[5]), Vector3D(q[6], q[7], q[8]),
                             Vector3D(q[9], q[10], q[11]), Vector3D(q[12], q[13], q[14]), Vector3D(q[15], q[16], q[17])};
  }
  return out;
}

void Particles::updateSurface() {                                // particles.cpp:393-402
  if (surfaceUpToTimestep) return;
  surface = getSurfacePrims(0.95 * rest_density, params_.h * 0.5);   // ISO_LEVEL_REST_DENSITY_RATIO, FSTEPSIZE_RATIO (particles.cpp:14,18)
  surfaceUpToTimestep = true;
}

double Particles::estimateDensityAt(Vector3D pos) const {
  const_cast<Particles*>(this)->syncMirror();                    // mirror_each_step == false: `ps` may be behind the device
  const double H = params_.h, H2 = H * H;
  double H9 = 1; for (int i = 0; i < 9; i++) H9 *= H;
  double density = 0.0;
  for (const Particle* p : ps) {
    const double dx = p->position->x - pos.x, dy = p->position->y - pos.y, dz = p->position->z - pos.z;
    const double r2 = dx * dx + dy * dy + dz * dz;
    if (r2 >= H2) continue;
    const double t = H2 - r2;
    density += 1.56668147106 * (t * t * t) / H9;
  }
  return density;
}

std::vector<double> Particles::estimateDensitiesAt(const std::vector<Vector3D>& points) {
  ensureUploaded();
  std::vector<double> q(3 * points.size()), out(points.size());
  for (size_t i = 0; i < points.size(); i++) { q[3*i] = points[i].x; q[3*i+1] = points[i].y; q[3*i+2] = points[i].z; }
  pbf_handle* hs = multi_ ? scratchHandle() : handle_;
  if (pbf_density_at(hs, points.size(), q.data(), out.data()) != PBF_OK) { std::cerr << "[pbf_b200] " << pbf_last_error(hs) << std::endl; std::exit(EXIT_FAILURE); }
  return out;
}

// ---- restart files ------------------------------------------------------------------------------
bool Particles::saveCheckpoint(const char* filename, std::string* error) const {
  const_cast<Particles*>(this)->syncMirror();
  FILE* f = fopen(filename, "wb");
  if (!f) { if (error) *error = std::string("cannot open ") + filename; return false; }
  const int64_t n = (int64_t)ps.size(), steps = steps_taken, psz = (int64_t)sizeof(PbfParams), ns = (int64_t)(spheres_.size() / 4),
                nt = (int64_t)(tris_.size() / 18);
  // the host mirror is exactly the fp32 device state, widened
  bool ok = fwrite("PBFCKPT2", 1, 8, f) == 8 && fwrite(&n, 8, 1, f) == 1 && fwrite(&steps, 8, 1, f) == 1 &&
            fwrite(&simulate_time, 8, 1, f) == 1 && fwrite(&rest_density, 8, 1, f) == 1 && fwrite(&psz, 8, 1, f) == 1 &&
            fwrite(&params_, sizeof(PbfParams), 1, f) == 1 && fwrite(&ns, 8, 1, f) == 1 &&
            (ns == 0 || fwrite(spheres_.data(), 8, spheres_.size(), f) == spheres_.size()) &&
            fwrite(&nt, 8, 1, f) == 1 && (nt == 0 || fwrite(tris_.data(), 8, tris_.size(), f) == tris_.size()) &&
            (n == 0 || (fwrite(pos_.data(), 8, 3 * (size_t)n, f) == 3 * (size_t)n && fwrite(vel_.data(), 8, 3 * (size_t)n, f) == 3 * (size_t)n &&
                        fwrite(rho_.data(), 8, (size_t)n, f) == (size_t)n));
  ok = (fclose(f) == 0) && ok;
  if (!ok && error) *error = std::string("short write to ") + filename;
  return ok;
}

Particles* Particles::loadCheckpoint(const char* filename, std::string* error, int device, bool quiet_ctor, const std::vector<int>* devices) {
  auto fail = [&](const std::string& m) -> Particles* { if (error) *error = m; return nullptr; };
  FILE* f = fopen(filename, "rb");
  if (!f) return fail(std::string("cannot open ") + filename);
  char magic[8]; int64_t n = 0, steps = 0, psz = 0, ns = 0, nt = 0; double t = 0, rho0 = 0; PbfParams prm;
  bool ok = fread(magic, 1, 8, f) == 8 && (memcmp(magic, "PBFCKPT2", 8) == 0 || memcmp(magic, "PBFCKPT1", 8) == 0);
  const bool v2 = ok && magic[7] == '2';
  ok = ok && fread(&n, 8, 1, f) == 1 && fread(&steps, 8, 1, f) == 1 &&
       fread(&t, 8, 1, f) == 1 && fread(&rho0, 8, 1, f) == 1 && fread(&psz, 8, 1, f) == 1 && psz == (int64_t)sizeof(PbfParams) &&
       fread(&prm, sizeof(PbfParams), 1, f) == 1 && fread(&ns, 8, 1, f) == 1 && n >= 0 && ns >= 0 && ns <= PBF_MAX_SPHERES;
  if (!ok) { fclose(f); return fail("not a PBFCKPT checkpoint (or written with another PbfParams layout)"); }
  std::vector<double> sph(4 * (size_t)ns), tri, buf(7 * (size_t)n);
  ok = ns == 0 || fread(sph.data(), 8, sph.size(), f) == sph.size();
  if (ok && v2) {
    ok = fread(&nt, 8, 1, f) == 1 && nt >= 0 && (uint64_t)nt <= PBF_MAX_TRIANGLES;
    if (ok) { tri.resize(18 * (size_t)nt); ok = nt == 0 || fread(tri.data(), 8, tri.size(), f) == tri.size(); }
  }
  ok = ok && (n == 0 || fread(buf.data(), 8, buf.size(), f) == buf.size());
  fclose(f);
  if (!ok) return fail("truncated checkpoint");
  Particles* ps = new Particles(rho0, &prm, device, quiet_ctor);
  if (devices) ps->setDevices(*devices);
  for (int64_t i = 0; i < n; i++)
    ps->addParticle(Vector3D(buf[3*i], buf[3*i+1], buf[3*i+2]), Vector3D(buf[3*n + 3*i], buf[3*n + 3*i+1], buf[3*n + 3*i+2]));
  for (int64_t i = 0; i < n; i++) ps->rho_[i] = buf[6*n + i];
  ps->simulate_time = t; ps->steps_taken = steps;
  if (ns) ps->setObstacleSpheres(sph);
  if (nt) ps->setObstacleTriangles(tri);
  return ps;
}

std::string Particles::paramsString() const {
  std::stringstream ss;
  ss << "Fluid simulation parameters: " << std::endl
     << "\tTime step: " << params_.dt << std::endl
     << "\tSPH Density estimate radius H: " << params_.h << std::endl
     << "\tNewton steps: " << params_.iterations << std::endl
     << "\tConstraint relaxation epsilon: " << params_.eps_relax << std::endl
     << "\tTensile artificial pressure coefficient K: " << params_.k_corr << std::endl
     << "\tTensile artificial pressure exponent N: " << params_.n_corr << std::endl
     << "\tVorticity confinement coefficient epsilon: " << params_.vort_eps << std::endl
     << "\tViscosity coefficient C: " << params_.visc_c << std::endl
     << "\tBackend: B200 CUDA (libpbf_b200), fp32, Jacobi XSPH" << (multi() ? ", " + std::to_string(devices_.size()) + " x-slabs" : std::string()) << std::endl;
  return ss.str();
}

// ---- XML ---------------------------------------------------------------------------------------------
namespace {
// Minimal streaming tokenizer for the particle schema: yields (tag name, text content) for leaf
// elements, ignoring attributes, comments, declarations and whitespace.
struct XmlLeafReader {
  std::ifstream in;
  explicit XmlLeafReader(const char* f) : in(f, std::ios::in | std::ios::binary) {}
  bool ok() const { return in.is_open(); }
  // reads the next tag; returns false at EOF.  closing=true for </tag>
  bool nextTag(std::string& name, bool& closing, std::string& text_before) {
    text_before.clear();
    int c;
    while ((c = in.get()) != EOF && c != '<') text_before.push_back((char)c);
    if (c == EOF) return false;
    std::string tag;
    while ((c = in.get()) != EOF && c != '>') tag.push_back((char)c);
    if (c == EOF) return false;
    if (!tag.empty() && (tag[0] == '?' || tag[0] == '!')) return nextTag(name, closing, text_before);
    closing = !tag.empty() && tag[0] == '/';
    size_t b = closing ? 1 : 0, e = b;
    while (e < tag.size() && !isspace((unsigned char)tag[e]) && tag[e] != '/') e++;
    name = tag.substr(b, e - b);
    return true;
  }
};

bool parse3(const std::string& s, double out[3]) {   // Application::stov (application.cpp:293-300)
  std::stringstream ss(s);
  return (bool)(ss >> out[0] >> out[1] >> out[2]);
}
}  // namespace

bool parse_particles_xml(const char* filename, std::vector<double>& pos, std::vector<double>& vel, double& rho0, std::string* error) {
  auto fail = [&](const std::string& m) { if (error) *error = m; return false; };
  XmlLeafReader r(filename);
  if (!r.ok()) return fail(std::string("cannot open ") + filename);
  pos.clear(); vel.clear();
  bool have_root = false, have_density = false, in_particle = false, got_pos = false, got_v = false;
  double p3[3] = {0, 0, 0}, v3[3] = {0, 0, 0};
  std::string name, text; bool closing;
  while (r.nextTag(name, closing, text)) {
    if (!closing) {
      if (name == "particles") have_root = true;
      else if (name == "particle") { in_particle = true; got_pos = got_v = false; }
      continue;
    }
    if (name == "density" && !in_particle) {
      try { rho0 = (double)std::stof(text); } catch (...) { return fail("bad <density>"); }   // stof: Q17
      have_density = true;
    } else if (name == "pos" && in_particle) {
      if (!parse3(text, p3)) return fail("bad <pos>: " + text);
      got_pos = true;
    } else if (name == "v" && in_particle) {
      if (!parse3(text, v3)) return fail("bad <v>: " + text);
      got_v = true;
    } else if (name == "particle") {
      if (!got_pos || !got_v) return fail("<particle> without <pos> or <v>");
      pos.insert(pos.end(), p3, p3 + 3); vel.insert(vel.end(), v3, v3 + 3);
      in_particle = false;
    }
  }
  if (!have_root) return fail("Not a particles file!");
  if (!have_density) return fail("missing <density>");
  return true;
}

Particles* load_particles_xml(const char* filename, std::string* error, const PbfParams* params, int device, bool quiet, const std::vector<int>* devices) {
  std::vector<double> pos, vel; double rho0 = 1000.0;
  if (!parse_particles_xml(filename, pos, vel, rho0, error)) return nullptr;
  Particles* particles = new Particles(rho0, params, device, quiet);
  if (devices) particles->setDevices(*devices);
  const size_t n = pos.size() / 3;
  for (size_t i = 0; i < n; i++)
    particles->addParticle(Vector3D(pos[3*i], pos[3*i+1], pos[3*i+2]), Vector3D(vel[3*i], vel[3*i+1], vel[3*i+2]));
  particles->estimateDensities();
  return particles;
}

}  // namespace pbfhost
