// particles_b200.cpp — see particles_b200.h.  Host C++ only; the GPU is reached through the C ABI.
#include "particles_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

namespace pbfhost {

Particles::Particles(double rho0, const PbfParams* params, int device)
    : simulate_time(0.0), rest_density(rho0), device_(device) {
  if (params) params_ = *params; else pbf_default_params(&params_);
  params_.rest_density = rho0;
  if (!quiet) { fprintf(stdout, "%s", paramsString().c_str()); fflush(stdout); }   // particles.h:114-116
}

Particles::~Particles() {
  for (Particle* p : ps) delete p;
  if (handle_) pbf_destroy(handle_);
}

const char* Particles::lastError() const { return handle_ ? pbf_last_error(handle_) : "no device handle"; }

void Particles::addParticle(Vector3D pos, Vector3D v) {
  if (uploaded_) { std::cerr << "[pbf_b200] addParticle after the first step is not supported" << std::endl; std::exit(EXIT_FAILURE); }
  ps.push_back(new Particle(pos, v, rest_density));
}

void Particles::ensureUploaded() {
  if (uploaded_) return;
  int rc = pbf_create(&params_, device_, &handle_);
  if (rc != PBF_OK) {   // the reference's error style is exit() (application.cpp:313-317); there is no CPU fallback
    std::cerr << "[pbf_b200] pbf_create failed (code " << rc << "): no CUDA device?" << std::endl;
    std::exit(EXIT_FAILURE);
  }
  const size_t n = ps.size();
  pos_.resize(3 * n); vel_.resize(3 * n); rho_.assign(n, 0.0);
  for (size_t i = 0; i < n; i++) {
    pos_[3*i] = ps[i]->position.x; pos_[3*i+1] = ps[i]->position.y; pos_[3*i+2] = ps[i]->position.z;
    vel_[3*i] = ps[i]->velocity.x; vel_[3*i+1] = ps[i]->velocity.y; vel_[3*i+2] = ps[i]->velocity.z;
  }
  // the mirror arrays live as long as the handle: page-lock them once so every step's read-back is a direct DMA
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
  if (n) pbf_set_readback(handle_, pos_.data(), vel_.data(), rho_.data());   // every step streams its result into the mirror
  uploaded_ = true;
}

void Particles::refreshMirror(bool already_streamed) {
  const size_t n = ps.size();
  // after a step the streaming read-back has the data on its way: pbf_sync completes it
  int rc = already_streamed ? pbf_sync(handle_) : pbf_download(handle_, pos_.data(), vel_.data(), rho_.data());
  if (rc != PBF_OK) { std::cerr << "[pbf_b200] step failed: " << lastError() << std::endl; std::exit(EXIT_FAILURE); }
  for (size_t i = 0; i < n; i++) {
    Particle* p = ps[i];
    p->position = Vector3D(pos_[3*i], pos_[3*i+1], pos_[3*i+2]);
    p->velocity = Vector3D(vel_[3*i], vel_[3*i+1], vel_[3*i+2]);
    p->density = rho_[i];
  }
}

void Particles::estimateDensities() {
  ensureUploaded();
  if (pbf_estimate_densities(handle_) != PBF_OK) { std::cerr << "[pbf_b200] " << lastError() << std::endl; std::exit(EXIT_FAILURE); }
  refreshMirror();
}

void Particles::setObstacleSpheres(const std::vector<double>& s) {
  spheres_ = s;
  if (handle_ && pbf_set_obstacle_spheres(handle_, spheres_.size() / 4, spheres_.data()) != PBF_OK) {
    std::cerr << "[pbf_b200] " << lastError() << std::endl; std::exit(EXIT_FAILURE);
  }
}

void Particles::setObstacleTriangles(const std::vector<double>& t) {
  tris_ = t;
  if (handle_ && pbf_set_obstacle_triangles(handle_, tris_.size() / 18, tris_.data()) != PBF_OK) {
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
  if (pbf_step(handle_, 1) != PBF_OK) { std::cerr << "[pbf_b200] " << lastError() << std::endl; std::exit(EXIT_FAILURE); }
  refreshMirror(/*already_streamed=*/ps.size() > 0);
  double ms = 0;
  pbf_stats(handle_, &avg_rho_first_iter, &avg_rho_final, &ms);
  if (!quiet) std::cout << "avg rho: " << avg_rho_first_iter << " => " << avg_rho_final << std::endl;   // particles.cpp:267,279,295
  surfaceUpToTimestep = false;                                   // particles.cpp:296
  steps_taken++;
}

void Particles::timeStep() { timeStep(params_.dt); }             // DEFAULT_DELTA_T, particles.cpp:299-301

std::vector<Particles::SurfaceTriangle> Particles::getSurfacePrims(double isolevel, double fStepSize) {
  ensureUploaded();
  const double lo[3] = {surface_min.x, surface_min.y, surface_min.z}, hi[3] = {surface_max.x, surface_max.y, surface_max.z};
  const double grad_eps = 0.001;                                 // GRADIENT_EPS, particles.cpp:16
  size_t nt = 0;
  if (pbf_extract_surface(handle_, lo, hi, isolevel, fStepSize, grad_eps, 0, nullptr, &nt) != PBF_OK) { std::cerr << "[pbf_b200] " << lastError() << std::endl; std::exit(EXIT_FAILURE); }
  std::vector<double> buf(18 * nt);
  if (nt && pbf_extract_surface(handle_, lo, hi, isolevel, fStepSize, grad_eps, nt, buf.data(), &nt) != PBF_OK) { std::cerr << "[pbf_b200] " << lastError() << std::endl; std::exit(EXIT_FAILURE); }
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
  const double H = params_.h, H2 = H * H;
  double H9 = 1; for (int i = 0; i < 9; i++) H9 *= H;
  double density = 0.0;
  for (const Particle* p : ps) {
    const double dx = p->position.x - pos.x, dy = p->position.y - pos.y, dz = p->position.z - pos.z;
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
  if (pbf_density_at(handle_, points.size(), q.data(), out.data()) != PBF_OK) { std::cerr << "[pbf_b200] " << lastError() << std::endl; std::exit(EXIT_FAILURE); }
  return out;
}

// ---- restart files ------------------------------------------------------------------------------
bool Particles::saveCheckpoint(const char* filename, std::string* error) const {
  FILE* f = fopen(filename, "wb");
  if (!f) { if (error) *error = std::string("cannot open ") + filename; return false; }
  const int64_t n = (int64_t)ps.size(), steps = steps_taken, psz = (int64_t)sizeof(PbfParams), ns = (int64_t)(spheres_.size() / 4);
  std::vector<double> buf(7 * (size_t)n);
  for (int64_t i = 0; i < n; i++) {                // from the host mirror: exactly the fp32 device state, widened
    const Particle* p = ps[i];
    buf[3*i] = p->position.x; buf[3*i+1] = p->position.y; buf[3*i+2] = p->position.z;
    buf[3*n + 3*i] = p->velocity.x; buf[3*n + 3*i+1] = p->velocity.y; buf[3*n + 3*i+2] = p->velocity.z;
    buf[6*n + i] = p->density;
  }
  bool ok = fwrite("PBFCKPT1", 1, 8, f) == 8 && fwrite(&n, 8, 1, f) == 1 && fwrite(&steps, 8, 1, f) == 1 &&
            fwrite(&simulate_time, 8, 1, f) == 1 && fwrite(&rest_density, 8, 1, f) == 1 && fwrite(&psz, 8, 1, f) == 1 &&
            fwrite(&params_, sizeof(PbfParams), 1, f) == 1 && fwrite(&ns, 8, 1, f) == 1 &&
            (ns == 0 || fwrite(spheres_.data(), 8, spheres_.size(), f) == spheres_.size()) &&
            (n == 0 || fwrite(buf.data(), 8, buf.size(), f) == buf.size());
  ok = (fclose(f) == 0) && ok;
  if (!ok && error) *error = std::string("short write to ") + filename;
  return ok;
}

Particles* Particles::loadCheckpoint(const char* filename, std::string* error, int device) {
  auto fail = [&](const std::string& m) -> Particles* { if (error) *error = m; return nullptr; };
  FILE* f = fopen(filename, "rb");
  if (!f) return fail(std::string("cannot open ") + filename);
  char magic[8]; int64_t n = 0, steps = 0, psz = 0, ns = 0; double t = 0, rho0 = 0; PbfParams prm;
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "PBFCKPT1", 8) == 0 && fread(&n, 8, 1, f) == 1 && fread(&steps, 8, 1, f) == 1 &&
            fread(&t, 8, 1, f) == 1 && fread(&rho0, 8, 1, f) == 1 && fread(&psz, 8, 1, f) == 1 && psz == (int64_t)sizeof(PbfParams) &&
            fread(&prm, sizeof(PbfParams), 1, f) == 1 && fread(&ns, 8, 1, f) == 1 && n >= 0 && ns >= 0 && ns <= PBF_MAX_SPHERES;
  if (!ok) { fclose(f); return fail("not a PBFCKPT1 checkpoint (or written with another PbfParams layout)"); }
  std::vector<double> sph(4 * (size_t)ns), buf(7 * (size_t)n);
  ok = (ns == 0 || fread(sph.data(), 8, sph.size(), f) == sph.size()) && (n == 0 || fread(buf.data(), 8, buf.size(), f) == buf.size());
  fclose(f);
  if (!ok) return fail("truncated checkpoint");
  Particles* ps = new Particles(rho0, &prm, device);
  for (int64_t i = 0; i < n; i++) {
    ps->addParticle(Vector3D(buf[3*i], buf[3*i+1], buf[3*i+2]), Vector3D(buf[3*n + 3*i], buf[3*n + 3*i+1], buf[3*n + 3*i+2]));
    ps->ps.back()->density = buf[6*n + i];
  }
  ps->simulate_time = t; ps->steps_taken = steps;
  if (ns) ps->setObstacleSpheres(sph);
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
     << "\tBackend: B200 CUDA (libpbf_b200), fp32, Jacobi XSPH" << std::endl;
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

Particles* load_particles_xml(const char* filename, std::string* error, const PbfParams* params, int device) {
  std::vector<double> pos, vel; double rho0 = 1000.0;
  if (!parse_particles_xml(filename, pos, vel, rho0, error)) return nullptr;
  Particles* particles = new Particles(rho0, params, device);
  const size_t n = pos.size() / 3;
  for (size_t i = 0; i < n; i++)
    particles->addParticle(Vector3D(pos[3*i], pos[3*i+1], pos[3*i+2]), Vector3D(vel[3*i], vel[3*i+1], vel[3*i+2]));
  particles->estimateDensities();
  return particles;
}

}  // namespace pbfhost
