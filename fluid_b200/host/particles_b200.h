// particles_b200.h — host-side adapter with the reference's `Particles` interface
// (SsnL/Fluid src/particles.h:105-140) on top of the C ABI (include/pbf_b200.h).
//
// The rest of the reference program touches the fluid only through this surface:
//   Application::load_particles  (application.cpp:302-344): Particles(rho0), addParticle, estimateDensities
//   PathTracer::fluid_simulate_* (pathtracer.cpp:444-480):  timeStep(), simulate_time
//   Particles::redraw            (particles.cpp:303-307):   ps[i]->getPosition(), getDensityBasedColor()
//   Particles::estimateDensityAt (particles.cpp:446-453):   marching cubes field
// so a maintainer swaps `#include "particles.h"` for this header.  The solver
// state lives on the GPU; `ps` is a host mirror refreshed by every timeStep().
#pragma once
#include <cstddef>
#include <string>
#include <vector>

#include "../../include/pbf_b200.h"
#include "../../include/pbf_b200_multi.h"

namespace pbfhost {

struct Vector3D {   // CGL::Vector3D's data layout (three doubles); only what the adapter needs
  double x, y, z;
  Vector3D() : x(0), y(0), z(0) {}
  Vector3D(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
};

struct Color { float r, g, b, a; };

class Particle {   // particles.h:19-91: the read-only part the visualiser / surfacer use
 public:
  // A VIEW: position / velocity / density live in the owner's contiguous AoS arrays (the buffers the C ABI reads and
  // the streaming read-back writes), so refreshing the mirror after a step costs nothing per particle.
  Vector3D& velocity;                                        // particles.h:21 (public member of the reference)
  Particle(Vector3D* p, Vector3D* v, const double* rho, double rho0) : velocity(*v), position(p), density(rho), rest_density(rho0) {}
  double getLatestDensityEstimate() const { return *density; }
  Vector3D getPosition() const { return *position; }
  Color getDensityBasedColor() const {   // particles.h:48-52
    double ratio = (*density - 0.8 * rest_density) / (0.4 * rest_density);
    double c = ratio < 0.0 ? 0.0 : (ratio > 1.0 ? 1.0 : ratio);
    return Color{1.0f, (float)(1.0 - c), (float)(1.0 - c), 1.0f};
  }
 private:
  friend struct Particles;
  Vector3D* position;
  const double* density;
  double rest_density;
};

struct Particles {
  std::vector<Particle*> ps;           // particles.h:106 (host mirror, original order)
  double simulate_time;                // particles.h:109
  const double rest_density;           // particles.h:110
  bool surfaceUpToTimestep = false;    // particles.h:112
  // particles.h:111 `std::vector<Primitive*> surface`: here the arguments of the reference's MarchingTriangle
  // constructor (p1 p2 p3 n1 n2 n3), which is what PathTracer::build_accel turns into primitives (pathtracer.cpp:248-253)
  struct SurfaceTriangle { Vector3D p1, p2, p3, n1, n2, n3; };
  std::vector<SurfaceTriangle> surface;
  // lattice of the surfacer: the reference hard-codes the Cornell box (particles.cpp:326-350)
  Vector3D surface_min = Vector3D(-1, 0, -1), surface_max = Vector3D(1, 1.5, 1);
  bool quiet = false;                  // the reference prints two lines per step (Q16); keep, but allow silence
  // Particle::initializeWithNewNeighbors (particles.cpp:165-173) warns on cerr about every particle with fewer than
  // NUM_NEIGHBOR_ALERT_THRESHOLD (18, particles.cpp:32) neighbours.  Kept (unless quiet): the first
  // neighbor_alert_max_lines lines are printed in the reference's wording and index order, then one summary line
  // (the reference would print one line per particle: 100 per step on p.xml, millions on a large free surface).
  int neighbor_alert_threshold = 18;
  size_t neighbor_alert_max_lines = 128;
  size_t neighbor_alerts_last_step = 0;   // how many particles were below the threshold in the last timeStep

  // quiet_ctor: the reference prints the parameter banner from the constructor (particles.h:114-116), i.e. before a
  // caller could set `quiet`
  explicit Particles(double rest_density = 1000.0, const PbfParams* params = nullptr, int device = 0, bool quiet_ctor = false);
  ~Particles();
  Particles(const Particles&) = delete;
  Particles& operator=(const Particles&) = delete;

  // Several GPUs of one box: call before the first step / estimateDensities.  With more than one device
  // the fluid is cut into x-slabs, one per device, behind the same object (pbf_create_multi, include/pbf_b200_multi.h):
  // halos and migrating particles travel as peer stores over NVLink, the slabs are re-balanced as the fluid flows, and
  // the result is bit-identical to one device.  An id may repeat (several slabs on one GPU; for testing).
  void setDevices(const std::vector<int>& device_ids);
  // The reference's `ps` always holds the latest state, so by default every timeStep() ends with the read-back into the
  // mirror (streamed behind the finalize kernels on one GPU; all slabs at once on several).  A windowless run that only
  // looks at the particles at the end (main.cpp -d: simulate, then render one frame) can switch that off and call
  // syncMirror() when it needs `ps`: the steps then run at the device rate (at 128M particles the read-back is 7 GB per step).
  bool mirror_each_step = true;
  void syncMirror();
  int numDevices() const { return devices_.empty() ? 1 : (int)devices_.size(); }
  void addParticle(Vector3D pos, Vector3D v);    // particles.h:118-120
  void timeStep(double delta_t);                 // particles.cpp:250-297 (delta_t must equal params.dt)
  void timeStep();                               // particles.cpp:299-301
  void estimateDensities();                      // particles.cpp:440-444
  // The reference collides against `BVHAccel* bvh` (particles.h:108, set by pathtracer.cpp:266).  Of that scene the
  // GPU step takes the box (PbfParams) and the StaticScene::Sphere primitives: rows (cx, cy, cz, r), at most
  // PBF_MAX_SPHERES.  May be called at any time; takes effect from the next timeStep().
  void setObstacleSpheres(const std::vector<double>& cx_cy_cz_r);
  // ... and its triangle primitives (meshes up to 2^22 triangles, device BVH): 18 doubles each, p1 p2 p3 n1 n2 n3 (pbf_set_obstacle_triangles)
  void setObstacleTriangles(const std::vector<double>& p1_p2_p3_n1_n2_n3);
  // Marching-cubes surface on the GPU (pbf_extract_surface): same triangles, same order as the reference's
  // getSurfacePrims (particles.cpp:352-391) / updateSurface (393-402: isolevel 0.95 rho0, step 0.5 H)
  std::vector<SurfaceTriangle> getSurfacePrims(double isolevel, double fStepSize);
  void updateSurface();
  double estimateDensityAt(Vector3D pos) const;  // particles.cpp:446-453 (host loop over the mirror, one point)
  // the same field for many points at once on the GPU (pbf_density_at): what a surfacer should call
  std::vector<double> estimateDensitiesAt(const std::vector<Vector3D>& points);
  // Restart file: magic "PBFCKPT2", int64 n, int64 steps, double simulate_time, double rho0,
  // int64 sizeof(PbfParams) + the struct, int64 number of obstacle spheres + rows, int64 number of obstacle
  // triangles + 18 doubles each, then pos[3n], vel[3n], density[n] as little-endian doubles in original particle
  // order ("PBFCKPT1" files, which had no triangle block, still load).  The device state is fp32 and its layout is
  // a pure function of (positions, velocities, ids), so a run continued from a checkpoint is bit-identical to
  // the uninterrupted run.
  bool saveCheckpoint(const char* filename, std::string* error = nullptr) const;
  static Particles* loadCheckpoint(const char* filename, std::string* error = nullptr, int device = 0, bool quiet = false,
                                   const std::vector<int>* devices = nullptr);   // nullptr on error
  long long steps_taken = 0;
  bool mirror_stale_ = false, readback_on_ = false;
  std::string paramsString() const;              // particles.cpp:420-438
  // the two numbers of the reference's "avg rho: a => b" line for the last step
  double avg_rho_first_iter = 0.0, avg_rho_final = 0.0;
  const PbfParams& params() const { return params_; }
  const char* lastError() const;

 private:
  void ensureUploaded();
  void refreshMirror(bool already_streamed = false);
  void rebind();
  void reportNeighborAlerts();
  PbfParams params_;
  int device_;
  pbf_handle* handle_ = nullptr;
  pbf_multi* multi_ = nullptr;         // set instead of handle_ when setDevices named more than one device
  std::vector<int> devices_;
  bool multi() const { return devices_.size() > 1; }
  // surfacer / density field in the multi-device case: a temporary single-device handle holding the mirror
  pbf_handle* scratchHandle();
  pbf_handle* scratch_ = nullptr; long long scratch_step_ = -1;
  bool uploaded_ = false;
  // the mirror: AoS xyz doubles in original particle order (Vector3D-compatible), what pbf_upload reads and the
  // streaming read-back writes; `ps` holds views into them (storage_ = the view objects, contiguous)
  std::vector<double> pos_, vel_, rho_, spheres_, tris_;
  std::vector<Particle> storage_;
};

// Application::load_particles (application.cpp:302-344): <particles><density>rho0</density><ps>
// <particle><pos>x y z</pos><v>x y z</v></particle>...  Density goes through float like stof (Q17).
// Streaming reader (no DOM), so multi-million-particle files are fine.  Returns nullptr on error.
Particles* load_particles_xml(const char* filename, std::string* error = nullptr, const PbfParams* params = nullptr, int device = 0, bool quiet = false,
                              const std::vector<int>* devices = nullptr);
inline Particles* load_checkpoint(const char* filename, std::string* error = nullptr, int device = 0, bool quiet = false, const std::vector<int>* devices = nullptr) {
  return Particles::loadCheckpoint(filename, error, device, quiet, devices);
}
// parse only: positions / velocities (AoS doubles) and rho0; false on error
bool parse_particles_xml(const char* filename, std::vector<double>& pos, std::vector<double>& vel, double& rho0, std::string* error);

}  // namespace pbfhost
