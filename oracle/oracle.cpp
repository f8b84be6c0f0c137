// oracle.cpp — CPU ORACLE for the path-tracing hot path of df07/mcp-raytracer.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under mcp_raytracer_b200/ may include, link, import or
// call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / the CPU baseline.
//
// What it is: a C++17 restatement of the reference's TypeScript algorithm, function by
// function, with the reference's numeric model: vectors live in gl-matrix 3.4.3
// Float32Array storage (every vector-producing op rounds to FP32 on store), scalars are
// FP64 JS numbers.  The reference itself cannot run in this
// image (no node / tsc / JS engine), so:
//
//   PARITY PINNING: the pure functions below are pinned against every known-answer vector
//   the reference's own Jest tests hold for this path (tests/test_oracle_reference_vectors.py
//   cites each tests/**/*.test.ts:line).  Converged pixel values, the adaptive-sampling exit
//   rule, gamma/quantisation and BVH tie ordering are asserted by NO reference test and no
//   reference render is reproducible (Math.random is unseedable): for those this oracle is
//   the authority and that part is "parity unpinned" (also stated in DESIGN.md).
//   gl-matrix is an un-vendored dependency (package.json:23, pinned 3.4.3 in
//   package-lock.json:3694-3698); its FP32-store behaviour is restated from its public API.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off: no FMA contraction, so a*b+c*d
// rounds like V8 does).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../include/rt_b200.h"

namespace orc {

static const double kInf = std::numeric_limits<double>::infinity();
static const double kPi = 3.141592653589793; // Math.PI

// ----------------------------------------------------------------------------------------
// Vec3 — src/geometry/vec3.ts:19-365 over gl-matrix vec3 (Float32Array(3)).
// ----------------------------------------------------------------------------------------
struct V3 {
  float x, y, z;
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
static inline float f32(double v) { return (float)v; }
// Vec3.create (vec3.ts:263-269): JS numbers stored into a Float32Array.
static inline V3 mk(double x, double y, double z) { return V3{f32(x), f32(y), f32(z)}; }
// vec3.ts:78-83 / gl-matrix add
static inline V3 add(V3 a, V3 b) { return mk((double)a.x + b.x, (double)a.y + b.y, (double)a.z + b.z); }
// vec3.ts:90-95
static inline V3 sub(V3 a, V3 b) { return mk((double)a.x - b.x, (double)a.y - b.y, (double)a.z - b.z); }
// vec3.ts:102-107: gl-matrix scale(out,a,s) with s a JS double
static inline V3 scale(V3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
// vec3.ts:114-119
static inline V3 mulv(V3 a, V3 b) { return mk((double)a.x * b.x, (double)a.y * b.y, (double)a.z * b.z); }
// vec3.ts:126-130: divide(t) = scale by 1/t
static inline V3 divs(V3 a, double t) { return scale(a, 1.0 / t); }
// vec3.ts:60-70: negate then canonicalise -0 -> +0
static inline V3 neg(V3 a) {
  V3 r{-a.x, -a.y, -a.z};
  if (r.x == 0) r.x = 0;
  if (r.y == 0) r.y = 0;
  if (r.z == 0) r.z = 0;
  return r;
}
// vec3.ts:132-136
static inline double len2(V3 v) { return (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z; }
// vec3.ts:138-140: gl-matrix length = Math.hypot
static inline double len(V3 v) { return std::hypot((double)v.x, (double)v.y, (double)v.z); }
// vec3.ts:152-157
static inline double dot(V3 a, V3 b) { return (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z; }
// vec3.ts:163-167 / gl-matrix cross
static inline V3 cross(V3 a, V3 b) {
  double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
  return mk(ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx);
}
// vec3.ts:228-232 / gl-matrix normalize: len=x²+y²+z²; if(len>0) len=1/sqrt(len); out=a*len
static inline V3 unit(V3 a) {
  double l = (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z;
  if (l > 0) l = 1.0 / std::sqrt(l);
  return mk(a.x * l, a.y * l, a.z * l);
}
// vec3.ts:174-186
static inline V3 reflect(V3 v, V3 n) {
  double dp = dot(v, n);
  V3 scaled = scale(n, 2 * dp);
  return sub(v, scaled);
}
// vec3.ts:193-209
static inline V3 refract(V3 v, V3 n, double eta) {
  double cosTheta = std::min(dot(neg(v), n), 1.0);
  V3 perp = scale(add(v, scale(n, cosTheta)), eta);
  V3 par = scale(n, -std::sqrt(std::fabs(1.0 - len2(perp))));
  return add(perp, par);
}
// vec3.ts:239-242
static inline double illuminance(V3 c) { return 0.299 * c.x + 0.587 * c.y + 0.114 * c.z; }

// ----------------------------------------------------------------------------------------
// Random numbers.  The reference draws from V8's Math.random (unseedable), in a fixed
// program order per path.  The oracle keeps that draw ORDER and offers two sources:
//   mode 0  "path-keyed Philox": Philox4x32-7 (seven rounds: the fewest that are crush-resistant, Salmon et al. SC'11;
//           ten until round 2 — the count is the CUDA path's RT_PHILOX_ROUNDS), key=(pixel index, sample index),
//           counter=(block, stream, seed_lo, seed_hi); stream b serves the rayColor call
//           entered with stats.bounces == b, and the camera-ray draws of a path are the first
//           draws of its stream 0.  A block yields FIVE 24-bit uniforms u*2^-24: the high 24
//           bits of each of the four words, then one assembled from the low bytes of words
//           0..2; draw i of a stream is uniform i%5 of block i/5.  The CUDA path uses the
//           same streams, so a GPU path and an oracle path see the same numbers.
//   mode 1  sequential xorshift128+ (one stream per render strip, like one Math.random per
//           worker thread), for independence checks.
// ----------------------------------------------------------------------------------------
#ifndef ORC_PHILOX_ROUNDS
#define ORC_PHILOX_ROUNDS 7
#endif
static const int kPhiloxRounds = ORC_PHILOX_ROUNDS;
static inline void philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4], int rounds) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < rounds; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
  int mode = 0;
  uint64_t seed = 0;
  // mode 0
  uint32_t key[2] = {0, 0};
  uint32_t stream = 0, idx = 0;
  uint32_t buf[5] = {0, 0, 0, 0, 0};
  // mode 1
  uint64_t s0 = 1, s1 = 2;
  uint64_t draws = 0;
  // mode 2: the caller supplies the sequence Math.random() returns (per-function parity hooks); past its end: 0.5
  const double* list = nullptr;
  uint64_t list_n = 0;

  void seed_sequential(uint64_t sd) {
    // splitmix64 expansion of the seed
    auto sm = [](uint64_t& x) {
      uint64_t z = (x += 0x9E3779B97F4A7C15ull);
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      return z ^ (z >> 31);
    };
    uint64_t x = sd;
    s0 = sm(x);
    s1 = sm(x);
    if (!s0 && !s1) s1 = 1;
  }
  void begin_path(uint32_t pixel, uint32_t sample) {
    key[0] = pixel;
    key[1] = sample;
    begin_stream(0);
  }
  void begin_stream(uint32_t s) {
    stream = s;
    idx = 0;
  }
  double next() {
    ++draws;
    if (mode == 2) return draws <= list_n ? list[draws - 1] : 0.5;
    if (mode == 0) {
      if (idx % 5u == 0) {
        uint32_t ctr[4] = {idx / 5u, stream, (uint32_t)seed, (uint32_t)(seed >> 32)};
        uint32_t w[4];
        philox4x32(ctr, key, w, kPhiloxRounds);
        for (int k = 0; k < 4; ++k) buf[k] = w[k] >> 8;
        buf[4] = ((w[0] & 0xffu) << 16) | ((w[1] & 0xffu) << 8) | (w[2] & 0xffu);
      }
      uint32_t v = buf[idx % 5u];
      ++idx;
      return (double)v * (1.0 / 16777216.0);
    }
    uint64_t a = s0, b = s1;
    s0 = b;
    a ^= a << 23;
    a ^= a >> 17;
    a ^= b ^ (b >> 26);
    s1 = a;
    return (double)((s0 + s1) >> 11) * (1.0 / 9007199254740992.0);
  }
};

// vec3.ts:272-278: Vec3.random(min,max)
static inline V3 randomVec(Rng& g, double mn, double mx) {
  double a = mn + (mx - mn) * g.next();
  double b = mn + (mx - mn) * g.next();
  double c = mn + (mx - mn) * g.next();
  return mk(a, b, c);
}
// vec3.ts:285-292
static inline V3 randomInUnitSphere(Rng& g) {
  for (;;) {
    V3 p = randomVec(g, -1, 1);
    if (len2(p) < 1) return p;
  }
}
// vec3.ts:325-337
static inline V3 randomCosineDirection(Rng& g) {
  double r1 = g.next(), r2 = g.next();
  double phi = 2 * kPi * r1;
  double s = std::sqrt(r2);
  return mk(std::cos(phi) * s, std::sin(phi) * s, std::sqrt(1 - r2));
}
// vec3.ts:345-351
static inline V3 randomToSphere(Rng& g, double radius, double distanceSquared) {
  double r1 = g.next(), r2 = g.next();
  double z = 1 + r2 * (std::sqrt(1 - radius * radius / distanceSquared) - 1);
  double phi = 2 * kPi * r1;
  return mk(std::cos(phi) * std::sqrt(1 - z * z), std::sin(phi) * std::sqrt(1 - z * z), z);
}
// vec3.ts:357-364
static inline V3 randomInUnitDisk(Rng& g) {
  for (;;) {
    double a = 2 * g.next() - 1;
    double b = 2 * g.next() - 1;
    V3 p = mk(a, b, 0);
    if (len2(p) < 1) return p;
  }
}

// ----------------------------------------------------------------------------------------
// Ray / Interval / AABB — src/geometry/ray.ts, interval.ts, aabb.ts
// ----------------------------------------------------------------------------------------
struct Ray {
  V3 o, d;
  V3 at(double t) const { return add(o, scale(d, t)); } // ray.ts:25-28
};
struct Interval {
  double mn, mx;
  bool surrounds(double x) const { return mn < x && x < mx; } // interval.ts:51-53 (strict)
  // the rest of the type (interval.ts:33-64); only `surrounds` is called on the render path
  double size() const { return mx - mn; }
  bool contains(double x) const { return mn <= x && x <= mx; }
  double clamp(double x) const { return x < mn ? mn : (x > mx ? mx : x); }
};
struct AABB {
  V3 mn, mx;
};
static inline AABB emptyBox() { return AABB{mk(kInf, kInf, kInf), mk(-kInf, -kInf, -kInf)}; } // aabb.ts:92-97
// aabb.ts:68-80
static inline AABB surroundingBox(const AABB& a, const AABB& b) {
  return AABB{mk(std::min((double)a.mn.x, (double)b.mn.x), std::min((double)a.mn.y, (double)b.mn.y),
                 std::min((double)a.mn.z, (double)b.mn.z)),
              mk(std::max((double)a.mx.x, (double)b.mx.x), std::max((double)a.mx.y, (double)b.mx.y),
                 std::max((double)a.mx.z, (double)b.mx.z))};
}

// Event counters for the algorithmic-work model.
struct Counters {
  uint64_t rays = 0, box_tests = 0;
  uint64_t sphere_miss = 0, sphere_hit = 0;
  uint64_t planar_treject = 0, quad_outside = 0, quad_hit = 0, plane_hit = 0;
  uint64_t rr = 0, background = 0, hits = 0;
  uint64_t lambert = 0, metal = 0, metal_fuzz0 = 0, dielectric = 0, light_pdf_evals = 0;
  uint64_t paths = 0, defocus = 0;
  void operator+=(const Counters& o) {
    const uint64_t* s = (const uint64_t*)&o;
    uint64_t* d = (uint64_t*)this;
    for (size_t i = 0; i < sizeof(Counters) / 8; ++i) d[i] += s[i];
  }
};
static thread_local Counters* tl_cnt = nullptr;
#define CNT(f)                 \
  do {                         \
    if (tl_cnt) ++tl_cnt->f;   \
  } while (0)

struct NoCount { // suspends event counting for the enclosing scope
  Counters* saved;
  NoCount() : saved(tl_cnt) { tl_cnt = nullptr; }
  ~NoCount() { tl_cnt = saved; }
};

// aabb.ts:30-59 — per-axis test, each axis against the ORIGINAL interval.
static inline bool aabbHit(const AABB& b, const Ray& r, Interval rayT) {
  CNT(box_tests);
  for (int a = 0; a < 3; ++a) {
    double invD = 1.0 / (double)r.d[a];
    double t0 = ((double)b.mn[a] - (double)r.o[a]) * invD;
    double t1 = ((double)b.mx[a] - (double)r.o[a]) * invD;
    if (invD < 0) std::swap(t0, t1);
    double tMin = t0 > rayT.mn ? t0 : rayT.mn;
    double tMax = t1 < rayT.mx ? t1 : rayT.mx;
    if (tMax <= tMin) return false;
  }
  return true;
}

// ----------------------------------------------------------------------------------------
// Hittables — src/geometry/hittable.ts, src/entities/{sphere,plane,quad}.ts
// ----------------------------------------------------------------------------------------
struct Material;
struct HitRecord {
  V3 p, normal;
  double t;
  bool frontFace;
  const Material* material;
  int objId;
};

struct Hittable {
  virtual ~Hittable() {}
  virtual bool hit(const Ray& r, Interval rayT, HitRecord& rec) const = 0;
  virtual AABB boundingBox() const = 0;
  // PDFHittable (hittable.ts:52-72); only Sphere and Quad have `pdf`
  virtual bool hasPdf() const { return false; }
  virtual double pdfValue(V3, V3) const { return 0; }
  virtual V3 pdfRandomVec(V3, Rng&) const { return V3{0, 0, 0}; }
};

struct ONB { // src/geometry/onbasis.ts:18-51
  V3 u, v, w;
  explicit ONB(V3 n) {
    w = unit(n);
    V3 a = std::fabs((double)w.x) > 0.9 ? mk(0, 1, 0) : mk(1, 0, 0);
    v = unit(cross(w, a));
    u = cross(w, v);
  }
  V3 local(V3 a) const { return add(add(scale(u, a.x), scale(v, a.y)), scale(w, a.z)); }
};

struct Sphere : Hittable { // src/entities/sphere.ts
  V3 center;
  double radius;
  const Material* material;
  int objId;
  AABB box;
  Sphere(V3 c, double r, const Material* m, int id) : center(c), radius(r), material(m), objId(id) {
    V3 rv = mk(r, r, r); // sphere.ts:26-29 (inverted when r<0)
    box = AABB{sub(center, rv), add(center, rv)};
  }
  bool hit(const Ray& r, Interval rayT, HitRecord& rec) const override { // sphere.ts:45-85
    V3 oc = sub(r.o, center);
    double a = len2(r.d);
    double halfB = dot(oc, r.d);
    double c = len2(oc) - radius * radius;
    double disc = halfB * halfB - a * c;
    if (disc < 0) { CNT(sphere_miss); return false; }
    double sqrtd = std::sqrt(disc);
    double root = (-halfB - sqrtd) / a;
    if (!rayT.surrounds(root)) {
      root = (-halfB + sqrtd) / a;
      if (!rayT.surrounds(root)) { CNT(sphere_miss); return false; }
    }
    CNT(sphere_hit);
    V3 p = r.at(root);
    V3 n = divs(sub(p, center), radius);
    bool front = dot(r.d, n) <= 0;
    if (!front) n = neg(n);
    rec = HitRecord{p, n, root, front, material, objId};
    return true;
  }
  AABB boundingBox() const override { return box; }
  bool hasPdf() const override { return true; }
  double pdfValue(V3 origin, V3 direction) const override { // sphere.ts:106-131
    HitRecord rec;
    NoCount nc; // a light-pdf evaluation is not a traced ray
    if (!hit(Ray{origin, direction}, Interval{0.001, kInf}, rec)) return 0;
    double d2 = len2(sub(center, origin));
    if (d2 <= radius * radius) return 1.0 / (4.0 * kPi);
    double cosTheta = std::sqrt(1 - radius * radius / d2);
    double solidAngle = 2 * kPi * (1 - cosTheta);
    return 1 / solidAngle;
  }
  V3 pdfRandomVec(V3 origin, Rng& g) const override { // sphere.ts:140-147
    V3 oc = sub(center, origin);
    double d2 = len2(oc);
    ONB uvw(oc);
    return uvw.local(randomToSphere(g, radius, d2));
  }
};

struct Plane : Hittable { // src/entities/plane.ts
  V3 q, u, v, normal, inverseNormal, w;
  double d;
  const Material* material;
  int objId;
  AABB box;
  Plane(V3 q_, V3 u_, V3 v_, const Material* m, int id) : q(q_), u(u_), v(v_), material(m), objId(id) {
    V3 cp = cross(u, v); // plane.ts:33-42
    normal = unit(cp);
    inverseNormal = neg(normal);
    d = dot(normal, q);
    w = divs(cp, len2(cp));
    box = computeBox();
  }
  AABB computeBox() const { // plane.ts:122-154
    const double eps = 1e-4;
    if (std::fabs((double)normal.x) > 0.9999) {
      double px = d / normal.x;
      return AABB{mk(px - eps, -kInf, -kInf), mk(px + eps, kInf, kInf)};
    } else if (std::fabs((double)normal.y) > 0.9999) {
      double py = d / normal.y;
      return AABB{mk(-kInf, py - eps, -kInf), mk(kInf, py + eps, kInf)};
    } else if (std::fabs((double)normal.z) > 0.9999) {
      double pz = d / normal.z;
      return AABB{mk(-kInf, -kInf, pz - eps), mk(kInf, kInf, pz + eps)};
    }
    return AABB{mk(-kInf, -kInf, -kInf), mk(kInf, kInf, kInf)};
  }
  // plane.ts:55-77
  bool intersect(const Ray& r, Interval rayT, double& t, double& alpha, double& beta) const {
    double denom = dot(normal, r.d);
    if (std::fabs(denom) < 1e-8) { CNT(planar_treject); return false; }
    t = (d - dot(normal, r.o)) / denom;
    if (!rayT.surrounds(t)) { CNT(planar_treject); return false; }
    V3 ip = r.at(t);
    V3 hp = sub(ip, q);
    alpha = dot(w, cross(hp, v));
    beta = dot(w, cross(u, hp));
    return true;
  }
  bool hit(const Ray& r, Interval rayT, HitRecord& rec) const override { // plane.ts:86-106
    double t, a, b;
    if (!intersect(r, rayT, t, a, b)) return false;
    CNT(plane_hit);
    V3 p = r.at(t);
    bool front = dot(r.d, normal) <= 0;
    rec = HitRecord{p, front ? normal : inverseNormal, t, front, material, objId};
    return true;
  }
  AABB boundingBox() const override { return box; }
};

struct Quad : Hittable { // src/entities/quad.ts
  Plane plane;
  V3 q, u, v;
  double area;
  const Material* material;
  int objId;
  AABB box;
  Quad(V3 q_, V3 u_, V3 v_, const Material* m, int id)
      : plane(q_, u_, v_, m, id), q(q_), u(u_), v(v_), material(m), objId(id) {
    area = len(cross(u, v)); // quad.ts:37
    // quad.ts:92-114
    V3 v1 = q, v2 = add(q, u), v3 = add(q, v), v4 = add(add(q, u), v);
    auto mn4 = [](double a, double b, double c, double d) { return std::min(std::min(a, b), std::min(c, d)); };
    auto mx4 = [](double a, double b, double c, double d) { return std::max(std::max(a, b), std::max(c, d)); };
    const double eps = 1e-4;
    box = AABB{mk(mn4(v1.x, v2.x, v3.x, v4.x) - eps, mn4(v1.y, v2.y, v3.y, v4.y) - eps, mn4(v1.z, v2.z, v3.z, v4.z) - eps),
               mk(mx4(v1.x, v2.x, v3.x, v4.x) + eps, mx4(v1.y, v2.y, v3.y, v4.y) + eps, mx4(v1.z, v2.z, v3.z, v4.z) + eps)};
  }
  bool hit(const Ray& r, Interval rayT, HitRecord& rec) const override { // quad.ts:50-76
    double t, alpha, beta;
    if (!plane.intersect(r, rayT, t, alpha, beta)) return false;
    if (alpha < 0 || alpha > 1 || beta < 0 || beta > 1) { CNT(quad_outside); return false; }
    CNT(quad_hit);
    V3 p = r.at(t);
    bool front = dot(r.d, plane.normal) <= 0;
    V3 n = front ? plane.normal : neg(plane.normal);
    rec = HitRecord{p, n, t, front, material, objId};
    return true;
  }
  AABB boundingBox() const override { return box; }
  bool hasPdf() const override { return true; }
  double pdfValue(V3 origin, V3 direction) const override { // quad.ts:123-140
    HitRecord rec;
    NoCount nc;
    if (!hit(Ray{origin, direction}, Interval{0.001, kInf}, rec)) return 0;
    double d2 = len2(sub(rec.p, origin));
    double cosine = std::fabs(dot(direction, rec.normal));
    return d2 / (area * cosine);
  }
  V3 pdfRandomVec(V3 origin, Rng& g) const override { // quad.ts:148-158
    double alpha = g.next();
    double beta = g.next();
    V3 rp = add(add(q, scale(u, alpha)), scale(v, beta));
    return unit(sub(rp, origin));
  }
};

struct HittableList : Hittable { // src/geometry/hittableList.ts
  std::vector<const Hittable*> objects;
  AABB boundingBox() const override { // :34-56
    if (objects.empty()) return emptyBox();
    AABB r = objects[0]->boundingBox();
    for (size_t i = 1; i < objects.size(); ++i) r = surroundingBox(r, objects[i]->boundingBox());
    return r;
  }
  bool hit(const Ray& r, Interval rayT, HitRecord& rec) const override { // :71-87
    bool any = false;
    Interval iv = rayT;
    HitRecord tmp;
    for (const Hittable* o : objects) {
      if (o->hit(r, iv, tmp)) {
        iv.mx = tmp.t;
        rec = tmp;
        any = true;
      }
    }
    return any;
  }
};

struct EmptyHittable : Hittable { // bvh.ts:8-11
  bool hit(const Ray&, Interval, HitRecord&) const override { return false; }
  AABB boundingBox() const override { return emptyBox(); }
};
static const EmptyHittable kEmpty;

struct BVHNode : Hittable { // src/geometry/bvh.ts
  const Hittable* left = nullptr;
  const Hittable* right = nullptr;
  AABB box;
  std::vector<std::unique_ptr<Hittable>> owned;
  static bool compareBoxes(const Hittable* a, const Hittable* b, int axis) { // :112-117
    return a->boundingBox().mn[axis] < b->boundingBox().mn[axis];
  }
  BVHNode(const std::vector<const Hittable*>& objects, size_t start, size_t end) { // :34-102
    std::vector<const Hittable*> list(objects.begin() + start, objects.begin() + end);
    bool have = false;
    AABB nb = emptyBox();
    for (const Hittable* o : list) {
      AABB b = o->boundingBox();
      nb = have ? surroundingBox(nb, b) : b;
      have = true;
    }
    double xe = (double)nb.mx.x - (double)nb.mn.x;
    double ye = (double)nb.mx.y - (double)nb.mn.y;
    double ze = (double)nb.mx.z - (double)nb.mn.z;
    int axis = 0;
    if (ye > xe && ye > ze) axis = 1;
    else if (ze > xe && ze > ye) axis = 2;
    size_t span = end - start;
    if (span == 1) {
      left = list[0];
      right = &kEmpty;
    } else if (span == 2) {
      if (compareBoxes(list[0], list[1], axis)) { left = list[0]; right = list[1]; }
      else { left = list[1]; right = list[0]; }
    } else if (span <= 4) {
      auto* ll = new HittableList();
      for (const Hittable* o : list) ll->objects.push_back(o);
      owned.emplace_back(ll);
      left = ll;
      right = &kEmpty;
    } else {
      // objectsList.sort((a,b) => compareBoxes(a,b,axis) ? -1 : 1): V8's TimSort only ever
      // asks "order < 0", so a comparator that answers 1 for equal keys sorts exactly like
      // a stable sort on `<`.
      std::stable_sort(list.begin(), list.end(),
                       [axis](const Hittable* a, const Hittable* b) { return compareBoxes(a, b, axis); });
      size_t mid = span / 2;
      auto* l = new BVHNode(list, 0, mid);
      auto* r = new BVHNode(list, mid, span);
      owned.emplace_back(l);
      owned.emplace_back(r);
      left = l;
      right = r;
    }
    box = surroundingBox(left->boundingBox(), right->boundingBox());
  }
  bool hit(const Ray& r, Interval rayT, HitRecord& rec) const override { // :128-146
    if (!aabbHit(box, r, rayT)) return false;
    bool hl = left->hit(r, rayT, rec);
    Interval ri = hl ? Interval{rayT.mn, rec.t} : rayT;
    HitRecord rr;
    bool hr = right->hit(r, ri, rr);
    if (hr) rec = rr;
    return hr || hl;
  }
  AABB boundingBox() const override { return box; }
};

// ----------------------------------------------------------------------------------------
// PDFs — src/geometry/pdf.ts
// ----------------------------------------------------------------------------------------
struct CosinePDF { // pdf.ts:32-51
  ONB uvw;
  explicit CosinePDF(V3 w) : uvw(w) {}
  double value(V3 direction) const {
    double c = dot(unit(direction), uvw.w);
    return c <= 0 ? 0 : c / kPi;
  }
  V3 generate(Rng& g) const { return uvw.local(randomCosineDirection(g)); }
};

// ----------------------------------------------------------------------------------------
// Materials — src/materials/*.ts
// ----------------------------------------------------------------------------------------
struct ScatterResult { // material.ts:14-23 (+ dielectric.ts:7-9 `reflected`)
  V3 attenuation;
  bool hasScattered = false;
  Ray scattered;
  bool hasPdf = false;
  V3 pdfNormal; // CosinePDF(rec.normal)
  bool reflected = false;
};
struct Material {
  virtual ~Material() {}
  virtual bool scatter(const Ray&, const HitRecord&, Rng&, ScatterResult&) const { return false; } // material.ts:50-52
  virtual V3 emitted(const HitRecord&) const { return mk(0, 0, 0); }                                 // material.ts:54-56
};
struct Lambertian : Material { // lambertian.ts:26-31
  V3 albedo;
  explicit Lambertian(V3 a) : albedo(a) {}
  bool scatter(const Ray&, const HitRecord& rec, Rng&, ScatterResult& out) const override {
    CNT(lambert);
    out = ScatterResult();
    out.attenuation = albedo;
    out.hasPdf = true;
    out.pdfNormal = rec.normal;
    return true;
  }
};
struct Metal : Material { // metal.ts
  V3 albedo;
  double fuzz;
  Metal(V3 a, double f) : albedo(a), fuzz(f < 1 ? std::max(0.0, f) : 1) {} // :20
  bool scatter(const Ray& rIn, const HitRecord& rec, Rng& g, ScatterResult& out) const override { // :29-50
    if (fuzz > 0) CNT(metal); else CNT(metal_fuzz0);
    V3 reflected = reflect(unit(rIn.d), rec.normal);
    V3 fr = fuzz > 0 ? add(reflected, scale(randomInUnitSphere(g), fuzz)) : reflected;
    if (dot(fr, rec.normal) <= 0) return false;
    out = ScatterResult();
    out.attenuation = albedo;
    out.hasScattered = true;
    out.scattered = Ray{rec.p, fr};
    return true;
  }
};
struct Dielectric : Material { // dielectric.ts
  double ior;
  explicit Dielectric(double i) : ior(i) {}
  static double reflectance(double cosine, double ratio) { // :93-98
    double r0 = (1 - ratio) / (1 + ratio);
    r0 = r0 * r0;
    return r0 + (1 - r0) * std::pow(1 - cosine, 5);
  }
  bool scatter(const Ray& rIn, const HitRecord& rec, Rng& g, ScatterResult& out) const override { // :44-84
    CNT(dielectric);
    double ratio = rec.frontFace ? (1.0 / ior) : ior;
    V3 ud = unit(rIn.d);
    double cosTheta = std::min(dot(neg(ud), rec.normal), 1.0);
    double sinTheta = std::sqrt(1.0 - cosTheta * cosTheta);
    bool cannotRefract = ratio * sinTheta > 1.0;
    V3 dir;
    bool refl;
    if (cannotRefract || reflectance(cosTheta, ratio) > g.next()) {
      dir = reflect(ud, rec.normal);
      refl = true;
    } else {
      dir = refract(ud, rec.normal, ratio);
      refl = false;
    }
    out = ScatterResult();
    out.attenuation = mk(1, 1, 1);
    out.hasScattered = true;
    out.scattered = Ray{rec.p, dir};
    out.reflected = refl;
    return true;
  }
};
struct DiffuseLight : Material { // diffuseLight.ts:29-31
  V3 emit;
  explicit DiffuseLight(V3 e) : emit(e) {}
  V3 emitted(const HitRecord&) const override { return emit; }
};
struct LayeredMaterial : Material { // layeredMaterial.ts
  std::unique_ptr<Dielectric> outer;
  const Material* inner;
  LayeredMaterial(Dielectric* o, const Material* i) : outer(o), inner(i) {}
  bool scatter(const Ray& rIn, const HitRecord& rec, Rng& g, ScatterResult& out) const override { // :36-53
    ScatterResult o;
    if (!outer->scatter(rIn, rec, g, o)) return false;
    if (o.reflected) { out = o; return true; }
    return inner->scatter(o.scattered, rec, g, out);
  }
  V3 emitted(const HitRecord& rec) const override { return inner->emitted(rec); } // :61-63
};
struct MixedMaterial : Material { // mixedMaterial.ts
  const Material *m1, *m2;
  double weight;
  MixedMaterial(const Material* a, const Material* b, double w) : m1(a), m2(b), weight(std::max(0.0, std::min(1.0, w))) {} // :29
  bool scatter(const Ray& rIn, const HitRecord& rec, Rng& g, ScatterResult& out) const override { // :38-45
    if (g.next() < weight) return m1->scatter(rIn, rec, g, out);
    return m2->scatter(rIn, rec, g, out);
  }
  V3 emitted(const HitRecord& rec) const override { // :52-57
    V3 e1 = scale(m1->emitted(rec), weight);
    V3 e2 = scale(m2->emitted(rec), 1.0 - weight);
    return add(e1, e2);
  }
};

// ----------------------------------------------------------------------------------------
// PixelStats / RenderStats — src/render-utils/renderStats.ts
// ----------------------------------------------------------------------------------------
struct PixelStats {
  V3 color = mk(0, 0, 0);
  int samples = 0;
  long long bounces = 0;
  double minBounces = kInf, maxBounces = 0;
  double sumIll = 0, sumIll2 = 0;
  double m1[3] = {0, 0, 0}, m2[3] = {0, 0, 0}; // oracle-only: exact per-channel moments
  void addSample(V3 rayColor, int b, bool calcIll) { // :76-88
    color = add(color, rayColor);
    samples++;
    bounces += b;
    minBounces = std::min(minBounces, (double)b);
    maxBounces = std::max(maxBounces, (double)b);
    if (calcIll) {
      double il = illuminance(rayColor);
      sumIll += il;
      sumIll2 += il * il;
    }
    double c[3] = {rayColor.x, rayColor.y, rayColor.z};
    for (int k = 0; k < 3; ++k) { m1[k] += c[k]; m2[k] += c[k] * c[k]; }
  }
};
struct RenderStats {
  uint64_t pixels = 0, samplesTotal = 0, bouncesTotal = 0;
  double samplesMin = kInf, samplesMax = 0, bouncesMin = kInf, bouncesMax = 0;
  void addPixel(const PixelStats& p) { // :21-35
    pixels++;
    samplesTotal += p.samples;
    samplesMin = std::min(samplesMin, (double)p.samples);
    samplesMax = std::max(samplesMax, (double)p.samples);
    bouncesTotal += p.bounces;
    bouncesMin = std::min(bouncesMin, p.minBounces);
    bouncesMax = std::max(bouncesMax, p.maxBounces);
  }
  void merge(const RenderStats& s) { // :42-64
    pixels += s.pixels;
    samplesTotal += s.samplesTotal;
    samplesMin = std::min(samplesMin, s.samplesMin);
    samplesMax = std::max(samplesMax, s.samplesMax);
    bouncesTotal += s.bouncesTotal;
    bouncesMin = std::min(bouncesMin, s.bouncesMin);
    bouncesMax = std::max(bouncesMax, s.bouncesMax);
  }
};

// ----------------------------------------------------------------------------------------
// Camera — src/camera.ts
// ----------------------------------------------------------------------------------------
struct Camera {
  // options (camera.ts:73-83 after the merge done by the host)
  int width, samples, depth, aBatch, rouletteDepth, mode;
  double aspect, aTolerance;
  bool roulette;
  int imageWidth, imageHeight;
  V3 center, pixel00Loc, pixelDeltaU, pixelDeltaV, u, v, w, defocusDiskU, defocusDiskV;
  V3 bgTop, bgBottom;
  double aperture, focusDistance;
  bool useAdaptiveSampling;
  const Hittable* world = nullptr;
  std::vector<const Hittable*> lights;

  // owned scene
  std::vector<std::unique_ptr<Material>> materials; // one tree per object, like scenes.ts:113
  std::vector<std::unique_ptr<Hittable>> objects;
  std::unique_ptr<BVHNode> bvh;

  void init(const rt_camera_desc& c, const rt_render_opts& o) { // camera.ts:107-166
    width = o.width; samples = o.samples; depth = o.depth; aBatch = o.a_batch;
    rouletteDepth = o.roulette_depth; mode = o.mode; aspect = o.aspect; aTolerance = o.a_tolerance;
    roulette = o.roulette != 0;
    imageWidth = width;
    imageHeight = (int)std::ceil(imageWidth / aspect);
    V3 from = mk(c.from[0], c.from[1], c.from[2]);
    V3 at = mk(c.at[0], c.at[1], c.at[2]);
    V3 up = mk(c.up[0], c.up[1], c.up[2]);
    center = from;
    bgTop = mk(c.background_top[0], c.background_top[1], c.background_top[2]);
    bgBottom = mk(c.background_bottom[0], c.background_bottom[1], c.background_bottom[2]);
    aperture = c.aperture;
    focusDistance = c.focus != 0 && !std::isnan(c.focus) ? c.focus : len(sub(from, at)); // `focus || ...`
    double theta = c.vfov * (kPi / 180);
    double h = std::tan(theta / 2);
    double viewportHeight = 2 * h * focusDistance;
    double aspectRatio = (double)imageWidth / imageHeight;
    double viewportWidth = viewportHeight * aspectRatio;
    w = unit(sub(from, at));
    u = unit(cross(up, w));
    v = cross(w, u);
    V3 viewportU = scale(u, viewportWidth);
    V3 viewportV = scale(v, -viewportHeight);
    pixelDeltaU = divs(viewportU, imageWidth);
    pixelDeltaV = divs(viewportV, imageHeight);
    V3 halfU = divs(viewportU, 2);
    V3 halfV = divs(viewportV, 2);
    V3 upperLeft = sub(sub(sub(center, scale(w, focusDistance)), halfU), halfV);
    pixel00Loc = add(upperLeft, scale(add(pixelDeltaU, pixelDeltaV), 0.5));
    defocusDiskU = scale(u, aperture / 2);
    defocusDiskV = scale(v, aperture / 2);
    useAdaptiveSampling = aTolerance > 0 && samples > 1;
  }

  Ray getRay(int i, int j, Rng& g, bool jitterAndDefocus = true) const { // camera.ts:176-210
    V3 pixelCenter = add(add(pixel00Loc, scale(pixelDeltaU, i)), scale(pixelDeltaV, j));
    V3 pixelSample = pixelCenter;
    if (jitterAndDefocus && samples > 1) {
      double px = -0.5 + g.next();
      double py = -0.5 + g.next();
      pixelSample = add(add(pixelCenter, scale(pixelDeltaU, px)), scale(pixelDeltaV, py));
    }
    V3 origin = center;
    V3 dir = sub(pixelSample, center);
    if (jitterAndDefocus && aperture > 0) {
      CNT(defocus);
      V3 rd = randomInUnitDisk(g);
      V3 offset = add(scale(defocusDiskU, rd.x), scale(defocusDiskV, rd.y));
      origin = add(center, offset);
      dir = sub(pixelSample, origin);
    }
    return Ray{origin, dir};
  }

  // The diffuse branch of rayColor (camera.ts:285-308): direction from the mixture pdf, its value, and the
  // scatter pdf's value for that direction.  One function so that rayColor and the parity hook share it.
  void diffuseBounce(V3 p, V3 pdfNormal, Rng& g, V3& direction, double& pdfValue, double& spv) const {
    CosinePDF cpdf(pdfNormal);
    // MixturePDF([scatterPdf, ...lights], [0.5, 0.5/n ...]) — camera.ts:287-288, pdf.ts:57-99
    size_t n = lights.size();
    double totalWeight = 0;
    totalWeight += 0.5;
    for (size_t k = 0; k < n; ++k) totalWeight += 0.5 / n;
    // generate
    {
      double rnd = g.next() * totalWeight;
      double partial = 0.5;
      int chosen = -1; // -1 = cosine
      bool found = rnd < partial;
      if (!found) {
        for (size_t k = 0; k < n; ++k) {
          partial += 0.5 / n;
          if (rnd < partial) { chosen = (int)k; found = true; break; }
        }
        if (!found) chosen = n ? (int)n - 1 : -1; // fallback to last PDF
      }
      direction = chosen < 0 ? cpdf.generate(g) : lights[chosen]->pdfRandomVec(p, g);
    }
    double sum = 0;
    sum += 0.5 * cpdf.value(direction);
    for (size_t k = 0; k < n; ++k) {
      CNT(light_pdf_evals);
      sum += (0.5 / n) * lights[k]->pdfValue(p, direction);
    }
    pdfValue = sum / totalWeight;
    spv = cpdf.value(direction);
  }

  V3 rayColor(const Ray& r, V3 throughput, int& bounces, Rng& g) const { // camera.ts:221-319
    if (bounces > 0) g.begin_stream((uint32_t)bounces); // stream 0 continues after the camera-ray draws
    if (bounces >= depth) return mk(0, 0, 0);
    if (roulette && bounces >= rouletteDepth) {
      CNT(rr);
      double mc = std::max(std::max((double)throughput.x, (double)throughput.y), (double)throughput.z);
      double p = std::min(mc, 0.95);
      if (g.next() > p) return mk(0, 0, 0);
      throughput = divs(throughput, p);
    }
    CNT(rays);
    HitRecord rec;
    if (!world->hit(r, Interval{0.001, kInf}, rec)) {
      CNT(background);
      V3 ud = unit(r.d);
      double a = 0.5 * (ud.y + 1.0);
      return mulv(add(scale(bgTop, 1.0 - a), scale(bgBottom, a)), throughput);
    }
    CNT(hits);
    V3 emitted = mulv(rec.material->emitted(rec), throughput);
    ScatterResult sr;
    if (!rec.material->scatter(r, rec, g, sr)) return emitted;
    bounces++;
    if (sr.hasScattered) {
      V3 nt = mulv(throughput, sr.attenuation);
      V3 sc = rayColor(sr.scattered, nt, bounces, g);
      return add(emitted, sc);
    }
    if (sr.hasPdf) {
      V3 direction;
      double pdfValue, spv;
      diffuseBounce(rec.p, sr.pdfNormal, g, direction, pdfValue, spv);
      Ray scattered{rec.p, direction};
      if (pdfValue <= 0.0001) return emitted;
      V3 brdf = scale(sr.attenuation, spv);
      V3 nt = divs(mulv(throughput, brdf), pdfValue);
      V3 inc = rayColor(scattered, nt, bounces, g);
      return add(emitted, inc);
    }
    return emitted;
  }

  V3 finalColor(const PixelStats& p) const { // camera.ts:326-340
    if (mode == RT_MODE_BOUNCES) {
      double avg = p.samples > 0 ? (double)p.bounces / p.samples : 0;
      return mk(0, 0, std::min(avg / depth, 1.0));
    }
    if (mode == RT_MODE_SAMPLES) return mk(std::min((double)p.samples / samples, 1.0), 0, 0);
    return divs(p.color, p.samples);
  }
  bool pixelConverged(const PixelStats& p) const { // camera.ts:348-368
    if (aTolerance <= 0 || samples <= 1 || p.samples < 2) return false;
    if (p.samples % aBatch != 0) return false;
    double mean = p.sumIll / p.samples;
    double variance = (p.sumIll2 - (p.sumIll * p.sumIll) / p.samples) / (p.samples - 1);
    if (variance <= 0 || std::isnan(variance)) return true;
    double ci = 1.96 * std::sqrt(variance) / std::sqrt((double)p.samples);
    return ci <= aTolerance * mean;
  }
  static uint8_t toU8Clamped(double v) { // Uint8ClampedArray store (ToUint8Clamp)
    if (!(v > 0)) return 0; // NaN, -0, negatives
    if (v >= 255) return 255;
    return (uint8_t)std::nearbyint(v); // v is already an integer after Math.floor
  }
  void writeColor(uint8_t* buf, int i, int j, V3 c) const { // camera.ts:455-472
    size_t off = ((size_t)j * imageWidth + i) * 3;
    buf[off + 0] = toU8Clamped(std::floor(255.999 * std::sqrt((double)c.x)));
    buf[off + 1] = toU8Clamped(std::floor(255.999 * std::sqrt((double)c.y)));
    buf[off + 2] = toU8Clamped(std::floor(255.999 * std::sqrt((double)c.z)));
  }

  RenderStats renderRegion(uint8_t* buf, float* linear, double* moments, rt_region reg, Rng& g) const { // :388-431
    int endX = std::min(reg.x + reg.width, imageWidth);
    int endY = std::min(reg.y + reg.height, imageHeight);
    RenderStats rs;
    for (int j = reg.y; j < endY; ++j) {
      for (int i = reg.x; i < endX; ++i) {
        PixelStats px;
        uint32_t pixelIndex = (uint32_t)j * (uint32_t)imageWidth + (uint32_t)i;
        while (px.samples < samples && !pixelConverged(px)) {
          g.begin_path(pixelIndex, (uint32_t)px.samples);
          CNT(paths);
          Ray r = getRay(i, j, g);
          int b = 0;
          V3 c = rayColor(r, mk(1, 1, 1), b, g);
          px.addSample(c, b, useAdaptiveSampling);
        }
        V3 fc = finalColor(px);
        if (buf) writeColor(buf, i, j, fc);
        size_t pi = (size_t)j * imageWidth + i;
        if (linear) { linear[pi * 3 + 0] = fc.x; linear[pi * 3 + 1] = fc.y; linear[pi * 3 + 2] = fc.z; }
        if (moments) {
          double* m = moments + pi * 8;
          m[0] = px.m1[0]; m[1] = px.m1[1]; m[2] = px.m1[2];
          m[3] = px.m2[0]; m[4] = px.m2[1]; m[5] = px.m2[2];
          m[6] = px.samples; m[7] = (double)px.bounces;
        }
        rs.addPixel(px);
      }
    }
    return rs;
  }
};

// ----------------------------------------------------------------------------------------
// Scene factory — src/scenes/scenes.ts:60-199 on the flattened description
// ----------------------------------------------------------------------------------------
static thread_local std::string tl_err;

static Material* buildMaterial(const rt_scene_desc& s, int idx, std::vector<std::unique_ptr<Material>>& pool,
                               rt_status& st, int depthGuard = 0) {
  if (idx < 0 || (uint32_t)idx >= s.n_materials || depthGuard > 64) {
    st = RT_ERR_MATERIAL_NOT_FOUND;
    tl_err = "Material not found: " + std::to_string(idx);
    return nullptr;
  }
  const double* c = s.mat_color + 3 * (size_t)idx;
  double p = s.mat_param[idx];
  Material* m = nullptr;
  switch (s.mat_type[idx]) {
    case RT_MAT_LAMBERT: m = new Lambertian(mk(c[0], c[1], c[2])); break;
    case RT_MAT_METAL: m = new Metal(mk(c[0], c[1], c[2]), p); break;
    case RT_MAT_GLASS: m = new Dielectric(p); break;
    case RT_MAT_LIGHT: m = new DiffuseLight(mk(c[0], c[1], c[2])); break;
    case RT_MAT_MIXED: {
      Material* a = buildMaterial(s, s.mat_child[2 * idx], pool, st, depthGuard + 1);
      if (!a) return nullptr;
      Material* b = buildMaterial(s, s.mat_child[2 * idx + 1], pool, st, depthGuard + 1);
      if (!b) return nullptr;
      m = new MixedMaterial(a, b, p);
      break;
    }
    case RT_MAT_LAYERED: {
      int oi = s.mat_child[2 * idx + 1];
      if (oi < 0 || (uint32_t)oi >= s.n_materials) {
        st = RT_ERR_MATERIAL_NOT_FOUND;
        tl_err = "Material not found: " + std::to_string(oi);
        return nullptr;
      }
      if (s.mat_type[oi] != RT_MAT_GLASS) {
        st = RT_ERR_NOT_DIELECTRIC;
        tl_err = "Material is not a dielectric: " + std::to_string(oi);
        return nullptr;
      }
      Material* inner = buildMaterial(s, s.mat_child[2 * idx], pool, st, depthGuard + 1);
      if (!inner) return nullptr;
      m = new LayeredMaterial(new Dielectric(s.mat_param[oi]), inner);
      break;
    }
    default:
      st = RT_ERR_UNKNOWN_MATERIAL_TYPE;
      tl_err = "Unknown material type: " + std::to_string((int)s.mat_type[idx]);
      return nullptr;
  }
  pool.emplace_back(m);
  return m;
}

static rt_status buildCamera(const rt_scene_desc& s, const rt_render_opts& o, Camera*& out) {
  if (s.n_objects == 0) { tl_err = "scene has no objects"; return RT_ERR_INVALID_ARGUMENT; }
  auto cam = std::make_unique<Camera>();
  std::vector<const Hittable*> list;
  for (uint32_t i = 0; i < s.n_objects; ++i) {
    rt_status st = RT_OK;
    Material* m = buildMaterial(s, s.obj_material[i], cam->materials, st);
    if (!m) return st;
    const double* p = s.obj_pos + 3 * (size_t)i;
    const double* uu = s.obj_u + 3 * (size_t)i;
    const double* vv = s.obj_v + 3 * (size_t)i;
    Hittable* h = nullptr;
    switch (s.obj_type[i]) {
      case RT_OBJ_SPHERE: h = new Sphere(mk(p[0], p[1], p[2]), s.obj_r[i], m, (int)i); break;
      case RT_OBJ_PLANE: h = new Plane(mk(p[0], p[1], p[2]), mk(uu[0], uu[1], uu[2]), mk(vv[0], vv[1], vv[2]), m, (int)i); break;
      case RT_OBJ_QUAD: h = new Quad(mk(p[0], p[1], p[2]), mk(uu[0], uu[1], uu[2]), mk(vv[0], vv[1], vv[2]), m, (int)i); break;
      default:
        tl_err = "Unknown object type: " + std::to_string((int)s.obj_type[i]);
        return RT_ERR_UNKNOWN_OBJECT_TYPE;
    }
    cam->objects.emplace_back(h);
    list.push_back(h);
  }
  cam->bvh = std::make_unique<BVHNode>(list, 0, list.size()); // BVHNode.fromList, scenes.ts:71
  cam->world = cam->bvh.get();
  for (uint32_t i = 0; i < s.n_objects; ++i) // scenes.ts:74-79
    if (s.obj_light[i] && list[i]->hasPdf()) cam->lights.push_back(list[i]);
  cam->init(s.camera, o);
  out = cam.release();
  return RT_OK;
}

// raytracer.ts:185-205
static std::vector<rt_region> divideIntoRegions(int W, int H, int count, int y0) {
  int rh = (int)std::ceil((double)H / count);
  std::vector<rt_region> out;
  for (int i = 0; i < count; ++i) {
    int sy = i * rh;
    int h = std::min(rh, H - sy);
    if (h <= 0) break;
    out.push_back(rt_region{0, y0 + sy, W, h});
  }
  return out;
}

// Flattened view of the reference BVH (for the device-side REFERENCE topology parity check).
static void countNodes(const Hittable* h, int& nodes, int& maxDepth, int d) {
  if (auto* b = dynamic_cast<const BVHNode*>(h)) {
    nodes++;
    maxDepth = std::max(maxDepth, d);
    countNodes(b->left, nodes, maxDepth, d + 1);
    countNodes(b->right, nodes, maxDepth, d + 1);
  }
}

} // namespace orc

// ========================================================================================
// C entry points (ctypes).  Names start with orc_ so they can never be mistaken for the
// product ABI in include/rt_b200.h.
// ========================================================================================
using namespace orc;

extern "C" {

const char* orc_last_error() { return tl_err.c_str(); }

int orc_camera_create(const rt_scene_desc* s, const rt_render_opts* o, void** out) {
  if (!s || !o || !out) { tl_err = "null argument"; return RT_ERR_INVALID_ARGUMENT; }
  Camera* c = nullptr;
  rt_status st = buildCamera(*s, *o, c);
  if (st != RT_OK) return st;
  *out = c;
  return RT_OK;
}
void orc_camera_destroy(void* cam) { delete (Camera*)cam; }

// out: [0]=W [1]=H [2]=n_lights [3]=bvh nodes [4]=bvh depth
void orc_camera_info(void* cam, int* out, float* vecs /* 9x3: center,p00,du,dv,u,v,w,ddu,ddv */, double* focus) {
  Camera* c = (Camera*)cam;
  out[0] = c->imageWidth; out[1] = c->imageHeight; out[2] = (int)c->lights.size();
  int nodes = 0, md = 0;
  countNodes(c->bvh.get(), nodes, md, 1);
  out[3] = nodes; out[4] = md;
  V3 vs[9] = {c->center, c->pixel00Loc, c->pixelDeltaU, c->pixelDeltaV, c->u, c->v, c->w, c->defocusDiskU, c->defocusDiskV};
  for (int i = 0; i < 9; ++i) { vecs[3 * i] = vs[i].x; vecs[3 * i + 1] = vs[i].y; vecs[3 * i + 2] = vs[i].z; }
  *focus = c->focusDistance;
}

// Camera.renderRegion.  threads<=1: the serial path (raytracer.ts:56-59).  threads>1: the
// reference's parallel path — row strips of ceil(H/threads), one per thread, static
// (raytracer.ts:60-90, :185-205), stats merged like RenderStats.merge.
// rng_mode 0 = path-keyed Philox (seed), 1 = sequential xorshift128+ (seed + strip).
// counters: optional uint64[sizeof(Counters)/8] summed event counts.
int orc_render_region(void* cam, const rt_region* reg, uint8_t* rgb8, size_t rgb8_len, float* linear,
                      double* moments, rt_stats* stats, int rng_mode, uint64_t seed, int threads,
                      uint64_t* counters) {
  Camera* c = (Camera*)cam;
  if (!c || !reg) { tl_err = "null argument"; return RT_ERR_INVALID_ARGUMENT; }
  if (rgb8 && rgb8_len < (size_t)c->imageWidth * c->imageHeight * 3) { tl_err = "rgb8 buffer too small"; return RT_ERR_BUFFER_TOO_SMALL; }
  int endY = std::min(reg->y + reg->height, c->imageHeight);
  int rows = std::max(0, endY - reg->y);
  std::vector<rt_region> strips;
  if (threads <= 1 || rows <= 1) strips.push_back(*reg);
  else {
    strips = divideIntoRegions(reg->width, rows, threads, reg->y);
    for (auto& s : strips) s.x = reg->x;
  }
  std::vector<RenderStats> rs(strips.size());
  std::vector<Counters> cn(strips.size());
  std::vector<std::thread> th;
  auto work = [&](size_t k) {
    Rng g;
    g.mode = rng_mode;
    g.seed = seed;
    if (rng_mode == 1) g.seed_sequential(seed * 0x100000001B3ull + k);
    tl_cnt = counters ? &cn[k] : nullptr;
    rs[k] = c->renderRegion(rgb8, linear, moments, strips[k], g);
    tl_cnt = nullptr;
  };
  if (strips.size() == 1) work(0);
  else {
    for (size_t k = 0; k < strips.size(); ++k) th.emplace_back(work, k);
    for (auto& t : th) t.join();
  }
  RenderStats m;
  Counters total;
  for (size_t k = 0; k < strips.size(); ++k) { m.merge(rs[k]); total += cn[k]; }
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->pixels = m.pixels;
    stats->samples_total = m.samplesTotal;
    stats->samples_min = std::isinf(m.samplesMin) ? INT32_MAX : (int32_t)m.samplesMin;
    stats->samples_max = (int32_t)m.samplesMax;
    stats->bounces_total = m.bouncesTotal;
    stats->bounces_min = std::isinf(m.bouncesMin) ? INT32_MAX : (int32_t)m.bouncesMin;
    stats->bounces_max = (int32_t)m.bouncesMax;
    stats->rays = total.rays;
  }
  if (counters) std::memcpy(counters, &total, sizeof(Counters));
  return RT_OK;
}
int orc_counters_len() { return (int)(sizeof(Counters) / 8); }

// Primary visibility through pixel centres (no jitter, no defocus) — parity hook.
int orc_trace_primary(void* cam, const rt_region* reg, int32_t* obj_id, double* t, float* normal, uint8_t* front) {
  Camera* c = (Camera*)cam;
  int endX = std::min(reg->x + reg->width, c->imageWidth);
  int endY = std::min(reg->y + reg->height, c->imageHeight);
  Rng g;
  for (int j = reg->y; j < endY; ++j)
    for (int i = reg->x; i < endX; ++i) {
      Ray r = c->getRay(i, j, g, false);
      HitRecord rec;
      bool h = c->world->hit(r, Interval{0.001, kInf}, rec);
      size_t pi = (size_t)j * c->imageWidth + i;
      if (obj_id) obj_id[pi] = h ? rec.objId : -1;
      if (t) t[pi] = h ? rec.t : kInf;
      if (normal) { normal[3 * pi] = h ? rec.normal.x : 0; normal[3 * pi + 1] = h ? rec.normal.y : 0; normal[3 * pi + 2] = h ? rec.normal.z : 0; }
      if (front) front[pi] = h ? rec.frontFace : 0;
    }
  return RT_OK;
}

// Brute-force closest hit over the object list (no BVH): independent cross-check of the
// BVH restatement (tests/geometry/bvh.test.ts:52-159 does the same against HittableList).
int orc_trace_rays(void* cam, int n, const float* origins, const float* dirs, double tmin, double tmax, int use_bvh,
                   int32_t* obj_id, double* t, float* normal, uint8_t* front) {
  Camera* c = (Camera*)cam;
  HittableList all;
  for (auto& o : c->objects) all.objects.push_back(o.get());
  for (int k = 0; k < n; ++k) {
    Ray r{V3{origins[3 * k], origins[3 * k + 1], origins[3 * k + 2]}, V3{dirs[3 * k], dirs[3 * k + 1], dirs[3 * k + 2]}};
    HitRecord rec;
    bool h = use_bvh ? c->world->hit(r, Interval{tmin, tmax}, rec) : all.hit(r, Interval{tmin, tmax}, rec);
    if (obj_id) obj_id[k] = h ? rec.objId : -1;
    if (t) t[k] = h ? rec.t : kInf;
    if (normal) { normal[3 * k] = h ? rec.normal.x : 0; normal[3 * k + 1] = h ? rec.normal.y : 0; normal[3 * k + 2] = h ? rec.normal.z : 0; }
    if (front) front[k] = h ? rec.frontFace : 0;
  }
  return RT_OK;
}

// rayColor for explicit rays (tests/camera.test.ts drives rayColor directly).
int orc_ray_color(void* cam, const float* origin, const float* dir, uint32_t pixel, uint32_t sample, uint64_t seed,
                  float* rgb, int* bounces) {
  Camera* c = (Camera*)cam;
  Rng g;
  g.seed = seed;
  g.begin_path(pixel, sample);
  int b = 0;
  V3 col = c->rayColor(Ray{V3{origin[0], origin[1], origin[2]}, V3{dir[0], dir[1], dir[2]}}, mk(1, 1, 1), b, g);
  rgb[0] = col.x; rgb[1] = col.y; rgb[2] = col.z;
  *bounces = b;
  return RT_OK;
}

// getRay with the path-keyed stream (tests/camera.test.ts:202-222, :537-587).
int orc_get_ray(void* cam, int i, int j, uint32_t sample, uint64_t seed, float* origin, float* dir) {
  Camera* c = (Camera*)cam;
  Rng g;
  g.seed = seed;
  g.begin_path((uint32_t)j * c->imageWidth + i, sample);
  Ray r = c->getRay(i, j, g);
  origin[0] = r.o.x; origin[1] = r.o.y; origin[2] = r.o.z;
  dir[0] = r.d.x; dir[1] = r.d.y; dir[2] = r.d.z;
  return RT_OK;
}

// ---- unit hooks for the reference's known-answer vectors --------------------------------
void orc_philox4x32(const uint32_t* ctr, const uint32_t* key, int rounds, uint32_t* out) { philox4x32(ctr, key, out, rounds); }
int orc_philox_rounds(void) { return kPhiloxRounds; } // rounds of the path-keyed streams (the CUDA path's RT_PHILOX_ROUNDS)

// uniforms of the path-keyed stream (so tests can predict draws)
void orc_stream_uniforms(uint32_t pixel, uint32_t sample, uint32_t stream, uint64_t seed, int n, double* out) {
  Rng g;
  g.seed = seed;
  g.begin_path(pixel, sample);
  g.begin_stream(stream);
  for (int i = 0; i < n; ++i) out[i] = g.next();
}

static Material* unitMat() { static Lambertian m(mk(0.5, 0.5, 0.5)); return &m; }

// returns 1 on hit; out = [t, px,py,pz, nx,ny,nz, frontFace]
int orc_sphere_hit(const double* c, double r, const double* o, const double* d, double tmin, double tmax, double* out) {
  Sphere s(mk(c[0], c[1], c[2]), r, unitMat(), 0);
  HitRecord rec;
  if (!s.hit(Ray{mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2])}, Interval{tmin, tmax}, rec)) return 0;
  double v[8] = {rec.t, rec.p.x, rec.p.y, rec.p.z, rec.normal.x, rec.normal.y, rec.normal.z, (double)rec.frontFace};
  std::memcpy(out, v, sizeof(v));
  return 1;
}
double orc_sphere_pdf_value(const double* c, double r, const double* o, const double* d) {
  Sphere s(mk(c[0], c[1], c[2]), r, unitMat(), 0);
  return s.pdfValue(mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2]));
}
void orc_sphere_pdf_random(const double* c, double r, const double* o, uint64_t seed, int n, float* out) {
  Sphere s(mk(c[0], c[1], c[2]), r, unitMat(), 0);
  Rng g; g.mode = 1; g.seed_sequential(seed);
  for (int i = 0; i < n; ++i) { V3 v = s.pdfRandomVec(mk(o[0], o[1], o[2]), g); out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z; }
}
// kind: RT_OBJ_PLANE or RT_OBJ_QUAD; out = [t, px,py,pz, nx,ny,nz, frontFace]
int orc_planar_hit(int kind, const double* q, const double* u, const double* v, const double* o, const double* d, double tmin,
                   double tmax, double* out) {
  HitRecord rec;
  Ray r{mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2])};
  bool h;
  if (kind == RT_OBJ_PLANE) { Plane p(mk(q[0], q[1], q[2]), mk(u[0], u[1], u[2]), mk(v[0], v[1], v[2]), unitMat(), 0); h = p.hit(r, Interval{tmin, tmax}, rec); }
  else { Quad p(mk(q[0], q[1], q[2]), mk(u[0], u[1], u[2]), mk(v[0], v[1], v[2]), unitMat(), 0); h = p.hit(r, Interval{tmin, tmax}, rec); }
  if (!h) return 0;
  double vals[8] = {rec.t, rec.p.x, rec.p.y, rec.p.z, rec.normal.x, rec.normal.y, rec.normal.z, (double)rec.frontFace};
  std::memcpy(out, vals, sizeof(vals));
  return 1;
}
// Plane ctor products + intersect: out = [nx,ny,nz, d, wx,wy,wz, hit, t, alpha, beta]
void orc_plane_intersect(const double* q, const double* u, const double* v, const double* o, const double* d, double tmin,
                         double tmax, double* out) {
  Plane p(mk(q[0], q[1], q[2]), mk(u[0], u[1], u[2]), mk(v[0], v[1], v[2]), unitMat(), 0);
  double t = 0, a = 0, b = 0;
  bool h = p.intersect(Ray{mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2])}, Interval{tmin, tmax}, t, a, b);
  double vals[11] = {p.normal.x, p.normal.y, p.normal.z, p.d, p.w.x, p.w.y, p.w.z, (double)h, t, a, b};
  std::memcpy(out, vals, sizeof(vals));
}
// bounding box of one object: out = [minx,miny,minz,maxx,maxy,maxz]
void orc_object_bbox(int kind, const double* q, const double* u, const double* v, double r, double* out) {
  AABB b;
  if (kind == RT_OBJ_SPHERE) b = Sphere(mk(q[0], q[1], q[2]), r, unitMat(), 0).boundingBox();
  else if (kind == RT_OBJ_PLANE) b = Plane(mk(q[0], q[1], q[2]), mk(u[0], u[1], u[2]), mk(v[0], v[1], v[2]), unitMat(), 0).boundingBox();
  else b = Quad(mk(q[0], q[1], q[2]), mk(u[0], u[1], u[2]), mk(v[0], v[1], v[2]), unitMat(), 0).boundingBox();
  double vals[6] = {b.mn.x, b.mn.y, b.mn.z, b.mx.x, b.mx.y, b.mx.z};
  std::memcpy(out, vals, sizeof(vals));
}
double orc_quad_pdf_value(const double* q, const double* u, const double* v, const double* o, const double* d) {
  Quad p(mk(q[0], q[1], q[2]), mk(u[0], u[1], u[2]), mk(v[0], v[1], v[2]), unitMat(), 0);
  return p.pdfValue(mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2]));
}
void orc_quad_pdf_random(const double* q, const double* u, const double* v, const double* o, uint64_t seed, int n, float* out) {
  Quad p(mk(q[0], q[1], q[2]), mk(u[0], u[1], u[2]), mk(v[0], v[1], v[2]), unitMat(), 0);
  Rng g; g.mode = 1; g.seed_sequential(seed);
  for (int i = 0; i < n; ++i) { V3 x = p.pdfRandomVec(mk(o[0], o[1], o[2]), g); out[3 * i] = x.x; out[3 * i + 1] = x.y; out[3 * i + 2] = x.z; }
}
int orc_aabb_hit(const double* mn, const double* mx, const double* o, const double* d, double tmin, double tmax) {
  return aabbHit(AABB{mk(mn[0], mn[1], mn[2]), mk(mx[0], mx[1], mx[2])}, Ray{mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2])}, Interval{tmin, tmax});
}
void orc_surrounding_box(const double* mn0, const double* mx0, const double* mn1, const double* mx1, double* out) {
  AABB b = surroundingBox(AABB{mk(mn0[0], mn0[1], mn0[2]), mk(mx0[0], mx0[1], mx0[2])}, AABB{mk(mn1[0], mn1[1], mn1[2]), mk(mx1[0], mx1[1], mx1[2])});
  double vals[6] = {b.mn.x, b.mn.y, b.mn.z, b.mx.x, b.mx.y, b.mx.z};
  std::memcpy(out, vals, sizeof(vals));
}
double orc_cosine_pdf_value(const double* n, const double* d) { return CosinePDF(mk(n[0], n[1], n[2])).value(mk(d[0], d[1], d[2])); }
void orc_cosine_pdf_generate(const double* n, uint64_t seed, int cnt, float* out) {
  CosinePDF p(mk(n[0], n[1], n[2]));
  Rng g; g.mode = 1; g.seed_sequential(seed);
  for (int i = 0; i < cnt; ++i) { V3 x = p.generate(g); out[3 * i] = x.x; out[3 * i + 1] = x.y; out[3 * i + 2] = x.z; }
}
void orc_random_cosine_direction(uint64_t seed, int cnt, float* out) {
  Rng g; g.mode = 1; g.seed_sequential(seed);
  for (int i = 0; i < cnt; ++i) { V3 x = randomCosineDirection(g); out[3 * i] = x.x; out[3 * i + 1] = x.y; out[3 * i + 2] = x.z; }
}
void orc_onb(const double* n, float* out) {
  ONB b(mk(n[0], n[1], n[2]));
  float v[9] = {b.u.x, b.u.y, b.u.z, b.v.x, b.v.y, b.v.z, b.w.x, b.w.y, b.w.z};
  std::memcpy(out, v, sizeof(v));
}
// MixturePDF.value with arbitrary component values/weights (pdf.ts:77-83)
double orc_mixture_value(int n, const double* values, const double* weights) {
  double total = 0;
  for (int i = 0; i < n; ++i) total += weights[i];
  double sum = 0;
  for (int i = 0; i < n; ++i) sum += weights[i] * values[i];
  return sum / total;
}
// MixturePDF.generate selection (pdf.ts:85-98): index chosen for uniform `u`
int orc_mixture_select(int n, const double* weights, double u) {
  double total = 0;
  for (int i = 0; i < n; ++i) total += weights[i];
  double rnd = u * total, partial = 0;
  for (int i = 0; i < n; ++i) { partial += weights[i]; if (rnd < partial) return i; }
  return n - 1;
}
void orc_reflect(const double* v, const double* n, float* out) { V3 r = reflect(mk(v[0], v[1], v[2]), mk(n[0], n[1], n[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void orc_refract(const double* v, const double* n, double eta, float* out) { V3 r = refract(mk(v[0], v[1], v[2]), mk(n[0], n[1], n[2]), eta); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void orc_unit(const double* v, float* out) { V3 r = unit(mk(v[0], v[1], v[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
double orc_dielectric_reflectance(double cosine, double ratio) { return Dielectric::reflectance(cosine, ratio); }
double orc_metal_fuzz_clamp(double f) { return Metal(mk(1, 1, 1), f).fuzz; }
double orc_mixed_weight_clamp(double w) { static Lambertian l(mk(0, 0, 0)); return MixedMaterial(&l, &l, w).weight; }

// Scatter one material root of a scene description at a synthetic hit.
// in: rIn origin/dir, hit p/normal/frontFace.  out = [scattered?, pdf?, att r,g,b, dir x,y,z, reflected]
int orc_material_scatter(const rt_scene_desc* s, int root, const double* ro, const double* rd, const double* p, const double* n,
                         int frontFace, uint64_t seed, uint32_t sample, double* out, float* emitted) {
  std::vector<std::unique_ptr<Material>> pool;
  rt_status st = RT_OK;
  Material* m = buildMaterial(*s, root, pool, st);
  if (!m) return -(int)st;
  HitRecord rec{mk(p[0], p[1], p[2]), mk(n[0], n[1], n[2]), 1.0, frontFace != 0, m, 0};
  Rng g; g.mode = 1; g.seed_sequential(seed * 1315423911ull + sample);
  ScatterResult sr;
  bool ok = m->scatter(Ray{mk(ro[0], ro[1], ro[2]), mk(rd[0], rd[1], rd[2])}, rec, g, sr);
  V3 e = m->emitted(rec);
  if (emitted) { emitted[0] = e.x; emitted[1] = e.y; emitted[2] = e.z; }
  if (!ok) return 0;
  double vals[9] = {(double)sr.hasScattered, (double)sr.hasPdf, sr.attenuation.x, sr.attenuation.y, sr.attenuation.z,
                    sr.hasScattered ? sr.scattered.d.x : 0.0, sr.hasScattered ? sr.scattered.d.y : 0.0,
                    sr.hasScattered ? sr.scattered.d.z : 0.0, (double)sr.reflected};
  std::memcpy(out, vals, sizeof(vals));
  return 1;
}

// ---- per-function hooks with EXPLICIT uniforms: the caller supplies the sequence Math.random() would return,
// so the CUDA device functions (rt_debug_* in include/rt_b200.h) can be driven with the very same numbers ----
static Rng listRng(const double* u, int n) { Rng g; g.mode = 2; g.list = u; g.list_n = (uint64_t)(n > 0 ? n : 0); return g; }

// material.scatter + material.emitted at a synthetic hit (materials/*.ts).  out = [scattered?, pdf?, att r,g,b,
// dir x,y,z, reflected]; returns 1 = scattered / pdf result, 0 = null (absorbed / light), < 0 = bad scene.
int orc_material_scatter_u(const rt_scene_desc* s, int root, const double* ro, const double* rd, const double* p, const double* n,
                           int frontFace, const double* uniforms, int n_uniforms, double* out, float* emitted, int* used) {
  std::vector<std::unique_ptr<Material>> pool;
  rt_status st = RT_OK;
  Material* m = buildMaterial(*s, root, pool, st);
  if (!m) return -(int)st;
  HitRecord rec{mk(p[0], p[1], p[2]), mk(n[0], n[1], n[2]), 1.0, frontFace != 0, m, 0};
  Rng g = listRng(uniforms, n_uniforms);
  ScatterResult sr;
  bool ok = m->scatter(Ray{mk(ro[0], ro[1], ro[2]), mk(rd[0], rd[1], rd[2])}, rec, g, sr);
  V3 e = m->emitted(rec);
  if (emitted) { emitted[0] = e.x; emitted[1] = e.y; emitted[2] = e.z; }
  if (used) *used = (int)g.draws;
  if (!ok) return 0;
  double vals[9] = {(double)sr.hasScattered, (double)sr.hasPdf, sr.attenuation.x, sr.attenuation.y, sr.attenuation.z,
                    sr.hasScattered ? sr.scattered.d.x : 0.0, sr.hasScattered ? sr.scattered.d.y : 0.0,
                    sr.hasScattered ? sr.scattered.d.z : 0.0, (double)sr.reflected};
  std::memcpy(out, vals, sizeof(vals));
  return 1;
}
// Camera.getRay(i, j) (camera.ts:176-210)
int orc_get_ray_u(void* cam, int i, int j, const double* uniforms, int n_uniforms, float* origin, float* dir, int* used) {
  Camera* c = (Camera*)cam;
  Rng g = listRng(uniforms, n_uniforms);
  Ray r = c->getRay(i, j, g);
  origin[0] = r.o.x; origin[1] = r.o.y; origin[2] = r.o.z;
  dir[0] = r.d.x; dir[1] = r.d.y; dir[2] = r.d.z;
  if (used) *used = (int)g.draws;
  return RT_OK;
}
// lights[k].pdfValue(origin, direction) of the camera's light list (quad.ts:123-140, sphere.ts:106-131)
double orc_light_pdf_value(void* cam, int k, const double* o, const double* d) {
  Camera* c = (Camera*)cam;
  if (k < 0 || (size_t)k >= c->lights.size()) return -1.0;
  return c->lights[(size_t)k]->pdfValue(mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2]));
}
// lights[k].pdfRandomVec(origin) (quad.ts:148-158, sphere.ts:140-147)
int orc_light_random_vec_u(void* cam, int k, const double* o, const double* uniforms, int n_uniforms, float* out, int* used) {
  Camera* c = (Camera*)cam;
  if (k < 0 || (size_t)k >= c->lights.size()) return RT_ERR_INVALID_ARGUMENT;
  Rng g = listRng(uniforms, n_uniforms);
  V3 v = c->lights[(size_t)k]->pdfRandomVec(mk(o[0], o[1], o[2]), g);
  out[0] = v.x; out[1] = v.y; out[2] = v.z;
  if (used) *used = (int)g.draws;
  return RT_OK;
}
// The diffuse branch of rayColor (camera.ts:285-308) at a hit point with normal n: out = [dir x,y,z, pdfValue,
// scatterPdfValue, continues (pdfValue > 0.0001)].
int orc_diffuse_bounce_u(void* cam, const double* p, const double* n, const double* uniforms, int n_uniforms, double* out, int* used) {
  Camera* c = (Camera*)cam;
  Rng g = listRng(uniforms, n_uniforms);
  V3 dir;
  double pdfValue, spv;
  c->diffuseBounce(mk(p[0], p[1], p[2]), mk(n[0], n[1], n[2]), g, dir, pdfValue, spv);
  out[0] = dir.x; out[1] = dir.y; out[2] = dir.z; out[3] = pdfValue; out[4] = spv; out[5] = pdfValue <= 0.0001 ? 0.0 : 1.0;
  if (used) *used = (int)g.draws;
  return RT_OK;
}

// finalColor + writeColorToBuffer on one colour (camera.ts:455-472)
void orc_write_color(const double* c, uint8_t* out) {
  out[0] = Camera::toU8Clamped(std::floor(255.999 * std::sqrt((double)f32(c[0]))));
  out[1] = Camera::toU8Clamped(std::floor(255.999 * std::sqrt((double)f32(c[1]))));
  out[2] = Camera::toU8Clamped(std::floor(255.999 * std::sqrt((double)f32(c[2]))));
}
// pixelConverged on explicit statistics (camera.ts:348-368)
int orc_pixel_converged(void* cam, int samples, double sumIll, double sumIll2) {
  PixelStats p;
  p.samples = samples; p.sumIll = sumIll; p.sumIll2 = sumIll2;
  return ((Camera*)cam)->pixelConverged(p);
}

// The vector / ray / interval layer on its own (vec3.ts, ray.ts:25-28, interval.ts), so the reference's
// tests/geometry/{vec3,ray,interval}.test.ts vectors can be replayed against it.
// op: 0 negate, 1 add, 2 subtract, 3 multiply(s), 4 multiplyVec, 5 divide(s), 6 cross, 7 unitVector -> out3;
//     8 lengthSquared, 9 length, 10 dot, 11 nearZero (vec3.ts:215-221), 12 illuminance -> return value.
double orc_vec3_op(int op, const double* a, const double* b, double s, float* out3) {
  const V3 A = mk(a[0], a[1], a[2]);
  const V3 B = b ? mk(b[0], b[1], b[2]) : mk(0, 0, 0);
  V3 r = mk(0, 0, 0);
  double v = 0;
  switch (op) {
    case 0: r = neg(A); break;
    case 1: r = add(A, B); break;
    case 2: r = sub(A, B); break;
    case 3: r = scale(A, s); break;
    case 4: r = mulv(A, B); break;
    case 5: r = divs(A, s); break;
    case 6: r = cross(A, B); break;
    case 7: r = unit(A); break;
    case 8: v = len2(A); break;
    case 9: v = len(A); break;
    case 10: v = dot(A, B); break;
    case 11: v = (std::fabs((double)A.x) < 1e-8 && std::fabs((double)A.y) < 1e-8 && std::fabs((double)A.z) < 1e-8) ? 1 : 0; break;
    case 12: v = illuminance(A); break;
    default: return std::nan("");
  }
  if (out3) { out3[0] = r.x; out3[1] = r.y; out3[2] = r.z; }
  return v;
}
void orc_ray_at(const double* o, const double* d, double t, float* out3) {
  Ray r{mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2])};
  V3 p = r.at(t);
  out3[0] = p.x; out3[1] = p.y; out3[2] = p.z;
}
// op: 0 size, 1 contains, 2 surrounds, 3 clamp
double orc_interval_op(int op, double mn, double mx, double x) {
  Interval iv{mn, mx};
  switch (op) {
    case 0: return iv.size();
    case 1: return iv.contains(x) ? 1 : 0;
    case 2: return iv.surrounds(x) ? 1 : 0;
    case 3: return iv.clamp(x);
    default: return std::nan("");
  }
}
// kind: 0 Vec3.random(mn, mx) (vec3.ts:272-278), 1 randomInUnitSphere (:285-292), 2 randomInUnitDisk (:357-364)
void orc_sample_vec3(int kind, uint64_t seed, double mn, double mx, int cnt, float* out) {
  Rng g; g.mode = 1; g.seed_sequential(seed);
  for (int i = 0; i < cnt; ++i) {
    V3 x = kind == 0 ? randomVec(g, mn, mx) : (kind == 1 ? randomInUnitSphere(g) : randomInUnitDisk(g));
    out[3 * i] = x.x; out[3 * i + 1] = x.y; out[3 * i + 2] = x.z;
  }
}

} // extern "C"
