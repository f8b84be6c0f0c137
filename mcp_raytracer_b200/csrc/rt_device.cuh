// rt_device.cuh — device-side building blocks of the sm_100a path tracer.
//
// Everything the reference evaluates per ray (src/camera.ts:221-319 and callees) as FP32
// device functions, plus guarded FP64 re-tests: a primitive test whose FP32 outcome is
// within its own error bound of a decision threshold (hit/miss, interval end, quad edge) or
// whose t is not good to ~2e-5 relative is re-evaluated with the reference's exact formula
// in FP64 from the reference's FP32-stored operands (ExactPrim).  That keeps primary-hit
// object ids identical to the reference's FP64-scalar arithmetic without paying FP64 on the
// common path.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "rt_types.h"

namespace rt {

#define RT_DEV __device__ __forceinline__

static constexpr float kEps32 = 5.9604645e-08f; // 2^-24
static constexpr float kRayTMin = 0.001f;       // camera.ts:249

// ---------------------------------------------------------------------------------------
// small vector helpers
// ---------------------------------------------------------------------------------------
struct V3 {
  float x, y, z;
};
RT_DEV V3 mk3(float x, float y, float z) { return V3{x, y, z}; }
RT_DEV V3 ld3(const float* p) { return V3{p[0], p[1], p[2]}; }
RT_DEV V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
RT_DEV V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
RT_DEV V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
RT_DEV V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
RT_DEV V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
RT_DEV float dot3(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
RT_DEV V3 cross3(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
RT_DEV V3 fma3(float s, V3 a, V3 b) { return V3{fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)}; } // s*a+b
RT_DEV float l1norm(V3 a) { return fabsf(a.x) + fabsf(a.y) + fabsf(a.z); }
RT_DEV V3 normalize3(V3 a) { // gl-matrix normalize: zero stays zero
  float l = dot3(a, a);
  float s = l > 0.f ? rsqrtf(l) : 0.f;
  // one Newton step brings rsqrtf (2 ulp) to ~0.5 ulp so unit vectors are unit to FP32 accuracy
  s = s * fmaf(-0.5f * l * s, s, 1.5f);
  return a * s;
}
RT_DEV float maxc(V3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }

RT_DEV F4 ldg4(const F4* p) {
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return F4{v.x, v.y, v.z, v.w};
}
RT_DEV I4 ldgi4(const I4* p) {
  int4 v = __ldg(reinterpret_cast<const int4*>(p));
  return I4{v.x, v.y, v.z, v.w};
}
RT_DEV I2 ldgi2(const I2* p) {
  int2 v = __ldg(reinterpret_cast<const int2*>(p));
  return I2{v.x, v.y};
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10, keyed (pixel, sample), counter (block, stream, seed_lo, seed_hi).
// stream 0 = camera ray, stream 1+b = the bounce entered with `bounces == b`.
// Draw i of a stream is word i&3 of block i>>2, mapped to (w>>8) * 2^-24.
// ---------------------------------------------------------------------------------------
struct Rng {
  uint32_t k0, k1, seed_lo, seed_hi, stream, block;
  uint32_t w0, w1, w2, w3;
  int avail;
  RT_DEV void begin_path(uint32_t pixel, uint32_t sample, uint32_t slo, uint32_t shi) {
    k0 = pixel; k1 = sample; seed_lo = slo; seed_hi = shi;
    begin_stream(0);
  }
  RT_DEV void begin_stream(uint32_t s) { stream = s; block = 0; avail = 0; }
  RT_DEV void refill() {
    uint32_t c0 = block, c1 = stream, c2 = seed_lo, c3 = seed_hi, a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
      uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
      c0 = h1 ^ c1 ^ a; c1 = l1; c2 = h0 ^ c3 ^ b; c3 = l0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    w0 = c0; w1 = c1; w2 = c2; w3 = c3;
    ++block;
    avail = 4;
  }
  RT_DEV float next() {
    if (avail == 0) refill();
    uint32_t w = w0;
    w0 = w1; w1 = w2; w2 = w3;
    --avail;
    return (float)(w >> 8) * 5.9604645e-08f;
  }
};

// ---------------------------------------------------------------------------------------
// Ray with the per-ray invariants the tests share
// ---------------------------------------------------------------------------------------
struct Ray {
  V3 o, d;
};
struct RayPre {
  float a, inva, l1d; // |d|^2, 1/|d|^2, L1 norm of d
  V3 idir;            // 1/d
  V3 oid;             // o * idir
};
RT_DEV RayPre precompute(const Ray& r, bool need_box) {
  RayPre p;
  p.a = dot3(r.d, r.d);
  p.inva = 1.0f / p.a;
  p.l1d = l1norm(r.d);
  if (need_box) {
    p.idir = V3{1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
    p.oid = r.o * p.idir;
  } else {
    p.idir = V3{0, 0, 0};
    p.oid = V3{0, 0, 0};
  }
  return p;
}

// ---------------------------------------------------------------------------------------
// Exact (FP64) re-tests: the reference's formulas on the reference's operands, no FMA
// contraction (__d*_rn), so results agree with V8 double arithmetic.
// ---------------------------------------------------------------------------------------
RT_DEV double ddot(const float* a, V3 b) {
  return __dadd_rn(__dadd_rn(__dmul_rn((double)a[0], (double)b.x), __dmul_rn((double)a[1], (double)b.y)),
                   __dmul_rn((double)a[2], (double)b.z));
}
RT_DEV double ddot(V3 a, V3 b) {
  return __dadd_rn(__dadd_rn(__dmul_rn((double)a.x, (double)b.x), __dmul_rn((double)a.y, (double)b.y)),
                   __dmul_rn((double)a.z, (double)b.z));
}
// One primitive, the reference's formula in FP64 over the open interval (tmin, tmax):
// sphere.ts:45-66, plane.ts:55-77 + quad.ts:60 (quad: inclusive inside test; plane: none).
__device__ __noinline__ bool prim_exact(const ExactPrim* e, Ray r, double tmin, double tmax, double& t_out) {
  if (e->type == OBJ_SPHERE) {
    V3 c = ld3(e->q);
    V3 oc = V3{__fsub_rn(r.o.x, c.x), __fsub_rn(r.o.y, c.y), __fsub_rn(r.o.z, c.z)};
    double a = ddot(r.d, r.d);
    double hb = ddot(oc, r.d);
    double rr = e->r;
    double cc = __dsub_rn(ddot(oc, oc), __dmul_rn(rr, rr));
    double disc = __dsub_rn(__dmul_rn(hb, hb), __dmul_rn(a, cc));
    if (disc < 0) return false;
    double sq = sqrt(disc);
    double root = (-hb - sq) / a;
    if (!(tmin < root && root < tmax)) {
      root = (-hb + sq) / a;
      if (!(tmin < root && root < tmax)) return false;
    }
    t_out = root;
    return true;
  }
  double denom = ddot(e->n, r.d);
  if (fabs(denom) < 1e-8) return false;
  double t = (e->D - ddot(e->n, r.o)) / denom;
  if (!(tmin < t && t < tmax)) return false;
  if (e->type == OBJ_QUAD) {
    // r.at(t): scale then add, each rounded to FP32 (ray.ts:25-28)
    V3 sd = V3{(float)__dmul_rn((double)r.d.x, t), (float)__dmul_rn((double)r.d.y, t), (float)__dmul_rn((double)r.d.z, t)};
    V3 ip = V3{__fadd_rn(r.o.x, sd.x), __fadd_rn(r.o.y, sd.y), __fadd_rn(r.o.z, sd.z)};
    V3 hp = V3{__fsub_rn(ip.x, e->q[0]), __fsub_rn(ip.y, e->q[1]), __fsub_rn(ip.z, e->q[2])};
    auto crossf = [](V3 a, V3 b) {
      double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
      return V3{(float)__dsub_rn(__dmul_rn(ay, bz), __dmul_rn(az, by)), (float)__dsub_rn(__dmul_rn(az, bx), __dmul_rn(ax, bz)),
                (float)__dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx))};
    };
    double alpha = ddot(e->w, crossf(hp, ld3(e->v)));
    double beta = ddot(e->w, crossf(ld3(e->u), hp));
    if (alpha < 0 || alpha > 1 || beta < 0 || beta > 1) return false;
  }
  t_out = t;
  return true;
}

// Decide in FP64 whether slot `slot` beats the current best (tbest from slot sbest).  When
// the two distances agree to FP32 resolution the incumbent is re-evaluated in FP64 too and
// the comparison is the reference's strict `<` (interval.ts:51-53): on an exact tie the
// primitive tested first keeps the hit, as in hittableList.ts:76-84 / bvh.ts:139-145.
__device__ __noinline__ bool exact_closer(const ExactPrim* ex, int slot, int sbest, float tbest, Ray r, float& t_out) {
  double tn;
  if (!prim_exact(ex + slot, r, 0.001, (double)CUDART_INF_F, tn)) return false;
  if (sbest >= 0) {
    double tb = (double)tbest;
    if (tn > tb * (1.0 + 4.0 * (double)kEps32)) return false;
    if (tn >= tb * (1.0 - 4.0 * (double)kEps32)) {
      double te;
      if (prim_exact(ex + sbest, r, 0.001, (double)CUDART_INF_F, te)) tb = te;
      if (!(tn < tb)) return false;
    }
  }
  t_out = (float)tn;
  return true;
}

// ---------------------------------------------------------------------------------------
// FP32 primitive tests.  Return 1 = hit (t_out valid), 0 = miss, 2 = ambiguous (re-test).
// ---------------------------------------------------------------------------------------
// sphere.ts:45-66 rewritten for FP32: disc/a = r^2 - |oc - (oc.d/a) d|^2 (cancellation-free
// form), roots -s -/+ sqrt(disc/a^2).
RT_DEV int sphere_fast(F4 s, const Ray& r, const RayPre& pre, float tmin, float tmax, float& t_out) {
  V3 oc = r.o - mk3(s.x, s.y, s.z);
  float hb = dot3(oc, r.d);
  float sp = hb * pre.inva;
  V3 l = fma3(-sp, r.d, oc);
  float l2 = dot3(l, l);
  float r2 = s.w * s.w;
  float da = r2 - l2;
  // Spheres subtending less than 1e-4 of their distance are below the resolution of the
  // reference's own FP32 positions; for the rest the error bound below is < 0.01 r^2.
  if (da < -0.01f * r2) return 0;
  float el = 4.f * kEps32 * fmaf(fabsf(sp), pre.l1d, l1norm(oc));
  float E = fmaf(2.f * l1norm(l), el, fmaf(3.f * el, el, 8.f * kEps32 * (r2 + l2)));
  if (da <= E) return da < -E ? 0 : 2;
  float sqd = sqrtf(da * pre.inva);
  float terr = fmaf(E * pre.inva, 0.5f / sqd, 4.f * kEps32 * (fabsf(sp) + sqd));
  float t = -sp - sqd;
  // near root first, then far root, each against the open interval (sphere.ts:59-65)
  if (!(t > tmin + terr && t < tmax - terr)) {
    if (fabsf(t - tmin) <= terr || fabsf(t - tmax) <= terr) return 2;
    t = -sp + sqd;
    if (!(t > tmin + terr && t < tmax - terr)) {
      if (fabsf(t - tmin) <= terr || fabsf(t - tmax) <= terr) return 2;
      return 0;
    }
  }
  if (terr > 2e-5f * t) return 2;
  t_out = t;
  return 1;
}

// plane.ts:55-77 / quad.ts:50-63.  p0 = (n, D), p1 = (A, q.A), p2 = (B, q.B).
RT_DEV int planar_fast(F4 p0, F4 p1, F4 p2, bool is_quad, const Ray& r, float tmin, float tmax, float& t_out) {
  V3 n = mk3(p0.x, p0.y, p0.z);
  float denom = dot3(n, r.d);
  if (fabsf(denom) < 1e-8f) return 0;
  float no = dot3(n, r.o);
  float inv = 1.0f / denom;
  float t = (p0.w - no) * inv;
  float terr = 4.f * kEps32 * (fabsf(p0.w) + fabsf(n.x * r.o.x) + fabsf(n.y * r.o.y) + fabsf(n.z * r.o.z)) * fabsf(inv) +
               4.f * kEps32 * fabsf(t);
  if (!(t > tmin + terr && t < tmax - terr)) {
    if (fabsf(t - tmin) <= terr || fabsf(t - tmax) <= terr) return 2;
    return 0;
  }
  if (is_quad) {
    V3 p = fma3(t, r.d, r.o);
    V3 A = mk3(p1.x, p1.y, p1.z), B = mk3(p2.x, p2.y, p2.z);
    float alpha = dot3(p, A) - p1.w;
    float beta = dot3(p, B) - p2.w;
    float ea = 4.f * kEps32 * (fabsf(p.x * A.x) + fabsf(p.y * A.y) + fabsf(p.z * A.z) + fabsf(p1.w)) + terr * fabsf(dot3(r.d, A));
    float eb = 4.f * kEps32 * (fabsf(p.x * B.x) + fabsf(p.y * B.y) + fabsf(p.z * B.z) + fabsf(p2.w)) + terr * fabsf(dot3(r.d, B));
    bool in = alpha >= ea && alpha <= 1.f - ea && beta >= eb && beta <= 1.f - eb;
    if (!in) {
      bool out = alpha < -ea || alpha > 1.f + ea || beta < -eb || beta > 1.f + eb;
      return out ? 0 : 2;
    }
  }
  t_out = t;
  return 1;
}

// One primitive slot against the ray; updates (tbest, sbest) on a closer hit.  The ray
// interval is always (0.001, tbest) — camera.ts:249 with the shrinking upper end of
// hittableList.ts:76-84 / bvh.ts:139-141.
template <class Scene>
RT_DEV void test_slot(const Scene& S, int slot, int type, F4 p0, F4 p1, F4 p2, const Ray& r, const RayPre& pre,
                      float& tbest, int& sbest) {
  float t;
  int res;
  if (type == OBJ_SPHERE) res = sphere_fast(p0, r, pre, kRayTMin, tbest, t);
  else res = planar_fast(p0, p1, p2, type == OBJ_QUAD, r, kRayTMin, tbest, t);
  if (res == 2) res = exact_closer(S.exact, slot, sbest, tbest, r, t) ? 1 : 0;
  if (res == 1) { tbest = t; sbest = slot; }
}

// ---------------------------------------------------------------------------------------
// Box tests
// ---------------------------------------------------------------------------------------
// SAH trees: true slab test, conservative (tmax padded by 4 ulp); NaNs from 0*inf are
// dropped by fminf/fmaxf.  Returns entry distance through tnear.
RT_DEV bool slab_hit(const float* mn, const float* mx, const RayPre& pre, float tmin, float tmax, float& tnear) {
  float x0 = fmaf(mn[0], pre.idir.x, -pre.oid.x), x1 = fmaf(mx[0], pre.idir.x, -pre.oid.x);
  float y0 = fmaf(mn[1], pre.idir.y, -pre.oid.y), y1 = fmaf(mx[1], pre.idir.y, -pre.oid.y);
  float z0 = fmaf(mn[2], pre.idir.z, -pre.oid.z), z1 = fmaf(mx[2], pre.idir.z, -pre.oid.z);
  float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
  float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
  tnear = tn;
  return tn <= tf * 1.0000005f + 1e-30f;
}
// REFERENCE trees: aabb.ts:30-59 verbatim in structure — each axis independently against the
// original interval, NaN comparisons fall through as "overlap" — padded so FP32 never
// rejects what FP64 accepts.
RT_DEV bool ref_box_hit(const float* mn, const float* mx, const Ray& r, const RayPre& pre, float tmin, float tmax) {
  const float o[3] = {r.o.x, r.o.y, r.o.z};
  const float id[3] = {pre.idir.x, pre.idir.y, pre.idir.z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float t0 = (mn[a] - o[a]) * id[a];
    float t1 = (mx[a] - o[a]) * id[a];
    if (id[a] < 0.f) { float tt = t0; t0 = t1; t1 = tt; }
    float lo = t0 > tmin ? t0 : tmin;
    float hi = t1 < tmax ? t1 : tmax;
    float pad = 8.f * kEps32 * (fabsf(lo) + fabsf(hi));
    if (hi + pad <= lo) return false;
  }
  return true;
}

// ---------------------------------------------------------------------------------------
// Closest hit over (tmin, inf): the device form of world.hit (camera.ts:249).
// ---------------------------------------------------------------------------------------
struct Node4 {
  F4 a, b, c, d;
};
RT_DEV Node4 load_node(const F4* nodes, int idx) {
  const F4* p = nodes + 4 * (size_t)idx;
  return Node4{ldg4(p), ldg4(p + 1), ldg4(p + 2), ldg4(p + 3)};
}

template <class Scene>
RT_DEV void intersect_leaf(const Scene& S, int ref, const Ray& r, const RayPre& pre, float tmin, float& tbest, int& sbest) {
  int v = ~ref;
  int first = v >> 6, count = ((v >> 4) & 3) + 1, mask = v & 15;
  for (int k = 0; k < count; ++k) {
    int slot = first + k;
    F4 p0 = ldg4(S.p0 + slot);
    if ((mask >> k) & 1) {
      int type = (ldgi2(S.slot_info + slot).y >> 30) & 3;
      test_slot(S, slot, type, p0, ldg4(S.p1 + slot), ldg4(S.p2 + slot), r, pre, tbest, sbest);
    } else {
      F4 z{0, 0, 0, 0};
      test_slot(S, slot, OBJ_SPHERE, p0, z, z, r, pre, tbest, sbest);
    }
  }
}

// LIST: every slot in object order; records come from shared memory.
struct SmemList {
  const F4* p0;
  const F4* p1;
  const F4* p2;
  const int* type;
  int n;
};
template <class Scene>
RT_DEV void trace_list(const Scene& S, const SmemList& L, const Ray& r, const RayPre& pre, float tmin, float& tbest, int& sbest) {
  for (int s = 0; s < L.n; ++s) test_slot(S, s, L.type[s], L.p0[s], L.p1[s], L.p2[s], r, pre, tbest, sbest);
}

// SAH: unbounded prefix, then stack traversal, nearer child first.
template <class Scene>
RT_DEV void trace_sah(const Scene& S, const Ray& r, const RayPre& pre, float tmin, float& tbest, int& sbest) {
  for (int s = 0; s < S.n_unbounded; ++s) {
    int type = (ldgi2(S.slot_info + s).y >> 30) & 3;
    test_slot(S, s, type, ldg4(S.p0 + s), ldg4(S.p1 + s), ldg4(S.p2 + s), r, pre, tbest, sbest);
  }
  if (S.n_nodes == 0) return;
  int stack[64];
  int sp = 0;
  int cur = 0;
  for (;;) {
    Node4 nd = load_node(S.nodes, cur);
    const float lmn[3] = {nd.a.x, nd.a.y, nd.a.z}, lmx[3] = {nd.a.w, nd.b.x, nd.b.y};
    const float rmn[3] = {nd.b.z, nd.b.w, nd.c.x}, rmx[3] = {nd.c.y, nd.c.z, nd.c.w};
    int lref = __float_as_int(nd.d.x), rref = __float_as_int(nd.d.y);
    float tl, tr;
    bool hl = lref != kEmptyRef && slab_hit(lmn, lmx, pre, tmin, tbest, tl);
    bool hr = rref != kEmptyRef && slab_hit(rmn, rmx, pre, tmin, tbest, tr);
    // leaves are intersected on the spot, nearer one first
    if (hl && hr && tr < tl) {
      if (rref < 0) { intersect_leaf(S, rref, r, pre, tmin, tbest, sbest); hr = false; }
      if (lref < 0) { if (tl <= tbest) intersect_leaf(S, lref, r, pre, tmin, tbest, sbest); hl = false; }
    } else {
      if (hl && lref < 0) { intersect_leaf(S, lref, r, pre, tmin, tbest, sbest); hl = false; }
      if (hr && rref < 0) { if (tr <= tbest) intersect_leaf(S, rref, r, pre, tmin, tbest, sbest); hr = false; }
    }
    if (hl && tl > tbest) hl = false;
    if (hr && tr > tbest) hr = false;
    if (hl && hr) {
      bool lfirst = tl <= tr;
      stack[sp++] = lfirst ? rref : lref;
      cur = lfirst ? lref : rref;
    } else if (hl) cur = lref;
    else if (hr) cur = rref;
    else {
      if (sp == 0) return;
      cur = stack[--sp];
    }
  }
}

// REFERENCE: depth-first, left subtree completely before the right box is (re)tested with
// the shrunken interval — bvh.ts:128-146.
template <class Scene>
RT_DEV void trace_ref(const Scene& S, const Ray& r, const RayPre& pre, float tmin, float& tbest, int& sbest) {
  int stack[64]; // node indices whose RIGHT child is pending
  int sp = 0;
  int cur = 0;
  bool enter_right = false; // cur's left side is done; evaluate its right child
  for (;;) {
    Node4 nd = load_node(S.nodes, cur);
    int next = -1;
    if (!enter_right) {
      const float lmn[3] = {nd.a.x, nd.a.y, nd.a.z}, lmx[3] = {nd.a.w, nd.b.x, nd.b.y};
      int lref = __float_as_int(nd.d.x), rref = __float_as_int(nd.d.y);
      if (rref != kEmptyRef) stack[sp++] = cur;
      if (lref != kEmptyRef && ref_box_hit(lmn, lmx, r, pre, tmin, tbest)) {
        if (lref < 0) intersect_leaf(S, lref, r, pre, tmin, tbest, sbest);
        else next = lref;
      }
    } else {
      const float rmn[3] = {nd.b.z, nd.b.w, nd.c.x}, rmx[3] = {nd.c.y, nd.c.z, nd.c.w};
      int rref = __float_as_int(nd.d.y);
      if (ref_box_hit(rmn, rmx, r, pre, tmin, tbest)) {
        if (rref < 0) intersect_leaf(S, rref, r, pre, tmin, tbest, sbest);
        else next = rref;
      }
    }
    if (next >= 0) { cur = next; enter_right = false; continue; }
    if (sp == 0) return;
    cur = stack[--sp];
    enter_right = true;
  }
}

// ---------------------------------------------------------------------------------------
// Surface at a hit: rec.p, rec.normal (face-forwarded), rec.frontFace
// (sphere.ts:68-84, plane.ts:93-97, quad.ts:64-67).
// ---------------------------------------------------------------------------------------
struct Surf {
  V3 p, n;
  bool front;
};
RT_DEV Surf surface_at(int type, F4 p0, const Ray& r, float t) {
  Surf s;
  s.p = fma3(t, r.d, r.o);
  V3 n;
  if (type == OBJ_SPHERE) n = (s.p - mk3(p0.x, p0.y, p0.z)) * (1.0f / p0.w);
  else n = mk3(p0.x, p0.y, p0.z);
  s.front = dot3(r.d, n) <= 0.f;
  s.n = s.front ? n : mk3(0.f - n.x, 0.f - n.y, 0.f - n.z); // 0-x: negate() canonicalises -0 (vec3.ts:60-70)
  return s;
}

// ---------------------------------------------------------------------------------------
// Sampling — vec3.ts:272-364, onbasis.ts:18-51, pdf.ts:32-51
// ---------------------------------------------------------------------------------------
struct Onb {
  V3 u, v, w;
};
RT_DEV Onb make_onb(V3 n) {
  Onb b;
  b.w = normalize3(n);
  V3 a = fabsf(b.w.x) > 0.9f ? mk3(0, 1, 0) : mk3(1, 0, 0);
  b.v = normalize3(cross3(b.w, a));
  b.u = cross3(b.w, b.v);
  return b;
}
RT_DEV V3 onb_local(const Onb& b, V3 a) { return fma3(a.x, b.u, fma3(a.y, b.v, b.w * a.z)); }
RT_DEV V3 random_cosine_direction(Rng& g) {
  float r1 = g.next(), r2 = g.next();
  float sn, cs;
  sincospif(2.f * r1, &sn, &cs);
  float s = sqrtf(r2);
  return mk3(cs * s, sn * s, sqrtf(1.f - r2));
}
RT_DEV V3 random_in_unit_sphere(Rng& g) {
  for (;;) {
    float x = fmaf(2.f, g.next(), -1.f), y = fmaf(2.f, g.next(), -1.f), z = fmaf(2.f, g.next(), -1.f);
    V3 p = mk3(x, y, z);
    if (dot3(p, p) < 1.f) return p;
  }
}
RT_DEV float cosine_pdf_value(V3 w_unit, V3 dir) {
  float c = dot3(normalize3(dir), w_unit);
  return c <= 0.f ? 0.f : c * 0.31830988618f;
}

// Light pdfs — quad.ts:123-158, sphere.ts:106-147.  A pdf evaluation is one single-primitive
// hit test, never a traced ray.
template <class Scene>
RT_DEV float light_pdf_value(const Scene& S, const DevLight& L, V3 origin, V3 dir) {
  Ray r{origin, dir};
  RayPre pre = precompute(r, false);
  float t;
  if (L.type == OBJ_QUAD) {
    int res = planar_fast(L.p0, L.p1, L.p2, true, r, kRayTMin, CUDART_INF_F, t);
    if (res == 2) res = exact_closer(S.exact, L.slot, -1, CUDART_INF_F, r, t) ? 1 : 0;
    if (res != 1) return 0.f;
    float d2 = t * t * pre.a; // |rec.p - origin|^2
    float cosine = fabsf(dot3(dir, mk3(L.p0.x, L.p0.y, L.p0.z)));
    return d2 / (L.area * cosine);
  }
  int res = sphere_fast(L.p0, r, pre, kRayTMin, CUDART_INF_F, t);
  if (res == 2) res = exact_closer(S.exact, L.slot, -1, CUDART_INF_F, r, t) ? 1 : 0;
  if (res != 1) return 0.f;
  V3 oc = mk3(L.p0.x, L.p0.y, L.p0.z) - origin;
  float d2 = dot3(oc, oc), r2 = L.radius * L.radius;
  if (d2 <= r2) return 0.07957747155f; // 1/(4 pi)
  float cos_theta = sqrtf(1.f - r2 / d2);
  return 1.f / (6.28318530718f * (1.f - cos_theta));
}
RT_DEV V3 light_random_vec(const DevLight& L, V3 origin, Rng& g) {
  if (L.type == OBJ_QUAD) {
    float alpha = g.next(), beta = g.next();
    V3 rp = fma3(beta, ld3(L.v), fma3(alpha, ld3(L.u), ld3(L.q)));
    return normalize3(rp - origin);
  }
  V3 oc = mk3(L.p0.x, L.p0.y, L.p0.z) - origin;
  float d2 = dot3(oc, oc);
  Onb b = make_onb(oc);
  float r1 = g.next(), r2 = g.next();
  float z = 1.f + r2 * (sqrtf(1.f - L.radius * L.radius / d2) - 1.f);
  float sn, cs;
  sincospif(2.f * r1, &sn, &cs);
  float s = sqrtf(1.f - z * z);
  return onb_local(b, mk3(cs * s, sn * s, z));
}

// ---------------------------------------------------------------------------------------
// Materials — one stackless walk over the material node table.  Mixed and Layered each
// tail-call exactly one child (mixedMaterial.ts:38-45, layeredMaterial.ts:36-53), so the
// recursion of the reference is a loop here.
// ---------------------------------------------------------------------------------------
enum : int { SCATTER_NONE = 0, SCATTER_SPECULAR = 1, SCATTER_DIFFUSE = 2 };
struct Scatter {
  int kind;
  V3 attenuation;
  V3 dir; // specular: scattered direction
};
RT_DEV V3 reflect3(V3 v, V3 n) { return fma3(-2.f * dot3(v, n), n, v); }

// dielectric.ts:44-84.  Returns true when the ray was reflected.
RT_DEV bool dielectric_dir(V3 din, V3 n, bool front, float ior, Rng& g, V3& out) {
  float ratio = front ? 1.0f / ior : ior;
  V3 ud = normalize3(din);
  float cos_t = fminf(-dot3(ud, n), 1.0f);
  float sin_t = sqrtf(fmaxf(0.f, 1.0f - cos_t * cos_t));
  bool cannot = ratio * sin_t > 1.0f;
  bool refl = cannot;
  if (!cannot) {
    float r0 = (1.f - ratio) / (1.f + ratio);
    r0 *= r0;
    float m = 1.f - cos_t, m2 = m * m;
    float reflectance = fmaf(1.f - r0, m2 * m2 * m, r0);
    refl = reflectance > g.next();
  }
  if (refl) out = reflect3(ud, n);
  else { // vec3.ts:193-209
    V3 perp = fma3(cos_t, n, ud) * ratio;
    float k = -sqrtf(fabsf(1.0f - dot3(perp, perp)));
    out = fma3(k, n, perp);
  }
  return refl;
}

template <class Scene>
RT_DEV Scatter scatter_material(const Scene& S, int root, V3 din, const Surf& sf, Rng& g) {
  Scatter out;
  out.kind = SCATTER_NONE;
  out.attenuation = mk3(0, 0, 0);
  out.dir = mk3(0, 0, 0);
  int node = root;
  for (;;) {
    I4 b = ldgi4(S.matB + node);
    F4 a = ldg4(S.matA + node);
    switch (b.x) {
      case MAT_MIXED:
        node = g.next() < a.w ? b.y : b.z;
        continue;
      case MAT_LAYERED: {
        V3 d;
        if (dielectric_dir(din, sf.n, sf.front, a.w, g, d)) {
          out.kind = SCATTER_SPECULAR;
          out.attenuation = mk3(1, 1, 1);
          out.dir = d;
          return out;
        }
        din = d; // inner.scatter(refracted ray, same rec)
        node = b.y;
        continue;
      }
      case MAT_LAMBERT:
        out.kind = SCATTER_DIFFUSE;
        out.attenuation = mk3(a.x, a.y, a.z);
        return out;
      case MAT_METAL: { // metal.ts:29-50
        V3 refl = reflect3(normalize3(din), sf.n);
        if (a.w > 0.f) refl = fma3(a.w, random_in_unit_sphere(g), refl);
        if (dot3(refl, sf.n) <= 0.f) return out; // absorbed
        out.kind = SCATTER_SPECULAR;
        out.attenuation = mk3(a.x, a.y, a.z);
        out.dir = refl;
        return out;
      }
      case MAT_GLASS: {
        V3 d;
        dielectric_dir(din, sf.n, sf.front, a.w, g, d);
        out.kind = SCATTER_SPECULAR;
        out.attenuation = mk3(1, 1, 1);
        out.dir = d;
        return out;
      }
      default: // MAT_LIGHT: DefaultMaterial.scatter -> null (material.ts:50-52)
        return out;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Camera rays — camera.ts:176-210 with the reference's FP32 rounding sequence
// (scale then add, each rounded: __fmul_rn/__fadd_rn block FMA contraction), so the ray is
// bit-identical to the oracle's for the same random numbers.
// ---------------------------------------------------------------------------------------
RT_DEV V3 mul_rn(V3 a, float s) { return V3{__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)}; }
RT_DEV V3 add_rn(V3 a, V3 b) { return V3{__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)}; }
RT_DEV V3 sub_rn(V3 a, V3 b) { return V3{__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)}; }

RT_DEV Ray camera_ray(const DevCamera& c, int i, int j, Rng& g, bool jitter_and_defocus) {
  V3 p00 = ld3(c.p00), du = ld3(c.du), dv = ld3(c.dv), center = ld3(c.center);
  V3 pc = add_rn(add_rn(p00, mul_rn(du, (float)i)), mul_rn(dv, (float)j));
  V3 ps = pc;
  if (jitter_and_defocus && c.jitter) {
    float px = -0.5f + g.next();
    float py = -0.5f + g.next();
    ps = add_rn(add_rn(pc, mul_rn(du, px)), mul_rn(dv, py));
  }
  Ray r;
  r.o = center;
  r.d = sub_rn(ps, center);
  if (jitter_and_defocus && c.defocus) {
    float x, y;
    do { // vec3.ts:357-364
      x = __fadd_rn(__fmul_rn(2.f, g.next()), -1.f);
      y = __fadd_rn(__fmul_rn(2.f, g.next()), -1.f);
    } while (!(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)) < 1.f));
    V3 off = add_rn(mul_rn(ld3(c.ddu), x), mul_rn(ld3(c.ddv), y));
    r.o = add_rn(center, off);
    r.d = sub_rn(ps, r.o);
  }
  return r;
}

} // namespace rt
