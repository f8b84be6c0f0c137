// rt_debug.cuh — per-function parity hooks (rt_debug_* of include/rt_b200.h).
//
// Each kernel runs ONE device function of the render path — the very template instantiated by the render
// kernels, not a copy — on explicit inputs, with the uniforms Math.random() would return supplied by the
// caller (ListRng) instead of a Philox stream.  tests/test_gpu_functions.py feeds the same inputs to the
// CPU oracle's *_u hooks and to these, on the reference's own Jest vectors.  One thread per record; nothing
// here is on the render path.
#pragma once
#include "rt_device.cuh"

namespace rt {

struct DbgHit { // a synthetic HitRecord + the incoming ray (src/geometry/hittable.ts:9-18)
  float ro[3], rd[3], p[3], n[3];
  int front, pad;
};
struct DbgScatterOut { // mirrors rt_debug_scatter_out
  int kind, used;
  float att[3], dir[3], emitted[3];
};
static constexpr int kDbgUniforms = 16; // RT_DEBUG_UNIFORMS

// material.scatter(rIn, rec) + material.emitted(rec) for material node `root` (src/materials/*.ts)
__global__ void k_debug_scatter(const DevScene S, int root, int n, const DbgHit* hits, const float* uniforms, DbgScatterOut* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const DbgHit h = hits[i];
  ListRng g{uniforms + (size_t)i * kDbgUniforms, kDbgUniforms, 0};
  Surf sf{ld3(h.p), ld3(h.n), h.front != 0};
  const I4 mb = ldgi4(S.matB + root);
  const F4 ma = ldg4(S.matA + root);
  const F4 me = ldg4(S.matE + root);
  Scatter sc;
  // the same dispatch as path_post (rt_megakernel.cu): plain Lambert and lights never enter the material walk
  if (mb.x == MAT_LAMBERT) { sc.kind = SCATTER_DIFFUSE; sc.attenuation = xyz(ma); sc.dir = mk3(0, 0, 0); }
  else if (mb.x == MAT_LIGHT) { sc.kind = SCATTER_NONE; sc.attenuation = mk3(0, 0, 0); sc.dir = mk3(0, 0, 0); }
  else sc = scatter_material(S, root, mb, ma, ld3(h.rd), sf, g);
  DbgScatterOut o;
  o.kind = sc.kind;
  o.used = g.used;
  o.att[0] = sc.attenuation.x; o.att[1] = sc.attenuation.y; o.att[2] = sc.attenuation.z;
  o.dir[0] = sc.dir.x; o.dir[1] = sc.dir.y; o.dir[2] = sc.dir.z;
  o.emitted[0] = me.x; o.emitted[1] = me.y; o.emitted[2] = me.z;
  out[i] = o;
}

// Camera.getRay(i, j) (src/camera.ts:176-210): out = origin.xyz, dir.xyz; used = uniforms consumed
__global__ void k_debug_get_ray(const DevScene S, int n, const int* ij, const float* uniforms, float* out, int* used) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  ListRng g{uniforms + (size_t)k * kDbgUniforms, kDbgUniforms, 0};
  const Ray r = camera_ray(S.cam, ij[2 * k], ij[2 * k + 1], g, true);
  float* o = out + 6 * (size_t)k;
  o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
  used[k] = g.used;
}

// lights[light].pdfValue(origin, direction) (src/entities/quad.ts:123-140, sphere.ts:106-131)
__global__ void k_debug_light_pdf(const DevScene S, int light, int n, const float* origin, const float* dir, float* value) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  value[k] = light_pdf_value(S, S.lights[light], ld3(origin + 3 * (size_t)k), ld3(dir + 3 * (size_t)k));
}

// lights[light].pdfRandomVec(origin) (src/entities/quad.ts:148-158, sphere.ts:140-147); uniforms[0..1] = (r1, r2)
__global__ void k_debug_light_random(const DevScene S, int light, int n, const float* origin, const float* uniforms, float* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float* u = uniforms + (size_t)k * kDbgUniforms;
  const V3 v = light_random_vec(S.lights[light], ld3(origin + 3 * (size_t)k), u[0], u[1]);
  out[3 * (size_t)k] = v.x; out[3 * (size_t)k + 1] = v.y; out[3 * (size_t)k + 2] = v.z;
}

// The diffuse branch of rayColor (src/camera.ts:285-308): out = dir.xyz, mixture pdf value, scatter pdf value,
// continues (pdf value > 0.0001); uniforms[0..2] = (component select, r1, r2)
__global__ void k_debug_diffuse(const DevScene S, int n, const float* p, const float* nrm, const float* uniforms, float* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const MixW mw = make_mixw(S);
  const float* u = uniforms + (size_t)k * kDbgUniforms;
  V3 dir;
  float cosv, pdf_value;
  diffuse_bounce(S, mw, ld3(p + 3 * (size_t)k), ld3(nrm + 3 * (size_t)k), u[0], u[1], u[2], dir, cosv, pdf_value);
  float* o = out + 6 * (size_t)k;
  o[0] = dir.x; o[1] = dir.y; o[2] = dir.z; o[3] = pdf_value; o[4] = cosv; o[5] = pdf_value > 0.0001f ? 1.f : 0.f;
}

static inline int dbg_blocks(int n) { return (n + 127) / 128; }
cudaError_t launch_debug_scatter(const DevScene& S, int root, int n, const void* hits, const float* uniforms, void* out, cudaStream_t st) {
  k_debug_scatter<<<dbg_blocks(n), 128, 0, st>>>(S, root, n, (const DbgHit*)hits, uniforms, (DbgScatterOut*)out);
  return cudaGetLastError();
}
cudaError_t launch_debug_get_ray(const DevScene& S, int n, const int* ij, const float* uniforms, float* out, int* used, cudaStream_t st) {
  k_debug_get_ray<<<dbg_blocks(n), 128, 0, st>>>(S, n, ij, uniforms, out, used);
  return cudaGetLastError();
}
cudaError_t launch_debug_light_pdf(const DevScene& S, int light, int n, const float* origin, const float* dir, float* value, cudaStream_t st) {
  k_debug_light_pdf<<<dbg_blocks(n), 128, 0, st>>>(S, light, n, origin, dir, value);
  return cudaGetLastError();
}
cudaError_t launch_debug_light_random(const DevScene& S, int light, int n, const float* origin, const float* uniforms, float* out, cudaStream_t st) {
  k_debug_light_random<<<dbg_blocks(n), 128, 0, st>>>(S, light, n, origin, uniforms, out);
  return cudaGetLastError();
}
cudaError_t launch_debug_diffuse(const DevScene& S, int n, const float* p, const float* nrm, const float* uniforms, float* out, cudaStream_t st) {
  k_debug_diffuse<<<dbg_blocks(n), 128, 0, st>>>(S, n, p, nrm, uniforms, out);
  return cudaGetLastError();
}

} // namespace rt
