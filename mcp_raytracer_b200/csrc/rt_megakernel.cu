// rt_megakernel.cu — register-resident path tracer ("megakernel" integrator) and the
// primary-visibility parity kernel, for sm_100a.
//
// Mapping: one CTA = one 16x16 image tile, one thread = one pixel, a warp = an 8x4 pixel
// block (coherent primary rays).  Each thread runs the reference's per-pixel loop
// (src/camera.ts:400-423): sample until `samples` or adaptive convergence, accumulate, write.
// The sample loop and the bounce recursion (src/camera.ts:221-319) are flattened into ONE
// loop with per-lane path regeneration: a lane whose path ended starts its next sample in
// the same iteration in which its neighbours trace their next bounce, so every iteration
// every live lane generates one Philox block, traces exactly one ray and shades one hit, and
// no lane idles waiting for the longest path of a sample.  All path state lives in registers;
// HBM sees the scene reads (L1/L2 resident) and 3 bytes per pixel of output.
#include "rt_device.cuh"

namespace rt {

static constexpr int kTile = 16;
static constexpr int kListMax = 64;
#ifndef RT_MIN_BLOCKS
#define RT_MIN_BLOCKS 2
#endif

RT_DEV unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
RT_DEV int warp_min(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
RT_DEV int warp_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// thread -> pixel inside the tile: warp w covers an 8x4 block
RT_DEV void tile_pixel(int& px, int& py) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  px = (w & 1) * 8 + (lane & 7);
  py = (w >> 1) * 4 + (lane >> 3);
}

template <int KIND>
RT_DEV bool closest_hit(const DevScene& S, const SmemList& L, const Ray& r, float& t, int& slot) {
  RayPre pre = precompute(r, KIND != BVH_LIST);
  t = CUDART_INF_F;
  slot = -1;
  if (KIND == BVH_LIST) trace_list(S, L, r, pre, t, slot);
  else if (KIND == BVH_SAH) trace_sah(S, r, pre, t, slot);
  else trace_ref(S, r, pre, t, slot);
  return slot >= 0;
}

struct ListSmem {
  F4 p0[kListMax], p1[kListMax], p2[kListMax], p3[kListMax];
  int type[kListMax];
  int mat[kListMax];
};

template <int KIND>
RT_DEV SmemList stage_list(const DevScene& S, ListSmem& sm) {
  SmemList L{nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  if (KIND == BVH_LIST) {
    for (int s = threadIdx.x; s < S.n_slots; s += blockDim.x) {
      sm.p0[s] = ldg4(S.p0 + s);
      sm.p1[s] = ldg4(S.p1 + s);
      sm.p2[s] = ldg4(S.p2 + s);
      sm.p3[s] = ldg4(S.p3 + s);
      I2 info = ldgi2(S.slot_info + s);
      sm.type[s] = (info.y >> 30) & 3;
      sm.mat[s] = info.x;
    }
    __syncthreads();
    L = SmemList{sm.p0, sm.p1, sm.p2, sm.p3, sm.type, S.n_slots};
  }
  return L;
}

// =========================================================================================
// render kernel.  FULL = adaptive sampling / render modes / moments output compiled in;
// the lean variant is the fixed-spp default-mode path the benchmark configs run.
// =========================================================================================
template <int KIND, bool FULL>
__global__ void __launch_bounds__(256, RT_MIN_BLOCKS) k_render_mega(const DevScene S, const RenderParams R) {
  __shared__ ListSmem sm;
  const SmemList L = stage_list<KIND>(S, sm);
  const DevCamera& cam = S.cam;

  const int tx = (R.x0 / kTile) + blockIdx.x, ty = (R.y0 / kTile) + blockIdx.y;
  const bool owned = R.part_count <= 1 || ((tx + ty) % R.part_count) == R.part_index;
  int lx, ly;
  tile_pixel(lx, ly);
  const int i = tx * kTile + lx, j = ty * kTile + ly;
  const bool active = owned && i >= R.x0 && i < R.x1 && j >= R.y0 && j < R.y1;
  const uint32_t pixel = (uint32_t)j * (uint32_t)cam.width + (uint32_t)i;

  // PixelStats (renderStats.ts:67-88)
  V3 color = mk3(0, 0, 0);
  int samples = 0;
  unsigned int bounces_sum = 0, rays = 0; // per pixel: < 2^32 for any sane spp * depth
  int min_b = 0x7fffffff, max_b = 0;
  double sum_ill = 0, sum_ill2 = 0;
  float m2x = 0, m2y = 0, m2z = 0; // sum of squares for the optional moments output

  // path state
  Ray ray{mk3(0, 0, 0), mk3(0, 0, 1)};
  V3 tp = mk3(1, 1, 1), radiance = mk3(0, 0, 0);
  int bounces = 0;
  bool need_path = true;
  Rng g;
  const int nl = S.n_lights;
  const float wl = nl > 0 ? 0.5f / (float)nl : 0.f;
  float total_w = 0.5f; // MixturePDF totalWeight: 0.5 + n * (0.5/n), summed like pdf.ts:71
  for (int k = 0; k < nl; ++k) total_w += wl;
  const float inv_total_w = 1.0f / total_w;

  while (active) {
    if (need_path) {
      // while (pixel.samples < samples && !pixelConverged(pixel)) — camera.ts:406
      bool stop = samples >= cam.samples;
      if (FULL && !stop && cam.adaptive && samples >= 2 && (samples % cam.a_batch) == 0) { // camera.ts:348-368
        double n = (double)samples;
        double mean = sum_ill / n;
        double var = (sum_ill2 - (sum_ill * sum_ill) / n) / (n - 1.0);
        if (var <= 0.0 || var != var) stop = true;
        else stop = 1.96 * sqrt(var) / sqrt(n) <= (double)cam.a_tol * mean;
      }
      if (stop) break;
      tp = mk3(1, 1, 1);
      radiance = mk3(0, 0, 0);
      bounces = 0;
    }
    // one Philox block per bounce, generated by all lanes together
    g.begin(pixel, (uint32_t)samples, (uint32_t)bounces, S.seed_lo, S.seed_hi);
    if (need_path) {
      ray = camera_ray(cam, i, j, g, true);
      need_path = false;
    }
    // ---- one rayColor call (camera.ts:221-319) ----
    bool done = bounces >= cam.depth;
    if (!done && cam.roulette && bounces >= cam.rr_depth) { // camera.ts:233-245
      float p = fminf(maxc(tp), 0.95f);
      done = g.next() > p;
      tp = tp * (1.0f / p);
    }
    if (!done) {
      float t;
      int slot;
      ++rays;
      if (!closest_hit<KIND>(S, L, ray, t, slot)) { // camera.ts:252-258
        V3 ud = normalize3(ray.d);
        float a = 0.5f * (ud.y + 1.0f);
        V3 bg = ld3(cam.bg_top) * (1.0f - a) + ld3(cam.bg_bottom) * a;
        radiance = radiance + bg * tp;
        done = true;
      } else {
        int type, root;
        F4 p0;
        if (KIND == BVH_LIST) { type = sm.type[slot]; root = sm.mat[slot]; p0 = sm.p0[slot]; }
        else { I2 info = ldgi2(S.slot_info + slot); type = (info.y >> 30) & 3; root = info.x; p0 = ldg4(S.p0 + slot); }
        const Surf sf = surface_at(type, p0, ray, t);
        const I4 mb = ldgi4(S.matB + root);
        const F4 ma = ldg4(S.matA + root);
        if (mb.w) { // emitted * throughput (camera.ts:261)
          F4 e = ldg4(S.matE + root);
          radiance = radiance + xyz(e) * tp;
        }
        Scatter sc;
        if (mb.x == MAT_LAMBERT) { sc.kind = SCATTER_DIFFUSE; sc.attenuation = xyz(ma); sc.dir = mk3(0, 0, 0); }
        else if (mb.x == MAT_LIGHT) { sc.kind = SCATTER_NONE; sc.attenuation = mk3(0, 0, 0); sc.dir = mk3(0, 0, 0); }
        else sc = scatter_material(S, root, mb, ma, ray.d, sf, g);
        if (sc.kind == SCATTER_NONE) done = true; // camera.ts:267-269 (bounce not counted)
        else {
          ++bounces;
          if (sc.kind == SCATTER_SPECULAR) { // camera.ts:275-282
            tp = tp * sc.attenuation;
            ray = Ray{sf.p, sc.dir};
          } else { // camera.ts:285-315 with MixturePDF([cosine, lights...], [0.5, 0.5/n ...]) — pdf.ts:57-99
            const Onb onb = make_onb(sf.n);
            const float rnd = g.next() * total_w;
            const float r1 = g.next(), r2 = g.next();
            V3 dir = onb_local(onb, cosine_direction(r1, r2));
            if (nl > 0) {
              float partial = 0.5f;
              int chosen = nl - 1;
              for (int k = 0; k < nl; ++k) {
                partial += wl;
                if (rnd < partial) { chosen = k; break; }
              }
              V3 ldir = light_random_vec(S.lights[chosen], sf.p, r1, r2);
              dir = sel3(rnd < 0.5f, dir, ldir);
            }
            const float cz = dot3(dir, onb.w); // all three generators return unit vectors
            const float cosv = cz <= 0.f ? 0.f : cz * 0.31830988618f;
            float sum = 0.5f * cosv;
            for (int k = 0; k < nl; ++k) sum = fmaf(wl, light_pdf_value(S, S.lights[k], sf.p, dir), sum);
            const float pdf_value = sum * inv_total_w;
            if (!(pdf_value > 0.0001f)) done = true; // camera.ts:298-301 (NaN also ends the path)
            else {
              tp = tp * (sc.attenuation * (cosv * rcp_approx(pdf_value)));
              ray = Ray{sf.p, dir};
            }
          }
        }
      }
    }
    if (done) { // pixel.add(rayColor, bounces, useAdaptiveSampling) — renderStats.ts:76-88
      color = color + radiance;
      ++samples;
      bounces_sum += (unsigned)bounces;
      min_b = min(min_b, bounces);
      max_b = max(max_b, bounces);
      if (FULL) {
        if (cam.adaptive) {
          double il = 0.299 * (double)radiance.x + 0.587 * (double)radiance.y + 0.114 * (double)radiance.z;
          sum_ill += il;
          sum_ill2 += il * il;
        }
        if (R.moments) {
          m2x = fmaf(radiance.x, radiance.x, m2x);
          m2y = fmaf(radiance.y, radiance.y, m2y);
          m2z = fmaf(radiance.z, radiance.z, m2z);
        }
      }
      need_path = true;
    }
  }

  if (active) {
    // finalColor (camera.ts:326-340) + writeColorToBuffer (camera.ts:455-472)
    V3 fc;
    if (FULL && cam.mode == 1) {
      float avg = samples > 0 ? (float)((double)bounces_sum / (double)samples) : 0.f;
      fc = mk3(0, 0, fminf(avg / (float)cam.depth, 1.0f));
    } else if (FULL && cam.mode == 2) {
      fc = mk3(fminf((float)samples / (float)cam.samples, 1.0f), 0, 0);
    } else {
      float inv = (float)(1.0 / (double)samples);
      fc = color * inv;
    }
    const size_t pi = (size_t)j * cam.width + i;
    if (R.rgb8) {
      const float c[3] = {fc.x, fc.y, fc.z};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = floor(255.999 * sqrt((double)c[k]));
        // Uint8ClampedArray store: NaN/negative -> 0, >255 -> 255
        R.rgb8[pi * 3 + k] = !(v > 0.0) ? 0 : (v >= 255.0 ? 255 : (uint8_t)v);
      }
    }
    if (R.linear) { R.linear[pi * 3] = fc.x; R.linear[pi * 3 + 1] = fc.y; R.linear[pi * 3 + 2] = fc.z; }
    if (FULL && R.moments) {
      float* m = R.moments + pi * 8;
      m[0] = color.x; m[1] = color.y; m[2] = color.z; m[3] = m2x; m[4] = m2y; m[5] = m2z;
      m[6] = (float)samples; m[7] = (float)bounces_sum;
    }
  }

  // RenderStats.addPixel (renderStats.ts:21-35), reduced per warp then one atomic each
  if (R.stats) {
    unsigned long long px = warp_sum(active ? 1ull : 0ull);
    unsigned long long ss = warp_sum(active ? (unsigned long long)samples : 0ull);
    unsigned long long bs = warp_sum(active ? (unsigned long long)bounces_sum : 0ull);
    unsigned long long rs = warp_sum(active ? (unsigned long long)rays : 0ull);
    int smin = warp_min(active ? samples : 0x7fffffff), smax = warp_max(active ? samples : 0);
    int bmin = warp_min(active ? min_b : 0x7fffffff), bmax = warp_max(active ? max_b : 0);
    if ((threadIdx.x & 31) == 0 && px) {
      atomicAdd(R.stats + kStatPixels, px);
      atomicAdd(R.stats + kStatSamples, ss);
      atomicAdd(R.stats + kStatBounces, bs);
      atomicAdd(R.stats + kStatRays, rs);
      atomicMin(R.stats + kStatSamplesMin, (unsigned long long)smin);
      atomicMax(R.stats + kStatSamplesMax, (unsigned long long)smax);
      atomicMin(R.stats + kStatBouncesMin, (unsigned long long)bmin);
      atomicMax(R.stats + kStatBouncesMax, (unsigned long long)bmax);
    }
  }
}

// =========================================================================================
// primary visibility (parity hook): pixel-centre rays, no jitter, no defocus
// =========================================================================================
template <int KIND>
__global__ void __launch_bounds__(256) k_trace_primary(const DevScene S, const RenderParams R, int* obj_id, float* t_out,
                                                       float* normal, uint8_t* front) {
  __shared__ ListSmem sm;
  const SmemList L = stage_list<KIND>(S, sm);
  const int tx = (R.x0 / kTile) + blockIdx.x, ty = (R.y0 / kTile) + blockIdx.y;
  int lx, ly;
  tile_pixel(lx, ly);
  const int i = tx * kTile + lx, j = ty * kTile + ly;
  if (!(i >= R.x0 && i < R.x1 && j >= R.y0 && j < R.y1)) return;
  Rng g{};
  Ray ray = camera_ray(S.cam, i, j, g, false);
  float t;
  int slot;
  bool hit = closest_hit<KIND>(S, L, ray, t, slot);
  const size_t pi = (size_t)j * S.cam.width + i;
  Surf sf{mk3(0, 0, 0), mk3(0, 0, 0), false};
  int obj = -1;
  if (hit) {
    I2 info = ldgi2(S.slot_info + slot);
    obj = info.y & 0x3fffffff;
    sf = surface_at((info.y >> 30) & 3, ldg4(S.p0 + slot), ray, t);
  }
  if (obj_id) obj_id[pi] = obj;
  if (t_out) t_out[pi] = hit ? t : CUDART_INF_F;
  if (normal) { normal[pi * 3] = sf.n.x; normal[pi * 3 + 1] = sf.n.y; normal[pi * 3 + 2] = sf.n.z; }
  if (front) front[pi] = hit && sf.front ? 1 : 0;
}

// =========================================================================================
// FP32 peak microbenchmark: 8 independent FFMA chains per thread
// =========================================================================================
__global__ void __launch_bounds__(256) k_fp32_peak(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// =========================================================================================
// launchers (called from rt_api.cu)
// =========================================================================================
static dim3 tile_grid(const RenderParams& R) {
  int tx0 = R.x0 / kTile, ty0 = R.y0 / kTile;
  int tx1 = (R.x1 - 1) / kTile, ty1 = (R.y1 - 1) / kTile;
  return dim3((unsigned)(tx1 - tx0 + 1), (unsigned)(ty1 - ty0 + 1), 1);
}

template <bool FULL>
static void launch_mega_kind(const DevScene& S, const RenderParams& R, dim3 grid, cudaStream_t st) {
  dim3 block(256);
  switch (S.bvh_kind) {
    case BVH_LIST: k_render_mega<BVH_LIST, FULL><<<grid, block, 0, st>>>(S, R); break;
    case BVH_SAH: k_render_mega<BVH_SAH, FULL><<<grid, block, 0, st>>>(S, R); break;
    default: k_render_mega<BVH_REFERENCE, FULL><<<grid, block, 0, st>>>(S, R); break;
  }
}

cudaError_t launch_render_mega(const DevScene& S, const RenderParams& R, cudaStream_t st) {
  if (R.x1 <= R.x0 || R.y1 <= R.y0) return cudaSuccess;
  dim3 grid = tile_grid(R);
  const bool full = S.cam.adaptive || S.cam.mode != 0 || R.moments != nullptr;
  if (full) launch_mega_kind<true>(S, R, grid, st);
  else launch_mega_kind<false>(S, R, grid, st);
  return cudaGetLastError();
}

cudaError_t launch_trace_primary(const DevScene& S, const RenderParams& R, int* obj_id, float* t, float* normal,
                                 uint8_t* front, cudaStream_t st) {
  if (R.x1 <= R.x0 || R.y1 <= R.y0) return cudaSuccess;
  dim3 grid = tile_grid(R), block(256);
  switch (S.bvh_kind) {
    case BVH_LIST: k_trace_primary<BVH_LIST><<<grid, block, 0, st>>>(S, R, obj_id, t, normal, front); break;
    case BVH_SAH: k_trace_primary<BVH_SAH><<<grid, block, 0, st>>>(S, R, obj_id, t, normal, front); break;
    default: k_trace_primary<BVH_REFERENCE><<<grid, block, 0, st>>>(S, R, obj_id, t, normal, front); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_fp32_peak(float* out, int blocks, int iters, cudaStream_t st) {
  k_fp32_peak<<<blocks, 256, 0, st>>>(out, iters, 0.999f, 0.001f);
  return cudaGetLastError();
}

} // namespace rt
