// rt_wavefront.cuh — wavefront integrator (RT_INTEGRATOR_WAVEFRONT), included at the end of
// rt_megakernel.cu so that it shares path_pre / path_post / closest_hit and the primitive tests: the
// two integrators run the same arithmetic on the same Philox streams and produce the same image.
//
// A pool of path slots lives in HBM (SoA, 80 B per slot).  One iteration of the host loop runs
//   k_wf_generate   every free slot takes the next (pixel, sample) pair, shoots the camera ray, applies the
//                   depth / roulette test of bounce 0 and queues the ray
//   k_wf_extend     persistent threads; closest hit for every queued ray.  Tree scenes: lanes pull ray ids
//                   from the queue on demand (warp-aggregated atomic) and traverse in converged
//                   inner-node / leaf rounds, so a finished lane takes a new ray instead of waiting for
//                   the warp's longest traversal.  Then each ray is binned by the material tag at its hit
//                   — miss or emitter-only (terminal), Lambertian, everything else — with one
//                   warp-ballot/prefix-compacted append per tag
//   k_wf_shade<TAG> one launch per tag over its compacted queue: emission, scatter, mixture-pdf light
//                   sampling (the light's single-primitive visibility test of quad.ts:123-140 lives
//                   here — the reference has no shadow rays), throughput update, then the depth / roulette
//                   test of the next bounce; surviving paths are queued for the next extend, finished
//                   ones add their radiance to the pixel's fixed-point accumulator and free the slot
// until no samples are left and every slot is free; k_wf_resolve then writes the pixels.
//
// Why it is not the default: per bounce it moves ~200 B of path state per ray through HBM/L2 and needs
// 5 launches per iteration, while the megakernel keeps the same state in registers; see
// profiles/README.md for the measured comparison on the BASELINE configs.
#pragma once

namespace rt {

enum : int { WF_TAG_TERMINAL = 0, WF_TAG_LAMBERT = 1, WF_TAG_OTHER = 2 };

// float4 / uint2 views of the POD arrays declared in rt_types.h (WfBuffers)
struct WfView {
  float4* ray_o;   // o.xyz, hit t
  float4* ray_d;   // d.xyz, hit slot (int bits)
  float4* tp;      // throughput.xyz, bounces (int bits)
  float4* rad;     // radiance.xyz, draws already taken from the current bounce's stream (int bits)
  uint2* pix;      // pixel index, sample index
};
RT_DEV WfView wf_view(const WfBuffers& W) {
  return WfView{reinterpret_cast<float4*>(W.ray_o), reinterpret_cast<float4*>(W.ray_d), reinterpret_cast<float4*>(W.tp),
                reinterpret_cast<float4*>(W.rad), reinterpret_cast<uint2*>(W.pix)};
}

// ---- pair index -> pixel of this GPU's share of the region ---------------------------------------
// Pairs are enumerated block-major over the 8x4 blocks this GPU owns so that consecutive slots start on
// neighbouring pixels of the same sample.  owned_blocks[] lists the block ids (host-built).
struct WfShare {
  const int* owned_blocks;
  int n_owned;
  int blocks_x;
};

RT_DEV void wf_store_path(const WfBuffers& W, int id, const PathState& ps, int used) {
  wf_view(W).ray_o[id] = make_float4(ps.ray.o.x, ps.ray.o.y, ps.ray.o.z, 0.f);
  wf_view(W).ray_d[id] = make_float4(ps.ray.d.x, ps.ray.d.y, ps.ray.d.z, 0.f);
  wf_view(W).tp[id] = make_float4(ps.tp.x, ps.tp.y, ps.tp.z, __int_as_float(ps.bounces));
  wf_view(W).rad[id] = make_float4(ps.radiance.x, ps.radiance.y, ps.radiance.z, __int_as_float(used));
}

// A finished path: pixel.add(rayColor, bounces) — exact fixed-point sums, stats.
RT_DEV void wf_finish(const RenderParams& R, const DevCamera& cam, uint2 px, const PathState& ps, unsigned& st_paths,
                      unsigned& st_bounces, int& st_bmin, int& st_bmax) {
  unsigned long long* a = R.accum + (size_t)px.x * 4;
  atomicAdd(a + 0, to_fixed(ps.radiance.x));
  atomicAdd(a + 1, to_fixed(ps.radiance.y));
  atomicAdd(a + 2, to_fixed(ps.radiance.z));
  ++st_paths;
  st_bounces += (unsigned)ps.bounces;
  st_bmin = min(st_bmin, ps.bounces);
  st_bmax = max(st_bmax, ps.bounces);
}

// warp-aggregated append of `id` to a queue for the lanes with pred set
RT_DEV void wf_push(int* queue, int* counter, bool pred, int id) {
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m == 0) return;
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs(m) - 1;
  int base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) queue[base + __popc(m & ((1u << lane) - 1u))] = id;
}

// -------------------------------------------------------------------------------------------------
// generate: free slots take new (pixel, sample) pairs
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wf_generate(const DevScene S, const RenderParams R, const WfBuffers W, const WfShare sh,
                                                     int n_free, const int* q_free_in) {
  const DevCamera& cam = S.cam;
  unsigned st_paths = 0, st_bounces = 0;
  int st_bmin = 0x7fffffff, st_bmax = 0;
  const int stride = gridDim.x * blockDim.x;
  for (int base_i = blockIdx.x * blockDim.x; base_i < n_free; base_i += stride) {
    const int i = base_i + threadIdx.x;
    const bool lane_on = i < n_free;
    // warp-aggregated grab of pair indices
    const unsigned m = __ballot_sync(0xffffffffu, lane_on);
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long pair = 0;
    if (m) {
      const int leader = __ffs(m) - 1;
      unsigned long long b = 0;
      if ((int)lane == leader) b = atomicAdd(reinterpret_cast<unsigned long long*>(W.counters + WFC_NEXT_PAIR), (unsigned long long)__popc(m));
      b = __shfl_sync(0xffffffffu, b, leader);
      pair = b + __popc(m & ((1u << lane) - 1u));
    }
    bool queued = false, still_free = false;
    int id = 0;
    if (lane_on) {
      id = q_free_in[i];
      if (pair >= W.total_pairs) still_free = false; // no work left: the slot simply retires (not re-queued)
      else {
        // pair -> (block, sample, pixel in block): 32 consecutive pairs = one 8x4 block at one sample
        const unsigned long long per_block = 32ull * (unsigned)cam.samples;
        const int ob = (int)(pair / per_block);
        const unsigned rem = (unsigned)(pair - (unsigned long long)ob * per_block);
        const int sample = (int)(rem >> 5), lp = (int)(rem & 31u);
        const int blk = sh.owned_blocks[ob];
        const int bx = blk % sh.blocks_x, by = blk / sh.blocks_x;
        const int pi_x = (R.x0 / kTile) * kTile + bx * 8 + (lp & 7), pi_y = (R.y0 / kTile) * kTile + by * 4 + (lp >> 3);
        if (pi_x >= R.x0 && pi_x < R.x1 && pi_y >= R.y0 && pi_y < R.y1) {
          const uint32_t pixel = (uint32_t)pi_y * (uint32_t)cam.width + (uint32_t)pi_x;
          PathState ps{Ray{mk3(0, 0, 0), mk3(0, 0, 1)}, mk3(1, 1, 1), mk3(0, 0, 0), 0};
          Rng g;
          g.begin(pixel, (uint32_t)sample, 0u, S.seed_lo, S.seed_hi);
          ps.ray = camera_ray(cam, pi_x, pi_y, g, true);
          const uint2 px = make_uint2(pixel, (uint32_t)sample);
          wf_view(W).pix[id] = px;
          if (path_pre(cam, ps, g)) { // depth 0 or roulette at bounce 0
            wf_finish(R, cam, px, ps, st_paths, st_bounces, st_bmin, st_bmax);
            still_free = true;
          } else {
            wf_store_path(W, id, ps, (int)((g.block - 1u) * 5u + (5u - (unsigned)g.avail)));
            queued = true;
          }
        } else still_free = true; // pixel outside the region: take another pair next iteration
      }
    }
    wf_push(W.q_extend, W.counters + WFC_N_EXTEND, queued, id);
    wf_push(W.q_free, W.counters + WFC_N_FREE, still_free, id);
  }
  flush_stats(R, 0u, cam.samples, st_paths, st_bounces, 0u, st_bmin, st_bmax);
}

// -------------------------------------------------------------------------------------------------
// extend: closest hit + binning by material tag
// -------------------------------------------------------------------------------------------------
RT_DEV int wf_tag_of(const DevScene& S, int slot) {
  if (slot < 0) return WF_TAG_TERMINAL;
  const int root = ldgi2(S.slot_info + slot).x;
  const int ty = ldgi4(S.matB + root).x;
  return ty == MAT_LIGHT ? WF_TAG_TERMINAL : (ty == MAT_LAMBERT ? WF_TAG_LAMBERT : WF_TAG_OTHER);
}
RT_DEV void wf_bin(const DevScene& S, const WfBuffers& W, bool done, int id, float t, int slot) {
  if (done) {
    W.ray_o[id].w = t;
    W.ray_d[id].w = __int_as_float(slot);
  }
  const int tag = done ? wf_tag_of(S, slot) : -1;
#pragma unroll
  for (int c = 0; c < WF_TAGS; ++c) wf_push(W.q_shade[c], W.counters + WFC_N_SHADE0 + c, tag == c, id);
}

template <int KIND>
__global__ void __launch_bounds__(256, 2) k_wf_extend(const DevScene S, const RenderParams R, const WfBuffers W, int n_rays,
                                                      const int* q_in) {
  __shared__ ListSmemData sm_data;
  const ListSmem sm = stage_list<KIND>(S, sm_data);
  const ListSmem& L = sm;
  unsigned st_rays = 0;
  if (KIND != BVH_SAH) {
    // brute-force list / reference tree: one whole query per ray, grid-stride
    const int stride = gridDim.x * blockDim.x;
    for (int base_i = blockIdx.x * blockDim.x; base_i < n_rays; base_i += stride) {
      const int i = base_i + threadIdx.x;
      const bool on = i < n_rays;
      int id = 0, slot = -1;
      float t = CUDART_INF_F;
      if (on) {
        id = q_in[i];
        const float4 o = wf_view(W).ray_o[id], d = wf_view(W).ray_d[id];
        closest_hit<KIND>(S, L, Ray{mk3(o.x, o.y, o.z), mk3(d.x, d.y, d.z)}, t, slot);
        ++st_rays;
      }
      wf_bin(S, W, on, id, t, slot);
    }
  } else {
    // persistent threads with a warp-level ray queue: lanes pull ray ids on demand
    const unsigned lane = threadIdx.x & 31u;
    TravStack stack;
    int id = 0;
    int st = ST_NONE;
    bool retired = false;
    Ray ray{mk3(0, 0, 0), mk3(0, 0, 1)};
    Trav tv{-1, 0, CUDART_INF_F, -1};
    BoxPre bp{mk3(0, 0, 0), mk3(0, 0, 0)};
    for (;;) {
      const bool want = st == ST_NONE && !retired;
      const unsigned m = __ballot_sync(0xffffffffu, want);
      if (m) {
        const int leader = __ffs(m) - 1;
        int base = 0;
        if ((int)lane == leader) base = atomicAdd(W.counters + WFC_EXTEND_CURSOR, __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (want) {
          const int i = base + __popc(m & ((1u << lane) - 1u));
          if (i >= n_rays) retired = true;
          else {
            id = q_in[i];
            const float4 o = wf_view(W).ray_o[id], d = wf_view(W).ray_d[id];
            ray = Ray{mk3(o.x, o.y, o.z), mk3(d.x, d.y, d.z)};
            trav_begin(S, ray, tv);
            bp = box_precompute(ray);
            ++st_rays;
            st = tv.cur >= 0 ? ST_TRACE : ST_HIT;
          }
        }
      }
      // finished rays: record the hit, bin by material tag
      wf_bin(S, W, st == ST_HIT, id, tv.tbest, tv.sbest);
      if (st == ST_HIT) st = ST_NONE;
      if (__all_sync(0xffffffffu, st == ST_NONE && retired)) break;
      // traversal bursts (see k_render_trav) until few lanes are left
      for (;;) {
        const unsigned tt = __ballot_sync(0xffffffffu, st == ST_TRACE);
        if (tt == 0) break;
        if (__popc(tt) <= R.trav_min_lanes) {
          if (__any_sync(0xffffffffu, st == ST_HIT || (st == ST_NONE && !retired))) break;
        }
        if (st == ST_TRACE) {
          int steps = 0;
          TravLeaves lv{0, 0, 0, 0};
#pragma unroll 1
          while (tv.cur >= 0 && lv.a == 0 && steps < R.trav_burst) { trav_inner(S, bp, tv, stack, lv); ++steps; }
          if (lv.a != 0) trav_leaves(S, ray, bp, tv, lv);
          st = tv.cur >= 0 ? ST_TRACE : ST_HIT;
        }
      }
    }
  }
  flush_stats(R, 0u, S.cam.samples, 0u, 0u, st_rays, 0x7fffffff, 0);
}

// -------------------------------------------------------------------------------------------------
// shade: one launch per material tag
// -------------------------------------------------------------------------------------------------
template <int KIND, int TAG>
__global__ void __launch_bounds__(256) k_wf_shade(const DevScene S, const RenderParams R, const WfBuffers W, int n, const int* q_in) {
  __shared__ ListSmemData sm_data;
  const ListSmem sm = stage_list<KIND>(S, sm_data);
  const DevCamera& cam = S.cam;
  const MixW mw = make_mixw(S);
  unsigned st_paths = 0, st_bounces = 0;
  int st_bmin = 0x7fffffff, st_bmax = 0;
  const int stride = gridDim.x * blockDim.x;
  for (int base_i = blockIdx.x * blockDim.x; base_i < n; base_i += stride) {
    const int i = base_i + threadIdx.x;
    const bool on = i < n;
    bool queued = false, freed = false;
    int id = 0;
    if (on) {
      id = q_in[i];
      const float4 o = wf_view(W).ray_o[id], d = wf_view(W).ray_d[id], tpv = wf_view(W).tp[id], rv = wf_view(W).rad[id];
      const uint2 px = wf_view(W).pix[id];
      PathState ps{Ray{mk3(o.x, o.y, o.z), mk3(d.x, d.y, d.z)}, mk3(tpv.x, tpv.y, tpv.z), mk3(rv.x, rv.y, rv.z), __float_as_int(tpv.w)};
      const float t = o.w;
      const int slot = __float_as_int(d.w);
      // resume the bounce's Philox stream after the draws taken before the trace (camera, roulette)
      Rng g;
      g.begin(px.x, px.y, (uint32_t)ps.bounces, S.seed_lo, S.seed_hi);
      for (int k = __float_as_int(rv.w); k > 0; --k) g.next();
      bool done = path_post<KIND>(S, &sm, mw, ps, g, t, slot);
      int used = 0;
      if (!done) { // the depth / roulette test of the next bounce, on its own stream
        g.begin(px.x, px.y, (uint32_t)ps.bounces, S.seed_lo, S.seed_hi);
        done = path_pre(cam, ps, g);
        used = (int)((g.block - 1u) * 5u + (5u - (unsigned)g.avail));
      }
      if (done) {
        wf_finish(R, cam, px, ps, st_paths, st_bounces, st_bmin, st_bmax);
        freed = true;
      } else {
        wf_store_path(W, id, ps, used);
        queued = true;
      }
    }
    wf_push(W.q_extend, W.counters + WFC_N_EXTEND, queued, id);
    wf_push(W.q_free, W.counters + WFC_N_FREE, freed, id);
  }
  flush_stats(R, 0u, cam.samples, st_paths, st_bounces, 0u, st_bmin, st_bmax);
}

// -------------------------------------------------------------------------------------------------
// resolve: accum -> pixels (finalColor + writeColorToBuffer), pixel stats
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wf_resolve(const DevScene S, const RenderParams R) {
  const DevCamera& cam = S.cam;
  const int w = R.x1 - R.x0, h = R.y1 - R.y0;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  bool on = idx < w * h;
  if (on) {
    const int i = R.x0 + idx % w, j = R.y0 + idx / w;
    on = R.part_count <= 1 || block_owner(i >> 3, j >> 2, (cam.width + 7) >> 3, R.part_count) == R.part_index;
    if (on) {
      const size_t pi = (size_t)j * cam.width + i;
      const unsigned long long* a = R.accum + pi * 4;
      write_pixel(R, pi, mk3(from_fixed(a[0], cam.samples), from_fixed(a[1], cam.samples), from_fixed(a[2], cam.samples)));
    }
  }
  flush_stats(R, on ? 1u : 0u, cam.samples, 0u, 0u, 0u, 0x7fffffff, 0);
}

__global__ void k_wf_fill_free(int* q, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) q[i] = i;
}

// -------------------------------------------------------------------------------------------------
// host driver
// -------------------------------------------------------------------------------------------------
template <int KIND>
static cudaError_t wf_run(const DevScene& S, const RenderParams& R, WfHost& H, int sms, cudaStream_t st, int* launches) {
  WfBuffers W = H.W;
  const WfShare sh{H.owned_blocks, H.n_owned, H.blocks_x};
  const int n_slots = W.n_slots;
  int cur = 0;
  int n_free = n_slots, n_extend = 0;
  cudaError_t e;
  k_wf_fill_free<<<(n_slots + 255) / 256, 256, 0, st>>>(H.q_free[0], n_slots);
  if ((e = cudaMemsetAsync(W.counters, 0, WFC_COUNT * sizeof(int), st)) != cudaSuccess) return e;
  *launches += 1;
  const int grid_cap = sms * 8;
  auto grid_for = [&](int n) { int g = (n + 255) / 256; return g < 1 ? 1 : (g > grid_cap ? grid_cap : g); };
  for (int iter = 0; iter < (1 << 24); ++iter) {
    // generate into (q_extend[cur], q_free[cur^1]); extend reads q_extend[cur]; shade appends to q_extend[cur^1] / q_free[cur^1]
    W.q_extend = H.q_extend[cur];
    W.q_free = H.q_free[cur ^ 1];
    if (n_free > 0) {
      k_wf_generate<<<grid_for(n_free), 256, 0, st>>>(S, R, W, sh, n_free, H.q_free[cur]);
      *launches += 1;
    }
    // the number of queued rays is only known on the device: read it back (one small sync per iteration)
    if ((e = cudaMemcpyAsync(H.h_counters, W.counters, WFC_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    n_extend = H.h_counters[WFC_N_EXTEND];
    const int n_free_after_gen = H.h_counters[WFC_N_FREE];
    if (n_extend == 0) {
      // nothing in flight: done when the pair cursor has passed the end, else only region-clipped pairs were drawn
      const unsigned long long next_pair = *reinterpret_cast<unsigned long long*>(H.h_counters + WFC_NEXT_PAIR);
      if (next_pair >= W.total_pairs || n_free_after_gen == 0) break;
      // re-generate with the slots that drew clipped pairs
      if ((e = cudaMemsetAsync(W.counters + WFC_N_FREE, 0, sizeof(int), st)) != cudaSuccess) return e;
      n_free = n_free_after_gen;
      cur ^= 1;
      // q_free[cur] now holds them; q_extend[cur] is empty
      continue;
    }
    // reset the counters this iteration's extend/shade will fill
    if ((e = cudaMemsetAsync(W.counters + WFC_N_EXTEND, 0, (WFC_N_FREE - WFC_N_EXTEND) * sizeof(int), st)) != cudaSuccess) return e;
    {
      int g = KIND == BVH_SAH ? sms * 2 : grid_for(n_extend);
      k_wf_extend<KIND><<<g, 256, 0, st>>>(S, R, W, n_extend, H.q_extend[cur]);
      *launches += 1;
    }
    if ((e = cudaMemcpyAsync(H.h_counters, W.counters, WFC_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    const int n0 = H.h_counters[WFC_N_SHADE0], n1 = H.h_counters[WFC_N_SHADE1], n2 = H.h_counters[WFC_N_SHADE2];
    // shade appends surviving rays to the OTHER extend queue; freed slots join the ones generate left over
    W.q_extend = H.q_extend[cur ^ 1];
    if (n0) { k_wf_shade<KIND, WF_TAG_TERMINAL><<<grid_for(n0), 256, 0, st>>>(S, R, W, n0, W.q_shade[0]); *launches += 1; }
    if (n1) { k_wf_shade<KIND, WF_TAG_LAMBERT><<<grid_for(n1), 256, 0, st>>>(S, R, W, n1, W.q_shade[1]); *launches += 1; }
    if (n2) { k_wf_shade<KIND, WF_TAG_OTHER><<<grid_for(n2), 256, 0, st>>>(S, R, W, n2, W.q_shade[2]); *launches += 1; }
    if ((e = cudaMemcpyAsync(H.h_counters, W.counters, WFC_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    n_free = H.h_counters[WFC_N_FREE];
    if ((e = cudaMemsetAsync(W.counters + WFC_N_FREE, 0, sizeof(int), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(W.counters + WFC_N_SHADE0, 0, WF_TAGS * sizeof(int), st)) != cudaSuccess) return e;
    cur ^= 1;
  }
  const int npx = (R.x1 - R.x0) * (R.y1 - R.y0);
  k_wf_resolve<<<(npx + 255) / 256, 256, 0, st>>>(S, R);
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_render_wavefront(const DevScene& S, const RenderParams& R, WfHost& H, int sms, cudaStream_t st, int* launches) {
  if (R.x1 <= R.x0 || R.y1 <= R.y0) return cudaSuccess;
  switch (S.bvh_kind) {
    case BVH_LIST: return wf_run<BVH_LIST>(S, R, H, sms, st, launches);
    case BVH_SAH: return wf_run<BVH_SAH>(S, R, H, sms, st, launches);
    default: return wf_run<BVH_REFERENCE>(S, R, H, sms, st, launches);
  }
}

} // namespace rt
