/*
 * rt_napi.c — N-API shim over the C ABI of include/rt_b200.h.
 *
 * UNVERIFIED IN THIS REPOSITORY'S BUILD IMAGE: there is no node toolchain and no node_api.h here
 * (SURVEY.md §0), so this file has never been compiled.  It is the binding a maintainer of
 * df07/mcp-raytracer would add; the same entry points are exercised end to end from Python ctypes
 * (mcp_raytracer_b200/_native.py) by the test-suite.
 *
 * Exports (all synchronous; wrap renderRegion in napi_async_work to keep generateImageBuffer's
 * Promise contract, src/raytracer.ts:39):
 *   createCamera(flat: FlatScene, opts: RenderOpts) -> external handle   [rt_camera_create]
 *   renderRegion(handle, region, pixelData: Uint8ClampedArray) -> stats   [rt_camera_render_region]
 *   cameraInfo(handle) -> { imageWidth, imageHeight, ... }                [rt_camera_get_info]
 *   destroyCamera(handle)                                                 [rt_camera_destroy]
 *   deviceCount() -> number                                               [rt_device_count]
 * A non-zero rt_status becomes a thrown JS Error carrying rt_last_error(), so upstream error
 * behaviour (src/scenes/scenes.ts:137,154,178,191,195) is unchanged.
 */
#include <node_api.h>
#include <string.h>

#include "../../include/rt_b200.h"

#define CHECK(env, call) do { if ((call) != napi_ok) { napi_throw_error((env), NULL, "N-API call failed: " #call); return NULL; } } while (0)

static napi_value throw_rt(napi_env env, rt_status st) {
  (void)st;
  napi_throw_error(env, NULL, rt_last_error());
  return NULL;
}

/* typed-array field of an object -> raw pointer (the JS side keeps the arrays alive during the call) */
static void* ta_field(napi_env env, napi_value obj, const char* name, size_t* len) {
  napi_value v; napi_typedarray_type ty; void* data = NULL; napi_value ab; size_t off;
  if (napi_get_named_property(env, obj, name, &v) != napi_ok) return NULL;
  if (napi_get_typedarray_info(env, v, &ty, len, &data, &ab, &off) != napi_ok) return NULL;
  return data;
}
static double num_field(napi_env env, napi_value obj, const char* name) {
  napi_value v; double d = 0; napi_get_named_property(env, obj, name, &v); napi_get_value_double(env, v, &d); return d;
}
static void vec3_field(napi_env env, napi_value obj, const char* name, double out[3]) {
  size_t n; double* p = (double*)ta_field(env, obj, name, &n);
  if (p && n >= 3) memcpy(out, p, 3 * sizeof(double));
}

static void finalize_camera(napi_env env, void* data, void* hint) { (void)env; (void)hint; rt_camera_destroy((rt_camera*)data); }

static napi_value CreateCamera(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2];
  CHECK(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  napi_value flat = argv[0], o = argv[1], cam;
  rt_scene_desc s; memset(&s, 0, sizeof(s));
  size_t n = 0;
  s.obj_type = (const uint8_t*)ta_field(env, flat, "objType", &n); s.n_objects = (uint32_t)n;
  s.obj_pos = (const double*)ta_field(env, flat, "objPos", &n);
  s.obj_u = (const double*)ta_field(env, flat, "objU", &n);
  s.obj_v = (const double*)ta_field(env, flat, "objV", &n);
  s.obj_r = (const double*)ta_field(env, flat, "objR", &n);
  s.obj_material = (const int32_t*)ta_field(env, flat, "objMaterial", &n);
  s.obj_light = (const uint8_t*)ta_field(env, flat, "objLight", &n);
  s.mat_type = (const uint8_t*)ta_field(env, flat, "matType", &n); s.n_materials = (uint32_t)n;
  s.mat_color = (const double*)ta_field(env, flat, "matColor", &n);
  s.mat_param = (const double*)ta_field(env, flat, "matParam", &n);
  s.mat_child = (const int32_t*)ta_field(env, flat, "matChild", &n);
  CHECK(env, napi_get_named_property(env, flat, "camera", &cam));
  s.camera.vfov = num_field(env, cam, "vfov"); s.camera.aperture = num_field(env, cam, "aperture"); s.camera.focus = num_field(env, cam, "focus");
  vec3_field(env, cam, "from", s.camera.from); vec3_field(env, cam, "at", s.camera.at); vec3_field(env, cam, "up", s.camera.up);
  vec3_field(env, cam, "backgroundTop", s.camera.background_top); vec3_field(env, cam, "backgroundBottom", s.camera.background_bottom);
  rt_render_opts r; memset(&r, 0, sizeof(r));
  r.width = (int32_t)num_field(env, o, "width"); r.aspect = num_field(env, o, "aspect"); r.samples = (int32_t)num_field(env, o, "samples");
  r.depth = (int32_t)num_field(env, o, "depth"); r.a_tolerance = num_field(env, o, "aTolerance"); r.a_batch = (int32_t)num_field(env, o, "aBatch");
  r.roulette = (int32_t)num_field(env, o, "roulette"); r.roulette_depth = (int32_t)num_field(env, o, "rouletteDepth");
  r.mode = (int32_t)num_field(env, o, "mode"); r.seed = (uint64_t)num_field(env, o, "seed");
  r.bvh = RT_BVH_AUTO; r.integrator = RT_INTEGRATOR_AUTO; r.device = (int32_t)num_field(env, o, "device");
  r.part_index = (int32_t)num_field(env, o, "partIndex"); r.part_count = (int32_t)num_field(env, o, "partCount");
  rt_camera* h = NULL;
  rt_status st = rt_camera_create(&s, &r, &h);
  if (st != RT_OK) return throw_rt(env, st);
  napi_value ext;
  CHECK(env, napi_create_external(env, h, finalize_camera, NULL, &ext));
  return ext;
}

static napi_value RenderRegion(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3];
  CHECK(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  rt_camera* h = NULL; CHECK(env, napi_get_value_external(env, argv[0], (void**)&h));
  rt_region reg = { (int32_t)num_field(env, argv[1], "x"), (int32_t)num_field(env, argv[1], "y"),
                    (int32_t)num_field(env, argv[1], "width"), (int32_t)num_field(env, argv[1], "height") };
  napi_typedarray_type ty; size_t len; void* data; napi_value ab; size_t off;
  CHECK(env, napi_get_typedarray_info(env, argv[2], &ty, &len, &data, &ab, &off)); /* Uint8ClampedArray, also over a SharedArrayBuffer */
  rt_stats st; rt_status rc = rt_camera_render_region(h, &reg, (uint8_t*)data, len, NULL, &st);
  if (rc != RT_OK) return throw_rt(env, rc);
  napi_value out, v;
  CHECK(env, napi_create_object(env, &out));
#define SETD(name, val) do { napi_create_double(env, (double)(val), &v); napi_set_named_property(env, out, name, v); } while (0)
  SETD("pixels", st.pixels); SETD("samplesTotal", st.samples_total); SETD("samplesMin", st.samples_min); SETD("samplesMax", st.samples_max);
  SETD("bouncesTotal", st.bounces_total); SETD("bouncesMin", st.bounces_min); SETD("bouncesMax", st.bounces_max);
  SETD("rays", st.rays); SETD("deviceMs", st.device_ms);
  return out;
}

static napi_value CameraInfo(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1];
  CHECK(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  rt_camera* h = NULL; CHECK(env, napi_get_value_external(env, argv[0], (void**)&h));
  rt_camera_info ci; rt_status rc = rt_camera_get_info(h, &ci);
  if (rc != RT_OK) return throw_rt(env, rc);
  napi_value out, v; CHECK(env, napi_create_object(env, &out));
  SETD("imageWidth", ci.image_width); SETD("imageHeight", ci.image_height); SETD("channels", ci.channels);
  SETD("nLights", ci.n_lights); SETD("focusDistance", ci.focus_distance); SETD("useAdaptiveSampling", ci.use_adaptive_sampling);
  return out;
}

static napi_value DestroyCamera(napi_env env, napi_callback_info info) {
  /* the external's finalizer owns destruction; explicit destroy is a no-op kept for symmetry */
  (void)info; napi_value u; napi_get_undefined(env, &u); return u;
}
static napi_value DeviceCount(napi_env env, napi_callback_info info) { (void)info; napi_value v; napi_create_int32(env, rt_device_count(), &v); return v; }
/* device buffers of destroyed cameras are cached for the next createCamera; a long-lived MCP server can hand them back */
static napi_value TrimDeviceCache(napi_env env, napi_callback_info info) { (void)info; napi_value v; napi_create_double(env, (double)rt_trim_device_cache(), &v); return v; }

static napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor d[] = {
    {"createCamera", 0, CreateCamera, 0, 0, 0, napi_default, 0}, {"renderRegion", 0, RenderRegion, 0, 0, 0, napi_default, 0},
    {"cameraInfo", 0, CameraInfo, 0, 0, 0, napi_default, 0}, {"destroyCamera", 0, DestroyCamera, 0, 0, 0, napi_default, 0},
    {"deviceCount", 0, DeviceCount, 0, 0, 0, napi_default, 0}, {"trimDeviceCache", 0, TrimDeviceCache, 0, 0, 0, napi_default, 0},
  };
  napi_define_properties(env, exports, sizeof(d) / sizeof(d[0]), d);
  return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
