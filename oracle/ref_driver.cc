/* oracle/ref_driver.cc -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product path).
 *
 * A small C ABI ("yref_*") around the UNMODIFIED reference (libYafaRay, compiled in place from
 * /root/reference by oracle/Makefile).  It builds a scene through the reference's public C API
 * (include/public_api/yafaray_c_api.h:193-224), takes the accelerator the reference itself
 * constructed in Scene::preprocess (src/scene/scene.cc:314-353) and runs ray batches through
 *   - Accelerator::intersect(ray, t_max)            (include/accelerator/accelerator.h:50, wrapper rule :91)
 *   - Accelerator::isShadowed(ray)                  (include/accelerator/accelerator.h:103-111)
 *   - Accelerator::isShadowedTransparentShadow(...) (include/accelerator/accelerator.h:113-120)
 * so that tests/ and bench.py's cpu_baseline leg can pin the C restatement (kd_oracle.c) and the CUDA
 * path against the real thing, and time the real thing on the host cores.
 *
 * It also exports the reference's own kd-tree (nodes + leaf primitive lists) in a flat form so the C
 * restatement can be checked bit-for-bit, ties included, on exactly the tree the reference traverses.
 * That needs the private members of AcceleratorKdTree; this one file is compiled with
 * -fno-access-control instead of editing any reference header.
 *
 * Ray record (8 floats, same as include/b200rt.h): ox oy oz tmin dx dy dz tmax.
 */
#include "yafaray_c_api.h"
#include "scene/scene.h"
#include "accelerator/accelerator.h"
#include "accelerator/accelerator_kdtree_original.h"
#include "geometry/object/object.h"
#include "geometry/primitive/primitive.h"
#include "geometry/primitive/primitive_instance.h"
#include "geometry/instance.h"
#include "geometry/ray.h"
#include "common/items.h"

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct RefScene
{
	yafaray_Logger *logger = nullptr;
	yafaray_Scene *scene = nullptr;
	yafaray_RenderControl *render_control = nullptr;
	std::vector<size_t> material_ids;
	int n_objects = 0;
	std::vector<const yafaray::Primitive *> prims;            // same order Scene::preprocess hands to the factory
	std::unordered_map<const yafaray::Primitive *, int32_t> prim_index;
	const yafaray::Accelerator *accel = nullptr;
	double build_seconds = 0.0;
	const float *times = nullptr; // Ray::time_ of the rays of the following trace calls (yref_set_times); nullptr = 0
};

void loggerSink(yafaray_LogLevel, size_t, const char *, const char *, void *) {}

template <typename F>
void parallelFor(size_t n, int n_threads, F &&f)
{
	if(n_threads <= 0) n_threads = static_cast<int>(std::thread::hardware_concurrency());
	if(n_threads <= 1 || n < 1024) { f(size_t{0}, n); return; }
	std::atomic<size_t> next{0};
	const size_t chunk = 4096;
	std::vector<std::thread> pool;
	for(int t = 0; t < n_threads; ++t)
		pool.emplace_back([&]() {
			for(;;)
			{
				const size_t b = next.fetch_add(chunk);
				if(b >= n) break;
				f(b, std::min(n, b + chunk));
			}
		});
	for(auto &th : pool) th.join();
}

inline yafaray::Ray makeRay(const float *r, float time = 0.f)
{
	return yafaray::Ray{yafaray::Point3f{{r[0], r[1], r[2]}}, yafaray::Vec3f{{r[4], r[5], r[6]}}, time, /*tmin*/ r[3], /*tmax*/ r[7]};
}

} // namespace

extern "C" {

void *yref_scene_create(int verbose)
{
	auto *s = new RefScene;
	s->logger = yafaray_createLogger("oracle", verbose ? nullptr : loggerSink, nullptr, verbose ? YAFARAY_DISPLAY_CONSOLE_NORMAL : YAFARAY_DISPLAY_CONSOLE_HIDDEN);
	yafaray_setConsoleVerbosityLevel(s->logger, verbose ? YAFARAY_LOG_LEVEL_VERBOSE : YAFARAY_LOG_LEVEL_MUTE);
	yafaray_setLogVerbosityLevel(s->logger, YAFARAY_LOG_LEVEL_MUTE);
	s->scene = yafaray_createScene(s->logger, "oracle_scene");
	s->render_control = yafaray_createRenderControl();
	yafaray_setRenderControlForNormalStart(s->render_control);
	return s;
}

void yref_scene_destroy(void *h)
{
	auto *s = static_cast<RefScene *>(h);
	if(!s) return;
	yafaray_destroyScene(s->scene);
	yafaray_destroyRenderControl(s->render_control);
	yafaray_destroyLogger(s->logger);
	delete s;
}

/* visibility: "normal" | "invisible" | "shadow_only" | "no_shadows"; transparency in [0,1] (shinydiffusemat).
 * Returns the index to pass as face material in yref_add_mesh, or -1. */
int yref_add_material(void *h, const char *visibility, float transparency)
{
	auto *s = static_cast<RefScene *>(h);
	yafaray_ParamMap *pm = yafaray_createParamMap();
	yafaray_setParamMapString(pm, "type", "shinydiffusemat");
	yafaray_setParamMapColor(pm, "color", 0.8, 0.6, 0.4, 1.0);
	yafaray_setParamMapFloat(pm, "diffuse_reflect", 1.0);
	yafaray_setParamMapFloat(pm, "transparency", transparency);
	yafaray_setParamMapFloat(pm, "transmit_filter", 1.0);
	yafaray_setParamMapString(pm, "visibility", visibility);
	size_t id = 0;
	const std::string name = "oracle_mat_" + std::to_string(s->material_ids.size());
	yafaray_ParamMapList *nodes = yafaray_createParamMapList();
	const yafaray_ResultFlags res = yafaray_createMaterial(s->scene, &id, name.c_str(), pm, nodes);
	yafaray_destroyParamMapList(nodes);
	yafaray_destroyParamMap(pm);
	if(res & YAFARAY_RESULT_ERROR_WHILE_CREATING) return -1;
	s->material_ids.push_back(id);
	return static_cast<int>(s->material_ids.size()) - 1;
}

/* idx: 4 uint32 per face; idx[4*f+3] == 0xFFFFFFFF marks a triangle, otherwise a quad.
 * face_material: index returned by yref_add_material per face (NULL => material 0). */
int yref_add_mesh(void *h, const float *xyz, size_t n_verts, const uint32_t *idx, size_t n_faces, const int32_t *face_material, const char *object_visibility)
{
	auto *s = static_cast<RefScene *>(h);
	if(s->material_ids.empty()) return -1;
	yafaray_ParamMap *pm = yafaray_createParamMap();
	yafaray_setParamMapString(pm, "type", "mesh");
	yafaray_setParamMapInt(pm, "num_vertices", static_cast<int>(n_verts));
	yafaray_setParamMapInt(pm, "num_faces", static_cast<int>(n_faces));
	yafaray_setParamMapString(pm, "visibility", object_visibility ? object_visibility : "normal");
	size_t object_id = 0;
	const std::string name = "oracle_mesh_" + std::to_string(s->n_objects++);
	const yafaray_ResultFlags res = yafaray_createObject(s->scene, &object_id, name.c_str(), pm);
	yafaray_destroyParamMap(pm);
	if(res & YAFARAY_RESULT_ERROR_WHILE_CREATING) return -2;
	for(size_t v = 0; v < n_verts; ++v) yafaray_addVertex(s->scene, object_id, xyz[3 * v], xyz[3 * v + 1], xyz[3 * v + 2]);
	for(size_t f = 0; f < n_faces; ++f)
	{
		const size_t mat = s->material_ids[face_material ? face_material[f] : 0];
		const uint32_t *i = idx + 4 * f;
		if(i[3] == 0xFFFFFFFFu) yafaray_addTriangle(s->scene, object_id, i[0], i[1], i[2], mat);
		else yafaray_addQuad(s->scene, object_id, i[0], i[1], i[2], i[3], mat);
	}
	yafaray_initObject(s->scene, object_id, s->material_ids[0]);
	return static_cast<int>(object_id);
}

/* Mesh with motion blur and / or to be used as the base of instances.  xyz1 / xyz2 non-NULL: a Bezier motion-blur mesh
 * ("motion_blur_bezier", three time steps given with yafaray_addVertexTimeStep, include/public_api/yafaray_c_api.h:214) over
 * [time_start, time_end].  is_base != 0: "is_base_object", the object itself is not rendered (src/scene/scene.cc:324), only its
 * instances are.  Returns the object id. */
int yref_add_mesh_ex(void *h, const float *xyz, const float *xyz1, const float *xyz2, size_t n_verts, const uint32_t *idx, size_t n_faces,
                     const int32_t *face_material, const char *object_visibility, int is_base, float time_start, float time_end)
{
	auto *s = static_cast<RefScene *>(h);
	if(s->material_ids.empty()) return -1;
	const bool bezier = xyz1 && xyz2;
	yafaray_ParamMap *pm = yafaray_createParamMap();
	yafaray_setParamMapString(pm, "type", "mesh");
	yafaray_setParamMapInt(pm, "num_vertices", static_cast<int>(n_verts));
	yafaray_setParamMapInt(pm, "num_faces", static_cast<int>(n_faces));
	yafaray_setParamMapString(pm, "visibility", object_visibility ? object_visibility : "normal");
	yafaray_setParamMapBool(pm, "is_base_object", is_base ? YAFARAY_BOOL_TRUE : YAFARAY_BOOL_FALSE);
	if(bezier)
	{
		yafaray_setParamMapBool(pm, "motion_blur_bezier", YAFARAY_BOOL_TRUE);
		yafaray_setParamMapFloat(pm, "time_range_start", time_start);
		yafaray_setParamMapFloat(pm, "time_range_end", time_end);
	}
	size_t object_id = 0;
	const std::string name = "oracle_mesh_" + std::to_string(s->n_objects++);
	const yafaray_ResultFlags res = yafaray_createObject(s->scene, &object_id, name.c_str(), pm);
	yafaray_destroyParamMap(pm);
	if(res & YAFARAY_RESULT_ERROR_WHILE_CREATING) return -2;
	const float *steps[3] = {xyz, xyz1, xyz2};
	for(int step = 0; step < (bezier ? 3 : 1); ++step)
		for(size_t v = 0; v < n_verts; ++v) yafaray_addVertexTimeStep(s->scene, object_id, steps[step][3 * v], steps[step][3 * v + 1], steps[step][3 * v + 2], static_cast<unsigned char>(step));
	for(size_t f = 0; f < n_faces; ++f)
	{
		const size_t mat = s->material_ids[face_material ? face_material[f] : 0];
		const uint32_t *i = idx + 4 * f;
		if(i[3] == 0xFFFFFFFFu) yafaray_addTriangle(s->scene, object_id, i[0], i[1], i[2], mat);
		else yafaray_addQuad(s->scene, object_id, i[0], i[1], i[2], i[3], mat);
	}
	yafaray_initObject(s->scene, object_id, s->material_ids[0]);
	return static_cast<int>(object_id);
}

/* One instance of object `object_id` with n_matrices obj_to_world matrices (16 doubles each, row major) at the given times:
 * 1 matrix = static instance, 3 = moving instance (Instance::hasMotionBlur, include/geometry/instance.h:48). */
int yref_add_instance(void *h, int object_id, const double *matrices, const float *times, int n_matrices)
{
	auto *s = static_cast<RefScene *>(h);
	const size_t instance_id = yafaray_createInstance(s->scene);
	if(!yafaray_addInstanceObject(s->scene, instance_id, static_cast<size_t>(object_id))) return -1;
	for(int k = 0; k < n_matrices; ++k)
		if(!yafaray_addInstanceMatrixArray(s->scene, instance_id, matrices + 16 * k, times[k])) return -2;
	return static_cast<int>(instance_id);
}

/* One "sphere" object (Object::factory -> SpherePrimitive, src/geometry/object/object.cc:80-90,
 * src/geometry/primitive/primitive_sphere.cc).  material: index returned by yref_add_material. */
int yref_add_sphere(void *h, float cx, float cy, float cz, float radius, int material, const char *object_visibility)
{
	auto *s = static_cast<RefScene *>(h);
	if(material < 0 || static_cast<size_t>(material) >= s->material_ids.size()) return -1;
	yafaray_ParamMap *pm = yafaray_createParamMap();
	yafaray_setParamMapString(pm, "type", "sphere");
	yafaray_setParamMapVector(pm, "center", cx, cy, cz);
	yafaray_setParamMapFloat(pm, "radius", radius);
	yafaray_setParamMapString(pm, "material", ("oracle_mat_" + std::to_string(material)).c_str());
	yafaray_setParamMapString(pm, "visibility", object_visibility ? object_visibility : "normal");
	size_t object_id = 0;
	const std::string name = "oracle_sphere_" + std::to_string(s->n_objects++);
	const yafaray_ResultFlags res = yafaray_createObject(s->scene, &object_id, name.c_str(), pm);
	yafaray_destroyParamMap(pm);
	if(res & YAFARAY_RESULT_ERROR_WHILE_CREATING) return -2;
	return static_cast<int>(object_id);
}

/* accel_type NULL => no accelerator params at all (reference default, like tests/test01).
 * depth/max_leaf/cost_ratio/empty_bonus < 0 => leave at the reference default. */
int yref_build(void *h, const char *accel_type, int depth, int max_leaf_size, float cost_ratio, float empty_bonus)
{
	auto *s = static_cast<RefScene *>(h);
	if(accel_type)
	{
		yafaray_ParamMap *pm = yafaray_createParamMap();
		yafaray_setParamMapString(pm, "type", accel_type);
		if(depth >= 0) yafaray_setParamMapInt(pm, "depth", depth);
		if(max_leaf_size >= 0) yafaray_setParamMapInt(pm, "max_leaf_size_", max_leaf_size);
		if(cost_ratio >= 0.f) yafaray_setParamMapFloat(pm, "cost_ratio", cost_ratio);
		if(empty_bonus >= 0.f) yafaray_setParamMapFloat(pm, "empty_bonus", empty_bonus);
		yafaray_setSceneAcceleratorParams(s->scene, pm);
		yafaray_destroyParamMap(pm);
	}
	const auto t0 = std::chrono::steady_clock::now();
	const yafaray_SceneModifiedFlags flags = yafaray_checkAndClearSceneModifiedFlags(s->scene);
	if(!yafaray_preprocessScene(s->scene, s->render_control, flags)) return -1;
	s->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	auto *scene = reinterpret_cast<yafaray::Scene *>(s->scene);
	s->accel = scene->getAccelerator();
	if(!s->accel) return -2;
	// identical gathering order to Scene::preprocess (src/scene/scene.cc:320-341): objects, then instances
	s->prims.clear();
	s->prim_index.clear();
	for(const auto &[object, object_name, object_enabled] : scene->getObjects())
	{
		if(!object || !object_enabled || object->getVisibility() == yafaray::Visibility::None || object->isBaseObject()) continue;
		const auto object_primitives{object->getPrimitives()};
		s->prims.insert(s->prims.end(), object_primitives.begin(), object_primitives.end());
	}
	for(const auto &instance : scene->instances_)
	{
		if(!instance) continue;
		const auto instance_primitives{instance->getPrimitives()};
		s->prims.insert(s->prims.end(), instance_primitives.begin(), instance_primitives.end());
	}
	for(size_t i = 0; i < s->prims.size(); ++i) s->prim_index[s->prims[i]] = static_cast<int32_t>(i);
	return 0;
}

/* Ray::time_ (include/geometry/ray.h:49) of the rays of the following trace calls, one float per ray; NULL = time 0. */
void yref_set_times(void *h, const float *times) { static_cast<RefScene *>(h)->times = times; }

double yref_build_seconds(void *h) { return static_cast<RefScene *>(h)->build_seconds; }
size_t yref_num_prims(void *h) { return static_cast<RefScene *>(h)->prims.size(); }

void yref_get_bound(void *h, float *out6)
{
	const auto b = static_cast<RefScene *>(h)->accel->getBound();
	for(int i = 0; i < 3; ++i) { out6[i] = b.a_[static_cast<yafaray::Axis>(i)]; out6[3 + i] = b.g_[static_cast<yafaray::Axis>(i)]; }
}

/* closest hit; out_prim = -1 on miss.  Returns wall seconds of the traced region. */
double yref_trace_closest(void *h, const float *rays, size_t n, float *out_t, float *out_u, float *out_v, int32_t *out_prim, int n_threads)
{
	auto *s = static_cast<RefScene *>(h);
	const auto t0 = std::chrono::steady_clock::now();
	parallelFor(n, n_threads, [&](size_t b, size_t e) {
		for(size_t i = b; i < e; ++i)
		{
			const yafaray::Ray ray{makeRay(rays + 8 * i, s->times ? s->times[i] : 0.f)};
			const float t_max = (ray.tmax_ >= 0.f) ? ray.tmax_ : std::numeric_limits<float>::max(); // accelerator.h:91
			const yafaray::IntersectData d{s->accel->intersect(ray, t_max)};
			if(d.isHit() && d.primitive_)
			{
				out_t[i] = d.t_max_; out_u[i] = d.uv_.u_; out_v[i] = d.uv_.v_;
				out_prim[i] = s->prim_index.at(d.primitive_);
			}
			else { out_t[i] = 0.f; out_u[i] = 0.f; out_v[i] = 0.f; out_prim[i] = -1; }
		}
	});
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* any-hit shadow through Accelerator::isShadowed; out_prim = occluder index or -1. */
double yref_trace_shadow(void *h, const float *rays, size_t n, uint8_t *out_shadowed, int32_t *out_prim, int n_threads)
{
	auto *s = static_cast<RefScene *>(h);
	const auto t0 = std::chrono::steady_clock::now();
	parallelFor(n, n_threads, [&](size_t b, size_t e) {
		for(size_t i = b; i < e; ++i)
		{
			const auto [shadowed, prim] = s->accel->isShadowed(makeRay(rays + 8 * i, s->times ? s->times[i] : 0.f));
			out_shadowed[i] = shadowed ? 1 : 0;
			if(out_prim) out_prim[i] = (shadowed && prim) ? s->prim_index.at(prim) : -1;
		}
	});
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* transparent shadow through Accelerator::isShadowedTransparentShadow (camera = nullptr); out_rgb 3 floats per ray. */
double yref_trace_tshadow(void *h, const float *rays, size_t n, int max_depth, uint8_t *out_shadowed, float *out_rgb, int n_threads)
{
	auto *s = static_cast<RefScene *>(h);
	const auto t0 = std::chrono::steady_clock::now();
	parallelFor(n, n_threads, [&](size_t b, size_t e) {
		for(size_t i = b; i < e; ++i)
		{
			const auto [shadowed, color, prim] = s->accel->isShadowedTransparentShadow(makeRay(rays + 8 * i, s->times ? s->times[i] : 0.f), max_depth, nullptr);
			out_shadowed[i] = shadowed ? 1 : 0;
			out_rgb[3 * i] = color.r_; out_rgb[3 * i + 1] = color.g_; out_rgb[3 * i + 2] = color.b_;
		}
	});
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* ---- export of the reference's own tree (only for type yafaray-kdtree-original) ----
 * node i: split[i] (interior) ; flags[i] exactly as AcceleratorKdTree::Node::flags_
 * (accelerator_kdtree_original.h:106-125); leaf i: first_ref[i] indexes refs[], nPrimitives = flags>>2. */
int64_t yref_tree_counts(void *h, int64_t *n_refs)
{
	auto *s = static_cast<RefScene *>(h);
	const auto *kd = dynamic_cast<const yafaray::AcceleratorKdTree *>(s->accel);
	if(!kd) return -1;
	int64_t refs = 0;
	for(uint32_t i = 0; i < kd->next_free_node_; ++i)
		if(kd->nodes_[i].isLeaf()) refs += kd->nodes_[i].nPrimitives();
	if(n_refs) *n_refs = refs;
	return kd->next_free_node_;
}

int yref_tree_export(void *h, float *split, uint32_t *flags, uint32_t *first_ref, uint32_t *refs)
{
	auto *s = static_cast<RefScene *>(h);
	const auto *kd = dynamic_cast<const yafaray::AcceleratorKdTree *>(s->accel);
	if(!kd) return -1;
	uint32_t cursor = 0;
	for(uint32_t i = 0; i < kd->next_free_node_; ++i)
	{
		const auto &node = kd->nodes_[i];
		flags[i] = node.flags_;
		first_ref[i] = 0;
		split[i] = 0.f;
		if(!node.isLeaf()) { split[i] = node.splitPos(); continue; }
		const uint32_t np = node.nPrimitives();
		first_ref[i] = cursor;
		if(np == 1) refs[cursor++] = static_cast<uint32_t>(s->prim_index.at(node.getOnePrimitive()));
		else for(uint32_t k = 0; k < np; ++k) refs[cursor++] = static_cast<uint32_t>(s->prim_index.at(node.primitives_[k]));
	}
	return 0;
}

int yref_hardware_threads(void) { return static_cast<int>(std::thread::hardware_concurrency()); }

} // extern "C"
