/* integration/accelerator_b200.cc -- see accelerator_b200.h.  Compiled inside a libYafaRay tree and linked with
 * libb200rt (include/b200rt.h).  Follows the conventions of the reference's accelerators
 * (src/accelerator/accelerator_kdtree_original.cc:28-63 for params/factory; error handling per SURVEY.md 8b:
 * no exception leaves this class, a CUDA failure is logged and the factory returns no accelerator, so that
 * SurfaceIntegrator::preprocess fails instead of silently rendering on a CPU path). */
#include "accelerator/accelerator_b200.h"
#include "b200rt.h"
#include "common/logger.h"
#include "geometry/object/object_mesh.h"
#include "geometry/primitive/primitive_face.h"
#include "geometry/primitive/primitive_instance.h"
#include "geometry/primitive/primitive_sphere.h"
#include "geometry/shape/shape_polygon.h"
#include <cstdio>
#include <cstdlib>
#include "geometry/instance.h"
#include "geometry/matrix.h"
#include "material/material.h"
#include "param/param.h"
#include "render/render_control.h"
#include <limits>
#include <memory>

namespace yafaray {

std::map<std::string, const ParamMeta *> AcceleratorB200::Params::getParamMetaMap()
{
	auto param_meta_map{ParentClassType_t::Params::getParamMetaMap()};
	PARAM_META(max_depth_);
	PARAM_META(max_leaf_size_);
	PARAM_META(cost_ratio_);
	PARAM_META(empty_bonus_);
	PARAM_META(device_);
	PARAM_META(num_threads_);
	PARAM_META(wavefront_fibers_);
	PARAM_META(wavefront_groups_);
	PARAM_META(wavefront_block_);
	PARAM_META(wavefront_stack_kb_);
	PARAM_META(tile_shard_index_);
	PARAM_META(tile_shard_count_);
	return param_meta_map;
}

AcceleratorB200::Params::Params(ParamResult &param_result, const ParamMap &param_map)
{
	PARAM_LOAD(max_depth_);
	PARAM_LOAD(max_leaf_size_);
	PARAM_LOAD(cost_ratio_);
	PARAM_LOAD(empty_bonus_);
	PARAM_LOAD(device_);
	PARAM_LOAD(num_threads_);
	PARAM_LOAD(wavefront_fibers_);
	PARAM_LOAD(wavefront_groups_);
	PARAM_LOAD(wavefront_block_);
	PARAM_LOAD(wavefront_stack_kb_);
	PARAM_LOAD(tile_shard_index_);
	PARAM_LOAD(tile_shard_count_);
}

ParamMap AcceleratorB200::getAsParamMap(bool only_non_default) const
{
	auto param_map{ParentClassType_t::getAsParamMap(only_non_default)};
	param_map.setParam("type", type().print());
	PARAM_SAVE(max_depth_);
	PARAM_SAVE(max_leaf_size_);
	PARAM_SAVE(cost_ratio_);
	PARAM_SAVE(empty_bonus_);
	PARAM_SAVE(device_);
	PARAM_SAVE(num_threads_);
	PARAM_SAVE(wavefront_fibers_);
	PARAM_SAVE(wavefront_groups_);
	PARAM_SAVE(wavefront_block_);
	PARAM_SAVE(wavefront_stack_kb_);
	PARAM_SAVE(tile_shard_index_);
	PARAM_SAVE(tile_shard_count_);
	return param_map;
}

std::pair<std::unique_ptr<Accelerator>, ParamResult> AcceleratorB200::factory(Logger &logger, const RenderControl *render_control, const std::vector<const Primitive *> &primitives, const ParamMap &param_map)
{
	auto param_result{class_meta::check<Params>(param_map, {"type"}, {})};
	auto accelerator{std::make_unique<AcceleratorB200>(logger, param_result, render_control, primitives, param_map)};
	if(param_result.notOk()) logger.logWarning(param_result.print<ThisClassType_t>("", {"type"}));
	if(!accelerator->ok())
	{
		// Scene::preprocess dereferences the returned accelerator unconditionally (src/scene/scene.cc:352-353) and the C API
		// ignores the result of the preprocess calls, so returning nullptr would crash the host application.  The failed
		// accelerator is returned instead: it has a zero bound, every query misses, the error has been logged at ERROR level
		// and the result carries YAFARAY_RESULT_ERROR_WHILE_CREATING.  No CPU traversal path exists in this class.
		param_result.flags_ |= ResultFlags{YAFARAY_RESULT_ERROR_WHILE_CREATING};
		logger.logError(getClassName(), ": no usable accelerator was built; every ray will miss");
	}
	return {std::move(accelerator), param_result};
}

namespace {

// per-face flag byte: the visibility tests of accelerator.h:126-127,138-139,152-154 and Material::isTransparent()
uint8_t faceFlags(const Primitive *primitive)
{
	const Visibility prim_visibility{primitive->getVisibility()};
	const Material *material{primitive->getMaterial()};
	const Visibility mat_visibility{material ? material->getVisibility() : Visibility{Visibility::Normal}};
	uint8_t flags = 0;
	if(prim_visibility.has(Visibility::Visible) && mat_visibility.has(Visibility::Visible)) flags |= B200RT_FACE_VISIBLE;
	if(prim_visibility.has(Visibility::CastsShadows) && mat_visibility.has(Visibility::CastsShadows)) flags |= B200RT_FACE_CASTS_SHADOWS;
	if(material && material->isTransparent()) flags |= B200RT_FACE_TRANSPARENT;
	return flags;
}

inline b200rt_ray toRay(const Ray &ray, float t_max)
{
	return {ray.from_[Axis::X], ray.from_[Axis::Y], ray.from_[Axis::Z], ray.tmin_, ray.dir_[Axis::X], ray.dir_[Axis::Y], ray.dir_[Axis::Z], t_max};
}

} //namespace

AcceleratorB200::AcceleratorB200(Logger &logger, ParamResult &param_result, const RenderControl *render_control, const std::vector<const Primitive *> &primitives, const ParamMap &param_map) : ParentClassType_t{logger, param_result, render_control, param_map}, params_{param_result, param_map}, primitives_{primitives}
{
	if(logger.isDebug()) logger.logDebug("**" + getClassName() + " params_:\n" + getAsParamMap(true).print());
	logger_.logInfo(getClassName(), ": Starting build (", primitives_.size(), " prims) on CUDA device ", params_.device_);
	// flatten the primitives: only static triangle / quad faces live in the GPU tree (SURVEY.md 8f N3 lists the rest)
	std::vector<float> xyz;
	std::vector<uint32_t> idx;
	std::vector<uint8_t> flags;
	xyz.reserve(primitives_.size() * 12);
	idx.reserve(primitives_.size() * 4);
	flags.reserve(primitives_.size());
	b200rt_build_params build_params{};
	build_params.max_depth = params_.max_depth_;
	build_params.max_leaf_size = params_.max_leaf_size_;
	build_params.cost_ratio = params_.cost_ratio_;
	build_params.empty_bonus = params_.empty_bonus_;
	build_params.build_threads = params_.num_threads_;
	b200rt_scene *scene = nullptr;
	int rc = b200rt_create(params_.device_, &build_params, &scene);
	// faces gathered so far go to the library before a sphere does, so that face ids follow the primitive order
	// diagnostic (B200_DUMP_SCENE=<file>): the flattened geometry as libb200rt receives it -- per add_mesh call one record
	// {uint64 n_verts, uint64 n_faces, float xyz[3 n_verts], uint32 idx[4 n_faces], uint8 flags[n_faces]}; read by tests/tools/dump_compare.py
	const std::unique_ptr<std::FILE, int (*)(std::FILE *)> dump_file{std::getenv("B200_DUMP_SCENE") ? std::fopen(std::getenv("B200_DUMP_SCENE"), "wb") : nullptr, &std::fclose};
	std::FILE *const dump{dump_file.get()};
	const auto flush_mesh{[&]() {
		if(dump && !flags.empty())
		{
			const uint64_t counts[2]{xyz.size() / 3, flags.size()};
			std::fwrite(counts, sizeof(uint64_t), 2, dump);
			std::fwrite(xyz.data(), sizeof(float), xyz.size(), dump);
			std::fwrite(idx.data(), sizeof(uint32_t), idx.size(), dump);
			std::fwrite(flags.data(), 1, flags.size(), dump);
		}
		if(rc == B200RT_OK && !flags.empty()) rc = b200rt_add_mesh(scene, xyz.data(), xyz.size() / 3, idx.data(), idx.size() / 4, flags.data());
		xyz.clear();
		idx.clear();
		flags.clear();
	}};
	// faces of one motion-blur mesh / moving instance gathered so far (b200rt_add_mesh_bezier / _moving)
	std::vector<float> mxyz[3];
	std::vector<uint32_t> midx;
	std::vector<uint8_t> mflags;
	const void *motion_key{nullptr}, *motion_transform{nullptr};
	bool motion_bezier{false};
	float motion_t0{0.f}, motion_t1{0.f}, motion_matrices[48];
	const auto instanceKey{[](const Primitive *p) -> const void * {
		const auto *instance_primitive{dynamic_cast<const PrimitiveInstance *>(p)};
		return instance_primitive ? static_cast<const void *>(&instance_primitive->getBaseInstance()) : nullptr;
	}};
	const auto flush_motion{[&]() {
		if(rc == B200RT_OK && !mflags.empty())
		{
			if(motion_bezier) rc = b200rt_add_mesh_bezier(scene, mxyz[0].data(), mxyz[1].data(), mxyz[2].data(), mxyz[0].size() / 3, midx.data(), midx.size() / 4, mflags.data(), motion_t0, motion_t1);
			else rc = b200rt_add_mesh_moving(scene, mxyz[0].data(), mxyz[0].size() / 3, midx.data(), midx.size() / 4, mflags.data(), motion_matrices, motion_t0, motion_t1);
		}
		for(auto &v : mxyz) v.clear();
		midx.clear();
		mflags.clear();
		motion_key = nullptr;
	}};
	const bool verify_extraction{std::getenv("B200_VERIFY_EXTRACTION") != nullptr};
	size_t verify_tests{0}, verify_hits{0}, verify_mismatches{0};
	for(const Primitive *primitive : primitives_)
	{
		if(render_control_ && render_control_->canceled())
		{
			if(scene) b200rt_destroy(scene);
			return;
		}
		// Instances: PrimitiveInstance::intersect hands the base primitive the instance's matrix, nested
		// instances multiply theirs in front (include/geometry/primitive/primitive_instance.h:83-91), and a mesh face then tests
		// the ray against obj_to_world * vertex (include/geometry/primitive/primitive_face.h:81-84).  A static instance
		// (Instance::hasMotionBlur() false, include/geometry/instance.h:48) always uses matrix 0, so its faces are uploaded
		// pre-transformed WITH THE REFERENCE'S OWN matrix product and matrix * point code: same floats, same hits.
		const Primitive *base{primitive};
		Matrix4f obj_to_world{1.f};
		bool transformed{false}, moving{false};
		int levels{0};
		const Instance *moving_instance{nullptr};
		while(const auto *instance_primitive{dynamic_cast<const PrimitiveInstance *>(base)})
		{
			const Instance &instance{instance_primitive->getBaseInstance()};
			if(instance.hasMotionBlur()) { moving = true; moving_instance = &instance; }
			obj_to_world = transformed ? obj_to_world * instance.getObjToWorldMatrix(0) : instance.getObjToWorldMatrix(0);
			transformed = true;
			++levels;
			base = &instance_primitive->getBasePrimitive();
		}
		if(const auto *sphere{dynamic_cast<const SpherePrimitive *>(base)})
		{
			// Spheres (src/geometry/primitive/primitive_sphere.cc:71-102): centre and radius are the primitive's own parameters;
			// the instance matrix is ignored for them, as the reference does (primitive_sphere.cc:77-81,104-122).
			Vec3f center{0.f};
			float radius{1.f};
			const ParamMap sphere_params{sphere->getAsParamMap(false)};
			sphere_params.getParam("center", center);
			sphere_params.getParam("radius", radius);
			flush_motion();
			flush_mesh();
			if(rc == B200RT_OK)
			{
				const float center_radius[4]{center[Axis::X], center[Axis::Y], center[Axis::Z], radius};
				const uint8_t sphere_flags{faceFlags(primitive)};
				face_flags_.push_back(sphere_flags);
				rc = b200rt_add_spheres(scene, center_radius, 1, &sphere_flags);
			}
			continue;
		}
		const auto *face{dynamic_cast<const FacePrimitive *>(base)};
		const int n_vertices{face ? face->numVertices() : 0};
		// Motion blur: a face of a Bezier motion-blur mesh (directly or under static instances) and a static face
		// directly under ONE moving instance go to the GPU with their three time steps / matrices (b200rt_add_mesh_bezier / _moving).
		// Not carried: moving instances of motion-blur meshes and nested instances with a moving level (the reference multiplies
		// per-ray matrices there, primitive_instance.h:88-91).
		const bool bezier{face && face->hasMotionBlur() && !moving};
		const bool moving_face{face && moving && !face->hasMotionBlur() && levels == 1};
		if(!face || (moving && !moving_face) || (n_vertices != 3 && n_vertices != 4))
		{
			logger_.logError(getClassName(), ": primitive kind not supported by the b200-kdtree accelerator (triangle and quad mesh faces -- static, Bezier motion blur, or under one moving instance --, spheres and static instances of them are); no accelerator created");
			if(scene) b200rt_destroy(scene);
			return;
		}
		if(bezier || moving_face)
		{
			// consecutive faces of one motion-blur mesh / one moving instance travel in one call (one matrix table entry per instance)
			const void *key{bezier ? reinterpret_cast<const void *>(face->getObjectHandle()) : static_cast<const void *>(moving_instance)};
			if(motion_key != key || motion_transform != (transformed ? instanceKey(primitive) : nullptr)) flush_motion();
			flush_mesh();
			motion_key = key;
			motion_transform = transformed ? instanceKey(primitive) : nullptr;
			motion_bezier = bezier;
			const uint8_t motion_flags{faceFlags(primitive)};
			face_flags_.push_back(motion_flags);
			mflags.push_back(motion_flags);
			const uint32_t first{static_cast<uint32_t>(mxyz[0].size() / 3)};
			for(int v = 0; v < 4; ++v) midx.push_back(v < n_vertices ? first + static_cast<uint32_t>(v) : 0xFFFFFFFFu);
			if(bezier)
			{
				// the mesh's three time steps as FacePrimitive::getVertex returns them (time step 1 already holds the Bezier control
				// points, object_mesh.cc:268-277); FacePrimitive::getObjectHandle() is the address of the face's MeshObject
				const auto *mesh{reinterpret_cast<const MeshObject *>(face->getObjectHandle())};
				motion_t0 = mesh->getTimeRangeStart(); motion_t1 = mesh->getTimeRangeEnd();
				for(unsigned char step = 0; step < 3; ++step)
					for(int v = 0; v < n_vertices; ++v)
					{
						const Point3f p{transformed ? face->getVertex(v, step, obj_to_world) : face->getVertex(v, step)};
						mxyz[step].push_back(p[Axis::X]); mxyz[step].push_back(p[Axis::Y]); mxyz[step].push_back(p[Axis::Z]);
					}
			}
			else
			{
				motion_t0 = moving_instance->getTimeRangeStart(); motion_t1 = moving_instance->getTimeRangeEnd();
				for(int v = 0; v < n_vertices; ++v)
				{
					const Point3f p{face->getVertex(v, 0)};
					mxyz[0].push_back(p[Axis::X]); mxyz[0].push_back(p[Axis::Y]); mxyz[0].push_back(p[Axis::Z]);
				}
				for(unsigned char step = 0; step < 3; ++step)
				{
					const Matrix4f &m{moving_instance->getObjToWorldMatrix(step)};
					for(int i = 0; i < 4; ++i)
						for(int j = 0; j < 4; ++j) motion_matrices[16 * step + 4 * i + j] = m[i][j];
				}
			}
			continue;
		}
		flush_motion();
		const uint32_t first_vertex{static_cast<uint32_t>(xyz.size() / 3)};
		for(int v = 0; v < n_vertices; ++v)
		{
			const Point3f p{transformed ? face->getVertex(v, 0, obj_to_world) : face->getVertex(v, 0)};
			xyz.push_back(p[Axis::X]); xyz.push_back(p[Axis::Y]); xyz.push_back(p[Axis::Z]);
		}
		for(int v = 0; v < 4; ++v) idx.push_back(v < n_vertices ? first_vertex + static_cast<uint32_t>(v) : 0xFFFFFFFFu);
		flags.push_back(faceFlags(primitive));
		face_flags_.push_back(flags.back());
		if(verify_extraction)
		{
			// diagnostic (B200_VERIFY_EXTRACTION=1): the vertices just extracted, tested with the reference's own polygon code, must
			// give bit for bit what the primitive's virtual intersect() gives -- for a ray through the face and for one that grazes it
			const float *p{xyz.data() + 3 * size_t(first_vertex)};
			const Point3f v_0{{p[0], p[1], p[2]}}, v_1{{p[3], p[4], p[5]}}, v_2{{p[6], p[7], p[8]}};
			const Vec3f normal{(v_1 - v_0) ^ (v_2 - v_0)};
			const Point3f centre{(v_0 + v_1 + v_2) * (1.f / 3.f)};
			for(int k = 0; k < 2; ++k)
			{
				const Point3f from{centre + normal * (k == 0 ? 1.f : 0.37f) + (v_1 - v_0) * (k == 0 ? 0.f : 0.21f)};
				const Vec3f dir{k == 0 ? -normal : -(normal + (v_2 - v_0) * 0.4f)};
				const auto expected{primitive->intersect(from, dir, 0.f)};
				std::pair<float, Uv<float>> got;
				if(n_vertices == 3) got = ShapePolygon<float, 3>{{v_0, v_1, v_2}}.intersect(from, dir);
				else got = ShapePolygon<float, 4>{{v_0, v_1, v_2, Point3f{{p[9], p[10], p[11]}}}}.intersect(from, dir);
				++verify_tests;
				if(got.first != expected.first || got.second.u_ != expected.second.u_ || got.second.v_ != expected.second.v_) ++verify_mismatches;
				if(expected.first > 0.f) ++verify_hits;
			}
		}
	}
	if(verify_extraction) logger_.logInfo(getClassName(), ": extraction check: ", verify_tests, " test rays, ", verify_hits, " hits, ", verify_mismatches, " differ from Primitive::intersect");
	flush_motion();
	flush_mesh();
	if(dump) std::fflush(dump);
	if(rc == B200RT_OK) rc = b200rt_build(scene);
	float bound[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
	if(rc == B200RT_OK) rc = b200rt_get_bound(scene, bound);
	if(rc != B200RT_OK)
	{
		logger_.logError(getClassName(), ": libb200rt failed (", rc, "): ", b200rt_last_error());
		if(scene) b200rt_destroy(scene);
		return;
	}
	scene_ = scene;
	bound_ = Bound<float>{{{bound[0], bound[1], bound[2]}}, {{bound[3], bound[4], bound[5]}}};
	b200rt_stats stats{};
	if(b200rt_get_stats(scene_, &stats) == B200RT_OK && logger_.isVerbose())
	{
		logger_.logVerbose(getClassName(), ": Stats (build ", stats.build_seconds, "s, upload ", stats.upload_seconds, "s)");
		logger_.logVerbose(getClassName(), ": Primitives in tree: ", stats.n_faces, " (", stats.n_triangles, " triangles, ", stats.n_quads, " quads)");
		logger_.logVerbose(getClassName(), ": Interior nodes: ", stats.n_interior, " / leaf nodes: ", stats.n_leaves, " (empty: ", stats.n_empty_leaves, ")");
		logger_.logVerbose(getClassName(), ": Leaf prims: ", stats.n_leaf_refs, ", depth: ", stats.max_depth, ", device bytes: ", stats.device_bytes);
	}
}

AcceleratorB200::~AcceleratorB200()
{
	idle_queues_.clear(); //pinned buffers go before the scene (and its CUDA context use) does
	if(scene_) b200rt_destroy(scene_);
}

// ---- the three virtual per-ray queries: batches of one, rays already wrapped by the caller ------------------
IntersectData AcceleratorB200::intersect(const Ray &ray, float t_max) const
{
	const b200rt_ray r{toRay(ray, t_max)};
	b200rt_hit hit{};
	IntersectData data;
	data.t_max_ = t_max;
	if(!scene_) return data;
	if(b200::RayQueue *queue{b200::RayQueue::current()}) hit = queue->closest(scene_, r, ray.time_); //on a fiber: joins the thread's next batch
	else if(++wf_per_ray_calls_, b200rt_trace_timed(scene_, B200RT_QUERY_CLOSEST, B200RT_RAYS_TREE_SPACE, &r, &ray.time_, 1, &hit, 0) != B200RT_OK) return data;
	if(hit.prim == B200RT_MISS) return data;
	data.t_hit_ = hit.t;
	data.t_max_ = hit.t;
	data.uv_ = {hit.u, hit.v};
	data.primitive_ = primitives_[hit.prim];
	return data;
}

IntersectData AcceleratorB200::intersectShadow(const Ray &ray, float t_max) const
{
	const b200rt_ray r{toRay(ray, t_max)};
	uint32_t occluder = B200RT_MISS;
	IntersectData data;
	if(!scene_) return data;
	if(b200::RayQueue *queue{b200::RayQueue::current()}) occluder = queue->shadow(scene_, r, ray.time_);
	else if(++wf_per_ray_calls_, b200rt_trace_timed(scene_, B200RT_QUERY_SHADOW, B200RT_RAYS_TREE_SPACE, &r, &ray.time_, 1, &occluder, 0) != B200RT_OK) return data;
	if(occluder == B200RT_MISS) return data;
	data.t_hit_ = 1.f; //any value > 0: callers only use isHit() and primitive_ (accelerator.h:110)
	data.primitive_ = primitives_[occluder];
	return data;
}

IntersectData AcceleratorB200::intersectTransparentShadow(const Ray &ray, int max_depth, float dist, const Camera *camera) const
{
	const b200rt_ray r{toRay(ray, dist)};
	b200rt_tshadow res{};
	IntersectData data;
	if(max_depth < 0) max_depth = 0; //behaves like 0 in the reference: "depth >= max_depth" holds for the first transparent caster (accelerator.h:161)
	if(!scene_) return data;
	const b200rt_hit *casters{res.transparent};
	std::vector<uint32_t> deep; //a shadow_depth above the 8 casters of b200rt_tshadow: a larger record, traced on its own
	if(max_depth > B200RT_TSHADOW_MAX)
	{
		if(!depth_clamp_logged_.exchange(true)) logger_.logInfo(getClassName(), ": shadow_depth ", max_depth, " exceeds the ", B200RT_TSHADOW_MAX, " casters of a queued result record; such rays are traced one call each (b200rt_trace_tshadow_deep)");
		deep.resize(4u + 4u * static_cast<size_t>(max_depth));
		++wf_per_ray_calls_;
		if(b200rt_trace_tshadow_deep(scene_, B200RT_RAYS_TREE_SPACE, &r, &ray.time_, 1, max_depth, max_depth, deep.data()) != B200RT_OK) return data;
		res.shadowed = deep[0]; res.n_transparent = deep[1]; res.occluder = deep[2];
		casters = reinterpret_cast<const b200rt_hit *>(deep.data() + 4);
	}
	else if(b200::RayQueue *queue{b200::RayQueue::current()}) res = queue->transparentShadow(scene_, r, ray.time_, max_depth);
	else if(++wf_per_ray_calls_, b200rt_trace_timed(scene_, B200RT_QUERY_TSHADOW, B200RT_RAYS_TREE_SPACE, &r, &ray.time_, 1, &res, max_depth) != B200RT_OK) return data;
	if(res.shadowed)
	{
		data.t_hit_ = 1.f;
		data.primitive_ = primitives_[res.occluder];
		return data;
	}
	// material evaluation stays on the host: colour = product of the transparencies of the distinct casters (accelerator.h:163-166)
	for(uint32_t k = 0; k < res.n_transparent; ++k)
	{
		const b200rt_hit &h{casters[k]};
		const Primitive *primitive{primitives_[h.prim]};
		const Point3f hit_point{ray.from_ + h.t * ray.dir_};
		const auto sp{primitive->getSurface(nullptr, hit_point, ray.time_, {h.u, h.v}, camera)};
		if(sp) data.color_ *= sp->getTransparency(ray.dir_, camera);
	}
	return data; //setNoHit(): t_hit_ = 0, primitive_ = nullptr, colour kept (accelerator_kdtree_common.h:246-250)
}

// ---- wavefront ray queues ---------------------------------------------------------------------------------
std::unique_ptr<b200::RayQueue> AcceleratorB200::acquireRayQueue() const
{
	{
		std::lock_guard<std::mutex> lock(queues_mutex_);
		if(!idle_queues_.empty())
		{
			auto queue{std::move(idle_queues_.back())};
			idle_queues_.pop_back();
			return queue;
		}
	}
	return std::make_unique<b200::RayQueue>(wavefrontFibers(), wavefrontGroups(), wavefrontStackBytes());
}

void AcceleratorB200::releaseRayQueue(std::unique_ptr<b200::RayQueue> queue) const
{
	if(!queue || !queue->ok()) return; //a failed queue is dropped
	addWavefrontStats(queue->stats());
	queue->resetStats();
	std::lock_guard<std::mutex> lock(queues_mutex_);
	idle_queues_.push_back(std::move(queue));
}

void AcceleratorB200::refreshFaceFlags() const
{
	if(!scene_ || face_flags_.size() != primitives_.size()) return;
	std::lock_guard<std::mutex> lock(queues_mutex_);
	size_t changed{0};
	for(size_t i = 0; i < primitives_.size(); ++i)
	{
		const uint8_t flags{faceFlags(primitives_[i])};
		if(flags != face_flags_[i]) { face_flags_[i] = flags; ++changed; }
	}
	if(changed == 0) return;
	if(b200rt_update_face_flags(scene_, face_flags_.data(), face_flags_.size()) != B200RT_OK) logger_.logError(getClassName(), ": b200rt_update_face_flags failed: ", b200rt_last_error());
	else logger_.logInfo(getClassName(), ": visibility / transparency of ", changed, " primitives changed since the build; flags updated on the device");
}

// ---- wavefront statistics ---------------------------------------------------------------------------------
void AcceleratorB200::addWavefrontStats(const b200::RayQueue::Stats &stats) const
{
	for(int kind = 0; kind < 3; ++kind) wf_rays_[kind] += stats.rays[kind];
	wf_batches_ += stats.batches;
	wf_calls_ += stats.calls;
	wf_switches_ += stats.switches;
	wf_trace_us_ += static_cast<uint64_t>(stats.trace_seconds * 1e6);
	wf_run_us_ += static_cast<uint64_t>(stats.run_seconds * 1e6);
}

void AcceleratorB200::logWavefrontStats() const
{
	const uint64_t closest{wf_rays_[0].exchange(0)}, shadow{wf_rays_[1].exchange(0)}, tshadow{wf_rays_[2].exchange(0)};
	const uint64_t batches{wf_batches_.exchange(0)}, calls{wf_calls_.exchange(0)}, per_ray{wf_per_ray_calls_.exchange(0)};
	const double trace_s{static_cast<double>(wf_trace_us_.exchange(0)) * 1e-6}, run_s{static_cast<double>(wf_run_us_.exchange(0)) * 1e-6};
	wf_switches_ = 0;
	// kernel launches since the last line: the flush combiner of libb200rt merges the flushes of all render threads (process-wide counter)
	const uint64_t launches_now{b200rt_launch_count()};
	const uint64_t launches{launches_now - wf_launches_seen_.exchange(launches_now)};
	logger_.logInfo(getClassName(), ": wavefront rays closest=", closest, " shadow=", shadow, " transparent-shadow=", tshadow, " in ", batches, " batches / ", calls,
					" libb200rt calls (", batches ? (closest + shadow + tshadow) / batches : 0, " rays per batch), ", trace_s, " thread-seconds inside libb200rt of ", run_s, " thread-seconds in the render workers; per-ray calls outside fibers: ", per_ray,
					"; kernel launches: ", launches, " (", launches ? (closest + shadow + tshadow) / launches : 0, " rays per launch)");
}

void AcceleratorB200::logQueueError(const std::string &what) const
{
	logger_.logError(getClassName(), ": wavefront ray queue failed (", what, "); the affected workers fall back to one-ray launches");
}

} //namespace yafaray
