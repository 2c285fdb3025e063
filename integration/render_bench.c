/* integration/render_bench.c -- a client of libYafaRay's PUBLIC C API only (yafaray_c_api.h), in the style of the
 * reference's tests/testNN clients: builds a synthetic height-field scene of the S1M-hf shape with a
 * few boxes standing on it, one point light and one area light, and renders it with the integrator and the accelerator
 * type named on the command line.  Used to time `yafaray_render` end to end with the stock CPU kd-tree against the
 * "b200-kdtree" accelerator and to compare the two images (tests/test_render.py, tools/render_compare.py).
 *
 *   render_bench <accelerator-type> <integrator> <cells> <width> <height> <aa_samples> <out.tga> [threads] [key=value ...]
 *     accelerator-type  yafaray-kdtree-original | yafaray-kdtree-multi-thread | b200-kdtree
 *     integrator        directlighting | pathtracing
 *     key=value         extra integer accelerator parameters (e.g. wavefront_fibers=0 wavefront_block=4)
 *     i:key=value       extra integer integrator parameters (e.g. i:AO_samples=32 i:bounces=5)
 *     b:key=0|1         extra boolean integrator parameters (e.g. b:do_AO=1)
 *     f:key=value       extra float integrator parameters (e.g. f:AO_distance=2.5)
 *     b:key=value       extra boolean integrator parameters (e.g. b:time_forced=1 f:time_forced_value=0.4)
 *     tile_shard=i/n    multi-GPU rendering: render only share i of n of the frame's tiles (render/tile_shard_b200.h); sets the
 *                       b200-kdtree parameters tile_shard_index / tile_shard_count, or B200_TILE_SHARD for the stock accelerators
 *     instances=n       n static instances (rotation about z + translation) of one box standing on the field; every third one
 *                       is an instance OF THE PREVIOUS INSTANCE (nested, include/geometry/primitive/primitive_instance.h:88-91)
 *     spheres=n         n objects of type "sphere" (SpherePrimitive) resting on / floating above the field
 *     rerender_hidden=1 after the render, replace the material of the boxes by an INVISIBLE one (no object is touched, so the scene keeps
 *                       its accelerator, src/scene/scene.cc:318) and render again into the same outputs: what the files hold is the
 *                       second frame -- the boxes and their shadows must be gone (AcceleratorB200::refreshFaceFlags)
 *     film_save=prefix  write the film's weighted sums as "<prefix> - node 0000.film" when the render ends (the reference's own
 *                       film_load_save_mode=save, src/render/imagefilm.cc:1099-1176); libyafaray_b200/film.py sums such films
 */
#define _GNU_SOURCE
#include "yafaray_c_api.h"
#include <dlfcn.h>
#include <math.h>
#include <signal.h>
#include <stdint.h>
#include <sys/time.h>
#include <ucontext.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static unsigned lcg_state = 12345u;
static double lcg(void)
{
	lcg_state = lcg_state * 1664525u + 1013904223u;
	return (double) (lcg_state >> 8) / 16777216.0;
}

static size_t add_box(yafaray_Scene *scene, const char *name, double x0, double y0, double z0, double x1, double y1, double z1, const char *material)
{
	size_t object_id = 0, material_id = 0;
	yafaray_ParamMap *pm = yafaray_createParamMap();
	yafaray_setParamMapInt(pm, "num_faces", 6);
	yafaray_setParamMapInt(pm, "num_vertices", 8);
	yafaray_setParamMapString(pm, "type", "mesh");
	yafaray_createObject(scene, &object_id, name, pm);
	for(int k = 0; k < 8; ++k) yafaray_addVertex(scene, object_id, (k & 4) ? x1 : x0, (k & 2) ? y1 : y0, (k & 1) ? z1 : z0);
	yafaray_getMaterialId(scene, &material_id, material);
	yafaray_addQuad(scene, object_id, 2, 0, 1, 3, material_id);
	yafaray_addQuad(scene, object_id, 3, 7, 6, 2, material_id);
	yafaray_addQuad(scene, object_id, 7, 5, 4, 6, material_id);
	yafaray_addQuad(scene, object_id, 0, 4, 5, 1, material_id);
	yafaray_addQuad(scene, object_id, 0, 2, 6, 4, material_id);
	yafaray_addQuad(scene, object_id, 5, 7, 3, 1, material_id);
	yafaray_initObject(scene, object_id, material_id);
	yafaray_destroyParamMap(pm);
	return object_id;
}

/* B200_PROF=1: a sampling profile of yafaray_render (SIGPROF on process CPU time, the interrupted program counter only -- the
 * render threads run on fibers, which unwinders do not like).  Prints the hottest (module, offset) pairs; resolve them with
 * addr2line / nm against the same binaries.  A diagnostic for "where does the CPU side of a b200-kdtree frame go". */
#define PROF_MAX (1 << 20)
static uintptr_t prof_pc[PROF_MAX];
static volatile int prof_n = 0;
static void prof_handler(int sig, siginfo_t *si, void *uc)
{
	(void) sig; (void) si;
	const int i = __atomic_fetch_add(&prof_n, 1, __ATOMIC_RELAXED);
	if(i < PROF_MAX) prof_pc[i] = (uintptr_t) ((ucontext_t *) uc)->uc_mcontext.gregs[REG_RIP];
}
static void prof_start(void)
{
	struct sigaction sa;
	memset(&sa, 0, sizeof sa);
	sa.sa_sigaction = prof_handler;
	sa.sa_flags = SA_SIGINFO | SA_RESTART;
	sigaction(SIGPROF, &sa, NULL);
	const struct itimerval it = {{0, 200}, {0, 200}};
	setitimer(ITIMER_PROF, &it, NULL);
}
static int prof_cmp(const void *a, const void *b) { const uintptr_t x = *(const uintptr_t *) a, y = *(const uintptr_t *) b; return x < y ? -1 : x > y; }
static void prof_stop(void)
{
	const struct itimerval off = {{0, 0}, {0, 0}};
	setitimer(ITIMER_PROF, &off, NULL);
	int n = prof_n < PROF_MAX ? prof_n : PROF_MAX;
	/* per function (dladdr symbol, else module + 256-byte bucket) */
	for(int i = 0; i < n; ++i)
	{
		Dl_info info;
		if(dladdr((void *) prof_pc[i], &info) && info.dli_saddr) prof_pc[i] = (uintptr_t) info.dli_saddr;
		else prof_pc[i] &= ~(uintptr_t) 255;
	}
	qsort(prof_pc, (size_t) n, sizeof(uintptr_t), prof_cmp);
	printf("PROF %d samples\n", n);
	for(int i = 0; i < n;)
	{
		int j = i;
		while(j < n && prof_pc[j] == prof_pc[i]) ++j;
		if((j - i) * 400 >= n) /* at least 0.25 % */
		{
			Dl_info info;
			memset(&info, 0, sizeof info);
			dladdr((void *) prof_pc[i], &info);
			printf("PROF %6.2f%% %s +0x%lx %s\n", 100.0 * (j - i) / n, info.dli_fname ? info.dli_fname : "?", (unsigned long) (prof_pc[i] - (uintptr_t) info.dli_fbase), info.dli_sname ? info.dli_sname : "");
		}
		i = j;
	}
}

int main(int argc, char **argv)
{
	if(argc < 8)
	{
		fprintf(stderr, "usage: %s <accelerator-type> <integrator> <cells> <width> <height> <aa_samples> <out.tga> [threads] [key=value ...]\n", argv[0]);
		return 2;
	}
	const char *accel = argv[1], *integrator = argv[2], *out_path = argv[7];
	const int cells = atoi(argv[3]), width = atoi(argv[4]), height = atoi(argv[5]), aa_samples = atoi(argv[6]);
	const int threads = argc > 8 ? atoi(argv[8]) : -1;
	const double scale = 10.0; /* scene spans [0,10]^2 */
	const char *film_save = NULL;

	yafaray_Logger *logger = yafaray_createLogger("", NULL, NULL, YAFARAY_DISPLAY_CONSOLE_NORMAL);
	yafaray_setConsoleLogColorsEnabled(logger, YAFARAY_BOOL_FALSE);
	yafaray_setConsoleVerbosityLevel(logger, YAFARAY_LOG_LEVEL_INFO);
	yafaray_Scene *scene = yafaray_createScene(logger, "scene");
	yafaray_ParamMap *pm = yafaray_createParamMap();
	yafaray_ParamMapList *pml = yafaray_createParamMapList();
	size_t material_id = 0, object_id = 0;

	/* materials */
	yafaray_setParamMapColor(pm, "color", 0.75f, 0.7f, 0.6f, 1.f);
	yafaray_setParamMapFloat(pm, "diffuse_reflect", 1.f);
	yafaray_setParamMapString(pm, "type", "shinydiffusemat");
	yafaray_createMaterial(scene, &material_id, "ground", pm, pml);
	yafaray_clearParamMap(pm);
	yafaray_setParamMapColor(pm, "color", 0.3f, 0.45f, 0.8f, 1.f);
	yafaray_setParamMapFloat(pm, "diffuse_reflect", 0.9f);
	yafaray_setParamMapFloat(pm, "specular_reflect", 0.15f);
	yafaray_setParamMapString(pm, "type", "shinydiffusemat");
	yafaray_createMaterial(scene, &material_id, "boxes", pm, pml);

	/* height-field: cells x cells quads split in two triangles, z as S1M-hf plus a little noise */
	const double t_scene = now();
	const int nv = cells + 1;
	yafaray_clearParamMap(pm);
	yafaray_setParamMapInt(pm, "num_faces", 2 * cells * cells);
	yafaray_setParamMapInt(pm, "num_vertices", nv * nv);
	yafaray_setParamMapString(pm, "type", "mesh");
	yafaray_createObject(scene, &object_id, "heightfield", pm);
	for(int j = 0; j < nv; ++j)
		for(int i = 0; i < nv; ++i)
		{
			const double x = (double) i / cells, y = (double) j / cells;
			const double z = 0.15 * sin(17.0 * x) * cos(13.0 * y) + 0.05 * sin(71.0 * x + 3.0 * y) + 0.002 * lcg();
			yafaray_addVertex(scene, object_id, scale * x, scale * y, scale * 0.35 * z);
		}
	yafaray_getMaterialId(scene, &material_id, "ground");
	for(int j = 0; j < cells; ++j)
		for(int i = 0; i < cells; ++i)
		{
			const int v00 = j * nv + i, v10 = v00 + 1, v01 = v00 + nv, v11 = v01 + 1;
			yafaray_addTriangle(scene, object_id, v00, v10, v11, material_id);
			yafaray_addTriangle(scene, object_id, v00, v11, v01, material_id);
		}
	yafaray_initObject(scene, object_id, material_id);
	/* boxes standing in the field (occluders for the shadow rays) */
	for(int k = 0; k < 12; ++k)
	{
		char name[32];
		snprintf(name, sizeof name, "box%02d", k);
		const double cx = scale * (0.15 + 0.7 * lcg()), cy = scale * (0.15 + 0.7 * lcg()), s = scale * (0.02 + 0.03 * lcg()), h = scale * (0.08 + 0.12 * lcg());
		add_box(scene, name, cx - s, cy - s, -0.6, cx + s, cy + s, h, "boxes");
	}

	/* static instances of one more box, on a ring around the centre of the field */
	int n_instances = 0;
	for(int a = 9; a < argc; ++a) if(strncmp(argv[a], "instances=", 10) == 0) n_instances = atoi(argv[a] + 10);
	if(n_instances > 0)
	{
		const double s = 0.018 * scale;
		const size_t pillar = add_box(scene, "pillar", -s, -s, -0.6, s, s, 0.16 * scale, "boxes");
		size_t previous = 0;
		for(int k = 0; k < n_instances; ++k)
		{
			const double angle = 6.283185307179586 * k / n_instances, c = cos(2.5 * angle), sn = sin(2.5 * angle);
			const size_t instance = yafaray_createInstance(scene);
			if(k % 3 == 2)
			{
				/* nested: the previous instance moved a little further out and up */
				yafaray_addInstanceOfInstance(scene, instance, previous);
				yafaray_addInstanceMatrix(scene, instance, 1., 0., 0., 0.035 * scale * cos(angle), 0., 1., 0., 0.035 * scale * sin(angle), 0., 0., 1., 0.05 * scale, 0., 0., 0., 1., 0.f);
			}
			else
			{
				const double r = scale * (0.22 + 0.12 * (k & 1));
				yafaray_addInstanceObject(scene, instance, pillar);
				yafaray_addInstanceMatrix(scene, instance, c, -sn, 0., 0.5 * scale + r * cos(angle), sn, c, 0., 0.5 * scale + r * sin(angle), 0., 0., 1.1, 0., 0., 0., 0., 1., 0.f);
			}
			previous = instance;
		}
	}

	/* motion blur: motion=n adds n boxes that deform over the frame (Bezier motion-blur meshes, three time steps)
	 * and n moving instances (three matrices) of one pillar that is itself not rendered ("is_base_object") */
	int n_motion = 0;
	for(int a = 9; a < argc; ++a) if(strncmp(argv[a], "motion=", 7) == 0) n_motion = atoi(argv[a] + 7);
	for(int k = 0; k < n_motion; ++k)
	{
		char name[32];
		snprintf(name, sizeof name, "blur%02d", k);
		const double cx = scale * (0.2 + 0.6 * lcg()), cy = scale * (0.2 + 0.6 * lcg()), s = scale * 0.03, h = scale * 0.2;
		size_t blur_id = 0, blur_material = 0;
		yafaray_clearParamMap(pm);
		yafaray_setParamMapInt(pm, "num_faces", 6);
		yafaray_setParamMapInt(pm, "num_vertices", 8);
		yafaray_setParamMapString(pm, "type", "mesh");
		yafaray_setParamMapBool(pm, "motion_blur_bezier", YAFARAY_BOOL_TRUE);
		yafaray_setParamMapFloat(pm, "time_range_start", 0.f);
		yafaray_setParamMapFloat(pm, "time_range_end", 1.f);
		yafaray_createObject(scene, &blur_id, name, pm);
		for(int step = 0; step < 3; ++step)
		{
			/* the box slides along x, rises and twists a little: every vertex has its own path */
			const double dx = 0.06 * scale * step, dz = 0.03 * scale * step * (2 - step), tw = 0.25 * step;
			for(int v = 0; v < 8; ++v)
			{
				const double x = (v & 4) ? s : -s, y = (v & 2) ? s : -s, z = (v & 1) ? h : -0.6;
				yafaray_addVertexTimeStep(scene, blur_id, cx + dx + x * cos(tw) - y * sin(tw), cy + x * sin(tw) + y * cos(tw), z + ((v & 1) ? dz : 0.), (unsigned char) step);
			}
		}
		yafaray_getMaterialId(scene, &blur_material, "boxes");
		yafaray_addQuad(scene, blur_id, 2, 0, 1, 3, blur_material);
		yafaray_addQuad(scene, blur_id, 3, 7, 6, 2, blur_material);
		yafaray_addQuad(scene, blur_id, 7, 5, 4, 6, blur_material);
		yafaray_addQuad(scene, blur_id, 0, 4, 5, 1, blur_material);
		yafaray_addQuad(scene, blur_id, 0, 2, 6, 4, blur_material);
		yafaray_addQuad(scene, blur_id, 5, 7, 3, 1, blur_material);
		yafaray_initObject(scene, blur_id, blur_material);
	}
	if(n_motion > 0)
	{
		const double s = 0.02 * scale;
		size_t base_id = 0, base_material = 0;
		yafaray_clearParamMap(pm);
		yafaray_setParamMapInt(pm, "num_faces", 6);
		yafaray_setParamMapInt(pm, "num_vertices", 8);
		yafaray_setParamMapString(pm, "type", "mesh");
		yafaray_setParamMapBool(pm, "is_base_object", YAFARAY_BOOL_TRUE);
		yafaray_createObject(scene, &base_id, "moving_base", pm);
		for(int v = 0; v < 8; ++v) yafaray_addVertex(scene, base_id, (v & 4) ? s : -s, (v & 2) ? s : -s, (v & 1) ? 0.18 * scale : -0.6);
		yafaray_getMaterialId(scene, &base_material, "boxes");
		yafaray_addQuad(scene, base_id, 2, 0, 1, 3, base_material);
		yafaray_addQuad(scene, base_id, 3, 7, 6, 2, base_material);
		yafaray_addQuad(scene, base_id, 7, 5, 4, 6, base_material);
		yafaray_addQuad(scene, base_id, 0, 4, 5, 1, base_material);
		yafaray_addQuad(scene, base_id, 0, 2, 6, 4, base_material);
		yafaray_addQuad(scene, base_id, 5, 7, 3, 1, base_material);
		yafaray_initObject(scene, base_id, base_material);
		for(int k = 0; k < n_motion; ++k)
		{
			const double px = scale * (0.15 + 0.7 * lcg()), py = scale * (0.15 + 0.7 * lcg());
			const size_t instance = yafaray_createInstance(scene);
			yafaray_addInstanceObject(scene, instance, base_id);
			for(int step = 0; step < 3; ++step)
			{
				const double ang = 0.5 * step + k, c = cos(ang), sn = sin(ang), tx = px + 0.05 * scale * step, ty = py - 0.03 * scale * step * step;
				yafaray_addInstanceMatrix(scene, instance, c, -sn, 0., tx, sn, c, 0., ty, 0., 0., 1. + 0.1 * step, 0., 0., 0., 0., 1., 0.5f * step);
			}
		}
	}

	/* spheres: objects of type "sphere", src/geometry/object/object.cc:80-90 */
	int n_spheres = 0;
	for(int a = 9; a < argc; ++a) if(strncmp(argv[a], "spheres=", 8) == 0) n_spheres = atoi(argv[a] + 8);
	for(int k = 0; k < n_spheres; ++k)
	{
		char name[32];
		snprintf(name, sizeof name, "sphere%03d", k);
		const double angle = 2.399963229728653 * k, rad = scale * 0.42 * sqrt((k + 0.5) / n_spheres);
		yafaray_clearParamMap(pm);
		yafaray_setParamMapString(pm, "type", "sphere");
		yafaray_setParamMapVector(pm, "center", 0.5 * scale + rad * cos(angle), 0.5 * scale + rad * sin(angle), scale * (0.06 + 0.05 * (k % 4)));
		yafaray_setParamMapFloat(pm, "radius", (float) (scale * (0.012 + 0.004 * (k % 5))));
		yafaray_setParamMapString(pm, "material", (k & 1) ? "boxes" : "ground");
		yafaray_createObject(scene, &object_id, name, pm);
	}

	/* lights */
	yafaray_clearParamMap(pm);
	yafaray_setParamMapColor(pm, "color", 1.f, 0.95f, 0.9f, 1.f);
	yafaray_setParamMapVector(pm, "from", 0.2 * scale, 0.1 * scale, 0.9 * scale);
	yafaray_setParamMapFloat(pm, "power", 60.f);
	yafaray_setParamMapString(pm, "type", "pointlight");
	yafaray_createLight(scene, "point", pm);
	yafaray_clearParamMap(pm);
	yafaray_setParamMapColor(pm, "color", 0.9f, 0.95f, 1.f, 1.f);
	yafaray_setParamMapVector(pm, "corner", 0.7 * scale, 0.6 * scale, 0.8 * scale);
	yafaray_setParamMapVector(pm, "point1", 0.9 * scale, 0.6 * scale, 0.8 * scale);
	yafaray_setParamMapVector(pm, "point2", 0.7 * scale, 0.8 * scale, 0.8 * scale);
	yafaray_setParamMapFloat(pm, "power", 25.f);
	yafaray_setParamMapInt(pm, "samples", 4);
	yafaray_setParamMapString(pm, "type", "arealight");
	yafaray_createLight(scene, "area", pm);

	yafaray_clearParamMap(pm);
	yafaray_setParamMapColor(pm, "color", 0.5f, 0.6f, 0.8f, 1.f);
	yafaray_setParamMapFloat(pm, "power", 0.4f);
	yafaray_setParamMapString(pm, "type", "constant");
	yafaray_defineBackground(scene, pm);

	/* accelerator: exactly what tests/test02/test02.c:5193-5197 does */
	yafaray_clearParamMap(pm);
	yafaray_setParamMapString(pm, "type", accel);
	for(int a = 9; a < argc; ++a)
	{
		char key[64];
		const char *eq = strchr(argv[a], '=');
		if(!eq || (size_t) (eq - argv[a]) >= sizeof key || argv[a][1] == ':') continue;
		memcpy(key, argv[a], (size_t) (eq - argv[a]));
		key[eq - argv[a]] = 0;
		if(strcmp(key, "film_save") == 0) { film_save = eq + 1; continue; }
		if(strcmp(key, "instances") == 0 || strcmp(key, "spheres") == 0 || strcmp(key, "motion") == 0 || strcmp(key, "rerender_hidden") == 0) continue; /* handled elsewhere */
		if(strcmp(key, "tile_shard") == 0)
		{
			int shard_index = 0, shard_count = 1;
			if(sscanf(eq + 1, "%d/%d", &shard_index, &shard_count) != 2 || shard_count < 1 || shard_index < 0 || shard_index >= shard_count)
			{
				fprintf(stderr, "render_bench: bad tile_shard '%s' (expected index/count)\n", eq + 1);
				return 2;
			}
			if(strcmp(accel, "b200-kdtree") == 0)
			{
				yafaray_setParamMapInt(pm, "tile_shard_index", shard_index);
				yafaray_setParamMapInt(pm, "tile_shard_count", shard_count);
			}
			else setenv("B200_TILE_SHARD", eq + 1, 1);
			continue;
		}
		yafaray_setParamMapInt(pm, key, atoi(eq + 1));
	}
	yafaray_setSceneAcceleratorParams(scene, pm);

	/* integrator */
	yafaray_clearParamMap(pm);
	yafaray_setParamMapString(pm, "type", integrator);
	yafaray_setParamMapInt(pm, "raydepth", 3);
	yafaray_setParamMapInt(pm, "threads", threads);
	if(strcmp(integrator, "pathtracing") == 0)
	{
		yafaray_setParamMapInt(pm, "bounces", 3);
		yafaray_setParamMapInt(pm, "path_samples", 1);
		yafaray_setParamMapString(pm, "caustic_type", "none");
	}
	for(int a = 9; a < argc; ++a)
	{
		char key[64];
		const char *eq = strchr(argv[a], '=');
		if(!eq || argv[a][1] != ':' || (size_t) (eq - argv[a] - 2) >= sizeof key) continue;
		memcpy(key, argv[a] + 2, (size_t) (eq - argv[a] - 2));
		key[eq - argv[a] - 2] = 0;
		if(argv[a][0] == 'i') yafaray_setParamMapInt(pm, key, atoi(eq + 1));
		else if(argv[a][0] == 'b') yafaray_setParamMapBool(pm, key, atoi(eq + 1) ? YAFARAY_BOOL_TRUE : YAFARAY_BOOL_FALSE);
		else if(argv[a][0] == 'f') yafaray_setParamMapFloat(pm, key, (float) atof(eq + 1));
		else if(argv[a][0] == 'b') yafaray_setParamMapBool(pm, key, atoi(eq + 1) ? YAFARAY_BOOL_TRUE : YAFARAY_BOOL_FALSE);
	}
	yafaray_SurfaceIntegrator *surface_integrator = yafaray_createSurfaceIntegrator(logger, "surface integrator", pm);

	/* film, layer, camera, output */
	yafaray_clearParamMap(pm);
	yafaray_setParamMapInt(pm, "width", width);
	yafaray_setParamMapInt(pm, "height", height);
	yafaray_setParamMapInt(pm, "AA_passes", 1);
	yafaray_setParamMapInt(pm, "AA_minsamples", aa_samples);
	yafaray_setParamMapInt(pm, "threads", threads);
	if(film_save)
	{
		yafaray_setParamMapString(pm, "film_load_save_mode", "save");
		yafaray_setParamMapString(pm, "film_load_save_path", film_save);
	}
	yafaray_Film *film = yafaray_createFilm(logger, surface_integrator, "film", pm);
	yafaray_clearParamMap(pm);
	yafaray_setParamMapString(pm, "exported_image_name", "Combined");
	yafaray_setParamMapString(pm, "exported_image_type", "ColorAlpha");
	yafaray_setParamMapString(pm, "image_type", "ColorAlpha");
	yafaray_setParamMapString(pm, "type", "combined");
	yafaray_defineLayer(film, pm);
	yafaray_clearParamMap(pm);
	yafaray_setParamMapFloat(pm, "focal", 1.1f);
	yafaray_setParamMapVector(pm, "from", 0.5 * scale, -0.45 * scale, 0.55 * scale);
	yafaray_setParamMapVector(pm, "to", 0.5 * scale, 0.5 * scale, 0.0);
	yafaray_setParamMapVector(pm, "up", 0.5 * scale, -0.45 * scale, 0.55 * scale + 1.0);
	yafaray_setParamMapInt(pm, "resx", width);
	yafaray_setParamMapInt(pm, "resy", height);
	yafaray_setParamMapString(pm, "type", "perspective");
	yafaray_defineCamera(film, pm);
	yafaray_clearParamMap(pm);
	yafaray_setParamMapString(pm, "image_path", out_path);
	yafaray_setParamMapBool(pm, "denoise_enabled", YAFARAY_BOOL_FALSE);
	yafaray_createOutput(film, "output_tga", pm);
	const double t_built = now();

	yafaray_RenderMonitor *render_monitor = yafaray_createRenderMonitor(NULL, NULL, YAFARAY_DISPLAY_CONSOLE_HIDDEN);
	yafaray_RenderControl *render_control = yafaray_createRenderControl();
	yafaray_setRenderControlForNormalStart(render_control);
	yafaray_SceneModifiedFlags flags = yafaray_checkAndClearSceneModifiedFlags(scene);
	yafaray_preprocessScene(scene, render_control, flags);
	const double t_pre = now();
	yafaray_preprocessSurfaceIntegrator(render_monitor, surface_integrator, render_control, scene);
	const int profile = getenv("B200_PROF") != NULL;
	if(profile) prof_start();
	const double t_render0 = now();
	yafaray_render(render_control, render_monitor, surface_integrator, film);
	double t_render1 = now();
	if(profile) prof_stop();
	int rerender_hidden = 0;
	for(int a = 9; a < argc; ++a) if(strncmp(argv[a], "rerender_hidden=", 16) == 0) rerender_hidden = atoi(argv[a] + 16);
	if(rerender_hidden)
	{
		yafaray_clearParamMap(pm);
		yafaray_setParamMapColor(pm, "color", 0.3f, 0.45f, 0.8f, 1.f);
		yafaray_setParamMapFloat(pm, "diffuse_reflect", 0.9f);
		yafaray_setParamMapString(pm, "visibility", "invisible");
		yafaray_setParamMapString(pm, "type", "shinydiffusemat");
		yafaray_createMaterial(scene, &material_id, "boxes", pm, pml); /* same name: replaces the material the faces refer to */
		flags = yafaray_checkAndClearSceneModifiedFlags(scene);
		yafaray_preprocessScene(scene, render_control, flags);
		yafaray_preprocessSurfaceIntegrator(render_monitor, surface_integrator, render_control, scene);
		yafaray_setRenderControlForNormalStart(render_control);
		yafaray_render(render_control, render_monitor, surface_integrator, film);
		t_render1 = now();
	}
	printf("RENDER_BENCH {\"accelerator\": \"%s\", \"integrator\": \"%s\", \"triangles\": %d, \"width\": %d, \"height\": %d, \"aa_samples\": %d, \"threads\": %d, "
	       "\"scene_seconds\": %.3f, \"preprocess_seconds\": %.3f, \"render_seconds\": %.3f}\n",
	       accel, integrator, 2 * cells * cells + 72 + (n_instances > 0 ? 6 * (n_instances + 1) : 0) + n_spheres + 12 * n_motion, width, height, aa_samples, threads, t_built - t_scene, t_pre - t_built, t_render1 - t_render0);
	yafaray_destroyRenderControl(render_control);
	yafaray_destroyRenderMonitor(render_monitor);
	yafaray_destroyFilm(film);
	yafaray_destroySurfaceIntegrator(surface_integrator);
	yafaray_destroyScene(scene);
	yafaray_destroyLogger(logger);
	yafaray_destroyParamMapList(pml);
	yafaray_destroyParamMap(pm);
	return 0;
}
