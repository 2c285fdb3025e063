/* integration/test_hook.h -- lets the reference's OWN test clients (tests/testNN/testNN.c, compiled unmodified from
 * /root/reference with `gcc -include integration/test_hook.h`) select the accelerator type and deterministic
 * threading from the environment, without editing them:
 *   B200_ACCEL_TYPE=b200-kdtree      -> yafaray_setSceneAcceleratorParams(type=...) before yafaray_preprocessScene,
 *                                       exactly what tests/test02/test02.c:5193-5197 does by hand
 *   B200_DETERMINISTIC=1             -> threads=1 on integrator and film, tiles_order=linear
 *   B200_AA_PASSES=n                 -> overrides AA_passes of the film (shorter renders on the per-ray path)
 *   B200_WAVEFRONT_FIBERS=n          -> accelerator parameter "wavefront_fibers" (0 = per-ray calls only)
 *   B200_WAVEFRONT_BLOCK=n           -> accelerator parameter "wavefront_block"
 * TEST INFRASTRUCTURE ONLY. */
#ifndef B200_TEST_HOOK_H
#define B200_TEST_HOOK_H
#include <stdlib.h>
#include "yafaray_c_api.h"

static yafaray_Bool b200_hook_preprocessScene(yafaray_Scene *scene, const yafaray_RenderControl *render_control, yafaray_SceneModifiedFlags flags)
{
	const char *type = getenv("B200_ACCEL_TYPE");
	if(type && *type)
	{
		yafaray_ParamMap *pm = yafaray_createParamMap();
		yafaray_setParamMapString(pm, "type", type);
		if(getenv("B200_WAVEFRONT_FIBERS")) yafaray_setParamMapInt(pm, "wavefront_fibers", atoi(getenv("B200_WAVEFRONT_FIBERS")));
		if(getenv("B200_WAVEFRONT_GROUPS")) yafaray_setParamMapInt(pm, "wavefront_groups", atoi(getenv("B200_WAVEFRONT_GROUPS")));
		if(getenv("B200_WAVEFRONT_BLOCK")) yafaray_setParamMapInt(pm, "wavefront_block", atoi(getenv("B200_WAVEFRONT_BLOCK")));
		yafaray_setSceneAcceleratorParams(scene, pm);
		yafaray_destroyParamMap(pm);
		flags = (yafaray_SceneModifiedFlags) (flags | yafaray_checkAndClearSceneModifiedFlags(scene));
	}
	return yafaray_preprocessScene(scene, render_control, flags);
}

static yafaray_SurfaceIntegrator *b200_hook_createSurfaceIntegrator(yafaray_Logger *logger, const char *name, const yafaray_ParamMap *param_map)
{
	if(getenv("B200_DETERMINISTIC")) yafaray_setParamMapInt((yafaray_ParamMap *) param_map, "threads", 1);
	return yafaray_createSurfaceIntegrator(logger, name, param_map);
}

static yafaray_Film *b200_hook_createFilm(yafaray_Logger *logger, yafaray_SurfaceIntegrator *surface_integrator, const char *name, const yafaray_ParamMap *param_map)
{
	if(getenv("B200_DETERMINISTIC"))
	{
		yafaray_setParamMapInt((yafaray_ParamMap *) param_map, "threads", 1);
		yafaray_setParamMapString((yafaray_ParamMap *) param_map, "tiles_order", "linear");
	}
	if(getenv("B200_AA_PASSES")) yafaray_setParamMapInt((yafaray_ParamMap *) param_map, "AA_passes", atoi(getenv("B200_AA_PASSES")));
	return yafaray_createFilm(logger, surface_integrator, name, param_map);
}

#define yafaray_preprocessScene b200_hook_preprocessScene
#define yafaray_createSurfaceIntegrator b200_hook_createSurfaceIntegrator
#define yafaray_createFilm b200_hook_createFilm
#endif
