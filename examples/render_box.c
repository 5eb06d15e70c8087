/* render_box.c — the C-ABI from plain C99: render one of the built-in boxes and write the image
 * main() of the reference would write (src/smallvcm.cxx:311-320).
 *
 *   gcc -std=c99 -Iinclude examples/render_box.c -Lsmallvcm_b200/lib -lsvcm \
 *       -Wl,-rpath,$PWD/smallvcm_b200/lib -o render_box
 *   ./render_box 1 vcm 512 16 out.bmp          (scene, algorithm, resolution, iterations, file)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "svcm.h"

static int algorithm_id(const char *name)
{
    static const char *names[] = { "lt", "ppm", "bpm", "bpt", "vcm", "pt", "el" };   /* SVCM_ALG_* order */
    int i;
    for(i = 0; i < 7; i++)
        if(strcmp(name, names[i]) == 0) return i;
    return -1;
}

int main(int argc, char **argv)
{
    const int scene_id    = argc > 1 ? atoi(argv[1]) : 1;
    const int algorithm   = algorithm_id(argc > 2 ? argv[2] : "vcm");
    const int resolution  = argc > 3 ? atoi(argv[3]) : 512;
    const int iterations  = argc > 4 ? atoi(argv[4]) : 4;
    const char *file      = argc > 5 ? argv[5] : "out.bmp";
    svcm_scene *scene_storage = NULL;
    const SvcmSceneDesc *scene = NULL;
    SvcmParams params;
    SvcmStats stats;
    svcm_renderer *renderer = NULL;
    int rc;

    if(algorithm < 0) { fprintf(stderr, "unknown algorithm\n"); return 2; }
    rc = svcm_scene_cornell_box(svcm_scene_config_mask(scene_id), resolution, resolution, &scene_storage, &scene);
    if(rc != SVCM_OK) { fprintf(stderr, "scene: %s\n", svcm_last_error(NULL)); return 2; }

    memset(&params, 0, sizeof(params));
    params.algorithm = algorithm;
    params.radiusFactor = 0.003f;            /* Config defaults, src/config.hxx:192-193 */
    params.radiusAlpha = 0.75f;
    params.seed = 1234;
    params.device = 0; params.rank = 0; params.worldSize = 1;
    rc = svcm_create(scene, &params, &renderer);
    if(rc != SVCM_OK) { fprintf(stderr, "create: %s\n", svcm_last_error(NULL)); svcm_scene_free(scene_storage); return 2; }

    /* minPathLength 0, maxPathLength 10: the driver's defaults (src/config.hxx:190-191) */
    rc = svcm_run_iterations(renderer, 0, iterations, 0u, 10u);
    if(rc == SVCM_OK) rc = svcm_save_image(renderer, file, 2.2f);
    if(rc == SVCM_OK) rc = svcm_get_stats(renderer, &stats);
    if(rc != SVCM_OK) fprintf(stderr, "render: %s\n", svcm_last_error(renderer));
    else printf("%s: %d iteration(s), %lld light vertices in the last one\n", file, iterations,
                (long long)stats.lightVertices);

    svcm_destroy(renderer);
    svcm_scene_free(scene_storage);
    return rc == SVCM_OK ? 0 : 2;
}
