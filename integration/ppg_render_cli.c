/* ppg_render_cli.c -- a plain C99 host of libppg_b200.so: what `mitsuba scene.xml` does with the guided_path plugin, through include/ppg.h alone
 * (no Mitsuba, no Python at run time).
 *
 *   python -m ppg_b200.convert scene.xml scene.ppgscene [width height]        once, on any machine (reads the reference's scene XML unchanged)
 *   ppg_render_cli scene.ppgscene out.pfm [-D name=value ...] [--device N] [--sdt tree.sdt] [--check]
 *
 * -D overrides an integrator parameter with the XML's names and value strings (same validation and error text as the plugin constructor,
 * GP:1014-1085): -D budget=127 -D budgetType=spp -D nee=kickstart ...  --check stops after loading the scene and validating the parameters
 * (no CUDA device needed).  The film is written as a little-endian PFM (RGB float32, bottom row first) and the per-iteration block of the
 * reference's log (GP:1176-1186, 1323-1326) goes to stderr.  Exit code: 0, or the negated ppg_status.
 *
 * Build:  make -C integration        (cc -std=c99 -I../include ppg_render_cli.c -L<csrc> -lppg_b200)
 */
#include <ppg.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int die(int rc, const char *what) {
    fprintf(stderr, "ppg_render_cli: %s: %s (status %d)\n", what, ppg_last_error(), rc);
    return rc < 0 ? -rc : 1;
}

/* "name=value" -> ppg_params_set; returns the status */
static int set_param(ppg_params *prm, const char *assignment) {
    char name[64]; const char *eq = strchr(assignment, '=');
    if (!eq || eq == assignment || (size_t) (eq - assignment) >= sizeof(name)) { fprintf(stderr, "ppg_render_cli: expected name=value, got '%s'\n", assignment); return PPG_ERR_INVALID_ARGUMENT; }
    memcpy(name, assignment, (size_t) (eq - assignment)); name[eq - assignment] = '\0';
    return ppg_params_set(prm, name, eq + 1);
}

static int write_pfm(const char *path, const float *rgb, int w, int h) {
    FILE *f = fopen(path, "wb");
    int y;
    if (!f) return -1;
    fprintf(f, "PF\n%d %d\n-1.0\n", w, h);
    for (y = h - 1; y >= 0; --y) fwrite(rgb + (size_t) y * w * 3, sizeof(float), (size_t) w * 3, f);
    return fclose(f);
}

int main(int argc, char **argv) {
    const char *scenePath = NULL, *outPath = NULL, *sdtPath = NULL;
    const char *overrides[64]; int nOverrides = 0, device = 0, checkOnly = 0, i, rc;
    ppg_scene_desc desc; ppg_scene_file *file = NULL; const char *props = NULL;
    ppg_params prm; ppg_integrator *h = NULL; ppg_stats st; float *rgb;

    for (i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-D") && i + 1 < argc) { if (nOverrides < 64) overrides[nOverrides++] = argv[++i]; }
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--sdt") && i + 1 < argc) sdtPath = argv[++i];
        else if (!strcmp(argv[i], "--check")) checkOnly = 1;
        else if (!scenePath) scenePath = argv[i];
        else if (!outPath) outPath = argv[i];
        else { fprintf(stderr, "ppg_render_cli: unexpected argument '%s'\n", argv[i]); return 2; }
    }
    if (!scenePath || (!outPath && !checkOnly)) {
        fprintf(stderr, "usage: ppg_render_cli scene.ppgscene out.pfm [-D name=value ...] [--device N] [--sdt tree.sdt] [--check]\n%s, C ABI version %d\n", ppg_description(), ppg_abi_version());
        return 2;
    }
    rc = ppg_scene_file_load(scenePath, &desc, &file, &props);
    if (rc) return die(rc, scenePath);

    /* CreateInstance(props): defaults, then the XML's <integrator> block ("name=value" lines), then the command line */
    ppg_params_default(&prm);
    if (props) {
        const char *p = props;
        while (*p) {
            char line[256]; const char *nl = strchr(p, '\n'); size_t n = nl ? (size_t) (nl - p) : strlen(p);
            if (n && n < sizeof(line)) { memcpy(line, p, n); line[n] = '\0'; rc = set_param(&prm, line); if (rc) { ppg_scene_file_free(file); return die(rc, line); } }
            p += n + (nl ? 1 : 0);
        }
    }
    for (i = 0; i < nOverrides; ++i) { rc = set_param(&prm, overrides[i]); if (rc) { ppg_scene_file_free(file); return die(rc, overrides[i]); } }
    rc = ppg_params_validate(&prm);
    if (rc) { ppg_scene_file_free(file); return die(rc, "parameters"); }
    fprintf(stderr, "scene: %u triangles, %u shapes, %u materials, %u area emitters%s, film %d x %d\n", desc.n_triangles, desc.n_shapes, desc.n_bsdfs, desc.n_emitters,
            desc.envmap.width ? " + environment map" : "", desc.camera.film_width, desc.camera.film_height);
    if (checkOnly) { ppg_scene_file_free(file); fprintf(stderr, "check ok\n"); return 0; }

    rc = ppg_create(&prm, device, &h);
    if (rc) { ppg_scene_file_free(file); return die(rc, "ppg_create"); }
    rc = ppg_set_scene(h, &desc);                       /* Scene -> HBM; the host arrays may go afterwards */
    {
        const int w = desc.camera.film_width, hgt = desc.camera.film_height;
        ppg_scene_file_free(file);
        if (rc) { ppg_destroy(h); return die(rc, "ppg_set_scene"); }
        rgb = (float *) malloc(sizeof(float) * 3 * (size_t) w * (size_t) hgt);
        if (!rgb) { ppg_destroy(h); fprintf(stderr, "ppg_render_cli: out of memory\n"); return 1; }
        rc = ppg_render(h, rgb, &st);                   /* Integrator::render(): PPG_OK, or PPG_ERR_CANCELLED with the partial film */
        if (rc && rc != PPG_ERR_CANCELLED) { free(rgb); ppg_destroy(h); return die(rc, "ppg_render"); }
        for (i = 0; i < st.n_iterations && i < PPG_MAX_ITERATIONS; ++i) {
            const ppg_iteration_stats *it = &st.iterations[i];
            fprintf(stderr, "ITERATION %d%s, %d passes, %.2f s, Var: %g | D-tree depth %d..%d (avg %.2f), nodes %llu..%llu (avg %.1f), stat. weight avg %.1f, %u S-tree leaves\n",
                    it->iteration, it->is_final ? " (FINAL)" : "", it->passes, it->seconds, it->variance, it->depth_min, it->depth_max, it->depth_avg,
                    (unsigned long long) it->nodes_min, (unsigned long long) it->nodes_max, it->nodes_avg, it->weight_avg, it->s_tree_leaves);
        }
        fprintf(stderr, "%llu paths, %llu vertices in %.3f s (%.1f Msamples/s on the device), %llu kernel launches\n", (unsigned long long) st.total_paths,
                (unsigned long long) st.total_vertices, st.render_seconds, st.render_device_ms > 0 ? (double) st.total_vertices / st.render_device_ms * 1e-3 : 0.0,
                (unsigned long long) st.kernel_launches);
        if (write_pfm(outPath, rgb, w, hgt)) { fprintf(stderr, "ppg_render_cli: cannot write %s\n", outPath); free(rgb); ppg_destroy(h); return 7; }
        free(rgb);
    }
    if (sdtPath) { const int rc2 = ppg_dump_sdtree(h, sdtPath); if (rc2) { ppg_destroy(h); return die(rc2, sdtPath); } }
    ppg_destroy(h);
    return rc == PPG_ERR_CANCELLED ? 5 : 0;
}
