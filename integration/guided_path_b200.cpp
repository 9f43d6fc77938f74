/*
 * guided_path_b200.cpp -- Mitsuba 0.5 integrator plugin that runs the B200-native guided path tracer (libppg_b200.so)
 * behind the plugin surface of the reference's GuidedPathTracer:
 *
 *     <integrator type="guided_path_b200"> ... the same parameters as type="guided_path" ... </integrator>
 *
 * What Mitsuba binds (and what this file provides):
 *   - extern "C" CreateInstance(const Properties &) / GetDescription()      MTS_EXPORT_PLUGIN, include/mitsuba/core/cobject.h:99-107
 *     (reference: mitsuba/src/integrators/path/guided_path.cpp:2422); loaded by src/libcore/plugin.cpp:40-96
 *   - Integrator::render(Scene*, RenderQueue*, const RenderJob*, int, int, int) -> bool     include/mitsuba/render/integrator.h:74-75
 *     (reference implementation guided_path.cpp:1516-1585) and Integrator::cancel()        integrator.h:77-84 (guided_path.cpp:1643-1648)
 *   - the constructor reads the XML parameters with the reference's names, defaults and validation (guided_path.cpp:1014-1085,
 *     src/librender/integrator.cpp:190-225) by forwarding every (name, value) pair to ppg_params_set.
 *
 * THIS FILE COMPILES ONLY AGAINST A MITSUBA 0.5 TREE (SCons + boost + xerces-c + OpenEXR are not available in the repository's build
 * container); build it as any other plugin, e.g. in mitsuba/src/integrators/SConscript:
 *     plugins += env.SharedLibrary('guided_path_b200', ['path/guided_path_b200.cpp'], LIBS = env['LIBS'] + ['ppg_b200'],
 *                                  CPPPATH = env['CPPPATH'] + ['#/../practical-path-guiding_b200/../include'])
 * Everything below the Mitsuba types is the plain C ABI of include/ppg.h, which IS built and tested in this repository.
 *
 * flatten(Scene *): Mitsuba 0.5 has no public accessor for the children of a live BSDF (twosided / mask / bumpmap keep their nested BSDF and
 * textures in protected members; include/mitsuba/render/bsdf.h offers getDiffuseReflectance(its) and little else), so materials cannot be
 * read back from the Scene object.  The scene's own XML says everything: flatten() converts scene->getSourceFile() once with the
 * repository's converter (`python -m ppg_b200.convert scene.xml scene.ppgscene`, or $PPG_B200_CONVERTER) into the flat array form
 * ppg_scene_file_load reads, and then takes whatever the host may have changed since loading from the LIVE objects: the sensor's
 * world transform, field of view, clip planes and the film's crop size.
 */
#include <mitsuba/render/scene.h>
#include <mitsuba/render/renderqueue.h>
#include <mitsuba/render/renderjob.h>
#include <mitsuba/core/statistics.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/fstream.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "ppg.h"

MTS_NAMESPACE_BEGIN

class GuidedPathTracerB200 : public Integrator {
public:
    GuidedPathTracerB200(const Properties &props) : Integrator(props), m_handle(NULL) {
        ppg_params_default(&m_params);
        /* the reference's parameter set (guided_path.cpp:1014-1085 + MonteCarloIntegrator); values travel as XML strings so that
           ppg_params_set applies exactly the reference's validation (an unknown enum string is an error, like Assert(false)) */
        static const char *strings[] = {"nee", "sampleCombination", "spatialFilter", "directionalFilter", "bsdfSamplingFractionLoss", "budgetType"};
        static const char *integers[] = {"sdTreeMaxMemory", "sTreeThreshold", "sppPerPass", "maxDepth", "rrDepth"};
        static const char *floats[] = {"dTreeThreshold", "bsdfSamplingFraction", "budget"};
        static const char *booleans[] = {"dumpSDTree", "strictNormals", "hideEmitters"};
        for (size_t i = 0; i < sizeof(strings) / sizeof(*strings); ++i)
            if (props.hasProperty(strings[i])) set(strings[i], props.getString(strings[i]));
        for (size_t i = 0; i < sizeof(integers) / sizeof(*integers); ++i)
            if (props.hasProperty(integers[i])) set(integers[i], formatString("%i", props.getInteger(integers[i])));
        for (size_t i = 0; i < sizeof(floats) / sizeof(*floats); ++i)
            if (props.hasProperty(floats[i])) set(floats[i], formatString("%.9g", (double) props.getFloat(floats[i])));
        for (size_t i = 0; i < sizeof(booleans) / sizeof(*booleans); ++i)
            if (props.hasProperty(booleans[i])) set(booleans[i], props.getBoolean(booleans[i]) ? "true" : "false");
        if (ppg_params_validate(&m_params) != PPG_OK)
            Log(EError, "%s", ppg_last_error());
        m_device = props.getInteger("device", -1);
    }

    GuidedPathTracerB200(Stream *stream, InstanceManager *manager) : Integrator(stream, manager), m_handle(NULL) {
        stream->read(&m_params, sizeof(m_params));
        m_device = stream->readInt();
    }

    virtual ~GuidedPathTracerB200() { if (m_handle) ppg_destroy(m_handle); }

    void serialize(Stream *stream, InstanceManager *manager) const {
        Integrator::serialize(stream, manager);
        stream->write(&m_params, sizeof(m_params));
        stream->writeInt(m_device);
    }

    bool preprocess(const Scene *, RenderQueue *, const RenderJob *, int, int, int) { return true; }

    /* Integrator::render (integrator.h:74-75; reference: guided_path.cpp:1516-1585) */
    bool render(Scene *scene, RenderQueue *queue, const RenderJob *job, int sceneResID, int sensorResID, int samplerResID) {
        ref<Sensor> sensor = scene->getSensor();
        ref<Film> film = sensor->getFilm();
        const Vector2i size = film->getCropSize();
        if (film->getCropOffset() != Point2i(0) || size != film->getSize())
            Log(EError, "guided_path_b200: crop windows are not supported");
        if (!scene->getMedia().empty())
            Log(EError, "guided_path_b200: participating media are not supported (nor by the reference, README.md:5-7)");

        ppg_scene_desc desc; ppg_scene_file *file = NULL;
        if (!flatten(scene, sensor, size, desc, file))
            return false;
        if (!m_handle && ppg_create(&m_params, m_device, &m_handle) != PPG_OK) {
            Log(EWarn, "guided_path_b200: %s", ppg_last_error());
            ppg_scene_file_free(file);
            return false;
        }
        int rc = ppg_set_scene(m_handle, &desc);
        ppg_scene_file_free(file);
        if (rc != PPG_OK) { Log(EWarn, "guided_path_b200: %s", ppg_last_error()); return false; }
        ppg_set_destination(m_handle, scene->getDestinationFile().string().c_str());       /* dumpSDTree writes <destination>-NN.sdt (GP:1191-1195) */

        /* progressive film: the reference puts every finished block into the film (renderproc.cpp:143-151) */
        Progress ctx; ctx.film = film.get(); ctx.queue = queue; ctx.job = job; ctx.size = size;
        ppg_set_film_callback(m_handle, &GuidedPathTracerB200::onFilm, &ctx);

        Log(EInfo, "Starting render job (%ix%i, CUDA device %i) ..", size.x, size.y, m_device);
        std::vector<float> rgb((size_t) size.x * size.y * 3);
        ppg_stats stats;
        rc = ppg_render(m_handle, &rgb[0], &stats);
        if (rc != PPG_OK && rc != PPG_ERR_CANCELLED) { Log(EWarn, "guided_path_b200: %s", ppg_last_error()); return false; }
        for (int i = 0; i < stats.n_iterations; ++i) {       /* the reference's per-iteration log lines (GP:1176-1186, 1323-1326) */
            const ppg_iteration_stats &s = stats.iterations[i];
            Log(EInfo, "ITERATION %d, %d passes%s: %.2f seconds, Total passes: %d, Var: %f; D-tree nodes avg %.1f, depth avg %.2f, stat. weight avg %.1f",
                s.iteration, s.passes, s.is_final ? " (FINAL)" : "", s.seconds, s.total_passes, s.variance, s.nodes_avg, s.depth_avg, s.weight_avg);
        }
        putFilm(film.get(), &rgb[0], size);
        queue->signalRefresh(job);
        return rc == PPG_OK;                                 /* a cancelled render returns false like the reference (GP:1270-1277) */
    }

    /* Integrator::cancel (integrator.h:77-84; reference: guided_path.cpp:1643-1648): asynchronous, from any thread */
    void cancel() { if (m_handle) ppg_cancel(m_handle); }

    void postprocess(const Scene *, RenderQueue *, const RenderJob *, int, int, int) { }

    std::string toString() const {
        std::ostringstream oss;
        oss << "GuidedPathTracerB200[" << endl << "  device = " << m_device << "," << endl << "  sppPerPass = " << m_params.spp_per_pass << "," << endl
            << "  sTreeThreshold = " << m_params.s_tree_threshold << "," << endl << "  budget = " << m_params.budget << endl << "]";
        return oss.str();
    }

    MTS_DECLARE_CLASS()
private:
    struct Progress { Film *film; RenderQueue *queue; const RenderJob *job; Vector2i size; };

    void set(const char *name, const std::string &value) {
        if (ppg_params_set(&m_params, name, value.c_str()) != PPG_OK)
            Log(EError, "%s", ppg_last_error());            /* Log(EError) throws, like the reference's Assert(false) / Log(EError) */
    }

    /* linear RGB (device or host pointer resolved by the caller) -> the film */
    static void putFilm(Film *film, const float *rgb, const Vector2i &size) {
        ref<Bitmap> bitmap = new Bitmap(Bitmap::ERGB, Bitmap::EFloat32, size);
        memcpy(bitmap->getFloat32Data(), rgb, (size_t) size.x * size.y * 3 * sizeof(float));
        film->setBitmap(bitmap);
    }

    static void onFilm(void *user, const float *rgb_dev, int width, int height, int /* passes */) {
        Progress *p = static_cast<Progress *>(user);
        std::vector<float> host((size_t) width * height * 3);
        if (ppg_copy_from_device(&host[0], rgb_dev, host.size() * sizeof(float)) != PPG_OK) return;
        putFilm(p->film, &host[0], p->size);
        p->queue->signalRefresh(p->job);
    }

    /* Scene -> ppg_scene_desc (see the header comment) */
    bool flatten(const Scene *scene, const Sensor *sensor, const Vector2i &size, ppg_scene_desc &desc, ppg_scene_file *&file) {
        const fs::path xml = scene->getSourceFile();
        fs::path flat = xml; flat.replace_extension(".ppgscene");
        if (!fs::exists(flat) || fs::last_write_time(flat) < fs::last_write_time(xml)) {
            const char *conv = getenv("PPG_B200_CONVERTER");
            const std::string cmd = std::string(conv ? conv : "python -m ppg_b200.convert") + " \"" + xml.string() + "\" \"" + flat.string() + "\"";
            Log(EInfo, "guided_path_b200: converting the scene: %s", cmd.c_str());
            if (std::system(cmd.c_str()) != 0) { Log(EWarn, "guided_path_b200: scene conversion failed"); return false; }
        }
        if (ppg_scene_file_load(flat.string().c_str(), &desc, &file, NULL) != PPG_OK) { Log(EWarn, "guided_path_b200: %s", ppg_last_error()); return false; }
        /* live sensor state (the GUI may have moved the camera): perspective.cpp:120-298 */
        const PerspectiveCamera *cam = dynamic_cast<const PerspectiveCamera *>(sensor);
        if (!cam) { Log(EWarn, "guided_path_b200: only the perspective sensor is supported"); ppg_scene_file_free(file); return false; }
        const Matrix4x4 &m = cam->getWorldTransform()->eval(0).getMatrix();
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) desc.camera.to_world[4 * r + c] = (float) m(r, c);
        desc.camera.x_fov_deg = (float) cam->getXFov();
        desc.camera.near_clip = (float) cam->getNearClip(); desc.camera.far_clip = (float) cam->getFarClip();
        desc.camera.film_width = size.x; desc.camera.film_height = size.y;
        /* Scene::getAABB() already holds kd-tree + sensor + emitter boxes (librender/scene.cpp:387-413) */
        const AABB &aabb = scene->getAABB();
        for (int i = 0; i < 3; ++i) { desc.aabb_min[i] = (float) aabb.min[i]; desc.aabb_max[i] = (float) aabb.max[i]; }
        return true;
    }

    ppg_params m_params;
    ppg_integrator *m_handle;
    int m_device;
};

MTS_IMPLEMENT_CLASS_S(GuidedPathTracerB200, false, Integrator)
MTS_EXPORT_PLUGIN(GuidedPathTracerB200, "Guided path tracer (B200)");
MTS_NAMESPACE_END
