// render_project — the drop-in path end to end in the reference's own language, no Python: loads a bisemutum project directory
// (host/project.hpp), hands it to the CUDA library through the C ABI, selects the renderer BY NAME as the engine does
// (project.toml `renderer = "..."` -> GraphicsManager::set_renderer, engine.cpp:147-148) and drives it like
// GraphicsManager::render_frame (graphics_manager.cpp:407-427): per frame prepare_renderer_per_frame_data, per camera
// prepare_renderer_per_camera_data + render_camera, then RenderGraph::execute. Writes the back buffer as a little-endian PFM.
//
//   render_project <project dir | model.gltf | model.glb> <out.pfm> [frames=16] [width height] [--merged] [--bloom threshold softness] [--renderer name]
//                  [--camera px py pz fx fy fz [yfov]] [--dir-light dx dy dz r g b strength] [--bounces n]
// A .gltf / .glb argument goes through the glTF importer (host/gltf.hpp = the editor's "Import Model (glTF)" action); a glTF import carries no
// camera or light (the reference ignores them too), so they come from the command line.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "project.hpp"
#include "gltf.hpp"

using namespace bi;

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <project dir> <out.pfm> [frames] [width height] [--merged] [--bloom threshold softness] [--renderer name]\n", argv[0]);
        return 2;
    }
    std::string dir = argv[1], out_path = argv[2], renderer_name = "CudaPathTracingRenderer";
    uint32_t frames = 16, width = 0, height = 0, accel = BPT_ACCEL_TWO_LEVEL;
    PostProcessVolume post;
    std::vector<std::string> pos;
    std::vector<float> cam_arg, light_arg;
    uint32_t bounces = 0;
    for (int i = 3; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--merged") accel = BPT_ACCEL_MERGED;
        else if (a == "--bloom" && i + 2 < argc) { post.bloom = true; post.bloom_threshold = (float)std::atof(argv[++i]); post.bloom_threshold_softness = (float)std::atof(argv[++i]); }
        else if (a == "--renderer" && i + 1 < argc) renderer_name = argv[++i];
        else if (a == "--camera" && i + 6 < argc) { for (int k = 0; k < 6; k++) cam_arg.push_back((float)std::atof(argv[++i])); if (i + 1 < argc && argv[i + 1][0] != '-') cam_arg.push_back((float)std::atof(argv[++i])); }
        else if (a == "--dir-light" && i + 7 < argc) { for (int k = 0; k < 7; k++) light_arg.push_back((float)std::atof(argv[++i])); }
        else if (a == "--bounces" && i + 1 < argc) bounces = (uint32_t)std::atoi(argv[++i]);
        else pos.push_back(a);
    }
    if (pos.size() >= 1) frames = (uint32_t)std::atoi(pos[0].c_str());
    if (pos.size() >= 3) { width = (uint32_t)std::atoi(pos[1].c_str()); height = (uint32_t)std::atoi(pos[2].c_str()); }

    project::Project prj;
    std::string err;
    auto ends_with = [&](const char* e) { return dir.size() >= std::strlen(e) && dir.compare(dir.size() - std::strlen(e), std::string::npos, e) == 0; };
    if (ends_with(".gltf") || ends_with(".glb")) {
        if (!project::import_gltf(dir, prj, err)) { std::fprintf(stderr, "import_gltf: %s\n", err.c_str()); return 1; }
    } else if (!project::load_project(dir, prj, err)) { std::fprintf(stderr, "load_project: %s\n", err.c_str()); return 1; }
    if (cam_arg.size() >= 6) {
        for (int k = 0; k < 3; k++) { prj.cam_position[k] = cam_arg[(size_t)k]; prj.cam_front[k] = cam_arg[(size_t)k + 3]; }
        if (cam_arg.size() > 6) prj.yfov = cam_arg[6];
    }
    if (light_arg.size() == 7) {                                                       // the light's direction is its transform's +Y axis (lights.cpp:52-63)
        DirectionalLightComponent l; l.color = {light_arg[3], light_arg[4], light_arg[5]}; l.strength = light_arg[6];
        LightTransform lt; lt.rotation[1] = light_arg[0]; lt.rotation[4] = light_arg[1]; lt.rotation[7] = light_arg[2];
        prj.lights.add(l, lt);
    }
    if (bounces) prj.path_tracing.max_bounces = bounces;
    if (!width || !height) { width = prj.target_width ? prj.target_width : 640; height = prj.target_height ? prj.target_height : 360; }

    bpt_config cfg{};
    cfg.device = 0; cfg.width = width; cfg.height = height;
    bpt_context* ctx = nullptr;
    bpt_status s = bpt_create(&cfg, &ctx);
    if (s != BPT_OK) { std::fprintf(stderr, "bpt_create: status %d (no CUDA device? there is no CPU path)\n", (int)s); return 1; }
    if ((s = project::upload_project(prj, ctx, accel, err)) != BPT_OK) { std::fprintf(stderr, "upload_project: %s (%s)\n", err.c_str(), bpt_last_error(ctx)); return 1; }

    gfx::GraphicsManager mgr(ctx);
    mgr.register_renderer<CudaPathTracingRenderer>();                                  // src/engine/register_renderer.cpp:9-11
    if (!mgr.set_renderer(renderer_name)) { std::fprintf(stderr, "renderer '%s' is not registered\n", renderer_name.c_str()); return 1; }
    auto* r = static_cast<CudaPathTracingRenderer*>(mgr.renderer());
    r->lights_ctx = prj.lights;
    r->settings.path_tracing = prj.path_tracing;
    r->post_process = post;
    std::vector<float> back((size_t)width * height * 4);
    r->back_buffer = back.data();

    gfx::Camera camera;
    camera.position = {prj.cam_position[0], prj.cam_position[1], prj.cam_position[2]};
    camera.front_dir = {prj.cam_front[0], prj.cam_front[1], prj.cam_front[2]};
    camera.up_dir = {prj.cam_up[0], prj.cam_up[1], prj.cam_up[2]};
    camera.yfov = prj.yfov; camera.near_z = prj.near_z; camera.far_z = prj.far_z;
    camera.projection_type = prj.orthographic ? gfx::ProjectionType::orthographic : gfx::ProjectionType::perspective;
    camera.set_target_extent(width, height);
    for (uint32_t f = 0; f < frames; f++) {
        camera.update_shader_params(f);
        r->path_tracing_pass.set_frame_count(f);
        r->prepare_renderer_per_frame_data();
        r->prepare_renderer_per_camera_data(camera);
        gfx::RenderGraph rg;
        r->render_camera(camera, rg);
        rg.execute();
        if (r->last_status() != BPT_OK) { std::fprintf(stderr, "frame %u: status %d: %s\n", f, (int)r->last_status(), bpt_last_error(ctx)); return 1; }
    }
    bpt_counters c{};
    bpt_get_counters(ctx, &c);
    FILE* fp = std::fopen(out_path.c_str(), "wb");
    if (!fp) { std::perror(out_path.c_str()); return 1; }
    std::fprintf(fp, "PF\n%u %u\n-1.0\n", width, height);
    for (uint32_t y = height; y-- > 0;)                                                // PFM rows run bottom to top
        for (uint32_t x = 0; x < width; x++) std::fwrite(&back[((size_t)y * width + x) * 4], sizeof(float), 3, fp);
    std::fclose(fp);
    std::printf("%s: %u frames of %ux%u, %zu drawables, %llu extend + %llu shadow rays -> %s\n", renderer_name.c_str(), frames, width, height,
                prj.drawables.size(), (unsigned long long)c.extend_rays, (unsigned long long)c.shadow_rays, out_path.c_str());
    bpt_destroy(ctx);
    return 0;
}
