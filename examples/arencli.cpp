// arencli — the caller of the hot path (examples/arencli.rs:29-62 of the reference), on the B200 core.
//   arencli INPUT [-t NUM] [--device N] [--spp-scale K]
// Reads a scene description (cb.json format), renders it with the PT renderer on the GPU, writes the PNG
// named by "outputfilename" and prints `Done! Time used: {:.4}s` like the reference.
// `-t/--thread` is accepted for compatibility (the reference sizes its rayon pool with it) and ignored.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "../include/arn_host.h"

int main(int argc, char** argv) {
    std::string input, out_override; int device = 0;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "-t") || !std::strcmp(argv[i], "--thread")) { i++; }
        else if (!std::strcmp(argv[i], "--device") && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "-o") && i + 1 < argc) out_override = argv[++i];
        else input = argv[i];
    }
    if (input.empty()) { std::fprintf(stderr, "usage: arencli INPUT [-t NUM] [--device N] [-o FILE]\n"); return 2; }
    arn_hscene* hs = nullptr; arn_hscene_create(&hs);
    arn_camera cam; arn_film film; arn_sampler smp; arn_pt_params prm; char outname[1024];
    int rc = arn_hscene_load_json(hs, input.c_str(), nullptr, &cam, &film, &smp, &prm, outname, sizeof outname);
    if (rc != ARN_OK) { std::fprintf(stderr, "Scene parsing failed: %s\n", arn_hscene_last_error(hs)); return 1; }   // arencli.rs:52
    if ((rc = arn_hscene_build(hs, ARN_BVH_SAH)) != ARN_OK) { std::fprintf(stderr, "BVH::new failed: %s\n", arn_hscene_last_error(hs)); return 1; }
    arn_ctx* ctx = nullptr;
    if ((rc = arn_ctx_create(device, &ctx)) != ARN_OK) { std::fprintf(stderr, "%s\n", arn_last_error(nullptr)); return 1; }
    arn_scene* scene = nullptr;
    if ((rc = arn_scene_upload(ctx, arn_hscene_desc(hs), &scene)) != ARN_OK) { std::fprintf(stderr, "%s\n", arn_last_error(ctx)); return 1; }
    std::printf("Rendering...\n");
    size_t w = (size_t)(film.crop_max_x - film.crop_min_x), h = (size_t)(film.crop_max_y - film.crop_min_y);
    float* acc = (float*)std::calloc(w * h * 4, sizeof(float));
    auto t0 = std::chrono::steady_clock::now();                      // the reference times render() incl. merge + PNG (arencli.rs:54-61)
    arn_stats st;
    rc = arn_render_pt(scene, &cam, &film, &smp, &prm, acc, &st);
    if (rc != ARN_OK) { std::fprintf(stderr, "render failed: %s\n", arn_last_error(ctx)); return 1; }
    const char* out = out_override.empty() ? outname : out_override.c_str();
    if (arn_save_png(out, acc, (uint32_t)w, (uint32_t)h) != ARN_OK) std::fprintf(stderr, "Path tracing result saving at %s failed\n", out);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("Done! Time used: %.4fs\n", secs);
    std::printf("  %llu camera samples, %llu rays, %.1f Mrays/s on the device (%.1f ms), %llu invalid samples\n", (unsigned long long)st.camera_rays,
                (unsigned long long)(st.extend_rays + st.shadow_rays + st.mis_rays), (st.extend_rays + st.shadow_rays + st.mis_rays) / st.gpu_ms / 1e3, st.gpu_ms,
                (unsigned long long)st.invalid_samples);
    std::free(acc); arn_scene_destroy(scene); arn_ctx_destroy(ctx); arn_hscene_destroy(hs);
    return 0;
}
