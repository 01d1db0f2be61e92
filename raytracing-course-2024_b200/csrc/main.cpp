// main.cpp -- the `raytracing-engine` CLI with the reference's argv (src/main.rs:28-73, run.sh:2):
//     raytracing-engine <scene.gltf|scene.txt> <width> <height> <samples> <out.ppm> [png_basename]
//     raytracing-engine <scene.txt> <out.ppm>                       (the text era's form: DIMENSIONS / SAMPLES come from the file)
// A ".txt" scene is the course's text format (rt_scene_load_text: own spec, reference HEAD dropped its parser, main.rs:48);
// width / height / samples <= 0 keep the file's DIMENSIONS / SAMPLES.  Scene ingest, rendering and output all go through the C ABI (include/rt_api.h); this file is what main.rs
// becomes once render_scene is the B200 library.  Differences from the reference, on purpose:
//   * errors are reported and the exit code is non-zero (the reference panics);
//   * the PPM is truncated, not appended to (main.rs:62-66 opens with append(true), which stacks a second image
//     behind an existing file); set RT_PPM_APPEND=1 to get the reference behaviour;
//   * the optional 6th argument (PNG dump, main.rs:68-72) writes `<basename>.png` through rt_write_png.
// Extra knobs are environment variables only, so the positional CLI stays drop-in:
//   RT_SEED (Philox key, default 0), RT_DEVICE (first CUDA device, default 0), RT_GPUS (GPUs of this box to shard the
//   samples over: one NCCL reduce of the framebuffer at the end, default 1), RT_STATS=1 (work counters).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/rt_api.h"

static long long env_ll(const char* name, long long dflt) { const char* v = std::getenv(name); return v && *v ? std::atoll(v) : dflt; }

int main(int argc, char** argv) {
    if (argc != 3 && argc < 6) {
        std::fprintf(stderr, "usage: %s <scene.gltf|scene.txt> <width> <height> <samples> <out.ppm> [png_basename]\n       %s <scene.txt> <out.ppm>\n", argv[0], argv[0]);
        return 2;
    }
    const char* input_scene = argv[1];
    int width = argc == 3 ? 0 : std::atoi(argv[2]), height = argc == 3 ? 0 : std::atoi(argv[3]), samples = argc == 3 ? 0 : std::atoi(argv[4]);
    const char* output_ppm = argc == 3 ? argv[2] : argv[5];
    const int device = (int)env_ll("RT_DEVICE", 0);

    int n_gpus = (int)env_ll("RT_GPUS", 1);
    if (n_gpus < 1) n_gpus = 1;
    std::vector<RtScene*> scenes((size_t)n_gpus, nullptr);
    auto destroy_all = [&]() { for (RtScene* s : scenes) if (s) rt_scene_destroy(s); };
    for (int g = 0; g < n_gpus; ++g) {                                                             // main.rs:45-47 on every device
        if (rt_scene_load(input_scene, width, height, samples, device + g, &scenes[(size_t)g]) != RT_OK) {
            std::fprintf(stderr, "error: %s\n", rt_last_error());
            destroy_all();
            return 1;
        }
    }
    if (rt_multi_init(scenes.data(), n_gpus) != RT_OK) {   // NCCL communicators are part of the set-up, like the BVH build (main.rs:45-47)
        std::fprintf(stderr, "error: %s\n", rt_last_error());
        destroy_all();
        return 1;
    }
    RtScene* scene = scenes[0];
    RtSceneInfo info;
    rt_scene_info(scene, &info);
    RtSceneDesc desc;
    rt_scene_get_desc(scene, &desc);                                                               // a text scene brings its own DIMENSIONS / SAMPLES
    width = desc.width; height = desc.height; samples = desc.samples;
    std::printf("Scene finite primitives: %d, light sources: %d\n", info.n_tris, info.n_lights);   // main.rs:49-53

    std::vector<uint8_t> rendered((size_t)width * (size_t)height * 3);
    RtRenderParams params = {};
    params.seed = (uint64_t)env_ll("RT_SEED", 0);
    params.collect_stats = (int)env_ll("RT_STATS", 0);
    RtStats stats;
    const auto start = std::chrono::steady_clock::now();                                           // main.rs:54
    if (rt_render_multi(scenes.data(), n_gpus, &params, rendered.data(), &stats) != RT_OK) {
        std::fprintf(stderr, "error: %s\n", rt_last_error());
        destroy_all();
        return 1;
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    std::printf("Rendering took %.6fs\n", secs);                                                   // main.rs:57-58
    const double msamples = (double)width * height * samples / secs / 1e6;
    std::printf("{\"msamples_per_s\": %.3f, \"n_gpus\": %d, \"kernel_ms\": %.3f, \"total_ms\": %.3f, \"render_ms\": %.3f, \"reduce_ms\": %.3f, \"resolve_ms\": %.3f", msamples,
                n_gpus, stats.kernel_ms, stats.total_ms, stats.render_ms, stats.reduce_ms, stats.resolve_ms);
    if (params.collect_stats) std::printf(", \"segments\": %llu", (unsigned long long)stats.segments);   // work counters exist only with RT_STATS=1
    std::printf("}\n");
    std::printf("Dumping to %s\n", output_ppm);                                                    // main.rs:59
    if (rt_write_ppm(output_ppm, width, height, rendered.data(), (int)env_ll("RT_PPM_APPEND", 0)) != RT_OK) {
        std::fprintf(stderr, "error: %s\n", rt_last_error());
        destroy_all();
        return 1;
    }
    if (argc > 6) {                                                                                // main.rs:68-72
        const std::string png = std::string(argv[6]) + ".png";
        if (rt_write_png(png.c_str(), width, height, rendered.data()) != RT_OK) {
            std::fprintf(stderr, "error: %s\n", rt_last_error());
            destroy_all();
            return 1;
        }
        std::printf("Image dumped to %s\n", png.c_str());
    }
    destroy_all();
    return 0;
}
